#!/usr/bin/env python
"""bench.py — greedy MI marker selection on synthetic discretised expression matrices.

    python bench.py --gpus N --steps K --warmup W [--workload c4] [--selector CIFE]
    python bench.py --impl reference ...        # CPU arm: the reference's numba path (1 thread)

A "step" is one greedy selection step (BASELINE.json metric: cell·gene MI evals/s per greedy
step): argmax of the candidate records -> commit -> column of the winner -> one streaming
pass over the whole matrix (joint histograms for H(x_j,x_i) and H(y,x_j,x_i)) -> entropies ->
score update.  value = N_cells * P_genes * K / t over all GPUs (genes are sharded; total work is
fixed, so scaling is "strong").  Inputs are generated on the device and are resident in HBM
when the timed region starts; `e2e` repeats the job from pinned HOST buffers through the public
constructor SparseInformationSet(X_scipy, y) (upload + layout + init + K steps + read-back of S
inside the timed region).  The selected sequence of every run is compared with the CPU
oracle's (tests/golden/large_*.npz); a mismatch makes the run invalid (exit code 1).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[3]: 1.3M mouse-brain-shaped sparse, 60 clusters (the target config)
    "c4": dict(N=1306127, P=27998, C=60, K=5, mean_density=0.07, seed=4,
               name="1.3M-mouse-brain-shaped synthetic sparse 1,306,127 x 27,998, 60 clusters, 7% nnz"),
    # BASELINE.json configs[2]: PBMC-68k-shaped sparse
    "c3": dict(N=68579, P=32738, C=10, K=5, mean_density=0.02, seed=3,
               name="PBMC-68k-shaped synthetic sparse 68,579 x 32,738, 10 clusters, 2% nnz"),
    # BASELINE.json configs[0] / [1]: small dense count matrices through the real discretiser
    "c1": dict(N=2730, P=3451, C=19, K=5, seed=1, dense=True,
               name="Paul15-shaped synthetic dense 2,730 x 3,451, 19 clusters, makeinfoset (k=5)"),
    "c2": dict(N=3005, P=19972, C=9, K=5, seed=2, dense=True,
               name="Zeisel-shaped synthetic dense 3,005 x 19,972, 9 clusters, makeinfoset (k=5)"),
    # BASELINE.json configs[4]: three points of the sweep (JMI), 5 / 10 / 20 bins
    "c5_k5": dict(N=1000000, P=20000, C=20, K=5, mean_density=0.05, seed=5,
                  name="sweep point 1e6 x 2e4, 20 clusters, 5 bins, 5% nnz"),
    "c5_k10": dict(N=100000, P=50000, C=20, K=10, mean_density=0.05, seed=5,
                   name="sweep point 1e5 x 5e4, 20 clusters, 10 bins, 5% nnz"),
    "c5_k20": dict(N=100000, P=50000, C=20, K=20, mean_density=0.05, seed=5,
                   name="sweep point 1e5 x 5e4, 20 clusters, 20 bins, 5% nnz"),
    # one eighth of the genes of c4: what each rank streams in the 8-GPU run (tuning aid)
    "c4_shard8": dict(N=1306127, P=3500, C=60, K=5, mean_density=0.07, seed=4,
                      name="c4 gene shard: 1,306,127 x 3,500, 60 clusters, 7% nnz"),
    # small case for quick functional runs
    "small": dict(N=20000, P=2000, C=12, K=5, mean_density=0.06, seed=7,
                  name="small synthetic sparse 20,000 x 2,000, 12 clusters"),
}
# the other BASELINE configurations, measured after the main line: (workload, selector, steps)
EXTRA = [("c1", "CIFE", 20), ("c2", "JMI", 50), ("c3", "CIFE", 50), ("c5_k5", "JMI", 100),
         ("c5_k10", "JMI", 100), ("c5_k20", "JMI", 100)]
METRIC = "cell_gene_mi_evals_per_s_per_greedy_step"
UNIT = "cell*gene/s"


def config_of(w, selector, steps):
    """The `config` object — identical in both arms, so that the driver can compare them.
    `l2`: how the timing rule "flush L2 or use inputs larger than L2" is met by the GPU arm —
    an estimate from the workload's shape (2 bytes per stored code, the fullest of 8 gene
    shards), the measured stream size is in details.l2."""
    cfg = {"workload": w["name"], "selector": selector, "bins": w["K"], "k": steps}
    if not w.get("dense"):
        per_gpu8 = 2.0 * w["N"] * w["P"] * w["mean_density"] / 8
        cfg["l2"] = ("inputs exceed L2 (no flush): >= %.2f GB streamed per step and GPU at 1..8 GPUs "
                     "vs 126 MB" % (per_gpu8 / 1e9)) if per_gpu8 > 2 * 126e6 else \
            "inputs may stay L2-resident between steps (the workload is that small)"
    else:
        cfg["l2"] = "inputs stay L2-resident between steps (the workload is that small)"
    return cfg


def workload_key(w):
    return next(k for k, v in WORKLOADS.items() if v is w)


def golden_sequence(workload, selector):
    """The CPU oracle's gene sequence for this workload (tests/golden/make_golden_large.py)."""
    try:
        pack = np.load(os.path.join(ROOT, "tests", "golden", "large_%s.npz" % workload))
        return [int(s) for s in pack[selector + "_S"]]
    except Exception:
        return None


def check_sequence(S, workload, selector):
    gold = golden_sequence(workload, selector)
    if gold is None:
        return {"against": None, "equal": None, "steps_compared": 0}
    n = min(len(S), len(gold))
    diff = next((i for i in range(n) if int(S[i]) != gold[i]), None)
    return {"against": "tests/golden/large_%s.npz (CPU oracle, reference's greedy loop)" % workload,
            "equal": diff is None, "steps_compared": n, "first_difference_at": diff}


class DenseWorkload:
    """Poisson-gamma counts (host) -> GPU quantile discretiser -> packed CSC engine."""

    def __init__(self, w):
        self.N, self.P, self.C, self.K, self.seed = w["N"], w["P"], w["C"], w["K"], w["seed"]

    def _host(self):
        from picturedrocks_b200.synth import poisson_gamma

        return poisson_gamma(self.N, self.P, self.C, self.seed)

    def cpu_labels(self):
        return self._host()[1]

    def cpu_csc(self, genes, y=None):
        """CPU arms only: codes of the listed genes from the CPU oracle's discretiser."""
        import scipy.sparse

        from oracle import oracle as O

        X = self._host()[0]
        return scipy.sparse.csc_matrix(O.quantile_discretize(X[:, list(genes)], self.K))

    def device_engine(self, lo, hi, comm=None):
        import torch

        from picturedrocks_b200.engine import Engine
        from picturedrocks_b200.markers.mutualinformation.infoset import discretize_to_device
        from picturedrocks_b200.synth import poisson_gamma

        X, y = poisson_gamma(self.N, self.P, self.C, self.seed)
        codes = discretize_to_device(X[:, lo:hi], self.K)
        eng = Engine.from_dense_codes(codes, self.N, gene_lo=lo, P_total=self.P, comm=comm)
        return eng, torch.from_numpy(y.astype(np.int32)).cuda()


def make_spec(w):
    from picturedrocks_b200.synth import DiscretisedSpec

    if w.get("dense"):
        return DenseWorkload(w)

    return DiscretisedSpec(N=w["N"], P=w["P"], C=w["C"], K=w["K"], mean_density=w["mean_density"],
                           seed=w["seed"])


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def known_traffic(workload, world, kernel):
    """dram bytes per launch of the dominant kernel from a committed ncu --set full capture of
    exactly this workload / shard size / kernel (profiles/roofline_traffic.json), else None."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            d = json.load(f)
        e = d.get("%s@%d" % (workload, world))
        return e["dram_bytes_per_launch"] if e and e.get("kernel") == kernel else None
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------
# CPU arms
# ------------------------------------------------------------------------------------------
def reference_sample(w, selector, n_genes, n_steps):
    """Bounded sample of a workload for the CPU arms: `n_genes` candidate genes evenly spaced
    over the matrix plus the genes the CPU oracle selects in the first `n_steps` steps (the
    x_j of the timed steps — real greedy steps, not a fixed median gene), all cells.
    Returns (X scipy CSC int64, y int64, list of x_j columns inside X, genes sampled)."""
    import scipy.sparse

    spec = make_spec(w)
    gold = golden_sequence(workload_key(w), selector) or []
    cand = sorted(set(int(g) for g in np.linspace(0, spec.P - 1, min(n_genes, spec.P)).astype(int)))
    xj = list(gold[:n_steps])
    in_cand = set(cand)
    genes = cand + [g for g in dict.fromkeys(xj) if g not in in_cand]
    if w.get("dense"):
        y = spec.cpu_labels()
        X = spec.cpu_csc(genes, y)
        X.data = X.data.astype(np.int64)
    else:
        from oracle import oracle as O

        indptr, rows, codes, y = O.synth_csc(spec, genes=genes)  # oracle/synth_cpu.c
        X = scipy.sparse.csc_matrix((codes.astype(np.int64), rows, indptr),
                                    shape=(spec.N, len(genes)))
    pos = {g: i for i, g in enumerate(genes)}
    if not xj:  # no golden sequence: the densest sampled genes stand in for the winners
        order = np.argsort(-np.diff(X.indptr))
        xj = [genes[int(i)] for i in order[: max(n_steps, 1)]]
    return X, y.astype(np.int64), [pos[g] for g in xj], len(genes)


def time_cpu_steps(infoset, xj_cols, warmup, steps, supervised=True):
    """The two entropy_wrt calls (plus the two scalar entropies) of one greedy step per x_j
    (iterative.py:50-56), timed on the host."""
    def step(j):
        infoset.entropy(np.array([j], dtype=np.int64))
        a = infoset.entropy_wrt(np.array([j], dtype=np.int64))
        if supervised:
            infoset.entropy(np.array([-1, j], dtype=np.int64))
            a = infoset.entropy_wrt(np.array([-1, j], dtype=np.int64))
        return a

    n = len(xj_cols)
    for i in range(warmup):
        step(xj_cols[i % n])
    t0 = time.perf_counter()
    for i in range(steps):
        step(xj_cols[(warmup + i) % n])
    return time.perf_counter() - t0


def cpu_factory():
    """(kind, constructor) of the CPU implementation that is timed: the reference's own files
    (numba JIT) when present — /root/reference, or the travelling copy oracle/_ref — else the
    oracle port on one thread."""
    from oracle import oracle as O
    from oracle import ref_shim

    if ref_shim.available():
        infoset_mod, _ = ref_shim.load("jit")
        return "reference", (lambda X, y: infoset_mod.SparseInformationSet(X, y))
    return "port", (lambda X, y: O.SparseInformationSet(X, y, nthreads=1))


def run_reference(args, w):
    """CPU arm.  `value` is the reference's OWN implementation (the three hot-path files of
    umangv/picturedrocks, numba JIT, single-threaded as the reference is) on a bounded gene
    sample of the same workload; the oracle port on all host cores is reported beside it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    cores = O.host_cores()  # (torchrun sets OMP_NUM_THREADS=1: the thread count is explicit)
    supervised = args.selector != "CIFEUnsup"
    n_steps = args.warmup + args.steps
    # calibration on a small sample (also the JIT warm-up), then a sample sized for the budget
    budget = float(os.environ.get("PRB_REF_BUDGET_S", "120"))
    X0, y0, xj0, g0 = reference_sample(w, args.selector, 24, 2)
    kind, make = cpu_factory()
    inf0 = make(X0, y0)
    time_cpu_steps(inf0, xj0, 1, 0, supervised)  # compile
    t_cal = time_cpu_steps(inf0, xj0, 0, 2, supervised) / 2.0
    G = args.ref_genes or int(max(16, min(1024, budget / max(t_cal / g0 * n_steps, 1e-9))))
    X, y, xj, n_g = reference_sample(w, args.selector, G, n_steps)
    dt = time_cpu_steps(make(X, y), xj, args.warmup, args.steps, supervised)
    value = w["N"] * n_g * args.steps / dt
    # the oracle port (same algorithm in C, per-gene loop over all host cores), same sample
    port = O.SparseInformationSet(X, y, nthreads=cores)
    dt_p = time_cpu_steps(port, xj, min(args.warmup, 2), args.steps, supervised)
    sample = ("%d of %d genes (evenly spaced + the winners of the timed steps) x all %d cells; per "
              "step the two entropy_wrt and two entropy calls of one %s step with the x_j the CPU "
              "oracle selects at that step" % (n_g, w["P"], w["N"], args.selector))
    what = ("umangv/picturedrocks infoset.py / iterative.py loaded unmodified (oracle/ref_shim.py), "
            "numba JIT, 1 thread — the reference is single-threaded" if kind == "reference"
            else "oracle port, 1 thread (reference files not present)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64/f64",
        "data": "synthetic", "config": config_of(w, args.selector, args.steps),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample,
                         "what": what, "host_cores": cores},
        "port_all_cores": {"value": w["N"] * n_g * args.steps / dt_p, "unit": UNIT, "cores": cores,
                           "kind": "port", "ms_per_step": 1e3 * dt_p / args.steps},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(w, selector, budget_s=20.0):
    """cpu_baseline of the main line (rank 0, N=1): the reference's numba path, 1 thread, on a
    bounded sample; the port on all cores beside it."""
    from oracle import oracle as O

    cores = O.host_cores()
    supervised = selector != "CIFEUnsup"
    X0, y0, xj0, g0 = reference_sample(w, selector, 16, 2)
    kind, make = cpu_factory()
    inf0 = make(X0, y0)
    time_cpu_steps(inf0, xj0, 1, 0, supervised)
    t_cal = time_cpu_steps(inf0, xj0, 0, 2, supervised) / 2.0
    reps = 3
    G = int(max(16, min(512, budget_s / max(t_cal / g0 * (reps + 1), 1e-9))))
    X, y, xj, n_g = reference_sample(w, selector, G, reps + 1)
    dt = time_cpu_steps(make(X, y), xj, 1, reps, supervised)
    port = O.SparseInformationSet(X, y, nthreads=cores)
    dt_p = time_cpu_steps(port, xj, 1, reps, supervised)
    return {"value": w["N"] * n_g * reps / dt, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": "%d of %d genes x all %d cells, %d greedy steps (x_j = the oracle's winners), "
                      "%s, 1 thread" % (n_g, w["P"], w["N"], reps,
                                        "numba JIT" if kind == "reference" else "C port"),
            "port_all_cores": {"value": w["N"] * n_g * reps / dt_p, "unit": UNIT, "cores": cores,
                               "kind": "port"}}


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
class Ctx:
    def __init__(self):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        if self.world > 1:
            # NCCL's own log lines go to stderr: stdout carries the one JSON line only
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def _reduce(self, v, op):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, v):
        return self._reduce(v, self.dist.ReduceOp.MAX)

    def sum_over_ranks(self, v):
        return self._reduce(v, self.dist.ReduceOp.SUM)


def measure_resident(ctx, wname, selector, steps, warmup, skip_stream=False, sampler=None):
    """W warm-up + K timed greedy steps on a matrix resident in HBM, then up to 10 further
    eager steps with CUDA events around every launch of the streaming kernel."""
    torch = ctx.torch
    import picturedrocks_b200 as pr
    import picturedrocks_b200.markers.mutualinformation.iterative as _it
    from picturedrocks_b200 import _native as nat
    from picturedrocks_b200.dist import Comm, partition_genes, partition_genes_weighted

    w = WORKLOADS[wname]
    spec = make_spec(w)
    # contiguous gene ranges of equal nnz (every rank counts the whole matrix once: 0.1 s);
    # the dense workloads (configs 1, 2: discretised on the device) are cut by gene count
    if ctx.world == 1:
        parts = [(0, spec.P)]
    elif hasattr(spec, "device_gene_nnz"):
        parts = partition_genes_weighted(spec.device_gene_nnz(), ctx.world)
    else:
        parts = partition_genes(spec.P, ctx.world)
    lo, hi = parts[ctx.rank]
    comm = Comm(shard_sizes=[b - a for a, b in parts]) if ctx.world > 1 else None
    eng, yd = spec.device_engine(lo, hi, comm=comm)
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    fs = getattr(pr.markers, selector)(inf)
    if skip_stream:
        eng._stream_v_launch = lambda *a, **k: None
        eng._stream_launch = lambda *a, **k: None
    _it.GRAPH_MIN_STEPS = min(_it.GRAPH_MIN_STEPS, max(warmup, 2))
    fs.autoselect(warmup)  # W untimed steps (the step's CUDA graph is captured here)
    ctx.barrier()
    nat.LAUNCHES = 0
    if sampler is not None:
        sampler.start()
    ctx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fs.autoselect(steps)
    ev1.record()
    ctx.barrier()
    clocks = sampler.stop() if sampler is not None else None
    graphed = fs._graph is not None
    launches = fs.graph_kernels_per_step * steps if graphed else nat.LAUNCHES
    ms = ctx.max_over_ranks(ev0.elapsed_time(ev1))
    # the dominant kernel, timed live with CUDA events on the launching stream: the same greedy
    # steps continue, launched eagerly (events cannot be read back from inside a replayed graph)
    n_inst = min(steps, 10)
    eng.profile_events = []
    ctx.barrier()
    iv0, iv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iv0.record()
    fs.autoselect(n_inst)
    iv1.record()
    ctx.barrier()
    events, eng.profile_events = eng.profile_events, None
    k_ms = None
    if events:
        k_ms = ctx.max_over_ranks(float(np.mean([a.elapsed_time(b) for a, b in events])))
    inst_ms = iv0.elapsed_time(iv1) / n_inst
    S = fs.S
    nnz_local = eng.nnz
    nnz_total = int(ctx.sum_over_ranks(float(nnz_local)))
    if eng.vmode and eng.v is not None:
        kernel = "k_stream_ell2" if eng.ell2 else ("k_stream_ell" if eng.ell else "k_stream_v")
        bytes_entry = 2.0 if (eng.ell or eng.v16) else 4.0
        stream_bytes = 128 * eng.v["rows"]
    else:
        kernel, bytes_entry, stream_bytes = "k_stream", 4.0, 4 * nnz_local
    return dict(w=w, spec=spec, ms=ms, steps=steps, warmup=warmup, S=S, graphed=graphed,
                launches=launches, k_ms=k_ms, inst_ms=inst_ms, n_events=len(events), kernel=kernel,
                bytes_entry=bytes_entry, nnz_local=nnz_local, nnz_total=nnz_total,
                stream_bytes=stream_bytes, clocks=clocks, lo=lo, hi=hi, comm=comm,
                value=spec.N * spec.P * steps / (ms * 1e-3), keep=(eng, yd, inf, fs))


def roofline_of(ctx, wname, m):
    """Roofline of the dominant (streaming) kernel of a measurement (slowest rank's launch
    time over this rank's shard; the shards are equal to within a gene)."""
    if not m["k_ms"]:
        return None
    peak, peak_src = measured_peak()
    alg = m["bytes_entry"] * m["nnz_local"]  # one entry per stored code, read once
    ach = alg / (m["k_ms"] * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": m["kernel"], "achieved": ach, "peak": peak, "unit": "GB/s",
            "frac": ach / peak, "traffic": known_traffic(wname, ctx.world, m["kernel"]),
            "algorithmic_bytes_per_launch": alg,
            "algorithmic_bytes_per_entry": m["bytes_entry"],
            "stream_bytes_per_launch_incl_padding": m["stream_bytes"],
            "avg_launch_ms": m["k_ms"], "launches_timed": m["n_events"], "peak_source": peak_src,
            "share_of_step": m["k_ms"] / (m["ms"] / m["steps"]),
            "timed_how": "CUDA events around each %s launch in %d further greedy steps launched "
                         "eagerly right after the timed region (%.3f ms/step eager vs %.3f ms/step "
                         "in the CUDA-graph timed region); max over ranks"
                         % (m["kernel"], m["n_events"], m["inst_ms"] or 0.0, m["ms"] / m["steps"]),
            "reference_layout_equiv_GBps": 5.0 * m["nnz_local"] / (m["k_ms"] * 1e-3) / 1e9}


def measure_e2e(ctx, selector, steps, m):
    """The same job from pinned HOST buffers through the public constructor."""
    torch = ctx.torch
    import scipy.sparse

    import picturedrocks_b200 as pr
    from picturedrocks_b200 import _native as nat

    eng, yd, inf, fs = m.pop("keep")
    spec, lo, hi, comm = m["spec"], m["lo"], m["hi"], m["comm"]
    nnz = m["nnz_local"]
    # a canonical scipy CSC over PINNED host arrays (scipy wants int64 indices beyond 2^31 entries)
    idt = torch.int64 if nnz >= (1 << 31) - 1 else torch.int32
    rows = torch.empty(max(nnz, 1), dtype=torch.int32, device="cuda")
    codes = torch.empty(max(nnz, 1), dtype=torch.uint8, device="cuda")
    nat.call("prb_unpack_csc", nat.ptr(eng.packed), nnz, nat.ptr(rows), nat.ptr(codes), nat.stream())
    h_rows = torch.empty(max(nnz, 1), dtype=idt, pin_memory=True)
    h_codes = torch.empty(max(nnz, 1), dtype=torch.uint8, pin_memory=True)
    h_indptr = torch.empty(hi - lo + 1, dtype=idt, pin_memory=True)
    step_c = 1 << 28
    for a in range(0, nnz, step_c):  # chunked: no second full-size device array of int64 row ids
        h_rows[a: a + step_c].copy_(rows[a: a + step_c])
    h_codes.copy_(codes)
    h_indptr.copy_(eng.indptr)
    y_host = yd.cpu().numpy().astype(np.int64)
    X = scipy.sparse.csc_matrix((h_codes.numpy()[:nnz], h_rows.numpy()[:nnz], h_indptr.numpy()),
                                shape=(spec.N, hi - lo), copy=False)
    pinned_in_place = (X.indices.ctypes.data == h_rows.data_ptr()
                       and X.data.ctypes.data == h_codes.data_ptr())
    # the resident run's buffers go back to torch's caching allocator (a warm allocator, as in
    # any long-running process); nothing of its contents is reused
    del rows, codes, fs, inf, eng
    ctx.barrier()
    t0 = time.perf_counter()
    inf2 = pr.markers.SparseInformationSet(X, y_host, gene_lo=lo, P_total=spec.P, comm=comm)
    fs2 = getattr(pr.markers, selector)(inf2)
    fs2.autoselect(steps)
    S_e2e = fs2.S  # device -> host read of the result
    torch.cuda.synchronize()
    dt = ctx.max_over_ranks(time.perf_counter() - t0)
    h2d = X.indices.nbytes + X.data.nbytes + X.indptr.nbytes + 4 * spec.N
    e2e = {"value": spec.N * spec.P * steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": int(h2d // steps), "d2h_bytes_per_step": 4,
           "seconds_total": dt, "h2d_bytes_total": int(h2d),
           "pcie_seconds_at_55GBps": h2d / 55e9, "host_arrays_pinned_in_place": pinned_in_place,
           "what": "scipy CSC over pinned host arrays (%s row ids, uint8 codes) -> "
                   "SparseInformationSet(X, y): chunked H2D overlapped with pack + canonical check "
                   "on device -> streamed layout -> %s init -> %d steps -> S to host"
                   % (str(idt).replace("torch.", ""), selector, steps),
           "same_sequence_as_resident_run": [int(s) for s in S_e2e[:steps]]
           == [int(s) for s in m["S"][:steps]]}
    del fs2, inf2
    return e2e


def run_ours(args, wname):
    ctx = Ctx()
    torch = ctx.torch
    w = WORKLOADS[wname]
    sampler = ClockSampler(ctx.local)
    m = measure_resident(ctx, wname, args.selector, args.steps, args.warmup, args.skip_stream,
                         sampler)
    roof = roofline_of(ctx, wname, m)
    seq = check_sequence(m["S"], wname, args.selector)
    e2e = None
    if not args.no_e2e:
        e2e = measure_e2e(ctx, args.selector, args.steps, m)
    m.pop("keep", None)
    torch.cuda.empty_cache()
    extra = []
    if not args.no_extra:
        for xn, xsel, xsteps in EXTRA:
            xw = WORKLOADS[xn]
            xm = measure_resident(ctx, xn, xsel, xsteps, 3)
            xr = roofline_of(ctx, xn, xm)
            xs = check_sequence(xm["S"], xn, xsel)
            xm.pop("keep", None)
            torch.cuda.empty_cache()
            extra.append({"config": config_of(xw, xsel, xsteps), "value": xm["value"], "unit": UNIT,
                          "ms_per_step": xm["ms"] / xsteps, "nnz": xm["nnz_total"],
                          "time_to_top_k_s": xm["ms"] * 1e-3,
                          "roofline": None if xr is None else {
                              k: xr[k] for k in ("kernel", "achieved", "peak", "frac", "avg_launch_ms",
                                                 "share_of_step", "algorithmic_bytes_per_launch")},
                          "sequence_check": xs})
    cpu = None
    if ctx.rank == 0 and ctx.world == 1 and not args.no_cpu:
        cpu = cpu_baseline_sample(w, args.selector)
    ok = (seq["equal"] is not False
          and all(x["sequence_check"]["equal"] is not False for x in extra)
          and (e2e is None or e2e["same_sequence_as_resident_run"]))
    if ctx.rank == 0:
        line = {
            "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": ctx.world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u8 codes / int32 counts / f64 entropies", "data": "synthetic",
            "config": config_of(w, args.selector, args.steps),
            "details": {"nnz": m["nnz_total"],
                        "parallelism": "genes sharded over %d GPU(s)%s" % (
                            ctx.world, ", contiguous ranges of equal nnz; per step one 16-byte "
                            "record per rank over NVLink peer memory" if ctx.world > 1 else ""),
                        "l2": "inputs exceed L2 (%.2f GB streamed per launch and GPU vs 126 MB)"
                              % (m["stream_bytes"] / 1e9),
                        "time_to_top_k_s": m["ms"] * 1e-3,
                        "time_to_top_k_note": "the K timed greedy steps on resident data; the "
                                              "selector's init pass (base scores from the layout's "
                                              "count tables) is outside, e2e includes it"},
            "clocks": m["clocks"], "e2e": e2e, "gpu_launches": m["launches"],
            "cuda_graph": m["graphed"], "roofline": roof, "cpu_baseline": cpu,
            "sequence_check": seq, "selected": [int(s) for s in m["S"][:8]],
            "extra_configs": extra,
        }
        if args.skip_stream:
            line["invalid"] = "overhead probe: the streaming kernel was not launched"
        elif not ok:
            line["invalid"] = "a selected gene sequence differs from the CPU oracle's"
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        from picturedrocks_b200.dist import shutdown

        shutdown()
    if not ok and not args.skip_stream:
        sys.exit(1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS))
    ap.add_argument("--selector", default="CIFE", choices=["CIFE", "JMI", "CIFEUnsup"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the other BASELINE configurations (extra_configs)")
    ap.add_argument("--ref-genes", type=int, default=0)
    ap.add_argument("--skip-stream", action="store_true",
                    help="tuning aid: time the step without its streaming kernel (invalid result)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload])
    else:
        run_ours(args, args.workload)


if __name__ == "__main__":
    main()
