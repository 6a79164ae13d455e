#!/usr/bin/env python
"""bench.py — greedy MI marker selection on synthetic discretised expression matrices.

    python bench.py --gpus N --steps K --warmup W [--workload c4] [--selector CIFE]
    python bench.py --impl reference ...        # CPU arm: the oracle port on host cores

A "step" is one greedy selection step (BASELINE.json metric: cell·gene MI evals/s per
greedy step): argmax -> commit -> per-cell state of the winner -> one streaming pass
over the whole matrix (k_stream_v: joint histograms for H(x_j,x_i) and H(y,x_j,x_i)) ->
entropies -> score update.  value = N_cells * P_genes * K / t over all GPUs (genes are sharded; total
work is fixed, so scaling is "strong").  Inputs are generated on the device and are
resident in HBM when the timed region starts; `e2e` repeats the job from pinned HOST
buffers (upload + table build + init + K steps + read-back of S inside the timed region).
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
    # BASELINE.json configs[4]: two points of the sweep (JMI); k = 10 / 20 go through the
    # generic shared-histogram kernel, k = 5 through the V layout
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
METRIC = "cell_gene_mi_evals_per_s_per_greedy_step"
UNIT = "cell*gene/s"


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
        """Reference arm only: codes of the listed genes from the CPU oracle's discretiser."""
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


def known_traffic(workload):
    """dram bytes per launch of the dominant kernel from a committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            d = json.load(f)
        return d.get(workload, {}).get("k_stream_v_dram_bytes_per_launch")
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
def run_reference(args, w):
    """CPU arm: the oracle port (oracle/oracle.c, a restatement of the reference's
    _sparse_entropy_wrt) on all host cores, on a bounded gene sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O

    spec = make_spec(w)
    cores = O.max_threads()
    G = args.ref_genes or int(min(512, max(64, 8 * cores)))
    y = spec.cpu_labels()
    # a dense-ish first gene so the master column is representative
    genes = list(range(G))
    X = spec.cpu_csc(genes, y)
    X.data = X.data.astype(np.int64)
    inf = O.SparseInformationSet(X, y, nthreads=cores)
    dens = np.diff(X.indptr)
    j = int(np.argsort(dens)[len(dens) // 2])  # median-density gene as x_j

    def step():
        a = inf.entropy_wrt(np.array([j], dtype=np.int64))
        b = inf.entropy_wrt(np.array([-1, j], dtype=np.int64))
        return a, b

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = spec.N * G * args.steps / dt
    sample = "%d of %d genes x all %d cells, both entropy_wrt calls of one step, %d OpenMP threads" % (
        G, spec.P, spec.N, cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int64/f64",
        "data": "synthetic",
        "config": {"workload": w["name"], "selector": args.selector, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def cpu_baseline_sample(spec, eng, y_host, budget_s=15.0):
    """Oracle port on a bounded gene sample taken from the resident matrix (rank 0, N=1)."""
    import torch

    from oracle import oracle as O
    from picturedrocks_b200 import _native as nat

    cores = O.max_threads()
    G = int(min(eng.P, max(64, 8 * cores)))
    X = eng.to_scipy_csc(G)  # original cell order, int64 data like the reference's
    ip = X.indptr
    inf = O.SparseInformationSet(X, y_host, nthreads=cores)
    dens = np.diff(ip)
    j = int(np.argsort(dens)[len(dens) // 2])
    reps, t_total = 0, 0.0
    while reps < 3 or (t_total < 2.0 and reps < 20):
        t0 = time.perf_counter()
        inf.entropy_wrt(np.array([j], dtype=np.int64))
        inf.entropy_wrt(np.array([-1, j], dtype=np.int64))
        t_total += time.perf_counter() - t0
        reps += 1
        if t_total > budget_s:
            break
    return {"value": spec.N * G * reps / t_total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d of %d genes x all %d cells, both entropy_wrt calls of one greedy step, "
                      "%d reps, %d OpenMP threads" % (G, spec.P, spec.N, reps, cores)}


def run_ours(args, w):
    import torch
    import torch.distributed as dist

    import picturedrocks_b200 as pr
    from picturedrocks_b200 import _native as nat
    from picturedrocks_b200.dist import Comm, partition_genes
    from picturedrocks_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        # NCCL's own log lines (e.g. NCCL_DEBUG=VERSION) go to stderr: stdout carries the one
        # JSON line only
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    spec = make_spec(w)
    parts = partition_genes(spec.P, world)
    lo, hi = parts[rank]
    if world > 1:
        comm = Comm(shard_sizes=[b - a for a, b in parts])
    Selector = getattr(pr.markers, args.selector)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident inputs ------------------------------------------------------------------
    eng, yd = spec.device_engine(lo, hi, comm=comm)
    inf = pr.markers.SparseInformationSet._from_engine(eng, yd)
    nnz_local = eng.nnz
    stream_bytes = 128 * eng.v["rows"] if eng.v is not None else 4 * nnz_local
    hot_kernel = "k_stream_v" if eng.vmode else "k_stream"
    fs = Selector(inf)
    if args.skip_stream:
        # overhead probe: every kernel of the step except the streaming pass (results are
        # meaningless; the JSON line is marked invalid)
        eng._stream_v_launch = lambda *a, **k: None
        eng._stream_launch = lambda *a, **k: None
    os.environ.setdefault("PRB_GRAPH_MIN_STEPS", "3")
    import picturedrocks_b200.markers.mutualinformation.iterative as _it
    _it.GRAPH_MIN_STEPS = min(_it.GRAPH_MIN_STEPS, max(args.warmup, 2))
    fs.autoselect(args.warmup)  # W untimed steps (the step's CUDA graph is captured here)
    barrier()
    # ---- timed region: K greedy steps, inputs resident in HBM ------------------------------
    sampler = ClockSampler(local)
    nat.LAUNCHES = 0
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    fs.autoselect(args.steps)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    graphed = fs._graph is not None
    launches = fs.graph_kernels_per_step * args.steps if graphed else nat.LAUNCHES
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # ---- the dominant kernel, timed live with CUDA events on the launching stream: the same
    # greedy steps continue, launched eagerly (events cannot be read back from inside a
    # replayed CUDA graph) -----------------------------------------------------------------
    n_inst = min(args.steps, 10)
    eng.profile_events = []
    barrier()
    iv0, iv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iv0.record()
    fs.autoselect(n_inst)
    iv1.record()
    barrier()
    events = eng.profile_events
    eng.profile_events = None
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in events])) if events else None
    inst_ms = iv0.elapsed_time(iv1) / n_inst
    S_dev = fs.S
    value = spec.N * spec.P * args.steps / (ms * 1e-3)

    # ---- e2e: same job from pinned HOST buffers through the public API ---------------------
    e2e = None
    if not args.no_e2e:
        rows = torch.empty(max(nnz_local, 1), dtype=torch.int32, device="cuda")
        codes = torch.empty(max(nnz_local, 1), dtype=torch.uint8, device="cuda")
        nat.call("prb_unpack_csc", nat.ptr(eng.packed), nnz_local, nat.ptr(rows), nat.ptr(codes),
                 nat.stream())
        h_rows = torch.empty(max(nnz_local, 1), dtype=torch.int32, pin_memory=True)
        h_codes = torch.empty(max(nnz_local, 1), dtype=torch.uint8, pin_memory=True)
        h_indptr = torch.empty(eng.P + 1, dtype=torch.int64, pin_memory=True)
        h_y = torch.empty(spec.N, dtype=torch.int32, pin_memory=True)
        h_rows.copy_(rows)
        h_codes.copy_(codes)
        h_indptr.copy_(eng.indptr)
        h_y.copy_(yd)
        # the resident run's buffers go back to torch's caching allocator (a warm allocator,
        # as in any long-running process); nothing of its contents is reused
        del rows, codes, fs, inf, eng
        barrier()
        t0 = time.perf_counter()
        eng2 = Engine.from_csc_arrays(h_indptr, h_rows[:nnz_local], h_codes[:nnz_local], spec.N,
                                      gene_lo=lo, P_total=spec.P, comm=comm)
        inf2 = pr.markers.SparseInformationSet._from_engine(eng2, h_y.to("cuda", non_blocking=True))
        fs2 = Selector(inf2)
        fs2.autoselect(args.steps)
        S_e2e = fs2.S  # device -> host read of the result
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        h2d = 5 * nnz_local + 8 * (hi - lo + 1) + 4 * spec.N
        e2e = {"value": spec.N * spec.P * args.steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d // args.steps), "d2h_bytes_per_step": 4,
               "seconds_total": dt,
               "what": "pinned host CSC -> H2D -> pack, V layout -> %s init -> %d steps "
                       "-> S to host"
                       % (args.selector, args.steps),
               "same_sequence_as_resident_run": S_e2e[: args.steps] == S_dev[args.warmup:][: args.steps]
               if args.warmup == 0 else None}
        eng_for_cpu, y_for_cpu = eng2, h_y.numpy().astype(np.int64)
    else:
        eng_for_cpu, y_for_cpu = eng, yd.cpu().numpy().astype(np.int64)

    # ---- roofline of the dominant kernel (k_stream_v) ---------------------------------------
    peak, peak_src = measured_peak()
    roofline = None
    if k_ms:
        alg_bytes = 4.0 * nnz_local  # one packed uint32 word per stored entry, read once
        achieved = alg_bytes / (k_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": hot_kernel, "achieved": achieved, "peak": peak,
                    "unit": "GB/s", "frac": achieved / peak, "traffic": known_traffic(args.workload),
                    "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": k_ms,
                    "launches_timed": len(events), "peak_source": peak_src,
                    "share_of_step": k_ms / (ms / args.steps),
                    "timed_how": "CUDA events around each " + hot_kernel + " launch in %d further greedy "
                                 "steps launched eagerly right after the timed region "
                                 "(%.3f ms/step eager vs %.3f ms/step in the CUDA-graph timed "
                                 "region)" % (n_inst, inst_ms, ms / args.steps),
                    "stream_bytes_per_launch_incl_row_padding": stream_bytes,
                    "reference_layout_equiv_GBps": 5.0 * nnz_local / (k_ms * 1e-3) / 1e9}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_sample(spec, eng_for_cpu, y_for_cpu)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8 codes / int32 counts / f64 entropies",
            "data": "synthetic",
            "config": {"workload": w["name"], "selector": args.selector, "bins": w["K"],
                       "nnz": int(nnz_local) if world == 1 else None,
                       "parallelism": "genes sharded over %d GPU(s)" % world,
                       "l2": "inputs exceed L2 (%.1f GB packed CSC per GPU vs 126 MB)"
                             % (4.0 * nnz_local / 1e9),
                       "time_to_top_k_s": ms * 1e-3, "k": args.steps,
                       "time_to_top_k_note": "the K timed greedy steps on resident data; the "
                                             "selector's init pass (base scores from the layout's "
                                             "count tables) is outside, e2e includes it"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "cuda_graph": graphed, "roofline": roofline,
            "cpu_baseline": cpu, "selected": S_dev[:8],
        }
        if args.skip_stream:
            line["invalid"] = "overhead probe: the streaming kernel was not launched"

        print(json.dumps(line), flush=True)
    if world > 1:
        from picturedrocks_b200.dist import shutdown

        shutdown()


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
    ap.add_argument("--ref-genes", type=int, default=0)
    ap.add_argument("--skip-stream", action="store_true",
                    help="tuning aid: time the step without its streaming kernel (invalid result)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, w)
    else:
        run_ours(args, w)


if __name__ == "__main__":
    main()
