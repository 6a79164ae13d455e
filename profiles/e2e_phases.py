"""Phases of the end-to-end job at config 4 from pinned host CSC arrays (int32 rows here), host timers with
synchronisation between phases; run on the B200 box through gpurun."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import picturedrocks_b200 as pr
from picturedrocks_b200 import _native as nat
from picturedrocks_b200.engine import Engine
from picturedrocks_b200.synth import DiscretisedSpec
spec = DiscretisedSpec(N=1306127, P=27998, C=60, K=5, mean_density=0.07, seed=4)
eng, yd = spec.device_engine()
nnz = eng.nnz
rows = torch.empty(nnz, dtype=torch.int32, device="cuda"); codes = torch.empty(nnz, dtype=torch.uint8, device="cuda")
nat.call("prb_unpack_csc", nat.ptr(eng.packed), nnz, nat.ptr(rows), nat.ptr(codes), nat.stream())
h_rows = torch.empty(nnz, dtype=torch.int32, pin_memory=True); h_codes = torch.empty(nnz, dtype=torch.uint8, pin_memory=True)
h_indptr = torch.empty(eng.P + 1, dtype=torch.int64, pin_memory=True); h_y = torch.empty(spec.N, dtype=torch.int32, pin_memory=True)
h_rows.copy_(rows); h_codes.copy_(codes); h_indptr.copy_(eng.indptr); h_y.copy_(yd)
del rows, codes, eng; torch.cuda.empty_cache(); torch.cuda.synchronize()
def lap(name, t=[time.perf_counter()]):
    torch.cuda.synchronize(); now = time.perf_counter(); print("%-28s %.1f ms" % (name, 1e3 * (now - t[0]))); t[0] = time.perf_counter()
lap("start")
eng2 = Engine.from_csc_arrays(h_indptr, h_rows, h_codes, spec.N)
lap("H2D + pack + Engine()")
inf2 = pr.markers.SparseInformationSet._from_engine(eng2, h_y.to("cuda", non_blocking=True))
lap("set_y (sort, V layout, cnt)")
fs2 = pr.markers.CIFE(inf2)
lap("CIFE init")
fs2.autoselect(50); S = fs2.S
lap("50 steps + S")
# ---- finer: time every native call and torch op group inside a fresh set_y
import collections
del fs2, inf2; eng2.v = None; torch.cuda.synchronize()
acc = collections.OrderedDict()
orig = nat.call
def timed(name, *a):
    torch.cuda.synchronize(); t0 = time.perf_counter(); orig(name, *a); torch.cuda.synchronize()
    acc[name] = acc.get(name, 0.0) + 1e3 * (time.perf_counter() - t0)
nat.call = timed
import picturedrocks_b200.engine as E
E.nat.call = timed
t0 = time.perf_counter(); eng2.set_y(h_y.to("cuda")); torch.cuda.synchronize(); tot = 1e3 * (time.perf_counter() - t0)
print("set_y again: %.1f ms; native calls:" % tot, {k: round(v, 1) for k, v in acc.items()})
