"""Build the APSP matrix of one synthetic config (used under ncu / clock sampling)."""
import argparse
import ctypes as C
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drs_abm_b200 import _lib, synth

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C2")
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--repeat", type=int, default=1)
ap.add_argument("--algorithm", type=int, default=0, help="0 FW, 1 SSSP, 2 auto")
args = ap.parse_args()
lib = _lib.load()
data = synth.config_graph(args.config)
h = C.c_void_p()
src, dst = _lib.as_i32(data.src), _lib.as_i32(data.dst)
tt, ln = _lib.as_f64(data.eattrs["ttmedian"]), _lib.as_f64(data.eattrs["length"])
_lib.check(lib.drs_graph_create(data.n, data.m, 0, _lib.ptr_i32(src), _lib.ptr_i32(dst), _lib.ptr_f64(tt),
                                _lib.ptr_f64(ln), C.byref(h)))
_lib.check(lib.drs_apsp_set_algorithm(h, args.algorithm))
for r in range(args.repeat):
    t0 = time.time()
    _lib.check(lib.drs_apsp_build(h, args.mode))
    dt = time.time() - t0
    npad = (data.n + 127) // 128 * 128
    print("build %d: %.4f s  %.3e relax/s" % (r, dt, npad ** 3 / dt), flush=True)
lib.drs_graph_destroy(h)
