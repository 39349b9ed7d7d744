"""First GPU check: ALU microbenchmark, FW correctness vs scipy Dijkstra, FW timing."""
import ctypes as C
import json
import os
import subprocess
import sys
import time

import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import dijkstra

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drs_abm_b200 import _lib, synth

lib = _lib.load()
out = {}


def make_graph(data):
    h = C.c_void_p()
    src, dst = _lib.as_i32(data.src), _lib.as_i32(data.dst)
    tt, ln = _lib.as_f64(data.eattrs["ttmedian"]), _lib.as_f64(data.eattrs["length"])
    _lib.check(lib.drs_graph_create(data.n, data.m, 0, _lib.ptr_i32(src), _lib.ptr_i32(dst), _lib.ptr_f64(tt),
                                    _lib.ptr_f64(ln), C.byref(h)))
    return h


def scipy_apsp(data):
    w = np.asarray(data.eattrs["ttmedian"], dtype=np.float64)
    A = csr_matrix((np.concatenate([w, w]), (np.concatenate([data.src, data.dst]), np.concatenate([data.dst, data.src]))),
                   shape=(data.n, data.n))
    return dijkstra(A, directed=True)


def rows(h, n):
    buf = np.empty((n, n), dtype=np.float64)
    _lib.check(lib.drs_apsp_rows_to_host(h, 0, n, _lib.ptr_f64(buf)))
    return buf


for v in range(3):
    r = C.c_double()
    _lib.check(lib.drs_alu_peak(v, 20000, C.byref(r)))
    out["alu_peak_variant_%d_relax_per_s" % v] = r.value
print(json.dumps(out), flush=True)

for name in ["C1", "C1_uniform", "C2", "C2_float"]:
    data = synth.config_graph(name)
    ref = scipy_apsp(data)
    for mode in ([0, 1] if "float" not in name else [0]):
        h = make_graph(data)
        t0 = time.time()
        _lib.check(lib.drs_apsp_build(h, mode))
        t1 = time.time()
        _lib.check(lib.drs_apsp_build(h, mode))
        t2 = time.time()
        got = rows(h, data.n)
        if "float" in name:
            err = np.max(np.abs(got - ref) / np.maximum(ref, 1e-9))
            ok = bool(err < 1e-5)
            out["%s_mode%d" % (name, mode)] = {"ok": ok, "max_rel_err": float(err), "build_s": t2 - t1, "first_build_s": t1 - t0}
        else:
            ok = bool(np.array_equal(got, ref))
            out["%s_mode%d" % (name, mode)] = {"ok": ok, "mismatch": int((got != ref).sum()), "build_s": t2 - t1, "first_build_s": t1 - t0}
        # path check on a few pairs
        o, d = synth.random_od_pairs(data.n, 2000, 1)
        tt = np.empty(2000); ln = np.empty(2000); hops = np.empty(2000, dtype=np.int32)
        _lib.check(lib.drs_query_pairs(h, 2000, _lib.ptr_i32(o), _lib.ptr_i32(d), _lib.ptr_f64(tt), _lib.ptr_f64(ln),
                                       hops.ctypes.data_as(C.POINTER(C.c_int32))))
        perr = np.max(np.abs(tt - ref[o, d]) / np.maximum(ref[o, d], 1e-9))
        out["%s_mode%d" % (name, mode)]["path_tt_max_rel_err"] = float(perr)
        out["%s_mode%d" % (name, mode)]["mean_hops"] = float(hops.mean())
        lib.drs_graph_destroy(h)
        print(name, mode, out["%s_mode%d" % (name, mode)], flush=True)

if "--c3" in sys.argv:
    data = synth.config_graph("C3")
    for mode in [0, 1]:
        h = make_graph(data)
        t0 = time.time()
        _lib.check(lib.drs_apsp_build(h, mode))
        t1 = time.time()
        npad = (data.n + 127) // 128 * 128
        out["C3_mode%d" % mode] = {"build_s": t1 - t0, "n": data.n, "relax_per_s": npad ** 3 / (t1 - t0)}
        # spot check 20 rows against scipy
        w = np.asarray(data.eattrs["ttmedian"], dtype=np.float64)
        A = csr_matrix((np.concatenate([w, w]), (np.concatenate([data.src, data.dst]), np.concatenate([data.dst, data.src]))),
                       shape=(data.n, data.n))
        idx = np.random.RandomState(0).randint(0, data.n, 20)
        ref = dijkstra(A, directed=True, indices=idx)
        buf = np.empty((1, data.n))
        bad = 0
        for k, i in enumerate(idx):
            _lib.check(lib.drs_apsp_rows_to_host(h, int(i), 1, _lib.ptr_f64(buf)))
            bad += int((buf[0] != ref[k]).sum())
        out["C3_mode%d" % mode]["spot_mismatch"] = bad
        lib.drs_graph_destroy(h)
        print("C3", mode, out["C3_mode%d" % mode], flush=True)

os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/gpu_check_fw.json", "w") as fh:
    json.dump(out, fh, indent=1)
print(json.dumps(out))
