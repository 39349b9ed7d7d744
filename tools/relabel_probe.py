"""How much does the vertex numbering of the input map matter to the SSSP build?  Builds the C3 matrix for several
relabelings of the same graph (row-major as generated, random, Morton, BFS, 4x2 tiles) and prints the kernel time.
Run on a GPU box: python tools/relabel_probe.py"""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from drs_abm_b200 import _lib, synth

data = synth.config_graph("C3")
n = data.n
lng = np.unique(np.asarray(data.vattrs["lng"], dtype=np.float64), return_inverse=True)[1].astype(np.int64)   # grid column
lat = np.unique(np.asarray(data.vattrs["lat"], dtype=np.float64), return_inverse=True)[1].astype(np.int64)   # grid row
lib = _lib.load()


def morton(x, y):
    def spread(v):
        v = v & 0xffff
        v = (v | (v << 8)) & 0x00ff00ff
        v = (v | (v << 4)) & 0x0f0f0f0f
        v = (v | (v << 2)) & 0x33333333
        v = (v | (v << 1)) & 0x55555555
        return v
    return spread(x) | (spread(y) << 1)


def hilbert(x, y, bits=8):
    """Hilbert-curve index of integer grid points (vectorised xy2d)."""
    x = x.astype(np.int64).copy(); y = y.astype(np.int64).copy()
    d = np.zeros_like(x)
    s = 1 << (bits - 1)
    while s > 0:
        rx = ((x & s) > 0).astype(np.int64)
        ry = ((y & s) > 0).astype(np.int64)
        d += s * s * ((3 * rx) ^ ry)
        # rotate
        flip = (ry == 0) & (rx == 1)
        x = np.where(flip, s - 1 - x, x); y = np.where(flip, s - 1 - y, y)
        swap = ry == 0
        x, y = np.where(swap, y, x), np.where(swap, x, y)
        s >>= 1
    return d


def bfs_order():
    import scipy.sparse as sp
    from scipy.sparse.csgraph import breadth_first_order
    A = sp.csr_matrix((np.ones(2 * data.m), (np.concatenate([data.src, data.dst]), np.concatenate([data.dst, data.src]))), shape=(n, n))
    order, _ = breadth_first_order(A, 0, directed=False)
    rest = np.setdiff1d(np.arange(n), order)
    return np.concatenate([order, rest])


def rcm_order():
    import scipy.sparse as sp
    from scipy.sparse.csgraph import reverse_cuthill_mckee
    A = sp.csr_matrix((np.ones(2 * data.m), (np.concatenate([data.src, data.dst]), np.concatenate([data.dst, data.src]))), shape=(n, n))
    return np.asarray(reverse_cuthill_mckee(A, symmetric_mode=True))


orders = {
    "as generated (row-major)": np.arange(n),
    "random": np.random.RandomState(0).permutation(n),
    "morton": np.argsort(morton(lng, lat), kind="stable"),
    "hilbert": np.argsort(hilbert(lng, lat), kind="stable"),
    "morton of 2x-coarse cells, row-major inside": np.argsort(morton(lng // 2, lat // 2) * 4 + (lat % 2) * 2 + (lng % 2), kind="stable"),
    "tiles 4x2": np.argsort((lat // 2) * 100000 * 8 + (lng // 4) * 8 + (lat % 2) * 4 + (lng % 4), kind="stable"),
    "tiles 8x4": np.argsort((lat // 4) * 100000 * 32 + (lng // 8) * 32 + (lat % 4) * 8 + (lng % 8), kind="stable"),
    "bfs": bfs_order(),
    "rcm": rcm_order(),
}
for name, order in orders.items():
    new_id = np.empty(n, dtype=np.int64)
    new_id[order] = np.arange(n)                   # old vertex order[k] becomes vertex k
    src = _lib.as_i32(new_id[data.src]); dst = _lib.as_i32(new_id[data.dst])
    tt, ln = _lib.as_f64(data.eattrs["ttmedian"]), _lib.as_f64(data.eattrs["length"])
    h = C.c_void_p()
    _lib.check(lib.drs_graph_create(n, data.m, 0, _lib.ptr_i32(src), _lib.ptr_i32(dst), _lib.ptr_f64(tt), _lib.ptr_f64(ln), C.byref(h)), lib)
    _lib.check(lib.drs_apsp_set_algorithm(h, _lib.APSP_SSSP), lib)
    ms = np.zeros(4); rounds = C.c_int32(0); best = 1e9
    for rep in range(3):
        _lib.check(lib.drs_apsp_build(h, _lib.MODE_F32), lib)
        _lib.check(lib.drs_apsp_last_profile(h, _lib.ptr_f64(ms), C.byref(rounds)), lib)
        best = min(best, float(ms[2]))
    lib.drs_graph_destroy(h)
    print(json.dumps({"numbering": name, "sssp_ms": round(best, 2)}), flush=True)
