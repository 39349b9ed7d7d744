"""SSSP-based APSP vs FW: equality and timing; delta sweep."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drs_abm_b200 import _lib, synth
from drs_abm_b200.routing import RoadGraph

lib = _lib.load()
for name in sys.argv[1:] or ["C2", "C3"]:
    data = synth.config_graph(name)
    Gf = RoadGraph(data, algorithm=_lib.APSP_FW, numbering="input")
    t0 = time.time(); Gf.build(); t_fw = time.time() - t0
    Gs = RoadGraph(data, algorithm=_lib.APSP_SSSP, numbering="input")
    t0 = time.time(); Gs.build(); t_s1 = time.time() - t0
    Gs._valid = False
    t0 = time.time(); Gs.build(); t_s2 = time.time() - t0
    prof = np.zeros(4); r = C.c_int32()
    lib.drs_apsp_last_profile(Gs._handle, _lib.ptr_f64(prof), C.byref(r))
    idx = np.random.RandomState(0).randint(0, data.n, 64)
    same = all(np.array_equal(Gf.dist_rows(int(i), 1), Gs.dist_rows(int(i), 1)) for i in idx)
    o, d = synth.random_od_pairs(data.n, 2000, 1)
    pf, ps = Gf.path_sums(o, d), Gs.path_sums(o, d)
    print(name, "fw %.3fs sssp %.3fs (first %.3fs) prof_ms %s rows_equal %s paths_equal %s" % (
        t_fw, t_s2, t_s1, prof.round(2).tolist(), same, all(np.array_equal(a, b) for a, b in zip(pf, ps))), flush=True)
    # delta sweep on a shard of rows
    npad = (data.n + 127) // 128 * 128
    nrows = min(npad, 8192)
    buf = torch.empty((nrows, npad), dtype=torch.float32, device="cuda")
    mean_w = float(np.mean(data.eattrs["ttmedian"]))
    for mult in (0.5, 1, 2, 4, 8, 16):
        torch.cuda.synchronize(); t0 = time.time()
        _lib.check(lib.drs_apsp_rows_sssp(Gs._handle, 0, C.c_void_p(buf.data_ptr()), npad, 0, nrows, mult * mean_w, None))
        torch.cuda.synchronize()
        print("   delta = %.1f x mean weight: %d rows in %.4f s -> full APSP %.3f s" % (mult, nrows, time.time() - t0, (time.time() - t0) * data.n / nrows), flush=True)
    Gf.close(); Gs.close()
