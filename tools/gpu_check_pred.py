"""Time the predecessor pass alone on the C3 matrix."""
import ctypes as C, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from drs_abm_b200 import _lib, synth
from drs_abm_b200.routing import RoadGraph
lib = _lib.load()
data = synth.config_graph("C3")
G = RoadGraph(data, numbering="input"); G.build()
dist = C.c_void_p(); pred = C.c_void_p(); ld = C.c_int64()
_lib.check(lib.drs_apsp_matrices(G._handle, C.byref(dist), C.byref(pred), C.byref(ld)))
out = torch.empty((ld.value, ld.value), dtype=torch.int32, device="cuda")
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.time()
    _lib.check(lib.drs_pred_build(G._handle, 0, dist, ld.value, 0, ld.value, C.c_void_p(out.data_ptr()), None))
    torch.cuda.synchronize(); dt = time.time() - t0
print("pred pass: %.2f ms  (lib %s) checksum %d" % (dt * 1e3, os.environ.get("DRS_B200_LIB", "default"), int(out.long().sum().item())))
