"""Two-GPU checks of the sharded all-pairs builds (NCCL broadcast, fused P2P pivot step, sharded SSSP, sharded
separator-partitioned build).
Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from drs_abm_b200 import _lib, apsp_dist, synth
    from drs_abm_b200.routing import RoadGraph
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    data = synth.grid_city(40, 7, 0.93, "int")                 # ~1.5k vertices: 12 blocks, uneven split impossible -> 6 + 6
    G = RoadGraph(data, part_target=200, second_level_min=1)    # ~8 parts and a second level on this small map
    tiles = apsp_dist.CudaTiles(G.upload(), _lib.MODE_F32)
    npad = (data.n + 127) // 128 * 128
    results = {}
    for name in ("fw", "sssp", "sep"):
        local, pred, row0 = apsp_dist.build_sharded(tiles, npad, rank, world, dist, algorithm=name)
        full, fpred = apsp_dist.gather_replica(local, pred, npad, world, dist, tiles)
        results[name] = (full.cpu().numpy(), fpred.cpu().numpy())
    comm = apsp_dist.P2PComm(tiles, npad, rank, world, dist)
    for rep in range(2):                                        # twice: signal / ack values must stay monotone across builds
        local, pred, row0 = apsp_dist.build_sharded_p2p(tiles, npad, rank, world, dist, comm)
    full, fpred = apsp_dist.gather_replica(local, pred, npad, world, dist, tiles)
    results["fw_p2p"] = (full.cpu().numpy(), fpred.cpu().numpy())
    comm.close()
    # the replica serves queries through the ordinary API
    G.adopt(full, fpred)
    o, d = synth.random_od_pairs(data.n, 200, 2)
    tt, ln, hops = G.path_sums(o, d)
    np.savez(os.path.join(out_dir, "r%d.npz" % rank), tt=tt, **{k + "_d": v[0] for k, v in results.items()},
             **{k + "_p": v[1] for k, v in results.items()})
    dist.destroy_process_group()


def test_sharded_builds_agree_with_single_gpu(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, ROOT)
    from drs_abm_b200 import synth
    from drs_abm_b200.routing import RoadGraph
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    data = synth.grid_city(40, 7, 0.93, "int")
    G = RoadGraph(data, algorithm=0, part_target=200, second_level_min=1)   # the workers' numbering
    want = G.dist_rows(0, G.n)
    o, d = synth.random_od_pairs(data.n, 200, 2)
    want_tt = G.path_sums(o, d)[0]
    dev = G.dev_ids(np.arange(G.n))                              # the raw shards are in the library's vertex numbering
    for rank in range(2):
        z = np.load(os.path.join(str(tmp_path), "r%d.npz" % rank))
        for k in ("fw", "sssp", "sep", "fw_p2p"):
            assert np.array_equal(G.host_cols(z[k + "_d"][dev]).astype(np.float64), want), (rank, k)
            assert np.array_equal(z[k + "_p"], z["fw_p"]), (rank, k)       # canonical predecessors do not depend on the path taken
        assert np.array_equal(z["tt"], want_tt)
    G.close()
