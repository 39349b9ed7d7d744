"""The multi-GPU orchestration (ownership, broadcast order, look-ahead) on CPU: two gloo ranks,
tile arithmetic emulated in numpy (test double for the CUDA building blocks)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TILE = 128


class NumpyTiles(object):
    """Test double of drs_abm_b200.apsp_dist.CudaTiles (same method contract) on torch CPU tensors."""

    def __init__(self, n, src, dst, w):
        import torch
        self.torch = torch
        self.n, self.src, self.dst, self.w = n, src, dst, w

    def empty(self, rows, cols):
        return self.torch.empty((rows, cols), dtype=self.torch.float32)

    def empty_pred(self, rows, cols):
        return self.torch.empty((rows, cols), dtype=self.torch.int32)

    def init(self, local, row0):
        a = local.numpy()
        a[:] = np.inf
        for i in range(a.shape[0]):
            if row0 + i < a.shape[1]:
                a[i, row0 + i] = 0
        for u, v, w in zip(self.src, self.dst, self.w):
            for (x, y) in ((u, v), (v, u)):
                if row0 <= x < row0 + a.shape[0]:
                    a[x - row0, y] = min(a[x - row0, y], w)

    @staticmethod
    def _minplus(A, B):
        return np.min(A[:, :, None] + B[None, :, :], axis=1)

    def pivot_row(self, panel, kb):
        p = panel.numpy()
        d = p[:, kb * TILE:(kb + 1) * TILE]
        for k in range(TILE):
            d[:] = np.minimum(d, d[:, [k]] + d[[k], :])
        for j in range(p.shape[1] // TILE):
            if j != kb:
                blk = p[:, j * TILE:(j + 1) * TILE]
                blk[:] = np.minimum(blk, self._minplus(d, blk))

    def update(self, local, panel, kb, skip_local, blo, bhi):
        a, p = local.numpy(), panel.numpy()
        d = p[:, kb * TILE:(kb + 1) * TILE]
        for bi in range(blo, bhi):
            if bi == skip_local:
                continue
            rows = a[bi * TILE:(bi + 1) * TILE]
            col = rows[:, kb * TILE:(kb + 1) * TILE]
            col[:] = np.minimum(col, self._minplus(col, d))
            for j in range(a.shape[1] // TILE):
                if j != kb:
                    blk = rows[:, j * TILE:(j + 1) * TILE]
                    blk[:] = np.minimum(blk, self._minplus(col, p[:, j * TILE:(j + 1) * TILE]))

    def sep_rows(self, local, row0):
        """numpy restatement of the separator-partitioned build (csrc/sep_apsp.cu, stages S1-S4) on the partition the
        library's own host code produces (drs_partition_order): rows [row0, row0 + rows) of the matrix, caller's numbering."""
        from drs_abm_b200.routing import partition_order
        n = self.n
        order, pp = partition_order(n, self.src, self.dst, None, None, 40)
        new = np.empty(n, dtype=np.int64); new[order] = np.arange(n)
        W = np.full((n, n), np.inf); np.fill_diagonal(W, 0.0)
        for u, v, w in zip(new[self.src], new[self.dst], self.w):
            W[u, v] = min(W[u, v], w); W[v, u] = min(W[v, u], w)
        sep0, nparts = int(pp[-1]), len(pp) - 1
        fin = np.isfinite(W) & ~np.eye(n, dtype=bool)

        def closure(A):
            A = A.copy()
            for k in range(A.shape[0]):
                A = np.minimum(A, A[:, [k]] + A[[k], :])
            return A
        lists = []
        for j in range(nparts):
            inter = np.arange(pp[j], pp[j + 1])
            assert not fin[np.ix_(inter, np.r_[0:pp[j], pp[j + 1]:sep0])].any()      # no arc between two interiors
            lists.append(list(sep0 + np.flatnonzero(fin[inter, sep0:].any(axis=0) | fin[sep0:, inter].any(axis=1))))
        listed = set(x for l in lists for x in l)
        for s in range(sep0, n):
            if s not in listed:
                lists[0].append(s)                                                   # orphan separator: any closed part
        M = [closure(W[np.ix_(q, q)]) for q in ([list(range(pp[j], pp[j + 1])) + lists[j] for j in range(nparts)])]   # S1
        dS = W[sep0:, sep0:].copy()
        for j in range(nparts):                                                       # S1x: S_j x S_j blocks
            nj, idx = pp[j + 1] - pp[j], np.array(lists[j], dtype=np.int64) - sep0
            if len(idx):
                dS[np.ix_(idx, idx)] = np.minimum(dS[np.ix_(idx, idx)], M[j][nj:, nj:])
        dS = closure(dS)                                                              # S2
        D = np.full((n, n - sep0), np.inf)                                            # S3: every vertex to every separator
        D[sep0:] = dS
        for i in range(nparts):
            ni, idx = pp[i + 1] - pp[i], np.array(lists[i], dtype=np.int64) - sep0
            if len(idx):
                D[pp[i]:pp[i + 1]] = self._minplus(M[i][:ni, ni:], dS[idx])
        Cm = np.full((n, n), np.inf)                                                  # S4
        Cm[:, sep0:] = D
        for j in range(nparts):
            nj, idx = pp[j + 1] - pp[j], np.array(lists[j], dtype=np.int64) - sep0
            if len(idx):
                Cm[:, pp[j]:pp[j + 1]] = self._minplus(D[:, idx], M[j][nj:, :nj])
            Cm[pp[j]:pp[j + 1], pp[j]:pp[j + 1]] = np.minimum(Cm[pp[j]:pp[j + 1], pp[j]:pp[j + 1]], M[j][:nj, :nj])
        a = local.numpy()
        a[:] = np.inf
        for i in range(a.shape[0]):
            u = row0 + i
            if u < n:
                a[i, :n] = Cm[new[u], new]
            elif u < a.shape[1]:
                a[i, u] = 0.0

    def pred(self, local, row0, out):
        out.zero_()

    def synchronize(self):
        pass


def _worker(rank, world, port, lookahead, result_dir, algorithm="fw"):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from drs_abm_b200 import apsp_dist, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    data = synth.grid_city(18, 3, 0.95, "int")          # 300-odd nodes -> 3 blocks: uneven 2 + 1 split
    npad = (data.n + TILE - 1) // TILE * TILE
    tiles = NumpyTiles(data.n, data.src, data.dst, np.asarray(data.eattrs["ttmedian"], dtype=np.float32))
    local, pred, row0 = apsp_dist.build_sharded(tiles, npad, rank, world, dist, lookahead=lookahead, with_pred=False,
                                                algorithm=algorithm)
    full, _ = apsp_dist.gather_replica(local, None, npad, world, dist, tiles)
    np.save(os.path.join(result_dir, "full_%d.npy" % rank), full.numpy())
    np.save(os.path.join(result_dir, "row0_%d.npy" % rank), np.array([row0, local.shape[0]]))
    dist.destroy_process_group()


@pytest.mark.parametrize("lookahead", [False, True])
def test_sharded_fw_two_ranks_gloo(tmp_path, lookahead):
    import torch.multiprocessing as mp
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra
    sys.path.insert(0, ROOT)
    from drs_abm_b200 import apsp_dist, synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, lookahead, str(tmp_path)), nprocs=2, join=True)
    data = synth.grid_city(18, 3, 0.95, "int")
    w = np.asarray(data.eattrs["ttmedian"])
    A = csr_matrix((np.concatenate([w, w]), (np.concatenate([data.src, data.dst]), np.concatenate([data.dst, data.src]))),
                   shape=(data.n, data.n))
    want = dijkstra(A, directed=True)
    for rank in range(2):
        full = np.load(os.path.join(str(tmp_path), "full_%d.npy" % rank))
        assert np.array_equal(full[:data.n, :data.n].astype(np.float64), want)
    r0 = np.load(os.path.join(str(tmp_path), "row0_0.npy")); r1 = np.load(os.path.join(str(tmp_path), "row0_1.npy"))
    assert r0[0] == 0 and r1[0] == r0[1] and r0[1] + r1[1] == full.shape[0]


def test_sharded_separator_build_two_ranks_gloo(tmp_path):
    """algorithm="sep" of the sharded build: every rank computes its own row block, no collective; the tile arithmetic is the
    numpy restatement of the stages on the partition drs_partition_order returns (host code of the library)."""
    import torch.multiprocessing as mp
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import dijkstra
    sys.path.insert(0, ROOT)
    from drs_abm_b200 import synth
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, False, str(tmp_path), "sep"), nprocs=2, join=True)
    data = synth.grid_city(18, 3, 0.95, "int")
    w = np.asarray(data.eattrs["ttmedian"])
    A = csr_matrix((np.concatenate([w, w]), (np.concatenate([data.src, data.dst]), np.concatenate([data.dst, data.src]))),
                   shape=(data.n, data.n))
    want = dijkstra(A, directed=True)
    for rank in range(2):
        full = np.load(os.path.join(str(tmp_path), "full_%d.npy" % rank))
        assert np.array_equal(full[:data.n, :data.n].astype(np.float64), want)


def test_block_partition():
    sys.path.insert(0, ROOT)
    from drs_abm_b200.apsp_dist import block_partition, owner_of
    assert block_partition(392, 8) == [49 * i for i in range(9)]
    st = block_partition(10, 4)
    assert st == [0, 3, 6, 8, 10]
    assert [owner_of(k, st) for k in range(10)] == [0, 0, 0, 1, 1, 1, 2, 2, 3, 3]
    assert block_partition(1, 2) == [0, 1, 1]
    with pytest.raises(ValueError):
        owner_of(10, st)
