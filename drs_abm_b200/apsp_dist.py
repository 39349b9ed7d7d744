"""Row-block-sharded blocked Floyd-Warshall over 1/2/4/8 GPUs of one host
(SURVEY.md section 8e): one process per GPU, torch.distributed for the plumbing.

Rank r owns a contiguous range of 128-row blocks of the distance matrix.  Round
kb: the owner of block-row kb closes the diagonal tile and updates its pivot row
panel (phases 1 and 2-row, ``drs_fw_pivot_row``), broadcasts the 128 x npad panel,
and every rank then updates its own rows (phase 2-column + phase 3,
``drs_fw_update``).  Column panels never move.  With look-ahead the owner of
round kb+1 updates just that block-row first, runs its pivot-row step and starts
the next broadcast on the communication stream while everybody (itself
included) is still busy with the rest of round kb, so the broadcast latency is
off the critical path.

The reference has no distributed code at all ("Distributed
communication backend: none"); this module is new functionality behind the same
matrix the single-GPU build produces (bit-identical: min-plus is exact and the
per-tile operation order does not depend on the sharding).

The tile arithmetic is delegated to a backend object so that the orchestration
(ownership, broadcast order, look-ahead) can be exercised on CPU with the gloo
backend in tests; the product backend, ``CudaTiles``, calls the C-ABI.
"""
import ctypes as C

import numpy as np

from . import _lib

TILE = _lib.FW_TILE


def block_partition(nb, world):
    """Contiguous block-row ranges: rank r owns [starts[r], starts[r+1])."""
    base, extra = divmod(nb, world)
    starts = [0]
    for r in range(world):
        starts.append(starts[-1] + base + (1 if r < extra else 0))
    return starts


def owner_of(kb, starts):
    for r in range(len(starts) - 1):
        if starts[r] <= kb < starts[r + 1]:
            return r
    raise ValueError("block %d outside the partition" % kb)


class CudaTiles(object):
    """Product backend: the drs_fw_* building blocks of libdrs_b200.so on torch CUDA tensors."""

    def __init__(self, graph_handle, mode):
        import torch
        self.torch = torch
        self.lib = _lib.load()
        self.g = graph_handle
        self.mode = mode
        self.dtype = torch.float32 if mode == _lib.MODE_F32 else torch.int32

    def _stream(self):
        return C.c_void_p(self.torch.cuda.current_stream().cuda_stream)

    def empty(self, rows, cols):
        return self.torch.empty((rows, cols), dtype=self.dtype, device="cuda")

    def empty_pred(self, rows, cols):
        return self.torch.empty((rows, cols), dtype=self.torch.int32, device="cuda")

    def init(self, local, row0):
        _lib.check(self.lib.drs_fw_init(self.g, self.mode, C.c_void_p(local.data_ptr()), local.shape[1], row0,
                                        local.shape[0], self._stream()), self.lib)

    def pivot_row(self, panel, kb):
        _lib.check(self.lib.drs_fw_pivot_row(self.mode, C.c_void_p(panel.data_ptr()), panel.shape[1], kb,
                                             self._stream()), self.lib)

    def update(self, local, panel, kb, skip_local, blo, bhi):
        _lib.check(self.lib.drs_fw_update(self.mode, C.c_void_p(local.data_ptr()), local.shape[1], local.shape[0],
                                          C.c_void_p(panel.data_ptr()), kb, skip_local, blo, bhi, self._stream()),
                   self.lib)

    def sssp_rows(self, local, row0):
        _lib.check(self.lib.drs_apsp_rows_sssp(self.g, self.mode, C.c_void_p(local.data_ptr()), local.shape[1], row0,
                                               local.shape[0], 0.0, self._stream()), self.lib)

    def sep_rows(self, local, row0):
        _lib.check(self.lib.drs_apsp_rows_sep(self.g, self.mode, C.c_void_p(local.data_ptr()), local.shape[1], row0,
                                              local.shape[0], self._stream()), self.lib)

    def pred(self, local, row0, out):
        _lib.check(self.lib.drs_pred_build(self.g, self.mode, C.c_void_p(local.data_ptr()), local.shape[1], row0,
                                           local.shape[0], C.c_void_p(out.data_ptr()), self._stream()), self.lib)

    def synchronize(self):
        self.torch.cuda.synchronize()


class P2PComm(object):
    """Per-rank communication block for the fused pivot step (drs_fw_pivot_row_p2p): signal word,
    acknowledgement words (one per sender), a time-out word and two pivot-panel buffers, allocated with
    cudaMalloc, exported by CUDA IPC and mapped by every other rank of the host."""
    HEADER = 1024          # bytes: [0] signal, [64 + 4*r] ack of rank r, [512] time-out flag

    def __init__(self, tiles, npad, rank, world, dist, timeout_ms=20000):
        self.lib, self.rank, self.world, self.npad = tiles.lib, rank, world, npad
        self.tiles = tiles
        self.timeout_ms = timeout_ms
        self.panel_bytes = TILE * npad * 4
        total = self.HEADER + 2 * self.panel_bytes
        base = C.c_void_p()
        handle = (C.c_ubyte * 64)()
        _lib.check(self.lib.drs_p2p_alloc(total, C.byref(base), handle), self.lib)
        self.base = base.value
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle))
        self.peer_base = {}
        for r in range(world):
            if r == rank:
                continue
            ptr = C.c_void_p()
            hb = (C.c_ubyte * 64).from_buffer_copy(handles[r])
            _lib.check(self.lib.drs_p2p_open(hb, C.byref(ptr)), self.lib)
            self.peer_base[r] = ptr.value
        self.peers = sorted(self.peer_base)
        self.builds = 0
        dist.barrier()

    # addresses inside a block
    def signal(self, base):
        return base

    def ack(self, base, sender):
        return base + 64 + 4 * sender

    def panel_addr(self, base, which):
        return base + self.HEADER + which * self.panel_bytes

    def local_panel(self, which):
        """torch view of one of this rank's panel buffers."""
        return _as_tensor(self.tiles.torch, self.panel_addr(self.base, which), (TILE, self.npad), self.tiles.dtype)

    def _ptr_array(self, values):
        arr = (C.c_void_p * max(len(values), 1))(*values)
        return arr

    def pivot_row(self, panel, kb, which, value):
        stream = self.tiles._stream()
        panels = self._ptr_array([self.panel_addr(self.peer_base[r], which) for r in self.peers])
        sigs = self._ptr_array([self.signal(self.peer_base[r]) for r in self.peers])
        _lib.check(self.lib.drs_fw_pivot_row_p2p(self.tiles.mode, C.c_void_p(panel.data_ptr()), panel.shape[1], kb,
                                                 len(self.peers), panels, sigs, value, stream), self.lib)

    def wait_signal(self, value):
        words = self._ptr_array([self.signal(self.base)])
        _lib.check(self.lib.drs_fw_wait_words(1, words, value, self.timeout_ms, C.c_void_p(self.base + 512),
                                              self.tiles._stream()), self.lib)

    def wait_acks(self, value):
        words = self._ptr_array([self.ack(self.base, r) for r in self.peers])
        _lib.check(self.lib.drs_fw_wait_words(len(self.peers), words, value, self.timeout_ms, C.c_void_p(self.base + 512),
                                              self.tiles._stream()), self.lib)

    def post_ack(self, value):
        words = self._ptr_array([self.ack(self.peer_base[r], self.rank) for r in self.peers])
        _lib.check(self.lib.drs_fw_post_words(len(self.peers), words, value, self.tiles._stream()), self.lib)

    def timed_out(self):
        torch = self.tiles.torch
        flag = _as_tensor(torch, self.base + 512, (1,), torch.int32)
        return bool(flag.item())

    def close(self):
        for r, ptr in self.peer_base.items():
            self.lib.drs_p2p_close(C.c_void_p(ptr))
        self.peer_base = {}
        if self.base:
            self.lib.drs_p2p_free(C.c_void_p(self.base))
            self.base = None


def _as_tensor(torch, addr, shape, dtype):
    """torch tensor over raw device memory owned by the library (no copy)."""
    class _Holder(object):
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f4" if dtype == torch.float32 else "<i4",
                                  "data": (int(addr), False), "version": 2, "strides": None}
    return torch.as_tensor(h, device="cuda")


def build_sharded_p2p(tiles, npad, rank, world, dist, comm, lookahead=True, with_pred=True):
    """Sharded FW whose pivot step stores its panel straight into the peers' buffers (fused compute +
    broadcast over NVLink); same schedule as build_sharded(algorithm="fw")."""
    nb = npad // TILE
    starts = block_partition(nb, world)
    b0, b1 = starts[rank], starts[rank + 1]
    row0 = b0 * TILE
    local = tiles.empty((b1 - b0) * TILE, npad)
    tiles.init(local, row0)
    bufs = [comm.local_panel(0), comm.local_panel(1)]
    base = comm.builds * (nb + 4) + 1
    comm.builds += 1
    dist.barrier()                      # nobody may still be using the buffers of an earlier build

    def own_panel(kb):
        return local[(kb - b0) * TILE:(kb - b0 + 1) * TILE]

    def pivot(kb):
        """Owner side: wait until the peers are done with the buffer panel kb goes to, then close + push."""
        if kb >= 2:
            comm.wait_acks(base + kb - 1)       # ack value of round kb-2
        comm.pivot_row(own_panel(kb), kb, kb % 2, base + kb + 1)

    if owner_of(0, starts) == rank:
        pivot(0)
    for kb in range(nb):
        owner = owner_of(kb, starts)
        if owner == rank:
            panel = own_panel(kb)
        else:
            comm.wait_signal(base + kb + 1)
            panel = bufs[kb % 2]
        skip = (kb - b0) if owner == rank else -1
        if kb + 1 < nb and owner_of(kb + 1, starts) == rank:
            lb = kb + 1 - b0
            if lookahead:
                tiles.update(local, panel, kb, skip, lb, lb + 1)
                pivot(kb + 1)
                tiles.update(local, panel, kb, skip, 0, lb)
                tiles.update(local, panel, kb, skip, lb + 1, b1 - b0)
            else:
                tiles.update(local, panel, kb, skip, 0, b1 - b0)
                pivot(kb + 1)
        else:
            tiles.update(local, panel, kb, skip, 0, b1 - b0)
        comm.post_ack(base + kb + 1)
    pred_local = None
    if with_pred:
        pred_local = tiles.empty_pred((b1 - b0) * TILE, npad)
        tiles.pred(local, row0, pred_local)
    tiles.synchronize()
    if comm.timed_out():
        raise RuntimeError("fused pivot broadcast timed out waiting for a peer (rank %d)" % rank)
    return local, pred_local, row0


def build_sharded(tiles, npad, rank=0, world=1, dist=None, lookahead=True, with_pred=True, algorithm="fw"):
    """Runs the sharded build.  ``tiles`` is a backend (CudaTiles), ``dist`` the
    torch.distributed module (or None when world == 1).

    algorithm "fw": blocked Floyd-Warshall with pivot-panel broadcasts (below).
    algorithm "sssp": every rank runs the delta-stepping searches of its own source rows;
    sources are independent, so there is no collective at all.
    algorithm "sep": the separator-partitioned build (csrc/sep_apsp.cu): every rank closes the parts and the separator
    graph (a few ms, replicated) and expands its own rows; no collective either.

    Returns (local, pred_local, row0): this rank's rows of the distance and
    predecessor matrices and the global index of its first row.
    """
    nb = npad // TILE
    starts = block_partition(nb, world)
    b0, b1 = starts[rank], starts[rank + 1]
    row0 = b0 * TILE
    local = tiles.empty((b1 - b0) * TILE, npad)
    if algorithm in ("sssp", "sep"):
        if algorithm == "sep":
            tiles.sep_rows(local, row0)
        else:
            tiles.sssp_rows(local, row0)
        pred_local = None
        if with_pred:
            pred_local = tiles.empty_pred((b1 - b0) * TILE, npad)
            tiles.pred(local, row0, pred_local)
        tiles.synchronize()
        return local, pred_local, row0
    tiles.init(local, row0)
    bufs = [tiles.empty(TILE, npad), tiles.empty(TILE, npad)] if world > 1 else [None, None]

    def own_panel(kb):
        return local[(kb - b0) * TILE:(kb - b0 + 1) * TILE]

    def bcast(kb):
        """Posts the broadcast of pivot panel kb; returns (panel tensor, work handle or None)."""
        owner = owner_of(kb, starts)
        panel = own_panel(kb) if owner == rank else bufs[kb % 2]
        if world == 1:
            return panel, None
        return panel, dist.broadcast(panel, src=owner, async_op=True)

    # round 0 pivot
    if owner_of(0, starts) == rank:
        tiles.pivot_row(own_panel(0), 0)
    cur = bcast(0)
    for kb in range(nb):
        panel, work = cur
        if work is not None:
            work.wait()
        owner = owner_of(kb, starts)
        skip = (kb - b0) if owner == rank else -1
        nxt = None
        if kb + 1 < nb:
            nowner = owner_of(kb + 1, starts)
            if lookahead and nowner == rank:
                # update only block-row kb+1, then its pivot step, then start the next broadcast
                lb = kb + 1 - b0
                tiles.update(local, panel, kb, skip, lb, lb + 1)
                tiles.pivot_row(own_panel(kb + 1), kb + 1)
                nxt = bcast(kb + 1)
                tiles.update(local, panel, kb, skip, 0, lb)
                tiles.update(local, panel, kb, skip, lb + 1, b1 - b0)
            elif lookahead:
                nxt = bcast(kb + 1)          # post the receive early; it overlaps with this round's update
                tiles.update(local, panel, kb, skip, 0, b1 - b0)
            else:
                tiles.update(local, panel, kb, skip, 0, b1 - b0)
                if nowner == rank:
                    tiles.pivot_row(own_panel(kb + 1), kb + 1)
                nxt = bcast(kb + 1)
        else:
            tiles.update(local, panel, kb, skip, 0, b1 - b0)
        cur = nxt
    pred_local = None
    if with_pred:
        pred_local = tiles.empty_pred((b1 - b0) * TILE, npad)
        tiles.pred(local, row0, pred_local)
    tiles.synchronize()
    return local, pred_local, row0


def gather_replica(local, pred_local, npad, world, dist, tiles):
    """All-gathers the shards so that every rank holds the full matrices (the per-tick
    dispatch kernels run on one GPU against a full copy)."""
    if world == 1:
        return local, pred_local
    nb = npad // TILE
    starts = block_partition(nb, world)
    rank = dist.get_rank()
    full = tiles.empty(npad, npad)
    parts = [full[starts[r] * TILE:starts[r + 1] * TILE] for r in range(world)]
    parts[rank].copy_(local)
    for r in range(world):          # shard sizes may differ by one block: broadcast part by part
        dist.broadcast(parts[r], src=r)
    full_pred = None
    if pred_local is not None:
        full_pred = tiles.empty_pred(npad, npad)
        pparts = [full_pred[starts[r] * TILE:starts[r + 1] * TILE] for r in range(world)]
        pparts[rank].copy_(pred_local)
        for r in range(world):
            dist.broadcast(pparts[r], src=r)
    return full, full_pred
