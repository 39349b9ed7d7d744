"""Separator-partitioned build (csrc/sep_apsp.cu, DRS_APSP_SEP) through the C-ABI: the matrix it writes is the matrix of
the oracle (scipy-free restatement of the reference's igraph Dijkstra, oracle/routing_oracle.py) and of the SSSP build,
bit for bit, on integer-weighted maps -- undirected, directed, disconnected, with and without coordinates, one part or
many, in both element types, whole or by row shards.  Replaces lib/RoutingEngine.py:87,199,212 and
lib/optimUtils.py:528,531,659 like the other builders."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden_map_path, oracle_graph, oracle_graph_from_data

pytestmark = pytest.mark.gpu


def _data(n, src, dst, tt, directed=False, coords=None):
    from drs_abm_b200.gml import RoadGraphData
    vattrs = {"name": [str(i) for i in range(n)]}
    if coords is not None:
        vattrs["lng"], vattrs["lat"] = [float(c) for c in coords[0]], [float(c) for c in coords[1]]
    tt = np.asarray(tt, dtype=np.float64)
    return RoadGraphData(n, directed, src, dst, vattrs, {"ttmedian": tt, "length": 10.0 * tt})


def _random_planar(n_side, seed, directed=False, coords=True, drop=0.1, islands=1, wmax=90):
    """Grid with random integer weights, a fraction of the edges dropped (components are kept: unreachable pairs are part
    of the case), optionally several disjoint copies, optionally one-way arcs with independent weights."""
    rs = np.random.RandomState(seed)
    src, dst, xs, ys = [], [], [], []
    for isl in range(islands):
        base = isl * n_side * n_side
        for i in range(n_side):
            for j in range(n_side):
                xs.append(j + isl * (n_side + 3)); ys.append(i)
                if j + 1 < n_side and rs.rand() > drop:
                    src.append(base + i * n_side + j); dst.append(base + i * n_side + j + 1)
                if i + 1 < n_side and rs.rand() > drop:
                    src.append(base + i * n_side + j); dst.append(base + (i + 1) * n_side + j)
    src, dst = np.array(src), np.array(dst)
    if directed:
        back = rs.rand(len(src)) < 0.8
        src, dst = np.concatenate([src, dst[back]]), np.concatenate([dst, src[back]])
    tt = rs.randint(5, wmax, size=len(src)).astype(np.float64)
    n = islands * n_side * n_side
    perm = rs.permutation(n)                     # the caller's numbering carries no locality
    inv = np.empty(n, dtype=np.int64); inv[perm] = np.arange(n)
    cs = (np.array(xs)[inv], np.array(ys)[inv]) if coords else None
    return _data(n, perm[src], perm[dst], tt, directed, cs)


def _matrices(data, target, mode, algorithms, levels=1):
    from drs_abm_b200 import _lib
    from drs_abm_b200.routing import RoadGraph
    out = {}
    for algo in algorithms:
        G = RoadGraph(data, mode=mode, algorithm=algo, numbering="partition", part_target=target,
                      second_level_min=1 if levels == 2 else 10 ** 9)
        out[algo] = G.dist_rows(0, G.n)
        if algo == _lib.APSP_SEP:
            info = [C.c_int32(0) for _ in range(8)]
            _lib.check(G._lib.drs_sep_info(G._handle, *[C.byref(x) for x in info[:5]]), G._lib)
            _lib.check(G._lib.drs_sep_info2(G._handle, C.byref(info[5]), C.byref(info[6]), C.byref(info[7])), G._lib)
            out["parts"], out["nsep"], out["groups"], out["top"], out["used16"] = info[0].value, info[1].value, info[5].value, info[6].value, info[7].value
        G.close()
    return out


@pytest.mark.parametrize("name", ["g8_int", "g20_int", "g20_uniform"])
@pytest.mark.parametrize("target", [16, 60, 4096])
@pytest.mark.parametrize("mode", [0, 1])
def test_sep_matrix_equals_the_oracle_on_the_golden_maps(name, target, mode):
    from drs_abm_b200 import _lib
    from drs_abm_b200.gml import read_gml
    got = _matrices(read_gml(golden_map_path(name)), target, mode, [_lib.APSP_SEP])
    assert np.array_equal(got[_lib.APSP_SEP], oracle_graph(name).dist_matrix())
    assert got["parts"] >= (1 if target == 4096 else 2)


CASES = {
    "undirected_24": dict(n_side=24, seed=3),
    "directed_20": dict(n_side=20, seed=4, directed=True),
    "no_coordinates_22": dict(n_side=22, seed=5, coords=False),
    "islands_3x12": dict(n_side=12, seed=6, islands=3, drop=0.2),
    "directed_islands_no_coordinates": dict(n_side=14, seed=7, islands=2, directed=True, coords=False, drop=0.25),
    # distances beyond 15 bits: the output tiles stay on the 32-bit path; and right at the edge of the 16-bit range
    "heavy_weights_24": dict(n_side=24, seed=8, wmax=9000),
    "edge_of_16_bits_24": dict(n_side=24, seed=9, wmax=1200),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("target", [24, 130])
@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("levels", [1, 2])
def test_sep_matrix_equals_oracle_and_sssp_on_irregular_maps(case, target, mode, levels):
    from drs_abm_b200 import _lib
    data = _random_planar(**CASES[case])
    got = _matrices(data, target, mode, [_lib.APSP_SEP, _lib.APSP_SSSP], levels)
    want = oracle_graph_from_data(data).dist_matrix()
    assert np.array_equal(got[_lib.APSP_SSSP], want)
    assert np.array_equal(got[_lib.APSP_SEP], want)              # +inf where unreachable, on both sides
    assert got["parts"] >= 3 and got["nsep"] > 0
    if levels == 2 and target == 24:
        assert got["groups"] >= 2 and 0 < got["top"] < got["nsep"]   # the separator graph really went through level 2
    finite = want[np.isfinite(want)]
    if mode == 1 or case == "heavy_weights_24":
        assert got["used16"] == 0                                    # int32 matrices and long distances: 32-bit tiles
    elif finite.max() < 16000:
        assert got["used16"] == 1                                    # VIADDMNMX.U16x2 tiles, same matrix


def test_sep_on_the_5k_city_equals_sssp_and_paths_follow():
    """C2 (5 040 vertices, 27 parts): whole matrix on the device equal to the SSSP build's; path queries, which read the
    matrix through the canonical predecessor rule, return the oracle's edge ids."""
    import torch
    from drs_abm_b200 import _lib, synth
    from drs_abm_b200.apsp_dist import CudaTiles, build_sharded
    from drs_abm_b200.routing import RoadGraph
    data = synth.config_graph("C2")
    G = RoadGraph(data, numbering="partition")
    h = G.upload()
    tiles = CudaTiles(h, _lib.MODE_F32)
    npad = (data.n + 127) // 128 * 128
    a, _, _ = build_sharded(tiles, npad, algorithm="sep", with_pred=False)
    b, _, _ = build_sharded(tiles, npad, algorithm="sssp", with_pred=False)
    assert torch.equal(a, b)
    # two row shards of unequal size give the same rows as the whole build
    top = torch.empty((1280, npad), dtype=torch.float32, device="cuda")
    rest = torch.empty((npad - 1280, npad), dtype=torch.float32, device="cuda")
    tiles.sep_rows(top, 0); tiles.sep_rows(rest, 1280); torch.cuda.synchronize()
    assert torch.equal(torch.cat([top, rest]), a)
    assert G.algorithm == _lib.APSP_AUTO
    G.build()                                                   # AUTO picks the partitioned build on this map
    prof = np.zeros(8)
    _lib.check(G._lib.drs_apsp_sep_profile(h, _lib.ptr_f64(prof)), G._lib)
    assert prof.sum() > 0
    O = oracle_graph_from_data(data)
    rs = np.random.RandomState(2)
    for _ in range(40):
        s, t = int(rs.randint(data.n)), int(rs.randint(data.n))
        assert G.epath(s, t) == O.epath(s, t)
    G.close()


@pytest.mark.parametrize("name", ["C1_float", "C2_float", "g12_float", "g10_dfloat"])
def test_sep_on_float_travel_times_stays_within_the_float_tolerance(name):
    """Float-weighted maps (the reference's own GMLs carry float ttmedian): the build re-associates the fp32 sums, so the matrix
    is compared with the fp64 oracle at the north star's 1e-5 (measured: a few 1e-7); the tiles stay 32-bit; AUTO selects the
    build; and the path queries -- fp64 sums along the path the matrix selects, hop by hop, up to ~140 hops on the 5k map --
    stay within the same tolerance of the true shortest travel time."""
    from drs_abm_b200 import _lib, synth
    from drs_abm_b200.gml import read_gml
    from drs_abm_b200.routing import RoadGraph
    data = read_gml(golden_map_path(name)) if name.startswith("g") else synth.config_graph(name)
    G = RoadGraph(data, part_target=64 if data.n < 1000 else None, second_level_min=1 if name == "C2_float" else None)
    G.build()                                                     # AUTO
    info = [C.c_int32(0) for _ in range(8)]
    _lib.check(G._lib.drs_sep_info(G._handle, *[C.byref(x) for x in info[:5]]), G._lib)
    _lib.check(G._lib.drs_sep_info2(G._handle, C.byref(info[5]), C.byref(info[6]), C.byref(info[7])), G._lib)
    prof = np.zeros(8)
    _lib.check(G._lib.drs_apsp_sep_profile(G._handle, _lib.ptr_f64(prof)), G._lib)
    assert info[0].value > 1 and info[4].value == 1 and prof.sum() > 0 and info[7].value == 0     # SEP ran, 32-bit tiles
    O = oracle_graph_from_data(data)
    rs = np.random.RandomState(4)
    rows = np.unique(rs.randint(0, data.n, 60))
    want = O.dist_rows(rows)
    got = np.stack([G.dist_rows(int(r), 1)[0] for r in rows])
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.max(np.abs(got[fin] - want[fin]) / np.maximum(want[fin], 1e-12)) < 1e-5
    s = rs.randint(0, data.n, 400); t = rs.randint(0, data.n, 400)
    tt, ln, hops = G.path_sums(s, t)
    true = O.dist_rows(np.unique(s))
    idx = {int(v): i for i, v in enumerate(np.unique(s))}
    ref = np.array([true[idx[int(a)], int(b)] for a, b in zip(s, t)])
    ok = np.isfinite(ref)
    assert np.max(np.abs(tt[ok] - ref[ok]) / np.maximum(ref[ok], 1e-12)) < 1e-5
    if name == "C2_float":
        assert hops.max() > 100
    G.close()


def test_partition_is_validated_and_follows_weight_updates():
    from drs_abm_b200 import _lib, synth
    from drs_abm_b200.routing import RoadGraph
    data = synth.config_graph("C1")
    G = RoadGraph(data, algorithm=_lib.APSP_SEP, numbering="partition", part_target=50)
    h = G.upload()
    lib = G._lib
    bad = _lib.as_i32([0, 200, 400])                             # two "interiors" that are joined by edges
    assert lib.drs_graph_set_partition(h, 2, _lib.ptr_i32(bad), 0, None) == _lib.ERR_ARG
    assert b"joins the interiors" in lib.drs_last_error()
    pp = _lib.as_i32(G._part_ptr)
    _lib.check(lib.drs_graph_set_partition(h, len(pp) - 1, _lib.ptr_i32(pp), 0, None), lib)
    first = G.dist_rows(0, G.n)
    tt = np.array(G._eattrs["ttmedian"]); tt[::3] += 7.0
    G.es["ttmedian"] = tt.tolist()                               # the igraph-style write invalidates the matrix
    second = G.dist_rows(0, G.n)
    from drs_abm_b200.gml import RoadGraphData
    d2 = RoadGraphData(data.n, data.directed, data.src, data.dst, data.vattrs, dict(data.eattrs, ttmedian=tt))
    assert np.array_equal(second, oracle_graph_from_data(d2).dist_matrix()) and not np.array_equal(first, second)
    # the same graph turned float-weighted: the next build leaves the 16-bit path and stays within the float tolerance
    used = C.c_int32(0)
    _lib.check(lib.drs_sep_info2(h, None, None, C.byref(used)), lib)
    assert used.value == 1
    tt2 = tt.copy(); tt2[::5] += 0.37
    G.es["ttmedian"] = tt2.tolist()
    third = G.dist_rows(0, G.n)
    _lib.check(lib.drs_sep_info2(h, None, None, C.byref(used)), lib)
    assert used.value == 0
    d3 = RoadGraphData(data.n, data.directed, data.src, data.dst, data.vattrs, dict(data.eattrs, ttmedian=tt2))
    want3 = oracle_graph_from_data(d3).dist_matrix()
    assert np.max(np.abs(third - want3) / np.maximum(want3, 1e-12)) < 1e-5
    G.close()


def test_graph_without_small_separators_falls_back_and_explicit_sep_stays_exact():
    """A random graph of average degree 16 has no small separators: the cover takes most of the vertices.  AUTO then leaves the
    separator build alone (drs_sep_info: usable = 0) and uses another builder; forcing APSP_SEP is still exact -- the
    construction does not depend on the separators being few, only its speed does."""
    from drs_abm_b200 import _lib
    from drs_abm_b200.routing import RoadGraph
    rs = np.random.RandomState(11)
    n, m = 500, 4000
    src, dst = rs.randint(0, n, m), rs.randint(0, n, m)
    keep = src != dst
    data = _data(n, src[keep], dst[keep], rs.randint(1, 50, size=int(keep.sum())).astype(np.float64))
    want = oracle_graph_from_data(data).dist_matrix()
    G = RoadGraph(data, part_target=64)
    G.build()
    info = [C.c_int32(0) for _ in range(5)]
    _lib.check(G._lib.drs_sep_info(G._handle, *[C.byref(x) for x in info]), G._lib)
    assert info[1].value * 3 > n and info[4].value == 0           # separators are most of the graph: not usable
    assert np.array_equal(G.dist_rows(0, n), want)
    G.close()
    G = RoadGraph(data, algorithm=_lib.APSP_SEP, part_target=64, second_level_min=1)
    assert np.array_equal(G.dist_rows(0, n), want)
    G.close()


def test_a_second_graph_reuses_the_blocks_of_the_first():
    """Device and pinned blocks of a closed graph go to a process-wide cache (drs_api.cu): a second graph of the same map builds
    and answers queries without a single cudaMalloc / cudaHostAlloc (drs_alloc_count), and drs_pool_trim gives the blocks back."""
    from drs_abm_b200 import _lib, synth
    from drs_abm_b200.routing import RoadGraph
    data = synth.config_graph("C2")
    lib = _lib.load()
    o, d = synth.random_od_pairs(data.n, 512, 5)
    _lib.check(lib.drs_pool_trim(), lib)                          # whatever earlier tests left behind
    G = RoadGraph(data)
    first = G.path_sums(o, d)[0]
    G.matrix_lookup(o, d)
    G.close()
    a0 = lib.drs_alloc_count()
    G = RoadGraph(data)
    second = G.path_sums(o, d)[0]
    G.matrix_lookup(o, d)
    G.close()
    assert lib.drs_alloc_count() == a0, "a fresh graph of the same size allocated device or pinned memory"
    assert np.array_equal(first, second)
    _lib.check(lib.drs_pool_trim(), lib)
    G = RoadGraph(data)
    G.build()
    assert lib.drs_alloc_count() > a0                              # after a trim the blocks come from the driver again
    G.close()


@pytest.mark.parametrize("tile_rows", ["128", "64", "32"])
def test_every_height_of_the_16_bit_output_tiles(tile_rows, monkeypatch):
    """DRS_SEP_TILE_ROWS selects the height of the 16-bit output tiles (64 rows = 128 threads, six CTAs per SM, is the default):
    same matrix at every height, whole and by row shards, undirected (mirrored tiles) and directed."""
    import torch
    from drs_abm_b200 import _lib, synth
    from drs_abm_b200.apsp_dist import CudaTiles, build_sharded
    from drs_abm_b200.routing import RoadGraph
    monkeypatch.setenv("DRS_SEP_TILE_ROWS", tile_rows)
    for data in (synth.config_graph("C2"), _random_planar(n_side=30, seed=12, directed=True)):
        G = RoadGraph(data, part_target=100, second_level_min=1)
        h = G.upload()
        tiles = CudaTiles(h, _lib.MODE_F32)
        npad = (data.n + 127) // 128 * 128
        a, _, _ = build_sharded(tiles, npad, algorithm="sep", with_pred=False)
        used = C.c_int32(0)
        _lib.check(G._lib.drs_sep_info2(h, None, None, C.byref(used)), G._lib)
        assert used.value == 1
        b, _, _ = build_sharded(tiles, npad, algorithm="sssp", with_pred=False)
        assert torch.equal(a, b)
        top = torch.empty((384, npad), dtype=torch.float32, device="cuda")
        tiles.sep_rows(top, 256); torch.cuda.synchronize()
        assert torch.equal(top, a[256:640])
        G.close()
