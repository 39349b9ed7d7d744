"""Drop-in routing engine: the reference's ``RoutingEngine`` call surface over the
B200 travel-time matrix.

Mirrors /root/reference lib/RoutingEngine.py:20-229 (same constructor signature,
method names, argument meaning, return shapes and error behaviour) and the
subset of the python-igraph ``Graph`` API that lib/Agents.py, lib/optimUtils.py,
lib/Optimization.py and the experiment scripts touch through ``router.G``
("G proxy").  Every shortest-path question is answered by
libdrs_b200.so (hand-written sm_100a CUDA behind a C-ABI); there is no CPU
fallback -- if the library or the GPU is missing the first query raises.
"""
import ctypes as C
import math
import os
import re

import numpy as np

from . import _lib
from .gml import RoadGraphData, read_gml

MPH_2_MPS = 0.44704            # lib/Constants.py:152
UNCERTAINTY_MULTIPLIER = 1.5   # lib/Constants.py:99

_NUM = re.compile(r"[-+]?(?:\d+\.?\d*|\.\d+)(?:[eE][-+]?\d+)?")


class Vertex(object):
    """igraph.Vertex stand-in: ``v['lng']``, ``v['lat']``, ``v['name']``, ``v.index``."""
    __slots__ = ("graph", "index")

    def __init__(self, graph, index):
        self.graph = graph
        self.index = int(index)

    def __getitem__(self, key):
        return self.graph._vattrs[key][self.index]

    def __setitem__(self, key, value):
        self.graph._own_vattr(key)[self.index] = value
        if key in ("name", "lng", "lat"):
            self.graph._by_name = self.graph._by_coord = None

    def attributes(self):
        return {k: v[self.index] for k, v in self.graph._vattrs.items()}

    def __eq__(self, other):
        return isinstance(other, Vertex) and other.graph is self.graph and other.index == self.index

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((id(self.graph), self.index))

    def __repr__(self):
        return "Vertex(%d, %r)" % (self.index, self.attributes())


class Edge(object):
    """igraph.Edge stand-in: ``.index/.source/.target/.tuple``, ``e['length']``, ``e['ttmedian'] = x``."""
    __slots__ = ("graph", "index")

    def __init__(self, graph, index):
        self.graph = graph
        self.index = int(index)

    @property
    def source(self):
        return int(self.graph._src[self.index])

    @property
    def target(self):
        return int(self.graph._dst[self.index])

    @property
    def tuple(self):
        return (self.source, self.target)

    def __getitem__(self, key):
        v = self.graph._eattrs[key][self.index]
        return float(v) if isinstance(v, (np.floating, np.integer)) else v

    def __setitem__(self, key, value):
        self.graph._set_edge_attr(key, self.index, value)

    def attributes(self):
        return {k: self[k] for k in self.graph._eattrs}


class VertexSeq(object):
    def __init__(self, graph):
        self.graph = graph

    def __len__(self):
        return self.graph.n

    def __iter__(self):
        return (Vertex(self.graph, i) for i in range(self.graph.n))

    def __getitem__(self, key):
        if isinstance(key, str):
            return list(self.graph._vattrs[key])
        i = int(key)
        if i < 0:
            i += self.graph.n
        if not 0 <= i < self.graph.n:
            raise IndexError("vertex index out of range")
        return Vertex(self.graph, i)

    def find(self, *args, **kwds):
        """``vs.find(name=...)`` / ``vs.find("...")`` / ``vs.find(i)``; unknown -> ValueError like igraph
        (relied on at lib/Agents.py:1508-1512)."""
        if args:
            if isinstance(args[0], str):
                kwds = dict(kwds, name=args[0])
            else:
                return self[int(args[0])]
        if list(kwds) == ["name"]:
            return Vertex(self.graph, self.graph.node_index(kwds["name"]))
        for i in range(self.graph.n):
            if all(self.graph._vattrs[k][i] == v for k, v in kwds.items()):
                return Vertex(self.graph, i)
        raise ValueError("no such vertex")


class EdgeSeq(object):
    def __init__(self, graph):
        self.graph = graph

    def __len__(self):
        return self.graph.m

    def __iter__(self):
        return (Edge(self.graph, i) for i in range(self.graph.m))

    def __getitem__(self, key):
        if isinstance(key, str):
            return [float(x) if isinstance(x, (np.floating, np.integer)) else x for x in self.graph._eattrs[key]]
        i = int(key)
        if i < 0:
            i += self.graph.m
        if not 0 <= i < self.graph.m:
            raise IndexError("edge index out of range")
        return Edge(self.graph, i)

    def __setitem__(self, key, values):
        """``G.es["ttmedian"] = array`` (generateMetadata.py:42-43)."""
        self.graph._set_edge_attr(key, None, values)


def _spread16(v):
    """Bits of a 16-bit integer array moved to the even positions (Morton interleave)."""
    v = v.astype(np.uint64) & np.uint64(0xffff)
    v = (v | (v << np.uint64(8))) & np.uint64(0x00ff00ff)
    v = (v | (v << np.uint64(4))) & np.uint64(0x0f0f0f0f)
    v = (v | (v << np.uint64(2))) & np.uint64(0x33333333)
    v = (v | (v << np.uint64(1))) & np.uint64(0x55555555)
    return v


def locality_order(n, src, dst, lng=None, lat=None, native=False):
    """Vertex ids in the order of the DEVICE numbering: order[k] = caller's id of device vertex k.

    The kernels index distance rows by vertex id, so neighbouring vertices should have neighbouring ids (the same
    50k-node map builds in 84 ms numbered along a Morton curve, 93 ms row-major, 141 ms numbered at random --
    profiles/r01_sssp_numbering_probe.jsonl).  With usable coordinates: Morton (Z-order) curve over (lng, lat)
    ranked to at most 16 bits each; otherwise reverse Cuthill-McKee over the adjacency."""
    try:
        x, y = np.ascontiguousarray(lng, dtype=np.float64), np.ascontiguousarray(lat, dtype=np.float64)
        if x.shape != (n,) or y.shape != (n,):
            raise ValueError
        if native:                                   # the same order from the library's C++ helper (for bindings without numpy;
                                                     # numpy's vectorised sorts are the faster of the two here)
            lib = _lib.load()
            order = np.empty(n, dtype=np.int32)
            if lib.drs_locality_order(n, _lib.ptr_f64(x), _lib.ptr_f64(y), _lib.ptr_i32(order)) == 0:
                return order
            raise ValueError
        if not (np.isfinite(x).all() and np.isfinite(y).all()):
            raise ValueError
        ux, qx = np.unique(x, return_inverse=True)      # rank of each coordinate among the distinct values: a regular
        uy, qy = np.unique(y, return_inverse=True)      # grid gets its exact Z-curve, an irregular map a density-equalised one
        if len(ux) == 1 and len(uy) == 1:
            raise ValueError
        if len(ux) > 65536:
            qx = qx.astype(np.int64) * 65536 // len(ux)
        if len(uy) > 65536:
            qy = qy.astype(np.int64) * 65536 // len(uy)
        return np.argsort(_spread16(qx) | (_spread16(qy) << np.uint64(1)), kind="stable").astype(np.int32)
    except (TypeError, ValueError):
        from scipy.sparse import csr_matrix
        from scipy.sparse.csgraph import reverse_cuthill_mckee
        if len(src) == 0:
            return np.arange(n, dtype=np.int32)
        A = csr_matrix((np.ones(2 * len(src), dtype=np.int8), (np.concatenate([src, dst]), np.concatenate([dst, src]))), shape=(n, n))
        return np.asarray(reverse_cuthill_mckee(A, symmetric_mode=True), dtype=np.int32)


# below this many separator vertices their graph is closed by the plain blocked Floyd-Warshall (a few rounds); above, by
# the partitioned build itself one level up (DRS_SEP_LEVEL2_MIN overrides)
SECOND_LEVEL_MIN_SEPARATORS = int(os.environ.get("DRS_SEP_LEVEL2_MIN", 1024))


# largest all-pairs matrix the bindings plan for (bytes): beyond it a map gets the plain locality numbering
MATRIX_BYTES_LIMIT = 120e9


def default_part_target(n):
    """Vertices per part of the separator partition: the expansion stage costs n^2 x (separators around a part, which
    grow like the square root of the part) and the separator closure (number of separators)^3, which shrinks with
    larger parts; on grid-like road maps the optimum is flat (measured on the 50k-node city: 128 .. 520 vertices per part
    build within 4 % of each other); a power of two keeps the part interiors inside whole 128-column tiles."""
    return int(min(2048, max(192, 1 << max(0, int(round(math.log2(max(n, 1) / 196.0)))))))


def partition_order(n, src, dst, lng=None, lat=None, target=None, second_level=False):
    """(order, part_ptr) of the part-major device numbering (include/drs_b200.h: drs_partition_order); with second_level
    also the group boundaries of the separators (relative to the first separator id)."""
    lib = _lib.load()
    target = int(target or default_part_target(n))
    src, dst = _lib.as_i32(src), _lib.as_i32(dst)
    x = y = None
    try:
        x, y = np.ascontiguousarray(lng, dtype=np.float64), np.ascontiguousarray(lat, dtype=np.float64)
        if x.shape != (n,) or y.shape != (n,):
            x = y = None
    except (TypeError, ValueError):
        x = y = None
    order = np.empty(n, dtype=np.int32)
    max_parts = max(4, 4 * (n // target + 1))
    part_ptr = np.zeros(max_parts + 1, dtype=np.int32)
    nparts, nsp = C.c_int32(0), C.c_int32(0)
    sep_ptr = np.zeros(max_parts + 1, dtype=np.int32)
    _lib.check(lib.drs_partition_order(n, len(src), _lib.ptr_i32(src), _lib.ptr_i32(dst),
                                       _lib.ptr_f64(x) if x is not None else None, _lib.ptr_f64(y) if y is not None else None,
                                       target, _lib.ptr_i32(order), _lib.ptr_i32(part_ptr), max_parts, C.byref(nparts),
                                       _lib.ptr_i32(sep_ptr), C.byref(nsp)), lib)
    if second_level:
        return order, part_ptr[:nparts.value + 1].copy(), sep_ptr[:nsp.value + 1].copy()
    return order, part_ptr[:nparts.value + 1].copy()


class RoadGraph(object):
    """The object callers know as ``router.G``: igraph look-alike backed by a device
    CSR graph and the resident travel-time / predecessor matrices.

    The matrix is built lazily for one weight attribute at a time and rebuilt
    after any write to that attribute (optimizationTesting.py:221-226 mutates
    ``ttmedian`` after load).
    """

    def __init__(self, data, mode=_lib.MODE_F32, algorithm=_lib.APSP_AUTO, pred=_lib.PRED_LAZY, numbering="auto",
                 part_target=None, second_level_min=None):
        if not isinstance(data, RoadGraphData):
            raise TypeError("RoadGraph needs RoadGraphData (use RoadGraph.Read_GML)")
        self.n, self.m = data.n, data.m
        self.directed = data.directed
        self._src, self._dst = data.src, data.dst
        # attribute columns are shared with `data` until the first write through this graph (copy on write: a fresh graph of a
        # 50k-vertex map would otherwise spend milliseconds copying lists it never changes)
        self._vattrs = dict(data.vattrs)
        self._coords = getattr(data, "coords", None)
        self._eattrs = {k: (np.asarray(v, dtype=np.float64) if isinstance(v, np.ndarray) else v) for k, v in data.eattrs.items()}
        self._owned = set()
        self.mode = mode
        self.pred = pred              # _lib.PRED_LAZY (canonical rule evaluated per hop of a path query) | PRED_MATRIX
        self.algorithm = algorithm    # _lib.APSP_FW (blocked Floyd-Warshall) | APSP_SSSP (N delta-stepping searches) | APSP_AUTO
        self._handle = None
        self._weight_attr = None      # attribute the device weights / matrix were built from
        self._lib = None
        self._by_name = self._by_coord = None    # name / coordinate indexes are built on first look-up
        # Device numbering of the vertices: "locality" = along a space-filling curve (locality_order), "input" = the
        # caller's.  Vertex ids are translated at every call into the library (dev_ids / host_cols); edge ids, and
        # with them adjacency order, canonical predecessors and paths, do not depend on it.
        # "partition" (the default) = part-major numbering over a vertex-separator partition (partition_order): the
        # interiors of the parts first, the separators last; parts follow the bisection tree, so it is a locality
        # numbering too, and it is what the separator-partitioned build (APSP_SEP) works on.  "locality" = the plain
        # space-filling curve of round 1.
        # "auto": "partition" when an all-pairs matrix of this map fits a GPU (the partition exists for the matrix build), else
        # "locality" (maps served by SSSP rows on demand, BASELINE config C5).
        if numbering not in ("auto", "partition", "locality", "input"):
            raise ValueError("numbering must be 'auto', 'partition', 'locality' or 'input'")
        if numbering == "auto":
            npad = (data.n + 127) // 128 * 128
            numbering = "partition" if 4.0 * npad * npad <= MATRIX_BYTES_LIMIT else "locality"
        self.numbering = numbering
        self.part_target = int(os.environ.get("DRS_SEP_TARGET", part_target or 0)) or None
        self._part_ptr = self._sep_part_ptr = None
        self.second_level_min = SECOND_LEVEL_MIN_SEPARATORS if second_level_min is None else int(second_level_min)
        self._new_of_old = self._old_of_new = None

    # ------------------------------------------------------------------ igraph surface
    @classmethod
    def Read_GML(cls, path, **kw):
        return cls(read_gml(path), **kw)

    @property
    def vs(self):
        return VertexSeq(self)

    @property
    def es(self):
        return EdgeSeq(self)

    def vcount(self):
        return self.n

    def ecount(self):
        return self.m

    def is_directed(self):
        return self.directed

    def get_shortest_paths(self, v, to=None, weights=None, output="vpath"):
        """``G.get_shortest_paths(o, to=d, weights=w, output='epath')[0]`` (lib/RoutingEngine.py:87)."""
        s, t = self._resolve(v), self._resolve(to)
        epath = self.epath(s, t, weights)
        if output == "epath":
            return [epath]
        vpath, cur = [s], s
        for e in epath:
            a, b = int(self._src[e]), int(self._dst[e])
            cur = b if a == cur else a
            vpath.append(cur)
        return [vpath if (epath or s == t) else []]

    def shortest_paths(self, source=None, target=None, weights=None):
        """``G.shortest_paths(src, tgt, weights=)`` -> list of lists (lib/optimUtils.py:528,531,659)."""
        srcs = list(source) if isinstance(source, (list, tuple)) else ([source] if source is not None else range(self.n))
        tgts = list(target) if isinstance(target, (list, tuple)) else ([target] if target is not None else range(self.n))
        si = [self._resolve(s) for s in srcs]
        ti = [self._resolve(t) for t in tgts]
        if not si or not ti:
            return [[] for _ in si]
        ss = np.repeat(np.asarray(si, dtype=np.int32), len(ti))
        tt = np.tile(np.asarray(ti, dtype=np.int32), len(si))
        d = self.matrix_lookup(ss, tt, weights)
        return [[float(x) for x in row] for row in d.reshape(len(si), len(ti))]

    distances = shortest_paths   # python-igraph >= 0.10 spelling

    # ------------------------------------------------------------------ node identity
    def _index_nodes(self):
        """(Re)builds the name -> vertex and (lng, lat) -> vertex dictionaries."""
        names = self._vattrs.get("name")
        self._by_name = {}
        if names is not None:
            for i in range(self.n - 1, -1, -1):
                self._by_name[names[i]] = i
        self._by_coord = {}
        if "lng" in self._vattrs and "lat" in self._vattrs:
            lng, lat = self._vattrs["lng"], self._vattrs["lat"]
            for i in range(self.n - 1, -1, -1):
                try:
                    self._by_coord[(float(lng[i]), float(lat[i]))] = i
                except (TypeError, ValueError):
                    pass

    def node_index(self, name):
        """Vertex id of a ``str((lng, lat))`` name.  The literal name wins; otherwise the two
        numbers in the string are matched against the (lng, lat) attributes, which makes
        ``"(3, 4)"``, ``"(3.0, 4.0)"`` and numpy-2 reprs such as
        ``"(np.float64(3.0), np.float64(4.0))"`` equivalent.
        Unknown -> ValueError (igraph's behaviour)."""
        if self._by_name is None:
            self._index_nodes()
        i = self._by_name.get(name)
        if i is not None:
            return i
        if isinstance(name, str):
            nums = _NUM.findall(name.replace("float64", "").replace("float32", "").replace("int64", ""))
            if len(nums) == 2:
                i = self._by_coord.get((float(nums[0]), float(nums[1])))
                if i is not None:
                    return i
        raise ValueError("no such vertex: %r" % (name,))

    def node_of_location(self, lng, lat):
        if self._by_coord is None:
            self._index_nodes()
        i = self._by_coord.get((float(lng), float(lat)))
        if i is None:
            return self.node_index(str((lng, lat)))
        return i

    def _resolve(self, v):
        if isinstance(v, Vertex):
            return v.index
        if isinstance(v, str):
            return self.node_index(v)
        i = int(v)
        if not 0 <= i < self.n:
            raise ValueError("vertex index %d out of range" % i)
        return i

    # ------------------------------------------------------------------ device state
    def _own_vattr(self, key):
        """The list of a vertex attribute, private to this graph from now on (created if absent)."""
        if ("v", key) not in self._owned:
            cur = self._vattrs.get(key)
            self._vattrs[key] = list(cur) if cur is not None else [None] * self.n
            self._owned.add(("v", key))
            if key in ("lng", "lat"):
                self._coords = None          # the numeric copies made at load time no longer describe this graph
        return self._vattrs[key]

    def _set_edge_attr(self, key, index, value):
        if index is None:
            vals = np.asarray(value, dtype=np.float64) if not isinstance(value, str) else value
            if isinstance(vals, np.ndarray) and vals.ndim == 0:
                vals = np.full(self.m, float(vals))
            if len(vals) != self.m:
                raise ValueError("attribute list length (%d) != edge count (%d)" % (len(vals), self.m))
            self._eattrs[key] = np.array(vals, dtype=np.float64)
            self._owned.add(("e", key))
        else:
            if key not in self._eattrs:
                self._eattrs[key] = np.zeros(self.m, dtype=np.float64)
                self._owned.add(("e", key))
            elif ("e", key) not in self._owned:
                cur = self._eattrs[key]
                self._eattrs[key] = np.array(cur, dtype=np.float64) if isinstance(cur, np.ndarray) else list(cur)
                self._owned.add(("e", key))
            self._eattrs[key][index] = value
        if key == self._weight_attr:
            self._weights_dirty = True
        self._exact_key = None

    _weights_dirty = False

    def _number(self):
        if self.numbering in ("locality", "partition") and self._new_of_old is None:
            order = None
            if self.numbering == "partition":
                lng, lat = self._coords if self._coords is not None else (self._vattrs.get("lng"), self._vattrs.get("lat"))
                order, self._part_ptr, self._sep_part_ptr = partition_order(self.n, self._src, self._dst, lng, lat,
                                                                            self.part_target, True)
            else:
                order = locality_order(self.n, self._src, self._dst, self._vattrs.get("lng"), self._vattrs.get("lat"))
            self._old_of_new = np.ascontiguousarray(order, dtype=np.int32)
            self._new_of_old = np.empty(self.n, dtype=np.int32)
            self._new_of_old[order] = np.arange(self.n, dtype=np.int32)

    def dev_ids(self, ids):
        """Caller's vertex ids -> the library's (int32 array).  Out of range -> ValueError (igraph's behaviour)."""
        a = _lib.as_i32(ids)
        if a.size and (int(a.min()) < 0 or int(a.max()) >= self.n):
            raise ValueError("vertex index out of range (n=%d)" % self.n)
        self._number()
        return a if self._new_of_old is None else np.ascontiguousarray(self._new_of_old[a])

    def host_cols(self, rows):
        """Columns of distance rows (last axis = device vertex ids, at least n wide) -> the caller's numbering, n wide."""
        self._number()
        rows = np.asarray(rows)
        return rows[..., :self.n] if self._new_of_old is None else rows[..., self._new_of_old]

    def _ensure(self, weights, build=True):
        """Device graph exists, carries `weights`, and (build=True) the APSP matrix is valid."""
        attr = weights if weights is not None else (self._weight_attr or "ttmedian")
        if self._valid and attr == self._weight_attr and not self._weights_dirty and self._lib is not None:
            return self._lib                                  # the common case of every query: nothing to do
        if not isinstance(attr, str):
            raise ValueError("weights must name an edge attribute")
        if attr not in self._eattrs:
            raise ValueError("no such edge attribute: %r" % (attr,))
        lib = self._lib or _lib.load()
        self._lib = lib
        tt = _lib.as_f64(self._eattrs[attr])
        if self._handle is None:
            h = C.c_void_p()
            length = _lib.as_f64(self._eattrs.get("length", np.zeros(self.m)))
            src, dst = self.dev_ids(self._src), self.dev_ids(self._dst)
            _lib.check(lib.drs_graph_create(self.n, self.m, 1 if self.directed else 0, _lib.ptr_i32(src),
                                            _lib.ptr_i32(dst), _lib.ptr_f64(tt), _lib.ptr_f64(length), C.byref(h)), lib)
            self._handle = h
            if self._part_ptr is not None:
                pp, sp = _lib.as_i32(self._part_ptr), _lib.as_i32(self._sep_part_ptr)
                two = len(sp) > 2 and int(self.n - pp[-1]) >= self.second_level_min
                _lib.check(lib.drs_graph_set_partition(h, len(pp) - 1, _lib.ptr_i32(pp), len(sp) - 1 if two else 0,
                                                       _lib.ptr_i32(sp) if two else None), lib)
            self._weight_attr = attr
            self._weights_dirty = False
            self._valid = False
        elif attr != self._weight_attr or self._weights_dirty:
            _lib.check(lib.drs_graph_set_weights(self._handle, _lib.ptr_f64(tt)), lib)
            self._weight_attr = attr
            self._weights_dirty = False
            self._valid = False
        if build and not self._valid:
            _lib.check(lib.drs_apsp_set_algorithm(self._handle, self.algorithm), lib)
            _lib.check(lib.drs_apsp_set_pred_mode(self._handle, self.pred), lib)
            _lib.check(lib.drs_apsp_build(self._handle, self.mode), lib)
            self._valid = True
        return lib

    _valid = False

    def upload(self, weights=None):
        """Creates the device CSR graph (and refreshes its weights) WITHOUT building the matrix;
        returns the drs_graph* handle.  Used by the multi-GPU build and by benchmarks."""
        self._ensure(weights, build=False)
        return self._handle

    def adopt(self, dist_tensor, pred_tensor):
        """Adopts externally built matrices (torch CUDA tensors from apsp_dist.gather_replica); pred_tensor may be None."""
        lib = self._ensure(None, build=False)
        _lib.check(lib.drs_apsp_adopt(self._handle, self.mode, C.c_void_p(dist_tensor.data_ptr()),
                                      C.c_void_p(pred_tensor.data_ptr()) if pred_tensor is not None else None,
                                      dist_tensor.shape[1]), lib)
        self._adopted = (dist_tensor, pred_tensor)     # keep the tensors alive
        self._valid = True

    def build(self, weights=None):
        """Builds (or rebuilds) the travel-time matrix now instead of at the first query."""
        self._ensure(weights)
        return self

    @property
    def handle(self):
        """drs_graph* with a valid matrix for the current weight attribute."""
        self._ensure(None)
        return self._handle

    def close(self):
        if self._handle is not None and self._lib is not None:
            self._lib.drs_graph_destroy(self._handle)
        self._handle = None
        self._valid = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ bulk queries
    def matrix_lookup(self, s, t, weights=None):
        """dist[s[i], t[i]] as float64 (inf = unreachable): one gather kernel."""
        lib = self._ensure(weights)
        s, t = self.dev_ids(s), self.dev_ids(t)
        out = np.empty(len(s), dtype=np.float64)
        _lib.check(lib.drs_query_matrix(self._handle, len(s), _lib.ptr_i32(s), _lib.ptr_i32(t), _lib.ptr_f64(out)), lib)
        return out

    def path_sums(self, s, t, weights=None):
        """(duration, length, hops) along the canonical paths: fp64 left-to-right sums."""
        lib = self._ensure(weights)
        s, t = self.dev_ids(s), self.dev_ids(t)
        n = len(s)
        tt, ln, hops = np.empty(n), np.empty(n), np.empty(n, dtype=np.int32)
        _lib.check(lib.drs_query_pairs(self._handle, n, _lib.ptr_i32(s), _lib.ptr_i32(t), _lib.ptr_f64(tt),
                                       _lib.ptr_f64(ln), _lib.ptr_i32(hops)), lib)
        return tt, ln, hops

    def sums_equal_matrix(self, weights=None):
        """True when the fp64 left-to-right sum along ANY shortest path equals the fp32 / int32 matrix entry bit for bit:
        every weight an integer and (n - 1) * max weight below 2^24 (sums of integers are exact in both)."""
        attr = weights if weights is not None else (self._weight_attr or "ttmedian")
        key = (attr, id(self._eattrs.get(attr)))
        if self._exact_key != key or self._weights_dirty:
            w = np.asarray(self._eattrs[attr], dtype=np.float64)
            self._exact = bool(w.size == 0 or (np.all(w == np.floor(w)) and float(w.max()) * max(self.n - 1, 1) < 2.0 ** 24))
            self._exact_key = key
        return self._exact

    _exact_key, _exact = None, False

    def pair_time(self, s, t, weights=None):
        """dist[s, t] as a Python float (inf = unreachable): one gather kernel on mapped host memory."""
        lib = self._ensure(weights)
        if not (0 <= s < self.n and 0 <= t < self.n):
            raise ValueError("vertex index out of range (n=%d)" % self.n)
        if self._new_of_old is not None:
            s, t = int(self._new_of_old[s]), int(self._new_of_old[t])
        io = self._scalar_io
        if io is None:
            io = self._scalar_io = (C.c_int32(), C.c_int32(), C.c_double())
        io[0].value, io[1].value = s, t
        _lib.check(lib.drs_query_matrix(self._handle, 1, C.byref(io[0]), C.byref(io[1]), C.byref(io[2])), lib)
        return io[2].value

    _scalar_io = None

    def pair_sums(self, s, t, weights=None):
        """Scalar form of path_sums for one (s, t): (duration, length, hops).  One kernel launch whose answer lands in
        mapped host memory (drs_query_pair); the reference asks one pair at a time (lib/Agents.py:898-923, 756, 1073)."""
        lib = self._ensure(weights)
        if not (0 <= s < self.n and 0 <= t < self.n):
            raise ValueError("vertex index out of range (n=%d)" % self.n)
        if self._new_of_old is not None:
            s, t = int(self._new_of_old[s]), int(self._new_of_old[t])
        out = self._scalar_out
        if out is None:
            out = self._scalar_out = (C.c_double(), C.c_double(), C.c_int32())
        _lib.check(lib.drs_query_pair(self._handle, s, t, C.byref(out[0]), C.byref(out[1]), C.byref(out[2])), lib)
        return out[0].value, out[1].value, out[2].value

    _scalar_out = None

    def epath(self, s, t, weights=None):
        lib = self._ensure(weights)
        if not (0 <= s < self.n and 0 <= t < self.n):
            raise ValueError("vertex index out of range (n=%d)" % self.n)
        if self._new_of_old is not None:
            s, t = int(self._new_of_old[s]), int(self._new_of_old[t])
        buf = self._path_buf
        if buf is None:
            buf = self._path_buf = np.empty(1024, dtype=np.int32)
        while True:
            n = C.c_int32(len(buf))
            reach = C.c_int32(0)
            rc = lib.drs_path(self._handle, int(s), int(t), _lib.ptr_i32(buf), C.byref(n), C.byref(reach))
            if rc == _lib.ERR_CAPACITY:
                buf = self._path_buf = np.empty(max(n.value, 2 * len(buf)), dtype=np.int32)
                continue
            _lib.check(rc, lib)
            return buf[:n.value].tolist()

    _path_buf = None

    def sssp(self, sources, weights=None):
        """Distances from the given sources WITHOUT the N x N matrix (batched delta-stepping SSSP);
        returns a len(sources) x n float64 array.  For networks too large for the matrix
        (BASELINE config C5) and for ``G.shortest_paths(src_list)``-style row queries."""
        lib = self._ensure(weights, build=False)
        src = self.dev_ids([self._resolve(s) for s in sources])
        out = np.empty((len(src), self.n), dtype=np.float64)
        _lib.check(lib.drs_sssp_batch(self._handle, self.mode, len(src), _lib.ptr_i32(src), _lib.ptr_f64(out)), lib)
        return self.host_cols(out)

    def dist_rows(self, row0, nrows, weights=None):
        """Rows row0 .. row0+nrows-1 of the travel-time matrix (caller's numbering, float64, inf = unreachable)."""
        lib = self._ensure(weights)
        if self._new_of_old is None:
            out = np.empty((nrows, self.n), dtype=np.float64)
            _lib.check(lib.drs_apsp_rows_to_host(self._handle, int(row0), int(nrows), _lib.ptr_f64(out)), lib)
            return out
        dev = self.dev_ids(np.arange(row0, row0 + nrows))
        if nrows * 4 >= self.n:                   # most of the matrix: one copy, then reorder on the host
            full = np.empty((self.n, self.n), dtype=np.float64)
            _lib.check(lib.drs_apsp_rows_to_host(self._handle, 0, self.n, _lib.ptr_f64(full)), lib)
            return self.host_cols(full[dev])
        out = np.empty((nrows, self.n), dtype=np.float64)
        for k, r in enumerate(dev):
            _lib.check(lib.drs_apsp_rows_to_host(self._handle, int(r), 1, _lib.ptr_f64(out[k:k + 1])), lib)
        return self.host_cols(out)


class RoutingEngine(object):
    """Same call surface as the reference class (lib/RoutingEngine.py:20-229).

    Attributes (as in the reference): G, cst_speed, rs, ttweight, hour, speeds.
    """

    def __init__(self, graph_loc=None, cst_speed=6, graph=None, ttweight='ttmedian', seed=None, initial_hour=None,
                 speed_table=None):
        if graph_loc is not None:                            # lib/RoutingEngine.py:32-35
            self.G = RoadGraph.Read_GML(graph_loc)
        elif graph is not None:
            self.G = graph if isinstance(graph, RoadGraph) else RoadGraph(graph)
        self.cst_speed = cst_speed
        if seed is None:                                     # :37-39
            seed = np.random.randint(0, 1000000)
            os.makedirs("output", exist_ok=True)
            print('ROUTING ENGINE SEED: {}'.format(seed), file=open('output/seeds.txt', 'a+'))
        self.rs = np.random.RandomState(seed)
        self.ttweight = ttweight
        if initial_hour is not None:                         # :43-48
            self.hour = initial_hour
            self.speeds = speed_table[speed_table["hour"] == self.hour].sample()
        else:
            self.hour = None
            self.speeds = None

    def update_hour(self, new_hour, speed_table=None):
        """lib/RoutingEngine.py:51-53 (the reference method is missing ``self`` and reads an undefined
        global; here the table is passed in)."""
        self.hour = new_hour
        if speed_table is not None:
            self.speeds = speed_table[speed_table["hour"] == self.hour].sample()

    def draw_link_travel_time(self, link_index, use_uncertainty=True):
        """lib/RoutingEngine.py:56-79.  Stays on the host: it consumes ``self.rs`` in edge order."""
        graph_link = self.G.es[link_index]
        if use_uncertainty:
            if self.speeds is not None:
                mu = float(self.speeds["edge_{}_speed_mean".format(int(link_index))])
                sigma = float(self.speeds["edge_{}_speed_stddev".format(int(link_index))])
                speed_draw = np.max([self.rs.normal(mu, sigma), 5])
                return graph_link["length"] / (speed_draw * MPH_2_MPS)
            delta = graph_link["slnshift"]
            mu = graph_link["slnmean"]
            sigma = UNCERTAINTY_MULTIPLIER * graph_link["slnsd"]
            return delta + np.exp(self.rs.normal(mu, sigma))
        return graph_link[self.ttweight]

    def _path(self, olng, olat, dlng, dlat):
        onodename = str((olng, olat))                        # :85-87
        dnodename = str((dlng, dlat))
        o, d = self.G.node_index(onodename), self.G.node_index(dnodename)
        return o, d, self.G.epath(o, d, self.ttweight)

    def get_routing(self, olng, olat, dlng, dlat, pre_drawn_tt=None, use_uncertainty=False):
        """lib/RoutingEngine.py:84-191: route dict with per-edge steps."""
        o, d, path = self._path(olng, olat, dlng, dlat)
        G = self.G
        length = G._eattrs["length"]
        route = dict()
        route['distance'] = sum([float(length[idx]) for idx in path])
        route['steps'] = list()
        route['duration'] = 0
        if pre_drawn_tt is not None:
            if pre_drawn_tt[2] != path:                      # :102-109 path must be reproducible
                print("Path found in routing doesn't match pre-drawn path!!")
                print("ROUTE ALREADY FOUND:\n{}\n".format(pre_drawn_tt[2]))
                print("ROUTE FOUND NOW:\n{}\n".format(path))
                return None
            route['duration'] = pre_drawn_tt[0]
            routing_path = pre_drawn_tt[1]
        else:
            routing_path = path
        cur = o
        lng, lat = G._vattrs["lng"], G._vattrs["lat"]
        for edge_index in routing_path:
            step = dict()
            step['distance'] = float(length[edge_index])
            step['mean_duration'] = self.draw_link_travel_time(edge_index, use_uncertainty=False)
            if pre_drawn_tt is not None:
                step['duration'] = pre_drawn_tt[1][edge_index]
            else:
                drawn = self.draw_link_travel_time(edge_index, use_uncertainty)
                step['duration'] = drawn
                route['duration'] += drawn
            if not use_uncertainty:
                assert step['duration'] == step['mean_duration']
            step['weight'] = step['duration']
            a, b = int(G._src[edge_index]), int(G._dst[edge_index])
            if cur == a:                                     # :146-153 orient undirected edges along the walk
                nxt = b
            else:
                assert cur == b
                nxt = a
            step['intersections'] = [{'location': (lng[cur], lat[cur])}, {'location': (lng[nxt], lat[nxt])}]
            route['steps'].append(step)
            cur = nxt
        route['weight'] = route['duration']
        if len(path) == 0:                                   # :173-189 stay-in-place dummy step
            assert o == d
            loc = (lng[o], lat[o])
            route['steps'].append({'distance': 0.0, 'duration': 0.0, 'mean_duration': 0.0, 'weight': 0.0,
                                   'intersections': [{'location': loc}, {'location': loc}]})
        return route

    def _sums(self, olng, olat, dlng, dlat):
        o = self.G.node_index(str((olng, olat)))
        d = self.G.node_index(str((dlng, dlat)))
        return self.G.pair_sums(o, d, self.ttweight)

    def get_distance(self, olng, olat, dlng, dlat):
        """lib/RoutingEngine.py:195-205: length along the TIME-shortest path (int 0 for an empty path)."""
        _, ln, hops = self._sums(olng, olat, dlng, dlat)
        return ln if hops > 0 else 0

    def get_duration(self, olng, olat, dlng, dlat):
        """lib/RoutingEngine.py:209-218: left-to-right fp64 sum of the path's weights (int 0 for an empty path;
        igraph returns an empty path for an unreachable target).  Where every weight is an integer (and n * max weight
        < 2^24) that sum is exact and equals the matrix entry, so the path need not be walked: one gather."""
        G = self.G
        if G.sums_equal_matrix(self.ttweight):
            o = G.node_index(str((olng, olat)))
            d = G.node_index(str((dlng, dlat)))
            if o == d:
                return 0
            tt = G.pair_time(o, d, self.ttweight)
            return tt if tt != float("inf") else 0
        tt, _, hops = self._sums(olng, olat, dlng, dlat)
        return tt if hops > 0 else 0

    def get_distance_duration(self, olng, olat, dlng, dlat, euclidean=False):
        """lib/RoutingEngine.py:221-226."""
        if euclidean:
            gc_dist = (6371000 * 2 * math.pi / 360 * np.sqrt(
                (math.cos((olat + dlat) * math.pi / 360) * (olng - dlng)) ** 2 + (olat - dlat) ** 2))
            return gc_dist, gc_dist / self.cst_speed
        tt, ln, hops = self._sums(olng, olat, dlng, dlat)
        return (ln, tt) if hops > 0 else (0, 0)

    def generateRandomLocation(self):
        """lib/RoutingEngine.py:228-229: ``rs.choice(G.vs)`` draws exactly one ``randint(0, n)``."""
        return self.G.vs[int(self.rs.randint(0, self.G.n))]

    def offer_queries(self, vehs, trips, num_closest_vehs=3):
        """The routing half of ``Fleet.generate_offer`` (lib/Agents.py:875-923) for a batch of arriving trips in ONE
        device query: per trip the ``num_closest_vehs`` vehicles closest by crow-flies distance among those with a
        free seat (``get_n_closest_vehs``, lib/Agents.py:1208-1223: equirectangular metres, stable argsort, full
        vehicles pushed to 1e10), ``get_duration(vehicle's next node -> origin) + delay`` for each of them
        (lib/Agents.py:898), and ``get_duration`` / ``get_distance`` origin -> destination (lib/Agents.py:904,923).
        The KDE-conditioned pdfs stay on the host.  ``trips`` need ``olng, olat, dlng, dlat``.

        Returns a list with one dict per trip: closest (vehicle objects), veh_dists, direct_wt (list), direct_tt, distance.
        """
        G = self.G
        locs = [veh.get_next_node() for veh in vehs]
        vlng = np.array([l[0][0] for l in locs], dtype=np.float64)
        vlat = np.array([l[0][1] for l in locs], dtype=np.float64)
        free = np.array([veh.n < veh.K for veh in vehs], dtype=bool)
        s, t, meta = [], [], []
        for trip in trips:
            # math.cos per vehicle, like the reference (numpy's vectorised cos may differ in the last bit)
            cosv = np.array([math.cos(x) for x in (vlat + trip.olat) * math.pi / 360])
            gc = 6371000 * 2 * math.pi / 360 * np.sqrt((cosv * (vlng - trip.olng)) ** 2 + (vlat - trip.olat) ** 2)
            dists = np.where(free, gc, 1e10)
            order = np.argsort(dists)[:num_closest_vehs]          # np.argsort: same (quicksort) order as the reference
            o = G.node_of_location(trip.olng, trip.olat)
            d = G.node_of_location(trip.dlng, trip.dlat)
            for k in order:
                s.append(G.node_of_location(vlng[k], vlat[k])); t.append(o)
            s.append(o); t.append(d)
            meta.append((order, dists[order]))
        tt, ln, hops = G.path_sums(s, t, self.ttweight) if s else (np.zeros(0), np.zeros(0), np.zeros(0, dtype=np.int32))
        out, pos = [], 0
        for (order, vd) in meta:
            k = len(order)
            wt = [(float(tt[pos + i]) if hops[pos + i] > 0 else 0) + locs[order[i]][1] for i in range(k)]
            od = pos + k
            out.append({"closest": [vehs[i] for i in order], "veh_dists": [float(x) for x in vd], "direct_wt": wt,
                        "direct_tt": float(tt[od]) if hops[od] > 0 else 0, "distance": float(ln[od]) if hops[od] > 0 else 0})
            pos += k + 1
        return out

    # ---- batched extensions (not in the reference; same results as calling the scalar methods in a loop)
    def get_durations(self, pairs):
        """pairs: iterable of (olng, olat, dlng, dlat) -> np.ndarray of durations."""
        o = [self.G.node_index(str((p[0], p[1]))) for p in pairs]
        d = [self.G.node_index(str((p[2], p[3]))) for p in pairs]
        tt, _, hops = self.G.path_sums(o, d, self.ttweight)
        return np.where(hops > 0, tt, 0.0)
