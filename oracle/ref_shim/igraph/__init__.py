"""Minimal python-igraph facade over networkx.  TEST INFRASTRUCTURE ONLY.

python-igraph is a third-party dependency of the reference that is neither
vendored nor installed.  This facade implements just the
subset of its API that lib/RoutingEngine.py, lib/optimUtils.py and
lib/Optimization.py (AlonsoMora) touch, so that those reference files can be
imported and executed UNMODIFIED in the authoring container by
oracle/make_golden.py to produce the fixtures under tests/golden/.

Shortest paths are computed by networkx's Dijkstra (an implementation
independent of both the oracle's scipy path and the CUDA kernels).  Which of
several equal-cost paths is returned follows networkx, not igraph; the golden
generator therefore only records path-dependent outputs where the shortest path
is unique.
"""
import networkx as nx

ALL = 3
OUT = 1
IN = 2


class Vertex(object):
    def __init__(self, graph, index):
        self.graph = graph
        self.index = index

    def __getitem__(self, key):
        return self.graph._vattrs[self.index].get(key)

    def __setitem__(self, key, value):
        self.graph._vattrs[self.index][key] = value

    def attributes(self):
        return dict(self.graph._vattrs[self.index])

    def neighbors(self, mode=ALL):
        return [Vertex(self.graph, i) for i in self.graph._neighbor_ids(self.index)]

    def __eq__(self, other):
        return isinstance(other, Vertex) and other.graph is self.graph and other.index == self.index

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((id(self.graph), self.index))


class Edge(object):
    def __init__(self, graph, index):
        self.graph = graph
        self.index = index

    @property
    def source(self):
        return self.graph._edges[self.index][0]

    @property
    def target(self):
        return self.graph._edges[self.index][1]

    @property
    def tuple(self):
        return self.graph._edges[self.index]

    def __getitem__(self, key):
        return self.graph._eattrs[self.index].get(key)

    def __setitem__(self, key, value):
        self.graph._eattrs[self.index][key] = value
        self.graph._nx = None


def _match(attrs, conds):
    for k, v in conds.items():
        if k.endswith("_in"):
            if attrs.get(k[:-3]) not in v:
                return False
        elif attrs.get(k) != v:
            return False
    return True


class VertexSeq(object):
    def __init__(self, graph, indices=None):
        self.graph = graph
        self._idx = indices

    def indices_list(self):
        return list(range(len(self.graph._vattrs))) if self._idx is None else list(self._idx)

    def __len__(self):
        return len(self.indices_list())

    def __iter__(self):
        return iter([Vertex(self.graph, i) for i in self.indices_list()])

    def __getitem__(self, key):
        if isinstance(key, str):
            return [self.graph._vattrs[i].get(key) for i in self.indices_list()]
        return Vertex(self.graph, self.indices_list()[int(key)])

    def find(self, *args, **kwds):
        if args:
            if isinstance(args[0], str):
                kwds = dict(kwds, name=args[0])
            else:
                return Vertex(self.graph, self.indices_list()[int(args[0])])
        if list(kwds.keys()) == ["name"] and self._idx is None:
            i = self.graph._name_index().get(kwds["name"])
            if i is None:
                raise ValueError("no such vertex: %r" % (kwds["name"],))
            return Vertex(self.graph, i)
        for i in self.indices_list():
            if _match(self.graph._vattrs[i], kwds):
                return Vertex(self.graph, i)
        raise ValueError("no such vertex")

    def select(self, **kwds):
        return VertexSeq(self.graph, [i for i in self.indices_list() if _match(self.graph._vattrs[i], kwds)])

    def subgraph(self):
        keep = self.indices_list()
        pos = {v: k for k, v in enumerate(keep)}
        g = Graph(directed=self.graph._directed)
        for v in keep:
            g._vattrs.append(dict(self.graph._vattrs[v]))
        for e, (a, b) in enumerate(self.graph._edges):
            if a in pos and b in pos:
                g._edges.append((pos[a], pos[b]))
                g._eattrs.append(dict(self.graph._eattrs[e]))
        return g


class EdgeSeq(object):
    def __init__(self, graph, indices=None):
        self.graph = graph
        self._idx = indices

    def indices_list(self):
        return list(range(len(self.graph._edges))) if self._idx is None else list(self._idx)

    def __len__(self):
        return len(self.indices_list())

    def __iter__(self):
        return iter([Edge(self.graph, i) for i in self.indices_list()])

    def __getitem__(self, key):
        if isinstance(key, str):
            return [self.graph._eattrs[i].get(key) for i in self.indices_list()]
        return Edge(self.graph, self.indices_list()[int(key)])

    def __setitem__(self, key, values):
        idx = self.indices_list()
        if not hasattr(values, "__len__"):
            values = [values] * len(idx)
        for i, v in zip(idx, values):
            self.graph._eattrs[i][key] = v
        self.graph._nx = None

    def _conds_ok(self, i, kwds):
        a, b = self.graph._edges[i]
        for k, v in kwds.items():
            if k == "_source":
                v = v.index if isinstance(v, Vertex) else v
                if a != v:
                    return False
            elif k == "_target":
                v = v.index if isinstance(v, Vertex) else v
                if b != v:
                    return False
            elif not _match(self.graph._eattrs[i], {k: v}):
                return False
        return True

    def find(self, **kwds):
        for i in self.indices_list():
            if self._conds_ok(i, kwds):
                return Edge(self.graph, i)
        raise ValueError("no such edge")

    def select(self, **kwds):
        return EdgeSeq(self.graph, [i for i in self.indices_list() if self._conds_ok(i, kwds)])


class Graph(object):
    def __init__(self, directed=False):
        self._directed = directed
        self._vattrs = []
        self._edges = []
        self._eattrs = []
        self._nx = None
        self._sssp = {}
        self._names = None

    # ------------------------------------------------------------ construction
    @classmethod
    def Read_GML(cls, path):
        G = nx.read_gml(path, label="id")
        g = cls(directed=G.is_directed())
        order = list(G.nodes())
        pos = {k: i for i, k in enumerate(order)}
        for k in order:
            attrs = {a: (float(v) if isinstance(v, (int, float)) else v) for a, v in G.nodes[k].items()}
            attrs["id"] = float(k)
            g._vattrs.append(attrs)
        for u, v, d in G.edges(data=True):
            g._edges.append((pos[u], pos[v]))
            g._eattrs.append({a: (float(x) if isinstance(x, (int, float)) else x) for a, x in d.items()})
        return g

    def add_vertex(self, name=None, **kwds):
        attrs = dict(kwds)
        if name is not None:
            attrs["name"] = name
        self._vattrs.append(attrs)
        self._names = None

    def _vid(self, v):
        if isinstance(v, Vertex):
            return v.index
        if isinstance(v, str):
            i = self._name_index().get(v)
            if i is None:
                raise ValueError("no such vertex: %r" % (v,))
            return i
        return int(v)

    def add_edge(self, source, target, **kwds):
        self._edges.append((self._vid(source), self._vid(target)))
        self._eattrs.append(dict(kwds))
        self._nx = None

    def copy(self):
        g = Graph(directed=self._directed)
        g._vattrs = [dict(a) for a in self._vattrs]
        g._edges = list(self._edges)
        g._eattrs = [dict(a) for a in self._eattrs]
        return g

    def delete_edges(self, edges=None):
        self._edges = []
        self._eattrs = []
        self._nx = None

    # ------------------------------------------------------------------ access
    @property
    def vs(self):
        return VertexSeq(self)

    @property
    def es(self):
        return EdgeSeq(self)

    def _name_index(self):
        if self._names is None or len(self._names) != len(self._vattrs):
            self._names = {a.get("name"): i for i, a in reversed(list(enumerate(self._vattrs)))}
        return self._names

    def _neighbor_ids(self, v):
        out = []
        for (a, b) in self._edges:
            if a == v:
                out.append(b)
            elif b == v and not self._directed:
                out.append(a)
        return sorted(out)

    def neighbors(self, v, mode=ALL):
        return self._neighbor_ids(self._vid(v))

    def incident(self, v, mode=OUT):
        v = self._vid(v)
        return [e for e, (a, b) in enumerate(self._edges) if a == v or b == v]

    # ---------------------------------------------------------- shortest paths
    def _nxgraph(self, weights):
        if self._nx is None or self._nx[0] != weights:
            G = nx.MultiDiGraph() if self._directed else nx.MultiGraph()
            G.add_nodes_from(range(len(self._vattrs)))
            for e, (a, b) in enumerate(self._edges):
                G.add_edge(a, b, key=e, w=self._eattrs[e][weights])
            self._nx = (weights, G)
            self._sssp = {}
        return self._nx[1]

    def _dist_from(self, G, s):
        # memoised single-source Dijkstra: the reference recomputes it on every call; the
        # values are identical, the golden generator just finishes in minutes instead of hours
        d = self._sssp.get(s)
        if d is None:
            d = nx.single_source_dijkstra_path_length(G, s, weight="w")
            self._sssp[s] = d
        return d

    def get_shortest_paths(self, v, to=None, weights=None, output="vpath"):
        G = self._nxgraph(weights)
        s, t = self._vid(v), self._vid(to)
        try:
            vpath = nx.dijkstra_path(G, s, t, weight="w")
        except nx.NetworkXNoPath:
            return [[]]
        if output == "vpath":
            return [vpath]
        epath = []
        for a, b in zip(vpath[:-1], vpath[1:]):
            best = min(G[a][b].items(), key=lambda kv: (kv[1]["w"], kv[0]))
            epath.append(best[0])
        return [epath]

    def shortest_paths(self, source=None, target=None, weights=None):
        G = self._nxgraph(weights)
        srcs = source if isinstance(source, (list, tuple)) else [source]
        tgts = target if isinstance(target, (list, tuple)) else [target]
        out = []
        for s in srcs:
            dist = self._dist_from(G, self._vid(s))
            out.append([dist.get(self._vid(t), float("inf")) for t in tgts])
        return out
