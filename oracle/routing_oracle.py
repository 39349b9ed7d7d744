"""CPU oracle for the routing half of the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (drs_abm_b200/) never does.

What it restates (paths relative to the nbailey/drs-abm checkout):
  * lib/RoutingEngine.py:20-229  -- class RoutingEngine (OracleRoutingEngine)
  * the python-igraph calls it makes (igraph is a third-party dependency that is
    not vendored and not installed here; version unpinned by the reference --
    SURVEY.md section 8c): Graph.Read_GML, vs.find, es[i],
    get_shortest_paths(..., output='epath'), shortest_paths(...).
    igraph's Dijkstra is restated with scipy.sparse.csgraph.dijkstra (C) for
    distances and with the canonical-predecessor rule below for the path.

Parity pin: the reference has no golden vectors or tests for this path
(SURVEY.md section 4).  This oracle is pinned against (a) the reference's own
lib/RoutingEngine.py executed in the authoring container on an igraph facade
(oracle/ref_shim, fixtures under tests/golden/ made by oracle/make_golden.py) and
(b) networkx Dijkstra distances.  Which of several equal-cost paths igraph
returns is unobservable here: "parity unpinned" for tie-breaking, which this
oracle DEFINES:

Canonical predecessor rule (algorithm independent): the last edge of the path
s -> t is the in-edge (u -> t, w, eid) minimising, lexicographically,
    (dist[s,u] + w,  dist[s,u],  eid).
For exact (integer) weights every tight edge has the same first key, so the rule
picks the tight in-neighbour closest to s -- the neighbour a Dijkstra settles
first, which is the parent igraph/networkx keep ("replace only on strictly
smaller distance") -- and breaks remaining ties by smallest edge id.
"""
import math

import numpy as np
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import dijkstra

MPH_2_MPS = 0.44704            # lib/Constants.py:152
UNCERTAINTY_MULTIPLIER = 1.5   # lib/Constants.py:99


class OracleGraph(object):
    """Plain-array road graph: what ig.Graph.Read_GML yields for the attributes the path uses."""

    def __init__(self, n, src, dst, tt, length, names, lng, lat, directed=False, extra_eattrs=None):
        self.n = int(n)
        self.src = np.asarray(src, dtype=np.int64)
        self.dst = np.asarray(dst, dtype=np.int64)
        self.m = len(self.src)
        self.tt = np.asarray(tt, dtype=np.float64).copy()
        self.length = np.asarray(length, dtype=np.float64)
        self.names = list(names)
        self.lng = list(lng)
        self.lat = list(lat)
        self.directed = bool(directed)
        self.extra = extra_eattrs or {}
        self.name_index = {nm: i for i, nm in enumerate(self.names)}
        # in-arcs per target, in edge-id order
        self.in_arcs = [[] for _ in range(self.n)]
        for e in range(self.m):
            u, v = int(self.src[e]), int(self.dst[e])
            self.in_arcs[v].append((u, e))
            if not self.directed and u != v:
                self.in_arcs[u].append((v, e))
        self._dist = None

    @classmethod
    def from_gml(cls, path):
        """Independent GML reader (networkx) used to cross-check the product's loader."""
        import networkx as nx
        G = nx.read_gml(path, label="id")
        nodes = list(G.nodes())
        idx = {k: i for i, k in enumerate(nodes)}
        src, dst, tt, ln = [], [], [], []
        for u, v, d in G.edges(data=True):
            src.append(idx[u]); dst.append(idx[v]); tt.append(float(d["ttmedian"])); ln.append(float(d["length"]))
        names = [G.nodes[k]["name"] for k in nodes]
        lng = [float(G.nodes[k]["lng"]) for k in nodes]
        lat = [float(G.nodes[k]["lat"]) for k in nodes]
        return cls(len(nodes), src, dst, tt, ln, names, lng, lat, directed=G.is_directed())

    # ---------------------------------------------------------------- distances
    def adjacency(self):
        if self.directed:
            r, c, w = self.src, self.dst, self.tt
        else:
            r = np.concatenate([self.src, self.dst]); c = np.concatenate([self.dst, self.src])
            w = np.concatenate([self.tt, self.tt])
        # multi-edges: keep the minimum weight (csr_matrix would SUM duplicates)
        order = np.lexsort((w, c, r))
        r, c, w = r[order], c[order], w[order]
        first = np.ones(len(r), dtype=bool)
        first[1:] = (r[1:] != r[:-1]) | (c[1:] != c[:-1])
        return csr_matrix((w[first], (r[first], c[first])), shape=(self.n, self.n))

    def invalidate(self):
        self._dist = None

    def dist_rows(self, sources):
        """fp64 Dijkstra distances from the given sources (igraph shortest_paths)."""
        return dijkstra(self.adjacency(), directed=True, indices=np.asarray(sources, dtype=np.int64))

    def dist_matrix(self):
        if self._dist is None:
            self._dist = dijkstra(self.adjacency(), directed=True)
        return self._dist

    # ------------------------------------------------------------ canonical path
    def pred_edge(self, drow, s, t, dtype=np.float64):
        """Canonical last edge of the path s->t given the distance row of s."""
        if s == t or not np.isfinite(drow[t]):
            return -1
        best = None
        for (u, e) in self.in_arcs[t]:
            du = dtype(drow[u])
            if not np.isfinite(du):
                continue
            key = (dtype(du + dtype(self.tt[e])), du, e)
            if best is None or key < best:
                best = key
        return -1 if best is None else best[2]

    def epath(self, s, t, drow=None, dtype=np.float64):
        """Edge ids of the canonical shortest path s->t in travel order ([] if s==t or unreachable)."""
        if drow is None:
            drow = self.dist_matrix()[s] if self._dist is not None else self.dist_rows([s])[0]
        if s == t or not np.isfinite(drow[t]):
            return []
        path = []
        cur = t
        while cur != s:
            e = self.pred_edge(drow, s, cur, dtype)
            assert e >= 0 and len(path) <= self.n
            path.append(e)
            a, b = int(self.src[e]), int(self.dst[e])
            cur = a if b == cur else b
        path.reverse()
        return path


class _Vertex(dict):
    """Stands in for an igraph Vertex: v['lng'], v['lat'], v['name'], v.index."""

    def __init__(self, index, name, lng, lat):
        dict.__init__(self, name=name, lng=lng, lat=lat)
        self.index = index


class OracleRoutingEngine(object):
    """Restatement of lib/RoutingEngine.py:20-229 over OracleGraph.

    Same constructor defaults and method results as the reference class; the
    igraph calls are replaced by OracleGraph.epath / dist_rows.
    """

    def __init__(self, graph, cst_speed=6, ttweight="ttmedian", seed=0, speeds=None):
        self.G = graph                      # lib/RoutingEngine.py:33-35
        self.cst_speed = cst_speed          # :36
        self.rs = np.random.RandomState(seed)   # :40
        self.ttweight = ttweight            # :41
        self.hour = None                    # :47-48
        self.speeds = speeds

    def _node(self, lng, lat):
        name = str((lng, lat))              # :85-86 -- vertex names are str((lng, lat))
        if name not in self.G.name_index:
            raise ValueError("no such vertex: %r" % name)   # igraph vs.find raises ValueError
        return self.G.name_index[name]

    def draw_link_travel_time(self, link_index, use_uncertainty=True):
        """lib/RoutingEngine.py:56-79."""
        if use_uncertainty:
            if self.speeds is not None:     # :61-68 hourly speed table branch
                mu = float(self.speeds["edge_{}_speed_mean".format(int(link_index))])
                sigma = float(self.speeds["edge_{}_speed_stddev".format(int(link_index))])
                speed_draw = np.max([self.rs.normal(mu, sigma), 5])
                return self.G.length[link_index] / (speed_draw * MPH_2_MPS)
            delta = self.G.extra["slnshift"][link_index]          # :71-75 shifted log-normal
            mu = self.G.extra["slnmean"][link_index]
            sigma = UNCERTAINTY_MULTIPLIER * self.G.extra["slnsd"][link_index]
            return delta + np.exp(self.rs.normal(mu, sigma))
        return float(self.G.tt[link_index])                       # :79

    def get_routing(self, olng, olat, dlng, dlat, pre_drawn_tt=None, use_uncertainty=False):
        """lib/RoutingEngine.py:84-191."""
        o, d = self._node(olng, olat), self._node(dlng, dlat)
        path = self.G.epath(o, d)                                  # :87
        route = {"distance": sum([float(self.G.length[e]) for e in path]), "steps": [], "duration": 0}  # :92-94
        if pre_drawn_tt is not None:                               # :96-112
            if pre_drawn_tt[2] != path:
                return None
            route["duration"] = pre_drawn_tt[0]
            routing_path = pre_drawn_tt[1]
        else:
            routing_path = path
        cur = o
        for e in routing_path:                                     # :120-166
            step = {"distance": float(self.G.length[e]),
                    "mean_duration": self.draw_link_travel_time(e, use_uncertainty=False)}
            if pre_drawn_tt is not None:
                step["duration"] = pre_drawn_tt[1][e]
            else:
                drawn = self.draw_link_travel_time(e, use_uncertainty)
                step["duration"] = drawn
                route["duration"] += drawn
            if not use_uncertainty:
                assert step["duration"] == step["mean_duration"]   # :134-135
            step["weight"] = step["duration"]
            a, b = int(self.G.src[e]), int(self.G.dst[e])
            if cur == a:                                           # :146-153 orientation of undirected edges
                nxt = b
            else:
                assert cur == b
                nxt = a
            step["intersections"] = [{"location": (self.G.lng[cur], self.G.lat[cur])},
                                     {"location": (self.G.lng[nxt], self.G.lat[nxt])}]
            route["steps"].append(step)
            cur = nxt
        route["weight"] = route["duration"]                        # :170
        if len(path) == 0:                                         # :173-189 zero-length dummy step
            assert o == d
            loc = (self.G.lng[o], self.G.lat[o])
            route["steps"].append({"distance": 0.0, "duration": 0.0, "mean_duration": 0.0, "weight": 0.0,
                                   "intersections": [{"location": loc}, {"location": loc}]})
        return route

    def get_distance(self, olng, olat, dlng, dlat):
        """lib/RoutingEngine.py:195-205: length along the TIME-shortest path."""
        total = 0
        for e in self.G.epath(self._node(olng, olat), self._node(dlng, dlat)):
            total += float(self.G.length[e])
        return total

    def get_duration(self, olng, olat, dlng, dlat):
        """lib/RoutingEngine.py:209-218: left-to-right fp64 sum of the path's edge weights."""
        duration = 0
        for e in self.G.epath(self._node(olng, olat), self._node(dlng, dlat)):
            duration += float(self.G.tt[e])
        return duration

    def get_distance_duration(self, olng, olat, dlng, dlat, euclidean=False):
        """lib/RoutingEngine.py:221-226."""
        if euclidean:
            gc = (6371000 * 2 * math.pi / 360 * np.sqrt(
                (math.cos((olat + dlat) * math.pi / 360) * (olng - dlng)) ** 2 + (olat - dlat) ** 2))
            return gc, gc / self.cst_speed
        return self.get_distance(olng, olat, dlng, dlat), self.get_duration(olng, olat, dlng, dlat)

    def generateRandomLocation(self):
        """lib/RoutingEngine.py:228-229: rs.choice(G.vs) == G.vs[rs.randint(0, n)]."""
        i = int(self.rs.randint(0, self.G.n))
        return _Vertex(i, self.G.names[i], self.G.lng[i], self.G.lat[i])
