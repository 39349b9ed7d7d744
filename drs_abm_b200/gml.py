"""GML road-network reader/writer (the on-disk format either side of the path).

Replaces ``ig.Graph.Read_GML(graph_loc)`` (lib/RoutingEngine.py:33).  Only the
subset igraph writes is handled: ``graph [ directed D node [ id .. k v .. ] edge
[ source .. target .. k v .. ] ]``.  Like igraph, numeric attributes come back as
Python floats and strings as str; vertex order is file order and ``source`` /
``target`` refer to node ``id`` values.
"""
import re

import array

import numpy as np

_TOKEN = re.compile(r'"([^"]*)"|(\[)|(\])|([^\s\[\]"]+)')


class RoadGraphData(object):
    """Plain-array form of a road network.

    n, m, directed; ``vattrs``: {name: list}, ``eattrs``: {name: np.ndarray or
    list}; ``src`` / ``dst``: int32 arrays of vertex indices (file order).
    """

    def __init__(self, n, directed, src, dst, vattrs, eattrs):
        self.n = int(n)
        self.directed = bool(directed)
        self.src = np.ascontiguousarray(src, dtype=np.int32)
        self.dst = np.ascontiguousarray(dst, dtype=np.int32)
        self.m = int(self.src.shape[0])
        self.vattrs = vattrs
        self.eattrs = eattrs
        # numeric (lng, lat) arrays beside the attribute lists, converted once at load time: the device numbering of every
        # RoadGraph made from this map reads them (None when the map has no usable coordinates)
        self.coords = None
        try:
            x = np.frombuffer(array.array("d", vattrs["lng"]), dtype=np.float64)
            y = np.frombuffer(array.array("d", vattrs["lat"]), dtype=np.float64)
            if x.shape == (self.n,) and y.shape == (self.n,):
                self.coords = (x, y)
        except (KeyError, TypeError, ValueError):
            pass


def _convert(tok_str, tok_num):
    if tok_str is not None:
        return tok_str
    try:
        return float(tok_num)
    except ValueError:
        return tok_num


def read_gml(path):
    with open(path, "r") as fh:
        text = fh.read()
    tokens = _TOKEN.findall(text)
    # tokens: tuples (string, '[', ']', bare)
    pos = 0
    ntok = len(tokens)

    def parse_list(pos):
        """Parses key/value pairs until the matching ']' ; returns (list_of_pairs, newpos)."""
        items = []
        while pos < ntok:
            s, lb, rb, bare = tokens[pos]
            if rb:
                return items, pos + 1
            key = bare
            pos += 1
            s, lb, rb, bare = tokens[pos]
            if lb:
                val, pos = parse_list(pos + 1)
            else:
                val = s if (s or not bare) else None
                if val is None:
                    try:
                        val = float(bare)
                    except ValueError:
                        val = bare
                pos += 1
            items.append((key, val))
        return items, pos

    top, _ = parse_list(0)
    graph = None
    for k, v in top:
        if k == "graph":
            graph = v
    if graph is None:
        raise ValueError("no 'graph [' block in %s" % path)
    directed = False
    nodes, edges = [], []
    for k, v in graph:
        if k == "directed":
            directed = bool(int(v))
        elif k == "node":
            nodes.append(dict(v))
        elif k == "edge":
            edges.append(dict(v))
    id2idx = {}
    for i, nd in enumerate(nodes):
        id2idx[int(nd.get("id", i))] = i
    vkeys = []
    for nd in nodes:
        for k in nd:
            if k not in vkeys:
                vkeys.append(k)
    vattrs = {k: [nd.get(k) for nd in nodes] for k in vkeys}
    ekeys = []
    for ed in edges:
        for k in ed:
            if k not in ekeys and k not in ("source", "target"):
                ekeys.append(k)
    src = np.array([id2idx[int(ed["source"])] for ed in edges], dtype=np.int32)
    dst = np.array([id2idx[int(ed["target"])] for ed in edges], dtype=np.int32)
    eattrs = {}
    for k in ekeys:
        col = [ed.get(k) for ed in edges]
        if all(isinstance(x, float) for x in col):
            col = np.array(col, dtype=np.float64)
        eattrs[k] = col
    return RoadGraphData(len(nodes), directed, src, dst, vattrs, eattrs)


def _fmt(v):
    if isinstance(v, str):
        return '"%s"' % v
    if isinstance(v, (int, np.integer)):
        return str(int(v))
    return repr(float(v))


def write_gml(path, data):
    """Writes igraph-style GML; vertex ``id`` = index."""
    with open(path, "w") as fh:
        fh.write("Creator \"drs_abm_b200\"\ngraph\n[\n  directed %d\n" % (1 if data.directed else 0))
        vkeys = [k for k in data.vattrs if k != "id"]
        for i in range(data.n):
            fh.write("  node\n  [\n    id %d\n" % i)
            for k in vkeys:
                fh.write("    %s %s\n" % (k, _fmt(data.vattrs[k][i])))
            fh.write("  ]\n")
        ekeys = list(data.eattrs)
        for e in range(data.m):
            fh.write("  edge\n  [\n    source %d\n    target %d\n" % (data.src[e], data.dst[e]))
            for k in ekeys:
                fh.write("    %s %s\n" % (k, _fmt(data.eattrs[k][e])))
            fh.write("  ]\n")
        fh.write("]\n")
