import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def unjf(x):
    """Inverse of make_golden.jf: JSON-safe float back to float."""
    return float("inf") if x == "inf" else x


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as fh:
        return json.load(fh)


def golden_map_path(name):
    return os.path.join(GOLDEN, name + ".gml")


_oracle_graphs = {}


def oracle_graph(name):
    """OracleGraph of a golden map, read by the oracle's own (networkx) GML reader."""
    from oracle.routing_oracle import OracleGraph
    if name not in _oracle_graphs:
        _oracle_graphs[name] = OracleGraph.from_gml(golden_map_path(name))
    return _oracle_graphs[name]


def oracle_graph_from_data(data):
    from oracle.routing_oracle import OracleGraph
    extra = {k: np.asarray(data.eattrs[k]) for k in ("slnshift", "slnmean", "slnsd") if k in data.eattrs}
    return OracleGraph(data.n, data.src, data.dst, data.eattrs["ttmedian"], data.eattrs["length"],
                       data.vattrs["name"], data.vattrs["lng"], data.vattrs["lat"], directed=data.directed,
                       extra_eattrs=extra)


class OReq(object):
    """Request object with the attribute surface the optimizers read (lib/Optimization.py:107-118,
    lib/optimUtils.py:95-108)."""

    def __init__(self, rid, olng, olat, dlng, dlat, Cep, Clp, Ced, Cld, assigned=False):
        self.id, self.olng, self.olat, self.dlng, self.dlat = rid, olng, olat, dlng, dlat
        self.Cep, self.Clp, self.Ced, self.Cld = Cep, Clp, Ced, Cld
        self.assigned = assigned

    def get_dropoff_time_window(self):
        return [self.Ced, self.Cld]


class OVeh(object):
    """Vehicle object with the attribute surface the optimizers read (lib/Optimization.py:121-129)."""

    def __init__(self, vid, K, loc, t, pax, wait):
        self.id, self.K = vid, K
        self._loc, self._t, self._pax, self._wait = loc, t, set(pax), set(wait)

    def get_passenger_requests(self):
        return set(self._pax), set(self._wait)

    def get_next_node(self):
        return self._loc, self._t
