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
                       data.vattrs["name"], data.vattrs.get("lng", [0.0] * data.n), data.vattrs.get("lat", [0.0] * data.n),
                       directed=data.directed,
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


@pytest.fixture(params=["split_default", "split_deep", "split_none"])
def eval_split(request, monkeypatch):
    """Runs a dispatch test under three settings of how the trip-evaluation kernels split a trip's ordering search
    into parallel searches (by the first 1, 2 or 3 stops): the default, as deep as possible for every trip, and no
    deeper than the first stop.  The split is a scheduling decision and must never change a result."""
    s2, s3 = {"split_default": (None, None), "split_deep": ("2", "3"), "split_none": ("99", "99")}[request.param]
    for name, val in (("DRS_EVAL_SPLIT2", s2), ("DRS_EVAL_SPLIT3", s3)):
        if val is None:
            monkeypatch.delenv(name, raising=False)
        else:
            monkeypatch.setenv(name, val)
    return request.param
