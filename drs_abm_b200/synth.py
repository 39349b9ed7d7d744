"""Synthetic road networks, fleets and demand.

The reference ships no input data (maps, speeds, requests and vehicles are all
git-ignored: /root/reference/.gitignore:7-11), so every configuration in
BASELINE.json is generated here with fixed seeds.  File formats follow the
reference's readers: GML with vertex attrs name/lng/lat and edge attrs
length/ttmedian/slnshift/slnmean/slnsd (lib/RoutingEngine.py:71-79,92,140,156);
vehicles CSV ``vid,vlng,vlat`` and requests CSV
``rid,olng,olat,dlng,dlat,Cep,Clp,Ced,Cld`` (optimizationTesting.py:24-42);
demand CSV ``olat,olng,dlat,dlng,count,pt_av,car_av`` (lib/ModeChoiceAlt.py:111-139).
"""
import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import connected_components

from .gml import RoadGraphData

LNG0, LAT0, STEP = -71.10, 42.30, 0.001

# name -> (n_side, keep, seed, weights)
CONFIGS = {
    "C1": (20, 1.0, 20, "int"),
    "C1_float": (20, 1.0, 20, "float"),
    "C1_uniform": (20, 1.0, 20, "uniform"),
    "C2": (71, 0.92, 5000, "int"),
    "C2_float": (71, 0.92, 5000, "float"),
    "C3": (224, 0.92, 50000, "int"),
    "C5": (448, 0.90, 200000, "int"),
}


def node_coords(i, j):
    return round(LNG0 + i * STEP, 6), round(LAT0 + j * STEP, 6)


def grid_city(n_side, seed, keep=1.0, weights="int"):
    """Undirected n_side x n_side grid; a fraction 1-keep of the edges is removed
    and only the largest connected component is kept.

    weights: "int"   ttmedian = randint(20, 61) s, length = 10 * ttmedian m
             "float" length = 100 * lognormal(0, 0.25) m, ttmedian = length / 9.0 s
             "uniform" every ttmedian = 30 s, length = 300 m (tie-stress case)
    """
    rs = np.random.RandomState(seed)
    idx = np.arange(n_side * n_side, dtype=np.int64).reshape(n_side, n_side)  # k = i*n_side + j
    right = np.stack([idx[:, :-1].ravel(), idx[:, 1:].ravel()], 1)   # (i,j)-(i,j+1)
    down = np.stack([idx[:-1, :].ravel(), idx[1:, :].ravel()], 1)    # (i,j)-(i+1,j)
    # edges "right then down per node": order by source node, right before down
    allE = np.concatenate([right, down], 0)
    kind = np.concatenate([np.zeros(len(right), np.int64), np.ones(len(down), np.int64)])
    order = np.lexsort((kind, allE[:, 0]))
    allE = allE[order]
    m0 = allE.shape[0]
    if keep < 1.0:
        mask = rs.rand(m0) < keep
        allE = allE[mask]
    n0 = n_side * n_side
    adj = coo_matrix((np.ones(len(allE)), (allE[:, 0], allE[:, 1])), shape=(n0, n0))
    ncomp, labels = connected_components(adj, directed=False)
    if ncomp > 1:
        big = np.argmax(np.bincount(labels))
        keep_nodes = labels == big
        emask = keep_nodes[allE[:, 0]] & keep_nodes[allE[:, 1]]
        allE = allE[emask]
        new_id = -np.ones(n0, np.int64)
        new_id[keep_nodes] = np.arange(keep_nodes.sum())
        old_ids = np.nonzero(keep_nodes)[0]
        allE = new_id[allE]
    else:
        old_ids = np.arange(n0)
    n = len(old_ids)
    m = allE.shape[0]
    if weights == "int":
        tt = rs.randint(20, 61, size=m).astype(np.float64)
        length = 10.0 * tt
    elif weights == "float":
        length = 100.0 * rs.lognormal(0.0, 0.25, size=m)
        tt = length / 9.0
    elif weights == "uniform":
        tt = np.full(m, 30.0)
        length = np.full(m, 300.0)
    else:
        raise ValueError("unknown weights %r" % (weights,))
    ii, jj = old_ids // n_side, old_ids % n_side
    lng = [node_coords(int(a), int(b))[0] for a, b in zip(ii, jj)]
    lat = [node_coords(int(a), int(b))[1] for a, b in zip(ii, jj)]
    names = [str((a, b)) for a, b in zip(lng, lat)]
    vattrs = {"name": names, "lng": lng, "lat": lat}
    # shifted-lognormal link travel-time parameters (lib/RoutingEngine.py:71-75): median = ttmedian
    slnshift = 0.5 * tt
    slnmean = np.log(0.5 * tt)
    slnsd = np.full(m, 0.2)
    eattrs = {"length": length, "ttmedian": tt, "slnshift": slnshift, "slnmean": slnmean, "slnsd": slnsd}
    return RoadGraphData(n, False, allE[:, 0], allE[:, 1], vattrs, eattrs)


def config_graph(name):
    n_side, keep, seed, weights = CONFIGS[name]
    return grid_city(n_side, seed, keep, weights)


def random_od_pairs(n, count, seed):
    rs = np.random.RandomState(seed)
    o = rs.randint(0, n, size=count)
    d = rs.randint(0, n, size=count)
    same = o == d
    d[same] = (d[same] + 1 + rs.randint(0, n - 1, size=same.sum())) % n
    return o.astype(np.int32), d.astype(np.int32)
