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
             "dfloat" like "float" but DIRECTED: two one-way arcs per street with independent lengths
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
    elif weights == "dfloat":
        # DIRECTED float map: every street is two one-way arcs with independent lengths (a -> b, then b -> a)
        allE = np.concatenate([allE, allE[:, ::-1]], 0)
        allE = allE[np.argsort(allE[:, 0], kind="stable")]   # arcs grouped by their tail vertex, like the other maps' edges
        m = allE.shape[0]
        length = 100.0 * rs.lognormal(0.0, 0.25, size=m)
        tt = length / 9.0
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
    return RoadGraphData(n, weights == "dfloat", allE[:, 0], allE[:, 1], vattrs, eattrs)


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


# --------------------------------------------------------------------------- fleet state
# Plan templates by number of planned stops L: A = on-board request (drop-off only),
# W = waiting request (pickup then drop-off).  Occupancy never exceeds 4.
_PLAN_TEMPLATES = {
    0: [],
    2: [("W0", 1), ("W0", -1)],
    4: [("W0", 1), ("W1", 1), ("W0", -1), ("W1", -1)],
    6: [("A0", -1), ("W0", 1), ("A1", -1), ("W0", -1), ("W1", 1), ("W1", -1)],
    8: [("A0", -1), ("A1", -1), ("W0", 1), ("A2", -1), ("W0", -1), ("A3", -1), ("W1", 1), ("W1", -1)],
}
MAX_WAIT = 600.0      # lib/Constants.py:145
MAX_DETOUR = 1.5      # lib/Constants.py:144
COEF_INVEH, COEF_WAIT = 1.0, 1.5


def _grid_index(data):
    lng = np.asarray(data.vattrs["lng"]); lat = np.asarray(data.vattrs["lat"])
    gi = np.rint((lng - LNG0) / STEP).astype(np.int64)
    gj = np.rint((lat - LAT0) / STEP).astype(np.int64)
    side = int(max(gi.max(), gj.max())) + 1
    table = -np.ones((side, side), dtype=np.int64)
    table[gi, gj] = np.arange(data.n)
    return gi, gj, table


def _near(rs, table, gi, gj, node, radius):
    side = table.shape[0]
    for _ in range(64):
        i = int(np.clip(gi[node] + rs.randint(-radius, radius + 1), 0, side - 1))
        j = int(np.clip(gj[node] + rs.randint(-radius, radius + 1), 0, side - 1))
        if table[i, j] >= 0 and table[i, j] != node:
            return int(table[i, j])
    return int((node + 1) % len(gi))


def fleet_scenario(data, n_veh, n_new, T, seed, tt_lookup, plan_stops=(0, 2, 4, 6, 8), capacity=4, radius=6):
    """Synthetic per-tick state (BASELINE config C4): ``n_veh`` vehicles with feasible plans of
    L planned stops (L drawn uniformly from ``plan_stops``) and ``n_new`` unassigned requests.

    ``tt_lookup(s, t)`` returns travel times for int arrays of vertex pairs (the resident
    matrix).  Returns a dict of plain arrays / lists:
      table    request table columns (o, d, Cep, Clp, Ced, Cld, Ts, Tr, pwt), one row per request
      new_rows rows of the unassigned requests
      vehs     list of dicts {vid, node, K, n, T, c, t0, route [(row, pod, node)], pax [rows], wait [rows]}
    """
    rs = np.random.RandomState(seed)
    gi, gj, table = _grid_index(data)
    vnode = rs.randint(0, data.n, size=n_veh)
    Ls = rs.choice(np.asarray(plan_stops), size=n_veh)
    rows_o, rows_d = [], []
    plans = []                                   # per vehicle: list of (row, pod)
    roles = []                                   # per request row: ("A"|"W"|"N", vehicle)
    for v in range(n_veh):
        tmpl = _PLAN_TEMPLATES[int(Ls[v])]
        names = {}
        cur = int(vnode[v])
        plan = []
        for (name, pod) in tmpl:
            if name not in names:
                row = len(rows_o)
                names[name] = row
                if name[0] == "A":               # on board: picked up somewhere behind the vehicle
                    o = _near(rs, table, gi, gj, int(vnode[v]), radius)
                    d = _near(rs, table, gi, gj, cur, radius)
                else:
                    o = _near(rs, table, gi, gj, cur, radius)
                    d = _near(rs, table, gi, gj, o, radius)
                rows_o.append(o); rows_d.append(d); roles.append((name[0], v))
            row = names[name]
            node = rows_o[row] if pod == 1 else rows_d[row]
            plan.append((row, pod, node))
            cur = node
        plans.append(plan)
    new_rows = []
    for _ in range(n_new):
        o = int(rs.randint(0, data.n))
        d = _near(rs, table, gi, gj, o, 2 * radius)
        new_rows.append(len(rows_o)); rows_o.append(o); rows_d.append(d); roles.append(("N", -1))
    rows_o = np.asarray(rows_o, dtype=np.int32); rows_d = np.asarray(rows_d, dtype=np.int32)
    nrows = len(rows_o)
    Ts = tt_lookup(rows_o, rows_d)
    # plan segment times and, for on-board requests, origin -> vehicle vertex
    seg_s, seg_t = [], []
    for v in range(n_veh):
        cur = int(vnode[v])
        for (_, _, node) in plans[v]:
            seg_s.append(cur); seg_t.append(node); cur = node
    seg_tt = tt_lookup(np.asarray(seg_s, dtype=np.int32), np.asarray(seg_t, dtype=np.int32)) if seg_s else np.zeros(0)
    a_rows = [r for r in range(nrows) if roles[r][0] == "A"]
    a_tt = tt_lookup(rows_o[a_rows], np.asarray([vnode[roles[r][1]] for r in a_rows], dtype=np.int32)) if a_rows else np.zeros(0)
    a_back = dict(zip(a_rows, a_tt))
    Cep = np.zeros(nrows); Clp = np.zeros(nrows); Tr = np.zeros(nrows); pwt = np.zeros(nrows)
    Ced = np.zeros(nrows); Cld = np.zeros(nrows)
    vehs = []
    k = 0
    for v in range(n_veh):
        t = 0.0
        n = sum(1 for (row, pod, _) in plans[v] if roles[row][0] == "A")
        n0 = n
        c = 0.0
        for (row, pod, node) in plans[v]:
            dt = float(seg_tt[k]); k += 1
            t += dt
            c += n * dt * COEF_INVEH
            n += pod
            if pod == 1:
                c += t * COEF_WAIT
                Cep[row] = T + t - float(rs.randint(0, 300))      # arrival inside [Cep, Cep + MAX_WAIT]
                pwt[row] = t if t > 0 else 1.0
        for (row, pod, node) in plans[v]:
            if roles[row][0] == "A" and pod == -1:
                Cep[row] = T - a_back[row] - float(rs.randint(0, 100))   # picked up at o, rode to the vehicle vertex
        vehs.append({"vid": v, "node": int(vnode[v]), "K": capacity, "n": n0, "T": float(T), "c": c, "t0": 0.0,
                     "route": plans[v],
                     "pax": [row for (row, pod, _) in plans[v] if roles[row][0] == "A"],
                     "wait": [row for (row, pod, _) in plans[v] if roles[row][0] == "W" and pod == 1]})
    for r in new_rows:
        Cep[r] = T - float(rs.randint(0, 30))
    Tr[:] = Cep
    Clp[:] = Cep + MAX_WAIT
    Ced[:] = Cep + Ts
    Cld[:] = Clp + MAX_DETOUR * Ts
    tab = {"o": rows_o, "d": rows_d, "Cep": Cep, "Clp": Clp, "Ced": Ced, "Cld": Cld, "Ts": np.asarray(Ts, dtype=np.float64),
           "Tr": Tr, "pwt": pwt}
    return {"table": tab, "new_rows": np.asarray(new_rows, dtype=np.int32), "vehs": vehs, "T": float(T)}
