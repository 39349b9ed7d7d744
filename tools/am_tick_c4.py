"""One Alonso-Mora tick at BASELINE config C4's scale through the drop-in API (AlonsoMora.optimizeAssignment):
10 000 vehicles (half of them carrying a rider), 2 000 new requests, on the 50k-node map.  Prints one JSON line per
trip-size cap with the stage times the reference prints (R-V graph, RTV graph, assignment).

    python tools/am_tick_c4.py [--vehs 10000] [--reqs 2000] [--caps 1,2,0] [--budget 600]

caps: 0 = vehicle capacity (the reference's behaviour).  Every cap runs under a wall-clock budget (SIGALRM) so
that a combinatorial blow-up in the host-side RTV enumeration cannot hang the box."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def build(args):
    import bench
    from drs_abm_b200 import synth
    from drs_abm_b200.routing import RoadGraph
    data = synth.config_graph("C3")
    t0 = time.time()
    G = RoadGraph(data).build()
    t_build = time.time() - t0
    lng, lat = data.vattrs["lng"], data.vattrs["lat"]
    rs = np.random.RandomState(4)
    T = 3600.0
    vehs = [bench._SimVeh(k, 4, int(rs.randint(data.n)), lng, lat) for k in range(args.vehs)]
    reqs = []
    # half of the fleet has just picked a rider up where it stands
    busy = rs.permutation(args.vehs)[: args.vehs // 2]
    o = np.array([vehs[v].node for v in busy])
    d = (o + 1 + rs.randint(0, data.n - 1, len(busy))) % data.n
    Ts = G.path_sums(o, d)[0]
    for k, v in enumerate(busy):
        if not np.isfinite(Ts[k]):
            continue
        rq = bench._SimReq(len(reqs), int(o[k]), int(d[k]), lng, lat, T - float(rs.randint(0, 300)), float(Ts[k]))
        rq.assigned = True
        reqs.append(rq)
        vehs[v].onboard.add(rq.id)
        vehs[v].plan = [(rq.id, -1, int(d[k]), T + float(Ts[k]))]
        vehs[v].node = int(o[k])

    # _SimVeh.position() reports the NEXT stop; a vehicle that has just picked up stands on the pickup vertex
    for v in vehs:
        nd = v.node
        v.get_next_node = (lambda nd=nd: ((lng[nd], lat[nd]), 0.0))
    o = rs.randint(0, data.n, args.reqs)
    d = (o + 1 + rs.randint(0, data.n - 1, args.reqs)) % data.n
    Ts = G.path_sums(o, d)[0]
    for k in range(args.reqs):
        if np.isfinite(Ts[k]):
            reqs.append(bench._SimReq(len(reqs), int(o[k]), int(d[k]), lng, lat, T - float(rs.randint(0, 30)), float(Ts[k])))
    return G, vehs, reqs, T, t_build


def one_cap(G, vehs, reqs, T, cap, out):
    from drs_abm_b200.dispatch import AlonsoMora
    params = {"routing_method": "enumerate", "verbose": False}
    if cap:
        params["max_trip_size"] = cap
    opt = AlonsoMora(params)
    if os.environ.get("AM_DUMP_ILP"):
        # keep the set-packing instance of this tick for solver experiments on a CPU-only machine
        from drs_abm_b200 import dispatch as Dm
        inner = Dm._solve_packing_arrays

        def dumping(nveh, nreq, veh, size, members, costs):
            np.savez_compressed(os.environ["AM_DUMP_ILP"] + "_cap%s.npz" % (cap or "K"), nveh=nveh, nreq=nreq, veh=veh, size=size,
                                members=members, costs=costs)
            return inner(nveh, nreq, veh, size, members, costs)
        Dm._solve_packing_arrays = dumping
    t0 = time.time()
    (routes, rej), meta = opt.optimizeAssignment(vehs, reqs, G, T)
    dt = time.time() - t0
    line = {"max_trip_size": cap or "capacity", "vehicles": len(vehs), "new_requests": sum(1 for r in reqs if not r.assigned),
            "tick_s": dt, "rv_s": meta.get("rv_time"), "rtv_s": meta.get("rtv_time"), "assignment_s": meta.get("solve_time"),
            "objective": meta.get("objective"), "vehicles_assigned": len(routes) if isinstance(routes, dict) else None,
            "rejected": len(rej) if rej is not None else None}
    with open(out, "a") as f:
        f.write(json.dumps(line) + "\n")
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vehs", type=int, default=10000)
    ap.add_argument("--reqs", type=int, default=2000)
    ap.add_argument("--caps", default="1,2,0")
    ap.add_argument("--budget", type=float, default=600.0)
    ap.add_argument("--out", default="gpurun_out/r02_am_tick_c4.jsonl")
    args = ap.parse_args()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    G, vehs, reqs, T, t_build = build(args)
    print(json.dumps({"apsp_and_setup_s": t_build, "vehicles": len(vehs), "requests_total": len(reqs)}), flush=True)
    for cap in [int(c) for c in args.caps.split(",")]:
        # twice: the first call of a cap pays the arena growth, the second is the steady state
        for rep in range(2):
            t0 = time.time()
            try:
                import signal

                def on_alarm(*a):
                    raise TimeoutError()
                signal.signal(signal.SIGALRM, on_alarm)
                signal.alarm(int(args.budget))
                one_cap(G, vehs, reqs, T, cap, args.out)
                signal.alarm(0)
            except TimeoutError:
                line = {"max_trip_size": cap or "capacity", "timeout_s": args.budget}
                with open(args.out, "a") as f:
                    f.write(json.dumps(line) + "\n")
                print(json.dumps(line), flush=True)
                return


if __name__ == "__main__":
    main()
