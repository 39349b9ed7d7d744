"""B200-native routing-and-dispatch core for nbailey/drs-abm.

Drop-in classes (same names and call surface as the reference):
    drs_abm_b200.routing.RoutingEngine      lib/RoutingEngine.py
    drs_abm_b200.routing.RoadGraph          the igraph object callers reach through ``router.G``
    drs_abm_b200.dispatch.AlonsoMora        lib/Optimization.py
    drs_abm_b200.dispatch.InsertionHeuristics   lib/Agents.py:1429-1572 (legacy insertion heuristic)
Multi-GPU builds: drs_abm_b200.apsp_dist.  Everything computes in libdrs_b200.so
(``python -m drs_abm_b200.build``); there is no CPU fallback.
"""
__version__ = "0.1.0"
