"""SASS opcode histogram of libdrs_b200.so per kernel family (cuobjdump -sass; no GPU needed).
    python tools/sass_opcodes.py > profiles/r02_sass_opcodes.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "drs_abm_b200", "libdrs_b200.so")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
FAMILIES = [("sep_expand16", "separator build, 16-bit output tiles (sep_apsp.cu)"), ("sep_expand", "separator build, 32-bit tiles: output + vertex->separator rows"),
            ("sep_fw_", "separator build, batched part / separator closures"), ("sep_", "separator build, fills / gathers / spread"),
            ("fw_", "Floyd-Warshall tiles (fw_kernels.cu)"), ("pred_kernel", "canonical predecessors"), ("sssp_batch", "SSSP, rows in L2 (default)"),
            ("sssp_smem", "SSSP, alternative kernels (row in shared memory / lane per arc)"), ("insertion_reach", "insertion reach test (bulk-copy staged)"),
            ("insertion_eval", "insertion pair evaluation"), ("insertion_fold", "insertion fold"), ("insertion_queue", "insertion sequential queue"),
            ("insertion_commit", "insertion commit"), ("plan_prep", "insertion plan prep"), ("rv_filter", "RV filter"), ("rv_eval", "RV evaluation"),
            ("trip_eval", "trip evaluation"), ("path_query", "path queries"), ("matrix_gather", "matrix look-ups")]
INTEREST = ["UBLKCP", "UBLKPF", "SYNCS", "LDG.E.ENL2.256", "LDGSTS", "FADD2", "FMNMX3", "VIADDMNMX", "ATOMS", "ATOMG", "RED", "LDS", "STS", "DADD", "DFMA",
            "DMUL", "DSETP", "SHFL", "VOTE", "BAR", "MEMBAR", "STG.E.EF", "UTMALDG", "UTCHMMA", "HMMA"]
hist = collections.OrderedDict((f[0], collections.Counter()) for f in FAMILIES)
total = collections.Counter()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = next((f[0] for f in FAMILIES if f[0] in m.group(1)), None)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        hist[cur][m.group(1)] += 1
        total[cur] += 1
print("# SASS opcodes of `drs_abm_b200/libdrs_b200.so` (sm_100a), by kernel family\n")
print("`python tools/sass_opcodes.py` (cuobjdump -sass; counts are static instruction counts over all template instantiations of a family).\n")
print("| family | instructions | " + " | ".join(INTEREST) + " |")
print("|---|---|" + "---|" * len(INTEREST))
for key, title in FAMILIES:
    h = hist[key]
    if not total[key]:
        continue
    cells = []
    for op in INTEREST:
        c = sum(v for k, v in h.items() if k.startswith(op))
        cells.append(str(c) if c else "")
    print("| %s | %d | %s |" % (title, total[key], " | ".join(cells)))
print("\nBulk-copy engine: `UBLKCP.S.G` = `cp.async.bulk.shared::cluster.global` (the insertion reach test stages a matrix row per request),")
print("`UBLKPF.L2` = `cp.async.bulk.prefetch.L2`, `SYNCS.*` = mbarrier arrive / try_wait.  `LDG.E.ENL2.256` = 256-bit global load (SSSP adjacency")
print("records).  `FADD2` / `FMNMX3` / `VIADDMNMX` = packed fp32x2 add, 3-input min, DPX add+min (Floyd-Warshall and separator-build tiles;")
print("`VIADDMNMX.U16x2`, two 16-bit relaxations per lane, in the 16-bit output tiles: %d static instances).  No `UTC*MMA` / `HMMA`:" % sum(v for k, v in hist["sep_expand16"].items() if k.startswith("VIADDMNMX.U16")))
print("min-plus is not a GEMM.")
