"""Key counters of an .ncu-rep as a markdown table (reads the report with `ncu -i ... --page raw --csv`; no GPU needed).
    python tools/ncu_summary.py gpurun_out/r02_targets.ncu-rep > profiles/r02_kernels_ncu.md"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
METRICS = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"),
           ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
           ("lts__t_sectors_srcunit_tex_op_read.sum", "L2->L1 sectors read"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
           ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
           ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
           ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
           ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
           ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
           ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
           ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
           ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long scoreboard (warps per issue)"),
           ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short scoreboard"),
           ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
           ("smsp__inst_executed.sum", "warp instructions"), ("launch__registers_per_thread", "registers / thread"),
           ("launch__grid_size", "grid"), ("launch__block_size", "block")]
kernels = rows[2:]
names = []
for r in kernels:
    nm = r[ix["Kernel Name"]]
    nm = nm.split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    names.append(nm)
print("# ncu --set full --clock-control none: key counters per launch (`%s`)\n" % rep)
print("| metric | " + " | ".join(names) + " |")
print("|---|" + "---|" * len(names))
for key, title in METRICS:
    if key not in ix:
        continue
    cells = []
    for r in kernels:
        v = r[ix[key]]
        try:
            f = float(v)
            v = ("%.3f" % f).rstrip("0").rstrip(".") if abs(f) < 1e6 else "%.4g" % f
        except ValueError:
            pass
        cells.append(v)
    print("| %s [%s] | %s |" % (title, units[ix[key]], " | ".join(cells)))
