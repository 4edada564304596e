"""Turns an .ncu-rep (ncu --set full --import-source on) into the text summary committed under profiles/.
   python tools/ncu_summarize.py gpurun_out/x.ncu-rep [--source N]   (N = launch index for the per-line table)"""
import csv, io, subprocess, sys

rep = sys.argv[1]
src_idx = int(sys.argv[sys.argv.index("--source") + 1]) if "--source" in sys.argv else None

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % of peak"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % of peak"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall: math pipe throttle / issue"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall: not selected / issue"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall: long scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall: short scoreboard / issue"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall: barrier / issue"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall: wait / issue"),
]
print(f"# {rep}")
for li, d in enumerate(data):
    name = d[ix["Kernel Name"]]
    grid = d[ix["Grid Size"]] if "Grid Size" in ix else ""
    block = d[ix["Block Size"]] if "Block Size" in ix else ""
    print(f"\n== launch {li}: {name}  grid {grid} block {block}")
    for k, label in KEYS:
        if k in ix:
            print(f"   {label:38s} {d[ix[k]]:>18s} {units[ix[k]]}")

if src_idx is not None:
    name = data[src_idx][ix["Kernel Name"]].split("(")[0].split("<")[0].split()[-1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id",
                          f"::regex:{name}:{src_idx + 1}"], capture_output=True, text=True).stdout
    cur, ie, lines = None, None, []
    for r in csv.reader(io.StringIO(out)):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if "Instructions Executed" in r:
            ie = r.index("Instructions Executed")
            continue
        if ie is None or len(r) <= ie or r[2] != "-":
            continue
        try:
            n = int(r[ie])
        except ValueError:
            continue
        if n > 0:
            lines.append((n, cur, r[0], r[1].strip()[:100]))
    tot = sum(x[0] for x in lines)
    print(f"\n== per source line, launch {src_idx} ({tot} warp instructions attributed); top 30")
    for n, f, l, sr in sorted(lines, reverse=True)[:30]:
        print(f"   {n * 100 / tot:5.1f}%  {f}:{l:>4s}  {sr}")
