#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics per kernel launch + SASS opcode mix per kernel (needs ncu on PATH)."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
units_per_step = None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
want = [
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active%"),
    ("launch__registers_per_thread", "regs"), ("launch__occupancy_limit_shared_mem", "occ_smem"),
    ("launch__occupancy_limit_registers", "occ_regs"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("l1tex__t_sector_hit_rate.pct", "l1hit%"), ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall_short_sb"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall_mio"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall_wait"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"),
    ("smsp__inst_executed.sum", "warp_inst"), ("launch__grid_size", "grid"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem_conflicts"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64pipe%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64cyc%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem_wavefront%"),
]
seen = collections.Counter()
for r in data:
    name = re.sub(r"\(.*", "", r[idx["Kernel Name"]])
    full = r[idx["Kernel Name"]][:70]
    seen[full] += 1
    if seen[full] > 1:
        continue
    print("----", full)
    line = []
    for k, short in want:
        if k in idx:
            v = r[idx[k]]
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            line.append(f"{short}={v}{units[idx[k]] if short in ('time','dram_rd','dram_wr') else ''}")
    print("   " + "  ".join(line))
    # every stall reason, cycles per issued instruction (sorted)
    st = []
    for h, i in idx.items():
        m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio", h)
        if m:
            try:
                st.append((float(r[i]), m.group(1)))
            except ValueError:
                pass
    print("   stalls/issue: " + "  ".join(f"{n}={v:.3f}" for v, n in sorted(st, reverse=True) if v >= 0.01))

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        secs.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
done = set()
for s in secs:
    if s["name"] in done or not s["rows"]:
        continue
    done.add(s["name"])
    h = s["rows"][0]
    ix = {n: i for i, n in enumerate(h)}
    ops, tot = collections.Counter(), 0
    for r in s["rows"][1:]:
        try:
            n = int(r[ix["Instructions Executed"]])
        except (ValueError, IndexError, KeyError):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
        ops[m.group(2).split(".")[0] if m else "?"] += n
        tot += n
    print("==== opcode mix:", s["name"][:70], "total warp-instr", tot)
    print("    " + ", ".join(f"{o}:{100.0*n/tot:.1f}%" for o, n in ops.most_common(22)))
