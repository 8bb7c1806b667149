#!/usr/bin/env python3
"""Summarise an `ncu --set full` report of the row/col kernels into profiles/<round>/ncu_full_row_col.md and
profiles/traffic.json (the per-launch DRAM traffic bench.py reports as roofline.traffic).
usage: ncu_summary.py REPORT.ncu-rep OUT.md "<command that was profiled>" [frames_per_launch]"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg.per_second', 'dram__cycles_elapsed.avg.per_second',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct', 'smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct',
        'smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_wait_per_warp_active.pct',
        'smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct', 'smsp__warp_issue_stalled_drain_per_warp_active.pct',
        'smsp__warp_issue_stalled_no_instruction_per_warp_active.pct', 'smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct']
SCALE = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0, 'Tbyte': 1e12}


def main():
    rep, out_md, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
    frames = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(io.StringIO(raw)))
    hdr, units, rows = r[0], r[1], r[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    N, E = 18432, 147456
    with open(out_md, "w") as f:
        f.write("# ncu --set full --clock-control none: `%s`\n\n" % cmd)
        f.write("One launch = %d frames (groups of 32), n=18432/m=2048 code, fp64. Algorithmic bytes per launch: "
                "row 16*E*F = %.3f GB, col (16*E+8.25*N)*F = %.3f GB.\n\n" % (frames, 16 * E * frames / 1e9, (16 * E + 8.25 * N) * frames / 1e9))
        names = [rows[i][idx['Kernel Name']].split('<')[0].replace('void ', '').replace('dnaldpc::', '') + " #%d" % i for i in range(len(rows))]
        f.write("| metric | unit | " + " | ".join(names) + " |\n|---|---|" + "---|" * len(rows) + "\n")
        for k in KEYS:
            if k in idx:
                f.write("| %s | %s | " % (k, units[idx[k]]) + " | ".join(rows[i][idx[k]] for i in range(len(rows))) + " |\n")

    def tot(x):
        return sum(float(x[idx[k]]) * SCALE[units[idx[k]]] for k in ('dram__bytes_read.sum', 'dram__bytes_write.sum'))
    tr = {}
    for tag in ("row", "col"):
        sel = [x for x in rows if tag + "_pass" in x[idx['Kernel Name']]]
        # launches of ticks in which the slots were empty (between two waves of a short batch) are not "a launch over
        # 4096 frames": keep the ones that moved at least half of the largest launch's bytes
        if sel:
            top = max(tot(x) for x in sel)
            sel = [x for x in sel if tot(x) >= 0.5 * top]
        if sel:
            tr[tag] = sum(tot(x) for x in sel) / len(sel)
            tr[tag + "_per_frame"] = tr[tag] / frames
    tr["frames_per_launch"] = frames
    tr["note"] = "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, " + os.path.relpath(out_md, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    json.dump(tr, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(tr))


if __name__ == "__main__":
    main()
