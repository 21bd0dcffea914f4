#!/usr/bin/env python
"""Condense an `ncu --set full` report into the few counters DESIGN.md / bench.py cite.

  python scripts/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>_ncu_full.txt

Per profiled launch: duration, DRAM read/write bytes (the `roofline.traffic` figure),
tensor-pipe activity, L2 traffic, issue-slot use, registers, the top warp-stall reasons.
Runs in the build container (ncu reads reports without a GPU).
"""
import csv
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed (max)"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes (all traffic)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor hmma-subpipe active cycles (avg/SM)"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor-memory (TMEM) active %"),
    ("sm__inst_executed_pipe_uniform.sum", "uniform-pipe instructions"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_scoreboard %"),
    ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stall barrier %"),
    ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "stall wait %"),
    ("smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "stall no_instruction %"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "stall short_scoreboard %"),
    ("smsp__warp_issue_stalled_membar_per_warp_active.pct", "stall membar %"),
    ("smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "stall sleeping %"),
    ("smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "stall branch_resolving %"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall math_pipe_throttle %"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "stall lg_throttle %"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        sys.exit("no launches in report")
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary of {rep}  ({len(rows) - 2} profiled launch(es))")
    for r in rows[2:]:
        print()
        print("kernel:", r[col["Kernel Name"]][:140])
        for name, label in KEEP:
            hits = [h for h in hdr if h == name or h.endswith("." + name)]
            for h in hits[:1]:
                print(f"  {label:48s} {r[col[h]]:>20s} {units[col[h]]}")
        try:
            rd = float(r[col["dram__bytes_read.sum"]].replace(",", ""))
            wr = float(r[col["dram__bytes_write.sum"]].replace(",", ""))
            mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            tot = rd * mult[units[col["dram__bytes_read.sum"]]] + wr * mult[units[col["dram__bytes_write.sum"]]]
            print(f"  {'DRAM traffic (read+write), bytes':48s} {tot:20.0f}")
        except Exception:
            pass


if __name__ == "__main__":
    main()
