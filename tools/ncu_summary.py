#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the few metrics the roofline
discussion uses.  Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/x.txt]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__t_bytes.sum", "L1 bytes"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__t_sectors.sum", "L2 sectors (32 B)"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 throughput % of peak"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64 pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers"),
    ("launch__grid_size", "grid"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stall long_scoreboard %"),
    ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stall math_pipe %"),
    ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "stall short_scoreboard %"),
    ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "stall wait %"),
    ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "stall lg_throttle %"),
    ("smsp__warp_issue_stalled_not_selected_per_warp_active.pct", "stall not_selected %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers)"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}")
    for r in rows[2:]:
        print(f"\n== {r[idx['Kernel Name']]}  (launch id {r[idx['ID']]})")
        for key, label in WANT:
            if key in idx:
                print(f"  {label:26s} {r[idx[key]]:>16s} {units[idx[key]]}")
        # warp-state sampling: where the warps spent their time (share of all samples)
        samp = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: float(r[i]) for h, i in idx.items()
                if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and r[i] not in ("", "n/a")}
        tot = sum(samp.values())
        if tot > 0:
            top = sorted(samp.items(), key=lambda kv: -kv[1])[:8]
            print("  warp-state samples         " + ", ".join(f"{k} {100.0 * v / tot:.0f} %" for k, v in top))


if __name__ == "__main__":
    main()
