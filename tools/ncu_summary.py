"""Summarise an .ncu-rep (ncu --set full) into the few numbers the profiles/ files quote: duration, DRAM bytes, L2 hit rate,
issue utilisation, occupancy limits and the top warp-stall reasons per issue slot.

    python tools/ncu_summary.py gpurun_out/foo.ncu-rep [> profiles/rNN_ncu_foo.txt]
"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    print(f"# {rep}: {len(rows) - 2} launch(es)")
    for n, r in enumerate(rows[2:], 1):
        print(f"\n## launch {n}: {r[idx['Kernel Name']]}")
        for k in KEYS:
            if k in idx:
                print(f"{k:80s} {r[idx[k]]:>24s} {units[idx[k]]}")
        top = sorted(((float(r[idx[h]] or 0), h) for h in stalls), reverse=True)[:6]
        for v, h in top:
            print(f"{h:80s} {v:24.3f} warps per issue slot")


if __name__ == "__main__":
    main()
