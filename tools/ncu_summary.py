#!/usr/bin/env python
"""Summarise one `ncu --set full` report (raw page) into the numbers DESIGN.md / bench.py cite:
  python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput % of peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "TENSOR PIPE active % of peak (while SM active)"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of peak (elapsed)"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak (gpu__dram_throughput)"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock during the capture"),
    ("sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) active % of peak"),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "tensor pipe inst % of peak"),
    ("sm__pipe_tensor_subpipe_umma_cycles_active.avg.pct_of_peak_sustained_active", "UMMA sub-pipe active % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem), CTAs/SM"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"source: `{rep}` (ncu --set full --clock-control none)\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"### `{name[:120]}`\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for key, label in KEYS:
            for i, h in enumerate(hdr):
                if h == key:
                    print(f"| {label} (`{key}`) | {r[i]} | {units[i]} |")
        stall = []
        for i, h in enumerate(hdr):
            if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h:
                try:
                    stall.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stall) or 1.0
        print("\nwarp stall samples: " + ", ".join(f"{h} {100 * v / tot:.0f}%" for v, h in sorted(stall, reverse=True)[:6]) + "\n")


if __name__ == "__main__":
    main()
