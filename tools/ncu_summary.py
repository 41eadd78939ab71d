#!/usr/bin/env python
"""Extract the metrics the roofline argument rests on from an `ncu --set full` report (read here, no GPU needed).
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/xxx.md"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__cycles_elapsed.max.per_second", "launch__shared_mem_per_block_dynamic"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none: `{path}`\n")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"## {name[:110]}\n")
        print("| metric | value | unit |\n|---|---:|---|")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"| {k} | {r[i]} | {units[i]} |")
        print()


if __name__ == "__main__":
    main(sys.argv[1])
