#!/usr/bin/env python
"""Summarise ncu output for profiles/ (developer tool, run in the CPU container on files brought back by gpurun).

  python tools/ncu_summary.py launches gpurun_out/r01_launches_c3.csv    -> per-kernel totals / shares of the step
  python tools/ncu_summary.py full gpurun_out/r01_collide_c3.ncu-rep     -> key counters of each captured launch
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum",
        "launch__grid_size", "launch__block_size"]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    seq = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}.get(u, 1.0)
        seq.append((row["Kernel Name"].split("(")[0].replace("void ", ""), v))
    agg = collections.OrderedDict()
    for n, v in seq:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print("# %s: %d launches, %.1f us total (ncu per-launch times: cold cache, serialised -- compare SHARES)" % (path, len(seq), tot))
    print("%-44s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print("%-44s %6d %12.1f %10.1f %7.3f" % (n[:44], a[0], a[1], a[1] / a[0], a[1] / tot))
    print("# last launches in order:")
    for n, v in seq[-24:]:
        print("#   %-44s %10.1f us" % (n[:44], v))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print("# %s" % path)
    for r in rows[2:]:
        print("kernel: %s" % r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print("  %-66s %s %s" % (k, r[hdr.index(k)], units[hdr.index(k)]))
        if "dram__bytes_read.sum" in hdr:
            def gb(k):
                v = float(r[hdr.index(k)].replace(",", ""))
                return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[units[hdr.index(k)]]
            print("  %-66s %.0f" % ("traffic_bytes_per_launch (dram read+write)", gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
