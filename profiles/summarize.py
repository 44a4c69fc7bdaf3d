"""Turn gpurun_out/ ncu artefacts into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv > profiles/launches_r1.txt
    python profiles/summarize.py full gpurun_out/prof_contract_r1.ncu-rep > profiles/ncu_contract_r1.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEY = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__cluster_dim_x", "smsp__cycles_active.avg",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__warps_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        ns = float(r[14])
        tot += ns
        name = re.sub(r"\(.*", "", r[4])
        if "contract" not in name:
            name = re.sub(r"<.*", "", name)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none: {len(rows)} launches, {tot / 1e6:.3f} ms total (cold-cache, serialised)")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{ns / 1e6:10.3f} ms {100 * ns / tot:5.1f}%  x{n:4d}  avg {ns / n / 1e3:9.1f} us  {k[:120]}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("=" * 100)
        print(r[hdr.index("Kernel Name")], " grid", r[hdr.index("Grid Size")] if "Grid Size" in hdr else "")
        for k in KEY:
            if k in hdr:
                print(f"  {k:75s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
