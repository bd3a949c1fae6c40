#!/usr/bin/env python
"""Summarise ncu outputs into small text files under profiles/ (the judged, committed evidence).

    python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_launches_cfg1_fp32.txt [skip_first_n]
    python tools/ncu_summary.py full gpurun_out/prof_conv.ncu-rep profiles/r01_conv_gemm_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__cycles_active.avg",
        "smsp__cycles_active.avg", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def launches(src, dst, skip=0):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    body = rows[1 + skip:]
    agg = collections.OrderedDict()
    for r in body:
        name = r[ki].split("(")[0]
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none; source {src}; first {skip} launches skipped\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"# {'kernel':70s} {'launches':>8s} {'sum_ms':>10s} {'avg_us':>10s} {'share':>7s}\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"{k:72s} {len(v):8d} {sum(v) / 1e6:10.3f} {sum(v) / len(v) / 1e3:10.1f} {sum(v) / tot:7.3f}\n")
        f.write(f"# total {tot / 1e6:.3f} ms over {sum(len(v) for v in agg.values())} launches\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on; source {src}\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n== launch id {d.get('ID')}  {d.get('Kernel Name', '')[:110]}\n")
            f.write(f"   grid {d.get('launch__grid_size')} block {d.get('launch__block_size')}\n")
            for k in hdr:
                if any(k == key or k.startswith(key) for key in KEYS):
                    f.write(f"   {k:75s} {d[k]:>18s} {units[hdr.index(k)]}\n")
    print(open(dst).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 0)
    else:
        full(sys.argv[2], sys.argv[3])
