"""Turns the raw ncu outputs brought back in gpurun_out/ into the small text summaries committed here.
  python profiles/summarize.py launches gpurun_out/r1_tc_launches.csv > profiles/r1_launches_tc.txt
  python profiles/summarize.py full gpurun_out/r1_gemm_tc.ncu-rep > profiles/r1_gemm_tc_ncu.txt"""
import csv
import subprocess
import sys
from collections import defaultdict

KEYS = ['Kernel Name', 'Grid Size', 'Block Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows[1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        name = r[ki].split("(")[0]
        tot[name] += v
        cnt[name] += 1
    unit = rows[1][hdr.index("Metric Unit")] if len(rows) > 1 else "?"
    s = sum(tot.values())
    print(f"# ncu --metrics gpu__time_duration.sum launch list: {sum(cnt.values())} launches, {s:.0f} {unit} total")
    print(f"# (cold-cache, serialised: compare SHARES with bench.py's kernel_time_shares, not absolutes)")
    print(f"{'kernel':70s} {'launches':>8s} {'total':>14s} {'share':>7s} {'avg':>10s}")
    for k in sorted(tot, key=lambda k: -tot[k]):
        print(f"{k[:70]:70s} {cnt[k]:8d} {tot[k]:14.0f} {tot[k] / s:7.3f} {tot[k] / cnt[k]:10.1f}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = [hdr.index(k) for k in KEYS if k in hdr]
    print(f"# ncu --set full --clock-control none: {path}")
    for r in rows[2:]:
        print("---")
        for i in idx:
            print(f"{hdr[i]}: {r[i]} {units[i]}")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
