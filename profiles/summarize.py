"""Turns gpurun_out ncu artefacts into the small, tracked summaries under profiles/.
    python profiles/summarize.py launches gpurun_out/launches_X.csv profiles/X_launches
    python profiles/summarize.py full gpurun_out/prof_X.ncu-rep profiles/X_ncu_full
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'smsp__inst_executed.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.max']


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith('==')]
    agg = collections.OrderedDict()
    tot = 0.0
    n = 0
    per = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = re.sub(r'\(.*', '', row['Kernel Name'])
        v = float(row['Metric Value'].replace(',', ''))
        v *= {'ns': 1, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(row['Metric Unit'], 1)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
        n += 1
        per.append((row['ID'], name, v))
    with open(dst + '.md', 'w') as f:
        f.write(f"# ncu launch list summary ({src})\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, "
                f"serialised: compare shares, not absolutes). {n} launches, {tot / 1e6:.3f} ms total.\n\n")
        f.write("| total ms | share | launches | avg us | kernel |\n|---:|---:|---:|---:|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {t / 1e6:.3f} | {100 * t / tot:.1f}% | {c} | {t / c / 1e3:.1f} | `{k[:120]}` |\n")
    with open(dst + '.csv', 'w') as f:
        f.write("id,kernel,duration_ns\n")
        for i, k, v in per:
            f.write(f"{i},\"{k[:100]}\",{v:.0f}\n")


def full(src, dst):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(dst + '.md', 'w') as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']][:140]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for w in KEEP[1:]:
                if w in idx:
                    f.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
            f.write("\n")


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3])
