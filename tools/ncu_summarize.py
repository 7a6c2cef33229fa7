"""Turn the ncu outputs a GPU visit brought back (gpurun_out/) into the tracked summaries under profiles/.

    python tools/ncu_summarize.py <tag> <round>     # e.g. r1c 01

 - gpurun_out/launches_<tag>.csv   (ncu --metrics gpu__time_duration.sum --csv)  -> profiles/r<round>_launches.csv + _summary.md
 - gpurun_out/prof_<tag>.ncu-rep   (ncu --set full)                              -> profiles/r<round>_ncu_full_summary.md
"""
import collections
import csv
import io
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rnd = sys.argv[1], sys.argv[2]
out = os.path.join(ROOT, 'profiles')
os.makedirs(out, exist_ok=True)

# ---- launch list ---------------------------------------------------------------------------------
src = os.path.join(ROOT, 'gpurun_out', f'launches_{tag}.csv')
if os.path.exists(src):
    lines = open(src).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = list(csv.DictReader(io.StringIO('\n'.join(lines[start:]))))
    shutil.copy(src, os.path.join(out, f'r{rnd}_launches.csv'))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(r['Metric Value'].replace(',', ''))
        unit = r['Metric Unit']
        us = v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v
        name = r['Kernel Name']
        name = name[5:] if name.startswith('void ') else name
        name = name.split('(')[0][:70]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(out, f'r{rnd}_launches_summary.md'), 'w') as f:
        f.write(f'# Round {int(rnd)} -- ncu launch list of `python bench.py --steps 3 --warmup 3 --no-cpu` (first 300 launches)\n\n'
                '`ncu --metrics gpu__time_duration.sum --clock-control none -c 300` (per-launch times are cold-cache and serialised: compare SHARES).\n'
                f'Raw CSV: profiles/r{rnd}_launches.csv\n\n| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n')
        for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'| `{name}` | {n} | {us / n:.1f} | {us:.0f} | {100 * us / total:.1f}% |\n')
    print('wrote launch summary,', len(agg), 'kernels')

# ---- full capture --------------------------------------------------------------------------------
# (several captures of the same command may be merged: extra tags after the round number, e.g. `r2f 01 r2g`; the first
# occurrence of a kernel wins)
reps = [os.path.join(ROOT, 'gpurun_out', f'prof_{t}.ncu-rep') for t in [tag] + sys.argv[3:]]
reps = [r for r in reps if os.path.exists(r)]
if reps:
    rows, seen, row_units = None, set(), {}
    for rep in reps:
        raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
        part = list(csv.reader(io.StringIO(raw)))
        if rows is None:
            rows = part[:2]
        ci = part[0].index('Kernel Name')
        for r in part[2:]:
            if r[ci] not in seen and part[0] == rows[0]:
                seen.add(r[ci])
                rows.append(r)
                row_units[id(r)] = part[1]          # ncu scales the units per report
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    want = [
        ('duration', 'gpu__time_duration.sum'),
        ('tensor pipe active % (of peak, while SM active)', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
        ('tensor pipe active % (of peak, elapsed)', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'),
        ('SM throughput %', 'sm__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('DRAM throughput %', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
        ('DRAM read', 'dram__bytes_read.sum'),
        ('DRAM write', 'dram__bytes_write.sum'),
        ('DRAM read rate', 'dram__bytes_read.sum.per_second'),
        ('L2 throughput %', 'lts__throughput.avg.pct_of_peak_sustained_elapsed'),
        ('registers/thread', 'launch__registers_per_thread'),
        ('dynamic smem/block', 'launch__shared_mem_per_block_dynamic'),
        ('grid', 'launch__grid_size'),
        ('block', 'launch__block_size'),
        ('achieved occupancy %', 'sm__warps_active.avg.pct_of_peak_sustained_active'),
        ('warp instructions', 'smsp__inst_executed.sum'),
        ('SM cycles elapsed', 'sm__cycles_elapsed.avg'),
    ]
    with open(os.path.join(out, f'r{rnd}_ncu_full_summary.md'), 'w') as f:
        f.write(f'# Round {int(rnd)} -- `ncu --set full --clock-control none --import-source on` of the hot kernels\n\n'
                'Command: `python bench.py --steps 3 --warmup 3 --no-cpu` (B200, 1 GPU), kernels `mlp_kernel|composite_kernel|mask_kernel`, '
                f'launch-skip 24, count 4 (+ a second capture of the two MLP kernels only).\nReport files: ' + ', '.join('gpurun_out/' + os.path.basename(r) for r in reps) + ' (scratch, not committed); numbers below are per launch.\n'
                '`traffic` of bench.py\'s roofline object = DRAM read + DRAM write of the kernel\'s row here.\n')
        for r in rows[2:]:
            f.write(f'\n## `{r[col["Kernel Name"]][:90]}`\n\n| metric | value |\n|---|---|\n')
            for label, key in want:
                if key in col:
                    f.write(f'| {label} | {r[col[key]]} {row_units[id(r)][col[key]]} |\n')
    # per-launch DRAM traffic of each kernel, read by bench.py for roofline.traffic
    import json

    def to_bytes(v, u):
        v = float(v.replace(',', ''))
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    traffic, tensor = {}, {}
    for r in rows[2:]:
        name = r[col['Kernel Name']]
        name = (name[5:] if name.startswith('void ') else name).split('(')[0]
        tensor[name] = float(r[col['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']].replace(',', ''))
        u = row_units[id(r)]
        traffic[name] = to_bytes(r[col['dram__bytes_read.sum']], u[col['dram__bytes_read.sum']]) + \
            to_bytes(r[col['dram__bytes_write.sum']], u[col['dram__bytes_write.sum']])
    json.dump({'source': 'ncu --set full, ' + ', '.join('gpurun_out/' + os.path.basename(r) for r in reps) + ', bench.py --steps 3 --warmup 3 --no-cpu', 'unit': 'bytes per launch',
               'dram_read_plus_write': traffic, 'tensor_pipe_active_pct': tensor}, open(os.path.join(out, f'r{rnd}_traffic.json'), 'w'), indent=1)
    print('wrote full summary,', len(rows) - 2, 'kernels')
