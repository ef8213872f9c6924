"""Turn ncu outputs brought back in gpurun_out/ into the committed summaries under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md "<title>"
  python tools/summarize_ncu.py full gpurun_out/prof_r1.ncu-rep profiles/r1_full.md "<title>"
"""
import collections
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'lts__t_sectors.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__warps_eligible.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size',
        'launch__waves_per_multiprocessor', 'sm__cycles_elapsed.avg.per_second',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(src, dst, title):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(',', ''))
        v = v / 1e3 if r[ui] == 'ns' else v * 1e3 if r[ui] == 'ms' else v
        name = r[ki].split('(')[0].replace('void ', '')
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    with open(dst, 'w') as fh:
        fh.write(f'# {title}\n\nSource: `{src}` (ncu --metrics gpu__time_duration.sum --clock-control none; '
                 'cold-cache, serialised launches: compare shares, not absolutes).\n\n')
        fh.write('| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n')
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            fh.write(f'| `{k[:70]}` | {cnt[k]} | {v:.1f} | {100 * v / T:.1f}% | {v / cnt[k]:.1f} |\n')
    print(open(dst).read())


def full(src, dst, title):
    raw = subprocess.run(['ncu', '-i', src, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, 'w') as fh:
        fh.write(f'# {title}\n\nSource: `{src}` (ncu --set full --clock-control none --import-source on), '
                 'read with `ncu -i ... --page raw --csv`.\n\n')
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            fh.write(f"## {d['Kernel Name'][:80]}  (launch id {d.get('ID')})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d:
                    fh.write(f'| {k} | {d[k]} | {units[hdr.index(k)]} |\n')
            try:
                t = float(d['gpu__time_duration.sum'].replace(',', ''))
                tu = units[hdr.index('gpu__time_duration.sum')]
                t_s = t * {'ns': 1e-9, 'us': 1e-6, 'ms': 1e-3, 's': 1.0}.get(tu, 1e-9)

                def gb(key):
                    v = float(d[key].replace(',', ''))
                    u = units[hdr.index(key)]
                    return v * {'byte': 1e-9, 'Kbyte': 1e-6, 'Mbyte': 1e-3, 'Gbyte': 1.0}.get(u, 1e-9)
                traffic = gb('dram__bytes_read.sum') + gb('dram__bytes_write.sum')
                fh.write(f'| **DRAM traffic (read+write)** | {traffic:.3f} | GB |\n')
                fh.write(f'| **DRAM GB/s over the launch** | {traffic / t_s:.0f} | GB/s |\n')
            except Exception as e:      # noqa
                fh.write(f'| derived | n/a ({e}) | |\n')
            fh.write('\n')
    print(open(dst).read()[:3000])


if __name__ == '__main__':
    {'launches': launches, 'full': full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])
