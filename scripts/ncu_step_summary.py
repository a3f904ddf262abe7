#!/usr/bin/env python
"""Reduce an `ncu --set full` report of ONE bench step (bench.py --ncu-range) to text the repo can carry.

Runs on the GPU box right after the capture:
  python scripts/ncu_step_summary.py <report.ncu-rep> <workload> <out_prefix>
writes  <out_prefix>_launches.csv   one row per launch: kernel, grid, duration, dram bytes, tensor/dram/sm %
        <out_prefix>_summary.json   per-kernel-name totals (launches, time, dram bytes, time-weighted %)
"""
import csv
import io
import json
import subprocess
import sys

COLS = {
    'dur_ns': 'gpu__time_duration.sum',
    'dram_rd': 'dram__bytes_read.sum',
    'dram_wr': 'dram__bytes_write.sum',
    'tensor_pct': 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
    'tensor_subpipe_pct': 'sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active',
    'dram_pct': 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
    'sm_pct': 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    'l2_pct': 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
    'l1_pct': 'l1tex__throughput.avg.pct_of_peak_sustained_active',
    'regs': 'launch__registers_per_thread',
    'grid': 'launch__grid_size',
    'block': 'launch__block_size',
    'smem_dyn': 'launch__shared_mem_per_block_dynamic',
    'warps_active_pct': 'sm__warps_active.avg.pct_of_peak_sustained_active',
    'ipc': 'sm__inst_executed.avg.per_cycle_elapsed',
    'l2_hit_pct': 'lts__t_sector_hit_rate.pct',
}


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    u = unit.lower()
    mult = {'byte': 1, 'bytes': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9, 'tbyte': 1e12}
    return v * mult.get(u, 1)


def to_ns(v, unit):
    v = float(v.replace(',', ''))
    mult = {'ns': 1, 'nsecond': 1, 'us': 1e3, 'usecond': 1e3, 'ms': 1e6, 'msecond': 1e6, 's': 1e9, 'second': 1e9}
    return v * mult.get(unit.lower(), 1)


def main():
    rep, workload, prefix = sys.argv[1:4]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    # first row = metric names, second = units, then one row per launch
    hdr = None
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr = i
            break
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    idx = {n: j for j, n in enumerate(names)}
    kcol = idx['Kernel Name']
    out_rows = []
    for r in data:
        if len(r) != len(names):
            continue
        full = r[kcol].split('(')[0].replace('void ', '').replace('lumi::', '').strip()
        rec = {'kernel': full.split('<')[0], 'variant': full}
        for k, m in COLS.items():
            if m not in idx or r[idx[m]] in ('', 'n/a'):
                rec[k] = None
                continue
            v, u = r[idx[m]], units[idx[m]]
            if k.startswith('dram_r') or k.startswith('dram_w') or k == 'smem_dyn':
                rec[k] = to_bytes(v, u)
            elif k == 'dur_ns':
                rec[k] = to_ns(v, u)
            else:
                rec[k] = float(v.replace(',', ''))
        out_rows.append(rec)
    with open(prefix + '_launches.csv', 'w', newline='') as f:
        w = csv.DictWriter(f, fieldnames=['kernel', 'variant'] + list(COLS))
        w.writeheader()
        w.writerows(out_rows)
    summ = {}
    for rec in out_rows:
        s = summ.setdefault(rec['kernel'], {'launches': 0, 'time_us': 0.0, 'dram_bytes': 0.0, '_tw': {}})
        d = rec['dur_ns'] or 0.0
        s['launches'] += 1
        s['time_us'] += d / 1e3
        s['dram_bytes'] += (rec['dram_rd'] or 0) + (rec['dram_wr'] or 0)
        for k in ('tensor_pct', 'dram_pct', 'sm_pct', 'l2_pct', 'warps_active_pct', 'ipc'):
            if rec[k] is not None:
                s['_tw'][k] = s['_tw'].get(k, 0.0) + rec[k] * d
    for s in summ.values():
        for k, v in s.pop('_tw').items():
            s[k + '_timeweighted'] = v / (s['time_us'] * 1e3) if s['time_us'] else None
    total = sum(s['time_us'] for s in summ.values())
    for s in summ.values():
        s['share_of_step'] = s['time_us'] / total if total else None
    try:
        full = json.load(open(prefix + '_summary.json'))
    except Exception:
        full = {}
    full[workload] = summ
    json.dump(full, open(prefix + '_summary.json', 'w'), indent=1, sort_keys=True)
    for k, s in sorted(summ.items(), key=lambda kv: -kv[1]['time_us']):
        print('%-32s n=%3d  %9.1f us  share %.3f  dram %8.1f MB  tensor %s' % (
            k[:32], s['launches'], s['time_us'], s['share_of_step'], s['dram_bytes'] / 1e6,
            s.get('tensor_pct_timeweighted')))


if __name__ == '__main__':
    main()
