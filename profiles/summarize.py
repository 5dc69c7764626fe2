"""Helpers to turn ncu exports into the short summaries committed under profiles/.
usage: python profiles/summarize.py launches <launches.csv> | raw <report.ncu-rep> | src <report.ncu-rep> [top]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'sm__cycles_active.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    for i, r in enumerate(rows):
        if 'Kernel Name' in r:
            hdr, start = r, i + 1
            break
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    agg = collections.OrderedDict()
    for r in rows[start:]:
        name = r[ki].split('(CU')[0].split('(const')[0].split('(float')[0][:60]
        v = float(r[vi].replace(',', ''))
        v = {'ns': v / 1e3, 'us': v, 'ms': v * 1e3}.get(r[ui], v)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"{'kernel':60s} {'n':>5s} {'total_us':>12s} {'avg_us':>10s} share")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:60s} {c:5d} {t:12.1f} {t / c:10.1f} {t / tot:6.3f}")


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index('Kernel Name')
    for r in rows[2:]:
        print('==', r[ki][:70])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:85s} {r[i]:>16s} {units[i]}")


def src(rep, top=25):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kern, cur = [], None
    for r in rows:
        if r and r[0] == 'Kernel Name':
            cur = {'name': r[1], 'rows': []}
            kern.append(cur)
        elif r and r[0] == 'Address':
            cur['hdr'] = r
        elif cur is not None and len(r) > 10:
            cur['rows'].append(r)
    for k in kern:
        h = k['hdr']
        si, ii = h.index('# Samples'), h.index('Instructions Executed')
        tot = sum(int(r[si]) for r in k['rows'])
        print('====', k['name'][:60], 'total samples', tot, 'instructions', len(k['rows']))
        best = sorted(enumerate(k['rows']), key=lambda x: -int(x[1][si]))[:top]
        for idx, r in sorted(best):
            st = {h[j]: int(r[j]) for j in range(h.index('stall_barrier'), h.index('stall_wait') + 1) if r[j] not in ('0', '')}
            print(f"{idx:5d} {r[1].strip()[:58]:58s} samples={r[si]:>7s} exec={r[ii]:>9s}", dict(sorted(st.items(), key=lambda kv: -kv[1])[:3]))


if __name__ == '__main__':
    cmd = sys.argv[1]
    if cmd == 'launches':
        launches(sys.argv[2])
    elif cmd == 'raw':
        raw(sys.argv[2])
    else:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
