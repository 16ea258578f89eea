"""Turn the raw ncu artefacts in gpurun_out/ into the small, tracked summaries under profiles/.

    python tools/summarize_profiles.py --tag r01

  gpurun_out/prof_*.ncu-rep  ->  profiles/ncu_<name>_<tag>.json  (selected raw metrics per captured launch)
  gpurun_out/launches.csv    ->  profiles/launches_<tag>.json    (per-kernel share of the timed region)
"""
import argparse
import collections
import csv
import glob
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__shared_mem_per_block_dynamic', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_tc_wavefronts_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'sm__cycles_elapsed.max', 'smsp__inst_executed.sum']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(unit, 1)


def summarize_rep(path, tag):
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return None
    hdr, units = rows[0], rows[1]
    recs = []
    for r in rows[2:]:
        d = {'kernel': r[hdr.index('Kernel Name')][:160]}
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                d[k] = dict(value=r[i], unit=units[i])
        if 'dram__bytes_read.sum' in d:
            d['dram_bytes_per_launch'] = (to_bytes(d['dram__bytes_read.sum']['value'], d['dram__bytes_read.sum']['unit']) +
                                          to_bytes(d['dram__bytes_write.sum']['value'], d['dram__bytes_write.sum']['unit']))
        recs.append(d)
    name = os.path.splitext(os.path.basename(path))[0].replace('prof_', '')
    dst = os.path.join(ROOT, 'profiles', f'ncu_{name}_{tag}.json')
    json.dump(dict(source=os.path.basename(path), note='ncu --set full --clock-control none; per-launch, cold cache',
                   launches=recs), open(dst, 'w'), indent=1)
    return dst


def summarize_launches(path, tag):
    lines = [l for l in open(path) if not l.startswith('==')]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in r:
        if len(row) <= vi:
            continue
        try:
            v = float(row[vi].replace(',', ''))
        except ValueError:
            continue
        name = re.sub(r'\(.*', '', row[ki])[:100]
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    ours = sum(v[1] for k, v in agg.items() if 'quip::' in k)
    table = [dict(kernel=k, launches=v[0], total_us=v[1] / 1e3, share=v[1] / tot, avg_us=v[1] / v[0] / 1e3)
             for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    dst = os.path.join(ROOT, 'profiles', f'launches_{tag}.json')
    json.dump(dict(source=os.path.basename(path),
                   note='ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off over the '
                        'timed region of bench.py (serialised, cold cache: compare shares, not absolutes)',
                   total_us=tot / 1e3, quip_kernels_share=ours / tot, kernels=table[:40]), open(dst, 'w'), indent=1)
    return dst


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--tag', default='r01')
    a = ap.parse_args()
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    for rep in sorted(glob.glob(os.path.join(ROOT, 'gpurun_out', 'prof_*.ncu-rep'))):
        print(summarize_rep(rep, a.tag))
    lp = os.path.join(ROOT, 'gpurun_out', 'launches.csv')
    if os.path.exists(lp):
        print(summarize_launches(lp, a.tag))
