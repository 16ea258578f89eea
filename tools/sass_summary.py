"""Count the Blackwell-specific SASS mnemonics per kernel of libquip_b200.so (profiles/sass_summary_<tag>.json):
UTC*MMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG / UTMASTG / UBLKCP (TMA), HMMA / IMMA (mma.sync), LDGSTS (cp.async),
LDSM (ldmatrix).  Runs anywhere cuobjdump is installed (no GPU needed):  python tools/sass_summary.py --tag r02"""
import argparse
import collections
import json
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = ['UTCHMMA', 'UTCQMMA', 'UTCIMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'UTCBAR', 'HMMA', 'IMMA', 'LDGSTS', 'LDSM',
            'SYNCS', 'MEMBAR', 'ACQBULK', 'UCGABAR']


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--tag', default='r02')
    a = ap.parse_args()
    lib = os.path.join(ROOT, 'quip_b200', 'libquip_b200.so')
    out = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
            cur = kernels.setdefault(name[:160], collections.Counter())
            continue
        if cur is None:
            continue
        m = re.search(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
        if m:
            op = m.group(1)
            cur['instructions'] += 1
            for p in PATTERNS:
                if op.startswith(p):
                    cur[p + ('.2CTA' if '.2CTA' in op else '')] += 1
    rows = [dict(kernel=k, **v) for k, v in kernels.items()]
    tc = [r for r in rows if any(k.startswith('UTC') and k != 'UTCBAR' for k in r)]
    dst = os.path.join(ROOT, 'profiles', f'sass_summary_{a.tag}.json')
    json.dump(dict(source='cuobjdump -sass quip_b200/libquip_b200.so (sm_100a)', kernels=len(rows),
                   kernels_with_tcgen05_mma=[r['kernel'] for r in tc], per_kernel=rows), open(dst, 'w'), indent=1)
    print(dst, len(rows), 'kernels;', len(tc), 'with UTC*MMA')
    for r in tc:
        print({k: v for k, v in r.items() if k != 'instructions'})


if __name__ == '__main__':
    main()
