#!/bin/bash
# Last confirmation on the final tree: smoke(), the default bench line, the reference arm.
set -u
out=gpurun_out/r2last
mkdir -p $out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke exit $?" | tee -a $out/summary.txt
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench default exit $?" | tee -a $out/summary.txt
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err
echo "bench reference exit $?" | tee -a $out/summary.txt
tail -1 $out/smoke.log
python - <<'PY'
import json
for n in ('bench_default','bench_reference'):
    try:
        d=json.loads(open(f'gpurun_out/r2last/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), round(d['ms_per_step'],2), d.get('e2e',{}).get('value'), d.get('config',{}).get('glue',{}).get('mode'),
              d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('share_of_step'), (d.get('cpu_baseline') or {}).get('value'), (d.get('cpu_baseline') or {}).get('cores'),
              {k:round(v['tokens_per_s']) for k,v in ((d.get('decode') or {}).get('graph_decode_batch_sweep') or {}).items() if 'tokens_per_s' in v})
    except Exception as e: print(n, 'failed', e)
PY
