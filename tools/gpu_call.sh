#!/bin/bash
# Final validation of the round: whole GPU suite, smoke(), the default bench line and the reference arm.
set -u
out=gpurun_out/r2final2
mkdir -p $out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -x -q > $out/tests.log 2>&1
echo "gpu suite exit $?" | tee -a $out/summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1
echo "smoke exit $?" | tee -a $out/summary.txt
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench default exit $?" | tee -a $out/summary.txt
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err
echo "bench reference exit $?" | tee -a $out/summary.txt
cp gpurun_out/parity_report.jsonl $out/ 2>/dev/null
tail -4 $out/tests.log | cut -c1-250
cat $out/smoke.log | tail -2
python - <<'PY'
import json
for n in ('bench_default','bench_reference'):
    try:
        d=json.loads(open(f'gpurun_out/r2final2/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), round(d['ms_per_step'],2), d.get('e2e',{}).get('value'), d.get('config',{}).get('glue',{}).get('mode'),
              d.get('roofline',{}).get('frac'), d.get('cpu_baseline'), d.get('steps'), d.get('timed_seconds'), (d.get('decode') or {}).get('hbm_frac'),
              ((d.get('decode') or {}).get('graph_decode') or {}).get('tokens_per_s'))
    except Exception as e: print(n, 'failed', e)
PY
tail -3 $out/bench_default.err
