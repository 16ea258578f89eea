#!/bin/bash
# Few-token passes with all token loads requested up front: whole GPU suite, micro-benchmark, bench line.
set -u
out=gpurun_out/r2fewtok
mkdir -p $out
rm -f gpurun_out/parity_report.jsonl
timeout 1200 python -m pytest tests -m gpu -q > $out/tests.log 2>&1
echo "gpu suite exit $?" | tee -a $out/summary.txt
timeout 300 python tools/microbench.py --what skinny --graph --out $out/mb_skinny.json > $out/mb_skinny.log 2>&1
echo "microbench exit $?" | tee -a $out/summary.txt
timeout 600 python bench.py --no-cpu-baseline > $out/bench.json 2> $out/bench.err
echo "bench exit $?" | tee -a $out/summary.txt
cp gpurun_out/parity_report.jsonl $out/ 2>/dev/null
grep -n "FAILED\|passed\|failed" $out/tests.log | tail -8 | cut -c1-250
grep "qlinear_forward" $out/mb_skinny.log | cut -c1-120
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2fewtok/bench.json').read().strip().splitlines()[-1])
print(round(d['value']), d['config']['glue']['mode'], d['roofline']['frac'], d['roofline'].get('share_of_step'))
dec=d['decode']
print({k:round(v['tokens_per_s']) for k,v in dec['graph_decode_batch_sweep'].items() if 'tokens_per_s' in v}, dec.get('hbm_frac'), dec.get('ms'), round(dec['graph_decode']['tokens_per_s']))
PY
