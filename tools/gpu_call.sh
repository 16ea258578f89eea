#!/bin/bash
# Folded gathers of the 4096-wide sides: tests, then the default bench.
set -u
out=gpurun_out/r2fold
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_staged.py tests/test_gpu_glue.py -m gpu -q > $out/tests.log 2>&1
echo "tests exit $?" | tee -a $out/summary.txt
timeout 900 python bench.py --no-cpu-baseline > $out/bench_default.json 2> $out/bench_default.err
echo "bench default exit $?" | tee -a $out/summary.txt
QUIP_FOLD_GATHERS=mlp timeout 600 python bench.py --no-decode --no-cpu-baseline > $out/bench_mlp_only.json 2> $out/bench_mlp_only.err
echo "bench (mlp-only folding) exit $?" | tee -a $out/summary.txt
tail -8 $out/tests.log | cut -c1-250
python - <<'PY'
import json
for n in ('bench_default','bench_mlp_only'):
    try:
        d=json.loads(open(f'gpurun_out/r2fold/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value']), round(d['ms_per_step'],2), d['config']['glue'].get('mode'), d['config']['glue'].get('rel_err_vs_hf_layers'), d['config']['glue'].get('nll_rel_diff_vs_hf_glue'))
    except Exception as e: print(n, 'failed', e)
PY
tail -3 $out/bench_default.err
