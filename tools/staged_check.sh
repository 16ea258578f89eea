#!/bin/bash
# First GPU call of the next round: verify and measure everything that was written without a GPU.
#   gpurun --timeout 1500 -- 'bash tools/staged_check.sh'
# Results land in gpurun_out/staged/.
set -u
out=gpurun_out/staged
mkdir -p $out
export QUIP_TEST_STAGED=1
timeout 600 python -m pytest tests/test_gpu_staged.py -m gpu -x -q > $out/staged_tests.log 2>&1
echo "staged tests exit $?" | tee -a $out/summary.txt
timeout 300 python tools/microbench.py --what glue --out $out/glue.json > $out/glue.log 2>&1
echo "glue microbench exit $?" | tee -a $out/summary.txt
timeout 420 python tools/microbench.py --what stack --out $out/stack.json > $out/stack.log 2>&1
echo "stack microbench exit $?" | tee -a $out/summary.txt
timeout 300 python tools/microbench.py --what batchdecode --out $out/batchdecode.json > $out/batchdecode.log 2>&1
echo "batched-decode microbench exit $?" | tee -a $out/summary.txt
QUIP_FUSED_LAYER=0 timeout 420 python bench.py --steps 8 --warmup 3 --no-decode --no-cpu-baseline > $out/bench_hf_glue.json 2> $out/bench_hf_glue.err
QUIP_FUSED_LAYER=1 timeout 420 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > $out/bench_fused_glue.json 2> $out/bench_fused_glue.err
python - <<'PY' | tee -a gpurun_out/staged/summary.txt
import json
for n in ('bench_hf_glue', 'bench_fused_glue'):
    try:
        d = json.loads(open(f'gpurun_out/staged/{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value']), 'tok/s', round(d['ms_per_step'], 2), 'ms/step', 'e2e', round(d['e2e']['value']),
              'decode', {k: round(v.get('tokens_per_s', 0)) for k, v in d.get('decode', {}).items() if isinstance(v, dict) and 'tokens_per_s' in v})
    except Exception as e:
        print(n, 'failed:', e)
PY
tail -5 $out/staged_tests.log
