#!/bin/bash
# Kronecker structure + OPT graph decode: new tests, the structure micro-benchmark, the two bench lines.
set -u
out=gpurun_out/r2kron
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_headline.py tests/test_gpu_decode.py -m gpu -q -k "kronecker or opt_graph" > $out/tests.log 2>&1
echo "tests exit $?" | tee -a $out/summary.txt
timeout 300 python tools/microbench.py --what structure --graph --out $out/mb_structure.json > $out/mb_structure.log 2>&1
echo "microbench exit $?" | tee -a $out/summary.txt
timeout 600 python bench.py --incoh kron --no-cpu-baseline > $out/bench_kron.json 2> $out/bench_kron.err
echo "bench kron exit $?" | tee -a $out/summary.txt
timeout 400 python bench.py --model opt1.3b --no-cpu-baseline > $out/bench_opt13b.json 2> $out/bench_opt13b.err
echo "bench opt1.3b exit $?" | tee -a $out/summary.txt
tail -5 $out/tests.log | cut -c1-250
grep qlinear_forward $out/mb_structure.log | cut -c1-200
python - <<'PY'
import json
for n in ('bench_kron','bench_opt13b'):
    try:
        d=json.loads(open(f'gpurun_out/r2kron/{n}.json').read().strip().splitlines()[-1])
        dec=d.get('decode',{})
        print(n, round(d['value']), round(d['ms_per_step'],2), d['config']['glue'].get('mode'), d['roofline']['frac'],
              {k:(round(v['tokens_per_s']) if isinstance(v,dict) and 'tokens_per_s' in v else None) for k,v in dec.items() if k in ('hf_decode','graph_decode')},
              dec.get('hbm_frac'), dec.get('ms'), dec.get('error'))
    except Exception as e: print(n, 'failed', e)
PY
tail -3 $out/bench_kron.err; tail -3 $out/bench_opt13b.err
