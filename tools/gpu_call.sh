#!/bin/bash
# GPU call 5 of round 2: reworked one-launch few-token sides.
set -u
out=gpurun_out/r2c5
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_decode.py -m gpu -q > $out/tests.log 2>&1
echo "layer/decode tests exit $?" | tee -a $out/summary.txt
timeout 600 python tools/microbench.py --what decode --graph --out $out/mb_decode.json > $out/mb_decode.log 2>&1
echo "microbench decode exit $?" | tee -a $out/summary.txt
timeout 900 python bench.py --steps 4 --no-cpu-baseline > $out/bench_default.json 2> $out/bench_default.err
echo "bench default exit $?" | tee -a $out/summary.txt
tail -12 $out/tests.log | cut -c1-200
grep -c qlinear_forward $out/mb_decode.log
