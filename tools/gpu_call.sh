#!/bin/bash
# GPU call 4 of round 2: one-launch few-token sides, real-quantisation bench, default bench.
set -u
out=gpurun_out/r2c4
mkdir -p $out
rm -f gpurun_out/parity_report.jsonl
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_decode.py tests/test_gpu_staged.py -m gpu -q > $out/tests.log 2>&1
echo "layer/decode tests exit $?" | tee -a $out/summary.txt
timeout 900 python tools/quantize_bench.py --layers 2 > $out/quantize_bench.json 2> $out/quantize_bench.err
echo "quantize bench exit $?" | tee -a $out/summary.txt
timeout 900 python bench.py --no-cpu-baseline > $out/bench_default.json 2> $out/bench_default.err
echo "bench default exit $?" | tee -a $out/summary.txt
timeout 600 python tools/microbench.py --what decode --graph --out $out/mb_decode.json > $out/mb_decode.log 2>&1
echo "microbench decode exit $?" | tee -a $out/summary.txt
cp gpurun_out/parity_report.jsonl $out/ 2>/dev/null
tail -12 $out/tests.log
cat $out/quantize_bench.json; tail -3 $out/quantize_bench.err
head -c 300 $out/bench_default.json; echo; tail -3 $out/bench_default.err
tail -5 $out/mb_decode.log
