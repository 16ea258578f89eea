#!/bin/bash
# GPU call 3 of round 2: new kernels (LDLQ, Hessian, vecquant), whole suite, default bench.
set -u
out=gpurun_out/r2c3
mkdir -p $out
rm -f gpurun_out/parity_report.jsonl
timeout 600 python -m pytest tests/test_gpu_quantize.py -m gpu -q > $out/quantize.log 2>&1
echo "quantize tests exit $?" | tee -a $out/summary.txt
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "vecquant or quant3linear" > $out/vecquant.log 2>&1
echo "vecquant tests exit $?" | tee -a $out/summary.txt
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_quantize.py > $out/tests.log 2>&1
echo "gpu suite exit $?" | tee -a $out/summary.txt
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err
echo "bench default exit $?" | tee -a $out/summary.txt
cp gpurun_out/parity_report.jsonl $out/ 2>/dev/null
tail -25 $out/quantize.log
tail -8 $out/vecquant.log
tail -6 $out/tests.log
head -c 300 $out/bench_default.json; echo; tail -3 $out/bench_default.err
