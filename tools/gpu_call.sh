#!/bin/bash
# Whole Llama-2-7B-shaped model (32 decoder layers, random init) quantized on the GPU, packed and evaluated.
set -u
out=gpurun_out/r2quant
mkdir -p $out
timeout 2400 python tools/quantize_bench.py --layers 32 --calib 4 --eval 2 > $out/quantize_bench_32.json 2> $out/quantize_bench_32.err
echo "quantize bench (32 layers) exit $?" | tee -a $out/summary.txt
cat $out/quantize_bench_32.json; tail -3 $out/quantize_bench_32.err | cut -c1-300
