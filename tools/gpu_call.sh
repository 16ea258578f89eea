#!/bin/bash
# One GPU call: whole -m gpu suite (staged tests included), the glue bisect, short benches with either glue.
#   gpurun --timeout 1500 -- 'bash tools/gpu_call.sh'
set -u
out=gpurun_out/r2c1
mkdir -p $out
rm -f gpurun_out/parity_report.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
QUIP_TEST_STAGED=1 timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_staged.py > $out/tests.log 2>&1
echo "gpu tests exit $?" | tee -a $out/summary.txt
QUIP_TEST_STAGED=1 timeout 600 python -m pytest tests/test_gpu_staged.py -m gpu -q > $out/staged.log 2>&1
echo "staged tests exit $?" | tee -a $out/summary.txt
timeout 300 python tools/glue_bisect.py > $out/glue_bisect.json 2> $out/glue_bisect.err
echo "glue bisect exit $?" | tee -a $out/summary.txt
QUIP_FUSED_LAYER=1 timeout 420 python bench.py --steps 8 --warmup 3 --no-decode --no-cpu-baseline > $out/bench_fused.json 2> $out/bench_fused.err
echo "bench fused exit $?" | tee -a $out/summary.txt
timeout 420 python bench.py --steps 8 --warmup 3 --no-decode --no-cpu-baseline > $out/bench_auto.json 2> $out/bench_auto.err
echo "bench auto exit $?" | tee -a $out/summary.txt
cp gpurun_out/parity_report.jsonl $out/ 2>/dev/null
tail -15 $out/tests.log
tail -8 $out/staged.log
head -c 1500 $out/bench_fused.json
