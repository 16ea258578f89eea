#!/bin/bash
set -u
out=gpurun_out/r2c8
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_staged.py -m gpu -q > $out/tests.log 2>&1
echo "tests exit $?" | tee -a $out/summary.txt
timeout 600 python tools/microbench.py --what decode --graph --out $out/mb_decode.json > $out/mb_decode.log 2>&1
echo "microbench decode exit $?" | tee -a $out/summary.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'side_fewtok|pass_fewtok|qgemv|gather_fewtok|skinny' --csv \
    --log-file $out/decode_launches.csv python tools/prof_decode.py > $out/prof_decode.log 2>&1
tail -4 $out/tests.log | cut -c1-200
