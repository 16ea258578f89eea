#!/bin/bash
# GPU call: decode-kernel durations + full capture of the one-launch side kernel; LDLQ timing after the bank-conflict fix.
set -u
out=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'side_fewtok|pass_fewtok|qgemv|gather_fewtok|skinny' --csv \
    --log-file $out/decode_launches_r02.csv python tools/prof_decode.py > $out/prof_decode.log 2>&1
echo "decode launch list exit $?" | tee -a $out/prof_summary2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:side_fewtok -s 2 -c 4 -o $out/prof_side_fewtok python tools/prof_decode.py > $out/prof_decode2.log 2>&1
echo "side_fewtok capture exit $?" | tee -a $out/prof_summary2.txt
timeout 600 python -m pytest tests/test_gpu_quantize.py tests/test_gpu_staged.py -m gpu -q > $out/r2c7_tests.log 2>&1
echo "tests exit $?" | tee -a $out/prof_summary2.txt
timeout 600 python tools/quantize_bench.py --layers 1 > $out/quantize_bench_r2c7.json 2> $out/quantize_bench_r2c7.err
echo "quantize bench exit $?" | tee -a $out/prof_summary2.txt
tail -5 $out/r2c7_tests.log | cut -c1-200
cat $out/quantize_bench_r2c7.json
grep -v "^==" $out/decode_launches_r02.csv | cut -d, -f5,12- | head -80
