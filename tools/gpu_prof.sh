#!/bin/bash
# Profiling call of round 2 (one GPU): launch list of the timed step, ncu --set full of the side kernels and of the new kernels.
set -u
out=gpurun_out
mkdir -p $out
# 1. every launch of two timed steps with its device time (graph nodes are profiled as kernels)
QUIP_PROFILE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $out/launches.csv python bench.py --steps 2 --warmup 3 --no-decode --no-cpu-baseline > $out/bench_ncu.log 2> $out/bench_ncu.err
echo "launch list exit $?" | tee -a $out/prof_summary.txt
# 2. every kernel of a 2048-token forward at the three layer shapes (the pack / plan kernels of the set-up come first: the
#    warm-up pass launches 3 + 5 + 5 library kernels, preceded by 3 pack kernels)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'side_fused|gather_kernel|pass_|qgemm_tc_kernel|rowsum' -s 13 -c 13 \
    -o $out/prof_sides_r02 python tools/prof_sides.py > $out/prof_sides.log 2>&1
echo "sides profile exit $?" | tee -a $out/prof_summary.txt
# 3. round-2 kernels
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'hessian|ldlq_block|greedy_block|side_fewtok|vecquant' -c 30 \
    -o $out/prof_r2new python tools/prof_new_kernels.py > $out/prof_new.log 2>&1
echo "new kernels profile exit $?" | tee -a $out/prof_summary.txt
ls -la $out/*.ncu-rep
