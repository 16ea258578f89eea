#!/bin/bash
# 8-GPU call: the layer pipeline of BASELINE configs[3] and [4] across one box.
#   gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_pp8.sh'
set -u
N=8
out=gpurun_out/pp8
mkdir -p $out
run() {
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > $out/$name.json 2> $out/$name.err
  echo "$name exit $?" | tee -a $out/summary.txt
  tail -c 1200 $out/$name.json; echo
}
run pp_llama70b --model llama70b --parallelism pp --steps 48 --warmup 8
run pp_opt30b --model opt30b --parallelism pp --steps 48 --warmup 8
grep -v Warning $out/pp_llama70b.err | tail -3
