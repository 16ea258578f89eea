#!/bin/bash
# Multi-GPU call: the layer pipeline (BASELINE configs[3], [4]) on N GPUs of one box.
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_pp.sh N'
set -u
N=${1:-2}
out=gpurun_out/pp${N}
mkdir -p $out
run() {  # name, args...
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N "$@" > $out/$name.json 2> $out/$name.err
  echo "$name exit $?" | tee -a $out/summary.txt
  tail -c 1500 $out/$name.json; echo
}
run pp_llama70b --model llama70b --parallelism pp --steps 16 --warmup 3
run pp_opt30b --model opt30b --parallelism pp --steps 16 --warmup 3
run pp_llama7b --model llama7b --parallelism pp --steps 16 --warmup 3
run dp_llama7b --steps 8 --warmup 3 --no-decode --no-cpu-baseline
tail -5 $out/pp_llama70b.err
