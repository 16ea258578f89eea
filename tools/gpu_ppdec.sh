#!/bin/bash
# Multi-GPU call: the layer pipeline with its token-by-token decode leg (PipelinedDecoder) on N GPUs of one box.
#   gpurun --gpus N --timeout 900 -- 'bash tools/gpu_ppdec.sh N'
set -u
N=${1:-2}
out=gpurun_out/ppdec${N}
mkdir -p $out
run() {  # name, args...
  name=$1; shift
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N "$@" > $out/$name.json 2> $out/$name.err
  echo "$name exit $?" | tee -a $out/summary.txt
  python - $out/$name.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(round(d['value']), d['pipeline']['stage_ms'], d.get('decode'))
except Exception as e:
    print('no line', e)
PY
}
run pp_llama70b --model llama70b --parallelism pp --steps 8 --warmup 3
run pp_opt30b --model opt30b --parallelism pp --steps 8 --warmup 3
tail -4 $out/pp_llama70b.err; tail -4 $out/pp_opt30b.err
