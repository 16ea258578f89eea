#!/bin/bash
# Data-parallel bench under torchrun exactly as the driver launches it.  usage: gpu_dp.sh N
set -u
N=${1:-4}
out=gpurun_out/dp${N}
mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 8 --warmup 3 > $out/bench.json 2> $out/bench.err
echo "dp$N exit $?" | tee -a $out/summary.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --impl reference --gpus $N --steps 8 --warmup 3 > $out/bench_reference.json 2> $out/bench_reference.err
echo "reference arm under torchrun exit $?" | tee -a $out/summary.txt
python - <<PY
import json
d=json.loads(open('$out/bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['config']['glue']['mode'], d['clocks'])
r=json.loads(open('$out/bench_reference.json').read().strip().splitlines()[-1])
print('reference', round(r['value'],1), r['steps'], r['cpu_baseline']['cores'])
PY
grep -v "Warning\|warn\|^\*\|OMP_NUM" $out/bench.err | tail -5
