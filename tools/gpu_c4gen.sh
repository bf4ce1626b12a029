#!/bin/bash
# C4 as BASELINE.json states it: population 512 x 20 generations, N = 4096, D = 8 -- the scoring work of 20 generations
# (fresh random population per generation, 3 restarts, max_iters = 10), on the GPUs of this box.
# Usage: bash tools/gpu_c4gen.sh <n_gpus>
N=${1:-1}
O=gpurun_out
mkdir -p $O
if [ "$N" = "1" ]; then
  timeout 2400 python bench.py --steps 3 --warmup 3 --no-cpu --c4-generations 20 > $O/bench_c4gen_n$N.json 2> $O/bench_c4gen_n$N.err
else
  timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu --c4-generations 20 > $O/bench_c4gen_n$N.json 2> $O/bench_c4gen_n$N.err
fi
echo "rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_c4gen_n$N.json').read().strip().splitlines()[-1])
    print('N=$N', d['strong_c4']['generations'], 'one generation', d['strong_c4']['fit_seconds'])
except Exception as e:
    print('no line', e)
PY
grep -v "^$" $O/bench_c4gen_n$N.err | tail -3
