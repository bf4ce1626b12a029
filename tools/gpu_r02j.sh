#!/bin/bash
# round 2, call J (8 GPUs): the bench line at N = 8 with its strong-scaling legs
O=gpurun_out
mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 5 --warmup 3 > $O/bench_r02j_n8.json 2> $O/bench_r02j_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_r02j_n8.json').read().strip().splitlines()[-1])
    print('N=8 value', d['value'], 'fit', d['fit']['seconds'], d['fit']['nlml_grad_evaluations'], 'c4 eval', d['strong_c4']['eval_step_seconds'], 'c4 fit', d['strong_c4']['fit_seconds'], d['strong_c4']['fit_nlml_grad_evaluations'])
except Exception as e:
    print('no bench line', e)
PY
grep -v "^$" $O/bench_r02j_n8.err | tail -5
