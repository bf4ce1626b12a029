#!/bin/bash
# round 2, call I (2 GPUs): the bench line at N = 2 with its strong-scaling legs
O=gpurun_out
mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_r02i_n2.json 2> $O/bench_r02i_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_r02i_n2.json').read().strip().splitlines()[-1])
    print('N=2 value', d['value'], 'fit', d['fit']['seconds'], d['fit']['nlml_grad_evaluations'], 'c4 eval', d['strong_c4']['eval_step_seconds'], 'c4 fit', d['strong_c4']['fit_seconds'], d['strong_c4']['fit_nlml_grad_evaluations'])
except Exception as e:
    print('no bench line', e)
PY
grep -v "^$" $O/bench_r02i_n2.err | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench_ref_r02i_n2.json 2> $O/bench_ref_r02i_n2.err; echo "bench ref n2 rc=$?"; cut -c1-300 $O/bench_ref_r02i_n2.json
