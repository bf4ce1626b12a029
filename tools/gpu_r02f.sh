#!/bin/bash
# round 2, call F (2 GPUs): the sharded scoring seam over NCCL (bit-identical to one GPU) and the bench line at N = 2 with its
# strong-scaling legs; first the per-warp zero skipping of the TMA kernels on one GPU (bit-identical, timing)
O=gpurun_out
mkdir -p $O
timeout 300 python tools/tma_probe.py 2048 256 > $O/tma_probe_c2_f.log 2>&1; echo "tma c2 rc=$?"; tail -6 $O/tma_probe_c2_f.log
timeout 600 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_lml.py tests/test_gpu_potrf.py -q > $O/pytest_r02f.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_r02f.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/sharded_fit_check.py > $O/sharded_fit_r02f.log 2>&1; echo "sharded rc=$?"; tail -3 $O/sharded_fit_r02f.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_r02f_n2.json 2> $O/bench_r02f_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_r02f_n2.json').read().strip().splitlines()[-1])
    print('N=2 value', d['value'], 'fit', d['fit']['seconds'], 'c4 eval', d['strong_c4']['eval_step_seconds'], 'c4 fit', d['strong_c4']['fit_seconds'])
except Exception as e:
    print('no bench line', e)
PY
tail -5 $O/bench_r02f_n2.err
