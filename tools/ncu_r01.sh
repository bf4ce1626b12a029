#!/bin/bash
# One gpurun call: GPU tests, bench line, ncu launch list of the SAME bench command, one --set full capture per kernel.
# Usage (from the repo root on the GPU box):  bash tools/ncu_r01.sh <tag>
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke_$TAG.log
tail -3 $O/pytest_gpu_$TAG.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-fit"
$BENCH > $O/bench_plain_$TAG.json 2> $O/bench_plain_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_$TAG.csv $BENCH > $O/ncu_launch_$TAG.log 2>&1
PROBE="python tools/perf_probe.py 2048 32 1"
$PROBE > $O/probe_plain_$TAG.log 2>&1 || exit 1
cap() {  # name, kernel regex, skip
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_$1_$TAG $PROBE > $O/ncu_$1_$TAG.log 2>&1
}
cap upd chol_update64_kernel 3    # 4th launch of the Cholesky update kernel
cap vupd inv_update64_kernel 3
cap kinv kinv64_kernel 0
cap trsm chol_trsm_kernel 2
cap grad poly_grad_kernel 0
cap build poly_build_kernel 0
cap potrf potrf_diag 3
python bench.py --steps 5 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
tail -c 600 $O/bench_$TAG.json
ls -la $O | tail -20
