#!/bin/bash
# One gpurun call: GPU tests, smoke, bench line, ncu launch list of the SAME bench command, one --set full capture per kernel.
# Usage (from the repo root on the GPU box):  bash tools/ncu_r02.sh <tag>
TAG=${1:-r02z}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -q --timeout 1500 > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
tail -3 $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke_$TAG.log
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-fit"
$BENCH > $O/bench_plain_$TAG.json 2> $O/bench_plain_$TAG.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_$TAG.csv $BENCH > $O/ncu_launch_$TAG.log 2>&1
PROBE="python tools/perf_probe.py 2048 32 1"
$PROBE > $O/probe_plain_$TAG.log 2>&1 || exit 1
cap() {  # name, kernel regex, skip, command...
  local name=$1 re=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c 1 -f -o $O/prof_${name}_$TAG "$@" > $O/ncu_${name}_$TAG.log 2>&1
}
cap kinv kinv64_tma_kernel 0 $PROBE
cap upd chol_update64_tma_kernel 3 $PROBE
cap vupd inv_update64_tma_kernel 3 $PROBE
cap grad poly_grad_kernel 0 $PROBE
cap build poly_build_kernel 0 $PROBE
cap potrf potrf_diag 3 $PROBE
C5="python tools/perf_c5.py 16384"
$C5 > $O/c5_$TAG.log 2>&1
cap c5upd chol_update64_tma_kernel 4 $C5          # the first K = 512 trailing update of the N = 16384 Cholesky (29 040 CTAs)
C3="python tools/bench_configs.py --configs c3 --cpu-budget 0"
$C3 > $O/c3_$TAG.log 2>&1
cap hpair hellinger_pair 1 $C3
cap hpre hellinger_precompute_kernel 0 $C3
python bench.py --steps 5 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err
tail -c 400 $O/bench_$TAG.json
ls $O | wc -l
