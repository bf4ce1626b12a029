#!/bin/bash
# round 2, call A: all GPU tests (no -x: every failure wanted), smoke, baseline bench line, C3 timing + ncu baseline of the
# Hellinger kernels before their rewrite.
TAG=${1:-r02a}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $O/smi_$TAG.txt
python -m pytest tests -m gpu -q --timeout 1200 > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu_$TAG.log
tail -25 $O/pytest_gpu_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke_$TAG.log
python bench.py --steps 5 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
tail -c 1500 $O/bench_$TAG.json
python tools/bench_configs.py --configs c3 --cpu-budget 5 --out $O/configs_c3_$TAG.json > $O/configs_c3_$TAG.log 2>&1; echo "c3 rc=$?"
tail -5 $O/configs_c3_$TAG.log
C3="python tools/bench_configs.py --configs c3 --cpu-budget 0"
ncu --set full --clock-control none --import-source on -k regex:hellinger_pair_kernel -c 1 -f -o $O/prof_hpair_$TAG $C3 > $O/ncu_hpair_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hellinger_precompute_kernel -c 1 -f -o $O/prof_hpre_$TAG $C3 > $O/ncu_hpre_$TAG.log 2>&1
ls -la $O | tail -20
