#!/bin/bash
# round 2, call E: TMA 4-CTA variant, all GPU tests, the new bench line, ncu evidence (C5 Cholesky update at N = 16384, launch list)
O=gpurun_out
mkdir -p $O
timeout 300 python tools/tma_probe.py 2048 256 > $O/tma_probe_c2_e.log 2>&1; echo "tma c2 rc=$?"; tail -6 $O/tma_probe_c2_e.log
python -m pytest tests -m gpu -q --timeout 1500 > $O/pytest_gpu_r02e.log 2>&1; echo "pytest rc=$?"; tail -6 $O/pytest_gpu_r02e.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke_r02e.log 2>&1; echo "smoke rc=$?"
python bench.py --steps 5 --warmup 3 > $O/bench_r02e.json 2> $O/bench_r02e.err; echo "bench rc=$?"; tail -c 3000 $O/bench_r02e.json; tail -5 $O/bench_r02e.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_r02e.json 2> $O/bench_ref_r02e.err; echo "bench ref rc=$?"; tail -c 600 $O/bench_ref_r02e.json
C5="python tools/perf_c5.py 16384"
$C5 > $O/c5_r02e.log 2>&1; tail -4 $O/c5_r02e.log
ncu --set full --clock-control none --import-source on -k regex:chol_update64 -s 40 -c 1 -f -o $O/prof_c5_upd_r02e $C5 > $O/ncu_c5_upd_r02e.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_c5_r02e.csv $C5 > $O/ncu_c5_launch_r02e.log 2>&1
ls -la $O | tail -15
