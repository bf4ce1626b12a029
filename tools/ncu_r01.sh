set -e
CMD="python tools/perf_probe.py 2048 32 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu0.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 60 -c 1 -f -o gpurun_out/prof_kinv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 1 -c 1 -f -o gpurun_out/prof_syrk $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:grad_kernel -c 1 -f -o gpurun_out/prof_grad $CMD > gpurun_out/ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:build_k_kernel -c 1 -f -o gpurun_out/prof_build $CMD > gpurun_out/ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:potrf_diag -s 3 -c 1 -f -o gpurun_out/prof_potrf $CMD > gpurun_out/ncu5.log 2>&1
tail -3 gpurun_out/plain.log; ls -la gpurun_out
