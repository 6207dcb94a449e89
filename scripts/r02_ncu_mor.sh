# ncu evidence for the fused batched-MOR kernel: launch list of the bench command + one full capture (1 GPU)
CMD="python bench.py --workload mor --steps 5 --warmup 3 --no-cpu"
$CMD > gpurun_out/r02_mor_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:tc_mor -c 200 --csv --log-file gpurun_out/r02_launches_mor.csv $CMD > gpurun_out/r02_mor_ncu_list.log 2>&1
$CMD > gpurun_out/r02_mor_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_mor_fused_kernel -s 1 -c 1 -o gpurun_out/r02_mor_fused $CMD > gpurun_out/r02_mor_ncu_full.log 2>&1
tail -2 gpurun_out/r02_mor_ncu_plain.log | cut -c1-400; tail -3 gpurun_out/r02_mor_ncu_list.log | cut -c1-200; tail -3 gpurun_out/r02_mor_ncu_full.log | cut -c1-200; ls -la gpurun_out/r02_mor_fused* gpurun_out/r02_launches_mor.csv
