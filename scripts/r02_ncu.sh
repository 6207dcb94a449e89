# ncu evidence for the round-2 PCG kernel: launch list of the bench command + one full capture (1 GPU)
CMD="python bench.py --steps 2 --warmup 1 --no-extra --no-cpu"
$CMD > gpurun_out/r02_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_10m.csv $CMD > gpurun_out/r02_ncu_list.log 2>&1
$CMD > gpurun_out/r02_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_pcg_coop_kernel -s 1 -c 1 -o gpurun_out/r02_pcg $CMD > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_plain.log | cut -c1-300; tail -3 gpurun_out/r02_ncu_list.log | cut -c1-200; tail -3 gpurun_out/r02_ncu_full.log | cut -c1-200; ls -la gpurun_out/r02_pcg* gpurun_out/r02_launches_bench_10m.csv
