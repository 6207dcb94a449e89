# ncu evidence for the final round-2 PCG kernel (index-ahead SpMV, merged x update): launch list of the bench command +
# one full capture of the TIMED launch (1 GPU).  --steps 1 --warmup 1: the captured launch (#2) is the timed step, so the
# plain run's iters_per_step is its iteration count.
CMD="python bench.py --steps 1 --warmup 1 --no-extra --no-cpu"
$CMD > gpurun_out/r02s4_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02s4_launches_bench_10m.csv $CMD > gpurun_out/r02s4_ncu_list.log 2>&1
$CMD > gpurun_out/r02s4_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_pcg_coop_kernel -s 1 -c 1 -o gpurun_out/r02s4_pcg $CMD > gpurun_out/r02s4_ncu_full.log 2>&1
tail -2 gpurun_out/r02s4_ncu_plain.log | cut -c1-400; tail -3 gpurun_out/r02s4_ncu_list.log | cut -c1-200; tail -3 gpurun_out/r02s4_ncu_full.log | cut -c1-200; ls -la gpurun_out/r02s4_pcg* gpurun_out/r02s4_launches_bench_10m.csv
