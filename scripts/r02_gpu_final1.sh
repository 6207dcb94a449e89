# full GPU suite + the N=1 bench line of the final round-2 code
timeout -k 10 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_final.log 2>&1; tail -3 gpurun_out/r02_gputests_final.log
timeout -k 10 900 python bench.py > gpurun_out/r02_bench_n1_final.json 2> gpurun_out/r02_bench_n1_final.err; tail -c 600 gpurun_out/r02_bench_n1_final.json; tail -3 gpurun_out/r02_bench_n1_final.err
