# single GPU: correctness of both PCG formulations of the dist kernel (world 1), then classic vs pipelined timings
export TC_DIST_TIMEOUT_S=10
export RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29531
timeout 300 python scripts/dist_check.py 2>&1 | tail -8
timeout 600 python scripts/pipe_bench.py --rows-per-gpu 1300000 --variants classic,sr,sr:316,sr:232,sr:416,sr:132,sr:232:0,pipe:316 --json gpurun_out/r02_pipe_n1_1m3.json 2>&1 | tail -12
timeout 600 python scripts/pipe_bench.py --rows-per-gpu 2560000 --variants classic,sr,sr:316,sr:232 --json gpurun_out/r02_pipe_n1_2m5.json 2>&1 | tail -6
timeout 600 python scripts/pipe_bench.py --rows-per-gpu 10350000 --variants classic,sr,sr:316,sr:232,sr:232:0 --json gpurun_out/r02_pipe_n1_10m.json 2>&1 | tail -6
