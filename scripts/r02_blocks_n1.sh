export TC_DIST_TIMEOUT_S=10 RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29561
timeout 300 python scripts/pipe_bench.py --rows-per-gpu 1300000 --variants classic,classic::::296 --blocks --json gpurun_out/r02_blocks_n1.json 2>&1 | grep "^{\|ok\|rror" | cut -c1-400
