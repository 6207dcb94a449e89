# N ranks: both PCG formulations against the single-GPU solver (dist_check), then classic vs single-reduction vs pipelined
N=${1:-2}
export TC_DIST_TIMEOUT_S=10
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/dist_check.py --json gpurun_out/r02_dist_check_modes_n$N.json > gpurun_out/r02_dist_check_modes_n$N.log 2>&1
grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r02_dist_check_modes_n$N.log | grep -B2 -A12 "Traceback\|rror\|ssert" | head -60
grep "cuboid\|csr body\|dist_check ok" gpurun_out/r02_dist_check_modes_n$N.log | head -14
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 scripts/pipe_bench.py --rows-per-gpu ${2:-1300000} --variants ${3:-classic,sr:232,sr:316,pipe:316} --json gpurun_out/r02_pipe_n${N}.json 2>&1 | grep "^{\|ok\|rror" | tail -12
