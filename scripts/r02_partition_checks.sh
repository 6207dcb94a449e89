# N ranks: the unmodified reference's Network.sim on a row-partitioned body (THETA_PCG, adaptive pairs, W-methods) vs one GPU
N=${1:-2}
export TC_DIST_TIMEOUT_S=10
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516 scripts/dropin_partitioned_check.py --json gpurun_out/r02_dropin_partitioned_n$N.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | grep "rank 0\|ok\|rror\|FAIL" | tail -8
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/dropin_partitioned_rk_check.py --json gpurun_out/r02_dropin_partitioned_rk_n$N.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | grep "rank 0\|ok\|rror\|FAIL" | tail -10
