# adaptive explicit pairs on a partitioned body at N ranks
N=${1:-2}
export TC_DIST_TIMEOUT_S=10
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 scripts/dropin_partitioned_rk_check.py --json gpurun_out/r02_dropin_partitioned_rk_n$N.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -25
