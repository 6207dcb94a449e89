export TC_DIST_TIMEOUT_S=10
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 scripts/dropin_partitioned_check.py --json gpurun_out/r02_dropin_partitioned_n$N.json > gpurun_out/r02_dropin_partitioned_n$N.log 2>&1
grep -v "^W\|^\*\*\*\|OMP_NUM\|Warning\|warn" gpurun_out/r02_dropin_partitioned_n$N.log | grep -B2 -A25 "Traceback\|rank 0/\|ok" | head -60
