# repeat the multi-rank check to look for intermittent failures; logs of failing runs are kept
N=${1:-2}; R=${2:-6}
export TC_DIST_TIMEOUT_S=10
for i in $(seq 1 $R); do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29550+i)) scripts/dist_check.py > gpurun_out/r02_dist_loop_$i.log 2>&1
  rc=$?
  if grep -q "dist_check ok" gpurun_out/r02_dist_loop_$i.log; then echo "run $i ok"; rm gpurun_out/r02_dist_loop_$i.log; else echo "run $i FAILED rc=$rc"; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/r02_dist_loop_$i.log | head -40; fi
done
