# N ranks: partitioned adaptive schemes check, then dist_check + the full bench line (scripts/r02_gpu_scale.sh)
N=${1:-8}
bash scripts/r02_gpu_batch9.sh $N
bash scripts/r02_gpu_scale.sh $N
