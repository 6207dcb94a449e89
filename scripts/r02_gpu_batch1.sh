set -x
python bench.py --dof 1e6 --steps 16 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b1_1m.json 2> gpurun_out/r02_b1_1m.err
python bench.py --dof 2.5e6 --steps 8 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b1_2p5m.json 2>> gpurun_out/r02_b1_1m.err
python bench.py --steps 8 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b1_10m.json 2>> gpurun_out/r02_b1_1m.err
RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29511 python scripts/dist_check.py > gpurun_out/r02_b1_distcheck_w1.log 2>&1
TC_FORCE_DIST=1 python bench.py --steps 8 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b1_dist_w1_10m.json 2> gpurun_out/r02_b1_dist_w1.err
tail -3 gpurun_out/r02_b1_distcheck_w1.log; tail -3 gpurun_out/r02_b1_dist_w1.err; tail -3 gpurun_out/r02_b1_1m.err
python - <<'PY'
import json
for f in ["r02_b1_1m","r02_b1_2p5m","r02_b1_10m","r02_b1_dist_w1_10m"]:
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"]/1e6, d["ms_per_step"], d["config"]["iters_per_step"], d["roofline"]["frac"], d["roofline"].get("phase_ms_block0") or d["config"].get("phase_ms_rank0",{}).get("block0"))
    except Exception as e: print(f, "ERR", e)
PY
