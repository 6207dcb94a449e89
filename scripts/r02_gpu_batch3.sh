# multi-GPU validation + first strong-scaling numbers (N = $1)
N=${1:-2}
export TC_DIST_TIMEOUT_S=5
RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29511 timeout 300 python scripts/dist_check.py 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py --json gpurun_out/r02_dist_check_n$N.json 2>&1 | grep -v "^W\|^\*\*\*" | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b3_n$N.json 2> gpurun_out/r02_b3_n$N.err
tail -3 gpurun_out/r02_b3_n$N.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b3_n1.json 2> gpurun_out/r02_b3_n1.err
python - <<PY
import json
for f in ["r02_b3_n1", "r02_b3_n$N"]:
    try:
        txt=[l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        d=json.loads(txt)
        print(f, round(d["value"]/1e6,1), "ms/step", round(d["ms_per_step"],3), "its", d["config"]["iters_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), "parity", d.get("parity",{}).get("max_rel"), d["config"].get("busy_wait_ms_per_rank"))
        ph=d["config"].get("phase_ms_rank0")
        if ph:
            its=d["config"]["iters_per_step"]*d["steps"]
            print("   last-block us/iter:", {k: round(1e3*v/its,2) for k,v in ph["last"].items()})
    except Exception as e: print(f, "ERR", e)
PY
