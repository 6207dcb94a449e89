N=${1:-2}
DOF=${2:-2.5e6}
export TC_DIST_TIMEOUT_S=10
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -3
TC_DIST_DIAG=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --dof $DOF --steps 16 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b5.json 2> gpurun_out/r02_b5.err
tail -2 gpurun_out/r02_b5.err | grep -v "OMP\|\*\*\*"
python - <<PY
import json
for f in ["r02_b5"]:
    try:
        txt=[l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        d=json.loads(txt)
        print(f, round(d["value"]/1e6,1), "ms/step", round(d["ms_per_step"],3), "its", d["config"]["iters_per_step"], "late", d["config"].get("late_ghosts"), "ghost_rows", d["config"]["ghost_rows"], "parity", d.get("parity",{}).get("max_rel"), "grid", d["config"]["grid"])
        ph=d["config"].get("phase_ms_rank0")
        its=d["config"]["iters_per_step"]*d["steps"]
        for k in ("block0","last"): print("   ", k, {a: round(1e3*v/its,2) for a,v in ph[k].items()})
    except Exception as e: print(f, "ERR", e)
PY
