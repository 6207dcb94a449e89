# full bench line at N GPUs (as the driver launches it) + dist_check; outputs under gpurun_out/
N=${1:-8}
export TC_DIST_TIMEOUT_S=10
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py --gpus 1 --steps 8 --warmup 3 > gpurun_out/r02_scale_n1.json 2> gpurun_out/r02_scale_n1.err
else
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/dist_check.py --json gpurun_out/r02_dist_check_n$N.json 2>&1 | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -20
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err
fi
tail -3 gpurun_out/r02_scale_n$N.err
python - <<PY
import json
f="r02_scale_n$N"
try:
    txt=[l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
    d=json.loads(txt)
    print(f, round(d["value"]/1e6,1), "ms/step", round(d["ms_per_step"],3), "its", d["config"]["iters_per_step"], "e2e", round(d["e2e"]["value"]/1e6,1), "parity", d.get("parity",{}).get("max_rel"))
    for k,v in d.get("extra",{}).items():
        if "error" in v: print("  ", k, v); continue
        print("  ", k, round(v["value"]/1e6,2), v.get("unit"), "ms/step", round(v["ms_per_step"],3), "parity", (v.get("parity") or {}).get("max_rel"), "frac", v.get("roofline_frac") or (v.get("roofline") or {}).get("frac"))
    ph=d["config"].get("phase_ms_rank0")
    if ph:
        its=d["config"]["iters_per_step"]*d["steps"]
        print("   last-block us/iter:", {k: round(1e3*v/its,2) for k,v in ph["last"].items()})
        print("   busy/wait:", d["config"]["busy_wait_ms_per_rank"])
except Exception as e: print(f, "ERR", e)
PY
