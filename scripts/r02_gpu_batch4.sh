# small-slab experiments at N=2 (1.25 M rows per GPU ~ the N=8 regime of the 10 M mesh) + single-GPU sizes
N=2
export TC_DIST_TIMEOUT_S=10
for BF in 0 1; do
  TC_DIST_BND_FIRST=$BF timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --dof 2.5e6 --steps 16 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b4_bf$BF.json 2> gpurun_out/r02_b4_bf$BF.err
  tail -2 gpurun_out/r02_b4_bf$BF.err | grep -v OMP
done
for D in 1e6 2.5e6 10e6; do
  timeout 600 python bench.py --dof $D --steps 16 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b4_n1_$D.json 2> gpurun_out/r02_b4_n1.err
done
python - <<PY
import json
for f in ["r02_b4_n1_1e6", "r02_b4_n1_2.5e6", "r02_b4_n1_10e6", "r02_b4_bf0", "r02_b4_bf1"]:
    try:
        txt=[l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1]
        d=json.loads(txt)
        print(f, round(d["value"]/1e6,1), "ms/step", round(d["ms_per_step"],3), "its", d["config"]["iters_per_step"], "frac", round(d["roofline"]["frac"],3), "parity", d.get("parity",{}).get("max_rel"), "grid", d["config"].get("grid"))
        ph=d["config"].get("phase_ms_rank0")
        its=d["config"]["iters_per_step"]*d["steps"]
        if ph:
            for k in ("block0","last"): print("   ", k, {a: round(1e3*v/its,2) for a,v in ph[k].items()})
        else:
            print("   ", {a: round(1e3*v/its,2) for a,v in d["roofline"]["phase_ms_block0"].items()})
    except Exception as e: print(f, "ERR", e)
PY
