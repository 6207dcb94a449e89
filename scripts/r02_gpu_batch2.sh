# barrier variants x block shapes (round-2 tuning run; outputs kept under gpurun_out/)
python scripts/barrier_bench.py gpurun_out/r02_barrier_bench_v3.json 2>&1 | tail -8
for T in 256 512 1024; do
  for D in 1e6 2.5e6 10e6; do
    TC_PCG_THREADS=$T python bench.py --dof $D --steps 8 --warmup 3 --no-extra --no-cpu > gpurun_out/r02_b2_${T}_${D}.json 2> gpurun_out/r02_b2_err.log || tail -3 gpurun_out/r02_b2_err.log
  done
done
python - <<'PY'
import json
for T in (256,512,1024):
    for D in ("1e6","2.5e6","10e6"):
        try:
            d=json.load(open(f"gpurun_out/r02_b2_{T}_{D}.json"))
            ph=d["roofline"]["phase_ms_block0"]; its=d["config"]["iters_per_step"]*d["steps"]
            print(T, D, round(d["value"]/1e6,1), round(d["roofline"]["frac"],3), {k: round(1e3*v/its,2) for k,v in ph.items()})
        except Exception as e: print(T, D, "ERR", e)
PY
python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
