for v in 0 1; do
TC_RHS_COL16=$v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-extra > gpurun_out/r02_rhs16_$v.json 2> gpurun_out/r02_rhs16_$v.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02_rhs16_$v.json") if l.startswith("{")][-1])
print("TC_RHS_COL16=$v value", round(d["value"]/1e6,1), "rk45", d.get("rk45_explicit"))
PY
done
