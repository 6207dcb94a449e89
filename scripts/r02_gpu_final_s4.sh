# full GPU suite + the N=1 bench line of the final code of round 2
timeout -k 10 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r02_gputests_final_s4.log 2>&1; tail -3 gpurun_out/r02_gputests_final_s4.log
timeout -k 10 600 python bench.py > gpurun_out/r02_bench_n1_final_s4.json 2> gpurun_out/r02_bench_n1_final_s4.err; tail -c 300 gpurun_out/r02_bench_n1_final_s4.json; tail -3 gpurun_out/r02_bench_n1_final_s4.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02_bench_n1_final_s4.json") if l.startswith("{")][-1])
print("N=1", round(d["value"]/1e6,1), "ms/step", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "frac_dram", d["roofline"].get("frac_dram"), "e2e", round(d["e2e"]["value"]/1e6,1), "rk45", d.get("rk45_explicit",{}).get("ms_per_attempt"), "parity", d.get("parity",{}).get("max_rel"), "cpu", d.get("cpu_baseline",{}).get("value"))
for k,v in d.get("extra",{}).items():
    print("  ", k, v.get("value") and round(v["value"]/1e6,2), "ms/step", v.get("ms_per_step"), "frac", v.get("roofline_frac") or (v.get("roofline") or {}).get("frac"), v.get("error"))
PY
