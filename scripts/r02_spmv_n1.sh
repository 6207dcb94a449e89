# single GPU: dist kernel (world 1) at two sizes + the bench line of the solver path; the development series of the
# product (row tail, indices ahead, profiler, block shape via TC_PCG_THREADS=768) in profiles/README.md came from this script
export TC_DIST_TIMEOUT_S=10 RANK=0 WORLD_SIZE=1 LOCAL_RANK=0 MASTER_ADDR=127.0.0.1 MASTER_PORT=29571
timeout 300 python scripts/dist_check.py 2>&1 | grep "FAILED\|dist_check ok" | cut -c1-200
for r in 1300000 10350000; do
timeout 300 python scripts/pipe_bench.py --rows-per-gpu $r --variants classic 2>&1 | grep "^{" | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('dist', d['rows_per_gpu'], 'grid', d['grid'], 'ms/step %.3f us/iter %.1f Mdof %.0f'%(d['ms_per_step'],d['us_per_iter'],d['mdof_steps_per_s']), d['us_per_iter_by_phase_last'])"
done
unset RANK WORLD_SIZE LOCAL_RANK MASTER_ADDR MASTER_PORT
timeout 900 python bench.py --steps 8 --warmup 3 --no-cpu > gpurun_out/r02_spmv_bench_n1.json 2> gpurun_out/r02_spmv_bench_n1.err
tail -3 gpurun_out/r02_spmv_bench_n1.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r02_spmv_bench_n1.json") if l.startswith("{")][-1])
print("solver N=1", round(d["value"]/1e6,1), "ms/step", round(d["ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]/1e6,1), "rk45", d.get("rk45_explicit",{}).get("ms_per_attempt"), "parity", d.get("parity",{}).get("max_rel"), "phases", d["roofline"].get("phase_ms_block0"))
for k,v in d.get("extra",{}).items():
    print("  ", k, v.get("value") and round(v["value"]/1e6,2), "ms/step", v.get("ms_per_step"), "frac", v.get("roofline_frac") or (v.get("roofline") or {}).get("frac"), v.get("error"))
PY
