import ctypes as C, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from thermca_b200 import _cabi
lib = _cabi.load_library()
z = np.load(os.path.join(ROOT, "tests", "golden", "link_programs.npz"))
tables = [(z[f"table{i}_x"], z[f"table{i}_y"]) for i in range(int(z["n_tables"]))]
tab_dir, arena, off = [], [], 0
for x, y in tables:
    tab_dir += [off, len(x)]; arena += [x, y]; off += 2 * len(x)
dev = torch.device("cuda")
up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
d_dir, d_arena = up(np.array(tab_dir), np.int32), up(np.concatenate(arena), np.float64)
p = lambda a: C.c_void_p(a.data_ptr())
d_in = torch.zeros(1, dtype=torch.float64, device=dev)
for k, name in enumerate(z["names"]):
    code, consts, result = z[f"{k}_code"], z[f"{k}_consts"], int(z[f"{k}_result"])
    d_code, d_consts = up(code.ravel(), np.int32), up(consts if len(consts) else np.zeros(1), np.float64)
    for m in (1, 32, 4096):
        a0 = torch.full((m,), 36.0, dtype=torch.float64, device=dev); a1 = torch.full((m,), 22.0, dtype=torch.float64, device=dev)
        out = torch.empty(m, dtype=torch.float64, device=dev)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        def run():
            _cabi.check(lib.tc_vm_eval(st, p(d_code), len(code), result, p(d_consts), p(d_dir), p(d_arena), p(d_in), p(a0), p(a1), m, p(out)))
        for _ in range(3): run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): run()
        e1.record(); torch.cuda.synchronize()
        print(f"{name:45s} {len(code):3d} instr  m={m:5d}: {1e3 * e0.elapsed_time(e1) / 20:8.1f} us per launch", flush=True)
    if k >= 5: break
