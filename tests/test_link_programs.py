"""Link-closure programs: every library film / conductance closure the tracer lowers,
checked against the values the unmodified reference closure returns (tests/golden/link_programs.npz,
oracle/make_golden_links.py), including the printed known answers of
docs/examples/heat_transfer_on_skin.ipynb:61-209."""
import ctypes as C
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "link_programs.npz")
KAT = {"radiation.therm_radn[kat]": 4.8055738452682135, "free_conv.vert_surf[kat]": 3.889609859462753,
       "forced_conv.cyl_cross_flow[kat]": 3.8022233532323804, "combd_film.mix_conv[kat]": 4.574476631734567}


def _load():
    z = np.load(GOLDEN)
    tables = [(z[f"table{i}_x"], z[f"table{i}_y"]) for i in range(int(z["n_tables"]))]
    progs = []
    for k, name in enumerate(z["names"]):
        progs.append((str(name), z[f"{k}_code"], z[f"{k}_consts"], int(z[f"{k}_result"]), z[f"{k}_expect"]))
    return progs, tables, z["t0"], z["t1"]


def test_host_interpreter_reproduces_reference_closures_bit_exactly():
    from thermca_b200.trace import run_program
    progs, tables, t0, t1 = _load()
    assert len(progs) == 28
    for name, code, consts, result, expect in progs:
        got = run_program([tuple(r) for r in code], list(consts), result, [t0, t1], [], tables)
        assert np.array_equal(got, expect, equal_nan=True), name
        if name in KAT:                                      # sample 0 is the notebook's (36, 22); its digits come
            assert abs(got[0] - KAT[name]) <= 4e-16 * KAT[name]     # from a scalar call (libm pow vs numpy's array pow)


def test_oracle_evaluator_is_independent_and_reproduces_reference_closures():
    """The oracle evaluates link programs with its own table-driven evaluator (oracle/program_oracle.py); it
    shares only the storage format (opcode numbering) with the product and is pinned against the values of the
    unmodified reference closures, bit for bit."""
    import program_oracle
    from thermca_b200.trace import OPS
    assert list(program_oracle.OPCODES) == list(OPS)
    src = open(program_oracle.__file__).read() + open(os.path.join(os.path.dirname(program_oracle.__file__),
                                                                   "rhs_oracle.py")).read()
    assert "import thermca_b200" not in src and "from thermca_b200" not in src
    progs, tables, t0, t1 = _load()
    for name, code, consts, result, expect in progs:
        got = program_oracle.eval_program(code, consts, result, [t0, t1], [], tables)
        assert np.array_equal(got, expect, equal_nan=True), name


@pytest.mark.gpu
def test_device_interpreter_matches_reference_closures():
    import torch
    from thermca_b200 import _cabi
    lib = _cabi.load_library()
    progs, tables, t0, t1 = _load()
    dev = torch.device("cuda")
    # table arena: x then y per table
    tab_dir, arena, off = [], [], 0
    for x, y in tables:
        tab_dir += [off, len(x)]
        arena += [x, y]
        off += 2 * len(x)
    up = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)
    d_dir, d_arena = up(np.array(tab_dir), np.int32), up(np.concatenate(arena), np.float64)
    d_t0, d_t1 = up(t0, np.float64), up(t1, np.float64)
    d_in = torch.zeros(1, dtype=torch.float64, device=dev)
    p = lambda a: C.c_void_p(a.data_ptr())
    worst = 0.0
    for name, code, consts, result, expect in progs:
        d_code = up(code.ravel(), np.int32)
        d_consts = up(consts if len(consts) else np.zeros(1), np.float64)
        out = torch.empty(len(t0), dtype=torch.float64, device=dev)
        _cabi.check(lib.tc_vm_eval(C.c_void_p(torch.cuda.current_stream().cuda_stream), p(d_code), len(code), result,
                                   p(d_consts), p(d_dir), p(d_arena), p(d_in), p(d_t0), p(d_t1), len(t0), p(out)))
        torch.cuda.synchronize()
        got = out.cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(expect)), name
        ok = ~np.isnan(expect)
        rel = np.abs(got[ok] - expect[ok]) / np.maximum(np.abs(expect[ok]), 1e-300)
        worst = max(worst, rel.max() if ok.any() else 0.0)
        # libm vs CUDA pow/exp/log differ by an ulp or two; everything else is the same IEEE operation
        assert rel.max() <= 2e-14, (name, rel.max())
        if name in KAT:
            assert abs(got[0] - KAT[name]) <= 1e-14 * KAT[name]
    print("worst relative deviation", worst)
