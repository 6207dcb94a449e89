"""Batched MOR ensemble: DMMA kernel vs a numpy restatement of solver.py:282-299 per scenario."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _numpy_rk4(A_body, A_films, B_T, film_surf, h0, tau, t_link, q_surf, X, t0, h, nsteps):
    S, r, B = X.shape
    F = A_films.shape[0]

    def f(t, X):
        films = h0 * (1.0 + 0.5 * np.sin(2 * np.pi * t / tau)) if F else np.zeros((0, B))
        out = np.empty_like(X)
        for s in range(S):
            z = A_body[s] @ X[s]
            flux = np.full(B, q_surf[s])
            for fi in range(F):
                z = z + films[fi][None, :] * (A_films[fi, s] @ X[s])
                if film_surf[fi] == s:
                    flux = flux + films[fi] * t_link[fi]
            out[s] = -z + B_T[s][:, None] * flux[None, :]
        return out
    t = t0
    for _ in range(nsteps):
        k1 = f(t, X); k2 = f(t + h / 2, X + h / 2 * k1); k3 = f(t + h / 2, X + h / 2 * k2); k4 = f(t + h, X + h * k3)
        X = X + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        t += h
    return X


def _random_case(S, r, F, B, seed):
    rng = np.random.default_rng(seed)
    def spd():
        Q = rng.standard_normal((r, r)) / np.sqrt(r)
        return Q @ Q.T + 0.1 * np.eye(r)
    A_body = np.array([spd() for _ in range(S)]) * 0.05
    A_films = np.array([[spd() * 1e-3 for _ in range(S)] for _ in range(F)]).reshape(F, S, r, r)
    B_T = rng.standard_normal((S, r)) * 1e-2
    film_surf = rng.integers(0, S, size=F).astype(np.int32)
    h0 = rng.uniform(5, 50, size=(F, B)); tau = rng.uniform(60, 600, size=(F, B))
    t_link = rng.uniform(-5, 5, size=F); q_surf = rng.uniform(0, 100, size=S)
    X0 = rng.standard_normal((S, r, B))
    return A_body, A_films, B_T, film_surf, h0, tau, t_link, q_surf, X0


@pytest.mark.parametrize("S,r,F,B", [(3, 40, 2, 96), (2, 200, 1, 130), (1, 17, 0, 64), (2, 64, 3, 64), (2, 33, 5, 50),
                                     (3, 200, 2, 192), (2, 198, 2, 70), (1, 136, 0, 66), (2, 150, 3, 64), (1, 128, 2, 64)])
@pytest.mark.parametrize("dmma", [3, 2, 1, 0])
def test_mor_batch_matches_numpy(S, r, F, B, dmma):
    from thermca_b200.mor_batch import MorEnsemble
    A_body, A_films, B_T, film_surf, h0, tau, t_link, q_surf, X0 = _random_case(S, r, F, B, 7)
    ens = MorEnsemble(A_body, A_films, B_T, film_surf, h0, tau, t_link, q_surf, X0)
    ens.time = 3.0
    ens.steps(3, 0.7, use_dmma=dmma)
    X = ens.state()
    Xref = _numpy_rk4(A_body, A_films, B_T, film_surf, h0, tau, t_link, q_surf, X0, 3.0, 0.7, 3)
    assert np.max(np.abs(X - Xref)) <= 1e-11 * np.max(np.abs(Xref))


@pytest.mark.parametrize("S,r,F,B,nsteps", [(3, 200, 2, 2000, 5), (1, 200, 2, 64, 8), (2, 96, 1, 1000, 3),
                                            (3, 200, 2, 4096, 2), (2, 200, 0, 70, 1), (1, 8, 2, 33, 4)])
def test_fused_kernel_equals_per_stage_kernels(S, r, F, B, nsteps):
    """The fused persistent kernel (mode 3) against the per-stage DMMA kernels (mode 2) at sizes a numpy loop does
    not finish quickly.  The shapes make units straddle CTAs in every way the schedule allows: more units than SMs
    (a unit starts on one CTA and ends on its neighbour), fewer units than SMs (one unit passes through up to 8
    CTAs), ragged last column tile, a single step."""
    from thermca_b200.mor_batch import MorEnsemble
    case = _random_case(S, r, F, B, 11)
    a = MorEnsemble(*case)
    b = MorEnsemble(*case)
    assert a.fused
    a.time = b.time = 1.5
    a.steps(nsteps, 0.3, use_dmma=3)
    b.steps(nsteps, 0.3, use_dmma=2)
    assert a.last_mode == 3 and b.last_mode == 2
    Xa, Xb = a.state(), b.state()
    assert np.isfinite(Xa).all()
    assert np.max(np.abs(Xa - Xb)) <= 1e-13 * np.max(np.abs(Xb))
    # a second call continues from the stored state and time
    a.steps(2, 0.3, use_dmma=3)
    b.steps(2, 0.3, use_dmma=2)
    assert np.max(np.abs(a.state() - b.state())) <= 1e-13 * np.max(np.abs(Xb))


def test_mor_batch_reproduces_reference_rhs_per_scenario(golden):
    """One scenario of the ensemble with the golden MOR part's operators == the oracle's MOR
    right-hand side (which is pinned to the reference) integrated with the same RK4."""
    from thermca_b200.mor_batch import MorEnsemble
    spec, ex = golden("plate_mor")
    p = spec["mor"][0]
    S, r = int(p["S"]), int(p["r"])
    F = int(p["F"])
    B = 64
    films = np.asarray(p["films"])
    h0 = np.repeat(films[:, None], B, axis=1) / 1.0
    tau = np.full((F, B), 1e30)                        # sin(2 pi t / tau) -> 0: constant films
    T = spec["init_temp"]
    t_link = T[p["link_elem_net_idxs"]] - float(p["ref_temp"])
    q_surf = spec["heat_src"][p["surf_net_idxs"]] / p["surf_areas"]
    x0 = np.asarray(p["x0"]).reshape(S, r)
    ens = MorEnsemble(p["A_body"], p["A_films"], p["B_T"], p["film_surf_idxs"], h0, tau, t_link, q_surf, x0)
    from rhs_oracle import OracleRHS
    orc = OracleRHS(spec)
    y = ex["y0"].copy()
    h = 0.5
    for _ in range(4):                                   # RK4 with the oracle RHS
        k1 = orc(0.0, y); k2 = orc(0.0, y + h / 2 * k1); k3 = orc(0.0, y + h / 2 * k2); k4 = orc(0.0, y + h * k3)
        y = y + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
    ens.steps(4, h)
    X = ens.state()
    assert np.max(np.abs(X[:, :, 0].ravel() - y)) <= 1e-10 * np.max(np.abs(y))
    assert np.array_equal(X[:, :, 0], X[:, :, B - 1])


def test_ensemble_on_reference_operators_matches_reference_run(golden):
    """Every scenario of an ensemble built from the reference's MORIO operators (golden plate_mor) with the
    golden's constant films reproduces the surface temperatures of the reference's own Network.sim run."""
    from thermca_b200.mor_batch import MorEnsemble
    spec, ex = golden("plate_mor")
    part = spec["mor"][0]
    ref = ex["result"]
    F, B = int(part["F"]), 64
    films = np.asarray(part["films"], dtype=np.float64)
    h0 = np.repeat(films[:, None], B, axis=1)
    tau = np.full((F, B), 1e300)                               # sin(2 pi t / tau) = 0: constant films
    link_T = np.asarray(spec["init_temp"])[np.asarray(part["link_elem_net_idxs"])]
    heat = np.asarray(spec["heat_src"])[np.asarray(part["surf_net_idxs"])]
    ens = MorEnsemble.from_part(part, link_T, heat, h0, tau)
    t_end = float(ref["times"][-1])
    ens.steps(int(round(t_end / 0.25)), 0.25)
    Ts = ens.surface_temps()                                   # (S, B)
    want = ref["net_temps"][-1][np.asarray(part["surf_net_idxs"])]
    assert np.abs(Ts - Ts[:, :1]).max() == 0.0                 # identical scenarios -> identical columns
    # the reference run is adaptive RK45 at rtol 1e-3; RK4 with h = 0.25 s is far more accurate
    assert np.allclose(Ts[:, 0] - 20.0, want - 20.0, rtol=2e-3, atol=1e-3)
    # and against a tight CPU integration of the same lowered network
    from rhs_oracle import integrate
    tight = integrate(spec, [0.0, t_end], 1e-10, 1e-12, "RK45")
    want_t = tight["net_temps"][-1][np.asarray(part["surf_net_idxs"])]
    assert np.allclose(Ts[:, 0], want_t, rtol=1e-8, atol=1e-8)
