"""Pin the CPU oracle against every in-tree golden vector / known-answer test of the reference for this path
 and against independent physics.  CPU only.

Each test cites the reference file:line whose assertion it restates.
"""
import os

import numpy as np
import pytest
from scipy.integrate import solve_ivp

import nqcdynamics_jl_b200 as nq
import oracle
from helpers import A, model_config

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _cfg(model, method=A.METHOD_FSSH, masses=None, **kw):
    natoms = model.natoms or 1
    masses = masses if masses is not None else np.ones(model.ndofs * natoms)
    base = dict(method=method, masses=masses, ntraj=1, dt=1.0)
    base.update(kw)
    return A.make_config(**model_config(model, **base))


# ---- ring polymers ------------------------------------------------------------------------------
def test_normal_mode_transform_doc_vectors():
    """docs/src/api/RingPolymerArrays/ringpolymerarrays.md:93-130 (B = 4)."""
    U = oracle.normal_mode_matrix(4)
    assert np.allclose(U.T @ np.ones(4), [2.0, 0.0, 0.0, 0.0], atol=1e-12)
    assert np.allclose(U @ np.ones(4), [1.70711, -0.707107, 0.292893, 0.707107], atol=1e-5)


@pytest.mark.parametrize("B", [1, 2, 3, 4, 5, 10, 16, 32])
def test_normal_mode_matrix_diagonalises_springs(B):
    """test/Core/ring_polymers.jl:13,22-27: U'U = I, |det U| = 1, U'SU = diag(lambda), eigvals(S) = NM springs."""
    omega_n = 0.7
    U = oracle.normal_mode_matrix(B)
    if B == 1:
        S = np.zeros((1, 1))
    elif B == 2:
        S = np.array([[2.0, -2.0], [-2.0, 2.0]])
    else:
        S = 2 * np.eye(B) - np.roll(np.eye(B), 1, axis=0) - np.roll(np.eye(B), -1, axis=0)
    S = S * omega_n ** 2 / 2                                           # ring_polymer.jl:46-57
    lam = (2 * omega_n * np.sin(np.arange(B) * np.pi / B)) ** 2 / 2     # ring_polymer.jl:59-60
    assert np.allclose(U.T @ U, np.eye(B), atol=1e-13)
    assert abs(abs(np.linalg.det(U)) - 1) < 1e-12
    assert np.allclose(U.T @ S @ U, np.diag(lam), atol=1e-12)
    assert np.allclose(np.sort(np.linalg.eigvalsh(S)), np.sort(lam), atol=1e-12)


def test_cayley_propagator():
    """test/Core/ring_polymers.jl:30-32 (full == half*half) and ring_polymer.jl:73-79 (definition)."""
    B, omega_n, dt = 10, 10 * 0.003, 0.1
    half = oracle.cayley(B, omega_n, dt, True)
    full = oracle.cayley(B, omega_n, dt, False)
    wk = 2 * omega_n * np.sin(np.arange(B) * np.pi / B)
    for k in range(B):
        assert np.allclose(full[k], half[k] @ half[k], atol=1e-14)
        Am = np.array([[0.0, 1.0], [-wk[k] ** 2, 0.0]])
        ref = np.linalg.inv(np.eye(2) - dt * Am / 2) @ (np.eye(2) + dt * Am / 2)
        assert np.allclose(full[k], ref, atol=1e-14)


def test_bcb_second_order_on_harmonic_ring_polymer():
    """test/Dynamics/algorithms/bcb.jl:18-43: BCB vs the analytic harmonic ring polymer, order 2 (atol 0.4)."""
    B, m, w = 10, 1837.4715941070515, 1.0e-2
    kT = 300 * 3.166811563e-6
    model = nq.Harmonic(m=m, ω=w)
    rng = np.random.default_rng(2)
    r0 = rng.normal(0.0, 0.01, B)
    v0 = rng.normal(0.0, np.sqrt(B * kT / m), B)
    tf = 200.0
    U = oracle.normal_mode_matrix(B)
    wk = np.sqrt((2 * B * kT * np.sin(np.arange(B) * np.pi / B)) ** 2 + w ** 2)
    rn, vn = U.T @ r0, U.T @ v0
    r_exact = U @ (rn * np.cos(wk * tf) + vn / wk * np.sin(wk * tf))
    errs = []
    dts = [8.0, 4.0, 2.0, 1.0]
    for dt in dts:
        cfg, keep = _cfg(model, A.METHOD_CLASSICAL, masses=[m], dt=dt, nbeads=B, temperature=kT)
        h = oracle.OracleEngine(cfg, keep)
        h.set_state(r0, v0)
        h.run(int(round(tf / dt)))
        errs.append(np.max(np.abs(h.get_state()["r"].ravel() - r_exact)))
    orders = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert np.all(np.abs(orders - 2.0) < 0.4), orders


# ---- calculator cache / models -------------------------------------------------------------------
@pytest.mark.parametrize("model,r", [
    (nq.TullyModelOne(), [-0.3]), (nq.TullyModelTwo(), [0.7]), (nq.TullyModelThree(), [-1.1]),
    (nq.DoubleWell(), [0.4]), (nq.ThreeStateMorse(), [3.4]),
    (nq.SpinBoson(nq.DebyeSpectralDensity(0.25, 0.5), 5, 0.1, 1.0), [0.3, -0.2, 0.5, 0.1, -0.7]),
    (nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(8, -0.0192, 0.0192)), [11.0]),
    (nq.AndersonHolstein(nq.ErpenbeckThoss(Γ=6.4e-3), nq.TrapezoidalRule(8, -0.0192, 0.0192)), [5.9]),     # test/Dynamics/iesh.jl:17-25
    (nq.AndersonHolstein(nq.ErpenbeckThoss(Γ=0.2 / 27.2114), nq.ShenviGaussLegendre(8, -0.9, 0.9)), [3.1]),  # iesh.md:85-105
])
def test_calculator_cache_identities(model, r):
    """test/Core/calculators.jl:99-108 (w = eigvals(V), |Z| = |eigvecs|, adiab = Z' dV Z) and
    test/Dynamics/fssh.jl:32-33 (NAC antisymmetric); derivative checked by finite differences."""
    cfg, keep = _cfg(model)
    ev = oracle.evaluate_model(cfg, r)
    w_np, Z_np = np.linalg.eigh(ev["V"])
    assert np.allclose(ev["w"], w_np, atol=1e-13)
    assert np.allclose(np.abs(ev["Z"]), np.abs(Z_np), atol=1e-10)
    for I in range(cfg.ndofs):
        assert np.allclose(ev["adiab"][I], ev["Z"].T @ ev["dV"][I] @ ev["Z"], atol=1e-13)
        assert np.allclose(ev["nac"][I], -ev["nac"][I].T, atol=1e-13)
        h = 1e-6
        rp, rm = np.array(r, dtype=float), np.array(r, dtype=float)
        rp[I] += h; rm[I] -= h
        fd = (oracle.evaluate_model(cfg, rp)["V"] - oracle.evaluate_model(cfg, rm)["V"]) / (2 * h)
        assert np.allclose(ev["dV"][I], fd, atol=1e-7)
        # d_ij = <phi_i | d/dR phi_j>: finite difference of the eigenvectors (gauge = continuity)
        Zp, Zm = oracle.evaluate_model(cfg, rp)["Z"], oracle.evaluate_model(cfg, rm)["Z"]
        for Zx in (Zp, Zm):
            Zx *= np.sign(np.sum(Zx * ev["Z"], axis=0))
        d_fd = ev["Z"].T @ (Zp - Zm) / (2 * h)
        assert np.allclose(ev["nac"][I], d_fd, atol=1e-5)


def test_double_well_populations_at_origin():
    """test/Dynamics/fssh.jl:28,47-51 ; ehrenfest.jl:26,40-43 ; DynamicsUtils.jl:40-62."""
    model = nq.DoubleWell()
    obs = (1 << A.OBS_DIABATIC_POP) | (1 << A.OBS_SIGMA)
    for method in (A.METHOD_FSSH, A.METHOD_EHRENFEST):
        cfg, keep = _cfg(model, method, masses=[2.0], observables=obs, nsave=1)
        h = oracle.OracleEngine(cfg, keep)
        sigma = np.zeros((1, 2, 2)); sigma[0, 0, 0] = 1.0
        h.set_state([0.0], [0.3], sigma, None, [1] if method == A.METHOD_FSSH else None)   # PureState(1, Adiabatic())
        assert np.allclose(h.get_state()["sigma"][0], [[1, 0], [0, 0]])
        assert np.allclose(h.observable_sum(A.OBS_DIABATIC_POP)[0], [0.5, 0.5], atol=1e-14)
        # diabatic PureState(1) -> adiabatic density with |rho| = 0.5 everywhere, exact round trip
        h2 = oracle.OracleEngine(*_cfg(model, method, masses=[2.0], observables=obs, nsave=1))
        rho = np.zeros((1, 2, 2)); rho[0, 0, 0] = 1.0
        h2.set_state_diabatic([0.0], [0.3], rho, None, [1] if method == A.METHOD_FSSH else None)
        s = h2.get_state()["sigma"][0]
        assert np.allclose(np.abs(s), 0.5, atol=1e-14)
        Z = oracle.evaluate_model(cfg, [0.0])["Z"]
        assert np.allclose(Z @ s @ Z.T, [[1, 0], [0, 0]], atol=1e-14)


def test_select_new_state_truth_table():
    """test/Dynamics/fssh.jl:53-63."""
    assert oracle.select_new_state([0.0, 1.0], 1, 0.37) == 2
    assert oracle.select_new_state([0.0, 1.0], 2, 0.37) == 2
    assert oracle.select_new_state([1.0, 0.0], 2, 0.37) == 1


@pytest.mark.parametrize("nbeads", [1, 5])
def test_rescale_velocity_accept_reject_and_energy(nbeads):
    """test/Dynamics/fssh.jl:65-74,87-121 (RP: :154-208): v = 0 -> frustrated, v = 1e5 -> accepted,
    dKE = -dE across an accepted hop."""
    model = nq.DoubleWell()
    m = 2.0
    cfg, keep = _cfg(model, masses=[m], nbeads=nbeads, temperature=1.0)
    r = np.full(nbeads, 0.37)
    ok, v, eig = oracle.unit_rescale(cfg, r, np.zeros(nbeads), 2, 1)
    assert not ok and np.all(v == 0.0)
    ok, v, eig = oracle.unit_rescale(cfg, r, np.full(nbeads, 1e5), 2, 1)
    assert ok
    v0 = np.full(nbeads, 2.0)
    ok, v1, eig = oracle.unit_rescale(cfg, r, v0, 2, 1)
    assert ok
    dE = eig[1] - eig[0]
    # centroid kinetic energy carries the hop (hopping velocity = centroid, rpsh.jl:30-37)
    dKE = 0.5 * m * (np.mean(v1) ** 2 - np.mean(v0) ** 2)
    assert abs(dKE + dE) < 1e-3 * abs(dE)
    # :vinversion reflects the velocity on a frustrated hop (surface_hopping.jl:155-164)
    cfg2, keep2 = _cfg(model, masses=[m], nbeads=nbeads, temperature=1.0, rescaling=A.RESCALE_VINVERSION)
    ok, v2, _ = oracle.unit_rescale(cfg2, r, np.full(nbeads, 1e-3), 2, 1)
    assert not ok and np.allclose(v2, -1e-3)


def test_fssh_energy_conservation_tully_two():
    """test/Dynamics/fssh.jl:123-133: TullyModelTwo, v = 100/2000, r = -10, dt = 0.1, E conserved to 1e-2."""
    cfg, keep = _cfg(nq.TullyModelTwo(), masses=[2000.0], dt=0.1, save_every=500, nsave=11,
                     observables=1 << A.OBS_TOTAL_ENERGY, seed=5)
    h = oracle.OracleEngine(cfg, keep)
    sigma = np.zeros((1, 2, 2)); sigma[0, 0, 0] = 1.0
    h.set_state([-10.0], [100.0 / 2000], sigma, None, [1])
    h.run(5000)
    E = h.observable_sum(A.OBS_TOTAL_ENERGY)[:, 0]
    assert abs(E[-1] - E[0]) < 1e-2 * abs(E[0])


# ---- electronic propagation ----------------------------------------------------------------------
def test_tsit5_tableau_order_conditions():
    """The Tsit5 tableau (OrdinaryDiffEq, external) satisfies the Runge-Kutta order conditions through order 5."""
    c = np.array([0, 0.161, 0.327, 0.9, 0.9800255409045097, 1.0])
    Am = np.zeros((6, 6))
    Am[1, 0] = 0.161
    Am[2, :2] = [-0.008480655492356989, 0.335480655492357]
    Am[3, :3] = [2.8971530571054935, -6.359448489975075, 4.3622954328695815]
    Am[4, :4] = [5.325864828439257, -11.748883564062828, 7.4955393428898365, -0.09249506636175525]
    Am[5, :5] = [5.86145544294642, -12.92096931784711, 8.159367898576159, -0.071584973281401, -0.028269050394068383]
    b = np.array([0.09646076681806523, 0.01, 0.4798896504144996, 1.379008574103742, -3.290069515436081, 2.324710524099774])
    assert np.allclose(Am.sum(1), c, atol=1e-14)
    conds = [(b.sum(), 1), (b @ c, 1 / 2), (b @ c ** 2, 1 / 3), (b @ Am @ c, 1 / 6), (b @ c ** 3, 1 / 4),
             (b @ (c * (Am @ c)), 1 / 8), (b @ Am @ c ** 2, 1 / 12), (b @ Am @ Am @ c, 1 / 24), (b @ c ** 4, 1 / 5),
             (b @ Am @ c ** 3, 1 / 20), (b @ Am @ Am @ Am @ c, 1 / 120)]
    for got, want in conds:
        assert abs(got - want) < 5e-15


def test_density_propagation_preserves_trace_and_matches_scipy():
    """test/Dynamics/electronic_dynamics.jl:35-47 (trace preserved) + the oracle's 5 Tsit5 sub-steps against a
    tight-tolerance integration of the same interpolated generator (electronic_dynamics.jl:55-116)."""
    rng = np.random.default_rng(4)
    n = 3
    E0, E1 = np.sort(rng.normal(size=n)) * 0.1, np.sort(rng.normal(size=n)) * 0.1
    g0, g1 = rng.normal(size=(n, n)) * 0.05, rng.normal(size=(n, n)) * 0.05
    g0, g1 = g0 - g0.T, g1 - g1.T
    psi = rng.normal(size=n) + 1j * rng.normal(size=n); psi /= np.linalg.norm(psi)
    sigma0 = np.outer(psi, psi.conj())
    t, dt = 3.0, 1.0
    out = oracle.propagate_density(E0, g0, t, E1, g1, t + dt, t, dt, sigma0)
    assert abs(np.trace(out) - 1.0) < 1e-12
    assert np.allclose(out, out.conj().T, atol=1e-13)

    def rhs(tau, y):
        s = (y[:n * n] + 1j * y[n * n:]).reshape(n, n)
        loc = (tau - t) / dt
        Amat = np.diag(E0 + (E1 - E0) * loc).astype(complex) - 1j * (g0 + (g1 - g0) * loc)
        ds = -1j * (Amat @ s - s @ Amat)
        return np.concatenate([ds.real.ravel(), ds.imag.ravel()])
    sol = solve_ivp(rhs, (t, t + dt), np.concatenate([sigma0.real.ravel(), sigma0.imag.ravel()]), rtol=1e-12, atol=1e-14)
    ref = (sol.y[:n * n, -1] + 1j * sol.y[n * n:, -1]).reshape(n, n)
    assert np.max(np.abs(out - ref)) < 1e-8      # Tsit5 at dt/5 truncation error, not rounding


def test_first_step_quirk_q1():
    """Q1 (bab_electronics.jl:35-39, electronic_dynamics.jl:118-127): on the first nuclear step the electronic
    generator ramps linearly from ZERO at t = 0 to the true (E, v.d) at t0 + dt."""
    cfg, keep = _cfg(nq.TullyModelOne(), A.METHOD_EHRENFEST, masses=[2000.0], diagnostics=1)
    h = oracle.OracleEngine(cfg, keep)
    rho = np.zeros((1, 2, 2)); rho[0, 0, 0] = 0.5; rho[0, 1, 1] = 0.5; rho[0, 0, 1] = rho[0, 1, 0] = 0.5
    h.set_state_diabatic([-0.4], [0.01], rho)
    s0 = h.get_state()["sigma"][0]
    h.run(1)
    d = h.diagnostics()
    vd = d["nac"][0, 0] * h.get_state()["v"].ravel()[0]
    expect = oracle.propagate_density(np.zeros(2), np.zeros((2, 2)), 0.0, d["eig"][0], vd, 1.0, 0.0, 1.0, s0)
    assert np.max(np.abs(h.get_state()["sigma"][0] - expect)) < 1e-14


def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors)."""
    assert list(oracle.philox_raw([0, 0, 0, 0], [0, 0])) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(oracle.philox_raw([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(oracle.philox_raw([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    u = np.array([oracle.philox_uniform(7, g, s) for g in range(50) for s in range(50)])
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.02


# ---- independent physics ---------------------------------------------------------------------------
def test_ehrenfest_matches_exact_diabatic_propagation():
    """Pins the NAC sign/formula, the density-matrix EOM and the diabatic<->adiabatic transformation at the
    physics level: the oracle's adiabatic-representation Ehrenfest trajectory converges to a tight-tolerance
    integration of the same mean-field equations written in the DIABATIC representation.  (Convergence is
    first order in dt, not second: quirk Q1 puts an O(dE * dt) phase error into the first step.)"""
    a, b, c, d, m = 0.01, 1.6, 0.005, 1.0, 2000.0

    def Vd(q):
        v11 = a * (1 - np.exp(-b * q)) if q > 0 else -a * (1 - np.exp(b * q))
        v12 = c * np.exp(-d * q * q)
        return np.array([[v11, v12], [v12, -v11]])

    def dVd(q):
        d11 = a * b * np.exp(-b * abs(q)); d12 = -2 * c * d * q * np.exp(-d * q * q)
        return np.array([[d11, d12], [d12, -d11]])

    def rhs(t, y):
        q, p = y[0], y[1]; cc = y[2:4] + 1j * y[4:6]
        F = -np.real(cc.conj() @ dVd(q) @ cc)
        dc = -1j * Vd(q) @ cc
        return [p / m, F, dc[0].real, dc[1].real, dc[0].imag, dc[1].imag]
    q0, k0, tf = -4.0, 10.0, 1600.0
    sol = solve_ivp(rhs, (0, tf), [q0, k0, 1, 0, 0, 0], rtol=1e-11, atol=1e-13, method="DOP853")
    exact = np.abs(sol.y[2:4, -1] + 1j * sol.y[4:6, -1]) ** 2
    errs = []
    for dt in (1.0, 0.5):
        n = int(tf / dt)
        cfg, keep = _cfg(nq.TullyModelOne(), A.METHOD_EHRENFEST, masses=[m], dt=dt, save_every=n, nsave=2,
                         observables=1 << A.OBS_DIABATIC_POP)
        h = oracle.OracleEngine(cfg, keep)
        rho = np.zeros((1, 2, 2)); rho[0, 0, 0] = 1
        h.set_state_diabatic([q0], [k0 / m], rho)
        h.run(n)
        errs.append(np.max(np.abs(h.observable_sum(A.OBS_DIABATIC_POP)[-1] - exact)))
        assert abs(h.get_state()["r"].ravel()[0] - sol.y[0, -1]) < 5e-3
    assert errs[0] < 5e-4 and errs[1] < errs[0] / 1.7


def test_spin_boson_ehrenfest_vs_gao_saller_curve():
    """test/Dynamics/ehrenfest.jl:110-144: Ohmic(2.5, 0.09), N = 100, beta = 5, eps = 0, Delta = 1, dt = 0.1,
    500 trajectories; <sigma_z>(t) = P11 - P12 within 0.2 of the digitised Gao/Saller JCTC 2020 Fig. 2b curve
    (tests/golden/gao_saller_jctc_2020_fig2b.csv, copied verbatim from the reference's test data)."""
    N, beta, T = 100, 5.0, 500
    model = nq.SpinBoson(nq.OhmicSpectralDensity(2.5, 0.09), N, 0.0, 1.0)
    w = model.bath_a
    rng = np.random.default_rng(2020)
    sr = np.sqrt(1 / (2 * w * np.tanh(beta * w / 2))); sv = np.sqrt(w / (2 * np.tanh(beta * w / 2)))
    r = rng.standard_normal((T, N)) * sr; v = rng.standard_normal((T, N)) * sv
    kw = model_config(model, method=A.METHOD_EHRENFEST, masses=np.ones(N), ntraj=T, dt=0.1, save_every=1, nsave=201,
                      observables=1 << A.OBS_POPCORR_DIABATIC)
    cfg, keep = A.make_config(**kw)
    h = oracle.OracleEngine(cfg, keep)
    rho = np.zeros((T, 2, 2)); rho[:, 0, 0] = 1
    h.set_state_diabatic(r, v, rho)
    h.run(200)
    pc = h.observable_sum(A.OBS_POPCORR_DIABATIC) / T          # [i + 2 j] = P_i(0) P_j(t)
    result = pc[:, 0] - pc[:, 2]                                # p[1,1] - p[1,2]
    data = np.loadtxt(os.path.join(GOLDEN, "gao_saller_jctc_2020_fig2b.csv"), delimiter=",")
    t = 0.1 * np.arange(201)
    # linear interpolation with linear extrapolation (extrapolation_bc = Line())
    ref = np.interp(t, data[:, 0], data[:, 1])
    lo = t < data[0, 0]
    ref[lo] = data[0, 1] + (t[lo] - data[0, 0]) * (data[1, 1] - data[0, 1]) / (data[1, 0] - data[0, 0])
    hi = t > data[-1, 0]
    ref[hi] = data[-1, 1] + (t[hi] - data[-1, 0]) * (data[-1, 1] - data[-2, 1]) / (data[-1, 0] - data[-2, 0])
    assert np.max(np.abs(result - ref)) < 0.2


def test_nrpmd_energy_conservation_second_order():
    """test/Dynamics/nrpmd.jl:40-83: RingPolymerMInt is symplectic and of order 2 -- the error of the NRPMD
    Hamiltonian (spring + kinetic + mapping potential, nrpmd.jl:124-139) quarters when dt halves; the total
    mapping population sum_j [(q_j^2 + p_j^2)/2 - gamma] stays 1."""
    B, T, g = 4, 3, 0.5
    rng = np.random.default_rng(0)
    r0 = 0.3 * rng.standard_normal((T, B, 1)); v0 = 0.5 * rng.standard_normal((T, B, 1))
    th = rng.random((T, B, 2)) * 2 * np.pi
    R = np.array([np.sqrt(2 + 2 * g), np.sqrt(2 * g)])          # nrpmd.jl:47-65, PureState(1)
    q0, p0 = np.cos(th) * R, np.sin(th) * R
    errs = []
    for dt in (0.02, 0.01, 0.005):
        n = int(round(2.0 / dt))
        kw = model_config(nq.DoubleWell(), method=A.METHOD_NRPMD, masses=[1.0], ntraj=T, dt=dt, nbeads=B, temperature=0.7,
                          save_every=n, nsave=2, observables=(1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_DIABATIC_POP),
                          per_trajectory=1, nrpmd_gamma=g)
        cfg, keep = A.make_config(**kw)
        h = oracle.OracleEngine(cfg, keep)
        h.set_state(r0, v0); h.set_mapping(q0, p0)
        h.run(n)
        E = h.observable_per_trajectory(A.OBS_TOTAL_ENERGY)[:, :, 0]
        errs.append(np.max(np.abs(E[:, 1] - E[:, 0])))
        pop = h.observable_per_trajectory(A.OBS_DIABATIC_POP)
        assert np.allclose(pop[:, 0], [1.0, 0.0], atol=1e-12) and np.allclose(pop.sum(axis=2), 1.0, atol=1e-10)
    orders = np.log2(np.array(errs[:-1]) / np.array(errs[1:]))
    assert np.all(np.abs(orders - 2.0) < 0.1), orders


# ---- IESH pieces -----------------------------------------------------------------------------------
def test_set_unoccupied_states():
    """test/Dynamics/iesh.jl:66-73."""
    assert list(oracle.unoccupied(31, np.arange(1, 16))) == list(range(16, 32))
    assert list(oracle.unoccupied(31, np.arange(6, 21))) == list(range(1, 6)) + list(range(21, 32))


def test_fast_determinant():
    """test/Core/FastDeterminant.jl:8-13: det! == det."""
    rng = np.random.default_rng(8)
    for n in (1, 2, 5, 15):
        M = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n))
        assert abs(oracle.complex_det(M) - np.linalg.det(M)) < 1e-10 * max(1.0, abs(np.linalg.det(M)))


def test_edc_decoherence_monotone_and_normalised():
    """test/Dynamics/test_decoherence_corrections.jl:15-29: unoccupied amplitudes shrink, the norm stays 1."""
    rng = np.random.default_rng(9)
    n = 6
    psi = rng.normal(size=n) + 1j * rng.normal(size=n); psi /= np.linalg.norm(psi)
    E = np.sort(rng.normal(size=n))
    out = oracle.edc(psi, 3, 1.0, E, 0.02, 0.1)
    others = [i for i in range(n) if i != 2]
    assert np.all(np.abs(out[others]) < np.abs(psi[others]))
    assert abs(np.linalg.norm(out) - 1.0) < 1e-13
    assert abs(out[2]) > abs(psi[2])


def test_hermitian_propagator_is_unitary_and_matches_expm():
    """wavefunction_dynamics.jl:25-58: U = V exp(-i lambda dt) V' from the complex Hermitian eigensolver."""
    from scipy.linalg import expm
    rng = np.random.default_rng(10)
    n = 9
    H = rng.normal(size=(n, n)) + 1j * rng.normal(size=(n, n)); H = (H + H.conj().T) / 2
    w, V = oracle.herm_eigh(H)
    U = V @ np.diag(np.exp(-1j * w * 0.3)) @ V.conj().T
    assert np.allclose(U, expm(-1j * H * 0.3), atol=1e-12)
    assert np.allclose(U.conj().T @ U, np.eye(n), atol=1e-12)


def test_symmetric_eigensolvers_agree_with_numpy():
    rng = np.random.default_rng(11)
    for n in (2, 3, 7, 31, 101):
        M = rng.normal(size=(n, n)); M = M + M.T
        for algo in (1, 2):
            if algo == 1 and n > 40:
                continue
            w, Z = oracle.sym_eigh(M, algo)
            assert np.allclose(w, np.linalg.eigvalsh(M), atol=1e-11)
            assert np.allclose(Z.T @ M @ Z, np.diag(w), atol=1e-10)
            assert np.allclose(Z.T @ Z, np.eye(n), atol=1e-12)


def _iesh_oracle(M, T, dt, nsteps, **extra):
    import nqcdynamics_jl_b200 as nq
    from helpers import A, model_config
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(M, -0.0192, 0.0192))
    kw = model_config(model, method=A.METHOD_IESH, masses=[2000.0], ntraj=T, dt=dt, rng=A.RNG_INJECTED, save_every=1,
                      nsave=nsteps + 1, per_trajectory=1, diagnostics=1,
                      observables=(1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_KINETIC) | (1 << A.OBS_POTENTIAL) |
                                  (1 << A.OBS_DISCRETE_STATE) | (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_DIABATIC_POP))
    kw.update(extra)
    cfg, keep = A.make_config(**kw)
    return model, oracle.OracleEngine(cfg, keep)


def test_iesh_hop_conserves_energy():
    """test/Dynamics/iesh.jl:122-152: across an accepted hop dKE = -dE and the Hamiltonian is unchanged."""
    from helpers import A
    M, T, nsteps = 30, 6, 60
    rng = np.random.default_rng(31)
    model, h = _iesh_oracle(M, T, 0.02, nsteps)
    n, ne = model.nstates, model.nelectrons
    psi = np.zeros((T, ne, n), dtype=complex)
    for t in range(T):
        q, _ = np.linalg.qr(np.eye(n, ne) + 0.3 * np.linalg.qr(rng.standard_normal((n, ne)) + 1j * rng.standard_normal((n, ne)))[0])
        psi[t] = q.T
    state = np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1))
    h.set_state(5.0 + 10.0 * rng.random(T), -np.abs(rng.standard_normal(T)) * 8e-3, psi.real, psi.imag, state)
    h.set_draws(rng.random((nsteps, T)) * 2e-6)         # tiny draws: a hop is attempted on almost every step
    h.run(nsteps)
    c = h.counters()
    assert c["hops"] > 5, c
    E = h.observable_per_trajectory(A.OBS_TOTAL_ENERGY)[:, :, 0]
    KE = h.observable_per_trajectory(A.OBS_KINETIC)[:, :, 0]
    PE = h.observable_per_trajectory(A.OBS_POTENTIAL)[:, :, 0]
    occ = np.sort(np.rint(h.observable_per_trajectory(A.OBS_DISCRETE_STATE)).astype(int), axis=2)
    hopped = np.any(occ[:, 1:] != occ[:, :-1], axis=2)
    assert hopped.sum() == c["hops"]
    # dt = 0.02: the integration error (incl. the force that is not refreshed after a hop) is ~1e-8, hop gaps are ~1e-3
    assert np.max(np.abs(E - E[:, :1])) < 1e-6
    dKE, dPE = np.diff(KE, axis=1)[hopped], np.diff(PE, axis=1)[hopped]
    assert np.all(np.abs(dPE) > 1e-5) and np.allclose(dKE, -dPE, rtol=1e-3)
    # populations: ne electrons in total in either representation; adiabatic occupations are 0/1 (iesh.jl:371-375)
    adi = h.observable_per_trajectory(A.OBS_ADIABATIC_POP); dia = h.observable_per_trajectory(A.OBS_DIABATIC_POP)
    assert np.allclose(adi.sum(axis=2), ne) and set(np.unique(adi)) <= {0.0, 1.0}
    assert np.allclose(dia.sum(axis=2), ne, atol=1e-9)


def test_iesh_ground_state_overlap_and_pruning():
    """test/Dynamics/iesh.jl:87-109: for the ground-state DynamicsVariables the overlap is the identity (det S = 1), so
    with zero velocity every hopping probability vanishes and a pruned and an unpruned run agree exactly."""
    from helpers import A
    M, T, nsteps = 30, 3, 5
    outs = []
    for est in (1, 0):
        model, h = _iesh_oracle(M, T, 1.0, nsteps, estimate_probability=est)
        n, ne = model.nstates, model.nelectrons
        psi = np.zeros((T, ne, n)); psi[:, np.arange(ne), np.arange(ne)] = 1.0
        state = np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1))
        h.set_state(np.array([3.0, 12.0, 21.0]), np.array([1e-4, -2e-4, 3e-4]), psi, None, state)
        h.set_draws(np.full((nsteps, T), 0.5))
        h.run(nsteps)
        outs.append((h.get_state(), h.counters(), h.hop_search_count()))
    (s1, c1, n1), (s0, c0, n0) = outs
    assert n1 == 0 and n0 == nsteps * T and c1["hops"] == c0["hops"] == 0
    assert np.array_equal(s1["sigma"], s0["sigma"]) and np.array_equal(s1["r"], s0["r"])
    assert np.allclose(np.abs(np.linalg.det(s1["sigma"][:, :ne, :])), 1.0, atol=1e-6)     # still close to adiabatic


def test_ehrenfest_na_conserves_energy_and_electrons():
    """test/Dynamics/ehrenfest_na.jl:24-52: Simulation{EhrenfestNA} on AndersonHolstein(MiaoSubotnik, TrapezoidalRule(30)),
    r = 21, v = 0, ground-state orbitals, dt = 10, tspan (0, 2000): var(total energy) < 1e-6; the electron count
    sum_m adiabatic_population[m] = ne is conserved (unitary propagation)."""
    from helpers import A
    M, T, nsteps = 30, 2, 200
    model, h = _iesh_oracle(M, T, 10.0, nsteps, method=A.METHOD_EHRENFEST_NA,
                            observables=(1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_KINETIC) | (1 << A.OBS_POTENTIAL) |
                                        (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_POSITION))
    n, ne = model.nstates, model.nelectrons
    psi = np.zeros((T, ne, n)); psi[:, np.arange(ne), np.arange(ne)] = 1.0
    h.set_state(np.array([21.0, 18.0]), np.array([0.0, -1e-4]), psi, None, None)
    h.run(nsteps)
    E = h.observable_per_trajectory(A.OBS_TOTAL_ENERGY)[:, :, 0]
    assert np.all(np.var(E, axis=1) < 1e-6)
    assert np.max(np.abs(E - E[:, :1])) < 2e-5 * np.max(np.abs(E))
    adi = h.observable_per_trajectory(A.OBS_ADIABATIC_POP)
    assert np.allclose(adi.sum(axis=2), ne, atol=1e-9) and np.all(adi > -1e-12) and np.all(adi < 1 + 1e-9)
    x = h.observable_per_trajectory(A.OBS_POSITION)[:, :, 0]
    assert np.max(np.abs(x[0] - 21.0)) > 1e-3            # the trajectory does move on the mean-field surface


def test_rp_ehrenfest_na_and_rpiesh_conserve_energy():
    """test/Dynamics/rp_ehrenfest_na.jl:10-41: RingPolymerSimulation{EhrenfestNA}(Atoms(2000), AndersonHolstein(MiaoSubotnik,
    TrapezoidalRule(30); fermi_level = 0.001), 4 beads), v = 0, r = 21 + N(0, 1), dt = 10, tspan (0, 2000), BCBWavefunction:
    var(total energy) < 1e-6 -- the reference's own assertion, on the oracle's step_rpiesh (bead forces, spring term, psi
    propagated with the previous centroid generator).  RPIESH without hops (huge draws) on the same system likewise."""
    import nqcdynamics_jl_b200 as nq
    from helpers import A, model_config, oracle_factory
    B, T, nsteps = 4, 2, 200
    rng = np.random.default_rng(5)
    model = nq.AndersonHolstein(nq.MiaoSubotnik(Γ=6.4e-3), nq.TrapezoidalRule(30, -0.0192, 0.0192), fermi_level=0.001)
    n, ne = model.nstates, model.nelectrons
    psi = np.zeros((T, ne, n)); psi[:, np.arange(ne), np.arange(ne)] = 1.0
    r = 21.0 + rng.standard_normal((T, B))
    obs = (1 << A.OBS_TOTAL_ENERGY) | (1 << A.OBS_ADIABATIC_POP) | (1 << A.OBS_POSITION)
    for method in (A.METHOD_EHRENFEST_NA, A.METHOD_IESH):
        kw = model_config(model, method=method, masses=[2000.0], ntraj=T, dt=10.0, nbeads=B, temperature=9.5e-4, rng=A.RNG_INJECTED,
                          save_every=1, nsave=nsteps + 1, observables=obs, per_trajectory=1)
        cfg, keep = A.make_config(**kw)
        h = oracle_factory()(cfg, keep)
        state = np.tile(np.arange(1, ne + 1, dtype=np.int32), (T, 1)) if method == A.METHOD_IESH else None
        h.set_state(r, np.zeros((T, B)), psi, None, state)
        if method == A.METHOD_IESH:
            h.set_draws(np.full((nsteps, T), 0.999999))
        h.run(nsteps)
        E = h.observable_per_trajectory(A.OBS_TOTAL_ENERGY)[:, :, 0]
        assert np.all(np.var(E, axis=1) < 1e-6), (method, np.var(E, axis=1))
        adi = h.observable_per_trajectory(A.OBS_ADIABATIC_POP)
        assert np.allclose(adi.sum(axis=2), ne, atol=1e-9)
        x = h.observable_per_trajectory(A.OBS_POSITION)[:, :, 0]
        assert np.max(np.abs(x[:, -1] - x[:, 0])) > 1e-3      # the centroid moves
        h.close()
