"""GPU parity tests of the Lindblad path against the golden vectors of the
unmodified reference forward, the reference's known answers, the torch oracle and the NumPy adjoint model.

Tolerances.  The reference integrates adaptively at atol = 1e-12; the numpy reference and the torch oracle already
differ by 2e-10 on final densities (a rounding-level change of an error norm near the accept threshold moves the
whole step sequence), so the forward bar is 1e-9.  The product's gradient is the discrete adjoint on the realised
grid: it matches the oracle with `freeze_steps=True` to 1e-7; against the oracle's full autograd gradient (which
also differentiates the step-size controller) only the reference's own reproducibility band can be asserted - the
test measures that band by perturbing the controls by 1e-13."""
import glob
import os

import numpy as np
import pytest

from tests.problems import GOLDEN, load_golden, numpy_hamiltonian

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def _costs(mod, d):
    N, ces = int(d["N"]), int(d["cost_eval_step"])
    return [mod.TargetDensityInfidelity(d["target_densities"], cost_multiplier=0.8),
            mod.ForbidDensities(d["forbidden_densities"], N, cost_eval_step=ces, cost_multiplier=0.4),
            mod.TargetDensityInfidelityTime(N, d["target_densities"], cost_eval_step=ces, cost_multiplier=0.3)]


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "lindblad_case_*.npz"))))
def test_golden_lindblad_cases(path):
    import qoc_b200.standard as std
    from oracle import qoc_oracle as orc
    from qoc_b200.core.plan import LindbladPlan
    d = np.load(path)
    cc = bool(d["complex_controls"])
    K, M, N, ces, T = int(d["K"]), int(d["M"]), int(d["N"]), int(d["cost_eval_step"]), float(d["T"])
    wh, wl = bool(d["with_hamiltonian"]), bool(d["with_lindblad"])
    ham = numpy_hamiltonian(d["h0"], d["drives"], cc) if wh else None
    ld = (lambda t: (d["gammas"], d["lindblad_ops"])) if wl else None
    plan = LindbladPlan(d["initial_densities"], _costs(std, d), T, N, hamiltonian=ham, lindblad_data=ld,
                        control_eval_count=M, control_count=K, complex_controls=cc, cost_eval_step=ces)
    err, grads, finals = plan.cost_and_grad(d["controls"])
    err_f, finals_f = plan.cost(d["controls"])
    stats = plan.stats()
    # forward: unmodified reference (golden)
    assert abs(err - float(d["error"])) < 1e-9 and abs(err_f - float(d["error"])) < 1e-9
    assert rel(finals, d["final_densities"]) < 1e-9 and rel(finals_f, d["final_densities"]) < 1e-9
    assert stats["accepted"] > 0 and stats["attempts"] >= stats["accepted"]
    # gradient: discrete adjoint on the realised grid = oracle with frozen step sizes
    o_ham = orc.make_hamiltonian(d["h0"], d["drives"], cc) if wh else None
    o_ld = orc.make_lindblad_data(d["gammas"], d["lindblad_ops"]) if wl else None
    _, g_frozen, _ = orc.lindblad_cost_and_grad(d["controls"], o_ham, o_ld, d["initial_densities"], _costs(orc, d), T, N,
                                                cost_eval_step=ces, freeze_steps=True)
    assert grads.shape == d["controls"].shape and grads.dtype == d["controls"].dtype
    assert rel(grads, g_frozen) < 1e-7, rel(grads, g_frozen)
    # finite differences of the reference forward (noisy: adaptive integrator, eps = 1e-5)
    assert rel(grads, d["fd_grad"]) < 1e-3
    # the oracle's full gradient: assert agreement within the band in which the oracle reproduces ITSELF
    _, g_full, _ = orc.lindblad_cost_and_grad(d["controls"], o_ham, o_ld, d["initial_densities"], _costs(orc, d), T, N,
                                              cost_eval_step=ces)
    band = 0.0
    for eps in (1e-13, -1e-13, 3e-13):
        _, g_p, _ = orc.lindblad_cost_and_grad(d["controls"] * (1 + eps), o_ham, o_ld, d["initial_densities"],
                                               _costs(orc, d), T, N, cost_eval_step=ces)
        band = max(band, rel(g_p, g_full))
    assert rel(grads, g_full) < max(10 * band, 1e-4), (rel(grads, g_full), band)
    plan.close()


def test_lindblad_known_answers():
    """amplitude damping (tests/test_core.py:124-148) and iSWAP on densities (:86-106) through the public API."""
    import qoc_b200 as qoc
    k = load_golden("lindblad_known.npz")
    gamma = float(k["ad_gamma"])
    sp = np.array([[0, 1], [0, 0]], dtype=complex)
    rho0 = k["ad_rho0"].astype(complex)[None]
    res = qoc.evolve_lindblad_discrete(1.0, rho0, 2, lindblad_data=lambda t: (np.array([gamma]), np.stack([sp])))
    assert rel(res.final_densities, k["ad_final"]) < 1e-9
    a0, b0 = rho0[0, 0, 0].real, rho0[0, 0, 1].real
    want = np.array([[1 - (1 - a0) * np.exp(-gamma), b0 * np.exp(-gamma / 2)],
                     [b0 * np.exp(-gamma / 2), (1 - a0) * np.exp(-gamma)]])
    assert np.allclose(res.final_densities[0], want, atol=1e-8)
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sy = np.array([[0, -1j], [1j, 0]])
    hm = 0.5 * (np.kron(sx, sx) + np.kron(sy, sy))
    init = np.eye(4, dtype=complex)[:, :, None]
    initd = np.matmul(init, np.conjugate(np.swapaxes(init, -1, -2)))
    res = qoc.evolve_lindblad_discrete(np.pi / 2, initd, 2, hamiltonian=lambda c, t: hm)
    assert rel(res.final_densities, k["iswap_final"]) < 1e-9
    u = np.array([[1, 0, 0, 0], [0, 0, -1j, 0], [0, -1j, 0, 0], [0, 0, 0, 1]])
    assert np.allclose(res.final_densities, u @ initd @ u.conj().T, atol=1e-8)


def test_adjoint_model_vs_cuda_same_algorithm():
    """the CUDA kernels against the NumPy model of the same algorithm on a larger random problem."""
    import qoc_b200.standard as std
    from oracle import lindblad_adjoint_model as lam
    from qoc_b200.core.plan import LindbladPlan
    rng = np.random.default_rng(5)
    n, D, K, M, N, L, T = 6, 3, 2, 7, 4, 2, 1.7
    h0 = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)); h0 = (h0 + h0.conj().T) / 2
    dr = np.triu(rng.standard_normal((K, n, n)) + 1j * rng.standard_normal((K, n, n)), 1) * 0.3
    u = (rng.standard_normal((M, K)) + 1j * rng.standard_normal((M, K))) * 0.4
    gam = rng.uniform(0.05, 0.3, L)
    ops = (rng.standard_normal((L, n, n)) + 1j * rng.standard_normal((L, n, n))) * 0.4
    def dens(count):
        out = []
        for _ in range(count):
            v = rng.standard_normal(n) + 1j * rng.standard_normal(n); v /= np.linalg.norm(v)
            out.append(np.outer(v, v.conj()))
        return np.array(out)
    rho0, targ = dens(D), dens(D)
    forb = np.array([dens(2) for _ in range(D)])
    costs = [std.TargetDensityInfidelity(targ), std.ForbidDensities(forb, N, cost_multiplier=0.5),
             std.ControlNorm(K, M, cost_multiplier=0.1)]
    plan = LindbladPlan(rho0, costs, T, N, hamiltonian=numpy_hamiltonian(h0, dr, True), lindblad_data=lambda t: (gam, ops),
                        control_eval_count=M, control_count=K, complex_controls=True)
    err, grads, finals = plan.cost_and_grad(u)
    dd = dr.conj().transpose(0, 2, 1)
    x = np.concatenate([u.real, u.imag], axis=1)
    model = lam.Model(h0, np.concatenate([dr + dd, 1j * (dr - dd)]), gam, ops, T, M)
    terms = [lam.DensityTerm(0, [targ[i][None] for i in range(D)], 1.0, False),
             lam.DensityTerm(1, [forb[i] for i in range(D)], 0.5 / ((N - 1) * D), True)]
    c2, g2, f2, st = lam.cost_and_grad(x, model, rho0, terms, T, N)
    cn, cg = costs[2].control_value_and_grad(u)
    assert abs(err - (c2 + cn)) < 1e-9
    assert rel(finals, f2) < 1e-9
    assert rel(grads, g2[:, :K] + 1j * g2[:, K:] + cg) < 1e-7
    assert rel(plan.intermediate_densities()[-1], f2) < 1e-9
    plan.close()


def test_grape_lindblad_runs_and_descends():
    """examples/1_transmon_pi_dechoerence.py (cfg2) shape through the public API: the optimiser loop runs on the
    GPU path and the error decreases."""
    import qoc_b200 as qoc
    from qoc_b200.standard import (Adam, TargetDensityInfidelity, get_annihilation_operator, get_creation_operator, SIGMA_Z)
    a, ad = get_annihilation_operator(2), get_creation_operator(2)
    h = lambda c, t: SIGMA_Z / 2 + c[0] * a + np.conjugate(c[0]) * ad
    rho0 = np.array([[[1, 0], [0, 0]]], dtype=complex)
    targ = np.array([[[0, 0], [0, 1]]], dtype=complex)
    res = qoc.grape_lindblad_discrete(1, 11, [TargetDensityInfidelity(targ)], 10.0, rho0, 2, complex_controls=True,
                                      hamiltonian=h, lindblad_data=lambda t: (np.array([1e-3]), np.stack([a])),
                                      iteration_count=15, max_control_norms=np.array([5.0]), optimizer=Adam(learning_rate=2e-2),
                                      log_iteration_step=0)
    first = qoc.evolve_lindblad_discrete(10.0, rho0, 2, controls=np.full((11, 1), 0.5 * (1 - 1j) / np.sqrt(2)),
                                         costs=[TargetDensityInfidelity(targ)], hamiltonian=h,
                                         lindblad_data=lambda t: (np.array([1e-3]), np.stack([a])))
    assert res.best_error < first.error
    assert res.best_final_densities.shape == (1, 2, 2) and res.best_controls.shape == (11, 1)


def test_grape_lindblad_trajectory_matches_oracle_driven_optimisation():
    """cfg2-like problem (examples/1_transmon_pi_dechoerence.py:22-60: n = 2, one density, amplitude damping, complex control):
    `grape_lindblad_discrete` with Adam for k iterations against the SAME host optimiser driven by the oracle's cost and
    frozen-grid gradient (the contract of include/qocb200.h).  The errors along the trajectory and the final controls must
    agree: gradient differences of 1e-7 move an Adam step by 1e-7 * learning rate, far below what the assertion allows."""
    import qoc_b200 as qoc
    import qoc_b200.standard as std
    from oracle import qoc_oracle as orc
    from qoc_b200.core.common import slap_controls, strip_controls
    n, T, M, N, k = 2, 10.0, 11, 3, 6
    a = np.diag(np.sqrt(np.arange(1, n)), 1).astype(complex)
    h0 = np.diag(np.arange(n) - 0.5).astype(complex)
    rho0 = np.zeros((1, n, n), dtype=complex); rho0[0, 0, 0] = 1
    targ = np.zeros((1, n, n), dtype=complex); targ[0, 1, 1] = 1
    gam = np.array([1e-3])
    u0 = 0.1 * (1 - 1j) / np.sqrt(2) * np.ones((M, 1), dtype=complex)
    errors = []

    class Rec(std.Adam):
        def run(self, function, iteration_count, initial_params, jacobian, args=()):
            def jac(params, *a_):
                g, term = jacobian(params, *a_)
                errors.append(a_[1].error)                   # reporter.error of this iteration
                return g, term
            return super().run(function, iteration_count, initial_params, jac, args=args)
    res = qoc.grape_lindblad_discrete(1, M, [std.TargetDensityInfidelity(targ)], T, rho0, N, complex_controls=True,
                                      hamiltonian=numpy_hamiltonian(h0, a[None], True), lindblad_data=lambda t: (gam, a[None]),
                                      initial_controls=u0.copy(), iteration_count=k, log_iteration_step=0, optimizer=Rec())
    # the same Adam on the host, fed by the oracle
    o_err = []
    oh, old = orc.make_hamiltonian(h0, a[None], True), orc.make_lindblad_data(gam, a[None])

    def o_fun(params, *a_):
        return 0.0, False

    def o_jac(params, *a_):
        u = slap_controls(True, params, (M, 1))
        e, g, _ = orc.lindblad_cost_and_grad(u, oh, old, rho0, [orc.TargetDensityInfidelity(targ)], T, N, freeze_steps=True)
        o_err.append(e)
        return strip_controls(True, g), False
    std.Adam().run(o_fun, k, strip_controls(True, u0.copy()), o_jac)
    assert len(errors) == k and len(o_err) == k
    assert np.abs(np.array(errors) - np.array(o_err)).max() < 1e-8, (errors, o_err)
    assert errors[-1] < errors[0] and res.best_error == min(errors)
