"""GPU parity tests: the CUDA path (through the C ABI, via qoc_b200's host mirror of the qoc API) against
the CPU oracle (oracle/qoc_oracle.py: torch complex128 restatement of the reference + torch.autograd) and
the golden vectors produced by the unmodified reference forward (tests/golden/).

Tolerance (BASELINE.json north_star): cost and gradient within 1e-10 relative in complex128 mode."""
import glob
import os

import numpy as np
import pytest

from tests.problems import (GOLDEN, Problem, golden_schroedinger_costs, load_golden, numpy_hamiltonian)

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300)


def oracle_mod():
    from oracle import qoc_oracle
    return qoc_oracle


def product():
    import qoc_b200.standard as std
    from qoc_b200.core.plan import SchroedingerPlan
    from qoc_b200.models import MagnusPolicy
    return std, SchroedingerPlan, {2: MagnusPolicy.M2, 4: MagnusPolicy.M4, 6: MagnusPolicy.M6}


# --- batched expm (qoc/standard/functions/expm.py:210-252) -----------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 3, 4, 7, 8, 13, 16, 24, 32, 48, 60, 64])
@pytest.mark.parametrize("norm", [0.3, 4.0, 30.0])
def test_expm_batched_vs_oracle(n, norm):
    import torch
    orc = oracle_mod()
    from qoc_b200.standard.functions import expm
    rng = np.random.default_rng(n * 100 + int(norm))
    batch = 5
    a = rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))
    a[0] = -1j * (a[0] + a[0].conj().T)                       # one anti-hermitian member
    for b in range(batch):
        a[b] *= norm * (0.5 + 0.5 * b / batch) / np.abs(a[b]).sum(axis=0).max()
    got = expm(a)
    for b in range(batch):
        want = orc.expm_pade(torch.as_tensor(a[b])).numpy()
        assert rel(got[b], want) < 1e-12, (n, norm, b)


def test_expm_golden_vectors():
    from qoc_b200.standard.functions import expm
    d = load_golden("unit_vectors.npz")
    for i in range(int(d["expm_count"])):
        got = expm(d["expm_in_%d" % i])
        assert rel(got, d["expm_out_%d" % i]) < 1e-12, i


@pytest.mark.parametrize("n", [2, 5, 8, 16, 31, 32, 60, 64])
@pytest.mark.parametrize("norm", [1.5, 12.0])
def test_expm_vjp_vs_autograd(n, norm):
    """reverse pass of the Pade graph in autograd's cotangent convention vs torch.autograd over the oracle's
    expm_pade: torch's .grad of Re<conj(ubar), U> is conj(abar)."""
    import torch
    orc = oracle_mod()
    from qoc_b200.standard.functions import expm_vjp
    rng = np.random.default_rng(n + 7)
    batch = 3
    a = rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))
    for b in range(batch):
        a[b] *= norm / np.abs(a[b]).sum(axis=0).max()
    ubar = rng.standard_normal((batch, n, n)) + 1j * rng.standard_normal((batch, n, n))
    out, abar = expm_vjp(a, ubar)
    for b in range(batch):
        at = torch.tensor(a[b], requires_grad=True)
        u = orc.expm_pade(at)
        # scalar L = Re sum(ubar * U): autograd-convention cotangent of U is ubar; torch grad = conj(abar)
        loss = torch.sum(torch.as_tensor(ubar[b]) * u).real
        loss.backward()
        assert rel(out[b], u.detach().numpy()) < 1e-12
        assert rel(abar[b], np.conj(at.grad.numpy())) < 1e-11, (n, norm, b)


# --- golden Schroedinger cases: reference forward + oracle gradient -------------------------------------------
@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "schroedinger_case_*.npz"))))
def test_golden_schroedinger_cases(path):
    d = np.load(path)
    std, Plan, pol = product()
    orc = oracle_mod()
    cc = bool(d["complex_controls"])
    K, M, N, order, ces, T = int(d["K"]), int(d["M"]), int(d["N"]), int(d["order"]), int(d["cost_eval_step"]), float(d["T"])
    plan = Plan(numpy_hamiltonian(d["h0"], d["drives"], cc), d["initial_states"], golden_schroedinger_costs(d, std), T, N,
                control_eval_count=M, control_count=K, complex_controls=cc, magnus_policy=pol[order], cost_eval_step=ces)
    err, grads, finals = plan.cost_and_grad(d["controls"])
    err_fwd, finals_fwd = plan.cost(d["controls"])
    # reference forward (golden)
    assert abs(err - float(d["error"])) <= RTOL * abs(float(d["error"]))
    assert abs(err_fwd - float(d["error"])) <= RTOL * abs(float(d["error"]))
    assert rel(finals, d["final_states"]) < RTOL
    assert rel(finals_fwd, d["final_states"]) < RTOL
    # oracle gradient (torch autograd over the reference's operation sequence)
    o_err, o_grad, _ = orc.schroedinger_cost_and_grad(d["controls"], orc.make_hamiltonian(d["h0"], d["drives"], cc),
                                                      d["initial_states"], golden_schroedinger_costs(d, orc), T, N,
                                                      order=order, cost_eval_step=ces)
    assert abs(err - o_err) <= RTOL * abs(o_err)
    assert rel(grads, o_grad) < RTOL
    if "fd_grad" in d.files:                                 # finite differences of the reference forward
        assert rel(grads, d["fd_grad"]) < 1e-6
    plan.close()


def test_golden_examples():
    """examples/0_transmon_pi.py and examples/tutorial.py at iteration 0 (reference forward)."""
    std, Plan, pol = product()
    d = load_golden("cfg1_transmon_pi.npz")
    from qoc_b200.standard import (TargetStateInfidelity, get_annihilation_operator, get_creation_operator, SIGMA_Z)
    a, ad = get_annihilation_operator(2), get_creation_operator(2)
    h = lambda c, t: SIGMA_Z / 2 + c[0] * a + np.conjugate(c[0]) * ad
    init = np.array([[[1], [0]]], dtype=np.complex128)
    target = np.array([[[0], [1]]], dtype=np.complex128)
    plan = Plan(h, init, [TargetStateInfidelity(target)], 10.0, 11, control_eval_count=11, control_count=1,
                complex_controls=True, magnus_policy=pol[2])
    err, grads, finals = plan.cost_and_grad(d["controls"])
    assert abs(err - float(d["error"])) < 1e-12
    assert rel(finals, d["final_states"]) < RTOL
    assert rel(grads, d["fd_grad"]) < 1e-6
    plan.close()
    t = load_golden("tutorial_iter0.npz")
    hm = numpy_hamiltonian(t["h0"].astype(complex), t["drives"].astype(complex), True)
    plan = Plan(hm, t["initial_states"], [TargetStateInfidelity(t["target_states"])], 15.0, 100,
                control_eval_count=100, control_count=2, complex_controls=True, magnus_policy=pol[2])
    err, grads, finals = plan.cost_and_grad(t["controls"])
    assert abs(err - float(t["error"])) < 1e-12
    assert abs(err - float(t["notebook_error"])) < 5e-10     # examples/tutorial.ipynb:313 prints 9 digits
    assert rel(finals, t["final_states"]) < RTOL
    plan.close()


def test_iswap_known_answer():
    """tests/test_core.py:450-469: H = (XX+YY)/2, T = pi/2, N = 1000, all Magnus orders."""
    std, Plan, pol = product()
    g = load_golden("iswap_schroedinger.npz")
    sx = np.array([[0, 1], [1, 0]], dtype=complex)
    sy = np.array([[0, -1j], [1j, 0]])
    H = (np.kron(sx, sx) + np.kron(sy, sy)) / 2
    init = np.eye(4, dtype=complex)[:, :, None]
    want = np.array([[1, 0, 0, 0], [0, 0, -1j, 0], [0, -1j, 0, 0], [0, 0, 0, 1]]).T[:, :, None]
    for order in (2, 4, 6):
        plan = Plan(lambda c, t: H, init, [], np.pi / 2, 1000, magnus_policy=pol[order])
        _, finals = plan.cost(None)
        assert np.allclose(finals, want, atol=1e-12)
        assert rel(finals, g["final_states_m%d" % order]) < RTOL
        plan.close()


# --- random problems vs oracle: Magnus orders, cost kinds, chunking, tape modes, stiff (s > 0) ------------------
CASES = [
    # n, slices, K, S, order, complex, F, stiff, ces, step_target, neglect
    (2, 10, 1, 1, 2, True, 0, 1.0, 1, False, False),
    (3, 17, 2, 2, 4, False, 1, 1.0, 1, False, False),
    (4, 23, 2, 2, 6, True, 2, 1.0, 2, True, False),
    (8, 40, 2, 4, 4, True, 3, 1.0, 3, True, True),
    (12, 31, 3, 3, 6, False, 2, 8.0, 1, False, True),
    (16, 50, 2, 4, 4, True, 4, 8.0, 5, True, False),
    (20, 29, 2, 2, 2, False, 2, 40.0, 1, False, False),
    (32, 30, 2, 8, 4, True, 3, 1.0, 1, False, False),
    (60, 24, 2, 4, 4, True, 6, 1.0, 1, False, False),
    (64, 20, 4, 4, 4, False, 0, 1.0, 1, False, False),
    (64, 12, 2, 3, 6, True, 2, 8.0, 2, True, False),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "n%d_N%d_K%d_S%d_M%d_%s_F%d_x%g_ces%d_%d%d" % (
    c[0], c[1], c[2], c[3], c[4], "c" if c[5] else "r", c[6], c[7], c[8], c[9], c[10]))
@pytest.mark.parametrize("mode", ["tape", "recompute", "onechunk"])
def test_random_problem_vs_oracle(case, mode):
    n, slices, K, S, order, cc, F, stiff, ces, step_target, neglect = case
    std, Plan, pol = product()
    orc = oracle_mod()
    p = Problem(n, slices, K, S, order, complex_controls=cc, F=F, seed=n + slices, stiff=stiff, cost_eval_step=ces,
                step_target=step_target, neglect_phase=neglect)
    kw = dict(store_tape=(mode != "recompute"))
    if mode == "onechunk":
        kw["chunks_per_member"] = 1
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M,
                control_count=K, complex_controls=cc, magnus_policy=pol[order], cost_eval_step=ces, **kw)
    err, grads, finals = plan.cost_and_grad(p.controls)
    err2, grads2, _ = plan.cost_and_grad(p.controls)          # plans are reusable and deterministic
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, cc),
                                                          p.initial_states, p.costs(orc), p.T, p.N, order=order,
                                                          cost_eval_step=ces)
    assert err == err2 and np.array_equal(grads, grads2)
    assert abs(err - o_err) <= RTOL * max(abs(o_err), 1e-3), (err, o_err)
    assert rel(finals, o_fin) < RTOL
    assert grads.shape == p.controls.shape and grads.dtype == p.controls.dtype
    assert rel(grads, o_grad) < RTOL, rel(grads, o_grad)
    plan.close()


def test_controls_off_grid_and_intermediate_states():
    """control_eval_count != system_eval_count (interpolation incl. extrapolation at the last nodes,
    qoc/core/mathmethods.py:54-59) and the stored psi_j (save_intermediate_states payload)."""
    std, Plan, pol = product()
    orc = oracle_mod()
    p = Problem(6, 37, 2, 2, 4, complex_controls=True, F=2, seed=5, M=11)
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M,
                control_count=2, complex_controls=True, magnus_policy=pol[4])
    err, grads, finals = plan.cost_and_grad(p.controls)
    o_err, o_grad, _ = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, True),
                                                      p.initial_states, p.costs(orc), p.T, p.N, order=4)
    assert abs(err - o_err) <= RTOL * abs(o_err)
    assert rel(grads, o_grad) < RTOL
    import torch
    _, _, trail = orc.evaluate_schroedinger(torch.as_tensor(p.controls), orc.make_hamiltonian(p.h0, p.drives, True),
                                            p.initial_states, p.costs(orc), p.T, p.N, order=4, keep_states=True)
    states = plan.intermediate_states()
    assert states.shape == (p.N, 2, 6, 1)
    assert rel(states, np.stack([t.numpy() for t in trail])) < RTOL
    plan.close()


def test_ensemble_mean_over_members():
    """build-side extension (cfg5): members differ in the drift; cost = mean over members of the single-member
    cost.  Oracle = python loop over members."""
    std, Plan, pol = product()
    orc = oracle_mod()
    p = Problem(8, 21, 2, 3, 2, complex_controls=False, F=0, seed=11)
    rng = np.random.default_rng(3)
    E = 5
    z = np.diag(np.arange(8) - 3.5).astype(complex)
    drifts = np.stack([p.h0 + d * z for d in rng.normal(0, 0.1, E)])
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M,
                control_count=2, magnus_policy=pol[2], ensemble_drifts=drifts)
    err, grads, finals = plan.cost_and_grad(p.controls)
    o_err, o_grad = 0.0, 0.0
    for e in range(E):
        v, g, _ = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(drifts[e], p.drives, False),
                                                 p.initial_states, p.costs(orc), p.T, p.N, order=2)
        o_err += v / E
        o_grad = o_grad + g / E
    assert abs(err - o_err) <= RTOL * abs(o_err)
    assert rel(grads, o_grad) < RTOL
    assert finals.shape == (E, 3, 8, 1)
    plan.close()


def test_full_size_properties():
    """BASELINE shape (n=64, 2000 slices, M4): size-independent properties instead of the (slow) oracle -
    unitarity of the propagators, norm preservation of every stored state, gradient vs a directional central
    difference of the GPU forward, determinism across chunkings."""
    std, Plan, pol = product()
    p = Problem(64, 2000, 4, 4, 4, complex_controls=False, F=0, seed=0)
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M,
                control_count=4, magnus_policy=pol[4])
    err, grads, finals = plan.cost_and_grad(p.controls)
    states = plan.intermediate_states()[..., 0]
    assert np.abs(np.linalg.norm(states, axis=-1) - 1).max() < 1e-11
    U = plan.propagators()[::97]
    assert np.abs(U @ U.conj().transpose(0, 2, 1) - np.eye(64)).max() < 1e-12
    rng = np.random.default_rng(1)
    v = rng.standard_normal(p.controls.shape)
    eps = 1e-5
    ep, _ = plan.cost(p.controls + eps * v)
    em, _ = plan.cost(p.controls - eps * v)
    fd = (ep - em) / (2 * eps)
    assert abs(fd - np.sum(grads * v)) < 1e-6 * max(1.0, abs(fd))
    plan.close()
    plan1 = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M,
                 control_count=4, magnus_policy=pol[4], chunks_per_member=7, store_tape=False)
    err1, grads1, _ = plan1.cost_and_grad(p.controls)
    assert abs(err - err1) < 1e-11
    assert rel(grads1, grads) < 1e-9
    plan1.close()


def test_errors_are_loud():
    std, Plan, pol = product()
    p = Problem(4, 5, 1, 1, 2)
    with pytest.raises(NotImplementedError):
        Plan(lambda c, t: p.h0 + c[0] ** 2 * p.drives[0], p.initial_states, [], p.T, p.N, control_eval_count=p.M,
             control_count=1, magnus_policy=pol[2])
    with pytest.raises(NotImplementedError):
        Plan(lambda c, t: p.h0 * (1 + t) + np.sin(c[0] * t) * p.drives[0], p.initial_states, [], p.T, p.N, control_eval_count=p.M,
             control_count=1, magnus_policy=pol[2])          # time-dependent AND non-linear in the controls


EDGE = [
    # n, slices, K, S, order, complex, F, stiff, ces, step_target, label
    (64, 1, 2, 1, 4, False, 0, 1.0, 1, False, "single_slice_S1_lowrank"),
    (64, 9, 7, 3, 4, False, 1, 1.0, 1, False, "KR7_commutators_off"),
    (64, 9, 4, 2, 4, True, 0, 1.0, 1, False, "KR8_complex_commutators_off"),
    (64, 9, 2, 5, 4, True, 2, 1.0, 1, False, "S5_dense_reverse"),
    (61, 7, 1, 4, 2, True, 4, 1.0, 3, True, "n61_M2_lowrank"),
    (64, 8, 2, 4, 6, False, 0, 1.0, 1, False, "M6_lowrank"),
    (64, 6, 2, 4, 4, False, 0, 6.0, 1, False, "mixed_s_lowrank_and_dense"),
    (5, 1, 1, 1, 4, True, 1, 1.0, 1, True, "n5_single_slice"),
]


@pytest.mark.parametrize("case", EDGE, ids=lambda c: c[-1])
def test_edge_cases_vs_oracle(case):
    n, slices, K, S, order, cc, F, stiff, ces, step_target, _ = case
    std, Plan, pol = product()
    orc = oracle_mod()
    p = Problem(n, slices, K, S, order, complex_controls=cc, F=F, seed=17 + n, stiff=stiff, cost_eval_step=ces,
                step_target=step_target)
    if case[-1] == "mixed_s_lowrank_and_dense":          # some slices need squarings (s > 0), some do not
        p.controls[: p.M // 2] *= 0.02
        p.h0 = p.h0 * 0.8
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M,
                control_count=K, complex_controls=cc, magnus_policy=pol[order], cost_eval_step=ces)
    err, grads, finals = plan.cost_and_grad(p.controls)
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, cc),
                                                          p.initial_states, p.costs(orc), p.T, p.N, order=order,
                                                          cost_eval_step=ces)
    assert abs(err - o_err) <= RTOL * max(abs(o_err), 1e-3)
    assert rel(finals, o_fin) < RTOL
    assert rel(grads, o_grad) < RTOL, rel(grads, o_grad)
    plan.close()


def test_lowrank_and_commutator_paths_match_dense(monkeypatch):
    """the rank-S reverse pass and the product-free Magnus M4 against the dense / product forms of the same kernels"""
    std, Plan, pol = product()
    p = Problem(64, 40, 4, 4, 4, complex_controls=False, F=2, seed=2)
    kw = dict(control_eval_count=p.M, control_count=4, magnus_policy=pol[4])
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, **kw)
    e1, g1, f1 = plan.cost_and_grad(p.controls)
    plan.close()
    monkeypatch.setenv("QOCB_NO_LOWRANK", "1")
    monkeypatch.setenv("QOCB_NO_COMM", "1")
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, **kw)
    e2, g2, f2 = plan.cost_and_grad(p.controls)
    plan.close()
    assert abs(e1 - e2) < 1e-13 and rel(f1, f2) < 1e-12 and rel(g1, g2) < 1e-11


VARIANTS = [{"QOCB_NO_NOPIV": "1"}, {"QOCB_NO_THREE_LEVEL": "1"}, {"QOCB_THREE_LEVEL_STEP": "0"}, {"QOCB_NO_TMA": "1"},
            {"QOCB_NO_PREMAGNUS": "1"}, {"QOCB_LOWRANK": "1"}, {"QOCB_NO_LOWRANK": "1", "QOCB_NO_NOPIV": "1"}]


@pytest.mark.parametrize("F,ces", [(0, 1), (2, 3)], ids=["final_cost", "step_costs"])
def test_fast_paths_match_general_paths(monkeypatch, F, ces):
    """Hermitian operators select the half products, the pivot-free LU, the one-product reverse stage and (with enough slices)
    the three-level boundary scheme; every A/B switch must reproduce the default result, and the default the oracle."""
    std, Plan, pol = product()
    orc = oracle_mod()
    p = Problem(64, 310, 3, 4, 4, complex_controls=False, F=F, seed=23, cost_eval_step=ces)
    kw = dict(control_eval_count=p.M, control_count=3, magnus_policy=pol[4], cost_eval_step=ces)
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, **kw)
    e0, g0, f0 = plan.cost_and_grad(p.controls)
    plan.close()
    o_err, o_grad, o_fin = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(p.h0, p.drives, False),
                                                          p.initial_states, p.costs(orc), p.T, p.N, order=4, cost_eval_step=ces)
    assert abs(e0 - o_err) <= RTOL * max(abs(o_err), 1e-3) and rel(g0, o_grad) < RTOL and rel(f0, o_fin) < RTOL
    for env in VARIANTS:
        with monkeypatch.context() as m:
            for k, v in env.items():
                m.setenv(k, v)
            plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, **kw)
            e1, g1, f1 = plan.cost_and_grad(p.controls)
            plan.close()
        assert abs(e1 - e0) < 1e-12 * max(abs(e0), 1e-3), env
        assert rel(f1, f0) < 1e-11 and rel(g1, g0) < 1e-10, (env, rel(g1, g0))


def test_ensemble_with_lowrank_path():
    std, Plan, pol = product()
    orc = oracle_mod()
    p = Problem(64, 6, 2, 2, 4, complex_controls=False, F=0, seed=4)
    rng = np.random.default_rng(8)
    z = np.diag(np.linspace(-1, 1, 64)).astype(complex) * 0.05
    drifts = np.stack([p.h0 + d * z for d in rng.normal(0, 1.0, 3)])
    plan = Plan(p.hamiltonian_numpy(), p.initial_states, p.costs(std), p.T, p.N, control_eval_count=p.M, control_count=2,
                magnus_policy=pol[4], ensemble_drifts=drifts)
    err, grads, _ = plan.cost_and_grad(p.controls)
    o_err, o_grad = 0.0, 0.0
    for e in range(3):
        v, g, _ = orc.schroedinger_cost_and_grad(p.controls, orc.make_hamiltonian(drifts[e], p.drives, False),
                                                 p.initial_states, p.costs(orc), p.T, p.N, order=4)
        o_err += v / 3
        o_grad = o_grad + g / 3
    assert abs(err - o_err) <= RTOL * abs(o_err) and rel(grads, o_grad) < RTOL
    plan.close()
