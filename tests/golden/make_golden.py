"""
make_golden.py - generate golden input/output vectors from the UNMODIFIED reference
(`/root/reference`, SchusterLab/qoc) for the GRAPE propagate-and-differentiate hot path.

The reference needs third-party packages that are absent in this image (HIPS autograd, h5py,
qutip, matplotlib, IPython).  Its *forward* arithmetic only uses `autograd.numpy` as a drop-in
for numpy, so this script installs stub modules (autograd.numpy := numpy; h5py/matplotlib/qutip/
IPython := empty shells) and then imports the real `qoc` package from /root/reference.  Every
number written here is therefore produced by the reference's own code:
  * qoc/core/schroedingerdiscrete.py:356-502   (_evaluate_schroedinger_discrete, one-slice step)
  * qoc/core/lindbladdiscrete.py:357-495       (_evaluate_lindblad_discrete, rhs)
  * qoc/core/mathmethods.py                    (interpolation, Magnus, Lindbladian, RKDP5)
  * qoc/standard/functions/expm.py:210-252     (expm_pade)
  * qoc/standard/costs/*.py                    (cost classes)
The reference's *backward* lives in autograd (absent), so gradients are pinned by central finite
differences of the reference forward (`fd_grad`, accuracy ~1e-8 relative) - an independent check of
the oracle's analytic gradients, not a 1e-10 pin.

Run (in the build container only; /root/reference does not exist on the GPU box):
    python tests/golden/make_golden.py
Outputs: tests/golden/*.npz (committed).
"""
import os
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"


def install_stubs():
    if not hasattr(np, "float_"):
        np.float_ = np.float64  # qoc/standard/utils/jsonutil.py:19 uses the removed alias
    if not hasattr(np, "string_"):
        np.string_ = np.bytes_
    ag = types.ModuleType("autograd")
    ag.numpy = np
    ext = types.ModuleType("autograd.extend")

    class Box(object):
        pass
    ext.Box = Box
    ext.defvjp = lambda *a, **k: None
    ext.primitive = lambda f: f
    ext.vspace = lambda x: None
    core = types.ModuleType("autograd.core")
    core.make_vjp = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("autograd is stubbed"))
    wrap = types.ModuleType("autograd.wrap_util")
    wrap.unary_to_nary = lambda f: f
    sys.modules.update({"autograd": ag, "autograd.numpy": np, "autograd.extend": ext,
                        "autograd.core": core, "autograd.wrap_util": wrap})
    for name in ("h5py", "matplotlib", "matplotlib.patches", "matplotlib.pyplot",
                 "matplotlib.gridspec", "IPython", "IPython.display"):
        sys.modules[name] = mock.MagicMock()
    qutip = types.ModuleType("qutip")
    qutip.__all__ = []
    sys.modules["qutip"] = qutip
    sys.path.insert(0, REF)


install_stubs()
import qoc  # noqa: E402  (the real reference package)
from qoc.core.schroedingerdiscrete import evolve_schroedinger_discrete  # noqa: E402
from qoc.core.lindbladdiscrete import evolve_lindblad_discrete  # noqa: E402
from qoc.core.mathmethods import (get_lindbladian, integrate_rkdp5, magnus_m2, magnus_m4,  # noqa: E402
                                  magnus_m6, interpolate_linear_set)
from qoc.core.common import clip_control_norms, gen_controls_flat  # noqa: E402
from qoc.models import MagnusPolicy  # noqa: E402
from qoc.standard import (TargetStateInfidelity, TargetStateInfidelityTime, ForbidStates,  # noqa: E402
                          TargetDensityInfidelity, TargetDensityInfidelityTime, ForbidDensities,
                          ControlNorm, ControlVariation, ControlBandwidthMax,
                          conjugate_transpose, get_annihilation_operator, get_creation_operator,
                          krons, matmuls, SIGMA_X, SIGMA_Y, SIGMA_Z, Adam, SGD)
from qoc.standard.functions.expm import expm_pade  # noqa: E402

POLICIES = {2: MagnusPolicy.M2, 4: MagnusPolicy.M4, 6: MagnusPolicy.M6}


def rand_herm(rng, n):
    x = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    return (x + x.conj().T) / 2


def haar_columns(rng, n, count):
    z = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    q, r = np.linalg.qr(z)
    q = q * (np.diag(r) / np.abs(np.diag(r)))
    return q[:, :count]


def make_hamiltonian(h0, drives, complex_controls):
    """H(u) = H0 + sum_k u_k D_k (real controls, D_k hermitian) or
    H0 + sum_k u_k C_k + conj(u_k) C_k^dagger (complex controls) - the shape of every
    hamiltonian in the reference's examples/tests (examples/0_transmon_pi.py:24-26)."""
    if complex_controls:
        def hamiltonian(controls, time):
            h = h0
            for k in range(len(drives)):
                h = h + controls[k] * drives[k] + np.conjugate(controls[k]) * drives[k].conj().T
            return h
    else:
        def hamiltonian(controls, time):
            h = h0
            for k in range(len(drives)):
                h = h + controls[k] * drives[k]
            return h
    return hamiltonian


def fd_grad(fn, controls, eps=1e-6):
    """central differences of the reference forward; complex controls -> d/dx + i d/dy
    (the optimiser convention after the wrapper's conjugate, schroedingerdiscrete.py:320-324)."""
    g = np.zeros_like(controls)
    it = np.nditer(controls, flags=["multi_index"])
    for _ in it:
        idx = it.multi_index
        parts = (1.0, 1j) if np.iscomplexobj(controls) else (1.0,)
        for part in parts:
            cp = controls.copy()
            cm = controls.copy()
            cp[idx] += eps * part
            cm[idx] -= eps * part
            d = (fn(cp) - fn(cm)) / (2 * eps)
            g[idx] += d * part
    return g


def schroedinger_case(seed, n, S, K, M, N, order, complex_controls, cost_eval_step, stiff, F=2,
                      with_time_cost=True, with_control_costs=True, neglect_phase=False, do_fd=True):
    rng = np.random.default_rng(seed)
    T = float(N - 1) * (0.37 if not stiff else 1.0)
    dt = T / (N - 1)
    scale = 8.0 if stiff else 1.0
    h0 = rand_herm(rng, n)
    h0 *= scale / (dt * np.abs(h0).sum(axis=0).max())
    drives = []
    for k in range(K):
        if complex_controls:
            c = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)), 1)
            c *= scale * 0.2 / (dt * np.abs(c + c.conj().T).sum(axis=0).max())
        else:
            c = rand_herm(rng, n)
            c *= scale * 0.2 / (dt * np.abs(c).sum(axis=0).max())
        drives.append(c)
    drives = np.stack(drives)
    if complex_controls:
        controls = (rng.normal(0, 0.5, (M, K)) + 1j * rng.normal(0, 0.5, (M, K))) / np.sqrt(2)
    else:
        controls = rng.normal(0, 0.5, (M, K))
    cols = haar_columns(rng, n, S)
    initial_states = np.stack([cols[:, s:s + 1] for s in range(S)])
    tcols = haar_columns(rng, n, n)
    target_states = np.stack([tcols[:, s:s + 1] for s in range(S)])
    forbidden = np.stack([np.stack([tcols[:, (S + s * F + f) % n][:, None] for f in range(F)])
                          for s in range(S)])
    max_norms = np.full(K, 2.5)
    costs = [TargetStateInfidelity(target_states, neglect_relative_pahse=neglect_phase, cost_multiplier=0.9),
             ForbidStates(forbidden, N, cost_eval_step=cost_eval_step, cost_multiplier=0.35)]
    cost_names = ["TargetStateInfidelity", "ForbidStates"]
    if with_time_cost:
        costs.append(TargetStateInfidelityTime(N, target_states, neglect_relative_pahse=neglect_phase,
                                               cost_eval_step=cost_eval_step, cost_multiplier=0.2))
        cost_names.append("TargetStateInfidelityTime")
    if with_control_costs:
        costs.append(ControlNorm(K, M, cost_multiplier=0.11, max_control_norms=max_norms))
        costs.append(ControlVariation(K, M, cost_multiplier=0.07, max_control_norms=max_norms, order=1))
        costs.append(ControlVariation(K, M, cost_multiplier=0.05, max_control_norms=max_norms, order=2))
        cost_names += ["ControlNorm", "ControlVariation1", "ControlVariation2"]
    hamiltonian = make_hamiltonian(h0, drives, complex_controls)

    def forward(c):
        r = evolve_schroedinger_discrete(T, hamiltonian, initial_states, N, controls=c,
                                         cost_eval_step=cost_eval_step, costs=costs,
                                         magnus_policy=POLICIES[order])
        return r.error
    res = evolve_schroedinger_discrete(T, hamiltonian, initial_states, N, controls=controls,
                                       cost_eval_step=cost_eval_step, costs=costs,
                                       magnus_policy=POLICIES[order])
    out = dict(n=n, S=S, K=K, M=M, N=N, order=order, complex_controls=complex_controls,
               cost_eval_step=cost_eval_step, T=T, h0=h0, drives=drives, controls=controls,
               initial_states=initial_states, target_states=target_states, forbidden_states=forbidden,
               max_control_norms=max_norms, cost_names=np.array(cost_names), neglect_phase=neglect_phase,
               error=res.error, final_states=res.final_states)
    if do_fd:
        out["fd_grad"] = fd_grad(forward, controls)
    return out


def lindblad_case(seed, n, D, K, M, N, complex_controls, cost_eval_step, L=2, F=2, do_fd=True,
                  with_hamiltonian=True, with_lindblad=True):
    rng = np.random.default_rng(seed)
    T = 0.6 * (N - 1)
    h0 = rand_herm(rng, n) * 0.7
    drives = []
    for k in range(K):
        if complex_controls:
            c = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)), 1) * 0.3
        else:
            c = rand_herm(rng, n) * 0.3
        drives.append(c)
    drives = np.stack(drives)
    if complex_controls:
        controls = (rng.normal(0, 0.5, (M, K)) + 1j * rng.normal(0, 0.5, (M, K))) / np.sqrt(2)
    else:
        controls = rng.normal(0, 0.5, (M, K))
    gammas = rng.uniform(0.05, 0.3, L)
    ops = (rng.standard_normal((L, n, n)) + 1j * rng.standard_normal((L, n, n))) * 0.5
    cols = haar_columns(rng, n, n)
    init = np.stack([np.outer(cols[:, d], cols[:, d].conj()) for d in range(D)])
    tcols = haar_columns(rng, n, n)
    target = np.stack([np.outer(tcols[:, d], tcols[:, d].conj()) for d in range(D)])
    forbidden = np.stack([np.stack([np.outer(tcols[:, (D + d * F + f) % n], tcols[:, (D + d * F + f) % n].conj())
                                    for f in range(F)]) for d in range(D)])
    costs = [TargetDensityInfidelity(target, cost_multiplier=0.8),
             ForbidDensities(forbidden, N, cost_eval_step=cost_eval_step, cost_multiplier=0.4),
             TargetDensityInfidelityTime(N, target, cost_eval_step=cost_eval_step, cost_multiplier=0.3)]
    hamiltonian = make_hamiltonian(h0, drives, complex_controls) if with_hamiltonian else None
    lindblad_data = (lambda time: (gammas, ops)) if with_lindblad else None

    def forward(c):
        r = evolve_lindblad_discrete(T, init, N, controls=c, cost_eval_step=cost_eval_step, costs=costs,
                                     hamiltonian=hamiltonian, lindblad_data=lindblad_data)
        return r.error
    res = evolve_lindblad_discrete(T, init, N, controls=controls, cost_eval_step=cost_eval_step, costs=costs,
                                   hamiltonian=hamiltonian, lindblad_data=lindblad_data)
    out = dict(n=n, D=D, K=K, M=M, N=N, complex_controls=complex_controls, cost_eval_step=cost_eval_step,
               T=T, h0=h0, drives=drives, controls=controls, gammas=gammas, lindblad_ops=ops,
               initial_densities=init, target_densities=target, forbidden_densities=forbidden,
               with_hamiltonian=with_hamiltonian, with_lindblad=with_lindblad,
               error=res.error, final_densities=res.final_densities)
    if do_fd:
        out["fd_grad"] = fd_grad(forward, controls, eps=1e-5)
    return out


def save(name, d):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote", path, "error" in d and d["error"])


def main():
    # --- unit-level vectors -------------------------------------------------------------------
    rng = np.random.default_rng(1234)
    unit = {}
    mats, exps = [], []
    for n, scale in ((2, 0.3), (3, 1.0), (5, 4.0), (8, 9.0), (8, 40.0), (16, 2.0), (16, 23.0)):
        a = -1j * rand_herm(rng, n)
        a *= scale / np.abs(a).sum(axis=0).max()
        mats.append(a)
        exps.append(expm_pade(a))
    g = rng.standard_normal((6, 6)) + 1j * rng.standard_normal((6, 6))      # non-normal input
    g *= 3.0 / np.abs(g).sum(axis=0).max()
    mats.append(g)
    exps.append(expm_pade(g))
    for i, (a, e) in enumerate(zip(mats, exps)):
        unit["expm_in_%d" % i] = a
        unit["expm_out_%d" % i] = e
    unit["expm_count"] = len(mats)
    # Magnus on a genuinely time-dependent generator (the reference test only covers constant A)
    a0 = -1j * rand_herm(rng, 4)
    a1 = -1j * rand_herm(rng, 4)
    a2 = -1j * rand_herm(rng, 4)
    afun = lambda t: a0 + t * a1 + t * t * a2
    unit["magnus_a0"], unit["magnus_a1"], unit["magnus_a2"] = a0, a1, a2
    unit["magnus_dt"], unit["magnus_t"] = 0.3, 0.45
    unit["magnus_m2"] = magnus_m2(afun, 0.3, 0.45)
    unit["magnus_m4"] = magnus_m4(afun, 0.3, 0.45)
    unit["magnus_m6"] = magnus_m6(afun, 0.3, 0.45)
    # interpolation incl. extrapolation on both sides
    xs = np.linspace(0, 3.0, 7)
    ys = rng.standard_normal((7, 2)) + 1j * rng.standard_normal((7, 2))
    q = np.array([-0.2, 0.0, 0.1, 0.5, 1.49, 1.5, 2.99, 3.0, 3.3])
    unit["interp_xs"], unit["interp_ys"], unit["interp_q"] = xs, ys, q
    unit["interp_out"] = np.stack([interpolate_linear_set(x, xs, ys) for x in q])
    # Lindbladian
    rho = rng.standard_normal((2, 4, 4)) + 1j * rng.standard_normal((2, 4, 4))
    hh = rand_herm(rng, 4)
    gam = rng.uniform(0.1, 1.0, 3)
    lops = rng.standard_normal((3, 4, 4)) + 1j * rng.standard_normal((3, 4, 4))
    unit["lind_rho"], unit["lind_h"], unit["lind_gam"], unit["lind_ops"] = rho, hh, gam, lops
    unit["lind_out"] = get_lindbladian(rho, gam, hh, lops)
    unit["lind_out_noh"] = get_lindbladian(rho, gam, None, lops)
    unit["lind_out_nol"] = get_lindbladian(rho, None, hh, None)
    # RKDP5 on the reference's exact ODE (tests/test_core.py:380-393) and on a linear complex system
    rhs = lambda x, y: ((-2 * x * y + 9 * x ** 2) / (2 * y + x ** 2 + 1))
    unit["rkdp5_ode_y1"] = integrate_rkdp5(rhs, np.array([10.]), 0, np.array((-3.,)))[0]
    lm = -1j * rand_herm(rng, 3) - 0.1 * np.eye(3)
    y0 = rng.standard_normal((2, 3, 3)) + 1j * rng.standard_normal((2, 3, 3))
    unit["rkdp5_lin_m"], unit["rkdp5_lin_y0"] = lm, y0
    unit["rkdp5_lin_y1"] = integrate_rkdp5(lambda x, y: np.matmul(lm, y) * (1 + 0.3 * x), np.array([0.8]), 0.1, y0)
    # clip (tests/test_core.py:11-19) on random complex controls
    cc = rng.standard_normal((9, 3)) * 3 + 1j * rng.standard_normal((9, 3)) * 3
    unit["clip_in"] = cc.copy()
    mx = np.array([1.5, 2.0, 4.0])
    clip_control_norms(cc, mx)
    unit["clip_max"], unit["clip_out"] = mx, cc
    # Adam / SGD trajectories on a fixed gradient sequence
    grads = rng.standard_normal((5, 6))
    ad = Adam(learning_rate=1e-2)
    ad.gradient_moment = np.zeros(6)
    ad.gradient_square_moment = np.zeros(6)
    ad.iteration_count = 0
    p = np.ones(6)
    traj = []
    for gi in grads:
        p = ad.update(gi, p)
        traj.append(p)
    unit["adam_grads"], unit["adam_traj"] = grads, np.stack(traj)
    ad2 = Adam(learning_rate=5e-2, clip_grads=0.5, scale_grads=2.0, learning_rate_decay=3.0)
    ad2.gradient_moment = np.zeros(6)
    ad2.gradient_square_moment = np.zeros(6)
    ad2.iteration_count = 0
    p = np.ones(6)
    traj = []
    for gi in grads:
        p = ad2.update(gi, p)
        traj.append(p)
    unit["adam2_traj"] = np.stack(traj)
    # ControlBandwidthMax value
    ctl = rng.standard_normal((32, 2))
    unit["cbm_controls"] = ctl
    unit["cbm_value"] = ControlBandwidthMax(2, 32, 10.0, np.array([0.4, 0.9]), cost_multiplier=0.7).cost(ctl, None, 0)
    save("unit_vectors", unit)

    # --- reference examples: iteration-0 forward -------------------------------------------
    # examples/0_transmon_pi.py:18-40 (cfg1): flat complex initial controls (common.py:110-142)
    a_op, ad_op = get_annihilation_operator(2), get_creation_operator(2)
    ham = lambda c, t: SIGMA_Z / 2 + c[0] * a_op + np.conjugate(c[0]) * ad_op
    init = np.stack((np.array([[1], [0]]),))
    targ = np.stack((np.array([[0], [1]]),))
    ctl = gen_controls_flat(True, 1, 11, 10, np.ones(1))
    r = evolve_schroedinger_discrete(10, ham, init, 11, controls=ctl, costs=[TargetStateInfidelity(targ)])
    f0 = lambda c: evolve_schroedinger_discrete(10, ham, init, 11, controls=c,
                                                costs=[TargetStateInfidelity(targ)]).error
    save("cfg1_transmon_pi", dict(controls=ctl, error=r.error, final_states=r.final_states,
                                  fd_grad=fd_grad(f0, ctl)))
    # examples/tutorial.py:44-160: n=4, M=N=100, T=15, M2 -> notebook log 9.99980846e-01
    PI_2 = 2 * np.pi
    W_T, W_C, CHI = PI_2 * 5.6640, PI_2 * 4.4526, PI_2 * -2.194
    ALPHA_BY_2, KAPPA_BY_2, CHIP_BY_2 = PI_2 * -2.36e-1, PI_2 * -3.7e-6, PI_2 * -1.9e-6
    A, AD, AI = get_annihilation_operator(2), get_creation_operator(2), np.eye(2)
    B, BD, BI = A, AD, AI
    HS = (W_C * krons(matmuls(AD, A), BI) + KAPPA_BY_2 * krons(matmuls(AD, AD, A, A), BI)
          + W_T * krons(AI, matmuls(BD, B)) + ALPHA_BY_2 * krons(AI, matmuls(BD, BD, B, B))
          + CHI * krons(matmuls(AD, A), matmuls(BD, B))
          + CHIP_BY_2 * krons(matmuls(AD, AD, A, A), matmuls(BD, B)))
    C0, C1 = krons(A, BI), krons(AI, B)
    hamt = lambda c, t: (HS + c[0] * C0 + np.conjugate(c[0]) * C0.conj().T
                         + c[1] * C1 + np.conjugate(c[1]) * C1.conj().T)
    z2, o2 = np.array([[1.], [0.]]), np.array([[0.], [1.]])
    init = np.stack((krons(z2, z2),))
    targ = np.stack((krons(o2, z2),))
    ctl = gen_controls_flat(True, 2, 100, 15, np.ones(2))
    r = evolve_schroedinger_discrete(15, hamt, init, 100, controls=ctl, costs=[TargetStateInfidelity(targ)])
    save("tutorial_iter0", dict(h0=HS, drives=np.stack([C0, C1]), controls=ctl, initial_states=init,
                                target_states=targ, error=r.error, final_states=r.final_states,
                                notebook_error=9.99980846e-01))

    # --- randomised Schroedinger cases -----------------------------------------------------
    cases = [
        dict(seed=0, n=3, S=1, K=1, M=6, N=6, order=2, complex_controls=False, cost_eval_step=1, stiff=False),
        dict(seed=1, n=4, S=2, K=2, M=5, N=9, order=4, complex_controls=True, cost_eval_step=2, stiff=False),
        dict(seed=2, n=5, S=3, K=2, M=9, N=7, order=6, complex_controls=True, cost_eval_step=3, stiff=False),
        dict(seed=3, n=6, S=2, K=3, M=7, N=8, order=4, complex_controls=False, cost_eval_step=1, stiff=True),
        dict(seed=4, n=8, S=4, K=2, M=8, N=8, order=2, complex_controls=True, cost_eval_step=1, stiff=True,
             neglect_phase=True),
        dict(seed=5, n=7, S=2, K=1, M=4, N=10, order=6, complex_controls=False, cost_eval_step=4, stiff=True),
        dict(seed=6, n=16, S=3, K=2, M=12, N=12, order=4, complex_controls=True, cost_eval_step=1, stiff=False,
             do_fd=False),
        dict(seed=7, n=20, S=2, K=2, M=10, N=16, order=2, complex_controls=False, cost_eval_step=5, stiff=True,
             do_fd=False),
    ]
    for i, kw in enumerate(cases):
        save("schroedinger_case_%d" % i, schroedinger_case(**kw))

    # iSWAP known answer (tests/test_core.py:450-469) through the reference itself
    hm = 0.5 * (np.kron(SIGMA_X, SIGMA_X) + np.kron(SIGMA_Y, SIGMA_Y))
    eye4 = np.eye(4, dtype=np.complex128)
    init = np.stack([eye4[:, i:i + 1] for i in range(4)])
    isw = {}
    for order, pol in POLICIES.items():
        r = evolve_schroedinger_discrete(np.pi / 2, lambda c, t: hm, init, 1000, magnus_policy=pol)
        isw["final_states_m%d" % order] = r.final_states
    save("iswap_schroedinger", isw)

    # --- Lindblad cases --------------------------------------------------------------------
    lcases = [
        dict(seed=10, n=2, D=1, K=1, M=5, N=2, complex_controls=True, cost_eval_step=1, L=1, F=1),
        dict(seed=11, n=3, D=2, K=2, M=4, N=4, complex_controls=False, cost_eval_step=1),
        dict(seed=12, n=4, D=2, K=1, M=6, N=5, complex_controls=True, cost_eval_step=2),
        dict(seed=13, n=3, D=1, K=1, M=4, N=3, complex_controls=False, cost_eval_step=1, with_lindblad=False),
    ]
    for i, kw in enumerate(lcases):
        save("lindblad_case_%d" % i, lindblad_case(**kw))
    # amplitude damping (tests/test_core.py:124-148) with fixed a0,b0; iSWAP on densities (:86-106)
    gamma = 2
    sp = np.array([[0, 1], [0, 0]])
    a0, b0 = 0.3, 0.4
    rho0 = np.array(((a0, b0), (b0, 1 - a0)))
    r = evolve_lindblad_discrete(1., np.stack((rho0,)), 2,
                                 lindblad_data=lambda t: (np.array((gamma,)), np.stack((sp,))))
    initd = np.matmul(init, conjugate_transpose(init))
    r2 = evolve_lindblad_discrete(np.pi / 2, initd, 2, hamiltonian=lambda c, t: hm)
    save("lindblad_known", dict(ad_rho0=rho0, ad_gamma=gamma, ad_final=r.final_densities,
                                iswap_final=r2.final_densities))


if __name__ == "__main__":
    main()
