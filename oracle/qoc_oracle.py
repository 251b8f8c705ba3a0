"""
qoc_oracle.py - CPU restatement (torch complex128 + torch.autograd) of the reference's GRAPE
propagate-and-differentiate hot path.

*** TEST INFRASTRUCTURE ONLY. ***  Nothing under `qoc_b200/` may import this module.  It is imported
by `tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s cpu_baseline / `--impl reference` legs,
and only as the checker / the timed CPU baseline - never as the product path.

Parity status: PINNED (forward) against outputs of the unmodified reference run in the build
container (`tests/golden/*.npz`, produced by `tests/golden/make_golden.py`, which imports
/root/reference with numpy standing in for `autograd.numpy`) and against the reference's own
known answers (iSWAP `tests/test_core.py:450-469`, amplitude damping `:124-148`, Lindbladian hand value
`:300-310`, RKDP5 exact ODE `:380-393`, cost known answers `tests/test_standard.py`, the recorded
`total error = 9.99980846e-01` of `examples/tutorial.ipynb:313`).  The reference's gradient is
produced by the third-party HIPS `autograd` package (setup.py:11, unpinned, absent from this image and
from /root/reference); no reference test pins a gradient value (`tests/test_core.py:567-571`).  Here the
gradient is torch reverse-mode AD over the same operation sequence, and is pinned to central finite
differences of the *reference* forward (golden `fd_grad`, ~1e-8) and to an independent hand-adjoint
(`oracle/adjoint_model.py`, ~1e-13).

Every function cites the reference lines it follows (paths relative to /root/reference).
Conventions: `grads` returned by `*_cost_and_grad` are dE/dx + i dE/dy for complex controls - the
value `_esdj_wrap` holds after its conjugate (qoc/core/schroedingerdiscrete.py:320-324), which is what
torch's `.grad` yields directly.
"""
import math

import numpy as np
import torch

CDT = torch.complex128

# --- qoc/standard/functions/expm.py:86-101 (Pade-13 coefficients), :192-207 (theta_13) ---------
PADE13_B = (64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
            129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
            40840800., 960960., 16380., 182., 1.)
THETA13 = 5.371920351148152


def one_norm(a):
    """qoc/standard/functions/expm.py:103-116 - max column sum of |a_ij|."""
    return torch.max(torch.sum(torch.abs(a), dim=0))


def pade13(a, ident):
    """qoc/standard/functions/expm.py:153-159."""
    b = PADE13_B
    a2 = a @ a
    a4 = a2 @ a2
    a6 = a2 @ a4
    u = a @ (a6 @ (b[13] * a6 + b[11] * a4 + b[9] * a2) + b[7] * a6 + b[5] * a4 + b[3] * a2) + b[1] * a
    v = a6 @ (b[12] * a6 + b[10] * a4 + b[8] * a2) + b[6] * a6 + b[4] * a4 + b[2] * a2 + b[0] * ident
    return u, v


def expm_pade(a):
    """qoc/standard/functions/expm.py:210-252.  The order loop there has no `break` (:230-234), so any
    norm below theta_13 ends at order 13; otherwise order 13 with s = max(0, ceil(log2(norm/theta_13)))
    (:238-241).  The net algorithm is always Pade-13; the scaling count is not differentiated."""
    norm = float(one_norm(a.detach()))
    scale = 0
    if not norm < THETA13:
        scale = max(0, int(math.ceil(math.log2(norm / THETA13))))
        a = a * (2.0 ** -scale)
    ident = torch.eye(a.shape[0], dtype=a.dtype)
    u, v = pade13(a, ident)
    r = torch.linalg.solve(v - u, v + u)            # :246
    for _ in range(scale):                         # :249-250
        r = r @ r
    return r


def commutator(a, b):
    """qoc/standard/functions/convenience.py:16-29."""
    return a @ b - b @ a


def rms_norm(x):
    """qoc/standard/functions/convenience.py:77-91 (returned as a real tensor; the reference value has
    an exactly-zero imaginary part)."""
    return torch.sqrt(torch.sum((x * torch.conj(x)).real) / x.numel())


# --- qoc/core/mathmethods.py:14-67 ---------------------------------------------------------------
def interpolate_linear_set(x, xs, ys):
    """Linear interpolation with two-point extrapolation outside [xs[0], xs[-1]] (:54-59); interior
    index = first k with x <= xs[k] (:64)."""
    if x <= xs[0]:
        i0, i1 = 0, 1
    elif x >= xs[-1]:
        i0, i1 = len(xs) - 2, len(xs) - 1
    else:
        i1 = int(np.argmax(x <= xs))
        i0 = i1 - 1
    return ys[i0] + ((ys[i1] - ys[i0]) / (xs[i1] - xs[i0])) * (x - xs[i0])


# --- qoc/core/mathmethods.py:72-164 --------------------------------------------------------------
_S3, _S15 = math.sqrt(3.0), math.sqrt(15.0)


def magnus(a, dt, t, order):
    """`a` maps time -> generator matrix.  order 2: :74-93, order 4: :100-122, order 6: :134-164."""
    if order == 2:
        return dt * a(t + dt * 0.5)
    if order == 4:
        a1 = a(t + dt * (0.5 - _S3 / 6))
        a2 = a(t + dt * (0.5 + _S3 / 6))
        return (dt / 2) * (a1 + a2) + (_S3 / 12) * (dt ** 2) * commutator(a2, a1)
    if order == 6:
        a1 = a(t + dt * (0.5 - _S15 / 10))
        a2 = a(t + dt * 0.5)
        a3 = a(t + dt * (0.5 + _S15 / 10))
        b1 = dt * a2
        b2 = (_S15 / 3) * dt * (a3 - a1)
        b3 = (10.0 / 3) * dt * (a3 - 2 * a2 + a1)
        c12 = commutator(b1, b2)
        return b1 + 0.5 * b3 + (1.0 / 240) * commutator(-20 * b1 - b3 + c12,
                                                       b2 - (1.0 / 60) * commutator(b1, 2 * b3 + c12))
    raise ValueError("Unrecognized magnus order {}".format(order))


# --- cost restatements: qoc/standard/costs/*.py ---------------------------------------------------
def _dagger(x):
    return torch.conj(torch.swapaxes(x, -1, -2))


class TargetStateInfidelity(object):
    """qoc/standard/costs/targetstateinfidelity.py:12-63."""
    requires_step_evaluation = False

    def __init__(self, target_states, neglect_relative_pahse=False, cost_multiplier=1.):
        t = torch.as_tensor(np.asarray(target_states), dtype=CDT)
        self.state_count = t.shape[0]
        self.target_states_dagger = _dagger(t)
        self.neglect = neglect_relative_pahse
        self.cost_multiplier = cost_multiplier
        self.norm = 1.0

    def cost(self, controls, states, step):
        ip = (self.target_states_dagger @ states)[:, 0, 0]
        if not self.neglect:
            tot = torch.sum(ip)
            fid = (tot * torch.conj(tot)).real / self.state_count ** 2
        else:
            fid = torch.sum((ip * torch.conj(ip)).real) / self.state_count
        return (1 - fid) / self.norm * self.cost_multiplier


class TargetStateInfidelityTime(TargetStateInfidelity):
    """qoc/standard/costs/targetstateinfidelitytime.py:13-73; normalised by
    cost_eval_count = (N-1)//cost_eval_step (:41)."""
    requires_step_evaluation = True

    def __init__(self, system_eval_count, target_states, neglect_relative_pahse=False,
                 cost_eval_step=1, cost_multiplier=1.):
        super().__init__(target_states, neglect_relative_pahse, cost_multiplier)
        self.norm = (system_eval_count - 1) // cost_eval_step


class ForbidStates(object):
    """qoc/standard/costs/forbidstates.py:12-81."""
    requires_step_evaluation = True

    def __init__(self, forbidden_states, system_eval_count, cost_eval_step=1, cost_multiplier=1.):
        self.fdag = [_dagger(torch.as_tensor(np.asarray(f), dtype=CDT)) for f in forbidden_states]
        self.norm = ((system_eval_count - 1) // cost_eval_step) * len(self.fdag)
        self.cost_multiplier = cost_multiplier

    def cost(self, controls, states, step):
        tot = 0
        for i, fd in enumerate(self.fdag):
            ip = (fd @ states[i])[:, 0, 0]
            tot = tot + torch.sum((ip * torch.conj(ip)).real) / fd.shape[0]
        return tot / self.norm * self.cost_multiplier


class TargetDensityInfidelity(object):
    """qoc/standard/costs/targetdensityinfidelity.py:12-69 (note the 1/hilbert_size, :65)."""
    requires_step_evaluation = False

    def __init__(self, target_densities, cost_multiplier=1.):
        t = torch.as_tensor(np.asarray(target_densities), dtype=CDT)
        self.density_count, self.hilbert_size = t.shape[0], t.shape[1]
        self.tdag = _dagger(t)
        self.cost_multiplier = cost_multiplier
        self.norm = 1.0

    def cost(self, controls, densities, step):
        prods = self.tdag @ densities
        tot = 0
        for p in prods:
            tot = tot + torch.abs(torch.trace(p))
        return (1 - tot / (self.density_count * self.hilbert_size)) / self.norm * self.cost_multiplier


class TargetDensityInfidelityTime(TargetDensityInfidelity):
    """qoc/standard/costs/targetdensityinfidelitytime.py:13-76; `requires_step_evaluation = False`
    there (:30), i.e. evaluated once at the final step but still divided by cost_eval_count."""
    requires_step_evaluation = False

    def __init__(self, system_eval_count, target_densities, cost_eval_step=1, cost_multiplier=1.):
        super().__init__(target_densities, cost_multiplier)
        self.norm = (system_eval_count - 1) // cost_eval_step


class ForbidDensities(object):
    """qoc/standard/costs/forbiddensities.py:12-85."""
    requires_step_evaluation = True

    def __init__(self, forbidden_densities, system_eval_count, cost_eval_step=1, cost_multiplier=1.):
        self.fdag = [_dagger(torch.as_tensor(np.asarray(f), dtype=CDT)) for f in forbidden_densities]
        self.norm = ((system_eval_count - 1) // cost_eval_step) * len(self.fdag)
        self.hilbert_size = self.fdag[0].shape[-1]
        self.cost_multiplier = cost_multiplier

    def cost(self, controls, densities, step):
        tot = 0
        for i, fd in enumerate(self.fdag):
            sub = 0
            for f in fd:
                ip = torch.trace(f @ densities[i]) / self.hilbert_size
                sub = sub + (ip * torch.conj(ip)).real
            tot = tot + sub / fd.shape[0]
        return tot / self.norm * self.cost_multiplier


class ControlNorm(object):
    """qoc/standard/costs/controlnorm.py:11-73."""
    requires_step_evaluation = False

    def __init__(self, control_count, control_eval_count, control_weights=None, cost_multiplier=1.,
                 max_control_norms=None):
        self.w = None if control_weights is None else torch.as_tensor(np.asarray(control_weights))
        self.size = control_eval_count * control_count
        self.mx = None if max_control_norms is None else torch.as_tensor(np.asarray(max_control_norms, dtype=np.float64))
        self.cost_multiplier = cost_multiplier

    def cost(self, controls, states, step):
        c = controls
        if self.mx is not None:
            c = c / self.mx
        if self.w is not None:
            c = c * self.w
        return torch.sum((c * torch.conj(c)).real) / self.size * self.cost_multiplier


class ControlVariation(object):
    """qoc/standard/costs/controlvariation.py:11-75."""
    requires_step_evaluation = False

    def __init__(self, control_count, control_eval_count, cost_multiplier=1., max_control_norms=None, order=1):
        self.mx = None if max_control_norms is None else torch.as_tensor(np.asarray(max_control_norms, dtype=np.float64))
        self.order = order
        self.norm = control_count * (control_eval_count - order) * (2 ** order)
        self.cost_multiplier = cost_multiplier

    def cost(self, controls, states, step):
        c = controls if self.mx is None else controls / self.mx
        d = torch.diff(c, n=self.order, dim=0)
        return torch.sum((d * torch.conj(d)).real) / self.norm * self.cost_multiplier


class ControlArea(object):
    """qoc/standard/costs/controlarea.py:11-67 (only the max_control_norms branch is executable in the
    reference: the other branch assigns `normalized_control` and then reads `normalized_controls`, :55-64)."""
    requires_step_evaluation = False

    def __init__(self, control_count, control_eval_count, cost_multiplier=1., max_control_norms=None):
        self.control_count = control_count
        self.size = control_count * control_eval_count
        self.mx = None if max_control_norms is None else torch.as_tensor(np.asarray(max_control_norms, dtype=np.float64))
        self.cost_multiplier = cost_multiplier

    def cost(self, controls, states, step):
        if self.mx is None:
            raise NameError("name 'normalized_controls' is not defined")
        c = controls / self.mx
        tot = 0
        for i in range(self.control_count):
            tot = tot + torch.abs(torch.sum(c[:, i]))
        return tot / self.size * self.cost_multiplier


class ControlBandwidthMax(object):
    """qoc/standard/costs/controlbandwidthmax.py:11-77."""
    requires_step_evaluation = False

    def __init__(self, control_count, control_eval_count, evolution_time, max_bandwidths, cost_multiplier=1.):
        self.max_bandwidths = np.asarray(max_bandwidths)
        self.control_count = control_count
        self.freqs = np.fft.fftfreq(control_eval_count, d=evolution_time / (control_eval_count - 1))
        self.cost_multiplier = cost_multiplier

    def cost(self, controls, states, step):
        tot = 0
        for i, bw in enumerate(self.max_bandwidths):
            mag = torch.abs(torch.fft.fft(controls[:, i]))
            idx = torch.as_tensor(np.nonzero(self.freqs >= bw)[0])
            pen = mag[idx]
            tot = tot + torch.sum(pen) / (idx.shape[0] * torch.max(pen))
        return tot / self.control_count * self.cost_multiplier


# --- Schroedinger evolution: qoc/core/schroedingerdiscrete.py:356-502 ------------------------------
def evolve_step_schroedinger(dt, hamiltonian, states, time, control_eval_times, controls, order):
    """One slice (:441-502): generator -1j*H(interp(t), t) (:483-486), Magnus (:488-497),
    expm (:499), U @ states (:500)."""
    def gen(t_):
        c = None if controls is None else interpolate_linear_set(t_, control_eval_times, controls)
        return -1j * hamiltonian(c, t_)
    return expm_pade(magnus(gen, dt, time, order)) @ states


def evaluate_schroedinger(controls, hamiltonian, initial_states, costs, evolution_time, system_eval_count,
                          order=2, cost_eval_step=1, keep_states=False):
    """Total cost of one evolution (:356-438).  Step costs at steps with step % cost_eval_step == 0 and
    step != 0, final step included (:405-415); propagation on every step but the last (:419-425); then
    non-step costs on the final states (:429-432).  Returns (error, final_states[, all_states])."""
    n_ctl = 0 if controls is None else controls.shape[0]
    control_eval_times = np.linspace(0, evolution_time, n_ctl)      # qoc/models/programstate.py:41
    dt = evolution_time / (system_eval_count - 1)                   # :44
    final_step = system_eval_count - 1
    states = torch.as_tensor(np.asarray(initial_states), dtype=CDT)
    step_costs = [c for c in costs if c.requires_step_evaluation]   # programstate.py:52-60
    error = 0
    trail = []
    for step in range(system_eval_count):
        if keep_states:
            trail.append(states.detach().clone())
        if step % cost_eval_step == 0 and step != 0:
            for c in step_costs:
                error = error + c.cost(controls, states, step)
        if step != final_step:
            states = evolve_step_schroedinger(dt, hamiltonian, states, step * dt, control_eval_times,
                                              controls, order)
    for c in costs:
        if not c.requires_step_evaluation:
            error = error + c.cost(controls, states, final_step)
    if keep_states:
        return error, states, trail
    return error, states


def schroedinger_cost_and_grad(controls, hamiltonian, initial_states, costs, evolution_time,
                               system_eval_count, order=2, cost_eval_step=1):
    """`ans_jacobian(_evaluate_schroedinger_discrete, 0)` (qoc/standard/utils/autogradutil.py:10-31)
    followed by the wrapper's conjugate (schroedingerdiscrete.py:320-324)."""
    c = torch.tensor(np.asarray(controls), requires_grad=True)
    error, states = evaluate_schroedinger(c, hamiltonian, initial_states, costs, evolution_time,
                                          system_eval_count, order, cost_eval_step)
    error.backward()
    return float(error.detach()), c.grad.numpy().copy(), states.detach().numpy().copy()


# --- Lindblad evolution: qoc/core/mathmethods.py:169-480, qoc/core/lindbladdiscrete.py:357-495 ------
def get_lindbladian(densities, dissipators=None, hamiltonian=None, operators=None):
    """qoc/core/mathmethods.py:169-206."""
    lind = 0
    if hamiltonian is not None:
        lind = -1j * commutator(hamiltonian, densities)
    if dissipators is not None and operators is not None:
        odag = _dagger(operators)
        oprod = odag @ operators
        for i in range(operators.shape[0]):
            lind = lind + dissipators[i] * (operators[i] @ densities @ odag[i]
                                            - 0.5 * (oprod[i] @ densities) - 0.5 * (densities @ oprod[i]))
    return lind


# Dormand-Prince tableau, qoc/core/mathmethods.py:211-260
_A = ((), (1 / 5,), (3 / 40, 9 / 40), (44 / 45, -56 / 15, 32 / 9),
      (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
      (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656))
_C = (0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1)
_B = (35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0)
_BH = (5179 / 57600, 0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40)
_D = (-12715105075 / 11282082432, 0, 87487479700 / 32700410799, -10690763975 / 1880347072,
      701980252875 / 199316789632, -1453857185 / 822651844, 69997945 / 29380423)
_ERROR_EXP = -1 / 5


def rkdp5_step(h, rhs, x0, y0, k1):
    """qoc/core/mathmethods.py:307-349."""
    ks = [k1]
    for i in range(1, 6):
        acc = 0
        for j in range(i):
            acc = acc + _A[i][j] * ks[j]
        ks.append(rhs(x0 + _C[i] * h, y0 + h * acc))
    y1 = y0 + h * (_B[0] * ks[0] + _B[2] * ks[2] + _B[3] * ks[3] + _B[4] * ks[4] + _B[5] * ks[5])
    ks.append(rhs(x0 + h, y1))
    y1h = y0 + h * (_BH[0] * ks[0] + _BH[2] * ks[2] + _BH[3] * ks[3] + _BH[4] * ks[4]
                    + _BH[5] * ks[5] + _BH[6] * ks[6])
    return ks, y1, y1h


def rkdp5_dense(ks, x0, x1, x_eval, y0, y1):
    """qoc/core/mathmethods.py:263-304 for a single evaluation point."""
    h = x1 - x0
    r2 = y1 - y0
    r3 = y0 + h * ks[0] - y1
    r4 = 2 * (y1 - y0) - h * (ks[0] + ks[6])
    r5 = h * (_D[0] * ks[0] + _D[2] * ks[2] + _D[3] * ks[3] + _D[4] * ks[4] + _D[5] * ks[5] + _D[6] * ks[6])
    th = (x_eval - x0) / h
    return y0 + th * (r2 + r3) - th ** 2 * (r3 - r4 - r5) - th ** 3 * (r4 + 2 * r5) + th ** 4 * r5


def integrate_rkdp5(rhs, x_final, x_initial, y_initial, atol=1e-12, stats=None, freeze_steps=False):
    """qoc/core/mathmethods.py:352-480 for one output point (`x_eval = [x_final]`, the only way the hot
    path calls it, lindbladdiscrete.py:427).  rtol = 0.  Step sizes are tensors so that autograd also
    differentiates the controller, as the reference's tape does; `freeze_steps=True` detaches them
    (used to quantify that contribution)."""
    t = lambda v: torch.as_tensor(v, dtype=torch.float64)
    f0 = rhs(t(x_initial), y_initial)
    d0, d1 = rms_norm(y_initial), rms_norm(f0)
    if d0 < 1e-5 or d1 < 1e-5:
        h0 = t(1e-6)
    else:
        h0 = 0.01 * d0 / d1
    f1 = rhs(x_initial + h0, y_initial + h0 * f0)
    d2 = rms_norm(f1 - f0) / h0
    if torch.maximum(d1, d2) <= 1e-15:
        h1 = torch.maximum(t(1e-6), h0 * 1e-3)
    else:
        h1 = torch.pow(0.01 / torch.maximum(d1, d2), 1 / 6)
    step = torch.minimum(100 * h0, h1)
    x_cur, y_cur, k1 = t(x_initial), y_initial, f0
    out = None
    while x_cur <= x_final:
        rejected = False
        while True:
            if freeze_steps:
                step = step.detach()
            ks, y1, y1h = rkdp5_step(step, rhs, x_cur, y_cur, k1)
            x_new = x_cur + step
            err = rms_norm((y1 - y1h) / atol)
            if stats is not None:
                stats["attempts"] = stats.get("attempts", 0) + 1
            if err < 1:
                if err == 0:
                    fac = t(10.)
                else:
                    fac = torch.minimum(t(10.), 0.9 * torch.pow(err, _ERROR_EXP))
                if rejected:
                    fac = torch.minimum(t(1.), fac)
                step_next = step * fac
                break
            rejected = True
            step = step * torch.maximum(t(0.2), 0.9 * torch.pow(err, _ERROR_EXP))
        if x_cur <= x_final and x_final <= x_new:
            out = rkdp5_dense(ks, x_cur, x_new, x_final, y_cur, y1)   # last hit wins (:462-472 appends)
        if stats is not None:
            stats["accepted"] = stats.get("accepted", 0) + 1
        x_cur, y_cur, k1, step = x_new, y1, ks[6], step_next
    return out


def evaluate_lindblad(controls, hamiltonian, lindblad_data, initial_densities, costs, evolution_time,
                      system_eval_count, cost_eval_step=1, stats=None, freeze_steps=False):
    """qoc/core/lindbladdiscrete.py:357-441 with the rhs closure of :486-492."""
    n_ctl = 0 if controls is None else controls.shape[0]
    control_eval_times = np.linspace(0, evolution_time, n_ctl)
    dt = evolution_time / (system_eval_count - 1)
    final_step = system_eval_count - 1
    densities = torch.as_tensor(np.asarray(initial_densities), dtype=CDT)
    step_costs = [c for c in costs if c.requires_step_evaluation]

    def rhs(time, rho):
        tv = float(time.detach()) if torch.is_tensor(time) else float(time)
        if controls is None:
            c = None
        else:
            # interpolate_linear_set with the (tensor) time kept differentiable: mathmethods.py:54-65
            xs = control_eval_times
            if tv <= xs[0]:
                i0, i1 = 0, 1
            elif tv >= xs[-1]:
                i0, i1 = len(xs) - 2, len(xs) - 1
            else:
                i1 = int(np.argmax(tv <= xs))
                i0 = i1 - 1
            c = controls[i0] + ((controls[i1] - controls[i0]) / (xs[i1] - xs[i0])) * (time - xs[i0])
        h = None if hamiltonian is None else hamiltonian(c, time)
        gam, ops = (None, None) if lindblad_data is None else lindblad_data(time)
        return get_lindbladian(rho, gam, h, ops)

    error = 0
    for step in range(system_eval_count):
        if step % cost_eval_step == 0 and step != 0:
            for c in step_costs:
                error = error + c.cost(controls, densities, step)
        if step != final_step:
            densities = integrate_rkdp5(rhs, step * dt + dt, step * dt, densities, stats=stats,
                                        freeze_steps=freeze_steps)
    for c in costs:
        if not c.requires_step_evaluation:
            error = error + c.cost(controls, densities, final_step)
    return error, densities


def lindblad_cost_and_grad(controls, hamiltonian, lindblad_data, initial_densities, costs, evolution_time,
                           system_eval_count, cost_eval_step=1, freeze_steps=False, stats=None):
    c = torch.tensor(np.asarray(controls), requires_grad=True)
    error, dens = evaluate_lindblad(c, hamiltonian, lindblad_data, initial_densities, costs, evolution_time,
                                    system_eval_count, cost_eval_step, stats=stats, freeze_steps=freeze_steps)
    error.backward()
    return float(error.detach()), c.grad.numpy().copy(), dens.detach().numpy().copy()


# --- helpers shared by tests / bench (problem construction in torch) -------------------------------
def make_hamiltonian(h0, drives, complex_controls):
    """H(u) = H0 + sum_k u_k D_k (real) or H0 + sum_k u_k C_k + conj(u_k) C_k^dagger (complex): the
    form of every hamiltonian in the reference's examples and tests (examples/0_transmon_pi.py:24-26,
    examples/tutorial.py:101-106, tests/test_core.py:529-531,582)."""
    h0 = torch.as_tensor(np.asarray(h0), dtype=CDT)
    dr = torch.as_tensor(np.asarray(drives), dtype=CDT)
    drd = _dagger(dr)

    def hamiltonian(controls, time):
        h = h0
        if controls is None:
            return h
        for k in range(dr.shape[0]):
            if complex_controls:
                h = h + controls[k] * dr[k] + torch.conj(controls[k]) * drd[k]
            else:
                h = h + controls[k] * dr[k]
        return h
    return hamiltonian


def make_lindblad_data(gammas, ops):
    g = torch.as_tensor(np.asarray(gammas, dtype=np.float64))
    o = torch.as_tensor(np.asarray(ops), dtype=CDT)
    return lambda time: (g, o)
