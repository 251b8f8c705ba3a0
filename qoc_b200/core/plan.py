"""
plan.py - host side of the drop-in boundary: turns the reference's program state (user `hamiltonian`
callable, cost objects, states) into a device plan of the C ABI (include/qocb200.h) and evaluates
cost / cost+gradient on the GPU.

What happens here, once per `grape_*` / `evolve_*` call:
  * Hamiltonian structure extraction.  The reference calls the opaque Python callable `hamiltonian(controls, t)`
    at every Magnus node of every slice (qoc/core/schroedingerdiscrete.py:483-486) and differentiates through
    it.  The CUDA path needs the operator structure, so the callable is probed: H0 = h(0), A_k = h(e_k) - H0,
    B_k = h(i e_k) - H0 (complex controls), and real-linearity in (Re u, Im u) plus time-independence are
    verified with random probes.  Every hamiltonian in the reference's examples/tests has this form
    (examples/0_transmon_pi.py:24-26, examples/tutorial.py:101-106, tests/test_core.py:529-531,582).
    A callable that uses its `time` argument is probed at every Magnus node of every
    slice instead and expanded over a small operator basis (`extract_time_dependent_structure`): the device
    then sees H = G0 + sum_c coef_c(node) A_c with per-node coefficients affine in the interpolated controls
    (qocb_set_node_map).  Anything else (non-linear in the controls, more than 16 operator channels) raises
    NotImplementedError - there is no CPU fallback.
  * cost recognition: the qoc cost classes provide device descriptors (`device_terms`) or, for control-only
    costs, an analytic host value+gradient (`control_value_and_grad`).
"""
import ctypes

import numpy as np

from qoc_b200 import _lib
from qoc_b200.models.enums import InterpolationPolicy, MagnusPolicy

_PROBE_RTOL = 1e-10


class NonlinearHamiltonian(NotImplementedError):
    """raised by the structure extraction when the callable is not affine in (Re u, Im u)."""


_MAX_CHANNELS = 16          # kMaxKR of the CUDA kernels
_NODES = {1: (0.5,), 2: (0.5 - np.sqrt(3.0) / 6, 0.5 + np.sqrt(3.0) / 6),
          3: (0.5 - np.sqrt(15.0) / 10, 0.5, 0.5 + np.sqrt(15.0) / 10)}          # mathmethods.py:88,113-114,148-150


def _unit_controls(control_count, complex_controls):
    dtype = np.complex128 if complex_controls else np.float64
    zero = np.zeros(control_count, dtype=dtype)
    units = []
    for part in ((1.0, 1j) if complex_controls else (1.0,)):
        for k in range(control_count):
            e = zero.copy()
            e[k] = part
            units.append(e)
    return zero, units


def _probe_times(evolution_time, system_eval_count=None, magnus_order=None, dense=64, seed=4321):
    """times at which a callable is checked for time dependence: the end points, `dense` evenly spread times, a few random
    ones and - when the slice grid is known - Magnus node times the reference evaluates the callable at
    (qoc/core/schroedingerdiscrete.py:483-497), including the first and last slices, so that windows, steps and envelopes
    that vanish at a handful of fixed probe times are still seen."""
    rng = np.random.default_rng(seed)
    T = float(evolution_time)
    times = [0.0, T, 0.37 * T, 0.731 * T]
    times += list(np.linspace(0.0, T, dense + 2)[1:-1])
    times += list(rng.uniform(0.0, T, 8))
    if system_eval_count is not None and magnus_order is not None and system_eval_count > 1:
        nsl = system_eval_count - 1
        dt = T / nsl
        picks = sorted(set([0, 1, nsl // 2, nsl - 2, nsl - 1] + [int(x) for x in rng.integers(0, nsl, 16)]))
        for j in picks:
            if 0 <= j < nsl:
                times += [(j + c) * dt for c in _NODES[magnus_order // 2]]
    return times


def _is_time_dependent(hamiltonian, zero, units, evolution_time, system_eval_count=None, magnus_order=None):
    times = _probe_times(evolution_time, system_eval_count, magnus_order)
    for u in [zero] + units:
        ref = np.asarray(hamiltonian(u, times[0]), dtype=np.complex128)
        tol = _PROBE_RTOL * max(1.0, np.abs(ref).max())
        for t in times[1:]:
            if not np.allclose(np.asarray(hamiltonian(u, t), dtype=np.complex128), ref, rtol=_PROBE_RTOL, atol=tol):
                return True
    return False


class _RealBasis(object):
    """Streaming rank-revealing Gram-Schmidt of complex matrices over the REALS (the device coefficients are real):
    `add(m)` returns the coefficient vector of m on the basis so far, extending the basis when the residual exceeds
    1e-12 * scale.  Earlier samples have zero weight on later basis vectors by construction."""

    def __init__(self, limit):
        self.vecs, self.limit, self.scale = [], limit, 1.0

    def add(self, m):
        v = np.concatenate([m.real.ravel(), m.imag.ravel()])
        self.scale = max(self.scale, np.abs(v).max())
        coef = np.zeros(len(self.vecs) + 1)
        r = v.copy()
        for _ in range(2):                       # re-orthogonalise once
            for k, b in enumerate(self.vecs):
                c = np.dot(b, r)
                coef[k] += c
                r -= c * b
        nr = np.linalg.norm(r)
        if nr > 1e-12 * self.scale * np.sqrt(v.size):
            if len(self.vecs) >= self.limit:
                raise NotImplementedError(
                    "the time dependence of this hamiltonian needs more than {} operator channels; not supported by "
                    "the CUDA path (no CPU fallback)".format(self.limit))
            self.vecs.append(r / nr)
            coef[-1] = nr
            return coef
        return coef[:-1]

    def matrices(self, shape):
        half = shape[0] * shape[1]
        return np.array([(b[:half] + 1j * b[half:]).reshape(shape) for b in self.vecs]).reshape((-1,) + tuple(shape))


def extract_time_dependent_structure(hamiltonian, control_count, complex_controls, evolution_time, system_eval_count,
                                     magnus_order, seed=1234):
    """Structure of a `hamiltonian(controls, time)` that is affine in the controls at every time:
    returns (G0, channels [KC x n x n], offset [N-1, q, KC], gain [N-1, q, KC, KR]) such that at Magnus node i of
    slice j   H = G0 + sum_c (offset[j,i,c] + sum_r gain[j,i,c,r] x_r) channels[c],   x = [Re u, Im u].
    The callable is evaluated at the node times the reference evaluates it at (schroedingerdiscrete.py:483-497);
    one shared real operator basis serves the drift variation and every control operator, built in one streaming
    pass (memory: the basis and the coefficients only)."""
    q = magnus_order // 2
    nsl = system_eval_count - 1
    dt = evolution_time / nsl
    times = np.array([[(j + c) * dt for c in _NODES[q]] for j in range(nsl)]).ravel()
    zero, units = _unit_controls(control_count, complex_controls) if control_count else (None, [])
    KR = len(units)
    g0 = np.array(hamiltonian(zero, times[0]), dtype=np.complex128)
    if g0.ndim != 2 or g0.shape[0] != g0.shape[1]:
        raise ValueError("hamiltonian(controls, time) must return a square matrix")
    basis = _RealBasis(_MAX_CHANNELS)
    rows = []                                   # per node: [offset coefficients, gain coefficients of control 0, ...]
    for t in times:
        d = np.asarray(hamiltonian(zero, t), dtype=np.complex128)
        row = [basis.add(d - g0)]
        for e in units:
            row.append(basis.add(np.asarray(hamiltonian(e, t), dtype=np.complex128) - d))
        rows.append(row)
    KC = len(basis.vecs)
    if KC == 0:                                 # H == G0 at every node
        return g0, np.zeros((KR,) + g0.shape, dtype=np.complex128)
    channels = basis.matrices(g0.shape)
    offset = np.zeros((len(times), KC))
    gain = np.zeros((len(times), KC, KR))
    for k, row in enumerate(rows):
        offset[k, :len(row[0])] = row[0]
        for r in range(KR):
            gain[k, :len(row[1 + r]), r] = row[1 + r]
    # affine in the controls?  (random controls at a few node times against the expansion)
    rng = np.random.default_rng(seed)
    scale = max(1.0, basis.scale, np.abs(g0).max())
    for trial in range(4 if KR else 0):
        k = int(rng.integers(len(times)))
        u = rng.standard_normal(control_count) * (1.0 + trial)
        if complex_controls:
            u = u + 1j * rng.standard_normal(control_count) * (1.0 + trial)
        x = np.concatenate([u.real, u.imag]) if complex_controls else u
        model = g0 + np.tensordot(offset[k] + gain[k] @ x, channels, axes=(0, 0))
        got = np.asarray(hamiltonian(u.astype(zero.dtype), times[k]), dtype=np.complex128)
        if not np.allclose(got, model, rtol=0, atol=_PROBE_RTOL * scale * (1 + np.abs(x).sum())):
            raise NonlinearHamiltonian(
                "hamiltonian(controls, time) is not affine in (Re u, Im u); a SchroedingerPlan needs that form - "
                "`make_schroedinger_plan` handles non-linear callables (no CPU fallback)")
    return (g0, channels, np.ascontiguousarray(offset.reshape(nsl, q, KC)),
            np.ascontiguousarray(gain.reshape(nsl, q, KC, KR)))


def extract_hamiltonian_structure(hamiltonian, control_count, complex_controls, evolution_time, hilbert_size=None,
                                  seed=1234, system_eval_count=None, magnus_order=None):
    """Returns (h0 [n x n], a_ops [KR x n x n]) with H(x) = H0 + sum_r x_r A_r, x = [Re u, Im u]; for a callable
    that depends on `time` (needs system_eval_count and magnus_order): the 4-tuple of
    `extract_time_dependent_structure`."""
    rng = np.random.default_rng(seed)
    t0, t1 = 0.0, 0.37 * evolution_time
    zero, units = _unit_controls(control_count, complex_controls) if control_count else (None, [])
    if _is_time_dependent(hamiltonian, zero, units, evolution_time, system_eval_count, magnus_order):
        if system_eval_count is None or magnus_order is None:
            raise NotImplementedError("time-dependent hamiltonian callables are not supported by this CUDA path yet "
                                      "(no CPU fallback)")
        return extract_time_dependent_structure(hamiltonian, control_count, complex_controls, evolution_time,
                                                system_eval_count, magnus_order, seed)
    if control_count == 0:
        h0 = np.asarray(hamiltonian(None, t0), dtype=np.complex128)
        return h0, np.zeros((0,) + h0.shape, dtype=np.complex128)
    dtype = zero.dtype
    h0 = np.array(hamiltonian(zero, t0), dtype=np.complex128)
    if h0.ndim != 2 or h0.shape[0] != h0.shape[1]:
        raise ValueError("hamiltonian(controls, time) must return a square matrix")
    a_ops = np.stack([np.array(hamiltonian(e, t0), dtype=np.complex128) - h0 for e in units])
    scale = max(1.0, np.abs(h0).max(), np.abs(a_ops).max())
    for trial in range(3):
        u = rng.standard_normal(control_count) * (1.0 + trial)
        if complex_controls:
            u = u + 1j * rng.standard_normal(control_count) * (1.0 + trial)
        x = np.concatenate([u.real, u.imag]) if complex_controls else u
        model = h0 + np.tensordot(x, a_ops, axes=(0, 0))
        for t in (t0, t1, float(rng.uniform(0.0, evolution_time)), float(evolution_time)):
            got = np.asarray(hamiltonian(u.astype(dtype), t), dtype=np.complex128)
            if not np.allclose(got, model, rtol=0, atol=_PROBE_RTOL * scale * (1 + np.abs(x).sum())):
                raise NonlinearHamiltonian(
                    "hamiltonian(controls, time) is not of the form H0 + sum_k Re(u_k) A_k + Im(u_k) B_k; a SchroedingerPlan "
                    "needs that form - `make_schroedinger_plan` handles non-linear callables (no CPU fallback)")
    return h0, a_ops


class SchroedingerPlan(object):
    """Device plan for `_evaluate_schroedinger_discrete` and its jacobian (schroedingerdiscrete.py:356-438,
    :318).  `cost(controls)` and `cost_and_grad(controls)` take controls in cost-function format
    ((M x K) float64 or complex128) and return what the reference seam returns."""

    def __init__(self, hamiltonian, initial_states, costs, evolution_time, system_eval_count,
                 control_eval_count=0, control_count=0, complex_controls=False,
                 magnus_policy=MagnusPolicy.M2, cost_eval_step=1,
                 interpolation_policy=InterpolationPolicy.LINEAR, device=0, store_tape=True,
                 chunks_per_member=0, ensemble_drifts=None, structure=None, slice_range=None, state_slice=None):
        if interpolation_policy != InterpolationPolicy.LINEAR:
            raise NotImplementedError("The interpolation policy {} is not yet supported for this method."
                                      "".format(interpolation_policy))
        if not isinstance(magnus_policy, MagnusPolicy):
            raise ValueError("Unrecognized magnus policy {}.".format(magnus_policy))
        self.lib = _lib.load()
        self._q = magnus_policy.order // 2
        initial_states = np.asarray(initial_states)
        # state sharding (qoc_b200/core/sharded.py): this plan holds states [s0, s1) of S_total; normalisations use S_total
        self.S_total = initial_states.shape[0]
        self.state_first = 0
        if state_slice is not None:
            s0, s1 = int(state_slice[0]), int(state_slice[1])
            if not 0 <= s0 < s1 <= self.S_total:
                raise ValueError("bad state slice {} of {} states".format(state_slice, self.S_total))
            if ensemble_drifts is not None or slice_range is not None:
                raise NotImplementedError("state sharding cannot be combined with ensembles or time-slice sharding")
            initial_states = initial_states[s0:s1]
            self.state_first = s0
        self.S, self.n = initial_states.shape[0], initial_states.shape[1]
        self.K = int(control_count)
        self.M = int(control_eval_count)
        self.N = int(system_eval_count)
        self.complex_controls = bool(complex_controls)
        self.KR = self.K * (2 if self.complex_controls else 1)
        self.costs = list(costs)
        if structure is None:
            structure = extract_hamiltonian_structure(hamiltonian, self.K, self.complex_controls, evolution_time,
                                                      system_eval_count=self.N, magnus_order=magnus_policy.order)
        self.structure = structure
        h0, a_ops = structure[0], structure[1]
        node_map = structure[2:] if len(structure) == 4 else None
        if node_map is not None and ensemble_drifts is not None:
            raise NotImplementedError("ensemble drifts with a time-dependent hamiltonian are not supported")
        if h0.shape[0] != self.n:
            raise ValueError("hamiltonian size {} does not match the states' hilbert size {}".format(h0.shape[0], self.n))
        if ensemble_drifts is not None:          # build-side extension: members differ in the drift only
            h0s = np.ascontiguousarray(ensemble_drifts, dtype=np.complex128)
        else:
            h0s = np.ascontiguousarray(h0[None], dtype=np.complex128)
        self.E = h0s.shape[0]
        pb = _lib.Problem(hilbert_size=self.n, state_count=self.S, control_count=self.KR,
                          control_eval_count=self.M, system_eval_count=self.N,
                          magnus_order=magnus_policy.order, cost_eval_step=int(cost_eval_step),
                          ensemble_count=self.E, device=int(device), store_tape=int(bool(store_tape)),
                          chunks_per_member=int(chunks_per_member),
                          slice_begin=0 if slice_range is None else int(slice_range[0]),
                          slice_end=0 if slice_range is None else int(slice_range[1]),
                          channel_count=0 if node_map is None else int(a_ops.shape[0]),
                          state_total=self.S_total if state_slice is not None else 0, state_first=self.state_first,
                          evolution_time=float(evolution_time))
        handle = ctypes.c_void_p()
        _lib.check(self.lib.qocb_plan_create(ctypes.byref(pb), ctypes.byref(handle)))
        self.handle = handle
        a_ops = np.ascontiguousarray(a_ops, dtype=np.complex128)
        _lib.check(self.lib.qocb_set_operators(handle, _lib.ptr(h0s), _lib.ptr(a_ops) if a_ops.shape[0] else None), handle)
        if node_map is not None:
            offset = np.ascontiguousarray(node_map[0], dtype=np.float64)
            gain = np.ascontiguousarray(node_map[1], dtype=np.float64)
            _lib.check(self.lib.qocb_set_node_map(handle, _lib.ptr(offset), _lib.ptr(gain) if self.KR else None), handle)
        psi0 = np.ascontiguousarray(initial_states.reshape(self.S, self.n), dtype=np.complex128)
        _lib.check(self.lib.qocb_set_states(handle, _lib.ptr(psi0)), handle)
        self.control_costs = []
        self.user_costs = []
        from qoc_b200.models.cost import Cost
        for c in self.costs:
            if (getattr(type(c), "device_terms", Cost.device_terms) is Cost.device_terms and
                    getattr(type(c), "control_value_and_grad", Cost.control_value_and_grad) is Cost.control_value_and_grad):
                # a user-defined Cost subclass (qoc/models/cost.py:5-51): only `cost(controls, states, step)` exists.  Its value and
                # a numeric cotangent of the final states are computed on the host around qocb_forward / qocb_backward.
                if getattr(c, "requires_step_evaluation", False):
                    raise NotImplementedError("The user-defined cost {} requires step evaluation; only final-step user costs are "
                                              "supported by the CUDA path (no CPU fallback).".format(c))
                if ensemble_drifts is not None or slice_range is not None or state_slice is not None:
                    raise NotImplementedError("user-defined costs cannot be combined with ensembles or sharded plans")
                self.user_costs.append(c)
                continue
            terms = c.device_terms(self.S_total, self.n)
            if state_slice is not None:                  # the vectors (and counts) of the local states only
                terms = [(kind, step, weight, np.asarray(vecs)[self.state_first:self.state_first + self.S],
                          None if counts is None else np.asarray(counts)[self.state_first:self.state_first + self.S])
                         for kind, step, weight, vecs, counts in terms]
            if not terms:
                if getattr(type(c), "control_value_and_grad", Cost.control_value_and_grad) is Cost.control_value_and_grad:
                    raise NotImplementedError("The cost {} has neither device terms nor an analytic control gradient "
                                              "(no CPU fallback).".format(c))
                self.control_costs.append(c)
            for kind, step, weight, vecs, counts in terms:
                vecs = np.ascontiguousarray(vecs, dtype=np.complex128)
                cnt = None if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
                _lib.check(self.lib.qocb_add_cost(handle, kind, step, float(weight), _lib.ptr(vecs),
                                                  None if cnt is None else cnt.ctypes.data_as(ctypes.c_void_p),
                                                  vecs.shape[1]), handle)

    # -- helpers ------------------------------------------------------------------------------------
    def _real_channels(self, controls):
        if self.KR == 0:
            return None
        controls = np.asarray(controls)
        if controls.shape != (self.M, self.K):
            raise ValueError("controls must have shape {}".format((self.M, self.K)))
        if self.complex_controls:
            return np.ascontiguousarray(np.concatenate([controls.real, controls.imag], axis=1), dtype=np.float64)
        return np.ascontiguousarray(controls, dtype=np.float64)

    def _control_costs(self, controls, want_grad):
        value, grad = 0.0, None
        for c in self.control_costs:
            v, g = c.control_value_and_grad(controls)
            value += v
            if want_grad and g is not None:
                grad = g if grad is None else grad + g
        return value, grad

    def _final_states(self, buf):
        fs = buf.reshape(self.E, self.S, self.n, 1)
        return fs[0] if self.E == 1 else fs

    # -- user-defined costs: value and cotangents by 4-point central differences of `cost(controls, states, step)` ---------
    def _user_value(self, controls, finals):
        return float(sum(np.real(c.cost(controls, finals, self.N - 1)) for c in self.user_costs))

    def _user_seed(self, controls, finals):
        """d c / d Re psi - i d c / d Im psi of the summed user costs at the final states (autograd's cotangent convention)."""
        seed = np.zeros((self.S, self.n), dtype=np.complex128)
        psi = np.array(finals, dtype=np.complex128)
        for s_ in range(self.S):
            for a in range(self.n):
                h = 1e-3 * max(1.0, abs(psi[s_, a, 0]))
                d = []
                for direction in (1.0, 1j):
                    vals = []
                    for k in (1.0, -1.0, 2.0, -2.0):
                        q = psi.copy()
                        q[s_, a, 0] += k * h * direction
                        vals.append(self._user_value(controls, q))
                    d.append((8.0 * (vals[0] - vals[1]) - (vals[2] - vals[3])) / (12.0 * h))
                seed[s_, a] = d[0] - 1j * d[1]
        return seed

    def _user_control_grad(self, controls, finals):
        """explicit dependence of the user costs on the controls (most have none: one probe decides)."""
        controls = np.asarray(controls)
        base = self._user_value(controls, finals)
        rng = np.random.default_rng(0)
        probe = controls + 1e-3 * (rng.standard_normal(controls.shape) + (1j * rng.standard_normal(controls.shape)
                                                                            if np.iscomplexobj(controls) else 0))
        if abs(self._user_value(probe, finals) - base) <= 1e-14 * max(1.0, abs(base)):
            return None
        g = np.zeros(controls.shape, dtype=controls.dtype)
        for idx in np.ndindex(*controls.shape):
            h = 1e-3 * max(1.0, abs(controls[idx]))
            parts = []
            for direction in ((1.0, 1j) if np.iscomplexobj(controls) else (1.0,)):
                vals = []
                for k in (1.0, -1.0, 2.0, -2.0):
                    q = controls.copy()
                    q[idx] += k * h * direction
                    vals.append(self._user_value(q, finals))
                parts.append((8.0 * (vals[0] - vals[1]) - (vals[2] - vals[3])) / (12.0 * h))
            g[idx] = parts[0] + (1j * parts[1] if len(parts) > 1 else 0)
        return g

    def _eval_with_user_costs(self, controls, want_grad):
        x = self._real_channels(controls)
        out = np.zeros(1)
        fs = np.empty((self.E, self.S, self.n), dtype=np.complex128)
        _lib.check(self.lib.qocb_forward(self.handle, _lib.ptr(x), _lib.ptr(out), _lib.ptr(fs)), self.handle)
        finals = self._final_states(fs)
        err = float(out[0]) + self._user_value(controls, finals)
        extra, extra_grad = self._control_costs(np.asarray(controls), want_grad)
        if not want_grad:
            return err + extra, None, finals
        seed = np.ascontiguousarray(self._user_seed(controls, finals))
        g = np.zeros((self.M, self.KR))
        _lib.check(self.lib.qocb_backward(self.handle, _lib.ptr(seed), _lib.ptr(g)), self.handle)
        grads = g[:, :self.K] + 1j * g[:, self.K:] if self.complex_controls else g
        ug = self._user_control_grad(controls, finals)
        if ug is not None:
            grads = grads + ug
        if extra_grad is not None:
            grads = grads + extra_grad
        return err + extra, grads, finals

    # -- the seam ---------------------------------------------------------------------------------------
    def cost(self, controls):
        if self.user_costs:
            err, _, finals = self._eval_with_user_costs(controls, False)
            return err, finals
        x = self._real_channels(controls)
        out = np.zeros(1)
        fs = np.empty((self.E, self.S, self.n), dtype=np.complex128)
        _lib.check(self.lib.qocb_cost(self.handle, _lib.ptr(x), _lib.ptr(out), _lib.ptr(fs)), self.handle)
        extra, _ = self._control_costs(controls, False) if controls is not None else (0.0, None)
        return float(out[0]) + extra, self._final_states(fs)

    def cost_and_grad(self, controls):
        """returns (error, grads, final_states); grads has controls' shape and dtype and is
        dE/dRe(u) + i dE/dIm(u) for complex controls (the value after the wrapper's conjugate,
        schroedingerdiscrete.py:323-324)."""
        if self.user_costs:
            return self._eval_with_user_costs(controls, True)
        x = self._real_channels(controls)
        out = np.zeros(1)
        g = np.zeros((self.M, self.KR))
        fs = np.empty((self.E, self.S, self.n), dtype=np.complex128)
        _lib.check(self.lib.qocb_cost_and_grad(self.handle, _lib.ptr(x), _lib.ptr(out), _lib.ptr(g), _lib.ptr(fs)),
                   self.handle)
        grads = g[:, :self.K] + 1j * g[:, self.K:] if self.complex_controls else g
        extra, extra_grad = self._control_costs(np.asarray(controls), True)
        if extra_grad is not None:
            grads = grads + extra_grad
        return float(out[0]) + extra, grads, self._final_states(fs)

    def cost_and_grad_autograd(self, controls):
        """the same (error, grads, finals) through HIPS autograd: the GPU evaluation is registered as an autograd primitive
        with defvjp (`make_autograd_primitive`) and differentiated with the reference's own operator
        (`ans_jacobian`, qoc/core/schroedingerdiscrete.py:318, followed by the conjugate for complex controls, :323-324).
        Used by the `_value_and_jacobian_*` seams whenever `autograd` is importable."""
        from qoc_b200.standard.utils import ans_jacobian, make_autograd_primitive
        prim = getattr(self, "_autograd_primitive", None)
        if prim is None:
            def value_and_grad(c):
                err, g, fin = self.cost_and_grad(c)
                self._autograd_finals = fin
                return err, (np.conjugate(g) if np.iscomplexobj(g) else g)      # autograd convention: d/dx - i d/dy
            prim = self._autograd_primitive = make_autograd_primitive(value_and_grad)
        controls = np.asarray(controls)
        error, jac = ans_jacobian(prim, 0)(controls)
        grads = np.conjugate(jac) if np.iscomplexobj(controls) else jac
        return float(error), grads, self._autograd_finals

    def final_states(self):
        """final states of the last evaluation (also of one enqueued through the device-resident hooks)."""
        fs = np.empty((self.E, self.S, self.n), dtype=np.complex128)
        _lib.check(self.lib.qocb_get_final_states(self.handle, _lib.ptr(fs)), self.handle)
        return self._final_states(fs)

    def node_grad(self):
        """per-node cotangents of the operator-channel coefficients of the last cost_and_grad: [N-1][q][KC] (member 0)."""
        KC = self.structure[1].shape[0]
        buf = np.zeros((self.N - 1, self._q, KC))
        _lib.check(self.lib.qocb_get_node_grad(self.handle, _lib.ptr(buf)), self.handle)
        return buf

    def intermediate_states(self):
        """[N][S][n][1] states of the last evaluation (member 0 unless E > 1: then [E][N][S][n][1])."""
        buf = np.empty((self.E, self.N, self.S, self.n), dtype=np.complex128)
        _lib.check(self.lib.qocb_get_states(self.handle, _lib.ptr(buf)), self.handle)
        buf = buf[..., None]
        return buf[0] if self.E == 1 else buf

    def propagators(self):
        buf = np.empty((self.E, self.N - 1, self.n, self.n), dtype=np.complex128)
        _lib.check(self.lib.qocb_get_propagators(self.handle, _lib.ptr(buf)), self.handle)
        return buf[0] if self.E == 1 else buf

    # -- device-resident benchmarking hooks --------------------------------------------------------------
    def upload(self, controls):
        _lib.check(self.lib.qocb_upload_controls(self.handle, _lib.ptr(self._real_channels(controls))), self.handle)

    def time_resident(self, with_grad=True, warmup=3, iters=10, flush_l2=True):
        total = np.zeros(1)
        stages = np.zeros(8)
        _lib.check(self.lib.qocb_time_resident(self.handle, int(with_grad), warmup, iters, int(flush_l2),
                                               _lib.ptr(total), _lib.ptr(stages)), self.handle)
        return float(total[0]), stages

    def launch_count(self, with_grad=True):
        return int(self.lib.qocb_launch_count(self.handle, int(with_grad)))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.qocb_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _interp_pair(x, xs):
    """the two control points and weights the reference's linear interpolation uses at time x, including its
    extrapolation rule outside [xs[0], xs[-1]] (qoc/core/mathmethods.py:54-65)."""
    M = len(xs)
    if x <= xs[0]:
        i0, i1 = 0, 1
    elif x >= xs[M - 1]:
        i0, i1 = M - 2, M - 1
    else:
        i1 = int(np.argmax(x <= xs))
        i0 = i1 - 1
    w1 = (x - xs[i0]) / (xs[i1] - xs[i0])
    return i0, i1, 1.0 - w1, w1


class NonlinearSchroedingerPlan(object):
    """`hamiltonian(controls, time)` callables that are NOT affine in the controls, e.g.
    H0 + u0^2 X + sin(u1 t) Y.  The reference differentiates through the callable with autograd; here

      * once per plan, the callable is sampled at random controls and node times and H - G0 is expanded over a small real
        operator basis {A_c}, c < 16 (anything that needs more channels raises NotImplementedError);
      * per evaluation the HOST calls the callable at every Magnus node of every slice - exactly where the reference calls it
        (qoc/core/schroedingerdiscrete.py:483-497) - with the linearly interpolated controls, projects H - G0 on the basis
        (residual checked) and uploads the channel coefficients; the same CUDA kernels propagate and differentiate with
        respect to those coefficients (`qocb_set_node_map` with control_count = 0, `qocb_get_node_grad`);
      * the chain rule through the callable uses a NUMERIC Jacobian of the coefficients with respect to (Re u, Im u) at each
        node: 4-point central differences (error O(h^4)), about 1e-11 relative for smooth callables.  Cost and final states
        keep the 1e-10 parity of the affine path; the gradient is as accurate as that Jacobian (tests assert 1e-8).

    The host work is O(nodes * control channels) Python calls per evaluation: a generality path, not a fast one."""

    def __init__(self, hamiltonian, initial_states, costs, evolution_time, system_eval_count, control_eval_count=0,
                 control_count=0, complex_controls=False, magnus_policy=MagnusPolicy.M2, cost_eval_step=1,
                 interpolation_policy=InterpolationPolicy.LINEAR, device=0, store_tape=True, seed=1234, **unused):
        from qoc_b200.models.cost import Cost
        self.hamiltonian = hamiltonian
        self.K, self.M, self.N = int(control_count), int(control_eval_count), int(system_eval_count)
        self.complex_controls = bool(complex_controls)
        self.KR = self.K * (2 if self.complex_controls else 1)
        self.T = float(evolution_time)
        self.q = magnus_policy.order // 2
        nsl = self.N - 1
        dt = self.T / nsl
        self.times = np.array([[(j + c) * dt for c in _NODES[self.q]] for j in range(nsl)])         # [N-1][q]
        xs = np.linspace(0, self.T, self.M)
        self.pairs = [[_interp_pair(t, xs) for t in row] for row in self.times]
        dtype = np.complex128 if self.complex_controls else np.float64
        rng = np.random.default_rng(seed)
        zero = np.zeros(self.K, dtype=dtype)
        self.g0 = np.array(hamiltonian(zero, float(self.times[0, 0])), dtype=np.complex128)
        n = self.g0.shape[0]
        basis = _RealBasis(_MAX_CHANNELS)
        probe_times = list(self.times.ravel()[np.linspace(0, self.times.size - 1, min(12, self.times.size)).astype(int)])
        for t in probe_times:
            basis.add(np.asarray(hamiltonian(zero, float(t)), dtype=np.complex128) - self.g0)
            for trial in range(10):
                u = rng.standard_normal(self.K) * (0.3 + 0.5 * trial)
                if self.complex_controls:
                    u = u + 1j * rng.standard_normal(self.K) * (0.3 + 0.5 * trial)
                basis.add(np.asarray(hamiltonian(u.astype(dtype), float(t)), dtype=np.complex128) - self.g0)
        self.KC = len(basis.vecs)
        if self.KC == 0:
            raise ValueError("the hamiltonian does not depend on controls or time: use a SchroedingerPlan")
        self.basis = np.array(basis.vecs)                                    # [KC][2 n^2] orthonormal over the reals
        self.scale = max(1.0, basis.scale)
        channels = basis.matrices(self.g0.shape)
        offset = np.zeros((nsl, self.q, self.KC))
        gain = np.zeros((nsl, self.q, self.KC, 0))
        for c in costs:
            if (getattr(type(c), "device_terms", Cost.device_terms) is Cost.device_terms and
                    getattr(type(c), "control_value_and_grad", Cost.control_value_and_grad) is Cost.control_value_and_grad):
                raise NotImplementedError("user-defined costs together with a non-linear hamiltonian are not supported")
        self.state_costs = [c for c in costs if getattr(type(c), "control_value_and_grad", Cost.control_value_and_grad)
                            is Cost.control_value_and_grad]
        self.control_costs = [c for c in costs if c not in self.state_costs]
        self.inner = SchroedingerPlan(None, initial_states, self.state_costs, evolution_time, system_eval_count,
                                      control_eval_count=0, control_count=0, complex_controls=False, magnus_policy=magnus_policy,
                                      cost_eval_step=cost_eval_step, interpolation_policy=interpolation_policy, device=device,
                                      store_tape=store_tape, structure=(self.g0, channels, offset, gain))
        self.S, self.n, self.E = self.inner.S, n, 1

    # -- host side of one evaluation ---------------------------------------------------------------------
    def _project(self, h):
        d = h - self.g0
        v = np.concatenate([d.real.ravel(), d.imag.ravel()])
        c = self.basis @ v
        if np.abs(v - c @ self.basis).max() > 1e-9 * self.scale:
            raise RuntimeError("hamiltonian(controls, time) left the operator basis found at plan creation (a term that the "
                               "probing did not excite): not representable on the CUDA path")
        return c

    def _coefficients(self, controls, want_jacobian):
        controls = np.asarray(controls)
        if controls.shape != (self.M, self.K):
            raise ValueError("controls must have shape {}".format((self.M, self.K)))
        nsl = self.N - 1
        coef = np.zeros((nsl, self.q, self.KC))
        jac = np.zeros((nsl, self.q, self.KC, self.KR)) if want_jacobian else None
        dirs = [1.0] * self.K + ([1j] * self.K if self.complex_controls else [])
        for j in range(nsl):
            for i in range(self.q):
                i0, i1, w0, w1 = self.pairs[j][i]
                u = controls[i0] * w0 + controls[i1] * w1
                t = float(self.times[j, i])
                coef[j, i] = self._project(np.asarray(self.hamiltonian(u, t), dtype=np.complex128))
                if not want_jacobian:
                    continue
                for r in range(self.KR):
                    k = r % self.K
                    h = 1e-3 * max(1.0, abs(u[k]))
                    e = np.zeros(self.K, dtype=u.dtype)
                    e[k] = dirs[r]

                    def f(s_, e=e, h=h):
                        return self._project(np.asarray(self.hamiltonian(u + s_ * h * e, t), dtype=np.complex128))
                    jac[j, i, :, r] = (8.0 * (f(1.0) - f(-1.0)) - (f(2.0) - f(-2.0))) / (12.0 * h)
        return coef, jac

    def _upload(self, coef):
        p = self.inner
        off = np.ascontiguousarray(coef, dtype=np.float64)
        _lib.check(p.lib.qocb_set_node_map(p.handle, _lib.ptr(off), None), p.handle)

    _control_costs = SchroedingerPlan._control_costs

    def cost(self, controls):
        coef, _ = self._coefficients(controls, False)
        self._upload(coef)
        err, finals = self.inner.cost(None)
        return err + self._control_costs(np.asarray(controls), False)[0], finals

    def cost_and_grad(self, controls):
        """(error, grads, final_states) as `SchroedingerPlan.cost_and_grad`."""
        controls = np.asarray(controls)
        coef, jac = self._coefficients(controls, True)
        self._upload(coef)
        err, _, finals = self.inner.cost_and_grad(None)
        gc = self.inner.node_grad()                                          # dE / d coef [N-1][q][KC]
        gx = np.zeros((self.M, self.KR))
        for j in range(self.N - 1):
            for i in range(self.q):
                i0, i1, w0, w1 = self.pairs[j][i]
                g = jac[j, i].T @ gc[j, i]
                gx[i0] += w0 * g
                gx[i1] += w1 * g
        grads = gx[:, :self.K] + 1j * gx[:, self.K:] if self.complex_controls else gx
        extra, extra_grad = self._control_costs(controls, True)
        if extra_grad is not None:
            grads = grads + extra_grad
        return err + extra, grads, finals

    cost_and_grad_autograd = SchroedingerPlan.cost_and_grad_autograd

    def intermediate_states(self):
        return self.inner.intermediate_states()

    def launch_count(self, with_grad=True):
        return self.inner.launch_count(with_grad)

    def close(self):
        self.inner.close()


def make_schroedinger_plan(hamiltonian, *args, **kw):
    """`SchroedingerPlan` when the callable is affine in the controls (time-dependent or not), `NonlinearSchroedingerPlan`
    otherwise - what the public programs (`grape_schroedinger_discrete`, `evolve_schroedinger_discrete`) build."""
    try:
        return SchroedingerPlan(hamiltonian, *args, **kw)
    except NonlinearHamiltonian:
        return NonlinearSchroedingerPlan(hamiltonian, *args, **kw)


# --- Lindblad --------------------------------------------------------------------------------------------
def extract_lindblad_structure(lindblad_data, evolution_time):
    """(gammas [L], operators [L x n x n]) of a time-independent `lindblad_data(time)` callable
    (qoc/core/lindbladdiscrete.py:486-492); time-dependent dissipation raises NotImplementedError."""
    if lindblad_data is None:
        return None, None
    g0, o0 = lindblad_data(0.0)
    g1, o1 = lindblad_data(0.37 * evolution_time)
    g0, o0 = np.asarray(g0, dtype=np.float64), np.asarray(o0, dtype=np.complex128)
    if not (np.allclose(g0, np.asarray(g1), rtol=_PROBE_RTOL, atol=0) and
            np.allclose(o0, np.asarray(o1), rtol=_PROBE_RTOL, atol=_PROBE_RTOL)):
        raise NotImplementedError("time-dependent lindblad_data callables are not supported by the CUDA path yet "
                                  "(no CPU fallback)")
    if o0.ndim != 3 or g0.shape[0] != o0.shape[0]:
        raise ValueError("lindblad_data(time) must return (dissipators [L], operators [L x n x n])")
    return np.ascontiguousarray(g0), np.ascontiguousarray(o0)


class LindbladPlan(object):
    """Device plan for `_evaluate_lindblad_discrete` and its jacobian (qoc/core/lindbladdiscrete.py:357-441, :322).

    Gradient semantics: the discrete adjoint of the Dormand-Prince map on the realised (accepted) step grid.  The
    reference's autograd tape also differentiates the step-size controller; those terms are rounding noise (see
    oracle/lindblad_adjoint_model.py and DESIGN.md section 5) and are deliberately not reproduced."""

    def __init__(self, initial_densities, costs, evolution_time, system_eval_count, hamiltonian=None,
                 lindblad_data=None, control_eval_count=0, control_count=0, complex_controls=False, cost_eval_step=1,
                 interpolation_policy=InterpolationPolicy.LINEAR, device=0, max_rk_steps=0, structure=None):
        if interpolation_policy != InterpolationPolicy.LINEAR:
            raise ValueError("The interpolation policy {} is not implemented for this method."
                             "".format(interpolation_policy))
        self.lib = _lib.load()
        rho0 = np.ascontiguousarray(initial_densities, dtype=np.complex128)
        self.D, self.n = rho0.shape[0], rho0.shape[1]
        self.K, self.M, self.N = int(control_count), int(control_eval_count), int(system_eval_count)
        self.complex_controls = bool(complex_controls)
        self.have_h = hamiltonian is not None
        self.KR = self.K * (2 if self.complex_controls else 1) if self.have_h else 0
        self.costs = list(costs)
        h0 = a_ops = None
        if self.have_h:
            if structure is None:
                structure = extract_hamiltonian_structure(hamiltonian, self.K, self.complex_controls, evolution_time)
            h0, a_ops = structure
            if h0.shape[0] != self.n:
                raise ValueError("hamiltonian size {} does not match the densities' hilbert size {}".format(h0.shape[0], self.n))
        gammas, lops = extract_lindblad_structure(lindblad_data, evolution_time)
        self.L = 0 if gammas is None else gammas.shape[0]
        pb = _lib.LindbladProblem(hilbert_size=self.n, density_count=self.D, control_count=self.KR,
                                  control_eval_count=self.M if self.KR else 0, system_eval_count=self.N,
                                  cost_eval_step=int(cost_eval_step), lindblad_count=self.L,
                                  have_hamiltonian=int(self.have_h), device=int(device), max_rk_steps=int(max_rk_steps),
                                  reserved0=0, reserved1=0, evolution_time=float(evolution_time))
        handle = ctypes.c_void_p()
        rc = self.lib.qocb_lindblad_create(ctypes.byref(pb), ctypes.byref(handle))
        if rc != 0:
            raise RuntimeError("qoc_b200 CUDA Lindblad path failed (rc={}): {}".format(
                rc, self.lib.qocb_lindblad_last_error(None).decode()))
        self.handle = handle
        h0c = None if h0 is None else np.ascontiguousarray(h0, dtype=np.complex128)
        aoc = None if a_ops is None or self.KR == 0 else np.ascontiguousarray(a_ops, dtype=np.complex128)
        self._check(self.lib.qocb_lindblad_set_operators(handle, _lib.ptr(h0c), _lib.ptr(aoc), _lib.ptr(gammas), _lib.ptr(lops)))
        self._check(self.lib.qocb_lindblad_set_densities(handle, _lib.ptr(rho0)))
        self.control_costs = []
        from qoc_b200.models.cost import Cost
        for c in self.costs:
            terms = c.device_terms_density(self.D, self.n) if hasattr(c, "device_terms_density") else None
            if not terms:
                # only genuine control-only costs (classes that override the analytic hook) may skip the device: a state
                # cost passed by mistake or a user-defined density cost must not be dropped silently
                hook = getattr(type(c), "control_value_and_grad", None)
                if terms is None or hook is None or hook is Cost.control_value_and_grad:
                    raise NotImplementedError("The cost {} has no B200 device descriptor for the Lindblad path (no CPU "
                                              "fallback): density costs need `device_terms_density`, control-only costs "
                                              "`control_value_and_grad`.".format(c))
                self.control_costs.append(c)
            for kind, step, weight, mats, counts in terms:
                mats = np.ascontiguousarray(mats, dtype=np.complex128)
                cnt = None if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
                self._check(self.lib.qocb_lindblad_add_cost(handle, kind, step, float(weight), _lib.ptr(mats),
                                                            None if cnt is None else cnt.ctypes.data_as(ctypes.c_void_p),
                                                            mats.shape[1]))

    def _check(self, rc):
        if rc != 0:
            msg = self.lib.qocb_lindblad_last_error(self.handle)
            raise RuntimeError("qoc_b200 CUDA Lindblad path failed (rc={}): {}".format(rc, msg.decode() if msg else "?"))

    _real_channels = SchroedingerPlan._real_channels
    _control_costs = SchroedingerPlan._control_costs
    cost_and_grad_autograd = SchroedingerPlan.cost_and_grad_autograd

    def cost(self, controls):
        x = self._real_channels(controls) if controls is not None else None
        out = np.zeros(1)
        fd = np.empty((self.D, self.n, self.n), dtype=np.complex128)
        self._check(self.lib.qocb_lindblad_cost(self.handle, _lib.ptr(x), _lib.ptr(out), _lib.ptr(fd)))
        extra, _ = self._control_costs(controls, False) if controls is not None else (0.0, None)
        return float(out[0]) + extra, fd

    def cost_and_grad(self, controls):
        """(error, grads, final_densities); grads = dE/dRe(u) + i dE/dIm(u) for complex controls."""
        x = self._real_channels(controls)
        out = np.zeros(1)
        fd = np.empty((self.D, self.n, self.n), dtype=np.complex128)
        controls = np.asarray(controls)
        if self.KR == 0:                                   # hamiltonian = None: the state costs do not see the controls
            self._check(self.lib.qocb_lindblad_cost(self.handle, None, _lib.ptr(out), _lib.ptr(fd)))
            grads = np.zeros(controls.shape, dtype=controls.dtype)
        else:
            g = np.zeros((self.M, self.KR))
            self._check(self.lib.qocb_lindblad_cost_and_grad(self.handle, _lib.ptr(x), _lib.ptr(out), _lib.ptr(g), _lib.ptr(fd)))
            grads = g[:, :self.K] + 1j * g[:, self.K:] if self.complex_controls else g
        extra, extra_grad = self._control_costs(controls, True)
        if extra_grad is not None:
            grads = grads + extra_grad
        return float(out[0]) + extra, grads, fd

    def stats(self):
        s = np.zeros(2, dtype=np.int64)
        self._check(self.lib.qocb_lindblad_stats(self.handle, s.ctypes.data_as(ctypes.c_void_p)))
        return {"attempts": int(s[0]), "accepted": int(s[1])}

    def intermediate_densities(self):
        buf = np.empty((self.N, self.D, self.n, self.n), dtype=np.complex128)
        self._check(self.lib.qocb_lindblad_get_densities(self.handle, _lib.ptr(buf)))
        return buf

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.qocb_lindblad_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
