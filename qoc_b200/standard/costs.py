"""
costs.py - the cost classes of qoc.standard.costs with the same constructor signatures and `.cost(...)`
values (NumPy, host), plus the two hooks the B200 path uses:

  * `device_terms(S, n)`          - state costs: descriptors for the CUDA warp-shuffle reductions
                                    (kind, step_cost, weight, vectors[S][F][n], counts[S]);
  * `control_value_and_grad(u)`   - control-only costs: value and analytic gradient dE/dx + i dE/dy
                                    (the convention the optimiser receives, schroedingerdiscrete.py:320-324).

References: qoc/standard/costs/{targetstateinfidelity,targetstateinfidelitytime,forbidstates,
targetdensityinfidelity,targetdensityinfidelitytime,forbiddensities,controlnorm,controlvariation,controlarea,
controlbandwidthmax}.py.
"""
import numpy as np

from qoc_b200.models.cost import Cost

KIND_TARGET_COHERENT, KIND_TARGET_INCOHERENT, KIND_FORBID = 0, 1, 2
LKIND_TARGET, LKIND_FORBID = 0, 1          # density costs (QOCB_LCOST_*)


def _dag(x):
    return np.conjugate(np.swapaxes(x, -1, -2))


def _abs2(z):
    return np.real(z * np.conjugate(z))


# --- state costs -------------------------------------------------------------------------------------
class TargetStateInfidelity(Cost):
    """1 - |sum_s <t_s|psi_s>|^2 / S^2, or with `neglect_relative_pahse` (sic, reference spelling)
    1 - sum_s |<t_s|psi_s>|^2 / S  (targetstateinfidelity.py:39-63)."""
    name = "target_state_infidelity"
    requires_step_evaluation = False

    def __init__(self, target_states, neglect_relative_pahse=False, cost_multiplier=1.):
        super().__init__(cost_multiplier=cost_multiplier)
        target_states = np.asarray(target_states)
        self.state_count = target_states.shape[0]
        self.target_states_dagger = _dag(target_states)
        self.neglect_relative_pahse = neglect_relative_pahse
        self._norm = 1.0

    def cost(self, controls, states, system_eval_step):
        ips = np.matmul(self.target_states_dagger, states)[:, 0, 0]
        if not self.neglect_relative_pahse:
            fidelity = _abs2(np.sum(ips)) / self.state_count ** 2
        else:
            fidelity = np.sum(_abs2(ips)) / self.state_count
        return (1 - fidelity) / self._norm * self.cost_multiplier

    def device_terms(self, state_count, hilbert_size):
        vecs = _dag(self.target_states_dagger).reshape(self.state_count, 1, hilbert_size)
        kind = KIND_TARGET_INCOHERENT if self.neglect_relative_pahse else KIND_TARGET_COHERENT
        return [(kind, int(self.requires_step_evaluation), self.cost_multiplier / self._norm,
                 np.ascontiguousarray(vecs, dtype=np.complex128), None)]


class TargetStateInfidelityTime(TargetStateInfidelity):
    """the same infidelity at every cost step, divided by cost_eval_count = (N-1)//cost_eval_step
    (targetstateinfidelitytime.py:33-73)."""
    name = "target_state_infidelity_time"
    requires_step_evaluation = True

    def __init__(self, system_eval_count, target_states, neglect_relative_pahse=False, cost_eval_step=1,
                 cost_multiplier=1.):
        super().__init__(np.stack(target_states), neglect_relative_pahse, cost_multiplier)
        self.cost_eval_count, _ = np.divmod(system_eval_count - 1, cost_eval_step)
        self._norm = self.cost_eval_count


class ForbidStates(Cost):
    """sum_s [sum_f |<f_sf|psi_s>|^2 / F_s] / (cost_eval_count * S)  (forbidstates.py:30-81)."""
    name = "forbid_states"
    requires_step_evaluation = True

    def __init__(self, forbidden_states, system_eval_count, cost_eval_step=1, cost_multiplier=1.):
        super().__init__(cost_multiplier=cost_multiplier)
        state_count = len(forbidden_states)
        count, _ = np.divmod(system_eval_count - 1, cost_eval_step)
        self.cost_normalization_constant = count * state_count
        self.forbidden_states_count = np.array([len(f) for f in forbidden_states])
        self.forbidden_states_dagger = [_dag(np.asarray(f)) for f in forbidden_states]

    def cost(self, controls, states, system_eval_step):
        total = 0
        for i, fdag in enumerate(self.forbidden_states_dagger):
            ips = np.matmul(fdag, states[i])[:, 0, 0]
            total = total + np.sum(_abs2(ips)) / self.forbidden_states_count[i]
        return total / self.cost_normalization_constant * self.cost_multiplier

    def device_terms(self, state_count, hilbert_size):
        fmax = int(self.forbidden_states_count.max())
        vecs = np.zeros((state_count, fmax, hilbert_size), dtype=np.complex128)
        for s, fdag in enumerate(self.forbidden_states_dagger):
            vecs[s, :fdag.shape[0]] = _dag(fdag)[:, :, 0]
        return [(KIND_FORBID, 1, self.cost_multiplier / self.cost_normalization_constant, vecs,
                 self.forbidden_states_count.astype(np.int32))]


# --- density costs (Lindblad path) -----------------------------------------------------------------
class TargetDensityInfidelity(Cost):
    """1 - sum_d |tr(T_d^dagger rho_d)| / (D n)  (targetdensityinfidelity.py:41-69)."""
    name = "target_density_infidelity"
    requires_step_evaluation = False

    def __init__(self, target_densities, cost_multiplier=1.):
        super().__init__(cost_multiplier=cost_multiplier)
        target_densities = np.asarray(target_densities)
        self.density_count = target_densities.shape[0]
        self.hilbert_size = target_densities.shape[1]
        self.target_densities_dagger = _dag(target_densities)
        self._norm = 1.0

    def cost(self, controls, densities, sytem_eval_step):
        prods = np.matmul(self.target_densities_dagger, densities)
        fid = sum(np.abs(np.trace(p)) for p in prods) / (self.density_count * self.hilbert_size)
        return (1 - fid) / self._norm * self.cost_multiplier

    def device_terms_density(self, density_count, hilbert_size):
        mats = _dag(self.target_densities_dagger).reshape(self.density_count, 1, hilbert_size, hilbert_size)
        return [(LKIND_TARGET, int(self.requires_step_evaluation), self.cost_multiplier / self._norm,
                 np.ascontiguousarray(mats, dtype=np.complex128), None)]


class TargetDensityInfidelityTime(TargetDensityInfidelity):
    """as above divided by cost_eval_count; evaluated on the final densities only because the reference sets
    requires_step_evaluation = False (targetdensityinfidelitytime.py:30)."""
    name = "target_density_infidelity_time"
    requires_step_evaluation = False

    def __init__(self, system_eval_count, target_densities, cost_eval_step=1, cost_multiplier=1.):
        super().__init__(np.stack(target_densities), cost_multiplier)
        self.cost_eval_count, _ = np.divmod(system_eval_count - 1, cost_eval_step)
        self._norm = self.cost_eval_count


class ForbidDensities(Cost):
    """sum_d [sum_f |tr(F_df^dagger rho_d) / n|^2 / F_d] / (cost_eval_count * D)  (forbiddensities.py:53-85)."""
    name = "forbid_densities"
    requires_step_evaluation = True

    def __init__(self, forbidden_densities, system_eval_count, cost_eval_step=1, cost_multiplier=1.):
        super().__init__(cost_multiplier=cost_multiplier)
        density_count = len(forbidden_densities)
        count, _ = np.divmod(system_eval_count - 1, cost_eval_step)
        self.cost_normalization_constant = count * density_count
        self.forbidden_densities_count = np.array([len(f) for f in forbidden_densities])
        self.forbidden_densities_dagger = [_dag(np.asarray(f)) for f in forbidden_densities]
        self.hilbert_size = self.forbidden_densities_dagger[0].shape[-1]

    def cost(self, controls, densities, system_eval_step):
        total = 0
        for i, fdag in enumerate(self.forbidden_densities_dagger):
            ips = np.trace(np.matmul(fdag, densities[i]), axis1=-2, axis2=-1) / self.hilbert_size
            total = total + np.sum(_abs2(ips)) / self.forbidden_densities_count[i]
        return total / self.cost_normalization_constant * self.cost_multiplier

    def device_terms_density(self, density_count, hilbert_size):
        fmax = int(self.forbidden_densities_count.max())
        mats = np.zeros((density_count, fmax, hilbert_size, hilbert_size), dtype=np.complex128)
        for d, fdag in enumerate(self.forbidden_densities_dagger):
            mats[d, :fdag.shape[0]] = _dag(fdag)
        return [(LKIND_FORBID, 1, self.cost_multiplier / self.cost_normalization_constant, mats,
                 self.forbidden_densities_count.astype(np.int32))]


# --- control-only costs: host value + analytic gradient -------------------------------------------------
class _ControlCost(Cost):
    requires_step_evaluation = False

    def device_terms(self, state_count, hilbert_size):
        return []

    def device_terms_density(self, density_count, hilbert_size):
        return []

    def cost(self, controls, states, system_eval_step):
        return self.control_value_and_grad(controls)[0]


class ControlNorm(_ControlCost):
    """sum |u w / max|^2 / (M K)  (controlnorm.py:48-73)."""
    name = "control_norm"

    def __init__(self, control_count, control_eval_count, control_weights=None, cost_multiplier=1.,
                 max_control_norms=None):
        super().__init__(cost_multiplier=cost_multiplier)
        self.control_weights = control_weights
        self.controls_size = control_eval_count * control_count
        self.max_control_norms = max_control_norms

    def control_value_and_grad(self, controls):
        scale = np.ones(controls.shape[1]) if self.max_control_norms is None else 1.0 / np.asarray(self.max_control_norms)
        scale = scale * (1.0 if self.control_weights is None else np.asarray(self.control_weights))
        c = controls * scale
        f = self.cost_multiplier / self.controls_size
        return np.sum(_abs2(c)) * f, 2 * f * c * scale


class ControlVariation(_ControlCost):
    """sum |diff^order(u / max)|^2 / (K (M - order) 2^order)  (controlvariation.py:47-75)."""
    name = "control_variation"

    def __init__(self, control_count, control_eval_count, cost_multiplier=1., max_control_norms=None, order=1):
        super().__init__(cost_multiplier=cost_multiplier)
        self.max_control_norms = max_control_norms
        self.diffs_size = control_count * (control_eval_count - order)
        self.order = order
        self.cost_normalization_constant = self.diffs_size * (2 ** self.order)

    def control_value_and_grad(self, controls):
        scale = np.ones(controls.shape[1]) if self.max_control_norms is None else 1.0 / np.asarray(self.max_control_norms)
        d = np.diff(controls * scale, axis=0, n=self.order)
        f = self.cost_multiplier / self.cost_normalization_constant
        g = 2 * f * d
        for _ in range(self.order):            # transpose of the forward-difference operator
            g = np.concatenate([-g[:1], g[:-1] - g[1:], g[-1:]], axis=0)
        return np.sum(_abs2(d)) * f, g * scale


class ControlArea(_ControlCost):
    """sum_k |sum_t u_k / max_k| / (M K)  (controlarea.py:43-67).  As in the reference, the branch without
    max_control_norms is not executable there (NameError, :55-64) and raises the same error here."""
    name = "control_area"

    def __init__(self, control_count, control_eval_count, cost_multiplier=1., max_control_norms=None):
        super().__init__(cost_multiplier=cost_multiplier)
        self.control_count = control_count
        self.control_size = control_count * control_eval_count
        self.max_control_norms = max_control_norms

    def control_value_and_grad(self, controls):
        if self.max_control_norms is None:
            raise NameError("name 'normalized_controls' is not defined")
        scale = 1.0 / np.asarray(self.max_control_norms)
        sums = np.sum(controls * scale, axis=0)
        mags = np.abs(sums)
        f = self.cost_multiplier / self.control_size
        direction = np.where(mags > 0, sums / np.where(mags > 0, mags, 1), 0)
        return np.sum(mags) * f, np.broadcast_to(f * direction * scale, controls.shape).copy()


class ControlBandwidthMax(_ControlCost):
    """FFT-band penalty (controlbandwidthmax.py:52-77): per control, sum of |fft| over freqs >= max_bandwidth,
    divided by (count * max of those magnitudes)."""
    name = "control_bandwidth_max"

    def __init__(self, control_count, control_eval_count, evolution_time, max_bandwidths, cost_multiplier=1.):
        super().__init__(cost_multiplier=cost_multiplier)
        self.max_bandwidths = max_bandwidths
        self.control_count = control_count
        dt = evolution_time / (control_eval_count - 1)
        self.freqs = np.fft.fftfreq(control_eval_count, d=dt)

    def control_value_and_grad(self, controls):
        M = controls.shape[0]
        total = 0.0
        grad = np.zeros(controls.shape, dtype=controls.dtype)
        f = self.cost_multiplier / self.control_count
        for i, bw in enumerate(self.max_bandwidths):
            spec = np.fft.fft(controls[:, i])
            mag = np.abs(spec)
            idx = np.nonzero(self.freqs >= bw)[0]
            pen = mag[idx]
            top = np.max(pen)
            total = total + np.sum(pen) / (idx.shape[0] * top)
            # d/d mag_i of  sum(pen) / (count * max(pen))
            coef = np.zeros(M)
            coef[idx] = 1.0 / (idx.shape[0] * top)
            coef[idx[np.argmax(pen)]] -= np.sum(pen) / (idx.shape[0] * top * top)
            phase = np.where(mag > 0, spec / np.where(mag > 0, mag, 1), 0)
            g = M * np.fft.ifft(coef * phase)                       # d/dx + i d/dy of sum coef_i |F_i|
            grad[:, i] = f * (g if np.iscomplexobj(controls) else np.real(g))
        return total * f, grad
