"""Seeded synthetic problems shared by the parity tests and bench.py.

Pure NumPy; nothing here touches /root/reference or the oracle."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rand_herm(rng, n):
    x = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    return (x + x.conj().T) / 2


def one_norm(a):
    return np.abs(a).sum(axis=0).max()


def haar_columns(rng, n, count):
    """`count` columns of Haar unitaries: one frame when count <= n, otherwise several independent frames side by side
    (cfg5 has S = 64 states in a 32-dimensional space; the reference places no constraint on the state list)."""
    blocks = []
    while True:
        z = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
        q, r = np.linalg.qr(z)
        q = q * (np.diag(r) / np.abs(np.diag(r)))
        blocks.append(q[:, :min(n, count)])
        count -= n
        if count <= 0:
            break
    return blocks[0] if len(blocks) == 1 else np.concatenate(blocks, axis=1)


class Problem(object):
    """random-control problem(n, N, K, S, order): H0 hermitian with ||dt H0||_1 = h0_norm, drives with
    ||dt H_k||_1 = drive_norm, controls ~ N(0, 0.5^2) on M = N points, dt = 1."""

    def __init__(self, n, slices, K, S, order, complex_controls=False, F=0, seed=0, stiff=1.0,
                 h0_norm=1.0, drive_norm=0.2, M=None, cost_eval_step=1, step_target=False, neglect_phase=False):
        rng = np.random.default_rng(seed)
        self.n, self.N, self.K, self.S, self.order = n, slices + 1, K, S, order
        self.M = self.N if M is None else M
        self.T = float(slices)
        self.complex_controls = complex_controls
        self.cost_eval_step = cost_eval_step
        h0 = rand_herm(rng, n)
        self.h0 = h0 * (stiff * h0_norm / one_norm(h0))
        drives = []
        for _ in range(K):
            if complex_controls:
                c = np.triu(rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)), 1)
                c = c * (stiff * drive_norm / max(one_norm(c + c.conj().T), 1e-300))
            else:
                c = rand_herm(rng, n)
                c = c * (stiff * drive_norm / one_norm(c))
            drives.append(c)
        self.drives = np.array(drives).reshape(K, n, n)
        if complex_controls:
            self.controls = (rng.standard_normal((self.M, K)) + 1j * rng.standard_normal((self.M, K))) * (0.5 / np.sqrt(2))
        else:
            self.controls = rng.standard_normal((self.M, K)) * 0.5
        cols = haar_columns(rng, n, min(n, S * (1 + F)) if S <= n else S)
        self.initial_states = np.ascontiguousarray(haar_columns(rng, n, S).T)[:, :, None]
        self.target_states = np.ascontiguousarray(cols[:, :S].T)[:, :, None]
        self.F = F
        self.forbidden_states = None
        if F > 0:
            fb = []
            for s in range(S):
                v = haar_columns(rng, n, min(F, n))
                fb.append(np.ascontiguousarray(v.T)[:, :, None])
            self.forbidden_states = np.array(fb)
        self.step_target = step_target
        self.neglect_phase = neglect_phase

    def hamiltonian_numpy(self):
        h0, dr, cc = self.h0, self.drives, self.complex_controls

        def hamiltonian(controls, time):
            h = h0
            if controls is None:
                return h
            for k in range(dr.shape[0]):
                h = h + controls[k] * dr[k]
                if cc:
                    h = h + np.conjugate(controls[k]) * dr[k].conj().T
            return h
        return hamiltonian

    # time-dependent variant: rotating drives and a modulated drift,
    #   H(u, t) = H0 + 0.3 cos(w0 t) D + sum_k [u_k e^{i w_k t} C_k + h.c.]           (complex controls)
    #   H(u, t) = H0 + 0.3 cos(w0 t) D + sum_k u_k [cos(w_k t) A_k + sin(w_k t) B_k]  (real controls)
    def _td_parts(self):
        rng = np.random.default_rng(1000 + self.n)
        d = rand_herm(rng, self.n)
        d = d * (one_norm(self.h0) / one_norm(d))
        b = []
        for k in range(self.K):
            x = rand_herm(rng, self.n)
            b.append(x * (one_norm(self.drives[k] + self.drives[k].conj().T) / one_norm(x) / 2))
        w = 2 * np.pi * (0.05 + 0.03 * np.arange(self.K + 1)) / 1.0
        return d, np.array(b).reshape(self.K, self.n, self.n), w

    def hamiltonian_td_numpy(self):
        h0, dr, cc = self.h0, self.drives, self.complex_controls
        d, b, w = self._td_parts()

        def hamiltonian(controls, time):
            h = h0 + 0.3 * np.cos(w[0] * time) * d
            if controls is None:
                return h
            for k in range(dr.shape[0]):
                if cc:
                    ph = np.exp(1j * w[k + 1] * time)
                    h = h + controls[k] * ph * dr[k] + np.conjugate(controls[k] * ph) * dr[k].conj().T
                else:
                    h = h + controls[k] * (np.cos(w[k + 1] * time) * dr[k] + np.sin(w[k + 1] * time) * b[k])
            return h
        return hamiltonian

    def hamiltonian_td_torch(self):
        import math
        import torch
        cdt = torch.complex128
        h0 = torch.as_tensor(self.h0, dtype=cdt)
        dr = torch.as_tensor(self.drives, dtype=cdt)
        drd = dr.conj().transpose(-1, -2)
        d0, b0, w = self._td_parts()
        d, b = torch.as_tensor(d0, dtype=cdt), torch.as_tensor(b0, dtype=cdt)
        cc = self.complex_controls

        def hamiltonian(controls, time):
            time = float(time)
            h = h0 + 0.3 * math.cos(w[0] * time) * d
            if controls is None:
                return h
            for k in range(dr.shape[0]):
                if cc:
                    ph = complex(math.cos(w[k + 1] * time), math.sin(w[k + 1] * time))
                    h = h + controls[k] * ph * dr[k] + torch.conj(controls[k] * ph) * drd[k]
                else:
                    h = h + controls[k] * (math.cos(w[k + 1] * time) * dr[k] + math.sin(w[k + 1] * time) * b[k])
            return h
        return hamiltonian

    def costs(self, mod):
        """cost objects from `mod` (qoc_b200.standard or the oracle module - same constructor signatures)."""
        out = []
        if self.step_target:
            out.append(mod.TargetStateInfidelityTime(self.N, self.target_states,
                                                     neglect_relative_pahse=self.neglect_phase,
                                                     cost_eval_step=self.cost_eval_step))
        else:
            out.append(mod.TargetStateInfidelity(self.target_states, neglect_relative_pahse=self.neglect_phase))
        if self.F > 0:
            out.append(mod.ForbidStates(self.forbidden_states, self.N, cost_eval_step=self.cost_eval_step,
                                        cost_multiplier=0.7))
        return out


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_schroedinger_costs(d, mod):
    """the cost list of tests/golden/make_golden.py:schroedinger_case rebuilt from the stored fields with the
    cost classes of `mod` (product or oracle)."""
    N, M, K = int(d["N"]), int(d["M"]), int(d["K"])
    ces = int(d["cost_eval_step"])
    neglect = bool(d["neglect_phase"])
    mx = d["max_control_norms"]
    costs = []
    for name in [str(x) for x in d["cost_names"]]:
        if name == "TargetStateInfidelity":
            costs.append(mod.TargetStateInfidelity(d["target_states"], neglect_relative_pahse=neglect, cost_multiplier=0.9))
        elif name == "ForbidStates":
            costs.append(mod.ForbidStates(d["forbidden_states"], N, cost_eval_step=ces, cost_multiplier=0.35))
        elif name == "TargetStateInfidelityTime":
            costs.append(mod.TargetStateInfidelityTime(N, d["target_states"], neglect_relative_pahse=neglect,
                                                       cost_eval_step=ces, cost_multiplier=0.2))
        elif name == "ControlNorm":
            costs.append(mod.ControlNorm(K, M, cost_multiplier=0.11, max_control_norms=mx))
        elif name == "ControlVariation1":
            costs.append(mod.ControlVariation(K, M, cost_multiplier=0.07, max_control_norms=mx, order=1))
        elif name == "ControlVariation2":
            costs.append(mod.ControlVariation(K, M, cost_multiplier=0.05, max_control_norms=mx, order=2))
        else:
            raise KeyError(name)
    return costs


def numpy_hamiltonian(h0, drives, complex_controls):
    def hamiltonian(controls, time):
        h = h0
        if controls is None:
            return h
        for k in range(drives.shape[0]):
            h = h + controls[k] * drives[k]
            if complex_controls:
                h = h + np.conjugate(controls[k]) * drives[k].conj().T
        return h
    return hamiltonian
