"""NumPy engines with the interface of `qoc_b200.core.sharded.units_evaluate` (forward / backward / pack), built on the
test-only adjoint model (oracle/adjoint_model.py): `NumpyMemberEngine` holds a block of ensemble members, `NumpyStateEngine`
a block of initial states (with the coherent-overlap exchange of TargetStateInfidelity).  They let the world_size-2 gloo
tests exercise the host-side protocol - partition, weights, the coherent all-reduce, payload layout - on CPU.
TEST INFRASTRUCTURE ONLY - never imported by the product."""
import numpy as np
import torch

from oracle import adjoint_model as am
from qoc_b200.core.sharded import slice_bounds


class NumpyMemberEngine(object):
    """members [b[rank], b[rank+1]) of an ensemble whose members differ in the drift; cost = mean over ALL members."""

    def __init__(self, rank, world, x, drifts, a_ops, psi0, terms, T, N, order, cost_eval_step=1):
        b = slice_bounds(len(drifts), world)
        self.mine = range(b[rank], b[rank + 1])
        self.weight = len(self.mine) / float(len(drifts))
        self.args = (x, drifts, a_ops, psi0, terms, T, N, order, cost_eval_step)

    def forward(self, with_grad):
        x, drifts, a_ops, psi0, terms, T, N, order, ces = self.args
        outs = [am.cost_and_grad(x, drifts[e], a_ops, psi0, terms, T, N, order, cost_eval_step=ces) for e in self.mine]
        self.cost = np.mean([o[0] for o in outs])                 # the device pipeline returns the mean over ITS members
        self.grad = np.mean([o[1] for o in outs], axis=0)
        return None

    def backward(self, coh):
        pass

    def pack(self, with_grad):
        g = self.grad if with_grad else np.zeros_like(self.grad)
        return torch.from_numpy(np.concatenate([g.ravel(), [self.cost]]) * self.weight)


class NumpyStateEngine(object):
    """states [b[rank], b[rank+1]) of the S initial states; every rank computes every slice propagator (they depend on the
    controls only) but sweeps its own states.  Normalisations use the TOTAL state count; the coherent target cost couples
    the states through T = sum_s <t_s|psi_s>: partial sums out of forward(), totals into backward()."""

    def __init__(self, rank, world, x, h0, a_ops, psi0, terms, T, N, order, cost_eval_step=1):
        self.rank = rank
        self.S_total = psi0.shape[0]
        b = slice_bounds(self.S_total, world)
        self.s0, self.s1 = b[rank], b[rank + 1]
        self.x, self.h0, self.a_ops, self.psi0, self.terms = np.asarray(x, dtype=float), h0, a_ops, psi0[self.s0:self.s1], terms
        self.T, self.N, self.order, self.ces = T, N, order, cost_eval_step
        self.M, self.KR = self.x.shape
        self.n = psi0.shape[1]
        self.idx, self.w = am.interp_table(T, self.M, N, order)
        self.dt = T / (N - 1)
        self.slots = [(k, ti) for k in range(1, N) for ti, t in enumerate(terms)
                      if t.kind == 0 and ((t.step and k % self.ces == 0) or (not t.step and k == N - 1))]

    def _active(self, t, k):
        return (t.step and k % self.ces == 0 and k != 0) or (not t.step and k == self.N - 1)

    def forward(self, with_grad):
        g0, g = -1j * self.h0, -1j * self.a_ops
        self.gens, self.tapes, self.us = [], [], []
        for j in range(self.N - 1):
            a = [g0 + np.tensordot(self.x[self.idx[j, i, 0]] * self.w[j, i, 0] + self.x[self.idx[j, i, 1]] * self.w[j, i, 1], g, axes=(0, 0))
                 for i in range(self.idx.shape[1])]
            u, tape = am.pade_fwd(am.magnus_fwd(a, self.dt, self.order))
            self.gens.append(a); self.tapes.append(tape); self.us.append(u)
        Sl = self.s1 - self.s0
        self.psi = np.zeros((self.N, Sl, self.n), dtype=complex)
        self.psi[0] = self.psi0
        for j in range(self.N - 1):
            self.psi[j + 1] = self.psi[j] @ self.us[j].T
        self.cost = 0.0
        self.seeds = np.zeros_like(self.psi)
        coh = np.zeros((len(self.slots), 2))
        for k in range(1, self.N):
            for ti, t in enumerate(self.terms):
                if not self._active(t, k):
                    continue
                S = self.S_total
                if t.kind == 0:
                    tot = sum(np.vdot(t.vectors[s][0], self.psi[k, s - self.s0]) for s in range(self.s0, self.s1))
                    coh[self.slots.index((k, ti))] = [tot.real, tot.imag]
                elif t.kind == 1:
                    if self.rank == 0:
                        self.cost += t.mult / t.norm                       # the constant 1 is counted once
                    for s in range(self.s0, self.s1):
                        ip = np.vdot(t.vectors[s][0], self.psi[k, s - self.s0])
                        self.cost -= t.mult / t.norm * abs(ip) ** 2 / S
                        self.seeds[k, s - self.s0] += -t.mult / t.norm * 2 * np.conj(ip) / S * np.conj(t.vectors[s][0])
                else:
                    for s in range(self.s0, self.s1):
                        F = t.vectors[s].shape[0]
                        for f in range(F):
                            ip = np.vdot(t.vectors[s][f], self.psi[k, s - self.s0])
                            self.cost += t.mult / t.norm * abs(ip) ** 2 / F
                            self.seeds[k, s - self.s0] += t.mult / t.norm * 2 * np.conj(ip) / F * np.conj(t.vectors[s][f])
        self.coh = torch.from_numpy(coh.ravel().copy())
        return self.coh

    def _coherent(self, coh):
        """value (rank 0 only) and seeds of the coherent terms from the all-reduced overlap sums"""
        tot = coh.numpy().reshape(-1, 2)
        S = self.S_total
        for (k, ti), (tr, tim) in zip(self.slots, tot):
            t, T_ = self.terms[ti], complex(tr, tim)
            if self.rank == 0:
                self.cost += t.mult / t.norm * (1 - abs(T_) ** 2 / S ** 2)
            for s in range(self.s0, self.s1):
                self.seeds[k, s - self.s0] += -t.mult / t.norm * 2 * np.conj(T_) / S ** 2 * np.conj(t.vectors[s][0])

    def backward(self, coh):
        self._coherent(coh)
        self._coherent_done = True
        lam = np.zeros_like(self.psi)
        lam[self.N - 1] = self.seeds[self.N - 1]
        for j in range(self.N - 2, -1, -1):
            lam[j] = lam[j + 1] @ self.us[j] + self.seeds[j]
        g = -1j * self.a_ops
        self.grad = np.zeros((self.M, self.KR))
        for j in range(self.N - 1):
            ubar = np.einsum("sa,sb->ab", lam[j + 1], self.psi[j])
            abar = am.magnus_bwd(self.gens[j], self.dt, self.order, am.pade_bwd(self.tapes[j], ubar))
            for i in range(self.idx.shape[1]):
                cbar = np.real(np.einsum("ab,rab->r", abar[i], g))
                self.grad[self.idx[j, i, 0]] += self.w[j, i, 0] * cbar
                self.grad[self.idx[j, i, 1]] += self.w[j, i, 1] * cbar

    def pack(self, with_grad):
        if not with_grad:                                              # forward only: the coherent value still needs the totals
            self._coherent(self.coh)
        g = self.grad if with_grad else np.zeros((self.M, self.KR))
        return torch.from_numpy(np.concatenate([g.ravel(), [self.cost]]))
