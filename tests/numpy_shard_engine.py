"""NumPy shard engine with the phase interface of `qoc_b200.core.sharded.CudaShardEngine`, built on the test-only
adjoint model (oracle/adjoint_model.py).  It lets the world_size-2 gloo tests exercise the host-side sharding
protocol (`sharded_evaluate`: partition, collective order, payload layouts, final reduction) on CPU.
TEST INFRASTRUCTURE ONLY - never imported by the product."""
import numpy as np
import torch

from oracle import adjoint_model as am
from qoc_b200.core.sharded import slice_bounds


def _planar(z):
    return np.concatenate([z.real.ravel(), z.imag.ravel()])


def _mat(v, n):
    return v[:n * n].reshape(n, n) + 1j * v[n * n:].reshape(n, n)


class NumpyShardEngine(object):
    def __init__(self, rank, world, x_shape, h0, a_ops, psi0, terms, T, N, order, cost_eval_step=1, node_map=None):
        """node_map = (offset [N-1, q, KC], gain [N-1, q, KC, KR]): a_ops are operator channels of a time-dependent
        hamiltonian with per-node coefficients offset + gain @ x(t) (qoc_b200/core/plan.py:extract_time_dependent_structure)"""
        self.node_map = node_map
        self.rank, self.world = rank, world
        self.h0, self.a_ops, self.psi0, self.terms = h0, a_ops, psi0, terms
        self.T, self.N, self.order, self.ces = T, N, order, cost_eval_step
        self.M, self.KR = x_shape
        self.S, self.n = psi0.shape
        b = slice_bounds(N - 1, world)
        self.j0, self.j1 = b[rank], b[rank + 1]
        self.owns_final = self.j1 == N - 1
        self.GM, self.VS = 2 * self.n * self.n, 2 * self.S * self.n
        self.RS = self.M * self.KR + 1 + self.VS
        self.idx, self.w = am.interp_table(T, self.M, N, order)
        self.dt = T / (N - 1)

    def upload(self, x):
        self.x = np.asarray(x, dtype=np.float64)

    def forward_local(self, with_grad):
        g0, g = -1j * self.h0, -1j * self.a_ops
        self.gens, self.tapes, self.us = [], [], []
        prop = np.eye(self.n, dtype=complex)
        for j in range(self.j0, self.j1):
            a = []
            for i in range(self.idx.shape[1]):
                c = self.x[self.idx[j, i, 0]] * self.w[j, i, 0] + self.x[self.idx[j, i, 1]] * self.w[j, i, 1]
                if self.node_map is not None:
                    c = self.node_map[0][j, i] + self.node_map[1][j, i] @ c
                a.append(g0 + np.tensordot(c, g, axes=(0, 0)))
            u, tape = am.pade_fwd(am.magnus_fwd(a, self.dt, self.order))
            self.gens.append(a); self.tapes.append(tape); self.us.append(u)
            prop = u @ prop
        return torch.from_numpy(_planar(prop))

    def _hits(self, k):
        step = k % self.ces == 0 and k != 0
        return step, k == self.N - 1

    def forward_finish(self, all_p):
        ap = all_p.numpy().reshape(self.world, self.GM)
        v = self.psi0.copy()
        for r in range(self.rank):
            v = v @ _mat(ap[r], self.n).T
        L = self.j1 - self.j0
        self.psi = np.zeros((L + 1, self.S, self.n), dtype=complex)
        self.psi[0] = v
        for j in range(L):
            self.psi[j + 1] = self.psi[j] @ self.us[j].T
        self.cost = 0.0
        self.seeds = np.zeros_like(self.psi)
        for k in range(self.j0, self.j1 + 1):
            st, fin = self._hits(k)
            for t in self.terms:
                if (t.step and st) or (not t.step and fin):
                    val, sd = t.value_and_seed(self.psi[k - self.j0])
                    self.seeds[k - self.j0] += sd
                    if k > self.j0:                       # values: states (j0, j1]; seeds: states [j0, j1) + final
                        self.cost += val

    def _costates(self, lam_in):
        L = self.j1 - self.j0
        lam = np.zeros_like(self.psi)
        lam[L] = lam_in + (self.seeds[L] if self.owns_final else 0)
        for j in range(L - 1, -1, -1):
            lam[j] = lam[j + 1] @ self.us[j] + self.seeds[j]
        return lam

    def backward_particular(self):
        return torch.from_numpy(_planar(self._costates(np.zeros((self.S, self.n), dtype=complex))[0]))

    def backward_finish(self, all_p, all_b):
        ap = all_p.numpy().reshape(self.world, self.GM)
        ab = all_b.numpy().reshape(self.world, self.VS)
        half = self.S * self.n
        v = np.zeros((self.S, self.n), dtype=complex)
        for r in range(self.world - 1, self.rank, -1):
            v = v @ _mat(ap[r], self.n) + (ab[r][:half] + 1j * ab[r][half:]).reshape(self.S, self.n)
        lam = self._costates(v)
        g = -1j * self.a_ops
        self.grad = np.zeros((self.M, self.KR))
        for jl in range(self.j1 - self.j0):
            j = self.j0 + jl
            ubar = np.einsum("sa,sb->ab", lam[jl + 1], self.psi[jl])
            abar = am.magnus_bwd(self.gens[jl], self.dt, self.order, am.pade_bwd(self.tapes[jl], ubar))
            for i in range(self.idx.shape[1]):
                cbar = np.real(np.einsum("ab,rab->r", abar[i], g))
                if self.node_map is not None:
                    cbar = self.node_map[1][j, i].T @ cbar
                self.grad[self.idx[j, i, 0]] += self.w[j, i, 0] * cbar
                self.grad[self.idx[j, i, 1]] += self.w[j, i, 1] * cbar

    def pack_result(self, with_grad):
        out = np.zeros(self.RS)
        if with_grad:
            out[:self.M * self.KR] = self.grad.ravel()
        out[self.M * self.KR] = self.cost
        if self.owns_final:
            fin = self.psi[-1]
            out[self.M * self.KR + 1:] = np.concatenate([np.stack([fin[s].real, fin[s].imag]).ravel() for s in range(self.S)])
        return torch.from_numpy(out)

    def unpack(self, host):
        cnt = self.M * self.KR
        fin = host[cnt + 1:].reshape(self.S, 2, self.n)
        return float(host[cnt]), host[:cnt].reshape(self.M, self.KR).copy(), (fin[:, 0] + 1j * fin[:, 1])[:, :, None]
