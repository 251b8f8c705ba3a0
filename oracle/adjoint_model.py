"""
adjoint_model.py - NumPy model of the *algorithm the CUDA kernels implement* for the Schroedinger GRAPE
cost+gradient, independent of torch.autograd.

*** TEST INFRASTRUCTURE ONLY (same rule as qoc_oracle.py). ***

Forward: per slice, interpolation table -> generator a_i = G0 + sum_r c_r G_r (G = -1j * H pieces) ->
Magnus M2/M4/M6 (qoc/core/mathmethods.py:72-164) -> Pade-13 scaling-and-squaring with an LU solve
(qoc/standard/functions/expm.py:210-252) -> state sweep and cost terms
(qoc/core/schroedingerdiscrete.py:356-438).
Backward: hand-written reverse mode over exactly that operation graph, in HIPS-autograd's cotangent
convention (no conjugations for holomorphic ops; real parameters take the real part), i.e. what
`ans_jacobian` (qoc/standard/utils/autogradutil.py:10-31) yields on the reference tape.
A second, independent expm adjoint via scipy.linalg.expm_frechet (`adjoint='frechet'`) cross-checks the
reverse-mode Pade formulas.

The control model is the real-linear structure the product extracts from the user's callable:
H(x) = H0 + sum_r x_r A_r with x = real controls (for complex controls x = [Re u, Im u]).
"""
import math

import numpy as np
import scipy.linalg as sla

B = (64764752532480000., 32382376266240000., 7771770303897600., 1187353796428800.,
     129060195264000., 10559470521600., 670442572800., 33522128640., 1323241920.,
     40840800., 960960., 16380., 182., 1.)
THETA13 = 5.371920351148152
S3, S15 = math.sqrt(3.0), math.sqrt(15.0)
NODES = {2: (0.5,), 4: (0.5 - S3 / 6, 0.5 + S3 / 6), 6: (0.5 - S15 / 10, 0.5, 0.5 + S15 / 10)}


def interp_table(T, M, N, order):
    """(i0, i1, w0, w1) per slice and Magnus node: value = y[i0]*w0 + y[i1]*w1, restating
    qoc/core/mathmethods.py:36-67 on control_eval_times = linspace(0, T, M) (programstate.py:41)."""
    xs = np.linspace(0, T, M)
    dt = T / (N - 1)
    q = len(NODES[order])
    idx = np.zeros((N - 1, q, 2), dtype=np.int32)
    w = np.zeros((N - 1, q, 2))
    for j in range(N - 1):
        for i, c in enumerate(NODES[order]):
            x = j * dt + dt * c
            if x <= xs[0]:
                i0, i1 = 0, 1
            elif x >= xs[-1]:
                i0, i1 = M - 2, M - 1
            else:
                i1 = int(np.argmax(x <= xs))
                i0 = i1 - 1
            w1 = (x - xs[i0]) / (xs[i1] - xs[i0])
            idx[j, i] = (i0, i1)
            w[j, i] = (1 - w1, w1)
    return idx, w


def comm(a, b):
    return a @ b - b @ a


def magnus_fwd(a, dt, order):
    if order == 2:
        return dt * a[0]
    if order == 4:
        return (dt / 2) * (a[0] + a[1]) + (S3 / 12) * dt * dt * comm(a[1], a[0])
    a1, a2, a3 = a
    b1 = dt * a2
    b2 = (S15 / 3) * dt * (a3 - a1)
    b3 = (10.0 / 3) * dt * (a3 - 2 * a2 + a1)
    c12 = comm(b1, b2)
    return b1 + 0.5 * b3 + (1.0 / 240) * comm(-20 * b1 - b3 + c12, b2 - (1.0 / 60) * comm(b1, 2 * b3 + c12))


def comm_bwd(a, b, cbar):
    """cotangents of C = a b - b a (unconjugated convention)."""
    return cbar @ b.T - b.T @ cbar, a.T @ cbar - cbar @ a.T


def magnus_bwd(a, dt, order, mbar):
    if order == 2:
        return [dt * mbar]
    if order == 4:
        f = (S3 / 12) * dt * dt
        a2bar, a1bar = comm_bwd(a[1], a[0], f * mbar)
        return [a1bar + (dt / 2) * mbar, a2bar + (dt / 2) * mbar]
    a1, a2, a3 = a
    b1 = dt * a2
    b2 = (S15 / 3) * dt * (a3 - a1)
    b3 = (10.0 / 3) * dt * (a3 - 2 * a2 + a1)
    c12 = comm(b1, b2)
    p = -20 * b1 - b3 + c12
    d = comm(b1, 2 * b3 + c12)
    qm = b2 - (1.0 / 60) * d
    b1bar = mbar.copy()
    b3bar = 0.5 * mbar
    pbar, qbar = comm_bwd(p, qm, (1.0 / 240) * mbar)
    b2bar = qbar.copy()
    dbar = -(1.0 / 60) * qbar
    t1, ebar = comm_bwd(b1, 2 * b3 + c12, dbar)
    b1bar = b1bar + t1 - 20 * pbar
    b3bar = b3bar + 2 * ebar - pbar
    c12bar = ebar + pbar
    t1, t2 = comm_bwd(b1, b2, c12bar)
    b1bar = b1bar + t1
    b2bar = b2bar + t2
    a1bar = -(S15 / 3) * dt * b2bar + (10.0 / 3) * dt * b3bar
    a2bar = dt * b1bar - (20.0 / 3) * dt * b3bar
    a3bar = (S15 / 3) * dt * b2bar + (10.0 / 3) * dt * b3bar
    return [a1bar, a2bar, a3bar]


def pade_fwd(m):
    """Pade-13 with scaling and squaring; returns U and the tape the backward needs."""
    norm = np.abs(m).sum(axis=0).max()
    s = 0
    if not norm < THETA13:
        s = max(0, int(math.ceil(math.log2(norm / THETA13))))
    a = m * (2.0 ** -s)
    n = a.shape[0]
    ident = np.eye(n)
    a2 = a @ a
    a4 = a2 @ a2
    a6 = a2 @ a4
    w1 = B[13] * a6 + B[11] * a4 + B[9] * a2
    x1 = B[12] * a6 + B[10] * a4 + B[8] * a2
    y = a6 @ w1 + B[7] * a6 + B[5] * a4 + B[3] * a2 + B[1] * ident
    uo = a @ y
    ve = a6 @ x1 + B[6] * a6 + B[4] * a4 + B[2] * a2 + B[0] * ident
    lu = sla.lu_factor(ve - uo)
    r = [sla.lu_solve(lu, ve + uo)]
    for _ in range(s):
        r.append(r[-1] @ r[-1])
    return r[-1], dict(s=s, a=a, a2=a2, a4=a4, a6=a6, w1=w1, x1=x1, y=y, lu=lu, r=r)


def pade_bwd(tape, ubar):
    """reverse mode through pade_fwd: returns the cotangent of the (unscaled) input matrix."""
    r = tape["r"]
    rbar = ubar
    for i in range(tape["s"], 0, -1):
        rbar = rbar @ r[i - 1].T + r[i - 1].T @ rbar
    pbar = sla.lu_solve(tape["lu"], rbar, trans=1)          # Q^{-T} rbar
    qbar = -pbar @ r[0].T
    uobar = pbar - qbar
    vebar = pbar + qbar
    a, a2, a4, a6, w1, x1, y = (tape[k] for k in ("a", "a2", "a4", "a6", "w1", "x1", "y"))
    abar = uobar @ y.T
    ybar = a.T @ uobar
    a6bar = B[7] * ybar + B[6] * vebar + ybar @ w1.T + vebar @ x1.T
    a4bar = B[5] * ybar + B[4] * vebar
    a2bar = B[3] * ybar + B[2] * vebar
    w1bar = a6.T @ ybar
    x1bar = a6.T @ vebar
    a6bar = a6bar + B[13] * w1bar + B[12] * x1bar
    a4bar = a4bar + B[11] * w1bar + B[10] * x1bar
    a2bar = a2bar + B[9] * w1bar + B[8] * x1bar
    a2bar = a2bar + a6bar @ a4.T
    a4bar = a4bar + a2.T @ a6bar
    a2bar = a2bar + a4bar @ a2.T + a2.T @ a4bar
    abar = abar + a2bar @ a.T + a.T @ a2bar
    return abar * (2.0 ** -tape["s"])


class CostTerm(object):
    """kind 0: mult/norm * (1 - |sum_s <v_s|psi_s>|^2 / S^2)   (targetstateinfidelity.py:52-56)
       kind 1: mult/norm * (1 - sum_s |<v_s|psi_s>|^2 / S)       (:57-61)
       kind 2: mult/norm * sum_s (1/F_s) sum_f |<v_sf|psi_s>|^2  (forbidstates.py:64-81)
       vectors: list over states of (F_s x n) arrays; step=True -> evaluated at every cost step."""
    def __init__(self, kind, vectors, mult, norm, step):
        self.kind, self.vectors, self.mult, self.norm, self.step = kind, vectors, mult, norm, step

    def value_and_seed(self, psi):
        S = psi.shape[0]
        seed = np.zeros_like(psi)
        if self.kind == 0:
            ips = np.array([np.vdot(self.vectors[s][0], psi[s]) for s in range(S)])
            tot = ips.sum()
            val = self.mult / self.norm * (1 - abs(tot) ** 2 / S ** 2)
            for s in range(S):
                seed[s] = -self.mult / self.norm * 2 * np.conj(tot) / S ** 2 * np.conj(self.vectors[s][0])
        elif self.kind == 1:
            val = 0.0
            for s in range(S):
                ip = np.vdot(self.vectors[s][0], psi[s])
                val += abs(ip) ** 2
                seed[s] = -self.mult / self.norm * 2 * np.conj(ip) / S * np.conj(self.vectors[s][0])
            val = self.mult / self.norm * (1 - val / S)
        else:
            val = 0.0
            for s in range(S):
                F = self.vectors[s].shape[0]
                for f in range(F):
                    ip = np.vdot(self.vectors[s][f], psi[s])
                    val += abs(ip) ** 2 / F
                    seed[s] += self.mult / self.norm * 2 * np.conj(ip) / F * np.conj(self.vectors[s][f])
            val = self.mult / self.norm * val
        return val, seed


def cost_and_grad(x, h0, a_ops, psi0, terms, T, N, order, cost_eval_step=1, adjoint="pade", chunks=1):
    """x: (M x KR) real controls; h0: (n x n); a_ops: (KR x n x n); psi0: (S x n).
    Returns (cost, grad (M x KR) real, final states (S x n)).  `chunks` > 1 exercises the chunked
    propagator scan (forward) and the affine costate recursion (backward) the multi-GPU path uses."""
    M, KR = x.shape
    dt = T / (N - 1)
    idx, w = interp_table(T, M, N, order)
    q = idx.shape[1]
    g0 = -1j * h0
    g = -1j * a_ops
    gens, tapes, us = [], [], []
    for j in range(N - 1):
        a = []
        for i in range(q):
            c = x[idx[j, i, 0]] * w[j, i, 0] + x[idx[j, i, 1]] * w[j, i, 1]
            a.append(g0 + np.tensordot(c, g, axes=(0, 0)))
        m = magnus_fwd(a, dt, order)
        u, tape = pade_fwd(m)
        gens.append(a)
        tapes.append(tape)
        us.append(u)
    S, n = psi0.shape
    psi = np.zeros((N, S, n), dtype=np.complex128)
    psi[0] = psi0
    if chunks > 1:
        bounds = np.linspace(0, N - 1, chunks + 1).astype(int)
        props = []
        for c in range(chunks):
            p = np.eye(n, dtype=np.complex128)
            for j in range(bounds[c], bounds[c + 1]):
                p = us[j] @ p
            props.append(p)
        for c in range(chunks):                     # sequential boundary states
            psi[bounds[c + 1]] = psi[bounds[c]] @ props[c].T
        for c in range(chunks):                     # independent local sweeps
            for j in range(bounds[c], bounds[c + 1] - 1):
                psi[j + 1] = psi[j] @ us[j].T
    else:
        for j in range(N - 1):
            psi[j + 1] = psi[j] @ us[j].T
    cost = 0.0
    seeds = np.zeros_like(psi)
    for step in range(N):
        for t in terms:
            hit = (t.step and step % cost_eval_step == 0 and step != 0) or (not t.step and step == N - 1)
            if hit:
                v, sd = t.value_and_seed(psi[step])
                cost += v
                seeds[step] += sd
    lam = np.zeros_like(psi)
    lam[N - 1] = seeds[N - 1]
    if chunks > 1:
        # affine recursion lam_j = lam_{j+1} U_j + seed_j: per-chunk particular part with zero incoming
        # costate (parallel), sequential combine over chunk boundaries, then parallel local sweeps.
        part = np.zeros((chunks, S, n), dtype=np.complex128)
        for c in range(chunks):
            l = np.zeros((S, n), dtype=np.complex128)
            for j in range(bounds[c + 1] - 1, bounds[c] - 1, -1):
                l = l @ us[j] + seeds[j]
            part[c] = l
        lam_b = [None] * (chunks + 1)
        lam_b[chunks] = seeds[N - 1].copy()
        for c in range(chunks - 1, -1, -1):
            lam_b[c] = lam_b[c + 1] @ props[c] + part[c]
        for c in range(chunks):
            lam[bounds[c + 1]] = lam_b[c + 1]
            for j in range(bounds[c + 1] - 1, bounds[c] - 1, -1):
                lam[j] = lam[j + 1] @ us[j] + seeds[j]
    else:
        for j in range(N - 2, -1, -1):
            lam[j] = lam[j + 1] @ us[j] + seeds[j]
    grad = np.zeros((M, KR))
    for j in range(N - 1):
        ubar = np.einsum("sa,sb->ab", lam[j + 1], psi[j])
        if adjoint == "pade":
            mbar = pade_bwd(tapes[j], ubar)
        else:
            m = magnus_fwd(gens[j], dt, order)
            mbar = sla.expm_frechet(m.T, ubar, compute_expm=False)
        abar = magnus_bwd(gens[j], dt, order, mbar)
        for i in range(q):
            cbar = np.real(np.einsum("ab,rab->r", abar[i], g))
            grad[idx[j, i, 0]] += w[j, i, 0] * cbar
            grad[idx[j, i, 1]] += w[j, i, 1] * cbar
    return cost, grad, psi[N - 1]
