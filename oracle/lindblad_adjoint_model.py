"""
lindblad_adjoint_model.py - NumPy model of the *algorithm the CUDA Lindblad kernels implement*
(SURVEY.md section 8 rows A12-A15): forward = the reference's per-interval adaptive Dormand-Prince 5(4)
integration of the Lindblad master equation (qoc/core/lindbladdiscrete.py:357-441, qoc/core/mathmethods.py:169-206,
:211-480); backward = hand-written reverse mode of the Runge-Kutta map on the REALISED grid (accepted step sizes and
positions held constant; dense output :263-304 and FSAL :477 included).

Why the controller is not differentiated.  The reference's autograd tape also differentiates the step-size
controller (step sizes are boxes).  Those extra terms are driven by err = rms((y1 - y1h) / 1e-12), a difference
taken at the rounding floor, so they are noise: perturbing the controls by 1e-13 (relative) changes the
oracle's full gradient by 1e-5 .. 1e-4 relative on benign problems and by O(1) on tests/golden/lindblad_case_0
(where the full gradient is 8000x off the finite-difference gradient of the reference forward), while the
frozen-grid gradient moves by < 1e-8.  The frozen-grid gradient differs from the full one by 4e-6 .. 5e-5 on
the benign golden cases, i.e. it lies inside the reference's own reproducibility band, and it agrees with
oracle/qoc_oracle.py `freeze_steps=True` to 1e-8 (tests/test_oracle_golden.py).

*** TEST INFRASTRUCTURE ONLY (same rule as qoc_oracle.py). ***

Cotangents follow HIPS-autograd's unconjugated convention (C = A B: Abar = Cbar B^T, Bbar = A^T Cbar;
real parameters take the real part).  The control model is the real-linear structure the product extracts:
H(x) = H0 + sum_r x_r A_r, x = real control channels ([Re u, Im u] for complex controls).
"""
import numpy as np

A_T = ((), (1 / 5,), (3 / 40, 9 / 40), (44 / 45, -56 / 15, 32 / 9),
       (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
       (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656))
C_T = (0, 1 / 5, 3 / 10, 4 / 5, 8 / 9, 1)
B_T = (35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84, 0)
BH_T = (5179 / 57600, 0, 7571 / 16695, 393 / 640, -92097 / 339200, 187 / 2100, 1 / 40)
D_T = (-12715105075 / 11282082432, 0, 87487479700 / 32700410799, -10690763975 / 1880347072,
       701980252875 / 199316789632, -1453857185 / 822651844, 69997945 / 29380423)
ATOL = 1e-12
EXPO = -1 / 5


def rms(x):
    return np.sqrt(np.sum((x * np.conj(x)).real) / x.size)


def rms_bar(x, r, rbar):
    """cotangent of x for r = rms(x)."""
    return rbar * np.conj(x) / (x.size * r)


class Model(object):
    def __init__(self, h0, a_ops, gammas, lops, T, M):
        """h0 (n x n) or None, a_ops (KR x n x n), gammas (L,), lops (L x n x n) or None."""
        self.h0, self.a_ops = h0, a_ops
        self.have_h = h0 is not None
        self.have_l = lops is not None
        if self.have_l:
            self.gam, self.L = np.asarray(gammas, dtype=float), np.asarray(lops, dtype=complex)
            self.Ld = np.conj(np.swapaxes(self.L, -1, -2))
            self.K = self.Ld @ self.L
        self.T, self.M = T, M
        self.xs = np.linspace(0, T, M) if M > 0 else None

    # -- controls at time t -------------------------------------------------------------------------
    def locate(self, t):
        xs = self.xs
        if t <= xs[0]:
            return 0, 1
        if t >= xs[-1]:
            return len(xs) - 2, len(xs) - 1
        i1 = int(np.argmax(t <= xs))
        return i1 - 1, i1

    def ham(self, x, t):
        if not self.have_h:
            return None, None
        if self.M == 0 or self.a_ops.shape[0] == 0:
            return self.h0, None
        i0, i1 = self.locate(t)
        slope = (x[i1] - x[i0]) / (self.xs[i1] - self.xs[i0])
        c = x[i0] + slope * (t - self.xs[i0])
        return self.h0 + np.tensordot(c, self.a_ops, axes=(0, 0)), (i0, i1, slope)

    # -- rhs and its vjp ------------------------------------------------------------------------------
    def rhs(self, x, t, rho):
        out = np.zeros_like(rho)
        if self.have_h:
            h, _ = self.ham(x, t)
            out = out + (-1j) * (h @ rho - rho @ h)
        if self.have_l:
            for l in range(self.L.shape[0]):
                out = out + self.gam[l] * (self.L[l] @ rho @ self.Ld[l] - 0.5 * (self.K[l] @ rho) - 0.5 * (rho @ self.K[l]))
        return out

    def rhs_vjp(self, x, t, rho, rbar, xbar):
        """returns (rho_bar, t_bar) and accumulates the control cotangent into xbar (M x KR)."""
        rho_bar = np.zeros_like(rho)
        tbar = 0.0
        if self.have_h:
            h, info = self.ham(x, t)
            ht = h.T
            rho_bar = rho_bar + (-1j) * (ht @ rbar - rbar @ ht)
            if info is not None:
                i0, i1, slope = info
                rt = np.swapaxes(rho, -1, -2)
                hbar = ((-1j) * (rbar @ rt - rt @ rbar)).sum(axis=0)
                cbar = np.real(np.einsum("ab,rab->r", hbar, self.a_ops))
                w = (t - self.xs[i0]) / (self.xs[i1] - self.xs[i0])
                xbar[i0] += cbar * (1 - w)
                xbar[i1] += cbar * w
                tbar = float(np.dot(cbar, slope))
        if self.have_l:
            for l in range(self.L.shape[0]):
                kt = self.K[l].T
                rho_bar = rho_bar + self.gam[l] * (self.L[l].T @ rbar @ np.conj(self.L[l]) - 0.5 * (kt @ rbar) - 0.5 * (rbar @ kt))
        return rho_bar, tbar

    # -- one RK attempt ------------------------------------------------------------------------------
    def attempt(self, x, t0, y0, k1, h):
        ks = [k1]
        zs = [None]
        for i in range(1, 6):
            acc = 0
            for j in range(i):
                acc = acc + A_T[i][j] * ks[j]
            zs.append(acc)
            ks.append(self.rhs(x, t0 + C_T[i] * h, y0 + h * acc))
        eb = B_T[0] * ks[0] + B_T[2] * ks[2] + B_T[3] * ks[3] + B_T[4] * ks[4] + B_T[5] * ks[5]
        y1 = y0 + h * eb
        ks.append(self.rhs(x, t0 + h, y1))
        ebh = (BH_T[0] * ks[0] + BH_T[2] * ks[2] + BH_T[3] * ks[3] + BH_T[4] * ks[4] + BH_T[5] * ks[5]
               + BH_T[6] * ks[6])
        y1h = y0 + h * ebh
        err = rms((y1 - y1h) / ATOL)
        return ks, zs, y1, y1h, err

    def attempt_bwd(self, x, t0, y0, k1, h, y1bar, k7bar, errbar, xbar, dense=None):
        """reverse of `attempt` (+ optional dense output seed).  Returns (y0bar, k1bar, hbar, t0bar)."""
        ks, zs, y1, y1h, err = self.attempt(x, t0, y0, k1, h)
        kbar = [np.zeros_like(y0) for _ in range(7)]
        kbar[6] = kbar[6] + k7bar
        y0bar = np.zeros_like(y0)
        y1bar = y1bar.copy()
        hbar, tbar = 0.0, 0.0
        if dense is not None:                      # out = dense(ks, t0, t0 + h, x_eval, y0, y1), seed outbar
            x_eval, outbar = dense
            th = (x_eval - t0) / h
            r2 = y1 - y0
            r3 = y0 + h * ks[0] - y1
            r4 = 2 * (y1 - y0) - h * (ks[0] + ks[6])
            sd = D_T[0] * ks[0] + D_T[2] * ks[2] + D_T[3] * ks[3] + D_T[4] * ks[4] + D_T[5] * ks[5] + D_T[6] * ks[6]
            r5 = h * sd
            # out = y0 + th (r2 + r3) - th^2 (r3 - r4 - r5) - th^3 (r4 + 2 r5) + th^4 r5
            c2, c3, c4, c5 = th, th - th ** 2, th ** 2 - th ** 3, th ** 2 - 2 * th ** 3 + th ** 4
            dth = ((r2 + r3) - 2 * th * (r3 - r4 - r5) - 3 * th ** 2 * (r4 + 2 * r5) + 4 * th ** 3 * r5)
            thbar = np.real(np.sum(outbar * dth))
            r2b, r3b, r4b, r5b = c2 * outbar, c3 * outbar, c4 * outbar, c5 * outbar
            y0bar = y0bar + outbar - r2b + r3b - 2 * r4b
            y1bar = y1bar + r2b - r3b + 2 * r4b
            kbar[0] = kbar[0] + h * r3b - h * r4b
            kbar[6] = kbar[6] - h * r4b
            hbar += np.real(np.sum(r3b * ks[0])) - np.real(np.sum(r4b * (ks[0] + ks[6]))) + np.real(np.sum(r5b * sd))
            for i in (0, 2, 3, 4, 5, 6):
                kbar[i] = kbar[i] + h * D_T[i] * r5b
            # th = (x_eval - t0) / h
            hbar += thbar * (-(x_eval - t0) / h ** 2)
            tbar += thbar * (-1.0 / h)
        if errbar != 0.0:
            e = (y1 - y1h) / ATOL
            dbar = rms_bar(e, err, errbar) / ATOL             # cotangent of (y1 - y1h)
            y1bar = y1bar + dbar
            # y1h = y0 + h * ebh
            y0bar = y0bar - dbar
            ebh = (BH_T[0] * ks[0] + BH_T[2] * ks[2] + BH_T[3] * ks[3] + BH_T[4] * ks[4] + BH_T[5] * ks[5]
                   + BH_T[6] * ks[6])
            hbar -= np.real(np.sum(dbar * ebh))
            for i in (0, 2, 3, 4, 5, 6):
                kbar[i] = kbar[i] - h * BH_T[i] * dbar
        # ks[6] = rhs(t0 + h, y1)
        rb, tb = self.rhs_vjp(x, t0 + h, y1, kbar[6], xbar)
        y1bar = y1bar + rb
        tbar += tb
        hbar += tb
        # y1 = y0 + h * eb
        eb = B_T[0] * ks[0] + B_T[2] * ks[2] + B_T[3] * ks[3] + B_T[4] * ks[4] + B_T[5] * ks[5]
        y0bar = y0bar + y1bar
        hbar += np.real(np.sum(y1bar * eb))
        for i in (0, 2, 3, 4, 5):
            kbar[i] = kbar[i] + h * B_T[i] * y1bar
        for i in range(5, 0, -1):
            zb, tb = self.rhs_vjp(x, t0 + C_T[i] * h, y0 + h * zs[i], kbar[i], xbar)
            y0bar = y0bar + zb
            hbar += np.real(np.sum(zb * zs[i])) + C_T[i] * tb
            tbar += tb
            for j in range(i):
                kbar[j] = kbar[j] + h * A_T[i][j] * zb
        return y0bar, kbar[0], hbar, tbar

    # -- one interval ----------------------------------------------------------------------------------
    def integrate(self, x, t0, tf, y, replay=None):
        """returns (y(tf), tape).  `replay` = a tape of an earlier run: every data-dependent decision
        (accept/reject, min/max branches, loop length) is taken from it instead of from the data, which makes the
        result a smooth function of the controls (used to validate the hand adjoint by finite differences)."""
        f0 = self.rhs(x, t0, y)
        d0, d1 = rms(y), rms(f0)
        small = (d0 < 1e-5 or d1 < 1e-5) if replay is None else replay["small"]
        h0 = 1e-6 if small else 0.01 * d0 / d1
        f1 = self.rhs(x, t0 + h0, y + h0 * f0)
        d2 = rms(f1 - f0) / h0
        d1_wins = (d1 >= d2) if replay is None else replay["d1_wins"]
        mx = d1 if d1_wins else d2
        tiny = (mx <= 1e-15) if replay is None else replay["tiny"]
        if tiny:
            h1_floor = (not h0 * 1e-3 > 1e-6) if replay is None else replay["h1_floor"]
            h1 = 1e-6 if h1_floor else h0 * 1e-3
        else:
            h1_floor = False
            h1 = (0.01 / mx) ** (1 / 6)
        h0_wins = (100 * h0 <= h1) if replay is None else replay["h0_wins"]
        step = 100 * h0 if h0_wins else h1
        tape = dict(t0=t0, tf=tf, y_in=y, small=small, tiny=tiny, h0=h0, h1=h1, d0=d0, d1=d1, d2=d2, f0=f0, f1=f1,
                    d1_wins=d1_wins, h1_floor=h1_floor, h0_wins=h0_wins, steps=[])
        xc, yc, k1 = t0, y, f0
        out = None
        si = 0
        while (xc <= tf) if replay is None else (si < len(replay["steps"])):
            rs = None if replay is None else replay["steps"][si]
            si += 1
            rejected = False
            attempts = []
            while True:
                ks, zs, y1, y1h, err = self.attempt(x, xc, yc, k1, step)
                attempts.append((step, err))
                accept = (err < 1) if rs is None else (len(attempts) == len(rs["attempts"]))
                if accept:
                    if rs is not None:
                        mode = rs["fac_mode"]
                        fac = 0.9 * err ** EXPO if mode == "pow" else rs["fac"]
                    else:
                        if err == 0:
                            fac, mode = 10.0, "const"
                        else:
                            raw = 0.9 * err ** EXPO
                            fac, mode = (10.0, "const") if raw >= 10.0 else (raw, "pow")
                        if rejected and fac >= 1.0:
                            fac, mode = 1.0, "const"
                    step_next = step * fac
                    break
                rejected = True
                raw = 0.9 * err ** EXPO
                floor = (raw <= 0.2) if rs is None else rs["rej_floor"][len(attempts) - 1]
                attempts[-1] = (step, err, floor)
                step = step * (0.2 if floor else raw)
            xn = xc + step
            hit = (xc <= tf and tf <= xn) if rs is None else rs["hit"]
            if hit:
                hh = xn - xc
                th = (tf - xc) / hh
                r2 = y1 - yc
                r3 = yc + hh * ks[0] - y1
                r4 = 2 * (y1 - yc) - hh * (ks[0] + ks[6])
                r5 = hh * (D_T[0] * ks[0] + D_T[2] * ks[2] + D_T[3] * ks[3] + D_T[4] * ks[4] + D_T[5] * ks[5] + D_T[6] * ks[6])
                out = yc + th * (r2 + r3) - th ** 2 * (r3 - r4 - r5) - th ** 3 * (r4 + 2 * r5) + th ** 4 * r5
            tape["steps"].append(dict(x=xc, y=yc, k1=k1, attempts=[(a[0], a[1]) for a in attempts], fac_mode=mode, fac=fac,
                                      hit=hit, rej_floor=[a[2] for a in attempts[:-1]]))
            xc, yc, k1, step = xn, y1, ks[6], step_next
        return out, tape

    def integrate_bwd_frozen(self, x, tape, outbar, xbar):
        """as integrate_bwd but with the realised step sizes and step positions treated as constants (the
        discrete adjoint of the RK map on the accepted grid): no controller terms, no rejected attempts."""
        steps = tape["steps"]
        last_hit = max(i for i, s in enumerate(steps) if s["hit"])
        ybar_next = np.zeros_like(outbar)
        k1bar_next = np.zeros_like(outbar)
        for i in range(last_hit, -1, -1):           # steps after the output step do not influence it
            s = steps[i]
            dense = (tape["tf"], outbar) if i == last_hit else None
            ybar_next, k1bar_next, _, _ = self.attempt_bwd(x, s["x"], s["y"], s["k1"], s["attempts"][-1][0], ybar_next,
                                                           k1bar_next, 0.0, xbar, dense)
        zb, _ = self.rhs_vjp(x, tape["t0"], tape["y_in"], k1bar_next, xbar)      # k1 of the first step = rhs(t0, y_in)
        return ybar_next + zb


class DensityTerm(object):
    """kind 0: w * (1 - sum_d |tr(T_d^dagger rho_d)| / (D n))      (targetdensityinfidelity.py:41-69)
       kind 1: w * sum_d (1/F_d) sum_f |tr(F_df^dagger rho_d) / n|^2   (forbiddensities.py:53-85)
       mats: list over densities of (F_d x n x n) arrays; w = multiplier / normalisation."""
    def __init__(self, kind, mats, w, step):
        self.kind, self.mats, self.w, self.step = kind, mats, w, step

    def value_and_seed(self, rho):
        D, n = rho.shape[0], rho.shape[1]
        seed = np.zeros_like(rho)
        if self.kind == 0:
            tot = 0.0
            for d in range(D):
                z = np.sum(np.conj(self.mats[d][0]) * rho[d])
                tot += abs(z)
                if abs(z) > 0:
                    seed[d] = -self.w / (D * n) * (np.conj(z) / abs(z)) * np.conj(self.mats[d][0])
            return self.w * (1 - tot / (D * n)), seed
        tot = 0.0
        for d in range(D):
            F = self.mats[d].shape[0]
            for f in range(F):
                ip = np.sum(np.conj(self.mats[d][f]) * rho[d]) / n
                tot += abs(ip) ** 2 / F
                seed[d] += self.w * 2 * np.conj(ip) / F * np.conj(self.mats[d][f]) / n
        return self.w * tot, seed


def cost_and_grad(x, model, rho0, terms, T, N, cost_eval_step=1, want_grad=True, replay=None, keep=None):
    """x: (M x KR) real control channels (or None).  Returns (cost, grad (M x KR), final densities, stats).
    `keep` (a list) receives the interval tapes; `replay` = such a list replays its decisions."""
    dt = T / (N - 1)
    rho = np.asarray(rho0, dtype=complex)
    tapes, cost = [], 0.0
    seeds = {}
    for step in range(N):
        st = step % cost_eval_step == 0 and step != 0
        fin = step == N - 1
        sd = np.zeros_like(rho)
        for t in terms:
            if (t.step and st) or (not t.step and fin):
                v, s_ = t.value_and_seed(rho)
                cost += v
                sd = sd + s_
        seeds[step] = sd
        if not fin:
            rho, tape = model.integrate(x, step * dt, step * dt + dt, rho, None if replay is None else replay[step])
            tapes.append(tape)
    if keep is not None:
        keep.extend(tapes)
    stats = dict(accepted=sum(len(t["steps"]) for t in tapes),
                 attempts=sum(len(s["attempts"]) for t in tapes for s in t["steps"]))
    if not want_grad or x is None:
        return cost, None, rho, stats
    xbar = np.zeros_like(x)
    rbar = seeds[N - 1]
    for step in range(N - 2, -1, -1):
        rbar = model.integrate_bwd_frozen(x, tapes[step], rbar, xbar) + seeds[step]
    return cost, xbar, rho, stats
