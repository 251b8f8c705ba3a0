// large.cuh - GRAPE hot path for Hilbert dimensions above 64 (up to 512), where one matrix set no longer fits the
// shared memory of one CTA.  Same algorithm as expm_slice.cuh / sweep.cuh (qoc/core/schroedingerdiscrete.py:356-502,
// qoc/core/mathmethods.py:72-122, qoc/standard/functions/expm.py:210-252 and the reverse of that graph), organised as
// BATCHED level-3 work over many time slices at once:
//   * every matrix product is one strided-batched complex GEMM over the slices of a batch: the DMMA tile kernel of
//     zgemm.cuh (64 x 64 tiles, cp.async double-buffered K panels; ~85 % of cuBLAS ZGEMM on this pipeline;
//     QOCB_LARGE_CUBLAS=1 switches to the library GEMM for A/B comparison);
//   * the Pade denominator solve is cuBLAS batched LU (zgetrf/zgetrs) - the one library call left on this path.  The forward uses R = P Q^-1 (P and Q are
//     polynomials in A and commute, so this equals the reference's Q^-1 P) because that form needs no transposes in
//     row-major storage; the reverse pass differentiates exactly this form;
//   * everything else - generator assembly from interpolated controls, Magnus combination, one-norm / scaling, the Pade
//     polynomial combinations, squaring selection, rank-S cotangent seeds, Magnus adjoint, contraction with the control
//     operators, state / costate sweeps with the fused cost reductions - are kernels of this file.
// The reverse pass recomputes the forward intermediates per batch (no tape across the evaluation), so memory stays
// O(batch) + the propagators.  Magnus M2, M4 (product-free when KR <= 6) and M6 (batched commutator chain).
// Matrices are interleaved complex128 (double2), row-major, dense n x n; state vectors stay planar ([s][2][n]) so the
// cost reductions of sweep.cuh are shared.
#pragma once
#include <cooperative_groups.h>
#include <cublas_v2.h>

namespace qocb {

constexpr int kLgThreads = kSweepThreads;      // the shared cost helpers stride by kSweepThreads
constexpr int kLgMaxSq = 6;                    // squarings kept per batch for the reverse pass

struct LgCoef {                                // interpolation of the controls at the Magnus nodes of the local slices
    const double *controls; const int *itab_idx; const double *itab_w; int KR, q;
    const double *nodecoef;                    // precomputed channel coefficients [slice][q][KR] (time-dependent operators) or nullptr
};

__device__ __forceinline__ double lg_coef(const LgCoef &c, int j, int i, int r) {
    if (c.nodecoef) return c.nodecoef[(size_t)(j * c.q + i) * c.KR + r];
    const int o = (j * c.q + i) * 2;
    return c.controls[c.itab_idx[o] * c.KR + r] * c.itab_w[o] + c.controls[c.itab_idx[o + 1] * c.KR + r] * c.itab_w[o + 1];
}

// node generators a_i[b] = G0 + sum_r c_{j,i,r} G_r for the slices j = j_begin + b of a batch
__global__ void k_lg_assemble(double2 *a1, double2 *a2, const double2 *G0, const double2 *G, LgCoef c, int j_begin, int B, int nn) {
    const size_t tot = (size_t)B * nn;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / nn), e = (int)(t - (size_t)b * nn);
        const double2 g0 = G0[e];
        double2 v1 = g0, v2 = g0;
        for (int r = 0; r < c.KR; ++r) {
            const double2 g = G[(size_t)r * nn + e];
            const double c1 = lg_coef(c, j_begin + b, 0, r);
            v1.x += c1 * g.x; v1.y += c1 * g.y;
            if (c.q > 1) { const double c2 = lg_coef(c, j_begin + b, 1, r); v2.x += c2 * g.x; v2.y += c2 * g.y; }
        }
        a1[t] = v1;
        if (c.q > 1) a2[t] = v2;
    }
}

// Magnus M6 combinations at the three Gauss-Legendre nodes (mathmethods.py:129-164): b1 = dt a2,
// b2 = (sqrt15/3) dt (a3 - a1), b3 = (10/3) dt (a3 - 2 a2 + a1); the drift cancels in b2 and b3
__global__ void k_lg_assemble6(double2 *b1, double2 *b2, double2 *b3, const double2 *G0, const double2 *G, LgCoef c,
                               int j_begin, int B, int nn, double dt) {
    const size_t tot = (size_t)B * nn;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / nn), e = (int)(t - (size_t)b * nn), j = j_begin + b;
        double2 v1 = make_double2(dt * G0[e].x, dt * G0[e].y), v2 = make_double2(0., 0.), v3 = v2;
        for (int r = 0; r < c.KR; ++r) {
            const double2 g = G[(size_t)r * nn + e];
            const double c1 = lg_coef(c, j, 0, r), c2 = lg_coef(c, j, 1, r), c3 = lg_coef(c, j, 2, r);
            const double w1 = dt * c2, w2 = (QOCB_S15 / 3.0) * dt * (c3 - c1), w3 = (10.0 / 3.0) * dt * (c3 - 2.0 * c2 + c1);
            v1.x += w1 * g.x; v1.y += w1 * g.y;
            v2.x += w2 * g.x; v2.y += w2 * g.y;
            v3.x += w3 * g.x; v3.y += w3 * g.y;
        }
        b1[t] = v1; b2[t] = v2; b3[t] = v3;
    }
}

// product-free Magnus M4 (see expm_slice.cuh): M[b] = dt G0 + sum_r dt/2 (c1_r + c2_r) G_r
//                                                    + f [ sum_r (c1_r - c2_r) C0_r + sum_{s<r} (c2_s c1_r - c2_r c1_s) C_sr ]
__global__ void k_lg_magnus4_comm(double2 *M, const double2 *G0, const double2 *G, const double2 *C0, const double2 *Cs, LgCoef c,
                                  int j_begin, int B, int nn, double dt) {
    const size_t tot = (size_t)B * nn;
    const double f = (QOCB_S3 / 12.0) * dt * dt;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / nn), e = (int)(t - (size_t)b * nn), j = j_begin + b;
        double2 v = make_double2(dt * G0[e].x, dt * G0[e].y);
        for (int r = 0; r < c.KR; ++r) {
            const double c1 = lg_coef(c, j, 0, r), c2 = lg_coef(c, j, 1, r);
            const double2 g = G[(size_t)r * nn + e], k = C0[(size_t)r * nn + e];
            const double wg = 0.5 * dt * (c1 + c2), wk = f * (c1 - c2);
            v.x += wg * g.x + wk * k.x; v.y += wg * g.y + wk * k.y;
        }
        for (int s_ = 0; s_ < c.KR; ++s_)
            for (int r = s_ + 1; r < c.KR; ++r) {
                const double w = f * (lg_coef(c, j, 1, s_) * lg_coef(c, j, 0, r) - lg_coef(c, j, 1, r) * lg_coef(c, j, 0, s_));
                const double2 k = Cs[(size_t)comm_pair(s_, r, c.KR) * nn + e];
                v.x += w * k.x; v.y += w * k.y;
            }
        M[t] = v;
    }
}

// adjoint of the product-free Magnus M4: node_grad from inner products of mbar[b] with G_r, C0_r, C_sr; one CTA per slice
__global__ void __launch_bounds__(256) k_lg_contract_comm(const double2 *mbar, const double2 *G, const double2 *C0, const double2 *Cs,
                                                          LgCoef c, double *node_grad, int j_begin, int nn, double dt) {
    __shared__ double red[256];
    __shared__ double dots[2 * kMaxCommKR + kMaxCommKR * (kMaxCommKR - 1) / 2];
    const int b = blockIdx.x, KR = c.KR, NPAIR = KR * (KR - 1) / 2, j = j_begin + b;
    const double2 *m = mbar + (size_t)b * nn;
    for (int q = 0; q < 2 * KR + NPAIR; ++q) {
        const double2 *g = q < KR ? G + (size_t)q * nn : (q < 2 * KR ? C0 + (size_t)(q - KR) * nn : Cs + (size_t)(q - 2 * KR) * nn);
        double s = 0.;
        for (int e = threadIdx.x; e < nn; e += 256) s += m[e].x * g[e].x - m[e].y * g[e].y;
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) dots[q] = red[0];
        __syncthreads();
    }
    const double f = (QOCB_S3 / 12.0) * dt * dt;
    for (int e = threadIdx.x; e < 2 * KR; e += 256) {
        const int node = e / KR, r = e % KR;
        double acc = dots[KR + r];
        for (int s_ = 0; s_ < KR; ++s_) {
            if (s_ == r) continue;
            const double k = s_ < r ? dots[2 * KR + comm_pair(s_, r, KR)] : -dots[2 * KR + comm_pair(r, s_, KR)];
            acc += lg_coef(c, j, node == 0 ? 1 : 0, s_) * k;              // node 0 pairs with c2, node 1 with c1
        }
        node_grad[((size_t)j * 2 + node) * KR + r] = 0.5 * dt * dots[r] + (node == 0 ? f : -f) * acc;
    }
}

// out = alpha x + beta y + gamma z (any of y, z may be null)
__global__ void k_lg_axpby(double2 *out, double alpha, const double2 *x, double beta, const double2 *y, double gamma, const double2 *z, size_t tot) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        double2 v = make_double2(alpha * x[t].x, alpha * x[t].y);
        if (y) { v.x += beta * y[t].x; v.y += beta * y[t].y; }
        if (z) { v.x += gamma * z[t].x; v.y += gamma * z[t].y; }
        out[t] = v;
    }
}

// one-norm, squaring count and scaling per slice (expm.py:103-116, :229-241); one CTA per slice, in place
// lu_flag (optional): raised when a scaled matrix has ||A||_1 >= QOCB_NOPIV_NORM - the block LU without inter-block
// pivoting (below) is only used under that bound
__global__ void __launch_bounds__(256) k_lg_norm_scale(double2 *A, int *sarr, int n, int *lu_flag = nullptr) {
    __shared__ double red[256];
    double2 *a = A + (size_t)blockIdx.x * n * n;
    double best = 0.;
    for (int c = threadIdx.x; c < n; c += 256) {
        double s = 0.;
        for (int r = 0; r < n; ++r) { const double2 v = a[(size_t)r * n + c]; s += sqrt(v.x * v.x + v.y * v.y); }
        best = fmax(best, s);
    }
    red[threadIdx.x] = best;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] = fmax(red[threadIdx.x], red[threadIdx.x + o]); __syncthreads(); }
    const double norm = red[0];
    int s = 0;
    if (!(norm < QOCB_THETA13)) { s = (int)ceil(log2(norm / QOCB_THETA13)); if (s < 0) s = 0; }
    if (threadIdx.x == 0) sarr[blockIdx.x] = s;
    const double scale = ldexp(1.0, -s);
    if (lu_flag != nullptr && threadIdx.x == 0 && !(norm * scale < QOCB_NOPIV_NORM)) *lu_flag = 1;
    for (int e = threadIdx.x; e < n * n; e += 256) { a[e].x *= scale; a[e].y *= scale; }
}

// W1, X1 and the additive parts of Y and Ve (expm.py:153-159)
__global__ void k_lg_poly(const double2 *A2, const double2 *A4, const double2 *A6, double2 *W1, double2 *X1, double2 *Y, double2 *Ve, int n, size_t tot) {
    const int nn = n * n;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int e = (int)(t % nn);
        const bool diag = (e / n) == (e % n);
        const double2 a2 = A2[t], a4 = A4[t], a6 = A6[t];
        W1[t] = make_double2(kB[13] * a6.x + kB[11] * a4.x + kB[9] * a2.x, kB[13] * a6.y + kB[11] * a4.y + kB[9] * a2.y);
        X1[t] = make_double2(kB[12] * a6.x + kB[10] * a4.x + kB[8] * a2.x, kB[12] * a6.y + kB[10] * a4.y + kB[8] * a2.y);
        Y[t] = make_double2(kB[7] * a6.x + kB[5] * a4.x + kB[3] * a2.x + (diag ? kB[1] : 0.), kB[7] * a6.y + kB[5] * a4.y + kB[3] * a2.y);
        Ve[t] = make_double2(kB[6] * a6.x + kB[4] * a4.x + kB[2] * a2.x + (diag ? kB[0] : 0.), kB[6] * a6.y + kB[4] * a4.y + kB[2] * a2.y);
    }
}

// P = Ve + Uo, Q = Ve - Uo
__global__ void k_lg_pq(const double2 *Ve, const double2 *Uo, double2 *P, double2 *Q, size_t tot) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const double2 v = Ve[t], u = Uo[t];
        P[t] = make_double2(v.x + u.x, v.y + u.y);
        Q[t] = make_double2(v.x - u.x, v.y - u.y);
    }
}

// dst[b] = (level < s_b) ? src[b] : dst[b]   (squaring / reverse-squaring selection per slice); gate: see k_zgemm
__global__ void k_lg_select(double2 *dst, const double2 *src, const int *sarr, int level, int nn, size_t tot, const int *gate = nullptr) {
    if (gate != nullptr && level >= *gate) return;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x)
        if (level < sarr[t / nn]) dst[t] = src[t];
}
// gated copy (squaring inputs kept for the reverse pass)
__global__ void k_lg_copy_gated(double2 *dst, const double2 *src, size_t tot, const int *gate, int level) {
    if (level >= *gate) return;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) dst[t] = src[t];
}
// largest squaring count of a batch -> *smax (device); counts above `cap` raise the device error flag
__global__ void k_lg_batch_max(const int *sarr, int B, int *smax, int cap, int *err_flag) {
    __shared__ int red[256];
    int m = 0;
    for (int b = threadIdx.x; b < B; b += 256) m = max(m, sarr[b]);
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] = max(red[threadIdx.x], red[threadIdx.x + o]); __syncthreads(); }
    if (threadIdx.x == 0) { *smax = red[0]; if (red[0] > cap) *err_flag = 1; }
}

// x[b] *= 2^-s_b
__global__ void k_lg_unscale(double2 *x, const int *sarr, int nn, size_t tot) {
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const double f = ldexp(1.0, -sarr[t / nn]);
        x[t].x *= f; x[t].y *= f;
    }
}

// ubar[b][a][c] = sum_s lam_{j+1,s}[a] psi_{j,s}[c]  (cotangent of U_j from states = U_j psi); vectors planar [step][s][2][n]
__global__ void k_lg_ubar(double2 *ubar, const double *psi, const double *lam, int j_begin, int B, int n, int S) {
    const int nn = n * n;
    const size_t tot = (size_t)B * nn, VS = (size_t)S * 2 * n;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / nn), e = (int)(t - (size_t)b * nn), a = e / n, c = e - a * n;
        const double *p = psi + (size_t)(j_begin + b) * VS, *l = lam + (size_t)(j_begin + b + 1) * VS;
        double sr = 0., si = 0.;
        for (int s = 0; s < S; ++s) {
            const double lr = l[s * 2 * n + a], li = l[s * 2 * n + n + a], pr = p[s * 2 * n + c], pi = p[s * 2 * n + n + c];
            sr += lr * pr - li * pi; si += lr * pi + li * pr;
        }
        ubar[t] = make_double2(sr, si);
    }
}

// ---- rank-S reverse pass (see lowrank.cuh for the algebra; here with R = P Q^-1, so the common factor sits on the right):
//   psit = Q^-1 psi,  lamp = R0^T lam,  pl = lam + lamp,  ml = lam - lamp,  a = A^T pl
//   uobar = pl psit^T, vebar = ml psit^T, ybar = a psit^T, abar1 = pl (Y psit)^T
//   Lb = [a, ml, A6^T a, A6^T ml],  R6 = [b7 psit + W1 psit, b6 psit + X1 psit, b13 psit, b12 psit],  R4, R2 likewise
// PSI[b][4][n] holds psi^T (the right-hand sides of the thin solve), LAM[b][n][4] the costates as columns.
__global__ void k_lr_load(double2 *PSI, double2 *LAM, const double *psi, const double *lam, int j_begin, int B, int n, int S) {
    const size_t tot = (size_t)B * n * 4, VS = (size_t)S * 2 * n;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / (4 * n)), r = (int)(t % (4 * n)), s = r / n, a = r % n;
        const double *p = psi + (size_t)(j_begin + b) * VS, *l = lam + (size_t)(j_begin + b + 1) * VS;
        const bool on = s < S;
        PSI[(size_t)b * 4 * n + (size_t)s * n + a] = on ? make_double2(p[s * 2 * n + a], p[s * 2 * n + n + a]) : make_double2(0., 0.);
        LAM[((size_t)b * n + a) * 4 + s] = on ? make_double2(l[s * 2 * n + a], l[s * 2 * n + n + a]) : make_double2(0., 0.);
    }
}

// from psit (PSI after the solve), lam, lamp: TMPV[:, 0:4] = pl, LEFT[:, 4:8] = ml, and the polynomial blocks of RIGHT
__global__ void k_lr_fill(const double2 *PSI, const double2 *LAM, const double2 *LAMP, double2 *PT, double2 *TMPV, double2 *LEFT,
                          double2 *RIGHT, int B, int n) {
    const size_t tot = (size_t)B * n * 4;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(t / (4 * n)), r = (int)(t % (4 * n)), a = r / 4, s = r % 4;
        const size_t row = (size_t)b * n + a;
        const double2 x = PSI[(size_t)b * 4 * n + (size_t)s * n + a], l = LAM[row * 4 + s], lp = LAMP[row * 4 + s];
        PT[row * 4 + s] = x;
        TMPV[row * 16 + s] = make_double2(l.x + lp.x, l.y + lp.y);
        LEFT[row * 48 + 4 + s] = make_double2(l.x - lp.x, l.y - lp.y);
        double2 *R = RIGHT + row * 48;
        const int cols[12] = {0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44};
        const int bi[12] = {3, 2, 9, 8, 5, 4, 11, 10, 7, 6, 13, 12};
#pragma unroll
        for (int q = 0; q < 12; ++q) R[cols[q] + s] = make_double2(kB[bi[q]] * x.x, kB[bi[q]] * x.y);
    }
}

// node_grad[(j*q + i)*KR + r] = Re sum_e abar_i[b][e] G_r[e]; one CTA per (slice of the batch, node)
__global__ void __launch_bounds__(256) k_lg_contract(const double2 *ab1, const double2 *ab2, const double2 *ab3, const double2 *G,
                                                     double *node_grad, int j_begin, int nn, int KR, int q) {
    __shared__ double red[256];
    const int b = blockIdx.x, i = blockIdx.y;
    const double2 *ab = (i == 0 ? ab1 : i == 1 ? ab2 : ab3) + (size_t)b * nn;
    for (int r = 0; r < KR; ++r) {
        const double2 *g = G + (size_t)r * nn;
        double s = 0.;
        for (int e = threadIdx.x; e < nn; e += 256) s += ab[e].x * g[e].x - ab[e].y * g[e].y;
        red[threadIdx.x] = s;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
        if (threadIdx.x == 0) node_grad[((size_t)(j_begin + b) * q + i) * KR + r] = red[0];
        __syncthreads();
    }
}

// ---- state / costate sweeps -----------------------------------------------------------------------------------------
// Same three-step chunk scheme as sweep.cuh (boundary pass over chunk propagators, independent local sweeps; affine
// recursion with particular parts backwards); the chunk propagators are one level of the batched pairwise GEMM tree.
// Every mat-vec streams its n x n matrix from HBM / L2: one warp per row with the rows of four passes in flight
// (register prefetch).  The costate sweeps read explicitly transposed copies U_j^T, P_c^T, so both directions use the
// coalesced row-streaming kernel.
__global__ void k_lg_transpose(double2 *dst, const double2 *src, int n, int batch) {
    __shared__ double2 tile[16][17];
    const int tiles = (n + 15) / 16;
    const size_t nn = (size_t)n * n;
    for (int w = blockIdx.x; w < batch * tiles * tiles; w += gridDim.x) {
        const int b = w / (tiles * tiles), t = w % (tiles * tiles), tr = t / tiles, tc = t % tiles;
        const int r = tr * 16 + threadIdx.y, c = tc * 16 + threadIdx.x;
        if (r < n && c < n) tile[threadIdx.y][threadIdx.x] = src[b * nn + (size_t)r * n + c];
        __syncthreads();
        const int r2 = tc * 16 + threadIdx.y, c2 = tr * 16 + threadIdx.x;
        if (r2 < n && c2 < n) dst[b * nn + (size_t)r2 * n + c2] = tile[threadIdx.x][threadIdx.y];
        __syncthreads();
    }
}

// out[s][a] = sum_b U[a][b] in[s][b]; vectors planar [s][2][n] in shared memory; ends with a barrier
// a_begin / a_end: output rows computed by this call (the cluster boundary passes split the rows over their CTAs)
template <int NPL>          // double2 per lane per row: n <= 32 * NPL
__device__ void lg_matvec_rows(double *out, const double *in, const double2 *U, int n, int S, int a_begin = 0, int a_end = -1) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = kLgThreads / 32, D = NPL <= 8 ? 4 : 2;
    if (a_end < 0) a_end = n;
    for (int s0 = 0; s0 < S; s0 += 4) {
        const int sc = min(4, S - s0);
        for (int a0 = a_begin + warp; a0 < a_end; a0 += NW * D) {
            double2 u[D][NPL];
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int a = a0 + d * NW;
#pragma unroll
                for (int k = 0; k < NPL; ++k) {
                    const int c = lane + 32 * k;
                    u[d][k] = (a < a_end && c < n) ? U[(size_t)a * n + c] : make_double2(0., 0.);
                }
            }
#pragma unroll
            for (int d = 0; d < D; ++d) {
                const int a = a0 + d * NW;
                double ar[4] = {0., 0., 0., 0.}, ai[4] = {0., 0., 0., 0.};
#pragma unroll
                for (int k = 0; k < NPL; ++k) {
                    const int c = lane + 32 * k;
                    if (c < n) {
#pragma unroll
                        for (int s = 0; s < 4; ++s)
                            if (s < sc) {
                                const double vr = in[(s0 + s) * 2 * n + c], vi = in[(s0 + s) * 2 * n + n + c];
                                ar[s] += u[d][k].x * vr - u[d][k].y * vi; ai[s] += u[d][k].x * vi + u[d][k].y * vr;
                            }
                    }
                }
#pragma unroll
                for (int s = 0; s < 4; ++s) {
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) { ar[s] += __shfl_xor_sync(0xffffffffu, ar[s], o); ai[s] += __shfl_xor_sync(0xffffffffu, ai[s], o); }
                    if (lane == 0 && s < sc && a < a_end) { out[(s0 + s) * 2 * n + a] = ar[s]; out[(s0 + s) * 2 * n + n + a] = ai[s]; }
                }
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void lg_matvec(double *out, const double *in, const double2 *U, int n, int S, int a_begin = 0, int a_end = -1) {
    if (n <= 128) lg_matvec_rows<4>(out, in, U, n, S, a_begin, a_end);
    else if (n <= 256) lg_matvec_rows<8>(out, in, U, n, S, a_begin, a_end);
    else lg_matvec_rows<16>(out, in, U, n, S, a_begin, a_end);
}

struct LgSweep {
    SweepArgs a;            // NP = n, N = local state count, chunk tables, cost tables, psi / lam / part / cost_part, ...
    const double2 *U, *UT;  // [N-1][n*n] propagators and their transposes
    const double2 *P, *PT;  // [nchunks][n*n] chunk propagators and their transposes
    int nchunks;
};

// (1) boundary states; grid = 1
__global__ void __launch_bounds__(kLgThreads) k_lg_boundary_fwd(LgSweep g) {
    extern __shared__ __align__(16) double sm_raw[];
    const SweepArgs &a = g.a;
    const int n = a.NP, S = a.S, VS = S * 2 * n;
    double *v0 = sm_raw, *v1 = v0 + VS;
    for (int i = threadIdx.x; i < VS; i += kLgThreads) { v0[i] = a.psi_in[i]; a.psi[i] = a.psi_in[i]; }
    __syncthreads();
    for (int c = 0; c < g.nchunks; ++c) {
        lg_matvec(v1, v0, g.P + (size_t)c * n * n, n, S);
        const int kend = a.chunk_begin[c + 1];
        for (int i = threadIdx.x; i < VS; i += kLgThreads) a.psi[(size_t)kend * VS + i] = v1[i];
        double *t = v0; v0 = v1; v1 = t;
        __syncthreads();
    }
}

// (2) local forward sweeps + cost values; grid = nchunks
__global__ void __launch_bounds__(kLgThreads) k_lg_sweep_fwd(LgSweep g) {
    extern __shared__ __align__(16) double sm_raw[];
    const SweepArgs &a = g.a;
    const int n = a.NP, S = a.S, VS = S * 2 * n;
    double *v0 = sm_raw, *v1 = v0 + VS, *ip = v1 + VS;
    const int c = blockIdx.x, jb = a.chunk_begin[c], je = a.chunk_begin[c + 1];
    for (int i = threadIdx.x; i < VS; i += kLgThreads) v0[i] = a.psi[(size_t)jb * VS + i];
    __syncthreads();
    double cost = 0.;
    for (int j = jb; j < je; ++j) {
        const int k = j + 1;
        if (k < je) {
            lg_matvec(v1, v0, g.U + (size_t)j * n * n, n, S);
            for (int i = threadIdx.x; i < VS; i += kLgThreads) a.psi[(size_t)k * VS + i] = v1[i];
        } else {
            for (int i = threadIdx.x; i < VS; i += kLgThreads) v1[i] = a.psi[(size_t)k * VS + i];
            __syncthreads();
        }
        const bool st = is_step_cost_state(k + a.j_off, a.ces), fin = (k + a.j_off == a.Nglob - 1);
        if (a.nterms > 0 && (st || fin)) {
            cost_inner_products(a, v1, ip, st, fin);
            cost += cost_value(a, ip, st, fin, k + a.j_off);
        }
        double *t = v0; v0 = v1; v1 = t;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.cost_part[c] = cost;
}

// (3a/3c) local backward sweeps; grid = nchunks
template <bool PARTICULAR>
__global__ void __launch_bounds__(kLgThreads) k_lg_sweep_bwd(LgSweep g) {
    extern __shared__ __align__(16) double sm_raw[];
    const SweepArgs &a = g.a;
    const int n = a.NP, S = a.S, VS = S * 2 * n;
    double *v0 = sm_raw, *v1 = v0 + VS, *ip = v1 + VS;
    const int c = blockIdx.x, jb = a.chunk_begin[c], je = a.chunk_begin[c + 1];
    for (int i = threadIdx.x; i < VS; i += kLgThreads) v0[i] = PARTICULAR ? 0. : a.lam[(size_t)je * VS + i];
    __syncthreads();
    const int jstop = PARTICULAR ? jb : jb + 1;
    for (int j = je - 1; j >= jstop; --j) {
        lg_matvec(v1, v0, g.UT + (size_t)j * n * n, n, S);
        if (a.nterms > 0 && is_step_cost_state(j + a.j_off, a.ces)) {
            cost_inner_products(a, a.psi + (size_t)j * VS, ip, true, false);
            cost_add_seed(a, ip, v1, true, false, j + a.j_off);
        }
        if (!PARTICULAR) for (int i = threadIdx.x; i < VS; i += kLgThreads) a.lam[(size_t)j * VS + i] = v1[i];
        double *t = v0; v0 = v1; v1 = t;
        __syncthreads();
    }
    if (PARTICULAR) for (int i = threadIdx.x; i < VS; i += kLgThreads) a.part[(size_t)c * VS + i] = v0[i];
}

// (3b) boundary costates; grid = 1
__global__ void __launch_bounds__(kLgThreads) k_lg_boundary_bwd(LgSweep g, int have_part) {
    extern __shared__ __align__(16) double sm_raw[];
    const SweepArgs &a = g.a;
    const int n = a.NP, S = a.S, VS = S * 2 * n;
    double *v0 = sm_raw, *v1 = v0 + VS, *ip = v1 + VS;
    for (int i = threadIdx.x; i < VS; i += kLgThreads) v0[i] = a.lam_in ? a.lam_in[i] : 0.;
    __syncthreads();
    if (a.nterms > 0 && a.add_final_seed) {
        const bool st = is_step_cost_state(a.N - 1 + a.j_off, a.ces);
        cost_inner_products(a, a.psi + (size_t)(a.N - 1) * VS, ip, st, true);
        cost_add_seed(a, ip, v0, st, true, a.N - 1 + a.j_off);
    }
    for (int i = threadIdx.x; i < VS; i += kLgThreads) a.lam[(size_t)(a.N - 1) * VS + i] = v0[i];
    for (int c = g.nchunks - 1; c >= 0; --c) {
        lg_matvec(v1, v0, g.PT + (size_t)c * n * n, n, S);
        if (have_part) { for (int i = threadIdx.x; i < VS; i += kLgThreads) v1[i] += a.part[(size_t)c * VS + i]; __syncthreads(); }
        const int kbeg = a.chunk_begin[c];
        for (int i = threadIdx.x; i < VS; i += kLgThreads) a.lam[(size_t)kbeg * VS + i] = v1[i];
        double *t = v0; v0 = v1; v1 = t;
        __syncthreads();
    }
    if (a.b_out) for (int i = threadIdx.x; i < VS; i += kLgThreads) a.b_out[i] = v0[i];
}

// ---- own batched LU of the Pade denominator for n > 64 -------------------------------------------------------------------
// Block LU with 64 x 64 blocks and no pivoting BETWEEN blocks, for anti-Hermitian arguments with ||A||_1 < 2.5 (see tile.cuh:
// the denominator then has a positive definite Hermitian part, so every leading principal block and every Schur
// complement is safely invertible).  Right-looking, per block column k:
//     D_k <- inverse of the current diagonal block         k_lg_blockinv: one CTA, the pivoted DMMA LU + solve of tile.cuh
//     U[k, j>k] <- D_k^-1 Q[k, j>k]                        batched DMMA GEMM (zgemm.cuh), in place (one row tile per CTA)
//     Q[i>k, j>k] -= Q[i>k, k] U[k, j>k]                   batched DMMA GEMM
// which leaves Q = L U with L block lower triangular (inverted diagonal blocks stored) and U unit block upper triangular.
// The solves X Q = P (forward pass) and X Q^T = R (reverse pass) are block substitutions made of the same GEMMs
// (capi.cu: lg_block_lu, lg_solve_right).  Everything else (general matrices, larger norms) stays on cuBLAS getrf / getrs.
template <class C>
__global__ void __launch_bounds__(C::NT) k_lg_blockinv(double2 *Q, int n, int k0, int bs, long long stride, int *lu_flag) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Smem<C> sm(smem_raw);
    double2 *q = Q + (size_t)blockIdx.x * stride + (size_t)k0 * n + k0;
    double mx = 0.;
    for (int idx = threadIdx.x; idx < C::NP * C::NP; idx += C::NT) {
        const int r = idx / C::NP, c = idx - r * C::NP;
        double2 v = make_double2(r == c ? 1.0 : 0.0, 0.0);          // identity padding decouples the unused rows / columns
        if (r < bs && c < bs) { v = q[(size_t)r * n + c]; mx = fmax(mx, fabs(v.x) + fabs(v.y)); }
        sm.X2[r * C::LD + c] = v.x; sm.X2[C::PLANE + r * C::LD + c] = v.y;
        sm.X1[r * C::LD + c] = r == c ? 1.0 : 0.0; sm.X1[C::PLANE + r * C::LD + c] = 0.0;
    }
    __syncthreads();
    lu_factor_blocked<C>(sm.X2, sm.piv, sm.piv + C::NP);
    lu_solve_blocked<C, false>(sm.X2, sm.piv, sm.X1, sm.X0);        // X0 = D^-1
    double mi = 0.;
    for (int idx = threadIdx.x; idx < C::NP * C::NP; idx += C::NT) {
        const int r = idx / C::NP, c = idx - r * C::NP;
        if (r < bs && c < bs) {
            const double2 v = make_double2(sm.X0[r * C::LD + c], sm.X0[C::PLANE + r * C::LD + c]);
            mi = fmax(mi, fabs(v.x) + fabs(v.y));
            q[(size_t)r * n + c] = v;
        }
    }
    // conditioning guard: max|D| * max|D^-1| * bs bounds the 1-norm condition number from above
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o)); mi = fmax(mi, __shfl_xor_sync(0xffffffffu, mi, o)); }
    __shared__ double gmx[32], gmi[32];
    if ((threadIdx.x & 31) == 0) { gmx[threadIdx.x >> 5] = mx; gmi[threadIdx.x >> 5] = mi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < C::NWARP; ++w) { mx = fmax(mx, gmx[w]); mi = fmax(mi, gmi[w]); }
        if (!(mx * mi * bs < 1e10)) *lu_flag = 1;
    }
}

// dst[b][r][c] = src[b][r][c] for r < rows, c < cols (row strides ldd / lds, batch strides sd / ss)
__global__ void k_lg_copy_rect(double2 *dst, const double2 *src, int rows, int cols, int ldd, int lds, long long sd, long long ss, int batch) {
    const size_t per = (size_t)rows * cols, tot = per * batch;
    for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < tot; t += (size_t)gridDim.x * blockDim.x) {
        const size_t b = t / per, e = t - b * per;
        const int r = (int)(e / cols), c = (int)(e - (size_t)r * cols);
        dst[b * sd + (size_t)r * ldd + c] = src[b * ss + (size_t)r * lds + c];
    }
}

// ---- boundary passes on a thread-block cluster ------------------------------------------------------------------------
// The passes over the chunk boundaries are chains of dependent mat-vecs with an n x n propagator (1 MB at n = 256): one CTA
// pulls that through a single SM's L2 port (64 us per step at n = 256; 2 x 157 steps = 15 % of a cfg4 evaluation).  Here the
// rows of each mat-vec are split over the kLgCluster CTAs of ONE cluster: every CTA keeps the whole vector in its shared
// memory, computes its slice of the result, writes that slice into the next-vector buffer of every CTA of the cluster through
// distributed shared memory, and the cluster barrier ends the step.
constexpr int kLgCluster = 8;
namespace cgx = cooperative_groups;

// out-slice [a0, a1) of v1 (all states) -> the v1 buffer of every CTA of the cluster
__device__ __forceinline__ void lg_cluster_share(cgx::cluster_group &cl, double *v1, int n, int S, int a0, int a1) {
    const int len = a1 - a0;
    for (unsigned r = 0; r < cl.num_blocks(); ++r) {
        if (r == cl.block_rank()) continue;
        double *dst = cl.map_shared_rank(v1, r);
        for (int i = threadIdx.x; i < S * 2 * len; i += kLgThreads) {
            const int sp = i / len, a = a0 + i - sp * len;           // sp = state * 2 + plane
            dst[sp * n + a] = v1[sp * n + a];
        }
    }
}

__global__ void __cluster_dims__(kLgCluster, 1, 1) __launch_bounds__(kLgThreads) k_lg_boundary_fwd_cl(LgSweep g) {
    extern __shared__ __align__(16) double sm_raw[];
    cgx::cluster_group cl = cgx::this_cluster();
    const SweepArgs &a = g.a;
    const int n = a.NP, S = a.S, VS = S * 2 * n;
    const int rank = (int)cl.block_rank(), per = (n + kLgCluster - 1) / kLgCluster;
    const int a0 = min(n, rank * per), a1 = min(n, a0 + per);
    double *v0 = sm_raw, *v1 = v0 + VS;
    for (int i = threadIdx.x; i < VS; i += kLgThreads) { v0[i] = a.psi_in[i]; if (rank == 0) a.psi[i] = a.psi_in[i]; }
    cl.sync();
    for (int c = 0; c < g.nchunks; ++c) {
        lg_matvec(v1, v0, g.P + (size_t)c * n * n, n, S, a0, a1);   // ends with a block barrier
        lg_cluster_share(cl, v1, n, S, a0, a1);
        const int kend = a.chunk_begin[c + 1];
        for (int i = threadIdx.x; i < S * 2 * (a1 - a0); i += kLgThreads) {      // this CTA's slice of the boundary state
            const int sp = i / (a1 - a0), x = a0 + i - sp * (a1 - a0);
            a.psi[(size_t)kend * VS + sp * n + x] = v1[sp * n + x];
        }
        cl.sync();                                                  // every slice of v1 has arrived everywhere; v0 is free
        double *t = v0; v0 = v1; v1 = t;
    }
}

__global__ void __cluster_dims__(kLgCluster, 1, 1) __launch_bounds__(kLgThreads) k_lg_boundary_bwd_cl(LgSweep g, int have_part) {
    extern __shared__ __align__(16) double sm_raw[];
    cgx::cluster_group cl = cgx::this_cluster();
    const SweepArgs &a = g.a;
    const int n = a.NP, S = a.S, VS = S * 2 * n;
    const int rank = (int)cl.block_rank(), per = (n + kLgCluster - 1) / kLgCluster;
    const int a0 = min(n, rank * per), a1 = min(n, a0 + per);
    double *v0 = sm_raw, *v1 = v0 + VS, *ip = v1 + VS;
    for (int i = threadIdx.x; i < VS; i += kLgThreads) v0[i] = a.lam_in ? a.lam_in[i] : 0.;
    __syncthreads();
    if (a.nterms > 0 && a.add_final_seed) {                         // every CTA seeds its own full copy (same arithmetic everywhere)
        const bool st = is_step_cost_state(a.N - 1 + a.j_off, a.ces);
        cost_inner_products(a, a.psi + (size_t)(a.N - 1) * VS, ip, st, true);
        cost_add_seed(a, ip, v0, st, true, a.N - 1 + a.j_off);
    }
    if (rank == 0) for (int i = threadIdx.x; i < VS; i += kLgThreads) a.lam[(size_t)(a.N - 1) * VS + i] = v0[i];
    cl.sync();
    for (int c = g.nchunks - 1; c >= 0; --c) {
        lg_matvec(v1, v0, g.PT + (size_t)c * n * n, n, S, a0, a1);
        const int kbeg = a.chunk_begin[c];
        for (int i = threadIdx.x; i < S * 2 * (a1 - a0); i += kLgThreads) {
            const int sp = i / (a1 - a0), x = a0 + i - sp * (a1 - a0);
            if (have_part) v1[sp * n + x] += a.part[(size_t)c * VS + sp * n + x];
            a.lam[(size_t)kbeg * VS + sp * n + x] = v1[sp * n + x];
        }
        __syncthreads();
        lg_cluster_share(cl, v1, n, S, a0, a1);
        cl.sync();
        double *t = v0; v0 = v1; v1 = t;
    }
    if (a.b_out && rank == 0) for (int i = threadIdx.x; i < VS; i += kLgThreads) a.b_out[i] = v0[i];
}

// time sharding: psi_in = P_{rank-1} .. P_0 psi0   /   lam_in from the later shards (allPT: transposed shard propagators)
__global__ void __launch_bounds__(kLgThreads) k_lg_prefix(const double2 *allP, const double *psi0, double *psi_in, int rank, int n, int S) {
    extern __shared__ __align__(16) double sm_raw[];
    const int VS = S * 2 * n;
    double *v0 = sm_raw, *v1 = v0 + VS;
    for (int i = threadIdx.x; i < VS; i += kLgThreads) v0[i] = psi0[i];
    __syncthreads();
    for (int r = 0; r < rank; ++r) { lg_matvec(v1, v0, allP + (size_t)r * n * n, n, S); double *t = v0; v0 = v1; v1 = t; }
    for (int i = threadIdx.x; i < VS; i += kLgThreads) psi_in[i] = v0[i];
}
__global__ void __launch_bounds__(kLgThreads) k_lg_suffix(const double2 *allPT, const double *allb, double *lam_in, int rank, int world, int n, int S) {
    extern __shared__ __align__(16) double sm_raw[];
    const int VS = S * 2 * n;
    double *v0 = sm_raw, *v1 = v0 + VS;
    for (int i = threadIdx.x; i < VS; i += kLgThreads) v0[i] = 0.;
    __syncthreads();
    for (int r = world - 1; r > rank; --r) {
        lg_matvec(v1, v0, allPT + (size_t)r * n * n, n, S);
        for (int i = threadIdx.x; i < VS; i += kLgThreads) v1[i] += allb[(size_t)r * VS + i];
        __syncthreads();
        double *t = v0; v0 = v1; v1 = t;
    }
    for (int i = threadIdx.x; i < VS; i += kLgThreads) lam_in[i] = v0[i];
}

}  // namespace qocb
