// small.cuh - batched matrix exponential for small Hilbert dimensions (n <= 8), register resident.
//
// Reference algorithm, step for step: qoc/standard/functions/expm.py:210-252 (expm_pade: one-norm, scaling count,
// Pade-13 numerator / denominator (:153-159), numpy.linalg.solve = zgesv with partial pivoting, repeated squaring).
//
// The CTA-per-matrix kernels of expm_slice.cuh pad every dimension below 8 to one 8 x 8 DMMA tile and round-trip
// intermediates through shared and global memory: at n = 2 that is 1.6e8 matrices/s, 0.3 % of the HBM roofline.  Here a
// matrix never leaves registers between its load and its store:
//   * n <= 2: ONE THREAD per matrix, everything fully unrolled (k_expm_thread);
//   * n = 3, 4: one matrix per group of T = 4 lanes, lane r owns row r of every intermediate; a product C = X Y is, per
//     lane, sum_k X[r][k] * (row k of Y, broadcast from lane k of the group by shuffles) - 4 n^2 DFMA and 2 n^2 shuffles
//     per lane, no padding waste at n = 3 (k_expm_rows; the template also covers n <= 8 with T = 8, where it loses to the
//     DMMA tile path on register pressure - see capi.cu).
// On B200 the FP64 tensor and vector pipes have the same peak, so plain DFMA costs nothing against DMMA here.
// I/O is the caller's layout, [batch][n][n] interleaved complex128: a lane reads and writes its row with 16-byte
// accesses, a warp covers 32 / T whole matrices = one contiguous block.  Algorithmic traffic: 32 n^2 bytes per matrix.
// The pivot rule is exactly izamax's (first maximum of |re| + |im|), the scaling count exactly the reference's.
#pragma once
#include "expm_slice.cuh"

namespace qocb {

// ---- one thread per matrix (n = 1, 2) ----------------------------------------------------------------------------------
template <int N> struct TMat { double r[N][N], i[N][N]; };

template <int N>
__device__ __forceinline__ void tmul(TMat<N> &c, const TMat<N> &a, const TMat<N> &b) {
#pragma unroll
    for (int x = 0; x < N; ++x)
#pragma unroll
        for (int y = 0; y < N; ++y) {
            double re = 0., im = 0.;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                re = fma(a.r[x][k], b.r[k][y], re); re = fma(-a.i[x][k], b.i[k][y], re);
                im = fma(a.r[x][k], b.i[k][y], im); im = fma(a.i[x][k], b.r[k][y], im);
            }
            c.r[x][y] = re; c.i[x][y] = im;
        }
}

template <int N>
__global__ void __launch_bounds__(128) k_expm_thread(const double2 *__restrict__ in, double2 *__restrict__ out, long long batch) {
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < batch; b += (long long)gridDim.x * blockDim.x) {
        TMat<N> A;
#pragma unroll
        for (int x = 0; x < N; ++x)
#pragma unroll
            for (int y = 0; y < N; ++y) { const double2 v = in[b * N * N + x * N + y]; A.r[x][y] = v.x; A.i[x][y] = v.y; }
        double norm = 0.;                                           // one-norm: largest column sum of |a_xy|
#pragma unroll
        for (int y = 0; y < N; ++y) {
            double cs = 0.;
#pragma unroll
            for (int x = 0; x < N; ++x) cs += sqrt(A.r[x][y] * A.r[x][y] + A.i[x][y] * A.i[x][y]);
            norm = fmax(norm, cs);
        }
        int s = 0;
        if (!(norm < QOCB_THETA13)) { s = (int)ceil(log2(norm / QOCB_THETA13)); if (s < 0) s = 0; }
        const double scale = ldexp(1.0, -s);
#pragma unroll
        for (int x = 0; x < N; ++x)
#pragma unroll
            for (int y = 0; y < N; ++y) { A.r[x][y] *= scale; A.i[x][y] *= scale; }
        TMat<N> A2, A4, A6, P, Q, T;
        tmul<N>(A2, A, A); tmul<N>(A4, A2, A2); tmul<N>(A6, A2, A4);
        // Y = A6 W1 + b7 A6 + b5 A4 + b3 A2 + b1 I,   W1 = b13 A6 + b11 A4 + b9 A2    -> T, then Uo = A Y -> Q
#pragma unroll
        for (int x = 0; x < N; ++x)
#pragma unroll
            for (int y = 0; y < N; ++y) { T.r[x][y] = kB[13] * A6.r[x][y] + kB[11] * A4.r[x][y] + kB[9] * A2.r[x][y]; T.i[x][y] = kB[13] * A6.i[x][y] + kB[11] * A4.i[x][y] + kB[9] * A2.i[x][y]; }
        tmul<N>(P, A6, T);
#pragma unroll
        for (int x = 0; x < N; ++x)
#pragma unroll
            for (int y = 0; y < N; ++y) {
                P.r[x][y] += kB[7] * A6.r[x][y] + kB[5] * A4.r[x][y] + kB[3] * A2.r[x][y] + (x == y ? kB[1] : 0.);
                P.i[x][y] += kB[7] * A6.i[x][y] + kB[5] * A4.i[x][y] + kB[3] * A2.i[x][y];
            }
        tmul<N>(Q, A, P);                                           // Uo
        // Ve = A6 X1 + b6 A6 + b4 A4 + b2 A2 + b0 I,  X1 = b12 A6 + b10 A4 + b8 A2
#pragma unroll
        for (int x = 0; x < N; ++x)
#pragma unroll
            for (int y = 0; y < N; ++y) { T.r[x][y] = kB[12] * A6.r[x][y] + kB[10] * A4.r[x][y] + kB[8] * A2.r[x][y]; T.i[x][y] = kB[12] * A6.i[x][y] + kB[10] * A4.i[x][y] + kB[8] * A2.i[x][y]; }
        tmul<N>(P, A6, T);
#pragma unroll
        for (int x = 0; x < N; ++x)
#pragma unroll
            for (int y = 0; y < N; ++y) {
                const double vr = P.r[x][y] + kB[6] * A6.r[x][y] + kB[4] * A4.r[x][y] + kB[2] * A2.r[x][y] + (x == y ? kB[0] : 0.);
                const double vi = P.i[x][y] + kB[6] * A6.i[x][y] + kB[4] * A4.i[x][y] + kB[2] * A2.i[x][y];
                const double ur = Q.r[x][y], ui = Q.i[x][y];
                P.r[x][y] = vr + ur; P.i[x][y] = vi + ui;           // P = Ve + Uo
                Q.r[x][y] = vr - ur; Q.i[x][y] = vi - ui;           // Q = Ve - Uo
            }
        // Q R = P: Gaussian elimination with partial pivoting on [Q | P] (zgesv)
#pragma unroll
        for (int j = 0; j < N; ++j) {
            int piv = j;
            double best = fabs(Q.r[j][j]) + fabs(Q.i[j][j]);
#pragma unroll
            for (int x = j + 1; x < N; ++x) { const double m = fabs(Q.r[x][j]) + fabs(Q.i[x][j]); if (m > best) { best = m; piv = x; } }
#pragma unroll
            for (int x = j + 1; x < N; ++x)
                if (piv == x) {
#pragma unroll
                    for (int y = 0; y < N; ++y) {
                        double t_;
                        t_ = Q.r[j][y]; Q.r[j][y] = Q.r[x][y]; Q.r[x][y] = t_; t_ = Q.i[j][y]; Q.i[j][y] = Q.i[x][y]; Q.i[x][y] = t_;
                        t_ = P.r[j][y]; P.r[j][y] = P.r[x][y]; P.r[x][y] = t_; t_ = P.i[j][y]; P.i[j][y] = P.i[x][y]; P.i[x][y] = t_;
                    }
                }
            const cplx inv = crecip({Q.r[j][j], Q.i[j][j]});
#pragma unroll
            for (int x = j + 1; x < N; ++x) {
                const cplx l = cmul({Q.r[x][j], Q.i[x][j]}, inv);
#pragma unroll
                for (int y = j + 1; y < N; ++y) { Q.r[x][y] -= l.r * Q.r[j][y] - l.i * Q.i[j][y]; Q.i[x][y] -= l.r * Q.i[j][y] + l.i * Q.r[j][y]; }
#pragma unroll
                for (int y = 0; y < N; ++y) { P.r[x][y] -= l.r * P.r[j][y] - l.i * P.i[j][y]; P.i[x][y] -= l.r * P.i[j][y] + l.i * P.r[j][y]; }
            }
        }
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            const cplx inv = crecip({Q.r[j][j], Q.i[j][j]});
#pragma unroll
            for (int y = 0; y < N; ++y) { const cplx v = cmul({P.r[j][y], P.i[j][y]}, inv); P.r[j][y] = v.r; P.i[j][y] = v.i; }
#pragma unroll
            for (int x = 0; x < j; ++x)
#pragma unroll
                for (int y = 0; y < N; ++y) { P.r[x][y] -= Q.r[x][j] * P.r[j][y] - Q.i[x][j] * P.i[j][y]; P.i[x][y] -= Q.r[x][j] * P.i[j][y] + Q.i[x][j] * P.r[j][y]; }
        }
        for (int q_ = 0; q_ < s; ++q_) { tmul<N>(T, P, P); P = T; }
#pragma unroll
        for (int x = 0; x < N; ++x)
#pragma unroll
            for (int y = 0; y < N; ++y) out[b * N * N + x * N + y] = make_double2(P.r[x][y], P.i[x][y]);
    }
}

// ---- one matrix per group of T lanes, lane r = row r -------------------------------------------------------
template <int N> struct Row { double r[N], i[N]; };

// c = (row `lr` of X) * Y, Y's rows broadcast from the lanes of the group; every lane of the warp must call
template <int N, int T>
__device__ __forceinline__ void rmul(Row<N> &c, const Row<N> &x, const Row<N> &y) {
    constexpr unsigned FULL = 0xffffffffu;
#pragma unroll
    for (int j = 0; j < N; ++j) { c.r[j] = 0.; c.i[j] = 0.; }
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const double xr = x.r[k], xi = x.i[k];
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double yr = __shfl_sync(FULL, y.r[j], k, T), yi = __shfl_sync(FULL, y.i[j], k, T);
            c.r[j] = fma(xr, yr, c.r[j]); c.r[j] = fma(-xi, yi, c.r[j]);
            c.i[j] = fma(xr, yi, c.i[j]); c.i[j] = fma(xi, yr, c.i[j]);
        }
    }
}

template <int N, int T>
__global__ void __launch_bounds__(128) k_expm_rows(const double2 *__restrict__ in, double2 *__restrict__ out, long long batch) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int PER_WARP = 32 / T;
    const int lane = threadIdx.x & 31, lr = lane % T, grp = lane / T;
    const bool row_ok = lr < N;
    const long long warps = ((long long)gridDim.x * blockDim.x) >> 5, warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (long long base = warp0 * PER_WARP; base < batch; base += warps * PER_WARP) {     // warp-uniform trip count
        const long long b = base + grp;
        const bool live = b < batch && row_ok;
        Row<N> A;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double2 v = make_double2(0., 0.);
            if (live) v = in[(b * N + lr) * N + j];
            A.r[j] = v.x; A.i[j] = v.y;
        }
        double norm = 0.;
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double cs = sqrt(A.r[j] * A.r[j] + A.i[j] * A.i[j]);
#pragma unroll
            for (int o = T / 2; o > 0; o >>= 1) cs += __shfl_xor_sync(FULL, cs, o, T);
            norm = fmax(norm, cs);
        }
        int s = 0;
        if (!(norm < QOCB_THETA13)) { s = (int)ceil(log2(norm / QOCB_THETA13)); if (s < 0) s = 0; }
        const double scale = ldexp(1.0, -s);
#pragma unroll
        for (int j = 0; j < N; ++j) { A.r[j] *= scale; A.i[j] *= scale; }
        Row<N> A2, A4, A6, P, Q, W;
        rmul<N, T>(A2, A, A); rmul<N, T>(A4, A2, A2); rmul<N, T>(A6, A2, A4);
#pragma unroll
        for (int j = 0; j < N; ++j) { W.r[j] = kB[13] * A6.r[j] + kB[11] * A4.r[j] + kB[9] * A2.r[j]; W.i[j] = kB[13] * A6.i[j] + kB[11] * A4.i[j] + kB[9] * A2.i[j]; }
        rmul<N, T>(P, A6, W);                                       // Y
#pragma unroll
        for (int j = 0; j < N; ++j) {
            P.r[j] += kB[7] * A6.r[j] + kB[5] * A4.r[j] + kB[3] * A2.r[j] + (j == lr ? kB[1] : 0.);
            P.i[j] += kB[7] * A6.i[j] + kB[5] * A4.i[j] + kB[3] * A2.i[j];
        }
        rmul<N, T>(Q, A, P);                                        // Uo = A Y
#pragma unroll
        for (int j = 0; j < N; ++j) { W.r[j] = kB[12] * A6.r[j] + kB[10] * A4.r[j] + kB[8] * A2.r[j]; W.i[j] = kB[12] * A6.i[j] + kB[10] * A4.i[j] + kB[8] * A2.i[j]; }
        rmul<N, T>(P, A6, W);
#pragma unroll
        for (int j = 0; j < N; ++j) {
            const double vr = P.r[j] + kB[6] * A6.r[j] + kB[4] * A4.r[j] + kB[2] * A2.r[j] + (j == lr ? kB[0] : 0.);
            const double vi = P.i[j] + kB[6] * A6.i[j] + kB[4] * A4.i[j] + kB[2] * A2.i[j];
            const double ur = Q.r[j], ui = Q.i[j];
            P.r[j] = vr + ur; P.i[j] = vi + ui; Q.r[j] = vr - ur; Q.i[j] = vi - ui;
        }
        // Q R = P with partial pivoting: rows are exchanged between lanes
#pragma unroll
        for (int j = 0; j < N; ++j) {
            double m = (lr >= j && lr < N) ? fabs(Q.r[j]) + fabs(Q.i[j]) : -1.0;      // izamax over rows j .. N-1 of column j
            int idx = lr;
#pragma unroll
            for (int o = T / 2; o > 0; o >>= 1) {
                const double m2 = __shfl_xor_sync(FULL, m, o, T);
                const int i2 = __shfl_xor_sync(FULL, idx, o, T);
                if (m2 > m || (m2 == m && i2 < idx)) { m = m2; idx = i2; }
            }
            const int src = lr == j ? idx : (lr == idx ? j : lr);    // swap rows j and idx
#pragma unroll
            for (int y = 0; y < N; ++y) {
                Q.r[y] = __shfl_sync(FULL, Q.r[y], src, T); Q.i[y] = __shfl_sync(FULL, Q.i[y], src, T);
                P.r[y] = __shfl_sync(FULL, P.r[y], src, T); P.i[y] = __shfl_sync(FULL, P.i[y], src, T);
            }
            const double pr = __shfl_sync(FULL, Q.r[j], j, T), pi = __shfl_sync(FULL, Q.i[j], j, T);
            const cplx inv = crecip({pr, pi});
            const cplx l = cmul({Q.r[j], Q.i[j]}, inv);
            const bool below = lr > j;
#pragma unroll
            for (int y = 0; y < N; ++y) {
                const double ur = __shfl_sync(FULL, Q.r[y], j, T), ui = __shfl_sync(FULL, Q.i[y], j, T);
                const double vr = __shfl_sync(FULL, P.r[y], j, T), vi = __shfl_sync(FULL, P.i[y], j, T);
                if (below) {
                    if (y > j) { Q.r[y] -= l.r * ur - l.i * ui; Q.i[y] -= l.r * ui + l.i * ur; }
                    P.r[y] -= l.r * vr - l.i * vi; P.i[y] -= l.r * vi + l.i * vr;
                }
            }
        }
#pragma unroll
        for (int j = N - 1; j >= 0; --j) {
            const double pr = __shfl_sync(FULL, Q.r[j], j, T), pi = __shfl_sync(FULL, Q.i[j], j, T);
            const cplx inv = crecip({pr, pi});
            const double ur = Q.r[j], ui = Q.i[j];                   // u_rj of this lane's row
#pragma unroll
            for (int y = 0; y < N; ++y) {
                if (lr == j) { const cplx v = cmul({P.r[y], P.i[y]}, inv); P.r[y] = v.r; P.i[y] = v.i; }
                const double vr = __shfl_sync(FULL, P.r[y], j, T), vi = __shfl_sync(FULL, P.i[y], j, T);
                if (lr < j) { P.r[y] -= ur * vr - ui * vi; P.i[y] -= ur * vi + ui * vr; }
            }
        }
        int smax = s;                                               // squarings: the warp runs the largest count of its matrices
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) smax = max(smax, __shfl_xor_sync(FULL, smax, o));
        for (int q_ = 0; q_ < smax; ++q_) {
            rmul<N, T>(W, P, P);
            if (q_ < s) P = W;
        }
        if (live) {
#pragma unroll
            for (int j = 0; j < N; ++j) out[(b * N + lr) * N + j] = make_double2(P.r[j], P.i[j]);
        }
    }
}

}  // namespace qocb
