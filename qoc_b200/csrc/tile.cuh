// tile.cuh - CTA-cooperative complex128 matrix primitives for sm_100a.
//
// One CTA owns one n x n complex matrix problem (n padded to NP).  Matrices are PLANAR (a real plane
// followed by an imaginary plane), row-major: dense NP x NP in global memory, row stride LD = NP + 4 in
// shared memory.  With LD = 4 (mod 8) every 64-bit fragment load of the FP64 tensor-core instruction
// (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4, the native FP64 MMA on sm_100a; tcgen05 has no f64 kind) is
// bank-conflict free for A, B, A^T and B^T operands alike.
//
// Element ownership: in every accumulator / epilogue / elementwise step a thread touches the SAME set of
// (row, col) pairs (the DMMA accumulator layout of its warp tile).  Dependent elementwise chains on
// global scratch matrices therefore need no barrier: a thread only ever re-reads what it wrote itself.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qocb {

template <int NP_, int WM_, int WN_>
struct Cfg {
    static constexpr int NP = NP_, WM = WM_, WN = WN_;
    static constexpr int NWARP = WM * WN, NT = 32 * NWARP;
    static constexpr int TM = NP / 8 / WM, TN = NP / 8 / WN;   // 8x8 tiles per warp
    static constexpr int LD = NP + 4;
    static constexpr int PLANE = NP * LD;                      // doubles per smem plane
    static constexpr int SMAT = 2 * PLANE;                     // doubles per smem matrix
    static constexpr int GPLANE = NP * NP;
    static constexpr int GMAT = 2 * NP * NP;                   // doubles per global matrix
    static constexpr int PARTS = NT / NP;                      // threads per column in triangular solves
    static_assert(NP % (8 * WM) == 0 && NP % (8 * WN) == 0, "warp tiling must divide NP");
    static_assert(NT % NP == 0 && PARTS >= 1 && PARTS <= 32 && (32 % PARTS) == 0, "bad PARTS");
};

template <class C>
struct Acc {
    double v[C::TM][C::TN][4];   // [0..1] real (col, col+1), [2..3] imaginary
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
            for (int j = 0; j < C::TN; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) v[i][j][e] = 0.0;
    }
};

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// acc += (+/-) op(A) * op(B), A and B resident in shared memory (planar, stride LD).
template <class C, bool TA, bool TB, bool NEG>
__device__ __forceinline__ void mma_smem(Acc<C> &acc, const double *__restrict__ A, const double *__restrict__ B) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int row0 = (warp / C::WN) * C::TM * 8 + g;
    const int col0 = (warp % C::WN) * C::TN * 8 + g;
#pragma unroll 2
    for (int kk = 0; kk < C::NP / 4; ++kk) {
        const int k = kk * 4 + t;
        double ar[C::TM], ai[C::TM], nai[C::TM], br[C::TN], bi[C::TN];
#pragma unroll
        for (int i = 0; i < C::TM; ++i) {
            const int r = row0 + i * 8;
            const int idx = TA ? (k * C::LD + r) : (r * C::LD + k);
            double xr = A[idx], xi = A[C::PLANE + idx];
            if (NEG) { xr = -xr; xi = -xi; }
            ar[i] = xr; ai[i] = xi; nai[i] = -xi;
        }
#pragma unroll
        for (int j = 0; j < C::TN; ++j) {
            const int c = col0 + j * 8;
            const int idx = TB ? (c * C::LD + k) : (k * C::LD + c);
            br[j] = B[idx]; bi[j] = B[C::PLANE + idx];
        }
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
            for (int j = 0; j < C::TN; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], ar[i], br[j]);
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
            for (int j = 0; j < C::TN; ++j) dmma884(acc.v[i][j][2], acc.v[i][j][3], ar[i], bi[j]);
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
            for (int j = 0; j < C::TN; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], nai[i], bi[j]);
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
            for (int j = 0; j < C::TN; ++j) dmma884(acc.v[i][j][2], acc.v[i][j][3], ai[i], br[j]);
    }
}

// Visit the element pairs this thread owns: f(tm, tn, row, col) owns (row, col) and (row, col + 1).
template <class C, class F>
__device__ __forceinline__ void for_owned(F &&f) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row0 = (warp / C::WN) * C::TM * 8 + (lane >> 2);
    const int col0 = (warp % C::WN) * C::TN * 8 + (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < C::TM; ++i)
#pragma unroll
        for (int j = 0; j < C::TN; ++j) f(i, j, row0 + i * 8, col0 + j * 8);
}

struct c2 { double r0, r1, i0, i1; };   // a complex element pair

template <class C> __device__ __forceinline__ c2 ldg2(const double *__restrict__ g, int row, int col) {
    const double2 r = *reinterpret_cast<const double2 *>(g + row * C::NP + col);
    const double2 i = *reinterpret_cast<const double2 *>(g + C::GPLANE + row * C::NP + col);
    return {r.x, r.y, i.x, i.y};
}
template <class C> __device__ __forceinline__ void stg2(double *__restrict__ g, int row, int col, const c2 &v) {
    *reinterpret_cast<double2 *>(g + row * C::NP + col) = make_double2(v.r0, v.r1);
    *reinterpret_cast<double2 *>(g + C::GPLANE + row * C::NP + col) = make_double2(v.i0, v.i1);
}
template <class C> __device__ __forceinline__ c2 lds2(const double *s, int row, int col) {
    const double2 r = *reinterpret_cast<const double2 *>(s + row * C::LD + col);
    const double2 i = *reinterpret_cast<const double2 *>(s + C::PLANE + row * C::LD + col);
    return {r.x, r.y, i.x, i.y};
}
template <class C> __device__ __forceinline__ void sts2(double *s, int row, int col, const c2 &v) {
    *reinterpret_cast<double2 *>(s + row * C::LD + col) = make_double2(v.r0, v.r1);
    *reinterpret_cast<double2 *>(s + C::PLANE + row * C::LD + col) = make_double2(v.i0, v.i1);
}
template <class C> __device__ __forceinline__ c2 accv(const Acc<C> &a, int i, int j) {
    return {a.v[i][j][0], a.v[i][j][1], a.v[i][j][2], a.v[i][j][3]};
}
__device__ __forceinline__ c2 operator+(const c2 &a, const c2 &b) { return {a.r0 + b.r0, a.r1 + b.r1, a.i0 + b.i0, a.i1 + b.i1}; }
__device__ __forceinline__ c2 operator-(const c2 &a, const c2 &b) { return {a.r0 - b.r0, a.r1 - b.r1, a.i0 - b.i0, a.i1 - b.i1}; }
__device__ __forceinline__ c2 operator*(double s, const c2 &a) { return {s * a.r0, s * a.r1, s * a.i0, s * a.i1}; }
__device__ __forceinline__ c2 czero() { return {0., 0., 0., 0.}; }
// real identity contribution on the diagonal of an owned pair
__device__ __forceinline__ c2 add_diag(c2 v, int row, int col, double d) {
    if (row == col) v.r0 += d;
    if (row == col + 1) v.r1 += d;
    return v;
}

// cooperative copy global (dense planar) -> shared (padded planar); caller provides the barriers
template <class C> __device__ __forceinline__ void g2s(double *__restrict__ s, const double *__restrict__ g) {
    constexpr int CH = C::GMAT / 2;            // 16-byte chunks
    constexpr int RCH = C::NP / 2;
    for (int idx = threadIdx.x; idx < CH; idx += C::NT) {
        const int plane = idx / (C::GPLANE / 2);
        const int rem = idx - plane * (C::GPLANE / 2);
        const int row = rem / RCH, cc = rem - row * RCH;
        const double2 v = reinterpret_cast<const double2 *>(g)[idx];
        *reinterpret_cast<double2 *>(s + plane * C::PLANE + row * C::LD + cc * 2) = v;
    }
}
template <class C> __device__ __forceinline__ void s2g(double *__restrict__ g, const double *__restrict__ s) {
    constexpr int CH = C::GMAT / 2;
    constexpr int RCH = C::NP / 2;
    for (int idx = threadIdx.x; idx < CH; idx += C::NT) {
        const int plane = idx / (C::GPLANE / 2);
        const int rem = idx - plane * (C::GPLANE / 2);
        const int row = rem / RCH, cc = rem - row * RCH;
        reinterpret_cast<double2 *>(g)[idx] = *reinterpret_cast<const double2 *>(s + plane * C::PLANE + row * C::LD + cc * 2);
    }
}

// ---- small complex helpers -------------------------------------------------------------------------
struct cplx { double r, i; };
__device__ __forceinline__ cplx cmul(cplx a, cplx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }
__device__ __forceinline__ cplx crecip(cplx a) {
    // Smith's algorithm (robust against overflow), as LAPACK's zladiv does for zgesv pivots
    if (fabs(a.r) >= fabs(a.i)) { const double q = a.i / a.r, d = a.r + a.i * q; return {1.0 / d, -q / d}; }
    const double q = a.r / a.i, d = a.r * q + a.i; return {q / d, -1.0 / d};
}

// ---- LU with partial pivoting, in place on a shared-memory matrix ---------------------------------------
// Follows zgesv's zgetf2 semantics (qoc/standard/functions/expm.py:246 -> numpy.linalg.solve -> LAPACK):
// pivot = first row maximising |re| + |im| (izamax), unit-lower L stored below the diagonal.
// red: shared scratch of >= 2 doubles; piv: shared int[NP].  Ends with a barrier.
template <class C>
__device__ void lu_factor_smem(double *__restrict__ Q, int *__restrict__ piv, double *__restrict__ red) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double *Qr = Q, *Qi = Q + C::PLANE;
    for (int k = 0; k < C::NP; ++k) {
        if (warp == 0) {
            double best = -1.0; int bi = k;
            for (int i = k + lane; i < C::NP; i += 32) {
                const double v = fabs(Qr[i * C::LD + k]) + fabs(Qi[i * C::LD + k]);
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_down_sync(0xffffffffu, best, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) piv[k] = bi;
        }
        __syncthreads();
        const int p = piv[k];
        if (p != k) {
            for (int c = tid; c < C::NP; c += C::NT) {
                double t0 = Qr[k * C::LD + c]; Qr[k * C::LD + c] = Qr[p * C::LD + c]; Qr[p * C::LD + c] = t0;
                t0 = Qi[k * C::LD + c]; Qi[k * C::LD + c] = Qi[p * C::LD + c]; Qi[p * C::LD + c] = t0;
            }
            __syncthreads();
        }
        const cplx inv = crecip({Qr[k * C::LD + k], Qi[k * C::LD + k]});
        const int m = C::NP - k - 1;           // trailing size
        // multipliers are formed on the fly from the unscaled column k; column k is scaled afterwards
        for (int e = tid; e < m * m; e += C::NT) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            const cplx l = cmul({Qr[i * C::LD + k], Qi[i * C::LD + k]}, inv);
            const cplx u = {Qr[k * C::LD + j], Qi[k * C::LD + j]};
            Qr[i * C::LD + j] -= l.r * u.r - l.i * u.i;
            Qi[i * C::LD + j] -= l.r * u.i + l.i * u.r;
        }
        __syncthreads();
        for (int i = k + 1 + tid; i < C::NP; i += C::NT) {
            const cplx l = cmul({Qr[i * C::LD + k], Qi[i * C::LD + k]}, inv);
            Qr[i * C::LD + k] = l.r; Qi[i * C::LD + k] = l.i;
        }
        // no barrier needed here: column k is not read again before the barrier after the next pivot search
    }
    __syncthreads();
}

// X <- Q^{-1} X (TRANS = false) or X <- Q^{-T} X (TRANS = true) with the factors from lu_factor_smem.
// PARTS threads cooperate on one column of X (rows split mod PARTS); the pivot-row value is broadcast
// inside the PARTS-lane group by shuffle, so no block barrier is needed inside the substitution loops.
template <class C, bool TRANS>
__device__ void lu_solve_smem(const double *__restrict__ LU, const int *__restrict__ piv, double *__restrict__ X) {
    const int tid = threadIdx.x;
    constexpr int P = C::PARTS;
    const int c = tid / P, part = tid % P;
    const int lane = tid & 31;
    const int gbase = lane - part;               // first lane of this column group
    const double *Lr = LU, *Li = LU + C::PLANE;
    double *Xr = X, *Xi = X + C::PLANE;
    if (!TRANS) {
        if (part == 0)
            for (int k = 0; k < C::NP; ++k) {
                const int p = piv[k];
                if (p != k) {
                    double t0 = Xr[k * C::LD + c]; Xr[k * C::LD + c] = Xr[p * C::LD + c]; Xr[p * C::LD + c] = t0;
                    t0 = Xi[k * C::LD + c]; Xi[k * C::LD + c] = Xi[p * C::LD + c]; Xi[p * C::LD + c] = t0;
                }
            }
        __syncwarp();
        // L y = x (unit lower)
        for (int k = 0; k < C::NP; ++k) {
            double xr = 0., xi = 0.;
            if (part == k % P) { xr = Xr[k * C::LD + c]; xi = Xi[k * C::LD + c]; }
            xr = __shfl_sync(0xffffffffu, xr, gbase + k % P);
            xi = __shfl_sync(0xffffffffu, xi, gbase + k % P);
            int i = k + 1; i += ((part - i) % P + P) % P;
            for (; i < C::NP; i += P) {
                const double lr = Lr[i * C::LD + k], li = Li[i * C::LD + k];
                Xr[i * C::LD + c] -= lr * xr - li * xi;
                Xi[i * C::LD + c] -= lr * xi + li * xr;
            }
        }
        // U z = y
        for (int k = C::NP - 1; k >= 0; --k) {
            double xr = 0., xi = 0.;
            if (part == k % P) {
                const cplx inv = crecip({Lr[k * C::LD + k], Li[k * C::LD + k]});
                const cplx v = cmul({Xr[k * C::LD + c], Xi[k * C::LD + c]}, inv);
                Xr[k * C::LD + c] = v.r; Xi[k * C::LD + c] = v.i; xr = v.r; xi = v.i;
            }
            xr = __shfl_sync(0xffffffffu, xr, gbase + k % P);
            xi = __shfl_sync(0xffffffffu, xi, gbase + k % P);
            for (int i = part; i < k; i += P) {
                const double ur = Lr[i * C::LD + k], ui = Li[i * C::LD + k];
                Xr[i * C::LD + c] -= ur * xr - ui * xi;
                Xi[i * C::LD + c] -= ur * xi + ui * xr;
            }
        }
    } else {
        // U^T y = x : forward substitution with rows of U
        for (int k = 0; k < C::NP; ++k) {
            double xr = 0., xi = 0.;
            if (part == k % P) {
                const cplx inv = crecip({Lr[k * C::LD + k], Li[k * C::LD + k]});
                const cplx v = cmul({Xr[k * C::LD + c], Xi[k * C::LD + c]}, inv);
                Xr[k * C::LD + c] = v.r; Xi[k * C::LD + c] = v.i; xr = v.r; xi = v.i;
            }
            xr = __shfl_sync(0xffffffffu, xr, gbase + k % P);
            xi = __shfl_sync(0xffffffffu, xi, gbase + k % P);
            int i = k + 1; i += ((part - i) % P + P) % P;
            for (; i < C::NP; i += P) {
                const double ur = Lr[k * C::LD + i], ui = Li[k * C::LD + i];
                Xr[i * C::LD + c] -= ur * xr - ui * xi;
                Xi[i * C::LD + c] -= ur * xi + ui * xr;
            }
        }
        // L^T z = y : backward substitution with rows of L (unit diagonal)
        for (int k = C::NP - 1; k >= 0; --k) {
            double xr = 0., xi = 0.;
            if (part == k % P) { xr = Xr[k * C::LD + c]; xi = Xi[k * C::LD + c]; }
            xr = __shfl_sync(0xffffffffu, xr, gbase + k % P);
            xi = __shfl_sync(0xffffffffu, xi, gbase + k % P);
            for (int i = part; i < k; i += P) {
                const double lr = Lr[k * C::LD + i], li = Li[k * C::LD + i];
                Xr[i * C::LD + c] -= lr * xr - li * xi;
                Xi[i * C::LD + c] -= lr * xi + li * xr;
            }
        }
        __syncwarp();
        if (part == 0)
            for (int k = C::NP - 1; k >= 0; --k) {
                const int p = piv[k];
                if (p != k) {
                    double t0 = Xr[k * C::LD + c]; Xr[k * C::LD + c] = Xr[p * C::LD + c]; Xr[p * C::LD + c] = t0;
                    t0 = Xi[k * C::LD + c]; Xi[k * C::LD + c] = Xi[p * C::LD + c]; Xi[p * C::LD + c] = t0;
                }
            }
    }
    __syncthreads();
}

// block-wide sum of NV doubles per thread -> result valid in all threads; red needs NV * NWARP doubles
template <class C, int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *__restrict__ red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
    }
    __syncthreads();
    if (lane == 0)
#pragma unroll
        for (int q = 0; q < NV; ++q) red[q * C::NWARP + warp] = v[q];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) {
        double s = 0.;
        for (int w = 0; w < C::NWARP; ++w) s += red[q * C::NWARP + w];
        v[q] = s;
    }
}

}  // namespace qocb
