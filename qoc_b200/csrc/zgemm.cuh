// zgemm.cuh - strided-batched complex128 GEMM on the FP64 tensor cores of sm_100a for the large-dimension path:
//     C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b],   op = identity or (unconjugated) transpose,
// row-major n x n `double2` operands in HBM / L2, any n (edges are zero-filled / masked).
// One CTA = one 64 x 64 (or, for thin operands of the rank-S reverse pass, 64 x 16) output tile; K is walked in panels of 32
// staged in shared memory by 16-byte cp.async (LDGSTS), double buffered; 8 warps in a 2 x 4 grid, each 32 x 16 outputs
// = 4 x 2 DMMA tiles (thin: 8 x 1 warps of 8 x 16); a complex product is 4 real
// mma.sync.m8n8k4.f64 per k-step (tcgen05 has no f64 kind).  Shared layouts are interleaved (re, im) with row strides
// 36 (A: rows m, cols k) and 66 (B: rows k, cols n) double2, which makes the 16-byte fragment loads of both operands
// bank-conflict free.
#pragma once
#include "tile.cuh"
#include "sweep.cuh"

namespace qocb {

constexpr int ZG_BM = 64, ZG_BK = 32, ZG_LDA = 36, ZG_NT = 256;
template <int BN> struct ZgTile {                 // BN = 64: 2 x 4 warps of 32 x 16; BN = 16 (thin operands): 8 x 1 warps of 8 x 16
    static constexpr int LDB = BN + 2;            // = 2 (mod 4) double2: conflict-free 16-byte B-fragment loads
    static constexpr int WM = BN == 64 ? 2 : 8, WN = BN == 64 ? 4 : 1, TM = ZG_BM / 8 / WM, TN = BN / 8 / WN;
    static constexpr size_t smem = sizeof(double2) * 2 * (ZG_BM * ZG_LDA + ZG_BK * LDB);
};

__device__ __forceinline__ void cp_async16_zfill(void *smem, const void *gmem, bool pred) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(sa), "l"(gmem), "r"(sz));
}

// C[b] (m x nc, ldc) = alpha * op(A[b]) (m x k) * op(B[b]) (k x nc) + beta * C[b]; row-major with leading dimensions.
// op(A)(i, kk) = TA ? A[kk * lda + i] : A[i * lda + kk];  op(B)(kk, c) = TB ? B[c * ldb + kk] : B[kk * ldb + c].
template <bool TA, bool TB, int BN>
__global__ void __launch_bounds__(ZG_NT) k_zgemm(const double2 *__restrict__ A, const double2 *__restrict__ B, double2 *C,
                                                 int m, int nc, int k, int lda, int ldb, int ldc, double alpha, double beta,
                                                 long long sA, long long sB, long long sC, const int *gate = nullptr, int gate_level = 0) {
    // gated launches (squarings of the scaling-and-squaring loop): the host enqueues a fixed number of them, the device skips
    // those beyond the largest squaring count of the batch - no host synchronisation to learn the count
    if (gate != nullptr && gate_level >= *gate) return;
    using Z = ZgTile<BN>;
    extern __shared__ __align__(16) unsigned char zg_raw[];
    double2 *As = reinterpret_cast<double2 *>(zg_raw);                       // [2][BM][LDA]
    double2 *Bs = As + 2 * ZG_BM * ZG_LDA;                                   // [2][BK][LDB]
    const int tiles_n = (nc + BN - 1) / BN;
    const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
    const int m0 = tm * ZG_BM, n0 = tn * BN;
    A += (size_t)blockIdx.y * sA; B += (size_t)blockIdx.y * sB; C += (size_t)blockIdx.y * sC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int row0 = (warp / Z::WN) * Z::TM * 8, col0 = (warp % Z::WN) * Z::TN * 8;

    auto load_panel = [&](int buf, int k0) {
        double2 *as = As + buf * ZG_BM * ZG_LDA, *bs = Bs + buf * ZG_BK * Z::LDB;
#pragma unroll
        for (int i = 0; i < (ZG_BM * ZG_BK) / ZG_NT; ++i) {
            const int e = tid + i * ZG_NT;
            int r, kk;
            if (TA) { kk = e / ZG_BM; r = e % ZG_BM; } else { r = e / ZG_BK; kk = e % ZG_BK; }
            const bool ok = (m0 + r < m) && (k0 + kk < k);
            const double2 *src = TA ? A + (size_t)(k0 + kk) * lda + m0 + r : A + (size_t)(m0 + r) * lda + k0 + kk;
            cp_async16_zfill(as + r * ZG_LDA + kk, ok ? src : A, ok);
        }
#pragma unroll
        for (int i = 0; i < (ZG_BK * BN + ZG_NT - 1) / ZG_NT; ++i) {
            const int e = tid + i * ZG_NT;
            if (e < ZG_BK * BN) {
                int kk, c;
                if (TB) { c = e / ZG_BK; kk = e % ZG_BK; } else { kk = e / BN; c = e % BN; }
                const bool ok = (n0 + c < nc) && (k0 + kk < k);
                const double2 *src = TB ? B + (size_t)(n0 + c) * ldb + k0 + kk : B + (size_t)(k0 + kk) * ldb + n0 + c;
                cp_async16_zfill(bs + kk * Z::LDB + c, ok ? src : B, ok);
            }
        }
    };

    // 3M complex products (tile.cuh: Acc): P1 = Ar Br, P2 = Ai Bi, P3 = (Ar + Ai)(Br + Bi); three DMMA per tile and k-step
    double acc[Z::TM][Z::TN][6];
#pragma unroll
    for (int i = 0; i < Z::TM; ++i)
#pragma unroll
        for (int j = 0; j < Z::TN; ++j)
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[i][j][q] = 0.0;

    const int panels = (k + ZG_BK - 1) / ZG_BK;
    load_panel(0, 0);
    cp_async_commit();
    for (int p = 0; p < panels; ++p) {
        const int buf = p & 1;
        if (p + 1 < panels) load_panel(buf ^ 1, (p + 1) * ZG_BK);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const double2 *as = As + buf * ZG_BM * ZG_LDA, *bs = Bs + buf * ZG_BK * Z::LDB;
#pragma unroll
        for (int kk = 0; kk < ZG_BK / 4; ++kk) {
            double2 a[Z::TM], b[Z::TN];
#pragma unroll
            for (int i = 0; i < Z::TM; ++i) a[i] = as[(row0 + i * 8 + g) * ZG_LDA + kk * 4 + t];
#pragma unroll
            for (int j = 0; j < Z::TN; ++j) b[j] = bs[(kk * 4 + t) * Z::LDB + col0 + j * 8 + g];
#pragma unroll
            double sa[Z::TM], sb[Z::TN];
#pragma unroll
            for (int i = 0; i < Z::TM; ++i) sa[i] = a[i].x + a[i].y;
#pragma unroll
            for (int j = 0; j < Z::TN; ++j) sb[j] = b[j].x + b[j].y;
#pragma unroll
            for (int i = 0; i < Z::TM; ++i)
#pragma unroll
                for (int j = 0; j < Z::TN; ++j) {
                    dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
                    dmma884(acc[i][j][2], acc[i][j][3], a[i].y, b[j].y);
                    dmma884(acc[i][j][4], acc[i][j][5], sa[i], sb[j]);
                }
        }
        __syncthreads();
    }
    cp_async_wait<0>();
#pragma unroll
    for (int i = 0; i < Z::TM; ++i)
#pragma unroll
        for (int j = 0; j < Z::TN; ++j) {
            const int r = m0 + row0 + i * 8 + g, c = n0 + col0 + j * 8 + 2 * t;
            if (r < m) {
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (c + q < nc) {
                        double2 *dst = C + (size_t)r * ldc + c + q;
                        double2 v = make_double2(alpha * (acc[i][j][q] - acc[i][j][2 + q]), alpha * (acc[i][j][4 + q] - acc[i][j][q] - acc[i][j][2 + q]));
                        if (beta != 0.0) { const double2 o = *dst; v.x += beta * o.x; v.y += beta * o.y; }
                        *dst = v;
                    }
            }
        }
}

}  // namespace qocb
