// lowrank.cuh - reverse pass of the Pade graph exploiting the rank of the incoming cotangent (NP = 64, S <= 4, s = 0).
//
// In GRAPE the cotangent of a slice propagator is an outer-product sum of S costate/state pairs,
//     ubar = sum_s lam_{j+1,s} psi_{j,s}^T        (S << n),
// and without squarings (s = 0) the reverse of the Pade-13 graph keeps that structure for most of its length.  With
// l = Q^-T lam, psi' = R0 psi = psi_{j+1}, p = psi + psi', m = psi - psi', a = A^T l (all n x S):
//     uobar = l p^T, vebar = l m^T, ybar = a p^T, abar1 = l (Y p)^T
//     a6bar = Lb R6^T,  a4bar(b-part) = Lb R4^T,  a2bar(b-part) = Lb R2^T        with
//     Lb = [a, l, A6^T a, A6^T l],  R6 = [b7 p + W1 p, b6 m + X1 m, b13 p, b12 m],
//     R4 = [b5 p, b4 m, b11 p, b10 m],  R2 = [b3 p, b2 m, b9 p, b8 m]                        (n x 4S each)
//     a2bar = Lb (R2 + A4 R6 + A2 R4)^T + (A2^T Lb) (A2 R6 + R4)^T + (A2^T A2^T Lb) R6^T      (rank 12 S <= 48)
//     mbar  = abar1 + a2bar A^T + A^T a2bar.
// Eleven of the thirteen n^3 products of the dense reverse pass (pade_backward, expm_slice.cuh) become n^2 x 8 or
// n^2 x 16 "thin" DMMA products plus one rank-48 product: 4.75 instead of 13 matmul-equivalents, and the transposed
// solve runs on one 8-column tile.  The algebra is the same graph (oracle/adjoint_model.py:pade_bwd), re-associated.
//
// Shared memory: X0 = staging of the tape matrix in use (then a2bar); LEFT [64][52] x 2 planes and TMP [64][12] x 2 in
// the X1 region; RIGHT [64][52] x 2 in the X2 region.
#pragma once
#include "tile.cuh"

namespace qocb {

constexpr int LR_LD = 52, LR_PL = 64 * LR_LD, LR_TLD = 12, LR_TPL = 64 * LR_TLD;

// acc (8 x 8 tile, rows ar0.., cols of B bc0..) += sign * op(A)[ar0:+8, ak0:+8] * B[bk0:+8, bc0:+8]; A is a C-layout
// matrix (stride C::LD), B a thin matrix with row stride LDB and plane stride PLB
template <class C, bool T, int MASK, bool NEG, int LDB, int PLB>
__device__ __forceinline__ void tile_mma_thin(c2 &acc, const double *A, int ar0, int ak0, const double *B, int bk0, int bc0) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double p1a = 0., p1b = 0., p2a = 0., p2b = 0., p3a = 0., p3b = 0.;   // 3M partial products (tile.cuh: Acc)
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
        double ar, ai;
        ld_afrag<C, T, MASK>(A, ar0, ak0 + 4 * ks, g, t, ar, ai);
        if (NEG) { ar = -ar; ai = -ai; }
        const int bidx = (bk0 + 4 * ks + t) * LDB + bc0 + g;
        const double br = B[bidx], bi = B[PLB + bidx];
        dmma884(p1a, p1b, ar, br);
        dmma884(p2a, p2b, ai, bi);
        dmma884(p3a, p3b, ar + ai, br + bi);
    }
    acc.r0 += p1a - p2a; acc.r1 += p1b - p2b;
    acc.i0 += p3a - p1a - p2a; acc.i1 += p3b - p1b - p2b;
}
template <int LD, int PL> __device__ __forceinline__ c2 ld_thin(const double *M, int r0, int c0) {
    const int lane = threadIdx.x & 31, row = r0 + (lane >> 2), col = c0 + (lane & 3) * 2;
    const double2 r = *reinterpret_cast<const double2 *>(M + row * LD + col);
    const double2 i = *reinterpret_cast<const double2 *>(M + PL + row * LD + col);
    return {r.x, r.y, i.x, i.y};
}
template <int LD, int PL> __device__ __forceinline__ void st_thin(double *M, int r0, int c0, const c2 &v) {
    const int lane = threadIdx.x & 31, row = r0 + (lane >> 2), col = c0 + (lane & 3) * 2;
    *reinterpret_cast<double2 *>(M + row * LD + col) = make_double2(v.r0, v.r1);
    *reinterpret_cast<double2 *>(M + PL + row * LD + col) = make_double2(v.i0, v.i1);
}

// one 8 x 8 output tile (row tile rt) of op(M) * B[:, c0:c0+8], M = 64 x 64 in C layout.  The three 3M partial products
// are accumulated over the whole contraction and recombined once (dependent DMMAs are three issues apart)
template <class C, bool TA, int LDB, int PLB>
__device__ __forceinline__ c2 thin_tile(const double *M, const double *B, int rt, int c0) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double p1a = 0., p1b = 0., p2a = 0., p2b = 0., p3a = 0., p3b = 0.;
#pragma unroll 4
    for (int ks = 0; ks < C::NP / 4; ++ks) {
        double ar, ai;
        ld_afrag<C, TA, MASK_NONE>(M, rt * 8, 4 * ks, g, t, ar, ai);
        const int bidx = (4 * ks + t) * LDB + c0 + g;
        const double br = B[bidx], bi = B[PLB + bidx];
        dmma884(p1a, p1b, ar, br);
        dmma884(p2a, p2b, ai, bi);
        dmma884(p3a, p3b, ar + ai, br + bi);
    }
    return {p1a - p2a, p1b - p2b, p3a - p1a - p2a, p3b - p1b - p2b};
}

// Two thin products with ONE matrix in one pass (Hermitian M only): v1 = M * B1[:, c1:c1+8] and v2 = M^T * B2[:, c2:c2+8],
// the second as conj(M * conj(B2)) since M^T = conj(M) - both read the same A fragments, and the six 3M chains interleave
template <class C, int LDB1, int PLB1, int LDB2, int PLB2>
__device__ __forceinline__ void thin_tile_pair(const double *M, const double *B1, int col1, const double *B2, int col2, int rt,
                                               c2 &v1, c2 &v2) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double p1a = 0., p1b = 0., p2a = 0., p2b = 0., p3a = 0., p3b = 0.;
    double q1a = 0., q1b = 0., q2a = 0., q2b = 0., q3a = 0., q3b = 0.;
#pragma unroll 4
    for (int ks = 0; ks < C::NP / 4; ++ks) {
        double ar, ai;
        ld_afrag<C, false, MASK_NONE>(M, rt * 8, 4 * ks, g, t, ar, ai);
        const double as = ar + ai;
        const int i1 = (4 * ks + t) * LDB1 + col1 + g, i2 = (4 * ks + t) * LDB2 + col2 + g;
        const double br = B1[i1], bi = B1[PLB1 + i1];
        const double cr = B2[i2], ci = -B2[PLB2 + i2];
        dmma884(p1a, p1b, ar, br);
        dmma884(q1a, q1b, ar, cr);
        dmma884(p2a, p2b, ai, bi);
        dmma884(q2a, q2b, ai, ci);
        dmma884(p3a, p3b, as, br + bi);
        dmma884(q3a, q3b, as, cr + ci);
    }
    v1 = {p1a - p2a, p1b - p2b, p3a - p1a - p2a, p3b - p1b - p2b};
    v2 = {q1a - q2a, q1b - q2b, -(q3a - q1a - q2a), -(q3b - q1b - q2b)};
}

// acc += L[:, lc0:lc0+K] * R[:, rc0:rc0+K]^T  (L, R thin matrices; K a multiple of 4), warp tiling of Cfg
template <class C, int LDL, int PLL, int LDR, int PLR>
__device__ __forceinline__ void mma_lowrank(Acc<C> &acc, const double *L, int lc0, const double *R, int rc0, int K) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    const int row0 = (warp / C::WN) * C::TM * 8 + g, col0 = (warp % C::WN) * C::TN * 8 + g;
    for (int k0 = 0; k0 < K; k0 += 4) {
        double ar[C::TM], ai[C::TM], br[C::TN], bi[C::TN];
#pragma unroll
        for (int i = 0; i < C::TM; ++i) { const int idx = (row0 + i * 8) * LDL + lc0 + k0 + t; ar[i] = L[idx]; ai[i] = L[PLL + idx]; }
#pragma unroll
        for (int j = 0; j < C::TN; ++j) { const int idx = (col0 + j * 8) * LDR + rc0 + k0 + t; br[j] = R[idx]; bi[j] = R[PLR + idx]; }
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
            for (int j = 0; j < C::TN; ++j) {
                dmma884(acc.v[i][j][0], acc.v[i][j][1], ar[i], br[j]);
                dmma884(acc.v[i][j][2], acc.v[i][j][3], ai[i], bi[j]);
                dmma884(acc.v[i][j][4], acc.v[i][j][5], ar[i] + ai[i], br[j] + bi[j]);
            }
    }
}

// W[:, 0:8] <- Q^-T W[:, 0:8] (LUi factors in LU, C layout; W thin with LR_LD / LR_PL); the row un-permutation is left
// to the caller: result row i belongs to row perm[i].  Diagonal tiles by warp 0, the other row tiles spread over the warps.
template <class C, int LDW = LR_LD, int PLW = LR_PL>
__device__ void lu_solve_thin_T(const double *LU, double *W) {
    constexpr int NB = C::NP / 8;
    const int warp = threadIdx.x >> 5;
    for (int kb = 0; kb < NB; ++kb) {                              // U^T y = b
        if (warp == 0) {
            c2 z = czero();
            tile_mma_thin<C, true, MASK_UINV, false, LDW, PLW>(z, LU, kb * 8, kb * 8, W, kb * 8, 0);
            __syncwarp();
            st_thin<LDW, PLW>(W, kb * 8, 0, z);
        }
        __syncthreads();
        for (int rt = kb + 1 + warp; rt < NB; rt += C::NWARP) {
            c2 a = ld_thin<LDW, PLW>(W, rt * 8, 0);
            tile_mma_thin<C, true, MASK_NONE, true, LDW, PLW>(a, LU, rt * 8, kb * 8, W, kb * 8, 0);
            st_thin<LDW, PLW>(W, rt * 8, 0, a);
        }
        __syncthreads();
    }
    for (int kb = NB - 1; kb >= 0; --kb) {                         // L^T z = y
        if (warp == 0) {
            c2 z = czero();
            tile_mma_thin<C, true, MASK_LINV, false, LDW, PLW>(z, LU, kb * 8, kb * 8, W, kb * 8, 0);
            __syncwarp();
            st_thin<LDW, PLW>(W, kb * 8, 0, z);
        }
        __syncthreads();
        for (int rt = warp; rt < kb; rt += C::NWARP) {
            c2 a = ld_thin<LDW, PLW>(W, rt * 8, 0);
            tile_mma_thin<C, true, MASK_NONE, true, LDW, PLW>(a, LU, rt * 8, kb * 8, W, kb * 8, 0);
            st_thin<LDW, PLW>(W, rt * 8, 0, a);
        }
        __syncthreads();
    }
}

}  // namespace qocb
