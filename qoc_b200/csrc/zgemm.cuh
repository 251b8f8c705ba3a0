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
#include <algorithm>
#include "tile.cuh"
#include "sweep.cuh"
#include "tma.cuh"

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

// ---- TMA-fed variant -----------------------------------------------------------------------------------------------------
// Same tiles, same warp layout, same 3M DMMA inner loop; the operand panels arrive as cp.async.bulk.tensor boxes (SASS
// UTMALDG) through a 3-stage mbarrier ring instead of per-thread 16-byte LDGSTS through a 2-stage one: one thread issues two
// boxes per K panel, nobody computes addresses or edge predicates (TMA zero-fills out-of-bounds rows / columns / K tails).
// The operands are described as rank-3 FLOAT64 tensors [batch][rows][2 * cols] (a double2 is two doubles), so a box of
// 2 * (w + pad) doubles by h rows lands with the padded row stride the conflict-free fragment loads want:
//     op(A) = A   : box 64 rows (m) x 36 double2 (k)  -> as[m][36]      op(A) = A^T : box 32 rows (k) x 66 double2 (m) -> as[k][66]
//     op(B) = B   : box 32 rows (k) x (BN + 2) (n)    -> bs[k][BN + 2]  op(B) = B^T : box BN rows (n) x 36 double2 (k) -> bs[n][36]
// (the pad columns hold the neighbouring data or zeros and are never read).  A transposed operand stays in its global
// orientation in shared memory - the 16-byte fragment loads are conflict free either way (8 lanes read 128 contiguous bytes).
constexpr int ZG_STAGES = 3, ZG_LDT = 66;
template <bool TA, bool TB, int BN> struct ZgTmaTile {
    static constexpr int A_ROWS = TA ? ZG_BK : ZG_BM, A_LD = TA ? ZG_LDT : ZG_LDA;                   // double2 units
    static constexpr int B_ROWS = TB ? BN : ZG_BK, B_LD = TB ? ZG_LDA : BN + 2;
    static constexpr int A_ELEMS = A_ROWS * A_LD, B_ELEMS = B_ROWS * B_LD;
    static constexpr int STAGE = ((A_ELEMS + B_ELEMS) * 16 + 127) / 128 * 128;                        // bytes, 128-byte aligned
    static constexpr int A_BYTES = (A_ELEMS * 16 + 127) / 128 * 128;
    static constexpr size_t smem = (size_t)ZG_STAGES * STAGE + 128 + ZG_STAGES * sizeof(uint64_t);
};
struct ZgMaps { CUtensorMap a, b; };

__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, int c0, int c1, int c2, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];\n"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

template <bool TA, bool TB, int BN>
__global__ void __launch_bounds__(ZG_NT) k_zgemm_tma(const __grid_constant__ ZgMaps maps, double2 *C, int m, int nc, int k, int ldc,
                                                     double alpha, double beta, long long sC, const int *gate = nullptr, int gate_level = 0) {
    if (gate != nullptr && gate_level >= *gate) return;
    using Z = ZgTile<BN>;
    using T = ZgTmaTile<TA, TB, BN>;
    extern __shared__ __align__(128) unsigned char zt_raw[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(zt_raw) + 127) & ~(uintptr_t)127);
    uint64_t *full = reinterpret_cast<uint64_t *>(base + (size_t)ZG_STAGES * T::STAGE);
    const int tiles_n = (nc + BN - 1) / BN;
    const int tm = blockIdx.x / tiles_n, tn = blockIdx.x % tiles_n;
    const int m0 = tm * ZG_BM, n0 = tn * BN, bz = blockIdx.y;
    C += (size_t)bz * sC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
    const int row0 = (warp / Z::WN) * Z::TM * 8, col0 = (warp % Z::WN) * Z::TN * 8;
    if (tid == 0) {
        for (int s_ = 0; s_ < ZG_STAGES; ++s_) mbar_init(full + s_, 1);
        mbar_init_fence();
    }
    __syncthreads();
    const int panels = (k + ZG_BK - 1) / ZG_BK;
    constexpr uint32_t kBytes = (uint32_t)((T::A_ELEMS + T::B_ELEMS) * 16);
    auto issue = [&](int p) {                               // one thread: both boxes of K panel p into stage p % ZG_STAGES
        unsigned char *st = base + (size_t)(p % ZG_STAGES) * T::STAGE;
        const int k0 = p * ZG_BK;
        fence_proxy_async();
        mbar_expect_tx(full + p % ZG_STAGES, kBytes);
        if (TA) tma_load_3d(st, &maps.a, 2 * m0, k0, bz, full + p % ZG_STAGES); else tma_load_3d(st, &maps.a, 2 * k0, m0, bz, full + p % ZG_STAGES);
        if (TB) tma_load_3d(st + T::A_BYTES, &maps.b, 2 * k0, n0, bz, full + p % ZG_STAGES); else tma_load_3d(st + T::A_BYTES, &maps.b, 2 * n0, k0, bz, full + p % ZG_STAGES);
    };
    if (tid == 0) for (int p = 0; p < ZG_STAGES - 1 && p < panels; ++p) issue(p);

    double acc[Z::TM][Z::TN][6];
#pragma unroll
    for (int i = 0; i < Z::TM; ++i)
#pragma unroll
        for (int j = 0; j < Z::TN; ++j)
#pragma unroll
            for (int q = 0; q < 6; ++q) acc[i][j][q] = 0.0;

    for (int p = 0; p < panels; ++p) {
        if (tid == 0 && p + ZG_STAGES - 1 < panels) issue(p + ZG_STAGES - 1);   // its stage was released by the barrier ending panel p - 1
        mbar_wait(full + p % ZG_STAGES, (uint32_t)((p / ZG_STAGES) & 1));
        const double2 *as = reinterpret_cast<const double2 *>(base + (size_t)(p % ZG_STAGES) * T::STAGE);
        const double2 *bs = reinterpret_cast<const double2 *>(base + (size_t)(p % ZG_STAGES) * T::STAGE + T::A_BYTES);
#pragma unroll
        for (int kk = 0; kk < ZG_BK / 4; ++kk) {
            double2 a[Z::TM], b[Z::TN];
#pragma unroll
            for (int i = 0; i < Z::TM; ++i) a[i] = TA ? as[(kk * 4 + t) * T::A_LD + row0 + i * 8 + g] : as[(row0 + i * 8 + g) * T::A_LD + kk * 4 + t];
#pragma unroll
            for (int j = 0; j < Z::TN; ++j) b[j] = TB ? bs[(col0 + j * 8 + g) * T::B_LD + kk * 4 + t] : bs[(kk * 4 + t) * T::B_LD + col0 + j * 8 + g];
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
        __syncthreads();                                    // every warp is done with this stage: it may be refilled
    }
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

namespace qocb_host {
// rank-3 map of a strided batch of row-major double2 matrices: [batch][rows][2 * cols doubles]; box = box_cols double2 x box_rows
inline bool make_zgemm_map(CUtensorMap *map, const double2 *base, int rows, int cols, int ld, long long batch_stride, int batch,
                           int box_cols, int box_rows) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn || !base || (reinterpret_cast<uintptr_t>(base) & 15) || rows <= 0 || cols <= 0) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)2 * cols, (cuuint64_t)rows, (cuuint64_t)std::max(batch, 1)};
    const cuuint64_t gstride[2] = {(cuuint64_t)ld * 16, (cuuint64_t)(batch > 1 ? batch_stride : (long long)rows * ld) * 16};
    const cuuint32_t box[3] = {(cuuint32_t)(2 * box_cols), (cuuint32_t)box_rows, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (box[0] > 256 || box[1] > 256 || gstride[1] == 0) return false;
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, const_cast<double2 *>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace qocb_host
