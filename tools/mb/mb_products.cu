// micro-benchmark of the shared-memory complex products of tile.cuh (clock64 per product, one CTA per SM)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../qoc_b200/csrc/tile.cuh"
using namespace qocb;
using C = Cfg<64, 2, 4>;

// variant 2: per-tile A rows, no selects
template <class C>
__device__ __forceinline__ void mma_herm_v2(HAcc &acc, const double *__restrict__ A, const double *__restrict__ B) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const bool five = warp < 4;
    int ra[5], cb[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) { int rt, ct; herm_tile(warp, q, rt, ct); cb[q] = ct * 8 + g; ra[q] = (rt * 8 + g) * C::LD; }
#pragma unroll 2
    for (int kk = 0; kk < C::NP / 4; ++kk) {
        const int k = kk * 4 + t;
        double ar[5], ai[5], sa[5], br[5], bi[5], sb[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int idx = k * C::LD + cb[q];
            br[q] = B[idx]; bi[q] = B[C::PLANE + idx]; sb[q] = br[q] + bi[q];
            ar[q] = A[ra[q] + k]; ai[q] = A[C::PLANE + ra[q] + k]; sa[q] = ar[q] + ai[q];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma884(acc.v[q][0], acc.v[q][1], ar[q], br[q]);
        if (five) dmma884(acc.v[4][0], acc.v[4][1], ar[4], br[4]);
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma884(acc.v[q][2], acc.v[q][3], ai[q], bi[q]);
        if (five) dmma884(acc.v[4][2], acc.v[4][3], ai[4], bi[4]);
#pragma unroll
        for (int q = 0; q < 4; ++q) dmma884(acc.v[q][4], acc.v[q][5], sa[q], sb[q]);
        if (five) dmma884(acc.v[4][4], acc.v[4][5], sa[4], sb[4]);
    }
}
// variant 3: v2 with all five tiles unconditionally (4-tile warps repeat a tile) - no predicated DMMA
template <class C>
__device__ __forceinline__ void mma_herm_v3(HAcc &acc, const double *__restrict__ A, const double *__restrict__ B) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    int ra[5], cb[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) { int rt, ct; herm_tile(warp, q, rt, ct); cb[q] = ct * 8 + g; ra[q] = (rt * 8 + g) * C::LD; }
#pragma unroll 2
    for (int kk = 0; kk < C::NP / 4; ++kk) {
        const int k = kk * 4 + t;
        double ar[5], ai[5], sa[5], br[5], bi[5], sb[5];
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            const int idx = k * C::LD + cb[q];
            br[q] = B[idx]; bi[q] = B[C::PLANE + idx]; sb[q] = br[q] + bi[q];
            ar[q] = A[ra[q] + k]; ai[q] = A[C::PLANE + ra[q] + k]; sa[q] = ar[q] + ai[q];
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) dmma884(acc.v[q][0], acc.v[q][1], ar[q], br[q]);
#pragma unroll
        for (int q = 0; q < 5; ++q) dmma884(acc.v[q][2], acc.v[q][3], ai[q], bi[q]);
#pragma unroll
        for (int q = 0; q < 5; ++q) dmma884(acc.v[q][4], acc.v[q][5], sa[q], sb[q]);
    }
}
// variant 4: v3 with register double buffering of the fragments
template <class C>
__device__ __forceinline__ void mma_herm_v4(HAcc &acc, const double *__restrict__ A, const double *__restrict__ B) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    int ra[5], cb[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) { int rt, ct; herm_tile(warp, q, rt, ct); cb[q] = ct * 8 + g + t * C::LD; ra[q] = (rt * 8 + g) * C::LD + t; }
    double ar[5], ai[5], br[5], bi[5];
#pragma unroll
    for (int q = 0; q < 5; ++q) { br[q] = B[cb[q]]; bi[q] = B[C::PLANE + cb[q]]; ar[q] = A[ra[q]]; ai[q] = A[C::PLANE + ra[q]]; }
#pragma unroll 4
    for (int kk = 0; kk < C::NP / 4; ++kk) {
        double nar[5], nai[5], nbr[5], nbi[5];
        const int kn = (kk + 1 < C::NP / 4) ? kk + 1 : kk;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            nbr[q] = B[cb[q] + kn * 4 * C::LD]; nbi[q] = B[C::PLANE + cb[q] + kn * 4 * C::LD];
            nar[q] = A[ra[q] + kn * 4]; nai[q] = A[C::PLANE + ra[q] + kn * 4];
        }
#pragma unroll
        for (int q = 0; q < 5; ++q) dmma884(acc.v[q][0], acc.v[q][1], ar[q], br[q]);
#pragma unroll
        for (int q = 0; q < 5; ++q) dmma884(acc.v[q][2], acc.v[q][3], ai[q], bi[q]);
#pragma unroll
        for (int q = 0; q < 5; ++q) dmma884(acc.v[q][4], acc.v[q][5], ar[q] + ai[q], br[q] + bi[q]);
#pragma unroll
        for (int q = 0; q < 5; ++q) { ar[q] = nar[q]; ai[q] = nai[q]; br[q] = nbr[q]; bi[q] = nbi[q]; }
    }
}
// variant 5: DMMA only (pipe bound of 5 tiles x 3 per k-step)
__device__ __forceinline__ void mma_pipe5(HAcc &acc, double a, double b) {
#pragma unroll 4
    for (int kk = 0; kk < 16; ++kk) {
#pragma unroll
        for (int e = 0; e < 3; ++e)
#pragma unroll
            for (int q = 0; q < 5; ++q) dmma884(acc.v[q][2 * e], acc.v[q][2 * e + 1], a, b);
    }
}
// full product, double-buffered fragments
template <class C>
__device__ __forceinline__ void mma_smem_db(Acc<C> &acc, const double *__restrict__ A, const double *__restrict__ B) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int a0 = ((warp / C::WN) * C::TM * 8 + g) * C::LD + t;
    const int b0 = (warp % C::WN) * C::TN * 8 + g + t * C::LD;
    double ar[C::TM], ai[C::TM], br[C::TN], bi[C::TN];
#pragma unroll
    for (int i = 0; i < C::TM; ++i) { ar[i] = A[a0 + i * 8 * C::LD]; ai[i] = A[C::PLANE + a0 + i * 8 * C::LD]; }
#pragma unroll
    for (int j = 0; j < C::TN; ++j) { br[j] = B[b0 + j * 8]; bi[j] = B[C::PLANE + b0 + j * 8]; }
#pragma unroll 4
    for (int kk = 0; kk < C::NP / 4; ++kk) {
        const int kn = (kk + 1 < C::NP / 4) ? kk + 1 : kk;
        double nar[C::TM], nai[C::TM], nbr[C::TN], nbi[C::TN], sa[C::TM], sb[C::TN];
#pragma unroll
        for (int i = 0; i < C::TM; ++i) { nar[i] = A[a0 + i * 8 * C::LD + kn * 4]; nai[i] = A[C::PLANE + a0 + i * 8 * C::LD + kn * 4]; sa[i] = ar[i] + ai[i]; }
#pragma unroll
        for (int j = 0; j < C::TN; ++j) { nbr[j] = B[b0 + j * 8 + kn * 4 * C::LD]; nbi[j] = B[C::PLANE + b0 + j * 8 + kn * 4 * C::LD]; sb[j] = br[j] + bi[j]; }
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
            for (int j = 0; j < C::TN; ++j) dmma884(acc.v[i][j][0], acc.v[i][j][1], ar[i], br[j]);
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
            for (int j = 0; j < C::TN; ++j) dmma884(acc.v[i][j][2], acc.v[i][j][3], ai[i], bi[j]);
#pragma unroll
        for (int i = 0; i < C::TM; ++i)
#pragma unroll
            for (int j = 0; j < C::TN; ++j) dmma884(acc.v[i][j][4], acc.v[i][j][5], sa[i], sb[j]);
#pragma unroll
        for (int i = 0; i < C::TM; ++i) { ar[i] = nar[i]; ai[i] = nai[i]; }
#pragma unroll
        for (int j = 0; j < C::TN; ++j) { br[j] = nbr[j]; bi[j] = nbi[j]; }
    }
}

template <int V>
__global__ void __launch_bounds__(256, 1) k_bench(long long *out, double *sink, int reps) {
    extern __shared__ double smem[];
    double *X0 = smem, *X1 = smem + C::SMAT, *X2 = smem + 2 * C::SMAT;
    for (int i = threadIdx.x; i < 3 * C::SMAT; i += blockDim.x) smem[i] = 1e-3 * ((i * 7919) % 1013) - 0.5;
    __syncthreads();
    double s = 0.0;
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (V == 0 || V == 6) {
            Acc<C> acc; acc.zero();
            if (V == 0) mma_smem<C, false, false, false>(acc, X2, X0); else mma_smem_db<C>(acc, X2, X0);
            for_owned<C>([&](int i, int j, int, int) { s += accv<C>(acc, i, j).r0; });
        } else {
            HAcc h; h.zero();
            if (V == 1) mma_herm<C>(h, X2, X0);
            if (V == 2) mma_herm_v2<C>(h, X2, X0);
            if (V == 3) mma_herm_v3<C>(h, X2, X0);
            if (V == 4) mma_herm_v4<C>(h, X2, X0);
            if (V == 5) mma_pipe5(h, X2[threadIdx.x], X0[threadIdx.x]);
            if (V == 7) { mma_herm<C>(h, X2, X0); __syncthreads(); herm_store<C, false>(X1, h); }
            for (int q = 0; q < 5; ++q) s += h.v[q][0] + h.v[q][2] + h.v[q][4];
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[V] = (t1 - t0) / reps;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    long long *out; double *sink;
    cudaMalloc(&out, 64 * sizeof(long long)); cudaMalloc(&sink, 148 * 256 * sizeof(double));
    cudaMemset(out, 0, 64 * sizeof(long long));
    const int smem = 3 * C::SMAT * sizeof(double), reps = 400;
#define RUN(V) cudaFuncSetAttribute(k_bench<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); k_bench<V><<<148, 256, smem>>>(out, sink, 10); k_bench<V><<<148, 256, smem>>>(out, sink, reps);
    RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7)
    long long h[64];
    cudaError_t e = cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(e));
    const char *nm[] = {"full mma_smem", "herm (selects, predicated 5th)", "herm v2 (per-tile A rows)", "herm v3 (5 tiles unconditional)", "herm v4 (v3 + double buffer)", "pipe only 15 DMMA/k-step", "full double-buffered", "herm + sync + herm_store"};
    for (int v = 0; v < 8; ++v) printf("V%d %-34s %6lld clk  %.2f us\n", v, nm[v], h[v], h[v] / 1965.0);
    return 0;
}
