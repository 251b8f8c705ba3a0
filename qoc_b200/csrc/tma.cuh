// tma.cuh - Tensor Memory Accelerator feed for the planar matrices of the n <= 64 path (sm_100a).
//
// Every matrix of that path lives in global memory as two dense NP x NP planes of doubles (tile.cuh) and in shared memory
// as two planes with row stride LD = NP + 4 (the padding makes every DMMA fragment load bank-conflict free).  One 2-D TMA
// box per plane moves it: the tensor map describes a buffer of matrices as ONE tall 2-D array
//     [rows = matrices * 2 planes * NP][cols = NP]          (row pitch NP * 8 bytes)
// and the box is NP rows x (NP + 4) columns.  The four columns beyond the tensor's extent are out of bounds, which TMA
// zero-fills - so the box lands in shared memory with exactly the padded row stride, one `cp.async.bulk.tensor.2d` (SASS
// UTMALDG) per plane, completion signalled on an mbarrier (complete_tx::bytes counts the whole box, padding included).
// The copies are asynchronous: k_forward issues the next slice's Magnus matrix and the chunk propagator while the LU /
// substitution / squarings of the current slice run, and waits on the mbarrier only where the data is consumed.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace qocb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// orders the generic-proxy accesses made visible to this thread (e.g. by a preceding __syncthreads) before its subsequent
// async-proxy operations: needed before a TMA write into shared memory that other threads have just read or written
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
// all state spaces: a thread's generic-proxy writes to GLOBAL memory before a later TMA read of them
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.b32 %0, 1, 0, p;\n"
            "}\n" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
// one 2-D box: global coordinates (c0 = column, c1 = row) -> shared memory at dst; completion on bar
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n"
                 ::"r"(smem_u32(dst)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// Issue the two plane boxes of matrix `index` of the buffer described by `map` into the padded shared-memory matrix
// `dst` (row stride LD = NP + 4).  Called by ONE thread, after a barrier that ended every generic access to `dst`.
template <int NP>
__device__ __forceinline__ void tma_fetch_matrix(double *dst, const CUtensorMap *map, long long index, uint64_t *bar) {
    constexpr int LD = NP + 4, PLANE = NP * LD;
    constexpr uint32_t kBoxBytes = (uint32_t)(PLANE * sizeof(double));
    fence_proxy_async();
    mbar_expect_tx(bar, 2 * kBoxBytes);
    const long long row = index * 2 * NP;
    tma_load_2d(dst, map, 0, (int)row, bar);
    tma_load_2d(dst + PLANE, map, 0, (int)(row + NP), bar);
}

}  // namespace qocb

// ---- host: tensor maps through the driver entry point (no -lcuda link dependency) --------------------------------
namespace qocb_host {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

// tensor map of a buffer of `matrices` planar NP x NP matrices (see the header comment); false if TMA is unusable here
// (no driver entry point, a buffer that is not 16-byte aligned, more rows than a 32-bit coordinate holds)
inline bool make_matrix_map(CUtensorMap *map, const double *base, long long matrices, int NP) {
    EncodeTiledFn fn = encode_tiled_fn();
    const unsigned long long rows = (unsigned long long)matrices * 2ull * (unsigned long long)NP;
    if (!fn || !base || matrices <= 0 || (reinterpret_cast<uintptr_t>(base) & 15) || rows >= (1ull << 31)) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)NP, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)NP * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)(NP + 4), (cuuint32_t)NP};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double *>(base), gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

}  // namespace qocb_host
