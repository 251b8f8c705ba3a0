// microbench.cu - FP64 pipe probes on B200 (sm_100a): DFMA, DMMA m8n8k4 / m16n8k8 / m16n8k16
// throughput and dependent-issue latency, plus cuBLAS DGEMM / ZGEMM sustained rates.
// These numbers are the FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 figure).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/microbench tools/microbench.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cuComplex.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double *d, const double *a, const double *b) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double *d, const double *a, const double *b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(d[0]), "+d"(d[1]), "+d"(d[2]), "+d"(d[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int ILP>
__global__ void k_dfma(double *out, int iters) {
    double acc[ILP];
    double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dmma884(double *out, int iters) {
    double d0[ILP], d1[ILP];
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { d0[i] = i; d1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(d0[i], d1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += d0[i] + d1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dmma1688(double *out, int iters) {
    double d[ILP][4];
    double a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    b[0] = 1e-3; b[1] = 2e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) d[i][j] = i + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma1688(d[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_dmma16816(double *out, int iters) {
    double d[ILP][4];
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (i + 1);
#pragma unroll
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) d[i][j] = i + j;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma16816(d[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) for (int j = 0; j < 4; ++j) s += d[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_khz\": %d}\n", p.name, sms, p.major, p.minor, p.clockRate);
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    const int iters = 20000;
    // throughput: many warps, ILP 8; latency: 1 warp / SM-subpartition, ILP 1
    for (int warps_per_cta : {4, 8, 16, 32}) {
        int thr = warps_per_cta * 32;
        float ms = time_ms([&] { k_dfma<8><<<sms * 2, thr>>>(out, iters); });
        double fl = 2.0 * 8 * iters * (double)thr * sms * 2;
        printf("{\"probe\": \"dfma\", \"warps_per_cta\": %d, \"ctas_per_sm\": 2, \"tflops\": %.2f}\n", warps_per_cta, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma884<8><<<sms * 2, thr>>>(out, iters); });
        fl = 2.0 * 256 * 8 * iters * (double)warps_per_cta * sms * 2;
        printf("{\"probe\": \"dmma_m8n8k4\", \"warps_per_cta\": %d, \"ctas_per_sm\": 2, \"tflops\": %.2f}\n", warps_per_cta, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma1688<8><<<sms * 2, thr>>>(out, iters); });
        fl = 2.0 * 1024 * 8 * iters * (double)warps_per_cta * sms * 2;
        printf("{\"probe\": \"dmma_m16n8k8\", \"warps_per_cta\": %d, \"ctas_per_sm\": 2, \"tflops\": %.2f}\n", warps_per_cta, fl / ms * 1e-9);
        ms = time_ms([&] { k_dmma16816<8><<<sms * 2, thr>>>(out, iters); });
        fl = 2.0 * 2048 * 8 * iters * (double)warps_per_cta * sms * 2;
        printf("{\"probe\": \"dmma_m16n8k16\", \"warps_per_cta\": %d, \"ctas_per_sm\": 2, \"tflops\": %.2f}\n", warps_per_cta, fl / ms * 1e-9);
    }
    {   // dependent-issue latency in ns per instruction (single warp per CTA, ILP 1)
        float ms = time_ms([&] { k_dfma<1><<<sms, 32>>>(out, iters); });
        printf("{\"probe\": \"dfma_dep_latency_ns\", \"value\": %.2f}\n", ms * 1e6 / iters);
        ms = time_ms([&] { k_dmma884<1><<<sms, 32>>>(out, iters); });
        printf("{\"probe\": \"dmma884_dep_latency_ns\", \"value\": %.2f}\n", ms * 1e6 / iters);
        ms = time_ms([&] { k_dmma1688<1><<<sms, 32>>>(out, iters); });
        printf("{\"probe\": \"dmma1688_dep_latency_ns\", \"value\": %.2f}\n", ms * 1e6 / iters);
        ms = time_ms([&] { k_dmma16816<1><<<sms, 32>>>(out, iters); });
        printf("{\"probe\": \"dmma16816_dep_latency_ns\", \"value\": %.2f}\n", ms * 1e6 / iters);
        // single warp per SMSP, ILP 8 (issue-limited rate of one warp)
        ms = time_ms([&] { k_dmma884<8><<<sms, 128>>>(out, iters); });
        printf("{\"probe\": \"dmma884_4warps_ilp8_ns_per_instr\", \"value\": %.3f}\n", ms * 1e6 / iters / 8);
        ms = time_ms([&] { k_dmma884<4><<<sms, 256>>>(out, iters); });
        printf("{\"probe\": \"dmma884_8warps_ilp4_ns_per_instr_per_warp\", \"value\": %.3f}\n", ms * 1e6 / iters / 4);
    }
    // cuBLAS DGEMM / ZGEMM
    cublasHandle_t h; cublasCreate(&h);
    {
        int n = 8192;
        double *A, *B, *C; CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&C, sizeof(double) * n * n));
        CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
        double al = 1, be = 0;
        float ms = time_ms([&] { cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 3);
        printf("{\"probe\": \"cublas_dgemm_8192\", \"tflops\": %.2f}\n", 2.0 * n * n * (double)n / ms * 1e-9);
        // sustained ~3 s
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        int reps = (int)(3000.0f / ms) + 1;
        cudaEventRecord(e0);
        for (int r = 0; r < reps; ++r) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float tot; cudaEventElapsedTime(&tot, e0, e1);
        printf("{\"probe\": \"cublas_dgemm_8192_sustained\", \"tflops\": %.2f, \"seconds\": %.2f}\n", 2.0 * n * n * (double)n * reps / tot * 1e-9, tot * 1e-3);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    {
        int n = 4096;
        cuDoubleComplex *A, *B, *C; CK(cudaMalloc(&A, 16ull * n * n)); CK(cudaMalloc(&B, 16ull * n * n)); CK(cudaMalloc(&C, 16ull * n * n));
        CK(cudaMemset(A, 0, 16ull * n * n)); CK(cudaMemset(B, 0, 16ull * n * n));
        cuDoubleComplex al = make_cuDoubleComplex(1, 0), be = make_cuDoubleComplex(0, 0);
        float ms = time_ms([&] { cublasZgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, B, n, &be, C, n); }, 3);
        printf("{\"probe\": \"cublas_zgemm_4096\", \"tflops\": %.2f}\n", 8.0 * n * n * (double)n / ms * 1e-9);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    for (int n : {32, 64, 128, 256}) {
        int batch = n <= 64 ? 4096 : (n == 128 ? 1024 : 256);
        cuDoubleComplex *A, *B, *C; size_t sz = 16ull * n * n * batch;
        CK(cudaMalloc(&A, sz)); CK(cudaMalloc(&B, sz)); CK(cudaMalloc(&C, sz));
        CK(cudaMemset(A, 0, sz)); CK(cudaMemset(B, 0, sz));
        cuDoubleComplex al = make_cuDoubleComplex(1, 0), be = make_cuDoubleComplex(0, 0);
        float ms = time_ms([&] { cublasZgemmStridedBatched(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &al, A, n, (long long)n * n, B, n, (long long)n * n, &be, C, n, (long long)n * n, batch); }, 5);
        printf("{\"probe\": \"cublas_zgemm_strided_batched\", \"n\": %d, \"batch\": %d, \"tflops\": %.2f, \"us_per_matrix\": %.3f}\n", n, batch, 8.0 * n * n * (double)n * batch / ms * 1e-9, ms * 1e3 / batch);
        cudaFree(A); cudaFree(B); cudaFree(C);
    }
    // plain device copy bandwidth (sanity vs MEASURED_PEAKS.json)
    {
        size_t bytes = 2ull << 30; char *a, *b; CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes));
        float ms = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 5);
        printf("{\"probe\": \"d2d_copy\", \"gbs\": %.1f}\n", 2.0 * bytes / ms * 1e-6);
        cudaFree(a); cudaFree(b);
    }
    printf("{\"done\": true}\n");
    return 0;
}
