// sweep.cuh - state / costate propagation across time slices and the cost reductions.
//
// Reference path: the `for system_eval_step` loop of _evaluate_schroedinger_discrete
// (qoc/core/schroedingerdiscrete.py:393-425) and the state costs (qoc/standard/costs/
// targetstateinfidelity.py:39-63, targetstateinfidelitytime.py:46-73, forbidstates.py:50-81).
//
// The time loop is cut into chunks.  With the chunk propagators P_c = prod U_j (built by the expm kernel)
// the forward recursion psi_{j+1} = U_j psi_j becomes: (1) a short sequential pass over chunk boundaries,
// (2) independent local sweeps of all chunks.  The costate recursion lam_j = U_j^T lam_{j+1} + seed_j is
// affine: (1) local sweeps with zero incoming costate give the particular parts, (2) a sequential pass over
// chunk boundaries, (3) local sweeps with the true incoming costates.  The same three-step scheme is what
// the multi-GPU time-slice sharding uses across ranks.
//
// Vectors are planar (NP reals then NP imaginaries).  psi / lam arrays: [member][step][state][2][NP].
#pragma once
#include <cuda_runtime.h>

namespace qocb {

struct CostTerm {
    int kind;          // 0 coherent target, 1 incoherent target, 2 forbid
    int step;          // 1: evaluated at every cost step, 0: final step only
    int fmax;          // vectors per state (padded)
    int vec_off;       // offset (in vectors) into the plan's vector pool: index (s * fmax + f)
    int cnt_off;       // offset into the counts pool: F_s per state
    int ip_off;        // offset into the shared inner-product scratch
    double w;          // cost_multiplier / normalisation
};

struct SweepArgs {
    int NP, S, N, E;               // padded dim, states, system_eval_count, ensemble members
    int ces;                       // cost_eval_step
    int nterms, ip_total;
    const CostTerm *terms;
    const double *vecs;            // [..][2][NP]
    const int *counts;
    const double *U;               // [E*(N-1)][2*NP*NP]
    const double *chunkP;          // [nchunks][2*NP*NP]
    const int *chunk_begin;        // [nchunks+1] work-item index w = e*(N-1)+j; chunks never straddle members
    const int *member_chunk0;      // [E+1] first chunk of each member
    double *psi, *lam;             // [E][N][S][2][NP]
    double *part;                  // [nchunks][S][2][NP]
    double *cost_part;             // [nchunks]
    const double *psi_in;          // [S][2][NP] state entering the first local slice (psi0 unless time-sharded)
    const double *lam_in;          // [S][2][NP] costate entering the last local state from later shards, or nullptr
    double *b_out;                 // [S][2][NP] k_boundary_bwd also writes the costate at the first local state here
    int j_off, Nglob;              // time sharding: global index of local slice 0; global system_eval_count.
                                   // N above is the LOCAL state count (local slices + 1)
    int add_final_seed;            // this shard owns the final state (seed of the final-step costs)
};

constexpr int kSweepThreads = 256;

__device__ __forceinline__ void load_mat_sweep(double *sU, const double *gU, int NP) {
    const int LDS = NP + 1, PL = NP * LDS;
    for (int idx = threadIdx.x; idx < 2 * NP * NP; idx += kSweepThreads) {
        const int plane = idx / (NP * NP), rem = idx % (NP * NP);
        sU[plane * PL + (rem / NP) * LDS + (rem % NP)] = gU[idx];
    }
}

// out[s][a] = sum_b U[a][b] in[s][b]   (TRANS: sum_b U[b][a] in[s][b]);  sU padded NP+1, vectors in smem
template <bool TRANS>
__device__ __forceinline__ void matvec_smem(double *out, const double *in, const double *sU, int NP, int S) {
    const int LDS = NP + 1, PL = NP * LDS;
    for (int o = threadIdx.x; o < S * NP; o += kSweepThreads) {
        const int s = o / NP, a = o % NP;
        const double *vr = in + s * 2 * NP, *vi = vr + NP;
        double xr = 0., xi = 0., yr = 0., yi = 0.;
        for (int b = 0; b < NP; ++b) {
            const int idx = TRANS ? (b * LDS + a) : (a * LDS + b);
            const double ur = sU[idx], ui = sU[PL + idx];
            xr = fma(ur, vr[b], xr); yr = fma(ui, vi[b], yr);
            xi = fma(ur, vi[b], xi); yi = fma(ui, vr[b], yi);
        }
        out[s * 2 * NP + a] = xr - yr;
        out[s * 2 * NP + NP + a] = xi + yi;
    }
}

__device__ __forceinline__ bool is_step_cost_state(int k, int ces) { return k != 0 && (k % ces) == 0; }

// inner products <v_{t,s,f} | psi_s> for every active term -> ip[2*(ip_off + s*fmax + f)]; one warp per vector
__device__ void cost_inner_products(const SweepArgs &a, const double *psi, double *ip, bool step_state, bool final_state) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = kSweepThreads / 32;
    for (int t = 0; t < a.nterms; ++t) {
        const CostTerm tm = a.terms[t];
        if (!((tm.step && step_state) || (!tm.step && final_state))) continue;
        for (int e = warp; e < a.S * tm.fmax; e += nw) {
            const int s = e / tm.fmax, f = e % tm.fmax;
            double pr = 0., pi = 0.;
            if (f < a.counts[tm.cnt_off + s]) {
                const double *v = a.vecs + (size_t)(tm.vec_off + e) * 2 * a.NP;
                const double *p = psi + s * 2 * a.NP;
                for (int b = lane; b < a.NP; b += 32) {       // conj(v) * psi
                    const double vr = v[b], vi = v[a.NP + b], xr = p[b], xi = p[a.NP + b];
                    pr += vr * xr + vi * xi;
                    pi += vr * xi - vi * xr;
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { pr += __shfl_xor_sync(0xffffffffu, pr, o); pi += __shfl_xor_sync(0xffffffffu, pi, o); }
            if (lane == 0) { ip[2 * (tm.ip_off + e)] = pr; ip[2 * (tm.ip_off + e) + 1] = pi; }
        }
    }
    __syncthreads();
}

// value of the active cost terms at one state set (thread 0 returns it; other threads return 0)
__device__ double cost_value(const SweepArgs &a, const double *ip, bool step_state, bool final_state) {
    double val = 0.;
    if (threadIdx.x == 0) {
        for (int t = 0; t < a.nterms; ++t) {
            const CostTerm tm = a.terms[t];
            if (!((tm.step && step_state) || (!tm.step && final_state))) continue;
            if (tm.kind == 0) {
                double tr = 0., ti = 0.;
                for (int s = 0; s < a.S; ++s) { tr += ip[2 * (tm.ip_off + s * tm.fmax)]; ti += ip[2 * (tm.ip_off + s * tm.fmax) + 1]; }
                val += tm.w * (1.0 - (tr * tr + ti * ti) / ((double)a.S * a.S));
            } else if (tm.kind == 1) {
                double acc = 0.;
                for (int s = 0; s < a.S; ++s) {
                    const double r = ip[2 * (tm.ip_off + s * tm.fmax)], i = ip[2 * (tm.ip_off + s * tm.fmax) + 1];
                    acc += r * r + i * i;
                }
                val += tm.w * (1.0 - acc / a.S);
            } else {
                double acc = 0.;
                for (int s = 0; s < a.S; ++s) {
                    const int F = a.counts[tm.cnt_off + s];
                    double sub = 0.;
                    for (int f = 0; f < F; ++f) {
                        const double r = ip[2 * (tm.ip_off + s * tm.fmax + f)], i = ip[2 * (tm.ip_off + s * tm.fmax + f) + 1];
                        sub += r * r + i * i;
                    }
                    acc += sub / F;
                }
                val += tm.w * acc;
            }
        }
    }
    return val;
}

// lam[s][a] += d cost / d psi_s[a] in autograd's convention (d/dx - i d/dy):  sum coef * conj(v[a])
__device__ void cost_add_seed(const SweepArgs &a, const double *ip, double *lam, bool step_state, bool final_state) {
    for (int t = 0; t < a.nterms; ++t) {
        const CostTerm tm = a.terms[t];
        if (!((tm.step && step_state) || (!tm.step && final_state))) continue;
        double tr = 0., ti = 0.;
        if (tm.kind == 0)
            for (int s = 0; s < a.S; ++s) { tr += ip[2 * (tm.ip_off + s * tm.fmax)]; ti += ip[2 * (tm.ip_off + s * tm.fmax) + 1]; }
        for (int o = threadIdx.x; o < a.S * a.NP; o += kSweepThreads) {
            const int s = o / a.NP, b = o % a.NP;
            const int F = a.counts[tm.cnt_off + s];
            double sr = 0., si = 0.;
            for (int f = 0; f < F; ++f) {
                const double *v = a.vecs + (size_t)(tm.vec_off + s * tm.fmax + f) * 2 * a.NP;
                double cr, ci;         // coefficient = scale * conj(ip or total)
                if (tm.kind == 0) { const double k = -2.0 * tm.w / ((double)a.S * a.S); cr = k * tr; ci = -k * ti; }
                else if (tm.kind == 1) { const double k = -2.0 * tm.w / a.S; cr = k * ip[2 * (tm.ip_off + s * tm.fmax)]; ci = -k * ip[2 * (tm.ip_off + s * tm.fmax) + 1]; }
                else { const double k = 2.0 * tm.w / F; cr = k * ip[2 * (tm.ip_off + s * tm.fmax + f)]; ci = -k * ip[2 * (tm.ip_off + s * tm.fmax + f) + 1]; }
                const double vr = v[b], vi = -v[a.NP + b];              // conj(v)
                sr += cr * vr - ci * vi;
                si += cr * vi + ci * vr;
            }
            lam[s * 2 * a.NP + b] += sr;
            lam[s * 2 * a.NP + a.NP + b] += si;
        }
    }
    __syncthreads();
}

// dynamic smem layout of the sweep kernels: U (2*NP*(NP+1)) | v0 (S*2*NP) | v1 (S*2*NP) | ip (2*ip_total)
__host__ __device__ inline size_t sweep_smem_bytes(int NP, int S, int ip_total) {
    return sizeof(double) * ((size_t)2 * NP * (NP + 1) + (size_t)4 * S * NP + (size_t)2 * (ip_total > 0 ? ip_total : 1));
}

// (1) boundary states: psi[b_{c+1}] = P_c psi[b_c], sequential over the chunks of one member; grid = E
__global__ void __launch_bounds__(kSweepThreads) k_boundary_fwd(SweepArgs a) {
    extern __shared__ double sm[];
    const int NP = a.NP, S = a.S, VS = S * 2 * NP;
    double *sU = sm, *v0 = sm + 2 * NP * (NP + 1), *v1 = v0 + VS;
    const int e = blockIdx.x;
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) v0[i] = a.psi_in[i];
    __syncthreads();
    double *psi_e = a.psi + (size_t)e * a.N * VS;
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) psi_e[i] = v0[i];
    for (int c = a.member_chunk0[e]; c < a.member_chunk0[e + 1]; ++c) {
        load_mat_sweep(sU, a.chunkP + (size_t)c * 2 * NP * NP, NP);
        __syncthreads();
        matvec_smem<false>(v1, v0, sU, NP, S);
        __syncthreads();
        const int kend = a.chunk_begin[c + 1] - e * (a.N - 1);      // state index at the end of chunk c
        for (int i = threadIdx.x; i < VS; i += kSweepThreads) psi_e[(size_t)kend * VS + i] = v1[i];
        double *t = v0; v0 = v1; v1 = t;
        __syncthreads();
    }
}

// (2) local forward sweeps + cost values; grid = nchunks
__global__ void __launch_bounds__(kSweepThreads) k_sweep_fwd(SweepArgs a) {
    extern __shared__ double sm[];
    const int NP = a.NP, S = a.S, VS = S * 2 * NP;
    double *sU = sm, *v0 = sm + 2 * NP * (NP + 1), *v1 = v0 + VS, *ip = v1 + VS;
    const int c = blockIdx.x;
    const int wb = a.chunk_begin[c], we = a.chunk_begin[c + 1];
    const int e = wb / (a.N - 1), jb = wb - e * (a.N - 1), je = we - e * (a.N - 1);
    double *psi_e = a.psi + (size_t)e * a.N * VS;
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) v0[i] = psi_e[(size_t)jb * VS + i];
    __syncthreads();
    double cost = 0.;
    for (int j = jb; j < je; ++j) {
        const int k = j + 1;                                          // state produced by slice j
        if (k < je) {
            load_mat_sweep(sU, a.U + (size_t)(e * (a.N - 1) + j) * 2 * NP * NP, NP);
            __syncthreads();
            matvec_smem<false>(v1, v0, sU, NP, S);
            __syncthreads();
            for (int i = threadIdx.x; i < VS; i += kSweepThreads) psi_e[(size_t)k * VS + i] = v1[i];
        } else {
            for (int i = threadIdx.x; i < VS; i += kSweepThreads) v1[i] = psi_e[(size_t)k * VS + i];   // boundary state
            __syncthreads();
        }
        const bool st = is_step_cost_state(k + a.j_off, a.ces), fin = (k + a.j_off == a.Nglob - 1);
        if (a.nterms > 0 && (st || fin)) {
            cost_inner_products(a, v1, ip, st, fin);
            cost += cost_value(a, ip, st, fin);
        }
        double *t = v0; v0 = v1; v1 = t;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.cost_part[c] = cost;
}

// (3a/3c) local backward sweeps.  PARTICULAR: zero incoming costate, result -> part[c], nothing stored.
// otherwise: incoming lam[je] read from the lam array (written by k_boundary_bwd), lam[j] stored for j in (jb, je).
template <bool PARTICULAR>
__global__ void __launch_bounds__(kSweepThreads) k_sweep_bwd(SweepArgs a) {
    extern __shared__ double sm[];
    const int NP = a.NP, S = a.S, VS = S * 2 * NP;
    double *sU = sm, *v0 = sm + 2 * NP * (NP + 1), *v1 = v0 + VS, *ip = v1 + VS;
    const int c = blockIdx.x;
    const int wb = a.chunk_begin[c], we = a.chunk_begin[c + 1];
    const int e = wb / (a.N - 1), jb = wb - e * (a.N - 1), je = we - e * (a.N - 1);
    const double *psi_e = a.psi + (size_t)e * a.N * VS;
    double *lam_e = a.lam + (size_t)e * a.N * VS;
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) v0[i] = PARTICULAR ? 0. : lam_e[(size_t)je * VS + i];
    __syncthreads();
    const int jstop = PARTICULAR ? jb : jb + 1;
    for (int j = je - 1; j >= jstop; --j) {
        load_mat_sweep(sU, a.U + (size_t)(e * (a.N - 1) + j) * 2 * NP * NP, NP);
        __syncthreads();
        matvec_smem<true>(v1, v0, sU, NP, S);                         // lam_j = U_j^T lam_{j+1}
        __syncthreads();
        const bool st = is_step_cost_state(j + a.j_off, a.ces);
        if (a.nterms > 0 && st) {                                     // + seed_j (state j < N-1: step costs only)
            cost_inner_products(a, psi_e + (size_t)j * VS, ip, true, false);
            cost_add_seed(a, ip, v1, true, false);
        }
        if (!PARTICULAR)
            for (int i = threadIdx.x; i < VS; i += kSweepThreads) lam_e[(size_t)j * VS + i] = v1[i];
        double *t = v0; v0 = v1; v1 = t;
        __syncthreads();
    }
    if (PARTICULAR)
        for (int i = threadIdx.x; i < VS; i += kSweepThreads) a.part[(size_t)c * VS + i] = v0[i];
}

// (3b) boundary costates, sequential over the chunks of one member (last to first); grid = E.
// lam[N-1] = seed_{N-1};  lam[b_c] = P_c^T lam[b_{c+1}] + part_c
__global__ void __launch_bounds__(kSweepThreads) k_boundary_bwd(SweepArgs a, int have_part) {
    extern __shared__ double sm[];
    const int NP = a.NP, S = a.S, VS = S * 2 * NP;
    double *sU = sm, *v0 = sm + 2 * NP * (NP + 1), *v1 = v0 + VS, *ip = v1 + VS;
    const int e = blockIdx.x;
    const double *psi_e = a.psi + (size_t)e * a.N * VS;
    double *lam_e = a.lam + (size_t)e * a.N * VS;
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) v0[i] = a.lam_in ? a.lam_in[i] : 0.;
    __syncthreads();
    if (a.nterms > 0 && a.add_final_seed) {
        const bool st = is_step_cost_state(a.N - 1 + a.j_off, a.ces);
        cost_inner_products(a, psi_e + (size_t)(a.N - 1) * VS, ip, st, true);
        cost_add_seed(a, ip, v0, st, true);
    }
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) lam_e[(size_t)(a.N - 1) * VS + i] = v0[i];
    for (int c = a.member_chunk0[e + 1] - 1; c >= a.member_chunk0[e]; --c) {
        load_mat_sweep(sU, a.chunkP + (size_t)c * 2 * NP * NP, NP);
        __syncthreads();
        matvec_smem<true>(v1, v0, sU, NP, S);
        __syncthreads();
        if (have_part)
            for (int i = threadIdx.x; i < VS; i += kSweepThreads) v1[i] += a.part[(size_t)c * VS + i];
        const int kbeg = a.chunk_begin[c] - e * (a.N - 1);
        for (int i = threadIdx.x; i < VS; i += kSweepThreads) lam_e[(size_t)kbeg * VS + i] = v1[i];
        double *t = v0; v0 = v1; v1 = t;
        __syncthreads();
    }
    if (a.b_out && e == 0)
        for (int i = threadIdx.x; i < VS; i += kSweepThreads) a.b_out[i] = v0[i];
}

// time sharding: state entering shard `rank` = P_{rank-1} ... P_0 psi0 (allP: [world][2*NP*NP]); grid = 1
__global__ void __launch_bounds__(kSweepThreads) k_prefix_states(const double *allP, const double *psi0, double *psi_in,
                                                                 int rank, int NP, int S) {
    extern __shared__ double sm[];
    const int VS = S * 2 * NP;
    double *sU = sm, *v0 = sm + 2 * NP * (NP + 1), *v1 = v0 + VS;
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) v0[i] = psi0[i];
    __syncthreads();
    for (int r = 0; r < rank; ++r) {
        load_mat_sweep(sU, allP + (size_t)r * 2 * NP * NP, NP);
        __syncthreads();
        matvec_smem<false>(v1, v0, sU, NP, S);
        __syncthreads();
        double *t = v0; v0 = v1; v1 = t;
    }
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) psi_in[i] = v0[i];
}

// time sharding: costate entering shard `rank` from the later shards: lam_in(world-1) = 0,
// lam_in(g) = P_{g+1}^T lam_in(g+1) + b_{g+1}   (allb: [world][S][2][NP]); grid = 1
__global__ void __launch_bounds__(kSweepThreads) k_suffix_costates(const double *allP, const double *allb, double *lam_in,
                                                                   int rank, int world, int NP, int S) {
    extern __shared__ double sm[];
    const int VS = S * 2 * NP;
    double *sU = sm, *v0 = sm + 2 * NP * (NP + 1), *v1 = v0 + VS;
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) v0[i] = 0.;
    __syncthreads();
    for (int r = world - 1; r > rank; --r) {
        load_mat_sweep(sU, allP + (size_t)r * 2 * NP * NP, NP);
        __syncthreads();
        matvec_smem<true>(v1, v0, sU, NP, S);
        __syncthreads();
        for (int i = threadIdx.x; i < VS; i += kSweepThreads) v1[i] += allb[(size_t)r * VS + i];
        __syncthreads();
        double *t = v0; v0 = v1; v1 = t;
    }
    for (int i = threadIdx.x; i < VS; i += kSweepThreads) lam_in[i] = v0[i];
}

}  // namespace qocb
