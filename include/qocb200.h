/*
 * qocb200.h - C ABI of the B200-native GRAPE propagate-and-differentiate hot path.
 *
 * The reference (SchusterLab/qoc) is pure Python and has no FFI; the seam this library replaces is
 *   (1) _evaluate_schroedinger_discrete(controls, pstate, reporter) -> error
 *       (qoc/core/schroedingerdiscrete.py:356-438, also used by evolve_schroedinger_discrete, :101), and
 *   (2) ans_jacobian(_evaluate_schroedinger_discrete, 0)(controls, pstate, reporter) -> (error, grads)
 *       (qoc/core/schroedingerdiscrete.py:318, qoc/standard/utils/autogradutil.py:10-31),
 * and the Lindblad twins (qoc/core/lindbladdiscrete.py:357-441, :322).  INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add at those two call sites.
 *
 * Conventions
 *   - every pointer is a HOST pointer unless the name ends in `_dev`; the caller owns all host buffers;
 *   - complex arrays are interleaved (re, im) doubles in NumPy C order (complex128);
 *   - controls are REAL channels: real controls -> KR = K; complex controls -> KR = 2K with
 *     x = [Re u_0..Re u_{K-1}, Im u_0..Im u_{K-1}] per control step, and the Hamiltonian is the real-linear form
 *     H(x) = H0 + sum_r x_r A_r (the Python layer extracts H0, A_r from the user's callable);
 *   - gradients are dE/dx_r (real).  For complex controls dE/dRe u + i dE/dIm u is what qoc's optimiser
 *     receives after the wrapper's conjugate (qoc/core/schroedingerdiscrete.py:320-324);
 *   - every function returns 0 on success; on failure a negative code, message via qocb_last_error();
 *   - a plan is not re-entrant; distinct plans are independent; calls block until outputs are on the host
 *     unless stated otherwise.  The plan owns all device memory and its stream.
 */
#ifndef QOCB200_H
#define QOCB200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qocb_plan qocb_plan;

typedef struct {
    int32_t hilbert_size;        /* n */
    int32_t state_count;         /* S */
    int32_t control_count;       /* KR: real control channels */
    int32_t control_eval_count;  /* M  (qoc/models/programstate.py:40-41) */
    int32_t system_eval_count;   /* N  (N-1 time slices, dt = T/(N-1), programstate.py:44) */
    int32_t magnus_order;        /* 2, 4 or 6  (qoc/models/magnuspolicy.py:8-26) */
    int32_t cost_eval_step;      /* >= 1 */
    int32_t ensemble_count;      /* E >= 1: members differ in the drift H0 only; cost = mean over members */
    int32_t device;              /* CUDA device ordinal */
    int32_t store_tape;          /* 1: keep Pade intermediates of the forward pass in HBM for the reverse pass;
                                    0: recompute them in the reverse pass (less memory, ~25% more flops) */
    int32_t chunks_per_member;   /* 0 = automatic */
    int32_t slice_begin;         /* time-slice sharding: this plan owns slices [slice_begin, slice_end) of the N-1 */
    int32_t slice_end;           /* (both 0 = the whole pulse, unsharded).  See the qocb_shard_* calls below */
    int32_t channel_count;       /* KC: operator channels of a time-dependent hamiltonian (qocb_set_node_map);
                                    0 = control_count (time-independent operators, one per real control channel) */
    int32_t state_total;         /* state sharding (qocb_state_shard_*): total number of initial states over all ranks; */
    int32_t state_first;         /* index of this plan's first state among them.  0 / 0 = all states are here */
    double evolution_time;       /* T */
} qocb_problem;

/* cost kinds (qoc/standard/costs): vectors are `state_count x fmax` column vectors of length n */
#define QOCB_COST_TARGET_COHERENT 0   /* w * (1 - |sum_s <t_s|psi_s>|^2 / S^2)   targetstateinfidelity.py:52-56 */
#define QOCB_COST_TARGET_INCOHERENT 1 /* w * (1 - sum_s |<t_s|psi_s>|^2 / S)     targetstateinfidelity.py:57-61 */
#define QOCB_COST_FORBID 2            /* w * sum_s (1/F_s) sum_f |<f_sf|psi_s>|^2  forbidstates.py:64-81         */

int qocb_plan_create(const qocb_problem *problem, qocb_plan **plan_out);
int qocb_plan_destroy(qocb_plan *plan);
const char *qocb_last_error(const qocb_plan *plan);           /* plan may be NULL: last creation error */

/* H0: [E][n][n] complex (E = ensemble_count), A: [KC][n][n] complex (KC = channel_count, or control_count if that is 0) */
int qocb_set_operators(qocb_plan *plan, const double *h0, const double *a_ops);
/* Time-dependent hamiltonian(controls, time) (the callable's `time` argument, qoc/core/schroedingerdiscrete.py:483-486),
   affine in the controls: at Magnus node i of slice j
       H = H0 + sum_c coef[j][i][c] A_c,   coef[j][i][c] = offset[j][i][c] + sum_r gain[j][i][c][r] x_r(t_ji),
   x_r = interpolated real control channel r.  offset: [N-1][q][KC], gain: [N-1][q][KC][KR] for ALL N-1 slices (a plan
   that owns a slice range reads its part); q = magnus_order / 2.  Requires channel_count > 0.  gain may be NULL when
   control_count is 0.  The gradient returned by qocb_cost_and_grad stays [M][KR]. */
int qocb_set_node_map(qocb_plan *plan, const double *offset, const double *gain);
/* psi0: [S][n] complex */
int qocb_set_states(qocb_plan *plan, const double *psi0);
/* vectors: [S][fmax][n] complex; counts: [S] number of valid vectors per state (NULL = fmax for all);
   weight = cost_multiplier / normalisation; step_cost != 0: evaluated at every cost step, else final step only */
int qocb_add_cost(qocb_plan *plan, int32_t kind, int32_t step_cost, double weight,
                  const double *vectors, const int32_t *counts, int32_t fmax);
int qocb_clear_costs(qocb_plan *plan);

/* controls: [M][KR] real.  cost: 1 double.  final_states: [E][S][n] complex or NULL. */
int qocb_cost(qocb_plan *plan, const double *controls, double *cost, double *final_states);
/* grad: [M][KR] real */
int qocb_cost_and_grad(qocb_plan *plan, const double *controls, double *cost, double *grad, double *final_states);
/* The same evaluation in two calls, for cost terms the host must evaluate itself (user-defined Cost subclasses,
   qoc/models/cost.py:5-51): qocb_forward runs the forward pass (keeping the reverse-pass tape) and returns the device costs'
   value and the final states; the host adds its own terms c(psi_N) and hands their cotangent to qocb_backward as
   final_seed[S][n] complex = d c / d Re psi - i d c / d Im psi (autograd's convention), or NULL; qocb_backward returns the
   gradient of (device costs + host terms) with respect to the controls.  Unsharded single-member plans only. */
int qocb_forward(qocb_plan *plan, const double *controls, double *cost, double *final_states);
int qocb_backward(qocb_plan *plan, const double *final_seed, double *grad);
/* all states of the last evaluation: [E][N][S][n] complex (the reference's save_intermediate_states payload,
   qoc/core/schroedingerdiscrete.py:395-402) */
int qocb_get_states(qocb_plan *plan, double *states);
/* final states of the last evaluation: [E][S][n] complex (reporter.final_states, qoc/core/schroedingerdiscrete.py:436);
   waits for the plan stream, so it also serves evaluations enqueued with qocb_run_resident */
int qocb_get_final_states(qocb_plan *plan, double *final_states);
/* per-node cotangents of the operator-channel coefficients of the last qocb_cost_and_grad: [E][N-1][q][KC], dE / d coef[j][i][c]
   (KC = channel_count, or control_count when that is 0).  With control_count = 0 and the coefficients uploaded as the offset
   table of qocb_set_node_map this is the device half of the chain rule through a hamiltonian(controls, time) callable that
   is not affine in the controls (qoc_b200/core/plan.py: NonlinearSchroedingerPlan). */
int qocb_get_node_grad(qocb_plan *plan, double *node_grad);
/* slice propagators U_j of the last evaluation: [E][N-1][n][n] complex */
int qocb_get_propagators(qocb_plan *plan, double *props);

/* device-resident pipeline for benchmarking: controls already in HBM, no host transfer, no sync.
   qocb_upload_controls stages them; qocb_run_resident enqueues the whole evaluation on the plan stream;
   qocb_sync waits; qocb_download_result copies cost/grad to the host. */
int qocb_upload_controls(qocb_plan *plan, const double *controls);
int qocb_run_resident(qocb_plan *plan, int32_t with_grad);
int qocb_sync(qocb_plan *plan);
int qocb_download_result(qocb_plan *plan, double *cost, double *grad);
/* time `iters` resident evaluations with CUDA events on the plan stream after `warmup` untimed ones.
   ms_total[0] = total ms; kernel_ms[8] = summed ms of each pipeline stage (expm forward = Magnus pass + k_forward, boundary
   fwd, sweep fwd, sweep bwd (all three), expm reverse (k_backward + Magnus adjoint pass), gather, finalize, propagator tree
   between the expm forward and the boundary pass); flush_l2 != 0 writes a 256 MiB buffer
   between evaluations (outside the stage timers, inside ms_total only if count_flush != 0). */
int qocb_time_resident(qocb_plan *plan, int32_t with_grad, int32_t warmup, int32_t iters, int32_t flush_l2,
                       double *ms_total, double *stage_ms);
/* enqueue a 256 MiB write on the plan stream (evicts the 126 MB L2 between timed iterations) */
int qocb_flush_l2(qocb_plan *plan);
/* number of kernel launches of one evaluation */
int qocb_launch_count(qocb_plan *plan, int32_t with_grad);
void *qocb_stream(qocb_plan *plan);                           /* cudaStream_t of the plan */

/* ---- time-slice sharding (one plan per rank, created with a slice range) ---------------------
   The four phase calls only ENQUEUE work on the plan stream.  `*_dev` pointers are DEVICE buffers owned by the
   caller: the host language allocates them and runs the collectives between the phases on the same stream
   (qoc_b200/core/sharded.py uses torch.distributed / NCCL over NVLink):
     1. qocb_shard_forward_local(plan, with_grad, shardP_dev)   Magnus + expm of the local slices; product of the
        local propagators -> shardP_dev[qocb_shard_matrix_doubles]
        -- all-gather shardP_dev -> allP_dev[world][matrix_doubles]
     2. qocb_shard_forward_finish(plan, allP_dev, rank)         incoming boundary state P_{rank-1}..P_0 psi0, local
        state sweeps, cost partial
     3. qocb_shard_backward_particular(plan, b_dev)             costate at the shard's first state for a zero incoming
        costate -> b_dev[qocb_shard_vector_doubles]             (the affine recursion's particular part)
        -- all-gather b_dev -> allb_dev[world][vector_doubles]
     4. qocb_shard_backward_finish(plan, allP_dev, allb_dev, rank, world)
        incoming costate by suffix combination, local costate sweeps, expm/Magnus adjoints, gradient scatter
     5. qocb_shard_pack_result(plan, with_grad, result_dev)     result_dev[qocb_shard_result_doubles] =
        [partial gradient M*KR | partial cost | final states S*2*NP planar, zeros unless this shard owns the last
        slice] -- all-reduce(sum) it.  A forward-only evaluation runs phases 1, 2, 5 (with_grad = 0). */
int qocb_shard_matrix_doubles(qocb_plan *plan);
int qocb_shard_vector_doubles(qocb_plan *plan);
int qocb_shard_forward_local(qocb_plan *plan, int32_t with_grad, double *shardP_dev);
int qocb_shard_forward_finish(qocb_plan *plan, const double *allP_dev, int32_t rank);
int qocb_shard_backward_particular(qocb_plan *plan, double *b_dev);
int qocb_shard_backward_finish(qocb_plan *plan, const double *allP_dev, const double *allb_dev, int32_t rank,
                               int32_t world);
int qocb_shard_result_doubles(qocb_plan *plan);
int qocb_shard_pack_result(qocb_plan *plan, int32_t with_grad, double *result_dev);

/* ---- sharding of the initial STATES (independent units; one plan per rank, created with state_count = the local
   states, state_total / state_first set, and the cost vectors of the local states only).  Every rank computes every slice
   propagator - they depend on the controls alone - and sweeps its own states; cost normalisations use state_total.  The only
   coupling is the coherent target infidelity 1 - |sum_s <t_s|psi_s>|^2 / S^2 (targetstateinfidelity.py:53-55):
     1. qocb_state_shard_forward(plan, with_grad, coh_dev)   expm + state sweeps; coh_dev[qocb_state_shard_coherent_doubles] =
        this rank's partial overlap sums (re, im per coherent term and cost step); the cost so far excludes those terms
        -- all-reduce(sum) coh_dev (skip when the count is 0)
     2. qocb_state_shard_finish(plan, with_grad, coh_dev)    coherent values from the totals (added once, by the rank with
        state_first = 0), then - with_grad - costate sweeps seeded from the totals, expm / Magnus adjoints, gradient
     3. qocb_shard_pack_result(plan, with_grad, result_dev)  [gradient | cost | ...] -- all-reduce(sum) the first M*KR + 1 doubles.
   The calls only enqueue on the plan stream.  Not combinable with time-slice sharding or ensembles. */
int qocb_state_shard_coherent_doubles(qocb_plan *plan);
int qocb_state_shard_forward(qocb_plan *plan, int32_t with_grad, double *coh_dev);
int qocb_state_shard_finish(qocb_plan *plan, int32_t with_grad, const double *coh_dev);

/* standalone batched matrix exponential (qoc/standard/functions/expm.py:210-252), bench / test hook.
   a, out: [batch][n][n] complex on the host. */
int qocb_expm_batched(int32_t n, int64_t batch, const double *a, double *out, int32_t device);
/* same with adjoint: given ubar [batch][n][n] (cotangent of the output, autograd convention) returns
   abar [batch][n][n] (cotangent of the input) - exercises the reverse pass in isolation */
int qocb_expm_vjp_batched(int32_t n, int64_t batch, const double *a, const double *ubar, double *out, double *abar,
                          int32_t device);
/* device-resident timing of the batched expm: returns ms per launch (best of `iters`) */
int qocb_expm_batched_time(int32_t n, int64_t batch, double norm_scale, int32_t iters, double *ms_best, int32_t device);

/* the same with `warmup` untimed launches and the total over exactly `iters` timed ones (bench.py --workload expm_batched_n*):
   ms_total / ms_best may be NULL.  n <= 4 runs the register-resident kernels (one matrix per thread for n <= 2, per 4-lane
   group for n = 3, 4) on [batch][n][n] interleaved complex128 with a 256 MiB L2 flush between launches. */
int qocb_expm_batched_bench(int32_t n, int64_t batch, double norm_scale, int32_t warmup, int32_t iters, double *ms_total,
                            double *ms_best, int32_t device);

/* ---- Lindblad path: _evaluate_lindblad_discrete (qoc/core/lindbladdiscrete.py:357-441) and its jacobian (:322) -------
   Densities are [D][n][n] complex128, interleaved (NumPy C order).  H(x) = H0 + sum_r x_r A_r as above; the dissipator is
   sum_l gamma_l (L_l rho L_l^dag - 1/2 {L_l^dag L_l, rho}) with time-independent (gamma_l, L_l).  Each of the N-1
   intervals is a fresh adaptive Dormand-Prince 5(4) integration (qoc/core/mathmethods.py:352-480).

   CONTRACT OF THE LINDBLAD RESULTS (tolerances are asserted in tests/test_gpu_lindblad.py):
     - cost and final densities: the reference's adaptive integration at atol = 1e-12, rtol = 0.  A rounding-level change of
       one error norm near the accept threshold moves the whole step sequence, so two correct implementations agree to the
       integrator's own tolerance, not to rounding: |cost - reference| and the final densities within 1e-9.
     - gradient: the DISCRETE ADJOINT OF THE DORMAND-PRINCE MAP ON THE REALISED STEP GRID - every accepted step (x, h) of the
       forward pass is held fixed and the Runge-Kutta stages, the 4th-order dense output at the interval end and the FSAL
       coupling are differentiated exactly.  The reference's autograd tape additionally differentiates the step-size
       controller (qoc/core/mathmethods.py:436-462), whose inputs are error norms taken at the rounding floor; those terms are
       noise (perturbing the controls by 1e-13 moves the reference-equivalent full gradient by 1e-5 .. 1e-4).  The gradient
       returned here matches the reference-equivalent gradient with the step grid frozen within 1e-7 relative, and the full
       one within its own reproducibility band.  It is NOT claimed at the 1e-10 of the Schroedinger path. */
typedef struct qocb_lplan qocb_lplan;

typedef struct {
    int32_t hilbert_size;        /* n */
    int32_t density_count;       /* D */
    int32_t control_count;       /* KR real control channels (0 with have_hamiltonian = 0) */
    int32_t control_eval_count;  /* M */
    int32_t system_eval_count;   /* N */
    int32_t cost_eval_step;
    int32_t lindblad_count;      /* L (0: lindblad_data = None) */
    int32_t have_hamiltonian;    /* 0: hamiltonian = None (lindbladdiscrete.py:480-484) */
    int32_t device;
    int32_t max_rk_steps;        /* capacity of the accepted-step tape of one evaluation; 0 = automatic */
    int32_t reserved0, reserved1;
    double evolution_time;
} qocb_lindblad_problem;

#define QOCB_LCOST_TARGET 0   /* w * (1 - sum_d |tr(T_d^dag rho_d)| / (D n))        targetdensityinfidelity.py:60-66 */
#define QOCB_LCOST_FORBID 1   /* w * sum_d (1/F_d) sum_f |tr(F_df^dag rho_d) / n|^2  forbiddensities.py:67-85        */

int qocb_lindblad_create(const qocb_lindblad_problem *problem, qocb_lplan **plan_out);
int qocb_lindblad_destroy(qocb_lplan *plan);
const char *qocb_lindblad_last_error(const qocb_lplan *plan);
/* h0: [n][n] (NULL without hamiltonian); a_ops: [KR][n][n]; gammas: [L]; lops: [L][n][n] */
int qocb_lindblad_set_operators(qocb_lplan *plan, const double *h0, const double *a_ops, const double *gammas,
                                const double *lops);
int qocb_lindblad_set_densities(qocb_lplan *plan, const double *rho0);
/* mats: [D][fmax][n][n]; counts: [D] or NULL; weight = cost_multiplier / normalisation; step_cost as above */
int qocb_lindblad_add_cost(qocb_lplan *plan, int32_t kind, int32_t step_cost, double weight, const double *mats,
                           const int32_t *counts, int32_t fmax);
/* controls: [M][KR]; final_densities: [D][n][n] or NULL; grad: [M][KR] */
int qocb_lindblad_cost(qocb_lplan *plan, const double *controls, double *cost, double *final_densities);
int qocb_lindblad_cost_and_grad(qocb_lplan *plan, const double *controls, double *cost, double *grad,
                                double *final_densities);
/* stats[0] = Runge-Kutta attempts, stats[1] = accepted steps of the last evaluation */
int qocb_lindblad_stats(qocb_lplan *plan, int64_t *stats);
/* densities at every system step of the last evaluation: [N][D][n][n] (save_intermediate_densities payload) */
int qocb_lindblad_get_densities(qocb_lplan *plan, double *densities);

const char *qocb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QOCB200_H */
