/*
 * rdfwi.h -- C ABI of the B200 (sm_100a) finite-difference wave solver that replaces the
 * time loop of the reference's red_diffeq/solvers/pde.py and the autograd tape behind it.
 *
 * The reference has no FFI for this path: its boundary is the Python class
 * red_diffeq.solvers.pde.FWIForward (solvers/pde.py:6-93).  The entry points below are what a
 * ctypes binding of that class calls instead of
 *   - FWIForward.forward / FWM          (solvers/pde.py:61-93)  -> rdfwi_forward
 *   - loss.backward() through the tape  (core/inversion.py:86)   -> rdfwi_backward
 *   - get_Abc / alpha / beta_dt planes  (solvers/pde.py:38-52, :63-71) -> rdfwi_coefficients (test hook)
 * INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*.
 *   - every large buffer (outputs, workspace, wavefield history) is allocated by the caller on the
 *     plan's device and passed in; the library only enqueues work on `stream` and never
 *     synchronises it.  A plan owns a few KB of device tables (geometry, wavelet).
 *   - all functions return 0 on success, a RDFWI_E* code otherwise; rdfwi_last_error() gives the
 *     message of the calling thread's last failure.
 *   - fp32 everywhere on the device (the reference computes in fp32).
 */
#ifndef RDFWI_H
#define RDFWI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RDFWI_VERSION 100

enum {
    RDFWI_OK = 0,
    RDFWI_EINVAL = 1, /* bad argument / unsupported geometry          */
    RDFWI_ECUDA = 2,  /* a CUDA runtime call failed                   */
    RDFWI_ESIZE = 3   /* caller-provided buffer too small             */
};

/* Acquisition geometry and discretisation: the reference's ctx dict (`pde:` YAML block,
 * solvers/pde.py:8-24) after adj_sr() (:54-59) and ricker() (:26-36) have been applied on the host. */
typedef struct rdfwi_survey {
    int32_t nz, nx;          /* unpadded velocity model, rows (depth) x columns                 */
    int32_t nbc;             /* sponge / replicate-padding width (solvers/pde.py:91)            */
    int32_t ns, nrec;        /* shots per model, receivers per shot                             */
    int32_t nt;              /* time levels                                                     */
    int32_t sample_temporal; /* keep every k-th level in the seismogram (solvers/pde.py:82)     */
    int32_t isz, igz;        /* padded-grid source / receiver row, in [0, nz+2*nbc)             */
    double dx, dt;           /* grid step (m), time step (s)                                    */
    const int32_t *isx;      /* host, (ns)   padded-grid source columns, in [0, nx+2*nbc)       */
    const int32_t *igx;      /* host, (nrec) padded-grid receiver columns (duplicates allowed)  */
    const double *wavelet;   /* host, (nt)   source time function, float64 as ricker() returns  */
} rdfwi_survey;

typedef struct rdfwi_plan_s *rdfwi_plan;

int rdfwi_version(void);
const char *rdfwi_last_error(void);

/* Creates a plan on the current CUDA device.  Copies the survey (pointers need not outlive the call). */
int rdfwi_plan_create(const rdfwi_survey *survey, rdfwi_plan *plan_out);
int rdfwi_plan_destroy(rdfwi_plan plan);

/* Tunables (all optional).  Keys:
 *   "engine"         0 = by problem size (default), 1 = per-level tiled kernels, 2 = cluster-resident time loop
 *   "adj_mode"       0 = adjoint on the engine the forward pass ran on (default), 1 = fused per-level adjoint (one launch
 *                    per level, imaging sums read-modify-written in HBM; needs no scratch history)
 *   "imaging"        cluster-resident engine: where the zero-lag imaging sums are formed.  0 / 2 = inside the adjoint sweep,
 *                    accumulators in tensor memory, forward history read once (default); 1 = split adjoint: the adjoint
 *                    field is written to a scratch history and a streaming kernel reads both histories
 *   "history_segment" 0 = keep every level; K >= 3 = keep a pair of levels every K levels and recompute K levels at a
 *                    time in the backward pass; K >= nt = keep nothing, the backward pass recomputes the forward field
 *                    chunk by chunk (must be set BEFORE the workspace / history sizes are queried; the `segment`
 *                    argument of forward/backward must equal it)
 *   "u_chunk_shots"  shots per chunk of the split adjoint (0 = auto: whole waves of co-resident clusters)
 *   "scratch_mb"     cap on one scratch history of the split adjoint, MB (0 = 40000; 55000 for the recompute tier)
 *   "cluster_size"   CTAs per cluster of the cluster-resident time loop (0 = smallest of 1..8 or 16 that fits)
 *   "cluster_rows"   rows marched per thread by the cluster-resident time loop: 13, 7, 5 or 4 (0 = auto: 13, or fewer on a
 *                    wider cluster when a launch has so few shots that each still gets its own co-resident cluster)
 *   "rows_per_thread" (1, 2, 4, 8; tile rows = 8x) / "adj_rows_per_thread"  z-rows marched per thread, per-level kernels
 *   "chunk_models"   models advanced together by the per-level forward (0 = auto)
 *   "timing"         1 = record CUDA events around each kernel class on the caller's stream (read back as "us_<class>",
 *                    "n_<class>" with class in forward, adjoint_field, imaging, adjoint_loop, adjoint_resident).  EXCEPTION to "the library
 *                    never synchronises": rdfwi_plan_get("us_*" / "n_*") waits (cudaEventSynchronize) for the recorded
 *                    events, i.e. for the timed work -- a measurement hook, never called by forward / backward themselves
 *   "perturb"        debug: seed (> 0) of pseudo-random per-warp delays (up to ~4 us) in front of every synchronisation point
 *                    of the cluster-resident time loop -- halo waits, early / late halo pushes, bulk-copy hand-over, sampling
 *                    warp -- run by separate kernel instantiations; results must not change (tests/test_gpu_perturb.py: the
 *                    substitute for racecheck, which is closed on the GPU pool).  0 = off (production kernels)
 * rdfwi_plan_get additionally answers "pitch", "nzp", "nxp", "nt_out", "cluster_size_used", "cluster_size_last", "cluster_rows_last" (what the
 * last cluster-resident launch ran),
 * "cluster_wave" (co-resident clusters of the forward configuration), "adj_split" (what the last backward ran: 0 per-level fused,
 * 1 cluster split, 2 cluster split on a recomputed forward history, 3 per-level split, 4 cluster resident -- the imaging sums
 * formed inside the adjoint sweep --, 5 cluster resident on a recomputed forward history), "u_chunk_used". */
int rdfwi_plan_set(rdfwi_plan plan, const char *key, int64_t value);
int rdfwi_plan_get(rdfwi_plan plan, const char *key, int64_t *value_out);

/* Floats per wavefield level of one shot in the library's internal pitched layout. */
size_t rdfwi_level_floats(rdfwi_plan plan);

/* Scratch the caller must provide to forward / backward / coefficients for a batch of B models. */
size_t rdfwi_workspace_bytes(rdfwi_plan plan, int32_t B);

/* Bytes of wavefield history rdfwi_forward must be given so that rdfwi_backward can run.
 *   segment == 0 : every level is kept (B*ns*nt levels)
 *   segment == K : the pair (p_{jK-2}, p_{jK-1}) in front of every segment j >= 1 of K levels is kept; the backward
 *                  pass recomputes one segment at a time into the workspace (one extra forward, memory / (K/2)).
 *                  0 bytes when K >= nt (single segment: nothing is kept, the backward pass recomputes the forward
 *                  field from the zero initial state); a NULL history is then accepted by forward and backward. */
size_t rdfwi_history_bytes(rdfwi_plan plan, int32_t B, int32_t segment);

/*
 * Forward modelling  (replaces FWIForward.forward, solvers/pde.py:88-93, after v_denorm_func).
 *   v_phys   device, (B, nz, nx) fp32 contiguous, velocity in m/s
 *   seis     device, (B, ns, ceil(nt/sample_temporal), nrec) fp32, written
 *   history  device or NULL.  NULL = forward only (the reference under torch.no_grad()).
 */
int rdfwi_forward(rdfwi_plan plan, const float *v_phys, int32_t B, float *seis,
                  void *workspace, size_t workspace_bytes,
                  void *history, size_t history_bytes, int32_t segment, void *stream);

/*
 * Adjoint pass: grad_v = d sum(seis * grad_seis) / d v_phys, including the sponge's dependence on
 * min(v) (solvers/pde.py:41) and the fold of the replicate padding (:91) -- i.e. exactly what
 * autograd returns for the input of FWIForward.FWM's caller.
 *   v_phys, history : the same buffers rdfwi_forward was given
 *   grad_seis       : device, same shape as seis
 *   grad_v          : device, (B, nz, nx) fp32, overwritten
 */
int rdfwi_backward(rdfwi_plan plan, const float *v_phys, int32_t B, const float *grad_seis,
                   float *grad_v, void *workspace, size_t workspace_bytes,
                   const void *history, size_t history_bytes, int32_t segment, void *stream);

/*
 * Test hook: the per-model coefficient data the step kernels consume.
 *   alpha_pad  device, (B, nzp, pitch) fp32     alpha = (v*dt/dx)^2 on the padded, pitched grid
 *   kappa_tab  device, (B, nbc+1) fp32          abc*dt profile, entry nbc is the interior's 0
 *   velmin     device, (B) fp32 ; argmin device, (B) int32 (row-major index into (nz, nx))
 *   beta_src   device, (B, ns) fp32             (v*dt)^2 at each source cell
 */
int rdfwi_coefficients(rdfwi_plan plan, const float *v_phys, int32_t B, float *alpha_pad,
                       float *kappa_tab, float *velmin, int32_t *argmin, float *beta_src,
                       void *workspace, size_t workspace_bytes, void *stream);

/*
 * Data misfit of the inversion loop in one pass (opt-in, beyond the drop-in operator): what
 * LossCalculator.observation_loss (core/losses.py:27-40) and its autograd compute with ~10 elementwise ATen kernels.
 *   seis, observed : device, (B, ns, nt_out, nrec) fp32 (the plan's seismogram shape), modelled / observed data
 *   mask           : device, same shape, 1 = observed trace sample, 0 = missing (core/inversion.py:66); NULL = all ones
 *   stats          : device, (B, 2) float64, written: [b][0] = sum |observed - seis| * mask, [b][1] = sum mask.
 *                    The reference's loss is stats[b][0] / max(stats[b][1], 1)  (losses.py:36; the mean when mask is NULL)
 *   sign_out       : device, same shape as seis (may alias seis) or NULL: mask * sign(seis - observed), i.e.
 *                    d stats[b][0] / d seis -- scaled by g_b / count_b it is the cotangent rdfwi_backward takes
 *   workspace      : >= 512 * B bytes of scratch (the forward workspace may be reused)
 */
int rdfwi_misfit_l1(rdfwi_plan plan, const float *seis, const float *observed, const float *mask, int32_t B,
                    double *stats, float *sign_out, void *workspace, size_t workspace_bytes, void *stream);

/* Number of kernel launches enqueued by this thread's last rdfwi_forward / rdfwi_backward / rdfwi_misfit_l1 call. */
int64_t rdfwi_last_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* RDFWI_H */
