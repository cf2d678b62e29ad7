// rdfwi_common.cuh -- shared declarations of the sm_100a wave-solver library (see include/rdfwi.h).
//
// Wavefield layout in HBM ("pitched periodic layout")
//   One time level of one shot is nzp rows of `pitch` floats, pitch = nxp rounded up to 4, so every
//   row starts 16-byte aligned and a thread owns one float4 = 4 consecutive cells of a row.
//   Columns nxp .. pitch-1 are *periodic images* of columns 0 .. pitch-nxp-1: they are computed with
//   the coefficients of the cells they mirror, so they always hold bit-identical copies and the
//   float4 holding the last real columns sees its right-hand neighbours in-register.  This is how the
//   reference's torch.roll wrap-around (solvers/pde.py:79) is honoured without a branch in the
//   stencil; only the four scalar x-neighbour loads and the row offsets use wrapped indices.
//   Rotating scratch levels are stored level(t)[shot][z][x]; the wavefield history is shot-major,
//   hist[shot][t][z][x] for t = 0 .. nt-1, so one shot's levels stream contiguously (the cluster-resident
//   kernels write / read them with 1-D bulk copies) and the per-level kernels address it with a shot stride.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#include <string>
#include <vector>

#include "rdfwi.h"

namespace rdfwi {

constexpr int kThreads = 256;      // threads per CTA of the step kernels
constexpr int kMinBlocks = 32;     // CTAs per model in the min/arg-min and velmin-term reductions
constexpr int kClusterThreads = 512;  // threads per CTA of the cluster-resident kernels (<= 128 registers each)
constexpr int kClusterRowsMax = 13;   // rows marched per thread there (coefficients stay in registers)

// Geometry every kernel needs, passed by value (lives in constant bank).
struct Grid {
    int nz, nx, nbc;    // unpadded model, sponge width
    int nzp, nxp;       // padded grid
    int pitch, q4;      // floats per row, float4 per row
    int ns, nrec, nt_out;
    int isz, igz;
    unsigned long long level;  // nzp * pitch floats
};

struct Plan {
    int device = 0;
    Grid g{};
    int nt = 0, st = 1;
    double dx = 0, dt = 0;
    float dx_f = 0, dt_f = 0;
    float a_f = 0, two_a_f = 0, log1e7_f = 0;  // get_Abc constants (solvers/pde.py:42-43)
    std::vector<float> wavelet;                // fp32 cast of the host wavelet (what `tensor * src[i]` uses)
    // device tables (owned)
    int *d_isx = nullptr;      // (ns)
    int *d_rec_ptr = nullptr;  // (nxp+1) CSR: receivers sitting in each padded column
    int *d_rec_idx = nullptr;  // (nrec)
    int *d_igx = nullptr;      // (nrec) padded column of every receiver
    bool rec_simple = false;   // no padded column holds more than one receiver
    float *d_r2 = nullptr;     // (nbc+1) (k*dx/a)^2, entry nbc = 0
    float *d_dkap = nullptr;   // (nbc+1) d(kappa*dt)/d(velmin) per profile entry, entry nbc = 0
    float *d_wavelet = nullptr;  // (nt) fp32 wavelet for the cluster-resident kernels
    // options
    int chunk_models = 0;   // 0 = auto
    int rows_per_thread = 4;
    int adj_rows_per_thread = 1;
    int engine = 0;         // 0 = auto, 1 = per-level kernels, 2 = cluster-resident time loop
    int cluster_size = 0;   // 0 = smallest cluster that fits
    int cluster_rows = 0;   // rows marched per thread of k_fwd_cluster: 0 = auto (13; 7 or 4 on wider clusters for few shots)
    int adj_mode = 0;         // 0 = auto (split: cluster u-field kernel + streaming imaging kernel), 1 = fused per-level adjoint (k_adj_step)
    int imaging = 0;          // cluster engine: where the imaging sums are formed -- 0 / 2 = inside the adjoint sweep, accumulators
                              // in tensor memory (default), 1 = streaming kernel over two histories (the split adjoint)
    int perturb = 0;          // debug: seed of the schedule perturbation of k_fwd_cluster (0 = off), see rdfwi.h
    int img_prefetch = 0;     // imaging kernel: levels ahead pulled into L2 (0 = default 4)
    long long *trace_ptr = nullptr;  // debug: device buffer for per-warp timeline stamps of k_fwd_cluster
    int last_u_chunk = 0;     // shots per chunk of the last split adjoint
    int last_split = 0;       // whether the last rdfwi_backward ran the split adjoint (reported by rdfwi_plan_get "adj_split")
    int u_chunk_shots = 0;    // shots whose adjoint-field history is in flight at once in split mode (0 = auto)
    // optional per-kernel-class timing with CUDA events on the caller's stream (rdfwi_plan_set "timing")
    int timing = 0;
    struct Span { cudaEvent_t a, b; int kind; };
    std::vector<Span> spans;  // kind: 0 forward time loop, 1 adjoint-field time loop, 2 imaging, 3 fused / per-level adjoint loop,
                              // 4 cluster-resident adjoint with the imaging sums in the sweep
    int history_segment = 0;  // 0 = keep every level; K >= 3 = checkpoint pairs every K levels, recompute in the backward pass
                              // (K >= nt: no history at all -- the backward pass recomputes the forward field chunk by chunk)
    long long scratch_mb = 0; // cap on ONE scratch history of the split adjoint, MB (0 = 40000)
    mutable int wave_keys[8] = {0}, wave_vals[8] = {0}, wave_n = 0;  // cached fwd_cluster_wave() per configuration
    int last_fwd_C = 0, last_fwd_rows = 0;  // configuration of the last k_fwd_cluster launch (reported by rdfwi_plan_get)
};

// Cluster-resident forward time loop (kernels_cluster.cu).
struct ClusterFwdArgs {
    const float *alpha;     // (B, nzp, pitch)
    const float *kap;       // (B, nbc+1)
    const float *beta_src;  // (B*ns)
    const int *isx;
    const int *rec_ptr;
    const int *rec_idx;
    const float *wavelet;   // (nt) device
    float *seis;            // (B*ns, nt_out, nrec)
    float *hist;            // [shot][t][z][x], t = 0..nt-2, or nullptr (indexed by the launch-local shot)
    int nshots, nt, st;
    long long *trace;       // debug: per-warp clock64 stamps (nullptr = off)
    int shot0;              // global index of the launch's first shot (seis / cot / Gb / model lookup)
    int adj_mode;           // 0 = forward wavefield, 1 = adjoint field in the u-variable -> hist, 2 = adjoint field with the
                            // imaging sums formed in the kernel (accumulators in tensor memory) -> Ga, Gk (see k_fwd_cluster)
    const float *phist;     // mode 2: forward history [shot][t][z][x] of the launch's shots (launch-local shot index)
    float *Ga, *Gk;         // mode 2: (B*ns, nzp, pitch) imaging planes per shot
    const float *cot;       // adjoint mode: (B*ns, nt_out, nrec) cotangent of the seismograms
    float *Gb;              // adjoint mode: (B*ns) sum_t u_t[src] w_t / alpha_src
    int slabrows, ngroups, wav_smem;  // filled by launch_fwd_cluster from the ClusterConfig
    unsigned perturb;       // debug: seed of pseudo-random per-warp delays at the synchronisation points (0 = off)
    int rows_flip;          // which CTAs take the nzp % C extra rows (filled by launch_fwd_cluster)
    const int *rec_col;     // (nrec) padded column of every receiver
    int rec_simple;         // no column holds more than one receiver: the receiver-major sampling / cotangent paths apply
};

struct ClusterConfig {
    int C = 0;        // CTAs per cluster = row slabs per shot
    int maxrows = 0;  // rows of the largest slab
    int ngroups = 0;  // row groups per CTA (threads = ngroups * q4), rmax rows per thread
    int slabrows = 0; // ngroups * rmax >= maxrows (rows allocated per buffer, halos excluded)
    size_t smem = 0;  // dynamic shared memory per CTA
    int rmax = 0;     // rows per thread (template instantiation)
    bool wav_smem = false;  // the wavelet is staged in shared memory
    bool img = false;       // sized for the resident imaging (MODE 2 of k_fwd_cluster: two 16-byte slots per thread more)
    int nthreads = 512;     // threads per CTA
};

// One time level of the tiled per-level kernel (kernels_tile.cu): forward field, or adjoint field in the u-variable.
struct StepArgs {
    const float *p1;        // level t-1 (adjoint: u_{t+1}) of the launch's first shot
    const float *p0;        // level t-2 (adjoint: u_{t+2})
    float *out;             // level t
    unsigned long long ss_p1, ss_p0, ss_out;  // floats between consecutive shots in p1 / p0 / out (0 = one shared zero level)
    const float *alpha;     // (B, nzp, pitch)   whole batch: indexed by the global model of a shot
    const float *kap;       // (B, nbc+1)
    const float *beta_src;  // (B*ns)
    const int *isx;
    const int *rec_ptr;
    const int *rec_idx;
    float *seis;            // forward: (B*ns, nt_out, nrec) or nullptr when this level is not sampled
    const float *cot;       // adjoint: (B*ns, nt_out, nrec) or nullptr when this level carries no cotangent
    float *Gb;              // adjoint: (B*ns) running sum_t u_t[src] w_t; divided by alpha_src at the last level
    int it_out;
    float w_t;
    int shot0, nshots;      // global index of the launch's first shot, shots in the launch (grid.z)
    int adj, last;          // adjoint-field mode; last level of the reverse loop (t = 0)
};

struct AdjArgs {
    const float *q1;   // q_{t+1}
    const float *q2;   // q_{t+2}
    float *out;        // q_t
    const float *pm1;  // forward level t-1 (zeros for t = 0)
    const float *alpha;
    const float *kap;
    const int *isx;
    const int *rec_ptr;
    const int *rec_idx;
    const float *cot;  // (nb, ns, nt_out, nrec) or nullptr when level t carries no cotangent
    int it_out;
    float w_t;
    float *Ga;  // (nb, nzp, pitch)  sum_t,s q_t (S-5) p_{t-1}
    float *Gk;  // (nb, nzp, pitch)  sum_t,s (q_{t+1}-q_t) p_{t-1}
    float *Gb;  // (nb, ns)          sum_t   q_t[src] w_t
    unsigned long long ss_pm1;  // floats between consecutive shots in pm1 (history or the zero level)
    int slices;                 // grid.z slices the shots are dealt over = imaging planes per model (Ga, Gk)
};

// cudaDeviceGetAttribute, asked once per (device, attribute): the answers never change, and the library is called inside
// CUDA-graph capture (core/inversion.py), where the fewer runtime queries the better.
inline cudaError_t device_attr(int *value, cudaDeviceAttr attr, int device)
{
    static int cache[64][3];
    static bool have[64][3] = {};
    const int slot = attr == cudaDevAttrMaxSharedMemoryPerBlockOptin ? 0 : attr == cudaDevAttrMultiProcessorCount ? 1 : 2;
    const int d = device & 63;
    if (have[d][slot]) { *value = cache[d][slot]; return cudaSuccess; }
    const cudaError_t e = cudaDeviceGetAttribute(value, attr, device);
    if (e == cudaSuccess) { cache[d][slot] = *value; have[d][slot] = true; }
    return e;
}

void set_error(const std::string &msg);
void count_launch();

// kernels_prologue.cu
cudaError_t launch_coefficients(const Plan &p, const float *v, int B, float *alpha_pad, float *kap, float *velmin,
                                int *argmin, float *beta_src, float *minpart, cudaStream_t st);
// kernels_tile.cu
cudaError_t launch_step_tile(const Plan &p, const StepArgs &a, cudaStream_t st);
// kernels_step.cu
cudaError_t launch_adj_step(const Plan &p, const AdjArgs &a, int nb, cudaStream_t st);
int adj_shot_slices(const Plan &p, int nb);  // imaging planes per model the per-level adjoint accumulates into
// kernels_cluster.cu
// nshots = 0: the throughput configuration (smallest cluster that fits, 13 rows per thread).  nshots > 0: the
// configuration for a launch of that many shots -- when they are so few that they leave most SMs idle, a wider cluster
// with fewer rows per thread (shorter sweeps, same arithmetic per cell) as long as all shots stay co-resident.
bool cluster_config(const Plan &p, ClusterConfig *cfg, int nshots = 0, bool img = false);
cudaError_t launch_fwd_cluster(const Plan &p, const ClusterConfig &cc, ClusterFwdArgs a, cudaStream_t st);
int fwd_cluster_wave(const Plan &p, const ClusterConfig &cc);  // co-resident clusters = shots in flight per wave
// kernels_imaging.cu: zero-lag imaging sums of `nshots` shots from the forward history and the adjoint-field history
cudaError_t launch_imaging(const Plan &p, const float *phist, const float *uhist, const float *alpha, const float *kap,
                           const float *beta_src, const float *Gb, float *Ga, float *Gk, int shot0, int nshots, int pshot0,
                           cudaStream_t st);
// kernels_epilogue.cu  (planes = imaging planes per model: 1 for the per-level engine, ns for the cluster engine)
cudaError_t launch_gradient_epilogue(const Plan &p, const float *v, int B, const float *Ga, const float *Gk,
                                     const float *Gb, int planes, const int *argmin, float *fold_tmp,
                                     double *vel_part, float *grad_v, cudaStream_t st);

// kernels_misfit.cu: per-model masked L1 sums + sign field of the seismogram residual (core/losses.py:27-40)
size_t misfit_scratch_bytes(int B);
cudaError_t launch_misfit_l1(const float *seis, const float *obs, const float *mask, int B, long long n, float *sign_out,
                             double *stats, double *part, cudaStream_t st);

}  // namespace rdfwi
