// rdfwi_api.cu -- C ABI (include/rdfwi.h) and the host-side time-loop orchestration.
//
// The reference's FWM (solvers/pde.py:61-86) is a Python loop of nt iterations over the whole (B, ns)
// batch.  Here the batch is advanced in *chunks* of models whose three live levels fit in the L2 cache:
// a chunk runs its whole time loop before the next one starts, so that p_{t-1} and p_{t-2} of a level
// are L2 hits and HBM only sees the history write (forward) / history read (adjoint).
#include <cmath>
#include <cstring>
#include <algorithm>

#include "rdfwi_common.cuh"

namespace rdfwi {

static thread_local std::string t_error;
static thread_local int64_t t_launches = 0;

void set_error(const std::string &msg) { t_error = msg; }
void count_launch() { ++t_launches; }

namespace {

#define RD_CUDA(call)                                                                              \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            cudaGetLastError(); /* a failed launch must not be reported again by the next call */  \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                         \
            return RDFWI_ECUDA;                                                                    \
        }                                                                                          \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

size_t align_up(size_t n, size_t a) { return (n + a - 1) / a * a; }

// RAII CUDA-event bracket around the launches of one kernel class (only when the plan's "timing" option is on)
struct Timed {
    Plan *p;
    cudaStream_t st;
    Plan::Span span{};
    bool on;
    Timed(const Plan &plan, int kind, cudaStream_t stream) : p(const_cast<Plan *>(&plan)), st(stream), on(plan.timing != 0)
    {
        if (!on) return;
        span.kind = kind;
        on = cudaEventCreate(&span.a) == cudaSuccess && cudaEventCreate(&span.b) == cudaSuccess;
        if (on) cudaEventRecord(span.a, st);
    }
    ~Timed()
    {
        if (!on) return;
        cudaEventRecord(span.b, st);
        p->spans.push_back(span);
    }
};

// sum (microseconds) and count of the recorded spans of one kind; synchronises on their events
void span_total(Plan *p, int kind, double *us, int *count)
{
    *us = 0.0; *count = 0;
    for (const Plan::Span &s : p->spans) {
        if (s.kind != kind) continue;
        float ms = 0.f;
        if (cudaEventSynchronize(s.b) == cudaSuccess && cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) { *us += 1e3 * ms; ++*count; }
    }
}

void clear_spans(Plan *p)
{
    for (const Plan::Span &s : p->spans) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    p->spans.clear();
}

int chunk_models(const Plan &p, int B)
{
    int nb = p.chunk_models;
    if (nb <= 0) {
        int l2 = 0;
        if (device_attr(&l2, cudaDevAttrL2CacheSize, p.device) != cudaSuccess || l2 <= 0) l2 = 64 << 20;
        // three live levels (p_{t-1}, p_{t-2}, p_t) of every shot of the chunk in ~40% of L2
        const double per_model = 3.0 * p.g.ns * (double)p.g.level * sizeof(float);
        nb = (int)std::max(1.0, std::floor(0.4 * l2 / per_model));
    }
    nb = std::min(nb, B);
    const int nchunks = (B + nb - 1) / nb;
    return (B + nchunks - 1) / nchunks;  // even out the chunks
}

// Carve-up of the caller's workspace.
struct Workspace {
    float *alpha = nullptr, *kap = nullptr, *velmin = nullptr, *beta_src = nullptr, *minpart = nullptr;
    int *argmin = nullptr;
    float *fields = nullptr;  // 3 rotating levels of one chunk
    float *zero = nullptr;    // one zero level of one chunk (p_{-1}, p_{-2})
    float *Ga = nullptr, *Gk = nullptr, *Gb = nullptr, *fold_tmp = nullptr;
    double *vel_part = nullptr;
    size_t bytes = 0;
    size_t chunk_level = 0;  // floats of one level of a full chunk
    int nb = 0;
    int g_planes = 1;
    float *seg_hist = nullptr;
    bool split = false;
    bool split_tile = false;  // split adjoint on the per-level engine: tiled adjoint-field levels + streaming imaging
    bool recompute = false;   // split adjoint on a forward history recomputed chunk by chunk (history_segment >= nt)
    bool resident = false;    // cluster engine, imaging sums formed inside the adjoint sweep (tensor memory): no adjoint-field history
    int u_chunk = 0;
    float *u_hist = nullptr;
    float *p_hist = nullptr;  // recompute mode: forward history of one chunk of shots
};

int cached_wave(const Plan &p, const ClusterConfig &cc) { return fwd_cluster_wave(p, cc); }  // cached per configuration

Workspace carve(const Plan &p, int B, void *base)
{
    Workspace w;
    const Grid &g = p.g;
    w.nb = chunk_models(p, B);
    w.chunk_level = (size_t)w.nb * g.ns * g.level;
    size_t off = 0;
    char *b = static_cast<char *>(base);
    auto take = [&](size_t nbytes) {
        char *ptr = b ? b + off : nullptr;
        off += align_up(nbytes, 256);
        return ptr;
    };
    w.alpha = (float *)take((size_t)B * g.level * 4);
    w.kap = (float *)take((size_t)B * (g.nbc + 1) * 4);
    w.velmin = (float *)take((size_t)B * 4);
    w.argmin = (int *)take((size_t)B * 4);
    w.beta_src = (float *)take((size_t)B * g.ns * 4);
    w.minpart = (float *)take((size_t)B * kMinBlocks * 2 * 4);
    w.fields = (float *)take(3 * w.chunk_level * 4);
    w.zero = (float *)take(w.chunk_level * 4);
    ClusterConfig fcc;
    // split adjoint (cluster u-field kernel + streaming imaging kernel) whenever the forward cluster kernel fits; histories
    // checkpointed in time and adj_mode = 1 run the per-level adjoint.  history_segment >= nt (a single segment: nothing is
    // kept) runs the split adjoint on a forward history recomputed chunk by chunk.
    const bool single_segment = p.history_segment >= p.nt;
    const bool cluster_ok = p.engine != 1 && (p.history_segment == 0 || single_segment);
    w.split = cluster_ok && (p.adj_mode == 0 || single_segment) && cluster_config(p, &fcc);
    w.recompute = w.split && single_segment;
    ClusterConfig icc;  // the same with room for the resident imaging's staging slots (may need a larger cluster)
    w.resident = w.split && p.imaging != 1 && cluster_config(p, &icc, 0, true);
    w.u_chunk = 0;
    if (w.split) {
        // Shots whose adjoint-field history is in flight at once: whole waves of co-resident clusters (33 four-CTA or 24
        // six-CTA clusters per wave), two waves when the scratch cap allows, evened out over the chunks (long records that
        // leave less than a wave per chunk are better served by history_segment >= nt: the operator's policy picks that).
        const int nshots = B * g.ns;
        const int wave = cached_wave(p, w.resident ? icc : fcc);
        const double per_shot = (double)p.nt * (double)g.level * sizeof(float);
        // cap on one scratch history: 40 GB beside a full forward history, 55 GB each for the two of the recompute tier
        // (resident imaging needs no adjoint-field history: the recompute tier's one scratch history may take 110 GB)
        const double cap = p.scratch_mb > 0 ? 1e6 * (double)p.scratch_mb : (single_segment ? (w.resident ? 110e9 : 55e9) : 40e9);
        int chunk = p.u_chunk_shots;
        if (chunk <= 0) {
            const int fit = std::max(1, (int)(cap / per_shot));
            chunk = fit >= 2 * wave ? 2 * wave : (fit >= wave ? wave : fit);
        }
        chunk = std::min(chunk, nshots);
        const int nchunks = (nshots + chunk - 1) / chunk;
        w.u_chunk = (nshots + nchunks - 1) / nchunks;
    }
    // per-level engine, every level kept: the adjoint is split too (tiled adjoint-field levels into a scratch history of a
    // chunk of shots, then the pointwise imaging kernel); adj_mode = 1 keeps the fused per-level adjoint
    w.split_tile = !w.split && p.history_segment == 0 && p.adj_mode == 0;
    if (w.split_tile) {
        const int nshots = B * g.ns;
        const double per_shot = (double)p.nt * (double)g.level * sizeof(float);
        const double cap = p.scratch_mb > 0 ? 1e6 * (double)p.scratch_mb : 40e9;
        int chunk = p.u_chunk_shots > 0 ? p.u_chunk_shots : std::max(1, (int)(cap / per_shot));
        chunk = std::min(chunk, nshots);
        const int nchunks = (nshots + chunk - 1) / chunk;
        w.u_chunk = (nshots + nchunks - 1) / nchunks;
    }
    // imaging planes per model: one per shot for the cluster engines, one per grid.z slice for the per-level adjoint
    w.g_planes = (w.split || w.split_tile) ? g.ns : adj_shot_slices(p, w.nb);
    w.Ga = (float *)take((size_t)B * w.g_planes * g.level * 4);
    w.Gk = (float *)take((size_t)B * w.g_planes * g.level * 4);
    w.Gb = (float *)take((size_t)B * g.ns * 4);
    w.fold_tmp = (float *)take((size_t)B * g.nz * g.nxp * 4);
    w.vel_part = (double *)take((size_t)B * kMinBlocks * 8);
    // checkpoint mode: the levels of one segment of one chunk, recomputed during the backward pass
    w.seg_hist = (p.history_segment > 0 && !w.recompute) ? (float *)take(w.chunk_level * (size_t)(p.history_segment - 1) * 4) : nullptr;
    // split adjoint: adjoint-field history of one chunk of shots (+ the recomputed forward history of the chunk)
    if ((w.split && !w.resident) || w.split_tile) w.u_hist = (float *)take((size_t)w.u_chunk * (size_t)p.nt * g.level * 4);
    if (w.recompute) w.p_hist = (float *)take(((size_t)w.u_chunk * (size_t)p.nt * g.level + (size_t)kClusterRowsMax * g.pitch) * 4);
    w.bytes = off;
    return w;
}

int check_common(rdfwi_plan plan, const void *v, int B, const void *ws, size_t ws_bytes)
{
    if (!plan) { set_error("null plan"); return RDFWI_EINVAL; }
    if (!v || B <= 0) { set_error("null velocity pointer or B <= 0"); return RDFWI_EINVAL; }
    if (!ws) { set_error("null workspace"); return RDFWI_EINVAL; }
    const Plan &p = *reinterpret_cast<Plan *>(plan);
    const size_t need = carve(p, B, nullptr).bytes;
    if (ws_bytes < need) {
        set_error("workspace too small: need " + std::to_string(need) + " bytes, got " + std::to_string(ws_bytes));
        return RDFWI_ESIZE;
    }
    if ((reinterpret_cast<uintptr_t>(ws) & 255) != 0) { set_error("workspace must be 256-byte aligned"); return RDFWI_EINVAL; }
    return RDFWI_OK;
}

int num_segments(const Plan &p, int segment) { return (p.nt + segment - 1) / segment; }

size_t history_floats(const Plan &p, int B, int segment)
{
    if (segment > 0)  // the pair (p_{jK-2}, p_{jK-1}) in front of every segment j >= 1
        return (size_t)B * p.g.ns * (size_t)std::max(num_segments(p, segment) - 1, 0) * 2 * p.g.level;
    // every level p_0 .. p_{nt-1}, and a few rows of padding: the resident adjoint's sweeps fetch whole 13-row columns of a
    // slab, the last of which may reach past the last level of the last shot (kernels_cluster.cu, fwd_sweep)
    return (size_t)B * p.g.ns * (size_t)p.nt * p.g.level + (size_t)kClusterRowsMax * p.g.pitch;
}

int check_segment(const Plan &p, int segment)
{
    if (segment != p.history_segment) {
        set_error("segment argument (" + std::to_string(segment) + ") differs from the plan's history_segment option (" +
                  std::to_string(p.history_segment) + "): set it with rdfwi_plan_set before sizing the workspace");
        return RDFWI_EINVAL;
    }
    return RDFWI_OK;
}

}  // namespace
}  // namespace rdfwi

using namespace rdfwi;

extern "C" {

int rdfwi_version(void) { return RDFWI_VERSION; }
const char *rdfwi_last_error(void) { return t_error.c_str(); }
int64_t rdfwi_last_launch_count(void) { return t_launches; }

int rdfwi_plan_create(const rdfwi_survey *s, rdfwi_plan *out)
{
    if (!s || !out) { set_error("null argument"); return RDFWI_EINVAL; }
    *out = nullptr;
    if (s->nz < 1 || s->nx < 1 || s->nbc < 2 || s->ns < 1 || s->nrec < 1 || s->nt < 1 || s->sample_temporal < 1) {
        set_error("invalid survey: need nz,nx,ns,nrec,nt,sample_temporal >= 1 and nbc >= 2");
        return RDFWI_EINVAL;
    }
    if (!s->isx || !s->igx || !s->wavelet) { set_error("null geometry / wavelet pointer"); return RDFWI_EINVAL; }
    if (!(s->dx > 0) || !(s->dt > 0)) { set_error("dx and dt must be positive"); return RDFWI_EINVAL; }
    const int nzp = s->nz + 2 * s->nbc, nxp = s->nx + 2 * s->nbc;
    if (nxp < 8 || nzp < 8) { set_error("padded grid must be at least 8 x 8"); return RDFWI_EINVAL; }
    if ((double)nzp * ((nxp + 3) / 4 * 4) > 2.0e9) { set_error("padded grid too large for 32-bit cell indices"); return RDFWI_EINVAL; }
    if (s->isz < 0 || s->isz >= nzp || s->igz < 0 || s->igz >= nzp) { set_error("source / receiver row outside the padded grid"); return RDFWI_EINVAL; }
    for (int i = 0; i < s->ns; ++i)
        if (s->isx[i] < 0 || s->isx[i] >= nxp) { set_error("source column outside the padded grid"); return RDFWI_EINVAL; }
    for (int i = 0; i < s->nrec; ++i)
        if (s->igx[i] < 0 || s->igx[i] >= nxp) { set_error("receiver column outside the padded grid"); return RDFWI_EINVAL; }

    Plan *p = new Plan();
    RD_CUDA(cudaGetDevice(&p->device));
    Grid &g = p->g;
    g.nz = s->nz; g.nx = s->nx; g.nbc = s->nbc; g.nzp = nzp; g.nxp = nxp;
    g.pitch = (nxp + 3) / 4 * 4; g.q4 = g.pitch / 4;
    g.ns = s->ns; g.nrec = s->nrec;
    g.nt_out = (s->nt + s->sample_temporal - 1) / s->sample_temporal;
    g.isz = s->isz; g.igz = s->igz;
    g.level = (unsigned long long)nzp * g.pitch;
    p->nt = s->nt; p->st = s->sample_temporal;
    p->dx = s->dx; p->dt = s->dt;
    p->dx_f = (float)s->dx; p->dt_f = (float)s->dt;
    // get_Abc: a = (nbc-1)*dx (python float), kappa = 3.0*velmin*np.log(1e7)/(2.0*a)   (solvers/pde.py:42-43)
    const double a_d = (double)(s->nbc - 1) * s->dx;
    p->a_f = (float)a_d;
    p->two_a_f = (float)(2.0 * a_d);
    p->log1e7_f = (float)std::log(10000000.0);
    p->wavelet.resize(s->nt);
    for (int t = 0; t < s->nt; ++t) p->wavelet[t] = (float)s->wavelet[t];

    // host tables
    std::vector<float> r2(s->nbc + 1), dkap(s->nbc + 1);
    const float dkappa0 = (3.0f * p->log1e7_f) / p->two_a_f;
    for (int k = 0; k < s->nbc; ++k) {
        volatile float r = ((float)k * p->dx_f);  // (dimrange*dx/a)**2, one rounding per op (:46)
        r = r / p->a_f;
        volatile float rr = r * r;
        r2[k] = rr;
        dkap[k] = (dkappa0 * rr) * p->dt_f;
    }
    r2[s->nbc] = 0.0f;
    dkap[s->nbc] = 0.0f;
    std::vector<int> rec_ptr(nxp + 1, 0), rec_idx(s->nrec);
    for (int r = 0; r < s->nrec; ++r) rec_ptr[s->igx[r] + 1]++;
    for (int x = 0; x < nxp; ++x) rec_ptr[x + 1] += rec_ptr[x];
    {
        std::vector<int> fill(rec_ptr.begin(), rec_ptr.end() - 1);
        for (int r = 0; r < s->nrec; ++r) rec_idx[fill[s->igx[r]]++] = r;
    }
    auto upload = [&](auto **dst, const void *src, size_t bytes) -> cudaError_t {
        cudaError_t e = cudaMalloc((void **)dst, bytes);
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
    };
    cudaError_t e = upload(&p->d_isx, s->isx, sizeof(int) * s->ns);
    if (e == cudaSuccess) e = upload(&p->d_rec_ptr, rec_ptr.data(), sizeof(int) * rec_ptr.size());
    if (e == cudaSuccess) e = upload(&p->d_rec_idx, rec_idx.data(), sizeof(int) * rec_idx.size());
    if (e == cudaSuccess) e = upload(&p->d_igx, s->igx, sizeof(int) * s->nrec);
    p->rec_simple = true;
    for (int x = 0; x < nxp; ++x) p->rec_simple = p->rec_simple && rec_ptr[x + 1] - rec_ptr[x] <= 1;
    if (e == cudaSuccess) e = upload(&p->d_r2, r2.data(), sizeof(float) * r2.size());
    if (e == cudaSuccess) e = upload(&p->d_dkap, dkap.data(), sizeof(float) * dkap.size());
    if (e == cudaSuccess) e = upload(&p->d_wavelet, p->wavelet.data(), sizeof(float) * p->wavelet.size());
    if (e != cudaSuccess) {
        set_error(std::string("plan tables: ") + cudaGetErrorString(e));
        rdfwi_plan_destroy(reinterpret_cast<rdfwi_plan>(p));
        return RDFWI_ECUDA;
    }
    *out = reinterpret_cast<rdfwi_plan>(p);
    return RDFWI_OK;
}

int rdfwi_plan_destroy(rdfwi_plan plan)
{
    if (!plan) return RDFWI_OK;
    Plan *p = reinterpret_cast<Plan *>(plan);
    DeviceGuard guard(p->device);
    clear_spans(p);
    cudaFree(p->d_isx); cudaFree(p->d_rec_ptr); cudaFree(p->d_rec_idx); cudaFree(p->d_igx); cudaFree(p->d_r2); cudaFree(p->d_dkap); cudaFree(p->d_wavelet);
    delete p;
    return RDFWI_OK;
}

int rdfwi_plan_set(rdfwi_plan plan, const char *key, int64_t value)
{
    if (!plan || !key) { set_error("null argument"); return RDFWI_EINVAL; }
    Plan *p = reinterpret_cast<Plan *>(plan);
    const std::string k(key);
    if (k == "chunk_models") { if (value < 0) goto bad; p->chunk_models = (int)value; }
    else if (k == "rows_per_thread") { if (value != 1 && value != 2 && value != 4 && value != 8) goto bad; p->rows_per_thread = (int)value; }
    else if (k == "adj_rows_per_thread") { if (value != 1 && value != 2) goto bad; p->adj_rows_per_thread = (int)value; }
    else if (k == "engine") { if (value < 0 || value > 2) goto bad; p->engine = (int)value; }
    else if (k == "history_segment") { if (value < 0 || value == 1 || value == 2) goto bad; p->history_segment = (int)value; }
    else if (k == "adj_mode") { if (value < 0 || value > 1) goto bad; p->adj_mode = (int)value; }
    else if (k == "imaging") { if (value < 0 || value > 2) goto bad; p->imaging = (int)value; }
    else if (k == "timing") { clear_spans(p); p->timing = value != 0; }  // (re)starts the per-kernel-class timers
    else if (k == "u_chunk_shots") { if (value < 0) goto bad; p->u_chunk_shots = (int)value; }
    else if (k == "scratch_mb") { if (value < 0) goto bad; p->scratch_mb = value; }
    else if (k == "perturb") { if (value < 0 || value > 0x7fffffff) goto bad; p->perturb = (int)value; }
    else if (k == "img_prefetch") { if (value < 0 || value > 64) goto bad; p->img_prefetch = (int)value; }
    else if (k == "trace_ptr") { p->trace_ptr = reinterpret_cast<long long *>(value); }
    else if (k == "cluster_size") { if (value < 0 || (value > 8 && value != 16)) goto bad; p->cluster_size = (int)value; }
    else if (k == "cluster_rows") { if (value != 0 && value != 4 && value != 5 && value != 7 && value != kClusterRowsMax) goto bad; p->cluster_rows = (int)value; }
    else { set_error("unknown option " + k); return RDFWI_EINVAL; }
    return RDFWI_OK;
bad:
    set_error("invalid value for option " + k);
    return RDFWI_EINVAL;
}

int rdfwi_plan_get(rdfwi_plan plan, const char *key, int64_t *out)
{
    if (!plan || !key || !out) { set_error("null argument"); return RDFWI_EINVAL; }
    Plan *p = reinterpret_cast<Plan *>(plan);
    const std::string k(key);
    if (k == "chunk_models") *out = p->chunk_models;
    else if (k == "rows_per_thread") *out = p->rows_per_thread;
    else if (k == "adj_rows_per_thread") *out = p->adj_rows_per_thread;
    else if (k == "engine") *out = p->engine;
    else if (k == "history_segment") *out = p->history_segment;
    else if (k == "adj_mode") *out = p->adj_mode;
    else if (k == "imaging") *out = p->imaging;
    else if (k.rfind("us_", 0) == 0 || k.rfind("n_", 0) == 0) {
        const bool want_us = k[0] == 'u';
        const std::string what = k.substr(want_us ? 3 : 2);
        const int kind = what == "forward" ? 0 : what == "adjoint_field" ? 1 : what == "imaging" ? 2 : what == "adjoint_loop" ? 3 : what == "adjoint_resident" ? 4 : -1;
        if (kind < 0) { set_error("unknown timer " + k); return RDFWI_EINVAL; }
        double us; int n;
        span_total(p, kind, &us, &n);
        *out = want_us ? (int64_t)(us + 0.5) : n;
    }
    else if (k == "adj_split") *out = p->last_split;
    else if (k == "perturb") *out = p->perturb;
    else if (k == "u_chunk_shots") *out = p->u_chunk_shots;
    else if (k == "u_chunk_used") *out = p->last_u_chunk;
    else if (k == "scratch_mb") *out = p->scratch_mb;
    else if (k == "cluster_wave") { ClusterConfig cc; *out = cluster_config(*p, &cc) ? cached_wave(*p, cc) : 0; }
    else if (k == "cluster_size") *out = p->cluster_size;
    else if (k == "cluster_rows") *out = p->cluster_rows;
    else if (k == "cluster_size_last") *out = p->last_fwd_C;
    else if (k == "cluster_rows_last") *out = p->last_fwd_rows;
    else if (k == "cluster_size_used") { ClusterConfig cc; *out = cluster_config(*p, &cc) ? cc.C : 0; }
    else if (k == "pitch") *out = p->g.pitch;
    else if (k == "nzp") *out = p->g.nzp;
    else if (k == "nxp") *out = p->g.nxp;
    else if (k == "nt_out") *out = p->g.nt_out;
    else { set_error("unknown option " + k); return RDFWI_EINVAL; }
    return RDFWI_OK;
}

size_t rdfwi_level_floats(rdfwi_plan plan) { return plan ? (size_t) reinterpret_cast<Plan *>(plan)->g.level : 0; }

size_t rdfwi_workspace_bytes(rdfwi_plan plan, int32_t B)
{
    if (!plan || B <= 0) return 0;
    return carve(*reinterpret_cast<Plan *>(plan), B, nullptr).bytes;
}

size_t rdfwi_history_bytes(rdfwi_plan plan, int32_t B, int32_t segment)
{
    if (!plan || B <= 0) return 0;
    return history_floats(*reinterpret_cast<Plan *>(plan), B, segment) * sizeof(float);
}

int rdfwi_coefficients(rdfwi_plan plan, const float *v, int32_t B, float *alpha_pad, float *kappa_tab, float *velmin,
                       int32_t *argmin, float *beta_src, void *ws, size_t ws_bytes, void *stream)
{
    int rc = check_common(plan, v, B, ws, ws_bytes);
    if (rc) return rc;
    if (!alpha_pad || !kappa_tab || !velmin || !argmin || !beta_src) { set_error("null output"); return RDFWI_EINVAL; }
    const Plan &p = *reinterpret_cast<Plan *>(plan);
    DeviceGuard guard(p.device);
    t_launches = 0;
    Workspace w = carve(p, B, ws);
    RD_CUDA(launch_coefficients(p, v, B, alpha_pad, kappa_tab, velmin, argmin, beta_src, w.minpart, (cudaStream_t)stream));
    return RDFWI_OK;
}

int rdfwi_forward(rdfwi_plan plan, const float *v, int32_t B, float *seis, void *ws, size_t ws_bytes, void *history,
                  size_t history_bytes, int32_t segment, void *stream)
{
    int rc = check_common(plan, v, B, ws, ws_bytes);
    if (rc) return rc;
    if (!seis) { set_error("null seismogram buffer"); return RDFWI_EINVAL; }
    const Plan &p = *reinterpret_cast<Plan *>(plan);
    const Grid &g = p.g;
    if (history) {
        if ((rc = check_segment(p, segment)) != RDFWI_OK) return rc;
        if (history_bytes < history_floats(p, B, segment) * sizeof(float)) { set_error("history buffer too small"); return RDFWI_ESIZE; }
    }
    DeviceGuard guard(p.device);
    cudaStream_t st = (cudaStream_t)stream;
    t_launches = 0;
    Workspace w = carve(p, B, ws);
    RD_CUDA(launch_coefficients(p, v, B, w.alpha, w.kap, w.velmin, w.argmin, w.beta_src, w.minpart, st));
    float *hist = static_cast<float *>(history);
    const int nt = p.nt;
    if (hist && w.recompute) hist = nullptr;  // nothing is kept: the backward pass recomputes the forward field
    const bool ckpt = hist && segment > 0;  // checkpointed history: per-level engine
    ClusterConfig cc;
    if (p.engine != 1 && !ckpt && cluster_config(p, &cc, B * g.ns)) {
        // cluster-resident time loop: one launch for all shots and all levels
        ClusterFwdArgs a{};
        a.alpha = w.alpha; a.kap = w.kap; a.beta_src = w.beta_src;
        a.isx = p.d_isx; a.rec_ptr = p.d_rec_ptr; a.rec_idx = p.d_rec_idx; a.wavelet = p.d_wavelet;
        a.seis = seis; a.hist = hist; a.trace = p.trace_ptr;
        a.nshots = B * g.ns; a.nt = nt; a.st = p.st;
        Timed timed(p, 0, st);
        RD_CUDA(launch_fwd_cluster(p, cc, a, st));
        return RDFWI_OK;
    }
    if (p.engine == 2 && !ckpt) { set_error("engine=2 (cluster-resident) requested but the grid does not fit a cluster"); return RDFWI_EINVAL; }
    Timed timed_loop(p, 0, st);
    if (hist) RD_CUDA(cudaMemsetAsync(w.zero, 0, w.chunk_level * sizeof(float), st));
    const int K = segment;
    const size_t ck_stride = ckpt ? (size_t)(num_segments(p, K) - 1) * 2 * g.level : 0;  // shot stride of the checkpoints

    for (int b0 = 0; b0 < B; b0 += w.nb) {
        const int nb = std::min(w.nb, B - b0);
        const size_t lvl = (size_t)nb * g.ns * g.level;  // floats per level of this chunk
        const size_t hstride = (size_t)nt * g.level;  // shot stride inside the history
        float *hbase = hist ? hist + (size_t)b0 * g.ns * hstride : nullptr;
        if (!hist || ckpt) RD_CUDA(cudaMemsetAsync(w.fields, 0, 3 * w.chunk_level * sizeof(float), st));
        (void)lvl;
        float *ck_base = ckpt ? hist + (size_t)b0 * g.ns * ck_stride : nullptr;
        auto level_ptr = [&](int t, unsigned long long *stride) -> float * {
            *stride = g.level;
            if (ckpt) {
                // levels jK-2 and jK-1 (j >= 1) live in the checkpoint slots, everything else rotates in scratch
                if (t >= 0) {
                    const int j = (t + 2) / K, which = t - (j * K - 2);  // which = 0 or 1 when t is a checkpoint level
                    if (j >= 1 && j * K < nt && (which == 0 || which == 1)) {
                        *stride = ck_stride;
                        return ck_base + ((size_t)(j - 1) * 2 + which) * g.level;
                    }
                }
                return w.fields + (size_t)((t + 3) % 3) * w.chunk_level;
            }
            if (hist) {
                if (t < 0) return w.zero;
                *stride = hstride;
                return hbase + (size_t)t * g.level;
            }
            return w.fields + (size_t)((t + 3) % 3) * w.chunk_level;
        };
        for (int t = 0; t < nt; ++t) {
            StepArgs a{};
            a.p1 = level_ptr(t - 1, &a.ss_p1);
            a.p0 = level_ptr(t - 2, &a.ss_p0);
            a.out = level_ptr(t, &a.ss_out);
            a.alpha = w.alpha; a.kap = w.kap; a.beta_src = w.beta_src;
            a.isx = p.d_isx; a.rec_ptr = p.d_rec_ptr; a.rec_idx = p.d_rec_idx;
            a.seis = (t % p.st == 0) ? seis : nullptr;
            a.it_out = t / p.st;
            a.w_t = p.wavelet[t];
            a.shot0 = b0 * g.ns; a.nshots = nb * g.ns;
            launch_step_tile(p, a, st);
        }
    }
    RD_CUDA(cudaGetLastError());
    return RDFWI_OK;
}

int rdfwi_backward(rdfwi_plan plan, const float *v, int32_t B, const float *cot, float *grad_v, void *ws, size_t ws_bytes,
                   const void *history, size_t history_bytes, int32_t segment, void *stream)
{
    int rc = check_common(plan, v, B, ws, ws_bytes);
    if (rc) return rc;
    if (!cot || !grad_v) { set_error("null cotangent / gradient"); return RDFWI_EINVAL; }
    const Plan &p = *reinterpret_cast<Plan *>(plan);
    const Grid &g = p.g;
    if ((rc = check_segment(p, segment)) != RDFWI_OK) return rc;
    if (!history && history_floats(p, B, segment) > 0) { set_error("null history"); return RDFWI_EINVAL; }  // (one segment needs none)
    if (history_bytes < history_floats(p, B, segment) * sizeof(float)) { set_error("history buffer too small"); return RDFWI_ESIZE; }
    DeviceGuard guard(p.device);
    cudaStream_t st = (cudaStream_t)stream;
    t_launches = 0;
    Workspace w = carve(p, B, ws);
    RD_CUDA(launch_coefficients(p, v, B, w.alpha, w.kap, w.velmin, w.argmin, w.beta_src, w.minpart, st));
    const float *hist = static_cast<const float *>(history);
    const int nt = p.nt;
    const bool ckpt = segment > 0 && !w.recompute;
    RD_CUDA(cudaMemsetAsync(w.Gb, 0, (size_t)B * g.ns * sizeof(float), st));
    ClusterConfig cc;
    const_cast<Plan &>(p).last_split = (!ckpt && w.split) ? (w.recompute ? 2 : 1) : 0;
    const_cast<Plan &>(p).last_u_chunk = w.u_chunk;
    if (!ckpt && w.resident && cluster_config(p, &cc, w.recompute ? std::min(w.u_chunk, B * g.ns) : B * g.ns, true)) {
        // resident imaging: ONE cluster-resident kernel per launch runs the adjoint field and forms the imaging sums in the
        // sweep (accumulators in tensor memory), reading the forward history once; no adjoint-field history, no imaging pass
        const int nshots = B * g.ns;
        const int step = w.recompute ? w.u_chunk : nshots;
        const_cast<Plan &>(p).last_split = w.recompute ? 5 : 4;
        for (int s0 = 0; s0 < nshots; s0 += step) {
            const int n = std::min(step, nshots - s0);
            ClusterFwdArgs a{};
            a.alpha = w.alpha; a.kap = w.kap; a.beta_src = w.beta_src;
            a.isx = p.d_isx; a.rec_ptr = p.d_rec_ptr; a.rec_idx = p.d_rec_idx; a.wavelet = p.d_wavelet;
            a.nshots = n; a.nt = nt; a.st = p.st; a.shot0 = s0; a.trace = p.trace_ptr;
            if (w.recompute) {  // forward field of the chunk, again, this time keeping every level (no seismograms)
                ClusterConfig fcc;
                if (!cluster_config(p, &fcc, n)) { set_error("no cluster configuration for the recomputed forward field"); return RDFWI_EINVAL; }
                a.seis = nullptr; a.hist = w.p_hist; a.adj_mode = 0;
                Timed timed(p, 0, st);
                RD_CUDA(launch_fwd_cluster(p, fcc, a, st));
            }
            a.seis = nullptr; a.hist = nullptr; a.adj_mode = 2; a.cot = cot; a.Gb = w.Gb;
            a.phist = w.recompute ? w.p_hist : hist + (size_t)s0 * nt * g.level;
            a.Ga = w.Ga; a.Gk = w.Gk;
            Timed timed(p, 4, st);
            RD_CUDA(launch_fwd_cluster(p, cc, a, st));
        }
        RD_CUDA(launch_gradient_epilogue(p, v, B, w.Ga, w.Gk, w.Gb, w.g_planes, w.argmin, w.fold_tmp, w.vel_part, grad_v, st));
        return RDFWI_OK;
    }
    if (!ckpt && w.split && cluster_config(p, &cc, std::min(w.u_chunk, B * g.ns))) {
        // split adjoint: per chunk of shots, (1) the cluster-resident kernel runs the adjoint field in the u-variable
        // and streams it to HBM, (2) a streaming kernel forms the imaging sums from the two histories
        const int nshots = B * g.ns;
        for (int s0 = 0; s0 < nshots; s0 += w.u_chunk) {
            const int n = std::min(w.u_chunk, nshots - s0);
            ClusterFwdArgs a{};
            a.alpha = w.alpha; a.kap = w.kap; a.beta_src = w.beta_src;
            a.isx = p.d_isx; a.rec_ptr = p.d_rec_ptr; a.rec_idx = p.d_rec_idx; a.wavelet = p.d_wavelet;
            a.nshots = n; a.nt = nt; a.st = p.st; a.shot0 = s0; a.trace = p.trace_ptr;
            if (w.recompute) {  // forward field of the chunk, again, this time keeping every level (no seismograms)
                a.seis = nullptr; a.hist = w.p_hist; a.adj_mode = 0;
                Timed timed(p, 0, st);
                RD_CUDA(launch_fwd_cluster(p, cc, a, st));
            }
            a.seis = nullptr; a.hist = w.u_hist; a.adj_mode = 1; a.cot = cot; a.Gb = w.Gb;
            {
                Timed timed(p, 1, st);
                RD_CUDA(launch_fwd_cluster(p, cc, a, st));
            }
            Timed timed(p, 2, st);
            RD_CUDA(launch_imaging(p, w.recompute ? w.p_hist : hist, w.u_hist, w.alpha, w.kap, w.beta_src, w.Gb, w.Ga, w.Gk, s0, n,
                                   w.recompute ? s0 : 0, st));
        }
        RD_CUDA(launch_gradient_epilogue(p, v, B, w.Ga, w.Gk, w.Gb, w.g_planes, w.argmin, w.fold_tmp, w.vel_part, grad_v, st));
        return RDFWI_OK;
    }
    if (p.engine == 2 && !ckpt && p.adj_mode == 0) { set_error("engine=2 (cluster-resident) requested but the grid does not fit a cluster"); return RDFWI_EINVAL; }
    RD_CUDA(cudaMemsetAsync(w.zero, 0, w.chunk_level * sizeof(float), st));
    if (!ckpt && w.split_tile) {
        // split adjoint on the per-level engine: per chunk of shots, nt tiled launches write the adjoint field u_t (slot k
        // of the scratch history = reverse level nt-1-k), then the imaging kernel streams both histories once
        const_cast<Plan &>(p).last_split = 3;
        const_cast<Plan &>(p).last_u_chunk = w.u_chunk;
        const int nshots = B * g.ns;
        const unsigned long long ustride = (unsigned long long)nt * g.level;
        for (int s0 = 0; s0 < nshots; s0 += w.u_chunk) {
            const int n = std::min(w.u_chunk, nshots - s0);
            {
                Timed timed(p, 1, st);
                for (int k = 0; k < nt; ++k) {
                    const int t = nt - 1 - k;
                    StepArgs a{};
                    a.p1 = k >= 1 ? w.u_hist + (size_t)(k - 1) * g.level : w.zero; a.ss_p1 = k >= 1 ? ustride : 0;
                    a.p0 = k >= 2 ? w.u_hist + (size_t)(k - 2) * g.level : w.zero; a.ss_p0 = k >= 2 ? ustride : 0;
                    a.out = w.u_hist + (size_t)k * g.level; a.ss_out = ustride;
                    a.alpha = w.alpha; a.kap = w.kap; a.beta_src = w.beta_src;
                    a.isx = p.d_isx; a.rec_ptr = p.d_rec_ptr; a.rec_idx = p.d_rec_idx;
                    a.cot = (t % p.st == 0) ? cot : nullptr;
                    a.it_out = t / p.st;
                    a.w_t = p.wavelet[t];
                    a.Gb = w.Gb;
                    a.shot0 = s0; a.nshots = n; a.adj = 1; a.last = k == nt - 1;
                    launch_step_tile(p, a, st);
                }
            }
            Timed timed(p, 2, st);
            RD_CUDA(launch_imaging(p, hist, w.u_hist, w.alpha, w.kap, w.beta_src, w.Gb, w.Ga, w.Gk, s0, n, 0, st));
        }
        RD_CUDA(launch_gradient_epilogue(p, v, B, w.Ga, w.Gk, w.Gb, w.g_planes, w.argmin, w.fold_tmp, w.vel_part, grad_v, st));
        return RDFWI_OK;
    }
    const int K = segment;
    const int nseg = ckpt ? num_segments(p, K) : 1;
    const size_t ck_stride = ckpt ? (size_t)(nseg - 1) * 2 * g.level : 0;
    Timed *timed_adj = new Timed(p, 3, st);
    // per-level adjoint: shots dealt over grid.z slices, one imaging plane per slice (w.g_planes may be larger when a
    // cluster configuration exists but the checkpointed path is taken; the planes actually used are pl_slices)
    const int pl_slices = std::min(w.g_planes, adj_shot_slices(p, w.nb));
    RD_CUDA(cudaMemsetAsync(w.Ga, 0, (size_t)B * w.g_planes * g.level * sizeof(float), st));
    RD_CUDA(cudaMemsetAsync(w.Gk, 0, (size_t)B * w.g_planes * g.level * sizeof(float), st));

    for (int b0 = 0; b0 < B; b0 += w.nb) {
        const int nb = std::min(w.nb, B - b0);
        const size_t lvl = (size_t)nb * g.ns * g.level;
        const size_t hstride = (size_t)nt * g.level;
        const float *hbase = hist + (size_t)b0 * g.ns * hstride;
        (void)lvl;
        RD_CUDA(cudaMemsetAsync(w.fields, 0, 3 * w.chunk_level * sizeof(float), st));  // q_{nt} = q_{nt+1} = 0
        const float *ck_base = ckpt ? hist + (size_t)b0 * g.ns * ck_stride : nullptr;
        // forward level t-1 as seen by adjoint level t: full history, or the segment's recomputed levels / checkpoints
        auto forward_level = [&](int t, int t0, unsigned long long *stride) -> const float * {
            *stride = g.level;
            if (t < 0) return w.zero;
            if (!ckpt) { *stride = hstride; return hbase + (size_t)t * g.level; }
            if (t >= t0) { *stride = (size_t)(K - 1) * g.level; return w.seg_hist + (size_t)(t - t0) * g.level; }
            const int j = t0 / K;  // t is t0-1 or t0-2: the checkpoint pair in front of segment j
            *stride = ck_stride;
            return ck_base + ((size_t)(j - 1) * 2 + (t - (t0 - 2))) * g.level;
        };
        for (int seg = nseg - 1; seg >= 0; --seg) {
            const int t0 = ckpt ? seg * K : 0, t1 = ckpt ? std::min(nt, t0 + K) : nt;
            if (ckpt) {
                // recompute p_{t0} .. p_{t1-2} from the checkpoint pair (zeros for the first segment)
                for (int t = t0; t <= t1 - 2; ++t) {
                    StepArgs f{};
                    f.p1 = forward_level(t - 1, t0, &f.ss_p1);
                    f.p0 = forward_level(t - 2, t0, &f.ss_p0);
                    f.out = const_cast<float *>(forward_level(t, t0, &f.ss_out));
                    f.alpha = w.alpha; f.kap = w.kap; f.beta_src = w.beta_src;
                    f.isx = p.d_isx; f.rec_ptr = p.d_rec_ptr; f.rec_idx = p.d_rec_idx;
                    f.seis = nullptr; f.it_out = 0;
                    f.w_t = p.wavelet[t];
                    f.shot0 = b0 * g.ns; f.nshots = nb * g.ns;
                    launch_step_tile(p, f, st);
                }
            }
            for (int t = t1 - 1; t >= t0; --t) {
                AdjArgs a;
                a.q1 = w.fields + (size_t)((t + 1) % 3) * w.chunk_level;
                a.q2 = w.fields + (size_t)((t + 2) % 3) * w.chunk_level;
                a.out = w.fields + (size_t)(t % 3) * w.chunk_level;
                a.pm1 = forward_level(t - 1, t0, &a.ss_pm1);
                a.alpha = w.alpha + (size_t)b0 * g.level;
                a.kap = w.kap + (size_t)b0 * (g.nbc + 1);
                a.isx = p.d_isx; a.rec_ptr = p.d_rec_ptr; a.rec_idx = p.d_rec_idx;
                a.cot = (t % p.st == 0) ? cot + (size_t)b0 * g.ns * g.nt_out * g.nrec : nullptr;
                a.it_out = t / p.st;
                a.w_t = p.wavelet[t];
                a.slices = pl_slices;
                a.Ga = w.Ga + (size_t)b0 * pl_slices * g.level;
                a.Gk = w.Gk + (size_t)b0 * pl_slices * g.level;
                a.Gb = w.Gb + (size_t)b0 * g.ns;
                launch_adj_step(p, a, nb, st);
            }
        }
    }
    delete timed_adj;
    // the per-level engine accumulated into pl_slices planes per model (one per grid.z slice of shots)
    RD_CUDA(launch_gradient_epilogue(p, v, B, w.Ga, w.Gk, w.Gb, pl_slices, w.argmin, w.fold_tmp, w.vel_part, grad_v, st));
    return RDFWI_OK;
}

int rdfwi_misfit_l1(rdfwi_plan plan, const float *seis, const float *observed, const float *mask, int32_t B,
                    double *stats, float *sign_out, void *ws, size_t ws_bytes, void *stream)
{
    if (!plan) { set_error("null plan"); return RDFWI_EINVAL; }
    if (!seis || !observed || !stats || B <= 0) { set_error("null seismogram / observed / stats pointer or B <= 0"); return RDFWI_EINVAL; }
    if (!ws || ws_bytes < misfit_scratch_bytes(B)) { set_error("workspace too small for the misfit partial sums"); return RDFWI_ESIZE; }
    if ((reinterpret_cast<uintptr_t>(ws) & 7) != 0) { set_error("workspace must be 8-byte aligned"); return RDFWI_EINVAL; }
    const Plan &p = *reinterpret_cast<Plan *>(plan);
    const long long n = (long long)p.g.ns * p.g.nt_out * p.g.nrec;
    if ((n & 3) == 0 && (((reinterpret_cast<uintptr_t>(seis) | reinterpret_cast<uintptr_t>(observed) | reinterpret_cast<uintptr_t>(mask) |
                           reinterpret_cast<uintptr_t>(sign_out)) & 15) != 0)) {
        set_error("seismogram buffers must be 16-byte aligned");
        return RDFWI_EINVAL;
    }
    DeviceGuard guard(p.device);
    t_launches = 0;
    RD_CUDA(launch_misfit_l1(seis, observed, mask, B, n, sign_out, stats, static_cast<double *>(ws), (cudaStream_t)stream));
    return RDFWI_OK;
}

}  // extern "C"
