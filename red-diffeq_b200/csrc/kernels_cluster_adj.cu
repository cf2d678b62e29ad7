// kernels_cluster_adj.cu -- cluster-resident reverse-time loop: adjoint field + zero-lag imaging sums.
//
// Replaces what PyTorch autograd replays from the tape of solvers/pde.py:78-85 (core/inversion.py:86).
// Formulation ("adjoint in the u-variable"): with u_t = alpha * q_t the adjoint recurrence
//     q_t = T1 q_{t+1} + S(alpha q_{t+1}) - T2 q_{t+2} + R^T g_t                       
// becomes
//     u_t = T1 u_{t+1} + alpha S(u_{t+1}) - T2 u_{t+2} + alpha R^T g_t
// which is the *forward* recurrence with another source term, so the same shared-memory sweep applies
// (stencil on the stored field itself, no alpha needed at neighbour cells).  The imaging sums become
//     Ga = (1/alpha) sum_t u_t (S-5) p_{t-1},   Gk = (1/alpha) sum_t (u_{t+1}-u_t) p_{t-1},
//     Gb = (1/alpha_src) sum_t u_t[src] w_t
// and stay in registers for the whole time loop (a thread owns the same cells at every level).
//
// Shared memory per CTA: two u slabs (with 2+2 halo rows, exchanged through DSMEM as in the forward
// kernel), one p slab (forward level t-1 with halos, streamed from the HBM history by 1-D bulk copies
// completing on an mbarrier, issued one level ahead), one alpha slab.
#include "cluster_ptx.cuh"

namespace rdfwi {
namespace {

// u_t over the thread's rows: same structure as fwd_sweep (kernels_cluster.cu) including the early halo
// sends, but FMA contraction is allowed (no bit-parity requirement on the adjoint) and alpha is streamed
// from global memory (L2-resident: one plane per model), because the imaging sums own the registers.
template <int RMAX, int PITCH, int DIR>
__device__ __forceinline__ void adj_sweep(float *__restrict__ smem, const int cur, const int prv, const int kap_off,
                                          const int pitch_rt, const int l0, const SweepThread &th,
                                          const float *__restrict__ alpha_row0, const float (&kapx)[4], const HaloPush &hp,
                                          const uint64_t *push_bar, const bool push_now)
{
    const int pitch = PITCH > 0 ? PITCH : pitch_rt;
    const int P = DIR * pitch;
    const float c2 = 4.0f / 3.0f, c3 = -1.0f / 12.0f;
    const float *cb = smem + cur + (2 + l0) * pitch + th.x;
    float *pb = smem + prv + (2 + l0) * pitch + th.x;
    const float *eLp = smem + cur + (2 + l0) * pitch + th.eL;
    const float *eRp = smem + cur + (2 + l0) * pitch + th.eR;
    const float *kz = smem + kap_off + l0;
    const float *push_dst = smem + prv + hp.dst + th.x;
    const float *ab = alpha_row0 + (long)l0 * pitch + th.x;  // alpha of row l0 (global)

    float4 alv[RMAX];
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        const int lr = l0 + DIR * r;
        alv[r] = (lr >= th.la && lr < th.lb) ? __ldg(reinterpret_cast<const float4 *>(ab + r * P)) : make_float4(1.f, 1.f, 1.f, 1.f);
    }
    float4 w0 = ld4(cb - 2 * P), w1 = ld4(cb - P), w2 = ld4(cb), w3 = ld4(cb + P);
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        const float4 w4 = ld4(cb + (r + 2) * P);
        const float4 old = ld4(pb + r * P);
        const float kapz = kz[DIR * r];
        float l2 = __shfl_up_sync(0xffffffffu, w2.z, 1);
        float l1 = __shfl_up_sync(0xffffffffu, w2.w, 1);
        float r0 = __shfl_down_sync(0xffffffffu, w2.x, 1);
        float r1 = __shfl_down_sync(0xffffffffu, w2.y, 1);
        if (th.edgeL) { l2 = eLp[r * P]; l1 = eLp[r * P + 1]; }
        if (th.edgeR) { r0 = eRp[r * P]; r1 = eRp[r * P + 1]; }
        const float e[8] = {l2, l1, w2.x, w2.y, w2.z, w2.w, r0, r1};
        float o[4];
        // two cells per instruction (FADD2 / FMUL2 / FFMA2); contraction is welcome here
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 2 * h;
            const float2 up1 = h ? make_float2(w1.z, w1.w) : make_float2(w1.x, w1.y);
            const float2 dn1 = h ? make_float2(w3.z, w3.w) : make_float2(w3.x, w3.y);
            const float2 up2 = h ? make_float2(w0.z, w0.w) : make_float2(w0.x, w0.y);
            const float2 dn2 = h ? make_float2(w4.z, w4.w) : make_float2(w4.x, w4.y);
            const float2 oldp = h ? make_float2(old.z, old.w) : make_float2(old.x, old.y);
            const float2 alp = h ? make_float2(alv[r].z, alv[r].w) : make_float2(alv[r].x, alv[r].y);
            const float2 cen = make_float2(e[j + 2], e[j + 3]);
            const float2 s1 = f2add(f2add(f2add(up1, dn1), make_float2(e[j + 1], e[j + 2])), make_float2(e[j + 3], e[j + 4]));
            const float2 s2 = f2add(f2add(f2add(up2, dn2), make_float2(e[j], e[j + 1])), make_float2(e[j + 4], e[j + 5]));
            const float2 lap = f2fma(make_float2(c2, c2), s1, f2mul(make_float2(c3, c3), s2));
            const float2 kp = make_float2(th.colsp[j] ? kapx[j] : kapz, th.colsp[j + 1] ? kapx[j + 1] : kapz);
            const float2 t1 = f2sub(f2fma(make_float2(-5.0f, -5.0f), alp, make_float2(2.0f, 2.0f)), kp);
            const float2 t2 = f2sub(make_float2(1.0f, 1.0f), kp);
            const float2 res = f2fma(alp, lap, f2sub(f2mul(t1, cen), f2mul(t2, oldp)));
            o[j] = res.x; o[j + 1] = res.y;
        }
        const int lr = l0 + DIR * r;
        const float4 out = make_float4(o[0], o[1], o[2], o[3]);
        if (lr >= th.la && lr < th.lb) st4(pb + r * P, out);
        if (r < 2 && push_now) st_async_v4(push_dst + r * P, push_bar, hp.cta, out);
        w0 = w1; w1 = w2; w2 = w3; w3 = w4;
    }
}

// imaging sums of one level, rows in the same marching order as the sweep:
//     Ga += u_t (S-5) p_{t-1},  Gk += (u_{t+1} - u_t) p_{t-1}
template <int RMAX, int PITCH, int DIR>
__device__ __forceinline__ void imaging_sweep(const float *__restrict__ smem, const int ucur, const int uprv, const int pbuf,
                                              const int pitch_rt, const int l0, const SweepThread &th, float4 (&Ga)[RMAX],
                                              float4 (&Gk)[RMAX])
{
    const int pitch = PITCH > 0 ? PITCH : pitch_rt;
    const int P = DIR * pitch;
    const float c2 = 4.0f / 3.0f, c3 = -1.0f / 12.0f;
    const float *pb = smem + pbuf + (2 + l0) * pitch + th.x;       // p_{t-1}, row l0
    const float *u1b = smem + ucur + (2 + l0) * pitch + th.x;      // u_{t+1}
    const float *utb = smem + uprv + (2 + l0) * pitch + th.x;      // u_t (just written by this thread)
    const float *eLp = smem + pbuf + (2 + l0) * pitch + th.eL;
    const float *eRp = smem + pbuf + (2 + l0) * pitch + th.eR;

    float4 v0 = ld4(pb - 2 * P), v1 = ld4(pb - P), v2 = ld4(pb), v3 = ld4(pb + P);
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        const float4 v4 = ld4(pb + (r + 2) * P);
        const float4 ut = ld4(utb + r * P);
        const float4 u1 = ld4(u1b + r * P);
        float l2 = __shfl_up_sync(0xffffffffu, v2.z, 1);
        float l1 = __shfl_up_sync(0xffffffffu, v2.w, 1);
        float r0 = __shfl_down_sync(0xffffffffu, v2.x, 1);
        float r1 = __shfl_down_sync(0xffffffffu, v2.y, 1);
        if (th.edgeL) { l2 = eLp[r * P]; l1 = eLp[r * P + 1]; }
        if (th.edgeR) { r0 = eRp[r * P]; r1 = eRp[r * P + 1]; }
        const float e[8] = {l2, l1, v2.x, v2.y, v2.z, v2.w, r0, r1};
        float ga[4] = {Ga[r].x, Ga[r].y, Ga[r].z, Ga[r].w};
        float gk[4] = {Gk[r].x, Gk[r].y, Gk[r].z, Gk[r].w};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 2 * h;
            const float2 up1 = h ? make_float2(v1.z, v1.w) : make_float2(v1.x, v1.y);
            const float2 dn1 = h ? make_float2(v3.z, v3.w) : make_float2(v3.x, v3.y);
            const float2 up2 = h ? make_float2(v0.z, v0.w) : make_float2(v0.x, v0.y);
            const float2 dn2 = h ? make_float2(v4.z, v4.w) : make_float2(v4.x, v4.y);
            const float2 utp = h ? make_float2(ut.z, ut.w) : make_float2(ut.x, ut.y);
            const float2 u1p = h ? make_float2(u1.z, u1.w) : make_float2(u1.x, u1.y);
            const float2 pc = make_float2(e[j + 2], e[j + 3]);
            const float2 s1 = f2add(f2add(f2add(up1, dn1), make_float2(e[j + 1], e[j + 2])), make_float2(e[j + 3], e[j + 4]));
            const float2 s2 = f2add(f2add(f2add(up2, dn2), make_float2(e[j], e[j + 1])), make_float2(e[j + 4], e[j + 5]));
            const float2 lp = f2fma(make_float2(c2, c2), s1, f2fma(make_float2(c3, c3), s2, f2mul(make_float2(-5.0f, -5.0f), pc)));
            const float2 gan = f2fma(utp, lp, make_float2(ga[j], ga[j + 1]));
            const float2 gkn = f2fma(f2sub(u1p, utp), pc, make_float2(gk[j], gk[j + 1]));
            ga[j] = gan.x; ga[j + 1] = gan.y;
            gk[j] = gkn.x; gk[j + 1] = gkn.y;
        }
        Ga[r] = make_float4(ga[0], ga[1], ga[2], ga[3]);
        Gk[r] = make_float4(gk[0], gk[1], gk[2], gk[3]);
        v0 = v1; v1 = v2; v2 = v3; v3 = v4;
    }
}

template <int RMAX, int PITCH>
__global__ void __launch_bounds__(kClusterThreads, 1) k_adj_cluster(ClusterAdjArgs a, Grid g)
{
    extern __shared__ __align__(128) float smem[];

    const int C = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / C, ncl = gridDim.x / C;
    const int base = g.nzp / C, rem = g.nzp % C;
    const int nrows = base + (rank < rem ? 1 : 0);
    const int r0 = rank * base + (rank < rem ? rank : rem);
    const int up = rank == 0 ? C - 1 : rank - 1;
    const int dn = rank == C - 1 ? 0 : rank + 1;
    const int nrows_up = base + (up < rem ? 1 : 0);

    const int pitch = PITCH > 0 ? PITCH : g.pitch;
    const int slab = (a.slabrows + 4) * pitch;   // u slabs: 2 halo rows, slab rows, 2 halo rows
    const int pslab0 = 2 * slab;                 // two forward-history slabs (levels t-1 / t-2 in flight)
    const int kap_off = 4 * slab;
    // mbarriers: [0..3] u halos [buffer][top|bottom], [4..5] history slabs
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + ((kap_off + a.slabrows + 3) & ~3));
    int *s_rec_ptr = reinterpret_cast<int *>(bars + 6);
    int *s_rec_idx = s_rec_ptr + g.nxp + 1;
    float *s_wav = reinterpret_cast<float *>(s_rec_idx + g.nrec);
    const bool wav_in_smem = a.wav_smem != 0;

    const int tid = threadIdx.x, lane_id = tid & 31;
    const int grp = tid / g.q4, col = tid - grp * g.q4;
    const bool active = grp < a.ngroups;
    const bool warp_active = (tid - lane_id) < a.ngroups * g.q4;
    SweepThread th;
    th.x = col * 4;
    th.la = grp * RMAX;
    th.lb = !active ? th.la : (th.la + RMAX < nrows ? th.la + RMAX : nrows);
    th.edgeL = lane_id == 0 || col == 0;
    th.edgeR = lane_id == 31 || col == g.q4 - 1;
    th.eL = col == 0 ? g.nxp - 2 : th.x - 2;
    th.eR = col == g.q4 - 1 ? pitch - g.nxp : th.x + 4;
    int xc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        xc[j] = th.x + j >= g.nxp ? th.x + j - g.nxp : th.x + j;
        th.colsp[j] = sponge_index(xc[j], g.nxp, g.nbc) >= 0;
    }
    th.src_lr = (g.isz - r0 >= th.la && g.isz - r0 < th.lb) ? g.isz - r0 : -1;
    th.rec_lr = (g.igz - r0 >= th.la && g.igz - r0 < th.lb) ? g.igz - r0 : -1;
    // marching directions, halo reads and early sends: as in k_fwd_cluster
    const int last_grp = (nrows - 1) / RMAX;
    const bool rev = __any_sync(0xffffffffu, active && grp == last_grp && last_grp > 0) != 0;
    const int la_eff = active ? th.la : last_grp * RMAX;
    const int lb_eff = active ? th.lb : nrows;
    const int l0 = rev ? (lb_eff > la_eff ? lb_eff - 1 : la_eff) : la_eff;
    th.lac = l0;
    bool rd_top, rd_bot;
    {
        const int lo = rev ? l0 - (RMAX + 1) : l0 - 2, hi = rev ? l0 + 2 : l0 + RMAX + 1;
        rd_top = __any_sync(0xffffffffu, lo < 0) != 0;
        rd_bot = __any_sync(0xffffffffu, hi >= nrows && lo < nrows + 2) != 0;
    }
    // the receiver row is patched in the epilogue (cotangent injection), so it cannot be sent early
    const bool rec_on_edge = g.igz - r0 >= 0 && g.igz - r0 < nrows && (g.igz - r0 < 2 || g.igz - r0 >= nrows - 2);
    HaloPush hp;
    hp.early = false; hp.dst = 0; hp.cta = 0; hp.bar = 0;
    if (active && !rec_on_edge) {
        if (!rev && th.la == 0 && th.lb >= 2) {
            hp.early = true; hp.dst = (2 + nrows_up) * pitch; hp.cta = (uint32_t)up; hp.bar = 1;
        } else if (rev && th.lb == nrows && th.lb - th.la >= 2) {
            hp.early = true; hp.dst = pitch; hp.cta = (uint32_t)dn; hp.bar = 0;
        }
    }
    int late = 0;
    for (int h = 0; h < 2; ++h) {
        if (h >= th.la && h < th.lb && !(hp.early && hp.bar == 1)) late |= 1 << h;
        const int rb = nrows - 2 + h;
        if (rb >= th.la && rb < th.lb && !(hp.early && hp.bar == 0)) late |= 4 << h;
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t hist_shot = (size_t)a.nt * g.level;  // the history keeps every level (the last one is not read here)
    const uint32_t halo_bytes = (uint32_t)(2 * pitch * sizeof(float));
    // u-halo phases consumed per shot: reverse level index k = nt-1-t; buffer roles alternate with k
    const int uses1 = a.nt / 2, uses0 = (a.nt - 1) / 2;
    // history loads per shot: levels nt-2 .. 0 -> slab (lvl & 1)
    const int loads1 = (a.nt - 1) / 2, loads0 = a.nt / 2;

    if (tid == 0) {
        for (int i = 0; i < 6; ++i) mbar_init(bars + i, 1);
    }
    for (int i = tid; i <= g.nxp; i += kClusterThreads) s_rec_ptr[i] = a.rec_ptr[i];
    for (int i = tid; i < g.nrec; i += kClusterThreads) s_rec_idx[i] = a.rec_idx[i];
    if (wav_in_smem)
        for (int i = tid; i < a.nt; i += kClusterThreads) s_wav[i] = a.wavelet[i];
    __syncthreads();

    // stream forward level `lvl` (rows r0-2 .. r0+nrows+1, periodic in z) of `shot` into history slab (lvl & 1)
    auto load_level = [&](const int shot, const int lvl) {
        const float *src = a.hist + (size_t)shot * hist_shot + (size_t)lvl * g.level;
        float *dst = smem + pslab0 + (lvl & 1) * slab;
        uint64_t *bar = bars + 4 + (lvl & 1);
        const uint32_t row_b = (uint32_t)(pitch * sizeof(float));
        const int top = r0 - 2 < 0 ? r0 - 2 + g.nzp : r0 - 2;                      // first halo row (wraps for rank 0)
        const int bot = r0 + nrows >= g.nzp ? r0 + nrows - g.nzp : r0 + nrows;     // first row below the slab
        mbar_expect_tx(bar, (uint32_t)(nrows + 4) * row_b);
        bulk_load(dst, src + (size_t)top * pitch, 2 * row_b, bar);
        bulk_load(dst + 2 * pitch, src + (size_t)r0 * pitch, (uint32_t)nrows * row_b, bar);
        bulk_load(dst + (2 + nrows) * pitch, src + (size_t)bot * pitch, 2 * row_b, bar);
    };

    int shot_iter = 0;
    for (int shot = cid; shot < a.nshots; shot += ncl, ++shot_iter) {
        const int b = shot / g.ns, s = shot - b * g.ns;
        // u_{nt} = u_{nt+1} = 0; sponge tables of this model
        for (int i = tid; i < 2 * slab; i += kClusterThreads) smem[i] = 0.0f;
        const float *kap_b = a.kap + (size_t)b * (g.nbc + 1);
        const float *alpha_row0 = a.alpha + (size_t)b * g.level + (size_t)r0 * pitch;  // alpha of the CTA's row 0
        for (int i = tid; i < a.slabrows; i += kClusterThreads) {
            const int kz = sponge_index(r0 + i, g.nzp, g.nbc);
            smem[kap_off + i] = (i < nrows && kz >= 0) ? kap_b[kz] : 0.0f;
        }
        float kapx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kx = sponge_index(xc[j], g.nxp, g.nbc);
            kapx[j] = kx >= 0 ? kap_b[kx] : 0.0f;
        }
        const int xs = a.isx[s];
        int src_lane = -1;  // lane of the (non-image) source cell if this thread owns it
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (th.src_lr >= 0 && th.x + j < g.nxp && th.x + j == xs) src_lane = j;
        float4 Ga[RMAX], Gk[RMAX];  // in marching order
#pragma unroll
        for (int r = 0; r < RMAX; ++r) { Ga[r] = zero4; Gk[r] = zero4; }
        float gb = 0.0f;
        fence_proxy_async();
        __syncthreads();
        cluster_sync_all();  // shot boundary
        if (tid == 0) {      // history prefetch runs two levels ahead
            if (a.nt >= 2) load_level(shot, a.nt - 2);
            if (a.nt >= 3) load_level(shot, a.nt - 3);
        }

        // one reverse level: u_{t+1} in buffer `cur`, u_{t+2} in `prv`, u_t overwrites u_{t+2}
        auto level = [&](const int t, const int cur, const int prv) {
            const int k = a.nt - 1 - t;  // reverse level counter, 0-based
            const int cbuf = cur == 0 ? 0 : 1, pbuf = 1 - cbuf;
            uint64_t *bar_top = bars + 2 * cbuf, *bar_bot = bars + 2 * cbuf + 1;
            const bool sends = t > 0;  // u_0 has no consumer
            if (tid == 0 && sends) {   // arm the barriers of the buffer written now, before the neighbours' rows can land
                mbar_expect_tx(bars + 2 * pbuf, halo_bytes);
                mbar_expect_tx(bars + 2 * pbuf + 1, halo_bytes);
            }
            if (k >= 1) {
                const uint32_t parity = (uint32_t)((shot_iter * (cbuf ? uses1 : uses0) + (k - 1) / 2) & 1);
                if (warp_active) {
                    if (rd_top) mbar_wait(bar_top, parity);
                    if (rd_bot) mbar_wait(bar_bot, parity);
                }
            }
            const int p0 = prv + 2 * pitch + th.x;
            if (warp_active) {
                const uint64_t *push_bar = bars + 2 * pbuf + hp.bar;
                if (rev) adj_sweep<RMAX, PITCH, -1>(smem, cur, prv, kap_off, pitch, l0, th, alpha_row0, kapx, hp, push_bar, hp.early && sends);
                else adj_sweep<RMAX, PITCH, 1>(smem, cur, prv, kap_off, pitch, l0, th, alpha_row0, kapx, hp, push_bar, hp.early && sends);
                if (th.rec_lr >= 0 && t % a.st == 0) {  // u_t[rec] += alpha * g_t  (adjoint of the gather, :83)
                    const float *gt = a.cot + ((size_t)shot * g.nt_out + t / a.st) * g.nrec;
                    float4 v = ld4(smem + p0 + th.rec_lr * pitch);
                    const float4 alv = __ldg(reinterpret_cast<const float4 *>(alpha_row0 + (size_t)th.rec_lr * pitch + th.x));
                    float add[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        for (int q = s_rec_ptr[xc[j]]; q < s_rec_ptr[xc[j] + 1]; ++q) add[j] += gt[s_rec_idx[q]];
                    v.x += alv.x * add[0]; v.y += alv.y * add[1]; v.z += alv.z * add[2]; v.w += alv.w * add[3];
                    st4(smem + p0 + th.rec_lr * pitch, v);
                }
                if (src_lane >= 0) {  // adjoint of the source injection (:81)
                    const float4 v = ld4(smem + p0 + th.src_lr * pitch);
                    gb += (src_lane == 0 ? v.x : src_lane == 1 ? v.y : src_lane == 2 ? v.z : v.w) * (wav_in_smem ? s_wav[t] : a.wavelet[t]);
                }
                if (late != 0 && sends) {
                    for (int h = 0; h < 2; ++h) {
                        if (late & (1 << h))
                            st_async_v4(smem + prv + (2 + nrows_up + h) * pitch + th.x, bars + 2 * pbuf + 1, (uint32_t)up,
                                        ld4(smem + p0 + h * pitch));
                        if (late & (4 << h))
                            st_async_v4(smem + prv + h * pitch + th.x, bars + 2 * pbuf, (uint32_t)dn,
                                        ld4(smem + p0 + (nrows - 2 + h) * pitch));
                    }
                }
            }
            if (t >= 1) {
                // forward level t-1 sits in history slab ((t-1) & 1); its load was issued two levels ago
                const int hs = (t - 1) & 1;
                const int nth = (a.nt - 2 - (t - 1)) / 2;  // how many loads into this slab preceded it in this shot
                const uint32_t hpar = (uint32_t)((shot_iter * (hs ? loads1 : loads0) + nth) & 1);
                mbar_wait(bars + 4 + hs, hpar);
                if (warp_active) {
                    if (rev) imaging_sweep<RMAX, PITCH, -1>(smem, cur, prv, pslab0 + hs * slab, pitch, l0, th, Ga, Gk);
                    else imaging_sweep<RMAX, PITCH, 1>(smem, cur, prv, pslab0 + hs * slab, pitch, l0, th, Ga, Gk);
                }
            }
            __syncthreads();  // level t complete CTA-wide; history slab ((t-1)&1) is free again
            if (tid == 0 && t >= 3) load_level(shot, t - 3);
        };
        for (int t = a.nt - 1, cur = 0; t >= 0; --t, cur = slab - cur) level(t, cur, slab - cur);

        // per-shot imaging planes (divided by alpha, see header) -- summed over shots by the epilogue kernels
        if (active) {
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                const int lr = rev ? l0 - r : l0 + r;
                if (lr >= th.la && lr < th.lb) {
                    const size_t off = (size_t)(r0 + lr) * pitch + th.x;
                    const float4 alv = __ldg(reinterpret_cast<const float4 *>(a.alpha + (size_t)b * g.level + off));
                    st4(a.Ga + (size_t)shot * g.level + off, make_float4(Ga[r].x / alv.x, Ga[r].y / alv.y, Ga[r].z / alv.z, Ga[r].w / alv.w));
                    st4(a.Gk + (size_t)shot * g.level + off, make_float4(Gk[r].x / alv.x, Gk[r].y / alv.y, Gk[r].z / alv.z, Gk[r].w / alv.w));
                }
            }
            if (src_lane >= 0) {
                const float4 alv = __ldg(reinterpret_cast<const float4 *>(alpha_row0 + (size_t)th.src_lr * pitch + th.x));
                a.Gb[shot] = gb / (src_lane == 0 ? alv.x : src_lane == 1 ? alv.y : src_lane == 2 ? alv.z : alv.w);
            }
        }
        __syncthreads();
    }
    cluster_sync_all();  // no CTA exits while a neighbour may still address its shared memory
}

template <int RMAX>
bool adj_config_rmax(const Plan &p, int max_smem, ClusterConfig *cfg)
{
    const Grid &g = p.g;
    const int groups_max = kClusterThreads / g.q4;
    if (groups_max < 1) return false;
    for (int C = 1; C <= 16; ++C) {
        if (C > 8 && C != 16) continue;  // 1..8 are portable cluster sizes, 16 needs the non-portable opt-in
        if (p.adj_cluster_size > 0 && C != p.adj_cluster_size) continue;
        if (g.nzp / C < 2) break;
        const int maxrows = (g.nzp + C - 1) / C;
        const int ngroups = (maxrows + RMAX - 1) / RMAX;
        if (ngroups > groups_max) continue;
        const int slabrows = ngroups * RMAX;
        size_t smem = ((size_t)4 * (slabrows + 4) * g.pitch + slabrows + 24 + g.nxp + 1 + g.nrec) * sizeof(float);
        if (smem > (size_t)max_smem) continue;
        cfg->wav_smem = smem + (size_t)p.nt * sizeof(float) <= (size_t)max_smem;
        if (cfg->wav_smem) smem += (size_t)p.nt * sizeof(float);
        cfg->C = C; cfg->maxrows = maxrows; cfg->ngroups = ngroups; cfg->slabrows = slabrows; cfg->smem = smem; cfg->rmax = RMAX;
        return true;
    }
    return false;
}

template <int RMAX, int PITCH>
cudaError_t launch_adj_cluster_t(const Plan &p, const ClusterConfig &cc, ClusterAdjArgs a, cudaStream_t st)
{
    auto kernel = k_adj_cluster<RMAX, PITCH>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cc.smem);
    if (e != cudaSuccess) return e;
    a.slabrows = cc.slabrows; a.ngroups = cc.ngroups; a.wav_smem = cc.wav_smem ? 1 : 0;
    if (cc.C > 8) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cc.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(kClusterThreads);
    cfg.dynamicSmemBytes = cc.smem;
    cfg.stream = st;
    int sms = 148;
    device_attr(&sms, cudaDevAttrMultiProcessorCount, p.device);
    cfg.gridDim = dim3((unsigned)(sms / cc.C * cc.C));
    int max_clusters = 0;
    e = cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &cfg);
    if (e != cudaSuccess) return e;
    if (max_clusters < 1) return cudaErrorLaunchOutOfResources;
    const int ncl = max_clusters < a.nshots ? max_clusters : a.nshots;
    cfg.gridDim = dim3((unsigned)(ncl * cc.C));
    e = cudaLaunchKernelEx(&cfg, kernel, a, p.g);
    count_launch();
    return e;
}

}  // namespace

// Largest rows-per-thread (fewest idle threads first) whose slabs fit the shared memory of some cluster size.
bool adj_cluster_config(const Plan &p, ClusterConfig *cfg)
{
    int max_smem = 0;
    if (device_attr(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p.device) != cudaSuccess) return false;
    ClusterConfig best;
    bool found = false;
    ClusterConfig c;
    // smallest cluster first (most work per synchronisation), then the most active threads
    auto consider = [&](bool ok) {
        if (ok && (!found || c.C < best.C || (c.C == best.C && c.ngroups > best.ngroups))) { best = c; found = true; }
    };
    consider(adj_config_rmax<7>(p, max_smem, &c));
    consider(adj_config_rmax<5>(p, max_smem, &c));
    consider(adj_config_rmax<10>(p, max_smem, &c));
    consider(adj_config_rmax<4>(p, max_smem, &c));
    if (found) *cfg = best;
    return found;
}

cudaError_t launch_adj_cluster(const Plan &p, const ClusterConfig &cc, ClusterAdjArgs a, cudaStream_t st)
{
#define RD_DISPATCH(R)                                                               \
    switch (p.g.pitch) {                                                             \
        case 312: return launch_adj_cluster_t<R, 312>(p, cc, a, st);                 \
        case 432: return launch_adj_cluster_t<R, 432>(p, cc, a, st);                 \
        default: return launch_adj_cluster_t<R, 0>(p, cc, a, st);                    \
    }
    switch (cc.rmax) {
        case 4: RD_DISPATCH(4)
        case 5: RD_DISPATCH(5)
        case 7: RD_DISPATCH(7)
        default: RD_DISPATCH(10)
    }
#undef RD_DISPATCH
}

}  // namespace rdfwi
