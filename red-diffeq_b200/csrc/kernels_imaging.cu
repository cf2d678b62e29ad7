// kernels_imaging.cu -- zero-lag imaging condition as a pointwise streaming kernel over two wavefield histories.
//
// Second half of the "split" adjoint: after the cluster-resident kernel has written the adjoint
// field u_t = alpha*q_t of a chunk of shots to HBM (slot k holds u_{nt-1-k}), this kernel forms, per shot and cell,
//     Ga = sum_t q_t (S-5) p_{t-1}            (d L / d alpha)
//     Gk = sum_t (q_{t+1} - q_t) p_{t-1}      (d L / d kappa)
// which is what autograd accumulates for alpha / kappa from the tape of solvers/pde.py:79 (core/inversion.py:86).
//
// No stencil is evaluated here.  The forward recurrence itself says
//     alpha (S-5) p_{t-1} = p_t - (2-kappa) p_{t-1} + (1-kappa) p_{t-2} - beta w_t e_src
// and summation by parts in time (p_{-1} = p_{-2} = 0, u_nt = u_{nt+1} = 0) moves the time shifts onto u:
//     alpha^2 Ga = sum_m p_m [ u_m - (2-kappa) u_{m+1} + (1-kappa) u_{m+2} ]  -  beta_src sum_t w_t u_t[src] e_src
//     alpha   Gk = sum_m p_m [ u_{m+2} - u_{m+1} ]
// so every cell needs only its own p_m and u_m, read once each: 8 bytes per cell-update, perfectly coalesced, no
// halos, the two sums and u_{m+1}, u_{m+2} in registers for the whole time loop.  (Checked against the stencil form
// in fp64: identical to 4e-15; in fp32 both forms are equally far from the fp64 result, 1.9e-5 on the OpenFWI case.)
// This is the HBM-bound part of the adjoint.
#include "rdfwi_common.cuh"

namespace rdfwi {
namespace {

constexpr int kImgThreads = 256;

__device__ __forceinline__ float4 ldg4_stream(const float *p)
{
    float4 v;  // read once: do not keep in L1
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

__device__ __forceinline__ int sponge_index(int i, int n, int nbc)
{
    return i < nbc ? nbc - 1 - i : (i >= n - nbc ? i - (n - nbc) : -1);
}

// grid = (float4 slots of a level / 256, shots of the chunk); a thread owns one float4 of one shot for all levels
__global__ void __launch_bounds__(kImgThreads) k_imaging(const float *__restrict__ phist, const float *__restrict__ uhist,
                                                         const float *__restrict__ alpha, const float *__restrict__ kap,
                                                         const float *__restrict__ beta_src, const float *__restrict__ Gb,
                                                         const int *__restrict__ isx, float *__restrict__ Ga,
                                                         float *__restrict__ Gk, Grid g, int nt, int shot0, int pshot0, int prefetch)
{
    const int i = blockIdx.x * kImgThreads + threadIdx.x;  // float4 slot
    if (i >= g.nzp * g.q4) return;
    const int shot_l = blockIdx.y, shot = shot0 + shot_l, b = shot / g.ns;
    const int z = i / g.q4, x = (i - z * g.q4) * 4;
    const size_t cell = (size_t)i * 4;
    const size_t lvl = g.level;
    const float *pl = phist + (size_t)(shot - pshot0) * nt * lvl + (size_t)(nt - 1) * lvl + cell;  // p_m, m = nt-1 .. 0
    const float *ul = uhist + (size_t)shot_l * nt * lvl + cell;                         // u_m = slot nt-1-m

    // kappa*dt of the four cells (columns override rows, solvers/pde.py:48-51)
    float two_mk[4], one_mk[4];
    {
        const float *kap_b = kap + (size_t)b * (g.nbc + 1);
        const int kz = sponge_index(z, g.nzp, g.nbc);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int xc = x + j >= g.nxp ? x + j - g.nxp : x + j;
            const int kx = sponge_index(xc, g.nxp, g.nbc);
            const float k = kap_b[kx >= 0 ? kx : (kz >= 0 ? kz : g.nbc)];
            two_mk[j] = 2.0f - k;
            one_mk[j] = 1.0f - k;
        }
    }
    float ga[4] = {0.f, 0.f, 0.f, 0.f}, gk[4] = {0.f, 0.f, 0.f, 0.f};
    float4 u1 = make_float4(0.f, 0.f, 0.f, 0.f), u2 = u1;  // u_{m+1}, u_{m+2}

#pragma unroll 4
    for (int m = nt - 1; m >= 0; --m, pl -= lvl, ul += lvl) {
        if (m >= prefetch) {  // pull a later level's lines into L2 (one request per 128-byte line)
            if ((threadIdx.x & 7) == 0) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pl - (size_t)prefetch * lvl));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ul + (size_t)prefetch * lvl));
            }
        }
        const float4 p = ldg4_stream(pl);
        const float4 u = ldg4_stream(ul);
        ga[0] += p.x * ((u.x - two_mk[0] * u1.x) + one_mk[0] * u2.x);
        ga[1] += p.y * ((u.y - two_mk[1] * u1.y) + one_mk[1] * u2.y);
        ga[2] += p.z * ((u.z - two_mk[2] * u1.z) + one_mk[2] * u2.z);
        ga[3] += p.w * ((u.w - two_mk[3] * u1.w) + one_mk[3] * u2.w);
        gk[0] += p.x * (u2.x - u1.x);
        gk[1] += p.y * (u2.y - u1.y);
        gk[2] += p.z * (u2.z - u1.z);
        gk[3] += p.w * (u2.w - u1.w);
        u2 = u1;
        u1 = u;
    }

    const float4 al = *reinterpret_cast<const float4 *>(alpha + (size_t)b * lvl + cell);
    const float a4[4] = {al.x, al.y, al.z, al.w};
    // source cell: the forward recurrence had the extra term beta_src w_t there (Gb holds sum_t w_t u_t[src] / alpha_src)
    if (z == g.isz) {
        const int xs = isx[shot - b * g.ns];
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (x + j == xs && x + j < g.nxp) ga[j] -= beta_src[shot] * (Gb[shot] * a4[j]);
    }
    const size_t off = (size_t)shot * lvl + cell;
    *reinterpret_cast<float4 *>(Ga + off) = make_float4(ga[0] / (a4[0] * a4[0]), ga[1] / (a4[1] * a4[1]), ga[2] / (a4[2] * a4[2]), ga[3] / (a4[3] * a4[3]));
    *reinterpret_cast<float4 *>(Gk + off) = make_float4(gk[0] / a4[0], gk[1] / a4[1], gk[2] / a4[2], gk[3] / a4[3]);
}

}  // namespace

// phist holds the forward history of shots pshot0, pshot0+1, ... (0 = the whole batch's history; shot0 = a chunk-local
// history recomputed in the backward pass)
cudaError_t launch_imaging(const Plan &p, const float *phist, const float *uhist, const float *alpha, const float *kap,
                           const float *beta_src, const float *Gb, float *Ga, float *Gk, int shot0, int nshots, int pshot0,
                           cudaStream_t st)
{
    const Grid &g = p.g;
    const int slots = g.nzp * g.q4;
    const dim3 grid((slots + kImgThreads - 1) / kImgThreads, nshots);
    const int pf = p.img_prefetch > 0 ? p.img_prefetch : 4;
    k_imaging<<<grid, kImgThreads, 0, st>>>(phist, uhist, alpha, kap, beta_src, Gb, p.d_isx, Ga, Gk, g, p.nt, shot0, pshot0, pf);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rdfwi
