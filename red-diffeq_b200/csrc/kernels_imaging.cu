// kernels_imaging.cu -- zero-lag imaging condition as a streaming kernel over two wavefield histories.
//
// Second half of the "split" adjoint: after the cluster-resident kernel has written the adjoint
// field u_t = alpha*q_t of a chunk of shots to HBM (slot k holds u_{nt-1-k}), this kernel forms, per shot and cell,
//     Ga = (1/alpha) sum_{t=1}^{nt-1} u_t (S-5) p_{t-1}            (d L / d alpha)
//     Gk = (1/alpha) sum_{t=1}^{nt-1} (u_{t+1} - u_t) p_{t-1}      (d L / d kappa; u_nt = 0)
// which is what autograd accumulates for alpha / kappa from the tape of solvers/pde.py:79 (core/inversion.py:86).
// Pure streaming: every level of both histories is read once (8 B per cell-update + tile halos), the sums live in
// registers for the whole time loop -- this is the HBM-bound part of the adjoint.
//
// Work decomposition: a CTA owns a tile of kTileRows x (kTileCols*4) cells of one shot for all levels; a thread owns
// a float4 x kRows rows.  z-neighbours are float4 loads of the rows above / below (L1 hits inside the tile),
// x-neighbours 4 scalar loads; indices wrap periodically like torch.roll.
#include "rdfwi_common.cuh"

namespace rdfwi {
namespace {

constexpr int kImgCols = 8;    // float4 columns per CTA  (32 cells: one 128-byte line per row)
constexpr int kImgRowGroups = 32;
constexpr int kImgRows = 1;    // rows per thread
constexpr int kImgPrefetch = 4;  // levels ahead that are pulled into L2
constexpr int kImgTileRows = kImgRowGroups * kImgRows;

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float lane(const float4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

__global__ void __launch_bounds__(kImgCols *kImgRowGroups, 3) k_imaging(const float *__restrict__ phist, const float *__restrict__ uhist,
                                                                    const float *__restrict__ alpha, float *__restrict__ Ga,
                                                                    float *__restrict__ Gk, Grid g, int nt, int shot0)
{
    const int col = blockIdx.x * kImgCols + (threadIdx.x % kImgCols);
    const int z0 = blockIdx.y * kImgTileRows + (threadIdx.x / kImgCols) * kImgRows;
    const int shot_l = blockIdx.z, shot = shot0 + shot_l;
    if (col >= g.q4 || z0 >= g.nzp) return;
    const int x = col * 4;
    const int xm2 = x - 2 < 0 ? x - 2 + g.nxp : x - 2, xm1 = x - 1 < 0 ? x - 1 + g.nxp : x - 1;
    const int xp4 = x + 4 >= g.nxp ? x + 4 - g.nxp : x + 4, xp5 = x + 5 >= g.nxp ? x + 5 - g.nxp : x + 5;
    int roff[kImgRows + 4];
#pragma unroll
    for (int k = 0; k < kImgRows + 4; ++k) {
        int z = z0 - 2 + k;
        z = z < 0 ? z + g.nzp : (z >= g.nzp ? z - g.nzp : z);
        roff[k] = z * g.pitch;
    }
    const size_t hshot = (size_t)(nt - 1) * g.level;
    const float *P = phist + (size_t)shot * hshot;      // p_t at P + t*level
    const float *U = uhist + (size_t)shot_l * hshot;    // u_t at U + (nt-1-t)*level, t >= 1
    const float c2 = 4.0f / 3.0f, c3 = -1.0f / 12.0f;

    float ga[kImgRows][4], gk[kImgRows][4];
    float4 unext[kImgRows];  // u_{t+1} of the owned cells
#pragma unroll
    for (int r = 0; r < kImgRows; ++r) {
        unext[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) { ga[r][j] = 0.f; gk[r][j] = 0.f; }
    }

    // one lane per 128-byte row segment of the tile pulls the lines of a later level into L2, so that the demand loads
    // below see L2 latency instead of HBM latency (bytes in flight without spending registers)
    const bool prefetcher = (threadIdx.x % kImgCols) == 0 && z0 < g.nzp;
    const int pf_off = roff[2] + x;
#pragma unroll 2
    for (int t = nt - 1; t >= 1; --t) {
        const float *pl = P + (size_t)(t - 1) * g.level;
        const float *ul = U + (size_t)(nt - 1 - t) * g.level;
        if (prefetcher && t - kImgPrefetch >= 1) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(pl - (size_t)kImgPrefetch * g.level + pf_off));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(ul + (size_t)kImgPrefetch * g.level + pf_off));
        }
        float4 rows[kImgRows + 4];
#pragma unroll
        for (int k = 0; k < kImgRows + 4; ++k) rows[k] = ldg4(pl + roff[k] + x);
#pragma unroll
        for (int r = 0; r < kImgRows; ++r) {
            if (z0 + r < g.nzp) {
                const float4 ut = ldg4(ul + roff[r + 2] + x);
                const float e[8] = {__ldg(pl + roff[r + 2] + xm2), __ldg(pl + roff[r + 2] + xm1), rows[r + 2].x, rows[r + 2].y,
                                    rows[r + 2].z, rows[r + 2].w, __ldg(pl + roff[r + 2] + xp4), __ldg(pl + roff[r + 2] + xp5)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float pc = e[j + 2];
                    const float s1 = ((lane(rows[r + 1], j) + lane(rows[r + 3], j)) + e[j + 1]) + e[j + 3];
                    const float s2 = ((lane(rows[r], j) + lane(rows[r + 4], j)) + e[j]) + e[j + 4];
                    const float lp = (c2 * s1 + c3 * s2) - 5.0f * pc;
                    const float uj = lane(ut, j);
                    ga[r][j] += uj * lp;
                    gk[r][j] += (lane(unext[r], j) - uj) * pc;
                }
                unext[r] = ut;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kImgRows; ++r) {
        if (z0 + r < g.nzp) {
            const float4 al = ldg4(alpha + (size_t)(shot / g.ns) * g.level + roff[r + 2] + x);
            const size_t off = (size_t)shot * g.level + roff[r + 2] + x;
            *reinterpret_cast<float4 *>(Ga + off) = make_float4(ga[r][0] / al.x, ga[r][1] / al.y, ga[r][2] / al.z, ga[r][3] / al.w);
            *reinterpret_cast<float4 *>(Gk + off) = make_float4(gk[r][0] / al.x, gk[r][1] / al.y, gk[r][2] / al.z, gk[r][3] / al.w);
        }
    }
}

}  // namespace

cudaError_t launch_imaging(const Plan &p, const float *phist, const float *uhist, const float *alpha, float *Ga, float *Gk,
                           int shot0, int nshots, cudaStream_t st)
{
    const Grid &g = p.g;
    const dim3 grid((g.q4 + kImgCols - 1) / kImgCols, (g.nzp + kImgTileRows - 1) / kImgTileRows, nshots);
    k_imaging<<<grid, kImgCols * kImgRowGroups, 0, st>>>(phist, uhist, alpha, Ga, Gk, g, p.nt, shot0);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rdfwi
