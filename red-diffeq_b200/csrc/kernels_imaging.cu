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

constexpr int kImgCols = 8;        // float4 columns per CTA  (32 cells: one 128-byte line per row)
constexpr int kImgRowGroups = 32;

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float lane(const float4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

// ncu on the first version (profiles/ncu_imaging_r1_full_b64.txt): L1 at 84 % (five row loads + four scalar neighbour
// loads per output row) and half of the instructions address arithmetic.  Hence: two rows per thread share a six-row
// register window, x-neighbours come from the adjacent lanes by shuffle (only the tile's first / last column loads
// them), and the level pointers advance by a constant stride.
template <int kImgRows, int kMinBlocks>
__global__ void __launch_bounds__(kImgCols *kImgRowGroups, kMinBlocks) k_imaging(const float *__restrict__ phist, const float *__restrict__ uhist,
                                                                       const float *__restrict__ alpha, float *__restrict__ Ga,
                                                                       float *__restrict__ Gk, Grid g, int nt, int shot0,
                                                                       int kImgPrefetch /* levels ahead pulled into L2 */)
{
    const int lcol = threadIdx.x % kImgCols;
    const int col = blockIdx.x * kImgCols + lcol;
    constexpr int kImgTileRows = kImgRowGroups * kImgRows;
    const int z0 = blockIdx.y * kImgTileRows + (threadIdx.x / kImgCols) * kImgRows;
    const int shot_l = blockIdx.z, shot = shot0 + shot_l;
    // out-of-range threads keep running (the shuffles below need every lane) on clamped coordinates; they store nothing
    const bool live = col < g.q4 && z0 < g.nzp;
    const int colc = col < g.q4 ? col : g.q4 - 1;
    const int zc = z0 < g.nzp ? z0 : 0;
    const int x = colc * 4;
    const bool edgeL = lcol == 0, edgeR = lcol == kImgCols - 1 || col >= g.q4 - 1;
    const int eL = x == 0 ? g.nxp - 2 : x - 2;
    const int eR = colc == g.q4 - 1 ? g.pitch - g.nxp : x + 4;
    int roff[kImgRows + 4];
#pragma unroll
    for (int k = 0; k < kImgRows + 4; ++k) {
        int z = zc - 2 + k;
        z = z < 0 ? z + g.nzp : (z >= g.nzp ? z - g.nzp : z);
        roff[k] = z * g.pitch;
    }
    const size_t hshot = (size_t)(nt - 1) * g.level;
    const long lvl = (long)g.level;
    const float *pl = phist + (size_t)shot * hshot + (size_t)(nt - 2) * g.level + x;  // p_{t-1} for t = nt-1, then -= level
    const float *ul = uhist + (size_t)shot_l * hshot + x;                              // u_t = slot nt-1-t, then += level
    const float c2 = 4.0f / 3.0f, c3 = -1.0f / 12.0f;

    float ga[kImgRows][4], gk[kImgRows][4];
    float4 unext[kImgRows];  // u_{t+1} of the owned cells
#pragma unroll
    for (int r = 0; r < kImgRows; ++r) {
        unext[r] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 4; ++j) { ga[r][j] = 0.f; gk[r][j] = 0.f; }
    }
    // one lane per 128-byte row segment pulls the lines of a later level into L2, so that the demand loads see L2
    // latency instead of HBM latency (bytes in flight without spending registers)
    const bool prefetcher = lcol == 0 && live;

#pragma unroll 2
    for (int t = nt - 1; t >= 1; --t, pl -= lvl, ul += lvl) {
        if (prefetcher && t - kImgPrefetch >= 1) {
#pragma unroll
            for (int r = 0; r < kImgRows; ++r) {
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pl - kImgPrefetch * lvl + roff[r + 2]));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(ul + kImgPrefetch * lvl + roff[r + 2]));
            }
        }
        float4 rows[kImgRows + 4];
#pragma unroll
        for (int k = 0; k < kImgRows + 4; ++k) rows[k] = ldg4(pl + roff[k]);
#pragma unroll
        for (int r = 0; r < kImgRows; ++r) {
            const float4 ut = ldg4(ul + roff[r + 2]);
            const float4 c = rows[r + 2];
            float l2 = __shfl_up_sync(0xffffffffu, c.z, 1), l1 = __shfl_up_sync(0xffffffffu, c.w, 1);
            float r0 = __shfl_down_sync(0xffffffffu, c.x, 1), r1 = __shfl_down_sync(0xffffffffu, c.y, 1);
            if (edgeL) { l2 = __ldg(pl - x + roff[r + 2] + eL); l1 = __ldg(pl - x + roff[r + 2] + eL + 1); }
            if (edgeR) { r0 = __ldg(pl - x + roff[r + 2] + eR); r1 = __ldg(pl - x + roff[r + 2] + eR + 1); }
            const float e[8] = {l2, l1, c.x, c.y, c.z, c.w, r0, r1};
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
    if (!live) return;
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
    const int R = p.img_rows == 2 ? 2 : 1;
    const int pf = p.img_prefetch > 0 ? p.img_prefetch : 2;  // measured: 2 -> 56.6 ms, 4 -> 58.7, 8 -> 67.3 per step
    const int tile_rows = kImgRowGroups * R;
    const dim3 grid((g.q4 + kImgCols - 1) / kImgCols, (g.nzp + tile_rows - 1) / tile_rows, nshots);
    if (R == 2) k_imaging<2, 2><<<grid, kImgCols * kImgRowGroups, 0, st>>>(phist, uhist, alpha, Ga, Gk, g, p.nt, shot0, pf);
    else if (p.img_rows == 1) k_imaging<1, 4><<<grid, kImgCols * kImgRowGroups, 0, st>>>(phist, uhist, alpha, Ga, Gk, g, p.nt, shot0, pf);
    else k_imaging<1, 3><<<grid, kImgCols * kImgRowGroups, 0, st>>>(phist, uhist, alpha, Ga, Gk, g, p.nt, shot0, pf);  // measured best
    count_launch();
    return cudaGetLastError();
}

}  // namespace rdfwi
