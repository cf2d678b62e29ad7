// kernels_step.cu -- one time level of the FUSED per-level adjoint (q-variable, imaging sums accumulated in place), all
// shots of a chunk.  Used with histories checkpointed in time and with option adj_mode = 1; the default per-level
// adjoint is the split one (kernels_tile.cu in adjoint-field mode + kernels_imaging.cu).
//
// Adjoint (replaces what autograd replays from the tape, core/inversion.py:86):
//     q_t  = T1 q_{t+1} + S(alpha q_{t+1}) - T2 q_{t+2}  (+ receiver cotangent of level t)
//     Ga  += q_t (S-5) p_{t-1}      Gk += (q_{t+1} - q_t) p_{t-1}      Gb[s] += q_t[src_s] w_t
//   summed over the shots of a model in registers, one read-modify-write of Ga/Gk per cell per level.
//
// Work decomposition: a thread owns a float4 (4 consecutive cells of a row) for R consecutive rows and
// loops over the ns shots of its model, so alpha / kappa / T1 / T2 are fetched once and shared by the
// shots.  Rows z-2..z+R+1 are loaded once as float4 and reused for the R rows (register z-marching);
// x-neighbours outside the float4 are 4 scalar loads that hit L1.  grid = (float4 slots, models).
#include <algorithm>

#include "rdfwi_common.cuh"

namespace rdfwi {
namespace {

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

__device__ __forceinline__ float lane(const float4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

// profile index of a padded column / row: distance into the sponge, or -1 in the interior band
__device__ __forceinline__ int sponge_index(int i, int n, int nbc)
{
    return i < nbc ? nbc - 1 - i : (i >= n - nbc ? i - (n - nbc) : -1);
}

struct Cols {
    int xm2, xm1, xp4, xp5;  // wrapped scalar-neighbour columns
    int xc[4];               // true (wrapped) column of each lane
};

__device__ __forceinline__ Cols make_cols(int x, int nxp)
{
    Cols c;
    c.xm2 = x - 2 < 0 ? x - 2 + nxp : x - 2;
    c.xm1 = x - 1 < 0 ? x - 1 + nxp : x - 1;
    c.xp4 = x + 4 >= nxp ? x + 4 - nxp : x + 4;
    c.xp5 = x + 5 >= nxp ? x + 5 - nxp : x + 5;
#pragma unroll
    for (int j = 0; j < 4; ++j) c.xc[j] = x + j >= nxp ? x + j - nxp : x + j;
    return c;
}

// kappa*dt of the 4 lanes of row z (columns override rows in the corners, solvers/pde.py:48-51)
__device__ __forceinline__ void load_kappa(const float *__restrict__ kap_b, const Cols &c, int z, const Grid &g, float kp[4])
{
    const int kz = sponge_index(z, g.nzp, g.nbc);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int kx = sponge_index(c.xc[j], g.nxp, g.nbc);
        const int k = kx >= 0 ? kx : (kz >= 0 ? kz : g.nbc);
        kp[j] = kap_b[k];
    }
}

// ------------------------------------------------------------------------------------------------ adjoint
template <int R>
__global__ void __launch_bounds__(kThreads) k_adj_step(AdjArgs a, Grid g)
{
    const int i = blockIdx.x * kThreads + threadIdx.x;
    const int ngroups = (g.nzp + R - 1) / R;
    if (i >= ngroups * g.q4) return;
    const int zg = i / g.q4;
    const int x = (i - zg * g.q4) * 4;
    const int z0 = zg * R;
    const int b = blockIdx.y;
    const Cols c = make_cols(x, g.nxp);

    int roff[R + 4];
#pragma unroll
    for (int k = 0; k < R + 4; ++k) {
        int z = z0 - 2 + k;
        z = z < 0 ? z + g.nzp : (z >= g.nzp ? z - g.nzp : z);
        roff[k] = z * g.pitch;
    }

    // alpha on the whole stencil footprint (S acts on alpha*q), T1/T2 on the owned cells
    const float *alpha_b = a.alpha + (size_t)b * g.level;
    const float *kap_b = a.kap + (size_t)b * (g.nbc + 1);
    float4 arow[R + 4];
#pragma unroll
    for (int k = 0; k < R + 4; ++k) arow[k] = ld4(alpha_b + roff[k] + x);
    float asc[R][4], t1[R][4], t2[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        asc[r][0] = alpha_b[roff[r + 2] + c.xm2];
        asc[r][1] = alpha_b[roff[r + 2] + c.xm1];
        asc[r][2] = alpha_b[roff[r + 2] + c.xp4];
        asc[r][3] = alpha_b[roff[r + 2] + c.xp5];
        float kp[4];
        load_kappa(kap_b, c, z0 + r, g, kp);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            t1[r][j] = (2.0f + (-5.0f * lane(arow[r + 2], j))) - kp[j];
            t2[r][j] = 1.0f - kp[j];
        }
    }

    const float c2 = 4.0f / 3.0f, c3 = -1.0f / 12.0f;
    float ga[R][4], gk[R][4];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < 4; ++j) { ga[r][j] = 0.0f; gk[r][j] = 0.0f; }

    for (int s = blockIdx.z; s < g.ns; s += gridDim.z) {  // one imaging plane per (model, grid.z slice)
        const size_t so = (size_t)(b * g.ns + s) * g.level;
        const float *__restrict__ Q1 = a.q1 + so;
        const float *__restrict__ Q2 = a.q2 + so;
        const float *__restrict__ PM = a.pm1 + (size_t)(b * g.ns + s) * a.ss_pm1;
        float *__restrict__ QO = a.out + so;

        float4 qrow[R + 4], prow[R + 4];
#pragma unroll
        for (int k = 0; k < R + 4; ++k) {
            qrow[k] = ld4(Q1 + roff[k] + x);
            prow[k] = ld4(PM + roff[k] + x);
        }
        float4 q2c[R];
        float qsc[R][4], psc[R][4];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            q2c[r] = ld4(Q2 + roff[r + 2] + x);
            qsc[r][0] = Q1[roff[r + 2] + c.xm2]; psc[r][0] = PM[roff[r + 2] + c.xm2];
            qsc[r][1] = Q1[roff[r + 2] + c.xm1]; psc[r][1] = PM[roff[r + 2] + c.xm1];
            qsc[r][2] = Q1[roff[r + 2] + c.xp4]; psc[r][2] = PM[roff[r + 2] + c.xp4];
            qsc[r][3] = Q1[roff[r + 2] + c.xp5]; psc[r][3] = PM[roff[r + 2] + c.xp5];
        }
        const int xs = a.isx[s];

#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int z = z0 + r;
            if (z < g.nzp) {
                // alpha*q on the row's 8-wide window and on the rows above / below
                float eq[8], ep[8];
                eq[0] = asc[r][0] * qsc[r][0]; eq[1] = asc[r][1] * qsc[r][1];
                eq[6] = asc[r][2] * qsc[r][2]; eq[7] = asc[r][3] * qsc[r][3];
                ep[0] = psc[r][0]; ep[1] = psc[r][1]; ep[6] = psc[r][2]; ep[7] = psc[r][3];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    eq[j + 2] = lane(arow[r + 2], j) * lane(qrow[r + 2], j);
                    ep[j + 2] = lane(prow[r + 2], j);
                }
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float n1 = lane(arow[r + 1], j) * lane(qrow[r + 1], j);
                    const float s1 = lane(arow[r + 3], j) * lane(qrow[r + 3], j);
                    const float n2 = lane(arow[r], j) * lane(qrow[r], j);
                    const float s2 = lane(arow[r + 4], j) * lane(qrow[r + 4], j);
                    const float sa = c2 * (((n1 + s1) + eq[j + 1]) + eq[j + 3]) + c3 * (((n2 + s2) + eq[j]) + eq[j + 4]);
                    const float qc = lane(qrow[r + 2], j);
                    o[j] = (t1[r][j] * qc + sa) - t2[r][j] * lane(q2c[r], j);
                }
                if (a.cot != nullptr && z == g.igz) {  // adjoint of the receiver gather (:83)
                    const float *d = a.cot + ((size_t)(b * g.ns + s) * g.nt_out + a.it_out) * g.nrec;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int xc = c.xc[j];
                        for (int k = a.rec_ptr[xc]; k < a.rec_ptr[xc + 1]; ++k) o[j] += d[a.rec_idx[k]];
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float pc = ep[j + 2];
                    const float lp = -5.0f * pc + (c2 * (((lane(prow[r + 1], j) + lane(prow[r + 3], j)) + ep[j + 1]) + ep[j + 3]) +
                                                   c3 * (((lane(prow[r], j) + lane(prow[r + 4], j)) + ep[j]) + ep[j + 4]));
                    ga[r][j] += o[j] * lp;
                    gk[r][j] += (lane(qrow[r + 2], j) - o[j]) * pc;
                }
                if (z == g.isz) {  // adjoint of the source injection (:81): one owner thread per shot
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (x + j < g.nxp && x + j == xs) a.Gb[b * g.ns + s] += o[j] * a.w_t;
                }
                st4(QO + roff[r + 2] + x, make_float4(o[0], o[1], o[2], o[3]));
            }
        }
    }

#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (z0 + r < g.nzp) {
            float *pa = a.Ga + ((size_t)b * gridDim.z + blockIdx.z) * g.level + roff[r + 2] + x;
            float *pk = a.Gk + ((size_t)b * gridDim.z + blockIdx.z) * g.level + roff[r + 2] + x;
            float4 va = ld4(pa), vk = ld4(pk);
            va.x += ga[r][0]; va.y += ga[r][1]; va.z += ga[r][2]; va.w += ga[r][3];
            vk.x += gk[r][0]; vk.y += gk[r][1]; vk.z += gk[r][2]; vk.w += gk[r][3];
            st4(pa, va);
            st4(pk, vk);
        }
    }
}

}  // namespace

// grid.z slices over which the shots of a model are dealt so that a launch has at least ~2 waves of CTAs;
// the adjoint keeps at least `min_per_slice` shots per slice (each slice owns an imaging plane: its read-modify-write
// per level is amortised over the slice's shots)
int shot_slices(const Plan &p, int blocks_xy, int nb, int min_per_slice)
{
    int sms = 148;
    device_attr(&sms, cudaDevAttrMultiProcessorCount, p.device);
    const int want = (2 * 2 * sms + blocks_xy * nb - 1) / (blocks_xy * nb);  // 2 waves at ~2 CTAs per SM
    int sz = std::min(want, std::max(1, p.g.ns / min_per_slice));
    return std::max(1, std::min(sz, p.g.ns));
}

int adj_shot_slices(const Plan &p, int nb)
{
    const Grid &g = p.g;
    const int R = p.adj_rows_per_thread;
    const int groups = (g.nzp + R - 1) / R;
    return shot_slices(p, (groups * g.q4 + kThreads - 1) / kThreads, nb, 4);
}

cudaError_t launch_adj_step(const Plan &p, const AdjArgs &a, int nb, cudaStream_t st)
{
    const Grid &g = p.g;
    const int R = p.adj_rows_per_thread;
    const int groups = (g.nzp + R - 1) / R;
    const dim3 grid((unsigned)((groups * g.q4 + kThreads - 1) / kThreads), (unsigned)nb, (unsigned)a.slices);
    switch (R) {
        case 1: k_adj_step<1><<<grid, kThreads, 0, st>>>(a, g); break;
        default: k_adj_step<2><<<grid, kThreads, 0, st>>>(a, g); break;
    }
    count_launch();
    return cudaSuccess;
}

}  // namespace rdfwi
