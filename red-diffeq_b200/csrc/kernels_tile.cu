// kernels_tile.cu -- one time level for grids that do not fit a thread-block cluster: HBM-streaming 2-D tiles.
//
// Forward (EXACT): the body of the reference's hot loop (solvers/pde.py:79-83), one rounding per reference tensor op in
// the reference's association order (__fmul_rn / __fadd_rn / __fsub_rn: no FMA contraction) => bit-identical fields.
// Adjoint field (ADJ): the same recurrence in the u-variable, u = alpha*q,
//     u_t = T1 u_{t+1} + alpha S(u_{t+1}) - T2 u_{t+2},   u_t[rec] += alpha * g_t,   Gb += u_t[src] w_t
// (the adjoint of solvers/pde.py:79-83 that autograd replays from the tape, core/inversion.py:86); fused
// multiply-adds allowed.  The levels it writes are the adjoint-field history the pointwise imaging kernel
// (kernels_imaging.cu) reads, so the per-level adjoint moves 12 + 8 bytes per cell-update and has no accumulators.
//
// Work decomposition: a CTA is 32 x 8 threads and owns a tile of 128 columns x 8R rows of ONE shot.  The tile of
// p_{t-1} with its two halo rows / columns on every side, and the tiles of p_{t-2} and alpha, are brought into shared
// memory with cp.async (16-byte copies that bypass L1 and cost no registers: ~52 KB in flight per CTA, three or four
// CTAs per SM -- far more than the ~32 KB per SM that HBM latency x bandwidth asks for); then a thread marches down R
// consecutive rows of its float4 column with the five stencil rows in registers, x-neighbours from shared memory.
// The periodic wrap of torch.roll is resolved when the copies are issued (row / column indices), never in the stencil.
// Every HBM byte is touched once per level: 12 B per cell-update (read p_{t-1}, p_{t-2}, write p_t).
#include "rdfwi_common.cuh"
#include "cluster_ptx.cuh"

namespace rdfwi {
namespace {

constexpr int kTileX = 32;            // float4 per tile row (128 columns) = one warp
constexpr int kTileY = 8;             // row groups per tile
constexpr int kTileW = 4 * kTileX;    // columns per tile
constexpr int kRowW = kTileW + 8;     // shared-memory row of p_{t-1}: [2 unused][2 left halo][128][2 right halo][2 unused]

__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float2 lds2(const float *p) { return *reinterpret_cast<const float2 *>(p); }
__device__ __forceinline__ void st4_stream(float *p, float4 v)
{
    // written once, read again a whole level later: keep it out of L1
    asm volatile("st.global.cg.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void cp_async16(float *sdst, const float *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async4(float *sdst, const float *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
constexpr size_t tile_smem_bytes(int R) { return ((size_t)(kTileY * R + 4) * kRowW + 2 * (size_t)kTileY * R * kTileW) * sizeof(float); }

// ZINT: no row of the tile or of its halo wraps around or leaves the grid (all but the first and last tile rows of a
// grid): row offsets are affine and nothing is clamped.  Instruction issue, not HBM, bounded the first version of this
// kernel (68 instructions per cell-update, 28 % of them floating point: profiles/ncu_tile_r1_v2_issue_bound.txt).
template <int R, bool ADJ, bool ZINT>
__device__ __forceinline__ void tile_body(const StepArgs &a, const Grid &g, float *tsm)
{
    constexpr int TR = kTileY * R;  // rows of the tile
    float *sP1 = tsm;                        // (TR+4) x kRowW
    float *sP0 = sP1 + (TR + 4) * kRowW;     // TR x 128
    float *sAl = sP0 + TR * kTileW;          // TR x 128

    const int lx = threadIdx.x, ly = threadIdx.y;
    const int x4 = blockIdx.x * kTileX + lx;
    const bool col_ok = x4 < g.q4;
    const int x = x4 * 4;
    const int tx0 = blockIdx.x * kTileW, tz0 = blockIdx.y * TR;
    const int shot_l = blockIdx.z, gshot = a.shot0 + shot_l;
    const int b = gshot / g.ns, s = gshot - b * g.ns;
    const int pitch = g.pitch;

    const float *__restrict__ P1 = a.p1 + (size_t)shot_l * a.ss_p1;
    const float *__restrict__ P0 = a.p0 + (size_t)shot_l * a.ss_p0;
    float *__restrict__ PO = a.out + (size_t)shot_l * a.ss_out;
    const float *__restrict__ alpha_b = a.alpha + (size_t)b * g.level;
    const float *__restrict__ kap_b = a.kap + (size_t)b * (g.nbc + 1);

    // rows are periodic in z (torch.roll); rows past the grid are clamped (their results are never stored)
    auto row_off = [&](int z) {
        if (!ZINT) {
            z = z < 0 ? z + g.nzp : (z >= g.nzp ? z - g.nzp : z);
            z = z >= g.nzp ? g.nzp - 1 : z;
        }
        return z * pitch;
    };
    // ---- issue every load of the tile ----------------------------------------------------------------------------
    if (col_ok) {
#pragma unroll
        for (int k = 0; k < (TR + 4 + kTileY - 1) / kTileY; ++k) {
            const int rr = ly + k * kTileY;
            if (rr < TR + 4) cp_async16(sP1 + rr * kRowW + 4 + 4 * lx, P1 + row_off(tz0 - 2 + rr) + x);
        }
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int rr = ly + k * kTileY;
            const int ro = row_off(tz0 + rr);
            cp_async16(sP0 + rr * kTileW + 4 * lx, P0 + ro + x);
            cp_async16(sAl + rr * kTileW + 4 * lx, alpha_b + ro + x);
        }
    }
    if (lx < 4) {
        // x-halo of the centre rows: the two columns left of the tile and right of its last float4 (periodic in x; the
        // image columns nxp..pitch-1 of the last float4 are followed by column pitch-nxp)
        const int nvalid = g.q4 - blockIdx.x * kTileX < kTileX ? g.q4 - blockIdx.x * kTileX : kTileX;
        const bool last_tile = blockIdx.x * kTileX + nvalid == g.q4;
        const int gcol = lx < 2 ? (blockIdx.x == 0 ? g.nxp - 2 : tx0 - 2) + lx : (last_tile ? pitch - g.nxp : tx0 + kTileW) + (lx - 2);
        const int scol = lx < 2 ? 2 + lx : 4 + 4 * nvalid + (lx - 2);
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const int rr = ly + k * kTileY;
            cp_async4(sP1 + (rr + 2) * kRowW + scol, P1 + row_off(tz0 + rr) + gcol);
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    // per-thread constants while the copies fly
    const int xq = col_ok ? x : 0;
    int xc[4];
    float kx[4];  // kappa*dt contributed by the column (columns override rows in the corners, solvers/pde.py:48-51)
    bool colsp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        xc[j] = xq + j >= g.nxp ? xq + j - g.nxp : xq + j;
        const int k = sponge_index(xc[j], g.nxp, g.nbc);
        colsp[j] = k >= 0;
        kx[j] = colsp[j] ? kap_b[k] : 0.0f;
    }
    const float c2 = 4.0f / 3.0f;    // fp32(4.0/3.0), the reference's python scalar cast by the tensor op
    const float c3 = -1.0f / 12.0f;
    const int lr0 = ly * R;          // first tile row of this thread
    const int zt = tz0 + lr0;        // its grid row
    float kz[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int z = zt + r;
        const int kzi = sponge_index((ZINT || z < g.nzp) ? z : g.nzp - 1, g.nzp, g.nbc);
        kz[r] = kzi >= 0 ? kap_b[kzi] : 0.0f;
    }
    // rows of this thread that carry a source or receivers (usually none)
    const bool special = (g.isz >= zt && g.isz < zt + R) || (g.igz >= zt && g.igz < zt + R);

    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    if (!col_ok) return;

    // ---- march down the R rows ---------------------------------------------------------------------------------
    const float *c1p = sP1 + lr0 * kRowW + 4 + 4 * lx;  // row z-2 of the thread's first row
    const float *p0p = sP0 + lr0 * kTileW + 4 * lx;
    const float *alp_p = sAl + lr0 * kTileW + 4 * lx;
    float *outp = PO + (size_t)zt * pitch + x;
    const bool swap = (lx & 8) != 0;
    float4 w0 = lds4(c1p), w1 = lds4(c1p + kRowW), w2 = lds4(c1p + 2 * kRowW), w3 = lds4(c1p + 3 * kRowW);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float4 w4 = lds4(c1p + (r + 4) * kRowW);
        // neighbour pairs without bank conflicts: lanes 0-7 / 16-23 read their left pair first, the others their right pair
        // (see fwd_sweep in kernels_cluster.cu: a warp's left pairs only touch banks 2, 3 mod 4, its right pairs 0, 1 mod 4)
        const float2 qa = lds2(c1p + (r + 2) * kRowW + (swap ? 4 : -2)), qb = lds2(c1p + (r + 2) * kRowW + (swap ? -2 : 4));
        const float2 lft = swap ? qb : qa, rgt = swap ? qa : qb;
        const float4 old = lds4(p0p + r * kTileW);
        const float4 al = lds4(alp_p + r * kTileW);
        const float e[8] = {lft.x, lft.y, w2.x, w2.y, w2.z, w2.w, rgt.x, rgt.y};
        float o[4];
        // two cells per instruction (FADD2 / FFMA2).  Forward: only the twelve additions of a cell are packed, the six
        // products stay scalar __fmul_rn so that ptxas cannot contract them into FFMA2 -- one rounding per reference op,
        // same association as solvers/pde.py:79 (+ is commutative) => bit-identical fields.
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 2 * h;
            const float2 up1 = h ? make_float2(w1.z, w1.w) : make_float2(w1.x, w1.y);
            const float2 dn1 = h ? make_float2(w3.z, w3.w) : make_float2(w3.x, w3.y);
            const float2 up2 = h ? make_float2(w0.z, w0.w) : make_float2(w0.x, w0.y);
            const float2 dn2 = h ? make_float2(w4.z, w4.w) : make_float2(w4.x, w4.y);
            const float2 oldp = h ? make_float2(old.z, old.w) : make_float2(old.x, old.y);
            const float2 alp = h ? make_float2(al.z, al.w) : make_float2(al.x, al.y);
            // (((p1[z-1] + p1[z+1]) + p1[x-1]) + p1[x+1]) and the same at distance 2
            const float2 s1 = f2add(f2add(f2add(up1, dn1), make_float2(e[j + 1], e[j + 2])), make_float2(e[j + 3], e[j + 4]));
            const float2 s2 = f2add(f2add(f2add(up2, dn2), make_float2(e[j], e[j + 1])), make_float2(e[j + 4], e[j + 5]));
            const float2 kp = make_float2(colsp[j] ? kx[j] : kz[r], colsp[j + 1] ? kx[j + 1] : kz[r]);
            float2 res;
            if (!ADJ) {
                const float2 lap = f2add(make_float2(__fmul_rn(c2, s1.x), __fmul_rn(c2, s1.y)),
                                         make_float2(__fmul_rn(c3, s2.x), __fmul_rn(c3, s2.y)));
                const float2 t1 = f2sub(f2add(make_float2(2.0f, 2.0f), make_float2(__fmul_rn(-5.0f, alp.x), __fmul_rn(-5.0f, alp.y))), kp);  // temp1 (:69)
                const float2 t2 = f2sub(make_float2(1.0f, 1.0f), kp);                                                                       // temp2 (:70)
                const float2 a1 = make_float2(__fmul_rn(t1.x, e[j + 2]), __fmul_rn(t1.y, e[j + 3]));
                const float2 a2 = make_float2(__fmul_rn(t2.x, oldp.x), __fmul_rn(t2.y, oldp.y));
                const float2 a3 = make_float2(__fmul_rn(alp.x, lap.x), __fmul_rn(alp.y, lap.y));
                res = f2add(f2sub(a1, a2), a3);
            } else {
                const float2 cen = make_float2(e[j + 2], e[j + 3]);
                const float2 lap = f2fma(make_float2(c2, c2), s1, f2mul(make_float2(c3, c3), s2));
                const float2 t1 = f2sub(f2fma(make_float2(-5.0f, -5.0f), alp, make_float2(2.0f, 2.0f)), kp);
                const float2 t2 = f2sub(make_float2(1.0f, 1.0f), kp);
                res = f2fma(alp, lap, f2sub(f2mul(t1, cen), f2mul(t2, oldp)));
            }
            o[j] = res.x; o[j + 1] = res.y;
        }
        const int z = zt + r;
        if (special) {
            const int xs = a.isx[s];
            if (!ADJ) {
                if (z == g.isz) {  // p[src] += beta_dt[src] * wavelet[t]   (:80-81); periodic images of the column included
                    const float src_add = __fmul_rn(a.beta_src[gshot], a.w_t);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (xc[j] == xs) o[j] = __fadd_rn(o[j], src_add);
                }
                if (a.seis != nullptr && z == g.igz) {  // sampled after the injection (:82-83)
                    float *d = a.seis + ((size_t)gshot * g.nt_out + a.it_out) * g.nrec;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (x + j < g.nxp)
                            for (int k = a.rec_ptr[x + j]; k < a.rec_ptr[x + j + 1]; ++k) d[a.rec_idx[k]] = o[j];
                }
            } else {
                const float av[4] = {al.x, al.y, al.z, al.w};
                if (a.cot != nullptr && z == g.igz) {  // adjoint of the receiver gather (:83), in the u-variable
                    const float *d = a.cot + ((size_t)gshot * g.nt_out + a.it_out) * g.nrec;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float acc = 0.0f;
                        for (int k = a.rec_ptr[xc[j]]; k < a.rec_ptr[xc[j] + 1]; ++k) acc += d[a.rec_idx[k]];
                        o[j] += av[j] * acc;
                    }
                }
                if (z == g.isz) {  // adjoint of the source injection (:81): one owner thread per shot
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (x + j < g.nxp && x + j == xs) {
                            const float acc = a.Gb[gshot] + o[j] * a.w_t;
                            a.Gb[gshot] = a.last ? acc / av[j] : acc;  // the last level leaves sum_t q_t[src] w_t
                        }
                }
            }
        }
        if (ZINT || z < g.nzp) st4_stream(outp + (size_t)r * pitch, make_float4(o[0], o[1], o[2], o[3]));
        w0 = w1; w1 = w2; w2 = w3; w3 = w4;
    }
}

template <int R, bool ADJ>
__global__ void __launch_bounds__(kTileX *kTileY, R <= 4 ? 3 : 2) k_step_tile(StepArgs a, Grid g)
{
    extern __shared__ __align__(16) float tsm[];
    const int tz0 = blockIdx.y * (kTileY * R);
    if (tz0 >= 2 && tz0 + kTileY * R + 2 <= g.nzp) tile_body<R, ADJ, true>(a, g, tsm);
    else tile_body<R, ADJ, false>(a, g, tsm);
}

template <bool ADJ>
cudaError_t launch_t(const Plan &p, const StepArgs &a, cudaStream_t st)
{
    const Grid &g = p.g;
    const int R = p.rows_per_thread;
    const dim3 block(kTileX, kTileY);
    const dim3 grid((unsigned)((g.q4 + kTileX - 1) / kTileX), (unsigned)((g.nzp + kTileY * R - 1) / (kTileY * R)), (unsigned)a.nshots);
    static bool attr_set[64] = {false};  // > 48 KB of dynamic shared memory needs the opt-in, once per device
    if (!attr_set[p.device & 63]) {
        cudaError_t e = cudaFuncSetAttribute(k_step_tile<1, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(1));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_tile<2, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(2));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_tile<4, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(4));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_step_tile<8, ADJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem_bytes(8));
        if (e != cudaSuccess) return e;
        attr_set[p.device & 63] = true;
    }
    switch (R) {
        case 1: k_step_tile<1, ADJ><<<grid, block, tile_smem_bytes(1), st>>>(a, g); break;
        case 2: k_step_tile<2, ADJ><<<grid, block, tile_smem_bytes(2), st>>>(a, g); break;
        case 8: k_step_tile<8, ADJ><<<grid, block, tile_smem_bytes(8), st>>>(a, g); break;
        default: k_step_tile<4, ADJ><<<grid, block, tile_smem_bytes(4), st>>>(a, g); break;
    }
    count_launch();
    return cudaSuccess;
}

}  // namespace

cudaError_t launch_step_tile(const Plan &p, const StepArgs &a, cudaStream_t st)
{
    return a.adj ? launch_t<true>(p, a, st) : launch_t<false>(p, a, st);
}

}  // namespace rdfwi
