// kernels_cluster.cu -- cluster-resident time loop: a whole shot lives in the shared memory of one
// thread-block cluster for all nt levels.
//
// Replaces the same reference code as kernels_step.cu (the hot loop solvers/pde.py:78-85 and, for the
// adjoint, the autograd replay behind core/inversion.py:86) with identical arithmetic, but removes the
// per-level HBM round trip of the wavefields: a padded OpenFWI shot is 2 x 387 KB (p_{t-1}, p_{t-2}),
// which fits the shared memory of a 4-CTA cluster.  One persistent launch runs all levels of all shots
// (clusters take shots round-robin); per level each CTA
//   1. updates its slab of rows in place (p_t overwrites p_{t-2}: a cell needs p_{t-2} only at itself),
//      z-marching in registers, x-neighbour pairs from the resident row (conflict-free LDS.64, fwd_sweep);
//   2. pushes its two first / last rows into the neighbours' halo rows through distributed shared memory
//      (st.async ... mbarrier::complete_tx), periodic in z like torch.roll -- no cluster barrier in the loop;
//   3. forward (MODE 0): an elected thread streams the slab to the wavefield history with a 1-D bulk copy
//      (cp.async.bulk shared -> global), overlapped with the next level; the time-invariant coefficient
//      rows (alpha, t1) come from TENSOR MEMORY, used as a per-thread scratchpad (cluster_ptx.cuh);
//      adjoint (MODE 2, the default): the forward history streams IN (cp.async into thread-private
//      slots) and the zero-lag imaging sums are formed in the same sweep, their accumulators in tensor
//      memory -- the adjoint field never leaves the chip.  (MODE 1 + kernels_imaging.cu: the split adjoint
//      of round 1, kept as an option and as the cross-check.)
// HBM traffic: the history, written once by the forward and read once by the adjoint: 8 B per
// forward+adjoint cell-update pair instead of the 28 B a streaming implementation moves.
#include <mutex>

#include "cluster_ptx.cuh"

namespace rdfwi {
namespace {

// ------------------------------------------------------------------------------------------------ forward
// Hot loop notes (from the ncu captures under profiles/): in its first versions the kernel lost half of its
// cycles at a per-level barrier.cluster (MEMBAR.ALL.GPU + skew) and was issue-bound.  Hence:
//   * PITCH is a template parameter for the production grids (0 = runtime pitch), so every row offset is
//     an immediate of the LDS/STS instruction;
//   * rows a thread does not own are still computed (from in-bounds garbage), only their store is predicated;
//   * source injection and receiver sampling run in a short epilogue, executed by the owner threads;
//   * there is NO cluster-wide barrier inside the time loop.  Halo rows travel as st.async stores that
//     complete transaction bytes on an mbarrier of the receiving CTA.  The row group that owns the top
//     rows marches downwards and the one that owns the bottom rows marches upwards, so both produce the
//     rows their neighbours need in their first two iterations and send them at once; the consumers wait
//     for them only at the start of the next level, a whole sweep later.  The write-after-read hazard on
//     a halo buffer is covered by the data dependency itself: a CTA sends level-t edge rows only after it
//     received (and read) its neighbour's level t-1 rows, which that neighbour sent after reading its own
//     halo.  Inside a CTA one __syncthreads per level orders the row groups.
//   * two sweep instantiations (down / up) still fit the instruction cache; four (x two buffer roles) did
//     not (ncu: stall_no_instruction), so buffer roles are runtime base pointers.

// IMG (resident imaging, MODE 2 of k_fwd_cluster): the sweep that produces u_t also adds this level's terms of the zero-lag
// imaging sums (kernels_imaging.cu has the formulas) for the thread's cells,
//     ga += p_t [ u_t - (2-kappa) u_{t+1} + (1-kappa) u_{t+2} ],      gk += p_t [ u_{t+2} - u_{t+1} ],
// with u_t = the row just computed, u_{t+1} its centre operand, u_{t+2} the value it overwrites.  The 8 accumulators per
// float4 live in TENSOR MEMORY -- 104 of the thread's 128 private columns -- because the register file is full (alpha alone
// is 52 registers); ATM > 0: alpha of the last ATM marching rows lives there as well and is read with the accumulators.
// p_t comes from the forward history in HBM.  Loaded into registers (LDG two rows ahead) ptxas sinks every load to ~10
// instructions before its first use whatever the source order, and the sweep waits a full L2 / HBM latency per row
// (65 ms per 64 models, 40 ms with the loads removed: profiles/resident_adjoint_r2.md).  So each thread owns two 16-byte
// slots in shared memory and fills them with cp.async two rows ahead: asynchronous, no registers, no scheduler to argue
// with; the slab of the level three sweeps ahead is pulled into L2 by one bulk prefetch per level.
#ifndef RDFWI_IMG_PF
#define RDFWI_IMG_PF 1
#endif
constexpr int kImgPrefetchLevels = RDFWI_IMG_PF;
// 16-byte shared-memory slots per thread staging the forward rows of the resident imaging (see fwd_sweep)
__host__ __device__ constexpr int img_ring_slots(const int rows) { return rows <= 7 ? rows : 2; }
// CTM (forward mode): the two time-invariant coefficient rows of a cell -- alpha and t1 = (2 - 5 alpha) - kappa, both
// rounded exactly as the reference rounds them -- live in tensor memory (104 columns per thread) instead of 52 registers
// for alpha plus four instructions per cell pair and level recomputing t1.
template <int RMAX, int PITCH, int DIR, bool EXACT, bool IMG, int ATM, bool CTM>
__device__ __forceinline__ void fwd_sweep(float *__restrict__ smem, const int cur, const int prv, const int kap_off,
                                          const int pitch_rt, const int l0, const SweepThread &th,
                                          const float4 (&al)[RMAX], const float (&kapx)[4], const HaloPush &hp,
                                          const uint64_t *push_bar, const bool push_now, const ImgThread &im)
{
    const int pitch = PITCH > 0 ? PITCH : pitch_rt;
    const int P = DIR * pitch;                                     // signed row step in marching order
    const float c2 = 4.0f / 3.0f, c3 = -1.0f / 12.0f;
    const float *cb = smem + cur + (2 + l0) * pitch + th.x;        // row l0 of p_{t-1}, this thread's float4
    float *pb = smem + prv + (2 + l0) * pitch + th.x;              // row l0 of p_{t-2}; p_t goes there
    const float *eLp = smem + cur + (2 + l0) * pitch + th.eL;
    const float *eRp = smem + cur + (2 + l0) * pitch + th.eR;
#ifndef RDFWI_PLAIN_PAIRS
    // Bank-conflict-free neighbour pairs.  A lane reads 8 bytes 16 bytes apart from its neighbours': the left pairs of a
    // warp only touch banks = 2, 3 (mod 4), the right pairs banks = 0, 1 (mod 4), so each LDS.64 is a 2-way conflict (4
    // wavefronts instead of 2; 8 of a row's 27).  Let lanes 0-7 / 16-23 read their LEFT pair and lanes 8-15 / 24-31 their
    // RIGHT pair in one instruction and the other way round in a second one: every 16-lane wavefront then covers all 32
    // banks once.  Two selects per pair put the values back in place (the kernel is not issue-bound).
    const float *pairA = th.swap ? eRp : eLp, *pairB = th.swap ? eLp : eRp;
#endif
    const float *kz = smem + kap_off + l0;
    const float *push_dst = smem + prv + hp.dst + th.x;
    const int nvalid = th.lb - th.la;
    // Forward-history rows of the thread.  Wide clusters with few rows per thread (launches of few shots) have room for one
    // slot per row: every row of level t-1 is requested while level t is swept, a whole level ahead (a 4-row sweep is shorter
    // than two L2 round trips).  The throughput configuration (13 rows, shared memory full) has two slots, two rows ahead.
    constexpr int RING = img_ring_slots(RMAX);
    const uint32_t ring0 = IMG ? smem_u32(im.ring) : 0, ring1 = ring0 + kClusterThreads * 16;
    if (IMG && im.first) {
        // rows the thread does not own are fetched all the same (immediate offsets: no address arithmetic per row); they lie
        // inside the history or the kClusterRowsMax rows of padding behind it, their sums are never written out.
        // (Only the first level of a shot starts its copies here: every other level's first rows were requested by the
        // sweep before it, as soon as their slots were free -- see the row loop.)
        if (RING == 2) {
            cp_async16_commit(ring1, im.pg + P);
            cp_async16_commit(ring0, im.pg);
        } else {
#pragma unroll
            for (int q = 0; q < RMAX; ++q) cp_async16_commit(ring0 + q * (kClusterThreads * 16), im.pg + q * P);
        }
    }

    float4 w0 = ld4(cb - 2 * P), w1 = ld4(cb - P), w2 = ld4(cb), w3 = ld4(cb + P);
#pragma unroll
    for (int r = 0; r < RMAX; ++r) {
        float ga[4], gk[4], alt[4];
        float4 pv;
        if (IMG) {
            tm_ld4(im.tm + 4 * r, ga);
            tm_ld4(im.tm + 4 * RMAX + 4 * r, gk);
            if (r >= RMAX - ATM) tm_ld4(im.tm + 8 * RMAX + 4 * (r - (RMAX - ATM)), alt);
            if (RING == 2) {
                const uint32_t slot = (r & 1) ? ring1 : ring0;
                // this row's copy has landed (the next one's may be in flight).  Commit order with 13 rows: ... row 11, row 12,
                // then the next level's row 1 (its slot is free after row 11) and row 0 (after row 12) -- hence row 0 waits for
                // all.  (The last row of a shot's last level has nothing committed behind it either.)
                if (r == 0 || (r == RMAX - 1 && !im.next)) cp_async_wait<0>();
                else cp_async_wait<1>();
                pv = lds4_volatile(slot);
                if (r + 2 < RMAX) cp_async16_commit(slot, im.pg + (r + 2) * P);
                else if (im.next) cp_async16_commit(slot, im.pg - im.level + (r & 1) * P);  // forward level t-1: the row of this slot's parity
            } else {
                const uint32_t slot = ring0 + r * (kClusterThreads * 16);
                if (r == 0) cp_async_wait<0>();  // all rows of this level were requested during the previous sweep
                pv = lds4_volatile(slot);
                if (im.next) cp_async16_commit(slot, im.pg - im.level + r * P);
            }
        }
        if (CTM) {
            tm_ld4(im.tm + 4 * r, ga);              // alpha row
            tm_ld4(im.tm + 4 * RMAX + 4 * r, gk);   // t1 row
        }
        const float4 w4 = ld4(cb + (r + 2) * P);
        const float4 old = ld4(pb + r * P);
        const float kapz = kz[DIR * r];
        float l2, l1, r0, r1;
        if (PITCH > 0 && RMAX > 4) {
            // (the 4-row few-shot configuration measured 10 % slower this way: it keeps the shuffles)
            // production grids (even nxp, see the dispatcher): the two x-neighbour pairs are 8-byte aligned in the row that
            // sits in shared memory anyway -- two LDS.64 per row for every lane, instead of four shuffles plus four
            // predicated edge loads behind a divergent branch (ncu: BSSY/BSYNC stalls, ~18 of ~100 instructions per row)
#ifdef RDFWI_PLAIN_PAIRS
            const float2 lp = *reinterpret_cast<const float2 *>(eLp + r * P);
            const float2 rp = *reinterpret_cast<const float2 *>(eRp + r * P);
#else
            const float2 qa = *reinterpret_cast<const float2 *>(pairA + r * P);
            const float2 qb = *reinterpret_cast<const float2 *>(pairB + r * P);
            const float2 lp = th.swap ? qb : qa, rp = th.swap ? qa : qb;
#endif
            l2 = lp.x; l1 = lp.y; r0 = rp.x; r1 = rp.y;
        } else {
            // generic pitch: x-neighbours outside the float4 come from the adjacent lanes' centre vectors
            l2 = __shfl_up_sync(0xffffffffu, w2.z, 1);
            l1 = __shfl_up_sync(0xffffffffu, w2.w, 1);
            r0 = __shfl_down_sync(0xffffffffu, w2.x, 1);
            r1 = __shfl_down_sync(0xffffffffu, w2.y, 1);
            if (th.edgeL) { l2 = eLp[r * P]; l1 = eLp[r * P + 1]; }
            if (th.edgeR) { r0 = eRp[r * P]; r1 = eRp[r * P + 1]; }
        }
        const float e[8] = {l2, l1, w2.x, w2.y, w2.z, w2.w, r0, r1};
        float o[4];
        if (IMG) {
            if (r >= RMAX - ATM) tm_wait_ld(ga, gk, alt);
            else tm_wait_ld(ga, gk);
        }
        if (CTM) tm_wait_ld(ga, gk);
        const float4 alr = CTM ? make_float4(ga[0], ga[1], ga[2], ga[3])
                               : ((IMG && r >= RMAX - ATM) ? make_float4(alt[0], alt[1], alt[2], alt[3]) : al[r]);
        // Two cells per instruction for the twelve additions of a cell (FADD2, IEEE round-to-nearest per lane, same
        // association as the reference); the six multiplications stay scalar so that ptxas cannot contract them
        // into FFMA2 (it does contract packed products, even with .rn) -- seismograms stay bit-identical.
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = 2 * h;
            const float2 up1 = h ? make_float2(w1.z, w1.w) : make_float2(w1.x, w1.y);
            const float2 dn1 = h ? make_float2(w3.z, w3.w) : make_float2(w3.x, w3.y);
            const float2 up2 = h ? make_float2(w0.z, w0.w) : make_float2(w0.x, w0.y);
            const float2 dn2 = h ? make_float2(w4.z, w4.w) : make_float2(w4.x, w4.y);
            const float2 oldp = h ? make_float2(old.z, old.w) : make_float2(old.x, old.y);
            const float2 alp = h ? make_float2(alr.z, alr.w) : make_float2(alr.x, alr.y);
            // (((p1[z-1] + p1[z+1]) + p1[x-1]) + p1[x+1]) and the same at distance 2   (:79; + is commutative)
            const float2 s1 = f2add(f2add(f2add(up1, dn1), make_float2(e[j + 1], e[j + 2])), make_float2(e[j + 3], e[j + 4]));
            const float2 s2 = f2add(f2add(f2add(up2, dn2), make_float2(e[j], e[j + 1])), make_float2(e[j + 4], e[j + 5]));
            // kappa = column profile in sponge columns, row profile elsewhere (get_Abc: columns override rows): selected
            // arithmetically -- 1 * kapz + 0 and 0 * kapz + kapx are exact -- so no predicate registers are tied up
            // (adjoint modes: kapx holds 1 - kappa_col and th.mz is negated, so the same FMA gives 1 - kappa directly)
            const float2 kp = f2fma(make_float2(th.mz[j], th.mz[j + 1]), make_float2(kapz, kapz), make_float2(kapx[j], kapx[j + 1]));
            float2 res;
            if (EXACT) {  // forward wavefield: one rounding per reference op
                // (products stay scalar FMULs: ptxas contracts a packed product feeding a packed sum into FFMA2 even with .rn --
                // and it also rewrites fma(a, b, -0) into that product first: measured, the checksum of the batch changed)
                const float2 lap = f2add(make_float2(__fmul_rn(c2, s1.x), __fmul_rn(c2, s1.y)),
                                         make_float2(__fmul_rn(c3, s2.x), __fmul_rn(c3, s2.y)));
                const float2 t1 = CTM ? make_float2(gk[j], gk[j + 1])
                                      : f2sub(f2add(make_float2(2.0f, 2.0f), make_float2(__fmul_rn(-5.0f, alp.x), __fmul_rn(-5.0f, alp.y))), kp);
                const float2 t2 = f2sub(make_float2(1.0f, 1.0f), kp);
                const float2 a1 = make_float2(__fmul_rn(t1.x, e[j + 2]), __fmul_rn(t1.y, e[j + 3]));
                const float2 a2 = make_float2(__fmul_rn(t2.x, oldp.x), __fmul_rn(t2.y, oldp.y));
                const float2 a3 = make_float2(__fmul_rn(alp.x, lap.x), __fmul_rn(alp.y, lap.y));
                res = f2add(f2sub(a1, a2), a3);
            } else {
                // adjoint field: no bit-parity requirement, so packed FMAs and the recurrence regrouped around
                //     L = (4/3) s1 - (1/12) s2 - 5 u1,    d = u1 - u2,    u = alpha L + (u1 + (1 - kappa) d)
                // (identical to (2 - 5 alpha - kappa) u1 - (1 - kappa) u2 + alpha lap: 8 instead of 10 packed operations).
                // It also hands the imaging terms over for free: u - (2-kappa) u1 + (1-kappa) u2 = alpha L, u2 - u1 = -d.
                const float2 cen = make_float2(e[j + 2], e[j + 3]);
                const float2 t2 = kp;  // 1 - kappa (see above)
                const float2 L = f2fma(make_float2(c2, c2), s1, f2fma(make_float2(c3, c3), s2, f2mul(cen, make_float2(-5.0f, -5.0f))));
                const float2 d = f2sub(cen, oldp);
                const float2 aL = f2mul(alp, L);
                res = f2add(aL, f2fma(d, t2, cen));
                if (IMG) {  // ga += p alpha L,  gk += p d  (the sign of gk is put right when the planes are written)
                    const float2 pp = h ? make_float2(pv.z, pv.w) : make_float2(pv.x, pv.y);
                    const float2 na = f2fma(pp, aL, make_float2(ga[j], ga[j + 1]));
                    const float2 nk = f2fma(pp, d, make_float2(gk[j], gk[j + 1]));
                    ga[j] = na.x; ga[j + 1] = na.y; gk[j] = nk.x; gk[j + 1] = nk.y;
                }
            }
            o[j] = res.x; o[j + 1] = res.y;
        }
        if (IMG) {
            tm_st4(im.tm + 4 * r, ga[0], ga[1], ga[2], ga[3]);
            tm_st4(im.tm + 4 * RMAX + 4 * r, gk[0], gk[1], gk[2], gk[3]);
        }
        const float4 out = make_float4(o[0], o[1], o[2], o[3]);
        if (r < nvalid) st4(pb + r * P, out);  // marching rows 0 .. nvalid-1 are the thread's own (both directions)
        if (r < 2 && push_now) st_async_v4(push_dst + r * P, push_bar, hp.cta, out);  // edge rows leave at once
        w0 = w1; w1 = w2; w2 = w3; w3 = w4;
    }
    if (IMG) tm_wait_st();  // the accumulators are read again a level later
}

// ADJ = false: forward wavefield p (source injection, receiver sampling -> seismograms, history of p).
// ADJ = true : adjoint field in the u-variable, u = alpha*q: the recurrence is the same, so the same
//              sweep runs it; level k of the loop is reverse time t = nt-1-k, the "source" is alpha * cotangent at the
//              receiver cells, the history receives u_{nt-1} .. u_1 (slot k), and sum_t u_t[src] w_t is accumulated
//              for the beta_dt term.  The imaging sums are formed afterwards by k_imaging from the two histories.
// One 512-thread CTA per SM (two 256-thread CTAs of two different shots per SM were measured slower and are gone).
//
// Debug option "perturb" (ClusterFwdArgs::perturb != 0): pseudo-random per-warp delays in front of every synchronisation
// point of a level -- halo waits, the sweep with its early halo pushes, the late pushes, the bulk-copy hand-over, the
// sampling / cotangent warp -- so that the orderings the design relies on (halo write-after-read covered by the data
// dependency, barriers armed one level ahead, phase parities) are exercised under schedules that never occur in an
// unperturbed run; results must stay bit-identical (tests/test_gpu_perturb.py).  compute-sanitizer is closed on the pool.
__device__ __noinline__ void jitter_sleep(const unsigned seed, const unsigned site, const unsigned t)
{
    unsigned h = seed ^ (blockIdx.x * 0x9E3779B9u) ^ ((threadIdx.x >> 5) * 0x85EBCA6Bu) ^ (t * 0xC2B2AE35u) ^ (site * 0x27D4EB2Fu);
    h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
    if ((h & 3) == 0) __nanosleep((h >> 8) & 0xFFF);  // a quarter of the (warp, level, site) triples sleep up to ~4 us
}
// PERT is a template parameter: even five out-of-line calls behind a uniform `seed != 0` branch per level cost the
// production kernels 3 ms of 38 (measured, profiles/README.md round 2), so they are compiled without any trace of it.
template <bool PERT>
__device__ __forceinline__ void jitter(const unsigned seed, const unsigned site, const unsigned t)
{
    if (PERT) jitter_sleep(seed, site, t);
}

// MODE = 2: the adjoint field as in MODE 1, but nothing is written to HBM per level: the imaging sums are formed inside the
//              sweep from the forward history (see fwd_sweep) and the kernel writes the per-shot planes Ga, Gk at the end of
//              a shot.  Replaces adjoint-field kernel + u-history + k_imaging (12 B of HBM traffic per cell-update) by one
//              kernel reading 4 B.  The cotangent injected at the receiver cells after the sweep is not in the row the sweep
//              saw; its share, sum_t p_t[rec] g_t, is kept per column in shared memory by the CTA's last warp (s_ginj).
template <int RMAX, int PITCH, int MODE, bool PERT>
__global__ void __launch_bounds__(kClusterThreads, 1) k_fwd_cluster(ClusterFwdArgs a, Grid g)
{
    constexpr int NT = kClusterThreads;
    constexpr bool ADJ = MODE >= 1, IMG = MODE == 2;
    // alpha rows kept in tensor memory beside the 8 accumulator columns per row (128 columns per thread)
#ifdef RDFWI_NO_FWD_TMEM
    constexpr bool CTM = false;
#else
    constexpr bool CTM = MODE == 0;   // forward mode: alpha and t1 rows in tensor memory (see fwd_sweep)
#endif
    constexpr int ATM = !IMG ? 0 : ((128 - 8 * RMAX) / 4 < RMAX ? (128 - 8 * RMAX) / 4 : RMAX);
    static_assert(!IMG || 8 * RMAX + 4 * ATM <= 128, "tensor-memory columns per thread");
    extern __shared__ __align__(128) float smem[];  // (no static shared memory: the launch sizes the block to the last byte)

    const int C = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
    const int cid = blockIdx.x / C, ncl = gridDim.x / C;
    // rows of the grid dealt over the CTAs of the cluster: nzp / C each, the remainder one row each to the first CTAs -- or
    // to the last ones (a.rows_flip, chosen by cluster_rows_flip()) when that keeps the source / receiver row two rows away
    // from every slab edge: a CTA whose patched row is an edge row cannot push its halo rows from inside the sweep, and
    // when a launch has few shots its neighbours (and theirs ...) spend a third of every level waiting for them
    const int base = g.nzp / C, rem = g.nzp % C;
    const int first_long = a.rows_flip ? C - rem : 0;  // CTAs first_long .. first_long + rem - 1 have base + 1 rows
    auto rows_of = [&](const int k) { return base + ((k >= first_long && k < first_long + rem) ? 1 : 0); };
    const int nrows = rows_of(rank);
    const int r0 = rank * base + (rank <= first_long ? 0 : (rank - first_long < rem ? rank - first_long : rem));
    const int up = rank == 0 ? C - 1 : rank - 1;
    const int dn = rank == C - 1 ? 0 : rank + 1;
    const int nrows_up = rows_of(up);

    const int pitch = PITCH > 0 ? PITCH : g.pitch;
    const int slab = (a.slabrows + 4) * pitch;  // floats per buffer: 2 halo rows, slab rows, 2 halo rows
    const int kap_off = 2 * slab;               // per-row sponge value of the CTA's rows (slabrows floats)
    // halo mbarriers: [buffer][top|bottom]; same offsets in every CTA of the cluster
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + ((kap_off + a.slabrows + 3) & ~3));
    // small read-only tables staged once per kernel so the epilogue never waits on global memory:
    // receiver CSR (columns -> receiver slots) and, when it fits, the wavelet
    uint32_t &s_tmem = *reinterpret_cast<uint32_t *>(bars + 4);  // base address of the tensor-memory allocation
    float *s_prow = reinterpret_cast<float *>(bars + 4) + 4;  // [2][pitch] receiver row of the forward level (MODE 2), 16-byte aligned
    float *s_ginj = s_prow + 2 * pitch;                   // [pitch] sum_t p_t[rec] g_t per column (MODE 2)
    int *s_rec_ptr = reinterpret_cast<int *>(s_ginj + pitch);
    int *s_rec_idx = s_rec_ptr + g.nxp + 1;
    int *s_rec_col = s_rec_idx + g.nrec;  // padded column of every receiver (the survey's igx)
    float *s_cot = reinterpret_cast<float *>(s_rec_col + g.nrec);  // [2][nxp] per-column cotangent sums (adjoint mode)
    float *s_raw = s_cot + 2 * g.nxp;  // [2][nrec] cotangent rows as they sit in HBM, landed by cp.async (adjoint mode)
    float *s_wav = s_raw + 2 * g.nrec;
    const bool wav_in_smem = a.wav_smem != 0;
    float *s_ring = smem + (((int)(s_wav - smem) + (wav_in_smem ? a.nt : 0) + 3) & ~3);  // MODE 2: [slots][threads] 16 bytes, see fwd_sweep
    const bool st1 = a.st == 1;  // every level is sampled (all configs of the reference): no integer division on the level's critical path

    const int tid = threadIdx.x, lane_id = tid & 31;
    const int grp = tid / g.q4, col = tid - grp * g.q4;
    const bool active = grp < a.ngroups;
    const bool warp_active = (tid - lane_id) < a.ngroups * g.q4;  // warps without any owner lane skip the sweep
    SweepThread th;
    th.x = col * 4;
    th.la = grp * RMAX;
    th.lb = !active ? th.la : (th.la + RMAX < nrows ? th.la + RMAX : nrows);
    th.swap = (lane_id & 8) != 0;
    th.edgeL = lane_id == 0 || col == 0;
    th.edgeR = lane_id == 31 || col == g.q4 - 1;
    th.eL = col == 0 ? g.nxp - 2 : th.x - 2;
    th.eR = col == g.q4 - 1 ? pitch - g.nxp : th.x + 4;
    int xc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        xc[j] = th.x + j >= g.nxp ? th.x + j - g.nxp : th.x + j;
        th.colsp[j] = sponge_index(xc[j], g.nxp, g.nbc) >= 0;
        th.mz[j] = th.colsp[j] ? 0.0f : (ADJ ? -1.0f : 1.0f);  // adjoint modes form 1 - kappa: (1 - kappa_col) - mz kappa_row
    }
    // rows this thread must handle in the epilogue (local row index, or -1)
    th.src_lr = (g.isz - r0 >= th.la && g.isz - r0 < th.lb) ? g.isz - r0 : -1;
    th.rec_lr = (g.igz - r0 >= th.la && g.igz - r0 < th.lb) ? g.igz - r0 : -1;
    // Warps holding lanes of the group that owns the CTA's last row march upwards from that row, all others
    // downwards from their first row; idle lanes shadow the last group (their stores are predicated off).
    const int last_grp = (nrows - 1) / RMAX;
    const bool rev = __any_sync(0xffffffffu, active && grp == last_grp && last_grp > 0) != 0;
    const int la_eff = active ? th.la : last_grp * RMAX;
    const int lb_eff = active ? th.lb : nrows;
    const int l0 = rev ? (lb_eff > la_eff ? lb_eff - 1 : la_eff) : la_eff;  // first row in marching order
    th.lac = l0;
    // which halos the warp reads during a sweep (rows < 0 / rows nrows, nrows+1)
    bool rd_top, rd_bot;
    {
        const int lo = rev ? l0 - (RMAX + 1) : l0 - 2, hi = rev ? l0 + 2 : l0 + RMAX + 1;
        rd_top = __any_sync(0xffffffffu, lo < 0) != 0;
        rd_bot = __any_sync(0xffffffffu, hi >= nrows && lo < nrows + 2) != 0;
    }
    // early halo pushes: the CTA's first / last two rows are marching rows 0,1 of their owner threads --
    // unless the source row is one of them (it is patched in the epilogue, after which it is sent)
    const int patched_row = (ADJ ? g.igz : g.isz) - r0;  // the row whose cells are modified in the epilogue
    const bool src_on_edge = patched_row >= 0 && patched_row < nrows && (patched_row < 2 || patched_row >= nrows - 2);
    HaloPush hp;
    hp.early = false; hp.dst = 0; hp.cta = 0; hp.bar = 0;
    if (active && !src_on_edge) {
        if (!rev && th.la == 0 && th.lb >= 2) {          // rows 0,1 -> bottom halo of the CTA above
            hp.early = true; hp.dst = (2 + nrows_up) * pitch; hp.cta = (uint32_t)up; hp.bar = 1;
        } else if (rev && th.lb == nrows && th.lb - th.la >= 2) {  // rows nrows-1, nrows-2 -> top halo of the CTA below
            hp.early = true; hp.dst = pitch; hp.cta = (uint32_t)dn; hp.bar = 0;
        }
    }
    // edge rows this thread owns but does not send early (bit h: top row h; bit 2+h: row nrows-2+h)
    int late = 0;
    for (int h = 0; h < 2; ++h) {
        if (h >= th.la && h < th.lb && !(hp.early && hp.bar == 1)) late |= 1 << h;
        const int rb = nrows - 2 + h;
        if (rb >= th.la && rb < th.lb && !(hp.early && hp.bar == 0)) late |= 4 << h;
    }
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t hist_shot = (size_t)a.nt * g.level;  // the history keeps every level
    const uint32_t halo_bytes = (uint32_t)(2 * pitch * sizeof(float));
    // halo phases consumed per shot from buffer 1 (levels 1,3,..) and buffer 0 (levels 2,4,..)
    const int uses1 = a.nt / 2, uses0 = (a.nt - 1) / 2;

    if (tid == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);
    }
    uint32_t tm = 0;
    if (IMG || CTM) {
        if (tid < 32) tm_alloc_all(&s_tmem);
        tm_fence_before_sync();
        __syncthreads();
        tm_fence_after_sync();
        tm = s_tmem + ((uint32_t)(((tid >> 5) & 3) * 32) << 16) + (uint32_t)((tid >> 7) * 128);
        if (IMG) {
#pragma unroll
            for (int c = 0; c < 8 * RMAX; c += 4) tm_st4(tm + c, 0.f, 0.f, 0.f, 0.f);
            tm_wait_st();
        }
    }
    for (int i = tid; i <= g.nxp; i += NT) s_rec_ptr[i] = a.rec_ptr[i];
    for (int i = tid; i < g.nrec; i += NT) { s_rec_idx[i] = a.rec_idx[i]; s_rec_col[i] = a.rec_col[i]; }
    if (ADJ)
        for (int i = tid; i < 2 * g.nxp; i += NT) s_cot[i] = 0.0f;  // (receiver-major path: columns without a receiver stay 0)
    const bool rec_simple = a.rec_simple != 0;  // at most one receiver per column (every configuration of the reference)
    if (wav_in_smem)
        for (int i = tid; i < a.nt; i += NT) s_wav[i] = a.wavelet[i];
    __syncthreads();

    int shot_iter = 0;
    for (int shot = cid; shot < a.nshots; shot += ncl, ++shot_iter) {
        const int gshot = a.shot0 + shot;  // index into per-shot inputs/outputs other than the history of this launch
        const int b = gshot / g.ns, s = gshot - b * g.ns;
        // p_{-1} = p_{-2} = 0 (halo rows included); sponge tables of this model
        for (int i = tid; i < 2 * slab; i += NT) smem[i] = 0.0f;
        const float *kap_b = a.kap + (size_t)b * (g.nbc + 1);
        for (int i = tid; i < a.slabrows; i += NT) {
            const int kz = sponge_index(r0 + i, g.nzp, g.nbc);
            smem[kap_off + i] = (i < nrows && kz >= 0) ? kap_b[kz] : 0.0f;
        }
        float4 al[RMAX];  // alpha of the thread's rows, in marching order (MODE 2: the last ATM rows live in tensor memory)
#pragma unroll
        for (int r = 0; r < RMAX; ++r) {
            const int lrow = rev ? l0 - r : l0 + r;
            const float4 av = (lrow >= th.la && lrow < th.lb) ? ld4(a.alpha + (size_t)b * g.level + (size_t)(r0 + lrow) * pitch + th.x) : zero4;
            if (CTM) tm_st4(tm + 4 * r, av.x, av.y, av.z, av.w);
            else if (r < RMAX - ATM) al[r] = av;
            else tm_st4(tm + 8 * RMAX + 4 * (r - (RMAX - ATM)), av.x, av.y, av.z, av.w);
        }
        if ((IMG && ATM > 0) || CTM) tm_wait_st();
        float kapx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kx = sponge_index(xc[j], g.nxp, g.nbc);
            kapx[j] = kx >= 0 ? kap_b[kx] : 0.0f;
            if (ADJ) kapx[j] = 1.0f - kapx[j];
        }
        const int xs = a.isx[s];
        int src_mask = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) src_mask |= (th.src_lr >= 0 && xc[j] == xs) ? (1 << j) : 0;
        const float bsrc = (ADJ && !IMG) ? 0.0f : a.beta_src[gshot];
        const float *ph_shot = IMG ? a.phist + (size_t)shot * hist_shot : nullptr;  // forward history of this shot (MODE 2)
        // adjoint mode: alpha of the receiver-row cells, lane of the (non-image) source cell, beta_dt accumulator
        float4 al_rec = zero4;
        int src_lane = -1;
        float gb = 0.0f;
        if (ADJ) {
            if (th.rec_lr >= 0) al_rec = ld4(a.alpha + (size_t)b * g.level + (size_t)(r0 + th.rec_lr) * pitch + th.x);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (th.src_lr >= 0 && th.x + j < g.nxp && th.x + j == xs) src_lane = j;
        }
        const bool has_rec_row = g.igz >= r0 && g.igz < r0 + nrows;
        // Cotangent staging, executed by the CTA's last warp (which owns no rows), two levels deep: the row of reverse level
        // tr is fetched from HBM with fire-and-forget 4-byte cp.async into s_raw[buf] while the level before it is being
        // swept, and summed per padded column (receivers may share a column) into s_cot[buf] one level later.  Nothing on
        // the way waits for an HBM round trip: fetched by dependent loads, the row cost ~10 000 cycles per level on
        // Marmousi-width grids -- more than the sweep (tools/trace_levels.py).
        auto fetch_cot = [&](const int tr, const int buf) {
            if (tr < 0 || (!st1 && tr % a.st != 0)) return;
            const float *gt = a.cot + ((size_t)gshot * g.nt_out + (st1 ? tr : tr / a.st)) * g.nrec;
            float *dst = s_raw + buf * g.nrec;
            for (int r = lane_id; r < g.nrec; r += 32)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst + r)), "l"(gt + r) : "memory");
            if (IMG) {  // the receiver row of the forward level the cotangent row pairs with
                const float *gp = ph_shot + (size_t)tr * g.level + (size_t)g.igz * pitch;
                float *pd = s_prow + buf * pitch;
                for (int c = lane_id; c < g.q4; c += 32)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(pd + 4 * c)), "l"(gp + 4 * c) : "memory");
            }
        };
        auto sum_cot = [&](const int tr, const int buf) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncwarp();
            if (tr < 0 || (!st1 && tr % a.st != 0)) return;
            const float *raw = s_raw + buf * g.nrec;
            float *dst = s_cot + buf * g.nxp;
            if (rec_simple) {
                // receiver-major: independent iterations, three per lane for 70 receivers -- the column-major loop below is ten
                // dependent shared-memory round trips per lane, 4 500 cycles per level on a warp that every level's barrier
                // waits for: longer than a whole 4-row sweep when the launch has few shots (tools/trace_levels.py, round 2)
#pragma unroll 2
                for (int r = lane_id; r < g.nrec; r += 32) {
                    const int xx = s_rec_col[r];
                    const float c = raw[r];
                    dst[xx] = c;
                    if (IMG) s_ginj[xx] += c * s_prow[buf * pitch + xx];
                }
                return;
            }
#pragma unroll 4
            for (int xx = lane_id; xx < g.nxp; xx += 32) {
                float acc = 0.0f;
                for (int k = s_rec_ptr[xx]; k < s_rec_ptr[xx + 1]; ++k) acc += raw[s_rec_idx[k]];
                dst[xx] = acc;
                if (IMG) s_ginj[xx] += acc * s_prow[buf * pitch + xx];
            }
        };
        if (ADJ && tid >= NT - 32 && has_rec_row) {  // level 0 of the loop is reverse time nt-1
            if (IMG)
                for (int xx = lane_id; xx < pitch; xx += 32) s_ginj[xx] = 0.0f;
            fetch_cot(a.nt - 1, 0);
            sum_cot(a.nt - 1, 0);
            fetch_cot(a.nt - 2, 1);
        }
        if (IMG && tid == 0)
            for (int d = 0; d < kImgPrefetchLevels && d < a.nt; ++d)
                l2_prefetch_bulk(ph_shot + (size_t)(a.nt - 1 - d) * g.level + (size_t)r0 * pitch, (uint32_t)(nrows * pitch * sizeof(float)));
        __syncthreads();
        cluster_sync_all();  // shot boundary: every CTA has finished the previous shot and cleared its buffers
        if (CTM && warp_active) {
            // t1 = (2 + (-5 alpha)) - kappa per cell, one rounding per reference operation (solvers/pde.py:69), kappa selected
            // as in the sweep; the sponge row table was written before the barrier above
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                float av[4];
                tm_ld4(tm + 4 * r, av);
                const float kapz = smem[kap_off + (rev ? l0 - r : l0 + r)];
                float dummy[4] = {0.f, 0.f, 0.f, 0.f};
                tm_wait_ld(av, dummy);
                float t1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float kp = __fmaf_rn(th.mz[j], kapz, kapx[j]);
                    t1[j] = __fsub_rn(__fadd_rn(2.0f, __fmul_rn(-5.0f, av[j])), kp);
                }
                tm_st4(tm + 4 * RMAX + 4 * r, t1[0], t1[1], t1[2], t1[3]);
            }
            tm_wait_st();
        }

        // one time level: p_{t-1} in buffer `cur`, p_{t-2} in `prv`, p_t overwrites p_{t-2}
        // optional per-warp timeline (debug option "trace_ptr"): clock64 stamps of levels 100..103 of cluster 0's first shot
        auto stamp = [&](const int t, const int phase) {
            if (a.trace != nullptr && cid == 0 && shot_iter == 0 && t >= 100 && t < 104 && lane_id == 0)
                a.trace[(((size_t)rank * 4 + (t - 100)) * (NT / 32) + (tid >> 5)) * 6 + phase] = clock64();
        };
        auto level = [&](const int t, const int cur, const int prv) {
            const int cbuf = cur == 0 ? 0 : 1, pbuf = 1 - cbuf;
            stamp(t, 0);
            jitter<PERT>(a.perturb, 0, (unsigned)t);
            uint64_t *bar_top = bars + 2 * cbuf, *bar_bot = bars + 2 * cbuf + 1;
            const bool sends = t + 1 < a.nt;  // the last level of a shot has no consumer
            // Arm the barriers of the buffer written NOW (consumed at level t+1) before the neighbours' rows can land: a
            // waiter that finds its phase already complete returns at once, whereas arming at the consumer's level start
            // raced with the waiters and cost them a ~3600-cycle try_wait time-out per level (tools/trace_levels.py).
            if (tid == 0 && sends) {
                mbar_expect_tx(bars + 2 * pbuf, halo_bytes);
                mbar_expect_tx(bars + 2 * pbuf + 1, halo_bytes);
            }
            if (t >= 1) {
                // the halos of `cur` were sent by the neighbours early in level t-1 (t = 0: zero initial state)
                const uint32_t parity = (uint32_t)((shot_iter * (cbuf ? uses1 : uses0) + (t - 1) / 2) & 1);
                if (warp_active) {
                    if (rd_top) mbar_wait(bar_top, parity);
                    if (rd_bot) mbar_wait(bar_bot, parity);
                }
            }
            stamp(t, 1);
            jitter<PERT>(a.perturb, 1, (unsigned)t);
            const int p0 = prv + 2 * pitch + th.x;
            const int trev = a.nt - 1 - t;  // adjoint mode: the reverse-time level this iteration computes
            // cotangent pipeline of the last warp: sum the row of the NEXT reverse level (fetched one level ago), then fetch
            // the row of the level after it
            if (ADJ && tid >= NT - 32 && has_rec_row && t + 1 < a.nt) {
                sum_cot(trev - 1, (t + 1) & 1);
                fetch_cot(trev - 2, t & 1);
            }
            if (IMG && tid == 0 && trev >= kImgPrefetchLevels)  // pull the slab of the forward level three sweeps ahead into L2
                l2_prefetch_bulk(ph_shot + (size_t)(trev - kImgPrefetchLevels) * g.level + (size_t)r0 * pitch, (uint32_t)(nrows * pitch * sizeof(float)));
            if (warp_active) {
                const uint64_t *push_bar = bars + 2 * pbuf + hp.bar;  // barrier of the buffer written now, at the receiver
                ImgThread im;
                im.pg = IMG ? ph_shot + (size_t)trev * g.level + (size_t)(r0 + l0) * pitch + th.x : nullptr;
                im.tm = tm;
                im.ring = s_ring + tid * 4;
                im.first = t == 0;
                im.next = trev > 0;
                im.level = (size_t)g.level;
                if (rev) fwd_sweep<RMAX, PITCH, -1, !ADJ, IMG, ATM, CTM>(smem, cur, prv, kap_off, pitch, l0, th, al, kapx, hp, push_bar, hp.early && sends, im);
                else fwd_sweep<RMAX, PITCH, 1, !ADJ, IMG, ATM, CTM>(smem, cur, prv, kap_off, pitch, l0, th, al, kapx, hp, push_bar, hp.early && sends, im);
                stamp(t, 2);
                jitter<PERT>(a.perturb, 2, (unsigned)t);
                if (ADJ) {
                    if (th.rec_lr >= 0 && (st1 || trev % a.st == 0)) {  // u_t[rec] += alpha * g_t  (adjoint of the gather, :83)
                        float4 v = ld4(smem + p0 + th.rec_lr * pitch);
                        const float *cc = s_cot + (t & 1) * g.nxp;
                        v.x += al_rec.x * cc[xc[0]]; v.y += al_rec.y * cc[xc[1]]; v.z += al_rec.z * cc[xc[2]]; v.w += al_rec.w * cc[xc[3]];
                        st4(smem + p0 + th.rec_lr * pitch, v);
                    }
                    if (src_lane >= 0) {  // adjoint of the source injection (:81)
                        const float4 v = ld4(smem + p0 + th.src_lr * pitch);
                        gb += (src_lane == 0 ? v.x : src_lane == 1 ? v.y : src_lane == 2 ? v.z : v.w) * (wav_in_smem ? s_wav[trev] : a.wavelet[trev]);
                    }
                } else if (src_mask != 0) {  // p[src] += beta_dt[src] * wavelet[t]   (solvers/pde.py:80-81)
                    const float src_add = __fmul_rn(bsrc, wav_in_smem ? s_wav[t] : a.wavelet[t]);
                    float4 v = ld4(smem + p0 + th.src_lr * pitch);
                    if (src_mask & 1) v.x = __fadd_rn(v.x, src_add);
                    if (src_mask & 2) v.y = __fadd_rn(v.y, src_add);
                    if (src_mask & 4) v.z = __fadd_rn(v.z, src_add);
                    if (src_mask & 8) v.w = __fadd_rn(v.w, src_add);
                    st4(smem + p0 + th.src_lr * pitch, v);
                }
                if (late != 0 && sends) {  // edge rows that could not leave from inside the sweep
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
            stamp(t, 3);
            jitter<PERT>(a.perturb, 3, (unsigned)t);
            if (a.hist != nullptr) {
                fence_proxy_async();             // slab writes -> visible to the bulk-copy engine
                if (tid == 0) bulk_wait_read();  // the copy of the previous level has finished reading its buffer
            }
            stamp(t, 4);
            __syncthreads();                     // rows of level t are complete CTA-wide
            stamp(t, 5);
            jitter<PERT>(a.perturb, 4, (unsigned)t);
            // Receiver sampling (solvers/pde.py:82-83, after the source injection) is done by the CTA's last warp -- which
            // usually owns no rows -- from the finished level while the other warps already sweep the next one (that
            // buffer is read-only until the barrier after next).  In the owner threads' epilogue it sat on the critical
            // path of the whole cluster: ~1900 of 9000 cycles per level (tools/trace_levels.py).
            if (!ADJ && a.seis != nullptr && tid >= NT - 32 && (st1 || t % a.st == 0) && g.igz >= r0 && g.igz < r0 + nrows) {
                float *seis_t = a.seis + ((size_t)gshot * g.nt_out + (st1 ? t : t / a.st)) * g.nrec;
                const float *row = smem + prv + (2 + g.igz - r0) * pitch;
                if (rec_simple) {
#pragma unroll 2
                    for (int r = lane_id; r < g.nrec; r += 32) seis_t[r] = row[s_rec_col[r]];  // coalesced
                } else {
                    for (int xx = lane_id; xx < g.nxp; xx += 32)
                        for (int k = s_rec_ptr[xx]; k < s_rec_ptr[xx + 1]; ++k) seis_t[s_rec_idx[k]] = row[xx];
                }
            }
            if (a.hist != nullptr && tid == 0)
                bulk_store(a.hist + (size_t)shot * hist_shot + (size_t)t * g.level + (size_t)r0 * pitch,
                           smem + prv + 2 * pitch, (uint32_t)(nrows * pitch * sizeof(float)));
        };
        for (int t = 0, cur = 0; t < a.nt; ++t, cur = slab - cur) level(t, cur, slab - cur);
        float gb_over_al = 0.0f;
        if (ADJ && src_lane >= 0) {
            const float4 alv = ld4(a.alpha + (size_t)b * g.level + (size_t)(r0 + th.src_lr) * pitch + th.x);
            gb_over_al = gb / (src_lane == 0 ? alv.x : src_lane == 1 ? alv.y : src_lane == 2 ? alv.z : alv.w);
            a.Gb[gshot] = gb_over_al;
        }
        if (a.hist != nullptr && tid == 0) bulk_wait_read();
        __syncthreads();
        if (IMG && warp_active) {
            // imaging planes of this shot (what k_imaging writes): Ga = sums / alpha^2, Gk = sums / alpha; the accumulators go
            // back to zero for the cluster's next shot
            const int nvalid = th.lb - th.la;
#pragma unroll
            for (int r = 0; r < RMAX; ++r) {
                float ga[4], gk[4];
                tm_ld4(tm + 4 * r, ga);
                tm_ld4(tm + 4 * RMAX + 4 * r, gk);
                tm_wait_ld(ga, gk);
                tm_st4(tm + 4 * r, 0.f, 0.f, 0.f, 0.f);
                tm_st4(tm + 4 * RMAX + 4 * r, 0.f, 0.f, 0.f, 0.f);
                if (r < nvalid) {
                    const int lrow = rev ? l0 - r : l0 + r;
                    const size_t off = (size_t)(r0 + lrow) * pitch + th.x;
                    const float4 av = ld4(a.alpha + (size_t)b * g.level + off);
                    const float a4[4] = {av.x, av.y, av.z, av.w};
                    if (lrow == th.rec_lr) {  // the injected cotangent's share: alpha * sum_t p_t g_t
#pragma unroll
                        for (int j = 0; j < 4; ++j) ga[j] += a4[j] * s_ginj[xc[j]];
                    }
                    if (lrow == th.src_lr && src_lane >= 0) {  // the forward recurrence had the extra term beta_src w_t at the source cell
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (j == src_lane) ga[j] -= bsrc * (gb_over_al * a4[j]);
                    }
                    st4(a.Ga + (size_t)gshot * g.level + off, make_float4(ga[0] / (a4[0] * a4[0]), ga[1] / (a4[1] * a4[1]), ga[2] / (a4[2] * a4[2]), ga[3] / (a4[3] * a4[3])));
                    st4(a.Gk + (size_t)gshot * g.level + off, make_float4(-gk[0] / a4[0], -gk[1] / a4[1], -gk[2] / a4[2], -gk[3] / a4[3]));
                }
            }
            tm_wait_st();
        }
        __syncthreads();
    }
    if (tid == 0) bulk_wait_all();
    cluster_sync_all();  // no CTA exits while a neighbour may still address its shared memory
    if ((IMG || CTM) && tid < 32) tm_free_all(s_tmem);
}

}  // namespace

// Smallest cluster (1..8 CTAs, or the non-portable 16) whose slabs fit when every thread marches R rows.
static bool cluster_config_rows(const Plan &p, const int R, const bool allow16, ClusterConfig *cfg, const bool img)
{
    const Grid &g = p.g;
    const int nthreads = kClusterThreads;
    const int groups_max = nthreads / g.q4;
    if (groups_max < 1) return false;
    int max_smem = 0;
    if (device_attr(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, p.device) != cudaSuccess) return false;
    for (int C = 1; C <= 16; ++C) {
        if (C > 8 && C != 16) continue;  // 1..8 are portable cluster sizes, 16 needs the non-portable opt-in
        // 16-CTA clusters (31-row slabs, 9 co-resident clusters): in round 1 they lost to the tiled per-level engine on the
        // grids that need them (interior 256^2: 1.02e11 vs 1.25e11 pairs/s at 16 shots); with the resident adjoint and the
        // coefficients in tensor memory they win by half (2.00e11 / 2.42e11 / 1.86e11 at 16 / 64 / 256 shots against
        // 1.25e11 / 1.54e11 / 0.91e11, profiles/sweep_r2b_n1.md), so they are part of the automatic choice now
        (void)allow16;
        if (p.cluster_size > 0 && C != p.cluster_size) continue;
        if (g.nzp / C < 2) break;
        const int maxrows = (g.nzp + C - 1) / C;
        const int ngroups = (maxrows + R - 1) / R;  // each thread marches R rows
        if (ngroups > groups_max) continue;
        const int slabrows = ngroups * R;  // >= maxrows: rows past the slab are computed but never stored
        size_t smem = ((size_t)2 * (slabrows + 4) * g.pitch + slabrows + 16 + 3 * g.pitch + g.nxp + 1 + 2 * g.nrec + 2 * g.nxp + 2 * g.nrec) * sizeof(float);
        if (img) smem += (size_t)img_ring_slots(R) * kClusterThreads * 16 + 16;  // resident imaging: slots staging the forward rows
        if (smem > (size_t)max_smem) continue;
        const size_t room = (size_t)max_smem;
        cfg->wav_smem = smem + (size_t)p.nt * sizeof(float) <= room;
        if (cfg->wav_smem) smem += (size_t)p.nt * sizeof(float);
        cfg->img = img;
        cfg->nthreads = nthreads;
        cfg->C = C; cfg->maxrows = maxrows; cfg->ngroups = ngroups; cfg->slabrows = slabrows; cfg->smem = smem;
        cfg->rmax = R;
        return true;
    }
    return false;
}

bool cluster_config(const Plan &p, ClusterConfig *cfg, int nshots, bool img)
{
    if (p.cluster_rows > 0) return cluster_config_rows(p, p.cluster_rows, true, cfg, img);  // forced (tests, tuning)
    if (!cluster_config_rows(p, kClusterRowsMax, false, cfg, img)) return false;
    if (nshots <= 0 || p.cluster_size != 0) return true;
    // Few shots (one model of the reference's configs has 5): the throughput configuration would occupy nshots * C of
    // the 148 SMs and every level would still cost a full 13-row sweep.  A level is latency-bound (one shot's level takes
    // the same time alone as among 33 co-resident ones), so spread a shot over more CTAs with fewer rows per thread --
    // as long as every shot of the launch still gets its own co-resident cluster.
    if (2 * nshots * cfg->C > 148) return true;
    // Among the wide configurations that give every shot its own co-resident cluster, the cheapest level: a level is the
    // sweep of the busiest scheduler -- (row-owning warps per scheduler) x (rows marched + ~1 row of start-up loads) --
    // and two warps per scheduler hide too little latency (measured, one OpenFWI model on 16-CTA clusters: 4 rows 3.73 ms,
    // 5 rows 3.39 ms, 6 rows 3.72 ms, 7 rows 3.66 ms per gradient; one Marmousi-shaped model: 5 rows 4.16, 7 rows 4.21, 6 rows 4.50).
    static const int kWideRows[3] = {4, 5, 7};
    float best = 1e30f;
    const int base_C = cfg->C;
    for (int i = 0; i < 3; ++i) {
        ClusterConfig wide;
        if (!cluster_config_rows(p, kWideRows[i], true, &wide, img) || wide.C <= base_C) continue;
        if (fwd_cluster_wave(p, wide) < nshots) continue;
        const int warps = (wide.ngroups * p.g.q4 + 31) / 32, per_sched = (warps + 3) / 4;
        const float cost = per_sched * (kWideRows[i] + 1.0f) + (per_sched < 3 ? 6.0f : 0.0f);
        if (cost < best) { best = cost; *cfg = wide; }
    }
    return true;
}

template <int R, int PITCH, int MODE, bool PERT>
static cudaError_t launch_fwd_cluster_t(const Plan &p, const ClusterConfig &cc, ClusterFwdArgs a, cudaStream_t st, int *wave_only)
{
    constexpr int NT = kClusterThreads;
    auto kernel = k_fwd_cluster<R, PITCH, MODE, PERT>;
    a.slabrows = cc.slabrows; a.ngroups = cc.ngroups; a.wav_smem = cc.wav_smem ? 1 : 0;

    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cc.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = cc.smem;
    cfg.stream = st;
    // persistent: as many clusters as can be co-resident (or one per shot if fewer shots).  The function attributes and
    // the occupancy answer are set / asked once per (instantiation, device, cluster size, shared memory).
    static std::mutex mu;                       // function attributes are per (function, device): raise them monotonically
    static size_t smem_set[64] = {0};
    static bool nonportable_set[64] = {false};
    static thread_local int c_dev = -1, c_C = 0, c_max = 0;
    static thread_local size_t c_smem = 0;
    cudaError_t e;
    if (c_dev != p.device || c_C != cc.C || c_smem != cc.smem) {
        {
            std::lock_guard<std::mutex> lock(mu);
            const int d = p.device & 63;
            if (cc.smem > smem_set[d]) {
                e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cc.smem);
                if (e != cudaSuccess) return e;
                smem_set[d] = cc.smem;
            }
            if (cc.C > 8 && !nonportable_set[d]) {
                e = cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
                if (e != cudaSuccess) return e;
                nonportable_set[d] = true;
            }
        }
        int sms = 148;
        device_attr(&sms, cudaDevAttrMultiProcessorCount, p.device);
        cfg.gridDim = dim3((unsigned)(sms / cc.C * cc.C));
        int q = 0;
        e = cudaOccupancyMaxActiveClusters(&q, kernel, &cfg);
        if (e != cudaSuccess) return e;
        if (q < 1) return cudaErrorLaunchOutOfResources;
        c_dev = p.device; c_C = cc.C; c_smem = cc.smem; c_max = q;
    }
    const int max_clusters = c_max;
    if (wave_only != nullptr) { *wave_only = max_clusters; return cudaSuccess; }
    const int ncl = max_clusters < a.nshots ? max_clusters : a.nshots;
    cfg.gridDim = dim3((unsigned)(ncl * cc.C));
    e = cudaLaunchKernelEx(&cfg, kernel, a, p.g);
    count_launch();
    return e;
}

template <int R, int PITCH>
static cudaError_t dispatch_fwd_cluster_rp(const Plan &p, const ClusterConfig &cc, const ClusterFwdArgs &a, cudaStream_t st, int *wave_only)
{
#ifndef RDFWI_DEV_FAST  // (development builds compile the production instantiations only)
    if (a.perturb != 0)  // debug instantiations (schedule perturbation)
        return a.adj_mode == 2   ? launch_fwd_cluster_t<R, PITCH, 2, true>(p, cc, a, st, wave_only)
               : a.adj_mode == 1 ? launch_fwd_cluster_t<R, PITCH, 1, true>(p, cc, a, st, wave_only)
                                 : launch_fwd_cluster_t<R, PITCH, 0, true>(p, cc, a, st, wave_only);
#endif
    return a.adj_mode == 2   ? launch_fwd_cluster_t<R, PITCH, 2, false>(p, cc, a, st, wave_only)
           : a.adj_mode == 1 ? launch_fwd_cluster_t<R, PITCH, 1, false>(p, cc, a, st, wave_only)
                             : launch_fwd_cluster_t<R, PITCH, 0, false>(p, cc, a, st, wave_only);
}

// throughput configuration only (13 rows, no perturbation): the sweep-table grids interior 128^2 and 256^2 (BASELINE config 5)
template <int PITCH>
static cudaError_t dispatch_fwd_cluster_sweep_grid(const Plan &p, const ClusterConfig &cc, const ClusterFwdArgs &a, cudaStream_t st, int *wave_only)
{
    return a.adj_mode == 2   ? launch_fwd_cluster_t<kClusterRowsMax, PITCH, 2, false>(p, cc, a, st, wave_only)
           : a.adj_mode == 1 ? launch_fwd_cluster_t<kClusterRowsMax, PITCH, 1, false>(p, cc, a, st, wave_only)
                             : launch_fwd_cluster_t<kClusterRowsMax, PITCH, 0, false>(p, cc, a, st, wave_only);
}

template <int R>
static cudaError_t dispatch_fwd_cluster_r(const Plan &p, const ClusterConfig &cc, const ClusterFwdArgs &a, cudaStream_t st, int *wave_only)
{
    // production grids get immediate row offsets (OpenFWI 310+2, Marmousi/Overthrust 430+2); the specialised kernels read
    // x-neighbour PAIRS from shared memory, which needs an even padded width
    switch ((p.g.nxp & 1) == 0 ? p.g.pitch : 0) {
        case 312: return dispatch_fwd_cluster_rp<R, 312>(p, cc, a, st, wave_only);
#ifndef RDFWI_DEV_FAST
        case 432: return dispatch_fwd_cluster_rp<R, 432>(p, cc, a, st, wave_only);
        case 368: if (R == kClusterRowsMax && a.perturb == 0) return dispatch_fwd_cluster_sweep_grid<368>(p, cc, a, st, wave_only);
                  return dispatch_fwd_cluster_rp<R, 0>(p, cc, a, st, wave_only);
        case 496: if (R == kClusterRowsMax && a.perturb == 0) return dispatch_fwd_cluster_sweep_grid<496>(p, cc, a, st, wave_only);
                  return dispatch_fwd_cluster_rp<R, 0>(p, cc, a, st, wave_only);
        default: return dispatch_fwd_cluster_rp<R, 0>(p, cc, a, st, wave_only);
#else
        default: return cudaErrorInvalidValue;
#endif
    }
}

static cudaError_t dispatch_fwd_cluster(const Plan &p, const ClusterConfig &cc, const ClusterFwdArgs &a, cudaStream_t st, int *wave_only)
{
    switch (cc.rmax) {  // rows marched per thread: 13 for throughput, 7 / 4 on wider clusters when the shots are few
        case kClusterRowsMax: return dispatch_fwd_cluster_r<kClusterRowsMax>(p, cc, a, st, wave_only);
#ifndef RDFWI_DEV_FAST
        case 7: return dispatch_fwd_cluster_r<7>(p, cc, a, st, wave_only);
#endif
        case 4: return dispatch_fwd_cluster_r<4>(p, cc, a, st, wave_only);
        case 5: return dispatch_fwd_cluster_r<5>(p, cc, a, st, wave_only);
        default: return cudaErrorInvalidValue;
    }
}

// Which CTAs of a C-CTA cluster take the nzp % C extra rows (0: the first ones, 1: the last ones): the choice that keeps the
// source row and the receiver row at least two rows inside a slab, if there is one (see k_fwd_cluster).
static int cluster_rows_flip(const Grid &g, const int C)
{
    const int base = g.nzp / C, rem = g.nzp % C;
    auto on_edge = [&](const int flip, const int row) {
        const int first_long = flip ? C - rem : 0;
        for (int k = 0, r0 = 0; k < C; ++k) {
            const int n = base + ((k >= first_long && k < first_long + rem) ? 1 : 0);
            if (row >= r0 && row < r0 + n) return row - r0 < 2 || row - r0 >= n - 2;
            r0 += n;
        }
        return false;
    };
    if (rem == 0) return 0;
    const bool bad0 = on_edge(0, g.isz) || on_edge(0, g.igz), bad1 = on_edge(1, g.isz) || on_edge(1, g.igz);
    return (bad0 && !bad1) ? 1 : 0;
}

cudaError_t launch_fwd_cluster(const Plan &p, const ClusterConfig &cc, ClusterFwdArgs a, cudaStream_t st)
{
    a.perturb = (unsigned)p.perturb;
    a.rows_flip = cluster_rows_flip(p.g, cc.C);
    a.rec_col = p.d_igx;
    a.rec_simple = p.rec_simple ? 1 : 0;
    const_cast<Plan &>(p).last_fwd_C = cc.C;
    const_cast<Plan &>(p).last_fwd_rows = cc.rmax;
    return dispatch_fwd_cluster(p, cc, a, st, nullptr);
}

// Clusters of this configuration that are co-resident on the device = shots in flight per "wave" of the persistent
// kernel (33 four-CTA clusters, 24 six-CTA clusters on a 148-SM B200: whole clusters per GPC).  Falls back to
// SMs / C when the occupancy query is not available (no device).
int fwd_cluster_wave(const Plan &p, const ClusterConfig &cc)
{
    const int key = ((cc.C * 64 + cc.rmax) * 1024 + cc.nthreads) * 2 + (cc.img ? 1 : 0);
    for (int i = 0; i < p.wave_n; ++i)
        if (p.wave_keys[i] == key) return p.wave_vals[i];
    ClusterFwdArgs a{};
    a.adj_mode = cc.img ? 2 : 1;
    int wave = 0;
    if (dispatch_fwd_cluster(p, cc, a, nullptr, &wave) != cudaSuccess || wave < 1) {
        cudaGetLastError();
        int sms = 148;
        device_attr(&sms, cudaDevAttrMultiProcessorCount, p.device);
        wave = sms / cc.C > 0 ? sms / cc.C : 1;
    }
    const int slot = p.wave_n < 8 ? p.wave_n++ : 7;
    p.wave_keys[slot] = key; p.wave_vals[slot] = wave;
    return wave;
}

}  // namespace rdfwi
