// cluster_ptx.cuh -- device helpers shared by the cluster-resident kernels: vector loads, cluster rank /
// barrier, distributed-shared-memory stores, 1-D bulk copies (TMA without a tensor map) and mbarriers.
#pragma once
#include "rdfwi_common.cuh"

namespace rdfwi {
namespace {

__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }
__device__ __forceinline__ float lane(const float4 &v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); }

__device__ __forceinline__ int sponge_index(int i, int n, int nbc)
{
    return i < nbc ? nbc - 1 - i : (i >= n - nbc ? i - (n - nbc) : -1);
}

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// store a float4 into the same shared-memory offset of another CTA of the cluster
__device__ __forceinline__ void st_cluster_v4(const float *local_ptr, uint32_t cta, float4 v)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(local_ptr)), "r"(cta));
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(remote), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ void bulk_store(float *gptr, const float *sptr, uint32_t bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gptr), "r"(smem_u32(sptr)), "r"(bytes)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }


// ---- packed fp32x2 arithmetic (Blackwell FADD2 / FMUL2 / FFMA2): two IEEE round-to-nearest fp32 results per
// issue slot.  NOTE: ptxas contracts a packed mul feeding a packed add into FFMA2 even with explicit .rn, so the
// bit-exact forward path only uses f2add (on operands that are not packed products); the adjoint uses all three.
__device__ __forceinline__ float2 f2add(const float2 a, const float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tadd.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 f2sub(const float2 a, const float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tsub.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 f2mul(const float2 a, const float2 b)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmul.rn.f32x2 rc, ra, rb;\n\tmov.b64 {%0, %1}, rc;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 f2fma(const float2 a, const float2 b, const float2 c)
{
    float2 r;
    asm("{.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;}"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return r;
}

// per-thread constants of the cluster-resident sweeps
struct SweepThread {
    int x, la, lb;            // first column, local row range [la, lb)
    int lac;                  // la clamped into the slab for lanes that own nothing
    bool edgeL, edgeR;
    bool swap;                // lanes 8-15 / 24-31: read the right neighbour pair first (conflict-free LDS.64, see fwd_sweep)
    int eL, eR;               // column offsets of the (x-2, x-1) / (x+4, x+5) pairs, periodic
    bool colsp[4];
    float mz[4];              // 0 in sponge columns (kappa follows the column profile), 1 elsewhere (row profile)
    int src_lr, rec_lr;
};

// asynchronous float4 store into another CTA's shared memory; completes 16 bytes on that CTA's mbarrier
// (local_ptr / local_bar are this CTA's addresses of the same objects: all CTAs share one smem layout)
__device__ __forceinline__ void st_async_v4(const float *local_ptr, const uint64_t *local_bar, uint32_t cta, float4 v)
{
    uint32_t raddr, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(local_ptr)), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(smem_u32(local_bar)), "r"(cta));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.f32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr),
                 "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(rbar)
                 : "memory");
}

struct HaloPush {             // per thread
    bool early;               // this thread sends its first two marching rows from inside the sweep
    int dst;                  // float offset (inside the written buffer) of the neighbour's halo row for marching row 0
    uint32_t cta;             // receiving CTA
    int bar;                  // 0 = the receiver's "top halo" barrier, 1 = its "bottom halo" barrier
};

// ---- tensor memory as a per-thread scratchpad (imaging accumulators of the resident adjoint).
// TMEM is 128 lanes x 512 columns of 32 bits per SM.  A warp can only touch the 32 lanes 32*(warp%4) .. +31, thread i of
// the warp lane 32*(warp%4)+i, so with 16 warps per CTA every thread privately owns 128 columns: [128*(warp/4), +128).
// Measured (tools/tmem_bench.cu, profiles/tmem_scratchpad_r2.txt): 270 B/clk/SM loads, 350 B/clk/SM stores -- twice the
// shared-memory pipe -- and a read-modify-write stream next to a saturating LDS/STS stream slows the latter by 10 %:
// the two are separate pipes.  Addresses are warp-uniform: bits 31:16 lane, 15:0 column.
__device__ __forceinline__ void tm_ld4(uint32_t taddr, float (&v)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, float a, float b, float c, float d)
{
    // no "memory" clobber: tensor memory is its own address space, and a clobber here would pin every shared-memory access
    // of the sweep to its row (the compiler could no longer hoist the next row's loads above this row's arithmetic)
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "f"(a), "f"(b), "f"(c), "f"(d));
}
// wait for the thread's outstanding tcgen05.ld; the loaded registers pass through the statement ("+f") so that no use of
// them can be scheduled above it -- the only ordering the compiler has to respect
__device__ __forceinline__ void tm_wait_ld(float (&a)[4], float (&b)[4])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;" : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(b[0]), "+f"(b[1]), "+f"(b[2]), "+f"(b[3]));
}
__device__ __forceinline__ void tm_wait_ld(float (&a)[4], float (&b)[4], float (&c)[4])
{
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(a[0]), "+f"(a[1]), "+f"(a[2]), "+f"(a[3]), "+f"(b[0]), "+f"(b[1]), "+f"(b[2]), "+f"(b[3]), "+f"(c[0]), "+f"(c[1]),
                   "+f"(c[2]), "+f"(c[3]));
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// whole-TMEM allocation by one warp (one CTA per SM: the slabs fill the shared memory); base address lands in *slot
__device__ __forceinline__ void tm_alloc_all(uint32_t *slot)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_free_all(uint32_t base)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base) : "memory");
}
__device__ __forceinline__ void tm_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 16-byte asynchronous copy global -> shared (L2 only), one commit group per copy; volatile statements keep their order
// among themselves (copy -> wait -> read of the slot -> next copy into it), no "memory" clobber for the reason above
__device__ __forceinline__ void cp_async16_commit(uint32_t saddr, const float *gptr)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.commit_group;" ::"r"(saddr), "l"(gptr));
}
template <int N>
__device__ __forceinline__ void cp_async_wait()
{
    asm volatile("cp.async.wait_group %0;" ::"n"(N));
}
__device__ __forceinline__ float4 lds4_volatile(uint32_t saddr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void l2_prefetch_bulk(const float *gptr, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// per-thread state of the resident imaging (k_fwd_cluster, MODE 2)
struct ImgThread {
    const float *pg;  // global: this thread's float4 of its first marching row of the forward level paired with this sweep
    uint32_t tm;      // TMEM address of the thread's columns: Ga [0, 4R), Gk [4R, 8R), alpha of the last ATM rows [8R, ..)
    float *ring;      // the thread's first 16-byte slot in shared memory (the second one is `threads` slots further)
    bool first, next; // first level of a shot (its first two rows are not in flight yet) / a further level follows
    size_t level;     // floats per level of the history
};

// ---- mbarrier + bulk load (global -> shared), used to stream the forward history into the adjoint
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_load(float *sptr, const float *gptr, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sptr)),
                 "l"(gptr), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

}  // namespace
}  // namespace rdfwi
