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
