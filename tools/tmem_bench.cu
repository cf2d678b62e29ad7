// tmem_bench.cu -- can tensor memory serve as a per-thread scratchpad (imaging accumulators of the fused adjoint)?
//
// One 512-thread CTA per SM (200 KB of dynamic shared memory forces that, like k_fwd_cluster).  Every thread owns
// TMEM lane 32*(warp%4)+lane and the 128 columns [128*(warp/4), +128).  Measured per mode, in SM cycles per
// warp-level instruction and bytes per cycle per SM:
//   0  tcgen05.ld.32x32b.x4 only          1  tcgen05.st.32x32b.x4 only
//   2  ld x4 -> wait -> fma -> st x4 (read-modify-write of 13 "rows" per pass, what the adjoint sweep would do)
//   3  mode 2 with two arrays (Ga, Gk) per row      4  LDS.128 + STS.128 stream alone (shared-memory pipe reference)
//   5  mode 3 and mode 4 interleaved (do the two pipes add up or collide?)
// Also checks that what a thread wrote is what it reads back after all other warps wrote theirs.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_bench tools/tmem_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ void tm_ld4(uint32_t taddr, float (&v)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]) : "r"(taddr));
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const float (&v)[4])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int R = 13;

__global__ void __launch_bounds__(512, 1) k_tmem(int mode, int iters, long long *cycles, float *out, int *errors)
{
    extern __shared__ __align__(16) float smem[];
    __shared__ uint32_t tbase_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tbase = tbase_s;
    // lane field = bits 31:16, column = bits 15:0
    const uint32_t mine = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
    for (int i = tid; i < 48 * 1024; i += 512) smem[i] = (float)i;

    // correctness: write a thread-unique pattern to all 128 columns, sync, read back
    for (int c = 0; c < 128; c += 4) {
        const float v[4] = {(float)(tid * 1000 + c), (float)(tid * 1000 + c + 1), (float)(tid * 1000 + c + 2), (float)(tid * 1000 + c + 3)};
        tm_st4(mine + c, v);
    }
    tm_wait_st();
    __syncthreads();
    int bad = 0;
    for (int c = 0; c < 128; c += 4) {
        float v[4];
        tm_ld4(mine + c, v);
        tm_wait_ld();
        for (int j = 0; j < 4; ++j) bad += v[j] != (float)(tid * 1000 + c + j);
    }
    if (bad) atomicAdd(errors, bad);
    // zero the accumulators
    for (int c = 0; c < 128; c += 4) {
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        tm_st4(mine + c, z);
    }
    tm_wait_st();
    __syncthreads();

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    float4 *s4 = reinterpret_cast<float4 *>(smem) + tid;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (mode == 0) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float v[4];
                tm_ld4(mine + 4 * r, v);
                tm_wait_ld();
                acc[0] += v[0]; acc[1] += v[1]; acc[2] += v[2]; acc[3] += v[3];
            }
        } else if (mode == 1) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float v[4] = {acc[0] + r, acc[1], acc[2], acc[3] + it};
                tm_st4(mine + 4 * r, v);
            }
            tm_wait_st();
        } else if (mode == 2 || mode == 3 || mode == 5) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float a[4], b[4];
                tm_ld4(mine + 4 * r, a);
                if (mode != 2) tm_ld4(mine + 52 + 4 * r, b);
                float4 w = make_float4(1.f, 1.f, 1.f, 1.f);
                if (mode == 5) w = s4[r * 512];
                tm_wait_ld();
                a[0] += w.x; a[1] += w.y * 2.f; a[2] += 3.f * w.z; a[3] += 4.f * w.w;
                tm_st4(mine + 4 * r, a);
                if (mode != 2) {
                    b[0] -= 1.f; b[1] -= 2.f; b[2] -= 3.f; b[3] -= 4.f;
                    tm_st4(mine + 52 + 4 * r, b);
                }
                if (mode == 5) s4[r * 512] = make_float4(w.x, w.y, w.z, w.w + 0.f);
            }
            tm_wait_st();
        } else if (mode == 4) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                float4 w = s4[r * 512];
                w.x += 1.f;
                s4[r * 512] = w;
            }
        }
    }
    const long long t1 = clock64();
    // read back the accumulators of modes 2/3/5: a[j] must be iters*(j+1) (when w == 1, or smem-driven in mode 5)
    if (mode == 2 || mode == 3) {
        int bad2 = 0;
        for (int r = 0; r < R; ++r) {
            float a[4];
            tm_ld4(mine + 4 * r, a);
            tm_wait_ld();
            for (int j = 0; j < 4; ++j) bad2 += a[j] != (float)(iters * (j + 1));
            if (mode == 3) {
                tm_ld4(mine + 52 + 4 * r, a);
                tm_wait_ld();
                for (int j = 0; j < 4; ++j) bad2 += a[j] != -(float)(iters * (j + 1));
            }
        }
        if (bad2) atomicAdd(errors, bad2);
    }
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 512 + tid] = acc[0] + acc[1] + acc[2] + acc[3] + s4[0].x;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase) : "memory");
}

int main()
{
    const int nblk = 148, iters = 2000;
    long long *d_cyc; float *d_out; int *d_err;
    CK(cudaMalloc(&d_cyc, nblk * sizeof(long long)));
    CK(cudaMalloc(&d_out, nblk * 512 * sizeof(float)));
    CK(cudaMalloc(&d_err, sizeof(int)));
    CK(cudaFuncSetAttribute(k_tmem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    const char *names[6] = {"ld.x4 only", "st.x4 only", "ld-fma-st one array", "ld-fma-st two arrays", "LDS.128+STS.128 only", "two arrays + LDS/STS"};
    // instructions per (thread, row) and bytes moved per (thread, row) for the rate columns
    const int tm_bytes[6] = {16, 16, 32, 64, 0, 64}, sm_bytes[6] = {0, 0, 0, 0, 32, 32};
    for (int mode = 0; mode < 6; ++mode) {
        CK(cudaMemset(d_err, 0, sizeof(int)));
        k_tmem<<<nblk, 512, 200 * 1024>>>(mode, iters, d_cyc, d_out, d_err);
        CK(cudaDeviceSynchronize());
        long long cyc[148]; int err;
        CK(cudaMemcpy(cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(&err, d_err, sizeof(int), cudaMemcpyDeviceToHost));
        long long mx = 0; for (int i = 0; i < nblk; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
        const double per_warp_row_sm = (double)mx / ((double)iters * R * 16);  // cycles per warp-row, SM-wide (16 warps)
        printf("mode %d  %-24s cycles/warp-row/SM %7.2f   TMEM B/clk/SM %7.1f   SMEM B/clk/SM %7.1f   errors %d\n", mode, names[mode],
               per_warp_row_sm, tm_bytes[mode] * 32 / per_warp_row_sm, sm_bytes[mode] * 32 / per_warp_row_sm, err);
    }
    return 0;
}
