// kernels_misfit.cu -- data misfit of the inversion loop, evaluated in one pass over the seismograms.
//
// Replaces, for callers that opt in (FWIForward.misfit), the chain of elementwise ATen kernels behind
//   LossCalculator.observation_loss   red_diffeq/core/losses.py:27-40   |target - predicted| (* mask), per-model sum, / count
// and its autograd (sign(predicted - target) * mask * g / count): one read of the modelled and the observed
// seismograms yields the per-model sums AND the sign field the adjoint pass needs as cotangent, so the
// (B, ns, nt, n_rec) residual never makes a round trip through HBM as a torch tensor.
// Deterministic: per-thread double accumulators, shuffle / shared-memory reduction in a fixed order, no atomics.
#include "rdfwi_common.cuh"

namespace rdfwi {
namespace {

// part[b][block][0] = sum |obs - seis| * mask, part[b][block][1] = sum mask over the block's elements of model b
__global__ void __launch_bounds__(kThreads) k_misfit_partial(const float *__restrict__ seis, const float *__restrict__ obs,
                                                             const float *__restrict__ mask, const long long n,
                                                             float *sign_out, double *__restrict__ part)
{
    const int b = blockIdx.y;
    const float *s = seis + (size_t)b * n, *o = obs + (size_t)b * n;
    const float *m = mask ? mask + (size_t)b * n : nullptr;
    float *d = sign_out ? sign_out + (size_t)b * n : nullptr;
    double acc = 0.0, cnt = 0.0;
    auto one = [&](const float sv, const float ov, const float mv, float *dst) {
        const float r = sv - ov;
        acc += (double)(fabsf(ov - sv) * mv);   // L1Loss(target, predicted) * mask   (losses.py:27-32)
        cnt += (double)mv;
        if (dst) *dst = (float)((r > 0.0f) - (r < 0.0f)) * mv;  // d|o - s| / ds = sign(s - o); torch.sign(0) = 0
    };
    const bool vec = (n & 3) == 0;  // every model then starts 16-byte aligned (the buffers come from the caching allocator)
    if (vec) {
        const long long n4 = n >> 2;
        for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n4; i += (long long)gridDim.x * kThreads) {
            const float4 sv = reinterpret_cast<const float4 *>(s)[i], ov = reinterpret_cast<const float4 *>(o)[i];
            const float4 mv = m ? reinterpret_cast<const float4 *>(m)[i] : make_float4(1.f, 1.f, 1.f, 1.f);
            float4 out;
            one(sv.x, ov.x, mv.x, &out.x); one(sv.y, ov.y, mv.y, &out.y); one(sv.z, ov.z, mv.z, &out.z); one(sv.w, ov.w, mv.w, &out.w);
            if (d) reinterpret_cast<float4 *>(d)[i] = out;
        }
    } else {
        for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kThreads)
            one(s[i], o[i], m ? m[i] : 1.0f, d ? d + i : nullptr);
    }
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, k);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, k);
    }
    __shared__ double sm[2][kThreads / 32];
    if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = acc; sm[1][threadIdx.x >> 5] = cnt; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) { acc += sm[0][w]; cnt += sm[1][w]; }
        part[((size_t)b * gridDim.x + blockIdx.x) * 2] = acc;
        part[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = cnt;
    }
}

// one thread per model: the partial sums in a fixed order
__global__ void k_misfit_finish(const double *__restrict__ part, const int nparts, const int B, double *__restrict__ stats)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double acc = 0.0, cnt = 0.0;
    for (int k = 0; k < nparts; ++k) { acc += part[((size_t)b * nparts + k) * 2]; cnt += part[((size_t)b * nparts + k) * 2 + 1]; }
    stats[2 * b] = acc;
    stats[2 * b + 1] = cnt;
}

}  // namespace

size_t misfit_scratch_bytes(int B) { return (size_t)B * kMinBlocks * 2 * sizeof(double); }

cudaError_t launch_misfit_l1(const float *seis, const float *obs, const float *mask, int B, long long n, float *sign_out,
                             double *stats, double *part, cudaStream_t st)
{
    k_misfit_partial<<<dim3(kMinBlocks, B), kThreads, 0, st>>>(seis, obs, mask, n, sign_out, part);
    count_launch();
    k_misfit_finish<<<(B + 63) / 64, 64, 0, st>>>(part, kMinBlocks, B, stats);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rdfwi
