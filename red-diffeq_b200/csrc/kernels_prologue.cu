// kernels_prologue.cu -- per-model coefficient data of the FD recurrence.
//
// Replaces, bit for bit in fp32 (against the reference run on the CPU: on CUDA PyTorch evaluates tensor / python_scalar as
// a * (1/b), which differs from the __fdiv_rn below by up to 1 ulp), the prologue of FWIForward.FWM / forward in the reference:
//   replicate padding ............. solvers/pde.py:91
//   alpha = (v*dt/dx)**2 .......... solvers/pde.py:63
//   velmin + sponge profile ....... solvers/pde.py:38-52 (get_Abc), kappa = abc*dt (:65)
//   beta_dt at the source cells ... solvers/pde.py:71, :81
// Every arithmetic op is an explicit round-to-nearest intrinsic (__fmul_rn / __fdiv_rn) so that no FMA
// contraction can change the association the reference's eager tensor ops have.
#include "rdfwi_common.cuh"

namespace rdfwi {
namespace {

struct MinPair {
    float v;
    int i;
};

__device__ __forceinline__ MinPair min_pair(MinPair a, MinPair b)
{
    // torch.min's backward routes the gradient to the first row-major occurrence
    return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}

// Stage 1: kMinBlocks partial (min, first arg-min) pairs per model.
__global__ void __launch_bounds__(kThreads) k_min_partial(const float *__restrict__ v, int n, float *__restrict__ part)
{
    const int b = blockIdx.y;
    const float *vb = v + (size_t)b * n;
    MinPair m{__int_as_float(0x7f800000), 0x7fffffff};
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const float x = vb[i];
        if (x < m.v) { m.v = x; m.i = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MinPair t{__shfl_xor_sync(0xffffffffu, m.v, o), __shfl_xor_sync(0xffffffffu, m.i, o)};
        m = min_pair(m, t);
    }
    __shared__ MinPair sm[kThreads / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) m = min_pair(m, sm[w]);
        float *o = part + ((size_t)b * gridDim.x + blockIdx.x) * 2;
        o[0] = m.v;
        o[1] = __int_as_float(m.i);
    }
}

// Stage 2: one CTA per model -- velmin, arg-min, kappa table, beta at the sources.
__global__ void __launch_bounds__(128) k_model_tables(const float *__restrict__ v, const float *__restrict__ part, int nparts,
                                                      const float *__restrict__ r2, const int *__restrict__ isx, Grid g,
                                                      float dt, float log1e7, float two_a, float *__restrict__ kap,
                                                      float *__restrict__ velmin, int *__restrict__ argmin,
                                                      float *__restrict__ beta_src)
{
    const int b = blockIdx.x;
    __shared__ float s_kappa0;
    if (threadIdx.x == 0) {
        MinPair m{part[(size_t)b * nparts * 2], __float_as_int(part[(size_t)b * nparts * 2 + 1])};
        for (int k = 1; k < nparts; ++k) {
            const float *p = part + ((size_t)b * nparts + k) * 2;
            m = min_pair(m, MinPair{p[0], __float_as_int(p[1])});
        }
        velmin[b] = m.v;
        argmin[b] = m.i;
        // kappa = 3.0 * velmin * np.log(1e7) / (2.0 * a)      (solvers/pde.py:43)
        s_kappa0 = __fdiv_rn(__fmul_rn(__fmul_rn(3.0f, m.v), log1e7), two_a);
    }
    __syncthreads();
    const float kappa0 = s_kappa0;
    for (int k = threadIdx.x; k <= g.nbc; k += blockDim.x) {
        // damp1d = kappa * (k*dx/a)**2 (:46); kappa_dt = abc * dt (:65).  Entry nbc is the interior (0).
        kap[(size_t)b * (g.nbc + 1) + k] = (k < g.nbc) ? __fmul_rn(__fmul_rn(kappa0, r2[k]), dt) : 0.0f;
    }
    for (int s = threadIdx.x; s < g.ns; s += blockDim.x) {
        int iz = g.isz - g.nbc; iz = iz < 0 ? 0 : (iz >= g.nz ? g.nz - 1 : iz);
        int ix = isx[s] - g.nbc; ix = ix < 0 ? 0 : (ix >= g.nx ? g.nx - 1 : ix);
        const float u = __fmul_rn(v[((size_t)b * g.nz + iz) * g.nx + ix], dt);
        beta_src[(size_t)b * g.ns + s] = __fmul_rn(u, u);  // beta_dt = (v*dt)**2   (:71)
    }
}

// alpha on the padded, pitched grid (image columns included).
__global__ void __launch_bounds__(kThreads) k_alpha_pad(const float *__restrict__ v, Grid g, float dt, float dx,
                                                        float *__restrict__ alpha)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= g.nzp * g.pitch) return;
    const int z = i / g.pitch;
    int x = i - z * g.pitch;
    if (x >= g.nxp) x -= g.nxp;  // periodic image column
    int iz = z - g.nbc; iz = iz < 0 ? 0 : (iz >= g.nz ? g.nz - 1 : iz);
    int ix = x - g.nbc; ix = ix < 0 ? 0 : (ix >= g.nx ? g.nx - 1 : ix);
    const float u = __fmul_rn(v[((size_t)b * g.nz + iz) * g.nx + ix], dt);
    const float w = __fdiv_rn(u, dx);
    alpha[(size_t)b * g.level + i] = __fmul_rn(w, w);
}

}  // namespace

cudaError_t launch_coefficients(const Plan &p, const float *v, int B, float *alpha_pad, float *kap, float *velmin,
                                int *argmin, float *beta_src, float *minpart, cudaStream_t st)
{
    const Grid &g = p.g;
    const int n = g.nz * g.nx;
    int nparts = (n + kThreads - 1) / kThreads;
    if (nparts > kMinBlocks) nparts = kMinBlocks;
    k_min_partial<<<dim3(nparts, B), kThreads, 0, st>>>(v, n, minpart);
    count_launch();
    k_model_tables<<<B, 128, 0, st>>>(v, minpart, nparts, p.d_r2, p.d_isx, g, p.dt_f, p.log1e7_f, p.two_a_f, kap, velmin,
                                      argmin, beta_src);
    count_launch();
    const int cells = g.nzp * g.pitch;
    k_alpha_pad<<<dim3((cells + kThreads - 1) / kThreads, B), kThreads, 0, st>>>(v, g, p.dt_f, p.dx_f, alpha_pad);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rdfwi
