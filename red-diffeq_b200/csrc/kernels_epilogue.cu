// kernels_epilogue.cu -- from the imaging sums (Ga, Gk, Gb) to d L / d v_phys.
//
// The reference gets this from autograd (no source lines of its own); the chain being restated is
//   alpha   = (v*dt/dx)**2        solvers/pde.py:63   -> d alpha / d v = 2 (v dt/dx) dt/dx
//   beta_dt = (v*dt)**2           solvers/pde.py:71   -> 2 (v dt) dt           (source cells only, :81)
//   kappa   = get_Abc(v)*dt       solvers/pde.py:41-46, :65 -> through velmin: MinBackward adds
//                                 sum_cells Gk * d kappa/d velmin to the arg-min pixel
//   v_pad   = replicate pad       solvers/pde.py:91   -> ReplicationPad2dBackward folds the halo into the edges
// Deterministic: no atomics, fixed summation order.
#include "rdfwi_common.cuh"

namespace rdfwi {
namespace {

__device__ __forceinline__ int sponge_index(int i, int n, int nbc)
{
    return i < nbc ? nbc - 1 - i : (i >= n - nbc ? i - (n - nbc) : -1);
}

// tmp[b][iz][x] = sum over the padded rows that replicate model row iz
__global__ void __launch_bounds__(kThreads) k_fold_rows(const float *__restrict__ Ga, Grid g, int planes,
                                                        float *__restrict__ tmp)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= g.nz * g.nxp) return;
    const int iz = i / g.nxp, x = i - iz * g.nxp;
    const int zlo = iz == 0 ? 0 : iz + g.nbc;
    const int zhi = iz == g.nz - 1 ? g.nzp - 1 : iz + g.nbc;
    float acc = 0.0f;
    for (int s = 0; s < planes; ++s) {  // fixed order over the shots of the model: deterministic
        const float *src = Ga + ((size_t)b * planes + s) * g.level + x;
        for (int z = zlo; z <= zhi; ++z) acc += src[(size_t)z * g.pitch];
    }
    tmp[(size_t)b * g.nz * g.nxp + i] = acc;
}

// grad[b][iz][ix] = d alpha/d v * sum over the padded columns that replicate model column ix
__global__ void __launch_bounds__(kThreads) k_fold_cols(const float *__restrict__ tmp, const float *__restrict__ v, Grid g,
                                                        float dt, float dx, float *__restrict__ grad)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * kThreads + threadIdx.x;
    if (i >= g.nz * g.nx) return;
    const int iz = i / g.nx, ix = i - iz * g.nx;
    const int xlo = ix == 0 ? 0 : ix + g.nbc;
    const int xhi = ix == g.nx - 1 ? g.nxp - 1 : ix + g.nbc;
    const float *src = tmp + ((size_t)b * g.nz + iz) * g.nxp;
    float acc = 0.0f;
    for (int x = xlo; x <= xhi; ++x) acc += src[x];
    const float u = v[(size_t)b * g.nz * g.nx + i] * dt;
    grad[(size_t)b * g.nz * g.nx + i] = acc * ((2.0f * (u / dx)) * (dt / dx));
}

// partial sums of Gk * d(kappa dt)/d(velmin) over the sponge cells of one model (double accumulation)
__global__ void __launch_bounds__(kThreads) k_velmin_partial(const float *__restrict__ Gk, const float *__restrict__ dkap,
                                                             Grid g, int planes, double *__restrict__ part)
{
    const int b = blockIdx.y;
    const float *src = Gk + (size_t)b * planes * g.level;
    double acc = 0.0;
    const int n = g.nzp * g.nxp;
    for (int i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) {
        const int z = i / g.nxp, x = i - z * g.nxp;
        const int kx = sponge_index(x, g.nxp, g.nbc);
        const int kz = sponge_index(z, g.nzp, g.nbc);
        const int k = kx >= 0 ? kx : kz;
        if (k >= 0) {
            float gk = 0.0f;
            for (int s = 0; s < planes; ++s) gk += src[(size_t)s * g.level + (size_t)z * g.pitch + x];
            acc += (double)(gk * dkap[k]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double sm[kThreads / 32];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kThreads / 32; ++w) acc += sm[w];
        part[(size_t)b * gridDim.x + blockIdx.x] = acc;
    }
}

// one thread per model: source-cell (beta_dt) terms and the velmin term, added in a fixed order
__global__ void k_finish(const double *__restrict__ part, int nparts, const float *__restrict__ Gb,
                         const float *__restrict__ v, const int *__restrict__ isx, const int *__restrict__ argmin, Grid g,
                         float dt, int B, float *__restrict__ grad)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float *gb = grad + (size_t)b * g.nz * g.nx;
    const float *vb = v + (size_t)b * g.nz * g.nx;
    int iz = g.isz - g.nbc; iz = iz < 0 ? 0 : (iz >= g.nz ? g.nz - 1 : iz);
    for (int s = 0; s < g.ns; ++s) {
        int ix = isx[s] - g.nbc; ix = ix < 0 ? 0 : (ix >= g.nx ? g.nx - 1 : ix);
        const float u = vb[iz * g.nx + ix] * dt;
        gb[iz * g.nx + ix] += Gb[(size_t)b * g.ns + s] * ((2.0f * u) * dt);
    }
    double acc = 0.0;
    for (int k = 0; k < nparts; ++k) acc += part[(size_t)b * nparts + k];
    gb[argmin[b]] += (float)acc;
}

}  // namespace

cudaError_t launch_gradient_epilogue(const Plan &p, const float *v, int B, const float *Ga, const float *Gk,
                                     const float *Gb, int planes, const int *argmin, float *fold_tmp,
                                     double *vel_part, float *grad_v, cudaStream_t st)
{
    const Grid &g = p.g;
    k_fold_rows<<<dim3((g.nz * g.nxp + kThreads - 1) / kThreads, B), kThreads, 0, st>>>(Ga, g, planes, fold_tmp);
    count_launch();
    k_fold_cols<<<dim3((g.nz * g.nx + kThreads - 1) / kThreads, B), kThreads, 0, st>>>(fold_tmp, v, g, p.dt_f, p.dx_f, grad_v);
    count_launch();
    k_velmin_partial<<<dim3(kMinBlocks, B), kThreads, 0, st>>>(Gk, p.d_dkap, g, planes, vel_part);
    count_launch();
    k_finish<<<(B + 63) / 64, 64, 0, st>>>(vel_part, kMinBlocks, Gb, v, p.d_isx, argmin, g, p.dt_f, B, grad_v);
    count_launch();
    return cudaGetLastError();
}

}  // namespace rdfwi
