/*
 * c_abi_example.c -- the C ABI of include/rdfwi.h used from plain C (no Python, no torch): one forward + adjoint of a small
 * survey with caller-owned device buffers.  Build and run on a B200 box:
 *
 *   gcc -std=c99 -Iinclude examples/c_abi_example.c -o /tmp/rdfwi_example \
 *       -Lred-diffeq_b200 -lrdfwi -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/red-diffeq_b200 -lm
 *
 * tests/test_host_logic.py compiles it (-Wall -Werror) and links it against librdfwi.so on the CPU-only build box, which
 * checks that the header is valid C and that every entry point used here resolves.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rdfwi.h"

/* the three CUDA runtime calls this example needs, declared here so that it compiles without the CUDA headers */
extern int cudaMalloc(void **ptr, size_t bytes);
extern int cudaMemcpy(void *dst, const void *src, size_t bytes, int kind);
extern int cudaDeviceSynchronize(void);
enum { H2D = 1, D2H = 2 };

#define CHECK(call)                                                                   \
    do {                                                                              \
        int rc_ = (call);                                                             \
        if (rc_ != 0) {                                                               \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, rdfwi_last_error()); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

int main(void)
{
    enum { NZ = 20, NX = 24, NBC = 12, NS = 3, NT = 150, B = 2 };
    const double dx = 10.0, dt = 0.001, f = 25.0;
    int32_t isx[NS], igx[NX];
    double wavelet[NT];
    /* what FWIForward.__init__ / ricker / adj_sr do on the host (reference solvers/pde.py:16-36, :54-59) */
    for (int s = 0; s < NS; ++s) isx[s] = (int32_t)nearbyint((double)s * (NX - 1) / (NS - 1)) + NBC;
    for (int r = 0; r < NX; ++r) igx[r] = r + NBC;
    {
        const int nw = 2 * (int)floor(2.2 / (f * dt) / 2.0) + 1, nc = nw / 2;
        for (int t = 0; t < NT; ++t) {
            const double k = t + 1, a = (nc - k + 1) * f * dt * 3.14159265358979323846;
            wavelet[t] = t < nw ? (1.0 - 2.0 * a * a) * exp(-a * a) : 0.0;
        }
    }
    rdfwi_survey sv = {NZ, NX, NBC, NS, NX, NT, 1, 1 + NBC, 1 + NBC, dx, dt, isx, igx, wavelet};
    rdfwi_plan plan = NULL;
    CHECK(rdfwi_plan_create(&sv, &plan));

    const size_t nv = (size_t)B * NZ * NX, nseis = (size_t)B * NS * NT * NX;
    float *v_host = (float *)malloc(nv * sizeof(float)), *g_host = (float *)malloc(nv * sizeof(float));
    for (size_t i = 0; i < nv; ++i) v_host[i] = 1500.0f + 3000.0f * (float)((i / NX) % NZ) / NZ;   /* velocity grows with depth */
    const size_t ws_bytes = rdfwi_workspace_bytes(plan, B), hist_bytes = rdfwi_history_bytes(plan, B, 0);
    void *v = NULL, *seis = NULL, *cot = NULL, *grad = NULL, *ws = NULL, *hist = NULL;
    if (cudaMalloc(&v, nv * 4) || cudaMalloc(&grad, nv * 4) || cudaMalloc(&seis, nseis * 4) || cudaMalloc(&cot, nseis * 4) ||
        cudaMalloc(&ws, ws_bytes) || cudaMalloc(&hist, hist_bytes ? hist_bytes : 16)) {
        fprintf(stderr, "cudaMalloc failed\n");
        return 1;
    }
    cudaMemcpy(v, v_host, nv * 4, H2D);
    CHECK(rdfwi_forward(plan, (const float *)v, B, (float *)seis, ws, ws_bytes, hist, hist_bytes, 0, NULL));
    /* cotangent = the seismograms themselves: gradient of 0.5 * ||d||^2 */
    cudaMemcpy(cot, seis, nseis * 4, 3 /* device to device */);
    CHECK(rdfwi_backward(plan, (const float *)v, B, (const float *)cot, (float *)grad, ws, ws_bytes, hist, hist_bytes, 0, NULL));
    cudaDeviceSynchronize();
    cudaMemcpy(g_host, grad, nv * 4, D2H);
    double norm = 0.0;
    for (size_t i = 0; i < nv; ++i) norm += (double)g_host[i] * g_host[i];
    printf("rdfwi %d: |d(0.5*||d||^2)/dv| = %.6e, kernel launches of the adjoint pass: %lld\n", rdfwi_version(), sqrt(norm),
           (long long)rdfwi_last_launch_count());
    CHECK(rdfwi_plan_destroy(plan));
    free(v_host);
    free(g_host);
    return 0;
}
