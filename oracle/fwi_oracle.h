/*
 * fwi_oracle.h -- CPU oracle for the 2-D acoustic FD forward solve and its
 * adjoint.  TEST INFRASTRUCTURE ONLY (see fwi_oracle_impl.h).  Only tests/,
 * __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it.
 *
 * Parity pin: the reference ships no golden vectors, so this oracle is pinned
 * by fixtures generated in the build container by importing the reference's
 * own red_diffeq/solvers/pde.py (tests/golden/make_golden.py, outputs under
 * tests/golden/).  Forward: bit-identical fp32 seismograms.  Gradient: the
 * closed-form adjoint vs the reference's autograd (tolerances in the tests).
 */
#ifndef FWI_ORACLE_H
#define FWI_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    int B;        /* velocity models in the batch                                  */
    int nz, nx;   /* unpadded model size (rows = depth, cols = lateral)            */
    int nbc;      /* sponge width, solvers/pde.py:91                               */
    int nzp, nxp; /* padded size nz+2*nbc, nx+2*nbc                                */
    int ns, nrec; /* shots per model, receivers                                    */
    int nt, st;   /* time levels, sample_temporal                                  */
    int isz, igz; /* padded source / receiver row, solvers/pde.py:54-59            */
    double dx, dt;
    const int *isx;        /* (ns)   padded source columns                         */
    const int *igx;        /* (nrec) padded receiver columns                       */
    const double *wavelet; /* (nt)   ricker(), solvers/pde.py:26-36, float64       */
} fwi_oracle_geom;

int fwi_oracle_forward_f32(const fwi_oracle_geom *g, const float *v, float *seis, float *hist);
int fwi_oracle_forward_f64(const fwi_oracle_geom *g, const double *v, double *seis, double *hist);
int fwi_oracle_gradient_f32(const fwi_oracle_geom *g, const float *v, const float *cot, float *seis, float *grad_v);
int fwi_oracle_gradient_f64(const fwi_oracle_geom *g, const double *v, const double *cot, double *seis, double *grad_v);
/* coefficient planes of model b=0 only: alpha, kappa, temp1, temp2, beta_dt (each nzp*nxp) */
int fwi_oracle_coeffs_f32(const fwi_oracle_geom *g, const float *v, float *planes5, float *velmin, int *argmin);
int fwi_oracle_threads(void);
/* OpenMP threads of the following calls (bench.py: launchers such as torchrun export OMP_NUM_THREADS=1) */
void fwi_oracle_set_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
