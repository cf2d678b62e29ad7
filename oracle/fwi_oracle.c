/*
 * fwi_oracle.c -- CPU oracle (TEST INFRASTRUCTURE ONLY, see fwi_oracle.h).
 * Build: oracle/build_oracle.py (gcc -O3 -fopenmp -ffp-contract=off, no fast-math).
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>
#include "fwi_oracle.h"

#define REAL float
#define SUFFIX _f32
#include "fwi_oracle_impl.h"
#undef REAL
#undef SUFFIX

#define REAL double
#define SUFFIX _f64
#include "fwi_oracle_impl.h"
#undef REAL
#undef SUFFIX

int fwi_oracle_coeffs_f32(const fwi_oracle_geom *g, const float *v, float *planes5, float *velmin, int *argmin)
{
    const size_t cells = (size_t)g->nzp * g->nxp;
    model_coeffs_f32(g, v, NULL, planes5, planes5 + cells, planes5 + 2 * cells, planes5 + 3 * cells,
                     planes5 + 4 * cells, NULL, velmin, argmin);
    return 0;
}

int fwi_oracle_threads(void) { return omp_get_max_threads(); }
void fwi_oracle_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
