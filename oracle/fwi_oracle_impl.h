/*
 * fwi_oracle_impl.h -- body of the CPU oracle, included twice by fwi_oracle.c
 * (once with REAL=float, once with REAL=double).
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the algorithm in the
 * reference's red_diffeq/solvers/pde.py; it exists to check the CUDA path and
 * to time a CPU baseline.  Nothing in the product package may call it.
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference repository root).
 *
 * Arithmetic contract: this translation unit is compiled with
 * -ffp-contract=off and without -ffast-math, so every '*', '+', '-', '/'
 * below is one IEEE round-to-nearest operation, evaluated in the order the
 * parentheses give -- the same order in which the reference's eager tensor
 * expression is evaluated, one rounding per tensor op.
 *
 * Internal field layout ("ghost layout"): a wavefield of one shot is stored as
 * (nzp+4) x (nxp+4) REALs; cell (z, x) lives at G_IDX(z, x) and the two ghost
 * rows / columns on each side hold periodic copies (the reference's
 * torch.roll wrap-around, solvers/pde.py:79).
 */

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUFFIX)

#define G_PX(g) ((size_t)(g)->nxp + 4)
#define G_CELLS(g) (((size_t)(g)->nzp + 4) * ((size_t)(g)->nxp + 4))
#define G_IDX(px, z, x) (((size_t)((z) + 2)) * (px) + (size_t)((x) + 2))

/* Periodic ghost refresh; stands in for torch.roll's wrap (solvers/pde.py:79). */
static void FN(refresh_ghosts)(REAL *u, int nzp, int nxp)
{
    const size_t px = (size_t)nxp + 4;
    for (int z = 0; z < nzp; ++z) {
        REAL *row = u + G_IDX(px, z, 0);
        row[-2] = row[nxp - 2];
        row[-1] = row[nxp - 1];
        row[nxp] = row[0];
        row[nxp + 1] = row[1];
    }
    memcpy(u + G_IDX(px, -2, -2), u + G_IDX(px, nzp - 2, -2), px * sizeof(REAL));
    memcpy(u + G_IDX(px, -1, -2), u + G_IDX(px, nzp - 1, -2), px * sizeof(REAL));
    memcpy(u + G_IDX(px, nzp, -2), u + G_IDX(px, 0, -2), px * sizeof(REAL));
    memcpy(u + G_IDX(px, nzp + 1, -2), u + G_IDX(px, 1, -2), px * sizeof(REAL));
}

/*
 * Coefficient planes of one velocity model.
 *   replicate padding ............ solvers/pde.py:91
 *   alpha = (v*dt/dx)**2 ......... solvers/pde.py:63
 *   sponge get_Abc ............... solvers/pde.py:38-52 (columns overwrite corners, :50-51)
 *   kappa, temp1, temp2, beta_dt . solvers/pde.py:65-71
 * Planes are (nzp x nxp), no ghosts.  Returns velmin and the row-major first
 * arg-min of the *unpadded* model (same pixel torch.min's backward on the
 * padded array folds into).
 */
static void FN(model_coeffs)(const fwi_oracle_geom *g, const REAL *v /* nz*nx */,
                             REAL *vpad, REAL *alpha, REAL *kappa, REAL *t1, REAL *t2,
                             REAL *beta, REAL *damp_over_velmin, REAL *velmin_out, int *argmin_out)
{
    const int nz = g->nz, nx = g->nx, nbc = g->nbc, nzp = g->nzp, nxp = g->nxp;
    const REAL dt = (REAL)g->dt, dx = (REAL)g->dx;

    REAL velmin = v[0];
    int amin = 0;
    for (int i = 1; i < nz * nx; ++i)
        if (v[i] < velmin) { velmin = v[i]; amin = i; }
    *velmin_out = velmin;
    *argmin_out = amin;

    /* get_Abc: a=(nbc-1)*dx; kappa0 = 3.0*velmin*log(1e7)/(2.0*a); prof[k]=kappa0*(k*dx/a)**2 */
    const double a_d = (double)(nbc - 1) * g->dx;
    const REAL a = (REAL)a_d;
    const REAL kappa0 = ((((REAL)3.0) * velmin) * (REAL)log(10000000.0)) / (REAL)(2.0 * a_d);
    /* d kappa0 / d velmin, used only by the gradient (autograd of :43) */
    const REAL dkappa0 = (((REAL)3.0) * (REAL)log(10000000.0)) / (REAL)(2.0 * a_d);
    REAL *prof = (REAL *)malloc(sizeof(REAL) * (size_t)nbc);
    REAL *dprof = (REAL *)malloc(sizeof(REAL) * (size_t)nbc);
    for (int k = 0; k < nbc; ++k) {
        REAL r = (((REAL)k) * dx) / a;
        REAL r2 = r * r;
        prof[k] = kappa0 * r2;
        dprof[k] = dkappa0 * r2;
    }

    for (int z = 0; z < nzp; ++z) {
        int iz = z - nbc; iz = iz < 0 ? 0 : (iz >= nz ? nz - 1 : iz);
        for (int x = 0; x < nxp; ++x) {
            int ix = x - nbc; ix = ix < 0 ? 0 : (ix >= nx ? nx - 1 : ix);
            const size_t c = (size_t)z * nxp + x;
            const REAL vv = v[(size_t)iz * nx + ix];
            REAL d, dd;
            if (x < nbc) { d = prof[nbc - 1 - x]; dd = dprof[nbc - 1 - x]; }
            else if (x >= nxp - nbc) { d = prof[x - (nxp - nbc)]; dd = dprof[x - (nxp - nbc)]; }
            else if (z < nbc) { d = prof[nbc - 1 - z]; dd = dprof[nbc - 1 - z]; }
            else if (z >= nzp - nbc) { d = prof[z - (nzp - nbc)]; dd = dprof[z - (nzp - nbc)]; }
            else { d = (REAL)0; dd = (REAL)0; }
            const REAL u = vv * dt;
            const REAL w = u / dx;
            const REAL al = w * w;
            const REAL kp = d * dt;
            if (vpad) vpad[c] = vv;
            alpha[c] = al;
            kappa[c] = kp;
            t1[c] = (((REAL)2) + (((REAL)-5.0) * al)) - kp;
            t2[c] = ((REAL)1) - kp;
            beta[c] = u * u;
            if (damp_over_velmin) damp_over_velmin[c] = dd * dt;
        }
    }
    free(prof);
    free(dprof);
}

/*
 * One forward time level for one shot (solvers/pde.py:79):
 *   p = temp1*p1 - temp2*p0 + alpha*(c2*(rolls +-1) + c3*(rolls +-2))
 * u1 = p_{t-1}, u0 = p_{t-2}; out may alias u0 (p needs p0 only at its own cell).
 */
static void FN(forward_level)(const fwi_oracle_geom *g, const REAL *alpha, const REAL *t1,
                              const REAL *t2, const REAL *u1, const REAL *u0, REAL *out, int row_parallel)
{
    const int nzp = g->nzp, nxp = g->nxp;
    const size_t px = G_PX(g);
    const ptrdiff_t sp = (ptrdiff_t)px;
    const REAL c2 = (REAL)(4.0 / 3.0), c3 = (REAL)(-1.0 / 12.0);
#pragma omp parallel for schedule(static) if (row_parallel)
    for (int z = 0; z < nzp; ++z) {
        const REAL *r1 = u1 + G_IDX(px, z, 0);
        const REAL *r0 = u0 + G_IDX(px, z, 0);
        REAL *ro = out + G_IDX(px, z, 0);
        const REAL *al = alpha + (size_t)z * nxp;
        const REAL *a1 = t1 + (size_t)z * nxp;
        const REAL *a2 = t2 + (size_t)z * nxp;
        for (int x = 0; x < nxp; ++x) {
            const REAL s1 = ((r1[x - sp] + r1[x + sp]) + r1[x - 1]) + r1[x + 1];
            const REAL s2 = ((r1[x - 2 * sp] + r1[x + 2 * sp]) + r1[x - 2]) + r1[x + 2];
            const REAL lap = (c2 * s1) + (c3 * s2);
            ro[x] = ((a1[x] * r1[x]) - (a2[x] * r0[x])) + (al[x] * lap);
        }
    }
}

/* shots run in parallel when there are enough of them; otherwise rows of one shot do */
static int FN(shot_parallel)(int nshots)
{
    const int nth = omp_get_max_threads();
    return nshots >= nth || 2 * nshots > nth;
}

/*
 * Time loop of one shot (solvers/pde.py:75-85).  With hist != NULL every level
 * p_t is written straight into its history slot; otherwise two rotating fields.
 */
static int FN(forward_shot)(const fwi_oracle_geom *g, const REAL *coef /* 5 planes of the model */, int s,
                            REAL *seis_shot, REAL *hist_shot, int row_parallel)
{
    const int nt = g->nt, st = g->st, nrec = g->nrec, nzp = g->nzp, nxp = g->nxp;
    const size_t cells = (size_t)nzp * nxp, gc = G_CELLS(g), px = G_PX(g);
    const REAL *alpha = coef, *t1 = coef + 2 * cells, *t2 = coef + 3 * cells, *beta = coef + 4 * cells;
    REAL *base = (REAL *)calloc(gc * 2, sizeof(REAL));
    if (!base) return 1;
    REAL *ua = base, *ub = base + gc; /* ua = p_{t-1}, ub = p_{t-2} */
    const size_t src = G_IDX(px, g->isz, g->isx[s]);
    const REAL bsrc = beta[(size_t)g->isz * nxp + g->isx[s]];
    for (int t = 0; t < nt; ++t) {
        REAL *out = hist_shot ? hist_shot + (size_t)t * gc : ub;
        FN(forward_level)(g, alpha, t1, t2, ua, ub, out, row_parallel);
        out[src] = out[src] + (bsrc * (REAL)g->wavelet[t]);      /* source injection, :80-81 */
        FN(refresh_ghosts)(out, nzp, nxp);
        if (t % st == 0) {                                        /* sampled after injection, :82-83 */
            REAL *d = seis_shot + (size_t)(t / st) * nrec;
            for (int r = 0; r < nrec; ++r) d[r] = out[G_IDX(px, g->igz, g->igx[r])];
        }
        if (hist_shot) { ub = ua; ua = out; }                      /* p0 = p1 ; p1 = p  (:84-85) */
        else { REAL *tmp = ua; ua = ub; ub = tmp; }
    }
    free(base);
    return 0;
}

/*
 * Forward modelling of a batch (solvers/pde.py:61-86 FWM, :88-93 forward).
 *   v     (B, nz, nx) physical velocity (after v_denorm_func)
 *   seis  (B, ns, nt_out, nrec)
 *   hist  NULL, or (B*ns) x nt ghost-layout fields: hist[(shot*nt + t)] = p_t
 */
int FN(fwi_oracle_forward)(const fwi_oracle_geom *g, const REAL *v, REAL *seis, REAL *hist)
{
    const int B = g->B, ns = g->ns, nt = g->nt, st = g->st, nrec = g->nrec;
    const size_t cells = (size_t)g->nzp * g->nxp, gc = G_CELLS(g);
    const int nt_out = (nt + st - 1) / st;
    int status = 0;

    REAL *coef = (REAL *)malloc((size_t)B * cells * sizeof(REAL) * 5);
    if (!coef) return 1;
    for (int b = 0; b < B; ++b) {
        REAL *c = coef + (size_t)b * cells * 5;
        REAL velmin; int amin;
        FN(model_coeffs)(g, v + (size_t)b * g->nz * g->nx, NULL, c, c + cells, c + 2 * cells,
                         c + 3 * cells, c + 4 * cells, NULL, &velmin, &amin);
    }
    const int sp = FN(shot_parallel)(B * ns);
    /* shots are independent PDE solves (solvers/pde.py:75-81) */
#pragma omp parallel for schedule(dynamic, 1) if (sp)
    for (int i = 0; i < B * ns; ++i) {
        int rc = FN(forward_shot)(g, coef + (size_t)(i / ns) * cells * 5, i % ns,
                                  seis + (size_t)i * nt_out * nrec,
                                  hist ? hist + (size_t)i * nt * gc : NULL, !sp);
        if (rc) status = rc;
    }
    free(coef);
    return status;
}

/*
 * Reverse-time loop of one shot: adjoint field q and the zero-lag imaging sums.
 *   q_t = T1 q_{t+1} + S(alpha q_{t+1}) - T2 q_{t+2} (+ cotangent at receivers)
 *   Ga += q_t * (S - 5)(p_{t-1});  Gk += q_t * (p_{t-2} - p_{t-1});  Gb += q_t[src] * w[t]
 * (what autograd accumulates for alpha / kappa / beta_dt of solvers/pde.py:63-71)
 */
static int FN(adjoint_shot)(const fwi_oracle_geom *g, const REAL *alpha, const REAL *t1, const REAL *t2, int s,
                            const REAL *cot_shot, const REAL *hist_shot, REAL *Ga, REAL *Gk, REAL *Gb_out,
                            int row_parallel)
{
    const int nt = g->nt, st = g->st, nrec = g->nrec, nzp = g->nzp, nxp = g->nxp;
    const size_t gc = G_CELLS(g), px = G_PX(g);
    const ptrdiff_t sp = (ptrdiff_t)px;
    const REAL c2 = (REAL)(4.0 / 3.0), c3 = (REAL)(-1.0 / 12.0);
    REAL *qbase = (REAL *)calloc(gc * 4, sizeof(REAL));
    if (!qbase) return 1;
    REAL *q1 = qbase;           /* q_{t+1} */
    REAL *q2 = qbase + gc;      /* q_{t+2}, overwritten by q_t */
    REAL *aq = qbase + 2 * gc;  /* alpha * q_{t+1}, ghosted */
    REAL *zero = qbase + 3 * gc; /* p_{-1} = p_{-2} = 0 */
    REAL Gb = (REAL)0;
    for (int t = nt - 1; t >= 0; --t) {
        const REAL *pm1 = t >= 1 ? hist_shot + (size_t)(t - 1) * gc : zero;
        const REAL *pm2 = t >= 2 ? hist_shot + (size_t)(t - 2) * gc : zero;
#pragma omp parallel for schedule(static) if (row_parallel)
        for (int z = 0; z < nzp; ++z) {
            const REAL *al = alpha + (size_t)z * nxp;
            const REAL *r1 = q1 + G_IDX(px, z, 0);
            REAL *ra = aq + G_IDX(px, z, 0);
            for (int x = 0; x < nxp; ++x) ra[x] = al[x] * r1[x];
        }
        FN(refresh_ghosts)(aq, nzp, nxp);
#pragma omp parallel for schedule(static) if (row_parallel)
        for (int z = 0; z < nzp; ++z) {
            const REAL *ra = aq + G_IDX(px, z, 0);
            const REAL *r1 = q1 + G_IDX(px, z, 0);
            REAL *r2 = q2 + G_IDX(px, z, 0);
            const REAL *a1 = t1 + (size_t)z * nxp, *a2 = t2 + (size_t)z * nxp;
            for (int x = 0; x < nxp; ++x) {
                const REAL s1 = ((ra[x - sp] + ra[x + sp]) + ra[x - 1]) + ra[x + 1];
                const REAL s2 = ((ra[x - 2 * sp] + ra[x + 2 * sp]) + ra[x - 2]) + ra[x + 2];
                r2[x] = ((a1[x] * r1[x]) + ((c2 * s1) + (c3 * s2))) - (a2[x] * r2[x]);
            }
        }
        if (t % st == 0) { /* receiver cotangent enters at the sampled levels (adjoint of :83) */
            const REAL *gt = cot_shot + (size_t)(t / st) * nrec;
            for (int r = 0; r < nrec; ++r) q2[G_IDX(px, g->igz, g->igx[r])] += gt[r];
        }
#pragma omp parallel for schedule(static) if (row_parallel)
        for (int z = 0; z < nzp; ++z) {
            const REAL *m1 = pm1 + G_IDX(px, z, 0);
            const REAL *m2 = pm2 + G_IDX(px, z, 0);
            const REAL *rq = q2 + G_IDX(px, z, 0);
            REAL *ga = Ga + (size_t)z * nxp, *gk = Gk + (size_t)z * nxp;
            for (int x = 0; x < nxp; ++x) {
                const REAL s1 = ((m1[x - sp] + m1[x + sp]) + m1[x - 1]) + m1[x + 1];
                const REAL s2 = ((m1[x - 2 * sp] + m1[x + 2 * sp]) + m1[x - 2]) + m1[x + 2];
                const REAL lap = (((REAL)-5.0) * m1[x]) + ((c2 * s1) + (c3 * s2));
                ga[x] += rq[x] * lap;
                gk[x] += rq[x] * (m2[x] - m1[x]);
            }
        }
        Gb += q2[G_IDX(px, g->isz, g->isx[s])] * (REAL)g->wavelet[t];
        REAL *tmp = q1; q1 = q2; q2 = tmp;
    }
    *Gb_out = Gb;
    free(qbase);
    return 0;
}

/*
 * Velocity gradient of  L = sum(seis * cot)  by the discrete adjoint of the
 * recurrence above.  The reference has no source for this: it is what
 * PyTorch autograd produces from the tape of solvers/pde.py:61-93
 * (MinBackward of :41, ReplicationPad2dBackward of :91 included); the closed
 * form is SURVEY.md appendix A.2 and is pinned against the reference's
 * autograd by tests/golden/.
 *   cot    (B, ns, nt_out, nrec)
 *   seis   optional output (B, ns, nt_out, nrec)
 *   grad_v (B, nz, nx)   d L / d v_phys
 * Models are processed in groups sized so that every host thread has a shot;
 * memory = group * ns * nt ghost fields of history.
 */
int FN(fwi_oracle_gradient)(const fwi_oracle_geom *g, const REAL *v, const REAL *cot,
                            REAL *seis, REAL *grad_v)
{
    const int B = g->B, ns = g->ns, nt = g->nt, st = g->st, nrec = g->nrec;
    const int nz = g->nz, nx = g->nx, nbc = g->nbc, nzp = g->nzp, nxp = g->nxp;
    const size_t cells = (size_t)nzp * nxp, gc = G_CELLS(g);
    const int nt_out = (nt + st - 1) / st;
    const REAL dt = (REAL)g->dt, dx = (REAL)g->dx;
    const int nth = omp_get_max_threads();
    int group = (nth + ns - 1) / ns;
    if (group > B) group = B;
    {   /* cap history at FWI_ORACLE_HIST_GB (default 40) */
        const char *e = getenv("FWI_ORACLE_HIST_GB");
        const double cap = (e ? atof(e) : 40.0) * 1e9;
        const double per_model = (double)ns * nt * (double)gc * sizeof(REAL);
        while (group > 1 && per_model * group > cap) --group;
    }
    const size_t gshots = (size_t)group * ns;

    REAL *hist = (REAL *)malloc(gshots * nt * gc * sizeof(REAL));
    REAL *seis_g = (REAL *)malloc(gshots * nt_out * nrec * sizeof(REAL));
    REAL *planes = (REAL *)malloc((size_t)group * cells * sizeof(REAL) * 7);
    REAL *Ga_s = (REAL *)malloc(gshots * cells * sizeof(REAL));
    REAL *Gk_s = (REAL *)malloc(gshots * cells * sizeof(REAL));
    REAL *Gb_s = (REAL *)malloc(gshots * sizeof(REAL));
    REAL *velmin = (REAL *)malloc((size_t)group * sizeof(REAL));
    int *amin = (int *)malloc((size_t)group * sizeof(int));
    int status = 0;
    if (!hist || !seis_g || !planes || !Ga_s || !Gk_s || !Gb_s || !velmin || !amin) { status = 1; goto done; }

    for (int b0 = 0; b0 < B; b0 += group) {
        const int nb = (B - b0) < group ? (B - b0) : group;
        const int nshots = nb * ns;
        /* planes per model: [alpha kappa t1 t2 beta | vpad d(kappa)/d(velmin)] */
        for (int m = 0; m < nb; ++m) {
            REAL *c = planes + (size_t)m * cells * 7;
            FN(model_coeffs)(g, v + (size_t)(b0 + m) * nz * nx, c + 5 * cells, c, c + cells, c + 2 * cells,
                             c + 3 * cells, c + 4 * cells, c + 6 * cells, &velmin[m], &amin[m]);
        }
        memset(Ga_s, 0, (size_t)nshots * cells * sizeof(REAL));
        memset(Gk_s, 0, (size_t)nshots * cells * sizeof(REAL));
        const int sp = FN(shot_parallel)(nshots);
#pragma omp parallel for schedule(dynamic, 1) if (sp)
        for (int i = 0; i < nshots; ++i) {
            const int m = i / ns, s = i % ns;
            const REAL *c = planes + (size_t)m * cells * 7;
            REAL *h = hist + (size_t)i * nt * gc;
            int rc = FN(forward_shot)(g, c, s, seis_g + (size_t)i * nt_out * nrec, h, !sp);
            if (!rc)
                rc = FN(adjoint_shot)(g, c, c + 2 * cells, c + 3 * cells, s,
                                      cot + ((size_t)(b0 + m) * ns + s) * nt_out * nrec, h,
                                      Ga_s + (size_t)i * cells, Gk_s + (size_t)i * cells, &Gb_s[i], !sp);
            if (rc) status = rc;
        }
        if (status) goto done;
        if (seis)
            memcpy(seis + (size_t)b0 * ns * nt_out * nrec, seis_g, (size_t)nshots * nt_out * nrec * sizeof(REAL));

        /* chain through alpha, beta_dt and kappa(velmin), then fold the replicate halo (:91) */
        for (int m = 0; m < nb; ++m) {
            const REAL *c = planes + (size_t)m * cells * 7;
            const REAL *vpad = c + 5 * cells, *dov = c + 6 * cells;
            REAL *gb = grad_v + (size_t)(b0 + m) * nz * nx;
            for (int i = 0; i < nz * nx; ++i) gb[i] = (REAL)0;
            double gvelmin = 0.0;
            for (int z = 0; z < nzp; ++z) {
                int iz = z - nbc; iz = iz < 0 ? 0 : (iz >= nz ? nz - 1 : iz);
                for (int x = 0; x < nxp; ++x) {
                    int ix = x - nbc; ix = ix < 0 ? 0 : (ix >= nx ? nx - 1 : ix);
                    const size_t cc = (size_t)z * nxp + x;
                    REAL ga = (REAL)0, gk = (REAL)0;
                    for (int s = 0; s < ns; ++s) {
                        ga += Ga_s[((size_t)m * ns + s) * cells + cc];
                        gk += Gk_s[((size_t)m * ns + s) * cells + cc];
                    }
                    const REAL u = vpad[cc] * dt;
                    const REAL dalpha = (((REAL)2) * (u / dx)) * (dt / dx);
                    gb[(size_t)iz * nx + ix] += ga * dalpha;
                    gvelmin += (double)(gk * dov[cc]);
                }
            }
            for (int s = 0; s < ns; ++s) {
                const size_t cc = (size_t)g->isz * nxp + g->isx[s];
                int iz = g->isz - nbc; iz = iz < 0 ? 0 : (iz >= nz ? nz - 1 : iz);
                int ix = g->isx[s] - nbc; ix = ix < 0 ? 0 : (ix >= nx ? nx - 1 : ix);
                const REAL u = vpad[cc] * dt;
                gb[(size_t)iz * nx + ix] += Gb_s[(size_t)m * ns + s] * ((((REAL)2) * u) * dt);
            }
            gb[amin[m]] += (REAL)gvelmin;
        }
    }
done:
    free(hist); free(seis_g); free(planes); free(Ga_s); free(Gk_s); free(Gb_s); free(velmin); free(amin);
    return status;
}

#undef G_PX
#undef G_CELLS
#undef G_IDX
#undef FN
#undef CAT
#undef CAT_
