/*
 * dgr_oracle_impl.h — body of the CPU oracle, included twice by dgr_oracle.c
 * (REAL=float, SUF=f32  and  REAL=double, SUF=f64).
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under dreamgaussian_b200/ may include, link or call this.
 *
 * It restates, in plain C, the algorithm of the rasterizer op that the reference calls at
 * /root/reference/gs_renderer.py:745-809.  That op (pip package diff_gaussian_rasterization, ashawkey fork)
 * is NOT vendored in the reference and has no pinned commit (readme.md:30-32) — so this is a
 * restatement of its published algorithm, anchored on the reference's in-tree maths:
 *   - quaternion -> rotation, cov3D = (R S)(R S)^T, packing xx,xy,xz,yy,yz,zz: gs_renderer.py:50-59,85-117,128-132
 *   - SH basis and signs, +0.5 offset, clamp >= 0:                        sh_utils.py:57-112, gs_renderer.py:782-793
 *   - view / projection conventions (row-vector, P[3,2]=1):                gs_renderer.py:629-671
 *   - output shapes CHW / [1,H,W], int32 radii:                            gs_renderer.py:800-822, main.py:203-208
 * PARITY UNPINNED: the reference ships no tests or golden vectors for this path.
 */

#ifndef FLAG_G
/* per-Gaussian ambiguity bits are OR-ed from several OpenMP threads */
#define FLAG_G(g, bit) __atomic_fetch_or(&c->ambig_g[(g)], (unsigned char)(bit), __ATOMIC_RELAXED)
/* a1 * a2 * T above which a depth tie between two consecutive contributors is treated as able to move other gradients */
#define DGR_ORACLE_TIE_MATERIAL 1e-4
#endif
#define CAT_(a, b) a##_##b
#define CAT(a, b) CAT_(a, b)
#define FN(name) CAT(name, SUF)

typedef struct {
    /* problem */
    int P, M, D, H, W, gx, gy;
    REAL tanfovx, tanfovy, mod, fx, fy;
    REAL bg[3], V[16], PM[16], campos[3];
    int has_sh, has_cov;
    /* inputs kept for backward (copied) */
    REAL *means, *shs, *colors, *opac, *scales, *rots, *cov3d_in;
    /* per-Gaussian state */
    REAL *px, *py, *depth, *conic /*3*/, *cov2d /*3, incl. low-pass*/, *rgb /*3*/, *cov3d /*6*/;
    int *radii; unsigned char *clamped /*3*/; int *rect /*4: minx miny maxx maxy*/;
    unsigned char *ambig_g;
    /* binning */
    size_t N; unsigned *point_list; unsigned *ranges /*2 per tile*/;
    /* per-pixel */
    unsigned *n_contrib; REAL *final_T; unsigned char *ambig_px; float *tie_slack;
    double eps;
} FN(Ctx);

static inline REAL FN(rmax)(REAL a, REAL b) { return a > b ? a : b; }
static inline REAL FN(rmin)(REAL a, REAL b) { return a < b ? a : b; }

/* sh_utils.py:57-100 (deg 0..3); sh laid out [M][3] per Gaussian (gs_renderer.py:209-212) */
static void FN(sh_basis)(int D, REAL x, REAL y, REAL z, REAL *b /*16*/) {
    b[0] = (REAL)DGR_SH_C0;
    if (D > 0) {
        b[1] = -(REAL)DGR_SH_C1 * y; b[2] = (REAL)DGR_SH_C1 * z; b[3] = -(REAL)DGR_SH_C1 * x;
        if (D > 1) {
            REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            b[4] = (REAL)DGR_SH_C2_0 * xy; b[5] = (REAL)DGR_SH_C2_1 * yz;
            b[6] = (REAL)DGR_SH_C2_2 * ((REAL)2 * zz - xx - yy);
            b[7] = (REAL)DGR_SH_C2_3 * xz; b[8] = (REAL)DGR_SH_C2_4 * (xx - yy);
            if (D > 2) {
                b[9]  = (REAL)DGR_SH_C3_0 * y * ((REAL)3 * xx - yy);
                b[10] = (REAL)DGR_SH_C3_1 * xy * z;
                b[11] = (REAL)DGR_SH_C3_2 * y * ((REAL)4 * zz - xx - yy);
                b[12] = (REAL)DGR_SH_C3_3 * z * ((REAL)2 * zz - (REAL)3 * xx - (REAL)3 * yy);
                b[13] = (REAL)DGR_SH_C3_4 * x * ((REAL)4 * zz - xx - yy);
                b[14] = (REAL)DGR_SH_C3_5 * z * (xx - yy);
                b[15] = (REAL)DGR_SH_C3_6 * x * (xx - (REAL)3 * yy);
            }
        }
    }
}

/* d basis / d(x,y,z) for the backward */
static void FN(sh_basis_grad)(int D, REAL x, REAL y, REAL z, REAL *dbx, REAL *dby, REAL *dbz) {
    for (int i = 0; i < 16; i++) { dbx[i] = dby[i] = dbz[i] = 0; }
    if (D > 0) {
        dby[1] = -(REAL)DGR_SH_C1; dbz[2] = (REAL)DGR_SH_C1; dbx[3] = -(REAL)DGR_SH_C1;
        if (D > 1) {
            REAL xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
            dbx[4] = (REAL)DGR_SH_C2_0 * y; dby[4] = (REAL)DGR_SH_C2_0 * x;
            dby[5] = (REAL)DGR_SH_C2_1 * z; dbz[5] = (REAL)DGR_SH_C2_1 * y;
            dbx[6] = (REAL)DGR_SH_C2_2 * (REAL)-2 * x; dby[6] = (REAL)DGR_SH_C2_2 * (REAL)-2 * y;
            dbz[6] = (REAL)DGR_SH_C2_2 * (REAL)4 * z;
            dbx[7] = (REAL)DGR_SH_C2_3 * z; dbz[7] = (REAL)DGR_SH_C2_3 * x;
            dbx[8] = (REAL)DGR_SH_C2_4 * (REAL)2 * x; dby[8] = (REAL)DGR_SH_C2_4 * (REAL)-2 * y;
            if (D > 2) {
                dbx[9] = (REAL)DGR_SH_C3_0 * (REAL)6 * xy; dby[9] = (REAL)DGR_SH_C3_0 * ((REAL)3 * xx - (REAL)3 * yy);
                dbx[10] = (REAL)DGR_SH_C3_1 * yz; dby[10] = (REAL)DGR_SH_C3_1 * xz; dbz[10] = (REAL)DGR_SH_C3_1 * xy;
                dbx[11] = (REAL)DGR_SH_C3_2 * (REAL)-2 * xy;
                dby[11] = (REAL)DGR_SH_C3_2 * ((REAL)4 * zz - xx - (REAL)3 * yy);
                dbz[11] = (REAL)DGR_SH_C3_2 * (REAL)8 * yz;
                dbx[12] = (REAL)DGR_SH_C3_3 * (REAL)-6 * xz; dby[12] = (REAL)DGR_SH_C3_3 * (REAL)-6 * yz;
                dbz[12] = (REAL)DGR_SH_C3_3 * ((REAL)6 * zz - (REAL)3 * xx - (REAL)3 * yy);
                dbx[13] = (REAL)DGR_SH_C3_4 * ((REAL)4 * zz - (REAL)3 * xx - yy);
                dby[13] = (REAL)DGR_SH_C3_4 * (REAL)-2 * xy;
                dbz[13] = (REAL)DGR_SH_C3_4 * (REAL)8 * xz;
                dbx[14] = (REAL)DGR_SH_C3_5 * (REAL)2 * xz; dby[14] = (REAL)DGR_SH_C3_5 * (REAL)-2 * yz;
                dbz[14] = (REAL)DGR_SH_C3_5 * (xx - yy);
                dbx[15] = (REAL)DGR_SH_C3_6 * ((REAL)3 * xx - (REAL)3 * yy);
                dby[15] = (REAL)DGR_SH_C3_6 * (REAL)-6 * xy;
            }
        }
    }
}

/* gs_renderer.py:85-106 WITHOUT the normalisation (the op consumes the quaternion as given) */
static void FN(quat_to_R)(const REAL *q, REAL R[9]) {
    REAL r = q[0], x = q[1], y = q[2], z = q[3];
    R[0] = (REAL)1 - (REAL)2 * (y * y + z * z); R[1] = (REAL)2 * (x * y - r * z); R[2] = (REAL)2 * (x * z + r * y);
    R[3] = (REAL)2 * (x * y + r * z); R[4] = (REAL)1 - (REAL)2 * (x * x + z * z); R[5] = (REAL)2 * (y * z - r * x);
    R[6] = (REAL)2 * (x * z - r * y); R[7] = (REAL)2 * (y * z + r * x); R[8] = (REAL)1 - (REAL)2 * (x * x + y * y);
}

typedef struct { unsigned tile; REAL depth; unsigned idx; } FN(Inst);
static int FN(inst_cmp)(const void *a, const void *b) {
    const FN(Inst) *x = (const FN(Inst) *)a, *y = (const FN(Inst) *)b;
    if (x->tile != y->tile) return x->tile < y->tile ? -1 : 1;
    if (x->depth != y->depth) return x->depth < y->depth ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);      /* stable: ties keep Gaussian-index order */
}

void FN(dgr_oracle_free)(FN(Ctx) *c) {
    if (!c) return;
    free(c->means); free(c->shs); free(c->colors); free(c->opac); free(c->scales); free(c->rots); free(c->cov3d_in);
    free(c->px); free(c->py); free(c->depth); free(c->conic); free(c->cov2d); free(c->rgb); free(c->cov3d);
    free(c->radii); free(c->clamped); free(c->rect); free(c->ambig_g);
    free(c->point_list); free(c->ranges); free(c->n_contrib); free(c->final_T); free(c->ambig_px); free(c->tie_slack);
    free(c);
}

static REAL *FN(dup)(const REAL *src, size_t n) {
    if (!src) return NULL;
    REAL *d = (REAL *)malloc(n * sizeof(REAL) + 8);
    memcpy(d, src, n * sizeof(REAL));
    return d;
}

/*
 * Forward.  Outputs: out_color[3*H*W] (CHW), out_depth[H*W], out_alpha[H*W], radii[P].
 * Returns an opaque context for the backward (NULL on bad arguments).
 * eps: relative width of the "a float32 implementation may legitimately decide otherwise" band around every
 * discrete decision (alpha >= 1/255, T-stop, depth ties, ceil of the radius ...): those pixels / Gaussians
 * are reported in ambig_px / ambig_g so a parity test can treat them separately.
 */
FN(Ctx) *FN(dgr_oracle_forward)(
    int P, int M, int D, int H, int W,
    double tanfovx, double tanfovy, double scale_modifier,
    const REAL *bg, const REAL *viewmatrix, const REAL *projmatrix, const REAL *campos,
    const REAL *means3D, const REAL *shs, const REAL *colors_precomp, const REAL *opacities,
    const REAL *scales, const REAL *rotations, const REAL *cov3D_precomp,
    int prefiltered, double eps,
    REAL *out_color, REAL *out_depth, REAL *out_alpha, int *out_radii)
{
    (void)prefiltered;
    if ((shs == NULL) == (colors_precomp == NULL)) return NULL;
    if ((cov3D_precomp == NULL) == (scales == NULL || rotations == NULL)) return NULL;
    FN(Ctx) *c = (FN(Ctx) *)calloc(1, sizeof(FN(Ctx)));
    c->P = P; c->M = M; c->D = D; c->H = H; c->W = W; c->eps = eps;
    c->gx = (W + DGR_TILE - 1) / DGR_TILE; c->gy = (H + DGR_TILE - 1) / DGR_TILE;
    c->tanfovx = (REAL)tanfovx; c->tanfovy = (REAL)tanfovy; c->mod = (REAL)scale_modifier;
    c->fx = (REAL)W / ((REAL)2 * c->tanfovx); c->fy = (REAL)H / ((REAL)2 * c->tanfovy);
    memcpy(c->bg, bg, 3 * sizeof(REAL)); memcpy(c->V, viewmatrix, 16 * sizeof(REAL));
    memcpy(c->PM, projmatrix, 16 * sizeof(REAL)); memcpy(c->campos, campos, 3 * sizeof(REAL));
    c->has_sh = shs != NULL; c->has_cov = cov3D_precomp != NULL;
    c->means = FN(dup)(means3D, (size_t)P * 3); c->shs = FN(dup)(shs, (size_t)P * M * 3);
    c->colors = FN(dup)(colors_precomp, (size_t)P * 3); c->opac = FN(dup)(opacities, P);
    c->scales = FN(dup)(scales, (size_t)P * 3); c->rots = FN(dup)(rotations, (size_t)P * 4);
    c->cov3d_in = FN(dup)(cov3D_precomp, (size_t)P * 6);
    size_t Pn = P > 0 ? P : 1;
    c->px = calloc(Pn, sizeof(REAL)); c->py = calloc(Pn, sizeof(REAL)); c->depth = calloc(Pn, sizeof(REAL));
    c->conic = calloc(Pn * 3, sizeof(REAL)); c->cov2d = calloc(Pn * 3, sizeof(REAL));
    c->rgb = calloc(Pn * 3, sizeof(REAL)); c->cov3d = calloc(Pn * 6, sizeof(REAL));
    c->radii = calloc(Pn, sizeof(int)); c->clamped = calloc(Pn * 3, 1); c->rect = calloc(Pn * 4, sizeof(int));
    c->ambig_g = calloc(Pn, 1);
    const REAL *V = c->V, *PM = c->PM;
    const int gx = c->gx, gy = c->gy;
    const REAL limx = (REAL)DGR_FOV_CLAMP * c->tanfovx, limy = (REAL)DGR_FOV_CLAMP * c->tanfovy;

    /* ---------------- A2: per-Gaussian preprocess ---------------- */
    size_t *touched = calloc(Pn, sizeof(size_t));
#pragma omp parallel for schedule(static)
    for (int g = 0; g < P; g++) {
        const REAL *p = means3D + 3 * (size_t)g;
        REAL t[3];
        for (int i = 0; i < 3; i++) t[i] = p[0] * V[0 + i] + p[1] * V[4 + i] + p[2] * V[8 + i] + V[12 + i];
        if (fabs((double)t[2] - (double)DGR_NEAR_CULL_Z) < eps * 10) c->ambig_g[g] |= 4;
        if (t[2] <= (REAL)DGR_NEAR_CULL_Z) continue;
        REAL ph[4];
        for (int i = 0; i < 4; i++) ph[i] = p[0] * PM[0 + i] + p[1] * PM[4 + i] + p[2] * PM[8 + i] + PM[12 + i];
        REAL pw = (REAL)1 / (ph[3] + (REAL)DGR_W_EPS);
        REAL ndcx = ph[0] * pw, ndcy = ph[1] * pw;
        /* cov3D */
        REAL S6[6];
        if (c->has_cov) { for (int i = 0; i < 6; i++) S6[i] = cov3D_precomp[6 * (size_t)g + i]; }
        else {
            REAL R[9]; FN(quat_to_R)(rotations + 4 * (size_t)g, R);
            REAL s[3] = { c->mod * scales[3 * (size_t)g], c->mod * scales[3 * (size_t)g + 1], c->mod * scales[3 * (size_t)g + 2] };
            REAL Mx[9]; /* M = R * diag(s) */
            for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) Mx[3 * i + k] = R[3 * i + k] * s[k];
            REAL Sg[9];
            for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
                REAL a = 0; for (int k = 0; k < 3; k++) a += Mx[3 * i + k] * Mx[3 * j + k]; Sg[3 * i + j] = a; }
            S6[0] = Sg[0]; S6[1] = Sg[1]; S6[2] = Sg[2]; S6[3] = Sg[4]; S6[4] = Sg[5]; S6[5] = Sg[8];
        }
        for (int i = 0; i < 6; i++) c->cov3d[6 * (size_t)g + i] = S6[i];
        /* EWA projection */
        REAL txtz = t[0] / t[2], tytz = t[1] / t[2];
        REAL tx = FN(rmin)(limx, FN(rmax)(-limx, txtz)) * t[2];
        REAL ty = FN(rmin)(limy, FN(rmax)(-limy, tytz)) * t[2];
        REAL tz = t[2];
        REAL J00 = c->fx / tz, J02 = -(c->fx * tx) / (tz * tz), J11 = c->fy / tz, J12 = -(c->fy * ty) / (tz * tz);
        REAL T0[3], T1[3];                         /* T = J * Rwv,  Rwv[i][k] = V[k][i] */
        for (int k = 0; k < 3; k++) { T0[k] = J00 * V[4 * k + 0] + J02 * V[4 * k + 2]; T1[k] = J11 * V[4 * k + 1] + J12 * V[4 * k + 2]; }
        REAL Sg[9] = { S6[0], S6[1], S6[2], S6[1], S6[3], S6[4], S6[2], S6[4], S6[5] };
        REAL ST0[3], ST1[3];
        for (int k = 0; k < 3; k++) { ST0[k] = Sg[3 * k] * T0[0] + Sg[3 * k + 1] * T0[1] + Sg[3 * k + 2] * T0[2];
                                      ST1[k] = Sg[3 * k] * T1[0] + Sg[3 * k + 1] * T1[1] + Sg[3 * k + 2] * T1[2]; }
        REAL cxx = T0[0] * ST0[0] + T0[1] * ST0[1] + T0[2] * ST0[2] + (REAL)DGR_COV2D_LOWPASS;
        REAL cxy = T0[0] * ST1[0] + T0[1] * ST1[1] + T0[2] * ST1[2];
        REAL cyy = T1[0] * ST1[0] + T1[1] * ST1[1] + T1[2] * ST1[2] + (REAL)DGR_COV2D_LOWPASS;
        REAL det = cxx * cyy - cxy * cxy;
        if (det == (REAL)0) continue;
        REAL di = (REAL)1 / det;
        REAL mid = (REAL)0.5 * (cxx + cyy);
        REAL disc = FN(rmax)((REAL)DGR_EIG_FLOOR, mid * mid - det);
        REAL lam1 = mid + (REAL)sqrt((double)disc), lam2 = mid - (REAL)sqrt((double)disc);
        REAL rr = (REAL)DGR_RADIUS_SIGMAS * (REAL)sqrt((double)FN(rmax)(lam1, lam2));
        int radius = (int)ceil((double)rr);
        if (fabs((double)rr - floor((double)rr + 0.5)) < 1e-4 * (double)rr) c->ambig_g[g] |= 4;
        REAL mx = ((ndcx + (REAL)1) * (REAL)W - (REAL)1) * (REAL)0.5;
        REAL my = ((ndcy + (REAL)1) * (REAL)H - (REAL)1) * (REAL)0.5;
        /* tile rect (C cast = truncation toward zero, then clamp to the grid) */
        REAL e[4] = { (mx - radius) / DGR_TILE, (my - radius) / DGR_TILE,
                      (mx + radius + DGR_TILE - 1) / DGR_TILE, (my + radius + DGR_TILE - 1) / DGR_TILE };
        int r4[4];
        for (int i = 0; i < 4; i++) {
            double ev = (double)e[i];
            if (fabs(ev - floor(ev + 0.5)) < 1e-4 * (1.0 + fabs(ev))) c->ambig_g[g] |= 4;
            int lim = (i & 1) ? gy : gx;
            int v = (ev >= 2147483000.0) ? lim : (ev <= -2147483000.0 ? 0 : (int)e[i]);
            r4[i] = v < 0 ? 0 : (v > lim ? lim : v);
        }
        size_t area = (size_t)(r4[2] - r4[0]) * (size_t)(r4[3] - r4[1]);
        if (r4[2] <= r4[0] || r4[3] <= r4[1]) continue;
        /* colour */
        REAL rgb[3];
        if (c->has_sh) {
            REAL dx = p[0] - c->campos[0], dy = p[1] - c->campos[1], dz = p[2] - c->campos[2];
            REAL il = (REAL)1 / (REAL)sqrt((double)(dx * dx + dy * dy + dz * dz));
            REAL b[16]; FN(sh_basis)(D, dx * il, dy * il, dz * il, b);
            int nb = (D + 1) * (D + 1);
            const REAL *sh = shs + (size_t)g * M * 3;
            for (int ch = 0; ch < 3; ch++) {
                REAL a = 0; for (int k = 0; k < nb; k++) a += b[k] * sh[3 * k + ch];
                a += (REAL)DGR_SH_OFFSET;
                if (fabs((double)a) < eps * 10) c->ambig_g[g] |= 8;
                c->clamped[3 * (size_t)g + ch] = a < 0; rgb[ch] = a < 0 ? (REAL)0 : a;
            }
        } else { for (int ch = 0; ch < 3; ch++) rgb[ch] = colors_precomp[3 * (size_t)g + ch]; }
        c->px[g] = mx; c->py[g] = my; c->depth[g] = tz; c->radii[g] = radius;
        c->conic[3 * (size_t)g] = cyy * di; c->conic[3 * (size_t)g + 1] = -cxy * di; c->conic[3 * (size_t)g + 2] = cxx * di;
        c->cov2d[3 * (size_t)g] = cxx; c->cov2d[3 * (size_t)g + 1] = cxy; c->cov2d[3 * (size_t)g + 2] = cyy;
        for (int ch = 0; ch < 3; ch++) c->rgb[3 * (size_t)g + ch] = rgb[ch];
        for (int i = 0; i < 4; i++) c->rect[4 * (size_t)g + i] = r4[i];
        touched[g] = area;
    }
    for (int g = 0; g < P; g++) out_radii[g] = c->radii[g];

    /* ---------------- A3: binning, stable (tile, depth, index) order ---------------- */
    size_t N = 0; for (int g = 0; g < P; g++) N += touched[g];
    c->N = N;
    FN(Inst) *inst = (FN(Inst) *)malloc((N ? N : 1) * sizeof(FN(Inst)));
    { size_t o = 0;
      for (int g = 0; g < P; g++) if (touched[g]) {
          const int *r = c->rect + 4 * (size_t)g;
          for (int y = r[1]; y < r[3]; y++) for (int x = r[0]; x < r[2]; x++) {
              inst[o].tile = (unsigned)(y * gx + x); inst[o].depth = c->depth[g]; inst[o].idx = (unsigned)g; o++; } } }
    qsort(inst, N, sizeof(FN(Inst)), FN(inst_cmp));
    c->point_list = malloc((N ? N : 1) * sizeof(unsigned));
    c->ranges = calloc((size_t)gx * gy * 2, sizeof(unsigned));
    for (size_t i = 0; i < N; i++) {
        c->point_list[i] = inst[i].idx;
        if (i == 0 || inst[i].tile != inst[i - 1].tile) c->ranges[2 * inst[i].tile] = (unsigned)i;
        if (i == N - 1 || inst[i].tile != inst[i + 1].tile) c->ranges[2 * inst[i].tile + 1] = (unsigned)(i + 1);
    }
    free(inst); free(touched);

    /* ---------------- A4: per-pixel front-to-back compositing ---------------- */
    size_t HW = (size_t)H * W;
    c->n_contrib = calloc(HW ? HW : 1, sizeof(unsigned)); c->final_T = calloc(HW ? HW : 1, sizeof(REAL));
    c->ambig_px = calloc(HW ? HW : 1, 1);
    c->tie_slack = calloc(HW ? HW : 1, sizeof(float));
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        int tx0 = (tile % gx) * DGR_TILE, ty0 = (tile / gx) * DGR_TILE;
        unsigned s = c->ranges[2 * tile], e = c->ranges[2 * tile + 1];
        for (int yy = ty0; yy < ty0 + DGR_TILE && yy < H; yy++) for (int xx = tx0; xx < tx0 + DGR_TILE && xx < W; xx++) {
            REAL T = 1, C[3] = { 0, 0, 0 }, Dp = 0, Wt = 0; unsigned n = 0, last = 0;
            REAL last_depth = -1, last_a = 0, last_T = 1; unsigned last_g = 0; int have_last = 0; unsigned char amb = 0;
            double slack = 0;           /* how far resolving this pixel's depth ties the other way can move its colour */
            unsigned tie_upto = 0;      /* list position (1-based) of the first member of the deepest MATERIAL depth tie */
            unsigned last_n = 0;
            for (unsigned i = s; i < e; i++) {
                unsigned g = c->point_list[i]; n++;
                REAL dx = c->px[g] - (REAL)xx, dy = c->py[g] - (REAL)yy;
                const REAL *co = c->conic + 3 * (size_t)g;
                REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                if (power > (REAL)-1e-7 && power < (REAL)1e-7) { amb |= 8; FLAG_G(g, 1); }
                if (power > 0) continue;
                REAL og = c->opac[g] * (REAL)exp((double)power);
                REAL a = FN(rmin)((REAL)DGR_ALPHA_MAX, og);
                /* a float32 conic carries ~1e-6 relative error and the three terms of the quadratic form cancel, so the
                   error of `power` (= relative error of alpha) scales with the magnitude of those terms */
                double mag = 0.5 * (fabs((double)co[0]) * (double)dx * (double)dx + fabs((double)co[2]) * (double)dy * (double)dy)
                           + fabs((double)co[1] * (double)dx * (double)dy);
                if (fabs((double)a - (double)DGR_ALPHA_MIN) < eps * (1.0 + mag) * (double)DGR_ALPHA_MIN) { amb |= 1; FLAG_G(g, 1); }
                if (a < (REAL)DGR_ALPHA_MIN) continue;
                REAL test_T = T * ((REAL)1 - a);
                if (fabs((double)test_T - (double)DGR_T_STOP) < eps * (double)DGR_T_STOP && (double)(a * T) > 1e-5) { amb |= 2; FLAG_G(g, 1); }
                if (test_T < (REAL)DGR_T_STOP) break;
                if (have_last && fabs((double)c->depth[g] - (double)last_depth) < 1e-6 * (double)c->depth[g]) {   /* ~8 float32 ulps */
                    amb |= 4; FLAG_G(g, 2); FLAG_G(last_g, 2);
                    /* Swapping the two leaves everything behind them untouched (the product of their (1 - alpha) is the same)
                       but changes the "sum behind" that every Gaussian IN FRONT of them sees by ~ a1 a2 T (s1 - s2): material
                       when both are reasonably opaque and little has been absorbed yet. */
                    double dc = 0;
                    for (int ch = 0; ch < 3; ch++) { double d = fabs((double)c->rgb[3 * (size_t)g + ch] - (double)c->rgb[3 * (size_t)last_g + ch]); if (d > dc) dc = d; }
                    slack += (double)last_a * (double)a * (double)last_T * dc;
                    if ((double)last_a * (double)a * (double)last_T > DGR_ORACLE_TIE_MATERIAL) tie_upto = last_n;
                }
                have_last = 1; last_depth = c->depth[g]; last_g = g; last_a = a; last_T = T; last_n = n;
                REAL w = a * T;
                for (int ch = 0; ch < 3; ch++) C[ch] += c->rgb[3 * (size_t)g + ch] * w;
                Dp += c->depth[g] * w; Wt += w; T = test_T; last = n;
            }
            if (tie_upto > 1) {
                /* propagate: every Gaussian compositing in front of a material tied pair at this pixel may legitimately see a
                   different gradient from a float32 implementation that resolves the tie the other way (bit 16) */
                unsigned m = 0;
                for (unsigned i = s; i < e && m + 1 < tie_upto; i++) {
                    unsigned g = c->point_list[i]; m++;
                    REAL dx = c->px[g] - (REAL)xx, dy = c->py[g] - (REAL)yy;
                    const REAL *co = c->conic + 3 * (size_t)g;
                    REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0) continue;
                    REAL a = FN(rmin)((REAL)DGR_ALPHA_MAX, c->opac[g] * (REAL)exp((double)power));
                    if (a < (REAL)DGR_ALPHA_MIN) continue;
                    FLAG_G(g, 16);
                }
            }
            size_t pix = (size_t)yy * W + xx;
            for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix] = C[ch] + T * c->bg[ch];
            out_depth[pix] = Dp; out_alpha[pix] = Wt;
            c->n_contrib[pix] = last; c->final_T[pix] = T; c->ambig_px[pix] = amb; c->tie_slack[pix] = (float)slack;
        }
    }
    return c;
}

void FN(dgr_oracle_get_flags)(const FN(Ctx) *c, unsigned char *ambig_px, unsigned char *ambig_g, unsigned long long *N) {
    if (ambig_px) memcpy(ambig_px, c->ambig_px, (size_t)c->H * c->W);
    if (ambig_g) memcpy(ambig_g, c->ambig_g, (size_t)c->P);
    if (N) *N = c->N;
}

/* per pixel: sum over its depth ties of a1 a2 T max|rgb1 - rgb2| = how far the other resolution of the ties moves the colour */
void FN(dgr_oracle_get_tie_slack)(const FN(Ctx) *c, float *out) { memcpy(out, c->tie_slack, (size_t)c->H * c->W * sizeof(float)); }

/* intermediate state, for kernel-by-kernel parity checks */
void FN(dgr_oracle_get_state)(const FN(Ctx) *c, REAL *px, REAL *py, REAL *depth, REAL *conic, REAL *rgb, int *rect) {
    size_t P = c->P;
    if (px) memcpy(px, c->px, P * sizeof(REAL)); if (py) memcpy(py, c->py, P * sizeof(REAL));
    if (depth) memcpy(depth, c->depth, P * sizeof(REAL)); if (conic) memcpy(conic, c->conic, 3 * P * sizeof(REAL));
    if (rgb) memcpy(rgb, c->rgb, 3 * P * sizeof(REAL)); if (rect) memcpy(rect, c->rect, 4 * P * sizeof(int));
}

/*
 * Backward (A5 + A6).  Upstream grads: dL_dcolor[3*H*W], dL_ddepth[H*W], dL_dalpha[H*W] (any may be NULL = 0).
 * Outputs (each may be NULL): dL_dmeans3D[P*3], dL_dmeans2D[P*3] (NDC units, z = 0), dL_dshs[P*M*3],
 * dL_dcolors[P*3], dL_dopacity[P], dL_dscales[P*3], dL_drotations[P*4], dL_dcov3D[P*6].
 */
void FN(dgr_oracle_backward)(
    const FN(Ctx) *c, const REAL *gC, const REAL *gD, const REAL *gA,
    REAL *dL_dmeans3D, REAL *dL_dmeans2D, REAL *dL_dshs, REAL *dL_dcolors, REAL *dL_dopacity,
    REAL *dL_dscales, REAL *dL_drotations, REAL *dL_dcov3D, double *moments_out /* [P*10] or NULL: the 10 per-Gaussian sums */)
{
    const int P = c->P, H = c->H, W = c->W, gx = c->gx, gy = c->gy, M = c->M, D = c->D;
    const size_t HW = (size_t)H * W;
    int nthreads = 1;
#ifdef _OPENMP
    nthreads = omp_get_max_threads();
#endif
    /* per-thread accumulators of the 10 per-Gaussian intermediates, in double:
       0,1 dL/dmean_px (pixel units)  2,3,4 dL/dconic (A, B = true d/dB, C)  5 dL/dopacity  6,7,8 dL/drgb  9 dL/ddepth */
    double *acc = calloc((size_t)nthreads * (P > 0 ? P : 1) * 10, sizeof(double));
#pragma omp parallel
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
        double *A = acc + (size_t)tid * P * 10;
#pragma omp for schedule(dynamic, 1)
        for (int tile = 0; tile < gx * gy; tile++) {
            int tx0 = (tile % gx) * DGR_TILE, ty0 = (tile / gx) * DGR_TILE;
            unsigned s = c->ranges[2 * tile];
            for (int yy = ty0; yy < ty0 + DGR_TILE && yy < H; yy++) for (int xx = tx0; xx < tx0 + DGR_TILE && xx < W; xx++) {
                size_t pix = (size_t)yy * W + xx;
                unsigned last = c->n_contrib[pix];
                REAL gc[3] = { gC ? gC[pix] : 0, gC ? gC[HW + pix] : 0, gC ? gC[2 * HW + pix] : 0 };
                REAL gd = gD ? gD[pix] : 0, ga = gA ? gA[pix] : 0;
                REAL T_final = c->final_T[pix], T = T_final;
                REAL accC[3] = { 0, 0, 0 }, accD = 0, accA = 0, la = 0, lc[3] = { 0, 0, 0 }, ld = 0;
                REAL bgdot = c->bg[0] * gc[0] + c->bg[1] * gc[1] + c->bg[2] * gc[2];
                for (unsigned k = last; k-- > 0;) {
                    unsigned g = c->point_list[s + k];
                    REAL dx = c->px[g] - (REAL)xx, dy = c->py[g] - (REAL)yy;
                    const REAL *co = c->conic + 3 * (size_t)g;
                    REAL power = (REAL)-0.5 * (co[0] * dx * dx + co[2] * dy * dy) - co[1] * dx * dy;
                    if (power > 0) continue;
                    REAL G = (REAL)exp((double)power);
                    REAL a = FN(rmin)((REAL)DGR_ALPHA_MAX, c->opac[g] * G);
                    if (a < (REAL)DGR_ALPHA_MIN) continue;
                    T = T / ((REAL)1 - a);
                    REAL dL_da = 0;
                    for (int ch = 0; ch < 3; ch++) {
                        accC[ch] = la * lc[ch] + ((REAL)1 - la) * accC[ch]; lc[ch] = c->rgb[3 * (size_t)g + ch];
                        dL_da += (lc[ch] - accC[ch]) * gc[ch];
                        A[10 * (size_t)g + 6 + ch] += (double)(a * T * gc[ch]);
                    }
                    accD = la * ld + ((REAL)1 - la) * accD; ld = c->depth[g];
                    dL_da += (ld - accD) * gd;
                    A[10 * (size_t)g + 9] += (double)(a * T * gd);
                    accA = la + ((REAL)1 - la) * accA;
                    dL_da += ((REAL)1 - accA) * ga;
                    dL_da *= T;
                    dL_da += (-T_final / ((REAL)1 - a)) * bgdot;
                    la = a;
                    REAL dL_dG = c->opac[g] * dL_da;            /* ALPHA_MAX clamp is NOT masked (UNVERIFIED-EXT) */
                    REAL gdx = co[0] * dx + co[1] * dy, gdy = co[2] * dy + co[1] * dx;
                    A[10 * (size_t)g + 0] += (double)(dL_dG * -G * gdx);
                    A[10 * (size_t)g + 1] += (double)(dL_dG * -G * gdy);
                    A[10 * (size_t)g + 2] += (double)((REAL)-0.5 * G * dx * dx * dL_dG);
                    A[10 * (size_t)g + 3] += (double)(-G * dx * dy * dL_dG);
                    A[10 * (size_t)g + 4] += (double)((REAL)-0.5 * G * dy * dy * dL_dG);
                    A[10 * (size_t)g + 5] += (double)(G * dL_da);
                }
            }
        }
    }
    for (int t = 1; t < nthreads; t++) {
        double *A = acc + (size_t)t * P * 10;
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < (size_t)P * 10; i++) acc[i] += A[i];
    }

    if (moments_out) memcpy(moments_out, acc, (size_t)P * 10 * sizeof(double));
    /* ---------------- A6: per-Gaussian chain rule back to the inputs ---------------- */
    const REAL *V = c->V, *PM = c->PM;
    const REAL limx = (REAL)DGR_FOV_CLAMP * c->tanfovx, limy = (REAL)DGR_FOV_CLAMP * c->tanfovy;
    if (dL_dmeans3D) memset(dL_dmeans3D, 0, (size_t)P * 3 * sizeof(REAL));
    if (dL_dmeans2D) memset(dL_dmeans2D, 0, (size_t)P * 3 * sizeof(REAL));
    if (dL_dshs) memset(dL_dshs, 0, (size_t)P * M * 3 * sizeof(REAL));
    if (dL_dcolors) memset(dL_dcolors, 0, (size_t)P * 3 * sizeof(REAL));
    if (dL_dopacity) memset(dL_dopacity, 0, (size_t)P * sizeof(REAL));
    if (dL_dscales) memset(dL_dscales, 0, (size_t)P * 3 * sizeof(REAL));
    if (dL_drotations) memset(dL_drotations, 0, (size_t)P * 4 * sizeof(REAL));
    if (dL_dcov3D) memset(dL_dcov3D, 0, (size_t)P * 6 * sizeof(REAL));
#pragma omp parallel for schedule(static)
    for (int g = 0; g < P; g++) {
        if (c->radii[g] <= 0) continue;
        const double *A = acc + 10 * (size_t)g;
        const REAL *p = c->means + 3 * (size_t)g;
        REAL dmean[3] = { 0, 0, 0 };
        /* (1) mean_px through PM and the 1/(w+eps) divide; the reported means2D grad is in NDC units */
        REAL gndc[2] = { (REAL)A[0] * (REAL)0.5 * (REAL)W, (REAL)A[1] * (REAL)0.5 * (REAL)H };
        if (dL_dmeans2D) { dL_dmeans2D[3 * (size_t)g] = gndc[0]; dL_dmeans2D[3 * (size_t)g + 1] = gndc[1]; }
        {
            REAL ph[4];
            for (int i = 0; i < 4; i++) ph[i] = p[0] * PM[0 + i] + p[1] * PM[4 + i] + p[2] * PM[8 + i] + PM[12 + i];
            REAL mw = (REAL)1 / (ph[3] + (REAL)DGR_W_EPS);
            REAL mul1 = ph[0] * mw * mw, mul2 = ph[1] * mw * mw;
            for (int k = 0; k < 3; k++)
                dmean[k] += (PM[4 * k + 0] * mw - PM[4 * k + 3] * mul1) * gndc[0] + (PM[4 * k + 1] * mw - PM[4 * k + 3] * mul2) * gndc[1];
        }
        /* (4) depth = t.z */
        for (int k = 0; k < 3; k++) dmean[k] += V[4 * k + 2] * (REAL)A[9];
        /* (3) colour */
        if (c->has_sh) {
            REAL dxr = p[0] - c->campos[0], dyr = p[1] - c->campos[1], dzr = p[2] - c->campos[2];
            REAL len = (REAL)sqrt((double)(dxr * dxr + dyr * dyr + dzr * dzr)), il = (REAL)1 / len;
            REAL x = dxr * il, y = dyr * il, z = dzr * il;
            REAL b[16], dbx[16], dby[16], dbz[16];
            FN(sh_basis)(D, x, y, z, b); FN(sh_basis_grad)(D, x, y, z, dbx, dby, dbz);
            int nb = (D + 1) * (D + 1);
            const REAL *sh = c->shs + (size_t)g * M * 3;
            REAL ddir[3] = { 0, 0, 0 };
            for (int ch = 0; ch < 3; ch++) {
                REAL gr = c->clamped[3 * (size_t)g + ch] ? (REAL)0 : (REAL)A[6 + ch];
                for (int k = 0; k < nb; k++) {
                    if (dL_dshs) dL_dshs[((size_t)g * M + k) * 3 + ch] = b[k] * gr;
                    ddir[0] += dbx[k] * sh[3 * k + ch] * gr; ddir[1] += dby[k] * sh[3 * k + ch] * gr; ddir[2] += dbz[k] * sh[3 * k + ch] * gr;
                }
            }
            /* normalize backward: d(v/|v|) */
            REAL dot = x * ddir[0] + y * ddir[1] + z * ddir[2];
            dmean[0] += (ddir[0] - x * dot) * il; dmean[1] += (ddir[1] - y * dot) * il; dmean[2] += (ddir[2] - z * dot) * il;
        } else if (dL_dcolors) { for (int ch = 0; ch < 3; ch++) dL_dcolors[3 * (size_t)g + ch] = (REAL)A[6 + ch]; }
        if (dL_dopacity) dL_dopacity[g] = (REAL)A[5];
        /* (2) conic -> cov2D -> (Sigma, T) */
        REAL a = c->cov2d[3 * (size_t)g], b_ = c->cov2d[3 * (size_t)g + 1], cc = c->cov2d[3 * (size_t)g + 2];
        REAL det = a * cc - b_ * b_, d2 = (REAL)1 / (det * det);
        REAL gAc = (REAL)A[2], gBc = (REAL)A[3], gCc = (REAL)A[4];
        REAL ga = d2 * (-cc * cc * gAc + b_ * cc * gBc - b_ * b_ * gCc);
        REAL gc_ = d2 * (-b_ * b_ * gAc + a * b_ * gBc - a * a * gCc);
        REAL gb = d2 * ((REAL)2 * b_ * cc * gAc - (det + (REAL)2 * b_ * b_) * gBc + (REAL)2 * a * b_ * gCc);
        REAL Gm[4] = { ga, (REAL)0.5 * gb, (REAL)0.5 * gb, gc_ };
        REAL t[3];
        for (int i = 0; i < 3; i++) t[i] = p[0] * V[0 + i] + p[1] * V[4 + i] + p[2] * V[8 + i] + V[12 + i];
        REAL txtz = t[0] / t[2], tytz = t[1] / t[2];
        int clx = (txtz < -limx || txtz > limx), cly = (tytz < -limy || tytz > limy);
        REAL tx = FN(rmin)(limx, FN(rmax)(-limx, txtz)) * t[2], ty = FN(rmin)(limy, FN(rmax)(-limy, tytz)) * t[2], tz = t[2];
        REAL J00 = c->fx / tz, J02 = -(c->fx * tx) / (tz * tz), J11 = c->fy / tz, J12 = -(c->fy * ty) / (tz * tz);
        REAL Tm[6];
        for (int k = 0; k < 3; k++) { Tm[k] = J00 * V[4 * k + 0] + J02 * V[4 * k + 2]; Tm[3 + k] = J11 * V[4 * k + 1] + J12 * V[4 * k + 2]; }
        const REAL *S6 = c->cov3d + 6 * (size_t)g;
        REAL Sg[9] = { S6[0], S6[1], S6[2], S6[1], S6[3], S6[4], S6[2], S6[4], S6[5] };
        /* dL/dSigma (full symmetric matrix) = T^T Gm T */
        REAL dS[9];
        for (int k = 0; k < 3; k++) for (int l = 0; l < 3; l++)
            dS[3 * k + l] = Tm[k] * (Gm[0] * Tm[l] + Gm[1] * Tm[3 + l]) + Tm[3 + k] * (Gm[2] * Tm[l] + Gm[3] * Tm[3 + l]);
        /* dL/dT = 2 Gm T Sigma */
        REAL TS[6];
        for (int r = 0; r < 2; r++) for (int k = 0; k < 3; k++)
            TS[3 * r + k] = Tm[3 * r] * Sg[k] + Tm[3 * r + 1] * Sg[3 + k] + Tm[3 * r + 2] * Sg[6 + k];
        REAL dT[6];
        for (int k = 0; k < 3; k++) { dT[k] = (REAL)2 * (Gm[0] * TS[k] + Gm[1] * TS[3 + k]); dT[3 + k] = (REAL)2 * (Gm[2] * TS[k] + Gm[3] * TS[3 + k]); }
        /* T = J Rwv: dL/dJ[a][i] = sum_k dT[a][k] * V[k][i] */
        REAL dJ00 = 0, dJ02 = 0, dJ11 = 0, dJ12 = 0;
        for (int k = 0; k < 3; k++) { dJ00 += dT[k] * V[4 * k + 0]; dJ02 += dT[k] * V[4 * k + 2]; dJ11 += dT[3 + k] * V[4 * k + 1]; dJ12 += dT[3 + k] * V[4 * k + 2]; }
        REAL tz2 = (REAL)1 / (tz * tz), tz3 = tz2 / tz;
        REAL dtx = clx ? (REAL)0 : -c->fx * tz2 * dJ02;
        REAL dty = cly ? (REAL)0 : -c->fy * tz2 * dJ12;
        REAL dtz = -c->fx * tz2 * dJ00 - c->fy * tz2 * dJ11 + (REAL)2 * c->fx * tx * tz3 * dJ02 + (REAL)2 * c->fy * ty * tz3 * dJ12;
        for (int k = 0; k < 3; k++) dmean[k] += V[4 * k + 0] * dtx + V[4 * k + 1] * dty + V[4 * k + 2] * dtz;
        if (dL_dmeans3D) for (int k = 0; k < 3; k++) dL_dmeans3D[3 * (size_t)g + k] = dmean[k];
        if (c->has_cov) {
            if (dL_dcov3D) { REAL *o = dL_dcov3D + 6 * (size_t)g;
                o[0] = dS[0]; o[1] = (REAL)2 * dS[1]; o[2] = (REAL)2 * dS[2]; o[3] = dS[4]; o[4] = (REAL)2 * dS[5]; o[5] = dS[8]; }
        } else {
            /* Sigma = M M^T, M = R diag(mod*s): dL/dM = 2 dS M */
            REAL R[9]; const REAL *q = c->rots + 4 * (size_t)g; FN(quat_to_R)(q, R);
            REAL s[3] = { c->mod * c->scales[3 * (size_t)g], c->mod * c->scales[3 * (size_t)g + 1], c->mod * c->scales[3 * (size_t)g + 2] };
            REAL Mx[9]; for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) Mx[3 * i + k] = R[3 * i + k] * s[k];
            REAL dM[9];
            for (int i = 0; i < 3; i++) for (int k = 0; k < 3; k++) {
                REAL v = 0; for (int j = 0; j < 3; j++) v += dS[3 * i + j] * Mx[3 * j + k]; dM[3 * i + k] = (REAL)2 * v; }
            REAL dR[9];
            for (int k = 0; k < 3; k++) {
                REAL v = 0; for (int i = 0; i < 3; i++) { v += dM[3 * i + k] * R[3 * i + k]; dR[3 * i + k] = dM[3 * i + k] * s[k]; }
                if (dL_dscales) dL_dscales[3 * (size_t)g + k] = v * c->mod;   /* exact chain rule (x scale_modifier) */
            }
            if (dL_drotations) {
                REAL r = q[0], x = q[1], y = q[2], z = q[3];
                REAL *o = dL_drotations + 4 * (size_t)g;
                o[0] = (REAL)2 * (-z * dR[1] + y * dR[2] + z * dR[3] - x * dR[5] - y * dR[6] + x * dR[7]);
                o[1] = (REAL)2 * (y * dR[1] + z * dR[2] + y * dR[3] - (REAL)2 * x * dR[4] - r * dR[5] + z * dR[6] + r * dR[7] - (REAL)2 * x * dR[8]);
                o[2] = (REAL)2 * ((REAL)-2 * y * dR[0] + x * dR[1] + r * dR[2] + x * dR[3] + z * dR[5] - r * dR[6] + z * dR[7] - (REAL)2 * y * dR[8]);
                o[3] = (REAL)2 * ((REAL)-2 * z * dR[0] - r * dR[1] + x * dR[2] + r * dR[3] - (REAL)2 * z * dR[4] + y * dR[5] + x * dR[6] + y * dR[7]);
            }
        }
    }
    free(acc);
}

#undef FN
#undef CAT
#undef CAT_
