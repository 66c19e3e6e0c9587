/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the Oceananigans.jl
 * NonhydrostaticModel tendency arithmetic on a RectilinearGrid.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libocean_b200.so) never links or calls it.
 *
 * PARITY PINNING: the reference is Julia and cannot execute in this image (no julia binary, no
 * network, dependencies un-vendored) and its tree holds no golden vectors for this path
 * (regression truth is a network DataDep, test/data_dependencies.jl:16-50).  This restatement is
 * pinned by the reference's analytic known-answer tests re-stated in tests/test_oracle_*.py
 * (coefficients, closure flux divergences, WENO smoothness properties, Poisson residuals,
 * Taylor-Green decay, incompressibility) -- NOT by reference outputs: "parity unpinned" beyond those.
 *
 * Structure mirrors the reference: every quantity is a pointwise function of (i,j,k) exactly like
 * the reference operators, each face flux is evaluated twice through the difference operator
 * (src/Operators/difference_operators.jl:20-27), one "kernel" per tendency
 * (src/Models/NonhydrostaticModels/compute_nonhydrostatic_tendencies.jl:107-150).
 *
 * Compiled twice: -DFT=double (suffix _f64) and -DFT=float (suffix _f32).
 * Build with -ffp-contract=off: fused multiply-adds appear only where the reference has @muladd.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef FT
#define FT double
#endif
#ifdef ORACLE_F80
/* extended precision (x87 long double): same formulas, used only to measure the rounding error of the Float64 evaluation */
#define SUF(n) n##_f80
#define FMA(a, b, c) fmal((a), (b), (c))
#define FABS(a) fabsl(a)
#define SQRT(a) sqrtl(a)
#define CBRT(a) cbrtl(a)
#define FMAX(a, b) fmaxl((a), (b))
#define FMIN(a, b) fminl((a), (b))
#elif defined(ORACLE_F32)
#define SUF(n) n##_f32
#define FMA(a, b, c) fmaf((a), (b), (c))
#define FABS(a) fabsf(a)
#define SQRT(a) sqrtf(a)
#define CBRT(a) cbrtf(a)
#define FMAX(a, b) fmaxf((a), (b))
#define FMIN(a, b) fminf((a), (b))
#else
#define SUF(n) n##_f64
#define FMA(a, b, c) fma((a), (b), (c))
#define FABS(a) fabs(a)
#define SQRT(a) sqrt(a)
#define CBRT(a) cbrt(a)
#define FMAX(a, b) fmax((a), (b))
#define FMIN(a, b) fmin((a), (b))
#endif

#define MAXTR 8
#define MAXCL 4
#define MAXBUF 6

enum { PERIODIC = 0, BOUNDED = 1, FLAT = 2 };
enum { ADV_NONE = 0, ADV_CENTERED = 1, ADV_WENO = 2 };
enum { CL_SCALAR = 1, CL_SMAG = 2, CL_AMD = 3 };
enum { BUOY_NONE = 0, BUOY_TRACER = 1, BUOY_SEAWATER = 2 };

/* A field = parent array (x fastest) + offsets of logical index 1 (src/Grids/new_data.jl:11-74). */
typedef struct {
    FT *p;
    int P[3]; /* parent sizes */
    int o[3]; /* parent index (0-based) of logical index 1 */
} ofield;

typedef struct {
    int N[3], H[3], topo[3];
    const FT *df[3]; /* face spacings   Δᶠ, value at logical i = df[d][i + dfo[d]] */
    const FT *dc[3]; /* center spacings Δᶜ */
    int dfo[3], dco[3];
    /* advection */
    int adv_kind, adv_buffer;
    const FT *weno_coeff; /* [MAXBUF+1][MAXBUF][MAXBUF]   coeff_p(buffer, stencil)[j]       */
    const FT *weno_beta;  /* [MAXBUF+1][MAXBUF][21]       smoothness_coefficients           */
    const FT *weno_cstar; /* [MAXBUF+1][MAXBUF]           optimal weights C*                */
    const FT *cen_coeff;  /* [MAXBUF+1][2*MAXBUF]         centered coeffs in application order */
    FT weno_eps;
    /* closures */
    int nclosures;
    int closure_kind[MAXCL];
    FT nu[MAXCL];
    FT kappa[MAXCL][MAXTR];
    FT Pr[MAXCL][MAXTR];
    ofield nue[MAXCL];
    ofield kappae[MAXCL][MAXTR];
    /* smagorinsky / amd constants (used by closure-field kernels) */
    FT cs[MAXCL], cb[MAXCL];
    int lilly[MAXCL];
    FT Cnu[MAXCL], Ckappa[MAXCL][MAXTR];
    int amd_has_cb[MAXCL];
    /* buoyancy */
    int buoy_kind, ib, iT, iS;
    FT g, alpha, beta;
    int has_coriolis;
    FT fcor;
    /* fields */
    int ntracers;
    ofield u, v, w, c[MAXTR];
    int has_pHY;
    ofield pHY;
    ofield Gu, Gv, Gw, Gc[MAXTR];
    /* VerticallyImplicitTimeDiscretization per closure (abstract_scalar_diffusivity_closure.jl:265-312) */
    int closure_vi[MAXCL];
} oparams;

typedef const oparams *P;

static inline FT at(const ofield *f, int i, int j, int k) {
    long ii = i - 1 + f->o[0], jj = j - 1 + f->o[1], kk = k - 1 + f->o[2];
    return f->p[ii + (long)f->P[0] * (jj + (long)f->P[1] * kk)];
}
static inline FT *ref(const ofield *f, int i, int j, int k) {
    long ii = i - 1 + f->o[0], jj = j - 1 + f->o[1], kk = k - 1 + f->o[2];
    return &f->p[ii + (long)f->P[0] * (jj + (long)f->P[1] * kk)];
}

/* spacings: src/Operators/spacings_and_areas_and_volumes.jl:122-188 (Flat => 1) */
static inline FT DF(P g, int d, int i) { return g->topo[d] == FLAT ? (FT)1 : g->df[d][i + g->dfo[d]]; }
static inline FT DC(P g, int d, int i) { return g->topo[d] == FLAT ? (FT)1 : g->dc[d][i + g->dco[d]]; }
#define dxf(i) DF(g, 0, i)
#define dxc(i) DC(g, 0, i)
#define dyf(j) DF(g, 1, j)
#define dyc(j) DC(g, 1, j)
#define dzf(k) DF(g, 2, k)
#define dzc(k) DC(g, 2, k)

/* ------------------------------------------------------------------------------------------
 * Advection schemes (src/Advection/weno_reconstruction.jl:107-140, centered_reconstruction.jl:11-28)
 * ------------------------------------------------------------------------------------------ */
typedef struct scheme {
    int kind, n;
    const struct scheme *buffer_scheme;
    const struct scheme *advecting;
} scheme;

static scheme cen_chain[MAXBUF + 1];  /* Centered{n}, buffer_scheme = Centered{n-1} */
static scheme weno_chain[MAXBUF + 1]; /* WENO{n}: buffer WENO{n-1} ... WENO{2} -> Centered{1}; advecting Centered{n-1} */
static int chains_built = 0;
static void build_chains(void) {
    if (chains_built) return;
    for (int n = 1; n <= MAXBUF; n++) {
        cen_chain[n].kind = ADV_CENTERED;
        cen_chain[n].n = n;
        cen_chain[n].buffer_scheme = n > 1 ? &cen_chain[n - 1] : NULL;
        cen_chain[n].advecting = NULL;
    }
    for (int n = 2; n <= MAXBUF; n++) {
        weno_chain[n].kind = ADV_WENO;
        weno_chain[n].n = n;
        /* order = 2n-1; order <= 3 (n == 2) => buffer_scheme = Centered(order=2) */
        weno_chain[n].buffer_scheme = n > 2 ? &weno_chain[n - 1] : &cen_chain[1];
        weno_chain[n].advecting = &cen_chain[n - 1]; /* Centered(order = 2n-2) */
    }
    chains_built = 1;
}

/* "ψ" is either an array or a function of (i,j,k): */
typedef FT (*getfn)(P g, const void *arg, int i, int j, int k);
static FT get_field(P g, const void *arg, int i, int j, int k) { (void)g; return at((const ofield *)arg, i, j, k); }
/* Ax_qᶠᶜᶜ etc. (src/Operators/products_between_fields_and_grid_metrics.jl:5-14) */
static FT get_Ax_u(P g, const void *arg, int i, int j, int k) { (void)i; return (dyc(j) * dzc(k)) * at((const ofield *)arg, i, j, k); }
static FT get_Ay_v(P g, const void *arg, int i, int j, int k) { (void)j; return (dxc(i) * dzc(k)) * at((const ofield *)arg, i, j, k); }
static FT get_Az_w(P g, const void *arg, int i, int j, int k) { (void)k; return (dxc(i) * dyc(j)) * at((const ofield *)arg, i, j, k); }

static inline FT getd(P g, getfn f, const void *arg, int dir, int i, int j, int k, int s) {
    return dir == 0 ? f(g, arg, i + s, j, k) : dir == 1 ? f(g, arg, i, j + s, k) : f(g, arg, i, j, k + s);
}

/* symmetric_interpolate_*ᶠ for Centered{n} (centered_reconstruction.jl:54-63;
 * reconstruction_coefficients.jl:134-165): @muladd sum of C_m * ψ[idx + m - n - 1], m = 1..2n */
static FT centered_face(P g, int n, int dir, int i, int j, int k, getfn f, const void *arg) {
    const FT *C = g->cen_coeff + n * 2 * MAXBUF;
    FT acc = C[0] * getd(g, f, arg, dir, i, j, k, -n);
    for (int m = 2; m <= 2 * n; m++) acc = FMA(C[m - 1], getd(g, f, arg, dir, i, j, k, m - n - 1), acc);
    return acc;
}

/* Z-WENO (weno_interpolants.jl:348-354, 516-533) at face index (i,j,k) along dir */
static FT weno_face(P g, int n, int dir, int left_bias, int i, int j, int k, getfn f, const void *arg) {
    FT S[2 * MAXBUF];
    for (int m = 0; m < 2 * n; m++) S[m] = getd(g, f, arg, dir, i, j, k, m - n); /* ψ[i-n .. i+n-1] */
    FT sub[MAXBUF][MAXBUF], beta[MAXBUF], alpha[MAXBUF], p[MAXBUF];
    /* sub-stencils S_r (weno_interpolants.jl:448-471), 1-based: Left: S[n-r+q], Right: S[n+1+r-(q-1)] */
    for (int r = 0; r < n; r++)
        for (int q = 0; q < n; q++) sub[r][q] = left_bias ? S[(n - r + q) - 1] : S[(n + 1 + r - q) - 1];
    for (int r = 0; r < n; r++) {
        const FT *C = g->weno_beta + (n * MAXBUF + r) * 21;
        FT ps[MAXBUF];
#ifdef ORACLE_F32
        FT hat = sub[r][n / 2]; /* ψ[buffer ÷ 2 + 1] (weno_interpolants.jl:276-280) */
        for (int q = 0; q < n; q++) ps[q] = sub[r][q] - hat;
#else
        for (int q = 0; q < n; q++) ps[q] = sub[r][q];
#endif
        /* nested quadratic form under @muladd (weno_interpolants.jl:204-220,265) */
        int c = 0;
        FT b = 0;
        for (int s = 0; s < n - 1; s++) {
            FT inner = C[c] * ps[s];
            for (int q = s + 1; q < n; q++) inner = FMA(C[c + q - s], ps[q], inner);
            b = (s == 0) ? ps[s] * inner : FMA(ps[s], inner, b);
            c += n - s;
        }
        b = FMA(ps[n - 1] * ps[n - 1], C[c], b);
        beta[r] = b;
    }
    FT tau; /* weno_interpolants.jl:324-328 */
    switch (n) {
        case 2: tau = FABS(beta[0] - beta[1]); break;
        case 3: tau = FABS(beta[0] - beta[2]); break;
        case 4: tau = FABS(beta[0] + 3 * beta[1] - 3 * beta[2] - beta[3]); break;
        case 5: tau = FABS(beta[0] + 2 * beta[1] - 6 * beta[2] + 2 * beta[3] + beta[4]); break;
        default: tau = FABS(beta[0] + 36 * beta[1] + 135 * beta[2] - 135 * beta[3] - 36 * beta[4] - beta[5]); break;
    }
    const FT *Cs = g->weno_cstar + n * MAXBUF;
    FT sum = 0;
    for (int r = 0; r < n; r++) {
        FT q = tau / (beta[r] + g->weno_eps); /* CPU newton_div == div_fast (newton_div.jl:57) */
        alpha[r] = Cs[r] * (1 + q * q);
        sum = (r == 0) ? alpha[r] : sum + alpha[r];
    }
    FT inv = 1 / sum;
    for (int r = 0; r < n; r++) {
        const FT *cp = g->weno_coeff + (n * MAXBUF + r) * MAXBUF;
        FT a = cp[0] * sub[r][0]; /* sum(coeff .* ψ), no muladd (weno_interpolants.jl:136-137) */
        for (int q = 1; q < n; q++) a = a + cp[q] * sub[r][q];
        p[r] = a;
    }
    FT res = (alpha[0] * inv) * p[0];
    for (int r = 1; r < n; r++) res = FMA(alpha[r] * inv, p[r], res);
    return res;
}

/* outside_*_halo predicates (topologically_conditional_interpolation.jl:52-58) */
static inline int outside_sym(int is_center, int i, int N, int H) {
    return is_center ? ((i >= H) & (i <= N + 1 - H)) : ((i >= H + 1) & (i <= N + 1 - H));
}
static inline int outside_biased(int is_center, int i, int N, int H) {
    return is_center ? ((i >= H) & (i <= N + 1 - (H - 1)) & (i >= H - 1) & (i <= N + 1 - H))
                     : ((i >= H + 1) & (i <= N + 1 - (H - 1)) & (i >= H) & (i <= N + 1 - H));
}

/* symmetric_interpolate for a scheme (no topology logic) */
static FT sym_plain(P g, const scheme *s, int dir, int is_center, int i, int j, int k, getfn f, const void *arg) {
    const scheme *c = s->kind == ADV_WENO ? s->advecting : s; /* upwind_biased_reconstruction.jl:89-94 */
    int ii = i, jj = j, kk = k;
    if (is_center) { if (dir == 0) ii++; else if (dir == 1) jj++; else kk++; } /* reconstruction_coefficients.jl:6-10 */
    return centered_face(g, c->n, dir, ii, jj, kk, f, arg);
}
static FT biased_plain(P g, const scheme *s, int dir, int is_center, int left, int i, int j, int k, getfn f, const void *arg) {
    int ii = i, jj = j, kk = k;
    if (is_center) { if (dir == 0) ii++; else if (dir == 1) jj++; else kk++; }
    if (s->kind == ADV_WENO) return weno_face(g, s->n, dir, left, ii, jj, kk, f, arg);
    return centered_face(g, s->n, dir, ii, jj, kk, f, arg); /* centered_reconstruction.jl:44-49 */
}

/* _symmetric_interpolate_* / _biased_interpolate_* with the Bounded fallback chain and Flat rule */
static FT sym_interp(P g, const scheme *s, int dir, int is_center, int i, int j, int k, getfn f, const void *arg) {
    if (g->topo[dir] == FLAT) return f(g, arg, i, j, k); /* flat_advective_fluxes.jl:32-49 */
    if (g->topo[dir] == BOUNDED) {
        int idx = dir == 0 ? i : dir == 1 ? j : k;
        while (!(s->kind == ADV_CENTERED && s->n == 1)) { /* HOADV */
            if (outside_sym(is_center, idx, g->N[dir], s->n)) break;
            s = s->buffer_scheme;
        }
    }
    return sym_plain(g, s, dir, is_center, i, j, k, f, arg);
}
static FT biased_interp(P g, const scheme *s, int dir, int is_center, int left, int i, int j, int k, getfn f, const void *arg) {
    if (g->topo[dir] == FLAT) return f(g, arg, i, j, k);
    if (g->topo[dir] == BOUNDED) {
        int idx = dir == 0 ? i : dir == 1 ? j : k;
        while (!(s->kind == ADV_CENTERED && s->n == 1)) {
            if (outside_biased(is_center, idx, g->N[dir], s->n)) break;
            s = s->buffer_scheme;
        }
    }
    return biased_plain(g, s, dir, is_center, left, i, j, k, f, arg);
}

static const scheme *top_scheme(P g) {
    build_chains();
    if (g->adv_kind == ADV_WENO) return &weno_chain[g->adv_buffer];
    if (g->adv_kind == ADV_CENTERED) return &cen_chain[g->adv_buffer];
    return NULL;
}

/* ------------------------------------------------------------------------------------------
 * Advective fluxes.  flux id: 0..8 momentum (Uu,Vu,Wu,Uv,Vv,Wv,Uw,Vw,Ww), tracer x,y,z
 * upwind_biased_advective_fluxes.jl:23-121 ; centered_advective_fluxes.jl:19-37
 * ------------------------------------------------------------------------------------------ */
enum { CEN = 1, FACE = 0 };

static FT mom_flux(P g, int adv_dir, int comp, int i, int j, int k) {
    /* adv_dir: direction of the advecting velocity (0 U,1 V,2 W); comp: advected component (0 u,1 v,2 w) */
    const scheme *s = top_scheme(g);
    if (!s || g->topo[adv_dir] == FLAT) return 0; /* flat_advective_fluxes.jl:9-27 */
    const ofield *Uf = adv_dir == 0 ? &g->u : adv_dir == 1 ? &g->v : &g->w;
    const ofield *q = comp == 0 ? &g->u : comp == 1 ? &g->v : &g->w;
    getfn Aq = adv_dir == 0 ? get_Ax_u : adv_dir == 1 ? get_Ay_v : get_Az_w;
    /* the advecting velocity is interpolated along `comp` (to the location of the flux), the advected
     * component along `adv_dir`.  When adv_dir == comp both are center interpolations. */
    int loc = (adv_dir == comp) ? CEN : FACE;
    if (s->kind == ADV_WENO) {
        FT ut = sym_interp(g, s, comp, loc, i, j, k, Aq, Uf);
        FT qr = biased_interp(g, s, adv_dir, loc, ut > 0, i, j, k, get_field, q);
        return ut * qr;
    } else {
        /* A * ℑ(U) * ℑ(q) with A at the flux location */
        FT A;
        if (adv_dir == 0) {          /* Ax = Δy * Δz at (·, LY, LZ) of the flux */
            FT dy = comp == 1 ? dyf(j) : dyc(j);
            FT dz = comp == 2 ? dzf(k) : dzc(k);
            A = dy * dz;
        } else if (adv_dir == 1) {   /* Ay = Δx * Δz */
            FT dx = comp == 0 ? dxf(i) : dxc(i);
            FT dz = comp == 2 ? dzf(k) : dzc(k);
            A = dx * dz;
        } else {                     /* Az = Δx * Δy */
            FT dx = comp == 0 ? dxf(i) : dxc(i);
            FT dy = comp == 1 ? dyf(j) : dyc(j);
            A = dx * dy;
        }
        FT ut = sym_interp(g, s, comp, loc, i, j, k, get_field, Uf);
        FT qr = sym_interp(g, s, adv_dir, loc, i, j, k, get_field, q);
        return A * ut * qr;
    }
}

static FT tracer_flux(P g, int dir, const ofield *c, int i, int j, int k) {
    const scheme *s = top_scheme(g);
    if (!s || g->topo[dir] == FLAT) return 0;
    const ofield *Uf = dir == 0 ? &g->u : dir == 1 ? &g->v : &g->w;
    FT A = dir == 0 ? dyc(j) * dzc(k) : dir == 1 ? dxc(i) * dzc(k) : dxc(i) * dyc(j);
    if (s->kind == ADV_WENO) {
        FT ut = at(Uf, i, j, k);
        FT cr = biased_interp(g, s, dir, FACE, ut > 0, i, j, k, get_field, c);
        return A * ut * cr;
    } else {
        FT Aq = A * at(Uf, i, j, k);
        return Aq * sym_interp(g, s, dir, FACE, i, j, k, get_field, c);
    }
}

/* δ of a flux function: forward (center-type result from face fluxes): f(i+1)-f(i);
 * backward (face-type result): f(i)-f(i-1).  Flat => 0 (difference_operators.jl:20-48) */
#define DELTA(expr_hi, expr_lo, d) (g->topo[d] == FLAT ? (FT)0 : ((expr_hi) - (expr_lo)))

/* momentum_advection_operators.jl:47-84 */
static FT div_Uu(P g, int i, int j, int k) {
    FT Vi = 1 / ((dxf(i) * dyc(j)) * dzc(k));
    return Vi * (DELTA(mom_flux(g, 0, 0, i, j, k), mom_flux(g, 0, 0, i - 1, j, k), 0) +
                 DELTA(mom_flux(g, 1, 0, i, j + 1, k), mom_flux(g, 1, 0, i, j, k), 1) +
                 DELTA(mom_flux(g, 2, 0, i, j, k + 1), mom_flux(g, 2, 0, i, j, k), 2));
}
static FT div_Uv(P g, int i, int j, int k) {
    FT Vi = 1 / ((dxc(i) * dyf(j)) * dzc(k));
    return Vi * (DELTA(mom_flux(g, 0, 1, i + 1, j, k), mom_flux(g, 0, 1, i, j, k), 0) +
                 DELTA(mom_flux(g, 1, 1, i, j, k), mom_flux(g, 1, 1, i, j - 1, k), 1) +
                 DELTA(mom_flux(g, 2, 1, i, j, k + 1), mom_flux(g, 2, 1, i, j, k), 2));
}
static FT div_Uw(P g, int i, int j, int k) {
    FT Vi = 1 / ((dxc(i) * dyc(j)) * dzf(k));
    return Vi * (DELTA(mom_flux(g, 0, 2, i + 1, j, k), mom_flux(g, 0, 2, i, j, k), 0) +
                 DELTA(mom_flux(g, 1, 2, i, j + 1, k), mom_flux(g, 1, 2, i, j, k), 1) +
                 DELTA(mom_flux(g, 2, 2, i, j, k), mom_flux(g, 2, 2, i, j, k - 1), 2));
}
/* tracer_advection_operators.jl:31-35 */
static FT div_Uc(P g, const ofield *c, int i, int j, int k) {
    FT Vi = 1 / ((dxc(i) * dyc(j)) * dzc(k));
    return Vi * (DELTA(tracer_flux(g, 0, c, i + 1, j, k), tracer_flux(g, 0, c, i, j, k), 0) +
                 DELTA(tracer_flux(g, 1, c, i, j + 1, k), tracer_flux(g, 1, c, i, j, k), 1) +
                 DELTA(tracer_flux(g, 2, c, i, j, k + 1), tracer_flux(g, 2, c, i, j, k), 2));
}

/* ------------------------------------------------------------------------------------------
 * Velocity / tracer gradients (src/TurbulenceClosures/velocity_tracer_gradients.jl:6-42),
 * ∂ = δ * (1/Δ) (src/Operators/derivative_operators.jl:20-30); δ in a Flat direction is 0.
 * ------------------------------------------------------------------------------------------ */
#define U_(i, j, k) at(&g->u, i, j, k)
#define V_(i, j, k) at(&g->v, i, j, k)
#define W_(i, j, k) at(&g->w, i, j, k)
#define FLATX (g->topo[0] == FLAT)
#define FLATY (g->topo[1] == FLAT)
#define FLATZ (g->topo[2] == FLAT)

static FT dx_u(P g, int i, int j, int k) { return (FLATX ? 0 : U_(i + 1, j, k) - U_(i, j, k)) * (1 / dxc(i)); } /* ccc */
static FT dy_v(P g, int i, int j, int k) { return (FLATY ? 0 : V_(i, j + 1, k) - V_(i, j, k)) * (1 / dyc(j)); }
static FT dz_w(P g, int i, int j, int k) { return (FLATZ ? 0 : W_(i, j, k + 1) - W_(i, j, k)) * (1 / dzc(k)); }
static FT dx_v(P g, int i, int j, int k) { return (FLATX ? 0 : V_(i, j, k) - V_(i - 1, j, k)) * (1 / dxf(i)); } /* ffc */
static FT dy_u(P g, int i, int j, int k) { return (FLATY ? 0 : U_(i, j, k) - U_(i, j - 1, k)) * (1 / dyf(j)); } /* ffc */
static FT dx_w(P g, int i, int j, int k) { return (FLATX ? 0 : W_(i, j, k) - W_(i - 1, j, k)) * (1 / dxf(i)); } /* fcf */
static FT dz_u(P g, int i, int j, int k) { return (FLATZ ? 0 : U_(i, j, k) - U_(i, j, k - 1)) * (1 / dzf(k)); } /* fcf */
static FT dy_w(P g, int i, int j, int k) { return (FLATY ? 0 : W_(i, j, k) - W_(i, j - 1, k)) * (1 / dyf(j)); } /* cff */
static FT dz_v(P g, int i, int j, int k) { return (FLATZ ? 0 : V_(i, j, k) - V_(i, j, k - 1)) * (1 / dzf(k)); } /* cff */

static FT S11(P g, int i, int j, int k) { return dx_u(g, i, j, k); }
static FT S22(P g, int i, int j, int k) { return dy_v(g, i, j, k); }
static FT S33(P g, int i, int j, int k) { return dz_w(g, i, j, k); }
static FT S12(P g, int i, int j, int k) { return (FT)0.5 * (dy_u(g, i, j, k) + dx_v(g, i, j, k)); }
static FT S13(P g, int i, int j, int k) { return (FT)0.5 * (dz_u(g, i, j, k) + dx_w(g, i, j, k)); }
static FT S23(P g, int i, int j, int k) { return (FT)0.5 * (dz_v(g, i, j, k) + dy_w(g, i, j, k)); }

/* two-point interpolations of a pointwise function (interpolation_operators.jl:8-71); Flat => identity */
typedef FT (*ptfn)(P g, const void *arg, int i, int j, int k);
static FT Ic(P g, int d, ptfn f, const void *a, int i, int j, int k) { /* ℑ to center: (f(i) + f(i+1))/2 */
    if (g->topo[d] == FLAT) return f(g, a, i, j, k);
    return (FT)0.5 * (f(g, a, i, j, k) + (d == 0 ? f(g, a, i + 1, j, k) : d == 1 ? f(g, a, i, j + 1, k) : f(g, a, i, j, k + 1)));
}
static FT If(P g, int d, ptfn f, const void *a, int i, int j, int k) { /* ℑ to face: (f(i-1) + f(i))/2 */
    if (g->topo[d] == FLAT) return f(g, a, i, j, k);
    return (FT)0.5 * ((d == 0 ? f(g, a, i - 1, j, k) : d == 1 ? f(g, a, i, j - 1, k) : f(g, a, i, j, k - 1)) + f(g, a, i, j, k));
}
/* nested double interpolation: outer(d2) of inner(d1), e.g. ℑxyᶠᶠᵃ = ℑyᵃᶠᵃ(ℑxᶠᵃᵃ f) */
typedef struct { int d1, face1; ptfn f; const void *a; } nest1;
static FT nest_eval(P g, const void *arg, int i, int j, int k) {
    const nest1 *n = (const nest1 *)arg;
    return n->face1 ? If(g, n->d1, n->f, n->a, i, j, k) : Ic(g, n->d1, n->f, n->a, i, j, k);
}
static FT I2(P g, int d_outer, int face_outer, int d_inner, int face_inner, ptfn f, const void *a, int i, int j, int k) {
    nest1 n = {d_inner, face_inner, f, a};
    return face_outer ? If(g, d_outer, nest_eval, &n, i, j, k) : Ic(g, d_outer, nest_eval, &n, i, j, k);
}
static FT pt_field(P g, const void *a, int i, int j, int k) { (void)g; return at((const ofield *)a, i, j, k); }

/* viscosity at the four stress locations for closure m (abstract_scalar_diffusivity_closure.jl:330-351) */
static FT nu_ccc(P g, int m, int i, int j, int k) { return g->closure_kind[m] == CL_SCALAR ? g->nu[m] : at(&g->nue[m], i, j, k); }
static FT nu_ffc(P g, int m, int i, int j, int k) { return g->closure_kind[m] == CL_SCALAR ? g->nu[m] : I2(g, 1, 1, 0, 1, pt_field, &g->nue[m], i, j, k); }
static FT nu_fcf(P g, int m, int i, int j, int k) { return g->closure_kind[m] == CL_SCALAR ? g->nu[m] : I2(g, 2, 1, 0, 1, pt_field, &g->nue[m], i, j, k); }
static FT nu_cff(P g, int m, int i, int j, int k) { return g->closure_kind[m] == CL_SCALAR ? g->nu[m] : I2(g, 2, 1, 1, 1, pt_field, &g->nue[m], i, j, k); }

/* Ax_q(viscous_flux): area * (-2 ν Σ)  (closure_kernel_operators.jl:20-40; abstract_scalar...:214-226) */
static FT F_ux(P g, int m, int i, int j, int k) { return (dyc(j) * dzc(k)) * (-2 * (nu_ccc(g, m, i, j, k) * S11(g, i, j, k))); }
static FT F_uy(P g, int m, int i, int j, int k) { return (dxf(i) * dzc(k)) * (-2 * (nu_ffc(g, m, i, j, k) * S12(g, i, j, k))); }
/* VerticallyImplicitTimeDiscretization on a z-Bounded grid (abstract_scalar_diffusivity_closure.jl:270-312): away from the
 * boundary faces k = 1, Nz+1 the explicit vertical fluxes keep only what the tridiagonal solve does not contain:
 * uz = -nu dx(w), vz = -nu dy(w), wz = 0, q_z = 0 */
static int vi_elide(P g, int m, int k) { return g->closure_vi[m] && g->topo[2] == BOUNDED && !((k == 1) | (k == g->N[2] + 1)); }
static FT F_uz(P g, int m, int i, int j, int k) {
    if (vi_elide(g, m, k)) return (dxf(i) * dyc(j)) * (-(nu_fcf(g, m, i, j, k) * dx_w(g, i, j, k)));
    return (dxf(i) * dyc(j)) * (-2 * (nu_fcf(g, m, i, j, k) * S13(g, i, j, k)));
}
static FT F_vx(P g, int m, int i, int j, int k) { return (dyf(j) * dzc(k)) * (-2 * (nu_ffc(g, m, i, j, k) * S12(g, i, j, k))); }
static FT F_vy(P g, int m, int i, int j, int k) { return (dxc(i) * dzc(k)) * (-2 * (nu_ccc(g, m, i, j, k) * S22(g, i, j, k))); }
static FT F_vz(P g, int m, int i, int j, int k) {
    if (vi_elide(g, m, k)) return (dxc(i) * dyf(j)) * (-(nu_cff(g, m, i, j, k) * dy_w(g, i, j, k)));
    return (dxc(i) * dyf(j)) * (-2 * (nu_cff(g, m, i, j, k) * S23(g, i, j, k)));
}
static FT F_wx(P g, int m, int i, int j, int k) { return (dyc(j) * dzf(k)) * (-2 * (nu_fcf(g, m, i, j, k) * S13(g, i, j, k))); }
static FT F_wy(P g, int m, int i, int j, int k) { return (dxc(i) * dzf(k)) * (-2 * (nu_cff(g, m, i, j, k) * S23(g, i, j, k))); }
static FT F_wz(P g, int m, int i, int j, int k) {
    if (vi_elide(g, m, k)) return (dxc(i) * dyc(j)) * (FT)0;
    return (dxc(i) * dyc(j)) * (-2 * (nu_ccc(g, m, i, j, k) * S33(g, i, j, k)));
}

static FT div_tau1(P g, int m, int i, int j, int k) {
    FT Vi = 1 / ((dxf(i) * dyc(j)) * dzc(k));
    return Vi * (DELTA(F_ux(g, m, i, j, k), F_ux(g, m, i - 1, j, k), 0) + DELTA(F_uy(g, m, i, j + 1, k), F_uy(g, m, i, j, k), 1) +
                 DELTA(F_uz(g, m, i, j, k + 1), F_uz(g, m, i, j, k), 2));
}
static FT div_tau2(P g, int m, int i, int j, int k) {
    FT Vi = 1 / ((dxc(i) * dyf(j)) * dzc(k));
    return Vi * (DELTA(F_vx(g, m, i + 1, j, k), F_vx(g, m, i, j, k), 0) + DELTA(F_vy(g, m, i, j, k), F_vy(g, m, i, j - 1, k), 1) +
                 DELTA(F_vz(g, m, i, j, k + 1), F_vz(g, m, i, j, k), 2));
}
static FT div_tau3(P g, int m, int i, int j, int k) {
    FT Vi = 1 / ((dxc(i) * dyc(j)) * dzf(k));
    return Vi * (DELTA(F_wx(g, m, i + 1, j, k), F_wx(g, m, i, j, k), 0) + DELTA(F_wy(g, m, i, j + 1, k), F_wy(g, m, i, j, k), 1) +
                 DELTA(F_wz(g, m, i, j, k), F_wz(g, m, i, j, k - 1), 2));
}

/* diffusivities at fcc / cfc / ccf for closure m, tracer t */
static FT kap(P g, int m, int t, int d, int i, int j, int k) {
    switch (g->closure_kind[m]) {
        case CL_SCALAR: return g->kappa[m][t];
        case CL_SMAG: return If(g, d, pt_field, &g->nue[m], i, j, k) / g->Pr[m][t]; /* smagorinsky.jl:139-141 */
        default: return If(g, d, pt_field, &g->kappae[m][t], i, j, k);
    }
}
static FT Q_x(P g, int m, int t, int i, int j, int k) {
    const ofield *c = &g->c[t];
    FT dc = (FLATX ? 0 : at(c, i, j, k) - at(c, i - 1, j, k)) * (1 / dxf(i));
    return (dyc(j) * dzc(k)) * (-kap(g, m, t, 0, i, j, k) * dc);
}
static FT Q_y(P g, int m, int t, int i, int j, int k) {
    const ofield *c = &g->c[t];
    FT dc = (FLATY ? 0 : at(c, i, j, k) - at(c, i, j - 1, k)) * (1 / dyf(j));
    return (dxc(i) * dzc(k)) * (-kap(g, m, t, 1, i, j, k) * dc);
}
static FT Q_z(P g, int m, int t, int i, int j, int k) {
    const ofield *c = &g->c[t];
    if (vi_elide(g, m, k)) return (dxc(i) * dyc(j)) * (FT)0;
    FT dc = (FLATZ ? 0 : at(c, i, j, k) - at(c, i, j, k - 1)) * (1 / dzf(k));
    return (dxc(i) * dyc(j)) * (-kap(g, m, t, 2, i, j, k) * dc);
}
static FT div_q(P g, int m, int t, int i, int j, int k) {
    FT Vi = 1 / ((dxc(i) * dyc(j)) * dzc(k));
    return Vi * (DELTA(Q_x(g, m, t, i + 1, j, k), Q_x(g, m, t, i, j, k), 0) + DELTA(Q_y(g, m, t, i, j + 1, k), Q_y(g, m, t, i, j, k), 1) +
                 DELTA(Q_z(g, m, t, i, j, k + 1), Q_z(g, m, t, i, j, k), 2));
}

/* ------------------------------------------------------------------------------------------
 * Buoyancy, Coriolis, hydrostatic pressure gradient
 * ------------------------------------------------------------------------------------------ */
/* buoyancy_perturbationᶜᶜᶜ: buoyancy_tracer.jl:12 ; linear_equation_of_state.jl:72-74 */
static FT bpert(P g, const void *unused, int i, int j, int k) {
    (void)unused;
    if (g->buoy_kind == BUOY_TRACER) return at(&g->c[g->ib], i, j, k);
    if (g->buoy_kind == BUOY_SEAWATER) return g->g * (g->alpha * at(&g->c[g->iT], i, j, k) - g->beta * at(&g->c[g->iS], i, j, k));
    return 0;
}
/* masked_Ay_qᶜᶠᶜ == Ay_qᶜᶠᶜ on a non-immersed grid (coriolis_schemes.jl:55-58) */
static FT pt_Ay_v(P g, const void *a, int i, int j, int k) { (void)a; return (dxc(i) * dzc(k)) * V_(i, j, k); }
static FT pt_Ax_u(P g, const void *a, int i, int j, int k) { (void)a; return (dyc(j) * dzc(k)) * U_(i, j, k); }

/* ------------------------------------------------------------------------------------------
 * Tendencies (nonhydrostatic_tendency_kernel_functions.jl:71-302)
 * ------------------------------------------------------------------------------------------ */
static FT Gu_point(P g, int i, int j, int k) {
    FT r = -div_Uu(g, i, j, k);
    r = r - 0; /* background advection */
    r = r - 0; /* metric terms: 0 on RectilinearGrid (curvature_metric_terms.jl:36-38) */
    r = r + 0; /* x_dot_g_b: gravity is -z (g_dot_b.jl:6-9) */
    if (g->has_coriolis) { /* coriolis_schemes.jl:67 ; f constant so ℑy(f) = 0.5*(f+f) */
        FT fbar = FLATY ? g->fcor : (FT)0.5 * (g->fcor + g->fcor);
        FT Ayv = I2(g, 1, 0, 0, 1, pt_Ay_v, NULL, i, j, k); /* ℑxyᶠᶜᵃ = ℑyᵃᶜᵃ(ℑxᶠᵃᵃ ·) */
        FT xf = -fbar * Ayv * (1 / (dxf(i) * dzc(k)));       /* Ay⁻¹ᶠᶜᶜ */
        r = r - xf;
    } else
        r = r - 0;
    if (g->has_pHY) r = r - (FLATX ? 0 : at(&g->pHY, i, j, k) - at(&g->pHY, i - 1, j, k)) * (1 / dxf(i));
    else r = r - 0;
    if (g->nclosures > 0) {
        FT t = div_tau1(g, 0, i, j, k);
        for (int m = 1; m < g->nclosures; m++) t = t + div_tau1(g, m, i, j, k);
        r = r - t;
    } else
        r = r - 0;
    return r;
}
static FT Gv_point(P g, int i, int j, int k) {
    FT r = -div_Uv(g, i, j, k);
    if (g->has_coriolis) {
        FT fbar = FLATX ? g->fcor : (FT)0.5 * (g->fcor + g->fcor);
        FT Axu = I2(g, 1, 1, 0, 0, pt_Ax_u, NULL, i, j, k); /* ℑxyᶜᶠᵃ = ℑyᵃᶠᵃ(ℑxᶜᵃᵃ ·) */
        FT yf = fbar * Axu * (1 / (dyf(j) * dzc(k)));        /* Ax⁻¹ᶜᶠᶜ */
        r = r - yf;
    }
    if (g->has_pHY) r = r - (FLATY ? 0 : at(&g->pHY, i, j, k) - at(&g->pHY, i, j - 1, k)) * (1 / dyf(j));
    if (g->nclosures > 0) {
        FT t = div_tau2(g, 0, i, j, k);
        for (int m = 1; m < g->nclosures; m++) t = t + div_tau2(g, m, i, j, k);
        r = r - t;
    }
    return r;
}
static FT Gw_point(P g, int i, int j, int k) {
    FT r = -div_Uw(g, i, j, k);
    /* maybe_z_dot_g_b: only when there is no hydrostatic pressure field (…kernel_functions.jl:171-173) */
    if (!g->has_pHY && g->buoy_kind != BUOY_NONE) r = r + If(g, 2, bpert, NULL, i, j, k);
    if (g->nclosures > 0) {
        FT t = div_tau3(g, 0, i, j, k);
        for (int m = 1; m < g->nclosures; m++) t = t + div_tau3(g, m, i, j, k);
        r = r - t;
    }
    return r;
}
static FT Gc_point(P g, int t, int i, int j, int k) {
    FT r = -div_Uc(g, &g->c[t], i, j, k);
    if (g->nclosures > 0) {
        FT q = div_q(g, 0, t, i, j, k);
        for (int m = 1; m < g->nclosures; m++) q = q + div_q(g, m, t, i, j, k);
        r = r - q;
    }
    return r;
}

/* one "kernel launch" per tendency over size(grid) (compute_nonhydrostatic_tendencies.jl:67-97) */
/* ------------------------------------------------------------------------------------------
 * implicit_step! (src/TurbulenceClosures/vertically_implicit_diffusion_solver.jl:60-136,196-225) through
 * solve_batched_tridiagonal_system_z! (src/Solvers/batched_tridiagonal_solver.jl:211-243), in place on the field.
 * which: 0 u (f,c,c), 1 v (c,f,c), 2 w (c,c,f), 3+t tracer t (c,c,c).  t_scratch: Nz+1 values per column.
 * inactive_cell / inactive_node / peripheral_node: src/Grids/inactive_node.jl:43-165.
 * ------------------------------------------------------------------------------------------ */
static int inactive_cell(P g, int i, int j, int k) {
    return ((g->topo[0] == BOUNDED) & ((i < 1) | (i > g->N[0]))) | ((g->topo[1] == BOUNDED) & ((j < 1) | (j > g->N[1]))) |
           ((g->topo[2] == BOUNDED) & ((k < 1) | (k > g->N[2])));
}
/* any = 1: peripheral_node (OR over the cells around the node), any = 0: inactive_node (AND) */
static int node_test(P g, int i, int j, int k, int fx, int fy, int fz, int any) {
    int r = any ? 0 : 1;
    for (int a = 0; a <= fx; a++)
        for (int b = 0; b <= fy; b++)
            for (int c = 0; c <= fz; c++) {
                int v = inactive_cell(g, i - a, j - b, k - c);
                r = any ? (r | v) : (r & v);
            }
    return r;
}
static FT strong(FT x, int keep) { return keep ? x : (FT)0; } /* x * Bool */
/* νzᶠᶜᶠ / νzᶜᶠᶠ / νzᶜᶜᶜ / κzᶜᶜᶠ at level kk (abstract_scalar_diffusivity_closure.jl:137-151): the constants of a ScalarDiffusivity, or
 * the eddy viscosity / diffusivity of Smagorinsky and AMD interpolated to the node the coefficient belongs to */
static FT ivd_coefficient(P g, int m, int which, int i, int j, int kk) {
    if (g->closure_kind[m] == CL_SCALAR) return which < 3 ? g->nu[m] : g->kappa[m][which - 3];
    if (which == 0) return nu_fcf(g, m, i, j, kk);
    if (which == 1) return nu_cff(g, m, i, j, kk);
    if (which == 2) return nu_ccc(g, m, i, j, kk);
    return kap(g, m, which - 3, 2, i, j, kk);
}
static FT ivd_upper(P g, int which, int lx, int ly, int lz, int i, int j, int k, FT dt) {
    FT sum = 0;
    int first = 1;
    for (int m = 0; m < g->nclosures; m++) {
        if (!g->closure_vi[m]) continue;
        FT d;
        if (!lz) {
            FT kap = strong(ivd_coefficient(g, m, which, i, j, k + 1), !node_test(g, i, j, k + 1, lx, ly, 1, 0));
            d = -dt * kap * ((1 / dzc(k)) * (1 / dzf(k + 1)));
            d = strong(d, !node_test(g, i, j, k + 1, lx, ly, 1, 1));
        } else {
            FT nu = strong(ivd_coefficient(g, m, which, i, j, k), !node_test(g, i, j, k, lx, ly, 0, 0));
            d = -dt * nu * ((1 / dzc(k)) * (1 / dzf(k)));
            d = strong(d, !node_test(g, i, j, k, lx, ly, 0, 1));
        }
        sum = first ? d : sum + d;
        first = 0;
    }
    return sum;
}
static FT ivd_lower(P g, int which, int lx, int ly, int lz, int i, int j, int kk, FT dt) {
    FT sum = 0;
    int first = 1;
    for (int m = 0; m < g->nclosures; m++) {
        if (!g->closure_vi[m]) continue;
        FT d;
        if (!lz) {
            int k = kk + 1;
            FT kap = strong(ivd_coefficient(g, m, which, i, j, k), !node_test(g, i, j, k, lx, ly, 1, 0));
            d = -dt * kap * ((1 / dzc(k)) * (1 / dzf(k)));
            d = strong(d, !node_test(g, i, j, kk, lx, ly, 0, 1));
        } else {
            int kp = kk + 2;
            FT nu = strong(ivd_coefficient(g, m, which, i, j, kp - 1), !node_test(g, i, j, kp - 1, lx, ly, 0, 0));
            d = -dt * nu * ((1 / dzc(kp)) * (1 / dzf(kp - 1)));
            d = strong(d, !node_test(g, i, j, kk, lx, ly, 0, 1));
        }
        sum = first ? d : sum + d;
        first = 0;
    }
    return sum;
}
static FT ivd_diag(P g, int which, int lx, int ly, int lz, int i, int j, int k, FT dt) {
    return (FT)1 - dt * (FT)0 - ivd_upper(g, which, lx, ly, lz, i, j, k, dt) - ivd_lower(g, which, lx, ly, lz, i, j, k - 1, dt);
}
void SUF(orc_implicit_step)(const oparams *g, int which, ofield *phi, FT dt, FT *t) {
    const int lx = which == 0, ly = which == 1, lz = which == 2;
    const int Nz = g->N[2];
#ifdef ORACLE_F32
    const FT tiny = 10 * 1.1920929e-07f;
#else
    const FT tiny = 10 * 2.220446049250313e-16;
#endif
    for (int j = 1; j <= g->N[1]; j++)
        for (int i = 1; i <= g->N[0]; i++) {
            FT beta = ivd_diag(g, which, lx, ly, lz, i, j, 1, dt);
            *ref(phi, i, j, 1) = at(phi, i, j, 1) / beta;
            for (int k = 2; k <= Nz; k++) {
                FT cm = ivd_upper(g, which, lx, ly, lz, i, j, k - 1, dt);
                FT bk = ivd_diag(g, which, lx, ly, lz, i, j, k, dt);
                FT am = ivd_lower(g, which, lx, ly, lz, i, j, k - 1, dt);
                t[k] = cm / beta;
                beta = bk - am * t[k];
                FT fk = at(phi, i, j, k);
                FT cand = (fk - am * at(phi, i, j, k - 1)) / beta;
                if (FABS(beta) > tiny) *ref(phi, i, j, k) = cand;
            }
            for (int k = Nz - 1; k >= 1; k--) *ref(phi, i, j, k) -= t[k + 1] * at(phi, i, j, k + 1);
        }
}

void SUF(orc_compute_tendencies)(const oparams *g) {
    build_chains();
    const int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= Nz; k++)
        for (int j = 1; j <= Ny; j++)
            for (int i = 1; i <= Nx; i++) *ref(&g->Gu, i, j, k) = Gu_point(g, i, j, k);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= Nz; k++)
        for (int j = 1; j <= Ny; j++)
            for (int i = 1; i <= Nx; i++) *ref(&g->Gv, i, j, k) = Gv_point(g, i, j, k);
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= Nz; k++)
        for (int j = 1; j <= Ny; j++)
            for (int i = 1; i <= Nx; i++) *ref(&g->Gw, i, j, k) = Gw_point(g, i, j, k);
    for (int t = 0; t < g->ntracers; t++) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int k = 1; k <= Nz; k++)
            for (int j = 1; j <= Ny; j++)
                for (int i = 1; i <= Nx; i++) *ref(&g->Gc[t], i, j, k) = Gc_point(g, t, i, j, k);
    }
}

/* single-point probes used by the known-answer tests */
FT SUF(orc_weno_face)(const oparams *g, int n, int dir, int left, const ofield *f, int i, int j, int k) {
    build_chains();
    return weno_face(g, n, dir, left, i, j, k, get_field, f);
}
FT SUF(orc_biased_interp)(const oparams *g, int dir, int is_center, int left, const ofield *f, int i, int j, int k) {
    return biased_interp(g, top_scheme(g), dir, is_center, left, i, j, k, get_field, f);
}
FT SUF(orc_sym_interp)(const oparams *g, int dir, int is_center, const ofield *f, int i, int j, int k) {
    return sym_interp(g, top_scheme(g), dir, is_center, i, j, k, get_field, f);
}
FT SUF(orc_div_tau)(const oparams *g, int comp, int i, int j, int k) {
    return comp == 0 ? div_tau1(g, 0, i, j, k) : comp == 1 ? div_tau2(g, 0, i, j, k) : div_tau3(g, 0, i, j, k);
}
FT SUF(orc_div_q)(const oparams *g, int t, int i, int j, int k) { return div_q(g, 0, t, i, j, k); }
/* WENO smoothness/weights probe (test/test_weno_smoothness.jl): returns beta[n], omega[n] */
void SUF(orc_weno_beta_omega)(const oparams *g, int n, const FT *sub /* n x n */, FT *beta, FT *omega) {
    FT alpha[MAXBUF];
    for (int r = 0; r < n; r++) {
        const FT *C = g->weno_beta + (n * MAXBUF + r) * 21;
        FT ps[MAXBUF];
#ifdef ORACLE_F32
        FT hat = sub[r * n + n / 2];
        for (int q = 0; q < n; q++) ps[q] = sub[r * n + q] - hat;
#else
        for (int q = 0; q < n; q++) ps[q] = sub[r * n + q];
#endif
        int c = 0;
        FT b = 0;
        for (int s = 0; s < n - 1; s++) {
            FT inner = C[c] * ps[s];
            for (int q = s + 1; q < n; q++) inner = FMA(C[c + q - s], ps[q], inner);
            b = (s == 0) ? ps[s] * inner : FMA(ps[s], inner, b);
            c += n - s;
        }
        b = FMA(ps[n - 1] * ps[n - 1], C[c], b);
        beta[r] = b;
    }
    FT tau;
    switch (n) {
        case 2: tau = FABS(beta[0] - beta[1]); break;
        case 3: tau = FABS(beta[0] - beta[2]); break;
        case 4: tau = FABS(beta[0] + 3 * beta[1] - 3 * beta[2] - beta[3]); break;
        case 5: tau = FABS(beta[0] + 2 * beta[1] - 6 * beta[2] + 2 * beta[3] + beta[4]); break;
        default: tau = FABS(beta[0] + 36 * beta[1] + 135 * beta[2] - 135 * beta[3] - 36 * beta[4] - beta[5]); break;
    }
    const FT *Cs = g->weno_cstar + n * MAXBUF;
    FT sum = 0;
    for (int r = 0; r < n; r++) {
        FT q = tau / (beta[r] + g->weno_eps);
        alpha[r] = Cs[r] * (1 + q * q);
        sum = (r == 0) ? alpha[r] : sum + alpha[r];
    }
    FT inv = 1 / sum;
    for (int r = 0; r < n; r++) omega[r] = alpha[r] * inv;
}

/* ------------------------------------------------------------------------------------------
 * Closure fields
 * ------------------------------------------------------------------------------------------ */
static FT pt_S12sq(P g, const void *a, int i, int j, int k) { (void)a; FT s = S12(g, i, j, k); return s * s; }
static FT pt_S13sq(P g, const void *a, int i, int j, int k) { (void)a; FT s = S13(g, i, j, k); return s * s; }
static FT pt_S23sq(P g, const void *a, int i, int j, int k) { (void)a; FT s = S23(g, i, j, k); return s * s; }
/* ∂z_b at ccf (buoyancy_operations): δz(b)/Δzᶠ */
static FT pt_dz_b(P g, const void *a, int i, int j, int k) {
    (void)a;
    return (FLATZ ? 0 : bpert(g, NULL, i, j, k) - bpert(g, NULL, i, j, k - 1)) * (1 / dzf(k));
}

/* _compute_smagorinsky_viscosity! (Smagorinskys/smagorinsky.jl:90-104; lilly_coefficient.jl:129-142;
 * scale_invariant_operators.jl:10-13; velocity_tracer_gradients.jl tr_Σ²) */
void SUF(orc_smagorinsky_viscosity)(const oparams *g, int m) {
    const int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= Nz; k++)
        for (int j = 1; j <= Ny; j++)
            for (int i = 1; i <= Nx; i++) {
                FT s11 = S11(g, i, j, k), s22 = S22(g, i, j, k), s33 = S33(g, i, j, k);
                FT tr = s11 * s11 + s22 * s22 + s33 * s33; /* tr_Σ² = Σ₁₁² + Σ₂₂² + Σ₃₃² */
                FT SS = tr + 2 * I2(g, 1, 0, 0, 0, pt_S12sq, NULL, i, j, k) + 2 * I2(g, 2, 0, 0, 0, pt_S13sq, NULL, i, j, k) +
                        2 * I2(g, 2, 0, 1, 0, pt_S23sq, NULL, i, j, k);
                FT D3 = dxc(i) * dyc(j) * dzc(k);
                FT Df = CBRT(D3);
                FT cs2;
                if (g->lilly[m]) {
                    FT N2 = Ic(g, 2, pt_dz_b, NULL, i, j, k);
                    FT N2p = FMAX((FT)0, N2);
                    FT s2 = (FT)1 - FMIN((FT)1, g->cb[m] * N2p / SS);
                    FT st = (SS == 0) ? (FT)0 : SQRT(s2);
                    cs2 = st * (g->cs[m] * g->cs[m]);
                } else
                    cs2 = g->cs[m] * g->cs[m];
                *ref(&g->nue[m], i, j, k) = cs2 * (Df * Df) * SQRT(2 * SS);
            }
}

/* ------------------------------------------------------------------------------------------
 * AnisotropicMinimumDissipation (turbulence_closure_implementations/anisotropic_minimum_dissipation.jl:161-358;
 * normalised gradients velocity_tracer_gradients.jl:126-260).  Filter widths are 2Δ at ccc, evaluated at the
 * index of the stencil point whatever its location (:231-243).
 * ------------------------------------------------------------------------------------------ */
#define DFX(i) ((FT)2 * dxc(i))
#define DFY(j) ((FT)2 * dyc(j))
#define DFZ(k) ((FT)2 * dzc(k))
static FT n_dx_u(P g, const void *a, int i, int j, int k) { (void)a; return dx_u(g, i, j, k); }
static FT n_dy_v(P g, const void *a, int i, int j, int k) { (void)a; return dy_v(g, i, j, k); }
static FT n_dz_w(P g, const void *a, int i, int j, int k) { (void)a; return dz_w(g, i, j, k); }
static FT n_dx_v(P g, const void *a, int i, int j, int k) { (void)a; return DFX(i) / DFY(j) * dx_v(g, i, j, k); }
static FT n_dy_u(P g, const void *a, int i, int j, int k) { (void)a; return DFY(j) / DFX(i) * dy_u(g, i, j, k); }
static FT n_dx_w(P g, const void *a, int i, int j, int k) { (void)a; return DFX(i) / DFZ(k) * dx_w(g, i, j, k); }
static FT n_dz_u(P g, const void *a, int i, int j, int k) { (void)a; return DFZ(k) / DFX(i) * dz_u(g, i, j, k); }
static FT n_dy_w(P g, const void *a, int i, int j, int k) { (void)a; return DFY(j) / DFZ(k) * dy_w(g, i, j, k); }
static FT n_dz_v(P g, const void *a, int i, int j, int k) { (void)a; return DFZ(k) / DFY(j) * dz_v(g, i, j, k); }
static FT n_S12(P g, const void *a, int i, int j, int k) { return (FT)0.5 * (n_dy_u(g, a, i, j, k) + n_dx_v(g, a, i, j, k)); }
static FT n_S13(P g, const void *a, int i, int j, int k) { return (FT)0.5 * (n_dz_u(g, a, i, j, k) + n_dx_w(g, a, i, j, k)); }
static FT n_S23(P g, const void *a, int i, int j, int k) { return (FT)0.5 * (n_dz_v(g, a, i, j, k) + n_dy_w(g, a, i, j, k)); }
#define SQFN(name, base) static FT name(P g, const void *a, int i, int j, int k) { FT x = base(g, a, i, j, k); return x * x; }
SQFN(n_dx_v2, n_dx_v) SQFN(n_dy_u2, n_dy_u) SQFN(n_dx_w2, n_dx_w) SQFN(n_dz_u2, n_dz_u) SQFN(n_dy_w2, n_dy_w) SQFN(n_dz_v2, n_dz_v)
#define PRFN(name, f1, f2) static FT name(P g, const void *a, int i, int j, int k) { return f1(g, a, i, j, k) * f2(g, a, i, j, k); }
PRFN(n_dx_v_S12, n_dx_v, n_S12) PRFN(n_dy_u_S12, n_dy_u, n_S12) PRFN(n_dx_w_S13, n_dx_w, n_S13) PRFN(n_dz_u_S13, n_dz_u, n_S13)
PRFN(n_dz_v_S23, n_dz_v, n_S23) PRFN(n_dy_w_S23, n_dy_w, n_S23)
/* ℑxyᶜᶜᵃ = ℑyᵃᶜᵃ(ℑxᶜᵃᵃ), ℑxzᶜᵃᶜ = ℑzᵃᵃᶜ(ℑxᶜᵃᵃ), ℑyzᵃᶜᶜ = ℑzᵃᵃᶜ(ℑyᵃᶜᵃ)  (interpolation_operators.jl:45-53) */
#define IXY(f) I2(g, 1, 0, 0, 0, f, NULL, i, j, k)
#define IXZ(f) I2(g, 2, 0, 0, 0, f, NULL, i, j, k)
#define IYZ(f) I2(g, 2, 0, 1, 0, f, NULL, i, j, k)

/* tracer gradients: norm_∂x_c = Δᶠx * ∂xᶠᶜᶜ c, ... */
static FT n_dx_c(P g, const void *a, int i, int j, int k) { const ofield *c = a; return DFX(i) * ((FLATX ? 0 : at(c, i, j, k) - at(c, i - 1, j, k)) * (1 / dxf(i))); }
static FT n_dy_c(P g, const void *a, int i, int j, int k) { const ofield *c = a; return DFY(j) * ((FLATY ? 0 : at(c, i, j, k) - at(c, i, j - 1, k)) * (1 / dyf(j))); }
static FT n_dz_c(P g, const void *a, int i, int j, int k) { const ofield *c = a; return DFZ(k) * ((FLATZ ? 0 : at(c, i, j, k) - at(c, i, j, k - 1)) * (1 / dzf(k))); }
SQFN(n_dx_c2, n_dx_c) SQFN(n_dy_c2, n_dy_c) SQFN(n_dz_c2, n_dz_c)
/* ∂x b at fcc etc. of the buoyancy perturbation */
static FT pt_dx_b(P g, const void *a, int i, int j, int k) { (void)a; return (FLATX ? 0 : bpert(g, NULL, i, j, k) - bpert(g, NULL, i - 1, j, k)) * (1 / dxf(i)); }
static FT pt_dy_b(P g, const void *a, int i, int j, int k) { (void)a; return (FLATY ? 0 : bpert(g, NULL, i, j, k) - bpert(g, NULL, i, j - 1, k)) * (1 / dyf(j)); }

static FT amd_delta2(P g, int i, int j, int k) {
    FT fx = DFX(i), fy = DFY(j), fz = DFZ(k);
    return 3 / (1 / (fx * fx) + 1 / (fy * fy) + 1 / (fz * fz));
}

void SUF(orc_amd_viscosity)(const oparams *g, int m) {
    const int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= Nz; k++)
        for (int j = 1; j <= Ny; j++)
            for (int i = 1; i <= Nx; i++) {
                const FT ux = dx_u(g, i, j, k), vy = dy_v(g, i, j, k), wz = dz_w(g, i, j, k);
                /* norm_tr_∇uᶜᶜᶜ (:295-316) */
                FT q = ux * ux + vy * vy + wz * wz + IXY(n_dx_v2) + IXY(n_dy_u2) + IXZ(n_dx_w2) + IXZ(n_dz_u2) + IYZ(n_dy_w2) + IYZ(n_dz_v2);
                FT nu = 0;
                if (q != 0) {
                    /* norm_uᵢₐ_uⱼₐ_Σᵢⱼᶜᶜᶜ (:251-289) */
                    FT r1 = ux * (ux * ux) + vy * IXY(n_dx_v2) + wz * IXZ(n_dx_w2) + 2 * ux * IXY(n_dx_v_S12) + 2 * ux * IXZ(n_dx_w_S13) +
                            2 * IXY(n_dx_v) * IXZ(n_dx_w) * IYZ(n_S23);
                    FT r2 = ux * IXY(n_dy_u2) + vy * (vy * vy) + wz * IYZ(n_dy_w2) + 2 * vy * IXY(n_dy_u_S12) +
                            2 * IXY(n_dy_u) * IYZ(n_dy_w) * IXZ(n_S13) + 2 * vy * IYZ(n_dy_w_S23);
                    FT r3 = ux * IXZ(n_dz_u2) + vy * IYZ(n_dz_v2) + wz * (wz * wz) + 2 * IXZ(n_dz_u) * IYZ(n_dz_v) * IXY(n_S12) +
                            2 * wz * IXZ(n_dz_u_S13) + 2 * wz * IYZ(n_dz_v_S23);
                    FT r = r1 + r2 + r3;
                    FT cbz = 0;
                    if (g->amd_has_cb[m]) { /* Cb_norm_wᵢ_bᵢᶜᶜᶜ (:320-333) */
                        FT wxbx = IXZ(n_dx_w) * DFX(i) * Ic(g, 0, pt_dx_b, NULL, i, j, k);
                        FT wyby = IYZ(n_dy_w) * DFY(j) * Ic(g, 1, pt_dy_b, NULL, i, j, k);
                        FT wzbz = wz * DFZ(k) * Ic(g, 2, pt_dz_b, NULL, i, j, k);
                        cbz = g->cb[m] * (wxbx + wyby + wzbz);
                    }
                    cbz = cbz / DFZ(k);
                    nu = -g->Cnu[m] * amd_delta2(g, i, j, k) * (r - cbz) / q;
                }
                *ref(&g->nue[m], i, j, k) = FMAX((FT)0, nu);
            }
}

void SUF(orc_amd_diffusivity)(const oparams *g, int m, int t) {
    const int Nx = g->N[0], Ny = g->N[1], Nz = g->N[2];
    const ofield *c = &g->c[t];
#define ICX(f) Ic(g, 0, f, c, i, j, k)
#define ICY(f) Ic(g, 1, f, c, i, j, k)
#define ICZ(f) Ic(g, 2, f, c, i, j, k)
#pragma omp parallel for collapse(2) schedule(static)
    for (int k = 1; k <= Nz; k++)
        for (int j = 1; j <= Ny; j++)
            for (int i = 1; i <= Nx; i++) {
                FT sigma = ICX(n_dx_c2) + ICY(n_dy_c2) + ICZ(n_dz_c2); /* norm_θᵢ²ᶜᶜᶜ (:355-357) */
                FT kap_ = 0;
                if (sigma != 0) {
                    const FT ux = dx_u(g, i, j, k), vy = dy_v(g, i, j, k), wz = dz_w(g, i, j, k);
                    /* norm_uᵢⱼ_cⱼ_cᵢᶜᶜᶜ (:335-353); the ℑxz of norm_∂y_w in cy_uy is the reference's own */
                    FT cx = ux * ICX(n_dx_c2) + IXY(n_dx_v) * ICX(n_dx_c) * ICY(n_dy_c) + IXZ(n_dx_w) * ICX(n_dx_c) * ICZ(n_dz_c);
                    FT cy = IXY(n_dy_u) * ICY(n_dy_c) * ICX(n_dx_c) + vy * ICY(n_dy_c2) + IXZ(n_dy_w) * ICY(n_dy_c) * ICZ(n_dz_c);
                    FT cz = IXZ(n_dz_u) * ICZ(n_dz_c) * ICX(n_dx_c) + IYZ(n_dz_v) * ICZ(n_dz_c) * ICY(n_dy_c) + wz * ICZ(n_dz_c2);
                    FT theta = cx + cy + cz;
                    kap_ = -g->Ckappa[m][t] * amd_delta2(g, i, j, k) * theta / sigma;
                }
                *ref(&g->kappae[m][t], i, j, k) = FMAX((FT)0, kap_);
            }
}

/* thread control for bench.py's CPU arm: torchrun exports OMP_NUM_THREADS=1, so the count is set explicitly and the number
 * actually in effect is what gets reported */
#if !defined(ORACLE_F32) && !defined(ORACLE_F80)
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int orc_get_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
#endif
