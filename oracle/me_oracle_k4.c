/*
 * me_oracle_k4.c — scalar CPU restatement of the shared-covariance step (csrc/me_k4_device.cuh).  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (metropolisengine_b200/) links, loads or calls this file; tests/ compare the CUDA path with it.
 *
 * The reference (metropolisengine/metropolis_engine.py, "ME") has no shared-covariance mode — it is one chain — so
 * this path has no reference-held vectors: PARITY IS PINNED ON THE REFERENCE'S STEP SEMANTICS (proposal law ME:274-302,
 * hard wall ME:247, decision ME:319-338, Robbins-Monro width ME:429-438, running means ME:404-414) restated here with
 * the pooled covariance in place of the per-chain one, and on the per-chain engine (itself pinned by goldens recorded
 * from the live reference) through the ensemble cross-check in tests/test_gpu_k4.py.
 *
 * What is restated, in the kernel's own operation order (explicit fma where the kernel uses fma, so an FP64 chain is
 * reproduced bit for bit once the tensor-core increments are given):
 *   stream      chain g, step s: normals 8c..8c+7 of the operand row from Philox4x32-7(counter (g_lo, g_hi, s, c)), two per
 *               output word by inverse CDF from a 4096-entry BF16 quantile table (12 index bits + sign per half-word);
 *               scalar draws from slot 0x10000: low half of word x -> the real parameter's normal, words z, w -> accept
 *               uniform (53 bits).  Integer arithmetic and table look-ups only: the operand is reproduced EXACTLY
 *   increments  Delta_n = sum_k bf16(Z_k) bf16(B_nk), accumulated in double and rounded to float (the tensor core
 *               accumulates in FP32 in an order of its own: compared with a tolerance, then INJECTED — level L-A of
 *               SURVEY §8c applied to this path)
 *   step        x'_n = fma(sigma, (double)Delta_n, x_n); mode sums per group of consecutive modes (2 groups: p0 + p1;
 *               4 groups: (p0 + p1) + (p2 + p3) — the kernel's build says which, me_k4_layout.SUM_GROUPS); a' = fma(sigma s_a, za, a);
 *               wall; E' = total(a', s0, s1); accept = dE <= 0 or (T != 0 and u <= exp(-dE / T)); sigma update by fma
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define K4O_MAX_NC 64

typedef struct {
    int nc;
    int use_wall;
    double consts[16];        /* cylinder functor: kappa, alpha, gamma, beta */
    double temp, target, ratio;
    int groups;               /* groups of consecutive modes the per-mode sums are accumulated over (me_k4_layout.SUM_GROUPS) */
} k4o_config;

typedef struct { int D, X, E, SIG, MEAN, OBSM, NOBS, NACC, STATUS, WORDS; } k4o_layout;

void k4o_layout_for(int nc, k4o_layout *L) {
    L->D = 1 + 2 * nc; L->X = 0; L->E = L->D; L->SIG = L->D + 1; L->MEAN = L->D + 2; L->OBSM = L->MEAN + L->D;
    L->NOBS = 2 + nc; L->NACC = L->OBSM + L->NOBS; L->STATUS = L->NACC + 1; L->WORDS = L->STATUS + 1;
}

/* ------------------------------------------------------------------ Philox4x32-7 */
static void philox7(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    for (int r = 0; r < 7; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static float as_float(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t as_u32(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

/* float -> BF16 (round to nearest even), returned as the float it represents */
float k4o_bf16(float v) {
    uint32_t b = as_u32(v);
    uint32_t r = b + 0x7fffu + ((b >> 16) & 1u);
    return as_float(r & 0xffff0000u);
}
/* double -> BF16 in one rounding (nearest even) */
float k4o_bf16_from_double(double v) {
    float f = (float)v;                       /* nearest float */
    /* repair double rounding: if f sits exactly on a BF16 tie but v does not, move f off the tie towards v */
    uint32_t b = as_u32(f);
    if ((b & 0xffffu) == 0x8000u && (double)f != v) {
        b += ((double)f < v) == !(b >> 31) ? 1u : (uint32_t)-1;
        f = as_float(b);
    }
    return k4o_bf16(f);
}

/* Two BF16 normals from one 32-bit random word (k4::normal_pair_bf16): the low and the high half each give one normal —
 * 12 bits index the table of the 4096 half-normal quantiles Phi^-1(1/2 + (i + 1/2) / 8192) (BF16 bit patterns), bit 15 of
 * the half is the sign.  The table is an INPUT here: the test builds it independently (scipy's inverse normal CDF) and
 * checks it against the library's before use. */
static void normal_pair(uint32_t w, const uint16_t *tab, float *z0, float *z1) {
    const uint32_t t0 = tab[w & 0xfffu], t1 = tab[(w >> 16) & 0xfffu];
    *z0 = as_float((t0 | (w & 0x8000u)) << 16);
    *z1 = as_float(((t1 << 16) | (w & 0x80000000u)));
}

/* normals of one chain and step: zb[K] (K = 2 n_c), the BF16 values the operand holds */
void k4o_normals(uint64_t seed, uint64_t chain, uint32_t step, int K, const uint16_t *tab, float *zb) {
    for (int c = 0; c < K / 8; c++) {
        uint32_t r[4];
        philox7((uint32_t)chain, (uint32_t)(chain >> 32), step, (uint32_t)c, (uint32_t)seed, (uint32_t)(seed >> 32), r);
        for (int w = 0; w < 4; w++) normal_pair(r[w], tab, &zb[8 * c + 2 * w], &zb[8 * c + 2 * w + 1]);
    }
}

/* scalar draws of one chain and step: the real parameter's normal (table, low half of word x) and the accept uniform */
void k4o_scalars(uint64_t seed, uint64_t chain, uint32_t step, const uint16_t *tab, double *za, double *u) {
    uint32_t r[4];
    philox7((uint32_t)chain, (uint32_t)(chain >> 32), step, 0x10000u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
    float a, b;
    normal_pair(r[0], tab, &a, &b);
    *za = (double)a;
    /* u53: (hi >> 5) 2^-27 + (lo >> 6) 2^-53, a 53-bit uniform in [0, 1) (both terms and their sum are exact) */
    *u = (double)(r[2] >> 5) * (1.0 / 134217728.0) + (double)(r[3] >> 6) * (1.0 / 9007199254740992.0);
}

/* Delta[n] = sum_k B[n][k] z[k] for n < N (B row-major [N][K], BF16 values as floats), double accumulation -> float;
 * scale[n] = sum_k |B[n][k] z[k]| (for the FP32-accumulation tolerance) */
void k4o_delta(int N, const float *B, const float *zb, float *delta, double *scale) {
    for (int n = 0; n < N; n++) {
        double acc = 0.0, sc = 0.0;
        for (int k = 0; k < N; k++) {
            const double t = (double)B[(long)n * N + k] * (double)zb[k];
            acc += t; sc += fabs(t);
        }
        delta[n] = (float)acc;
        if (scale) scale[n] = sc;
    }
}

/* cylinder functor, the kernel's sequence (k4::EnergyCylinder) */
static void mode(double q, double re, double im, double *s0, double *s1) {
    const double m2 = fma(re, re, im * im);
    *s0 += m2;
    *s1 = fma(q * q, m2, *s1);
}
static double total(double a, double s0, double s1, const double *k, int nc) {
    const double a2 = a * a;
    const double inner = fma(k[1], s0, (k[2] * (1.0 + a2)) * s1);
    const double quad = fma(k[0], a2, inner);
    return fma(k[3] / (2.0 * (double)nc), s0 * s0, quad);
}

/* e^x for x <= 0 as the kernel evaluates it (me::exp_nonpos: 64-entry 2^(j/64) table + degree-5 polynomial); the
 * result only enters the comparison u <= prob */
static double exp_nonpos(double x) {
    if (x < -700.0) x = -700.0;
    const double kd = fma(x, 0x1.71547p+6, 6755399441055744.0);
    int64_t kbits; memcpy(&kbits, &kd, 8);
    const int k = (int)(int32_t)(kbits & 0xffffffff);
    const double kf = kd - 6755399441055744.0;
    double r = fma(kf, -1.083042469326756e-02, x);
    r = fma(kf, -2.9815858269852933e-12, r);
    double p = fma(r, 0x1.11111p-7, 4.1666666666666664e-2);
    p = fma(r, p, 0.16666666666666666);
    p = fma(r, p, 0.5);
    p = fma(r * r, p, r);
    const double t = exp2((double)(k & 63) * 0.015625);
    const double v = fma(t, p, t);
    return ldexp(v, k >> 6);
}

/* initial state of one chain (k4::init_body): energy from one pass over the modes */
void k4o_init(const k4o_config *c, double *st, const double *x0, double sigma0) {
    k4o_layout L;
    k4o_layout_for(c->nc, &L);
    const int nc = c->nc;
    memset(st, 0, sizeof(double) * L.WORDS);
    double s0 = 0.0, s1 = 0.0;
    st[L.X] = x0[0]; st[L.MEAN] = x0[0];
    for (int j = 0; j < nc; j++) {
        const double re = x0[1 + j], im = x0[1 + nc + j];
        st[L.X + 1 + j] = re; st[L.X + 1 + nc + j] = im;
        st[L.MEAN + 1 + j] = re; st[L.MEAN + 1 + nc + j] = im;
        mode((double)(j - nc / 2), re, im, &s0, &s1);
        st[L.OBSM + 1 + j] = hypot(re, im);
    }
    st[L.OBSM] = fabs(x0[0]);
    st[L.OBSM + 1 + nc] = x0[0] * x0[0];
    st[L.E] = total(x0[0], s0, s1, c->consts, nc);
    st[L.SIG] = sigma0;
}

/* one step of one chain with INJECTED increments (interleaved Re, Im: delta[2j], delta[2j+1]), real-parameter normal
 * and accept uniform; n_meas = measure_step_counter.  Returns the accept decision. */
int k4o_step(const k4o_config *c, double *st, const float *delta, double za, double u, double s_a, int64_t n_meas) {
    k4o_layout L;
    k4o_layout_for(c->nc, &L);
    const int groups = c->groups == 4 ? 4 : 2;   /* the kernel's column groups per chain (k4::EPI_GROUPS) */
    const int nc = c->nc, quarter = nc / groups;
    const double sig = st[L.SIG];
    double f = (double)n_meas / (double)(1 + nc);
    if (!(f > 200.0)) f = 200.0;
    const double g_up = c->ratio * (1 - c->target) / f, g_down = c->ratio * c->target / f;
    double xr[K4O_MAX_NC], xi[K4O_MAX_NC], part0[4], part1[4];
    part0[2] = part0[3] = part1[2] = part1[3] = 0.0;
    for (int g = 0; g < groups; g++) {
        double s0 = 0.0, s1 = 0.0, q = (double)(g * quarter - nc / 2);
        for (int jj = 0; jj < quarter; jj++) {
            const int j = g * quarter + jj;
            xr[j] = fma(sig, (double)delta[2 * j], st[L.X + 1 + j]);
            xi[j] = fma(sig, (double)delta[2 * j + 1], st[L.X + 1 + nc + j]);
            mode(q, xr[j], xi[j], &s0, &s1);
            q += 1.0;
        }
        part0[g] = s0; part1[g] = s1;
    }
    const double t0 = groups == 4 ? (part0[0] + part0[1]) + (part0[2] + part0[3]) : part0[0] + part0[1];
    const double t1 = groups == 4 ? (part1[0] + part1[1]) + (part1[2] + part1[3]) : part1[0] + part1[1];
    const double a_new = fma(sig * s_a, za, st[L.X]);
    int accept = 0;
    const int wall = c->use_wall && fabs(a_new) >= 1.0;
    if (!wall) {
        const double e_new = total(a_new, t0, t1, c->consts, nc);
        const double diff = e_new - st[L.E];
        const double inv_temp = c->temp != 0 ? 1.0 / c->temp : 0.0;
        const double prob = exp_nonpos(fmin(-diff * inv_temp, 0.0));
        accept = (diff <= 0) || ((c->temp != 0) && (diff == diff) && (u <= prob));
        if (accept) {
            st[L.E] = e_new; st[L.X] = a_new; st[L.NACC] += 1.0;
            for (int j = 0; j < nc; j++) { st[L.X + 1 + j] = xr[j]; st[L.X + 1 + nc + j] = xi[j]; }
        }
    }
    st[L.SIG] = accept ? fma(sig, g_up, sig) : fma(sig, -g_down, sig);
    return accept;
}

/* measure of one chain (k4_measure): running means, observable means; n = counter AFTER the increment */
void k4o_measure(const k4o_config *c, double *st, int64_t n) {
    k4o_layout L;
    k4o_layout_for(c->nc, &L);
    const int nc = c->nc;
    const double dn = (double)n, inv_n = 1.0 / dn, shrink = (dn - 1.0) * inv_n;
    const double a = st[L.X];
    st[L.MEAN] = st[L.MEAN] * shrink + a * inv_n;
    st[L.OBSM] = st[L.OBSM] * shrink + fabs(a) * inv_n;
    st[L.OBSM + 1 + nc] = st[L.OBSM + 1 + nc] * shrink + (a * a) * inv_n;
    for (int j = 0; j < nc; j++) {
        const double re = st[L.X + 1 + j], im = st[L.X + 1 + nc + j];
        st[L.MEAN + 1 + j] = st[L.MEAN + 1 + j] * shrink + re * inv_n;
        st[L.MEAN + 1 + nc + j] = st[L.MEAN + 1 + nc + j] * shrink + im * inv_n;
        st[L.OBSM + 1 + j] = st[L.OBSM + 1 + j] * shrink + hypot(re, im) * inv_n;
    }
}
