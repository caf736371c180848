/*
 * me_oracle.c — scalar CPU restatement of the MetropolisEngine hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Nothing in the product (metropolisengine_b200/) links, loads or calls this file.  It is the checker that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg compare the CUDA path against.
 *
 * Parity pinned: tests/test_oracle_c.py replays every fixture under tests/golden/ (recorded from the
 * unmodified reference by tests/golden/make_golden.py) through meo_run() in injected-draw mode and requires
 * identical accept/reject decisions and all state within 1e-13 relative of the reference.
 *
 * What is restated (reference = /root/reference/metropolisengine/metropolis_engine.py, "ME"; prose spec in
 * SURVEY.md Appendix A):
 *   step        ME:241-259 (mixed), ME:225-239 (all-real), ME:209-223 (all-complex)
 *   decision    ME:319-338     ties accept; T==0 rejects uphill without drawing; u <= exp(-1*diff/T)
 *   sigma       ME:429-456     f = max(n_measure/m, 200); c = sigma*ratio; +c(1-p)/f or -c p/f
 *   measure     ME:342-427     running means, Haario covariance + sigma^2/n regulariser once n > 50
 *   observables ME:458-463     |x_i|, |c_j|, x_i^2
 * Two draw sources:
 *   MEO_INJECT  recorded increments delta[step][d] and uniforms u[step] (NaN = none drawn)  — parity level L-A
 *   MEO_PHILOX  Philox4x32-7 + Box-Muller + per-chain Cholesky factors — the SAME stream definition the CUDA
 *               kernels use (key = seed, counter = (chain_lo, chain_hi, step, slot)), so a CUDA ensemble can be
 *               checked chain by chain.  (This part has no reference counterpart: the reference draws from
 *               numpy's global MT19937, ME:268,300.)
 *
 * State layout (one column of doubles per chain, offsets from meo_layout()), d = n_r + 2 n_c,
 * parameter order [real..., Re c..., Im c...] (the reference's embedding order, ME:288):
 *   X d | E 1 | SIG 2 (sigma_r, sigma_c) | MEAN d | COVR n_r(n_r+1)/2 lower-packed | COVC n_c^2 |
 *   OBSM 2n_r+n_c | FACR n_r(n_r+1)/2 | FACC n_c^2 | NACC 1 | STATUS 1
 * Hermitian packing (COVC, FACC): strictly-lower pairs p = i(i-1)/2 + j (j < i) at words [2p]=Re, [2p+1]=Im,
 * then the n_c real diagonal entries at offset n_c(n_c-1).
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared  (no FMA contraction: IEEE operation order matters).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MEO_MAX_D 160
#define MEO_INJECT 0
#define MEO_PHILOX 1
#define MEO_XOSHIRO 2        /* an unrelated generator (xoshiro256++, sequential per chain) feeding the same Box-Muller and
                                proposal code: shows which properties of a run belong to the algorithm, not to Philox */

enum { MEO_E_X2 = 0, MEO_E_XY = 1, MEO_E_MIXED = 2, MEO_E_CYL = 3, MEO_E_CALLBACK = 100 };

typedef double (*meo_energy_cb)(const double *x, int n_r, int n_c);
typedef int (*meo_reject_cb)(const double *x, int n_r, int n_c);

typedef struct {
    int n_r, n_c;
    int energy_id;
    int use_reject;          /* cylinder: |a| >= 1 ; callback: reject_cb */
    double consts[16];
    double temp, target, ratio;
    int m;
    meo_energy_cb energy_cb;
    meo_reject_cb reject_cb;
    int frozen;              /* experiment switch (tests/scripts/pooled_variance_offset.py): 1 = no width adaptation and no
                                covariance recursion from now on (means / observable means still run) */
} meo_config;

typedef struct {
    int X, E, SIG, MEAN, COVR, COVC, OBSM, FACR, FACC, NACC, STATUS, WORDS;
} meo_offsets;

void meo_layout(int n_r, int n_c, meo_offsets *o) {
    int d = n_r + 2 * n_c, w = 0;
    o->X = w; w += d;
    o->E = w; w += 1;
    o->SIG = w; w += 2;
    o->MEAN = w; w += d;
    o->COVR = w; w += n_r * (n_r + 1) / 2;
    o->COVC = w; w += n_c * n_c;
    o->OBSM = w; w += 2 * n_r + n_c;
    o->FACR = w; w += n_r * (n_r + 1) / 2;
    o->FACC = w; w += n_c * n_c;
    o->NACC = w; w += 1;
    o->STATUS = w; w += 1;
    o->WORDS = w;
}

/* ------------------------------------------------------------------ energies (oracle/energies.py order) */
static double energy_eval(const meo_config *c, const double *x) {
    const int n_r = c->n_r, n_c = c->n_c;
    const double *k = c->consts;
    switch (c->energy_id) {
    case MEO_E_X2:
        return x[0] * x[0];
    case MEO_E_XY:
        return k[0] * (x[0] * x[0] + x[1] * x[1]);
    case MEO_E_MIXED: {
        double area = 0.0, s = 0.0;
        for (int i = 0; i < n_r; i++) { double e = 1.0 - x[i]; area = area + k[0] * (e * e); }
        for (int j = 0; j < n_c; j++) {
            double re = x[n_r + j], im = x[n_r + n_c + j];
            double a = re * re + im * im;
            s = s + (k[1] * a + k[2] * (a * a));
        }
        double w = x[0] * x[1];
        return area + (k[3] != 0.0 ? fabs(w) : w) * (s / (double)n_c);
    }
    case MEO_E_CYL: {
        double a = x[0], a2 = a * a, quad = 0.0, tot = 0.0;
        for (int j = 0; j < n_c; j++) {
            double kk = (double)(j - n_c / 2);
            double re = x[n_r + j], im = x[n_r + n_c + j];
            double m2 = re * re + im * im;
            quad = quad + (k[1] + (k[2] * (kk * kk)) * (1.0 + a2)) * m2;
            tot = tot + m2;
        }
        return (k[0] * a2 + quad) + (k[3] / (2.0 * (double)n_c)) * (tot * tot);
    }
    default:
        return c->energy_cb(x, n_r, n_c);
    }
}

static int reject_eval(const meo_config *c, const double *x) {
    if (!c->use_reject) return 0;
    if (c->energy_id == MEO_E_CYL) return fabs(x[0]) >= 1.0;
    if (c->reject_cb) return c->reject_cb(x, c->n_r, c->n_c);
    return 0;
}

/* ------------------------------------------------------------------ Philox4x32-7 + Gaussian pairs */
#define MEO_PHILOX_ROUNDS 7      /* the stream definition of the CUDA kernels (me_device.cuh, ME_PHILOX_ROUNDS) */

static void philox4x32_r(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, int rounds,
                         uint32_t out[4]) {
    for (int r = 0; r < rounds; r++) {
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

void meo_philox(uint64_t seed, uint64_t chain, uint32_t step, uint32_t slot, uint32_t out[4]) {
    philox4x32_r((uint32_t)chain, (uint32_t)(chain >> 32), step, slot, (uint32_t)seed, (uint32_t)(seed >> 32),
                 MEO_PHILOX_ROUNDS, out);
}

/* any round count, for the Random123 known-answer vectors (7 and 10 rounds) */
void meo_philox_rounds(uint64_t seed, uint64_t chain, uint32_t step, uint32_t slot, int rounds, uint32_t out[4]) {
    philox4x32_r((uint32_t)chain, (uint32_t)(chain >> 32), step, slot, (uint32_t)seed, (uint32_t)(seed >> 32), rounds, out);
}

/* sin(pi t), cos(pi t) for t in [0,2): exact octant reduction, then libm on |arg| <= pi/4 */
static void sincospi_d(double t, double *s, double *c) {
    double q = floor(2.0 * t + 0.5);           /* nearest multiple of 1/2 (half-way cases up) */
    double r = t - 0.5 * q;                    /* exact, |r| <= 1/4 */
    double sr = sin(M_PI * r), cr = cos(M_PI * r);
    switch (((int)q) & 3) {
    case 0: *s = sr; *c = cr; break;
    case 1: *s = cr; *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
    }
}

/* Stream definition (shared with the CUDA kernels, me_device.cuh):
 *   per step and chain, Philox call q = 0 .. ceil(D/2)-1 with counter (chain_lo, chain_hi, step, q), key = seed,
 *   output words (x, y, z, w):
 *     radius uniform  u1 = (K + 1/2) 2^-52          in (0,1),  K = y : x[31:12]   (52 bits)
 *     angle           t  = z * 2^-31                in [0,2)      (sin/cos of pi*t)
 *     normals         z_{2q} = sqrt(-2 ln u1) cos(pi t),  z_{2q+1} = sqrt(-2 ln u1) sin(pi t)
 *   accept uniform    u = (A + 1/2) 2^-44 in (0,1),  A = w_0 : x_0[11:0]  (44 bits of call 0 the normals do not use).
 *   All bits used are distinct output bits of the generator. */
void meo_normal_pair(uint64_t seed, uint64_t chain, uint32_t step, uint32_t slot, double *z0, double *z1) {
    uint32_t r[4];
    meo_philox(seed, chain, step, slot, r);
    uint64_t K = ((uint64_t)r[1] << 20) | (uint64_t)(r[0] >> 12);
    double u1 = ((double)K + 0.5) * (1.0 / 4503599627370496.0);  /* exact: 2K+1 < 2^53 */
    double t = (double)r[2] * (1.0 / 2147483648.0);             /* [0,2) */
    double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi_d(t, &s, &c);
    *z0 = rad * c;
    *z1 = rad * s;
}

/* One-parameter shapes (d = 1): ONE call serves the steps 2P and 2P + 1 (me_device.cuh, ShareCall): counter
 * (chain_lo, chain_hi, P, 0); radius uniform from y alone, K = y : 0x80000; angle t = z 2^-31; step 2P proposes with the
 * cosine normal, step 2P + 1 with the sine normal; accept uniform A = w : 0x800 (step 2P), A = x : 0x800 (step 2P + 1). */
double meo_shared_normal(uint64_t seed, uint64_t chain, uint32_t step) {
    uint32_t r[4];
    meo_philox(seed, chain, step >> 1, 0, r);
    uint64_t K = ((uint64_t)r[1] << 20) | (uint64_t)0x80000u;
    double u1 = ((double)K + 0.5) * (1.0 / 4503599627370496.0);
    double t = (double)r[2] * (1.0 / 2147483648.0);
    double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi_d(t, &s, &c);
    return (step & 1u) ? rad * s : rad * c;
}

double meo_shared_uniform(uint64_t seed, uint64_t chain, uint32_t step) {
    uint32_t r[4];
    meo_philox(seed, chain, step >> 1, 0, r);
    uint64_t bits = ((uint64_t)((step & 1u) ? r[0] : r[3]) << 12) | (uint64_t)0x800u;
    return ((double)bits + 0.5) * (1.0 / 17592186044416.0);     /* 2^-44 */
}

double meo_uniform(uint64_t seed, uint64_t chain, uint32_t step, int n_calls) {
    uint32_t r[4];
    (void)n_calls;
    meo_philox(seed, chain, step, 0, r);
    uint64_t bits = ((uint64_t)r[3] << 12) | (uint64_t)(r[0] & 0xfffu);
    return ((double)bits + 0.5) * (1.0 / 17592186044416.0);     /* 2^-44 */
}

/* xoshiro256++ (Blackman & Vigna), state seeded by splitmix64 of (seed, chain); one stream per chain, kept in a small
 * table keyed by the caller (the experiment driver runs one chain per object, sequentially) */
static uint64_t xo_s[4];
static uint64_t xo_rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
void meo_xoshiro_seed(uint64_t seed, uint64_t chain) {
    uint64_t z = seed * 0x9E3779B97F4A7C15ull + chain * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull;
    for (int i = 0; i < 4; i++) {
        z += 0x9E3779B97F4A7C15ull;
        uint64_t t = z;
        t = (t ^ (t >> 30)) * 0xBF58476D1CE4E5B9ull;
        t = (t ^ (t >> 27)) * 0x94D049BB133111EBull;
        xo_s[i] = t ^ (t >> 31);
    }
}
void meo_xoshiro_state(uint64_t *io, int set) { for (int i = 0; i < 4; i++) { if (set) xo_s[i] = io[i]; else io[i] = xo_s[i]; } }
static uint64_t xo_next(void) {
    const uint64_t r = xo_rotl(xo_s[0] + xo_s[3], 23) + xo_s[0];
    const uint64_t t = xo_s[1] << 17;
    xo_s[2] ^= xo_s[0]; xo_s[3] ^= xo_s[1]; xo_s[1] ^= xo_s[2]; xo_s[0] ^= xo_s[3];
    xo_s[2] ^= t; xo_s[3] = xo_rotl(xo_s[3], 45);
    return r;
}
static double xo_uniform(void) { return ((double)(xo_next() >> 11) + 0.5) * (1.0 / 9007199254740992.0); }
static void xo_normal_pair(double *z0, double *z1) {
    const double u1 = xo_uniform(), t = 2.0 * xo_uniform();
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi_d(t, &s, &c);
    *z0 = rad * c; *z1 = rad * s;
}

/* ------------------------------------------------------------------ Cholesky factors of the proposal covariances */
static int herm_lo(int i, int j) { return 2 * (i * (i - 1) / 2 + j); }   /* j < i */

/* returns 0 ok, 1 if a pivot was not positive (reference: numpy raises / warns, ME:270) */
int meo_refactor(int n_r, int n_c, double *st, const meo_offsets *o) {
    int bad = 0;
    double *C = st + o->COVR, *L = st + o->FACR;
    for (int i = 0; i < n_r; i++) {
        for (int j = 0; j <= i; j++) {
            double s = C[i * (i + 1) / 2 + j];
            for (int k = 0; k < j; k++) s -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
            if (i == j) {
                if (!(s > 0.0)) { bad = 1; s = 0.0; }
                L[i * (i + 1) / 2 + i] = sqrt(s);
            } else {
                double piv = L[j * (j + 1) / 2 + j];
                L[i * (i + 1) / 2 + j] = piv > 0.0 ? s / piv : 0.0;
            }
        }
    }
    double *H = st + o->COVC, *G = st + o->FACC;
    int dg = n_c * (n_c - 1);
    for (int i = 0; i < n_c; i++) {
        for (int j = 0; j <= i; j++) {
            if (i == j) {
                double s = H[dg + i];
                for (int k = 0; k < j; k++) {
                    double re = G[herm_lo(i, k)], im = G[herm_lo(i, k) + 1];
                    s -= re * re + im * im;
                }
                if (!(s > 0.0)) { bad = 1; s = 0.0; }
                G[dg + i] = sqrt(s);
            } else {
                double sre = H[herm_lo(i, j)], sim = H[herm_lo(i, j) + 1];
                for (int k = 0; k < j; k++) {   /* s -= L_ik * conj(L_jk) */
                    double are = G[herm_lo(i, k)], aim = G[herm_lo(i, k) + 1];
                    double bre = G[herm_lo(j, k)], bim = G[herm_lo(j, k) + 1];
                    sre -= are * bre + aim * bim;
                    sim -= aim * bre - are * bim;
                }
                double piv = G[dg + j];
                G[herm_lo(i, j)] = piv > 0.0 ? sre / piv : 0.0;
                G[herm_lo(i, j) + 1] = piv > 0.0 ? sim / piv : 0.0;
            }
        }
    }
    return bad;
}

/* ------------------------------------------------------------------ initialisation (ME:40-125) */
void meo_init(const meo_config *c, double *st, const double *x0, double sigma0, const double *cov_r /*n_r*n_r or NULL*/,
              const double *cov_c_re, const double *cov_c_im /*n_c*n_c or NULL*/) {
    meo_offsets o;
    const int n_r = c->n_r, n_c = c->n_c, d = n_r + 2 * n_c;
    meo_layout(n_r, n_c, &o);
    memset(st, 0, sizeof(double) * o.WORDS);
    for (int i = 0; i < d; i++) { st[o.X + i] = x0[i]; st[o.MEAN + i] = x0[i]; }
    st[o.SIG] = sigma0; st[o.SIG + 1] = sigma0;
    for (int i = 0; i < n_r; i++)
        for (int j = 0; j <= i; j++)
            st[o.COVR + i * (i + 1) / 2 + j] = cov_r ? cov_r[i * n_r + j] : (i == j ? 1.0 : 0.0);
    for (int i = 0; i < n_c; i++) {
        st[o.COVC + n_c * (n_c - 1) + i] = cov_c_re ? cov_c_re[i * n_c + i] : 1.0;
        for (int j = 0; j < i; j++) {
            st[o.COVC + herm_lo(i, j)] = cov_c_re ? cov_c_re[i * n_c + j] : 0.0;
            st[o.COVC + herm_lo(i, j) + 1] = cov_c_im ? cov_c_im[i * n_c + j] : 0.0;
        }
    }
    for (int i = 0; i < n_r; i++) st[o.OBSM + i] = fabs(x0[i]);
    for (int j = 0; j < n_c; j++) st[o.OBSM + n_r + j] = hypot(x0[n_r + j], x0[n_r + n_c + j]);
    for (int i = 0; i < n_r; i++) st[o.OBSM + n_r + n_c + i] = x0[i] * x0[i];
    st[o.E] = energy_eval(c, x0);
    if (meo_refactor(n_r, n_c, st, &o)) st[o.STATUS] = 1.0;
}

/* ------------------------------------------------------------------ measure (ME:342-427) */
static void measure(const meo_config *c, double *st, const meo_offsets *o, int64_t n) {
    const int n_r = c->n_r, n_c = c->n_c;
    const double dn = (double)n, dn1 = (double)(n - 1), dn2 = (double)(n - 2);
    double *x = st + o->X, *mean = st + o->MEAN;
    double old[MEO_MAX_D];
    const double shrink = dn1 / dn;
    if (n_r) {
        for (int i = 0; i < n_r; i++) old[i] = mean[i];
        for (int i = 0; i < n_r; i++) { mean[i] = mean[i] * shrink; mean[i] = mean[i] + x[i] / dn; }
        if (n > 50 && !c->frozen) {
            const double sig = st[o->SIG];
            const double small = (sig * sig) / dn;
            const double decay = dn2 / dn1, grow = dn / dn1;
            double *C = st + o->COVR;
            for (int i = 0; i < n_r; i++)
                for (int j = 0; j <= i; j++) {
                    double v = C[i * (i + 1) / 2 + j] * decay;
                    double add = ((old[i] * old[j] - grow * (mean[i] * mean[j])) + (x[i] * x[j]) / dn1)
                                 + (i == j ? small : 0.0);
                    C[i * (i + 1) / 2 + j] = v + add;
                }
        }
    }
    if (n_c) {
        /* numpy divides a complex array by a real scalar as multiplication by the reciprocal */
        const double inv_n = 1.0 / dn, inv_n1 = 1.0 / dn1;
        double *xr = x + n_r, *xi = x + n_r + n_c, *mr = mean + n_r, *mi = mean + n_r + n_c;
        double *or_ = old, *oi = old + n_c;
        for (int j = 0; j < n_c; j++) { or_[j] = mr[j]; oi[j] = mi[j]; }
        for (int j = 0; j < n_c; j++) {
            mr[j] = mr[j] * shrink; mi[j] = mi[j] * shrink;
            mr[j] = mr[j] + xr[j] * inv_n; mi[j] = mi[j] + xi[j] * inv_n;
        }
        if (n > 50 && !c->frozen) {
            const double sig = st[o->SIG + 1];
            const double small = (sig * sig) / dn;
            const double decay = dn2 / dn1, grow = dn / dn1;
            double *H = st + o->COVC;
            const int dg = n_c * (n_c - 1);
            for (int i = 0; i < n_c; i++)
                for (int j = 0; j <= i; j++) {
                    /* outer(a, conj b)[i][j] = a_i * conj(b_j): re = ar*br + ai*bi ; im = ai*br - ar*bi */
                    double o_re = or_[i] * or_[j] + oi[i] * oi[j], o_im = oi[i] * or_[j] - or_[i] * oi[j];
                    double m_re = mr[i] * mr[j] + mi[i] * mi[j], m_im = mi[i] * mr[j] - mr[i] * mi[j];
                    double x_re = xr[i] * xr[j] + xi[i] * xi[j], x_im = xi[i] * xr[j] - xr[i] * xi[j];
                    double a_re = ((o_re - grow * m_re) + x_re * inv_n1) + (i == j ? small : 0.0);
                    double a_im = ((o_im - grow * m_im) + x_im * inv_n1);
                    if (i == j) {
                        H[dg + i] = H[dg + i] * decay + a_re;
                    } else {
                        H[herm_lo(i, j)] = H[herm_lo(i, j)] * decay + a_re;
                        H[herm_lo(i, j) + 1] = H[herm_lo(i, j) + 1] * decay + a_im;
                    }
                }
        }
    }
    double *om = st + o->OBSM;
    for (int i = 0; i < n_r; i++) om[i] = om[i] * shrink + fabs(x[i]) / dn;
    for (int j = 0; j < n_c; j++) om[n_r + j] = om[n_r + j] * shrink + hypot(x[n_r + j], x[n_r + n_c + j]) / dn;
    for (int i = 0; i < n_r; i++) om[n_r + n_c + i] = om[n_r + n_c + i] * shrink + (x[i] * x[i]) / dn;
    if (n > 50 && !c->frozen && meo_refactor(n_r, n_c, st, o)) st[o->STATUS] = 1.0;
}

/* ------------------------------------------------------------------ the driver
 * Runs n_blocks x (spm steps [+ one measure if do_measure]) on ONE chain.
 *   mode MEO_INJECT: delta[(step)*d + k], u[step]
 *   mode MEO_PHILOX: seed, chain_id, step0 (global index of the first step)
 * n_measure: in/out measure_step_counter (starts at 1, ME:73).
 * accept_out[step] (may be NULL); ts_out rows of (d + 3): x[d], E, sigma_r, sigma_c (may be NULL).
 */
int meo_run_group(const meo_config *c, double *st, int mode, int64_t n_blocks, int64_t spm, int do_measure,
                  int64_t *n_measure, const double *delta, const double *u, uint64_t seed, uint64_t chain_id,
                  uint64_t step0, unsigned char *accept_out, double *ts_out, int group);

int meo_run(const meo_config *c, double *st, int mode, int64_t n_blocks, int64_t spm, int do_measure,
            int64_t *n_measure, const double *delta, const double *u, uint64_t seed, uint64_t chain_id,
            uint64_t step0, unsigned char *accept_out, double *ts_out) {
    return meo_run_group(c, st, mode, n_blocks, spm, do_measure, n_measure, delta, u, seed, chain_id, step0,
                         accept_out, ts_out, 0);
}

/* group: 0 = step_all; for mixed engines 1 = step_real_group (ME:225-239), 2 = step_complex_group (ME:209-223):
 * only that block is proposed and only that group's width adapts (ME:440-456).
 * 3 / 4 = the two halves of step_complex_group under complex_sample_method="magnitude-phase" (ME:168-207):
 *   3 magnitude: c'_j = rect(N(|c_j|, s_j), arg c_j), s_j = sigma_c^2 C_jj used as a standard deviation (ME:304-310),
 *     adapts the complex width like group 2;
 *   4 phase: c'_j = rect(|c_j|, U(-pi, pi)) (ME:312-317), never adapts a width (ME:194-207).
 *   Injected mode: the records of these two hold the ABSOLUTE proposed values.
 *   Philox mode: magnitude j uses the step's normal z_j; phase j uses the angle word of Philox call j,
 *     theta = pi (word 2^-31) - pi. */
int meo_run_group(const meo_config *c, double *st, int mode, int64_t n_blocks, int64_t spm, int do_measure,
                  int64_t *n_measure, const double *delta, const double *u, uint64_t seed, uint64_t chain_id,
                  uint64_t step0, unsigned char *accept_out, double *ts_out, int group) {
    meo_offsets o;
    const int n_r = c->n_r, n_c = c->n_c, d = n_r + 2 * n_c;
    if (d > MEO_MAX_D) return -1;
    meo_layout(n_r, n_c, &o);
    const int kind = (n_r && n_c) ? 0 : (n_r ? 1 : 2);      /* 0 mixed, 1 all-real, 2 all-complex */
    double *x = st + o.X;
    int64_t n = *n_measure;
    int64_t s = 0;
    const double p = c->target;
    for (int64_t b = 0; b < n_blocks; b++) {
        for (int64_t k = 0; k < spm; k++, s++) {
            double prop[MEO_MAX_D];
            const double sr = st[o.SIG], sc = st[o.SIG + 1];
            if (mode == MEO_INJECT) {
                for (int i = 0; i < d; i++) prop[i] = delta[s * d + i] + (group >= 3 ? 0.0 : x[i]);
            } else if (group >= 3) {
                const uint32_t step = (uint32_t)(step0 + (uint64_t)s);
                const double *G = st + o.FACC;
                const int dg = n_c * (n_c - 1);
                for (int i = 0; i < n_r; i++) prop[i] = x[i];
                for (int j = 0; j < n_c; j++) {
                    const double re = x[n_r + j], im = x[n_r + n_c + j];
                    const double mag = hypot(re, im), ph = atan2(im, re);
                    if (group == 3) {
                        double z[2], cjj = G[dg + j] * G[dg + j];      /* C_jj = sum_k |G_jk|^2 */
                        for (int k = 0; k < j; k++)
                            cjj += G[herm_lo(j, k)] * G[herm_lo(j, k)] + G[herm_lo(j, k) + 1] * G[herm_lo(j, k) + 1];
                        meo_normal_pair(seed, chain_id, step, (uint32_t)(j / 2), &z[0], &z[1]);
                        const double nm = mag + z[j & 1] * ((sc * sc) * cjj);
                        prop[n_r + j] = nm * cos(ph);
                        prop[n_r + n_c + j] = nm * sin(ph);
                    } else {
                        uint32_t r[4];
                        meo_philox(seed, chain_id, step, (uint32_t)j, r);
                        double sn, cs;
                        sincospi_d((double)r[2] * (1.0 / 2147483648.0), &sn, &cs);   /* theta = pi t - pi */
                        prop[n_r + j] = mag * -cs;
                        prop[n_r + n_c + j] = mag * -sn;
                    }
                }
            } else {
                double z[MEO_MAX_D + 1];
                const uint32_t step = (uint32_t)(step0 + (uint64_t)s);
                for (int q = 0; q < (d + 1) / 2; q++) {
                    if (mode == MEO_XOSHIRO) xo_normal_pair(&z[2 * q], &z[2 * q + 1]);
                    else if (d == 1) { z[0] = meo_shared_normal(seed, chain_id, step); z[1] = 0.0; }
                    else meo_normal_pair(seed, chain_id, step, (uint32_t)q, &z[2 * q], &z[2 * q + 1]);
                }
                /* real block: sigma_r * (L z) ; z[0..n_r) */
                const double *L = st + o.FACR;
                for (int i = 0; i < n_r; i++) {
                    double acc = 0.0;
                    for (int j = 0; j <= i; j++) acc = acc + L[i * (i + 1) / 2 + j] * z[j];
                    prop[i] = x[i] + sr * acc;
                }
                /* complex block: sigma_c * conj(L) xi, xi_j = (z[n_r+2j] + i z[n_r+2j+1]) / sqrt(2) */
                const double *G = st + o.FACC;
                const int dg = n_c * (n_c - 1);
                const double rs = 0.70710678118654752440;
                for (int i = 0; i < n_c; i++) {
                    double are = 0.0, aim = 0.0;
                    for (int j = 0; j < i; j++) {
                        double lre = G[herm_lo(i, j)], lim = -G[herm_lo(i, j) + 1];     /* conj */
                        double zre = z[n_r + 2 * j], zim = z[n_r + 2 * j + 1];
                        are = are + (lre * zre - lim * zim);
                        aim = aim + (lre * zim + lim * zre);
                    }
                    are = are + G[dg + i] * z[n_r + 2 * i];
                    aim = aim + G[dg + i] * z[n_r + 2 * i + 1];
                    prop[n_r + i] = x[n_r + i] + (sc * rs) * are;
                    prop[n_r + n_c + i] = x[n_r + n_c + i] + (sc * rs) * aim;
                }
            }
            if (kind == 0 && group != 0)
                for (int i = 0; i < d; i++)
                    if ((i < n_r) != (group == 1)) prop[i] = x[i];
            int accept = 0;
            if (!reject_eval(c, prop)) {
                double e_new = energy_eval(c, prop);
                double diff = e_new - st[o.E];
                if (diff <= 0) accept = 1;
                else if (c->temp == 0) accept = 0;
                else {
                    double uu = (mode == MEO_INJECT) ? u[s] : (mode == MEO_XOSHIRO) ? xo_uniform()
                              : (d == 1) ? meo_shared_uniform(seed, chain_id, (uint32_t)(step0 + (uint64_t)s))
                              : meo_uniform(seed, chain_id, (uint32_t)(step0 + (uint64_t)s), (d + 1) / 2);
                    accept = uu <= exp(-1 * diff / c->temp);
                }
                if (accept) { st[o.E] = e_new; for (int i = 0; i < d; i++) x[i] = prop[i]; st[o.NACC] += 1.0; }
            }
            /* sigma adaptation (ME:429-456) */
            {
                double f = (double)n / (double)c->m; if (!(f > 200.0)) f = 200.0;
                const int grouped = (kind == 0 && group != 0);
                double *sg = grouped ? &st[o.SIG + (group == 1 ? 0 : 1)] : ((kind == 2) ? &st[o.SIG + 1] : &st[o.SIG]);
                double cc = (*sg) * c->ratio;
                if (group == 4 || c->frozen) { /* phase redraw: no adaptation; frozen: experiment switch */ }
                else if (accept) *sg = *sg + (cc * (1 - p)) / f;
                else *sg = *sg - (cc * p) / f;
                if (kind == 0 && !grouped) { st[o.SIG + 1] = st[o.SIG]; if (!(st[o.SIG] > 0)) st[o.STATUS] = 2.0; }
            }
            if (accept_out) accept_out[s] = (unsigned char)accept;
        }
        if (do_measure) {
            n += 1;
            measure(c, st, &o, n);
            if (ts_out) {
                double *row = ts_out + b * (d + 3);
                for (int i = 0; i < d; i++) row[i] = x[i];
                row[d] = st[o.E]; row[d + 1] = st[o.SIG]; row[d + 2] = st[o.SIG + 1];
            }
        }
    }
    *n_measure = n;
    return 0;
}
