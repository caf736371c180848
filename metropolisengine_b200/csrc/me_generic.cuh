/* me_generic.cuh — runtime-shape kernels for parameter spaces too large for the register-resident fused kernel
 * (D = n_r + 2 n_c > 32; e.g. the cylinder shape 1 real + 64 complex with PER-CHAIN adaptive covariance, the
 * reference's own algorithm).  One thread per chain, every per-chain quantity stays in the chain-minor state block
 * in global memory (each word access is coalesced across the warp).  The step is unfused:
 *     gk_propose -> gk_energy (built-in functor) or the caller's torch callable -> gk_accept,   gk_measure
 * Same arithmetic, operation order and Philox stream definition as me_device.cuh / oracle/me_oracle.c; compiled with
 * -fmad=false, so it is also the parity instantiation for these shapes (draw injection through inj_delta / inj_u).
 * Reference: proposal metropolis_engine.py:261-302, decision ME:319-338, sigma ME:429-456, measure ME:342-427.
 * The per-chain Cholesky (O(n_c^3) per measure, all operands in global memory) replaces numpy's per-STEP SVD.
 */
#ifndef ME_GENERIC_CUH
#define ME_GENERIC_CUH

#include "me_kernels.cuh"

#ifdef ME_NVRTC     /* the public header is not available to NVRTC: the functor ids of include/me_b200.h (me_generic.cu checks them) */
enum { ME_ENERGY_X2 = 0, ME_ENERGY_XY_WELL = 1, ME_ENERGY_MIXED_WELL = 2, ME_ENERGY_CYLINDER = 3, ME_ENERGY_EXTERNAL = 99,
       ME_ENERGY_USER = 100 };
#endif

namespace megk {


struct GLay {
    int nr, nc, d, X, E, SIG, MEAN, COVR, COVC, OBSM, FACR, FACC, NACC, STATUS;
    __device__ GLay(int nr_, int nc_) : nr(nr_), nc(nc_), d(nr_ + 2 * nc_) {
        X = 0; E = d; SIG = d + 1; MEAN = d + 3; COVR = MEAN + d; COVC = COVR + nr * (nr + 1) / 2;
        OBSM = COVC + nc * nc; FACR = OBSM + 2 * nr + nc; FACC = FACR + nr * (nr + 1) / 2; NACC = FACC + nc * nc;
        STATUS = NACC + 1;
    }
};
__device__ __forceinline__ int g_tri(int i, int j) { return i * (i + 1) / 2 + j; }
__device__ __forceinline__ int g_hlo(int i, int j) { return 2 * (i * (i - 1) / 2 + j); }

#define ST(w) st[(long long)(w) * ld + ch]

/* Cholesky factors of the chain's covariances, in place in global memory (me::refactor for runtime shapes).
   (Blocked variants with the same bits were measured slower at 1r+64c, 32,768 chains, 100 steps + 10 measures: four entries
   of a row sharing the loads of G_ik 310 ms, four rows sharing the loads of row j 242 ms, against 191 ms for this loop.) */
__device__ int g_refactor(double *st, long long ld, long long ch, const GLay &L) {
    int bad = 0;
    for (int i = 0; i < L.nr; i++)
        for (int j = 0; j <= i; j++) {
            double a = ST(L.COVR + g_tri(i, j));
            for (int k = 0; k < j; k++) a -= ST(L.FACR + g_tri(i, k)) * ST(L.FACR + g_tri(j, k));
            if (i == j) {
                if (!(a > 0.0)) { bad = 1; a = 0.0; }
                ST(L.FACR + g_tri(i, i)) = sqrt(a);
            } else {
                const double piv = ST(L.FACR + g_tri(j, j));
                ST(L.FACR + g_tri(i, j)) = piv > 0.0 ? a / piv : 0.0;
            }
        }
    const int dg = L.nc * (L.nc - 1);
    for (int i = 0; i < L.nc; i++)
        for (int j = 0; j <= i; j++) {
            if (i == j) {
                double a = ST(L.COVC + dg + i);
                for (int k = 0; k < j; k++) {
                    const double re = ST(L.FACC + g_hlo(i, k)), im = ST(L.FACC + g_hlo(i, k) + 1);
                    a -= re * re + im * im;
                }
                if (!(a > 0.0)) { bad = 1; a = 0.0; }
                ST(L.FACC + dg + i) = sqrt(a);
            } else {
                double are = ST(L.COVC + g_hlo(i, j)), aim = ST(L.COVC + g_hlo(i, j) + 1);
                for (int k = 0; k < j; k++) {
                    const double pr = ST(L.FACC + g_hlo(i, k)), pi = ST(L.FACC + g_hlo(i, k) + 1);
                    const double qr = ST(L.FACC + g_hlo(j, k)), qi = ST(L.FACC + g_hlo(j, k) + 1);
                    are -= pr * qr + pi * qi;
                    aim -= pi * qr - pr * qi;
                }
                const double piv = ST(L.FACC + dg + j);
                ST(L.FACC + g_hlo(i, j)) = piv > 0.0 ? are / piv : 0.0;
                ST(L.FACC + g_hlo(i, j) + 1) = piv > 0.0 ? aim / piv : 0.0;
            }
        }
    return bad;
}

#ifdef ME_GENERIC_USER
/* A user functor (CUDA text compiled at run time, me_set_energy_source) for a runtime shape: the contract is the fused
 * kernels' — me_user_energy(x[ME_NR], c_re[ME_NC], c_im[ME_NC], consts) — so the strided parameter vector is gathered into
 * per-thread arrays first (D > 32: they live in local memory; the functor's cost dominates anyway). */
__device__ __forceinline__ void g_user_gather(const double *v, long long ld, double *x, double *cr, double *ci) {
    for (int i = 0; i < ME_NR; i++) x[i] = v[(long long)i * ld];
    for (int j = 0; j < ME_NC; j++) { cr[j] = v[(long long)(ME_NR + j) * ld]; ci[j] = v[(long long)(ME_NR + ME_NC + j) * ld]; }
}
__device__ double g_user_energy(const double *v, long long ld, const double *k) {
    double x[ME_NR > 0 ? ME_NR : 1], cr[ME_NC > 0 ? ME_NC : 1], ci[ME_NC > 0 ? ME_NC : 1];
    g_user_gather(v, ld, x, cr, ci);
    return me_user_energy(x, cr, ci, k);
}
__device__ bool g_user_reject(const double *v, long long ld, const double *k) {
#ifdef ME_GENERIC_USER_REJECT
    double x[ME_NR > 0 ? ME_NR : 1], cr[ME_NC > 0 ? ME_NC : 1], ci[ME_NC > 0 ? ME_NC : 1];
    g_user_gather(v, ld, x, cr, ci);
    return me_user_reject(x, cr, ci, k);
#else
    return false;
#endif
}
#endif

/* built-in functors over a strided parameter vector v[i * ld] (same operation order as me_energies.cuh) */
__device__ double g_energy(int id, const double *v, long long ld, int nr, int nc, const double *k) {
#define V(i) v[(long long)(i) * ld]
    switch (id) {
    case ME_ENERGY_X2: return V(0) * V(0);
    case ME_ENERGY_XY_WELL: return k[0] * (V(0) * V(0) + V(1) * V(1));
    case ME_ENERGY_MIXED_WELL: {
        double area = 0.0, s = 0.0;
        for (int i = 0; i < nr; i++) { const double e = 1.0 - V(i); area = area + k[0] * (e * e); }
        for (int j = 0; j < nc; j++) {
            const double re = V(nr + j), im = V(nr + nc + j);
            const double a = re * re + im * im;
            s = s + (k[1] * a + k[2] * (a * a));
        }
        const double w = V(0) * V(1);
        return area + (k[3] != 0.0 ? fabs(w) : w) * (s / (double)nc);
    }
    case ME_ENERGY_CYLINDER: {
        const double a2 = V(0) * V(0);
        double quad = 0.0, tot = 0.0;
        for (int j = 0; j < nc; j++) {
            const double q = (double)(j - nc / 2);
            const double re = V(nr + j), im = V(nr + nc + j);
            const double m2 = re * re + im * im;
            quad = quad + (k[1] + (k[2] * (q * q)) * (1.0 + a2)) * m2;
            tot = tot + m2;
        }
        return (k[0] * a2 + quad) + (k[3] / (2.0 * (double)nc)) * (tot * tot);
    }
#ifdef ME_GENERIC_USER
    case ME_ENERGY_USER: return g_user_energy(v, ld, k);
#endif
    default: return 0.0;
    }
#undef V
}

/* hard wall of a proposal (ME:247): the cylinder functor's |a| >= 1, or the user functor's me_user_reject */
__device__ __forceinline__ bool g_wall(const MeParams &p, const double *v) {
    if (!p.use_reject) return false;
#ifdef ME_GENERIC_USER
    if (p.energy_id == ME_ENERGY_USER) return g_user_reject(v, p.ld, p.consts);
#endif
    return p.energy_id == ME_ENERGY_CYLINDER && fabs(v[0]) >= 1.0;
}

__device__ __forceinline__ void gk_energy_body(const MeParams &p) {
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_chains) return;
    const double *v = p.prop + ch;
    p.e_out[ch] = g_energy(p.energy_id, v, p.ld, p.n_real, p.n_complex, p.consts);
    if (p.rej_out)
        p.rej_out[ch] = g_wall(p, v) ? 1 : 0;
}

__device__ __forceinline__ void gk_init_body(const MeParams &p) {
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_chains) return;
    const GLay L(p.n_real, p.n_complex);
    const long long ld = p.ld;
    double *st = p.state;
    const int nr = L.nr, nc = L.nc;
    for (int i = 0; i < L.d; i++) {
        const double v = p.x0_broadcast ? p.x0[i] : p.x0[(long long)i * ld + ch];
        ST(L.X + i) = v;
        ST(L.MEAN + i) = v;
    }
    ST(L.SIG) = p.sigma0; ST(L.SIG + 1) = p.sigma0;
    for (int i = 0; i < nr; i++)
        for (int j = 0; j <= i; j++) ST(L.COVR + g_tri(i, j)) = p.cov_r0 ? p.cov_r0[i * nr + j] : (i == j ? 1.0 : 0.0);
    const int dg = nc * (nc - 1);
    for (int i = 0; i < nc; i++) {
        ST(L.COVC + dg + i) = p.cov_c0_re ? p.cov_c0_re[i * nc + i] : 1.0;
        for (int j = 0; j < i; j++) {
            ST(L.COVC + g_hlo(i, j)) = p.cov_c0_re ? p.cov_c0_re[i * nc + j] : 0.0;
            ST(L.COVC + g_hlo(i, j) + 1) = p.cov_c0_im ? p.cov_c0_im[i * nc + j] : 0.0;
        }
    }
    for (int i = 0; i < nr; i++) ST(L.OBSM + i) = fabs(ST(L.X + i));
    for (int j = 0; j < nc; j++) ST(L.OBSM + nr + j) = hypot(ST(L.X + nr + j), ST(L.X + nr + nc + j));
    for (int i = 0; i < nr; i++) ST(L.OBSM + nr + nc + i) = ST(L.X + i) * ST(L.X + i);
    const double e = p.have_e0 ? p.e_new[ch] : g_energy(p.energy_id, st + ch, ld, nr, nc, p.consts);
    ST(L.E) = e;
    ST(L.NACC) = 0.0;
    int status = 0;
    if (e != e) status |= ME_STATUS_ENERGY_NAN;
    if (g_refactor(st, ld, ch, L)) status |= ME_STATUS_NOT_PSD;
    ST(L.STATUS) = (double)status;
}

/* proposal (ME:261-302): x' = x + sigma_r L z ; c' = c + sigma_c conj(G) xi, xi = (z + i z')/sqrt2 */
__device__ void g_propose(const MeParams &p, const me::MathTables &tables, long long ch, unsigned step) {
    const GLay L(p.n_real, p.n_complex);
    const long long ld = p.ld;
    double *st = p.state;
    double *pr = p.prop + ch;
    const int nr = L.nr, nc = L.nc, d = L.d;
    if (p.inj_delta != nullptr) {
        const bool absolute = nc > 0 && p.group >= 3;     /* magnitude / phase records hold the proposal itself */
        for (int i = 0; i < d; i++)
            pr[(long long)i * ld] = p.inj_delta[(long long)i * ld + ch] + (absolute ? 0.0 : ST(L.X + i));
        if (nr > 0 && nc > 0 && p.group != 0)
            for (int i = 0; i < d; i++)
                if ((i < nr) != (p.group == 1)) pr[(long long)i * ld] = ST(L.X + i);
        return;
    }
    /* normals go to the scratch block first (z_k at scratch[k]) */
    double *z = p.scratch + ch;
    const me::Rng rng(p, p.chain_offset + (unsigned long long)ch);
    for (int q = 0; q < (d + 1) / 2; q++) {
        const me::U4 r = rng.bits(step, (unsigned)q);
        double z0, z1;
        me::Rng::box_muller<true>(r, tables, z0, z1);
        z[(long long)(2 * q) * ld] = z0;
        if (2 * q + 1 < d) z[(long long)(2 * q + 1) * ld] = z1;
    }
    if (nc > 0 && p.group >= 3) {
        /* magnitude-phase moves (ME:168-207, 304-317; same arithmetic as me::propose_magnitudes / propose_phases):
         * 3: |c_j|' = |c_j| + z_j sigma_c^2 C_jj (the reference's variance-as-deviation, ME:305,310),
         *    C_jj = sum_k |G_jk|^2;  4: c_j' = |c_j| e^{i theta}, theta from the angle word of Philox call j. */
        const int dg = nc * (nc - 1);
        const double sc2 = ST(L.SIG + 1) * ST(L.SIG + 1);
        for (int i = 0; i < nr; i++) pr[(long long)i * ld] = ST(L.X + i);
        for (int j = 0; j < nc; j++) {
            const double re = ST(L.X + nr + j), im = ST(L.X + nr + nc + j);
            const double mag = hypot(re, im);
            if (p.group == 3) {
                double cjj = ST(L.FACC + dg + j) * ST(L.FACC + dg + j);
                for (int k = 0; k < j; k++)
                    cjj += ST(L.FACC + g_hlo(j, k)) * ST(L.FACC + g_hlo(j, k))
                         + ST(L.FACC + g_hlo(j, k) + 1) * ST(L.FACC + g_hlo(j, k) + 1);
                const double nm = mag + z[(long long)j * ld] * (sc2 * cjj);
                const double ratio = nm / mag;
                const bool neg0 = __double2hiint(re) < 0;      /* zero modulus: atan2's signed zeros (see me_device.cuh) */
                pr[(long long)(nr + j) * ld] = mag > 0.0 ? re * ratio : (neg0 ? -nm : nm);
                pr[(long long)(nr + nc + j) * ld] = mag > 0.0 ? im * ratio
                                                              : nm * copysign(neg0 ? 1.2246467991473532e-16 : 0.0, im);
            } else {
                const me::U4 r = rng.bits(step, (unsigned)j);
                double sn, cs;
                me::sincospi_bits(r.z, sn, cs);
                pr[(long long)(nr + j) * ld] = mag * -cs;
                pr[(long long)(nr + nc + j) * ld] = mag * -sn;
            }
        }
        return;
    }
    const double sr = ST(L.SIG), sc = ST(L.SIG + 1) * 0.70710678118654752440;
    for (int i = 0; i < nr; i++) {
        double acc = 0.0;
        for (int j = 0; j <= i; j++) acc = acc + ST(L.FACR + g_tri(i, j)) * z[(long long)j * ld];
        pr[(long long)i * ld] = ST(L.X + i) + sr * acc;
    }
    const int dg = nc * (nc - 1);
    for (int i = 0; i < nc; i++) {
        double are = 0.0, aim = 0.0;
        for (int j = 0; j < i; j++) {
            const double lre = ST(L.FACC + g_hlo(i, j)), lim = -ST(L.FACC + g_hlo(i, j) + 1);
            const double zre = z[(long long)(nr + 2 * j) * ld], zim = z[(long long)(nr + 2 * j + 1) * ld];
            are = are + (lre * zre - lim * zim);
            aim = aim + (lre * zim + lim * zre);
        }
        are = are + ST(L.FACC + dg + i) * z[(long long)(nr + 2 * i) * ld];
        aim = aim + ST(L.FACC + dg + i) * z[(long long)(nr + 2 * i + 1) * ld];
        pr[(long long)(nr + i) * ld] = ST(L.X + nr + i) + sc * are;
        pr[(long long)(nr + nc + i) * ld] = ST(L.X + nr + nc + i) + sc * aim;
    }
    if (nr > 0 && nc > 0 && p.group != 0)          /* group-wise step: the other block keeps its value */
        for (int i = 0; i < d; i++)
            if ((i < nr) != (p.group == 1)) pr[(long long)i * ld] = ST(L.X + i);
}
__device__ __forceinline__ void gk_propose_body(const MeParams &p) {
    __shared__ me::MathTables tables;
    me::init_math_tables(tables);
    __syncthreads();
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_chains) return;
    g_propose(p, tables, ch, (unsigned)p.step0);
}

/* decision + sigma adaptation (ME:247-258, 319-338, 429-456); e_new / wall of THIS chain are passed in */
__device__ void g_accept(const MeParams &p, const me::MathTables &tables, long long ch, unsigned step, long long n_meas,
                         double e_new, bool wall) {
    const GLay L(p.n_real, p.n_complex);
    const long long ld = p.ld;
    double *st = p.state;
    const int kind = (L.nr > 0 && L.nc > 0) ? 0 : (L.nr > 0 ? 1 : 2);
    const bool grouped = kind == 0 && p.group != 0;
    const int sidx = grouped ? (p.group == 1 ? 0 : 1) : (kind == 2 ? 1 : 0);
    double sg = ST(L.SIG + sidx);
    int status = (int)ST(L.STATUS);
    const me::Gains g = me::make_gains(n_meas, p);
    bool accept = false;
    if (!wall) {
        if (e_new != e_new) status |= ME_STATUS_ENERGY_NAN;
        const double diff = e_new - ST(L.E);
        double u = 0.0;
        if (p.inj_u != nullptr) u = p.inj_u[ch];
        else if (diff > 0 && p.temp != 0) {
            const me::Rng rng(p, p.chain_offset + (unsigned long long)ch);
            me::Spare sp;
            me::Rng::keep_spare(rng.bits(step, 0u), 0, sp);
            u = me::Rng::accept_uniform(sp);
        }
        accept = me::decide<true>(diff, u, p, tables, g.hot);
        if (accept) {
            ST(L.E) = e_new;
            for (int i = 0; i < L.d; i++) ST(L.X + i) = p.prop[(long long)i * ld + ch];
            ST(L.NACC) += 1.0;
        }
    }
    if (!(L.nc > 0 && p.group == 4)) sg = me::adapt_sigma<true>(sg, accept, g, p);     /* phase redraw: ME:194-207 */
    ST(L.SIG + sidx) = sg;
    if (kind == 0 && !grouped) {
        ST(L.SIG + 1) = sg;
        if (!(sg > 0)) status |= ME_STATUS_SIGMA_NONPOS;
    }
    ST(L.STATUS) = (double)status;
    if (p.last_accept) p.last_accept[ch] = (unsigned char)accept;
}
__device__ __forceinline__ void gk_accept_body(const MeParams &p) {
    __shared__ me::MathTables tables;
    me::init_math_tables(tables);
    __syncthreads();
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_chains) return;
    g_accept(p, tables, ch, (unsigned)p.step0, p.n_meas0, p.e_new[ch], p.rej != nullptr && p.rej[ch] != 0);
}

/* measure (ME:342-427): the strict-order arithmetic of me::measure_update for runtime shapes; n = the counter AFTER the
 * increment, ts_row = the time-series row this measure writes. */
__device__ void g_measure(const MeParams &p, long long ch, long long n, long long ts_row) {
    const GLay L(p.n_real, p.n_complex);
    const long long ld = p.ld;
    double *st = p.state;
    double *old = p.scratch + ch;          /* old[i] at scratch[i][ch] */
    const int nr = L.nr, nc = L.nc, d = L.d;
    const double dn = (double)n, dn1 = (double)(n - 1), dn2 = (double)(n - 2);
    const double inv_n = 1.0 / dn, inv_n1 = 1.0 / dn1;
    const double shrink = dn1 / dn, decay = dn2 / dn1, grow = dn / dn1;
    const bool adapt_cov = n > 50;
    int status = (int)ST(L.STATUS);
    for (int i = 0; i < d; i++) old[(long long)i * ld] = ST(L.MEAN + i);
    for (int i = 0; i < nr; i++) {
        double m = ST(L.MEAN + i) * shrink;
        m = m + ST(L.X + i) / dn;
        ST(L.MEAN + i) = m;
    }
    if (nr > 0 && adapt_cov) {
        const double sig = ST(L.SIG), small = (sig * sig) / dn;
        for (int i = 0; i < nr; i++)
            for (int j = 0; j <= i; j++) {
                const double v = ST(L.COVR + g_tri(i, j)) * decay;
                const double add = ((old[(long long)i * ld] * old[(long long)j * ld] - grow * (ST(L.MEAN + i) * ST(L.MEAN + j)))
                                    + (ST(L.X + i) * ST(L.X + j)) / dn1) + (i == j ? small : 0.0);
                ST(L.COVR + g_tri(i, j)) = v + add;
            }
    }
    for (int j = 0; j < nc; j++) {
        double mr = ST(L.MEAN + nr + j) * shrink, mi = ST(L.MEAN + nr + nc + j) * shrink;
        mr = mr + ST(L.X + nr + j) * inv_n;
        mi = mi + ST(L.X + nr + nc + j) * inv_n;
        ST(L.MEAN + nr + j) = mr;
        ST(L.MEAN + nr + nc + j) = mi;
    }
    if (nc > 0 && adapt_cov) {
        const double sig = ST(L.SIG + 1), small = (sig * sig) / dn;
        const int dg = nc * (nc - 1);
#define OR(i) old[(long long)(nr + (i)) * ld]
#define OI(i) old[(long long)(nr + nc + (i)) * ld]
#define MR(i) ST(L.MEAN + nr + (i))
#define MI(i) ST(L.MEAN + nr + nc + (i))
#define XR(i) ST(L.X + nr + (i))
#define XI(i) ST(L.X + nr + nc + (i))
        for (int i = 0; i < nc; i++)
            for (int j = 0; j <= i; j++) {
                const double o_re = OR(i) * OR(j) + OI(i) * OI(j), o_im = OI(i) * OR(j) - OR(i) * OI(j);
                const double m_re = MR(i) * MR(j) + MI(i) * MI(j), m_im = MI(i) * MR(j) - MR(i) * MI(j);
                const double x_re = XR(i) * XR(j) + XI(i) * XI(j), x_im = XI(i) * XR(j) - XR(i) * XI(j);
                const double a_re = ((o_re - grow * m_re) + x_re * inv_n1) + (i == j ? small : 0.0);
                const double a_im = ((o_im - grow * m_im) + x_im * inv_n1);
                if (i == j) {
                    ST(L.COVC + dg + i) = ST(L.COVC + dg + i) * decay + a_re;
                } else {
                    ST(L.COVC + g_hlo(i, j)) = ST(L.COVC + g_hlo(i, j)) * decay + a_re;
                    ST(L.COVC + g_hlo(i, j) + 1) = ST(L.COVC + g_hlo(i, j) + 1) * decay + a_im;
                }
            }
#undef OR
#undef OI
#undef MR
#undef MI
#undef XR
#undef XI
    }
    for (int i = 0; i < nr; i++) ST(L.OBSM + i) = ST(L.OBSM + i) * shrink + fabs(ST(L.X + i)) / dn;
    for (int j = 0; j < nc; j++)
        ST(L.OBSM + nr + j) = ST(L.OBSM + nr + j) * shrink + hypot(ST(L.X + nr + j), ST(L.X + nr + nc + j)) / dn;
    for (int i = 0; i < nr; i++) ST(L.OBSM + nr + nc + i) = ST(L.OBSM + nr + nc + i) * shrink + (ST(L.X + i) * ST(L.X + i)) / dn;
    if (adapt_cov && g_refactor(st, ld, ch, L)) status |= ME_STATUS_NOT_PSD;
    ST(L.STATUS) = (double)status;
    if (p.record) {
        const int kind = (nr > 0 && nc > 0) ? 0 : (nr > 0 ? 1 : 2);
        const int tscols = d + (kind == 0 ? 3 : 2);
        double *row = p.ts + ts_row * (long long)tscols * ld + ch;
        for (int i = 0; i < d; i++) __stcs(row + (long long)i * ld, ST(L.X + i));
        __stcs(row + (long long)d * ld, ST(L.E));
        if (kind == 0) {
            __stcs(row + (long long)(d + 1) * ld, ST(L.SIG));
            __stcs(row + (long long)(d + 2) * ld, ST(L.SIG + 1));
        } else {
            __stcs(row + (long long)(d + 1) * ld, ST(L.SIG + (kind == 2 ? 1 : 0)));
        }
    }
}

/* The whole schedule of a runtime-shape engine with a built-in functor in ONE launch (the fused kernel's counterpart for
 * D > 32): n_blocks x (spm x [propose -> energy + wall -> decide] [+ measure]).  One thread per chain; the chain's state, its
 * 64 x 64 factor, the proposal and the normals stay in global memory (chain-minor: every word access of a warp is one
 * coalesced 256-byte line), so a step streams the factor once: 8 (n_r(n_r+1)/2 + n_c^2) bytes per chain-step — HBM bound. */
__device__ __forceinline__ void gk_run_body(const MeParams &p) {
    __shared__ me::MathTables tables;
    me::init_math_tables(tables);
    __syncthreads();
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_chains) return;
    long long n = p.n_meas0;
    unsigned step = (unsigned)p.step0;
    for (long long b = 0; b < p.n_blocks; b++) {
        for (long long k = 0; k < p.spm; k++, step++) {
            g_propose(p, tables, ch, step);
            const double *v = p.prop + ch;
            const bool wall = g_wall(p, v);
            const double e_new = wall ? 0.0 : g_energy(p.energy_id, v, p.ld, p.n_real, p.n_complex, p.consts);
            g_accept(p, tables, ch, step, n, e_new, wall);
        }
        if (p.do_measure) {
            n += 1;
            g_measure(p, ch, n, p.ts_row0 + b);
        }
    }
}

}  // namespace megk

/* entry points: ahead of time (built-in functors, me_generic.cu) under their own names; compiled at run time with a user
   functor (NVRTC, me_api.cu) as extern "C" me_gk_* */
#ifdef ME_NVRTC
#define ME_GK_KERNEL(name) extern "C" __global__ void me_##name(const __grid_constant__ MeParams p) { megk::name##_body(p); }
#else
#define ME_GK_KERNEL(name) __global__ void name(const __grid_constant__ MeParams p) { megk::name##_body(p); }
#endif
ME_GK_KERNEL(gk_energy)
ME_GK_KERNEL(gk_init)
ME_GK_KERNEL(gk_propose)
ME_GK_KERNEL(gk_accept)
ME_GK_KERNEL(gk_run)

#endif
