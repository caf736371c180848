/* me_math.cuh — FP64 special functions specialised for the step kernel's argument ranges.
 *
 * The generic libdevice log/exp spend a third of their instructions on argument classes that cannot occur
 * here (denormals, negatives, infinities) and materialise their polynomial constants through UMOV pairs.  These
 * versions are table-driven (tables built once per CTA in shared memory), branch-free and accurate to about
 * 1 ulp on their stated domains:
 *     neg2log_unit(u)   = -2 ln u           for u in [2^-53, 1]      (Box-Muller radius)
 *     exp_nonpos(x)     = e^x               for x <= 0               (Metropolis acceptance probability, ME:334)
 * Fed to NVRTC as text: no #includes.
 */
#ifndef ME_MATH_CUH
#define ME_MATH_CUH

namespace me {

/* Polynomial coefficients live in the constant bank so that DFMA/DADD take them as c[bank][offset] operands;
 * immediates would be re-materialised through UMOV / IMAD.MOV pairs on every use (a quarter of the issue slots
 * of the first version of this kernel, see profiles/). */
__constant__ double me_kc[32] = {
    /* 0..4  log1p series coefficients that are not exact binary fractions: 1/7, -1/6, 1/5, 1/3, (unused) */
    0.14285714285714285, -0.16666666666666666, 0.2, 0.33333333333333331, 0.0,
    /* 5..6  ln2_lo, ln2_hi */
    1.9082149292705877e-10, 0.6931471803691238,
    /* 7..10 exp: 64/ln2, -ln2_hi/64, -ln2_lo/64, 1/120 ; 11..12: 1/24, 1/6 */
    92.332482616893657, -1.083042469326756e-02, -2.9815858269852933e-12, 8.3333333333333332e-3,
    4.1666666666666664e-2, 0.16666666666666666,
    /* 13..19 sin(pi r)/r in s = r^2, r in [-1/4, 1/4] (near-minimax, rel. error 4e-17) */
    3.141592653589793, -5.167712780049954, 2.5501640398733763, -0.5992645289396449, 0.08214586918000175,
    -0.007370021586907771, 0.000461531855383581,
    /* 20..27 cos(pi r) in s (rel. error 2e-17) */
    1.0, -4.934802200544679, 4.058712126416747, -1.3352627688519174, 0.23533063019088787, -0.025806885652951306,
    0.0019294657440800042, -0.00010356747255199479,
    0.0, 0.0, 0.0, 0.0};

struct MathTables {
    double logt[128][2];   /* interval i of the mantissa [1 + i/128, 1 + (i+1)/128): {rc_i, l_i} with
                              rc_i = float(1 / upper edge), l_i = -ln(rc_i * (i >= 53 ? 2 : 1)) */
    double exp2t[64];      /* 2^(j/64) */
};

/* Called by every thread of the CTA before any use; the caller synchronises afterwards. */
__device__ __forceinline__ void init_math_tables(MathTables &T) {
    for (int i = threadIdx.x; i < 128; i += blockDim.x) {
        const double c = 1.0 + (double)(i + 1) * 0.0078125;
        const double rc = (double)(float)(1.0 / c);
        T.logt[i][0] = rc;
        T.logt[i][1] = -log(i >= 53 ? rc * 2.0 : rc);
    }
    for (int j = threadIdx.x; j < 64; j += blockDim.x) T.exp2t[j] = exp2((double)j * 0.015625);
}

/* -2 ln(u), u in [2^-53, 1].  u = 2^e m, m in [1,2); interval i = top 7 mantissa bits; r = m rc_i - 1 in
 * [-2^-7, 2^-24]; ln m = l_i + ln2 [i >= 53] + log1p(r) with the mantissa range folded to [0.71, 1.42) so that
 * u -> 1 keeps full relative accuracy (top interval has rc = 1/2, l = 0 exactly). */
__device__ __forceinline__ double neg2log_unit(double u, const MathTables &T) {
    const int hi = __double2hiint(u), lo = __double2loint(u);
    const int i = (hi >> 13) & 127;
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const int e = (hi >> 20) - 1023 + (i >= 53 ? 1 : 0);
    const double rc = T.logt[i][0], l = T.logt[i][1];
    const double r = fma(m, rc, -1.0);
    /* log1p(r) = r - r^2/2 + r^3/3 - ... - r^8/8 ; |r| <= 2^-7 -> truncation below 2^-66.  Horner on purpose:
       the Estrin forms of this and of the sin/cos/exp polynomials (dependency depth 4 instead of 8) were measured
       and are NOT faster here (7.3e10 vs 7.6e10 chain-steps/s, 164 instead of 114 registers uncapped). */
    double p = fma(r, -0.125, me_kc[0]);
    p = fma(r, p, me_kc[1]);
    p = fma(r, p, me_kc[2]);
    p = fma(r, p, -0.25);
    p = fma(r, p, me_kc[3]);
    p = fma(r, p, -0.5);
    p = fma(r * r, p, r);
    const double ed = __hiloint2double(0x43300000, e ^ 0x80000000) - 4503601774854144.0;   /* (double)e */
    /* ln u = e ln2_hi + (l + (p + e ln2_lo)); returns -2 ln u >= 0 */
    const double t = fma(ed, me_kc[5], p) + l;                     /* ln2_lo */
    const double ln_u = fma(ed, me_kc[6], t);                      /* ln2_hi (low 21 bits zero) */
    return fmax(-2.0 * ln_u, 0.0);
}

/* e^x for x <= 0 (clamped at -700: the result is only compared with a uniform on a 2^-53 grid).
 * k = round(64 x / ln2), x = k ln2/64 + r, |r| <= ln2/128; e^x = 2^(k>>6) 2^((k&63)/64) e^r. */
__device__ __forceinline__ double exp_nonpos(double x, const MathTables &T) {
    x = fmax(x, -700.0);
    const double kd = fma(x, me_kc[7], 6755399441055744.0);                /* 64/ln2, magic 1.5 * 2^52 */
    const int k = __double2loint(kd);
    const double kf = kd - 6755399441055744.0;
    double r = fma(kf, me_kc[8], x);                                        /* ln2_hi/64 (32 significant bits) */
    r = fma(kf, me_kc[9], r);                                               /* ln2_lo/64 */
    double p = fma(r, me_kc[10], me_kc[11]);
    p = fma(r, p, me_kc[12]);
    p = fma(r, p, 0.5);
    p = fma(r * r, p, r);                                                   /* e^r - 1 */
    const double t = T.exp2t[k & 63];
    const double v = fma(t, p, t);
    return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
}

/* sin(pi t), cos(pi t) for t in [0, 2): exact reduction to r in [-1/4, 1/4] around the nearest multiple of 1/2
 * (magic-number rounding, no F2I / FRND), polynomials in r^2, quadrant fix-up by selects. */
__device__ __forceinline__ void sincospi_02(double t, double &sn, double &cs) {
    const double kd = (t + t) + 6755399441055744.0;
    const int q = __double2loint(kd);
    const double r = fma(kd - 6755399441055744.0, -0.5, t);
    const double s = r * r;
    double ps = fma(s, me_kc[19], me_kc[18]);
    ps = fma(s, ps, me_kc[17]);
    ps = fma(s, ps, me_kc[16]);
    ps = fma(s, ps, me_kc[15]);
    ps = fma(s, ps, me_kc[14]);
    ps = fma(s, ps, me_kc[13]);
    const double sr = r * ps;
    double pc = fma(s, me_kc[27], me_kc[26]);
    pc = fma(s, pc, me_kc[25]);
    pc = fma(s, pc, me_kc[24]);
    pc = fma(s, pc, me_kc[23]);
    pc = fma(s, pc, me_kc[22]);
    pc = fma(s, pc, me_kc[21]);
    const double cr = fma(s, pc, 1.0);
    /* q & 3: 0 -> (sr, cr), 1 -> (cr, -sr), 2 -> (-sr, -cr), 3 -> (-cr, sr) */
    const double a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
    sn = (q & 2) ? -a : a;
    cs = ((q + 1) & 2) ? -b : b;
}

}  // namespace me

#endif
