/* me_math.cuh — FP64 special functions specialised for the step kernel's argument ranges.
 *
 * The generic libdevice log/exp spend a third of their instructions on argument classes that cannot occur
 * here (denormals, negatives, infinities) and materialise their polynomial constants through UMOV pairs.  These
 * versions are table-driven (tables built once per CTA in shared memory), branch-free and accurate to about
 * 1 ulp on their stated domains:
 *     neg2log_unit(u)   = -2 ln u           for u in [2^-53, 1)      (Box-Muller radius squared)
 *     sqrt_pos(w)       = sqrt(w)           for normal w > 0         (Box-Muller radius)
 *     sincospi_bits(z)  = sin, cos(pi z 2^-31)  for a 32-bit integer z (Box-Muller angle; integer octant reduction)
 *     exp_nonpos(x)     = e^x               for x <= 0               (Metropolis acceptance probability, ME:334)
 * Fed to NVRTC as text: no #includes.
 */
#ifndef ME_MATH_CUH
#define ME_MATH_CUH

namespace me {

/* Polynomial coefficients live in the constant bank so that DFMA/DADD take them as c[bank][offset] operands;
 * immediates would be re-materialised through UMOV / IMAD.MOV pairs on every use (a quarter of the issue slots
 * of the first version of this kernel, see profiles/).  Exception: a double whose low 32 bits are zero IS a
 * single-instruction immediate, and the leading coefficient of each polynomial only needs ~20 bits — so those are
 * written as truncated literals (ME_C_*), which removes one LDC per polynomial from the step loop. */
#define ME_C_SIN6 0x1.e3f38p-12        /* 4.6153e-4  (kc[19] truncated; effect on sin(pi r) < 2e-18) */
#define ME_C_COS7 (-0x1.b264bp-14)     /* -1.0357e-4 (kc[27] truncated; effect < 2e-19) */
#define ME_C_EXP5 0x1.11111p-7         /* 1/120 truncated; effect on e^r < 1e-20 */
#define ME_C_2LN2HI 0x1.62e42p+0       /* 2 ln2_hi */
#define ME_C_64_LN2 0x1.71547p+6       /* 64/ln2, leading 21 bits (the reduction uses the k it produced, so only |r| grows, by 3e-7) */
#define ME_C_ANGLE 2097152.25          /* 2^21 + 1/4: bias of the integer angle reduction (polynomial sin/cos, parity build) */
#define ME_C_ANGLE_TAB 4503599629467648.0   /* 2^52 + 2^21: bias of the table-driven sin/cos residual (throughput build) */
#define ME_C_UNIT 0.99999999999999988898   /* 1 - 2^-53: bias of the radius uniform */

/* Constants that cannot be immediates (non-zero low word) and would otherwise be re-materialised inside the step loop
 * (UMOV pairs / LDC): the fused kernel loads them once and pins them in registers. */
struct Pins {
    double unit, angle, k64;
};

/* Table of the table-driven sin/cos (throughput build): 1024 intervals of the angle pi t, t = z 2^-31 in [0, 2):
 * {sin a_i, cos a_i} at the interval midpoints a_i = (i + 1/2) 2 pi / 1024, computed by the host library in long double
 * with entry i + 512 = -entry i EXACTLY (so that z and z ^ 0x80000000 give exactly opposite normals: the proposal is
 * symmetric bit for bit).  16 KB, stored right behind the log table. */
#define ME_SINTAB_ENTRIES 1024

/* Table of the table-driven log: 1024 intervals of the mantissa [1 + i/1024, 1 + (i+1)/1024): {rc_i, -2 l_i} with
 * rc_i = float(1 / upper edge), l_i = -ln(rc_i * (i >= 424 ? 2 : 1)).  16 KB, computed once per device by the host
 * library in long double (me_api.cu) and read through the L1 / read-only path: no per-CTA initialisation, and the finer
 * intervals (|r| <= 2^-10) cut the log1p polynomial from degree 8 to degree 5. */
#define ME_LOGTAB_ENTRIES 1024
__constant__ double me_kc[32] = {
    /* 0..4  log1p series coefficients that are not exact binary fractions: 1/7, -1/6, 1/5, 1/3, (unused) */
    0.14285714285714285, -0.16666666666666666, 0.2, 0.33333333333333331, 0.0,
    /* 5..6  -ln2_lo (ln2 = ln2_hi + ln2_lo, ln2_hi = 0x1.62e42p-1 has 32 zero low bits), (unused) */
    -4.7493250390316726e-07, 0.0,
    /* 7..10 exp: (unused), -ln2_hi/64, -ln2_lo/64, (unused) ; 11..12: 1/24, 1/6 */
    0.0, -1.083042469326756e-02, -2.9815858269852933e-12, 0.0,
    4.1666666666666664e-2, 0.16666666666666666,
    /* 13..19 sin(pi r)/r in s = r^2, r in [-1/4, 1/4] (near-minimax, rel. error 4e-17); [19] is ME_C_SIN6 */
    3.141592653589793, -5.167712780049954, 2.5501640398733763, -0.5992645289396449, 0.08214586918000175,
    -0.007370021586907771, 0.000461531855383581,
    /* 20..27 cos(pi r) in s (rel. error 2e-17); [27] is ME_C_COS7 */
    1.0, -4.934802200544679, 4.058712126416747, -1.3352627688519174, 0.23533063019088787, -0.025806885652951306,
    0.0019294657440800042, -0.00010356747255199479,
    /* 28..31 residual rotation of the table-driven sin/cos in the INTEGER residual v = (z mod 2^22) - 2^21 (angle
       b = kappa v, kappa = 2 pi / 2^32): kappa, -kappa^3/6, -kappa^2/2, kappa^4/24 */
    1.4629180792671596e-09, -5.218056424438286e-28, -1.0700646533233578e-18, 1.90839727048673e-37};

struct MathTables {
    double exp2t[64];      /* 2^(j/64) */
    double pins[4];        /* ME_C_UNIT, ME_C_ANGLE, ME_C_64_LN2 (see Pins) */
};

/* Called by every thread of the CTA before any use; the caller synchronises afterwards. */
__device__ __forceinline__ void init_math_tables(MathTables &T) {
    for (int j = threadIdx.x; j < 64; j += blockDim.x) T.exp2t[j] = exp2((double)j * 0.015625);
    if (threadIdx.x == 0) { T.pins[0] = ME_C_UNIT; T.pins[1] = ME_C_ANGLE_TAB; T.pins[2] = ME_C_64_LN2; T.pins[3] = ME_C_ANGLE; }
}

/* After the table barrier: the pinned constants as register values.  Coming from shared memory they are opaque to
 * ptxas, which would otherwise re-materialise the literals inside the step loop. */
__device__ __forceinline__ Pins load_pins(const MathTables &T) {
    Pins c;
    c.unit = T.pins[0]; c.angle = T.pins[1]; c.k64 = T.pins[2];
    return c;
}

/* -2 ln(u), u in [2^-53, 1).  u = 2^e m, m in [1,2); interval i = top 10 mantissa bits; r = m rc_i - 1 in
 * [-2^-10, 2^-24]; ln m = l_i + ln2 [i >= 424] + log1p(r) with the mantissa range folded to [0.71, 1.42) so that
 * u -> 1 keeps full relative accuracy (top interval has rc = 1/2, l = 0 exactly).  The fold is an integer carry:
 * adding 600 to the 10-bit interval field of the high word overflows into the exponent field exactly when i >= 424.
 * The result is > 0 for every u < 1 (|.| only clears a sign that rounding could set for u within 2^-50 of 1).
 * Absolute error < 5e-13 (relative < 2.4e-10 where -2 ln u < 2^-9; checked against long-double log). */
/* Where the log table is read from: global memory through the read-only path (small shapes: one-warp CTAs that live for
 * one work item), or a per-CTA copy in shared memory (larger shapes: few, long-lived CTAs with two warps per
 * sub-partition, where the shorter LDS latency matters). */
struct LogTabGlobal {
    const double2 *base;          /* log table, followed by the sin/cos table */
    __device__ __forceinline__ double2 at(unsigned byte_off) const {
        return __ldg(reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(base) + byte_off));
    }
    __device__ __forceinline__ double2 sincos_at(unsigned byte_off) const { return at(byte_off + ME_LOGTAB_ENTRIES * 16u); }
};
struct LogTabShared {
    const double2 *base;          /* points into a __shared__ array holding both tables */
    __device__ __forceinline__ double2 at(unsigned byte_off) const {
        return *reinterpret_cast<const double2 *>(reinterpret_cast<const char *>(base) + byte_off);
    }
    __device__ __forceinline__ double2 sincos_at(unsigned byte_off) const { return at(byte_off + ME_LOGTAB_ENTRIES * 16u); }
};

template <class Tab>
__device__ __forceinline__ double neg2log_unit(double u, const Tab &logtab) {
    const int hi = __double2hiint(u), lo = __double2loint(u);
    /* {rc, -2 l}; byte offset added as a 32-bit quantity (a 64-bit index would be scaled with IMAD.WIDE on the pipe the
       Philox rounds need) */
    const double2 e = logtab.at((unsigned)((hi >> 6) & ((ME_LOGTAB_ENTRIES - 1) << 4)));
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    /* n = 1023 - (biased exponent + [i >= 424]) >= 0: minus the exponent of the folded mantissa, as one subtract+shift */
    const int n = (int)((unsigned)(0x3ff69fff - hi) >> 20);
    const double r = fma(m, e.x, -1.0);
    /* log1p(r) = r - r^2/2 + r^3/3 ; |r| <= 2^-10 -> truncation r^4/4 < 2.3e-13 absolute (accuracy budget of the draw
       stage: 1e-10 — the proposal is symmetric whatever the accuracy, and the Metropolis threshold is only trusted to
       4e-5 T before the exact fallback, see accept_window) */
    double p = fma(r, me_kc[3], -0.5);
    p = fma(r * r, p, r);
    const double nd = __hiloint2double(0x43300000, n) - 4503599627370496.0;              /* (double)n */
    /* -2 ln u = n (2 ln2_hi) + (-2 l + -2 (p - n ln2_lo)) */
    const double q = fma(nd, me_kc[5], p);                         /* -ln2_lo */
    const double t = fma(q, -2.0, e.y);
    const double w = fma(nd, ME_C_2LN2HI, t);                      /* exact product (11 x 21 bits) */
    return __hiloint2double(__double2hiint(w) & 0x7fffffff, __double2loint(w));
}

/* sqrt(w) for a normal, strictly positive w: MUFU.RSQ64H seed (2^-22) and one Newton step for 1/sqrt (no range check,
 * no slow path — the Box-Muller argument is in [2e-16, 74]). */
__device__ __forceinline__ double sqrt_pos(double w) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
    const double e = fma(-w, y * y, 1.0);
    const double yh = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));       /* y / 2 */
    const double y1 = fma(yh, e, y);
    return w * y1;                     /* relative error (3/8) e^2 < 1e-13: inside the draw stage's accuracy budget */
}

/* the same with the final correction step (relative error < 2^-52.5) */
__device__ __forceinline__ double sqrt_pos_full(double w) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(w));
    const double e = fma(-w, y * y, 1.0);
    const double yh = __hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y));
    const double y1 = fma(yh, e, y);
    const double s = w * y1;
    const double y1h = __hiloint2double(__double2hiint(y1) - 0x00100000, __double2loint(y1));
    return fma(fma(-s, s, w), y1h, s);
}

/* 1 / sqrt(d) for a normal, strictly positive d: MUFU.RSQ64H seed + two Newton steps (relative error ~1e-16).  Used by
 * the throughput build's Cholesky refactorisation, where it replaces a libdevice sqrt and a division per pivot. */
__device__ __forceinline__ double rsqrt_pos(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y * y, 1.0);
    y = fma(__hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y)), e, y);
    e = fma(-d, y * y, 1.0);
    return fma(__hiloint2double(__double2hiint(y) - 0x00100000, __double2loint(y)), e, y);
}

/* e^x for x <= 0 (clamped at -700: the result is only compared with a uniform on a 2^-44 grid).  x > 0 or NaN
 * gives an unspecified value without trapping (the caller ignores it: a downhill move is accepted anyway).
 * k = round(64 x / ln2), x = k ln2/64 + r, |r| <= ln2/128; e^x = 2^(k>>6) 2^((k&63)/64) e^r. */
__device__ __forceinline__ double exp_nonpos(double x, const MathTables &T, const double k64 = ME_C_64_LN2) {
    x = x < -700.0 ? -700.0 : x;
    const double kd = fma(x, k64, 6755399441055744.0);                     /* 64/ln2, magic 1.5 * 2^52 */
    const int k = __double2loint(kd);
    const double kf = kd - 6755399441055744.0;
    double r = fma(kf, me_kc[8], x);                                        /* ln2_hi/64 (32 significant bits) */
    r = fma(kf, me_kc[9], r);                                               /* ln2_lo/64 */
    double p = fma(r, ME_C_EXP5, me_kc[11]);
    p = fma(r, p, me_kc[12]);
    p = fma(r, p, 0.5);
    p = fma(r * r, p, r);                                                   /* e^r - 1 */
    const double t = T.exp2t[k & 63];
    const double v = fma(t, p, t);
    return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
}

/* sin(pi t), cos(pi t) for t = z 2^-31 in [0, 2), z a 32-bit integer.  The octant reduction is done on the integer:
 * zz = z + 2^29; quadrant q = zz >> 30 (round-half-up of 2t); r = t - q/2 = (zz mod 2^30) 2^-31 - 1/4 in [-1/4, 1/4),
 * exact, one DADD.  Polynomials in r^2; quadrant fix-up = one swap and two sign-bit XORs. */
__device__ __forceinline__ void sincospi_bits(unsigned z, double &sn, double &cs, const double bias = ME_C_ANGLE) {
    const unsigned zz = z + 0x20000000u;
    const double r = __hiloint2double(0x41400000, (int)(zz & 0x3fffffffu)) - bias;
    const double s = r * r;
    double ps = fma(s, ME_C_SIN6, me_kc[18]);
    ps = fma(s, ps, me_kc[17]);
    ps = fma(s, ps, me_kc[16]);
    ps = fma(s, ps, me_kc[15]);
    ps = fma(s, ps, me_kc[14]);
    ps = fma(s, ps, me_kc[13]);
    const double sr = r * ps;
    double pc = fma(s, ME_C_COS7, me_kc[26]);
    pc = fma(s, pc, me_kc[25]);
    pc = fma(s, pc, me_kc[24]);
    pc = fma(s, pc, me_kc[23]);
    pc = fma(s, pc, me_kc[22]);
    pc = fma(s, pc, me_kc[21]);
    const double cr = fma(s, pc, 1.0);
    /* q & 3: 0 -> (sr, cr), 1 -> (cr, -sr), 2 -> (-sr, -cr), 3 -> (-cr, sr) */
    const bool odd = (zz & 0x40000000u) != 0;
    const double a = odd ? cr : sr, b = odd ? sr : cr;
    sn = __hiloint2double(__double2hiint(a) ^ (int)(zz & 0x80000000u), __double2loint(a));
    cs = __hiloint2double(__double2hiint(b) ^ (int)((zz + 0x40000000u) & 0x80000000u), __double2loint(b));
}

/* sin(pi t), cos(pi t) for t = z 2^-31, table-driven (throughput build): interval i = z >> 22 of 1024, midpoint values
 * {sin a_i, cos a_i} from the table, residual angle b = kappa v with the exact integer v = (z mod 2^22) - 2^21 in
 * [-2^21, 2^21) (one DADD on a mantissa-assembled double), |b| <= pi/1024:
 *     sin b = v (kappa - kappa^3/6 v^2)          (b^5/120 < 2.3e-15)
 *     cos b - 1 = v^2 (-kappa^2/2 + kappa^4/24 v^2)  (b^6/720 < 2e-18)
 *     sin(a + b) = sa + (sa (cos b - 1) + ca sin b),   cos(a + b) = ca + (ca (cos b - 1) - sa sin b).
 * 10 FP64 instructions and no quadrant fix-up (the table covers the full circle), against 17 + 8 integer ones for the
 * polynomial version; absolute error < 1e-14. */
template <class Tab>
__device__ __forceinline__ void sincospi_tab(unsigned z, double &sn, double &cs, const Tab &tab, const double bias) {
    const double2 e = tab.sincos_at((z >> 18) & ((ME_SINTAB_ENTRIES - 1) << 4));        /* {sin a_i, cos a_i} */
    const double v = __hiloint2double(0x43300000, (int)(z & 0x003fffffu)) - bias;        /* exact */
    const double v2 = v * v;
    const double sb = v * fma(v2, me_kc[29], me_kc[28]);
    const double cm = v2 * fma(v2, me_kc[31], me_kc[30]);
    sn = fma(e.x, cm, fma(e.y, sb, e.x));
    cs = fma(e.y, cm, fma(-e.x, sb, e.y));
}

}  // namespace me

#endif
