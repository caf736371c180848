/* me_k4.cu — shared-covariance step kernel for large parameter spaces (BASELINE config 4: 1 real + 64 complex
 * Fourier-mode coefficients, 32,768 chains), tcgen05 tensor cores.
 *
 * Same Metropolis step as me_device.cuh (reference metropolis_engine.py:241-259: proposal, hard wall ME:247, energy
 * ME:250, decision ME:319-338, Robbins-Monro width ME:429-438), but the proposal covariance of the complex block is
 * SHARED by all chains (pooled at measure boundaries), so the proposal increments of a tile of 128 chains are one
 * dense contraction
 *        Delta[128 chains x 128] = Z[128 chains x 128 normals] . B^T[128 x 128]
 * with B the real embedding of conj(G)/sqrt2, C_c = G G^H (ME:288-302: w ~ CN(0, sigma^2 conj(C_c))).  That
 * contraction runs on the 5th-generation tensor cores: Z is generated in-kernel (Philox4x32-10 + Box-Muller) straight
 * into the UMMA canonical K-major shared-memory layout as BF16, B is staged once per CTA, the FP32 accumulator lives
 * in TMEM with TMEM lane = chain, and the epilogue (tcgen05.ld) gives every thread the increments of its own chain
 * for the FP64 energy sum.  Reduced precision only perturbs the proposal shape; the proposal stays symmetric
 * (signs of the normals come from independent random bits), so detailed balance is exact; state, energy, energy
 * difference and the accept test are FP64.
 *
 * Layout of the per-chain state block state[word*ld + chain] (shared-covariance engines keep no per-chain covariance):
 *   X 129 (a, Re c_0..63, Im c_0..63) | E | SIG | MEAN 129 | OBSM 66 | NACC | STATUS      = 328 words
 * Inside the kernel the complex block is held in shared memory in interleaved order n = 2j (Re c_j), 2j+1 (Im c_j),
 * [n][chain], 128 KB per tile of 128 chains.
 */
#include <cuda_runtime.h>
#include <cuda_bf16.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/me_b200.h"
#include "me_params.h"
#include "me_math.cuh"

namespace {

#ifndef K4_INT_CVT
#define K4_INT_CVT 0       /* 1: FP32->FP64 with integer instructions instead of F2F (XU pipe); measured: no gain (115.0 vs
                              113.0 us per 10 steps at 32,768 chains) — the kernel is bound by its barrier/MMA-wait structure */
#endif
constexpr int K4_NC = 64;
constexpr int K4_N = 2 * K4_NC;          /* embedded real dimension = MMA N = MMA K */
constexpr int K4_TILE = 128;             /* chains per tile = MMA M = TMEM lanes */
constexpr int K4_THREADS = 512;
constexpr int K4_D = 1 + K4_N;

/* state-block word offsets */
constexpr int K4_X = 0, K4_E = K4_D, K4_SIG = K4_D + 1, K4_MEAN = K4_D + 2, K4_OBSM = K4_MEAN + K4_D,
              K4_NOBS = 2 + K4_NC, K4_NACC = K4_OBSM + K4_NOBS, K4_STATUS = K4_NACC + 1, K4_WORDS = K4_STATUS + 1;

struct K4Params {
    double *state;
    long long ld, n_chains;
    unsigned long long chain_offset;
    unsigned rk[20];
    unsigned long long step0;
    long long n_steps;
    long long chains_per_cta;      /* step kernel: contiguous chains per CTA (a multiple of 32) */
    long long n_meas;              /* measure_step_counter (for the Robbins-Monro gain) */
    double temp, inv_temp, target, ratio;
    int m;
    double kappa, alpha, gamma, beta;   /* cylinder energy constants */
    int use_wall;
    const void *factor;            /* B operand, BF16, UMMA canonical K-major layout [16 k-chunks][128 n][8] */
    const double *s_a;             /* device scalar: shared proposal std of the real parameter */
    unsigned char *last_accept;
    float *dbg_z;                  /* optional [128 k][ld]: the normals of the FIRST step of the launch (tests) */
    float *dbg_delta;              /* optional [128 n][ld]: the tensor-core increments of the first step */
    /* measure */
    double *ts;
    long long ts_row;
    int record;
};

/* -------------------------------------------------------------------------------------------- PTX helpers */
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
/* bounded wait: a wrong descriptor must not hang the GPU — trap instead */
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

/* D[tmem] (+)= A[smem] . B[smem]^T, BF16 inputs, FP32 accumulate, M = 128, N = 128, K = 16 */
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

/* K-major, no swizzle: core matrix = 8 rows x 16 B contiguous; row groups 128 B apart (SBO), the two 16-byte
 * K chunks of one K=16 MMA 2048 B apart (LBO); descriptor fields in 16-byte units; version 1 (Blackwell). */
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)(2048 >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) |
           (1ull << 46);
}
/* instruction descriptor: D = F32, A = B = BF16, both K-major, N = 128, M = 128 */
constexpr uint32_t K4_IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(K4_N >> 3) << 17) |
                              ((uint32_t)(K4_TILE >> 4) << 24);

/* -------------------------------------------------------------------------------------------- RNG (FP32 path) */
struct U4 { unsigned x, y, z, w; };
__device__ __forceinline__ U4 philox(unsigned c0, unsigned c1, unsigned c2, unsigned c3, const unsigned *rk) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
        const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
        const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ rk[2 * r];
        const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ rk[2 * r + 1];
        c0 = n0; c1 = (unsigned)p1; c2 = n2; c3 = (unsigned)p0;
    }
    U4 o; o.x = c0; o.y = c1; o.z = c2; o.w = c3;
    return o;
}
/* two normals from 32 random bits: radius uniform from the high 16 bits, angle from the low 16 (2 quadrant bits + 14-bit
 * fraction).  The operand these normals feed is BF16 (8 significant bits), so a 2^-16 grid for the radius uniform and a
 * 1e-4 rad grid for the angle are already below its rounding; the radius is capped at sqrt(2 ln 2^16) = 4.7.  Half the
 * Philox calls of a 64-bit recipe — the generator's IMAD.WIDE rounds are the largest single cost of the step kernel.
 * Built for the SM's scarcest pipe: no integer->float conversions (the uniforms are assembled as float mantissas),
 * sin/cos as FP32 polynomials after an integer quadrant reduction, so the only XU operations left are one MUFU.LG2 and
 * one MUFU.SQRT per PAIR (was five: 2 I2F, LG2, RSQ, SIN, COS).  The two signs and the sin/cos swap come from
 * independent bits, so the pair's law is exactly symmetric whatever the accuracy of the approximations: the proposal
 * stays symmetric and detailed balance exact. */
__device__ __forceinline__ void normal_pair_f32(unsigned bits, float &z0, float &z1) {
    const float u = 2.0f - __uint_as_float(0x3f800000u | ((bits >> 16) << 7));     /* (0, 1], multiples of 2^-16 */
    float lg;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(u));                         /* u >= 2^-16: no denormal path */
    const float w = -1.3862943611f * lg;                                            /* -2 ln u >= 0 */
    float rad;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(w));
    const unsigned zz = (bits << 16) + 0x20000000u;                                 /* quadrant = zz >> 30 (rounded) */
    const float v = __uint_as_float(0x3f800000u | ((zz >> 7) & 0x007fffffu)) - 1.5f;   /* [-1/2, 1/2): angle (pi/2) v */
    const float q = v * v;
    float ps = fmaf(q, -0.0046817541f, 0.0796926263f);                              /* sin((pi/2) v) / v */
    ps = fmaf(q, ps, -0.6459640975f);
    ps = fmaf(q, ps, 1.5707963268f);
    const float sr = v * ps;
    float pc = fmaf(q, 0.0009192603f, -0.0208634807f);                              /* cos((pi/2) v) */
    pc = fmaf(q, pc, 0.2536695079f);
    pc = fmaf(q, pc, -1.2337005501f);
    const float cr = fmaf(q, pc, 1.0f);
    /* quadrant 0: (cos, sin) = (cr, sr); 1: (-sr, cr); 2: (-cr, -sr); 3: (sr, -cr) */
    const bool odd = (zz & 0x40000000u) != 0;
    const float cs = odd ? sr : cr, sn = odd ? cr : sr;
    z0 = rad * __uint_as_float(__float_as_uint(cs) ^ ((zz + 0x40000000u) & 0x80000000u));
    z1 = rad * __uint_as_float(__float_as_uint(sn) ^ (zz & 0x80000000u));
}
__device__ __forceinline__ double u53(unsigned hi, unsigned lo) {
    const double a = __hiloint2double(0x43300000 - (27 << 20), (int)(hi >> 5)) - 33554432.0;
    const double b = __hiloint2double(0x43300000 - (53 << 20), (int)(lo >> 6)) - 0.5;
    return a + b;
}
__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const unsigned *>(&v);
}

/* FP32 bit pattern -> double, exact for normal values, with integer instructions only (F2F.F64.F32 runs on the XU pipe
 * at a quarter of a warp per cycle, and the epilogues convert 64 increments per thread and step).  +-0 and denormals map
 * to +-2^-126-sized values, which are added to O(1) coordinates. */
__device__ __forceinline__ double f32_bits_to_f64(uint32_t f) {
#if K4_INT_CVT
    const uint32_t hi = ((((f << 1) >> 4) + 0x38000000u) | (f & 0x80000000u));
    return __hiloint2double((int)hi, (int)(f << 29));
#else
    return (double)__uint_as_float(f);
#endif
}

/* cylinder-style energy from its sufficient statistics (same functional form as me::EnergyCylinder) */
__device__ __forceinline__ double k4_energy(double a, double tot, double qsum, const K4Params &p) {
    const double a2 = a * a;
    return (p.kappa * a2 + (p.alpha * tot + p.gamma * (1.0 + a2) * qsum)) + (p.beta / (2.0 * K4_NC)) * (tot * tot);
}

/* -------------------------------------------------------------------------------------------- the step kernel */
struct K4Smem {
    double xs[K4_N][K4_TILE];            /* complex block, interleaved [n][chain]              128 KB */
    alignas(1024) unsigned char zs[K4_TILE * K4_N * 2];   /* A operand (normals), BF16           32 KB */
    alignas(1024) unsigned char ls[K4_N * K4_N * 2];      /* B operand (factor), BF16            32 KB */
    double part[4][2][K4_TILE];
    double a_s[K4_TILE], e_s[K4_TILE], sig_s[K4_TILE], za_s[2][K4_TILE], u_s[2][K4_TILE], nacc_s[K4_TILE];
    int acc_s[K4_TILE];
    int status_s[K4_TILE];
    me::MathTables tables;
    uint64_t mbar;
    uint32_t tmem_slot;
};

__global__ void __launch_bounds__(K4_THREADS, 1) k4_steps(const __grid_constant__ K4Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    K4Smem &S = *reinterpret_cast<K4Smem *>(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = 32 * (warp & 3) + lane;        /* chain within the tile = TMEM lane this warp may read */
    const int g = warp >> 2;                     /* column group: embedded coordinates [32g, 32g+32) = modes [16g, 16g+16) */
    const long long ld = p.ld;

    me::init_math_tables(S.tables);
    if (tid == 0) {
        mbar_init(&S.mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&S.tmem_slot, K4_N);
    /* stage the shared factor once per CTA */
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.factor);
        uint4 *dst = reinterpret_cast<uint4 *>(S.ls);
        for (int i = tid; i < K4_N * K4_N * 2 / 16; i += K4_THREADS) dst[i] = src[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = S.tmem_slot;
    const double s_a = *p.s_a;
    const double q0 = (double)(16 * g - K4_NC / 2);      /* wavenumber of this thread's first mode */
    const uint32_t zs_addr = smem_u32(S.zs), ls_addr = smem_u32(S.ls);
    uint32_t parity = 0;

    double f = (double)p.n_meas / (double)p.m;
    if (!(f > 200.0)) f = 200.0;
    const double g_up = p.ratio * (1 - p.target) / f, g_down = p.ratio * p.target / f;

    /* Each CTA owns a contiguous range of chains (a multiple of 32) and walks it in tiles of up to 128: 32,768 chains on
       147 CTAs are 224 chains each = one full tile and one with 96 active rows, whose fourth row of warps skips the
       generator and the epilogues (1.75 tile-times instead of the 2 that whole tiles dealt round-robin cost).  The MMA
       always runs M = 128; the accumulator rows of inactive chains are never read. */
    const long long range_lo = (long long)blockIdx.x * p.chains_per_cta;
    const long long range_hi = range_lo + p.chains_per_cta < p.n_chains ? range_lo + p.chains_per_cta : p.n_chains;
    for (long long base = range_lo; base < range_hi; base += K4_TILE) {
        const int cnt = (int)(range_hi - base < K4_TILE ? range_hi - base : K4_TILE);      /* multiple of 32 */
        const bool act = m < cnt;                                                          /* warp-uniform */
        const long long ch = act ? base + m : base;
        const unsigned long long gch = p.chain_offset + (unsigned long long)ch;
        const unsigned c0 = (unsigned)gch, c1 = (unsigned)(gch >> 32);
        /* load the tile's state */
        if (act) {
#pragma unroll 4
            for (int jj = 0; jj < 16; jj++) {
                const int j = 16 * g + jj;
                S.xs[2 * j][m] = p.state[(long long)(K4_X + 1 + j) * ld + ch];
                S.xs[2 * j + 1][m] = p.state[(long long)(K4_X + 1 + K4_NC + j) * ld + ch];
            }
        }
        if (g == 0 && act) {
            S.a_s[m] = p.state[(long long)K4_X * ld + ch];
            S.e_s[m] = p.state[(long long)K4_E * ld + ch];
            S.sig_s[m] = p.state[(long long)K4_SIG * ld + ch];
            S.nacc_s[m] = p.state[(long long)K4_NACC * ld + ch];
            S.status_s[m] = (int)p.state[(long long)K4_STATUS * ld + ch];
            S.acc_s[m] = 0;
        }
        __syncthreads();

        /* One step = generator (Z tile + scalars) -> MMA -> epilogue 1 (energy statistics) -> decision -> epilogue 2.
           The phases are skewed so that the pipes overlap: the Z tile of step s+1 is generated right after the MMA of
           step s has finished with the operand buffer — in the same barrier interval as epilogue 1 of step s, so the
           integer-heavy generator and the FP64-heavy epilogue of different warps run side by side — and the MMA of step
           s+1 is issued before the decision and epilogue 2 of step s, which hide its latency.  Two barriers per step. */
        auto generate = [&](unsigned step, int slot, bool first) {
            if (act) {
                /* Z tile: 32 normals per thread, BF16, canonical K-major layout (16-byte chunk kc of row m at
                   kc*2048 + m*16: consecutive lanes write consecutive 16 B); one Philox call = 8 normals = one chunk */
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float dz[8];
                    const U4 r = philox(c0, c1, step, (unsigned)(4 * g + i), p.rk);
                    normal_pair_f32(r.x, dz[0], dz[1]);
                    normal_pair_f32(r.y, dz[2], dz[3]);
                    normal_pair_f32(r.z, dz[4], dz[5]);
                    normal_pair_f32(r.w, dz[6], dz[7]);
                    uint4 v;
                    v.x = pack_bf16(dz[0], dz[1]);
                    v.y = pack_bf16(dz[2], dz[3]);
                    v.z = pack_bf16(dz[4], dz[5]);
                    v.w = pack_bf16(dz[6], dz[7]);
                    *reinterpret_cast<uint4 *>(S.zs + (4 * g + i) * 2048 + m * 16) = v;
                    if (first && p.dbg_z != nullptr) {
#pragma unroll
                        for (int k = 0; k < 8; k++)
                            p.dbg_z[(long long)(32 * g + 8 * i + k) * ld + ch] = __bfloat162float(__float2bfloat16_rn(dz[k]));
                    }
                }
                if (g == 0) {        /* draws of the real parameter and of the accept test */
                    const U4 r = philox(c0, c1, step, 32u, p.rk);
                    float za, zb;
                    normal_pair_f32(r.x, za, zb);
                    S.za_s[slot][m] = (double)za;
                    S.u_s[slot][m] = u53(r.z, r.w);
                }
            }
        };
        auto issue_mma = [&]() {     /* Delta = Z . B^T on the tensor cores: 8 x (M128 N128 K16), accumulator in TMEM */
            if (warp == 0) {
                tc_fence_after();
                if (lane == 0) {
#pragma unroll
                    for (int k = 0; k < K4_N / 16; k++)
                        umma_bf16(tmem_d, umma_desc(zs_addr + k * 4096), umma_desc(ls_addr + k * 4096), K4_IDESC,
                                  k > 0 ? 1u : 0u);
                    umma_commit(&S.mbar);
                }
                __syncwarp();
            }
        };
        if (p.n_steps > 0) {
            generate((unsigned)p.step0, 0, true);
            fence_async_smem();          /* generic-proxy stores -> visible to the tensor-core (async) proxy */
            __syncthreads();
            issue_mma();
        }
        for (long long s = 0; s < p.n_steps; s++) {
            const unsigned step = (unsigned)(p.step0 + (unsigned long long)s);
            const int slot = (int)(s & 1);
            mbar_wait(&S.mbar, parity);
            parity ^= 1u;
            tc_fence_after();

            /* ---- epilogue 1: thread (m, g) owns 32 increments of chain m; partial energy statistics */
            uint32_t raw[32];
            if (act) tmem_ld32(tmem_d + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(32 * g), raw);
            tc_fence_before();
            if (s + 1 < p.n_steps) generate(step + 1u, slot ^ 1, false);     /* the MMA is done with the operand buffer */
            if (act && p.dbg_delta != nullptr && s == 0) {
#pragma unroll
                for (int k = 0; k < 32; k++) p.dbg_delta[(long long)(32 * g + k) * ld + ch] = __uint_as_float(raw[k]);
            }
            const double sig = S.sig_s[m];
            /* sum_j q_j^2 |c_j|^2 with q_j = q0 + jj: three sums with compile-time weights (1, jj, jj^2) and one
               combination per thread — no integer->double conversion per mode (XU pipe) */
            if (act) {
                double tot = 0.0, t1 = 0.0, t2 = 0.0;
#pragma unroll
                for (int jj = 0; jj < 16; jj++) {
                    const int j = 16 * g + jj;
                    const double re = fma(sig, f32_bits_to_f64(raw[2 * jj]), S.xs[2 * j][m]);
                    const double im = fma(sig, f32_bits_to_f64(raw[2 * jj + 1]), S.xs[2 * j + 1][m]);
                    const double m2 = fma(re, re, im * im);
                    tot += m2;
                    t1 = fma((double)jj, m2, t1);
                    t2 = fma((double)(jj * jj), m2, t2);
                }
                S.part[g][0][m] = tot;
                S.part[g][1][m] = fma(q0 * q0, tot, fma(2.0 * q0, t1, t2));
            }
            fence_async_smem();
            __syncthreads();             /* A: statistics ready, next Z tile visible, accumulator read by everyone */
            if (s + 1 < p.n_steps) issue_mma();

            /* ---- decision (one thread per chain): ME:247-258 */
            if (g == 0 && act) {
                const double t_all = (S.part[0][0][m] + S.part[1][0][m]) + (S.part[2][0][m] + S.part[3][0][m]);
                const double q_all = (S.part[0][1][m] + S.part[1][1][m]) + (S.part[2][1][m] + S.part[3][1][m]);
                const double a_new = fma(sig * s_a, S.za_s[slot][m], S.a_s[m]);
                bool accept = false;
                const bool wall = p.use_wall && fabs(a_new) >= 1.0;
                if (!wall) {
                    const double e_new = k4_energy(a_new, t_all, q_all, p);
                    if (e_new != e_new) S.status_s[m] |= ME_STATUS_ENERGY_NAN;
                    const double diff = e_new - S.e_s[m];
                    const double prob = me::exp_nonpos(fmin(-diff * p.inv_temp, 0.0), S.tables);
                    /* a NaN difference rejects, as in the reference (`uniform <= exp(nan)` is False, ME:327-338): fmin
                       would turn it into prob = 1 */
                    accept = (diff <= 0) | ((p.temp != 0) & (diff == diff) & (S.u_s[slot][m] <= prob));
                    if (accept) { S.e_s[m] = e_new; S.a_s[m] = a_new; S.nacc_s[m] += 1.0; }
                }
                const double sg = accept ? fma(sig, g_up, sig) : fma(sig, -g_down, sig);
                S.sig_s[m] = sg;
                if (!(sg > 0)) S.status_s[m] |= ME_STATUS_SIGMA_NONPOS;
                S.acc_s[m] = accept ? 1 : 0;
            }
            __syncthreads();             /* B */

            /* ---- epilogue 2: accepted chains take the increments (still in registers) */
            if (act && S.acc_s[m]) {
#pragma unroll
                for (int jj = 0; jj < 16; jj++) {
                    const int j = 16 * g + jj;
                    S.xs[2 * j][m] = fma(sig, f32_bits_to_f64(raw[2 * jj]), S.xs[2 * j][m]);
                    S.xs[2 * j + 1][m] = fma(sig, f32_bits_to_f64(raw[2 * jj + 1]), S.xs[2 * j + 1][m]);
                }
            }
        }

        /* store the tile's state */
        __syncthreads();
        if (act) {
#pragma unroll 4
            for (int jj = 0; jj < 16; jj++) {
                const int j = 16 * g + jj;
                p.state[(long long)(K4_X + 1 + j) * ld + ch] = S.xs[2 * j][m];
                p.state[(long long)(K4_X + 1 + K4_NC + j) * ld + ch] = S.xs[2 * j + 1][m];
            }
        }
        if (g == 0 && act) {
            p.state[(long long)K4_X * ld + ch] = S.a_s[m];
            p.state[(long long)K4_E * ld + ch] = S.e_s[m];
            p.state[(long long)K4_SIG * ld + ch] = S.sig_s[m];
            p.state[(long long)K4_NACC * ld + ch] = S.nacc_s[m];
            p.state[(long long)K4_STATUS * ld + ch] = (double)S.status_s[m];
            if (p.last_accept && p.n_steps > 0) p.last_accept[ch] = (unsigned char)S.acc_s[m];
        }
        __syncthreads();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, K4_N);
}

/* -------------------------------------------------------------------------------------------- init / measure */
__global__ void k4_init(K4Params p, const double *x0, int broadcast, double sigma0) {
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.n_chains) return;
    const long long ld = p.ld;
    double tot = 0.0, qsum = 0.0;
    const double a = broadcast ? x0[0] : x0[ch];
    p.state[(long long)K4_X * ld + ch] = a;
    p.state[(long long)K4_MEAN * ld + ch] = a;
    for (int j = 0; j < K4_NC; j++) {
        const double re = broadcast ? x0[1 + j] : x0[(long long)(1 + j) * ld + ch];
        const double im = broadcast ? x0[1 + K4_NC + j] : x0[(long long)(1 + K4_NC + j) * ld + ch];
        p.state[(long long)(K4_X + 1 + j) * ld + ch] = re;
        p.state[(long long)(K4_X + 1 + K4_NC + j) * ld + ch] = im;
        p.state[(long long)(K4_MEAN + 1 + j) * ld + ch] = re;
        p.state[(long long)(K4_MEAN + 1 + K4_NC + j) * ld + ch] = im;
        const double m2 = re * re + im * im, q = (double)(j - K4_NC / 2);
        tot += m2;
        qsum += q * q * m2;
        p.state[(long long)(K4_OBSM + 1 + j) * ld + ch] = hypot(re, im);
    }
    p.state[(long long)K4_OBSM * ld + ch] = fabs(a);
    p.state[(long long)(K4_OBSM + 1 + K4_NC) * ld + ch] = a * a;
    p.state[(long long)K4_E * ld + ch] = k4_energy(a, tot, qsum, p);
    p.state[(long long)K4_SIG * ld + ch] = sigma0;
    p.state[(long long)K4_NACC * ld + ch] = 0.0;
    p.state[(long long)K4_STATUS * ld + ch] = 0.0;
}

/* measure (ME:342-356 without the per-chain covariance, which is shared): running means (ME:404-410), observable
 * means (ME:412-414, 458-463), one time-series row [129 params, E, sigma].  n = counter after the increment. */
__global__ void k4_measure(K4Params p) {
    /* one thread per (slot, chain): slot j < 64 = complex mode j, slot 64 = real parameter + energy + sigma */
    const long long ch = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (ch >= p.n_chains) return;
    const long long ld = p.ld;
    const double dn = (double)p.n_meas, inv_n = 1.0 / dn, shrink = (dn - 1.0) * inv_n;
    double *row = p.record ? p.ts + p.ts_row * (long long)(K4_D + 2) * ld + ch : nullptr;
    if (j == K4_NC) {
        const double a = p.state[(long long)K4_X * ld + ch];
        double *mp = &p.state[(long long)K4_MEAN * ld + ch];
        *mp = *mp * shrink + a * inv_n;
        double *o0 = &p.state[(long long)K4_OBSM * ld + ch], *o1 = &p.state[(long long)(K4_OBSM + 1 + K4_NC) * ld + ch];
        *o0 = *o0 * shrink + fabs(a) * inv_n;
        *o1 = *o1 * shrink + (a * a) * inv_n;
        if (row) {
            __stcs(row, a);
            __stcs(row + (long long)K4_D * ld, p.state[(long long)K4_E * ld + ch]);
            __stcs(row + (long long)(K4_D + 1) * ld, p.state[(long long)K4_SIG * ld + ch]);
        }
        return;
    }
    const double re = p.state[(long long)(K4_X + 1 + j) * ld + ch];
    const double im = p.state[(long long)(K4_X + 1 + K4_NC + j) * ld + ch];
    double *mr = &p.state[(long long)(K4_MEAN + 1 + j) * ld + ch];
    double *mi = &p.state[(long long)(K4_MEAN + 1 + K4_NC + j) * ld + ch];
    *mr = *mr * shrink + re * inv_n;
    *mi = *mi * shrink + im * inv_n;
    double *ob = &p.state[(long long)(K4_OBSM + 1 + j) * ld + ch];
    *ob = *ob * shrink + hypot(re, im) * inv_n;
    if (row) {
        __stcs(row + (long long)(1 + j) * ld, re);
        __stcs(row + (long long)(1 + K4_NC + j) * ld, im);
    }
}

/* Pooled moments of the current states, deterministic two-stage reduction (no atomics: the covariance feeds the
 * proposals, so run-to-run bit reproducibility needs a fixed summation order).
 *
 * With Y = [Re c; Im c] (128 rows) about the shift, everything the complex second moment needs is in the LOWER triangle of
 * the real symmetric S = sum_chains Y Y^T (128 x 128):
 *     Re (c c^H)_ij = S[i][j] + S[64+i][64+j],    Im (c c^H)_ij = S[64+i][j] - S[64+j][i]
 * — a rank-k update with half the flops of the full complex outer product (8,256 instead of 16,384 FMAs per chain).
 * Stage 1: CTA b sums its slice of chains.  Chains are staged 32 at a time in shared memory as Ys[k][row]; thread t < 136
 *          owns the 8 x 8 register tile (ti, tj), tj <= ti, of S (64 independent FMA chains per thread: the FP64 pipe
 *          stays full with ~1 warp per sub-partition); threads 136..255 keep the column sums of Y.
 *          part[b]: [0] chains, [1] sum sigma, [2] sum a, [3] sum a^2, [4..132) sum Y, [132..) S row-major (lower part).
 * Stage 2: fixed-order sum over the CTAs (2a), emitted in the complex layout the host accumulates (2b):
 *          out[K4_MOMW] (double2): [0] chains, [1] sum sigma, [2] sum a, [3] sum a^2, [4..68) sum c, [68..) sum c c^H. */
constexpr int K4_MOMW = 4 + K4_NC + K4_NC * K4_NC;
constexpr int K4_MOM_CHUNK = 32;
constexpr int K4_PARTW = 4 + K4_N + K4_N * K4_N;        /* doubles per CTA partial */

__global__ void __launch_bounds__(256) k4_moments_stage1(K4Params p, const double *shift, double *part,
                                                         long long chains_per_cta) {
    /* Ys[k][pos(row)]: every block of 8 rows is followed by 2 pad doubles, so that the 16-byte reads of lanes that own
       neighbouring tiles (80 B apart) fall into distinct banks; the row length 162 keeps the staging stores (same
       row, consecutive k) at 4-way instead of 32-way conflicts. */
    constexpr int YLD = K4_N + 2 * (K4_N / 8) + 2;       /* 162 */
    __shared__ __align__(16) double Ys[K4_MOM_CHUNK][YLD];   /* 41 KB */
    auto pos = [](int row) { return row + 2 * (row >> 3); };
    __shared__ double red[256];
    const int tid = threadIdx.x;
    const long long lo = (long long)blockIdx.x * chains_per_cta;
    long long hi = lo + chains_per_cta;
    if (hi > p.n_chains) hi = p.n_chains;
    const long long ld = p.ld;
    /* tile of thread t < 136: row-major enumeration of the lower triangle of the 16 x 16 tile grid */
    int ti = 0, tj = 0;
    {
        int t = tid < 136 ? tid : 0;
        while (t > ti) { t -= ti + 1; ti++; }
        tj = t;
    }
    const bool tile_thread = tid < 136;
    double acc[8][8];
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) acc[a][b] = 0.0;
    double colsum = 0.0, colsum2 = 0.0;                   /* thread 136 + r: sum of Y[r] (and of Y[120 + r] for r < 8) */
    double sa = 0.0, sa2 = 0.0, ssig = 0.0;               /* threads < 32 */
    for (long long base = lo; base < hi; base += K4_MOM_CHUNK) {
        const int cnt = (int)((hi - base) < K4_MOM_CHUNK ? (hi - base) : K4_MOM_CHUNK);
        __syncthreads();
        /* stage: consecutive threads read consecutive chains of one state word (coalesced), write Ys[k][row] */
        for (int e = tid; e < K4_N * K4_MOM_CHUNK; e += 256) {
            const int row = e / K4_MOM_CHUNK, k = e % K4_MOM_CHUNK;
            Ys[k][pos(row)] = k < cnt ? p.state[(long long)(K4_X + 1 + row) * ld + base + k] - shift[1 + row] : 0.0;
        }
        if (tid < cnt) {
            const double a = p.state[(long long)K4_X * ld + base + tid] - shift[0];
            sa += a; sa2 += a * a; ssig += p.state[(long long)K4_SIG * ld + base + tid];
        }
        __syncthreads();
        if (tile_thread) {
#pragma unroll 4
            for (int k = 0; k < K4_MOM_CHUNK; k++) {
                double ya[8], yb[8];
                const double2 *pa = reinterpret_cast<const double2 *>(&Ys[k][10 * ti]);     /* pos(8 ti) */
                const double2 *pb = reinterpret_cast<const double2 *>(&Ys[k][10 * tj]);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const double2 va = pa[q], vb = pb[q];
                    ya[2 * q] = va.x; ya[2 * q + 1] = va.y; yb[2 * q] = vb.x; yb[2 * q + 1] = vb.y;
                }
#pragma unroll
                for (int a = 0; a < 8; a++)
#pragma unroll
                    for (int b = 0; b < 8; b++) acc[a][b] = fma(ya[a], yb[b], acc[a][b]);
            }
        } else {
            const int r = tid - 136;
            for (int k = 0; k < K4_MOM_CHUNK; k++) colsum += Ys[k][pos(r)];
            if (r < 8)
                for (int k = 0; k < K4_MOM_CHUNK; k++) colsum2 += Ys[k][pos(120 + r)];
        }
    }
    double *out = part + (long long)blockIdx.x * K4_PARTW;
    if (tile_thread) {
#pragma unroll
        for (int a = 0; a < 8; a++)
#pragma unroll
            for (int b = 0; b < 8; b++) out[4 + K4_N + (8 * ti + a) * K4_N + (8 * tj + b)] = acc[a][b];
    }
    if (!tile_thread) {
        out[4 + (tid - 136)] = colsum;
        if (tid - 136 < 8) out[4 + 120 + (tid - 136)] = colsum2;
    }
    /* the three scalar sums: fixed-order trees */
    for (int which = 0; which < 3; which++) {
        __syncthreads();
        red[tid] = (tid < K4_MOM_CHUNK) ? (which == 0 ? ssig : (which == 1 ? sa : sa2)) : 0.0;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) { if (tid < o) red[tid] += red[tid + o]; __syncthreads(); }
        if (tid == 0) out[1 + which] = red[0];
    }
    if (tid == 0) out[0] = (double)(hi > lo ? hi - lo : 0);
}

/* stage 2a: total[idx] = sum over the CTA partials in CTA order (one thread per word, coalesced across threads;
 * the upper triangle of S is never read, its threads idle) */
__global__ void k4_moments_stage2a(const double *part, int n_parts, double *total) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K4_PARTW) return;
    if (idx >= 4 + K4_N) {
        const int e = idx - 4 - K4_N;
        if (e / K4_N < e % K4_N) return;
    }
    /* loads in batches of 16 (independent), adds in CTA order (fixed summation order) */
    double t = 0.0;
    int b = 0;
    for (; b + 16 <= n_parts; b += 16) {
        double v[16];
#pragma unroll
        for (int q = 0; q < 16; q++) v[q] = part[(long long)(b + q) * K4_PARTW + idx];
#pragma unroll
        for (int q = 0; q < 16; q++) t += v[q];
    }
    for (; b < n_parts; b++) t += part[(long long)b * K4_PARTW + idx];
    total[idx] = t;
}
/* stage 2b: the complex layout the host accumulates.  Optionally (single-GPU fast path) the running moments are
 * advanced here, mom[w] += inc[w] for w != 1, and a snapshot [mom (MOMW) | inc[0], inc[1]] is written for a factor
 * refresh that runs asynchronously on another stream. */
__global__ void k4_moments_stage2b(const double *total, double2 *out, double2 *mom, double2 *snap) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= K4_MOMW) return;
    double2 v;
    if (w < 4) v = make_double2(total[w], 0.0);
    else if (w < 4 + K4_NC) { const int i = w - 4; v = make_double2(total[4 + i], total[4 + K4_NC + i]); }
    else {
        const int e = w - 4 - K4_NC, i = e / K4_NC, j = e % K4_NC;
        auto S = [&](int r, int c) { return total[4 + K4_N + (r >= c ? r * K4_N + c : c * K4_N + r)]; };   /* symmetric */
        v = make_double2(S(i, j) + S(K4_NC + i, K4_NC + j), S(K4_NC + i, j) - S(K4_NC + j, i));
    }
    out[w] = v;
    if (mom != nullptr) {
        double2 m = mom[w];
        if (w != 1) { m.x += v.x; m.y += v.y; mom[w] = m; }
        if (snap != nullptr) {
            snap[w] = m;
            if (w < 2) snap[K4_MOMW + w] = v;
        }
    }
}

/* 1 / sqrt(d) for a normal positive d: MUFU.RSQ64H seed + two Newton steps (relative error ~1e-16) */
__device__ __forceinline__ double k4_rsqrt(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y * y, 1.0);
    y = fma(0.5 * y, e, y);
    e = fma(-d, y * y, 1.0);
    return fma(0.5 * y, e, y);
}

/* Pooled covariance -> shared proposal factor, one CTA (runs once per measure after the 50th, ME:389,396).
 * mom (complex, as double pairs): [0] sample count N, [2] sum a, [3] sum a^2, [4..68) sum c, [68..) sum c c^H
 * (about a fixed shift); inc: [0] chains measured now, [1] sum of their sigma.  Computes
 *   C_c = (S2 - S1 S1^H / N)/(N-1) + small I,  small = mean(sigma)^2 / n   (the regulariser of ME:418,425),
 * its Cholesky factor G, the BF16 UMMA operand of me_k4_step, and the same for the real parameter.
 * Cholesky: left-looking by columns, 4 threads per row splitting the dot product (fixed order + shuffle tree), two
 * barriers per column, pivot through one reciprocal square root.  status: nonzero if a pivot was not positive. */
__global__ void __launch_bounds__(256) k4_refactor(const double2 *mom, const double2 *inc, long long n_meas,
                                                   double2 *cov_c, double *cov_a, __nv_bfloat16 *factor, double *s_a,
                                                   int *status) {
    constexpr int LDA = K4_NC + 1;                       /* padded row length (double2) */
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    double2 *A = reinterpret_cast<double2 *>(smem_raw);  /* A[i * LDA + j] */
    __shared__ double2 col[K4_NC];
    __shared__ int bad;
    const int tid = threadIdx.x, row = tid >> 2, part = tid & 3;
    const double N = mom[0].x;
    const double sm = inc[1].x / inc[0].x;
    const double small = sm * sm / (double)n_meas;
    const double inv_n = 1.0 / N, inv_n1 = 1.0 / (N - 1.0);
    const double2 *s1 = mom + 4, *s2 = mom + 4 + K4_NC;
    if (tid == 0) bad = 0;
    for (int e = tid; e < K4_NC * K4_NC; e += blockDim.x) {
        const int i = e / K4_NC, j = e % K4_NC;
        /* s1_i conj(s1_j) */
        const double pr = s1[i].x * s1[j].x + s1[i].y * s1[j].y, pi = s1[i].y * s1[j].x - s1[i].x * s1[j].y;
        double2 v;
        v.x = (s2[e].x - pr * inv_n) * inv_n1 + (i == j ? small : 0.0);
        v.y = (s2[e].y - pi * inv_n) * inv_n1;
        A[i * LDA + j] = v;
        cov_c[e] = v;
    }
    if (tid == 0) {
        const double va = (mom[3].x - mom[2].x * mom[2].x * inv_n) * inv_n1 + small;
        *cov_a = va;
        *s_a = sqrt(va);
    }
    __syncthreads();
    for (int j = 0; j < K4_NC; j++) {
        /* v_i = A_ij - sum_{k<j} G_ik conj(G_jk), rows i >= j; the four parts of a row are adjacent lanes */
        double ar = 0.0, ai = 0.0;
        if (row >= j) {
            for (int k = part; k < j; k += 4) {
                const double2 pq = A[row * LDA + k], q = A[j * LDA + k];
                ar = fma(pq.x, q.x, fma(pq.y, q.y, ar));
                ai = fma(pq.y, q.x, fma(-pq.x, q.y, ai));
            }
        }
        ar += __shfl_xor_sync(0xffffffffu, ar, 1); ai += __shfl_xor_sync(0xffffffffu, ai, 1);
        ar += __shfl_xor_sync(0xffffffffu, ar, 2); ai += __shfl_xor_sync(0xffffffffu, ai, 2);
        if (part == 0 && row >= j) {
            const double2 a0 = A[row * LDA + j];
            col[row] = make_double2(a0.x - ar, a0.y - ai);
        }
        __syncthreads();
        double d = col[j].x;
        if (!(d > 0.0)) { if (tid == 0) bad = 1; d = small > 0.0 ? small : 1e-300; }
        const double inv = k4_rsqrt(d);
        if (part == 0 && row >= j) {
            const double2 v = col[row];
            A[row * LDA + j] = row == j ? make_double2(d * inv, 0.0) : make_double2(v.x * inv, v.y * inv);
        }
        __syncthreads();
    }
    /* B[2i][2j] = Gr/sqrt2, B[2i][2j+1] = Gi/sqrt2, B[2i+1][2j] = -Gi/sqrt2, B[2i+1][2j+1] = Gr/sqrt2; stored
       BF16 at [k/8][n][k%8] */
    const double rs = 0.70710678118654752440;
    for (int e = tid; e < K4_N * K4_N; e += blockDim.x) {
        const int nrow = e / K4_N, k = e % K4_N;
        const int i = nrow >> 1, jj = k >> 1;
        double v = 0.0;
        if (jj <= i) {
            const double2 gij = A[i * LDA + jj];
            const bool ro = nrow & 1, ko = k & 1;
            v = (ro == ko) ? gij.x : (ro ? -gij.y : gij.y);
            if (jj == i && ro != ko) v = 0.0;       /* diagonal of G is real */
        }
        factor[(k >> 3) * (K4_N * 8) + nrow * 8 + (k & 7)] = __double2bfloat16(v * rs);
    }
    if (tid == 0 && status) *status = bad;
}

}  // namespace

/* ============================================================================================ C ABI */
struct me_k4 {
    me_k4_config cfg;
    double *state = nullptr;
    const void *factor = nullptr;
    unsigned char *last_accept = nullptr;
    long long n_measure = 1;
    unsigned long long step = 0;
    int n_sm = 148;
    int reserved_sms = 0;          /* SMs the step kernel leaves free (for a concurrent factor refresh) */
    std::string err;
};

static std::string g_k4_create_error;
static int k4_fail(me_k4 *e, int code, const std::string &msg) {
    if (e) e->err = msg; else g_k4_create_error = msg;
    return code;
}

static void k4_base(me_k4 *e, K4Params &p) {
    memset(&p, 0, sizeof(p));
    p.state = e->state;
    p.ld = e->cfg.n_chains;
    p.n_chains = e->cfg.n_chains;
    p.chain_offset = (unsigned long long)e->cfg.chain_offset;
    for (int r = 0; r < 10; r++) {
        p.rk[2 * r] = (unsigned)e->cfg.seed + (unsigned)r * 0x9E3779B9u;
        p.rk[2 * r + 1] = (unsigned)(e->cfg.seed >> 32) + (unsigned)r * 0xBB67AE85u;
    }
    p.step0 = e->step;
    p.n_meas = e->n_measure;
    p.temp = e->cfg.temp;
    p.inv_temp = e->cfg.temp != 0 ? 1.0 / e->cfg.temp : 0.0;
    p.target = e->cfg.target_acceptance;
    p.ratio = e->cfg.ratio;
    p.m = 1 + K4_NC;
    p.kappa = e->cfg.consts[0]; p.alpha = e->cfg.consts[1]; p.gamma = e->cfg.consts[2]; p.beta = e->cfg.consts[3];
    p.use_wall = e->cfg.use_reject;
    p.factor = e->factor;
    p.last_accept = e->last_accept;
}

extern "C" {

int me_k4_layout_get(me_k4_layout *o) {
    if (!o) return ME_ERR_INVALID;
    o->X = K4_X; o->E = K4_E; o->SIG = K4_SIG; o->MEAN = K4_MEAN; o->OBSM = K4_OBSM; o->NACC = K4_NACC;
    o->STATUS = K4_STATUS; o->WORDS = K4_WORDS; o->D = K4_D; o->TS_COLS = K4_D + 2; o->N_COMPLEX = K4_NC;
    o->TILE = K4_TILE; o->FACTOR_BYTES = K4_N * K4_N * 2; o->MOM_WORDS = K4_MOMW;
    o->MOM_SCRATCH_PER_SM = K4_PARTW;
    return ME_OK;
}

int me_k4_create(const me_k4_config *cfg, me_k4 **out) {
    if (!cfg || !out) return k4_fail(nullptr, ME_ERR_INVALID, "null argument");
    if (cfg->n_real != 1 || cfg->n_complex != K4_NC)
        return k4_fail(nullptr, ME_ERR_UNSUPPORTED, "the shared-covariance tensor-core path is built for 1 real + 64 complex parameters");
    if (cfg->n_chains <= 0 || cfg->n_chains % K4_TILE != 0)
        return k4_fail(nullptr, ME_ERR_INVALID, "n_chains must be a positive multiple of 128 (one MMA tile = 128 chains)");
    if (!(cfg->temp >= 0)) return k4_fail(nullptr, ME_ERR_INVALID, "temp must be >= 0 (reference: assert, ME:92)");
    me_k4 *e = new me_k4();
    e->cfg = *cfg;
    int n_sm = 0;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && n_sm > 0) e->n_sm = n_sm;
    else cudaGetLastError();
    *out = e;
    return ME_OK;
}

int me_k4_destroy(me_k4 *e) { delete e; return ME_OK; }

int me_k4_bind(me_k4 *e, double *state, const void *factor_bf16, unsigned char *last_accept) {
    if (!e || !state || !factor_bf16) return ME_ERR_INVALID;
    e->state = state; e->factor = factor_bf16; e->last_accept = last_accept;
    return ME_OK;
}

int me_k4_set_factor(me_k4 *e, const void *factor_bf16) {
    if (!e || !factor_bf16) return ME_ERR_INVALID;
    e->factor = factor_bf16;
    return ME_OK;
}

int me_k4_set_reserved_sms(me_k4 *e, int32_t n) {
    if (!e || n < 0) return ME_ERR_INVALID;
    e->reserved_sms = n;
    return ME_OK;
}

int me_k4_init(me_k4 *e, const double *x0, int32_t broadcast, double sigma0, void *stream) {
    if (!e || !e->state || !x0) return ME_ERR_INVALID;
    K4Params p;
    k4_base(e, p);
    int prev = -1; cudaGetDevice(&prev); cudaSetDevice(e->cfg.device);
    const int block = 128, grid = (int)((e->cfg.n_chains + block - 1) / block);
    k4_init<<<grid, block, 0, (cudaStream_t)stream>>>(p, x0, broadcast, sigma0);
    cudaError_t ce = cudaGetLastError();
    cudaSetDevice(prev);
    e->n_measure = 1; e->step = 0;
    if (ce != cudaSuccess) return k4_fail(e, ME_ERR_CUDA, std::string("k4_init: ") + cudaGetErrorString(ce));
    return ME_OK;
}

int me_k4_step(me_k4 *e, int64_t n_steps, const double *s_a, float *dbg_z, float *dbg_delta, void *stream) {
    if (!e || !e->state || !s_a) return ME_ERR_INVALID;
    if (n_steps <= 0) return ME_OK;
    if (e->step + (unsigned long long)n_steps >= 0xffffffffull) return k4_fail(e, ME_ERR_INVALID, "step counter overflow");
    K4Params p;
    k4_base(e, p);
    p.n_steps = n_steps; p.s_a = s_a; p.dbg_z = dbg_z; p.dbg_delta = dbg_delta;
    int prev = -1; cudaGetDevice(&prev); cudaSetDevice(e->cfg.device);
    /* the attribute is per device: set it before every launch (a host-side table write, no device work) */
    cudaError_t ce = cudaFuncSetAttribute(k4_steps, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(K4Smem));
    if (ce == cudaSuccess) {
        const int avail = e->n_sm - e->reserved_sms > 0 ? e->n_sm - e->reserved_sms : 1;
        long long per = (e->cfg.n_chains + avail - 1) / avail;
        per = (per + 31) / 32 * 32;
        p.chains_per_cta = per;
        const int grid = (int)((e->cfg.n_chains + per - 1) / per);
        k4_steps<<<grid, K4_THREADS, sizeof(K4Smem), (cudaStream_t)stream>>>(p);
        ce = cudaGetLastError();
    }
    cudaSetDevice(prev);
    if (ce != cudaSuccess) return k4_fail(e, ME_ERR_CUDA, std::string("k4_steps: ") + cudaGetErrorString(ce));
    e->step += (unsigned long long)n_steps;
    return ME_OK;
}

int me_k4_measure(me_k4 *e, double *ts, int64_t ts_row, void *stream) {
    if (!e || !e->state) return ME_ERR_INVALID;
    e->n_measure += 1;
    K4Params p;
    k4_base(e, p);
    p.ts = ts; p.ts_row = ts_row; p.record = ts != nullptr;
    int prev = -1; cudaGetDevice(&prev); cudaSetDevice(e->cfg.device);
    const int block = 256;
    const dim3 grid((unsigned)((e->cfg.n_chains + block - 1) / block), K4_NC + 1);
    k4_measure<<<grid, block, 0, (cudaStream_t)stream>>>(p);
    cudaError_t ce = cudaGetLastError();
    cudaSetDevice(prev);
    if (ce != cudaSuccess) return k4_fail(e, ME_ERR_CUDA, std::string("k4_measure: ") + cudaGetErrorString(ce));
    return ME_OK;
}

int me_k4_moments(me_k4 *e, const double *shift, double *scratch, int64_t scratch_doubles, double *inc, double *mom_accum,
                  double *snapshot, void *stream) {
    if (!e || !e->state || !shift || !scratch || !inc) return ME_ERR_INVALID;
    const int n_parts = e->n_sm < 1 ? 1 : e->n_sm;
    if (scratch_doubles < (int64_t)(n_parts + 1) * K4_PARTW) return k4_fail(e, ME_ERR_INVALID, "moments scratch too small");
    K4Params p;
    k4_base(e, p);
    const long long per = (e->cfg.n_chains + n_parts - 1) / n_parts;
    int prev = -1; cudaGetDevice(&prev); cudaSetDevice(e->cfg.device);
    k4_moments_stage1<<<n_parts, 256, 0, (cudaStream_t)stream>>>(p, shift, scratch, per);
    double *total = scratch + (long long)n_parts * K4_PARTW;
    k4_moments_stage2a<<<(K4_PARTW + 127) / 128, 128, 0, (cudaStream_t)stream>>>(scratch, n_parts, total);
    k4_moments_stage2b<<<(K4_MOMW + 127) / 128, 128, 0, (cudaStream_t)stream>>>(total, reinterpret_cast<double2 *>(inc),
                                                                             reinterpret_cast<double2 *>(mom_accum),
                                                                             reinterpret_cast<double2 *>(snapshot));
    cudaError_t ce = cudaGetLastError();
    cudaSetDevice(prev);
    if (ce != cudaSuccess) return k4_fail(e, ME_ERR_CUDA, std::string("k4_moments: ") + cudaGetErrorString(ce));
    return ME_OK;
}

int me_k4_refactor(me_k4 *e, const double *mom, const double *inc, int64_t n_measure, double *cov_c, double *cov_a,
                   void *factor_bf16, double *s_a, int32_t *status, void *stream) {
    if (!e || !mom || !inc || !cov_c || !cov_a || !factor_bf16 || !s_a) return ME_ERR_INVALID;
    if (n_measure <= 0) n_measure = e->n_measure;
    int prev = -1; cudaGetDevice(&prev); cudaSetDevice(e->cfg.device);
    const int smem = K4_NC * (K4_NC + 1) * (int)sizeof(double2);
    cudaError_t ce = cudaFuncSetAttribute(k4_refactor, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (ce == cudaSuccess) {
        k4_refactor<<<1, 256, smem, (cudaStream_t)stream>>>(reinterpret_cast<const double2 *>(mom),
                                                            reinterpret_cast<const double2 *>(inc), n_measure,
                                                            reinterpret_cast<double2 *>(cov_c), cov_a,
                                                            reinterpret_cast<__nv_bfloat16 *>(factor_bf16), s_a, status);
        ce = cudaGetLastError();
    }
    cudaSetDevice(prev);
    if (ce != cudaSuccess) return k4_fail(e, ME_ERR_CUDA, std::string("k4_refactor: ") + cudaGetErrorString(ce));
    return ME_OK;
}

int me_k4_get_counters(me_k4 *e, int64_t *n_measure, uint64_t *step) {
    if (!e) return ME_ERR_INVALID;
    if (n_measure) *n_measure = e->n_measure;
    if (step) *step = e->step;
    return ME_OK;
}

const char *me_k4_last_error(me_k4 *e) { return e ? e->err.c_str() : g_k4_create_error.c_str(); }

}  // extern "C"
