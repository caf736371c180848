// Microbenchmark: do FP64 and integer instructions share dispatch cycles on sm_100a?
// Each warp runs ND independent DFMA chains and NI independent integer chains (IMAD.WIDE or LOP3) per iteration,
// 4 or 16 warps per SM.  If a warp-wide DFMA blocks its sub-partition's dispatch port for 2 cycles, the cost per
// iteration per SMSP is 2*ND + NI (x warps per SMSP); if only the FP64 pipe is busy for 2 cycles and the port is
// free for other pipes, it is max(2*ND, ND + NI).
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NI, int KIND>
__global__ void mix(double *out, long long iters, long long *cycles) {
    double a[ND > 0 ? ND : 1];
    unsigned b[NI > 0 ? NI : 1];
    for (int i = 0; i < ND; i++) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < NI; i++) b[i] = threadIdx.x * 77u + i;
    const double m = 0.99999988, c = 1.25e-7;
    long long t0 = clock64();
    for (long long it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < (ND > NI ? ND : NI); i++) {
            if (i < ND) a[i] = fma(a[i], m, c);
            if (i < NI) {
                if (KIND == 0) { unsigned long long p = (unsigned long long)b[i] * 0xD2511F53u; b[i] = (unsigned)(p >> 32) + (unsigned)p; }
                else if (KIND == 1) b[i] = (b[i] ^ 0x9E3779B9u) + (b[i] >> 3);          // LOP3 + SHF/IADD
                else if (KIND == 2) b[i] = b[i] * 0xCD9E8D57u + 12345u;                 // IMAD (32-bit), affine: ptxas composes it
                else b[i] = b[i] * b[i] + 12345u;                                       // IMAD (32-bit), quadratic: stays in the loop
            }
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ND; i++) s += a[i];
    for (int i = 0; i < NI; i++) s += (double)b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int ND, int NI, int KIND> void run(int warps_per_sm) {
    double *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    long long iters = 20000;
    for (int r = 0; r < 2; r++) { mix<ND, NI, KIND><<<148, 32 * warps_per_sm>>>(out, iters, cyc); cudaDeviceSynchronize(); }
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const char *kn = KIND == 0 ? "IMAD.WIDE+IADD" : (KIND == 1 ? "LOP3+SHF+IADD" : (KIND == 2 ? "IMAD32" : "IMAD32 x*x+c"));
    printf("ND=%2d DFMA + NI=%2d %-15s warps/SM=%2d: %7.2f cycles/iter/warp -> %.2f cycles per SMSP-iteration\n", ND, NI, kn,
           warps_per_sm, (double)h / iters, (double)h / iters / (warps_per_sm / 4.0) );
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<8, 0, 2>(16); run<0, 8, 2>(16); run<8, 8, 2>(16); run<8, 16, 2>(16); run<4, 16, 2>(16);
    run<0, 8, 0>(16); run<8, 8, 0>(16);
    run<0, 8, 1>(16); run<8, 8, 1>(16);
    run<8, 8, 2>(8); run<8, 8, 2>(12);
    run<0, 16, 3>(16); run<8, 16, 3>(16); run<8, 8, 3>(16); run<8, 32, 3>(16); run<0, 32, 3>(16);
    return 0;
}
