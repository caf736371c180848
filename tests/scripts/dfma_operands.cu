// Microbenchmark: does a warp-wide DFMA cost more than 2 FP64-pipe cycles when its operands are three distinct
// 64-bit REGISTERS (six 32-bit register reads) rather than one register and two immediates / constant-bank values?
// 8 independent chains per warp, 16 warps per SM (4 per sub-partition).
#include <cstdio>
#include <cuda_runtime.h>
template <int FORM>
__global__ void k(double *out, const double *in, long long iters, long long *cycles) {
    double a[8], b[8], c[8];
    for (int i = 0; i < 8; i++) { a[i] = threadIdx.x * 1e-3 + i; b[i] = in[i] - threadIdx.x * 1e-12; c[i] = in[8 + i] + threadIdx.x * 1e-13; }   // per-lane values: vector registers
    const double u = in[3];                                                // warp-uniform: lands in a uniform register
    long long t0 = clock64();
    for (long long it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (FORM == 0) a[i] = fma(a[i], 0.99999988, 1.25e-7);          // register, immediate/constant, constant
            else if (FORM == 1) a[i] = fma(a[i], b[i], c[i]);              // three distinct registers per chain
            else if (FORM == 2) a[i] = fma(a[i], b[0], c[0]);              // three registers, two shared by all chains
            else if (FORM == 3) a[i] = fma(b[i], c[i], a[i]);              // accumulate form (dot product)
            else if (FORM == 4) a[i] = fma(a[i], b[i], 1.25e-7);           // two registers + immediate (Horner step)
            else if (FORM == 5) a[i] = fma(a[i], b[i], u);                 // two registers + uniform register
            else a[i] = a[i] * b[i];                                       // DMUL, two registers
        }
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int FORM> void run(const char *name) {
    double *out, *in; long long *cyc; cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&in, 16 * 8); cudaMalloc(&cyc, 8);
    double h_in[16]; for (int i = 0; i < 16; i++) h_in[i] = i < 8 ? 0.99999988 - i * 1e-9 : 1.25e-7 + i * 1e-10;
    cudaMemcpy(in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    long long iters = 20000, h;
    for (int r = 0; r < 2; r++) { k<FORM><<<148, 512>>>(out, in, iters, cyc); cudaDeviceSynchronize(); }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-46s %7.2f cycles per 8 DFMA per warp -> %.2f cycles per DFMA per sub-partition\n", name, (double)h / iters,
           (double)h / iters / 4.0 / 8.0);
    cudaFree(out); cudaFree(in); cudaFree(cyc);
}
int main() {
    run<0>("fma(a, imm, imm)");
    run<1>("fma(a, b_i, c_i)   three registers");
    run<2>("fma(a, b_0, c_0)   shared multiplier/addend");
    run<3>("fma(b_i, c_i, a)   accumulate form");
    run<4>("fma(a, b_i, imm)   two registers + immediate");
    run<5>("fma(a, b_i, UR)    two registers + uniform reg");
    run<6>("a * b_i            DMUL, two registers");
    return 0;
}
