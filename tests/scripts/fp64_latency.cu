// Microbenchmark: dependent-issue latency and per-warp throughput of DFMA / IMAD.WIDE / LOP3 chains on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void dfma_chain(double *out, long long iters, long long *cycles) {
    double a[ILP];
    for (int i = 0; i < ILP; i++) a[i] = threadIdx.x * 1e-3 + i;
    const double m = 0.99999988, c = 1.25e-7;
    long long t0 = clock64();
    for (long long it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) a[i] = fma(a[i], m, c);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < ILP; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
__global__ void imad_chain(unsigned *out, long long iters, long long *cycles) {
    unsigned a = threadIdx.x + 1, b = 77;
    long long t0 = clock64();
    for (long long it = 0; it < iters; it++) {
        unsigned long long p = (unsigned long long)a * 0xD2511F53u;
        a = (unsigned)(p >> 32) ^ (unsigned)p ^ b;
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}
template <int ILP> void run(int warps_per_sm) {
    double *out; long long *cyc; cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
    long long iters = 20000;
    dfma_chain<ILP><<<148, 32 * warps_per_sm>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    dfma_chain<ILP><<<148, 32 * warps_per_sm>>>(out, iters, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA ILP=%d warps/SM=%2d: %.2f cycles per iteration (%.2f per DFMA)\n", ILP, warps_per_sm, (double)h / iters, (double)h / iters / ILP);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<1>(1); run<2>(1); run<4>(1); run<8>(1);
    run<1>(4); run<2>(4); run<4>(4); run<1>(16); run<2>(16); run<4>(16); run<8>(16);
    unsigned *o; long long *cyc; cudaMalloc(&o, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    imad_chain<<<148, 32>>>(o, 20000, cyc); cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("IMAD.WIDE+LOP3 dependent round: %.2f cycles\n", (double)h / 20000);
    return 0;
}
