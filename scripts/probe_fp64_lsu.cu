// Does FP64 arithmetic overlap with shared-memory traffic on one SM?  (B200, sm_100a)
//   mode 0: every warp runs ITER x 64 DFMA (16 independent chains)
//   mode 1: every warp runs ITER x NL 128-bit shared-memory accesses (half loads, half stores, conflict-free)
//   mode 2: both, interleaved in the same instruction stream, no dependency between the two
//   mode 3: even warps run mode 0, odd warps mode 1 (same per-warp counts)
// 2 CTAs x 256 threads per SM (the geometry of k_wf<double>), all SMs.  Prints cycles per iteration.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_fp64_lsu probe_fp64_lsu.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int NL>
__global__ void __launch_bounds__(256, 2) k(double* out, long long* cyc, int iters, double seed) {
    extern __shared__ __align__(16) double2 sm[];
    const int tid = threadIdx.x;
    double a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + i + tid;
    const double m = 1.0000001, c = 1e-9;
    double2 v[NL / 2];
#pragma unroll
    for (int i = 0; i < NL / 2; ++i) { v[i].x = tid + i; v[i].y = tid - i; }
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm) + tid * 16;
    const bool do_f = MODE == 0 || MODE == 2 || (MODE == 3 && ((tid >> 5) & 1) == 0);
    const bool do_l = MODE == 1 || MODE == 2 || (MODE == 3 && ((tid >> 5) & 1) == 1);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (do_l) {
#pragma unroll
            for (int i = 0; i < NL / 2; ++i)
                asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(base + i * 4096), "d"(v[i].x), "d"(v[i].y) : "memory");
        }
        if (do_f) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
        }
        if (do_l) {
#pragma unroll
            for (int i = 0; i < NL / 2; ++i)
                asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v[i].x), "=d"(v[i].y) : "r"(base + i * 4096) : "memory");
        }
        if (do_f) {
#pragma unroll
            for (int r = 0; r < 2; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
    for (int i = 0; i < NL / 2; ++i) s += v[i].x + v[i].y;
    out[blockIdx.x * 256 + tid] = s;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE, int NL> void run(const char* name, double* out, long long* cyc, int blocks, int iters) {
    cudaFuncSetAttribute(k<MODE, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<MODE, NL><<<blocks, 256, 65536>>>(out, cyc, iters, 1.0);
    cudaDeviceSynchronize();
    k<MODE, NL><<<blocks, 256, 65536>>>(out, cyc, iters, 1.0);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i];
    printf("%-58s NL=%2d: %8.1f cycles/iter  (%s)\n", name, NL, avg / blocks / iters, cudaGetErrorString(e));
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024 * 256); cudaMalloc(&cyc, 8 * 1024);
    const int blocks = 2 * sms, iters = 2000;
    printf("%d SMs, %d CTAs of 256 threads, per iteration and thread: 64 DFMA and/or NL 128-bit shared-memory accesses\n", sms, blocks);
    run<0, 8>("mode 0: DFMA only", out, cyc, blocks, iters);
    run<1, 8>("mode 1: shared memory only", out, cyc, blocks, iters);
    run<2, 8>("mode 2: both in every warp", out, cyc, blocks, iters);
    run<3, 8>("mode 3: even warps DFMA, odd warps shared memory", out, cyc, blocks, iters);
    run<1, 16>("mode 1: shared memory only", out, cyc, blocks, iters);
    run<2, 16>("mode 2: both in every warp", out, cyc, blocks, iters);
    run<3, 16>("mode 3: even warps DFMA, odd warps shared memory", out, cyc, blocks, iters);
    run<1, 32>("mode 1: shared memory only", out, cyc, blocks, iters);
    run<2, 32>("mode 2: both in every warp", out, cyc, blocks, iters);
    return 0;
}
