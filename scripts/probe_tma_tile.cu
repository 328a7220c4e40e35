// What does it cost an SM to move a 64 KB tile L2 -> registers -> L2, per thread with ld.global.cg / st.global (the way k_wf
// does it), against one bulk asynchronous copy (TMA, cp.async.bulk) into shared memory + ld.shared / st.shared + a bulk copy
// back?  Two CTAs of 256 threads per SM, every CTA owns a tile of 4096 complex128 samples of an L2-resident buffer.
//   variant 0: LSU      row tile (64 KB contiguous)            variant 1: LSU      column tile (256 segments of 256 B, pitch 4 KB)
//   variant 2: bulk copy row tile (one 64 KB copy each way)     variant 3: bulk copy column tile (256 copies of 256 B each way)
// each with WORK = 0 / 512 dependent-free DFMA per thread and iteration between the load and the store.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/probe_tma_tile probe_tma_tile.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ unsigned sa(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sa(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(unsigned long long* b, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sa(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned par) {
    asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(sa(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sa(dst)), "l"(src), "r"(bytes), "r"(sa(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(sa(src)), "r"(bytes) : "memory");
}
template <int VAR, int WORK>
__global__ void __launch_bounds__(256, 2) k(double2* buf, long long* cyc, int iters) {
    extern __shared__ __align__(128) double2 sm[];
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    constexpr bool COL = (VAR & 1) != 0, TMA = VAR >= 2;
    // row tile: 4096 contiguous samples; column tile: rows of 16 samples (256 B) at a pitch of 256 samples inside a 1 MiB waveform
    double2* base = COL ? buf + (size_t)(blockIdx.x / 16) * 65536 + (blockIdx.x % 16) * 16 : buf + (size_t)blockIdx.x * 4096;
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    double2 v[16];
    const double m = 1.0000001, c = 1e-9;
    unsigned par = 0;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (!TMA) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const double2* p = COL ? base + (size_t)(tid / 16 + q * 16) * 256 + (tid % 16) : base + tid + q * 256;
                v[q] = __ldcg(p);
            }
        } else {
            if (COL) {
                if (tid == 0) mbar_expect(&bar, 65536);
                __syncwarp();
                bulk_g2s(sm + tid * 16, base + (size_t)tid * 256, 256, &bar);
            } else if (tid == 0) {
                mbar_expect(&bar, 65536);
                bulk_g2s(sm, base, 65536, &bar);
            }
            mbar_wait(&bar, par); par ^= 1;
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = sm[COL ? (tid / 16 + q * 16) * 16 + (tid % 16) : tid + q * 256];
        }
#pragma unroll
        for (int w = 0; w < WORK / 32; ++w)
#pragma unroll
            for (int q = 0; q < 16; ++q) { v[q].x = fma(v[q].x, m, c); v[q].y = fma(v[q].y, m, c); }
        if (!TMA) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                double2* p = COL ? base + (size_t)(tid / 16 + q * 16) * 256 + (tid % 16) : base + tid + q * 256;
                *p = v[q];
            }
            __threadfence();                                       // (k_wf: the release before the team barrier)
            __syncthreads();
        } else {
            __syncthreads();                                       // everyone has read the landing buffer
#pragma unroll
            for (int q = 0; q < 16; ++q) sm[COL ? (tid / 16 + q * 16) * 16 + (tid % 16) : tid + q * 256] = v[q];
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            if (COL) bulk_s2g(base + (size_t)tid * 256, sm + tid * 16, 256);
            else if (tid == 0) bulk_s2g(base, sm, 65536);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            __syncthreads();
        }
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int VAR, int WORK> void run(const char* name, double2* buf, long long* cyc, int blocks, int iters) {
    cudaFuncSetAttribute(k<VAR, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    for (int rep = 0; rep < 2; ++rep) { k<VAR, WORK><<<blocks, 256, 65536>>>(buf, cyc, iters); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    long long h[1024];
    cudaMemcpy(h, cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i];
    printf("%-44s DFMA/thread %4d: %8.1f cycles per tile round trip  (%s)\n", name, WORK, avg / blocks / iters, cudaGetErrorString(e));
}
int main() {
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = (2 * sms / 16) * 16, iters = 300;
    double2* buf; long long* cyc;
    cudaMalloc(&buf, (size_t)blocks * 4096 * 16); cudaMemset(buf, 0, (size_t)blocks * 4096 * 16); cudaMalloc(&cyc, 8 * 1024);
    printf("%d SMs, %d CTAs x 256 threads (2 per SM), tile = 4096 complex128 (64 KB) in and out per iteration\n", sms, blocks);
    run<0, 0>("LSU  row tile", buf, cyc, blocks, iters);
    run<1, 0>("LSU  column tile", buf, cyc, blocks, iters);
    run<2, 0>("bulk row tile (1 x 64 KB)", buf, cyc, blocks, iters);
    run<3, 0>("bulk column tile (256 x 256 B)", buf, cyc, blocks, iters);
    run<0, 512>("LSU  row tile", buf, cyc, blocks, iters);
    run<1, 512>("LSU  column tile", buf, cyc, blocks, iters);
    run<2, 512>("bulk row tile (1 x 64 KB)", buf, cyc, blocks, iters);
    run<3, 512>("bulk column tile (256 x 256 B)", buf, cyc, blocks, iters);
    run<0, 2048>("LSU  row tile", buf, cyc, blocks, iters);
    run<1, 2048>("LSU  column tile", buf, cyc, blocks, iters);
    run<2, 2048>("bulk row tile (1 x 64 KB)", buf, cyc, blocks, iters);
    run<3, 2048>("bulk column tile (256 x 256 B)", buf, cyc, blocks, iters);
    return 0;
}
