// Probe (experiment, not product): how many 16-CTA clusters are co-resident on a B200 for several CTA shapes, where the
// hardware puts them, and what an all-to-all push of a 64 KB tile through distributed shared memory costs per CTA when
// two CTAs of different clusters share an SM and the FP64 pipe is busy.  Informs the layout of k_wfd.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/build/probe_dsmem scripts/probe_dsmem.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned cl_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cl_size() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cl_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cl_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned mapa(unsigned a, unsigned rank) {
    unsigned r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cl(unsigned a, double2 v) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(v.x), "d"(v.y) : "memory");
}

constexpr int PITCH = 273;   // elements per line (256 + one pad per 16 + 1)
__device__ __forceinline__ int pad(int a) { return a + (a >> 4); }

// MODE 0: barriers only; 1: push to the 16 peers (transposition pattern); 2: same stores, all into my own tile;
// 3: push, each warp store instruction writes 512 contiguous bytes of one peer (best case)
template <int MODE, int NT>
__global__ void __launch_bounds__(NT, 512 / NT) k_push(int iters, int work, long long* cyc, unsigned* smids, double* sink) {
    extern __shared__ __align__(16) unsigned char smem[];
    double2* T = reinterpret_cast<double2*>(smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int me = (int)cl_rank();
    const int tr = lane & 15, line = 2 * warp + (lane >> 4);
    const int lines = NT / 16;
    for (int i = tid; i < lines * PITCH; i += NT) T[i] = make_double2(1.0 + i * 1e-9, 0.5);
    __syncthreads();
    cl_arrive(); cl_wait();
    if (tid == 0) { unsigned s; asm volatile("mov.u32 %0, %%smid;" : "=r"(s)); smids[blockIdx.x] = s; }
    const unsigned tbase = (unsigned)__cvta_generic_to_shared(T);
    cl_arrive();                                                 // plays the role of X2 of a previous iteration
    const long long t0 = clock64();
    double acc = 0;
    for (int it = 0; it < iters; ++it) {
        cl_wait();                                               // X2: data landed
        double2 v[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) v[q] = T[line * PITCH + pad(tr + 16 * q)];
        for (int k = 0; k < work; ++k) {
#pragma unroll
            for (int q = 0; q < 16; ++q) { v[q].x = fma(v[q].x, 0.999999, v[q].y); v[q].y = fma(v[q].y, 1.000001, -1e-9 * v[q].x); }
        }
        cl_arrive();                                             // X1: my tile is free
        cl_wait();
        if (MODE != 0) {
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                unsigned a;
                if (MODE == 3) a = tbase + (unsigned)(((warp * 16 + me) * 32 + lane) * 16);
                else a = tbase + (unsigned)((tr * PITCH + pad(16 * me + line)) * 16);     // line tr of CTA q, element 16 me + line
                const unsigned dst = (MODE == 2) ? (unsigned)me : (unsigned)q % cl_size();
                st_cl(mapa(a, dst), v[q]);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 16; ++q) acc += v[q].x;
        }
        cl_arrive();                                             // X2
    }
    cl_wait();
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = (t1 - t0) / iters;
    if (acc == 123.456) sink[0] = acc;
    cl_arrive(); cl_wait();
}

template <int MODE, int NT>
void run(const char* name, int csize, int iters, int work, int nclusters_cap) {
    auto kern = k_push<MODE, NT>;
    const size_t smem = (size_t)(NT / 16) * PITCH * 16 + (NT == 256 ? 40 * 1024 : 80 * 1024);   // tile + the room stash/tables take
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = smem; cfg.stream = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cfg.gridDim = dim3(csize * 64);
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    if (e != cudaSuccess) { printf("%-28s cluster %2d x %3d thr, smem %6zu: occupancy query failed: %s\n", name, csize, NT, smem, cudaGetErrorString(e)); cudaGetLastError(); return; }
    int ncl = n;
    if (nclusters_cap > 0 && ncl > nclusters_cap) ncl = nclusters_cap;
    if (ncl < 1) { printf("%-28s cluster %2d x %3d thr: 0 clusters fit\n", name, csize, NT); return; }
    cfg.gridDim = dim3(csize * ncl);
    long long* cyc; unsigned* smids; double* sink;
    CK(cudaMalloc(&cyc, sizeof(long long) * csize * ncl));
    CK(cudaMalloc(&smids, sizeof(unsigned) * csize * ncl));
    CK(cudaMalloc(&sink, 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaLaunchKernelEx(&cfg, kern, 10, work, cyc, smids, sink));   // warm-up
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    CK(cudaLaunchKernelEx(&cfg, kern, iters, work, cyc, smids, sink));
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(csize * ncl); std::vector<unsigned> hs(csize * ncl);
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hs.data(), smids, sizeof(unsigned) * hs.size(), cudaMemcpyDeviceToHost));
    long long mn = h[0], mx = h[0]; double avg = 0;
    for (auto c : h) { mn = std::min(mn, c); mx = std::max(mx, c); avg += (double)c; }
    avg /= h.size();
    std::vector<unsigned> u(hs); std::sort(u.begin(), u.end()); u.erase(std::unique(u.begin(), u.end()), u.end());
    printf("%-28s cluster %2d x %3d thr, smem %6zu B: max active clusters %2d, ran %2d on %3zu SMs | work %3d | cycles/iter min %lld avg %.0f max %lld | %.3f us/iter\n",
           name, csize, NT, smem, n, ncl, u.size(), work, mn, avg, mx, 1e3 * ms / iters);
    if (MODE == 0 && work == 0 && nclusters_cap == 0) {
        for (int c = 0; c < ncl; ++c) {
            printf("   cluster %2d SMs:", c);
            std::vector<unsigned> s(hs.begin() + c * csize, hs.begin() + (c + 1) * csize); std::sort(s.begin(), s.end());
            for (auto x : s) printf(" %u", x);
            printf("\n");
        }
    }
    cudaFree(cyc); cudaFree(smids); cudaFree(sink);
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("%s: %d SMs, %zu B smem/SM, %zu B smem/block optin, clock %d kHz\n", pr.name, pr.multiProcessorCount, pr.sharedMemPerMultiprocessor,
           pr.sharedMemPerBlockOptin, pr.clockRate);
    const int it = 2000;
    run<0, 256>("barriers only", 16, it, 0, 0);
    run<0, 256>("barriers only", 8, it, 0, 0);
    run<0, 512>("barriers only, 512 thr", 8, it, 0, 0);
    run<0, 512>("barriers only, 512 thr", 16, it, 0, 0);
    run<0, 256>("barriers only", 4, it, 0, 0);
    run<0, 256>("barriers only", 2, it, 0, 0);
    run<1, 256>("push (transposition)", 16, it, 0, 0);
    run<2, 256>("same stores, own tile", 16, it, 0, 0);
    run<3, 256>("push, 512 B contiguous", 16, it, 0, 0);
    run<1, 256>("push, ONE cluster", 16, it, 0, 1);
    run<1, 256>("push, 7 clusters", 16, it, 0, 7);
    // with FP64 work: 16 points x `work` x 2 DFMA per thread and iteration (a half step of the SSFM is ~85 DFMA per point = work 42)
    for (int w : {10, 20, 42, 64}) {
        run<0, 256>("barriers + fp64 work", 16, it / 4, w, 0);
        run<1, 256>("push + fp64 work", 16, it / 4, w, 0);
    }
    run<1, 256>("push + fp64 work, 7 clusters", 16, it / 4, 42, 7);
    run<0, 256>("work only, 7 clusters", 16, it / 4, 42, 7);
    return 0;
}
