// Welch power spectral density of a batch of rows on the device.
//
// Replaces scipy.signal.welch as the reference calls it for `signal.psd()` (opticomlib/typing.py:1899-1902) and
// `utils.get_psd` (opticomlib/utils.py:2074-2079): segments of nperseg samples with 50 % overlap, periodic Hann window,
// no detrending, scaling='spectrum' (1 / (sum w)^2), two-sided, mean over the segments, then fftshift of the bins.
// The transforms are the in-register / shared-memory Stockham passes of fft_core.cuh (one nperseg-point transform per CTA
// pass, 16 points per thread); a CTA walks several segments of one row, accumulates |X_k|^2 in registers and adds its share
// to the row's spectrum once.  HBM traffic: every sample is read twice (overlap), the output is nperseg doubles per row.
#include <cmath>
#include <mutex>
#include <string>

#include "../../include/ssfm_b200.h"
#include "fft_core.cuh"
#include "ssfm_internal.h"

namespace ssfm_spec {
using namespace ssfm;

template <int M>
__global__ void __launch_bounds__(M / 16) k_welch(const double2* __restrict__ x, double* __restrict__ psd, const double2* __restrict__ tw_g,
                                                   long long n, int nseg, int seg_per_cta, double scale) {
    constexpr int E = 16, NT = M / E;
    typedef RowExchange<M, E> X;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* sm = reinterpret_cast<double2*>(smem_raw);                          // exchange buffer (padded)
    double2* tw = sm + X::size;                                                  // pass tables
    const int t = threadIdx.x;
    const long long row = blockIdx.y;
    for (int i = t; i < fft_plan<M, E>::table_size; i += NT) tw[i] = tw_g[i];
    double win[E], acc[E];
#pragma unroll
    for (int q = 0; q < E; ++q) {
        win[q] = 0.5 - 0.5 * cospi(2.0 * (double)(t + q * NT) / (double)M);      // scipy.signal.get_window('hann', M): periodic
        acc[q] = 0.0;
    }
    __syncthreads();
    const double2* xr = x + row * n;
    const int s0 = blockIdx.x * seg_per_cta, s1 = min(nseg, s0 + seg_per_cta);
    for (int s = s0; s < s1; ++s) {
        const double2* seg = xr + (long long)s * (M / 2);
        double2 v[E];
#pragma unroll
        for (int q = 0; q < E; ++q) { const double2 a = seg[t + q * NT]; v[q].x = a.x * win[q]; v[q].y = a.y * win[q]; }
        fft_passes<double, M, -1, X, E>::run(v, sm, tw, t);
#pragma unroll
        for (int q = 0; q < E; ++q) acc[q] += v[q].x * v[q].x + v[q].y * v[q].y;
        X::sync();                                                              // the exchange buffer is reused by the next segment
    }
    if (s1 > s0) {
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int k = t + q * NT;                                           // bin k -> fftshift position (k + M/2) mod M
            atomicAdd(psd + row * M + ((k + M / 2) & (M - 1)), acc[q] * scale);
        }
    }
}

struct Tables { void* tw[12] = {nullptr}; };
Tables& tables_of(int device) { static Tables t[64]; return t[(device >= 0 && device < 64) ? device : 0]; }
std::mutex g_mu;

template <int M>
int launch(const double2* x, double* psd, long long rows, long long n, int device, cudaStream_t st) {
    int lg = 0; while ((1 << lg) < M) ++lg;
    void*& tw = tables_of(device).tw[lg - 4];
    {
        std::lock_guard<std::mutex> lock(g_mu);
        if (!tw) { const int rc = ssfm_internal_pass_tables_f64(&tw, M, st); if (rc) return rc; }
    }
    const int nseg = (int)((n - M / 2) / (M / 2));
    double wsum = 0.0;
    for (int i = 0; i < M; ++i) wsum += 0.5 - 0.5 * std::cos(2.0 * 3.14159265358979323846 * i / M);
    const double scale = 1.0 / (wsum * wsum) / nseg;
    cudaError_t e = cudaMemsetAsync(psd, 0, sizeof(double) * (size_t)rows * M, st);
    if (e == cudaSuccess) {
        int per = 1;                                                            // enough CTAs to fill the chip, few atomics otherwise
        while ((long long)rows * ((nseg + per - 1) / per) > 4096 && per < nseg) per *= 2;
        dim3 grid((unsigned)((nseg + per - 1) / per), (unsigned)rows);
        const size_t smem = sizeof(double2) * (size_t)(RowExchange<M, 16>::size + fft_plan<M, 16>::table_size);
        static bool attr[64] = {false};
        if (!attr[device & 63]) { cudaFuncSetAttribute(k_welch<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr[device & 63] = true; }
        k_welch<M><<<grid, M / 16, smem, st>>>(x, psd, (const double2*)tw, n, nseg, per, scale);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { ssfm_err_slot = std::string("welch: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }
    return SSFM_OK;
}

}  // namespace ssfm_spec

extern "C" int ssfm_welch_psd(const void* x_dev, double* psd_dev, int64_t n_rows, int64_t n, int32_t nperseg, int32_t device,
                              void* stream) {
    using namespace ssfm_spec;
    if (!x_dev || !psd_dev) { ssfm_err_slot = "null buffer"; return SSFM_ERR_INVALID; }
    if (n_rows < 1 || n_rows > 65535 || n < nperseg) { ssfm_err_slot = "welch: 1 <= n_rows <= 65535 and n_samples >= nperseg expected"; return SSFM_ERR_INVALID; }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { ssfm_err_slot = std::string("welch: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)stream;
    const double2* x = (const double2*)x_dev;
    switch (nperseg) {
        case 256: return launch<256>(x, psd_dev, n_rows, n, device, st);
        case 512: return launch<512>(x, psd_dev, n_rows, n, device, st);
        case 1024: return launch<1024>(x, psd_dev, n_rows, n, device, st);
        case 2048: return launch<2048>(x, psd_dev, n_rows, n, device, st);
        default: ssfm_err_slot = "welch: nperseg must be 256, 512, 1024 or 2048 (the reference uses min(2048, n_samples))"; return SSFM_ERR_UNSUPPORTED;
    }
}
