// Zero-phase cascaded-biquad filtering on the device (LPF / BPF hot path).
//
// Replaces scipy.signal.sosfiltfilt as called by the reference at opticomlib/devices.py:820-823
// (BPF) and 1365-1368 (LPF): odd extension by `edge` samples, steady-state initial conditions
// scaled by the first sample, forward cascade, backward cascade, strip the extension.
//
// v0 kernel: one thread per (row, real|imag component) walks the recurrence sequentially.  It is
// exact (same operation order as the SciPy loop, FMA contraction disabled) and works for any N;
// the FFT-domain fast path for power-of-two rows reuses the SSFM transforms (see filt_fft.cu).
#include <cmath>
#include <cstring>
#include <string>

#include "../../include/ssfm_b200.h"
#include <cuda_runtime.h>

namespace ssfm_filt {

constexpr int MAX_SECTIONS = 8;

struct Sos {
    double c[MAX_SECTIONS][6];
    double zi[MAX_SECTIONS][2];
    int n_sections;
    int edge;
};

__device__ __forceinline__ double ext_at(const double* x, long long n, int edge, long long i) {
    // odd extension: 2*x[0] - x[edge-i] | x | 2*x[n-1] - x[n-2-m]   (component stride 2: interleaved complex)
    if (i < edge) return 2.0 * x[0] - x[2 * (edge - i)];
    i -= edge;
    if (i < n) return x[2 * i];
    i -= n;
    return 2.0 * x[2 * (n - 1)] - x[2 * (n - 2 - i)];
}

__device__ __forceinline__ double cascade(const Sos& f, double v, double (&z)[MAX_SECTIONS][2]) {
#pragma unroll 1
    for (int s = 0; s < f.n_sections; ++s) {
        const double y = __dadd_rn(__dmul_rn(f.c[s][0], v), z[s][0]);
        z[s][0] = __dadd_rn(__dsub_rn(__dmul_rn(f.c[s][1], v), __dmul_rn(f.c[s][4], y)), z[s][1]);
        z[s][1] = __dsub_rn(__dmul_rn(f.c[s][2], v), __dmul_rn(f.c[s][5], y));
        v = y;
    }
    return v;
}

// x, y: [rows][n] complex128 (interleaved); ws: [rows*2][n + 2*edge] doubles
__global__ void k_filtfilt_seq(const double* __restrict__ x, double* __restrict__ y, double* __restrict__ ws,
                               long long rows, long long n, Sos f) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= rows * 2) return;
    const long long row = tid >> 1;
    const int comp = (int)(tid & 1);
    const double* xr = x + row * n * 2 + comp;
    double* yr = y + row * n * 2 + comp;
    const long long len = n + 2 * (long long)f.edge;
    double* w = ws + tid * len;

    double z[MAX_SECTIONS][2];
    const double x0 = ext_at(xr, n, f.edge, 0);
    for (int s = 0; s < f.n_sections; ++s) { z[s][0] = __dmul_rn(f.zi[s][0], x0); z[s][1] = __dmul_rn(f.zi[s][1], x0); }
    for (long long i = 0; i < len; ++i) w[i] = cascade(f, ext_at(xr, n, f.edge, i), z);

    const double y0 = w[len - 1];
    for (int s = 0; s < f.n_sections; ++s) { z[s][0] = __dmul_rn(f.zi[s][0], y0); z[s][1] = __dmul_rn(f.zi[s][1], y0); }
    for (long long i = len - 1; i >= 0; --i) {
        const double v = cascade(f, w[i], z);
        const long long k = i - f.edge;
        if (k >= 0 && k < n) yr[2 * k] = v;
    }
}

int make_sos(Sos& f, const double* sos, int S, std::string& err) {
    if (S < 1 || S > MAX_SECTIONS) { err = "n_sections must be in [1, 8]"; return SSFM_ERR_INVALID; }
    std::memset(&f, 0, sizeof(f));
    f.n_sections = S;
    int nb = 0, na = 0;
    double scale = 1.0;
    for (int s = 0; s < S; ++s) {
        const double* c = sos + 6 * s;
        if (c[3] != 1.0) { err = "sos rows must be normalised (a0 == 1)"; return SSFM_ERR_INVALID; }
        for (int k = 0; k < 6; ++k) f.c[s][k] = c[k];
        nb += (c[2] == 0.0); na += (c[5] == 0.0);
        // steady state of the unit-step response (scipy.signal.sosfilt_zi / lfilter_zi), closed form
        const double B0 = c[1] - c[4] * c[0], B1 = c[2] - c[5] * c[0];
        const double z0 = (B0 + B1) / (1.0 + c[4] + c[5]);
        f.zi[s][0] = scale * z0;
        f.zi[s][1] = scale * (B1 - c[5] * z0);
        scale *= (c[0] + c[1] + c[2]) / (c[3] + c[4] + c[5]);
    }
    f.edge = 3 * (2 * S + 1 - (nb < na ? nb : na));
    return SSFM_OK;
}

}  // namespace ssfm_filt

extern thread_local std::string ssfm_err_slot;

extern "C" int ssfm_filtfilt_sos(void* x_dev, void* y_dev, int64_t n_rows, int64_t n, const double* sos_host,
                                 int32_t n_sections, int32_t device, void* stream) {
    using namespace ssfm_filt;
    std::string err;
    if (!x_dev || !y_dev || !sos_host) { ssfm_err_slot = "null buffer or sos"; return SSFM_ERR_INVALID; }
    if (n_rows < 1 || n < 1) { ssfm_err_slot = "n_rows and n_samples must be >= 1"; return SSFM_ERR_INVALID; }
    Sos f;
    int rc = make_sos(f, sos_host, n_sections, err);
    if (rc) { ssfm_err_slot = err; return rc; }
    if (n <= f.edge) {
        ssfm_err_slot = "The length of the input vector x must be greater than padlen, which is " + std::to_string(f.edge) + ".";
        return SSFM_ERR_INVALID;
    }
    cudaError_t e = cudaSetDevice(device);
    cudaStream_t st = (cudaStream_t)stream;
    double* ws = nullptr;
    const size_t len = (size_t)n + 2 * (size_t)f.edge;
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&ws, sizeof(double) * len * (size_t)n_rows * 2, st);
    if (e == cudaSuccess) {
        const long long threads = n_rows * 2;
        k_filtfilt_seq<<<(unsigned)((threads + 63) / 64), 64, 0, st>>>((const double*)x_dev, (double*)y_dev, ws, n_rows, n, f);
        e = cudaGetLastError();
    }
    if (ws) cudaFreeAsync(ws, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { ssfm_err_slot = std::string("filtfilt: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }
    return SSFM_OK;
}
