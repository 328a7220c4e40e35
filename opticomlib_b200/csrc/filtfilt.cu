// Zero-phase cascaded-biquad filtering on the device (LPF / BPF hot path).
//
// Replaces scipy.signal.sosfiltfilt as called by the reference at opticomlib/devices.py:820-823
// (BPF) and 1365-1368 (LPF): odd extension by `edge` samples, steady-state initial conditions
// scaled by the first sample, forward cascade, backward cascade, strip the extension.
//
// Two device paths, both exact to rounding (no CPU path):
//   * FFT path (rows of 2^8..2^22 samples): the forward-backward cascade is a linear filter with
//     the real, zero-phase response |H(e^{jw})|^2.  Away from the ends its output equals the circular
//     convolution  IFFT(|H|^2 FFT(x))  (computed with the SSFM transform kernels, 3 kernels, all rows
//     in parallel); the deviation comes from the start-up transients of the two recursions and decays
//     like rho^n (rho = largest pole radius).  The first and last K samples (rho^K < 1e-19) are then
//     recomputed exactly with the recursion itself on a short segment (k_filtfilt_edges).
//   * sequential path (any other length, or K too large for the row): one thread per
//     (row, real|imag) walks the whole recursion, same operation order as the SciPy loop.
#include <cmath>
#include <cstring>
#include <string>

#include "../../include/ssfm_b200.h"
#include "ssfm_internal.h"

namespace ssfm_filt {

__device__ __forceinline__ double cascade(const Sos& f, double v, double (&z)[MAX_SECTIONS][2]) {
#pragma unroll 1
    for (int s = 0; s < f.n_sections; ++s) {   // direct form II transposed, products and sums rounded separately
        const double y = __dadd_rn(__dmul_rn(f.c[s][0], v), z[s][0]);
        z[s][0] = __dadd_rn(__dsub_rn(__dmul_rn(f.c[s][1], v), __dmul_rn(f.c[s][4], y)), z[s][1]);
        z[s][1] = __dsub_rn(__dmul_rn(f.c[s][2], v), __dmul_rn(f.c[s][5], y));
        v = y;
    }
    return v;
}
__device__ __forceinline__ void init_state(const Sos& f, double x0, double (&z)[MAX_SECTIONS][2]) {
    for (int s = 0; s < f.n_sections; ++s) { z[s][0] = __dmul_rn(f.zi[s][0], x0); z[s][1] = __dmul_rn(f.zi[s][1], x0); }
}

// odd extension of one component of an interleaved complex row: 2*x[0]-x[edge-i] | x | 2*x[n-1]-x[n-2-m]
__device__ __forceinline__ double ext_at(const double* x, long long n, int edge, long long i) {
    if (i < edge) return 2.0 * x[0] - x[2 * (edge - i)];
    i -= edge;
    if (i < n) return x[2 * i];
    i -= n;
    return 2.0 * x[2 * (n - 1)] - x[2 * (n - 2 - i)];
}

// ---- sequential path: x, y [rows][n] complex128; ws [rows*2][n + 2*edge] doubles ---------------------
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
    init_state(f, ext_at(xr, n, f.edge, 0), z);
    for (long long i = 0; i < len; ++i) w[i] = cascade(f, ext_at(xr, n, f.edge, i), z);
    init_state(f, w[len - 1], z);
    for (long long i = len - 1; i >= 0; --i) {
        const double v = cascade(f, w[i], z);
        const long long k = i - f.edge;
        if (k >= 0 && k < n) yr[2 * k] = v;
    }
}

// ---- FFT path, step 1: save the first and last L samples of every row (the circular pass may run in place)
// seg [rows][2][L] complex128
__global__ void k_save_edges(const double2* __restrict__ x, double2* __restrict__ seg, long long rows, long long n, int L) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 2 * L) return;
    const long long row = i / (2 * L);
    const int side = (int)((i / L) & 1), j = (int)(i % L);
    seg[i] = x[row * n + (side ? n - L + j : j)];
}

// ---- FFT path, step 3: exact recursion on the two end segments, overwrite the first / last K outputs.
// One thread per (row, side, component).  Head: exact forward pass from the true start (odd extension,
// zi*ext[0]); the backward pass starts at the cut with the steady-state guess, whose error has decayed
// by rho^(L-K) when it reaches sample K.  Tail: mirror image (approximate forward start at the cut, exact
// odd extension and backward pass from the true end).
__global__ void k_filtfilt_edges(const double* __restrict__ seg, double* __restrict__ y, double* __restrict__ ws,
                                 long long rows, long long n, int L, int K, Sos f) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= rows * 4) return;
    const long long row = tid >> 2;
    const int side = (int)((tid >> 1) & 1), comp = (int)(tid & 1);
    const double* s = seg + ((row * 2 + side) * (long long)L) * 2 + comp;   // s[2*i]: sample i of the segment
    const int E = f.edge, len = L + E;
    double* w = ws + tid * (long long)len;
    double* yr = y + row * n * 2 + comp;
    double z[MAX_SECTIONS][2];
    if (side == 0) {
        const double x0 = s[0];
        init_state(f, 2.0 * x0 - s[2 * E], z);                                  // ext[0] = 2 x[0] - x[edge]
        for (int i = 0; i < E; ++i) w[i] = cascade(f, 2.0 * x0 - s[2 * (E - i)], z);
        for (int i = 0; i < L; ++i) w[E + i] = cascade(f, s[2 * i], z);
        init_state(f, w[len - 1], z);
        for (int i = len - 1; i >= E; --i) {
            const double v = cascade(f, w[i], z);
            if (i - E < K) yr[2 * (long long)(i - E)] = v;
        }
    } else {
        const double xl = s[2 * (L - 1)];
        init_state(f, s[0], z);
        for (int i = 0; i < L; ++i) w[i] = cascade(f, s[2 * i], z);
        for (int m = 0; m < E; ++m) w[L + m] = cascade(f, 2.0 * xl - s[2 * (L - 2 - m)], z);
        init_state(f, w[len - 1], z);
        for (int i = len - 1; i >= L - K; --i) {
            const double v = cascade(f, w[i], z);
            if (i < L) yr[2 * (n - L + i)] = v;
        }
    }
}

int make_sos(Sos& f, const double* sos, int S, double* rho_out, std::string& err) {
    if (S < 1 || S > MAX_SECTIONS) { err = "n_sections must be in [1, 8]"; return SSFM_ERR_INVALID; }
    std::memset(&f, 0, sizeof(f));
    f.n_sections = S;
    int nb = 0, na = 0;
    double scale = 1.0, rho = 0.0;
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
        // pole radius of z^2 + a1 z + a2
        const double disc = c[4] * c[4] - 4.0 * c[5];
        const double r = disc < 0 ? std::sqrt(c[5]) : 0.5 * (std::fabs(c[4]) + std::sqrt(disc));
        if (r > rho) rho = r;
    }
    f.edge = 3 * (2 * S + 1 - (nb < na ? nb : na));
    *rho_out = rho;
    return SSFM_OK;
}

}  // namespace ssfm_filt

extern "C" int ssfm_filtfilt_sos(void* x_dev, void* y_dev, int64_t n_rows, int64_t n, const double* sos_host,
                                 int32_t n_sections, int32_t device, void* stream) {
    using namespace ssfm_filt;
    std::string err;
    if (!x_dev || !y_dev || !sos_host) { ssfm_err_slot = "null buffer or sos"; return SSFM_ERR_INVALID; }
    if (n_rows < 1 || n < 1) { ssfm_err_slot = "n_rows and n_samples must be >= 1"; return SSFM_ERR_INVALID; }
    Sos f;
    double rho = 0;
    int rc = make_sos(f, sos_host, n_sections, &rho, err);
    if (rc) { ssfm_err_slot = err; return rc; }
    if (n <= f.edge) {
        ssfm_err_slot = "The length of the input vector x must be greater than padlen, which is " + std::to_string(f.edge) + ".";
        return SSFM_ERR_INVALID;
    }
    cudaError_t e = cudaSetDevice(device);
    cudaStream_t st = (cudaStream_t)stream;
    if (e != cudaSuccess) { ssfm_err_slot = std::string("filtfilt: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }

    // transient length: rho^K < 1e-19, with head-room for the polynomial factors of clustered Bessel poles
    long long K = -1;
    if (rho > 0 && rho < 1) K = (long long)std::ceil(1.3 * std::log(1e-19) / std::log(rho)) + 64;
    else if (rho == 0) K = 64;
    const bool pow2 = (n & (n - 1)) == 0 && n >= 256 && n <= (1ll << 22);
    const bool fft_path = pow2 && K > 0 && 4 * K + 2 * f.edge <= n && !getenv("SSFM_FILTFILT_SEQUENTIAL");

    double* ws = nullptr;
    double2* seg = nullptr;
    if (fft_path) {
        const int L = (int)(2 * K), Ki = (int)K;
        e = cudaMallocAsync((void**)&seg, sizeof(double2) * (size_t)n_rows * 2 * L, st);
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&ws, sizeof(double) * (size_t)n_rows * 4 * (L + f.edge), st);
        if (e == cudaSuccess) {
            const long long cnt = n_rows * 2 * L;
            k_save_edges<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>((const double2*)x_dev, seg, n_rows, n, L);
            if (y_dev != x_dev) e = cudaMemcpyAsync(y_dev, x_dev, sizeof(double2) * (size_t)n_rows * n, cudaMemcpyDeviceToDevice, st);
        }
        if (e == cudaSuccess) {
            rc = ssfm_internal_zero_phase_circular(device, n, n_rows, f, y_dev, st);
            if (!rc) {
                const long long threads = n_rows * 4;
                k_filtfilt_edges<<<(unsigned)((threads + 63) / 64), 64, 0, st>>>((const double*)seg, (double*)y_dev, ws,
                                                                              n_rows, n, L, Ki, f);
                e = cudaGetLastError();
            }
        }
    } else {
        const size_t len = (size_t)n + 2 * (size_t)f.edge;
        e = cudaMallocAsync((void**)&ws, sizeof(double) * len * (size_t)n_rows * 2, st);
        if (e == cudaSuccess) {
            const long long threads = n_rows * 2;
            k_filtfilt_seq<<<(unsigned)((threads + 63) / 64), 64, 0, st>>>((const double*)x_dev, (double*)y_dev, ws, n_rows, n, f);
            e = cudaGetLastError();
        }
    }
    if (ws) cudaFreeAsync(ws, st);
    if (seg) cudaFreeAsync(seg, st);
    if (rc) return rc;
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { ssfm_err_slot = std::string("filtfilt: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }
    return SSFM_OK;
}
