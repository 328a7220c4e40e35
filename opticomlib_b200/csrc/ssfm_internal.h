// Internal (non-ABI) interfaces shared by the translation units of the extension.
#pragma once
#include <cuda_runtime.h>
#include <string>

namespace ssfm_filt {
constexpr int MAX_SECTIONS = 8;
struct Sos {
    double c[MAX_SECTIONS][6];   // b0 b1 b2 a0 a1 a2 per section
    double zi[MAX_SECTIONS][2];  // steady state of the unit-step response, scaled by the DC gain of earlier sections
    int n_sections;
    int edge;                    // odd-extension length used by scipy.signal.sosfiltfilt
};
}  // namespace ssfm_filt

extern thread_local std::string ssfm_err_slot;

// Circular convolution of rows of n complex128 samples with the zero-phase response |H(e^{jw})|^2 of the cascade `f`,
// computed as IFFT(|H|^2 FFT(y)) with the SSFM transform kernels (n a power of two in [2^8, 2^22]).
// prepare: tables and the response for rows of n samples, for at most `max_rows` rows per apply (cached per device, n and
// max_rows).  apply: `rows` <= max_rows rows at y_dev, read from src_dev when given (out of place), no synchronisation: one
// launch of the persistent kernel (rows in flight L2-resident: 1 read + 1 write of HBM per row) for n in [2^12, 2^20], three
// streaming kernels otherwise.  SSFM_* codes.
int ssfm_internal_transfer_prepare(int device, long long n, long long max_rows, const ssfm_filt::Sos& f, void** plan_out,
                                   cudaStream_t st);
int ssfm_internal_transfer_apply(void* plan, void* y_dev, long long rows, cudaStream_t st, const void* src_dev = nullptr,
                                 int one_launch = 1);

// Pass tables (+ the 256-entry sincos table behind them) of the M-point complex128 transform of fft_core.cuh, device memory owned
// by the caller.
int ssfm_internal_pass_tables_f64(void** out, int M, cudaStream_t st);

