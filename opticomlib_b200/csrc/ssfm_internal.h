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

// y[rows][n] (complex128, device) <- circular convolution of each row with the zero-phase response
// |H(e^{jw})|^2 of the cascade `f`, computed as IFFT(|H|^2 FFT(y)) with the SSFM transform kernels.
// n must be a power of two in [2^8, 2^22].  Returns an SSFM_* code.
int ssfm_internal_zero_phase_circular(int device, long long n, long long rows, const ssfm_filt::Sos& f,
                                      void* y_dev, cudaStream_t st);
