// Zero-phase cascaded-biquad filtering on the device (LPF / BPF hot path).
//
// Replaces scipy.signal.sosfiltfilt as called by the reference at opticomlib/devices.py:820-823
// (BPF) and 1365-1368 (LPF): odd extension by `edge` samples, steady-state initial conditions
// scaled by the first sample, forward cascade, backward cascade, strip the extension.
//
// Two device paths, both exact to rounding (no CPU path):
//   * FFT path (rows of 2^8..2^22 samples): the forward-backward cascade is a linear filter with
//     the real, zero-phase response |H(e^{jw})|^2.  Away from the ends its output equals the circular
//     convolution  IFFT(|H|^2 FFT(x))  (computed with the SSFM transform kernels); the deviation comes from
//     the start-up transients of the two recursions and decays like rho^n (rho = largest pole radius).
//     The first and last K samples (rho^K < 1e-19) are then recomputed exactly with the recursion itself on
//     a short segment (k_filtfilt_edges).
//     Schedule (round 2): the rows are cut into chunks of <= 32 MiB and the three transform kernels (plus
//     the optional pack / unpack stages of the photodetector path) run back to back on one chunk, so the
//     chunk stays in the 126 MB L2 between them: HBM sees ONE read of the input and ONE write of the
//     output (round 1 streamed every row through HBM four times).  The end segments of ALL rows are saved
//     first and their recursions run on a side stream while the chunks go through (one launch, one wave of
//     threads, latency-bound: ~0.4 ms that used to sit on the critical path); a last small kernel writes
//     the exact end samples over the circular ones.
//   * overlap-save path (any row length >= 4096 when K <= 1024 -- the receiver filters of the reference's examples have
//     K = 360 .. 850): the zero-phase response is K samples long on either side, so nothing forces a transform of the whole
//     row.  Every CTA filters ONE block of 4096 samples on its own -- 4096-point transform, x |H|^2, inverse transform, all in
//     registers and shared memory (fft_core.cuh) -- and keeps the 4096 - 2K samples in the middle, whose circular wrap-around
//     is below rho^K < 1e-19.  No team of CTAs, no barrier between CTAs, no second pass: HBM sees one read of the input (the
//     halos of neighbouring blocks hit L2) and one write of the output, and the photodetector square law and the sampler
//     run in the load and the store of the same kernel (k_ols).  The ends of a row are handled as above.
//   * sequential path (any other length, or K too large for the row): one thread per
//     (row, real|imag) walks the whole recursion, same operation order as the SciPy loop.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/ssfm_b200.h"
#include "fft_core.cuh"
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


// ---------------------------------------------------------------------------------------------------------------
// Photodetector front end fused in front of the filter (reference PD, devices.py:1514-1552, followed by LPF 1363-1368)
// and sampler behind it (SAMPLER, devices.py:1871-1891: output[instant :: sps]).
//   signal current  i_sig  = r sum_pol |E|^2                                  (devices.py:1514-1517)
//   noise current   i_n    = r sum_pol (2 Re(E n*) + |n|^2) + extra + i_dark  (typing.py:1341-1342 product rule; 1530-1545)
// packed as (re, im) = R_load (i_sig, i_n) of ONE complex row, so that one transform filters both separately, as the
// reference does (devices.py:1365-1368).
struct PdSrc {
    const double2* field;     // [rows][n_pol][n]
    const double2* noise;     // same shape or null
    const double* extra;      // [rows][n] additive noise current (thermal + shot) or null
    double r, r_load, i_dark;
    int n_pol;
    int noise_out;            // 1: a noise row is produced (im part), 0: im = 0
};
__device__ __forceinline__ double2 pd_sample(const PdSrc& s, long long row, long long n, long long i) {
    double sig = 0.0, noi = 0.0;
    for (int p = 0; p < s.n_pol; ++p) {
        const double2 e = s.field[(row * s.n_pol + p) * n + i];
        sig += e.x * e.x + e.y * e.y;
        if (s.noise) {
            const double2 z = s.noise[(row * s.n_pol + p) * n + i];
            noi += 2.0 * (e.x * z.x + e.y * z.y) + (z.x * z.x + z.y * z.y);
        }
    }
    double2 o;
    o.x = s.r_load * (s.r * sig);
    o.y = s.noise_out ? s.r_load * (s.r * noi + (s.extra ? s.extra[row * n + i] : 0.0) + s.i_dark) : 0.0;
    return o;
}
__global__ void k_pd_pack(PdSrc s, double2* __restrict__ dst, long long row0, long long rows, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * n) return;
    dst[i] = pd_sample(s, row0 + i / n, n, i % n);
}
// end segments of the PACKED rows straight from the optical field (so that they can be saved before any chunk is packed)
__global__ void k_pd_save_edges(PdSrc s, double2* __restrict__ seg, long long rows, long long n, int L) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 2 * L) return;
    const long long row = i / (2 * L);
    const int side = (int)((i / L) & 1), j = (int)(i % L);
    seg[i] = pd_sample(s, row, n, side ? n - L + j : j);
}
// filtered packed rows -> real outputs, decimated: out[row][j] = re/im of y[row][offset + j*stride]
__global__ void k_unpack_real(const double2* __restrict__ y, double* __restrict__ out_sig, double* __restrict__ out_noise,
                              long long row0, long long rows, long long n, long long offset, long long stride, long long m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * m) return;
    const long long row = i / m, j = i % m;
    const double2 v = y[row * n + offset + j * stride];
    out_sig[(row0 + row) * m + j] = v.x;
    if (out_noise) out_noise[(row0 + row) * m + j] = v.y;
}

// ---- FFT path: exact recursion on the two end segments of every row, results into edge_out[rows][2][K] (complex).
// One thread per (row, side, component).  Head: exact forward pass from the true start (odd extension,
// zi*ext[0]); the backward pass starts at the cut with the steady-state guess, whose error has decayed
// by rho^(L-K) when it reaches sample K.  Tail: mirror image (approximate forward start at the cut, exact
// odd extension and backward pass from the true end).
// The recursions are one dependent chain per thread, but their INPUTS are not: samples are fetched in independent batches of
// EDGE_BATCH loads (all in flight together) ahead of the chain, so a step costs its arithmetic latency, not an L2 round trip
// (the first version loaded w[i] inside the chain: ~780 cycles per step, 2.1 ms per call whatever the batch size).
constexpr int EDGE_BATCH = 8;
__global__ void k_filtfilt_edges_out(const double* __restrict__ seg, double* __restrict__ edge_out, double* __restrict__ ws,
                                     long long rows, int L, int K, Sos f) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= rows * 4) return;
    const long long row = tid >> 2;
    const int side = (int)((tid >> 1) & 1), comp = (int)(tid & 1);
    const double* s = seg + ((row * 2 + side) * (long long)L) * 2 + comp;   // s[2*i]: sample i of the segment
    const int E = f.edge, len = L + E;
    double* w = ws + tid * (long long)len;
    double* o = edge_out + ((row * 2 + side) * (long long)K) * 2 + comp;    // o[2*i]: output sample i of this end
    double z[MAX_SECTIONS][2];
    double buf[EDGE_BATCH];
    // forward pass over the odd-extended segment: ext[i] = 2 x0 - s[E - i] (head, i < E) | s[i - E] ... | 2 xl - s[L - 2 - m] (tail)
    const double x0 = s[0], xl = s[2 * (L - 1)];
    auto ext = [&](int i) -> double {
        if (side == 0) return i < E ? 2.0 * x0 - s[2 * (E - i)] : s[2 * (i - E)];
        return i < L ? s[2 * i] : 2.0 * xl - s[2 * (L - 2 - (i - L))];
    };
    init_state(f, ext(0), z);                                               // head: exact start; tail: steady-state guess at the cut
    for (int i0 = 0; i0 < len; i0 += EDGE_BATCH) {
        const int nb = min(EDGE_BATCH, len - i0);
#pragma unroll
        for (int k = 0; k < EDGE_BATCH; ++k) buf[k] = k < nb ? ext(i0 + k) : 0.0;
#pragma unroll
        for (int k = 0; k < EDGE_BATCH; ++k) if (k < nb) w[i0 + k] = cascade(f, buf[k], z);
    }
    // backward pass: head keeps output samples 0 .. K-1 (w index E + j), tail keeps n-K .. n-1 (w index L-K + j)
    const int lo = side == 0 ? E : L - K;                                   // lowest w index whose output is wanted
    const int keep0 = side == 0 ? E : L - K, keep1 = side == 0 ? E + K : L; // outputs for w indices in [keep0, keep1)
    init_state(f, w[len - 1], z);
    for (int i1 = len - 1; i1 >= lo; i1 -= EDGE_BATCH) {
        const int nb = min(EDGE_BATCH, i1 - lo + 1);
#pragma unroll
        for (int k = 0; k < EDGE_BATCH; ++k) buf[k] = k < nb ? w[i1 - k] : 0.0;
#pragma unroll
        for (int k = 0; k < EDGE_BATCH; ++k) {
            if (k < nb) {
                const double v = cascade(f, buf[k], z);
                const int i = i1 - k;
                if (i >= keep0 && i < keep1) o[2 * (i - keep0)] = v;
            }
        }
    }
}
// exact end samples over the circular ones: complex output rows, or the decimated real outputs of the photodetector path
__global__ void k_scatter_edges(const double2* __restrict__ edge_out, double2* __restrict__ y, double* __restrict__ out_sig,
                                double* __restrict__ out_noise, long long rows, long long n, int K, long long offset,
                                long long stride, long long m) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * 2 * K) return;
    const long long row = i / (2 * K);
    const int side = (int)((i / K) & 1), j = (int)(i % K);
    const long long pos = side ? n - K + j : j;
    const double2 v = edge_out[i];
    if (y) { y[row * n + pos] = v; return; }
    if (pos < offset || (pos - offset) % stride) return;
    const long long jj = (pos - offset) / stride;
    if (jj >= m) return;
    out_sig[row * m + jj] = v.x;
    if (out_noise) out_noise[row * m + jj] = v.y;
}

// ---- overlap-save path: one CTA = one block of OLS_M samples of one row ---------------------------------------------------
// Block b of a row filters samples [b L - K, b L - K + OLS_M) (indices wrap around the row: only the first and last K outputs of
// a row see the wrap, and those are replaced by the exact end segments) and stores outputs [b L, (b + 1) L), L = OLS_M - 2 K.
constexpr int OLS_M = 4096, OLS_E = 16, OLS_NT = OLS_M / OLS_E;
struct OlsArgs {
    const double2* x;        // complex rows [rows][n] (pd == 0)
    PdSrc pd;                // photodetector front end (PD == true)
    double2* y;              // complex output rows [rows][n], or null:
    double* out_sig;         //   decimated real outputs [row0 + rows][m]: samples offset, offset + stride, ...
    double* out_noise;
    const double2* tw;       // pass tables of the OLS_M-point transform (global memory, L1-resident)
    const double* h2;        // |H(e^{j 2 pi k / OLS_M})|^2 / OLS_M
    long long n, row0, offset, stride, m;
    int K, L, nb;            // halo, outputs per block, blocks per row
    int wave;                // CTAs resident at once (L2 prefetch distance)
};
__global__ void k_fill_h2(double* out, Sos f) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= OLS_M) return;
    double s, c;
    sincospi(2.0 * (double)k / (double)OLS_M, &s, &c);          // z^-1 = (c, -s), z^-2 = (c2, -s2)
    const double c2 = c * c - s * s, s2 = 2.0 * s * c;
    double g = 1.0 / (double)OLS_M;
    for (int i = 0; i < f.n_sections; ++i) {
        const double* q = f.c[i];
        const double nr = q[0] + q[1] * c + q[2] * c2, ni = -(q[1] * s + q[2] * s2);
        const double dr = q[3] + q[4] * c + q[5] * c2, di = -(q[4] * s + q[5] * s2);
        g *= (nr * nr + ni * ni) / (dr * dr + di * di);
    }
    out[k] = g;
}
#ifndef OLS_TW
#define OLS_TW ssfm::ChainTwiddles
#endif
template <bool PD>
__global__ void __launch_bounds__(OLS_NT, 2) k_ols(OlsArgs a) {
    typedef ssfm::RowExchange<OLS_M, OLS_E> X;
    extern __shared__ __align__(16) unsigned char ols_smem[];
    double2* sm = reinterpret_cast<double2*>(ols_smem);
    const int tid = threadIdx.x;
    const long long row = blockIdx.x / a.nb;
    const int b = (int)(blockIdx.x % a.nb);
    const long long s0 = (long long)b * a.L - a.K;
    double2 v[OLS_E];
    if (!PD) {
#pragma unroll
        for (int q = 0; q < OLS_E; ++q) {
            long long i = s0 + tid + q * OLS_NT;
            i = i < 0 ? i + a.n : (i >= a.n ? i - a.n : i);
            v[q] = a.x[row * a.n + i];
        }
        // the block that this CTA slot runs next (one grid wave ahead) goes to L2 while this one computes
        const long long ahead = (long long)blockIdx.x + a.wave;
        if ((tid & 7) == 0 && ahead < (long long)gridDim.x) {
            const long long r2 = ahead / a.nb, st2 = (ahead % a.nb) * (long long)a.L - a.K;
#pragma unroll
            for (int q = 0; q < OLS_E; ++q) {
                const long long i = st2 + tid + q * OLS_NT;
                if (i >= 0 && i < a.n) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.x + r2 * a.n + i));
            }
        }
    } else {
        // square law and beat terms of pd_sample(), four samples at a time with their loads issued together (one sample at
        // a time the block waits for 16 dependent trips to memory)
        const PdSrc& s = a.pd;
        const long long r = a.row0 + row;
        if (!s.noise) {                                                 // no optical noise: the 16 loads of a polarisation together
            long long idx[OLS_E];
            double sig[OLS_E];
#pragma unroll
            for (int q = 0; q < OLS_E; ++q) {
                const long long i = s0 + tid + q * OLS_NT;
                idx[q] = i < 0 ? i + a.n : (i >= a.n ? i - a.n : i);
                sig[q] = 0.0;
            }
            for (int p = 0; p < s.n_pol; ++p) {
                const double2* fp = s.field + (r * s.n_pol + p) * a.n;
                double2 e[OLS_E];
#pragma unroll
                for (int q = 0; q < OLS_E; ++q) e[q] = fp[idx[q]];
#pragma unroll
                for (int q = 0; q < OLS_E; ++q) sig[q] += e[q].x * e[q].x + e[q].y * e[q].y;
            }
#pragma unroll
            for (int q = 0; q < OLS_E; ++q) {
                v[q].x = s.r_load * (s.r * sig[q]);
                v[q].y = s.noise_out ? s.r_load * (s.r * 0.0 + (s.extra ? s.extra[r * a.n + idx[q]] : 0.0) + s.i_dark) : 0.0;
            }
        } else
#pragma unroll
        for (int q0 = 0; q0 < OLS_E; q0 += 4) {
            long long idx[4];
            double sig[4] = {0.0, 0.0, 0.0, 0.0}, noi[4] = {0.0, 0.0, 0.0, 0.0}, ex[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                long long i = s0 + tid + (q0 + k) * OLS_NT;
                idx[k] = i < 0 ? i + a.n : (i >= a.n ? i - a.n : i);
            }
            if (s.extra && s.noise_out) {
#pragma unroll
                for (int k = 0; k < 4; ++k) ex[k] = s.extra[r * a.n + idx[k]];
            }
            for (int p = 0; p < s.n_pol; ++p) {
                const double2* fp = s.field + (r * s.n_pol + p) * a.n;
                double2 e[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) e[k] = fp[idx[k]];
                if (s.noise) {
                    const double2* zp = s.noise + (r * s.n_pol + p) * a.n;
                    double2 z[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) z[k] = zp[idx[k]];
#pragma unroll
                    for (int k = 0; k < 4; ++k) noi[k] += 2.0 * (e[k].x * z[k].x + e[k].y * z[k].y) + (z[k].x * z[k].x + z[k].y * z[k].y);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) sig[k] += e[k].x * e[k].x + e[k].y * e[k].y;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[q0 + k].x = s.r_load * (s.r * sig[k]);
                v[q0 + k].y = s.noise_out ? s.r_load * (s.r * noi[k] + ex[k] + s.i_dark) : 0.0;
            }
        }
    }
    ssfm::fft_passes<double, OLS_M, -1, X, OLS_E, 1, OLS_TW>::run(v, sm, a.tw, tid);
#pragma unroll
    for (int q = 0; q < OLS_E; ++q) {
        const double g = __ldg(a.h2 + tid + q * OLS_NT);
        v[q].x *= g; v[q].y *= g;
    }
    ssfm::fft_passes<double, OLS_M, +1, X, OLS_E, 1, OLS_TW>::run(v, sm, a.tw, tid);
#pragma unroll
    for (int q = 0; q < OLS_E; ++q) {
        const int j = tid + q * OLS_NT;
        const long long pos = s0 + j;
        if (j < a.K || j >= a.K + a.L || pos >= a.n) continue;
        if (!PD || a.y) { a.y[row * a.n + pos] = v[q]; continue; }
        if (pos < a.offset || (pos - a.offset) % a.stride) continue;
        const long long jj = (pos - a.offset) / a.stride;
        if (jj >= a.m) continue;
        a.out_sig[(a.row0 + row) * a.m + jj] = v[q].x;
        if (a.out_noise) a.out_noise[(a.row0 + row) * a.m + jj] = v[q].y;
    }
}

// ---- photodetector -> low-pass -> SAMPLER (reference PD devices.py:1514-1552, its closing LPF 1363-1368, SAMPLER 1871-1891:
// output[instant :: sps]) when the sampler keeps one sample in `stride` (stride >= 32): the outputs wanted are
// so few that evaluating the zero-phase response as a FIR filter AT those samples only -- 2K + 1 taps each, the taps being the
// response of k_ols to a unit impulse -- costs (2K + 1)/stride multiply-adds per input sample and component (26 for K = 848,
// stride 64) against ~130 FP64 instructions per input sample for the block transforms, and needs no exchange at all.
// One CTA = 64 consecutive outputs of one row.  The span of input samples they need (stride * 63 + 2K + 1) is detected (square
// law, beat terms) straight into shared memory as a matrix X[kk][a] = sample (stride * a + kk) of the span, so that
//   y[j] = sum_kk sum_kb h[kb * stride + kk] X[kk][j + kb]
// is, for every kk, a short FIR (ceil((2K + 1)/stride) taps) along a row of X.  A thread owns 4 consecutive outputs and a few
// values of kk: per tap it loads ONE new sample and ONE tap for 8 multiply-adds (a sliding window in registers), so the kernel
// is bound by the FP64 pipe, not by shared memory (one load per multiply-add in the naive form: measured 5.4 ms for
// 1024 x 2^18 samples, no faster than the block transforms).  Rows of X are stored de-interleaved by 4 (column a at
// (a % 4) * AQ + a / 4), which makes the window loads of the 16 threads of a group consecutive 16-byte words.
constexpr int FIR_NT = 256, FIR_TO = 64, FIR_R = 4, FIR_TPG = FIR_TO / FIR_R, FIR_G = FIR_NT / FIR_TPG;   // 16 threads x 16 groups
struct FirArgs {
    PdSrc pd;
    const double2* taps;     // taps[k].x = zero-phase impulse response at lag k - K, k = 0 .. 2K (output of k_ols on an impulse at K)
    double* out_sig;
    double* out_noise;
    long long n, m, offset, stride, row0;
    int K, AQ, kb_total;     // halo; quarter pitch of a row of X (even); ceil((2K + 1) / stride) rounded up to a multiple of 4
};
__device__ __forceinline__ int fir_col(int a, int AQ) { return (a & 3) * AQ + (a >> 2); }
template <bool SIMPLE>       // SIMPLE: one polarisation, no optical noise field (the common receiver): 8 samples per thread in flight
__global__ void __launch_bounds__(FIR_NT, 2) k_pd_fir(FirArgs a) {
    extern __shared__ __align__(16) unsigned char fir_smem[];
    const int S = (int)a.stride, ntap = 2 * a.K + 1, A = 4 * a.AQ + 1;  // (odd pitch: the stores of the detection phase spread over the banks)
    double2* xs = reinterpret_cast<double2*>(fir_smem);                 // [stride][4 * AQ + 1] detected samples (signal, noise)
    double* hs = reinterpret_cast<double*>(xs + (size_t)S * A);         // [kb_total * stride] taps, zero-padded
    const int tid = threadIdx.x;
    const long long per_row = (a.m + FIR_TO - 1) / FIR_TO;
    const long long row = blockIdx.x / per_row;
    const long long j0 = (blockIdx.x % per_row) * FIR_TO;               // first output of this CTA
    const long long g0 = a.offset + (long long)S * j0 - a.K;            // first input sample of the span
    const int span = S * (FIR_TO - 1) + ntap;
    const int fill = S * (FIR_TO + a.kb_total + 2);                     // every column the windows touch is written (zeros beyond the span:
    for (int i = tid; i < a.kb_total * S; i += FIR_NT) {               // they meet zero taps, but 0 x garbage could be NaN)
        const int kk = i / a.kb_total, kb = i % a.kb_total, k = kb * S + kk;   // hs[kk][kb] = tap kb * stride + kk
        hs[i] = k < ntap ? a.taps[k].x : 0.0;
    }
    const PdSrc& s = a.pd;
    const long long r = a.row0 + row;
    constexpr int NL = SIMPLE ? 8 : 2;
    // sample u of the span = row (u mod stride), column (u / stride) of X; the pair advances by FIR_NT without a division
    const int dkk = FIR_NT % S, dcol = FIR_NT / S;
    int kk_u = tid % S, col_u = tid / S;
    const double2* f0 = s.field + (r * s.n_pol) * a.n;
    const double2* f1 = f0 + a.n;
    const double2* z0 = s.noise ? s.noise + (r * s.n_pol) * a.n : nullptr;
    const double2* z1 = z0 ? z0 + a.n : nullptr;
    const double* xe = (s.extra && s.noise_out) ? s.extra + r * a.n : nullptr;
    const bool pol2 = s.n_pol > 1;
#pragma unroll 1
    for (int u0 = tid; u0 < fill; u0 += NL * FIR_NT) {
        double2 e[NL][SIMPLE ? 1 : 2], z[NL][SIMPLE ? 1 : 2];
        double ex[NL];
        bool in[NL];
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            const int u = u0 + i * FIR_NT;
            const long long gi = g0 + u;
            in[i] = u < span && gi >= 0 && gi < a.n;
            ex[i] = 0.0;
            e[i][0] = in[i] ? f0[gi] : make_double2(0.0, 0.0);
            if (!SIMPLE) {
                e[i][1] = (in[i] && pol2) ? f1[gi] : make_double2(0.0, 0.0);
                z[i][0] = (in[i] && z0) ? z0[gi] : make_double2(0.0, 0.0);
                z[i][1] = (in[i] && z0 && pol2) ? z1[gi] : make_double2(0.0, 0.0);
            }
            if (in[i] && xe) ex[i] = xe[gi];
        }
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            const int u = u0 + i * FIR_NT;
            if (u < fill) {
                double sig = 0.0, noi = 0.0;
#pragma unroll
                for (int p = 0; p < (SIMPLE ? 1 : 2); ++p) {
                    sig += e[i][p].x * e[i][p].x + e[i][p].y * e[i][p].y;
                    if (!SIMPLE) noi += 2.0 * (e[i][p].x * z[i][p].x + e[i][p].y * z[i][p].y) + (z[i][p].x * z[i][p].x + z[i][p].y * z[i][p].y);
                }
                double2 v;                                              // (samples outside the row only reach outputs within K of its
                v.x = in[i] ? s.r_load * (s.r * sig) : 0.0;             //  ends, which the exact end segments replace)
                v.y = (in[i] && s.noise_out) ? s.r_load * (s.r * noi + ex[i] + s.i_dark) : 0.0;
                xs[(size_t)kk_u * A + fir_col(col_u, a.AQ)] = v;
            }
            kk_u += dkk; col_u += dcol;
            if (kk_u >= S) { kk_u -= S; ++col_u; }
        }
    }
    __syncthreads();
    const int t = tid % FIR_TPG, g = tid / FIR_TPG;                     // outputs 4 t .. 4 t + 3; rows kk = g, g + 16, ...
    double2 acc[FIR_R];
#pragma unroll
    for (int q = 0; q < FIR_R; ++q) acc[q] = make_double2(0.0, 0.0);
    const int kb4 = a.kb_total;                                         // (a multiple of 4: the host pads with zero taps)
#pragma unroll 1
    for (int kk = g; kk < S; kk += FIR_G) {
        // column 4 (t + q) + c of row kk sits at xr[c * AQ + t + q]: four pointers, one index
        const double2* p0 = xs + (size_t)kk * A + t;
        const double2* p1 = p0 + a.AQ;
        const double2* p2 = p1 + a.AQ;
        const double2* p3 = p2 + a.AQ;
        const double2* hq = reinterpret_cast<const double2*>(hs + (size_t)kk * kb4);   // taps of this row, consecutive in kb
        double2 w0 = p0[0], w1 = p1[0], w2 = p2[0];
#pragma unroll 2
        for (int q = 0; q < kb4 / 4; ++q) {
            const double2 h01 = hq[2 * q], h23 = hq[2 * q + 1];
            const double2 w3 = p3[q], w4 = p0[q + 1], w5 = p1[q + 1], w6 = p2[q + 1];
#define FIR_TAP(h, a0, a1, a2, a3)                                                    \
            acc[0].x = fma(h, a0.x, acc[0].x); acc[0].y = fma(h, a0.y, acc[0].y);     \
            acc[1].x = fma(h, a1.x, acc[1].x); acc[1].y = fma(h, a1.y, acc[1].y);     \
            acc[2].x = fma(h, a2.x, acc[2].x); acc[2].y = fma(h, a2.y, acc[2].y);     \
            acc[3].x = fma(h, a3.x, acc[3].x); acc[3].y = fma(h, a3.y, acc[3].y);
            FIR_TAP(h01.x, w0, w1, w2, w3)
            FIR_TAP(h01.y, w1, w2, w3, w4)
            FIR_TAP(h23.x, w2, w3, w4, w5)
            FIR_TAP(h23.y, w3, w4, w5, w6)
#undef FIR_TAP
            w0 = w4; w1 = w5; w2 = w6;
        }
    }
    __syncthreads();                                                    // the sample matrix is free: partial sums go there
    double2* red = xs;                                                  // [FIR_G][FIR_TO]
#pragma unroll
    for (int q = 0; q < FIR_R; ++q) red[g * FIR_TO + 4 * t + q] = acc[q];
    __syncthreads();
    if (tid < FIR_TO && j0 + tid < a.m) {
        double sx = 0.0, sy = 0.0;
#pragma unroll
        for (int q = 0; q < FIR_G; ++q) { sx += red[q * FIR_TO + tid].x; sy += red[q * FIR_TO + tid].y; }
        a.out_sig[r * a.m + j0 + tid] = sx;
        if (a.out_noise) a.out_noise[r * a.m + j0 + tid] = sy;
    }
}

// Philox4x32-10 counter-based generator + Box-Muller: N(0, 1) doubles, reproducible from (seed, element index) alone,
// whatever the launch geometry (the reference draws its noise from NumPy's global stream: devices.py:933, 1523, 1527 --
// a device-side generator can only be validated statistically).
__device__ __forceinline__ void philox4x32_10(unsigned int (&c)[4], unsigned int k0, unsigned int k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned long long p0 = (unsigned long long)0xD2511F53u * c[0], p1 = (unsigned long long)0xCD9E8D57u * c[2];
        const unsigned int n0 = (unsigned int)(p1 >> 32) ^ c[1] ^ k0, n1 = (unsigned int)p1;
        const unsigned int n2 = (unsigned int)(p0 >> 32) ^ c[3] ^ k1, n3 = (unsigned int)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// two independent N(0,1) values for counter `idx` of stream (seed, sub)
__device__ __forceinline__ double2 philox_normal2(unsigned long long seed, unsigned int sub, unsigned long long idx) {
    unsigned int c[4] = {(unsigned int)idx, (unsigned int)(idx >> 32), sub, 0u};
    philox4x32_10(c, (unsigned int)seed, (unsigned int)(seed >> 32));
    const unsigned long long a = ((unsigned long long)c[0] << 32) | c[1], b = ((unsigned long long)c[2] << 32) | c[3];
    const double u1 = ((double)(a >> 11) + 0.5) * (1.0 / 9007199254740992.0);     // (0, 1)
    const double u2 = ((double)(b >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double rad = sqrt(-2.0 * log(u1));
    double sn, cs;
    sincospi(2.0 * u2, &sn, &cs);
    double2 o; o.x = rad * cs; o.y = rad * sn;
    return o;
}
// out[i] = mean + sigma * N(0,1), i < count (real doubles; pairs share one Philox call)
__global__ void k_gaussian(double* __restrict__ out, long long count, double mean, double sigma, unsigned long long seed, unsigned int sub) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (2 * i >= count) return;
    const double2 g = philox_normal2(seed, sub, (unsigned long long)i);
    out[2 * i] = mean + sigma * g.x;
    if (2 * i + 1 < count) out[2 * i + 1] = mean + sigma * g.y;
}
// EDFA: out[row][pol][i] = sqrt(G) in[row or 0][pol][i] + sigma (n1 + j n2)   (devices.py:921-936; y polarisation of a
// one-polarisation input is zero signal + ASE).  in_rows = 1 broadcasts one input waveform to every realisation.
__global__ void k_edfa(const double2* __restrict__ in, double2* __restrict__ out, long long rows, long long in_rows, int in_pol,
                       int out_pol, long long n, double g_amp, double sigma, unsigned long long seed, long long first_row) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * out_pol * n) return;
    const long long row = i / (out_pol * n);
    const int pol = (int)((i / n) % out_pol);
    const long long k = i % n;
    double2 e; e.x = 0.0; e.y = 0.0;
    if (pol < in_pol) { const double2 a = in[((in_rows == 1 ? 0 : row) * in_pol + pol) * n + k]; e.x = g_amp * a.x; e.y = g_amp * a.y; }
    // the counter is the element's index in the WHOLE batch, so a batch generated chunk by chunk equals the batch generated at once
    const double2 g = philox_normal2(seed, 0x45444641u, (unsigned long long)(first_row * out_pol * n + i));
    e.x += sigma * g.x; e.y += sigma * g.y;
    out[i] = e;
}

}  // namespace ssfm_filt

namespace {

using namespace ssfm_filt;

void pool_keeps_memory(int device) {                 // stream-ordered workspaces: do not give the memory back at every sync
    static bool done[64] = {false};
    if (device < 0 || device >= 64 || done[device]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[device] = true;
}

struct SideStream { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, join = nullptr; };
SideStream& side_of(int device) {
    static SideStream s[64];
    SideStream& x = s[(device >= 0 && device < 64) ? device : 0];
    if (!x.st) {
        cudaStreamCreateWithFlags(&x.st, cudaStreamNonBlocking);
        cudaEventCreateWithFlags(&x.fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&x.join, cudaEventDisableTiming);
    }
    return x;
}

// One chunk of rows stays L2-resident across the passes over it: 24 ... 40 MiB, and within that range the row count whose
// transform launches (n / 4096 CTAs per row, two 256-thread CTAs per SM) come closest to whole waves of CTAs -- a launch of
// 1.7 waves leaves the chip 15 % idle, and every chunk pays that three times.
long long chunk_rows(long long rows, long long n, int device) {
    const size_t row_bytes = 16 * (size_t)n;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const double slots = 2.0 * sms, per_row = (double)std::max<long long>(1, n / 4096);
    const long long lo = std::max<long long>(1, (long long)((24u << 20) / row_bytes)), hi = std::max<long long>(lo, (long long)((40u << 20) / row_bytes));
    long long best = lo;
    double best_eff = 0.0;
    for (long long r = lo; r <= hi; ++r) {
        const double w = r * per_row / slots, eff = w / std::ceil(w);
        if (eff >= best_eff - 1e-12) { best_eff = std::max(best_eff, eff); best = r; }
    }
    return std::max<long long>(1, std::min(rows, best));
}

constexpr size_t OLS_SMEM = sizeof(double2) * (size_t)ssfm::RowExchange<OLS_M, OLS_E>::size;
// pass tables of the OLS_M-point transform, one copy per device (64 KB, read through L1 by k_ols: measured on B200, a folded
// 8 KB table of the unit circle in shared memory was slower -- BPF of 1024 x 2^18 samples 3.84 against 3.47 ms)
const double2* ols_tables(int device, cudaStream_t st) {
    static std::mutex mu;
    static void* tab[64] = {nullptr};
    std::lock_guard<std::mutex> lock(mu);
    void*& t = tab[(device >= 0 && device < 64) ? device : 0];
    if (!t) {
        if (ssfm_internal_pass_tables_f64(&t, OLS_M, st) != SSFM_OK) { t = nullptr; return nullptr; }
        cudaFuncSetAttribute(k_ols<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OLS_SMEM);
        cudaFuncSetAttribute(k_ols<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)OLS_SMEM);
    }
    return (const double2*)t;
}

// The whole zero-phase filter of `rows` rows of n samples.  Source: complex rows x (pd == null) or the photodetector
// front end *pd; destination: complex rows y (may alias x), or -- out_sig != null -- real rows out_sig / out_noise
// holding samples offset, offset + stride, ... (m per row).  `y` is then a scratch buffer of >= chunk rows.
int zero_phase(const Sos& f, double rho, const double2* x, const PdSrc* pd, double2* y, double* out_sig, double* out_noise,
               long long rows, long long n, long long offset, long long stride, long long m, int device, cudaStream_t st) {
    cudaError_t e = cudaSuccess;
    int rc = SSFM_OK;
    pool_keeps_memory(device);
    // transient length: rho^K < 1e-19, with head-room for the polynomial factors of clustered Bessel poles
    long long K = -1;
    if (rho > 0 && rho < 1) K = (long long)std::ceil(1.3 * std::log(1e-19) / std::log(rho)) + 64;
    else if (rho == 0) K = 64;
    const bool pow2 = (n & (n - 1)) == 0 && n >= 256 && n <= (1ll << 22);
    const bool ends_fit = K > 0 && 4 * K + 2 * f.edge <= n && !getenv("SSFM_FILTFILT_SEQUENTIAL");
    // overlap-save (k_ols): blocks of 4096 samples with a halo of Kp >= K on either side, any row length; taken while at
    // least half of every block is output (Kp <= 1024)
    const long long Kp = (K + 7) & ~7ll;
    const bool ols = ends_fit && Kp <= 1024 && n >= OLS_M && !getenv("SSFM_FILTFILT_NO_OLS");
    const bool fft_path = ends_fit && (pow2 || ols);
    const bool real_out = out_sig != nullptr;
    // rows of 2^12 .. 2^20 samples go through the persistent kernel in ONE launch (it keeps the rows in flight L2-resident by
    // itself; at most 8 GiB of packed rows at a time); other lengths through the three streaming kernels in L2-sized chunks
    // (small batches leave most teams of the persistent kernel idle: measured on B200, 16 rows of 2^18 samples take 0.57 ms in
    // one launch against 0.48 ms in chunks, 1024 rows 8.3 against 11.2 ms -- so the one-launch path starts at 2^25 samples)
    // (the photodetector path packs the whole batch first, which costs it the chunks' L2 residency: 128 rows of 2^18 take 2.06 ms
    // in one launch against 1.75 ms in chunks, 1024 rows 10.1 against 12.4 ms -- its threshold is 2^27 samples)
    const bool one_launch = fft_path && !ols && n >= 4096 && n <= (1ll << 20) && rows * n >= (pd ? (1ll << 27) : (1ll << 25)) &&
                            !getenv("SSFM_TRANSFER_MULTILAUNCH");
    const long long chunk = !fft_path ? rows
                          : ols ? std::max<long long>(1, std::min<long long>(rows, (long long)((32u << 20) / (16 * (size_t)n))))
                          : one_launch ? std::max<long long>(1, std::min<long long>(rows, (long long)((8ull << 30) / (16 * (size_t)n))))
                                       : chunk_rows(rows, n, device);

    double* ws = nullptr;
    double2 *seg = nullptr, *edge_out = nullptr, *scratch = nullptr;
    auto cleanup = [&]() {
        if (ws) cudaFreeAsync(ws, st);
        if (seg) cudaFreeAsync(seg, st);
        if (edge_out) cudaFreeAsync(edge_out, st);
        if (scratch) cudaFreeAsync(scratch, st);
    };
    auto cuda_fail = [&](cudaError_t err) { cleanup(); ssfm_err_slot = std::string("filtfilt: ") + cudaGetErrorString(err); return SSFM_ERR_CUDA; };

    const bool in_place = !pd && !real_out && (const void*)x == (const void*)y;
    if (ols ? in_place : (real_out || (!fft_path && pd))) {   // packed rows live in a scratch buffer (one chunk; all rows on the sequential
                                                       // path); overlap-save needs one only in place (blocks read their neighbours' halos)
        e = cudaMallocAsync((void**)&scratch, sizeof(double2) * (size_t)chunk * n, st);
        if (e != cudaSuccess) return cuda_fail(e);
    }
    if (fft_path) {
        const int L = (int)(2 * K), Ki = (int)K;
        static std::mutex enqueue_mu;                  // the side stream and its two events are shared by the callers of a device
        std::lock_guard<std::mutex> lock(enqueue_mu);
        SideStream& sd = side_of(device);
        e = cudaMallocAsync((void**)&seg, sizeof(double2) * (size_t)rows * 2 * L, st);
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&edge_out, sizeof(double2) * (size_t)rows * 2 * Ki, st);
        if (e == cudaSuccess) e = cudaMallocAsync((void**)&ws, sizeof(double) * (size_t)rows * 4 * (L + f.edge), st);
        if (e != cudaSuccess) return cuda_fail(e);
        {   // end segments of every row, then their recursions on the side stream while the chunks go through
            const long long cnt = rows * 2 * L;
            if (pd) k_pd_save_edges<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(*pd, seg, rows, n, L);
            else k_save_edges<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(x, seg, rows, n, L);
            cudaEventRecord(sd.fork, st);
            cudaStreamWaitEvent(sd.st, sd.fork, 0);
            const long long threads = rows * 4;
            k_filtfilt_edges_out<<<(unsigned)((threads + 63) / 64), 64, 0, sd.st>>>((const double*)seg, (double*)edge_out, ws, rows, L, Ki, f);
            cudaEventRecord(sd.join, sd.st);
        }
        void* plan = nullptr;
        if (ols) {
            const double2* tw = ols_tables(device, st);
            double* h2 = nullptr;
            if (!tw) rc = SSFM_ERR_CUDA;
            if (!rc && (e = cudaMallocAsync((void**)&h2, sizeof(double) * OLS_M, st)) != cudaSuccess) { cudaStreamWaitEvent(st, sd.join, 0); return cuda_fail(e); }
            if (!rc) {
                k_fill_h2<<<OLS_M / 256, 256, 0, st>>>(h2, f);
                OlsArgs a{};
                a.tw = tw; a.h2 = h2; a.n = n; a.offset = offset; a.stride = stride; a.m = m;
                a.K = (int)Kp; a.L = OLS_M - 2 * (int)Kp; a.nb = (int)((n + a.L - 1) / a.L);
                const size_t smem = OLS_SMEM;
                {
                    int sms = 148;
                    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
                    a.wave = 2 * sms;
                }
                // sampler with a large stride: evaluate the response as a FIR filter at the wanted samples only (k_pd_fir); its
                // taps are what the block kernel returns for a unit impulse
                const bool fir = pd && real_out && stride >= 32 && stride <= 64 && Kp <= 1000 && !getenv("SSFM_PD_NO_FIR");
                if (fir) {
                    double2* imp = nullptr;
                    if ((e = cudaMallocAsync((void**)&imp, sizeof(double2) * 2 * OLS_M, st)) != cudaSuccess) {
                        cudaFreeAsync(h2, st); cudaStreamWaitEvent(st, sd.join, 0); return cuda_fail(e);
                    }
                    cudaMemsetAsync(imp, 0, sizeof(double2) * 2 * OLS_M, st);
                    const double one = 1.0;
                    cudaMemcpyAsync(&imp[Kp].x, &one, sizeof(double), cudaMemcpyHostToDevice, st);   // impulse at sample K of a 4096-sample row
                    OlsArgs ia = a;
                    ia.x = imp; ia.y = imp + OLS_M; ia.n = OLS_M; ia.nb = (OLS_M + ia.L - 1) / ia.L; ia.row0 = 0; ia.wave = 1 << 30;
                    ia.offset = 0; ia.stride = 1; ia.m = OLS_M;
                    k_ols<false><<<(unsigned)ia.nb, OLS_NT, smem, st>>>(ia);
                    FirArgs fa{};
                    fa.pd = *pd; fa.taps = imp + OLS_M; fa.out_sig = out_sig; fa.out_noise = out_noise;
                    fa.n = n; fa.m = m; fa.offset = offset; fa.stride = stride; fa.row0 = 0; fa.K = (int)Kp;
                    fa.kb_total = (int)(((2 * Kp + 1 + stride - 1) / stride + 3) & ~3ll);   // padded to a multiple of 4 with zero taps
                    fa.AQ = (((FIR_TO + fa.kb_total + 2 + 3) / 4) + 1) & ~1;        // columns 0 .. FIR_TO + kb_total + 1 of X, de-interleaved by 4
                    const size_t fsm = sizeof(double2) * (size_t)stride * (4 * fa.AQ + 1) + sizeof(double) * (size_t)fa.kb_total * stride;
                    const bool simple = pd->n_pol == 1 && !pd->noise;
                    static bool fir_attr[64] = {false};
                    if (!fir_attr[(device >= 0 && device < 64) ? device : 0]) {
                        cudaFuncSetAttribute(k_pd_fir<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
                        cudaFuncSetAttribute(k_pd_fir<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 113 * 1024);
                        fir_attr[(device >= 0 && device < 64) ? device : 0] = true;
                    }
                    const long long per_row = (m + FIR_TO - 1) / FIR_TO;
                    const long long rows_per = std::max<long long>(1, std::min<long long>(rows, 0x7fffffffll / per_row));
                    if (fsm <= 113 * 1024) {
                        for (long long r0 = 0; r0 < rows; r0 += rows_per) {
                            fa.row0 = r0;
                            const unsigned grid = (unsigned)(std::min(rows_per, rows - r0) * per_row);
                            if (simple) k_pd_fir<true><<<grid, FIR_NT, fsm, st>>>(fa);
                            else k_pd_fir<false><<<grid, FIR_NT, fsm, st>>>(fa);
                        }
                    }
                    e = cudaGetLastError();
                    cudaFreeAsync(imp, st);
                    if (fsm > 113 * 1024) e = cudaErrorInvalidValue;
                    if (e != cudaSuccess) { cudaFreeAsync(h2, st); cudaStreamWaitEvent(st, sd.join, 0); return cuda_fail(e); }
                }
                const long long per = in_place ? chunk : std::max<long long>(1, std::min<long long>(rows, (long long)(0x7fffffffll / a.nb)));
                for (long long r0 = 0; r0 < rows && !fir; r0 += per) {
                    const long long nr = std::min(per, rows - r0);
                    a.row0 = r0;
                    if (pd) {
                        a.pd = *pd; a.y = nullptr; a.out_sig = out_sig; a.out_noise = out_noise;
                        k_ols<true><<<(unsigned)(nr * a.nb), OLS_NT, smem, st>>>(a);
                    } else {
                        a.x = x + (size_t)r0 * n; a.y = in_place ? scratch : y + (size_t)r0 * n;
                        k_ols<false><<<(unsigned)(nr * a.nb), OLS_NT, smem, st>>>(a);
                        if (in_place) cudaMemcpyAsync(y + (size_t)r0 * n, scratch, sizeof(double2) * (size_t)nr * n, cudaMemcpyDeviceToDevice, st);
                    }
                }
                e = cudaGetLastError();
                cudaFreeAsync(h2, st);
                if (e != cudaSuccess) { cudaStreamWaitEvent(st, sd.join, 0); return cuda_fail(e); }
            }
        } else
        rc = ssfm_internal_transfer_prepare(device, n, chunk, f, &plan, st);
        for (long long r0 = 0; r0 < rows && !rc && !ols; r0 += chunk) {
            const long long nr = std::min(chunk, rows - r0);
            double2* buf = real_out ? scratch : y + (size_t)r0 * n;
            const double2* src = nullptr;
            if (pd) k_pd_pack<<<(unsigned)((nr * n + 255) / 256), 256, 0, st>>>(*pd, buf, r0, nr, n);
            else if (buf != x + (size_t)r0 * n) src = x + (size_t)r0 * n;      // out of place: the transfer reads the source itself
            rc = ssfm_internal_transfer_apply(plan, buf, nr, st, src, one_launch ? 1 : 0);
            if (!rc && real_out)
                k_unpack_real<<<(unsigned)((nr * m + 255) / 256), 256, 0, st>>>(buf, out_sig, out_noise, r0, nr, n, offset, stride, m);
        }
        cudaStreamWaitEvent(st, sd.join, 0);             // (also on the error path: the workspaces are freed on `st`)
        if (!rc) {
            const long long cnt = rows * 2 * Ki;
            k_scatter_edges<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(edge_out, real_out ? nullptr : y, out_sig, out_noise, rows, n, Ki,
                                                                        offset, stride, m);
            e = cudaGetLastError();
        }
    } else {
        const size_t len = (size_t)n + 2 * (size_t)f.edge;
        e = cudaMallocAsync((void**)&ws, sizeof(double) * len * (size_t)rows * 2, st);
        if (e != cudaSuccess) return cuda_fail(e);
        const double2* src = x;
        if (pd) { k_pd_pack<<<(unsigned)((rows * n + 255) / 256), 256, 0, st>>>(*pd, scratch, 0, rows, n); src = scratch; }
        double2* dst = real_out ? scratch : y;
        if (real_out && !pd) { cleanup(); ssfm_err_slot = "filtfilt: real outputs need the photodetector front end"; return SSFM_ERR_INVALID; }
        const long long threads = rows * 2;
        k_filtfilt_seq<<<(unsigned)((threads + 63) / 64), 64, 0, st>>>((const double*)src, (double*)dst, ws, rows, n, f);
        if (real_out) k_unpack_real<<<(unsigned)((rows * m + 255) / 256), 256, 0, st>>>(dst, out_sig, out_noise, 0, rows, n, offset, stride, m);
        e = cudaGetLastError();
    }
    cleanup();
    if (rc) return rc;
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { ssfm_err_slot = std::string("filtfilt: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }
    return SSFM_OK;
}

int parse_filter(Sos& f, double* rho, const double* sos_host, int n_sections, long long n) {
    std::string err;
    if (!sos_host) { ssfm_err_slot = "null sos"; return SSFM_ERR_INVALID; }
    int rc = make_sos(f, sos_host, n_sections, rho, err);
    if (rc) { ssfm_err_slot = err; return rc; }
    if (n <= f.edge) {
        ssfm_err_slot = "The length of the input vector x must be greater than padlen, which is " + std::to_string(f.edge) + ".";
        return SSFM_ERR_INVALID;
    }
    return SSFM_OK;
}

}  // namespace

extern "C" int ssfm_filtfilt_sos(void* x_dev, void* y_dev, int64_t n_rows, int64_t n, const double* sos_host,
                                 int32_t n_sections, int32_t device, void* stream) {
    if (!x_dev || !y_dev || !sos_host) { ssfm_err_slot = "null buffer or sos"; return SSFM_ERR_INVALID; }
    if (n_rows < 1 || n < 1) { ssfm_err_slot = "n_rows and n_samples must be >= 1"; return SSFM_ERR_INVALID; }
    Sos f;
    double rho = 0;
    int rc = parse_filter(f, &rho, sos_host, n_sections, n);
    if (rc) return rc;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { ssfm_err_slot = std::string("filtfilt: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }
    return zero_phase(f, rho, (const double2*)x_dev, nullptr, (double2*)y_dev, nullptr, nullptr, n_rows, n, 0, 1, n, device,
                      (cudaStream_t)stream);
}

extern "C" int ssfm_pd_lpf(const void* field_dev, const void* noise_dev, const double* extra_noise_dev, double* out_signal_dev,
                           double* out_noise_dev, int64_t n_rows, int32_t n_pol, int64_t n, double responsivity, double r_load,
                           double i_dark, const double* sos_host, int32_t n_sections, int64_t sample_offset,
                           int64_t sample_stride, int32_t device, void* stream) {
    if (!field_dev || !out_signal_dev) { ssfm_err_slot = "null field or output"; return SSFM_ERR_INVALID; }
    if (n_rows < 1 || n < 1 || (n_pol != 1 && n_pol != 2)) { ssfm_err_slot = "n_rows, n_samples >= 1 and n_pol in {1, 2} expected"; return SSFM_ERR_INVALID; }
    if (sample_stride < 1 || sample_offset < 0 || sample_offset >= n) { ssfm_err_slot = "sampler: 0 <= offset < n_samples and stride >= 1 expected"; return SSFM_ERR_INVALID; }
    if ((noise_dev || extra_noise_dev) && !out_noise_dev) { ssfm_err_slot = "noise inputs need out_noise"; return SSFM_ERR_INVALID; }
    Sos f;
    double rho = 0;
    int rc = parse_filter(f, &rho, sos_host, n_sections, n);
    if (rc) return rc;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { ssfm_err_slot = std::string("pd_lpf: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }
    PdSrc s;
    s.field = (const double2*)field_dev; s.noise = (const double2*)noise_dev; s.extra = extra_noise_dev;
    s.r = responsivity; s.r_load = r_load; s.i_dark = out_noise_dev ? i_dark : 0.0; s.n_pol = n_pol;
    s.noise_out = out_noise_dev ? 1 : 0;
    const long long m = (n - sample_offset + sample_stride - 1) / sample_stride;
    return zero_phase(f, rho, nullptr, &s, nullptr, out_signal_dev, out_noise_dev, n_rows, n, sample_offset, sample_stride, m, device,
                      (cudaStream_t)stream);
}

extern "C" int ssfm_gaussian_noise(double* out_dev, int64_t count, double mean, double sigma, uint64_t seed, uint32_t substream,
                                   int32_t device, void* stream) {
    if (!out_dev || count < 0) { ssfm_err_slot = "null buffer or negative count"; return SSFM_ERR_INVALID; }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess && count > 0) {
        const long long pairs = (count + 1) / 2;
        ssfm_filt::k_gaussian<<<(unsigned)((pairs + 255) / 256), 256, 0, (cudaStream_t)stream>>>(out_dev, count, mean, sigma, seed, substream);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { ssfm_err_slot = std::string("gaussian_noise: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }
    return SSFM_OK;
}

extern "C" int ssfm_edfa(const void* in_dev, void* out_dev, int64_t n_rows, int64_t in_rows, int32_t in_pol, int32_t out_pol,
                         int64_t n, double gain_db, double p_ase_w, uint64_t seed, int64_t first_row, int32_t device, void* stream) {
    if (!in_dev || !out_dev) { ssfm_err_slot = "null buffer"; return SSFM_ERR_INVALID; }
    if (n_rows < 1 || n < 1 || (in_rows != 1 && in_rows != n_rows) || in_pol < 1 || in_pol > 2 || out_pol < in_pol || out_pol > 2) {
        ssfm_err_slot = "edfa: rows >= 1, in_rows in {1, rows}, 1 <= in_pol <= out_pol <= 2 expected"; return SSFM_ERR_INVALID;
    }
    if (p_ase_w < 0) { ssfm_err_slot = "edfa: negative ASE power"; return SSFM_ERR_INVALID; }
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        const long long cnt = n_rows * out_pol * n;
        const double g_amp = std::sqrt(std::pow(10.0, gain_db / 10.0));          // np.sqrt(idb(G)), devices.py:921
        const double sigma = std::sqrt(p_ase_w / 4.0);                            // np.sqrt(P_ase/4), devices.py:933
        ssfm_filt::k_edfa<<<(unsigned)((cnt + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const double2*)in_dev, (double2*)out_dev, n_rows, in_rows,
                                                                                        in_pol, out_pol, n, g_amp, sigma, seed, first_row);
        e = cudaGetLastError();
    }
    if (e != cudaSuccess) { ssfm_err_slot = std::string("edfa: ") + cudaGetErrorString(e); return SSFM_ERR_CUDA; }
    return SSFM_OK;
}
