// Split-step Fourier kernels (FIBER / DBP hot path) for sm_100a.
//
// One propagation step of reference opticomlib/devices.py:1172-1196 over a batch of waveforms,
// each waveform = P polarisation rows of N = N1*N2 complex samples viewed as an N1 x N2 matrix
// (sample n = n1*N2 + n2):
//
//   k_col_fwd : first Kerr half step  A *= exp(j (h/2) gamma |A|^2)           (devices.py:1175-1177)
//               + stash of the phase (the reference reuses the start-of-step N^ for the second half)
//               + N1-point transforms down the columns + four-step twiddle W_N^{n2*k1}
//   k_row     : N2-point transforms along the rows (spectrum left in transposed order,
//               bin k = k1 + N1*k2 at [k1][k2]), linear operator exp(D~(w_k) h)  (devices.py:1145,1179),
//               inverse N2-point transforms                                      (devices.py:1178-1180)
//   k_col_inv : conjugate twiddle + inverse N1-point column transforms, 1/N,
//               second Kerr half step with the stashed phase (devices.py:1181),
//               max |A|^2 per waveform, and -- in the last tile of a waveform to finish --
//               the float32/float64 step-size controller                        (devices.py:1173,1193-1196)
//
// Memory traffic per sample and step: 3 field reads + 3 field writes + 1 stash write + 1 stash read.
#pragma once
#include "fft_core.cuh"
#include <type_traits>

namespace ssfm {

// Per-waveform controller state.  z and h are stored as double but always hold values of the
// compute real type R (float values are exactly representable), so the bookkeeping is bit-exact
// float32 in fp32 mode, as in the reference (devices.py:1155-1162, 1193-1196).
struct Ctrl {
    double z;                   // position reached [km]
    double h;                   // size of the NEXT step [km]
    unsigned long long pmax;    // max |A|^2 accumulator (bit pattern of R, ordered as unsigned)
    int steps;                  // steps taken
    int done;                   // z >= length
    unsigned int arrived;       // tiles of this waveform that finished k_col_inv in the current step
    int pad;
};

template <typename R>
struct Params {
    typedef typename cx_of<R>::type C;
    C* field;            // [B][P][N], updated in place
    const C* field_in;   // k_wf only: read the waveforms from here instead (out-of-place transfer functions); null = field
    R* stash;            // [B][P][N] Kerr phase of the current step
    Ctrl* ctrl;          // [B]
    int* active;         // number of waveforms with done == 0
    unsigned int* ticket;// start-order ticket counter of the fused column kernel
    unsigned long long* slots;   // [B][tiles*P][2] self-validating max words of SYNC_LL
    double* hlog;        // [B][hlog_cap] step sizes actually taken (may be null)
    const C* tw_col;     // pass tables of the N1-point transform
    const C* tw_row;     // pass tables of the N2-point transform
    const C* tw_lo;      // W_N^i,          i < 2^lo_bits
    const C* tw_hi;      // W_N^(i*2^lo_bits)
    const C* tw_full;    // optional full four-step table W_N^{n2*k1} at [k1][n2] (N <= 2^20): one coalesced L2 load
                         // instead of two table look-ups and a complex product
    int lin_sep;         // k_wf<double> with rows of 256 bins and beta_3 = 0: linear operator in separable form (ssfm_wf.cuh: lin_sep_table)
    int tw_chain;        // 1: four-step twiddles by recurrence (two interleaved chains seeded from tw_lo / tw_hi): 16 complex
                         // products per thread instead of 16 loads of 16 B from L2 -- the full table is 2 x 16 B of L2 traffic per
                         // sample and step next to the field's 4 x 16 B, and the L2 <-> SM path is what k_wf<double> waits for
    int l2_ahead;        // > 0: each CTA of the fused column kernel prefetches into L2 the tile that the CTA
                         // `l2_ahead` tickets later will load (about one CTA lifetime ahead)
    int small_phase;     // 1: every Kerr phase is <= 0.05 rad (adaptive mode with phi_max <= 0.05): short Taylor sincos
    const C* xfer;       // optional transfer function H[k] in transposed order ([k1][k2], bin k1 + N1*k2):
                         // when set, the row kernel multiplies by it instead of exp(D~ h) (filters, DM)
    int lo_bits;
    // long waveforms (N = N0 x N_l, ssfm_long.cu): the OUTER launches see one "waveform" of N0 rows x n2 columns (this
    // rank's slice of the N_l columns, global column = n2 + n2_off, twiddle tables built for the global N); the INNER
    // launches (inner = 1) transform the N0/G rows of N_l samples this rank owns after the exchange: every row shares
    // the controller state of waveform 0 (h, done), the row kernel maps its bins to global bins
    // k = bin_off + row + bin_mul (k1 + N1 k2) with the fftfreq wrap at n_glob, and k_col_inv leaves the controller alone.
    int n2_off, inner, bin_mul, bin_off, n_glob;
    // long waveforms spread over several GPUs, exchange fused into the kernels: instead of storing a stage's result locally
    // (and moving it with a collective afterwards) the column kernels store every element straight into the buffer of the
    // rank that owns it in the NEXT stage's layout -- peer memory over NVLink (CUDA IPC), 16-byte elements in runs of T
    // columns.  peer_mode 1 (outer stage -> rows layout): element (ka, nb) goes to rank ka >> peer_shift at
    // [ka & mask][peer_base + nb] with row pitch peer_pitch = N_l; peer_mode 2 (inner k_col_inv -> time layout): element nb of
    // my row goes to rank nb >> peer_shift at [peer_base + row][nb & mask] with pitch peer_pitch = columns per rank.
    typename cx_of<R>::type* peer[8];
    int peer_mode, peer_shift, peer_pitch, peer_base;
    const R* dim_tab;    // k_wf: imag(D~(w_k)) per bin in transposed order [k1][k2], filled once per propagation by k_fill_dim with
                         // exactly the operations of the inline evaluation (bit-identical), so that the row phase spends one
                         // L2 load instead of ~10 arithmetic instructions per sample and step
    int fwd_only;        // 1: the row kernel stops after the forward transforms and stores the spectrum (transposed order):
                         // used once per plan to build the chirp spectra of the arbitrary-length transform
    int defer_ctrl;      // 1: k_col_inv only accumulates max|A|^2 in ctrl.pmax; the controller runs later (k_ctrl_step), after
                         // the maxima of all ranks have been combined
    int n, n1, n2, log2_n2;
    int n_pol;
    int batch;           // waveforms in this launch
    int hlog_cap;
    int adaptive;        // 1: h = phi_max / max(|gamma| |A|^2) after every step
    int has_nl;          // gamma != 0
    int max_steps;       // safety stop
    int debug;           // bit0: no barrier spin (timing experiments only), bit1: blockIdx instead of tickets
    R gamma, abs_gamma, phi_max, length;
    R att_half;          // -alpha_lin/2            (devices.py:1145)
    R c2;                // imag(1j/2*beta_2)       (devices.py:1145)
    R c3;                // imag(1j/6*beta_3)
    double wscale;       // ((1.0/(N*dt))*2*pi)*1e-12: omega grid of typing.py:1641 and devices.py:1144 folded
                         // into one factor (differs from the three separate roundings by <= 2 ulp of
                         // float64, i.e. < 1e-12 rad on the largest phase; invisible after the cast to float32)
    R inv_n;
};

// ---- rounding-exact scalar helpers (no FMA contraction where the reference has none) ----------
__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
// sin/cos for the phase rotations.  float: CUDA's sincosf (full range, <= 2 ulp).
// double: 256-entry table of exp(j 2 pi i/256) in shared memory + a three-part Cody-Waite reduction
// to |r| <= pi/256 + degree-5/6 Taylor polynomials: 17 FP64 instructions instead of ~40 for
// sincos(), absolute error < 3e-16 for |x| < 1e7 rad (the linear-operator phase is O(1e3) rad).
constexpr int SC_N = 256;
// float: the argument is the float32 phase the reference feeds to exp() (O(1e3) rad for the linear operator), so the
// reduction must be exact for THAT value: it is done in double (two FMAs), the rest in float32 -- 256-entry table +
// degree-3/4 Taylor on |r| <= pi/256 (truncation < 3e-12), about 16 instructions instead of ~45 for sincosf's
// slow path; absolute error < 1.5e-7.
__device__ __forceinline__ void sincos_r(float x, const float2* tab, float* s, float* c) {
    const double magic = 6755399441055744.0;                       // 1.5 * 2^52
    const double xd = (double)x;
    const double m = fma(xd, 40.74366543152521, magic);            // x * 256/(2 pi), integer part in the low bits
    const int idx = __double2loint(m) & (SC_N - 1);
    const float r = (float)fma(-(m - magic), 0.02454369260617026, xd);   // x - k * 2 pi/256
    const float r2 = r * r;
    const float sr = fmaf(r * r2, -1.6666667e-1f, r);
    const float cr = fmaf(r2, fmaf(r2, 4.1666668e-2f, -0.5f), 1.0f);
    const float2 e = tab[idx];                                     // (cos, sin) of 2 pi idx/256
    *c = fmaf(e.x, cr, -(e.y * sr));
    *s = fmaf(e.y, cr, e.x * sr);
}
__device__ __forceinline__ void sincos_r(double x, const double2* tab, double* s, double* c) {
    const double magic = 6755399441055744.0;                       // 1.5 * 2^52
    const double m = fma(x, 40.74366543152521, magic);             // x * 256/(2 pi), integer part in the low bits
    const int idx = __double2loint(m) & (SC_N - 1);
    const double k = m - magic;
    double r = fma(-k, 0x1.921fb54400000p-6, x);                   // 2 pi/256 = C1 + C2 + C3
    r = fma(-k, 0x1.0b4611a600000p-40, r);
    r = fma(-k, 0x1.3198a2e037073p-75, r);
    const double r2 = r * r;
    const double sr = fma(r * r2, fma(r2, 8.3333333333333332e-3, -1.6666666666666666e-1), r);
    const double cr = fma(r2, fma(r2, fma(r2, -1.3888888888888889e-3, 4.1666666666666664e-2), -0.5), 1.0);
    const double2 e = tab[idx];                                    // (cos, sin) of 2 pi idx/256
    *c = fma(e.x, cr, -(e.y * sr));
    *s = fma(e.y, cr, e.x * sr);
}
// |x| <= 0.05 rad: Taylor series to x^7 / x^8 (truncation < 1e-17), 9 instructions, no table
template <typename R> __device__ __forceinline__ void sincos_small(R x, R* s, R* c) {
    const R x2 = x * x;
    *s = x + x * x2 * ((R)-1.6666666666666666e-1 + x2 * ((R)8.3333333333333332e-3 + x2 * (R)-1.9841269841269841e-4));
    *c = (R)1 + x2 * ((R)-0.5 + x2 * ((R)4.1666666666666664e-2 + x2 * ((R)-1.3888888888888889e-3 + x2 * (R)2.4801587301587302e-5)));
}
// compile-time selection; the kernels unswitch their unrolled Kerr loops on Params::small_phase with
//   auto body = [&](auto small) { ... kerr_sincos<decltype(small)::value>(...) ... };
//   if (p.small_phase) body(std::true_type{}); else body(std::false_type{});
template <bool SMALL, typename R, typename C>
__device__ __forceinline__ void kerr_sincos(R x, const C* tab, R* s, R* c) {
    if constexpr (SMALL) sincos_small<R>(x, s, c); else sincos_r(x, tab, s, c);
}
__device__ __forceinline__ float  exp_r(float x)  { return expf(x); }
__device__ __forceinline__ double exp_r(double x) { return exp(x); }
// w**3 as numpy.power gives it (correctly rounded cube): exact product in higher precision
__device__ __forceinline__ float  cube_r(float w)  { double d = (double)w; return (float)(d * d * d); }
__device__ __forceinline__ double cube_r(double w) { return w * w * w; }

__device__ __forceinline__ unsigned long long ord_bits(float p)  { return (unsigned long long)__float_as_uint(p); }
__device__ __forceinline__ unsigned long long ord_bits(double p) { return (unsigned long long)__double_as_longlong(p); }
template <typename R> __device__ __forceinline__ R from_bits(unsigned long long b);
template <> __device__ __forceinline__ float  from_bits<float>(unsigned long long b)  { return __uint_as_float((unsigned int)b); }
template <> __device__ __forceinline__ double from_bits<double>(unsigned long long b) { return __longlong_as_double((long long)b); }

template <typename R> __device__ __forceinline__ R pw_nan();
template <> __device__ __forceinline__ float  pw_nan<float>()  { return __uint_as_float(0x7fc00000u); }
template <> __device__ __forceinline__ double pw_nan<double>() { return __longlong_as_double(0x7ff8000000000000ll); }

// NaN-propagating max on the bit pattern: NaNs have the largest patterns among non-negative values,
// so a NaN power wins the atomicMax and poisons h exactly like numpy's max would.
template <typename R> __device__ __forceinline__ R block_max_bits(R v, unsigned long long* red) {
    unsigned long long b = ord_bits(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
        b = other > b ? other : b;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
    if (lane == 0) red[warp] = b;
    __syncthreads();
    if (warp == 0) {
        b = lane < nwarps ? red[lane] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
            b = other > b ? other : b;
        }
        if (lane == 0) red[0] = b;
    }
    __syncthreads();
    return from_bits<R>(red[0]);
}

// Step-size controller (devices.py:1173, 1193-1196) as a pure function of the state before the
// step and the waveform's max |A|^2 after it; every rounding is one IEEE operation in R.
template <typename R> struct CtrlNext { R z, h; int done; };
template <typename R>
__device__ __forceinline__ CtrlNext<R> controller_next(const Params<R>& p, R z, R h, int steps, R pmax) {
    CtrlNext<R> n;
    n.z = add_rn(z, h);                                           // z += h_
    R hn = h;
    if (p.adaptive) hn = p.phi_max / mul_rn(p.abs_gamma, pmax);   // phi_max / max(|gamma| |A|^2)
    const R rem = p.length - n.z;
    n.h = (rem < hn) ? rem : hn;                                  // python min(h_, length - z)
    n.done = !(n.z < p.length) || (steps + 1 >= p.max_steps);
    return n;
}
// One thread per waveform per step stores the new state.
template <typename R>
__device__ __forceinline__ void controller_commit(const Params<R>& p, int b, R h_taken, int steps, const CtrlNext<R>& n) {
    Ctrl& c = p.ctrl[b];
    if (p.hlog && steps < p.hlog_cap) p.hlog[(size_t)b * p.hlog_cap + steps] = (double)h_taken;
    c.z = (double)n.z; c.h = (double)n.h; c.pmax = 0ull; c.arrived = 0u;
    c.done = n.done;
    if (n.done) atomicSub(p.active, 1);
    __threadfence();                                              // publish the new state, then the step count
    *reinterpret_cast<volatile int*>(&c.steps) = steps + 1;       // (waiters of k_col_mid<SYNC_GLOBAL> spin on it)
}
template <typename R>
__device__ __forceinline__ void controller_update(const Params<R>& p, int b, R pmax) {
    const Ctrl& c = p.ctrl[b];
    const R h = (R)c.h;
    const int s = c.steps;
    const CtrlNext<R> n = controller_next<R>(p, (R)c.z, h, s, pmax);
    controller_commit<R>(p, b, h, s, n);
}

// imag(D~(w_k)) = imag(1j/2 beta_2) w^2 + imag(1j/6 beta_3) w^3 on the fftfreq grid (devices.py:1144-1145), one value per bin in
// transposed order: pos = k1 * N2 + k2 holds bin k = k1 + N1 k2
template <typename R>
__global__ void k_fill_dim(Params<R> p, R* out) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= p.n) return;
    int k = pos / p.n2 + p.n1 * (pos % p.n2);
    k = (k < (p.n >> 1)) ? k : k - p.n;
    const R wk = (R)((double)k * p.wscale);
    out[pos] = add_rn(mul_rn(p.c2, mul_rn(wk, wk)), mul_rn(p.c3, cube_r(wk)));
}

// deferred controller step of waveform 0 (long waveforms: ctrl.pmax holds the max over all ranks by now)
template <typename R>
__global__ void k_ctrl_step(Params<R> p) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && !p.ctrl[0].done) controller_update<R>(p, 0, from_bits<R>(p.ctrl[0].pmax));
}

// Barrier between the GPUs that share one long waveform (fused exchange): every rank posts the epoch into its slot of
// every peer's flag array (peer memory, system-scope release) and waits until all peers have posted theirs.  The
// stores of the preceding kernel on this stream -- including its stores into peer memory -- are ordered before the post.
static __global__ void k_xbar(unsigned int* const* peer_flags, unsigned int* my_flags, int me, int ranks, unsigned int epoch) {
    const int r = threadIdx.x;
    if (r < ranks) {
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flags[r] + me), "r"(epoch) : "memory");
        unsigned int seen;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(my_flags + r) : "memory");
        } while ((int)(seen - epoch) < 0);
    }
}

// Ask for SSFM_THREADS_PER_SM resident threads per SM (512 -> at most 128 registers per thread):
// several small CTAs per SM so that one CTA's global loads overlap another's transform.
#ifndef SSFM_THREADS_PER_SM
#define SSFM_THREADS_PER_SM (SSFM_E64 == 16 ? 512 : 768)
#endif
#ifndef SSFM_THREADS_PER_SM_F32
#define SSFM_THREADS_PER_SM_F32 768
#endif
__host__ __device__ constexpr int min_ctas_for(int threads, int per_sm) {
    return threads >= per_sm ? 1 : (per_sm / threads > 16 ? 16 : per_sm / threads);
}
template <typename R> __host__ __device__ constexpr int min_ctas(int threads) {
    return min_ctas_for(threads, sizeof(R) == 4 ? SSFM_THREADS_PER_SM_F32 : SSFM_THREADS_PER_SM);
}

// cp.async (LDGSTS): global -> shared without staging registers; used to prefetch the Kerr-phase
// stash of a tile while the inverse column transforms run.
template <int BYTES>
__device__ __forceinline__ void cp_async(void* smem_dst, const void* gmem_src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;\n" ::"r"(d), "l"(gmem_src), "n"(BYTES));
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

template <typename C>
__device__ __forceinline__ void load_tables(C* dst, const C* __restrict__ src, int count) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
}

// Four-step twiddles by recurrence: the seeds (two table look-ups and a product each) can be fetched before a barrier wait,
// the sixteen products after the field has arrived.
template <typename R> struct FsSeeds { typename cx_of<R>::type wa, wb, rho2; };
template <typename R, int E, int M>
__device__ __forceinline__ FsSeeds<R> fourstep_seeds(const Params<R>& p, int n2, int t) {
    typedef typename cx_of<R>::type C;
    const unsigned mask = (1u << p.lo_bits) - 1u;
    const unsigned i0 = (unsigned)n2 * (unsigned)t, i1 = (unsigned)n2 * (unsigned)(M / E);   // < N
    FsSeeds<R> s;
    s.wa = cmul(__ldg(p.tw_lo + (i0 & mask)), __ldg(p.tw_hi + (i0 >> p.lo_bits)));
    const C rho = cmul(__ldg(p.tw_lo + (i1 & mask)), __ldg(p.tw_hi + (i1 >> p.lo_bits)));
    s.rho2 = cmul(rho, rho);
    s.wb = cmul(s.wa, rho);
    return s;
}
template <bool CONJ, typename R, int E>
__device__ __forceinline__ void apply_fourstep_chain(typename cx_of<R>::type (&v)[E], FsSeeds<R> s) {
#pragma unroll
    for (int q = 0; q < E; q += 2) {
        v[q] = CONJ ? cmulc(v[q], s.wa) : cmul(v[q], s.wa);
        v[q + 1] = CONJ ? cmulc(v[q + 1], s.wb) : cmul(v[q + 1], s.wb);
        if (q + 2 < E) { s.wa = cmul(s.wa, s.rho2); s.wb = cmul(s.wb, s.rho2); }
    }
}

// v[q] *= W_N^{n2*k1} (or its conjugate), k1 = t + q*M/E.  The table choice is made ONCE outside the
// unrolled loop (a per-element branch doubles the code of every iteration and costs ~10 %).
template <bool CONJ, typename R, int E, int M>
__device__ __forceinline__ void apply_fourstep(const Params<R>& p, typename cx_of<R>::type (&v)[E], int n2, int t) {
    typedef typename cx_of<R>::type C;
    if (p.tw_chain) {                                    // W_N^{n2 (t + q M/E)} = W_N^{n2 t} (W_N^{n2 M/E})^q, even and odd q as two chains
        apply_fourstep_chain<CONJ, R, E>(v, fourstep_seeds<R, E, M>(p, n2, t));
    } else if (p.tw_full) {                              // full table [k1][n2]: one coalesced L2 load per point
        const C* __restrict__ base = p.tw_full + (size_t)t * p.n2 + n2;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const C w = __ldg(base + (size_t)q * (M / E) * p.n2);
            v[q] = CONJ ? cmulc(v[q], w) : cmul(v[q], w);
        }
    } else {                                             // two sqrt(N) tables: W_N^lo * W_N^(hi*2^lo_bits)
        const unsigned mask = (1u << p.lo_bits) - 1u;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const unsigned idx = (unsigned)n2 * (unsigned)(t + q * (M / E));   // < N
            const C w = cmul(__ldg(p.tw_lo + (idx & mask)), __ldg(p.tw_hi + (idx >> p.lo_bits)));
            v[q] = CONJ ? cmulc(v[q], w) : cmul(v[q], w);
        }
    }
}

// destination of element (r, col) of the local matrix of matrix-row pitch p.n2 (see Params::peer)
template <typename R>
__device__ __forceinline__ typename cx_of<R>::type* exchange_dst(const Params<R>& p, typename cx_of<R>::type* rowp, int r, int col, int bp) {
    if (p.peer_mode == 1)
        return p.peer[r >> p.peer_shift] + (size_t)(r & ((1 << p.peer_shift) - 1)) * p.peer_pitch + p.peer_base + col;
    if (p.peer_mode == 2) {
        const int nb = r * p.n2 + col;
        return p.peer[nb >> p.peer_shift] + (size_t)(p.peer_base + bp) * p.peer_pitch + (nb & ((1 << p.peer_shift) - 1));
    }
    return rowp + (size_t)r * p.n2 + col;
}

// ---------------------------------------------------------------------------------------------
// initial max |A|^2 per waveform (first step size, devices.py:1155-1156)
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void k_power_max(Params<R> p, int blocks_per_wf) {
    typedef typename cx_of<R>::type C;
    __shared__ unsigned long long red[32];
    const int b = blockIdx.x / blocks_per_wf, part = blockIdx.x % blocks_per_wf;
    const size_t len = (size_t)p.n_pol * p.n;
    const C* a = p.field + (size_t)b * len;
    R m = 0;
    bool nan = false;
    for (size_t i = (size_t)part * blockDim.x + threadIdx.x; i < len; i += (size_t)blocks_per_wf * blockDim.x) {
        C v = a[i];
        R pw = v.x * v.x + v.y * v.y;
        nan |= (pw != pw);
        m = pw > m ? pw : m;
    }
    if (nan) m = pw_nan<R>();
    m = block_max_bits<R>(m, red);
    if (threadIdx.x == 0) atomicMax(&p.ctrl[b].pmax, ord_bits(m));
}

// first step size and controller reset (devices.py:1155-1161)
template <typename R>
__global__ void k_ctrl_init(Params<R> p, int fixed, R h_fixed, int single_step) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.batch) return;
    Ctrl& c = p.ctrl[b];
    R h;
    if (fixed) h = h_fixed;
    else if (single_step) h = p.length;                         // no dispersion or no Kerr effect
    else h = p.phi_max / mul_rn(p.abs_gamma, from_bits<R>(c.pmax));
    h = (p.length < h) ? p.length : h;                          // python min(h_, length)
    c.z = 0.0; c.h = (double)h; c.pmax = 0ull; c.steps = 0; c.arrived = 0u;
    c.done = !((R)0 < p.length) || p.max_steps <= 0;
    if (c.done) atomicSub(p.active, 1);
}

// ---------------------------------------------------------------------------------------------
// column pass, forward
// ---------------------------------------------------------------------------------------------
template <typename R, int M, int T>
__global__ void __launch_bounds__(T * (M / points_per_thread<R>::value), min_ctas<R>(T * (M / points_per_thread<R>::value))) k_col_fwd(Params<R> p) {
    typedef typename cx_of<R>::type C;
    constexpr int E = points_per_thread<R>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);                   // [M][T] exchange tile
    C* tw = sm + M * T;                                       // pass tables

    const int tiles = p.n2 / T;
    const int tile = blockIdx.x % tiles, row = blockIdx.x / tiles;   // row = b*P + pol
    const int b = p.inner ? 0 : row / p.n_pol;
    const Ctrl ctl = p.ctrl[b];
    if (ctl.done) return;

    const int c = threadIdx.x % T, t = threadIdx.x / T;
    const int n2 = tile * T + c;
    load_tables(tw, p.tw_col, fft_plan<M, E>::table_size + SC_N);
    const C* sct = tw + fft_plan<M, E>::table_size;            // sincos table rides behind the pass tables

    C* rowp = p.field + (size_t)row * p.n;
    R* strow = p.stash + (size_t)row * p.n;
    C v[E];
#pragma unroll
    for (int q = 0; q < E; ++q) v[q] = rowp[(size_t)(t + q * (M / E)) * p.n2 + n2];
    __syncthreads();                                            // tables (incl. the sincos table) are in place

    if (p.has_nl) {
        const R hh = (R)ctl.h / (R)2;                           // h_/2
        auto kerr = [&](auto small) {
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const R pw = v[q].x * v[q].x + v[q].y * v[q].y; // |A|^2
                const R ph = mul_rn(hh, mul_rn(p.gamma, pw));   // (h_/2) * (gamma |A|^2)
                strow[(size_t)(t + q * (M / E)) * p.n2 + n2] = ph;
                R s, co; kerr_sincos<decltype(small)::value>(ph, sct, &s, &co);
                v[q] = cmul(v[q], mk<R>(co, s));
            }
        };
        if (sizeof(R) == 8 && p.small_phase) kerr(std::true_type{}); else kerr(std::false_type{});   // fp32: sincosf is cheap
    }
    fft_passes<R, M, -1, ColExchange<T>, E>::run(v, sm + c, tw, t);
    apply_fourstep<false, R, E, M>(p, v, n2 + p.n2_off, t);
#pragma unroll
    for (int q = 0; q < E; ++q) *exchange_dst<R>(p, rowp, t + q * (M / E), n2, row) = v[q];
}

// ---------------------------------------------------------------------------------------------
// row pass: forward transform, linear operator, inverse transform
// ---------------------------------------------------------------------------------------------
template <typename R, int M, int G>
__global__ void __launch_bounds__(G * (M / points_per_thread<R>::value), min_ctas<R>(G * (M / points_per_thread<R>::value))) k_row(Params<R> p) {
    typedef typename cx_of<R>::type C;
    constexpr int E = points_per_thread<R>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PM = RowExchange<M, E>::size;
    C* sm = reinterpret_cast<C*>(smem_raw);                   // [G][PM] private exchange buffers
    C* tw = sm + G * PM;

    const int g = threadIdx.x / (M / E), t = threadIdx.x % (M / E);
    const long long grow = (long long)blockIdx.x * G + g;       // global row index over [B*P][N1]
    const int bp = (int)(grow / p.n1), k1 = (int)(grow % p.n1);
    const int b = p.inner ? 0 : bp / p.n_pol;
    const Ctrl ctl = p.ctrl[b];
    if (ctl.done) return;                                       // G divides N1: uniform per block

    load_tables(tw, p.tw_row, fft_plan<M, E>::table_size + SC_N);
    const C* sct = tw + fft_plan<M, E>::table_size;
    C* base = p.field + (size_t)bp * p.n + (size_t)k1 * p.n2;
    C v[E];
#pragma unroll
    for (int q = 0; q < E; ++q) v[q] = base[t + q * (M / E)];
    __syncthreads();
    fft_passes<R, M, -1, RowExchange<M, E>, E>::run(v, sm + g * PM, tw, t);
    if (p.fwd_only) {
#pragma unroll
        for (int q = 0; q < E; ++q) base[t + q * (M / E)] = v[q];
        return;
    }

    if (p.xfer) {   // arbitrary transfer function (zero-phase filters: |H|^2; DM: exp(j w^2 D/2))
        const C* __restrict__ hrow = p.xfer + (size_t)k1 * p.n2;
#pragma unroll
        for (int q = 0; q < E; ++q) v[q] = cmul(v[q], __ldg(hrow + t + q * (M / E)));
    } else {        // exp(D~ h): real part -alpha/2*h (attenuation), imaginary part (b2/2 w^2 + b3/6 w^3) h
        const R h = (R)ctl.h;                                   // (the attenuation exp(-alpha/2 h) of D~ is a per-step
        const int half = p.n_glob >> 1;                         //  scalar: it is folded into the 1/N of the column pass)
        const int kbase = p.inner ? p.bin_off + bp : 0;         // long waveforms: this row is outer bin ka = bin_off + row
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const int k2 = t + q * (M / E);
            int k = kbase + p.bin_mul * (k1 + p.n1 * k2);       // transposed-order bin index
            k = (k < half) ? k : k - p.n_glob;                  // fftfreq ordering
            const R w = (R)((double)k * p.wscale);              // rad/ps, = fftfreq*2*pi*1e-12 (see Params::wscale)
            const R dim = add_rn(mul_rn(p.c2, mul_rn(w, w)), mul_rn(p.c3, cube_r(w)));
            const R ph = mul_rn(dim, h);
            R s, co; sincos_r(ph, sct, &s, &co);
            v[q] = cmul(v[q], mk<R>(co, s));
        }
    }
    fft_passes<R, M, +1, RowExchange<M, E>, E>::run(v, sm + g * PM, tw, t);
#pragma unroll
    for (int q = 0; q < E; ++q) base[t + q * (M / E)] = v[q];
}

// ---------------------------------------------------------------------------------------------
// column pass, inverse, second Kerr half step, power max, controller
// ---------------------------------------------------------------------------------------------
template <typename R, int M, int T>
__global__ void __launch_bounds__(T * (M / points_per_thread<R>::value), min_ctas<R>(T * (M / points_per_thread<R>::value))) k_col_inv(Params<R> p) {
    typedef typename cx_of<R>::type C;
    constexpr int E = points_per_thread<R>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long red[32];
    C* sm = reinterpret_cast<C*>(smem_raw);
    C* tw = sm + M * T;
    R* st_sm = reinterpret_cast<R*>(tw + fft_plan<M, E>::table_size + SC_N);   // [E][threads] stash prefetch

    const int tiles = p.n2 / T;
    const int tile = blockIdx.x % tiles, row = blockIdx.x / tiles;
    const int b = p.inner ? 0 : row / p.n_pol;
    if (p.ctrl[b].done) return;
    const R sc = p.inv_n * exp_r(mul_rn(p.att_half, (R)p.ctrl[b].h));   // 1/N and exp(-alpha/2 h) (real part of D~ h)

    const int c = threadIdx.x % T, t = threadIdx.x / T;
    const int n2 = tile * T + c;
    load_tables(tw, p.tw_col, fft_plan<M, E>::table_size + SC_N);
    const C* sct = tw + fft_plan<M, E>::table_size;            // sincos table rides behind the pass tables

    C* __restrict__ rowp = p.field + (size_t)row * p.n;
    const R* __restrict__ strow = p.stash + (size_t)row * p.n;
    if (p.has_nl) {
#pragma unroll
        for (int q = 0; q < E; ++q)
            cp_async<sizeof(R)>(st_sm + q * (T * (M / E)) + threadIdx.x, strow + (size_t)(t + q * (M / E)) * p.n2 + n2);
        cp_async_commit();
    }
    C v[E];
#pragma unroll
    for (int q = 0; q < E; ++q) v[q] = rowp[(size_t)(t + q * (M / E)) * p.n2 + n2];
    apply_fourstep<true, R, E, M>(p, v, n2 + p.n2_off, t);
    __syncthreads();
    fft_passes<R, M, +1, ColExchange<T>, E>::run(v, sm + c, tw, t);

    R pm = 0;
    bool nan = false;
    cp_async_wait_all();
#pragma unroll
    for (int q = 0; q < E; ++q) {
        const size_t off = (size_t)(t + q * (M / E)) * p.n2 + n2;
        C a = v[q];
        a.x *= sc; a.y *= sc;                                   // numpy ifft scaling (exact: N = 2^n) x attenuation of the step
        if (p.has_nl) {
            R s, co; sincos_r(st_sm[q * (T * (M / E)) + threadIdx.x], sct, &s, &co);
            a = cmul(a, mk<R>(co, s));
        }
        const R pw = a.x * a.x + a.y * a.y;
        nan |= (pw != pw);
        pm = pw > pm ? pw : pm;
        if (p.peer_mode) *exchange_dst<R>(p, rowp, t + q * (M / E), n2, row) = a;
        else rowp[off] = a;
    }
    if (nan) pm = pw_nan<R>();
    if (p.inner) return;                                        // inner transform of a long waveform: no controller
    pm = block_max_bits<R>(pm, red);

    if (threadIdx.x == 0) {
        atomicMax(&p.ctrl[b].pmax, ord_bits(pm));
        if (p.defer_ctrl) return;
        __threadfence();
        const unsigned total = (unsigned)(tiles * p.n_pol);
        const unsigned prev = atomicAdd(&p.ctrl[b].arrived, 1u);
        if (prev + 1u == total) {                               // last tile of this waveform
            __threadfence();
            const R all = from_bits<R>(atomicMax(&p.ctrl[b].pmax, 0ull));
            controller_update<R>(p, b, all);
        }
    }
}


// ---------------------------------------------------------------------------------------------
// fused column pass: end of step s and start of step s+1 in one visit of the tile
//   conj twiddle, inverse column transforms, 1/N, max |A|^2  ->  step-size controller
//   ->  rotation by the stashed phase of step s PLUS the first Kerr half step of step s+1
//       (one sincos), stash of the new phase, forward column transforms, twiddle.
// Field traffic per step drops from 3R+3W to 2R+2W (the ideal of SURVEY.md §8(d)).
//
// The controller needs the max |A|^2 over the WHOLE waveform (devices.py:1194) before the next
// Kerr half step can be applied, i.e. a barrier over the tiles of one waveform in mid-kernel:
//   SYNC_FIXED   fixed step size: h' does not depend on the field, no barrier at all; every tile
//                derives the new state locally, tile 0 commits it once all tiles have read the old one.
//   SYNC_CLUSTER the tiles of a waveform form one thread-block cluster (<= 16 CTAs): maxima are
//                exchanged through distributed shared memory, one hardware cluster barrier.
//   SYNC_LL      (default) any number of tiles: every tile publishes its maximum as self-validating
//                64-bit words {32 bits of the value, step tag} (the NCCL "LL" idea: the flag travels with the
//                data, so no fence and no atomic is needed), every tile polls the words of its waveform and
//                runs the controller itself.  One store + one polling round trip instead of
//                atomicMax / fence / atomicAdd / spin / fence / reload.
//   SYNC_GLOBAL  any number of tiles: atomics in global memory and a spin wait.  CTAs take a ticket
//                when they start and the ticket -- not blockIdx -- selects the tile, so the tiles of a
//                waveform start in order and the lowest unfinished waveform always has all its tiles
//                resident: no deadlock as long as tiles-per-waveform <= resident CTAs (host check).
//                (SYNC_LL uses the same tickets.)
// ---------------------------------------------------------------------------------------------
enum { SYNC_FIXED = 0, SYNC_CLUSTER = 1, SYNC_GLOBAL = 2, SYNC_LL = 3 };
__host__ __device__ constexpr bool col_mid_tables_in_smem(int M) { return M < 1024; }

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_arrive_release() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait_acquire() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void st_cluster_u64(void* local_smem, unsigned rank, unsigned long long v) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(local_smem);
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(ra), "l"(v) : "memory");
}

template <typename R, int M, int T, int SYNC>
__global__ void __launch_bounds__(T * (M / points_per_thread<R>::value), min_ctas<R>(T * (M / points_per_thread<R>::value))) k_col_mid(Params<R> p) {
    typedef typename cx_of<R>::type C;
    constexpr int E = points_per_thread<R>::value;
    constexpr int NT = T * (M / E);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long red[32];
    __shared__ unsigned long long cl_max[16];
    __shared__ unsigned int s_ticket;
    __shared__ int s_done, s_steps;
    __shared__ double s_hnext, s_z;
    // Pass tables live in shared memory for M < 1024; for longer transforms they are read through L1 from
    // global memory instead, which keeps the CTA under half of the SM's shared memory (two resident CTAs,
    // so that the 256 tiles of a single 2^20-sample waveform fit on the chip for the in-kernel barrier).
    constexpr bool kTabSmem = col_mid_tables_in_smem(M);
    constexpr int kTabCount = (kTabSmem ? fft_plan<M, E>::table_size : 0) + SC_N;
    C* sm = reinterpret_cast<C*>(smem_raw);
    C* tws = sm + M * T;                                                      // [pass tables][sincos table]
    R* st_sm = reinterpret_cast<R*>(tws + kTabCount);                          // [E][threads] stash prefetch
    const C* tw = kTabSmem ? tws : p.tw_col;
    const C* sct = tws + (kTabSmem ? fft_plan<M, E>::table_size : 0);

    // Thread 0 takes the ticket (SYNC_GLOBAL / SYNC_LL) and reads the controller state ONCE for the whole
    // CTA (only then does it report "state read" in SYNC_FIXED), so no thread can see a state committed by
    // a faster tile of the same waveform.  The state loads overlap the tile loads issued below; the
    // values are handed over at the barrier in front of the transforms.
    constexpr bool kTicket = (SYNC == SYNC_GLOBAL || SYNC == SYNC_LL);
    const int tiles = p.n2 / T;
    const unsigned total = (unsigned)(tiles * p.n_pol);
    int blk = blockIdx.x;
    if (kTicket) {
        if (threadIdx.x == 0) {
            const unsigned int tk = atomicAdd(p.ticket, 1u);
            if (tk == gridDim.x - 1) *p.ticket = 0u;          // last CTA to start re-arms the counter
            s_ticket = tk;
        }
        __syncthreads();
        blk = (int)s_ticket;
    }
    const int tile = blk % tiles, row = blk / tiles;
    const int b = row / p.n_pol;
    Ctrl* ctl = p.ctrl + b;
    if (p.l2_ahead > 0 && blk + p.l2_ahead < (int)gridDim.x) {   // warm L2 for a CTA that starts one lifetime later
        const int pb = blk + p.l2_ahead;
        const int ptile = pb % tiles, prow = pb / tiles;
        if (!p.ctrl[prow / p.n_pol].done) {
            const char* fa = reinterpret_cast<const char*>(p.field + (size_t)prow * p.n + (size_t)ptile * T);
            const char* sa = reinterpret_cast<const char*>(p.stash + (size_t)prow * p.n + (size_t)ptile * T);
            constexpr int FL = (T * (int)sizeof(C) + 127) / 128, SL = (T * (int)sizeof(R) + 127) / 128;   // 128-B lines per row segment
            for (int i = threadIdx.x; i < M * FL; i += NT)
                prefetch_l2(fa + (size_t)(i / FL) * p.n2 * sizeof(C) + (i % FL) * 128);
            if (p.has_nl)
                for (int i = threadIdx.x; i < M * SL; i += NT)
                    prefetch_l2(sa + (size_t)(i / SL) * p.n2 * sizeof(R) + (i % SL) * 128);
        }
    }
    if (threadIdx.x == 0) {
        s_done = *reinterpret_cast<volatile int*>(&ctl->done);
        s_steps = *reinterpret_cast<volatile int*>(&ctl->steps);
        s_z = *reinterpret_cast<volatile double*>(&ctl->z);
        s_hnext = *reinterpret_cast<volatile double*>(&ctl->h);
        if (SYNC == SYNC_FIXED) { __threadfence(); atomicAdd(&ctl->arrived, 1u); }
    }

    const int c = threadIdx.x % T, t = threadIdx.x / T;
    const int n2 = tile * T + c;
    load_tables(tws, p.tw_col + (kTabSmem ? 0 : fft_plan<M, E>::table_size), kTabCount);

    C* __restrict__ rowp = p.field + (size_t)row * p.n;
    R* __restrict__ strow = p.stash + (size_t)row * p.n;
    if (p.has_nl) {
#pragma unroll
        for (int q = 0; q < E; ++q)
            cp_async<sizeof(R)>(st_sm + q * NT + threadIdx.x, strow + (size_t)(t + q * (M / E)) * p.n2 + n2);
        cp_async_commit();
    }
    C v[E];
#pragma unroll
    for (int q = 0; q < E; ++q) v[q] = rowp[(size_t)(t + q * (M / E)) * p.n2 + n2];
    apply_fourstep<true, R, E, M>(p, v, n2 + p.n2_off, t);
    __syncthreads();
    const int step_before = s_steps;
    const R z_before = (R)s_z, h_before = (R)s_hnext;
    if (s_done) {                                              // uniform over the waveform (and its cluster)
        cp_async_wait_all();
        if (SYNC == SYNC_FIXED && tile == 0 && (row % p.n_pol) == 0 && threadIdx.x == 0) {
            while (*reinterpret_cast<volatile unsigned int*>(&ctl->arrived) < total) __nanosleep(32);
            ctl->arrived = 0u;
        }
        return;
    }
    fft_passes<R, M, +1, ColExchange<T>, E>::run(v, sm + c, tw, t);

    const R sc = p.inv_n * exp_r(mul_rn(p.att_half, h_before));   // 1/N (exact: N = 2^n) and exp(-alpha/2 h) (real part of D~ h)
#pragma unroll
    for (int q = 0; q < E; ++q) { v[q].x *= sc; v[q].y *= sc; }

    CtrlNext<R> nx;
    if (SYNC == SYNC_FIXED) {
        nx = controller_next<R>(p, z_before, h_before, step_before, (R)0);
    } else {
        R pm = 0;
        bool nan = false;
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const R pw = v[q].x * v[q].x + v[q].y * v[q].y;    // the Kerr rotations do not change |A|
            nan |= (pw != pw);
            pm = pw > pm ? pw : pm;
        }
        if (nan) pm = pw_nan<R>();
        if (SYNC == SYNC_LL) {
            // block max -> warp 0 publishes it and polls the words of the whole waveform
            constexpr int NW = sizeof(R) / 4;                  // 64-bit words per slot: {value bits 63..32 | tag}, {bits 31..0 | tag}
            unsigned long long bits = ord_bits(pm);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, bits, o); bits = x > bits ? x : bits; }
            const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
            if (lane == 0) red[warp] = bits;
            __syncthreads();
            if (warp == 0) {
                constexpr int NWARPS = (NT + 31) / 32;
                bits = lane < NWARPS ? red[lane] : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, bits, o); bits = x > bits ? x : bits; }
                const unsigned long long tag = (unsigned long long)(unsigned)(step_before + 1);
                volatile unsigned long long* wf = p.slots + (size_t)b * total * 2;
                const int me = (row % p.n_pol) * tiles + tile;
                if (lane < NW) {
                    const unsigned long long part = (NW == 1) ? (bits & 0xffffffffull) : (lane == 0 ? (bits >> 32) : (bits & 0xffffffffull));
                    wf[me * 2 + lane] = (part << 32) | tag;
                }
                unsigned long long best = 0ull;
                const int nwords = (int)total * NW;
                for (int base = 0; base < nwords; base += 32) {
                    const int idx = base + lane;
                    const bool have = idx < nwords;
                    unsigned long long w = tag;
                    for (;;) {
                        if (have) w = wf[(idx / NW) * 2 + (idx % NW)];
                        if (__all_sync(0xffffffffu, !have || (w & 0xffffffffull) == tag)) break;
                        __nanosleep(20);
                    }
                    unsigned long long val = have ? (w >> 32) : 0ull;
                    if (NW == 2) {
                        const unsigned long long other = __shfl_xor_sync(0xffffffffu, val, 1);
                        val = (lane & 1) ? 0ull : ((val << 32) | other);
                    }
                    best = val > best ? val : best;
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, best, o); best = x > best ? x : best; }
                if (lane == 0) red[0] = best;
            }
            __syncthreads();
            nx = controller_next<R>(p, z_before, h_before, step_before, from_bits<R>(red[0]));
            if (tile == 0 && (row % p.n_pol) == 0 && threadIdx.x == 0) controller_commit<R>(p, b, h_before, step_before, nx);
        } else {
        pm = block_max_bits<R>(pm, red);
        if (SYNC == SYNC_CLUSTER) {
            const unsigned rank = cluster_ctarank();
            if (threadIdx.x < total) st_cluster_u64(&cl_max[rank], threadIdx.x, ord_bits(pm));
            cluster_arrive_release();
            cluster_wait_acquire();
            unsigned long long m = 0ull;
            for (unsigned i = 0; i < total; ++i) m = cl_max[i] > m ? cl_max[i] : m;
            nx = controller_next<R>(p, z_before, h_before, step_before, from_bits<R>(m));
            if (rank == 0 && threadIdx.x == 0) controller_commit<R>(p, b, h_before, step_before, nx);
        } else {
            if (threadIdx.x == 0) {
                atomicMax(&ctl->pmax, ord_bits(pm));
                __threadfence();
                const unsigned prev = atomicAdd(&ctl->arrived, 1u);
                if (prev + 1u == total) {
                    __threadfence();
                    const R all = from_bits<R>(atomicMax(&ctl->pmax, 0ull));
                    const CtrlNext<R> n0 = controller_next<R>(p, z_before, h_before, step_before, all);
                    controller_commit<R>(p, b, h_before, step_before, n0);
                    s_done = n0.done; s_hnext = (double)n0.h;
                } else {
                    while (*reinterpret_cast<volatile int*>(&ctl->steps) == step_before) __nanosleep(32);
                    __threadfence();
                    s_done = *reinterpret_cast<volatile int*>(&ctl->done);
                    s_hnext = *reinterpret_cast<volatile double*>(&ctl->h);
                }
            }
            __syncthreads();
            nx.done = s_done; nx.h = (R)s_hnext; nx.z = 0;
        }
        }
    }
    cp_async_wait_all();

    if (nx.done) {                                              // last step of this waveform: time domain out
#pragma unroll
        for (int q = 0; q < E; ++q) {
            const size_t off = (size_t)(t + q * (M / E)) * p.n2 + n2;
            if (p.has_nl) {
                R s, co; sincos_r(st_sm[q * NT + threadIdx.x], sct, &s, &co);
                v[q] = cmul(v[q], mk<R>(co, s));
            }
            rowp[off] = v[q];
        }
    } else {
        if (p.has_nl) {
            const R hh = nx.h / (R)2;                           // h_/2 of the NEXT step
            auto kerr = [&](auto small) {
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const size_t off = (size_t)(t + q * (M / E)) * p.n2 + n2;
                    const R pw = v[q].x * v[q].x + v[q].y * v[q].y;
                    const R ph = mul_rn(hh, mul_rn(p.gamma, pw));   // first half step of the next step
                    const R tot = st_sm[q * NT + threadIdx.x] + ph; // + second half step of this one
                    strow[off] = ph;
                    R s, co; kerr_sincos<decltype(small)::value>(tot, sct, &s, &co);
                    v[q] = cmul(v[q], mk<R>(co, s));
                }
            };
            if (sizeof(R) == 8 && p.small_phase) kerr(std::true_type{}); else kerr(std::false_type{});
        }
        fft_passes<R, M, -1, ColExchange<T>, E>::run(v, sm + c, tw, t);
        apply_fourstep<false, R, E, M>(p, v, n2 + p.n2_off, t);
#pragma unroll
        for (int q = 0; q < E; ++q) *exchange_dst<R>(p, rowp, t + q * (M / E), n2, row) = v[q];
    }
    if (SYNC == SYNC_FIXED && tile == 0 && (row % p.n_pol) == 0 && threadIdx.x == 0) {
        // every tile of the waveform has read the old state by now (or will within microseconds)
        while (*reinterpret_cast<volatile unsigned int*>(&ctl->arrived) < total) __nanosleep(32);
        controller_commit<R>(p, b, h_before, step_before, nx);
    }
}


// ---------------------------------------------------------------------------------------------
// Arbitrary (non power-of-two) lengths: the reference accepts any N (numpy.fft).  The N-point transforms of
// devices.py:1178-1180 are evaluated as chirp-z (Bluestein) convolutions of length M = 2^ceil(log2(2N-1)) with the
// power-of-two kernels above:   DFT_N(x)[k] = w[k] * sum_n (x[n] w[n]) conj(w)[k-n],   w[n] = exp(-j pi n^2 / N).
// Per split step:  k_bs_open  (first Kerr half step, * w, zero padding)  ->  FFT_M, * FFT_M(conj w), IFFT_M
//   ->  k_bs_mid  (* exp(j imag(D~) h); the post-chirp of the forward and the pre-chirp of the inverse transform cancel)
//   ->  FFT_M, * FFT_M(w), IFFT_M  ->  k_bs_close (* conj(w)/N, attenuation, second Kerr half step, max |A|^2)
//   ->  k_ctrl_steps (controller of every waveform).  About ten passes over 2-4 N samples: a completeness path.
// ---------------------------------------------------------------------------------------------
template <typename R>
__global__ void k_bs_chirp(typename cx_of<R>::type* wtab, int n) {      // w[i] = exp(-j pi i^2 / n), i^2 reduced mod 2n exactly
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long long r = ((long long)i * i) % (2ll * n);
    double s, c;
    sincospi((double)r / (double)n, &s, &c);
    wtab[i] = mk<R>((R)c, (R)(-s));
}
template <typename R>
__global__ void k_bs_kernel_row(typename cx_of<R>::type* row, const typename cx_of<R>::type* wtab, int n, int m, int conj_w) {
    // b[i] = conj(w)[i] (or w[i]) at circular index i and m - i, zero elsewhere
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    typename cx_of<R>::type v = mk<R>((R)0, (R)0);
    int d = -1;
    if (i < n) d = i; else if (m - i < n) d = m - i;
    if (d >= 0) { v = wtab[d]; if (conj_w) v.y = -v.y; }
    row[i] = v;
}
template <typename R>
__global__ void k_bs_open(Params<R> p, typename cx_of<R>::type* wb, const typename cx_of<R>::type* __restrict__ wtab, int m,
                          const typename cx_of<R>::type* __restrict__ sct) {
    typedef typename cx_of<R>::type C;
    const int row = blockIdx.y, b = row / p.n_pol;
    const Ctrl ctl = p.ctrl[b];
    if (ctl.done) return;
    const R hh = (R)ctl.h / (R)2;
    const C* __restrict__ f = p.field + (size_t)row * p.n;
    R* __restrict__ st = p.stash + (size_t)row * p.n;
    C* __restrict__ w = wb + (size_t)row * m;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        C a = mk<R>((R)0, (R)0);
        if (i < p.n) {
            a = f[i];
            if (p.has_nl) {
                const R pw = a.x * a.x + a.y * a.y;
                const R ph = mul_rn(hh, mul_rn(p.gamma, pw));
                st[i] = ph;
                R s, c; sincos_r(ph, sct, &s, &c);
                a = cmul(a, mk<R>(c, s));
            }
            a = cmul(a, wtab[i]);
        }
        w[i] = a;
    }
}
template <typename R>
__global__ void k_bs_mid(Params<R> p, typename cx_of<R>::type* wb, int m, const typename cx_of<R>::type* __restrict__ sct) {
    typedef typename cx_of<R>::type C;
    const int row = blockIdx.y, b = row / p.n_pol;
    const Ctrl ctl = p.ctrl[b];
    if (ctl.done) return;
    const R h = (R)ctl.h;
    C* __restrict__ w = wb + (size_t)row * m;
    const int pos = (p.n + 1) >> 1;                             // numpy.fft.fftfreq: bins 0 .. ceil(n/2)-1 are non-negative
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) {
        C a = mk<R>((R)0, (R)0);
        if (i < p.n) {
            if (p.xfer) {                                       // arbitrary transfer function H[k], numpy bin order (DM, FBG)
                a = cmul(w[i], p.xfer[i]);
            } else {
                const int k = (i < pos) ? i : i - p.n;
                const R wk = (R)((double)k * p.wscale);
                const R dim = add_rn(mul_rn(p.c2, mul_rn(wk, wk)), mul_rn(p.c3, cube_r(wk)));
                R s, c; sincos_r(mul_rn(dim, h), sct, &s, &c);
                a = cmul(w[i], mk<R>(c, s));
            }
        }
        w[i] = a;
    }
}
template <typename R>
__global__ void k_bs_close(Params<R> p, const typename cx_of<R>::type* wb, const typename cx_of<R>::type* __restrict__ wtab, int m,
                           const typename cx_of<R>::type* __restrict__ sct) {
    typedef typename cx_of<R>::type C;
    __shared__ unsigned long long red[32];
    const int row = blockIdx.y, b = row / p.n_pol;
    const Ctrl ctl = p.ctrl[b];
    if (ctl.done) return;
    const R sc = p.inv_n * exp_r(mul_rn(p.att_half, (R)ctl.h));
    C* __restrict__ f = p.field + (size_t)row * p.n;
    const R* __restrict__ st = p.stash + (size_t)row * p.n;
    const C* __restrict__ w = wb + (size_t)row * m;
    R pm = 0;
    bool nan = false;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.n; i += gridDim.x * blockDim.x) {
        C a = cmulc(w[i], wtab[i]);
        a.x *= sc; a.y *= sc;
        if (p.has_nl) {
            R s, c; sincos_r(st[i], sct, &s, &c);
            a = cmul(a, mk<R>(c, s));
        }
        const R pw = a.x * a.x + a.y * a.y;
        nan |= (pw != pw);
        pm = pw > pm ? pw : pm;
        f[i] = a;
    }
    if (nan) pm = pw_nan<R>();
    pm = block_max_bits<R>(pm, red);
    if (threadIdx.x == 0) atomicMax(&p.ctrl[b].pmax, ord_bits(pm));
}
template <typename R>
__global__ void k_ctrl_steps(Params<R> p) {                    // deferred controller step of every unfinished waveform
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < p.batch && !p.ctrl[b].done) controller_update<R>(p, b, from_bits<R>(p.ctrl[b].pmax));
}

}  // namespace ssfm
