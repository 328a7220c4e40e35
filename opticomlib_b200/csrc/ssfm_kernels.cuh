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
    R* stash;            // [B][P][N] Kerr phase of the current step
    Ctrl* ctrl;          // [B]
    int* active;         // number of waveforms with done == 0
    unsigned int* ticket;// start-order ticket counter of the fused column kernel
    double* hlog;        // [B][hlog_cap] step sizes actually taken (may be null)
    const C* tw_col;     // pass tables of the N1-point transform
    const C* tw_row;     // pass tables of the N2-point transform
    const C* tw_lo;      // W_N^i,          i < 2^lo_bits
    const C* tw_hi;      // W_N^(i*2^lo_bits)
    int lo_bits;
    int n, n1, n2, log2_n2;
    int n_pol;
    int batch;           // waveforms in this launch
    int hlog_cap;
    int adaptive;        // 1: h = phi_max / max(|gamma| |A|^2) after every step
    int has_nl;          // gamma != 0
    int max_steps;       // safety stop
    R gamma, abs_gamma, phi_max, length;
    R att_half;          // -alpha_lin/2            (devices.py:1145)
    R c2;                // imag(1j/2*beta_2)       (devices.py:1145)
    R c3;                // imag(1j/6*beta_3)
    double fval;         // 1.0/(N*dt): numpy.fft.fftfreq scale (typing.py:1641)
    R inv_n;
};

// ---- rounding-exact scalar helpers (no FMA contraction where the reference has none) ----------
__device__ __forceinline__ float  mul_rn(float a, float b)   { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float  add_rn(float a, float b)   { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ void sincos_r(float x, float* s, float* c)    { sincosf(x, s, c); }
__device__ __forceinline__ void sincos_r(double x, double* s, double* c) { sincos(x, s, c); }
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

// Step-size controller, one thread per waveform per step (devices.py:1173, 1193-1196).
template <typename R>
__device__ __forceinline__ void controller_update(const Params<R>& p, int b, R pmax) {
    Ctrl& c = p.ctrl[b];
    R z = (R)c.z, h = (R)c.h;
    z = add_rn(z, h);                                   // z += h_
    const int s = c.steps;
    if (p.hlog && s < p.hlog_cap) p.hlog[(size_t)b * p.hlog_cap + s] = (double)h;
    R hn = h;
    if (p.adaptive) hn = p.phi_max / mul_rn(p.abs_gamma, pmax);   // phi_max / max(|gamma| |A|^2)
    const R rem = p.length - z;
    hn = (rem < hn) ? rem : hn;                         // python min(h_, length - z)
    const int done = !(z < p.length) || (s + 1 >= p.max_steps);
    c.z = (double)z; c.h = (double)hn; c.pmax = 0ull; c.arrived = 0u;
    c.done = done;
    if (done) atomicSub(p.active, 1);
    __threadfence();                                    // publish the new state, then the step count
    *reinterpret_cast<volatile int*>(&c.steps) = s + 1; // (waiters of k_col_mid spin on it)
}

// Two resident CTAs per SM (<= 128 registers) whenever a CTA has at most 256 threads: one CTA's
// global loads overlap the other's transform.
__host__ __device__ constexpr int min_ctas(int threads) { return threads <= 256 ? 2 : 1; }

template <typename C>
__device__ __forceinline__ void load_tables(C* dst, const C* __restrict__ src, int count) {
    for (int i = threadIdx.x; i < count; i += blockDim.x) dst[i] = src[i];
}

template <typename R>
__device__ __forceinline__ typename cx_of<R>::type fourstep_twiddle(const Params<R>& p, int n2, int k1) {
    const unsigned idx = (unsigned)n2 * (unsigned)k1;   // < N
    typename cx_of<R>::type lo = __ldg(p.tw_lo + (idx & ((1u << p.lo_bits) - 1u)));
    typename cx_of<R>::type hi = __ldg(p.tw_hi + (idx >> p.lo_bits));
    return cmul(lo, hi);
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
__global__ void __launch_bounds__(T * (M / 16), min_ctas(T * (M / 16))) k_col_fwd(Params<R> p) {
    typedef typename cx_of<R>::type C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    C* sm = reinterpret_cast<C*>(smem_raw);                   // [M][T] exchange tile
    C* tw = sm + M * T;                                       // pass tables

    const int tiles = p.n2 / T;
    const int tile = blockIdx.x % tiles, row = blockIdx.x / tiles;   // row = b*P + pol
    const int b = row / p.n_pol;
    const Ctrl ctl = p.ctrl[b];
    if (ctl.done) return;

    const int c = threadIdx.x % T, t = threadIdx.x / T;
    const int n2 = tile * T + c;
    load_tables(tw, p.tw_col, fft_plan<M>::table_size);

    C* rowp = p.field + (size_t)row * p.n;
    R* strow = p.stash + (size_t)row * p.n;
    C v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = rowp[(size_t)(t + q * (M / 16)) * p.n2 + n2];

    if (p.has_nl) {
        const R hh = (R)ctl.h / (R)2;                           // h_/2
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const R pw = v[q].x * v[q].x + v[q].y * v[q].y;     // |A|^2
            const R ph = mul_rn(hh, mul_rn(p.gamma, pw));       // (h_/2) * (gamma |A|^2)
            strow[(size_t)(t + q * (M / 16)) * p.n2 + n2] = ph;
            R s, co; sincos_r(ph, &s, &co);
            v[q] = cmul(v[q], mk<R>(co, s));
        }
    }
    __syncthreads();
    fft_passes<R, M, -1, ColExchange<T> >::run(v, sm + c, tw, t);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k1 = t + q * (M / 16);
        v[q] = cmul(v[q], fourstep_twiddle<R>(p, n2, k1));
        rowp[(size_t)k1 * p.n2 + n2] = v[q];
    }
}

// ---------------------------------------------------------------------------------------------
// row pass: forward transform, linear operator, inverse transform
// ---------------------------------------------------------------------------------------------
template <typename R, int M, int G>
__global__ void __launch_bounds__(G * (M / 16), min_ctas(G * (M / 16))) k_row(Params<R> p) {
    typedef typename cx_of<R>::type C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PM = pad16(M) + 1;
    C* sm = reinterpret_cast<C*>(smem_raw);                   // [G][PM] private exchange buffers
    C* tw = sm + G * PM;

    const int g = threadIdx.x / (M / 16), t = threadIdx.x % (M / 16);
    const long long grow = (long long)blockIdx.x * G + g;       // global row index over [B*P][N1]
    const int bp = (int)(grow / p.n1), k1 = (int)(grow % p.n1);
    const int b = bp / p.n_pol;
    const Ctrl ctl = p.ctrl[b];
    if (ctl.done) return;                                       // G divides N1: uniform per block

    load_tables(tw, p.tw_row, fft_plan<M>::table_size);
    C* base = p.field + (size_t)bp * p.n + (size_t)k1 * p.n2;
    C v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) v[q] = base[t + q * (M / 16)];
    __syncthreads();
    fft_passes<R, M, -1, RowExchange<M> >::run(v, sm + g * PM, tw, t);

    {   // exp(D~ h): real part -alpha/2*h (attenuation), imaginary part (b2/2 w^2 + b3/6 w^3) h
        const R h = (R)ctl.h;
        const R att = exp_r(mul_rn(p.att_half, h));
        const int half = p.n >> 1;
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int k2 = t + q * (M / 16);
            int k = k1 + p.n1 * k2;                             // transposed-order bin index
            k = (k < half) ? k : k - p.n;                       // fftfreq ordering
            const double w64 = ((double)k * p.fval * 2.0) * 3.141592653589793 * 1e-12;   // rad/ps
            const R w = (R)w64;
            const R dim = add_rn(mul_rn(p.c2, mul_rn(w, w)), mul_rn(p.c3, cube_r(w)));
            const R ph = mul_rn(dim, h);
            R s, co; sincos_r(ph, &s, &co);
            v[q] = cmul(v[q], mk<R>(att * co, att * s));
        }
    }
    fft_passes<R, M, +1, RowExchange<M> >::run(v, sm + g * PM, tw, t);
#pragma unroll
    for (int q = 0; q < 16; ++q) base[t + q * (M / 16)] = v[q];
}

// ---------------------------------------------------------------------------------------------
// column pass, inverse, second Kerr half step, power max, controller
// ---------------------------------------------------------------------------------------------
template <typename R, int M, int T>
__global__ void __launch_bounds__(T * (M / 16), min_ctas(T * (M / 16))) k_col_inv(Params<R> p) {
    typedef typename cx_of<R>::type C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long red[32];
    C* sm = reinterpret_cast<C*>(smem_raw);
    C* tw = sm + M * T;

    const int tiles = p.n2 / T;
    const int tile = blockIdx.x % tiles, row = blockIdx.x / tiles;
    const int b = row / p.n_pol;
    if (p.ctrl[b].done) return;

    const int c = threadIdx.x % T, t = threadIdx.x / T;
    const int n2 = tile * T + c;
    load_tables(tw, p.tw_col, fft_plan<M>::table_size);

    C* rowp = p.field + (size_t)row * p.n;
    const R* strow = p.stash + (size_t)row * p.n;
    C v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k1 = t + q * (M / 16);
        v[q] = cmulc(rowp[(size_t)k1 * p.n2 + n2], fourstep_twiddle<R>(p, n2, k1));
    }
    __syncthreads();
    fft_passes<R, M, +1, ColExchange<T> >::run(v, sm + c, tw, t);

    R pm = 0;
    bool nan = false;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const size_t off = (size_t)(t + q * (M / 16)) * p.n2 + n2;
        C a = v[q];
        a.x *= p.inv_n; a.y *= p.inv_n;                         // numpy ifft scaling (exact: N = 2^n)
        if (p.has_nl) {
            R s, co; sincos_r(strow[off], &s, &co);
            a = cmul(a, mk<R>(co, s));
        }
        const R pw = a.x * a.x + a.y * a.y;
        nan |= (pw != pw);
        pm = pw > pm ? pw : pm;
        rowp[off] = a;
    }
    if (nan) pm = pw_nan<R>();
    pm = block_max_bits<R>(pm, red);

    if (threadIdx.x == 0) {
        atomicMax(&p.ctrl[b].pmax, ord_bits(pm));
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
//   conj twiddle, inverse column transforms, 1/N, max |A|^2  ->  per-waveform barrier + controller
//   ->  rotation by the stashed phase of step s PLUS the first Kerr half step of step s+1
//       (one sincos), stash of the new phase, forward column transforms, twiddle.
// Field traffic per step drops from 3R+3W to 2R+2W (the ideal of SURVEY.md §8(d)).
//
// The barrier spans the tiles of ONE waveform (they need its global max, devices.py:1194).  CTAs
// take a ticket when they start, and the ticket -- not blockIdx -- selects the tile, so the tiles
// of a waveform are started in order and the lowest unfinished waveform always has all its tiles
// resident: no deadlock as long as tiles-per-waveform <= resident CTAs (checked by the host).
// ---------------------------------------------------------------------------------------------
template <typename R, int M, int T>
__global__ void __launch_bounds__(T * (M / 16), min_ctas(T * (M / 16))) k_col_mid(Params<R> p) {
    typedef typename cx_of<R>::type C;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long red[32];
    __shared__ unsigned int s_ticket;
    C* sm = reinterpret_cast<C*>(smem_raw);
    C* tw = sm + M * T;

    if (threadIdx.x == 0) {
        const unsigned int tk = atomicAdd(p.ticket, 1u);
        if (tk == gridDim.x - 1) *p.ticket = 0u;             // last CTA to start re-arms the counter
        s_ticket = tk;
    }
    __syncthreads();
    const int blk = (int)s_ticket;
    const int tiles = p.n2 / T;
    const int tile = blk % tiles, row = blk / tiles;
    const int b = row / p.n_pol;
    Ctrl* ctl = p.ctrl + b;
    if (ctl->done) return;
    const int step_before = ctl->steps;

    const int c = threadIdx.x % T, t = threadIdx.x / T;
    const int n2 = tile * T + c;
    load_tables(tw, p.tw_col, fft_plan<M>::table_size);

    C* rowp = p.field + (size_t)row * p.n;
    R* strow = p.stash + (size_t)row * p.n;
    C v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k1 = t + q * (M / 16);
        v[q] = cmulc(rowp[(size_t)k1 * p.n2 + n2], fourstep_twiddle<R>(p, n2, k1));
    }
    __syncthreads();
    fft_passes<R, M, +1, ColExchange<T> >::run(v, sm + c, tw, t);

    R pm = 0;
    bool nan = false;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        v[q].x *= p.inv_n; v[q].y *= p.inv_n;                  // numpy ifft scaling (exact: N = 2^n)
        const R pw = v[q].x * v[q].x + v[q].y * v[q].y;        // the Kerr rotations do not change |A|
        nan |= (pw != pw);
        pm = pw > pm ? pw : pm;
    }
    if (nan) pm = pw_nan<R>();
    pm = block_max_bits<R>(pm, red);

    // ---- per-waveform barrier; the last tile to arrive runs the controller ------------------------
    if (threadIdx.x == 0) {
        atomicMax(&ctl->pmax, ord_bits(pm));
        __threadfence();
        const unsigned total = (unsigned)(tiles * p.n_pol);
        const unsigned prev = atomicAdd(&ctl->arrived, 1u);
        if (prev + 1u == total) {
            __threadfence();
            const R all = from_bits<R>(atomicMax(&ctl->pmax, 0ull));
            controller_update<R>(p, b, all);
        } else {
            while (*reinterpret_cast<volatile int*>(&ctl->steps) == step_before) __nanosleep(64);
        }
        __threadfence();
    }
    __syncthreads();
    const int done = *reinterpret_cast<volatile int*>(&ctl->done);
    const R hh = (R)(*reinterpret_cast<volatile double*>(&ctl->h)) / (R)2;    // h_/2 of the NEXT step

    if (done) {                                                 // last step of this waveform: time domain out
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const size_t off = (size_t)(t + q * (M / 16)) * p.n2 + n2;
            if (p.has_nl) {
                R s, co; sincos_r(strow[off], &s, &co);
                v[q] = cmul(v[q], mk<R>(co, s));
            }
            rowp[off] = v[q];
        }
        return;
    }
    if (p.has_nl) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const size_t off = (size_t)(t + q * (M / 16)) * p.n2 + n2;
            const R pw = v[q].x * v[q].x + v[q].y * v[q].y;
            const R ph = mul_rn(hh, mul_rn(p.gamma, pw));       // first half step of the next step
            const R tot = strow[off] + ph;                      // + second half step of this one
            strow[off] = ph;
            R s, co; sincos_r(tot, &s, &co);
            v[q] = cmul(v[q], mk<R>(co, s));
        }
    }
    fft_passes<R, M, -1, ColExchange<T> >::run(v, sm + c, tw, t);
#pragma unroll
    for (int q = 0; q < 16; ++q) {
        const int k1 = t + q * (M / 16);
        v[q] = cmul(v[q], fourstep_twiddle<R>(p, n2, k1));
        rowp[(size_t)k1 * p.n2 + n2] = v[q];
    }
}

}  // namespace ssfm
