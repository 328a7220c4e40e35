// In-register / shared-memory Stockham FFT building blocks for sm_100a.
//
// Every thread owns E complex points (E = 16 for complex64, 8 for complex128) of one M-point
// transform (M = 2^m, 16 <= M <= 2048), always in the "load layout"   v[q] = x[t + q*M/E],
// t = thread index inside the transform.  A transform is a sequence of radix passes
// (E, E, ..., tail radix 2|4|8|16); between two passes the points are exchanged through shared memory.  Natural order in, natural order out,
// and the output is again in the load layout, so a forward transform, a point-wise operator and
// an inverse transform chain without touching memory (that is what the SSFM row kernel does).
//
// No tensor cores: the transform is not a dense contraction.  Arithmetic is plain FP32/FP64 FMA
// pipe work; the exchanges are 64/128-bit shared-memory accesses, conflict-free by construction
// (interleaved columns, or a one-in-sixteen padding for contiguous rows).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ssfm {

template <typename R> struct cx_of;
template <> struct cx_of<float>  { typedef float2  type; };
template <> struct cx_of<double> { typedef double2 type; };

template <typename R> __device__ __forceinline__ typename cx_of<R>::type mk(R x, R y) {
    typename cx_of<R>::type c; c.x = x; c.y = y; return c;
}
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) {
    C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r;
}
// a * conj(b)
template <typename C> __device__ __forceinline__ C cmulc(C a, C b) {
    C r; r.x = a.x * b.x + a.y * b.y; r.y = a.y * b.x - a.x * b.y; return r;
}
// multiply by  S*j  (S = -1: forward transform, S = +1: inverse)
template <int S, typename C> __device__ __forceinline__ C mulj(C a) {
    C r;
    if (S > 0) { r.x = -a.y; r.y = a.x; } else { r.x = a.y; r.y = -a.x; }
    return r;
}
// multiply by the stored (forward, e^{-j..}) twiddle for S = -1, by its conjugate for S = +1
template <int S, typename C> __device__ __forceinline__ C twmul(C a, C w) {
    return (S < 0) ? cmul(a, w) : cmulc(a, w);
}

// ---------------------------------------------------------------------------------------------
// Small DFTs on register arrays, natural order in and out.  S = -1 forward (e^{-j2pi/R}).
// ---------------------------------------------------------------------------------------------
template <int S, typename C> __device__ __forceinline__ void dft2(C& a0, C& a1) {
    C t = a0; a0 = cadd(t, a1); a1 = csub(t, a1);
}

template <int S, typename C> __device__ __forceinline__ void dft4(C& a0, C& a1, C& a2, C& a3) {
    C t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mulj<S>(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2);
    a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

// multiply by W8^1 = (1 + S j)/sqrt2 and W8^3 = (-1 + S j)/sqrt2
template <int S, typename C> __device__ __forceinline__ C mulw8_1(C a) {
    typedef decltype(a.x) R;
    const R h = (R)0.70710678118654752440;
    C r;
    if (S < 0) { r.x = (a.x + a.y) * h; r.y = (a.y - a.x) * h; }
    else       { r.x = (a.x - a.y) * h; r.y = (a.y + a.x) * h; }
    return r;
}
template <int S, typename C> __device__ __forceinline__ C mulw8_3(C a) {
    typedef decltype(a.x) R;
    const R h = (R)0.70710678118654752440;
    C r;
    if (S < 0) { r.x = (a.y - a.x) * h; r.y = -(a.x + a.y) * h; }
    else       { r.x = -(a.x + a.y) * h; r.y = (a.x - a.y) * h; }
    return r;
}

template <int S, typename C> __device__ __forceinline__ void dft8(C (&a)[8]) {
    // even / odd 4-point transforms, then the W8^k combination
    dft4<S>(a[0], a[2], a[4], a[6]);
    dft4<S>(a[1], a[3], a[5], a[7]);
    C o1 = mulw8_1<S>(a[3]);
    C o2 = mulj<S>(a[5]);
    C o3 = mulw8_3<S>(a[7]);
    C e0 = a[0], e1 = a[2], e2 = a[4], e3 = a[6], o0 = a[1];
    a[0] = cadd(e0, o0); a[4] = csub(e0, o0);
    a[1] = cadd(e1, o1); a[5] = csub(e1, o1);
    a[2] = cadd(e2, o2); a[6] = csub(e2, o2);
    a[3] = cadd(e3, o3); a[7] = csub(e3, o3);
}

template <int S, typename C> __device__ __forceinline__ void dft16(C (&a)[16]) {
    typedef decltype(a[0].x) R;
    const R c1 = (R)0.92387953251128675613;  // cos(pi/8)
    const R s1 = (R)0.38268343236508977173;  // sin(pi/8)
    // stage A: four 4-point transforms over n1 (stride 4), results A_{n2}[k1] left at a[4*k1 + n2]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) dft4<S>(a[n2], a[4 + n2], a[8 + n2], a[12 + n2]);
    // twiddles W16^{n2*k1}
    C w1 = mk<R>(c1, -s1), w3 = mk<R>(s1, -c1);  // forward values; twmul conjugates for S=+1
    a[5]  = twmul<S>(a[5], w1);            // n2=1,k1=1 : W^1
    a[9]  = mulw8_1<S>(a[9]);              // n2=1,k1=2 : W^2
    a[13] = twmul<S>(a[13], w3);           // n2=1,k1=3 : W^3
    a[6]  = mulw8_1<S>(a[6]);              // n2=2,k1=1 : W^2
    a[10] = mulj<S>(a[10]);                // n2=2,k1=2 : W^4
    a[14] = mulw8_3<S>(a[14]);             // n2=2,k1=3 : W^6
    a[7]  = twmul<S>(a[7], w3);            // n2=3,k1=1 : W^3
    a[11] = mulw8_3<S>(a[11]);             // n2=3,k1=2 : W^6
    {                                      // n2=3,k1=3 : W^9 = -W^1
        C t = twmul<S>(a[15], w1); a[15].x = -t.x; a[15].y = -t.y;
    }
    // stage B: for each k1 a 4-point transform over n2; X[k1 + 4*k2] lands at a[4*k1 + k2]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) dft4<S>(a[4 * k1], a[4 * k1 + 1], a[4 * k1 + 2], a[4 * k1 + 3]);
    // transpose the 4x4 register tile so that a[k] = X[k]
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1)
#pragma unroll
        for (int k2 = k1 + 1; k2 < 4; ++k2) { C t = a[4 * k1 + k2]; a[4 * k1 + k2] = a[4 * k2 + k1]; a[4 * k2 + k1] = t; }
}

// ---------------------------------------------------------------------------------------------
// Points per thread.  E = 16 for complex64 (32 data registers); E = 8 for complex128 (32 data
// registers as well): with 16 complex128 points a thread needs 128 registers, two 256-thread CTAs
// fill the register file and the SM runs 16 warps -- measured latency-bound.  Eight points per thread
// cost one more radix pass per transform but allow 24 warps per SM.
// ---------------------------------------------------------------------------------------------
#ifndef SSFM_E64
#define SSFM_E64 16   // measured: 8 points per thread (24 warps/SM) gains 5 % in the column kernel but loses 17 % in the
#endif                // row kernel (one more pass = more shared-memory wavefronts; the LSU path is the busiest unit)
template <typename R> struct points_per_thread { static constexpr int value = sizeof(R) == 8 ? SSFM_E64 : 16; };

// ---------------------------------------------------------------------------------------------
// Pass tables.  For a pass (Ns, R) with Ns > 1 the table holds W_{Ns*R}^{r*jm} at [(r-1)*Ns + jm],
// r = 1..R-1, jm = 0..Ns-1 (forward sign).  Tables of consecutive passes are concatenated.
// ---------------------------------------------------------------------------------------------
template <int M, int E> struct fft_plan {
    __host__ __device__ static constexpr int radix_at(int ns) { return (M / ns >= E) ? E : (M / ns); }
    __host__ __device__ static constexpr int table_size_from(int ns) {
        return (ns >= M) ? 0 : ((ns > 1 ? (radix_at(ns) - 1) * ns : 0) + table_size_from(ns * radix_at(ns)));
    }
    static constexpr int table_size = table_size_from(1);
    __host__ __device__ static constexpr int table_offset(int ns_target, int ns = 1) {  // offset of the table of pass `ns_target`
        return (ns >= ns_target) ? 0 : ((ns > 1 ? (radix_at(ns) - 1) * ns : 0) + table_offset(ns_target, ns * radix_at(ns)));
    }
};

template <int E> __host__ __device__ constexpr int pad_e(int a) { return a + a / E; }

// Exchange policies -----------------------------------------------------------------------------
// Column transforms: T transforms interleaved, element i of column c at sm[i*T + c]; the threads of
// one transform sit in different warps -> block-wide barrier.
template <int T> struct ColExchange {
    static constexpr int stride = T;
    __device__ static __forceinline__ int idx(int a) { return a * T; }
    __device__ static __forceinline__ void sync() { __syncthreads(); }
};
// The same layout for T = 16 columns handled as two independent halves of 8 columns by threads 0..127 / 128..255 of a 256-thread
// CTA: each half synchronises on its own named barrier, so the halves drift apart and one half's butterflies overlap the other
// half's exchange (and each barrier waits for 4 warps instead of 8).
template <int T> struct ColExchangeHalves {
    static constexpr int stride = T;
    __device__ static __forceinline__ int idx(int a) { return a * T; }
    __device__ static __forceinline__ void sync() { asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)(threadIdx.x >> 7)) : "memory"); }
};
// Row transforms: one padded private buffer per transform (one pad slot every E points keeps both the
// scattered writes and the contiguous reads conflict-free); the M/E threads of one transform are
// consecutive, so when they fit in a warp a warp barrier is enough.
template <int M, int E> struct RowExchange {
    static constexpr int stride = 1;
    static constexpr int size = pad_e<E>(M) + 1;
    __device__ static __forceinline__ int idx(int a) { return pad_e<E>(a); }
    __device__ static __forceinline__ void sync() {
        if (M / E <= 32) __syncwarp(); else __syncthreads();
    }
};

// How a pass multiplies its inputs by W_{Ns*Rr}^{r*jm}, r = 1 .. Rr-1.  TableTwiddles: one table entry per r (the pass tables
// of fft_plan; shared memory, or global memory read through L1).  ChainTwiddles: only W^{jm} is fetched, the powers are built
// with 14 complex products arranged as four chains of depth <= 5 -- for kernels whose LSU path is busier than their FMA pipe.
struct TableTwiddles {
    template <int M, int E, int Ns, int S, int Rr, typename C>
    __device__ static __forceinline__ void apply(C (&a)[Rr], const C* tw, int jm) {
        const C* tp = tw + fft_plan<M, E>::table_offset(Ns) + jm;
#pragma unroll
        for (int r = 1; r < Rr; ++r) a[r] = twmul<S>(a[r], tp[(r - 1) * Ns]);
    }
};
struct ChainTwiddles {
    template <int M, int E, int Ns, int S, int Rr, typename C>
    __device__ static __forceinline__ void apply(C (&a)[Rr], const C* tw, int jm) {
        if constexpr (Rr != 16) TableTwiddles::apply<M, E, Ns, S, Rr>(a, tw, jm);
        else {
            C w[4];
            w[0] = tw[fft_plan<M, E>::table_offset(Ns) + jm];          // W^{jm} (the r = 1 row of the pass table)
            w[1] = cmul(w[0], w[0]);
            w[2] = cmul(w[1], w[0]);
            w[3] = cmul(w[1], w[1]);
            const C w4 = w[3];
#pragma unroll
            for (int r = 1; r < 16; ++r) {
                C& wr = w[(r - 1) & 3];
                if (r > 4) wr = cmul(wr, w4);                           // W^{r jm} = W^{(r-4) jm} W^{4 jm}
                a[r] = twmul<S>(a[r], wr);
            }
        }
    }
};

// One M-point transform on v[E] (load layout v[q] = x[t + q*M/E]).  `sm` points at this transform's exchange
// buffer (already offset by the column for ColExchange), `tw` at the pass tables in shared memory.
template <typename R, int M, int S, typename X, int E, int Ns = 1, typename TW = TableTwiddles>
struct fft_passes {
    typedef typename cx_of<R>::type C;
    static constexpr int Rr = fft_plan<M, E>::radix_at(Ns);
    static constexpr int NB = E / Rr;
    __device__ static __forceinline__ void run(C (&v)[E], C* sm, const C* tw, int t) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = t + i * (M / E);
            C a[Rr];
#pragma unroll
            for (int r = 0; r < Rr; ++r) a[r] = v[i + r * NB];
            if constexpr (Ns > 1) TW::template apply<M, E, Ns, S, Rr>(a, tw, j & (Ns - 1));
            if constexpr (Rr == 16) dft16<S>(a);
            else if constexpr (Rr == 8) dft8<S>(a);
            else if constexpr (Rr == 4) dft4<S>(a[0], a[1], a[2], a[3]);
            else dft2<S>(a[0], a[1]);
#pragma unroll
            for (int r = 0; r < Rr; ++r) v[i + r * NB] = a[r];
        }
        if constexpr (Ns * Rr < M) {  // exchange, then the next pass (only full-radix passes get here: NB == 1, j == t)
            const int j0 = (t / Ns) * (Ns * Rr) + (t & (Ns - 1));
            X::sync();  // every reader of the previous exchange is done
#pragma unroll
            for (int r = 0; r < Rr; ++r) sm[X::idx(j0 + r * Ns)] = v[r];
            X::sync();
#pragma unroll
            for (int q = 0; q < E; ++q) v[q] = sm[X::idx(t + q * (M / E))];
            fft_passes<R, M, S, X, E, Ns * Rr, TW>::run(v, sm, tw, t);
        }
    }
};
template <typename R, int M, int S, typename X, int E, typename TW>
struct fft_passes<R, M, S, X, E, M, TW> {
    typedef typename cx_of<R>::type C;
    __device__ static __forceinline__ void run(C (&)[E], C*, const C*, int) {}
};

}  // namespace ssfm
