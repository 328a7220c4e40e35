// k_wf: the whole FIBER / DBP propagation of a batch as ONE persistent kernel (sm_100a).
//
// Reference loop: opticomlib/devices.py:1155-1196.  The multi-launch schedule of ssfm_kernels.cuh streams
// every waveform through HBM twice per step (k_row, k_col_mid).  Here a TEAM of `total` co-resident CTAs
// (total = n_pol * N/4096) adopts one waveform and carries it through ALL of its steps before it takes
// the next one, so that
//   * the field of the waveforms in flight (teams x N samples: a few MiB .. 37 MiB) never leaves the
//     126 MB L2 -- HBM sees one read and one write of the field per PROPAGATION, not four per step;
//   * the Kerr-phase stash of a tile lives in the shared memory of the CTA that owns the tile (the same
//     CTA visits the same tile every step): the stash traffic of the multi-launch schedule is gone;
//   * pass tables and the sincos table are loaded once per CTA; there are no launches, tickets or host
//     polls inside a propagation, and the step-size controller runs redundantly in every CTA.
//
// Per step each CTA runs a ROW phase (G rows of the N1 x N2 matrix: forward transform, exp(D~ h),
// inverse transform) and a COLUMN phase (T columns: inverse transform, 1/N, max|A|^2 -> team exchange
// -> controller -> merged Kerr rotation of the second half step of step s and the first half step of
// step s+1 -> forward transform), separated by team barriers (a monotonic arrival counter in L2).
// Loads of the field bypass L1 (ld.global.cg): the data was written by other SMs one phase earlier.
//
// Everything here is FP64/FP32 FMA-pipe arithmetic on L2-resident data; no tensor cores (no dense
// contraction on this path).
#pragma once
#include "ssfm_kernels.cuh"

namespace ssfm {

template <typename R>
struct WfArgs {
    unsigned int* bar;            // [n_teams][32]  monotonic arrival counter of the team barrier (one 128-B line each)
    unsigned long long* mail;     // [n_teams][16]  (sequence << 32 | waveform) handed out by CTA 0 of the team
    unsigned long long* slots;    // [n_teams][2][total][2]  self-validating max words, double-buffered by exchange parity
    unsigned int* next_wf;        // next waveform to hand out (dynamic assignment: step counts differ per waveform)
    long long budget;             // stop every waveform after this many steps in this call (max_steps of the C-ABI)
    int n_teams;
    int fixed, single, resume;
    R h_fixed;
};

__host__ __device__ constexpr int wf_cmax(int a, int b) { return a > b ? a : b; }

template <typename R, int M1, int M2>
struct wf_geom {
    typedef typename cx_of<R>::type C;
    static constexpr int E = 16, NT = 256;
    static constexpr int T = 4096 / M1, G = 4096 / M2;
    static constexpr int PM = RowExchange<M2, E>::size;
    static constexpr int TAB1 = fft_plan<M1, E>::table_size;
    static constexpr int TAB2 = (M1 == M2) ? 0 : fft_plan<M2, E>::table_size;
    static constexpr int XB = wf_cmax(M1 * T, G * PM);
    static constexpr size_t smem_full = sizeof(C) * (size_t)(XB + TAB1 + TAB2 + SC_N) + sizeof(R) * (size_t)(E * NT);
    // Two CTAs per SM need <= (228 KB / 2 - 1 KB) each.  When the pass tables do not fit next to the exchange buffer
    // and the stash (fp64, transforms of 1024 points or N1 != N2 >= 512) they are read through L1 from global memory.
    static constexpr bool TABS = smem_full <= 115712;
    static constexpr size_t smem = TABS ? smem_full : smem_full - sizeof(C) * (size_t)(TAB1 + TAB2);
};

__device__ __forceinline__ unsigned int ld_volatile_u32(const unsigned int* p) {
    return *reinterpret_cast<const volatile unsigned int*>(p);
}

template <typename R, int M1, int M2, bool SMALL>
__global__ void __launch_bounds__(256, (sizeof(R) == 8 ? 2 : 3)) k_wf(Params<R> p, WfArgs<R> a) {
    typedef typename cx_of<R>::type C;
    typedef wf_geom<R, M1, M2> GEO;
    constexpr int E = GEO::E, NT = GEO::NT, T = GEO::T, G = GEO::G, PM = GEO::PM;
    static_assert(points_per_thread<R>::value == 16, "k_wf assumes 16 points per thread");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long red[32];
    __shared__ unsigned int s_w;
    C* xb = reinterpret_cast<C*>(smem_raw);                    // exchange buffer: column tile [M1][T] or G padded rows
    constexpr bool TABS = GEO::TABS;
    C* tw1s = xb + GEO::XB;                                    // pass tables of the N1-point (column) transforms
    C* tw2s = tw1s + (TABS ? GEO::TAB1 : 0);                   // ... of the N2-point (row) transforms when N2 != N1
    C* sct = tw2s + (TABS ? GEO::TAB2 : 0);                    // sincos table
    R* st_sm = reinterpret_cast<R*>(sct + SC_N);               // [E][NT] Kerr phase of the current step of MY tile
    const C* tw1 = TABS ? tw1s : p.tw_col;
    const C* tw2 = TABS ? ((M1 == M2) ? tw1s : tw2s) : p.tw_row;

    const int tid = threadIdx.x;
    const int team = blockIdx.x % a.n_teams, me = blockIdx.x / a.n_teams;
    const int tiles = p.n2 / T;                                // column tiles (= row groups) per polarisation
    const unsigned total = (unsigned)(tiles * p.n_pol);        // CTAs per team
    const int pol = me / tiles, tile = me % tiles;
    const int c = tid % T, t = tid / T;                        // column phase: column c of the tile, thread t of its transform
    const int n2 = tile * T + c;
    const int g = tid / (M2 / E), tr = tid % (M2 / E);         // row phase: row g of the group, thread tr of its transform
    const int k1 = tile * G + g;

    if (TABS) {
        for (int i = tid; i < GEO::TAB1; i += NT) tw1s[i] = p.tw_col[i];
        for (int i = tid; i < GEO::TAB2; i += NT) tw2s[i] = p.tw_row[i];
    }
    for (int i = tid; i < SC_N; i += NT) sct[i] = p.tw_col[GEO::TAB1 + i];

    unsigned int bar_target = 0, xchg = 0, seq = 0;
    unsigned int* bar = a.bar + team * 32;

    auto team_barrier = [&]() {
        __syncthreads();
        if (tid == 0) {
            bar_target += total;
            __threadfence();
            atomicAdd(bar, 1u);
            while ((int)(ld_volatile_u32(bar) - bar_target) < 0) __nanosleep(20);
            __threadfence();
        }
        __syncthreads();
    };

    // max over the team of a per-thread value (NaN wins, like numpy's max): block reduction, one self-validating
    // word (two for double) per CTA -- {32 value bits | 32-bit exchange tag}: the flag travels with the data, so no
    // fence and no atomic is needed -- and every thread polls a share of the team's words.
    auto team_max = [&](R pm) -> R {
        constexpr int NW = sizeof(R) / 4;
        ++xchg;
        const unsigned long long tag = (unsigned long long)xchg;
        volatile unsigned long long* wf = a.slots + ((size_t)(team * 2 + (xchg & 1u)) * total) * 2;
        const R mine = block_max_bits<R>(pm, red);
        const unsigned long long bits = ord_bits(mine);
        if (tid < NW) {
            const unsigned long long part = (NW == 1) ? (bits & 0xffffffffull) : (tid == 0 ? (bits >> 32) : (bits & 0xffffffffull));
            wf[me * 2 + tid] = (part << 32) | tag;
        }
        unsigned long long best = 0ull;
        const int nwords = (int)total * NW;
        for (int base = 0; base < nwords; base += NT) {         // uniform trip count; lanes pair up (hi, lo) for double
            const int idx = base + tid;
            const bool have = idx < nwords;
            unsigned long long x = tag;
            if (have) {
                for (;;) {
                    x = wf[(idx / NW) * 2 + (idx % NW)];
                    if ((x & 0xffffffffull) == tag) break;
                    __nanosleep(20);
                }
            }
            unsigned long long val = have ? (x >> 32) : 0ull;
            if (NW == 2) {
                const unsigned long long other = __shfl_xor_sync(0xffffffffu, val, 1);
                val = (tid & 1) ? 0ull : ((val << 32) | other);
            }
            best = val > best ? val : best;
        }
        __syncthreads();                                        // red[0] of the first reduction has been read by everyone
        return block_max_bits<R>(from_bits<R>(best), red);
    };

    for (;;) {
        // ------------------------------------------------------------------ next waveform of this team
        ++seq;
        if (tid == 0) {
            volatile unsigned long long* mb = a.mail + team * 16;
            unsigned int wn;
            if (me == 0) {
                wn = atomicAdd(a.next_wf, 1u);
                *mb = ((unsigned long long)seq << 32) | (unsigned long long)wn;
            } else {
                unsigned long long m;
                for (;;) { m = *mb; if ((unsigned int)(m >> 32) == seq) break; __nanosleep(40); }
                wn = (unsigned int)m;
            }
            s_w = wn;
        }
        __syncthreads();
        const unsigned int w = s_w;
        __syncthreads();
        if (w >= (unsigned int)p.batch) break;

        C* __restrict__ rowp = p.field + ((size_t)w * p.n_pol + pol) * p.n;
        C* __restrict__ rbase = rowp + (size_t)k1 * p.n2;
        C v[E];
        R z, h;
        int steps;
        long long taken = 0;

        // ------------------------------------------------------------------ prologue: first step size, first Kerr
        // half step (devices.py:1155-1161, 1175-1177), forward column transforms, four-step twiddle
#pragma unroll
        for (int q = 0; q < E; ++q) v[q] = __ldcg(rowp + (size_t)(t + q * (M1 / E)) * p.n2 + n2);
        if (a.resume) {
            const Ctrl cs = p.ctrl[w];
            if (cs.done) continue;
            z = (R)cs.z; h = (R)cs.h; steps = cs.steps;
        } else {
            R h0;
            if (a.fixed) h0 = a.h_fixed;
            else if (a.single) h0 = p.length;                   // no dispersion or no Kerr effect: one step
            else {
                R pm = 0;
                bool nan = false;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const R pw = v[q].x * v[q].x + v[q].y * v[q].y;
                    nan |= (pw != pw);
                    pm = pw > pm ? pw : pm;
                }
                if (nan) pm = pw_nan<R>();
                h0 = p.phi_max / mul_rn(p.abs_gamma, team_max(pm));
            }
            h = (p.length < h0) ? p.length : h0;                // python min(h_, length)
            z = 0; steps = 0;
            const int done0 = !((R)0 < p.length) || a.budget <= 0;
            if (me == 0 && tid == 0) {
                Ctrl& cs = p.ctrl[w];
                cs.z = 0.0; cs.h = (double)h; cs.pmax = 0ull; cs.steps = 0; cs.arrived = 0u; cs.done = !((R)0 < p.length);
            }
            if (done0) continue;
        }
        if (p.has_nl) {
            const R hh = h / (R)2;                              // h_/2
#pragma unroll
            for (int q = 0; q < E; ++q) {
                const R pw = v[q].x * v[q].x + v[q].y * v[q].y; // |A|^2
                const R ph = mul_rn(hh, mul_rn(p.gamma, pw));   // (h_/2) * (gamma |A|^2)
                st_sm[q * NT + tid] = ph;
                R s, co; kerr_sincos<SMALL>(ph, sct, &s, &co);
                v[q] = cmul(v[q], mk<R>(co, s));
            }
        }
        fft_passes<R, M1, -1, ColExchange<T>, E>::run(v, xb + c, tw1, t);
        apply_fourstep<false, R, E, M1>(p, v, n2, t);
#pragma unroll
        for (int q = 0; q < E; ++q) rowp[(size_t)(t + q * (M1 / E)) * p.n2 + n2] = v[q];
        team_barrier();

        for (;;) {
            // -------------------------------------------------------------- row phase (devices.py:1178-1180)
#pragma unroll
            for (int q = 0; q < E; ++q) v[q] = __ldcg(rbase + tr + q * (M2 / E));
            fft_passes<R, M2, -1, RowExchange<M2, E>, E>::run(v, xb + g * PM, tw2, tr);
            {
                const int half = p.n >> 1;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const int k2 = tr + q * (M2 / E);
                    int k = k1 + p.n1 * k2;                     // transposed-order bin index
                    k = (k < half) ? k : k - p.n;               // fftfreq ordering
                    const R wk = (R)((double)k * p.wscale);     // rad/ps (see Params::wscale)
                    const R dim = add_rn(mul_rn(p.c2, mul_rn(wk, wk)), mul_rn(p.c3, cube_r(wk)));
                    const R ph = mul_rn(dim, h);
                    R s, co; sincos_r(ph, sct, &s, &co);
                    v[q] = cmul(v[q], mk<R>(co, s));
                }
            }
            fft_passes<R, M2, +1, RowExchange<M2, E>, E>::run(v, xb + g * PM, tw2, tr);
#pragma unroll
            for (int q = 0; q < E; ++q) rbase[tr + q * (M2 / E)] = v[q];
            team_barrier();

            // -------------------------------------------------------------- column phase: end of step s
#pragma unroll
            for (int q = 0; q < E; ++q) v[q] = __ldcg(rowp + (size_t)(t + q * (M1 / E)) * p.n2 + n2);
            apply_fourstep<true, R, E, M1>(p, v, n2, t);
            fft_passes<R, M1, +1, ColExchange<T>, E>::run(v, xb + c, tw1, t);
            const R sc = p.inv_n * exp_r(mul_rn(p.att_half, h)); // 1/N (exact) and exp(-alpha/2 h) (real part of D~ h)
#pragma unroll
            for (int q = 0; q < E; ++q) { v[q].x *= sc; v[q].y *= sc; }
            R pmax = 0;
            if (p.adaptive) {                                   // devices.py:1194: max over the whole waveform
                R pm = 0;
                bool nan = false;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const R pw = v[q].x * v[q].x + v[q].y * v[q].y;   // the Kerr rotations do not change |A|
                    nan |= (pw != pw);
                    pm = pw > pm ? pw : pm;
                }
                if (nan) pm = pw_nan<R>();
                pmax = team_max(pm);
            }
            const CtrlNext<R> nx = controller_next<R>(p, z, h, steps, pmax);
            ++taken;
            const bool stop = nx.done || taken >= a.budget;
            if (me == 0 && tid == 0) {
                Ctrl& cs = p.ctrl[w];
                if (p.hlog && steps < p.hlog_cap) p.hlog[(size_t)w * p.hlog_cap + steps] = (double)h;
                cs.z = (double)nx.z; cs.h = (double)nx.h; cs.steps = steps + 1; cs.done = nx.done;
            }
            if (stop) {                                         // second Kerr half step, time domain out
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    if (p.has_nl) {
                        R s, co; kerr_sincos<SMALL>(st_sm[q * NT + tid], sct, &s, &co);
                        v[q] = cmul(v[q], mk<R>(co, s));
                    }
                    rowp[(size_t)(t + q * (M1 / E)) * p.n2 + n2] = v[q];
                }
                break;
            }
            // -------------------------------------------------------------- ... and start of step s+1
            if (p.has_nl) {
                const R hh = nx.h / (R)2;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const R pw = v[q].x * v[q].x + v[q].y * v[q].y;
                    const R ph = mul_rn(hh, mul_rn(p.gamma, pw));   // first half step of the next step
                    const R tot = st_sm[q * NT + tid] + ph;         // + second half step of this one
                    st_sm[q * NT + tid] = ph;
                    R s, co; kerr_sincos<SMALL>(tot, sct, &s, &co);
                    v[q] = cmul(v[q], mk<R>(co, s));
                }
            }
            fft_passes<R, M1, -1, ColExchange<T>, E>::run(v, xb + c, tw1, t);
            apply_fourstep<false, R, E, M1>(p, v, n2, t);
#pragma unroll
            for (int q = 0; q < E; ++q) rowp[(size_t)(t + q * (M1 / E)) * p.n2 + n2] = v[q];
            z = nx.z; h = nx.h; ++steps;
            team_barrier();
        }
    }
}

}  // namespace ssfm
