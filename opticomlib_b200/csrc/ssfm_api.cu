// C-ABI of the split-step Fourier engine (see include/ssfm_b200.h) and the host-side step loop.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <algorithm>
#include <cmath>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/ssfm_b200.h"
#include "ssfm_kernels.cuh"
#include "ssfm_internal.h"
#include "ssfm_wf.h"
#include <map>
#include <mutex>
#include <tuple>

using namespace ssfm;

thread_local std::string ssfm_err_slot;   // shared with filtfilt.cu
std::atomic<long long> ssfm_launches{0};   // kernels launched by this library (bench.py's gpu_launches); host threads

namespace {

#define g_err ssfm_err_slot

int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU_TRY(expr)                                                                         \
    do {                                                                                     \
        cudaError_t e__ = (expr);                                                            \
        if (e__ != cudaSuccess)                                                              \
            return fail(SSFM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

int ilog2(long long v) { int l = 0; while ((1ll << l) < v) ++l; return l; }

constexpr int kMaxDevices = 64;
int current_device() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < kMaxDevices) ? d : 0; }

// compile-time geometry ------------------------------------------------------------------------
// Column tile width T (columns per CTA) and row group G (rows per CTA) for transforms of M points
// with E points per thread: aim at 256-thread CTAs, keep at least 2 columns (32 B of complex128) per
// row segment, never more than 32, and never more rows/columns than the matrix has.
constexpr int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
#ifndef SSFM_COL_THREADS
#define SSFM_COL_THREADS 256      // threads per CTA of the column kernels (experiments: 128 = four smaller CTAs per SM)
#endif
constexpr int col_tile(int M, int E) { return clampi(clampi(SSFM_COL_THREADS * E / M, 2, 32), 1, M); }
constexpr int row_group(int M, int E) { return clampi(256 * E / M, 1, M / 2); }
template <typename R> constexpr int col_tile_of(int M) { return col_tile(M, points_per_thread<R>::value); }
template <typename R> constexpr int row_group_of(int M) { return row_group(M, points_per_thread<R>::value); }
int col_tile_rt(int M, int dtype) { return dtype == SSFM_C64 ? col_tile_of<float>(M) : col_tile_of<double>(M); }

// twiddle tables, built on the device in double and rounded once to R ---------------------------
template <typename R>
__global__ void k_build_unit_roots(typename cx_of<R>::type* out, int count, long long stride, long long period) {
    // out[i] = exp(-2 pi j * (i*stride) / period)
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const long long m = ((long long)i * stride) % period;
    double s, c;
    sincospi(2.0 * (double)m / (double)period, &s, &c);
    out[i] = mk<R>((R)c, (R)(-s));
}

template <typename R>
__global__ void k_build_fourstep(typename cx_of<R>::type* out, int n, int n2) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;   // out[k1*N2 + n2] = exp(-2 pi j n2 k1 / N)
    if (pos >= n) return;
    const long long m = ((long long)(pos / n2) * (pos % n2)) % n;
    double s, c;
    sincospi(2.0 * (double)m / (double)n, &s, &c);
    out[pos] = mk<R>((R)c, (R)(-s));
}

template <typename R>
__global__ void k_build_pass_table(typename cx_of<R>::type* out, int ns, int radix) {
    // out[(r-1)*ns + jm] = W_{ns*radix}^{r*jm}
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (radix - 1) * ns) return;
    const int r = i / ns + 1, jm = i % ns;
    const long long period = (long long)ns * radix;
    const long long m = ((long long)r * jm) % period;
    double s, c;
    sincospi(2.0 * (double)m / (double)period, &s, &c);
    out[i] = mk<R>((R)c, (R)(-s));
}

template <typename R>
__global__ void k_build_sincos_table(typename cx_of<R>::type* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // out[i] = (cos, sin)(2 pi i / SC_N)
    if (i >= SC_N) return;
    double s, c;
    sincospi(2.0 * (double)i / (double)SC_N, &s, &c);
    out[i] = mk<R>((R)c, (R)s);
}

int pass_table_size(int M, int E) {
    int total = 0;
    for (int ns = 1; ns < M;) {
        const int r = (M / ns >= E) ? E : M / ns;
        if (ns > 1) total += (r - 1) * ns;
        ns *= r;
    }
    return total;
}

template <typename R>
int build_pass_tables(void** out, int M, cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    constexpr int E = points_per_thread<R>::value;
    const int total = pass_table_size(M, E);
    C* d = nullptr;
    CU_TRY(cudaMalloc(&d, sizeof(C) * (size_t)(total + SC_N)));
    k_build_sincos_table<R><<<(SC_N + 127) / 128, 128, 0, st>>>(d + total);
    int off = 0;
    for (int ns = 1; ns < M;) {
        const int r = (M / ns >= E) ? E : M / ns;
        if (ns > 1) {
            const int cnt = (r - 1) * ns;
            k_build_pass_table<R><<<(cnt + 127) / 128, 128, 0, st>>>(d + off, ns, r);
            off += cnt;
        }
        ns *= r;
    }
    CU_TRY(cudaGetLastError());
    *out = d;
    return SSFM_OK;
}

}  // namespace

struct ssfm_plan_s {
    int device = 0, dtype = 0, n_pol = 1;
    long long n = 0, batch = 0;
    int log2n = 0, n1 = 0, n2 = 0;
    void *tw_col = nullptr, *tw_row = nullptr, *tw_lo = nullptr, *tw_hi = nullptr, *tw_full = nullptr;
    void* stash = nullptr;       // Kerr-phase stash of the multi-launch schedule, allocated on first use (k_wf keeps it in shared memory)
    bool propagates = false;     // false: plan created for transfer functions only
    void* xfer = nullptr;        // transfer function table (transposed order), allocated on first use
    Ctrl* ctrl = nullptr;
    int* active = nullptr;       // one counter per chunk
    unsigned int* ticket = nullptr;
    unsigned long long* slots = nullptr;   // SYNC_LL words, [batch][tiles*P][2]
    size_t slots_bytes = 0;
    int fused = 1;               // 1: fused column kernel (2R+2W per step) when its barrier fits on the chip
    int num_sms = 0;
    int debug = 0;
    int use_tw_full = -1;        // four-step twiddles: -1 auto, 0 two tables, 1 full table, 2 recurrence
    int lin_sep = -1;            // k_wf: separable linear operator where it applies (-1 auto = on, 0 off)
    int l2_ahead = 0;
    int persistent = 1;          // 1: whole propagation as one persistent kernel (k_wf) when the geometry allows it
    int teams_cap = 0;           // k_wf: at most this many teams (0 = as many as fit)
    int placement = -1;          // k_wf: SM-aware team placement (-1 = auto)
    void* dim_tab = nullptr;     // k_wf: imag(D~) per bin (transposed order), refilled by every propagation
    void* tstash = nullptr;      // k_wf, multi-tile cluster teams: Kerr phase of the waveforms in flight (allocated on first use)
    size_t tstash_bytes = 0;
    const unsigned int* s_ready = nullptr;   // ssfm_propagate_streamed: arrival / completion counters of the call in progress
    unsigned int* s_done = nullptr;
    long long s_chunk_rows = 0;
    int async_mode = 0;          // 1: ssfm_propagate returns once the persistent kernel is enqueued (host pipelines)
    int cluster = -1;            // k_wf: teams as thread-block clusters (-1 = auto, 0 = never, 1 = always when possible)
    void* wf_sync = nullptr;     // k_wf barriers / mailboxes / max words
    cudaEvent_t wf_ev[2] = {nullptr, nullptr};
    cudaStream_t wf_side = nullptr;   // k_wf: side stream of the launch that fills the slots the clusters leave
    cudaEvent_t wf_ev_side = nullptr;
    cudaStream_t peek_stream = nullptr;   // ssfm_peek_state: a copy stream of its own + one pinned record
    Ctrl* peek_host = nullptr;
    int last_kind = 0;           // schedule of the last propagate: 0 none, 1 multi-launch, 2 k_wf
    int last_teams = 0;
    int lo_bits = 0;             // split of the four-step twiddle tables (0 = log2 n2)
    // long waveforms (ssfm_long_*): this plan is the OUTER N0-point stage over this rank's column slice
    long long long_n = 0;        // global transform length N = N0 x N_l (0 = ordinary plan)
    int long_ranks = 1, long_rank = 0;
    ssfm_plan_t inner = nullptr; // N_l-point transforms of the N0 / ranks rows this rank owns after the exchange
                                 // (chirp plans: the M-point transforms of the padded rows)
    // long waveforms over several GPUs with the exchange fused into the kernels (ssfm_long_p2p_*)
    void* xbuf = nullptr;        // [time-layout field][rows-layout buffer][flags], exported to the peers through CUDA IPC
    size_t xlocal = 0;           // bytes of one layout
    void* peer_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // xbuf of every rank, mapped here
    unsigned int** d_peer_flags = nullptr;
    int p2p = 0;
    unsigned int epoch = 0;
    // arbitrary lengths (chirp-z / Bluestein): n is not a power of two; rows are padded to chirp_m = 2^ceil(log2(2n-1))
    long long chirp_m = 0;
    void *wb = nullptr, *wtab = nullptr, *xf_fwd = nullptr, *xf_inv = nullptr;   // padded work rows, chirp, chirp spectra
    int n_active = 0;
    double* hlog = nullptr;
    int hlog_cap = 0;
    int* active_host = nullptr;  // pinned, 2 slots
    cudaEvent_t ev[2] = {nullptr, nullptr};
    long long chunk = 0;
    int burst = 8;
    bool have_state = false;
    ssfm_fiber_params last{};
};

static int plan_create_impl(ssfm_plan_t* out, int64_t n, int32_t n_pol, int64_t batch, int32_t dtype, int32_t device,
                            bool with_stash);
static int chirp_plan_create(ssfm_plan_t* out, int64_t n, int32_t n_pol, int64_t batch, int32_t dtype, int32_t device);
template <typename R>
static int chirp_propagate_t(ssfm_plan_t pl, void* field, const ssfm_fiber_params& prm, long long max_steps, int resume,
                             cudaStream_t st);
template <typename R>
static int chirp_apply_transfer_t(ssfm_plan_t pl, void* field, const void* h_dev, cudaStream_t st);

namespace {

template <typename R, int M>
int launch_col_fwd(const Params<R>& p, int nblocks, cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    constexpr int T = col_tile_of<R>(M);
    constexpr int E = points_per_thread<R>::value;
    const size_t smem = sizeof(C) * (size_t)(M * T + fft_plan<M, E>::table_size + SC_N);
    static bool attr_dev[kMaxDevices] = {false};     // function attributes are per device
    bool& attr = attr_dev[current_device()];
    if (!attr) { CU_TRY(cudaFuncSetAttribute(k_col_fwd<R, M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    k_col_fwd<R, M, T><<<nblocks, T * (M / E), smem, st>>>(p);
    ++ssfm_launches;
    return SSFM_OK;
}
template <typename R, int M>
int launch_col_inv(const Params<R>& p, int nblocks, cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    constexpr int T = col_tile_of<R>(M);
    constexpr int E = points_per_thread<R>::value;
    const size_t smem = sizeof(C) * (size_t)(M * T + fft_plan<M, E>::table_size + SC_N) + sizeof(R) * (size_t)(M * T);
    static bool attr_dev[kMaxDevices] = {false};     // function attributes are per device
    bool& attr = attr_dev[current_device()];
    if (!attr) { CU_TRY(cudaFuncSetAttribute(k_col_inv<R, M, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    k_col_inv<R, M, T><<<nblocks, T * (M / E), smem, st>>>(p);
    ++ssfm_launches;
    return SSFM_OK;
}
template <typename R, int M>
size_t col_mid_smem() {
    typedef typename cx_of<R>::type C;
    constexpr int T = col_tile_of<R>(M);
    constexpr int E = points_per_thread<R>::value;
    return sizeof(C) * (size_t)(M * T + (col_mid_tables_in_smem(M) ? fft_plan<M, E>::table_size : 0) + SC_N) +
           sizeof(R) * (size_t)(M * T);
}
template <typename R, int M, int SYNC>
int launch_col_mid(const Params<R>& p, int nblocks, int cluster, cudaStream_t st) {
    constexpr int T = col_tile_of<R>(M);
    constexpr int E = points_per_thread<R>::value;
    const size_t smem = col_mid_smem<R, M>();
    static bool attr_dev[kMaxDevices] = {false};     // function attributes are per device
    bool& attr = attr_dev[current_device()];
    if (!attr) {
        CU_TRY(cudaFuncSetAttribute(k_col_mid<R, M, T, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (SYNC == SYNC_CLUSTER)
            CU_TRY(cudaFuncSetAttribute(k_col_mid<R, M, T, SYNC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        attr = true;
    }
    if (SYNC == SYNC_CLUSTER) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)nblocks); cfg.blockDim = dim3(T * (M / E));
        cfg.dynamicSmemBytes = smem; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CU_TRY(cudaLaunchKernelEx(&cfg, k_col_mid<R, M, T, SYNC>, p));
    } else {
        k_col_mid<R, M, T, SYNC><<<nblocks, T * (M / E), smem, st>>>(p);
    }
    ++ssfm_launches;
    return SSFM_OK;
}
// How can the fused kernel synchronise the `group` tiles of one waveform on this device?
//   want = SYNC_CLUSTER : a cluster of `group` CTAs if it can be scheduled (group <= 16)
//   want = SYNC_GLOBAL / SYNC_LL : that protocol if `group` CTAs are resident at once
//   *mode = -1 when nothing fits (use the unfused schedule).
template <typename R, int M>
int col_mid_sync_mode(int num_sms, long long group, int want, int* mode) {
    constexpr int T = col_tile_of<R>(M);
    constexpr int E = points_per_thread<R>::value;
    const size_t smem = col_mid_smem<R, M>();
    *mode = -1;
    if (want == SYNC_CLUSTER && group <= 16) {
        CU_TRY(cudaFuncSetAttribute(k_col_mid<R, M, T, SYNC_CLUSTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(cudaFuncSetAttribute(k_col_mid<R, M, T, SYNC_CLUSTER>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)(group * 64)); cfg.blockDim = dim3(T * (M / E)); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)group; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int nclusters = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_col_mid<R, M, T, SYNC_CLUSTER>, &cfg);
        if (e == cudaSuccess && nclusters > 0) { *mode = SYNC_CLUSTER; return SSFM_OK; }
        (void)cudaGetLastError();
    }
    int per_sm = 0;
    if (want == SYNC_GLOBAL) {
        CU_TRY(cudaFuncSetAttribute(k_col_mid<R, M, T, SYNC_GLOBAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_col_mid<R, M, T, SYNC_GLOBAL>, T * (M / E), smem));
        if (group <= (long long)per_sm * num_sms) *mode = SYNC_GLOBAL;
    } else {
        CU_TRY(cudaFuncSetAttribute(k_col_mid<R, M, T, SYNC_LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_col_mid<R, M, T, SYNC_LL>, T * (M / E), smem));
        if (group <= (long long)per_sm * num_sms) *mode = SYNC_LL;
        if (getenv("SSFM_DEBUG")) fprintf(stderr, "[ssfm] col_mid LL: M=%d T=%d smem=%zu per_sm=%d group=%lld mode=%d\n", M, T, smem, per_sm, group, *mode);
    }
    return SSFM_OK;
}

template <typename R, int M>
int launch_row(const Params<R>& p, int nblocks, cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    constexpr int G = row_group_of<R>(M);
    constexpr int E = points_per_thread<R>::value;
    const size_t smem = sizeof(C) * (size_t)(G * RowExchange<M, E>::size + fft_plan<M, E>::table_size + SC_N);
    static bool attr_dev[kMaxDevices] = {false};     // function attributes are per device
    bool& attr = attr_dev[current_device()];
    if (!attr) { CU_TRY(cudaFuncSetAttribute(k_row<R, M, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; }
    k_row<R, M, G><<<nblocks, G * (M / E), smem, st>>>(p);
    ++ssfm_launches;
    return SSFM_OK;
}

#define SSFM_FOR_M(X) X(16) X(32) X(64) X(128) X(256) X(512) X(1024) X(2048)

enum ColKind { COL_FWD, COL_INV, COL_MID };

template <typename R>
int enqueue_col(const Params<R>& p, ColKind kind, cudaStream_t st, int sync = SYNC_GLOBAL) {
    const long long rows = (long long)p.batch * p.n_pol;
    switch (p.n1) {
#define X(M) case M: {                                                             \
            const int tiles = p.n2 / col_tile_of<R>(M);                                 \
            const int nb = (int)(rows * tiles);                                    \
            if (kind == COL_FWD) return launch_col_fwd<R, M>(p, nb, st);           \
            if (kind == COL_INV) return launch_col_inv<R, M>(p, nb, st);           \
            if (sync == SYNC_FIXED) return launch_col_mid<R, M, SYNC_FIXED>(p, nb, 1, st);              \
            if (sync == SYNC_CLUSTER) return launch_col_mid<R, M, SYNC_CLUSTER>(p, nb, tiles * p.n_pol, st); \
            if (sync == SYNC_LL) return launch_col_mid<R, M, SYNC_LL>(p, nb, 1, st);                    \
            return launch_col_mid<R, M, SYNC_GLOBAL>(p, nb, 1, st); }
        SSFM_FOR_M(X)
#undef X
        default: return fail(SSFM_ERR_UNSUPPORTED, "unsupported column transform size");
    }
}

template <typename R>
int enqueue_row(const Params<R>& p, cudaStream_t st) {
    const long long rows = (long long)p.batch * p.n_pol;
    switch (p.n2) {
#define X(M) case M: return launch_row<R, M>(p, (int)(rows * p.n1 / row_group_of<R>(M)), st);
        SSFM_FOR_M(X)
#undef X
        default: return fail(SSFM_ERR_UNSUPPORTED, "unsupported row transform size");
    }
}

template <typename R>
int fused_sync_mode(int n1, int num_sms, long long group, int want, int* mode) {
    switch (n1) {
#define X(M) case M: return col_mid_sync_mode<R, M>(num_sms, group, want, mode);
        SSFM_FOR_M(X)
#undef X
        default: return fail(SSFM_ERR_UNSUPPORTED, "unsupported column transform size");
    }
}

template <typename R>
Params<R> base_params(ssfm_plan_t pl, const ssfm_fiber_params& prm, bool& fixed, bool& single) {
    typedef typename cx_of<R>::type C;
    // scalar casts exactly as devices.py:1137-1142 (division in double first, then rounded to R)
    const R a_lin = (R)(prm.alpha_db_km / 4.343);
    const R b2 = (R)prm.beta2_ps2_km, b3 = (R)prm.beta3_ps3_km, g = (R)prm.gamma_w_km;
    const R L = (R)prm.length_km, pm = (R)prm.phi_max_rad;
    fixed = !std::isnan(prm.h_km);
    single = !fixed && ((b2 == (R)0 && b3 == (R)0) || g == (R)0);

    Params<R> base;
    std::memset(&base, 0, sizeof(base));
    base.tw_col = (const C*)pl->tw_col; base.tw_row = (const C*)pl->tw_row;
    base.tw_lo = (const C*)pl->tw_lo;   base.tw_hi = (const C*)pl->tw_hi;
    // four-step twiddles: auto = recurrence for complex128 (measured on B200, config #3: 5.34e10 against 5.23e10 with the full
    // table and 4.75e10 with two tables; the error stays ~1e-14), full table for complex64 (the recurrence doubles its error)
    const int tw_mode = pl->use_tw_full < 0 ? (sizeof(R) == 8 ? 2 : 1) : pl->use_tw_full;
    base.tw_full = tw_mode == 1 ? (const C*)pl->tw_full : nullptr;
    base.tw_chain = tw_mode == 2 ? 1 : 0;
    // separable linear operator of k_wf (complex128, beta_3 = 0, rows of 256 bins): plan option "lin_sep", -1 = auto (on)
    base.lin_sep = (sizeof(R) == 8 && pl->lin_sep != 0 && b3 == (R)0 && pl->n2 == 256 && !pl->long_n) ? 1 : 0;
    base.small_phase = (!fixed && !single && pm <= (R)0.05 && pm >= (R)0) ? 1 : 0;   // |Kerr phase| <= phi_max in adaptive mode
    base.lo_bits = pl->lo_bits ? pl->lo_bits : ilog2(pl->n2);
    base.n = (int)pl->n; base.n1 = pl->n1; base.n2 = pl->n2; base.log2_n2 = ilog2(pl->n2);
    base.n_glob = (int)pl->n; base.bin_mul = 1;
    base.n_pol = pl->n_pol;
    base.hlog_cap = pl->hlog_cap;
    base.adaptive = fixed ? 0 : 1;
    base.has_nl = (g != (R)0) ? 1 : 0;
    base.max_steps = 1 << 30;
    base.debug = pl->debug;
    base.l2_ahead = pl->l2_ahead;
    base.gamma = g; base.abs_gamma = std::fabs(g); base.phi_max = pm; base.length = L;
    base.att_half = -a_lin / (R)2;
    base.c2 = (R)0.5 * b2;                       // imag(1j/2 * beta_2): exact scaling
    base.c3 = (R)(1.0 / 6.0) * b3;               // imag(1j/6 * beta_3): R(1/6) times beta_3, rounded once
    base.wscale = ((1.0 / ((double)pl->n * prm.dt_s)) * 2.0) * 3.141592653589793 * 1e-12;
    base.inv_n = (R)1 / (R)pl->n;
    return base;
}

int ensure_stash(ssfm_plan_t pl) {
    if (pl->stash) return SSFM_OK;
    const size_t rsz = pl->dtype == SSFM_C64 ? 4 : 8;
    cudaError_t e = cudaMalloc(&pl->stash, (size_t)pl->batch * pl->n_pol * (size_t)pl->n * rsz);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return fail(SSFM_ERR_NOMEM, std::string("Kerr-phase stash: ") + cudaGetErrorString(e));
    }
    return SSFM_OK;
}

template <typename R>
int propagate_t(ssfm_plan_t pl, void* field, const ssfm_fiber_params& prm, long long max_steps, int resume,
                cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    bool fixed, single;
    const Params<R> base = base_params<R>(pl, prm, fixed, single);

    const long long B = pl->batch;
    const long long chunk = (pl->chunk > 0 && pl->chunk < B) ? pl->chunk : B;
    const size_t wf_elems = (size_t)pl->n_pol * (size_t)pl->n;
    const long long budget = max_steps > 0 ? max_steps : (1ll << 40);
    bool use_fused = pl->fused != 0;
    int sync = SYNC_FIXED;                                 // fixed step: no barrier in the fused kernel
    if (use_fused && !fixed) {                             // adaptive: the tiles of a waveform must synchronise
        const long long group = (long long)pl->n_pol * (pl->n2 / col_tile_rt(pl->n1, pl->dtype));
        // plan option "fused": 1 = LL words (default), 2 = atomics + spin, 3 = thread-block clusters (16-CTA clusters
        // schedule poorly: measured slower than either global-memory protocol, so they are opt-in only)
        const int want = pl->fused == 3 ? SYNC_CLUSTER : (pl->fused == 2 ? SYNC_GLOBAL : SYNC_LL);
        int rc = fused_sync_mode<R>(pl->n1, pl->num_sms, group, want, &sync);
        if (rc) return rc;
        if (sync < 0) use_fused = false;
    }
    if (pl->persistent && pl->wf_sync) {
        // One cooperative launch carries every waveform through all of its steps (ssfm_wf.cuh).
        Params<R> p = base;
        p.field = (C*)field; p.stash = (R*)pl->stash; p.ctrl = pl->ctrl; p.active = pl->active;
        p.ticket = pl->ticket; p.slots = pl->slots; p.hlog = pl->hlog; p.batch = (int)B;
        if (!pl->dim_tab) CU_TRY(cudaMalloc(&pl->dim_tab, sizeof(R) * (size_t)pl->n));
        k_fill_dim<R><<<(unsigned)((pl->n + 255) / 256), 256, 0, st>>>(p, (R*)pl->dim_tab);
        ++ssfm_launches;
        p.dim_tab = (const R*)pl->dim_tab;
        WfLaunch l{};
        l.sync_buf = pl->wf_sync; l.num_sms = pl->num_sms;
        l.fixed = fixed ? 1 : 0; l.single = single ? 1 : 0; l.resume = resume ? 1 : 0;
        l.h_fixed = fixed ? prm.h_km : 0.0;
        l.budget = budget; l.teams_cap = pl->teams_cap; l.placement = pl->placement; l.cluster = pl->cluster;
        l.ev0 = pl->wf_ev[0]; l.ev1 = pl->wf_ev[1];
        l.side = pl->wf_side; l.ev_side = pl->wf_ev_side;
        {   // waveforms of 32 .. 64 tiles without adaptive step control run as 16-CTA clusters with several tiles per CTA
            const long long units = (long long)pl->n_pol * (pl->n / 4096);
            if (p.has_nl && units >= 2 && units <= 64 && pl->n <= (1ll << 18) && pl->cluster != 0) {
                if (!pl->tstash) {                                  // (<= 16 tiles: the small clusters that fill the slots 16-CTA clusters leave)
                    const size_t bytes = (size_t)(units > 16 ? 16 : (getenv("SSFM_STASH_TEAMS") ? atoi(getenv("SSFM_STASH_TEAMS")) : 40)) * (size_t)units * 4096 * sizeof(R);
                    if (cudaMalloc(&pl->tstash, bytes) == cudaSuccess) pl->tstash_bytes = bytes;
                    else { (void)cudaGetLastError(); pl->tstash = nullptr; }
                }
                l.tstash = pl->tstash; l.tstash_bytes = pl->tstash_bytes;
            }
        }
        l.ready = pl->s_ready; l.done = pl->s_done; l.chunk_rows = (int)pl->s_chunk_rows;
        int teams = 0;
        const int rc = wf_propagate<R>(p, l, &teams, st);
        if (rc == SSFM_OK) {
            if (!pl->async_mode) CU_TRY(cudaStreamSynchronize(st));
            pl->have_state = true; pl->last = prm; pl->last_kind = 2; pl->last_teams = teams;
            return SSFM_OK;
        }
        if (rc != SSFM_ERR_UNSUPPORTED) return rc;
    }
    if (pl->s_ready) return SSFM_ERR_UNSUPPORTED;         // a streamed batch needs the persistent kernel (nothing has been enqueued)
    pl->last_kind = 1;
    { const int rs = ensure_stash(pl); if (rs) return rs; }
    int ci = 0;
    for (long long b0 = 0; b0 < B; b0 += chunk, ++ci) {
        const long long nb = (B - b0 < chunk) ? (B - b0) : chunk;
        Params<R> p = base;
        p.field = (C*)field + (size_t)b0 * wf_elems;
        p.stash = (R*)pl->stash + (size_t)b0 * wf_elems;
        p.ctrl = pl->ctrl + b0;
        p.active = pl->active + ci;
        p.ticket = pl->ticket;
        p.slots = pl->slots + (size_t)b0 * pl->n_pol * (size_t)(pl->n2 / col_tile_rt(pl->n1, pl->dtype)) * 2;
        p.hlog = pl->hlog ? pl->hlog + (size_t)b0 * pl->hlog_cap : nullptr;
        p.batch = (int)nb;

        if (!resume) {
            const int nb_i = (int)nb;
            CU_TRY(cudaMemcpyAsync(p.active, &nb_i, sizeof(int), cudaMemcpyHostToDevice, st));
            CU_TRY(cudaMemsetAsync(p.ctrl, 0, sizeof(Ctrl) * (size_t)nb, st));
            if (ci == 0) CU_TRY(cudaMemsetAsync(pl->slots, 0, pl->slots_bytes, st));     // step tags restart at 1
            if (!fixed && !single) {
                int per = (int)((wf_elems + 256 * 16 - 1) / (256 * 16));
                if (per > 64) per = 64;
                if (per < 1) per = 1;
                k_power_max<R><<<(unsigned)(nb * per), 256, 0, st>>>(p, per);
                ++ssfm_launches;
            }
            k_ctrl_init<R><<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(p, fixed ? 1 : 0, fixed ? (R)prm.h_km : (R)0,
                                                                      single ? 1 : 0);
            ++ssfm_launches;
            CU_TRY(cudaGetLastError());
        } else {
            // re-arm: waveforms that have not reached `length` continue; recount them on the host
            std::vector<Ctrl> h((size_t)nb);
            CU_TRY(cudaMemcpyAsync(h.data(), p.ctrl, sizeof(Ctrl) * (size_t)nb, cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaStreamSynchronize(st));
            int act = 0;
            for (auto& c : h) act += c.done ? 0 : 1;
            CU_TRY(cudaMemcpyAsync(p.active, &act, sizeof(int), cudaMemcpyHostToDevice, st));
            CU_TRY(cudaStreamSynchronize(st));
            if (act == 0) continue;
        }

        // One step = row kernel + column kernel(s).  Fused schedule (2R+2W per step):
        //   col_fwd | row, col_mid | row, col_mid | ...   (col_mid = end of step s + start of step s+1)
        // and the last budgeted step ends with col_inv so that the field is back in the time domain.
        // Unfused schedule (3R+3W): col_fwd, row, col_inv per step.
        long long enq = 0;
        int slot = 0;
        bool pending = false, finished = false;
        if (use_fused) { int rc = enqueue_col<R>(p, COL_FWD, st); if (rc) return rc; }
        while (!finished) {
            long long nsteps = pl->burst;
            if (enq + nsteps > budget) nsteps = budget - enq;
            for (long long s = 0; s < nsteps; ++s) {
                int rc = SSFM_OK;
                if (use_fused) {
                    rc = enqueue_row<R>(p, st);
                    if (!rc) rc = enqueue_col<R>(p, (enq + s + 1 == budget) ? COL_INV : COL_MID, st, sync);
                } else {
                    rc = enqueue_col<R>(p, COL_FWD, st);
                    if (!rc) rc = enqueue_row<R>(p, st);
                    if (!rc) rc = enqueue_col<R>(p, COL_INV, st);
                }
                if (rc) return rc;
            }
            enq += nsteps;
            CU_TRY(cudaGetLastError());
            CU_TRY(cudaMemcpyAsync(&pl->active_host[slot], p.active, sizeof(int), cudaMemcpyDeviceToHost, st));
            CU_TRY(cudaEventRecord(pl->ev[slot], st));
            if (pending) {
                CU_TRY(cudaEventSynchronize(pl->ev[slot ^ 1]));
                if (pl->active_host[slot ^ 1] == 0) finished = true;
            }
            pending = true;
            slot ^= 1;
            if (!finished && enq >= budget) {
                CU_TRY(cudaEventSynchronize(pl->ev[slot ^ 1]));
                finished = true;
            }
        }
    }
    CU_TRY(cudaStreamSynchronize(st));
    pl->have_state = true;
    pl->last = prm;
    return SSFM_OK;
}


// average device time of the three kernels of one step, CUDA events on the launching stream
template <typename R>
int time_kernels_t(ssfm_plan_t pl, void* field, const ssfm_fiber_params& prm_in, int reps, float* ms3, cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    ssfm_fiber_params prm = prm_in;
    if (std::isnan(prm.h_km)) prm.h_km = 1e-3;
    prm.length_km = 1e30;                                 // never finishes: every launch does full work
    bool fixed, single;
    { const int rs = ensure_stash(pl); if (rs) return rs; }
    Params<R> p = base_params<R>(pl, prm, fixed, single);
    p.field = (C*)field; p.stash = (R*)pl->stash; p.ctrl = pl->ctrl; p.active = pl->active; p.hlog = nullptr;
    p.batch = (int)pl->batch;
    const int nb = (int)pl->batch;
    CU_TRY(cudaMemcpyAsync(p.active, &nb, sizeof(int), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(p.ctrl, 0, sizeof(Ctrl) * (size_t)nb, st));
    k_ctrl_init<R><<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(p, 1, (R)prm.h_km, 0);
    p.ticket = pl->ticket;
    p.slots = pl->slots;
    CU_TRY(cudaMemsetAsync(pl->slots, 0, pl->slots_bytes, st));
    bool use_fused = pl->fused != 0;
    int sync = SYNC_FIXED;
    if (use_fused && prm_in.phi_max_rad >= 0) {            // time the adaptive-mode kernel (same work, plus the barrier)
        const long long group = (long long)pl->n_pol * (pl->n2 / col_tile_rt(pl->n1, pl->dtype));
        const int want = pl->fused == 3 ? SYNC_CLUSTER : (pl->fused == 2 ? SYNC_GLOBAL : SYNC_LL);
        int rc0 = fused_sync_mode<R>(pl->n1, pl->num_sms, group, want, &sync);
        if (rc0) return rc0;
        if (sync < 0) use_fused = false;
        p.adaptive = 1;                                     // controller follows phi_max / max; length is 1e30
    }
    cudaEvent_t ev[4];
    for (auto& e : ev) CU_TRY(cudaEventCreate(&e));
    double acc[3] = {0, 0, 0};
    int rc = SSFM_OK;
    if (use_fused) {   // steady state: row + fused column kernel; the opening col_fwd is timed once
        CU_TRY(cudaEventRecord(ev[0], st));
        rc = enqueue_col<R>(p, COL_FWD, st);
        CU_TRY(cudaEventRecord(ev[1], st));
        CU_TRY(cudaEventSynchronize(ev[1]));
        float ms0 = 0; CU_TRY(cudaEventElapsedTime(&ms0, ev[0], ev[1]));
        acc[0] = (double)ms0 * (reps > 0 ? reps : 1);
    }
    for (int r = -2; r < reps && !rc; ++r) {                 // two untimed warm-up steps
        CU_TRY(cudaEventRecord(ev[0], st));
        if (!use_fused) rc = enqueue_col<R>(p, COL_FWD, st);
        CU_TRY(cudaEventRecord(ev[1], st));
        if (!rc) rc = enqueue_row<R>(p, st);
        CU_TRY(cudaEventRecord(ev[2], st));
        if (!rc) rc = enqueue_col<R>(p, use_fused ? COL_MID : COL_INV, st, sync);
        CU_TRY(cudaEventRecord(ev[3], st));
        CU_TRY(cudaEventSynchronize(ev[3]));
        if (r >= 0)
            for (int k = use_fused ? 1 : 0; k < 3; ++k) { float ms = 0; CU_TRY(cudaEventElapsedTime(&ms, ev[k], ev[k + 1])); acc[k] += ms; }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    for (int k = 0; k < 3; ++k) ms3[k] = (float)(acc[k] / (reps > 0 ? reps : 1));
    pl->have_state = false;
    return rc;
}


// ---- transfer functions: y = IFFT(H * FFT(y)) per row, spectrum kept in transposed order ------------
// |H(e^{jw})|^2 of a biquad cascade, written straight into the transposed layout (bin k1 + N1*k2 at [k1][k2])
template <typename R>
__global__ void k_fill_xfer_sos(typename cx_of<R>::type* out, int n, int n1, int n2, ssfm_filt::Sos f) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    const int k1 = pos / n2, k2 = pos % n2;
    const int k = k1 + n1 * k2;
    double s, c;
    sincospi(2.0 * (double)k / (double)n, &s, &c);             // z^-1 = (c, -s), z^-2 = (c2, -s2)
    const double c2 = c * c - s * s, s2 = 2.0 * s * c;
    double g = 1.0;
    for (int i = 0; i < f.n_sections; ++i) {
        const double* q = f.c[i];
        const double nr = q[0] + q[1] * c + q[2] * c2, ni = -(q[1] * s + q[2] * s2);
        const double dr = q[3] + q[4] * c + q[5] * c2, di = -(q[4] * s + q[5] * s2);
        g *= (nr * nr + ni * ni) / (dr * dr + di * di);
    }
    out[pos] = mk<R>((R)g, (R)0);
}
// user table H[k] in natural bin order -> transposed layout
template <typename R>
__global__ void k_transpose_xfer(typename cx_of<R>::type* out, const typename cx_of<R>::type* in, int n, int n1, int n2) {
    const int pos = blockIdx.x * blockDim.x + threadIdx.x;
    if (pos >= n) return;
    out[pos] = in[(pos / n2) + n1 * (pos % n2)];
}

template <typename R>
Params<R> transfer_params(ssfm_plan_t pl, void* field, long long rows) {
    typedef typename cx_of<R>::type C;
    ssfm_fiber_params prm{};
    prm.dt_s = 1.0; prm.length_km = 1e30; prm.phi_max_rad = 0.01; prm.h_km = 1.0;   // linear, one "step" per application
    bool fixed, single;
    Params<R> p = base_params<R>(pl, prm, fixed, single);
    p.field = (C*)field; p.stash = nullptr; p.ctrl = pl->ctrl; p.active = pl->active; p.hlog = nullptr;
    p.ticket = pl->ticket; p.batch = (int)rows; p.xfer = (const C*)pl->xfer;
    return p;
}
// controller records of a transfer plan: fixed step 1 towards an unreachable length, so every application is "one more step"
// of every row and nothing has to be re-armed between applications (the filter pipeline applies the plan once per chunk)
template <typename R>
int transfer_arm(ssfm_plan_t pl, cudaStream_t st) {
    Params<R> p = transfer_params<R>(pl, nullptr, pl->batch);
    const int nb = (int)pl->batch;
    CU_TRY(cudaMemcpyAsync(p.active, &nb, sizeof(int), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(p.ctrl, 0, sizeof(Ctrl) * (size_t)nb, st));
    k_ctrl_init<R><<<(unsigned)((nb + 127) / 128), 128, 0, st>>>(p, 1, (R)1, 0);
    ++ssfm_launches;
    CU_TRY(cudaGetLastError());
    return SSFM_OK;
}
// y = IFFT(H FFT(x)) for `rows` rows: ONE launch of the persistent kernel when the geometry has one (2^12 .. 2^20 samples) --
// teams carry a row through forward column transforms, row transforms x H x inverse row transforms and inverse column
// transforms with the rows in flight L2-resident, so HBM sees one read (from `src` when given: out of place) and one write per
// row -- else the three streaming kernels (then `src` is copied into `field` first).  Nothing is synchronised.
template <typename R>
int apply_transfer_rows(ssfm_plan_t pl, void* field, long long rows, cudaStream_t st, bool arm = true, const void* src = nullptr,
                        bool one_launch = true) {
    typedef typename cx_of<R>::type C;
    if (one_launch && pl->persistent && pl->wf_sync && pl->n >= 4096 && pl->n <= (1ll << 20) && !getenv("SSFM_TRANSFER_MULTILAUNCH")) {
        ssfm_fiber_params prm{};
        prm.dt_s = 1.0; prm.length_km = 1.0; prm.phi_max_rad = 0.01; prm.h_km = 1.0;   // linear, exactly one "step" of length 1
        bool fixed, single;
        Params<R> p = base_params<R>(pl, prm, fixed, single);
        p.field = (C*)field; p.field_in = (src && src != field) ? (const C*)src : nullptr;
        p.stash = nullptr; p.ctrl = pl->ctrl; p.active = pl->active; p.hlog = nullptr;
        p.ticket = pl->ticket; p.slots = pl->slots; p.batch = (int)rows; p.xfer = (const C*)pl->xfer;
        WfLaunch l{};
        l.sync_buf = pl->wf_sync; l.num_sms = pl->num_sms;
        l.fixed = 1; l.single = 0; l.resume = 0; l.h_fixed = 1.0;
        l.budget = 1; l.teams_cap = pl->teams_cap; l.placement = pl->placement; l.cluster = pl->cluster;
        l.ev0 = pl->wf_ev[0]; l.ev1 = pl->wf_ev[1];
        l.side = pl->wf_side; l.ev_side = pl->wf_ev_side;
        int teams = 0;
        const int rc = wf_propagate<R>(p, l, &teams, st);
        if (rc == SSFM_OK) { pl->have_state = false; pl->last_kind = 2; pl->last_teams = teams; return SSFM_OK; }
        if (rc != SSFM_ERR_UNSUPPORTED) return rc;
    }
    if (src && src != field)
        CU_TRY(cudaMemcpyAsync(field, src, sizeof(C) * (size_t)rows * (size_t)pl->n, cudaMemcpyDeviceToDevice, st));
    if (arm) { const int ra = transfer_arm<R>(pl, st); if (ra) return ra; }
    const Params<R> p = transfer_params<R>(pl, field, rows);
    int rc = enqueue_col<R>(p, COL_FWD, st);
    if (!rc) rc = enqueue_row<R>(p, st);
    if (!rc) rc = enqueue_col<R>(p, COL_INV, st);
    if (rc) return rc;
    CU_TRY(cudaGetLastError());
    pl->have_state = false;
    pl->last_kind = 1;
    return SSFM_OK;
}
template <typename R>
int apply_transfer_t(ssfm_plan_t pl, void* field, cudaStream_t st) {
    const bool big = pl->batch * pl->n_pol * pl->n >= (1ll << 25);          // small batches: the three streaming kernels are faster
    return apply_transfer_rows<R>(pl, field, pl->batch, st, true, nullptr, big);
}

int ensure_xfer(ssfm_plan_t pl) {
    if (pl->xfer) return SSFM_OK;
    const size_t csz = pl->dtype == SSFM_C64 ? 8 : 16;
    CU_TRY(cudaMalloc(&pl->xfer, csz * (size_t)pl->n));
    return SSFM_OK;
}

std::mutex g_filter_mu;
std::map<std::tuple<int, long long, long long>, ssfm_plan_t> g_filter_plans;   // (device, n, rows) -> plan without stash

}  // namespace

int ssfm_internal_pass_tables_f64(void** out, int M, cudaStream_t st) { return build_pass_tables<double>(out, M, st); }

int ssfm_internal_transfer_prepare(int device, long long n, long long max_rows, const ssfm_filt::Sos& f, void** plan_out,
                                   cudaStream_t st) {
    std::lock_guard<std::mutex> lock(g_filter_mu);
    const auto key = std::make_tuple(device, n, max_rows);
    ssfm_plan_t pl = nullptr;
    auto it = g_filter_plans.find(key);
    if (it == g_filter_plans.end()) {
        if (g_filter_plans.size() >= 8) {                        // plans are small (tables only) but keep the map bounded
            ssfm_plan_destroy(g_filter_plans.begin()->second);
            g_filter_plans.erase(g_filter_plans.begin());
        }
        int rc = plan_create_impl(&pl, n, 1, max_rows, SSFM_C128, device, false);
        if (rc) return rc;
        g_filter_plans[key] = pl;
    } else {
        pl = it->second;
    }
    CU_TRY(cudaSetDevice(device));
    int rc = ensure_xfer(pl);
    if (rc) return rc;
    k_fill_xfer_sos<double><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((double2*)pl->xfer, (int)n, pl->n1, pl->n2, f);
    ++ssfm_launches;
    rc = transfer_arm<double>(pl, st);
    if (rc) return rc;
    *plan_out = pl;
    return SSFM_OK;
}

int ssfm_internal_transfer_apply(void* plan, void* y_dev, long long rows, cudaStream_t st, const void* src_dev, int one_launch) {
    ssfm_plan_t pl = (ssfm_plan_t)plan;
    if (!pl || rows < 1 || rows > pl->batch) return fail(SSFM_ERR_INVALID, "transfer_apply: bad plan or row count");
    return apply_transfer_rows<double>(pl, y_dev, rows, st, false, src_dev, one_launch != 0);
}

extern "C" {

int ssfm_apply_transfer(ssfm_plan_t pl, void* field, const void* h_dev, void* stream) {
    if (!pl || !field || !h_dev) return fail(SSFM_ERR_INVALID, "null plan, field or transfer function");
    if (pl->long_n) return fail(SSFM_ERR_UNSUPPORTED, "transfer functions are not available for long-waveform plans");
    if (pl->chirp_m) {
        CU_TRY(cudaSetDevice(pl->device));
        const int rc = pl->dtype == SSFM_C64 ? chirp_apply_transfer_t<float>(pl, field, h_dev, (cudaStream_t)stream)
                                             : chirp_apply_transfer_t<double>(pl, field, h_dev, (cudaStream_t)stream);
        if (rc) return rc;
        CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
        return SSFM_OK;
    }
    CU_TRY(cudaSetDevice(pl->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = ensure_xfer(pl);
    if (rc) return rc;
    const unsigned nb = (unsigned)((pl->n + 255) / 256);
    if (pl->dtype == SSFM_C64) {
        k_transpose_xfer<float><<<nb, 256, 0, st>>>((float2*)pl->xfer, (const float2*)h_dev, (int)pl->n, pl->n1, pl->n2);
        ++ssfm_launches;
        rc = apply_transfer_t<float>(pl, field, st);
    } else {
        k_transpose_xfer<double><<<nb, 256, 0, st>>>((double2*)pl->xfer, (const double2*)h_dev, (int)pl->n, pl->n1, pl->n2);
        ++ssfm_launches;
        rc = apply_transfer_t<double>(pl, field, st);
    }
    if (rc) return rc;
    CU_TRY(cudaStreamSynchronize(st));
    return SSFM_OK;
}

int64_t ssfm_launch_count(void) { return ssfm_launches; }

int ssfm_time_step_kernels(ssfm_plan_t pl, void* field, const ssfm_fiber_params* prm, int32_t reps, float* ms3,
                           void* stream) {
    if (!pl || !field || !prm || !ms3 || reps < 1) return fail(SSFM_ERR_INVALID, "null argument or reps < 1");
    if (pl->chirp_m || pl->long_n) return fail(SSFM_ERR_UNSUPPORTED, "per-kernel timing needs a power-of-two plan of at most 2^22 samples");
    CU_TRY(cudaSetDevice(pl->device));
    if (pl->dtype == SSFM_C64) return time_kernels_t<float>(pl, field, *prm, reps, ms3, (cudaStream_t)stream);
    return time_kernels_t<double>(pl, field, *prm, reps, ms3, (cudaStream_t)stream);
}

int ssfm_abi_version(void) { return SSFM_ABI_VERSION; }
const char* ssfm_last_error(void) { return g_err.c_str(); }

int ssfm_plan_create(ssfm_plan_t* out, int64_t n, int32_t n_pol, int64_t batch, int32_t dtype, int32_t device) {
    return plan_create_impl(out, n, n_pol, batch, dtype, device, true);
}

}  // extern "C"

static int plan_create_impl(ssfm_plan_t* out, int64_t n, int32_t n_pol, int64_t batch, int32_t dtype, int32_t device,
                            bool with_stash) {
    if (!out) return fail(SSFM_ERR_INVALID, "plan pointer is null");
    *out = nullptr;
    if (n_pol != 1 && n_pol != 2) return fail(SSFM_ERR_INVALID, "n_pol must be either 1 or 2");
    if (batch < 1) return fail(SSFM_ERR_INVALID, "n_waveforms must be >= 1");
    if (dtype != SSFM_C64 && dtype != SSFM_C128) return fail(SSFM_ERR_INVALID, "dtype must be SSFM_C64 or SSFM_C128");
    if (with_stash && n >= 2 && n <= (1ll << 21) && (n < 256 || (n & (n - 1))))
        return chirp_plan_create(out, n, n_pol, batch, dtype, device);      // any other length: chirp-z transforms
    if (n < 256 || n > (1ll << 22) || (n & (n - 1)))
        return fail(SSFM_ERR_UNSUPPORTED, "n_samples must be in [2, 2^21], or a power of two up to 2^22 "
                                          "(longer power-of-two waveforms: ssfm_long_plan_create)");
    CU_TRY(cudaSetDevice(device));
    ssfm_plan_s* pl = new ssfm_plan_s();
    pl->device = device; pl->dtype = dtype; pl->n_pol = n_pol; pl->n = n; pl->batch = batch;
    pl->log2n = ilog2(n);
    pl->n1 = 1 << (pl->log2n / 2);
    pl->n2 = (int)(n / pl->n1);
    pl->hlog_cap = 4096;
    if ((size_t)batch * pl->hlog_cap * sizeof(double) > (256u << 20)) pl->hlog_cap = (int)((256u << 20) / (batch * sizeof(double)));
    if (pl->hlog_cap < 16) pl->hlog_cap = 16;

    const size_t rsz = dtype == SSFM_C64 ? 4 : 8, csz = 2 * rsz;
    const size_t elems = (size_t)batch * n_pol * n;
    cudaError_t e;
    pl->propagates = with_stash;
    e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->ctrl, sizeof(Ctrl) * (size_t)batch);
    if (e == cudaSuccess) { pl->n_active = (int)batch; e = cudaMalloc((void**)&pl->active, sizeof(int) * (size_t)batch); }
    pl->slots_bytes = sizeof(unsigned long long) * 2 * (size_t)batch * n_pol * (size_t)(pl->n2 / col_tile_rt(pl->n1, pl->dtype));
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->slots, pl->slots_bytes);
    if (e == cudaSuccess) e = cudaMemset(pl->slots, 0, pl->slots_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->ticket, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(pl->ticket, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->hlog, sizeof(double) * (size_t)batch * pl->hlog_cap);
    if (e == cudaSuccess) e = cudaMalloc(&pl->tw_lo, csz * (size_t)pl->n2);
    if (e == cudaSuccess) e = cudaMalloc(&pl->tw_hi, csz * (size_t)pl->n1);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&pl->active_host, 2 * sizeof(int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev[1], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&pl->wf_sync, WF_SYNC_BYTES);          // (transfer plans run through k_wf too)
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&pl->wf_side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->wf_ev_side, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&pl->wf_ev[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&pl->wf_ev[1]);
    if (e != cudaSuccess) {
        ssfm_plan_destroy(pl);
        return fail(e == cudaErrorMemoryAllocation ? SSFM_ERR_NOMEM : SSFM_ERR_CUDA,
                    std::string("plan allocation: ") + cudaGetErrorString(e));
    }
    CU_TRY(cudaMemset(pl->ctrl, 0, sizeof(Ctrl) * (size_t)batch));
    CU_TRY(cudaMemset(pl->hlog, 0, sizeof(double) * (size_t)batch * pl->hlog_cap));
    int rc;
    if (dtype == SSFM_C64) {
        k_build_unit_roots<float><<<(pl->n2 + 127) / 128, 128>>>((float2*)pl->tw_lo, pl->n2, 1, n);
        k_build_unit_roots<float><<<(pl->n1 + 127) / 128, 128>>>((float2*)pl->tw_hi, pl->n1, pl->n2, n);
        rc = build_pass_tables<float>(&pl->tw_col, pl->n1, 0);
        if (!rc) rc = build_pass_tables<float>(&pl->tw_row, pl->n2, 0);
    } else {
        k_build_unit_roots<double><<<(pl->n2 + 127) / 128, 128>>>((double2*)pl->tw_lo, pl->n2, 1, n);
        k_build_unit_roots<double><<<(pl->n1 + 127) / 128, 128>>>((double2*)pl->tw_hi, pl->n1, pl->n2, n);
        rc = build_pass_tables<double>(&pl->tw_col, pl->n1, 0);
        if (!rc) rc = build_pass_tables<double>(&pl->tw_row, pl->n2, 0);
    }
    if (!rc && n <= (1ll << 20)) {   // full four-step table W_N^{n2*k1} (16 MiB of complex128 at N = 2^20; it lives in L2)
        cudaError_t ef = cudaMalloc(&pl->tw_full, csz * (size_t)n);
        if (ef != cudaSuccess) { (void)cudaGetLastError(); pl->tw_full = nullptr; }
        else if (dtype == SSFM_C64) k_build_fourstep<float><<<(unsigned)((n + 255) / 256), 256>>>((float2*)pl->tw_full, (int)n, pl->n2);
        else k_build_fourstep<double><<<(unsigned)((n + 255) / 256), 256>>>((double2*)pl->tw_full, (int)n, pl->n2);
    }
    if (rc) { ssfm_plan_destroy(pl); return rc; }
    cudaError_t es = cudaDeviceSynchronize();
    if (es != cudaSuccess) { ssfm_plan_destroy(pl); return fail(SSFM_ERR_CUDA, std::string("table build: ") + cudaGetErrorString(es)); }
    *out = pl;
    return SSFM_OK;
}

extern "C" {

int ssfm_plan_destroy(ssfm_plan_t pl) {
    if (!pl) return SSFM_OK;
    cudaSetDevice(pl->device);
    cudaFree(pl->stash); cudaFree(pl->xfer); cudaFree(pl->ctrl); cudaFree(pl->active); cudaFree(pl->hlog); cudaFree(pl->ticket); cudaFree(pl->slots);
    cudaFree(pl->tw_col); cudaFree(pl->tw_row); cudaFree(pl->tw_lo); cudaFree(pl->tw_hi); cudaFree(pl->tw_full);
    if (pl->active_host) cudaFreeHost(pl->active_host);
    if (pl->ev[0]) cudaEventDestroy(pl->ev[0]);
    if (pl->ev[1]) cudaEventDestroy(pl->ev[1]);
    cudaFree(pl->wf_sync);
    cudaFree(pl->wb); cudaFree(pl->wtab); cudaFree(pl->xf_fwd); cudaFree(pl->xf_inv);
    for (int r = 0; r < 8; ++r)
        if (pl->peer_base[r] && pl->peer_base[r] != pl->xbuf) cudaIpcCloseMemHandle(pl->peer_base[r]);
    cudaFree(pl->xbuf); cudaFree(pl->d_peer_flags); cudaFree(pl->dim_tab); cudaFree(pl->tstash);
    if (pl->peek_stream) cudaStreamDestroy(pl->peek_stream);
    if (pl->peek_host) cudaFreeHost(pl->peek_host);
    if (pl->wf_side) cudaStreamDestroy(pl->wf_side);
    if (pl->wf_ev_side) cudaEventDestroy(pl->wf_ev_side);
    if (pl->inner) ssfm_plan_destroy(pl->inner);
    if (pl->wf_ev[0]) cudaEventDestroy(pl->wf_ev[0]);
    if (pl->wf_ev[1]) cudaEventDestroy(pl->wf_ev[1]);
    delete pl;
    return SSFM_OK;
}

int ssfm_plan_set_option(ssfm_plan_t pl, const char* name, int64_t value) {
    if (!pl || !name) return fail(SSFM_ERR_INVALID, "null plan or option name");
    const std::string k(name);
    if (k == "chunk_waveforms") { if (value < 0) return fail(SSFM_ERR_INVALID, "chunk_waveforms < 0"); pl->chunk = value; }
    else if (k == "burst_steps") { if (value < 1 || value > 4096) return fail(SSFM_ERR_INVALID, "burst_steps out of range"); pl->burst = (int)value; }
    else if (k == "fused") { pl->fused = (int)value; }   // 0 unfused, 1 fused (LL barrier), 2 fused (atomic barrier), 3 fused (cluster barrier when possible)
    else if (k == "debug") { pl->debug = (int)value; }
    else if (k == "persistent") { pl->persistent = value ? 1 : 0; }
    else if (k == "async") { pl->async_mode = value ? 1 : 0; }
    else if (k == "cluster") { pl->cluster = value < 0 ? -1 : (value ? 1 : 0); }
    else if (k == "placement") { pl->placement = value < 0 ? -1 : (value ? 1 : 0); }
    else if (k == "teams") { if (value < 0) return fail(SSFM_ERR_INVALID, "teams < 0"); pl->teams_cap = (int)value; }
    else if (k == "lin_sep") { pl->lin_sep = value < 0 ? -1 : (value ? 1 : 0); }
    else if (k == "tw_full") { pl->use_tw_full = value < 0 ? -1 : ((value == 2) ? 2 : (value ? 1 : 0)); }
    else if (k == "l2_ahead") { pl->l2_ahead = (int)value; }
    else return fail(SSFM_ERR_INVALID, "unknown option '" + k + "'");
    return SSFM_OK;
}

int ssfm_propagate(ssfm_plan_t pl, void* field, const ssfm_fiber_params* prm, int64_t max_steps, int32_t resume,
                   void* stream) {
    if (!pl || !field || !prm) return fail(SSFM_ERR_INVALID, "null plan, field or params");
    if (!pl->propagates && !pl->chirp_m) return fail(SSFM_ERR_INVALID, "this plan was created for transfer functions only");
    if (pl->long_n) return fail(SSFM_ERR_INVALID, "long-waveform plans are driven through ssfm_long_*");
    if (!(prm->dt_s > 0)) return fail(SSFM_ERR_INVALID, "dt_s must be > 0");
    if (max_steps < 0) return fail(SSFM_ERR_INVALID, "max_steps < 0");
    if (resume && !pl->have_state) return fail(SSFM_ERR_INVALID, "resume requested but the plan holds no controller state");
    CU_TRY(cudaSetDevice(pl->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (pl->chirp_m) {
        if (pl->dtype == SSFM_C64) return chirp_propagate_t<float>(pl, field, *prm, max_steps, resume, st);
        return chirp_propagate_t<double>(pl, field, *prm, max_steps, resume, st);
    }
    if (pl->dtype == SSFM_C64) return propagate_t<float>(pl, field, *prm, max_steps, resume, st);
    return propagate_t<double>(pl, field, *prm, max_steps, resume, st);
}

int ssfm_get_state(ssfm_plan_t pl, int32_t* steps, double* z, double* h_next, int32_t* done) {
    if (!pl) return fail(SSFM_ERR_INVALID, "null plan");
    CU_TRY(cudaSetDevice(pl->device));
    std::vector<Ctrl> h((size_t)pl->batch);
    CU_TRY(cudaMemcpy(h.data(), pl->ctrl, sizeof(Ctrl) * h.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < h.size(); ++i) {
        if (steps) steps[i] = h[i].steps;
        if (z) z[i] = h[i].z;
        if (h_next) h_next[i] = h[i].h;
        if (done) done[i] = h[i].done;
    }
    return SSFM_OK;
}

int ssfm_plan_get_option(ssfm_plan_t pl, const char* name, int64_t* value) {
    if (!pl || !name || !value) return fail(SSFM_ERR_INVALID, "null plan, option name or value");
    const std::string k(name);
    if (k == "hlog_cap") *value = pl->hlog_cap;
    else if (k == "chunk_waveforms") *value = pl->chunk;
    else if (k == "burst_steps") *value = pl->burst;
    else if (k == "fused") *value = pl->fused;
    else if (k == "persistent") *value = pl->persistent;
    else if (k == "cluster") *value = pl->cluster;
    else if (k == "placement") *value = pl->placement;
    else if (k == "teams") *value = pl->teams_cap;
    else if (k == "n1") *value = pl->n1;
    else if (k == "n2") *value = pl->n2;
    else return fail(SSFM_ERR_INVALID, "unknown option '" + k + "'");
    return SSFM_OK;
}

int ssfm_peek_state(ssfm_plan_t pl, int64_t row, int32_t* steps, double* z, int32_t* done) {
    if (!pl || row < 0 || row >= pl->batch) return fail(SSFM_ERR_INVALID, "null plan or row out of range");
    CU_TRY(cudaSetDevice(pl->device));
    if (!pl->peek_stream) {
        CU_TRY(cudaStreamCreateWithFlags(&pl->peek_stream, cudaStreamNonBlocking));
        CU_TRY(cudaHostAlloc((void**)&pl->peek_host, sizeof(Ctrl), cudaHostAllocDefault));
    }
    // a copy on its own stream: it does not wait for the propagation running on the caller's stream; the controller record is
    // rewritten once per step by one thread (z, h, steps, done: each an aligned word, so a torn read shows at worst the
    // previous step's value of one field -- good enough for a progress display, which is all this is for)
    CU_TRY(cudaMemcpyAsync(pl->peek_host, pl->ctrl + row, sizeof(Ctrl), cudaMemcpyDeviceToHost, pl->peek_stream));
    CU_TRY(cudaStreamSynchronize(pl->peek_stream));
    if (steps) *steps = pl->peek_host->steps;
    if (z) *z = pl->peek_host->z;
    if (done) *done = pl->peek_host->done;
    return SSFM_OK;
}

int ssfm_get_last_timing(ssfm_plan_t pl, int32_t* kind, int32_t* teams, float* kernel_ms) {
    if (!pl) return fail(SSFM_ERR_INVALID, "null plan");
    if (kind) *kind = pl->last_kind;
    if (teams) *teams = pl->last_kind == 2 ? pl->last_teams : 0;
    if (kernel_ms) {
        *kernel_ms = 0.0f;
        if (pl->last_kind == 2) {
            CU_TRY(cudaSetDevice(pl->device));
            CU_TRY(cudaEventSynchronize(pl->wf_ev[1]));
            CU_TRY(cudaEventElapsedTime(kernel_ms, pl->wf_ev[0], pl->wf_ev[1]));
        }
    }
    return SSFM_OK;
}

int ssfm_copy_state_async(ssfm_plan_t pl, void* dst_host, void* stream) {
    if (!pl || !dst_host) return fail(SSFM_ERR_INVALID, "null plan or buffer");
    CU_TRY(cudaSetDevice(pl->device));
    CU_TRY(cudaMemcpyAsync(dst_host, pl->ctrl, sizeof(Ctrl) * (size_t)pl->batch, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return SSFM_OK;
}

// ---- streamed batches: one persistent launch for a batch whose chunks are still on their way from the host ----------------
namespace {
typedef int (*stream_memop_fn)(void* /*CUstream*/, unsigned long long /*CUdeviceptr*/, unsigned int, unsigned int);
stream_memop_fn driver_entry(const char* name) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
        (void)cudaGetLastError();
        return nullptr;
    }
    return (stream_memop_fn)fn;
}
}  // namespace

int ssfm_stream_write_u32(void* stream, void* dev_ptr, uint32_t value) {
    static stream_memop_fn fn = driver_entry("cuStreamWriteValue32");
    if (!fn) return fail(SSFM_ERR_UNSUPPORTED, "cuStreamWriteValue32 is not available from this driver");
    if (!dev_ptr) return fail(SSFM_ERR_INVALID, "null pointer");
    const int rc = fn(stream, (unsigned long long)dev_ptr, value, 0u /* CU_STREAM_WRITE_VALUE_DEFAULT */);
    if (rc) return fail(SSFM_ERR_CUDA, "cuStreamWriteValue32 failed with driver error " + std::to_string(rc));
    return SSFM_OK;
}

int ssfm_stream_wait_geq_u32(void* stream, const void* dev_ptr, uint32_t value) {
    static stream_memop_fn fn = driver_entry("cuStreamWaitValue32");
    if (!fn) return fail(SSFM_ERR_UNSUPPORTED, "cuStreamWaitValue32 is not available from this driver");
    if (!dev_ptr) return fail(SSFM_ERR_INVALID, "null pointer");
    const int rc = fn(stream, (unsigned long long)dev_ptr, value, 0u /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (rc) return fail(SSFM_ERR_CUDA, "cuStreamWaitValue32 failed with driver error " + std::to_string(rc));
    return SSFM_OK;
}

int ssfm_propagate_streamed(ssfm_plan_t pl, void* field, const ssfm_fiber_params* prm, const uint32_t* ready_dev,
                            uint32_t* done_dev, int64_t chunk_rows, void* stream) {
    if (!pl || !field || !prm || !ready_dev || !done_dev) return fail(SSFM_ERR_INVALID, "null plan, field, params or counters");
    if (chunk_rows < 1) return fail(SSFM_ERR_INVALID, "chunk_rows must be >= 1");
    if (!pl->propagates || pl->chirp_m || pl->long_n || !pl->persistent || !pl->wf_sync)
        return SSFM_ERR_UNSUPPORTED;                       // only the persistent kernel adopts waveforms as they arrive
    pl->s_ready = ready_dev; pl->s_done = done_dev; pl->s_chunk_rows = chunk_rows;
    const int was_async = pl->async_mode;
    pl->async_mode = 1;                                    // the caller enqueues the copies AFTER this call returns
    const int rc = ssfm_propagate(pl, field, prm, 0, 0, stream);
    pl->async_mode = was_async;
    pl->s_ready = nullptr; pl->s_done = nullptr; pl->s_chunk_rows = 0;
    return rc;
}

int ssfm_get_step_log(ssfm_plan_t pl, double* out, int64_t cap) {
    if (!pl || !out || cap < 1) return fail(SSFM_ERR_INVALID, "null plan/buffer or cap < 1");
    CU_TRY(cudaSetDevice(pl->device));
    const int64_t w = cap < pl->hlog_cap ? cap : pl->hlog_cap;
    CU_TRY(cudaMemcpy2D(out, sizeof(double) * (size_t)cap, pl->hlog, sizeof(double) * (size_t)pl->hlog_cap,
                        sizeof(double) * (size_t)w, (size_t)pl->batch, cudaMemcpyDeviceToHost));
    return SSFM_OK;
}

int ssfm_fiber_host(ssfm_plan_t pl, const void* in, void* outp, const ssfm_fiber_params* prm, void* stream) {
    if (!pl || !in || !outp || !prm) return fail(SSFM_ERR_INVALID, "null plan, buffer or params");
    CU_TRY(cudaSetDevice(pl->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = (size_t)pl->batch * pl->n_pol * pl->n * (pl->dtype == SSFM_C64 ? 8 : 16);
    void* d = nullptr;
    CU_TRY(cudaMalloc(&d, bytes));
    cudaError_t e = cudaMemcpyAsync(d, in, bytes, cudaMemcpyHostToDevice, st);
    int rc = SSFM_OK;
    if (e != cudaSuccess) rc = fail(SSFM_ERR_CUDA, std::string("H2D: ") + cudaGetErrorString(e));
    if (!rc) rc = ssfm_propagate(pl, d, prm, 0, 0, stream);
    if (!rc) {
        e = cudaMemcpyAsync(outp, d, bytes, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) rc = fail(SSFM_ERR_CUDA, std::string("D2H: ") + cudaGetErrorString(e));
    }
    cudaFree(d);
    return rc;
}

}  // extern "C"

// =================================================================================================
// Long waveforms: N = N0 x N_l, beyond one two-pass transform (N > 2^22) and/or spread over several
// GPUs (BASELINE config #5).  Sample n = na N_l + nb, bin k = ka + N0 kb.  In the time domain rank g holds
// the columns nb in [g N_l/G, (g+1) N_l/G) of the N0 x N_l matrix for every na (local array [N0][N_l/G]);
// in the frequency domain it holds the rows ka in [g N0/G, (g+1) N0/G) for every nb (local array
// [N0/G][N_l]).  One split step is
//   OUTER (this plan; the column kernels of ssfm_kernels.cuh with M = N0 and the global twiddle W_N^{nb ka}):
//         end of the previous step / Kerr half steps / N0-point column transforms       -- local, 1R + 1W
//   exchange  [N0][N_l/G] -> [N0/G][N_l]   (all-to-all between the ranks; the identity for one rank)
//   INNER (pl->inner): N_l-point transforms of the rank's rows by the ordinary two-pass kernels with the
//         linear operator exp(D~(w_k) h) at global bin k = ka + N0 (k1 + N1 k2) in the middle     -- local, 3R + 3W
//   exchange back.
// The exchange itself belongs to the host side (torch.distributed all_to_all_single over NCCL/NVLink,
// opticomlib_b200/longwave.py): the library has no NCCL dependency.  The spectrum stays in (doubly)
// transposed order throughout; the linear operator is point-wise, so no transposition pass exists.
// =================================================================================================
namespace {

template <typename R>
Params<R> long_outer_params(ssfm_plan_t pl, void* field) {
    typedef typename cx_of<R>::type C;
    bool fixed, single;
    Params<R> p = base_params<R>(pl, pl->last, fixed, single);
    p.field = (C*)field; p.stash = (R*)pl->stash; p.ctrl = pl->ctrl; p.active = pl->active;
    p.ticket = pl->ticket; p.slots = pl->slots; p.hlog = pl->hlog; p.batch = 1;
    p.tw_full = nullptr; p.tw_chain = 0;
    p.n2_off = pl->long_rank * pl->n2;
    p.n_glob = (int)pl->long_n;
    p.inv_n = (R)1 / (R)pl->long_n;
    p.defer_ctrl = 1;
    if (pl->p2p) {                                              // time layout lives in xbuf; results go to the owners' rows buffers
        p.field = (C*)pl->xbuf;
        for (int r = 0; r < pl->long_ranks; ++r) p.peer[r] = (C*)((char*)pl->peer_base[r] + pl->xlocal);
        p.peer_mode = 1;
        p.peer_shift = ilog2(pl->n1 / pl->long_ranks);
        p.peer_pitch = (int)(pl->long_n / pl->n1);              // N_l
        p.peer_base = pl->long_rank * pl->n2;
    }
    return p;
}

template <typename R>
Params<R> long_inner_params(ssfm_plan_t pl, void* rows) {
    typedef typename cx_of<R>::type C;
    ssfm_plan_t pi = pl->inner;
    bool fixed, single;
    Params<R> p = base_params<R>(pi, pl->last, fixed, single);
    p.field = (C*)rows; p.stash = nullptr; p.ctrl = pl->ctrl; p.active = pl->active;   // controller state of the OUTER plan
    p.ticket = pi->ticket; p.slots = pi->slots; p.hlog = nullptr; p.batch = (int)pi->batch;
    p.inner = 1;
    p.bin_mul = pl->n1;                                        // N0
    p.bin_off = pl->long_rank * (int)pi->batch;                // first outer bin ka of this rank
    p.n_glob = (int)pl->long_n;
    p.wscale = ((1.0 / ((double)pl->long_n * pl->last.dt_s)) * 2.0) * 3.141592653589793 * 1e-12;
    p.inv_n = (R)1; p.att_half = (R)0;                          // 1/N and the attenuation are applied once, by the outer stage
    p.has_nl = 0;
    if (pl->p2p) p.field = (C*)((char*)pl->xbuf + pl->xlocal);  // my rows buffer (filled by the peers' outer kernels)
    return p;
}

template <typename R>
int long_begin_t(ssfm_plan_t pl, void* field, cudaStream_t st) {
    bool fixed, single;
    (void)base_params<R>(pl, pl->last, fixed, single);
    Params<R> p = long_outer_params<R>(pl, field);
    const int one = 1;
    CU_TRY(cudaMemcpyAsync(p.active, &one, sizeof(int), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemsetAsync(p.ctrl, 0, sizeof(Ctrl), st));
    if (!fixed && !single) {                                    // local max |A|^2 -> ctrl.pmax (combined over the ranks by the host)
        k_power_max<R><<<64, 256, 0, st>>>(p, 64);
        ++ssfm_launches;
    }
    CU_TRY(cudaGetLastError());
    return SSFM_OK;
}

template <typename R>
int long_ctrl_t(ssfm_plan_t pl, int init, cudaStream_t st) {
    bool fixed, single;
    (void)base_params<R>(pl, pl->last, fixed, single);
    Params<R> p = long_outer_params<R>(pl, nullptr);
    if (init) k_ctrl_init<R><<<1, 32, 0, st>>>(p, fixed ? 1 : 0, fixed ? (R)pl->last.h_km : (R)0, single ? 1 : 0);
    else k_ctrl_step<R><<<1, 32, 0, st>>>(p);
    ++ssfm_launches;
    CU_TRY(cudaGetLastError());
    return SSFM_OK;
}

template <typename R>
int long_outer_t(ssfm_plan_t pl, void* field, int stage, cudaStream_t st) {
    Params<R> p = long_outer_params<R>(pl, field);
    if (stage == 0) return enqueue_col<R>(p, COL_FWD, st);
    if (stage == 2) { p.peer_mode = 0; return enqueue_col<R>(p, COL_INV, st); }   // time domain out: local stores
    return enqueue_col<R>(p, COL_MID, st, SYNC_FIXED);         // fixed step only: no max, the device controller advances alone
}

template <typename R>
int long_inner_t(ssfm_plan_t pl, void* rows, cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    Params<R> p = long_inner_params<R>(pl, rows);
    int rc = enqueue_col<R>(p, COL_FWD, st);
    if (!rc) rc = enqueue_row<R>(p, st);
    if (pl->p2p) {                                              // last pass stores into the owners' time-layout fields
        for (int r = 0; r < pl->long_ranks; ++r) p.peer[r] = (C*)pl->peer_base[r];
        p.peer_mode = 2;
        p.peer_shift = ilog2(pl->n2);                           // columns per rank
        p.peer_pitch = pl->n2;
        p.peer_base = pl->long_rank * (int)pl->inner->batch;
    }
    if (!rc) rc = enqueue_col<R>(p, COL_INV, st);
    return rc;
}

}  // namespace

extern "C" {

int ssfm_long_plan_create(ssfm_plan_t* out, int64_t n_global, int32_t n_outer, int32_t n_ranks, int32_t rank,
                          int32_t dtype, int32_t device) {
    if (!out) return fail(SSFM_ERR_INVALID, "plan pointer is null");
    *out = nullptr;
    if (dtype != SSFM_C64 && dtype != SSFM_C128) return fail(SSFM_ERR_INVALID, "dtype must be SSFM_C64 or SSFM_C128");
    if (n_global < (1ll << 12) || n_global > (1ll << 30) || (n_global & (n_global - 1)))
        return fail(SSFM_ERR_UNSUPPORTED, "long waveforms: n_samples must be a power of two in [2^12, 2^30]");
    if (n_outer < 16 || n_outer > 2048 || (n_outer & (n_outer - 1)))
        return fail(SSFM_ERR_UNSUPPORTED, "long waveforms: the outer factor must be a power of two in [16, 2048]");
    if (n_ranks < 1 || (n_ranks & (n_ranks - 1)) || rank < 0 || rank >= n_ranks || n_outer % n_ranks)
        return fail(SSFM_ERR_INVALID, "long waveforms: ranks must be a power of two dividing the outer factor");
    const long long nl = n_global / n_outer;
    if (nl < 256 || nl > (1ll << 22)) return fail(SSFM_ERR_UNSUPPORTED, "long waveforms: inner length out of range [2^8, 2^22]");
    const long long n2loc = nl / n_ranks;
    if (n2loc < 32) return fail(SSFM_ERR_UNSUPPORTED, "long waveforms: fewer than 32 columns per rank");
    CU_TRY(cudaSetDevice(device));
    ssfm_plan_s* pl = new ssfm_plan_s();
    pl->device = device; pl->dtype = dtype; pl->n_pol = 1; pl->batch = 1;
    pl->n1 = n_outer; pl->n2 = (int)n2loc; pl->n = (long long)n_outer * n2loc; pl->log2n = ilog2(pl->n);
    pl->long_n = n_global; pl->long_ranks = n_ranks; pl->long_rank = rank;
    pl->lo_bits = (ilog2(n_global) + 1) / 2;
    pl->hlog_cap = 1 << 16;
    pl->persistent = 0;
    const size_t rsz = dtype == SSFM_C64 ? 4 : 8, csz = 2 * rsz;
    const long long n_lo = 1ll << pl->lo_bits, n_hi = n_global >> pl->lo_bits;
    cudaError_t e = cudaMalloc(&pl->stash, (size_t)pl->n * rsz);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->ctrl, sizeof(Ctrl));
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->active, sizeof(int));
    pl->slots_bytes = 64;
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->slots, pl->slots_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->ticket, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaMemset(pl->ticket, 0, sizeof(unsigned int));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->hlog, sizeof(double) * (size_t)pl->hlog_cap);
    if (e == cudaSuccess) e = cudaMalloc(&pl->tw_lo, csz * (size_t)n_lo);
    if (e == cudaSuccess) e = cudaMalloc(&pl->tw_hi, csz * (size_t)n_hi);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&pl->active_host, 2 * sizeof(int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev[0], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev[1], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreate(&pl->wf_ev[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&pl->wf_ev[1]);
    if (e != cudaSuccess) {
        ssfm_plan_destroy(pl);
        return fail(e == cudaErrorMemoryAllocation ? SSFM_ERR_NOMEM : SSFM_ERR_CUDA,
                    std::string("long plan allocation: ") + cudaGetErrorString(e));
    }
    CU_TRY(cudaMemset(pl->ctrl, 0, sizeof(Ctrl)));
    CU_TRY(cudaMemset(pl->hlog, 0, sizeof(double) * (size_t)pl->hlog_cap));
    int rc;
    if (dtype == SSFM_C64) {
        k_build_unit_roots<float><<<(unsigned)((n_lo + 127) / 128), 128>>>((float2*)pl->tw_lo, (int)n_lo, 1, n_global);
        k_build_unit_roots<float><<<(unsigned)((n_hi + 127) / 128), 128>>>((float2*)pl->tw_hi, (int)n_hi, n_lo, n_global);
        rc = build_pass_tables<float>(&pl->tw_col, pl->n1, 0);
    } else {
        k_build_unit_roots<double><<<(unsigned)((n_lo + 127) / 128), 128>>>((double2*)pl->tw_lo, (int)n_lo, 1, n_global);
        k_build_unit_roots<double><<<(unsigned)((n_hi + 127) / 128), 128>>>((double2*)pl->tw_hi, (int)n_hi, n_lo, n_global);
        rc = build_pass_tables<double>(&pl->tw_col, pl->n1, 0);
    }
    if (!rc) rc = plan_create_impl(&pl->inner, nl, 1, n_outer / n_ranks, dtype, device, false);
    if (rc) { ssfm_plan_destroy(pl); return rc; }
    cudaError_t es = cudaDeviceSynchronize();
    if (es != cudaSuccess) { ssfm_plan_destroy(pl); return fail(SSFM_ERR_CUDA, std::string("table build: ") + cudaGetErrorString(es)); }
    *out = pl;
    return SSFM_OK;
}

int ssfm_long_p2p_export(ssfm_plan_t pl, void* handle64) {
    if (!pl || !pl->long_n || !handle64) return fail(SSFM_ERR_INVALID, "null argument or not a long-waveform plan");
    CU_TRY(cudaSetDevice(pl->device));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!pl->xbuf) {
        pl->xlocal = (size_t)pl->n * (pl->dtype == SSFM_C64 ? 8 : 16);
        if ((long long)(pl->long_n / pl->n1) / pl->long_ranks < pl->inner->n2)
            return fail(SSFM_ERR_UNSUPPORTED, "fused exchange: more ranks than rows of the inner transform matrix");
        CU_TRY(cudaMalloc(&pl->xbuf, 2 * pl->xlocal + 256));
        CU_TRY(cudaMemset((char*)pl->xbuf + 2 * pl->xlocal, 0, 256));
    }
    cudaIpcMemHandle_t h;
    CU_TRY(cudaIpcGetMemHandle(&h, pl->xbuf));
    std::memcpy(handle64, &h, 64);
    return SSFM_OK;
}

int ssfm_long_p2p_import(ssfm_plan_t pl, const void* handles) {
    if (!pl || !pl->long_n || !handles || !pl->xbuf) return fail(SSFM_ERR_INVALID, "null argument, or ssfm_long_p2p_export not called");
    if (pl->long_ranks > 8) return fail(SSFM_ERR_UNSUPPORTED, "fused exchange supports up to 8 ranks");
    CU_TRY(cudaSetDevice(pl->device));
    unsigned int* flags[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    for (int r = 0; r < pl->long_ranks; ++r) {
        if (r == pl->long_rank) pl->peer_base[r] = pl->xbuf;
        else if (!pl->peer_base[r]) {
            cudaIpcMemHandle_t h;
            std::memcpy(&h, (const char*)handles + 64 * r, 64);
            CU_TRY(cudaIpcOpenMemHandle(&pl->peer_base[r], h, cudaIpcMemLazyEnablePeerAccess));
        }
        flags[r] = (unsigned int*)((char*)pl->peer_base[r] + 2 * pl->xlocal);
    }
    if (!pl->d_peer_flags) CU_TRY(cudaMalloc((void**)&pl->d_peer_flags, sizeof(flags)));
    CU_TRY(cudaMemcpy(pl->d_peer_flags, flags, sizeof(flags), cudaMemcpyHostToDevice));
    pl->p2p = 1;
    return SSFM_OK;
}

int ssfm_long_p2p_copy(ssfm_plan_t pl, void* field_user, int32_t to_internal, void* stream) {
    if (!pl || !pl->p2p || !field_user) return fail(SSFM_ERR_INVALID, "null argument or fused exchange not set up");
    CU_TRY(cudaSetDevice(pl->device));
    if (to_internal) CU_TRY(cudaMemcpyAsync(pl->xbuf, field_user, pl->xlocal, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    else CU_TRY(cudaMemcpyAsync(field_user, pl->xbuf, pl->xlocal, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return SSFM_OK;
}

int ssfm_long_xbar(ssfm_plan_t pl, void* stream) {
    if (!pl || !pl->p2p) return fail(SSFM_ERR_INVALID, "null plan or fused exchange not set up");
    CU_TRY(cudaSetDevice(pl->device));
    ++pl->epoch;
    k_xbar<<<1, 32, 0, (cudaStream_t)stream>>>(pl->d_peer_flags, (unsigned int*)((char*)pl->xbuf + 2 * pl->xlocal), pl->long_rank,
                                               pl->long_ranks, pl->epoch);
    ++ssfm_launches;
    CU_TRY(cudaGetLastError());
    return SSFM_OK;
}

#define LONG_DISPATCH(pl, call_f, call_d) ((pl)->dtype == SSFM_C64 ? (call_f) : (call_d))

int ssfm_long_begin(ssfm_plan_t pl, void* field_local, const ssfm_fiber_params* prm, void* stream) {
    if (!pl || !pl->long_n || !field_local || !prm) return fail(SSFM_ERR_INVALID, "null argument or not a long-waveform plan");
    if (!(prm->dt_s > 0)) return fail(SSFM_ERR_INVALID, "dt_s must be > 0");
    CU_TRY(cudaSetDevice(pl->device));
    pl->last = *prm;
    pl->have_state = true;
    pl->last_kind = 1;
    cudaStream_t st = (cudaStream_t)stream;
    return LONG_DISPATCH(pl, long_begin_t<float>(pl, field_local, st), long_begin_t<double>(pl, field_local, st));
}

int ssfm_long_pmax(ssfm_plan_t pl, double* value, int32_t set, void* stream) {
    if (!pl || !pl->long_n || !value) return fail(SSFM_ERR_INVALID, "null argument or not a long-waveform plan");
    CU_TRY(cudaSetDevice(pl->device));
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long bits = 0;
    if (set) {                                                   // value holds a number of the compute real type
        if (pl->dtype == SSFM_C64) { float f = (float)*value; unsigned int u; std::memcpy(&u, &f, 4); bits = u; }
        else std::memcpy(&bits, value, 8);
        CU_TRY(cudaMemcpyAsync(&pl->ctrl->pmax, &bits, sizeof(bits), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaStreamSynchronize(st));
    } else {
        CU_TRY(cudaMemcpyAsync(&bits, &pl->ctrl->pmax, sizeof(bits), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        if (pl->dtype == SSFM_C64) { float f; unsigned int u = (unsigned int)bits; std::memcpy(&f, &u, 4); *value = (double)f; }
        else std::memcpy(value, &bits, 8);
    }
    return SSFM_OK;
}

int ssfm_long_ctrl(ssfm_plan_t pl, int32_t init, void* stream) {
    if (!pl || !pl->long_n) return fail(SSFM_ERR_INVALID, "null plan or not a long-waveform plan");
    CU_TRY(cudaSetDevice(pl->device));
    cudaStream_t st = (cudaStream_t)stream;
    return LONG_DISPATCH(pl, long_ctrl_t<float>(pl, init, st), long_ctrl_t<double>(pl, init, st));
}

int ssfm_long_outer(ssfm_plan_t pl, void* field_local, int32_t stage, void* stream) {
    if (!pl || !pl->long_n || !field_local) return fail(SSFM_ERR_INVALID, "null argument or not a long-waveform plan");
    if (stage < 0 || stage > 2) return fail(SSFM_ERR_INVALID, "stage must be 0 (open), 1 (mid) or 2 (close)");
    CU_TRY(cudaSetDevice(pl->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = LONG_DISPATCH(pl, long_outer_t<float>(pl, field_local, stage, st), long_outer_t<double>(pl, field_local, stage, st));
    if (!rc) CU_TRY(cudaGetLastError());
    return rc;
}

int ssfm_long_inner(ssfm_plan_t pl, void* rows_local, void* stream) {
    if (!pl || !pl->long_n || !rows_local) return fail(SSFM_ERR_INVALID, "null argument or not a long-waveform plan");
    CU_TRY(cudaSetDevice(pl->device));
    cudaStream_t st = (cudaStream_t)stream;
    int rc = LONG_DISPATCH(pl, long_inner_t<float>(pl, rows_local, st), long_inner_t<double>(pl, rows_local, st));
    if (!rc) CU_TRY(cudaGetLastError());
    return rc;
}

}  // extern "C"

// =================================================================================================
// Arbitrary lengths (chirp-z): see the kernels k_bs_* in ssfm_kernels.cuh.
// =================================================================================================
namespace {

// FFT_M -> * xf -> IFFT_M of the padded rows with the power-of-two kernels (xf already in transposed order)
template <typename R>
int chirp_transfer(ssfm_plan_t pl, const void* xf, cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    ssfm_plan_t pi = pl->inner;
    ssfm_fiber_params prm{};
    prm.dt_s = 1.0; prm.length_km = 1e30; prm.phi_max_rad = 0.01; prm.h_km = 1.0;
    bool fixed, single;
    Params<R> p = base_params<R>(pi, prm, fixed, single);
    p.field = (C*)pl->wb; p.stash = nullptr; p.ctrl = pi->ctrl; p.active = pi->active; p.hlog = nullptr;
    p.ticket = pi->ticket; p.slots = pi->slots; p.batch = (int)pi->batch; p.xfer = (const C*)xf;
    p.inner = 1; p.has_nl = 0; p.att_half = (R)0;              // dummy controller word (done = 0), no scaling but 1/M
    int rc = enqueue_col<R>(p, COL_FWD, st);
    if (!rc) rc = enqueue_row<R>(p, st);
    if (!rc) rc = enqueue_col<R>(p, COL_INV, st);
    return rc;
}

template <typename R>
int chirp_build_tables(ssfm_plan_t pl) {
    typedef typename cx_of<R>::type C;
    ssfm_plan_t pi = pl->inner;
    const int n = (int)pl->n, m = (int)pl->chirp_m;
    k_bs_chirp<R><<<(n + 255) / 256, 256>>>((C*)pl->wtab, n);
    for (int which = 0; which < 2; ++which) {                   // forward: FFT_M(conj w); inverse: FFT_M(w)
        C* dst = (C*)(which == 0 ? pl->xf_fwd : pl->xf_inv);
        k_bs_kernel_row<R><<<(m + 255) / 256, 256>>>(dst, (const C*)pl->wtab, n, m, which == 0 ? 1 : 0);
        ssfm_fiber_params prm{};
        prm.dt_s = 1.0; prm.length_km = 1e30; prm.phi_max_rad = 0.01; prm.h_km = 1.0;
        bool fixed, single;
        Params<R> p = base_params<R>(pi, prm, fixed, single);
        p.field = dst; p.ctrl = pi->ctrl; p.active = pi->active; p.ticket = pi->ticket; p.slots = pi->slots;
        p.batch = 1; p.inner = 1; p.has_nl = 0; p.fwd_only = 1;
        int rc = enqueue_col<R>(p, COL_FWD, 0);
        if (!rc) rc = enqueue_row<R>(p, 0);                     // spectrum of the chirp, left in transposed order in place
        if (rc) return rc;
    }
    CU_TRY(cudaGetLastError());
    return SSFM_OK;
}

}  // namespace

static int chirp_plan_create(ssfm_plan_t* out, int64_t n, int32_t n_pol, int64_t batch, int32_t dtype, int32_t device) {
    CU_TRY(cudaSetDevice(device));
    long long m = 256;
    while (m < 2 * n - 1) m <<= 1;
    ssfm_plan_s* pl = new ssfm_plan_s();
    pl->device = device; pl->dtype = dtype; pl->n_pol = n_pol; pl->n = n; pl->batch = batch;
    pl->chirp_m = m; pl->persistent = 0;
    pl->hlog_cap = 4096;
    if ((size_t)batch * pl->hlog_cap * sizeof(double) > (256u << 20)) pl->hlog_cap = (int)((256u << 20) / (batch * sizeof(double)));
    if (pl->hlog_cap < 16) pl->hlog_cap = 16;
    const size_t rsz = dtype == SSFM_C64 ? 4 : 8, csz = 2 * rsz;
    const size_t rows = (size_t)batch * n_pol;
    cudaError_t e = cudaMalloc(&pl->stash, rows * (size_t)n * rsz);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->ctrl, sizeof(Ctrl) * (size_t)batch);
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->active, sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc((void**)&pl->hlog, sizeof(double) * (size_t)batch * pl->hlog_cap);
    if (e == cudaSuccess) e = cudaMalloc(&pl->wb, rows * (size_t)m * csz);
    if (e == cudaSuccess) e = cudaMalloc(&pl->wtab, (size_t)n * csz);
    if (e == cudaSuccess) e = cudaMalloc(&pl->xf_fwd, (size_t)m * csz);
    if (e == cudaSuccess) e = cudaMalloc(&pl->xf_inv, (size_t)m * csz);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&pl->active_host, 2 * sizeof(int), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&pl->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaEventCreate(&pl->wf_ev[0]);
    if (e == cudaSuccess) e = cudaEventCreate(&pl->wf_ev[1]);
    if (e != cudaSuccess) {
        ssfm_plan_destroy(pl);
        return fail(e == cudaErrorMemoryAllocation ? SSFM_ERR_NOMEM : SSFM_ERR_CUDA,
                    std::string("chirp plan allocation: ") + cudaGetErrorString(e));
    }
    CU_TRY(cudaMemset(pl->ctrl, 0, sizeof(Ctrl) * (size_t)batch));
    CU_TRY(cudaMemset(pl->hlog, 0, sizeof(double) * (size_t)batch * pl->hlog_cap));
    int rc = plan_create_impl(&pl->inner, m, 1, (int64_t)rows, dtype, device, false);
    if (!rc) rc = dtype == SSFM_C64 ? chirp_build_tables<float>(pl) : chirp_build_tables<double>(pl);
    if (rc) { ssfm_plan_destroy(pl); return rc; }
    cudaError_t es = cudaDeviceSynchronize();
    if (es != cudaSuccess) { ssfm_plan_destroy(pl); return fail(SSFM_ERR_CUDA, std::string("chirp tables: ") + cudaGetErrorString(es)); }
    *out = pl;
    return SSFM_OK;
}

template <typename R>
static int chirp_propagate_t(ssfm_plan_t pl, void* field, const ssfm_fiber_params& prm, long long max_steps, int resume,
                             cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    bool fixed, single;
    Params<R> p = base_params<R>(pl, prm, fixed, single);
    ssfm_plan_t pi = pl->inner;
    const int E = points_per_thread<R>::value;
    const C* sct = (const C*)pi->tw_col + pass_table_size(pi->n1, E);          // sincos table behind the inner pass tables
    p.field = (C*)field; p.stash = (R*)pl->stash; p.ctrl = pl->ctrl; p.active = pl->active; p.hlog = pl->hlog;
    p.batch = (int)pl->batch;
    p.n_glob = (int)pl->n;
    p.inv_n = (R)1 / (R)pl->n;
    const int m = (int)pl->chirp_m;
    const int rows = (int)(pl->batch * pl->n_pol);
    const long long budget = max_steps > 0 ? max_steps : (1ll << 40);
    pl->last_kind = 1;
    int act = (int)pl->batch;
    if (!resume) {
        CU_TRY(cudaMemcpyAsync(p.active, &act, sizeof(int), cudaMemcpyHostToDevice, st));
        CU_TRY(cudaMemsetAsync(p.ctrl, 0, sizeof(Ctrl) * (size_t)pl->batch, st));
        if (!fixed && !single) {
            k_power_max<R><<<(unsigned)(pl->batch * 8), 256, 0, st>>>(p, 8);
            ++ssfm_launches;
        }
        k_ctrl_init<R><<<(unsigned)((pl->batch + 127) / 128), 128, 0, st>>>(p, fixed ? 1 : 0, fixed ? (R)prm.h_km : (R)0, single ? 1 : 0);
        ++ssfm_launches;
    } else {
        std::vector<Ctrl> h((size_t)pl->batch);
        CU_TRY(cudaMemcpyAsync(h.data(), p.ctrl, sizeof(Ctrl) * h.size(), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
        act = 0;
        for (auto& c : h) act += c.done ? 0 : 1;
        CU_TRY(cudaMemcpyAsync(p.active, &act, sizeof(int), cudaMemcpyHostToDevice, st));
    }
    CU_TRY(cudaMemcpyAsync(&pl->active_host[0], p.active, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    const dim3 grid((unsigned)std::min<long long>((m + 255) / 256, 1024), (unsigned)rows);
    for (long long s = 0; s < budget && pl->active_host[0] > 0; ++s) {
        k_bs_open<R><<<grid, 256, 0, st>>>(p, (C*)pl->wb, (const C*)pl->wtab, m, sct);
        int rc = chirp_transfer<R>(pl, pl->xf_fwd, st);
        if (rc) return rc;
        k_bs_mid<R><<<grid, 256, 0, st>>>(p, (C*)pl->wb, m, sct);
        rc = chirp_transfer<R>(pl, pl->xf_inv, st);
        if (rc) return rc;
        k_bs_close<R><<<grid, 256, 0, st>>>(p, (const C*)pl->wb, (const C*)pl->wtab, m, sct);
        k_ctrl_steps<R><<<(unsigned)((pl->batch + 127) / 128), 128, 0, st>>>(p);
        ssfm_launches += 4;
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaMemcpyAsync(&pl->active_host[0], p.active, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU_TRY(cudaStreamSynchronize(st));
    }
    pl->have_state = true;
    pl->last = prm;
    return SSFM_OK;
}

// row <- ifft_N(fft_N(row) * H) for a length that is not a power of two: the chirp-z pipeline without the Kerr steps
template <typename R>
static int chirp_apply_transfer_t(ssfm_plan_t pl, void* field, const void* h_dev, cudaStream_t st) {
    typedef typename cx_of<R>::type C;
    ssfm_fiber_params prm{};
    prm.dt_s = 1.0; prm.length_km = 1e30; prm.phi_max_rad = 0.01; prm.h_km = 1.0;
    bool fixed, single;
    Params<R> p = base_params<R>(pl, prm, fixed, single);      // gamma = 0: no Kerr rotation, alpha = 0: no attenuation
    ssfm_plan_t pi = pl->inner;
    const C* sct = (const C*)pi->tw_col + pass_table_size(pi->n1, points_per_thread<R>::value);
    p.field = (C*)field; p.stash = (R*)pl->stash; p.ctrl = pl->ctrl; p.active = pl->active; p.hlog = nullptr;
    p.batch = (int)pl->batch; p.n_glob = (int)pl->n; p.inv_n = (R)1 / (R)pl->n; p.xfer = (const C*)h_dev;
    CU_TRY(cudaMemsetAsync(p.ctrl, 0, sizeof(Ctrl) * (size_t)pl->batch, st));   // done = 0, h = 0
    const int m = (int)pl->chirp_m, rows = (int)(pl->batch * pl->n_pol);
    const dim3 grid((unsigned)std::min<long long>((m + 255) / 256, 1024), (unsigned)rows);
    k_bs_open<R><<<grid, 256, 0, st>>>(p, (C*)pl->wb, (const C*)pl->wtab, m, sct);
    int rc = chirp_transfer<R>(pl, pl->xf_fwd, st);
    if (rc) return rc;
    k_bs_mid<R><<<grid, 256, 0, st>>>(p, (C*)pl->wb, m, sct);
    rc = chirp_transfer<R>(pl, pl->xf_inv, st);
    if (rc) return rc;
    k_bs_close<R><<<grid, 256, 0, st>>>(p, (const C*)pl->wb, (const C*)pl->wtab, m, sct);
    ssfm_launches += 3;
    CU_TRY(cudaGetLastError());
    pl->have_state = false;
    return SSFM_OK;
}
