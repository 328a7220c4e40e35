// k_wf: the whole FIBER / DBP propagation of a batch as ONE persistent kernel (sm_100a).
//
// Reference loop: opticomlib/devices.py:1155-1196.  The multi-launch schedule of ssfm_kernels.cuh streams
// every waveform through HBM twice per step (k_row, k_col_mid).  Here a TEAM of `total` co-resident CTAs
// (total = n_pol * N/4096, N = 2^12 .. 2^20) adopts waveforms -- handed out dynamically, since step counts
// differ per waveform -- and carries each through ALL of its steps, so that
//   * the field of the waveforms in flight (teams x N samples: a few MiB .. ~30 MiB) never leaves the 126 MB
//     L2 -- HBM sees one read and one write of the field per PROPAGATION, not four per step;
//   * the Kerr-phase stash of a tile lives in the shared memory of the CTA that owns the tile (the same CTA
//     visits the same tile every step): the stash traffic of the multi-launch schedule is gone;
//   * pass tables and the sincos table are loaded once per CTA; there are no launches, tickets or host
//     polls inside a propagation, and the step-size controller runs redundantly in every CTA.
//
// The same kernel applies an arbitrary transfer function (zero-phase filters, DM): one fixed step with gamma = alpha = 0 and the
// row phase multiplying by a table (Params::xfer) is  FFT -> x H -> IFFT  in one pass, out of place if Params::field_in is set.
//
// Per step each CTA runs a ROW phase (G rows of the N1 x N2 matrix: forward transform, exp(D~ h),
// inverse transform) and a COLUMN phase (T columns: inverse transform, 1/N, max|A|^2 -> team exchange
// -> controller -> merged Kerr rotation of the second half step of step s and the first half step of
// step s+1 -> forward transform).  The phases of one waveform are separated by team barriers.
//
// Four variants of the team structure (template parameter TM; CL := TM == 1 or 3, MC := TM == 2, MT := TM == 3):
//   MT          (TM == 3) a team is ONE cluster of 2 .. 16 CTAs whatever the waveform length and every CTA carries several
//               4096-sample tiles through each phase (a loop around the phase body): barrier, exchange and load latencies are paid
//               once per phase, not once per tile.  The Kerr phase of the tiles in flight lives in an L2-resident team stash
//               (WfArgs::tstash) and is staged into shared memory with cp.async; adaptive step control makes two passes over the
//               CTA's tiles per column phase (end of step -> store -> team maximum -> re-read -> merged rotation -> forward
//               transforms).  Used for waveforms of 32 .. 64 tiles (one 16-CTA cluster each) and, as clusters of 2, for the CTA
//               slots that 16-CTA clusters cannot use;
//   MC          teams of 32 .. 256 CTAs (N = 2^17 .. 2^20) made of thread-block clusters of 8: the team barrier is the hardware
//               barrier inside every cluster plus ONE flag hop between the cluster leaders (4 .. 32 arrivals on the counter in
//               L2 instead of 32 .. 256, and only the leaders poll); the maxima travel through st.async inside a cluster and as
//               self-validating words between the leaders.  Launched cooperatively (all clusters of a team must be resident).
//   CL          teams of <= 16 CTAs are thread-block clusters: hardware cluster barrier (arrive.release /
//               wait.acquire by every thread); per-WARP maxima to every CTA of the cluster with st.async, counted by the
//               destination's mbarrier (no barrier and no fence in the middle of the column phase); the waveform index
//               through distributed shared memory.  1.4x (fp64) .. 1.6x (fp32) faster per team, but a 16-CTA cluster needs
//               16 distinct SMs of one GPC, so fewer teams are co-resident than the chip has CTA slots;
//   TM == 0     any team size: a monotonic arrival counter in L2 (red.release.gpu / relaxed poll + fence.acq_rel.gpu),
//               maxima as self-validating 64-bit words {value bits | exchange tag}, waveform index through a mailbox
//               word; launched cooperatively over every CTA slot, or as the second launch that fills the slots the
//               clusters leave (ssfm_wf_impl.inl).
//
// Tried and dropped (measured on B200): letting a team multiplex two or three waveforms
// ("slots") so that it runs a phase of waveform B while the barrier of waveform A completes.  The barrier waits
// shrank from ~5 k to ~2 k cycles per phase but the time moved into the max|A|^2 exchange in the middle of the
// column phase (where the registers ARE live and the CTA cannot switch), and the stash had to travel through
// L2: 4.2e10 instead of 4.6e10 sample*steps/s.
//
// Everything here is FP64/FP32 FMA-pipe arithmetic on L2-resident data; no tensor cores (no dense
// contraction on this path).  Loads of the field bypass L1 (ld.global.cg): the data was written by other
// SMs one phase earlier.
#pragma once
#include "ssfm_kernels.cuh"

namespace ssfm {

template <typename R>
struct WfArgs {
    unsigned int* bar;            // [n_teams][32]  monotonic arrival counter of the team barrier (one 128-B line each)
    unsigned long long* mail;     // [n_teams][16]  (sequence << 32 | waveform) handed out by CTA 0 of the team
    unsigned long long* slots;    // [n_teams][2][total][2]  self-validating max words, double-buffered by parity
    unsigned int* next_wf;        // next waveform to hand out (dynamic assignment: step counts differ per waveform)
    unsigned int* sm_cnt;         // [1024] CTAs of this launch that registered on each SM (team placement)
    unsigned int* grid_bar;       // arrival counter of the one grid-wide barrier that ends the registration
    long long budget;             // stop every waveform after this many steps in this call (max_steps of the C-ABI)
    int occ;                      // CTAs per SM (the grid is occ x num_sms, all co-resident)
    int placement;                // 1 = SM-aware team placement (default), 0 = by blockIdx (experiments)
    int n_teams;
    int fixed, single, resume;
    R h_fixed;
    R* tstash;                    // TM == 3: Kerr phase of the waveforms in flight, [n_teams][units][16][256] (L2-resident)
    int draw_min;                 // TM == 3 as the slower second launch: stop drawing when fewer waveforms than this are left
    // streamed batches (ssfm_propagate_streamed): the waveforms arrive from the host WHILE the kernel runs.  ready[0] = number of
    // waveforms whose samples have landed (written in stream order behind every chunk's copy): a drawn waveform waits for it.
    // done[c] counts the 4096-sample tiles of chunk c (chunk_rows waveforms each) whose final samples are stored: the stream
    // that copies chunk c back to the host waits for it to reach chunk rows x tiles per waveform.  Null: nothing of this.
    const unsigned int* ready;
    unsigned int* done;
    int chunk_rows;
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
    // separable linear operator (complex128, rows of 256 bins, beta_3 = 0): exp(j theta (M1 k2')^2) for |k2'| <= M2/2
    static constexpr int LSEP = (sizeof(R) == 8 && M2 == 256) ? (M2 / 2 + 1) : 0;
    static constexpr size_t smem_full = sizeof(C) * (size_t)(XB + TAB1 + TAB2 + SC_N + LSEP) + sizeof(R) * (size_t)(E * NT);
    // Two CTAs per SM need <= (228 KB / 2 - 1 KB) each.  When the pass tables do not fit next to the exchange buffer
    // and the stash (fp64, transforms of 1024 points or N1 != N2 >= 512) they are read through L1 from global memory.
    static constexpr bool TABS = smem_full <= 115712;
    static constexpr size_t smem = TABS ? smem_full : smem_full - sizeof(C) * (size_t)(TAB1 + TAB2);
};

// release-arrive / relaxed poll + acquire fence (PTX memory model, gpu scope).  The release is cumulative: it also
// covers the stores of the other threads of the CTA that were ordered before it by the preceding bar.sync -- the
// same reasoning cooperative-groups grid synchronisation rests on, without the sequentially-consistent fence.
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
    asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }
__device__ __forceinline__ unsigned int ld_relaxed_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// waveform w of a streamed batch: wait until the host-to-device copy of its chunk has been flagged (one thread)
template <typename A> __device__ __forceinline__ void wait_arrival(const A& a, unsigned int w, int batch) {
    if (a.ready && w < (unsigned int)batch) {
        while (ld_relaxed_sys_u32(a.ready) <= w) __nanosleep(500);
        asm volatile("fence.acq_rel.sys;" ::: "memory");
    }
}

enum { WF_GRAB = 0, WF_ROW = 1, WF_COL = 2, WF_END = 3 };
#ifndef SSFM_WF_CTAS_F32
#define SSFM_WF_CTAS_F32 3      // resident CTAs per SM asked of the compiler for complex64 (80 registers per thread)
#endif

// State of the team's current waveform, uniform over the CTA and over the team.  Its home is SHARED memory (two versions,
// written by thread 0 before a team barrier and read by everyone after it): carried in registers across a phase it gets
// spilled to local memory, and since the cluster barrier invalidates L1 every phase then began with a serialised L2 round
// trip before the first field load could be issued (round-1 ncu source page: ~4 % of all warp samples).
template <typename R>
struct WfShared {
    unsigned int w;               // waveform index
    unsigned int xchg;            // max|A|^2 exchanges of this team so far (mbarrier phase parity / tag of the flag-based words)
    int steps;                    // steps taken so far
    long long taken;              // steps taken in this call (budget)
    R z, h;                       // position reached, size of the step in progress
};

__device__ __forceinline__ unsigned cluster_id_x() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void st_cluster_u32(void* local_smem, unsigned rank, unsigned v) {
    const unsigned a = (unsigned)__cvta_generic_to_shared(local_smem);
    unsigned ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}

// mbarrier in shared memory + st.async through distributed shared memory: the data word and its "arrived" signal travel
// together (complete_tx on the destination CTA's barrier), so the max|A|^2 exchange of a cluster team needs neither the cluster
// barrier nor a release fence in the middle of the column phase (where the round-1 kernel stalled ~1.5 k cycles in MEMBAR).
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
// store `v` into `local_slot` of CTA `rank` and count 8 bytes on that CTA's copy of `local_bar`
__device__ __forceinline__ void st_async_u64(void* local_slot, unsigned long long* local_bar, unsigned rank, unsigned long long v) {
    unsigned ra, rb;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"((unsigned)__cvta_generic_to_shared(local_slot)), "r"(rank));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"((unsigned)__cvta_generic_to_shared(local_bar)), "r"(rank));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(ra), "l"(v), "r"(rb) : "memory");
}

// CL = true: the team is one thread-block cluster (teams of <= 16 CTAs): the team barrier is the hardware cluster barrier
// (arrive.release / wait.acquire, executed by every thread), the per-CTA maxima and the waveform index travel through
// distributed shared memory, and no cooperative launch, registration or global-memory flag is needed.
template <typename R, int M1, int M2, bool SMALL, int TM>
__global__ void __launch_bounds__(256, (sizeof(R) == 8 ? 2 : SSFM_WF_CTAS_F32)) k_wf(Params<R> p, WfArgs<R> a) {
    typedef typename cx_of<R>::type C;
    constexpr bool CL = (TM == 1 || TM == 3), MC = (TM == 2), MT = (TM == 3);
    typedef wf_geom<R, M1, M2> GEO;
    // pass twiddles: complex128 builds the 15 powers of a radix-16 pass from ONE table entry (14 complex products) instead of
    // 15 shared-memory loads of 16 B -- its LSU path is busier than its FP64 pipe; complex64 keeps the table
#ifdef SSFM_WF_NO_HALVES
    constexpr bool HALVES = false;
#else
    // 256-point column transforms in complex128: two half-CTAs with their own barriers (measured on B200, config #3: fp64 5.78e10 ->
    // 5.89e10; fp32, three CTAs per SM, 1.02e11 -> 9.2e10: off)
#ifdef SSFM_WF_HALVES_T8    // (512-point column transforms, halves of 4 columns = 64-byte segments of global memory: DBP of 128 x 2^18
                            //  3.87e10 -> 3.43e10, the same loss as in fp32 -- a warp access then touches 8 lines instead of 4)
    constexpr bool HALVES = ((GEO::T == 16 || GEO::T == 8) && GEO::NT == 256 && sizeof(R) == 8);
#else
    constexpr bool HALVES = (GEO::T == 16 && GEO::NT == 256 && sizeof(R) == 8);
#endif
#endif
    typedef typename std::conditional<HALVES, ColExchangeHalves<GEO::T>, ColExchange<GEO::T>>::type CX;
#ifdef SSFM_WF_TABLE_TWIDDLES
    typedef TableTwiddles WFTW;
#else
    typedef typename std::conditional<sizeof(R) == 8, ChainTwiddles, TableTwiddles>::type WFTW;
#endif
    constexpr int E = GEO::E, NT = GEO::NT, T = GEO::T, G = GEO::G, PM = GEO::PM;
    static_assert(points_per_thread<R>::value == 16, "k_wf assumes 16 points per thread");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ unsigned long long red[32];
    __shared__ unsigned int s_w, s_wcl[2];
    __shared__ unsigned long long cl_max[16 * (256 / 32)];     // [CTA of the cluster][warp] maxima (st.async from every CTA)
    __shared__ __align__(8) unsigned long long mb_max;         // mbarrier counting the bytes that arrive in cl_max
    __shared__ WfShared<R> sh[2];
    C* xb = reinterpret_cast<C*>(smem_raw);                    // exchange buffer: column tile [M1][T] or G padded rows
    constexpr bool TABS = GEO::TABS;
    C* tw1s = xb + GEO::XB;                                    // pass tables of the N1-point (column) transforms
    C* tw2s = tw1s + (TABS ? GEO::TAB1 : 0);                   // ... of the N2-point (row) transforms when N2 != N1
    C* sct = tw2s + (TABS ? GEO::TAB2 : 0);                    // sincos table
    C* btab = sct + SC_N;                                      // [LSEP] separable linear operator, see lin_sep_table below
    R* st_sm = reinterpret_cast<R*>(btab + GEO::LSEP);         // [E][NT] Kerr phase of the current step of MY tile
    const C* tw1 = TABS ? tw1s : p.tw_col;
    const C* tw2 = TABS ? ((M1 == M2) ? tw1s : tw2s) : p.tw_row;

    const int tid = threadIdx.x;
    constexpr int tiles = M2 / T;                              // column tiles (= row groups) per polarisation (p.n2 == M2, p.n == M1*M2)
    constexpr int NN = M1 * M2;
    const unsigned units = (unsigned)(tiles * p.n_pol);        // 4096-point tiles of one waveform
    unsigned total = units;                                    // CTAs per team (TM == 3: the cluster size, set below)

    // ---- team placement.  A team is bulk-synchronous, so its CTAs should run at the same speed, and a CTA's
    // speed depends on what the OTHER CTAs of its SM are doing.  The grid fills every CTA slot of the chip; each
    // CTA registers on its SM, and once all have, teams are laid out so that the `occ` CTAs of an SM belong to
    // `occ` different teams that share the same group of `total` SMs: every CTA of a team then has exactly the
    // same neighbours (the lock-stepped CTAs of the sibling teams).  Measured: +16 % in fp32 (three CTAs per SM),
    // -3 % in fp64 (two per SM), so the host enables it for fp32 only; teams larger than the SM count are
    // spread by blockIdx.
    __shared__ unsigned int s_slot, s_rank;
    int team, me;
    unsigned csz = 1u, crank = 0u, ncl = 1u, sub = 0u;       // CTAs per cluster, my rank in it; clusters per team, my cluster's index
    if (CL || MC) {
        crank = cluster_ctarank();
        asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(csz));
        if (CL) {
            team = (int)cluster_id_x();
            me = (int)crank;
            if (MT) total = csz;                               // every CTA carries units / csz tiles through each phase
        } else {
            ncl = total / csz;
            team = (int)(cluster_id_x() / ncl);
            sub = cluster_id_x() % ncl;
            me = (int)(sub * csz + crank);
        }
        if (tid == 0) {
            mbar_init(&mb_max, 1u);                            // one arrival per exchange: thread 0's expect_tx
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_arrive_release();                              // every CTA of the cluster has started (its shared memory and its
        cluster_wait_acquire();                                // barrier exist) before anyone stores into it through DSMEM
    } else {
    unsigned int smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (tid == 0) {
        s_rank = 0u;
        s_slot = atomicAdd(a.sm_cnt + smid, 1u);
        __threadfence();
        atomicAdd(a.grid_bar, 1u);
        while (ld_relaxed_u32(a.grid_bar) < gridDim.x) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
    {
        unsigned int lower = 0;                                // SMs with a smaller id that host CTAs of this launch
        for (unsigned int i = tid; i < smid; i += NT) lower += (ld_relaxed_u32(a.sm_cnt + i) != 0u) ? 1u : 0u;
        if (lower) atomicAdd(&s_rank, lower);
    }
    __syncthreads();
    const unsigned int n_sm = gridDim.x / (unsigned)a.occ;
    if (!a.placement || total > n_sm) {                        // by blockIdx: consecutive CTAs spread over the SMs
        team = (int)(blockIdx.x % (unsigned)a.n_teams);
        me = (int)(blockIdx.x / (unsigned)a.n_teams);
        if (me >= (int)total) team = a.n_teams;
    } else {
        const unsigned int grp = s_rank / total;
        team = (int)(grp * (unsigned)a.occ + s_slot);
        me = (int)(s_rank % total);
        if (grp >= n_sm / total) team = a.n_teams;             // SMs beyond the last full group stay idle
    }
    if (team >= a.n_teams) return;
    }

    const int pol = me / tiles, tile = me % tiles;
    constexpr int HT = T / 2;                                           // (HALVES: columns 0..T/2-1 on warps 0..3, the rest on warps 4..7)
    const int c = HALVES ? ((tid & (HT - 1)) | ((tid >> 7) * HT)) : tid % T;   // column phase: column c of the tile, thread t of its transform
    const int t = HALVES ? ((tid & 127) / HT) : tid / T;
    const int n2 = tile * T + c;
    const int g = tid / (M2 / E), tr = tid % (M2 / E);         // row phase: row g of the group, thread tr of its transform
    const int k1 = tile * G + g;

    if (TABS) {
        for (int i = tid; i < GEO::TAB1; i += NT) tw1s[i] = p.tw_col[i];
        for (int i = tid; i < GEO::TAB2; i += NT) tw2s[i] = p.tw_row[i];
    }
    for (int i = tid; i < SC_N; i += NT) sct[i] = p.tw_col[GEO::TAB1 + i];

#ifdef SSFM_WF_PROFILE
    long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pn = 0;
#define WF_T0(name) const long long name = clock64()
#define WF_ACC(i, name) prof[i] += clock64() - name
#else
#define WF_T0(name) do { } while (0)
#define WF_ACC(i, name) do { } while (0)
#endif

    // team barrier, split: arrive after the phase's stores, wait before the next phase's loads
    unsigned int* const bar = a.bar + (size_t)team * 32;
    unsigned int bar_target = 0u, xchg = 0u, seq = 0u, ver = 0u;   // (xchg is reloaded from shared memory in every phase)
    int state = WF_GRAB;
    auto bar_arrive = [&]() {
        if (CL) { cluster_arrive_release(); return; }           // every thread releases its own stores
        if (MC) { cluster_arrive_release(); bar_target += ncl; return; }
        if (total == 1u) return;                                // a team of one CTA: the CTA barrier in bar_wait is enough
        __syncthreads();                                        // every thread's stores are ordered before the release
        bar_target += total;                                  // (one arrival per WARP instead -- no CTA barrier, each warp releases
        if (tid == 0) red_release_add_u32(bar, 1u);             //  its own stores -- was measured 25-35 % slower: 8x the atomics and polls)
    };
    auto bar_wait = [&]() {
        if (CL) { cluster_wait_acquire(); return; }
        if (MC) {                                               // hardware barrier in the cluster, one flag hop between the leaders,
            cluster_wait_acquire();                             // hardware barrier again to pass the news on
            if (crank == 0u && tid == 0) {
                red_release_add_u32(bar, 1u);                   // (cumulative: covers the stores the cluster barrier made visible to me)
                while ((int)(ld_relaxed_u32(bar) - bar_target) < 0) { }
                fence_acq_rel_gpu();
            }
            cluster_arrive_release();
            cluster_wait_acquire();
            return;
        }
        if (total == 1u) { __syncthreads(); return; }           // stores (write-through) -> bar.sync -> ld.global.cg of the same CTA
        if (tid == 0) {
            while ((int)(ld_relaxed_u32(bar) - bar_target) < 0) { if (total > 32u) __nanosleep(20); }   // big teams: ease off L2
            fence_acq_rel_gpu();
        }
        __syncthreads();
    };

    // max over the team of a per-thread value (NaN wins, like numpy's max): block reduction, one self-validating
    // word (two for double) per CTA -- {32 value bits | 32-bit exchange tag}: the flag travels with the data, so no
    // fence and no atomic is needed -- and the first warps poll the team's words.
    auto team_max = [&](R pm) -> R {
        constexpr int NW = sizeof(R) / 4;
        const int warp = tid >> 5, lane = tid & 31;
        unsigned long long bits = ord_bits(pm);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, bits, o); bits = x > bits ? x : bits; }
        if (CL || MC) {
            // Every WARP sends its maximum to slot [rank][warp] of every CTA of the cluster with st.async; the bytes are counted
            // by the destination's mbarrier, whose phase completes when thread 0's expect_tx and all csz x 8 words are in.
            // No CTA barrier, no cluster barrier, no fence.  Two exchanges of a cluster are always separated by a cluster barrier
            // that every thread passes after it has read cl_max (every phase boundary is one; between the last exchange of a
            // waveform and the first of the next one the waveform draw provides it), so one buffer and one barrier are enough --
            // without that separation a fast CTA's words of round r+1 could land in a slow CTA that still waits for round r.
            constexpr int NWARP = NT / 32;
            const unsigned par = xchg & 1u;
            ++xchg;
            if (tid == 0) mbar_expect_tx(&mb_max, csz * NWARP * 8u);
            if (lane < (int)csz) st_async_u64(&cl_max[crank * NWARP + warp], &mb_max, (unsigned)lane, bits);
            mbar_wait(&mb_max, par);
            unsigned long long m = 0ull;
            for (int i = lane; i < (int)csz * NWARP; i += 32) { const unsigned long long x = cl_max[i]; m = x > m ? x : m; }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, m, o); m = x > m ? x : m; }
            if (CL) return from_bits<R>(m);
            // MC: m is the maximum of my cluster; the cluster leaders publish it as self-validating words and warp 0 of every
            // CTA polls the ncl words of the team
            const unsigned long long tag = (unsigned long long)xchg;
            volatile unsigned long long* wf = a.slots + ((size_t)(team * 2 + (xchg & 1u)) * total) * 2;
            if (warp == 0) {
                if (crank == 0u && lane < NW) {
                    const unsigned long long part = (NW == 1) ? (m & 0xffffffffull) : (lane == 0 ? (m >> 32) : (m & 0xffffffffull));
                    wf[sub * 2 + lane] = (part << 32) | tag;
                }
                const int nwords = (int)ncl * NW;
                unsigned long long best = 0ull;
                for (int base = 0; base < nwords; base += 32) {
                    const int idx = base + lane;
                    const bool have = idx < nwords;
                    unsigned long long x = tag;
                    if (have) { for (;;) { x = wf[(idx / NW) * 2 + (idx % NW)]; if ((x & 0xffffffffull) == tag) break; } }
                    unsigned long long val = have ? (x >> 32) : 0ull;
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
            const unsigned long long all = red[0];
            __syncthreads();                                    // red[0] is free for the next exchange
            return from_bits<R>(all);
        }
        ++xchg;
        const unsigned long long tag = (unsigned long long)xchg;
        volatile unsigned long long* wf = a.slots + ((size_t)(team * 2 + (xchg & 1u)) * total) * 2;
        if (lane == 0) red[warp] = bits;
        __syncthreads();
        if (total == 1u) {                                      // a team of one CTA: the block maximum is the answer
            unsigned long long b = red[0];
#pragma unroll
            for (int i = 1; i < NT / 32; ++i) b = red[i] > b ? red[i] : b;
            __syncthreads();                                    // red[] is free for the next exchange
            return from_bits<R>(b);
        }
        const int nwords = (int)total * NW;
        unsigned long long best = 0ull;
        if (warp * 32 < nwords) {                               // the first warps publish (warp 0) and poll
            if (warp == 0) {
                bits = lane < NT / 32 ? red[lane] : 0ull;
#pragma unroll
                for (int o = 8; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, bits, o); bits = x > bits ? x : bits; }
                bits = __shfl_sync(0xffffffffu, bits, 0);
                if (lane < NW) {
                    const unsigned long long part = (NW == 1) ? (bits & 0xffffffffull) : (lane == 0 ? (bits >> 32) : (bits & 0xffffffffull));
                    wf[me * 2 + lane] = (part << 32) | tag;
                }
            }
            for (int base = 0; base < nwords; base += NT) {     // lanes pair up (hi, lo) for double
                const int idx = base + tid;
                const bool have = idx < nwords;
                unsigned long long x = tag;
                if (have) {
                    for (;;) {
                        x = wf[(idx / NW) * 2 + (idx % NW)];
                        if ((x & 0xffffffffull) == tag) break;
                        if (total > 32u) __nanosleep(20);
                    }
                }
                unsigned long long val = have ? (x >> 32) : 0ull;
                if (NW == 2) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, val, 1);
                    val = (tid & 1) ? 0ull : ((val << 32) | other);
                }
                best = val > best ? val : best;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(0xffffffffu, best, o); best = x > best ? x : best; }
        }
        __syncthreads();                                        // red[] of the block reduction has been consumed
        if (lane == 0) red[warp] = best;
        __syncthreads();
        best = red[0];
#pragma unroll
        for (int i = 1; i < NT / 32; ++i) best = red[i] > best ? red[i] : best;
        __syncthreads();                                        // red[] is free for the next exchange
        return from_bits<R>(best);
    };

    // A waveform that is skipped (already done under `resume`, zero length, no step budget) passes through no team barrier,
    // so the leader of a flag-based team could draw again and overwrite the mailbox word before a slower CTA has read the
    // current draw: synchronise the team once before the next draw (cluster teams: every draw has its own cluster barrier).
    auto grab_fence = [&]() {
        if (!CL) { bar_arrive(); bar_wait(); }
    };
    // streamed batches: this CTA's `ntiles` tiles of waveform w hold their final samples (every thread's stores are ordered before
    // the count by the CTA barrier and the system-scope fence of thread 0; the reader is a copy engine)
    auto mark_done = [&](unsigned int w, unsigned int ntiles) {
        if (!a.done) return;
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            atomicAdd(a.done + w / (unsigned int)a.chunk_rows, ntiles);
        }
    };
    const unsigned int my_tiles = MT ? (units - (unsigned)me + total - 1u) / total : 1u;
    // Separable linear operator (Params::lin_sep: complex128, beta_3 = 0, rows of 256 bins).  With the signed bin index
    // s = k1 + M1 k2' (k2' = k2 below M2/2, k2 - M2 above: the fftfreq wrap) the phase of devices.py:1145,1179 is
    //   imag(D~) h = theta s^2 = theta k1^2 + 2 theta M1 k1 k2' + theta M1^2 k2'^2,     theta = h c2 wscale^2,
    // so exp(j imag(D~) h) = A(k1) G(k1)^k2' B(|k2'|): B is a table of M2/2 + 1 entries that the CTA fills as soon as the step
    // size is known (before the team barrier, i.e. for free), and along a thread's 16 bins (k2' = tr + 16 q', q' = -8 .. 7) the
    // rest is a geometric sequence: two sincos and 15 complex products per thread instead of 16 loads of imag(D~) from L2 --
    // whose latency nothing hid, the cluster barrier having just emptied L1 -- and 16 sincos.  Differs from the tabulated
    // evaluation by rounding (~1e-13 rad on phases of ~1e3 rad); complex64 keeps the table: there the reference's own float32
    // rounding of the phase must be reproduced.
    auto lin_sep_table = [&](R hstep) {
        if constexpr (GEO::LSEP > 0) {
            if (p.lin_sep && tid < GEO::LSEP) {
                const double th = (double)hstep * (double)p.c2 * p.wscale * p.wscale;
                const double m = (double)tid * (double)M1;
                double sn, cs;
                sincos_r(th * m * m, sct, &sn, &cs);
                btab[tid] = mk<R>((R)cs, (R)sn);
            }
        }
    };
    auto lin_sep_apply = [&](C (&v)[E], R hstep, int row_k1) {
        if constexpr (GEO::LSEP > 0) {
            constexpr int ST = M2 / E;
            const double th = (double)hstep * (double)p.c2 * p.wscale * p.wscale;
            const double k1d = (double)row_k1;
            double sn, cs;
            sincos_r(th * k1d * (k1d + 2.0 * (double)M1 * (double)(tr - 8 * ST)), sct, &sn, &cs);   // A G^(tr - 8 ST)
            C wa = mk<R>((R)cs, (R)sn);
            sincos_r(th * 2.0 * (double)M1 * (double)ST * k1d, sct, &sn, &cs);                        // G^ST
            const C rho = mk<R>((R)cs, (R)sn);
            const C rho2 = cmul(rho, rho);
            C wb = cmul(wa, rho);
#pragma unroll
            for (int i = 0; i < E; i += 2) {                    // i <-> q' = i - 8 <-> register q = (i + 8) & 15
                const int m0 = i < 8 ? (8 - i) * ST - tr : (i - 8) * ST + tr;
                const int m1 = i + 1 < 8 ? (7 - i) * ST - tr : (i - 7) * ST + tr;
                v[(i + 8) & 15] = cmul(v[(i + 8) & 15], cmul(wa, btab[m0]));
                v[(i + 9) & 15] = cmul(v[(i + 9) & 15], cmul(wb, btab[m1]));
                if (i + 2 < E) { wa = cmul(wa, rho2); wb = cmul(wb, rho2); }
            }
        }
    };
    auto sh_read = [&](unsigned v) -> WfShared<R> {
        const volatile WfShared<R>* q = &sh[v & 1u];
        WfShared<R> r;
        r.w = q->w; r.xchg = q->xchg; r.steps = q->steps; r.taken = q->taken; r.z = q->z; r.h = q->h;
        return r;
    };

    if (tid == 0) { sh[0].xchg = 0u; sh[1].xchg = 0u; }        // (the draw below synchronises the CTA before anyone reads it)
    while (state != WF_END) {
        {
            if (state == WF_GRAB) {
                // ---------------------------------------------------------- next waveform of this team
                ++seq;
                if (CL) {                                       // CTA 0 of the cluster draws the waveform and posts it to every CTA
                    if (me == 0) {
                        if (tid == 0) {
                            // a team that is slower per waveform than the teams of the main launch leaves the last waveforms
                            // to them (a late draw would finish long after everybody else)
                            if (MT && a.draw_min > 0 && (int)(ld_relaxed_u32(a.next_wf) + (unsigned)a.draw_min) > p.batch) s_w = 0xffffffffu;
                            else s_w = atomicAdd(a.next_wf, 1u);
                            wait_arrival(a, s_w, p.batch);
                        }
                        __syncthreads();
                        if (tid < (int)total) st_cluster_u32(&s_wcl[seq & 1u], (unsigned)tid, s_w);
                    }
                    cluster_arrive_release();
                    cluster_wait_acquire();
                    if (tid == 0) s_w = s_wcl[seq & 1u];
                } else if (tid == 0) {
                    volatile unsigned long long* mb = a.mail + (size_t)team * 16;
                    unsigned int wn;
                    if (me == 0) {
                        wn = atomicAdd(a.next_wf, 1u);
                        wait_arrival(a, wn, p.batch);
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
                if (MC) {                                       // the draw of a multi-cluster team goes through the mailbox and is no
                    cluster_arrive_release();                   // barrier: separate the previous waveform's last max exchange from
                    cluster_wait_acquire();                     // this one's first inside every cluster (see team_max)
                }
                if (w >= (unsigned int)p.batch) { state = WF_END; continue; }
                xchg = ((const volatile WfShared<R>*)&sh[ver & 1u])->xchg;

                // ---------------------------------------------------------- prologue: first step size, first Kerr
                // half step (devices.py:1155-1161, 1175-1177), forward column transforms, four-step twiddle
                R z0 = 0, h_first;
                int steps0 = 0;
                if (a.resume) {
                    const Ctrl cs = p.ctrl[w];
                    if (cs.done) { mark_done(w, my_tiles); grab_fence(); continue; }    // stays in WF_GRAB: next waveform
                    z0 = (R)cs.z; h_first = (R)cs.h; steps0 = cs.steps;
                }
                if constexpr (MT) {
                    // several tiles per CTA
                    if (!a.resume) {
                        R h0 = a.fixed ? a.h_fixed : p.length;
                        if (!a.fixed && !a.single) {            // devices.py:1155-1156: max over the whole waveform, one more read of it
                            R pm = 0;
                            bool nan = false;
#pragma unroll 1
                            for (unsigned u = (unsigned)me; u < units; u += total) {
                                const int un2 = (int)(u % (unsigned)tiles) * T + c;
                                const C* __restrict__ inp = (p.field_in ? p.field_in : p.field) + ((size_t)w * p.n_pol + u / (unsigned)tiles) * NN;
                                C v[E];
#pragma unroll
                                for (int q = 0; q < E; ++q) v[q] = __ldcg(inp + (size_t)(t + q * (M1 / E)) * M2 + un2);
#pragma unroll
                                for (int q = 0; q < E; ++q) {
                                    const R pw = v[q].x * v[q].x + v[q].y * v[q].y;
                                    nan |= (pw != pw);
                                    pm = pw > pm ? pw : pm;
                                }
                            }
                            if (nan) pm = pw_nan<R>();
                            h0 = p.phi_max / mul_rn(p.abs_gamma, team_max(pm));
                        }
                        h_first = (p.length < h0) ? p.length : h0;
                        const int done0 = !((R)0 < p.length) || a.budget <= 0;
                        if (me == 0 && tid == 0) {
                            Ctrl& cs = p.ctrl[w];
                            cs.z = 0.0; cs.h = (double)h_first; cs.pmax = 0ull; cs.steps = 0; cs.arrived = 0u; cs.done = !((R)0 < p.length);
                        }
                        if (done0) { mark_done(w, my_tiles); grab_fence(); continue; }
                    }
                    if (tid == 0) {
                        WfShared<R>& o = sh[(ver ^ 1u) & 1u];
                        o.w = w; o.xchg = xchg; o.steps = steps0; o.taken = 0; o.z = z0; o.h = h_first;
                    }
                    ver ^= 1u;
                    const R hh = h_first / (R)2;
#pragma unroll 1
                    for (unsigned u = (unsigned)me; u < units; u += total) {
                        const int un2 = (int)(u % (unsigned)tiles) * T + c;
                        C* __restrict__ rowp = p.field + ((size_t)w * p.n_pol + u / (unsigned)tiles) * NN;
                        const C* __restrict__ inp = p.field_in ? p.field_in + ((size_t)w * p.n_pol + u / (unsigned)tiles) * NN : rowp;
                        R* __restrict__ ts = a.tstash + ((size_t)team * units + u) * (size_t)(E * NT);
                        C v[E];
#pragma unroll
                        for (int q = 0; q < E; ++q) v[q] = __ldcg(inp + (size_t)(t + q * (M1 / E)) * M2 + un2);
                        if (p.has_nl) {
#pragma unroll
                            for (int q = 0; q < E; ++q) {
                                const R pw = v[q].x * v[q].x + v[q].y * v[q].y;
                                const R ph = mul_rn(hh, mul_rn(p.gamma, pw));
                                ts[q * NT + tid] = ph;
                                R sn, co; kerr_sincos<SMALL>(ph, sct, &sn, &co);
                                v[q] = cmul(v[q], mk<R>(co, sn));
                            }
                        }
                        fft_passes<R, M1, -1, CX, E, 1, WFTW>::run(v, xb + c, tw1, t);
                        apply_fourstep<false, R, E, M1>(p, v, un2, t);
#pragma unroll
                        for (int q = 0; q < E; ++q) rowp[(size_t)(t + q * (M1 / E)) * M2 + un2] = v[q];
                    }
                    lin_sep_table(h_first);
                    bar_arrive();
                    state = WF_ROW;
                    continue;
                }
                C* __restrict__ rowp = p.field + ((size_t)w * p.n_pol + pol) * NN;
                const C* __restrict__ inp = p.field_in ? p.field_in + ((size_t)w * p.n_pol + pol) * NN : rowp;   // out-of-place transfer
                C v[E];
#pragma unroll
                for (int q = 0; q < E; ++q) v[q] = __ldcg(inp + (size_t)(t + q * (M1 / E)) * M2 + n2);
                if (!a.resume) {
                    R h0;
                    if (a.fixed) h0 = a.h_fixed;
                    else if (a.single) h0 = p.length;           // no dispersion or no Kerr effect: one step
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
                    h_first = (p.length < h0) ? p.length : h0;  // python min(h_, length)
                    const int done0 = !((R)0 < p.length) || a.budget <= 0;
                    if (me == 0 && tid == 0) {
                        Ctrl& cs = p.ctrl[w];
                        cs.z = 0.0; cs.h = (double)h_first; cs.pmax = 0ull; cs.steps = 0; cs.arrived = 0u; cs.done = !((R)0 < p.length);
                    }
                    if (done0) { mark_done(w, my_tiles); grab_fence(); continue; }
                }
                if (tid == 0) {                                  // (the last readers of this version are two team barriers back)
                    WfShared<R>& o = sh[(ver ^ 1u) & 1u];
                    o.w = w; o.xchg = xchg; o.steps = steps0; o.taken = 0; o.z = z0; o.h = h_first;
                }
                ver ^= 1u;
                if (p.has_nl) {
                    const R hh = h_first / (R)2;                // h_/2
#pragma unroll
                    for (int q = 0; q < E; ++q) {
                        const R pw = v[q].x * v[q].x + v[q].y * v[q].y; // |A|^2
                        const R ph = mul_rn(hh, mul_rn(p.gamma, pw));   // (h_/2) * (gamma |A|^2)
                        st_sm[q * NT + tid] = ph;
                        R sn, co; kerr_sincos<SMALL>(ph, sct, &sn, &co);
                        v[q] = cmul(v[q], mk<R>(co, sn));
                    }
                }
                fft_passes<R, M1, -1, CX, E, 1, WFTW>::run(v, xb + c, tw1, t);
                apply_fourstep<false, R, E, M1>(p, v, n2, t);
#pragma unroll
                for (int q = 0; q < E; ++q) rowp[(size_t)(t + q * (M1 / E)) * M2 + n2] = v[q];
                lin_sep_table(h_first);
                bar_arrive();
                state = WF_ROW;
                continue;
            }

            if (state == WF_ROW) {
                // ---------------------------------------------------------- row phase (devices.py:1178-1180)
                WF_T0(t_r0);
                bar_wait();
                WF_ACC(0, t_r0);
                const WfShared<R> S = sh_read(ver);
                if constexpr (MT) {
                    const R h = S.h;
#pragma unroll 1
                    for (unsigned u = (unsigned)me; u < units; u += total) {
                        const int uk1 = (int)(u % (unsigned)tiles) * G + g;
                        C* __restrict__ rbase = p.field + ((size_t)S.w * p.n_pol + u / (unsigned)tiles) * NN + (size_t)uk1 * M2;
                        C v[E];
#pragma unroll
                        for (int q = 0; q < E; ++q) v[q] = __ldcg(rbase + tr + q * (M2 / E));
                        fft_passes<R, M2, -1, RowExchange<M2, E>, E, 1, WFTW>::run(v, xb + g * PM, tw2, tr);
                        if (p.xfer) {
                            const C* __restrict__ hrow = p.xfer + (size_t)uk1 * M2;
#pragma unroll
                            for (int q = 0; q < E; ++q) v[q] = cmul(v[q], __ldg(hrow + tr + q * (M2 / E)));
                        } else if (GEO::LSEP > 0 && p.lin_sep) {
                            lin_sep_apply(v, h, uk1);
                        } else {
                            const R* __restrict__ drow = p.dim_tab + (size_t)uk1 * M2;
#pragma unroll
                            for (int q = 0; q < E; ++q) {
                                const R ph = mul_rn(__ldg(drow + tr + q * (M2 / E)), h);
                                R sn, co; sincos_r(ph, sct, &sn, &co);
                                v[q] = cmul(v[q], mk<R>(co, sn));
                            }
                        }
                        fft_passes<R, M2, +1, RowExchange<M2, E>, E, 1, WFTW>::run(v, xb + g * PM, tw2, tr);
#pragma unroll
                        for (int q = 0; q < E; ++q) rbase[tr + q * (M2 / E)] = v[q];
                    }
                    bar_arrive();
                    state = WF_COL;
                    WF_ACC(1, t_r0);
                    continue;
                }
                C* __restrict__ rbase = p.field + ((size_t)S.w * p.n_pol + pol) * NN + (size_t)k1 * M2;
                const R h = S.h;
                C v[E];
#pragma unroll
                for (int q = 0; q < E; ++q) v[q] = __ldcg(rbase + tr + q * (M2 / E));
                fft_passes<R, M2, -1, RowExchange<M2, E>, E, 1, WFTW>::run(v, xb + g * PM, tw2, tr);
                if (p.xfer) {                                   // an arbitrary transfer function H[k] in transposed order: zero-phase
                    const C* __restrict__ hrow = p.xfer + (size_t)k1 * M2;      // filters (|H|^2), DM, FBG -- one pass over the rows,
#pragma unroll                                                  // FFT -> x H -> IFFT with the waveforms in flight L2-resident
                    for (int q = 0; q < E; ++q) v[q] = cmul(v[q], __ldg(hrow + tr + q * (M2 / E)));
                } else if (GEO::LSEP > 0 && p.lin_sep) {
                    lin_sep_apply(v, h, k1);
                } else {
                    const R* __restrict__ drow = p.dim_tab + (size_t)k1 * M2;   // imag(D~) of my row's bins (k_fill_dim)
#pragma unroll
                    for (int q = 0; q < E; ++q) {
                        const R ph = mul_rn(__ldg(drow + tr + q * (M2 / E)), h);
                        R sn, co; sincos_r(ph, sct, &sn, &co);
                        v[q] = cmul(v[q], mk<R>(co, sn));
                    }
                }
                fft_passes<R, M2, +1, RowExchange<M2, E>, E, 1, WFTW>::run(v, xb + g * PM, tw2, tr);
#pragma unroll
                for (int q = 0; q < E; ++q) rbase[tr + q * (M2 / E)] = v[q];
                bar_arrive();
                state = WF_COL;
                WF_ACC(1, t_r0);
                continue;
            }

            // -------------------------------------------------------------- column phase: end of step s ...
            WF_T0(t_c0);
            FsSeeds<R> fs_seed;                                 // (MT: several tiles, seeds per tile inside the loop)
            if (!MT && p.tw_chain) fs_seed = fourstep_seeds<R, E, M1>(p, n2, t);   // their L2 round trip hides behind the barrier wait
            bar_wait();
            WF_ACC(2, t_c0);
            const WfShared<R> S = sh_read(ver);
            if constexpr (MT) {
                const R z = S.z, h = S.h;
                const int steps = S.steps;
                xchg = S.xchg;
                const R sc = p.inv_n * exp_r(mul_rn(p.att_half, h));
                R pmax = 0;
                if (p.adaptive) {
                    // Adaptive step control needs max|A|^2 over the WHOLE waveform before any tile's Kerr rotation: a first pass
                    // over my tiles ends step s (inverse transforms, 1/N, attenuation, maximum) and stores the time-domain tile,
                    // the second pass below re-reads it (each thread its own samples) for the merged rotation and the forward
                    // transforms -- one more read and write of the field through L2 per step than the fixed-step schedule.
                    R pm = 0;
                    bool nan = false;
#pragma unroll 1
                    for (unsigned u = (unsigned)me; u < units; u += total) {
                        const int un2 = (int)(u % (unsigned)tiles) * T + c;
                        C* __restrict__ rowp = p.field + ((size_t)S.w * p.n_pol + u / (unsigned)tiles) * NN;
                        C v[E];
#pragma unroll
                        for (int q = 0; q < E; ++q) v[q] = __ldcg(rowp + (size_t)(t + q * (M1 / E)) * M2 + un2);
                        apply_fourstep<true, R, E, M1>(p, v, un2, t);
                        fft_passes<R, M1, +1, CX, E, 1, WFTW>::run(v, xb + c, tw1, t);
#pragma unroll
                        for (int q = 0; q < E; ++q) {
                            v[q].x *= sc; v[q].y *= sc;
                            const R pw = v[q].x * v[q].x + v[q].y * v[q].y;
                            nan |= (pw != pw);
                            pm = pw > pm ? pw : pm;
                            rowp[(size_t)(t + q * (M1 / E)) * M2 + un2] = v[q];
                        }
                    }
                    if (nan) pm = pw_nan<R>();
                    WF_T0(t_x0);
                    pmax = team_max(pm);
                    WF_ACC(3, t_x0);
                }
                const bool two_pass = p.adaptive != 0;
                const CtrlNext<R> nx = controller_next<R>(p, z, h, steps, pmax);
                const long long taken = S.taken + 1;
#ifdef SSFM_WF_PROFILE
                ++pn;
#endif
                const bool stop = nx.done || taken >= a.budget;
                if (me == 0 && tid == 0) {
                    Ctrl& cs = p.ctrl[S.w];
                    if (p.hlog && steps < p.hlog_cap) p.hlog[(size_t)S.w * p.hlog_cap + steps] = (double)h;
                    cs.z = (double)nx.z; cs.h = (double)nx.h; cs.steps = steps + 1; cs.done = nx.done;
                }
                const R hh = nx.h / (R)2;
#pragma unroll 1
                for (unsigned u = (unsigned)me; u < units; u += total) {
                    const int un2 = (int)(u % (unsigned)tiles) * T + c;
                    C* __restrict__ rowp = p.field + ((size_t)S.w * p.n_pol + u / (unsigned)tiles) * NN;
                    R* __restrict__ ts = a.tstash + ((size_t)team * units + u) * (size_t)(E * NT);
                    C v[E];
#pragma unroll
                    for (int q = 0; q < E; ++q) v[q] = __ldcg(rowp + (size_t)(t + q * (M1 / E)) * M2 + un2);
                    if (p.has_nl) {                                 // the tile's Kerr phase: global -> shared behind the inverse transforms
#pragma unroll                                                      // (every thread copies and reads its own entries only)
                        for (int q = 0; q < E; ++q) cp_async<sizeof(R)>(&st_sm[q * NT + tid], ts + q * NT + tid);
                        cp_async_commit();
                    }
                    if (!two_pass) {
                        apply_fourstep<true, R, E, M1>(p, v, un2, t);
                        fft_passes<R, M1, +1, CX, E, 1, WFTW>::run(v, xb + c, tw1, t);
#pragma unroll
                        for (int q = 0; q < E; ++q) { v[q].x *= sc; v[q].y *= sc; }
                    }
                    if (p.has_nl) cp_async_wait_all();
                    if (stop) {
#pragma unroll
                        for (int q = 0; q < E; ++q) {
                            if (p.has_nl) {
                                R sn, co; kerr_sincos<SMALL>(st_sm[q * NT + tid], sct, &sn, &co);
                                v[q] = cmul(v[q], mk<R>(co, sn));
                            }
                            rowp[(size_t)(t + q * (M1 / E)) * M2 + un2] = v[q];
                        }
                        continue;
                    }
                    if (p.has_nl) {
#pragma unroll
                        for (int q = 0; q < E; ++q) {
                            const R pw = v[q].x * v[q].x + v[q].y * v[q].y;
                            const R ph = mul_rn(hh, mul_rn(p.gamma, pw));
                            const R tot = st_sm[q * NT + tid] + ph;
                            ts[q * NT + tid] = ph;
                            R sn, co; kerr_sincos<SMALL>(tot, sct, &sn, &co);
                            v[q] = cmul(v[q], mk<R>(co, sn));
                        }
                    }
                    fft_passes<R, M1, -1, CX, E, 1, WFTW>::run(v, xb + c, tw1, t);
                    apply_fourstep<false, R, E, M1>(p, v, un2, t);
#pragma unroll
                    for (int q = 0; q < E; ++q) rowp[(size_t)(t + q * (M1 / E)) * M2 + un2] = v[q];
                }
                if (stop) {
                    if (tid == 0) sh[ver & 1u].xchg = xchg;
                    mark_done(S.w, my_tiles);
                    state = WF_GRAB;
                    WF_ACC(4, t_c0);
                    continue;
                }
                if (tid == 0) {
                    WfShared<R>& o = sh[(ver ^ 1u) & 1u];
                    o.w = S.w; o.xchg = xchg; o.steps = steps + 1; o.taken = taken; o.z = nx.z; o.h = nx.h;
                }
                ver ^= 1u;
                lin_sep_table(nx.h);
                bar_arrive();
                state = WF_ROW;
                WF_ACC(4, t_c0);
                continue;
            }
            C* __restrict__ rowp = p.field + ((size_t)S.w * p.n_pol + pol) * NN;
            const R z = S.z, h = S.h;
            const int steps = S.steps;
            xchg = S.xchg;
            C v[E];
#pragma unroll
            for (int q = 0; q < E; ++q) v[q] = __ldcg(rowp + (size_t)(t + q * (M1 / E)) * M2 + n2);
            if (p.tw_chain) apply_fourstep_chain<true, R, E>(v, fs_seed);
            else apply_fourstep<true, R, E, M1>(p, v, n2, t);
            fft_passes<R, M1, +1, CX, E, 1, WFTW>::run(v, xb + c, tw1, t);
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
                WF_T0(t_x0);
                pmax = team_max(pm);
                WF_ACC(3, t_x0);
            }
            const CtrlNext<R> nx = controller_next<R>(p, z, h, steps, pmax);
            const long long taken = S.taken + 1;
#ifdef SSFM_WF_PROFILE
            ++pn;
#endif
            const bool stop = nx.done || taken >= a.budget;
            if (me == 0 && tid == 0) {
                Ctrl& cs = p.ctrl[S.w];
                if (p.hlog && steps < p.hlog_cap) p.hlog[(size_t)S.w * p.hlog_cap + steps] = (double)h;
                cs.z = (double)nx.z; cs.h = (double)nx.h; cs.steps = steps + 1; cs.done = nx.done;
            }
            if (stop) {                                         // second Kerr half step, time domain out
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    if (p.has_nl) {
                        R sn, co; kerr_sincos<SMALL>(st_sm[q * NT + tid], sct, &sn, &co);
                        v[q] = cmul(v[q], mk<R>(co, sn));
                    }
                    rowp[(size_t)(t + q * (M1 / E)) * M2 + n2] = v[q];
                }
                if (tid == 0) sh[ver & 1u].xchg = xchg;         // (adaptive mode: every warp read this version before its exchange)
                mark_done(S.w, my_tiles);
                state = WF_GRAB;
                WF_ACC(4, t_c0);
                continue;
            }
            // -------------------------------------------------------------- ... and start of step s+1
            if (tid == 0) {                                     // next version of the state (read after the next team barrier)
                WfShared<R>& o = sh[(ver ^ 1u) & 1u];
                o.w = S.w; o.xchg = xchg; o.steps = steps + 1; o.taken = taken; o.z = nx.z; o.h = nx.h;
            }
            ver ^= 1u;
            if (p.has_nl) {
                const R hh = nx.h / (R)2;
#pragma unroll
                for (int q = 0; q < E; ++q) {
                    const R pw = v[q].x * v[q].x + v[q].y * v[q].y;
                    const R ph = mul_rn(hh, mul_rn(p.gamma, pw));   // first half step of the next step
                    const R tot = st_sm[q * NT + tid] + ph;         // + second half step of this one
                    st_sm[q * NT + tid] = ph;                       // (each thread re-reads only what it wrote itself)
                    R sn, co; kerr_sincos<SMALL>(tot, sct, &sn, &co);
                    v[q] = cmul(v[q], mk<R>(co, sn));
                }
            }
            fft_passes<R, M1, -1, CX, E, 1, WFTW>::run(v, xb + c, tw1, t);
            if (p.tw_chain) apply_fourstep_chain<false, R, E>(v, fs_seed);   // (the seeds of the phase's first twiddle: a second look-up
            else apply_fourstep<false, R, E, M1>(p, v, n2, t);               //  would miss L1, which the barrier wait emptied)
#pragma unroll
            for (int q = 0; q < E; ++q) rowp[(size_t)(t + q * (M1 / E)) * M2 + n2] = v[q];
            lin_sep_table(nx.h);
            bar_arrive();
            state = WF_ROW;
            WF_ACC(4, t_c0);
        }
    }
#ifdef SSFM_WF_PROFILE
    if (tid == 0 && me == 1 && pn > 0)
        printf("[k_wf profile] team %d steps %lld cycles/step: row wait %lld | row phase %lld | col wait %lld | exchange %lld | col phase %lld\n",
               team, pn, prof[0] / pn, (prof[1] - prof[0]) / pn, prof[2] / pn, prof[3] / pn, (prof[4] - prof[2]) / pn);
#endif
}

}  // namespace ssfm
