// Launch side of the persistent whole-propagation kernel k_wf (ssfm_wf.cuh).  Included by ssfm_wf_f32.cu and ssfm_wf_f64.cu
// (one translation unit per precision: the instantiations compile in parallel), with SSFM_WF_REAL = float | double.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/ssfm_b200.h"
#include "ssfm_wf.cuh"
#include "ssfm_wf.h"
#include "ssfm_internal.h"

extern std::atomic<long long> ssfm_launches;

namespace ssfm {
namespace {

int wf_fail(int code, const std::string& msg) { ssfm_err_slot = msg; return code; }

#define WF_TRY(expr)                                                                            \
    do {                                                                                        \
        cudaError_t e__ = (expr);                                                               \
        if (e__ != cudaSuccess)                                                                 \
            return wf_fail(SSFM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

// cluster variant: one thread-block cluster per team (teams of <= 16 CTAs).  16-CTA clusters must sit inside one GPC, so
// fewer of them are co-resident than the chip has CTA slots (measured on B200: 14 of 18 possible teams in fp64, 21 of 27 in
// fp32); the slots they leave are filled by a second launch of the flag-based variant on a side stream, whose teams draw
// from the SAME waveform counter.  (That launch needs no co-residency guarantee: if some of its CTAs are not scheduled
// at once its teams simply wait until the cluster kernel frees slots.)
template <typename R, int M1, int M2, bool SMALL>
int wf_launch_cluster(const Params<R>& p, const WfLaunch& l, int* teams_out, cudaStream_t st, int coop_ctas) {
    typedef wf_geom<R, M1, M2> GEO;
    auto kern = k_wf<R, M1, M2, SMALL, 1>;
    const int total = p.n_pol * (p.n2 / GEO::T);
    static int max_clusters_dev[64][17] = {{0}};                 // per instantiation, device and cluster size
    int dev = 0;
    cudaGetDevice(&dev);
    int (&max_clusters)[17] = max_clusters_dev[(dev >= 0 && dev < 64) ? dev : 0];
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(GEO::NT); cfg.dynamicSmemBytes = GEO::smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)total; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (const char* pol = getenv("SSFM_CLUSTER_POLICY")) {      // experiments: 1 = spread, 2 = load balancing
        at[1].id = cudaLaunchAttributeClusterSchedulingPolicyPreference;
        at[1].val.clusterSchedulingPolicyPreference = (cudaClusterSchedulingPolicy)atoi(pol);
        cfg.numAttrs = 2;
    }
    if (max_clusters[total] == 0) {
        WF_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEO::smem));
        WF_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cfg.gridDim = dim3((unsigned)(total * 64));
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
        max_clusters[total] = n > 0 ? n : -1;
    }
    long long teams = max_clusters[total];
    if (getenv("SSFM_DEBUG"))
        fprintf(stderr, "[ssfm] k_wf cluster<%d,%d,%d>: %d CTAs per cluster, %lld clusters fit (cooperative: %d CTAs)\n", (int)sizeof(R),
                M1, M2, total, teams, coop_ctas);
    if (teams < 1) return SSFM_ERR_UNSUPPORTED;
    long long fill = (l.side && l.ev_side && l.ev0) ? (coop_ctas - teams * total) / total : 0;   // flag-based teams in the free slots
    if (l.cluster < 1 && (teams + fill) * total * 10 < (long long)coop_ctas * 8) return SSFM_ERR_UNSUPPORTED;   // < 80 % of the chip
    if (teams > p.batch) teams = p.batch;
    if (l.teams_cap > 0 && teams > l.teams_cap) teams = l.teams_cap;
    if (const char* e = getenv("SSFM_CL_CAP")) { const long long c = atoll(e); if (c > 0 && teams > c) teams = c; }   // experiments
    if (fill > p.batch - teams) fill = p.batch - teams;
    if (l.teams_cap > 0 && fill > l.teams_cap - teams) fill = l.teams_cap - teams;
    const size_t head = 4096 + 256;
    const size_t need = head + (size_t)fill * 256 + (size_t)fill * 2 * total * 16;
    if (need > WF_SYNC_BYTES) fill = 0;
    WF_TRY(cudaMemsetAsync(l.sync_buf, 0, fill > 0 ? need : head, st));
    WfArgs<R> a;
    std::memset(&a, 0, sizeof(a));
    char* sb = (char*)l.sync_buf;
    a.next_wf = (unsigned int*)(sb + 4096 + 128);
    a.budget = l.budget;
    a.n_teams = (int)teams;
    a.fixed = l.fixed; a.single = l.single; a.resume = l.resume;
    a.h_fixed = (R)l.h_fixed;
    a.ready = l.ready; a.done = l.done; a.chunk_rows = l.chunk_rows;
    a.occ = 1; a.placement = 0;
    cfg.gridDim = dim3((unsigned)(teams * total));
    if (l.ev0) WF_TRY(cudaEventRecord(l.ev0, st));
    WF_TRY(cudaLaunchKernelEx(&cfg, kern, p, a));
    ++ssfm_launches;
    // The slots the 16-CTA clusters cannot use: small clusters (2 CTAs by default) that carry a waveform with several tiles per
    // CTA (k_wf<.., TM = 3>) -- hardware barrier, no flags through L2, barrier and exchange waits amortised over the tiles of a
    // phase -- when the team stash is there; else flag-based teams of `total` CTAs.
    int fcs = 2;
    if (const char* e = getenv("SSFM_FILL_CS")) fcs = atoi(e);
    const long long free_slots = (l.side && l.ev_side && l.ev0) ? coop_ctas - teams * total : 0;
    const size_t per_team = (size_t)total * 4096 * sizeof(R);
    // (measured on B200, config #3 in fp64: 5.77e10 with clusters of 2 against 5.33e10 with flag-based fill teams, 5.46e10 with
    // clusters of 4, 4.78e10 with single CTAs; a waveform takes such a team ~6x as long as a 16-CTA cluster: the small clusters stop
    // drawing when fewer than teams x slow waveforms are left (WfArgs::draw_min), and batches below twice that -- where they would
    // hardly draw at all -- keep the flag-based teams; the 256-row chunks of the host pipeline are above it)
    const long long slow = ((long long)(total / (fcs > 0 ? fcs : 1)) * 8 + 9) / 10;
    if (fcs >= 1 && fcs <= 8 && total % fcs == 0 && total > fcs && (l.tstash || !p.has_nl) && free_slots >= fcs &&
        p.batch >= (getenv("SSFM_FILL_MIN") ? atoll(getenv("SSFM_FILL_MIN")) : 2 * teams * slow)) {
        auto kf = k_wf<R, M1, M2, SMALL, 3>;
        static bool attr_done[64] = {false};
        if (!attr_done[(dev >= 0 && dev < 64) ? dev : 0]) {
            WF_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEO::smem));
            attr_done[(dev >= 0 && dev < 64) ? dev : 0] = true;
        }
        long long ft = free_slots / fcs;
        if (p.has_nl && ft > (long long)(l.tstash_bytes / per_team)) ft = (long long)(l.tstash_bytes / per_team);
        if (ft > p.batch - teams) ft = p.batch - teams;
        if (l.teams_cap > 0 && ft > l.teams_cap - teams) ft = l.teams_cap - teams;
        if (ft > 0) {
            WfArgs<R> b = a;
            b.n_teams = (int)ft;
            b.tstash = (R*)l.tstash;
            // a waveform takes a team of fcs CTAs about total/fcs x 0.8 times as long as a 16-CTA cluster (`slow` above)
            b.draw_min = (int)(teams * ((total / fcs) * 8 + 9) / 10);
            cudaLaunchConfig_t fc{};
            fc.blockDim = dim3(GEO::NT); fc.dynamicSmemBytes = GEO::smem; fc.stream = l.side;
            cudaLaunchAttribute fa[1];
            fa[0].id = cudaLaunchAttributeClusterDimension;
            fa[0].val.clusterDim.x = (unsigned)fcs; fa[0].val.clusterDim.y = 1; fa[0].val.clusterDim.z = 1;
            fc.attrs = fa; fc.numAttrs = 1;
            fc.gridDim = dim3((unsigned)(ft * fcs));
            WF_TRY(cudaStreamWaitEvent(l.side, l.ev0, 0));
            WF_TRY(cudaLaunchKernelEx(&fc, kf, p, b));
            WF_TRY(cudaEventRecord(l.ev_side, l.side));
            WF_TRY(cudaStreamWaitEvent(st, l.ev_side, 0));
            ++ssfm_launches;
            fill = ft;
        } else fill = 0;
    } else if (fill > 0) {
        WfArgs<R> b = a;
        b.sm_cnt = (unsigned int*)sb;
        b.grid_bar = (unsigned int*)(sb + 4096);
        b.bar = (unsigned int*)(sb + head);
        b.mail = (unsigned long long*)(sb + head + (size_t)fill * 128);
        b.slots = (unsigned long long*)(sb + head + (size_t)fill * 256);
        b.n_teams = (int)fill;
        WF_TRY(cudaStreamWaitEvent(l.side, l.ev0, 0));
        k_wf<R, M1, M2, SMALL, 0><<<(unsigned)(fill * total), GEO::NT, GEO::smem, l.side>>>(p, b);
        WF_TRY(cudaGetLastError());
        WF_TRY(cudaEventRecord(l.ev_side, l.side));
        WF_TRY(cudaStreamWaitEvent(st, l.ev_side, 0));
        ++ssfm_launches;
    }
    if (l.ev1) WF_TRY(cudaEventRecord(l.ev1, st));
    if (teams_out) *teams_out = (int)(teams + fill);
    return SSFM_OK;
}

// cluster teams with SEVERAL tiles per CTA (waveforms of 32 .. 64 tiles: 2^17, 2^18 samples, or 2^16 / 2^17 with two polarisations;
// fixed or adaptive steps -- the latter with two passes per column phase -- or a transfer function): a team is ONE cluster of 16 CTAs whatever the
// waveform length, every CTA carries units/16 tiles through each phase, and the team barrier is the hardware cluster barrier
// instead of 64 arrivals on a counter in L2 (measured on B200: 64-CTA flag-based teams spend ~40 % of a step in their two
// barriers).  The Kerr phase of the tiles in flight lives in a small L2-resident buffer (WfLaunch::tstash) instead of shared
// memory.  The CTA slots that 16-CTA clusters cannot use (a cluster needs 16 SMs of one GPC: 14 clusters fit) are filled by one
// launch of flag-based teams drawing from the same waveform counter, as in wf_launch_cluster.
template <typename R, int M1, int M2, bool SMALL>
int wf_launch_mt(const Params<R>& p, const WfLaunch& l, int* teams_out, cudaStream_t st, int coop_ctas) {
    typedef wf_geom<R, M1, M2> GEO;
    int CS = 16;
    if (const char* e = getenv("SSFM_MT_CS")) CS = atoi(e);        // experiments: clusters of 2 / 4 / 8
    if (CS < 1 || CS > 16 || (CS & (CS - 1))) CS = 16;
    auto kern = k_wf<R, M1, M2, SMALL, 3>;
    const long long units = (long long)p.n_pol * (p.n2 / GEO::T);
    if (units <= CS || units % CS || (p.has_nl && !l.tstash)) return SSFM_ERR_UNSUPPORTED;
    static int max_clusters_dev[64][17];
    static bool init = false;
    if (!init) { for (auto& r : max_clusters_dev) for (int& v : r) v = 0; init = true; }
    int dev = 0;
    cudaGetDevice(&dev);
    int& max_clusters = max_clusters_dev[(dev >= 0 && dev < 64) ? dev : 0][CS];
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(GEO::NT); cfg.dynamicSmemBytes = GEO::smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (max_clusters == 0) {
        WF_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEO::smem));
        WF_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cfg.gridDim = dim3((unsigned)(CS * 64));
        int n = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
        max_clusters = n > 0 ? n : -1;
    }
    long long teams = max_clusters;
    const size_t per_team = (size_t)units * 4096 * sizeof(R);
    if (p.has_nl && teams > (long long)(l.tstash_bytes / per_team)) teams = (long long)(l.tstash_bytes / per_team);
    if (getenv("SSFM_DEBUG"))
        fprintf(stderr, "[ssfm] k_wf multi-tile cluster<%d,%d,%d>: %lld tiles per waveform, %d per CTA, %lld clusters\n", (int)sizeof(R), M1, M2,
                units, (int)(units / CS), teams);
    if (teams < 1) return SSFM_ERR_UNSUPPORTED;
    // small batches keep more of the chip busy with flag-based teams (units CTAs per waveform instead of 16)
    if (l.cluster < 1 && p.batch * CS * 14 < (long long)coop_ctas * 10 && p.batch < teams) return SSFM_ERR_UNSUPPORTED;
    if (getenv("SSFM_DEBUG")) fprintf(stderr, "[ssfm] multi-tile: %lld teams of %d CTAs\n", teams, CS);
    long long fill = (l.side && l.ev_side && l.ev0) ? (coop_ctas - teams * CS) / units : 0;   // flag-based teams in the free slots
    if (teams > p.batch) teams = p.batch;
    if (l.teams_cap > 0 && teams > l.teams_cap) teams = l.teams_cap;
    if (fill > p.batch - teams) fill = p.batch - teams;
    if (l.teams_cap > 0 && fill > l.teams_cap - teams) fill = l.teams_cap - teams;
    if (getenv("SSFM_MT_NOFILL")) fill = 0;
    const size_t head = 4096 + 256;
    const size_t need = head + (size_t)fill * 256 + (size_t)fill * 2 * units * 16;
    if (need > WF_SYNC_BYTES) fill = 0;
    WF_TRY(cudaMemsetAsync(l.sync_buf, 0, fill > 0 ? need : head, st));
    WfArgs<R> a;
    std::memset(&a, 0, sizeof(a));
    char* sb = (char*)l.sync_buf;
    a.next_wf = (unsigned int*)(sb + 4096 + 128);
    a.budget = l.budget;
    a.n_teams = (int)teams;
    a.fixed = l.fixed; a.single = l.single; a.resume = l.resume;
    a.h_fixed = (R)l.h_fixed;
    a.ready = l.ready; a.done = l.done; a.chunk_rows = l.chunk_rows;
    a.occ = 1; a.placement = 0;
    a.tstash = (R*)l.tstash;
    cfg.gridDim = dim3((unsigned)(teams * CS));
    if (l.ev0) WF_TRY(cudaEventRecord(l.ev0, st));
    WF_TRY(cudaLaunchKernelEx(&cfg, kern, p, a));
    ++ssfm_launches;
    if (fill > 0) {
        auto kfill = k_wf<R, M1, M2, SMALL, 0>;
        static bool fill_attr[64] = {false};
        if (!fill_attr[(dev >= 0 && dev < 64) ? dev : 0]) {
            WF_TRY(cudaFuncSetAttribute(kfill, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEO::smem));
            fill_attr[(dev >= 0 && dev < 64) ? dev : 0] = true;
        }
        WfArgs<R> b = a;
        b.tstash = nullptr;
        b.sm_cnt = (unsigned int*)sb;
        b.grid_bar = (unsigned int*)(sb + 4096);
        b.bar = (unsigned int*)(sb + head);
        b.mail = (unsigned long long*)(sb + head + (size_t)fill * 128);
        b.slots = (unsigned long long*)(sb + head + (size_t)fill * 256);
        b.n_teams = (int)fill;
        WF_TRY(cudaStreamWaitEvent(l.side, l.ev0, 0));
        kfill<<<(unsigned)(fill * units), GEO::NT, GEO::smem, l.side>>>(p, b);
        WF_TRY(cudaGetLastError());
        WF_TRY(cudaEventRecord(l.ev_side, l.side));
        WF_TRY(cudaStreamWaitEvent(st, l.ev_side, 0));
        ++ssfm_launches;
    }
    if (l.ev1) WF_TRY(cudaEventRecord(l.ev1, st));
    if (teams_out) *teams_out = (int)(teams + fill);
    return SSFM_OK;
}

// multi-cluster teams (32 .. 256 CTAs): clusters of 8 CTAs, hardware barrier inside a cluster, one flag hop between the cluster
// leaders.  All clusters of a team must be resident at once, so the launch is cooperative and no larger than the number of
// clusters that fit on the chip (33 clusters of 8 with two 256-thread CTAs per SM on a B200: four teams of 2^18 samples, one of
// 2^20).
template <typename R, int M1, int M2, bool SMALL>
int wf_launch_mc(const Params<R>& p, const WfLaunch& l, int* teams_out, cudaStream_t st) {
    typedef wf_geom<R, M1, M2> GEO;
    constexpr int CS = 8;
    auto kern = k_wf<R, M1, M2, SMALL, 2>;
    const long long total = (long long)p.n_pol * (p.n2 / GEO::T);
    if (total <= 16 || total % CS) return SSFM_ERR_UNSUPPORTED;
    static int max_clusters_dev[64];
    static bool init = false;
    if (!init) { for (int& v : max_clusters_dev) v = 0; init = true; }
    int dev = 0;
    cudaGetDevice(&dev);
    int& max_clusters = max_clusters_dev[(dev >= 0 && dev < 64) ? dev : 0];
    cudaLaunchConfig_t cfg{};
    cfg.blockDim = dim3(GEO::NT); cfg.dynamicSmemBytes = GEO::smem; cfg.stream = st;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeCooperative;
    at[1].val.cooperative = 1;
    cfg.attrs = at; cfg.numAttrs = 2;
    if (max_clusters == 0) {
        WF_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEO::smem));
        cfg.gridDim = dim3((unsigned)(CS * 64));
        int n = 0;
        cfg.numAttrs = 1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        cfg.numAttrs = 2;
        if (e != cudaSuccess) { (void)cudaGetLastError(); n = 0; }
        max_clusters = n > 0 ? n : -1;
    }
    const long long ncl = total / CS;
    long long teams = max_clusters / ncl;
    if (getenv("SSFM_DEBUG"))
        fprintf(stderr, "[ssfm] k_wf multi-cluster<%d,%d,%d>: %lld clusters of %d per team, %d clusters fit -> %lld teams\n", (int)sizeof(R), M1, M2,
                ncl, CS, max_clusters, teams);
    if (teams < 1) return SSFM_ERR_UNSUPPORTED;
    if (teams > p.batch) teams = p.batch;
    if (l.teams_cap > 0 && teams > l.teams_cap) teams = l.teams_cap;
    const size_t head = 4096 + 256;
    const size_t ts = (size_t)teams;
    const size_t need = head + ts * 256 + ts * 2 * total * 16;
    if (need > WF_SYNC_BYTES) return SSFM_ERR_UNSUPPORTED;
    WF_TRY(cudaMemsetAsync(l.sync_buf, 0, need, st));
    WfArgs<R> a;
    std::memset(&a, 0, sizeof(a));
    char* sb = (char*)l.sync_buf;
    a.next_wf = (unsigned int*)(sb + 4096 + 128);
    a.bar = (unsigned int*)(sb + head);
    a.mail = (unsigned long long*)(sb + head + ts * 128);
    a.slots = (unsigned long long*)(sb + head + ts * 256);
    a.budget = l.budget;
    a.n_teams = (int)teams;
    a.fixed = l.fixed; a.single = l.single; a.resume = l.resume;
    a.h_fixed = (R)l.h_fixed;
    a.ready = l.ready; a.done = l.done; a.chunk_rows = l.chunk_rows;
    a.occ = 1; a.placement = 0;
    cfg.gridDim = dim3((unsigned)(teams * total));
    if (l.ev0) WF_TRY(cudaEventRecord(l.ev0, st));
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p, a);
    if (e != cudaSuccess) {                                  // e.g. cooperative + cluster launch refused: the flag-based teams take over
        (void)cudaGetLastError();
        if (getenv("SSFM_DEBUG")) fprintf(stderr, "[ssfm] multi-cluster launch refused: %s\n", cudaGetErrorString(e));
        return SSFM_ERR_UNSUPPORTED;
    }
    ++ssfm_launches;
    if (l.ev1) WF_TRY(cudaEventRecord(l.ev1, st));
    if (teams_out) *teams_out = (int)teams;
    return SSFM_OK;
}

template <typename R, int M1, int M2, bool SMALL>
int wf_launch(const Params<R>& p, const WfLaunch& l, int* teams_out, cudaStream_t st) {
    typedef wf_geom<R, M1, M2> GEO;
    auto kern = k_wf<R, M1, M2, SMALL, 0>;
    static int per_sm_dev[64];                                  // per instantiation and device (function attributes are per device)
    static bool per_sm_init = false;
    if (!per_sm_init) { for (int& v : per_sm_dev) v = -1; per_sm_init = true; }
    int dev0 = 0;
    cudaGetDevice(&dev0);
    int& per_sm = per_sm_dev[(dev0 >= 0 && dev0 < 64) ? dev0 : 0];
    if (per_sm < 0) {
        WF_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GEO::smem));
        int v = 0;
        WF_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, GEO::NT, GEO::smem));
        per_sm = v;
    }
    const long long total = (long long)p.n_pol * (p.n2 / GEO::T);       // CTAs of one team
    // placement of k_wf: teams no larger than the SM count share groups of `total` SMs (one CTA of each per SM)
    const int placement = l.placement >= 0 ? l.placement : (sizeof(R) == 4 ? 1 : 0);
    long long teams = (placement && total <= l.num_sms) ? (l.num_sms / total) * per_sm : (long long)per_sm * l.num_sms / total;
    if (teams > p.batch) teams = p.batch;
    if (l.teams_cap > 0 && teams > l.teams_cap) teams = l.teams_cap;
    if (teams < 1) return SSFM_ERR_UNSUPPORTED;                          // one waveform does not fit on the chip
    if constexpr (M1 * M2 == (1 << 16)) {                                // experiments: 16-tile waveforms on multi-tile clusters of SSFM_MT_CS CTAs
        if (getenv("SSFM_MT_CS") && l.cluster != 0 && total == 16) {
            const int rc = wf_launch_mt<R, M1, M2, SMALL>(p, l, teams_out, st, per_sm * l.num_sms);
            if (rc != SSFM_ERR_UNSUPPORTED) return rc;
        }
    }
    if constexpr (M2 <= 256) {                                           // teams of <= 16 CTAs can be thread-block clusters
        if (l.cluster != 0 && total <= 16 && total >= 2 && (total & (total - 1)) == 0) {
            const int rc = wf_launch_cluster<R, M1, M2, SMALL>(p, l, teams_out, st, per_sm * l.num_sms);
            if (rc != SSFM_ERR_UNSUPPORTED) return rc;
        }
    }
    if constexpr (M1 * M2 >= (1 << 16) && M1 * M2 <= (1 << 18)) {        // 32 .. 64 tiles per waveform
        // (measured on B200 against flag-based / multi-cluster teams: fixed step 2^18 +6 %, 2^17 +9 %; adaptive steps -- two passes per
        //  column phase -- fp64 +4 % / +6 %, fp32 +0 % / +16 %)
        if (l.cluster != 0 && total > 16 && total <= 64 && !getenv("SSFM_NO_MT")) {
            const int rc = wf_launch_mt<R, M1, M2, SMALL>(p, l, teams_out, st, per_sm * l.num_sms);
            if (rc != SSFM_ERR_UNSUPPORTED) return rc;
        }
    }
    if constexpr (M1 * M2 >= (1 << 16)) {                                // (2^16 with two polarisations, 2^17 .. 2^20)
        // measured on B200 (scripts/exp_mc.py): one 2^20-sample waveform, fp32: 4.97 ms against 6.18 ms with flag-based teams
        // (+24 %); fp64: 7.57 against 7.40 ms (its phases are longer, the barrier is a smaller share, and clusters of 8 leave the
        // chip 32 CTAs short of the cooperative grid) -- so "auto" (-1) takes this path for fp32 only, cluster = 1 forces it
        if ((l.cluster > 0 || (l.cluster < 0 && sizeof(R) == 4)) && total > 16 && total <= l.num_sms * (long long)per_sm) {
            const int rc = wf_launch_mc<R, M1, M2, SMALL>(p, l, teams_out, st);
            if (rc != SSFM_ERR_UNSUPPORTED) return rc;
        }
    }
    const long long nslots = 1;
    if (getenv("SSFM_DEBUG"))
        fprintf(stderr, "[ssfm] k_wf<%d,%d,%d,%d>: smem %zu, %d CTAs/SM, team %lld CTAs, %lld teams x %lld slots\n", (int)sizeof(R),
                M1, M2, (int)SMALL, (size_t)GEO::smem, per_sm, total, teams, nslots);

    // sync scratch: [sm_cnt: 4 KB][grid_bar: 128 B][next_wf: 128 B][bar: ts x 128 B][mail: ts x 128 B]
    //               [max words: ts x 2 x total x 16 B],  ts = teams x slots
    const size_t head = 4096 + 256;
    const size_t ts = (size_t)teams * nslots;
    const size_t need = head + ts * 256 + ts * 2 * total * 16;
    if (need > WF_SYNC_BYTES) return SSFM_ERR_UNSUPPORTED;
    WF_TRY(cudaMemsetAsync(l.sync_buf, 0, need, st));
    WfArgs<R> a;
    std::memset(&a, 0, sizeof(a));
    char* sb = (char*)l.sync_buf;
    a.sm_cnt = (unsigned int*)sb;
    a.grid_bar = (unsigned int*)(sb + 4096);
    a.next_wf = (unsigned int*)(sb + 4096 + 128);
    a.bar = (unsigned int*)(sb + head);
    a.mail = (unsigned long long*)(sb + head + ts * 128);
    a.slots = (unsigned long long*)(sb + head + ts * 256);

    a.occ = per_sm;
    a.placement = placement;
    a.budget = l.budget;
    a.n_teams = (int)teams;
    a.fixed = l.fixed; a.single = l.single; a.resume = l.resume;
    a.h_fixed = (R)l.h_fixed;
    a.ready = l.ready; a.done = l.done; a.chunk_rows = l.chunk_rows;
    Params<R> pp = p;
    void* args[2] = {(void*)&pp, (void*)&a};
    if (l.ev0) WF_TRY(cudaEventRecord(l.ev0, st));
    // the grid fills every CTA slot of the chip (co-residency is what the cooperative launch guarantees)
    WF_TRY(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)(per_sm * l.num_sms)), dim3(GEO::NT), args, GEO::smem, st));
    if (l.ev1) WF_TRY(cudaEventRecord(l.ev1, st));
    ++ssfm_launches;
    if (teams_out) *teams_out = (int)(teams * nslots);
    return SSFM_OK;
}

template <typename R, int M1, int M2>
int wf_launch_small(const Params<R>& p, const WfLaunch& l, int* teams_out, cudaStream_t st) {
    if (p.small_phase) return wf_launch<R, M1, M2, true>(p, l, teams_out, st);   // |Kerr phase| <= 0.05 rad: short Taylor sincos
    return wf_launch<R, M1, M2, false>(p, l, teams_out, st);
}

}  // namespace

template <typename R>
int wf_propagate(const Params<R>& p, const WfLaunch& l, int* teams_out, cudaStream_t st) {
#define WF_CASE(A, B) if (p.n1 == A && p.n2 == B) return wf_launch_small<R, A, B>(p, l, teams_out, st);
    WF_CASE(64, 64) WF_CASE(64, 128)      // 2^12, 2^13: teams of one / two CTAs (per polarisation)
    WF_CASE(128, 128) WF_CASE(128, 256) WF_CASE(256, 256) WF_CASE(256, 512) WF_CASE(512, 512) WF_CASE(512, 1024)
    WF_CASE(1024, 1024)
#undef WF_CASE
    return SSFM_ERR_UNSUPPORTED;
}

template int wf_propagate<SSFM_WF_REAL>(const Params<SSFM_WF_REAL>&, const WfLaunch&, int*, cudaStream_t);

}  // namespace ssfm
