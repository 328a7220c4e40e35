// Host-side interface of the persistent whole-propagation kernel (ssfm_wf.cuh / ssfm_wf_impl.inl).
#pragma once
#include <cuda_runtime.h>
#include "ssfm_kernels.cuh"

namespace ssfm {

struct WfLaunch {
    void* sync_buf;          // plan-owned scratch for barriers, mailboxes and max words (WF_SYNC_BYTES)
    int num_sms;
    int fixed, single, resume;
    double h_fixed;
    long long budget;        // steps per waveform in this call (> 0)
    int teams_cap;           // 0 = as many teams as fit on the chip
    int placement;           // 1 = SM-aware team placement, 0 = by blockIdx, -1 = auto (measured: on for fp32 with three
                             // CTAs per SM, +16 %; off for fp64 with two, where it costs 3 %)
    int cluster;             // teams of <= 16 CTAs as thread-block clusters: 1 = always, 0 = never, -1 = when >= 80 % of the
                             // CTA slots of the cooperative variant can be filled with clusters
    cudaStream_t side;       // side stream + event for the launch that fills the CTA slots the clusters leave (may be null)
    cudaEvent_t ev_side;
    cudaEvent_t ev0, ev1;    // recorded around the launch on the stream (may be null)
    void* tstash;            // multi-tile cluster teams: Kerr phase of the waveforms in flight (may be null: variant not used)
    size_t tstash_bytes;
    const unsigned int* ready;   // streamed batches (ssfm_propagate_streamed): arrival counter, completion counters per chunk
    unsigned int* done;
    int chunk_rows;
};
constexpr size_t WF_SYNC_BYTES = 1u << 20;

// Runs the whole propagation of p.batch waveforms as one cooperative launch of k_wf.
// Returns SSFM_ERR_UNSUPPORTED (no error text) when the geometry has no k_wf instantiation or one
// waveform's team does not fit on the chip; the caller then uses the multi-launch schedule.
// *in_flight = waveforms in flight (= teams).
template <typename R> int wf_propagate(const Params<R>& p, const WfLaunch& l, int* in_flight, cudaStream_t st);

}  // namespace ssfm
