/* ssfm_b200.h -- C ABI of the B200-native split-step Fourier engine.
 *
 * Drop-in boundary for ONE hot path of armando-palacio/opticomlib (v2.0.3):
 *   opticomlib.devices.FIBER   (opticomlib/devices.py:1038-1206)
 *   opticomlib.devices.DBP     (opticomlib/devices.py:1209-1283)
 *   opticomlib.devices.LPF     (opticomlib/devices.py:1286-1375)
 *   opticomlib.devices.BPF     (opticomlib/devices.py:788-826)
 *
 * The reference has no FFI of its own (it is pure Python; its only accelerator hook is an
 * `import cupy` inside FIBER, devices.py:1114-1134), so each entry point below names the block of
 * reference statements it replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on
 * the reference side.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types.
 *   - all "dev" pointers are CUDA device pointers on the plan's device, owned by the caller and
 *     borrowed for the duration of the call; "host" pointers are ordinary host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every function returns SSFM_OK (0) or a negative error code and never throws; a message
 *     for the last error of the calling thread is available from ssfm_last_error().
 *   - complex samples are interleaved (re, im); SSFM_C64 = 2 x float32, SSFM_C128 = 2 x float64.
 */
#ifndef SSFM_B200_H
#define SSFM_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SSFM_API __attribute__((visibility("default")))
#else
#define SSFM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define SSFM_OK              0
#define SSFM_ERR_INVALID    -1   /* bad argument (maps to ValueError / TypeError in the wrapper) */
#define SSFM_ERR_CUDA       -2   /* CUDA runtime error (RuntimeError) */
#define SSFM_ERR_UNSUPPORTED -3  /* size not supported by this build (ValueError) */
#define SSFM_ERR_NOMEM      -4

#define SSFM_C64   0   /* float32 arithmetic: the reference as shipped (devices.py:1137-1147) */
#define SSFM_C128  1   /* float64 arithmetic: the dtype-lifted algorithm (devices.py:2440-2486) */

#define SSFM_ABI_VERSION 1

typedef struct ssfm_plan_s* ssfm_plan_t;

/* Fibre and step-control arguments of FIBER (devices.py:1038-1048), in the reference's units. */
typedef struct ssfm_fiber_params {
    double dt_s;           /* sampling period gv.dt [s]  (typing.py:1641 via input.w()) */
    double length_km;      /* length   [km] */
    double alpha_db_km;    /* alpha    [dB/km]; divided by the literal 4.343 inside (devices.py:1137) */
    double beta2_ps2_km;   /* beta_2   [ps^2/km] */
    double beta3_ps3_km;   /* beta_3   [ps^3/km] */
    double gamma_w_km;     /* gamma    [1/(W km)] */
    double phi_max_rad;    /* phi_max  [rad] (code default 0.01, devices.py:1044) */
    double h_km;           /* fixed step [km]; NaN selects the adaptive rule of devices.py:1156,1194 */
} ssfm_fiber_params;

SSFM_API int         ssfm_abi_version(void);
SSFM_API const char* ssfm_last_error(void);

/* Plan for `n_waveforms` independent waveforms of `n_pol` (1|2) polarisation rows of `n_samples`
 * complex samples each.  Powers of two 2^8 <= n <= 2^22 run on the two-pass transform kernels; any
 * other length 2 <= n <= 2^21 (the reference accepts any N, numpy.fft) runs the same statements with
 * chirp-z (Bluestein) transforms built on those kernels -- correct but about an order of magnitude
 * slower per sample; longer power-of-two waveforms: ssfm_long_plan_create.  Holds the twiddle tables,
 * the Kerr-phase stash and the per-waveform controller state on `device`. */
SSFM_API int ssfm_plan_create(ssfm_plan_t* plan, int64_t n_samples, int32_t n_pol, int64_t n_waveforms,
                     int32_t dtype, int32_t device);
SSFM_API int ssfm_plan_destroy(ssfm_plan_t plan);

/* Tunables: "persistent" (1 = run the whole propagation as one persistent kernel whose teams of CTAs keep
 * the waveforms in flight resident in L2 -- the default whenever a waveform's team fits on the chip;
 * 0 = multi-launch schedule), "teams" (cap on the teams of CTAs = waveforms in flight of the persistent kernel, 0 = auto),
 * "chunk_waveforms" (multi-launch schedule: waveforms propagated together, 0 = all), "burst_steps",
 * "fused" (multi-launch schedule: 0 three kernels per step, 1..3 two kernels per step). */
SSFM_API int ssfm_plan_set_option(ssfm_plan_t plan, const char* name, int64_t value);
/* Read a tunable back, or one of the read-only properties "hlog_cap" (entries per waveform of the step-size log),
 * "n1", "n2" (the two-pass split of the transform). */
SSFM_API int ssfm_plan_get_option(ssfm_plan_t plan, const char* name, int64_t* value);

/* FIBER hot loop, devices.py:1155-1196, in place on field_dev[n_waveforms][n_pol][n_samples].
 * DBP (devices.py:1280-1283) is the same call with alpha, beta_2, beta_3, gamma negated by the caller.
 *   max_steps : stop every waveform after this many further steps (0 = run to z >= length)
 *   resume    : 0 = start at z = 0 (first step size from devices.py:1155-1159);
 *               1 = continue from the controller state left by the previous call on this plan
 * Returns when the field is final on `stream` (the call synchronises on its own work). */
SSFM_API int ssfm_propagate(ssfm_plan_t plan, void* field_dev, const ssfm_fiber_params* prm,
                   int64_t max_steps, int32_t resume, void* stream);

/* Controller state after ssfm_propagate, copied to host arrays of n_waveforms entries
 * (any pointer may be NULL): steps taken, z reached [km], size of the next step [km],
 * 1 if z >= length. */
SSFM_API int ssfm_get_state(ssfm_plan_t plan, int32_t* steps_host, double* z_host, double* h_next_host,
                   int32_t* done_host);
/* Host pipelines: with plan option "async" = 1, ssfm_propagate returns as soon as the persistent kernel is enqueued on
 * `stream` (when the multi-launch schedule has to be used the call still blocks).  ssfm_copy_state_async enqueues a copy of
 * the raw per-waveform controller records to (pinned) host memory on the same stream: n_waveforms records of 40 bytes
 * { double z; double h_next; uint64 scratch; int32 steps; int32 done; uint32 scratch; int32 pad }. */
SSFM_API int ssfm_copy_state_async(ssfm_plan_t plan, void* records_host, void* stream);
/* Streamed batches: ONE persistent launch carries a batch whose chunks are still being copied from the host, so the copies of
 * the whole batch overlap its propagation and the kernel keeps every team busy across chunk boundaries (replaces the per-chunk
 * launches of the host pipeline behind FIBER / fiber_batch for host buffers: reference call devices.py:1038 on NumPy arrays).
 *   ready_dev[0]   number of waveforms (rows of `field`, in order) whose samples are on the device; the caller advances it in
 *                  stream order behind every host-to-device copy (ssfm_stream_write_u32); the kernel draws waveform w only
 *                  once ready_dev[0] > w;
 *   done_dev[c]    incremented by the kernel, reaches rows_in_chunk(c) x n_pol x n_samples / 4096 when every sample of chunk c
 *                  (chunk_rows consecutive waveforms; the last chunk may be shorter) holds its final value; the caller makes the
 *                  stream of the device-to-host copy of chunk c wait for that (ssfm_stream_wait_geq_u32).
 * Both arrays are device memory, zeroed by the caller before the call.  The call never blocks; SSFM_ERR_UNSUPPORTED (nothing
 * enqueued) when the plan's geometry has no persistent kernel. */
SSFM_API int ssfm_propagate_streamed(ssfm_plan_t plan, void* field_dev, const ssfm_fiber_params* prm, const uint32_t* ready_dev,
                                     uint32_t* done_dev, int64_t chunk_rows, void* stream);
/* Stream memory operations (CUDA driver cuStreamWriteValue32 / cuStreamWaitValue32 >=) for the two counters above. */
SSFM_API int ssfm_stream_write_u32(void* stream, void* dev_ptr, uint32_t value);
SSFM_API int ssfm_stream_wait_geq_u32(void* stream, const void* dev_ptr, uint32_t value);
/* Progress of one waveform WHILE an asynchronous ssfm_propagate (option "async") is running: the controller record of
 * `row` is copied on a stream of its own (it does not wait for the propagation).  The reference updates a tqdm bar after
 * every step (devices.py:1164-1170, 1188-1191); this is the device-side counter such a bar polls. */
SSFM_API int ssfm_peek_state(ssfm_plan_t plan, int64_t row, int32_t* steps_host, double* z_host, int32_t* done_host);
/* Schedule used by the last ssfm_propagate on this plan: *kind = 1 multi-launch, 2 persistent kernel
 * (then *teams = waveforms in flight and *kernel_ms = device time of that one launch, CUDA events on the
 * launching stream).  Any pointer may be NULL.  Measurement hook for bench.py's roofline object. */
SSFM_API int ssfm_get_last_timing(ssfm_plan_t plan, int32_t* kind, int32_t* teams, float* kernel_ms);
/* Step sizes taken, hlog_host[n_waveforms][cap] (rows are filled up to min(steps, cap)). */
SSFM_API int ssfm_get_step_log(ssfm_plan_t plan, double* hlog_host, int64_t cap);

/* Measurement hooks (bench.py): number of kernels this library has launched in this process, and
 * the average device time [ms] of the three kernels of one split step (column forward, row,
 * column inverse) measured with CUDA events on `stream` over `reps` steps on field_dev (the field
 * is advanced by `reps` fixed steps of prm->h_km or 1e-3 km; use a scratch copy). */
SSFM_API int64_t ssfm_launch_count(void);
SSFM_API int ssfm_time_step_kernels(ssfm_plan_t plan, void* field_dev, const ssfm_fiber_params* prm,
                           int32_t reps, float* ms_out3, void* stream);

/* Same as ssfm_propagate but with HOST buffers: copies field_in_host to the device (casting is the
 * caller's job: the buffer must already have the plan's dtype), propagates, copies the result to
 * field_out_host.  This is the block devices.py:1147-1204 (asarray -> loop -> .get()). */
SSFM_API int ssfm_fiber_host(ssfm_plan_t plan, const void* field_in_host, void* field_out_host,
                    const ssfm_fiber_params* prm, void* stream);

/* FFT-domain transfer function, in place on field_dev[n_waveforms][n_pol][n_samples]:
 *   row <- ifft(fft(row) * H),  H = h_dev[n_samples] (plan dtype, numpy bin order k = 0..N-1).
 * This is the operation of DM (devices.py:1025-1029) and of the apply step of FBG (devices.py:2314-2316);
 * the zero-phase filters use it internally with H = |H_sos|^2. */
SSFM_API int ssfm_apply_transfer(ssfm_plan_t plan, void* field_dev, const void* h_dev, void* stream);

/* ---- long waveforms: N = N0 x N_l beyond one two-pass transform (N > 2^22) and/or one waveform spread over
 * `n_ranks` GPUs (BASELINE config #5).  The same statements of devices.py:1155-1196; only the transform is split:
 * rank g holds columns [g N_l/G, (g+1) N_l/G) of the N0 x N_l sample matrix (local [N0][N_l/G], sample n = na N_l + nb)
 * in the time domain and rows [g N0/G, (g+1) N0/G) (local [N0/G][N_l]) in the frequency domain.  The caller moves
 * the data between the two layouts (an all-to-all between the ranks; nothing for one rank) -- the library itself has
 * no NCCL dependency -- and drives one split step as
 *     ssfm_long_outer(stage)  ->  exchange  ->  ssfm_long_inner  ->  exchange back  ->  ssfm_long_outer(stage) ...
 * stage 0 "open": first Kerr half step + forward outer transform; stage 2 "close": inverse outer transform + second
 * Kerr half step + local max|A|^2 into the controller (then ssfm_long_pmax get / all-reduce MAX / set when n_ranks > 1,
 * and ssfm_long_ctrl(0) runs the step-size controller); stage 1 "mid" (fixed step only): close + open in one pass,
 * the device controller advances by itself and the last "mid" leaves the field in the time domain.
 * ssfm_get_state / ssfm_get_step_log report z, h, steps, done as for ordinary plans. */
SSFM_API int ssfm_long_plan_create(ssfm_plan_t* plan, int64_t n_samples_global, int32_t n_outer, int32_t n_ranks,
                          int32_t rank, int32_t dtype, int32_t device);
/* Reset the controller for a new propagation; adaptive mode: local max|A|^2 of field_local -> ssfm_long_pmax.
 * Follow with ssfm_long_ctrl(plan, 1, ...) (first step size, devices.py:1155-1161). */
SSFM_API int ssfm_long_begin(ssfm_plan_t plan, void* field_local_dev, const ssfm_fiber_params* prm, void* stream);
SSFM_API int ssfm_long_pmax(ssfm_plan_t plan, double* value_host, int32_t set, void* stream);
SSFM_API int ssfm_long_ctrl(ssfm_plan_t plan, int32_t init, void* stream);
SSFM_API int ssfm_long_outer(ssfm_plan_t plan, void* field_local_dev, int32_t stage, void* stream);
SSFM_API int ssfm_long_inner(ssfm_plan_t plan, void* rows_local_dev, void* stream);
/* Exchange fused into the kernels (several GPUs of one node, one process each).  Every rank exports the CUDA IPC handle
 * (64 bytes) of its exchange buffer, the caller gathers the handles of all ranks (rank order) and every rank imports them.
 * From then on the stage kernels store each result element straight into the buffer of the rank that owns it in the next
 * stage's layout -- peer memory over NVLink, no collective and no re-layout copy -- and the caller replaces each
 * exchange by ssfm_long_xbar (a flag barrier between the GPUs, on the stream).  The time-domain field lives in the
 * library's buffer: ssfm_long_p2p_copy(to_internal = 1) before ssfm_long_begin, (0) after the last stage; the
 * field / rows pointers of ssfm_long_begin / _outer / _inner are then ignored. */
SSFM_API int ssfm_long_p2p_export(ssfm_plan_t plan, void* handle64_host);
SSFM_API int ssfm_long_p2p_import(ssfm_plan_t plan, const void* handles_host);
SSFM_API int ssfm_long_p2p_copy(ssfm_plan_t plan, void* field_local_dev, int32_t to_internal, void* stream);
SSFM_API int ssfm_long_xbar(ssfm_plan_t plan, void* stream);

/* Zero-phase cascaded-biquad filtering (scipy.signal.sosfiltfilt as called at devices.py:820-823 and
 * 1365-1368): x_dev[n_rows][n_samples] complex128 -> y_dev (may alias x_dev).
 * sos_host[n_sections][6] = b0 b1 b2 a0 a1 a2 (a0 == 1) from scipy.signal.bessel(..., output='sos').
 * Real signals are passed as complex with zero imaginary part, or two real rows packed as re/im. */
SSFM_API int ssfm_filtfilt_sos(void* x_dev, void* y_dev, int64_t n_rows, int64_t n_samples,
                      const double* sos_host, int32_t n_sections, int32_t device, void* stream);

/* Photodetector front end + zero-phase low-pass + sampler in one call (reference PD, devices.py:1514-1552, whose last
 * statement is LPF(output, BW), devices.py:1363-1368; SAMPLER, devices.py:1871-1891, output[instant :: sps]).
 *   signal:  R_load * r * sum_pol |E|^2                                      -> out_signal_dev[n_rows][m]   (float64)
 *   noise:   R_load * (r * sum_pol (2 Re(E n*) + |n|^2) + extra + i_dark)    -> out_noise_dev[n_rows][m]    (float64, optional)
 * both filtered separately by the zero-phase cascade `sos_host` (as the reference filters signal and noise separately)
 * and sampled at sample_offset + j * sample_stride, j < m = ceil((n_samples - sample_offset) / sample_stride)
 * (offset 0, stride 1: every sample).  field_dev / noise_dev: complex128 [n_rows][n_pol][n_samples] (noise_dev may be
 * null); extra_noise_dev: float64 [n_rows][n_samples] additive noise CURRENT (thermal + shot; may be null -- see
 * ssfm_gaussian_noise).  out_noise_dev may be null when there is no noise input; i_dark is then ignored, as in the
 * reference's include_noise='none'. */
SSFM_API int ssfm_pd_lpf(const void* field_dev, const void* noise_dev, const double* extra_noise_dev, double* out_signal_dev,
                double* out_noise_dev, int64_t n_rows, int32_t n_pol, int64_t n_samples, double responsivity, double r_load,
                double i_dark, const double* sos_host, int32_t n_sections, int64_t sample_offset, int64_t sample_stride,
                int32_t device, void* stream);

/* out_dev[i] = mean + sigma * N(0,1), i < count: Philox4x32-10 + Box-Muller, a pure function of (seed, substream, i).
 * Stands in for the np.random.normal / randn draws of EDFA (devices.py:933) and PD (devices.py:1523, 1527), which come from
 * NumPy's global stream and cannot be reproduced bit for bit on a device. */
SSFM_API int ssfm_gaussian_noise(double* out_dev, int64_t count, double mean, double sigma, uint64_t seed, uint32_t substream,
                        int32_t device, void* stream);

/* EDFA gain and ASE on the device (reference EDFA, devices.py:921-936, without its optional BPF -- call ssfm_filtfilt_sos):
 *   out[row][pol][i] = sqrt(10^(gain_db/10)) * in[row][pol][i] + sqrt(p_ase_w/4) * (n1 + j n2),  n1, n2 ~ N(0,1) (Philox)
 * in_dev: complex128 [in_rows][in_pol][n_samples] with in_rows == n_rows, or in_rows == 1 to amplify ONE waveform into n_rows
 * independent noise realisations (the Monte-Carlo batch of BASELINE config #3 without any host->device copy);
 * out_dev: complex128 [n_rows][out_pol][n_samples], out_pol >= in_pol (the reference always returns two polarisations: a
 * polarisation the input does not have carries ASE only).  p_ase_w = idb(NF) h f0 (idb(G) - 1) fs (devices.py:930).
 * first_row: index of out_dev's first row in the whole batch (the noise of row b is a function of seed and first_row + b
 * only, so a batch produced in chunks equals the batch produced at once). */
SSFM_API int ssfm_edfa(const void* in_dev, void* out_dev, int64_t n_rows, int64_t in_rows, int32_t in_pol, int32_t out_pol,
              int64_t n_samples, double gain_db, double p_ase_w, uint64_t seed, int64_t first_row, int32_t device, void* stream);

/* Welch power spectral density of every row, as the reference computes it for `signal.psd()` (typing.py:1899-1902) and
 * `utils.get_psd` (utils.py:2074-2079): scipy.signal.welch(x, nperseg, scaling='spectrum', return_onesided=False,
 * detrend=False) -- periodic Hann window, 50 % overlap, mean over the segments -- followed by fftshift of the bins.
 * x_dev: complex128 [n_rows][n_samples]; psd_dev: float64 [n_rows][nperseg] (bin order of fftshift(fftfreq(nperseg)));
 * nperseg in {256, 512, 1024, 2048} (the reference uses min(2048, n_samples)). */
SSFM_API int ssfm_welch_psd(const void* x_dev, double* psd_dev, int64_t n_rows, int64_t n_samples, int32_t nperseg,
                   int32_t device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SSFM_B200_H */
