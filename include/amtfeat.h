/*
 * libamtfeat -- C-ABI of the B200-native amt_tools.features front end.
 *
 * The reference (cwitkowitz/amt-tools) has NO native / FFI interface: its boundary is the Python class
 * API of amt_tools.features.FeatureModule.  Every entry point below therefore cites the reference
 * *method* it replaces (paths under /root/reference/amt_tools/features/).  A maintainer binds these
 * with ctypes; amt_tools_b200/_lib.py is that binding.
 *
 * Conventions
 *   - plain C types only; no torch / CUDA types in signatures (a stream is passed as void*, i.e. a
 *     cudaStream_t / CUstream handle; NULL = legacy default stream).
 *   - every function returns an int status (0 = AMTFEAT_OK) unless it returns a value that cannot
 *     fail; on failure amtfeat_last_error() (thread-local) holds a message.  Nothing throws across
 *     the ABI, nothing calls exit().
 *   - a plan's tables are immutable after creation: amtfeat_process* may be called concurrently on
 *     distinct (stream, workspace) pairs; the calls serialise only their short host-side enqueue on
 *     a plan lock (the clip-descriptor ring, the fork / join events of a call slot and the profiling
 *     records of amtfeat_profile_* are plan state under that lock).  The caller owns audio, output
 *     and workspace buffers.
 *   - device = -1 builds a host-only plan (all integer / time queries work, no GPU needed);
 *     amtfeat_process* on such a plan fails with AMTFEAT_ERR_NO_DEVICE.  There is no CPU compute path.
 */
#ifndef AMTFEAT_H_
#define AMTFEAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMTFEAT_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define AMTFEAT_API __attribute__((visibility("default")))
#else
#define AMTFEAT_API
#endif

enum {
    AMTFEAT_OK = 0,
    AMTFEAT_ERR_INVALID = 1,   /* bad argument / unsupported configuration (Python: ValueError)   */
    AMTFEAT_ERR_CUDA = 2,      /* a CUDA runtime call failed                                       */
    AMTFEAT_ERR_NO_DEVICE = 3, /* compute requested on a host-only plan                            */
    AMTFEAT_ERR_WORKSPACE = 4  /* workspace too small                                              */
};

/* Which reference FeatureModule a plan mirrors. */
enum {
    AMTFEAT_WAVEFORM = 0, /* waveform.py:14  WaveformWrapper  -> (win, T) frames                   */
    AMTFEAT_STFT = 1,     /* stft.py:11      STFT             -> (1, n_fft/2+1, T)                 */
    AMTFEAT_MEL = 2,      /* mel.py:11       MelSpec          -> (1, n_mels, T)                    */
    AMTFEAT_VQT = 3,      /* vqt.py:17 VQT / cqt.py:7 CQT (gamma = 0) -> (1, n_bins, T)            */
    AMTFEAT_HVQT = 4,     /* hvqt.py:12 HVQT / hcqt.py:7 HCQT -> (H, n_bins, T), shared ladder     */
    AMTFEAT_POWER = 5     /* power.py:12     SignalPower      -> (T,)                              */
};

#define AMTFEAT_MAX_HARMONICS 16

/* Constructor arguments of the reference modules (stft.py:15, mel.py:15, vqt.py:21, hvqt.py:16,
 * power.py:16, waveform.py:18), with the Python-side defaults already resolved. */
typedef struct amtfeat_config {
    int32_t kind;
    int32_t hop_length;
    double sample_rate;
    int32_t decibels;       /* common.py:203-230 post_proc / power.py:52-55                         */
    int32_t center;         /* waveform.py:41                                                        */
    int32_t win_length;     /* waveform / stft / mel / power                                         */
    int32_t n_fft;          /* stft / mel (power of two, 8..2048)                                   */
    int32_t n_mels;         /* mel.py:37                                                             */
    int32_t htk;            /* mel.py:38                                                             */
    int32_t n_bins;         /* vqt.py:49                                                             */
    int32_t bins_per_octave;
    double fmin;            /* vqt.py:41-46 (already defaulted to note C1)                           */
    double gamma;           /* vqt.py:52-58 (already defaulted); 0 for CQT / HCQT                    */
    int32_t n_harmonics;    /* hvqt.py:38-42; 1 for VQT                                              */
    int32_t n_decim_taps;   /* 0 = built-in soxr-HQ-class design; else length of decim_taps (odd)    */
    double harmonics[AMTFEAT_MAX_HARMONICS]; /* sorted ascending (hvqt.py:40)                        */
    const double *decim_taps; /* optional 2:1 decimator taps, DC gain 1, linear phase (may be NULL)  */
} amtfeat_config;

typedef struct amtfeat_plan amtfeat_plan;

AMTFEAT_API int amtfeat_version(void);
AMTFEAT_API const char *amtfeat_last_error(void);

/* FeatureModule.__init__ of the kind named in cfg->kind.  device >= 0: CUDA ordinal; -1: host-only. */
AMTFEAT_API int amtfeat_plan_create(const amtfeat_config *cfg, int device, amtfeat_plan **out);
AMTFEAT_API void amtfeat_plan_destroy(amtfeat_plan *plan);

/* get_expected_frames (common.py:41-66, waveform.py:43-66, vqt.py:102-134, hvqt.py:60-83). */
AMTFEAT_API int64_t amtfeat_expected_frames(const amtfeat_plan *plan, int64_t num_samples);
/* Frames process_audio actually returns for num_samples (what librosa would give; equals
 * amtfeat_expected_frames for every configuration of the reference's examples). */
AMTFEAT_API int64_t amtfeat_output_frames(const amtfeat_plan *plan, int64_t num_samples);
/* get_sample_range (common.py:68-97, waveform.py:68-96, vqt.py:136-165, hvqt.py:85-105): the range is
 * the inclusive interval [*lo, *hi]; num_frames <= 0 gives lo = hi = 0 (the reference's array([0])). */
AMTFEAT_API int amtfeat_sample_range(const amtfeat_plan *plan, int64_t num_frames, int64_t *lo, int64_t *hi);
/* get_num_samples_required (common.py:99-112). */
AMTFEAT_API int64_t amtfeat_num_samples_required(const amtfeat_plan *plan);
/* get_times (common.py:232-258, waveform.py:155-185, vqt.py:197-227, hvqt.py:148-168): writes
 * amtfeat_expected_frames(num_samples) float64 values, bit-exact with librosa.frames_to_time. */
AMTFEAT_API int amtfeat_times(const amtfeat_plan *plan, int64_t num_samples, int at_start, double *out, int64_t capacity);
/* VQT.get_early_ds_count (vqt.py:64-100) for harmonic index h (0 for VQT / CQT). */
AMTFEAT_API int amtfeat_early_ds_count(const amtfeat_plan *plan, int harmonic_index);
/* get_num_channels / get_feature_size (common.py:284-308 and overrides). */
AMTFEAT_API int amtfeat_num_channels(const amtfeat_plan *plan);
AMTFEAT_API int amtfeat_feature_size(const amtfeat_plan *plan);
/* Shape of process_audio's result for one clip: ndim in {1, 2, 3}; empty audio follows the
 * reference quirks (stft.py:59 -> (1, n_fft, 0); mel.py:57; waveform.py:138). */
AMTFEAT_API int amtfeat_out_shape(const amtfeat_plan *plan, int64_t num_samples, int64_t shape[3], int *ndim);
/* JSON description of the plan (ladder levels, n_fft per level, nnz, taps, ...) for tests / docs. */
AMTFEAT_API int amtfeat_plan_describe(const amtfeat_plan *plan, char *buf, size_t capacity);

/* JSON description of how ONE clip of num_samples is laid out (tests / docs): frames stored and computed, the frames each
 * harmonic's dB maximum runs over (hvqt.py:123-128), the ladder level lengths and -- for harmonics librosa early-downsamples
 * by 2^eds >= 4 in one resample call (vqt.py:183) -- the geometry of the exact ladders' tails. */
AMTFEAT_API int amtfeat_clip_describe(const amtfeat_plan *plan, int64_t num_samples, char *buf, size_t capacity);

/* Device workspace needed by amtfeat_process for a batch with these clip lengths. */
AMTFEAT_API size_t amtfeat_workspace_bytes(const amtfeat_plan *plan, int batch, const int64_t *num_samples);

/*
 * process_audio (stft.py:42, mel.py:40, vqt.py:167, hvqt.py:107, power.py:31, waveform.py:121) for a
 * ragged batch of clips, all on the device.
 *   d_audio        device float32; clip b occupies [in_offsets[b], in_offsets[b] + num_samples[b])
 *                  (element offsets, each a multiple of 4)
 *   d_out          device float32; clip b's (C, F, T_b) block, T contiguous, starts at out_offsets[b]
 *   in_offsets, num_samples, out_offsets   HOST arrays of length batch
 *   d_workspace    device scratch of at least amtfeat_workspace_bytes(...)
 * Work is enqueued on `stream`; the call does not synchronise.
 */
AMTFEAT_API int amtfeat_process(const amtfeat_plan *plan, const float *d_audio, const int64_t *in_offsets,
                    const int64_t *num_samples, const int64_t *out_offsets, int batch, float *d_out,
                    void *d_workspace, size_t workspace_bytes, void *stream);

/*
 * One long track computed as chunks (several GPUs, or one chunk after the other).  The reference normalises a track by ITS
 * maximum (`ref=np.max`, features/common.py:199, 224-225; mel.py:94; power.py:55), so chunks computed apart need one
 * exchange step: the maximum over all chunks of C floats.
 *   amtfeat_process_raw       amtfeat_process without the dB epilogue: a dB plan leaves 10 log10(max(amin, power)) (no
 *                             reference, no floor, no rescaling) in d_out; a linear plan leaves its final values
 *   amtfeat_range_reference   d_ref[c] = max(d_ref[c], max over rows and frames [t_begin, t_end) of channel c of the
 *                             (C, F, frames) block at d_block); the caller initialises d_ref (C floats) with -inf.
 *                             log10 is monotonic: this is the log of the maximum, bit for bit what amtfeat_process uses
 *   amtfeat_range_finish      frames [t_begin, t_end) of the raw block -> frames [t_dst, ...) of the (C, F, dst_frames) block
 *                             at d_dst, through max(v - ref[c], -80) / 80 + 1 (SignalPower: no rescaling; linear plans: copy)
 * Deviation: an HVQT / HCQT harmonic whose own VQT is a frame or two longer than the common frame count (hvqt.py:123-128) has
 * its maximum taken over the stored frames only on this path.
 */
AMTFEAT_API int amtfeat_process_raw(const amtfeat_plan *plan, const float *d_audio, const int64_t *in_offsets,
                    const int64_t *num_samples, const int64_t *out_offsets, int batch, float *d_out,
                    void *d_workspace, size_t workspace_bytes, void *stream);
AMTFEAT_API int amtfeat_range_reference(const amtfeat_plan *plan, const float *d_block, int64_t frames, int64_t t_begin,
                    int64_t t_end, float *d_ref, void *stream);
AMTFEAT_API int amtfeat_range_finish(const amtfeat_plan *plan, const float *d_block, int64_t frames, int64_t t_begin,
                    int64_t t_end, const float *d_ref, float *d_dst, int64_t dst_frames, int64_t t_dst, void *stream);

/*
 * Same, with HOST buffers (pinned for full PCIe speed): copies the audio to d_audio, runs
 * amtfeat_process, copies the features back to h_out -- all stream-ordered on `stream`, no sync.
 * d_audio / d_out are caller-provided device staging buffers large enough for the batch.
 */
AMTFEAT_API int amtfeat_process_host(const amtfeat_plan *plan, const float *h_audio, const int64_t *in_offsets,
                         const int64_t *num_samples, const int64_t *out_offsets, int batch, float *h_out,
                         int64_t audio_elems, int64_t out_elems, float *d_audio, float *d_out,
                         void *d_workspace, size_t workspace_bytes, void *stream);

/*
 * Pipelined host executor for bulk precompute (TranscriptionDataset.calculate_feats, datasets/common.py:212-295, applied to
 * a whole corpus shard): three in-order streams (upload, compute, download) and `nslots` sets of device staging buffers,
 * chained with events, so the download of batch i overlaps the kernels of batch i+1 and the upload of batch i+2 and the
 * PCIe link stays busy in both directions.  Host buffers should be pinned.  One pipeline per GPU; plans of any module
 * kind may share it (the plan is named per submission).
 *   submit   enqueues H2D -> amtfeat_process -> D2H for one ragged batch and returns immediately with a ticket
 *   wait     blocks until the features of that ticket (ticket < 0: everything submitted) are in h_out
 * A slot's buffers are reused after `nslots` submissions; the event chain orders the reuse, the caller only has to keep
 * h_audio / h_out alive until wait() returns.
 */
typedef struct amtfeat_pipeline amtfeat_pipeline;
AMTFEAT_API int amtfeat_pipeline_create(int device, int nslots, int64_t max_audio_elems, int64_t max_out_elems,
                                        size_t max_workspace_bytes, amtfeat_pipeline **out);
AMTFEAT_API int amtfeat_pipeline_submit(amtfeat_pipeline *pipe, const amtfeat_plan *plan, const float *h_audio,
                                        const int64_t *in_offsets, const int64_t *num_samples, const int64_t *out_offsets,
                                        int batch, float *h_out, int64_t audio_elems, int64_t out_elems, int64_t *ticket);
AMTFEAT_API int amtfeat_pipeline_wait(amtfeat_pipeline *pipe, int64_t ticket);
AMTFEAT_API void amtfeat_pipeline_destroy(amtfeat_pipeline *pipe);

/* Number of kernel launches amtfeat_process enqueues for this batch (bench.py's gpu_launches). */
AMTFEAT_API int amtfeat_launch_count(const amtfeat_plan *plan, int batch, const int64_t *num_samples);

/*
 * Consumer-side helper: tools.framify_activations (amt_tools/tools/utils.py:2922-2984) on the device, so
 * that TabCNN.pre_proc (models/tabcnn.py:123-127) needs no host round trip.  d_in is (rows, num_frames) row-major, d_out is
 * (rows, hops, win_length) with hops = amtfeat_framify_hops(...); zero padding as librosa.util.pad_center.
 */
AMTFEAT_API int64_t amtfeat_framify_hops(int64_t num_frames, int win_length, int hop_length, int pad);
AMTFEAT_API int amtfeat_framify(const float *d_in, int64_t rows, int64_t num_frames, int win_length, int hop_length, int pad,
                                float *d_out, void *stream);

/*
 * Audio ingest on the device: what tools.load_normalize_audio (amt_tools/tools/io.py:50-87) does after
 * decoding the file -- librosa.to_mono, librosa.resample(res_type='kaiser_best' | 'kaiser_fast') (resampy's windowed-sinc
 * interpolation) and tools.rms_norm (amt_tools/tools/utils.py:2789-2814).  A resampler owns the interpolation table of one
 * (sr_orig, sr_new, filter) triple; device < 0 builds a host-only object (lengths / table queries, no compute).
 * Clips are packed like amtfeat_process inputs: element offsets + lengths.  The workspace holds per-clip descriptors.
 */
enum { AMTFEAT_RES_KAISER_BEST = 0, AMTFEAT_RES_KAISER_FAST = 1 };
typedef struct amtfeat_resampler amtfeat_resampler;
AMTFEAT_API int amtfeat_resampler_create(double sr_orig, double sr_new, int filter, int device, amtfeat_resampler **out);
AMTFEAT_API void amtfeat_resampler_destroy(amtfeat_resampler *r);
/* ceil(num_samples * sr_new / sr_orig): librosa.resample(fix=True) pads resampy's int(n * ratio) samples with zeros to that length */
AMTFEAT_API int64_t amtfeat_resampler_out_len(const amtfeat_resampler *r, int64_t num_samples);
/* Copies up to `capacity` entries of the half window (scaled by the ratio when downsampling); returns its length.
 * num_table = table samples per zero crossing, index_step = stride per input sample (either may be NULL). */
AMTFEAT_API int64_t amtfeat_resampler_table(const amtfeat_resampler *r, double *win, int64_t capacity, int *num_table, int *index_step);
AMTFEAT_API size_t amtfeat_ingest_workspace_bytes(int batch);
AMTFEAT_API int amtfeat_resample(const amtfeat_resampler *r, const float *d_in, const int64_t *in_offsets, const int64_t *num_samples,
                                 int batch, float *d_out, const int64_t *out_offsets, void *d_ws, size_t ws_bytes, void *stream);
/* 16-bit PCM -> float32 on the device, d_out[i] = d_pcm[i] * scale (librosa.load / soundfile scale: 1 / 32768; tools/io.py:78): a
 * device-resident consumer uploads half the bytes.  d_pcm 8-byte aligned, d_out 16-byte aligned. */
AMTFEAT_API int amtfeat_pcm16_to_float(const int16_t *d_pcm, int64_t num_samples, float scale, float *d_out, void *stream);
/* d_in is (channels, num_samples) row-major; d_out receives the channel mean */
AMTFEAT_API int amtfeat_to_mono(const float *d_in, int64_t num_samples, int channels, float *d_out, void *stream);
/* in place: clip / sqrt(mean(clip^2)); an all-zero (or empty) clip is left unchanged */
AMTFEAT_API int amtfeat_rms_norm(float *d_audio, const int64_t *offsets, const int64_t *num_samples, int batch, void *d_ws,
                                 size_t ws_bytes, void *stream);

/*
 * Measurement hook (no reference counterpart): when enabled, amtfeat_process records a CUDA event pair
 * around every kernel it launches on the launching stream; amtfeat_profile_read waits for them and
 * writes {"<kernel>": {"ms": total, "launches": n}, ...} as JSON, then clears the records.
 * Single-threaded use only.
 */
AMTFEAT_API int amtfeat_profile_enable(amtfeat_plan *plan, int enable);
AMTFEAT_API int amtfeat_profile_read(amtfeat_plan *plan, char *buf, size_t capacity);

#ifdef __cplusplus
}
#endif
#endif /* AMTFEAT_H_ */
