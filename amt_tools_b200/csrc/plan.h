// Internal plan representation shared by the host-side table builder (host_plan.cpp), the kernel
// launchers (kernels_*.cu) and the C-ABI (api.cpp).  Not part of the public interface.
#pragma once

#include <cstdint>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/amtfeat.h"

namespace amtfeat {

constexpr int kMaxLevels = 16;      // ladder depth (octaves + early downsampling)
constexpr int kMaxAlt = 3;          // exact ("one-shot early downsampling") ladders, one per distinct eds >= 2 among the harmonics
constexpr int kWarpsPerCta = 8;     // every FFT kernel runs 8 warps, one "unit" (1024 complex points) each
constexpr int kThreads = kWarpsPerCta * 32;

struct cfloat {
    float x, y;
};

// Per-clip launch metadata (device array, one per clip of the batch).
struct ClipMeta {
    int64_t in_off;                 // element offset of the clip inside d_audio
    int64_t n;                      // samples
    int64_t out_off;                // element offset of the clip's (C, F, T) block inside d_out
    int64_t lvl_off[kMaxLevels];    // level 0: == in_off (d_audio); level >= 1: offset inside the ladder buffer
    // Exact ladders (harmonics librosa early-downsamples by 2^eds >= 4 in ONE resample call, vqt.py:183): only the end of
    // each level is stored, from sample alt_first on; alt_off is the VIRTUAL offset (storage offset - alt_first) inside the
    // ladder buffer, so that sample m of the level is ladder[alt_off + m] for every m >= alt_first.  The first frames of a
    // level need the exact signal too (the cascade also drops the ringing BEFORE the first sample): head piece.
    int64_t alt_off[kMaxAlt][kMaxLevels];
    int64_t alt_hoff[kMaxAlt][kMaxLevels];   // head piece [0, alt_hlen) of the level (real offset)
    int32_t alt_first[kMaxAlt][kMaxLevels];
    int32_t alt_hlen[kMaxAlt][kMaxLevels];
    int32_t lvl_len[kMaxLevels];    // samples at each ladder level
    int32_t alt_th[kMaxLevels];     // frames [0, alt_th) of a level are served from the exact ladders (0: none) ...
    int32_t alt_t0[kMaxLevels];     // ... and the frames from alt_t0 on (INT32_MAX: none)
    int32_t t_max[AMTFEAT_MAX_HARMONICS];   // frames the dB maximum of a channel runs over: the harmonic's own, untrimmed VQT (hvqt.py:123-128)
    int32_t T;                      // output frames
    int32_t T_all;                  // frames to compute: max over channels of t_max (== T unless decibels on an HVQT)
};

// One row of a sparsified frequency-domain wavelet basis (librosa __vqt_filter_fft + sparsify_rows).
struct CqtRow {
    int32_t chan;      // output channel (harmonic index)
    int32_t bin;       // output frequency bin
    float inv_len;     // 1 / length (the `V /= sqrt(lengths)` of librosa.vqt, squared, applied to power)
    int32_t col0;      // first kept FFT bin
    int32_t cnt;       // kept band width (dropped entries inside the band are stored as zeros)
    int32_t woff;      // offset into the weight array
};

// Up to four adjacent rows of one (harmonic, octave) run, projected together by the lanes that share a
// frame: every FFT bin fetched from shared memory feeds four complex MACs, and the block's weights are a
// warp-uniform stream of [step][4 rows] complex values (two 16-byte loads per step).
//
// Harmonics whose frequencies are an octave apart (h = 0.5, 1, 2 of an HCQT) meet on the same ladder level with
// the same normalised frequencies, i.e. the same wavelet rows: such rows are projected ONCE and stored to every
// destination channel (`ndst` > 1).
constexpr int kMaxDst = 4;
struct CqtBlock4 {
    int32_t col0;       // first FFT bin of the union band of the rows
    int32_t steps;      // width of the union band (weights outside a row's own band are stored as zeros)
    int32_t woff;       // offset into weights4, in float4 (2 float4 per step)
    int32_t ndst;       // number of destination channels sharing these rows
    int32_t chan[kMaxDst];      // output channel (harmonic index) of each destination
    int32_t off[kMaxDst][4];    // chan * F + bin of each row per destination, -1 if absent
    float inv[4];       // 1 / length of each row (applied to the power)
};

static_assert(sizeof(CqtBlock4) % 16 == 0, "block descriptors are staged into shared memory in 16-byte granules");

// All rows that consume the FFT frames of one (ladder level, n_fft) pair.
struct CqtItem {
    int32_t level, nfft, hop, blk0, nblk, kmin, kmax, nrows;
    int32_t row0, kmax_true;    // per-row tables (small-n_fft fallback kernel); last bin with a non-zero weight
    int32_t woff0, wcount;      // the item's slice of weights4 (float4 units), staged into shared memory per CTA
    int32_t nuniq, alt;         // distinct rows after merging rows shared by several harmonics; 0: shared ladder, a >= 1: exact ladder a - 1 (tail frames only)
};

// Items whose hop is a small fraction of n_fft (the deep ladder levels: consecutive frames share all but `hop` samples)
// are computed by cqt_slide_kernel -- a sliding DFT on the item's band -- instead of one FFT per frame.
constexpr int kSlideMaxHop = 16;
inline bool is_slide_item(const CqtItem &it) {
    const int kb = it.kmax - it.kmin + 1;
    return it.hop >= 1 && it.hop <= kSlideMaxHop && (it.hop & (it.hop - 1)) == 0 && it.nfft >= 128 && it.nfft / it.hop >= 16 && kb <= kThreads;
}

struct cfloat4 {
    float ar, ai, br, bi;
};

// FFT constant tables for one complex length NC = n_fft / 2.
struct FftTables {
    std::vector<cfloat> tw1;   // [k1][n2] = exp(-2 pi i k1 n2 / NC), inter-pass twiddles
    std::vector<cfloat> tw2;   // [k] = exp(-i pi k / NC), k = 0..NC, real-FFT split twiddles
    cfloat *d_tw1 = nullptr, *d_tw2 = nullptr;
};

struct HarmonicInfo {
    double fmin;
    int eds_ref;   // reference formula, vqt.py:64-100 (frame counts / sample ranges)
    int eds_lib;   // librosa >= 0.10 formula (signal path)
    int alt;       // -1: served by the shared ladder alone; a >= 0: tail frames come from exact ladder a
};

// One exact ladder: levels eds .. eds + n_oct - 1; level eds is ONE 2^eds : 1 decimation of the audio (librosa
// __early_downsample -> resample), deeper levels are 2:1 steps from it.  Only the tail of each level is ever computed.
struct AltLadder {
    int eds = 0;
    std::vector<double> taps;      // one-shot 2^eds : 1 decimator, x sqrt(2^eds) folded in, odd length
    double *d_taps = nullptr;
};

// Per-clip geometry of the exact ladders (host arithmetic; see clip_tail_layout in host_plan.cpp).
struct TailLayout {
    int32_t th[kMaxLevels];                  // frames [0, th) are served from the exact ladders (0: none)
    int32_t t0[kMaxLevels];                  // ... and the frames from t0 on (INT32_MAX: none at this level)
    int32_t first[kMaxAlt][kMaxLevels];      // tail piece: first stored sample of the level (multiple of 4), -1: level unused
    int32_t count[kMaxAlt][kMaxLevels];      // tail piece: stored samples, len - first
    int32_t hlen[kMaxAlt][kMaxLevels];       // head piece: samples [0, hlen); 0: none; -1: the tail piece holds the whole level
    int32_t hsafe[kMaxLevels];               // shared level == exact level on [hsafe, dev)
    int32_t dev[kMaxLevels];
};

struct Plan {
    amtfeat_config cfg{};
    int device = -1;
    int C = 1, F = 1;

    // STFT family
    std::vector<float> window;                 // n_fft, periodic Hann of win_length, centre padded
    std::vector<int32_t> mel_start, mel_cnt, mel_off;
    std::vector<float> mel_w;

    // VQT family
    int n_oct = 0, n_filters = 0, n_levels = 0;
    std::vector<HarmonicInfo> harm;
    std::vector<float> taps;                   // 2:1 decimator, includes the sqrt(2) of `scale=True`
    // Fast-convolution form of the same decimator (decimate_fft_kernel): per bin k = 0..512 of the folded
    // (decimated) 1024-point spectrum, (Ha, Hb) = (H[k], conj(H[1024 - k])) / 2048 with H the 2048-point DFT of the
    // float32 taps.  Empty when the taps are too long for a 2048-point block (the direct kernel is used then).
    std::vector<cfloat4> decim_hh;
    std::vector<double> taps64;                // the same 2:1 taps, unrounded (exact-ladder tail kernel)
    std::vector<double> decim_h64;             // float64 fast-convolution form: H[k], k < 2048 (re, im), times 1 / 2048
    std::vector<double> decim_tw64;            // exp(-2 pi i m / 2048), m < 1024 (re, im)
    int decim_mode = 0;                        // 0: float64 fast convolution (default); AMTFEAT_DECIM=fft32 -> 1, =direct -> 2 (float32 forms, A/B)
    bool serial_launch = false;                // AMTFEAT_SERIAL=1: no side stream, every launch in order on the caller's stream (isolated per-kernel timing)
    bool meta_memcpy = false;                  // AMTFEAT_META_MEMCPY=1: clip descriptors by cudaMemcpyAsync instead of copy_meta_kernel (A/B)
    bool slide_off = false;                    // AMTFEAT_SLIDE=0 keeps every item on the FFT-per-frame kernel (tests / A-B)
    std::vector<CqtRow> rows;                  // per-row description (host only; tests / describe)
    std::vector<cfloat> weights;               // per-row weights (host only)
    std::vector<CqtBlock4> blocks;
    std::vector<cfloat4> weights4;
    std::vector<CqtItem> items;                // shared-ladder items first (sorted by n_fft, level), then the exact-ladder items
    std::vector<AltLadder> alts;               // exact ladders (empty unless some harmonic has eds_lib >= 2)
    uint32_t alt_mask = 0;                     // channels whose tail frames come from an exact ladder
    int32_t alt_nfft_max[kMaxLevels] = {};     // widest / narrowest transform among the exact-ladder items of a level (0: none)
    int32_t alt_nfft_min[kMaxLevels] = {};
    bool exact_eds = true;                     // AMTFEAT_EXACT_EDS=0: every harmonic from the shared ladder alone (A/B, DESIGN "Known deviations")
    std::vector<int32_t> item_kmax_true;       // last FFT bin with a non-zero weight, per item (describe / tests)
    // mel projection in segment form: FFT bin k feeds the rising slope of filter seg(k) and the falling slope of
    // filter seg(k) - 1.  mel_ww holds (up, down) per bin, padded lane-major [group of 32 segments][step][lane].
    std::vector<int32_t> mel_seg_start;        // first FFT bin of each of the n_mels + 1 segments
    std::vector<cfloat> mel_ww;
    std::vector<int32_t> mel_gsteps, mel_goff; // per group: steps, offset into mel_ww

    std::map<int, FftTables> fft;              // keyed by NC

    // device copies
    float *d_window = nullptr, *d_mel_w = nullptr, *d_taps = nullptr;
    cfloat4 *d_decim_hh = nullptr;
    double *d_decim_h64 = nullptr, *d_decim_tw64 = nullptr, *d_taps64 = nullptr;
    int32_t *d_mel_start = nullptr, *d_mel_cnt = nullptr, *d_mel_off = nullptr;
    CqtRow *d_rows = nullptr;
    cfloat *d_weights = nullptr;
    CqtBlock4 *d_blocks = nullptr;
    cfloat4 *d_weights4 = nullptr;
    CqtItem *d_items = nullptr;                // sorted by nfft so each kernel instantiation sees a contiguous slice
    cfloat *d_mel_ww = nullptr;
    int32_t *d_mel_seg_start = nullptr;
    int32_t *d_mel_gsteps = nullptr, *d_mel_goff = nullptr;
    std::vector<void *> d_allocs;
    // VQT family: the decimation ladder, the sliding-DFT items and the exact-ladder tails run on a high-priority side stream
    // underneath the projection launches of the caller's stream.  Consecutive calls alternate between kCallSlots
    // (side stream, fork / join events) sets, so that the ladder of call i + 1 does not queue behind the sliding-DFT items
    // of call i; the host-side enqueue of a call is serialised per plan (call_mu), which makes reusing the events safe.
    static constexpr int kCallSlots = 2;
    void *side_stream[kCallSlots] = {};
    void *tail_stream[kCallSlots] = {};        // exact-ladder pieces (plans with alts): they read the audio only, so they run beside the shared ladder
    void *call_events[kCallSlots][5] = {};     // fork, mid, all, side join, tail join
    mutable unsigned call_next = 0;
    mutable std::mutex call_mu;

    // Pinned staging ring for the per-call clip descriptors: a cudaMemcpyAsync from pageable memory makes the host wait for
    // the stream to drain first, which would serialise a caller that queues an upload and then amtfeat_process behind it.
    static constexpr int kMetaSlots = 8;
    mutable char *meta_ring = nullptr;
    mutable size_t meta_slot_bytes = 0;
    mutable unsigned meta_next = 0;
    mutable void *meta_events[kMetaSlots] = {};

    // optional per-kernel timing (amtfeat_profile_*): CUDA event pairs recorded around every launch of
    // amtfeat_process on the launching stream (under call_mu, like the rest of a call's enqueue).
    struct ProfRec { std::string name; void *e0, *e1; };
    mutable bool prof_enabled = false;
    mutable std::vector<ProfRec> prof;
};

// ---- host_plan.cpp ----
void set_error(const std::string &msg);
int build_plan_tables(Plan &p);                 // fills every host table, validates the configuration
int64_t expected_frames(const Plan &p, int64_t n);
int64_t output_frames(const Plan &p, int64_t n);
void level_lengths(const Plan &p, int64_t n, int32_t *len /* kMaxLevels */);
void clip_tail_layout(const Plan &p, int64_t n, int64_t T_all, TailLayout &tl);
int64_t harmonic_frames(const Plan &p, int h, int64_t n);   // frames of harmonic h's own (untrimmed) VQT
int sample_range(const Plan &p, int64_t frames, int64_t *lo, int64_t *hi);
std::string describe(const Plan &p);
const char *decimator_name(const Plan &p);
std::string describe_clip(const Plan &p, int64_t n);

// ---- kernels_*.cu ----
int upload_plan(Plan &p);
void free_plan_device(Plan &p);
size_t workspace_bytes(const Plan &p, int batch, const int64_t *n);
int launch_count(const Plan &p, int batch, const int64_t *n);
int profile_read(const Plan &p, std::string &json);
int framify(const float *d_in, long long rows, long long T, int win, int hop, long long lpad, long long hops, float *d_out,
            void *stream);
int process(const Plan &p, const float *d_audio, const int64_t *in_off, const int64_t *n, const int64_t *out_off,
            int batch, float *d_out, void *d_ws, size_t ws_bytes, void *stream, bool defer_epilogue = false, int *max_frames = nullptr);
bool has_db_epilogue(const Plan &p);
size_t stft_smem_bytes(const Plan &p);   // dynamic shared memory of stft_kernel for this (STFT / MEL) configuration
int epilogue_out(const Plan &p, const float *d_out, void *d_ws, int batch, int maxT, float *dst, bool few_ctas, void *stream);
// chunks of a long track: reference level of a frame range of a raw block, and the epilogue with an external reference
int range_reference(const Plan &p, const float *d_block, int64_t frames, int64_t t_begin, int64_t t_end, float *d_ref, void *stream);
int range_finish(const Plan &p, const float *d_block, int64_t frames, int64_t t_begin, int64_t t_end, const float *d_ref, float *d_dst,
                 int64_t dst_frames, int64_t t_dst, void *stream);

}  // namespace amtfeat
