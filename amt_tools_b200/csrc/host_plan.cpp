// Host side of a plan: validation, constant-table design (window, mel filterbank, decimator taps,
// sparsified frequency-domain wavelet bases, FFT twiddles) and the integer / float64 frame arithmetic
// of the reference's FeatureModule API.  No CUDA here.
//
// The table design restates, in double precision, what the reference obtains from librosa at run
// time (paths under /root/reference/amt_tools/features/):
//   window        stft.py:66   librosa.stft(window='hann')           periodic Hann, centre-padded to n_fft
//   mel           mel.py:64    librosa.filters.mel(norm='slaney')    Slaney / HTK scale
//   wavelet basis vqt.py:183   librosa.vqt -> __vqt_filter_fft       wavelet(), FFT, sparsify_rows(0.01)
//   decimator     vqt.py:183   librosa.resample(res_type='soxr_hq')  Kaiser-windowed sinc, soxr HQ recipe
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <complex>
#include <cstdio>
#include <cstring>
#include <climits>
#include <sstream>
#include <tuple>

#include "plan.h"

namespace amtfeat {

static thread_local std::string g_error;
void set_error(const std::string &msg) { g_error = msg; }
const char *last_error_cstr() { return g_error.c_str(); }

static const double kPi = 3.14159265358979323846264338327950288;
static const double kHannBandwidth = 1.50018310546875;  // librosa.filters.window_bandwidth('hann')

static inline int64_t floordiv(int64_t a, int64_t b) {
    int64_t q = a / b, r = a % b;
    return (r != 0 && ((r < 0) != (b < 0))) ? q - 1 : q;
}
static inline bool is_pow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }
static int num_two_factors(int64_t x) {
    if (x <= 0) return 0;
    int n = 0;
    while ((x & 1) == 0) { ++n; x >>= 1; }
    return n;
}
static bool is_wave_kind(int k) {
    return k == AMTFEAT_WAVEFORM || k == AMTFEAT_STFT || k == AMTFEAT_MEL || k == AMTFEAT_POWER;
}

// ------------------------------------------------------------------------------------------------
// frame / sample arithmetic
// ------------------------------------------------------------------------------------------------

static int64_t vqt_expected(const Plan &p, int h, int64_t n) {
    // vqt.py:118-132 : int(min(ceil(n / 2^k) // (hop // 2^k)) + 1), k = eds .. eds + n_oct - 1
    const int eds = p.harm[h].eds_ref;
    double best = INFINITY;
    for (int k = eds; k < eds + p.n_oct; ++k) {
        double sig = std::ceil((double)n / std::ldexp(1.0, k));
        int64_t hopk = p.cfg.hop_length >> k;
        double hops = hopk > 0 ? std::floor(sig / (double)hopk) : INFINITY;
        best = std::min(best, hops + 1.0);
    }
    return std::isfinite(best) ? (int64_t)best : 0;
}

int64_t expected_frames(const Plan &p, int64_t n) {
    const amtfeat_config &c = p.cfg;
    if (is_wave_kind(c.kind)) {
        if (c.center || n == 0) return n == 0 ? 0 : 1 + n / c.hop_length;           // common.py:62-64
        return 1 + (floordiv(std::max<int64_t>(0, n - c.win_length) - 1, c.hop_length) + 1);  // waveform.py:64
    }
    int64_t best = INT64_MAX;
    for (size_t h = 0; h < p.harm.size(); ++h) best = std::min(best, vqt_expected(p, (int)h, n));  // hvqt.py:77-81
    return best;
}

void level_lengths(const Plan &p, int64_t n, int32_t *len) {
    int64_t cur = n;
    for (int l = 0; l < kMaxLevels; ++l) {
        len[l] = (int32_t)cur;
        cur = (cur + 1) / 2;  // librosa.resample output length ceil(n * ratio)
    }
}

static int64_t vqt_lib_frames(const Plan &p, int h, int64_t n) {
    if (n == 0) return 0;
    int32_t len[kMaxLevels];
    level_lengths(p, n, len);
    int64_t best = INT64_MAX;
    for (int i = 0; i < p.n_oct; ++i) {
        int l = p.harm[h].eds_lib + i;
        int64_t hopl = p.cfg.hop_length >> l;
        best = std::min<int64_t>(best, 1 + len[l] / hopl);  // centred STFT of the level signal
    }
    return best;
}

int64_t harmonic_frames(const Plan &p, int h, int64_t n) { return vqt_lib_frames(p, h, n); }

// Geometry of the exact ladders for one clip of n samples (T_all = frames the kernels compute).
//
// A harmonic with eds >= 2 is early-downsampled by librosa in ONE resample(2^eds -> 1) call, then decimated 2:1 from
// octave to octave; the shared ladder reaches the same levels by cascaded 2:1 steps.  In the pass band the two agree to
// filter ripple (< 2e-7 of the peak), but every cascaded step keeps only samples 0 .. ceil(len / 2) - 1 of its (zero
// phase) output, as librosa's own octave steps do: the filter's ringing before the first and after the last sample
// is dropped, whereas the one-shot filter carries it through.  With a 2:1 output m reading inputs 2m - D .. 2m + D,
// shared level l equals the exact one on [hsafe[l], dev[l]):
//     hsafe[1] = 0,      hsafe[l] = ceil((hsafe[l-1] + D) / 2)        (tends to D)
//     dev[1] = len[1],   dev[l]   = ceil((dev[l-1] - D) / 2)          (tends to len - D)
// Frames whose window leaves that interval -- the first th and the frames from t0 on -- are recomputed from an exact
// ladder of which only the head [0, hlen) and the tail [first, len) of every level are ever built: level eds in one
// 2^eds : 1 pass over the audio, deeper levels 2:1 from it.  When the two pieces of a level meet, the whole level is built
// (first = 0, hlen = -1: the head is read from the tail piece).
void clip_tail_layout(const Plan &p, int64_t n, int64_t T_all, TailLayout &tl) {
    for (int l = 0; l < kMaxLevels; ++l) {
        tl.t0[l] = INT32_MAX;
        tl.th[l] = 0;
        tl.dev[l] = INT32_MAX;
        tl.hsafe[l] = 0;
        for (int a = 0; a < kMaxAlt; ++a) { tl.first[a][l] = -1; tl.count[a][l] = 0; tl.hlen[a][l] = 0; }
    }
    if (p.alts.empty() || n <= 0) return;
    int32_t len[kMaxLevels];
    level_lengths(p, n, len);
    const int64_t D = ((int64_t)p.taps.size() - 1) / 2;
    int64_t dev = len[1], hs = 0;
    tl.dev[0] = len[0];
    tl.dev[1] = len[1];
    for (int l = 2; l < p.n_levels; ++l) {
        dev = dev - D <= 0 ? 0 : (dev - D + 1) / 2;
        hs = (hs + D + 1) / 2;
        tl.dev[l] = (int32_t)dev;
        tl.hsafe[l] = (int32_t)std::min<int64_t>(hs, len[l]);
    }
    for (int l = 0; l < p.n_levels; ++l) {
        if (p.alt_nfft_max[l] <= 0) continue;
        const int64_t hop = p.cfg.hop_length >> l, half = p.alt_nfft_max[l] / 2;
        // the frames of a tile / chunk of the kernel that holds the shared copy of these rows are skipped or kept together
        const int64_t A = std::max<int64_t>(32, 16384 / p.alt_nfft_min[l]);
        const int64_t room = (int64_t)tl.dev[l] - half;                           // frame t stays below dev iff t * hop <= room
        const int64_t first_alt = room < 0 ? 0 : room / hop + 1;
        int64_t t0 = first_alt / A * A;
        int64_t th = (tl.hsafe[l] + half + hop - 1) / hop;                         // frame t starts at or after hsafe iff t * hop >= hsafe + half
        th = (th + A - 1) / A * A;
        if (th >= t0) { th = 0; t0 = 0; }                                          // the two regions meet: every frame is exact
        tl.t0[l] = t0 >= T_all ? INT32_MAX : (int32_t)t0;
        tl.th[l] = (int32_t)std::min<int64_t>(th, (T_all + A - 1) / A * A);
    }
    for (size_t a = 0; a < p.alts.size(); ++a) {
        const int e = p.alts[a].eds;
        // head pieces, deepest level first: what the level's own head frames read, and what the next level's head needs
        int64_t hneed[kMaxLevels] = {};
        int64_t h_next = 0;
        for (int l = e + p.n_oct - 1; l >= e; --l) {
            const int64_t hop = p.cfg.hop_length >> l;
            int64_t h = tl.th[l] > 0 ? (int64_t)(tl.th[l] - 1) * hop + p.alt_nfft_max[l] / 2 : 0;
            if (h_next > 0) h = std::max(h, 2 * h_next + D - 1);
            h = std::min<int64_t>((h + 3) / 4 * 4, len[l]);
            hneed[l] = h;
            h_next = h;
        }
        int64_t need_next = -1;
        for (int l = e + p.n_oct - 1; l >= e; --l) {
            int64_t need = INT64_MAX;
            if (tl.t0[l] != INT32_MAX) {
                const int64_t hop = p.cfg.hop_length >> l;
                need = std::max<int64_t>(0, (int64_t)tl.t0[l] * hop - p.alt_nfft_max[l] / 2);
            }
            if (need_next >= 0) need = std::min(need, std::max<int64_t>(0, 2 * need_next - D));
            if (need != INT64_MAX) need = need / 4 * 4;
            if (need != INT64_MAX && hneed[l] >= need) {       // head and tail pieces meet: one piece, the whole level
                need = 0;
                tl.hlen[a][l] = hneed[l] > 0 ? -1 : 0;
            } else {
                tl.hlen[a][l] = (int32_t)hneed[l];
            }
            if (need == INT64_MAX) continue;
            tl.first[a][l] = (int32_t)need;
            tl.count[a][l] = (int32_t)std::max<int64_t>(0, (int64_t)len[l] - need);
            need_next = need;
        }
    }
}

static int64_t padded_uncentered(const Plan &p, int64_t n) {
    // common.py:141-166 frame_pad: pad to a multiple of win (if n <= win) else of hop
    const int64_t required = p.cfg.win_length;  // get_sample_range(1)[-1] of the non-centred wrapper
    int64_t divisor = n > required ? p.cfg.hop_length : required;
    return (n + divisor - 1) / divisor * divisor;
}

int64_t output_frames(const Plan &p, int64_t n) {
    const amtfeat_config &c = p.cfg;
    if (n == 0) return 0;
    switch (c.kind) {
        case AMTFEAT_STFT:
        case AMTFEAT_MEL: {
            int64_t np_ = c.center ? n + 2 * (c.n_fft / 2) : padded_uncentered(p, n);
            if (np_ < c.n_fft) return -1;  // librosa.stft raises for too-short uncentred input
            return 1 + (np_ - c.n_fft) / c.hop_length;
        }
        case AMTFEAT_WAVEFORM:
        case AMTFEAT_POWER: {
            int64_t np_ = c.center ? n + 2 * (c.win_length / 2) : padded_uncentered(p, n);
            if (np_ < c.win_length) return -1;
            return 1 + (np_ - c.win_length) / c.hop_length;
        }
        default: {
            // librosa __early_downsample: "Input signal length=%d is too short for %d-octave CQT" when a harmonic is downsampled
            // early by 2^eds and the clip is shorter than that factor (the reference's process_audio then raises)
            for (size_t h = 0; h < p.harm.size(); ++h)
                if (p.harm[h].eds_lib > 0 && n < ((int64_t)1 << p.harm[h].eds_lib)) return -1;
            int64_t best = c.kind == AMTFEAT_HVQT ? expected_frames(p, n) : INT64_MAX;  // hvqt.py:123-128 trims
            for (size_t h = 0; h < p.harm.size(); ++h) best = std::min(best, vqt_lib_frames(p, (int)h, n));
            return best;
        }
    }
}

int sample_range(const Plan &p, int64_t frames, int64_t *lo, int64_t *hi) {
    const amtfeat_config &c = p.cfg;
    const int64_t hop = c.hop_length;
    if (is_wave_kind(c.kind)) {
        if (c.center || frames == 0) {  // common.py:86-95
            if (frames <= 0) { *lo = *hi = 0; return AMTFEAT_OK; }
            *hi = frames * hop - 1;
            *lo = std::max<int64_t>(1, *hi - hop + 1);
        } else if (frames == 1) {  // waveform.py:88-90
            *lo = 1; *hi = c.win_length;
        } else {  // waveform.py:92-94
            int64_t base = c.win_length + (frames - 2) * hop;
            *lo = 1 + base; *hi = hop + base;
        }
        return AMTFEAT_OK;
    }
    // vqt.py:152-163 ; hvqt.py:103 uses the highest harmonic
    const int64_t f = (int64_t)1 << p.harm.back().eds_ref;
    *hi = (floordiv(frames * hop, f) - 1) * f;
    *lo = std::max<int64_t>(1, *hi - hop + 1);
    return AMTFEAT_OK;
}

// ------------------------------------------------------------------------------------------------
// table design
// ------------------------------------------------------------------------------------------------

// numpy's pairwise summation of a contiguous float32 vector (umath loops_utils pairwise_sum: 8 accumulators on blocks of at
// most 128, recursive halving above), so that np.sum(mags) is reproduced to the bit.
static float np_pairwise_sum_f32(const float *a, size_t n) {
    if (n < 8) {
        float res = 0.f;
        for (size_t i = 0; i < n; ++i) res += a[i];
        return res;
    }
    if (n <= 128) {
        float r[8];
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        size_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; ++j) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) res += a[i];
        return res;
    }
    size_t n2 = n / 2;
    n2 -= n2 % 8;
    return np_pairwise_sum_f32(a, n2) + np_pairwise_sum_f32(a + n2, n - n2);
}

static void fft_inplace(std::vector<std::complex<double>> &a) {  // forward DFT, radix-2, n power of two
    const size_t n = a.size();
    for (size_t i = 1, j = 0; i < n; ++i) {
        size_t bit = n >> 1;
        for (; j & bit; bit >>= 1) j ^= bit;
        j ^= bit;
        if (i < j) std::swap(a[i], a[j]);
    }
    for (size_t len = 2; len <= n; len <<= 1) {
        for (size_t i = 0; i < n; i += len)
            for (size_t k = 0; k < len / 2; ++k) {
                double ang = -2.0 * kPi * (double)k / (double)len;
                std::complex<double> w(std::cos(ang), std::sin(ang));
                std::complex<double> u = a[i + k], v = a[i + k + len / 2] * w;
                a[i + k] = u + v;
                a[i + k + len / 2] = u - v;
            }
    }
}

static void build_window(Plan &p) {
    const int n_fft = p.cfg.n_fft, win = p.cfg.win_length;
    p.window.assign(n_fft, 0.f);
    const int lpad = (n_fft - win) / 2;
    for (int i = 0; i < win; ++i) {
        double w = win == 1 ? 1.0 : 0.5 - 0.5 * std::cos(2.0 * kPi * (double)i / (double)win);
        p.window[lpad + i] = (float)w;
    }
}

static double hz_to_mel(double f, bool htk) {
    if (htk) return 2595.0 * std::log10(1.0 + f / 700.0);
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return f >= min_log_hz ? min_log_mel + std::log(f / min_log_hz) / logstep : f / f_sp;
}
static double mel_to_hz(double m, bool htk) {
    if (htk) return 700.0 * (std::pow(10.0, m / 2595.0) - 1.0);
    const double f_sp = 200.0 / 3, min_log_hz = 1000.0, min_log_mel = min_log_hz / f_sp, logstep = std::log(6.4) / 27.0;
    return m >= min_log_mel ? min_log_hz * std::exp(logstep * (m - min_log_mel)) : f_sp * m;
}

static void build_mel(Plan &p) {
    const amtfeat_config &c = p.cfg;
    const int n_mels = c.n_mels, nb = c.n_fft / 2 + 1;
    const bool htk = c.htk != 0;
    const double fmax = c.sample_rate / 2.0;
    std::vector<double> mel_f(n_mels + 2);
    const double m_lo = hz_to_mel(0.0, htk), m_hi = hz_to_mel(fmax, htk);
    for (int i = 0; i < n_mels + 2; ++i) {
        double m = (i == n_mels + 1) ? m_hi : m_lo + (m_hi - m_lo) * (double)i / (double)(n_mels + 1);
        mel_f[i] = mel_to_hz(m, htk);
    }
    p.mel_start.assign(n_mels, 0);
    p.mel_cnt.assign(n_mels, 0);
    p.mel_off.assign(n_mels, 0);
    p.mel_w.clear();
    for (int i = 0; i < n_mels; ++i) {
        const double fd0 = mel_f[i + 1] - mel_f[i], fd1 = mel_f[i + 2] - mel_f[i + 1];
        const double enorm = 2.0 / (mel_f[i + 2] - mel_f[i]);
        int first = -1, last = -1;
        std::vector<float> row(nb);
        for (int k = 0; k < nb; ++k) {
            double fk = (double)k * c.sample_rate / (double)c.n_fft;  // rfftfreq
            double lower = -(mel_f[i] - fk) / fd0, upper = (mel_f[i + 2] - fk) / fd1;
            float w = (float)std::max(0.0, std::min(lower, upper));   // weights array is float32
            w = (float)((double)w * enorm);                           // weights *= enorm[:, newaxis]
            row[k] = w;
            if (w != 0.f) { if (first < 0) first = k; last = k; }
        }
        p.mel_off[i] = (int32_t)p.mel_w.size();
        if (first >= 0) {
            p.mel_start[i] = first;
            p.mel_cnt[i] = last - first + 1;
            p.mel_w.insert(p.mel_w.end(), row.begin() + first, row.begin() + last + 1);
        }
    }
    // segment form (see plan.h): every FFT bin has at most two non-zero filters, m and m + 1
    const int nseg = n_mels + 1;
    std::vector<int> seg_of(nb, -1);
    std::vector<float> up(nb, 0.f), down(nb, 0.f);
    {
        std::vector<std::vector<std::pair<int, float>>> nz(nb);
        for (int i = 0; i < n_mels; ++i)
            for (int j = 0; j < p.mel_cnt[i]; ++j) {
                float w = p.mel_w[p.mel_off[i] + j];
                if (w != 0.f) nz[p.mel_start[i] + j].push_back({i, w});
            }
        for (int k = 0; k < nb; ++k) {
            const double fk = (double)k * c.sample_rate / (double)c.n_fft;
            if (nz[k].size() == 2 && nz[k][1].first == nz[k][0].first + 1) {
                seg_of[k] = nz[k][1].first;
                up[k] = nz[k][1].second;
                down[k] = nz[k][0].second;
            } else if (nz[k].size() == 1) {
                const int i = nz[k][0].first;
                if (fk < mel_f[i + 1]) { seg_of[k] = i; up[k] = nz[k][0].second; }
                else { seg_of[k] = i + 1; down[k] = nz[k][0].second; }
            } else if (!nz[k].empty()) {
                p.mel_ww.clear();
                p.mel_seg_start.clear();
                return;  // build_plan_tables reports the unsupported filterbank
            }
        }
    }
    // bins with no filter inherit the previous segment with zero weights so that segments stay contiguous
    int cur = 0;
    for (int k = 0; k < nb; ++k) {
        if (seg_of[k] < 0) seg_of[k] = cur;
        if (seg_of[k] < cur) { p.mel_ww.clear(); p.mel_seg_start.clear(); return; }
        cur = seg_of[k];
    }
    p.mel_seg_start.assign(nseg, nb);
    std::vector<int> seg_cnt(nseg, 0);
    for (int k = nb - 1; k >= 0; --k) { p.mel_seg_start[seg_of[k]] = k; seg_cnt[seg_of[k]]++; }
    for (int sgi = 0; sgi < nseg; ++sgi) if (seg_cnt[sgi] == 0) p.mel_seg_start[sgi] = 0;
    const int ngroups = (nseg + 31) / 32;
    p.mel_gsteps.assign(ngroups, 0);
    p.mel_goff.assign(ngroups, 0);
    p.mel_ww.clear();
    for (int gidx = 0; gidx < ngroups; ++gidx) {
        int steps = 0;
        for (int l = 0; l < 32 && 32 * gidx + l < nseg; ++l) steps = std::max(steps, seg_cnt[32 * gidx + l]);
        p.mel_gsteps[gidx] = steps;
        p.mel_goff[gidx] = (int32_t)p.mel_ww.size();
        p.mel_ww.resize(p.mel_ww.size() + (size_t)std::max(steps, 1) * 32, cfloat{0.f, 0.f});
        for (int l = 0; l < 32 && 32 * gidx + l < nseg; ++l) {
            const int sgi = 32 * gidx + l;
            for (int j = 0; j < seg_cnt[sgi]; ++j) {
                const int k = p.mel_seg_start[sgi] + j;
                p.mel_ww[(size_t)p.mel_goff[gidx] + (size_t)j * 32 + l] = cfloat{up[k], down[k]};
            }
        }
    }
}

static double bessel_i0(double x) {
    double sum = 1.0, term = 1.0, q = x * x / 4.0;
    for (int k = 1; k < 500; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}

// soxr 'HQ' 2:1 decimator (libsoxr 0.1.3: soxr_quality_spec + lsx_design_lpf): 20-bit precision,
// passband to (1 - .05 / TO_3dB(rej)) * Nyq_out, stopband from Nyq_out, Kaiser window.
static std::vector<double> design_decimator(int factor = 2) {
    static const double coefs[10][4] = {
        {-6.784957e-10, 1.02856e-05, 0.1087556, -0.8988365 + .001}, {-6.897885e-10, 1.027433e-05, 0.10876, -0.8994658 + .002},
        {-1.000683e-09, 1.030092e-05, 0.1087677, -0.9007898 + .003}, {-3.654474e-10, 1.040631e-05, 0.1087085, -0.8977766 + .006},
        {8.106988e-09, 6.983091e-06, 0.1091387, -0.9172048 + .015},  {9.519571e-09, 7.272678e-06, 0.1090068, -0.9140768 + .025},
        {-5.626821e-09, 1.342186e-05, 0.1083999, -0.9065452 + .05},  {-9.965946e-08, 5.073548e-05, 0.1040967, -0.7672778 + .085},
        {1.604808e-07, -5.856462e-05, 0.1185998, -1.34824 + .1},     {-1.511964e-07, 6.363034e-05, 0.1064627, -0.9876665 + .18}};
    const double bits = 20.0, l2db = 20.0 * std::log10(2.0);
    const double rej = bits * l2db;
    const double to3db = (1.6e-6 * rej - 7.5e-4) * rej + .646;
    const double att = (bits + 1) * l2db;
    double Fp = (1.0 - .05 / to3db) / (double)factor, Fs = 1.0 / (double)factor;  // relative to the input Nyquist
    double tr_bw = std::min(.5 * (Fs - Fp), .5 * Fs);
    const double Fc = Fs - tr_bw;
    const double realm = std::log(tr_bw * .5 / Fc / .0005) / std::log(2.0);
    const int i0 = std::min(9, std::max(0, (int)realm)), i1 = std::min(9, std::max(0, 1 + (int)realm));
    const double b0 = ((coefs[i0][0] * att + coefs[i0][1]) * att + coefs[i0][2]) * att + coefs[i0][3];
    const double b1 = ((coefs[i1][0] * att + coefs[i1][1]) * att + coefs[i1][2]) * att + coefs[i1][3];
    const double beta = b0 + (b1 - b0) * (realm - (int)realm);
    const double a = ((.0007528358 - 1.577737e-05 * beta) * beta + .6248022) * beta + .06186902;
    int num_taps = (int)std::ceil(a / tr_bw + 1);
    num_taps = (num_taps + 4 - 2) / 4 * 4 + 1;  // 1 (mod 4)
    const int m = num_taps - 1;
    const double mult1 = 1.0 / (.5 * m + .5), i0b = bessel_i0(beta);
    std::vector<double> h(num_taps);
    double sum = 0;
    for (int i = 0; i < num_taps; ++i) {
        double z = i - .5 * m, x = z * kPi, y = z * mult1;
        double s = x != 0 ? std::sin(Fc * x) / x : Fc;
        h[i] = s * bessel_i0(beta * std::sqrt(std::max(0.0, 1 - y * y))) / i0b;
        sum += h[i];
    }
    for (double &v : h) v /= sum;
    return h;
}

static void build_fft_tables(Plan &p, int nfft) {
    const int NC = nfft / 2;
    if (p.fft.count(NC)) return;
    int R1, R2;
    switch (NC) {
        case 1024: R1 = 32; R2 = 32; break;
        case 512: R1 = 16; R2 = 32; break;
        case 256: R1 = 16; R2 = 16; break;
        case 128: R1 = 8; R2 = 16; break;
        case 64: R1 = 8; R2 = 8; break;
        case 32: R1 = 4; R2 = 8; break;
        case 16: R1 = 4; R2 = 4; break;
        case 8: R1 = 2; R2 = 4; break;
        default: R1 = 2; R2 = 2; break;  // 4
    }
    FftTables t;
    t.tw1.resize(NC);
    for (int k1 = 0; k1 < R1; ++k1)
        for (int n2 = 0; n2 < R2; ++n2) {
            double ang = -2.0 * kPi * (double)(k1 * n2) / (double)NC;
            t.tw1[k1 * R2 + n2] = {(float)std::cos(ang), (float)std::sin(ang)};
        }
    t.tw2.assign(NC + 2, cfloat{0.f, 0.f});   // k = 0 .. NC, padded to a whole number of 16-byte granules
    for (int k = 0; k <= NC; ++k) {
        double ang = -kPi * (double)k / (double)NC;
        t.tw2[k] = {(float)std::cos(ang), (float)std::sin(ang)};
    }
    p.fft[NC] = std::move(t);
}

static int build_vqt(Plan &p) {
    const amtfeat_config &c = p.cfg;
    const int bpo = c.bins_per_octave, n_bins = c.n_bins;
    if (n_bins < 1 || bpo < 1) { set_error("n_bins and bins_per_octave must be positive"); return AMTFEAT_ERR_INVALID; }
    if (c.fmin <= 0 || c.gamma < 0) { set_error("fmin must be positive and gamma non-negative"); return AMTFEAT_ERR_INVALID; }
    p.n_oct = (int)std::ceil((double)n_bins / bpo);
    p.n_filters = std::min(bpo, n_bins);
    const double sr0 = c.sample_rate, nyq = sr0 / 2.0;
    const double r = std::pow(2.0, 1.0 / bpo);
    const double alpha_lib = (r * r - 1) / (r * r + 1);       // librosa >= 0.10 relative bandwidth
    const double alpha_ref = std::pow(2.0, 1.0 / bpo) - 1;    // vqt.py:49 (reference's own, old convention)
    const double Q = 1.0 / alpha_lib;
    if (num_two_factors(c.hop_length) < p.n_oct - 1) {
        char b[160];
        snprintf(b, sizeof b, "hop_length must be a positive integer multiple of 2^%d for %d-octave CQT/VQT", p.n_oct - 1, p.n_oct);
        set_error(b);
        return AMTFEAT_ERR_INVALID;
    }
    std::vector<double> tapsd;
    if (c.n_decim_taps > 0 && c.decim_taps) {
        if (c.n_decim_taps % 2 == 0) { set_error("decimator taps must have odd length (linear phase, integer delay)"); return AMTFEAT_ERR_INVALID; }
        tapsd.assign(c.decim_taps, c.decim_taps + c.n_decim_taps);
    } else {
        tapsd = design_decimator();
    }
    if (((tapsd.size() - 1) / 2) % 2 != 0) {  // keep the group delay even so both polyphase branches are integer-aligned
        tapsd.insert(tapsd.begin(), 0.0);
        tapsd.push_back(0.0);
    }
    p.taps.resize(tapsd.size());
    p.taps64.resize(tapsd.size());
    for (size_t i = 0; i < tapsd.size(); ++i) {
        p.taps64[i] = tapsd[i] * std::sqrt(2.0);
        p.taps[i] = (float)p.taps64[i];
    }
    // frequency response of the (float32) taps for the fast-convolution decimator: 2048-point blocks, every other
    // output kept => the 2048-point spectrum folds onto 1024 points; a block yields 1024 - D outputs
    p.decim_hh.clear();
    {
        const char *env = std::getenv("AMTFEAT_DECIM");
        p.decim_mode = !env ? 0 : std::string(env) == "direct" ? 2 : std::string(env) == "fft32" ? 1 : 0;
        const char *env_s = std::getenv("AMTFEAT_SLIDE");
        p.slide_off = env_s && std::string(env_s) == "0";
        const char *env_q = std::getenv("AMTFEAT_SERIAL");
        p.serial_launch = env_q && std::string(env_q) == "1";
        const char *env_x = std::getenv("AMTFEAT_EXACT_EDS");
        p.exact_eds = !(env_x && std::string(env_x) == "0");
        const int nt = (int)p.taps.size(), D = (nt - 1) / 2;
        if (1024 - D >= 256) {
            build_fft_tables(p, 2048);
            std::vector<std::complex<double>> H(1025);
            for (int k = 0; k <= 1024; ++k) {
                std::complex<double> acc(0, 0);
                for (int j = 0; j < nt; ++j) {
                    const double ang = -2.0 * kPi * (double)((long long)j * k % 2048) / 2048.0;
                    acc += (double)p.taps[j] * std::complex<double>(std::cos(ang), std::sin(ang));
                }
                H[k] = acc / 2048.0;   // 1/2 of the fold, 1/1024 of the inverse transform
            }
            p.decim_hh.resize(513);
            for (int k = 0; k <= 512; ++k)
                p.decim_hh[k] = cfloat4{(float)H[k].real(), (float)H[k].imag(), (float)H[1024 - k].real(), (float)-H[1024 - k].imag()};
            // float64 form (decimate_fft64_kernel): the full 2048-point response of the (unrounded) taps, Hermitian
            // extension, same 1 / 2048, and exp(-2 pi i m / 2048), m < 1024, for its transforms
            std::vector<std::complex<double>> H64(1025);
            for (int k = 0; k <= 1024; ++k) {
                std::complex<double> acc(0, 0);
                for (int j = 0; j < nt; ++j) {
                    const double ang = -2.0 * kPi * (double)((long long)j * k % 2048) / 2048.0;
                    acc += tapsd[j] * std::sqrt(2.0) * std::complex<double>(std::cos(ang), std::sin(ang));   // the float64 design, unrounded
                }
                H64[k] = acc / 2048.0;
            }
            p.decim_h64.resize(2 * 2048);
            for (int k = 0; k < 2048; ++k) {
                const std::complex<double> v = k <= 1024 ? H64[k] : std::conj(H64[2048 - k]);
                p.decim_h64[2 * k] = v.real();
                p.decim_h64[2 * k + 1] = v.imag();
            }
            // per-pass twiddle tables of its Stockham transforms (kernels.cu decimate_fft64_kernel), each [r - 1][k] =
            // exp(-2 pi i r k / (R Ns)), k < Ns: forward 2048 = 16 x 16 x 8 (passes B, C), inverse 1024 = 8 x 8 x 16 (passes E, F)
            p.decim_tw64.clear();
            auto push_tw = [&](long long num, long long den) {   // exp(-2 pi i num / den), exact at the octant points
                num %= den;
                const double ang = -2.0 * kPi * (double)num / (double)den;
                double c = std::cos(ang), sn = std::sin(ang);
                if ((4 * num) % den == 0) {
                    const int q = (int)(4 * num / den);
                    c = q == 0 ? 1.0 : q == 2 ? -1.0 : 0.0;
                    sn = q == 1 ? -1.0 : q == 3 ? 1.0 : 0.0;
                }
                p.decim_tw64.push_back(c);
                p.decim_tw64.push_back(sn);
            };
            auto push_pass = [&](int R, int Ns) {
                for (int r = 1; r < R; ++r)
                    for (int k = 0; k < Ns; ++k) push_tw((long long)r * k, (long long)R * Ns);
            };
            push_pass(16, 16);    // B:  240 entries at 0
            push_pass(8, 256);    // C: 1792 entries at 240
            push_pass(8, 8);      // E:   56 entries at 2032
            push_pass(16, 64);    // F:  960 entries at 2088
        }
    }

    p.harm.clear();
    p.alts.clear();
    p.alt_mask = 0;
    for (int l = 0; l < kMaxLevels; ++l) p.alt_nfft_max[l] = p.alt_nfft_min[l] = 0;
    p.n_levels = 0;
    std::map<std::tuple<int, int, int>, std::vector<CqtRow>> groups;  // (0 = shared ladder | a + 1 = exact ladder a, nfft, level) -> rows
    for (int h = 0; h < c.n_harmonics; ++h) {
        HarmonicInfo hi;
        hi.fmin = c.harmonics[h] * c.fmin;  // hvqt.py:47
        std::vector<double> freqs(n_bins);
        for (int i = 0; i < n_bins; ++i) freqs[i] = hi.fmin * std::pow(2.0, (double)i / bpo);
        const double fmax = freqs[n_bins - 1];
        // reference: vqt.py:81-98
        const double cutoff_ref = fmax * (1 + 0.5 * kHannBandwidth * alpha_ref) + 0.5 * c.gamma;
        // librosa.filters.wavelet_lengths
        const double cutoff_lib = fmax * (1 + 0.5 * kHannBandwidth / Q) + 0.5 * c.gamma;
        if (cutoff_lib > nyq) {
            char b[200];
            snprintf(b, sizeof b, "Wavelet basis with max frequency=%g would exceed the Nyquist frequency=%g. Try reducing the number of frequency bins.", fmax, nyq);
            set_error(b);
            return AMTFEAT_ERR_INVALID;
        }
        auto eds_of = [&](double cutoff) {
            int c1 = std::max(0, (int)(std::ceil(std::log2(nyq / cutoff)) - 1) - 1);
            int c2 = std::max(0, num_two_factors(c.hop_length) - p.n_oct + 1);
            return std::min(c1, c2);
        };
        hi.eds_ref = eds_of(cutoff_ref);
        hi.eds_lib = eds_of(cutoff_lib);
        hi.alt = -1;
        if (hi.eds_lib >= 2 && p.exact_eds) {
            // librosa downsamples this harmonic by 2^eds in ONE resample call (constantq.py __early_downsample): its tail
            // frames come from an exact ladder (clip_tail_layout); harmonics with the same eds share one
            for (size_t a = 0; a < p.alts.size(); ++a)
                if (p.alts[a].eds == hi.eds_lib) hi.alt = (int)a;
            if (hi.alt < 0) {
                if ((int)p.alts.size() >= kMaxAlt) { set_error("too many distinct early-downsampling factors among the harmonics"); return AMTFEAT_ERR_INVALID; }
                AltLadder al;
                al.eds = hi.eds_lib;
                const int factor = 1 << hi.eds_lib;
                std::vector<double> t = design_decimator(factor);
                al.taps.resize(t.size());
                for (size_t i = 0; i < t.size(); ++i) al.taps[i] = t[i] * std::sqrt((double)factor);
                hi.alt = (int)p.alts.size();
                p.alts.push_back(std::move(al));
            }
            p.alt_mask |= 1u << h;
        }
        p.harm.push_back(hi);
        const int eds = hi.eds_lib;
        const double sr_post = sr0 / std::ldexp(1.0, eds);
        if (eds + p.n_oct > kMaxLevels) { set_error("too many ladder levels"); return AMTFEAT_ERR_INVALID; }
        p.n_levels = std::max(p.n_levels, eds + p.n_oct);
        for (int i = 0; i < p.n_oct; ++i) {
            const int lo = std::max(0, n_bins - p.n_filters * (i + 1)), hi_bin = n_bins - p.n_filters * i;
            const int level = eds + i;
            const double my_sr = sr0 / std::ldexp(1.0, level);
            double max_len = 0;
            std::vector<double> lens(hi_bin - lo);
            for (int k = lo; k < hi_bin; ++k) {
                lens[k - lo] = Q * my_sr / (freqs[k] + c.gamma / alpha_lib);
                max_len = std::max(max_len, lens[k - lo]);
            }
            const int nfft = (int)std::ldexp(1.0, (int)std::ceil(std::log2(max_len)));
            if (nfft < 8 || nfft > 2048) {
                char b[160];
                snprintf(b, sizeof b, "octave %d of harmonic %d needs n_fft=%d; supported range is 8..2048", i, h, nfft);
                set_error(b);
                return AMTFEAT_ERR_INVALID;
            }
            build_fft_tables(p, nfft);
            const double oct_scale = std::sqrt(sr_post / my_sr);  // fft_basis *= sqrt(sr / my_sr)
            for (int k = lo; k < hi_bin; ++k) {
                const double ilen = lens[k - lo];
                // n = arange(-ilen // 2, ilen // 2): floor(-ilen/2) .. floor(ilen/2) - 1
                const long n0 = (long)std::floor(-ilen / 2.0), n1 = (long)std::floor(ilen / 2.0);
                const int m = (int)(n1 - n0);
                std::vector<std::complex<double>> sig(m);
                double l1 = 0;
                for (int j = 0; j < m; ++j) {
                    double ph = (double)(n0 + j) * 2.0 * kPi * freqs[k] / my_sr;
                    double w = m == 1 ? 1.0 : 0.5 - 0.5 * std::cos(2.0 * kPi * (double)j / (double)m);
                    sig[j] = std::complex<double>(std::cos(ph), std::sin(ph)) * w;
                    l1 += std::abs(sig[j]);
                }
                std::vector<std::complex<double>> a(nfft, 0.0);
                const int lpad = (nfft - m) / 2;
                const double renorm = ilen / (double)nfft;  // basis *= lengths / n_fft
                for (int j = 0; j < m; ++j) {
                    std::complex<double> v = sig[j] / l1;
                    std::complex<float> v32((float)v.real(), (float)v.imag());           // dtype=complex64
                    std::complex<double> s = std::complex<double>(v32.real(), v32.imag()) * renorm;
                    a[lpad + j] = std::complex<double>((float)s.real(), (float)s.imag());  // stays complex64
                }
                fft_inplace(a);
                const int nb = nfft / 2 + 1;
                // sparsify_rows(quantile = 0.01) in the arithmetic librosa runs it in: fft_basis is complex64, so np.abs, np.sum
                // (pairwise), the division and np.cumsum (sequential) are all float32.  The threshold decision sits on a
                // cumulative sum of ~1000 terms: for some rows the 0.01 crossing is decided by 1e-6 of its value, which float64
                // arithmetic here would settle differently from the float32 the reference (and the oracle) uses.
#ifdef AMT_SPARSIFY_F64
                typedef double sp_t;
#else
                typedef float sp_t;
#endif
                std::vector<sp_t> mags(nb);
                for (int q = 0; q < nb; ++q) {
                    a[q] = std::complex<double>((float)a[q].real(), (float)a[q].imag());
                    mags[q] = sizeof(sp_t) == 4 ? (sp_t)hypotf((float)a[q].real(), (float)a[q].imag()) : (sp_t)std::abs(a[q]);
                }
                sp_t norm = 0;
                if (sizeof(sp_t) == 4) {
                    std::vector<float> mf(mags.begin(), mags.end());
                    norm = (sp_t)np_pairwise_sum_f32(mf.data(), (size_t)nb);
                } else {
                    for (int q = 0; q < nb; ++q) norm += mags[q];
                }
                std::vector<sp_t> srt(mags);
                std::sort(srt.begin(), srt.end());
                sp_t cum = 0, thr = srt.back();
                for (int q = 0; q < nb; ++q) {
                    cum += srt[q] / norm;
                    if (!(cum < (sp_t)0.01)) { thr = srt[q]; break; }
                }
                int first = -1, last = -1;
                for (int q = 0; q < nb; ++q)
                    if (mags[q] >= thr) { if (first < 0) first = q; last = q; }
                CqtRow row;
                row.chan = h;
                row.bin = k;
                const double len_post = Q * sr_post / (freqs[k] + c.gamma / alpha_lib);  // lengths at the post-eds rate
                row.inv_len = (float)(1.0 / len_post);
                row.col0 = first;
                row.cnt = last - first + 1;
                row.woff = (int32_t)p.weights.size();
                for (int q = first; q <= last; ++q) {
                    std::complex<double> w = mags[q] >= thr ? a[q] * oct_scale : std::complex<double>(0, 0);
                    p.weights.push_back({(float)w.real(), (float)w.imag()});
                }
                groups[std::make_tuple(0, nfft, level)].push_back(row);
                if (hi.alt >= 0) {
                    groups[std::make_tuple(hi.alt + 1, nfft, level)].push_back(row);
                    p.alt_nfft_max[level] = std::max(p.alt_nfft_max[level], nfft);
                    p.alt_nfft_min[level] = p.alt_nfft_min[level] ? std::min(p.alt_nfft_min[level], nfft) : nfft;
                }
            }
        }
    }
    p.rows.clear();
    p.items.clear();
    p.item_kmax_true.clear();
    p.blocks.clear();
    p.weights4.clear();
    for (auto &g : groups) {
        CqtItem it{};
        it.alt = std::get<0>(g.first);
        it.nfft = std::get<1>(g.first);
        it.level = std::get<2>(g.first);
        it.hop = c.hop_length >> it.level;
        it.nrows = (int32_t)g.second.size();
        it.kmin = INT32_MAX;
        it.kmax = 0;
        const int NC = it.nfft / 2;
        // Rows of different harmonics that are the same wavelet (same band, same weights once the octave scale
        // sqrt(sr / my_sr) and 1 / length are folded together) are projected once: `uniq` keeps the first
        // occurrence and the list of (chan, bin) destinations it serves.
        const std::vector<CqtRow> &rs = g.second;
        struct Uniq { size_t row; std::vector<std::pair<int, int>> dst; };
        std::vector<Uniq> uniq;
        for (size_t i = 0; i < rs.size(); ++i) {
            int found = -1;
            for (size_t u = 0; u < uniq.size() && found < 0; ++u) {
                const CqtRow &a = rs[uniq[u].row], &b = rs[i];
                if (a.chan == b.chan || a.col0 != b.col0 || a.cnt != b.cnt || (int)uniq[u].dst.size() >= kMaxDst) continue;
                const double sa = std::sqrt((double)a.inv_len), sb = std::sqrt((double)b.inv_len);
                bool same = true;
                double ref = 0;
                for (int q = 0; q < a.cnt; ++q) ref = std::max(ref, std::hypot((double)p.weights[a.woff + q].x, (double)p.weights[a.woff + q].y) * sa);
                for (int q = 0; q < a.cnt && same; ++q) {
                    const cfloat wa = p.weights[a.woff + q], wb = p.weights[b.woff + q];
                    if (std::fabs(wa.x * sa - wb.x * sb) > 4e-7 * ref || std::fabs(wa.y * sa - wb.y * sb) > 4e-7 * ref) same = false;
                    if ((wa.x == 0.f && wa.y == 0.f) != (wb.x == 0.f && wb.y == 0.f)) same = false;  // identical kept set
                }
                if (same) found = (int)u;
            }
            if (found >= 0) uniq[found].dst.push_back({rs[i].chan, rs[i].bin});
            else uniq.push_back({i, {{rs[i].chan, rs[i].bin}}});
        }
        // group up to four adjacent unique rows (adjacent bins in every destination) into blocks
        it.blk0 = (int32_t)p.blocks.size();
        auto adjacent = [&](const Uniq &a, const Uniq &b, int d) {
            if (a.dst.size() != b.dst.size()) return false;
            for (size_t k = 0; k < a.dst.size(); ++k)
                if (b.dst[k].first != a.dst[k].first || b.dst[k].second != a.dst[k].second + d) return false;
            return true;
        };
        for (size_t i = 0; i < uniq.size();) {
            size_t n = 1;
            while (n < 4 && i + n < uniq.size() && adjacent(uniq[i], uniq[i + n], (int)n)) ++n;
            CqtBlock4 bl{};
            int lo = INT32_MAX, hi = 0;
            for (size_t r = 0; r < n; ++r) {
                const CqtRow &rw = rs[uniq[i + r].row];
                lo = std::min(lo, rw.col0);
                hi = std::max(hi, rw.col0 + rw.cnt);
            }
            hi = std::min(hi, NC + 1);
            bl.col0 = lo;
            bl.steps = hi - lo;
            bl.woff = (int32_t)p.weights4.size();
            bl.ndst = (int32_t)uniq[i].dst.size();
            for (int d = 0; d < kMaxDst; ++d) {
                bl.chan[d] = d < bl.ndst ? uniq[i].dst[d].first : 0;
                for (int r = 0; r < 4; ++r)
                    bl.off[d][r] = (d < bl.ndst && r < (int)n) ? uniq[i + r].dst[d].first * p.F + uniq[i + r].dst[d].second : -1;
            }
            for (int r = 0; r < 4; ++r) bl.inv[r] = r < (int)n ? rs[uniq[i + r].row].inv_len : 0.f;
            p.weights4.resize(p.weights4.size() + (size_t)bl.steps * 2, cfloat4{0, 0, 0, 0});
            for (int st = 0; st < bl.steps; ++st) {
                const int col = lo + st;
                float w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                for (size_t r = 0; r < n; ++r) {
                    const CqtRow &rw = rs[uniq[i + r].row];
                    if (col >= rw.col0 && col < rw.col0 + rw.cnt) {
                        const cfloat v = p.weights[rw.woff + (col - rw.col0)];
                        w[2 * r] = v.x;
                        w[2 * r + 1] = v.y;
                    }
                }
                // row pairs side by side: (re0, re1, im0, im1), (re2, re3, im2, im3) -- operands of the packed FFMA2 projection
                p.weights4[(size_t)bl.woff + 2 * st] = cfloat4{w[0], w[2], w[1], w[3]};
                p.weights4[(size_t)bl.woff + 2 * st + 1] = cfloat4{w[4], w[6], w[5], w[7]};
            }
            it.kmin = std::min(it.kmin, lo);
            it.kmax = std::max(it.kmax, hi - 1);
            p.blocks.push_back(bl);
            i += n;
        }
        it.nblk = (int32_t)p.blocks.size() - it.blk0;
        it.nuniq = (int32_t)uniq.size();
        it.woff0 = it.nblk ? p.blocks[it.blk0].woff : 0;
        it.wcount = (int32_t)p.weights4.size() - it.woff0;
        int ktrue = 0;
        it.row0 = (int32_t)p.rows.size();
        for (const CqtRow &rw : rs) {
            p.rows.push_back(rw);
            ktrue = std::max(ktrue, rw.col0 + rw.cnt - 1);
        }
        p.item_kmax_true.push_back(ktrue);
        it.kmax_true = ktrue;
        p.items.push_back(it);
    }
    return AMTFEAT_OK;
}

int build_plan_tables(Plan &p) {
    {
        const char *env_m = std::getenv("AMTFEAT_META_MEMCPY");
        p.meta_memcpy = env_m && std::string(env_m) == "1";
    }
    amtfeat_config &c = p.cfg;
    if (c.hop_length <= 0) { set_error("hop_length must be a positive integer"); return AMTFEAT_ERR_INVALID; }
    if (!(c.sample_rate > 0)) { set_error("sample_rate must be positive"); return AMTFEAT_ERR_INVALID; }
    switch (c.kind) {
        case AMTFEAT_WAVEFORM:
        case AMTFEAT_POWER:
            if (c.win_length <= 0) { set_error("win_length must be positive"); return AMTFEAT_ERR_INVALID; }
            p.C = 1;
            p.F = c.kind == AMTFEAT_POWER ? 1 : c.win_length;
            return AMTFEAT_OK;
        case AMTFEAT_STFT:
        case AMTFEAT_MEL:
            if (!is_pow2(c.n_fft) || c.n_fft < 8 || c.n_fft > 2048) {
                set_error("n_fft must be a power of two in [8, 2048]");
                return AMTFEAT_ERR_INVALID;
            }
            if (c.win_length <= 0 || c.win_length > c.n_fft) { set_error("win_length must be in [1, n_fft]"); return AMTFEAT_ERR_INVALID; }
            p.C = 1;
            build_window(p);
            build_fft_tables(p, c.n_fft);
            if (c.kind == AMTFEAT_MEL) {
                if (c.n_mels <= 0) { set_error("n_mels must be positive"); return AMTFEAT_ERR_INVALID; }
                build_mel(p);
                if (p.mel_ww.empty()) { set_error("mel filterbank is not a chain of overlapping triangles (n_mels too large for n_fft?)"); return AMTFEAT_ERR_INVALID; }
                p.F = c.n_mels;
            } else {
                p.F = c.n_fft / 2 + 1;
            }
            // the kernel's shared-memory tiles depend on the configuration alone: reject what cannot run when the module is
            // constructed (a ValueError there), not at the first process_audio call
            if (stft_smem_bytes(p) > 227 * 1024) {
                set_error("hop_length / n_mels too large for the shared-memory tiles of this n_fft");
                return AMTFEAT_ERR_INVALID;
            }
            return AMTFEAT_OK;
        case AMTFEAT_VQT:
        case AMTFEAT_HVQT:
            if (c.n_harmonics < 1 || c.n_harmonics > AMTFEAT_MAX_HARMONICS) { set_error("n_harmonics out of range"); return AMTFEAT_ERR_INVALID; }
            p.C = c.n_harmonics;
            p.F = c.n_bins;
            return build_vqt(p);
        default:
            set_error("unknown module kind");
            return AMTFEAT_ERR_INVALID;
    }
}

const char *decimator_name(const Plan &p) {
    if (p.decim_mode == 2 || (p.decim_mode == 1 && p.decim_hh.empty()) || (p.decim_mode == 0 && p.decim_h64.empty())) return "direct";
    return p.decim_mode == 1 ? "fft32" : "fft64";
}

std::string describe(const Plan &p) {
    const amtfeat_config &c = p.cfg;
    std::ostringstream o;
    o << "{\"kind\": " << c.kind << ", \"channels\": " << p.C << ", \"feature_size\": " << p.F << ", \"device\": " << p.device;
    if (c.kind == AMTFEAT_MEL) o << ", \"mel_nnz\": " << p.mel_w.size();
    if (c.kind == AMTFEAT_VQT || c.kind == AMTFEAT_HVQT) {
        o << ", \"n_octaves\": " << p.n_oct << ", \"n_levels\": " << p.n_levels << ", \"decim_taps\": " << p.taps.size() << ", \"decimator\": \"" << decimator_name(p) << "\""
          << ", \"basis_nnz\": " << p.weights.size() << ", \"padded_block_nnz\": " << p.weights4.size() * 2 << ", \"exact_ladders\": [";
        for (size_t a = 0; a < p.alts.size(); ++a) o << (a ? ", " : "") << "{\"eds\": " << p.alts[a].eds << ", \"taps\": " << p.alts[a].taps.size() << "}";
        o << "], \"alt_mask\": " << p.alt_mask << ", \"eds_ref\": [";
        for (size_t h = 0; h < p.harm.size(); ++h) o << (h ? ", " : "") << p.harm[h].eds_ref;
        o << "], \"eds_lib\": [";
        for (size_t h = 0; h < p.harm.size(); ++h) o << (h ? ", " : "") << p.harm[h].eds_lib;
        o << "], \"items\": [";
        for (size_t i = 0; i < p.items.size(); ++i) {
            const CqtItem &it = p.items[i];
            o << (i ? ", " : "") << "{\"level\": " << it.level << ", \"n_fft\": " << it.nfft << ", \"hop\": " << it.hop
              << ", \"rows\": " << it.nrows << ", \"blocks\": " << it.nblk << ", \"unique_rows\": " << it.nuniq << ", \"kmin\": " << it.kmin << ", \"kmax\": " << p.item_kmax_true[i] << ", \"kmax_padded\": " << it.kmax
              << ", \"slide\": " << ((!p.slide_off && is_slide_item(it)) ? 1 : 0) << ", \"alt\": " << it.alt << ", \"block_steps\": [";
            for (int b = 0; b < it.nblk; ++b) o << (b ? ", " : "") << p.blocks[it.blk0 + b].steps;
            o << "]}";
        }
        o << "]";
        if (const char *e = std::getenv("AMTFEAT_DESCRIBE_ROWS")) {
            if (std::string(e) == "1") {     // per basis row: channel, bin, first kept FFT bin, band width, kept entries (tests / debugging)
                o << ", \"rows\": [";
                for (size_t r = 0; r < p.rows.size(); ++r) {
                    const CqtRow &rw = p.rows[r];
                    int nnz = 0;
                    for (int q = 0; q < rw.cnt; ++q) nnz += (p.weights[rw.woff + q].x != 0.f || p.weights[rw.woff + q].y != 0.f);
                    o << (r ? ", " : "") << "[" << rw.chan << ", " << rw.bin << ", " << rw.col0 << ", " << rw.cnt << ", " << nnz << "]";
                }
                o << "]";
            }
        }
    }
    o << "}";
    return o.str();
}

std::string describe_clip(const Plan &p, int64_t n) {
    std::ostringstream o;
    const int64_t T = output_frames(p, n);
    o << "{\"frames\": " << T;
    if (p.cfg.kind == AMTFEAT_VQT || p.cfg.kind == AMTFEAT_HVQT) {
        int64_t T_all = T;
        o << ", \"harmonic_frames\": [";
        for (size_t h = 0; h < p.harm.size(); ++h) {
            const int64_t th = harmonic_frames(p, (int)h, n);
            if (p.cfg.kind == AMTFEAT_HVQT && p.cfg.decibels) T_all = std::max(T_all, th);
            o << (h ? ", " : "") << th;
        }
        int32_t len[kMaxLevels];
        level_lengths(p, n, len);
        o << "], \"frames_computed\": " << T_all << ", \"decim_delay\": " << ((int64_t)p.taps.size() - 1) / 2 << ", \"level_len\": [";
        for (int l = 0; l < p.n_levels; ++l) o << (l ? ", " : "") << len[l];
        TailLayout tl;
        clip_tail_layout(p, n, T_all, tl);
        o << "], \"dev\": [";
        for (int l = 0; l < p.n_levels; ++l) o << (l ? ", " : "") << (tl.dev[l] == INT32_MAX ? -1 : tl.dev[l]);
        o << "], \"hsafe\": [";
        for (int l = 0; l < p.n_levels; ++l) o << (l ? ", " : "") << tl.hsafe[l];
        o << "], \"alt_th\": [";
        for (int l = 0; l < p.n_levels; ++l) o << (l ? ", " : "") << tl.th[l];
        o << "], \"alt_t0\": [";
        for (int l = 0; l < p.n_levels; ++l) o << (l ? ", " : "") << (tl.t0[l] == INT32_MAX ? -1 : tl.t0[l]);
        o << "], \"alt_nfft_max\": [";
        for (int l = 0; l < p.n_levels; ++l) o << (l ? ", " : "") << p.alt_nfft_max[l];
        o << "], \"exact_ladders\": [";
        for (size_t a = 0; a < p.alts.size(); ++a) {
            o << (a ? ", " : "") << "{\"eds\": " << p.alts[a].eds << ", \"first\": [";
            for (int l = 0; l < p.n_levels; ++l) o << (l ? ", " : "") << tl.first[a][l];
            o << "], \"count\": [";
            for (int l = 0; l < p.n_levels; ++l) o << (l ? ", " : "") << tl.count[a][l];
            o << "], \"head\": [";
            for (int l = 0; l < p.n_levels; ++l) o << (l ? ", " : "") << tl.hlen[a][l];
            o << "]}";
        }
        o << "]";
    }
    o << "}";
    return o.str();
}

}  // namespace amtfeat
