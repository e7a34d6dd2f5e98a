// Audio ingest on the device -- the step *before* the hot path:
//   tools.load_normalize_audio (amt_tools/tools/io.py:50-87) = librosa.load(sr=fs, mono=True, res_type='kaiser_best')
//   followed by tools.rms_norm (amt_tools/tools/utils.py:2789-2814).
// File decoding stays on the host; what runs here is everything after it: channel mean (librosa.to_mono), band-limited
// sinc interpolation (librosa.resample -> resampy, whose published algorithm is restated: a Kaiser-windowed sinc table
// with 2^precision samples per zero crossing, linearly interpolated between table entries, walked with stride
// int(scale * 2^precision) on the left and right wing of every output sample), and the RMS normalisation.
//
// resampy is not installed in this image: "parity unpinned" for the resampler (oracle/ingest.py restates the same
// published algorithm in numpy; analytic sinusoid / DC known answers pin the scale chain).  rms_norm and to_mono are
// pinned by the reference's own formulas.
#include <cuda_runtime.h>

#include "device_guard.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "plan.h"

namespace amtfeat {

const char *last_error_cstr();

struct Resampler {
    int device = -1;
    double sr_orig = 0, sr_new = 0, ratio = 1, scale = 1, rolloff = 0;
    int num_zeros = 0, num_table = 0, index_step = 0, nwin = 0;
    std::vector<double> win, delta;        // interp_win (already scaled by the ratio when downsampling), interp_delta
    double2 *d_tab = nullptr;              // (win, delta) pairs
    // Phase-table form (integer sample rates: the fractional position repeats every q outputs, see resample_poly_kernel)
    bool poly = false;
    long long p = 0, q = 0;                // time_increment = sr_orig / sr_new = p / q in lowest terms
    int Q = 0, P = 0;                      // super-period: Q = q m outputs <-> P = p m inputs (m makes Q >= 64)
    int LT = 0, NTAP = 0;                  // taps of the left wing (x[n - LT + 1 .. n]) and in total, per phase, zero padded to the longest
    int PADW = 0, WROW = 0;                // zero taps on either side of a weight row, row stride (PADW + NTAP + PADW)
    int group = 0;                         // phases a warp accumulates together (1, 2 or 4); PB = 8 * group
    int PB = 0, pitch = 0;                 // phases per pass of a CTA, row pitch (floats, odd) of its input tile
    std::vector<double> W;                 // [q][WROW] float64 weights, win[offset + i step] + eta delta[offset + i step] per tap
    double *d_W = nullptr;
};

static double bessel_i0_(double x) {
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 500; ++k) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-18 * sum) break;
    }
    return sum;
}

// resampy.filters.sinc_window(num_zeros, precision, window=kaiser(beta), rolloff): right half of the symmetric
// interpolation filter, n = 2^precision * num_zeros, n + 1 entries.
static void sinc_window(int num_zeros, int precision, double beta, double rolloff, std::vector<double> &out) {
    const double kPi = 3.14159265358979323846264338327950288;
    const int num_bits = 1 << precision, n = num_bits * num_zeros;
    out.resize((size_t)n + 1);
    const double i0b = bessel_i0_(beta);
    for (int i = 0; i <= n; ++i) {
        const double x = rolloff * ((double)num_zeros * (double)i / (double)n);   // np.linspace(0, num_zeros, n + 1)
        const double sinc = x == 0.0 ? 1.0 : std::sin(kPi * x) / (kPi * x);
        const double r = (double)i / (double)n;                                    // scipy.signal.kaiser(2n + 1, beta)[n:]
        const double taper = bessel_i0_(beta * std::sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
        out[i] = taper * rolloff * sinc;
    }
}

constexpr int kPolyLanes = 32, kPolyWarps = 8;                    // super-periods per CTA (lanes), warps
constexpr size_t kPolySmemMax = 110 * 1024, kPolyTableMax = 32u << 20;   // two CTAs per SM

// Phase table of the polyphase form.  With integer sample rates time_register = t p / q, so output t uses
//   n = floor(t p / q),  frac = (t p mod q) / q:  q distinct phases, each a fixed FIR over x[n - LT + 1 .. n + RT]
// whose taps are exactly the weights the per-sample walk forms (left wing: offset = int(scale frac 2^prec), eta its fraction,
// taps win[offset + i step] + eta delta[offset + i step], i < (nwin - offset) / step; right wing with scale - scale frac).
// The walk's own truncation at the signal's ends (i <= n, n + 1 + k < n_in) equals zero extension of x.
static void resampler_build_phases(Resampler &r) {
    r.poly = false;
    if (const char *e = std::getenv("AMTFEAT_RESAMPLE")) if (std::string(e) == "direct") return;
    const double a = std::round(r.sr_orig), b = std::round(r.sr_new);
    if (std::fabs(a - r.sr_orig) > 1e-9 || std::fabs(b - r.sr_new) > 1e-9 || a < 1 || b < 1 || a > 1e9 || b > 1e9) return;
    long long x = (long long)a, y = (long long)b;
    while (y) { const long long t = x % y; x = y; y = t; }
    const long long p = (long long)a / x, q = (long long)b / x;
    if (q > 2048) return;
    const long long m = q >= 64 ? 1 : (64 + q - 1) / q;
    if (p * m > (1 << 20)) return;
    std::vector<std::vector<double>> wl((size_t)q), wr((size_t)q);
    int LT = 0, RT = 0;
    for (long long ph = 0; ph < q; ++ph) {
        const double fracpos = (double)((ph * p) % q) / (double)q;
        for (int wing = 0; wing < 2; ++wing) {
            const double frac = wing == 0 ? r.scale * fracpos : r.scale - r.scale * fracpos;
            const double index_frac = frac * r.num_table;
            const int offset = (int)index_frac;
            const double eta = index_frac - offset;
            const int cnt = (r.nwin - offset) / r.index_step;
            std::vector<double> &w = wing == 0 ? wl[(size_t)ph] : wr[(size_t)ph];
            w.resize((size_t)std::max(cnt, 0));
            for (int i = 0; i < cnt; ++i) w[(size_t)i] = r.win[(size_t)offset + (size_t)i * r.index_step] + eta * r.delta[(size_t)offset + (size_t)i * r.index_step];
            (wing == 0 ? LT : RT) = std::max(wing == 0 ? LT : RT, cnt);
        }
    }
    if (LT < 1) return;
    const int NTAP = LT + RT;
    // A CTA walks the phases in passes of PB = 8 warps x `group`: per pass it stages the input tile (32 super-periods, one row
    // each) and the PB weight rows in shared memory.  A warp walks the input columns of its `group` consecutive phases together;
    // phase u lags the first by n(a + u) - n(a) columns, which PADW zero taps on either side of every row absorb.
    int group = 0, PADW = 0, WROW = 0, pitch = 0;
    for (int cand = 4; cand >= 1; cand >>= 1) {
        const int padw = (int)(((long long)(cand - 1) * p + q - 1) / q) + 1;
        const int wrow = (padw + NTAP + padw + 1) & ~1;
        const long long span = ((long long)kPolyWarps * cand * p + q - 1) / q + NTAP + 2;
        const long long pt = span | 1;                                        // odd pitch: the lanes of a warp read one column of 32 rows
        if ((size_t)kPolyLanes * pt * sizeof(float) + (size_t)kPolyWarps * cand * wrow * sizeof(double) <= kPolySmemMax) {
            group = cand; PADW = padw; WROW = wrow; pitch = (int)pt;
            break;
        }
    }
    if (!group || (size_t)q * WROW * sizeof(double) > kPolyTableMax) return;
    r.W.assign((size_t)q * WROW, 0.0);
    for (long long ph = 0; ph < q; ++ph) {
        double *row = r.W.data() + (size_t)ph * WROW + PADW;
        for (int i = 0; i < (int)wl[(size_t)ph].size(); ++i) row[LT - 1 - i] = wl[(size_t)ph][(size_t)i];      // tap d reads x[n - (LT - 1) + d]
        for (int k = 0; k < (int)wr[(size_t)ph].size(); ++k) row[LT + k] = wr[(size_t)ph][(size_t)k];
    }
    const int PB = kPolyWarps * group;
    r.group = group;
    r.p = p; r.q = q; r.Q = (int)(q * m); r.P = (int)(p * m); r.LT = LT; r.NTAP = NTAP; r.PADW = PADW; r.WROW = WROW; r.PB = PB; r.pitch = pitch;
    r.poly = true;
}

int resampler_build(Resampler &r, double sr_orig, double sr_new, int filter) {
    if (!(sr_orig > 0) || !(sr_new > 0)) { set_error("sample rates must be positive"); return AMTFEAT_ERR_INVALID; }
    int precision = 9;
    double beta;
    if (filter == AMTFEAT_RES_KAISER_BEST) { r.num_zeros = 64; beta = 14.769656459379492; r.rolloff = 0.9475937167399596; }
    else if (filter == AMTFEAT_RES_KAISER_FAST) { r.num_zeros = 16; beta = 8.555504641634386; r.rolloff = 0.85; }
    else { set_error("unknown resampling filter"); return AMTFEAT_ERR_INVALID; }
    r.sr_orig = sr_orig; r.sr_new = sr_new;
    r.ratio = sr_new / sr_orig;
    r.scale = std::min(1.0, r.ratio);
    r.num_table = 1 << precision;
    sinc_window(r.num_zeros, precision, beta, r.rolloff, r.win);
    r.nwin = (int)r.win.size();
    if (r.ratio < 1.0) for (double &w : r.win) w *= r.ratio;
    r.delta.resize(r.win.size());
    for (size_t i = 0; i + 1 < r.win.size(); ++i) r.delta[i] = r.win[i + 1] - r.win[i];
    r.delta.back() = 0.0;                                                          // np.diff(..., append=interp_win[-1])
    r.index_step = (int)(r.scale * r.num_table);
    if (r.index_step < 1) { set_error("sample-rate ratio too small for the interpolation table"); return AMTFEAT_ERR_INVALID; }
    resampler_build_phases(r);
    return AMTFEAT_OK;
}

// librosa.resample(..., fix=True): resampy allocates int(n * ratio) samples and librosa.util.fix_length pads the result with
// zeros to ceil(n * ratio) -- one extra (zero) sample whenever n * ratio is not an integer
int64_t resampler_out_len(const Resampler &r, int64_t n) { return (int64_t)std::ceil((double)n * r.ratio); }

#define AMT_CUDA(call)                                                                                 \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                             \
            return AMTFEAT_ERR_CUDA;                                                                   \
        }                                                                                              \
    } while (0)

int resampler_upload(Resampler &r, int device) {
    r.device = device;
    if (device < 0) return AMTFEAT_OK;
    DeviceGuard guard(device);
    AMT_CUDA(guard.status);
    std::vector<double2> tab(r.win.size());
    for (size_t i = 0; i < tab.size(); ++i) tab[i] = make_double2(r.win[i], r.delta[i]);
    AMT_CUDA(cudaMalloc(reinterpret_cast<void **>(&r.d_tab), tab.size() * sizeof(double2)));
    AMT_CUDA(cudaMemcpy(r.d_tab, tab.data(), tab.size() * sizeof(double2), cudaMemcpyHostToDevice));
    if (r.poly) {
        AMT_CUDA(cudaMalloc(reinterpret_cast<void **>(&r.d_W), r.W.size() * sizeof(double)));
        AMT_CUDA(cudaMemcpy(r.d_W, r.W.data(), r.W.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    return AMTFEAT_OK;
}

void resampler_free(Resampler &r) {
    if (r.d_tab || r.d_W) {
        DeviceGuard guard(r.device);
        if (r.d_tab) cudaFree(r.d_tab);
        if (r.d_W) cudaFree(r.d_W);
        r.d_tab = nullptr;
        r.d_W = nullptr;
    }
}

struct IngestClip {
    long long in_off, n_in, out_off, n_out;
};

// One thread per output sample.  Weights and the accumulation are float64 (the reference forms float64 weights and adds
// float64 products); the wings read neighbouring input samples, which L1 / L2 serve.
__global__ void __launch_bounds__(256) resample_kernel(const float *__restrict__ in, float *__restrict__ out,
                                                       const IngestClip *__restrict__ clips, const double2 *__restrict__ tab,
                                                       int nwin, int num_table, int index_step, double scale, double time_increment,
                                                       double ratio) {
    const IngestClip c = clips[blockIdx.y];
    const float *x = in + c.in_off;
    const long long n_res = (long long)((double)c.n_in * ratio);      // samples resampy itself produces; the rest is fix_length's zero padding
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < c.n_out; t += (long long)gridDim.x * blockDim.x) {
        if (t >= n_res) {
            out[c.out_off + t] = 0.f;
            continue;
        }
        const double time_register = (double)t * time_increment;
        const long long n = (long long)time_register;
        double acc = 0.0;
        {   // left wing: x[n], x[n - 1], ...
            const double frac = scale * (time_register - (double)n);
            const double index_frac = frac * num_table;
            const int offset = (int)index_frac;
            const double eta = index_frac - offset;
            const long long i_max = min(n + 1, (long long)((nwin - offset) / index_step));
            for (long long i = 0; i < i_max; ++i) {
                const double2 w = __ldg(tab + offset + i * index_step);
                acc = fma(w.x + eta * w.y, (double)__ldg(x + n - i), acc);
            }
        }
        {   // right wing: x[n + 1], x[n + 2], ...
            const double frac = scale - scale * (time_register - (double)n);
            const double index_frac = frac * num_table;
            const int offset = (int)index_frac;
            const double eta = index_frac - offset;
            const long long k_max = min(c.n_in - n - 1, (long long)((nwin - offset) / index_step));
            for (long long k = 0; k < k_max; ++k) {
                const double2 w = __ldg(tab + offset + k * index_step);
                acc = fma(w.x + eta * w.y, (double)__ldg(x + n + k + 1), acc);
            }
        }
        out[c.out_off + t] = (float)acc;
    }
}

// Polyphase form of the same resampler (integer sample rates; resampler_build_phases).
//   A CTA takes 32 consecutive super-periods of one clip -- lane g owns super-period gb + g, i.e. the outputs
//   t = (gb + g) Q + a, a < Q -- and walks the phases a in passes of PB.  Per pass the inputs the 32 x PB outputs read are staged
//   once in shared memory, one row per super-period with an ODD pitch: the lanes of a warp (same phase, hence the same,
//   warp-uniform weights) read their rows at the same column, conflict free.  A warp accumulates kPolyGroup phases at once
//   in ONE walk over their common input columns (every sample read and converted once, feeding all phases; independent
//   float64 chains); weights are fetched as aligned pairs of taps from the phase's contiguous row.
//   Every input sample is loaded from global memory once per pass, every weight row once per CTA.
struct PolyParams {
    const float *in;
    float *out;
    const IngestClip *clips;
    const double *W;
    long long p, q;
    int Q, P, LT, NTAP, PADW, WROW, pitch;
    double ratio;
};

template <int GROUP>
__global__ void __launch_bounds__(kPolyWarps * 32) resample_poly_kernel(const PolyParams pp) {
    constexpr int PB = kPolyWarps * GROUP;
    extern __shared__ __align__(16) float tile[];
    double *wsm = reinterpret_cast<double *>(tile + (size_t)kPolyLanes * pp.pitch + ((kPolyLanes * pp.pitch) & 1));   // [PB][WROW]
    const IngestClip c = pp.clips[blockIdx.y];
    const long long gb = (long long)blockIdx.x * kPolyLanes;
    if (gb * pp.Q >= c.n_out) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *x = pp.in + c.in_off;
    float *y = pp.out + c.out_off;
    const long long n_res = (long long)((double)c.n_in * pp.ratio);       // samples resampy itself produces; the rest is fix_length's zero padding
    const long long tbase = (gb + lane) * pp.Q;
    for (int a0 = 0; a0 < pp.Q; a0 += PB) {
        const int a1 = min(pp.Q, a0 + PB);
        const long long n_lo = ((long long)a0 * pp.p) / pp.q, n_hi = ((long long)(a1 - 1) * pp.p) / pp.q;
        const int lrow = (int)(n_hi - n_lo) + pp.NTAP + 1;
        __syncthreads();                                                   // the previous pass is done with the tile and the weights
        for (int r = warp; r < kPolyLanes; r += kPolyWarps) {
            const long long j0 = (gb + r) * pp.P + n_lo - (pp.LT - 1);
            float *row = tile + (size_t)r * pp.pitch;
            for (int k = lane; k < lrow; k += 32) {
                const long long j = j0 + k;
                row[k] = (j >= 0 && j < c.n_in) ? __ldg(x + j) : 0.f;
            }
        }
        for (int r = warp; r < PB; r += kPolyWarps) {                      // phases past the end repeat the last one (results dropped)
            const double *src = pp.W + (size_t)(min(a0 + r, pp.Q - 1) % pp.q) * pp.WROW;
            double *dst = wsm + (size_t)r * pp.WROW;
            for (int k = lane; k < pp.WROW; k += 32) dst[k] = __ldg(src + k);
        }
        __syncthreads();
        // one walk over the input columns of the warp's GROUP phases: every sample is read and converted ONCE and feeds all of
        // them (phase u multiplies column col by its tap col - shift_u; the zero taps around a row make that unconditional);
        // the weights are warp-uniform shared-memory reads (broadcast), the samples conflict free (odd pitch)
        const int abase = a0 + warp * GROUP;
        if (abase < a1) {
            const double *wrow[GROUP];
            double acc[GROUP];
            const long long n_first = ((long long)abase * pp.p) / pp.q;
            int shift_max = 0;
#pragma unroll
            for (int u = 0; u < GROUP; ++u) {
                const int a = min(abase + u, a1 - 1);
                const int shift = (int)(((long long)a * pp.p) / pp.q - n_first);
                wrow[u] = wsm + (size_t)(warp * GROUP + min(u, a1 - 1 - abase)) * pp.WROW + pp.PADW - shift;
                shift_max = max(shift_max, shift);
                acc[u] = 0.0;
            }
            const float *xr = tile + (size_t)lane * pp.pitch + (int)(n_first - n_lo);
            const int ncol = pp.NTAP + shift_max;
#pragma unroll 4
            for (int col = 0; col < ncol; ++col) {
                const double xv = (double)xr[col];
#pragma unroll
                for (int u = 0; u < GROUP; ++u) acc[u] = fma(wrow[u][col], xv, acc[u]);
            }
#pragma unroll
            for (int u = 0; u < GROUP; ++u) {
                const int a = abase + u;
                const long long t = tbase + a;
                if (a < a1 && t < c.n_out) y[t] = t < n_res ? (float)acc[u] : 0.f;
            }
        }
    }
}

// librosa.to_mono: mean over the channel axis of a (channels, n) clip.
__global__ void __launch_bounds__(256) to_mono_kernel(const float *__restrict__ in, float *__restrict__ out, long long n, int channels) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int c = 0; c < channels; ++c) s += __ldg(in + (long long)c * n + i);
        out[i] = s / (float)channels;
    }
}

// tools.rms_norm: audio / sqrt(mean(audio ** 2)) unless the RMS is zero.  Pass 1 accumulates the sum of squares per clip in
// float64 (one atomic per CTA); pass 2 divides by the float32 RMS, as the reference's float32 arithmetic does.
__global__ void __launch_bounds__(256) sumsq_kernel(const float *__restrict__ x, const IngestClip *__restrict__ clips, double *__restrict__ acc) {
    const IngestClip c = clips[blockIdx.y];
    const float *p = x + c.in_off;
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n_in; i += (long long)gridDim.x * blockDim.x) {
        const float v = __ldg(p + i);
        s += (double)(v * v);   // float32 square, as audio ** 2 on a float32 array
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __shared__ double part[8];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += part[w];
        if (tot != 0.0) atomicAdd(acc + blockIdx.y, tot);
    }
}

__global__ void __launch_bounds__(256) rms_scale_kernel(float *__restrict__ x, const IngestClip *__restrict__ clips, const double *__restrict__ acc) {
    const IngestClip c = clips[blockIdx.y];
    if (c.n_in == 0) return;
    const float rms = sqrtf((float)(acc[blockIdx.y] / (double)c.n_in));
    if (!(rms > 0.f)) return;
    float *p = x + c.in_off;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < c.n_in; i += (long long)gridDim.x * blockDim.x) p[i] = p[i] / rms;
}

static unsigned grid_for(long long n) { return (unsigned)std::max<long long>(1, std::min<long long>(148 * 16, (n + 255) / 256)); }

size_t ingest_workspace_bytes(int batch) { return (size_t)batch * (sizeof(IngestClip) + sizeof(double)) + 256; }

static int stage_clips(const int64_t *in_off, const int64_t *n_in, const int64_t *out_off, const int64_t *n_out, int batch, void *d_ws,
                       size_t ws_bytes, cudaStream_t st, IngestClip **d_clips, double **d_acc) {
    if (ws_bytes < ingest_workspace_bytes(batch)) { set_error("workspace too small"); return AMTFEAT_ERR_WORKSPACE; }
    std::vector<IngestClip> h(batch);
    for (int b = 0; b < batch; ++b) h[b] = IngestClip{in_off[b], n_in[b], out_off ? out_off[b] : 0, n_out ? n_out[b] : 0};
    char *ws = static_cast<char *>(d_ws);
    *d_acc = reinterpret_cast<double *>(ws);                                  // 8-byte aligned: the workspace base is
    *d_clips = reinterpret_cast<IngestClip *>(ws + (((size_t)batch * sizeof(double) + 255) / 256) * 256);
    AMT_CUDA(cudaMemcpyAsync(*d_clips, h.data(), h.size() * sizeof(IngestClip), cudaMemcpyHostToDevice, st));
    AMT_CUDA(cudaStreamSynchronize(st));   // `h` is pageable stack-owned memory: do not let it die under the copy
    return AMTFEAT_OK;
}

int resample_run(const Resampler &r, const float *d_in, const int64_t *in_off, const int64_t *n_in, int batch, float *d_out,
                 const int64_t *out_off, void *d_ws, size_t ws_bytes, void *stream) {
    if (r.device < 0) { set_error("host-only resampler: no CUDA device (there is no CPU compute path)"); return AMTFEAT_ERR_NO_DEVICE; }
    if (batch <= 0) return AMTFEAT_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    std::vector<int64_t> n_out(batch);
    long long maxo = 0;
    for (int b = 0; b < batch; ++b) { n_out[b] = resampler_out_len(r, n_in[b]); maxo = std::max<long long>(maxo, n_out[b]); }
    if (maxo == 0) return AMTFEAT_OK;
    IngestClip *d_clips; double *d_acc;
    int rc = stage_clips(in_off, n_in, out_off, n_out.data(), batch, d_ws, ws_bytes, st, &d_clips, &d_acc);
    if (rc) return rc;
    if (r.poly) {
        PolyParams pp{d_in, d_out, d_clips, r.d_W, r.p, r.q, r.Q, r.P, r.LT, r.NTAP, r.PADW, r.WROW, r.pitch, r.ratio};
        const size_t tile_floats = (size_t)kPolyLanes * r.pitch + (((size_t)kPolyLanes * r.pitch) & 1);
        const size_t smem = tile_floats * sizeof(float) + (size_t)kPolyWarps * r.group * r.WROW * sizeof(double);
        static bool attr_set[64] = {};
        if (r.device >= 0 && r.device < 64 && !attr_set[r.device]) {
            AMT_CUDA(cudaFuncSetAttribute(resample_poly_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPolySmemMax + 1024));
            AMT_CUDA(cudaFuncSetAttribute(resample_poly_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPolySmemMax + 1024));
            AMT_CUDA(cudaFuncSetAttribute(resample_poly_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPolySmemMax + 1024));
            attr_set[r.device] = true;
        }
        const long long supers = (maxo + r.Q - 1) / r.Q;
        dim3 grid((unsigned)((supers + kPolyLanes - 1) / kPolyLanes), batch);
        if (r.group == 4) resample_poly_kernel<4><<<grid, kPolyWarps * 32, smem, st>>>(pp);
        else if (r.group == 2) resample_poly_kernel<2><<<grid, kPolyWarps * 32, smem, st>>>(pp);
        else resample_poly_kernel<1><<<grid, kPolyWarps * 32, smem, st>>>(pp);
        AMT_CUDA(cudaGetLastError());
        return AMTFEAT_OK;
    }
    dim3 grid(grid_for(maxo), batch);
    resample_kernel<<<grid, 256, 0, st>>>(d_in, d_out, d_clips, r.d_tab, r.nwin, r.num_table, r.index_step, r.scale, 1.0 / r.ratio, r.ratio);
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

// 16-bit PCM (what a WAV file holds) -> float32 on the device: x * scale (soundfile / librosa.load: scale = 1 / 32768; a caller
// that normalises by a known RMS folds it in).  Halves the host-to-device bytes of a device-resident consumer.
__global__ void __launch_bounds__(256) pcm16_kernel(const short *__restrict__ in, float *__restrict__ out, long long n, float scale) {
    const long long n4 = n >> 2;
    const short4 *in4 = reinterpret_cast<const short4 *>(in);
    float4 *out4 = reinterpret_cast<float4 *>(out);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const short4 v = __ldg(in4 + i);
        out4[i] = make_float4(v.x * scale, v.y * scale, v.z * scale, v.w * scale);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) out[(n4 << 2) + threadIdx.x] = in[(n4 << 2) + threadIdx.x] * scale;
}

int pcm16_run(const short *d_in, int64_t n, float scale, float *d_out, void *stream) {
    if (n <= 0) return AMTFEAT_OK;
    if ((reinterpret_cast<uintptr_t>(d_in) & 7) || (reinterpret_cast<uintptr_t>(d_out) & 15)) {
        set_error("pcm16 input must be 8-byte aligned and the float32 output 16-byte aligned");
        return AMTFEAT_ERR_INVALID;
    }
    pcm16_kernel<<<grid_for(n / 4 + 1), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_in, d_out, n, scale);
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

int to_mono_run(const float *d_in, int64_t n, int channels, float *d_out, void *stream) {
    if (channels < 1) { set_error("channels must be >= 1"); return AMTFEAT_ERR_INVALID; }
    if (n <= 0) return AMTFEAT_OK;
    to_mono_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_in, d_out, n, channels);
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

int rms_norm_run(float *d_audio, const int64_t *off, const int64_t *n, int batch, void *d_ws, size_t ws_bytes, void *stream) {
    if (batch <= 0) return AMTFEAT_OK;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long maxn = 0;
    for (int b = 0; b < batch; ++b) maxn = std::max<long long>(maxn, n[b]);
    if (maxn == 0) return AMTFEAT_OK;
    IngestClip *d_clips; double *d_acc;
    int rc = stage_clips(off, n, nullptr, nullptr, batch, d_ws, ws_bytes, st, &d_clips, &d_acc);
    if (rc) return rc;
    AMT_CUDA(cudaMemsetAsync(d_acc, 0, (size_t)batch * sizeof(double), st));
    dim3 grid(grid_for(maxn), batch);
    sumsq_kernel<<<grid, 256, 0, st>>>(d_audio, d_clips, d_acc);
    rms_scale_kernel<<<grid, 256, 0, st>>>(d_audio, d_clips, d_acc);
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

}  // namespace amtfeat

struct amtfeat_resampler {
    amtfeat::Resampler r;
};

extern "C" {

int amtfeat_resampler_create(double sr_orig, double sr_new, int filter, int device, amtfeat_resampler **out) {
    if (!out) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    *out = nullptr;
    amtfeat_resampler *h = new (std::nothrow) amtfeat_resampler();
    if (!h) { amtfeat::set_error("out of memory"); return AMTFEAT_ERR_INVALID; }
    int rc;
    try {
        rc = amtfeat::resampler_build(h->r, sr_orig, sr_new, filter);
        if (rc == AMTFEAT_OK) rc = amtfeat::resampler_upload(h->r, device);
    } catch (const std::exception &e) {
        amtfeat::set_error(std::string("exception: ") + e.what());
        rc = AMTFEAT_ERR_INVALID;
    }
    if (rc != AMTFEAT_OK) {
        amtfeat::resampler_free(h->r);
        delete h;
        return rc;
    }
    *out = h;
    return AMTFEAT_OK;
}

void amtfeat_resampler_destroy(amtfeat_resampler *r) {
    if (!r) return;
    amtfeat::resampler_free(r->r);
    delete r;
}

int64_t amtfeat_resampler_out_len(const amtfeat_resampler *r, int64_t num_samples) {
    if (!r || num_samples < 0) return -1;
    return amtfeat::resampler_out_len(r->r, num_samples);
}

int64_t amtfeat_resampler_table(const amtfeat_resampler *r, double *win, int64_t capacity, int *num_table, int *index_step) {
    if (!r) return -1;
    if (num_table) *num_table = r->r.num_table;
    if (index_step) *index_step = r->r.index_step;
    const int64_t n = (int64_t)r->r.win.size();
    if (win) std::memcpy(win, r->r.win.data(), (size_t)std::min<int64_t>(n, std::max<int64_t>(0, capacity)) * sizeof(double));
    return n;
}

size_t amtfeat_ingest_workspace_bytes(int batch) { return amtfeat::ingest_workspace_bytes(batch < 0 ? 0 : batch); }

int amtfeat_resample(const amtfeat_resampler *r, const float *d_in, const int64_t *in_offsets, const int64_t *num_samples, int batch,
                     float *d_out, const int64_t *out_offsets, void *d_ws, size_t ws_bytes, void *stream) {
    if (!r || (batch > 0 && (!d_in || !in_offsets || !num_samples || !d_out || !out_offsets || !d_ws))) {
        amtfeat::set_error("null argument");
        return AMTFEAT_ERR_INVALID;
    }
    for (int b = 0; b < batch; ++b)
        if (num_samples[b] < 0) { amtfeat::set_error("negative clip length"); return AMTFEAT_ERR_INVALID; }
    try {
        return amtfeat::resample_run(r->r, d_in, in_offsets, num_samples, batch, d_out, out_offsets, d_ws, ws_bytes, stream);
    } catch (const std::exception &e) {
        amtfeat::set_error(std::string("exception: ") + e.what());
        return AMTFEAT_ERR_INVALID;
    }
}

int amtfeat_pcm16_to_float(const int16_t *d_pcm, int64_t num_samples, float scale, float *d_out, void *stream) {
    if (num_samples > 0 && (!d_pcm || !d_out)) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    return amtfeat::pcm16_run(reinterpret_cast<const short *>(d_pcm), num_samples, scale, d_out, stream);
}

int amtfeat_to_mono(const float *d_in, int64_t num_samples, int channels, float *d_out, void *stream) {
    if (num_samples > 0 && (!d_in || !d_out)) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    return amtfeat::to_mono_run(d_in, num_samples, channels, d_out, stream);
}

int amtfeat_rms_norm(float *d_audio, const int64_t *offsets, const int64_t *num_samples, int batch, void *d_ws, size_t ws_bytes,
                     void *stream) {
    if (batch > 0 && (!d_audio || !offsets || !num_samples || !d_ws)) { amtfeat::set_error("null argument"); return AMTFEAT_ERR_INVALID; }
    for (int b = 0; b < batch; ++b)
        if (num_samples[b] < 0) { amtfeat::set_error("negative clip length"); return AMTFEAT_ERR_INVALID; }
    try {
        return amtfeat::rms_norm_run(d_audio, offsets, num_samples, batch, d_ws, ws_bytes, stream);
    } catch (const std::exception &e) {
        amtfeat::set_error(std::string("exception: ") + e.what());
        return AMTFEAT_ERR_INVALID;
    }
}

}  // extern "C"
