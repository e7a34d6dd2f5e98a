// sm_100a kernels of libamtfeat and their launchers.
//
//   K1/K2/K3  stft_kernel<NC, MODE>   frame + window + real FFT + |.|^2 (+ mel projection) + per-clip max
//   K4        decimate_kernel         2:1 polyphase FIR decimator (soxr-HQ-class taps, x sqrt(2))
//   K1/K5     cqt_kernel<NC>          rectangular-window FFT of ladder frames + sparse complex basis
//                                     projection, shared by every harmonic that uses the same (level, n_fft)
//   K6        db_epilogue_kernel      (x - ref_dB) floor -80, /80 + 1
//   K7        power_kernel            SignalPower;  frames_kernel: WaveformWrapper framing
//
// All FFT kernels: 8 warps per CTA, one warp per 1024-complex-point unit (fft_device.cuh), audio tile
// staged once in shared memory, outputs transposed through shared memory so global stores are
// T-contiguous.  FP32 SIMT throughout (no tensor cores: the projections are sparse).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "device_guard.h"
#include "fft_device.cuh"
#include "plan.h"

// A/B knobs (tools/variants.py); the defaults are the shipped configuration
#ifndef AMT_CQT_TPC_MAX
#define AMT_CQT_TPC_MAX 8      // most tiles one CQT CTA walks
#endif
#ifndef AMT_CQT_WAVES
#define AMT_CQT_WAVES 4        // ... while keeping at least this many waves of CTAs
#endif
#ifndef AMT_LADDER_OVERLAP
#define AMT_LADDER_OVERLAP 1   // decimation ladder on a side stream underneath the projection of the shallower levels
#endif
#ifndef AMT_MID_LEVEL
#define AMT_MID_LEVEL 3        // projection launches are cut by ladder depth: level 0 | 1 .. AMT_MID_LEVEL | deeper
#endif
#ifndef AMT_SLIDE_RESEED
#define AMT_SLIDE_RESEED 8     // hop <= 4: frames between exact re-seeds of the float32 sliding-DFT phase P (recurrence P *= W^(k hop) in between)
#endif                         // (hop 8 / 16: the recurrence runs in float64 instead, see slide_run)
#ifndef AMT_SLIDE_SPLIT
#define AMT_SLIDE_SPLIT 1      // 0: one launch for all sliding items; 1: bands of at most 128 bins | wider; 2: one launch per CTA size
#endif
#ifndef AMT_SLIDE_TILE
#define AMT_SLIDE_TILE 4096    // samples of the level signal a sliding-DFT tile advances over (at most 1024 frames)
#endif
#ifndef AMT_CQT_MINB
#define AMT_CQT_MINB 2         // resident CTAs per SM the FFT-per-frame kernels are compiled for (3 caps them at 80 registers: measured slower, profiles/)
#endif
#ifndef AMT_STFT_MINB
#define AMT_STFT_MINB 2
#endif
#ifndef AMT_FFT_MAXREG
#define AMT_FFT_MAXREG 0       // > 0: register cap of the FFT kernels (instead of the 128 that two CTAs of 256 threads per SM allow): 112 leaves
#endif                         // room for one CTA of an HBM-bound kernel (dB epilogue, 32 registers) beside two FFT CTAs on an SM
#if AMT_FFT_MAXREG > 0
#define AMT_FFT_BOUNDS(minb) __maxnreg__(AMT_FFT_MAXREG)
#else
#define AMT_FFT_BOUNDS(minb) __launch_bounds__(kThreads, minb)
#endif
#ifndef AMT_EPI_MINB
#define AMT_EPI_MINB 8         // resident CTAs per SM the dB epilogue is compiled for (8: 32 registers, 64 warps per SM)
#endif
#ifndef AMT_EPI_UNROLL
#define AMT_EPI_UNROLL 2       // independent 16-byte loads in flight per thread of the dB epilogue
#endif
#ifndef AMT_EPI_CS
#define AMT_EPI_CS 0           // 1: streaming (evict-first) loads and stores in the dB epilogue
#endif
#ifndef AMT_TAIL_STREAM
#define AMT_TAIL_STREAM 1      // exact-ladder pieces on their own side stream (0: behind the sliding-DFT launches on the ladder's stream)
#endif
#ifndef AMT_DBG_SKIP
#define AMT_DBG_SKIP 0         // timing experiments only (results are wrong): 1 = no projection, 2 = no FFT, 3 = neither
#endif

namespace amtfeat {

#define AMT_CUDA(call)                                                                                 \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                             \
            return AMTFEAT_ERR_CUDA;                                                                   \
        }                                                                                              \
    } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ------------------------------------------------------------------------------------------------
// shared helpers (device)
// ------------------------------------------------------------------------------------------------

// Stage the samples needed by TT consecutive frames of one clip into shared memory, zero-filling
// everything outside [0, n) (this realises centre padding / frame padding without a padded copy).
// Overlapping frames (hop <= nfft): one contiguous span, float4 loads; frame f starts at shift + f * hop.
// Disjoint frames (hop > nfft): TT separate segments, frame f starts at f * nfft (shift = 0).
__device__ __forceinline__ void load_tile(float *tile, const float *__restrict__ src, long long n, long long s0, int hop,
                                          int nfft, int TT, int tid, int &shift, int &fstride) {
    if (hop <= nfft) {
        fstride = hop;
        shift = (int)(((s0 % 4) + 4) % 4);
        const long long a0 = s0 - shift;
        const int span = (TT - 1) * hop + nfft + shift;
        const int nvec = (span + 3) >> 2;
        float4 *tile4 = reinterpret_cast<float4 *>(tile);
        for (int i = tid; i < nvec; i += kThreads) {
            const long long g = a0 + 4ll * i;
            float4 v;
            if (g >= 0 && g + 3 < n) {
                v = __ldg(reinterpret_cast<const float4 *>(src + g));
            } else {
                v.x = (g >= 0 && g < n) ? __ldg(src + g) : 0.f;
                v.y = (g + 1 >= 0 && g + 1 < n) ? __ldg(src + g + 1) : 0.f;
                v.z = (g + 2 >= 0 && g + 2 < n) ? __ldg(src + g + 2) : 0.f;
                v.w = (g + 3 >= 0 && g + 3 < n) ? __ldg(src + g + 3) : 0.f;
            }
            tile4[i] = v;
        }
    } else {
        fstride = nfft;
        shift = 0;
        const int total = TT * nfft;
        for (int e = tid; e < total; e += kThreads) {
            const int f = e / nfft, r = e - f * nfft;
            const long long g = s0 + (long long)f * hop + r;
            tile[e] = (g >= 0 && g < n) ? __ldg(src + g) : 0.f;
        }
    }
}

// Asynchronous form of load_tile for overlapping frames (hop <= nfft): 16-byte cp.async with zero fill outside [0, n)
// (the tile start is 4-element aligned relative to the clip, so a vector is either wholly before the clip or has a
// valid prefix).  The caller commits / waits the group.
__device__ __forceinline__ void load_tile_async(float *tile, const float *__restrict__ src, long long n, long long s0, int hop,
                                                int nfft, int TT, int tid, int &shift) {
    shift = (int)(((s0 % 4) + 4) % 4);
    const long long a0 = s0 - shift;
    const int span = (TT - 1) * hop + nfft + shift;
    const int nvec = (span + 3) >> 2;
    for (int i = tid; i < nvec; i += kThreads) {
        const long long g = a0 + 4ll * i;
        int valid = 0;
        if (g >= 0 && g < n) valid = (int)(n - g < 4 ? n - g : 4) * 4;
        const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(tile + 4 * i));
        const float *gp = src + (valid ? g : 0);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gp), "r"(valid) : "memory");
    }
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------
// K1 + K2 + K3 : STFT / MelSpec
// ------------------------------------------------------------------------------------------------

struct StftParams {
    const float *audio;
    float *out;
    const ClipMeta *meta;
    float *maxbuf;
    const float *window;
    const float2 *tw1, *tw2;
    const int *mel_seg_start, *mel_gsteps, *mel_goff;
    const float2 *mel_ww;
    int hop, pad, n_mels, decibels, tile_floats, scr_floats, tiles_per_cta;
};

constexpr int MODE_STFT = 0, MODE_MEL = 1;

// Input point n (samples 2n, 2n + 1) of the frame that starts at sample s0 of a clip of `len` samples, zero outside the
// clip (this realises centre / frame padding without a padded copy).  Coalesced straight from global memory (the 4x
// frame overlap is served by L1 / L2); `interior` (warp-uniform) skips the bounds checks.
__device__ __forceinline__ float2 load_pair(const float *__restrict__ src, long long len, long long s0, int n, bool interior,
                                            bool vec_ok) {
    const long long i = s0 + 2 * n;
    if (interior && vec_ok) return __ldg(reinterpret_cast<const float2 *>(src + i));
    float2 v;
    v.x = (i >= 0 && i < len) ? __ldg(src + i) : 0.f;
    v.y = (i + 1 >= 0 && i + 1 < len) ? __ldg(src + i + 1) : 0.f;
    return v;
}

// One work item of the mel projection (segment form, see the kernel): the lanes take 32 consecutive segments of group
// `grp` in lock step over the group's widest segment and accumulate FC frames starting at frame f0.
template <int FC, int PT>
__device__ __forceinline__ void mel_item(const float *Pbuf, float *s_U, float *s_V, const StftParams &p, int grp, int f0, int lane,
                                         int nseg, int NC) {
    const int sgm = grp * 32 + lane;
    const int steps = __ldg(p.mel_gsteps + grp);
    const float2 *ww = p.mel_ww + __ldg(p.mel_goff + grp) + lane;
    const int k0 = sgm < nseg ? __ldg(p.mel_seg_start + sgm) : 0;
    float u[FC], v[FC];
#pragma unroll
    for (int f = 0; f < FC; ++f) u[f] = v[f] = 0.f;
#pragma unroll 2
    for (int j = 0; j < steps; ++j) {
        const float2 wv = __ldg(ww + j * 32);
        const float *pp = Pbuf + min(k0 + j, NC) * PT + f0;
#pragma unroll
        for (int f = 0; f < FC; ++f) {
            const float x = pp[f];
            u[f] = fmaf(wv.x, x, u[f]);
            v[f] = fmaf(wv.y, x, v[f]);
        }
    }
    if (sgm < nseg) {
#pragma unroll
        for (int f = 0; f < FC; ++f) {
            s_U[sgm * PT + f0 + f] = u[f];
            s_V[sgm * PT + f0 + f] = v[f];
        }
    }
}

// Stores `rows` x TT values of a tile as T-contiguous rows, VW frames per thread and store (VW = 4 / 2 when the clip's rows
// are 16 / 8-byte aligned; a tile starts at a multiple of 8 frames, so a vector never straddles the end of the clip).
// `load(row, t)` fetches the raw value, `conv` maps it to the stored one; returns the thread's maximum raw value.
template <int VW, int TT, typename LoadFn, typename ConvFn>
__device__ __forceinline__ float store_tile(int rows, float *__restrict__ out, int T, int t0, int tid, LoadFn load, ConvFn conv) {
    constexpr int QT = TT / VW;
    float vmax = 0.f;
    for (int idx = tid; idx < rows * QT; idx += kThreads) {
        const int t = (idx % QT) * VW, r = idx / QT;
        if (t0 + t < T) {
            float v[VW];
#pragma unroll
            for (int j = 0; j < VW; ++j) {
                const float x = load(r, t + j);
                vmax = fmaxf(vmax, x);
                v[j] = conv(x);
            }
            float *dst = out + (long long)r * T + t0 + t;
            if (VW == 4) *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            else if (VW == 2) *reinterpret_cast<float2 *>(dst) = make_float2(v[0], v[1]);
            else dst[0] = v[0];
        }
    }
    return vmax;
}

template <int NC, int MODE>
__global__ void AMT_FFT_BOUNDS(AMT_STFT_MINB) stft_kernel(const StftParams p) {
    using L = FftLayout<NC>;
    constexpr int G = L::G, S = L::S, TT = kWarpsPerCta * G, PT = TT + 1, NFFT = 2 * NC;
    extern __shared__ __align__(16) float smem[];
    float2 *s_tw1 = reinterpret_cast<float2 *>(smem);                // NC
    float2 *s_tw2 = s_tw1 + NC;                                      // NC/2 (k = 0 .. NC/2 - 1)
    float *s_scr = reinterpret_cast<float *>(s_tw2 + NC / 2);        // 8 * WARP_PITCH; later Pbuf[k][frame] (+ mel staging)
    float *s_tile = s_scr + p.scr_floats;                            // audio tile

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ClipMeta *cm = p.meta + blockIdx.y;
    const int T = cm->T;
    // A CTA walks `tiles_per_cta` consecutive tiles of one clip: the twiddles are staged once, and the audio of the
    // next tile streams into shared memory (cp.async) while the current tile is projected and stored.
    const int ntiles = (T + TT - 1) / TT;
    const int tile_begin = blockIdx.x * p.tiles_per_cta;
    if (tile_begin >= ntiles) return;
    const int tile_end = min(ntiles, tile_begin + p.tiles_per_cta);
    const float *src = p.audio + cm->in_off;
    const bool overlap = p.hop <= NFFT;
    float *out = p.out + cm->out_off;
    float vmax = 0.f;
    const int vw = ((T & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) ? 4
                   : ((T & 1) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0) ? 2 : 1;

    for (int i = tid; i < NC; i += kThreads) s_tw1[i] = p.tw1[i];
    for (int i = tid; i < NC / 2; i += kThreads) s_tw2[i] = p.tw2[i];
    int shift = 0, fstride = overlap ? p.hop : NFFT;
    if (overlap) {
        load_tile_async(s_tile, src, cm->n, (long long)tile_begin * TT * p.hop - p.pad, p.hop, NFFT, TT, tid, shift);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    }

    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int t0 = tile * TT;
        if (overlap) {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        } else {
            __syncthreads();   // previous iteration's readers of the scratch region are done
            load_tile(s_tile, src, cm->n, (long long)t0 * p.hop - p.pad, p.hop, NFFT, TT, tid, shift, fstride);
        }
        __syncthreads();       // tile (and, first time, the twiddles) visible; previous iteration's staging retired

        float2 *scr = reinterpret_cast<float2 *>(s_scr + warp * L::WARP_PITCH);
        {
            const float2 *win2 = reinterpret_cast<const float2 *>(p.window);
            const bool vec_ok = ((shift | fstride) & 1) == 0;
            if (vec_ok) {
                warp_fft_unit<NC>(scr, s_tw1, lane, [&](int g, int n) {
                    const float2 x = *reinterpret_cast<const float2 *>(s_tile + shift + (warp * G + g) * fstride + 2 * n);
                    const float2 w = __ldg(win2 + n);
                    return make_float2(x.x * w.x, x.y * w.y);
                });
            } else {
                warp_fft_unit<NC>(scr, s_tw1, lane, [&](int g, int n) {
                    const float *x = s_tile + shift + (warp * G + g) * fstride + 2 * n;
                    const float2 w = __ldg(win2 + n);
                    return make_float2(x[0] * w.x, x[1] * w.y);
                });
            }
        }

        // real-FFT split -> power spectrum in registers; after the barrier (scratch retired) it is written TRANSPOSED as
        // Pbuf[k][frame] (frame fastest, odd pitch) so both the store below and the consumers are conflict free
        float *Pbuf = s_scr;
        {
            float pw[16][2];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int q = lane + 32 * j;
                const int g = q / (NC / 2), k = q % (NC / 2);
                const float2 A = scr[g * S + k], B = scr[g * S + ((NC - k) & (NC - 1))];
                float2 E, Tw;
                rfft_split(A, B, s_tw2[k], E, Tw);
                const float ax = E.x + Tw.x, ay = E.y + Tw.y, bx = E.x - Tw.x, by = E.y - Tw.y;
                pw[j][0] = fmaf(ax, ax, ay * ay);
                pw[j][1] = fmaf(bx, bx, by * by);
            }
            float pmid = 0.f;
            if (lane < G) {
                const float2 A = scr[lane * S + NC / 2];
                pmid = fmaf(A.x, A.x, A.y * A.y);
            }
            __syncthreads();   // every warp is done with the audio tile and its FFT scratch
            if (overlap && tile + 1 < tile_end) {
                load_tile_async(s_tile, src, cm->n, (long long)(t0 + TT) * p.hop - p.pad, p.hop, NFFT, TT, tid, shift);
                asm volatile("cp.async.commit_group;\n" ::: "memory");
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int q = lane + 32 * j;
                const int g = q / (NC / 2), k = q % (NC / 2);
                Pbuf[k * PT + warp * G + g] = pw[j][0];
                Pbuf[(NC - k) * PT + warp * G + g] = pw[j][1];
            }
            if (lane < G) Pbuf[(NC / 2) * PT + warp * G + lane] = pmid;
        }
        __syncthreads();

        if (MODE == MODE_MEL) {
            // Sparse mel projection, segment form: FFT bin k lies between two mel centres and feeds exactly the rising slope
            // of filter seg(k) and the falling slope of filter seg(k) - 1, so every power value is read once:
            //   U[s] = sum_{k in segment s} up_k P[k],  V[s] = sum down_k P[k],  mel[m] = U[m] + V[m + 1].
            // One thread per (segment, chunk of 8 frames); the (up, down) weights are padded lane-major per group of 32 segments.
            // (A host-balanced schedule of (group, 2/4/8-frame chunk) items over the warps was measured: slower, the extra
            // weight loads cost more than the imbalance the second resident CTA already hides.)
            float *s_U = Pbuf + (NC + 1) * PT, *s_V = s_U + (p.n_mels + 1) * PT;
            const int nseg = p.n_mels + 1;
            {
                const int ngroups = (nseg + 31) >> 5;
                constexpr int NCHUNK = TT / 8;
                for (int w = warp; w < ngroups * NCHUNK; w += kWarpsPerCta)
                    mel_item<8, PT>(Pbuf, s_U, s_V, p, w / NCHUNK, (w % NCHUNK) * 8, lane, nseg, NC);
            }
            __syncthreads();
            auto load = [&](int m, int t) { return s_U[m * PT + t] + s_V[(m + 1) * PT + t]; };
            auto conv = [&](float v) { return p.decibels ? db10(fmaxf(1e-10f, v)) : v; };
            if (vw == 4) vmax = fmaxf(vmax, store_tile<4, TT>(p.n_mels, out, T, t0, tid, load, conv));
            else if (vw == 2) vmax = fmaxf(vmax, store_tile<2, TT>(p.n_mels, out, T, t0, tid, load, conv));
            else vmax = fmaxf(vmax, store_tile<1, TT>(p.n_mels, out, T, t0, tid, load, conv));
        } else {
            auto load = [&](int k, int t) { return Pbuf[k * PT + t]; };
            auto conv = [&](float v) { return p.decibels ? db10(fmaxf(1e-10f, v)) : sqrtf(v); };
            if (vw == 4) vmax = fmaxf(vmax, store_tile<4, TT>(NC + 1, out, T, t0, tid, load, conv));
            else if (vw == 2) vmax = fmaxf(vmax, store_tile<2, TT>(NC + 1, out, T, t0, tid, load, conv));
            else vmax = fmaxf(vmax, store_tile<1, TT>(NC + 1, out, T, t0, tid, load, conv));
        }
    }
    if (p.decibels) {
        vmax = warp_max(vmax);
        if (lane == 0) atomic_max_nonneg(p.maxbuf + blockIdx.y, vmax);
    }
}

// ------------------------------------------------------------------------------------------------
// K7 : SignalPower (power.py:31-57) and WaveformWrapper framing (waveform.py:121-153)
// ------------------------------------------------------------------------------------------------

__global__ void __launch_bounds__(kThreads) power_kernel(const float *__restrict__ audio, float *__restrict__ out,
                                                          const ClipMeta *__restrict__ meta, float *maxbuf, int hop, int win,
                                                          int pad, int decibels) {
    const ClipMeta cm = meta[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = blockIdx.x * kWarpsPerCta + warp;
    if (t >= cm.T) return;
    const float *src = audio + cm.in_off;
    const long long s0 = (long long)t * hop - pad;
    float acc = 0.f;
    for (int i = lane; i < win; i += 32) {
        const long long g = s0 + i;
        const float x = (g >= 0 && g < cm.n) ? __ldg(src + g) : 0.f;
        acc = fmaf(x, x, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const float pw = acc / (float)win;
        if (decibels) {
            // amplitude_to_db applied to the power (power.py:55): 10 log10(max(1e-10, pw^2)) - ref
            const float sq = pw * pw;
            out[cm.out_off + t] = db10(fmaxf(1e-10f, sq));
            atomic_max_nonneg(maxbuf + blockIdx.y, sq);
        } else {
            out[cm.out_off + t] = pw;
        }
    }
}

__global__ void __launch_bounds__(kThreads) frames_kernel(const float *__restrict__ audio, float *__restrict__ out,
                                                           const ClipMeta *__restrict__ meta, int hop, int win, int pad) {
    const ClipMeta cm = meta[blockIdx.z];
    const int t = blockIdx.x * 32 + (threadIdx.x & 31);
    const int j = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (t >= cm.T || j >= win) return;
    const long long g = (long long)t * hop - pad + j;
    out[cm.out_off + (long long)j * cm.T + t] = (g >= 0 && g < cm.n) ? __ldg(audio + cm.in_off + g) : 0.f;
}

// Device-side `tools.framify_activations` (amt_tools/tools/utils.py:2922-2984, used by TabCNN.pre_proc,
// models/tabcnn.py:123-127): out[r][i][w] = padded[r][i * hop + w], zero padding of `lpad` frames on the left.
__global__ void __launch_bounds__(kThreads) framify_kernel(const float *__restrict__ in, float *__restrict__ out, long long rows,
                                                            long long T, int win, int hop, long long lpad, long long hops) {
    const long long per_row = hops * win, total = rows * per_row;
    for (long long idx = (long long)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (long long)gridDim.x * kThreads) {
        const long long r = idx / per_row, rem = idx - r * per_row;
        const long long i = rem / win;
        const int w = (int)(rem - i * win);
        const long long t = i * hop + w - lpad;
        out[idx] = (t >= 0 && t < T) ? __ldg(in + r * T + t) : 0.f;
    }
}

int framify(const float *d_in, long long rows, long long T, int win, int hop, long long lpad, long long hops, float *d_out,
            void *stream) {
    const long long total = rows * hops * win;
    if (total <= 0) return AMTFEAT_OK;
    const unsigned grid = (unsigned)std::min<long long>((total + kThreads - 1) / kThreads, 148 * 32);
    framify_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_in, d_out, rows, T, win, hop, lpad, hops);
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

// ------------------------------------------------------------------------------------------------
// K6 : dB epilogue (common.py:199 + 224-225, mel.py:94, power.py:55)
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ float db_finish(float v, float ref_db, int scale01) {
    v = fmaxf(v - ref_db, -80.0f);
    return scale01 ? v / 80.0f + 1.0f : v;
}

// `dst` == `src`: in place (device-resident consumers).  `dst` = mapped pinned host memory (amtfeat_pipeline_*, same element offsets):
// the features are read from HBM once and leave for the host in the same pass -- no second pass over HBM and no separate copy.
__global__ void __launch_bounds__(kThreads, AMT_EPI_MINB) db_epilogue_kernel(const float *src, float *dst, const ClipMeta *__restrict__ meta,
                                                                const float *__restrict__ maxbuf, int C, int F, int scale01) {
    const int seg = blockIdx.y, b = seg / C, c = seg % C;
    const ClipMeta *cm = meta + b;
    const long long count = (long long)F * cm->T;
    const float *in = src + cm->out_off + (long long)c * count;
    float *o = dst + cm->out_off + (long long)c * count;
    const float ref_db = db10(fmaxf(1e-10f, maxbuf[seg]));
    // scalar head up to 16-byte alignment, float4 body, scalar tail (src and dst share the element offset, hence the alignment)
    const long long head = min(count, (long long)(((16 - (reinterpret_cast<uintptr_t>(o) & 15)) & 15) >> 2));
    const long long nvec = (count - head) >> 2;
    const long long tail0 = head + (nvec << 2);
    const float4 *i4 = reinterpret_cast<const float4 *>(in + head);
    float4 *o4 = reinterpret_cast<float4 *>(o + head);
    auto finish4 = [&](float4 v) {
        v.x = db_finish(v.x, ref_db, scale01);
        v.y = db_finish(v.y, ref_db, scale01);
        v.z = db_finish(v.z, ref_db, scale01);
        v.w = db_finish(v.w, ref_db, scale01);
        return v;
    };
    // AMT_EPI_UNROLL independent 16-byte loads in flight per thread at 64 resident warps per SM (measured, in place over 1.48 GB:
    // one load, 48 warps 0.585 ms; two loads, 64 warps 0.422 ms = 7.0 TB/s; four loads at 32 / 48 warps 0.486 / 0.436 ms)
    constexpr int U = AMT_EPI_UNROLL;
    const long long stride = (long long)gridDim.x * kThreads;
    long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) v[u] = AMT_EPI_CS ? __ldcs(i4 + i + u * stride) : i4[i + u * stride];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (AMT_EPI_CS) __stcs(o4 + i + u * stride, finish4(v[u]));
            else o4[i + u * stride] = finish4(v[u]);
        }
    }
    for (; i < nvec; i += stride) o4[i] = finish4(i4[i]);
    if (blockIdx.x == 0) {
        if (threadIdx.x < head) o[threadIdx.x] = db_finish(in[threadIdx.x], ref_db, scale01);
        if (tail0 + threadIdx.x < count) o[tail0 + threadIdx.x] = db_finish(in[tail0 + threadIdx.x], ref_db, scale01);
    }
}

// ------------------------------------------------------------------------------------------------
// Long tracks cut into chunks: `ref=np.max` (common.py:199, 224-225) runs over the WHOLE track, so a track whose
// chunks are computed apart (several GPUs, or one after the other) needs the reference level of every chunk's own frames
// before any element can be finished.  range_max_kernel reduces the raw (un-referenced) log values of the frames a chunk
// keeps (its halo frames belong to the neighbours); after the maxima of all chunks are combined (a max over C floats: the one
// exchange step of the path), range_finish_kernel applies max(v - ref, -80) / 80 + 1 to the kept frames while moving them to
// their place in the track's (C, F, T) block.  db10 is monotonic, so the maximum of the raw log values IS the log of the
// maximum: the reference is bit-identical with the one db_epilogue_kernel forms from the raw maximum.
// ------------------------------------------------------------------------------------------------

__device__ __forceinline__ void atomic_max_float(float *addr, float v) {
    int *ai = reinterpret_cast<int *>(addr);
    int old = *ai;
    while (__int_as_float(old) < v) {
        const int assumed = old;
        old = atomicCAS(ai, assumed, __float_as_int(v));
        if (old == assumed) break;
    }
}

// grid (frame tiles, channels): a thread walks the rows of its frames (coalesced along T, four rows in flight)
__global__ void __launch_bounds__(kThreads) range_max_kernel(const float *__restrict__ blk, int F, long long T, long long t0, long long t1,
                                                              float *ref) {
    const float *src = blk + (long long)blockIdx.y * F * T;
    float m = -INFINITY;
    for (long long t = t0 + (long long)blockIdx.x * kThreads + threadIdx.x; t < t1; t += (long long)gridDim.x * kThreads) {
        const float *q = src + t;
        int f = 0;
        for (; f + 4 <= F; f += 4) {
            const float a = q[(long long)f * T], b = q[(long long)(f + 1) * T], c = q[(long long)(f + 2) * T], d = q[(long long)(f + 3) * T];
            m = fmaxf(m, fmaxf(fmaxf(a, b), fmaxf(c, d)));
        }
        for (; f < F; ++f) m = fmaxf(m, q[(long long)f * T]);
    }
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0 && m > -INFINITY) atomic_max_float(ref + blockIdx.y, m);
}

// grid (frame tiles, rows, channels); mode 0: copy (linear features), 1: dB rescaled to [0, 1], 2: dB (SignalPower).  A thread
// moves four frames a tile apart per step (four independent loads in flight; source and destination rows start at arbitrary
// frames, so the accesses stay 4 bytes wide and coalesced).
__global__ void __launch_bounds__(kThreads) range_finish_kernel(const float *__restrict__ blk, int F, long long T, long long t0, long long t1,
                                                                 const float *__restrict__ ref, float *__restrict__ dst, long long T_dst,
                                                                 long long t_dst, int mode) {
    const int c = blockIdx.z;
    const float ref_db = mode ? ref[c] : 0.f;
    const long long step = (long long)gridDim.x * kThreads * 4;
    for (int f = blockIdx.y; f < F; f += gridDim.y) {
        const float *src = blk + ((long long)c * F + f) * T;
        float *o = dst + ((long long)c * F + f) * T_dst + t_dst - t0;
        for (long long t = t0 + (long long)blockIdx.x * kThreads * 4 + threadIdx.x; t < t1; t += step) {
            float v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = t + u * kThreads < t1 ? src[t + u * kThreads] : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (t + u * kThreads < t1) o[t + u * kThreads] = mode ? db_finish(v[u], ref_db, mode == 1) : v[u];
        }
    }
}

static int range_mode(const Plan &p) { return !has_db_epilogue(p) ? 0 : p.cfg.kind == AMTFEAT_POWER ? 2 : 1; }

int range_reference(const Plan &p, const float *d_block, int64_t frames, int64_t t_begin, int64_t t_end, float *d_ref, void *stream) {
    if (p.device < 0) { set_error("host-only plan: no CUDA device (there is no CPU compute path)"); return AMTFEAT_ERR_NO_DEVICE; }
    if (t_begin < 0 || t_end > frames || t_begin > t_end) { set_error("frame range outside the block"); return AMTFEAT_ERR_INVALID; }
    if (t_begin == t_end || !has_db_epilogue(p)) return AMTFEAT_OK;
    DeviceGuard guard(p.device);
    const int64_t w = t_end - t_begin;
    dim3 grid((unsigned)std::min<int64_t>((w + kThreads - 1) / kThreads, 148 * 8), p.C);
    range_max_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_block, p.F, frames, t_begin, t_end, d_ref);
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

int range_finish(const Plan &p, const float *d_block, int64_t frames, int64_t t_begin, int64_t t_end, const float *d_ref, float *d_dst,
                 int64_t dst_frames, int64_t t_dst, void *stream) {
    if (p.device < 0) { set_error("host-only plan: no CUDA device (there is no CPU compute path)"); return AMTFEAT_ERR_NO_DEVICE; }
    if (t_begin < 0 || t_end > frames || t_begin > t_end || t_dst < 0 || t_dst + (t_end - t_begin) > dst_frames) {
        set_error("frame range outside the source or the destination block");
        return AMTFEAT_ERR_INVALID;
    }
    if (t_begin == t_end) return AMTFEAT_OK;
    DeviceGuard guard(p.device);
    const int64_t w = t_end - t_begin;
    dim3 grid((unsigned)std::min<int64_t>((w + 4 * kThreads - 1) / (4 * kThreads), 148 * 8), (unsigned)std::min(p.F, 1024), p.C);
    range_finish_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_block, p.F, frames, t_begin, t_end, d_ref, d_dst,
                                                                                        dst_frames, t_dst, range_mode(p));
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

// ------------------------------------------------------------------------------------------------
// K4 : 2:1 decimator.  y[m] = sum_k h[k] x[2m + D - k], D = (ntaps - 1) / 2 (even), zero extension,
// h already scaled by sqrt(2) (librosa.resample(scale=True)).  Split into the two polyphase branches
// so every shared-memory access is unit stride.
// ------------------------------------------------------------------------------------------------

constexpr int kDecR = 16;                       // consecutive outputs per thread (register sliding window)
constexpr int kDecTile = kThreads * kDecR;      // outputs per CTA

// Skewed shared-memory index: 4 extra floats per 16.  A thread's window starts at 16 * lane + c, i.e. 16-byte quad
// 4 * lane + c / 4; with the skew the quad becomes 5 * lane + const (mod 8 bank groups), a bijection for every
// alignment c, so the 16-byte loads of a quarter warp never collide.
__device__ __forceinline__ int dec_phys(int i) { return i + ((i >> 4) << 2); }

__host__ __device__ inline int dec_front_pad(int D) { return ((((7 - D) % 4) + 4) % 4) + 8; }
__host__ __device__ inline int dec_jtot(int D) { return (D + 1 + 7) / 8 * 8; }

// One polyphase branch: acc[r] += sum_j h[j] * A[wbase - j + r], A skewed in shared memory.
__device__ __forceinline__ void dec_branch(const float *__restrict__ A, const float *__restrict__ h, int jtot, int wbase,
                                           float (&acc)[kDecR]) {
#pragma unroll 1
    for (int j0 = 0; j0 < jtot; j0 += 8) {
        const float4 h0 = *reinterpret_cast<const float4 *>(h + j0), h1 = *reinterpret_cast<const float4 *>(h + j0 + 4);
        const float hv[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        float v[24];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
            const float4 x = *reinterpret_cast<const float4 *>(A + dec_phys(wbase - 7 - j0 + 4 * q));
            v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
        }
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
#pragma unroll
            for (int r = 0; r < kDecR; ++r) acc[r] = fmaf(hv[jj], v[7 - jj + r], acc[r]);
    }
}

__global__ void __launch_bounds__(kThreads) decimate_kernel(const float *__restrict__ audio, float *__restrict__ ladder,
                                                             const ClipMeta *__restrict__ meta, const float *__restrict__ taps,
                                                             int ntaps, int level_out) {
    extern __shared__ __align__(16) float smem[];
    const ClipMeta *cm = meta + blockIdx.y;
    const int len_out = cm->lvl_len[level_out], len_in = cm->lvl_len[level_out - 1];
    const int m0 = blockIdx.x * kDecTile;
    if (m0 >= len_out) return;
    const float *src = (level_out == 1 ? audio : ladder) + cm->lvl_off[level_out - 1];
    float *dst = ladder + cm->lvl_off[level_out];
    const int D = (ntaps - 1) / 2;                 // even (host pads the taps), so `base` below is even
    const int FP = dec_front_pad(D), jtot = dec_jtot(D);
    const int len = FP + kDecTile + D + 8;         // logical length of each branch array
    const int plen = (dec_phys(len) + 7) & ~3;
    float *E = smem, *O2 = E + plen, *he = O2 + plen, *ho = he + jtot;
    const long long base = 2ll * m0 - D;
    // E[FP + e] = x[base + 2e],  O2[FP + e] = x[base + 2e - 1]   (e = 0 .. kDecTile + D - 1)
    for (int e = threadIdx.x; e < kDecTile + D; e += kThreads) {
        const long long g = base + 2ll * e;
        float2 v;
        if (g >= 0 && g + 1 < len_in) {
            v = __ldg(reinterpret_cast<const float2 *>(src + g));
        } else {
            v.x = (g >= 0 && g < len_in) ? __ldg(src + g) : 0.f;
            v.y = (g + 1 >= 0 && g + 1 < len_in) ? __ldg(src + g + 1) : 0.f;
        }
        E[dec_phys(FP + e)] = v.x;
        O2[dec_phys(FP + e + 1)] = v.y;
    }
    for (int i = threadIdx.x; i < FP; i += kThreads) { E[dec_phys(i)] = 0.f; O2[dec_phys(i)] = 0.f; }
    if (threadIdx.x == 0) O2[dec_phys(FP)] = 0.f;
    for (int i = threadIdx.x; i < 8; i += kThreads) { E[dec_phys(FP + kDecTile + D + i)] = 0.f; }
    for (int j = threadIdx.x; j < jtot; j += kThreads) {
        he[j] = (2 * j < ntaps) ? __ldg(taps + 2 * j) : 0.f;
        ho[j] = (2 * j + 1 < ntaps) ? __ldg(taps + 2 * j + 1) : 0.f;
    }
    __syncthreads();
    float acc[kDecR];
#pragma unroll
    for (int r = 0; r < kDecR; ++r) acc[r] = 0.f;
    const int wbase = FP + kDecR * threadIdx.x + D;   // FP + D - 7 is a multiple of 4 by construction
    dec_branch(E, he, jtot, wbase, acc);
    dec_branch(O2, ho, jtot, wbase, acc);
    const int m = m0 + kDecR * threadIdx.x;
    if (m + kDecR <= len_out) {
#pragma unroll
        for (int q = 0; q < kDecR / 4; ++q)
            *reinterpret_cast<float4 *>(dst + m + 4 * q) = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
    } else {
#pragma unroll
        for (int r = 0; r < kDecR; ++r)
            if (m + r < len_out) dst[m + r] = acc[r];
    }
}

// ------------------------------------------------------------------------------------------------
// K4 (fast form) : the same 2:1 decimator as overlap-save fast convolution.
//   A block is 2048 input samples u[n] = x[2 m0 - D + n]; its circular convolution with the taps is exact for
//   n >= 2 D, and keeping every other sample folds the 2048-point spectrum C = U H onto 1024 points:
//   Cd[k] = (C[k] + conj(C[1024 - k])) / 2, y[m0 + j - D] = irfft_1024(Cd)[j], j = D .. 1023  (1024 - D outputs).
//   One warp takes TWO blocks: forward real FFT of each (one 1024-point complex warp unit + split, fused with the
//   multiplication by H and the fold), then ONE 1024-point complex transform inverts both at once
//   (S = CdA + i CdB with Hermitian extension; IDFT(S) = conj(DFT(conj S)) = ydA + i ydB).
//   ~100 flop per output sample instead of 2 * 389 for the direct form.
// ------------------------------------------------------------------------------------------------

constexpr int kDfWarps = 6, kDfThreads = kDfWarps * 32;
constexpr int kDfCd = 520;   // float2 per warp for the folded spectrum of the first block (513 used)

struct DecFftParams {
    const float *audio;
    float *ladder;
    const ClipMeta *meta;
    const float2 *tw1, *tw2;
    const float4 *hh;
    int level_out, D, M;
};

// Folded, filtered spectrum of the block whose packed complex FFT sits in scr: Cd[k], k = lane + 32 j.
__device__ __forceinline__ void dec_fold(const float2 *scr, const DecFftParams &p, int lane, float2 (&cd)[17]) {
#pragma unroll
    for (int j = 0; j < 17; ++j) {
        const int k = lane + 32 * j;
        if (k <= 512) {
            const float2 A = scr[k], B = scr[(1024 - k) & 1023];
            float2 E, T;
            rfft_split(A, B, __ldg(p.tw2 + k), E, T);
            const float4 h = __ldg(p.hh + k);
            const float2 up = make_float2(E.x + T.x, E.y + T.y), um = make_float2(E.x - T.x, E.y - T.y);
            const float2 c0 = cmul(make_float2(h.x, h.y), up), c1 = cmul(make_float2(h.z, h.w), um);
            cd[j] = make_float2(c0.x + c1.x, c0.y + c1.y);
        }
    }
}

__global__ void __launch_bounds__(kDfThreads, 2) decimate_fft_kernel(const DecFftParams p) {
    using L = FftLayout<1024>;
    constexpr int WP2 = L::WARP_PITCH / 2;
    extern __shared__ __align__(16) float smem[];
    float2 *s_tw1 = reinterpret_cast<float2 *>(smem);      // 1024
    float2 *s_scr = s_tw1 + 1024;                          // kDfWarps * WP2
    float2 *s_cd = s_scr + kDfWarps * WP2;                 // kDfWarps * kDfCd
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ClipMeta *cm = p.meta + blockIdx.y;
    const int len_out = cm->lvl_len[p.level_out], len_in = cm->lvl_len[p.level_out - 1];
    if ((long long)blockIdx.x * kDfWarps * 2 * p.M >= len_out) return;
    for (int i = tid; i < 1024; i += kDfThreads) s_tw1[i] = p.tw1[i];
    __syncthreads();
    const long long mA = ((long long)blockIdx.x * kDfWarps + warp) * 2 * p.M, mB = mA + p.M;
    if (mA >= len_out) return;
    const float *src = (p.level_out == 1 ? p.audio : p.ladder) + cm->lvl_off[p.level_out - 1];
    float *dst = p.ladder + cm->lvl_off[p.level_out];
    float2 *scr = s_scr + warp * WP2, *cdA = s_cd + warp * kDfCd;

    // The three transforms (forward of block A, forward of block B, inverse of both) run through ONE copy of the FFT
    // code: the kernel is a single pass over straight-line code, so on the short deep ladder levels its time is
    // instruction-fetch latency, and a third of the code means a third of the cold misses.
    const bool haveB = mB < len_out;
    float2 cd[17];
#pragma unroll 1
    for (int ph = 0; ph < 3; ++ph) {
        if (ph == 1 && !haveB) {
#pragma unroll
            for (int j = 0; j < 17; ++j) cd[j] = make_float2(0.f, 0.f);
        } else {
            const long long base = 2 * (ph == 0 ? mA : mB) - p.D;               // even: D is even
            const bool interior = base >= 0 && base + 2048 <= len_in;
            warp_fft_unit<1024>(scr, s_tw1, lane, [&](int, int n) {
                if (ph == 2) return scr[(n >> 5) * 33 + (n & 31)];
                const long long i = base + 2 * n;
                if (interior) return __ldg(reinterpret_cast<const float2 *>(src + i));
                float2 v;
                v.x = (i >= 0 && i < len_in) ? __ldg(src + i) : 0.f;
                v.y = (i + 1 >= 0 && i + 1 < len_in) ? __ldg(src + i + 1) : 0.f;
                return v;
            });
            if (ph == 2) break;
            dec_fold(scr, p, lane, cd);
        }
        if (ph == 0) {
#pragma unroll
            for (int j = 0; j < 17; ++j)
                if (lane + 32 * j <= 512) cdA[lane + 32 * j] = cd[j];
            __syncwarp();
        } else {
            __syncwarp();
            // conj(S) in the padded [n1][n2] layout the transform's first pass reads and overwrites in place
#pragma unroll
            for (int j = 0; j < 17; ++j) {
                const int k = lane + 32 * j;
                if (k <= 512) {
                    const float2 a = cdA[k], b = cd[j];
                    scr[(k >> 5) * 33 + (k & 31)] = make_float2(a.x - b.y, -a.y - b.x);
                    if (k >= 1 && k <= 511) {
                        const int n = 1024 - k;
                        scr[(n >> 5) * 33 + (n & 31)] = make_float2(a.x + b.y, a.y - b.x);
                    }
                }
            }
            __syncwarp();
        }
    }
    // scr[n] = conj(ydA[n] + i ydB[n])
#pragma unroll 4
    for (int j = 0; j < 32; ++j) {
        const int n = lane + 32 * j;
        if (n >= p.D) {
            const float2 r = scr[n];
            const long long ma = mA + n - p.D, mb = mB + n - p.D;
            if (ma < len_out) dst[ma] = r.x;
            if (haveB && mb < len_out) dst[mb] = -r.y;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4 (default form) : the same overlap-save decimator in FLOAT64.
//   The float32 forms above are ~1e-7 of the block amplitude off per ladder step, white, and that noise stays in the
//   band while the strong high-frequency content it came from is filtered away: on the deep levels it was the largest
//   term of the dB error (measured: 6e-4 .. 1e-3 dB on bins within 60 dB of the maximum; with an exact ladder the rest
//   of the path is at 2e-4 .. 3e-4).  Here every level is the correctly rounded float32 of an exact decimation.
//   One CTA takes the two blocks as ONE complex signal z = uA + i uB (h is real, so z (*) h = uA (*) h + i uB (*) h):
//   2048-point complex DFT, spectral product with H, fold W[k] + W[k + 1024] (keeps every other output sample),
//   1024-point inverse DFT -> ydA + i ydB.  Stockham autosort passes in shared memory (radix 4, one radix 2), twiddles from
//   a float64 table; ~100 float64 flop per output sample.
// ------------------------------------------------------------------------------------------------

constexpr int kD64Threads = 128;

struct Dec64Params {
    const float *audio;
    float *ladder;
    const ClipMeta *meta;
    const double2 *tw;   // per-pass twiddle tables (host_plan.cpp decim_tw64), see tw64_off
    const double2 *H;    // response of the taps, k < 2048, times 1 / 2048
    int level_out, D, M, pairs_per_cta;
};

__device__ __forceinline__ double2 dmul(double2 a, double2 b) {
    return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ double2 dadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 dsub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

// exp(-2 pi i m / 32), m = 0..15, float64
__device__ __forceinline__ double2 w32d(int m) {
    switch (m) {
        case 0: return make_double2(1.0, -0.0);
        case 1: return make_double2(0.98078528040323043, -0.19509032201612825);
        case 2: return make_double2(0.92387953251128674, -0.38268343236508978);
        case 3: return make_double2(0.83146961230254524, -0.55557023301960218);
        case 4: return make_double2(0.70710678118654757, -0.70710678118654757);
        case 5: return make_double2(0.55557023301960218, -0.83146961230254524);
        case 6: return make_double2(0.38268343236508978, -0.92387953251128674);
        case 7: return make_double2(0.19509032201612825, -0.98078528040323043);
        case 8: return make_double2(0.0, -1.0);
        case 9: return make_double2(-0.19509032201612825, -0.98078528040323043);
        case 10: return make_double2(-0.38268343236508978, -0.92387953251128674);
        case 11: return make_double2(-0.55557023301960218, -0.83146961230254524);
        case 12: return make_double2(-0.70710678118654757, -0.70710678118654757);
        case 13: return make_double2(-0.83146961230254524, -0.55557023301960218);
        case 14: return make_double2(-0.92387953251128674, -0.38268343236508978);
        default: return make_double2(-0.98078528040323043, -0.19509032201612825);
    }
}

// In-register forward DFT of R points, float64 (radix-2 decimation in frequency; output k lives in register brev<R>(k)).
template <int R> __device__ __forceinline__ void dft_regs(double2 (&v)[R]) {
#pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int blk = 0; blk < R; blk += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const double2 a = v[blk + j], b = v[blk + j + half];
                v[blk + j] = dadd(a, b);
                const double2 d = dsub(a, b);
                const int m = j * (16 / half);  // twiddle exp(-2 pi i j / (2 half)) = w32d(m)
                if (m == 0) v[blk + j + half] = d;
                else if (m == 8) v[blk + j + half] = make_double2(d.y, -d.x);
                else if (m == 4) v[blk + j + half] = make_double2((d.x + d.y) * 0.70710678118654757, (d.y - d.x) * 0.70710678118654757);
                else if (m == 12) v[blk + j + half] = make_double2((d.y - d.x) * 0.70710678118654757, -(d.x + d.y) * 0.70710678118654757);
                else v[blk + j + half] = dmul(d, w32d(m));
            }
        }
    }
}

// Twiddle tables of the passes below, each [r - 1][k] = exp(-2 pi i r k / (R Ns)), k < Ns:
//   forward 2048 = 16 x 16 x 8:  pass B (R 16, Ns 16) at 0 (240 entries), pass C (R 8, Ns 256) at 240 (1792)
//   inverse 1024 = 8 x 8 x 16:   pass E (R 8, Ns 8) at 2032 (56), pass F (R 16, Ns 64) at 2088 (960)
constexpr int kTw64B = 0, kTw64C = 240, kTw64E = 2032, kTw64F = 2088, kTw64Total = 3048;
// shared-memory index padding: one extra 16-byte slot per 16 (forward) / per 8 (inverse) entries, so that the strided stores of
// a pass (thread j writes entries 16 j + r, or 8 t + s) spread over all eight 16-byte bank groups
__device__ __forceinline__ int pf64(int i) { return i + (i >> 4); }
__device__ __forceinline__ int pi64(int i) { return i + (i >> 3); }
constexpr int kD64Buf = 2048 + 2048 / 16;   // double2 entries of the one (in-place) buffer

// One CTA (128 threads, 16 complex float64 values per thread) per pair of blocks: every pass is one register-resident radix-16 /
// radix-8 DFT per thread with the Stockham index map; the data crosses shared memory twice per transform, in place.
__global__ void __launch_bounds__(kD64Threads, 4) decimate_fft64_kernel(const Dec64Params p) {
    extern __shared__ __align__(16) double2 buf[];
    const int t = threadIdx.x;
    const ClipMeta *cm = p.meta + blockIdx.y;
    const int len_out = cm->lvl_len[p.level_out], len_in = cm->lvl_len[p.level_out - 1];
    const int npairs = (int)(((long long)len_out + 2 * p.M - 1) / (2 * p.M));
    const int pair0 = blockIdx.x * p.pairs_per_cta;
    if (pair0 >= npairs) return;
    const int pair1 = min(npairs, pair0 + p.pairs_per_cta);
    const float *src = (p.level_out == 1 ? p.audio : p.ladder) + cm->lvl_off[p.level_out - 1];
    float *dst = p.ladder + cm->lvl_off[p.level_out];
    for (int pair = pair0; pair < pair1; ++pair) {
        const long long mA = (long long)pair * 2 * p.M, mB = mA + p.M;
        const bool haveB = mB < len_out;
        const long long baseA = 2 * mA - p.D, baseB = 2 * mB - p.D;
        const bool interior = baseA >= 0 && baseB + 2048 <= len_in;
        double2 v[16];
        // ---- forward 2048, pass A (radix 16, Ns = 1): z[t + 128 r] = uA + i uB straight from global memory
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const int n = t + 128 * r;
            float xa, xb;
            if (interior) {
                xa = __ldg(src + baseA + n);
                xb = __ldg(src + baseB + n);
            } else {
                const long long ia = baseA + n, ib = baseB + n;
                xa = (ia >= 0 && ia < len_in) ? __ldg(src + ia) : 0.f;
                xb = (haveB && ib >= 0 && ib < len_in) ? __ldg(src + ib) : 0.f;
            }
            v[r] = make_double2((double)xa, (double)xb);
        }
        dft_regs<16>(v);
        if (pair != pair0) __syncthreads();          // the previous pair's last reads of the buffer are done
#pragma unroll
        for (int r = 0; r < 16; ++r) buf[pf64(16 * t + r)] = v[brev<16>(r)];
        __syncthreads();
        // ---- pass B (radix 16, Ns = 16)
        {
            const int k = t & 15;
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = buf[pf64(t + 128 * r)];
#pragma unroll
            for (int r = 1; r < 16; ++r) v[r] = dmul(v[r], __ldg(p.tw + kTw64B + (r - 1) * 16 + k));
            dft_regs<16>(v);
            __syncthreads();
            const int j0 = (t - k) * 16 + k;
#pragma unroll
            for (int r = 0; r < 16; ++r) buf[pf64(j0 + 16 * r)] = v[brev<16>(r)];
        }
        __syncthreads();
        // ---- pass C (radix 8, Ns = 256): butterflies j = t and t + 128; outputs Z[j + 256 r] stay in registers, then the spectral
        //      product and the fold W[k] + W[k + 1024] (keeps every other output sample): Yd[j + 256 r'], r' < 4, conjugated for
        //      the inverse transform (IDFT = conj(DFT(conj .)))
        double2 y[8];                                  // y[s] = conj(Yd[t + 128 s])
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int j = t + 128 * hh;
            double2 z[8];
#pragma unroll
            for (int r = 0; r < 8; ++r) z[r] = buf[pf64(j + 256 * r)];
#pragma unroll
            for (int r = 1; r < 8; ++r) z[r] = dmul(z[r], __ldg(p.tw + kTw64C + (r - 1) * 256 + j));
            dft_regs<8>(z);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const double2 w0 = dmul(z[brev<8>(rr)], __ldg(p.H + j + 256 * rr));
                const double2 w1 = dmul(z[brev<8>(rr + 4)], __ldg(p.H + j + 256 * (rr + 4)));
                y[2 * rr + hh] = make_double2(w0.x + w1.x, -(w0.y + w1.y));
            }
        }
        // ---- inverse 1024, pass D (radix 8, Ns = 1) on the registers
        dft_regs<8>(y);
        __syncthreads();                               // pass C's reads of the buffer are done
#pragma unroll
        for (int s = 0; s < 8; ++s) buf[pi64(8 * t + s)] = y[brev<8>(s)];
        __syncthreads();
        // ---- pass E (radix 8, Ns = 8)
        {
            const int k = t & 7;
#pragma unroll
            for (int s = 0; s < 8; ++s) y[s] = buf[pi64(t + 128 * s)];
#pragma unroll
            for (int s = 1; s < 8; ++s) y[s] = dmul(y[s], __ldg(p.tw + kTw64E + (s - 1) * 8 + k));
            dft_regs<8>(y);
            __syncthreads();
            const int j0 = (t - k) * 8 + k;
#pragma unroll
            for (int s = 0; s < 8; ++s) buf[pi64(j0 + 8 * s)] = y[brev<8>(s)];
        }
        __syncthreads();
        // ---- pass F (radix 16, Ns = 64): 64 butterflies; output n = t + 64 r is conj(ydA[n] + i ydB[n]), stored for n >= D
        if (t < 64) {
#pragma unroll
            for (int r = 0; r < 16; ++r) v[r] = buf[pi64(t + 64 * r)];
#pragma unroll
            for (int r = 1; r < 16; ++r) v[r] = dmul(v[r], __ldg(p.tw + kTw64F + (r - 1) * 64 + t));
            dft_regs<16>(v);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int n = t + 64 * r;
                if (n >= p.D) {
                    const double2 o = v[brev<16>(r)];
                    const long long ma = mA + n - p.D, mb = mB + n - p.D;
                    if (ma < len_out) dst[ma] = (float)o.x;
                    if (haveB && mb < len_out) dst[mb] = (float)(-o.y);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Exact ladders (harmonics with eds >= 2): one level of the TAIL of an exact ladder, direct form, float64 accumulation.
//   level_in == 0: ONE factor : 1 decimation of the audio (librosa __early_downsample: a single resample call);
//   level_in >= 1: the 2:1 step from the previous exact level.  y[m] = sum_k h[k] x[factor m + D - k], zero extension.
// Only samples m >= alt_first[level_out] are produced (clip_tail_layout): a few thousand per clip.
// ------------------------------------------------------------------------------------------------

struct TailParams {
    const float *audio;
    float *ladder;
    const ClipMeta *meta;
    const double *taps;
    int ntaps, factor, level_in, level_out, alt;         // blockIdx.z = 1: the head piece [0, alt_hlen) instead of the tail piece [alt_first, len)
};

// A CTA produces kTailOut = 4 x kThreads consecutive outputs, four per thread (64 apart: conflict-free shared-memory reads that
// share every tap load).  Polyphase form: tap k = F q + ph of output m reads x[F (m - q) + D - ph], so the input span is staged
// in shared memory de-interleaved by phase: X[ph][j] = x[F (m0 + j - nq + 1) + D - ph].  Products of four taps are summed in
// float32 (exact products of float32 data and float32-rounded taps, three additions) and folded into a float64 accumulator:
// the result carries one float32 rounding per 4 taps of its 4-tap partial sums and none from the long summation -- as good as
// the final rounding to float32 -- at a quarter of the float64 operations and shared-memory wavefronts of a plain float64 loop.
// PT = 4 outputs per thread amortises the tap loads; PT = 1 (four times the CTAs) when the pieces of a batch are too few to fill
// the GPU otherwise (short clips: the launch is latency bound, one thread walking all taps of four outputs).
// EXACT: float64 taps and float64 products (one DFMA per tap).  Used for one-shot factors of 8 and more, where almost all of the
// input's energy lies in the stop band: the float32 rounding of the taps and of the 4-tap partial sums is relative to that full-band
// amplitude, and showed as 1.5e-3 dB on a CQT whose top bin sits at 2 % of the Nyquist frequency (tools/fuzz_oracle.py).
template <int PT, bool EXACT>
__global__ void __launch_bounds__(kThreads) tail_decimate_kernel(const TailParams p) {
    constexpr int kTailPerThread = PT, kTailOut = PT * kThreads;
    extern __shared__ __align__(16) float tsm[];
    const ClipMeta *cm = p.meta + blockIdx.y;
    const bool head = blockIdx.z != 0;
    long long first, end;
    float *dst;
    if (head) {
        // a head piece that is part of the tail piece (alt_hoff == alt_off: whole level stored) is produced by the tail pass
        if (cm->alt_hlen[p.alt][p.level_out] <= 0 || cm->alt_first[p.alt][p.level_out] == 0) return;
        first = 0;
        end = cm->alt_hlen[p.alt][p.level_out];
        dst = p.ladder + cm->alt_hoff[p.alt][p.level_out];
    } else {
        first = cm->alt_first[p.alt][p.level_out];
        if (first < 0) return;
        end = cm->lvl_len[p.level_out];
        dst = p.ladder + cm->alt_off[p.alt][p.level_out];
    }
    const long long m0 = first + (long long)blockIdx.x * kTailOut;
    if (m0 >= end) return;
    const float *src;
    long long lo, hi;   // samples of the input piece that exist; everything outside reads as zero (true for < 0 and >= len; the
                        // geometry guarantees that nothing else is ever asked for, tests/test_exact_ladder.py)
    if (p.level_in == 0) {
        src = p.audio + cm->in_off;
        lo = 0;
        hi = cm->n;
    } else if (head && cm->alt_first[p.alt][p.level_in] != 0) {
        src = p.ladder + cm->alt_hoff[p.alt][p.level_in];
        lo = 0;
        hi = cm->alt_hlen[p.alt][p.level_in];
    } else {
        src = p.ladder + cm->alt_off[p.alt][p.level_in];
        lo = cm->alt_first[p.alt][p.level_in];
        hi = cm->lvl_len[p.level_in];
    }
    const int F = p.factor, D = (p.ntaps - 1) / 2;
    const int nq = ((p.ntaps + F - 1) / F + 3) & ~3;      // taps per phase, padded to a multiple of 4 with zeros
    const int J = kTailOut + nq - 1, JP = J | 1;          // entries per phase (odd pitch)
    // hs[ph * nq + q] = taps[F q + ph] (16-byte aligned rows: float4 loads); EXACT: the same table in float64
    float *hs = tsm, *X = tsm + (size_t)F * nq * (EXACT ? 2 : 1);
    double *hd = reinterpret_cast<double *>(tsm);
    for (int i = threadIdx.x; i < F * nq; i += kThreads) {
        const int ph = i / nq, q = i - ph * nq, k = F * q + ph;
        const double h = k < p.ntaps ? __ldg(p.taps + k) : 0.0;
        if (EXACT) hd[i] = h;
        else hs[i] = (float)h;
    }
    const long long g0 = (long long)F * (m0 - nq + 1) + D;     // X[ph][j] = x[g0 + F j - ph]
    for (int i = threadIdx.x; i < F * J; i += kThreads) {
        const int u = i - (F - 1);                                // u = F j - ph, -(F - 1) <= u <= F (J - 1)
        const int ph = (F - 1) - ((u + F - 1) % F), j = (u + ph) / F;
        const long long g = g0 + u;
        X[ph * JP + j] = (g >= lo && g < hi) ? src[g] : 0.f;
    }
    __syncthreads();
    double acc[kTailPerThread];
#pragma unroll
    for (int r = 0; r < kTailPerThread; ++r) acc[r] = 0.0;
    for (int ph = 0; ph < F; ++ph) {
        const float *xp = X + ph * JP + threadIdx.x + nq - 1, *hp = hs + ph * nq;
        if (EXACT) {
            const double *hq = hd + ph * nq;
#pragma unroll 4
            for (int q = 0; q < nq; ++q) {
                const double h = hq[q];
#pragma unroll
                for (int r = 0; r < kTailPerThread; ++r) acc[r] = fma(h, (double)xp[r * kThreads - q], acc[r]);
            }
            continue;
        }
        for (int q = 0; q < nq; q += 4) {
            const float4 h4 = *reinterpret_cast<const float4 *>(hp + q);
#pragma unroll
            for (int r = 0; r < kTailPerThread; ++r) {
                const float *x = xp + r * kThreads - q;
                acc[r] += (double)(fmaf(h4.x, x[0], h4.y * x[-1]) + fmaf(h4.z, x[-2], h4.w * x[-3]));
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kTailPerThread; ++r) {
        const long long m = m0 + threadIdx.x + r * kThreads;
        if (m < end) dst[m] = (float)acc[r];
    }
}

// ------------------------------------------------------------------------------------------------
// K1 + K5 : CQT / VQT / HCQT response of one (ladder level, n_fft) item
// ------------------------------------------------------------------------------------------------

constexpr int kDbufPad = 2;    // even Dbuf pitch: the two frames of a lane are one aligned 16-byte load
// Projection of one block of 4 rows for the TWO consecutive frames (t, t + 1) this lane owns: every step is one 16-byte load of
// the band spectra of both frames + the block's two warp-uniform weight vectors feeding 16 packed FFMA2; |.|^2 / L -> dB or
// magnitude -> 8-byte stores (vec2: the clip's rows are 8-byte aligned and T is even) + lazy per-(clip, channel) maximum.
// `skip`: channels whose frames of this tile / chunk come from an exact ladder instead (group-uniform); `tmax`: per channel, the
// frames its dB maximum runs over (frames in [T, tmax) are computed for the maximum only, hvqt.py:123-128).
__device__ __forceinline__ void project_block2(const CqtBlock4 *bl, const float4 *wt, const float2 *Dp, int DP, float *out, int T, int t,
                                               int decibels, int *s_max, unsigned gmask, int glanes, bool leader, bool vec2,
                                               unsigned skip, const int *__restrict__ tmax) {
    const int ndst = bl->ndst;
    unsigned live_d = (1u << ndst) - 1u;
    if (skip) {     // only the first / last tiles of a clip, and only for plans with exact ladders
        live_d = 0;
        for (int d = 0; d < ndst; ++d)
            if (!((skip >> bl->chan[d]) & 1u)) live_d |= 1u << d;
        if (!live_d) return;
    }
    const int steps = bl->steps;
    const float2 z2 = make_float2(0.f, 0.f);
    float2 rA01 = z2, nA01 = z2, iA01 = z2, jA01 = z2, rA23 = z2, nA23 = z2, iA23 = z2, jA23 = z2;
    float2 rB01 = z2, nB01 = z2, iB01 = z2, jB01 = z2, rB23 = z2, nB23 = z2, iB23 = z2, jB23 = z2;
#pragma unroll 2
    for (int s = 0; s < steps; ++s) {
        const float4 dd = *reinterpret_cast<const float4 *>(Dp + s * DP);
        const float4 wa = wt[2 * s], wb = wt[2 * s + 1];
        const float2 w0 = make_float2(wa.x, wa.y), w1 = make_float2(wa.z, wa.w), w2 = make_float2(wb.x, wb.y), w3 = make_float2(wb.z, wb.w);
        const float2 ax = make_float2(dd.x, dd.x), ay = make_float2(dd.y, dd.y), bx = make_float2(dd.z, dd.z), by = make_float2(dd.w, dd.w);
        rA01 = ffma2(w0, ax, rA01); nA01 = ffma2(w1, ay, nA01); iA01 = ffma2(w0, ay, iA01); jA01 = ffma2(w1, ax, jA01);
        rA23 = ffma2(w2, ax, rA23); nA23 = ffma2(w3, ay, nA23); iA23 = ffma2(w2, ay, iA23); jA23 = ffma2(w3, ax, jA23);
        rB01 = ffma2(w0, bx, rB01); nB01 = ffma2(w1, by, nB01); iB01 = ffma2(w0, by, iB01); jB01 = ffma2(w1, bx, jB01);
        rB23 = ffma2(w2, bx, rB23); nB23 = ffma2(w3, by, nB23); iB23 = ffma2(w2, by, iB23); jB23 = ffma2(w3, bx, jB23);
    }
    const float4 inv = *reinterpret_cast<const float4 *>(bl->inv);
    float pa[4], pb[4];
    {
        const float x0 = rA01.x - nA01.x, y0 = iA01.x + jA01.x, x1 = rA01.y - nA01.y, y1 = iA01.y + jA01.y;
        const float x2 = rA23.x - nA23.x, y2 = iA23.x + jA23.x, x3 = rA23.y - nA23.y, y3 = iA23.y + jA23.y;
        pa[0] = fmaf(x0, x0, y0 * y0) * inv.x; pa[1] = fmaf(x1, x1, y1 * y1) * inv.y;
        pa[2] = fmaf(x2, x2, y2 * y2) * inv.z; pa[3] = fmaf(x3, x3, y3 * y3) * inv.w;
    }
    {
        const float x0 = rB01.x - nB01.x, y0 = iB01.x + jB01.x, x1 = rB01.y - nB01.y, y1 = iB01.y + jB01.y;
        const float x2 = rB23.x - nB23.x, y2 = iB23.x + jB23.x, x3 = rB23.y - nB23.y, y3 = iB23.y + jB23.y;
        pb[0] = fmaf(x0, x0, y0 * y0) * inv.x; pb[1] = fmaf(x1, x1, y1 * y1) * inv.y;
        pb[2] = fmaf(x2, x2, y2 * y2) * inv.z; pb[3] = fmaf(x3, x3, y3 * y3) * inv.w;
    }
    const bool liveA = t < T, liveB = t + 1 < T;
    if (liveA) {
        float va[4], vb[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            va[r] = decibels ? db10(fmaxf(1e-10f, pa[r])) : sqrtf(pa[r]);
            vb[r] = decibels ? db10(fmaxf(1e-10f, pb[r])) : sqrtf(pb[r]);
        }
        float *ot = out + t;
        const bool pair = vec2 && liveB;
        // rows shared by several harmonics are stored to each of them (one 16-byte descriptor load per destination)
        for (int d = 0; d < ndst; ++d) {
            if (!((live_d >> d) & 1u)) continue;
            const int4 off = *reinterpret_cast<const int4 *>(bl->off[d]);
            const int o4[4] = {off.x, off.y, off.z, off.w};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (o4[r] >= 0) {
                    float *q = ot + (long long)o4[r] * T;
                    if (pair) {
                        *reinterpret_cast<float2 *>(q) = make_float2(va[r], vb[r]);
                    } else {
                        q[0] = va[r];
                        if (liveB) q[1] = vb[r];
                    }
                }
            }
        }
    }
    if (decibels) {
        float vmax = liveA ? fmaxf(fmaxf(pa[0], pa[1]), fmaxf(pa[2], pa[3])) : 0.f;
        if (liveB) vmax = fmaxf(vmax, fmaxf(fmaxf(pb[0], pb[1]), fmaxf(pb[2], pb[3])));
        int have = 0x7fffffff;
        for (int d = 0; d < ndst; ++d)
            if ((live_d >> d) & 1u) have = min(have, s_max[bl->chan[d]]);
        if (__any_sync(gmask, __float_as_int(vmax) > have)) {
            for (int o = glanes / 2; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(gmask, vmax, o));
            if (leader)
                for (int d = 0; d < ndst; ++d)
                    if ((live_d >> d) & 1u) atomicMax(&s_max[bl->chan[d]], __float_as_int(vmax));
        }
        if (t + 1 >= T) {
            // frames past the trimmed output: they only count for the maximum of a harmonic whose own VQT is that long
            const float ma = fmaxf(fmaxf(pa[0], pa[1]), fmaxf(pa[2], pa[3])), mb = fmaxf(fmaxf(pb[0], pb[1]), fmaxf(pb[2], pb[3]));
            for (int d = 0; d < ndst; ++d) {
                if (!((live_d >> d) & 1u)) continue;
                const int lim = __ldg(tmax + bl->chan[d]);
                if (t >= T && t < lim) atomicMax(&s_max[bl->chan[d]], __float_as_int(ma));
                if (t + 1 >= T && t + 1 < lim) atomicMax(&s_max[bl->chan[d]], __float_as_int(mb));
            }
        }
    }
}

struct CqtParams {
    const float *audio, *ladder;
    float *out;
    const ClipMeta *meta;
    float *maxbuf;
    const CqtItem *items;
    const CqtBlock4 *blocks;
    const float4 *weights4;
    const CqtRow *rows;
    const float2 *weights;
    const float2 *tw1, *tw2;
    int C, F, decibels, tile_floats, stage_blocks, stage_rows, dbuf_off, w_off, blk_off, region_floats, tiles_per_cta;
    unsigned alt_mask;   // channels whose tail frames come from an exact ladder (Plan::alt_mask)
};

// Level signal an item frames: the shared ladder, or (exact-ladder item) the head / tail piece of an exact level (the tail
// through its virtual offset).  `len` is the length to zero-fill beyond: the level's, or the head piece's.
__device__ __forceinline__ const float *item_signal(const CqtItem &it, const ClipMeta *cm, const float *audio, const float *ladder,
                                                    bool head, long long &len) {
    len = cm->lvl_len[it.level];
    if (it.alt) {
        if (head) {
            len = cm->alt_hlen[it.alt - 1][it.level];
            return ladder + cm->alt_hoff[it.alt - 1][it.level];
        }
        return ladder + cm->alt_off[it.alt - 1][it.level];
    }
    return (it.level == 0 ? audio : ladder) + cm->lvl_off[it.level];
}

// Frames an item computes, as tiles of TT frames.  Shared-ladder item: tiles 0 .. over [0, Tall); its copy of exact-ladder rows
// skips the tiles inside [0, th) and [t0, Tall).  Exact-ladder item: `nhead` tiles over [0, min(th, Tall)), then the tiles from t0.
struct TileMap {
    int th, t0, nhead, ntiles, TT;
    bool alt;
    __device__ __forceinline__ TileMap(const CqtItem &it, const ClipMeta *cm, int Tall, int TT_) : TT(TT_), alt(it.alt != 0) {
        th = cm->alt_th[it.level];
        t0 = cm->alt_t0[it.level];
        if (alt) {
            nhead = (min(th, Tall) + TT - 1) / TT;
            ntiles = nhead + (t0 < Tall ? (Tall - t0 + TT - 1) / TT : 0);
        } else {
            nhead = 0;
            ntiles = (Tall + TT - 1) / TT;
        }
    }
    __device__ __forceinline__ bool head(int tile) const { return alt && tile < nhead; }
    __device__ __forceinline__ int first_frame(int tile) const { return (!alt || tile < nhead) ? tile * TT : t0 + (tile - nhead) * TT; }
    // frames of a head tile end at th (th is a multiple of every tile size); everything else runs to Tall
    __device__ __forceinline__ int frame_end(int tile, int Tall) const { return head(tile) ? min(th, Tall) : Tall; }
    __device__ __forceinline__ unsigned skip(int t, unsigned alt_mask) const { return (!alt && (t < th || t >= t0)) ? alt_mask : 0u; }
};

// Shared prologue of both CQT kernels: stage the level signal, run the warp FFT unit.
template <int NC>
__device__ __forceinline__ void cqt_fft_phase(const CqtParams &p, const CqtItem &it, const ClipMeta *cm, int t0, bool head, float2 *s_tw1,
                                              float2 *s_tw2, float *s_scr, float *s_tile) {
    using L = FftLayout<NC>;
    constexpr int G = L::G, TT = kWarpsPerCta * G, NFFT = 2 * NC, WP2 = L::WARP_PITCH / 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if !AMT_TW_GLOBAL
    for (int i = tid; i < NC; i += kThreads) s_tw1[i] = p.tw1[i];
    for (int i = tid; i <= NC; i += kThreads) s_tw2[i] = p.tw2[i];
#endif
    long long len;
    const float *src = item_signal(it, cm, p.audio, p.ladder, head, len);
    int shift, fstride;
    load_tile(s_tile, src, len, (long long)t0 * it.hop - NC, it.hop, NFFT, TT, tid, shift, fstride);
    __syncthreads();
    float2 *scr = reinterpret_cast<float2 *>(s_scr) + warp * WP2;
    const bool vec_ok = ((shift | fstride) & 1) == 0;
    const float2 *tw1 = AMT_TW_GLOBAL ? p.tw1 : s_tw1;
    if (vec_ok) {
        warp_fft_unit<NC, AMT_TW_GLOBAL != 0>(scr, tw1, lane, [&](int g, int n) {
            return *reinterpret_cast<const float2 *>(s_tile + shift + (warp * G + g) * fstride + 2 * n);
        });
    } else {
        warp_fft_unit<NC, AMT_TW_GLOBAL != 0>(scr, tw1, lane, [&](int g, int n) {
            const float *x = s_tile + shift + (warp * G + g) * fstride + 2 * n;
            return make_float2(x[0], x[1]);
        });
    }
}

// Main kernel (n_fft >= 128).  A CTA walks `tiles_per_cta` consecutive tiles of one (clip, item): the twiddles are staged
// once, and the audio of the next tile streams into its own shared-memory region (cp.async, zero fill) while the current
// tile is projected and stored.
// Phase A: per-warp FFT, then the real-FFT split restricted to the band [kmin, kmax] the item's rows touch; the band
//          is written TRANSPOSED into Dbuf[k - kmin][frame] (frame fastest, pitch TT + 1), which aliases the FFT scratch.
// Phase B: lanes run along frames, so every D fetch is a contiguous, conflict-free line; a group of FL lanes takes one
//          block of up to 4 rows and streams its warp-uniform weights ([step][4 rows], two 16-byte loads per step):
//          each D value feeds 4 complex MACs.  Results go through a staging tile (also inside the retired scratch) so
//          that the global stores are T-contiguous (16-byte vectors when the clip's rows are 16-byte aligned).
template <int NC, bool HALF>
__global__ void AMT_FFT_BOUNDS(AMT_CQT_MINB) cqt_kernel(const CqtParams p) {
    using L = FftLayout<NC>;
    constexpr int G = L::G, S = L::S, TT = kWarpsPerCta * G, WP2 = L::WARP_PITCH / 2, NFFT = 2 * NC;
    constexpr int FL = TT < 32 ? TT : 32;       // lanes along frames
    constexpr int NCHUNK = TT / FL;             // frame chunks per tile
    constexpr int DP = TT + kDbufPad;           // Dbuf pitch in float2
    extern __shared__ __align__(16) float smem[];
    float2 *s_tw1 = reinterpret_cast<float2 *>(smem);                 // NC
    float2 *s_tw2 = s_tw1 + NC;                                       // NC + 2 (k = 0 .. NC, padded)
    float *s_reg = reinterpret_cast<float *>(s_tw2 + NC + 2);         // FFT scratch; later Dbuf | weights | blocks
    float *s_tile = s_reg + p.region_floats;                          // audio tile (dedicated: prefetched during phase B)
    __shared__ int s_max[AMTFEAT_MAX_HARMONICS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ClipMeta *cm = p.meta + blockIdx.y;
    const int T = cm->T, Tall = cm->T_all;
    const CqtItem it = p.items[blockIdx.z];
    // an exact-ladder item covers the first and the last frames of its level; the shared copy of those rows skips them
    const TileMap tm(it, cm, Tall, TT);
    const int tile_begin = blockIdx.x * p.tiles_per_cta;
    if (tile_begin >= tm.ntiles) return;
    const int tile_end = min(tm.ntiles, tile_begin + p.tiles_per_cta);
    long long len;
    const float *src = item_signal(it, cm, p.audio, p.ladder, tm.head(tile_begin), len);
    const bool overlap = it.hop <= NFFT;
    if (tid < AMTFEAT_MAX_HARMONICS) s_max[tid] = 0;

    // prologue: twiddles and the first audio tile, all asynchronous (one exposed latency instead of three)
    {
        const float4 *g1 = reinterpret_cast<const float4 *>(p.tw1), *g2 = reinterpret_cast<const float4 *>(p.tw2);
        for (int i = tid; i < NC / 2; i += kThreads) {
            const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<float4 *>(s_tw1) + i));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(g1 + i) : "memory");
        }
        for (int i = tid; i < (NC + 2) / 2; i += kThreads) {   // the table is padded to NC + 2 entries on the host
            const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<float4 *>(s_tw2) + i));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(g2 + i) : "memory");
        }
    }
    int shift = 0, fstride = overlap ? it.hop : NFFT;
    if (overlap) load_tile_async(s_tile, src, len, (long long)tm.first_frame(tile_begin) * it.hop - NC, it.hop, NFFT, TT, tid, shift);
    asm volatile("cp.async.commit_group;\n" ::: "memory");

    // phase-B layout of the region
    float2 *Dbuf = reinterpret_cast<float2 *>(s_reg + p.dbuf_off);
    float4 *s_w = reinterpret_cast<float4 *>(s_reg + p.w_off);
    const CqtBlock4 *s_blk = reinterpret_cast<const CqtBlock4 *>(s_reg + p.blk_off);   // the item's block descriptors
    float *out = p.out + cm->out_off;
    const int kb = it.kmax - it.kmin + 1;

    for (int tile = tile_begin; tile < tile_end; ++tile) {
        const int t0 = tm.first_frame(tile);
        const unsigned skip = tm.skip(t0, p.alt_mask);
        if (overlap) {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::: "memory");
            __syncthreads();   // previous iteration's readers of the region are done
            src = item_signal(it, cm, p.audio, p.ladder, tm.head(tile), len);
            load_tile(s_tile, src, len, (long long)t0 * it.hop - NC, it.hop, NFFT, TT, tid, shift, fstride);
        }
        __syncthreads();       // tile (and, first time, the twiddles) visible; previous iteration's staging retired

        float2 *scr = reinterpret_cast<float2 *>(s_reg) + warp * WP2;
        if (AMT_DBG_SKIP & 2) {
        } else if (((shift | fstride) & 1) == 0) {
            warp_fft_unit<NC>(scr, s_tw1, lane, [&](int g, int n) {
                return *reinterpret_cast<const float2 *>(s_tile + shift + (warp * G + g) * fstride + 2 * n);
            });
        } else {
            warp_fft_unit<NC>(scr, s_tw1, lane, [&](int g, int n) {
                const float *x = s_tile + shift + (warp * G + g) * fstride + 2 * n;
                return make_float2(x[0], x[1]);
            });
        }

        // Real-FFT split on the band, held in registers across the barrier that retires the scratch, then written transposed
        // into Dbuf.  Only the first nj = ceil(kb / 32) register columns are touched (warp-uniform early exit).
        {
            // HALF: every item of the launch has a band of at most half the spectrum (the host checks) -> half the registers
            constexpr int JFULL = (NC + 1 + 31) / 32, JB = HALF ? (JFULL + 1) / 2 : JFULL;
            const int nj = (kb + 31) >> 5;
            float2 X[JB][G];
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                if (j >= nj) break;
                const int kk = lane + 32 * j;
                if (kk < kb) {
                    const int k = it.kmin + kk;
                    const float2 tw = s_tw2[k];
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const float2 A = scr[g * S + (k & (NC - 1))], B = scr[g * S + ((NC - k) & (NC - 1))];
                        float2 E, Tw;
                        rfft_split(A, B, tw, E, Tw);
                        X[j][g] = make_float2(E.x + Tw.x, E.y + Tw.y);
                    }
                }
            }
            __syncthreads();
            // stream the item's weights and block descriptors into the retired scratch (L1 bypass) while D is being written,
            // then the next tile's audio into the tile region (a second group: phase B only waits for the first)
            {
                const float4 *wsrc = p.weights4 + it.woff0;
                for (int i = tid; i < it.wcount; i += kThreads) {
                    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(s_w + i));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(wsrc + i) : "memory");
                }
                const float4 *bsrc = reinterpret_cast<const float4 *>(p.blocks + it.blk0);
                const int n16 = it.nblk * (int)(sizeof(CqtBlock4) / 16);
                for (int i = tid; i < n16; i += kThreads) {
                    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(reinterpret_cast<const float4 *>(s_blk) + i));
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(bsrc + i) : "memory");
                }
                asm volatile("cp.async.commit_group;\n" ::: "memory");
                if (overlap && tile + 1 < tile_end) {
                    src = item_signal(it, cm, p.audio, p.ladder, tm.head(tile + 1), len);
                    load_tile_async(s_tile, src, len, (long long)tm.first_frame(tile + 1) * it.hop - NC, it.hop, NFFT, TT, tid, shift);
                }
                asm volatile("cp.async.commit_group;\n" ::: "memory");
            }
#pragma unroll
            for (int j = 0; j < JB; ++j) {
                if (j >= nj) break;
                const int kk = lane + 32 * j;
                if (kk < kb) {
#pragma unroll
                    for (int g = 0; g < G; ++g) Dbuf[kk * DP + warp * G + g] = X[j][g];
                }
            }
            asm volatile("cp.async.wait_group 1;\n" ::: "memory");
        }
        __syncthreads();

        // Projection.  The lanes of a group hold consecutive frames, so the stores of one row are already T-contiguous
        // (FL * 4 bytes per row and group): results go straight to global memory.
        {
            constexpr int LPB = FL / 2;              // lanes per block: every lane owns two consecutive frames
            const int sub2 = tid / LPB, lt2 = tid % LPB;
            const unsigned gm2 = LPB == 32 ? 0xffffffffu : (((1u << LPB) - 1u) << ((lane / LPB) * LPB));
            const bool vec2 = ((T & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
            for (int w = sub2; w < it.nblk * NCHUNK; w += kThreads / LPB) {
                const int bi = w / NCHUNK, ch = w % NCHUNK;
                const CqtBlock4 *bl = s_blk + bi;
                project_block2(bl, s_w + (bl->woff - it.woff0), Dbuf + (bl->col0 - it.kmin) * DP + ch * FL + 2 * lt2, DP, out, T,
                               t0 + ch * FL + 2 * lt2, p.decibels, s_max, gm2, LPB, lt2 == 0, vec2, skip, cm->t_max);
            }
        }
        // no barrier here: the next iteration's top-of-loop barrier orders these reads before the next FFT's scratch writes
    }
    __syncthreads();
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (p.decibels && tid < p.C && s_max[tid] != 0) atomicMax(reinterpret_cast<int *>(p.maxbuf) + blockIdx.y * p.C + tid, s_max[tid]);
}

// ------------------------------------------------------------------------------------------------
// K1 + K5, deep ladder levels : sliding DFT (cqt_slide_kernel)
//
// On the deep levels of the ladder hop << n_fft (HCQT: hop 16 ... 2 against n_fft 1024), so consecutive rectangular frames
// share all but `hop` samples and one FFT per frame recomputes almost everything.  With W = exp(-2 pi i / N) and
// P_t[k] = W^(k t hop) the frame spectrum is  X_t[k] = conj(P_t[k]) * B_t[k],
//     B_t[k] = sum over the frame's samples x[j] W^(k (j + N/2)),    (absolute phase: W^(k N) = 1, so a sample enters
//     B_(t+1)[k] = B_t[k] + P_t[k] * sum_{m < hop} d_t[m] W^(k m),    and leaves with the SAME factor)
//     d_t[m] = x[t hop + N/2 + m] - x[t hop - N/2 + m].
// One thread owns one bin k of the item's band and walks the frames of a tile: 2 hop real-by-complex MACs with
// register-resident twiddles + a handful of complex operations per frame instead of an N-point FFT shared by the band
// (n_fft 1024, hop 8, 207 band bins: ~36 instructions per bin and frame).  Numerics: B is only ever ADDED to (no
// multiplicative state), with Kahan compensation; P is re-seeded from the exact table every 8 frames; a tile starts
// from B = 0 and a lead-in of N / hop sample groups (the first window), so errors never outlive a tile.  Every 32 frames
// the band spectra are handed to the same blocked projection as cqt_kernel through Dbuf[k][frame].
// ------------------------------------------------------------------------------------------------

constexpr int kSlideMaxItems = 24;          // items per launch (grid.z)
#ifndef AMT_SLIDE_FL
#define AMT_SLIDE_FL 32        // frames per projection chunk
#endif
#ifndef AMT_SLIDE_BUFS
#define AMT_SLIDE_BUFS 1       // Dbuf copies: 2 = chunks alternate between two buffers and need ONE barrier each instead of two.  Measured
                               // (c5): 16-frame chunks x 2 copies (same shared memory) 1.31 ms against 1.19 ms -- the shorter projection
                               // chunks cost more than the barrier saves; 32-frame chunks x 2 copies no longer fit two CTAs per SM
#endif
#ifndef AMT_SLIDE_CTAS
#define AMT_SLIDE_CTAS 2
#endif
constexpr int kSlideFL = AMT_SLIDE_FL;      // frames per projection chunk
constexpr int kSlideDP = kSlideFL + kDbufPad;   // Dbuf pitch (float2)
constexpr int kSlideBufs = AMT_SLIDE_BUFS;

struct SlideParams {
    const float *audio, *ladder;
    float *out;
    const ClipMeta *meta;
    float *maxbuf;
    const CqtItem *items;
    const CqtBlock4 *blocks;
    const float4 *weights4;
    int C, decibels, w_off, blk_off, x_off;
    unsigned alt_mask;
    int idx[kSlideMaxItems];
    const float2 *tw2[kSlideMaxItems];
};

__host__ __device__ inline int slide_tile_frames(int hop) { return AMT_SLIDE_TILE / hop < 1024 ? AMT_SLIDE_TILE / hop : 1024; }

// W_N^e from the half table tw2[k] = exp(-i pi k / NC), k = 0 .. NC (N = 2 NC)
__device__ __forceinline__ float2 tw_full(const float2 *__restrict__ tw2, int e, int NC) {
    e &= 2 * NC - 1;
    if (e <= NC) return __ldg(tw2 + e);
    const float2 w = __ldg(tw2 + 2 * NC - e);
    return make_float2(w.x, -w.y);
}

template <int H>
__device__ __forceinline__ void slide_run(const SlideParams &p, const CqtItem &it, const float2 *__restrict__ tw2,
                                          const ClipMeta *cm, int t0, int tend, int Tt, float *s_reg, int *s_max) {
    constexpr int FL = kSlideFL, DP = kSlideDP, HP = (H + 1) / 2;
    // An error of P scales the whole increment of a frame (hop samples of the FULL signal), so it matters most where the hop is
    // large: measured on the HCQT, hop 16 / 8 with a float32 recurrence re-seeded every 8 frames were the least accurate items of
    // the plan (9e-4 dB on bins 60 dB down).  There the recurrence runs in float64 on the otherwise idle FP64 pipe (4 operations
    // + 2 conversions per frame against 4 hop float32 flops): exact to 1e-16 per step, never re-seeded, rounded once per frame.
    constexpr bool PD = H >= 8;
    constexpr int RESEED = AMT_SLIDE_RESEED;
    const int tid = threadIdx.x, NT = blockDim.x;
    const int N = it.nfft, NC = N >> 1, Q = N / H;
    const int kb = it.kmax - it.kmin + 1;
    const bool active = tid < kb;
    const int k = it.kmin + tid;
    const int T = cm->T;
    // the shared copy of exact-ladder rows skips their first and last frames (chunk-uniform: both bounds are multiples of 32)
    const int sk_th = it.alt ? 0 : cm->alt_th[it.level], sk_t0 = it.alt ? 0x7fffffff : cm->alt_t0[it.level];
    float2 *Dbuf = reinterpret_cast<float2 *>(s_reg);
    const float4 *s_w = reinterpret_cast<const float4 *>(s_reg + p.w_off);
    const CqtBlock4 *s_blk = reinterpret_cast<const CqtBlock4 *>(s_reg + p.blk_off);
    float *s_x = s_reg + p.x_off;

    // W^(k m), m < H, as (re, re) / (im, im) pairs of consecutive samples: one packed FFMA2 per pair and component
    float2 wre[HP], wim[HP];
#pragma unroll
    for (int q = 0; q < HP; ++q) {
        const float2 w0 = tw_full(tw2, k * (2 * q), NC);
        const float2 w1 = (2 * q + 1 < H) ? tw_full(tw2, k * (2 * q + 1), NC) : make_float2(0.f, 0.f);
        wre[q] = make_float2(w0.x, w1.x);
        wim[q] = make_float2(w0.y, w1.y);
    }
    const float2 R = tw_full(tw2, k * H, NC);
    float2 B = make_float2(0.f, 0.f), cmp = make_float2(0.f, 0.f), P = make_float2(1.f, 0.f);
    auto seed = [&](int t) { return tw_full(tw2, k * ((t * H) & (N - 1)), NC); };
    double2 Pd = make_double2(1.0, 0.0), Rd = make_double2(1.0, 0.0);
    if (PD) {
        sincospi(-2.0 * (double)((k * H) & (N - 1)) / (double)N, &Rd.y, &Rd.x);
        sincospi(-2.0 * (double)((k * (((t0 - Q) * H) & (N - 1))) & (N - 1)) / (double)N, &Pd.y, &Pd.x);
        P = make_float2((float)Pd.x, (float)Pd.y);
    }
    // B += P * sum_m xs[m] W^(k m);  P *= W^(k H)
    auto step = [&](const float *xs) {
        float2 are = make_float2(0.f, 0.f), aim = make_float2(0.f, 0.f);
        if (H >= 4) {
#pragma unroll
            for (int q4 = 0; q4 < H / 4; ++q4) {
                const float4 v = *reinterpret_cast<const float4 *>(xs + 4 * q4);
                are = ffma2(make_float2(v.x, v.y), wre[2 * q4], are);
                aim = ffma2(make_float2(v.x, v.y), wim[2 * q4], aim);
                are = ffma2(make_float2(v.z, v.w), wre[2 * q4 + 1], are);
                aim = ffma2(make_float2(v.z, v.w), wim[2 * q4 + 1], aim);
            }
        } else if (H == 2) {
            const float2 v = *reinterpret_cast<const float2 *>(xs);
            are = ffma2(v, wre[0], are);
            aim = ffma2(v, wim[0], aim);
        } else {
            are.x = xs[0] * wre[0].x;
            aim.x = xs[0] * wim[0].x;
        }
        const float dx = are.x + are.y, dy = aim.x + aim.y;
        const float tx = fmaf(P.x, dx, -P.y * dy), ty = fmaf(P.x, dy, P.y * dx);
        const float yx = tx - cmp.x, yy = ty - cmp.y;          // Kahan-compensated B += (tx, ty)
        const float nx = B.x + yx, ny = B.y + yy;
        cmp.x = (nx - B.x) - yx;
        cmp.y = (ny - B.y) - yy;
        B.x = nx;
        B.y = ny;
        if (PD) {
            Pd = make_double2(fma(Pd.x, Rd.x, -Pd.y * Rd.y), fma(Pd.x, Rd.y, Pd.y * Rd.x));
            P = make_float2((float)Pd.x, (float)Pd.y);
        } else {
            P = cmul(P, R);
        }
    };

    // lead-in: the first window of the tile enters sample group by sample group
    if (active) {
#pragma unroll 2
        for (int u = 0; u < Q; ++u) {
            if (!PD && (u & (RESEED - 1)) == 0) P = seed(t0 - Q + u);
            step(s_x + u * H);
        }
    }
    __syncthreads();
    // in place x[i] <- x[i + N] - x[i]  (entering minus leaving sample), one residue class mod N per thread
    for (int r = tid; r < N; r += NT) {
        float prev = s_x[r];
        for (int i = r; i < Tt * H; i += N) {
            const float nxt = s_x[i + N];
            s_x[i] = nxt - prev;
            prev = nxt;
        }
    }
    __syncthreads();

    float *out = p.out + cm->out_off;
    // Chunks alternate between kSlideBufs copies of Dbuf (each kb x DP): with two, the sliding phase of chunk i + 1 writes the copy
    // the projection of chunk i is NOT reading, so the only barrier a chunk needs is the one between its own two phases -- a
    // thread passes it only after finishing the projection of the previous chunk, which orders that projection before the
    // sliding phase (one chunk later) that overwrites its copy.  A slow warp of one phase is absorbed by the other.
    const int buf_stride = kb * DP;
    int chunk = 0;
    for (int c0 = 0; c0 < Tt && t0 + c0 < tend; c0 += FL, ++chunk) {
        const unsigned skip = (t0 + c0 < sk_th || t0 + c0 >= sk_t0) ? p.alt_mask : 0u;
        float2 *Dcur = Dbuf + (kSlideBufs > 1 ? (chunk & 1) * buf_stride : 0);
        if (active) {
            float2 *dp = Dcur + tid * DP;
#pragma unroll 4
            for (int f = 0; f < FL; ++f) {
                if (!PD && (f & (RESEED - 1)) == 0) P = seed(t0 + c0 + f);
                dp[f] = make_float2(fmaf(P.x, B.x, P.y * B.y), fmaf(P.x, B.y, -P.y * B.x));   // conj(P) * B
                step(s_x + (c0 + f) * H);
            }
        }
        __syncthreads();
        {
            constexpr int LPB = FL / 2;
            const int sub2 = tid / LPB, lt2 = tid % LPB;
            const unsigned gm2 = LPB == 32 ? 0xffffffffu : (((1u << LPB) - 1u) << (((tid & 31) / LPB) * LPB));
            const bool vec2 = ((T & 1) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
            for (int w = sub2; w < it.nblk; w += NT / LPB) {
                const CqtBlock4 *bl = s_blk + w;
                project_block2(bl, s_w + (bl->woff - it.woff0), Dcur + (bl->col0 - it.kmin) * DP + 2 * lt2, DP, out, T, t0 + c0 + 2 * lt2,
                               p.decibels, s_max, gm2, LPB, lt2 == 0, vec2, skip, cm->t_max);
            }
        }
        if (kSlideBufs == 1) __syncthreads();
    }
}

__global__ void AMT_FFT_BOUNDS(AMT_SLIDE_CTAS) cqt_slide_kernel(const SlideParams p) {
    extern __shared__ __align__(16) float smem[];
    __shared__ int s_max[AMTFEAT_MAX_HARMONICS];
    const int tid = threadIdx.x, NT = blockDim.x;
    const CqtItem it = p.items[p.idx[blockIdx.z]];
    const float2 *tw2 = p.tw2[blockIdx.z];
    const ClipMeta *cm = p.meta + blockIdx.y;
    const int Tall = cm->T_all, Tt = slide_tile_frames(it.hop);
    // tiles of Tt frames; an exact-ladder item covers the first and the last frames of its level only (TileMap)
    const TileMap tm(it, cm, Tall, Tt);
    if ((int)blockIdx.x >= tm.ntiles) return;
    const bool head = tm.head(blockIdx.x);
    const int t0 = tm.first_frame(blockIdx.x), tend = tm.frame_end(blockIdx.x, Tall);
    if (tid < AMTFEAT_MAX_HARMONICS) s_max[tid] = 0;
    {
        // the tile's samples [t0 hop - N/2, (t0 + Tt) hop + N/2), zero outside the level signal (16-byte cp.async with zero
        // fill: the start is a multiple of 4 samples), and the item's weights and block descriptors
        long long len;
        const float *src = item_signal(it, cm, p.audio, p.ladder, head, len);
        const long long j0 = (long long)t0 * it.hop - it.nfft / 2;
        const int nvec = (Tt * it.hop + it.nfft) >> 2;
        float *s_x = smem + p.x_off;
        for (int i = tid; i < nvec; i += NT) {
            const long long g = j0 + 4ll * i;
            int valid = 0;
            if (g >= 0 && g < len) valid = (int)(len - g < 4 ? len - g : 4) * 4;
            const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(s_x + 4 * i));
            const float *gp = src + (valid ? g : 0);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(dst), "l"(gp), "r"(valid) : "memory");
        }
        float4 *s_w = reinterpret_cast<float4 *>(smem + p.w_off);
        const float4 *wsrc = p.weights4 + it.woff0;
        for (int i = tid; i < it.wcount; i += NT) {
            const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(s_w + i));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(wsrc + i) : "memory");
        }
        float4 *s_b = reinterpret_cast<float4 *>(smem + p.blk_off);
        const float4 *bsrc = reinterpret_cast<const float4 *>(p.blocks + it.blk0);
        const int n16 = it.nblk * (int)(sizeof(CqtBlock4) / 16);
        for (int i = tid; i < n16; i += NT) {
            const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(s_b + i));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(bsrc + i) : "memory");
        }
        asm volatile("cp.async.commit_group;\n" ::: "memory");
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    }
    __syncthreads();
    switch (it.hop) {
        case 16: slide_run<16>(p, it, tw2, cm, t0, tend, Tt, smem, s_max); break;
        case 8: slide_run<8>(p, it, tw2, cm, t0, tend, Tt, smem, s_max); break;
        case 4: slide_run<4>(p, it, tw2, cm, t0, tend, Tt, smem, s_max); break;
        case 2: slide_run<2>(p, it, tw2, cm, t0, tend, Tt, smem, s_max); break;
        default: slide_run<1>(p, it, tw2, cm, t0, tend, Tt, smem, s_max); break;
    }
    __syncthreads();   // the last chunk's projection (its maxima in s_max) is complete
    if (p.decibels && tid < p.C && s_max[tid] != 0) atomicMax(reinterpret_cast<int *>(p.maxbuf) + blockIdx.y * p.C + tid, s_max[tid]);
}

// Fallback for tiny transforms (n_fft <= 64: the lowest octaves of a VQT with a large gamma): one thread per
// (row, chunk of 8 frames) on the per-row tables.  Negligible share of any workload.
template <int NC>
__global__ void __launch_bounds__(kThreads, 2) cqt_small_kernel(const CqtParams p) {
    using L = FftLayout<NC>;
    constexpr int G = L::G, S = L::S, TT = kWarpsPerCta * G, WP2 = L::WARP_PITCH / 2;
    extern __shared__ __align__(16) float smem[];
    float2 *s_tw1 = reinterpret_cast<float2 *>(smem);
    float2 *s_tw2 = s_tw1 + NC;
    float *s_scr = reinterpret_cast<float *>(s_tw2 + NC + 2);
    float *s_tile = s_scr + kWarpsPerCta * L::WARP_PITCH;
    float *s_stage = s_tile + p.tile_floats;                          // stage_rows * (TT + 1)
    __shared__ int s_max[AMTFEAT_MAX_HARMONICS];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const ClipMeta *cm = p.meta + blockIdx.y;
    const int T = cm->T, Tall = cm->T_all;
    const CqtItem it = p.items[blockIdx.z];
    const TileMap tm(it, cm, Tall, TT);
    if ((int)blockIdx.x >= tm.ntiles) return;
    const int t0 = tm.first_frame(blockIdx.x), tend = tm.frame_end(blockIdx.x, Tall);
    if (tid < AMTFEAT_MAX_HARMONICS) s_max[tid] = 0;
    cqt_fft_phase<NC>(p, it, cm, t0, tm.head(blockIdx.x), s_tw1, s_tw2, s_scr, s_tile);

    float2 *scr = reinterpret_cast<float2 *>(s_scr) + warp * WP2;
    const int kb = it.kmax_true - it.kmin + 1;
    {
        constexpr int JMAX = (G * (NC + 1) + 31) / 32;
        float2 X[JMAX];
        const int nband = G * kb;
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const int q = lane + 32 * j;
            if (q < nband) {
                const int g = q / kb, k = it.kmin + (q - g * kb);
                const float2 A = scr[g * S + (k & (NC - 1))], B = scr[g * S + ((NC - k) & (NC - 1))];
                float2 E, Tw;
                rfft_split(A, B, AMT_TW_GLOBAL ? __ldg(p.tw2 + k) : s_tw2[k], E, Tw);
                X[j] = make_float2(E.x + Tw.x, E.y + Tw.y);
            }
        }
        __syncwarp();
#pragma unroll
        for (int j = 0; j < JMAX; ++j) {
            const int q = lane + 32 * j;
            if (q < nband) {
                const int g = q / kb;
                scr[g * S + (q - g * kb)] = X[j];
            }
        }
    }
    __syncthreads();

    const float2 *D = reinterpret_cast<const float2 *>(s_scr);
    float *out = p.out + cm->out_off;
    for (int rc = 0; rc < it.nrows; rc += p.stage_rows) {
        const int nr = min(p.stage_rows, it.nrows - rc);
        for (int w = tid; w < nr * G; w += kThreads) {
            const int rl = w % nr, ch = w / nr;
            const CqtRow row = p.rows[it.row0 + rc + rl];
            const float2 *wts = p.weights + row.woff;
            const int cbase = row.col0 - it.kmin;
            float2 acc[8];
#pragma unroll
            for (int f = 0; f < 8; ++f) acc[f] = make_float2(0.f, 0.f);
            for (int j = 0; j < row.cnt; ++j) {
                const float2 wv = __ldg(wts + j);
#pragma unroll
                for (int f = 0; f < 8; ++f) {
                    const int t = ch * 8 + f;
                    const float2 d = D[(t / G) * WP2 + (t % G) * S + cbase + j];
                    acc[f].x = fmaf(wv.x, d.x, acc[f].x);
                    acc[f].x = fmaf(-wv.y, d.y, acc[f].x);
                    acc[f].y = fmaf(wv.x, d.y, acc[f].y);
                    acc[f].y = fmaf(wv.y, d.x, acc[f].y);
                }
            }
            float vmax = 0.f;
            const bool alt_row = ((p.alt_mask >> row.chan) & 1u) != 0;
            const int tlim = min(tend, p.decibels ? cm->t_max[row.chan] : T);
#pragma unroll
            for (int f = 0; f < 8; ++f) {
                const float pw = fmaf(acc[f].x, acc[f].x, acc[f].y * acc[f].y) * row.inv_len;
                const int t = t0 + ch * 8 + f;
                if (t < tlim && !(alt_row && tm.skip(t, 1u))) vmax = fmaxf(vmax, pw);
                s_stage[rl * (TT + 1) + ch * 8 + f] = p.decibels ? db10(fmaxf(1e-10f, pw)) : sqrtf(pw);
            }
            if (p.decibels) atomicMax(&s_max[row.chan], __float_as_int(vmax));
        }
        __syncthreads();
        for (int idx = tid; idx < nr * TT; idx += kThreads) {
            const int t = idx % TT, rl = idx / TT;
            if (t0 + t < min(T, tend)) {
                const CqtRow *row = p.rows + it.row0 + rc + rl;
                const int chan = __ldg(&row->chan);
                if (!(((p.alt_mask >> chan) & 1u) && tm.skip(t0 + t, 1u)))
                    out[((long long)chan * p.F + __ldg(&row->bin)) * T + t0 + t] = s_stage[rl * (TT + 1) + t];
            }
        }
        __syncthreads();
    }
    if (p.decibels && tid < p.C && s_max[tid] != 0) atomicMax(reinterpret_cast<int *>(p.maxbuf) + blockIdx.y * p.C + tid, s_max[tid]);
}

// Per-call clip descriptors: pinned host ring slot -> workspace, read over the bus by a few small CTAs.  A kernel, not a
// cudaMemcpyAsync: a copy would queue on the host-to-device copy engine BEHIND the caller's bulk audio upload of the next
// batch (measured: the device-consumer loop ran at upload + compute instead of max(upload, compute)).
__global__ void __launch_bounds__(kThreads) copy_meta_kernel(const int4 *__restrict__ src, int4 *__restrict__ dst, int n16) {
    const int i = blockIdx.x * kThreads + threadIdx.x;     // one 16-byte word per thread: every bus read of the copy is in flight at once
    if (i < n16) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// host: upload, workspace layout, launch
// ------------------------------------------------------------------------------------------------

static bool is_vqt_kind_cfg(int kind) { return kind == AMTFEAT_VQT || kind == AMTFEAT_HVQT; }

template <typename Tp> static int upload_vec(Plan &p, const std::vector<Tp> &h, Tp **d) {
    *d = nullptr;
    if (h.empty()) return AMTFEAT_OK;
    AMT_CUDA(cudaMalloc(reinterpret_cast<void **>(d), h.size() * sizeof(Tp)));
    p.d_allocs.push_back(*d);
    AMT_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(Tp), cudaMemcpyHostToDevice));
    return AMTFEAT_OK;
}

template <int NC> static size_t stft_scr_floats(int n_mels, bool mel) {
    using L = FftLayout<NC>;
    const size_t PT = kWarpsPerCta * L::G + 1;
    size_t region = (size_t)kWarpsPerCta * L::WARP_PITCH;                                  // FFT scratch ...
    region = std::max(region, (size_t)(NC + 1) * PT + (mel ? 2 * (size_t)(n_mels + 1) * PT : 0));  // ... reused as Pbuf + mel staging
    return (region + 3) / 4 * 4;
}
template <int NC> static size_t cqt_smem(int tile_floats, int stage_rows) {
    using L = FftLayout<NC>;
    size_t fl = 2 * NC + 2 * (NC + 2) + (size_t)kWarpsPerCta * L::WARP_PITCH + tile_floats +
                (size_t)stage_rows * (kWarpsPerCta * L::G + 1) + stage_rows;
    return fl * sizeof(float);
}
template <int NC> static constexpr bool cqt_use_blocks() { return NC >= 64; }
template <int NC> static cudaError_t cqt_set_attr() {
    if (cqt_use_blocks<NC>()) {
        cudaError_t e = cudaFuncSetAttribute(cqt_kernel<(NC >= 64 ? NC : 64), false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return e;
        return cudaFuncSetAttribute(cqt_kernel<(NC >= 64 ? NC : 64), true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    }
    return cudaFuncSetAttribute(cqt_small_kernel<(NC < 64 ? NC : 32)>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
}
static int tile_floats_for(int TT, int hop, int nfft) {
    int fl = hop <= nfft ? (TT - 1) * hop + nfft + 8 : TT * nfft;
    return (fl + 3) / 4 * 4;
}

template <int NC> static int set_attrs() {
    AMT_CUDA(cudaFuncSetAttribute(stft_kernel<NC, MODE_STFT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AMT_CUDA(cudaFuncSetAttribute(stft_kernel<NC, MODE_MEL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    AMT_CUDA(cqt_set_attr<NC>());
    return AMTFEAT_OK;
}

int upload_plan(Plan &p) {
    DeviceGuard guard(p.device);
    AMT_CUDA(guard.status);   // reports an invalid ordinal
    if (is_vqt_kind_cfg(p.cfg.kind)) {
        // highest priority: the ladder's few CTAs take the next free SM slots instead of queueing behind a projection grid
        int prio_lo = 0, prio_hi = 0;
        AMT_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        for (int sl = 0; sl < Plan::kCallSlots; ++sl) {
            cudaStream_t side;
            AMT_CUDA(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_hi));
            p.side_stream[sl] = side;
            cudaStream_t tail;
            AMT_CUDA(cudaStreamCreateWithPriority(&tail, cudaStreamNonBlocking, prio_hi));
            p.tail_stream[sl] = tail;
            for (void *&e : p.call_events[sl]) AMT_CUDA(cudaEventCreateWithFlags(reinterpret_cast<cudaEvent_t *>(&e), cudaEventDisableTiming));
        }
    }
    int rc;
    if ((rc = set_attrs<1024>()) || (rc = set_attrs<512>()) || (rc = set_attrs<256>()) || (rc = set_attrs<128>()) ||
        (rc = set_attrs<64>()) || (rc = set_attrs<32>()) || (rc = set_attrs<16>()) || (rc = set_attrs<8>()) ||
        (rc = set_attrs<4>()))
        return rc;
    AMT_CUDA(cudaFuncSetAttribute(decimate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    AMT_CUDA(cudaFuncSetAttribute(decimate_fft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
    AMT_CUDA(cudaFuncSetAttribute(decimate_fft64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    AMT_CUDA((cudaFuncSetAttribute(tail_decimate_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    AMT_CUDA((cudaFuncSetAttribute(tail_decimate_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    AMT_CUDA((cudaFuncSetAttribute(tail_decimate_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    AMT_CUDA((cudaFuncSetAttribute(tail_decimate_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)));
    AMT_CUDA(cudaFuncSetAttribute(cqt_slide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    if ((rc = upload_vec(p, p.window, &p.d_window))) return rc;
    if ((rc = upload_vec(p, p.mel_start, &p.d_mel_start))) return rc;
    if ((rc = upload_vec(p, p.mel_cnt, &p.d_mel_cnt))) return rc;
    if ((rc = upload_vec(p, p.mel_off, &p.d_mel_off))) return rc;
    if ((rc = upload_vec(p, p.mel_w, &p.d_mel_w))) return rc;
    if ((rc = upload_vec(p, p.taps, &p.d_taps))) return rc;
    if ((rc = upload_vec(p, p.taps64, &p.d_taps64))) return rc;
    if ((rc = upload_vec(p, p.decim_hh, &p.d_decim_hh))) return rc;
    if ((rc = upload_vec(p, p.decim_h64, &p.d_decim_h64))) return rc;
    if ((rc = upload_vec(p, p.decim_tw64, &p.d_decim_tw64))) return rc;
    for (AltLadder &al : p.alts)
        if ((rc = upload_vec(p, al.taps, &al.d_taps))) return rc;
    if ((rc = upload_vec(p, p.rows, &p.d_rows))) return rc;
    if ((rc = upload_vec(p, p.weights, &p.d_weights))) return rc;
    if ((rc = upload_vec(p, p.blocks, &p.d_blocks))) return rc;
    if ((rc = upload_vec(p, p.weights4, &p.d_weights4))) return rc;
    if ((rc = upload_vec(p, p.mel_ww, &p.d_mel_ww))) return rc;
    if ((rc = upload_vec(p, p.mel_seg_start, &p.d_mel_seg_start))) return rc;
    if ((rc = upload_vec(p, p.mel_gsteps, &p.d_mel_gsteps))) return rc;
    if ((rc = upload_vec(p, p.mel_goff, &p.d_mel_goff))) return rc;
    if ((rc = upload_vec(p, p.items, &p.d_items))) return rc;
    for (auto &kv : p.fft) {
        if ((rc = upload_vec(p, kv.second.tw1, &kv.second.d_tw1))) return rc;
        if ((rc = upload_vec(p, kv.second.tw2, &kv.second.d_tw2))) return rc;
    }
    return AMTFEAT_OK;
}

// Sums the recorded event pairs per kernel name into `json` and clears the records.
int profile_read(const Plan &p, std::string &json) {
    DeviceGuard guard(p.device);
    std::lock_guard<std::mutex> lock(p.call_mu);
    std::map<std::string, std::pair<double, int>> acc;
    int rc = AMTFEAT_OK;
    for (auto &r : p.prof) {
        cudaEvent_t e0 = (cudaEvent_t)r.e0, e1 = (cudaEvent_t)r.e1;
        float ms = 0.f;
        if (cudaEventSynchronize(e1) != cudaSuccess || cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) {
            set_error("profile_read: an event pair could not be read");
            rc = AMTFEAT_ERR_CUDA;
        }
        acc[r.name].first += ms;
        acc[r.name].second += 1;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    p.prof.clear();
    json = "{";
    bool first = true;
    for (auto &kv : acc) {
        char b[256];
        snprintf(b, sizeof b, "%s\"%s\": {\"ms\": %.6f, \"launches\": %d}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
        json += b;
        first = false;
    }
    json += "}";
    return rc;
}

void free_plan_device(Plan &p) {
    if (p.device < 0) return;
    DeviceGuard guard(p.device);
    for (int sl = 0; sl < Plan::kCallSlots; ++sl) {
        for (void **sp : {&p.side_stream[sl], &p.tail_stream[sl]}) {
            if (*sp) {
                cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(*sp));
                cudaStreamDestroy(reinterpret_cast<cudaStream_t>(*sp));
                *sp = nullptr;
            }
        }
        for (void *&e : p.call_events[sl])
            if (e) { cudaEventDestroy(reinterpret_cast<cudaEvent_t>(e)); e = nullptr; }
    }
    for (auto &r : p.prof) {
        cudaEventDestroy((cudaEvent_t)r.e0);
        cudaEventDestroy((cudaEvent_t)r.e1);
    }
    p.prof.clear();
    if (p.meta_ring) {
        for (void *&e : p.meta_events)
            if (e) { cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(e)); cudaEventDestroy(reinterpret_cast<cudaEvent_t>(e)); e = nullptr; }
        cudaFreeHost(p.meta_ring);
        p.meta_ring = nullptr;
        p.meta_slot_bytes = 0;
    }
    for (void *d : p.d_allocs) cudaFree(d);
    p.d_allocs.clear();
}

// Event pair around one launch (amtfeat_profile_*): recorded under Plan::call_mu, like everything process() enqueues.
struct ProfScope {
    const Plan &p;
    cudaStream_t st;
    cudaEvent_t e1 = nullptr;
    ProfScope(const Plan &plan, const char *name, cudaStream_t s) : p(plan), st(s) {
        if (!p.prof_enabled) return;
        cudaEvent_t e0;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        p.prof.push_back({name, e0, e1});
    }
    ~ProfScope() {
        if (e1) cudaEventRecord(e1, st);
    }
};

static bool is_vqt_kind(const Plan &p) { return p.cfg.kind == AMTFEAT_VQT || p.cfg.kind == AMTFEAT_HVQT; }
bool has_db_epilogue(const Plan &p) { return p.cfg.decibels && p.cfg.kind != AMTFEAT_WAVEFORM; }

// Frames of one clip: T stored, T_all computed (the dB maximum of a harmonic runs over its own, untrimmed VQT: hvqt.py:123-128
// converts every harmonic to dB before trimming it to the common frame count).
static bool clip_frames(const Plan &p, int64_t n, ClipMeta &cm) {
    const int64_t T = output_frames(p, n);
    if (T < 0) return false;
    cm.T = cm.T_all = (int32_t)T;
    for (int c = 0; c < AMTFEAT_MAX_HARMONICS; ++c) cm.t_max[c] = (int32_t)T;
    if (p.cfg.kind == AMTFEAT_HVQT && p.cfg.decibels && T > 0) {
        for (size_t h = 0; h < p.harm.size(); ++h) {
            cm.t_max[h] = (int32_t)std::max<int64_t>(T, harmonic_frames(p, (int)h, n));
            cm.T_all = std::max(cm.T_all, cm.t_max[h]);
        }
    }
    return true;
}

// workspace: [ClipMeta x B][maxbuf float x B*C][ladder levels 1..n_levels-1][tails of the exact ladders]
struct WsLayout {
    size_t meta_off = 0, max_off = 0, ladder_off = 0, total = 0;
    bool ok = true;
};
static WsLayout ws_layout(const Plan &p, int batch, const int64_t *n, std::vector<ClipMeta> *metas) {
    WsLayout w;
    w.meta_off = 0;
    w.max_off = align_up((size_t)batch * sizeof(ClipMeta), 256);
    w.ladder_off = align_up(w.max_off + (size_t)batch * p.C * sizeof(float), 256);
    size_t ladder_elems = 0;
    std::vector<ClipMeta> local;
    std::vector<ClipMeta> &ms = metas ? *metas : local;
    ms.assign(batch, ClipMeta{});
    for (int b = 0; b < batch; ++b) {
        if (!clip_frames(p, n[b], ms[b])) w.ok = false;
        for (int l = 0; l < kMaxLevels; ++l) {
            ms[b].alt_t0[l] = INT32_MAX;
            ms[b].alt_th[l] = 0;
            for (int a = 0; a < kMaxAlt; ++a) { ms[b].alt_first[a][l] = -1; ms[b].alt_hlen[a][l] = 0; }
        }
    }
    if (is_vqt_kind(p)) {
        for (int b = 0; b < batch; ++b) level_lengths(p, n[b], ms[b].lvl_len);
        for (int l = 1; l < p.n_levels; ++l)
            for (int b = 0; b < batch; ++b) {
                ms[b].lvl_off[l] = (int64_t)ladder_elems;
                ladder_elems += align_up((size_t)ms[b].lvl_len[l], 4);
            }
        if (!p.alts.empty()) {
            TailLayout tl;
            for (int b = 0; b < batch; ++b) {
                clip_tail_layout(p, n[b], ms[b].T_all, tl);
                for (int l = 0; l < kMaxLevels; ++l) { ms[b].alt_t0[l] = tl.t0[l]; ms[b].alt_th[l] = tl.th[l]; }
                for (size_t a = 0; a < p.alts.size(); ++a)
                    for (int l = 0; l < kMaxLevels; ++l) {
                        ms[b].alt_first[a][l] = tl.first[a][l];
                        if (tl.first[a][l] >= 0) {
                            ms[b].alt_off[a][l] = (int64_t)ladder_elems - tl.first[a][l];   // virtual: sample m lives at alt_off + m
                            ladder_elems += align_up((size_t)tl.count[a][l], 4) + 4;
                        }
                        if (tl.hlen[a][l] > 0) {
                            ms[b].alt_hoff[a][l] = (int64_t)ladder_elems;
                            ms[b].alt_hlen[a][l] = tl.hlen[a][l];
                            ladder_elems += align_up((size_t)tl.hlen[a][l], 4) + 4;
                        } else if (tl.hlen[a][l] < 0) {      // the tail piece holds the whole level (first == 0)
                            ms[b].alt_hoff[a][l] = ms[b].alt_off[a][l];
                            ms[b].alt_hlen[a][l] = ms[b].lvl_len[l];
                        }
                    }
            }
        }
    }
    w.total = w.ladder_off + ladder_elems * sizeof(float) + 256;
    return w;
}

size_t workspace_bytes(const Plan &p, int batch, const int64_t *n) { return ws_layout(p, batch, n, nullptr).total; }

static int slide_class(const CqtItem &it) {
    const int kb = it.kmax - it.kmin + 1;
    return AMT_SLIDE_SPLIT == 0 ? 0 : AMT_SLIDE_SPLIT == 1 ? (kb <= 128 ? 0 : 1) : (kb + 31) / 32;
}

// Launch classes of the VQT-family items: 0 = level 0 | 1 = levels 1 .. AMT_MID_LEVEL | 2 = deeper (FFT per frame, caller's stream,
// each class waits only for the ladder levels it reads), 3 = sliding DFT (side stream), 4 / 5 = exact-ladder tails (FFT per
// frame / sliding DFT, side stream).
static int item_class(const Plan &p, const CqtItem &it, bool overlap) {
    const bool slide = !p.slide_off && is_slide_item(it);
    if (it.alt) return slide ? 5 : 4;
    if (slide) return 3;
    return !overlap ? 0 : it.level == 0 ? 0 : it.level <= AMT_MID_LEVEL ? 1 : 2;
}

int launch_count(const Plan &p, int batch, const int64_t *n) {
    (void)batch;
    (void)n;
    int k = 0;
    if (is_vqt_kind(p)) {
        k += p.n_levels - 1;
        const bool overlap = AMT_LADDER_OVERLAP && !p.serial_launch && p.side_stream[0] != nullptr && p.n_levels > 1;
        // one FFT-per-frame launch per run of equal n_fft and equal class; one sliding-DFT launch per CTA-size class
        int last = -1, last_cls = -1;
        std::map<std::pair<int, int>, int> sl;
        for (const CqtItem &it : p.items) {
            const int cls = item_class(p, it, overlap);
            if (cls == 3 || cls == 5) { ++sl[{cls, slide_class(it)}]; last = -1; continue; }
            if (it.nfft != last || cls != last_cls) { ++k; last = it.nfft; last_cls = cls; }
        }
        for (auto &kv : sl) k += (kv.second + kSlideMaxItems - 1) / kSlideMaxItems;
        for (const AltLadder &al : p.alts) { (void)al; k += p.n_oct; }   // one launch per level of an exact ladder (head and tail pieces)
    } else {
        k += 1;
    }
    if (p.cfg.decibels && p.cfg.kind != AMTFEAT_WAVEFORM) k += 1;
    return k;
}

template <int NC> static size_t stft_smem_nc(const Plan &p) {
    using L = FftLayout<NC>;
    const int TT = kWarpsPerCta * L::G;
    const size_t scr = stft_scr_floats<NC>(p.cfg.n_mels, p.cfg.kind == AMTFEAT_MEL);
    return (size_t)(2 * NC + NC + scr + tile_floats_for(TT, p.cfg.hop_length, 2 * NC)) * sizeof(float);
}
size_t stft_smem_bytes(const Plan &p) {
    switch (p.cfg.n_fft / 2) {
        case 1024: return stft_smem_nc<1024>(p);
        case 512: return stft_smem_nc<512>(p);
        case 256: return stft_smem_nc<256>(p);
        case 128: return stft_smem_nc<128>(p);
        case 64: return stft_smem_nc<64>(p);
        case 32: return stft_smem_nc<32>(p);
        case 16: return stft_smem_nc<16>(p);
        case 8: return stft_smem_nc<8>(p);
        default: return stft_smem_nc<4>(p);
    }
}

template <int NC>
static int launch_stft(const Plan &p, const StftParams &sp_in, int batch, int maxT, cudaStream_t st) {
    using L = FftLayout<NC>;
    const int TT = kWarpsPerCta * L::G;
    const bool mel = p.cfg.kind == AMTFEAT_MEL;
    StftParams sp = sp_in;
    sp.scr_floats = (int)stft_scr_floats<NC>(sp.n_mels, mel);
    sp.tile_floats = tile_floats_for(TT, sp.hop, 2 * NC);
    const size_t smem = (size_t)(2 * NC + NC + sp.scr_floats + sp.tile_floats) * sizeof(float);
    if (smem > 227 * 1024) { set_error("hop_length / n_mels too large for the shared-memory tiles of this n_fft"); return AMTFEAT_ERR_INVALID; }
    // tiles per CTA: amortise the per-CTA prologue while keeping >= ~4 waves of 2 CTAs per SM
    const int ntiles = (maxT + TT - 1) / TT;
    const long long total_tiles = (long long)ntiles * batch;
    sp.tiles_per_cta = (int)std::max<long long>(1, std::min<long long>(8, total_tiles / (148 * 2 * 4)));
    {
        static const int forced = [] { const char *e = std::getenv("AMTFEAT_STFT_TPC"); return e ? atoi(e) : 0; }();
        if (forced > 0) sp.tiles_per_cta = forced;
    }
    dim3 grid((ntiles + sp.tiles_per_cta - 1) / sp.tiles_per_cta, batch);
    ProfScope ps(p, mel ? "stft_kernel_mel" : "stft_kernel_mag", st);
    if (mel) stft_kernel<NC, MODE_MEL><<<grid, kThreads, smem, st>>>(sp);
    else stft_kernel<NC, MODE_STFT><<<grid, kThreads, smem, st>>>(sp);
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

// FFT-per-frame launch over the items [item0, item0 + nitems) (one n_fft, one class); `frames` = most frames any
// (clip, item) of the launch computes (T_all, or the length of the tail for exact-ladder items).
template <int NC>
static int launch_cqt(const Plan &p, CqtParams cp, int item0, int nitems, int batch, int frames, bool tail, cudaStream_t st) {
    using L = FftLayout<NC>;
    const int TT = kWarpsPerCta * L::G;
    int maxhop = 0, maxrows = 0, maxblk = 0, maxkb = 0;
    for (int i = item0; i < item0 + nitems; ++i) {
        maxhop = std::max(maxhop, p.items[i].hop);
        maxrows = std::max(maxrows, p.items[i].nrows);
        maxblk = std::max(maxblk, p.items[i].nblk);
        maxkb = std::max(maxkb, p.items[i].kmax - p.items[i].kmin + 1);
    }
    cp.tile_floats = tile_floats_for(TT, maxhop, 2 * NC);
    cp.items = p.d_items + item0;
    const FftTables &ft = p.fft.at(NC);
    cp.tw1 = reinterpret_cast<const float2 *>(ft.d_tw1);
    cp.tw2 = reinterpret_cast<const float2 *>(ft.d_tw2);
    const int ntiles = (frames + TT - 1) / TT;
    dim3 grid(ntiles, batch, nitems);
    const std::string nm = tail ? std::string("tail_cqt_kernel") : "cqt_kernel_nfft" + std::to_string(2 * NC);
    size_t smem;
    if (cqt_use_blocks<NC>()) {
        // phase B reuses the retired FFT scratch: Dbuf | the item's weights | its block descriptors.  The audio
        // tile keeps its own region so that the next tile can stream in during phase B.
        const int scratch_floats = kWarpsPerCta * L::WARP_PITCH;
        const int fixed_floats = 2 * NC + 2 * (NC + 2);
        int maxw = 0;
        for (int i = item0; i < item0 + nitems; ++i) maxw = std::max(maxw, p.items[i].wcount);
        const int dbuf_floats = (maxkb * (TT + kDbufPad) * 2 + 3) / 4 * 4;
        const int blk_floats = maxblk * (int)(sizeof(CqtBlock4) / 4);
        cp.stage_blocks = 0;
        cp.stage_rows = 0;
        cp.dbuf_off = 0;
        cp.w_off = dbuf_floats;
        cp.blk_off = cp.w_off + maxw * 4;
        cp.region_floats = (std::max(scratch_floats, cp.blk_off + blk_floats) + 3) / 4 * 4;
        smem = (size_t)(fixed_floats + cp.region_floats + cp.tile_floats) * sizeof(float);
        // tiles per CTA: amortise the per-CTA prologue while keeping >= ~4 waves of 2 CTAs per SM
        const long long total = (long long)ntiles * batch * nitems;
        cp.tiles_per_cta = (int)std::max<long long>(1, std::min<long long>(AMT_CQT_TPC_MAX, total / (148 * 2 * AMT_CQT_WAVES)));
        grid.x = (ntiles + cp.tiles_per_cta - 1) / cp.tiles_per_cta;
    } else {
        cp.stage_blocks = 0;
        cp.stage_rows = std::max(1, std::min(maxrows, 3072 / (TT + 1)));
        smem = cqt_smem<NC>(cp.tile_floats, cp.stage_rows);
    }
    if (smem > 226 * 1024) { set_error("hop_length too large for the shared-memory audio tile"); return AMTFEAT_ERR_INVALID; }
    ProfScope ps(p, nm.c_str(), st);
    if (cqt_use_blocks<NC>()) {
        constexpr int JFULL = (NC + 1 + 31) / 32;
        if ((maxkb + 31) / 32 <= (JFULL + 1) / 2) cqt_kernel<(NC >= 64 ? NC : 64), true><<<grid, kThreads, smem, st>>>(cp);
        else cqt_kernel<(NC >= 64 ? NC : 64), false><<<grid, kThreads, smem, st>>>(cp);
    }
    else cqt_small_kernel<(NC < 64 ? NC : 32)><<<grid, kThreads, smem, st>>>(cp);
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

static int launch_cqt_nc(const Plan &p, const CqtParams &cp, int NC, int item0, int nitems, int batch, int frames, bool tail, cudaStream_t st) {
    switch (NC) {
        case 1024: return launch_cqt<1024>(p, cp, item0, nitems, batch, frames, tail, st);
        case 512: return launch_cqt<512>(p, cp, item0, nitems, batch, frames, tail, st);
        case 256: return launch_cqt<256>(p, cp, item0, nitems, batch, frames, tail, st);
        case 128: return launch_cqt<128>(p, cp, item0, nitems, batch, frames, tail, st);
        case 64: return launch_cqt<64>(p, cp, item0, nitems, batch, frames, tail, st);
        case 32: return launch_cqt<32>(p, cp, item0, nitems, batch, frames, tail, st);
        case 16: return launch_cqt<16>(p, cp, item0, nitems, batch, frames, tail, st);
        case 8: return launch_cqt<8>(p, cp, item0, nitems, batch, frames, tail, st);
        default: return launch_cqt<4>(p, cp, item0, nitems, batch, frames, tail, st);
    }
}

// Sliding-DFT launch over the items `idx` (all of them pass is_slide_item); frames[i] = most frames a clip computes of item idx[i].
static int launch_slide(const Plan &p, const CqtParams &cp, const std::vector<int> &idx, const std::vector<int> &frames, int batch,
                        bool tail, cudaStream_t st) {
    for (size_t i0 = 0; i0 < idx.size(); i0 += kSlideMaxItems) {
        const int cnt = (int)std::min<size_t>(kSlideMaxItems, idx.size() - i0);
        SlideParams sp{};
        sp.audio = cp.audio; sp.ladder = cp.ladder; sp.out = cp.out; sp.meta = cp.meta; sp.maxbuf = cp.maxbuf;
        sp.items = p.d_items; sp.blocks = cp.blocks; sp.weights4 = cp.weights4; sp.C = cp.C; sp.decibels = cp.decibels;
        sp.alt_mask = cp.alt_mask;
        int maxkb = 0, maxw = 0, maxblk = 0, maxx = 0, tiles = 0;
        for (int i = 0; i < cnt; ++i) {
            const CqtItem &it = p.items[idx[i0 + i]];
            sp.idx[i] = idx[i0 + i];
            sp.tw2[i] = reinterpret_cast<const float2 *>(p.fft.at(it.nfft / 2).d_tw2);
            const int Tt = slide_tile_frames(it.hop);
            maxkb = std::max(maxkb, it.kmax - it.kmin + 1);
            maxw = std::max(maxw, it.wcount);
            maxblk = std::max(maxblk, it.nblk);
            maxx = std::max(maxx, Tt * it.hop + it.nfft);
            tiles = std::max(tiles, (frames[i0 + i] + Tt - 1) / Tt);
        }
        if (tiles == 0) continue;
        const int threads = (maxkb + 31) / 32 * 32;
        sp.w_off = (kSlideBufs * maxkb * kSlideDP * 2 + 3) / 4 * 4;
        sp.blk_off = sp.w_off + maxw * 4;
        sp.x_off = (sp.blk_off + maxblk * (int)(sizeof(CqtBlock4) / 4) + 3) / 4 * 4;
        const size_t smem = (size_t)(sp.x_off + maxx + 8) * sizeof(float);
        if (smem > 226 * 1024) { set_error("sliding-DFT tile does not fit in shared memory"); return AMTFEAT_ERR_INVALID; }
        dim3 grid(tiles, batch, cnt);
        ProfScope ps(p, tail ? "tail_slide_kernel" : "cqt_slide_kernel", st);
        cqt_slide_kernel<<<grid, threads, smem, st>>>(sp);
        AMT_CUDA(cudaGetLastError());
    }
    return AMTFEAT_OK;
}

// Joins the side stream back into the caller's stream on every exit path of the VQT family (error returns included), so
// that no call leaves work dangling behind the caller's back.
struct SideJoin {
    cudaStream_t st, side;
    cudaEvent_t ev;
    bool armed = false;
    ~SideJoin() {
        if (!armed || side == st) return;
        cudaEventRecord(ev, side);
        cudaStreamWaitEvent(st, ev, 0);
    }
};

// The dB epilogue of a batch whose producers ran with defer_epilogue, written to `dst` (same element offsets as d_out; mapped
// pinned host memory for the pipelined executor) on `stream`.  A handful of CTAs: the store stream is PCIe-bound, and the SMs
// are busy with the next batch.
int epilogue_out(const Plan &p, const float *d_out, void *d_ws, int batch, int maxT, float *dst, bool few_ctas, void *stream) {
    if (!has_db_epilogue(p) || batch <= 0 || maxT <= 0) return AMTFEAT_OK;
    DeviceGuard guard(p.device);
    char *ws = static_cast<char *>(d_ws);
    const ClipMeta *d_meta = reinterpret_cast<const ClipMeta *>(ws);
    const float *d_max = reinterpret_cast<const float *>(ws + align_up((size_t)batch * sizeof(ClipMeta), 256));
    const int64_t maxcount = (int64_t)p.F * maxT;
    unsigned gx = (unsigned)std::min<int64_t>(1024, (maxcount + kThreads * 4 - 1) / (kThreads * 4));
    if (few_ctas) {
        static const int total = [] { const char *e = std::getenv("AMTFEAT_PIPE_CTAS"); return e ? std::max(1, atoi(e)) : 2 * 148; }();
        gx = std::min<unsigned>(gx, std::max(1, (total + batch * p.C - 1) / (batch * p.C)));
    }
    dim3 grid(std::max(1u, gx), batch * p.C);
    std::lock_guard<std::mutex> lock(p.call_mu);
    ProfScope ps(p, few_ctas ? "db_epilogue_to_host_kernel" : "db_epilogue_kernel", reinterpret_cast<cudaStream_t>(stream));
    db_epilogue_kernel<<<grid, kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(d_out, dst, d_meta, d_max, p.C, p.F,
                                                                                        p.cfg.kind == AMTFEAT_POWER ? 0 : 1);
    AMT_CUDA(cudaGetLastError());
    return AMTFEAT_OK;
}

int process(const Plan &p, const float *d_audio, const int64_t *in_off, const int64_t *n, const int64_t *out_off,
            int batch, float *d_out, void *d_ws, size_t ws_bytes, void *stream, bool defer_epilogue, int *max_frames) {
    if (p.device < 0) { set_error("host-only plan: no CUDA device (there is no CPU compute path)"); return AMTFEAT_ERR_NO_DEVICE; }
    if (batch <= 0) return AMTFEAT_OK;
    if ((reinterpret_cast<uintptr_t>(d_audio) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15) || (reinterpret_cast<uintptr_t>(d_ws) & 255)) {
        set_error("d_audio / d_out must be 16-byte aligned and the workspace 256-byte aligned");
        return AMTFEAT_ERR_INVALID;
    }
    DeviceGuard guard(p.device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const amtfeat_config &c = p.cfg;
    std::vector<ClipMeta> metas;
    const WsLayout w = ws_layout(p, batch, n, &metas);
    if (!w.ok) { set_error("input too short (an uncentered frame longer than the padded signal, or a clip shorter than the early-downsampling factor of a CQT / VQT: librosa raises 'Input signal length=... is too short for ...-octave CQT')"); return AMTFEAT_ERR_INVALID; }
    if (ws_bytes < w.total) { set_error("workspace too small"); return AMTFEAT_ERR_WORKSPACE; }
    int maxT = 0, maxTall = 0;
    int64_t maxn = 0;
    for (int b = 0; b < batch; ++b) {
        if (in_off[b] % 4 != 0) { set_error("clip offsets must be multiples of 4 elements"); return AMTFEAT_ERR_INVALID; }
        metas[b].in_off = in_off[b];
        metas[b].lvl_off[0] = in_off[b];
        metas[b].lvl_len[0] = (int32_t)n[b];
        metas[b].n = n[b];
        metas[b].out_off = out_off[b];
        maxT = std::max<int>(maxT, metas[b].T);
        maxTall = std::max<int>(maxTall, metas[b].T_all);
        maxn = std::max(maxn, n[b]);
    }
    if (max_frames) *max_frames = maxT;
    if (maxT == 0) return AMTFEAT_OK;
    char *ws = static_cast<char *>(d_ws);
    ClipMeta *d_meta = reinterpret_cast<ClipMeta *>(ws + w.meta_off);
    float *d_max = reinterpret_cast<float *>(ws + w.max_off);
    float *d_ladder = reinterpret_cast<float *>(ws + w.ladder_off);
    // everything below is enqueued under the plan's call lock: the descriptor ring, the fork / join events of the call slot and
    // the profiling records are plan state; concurrent callers of one plan serialise their (short) host-side enqueue here
    std::lock_guard<std::mutex> call_lock(p.call_mu);
    {
        // descriptors go through a pinned ring slot (truly asynchronous copy); a slot is reused once its last copy has completed
        const size_t need = metas.size() * sizeof(ClipMeta);
        if (need > p.meta_slot_bytes) {
            for (void *&e : p.meta_events) {
                if (e) AMT_CUDA(cudaEventSynchronize(reinterpret_cast<cudaEvent_t>(e)));
                else AMT_CUDA(cudaEventCreateWithFlags(reinterpret_cast<cudaEvent_t *>(&e), cudaEventDisableTiming));
            }
            if (p.meta_ring) cudaFreeHost(p.meta_ring);
            p.meta_ring = nullptr;
            p.meta_slot_bytes = 0;
            const size_t slot = align_up(std::max<size_t>(need, 64 * 1024), 4096);
            AMT_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&p.meta_ring), slot * Plan::kMetaSlots, cudaHostAllocDefault));
            p.meta_slot_bytes = slot;
        }
        const unsigned slot = p.meta_next++ % Plan::kMetaSlots;
        cudaEvent_t ev = reinterpret_cast<cudaEvent_t>(p.meta_events[slot]);
        AMT_CUDA(cudaEventSynchronize(ev));
        char *h = p.meta_ring + (size_t)slot * p.meta_slot_bytes;
        std::memcpy(h, metas.data(), need);
        static_assert(sizeof(ClipMeta) % 16 == 0, "ClipMeta is copied in 16-byte words");
        if (p.meta_memcpy) {
            AMT_CUDA(cudaMemcpyAsync(d_meta, h, need, cudaMemcpyHostToDevice, st));
        } else {
            // pinned memory is device-addressable at its host address (unified addressing)
            const int n16 = (int)(need / 16);
            copy_meta_kernel<<<(n16 + kThreads - 1) / kThreads, kThreads, 0, st>>>(reinterpret_cast<const int4 *>(h), reinterpret_cast<int4 *>(d_meta), n16);
            AMT_CUDA(cudaGetLastError());
        }
        AMT_CUDA(cudaEventRecord(ev, st));
    }
    if (c.decibels) AMT_CUDA(cudaMemsetAsync(d_max, 0, (size_t)batch * p.C * sizeof(float), st));
    int rc = AMTFEAT_OK;
    int scale01 = 1;

    if (c.kind == AMTFEAT_STFT || c.kind == AMTFEAT_MEL) {
        const int NC = c.n_fft / 2;
        const FftTables &ft = p.fft.at(NC);
        StftParams sp{};
        sp.audio = d_audio; sp.out = d_out; sp.meta = d_meta; sp.maxbuf = d_max; sp.window = p.d_window;
        sp.tw1 = reinterpret_cast<const float2 *>(ft.d_tw1); sp.tw2 = reinterpret_cast<const float2 *>(ft.d_tw2);
        sp.mel_seg_start = p.d_mel_seg_start; sp.mel_gsteps = p.d_mel_gsteps; sp.mel_goff = p.d_mel_goff;
        sp.mel_ww = reinterpret_cast<const float2 *>(p.d_mel_ww);
        sp.hop = c.hop_length; sp.pad = c.center ? c.n_fft / 2 : 0; sp.n_mels = c.n_mels; sp.decibels = c.decibels;
        switch (NC) {
            case 1024: rc = launch_stft<1024>(p, sp, batch, maxT, st); break;
            case 512: rc = launch_stft<512>(p, sp, batch, maxT, st); break;
            case 256: rc = launch_stft<256>(p, sp, batch, maxT, st); break;
            case 128: rc = launch_stft<128>(p, sp, batch, maxT, st); break;
            case 64: rc = launch_stft<64>(p, sp, batch, maxT, st); break;
            case 32: rc = launch_stft<32>(p, sp, batch, maxT, st); break;
            case 16: rc = launch_stft<16>(p, sp, batch, maxT, st); break;
            case 8: rc = launch_stft<8>(p, sp, batch, maxT, st); break;
            default: rc = launch_stft<4>(p, sp, batch, maxT, st); break;
        }
        if (rc) return rc;
    } else if (c.kind == AMTFEAT_POWER) {
        dim3 grid((maxT + kWarpsPerCta - 1) / kWarpsPerCta, batch);
        ProfScope ps(p, "power_kernel", st);
        power_kernel<<<grid, kThreads, 0, st>>>(d_audio, d_out, d_meta, d_max, c.hop_length, c.win_length,
                                                c.center ? c.win_length / 2 : 0, c.decibels);
        AMT_CUDA(cudaGetLastError());
        scale01 = 0;
    } else if (c.kind == AMTFEAT_WAVEFORM) {
        dim3 grid((maxT + 31) / 32, (c.win_length + 7) / 8, batch);
        frames_kernel<<<grid, kThreads, 0, st>>>(d_audio, d_out, d_meta, c.hop_length, c.win_length,
                                                 c.center ? c.win_length / 2 : 0);
        AMT_CUDA(cudaGetLastError());
        return AMTFEAT_OK;
    } else {
        // Decimation ladder (levels 1 .. n_levels-1) on one of the plan's side streams, FFT-per-frame launches on the caller's
        // stream.  The deep ladder levels are short and latency-bound, so the FFT-per-frame launches are cut by ladder depth
        // (item_class) and each class only waits for the ladder levels it reads: the ladder runs underneath the projection of
        // the shallower levels.  The sliding-DFT items and the exact-ladder tails follow the ladder on the side stream.
        const int ntaps = (int)p.taps.size(), D = (ntaps - 1) / 2;
        const int dlen = dec_front_pad(D) + kDecTile + D + 8;
        const int dplen = ((dlen + ((dlen >> 4) << 2)) + 7) & ~3;
        const size_t dsmem = (size_t)(2 * dplen + 2 * dec_jtot(D)) * sizeof(float);
        int64_t len = maxn;
        const int mode = (p.decim_mode == 0 && !p.decim_h64.empty()) ? 0 : (p.decim_mode <= 1 && !p.decim_hh.empty()) ? 1 : 2;
        const bool overlap = AMT_LADDER_OVERLAP && !p.serial_launch && p.side_stream[0] != nullptr && p.n_levels > 1;
        const unsigned call_slot = p.call_next++ % Plan::kCallSlots;
        cudaStream_t lst = overlap ? reinterpret_cast<cudaStream_t>(p.side_stream[call_slot]) : st;
        cudaEvent_t ev_fork = reinterpret_cast<cudaEvent_t>(p.call_events[call_slot][0]), ev_mid = reinterpret_cast<cudaEvent_t>(p.call_events[call_slot][1]),
                    ev_all = reinterpret_cast<cudaEvent_t>(p.call_events[call_slot][2]), ev_side = reinterpret_cast<cudaEvent_t>(p.call_events[call_slot][3]);
        constexpr int kMidLevel = AMT_MID_LEVEL;           // classes: level 0 | 1 .. kMidLevel | deeper
        // the deepest level the FFT-per-frame launches of the caller's stream read
        int max_fft_level = 0;
        for (const CqtItem &it : p.items)
            if (item_class(p, it, overlap) <= 2) max_fft_level = std::max(max_fft_level, (int)it.level);
        const int deep_level = std::min(p.n_levels - 1, std::max(max_fft_level, kMidLevel));
        // the exact-ladder pieces and their items depend on the audio alone: their own stream, beside the shared ladder and the
        // sliding-DFT launches (on short clips that chain was the critical path of the call)
        cudaStream_t tst = (overlap && !p.alts.empty() && AMT_TAIL_STREAM) ? reinterpret_cast<cudaStream_t>(p.tail_stream[call_slot]) : lst;
        SideJoin join{st, lst, ev_side};
        SideJoin join_tail{st, tst, reinterpret_cast<cudaEvent_t>(p.call_events[call_slot][4])};
        if (overlap) {
            AMT_CUDA(cudaEventRecord(ev_fork, st));        // clip descriptors / cleared maxima are in place
            AMT_CUDA(cudaStreamWaitEvent(lst, ev_fork, 0));
            join.armed = true;
            if (tst != lst) {
                AMT_CUDA(cudaStreamWaitEvent(tst, ev_fork, 0));
                join_tail.armed = true;
            }
        }
        for (int l = 1; l < p.n_levels; ++l) {
            len = (len + 1) / 2;
            if (mode == 0) {
                Dec64Params dp{};
                dp.audio = d_audio; dp.ladder = d_ladder; dp.meta = d_meta;
                dp.tw = reinterpret_cast<const double2 *>(p.d_decim_tw64); dp.H = reinterpret_cast<const double2 *>(p.d_decim_h64);
                dp.level_out = l; dp.D = D; dp.M = 1024 - D;
                const int64_t npairs = (len + 2 * dp.M - 1) / (2 * dp.M);
                dp.pairs_per_cta = (int)std::max<int64_t>(1, std::min<int64_t>(8, npairs * batch / (148 * 4 * 2)));
                const size_t fsmem = (size_t)kD64Buf * sizeof(double2);
                dim3 grid((unsigned)((npairs + dp.pairs_per_cta - 1) / dp.pairs_per_cta), batch);
                ProfScope ps(p, "decimate_fft64_kernel", lst);
                decimate_fft64_kernel<<<grid, kD64Threads, fsmem, lst>>>(dp);
            } else if (mode == 1) {
                // (both float32 forms have a ~20 us latency floor per launch on the short, deep levels: one warp runs three 1024-point
                // transforms back to back / one thread runs 16 x 389 MACs; mixing the forms per level was measured and is no faster)
                using L = FftLayout<1024>;
                DecFftParams dp{};
                dp.audio = d_audio; dp.ladder = d_ladder; dp.meta = d_meta; dp.hh = reinterpret_cast<const float4 *>(p.d_decim_hh);
                const FftTables &ft = p.fft.at(1024);
                dp.tw1 = reinterpret_cast<const float2 *>(ft.d_tw1); dp.tw2 = reinterpret_cast<const float2 *>(ft.d_tw2);
                dp.level_out = l; dp.D = D; dp.M = 1024 - D;
                const size_t fsmem = (size_t)(1024 + kDfWarps * (L::WARP_PITCH / 2) + kDfWarps * kDfCd) * sizeof(float2);
                dim3 grid((unsigned)((len + (int64_t)kDfWarps * 2 * dp.M - 1) / ((int64_t)kDfWarps * 2 * dp.M)), batch);
                ProfScope ps(p, "decimate_fft_kernel", lst);
                decimate_fft_kernel<<<grid, kDfThreads, fsmem, lst>>>(dp);
            } else {
                dim3 grid((unsigned)((len + kDecTile - 1) / kDecTile), batch);
                ProfScope ps(p, "decimate_kernel", lst);
                decimate_kernel<<<grid, kThreads, dsmem, lst>>>(d_audio, d_ladder, d_meta, p.d_taps, ntaps, l);
            }
            AMT_CUDA(cudaGetLastError());
            if (overlap && l == std::min(kMidLevel, p.n_levels - 1)) AMT_CUDA(cudaEventRecord(ev_mid, lst));
            if (overlap && l == deep_level) AMT_CUDA(cudaEventRecord(ev_all, lst));
        }
        CqtParams cp{};
        cp.audio = d_audio; cp.ladder = d_ladder; cp.out = d_out; cp.meta = d_meta; cp.maxbuf = d_max;
        cp.rows = p.d_rows; cp.weights = reinterpret_cast<const float2 *>(p.d_weights);
        cp.blocks = p.d_blocks; cp.weights4 = reinterpret_cast<const float4 *>(p.d_weights4);
        cp.C = p.C; cp.F = p.F; cp.decibels = c.decibels; cp.alt_mask = p.alt_mask;
        // tiles (of TT frames) an exact-ladder item computes at most, as frames: head tiles + tail tiles, longest clip of the batch
        auto tail_frames = [&](const CqtItem &it, int TT) {
            int tiles = 0;
            for (int b = 0; b < batch; ++b) {
                const ClipMeta &m = metas[b];
                int t = (std::min(m.alt_th[it.level], m.T_all) + TT - 1) / TT;
                if (m.alt_t0[it.level] < m.T_all) t += (m.T_all - m.alt_t0[it.level] + TT - 1) / TT;
                tiles = std::max(tiles, t);
            }
            return tiles * TT;
        };
        // Sliding-DFT launches of one class (3: shared ladder, 5: exact-ladder tails) on the side stream, one launch per CTA-size
        // class (narrow bands first: they hold the longest tiles), so that a narrow band does not carry the idle warps, registers
        // and shared memory of the widest one.  CTAs are handed out in grid order (z slowest): the items with the longest tiles
        // (smallest hop: most frames and the longest lead-in per tile) go first so that they do not form the tail of the launch.
        auto launch_slides = [&](int cls, cudaStream_t s) -> int {
            std::map<int, std::vector<int>> classes;
            for (size_t i = 0; i < p.items.size(); ++i)
                if (item_class(p, p.items[i], overlap) == cls) classes[slide_class(p.items[i])].push_back((int)i);
            for (auto &kv : classes) {
                std::vector<int> &idx = kv.second;
                std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return p.items[a].hop < p.items[b].hop; });
                std::vector<int> frames;
                for (int i : idx) frames.push_back(cls == 5 ? tail_frames(p.items[i], slide_tile_frames(p.items[i].hop)) : maxTall);
                const int r = launch_slide(p, cp, idx, frames, batch, cls == 5, s);
                if (r) return r;
            }
            return AMTFEAT_OK;
        };
        // FFT-per-frame launches of one class: items are sorted by (ladder, n_fft, level); a launch takes a run of equal n_fft
        auto launch_ffts = [&](int cls, cudaStream_t s) -> int {
            size_t i0 = 0;
            while (i0 < p.items.size()) {
                size_t i1 = i0;
                const int c0 = item_class(p, p.items[i0], overlap);
                int frames = 0;
                while (i1 < p.items.size() && p.items[i1].nfft == p.items[i0].nfft && p.items[i1].alt == p.items[i0].alt &&
                       item_class(p, p.items[i1], overlap) == c0) {
                    frames = std::max(frames, cls == 4 ? tail_frames(p.items[i1], 16384 / p.items[i1].nfft) : maxTall);
                    ++i1;
                }
                if (c0 == cls && frames > 0) {
                    const int r = launch_cqt_nc(p, cp, p.items[i0].nfft / 2, (int)i0, (int)(i1 - i0), batch, frames, cls == 4, s);
                    if (r) return r;
                }
                i0 = i1;
            }
            return AMTFEAT_OK;
        };
        if ((rc = launch_slides(3, lst))) return rc;
        // exact ladders: the tail of every level (level eds in one 2^eds : 1 pass over the audio), then their items
        for (size_t a = 0; a < p.alts.size(); ++a) {
            const AltLadder &al = p.alts[a];
            for (int l = al.eds; l < al.eds + p.n_oct; ++l) {
                TailParams tp{};
                tp.audio = d_audio; tp.ladder = d_ladder; tp.meta = d_meta; tp.alt = (int)a; tp.level_out = l;
                if (l == al.eds) { tp.taps = al.d_taps; tp.ntaps = (int)al.taps.size(); tp.factor = 1 << al.eds; tp.level_in = 0; }
                else { tp.taps = p.d_taps64; tp.ntaps = (int)p.taps64.size(); tp.factor = 2; tp.level_in = l - 1; }
                int count = 0;       // one launch: tail pieces (z = 0) and head pieces (z = 1) of all clips
                for (int b = 0; b < batch; ++b) {
                    if (metas[b].alt_first[a][l] > 0) count = std::max(count, metas[b].alt_hlen[a][l]);
                    if (metas[b].alt_first[a][l] >= 0) count = std::max(count, metas[b].lvl_len[l] - metas[b].alt_first[a][l]);
                }
                if (count <= 0) continue;
                // four outputs per thread unless that leaves the GPU mostly empty (CTAs of a launch: pieces x clips x 2)
                const int out4 = 4 * kThreads;
                const bool wide = (long long)((count + out4 - 1) / out4) * batch * 2 >= 2 * 148;
                const int tail_out = wide ? out4 : kThreads;
                dim3 grid((count + tail_out - 1) / tail_out, batch, 2);
                const int nq = ((tp.ntaps + tp.factor - 1) / tp.factor + 3) & ~3;
                const bool exact = tp.factor >= 8;        // float64 taps and products where the stop band holds nearly all the energy
                const size_t tsmem = ((size_t)tp.factor * ((tail_out + nq - 1) | 1) + (size_t)tp.factor * nq * (exact ? 2 : 1)) * sizeof(float);
                if (tsmem > 200 * 1024) { set_error("one-shot early-downsampling filter too long for the tail kernel"); return AMTFEAT_ERR_INVALID; }
                ProfScope ps(p, "tail_decimate_kernel", tst);
                if (exact) {
                    if (wide) tail_decimate_kernel<4, true><<<grid, kThreads, tsmem, tst>>>(tp);
                    else tail_decimate_kernel<1, true><<<grid, kThreads, tsmem, tst>>>(tp);
                } else {
                    if (wide) tail_decimate_kernel<4, false><<<grid, kThreads, tsmem, tst>>>(tp);
                    else tail_decimate_kernel<1, false><<<grid, kThreads, tsmem, tst>>>(tp);
                }
                AMT_CUDA(cudaGetLastError());
            }
        }
        if (!p.alts.empty()) {
            if ((rc = launch_ffts(4, tst))) return rc;
            if ((rc = launch_slides(5, tst))) return rc;
        }
        for (int cls = 0; cls < (overlap ? 3 : 1); ++cls) {
            if (cls == 1) AMT_CUDA(cudaStreamWaitEvent(st, ev_mid, 0));
            if (cls == 2) AMT_CUDA(cudaStreamWaitEvent(st, ev_all, 0));
            if ((rc = launch_ffts(cls, st))) return rc;
        }
        // `join` (destructor) gives the caller's stream a dependency on everything the side stream was given
    }
    if (c.decibels && !defer_epilogue) {
        int64_t maxcount = (int64_t)p.F * maxT;
        unsigned gx = (unsigned)std::min<int64_t>(1024, (maxcount + kThreads * 4 - 1) / (kThreads * 4));
        dim3 grid(std::max(1u, gx), batch * p.C);
        ProfScope ps(p, "db_epilogue_kernel", st);
        db_epilogue_kernel<<<grid, kThreads, 0, st>>>(d_out, d_out, d_meta, d_max, p.C, p.F, scale01);
        AMT_CUDA(cudaGetLastError());
    }
    return AMTFEAT_OK;
}

}  // namespace amtfeat
