// Warp-level FFT building blocks (FP32 SIMT, sm_100a).
//
// One warp owns a "unit" of 1024 complex points = G = 1024 / NC independent complex FFTs of length
// NC (NC = n_fft / 2: a real frame of n_fft samples is packed as NC complex points).  Every FFT is two
// register-resident passes (radix R1 then radix R2, R1 * R2 = NC, both <= 32) with ONE shared-memory
// transpose in between, so each thread always carries 32 complex values and no block-wide barrier is
// needed inside the transform (only __syncwarp).
#pragma once

#include <cuda_runtime.h>

#ifndef AMT_FFT_PACKED
#define AMT_FFT_PACKED 1
#endif
#ifndef AMT_TW_GLOBAL
#define AMT_TW_GLOBAL 0
#endif

namespace amtfeat {

template <int NC> struct FftCfg;
template <> struct FftCfg<1024> { static constexpr int R1 = 32, R2 = 32; };
template <> struct FftCfg<512>  { static constexpr int R1 = 16, R2 = 32; };
template <> struct FftCfg<256>  { static constexpr int R1 = 16, R2 = 16; };
template <> struct FftCfg<128>  { static constexpr int R1 = 8,  R2 = 16; };
template <> struct FftCfg<64>   { static constexpr int R1 = 8,  R2 = 8;  };
template <> struct FftCfg<32>   { static constexpr int R1 = 4,  R2 = 8;  };
template <> struct FftCfg<16>   { static constexpr int R1 = 4,  R2 = 4;  };
template <> struct FftCfg<8>    { static constexpr int R1 = 2,  R2 = 4;  };
template <> struct FftCfg<4>    { static constexpr int R1 = 2,  R2 = 2;  };

// Shared-memory footprint (in float2) of one FFT inside a warp's scratch: R1 rows of R2 + 1 (padded
// so the transposed read of pass 2 is bank-conflict free).  NC + R1 >= NC + 1, so the same region
// later holds the NC + 1 output bins.
template <int NC> struct FftLayout {
    static constexpr int R1 = FftCfg<NC>::R1, R2 = FftCfg<NC>::R2;
    static constexpr int G = 1024 / NC;              // FFTs per warp unit
    static constexpr int S = R1 * (R2 + 1);          // float2 per FFT region
    static constexpr int SCR = G * S;                // float2 per warp
    static constexpr int WARP_PITCH = 2 * SCR + 4;   // floats; +4 keeps rows of different warps on different banks
};

// Packed FP32x2 arithmetic (sm_100a FADD2 / FFMA2): one issue slot for both components of a complex value.  The FFT
// kernels are issue-bound, not FMA-pipe-bound (profiles/), so halving the slots of the complex adds pays directly.
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n add.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd;}\n"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n sub.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd;}\n"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mul.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd;}\n"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
        " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}\n"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// exp(-2 pi i m / 32), m = 0..15 (round-to-nearest float literals of the exact values)
__device__ __forceinline__ float2 w32(int m) {
    switch (m) {
        case 0: return make_float2(1.0f, -0.0f);
        case 1: return make_float2(0.98078528040323043f, -0.19509032201612825f);
        case 2: return make_float2(0.92387953251128674f, -0.38268343236508978f);
        case 3: return make_float2(0.83146961230254524f, -0.55557023301960218f);
        case 4: return make_float2(0.70710678118654757f, -0.70710678118654757f);
        case 5: return make_float2(0.55557023301960218f, -0.83146961230254524f);
        case 6: return make_float2(0.38268343236508978f, -0.92387953251128674f);
        case 7: return make_float2(0.19509032201612825f, -0.98078528040323043f);
        case 8: return make_float2(0.0f, -1.0f);
        case 9: return make_float2(-0.19509032201612825f, -0.98078528040323043f);
        case 10: return make_float2(-0.38268343236508978f, -0.92387953251128674f);
        case 11: return make_float2(-0.55557023301960218f, -0.83146961230254524f);
        case 12: return make_float2(-0.70710678118654757f, -0.70710678118654757f);
        case 13: return make_float2(-0.83146961230254524f, -0.55557023301960218f);
        case 14: return make_float2(-0.92387953251128674f, -0.38268343236508978f);
        default: return make_float2(-0.98078528040323043f, -0.19509032201612825f);
    }
}

template <int R> __device__ __forceinline__ constexpr int brev(int i) {
    int r = 0;
    for (int b = 1; b < R; b <<= 1) { r = (r << 1) | (i & 1); i >>= 1; }
    return r;
}

// In-register forward DFT of R points (radix-2 decimation in frequency).  Output index k lives in
// register brev<R>(k).  All loop indices are compile-time after unrolling, so the twiddle switch and
// the trivial-twiddle shortcuts fold away.
template <int R> __device__ __forceinline__ void fft_regs(float2 (&v)[R]) {
#pragma unroll
    for (int half = R / 2; half >= 1; half >>= 1) {
#pragma unroll
        for (int blk = 0; blk < R; blk += 2 * half) {
#pragma unroll
            for (int j = 0; j < half; ++j) {
                const float2 a = v[blk + j], b = v[blk + j + half];
#if AMT_FFT_PACKED
                v[blk + j] = fadd2(a, b);
                const float2 d = fsub2(a, b);
#else
                v[blk + j] = make_float2(a.x + b.x, a.y + b.y);
                const float2 d = make_float2(a.x - b.x, a.y - b.y);
#endif
                const int m = j * (16 / half);  // twiddle exp(-2 pi i j / (2 half)) = w32(m)
                if (m == 0) {
                    v[blk + j + half] = d;
                } else if (m == 8) {  // times -i
                    v[blk + j + half] = make_float2(d.y, -d.x);
                } else if (m == 4) {  // times (1 - i) / sqrt(2)
                    v[blk + j + half] = make_float2((d.x + d.y) * 0.70710678118654757f, (d.y - d.x) * 0.70710678118654757f);
                } else if (m == 12) {  // times (-1 - i) / sqrt(2)
                    v[blk + j + half] = make_float2((d.y - d.x) * 0.70710678118654757f, -(d.x + d.y) * 0.70710678118654757f);
                } else {
                    v[blk + j + half] = cmul(d, w32(m));
                }
            }
        }
    }
}

// Forward complex FFTs of one warp unit.  `load(g, n)` returns input point n of FFT g.
// On return (after the trailing __syncwarp) scr[g * S + k] holds bin k of FFT g, k = 0..NC-1.
template <int NC, bool TWG = false, typename LoadFn>
__device__ __forceinline__ void warp_fft_unit(float2 *__restrict__ scr, const float2 *__restrict__ tw1, int lane, LoadFn load) {
    using L = FftLayout<NC>;
    constexpr int R1 = L::R1, R2 = L::R2, S = L::S;
    constexpr int U1 = 32 / R1, U2 = 32 / R2;
    // pass 1: R1-point DFTs over n1 (input stride R2), twiddle, store transposed
#pragma unroll
    for (int u = 0; u < U1; ++u) {
        const int q = lane + 32 * u;
        const int g = q / R2, n2 = q % R2;
        float2 v[R1];
#pragma unroll
        for (int n1 = 0; n1 < R1; ++n1) v[n1] = load(g, R2 * n1 + n2);
        fft_regs<R1>(v);
#pragma unroll
        for (int i = 0; i < R1; ++i) {
            const int k1 = brev<R1>(i);
            const float2 y = (k1 == 0) ? v[i] : cmul(v[i], TWG ? __ldg(tw1 + k1 * R2 + n2) : tw1[k1 * R2 + n2]);
            scr[g * S + k1 * (R2 + 1) + n2] = y;
        }
    }
    __syncwarp();
    // pass 2: R2-point DFTs over n2; all reads complete before the in-place natural-order store
    float2 w[U2][R2];
#pragma unroll
    for (int u = 0; u < U2; ++u) {
        const int q = lane + 32 * u;
        const int g = q / R1, k1 = q % R1;
#pragma unroll
        for (int n2 = 0; n2 < R2; ++n2) w[u][n2] = scr[g * S + k1 * (R2 + 1) + n2];
    }
    __syncwarp();
#pragma unroll
    for (int u = 0; u < U2; ++u) {
        const int q = lane + 32 * u;
        const int g = q / R1, k1 = q % R1;
        fft_regs<R2>(w[u]);
#pragma unroll
        for (int i = 0; i < R2; ++i) scr[g * S + k1 + R1 * brev<R2>(i)] = w[u][i];
    }
    __syncwarp();
}

// Real-FFT split: from packed bins A = Z[k], B = Z[(NC - k) % NC] and t = exp(-i pi k / NC) compute
// E = (A + conj B) / 2 and T = t * (A - conj B) / (2i); then X[k] = E + T and X[NC - k] = conj(E - T).
__device__ __forceinline__ void rfft_split(float2 A, float2 B, float2 t, float2 &E, float2 &T) {
    E = make_float2(0.5f * (A.x + B.x), 0.5f * (A.y - B.y));
    const float2 O = make_float2(0.5f * (A.y + B.y), -0.5f * (A.x - B.x));
    T = cmul(t, O);
}

// 10 * log10(x) for x > 0 through the MUFU lg2 path (absolute error ~1e-6 dB, far below the 1e-3 dB bar).
// __fmul_rn keeps the product from being contracted into an FMA by a caller (the epilogue's `x - ref` must be exactly 0
// at the maximum, common.py:224-225).
// Every caller passes x >= 1e-10 (a normal number), so the flush-to-zero form of lg2.approx returns the same bits as the
// default form while skipping its subnormal-input guard (3 of 6 instructions).
#ifndef AMT_DB_FTZ
#define AMT_DB_FTZ 1
#endif
__device__ __forceinline__ float db10(float x) {
#if AMT_DB_FTZ
    float l;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(x));
    return __fmul_rn(3.0102999566398120f, l);
#else
    return __fmul_rn(3.0102999566398120f, __log2f(x));
#endif
}

__device__ __forceinline__ void atomic_max_nonneg(float *addr, float v) {
    atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));  // valid ordering for non-negative floats
}

}  // namespace amtfeat
