// Micro-benchmark: scalar FFMA vs packed FFMA2 / FADD2 issue throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{.reg .b64 ra, rb, rc, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n mov.b64 rc, {%6,%7};\n"
        " fma.rn.f32x2 rd, ra, rb, rc;\n mov.b64 {%0,%1}, rd;}\n"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    float2 d;
    asm("{.reg .b64 ra, rb, rd;\n mov.b64 ra, {%2,%3};\n mov.b64 rb, {%4,%5};\n add.rn.f32x2 rd, ra, rb;\n mov.b64 {%0,%1}, rd;}\n"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}

constexpr int CH = 8;
template <int MODE> __global__ void __launch_bounds__(256) k(float2 *out, int iters, float2 m, float2 a) {
    float2 acc[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) acc[c] = make_float2(threadIdx.x + c, threadIdx.x - c);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int c = 0; c < CH; ++c) {
            if (MODE == 0) { acc[c].x = fmaf(acc[c].x, m.x, a.x); acc[c].y = fmaf(acc[c].y, m.y, a.y); }
            else if (MODE == 1) acc[c] = ffma2(acc[c], m, a);
            else if (MODE == 2) { acc[c].x += a.x; acc[c].y += a.y; }
            else acc[c] = fadd2(acc[c], a);
        }
    }
    float2 s = make_float2(0, 0);
#pragma unroll
    for (int c = 0; c < CH; ++c) { s.x += acc[c].x; s.y += acc[c].y; }
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, float2 *d) {
    const int iters = 4096, grid = 148 * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(d, iters, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f));
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, 256>>>(d, iters, make_float2(1.0001f, 0.9999f), make_float2(0.5f, 0.25f));
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 5.0 * grid * 256.0 * iters * CH * 2;  // scalar element-ops
    printf("%-8s %8.3f ms  %7.2f T elem-op/s  (per SM per clk @1.965GHz: %.1f)\n", name, ms, ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float2 *d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float2));
    run<0>("FFMA", d); run<1>("FFMA2", d); run<2>("FADD", d); run<3>("FADD2", d);
    return 0;
}
