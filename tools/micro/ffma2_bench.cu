// Micro-benchmark (development aid): does fma.rn.f32x2 relieve the issue port on B200?
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s\n", cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_ffma(float *out, int iters) {
    float v[8]; const float a = 0.999f + 1e-6f * threadIdx.x, c = 1e-3f;
    for (int j = 0; j < 8; j++) v[j] = 0.5f + 0.01f * j;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __fmaf_rn(v[j], a, c);
    float s = 0; for (int j = 0; j < 8; j++) s += v[j];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_ffma2(float *out, int iters) {
    float2 v[8]; const float2 a = make_float2(0.999f + 1e-6f * threadIdx.x, 0.998f), c = make_float2(1e-3f, 2e-3f);
    for (int j = 0; j < 8; j++) v[j] = make_float2(0.5f + 0.01f * j, 0.4f);
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < 8; j++) v[j] = __ffma2_rn(v[j], a, c);
    float s = 0; for (int j = 0; j < 8; j++) s += v[j].x + v[j].y;
    if (s == 123.456f) out[0] = s;
}
// 4 FFMA(2) + 4 integer ALU ops per step: is the total bound by issue slots or by pipes?
__global__ void k_mix(float *out, int iters) {
    float v[4]; unsigned u[4]; const float a = 0.999f + 1e-6f * threadIdx.x, c = 1e-3f;
    for (int j = 0; j < 4; j++) { v[j] = 0.5f + 0.01f * j; u[j] = threadIdx.x + j; }
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { v[j] = __fmaf_rn(v[j], a, c); v[j] = __fmaf_rn(v[j], a, c); u[j] = (u[j] ^ (u[j] >> 3)) + 0x9E3779B9u; }
    float s = 0; for (int j = 0; j < 4; j++) s += v[j] + u[j];
    if (s == 123.456f) out[0] = s;
}
__global__ void k_mix2(float *out, int iters) {
    float2 v[4]; unsigned u[4]; const float2 a = make_float2(0.999f + 1e-6f * threadIdx.x, 0.998f), c = make_float2(1e-3f, 2e-3f);
    for (int j = 0; j < 4; j++) { v[j] = make_float2(0.5f + 0.01f * j, 0.3f); u[j] = threadIdx.x + j; }
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) { v[j] = __ffma2_rn(v[j], a, c); u[j] = (u[j] ^ (u[j] >> 3)) + 0x9E3779B9u; }
    float s = 0; for (int j = 0; j < 4; j++) s += v[j].x + v[j].y + u[j];
    if (s == 123.456f) out[0] = s;
}
template <typename K> float timeit(K k, float *d, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148 * 8, 256>>>(d, 16);
    float best = 1e30f;
    for (int r = 0; r < 3; r++) { cudaEventRecord(e0); k<<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
int main() {
    float *d; CK(cudaMalloc(&d, 4));
    const int iters = 50000; const double thr = 148.0 * 8 * 256;
    float t1 = timeit(k_ffma, d, iters), t2 = timeit(k_ffma2, d, iters), t3 = timeit(k_mix, d, iters), t4 = timeit(k_mix2, d, iters);
    printf("FFMA : %.1f G fma-lanes/s (%.2f ms)\n", thr * 8 * iters / t1 * 1e-6, t1);
    printf("FFMA2: %.1f G fma-lanes/s (%.2f ms)  -> %.2fx per instruction\n", thr * 16 * iters / t2 * 1e-6, t2, 2 * t1 / t2);
    printf("mix  (8 FFMA + 12 int ops): %.2f ms\n", t3);
    printf("mix2 (4 FFMA2 + 12 int ops): %.2f ms  speedup %.2fx\n", t4, t3 / t4);
    CK(cudaGetLastError());
    return 0;
}
