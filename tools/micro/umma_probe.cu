// Characterisation probe for the arithmetic of ONE tcgen05.mma kind::tf32 (M = 128, N = 16, K = 8): D_out = A * B + D_in on
// TF32-exact inputs with widely spread exponents; the vectors are dumped for tools/micro/umma_model.py, which looks for
// an integer model (alignment, truncation, rounding) that reproduces the hardware bit for bit.
// nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/umma_probe.cu -o /tmp/umma_probe
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t v[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t v[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc], kind::tf32, issued by ONE thread
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!ok);
}

// canonical K-major no-swizzle tile of an [Npad x Kpad] operand (32-bit elements): 8 x 16 B core matrices,
// adjacent along K (LBO = 128 B), 8-row groups SBO = Kpad / 4 * 128 B apart
__host__ __device__ inline int b_offset_floats(int n, int k, int Kpad) {
    return (n >> 3) * (Kpad / 4) * 32 + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
}
__device__ __forceinline__ uint64_t b_desc(uint32_t saddr, int Kpad) {
    const uint64_t lbo = 128 >> 4, sbo = (uint64_t)((Kpad / 4) * 128) >> 4;
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46);
}
__host__ __device__ inline uint32_t idesc_tf32(int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}


__global__ void __launch_bounds__(128) umma_probe(const float *A, const float *B, const float *Din, float *Dout, int trials) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int t = threadIdx.x, warp = t >> 5;
    if (warp == 0) tmem_alloc(&tslot, 32);
    if (t == 0) mbar_init(&bar, 1);
    __syncthreads();
    const uint32_t tbase = tslot;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    for (int tr = 0; tr < trials; tr++) {
        for (int i = t; i < 16 * 8; i += 128) {                 // B[k][n] -> canonical K-major tile
            const int n = i / 8, k = i - n * 8;
            sm[b_offset_floats(n, k, 8)] = B[(tr * 8 + k) * 16 + n];
        }
        uint32_t a[8], d0[8], d1[8];
        for (int q = 0; q < 8; q++) a[q] = __float_as_uint(A[(tr * 128 + t) * 8 + q]);
        for (int q = 0; q < 8; q++) { d0[q] = __float_as_uint(Din[(tr * 128 + t) * 16 + q]); d1[q] = __float_as_uint(Din[(tr * 128 + t) * 16 + 8 + q]); }
        tmem_st8(lane_base + 0, a);
        tmem_st8(lane_base + 16, d0);
        tmem_st8(lane_base + 24, d1);
        tmem_wait_st();
        fence_async_smem();
        tc_fence_before();
        __syncthreads();
        if (t == 0) {
            tc_fence_after();
            umma_tf32_ts(tbase + 16, tbase + 0, b_desc(smem_u32(sm), 8), idesc_tf32(16), 1);
            umma_commit(&bar);
        }
        mbar_wait(&bar, tr & 1);
        tc_fence_after();
        uint32_t v[8];
        tmem_ld8(lane_base + 16, v); tmem_wait_ld();
        for (int q = 0; q < 8; q++) Dout[(tr * 128 + t) * 16 + q] = __uint_as_float(v[q]);
        tmem_ld8(lane_base + 24, v); tmem_wait_ld();
        for (int q = 0; q < 8; q++) Dout[(tr * 128 + t) * 16 + 8 + q] = __uint_as_float(v[q]);
        tc_fence_before();
        __syncthreads();
    }
    if (warp == 0) tmem_dealloc(tbase, 32);
}

static float tf32_rand(int emin, int emax) {            // sign * 11-bit significand * 2^e
    const int e = emin + rand() % (emax - emin + 1);
    const int m = 1024 + rand() % 1024;
    const float v = ldexpf((float)m, e - 10);
    return (rand() & 1) ? -v : v;
}
static float f32_rand(int emin, int emax) {
    const int e = emin + rand() % (emax - emin + 1);
    const int m = (1 << 23) + (rand() & 0x7fffff);
    const float v = ldexpf((float)m, e - 23);
    return (rand() & 1) ? -v : v;
}

int main(int argc, char **argv) {
    const int trials = 256;
    const char *out = argc > 1 ? argv[1] : "umma_probe.bin";
    std::vector<float> A(trials * 128 * 8), B(trials * 8 * 16), Din(trials * 128 * 16), Dout(trials * 128 * 16);
    srand(12345);
    for (int tr = 0; tr < trials; tr++) {
        const int mode = tr % 4;        // 0: narrow exponents, D = 0; 1: wide exponents, D = 0; 2: narrow + D; 3: wide + D
        const int span = (mode & 1) ? 14 : 2;
        const bool full = (tr / 4) % 2 == 1;       // every other block of 4 trials: operands with all 24 significand bits set at random
        for (int i = 0; i < 128 * 8; i++) A[tr * 1024 + i] = full ? f32_rand(-span, span) : tf32_rand(-span, span);
        for (int i = 0; i < 8 * 16; i++) B[tr * 128 + i] = full ? f32_rand(-span, span) : tf32_rand(-span, span);
        for (int i = 0; i < 128 * 16; i++) Din[tr * 2048 + i] = (mode & 2) ? f32_rand(-2 * span, 2 * span + 3) : 0.0f;
        if (tr % 16 == 9) for (int i = 0; i < 128 * 8; i++) if (rand() % 3 == 0) A[tr * 1024 + i] = 0.0f;   // sparse rows
    }
    float *dA, *dB, *dDi, *dDo;
    CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dDi, Din.size() * 4)); CK(cudaMalloc(&dDo, Dout.size() * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dDi, Din.data(), Din.size() * 4, cudaMemcpyHostToDevice));
    umma_probe<<<1, 128, 1024>>>(dA, dB, dDi, dDo, trials);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(Dout.data(), dDo, Dout.size() * 4, cudaMemcpyDeviceToHost));
    FILE *f = fopen(out, "wb");
    const int hdr[4] = {trials, 128, 8, 16};
    fwrite(hdr, 4, 4, f);
    fwrite(A.data(), 4, A.size(), f); fwrite(B.data(), 4, B.size(), f); fwrite(Din.data(), 4, Din.size(), f); fwrite(Dout.data(), 4, Dout.size(), f);
    fclose(f);
    // sanity: compare with the double-precision dot product
    double maxrel = 0;
    for (int tr = 0; tr < trials; tr++)
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < 16; n++) {
                double r = Din[(tr * 128 + m) * 16 + n], s = fabs(r);
                for (int k = 0; k < 8; k++) { const double p = (double)A[(tr * 128 + m) * 8 + k] * B[(tr * 8 + k) * 16 + n]; r += p; s += fabs(p); }
                maxrel = fmax(maxrel, fabs(r - Dout[(tr * 128 + m) * 16 + n]) / (s + 1e-300));
            }
    printf("wrote %s; max |err| / sum |terms| = %.3e\n", out, maxrel);
    return 0;
}
