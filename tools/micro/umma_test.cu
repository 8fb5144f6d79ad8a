// Bring-up test for the tcgen05 building block used by the GNN_BP4 tensor-core path:
//   D[128 x N] = A[128 x K] * W[K x N]   with A rows held by threads (one row per thread),
//   A staged in TMEM by tcgen05.st (hi / lo TF32 split), W in shared memory in the canonical
//   K-major no-swizzle UMMA layout (hi / lo), D read back with tcgen05.ld.
// nvcc -O2 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/micro/umma_test.cu -o /tmp/umma_test
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

__global__ void __launch_bounds__(128) umma_test(const float *A, const float *W, float *D, int K, int N, int Kpad, int Npad, int split) {
    extern __shared__ __align__(128) float sm[];
    float *Bh = sm, *Bl = sm + Npad * Kpad;
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    const int t = threadIdx.x, warp = t >> 5;
    if (warp == 0) tmem_alloc(&tslot, 256);
    for (int i = t; i < Npad * Kpad; i += 128) {
        const int n = i / Kpad, k = i - n * Kpad;
        const float w = (n < N && k < K) ? W[k * N + n] : 0.0f;
        const float hi = split ? __uint_as_float(__float_as_uint(w) & 0xffffe000u) : w;
        Bh[b_offset_floats(n, k, Kpad)] = hi;
        Bl[b_offset_floats(n, k, Kpad)] = w - hi;
    }
    if (t == 0) mbar_init(&bar, 1);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tbase = tslot;
    const uint32_t lane_base = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t colAh = 0, colAl = Kpad, colD = 2 * Kpad;
    // A row -> TMEM (hi, lo)
    for (int k0 = 0; k0 < Kpad; k0 += 8) {
        uint32_t hi[8], lo[8];
        for (int q = 0; q < 8; q++) {
            const float a = (k0 + q < K) ? A[t * K + k0 + q] : 0.0f;
            const float h = split ? __uint_as_float(__float_as_uint(a) & 0xffffe000u) : a;
            hi[q] = __float_as_uint(h); lo[q] = __float_as_uint(a - h);
        }
        tmem_st8(lane_base + colAh + k0, hi);
        tmem_st8(lane_base + colAl + k0, lo);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (t == 0) {
        tc_fence_after();
        const uint32_t idesc = idesc_tf32(Npad);
        const uint32_t bh = smem_u32(Bh), bl = smem_u32(Bl);
        for (int s = 0; s < Kpad / 8; s++) {
            const uint64_t dh = b_desc(bh + s * 256, Kpad), dl = b_desc(bl + s * 256, Kpad);
            umma_tf32_ts(tbase + colD, tbase + colAh + s * 8, dh, idesc, s > 0);
            if (split) {
                umma_tf32_ts(tbase + colD, tbase + colAl + s * 8, dh, idesc, 1);
                umma_tf32_ts(tbase + colD, tbase + colAh + s * 8, dl, idesc, 1);
            }
        }
        umma_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    for (int n0 = 0; n0 < Npad; n0 += 8) {
        uint32_t v[8];
        tmem_ld8(lane_base + colD + n0, v);
        tmem_wait_ld();
        for (int q = 0; q < 8; q++) if (n0 + q < N) D[t * N + n0 + q] = __uint_as_float(v[q]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, 256);
}

int main() {
    const int cases[][2] = {{24, 48}, {20, 40}, {40, 20}, {60, 40}, {20, 80}, {41, 40}};
    int bad = 0;
    for (auto &c : cases) {
        const int K = c[0], N = c[1], Kpad = (K + 7) & ~7, Npad = (N + 15) & ~15;
        std::vector<float> A(128 * K), W(K * N), D(128 * N);
        srand(K * 100 + N);
        for (auto &x : A) x = (float)rand() / RAND_MAX * 4.0f - 2.0f;
        for (auto &x : W) x = (float)rand() / RAND_MAX * 2.0f - 1.0f;
        float *dA, *dW, *dD;
        CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dW, W.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4));
        CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
        for (int split = 0; split < 2; split++) {
            CK(cudaMemset(dD, 0, D.size() * 4));
            const size_t smem = (size_t)2 * Npad * Kpad * 4;
            CK(cudaFuncSetAttribute(umma_test, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            umma_test<<<1, 128, smem>>>(dA, dW, dD, K, N, Kpad, Npad, split);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
            double maxerr = 0, maxref = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < N; n++) {
                    double r = 0;
                    for (int k = 0; k < K; k++) r += (double)A[m * K + k] * (double)W[k * N + n];
                    maxerr = fmax(maxerr, fabs(r - (double)D[m * N + n]));
                    maxref = fmax(maxref, fabs(r));
                }
            printf("K=%d N=%d split=%d  max |err| = %.3e  (max |ref| = %.2f)\n", K, N, split, maxerr, maxref);
            if (maxerr > (split ? 2e-5 : 2e-2)) bad++;
        }
        cudaFree(dA); cudaFree(dW); cudaFree(dD);
    }
    printf(bad ? "FAILED (%d)\n" : "OK\n", bad);
    return bad ? 1 : 0;
}
