// Hardware probe for the tcgen05 building blocks in ../tc_common.cuh (development tool, not product).
//   tc_probe <mode> <N> <K>      mode: 0 = SS cta_group::1, 1 = TS cta_group::1,
//                                      2 = TS cta_group::2 (M=256 over a CTA pair), 3 = SS cta_group::2
// Computes D[M][N] = A[M][K] * B[N][K]^T with small-integer bf16 inputs (exact in fp32) and compares
// every element with the host result.  Built with a spin limit so a protocol bug traps instead of hanging.
#define UMNN_TC_SPIN_LIMIT 200000000LL
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../tc_common.cuh"

using namespace umnn::tc;

#define CK(x)                                                                         \
    do {                                                                              \
        cudaError_t e = (x);                                                          \
        if (e != cudaSuccess) {                                                       \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                  \
        }                                                                             \
    } while (0)

constexpr int kAcol0 = 256;  // TMEM column where the A operand starts (accumulator at column 0)

// B (and A in SS mode) in shared memory: core matrices [k/8][row/8] of 128 bytes, row r at +16 r
__device__ __forceinline__ uint32_t core_offset(int row, int k, int rows) {
    return (uint32_t)(((k >> 3) * (rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k & 7) * 2);
}

template <int CG, bool A_TMEM>
__global__ void __launch_bounds__(128) probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                    float* __restrict__ D, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_holder;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
    const int pair = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int Mtile = 128 * CG;
    const int Nloc = N / CG;                       // rows of B held by this CTA
    uint8_t* Bs = smem;                            // Nloc x K bf16
    uint8_t* As = smem + (size_t)Nloc * K * 2;     // 128 x K bf16 (SS only)
    const __nv_bfloat16* Ag = A + ((size_t)pair * Mtile + rank * 128) * K;   // this CTA's 128 rows
    const __nv_bfloat16* Bg = B + (size_t)rank * Nloc * K;                   // this CTA's half of N

    for (int i = tid; i < Nloc * K; i += 128) {
        const int n = i / K, k = i % K;
        *reinterpret_cast<__nv_bfloat16*>(Bs + core_offset(n, k, Nloc)) = Bg[(size_t)n * K + k];
    }
    if (!A_TMEM)
        for (int i = tid; i < 128 * K; i += 128) {
            const int r = i / K, k = i % K;
            *reinterpret_cast<__nv_bfloat16*>(As + core_offset(r, k, 128)) = Ag[(size_t)r * K + k];
        }
    fence_proxy_async_smem();

    if (warp == 0) tmem_alloc<CG>(&tmem_holder, 512);
    if (tid == 0) {
        mbar_init(&done_bar, 1);
        fence_mbar_init();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_holder;

    if (A_TMEM) {
        // thread tid owns row tid of this CTA's A tile: 8 packed columns per K=16 block
        for (int kb = 0; kb < K / 16; ++kb) {
            uint32_t v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float e = __bfloat162float(Ag[(size_t)tid * K + kb * 16 + 2 * j]);
                const float o = __bfloat162float(Ag[(size_t)tid * K + kb * 16 + 2 * j + 1]);
                v[j] = pack_bf16x2(e, o);
            }
            tmem_st8(tmem_addr(tbase, warp * 32, kAcol0 + kb * 8), v);
        }
        tmem_st_wait();
        tc_fence_before_sync();
        __syncthreads();
        if (CG == 2) cluster_sync_all();
        tc_fence_after_sync();
    }

    if (rank == 0 && warp == 0 && lane == 0) {
        const uint32_t idesc = make_idesc_bf16_f32(Mtile, N);
        for (int kb = 0; kb < K / 16; ++kb) {
            const uint64_t bdesc = make_smem_desc(smem_u32(Bs) + kb * 2 * (Nloc / 8) * 128, (Nloc / 8) * 128, 128);
            if (A_TMEM) {
                mma_ts<CG>(tbase, tbase + kAcol0 + kb * 8, bdesc, idesc, kb > 0);
            } else {
                const uint64_t adesc = make_smem_desc(smem_u32(As) + kb * 2 * 16 * 128, 16 * 128, 128);
                mma_ss<CG>(tbase, adesc, bdesc, idesc, kb > 0);
            }
        }
        mma_commit<CG>(&done_bar);
    }
    mbar_wait(&done_bar, 0, 1);
    tc_fence_after_sync();

    float* Dg = D + ((size_t)pair * Mtile + rank * 128 + tid) * N;
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_addr(tbase, warp * 32, c), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c + j < N) Dg[c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) tmem_dealloc<CG>(tbase, 512);
}

// ---- MN-major operands: D[M][N] = sum_r At[r][M] * Bt[r][N]  (both operands stored [K=r][MN] row-major in
//      global; staged as core matrices [k8][mn8][8 r][8 mn] so that each 16-byte granule is 8 consecutive MN
//      elements of one K row), idesc a_major = b_major = 1
__device__ __forceinline__ uint32_t core_offset_mn(int r, int mn, int mn_total) {
    return (uint32_t)(((r >> 3) * (mn_total >> 3) + (mn >> 3)) * 128 + (r & 7) * 16 + (mn & 7) * 2);
}

template <int CG>
__global__ void __launch_bounds__(128) probe_mn_kernel(const __nv_bfloat16* __restrict__ At, const __nv_bfloat16* __restrict__ Bt,
                                                       float* __restrict__ D, int N, int K) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t done_bar;
    __shared__ uint32_t tmem_holder;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0;
    const int M = 128 * CG;
    const int Nloc = N / CG;
    uint8_t* As = smem;                                // 128 (M half) x K
    uint8_t* Bs = smem + (size_t)128 * K * 2;          // Nloc x K
    for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / 128, m = i % 128;
        *reinterpret_cast<__nv_bfloat16*>(As + core_offset_mn(r, m, 128)) = At[(size_t)r * M + rank * 128 + m];
    }
    for (int i = tid; i < Nloc * K; i += 128) {
        const int r = i / Nloc, n = i % Nloc;
        *reinterpret_cast<__nv_bfloat16*>(Bs + core_offset_mn(r, n, Nloc)) = Bt[(size_t)r * N + rank * Nloc + n];
    }
    fence_proxy_async_smem();
    if (warp == 0) tmem_alloc<CG>(&tmem_holder, 512);
    if (tid == 0) { mbar_init(&done_bar, 1); fence_mbar_init(); }
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    tc_fence_after_sync();
    const uint32_t tbase = tmem_holder;
    if (rank == 0 && warp == 0 && lane == 0) {
        const uint32_t idesc = make_idesc_bf16_f32(M, N) | (1u << 15) | (1u << 16);
        for (int kb = 0; kb < K / 16; ++kb) {
            // per K=16 block: 2 k8 groups; k8 stride = (MN/8)*128 bytes, mn8 stride = 128 bytes
            const uint64_t adesc = make_smem_desc(smem_u32(As) + kb * 2 * 16 * 128, 16 * 128, 128);
            const uint64_t bdesc = make_smem_desc(smem_u32(Bs) + kb * 2 * (Nloc / 8) * 128, (Nloc / 8) * 128, 128);
            mma_ss<CG>(tbase, adesc, bdesc, idesc, kb > 0);
        }
        mma_commit<CG>(&done_bar);
    }
    mbar_wait(&done_bar, 0, 1);
    tc_fence_after_sync();
    float* Dg = D + ((size_t)rank * 128 + tid) * N;
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem_addr(tbase, warp * 32, c), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
            if (c + j < N) Dg[c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before_sync();
    __syncthreads();
    if (CG == 2) cluster_sync_all();
    if (warp == 0) tmem_dealloc<CG>(tbase, 512);
}

template <int CG>
static int run_mn(int N, int K, int variant) {
    const int M = 128 * CG;
    std::vector<__nv_bfloat16> hA((size_t)K * M), hB((size_t)K * N);
    std::vector<float> fA((size_t)K * M), fB((size_t)K * N);
    srand(4321);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 9 - 4); hA[i] = __float2bfloat16(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 9 - 4); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, (size_t)M * N * 4));
    const size_t smem = (size_t)128 * K * 2 + (size_t)(N / CG) * K * 2 + 1024;
    auto kern = probe_mn_kernel<CG>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CG);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    (void)variant;
    CK(cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16*)dA, (const __nv_bfloat16*)dB, dD, N, K));
    CK(cudaDeviceSynchronize());
    std::vector<float> hD((size_t)M * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0; int shown = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < K; ++k) ref += fA[(size_t)k * M + m] * fB[(size_t)k * N + n];
            const float got = hD[(size_t)m * N + n];
            if (!(got == ref)) { ++bad; if (shown < 8) { printf("  mismatch m=%d n=%d got=%g ref=%g\n", m, n, got, ref); ++shown; } }
        }
    printf("probe MN-major CG=%d N=%d K=%d M=%d: %ld / %ld mismatches -> %s\n", CG, N, K, M, bad, (long)M * N, bad ? "FAIL" : "PASS");
    return bad ? 1 : 0;
}

template <int CG, bool A_TMEM>
static int run(int N, int K, int pairs) {
    const int M = 128 * CG * pairs;
    std::vector<__nv_bfloat16> hA((size_t)M * K), hB((size_t)N * K);
    std::vector<float> fA((size_t)M * K), fB((size_t)N * K);
    srand(1234);
    for (size_t i = 0; i < hA.size(); ++i) { fA[i] = (float)(rand() % 9 - 4); hA[i] = __float2bfloat16(fA[i]); }
    for (size_t i = 0; i < hB.size(); ++i) { fB[i] = (float)(rand() % 9 - 4); hB[i] = __float2bfloat16(fB[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dD;
    CK(cudaMalloc(&dA, hA.size() * 2));
    CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dD, (size_t)M * N * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xFF, (size_t)M * N * 4));
    const size_t smem = (size_t)(N / CG) * K * 2 + (A_TMEM ? 0 : (size_t)128 * K * 2) + 1024;
    auto kern = probe_kernel<CG, A_TMEM>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(CG * pairs);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kern, (const __nv_bfloat16*)dA, (const __nv_bfloat16*)dB, dD, N, K));
    CK(cudaDeviceSynchronize());
    std::vector<float> hD((size_t)M * N);
    CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
    long bad = 0;
    int shown = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            float ref = 0.f;
            for (int k = 0; k < K; ++k) ref += fA[(size_t)m * K + k] * fB[(size_t)n * K + k];
            const float got = hD[(size_t)m * N + n];
            if (!(got == ref)) {
                ++bad;
                if (shown < 8) { printf("  mismatch m=%d n=%d got=%g ref=%g\n", m, n, got, ref); ++shown; }
            }
        }
    printf("probe CG=%d A_TMEM=%d N=%d K=%d M=%d: %ld / %ld mismatches -> %s\n", CG, (int)A_TMEM, N, K, M, bad,
           (long)M * N, bad ? "FAIL" : "PASS");
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad ? 1 : 0;
}

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int N = argc > 2 ? atoi(argv[2]) : 208;
    const int K = argc > 3 ? atoi(argv[3]) : 64;
    const int pairs = argc > 4 ? atoi(argv[4]) : 2;
    switch (mode) {
        case 0: return run<1, false>(N, K, pairs);
        case 1: return run<1, true>(N, K, pairs);
        case 2: return run<2, true>(N, K, pairs);
        case 3: return run<2, false>(N, K, pairs);
        case 4: return run_mn<1>(N, K, pairs);
        case 5: return run_mn<2>(N, K, pairs);
    }
    return 3;
}
