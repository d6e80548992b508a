// umma_shift.cu -- two hardware questions behind the halo-tile convolution kernels (round 2), answered on a B200:
//
//  Q1  Can the A operand of tcgen05.mma start at a shared-memory address that is a multiple of 128 bytes but NOT of 1024
//      (the SWIZZLE_128B repeat)?  The halo-tile kernels keep ONE tile of activations (pixel rows of 128 bytes, written by
//      TMA with the 128-byte swizzle) and form the A operand of tap (dy, dx) by shifting the descriptor's start address by
//      (dy * pitch + dx) rows.  Two candidate encodings: base_offset = 0, or base_offset = (start >> 7) & 7.
//  Q2  Cycles per tcgen05.mma (M = 128, K = 16, fp16) as a function of N: is a narrow N (48) at the 128 * N / 256 floor
//      (B300_MICROARCH.md) or limited by the shared-memory reads of the A operand?
//
//  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/exp/umma_shift.bin scripts/exp/umma_shift.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../sc2-benchmark_b200/csrc/tc_common.cuh"

using namespace sc2::tc;

constexpr int kRows = 512;  // pixel rows in the A buffer (64 KB)
constexpr int kN = 64;

__device__ __forceinline__ uint64_t desc_with_base_offset(uint32_t saddr, uint32_t bo) {
    return make_smem_desc(saddr) | (static_cast<uint64_t>(bo & 7u) << 49);
}

// D[128, 64] = A[shift .. shift + 127][0..63] * B[0..63][0..63]^T, A rows written with the absolute-address 128-byte swizzle
__global__ void __launch_bounds__(128, 1) shift_kernel(float *out, int shift, int use_base_offset) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *a_buf = smem, *b_buf = smem + kRows * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(b_buf + kN * 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const uint32_t a_base = smem_u32(a_buf), b_base = smem_u32(b_buf);
    for (int i = threadIdx.x; i < kRows * 64; i += blockDim.x) {
        const int r = i / 64, k = i % 64;
        const float v = static_cast<float>(((r * 7 + k * 3) % 13) - 6);
        const uint32_t row_addr = a_base + r * 128;
        const uint32_t phys = row_addr + ((((k / 8) ^ ((row_addr >> 7) & 7)) << 4)) + (k % 8) * 2;
        *reinterpret_cast<__half *>(a_buf + (phys - a_base)) = __float2half(v);
    }
    for (int i = threadIdx.x; i < kN * 64; i += blockDim.x) {
        const int n = i / 64, k = i % 64;
        const float v = static_cast<float>(((n * 5 + k) % 11) - 5);
        const uint32_t row_addr = b_base + n * 128;
        const uint32_t phys = row_addr + ((((k / 8) ^ ((row_addr >> 7) & 7)) << 4)) + (k % 8) * 2;
        *reinterpret_cast<__half *>(b_buf + (phys - b_base)) = __float2half(v);
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, 64);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x < 32) {
        if (elect_one()) {
            const uint32_t a_addr = a_base + shift * 128;
            const uint32_t bo = use_base_offset ? ((a_addr >> 7) & 7u) : 0u;
            const uint64_t a_desc = desc_with_base_offset(a_addr, bo), b_desc = make_smem_desc(b_base);
            for (int k = 0; k < 4; ++k) umma_f16(tmem, a_desc + 2 * k, b_desc + 2 * k, make_idesc(kN), k > 0 ? 1u : 0u);
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < kN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        for (int e = 0; e < 32; ++e) out[(warp * 32 + lane) * kN + c0 + e] = __uint_as_float(v[e]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tcgen05_fence_after();
        tmem_dealloc(tmem, 64);
    }
}


// Q3  SWIZZLE_64B operands written by hand: rows of 64 bytes (32 fp16), 16-byte unit j of row r at ((j ^ ((r >> 1) & 3)) << 4),
//     descriptor layout type 4, 8-row groups 512 bytes apart.  D[128, 64] = A[128, 32] * B[64, 32]^T (two K steps).
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3ffff) >> 4);
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;
    return d;
}
__global__ void __launch_bounds__(128, 1) sw64_kernel(float *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *a_buf = smem, *b_buf = smem + 128 * 64;
    uint64_t *bar = reinterpret_cast<uint64_t *>(b_buf + kN * 64);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    for (int i = threadIdx.x; i < 128 * 32; i += blockDim.x) {
        const int r = i / 32, k = i % 32;
        const uint32_t off = r * 64 + ((((k / 8) ^ ((r >> 1) & 3)) << 4)) + (k % 8) * 2;
        *reinterpret_cast<__half *>(a_buf + off) = __float2half(static_cast<float>(((r * 7 + k * 3) % 13) - 6));
    }
    for (int i = threadIdx.x; i < kN * 32; i += blockDim.x) {
        const int n = i / 32, k = i % 32;
        const uint32_t off = n * 64 + ((((k / 8) ^ ((n >> 1) & 3)) << 4)) + (k % 8) * 2;
        *reinterpret_cast<__half *>(b_buf + off) = __float2half(static_cast<float>(((n * 5 + k) % 11) - 5));
    }
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, 64);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x < 32) {
        if (elect_one()) {
            const uint64_t a_desc = make_smem_desc_sw64(smem_u32(a_buf)), b_desc = make_smem_desc_sw64(smem_u32(b_buf));
            for (int k = 0; k < 2; ++k) umma_f16(tmem, a_desc + 2 * k, b_desc + 2 * k, make_idesc(kN), k > 0 ? 1u : 0u);
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < kN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, v);
        for (int e = 0; e < 32; ++e) out[(warp * 32 + lane) * kN + c0 + e] = __uint_as_float(v[e]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tcgen05_fence_after();
        tmem_dealloc(tmem, 64);
    }
}

// cycles per MMA: `iters` back-to-back tcgen05.mma (M = 128, N, K = 16) on resident operands; variant 1 alternates three
// (A, B) pairs like the split-fp16 passes
template <int N>
__global__ void __launch_bounds__(128, 1) rate_kernel(long long *cycles, int iters, int variant) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * kABytes + 2 * 256 * 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    for (int i = threadIdx.x; i < (2 * kABytes + 2 * 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x < 32) {
        long long t0 = 0, t1 = 0;
        if (elect_one()) {
            const uint32_t a0 = smem_u32(smem), a1 = a0 + kABytes, b0 = a0 + 2 * kABytes, b1 = b0 + 256 * 128;
            const uint64_t da0 = make_smem_desc(a0), da1 = make_smem_desc(a1), db0 = make_smem_desc(b0), db1 = make_smem_desc(b1);
            t0 = clock64();
            for (int i = 0; i < iters; ++i) {
                const int k = i & 3;
                if (variant == 0) {
                    umma_f16(tmem, da0 + 2 * k, db0 + 2 * k, make_idesc(N), 1u);
                } else {
                    umma_f16(tmem, da0 + 2 * k, db0 + 2 * k, make_idesc(N), 1u);
                    umma_f16(tmem + 256, da0 + 2 * k, db1 + 2 * k, make_idesc(N), 1u);
                    umma_f16(tmem + 256, da1 + 2 * k, db0 + 2 * k, make_idesc(N), 1u);
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0);
        t1 = clock64();
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;  // (lane 0 may not be the elected lane: its t0 is 0 then)
        long long tt0 = __shfl_sync(0xffffffffu, t0, 0);
        for (int l = 1; l < 32; ++l) {
            const long long c = __shfl_sync(0xffffffffu, t0, l);
            if (c > tt0) tt0 = c;
        }
        if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - tt0;
    } else {
        mbar_wait(bar, 0);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        tcgen05_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

#define CK(x)                                                                   \
    do {                                                                        \
        cudaError_t e_ = (x);                                                   \
        if (e_ != cudaSuccess) {                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return 1;                                                           \
        }                                                                       \
    } while (0)

template <int N>
static int run_rate(int grid, int variant) {
    const int smem = 2 * kABytes + 2 * 256 * 128 + 1024 + 64;
    CK(cudaFuncSetAttribute(rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    long long *d;
    CK(cudaMalloc(&d, sizeof(long long) * grid));
    const int iters = 4000;
    rate_kernel<N><<<grid, 128, smem>>>(d, iters, variant);
    CK(cudaDeviceSynchronize());
    rate_kernel<N><<<grid, 128, smem>>>(d, iters, variant);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(grid);
    CK(cudaMemcpy(h.data(), d, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (long long c : h) mx = c > mx ? c : mx;
    const int mmas = iters * (variant ? 3 : 1);
    printf("rate N=%3d grid=%3d variant=%d : %.1f cycles/MMA (floor 128*N/256 = %.1f)\n", N, grid, variant,
           static_cast<double>(mx) / mmas, 128.0 * N / 256.0);
    cudaFree(d);
    return 0;
}

int main() {
    const int smem = kRows * 128 + kN * 128 + 1024 + 64;
    CK(cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    float *d_out;
    CK(cudaMalloc(&d_out, 128 * kN * sizeof(float)));
    std::vector<float> h(128 * kN);
    const int shifts[] = {0, 8, 1, 3, 7, 9, 63, 65, 130, 197};
    for (int ubo = 0; ubo < 2; ++ubo)
        for (int s : shifts) {
            shift_kernel<<<1, 128, smem>>>(d_out, s, ubo);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
            int bad = 0;
            for (int m = 0; m < 128; ++m)
                for (int n = 0; n < kN; ++n) {
                    float want = 0.f;
                    for (int k = 0; k < 64; ++k)
                        want += static_cast<float>((((s + m) * 7 + k * 3) % 13) - 6) * static_cast<float>(((n * 5 + k) % 11) - 5);
                    if (want != h[m * kN + n]) ++bad;
                }
            printf("shift %3d base_offset=%s : %d / %d mismatches\n", s, ubo ? "(addr>>7)&7" : "0", bad, 128 * kN);
        }
    {
        CK(cudaFuncSetAttribute(sw64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        sw64_kernel<<<1, 128, 32768>>>(d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(h.data(), d_out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < kN; ++n) {
                float want = 0.f;
                for (int k = 0; k < 32; ++k) want += static_cast<float>(((m * 7 + k * 3) % 13) - 6) * static_cast<float>(((n * 5 + k) % 11) - 5);
                if (want != h[m * kN + n]) ++bad;
            }
        printf("SWIZZLE_64B manual layout : %d / %d mismatches\n", bad, 128 * kN);
    }
    if (getenv("SC2_EXP_SKIP_RATE")) return 0;
    cudaFree(d_out);
    for (int grid : {1, 148}) {
        if (run_rate<32>(grid, 0) || run_rate<48>(grid, 0) || run_rate<64>(grid, 0) || run_rate<96>(grid, 0) || run_rate<128>(grid, 0) ||
            run_rate<192>(grid, 0) || run_rate<256>(grid, 0))
            return 1;
        if (run_rate<48>(grid, 1) || run_rate<96>(grid, 1) || run_rate<128>(grid, 1) || run_rate<192>(grid, 1)) return 1;
    }
    return 0;
}
