// conv_ga_first.cu -- first layer of the analysis transform g_a + its GDN1, back to back in ONE kernel (fp32-grade, tcgen05).
//
// Conv2d(3 -> C, k5, s2, p2) on the NCHW image followed by GDN1(C) (sc2bench/models/layer.py:476-478).  Round 1 ran them as two
// launches with a 2.4 GB round trip of x1 through HBM (0.77 + 0.72 ms per 256 images); both are HBM-bound, so the fused kernel
// writes only y1: image in (fp32, or uint8 with the ToTensor / Normalize of the data loader as a 3 x 256 look-up table), y1 out as
// split fp16 parity planes, which is what the stride-2 halo kernel (conv_ga_halo.cu) reads next.
//
//   patches   One TMA box {4 tw + 4, 4 th + 1, 3} of the image per tile (zero fill = the conv's padding) lands in shared memory;
//             four producer warps (thread = output pixel) gather the 5x5x3 patch from it with conflict-free 16-byte loads, split
//             every value into (hi, lo) fp16 and write the A tile in the UMMA K-major layout: K = 75 -> 64 (SWIZZLE_128B rows)
//             + 16 (SWIZZLE_64B rows).  Round 1 gathered from global memory, one thread's 15 row segments at a time.
//   GEMMs     stacked weights [hi; lo] resident in shared memory: one N = 2n MMA gives hi.hi | hi.lo, one N = n MMA adds lo.hi
//             (see conv_ga_halo.cu).  The conv accumulator is drained once: x stays in registers, |x| (hi, lo) becomes the A
//             operand of the gamma GEMM (n = 96: 64 + 32 channels), y = x / (beta + gamma.|x|) is split and TMA-stored.
//   roles     warp 0: patch TMA + tile scheduler; warp 1: MMA issuer; warps 2..17: epilogue; warps 18..21: im2col producers.
#include "tc_common.cuh"

namespace sc2 {
namespace gaf {

using namespace sc2::tc;

constexpr int kEpiWarps = 16;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kParts = kEpiWarps / 4;  // column parts of the epilogue (x 4 TMEM lane quarters)
constexpr int kThreads = (2 + kEpiWarps + 4) * 32;  // warp 0 TMA, warp 1 MMA, epilogue warps, 4 im2col producer warps
constexpr float kLoScale = 2048.0f;
constexpr float kLoInv = 1.0f / 2048.0f;
constexpr int kK = 80;            // 3 * 5 * 5 = 75 padded to 5 K steps
constexpr int kA1Bytes = 128 * 64;  // second K chunk of the A tile: 128 rows of 64 bytes (SWIZZLE_64B)

struct Params {
    int batch, h_in, w_in;
    int hp, wp;           // parity-plane geometry = output size / 2
    int tiles_x, tiles_y, tw, th;
    int box_w, box_h;     // patch box: 4 tw + 4 columns (uint8: rounded up to 16), 4 th + 1 rows
    int patch_bytes;      // one patch buffer, multiple of 1024
    int c_out, out_c, stage_c, stage_plane;
    int off_gamma, off_a, off_ag, off_patch, off_bar, off_lut;
    const float *beta;
    const float *lut;     // uint8 input: [3][256] value of each byte after ToTensor + Normalize; nullptr: fp32 input
    int *tile_counter;
    TraceSink trace;
};

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fast_rcp(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return fmaf(r, fmaf(-v, r, 1.0f), r);
}

__device__ __forceinline__ void split8(const float *f, uint4 &h, uint4 &l) {
    uint32_t *hw = reinterpret_cast<uint32_t *>(&h), *lw = reinterpret_cast<uint32_t *>(&l);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 hh = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn((f[2 * e] - back.x) * kLoScale, (f[2 * e + 1] - back.y) * kLoScale);
        hw[e] = *reinterpret_cast<const uint32_t *>(&hh);
        lw[e] = *reinterpret_cast<const uint32_t *>(&ll);
    }
}

// K-major operand descriptor with the 64-byte swizzle: rows of 64 bytes, 8-row groups 512 bytes apart
__device__ __forceinline__ uint64_t make_smem_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3ffff) >> 4);
    d |= static_cast<uint64_t>(512 >> 4) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;  // SWIZZLE_64B
    return d;
}

// N: output channels padded to a multiple of 16; U8: uint8 image + look-up table
template <int N, bool U8>
__global__ void __launch_bounds__(kThreads, 1)
ga_first_gdn_kernel(const __grid_constant__ CUtensorMap map_img, const __grid_constant__ CUtensorMap map_w0,
                    const __grid_constant__ CUtensorMap map_w1, const __grid_constant__ CUtensorMap map_g0,
                    const __grid_constant__ CUtensorMap map_g1, const __grid_constant__ CUtensorMap map_o_hi,
                    const __grid_constant__ CUtensorMap map_o_lo, const __grid_constant__ Params p) {
    constexpr int kW0 = 2 * N * 128, kW1 = 2 * N * 64;  // stacked weights: K chunk 0 (64 wide, SW128), chunk 1 (32 wide, SW64)
    constexpr int kGC0 = N >= 64 ? 64 : N;              // gamma GEMM: channels in chunk 0 / chunk 1
    constexpr int kGC1 = N - kGC0;                      // 0, 16 or 32
    constexpr int kSub = N / 8;                               // 8-channel sub-units of the epilogue
    constexpr int kSubMax = (kSub + kParts - 1) / kParts;      // at most this many per thread
    static_assert(N % 16 == 0 && N >= 16 && N <= 96, "N");
    extern __shared__ uint8_t smem_raw[];
    // (aligned by OFFSET, not through an integer cast: the compiler keeps the shared address space -> LDS / STS, not generic LD / ST)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *s_w = smem;                    // [W chunk 0][W chunk 1]
    uint8_t *s_gamma = smem + p.off_gamma;  // [gamma chunk 0][gamma chunk 1]
    uint8_t *s_a = smem + p.off_a;          // A tile: [hi 0 (16 KB)][lo 0][hi 1 (8 KB)][lo 1]
    uint8_t *s_ag = smem + p.off_ag;        // |x| operand, same layout; aliased by the output staging tile
    uint8_t *s_patch = smem + p.off_patch;  // 2 patch buffers
    uint64_t *patch_full = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *patch_empty = patch_full + 2;
    uint64_t *a_full = patch_empty + 2;
    uint64_t *a_empty = a_full + 1;
    uint64_t *acc_full = a_empty + 1;
    uint64_t *acc_empty = acc_full + 1;
    uint64_t *ag_full = acc_empty + 1;
    uint64_t *g_full = ag_full + 1;
    uint64_t *w_full = g_full + 1;
    uint64_t *stage_free = w_full + 1;  // the bulk stores of the previous tile have read the staging tile (= the |x| operand's memory)
    uint64_t *y_done = stage_free + 1;  // every epilogue thread has written its part of the staging tile
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(y_done + 1);
    float *s_beta = reinterpret_cast<float *>(tmem_slot + 2);  // 16-byte aligned (13 barriers = 104 bytes, + 8)
    const float *s_lut = reinterpret_cast<const float *>(smem + p.off_lut);
    TileSched sched;
    sched.bind(reinterpret_cast<uint8_t *>(s_beta + N), p.tile_counter, p.tiles_x * p.tiles_y * p.batch * 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rows = p.tw * p.th;
    const int tiles_xy = p.tiles_x * p.tiles_y;
    const unsigned long long trace_t0 = p.trace.buf ? trace_now() : 0ull;
    int trace_tiles = 0;

    if (threadIdx.x < N) s_beta[threadIdx.x] = static_cast<int>(threadIdx.x) < p.c_out ? __ldg(p.beta + threadIdx.x) : 1.0f;
    if (U8)
        for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) reinterpret_cast<float *>(smem + p.off_lut)[i] = __ldg(p.lut + i);
    // chunk 1 of the A tile holds 16 valid K values per 64-byte row; the other half must be zero (its weights are zero, but
    // 0 * NaN from uninitialised memory would not be) -- written once, the producers only ever touch the first half
    for (int i = threadIdx.x; i < 2 * kA1Bytes / 16; i += blockDim.x) reinterpret_cast<uint4 *>(s_a + 2 * kABytes)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        sched.init(1 + kEpiWarps + 4);  // consumers: MMA warp, epilogue warps, 4 producer warps
        tma_prefetch_desc(&map_img);
        tma_prefetch_desc(&map_o_hi);
        tma_prefetch_desc(&map_o_lo);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&patch_full[s], 1);
            mbar_init(&patch_empty[s], 128);
        }
        mbar_init(a_full, 128);
        mbar_init(a_empty, 1);
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, kEpiThreads);
        mbar_init(ag_full, kEpiThreads);
        mbar_init(g_full, 1);
        mbar_init(w_full, 1);
        mbar_init(stage_free, 1);
        mbar_init(y_done, kEpiThreads);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    fence_proxy_async();  // the zero fill above is read by the MMA (async proxy)
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t kGammaCol = 2 * N;

    if (warp == 0) {
        // =============================== weights (once), patch TMA, tile scheduler ===============================
        if (elect_one()) {
            mbar_expect_tx(w_full, static_cast<uint32_t>(kW0 + kW1 + kW0 + (kGC1 ? kW1 : 0)));
            tma_load_2d(&map_w0, w_full, s_w, 0, 0);
            tma_load_2d(&map_w1, w_full, s_w + kW0, 64, 0);
            tma_load_2d(&map_g0, w_full, s_gamma, 0, 0);
            if (kGC1) tma_load_2d(&map_g1, w_full, s_gamma + kW0, 64, 0);
            int tile = sched.claim(0);
            for (uint32_t qn = 0;; ++qn) {
                sched.publish(qn, tile);
                if (tile < 0) break;
                const int next_tile = sched.claim(qn + 1);
                const uint32_t s = qn & 1u, ph = (qn >> 1) & 1u;
                const int sp = tile % tiles_xy, img = tile / tiles_xy;  // img = b * 4 + py * 2 + px
                const int b = img >> 2, py = (img >> 1) & 1, px = img & 1;
                const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
                mbar_wait(&patch_empty[s], ph ^ 1u);
                mbar_expect_tx(&patch_full[s], static_cast<uint32_t>(p.box_w * p.box_h * 3 * (U8 ? 1 : 4)));
                // TMA wants the box to start on a 16-byte boundary of the innermost dimension: round the first column down
                // (the producers add the remainder back, patch_col0 below)
                const int col = 4 * x0 + 2 * px - 2;
                tma_load_4d(&map_img, &patch_full[s], s_patch + s * p.patch_bytes, U8 ? ((col >> 4) << 4) : ((col >> 2) << 2),
                            4 * y0 + 2 * py - 2, 0, b);
                tile = next_tile;
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc_stack = make_idesc(2 * N), idesc_n = make_idesc(N);
        mbar_wait(w_full, 0);
        for (uint32_t lt = 0; sched.next(lt, lane) >= 0; ++lt) {
            mbar_wait(acc_empty, (lt & 1u) ^ 1u);
            mbar_wait(a_full, lt & 1u);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t a = smem_u32(s_a), w = smem_u32(s_w);
                // K chunk 0: 4 steps in 128-byte rows; chunk 1: 1 step in 64-byte rows
                const uint64_t a_hi0 = make_smem_desc(a), a_lo0 = make_smem_desc(a + kABytes), w0 = make_smem_desc(w);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    umma_f16(tmem_base, a_hi0 + 2 * k, w0 + 2 * k, idesc_stack, k > 0 ? 1u : 0u);
                    umma_f16(tmem_base + N, a_lo0 + 2 * k, w0 + 2 * k, idesc_n, 1u);
                }
                const uint64_t a_hi1 = make_smem_desc_sw64(a + 2 * kABytes), a_lo1 = make_smem_desc_sw64(a + 2 * kABytes + kA1Bytes);
                const uint64_t w1 = make_smem_desc_sw64(w + kW0);
                umma_f16(tmem_base, a_hi1, w1, idesc_stack, 1u);
                umma_f16(tmem_base + N, a_lo1, w1, idesc_n, 1u);
                umma_commit(a_empty);
                umma_commit(acc_full);
            }
            __syncwarp();
            // ---- gamma GEMM ----
            mbar_wait(ag_full, lt & 1u);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t ag = smem_u32(s_ag), gm = smem_u32(s_gamma);
                const uint64_t a_hi0 = make_smem_desc(ag), a_lo0 = make_smem_desc(ag + kABytes), g0 = make_smem_desc(gm);
#pragma unroll
                for (int k = 0; k < kGC0 / 16; ++k) {
                    umma_f16(tmem_base + kGammaCol, a_hi0 + 2 * k, g0 + 2 * k, idesc_stack, k > 0 ? 1u : 0u);
                    umma_f16(tmem_base + kGammaCol + N, a_lo0 + 2 * k, g0 + 2 * k, idesc_n, 1u);
                }
                if (kGC1) {
                    const uint64_t a_hi1 = make_smem_desc_sw64(ag + 2 * kABytes), a_lo1 = make_smem_desc_sw64(ag + 2 * kABytes + kA1Bytes);
                    const uint64_t g1 = make_smem_desc_sw64(gm + kW0);
#pragma unroll
                    for (int k = 0; k < kGC1 / 16; ++k) {
                        umma_f16(tmem_base + kGammaCol, a_hi1 + 2 * k, g1 + 2 * k, idesc_stack, 1u);
                        umma_f16(tmem_base + kGammaCol + N, a_lo1 + 2 * k, g1 + 2 * k, idesc_n, 1u);
                    }
                }
                umma_commit(g_full);
            }
            __syncwarp();
        }
    } else if (warp < 2 + kEpiWarps) {
        // =============================== epilogue warps (2..17): 4 column parts x 4 TMEM lane quarters ===============================
        // The epilogue is this kernel's critical path: ~1300 dependent instructions per warp and tile with 8 warps (ncu:
        // profiles/r2g_*); sixteen warps of 8-channel sub-units halve the chain.
        const int part = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int m = quarter * 32 + lane;
        const bool issuer = threadIdx.x == 64;
        const int s_begin = part * kSub / kParts, s_end = (part + 1) * kSub / kParts;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        __half *st_hi = reinterpret_cast<__half *>(s_ag);
        __half *st_lo = reinterpret_cast<__half *>(s_ag + p.stage_plane);
        for (uint32_t lt = 0;; ++lt) {
            const int tile = sched.next(lt, lane);
            if (tile < 0) break;
            ++trace_tiles;
            const int sp = tile % tiles_xy, img = tile / tiles_xy;
            const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
            float x[kSubMax][8];
            // (mbarriers instead of two block-wide bar.sync per tile: only the store-issuing thread waits for the whole tile; the
            // other 15 warps go on to the next tile's accumulators -- barrier stalls were 13 % of this kernel's samples)
            if (issuer) {
                tma_store_wait_read();
                mbar_arrive(stage_free);
            }
            mbar_wait(acc_full, lt & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int ui = 0; ui < kSubMax; ++ui) {
                uint32_t d0[8], d1[8];
                if (s_begin + ui < s_end) {
                    tmem_ld8_nowait(lane_addr + (s_begin + ui) * 8, d0);
                    tmem_ld8_nowait(lane_addr + N + (s_begin + ui) * 8, d1);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 8; ++e) x[ui][e] = fmaf(__uint_as_float(d1[e]), kLoInv, __uint_as_float(d0[e]));
                }
            }
            tcgen05_fence_before();
            mbar_arrive(acc_empty);
            mbar_wait(stage_free, lt & 1u);  // the |x| tile below overwrites the previous tile's staged output
#pragma unroll
            for (int ui = 0; ui < kSubMax; ++ui) {
                const int su = s_begin + ui;
                if (su < s_end) {
                    float a[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) a[e] = fabsf(x[ui][e]);
                    uint4 h, l;
                    split8(a, h, l);
                    if (su < 8) {  // chunk 0 (channels 0..63): 128-byte rows, 16-byte unit j XOR (row & 7)
                        const int phys = (su ^ (m & 7)) << 4;
                        *reinterpret_cast<uint4 *>(s_ag + m * 128 + phys) = h;
                        *reinterpret_cast<uint4 *>(s_ag + kABytes + m * 128 + phys) = l;
                    } else {       // chunk 1: 64-byte rows, unit j XOR ((row >> 1) & 3)
                        const int phys = ((su - 8) ^ ((m >> 1) & 3)) << 4;
                        *reinterpret_cast<uint4 *>(s_ag + 2 * kABytes + m * 64 + phys) = h;
                        *reinterpret_cast<uint4 *>(s_ag + 2 * kABytes + kA1Bytes + m * 64 + phys) = l;
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(ag_full);
            mbar_wait(g_full, lt & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int ui = 0; ui < kSubMax; ++ui) {
                const int c = (s_begin + ui) * 8;
                if (s_begin + ui < s_end) {
                    uint32_t d0[8], d1[8];
                    tmem_ld8_nowait(lane_addr + kGammaCol + c, d0);
                    tmem_ld8_nowait(lane_addr + kGammaCol + N + c, d1);
                    tmem_ld_wait();
                    if (m < rows && c < p.out_c) {
                        float y[8], bt[8];
                        *reinterpret_cast<float4 *>(&bt[0]) = *reinterpret_cast<const float4 *>(s_beta + c);
                        *reinterpret_cast<float4 *>(&bt[4]) = *reinterpret_cast<const float4 *>(s_beta + c + 4);
                        const bool whole = c + 8 <= p.c_out;  // (uniform) every channel of the sub-unit is real
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const float norm = fmaf(__uint_as_float(d1[e]), kLoInv, __uint_as_float(d0[e])) + bt[e];
                            // x * (1 / norm), like the reference; 1 / norm = MUFU.RCP + one Newton step (< 1 ulp)
                            y[e] = (whole || c + e < p.c_out) ? x[ui][e] * fast_rcp(norm) : 0.0f;
                        }
                        uint4 h, l;
                        split8(y, h, l);
                        *reinterpret_cast<uint4 *>(st_hi + m * p.stage_c + c) = h;
                        *reinterpret_cast<uint4 *>(st_lo + m * p.stage_c + c) = l;
                    }
                }
            }
            tcgen05_fence_before();
            fence_proxy_async();
            mbar_arrive(y_done);
            if (issuer) {
                mbar_wait(y_done, lt & 1u);
                tma_store_4d(&map_o_hi, st_hi, 0, x0, y0, img);
                tma_store_4d(&map_o_lo, st_lo, 0, x0, y0, img);
                tma_store_commit();
            }
        }
        if (issuer) tma_store_wait_all();
    } else if (warp < 2 + kEpiWarps + 4) {
        // =============================== im2col producers (4 warps): thread = tile row = output pixel ===============================
        const int m = (warp - 2 - kEpiWarps) * 32 + lane;
        const int ty = m / p.tw, tx = m - ty * p.tw;
        const int row_elems = p.box_w, chan_elems = p.box_w * p.box_h;
        for (uint32_t lt = 0;; ++lt) {
            const int tile = sched.next(lt, lane);
            if (tile < 0) break;
            const uint32_t s = lt & 1u;
            // first image column this tile needs, relative to the (16-byte aligned) box origin: 0 or 2 floats, 0..14 bytes
            const int col = 4 * ((tile % tiles_xy) % p.tiles_x) * p.tw + 2 * ((tile / tiles_xy) & 1) - 2;
            const int patch_col0 = U8 ? (col & 15) : (col & 3);
            mbar_wait(&patch_full[s], (lt >> 1) & 1u);
            mbar_wait(a_empty, (lt & 1u) ^ 1u);  // the previous tile's conv MMAs have read the A tile
            if (m < rows) {
                // K order = (c, dy, dx), like weight.reshape(c_out, -1); 8 values -> one 16-byte unit of the hi and of the lo tile
                const uint8_t *patch = s_patch + s * p.patch_bytes;
                // (uint8 input) which of this pixel's 5 rows / 5 columns lie inside the image
                const int img_t = tile / tiles_xy, sp_t = tile % tiles_xy;
                const int iy0 = 4 * ((sp_t / p.tiles_x) * p.th + ty) + 2 * ((img_t >> 1) & 1) - 2;
                const int ix0 = 4 * ((sp_t % p.tiles_x) * p.tw + tx) + 2 * (img_t & 1) - 2;
                uint32_t col_ok = 0;
#pragma unroll
                for (int dx = 0; dx < 5; ++dx) col_ok |= (static_cast<uint32_t>(ix0 + dx) < static_cast<uint32_t>(p.w_in) ? 1u : 0u) << dx;
                // two phases of 8 row segments = 40 K values = 5 sixteen-byte units each (5 * 8 = 8 * 5): half the live registers
#pragma unroll
                for (int phase = 0; phase < 2; ++phase) {
                    float f[40];
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int r = phase * 8 + rr;  // row segment (c, dy) = (r / 5, r % 5); r = 15 is the zero padding K 75..79
                        float *o = &f[rr * 5];
                        if (r >= 15) {
#pragma unroll
                            for (int i = 0; i < 5; ++i) o[i] = 0.0f;
                        } else {
                            const int c = r / 5, dy = r % 5;
                            const int e0 = c * chan_elems + (4 * ty + dy) * row_elems + 4 * tx + (patch_col0 & ~3);
                            if (U8) {  // 5 bytes starting 0 or 2 bytes into an aligned word
                                const uint32_t w0 = *reinterpret_cast<const uint32_t *>(patch + e0);
                                const uint32_t w1 = *reinterpret_cast<const uint32_t *>(patch + e0 + 4);
                                const uint32_t sh = static_cast<uint32_t>(patch_col0 & 3) * 8u;
                                const uint32_t w4 = __funnelshift_r(w0, w1, sh), b4 = (w1 >> sh) & 255u;
                                // the conv pads with ZEROS of the normalised image, but TMA fills bytes outside the image with 0
                                // and lut[0] = -mean / std: mask those positions instead of looking them up
                                const uint32_t ok = static_cast<uint32_t>(iy0 + dy) < static_cast<uint32_t>(p.h_in) ? col_ok : 0u;
                                o[0] = (ok & 1u) ? s_lut[c * 256 + (w4 & 255u)] : 0.0f;
                                o[1] = (ok & 2u) ? s_lut[c * 256 + ((w4 >> 8) & 255u)] : 0.0f;
                                o[2] = (ok & 4u) ? s_lut[c * 256 + ((w4 >> 16) & 255u)] : 0.0f;
                                o[3] = (ok & 8u) ? s_lut[c * 256 + (w4 >> 24)] : 0.0f;
                                o[4] = (ok & 16u) ? s_lut[c * 256 + b4] : 0.0f;
                            } else {   // 5 floats starting 0 or 2 floats into an aligned float4: two conflict-free 16-byte loads
                                const float4 v0 = *reinterpret_cast<const float4 *>(patch + 4 * e0);
                                const float4 v1 = *reinterpret_cast<const float4 *>(patch + 4 * e0 + 16);
                                if (patch_col0 & 2) {
                                    o[0] = v0.z; o[1] = v0.w; o[2] = v1.x; o[3] = v1.y; o[4] = v1.z;
                                } else {
                                    o[0] = v0.x; o[1] = v0.y; o[2] = v0.z; o[3] = v0.w; o[4] = v1.x;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int gg = 0; gg < 5; ++gg) {
                        const int g = phase * 5 + gg;  // 16-byte unit of the K dimension: 0..7 chunk 0, 8..9 chunk 1
                        uint4 h, l;
                        split8(&f[gg * 8], h, l);
                        if (g < 8) {
                            const int phys = (g ^ (m & 7)) << 4;
                            *reinterpret_cast<uint4 *>(s_a + m * 128 + phys) = h;
                            *reinterpret_cast<uint4 *>(s_a + kABytes + m * 128 + phys) = l;
                        } else {  // K 64..79: 64-byte rows
                            const int phys = ((g - 8) ^ ((m >> 1) & 3)) << 4;
                            *reinterpret_cast<uint4 *>(s_a + 2 * kABytes + m * 64 + phys) = h;
                            *reinterpret_cast<uint4 *>(s_a + 2 * kABytes + kA1Bytes + m * 64 + phys) = l;
                        }
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(a_full);
            mbar_arrive(&patch_empty[s]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 64) trace_emit(p.trace, TRACE_CONV_FIRST, trace_t0, trace_tiles);
}

template <int N, bool U8>
static int launch(const CUtensorMap *maps, const Params &p, int smem, cudaStream_t st) {
    static std::atomic<uint64_t> configured{0};
    if (int rc = ensure_dyn_smem(ga_first_gdn_kernel<N, U8>, 227 * 1024, configured)) return rc;
    const int64_t total = static_cast<int64_t>(p.tiles_x) * p.tiles_y * p.batch * 4;
    if (total > 0x7fffffff) return SC2_ERR_UNSUPPORTED;
    const int grid = total < persistent_grid() ? static_cast<int>(total) : persistent_grid();
    ga_first_gdn_kernel<N, U8><<<grid, kThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], maps[6], p);
    SC2_LAUNCH_CHECK("ga_first_gdn_kernel");
    return SC2_OK;
}

// 2-D K-major weight map with a selectable box width / swizzle: [rows, k_total] fp16, box {box_k, rows}
static int make_kmajor_map(CUtensorMap *m, const void *base, int k_total, int rows, int box_k, CUtensorMapSwizzle sw) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SC2_ERR_CUDA;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(k_total), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(k_total) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_k), static_cast<cuuint32_t>(rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SC2_OK : SC2_ERR_INVALID_ARG;
}

}  // namespace gaf
}  // namespace sc2

extern "C" {

int sc2_ga_first_conv_gdn(const void *image, int image_is_u8, const float *lut, int batch, int h_in, int w_in, int c_out,
                          const void *w_stack, const void *gamma_stack, const float *beta, void *out_hi, void *out_lo, int out_c,
                          int32_t *tile_counter, sc2_stream_t stream) {
    using namespace sc2::gaf;
    if (!image || !w_stack || !gamma_stack || !beta || !out_hi || !out_lo || batch < 1) return SC2_ERR_INVALID_ARG;
    if (image_is_u8 && !lut) return SC2_ERR_INVALID_ARG;
    const int n = sc2_ga_halo_n(c_out);
    if (!n) return SC2_ERR_UNSUPPORTED;
    if (out_c % 8 || out_c < c_out || out_c > n) return SC2_ERR_INVALID_ARG;
    // geometry: Conv2d(3 -> c_out, k5, s2, p2); output (and both parity-plane dimensions) must be even, rows 16-byte aligned
    const int h_out = (h_in + 4 - 5) / 2 + 1, w_out = (w_in + 4 - 5) / 2 + 1;
    if (h_out < 2 || w_out < 2 || (h_out & 1) || (w_out & 1)) return SC2_ERR_UNSUPPORTED;
    if (image_is_u8 ? (w_in % 16 != 0) : (w_in % 4 != 0)) return SC2_ERR_UNSUPPORTED;
    Params p;
    p.batch = batch; p.h_in = h_in; p.w_in = w_in;
    p.hp = h_out / 2; p.wp = w_out / 2;
    const int w0 = 2 * n * 128, w1 = 2 * n * 64;
    const int a_bytes = 2 * kABytes + 2 * kA1Bytes;
    int smem = 0;
    // tile: tw <= 16 columns x th rows, th * tw <= 128, rows balanced over the plane; th shrinks until the patch buffers fit
    const int n_col = (p.wp + 15) / 16;
    p.tw = (p.wp + n_col - 1) / n_col;
    p.tiles_x = n_col;
    int th_max = 128 / p.tw;
    if (th_max > p.hp) th_max = p.hp;
    if (th_max > 16) th_max = 16;
    for (;; --th_max) {
        if (th_max < 1) return SC2_ERR_UNSUPPORTED;
        const int n_row = (p.hp + th_max - 1) / th_max;
        p.th = (p.hp + n_row - 1) / n_row;
        p.tiles_y = n_row;
        // patch box: the 4 tw + 1 columns a tile reads, plus the columns the 16-byte alignment of the box origin adds on the left
        p.box_w = image_is_u8 ? (4 * p.tw + 1 + 14 + 4 + 15) / 16 * 16 : 4 * p.tw + 4;
        p.box_h = 4 * p.th + 1;
        p.patch_bytes = (p.box_w * p.box_h * 3 * (image_is_u8 ? 1 : 4) + 64 + 1023) / 1024 * 1024;  // + slack: the last thread's extra word
        p.stage_c = (out_c / 8) % 2 == 0 ? out_c + 8 : out_c;
        p.stage_plane = (p.th * p.tw * p.stage_c * 2 + 127) / 128 * 128;
        const int ag_region = ((a_bytes > 2 * p.stage_plane ? a_bytes : 2 * p.stage_plane) + 1023) / 1024 * 1024;
        p.off_gamma = w0 + w1;
        p.off_a = p.off_gamma + w0 + w1;
        p.off_ag = p.off_a + a_bytes;
        p.off_patch = p.off_ag + ag_region;
        p.off_bar = p.off_patch + 2 * p.patch_bytes;
        p.off_lut = (p.off_bar + 13 * 8 + 24 + n * 4 + kTileSchedBytes + 16 + 15) / 16 * 16;
        smem = p.off_lut + (image_is_u8 ? 3 * 256 * 4 : 0) + 1024;
        if (smem <= 227 * 1024) break;
    }
    p.c_out = c_out; p.out_c = out_c;
    p.beta = beta; p.lut = lut;
    p.tile_counter = tile_counter;
    p.trace = sc2::trace_sink();
    CUtensorMap maps[7];
    {   // image [batch, 3, h_in, w_in] seen as {w, h, c, batch}; box {box_w, box_h, 3, 1}, no swizzle, zero fill outside = padding
        EncodeTiledFn fn = get_encode_fn();
        if (!fn) return SC2_ERR_CUDA;
        const cuuint64_t es = image_is_u8 ? 1 : 4;
        cuuint64_t dims[4] = {static_cast<cuuint64_t>(w_in), static_cast<cuuint64_t>(h_in), 3, static_cast<cuuint64_t>(batch)};
        cuuint64_t strides[3] = {static_cast<cuuint64_t>(w_in) * es, static_cast<cuuint64_t>(w_in) * h_in * es,
                                 static_cast<cuuint64_t>(w_in) * h_in * 3 * es};
        cuuint32_t box[4] = {static_cast<cuuint32_t>(p.box_w), static_cast<cuuint32_t>(p.box_h), 3, 1};
        cuuint32_t estr[4] = {1, 1, 1, 1};
        if (p.box_w > 256 || p.box_h > 256) return SC2_ERR_UNSUPPORTED;
        CUresult r = fn(&maps[0], image_is_u8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void *>(image),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return SC2_ERR_INVALID_ARG;
    }
    // conv weights [2n, 80]: chunk 0 = K 0..63 (128-byte swizzle), chunk 1 = K 64..95 (64-byte swizzle; beyond 80: zero fill)
    int rc = make_kmajor_map(&maps[1], w_stack, kK, 2 * n, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_kmajor_map(&maps[2], w_stack, kK, 2 * n, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    // gamma [2n, n]
    rc = make_kmajor_map(&maps[3], gamma_stack, n, 2 * n, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = make_kmajor_map(&maps[4], gamma_stack, n, 2 * n, 32, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = make_nhwc_map(&maps[5], out_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out_c, p.wp, p.hp, batch * 4, p.stage_c, p.tw, p.th, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_nhwc_map(&maps[6], out_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out_c, p.wp, p.hp, batch * 4, p.stage_c, p.tw, p.th, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    cudaStream_t st = sc2::as_stream(stream);
#define SC2_GAF_DISPATCH(NN)                                                           \
    case NN: return image_is_u8 ? launch<NN, true>(maps, p, smem, st) : launch<NN, false>(maps, p, smem, st);
    switch (n) {
        SC2_GAF_DISPATCH(16)
        SC2_GAF_DISPATCH(32)
        SC2_GAF_DISPATCH(48)
        SC2_GAF_DISPATCH(64)
        SC2_GAF_DISPATCH(80)
        default: return image_is_u8 ? launch<96, true>(maps, p, smem, st) : launch<96, false>(maps, p, smem, st);
    }
#undef SC2_GAF_DISPATCH
}

}  // extern "C"
