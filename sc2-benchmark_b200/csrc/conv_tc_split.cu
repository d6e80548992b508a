// conv_tc_split.cu -- fp32-grade implicit-GEMM convolution on the tcgen05 tensor cores ("split fp16", 3 MMA passes).
//
// The analysis transform g_a (sc2bench/models/layer.py:475-484) feeds round(y - median): a latent error flips symbols,
// so g_a needs fp32-grade arithmetic while fp16/bf16/tf32 single-pass tensor-core math is 1e-3..1e-4.
// Here every operand a is carried as two fp16 numbers, a = hi + lo / 2048 with hi = fp16(a), lo = fp16((a - hi) * 2048)
// (22 mantissa bits; the 2^11 scale keeps lo in fp16's normal range), and a product sum is evaluated as
//     D0 = sum hi_a hi_b,   D1 = sum (hi_a lo_b + lo_a hi_b),   result = D0 + D1 / 2048        (lo_a lo_b ~ 2^-22 dropped)
// with three tcgen05.mma (kind::f16, fp32 accumulation in TMEM) per K step into two accumulators.  A CPU emulation
// gives 1.8e-7 latent error on the Entropic-Student encoder (torch's own fp32: 3.6e-7) and no symbol mismatch.
//
// Same structure as conv_tc.cu: A tiles are shifted 4-D TMA boxes of the NHWC activation planes (hi and lo),
// B tiles are 2-D boxes of the tap-major packed weights (hi and lo), one elected thread issues the MMAs.
// Stride-2 convolutions read PARITY PLANES: the producer of their input writes pixel (2Y+py, 2X+px) to plane py*2+px at
// (Y, X), so tap (dy, dx) of a stride-2 conv is again a unit-stride shifted box of one plane.
// Epilogues (thread = output pixel, straight from TMEM): store split | GDN1 (x / (beta + gamma.|x|), |x| formed in smem
// by the epilogue warps on both halves) | quantise to int32 symbols in coder (NCHW) order.
#include "tc_common.cuh"

namespace sc2 {
namespace tcs {

using namespace sc2::tc;

constexpr int kNumThreads = 192;
constexpr int kMaxTaps = 25;
constexpr float kLoScale = 2048.0f;
constexpr float kLoInv = 1.0f / 2048.0f;
// TMEM: 512 columns.  The hi*hi sum D0 is spread over up to 6 accumulators ("groups" of taps) that are added in fp32 by the
// epilogue: tcgen05 accumulation truncates, so one accumulator over K = 2400 (150 K-steps) loses ~3e-6; chunked
// summation brings it back to FFMA-grade.  The cross terms D1 are 2^-11 smaller and share one accumulator.
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kD1Col = 384;

// MODE_GDN_SPLIT: the GDN of the CompressAI zoo codecs, y = x / sqrt(beta + gamma . x^2) (compressai.layers.GDN [mem]; the
// transform warps square the A tile in shared memory instead of taking |.|, the epilogue multiplies by rsqrt).
enum Mode { MODE_STORE_SPLIT = 0, MODE_GDN1_SPLIT = 1, MODE_QUANT = 2, MODE_GDN_SPLIT = 3 };
enum Act { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

struct Tap {
    int8_t plane, dx, dy, pad_;
};

struct Params {
    int tiles_x, tiles_y, tw, th;
    int n_taps;
    Tap taps[kMaxTaps];
    int k_chunks, k_steps_last;
    int groups;        // D0 accumulator groups (taps are dealt to groups in order)
    int planes;        // input planes per output image (1, or 4 parity planes)
    int images;        // output images (grid-stride tile loop covers tiles_x * tiles_y * images tiles)
    int c_out;         // valid output channels (<= N_TILE)
    int h_out, w_out;
    const float *beta;
    const float *medians;
    __half *out_hi, *out_lo;
    int out_c;         // channel pitch of the output planes
    int stage_c;       // channel pitch of the staging tile: out_c padded to an ODD number of 16-byte units per row
    int32_t *out_sym;
    const __half *x_hi, *x_lo;
    int *tile_counter;  // zeroed by the caller: dynamic tile schedule; nullptr: static
    TraceSink trace;    // diagnostics (common.cuh)
    // ---- sc2_tc_split_conv_ex ----
    int in5d;           // stride-2 taps straight from an NHWC tensor: 5-D map (px * c_in + c, X, py, Y, image), no parity planes
    int c_in;
    int sym_c_total;    // MODE_QUANT: channels of the whole symbol tensor; this launch writes [sym_c_off, sym_c_off + c_out)
    int sym_c_off;
    int act;            // STORE: activation after the bias (ReLU / LeakyReLU of the hyper-analysis h_a)
    float slope;
};

// B_RES: 1x1 convolutions (one tap, <= 2 K chunks) keep the whole weight matrix resident in shared memory for the life of
// the persistent CTA and stream only A tiles through the ring (re-loading B per tile made those layers L2-bound).
template <int N_TILE, int STAGES, bool B_RES>
struct Smem {
    static constexpr int kBBytes = N_TILE * 128;
    static constexpr int kResBytes = B_RES ? 2 * 2 * kBBytes : 0;          // [chunk][hi, lo]
    static constexpr int kStageBytes = 2 * kABytes + (B_RES ? 0 : 2 * kBBytes);
    static constexpr int kRingBytes = STAGES * kStageBytes;
    // output staging: dense [128 rows][N_TILE channels] fp16 for the hi and the lo plane (TMA store source; in the GDN
    // mode the x tile is TMA-loaded into it first and y overwrites x in place)
    // (rows are padded by 16 bytes: with a pitch of 192 bytes the 16-byte accesses of 8 consecutive rows fall on two bank
    // groups -- 60 M conflict cycles per launch in GDN1(96); the padding channels lie outside the tensor, so the bulk stores
    // skip them and the x-tile loads zero-fill them)
    static constexpr int kStagePlane = kTileM * (N_TILE + 8) * 2;
    static constexpr int kStagingBytes = 2 * kStagePlane;
    static constexpr int kStagingOffset = kResBytes + kRingBytes;
    static constexpr int kBarOffset = kStagingOffset + kStagingBytes;
    static constexpr int kSchedOffset = kBarOffset + (3 * STAGES + 6) * 8 + 16;
    static constexpr int kBetaOffset = kSchedOffset + kTileSchedBytes;  // float[N_TILE]: beta, 1.0 beyond c_out
    static constexpr int kTotal = kBetaOffset + N_TILE * 4;
    static_assert(kTotal + 1024 <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ float fast_rcp(float v) {  // v > 0 (a GDN norm): reciprocal to within 1 ulp in 3 instructions
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return fmaf(r, fmaf(-v, r, 1.0f), r);
}

__device__ __forceinline__ float fast_rsqrt(float v) {  // v > 0: MUFU.RSQ + one Newton step
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r * fmaf(-0.5f * v * r, r, 1.5f);
}

__device__ __forceinline__ void split_store8(const float (&f)[8], __half *hi_ptr, __half *lo_ptr) {
    uint4 h, l;
    uint32_t *hw = reinterpret_cast<uint32_t *>(&h), *lw = reinterpret_cast<uint32_t *>(&l);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const __half2 hh = __floats2half2_rn(f[2 * e], f[2 * e + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn((f[2 * e] - back.x) * kLoScale, (f[2 * e + 1] - back.y) * kLoScale);
        hw[e] = *reinterpret_cast<const uint32_t *>(&hh);
        lw[e] = *reinterpret_cast<const uint32_t *>(&ll);
    }
    *reinterpret_cast<uint4 *>(hi_ptr) = h;
    *reinterpret_cast<uint4 *>(lo_ptr) = l;
}

template <int N_TILE, int STAGES, int MODE, bool B_RES>
__global__ void __launch_bounds__((MODE == MODE_GDN1_SPLIT || MODE == MODE_GDN_SPLIT) ? 448 : 320, 1)
tc_split_conv_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                     const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                     const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                     const __grid_constant__ Params p) {
    // Persistent: the CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  With one D0 group the two 256-column
    // halves of TMEM alternate between tiles (epilogue of tile i overlaps the MMAs of tile i + 1); with several D0
    // groups (long K) the whole TMEM belongs to one tile at a time.
    using L = Smem<N_TILE, STAGES, B_RES>;
    constexpr bool kGdn = MODE == MODE_GDN1_SPLIT || MODE == MODE_GDN_SPLIT;
    constexpr uint32_t kSlot = N_TILE <= 64 ? 64 : 128;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem_res = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *smem = smem_res + L::kResBytes;  // the ring
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_res + L::kBarOffset);
    uint64_t *empty = full + STAGES;
    uint64_t *xform = empty + STAGES;
    uint64_t *acc_full = xform + STAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *b_full = acc_empty + 2;
    uint64_t *x_full = b_full + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(x_full + 1);
    uint8_t *staging = smem_res + L::kStagingOffset;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rows = p.tw * p.th;
    const int k_iters = p.n_taps * p.k_chunks;
    const int tiles_xy = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_xy * p.images;
    const uint32_t acc_stages = p.groups == 1 ? 2u : 1u;
    const unsigned long long trace_t0 = p.trace.buf ? trace_now() : 0ull;
    int trace_tiles = 0;
    TileSched sched;
    sched.bind(smem_res + L::kSchedOffset, p.tile_counter, total_tiles);

    float *s_beta = reinterpret_cast<float *>(smem_res + L::kBetaOffset);
    // GDN: beta (1.0 beyond c_out keeps the reciprocal finite); otherwise the convolution's bias (0 when there is none)
    if (threadIdx.x < N_TILE)
        s_beta[threadIdx.x] = (p.beta && static_cast<int>(threadIdx.x) < p.c_out) ? __ldg(p.beta + threadIdx.x) : (kGdn ? 1.0f : 0.0f);

    if (threadIdx.x == 0) {
        sched.init(kGdn ? 13 : 9);  // consumers: MMA warp, 8 epilogue warps (, 4 transform warps)
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&xform[s], 128);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 256);
        }
        mbar_init(b_full, 1);
        mbar_init(x_full, 1);
        if (MODE != MODE_QUANT) {
            tma_prefetch_desc(&map_o_hi);
            tma_prefetch_desc(&map_o_lo);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================== TMA producer ===============================
        if (elect_one()) {
            const uint32_t stage_tx = static_cast<uint32_t>(2 * rows * 128 + (B_RES ? 0 : 2 * L::kBBytes));
            if (B_RES) {  // the whole (single-tap) weight matrix, once per CTA
                mbar_expect_tx(b_full, static_cast<uint32_t>(p.k_chunks * 2 * L::kBBytes));
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    tma_load_2d(&map_b_hi, b_full, smem_res + (2 * kc) * L::kBBytes, kc * kBlockK, 0);
                    tma_load_2d(&map_b_lo, b_full, smem_res + (2 * kc + 1) * L::kBBytes, kc * kBlockK, 0);
                }
            }
            uint32_t it = 0;
            int tile = sched.claim(0);
            for (uint32_t qn = 0;; ++qn) {
                sched.publish(qn, tile);
                if (tile < 0) break;
                const int next_tile = sched.claim(qn + 1);  // claimed early: the atomic's latency hides behind this tile's loads
                const int sp = tile % tiles_xy, img = tile / tiles_xy;
                const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
                for (int t = 0; t < p.n_taps; ++t) {
                    const Tap tap = p.taps[t];
                    for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
                        const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                        mbar_wait(&empty[s], ph ^ 1u);
                        uint8_t *dst = smem + s * L::kStageBytes;
                        mbar_expect_tx(&full[s], stage_tx);
                        const int cx = x0 + tap.dx, cy = y0 + tap.dy, cz = img * p.planes + tap.plane;
                        if (p.in5d) {
                            const int cc = (tap.plane & 1) * p.c_in + kc * kBlockK;
                            tma_load_5d(&map_a_hi, &full[s], dst, cc, cx, tap.plane >> 1, cy, img);
                            tma_load_5d(&map_a_lo, &full[s], dst + kABytes, cc, cx, tap.plane >> 1, cy, img);
                        } else {
                            tma_load_4d(&map_a_hi, &full[s], dst, kc * kBlockK, cx, cy, cz);
                            tma_load_4d(&map_a_lo, &full[s], dst + kABytes, kc * kBlockK, cx, cy, cz);
                        }
                        if (!B_RES) {
                            tma_load_2d(&map_b_hi, &full[s], dst + 2 * kABytes, kc * kBlockK, t * N_TILE);
                            tma_load_2d(&map_b_lo, &full[s], dst + 2 * kABytes + L::kBBytes, kc * kBlockK, t * N_TILE);
                        }
                    }
                }
                tile = next_tile;
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc = make_idesc(N_TILE);
        uint32_t it = 0;
        if (B_RES) mbar_wait(b_full, 0);
        for (uint32_t lt = 0; sched.next(lt, lane) >= 0; ++lt) {
            const uint32_t as = lt % acc_stages, aph = (lt / acc_stages) & 1u;
            mbar_wait(&acc_empty[as], aph ^ 1u);
            tcgen05_fence_after();
            const uint32_t d1_col = p.groups == 1 ? as * 256u + 128u : kD1Col;
            uint32_t started = 0;  // bit g: D0 group g holds data; bit 31: D1 does (tracked by every lane)
            for (int t = 0; t < p.n_taps; ++t)
                for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
                    const uint32_t g = static_cast<uint32_t>(t * p.groups / p.n_taps);
                    const uint32_t d0_col = p.groups == 1 ? as * 256u : g * kSlot;
                    const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                    mbar_wait(kGdn ? &xform[s] : &full[s], ph);
                    tcgen05_fence_after();
                    if (elect_one()) {
                        const uint32_t base = smem_u32(smem + s * L::kStageBytes);
                        const uint64_t a_hi = make_smem_desc(base), a_lo = make_smem_desc(base + kABytes);
                        const uint32_t b_base = B_RES ? smem_u32(smem_res + (2 * kc) * L::kBBytes) : base + 2 * kABytes;
                        const uint64_t b_hi = make_smem_desc(b_base), b_lo = make_smem_desc(b_base + L::kBBytes);
                        const int k_steps = (kc == p.k_chunks - 1) ? p.k_steps_last : kBlockK / 16;
                        for (int k = 0; k < k_steps; ++k) {
                            const uint32_t acc0 = k > 0 ? 1u : ((started >> g) & 1u), acc1 = k > 0 ? 1u : (started >> 31);
                            umma_f16(tmem_base + d0_col, a_hi + 2 * k, b_hi + 2 * k, idesc, acc0);  // D0[g] += hi * hi
                            umma_f16(tmem_base + d1_col, a_hi + 2 * k, b_lo + 2 * k, idesc, acc1);  // D1 += hi * lo
                            umma_f16(tmem_base + d1_col, a_lo + 2 * k, b_hi + 2 * k, idesc, 1u);    // D1 += lo * hi
                        }
                        umma_commit(&empty[s]);
                        if (t == p.n_taps - 1 && kc == p.k_chunks - 1) umma_commit(&acc_full[as]);
                    }
                    started |= (1u << g) | 0x80000000u;
                    __syncwarp();
                }
        }
    } else if (warp < 10) {
        // =============================== epilogue warps (2..9): two warpgroups, alternate 32-column chunks ===============================
        const int half = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const int ty = row / p.tw, tx = row - ty * p.tw;
        const bool issuer = threadIdx.x == 64;  // first epilogue thread: owns the bulk-store group and the x-tile loads
        __half *st_hi = reinterpret_cast<__half *>(staging), *st_lo = reinterpret_cast<__half *>(staging + L::kStagePlane);
        for (uint32_t lt = 0;; ++lt) {
            const int tile = sched.next(lt, lane);
            if (tile < 0) break;
            ++trace_tiles;
            const uint32_t as = lt % acc_stages, aph = (lt / acc_stages) & 1u;
            const int sp = tile % tiles_xy, img = tile / tiles_xy;
            const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
            const int oy = y0 + ty, ox = x0 + tx;
            const bool valid = row < rows && oy < p.h_out && ox < p.w_out;
            if (MODE != MODE_QUANT) {
                // the staging buffer is free once the previous tile's bulk stores have read it
                if (issuer) {
                    tma_store_wait_read();
                    if (kGdn) {  // x tile (hi, lo) -> staging; y will overwrite it in place
                        mbar_expect_tx(x_full, static_cast<uint32_t>(2 * rows * p.stage_c * 2));
                        tma_load_4d(&map_x_hi, x_full, st_hi, 0, x0, y0, img);
                        tma_load_4d(&map_x_lo, x_full, st_lo, 0, x0, y0, img);
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (kGdn) mbar_wait(x_full, lt & 1u);
            }
            mbar_wait(&acc_full[as], aph);
            tcgen05_fence_after();
            const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
            const uint32_t d0_base = p.groups == 1 ? as * 256u : 0u;
            const uint32_t d1_col = p.groups == 1 ? as * 256u + 128u : kD1Col;
#pragma unroll 1
            for (int c0 = half * 32; c0 < N_TILE; c0 += 64) {
                uint32_t d0[32], d1[32];
                tmem_ld32(lane_addr + d0_base + c0, d0);
                for (int g = 1; g < p.groups; ++g) {  // chunked summation of the hi*hi partial sums, fp32 round-to-nearest
                    uint32_t dg[32];
                    tmem_ld32(lane_addr + g * kSlot + c0, dg);
#pragma unroll
                    for (int e = 0; e < 32; ++e) d0[e] = __float_as_uint(__uint_as_float(d0[e]) + __uint_as_float(dg[e]));
                }
                tmem_ld32(lane_addr + d1_col + c0, d1);
                if (MODE == MODE_QUANT) {
                    if (!valid) continue;
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        const int c = c0 + e;
                        if (c < p.c_out) {
                            const float v = __uint_as_float(d0[e]) + __uint_as_float(d1[e]) * kLoInv + s_beta[c];
                            const float med = p.medians ? __ldg(p.medians + c) : 0.0f;
                            p.out_sym[((static_cast<int64_t>(img) * p.sym_c_total + p.sym_c_off + c) * p.h_out + oy) * p.w_out + ox] =
                                __float2int_rn(rintf(v - med));
                        }
                    }
                } else if (row < rows) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int c = c0 + 8 * g;
                        if (c >= p.out_c) continue;
                        float f[8];
                        const bool whole = c + 8 <= p.c_out;  // (uniform) every channel of the group is real
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            float v = fmaf(__uint_as_float(d1[8 * g + e]), kLoInv, __uint_as_float(d0[8 * g + e]));
                            if (!kGdn) {
                                v += s_beta[c + e];
                                if (p.act == ACT_RELU) v = fmaxf(v, 0.0f);
                                else if (p.act == ACT_LEAKY) v = v > 0.0f ? v : v * p.slope;
                            }
                            f[e] = (whole || c + e < p.c_out) ? v : 0.0f;
                        }
                        __half *ph = st_hi + row * p.stage_c + c, *pl = st_lo + row * p.stage_c + c;
                        if (kGdn) {
                            const uint4 xh = *reinterpret_cast<const uint4 *>(ph);
                            const uint4 xl = *reinterpret_cast<const uint4 *>(pl);
                            const __half2 *xhh = reinterpret_cast<const __half2 *>(&xh), *xlh = reinterpret_cast<const __half2 *>(&xl);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 a = __half22float2(xhh[e]), b = __half22float2(xlh[e]);
                                const float x0f = fmaf(b.x, kLoInv, a.x), x1f = fmaf(b.y, kLoInv, a.y);
                                if (MODE == MODE_GDN_SPLIT) {
                                    f[2 * e] = x0f * fast_rsqrt(f[2 * e] + s_beta[c + 2 * e]);
                                    f[2 * e + 1] = x1f * fast_rsqrt(f[2 * e + 1] + s_beta[c + 2 * e + 1]);
                                } else {
                                    // x * (1 / norm), like the reference; 1 / norm = MUFU.RCP + one Newton step (< 1 ulp)
                                    f[2 * e] = x0f * fast_rcp(f[2 * e] + s_beta[c + 2 * e]);
                                    f[2 * e + 1] = x1f * fast_rcp(f[2 * e + 1] + s_beta[c + 2 * e + 1]);
                                }
                            }
                        }
                        split_store8(f, ph, pl);
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(&acc_empty[as]);
            if (MODE != MODE_QUANT) {
                fence_proxy_async();
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (issuer) {
                    // (a transposed convolution's parity sub-grid is the same store through a strided view of the output)
                    tma_store_4d(&map_o_hi, st_hi, 0, x0, y0, img);
                    tma_store_4d(&map_o_lo, st_lo, 0, x0, y0, img);
                    tma_store_commit();
                }
            }
        }
        if (MODE != MODE_QUANT && issuer) tma_store_wait_all();
    } else if (kGdn) {
        // =============================== |x| transform warps (10..13) ===============================
        // |a| = |hi| + sign(hi) * lo / 2048: clear hi's sign bits, flip lo's where hi was negative
        const int row = (warp - 10) * 32 + lane;
        uint32_t it = 0;
        for (uint32_t lt = 0; sched.next(lt, lane) >= 0; ++lt)
            for (int k_it = 0; k_it < k_iters; ++k_it, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                if (row < rows) {
                    uint4 *rh = reinterpret_cast<uint4 *>(smem + s * L::kStageBytes + row * 128);
                    uint4 *rl = reinterpret_cast<uint4 *>(smem + s * L::kStageBytes + kABytes + row * 128);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        // rotate the 16-byte chunk with the row: 8 neighbouring rows hit 8 different bank groups
                        const int cc = (c + row) & 7;
                        uint4 h = rh[cc], l = rl[cc];
                        if (MODE == MODE_GDN_SPLIT) {  // x^2 in fp32, split again
                            uint32_t *hw = reinterpret_cast<uint32_t *>(&h), *lw = reinterpret_cast<uint32_t *>(&l);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 a = __half22float2(*reinterpret_cast<const __half2 *>(&hw[e]));
                                const float2 b = __half22float2(*reinterpret_cast<const __half2 *>(&lw[e]));
                                const float x0f = fmaf(b.x, kLoInv, a.x), x1f = fmaf(b.y, kLoInv, a.y);
                                const float s0 = x0f * x0f, s1 = x1f * x1f;
                                const __half2 hh = __floats2half2_rn(s0, s1);
                                const float2 back = __half22float2(hh);
                                const __half2 ll = __floats2half2_rn((s0 - back.x) * kLoScale, (s1 - back.y) * kLoScale);
                                hw[e] = *reinterpret_cast<const uint32_t *>(&hh);
                                lw[e] = *reinterpret_cast<const uint32_t *>(&ll);
                            }
                        } else {
                            l.x ^= h.x & 0x80008000u; l.y ^= h.y & 0x80008000u; l.z ^= h.z & 0x80008000u; l.w ^= h.w & 0x80008000u;
                            h.x &= 0x7fff7fffu; h.y &= 0x7fff7fffu; h.z &= 0x7fff7fffu; h.w &= 0x7fff7fffu;
                        }
                        rh[cc] = h;
                        rl[cc] = l;
                    }
                }
                fence_proxy_async();
                mbar_arrive(&xform[s]);
            }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (threadIdx.x == 64) trace_emit(p.trace, TRACE_CONV_SPLIT, trace_t0, trace_tiles);
}

// ---- im2col for the first layer (c_in = 3): fp32 NCHW image -> split fp16 patches in parity-plane pixel order -------
// out planes [batch * 4, H/2... ] are indexed (b, py, px, Y, X) with output pixel (oy, ox) = (2Y + py, 2X + px);
// K index = (c * kh + dy) * kw + dx, zero padded to k_pad.
__global__ void patchify_split_kernel(const float *__restrict__ x, __half *__restrict__ out_hi, __half *__restrict__ out_lo,
                                      int c_in, int h_in, int w_in, int kh, int kw, int stride, int pad, int hp, int wp,
                                      int k_pad, int64_t total_groups, int plain) {
    const int groups = k_pad / 8;
    const int K = c_in * kh * kw;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total_groups;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % groups);
        int64_t pixel = i / groups;  // ((b * 4 + parity) * hp + Y) * wp + X;  plain: (b * hp + oy) * wp + ox
        const int X = static_cast<int>(pixel % wp);
        pixel /= wp;
        const int Y = static_cast<int>(pixel % hp);
        pixel /= hp;
        const int parity = plain ? 0 : static_cast<int>(pixel & 3);
        const int64_t b = plain ? pixel : pixel >> 2;
        const int oy = plain ? Y : 2 * Y + (parity >> 1), ox = plain ? X : 2 * X + (parity & 1);
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = g * 8 + e;
            float v = 0.0f;
            if (k < K) {
                const int c = k / (kh * kw);
                const int r = k - c * kh * kw;
                const int dy = r / kw, dx = r - dy * kw;
                const int iy = oy * stride - pad + dy, ix = ox * stride - pad + dx;
                if (iy >= 0 && iy < h_in && ix >= 0 && ix < w_in) v = __ldg(x + ((b * c_in + c) * h_in + iy) * w_in + ix);
            }
            f[e] = v;
        }
        split_store8(f, out_hi + i * 8, out_lo + i * 8);
    }
}

template <int N_TILE, int STAGES, int MODE, bool B_RES>
static int launch(const CUtensorMap *maps, const Params &p, int images, cudaStream_t st) {
    const CUtensorMap &mah = maps[0], &mal = maps[1], &mbh = maps[2], &mbl = maps[3];
    using L = Smem<N_TILE, STAGES, B_RES>;
    const int smem = uniform_smem(L::kTotal + 1024);  // + slack for the manual 1024-byte alignment
    static std::atomic<uint64_t> configured{0};  // per device ordinal
    if (int rc = ensure_dyn_smem(tc_split_conv_kernel<N_TILE, STAGES, MODE, B_RES>, smem, configured)) return rc;
    const int64_t total = static_cast<int64_t>(p.tiles_x) * p.tiles_y * images;
    if (total > 0x7fffffff) return SC2_ERR_UNSUPPORTED;
    const int grid = total < persistent_grid() ? static_cast<int>(total) : persistent_grid();
    tc_split_conv_kernel<N_TILE, STAGES, MODE, B_RES><<<grid, (MODE == MODE_GDN1_SPLIT || MODE == MODE_GDN_SPLIT) ? 448 : 320, smem, st>>>(mah, mal, mbh, mbl, maps[4], maps[5], maps[6], maps[7], p);
    SC2_LAUNCH_CHECK("tc_split_conv_kernel");
    return SC2_OK;
}

template <int N_TILE, int STAGES, int STAGES_RES>
static int dispatch_mode(int mode, const CUtensorMap *maps, const Params &p, int images, cudaStream_t st) {
    const bool res = p.n_taps == 1 && p.k_chunks <= 2;
    switch (mode) {
        case MODE_STORE_SPLIT:
            return res ? launch<N_TILE, STAGES_RES, MODE_STORE_SPLIT, true>(maps, p, images, st)
                       : launch<N_TILE, STAGES, MODE_STORE_SPLIT, false>(maps, p, images, st);
        case MODE_GDN1_SPLIT:
            return res ? launch<N_TILE, STAGES_RES, MODE_GDN1_SPLIT, true>(maps, p, images, st)
                       : launch<N_TILE, STAGES, MODE_GDN1_SPLIT, false>(maps, p, images, st);
        case MODE_GDN_SPLIT:
            return res ? launch<N_TILE, STAGES_RES, MODE_GDN_SPLIT, true>(maps, p, images, st)
                       : launch<N_TILE, STAGES, MODE_GDN_SPLIT, false>(maps, p, images, st);
        default:
            return res ? launch<N_TILE, STAGES_RES, MODE_QUANT, true>(maps, p, images, st)
                       : launch<N_TILE, STAGES, MODE_QUANT, false>(maps, p, images, st);
    }
}

}  // namespace tcs
}  // namespace sc2

extern "C" {

int sc2_tc_split_n_tile(int c_out) {
    const int tiles[] = {32, 48, 64, 96, 128};
    for (int t : tiles)
        if (c_out <= t) return t;
    return 0;
}

int sc2_tc_split_conv_ex(const sc2_tc_split_ex_desc *d, const void *x_hi, const void *x_lo, const void *w_hi, const void *w_lo,
                         const float *vec, const float *medians, const void *gdn_x_hi, const void *gdn_x_lo, void *out_hi,
                         void *out_lo, int32_t *out_sym, int32_t *tile_counter, sc2_stream_t stream) {
    using namespace sc2::tcs;
    if (!d || !x_hi || !x_lo || !w_hi || !w_lo) return SC2_ERR_INVALID_ARG;
    if (d->images < 1 || d->c_in < 16 || d->c_in % 16 || d->c_out < 1 || d->n_off < 0 || d->n_off % 8) return SC2_ERR_INVALID_ARG;
    if (d->stride != 1 && d->stride != 2) return SC2_ERR_UNSUPPORTED;
    if (d->kh * d->kw > kMaxTaps || d->kh < 1 || d->kw < 1) return SC2_ERR_UNSUPPORTED;
    const int n_tile = sc2_tc_split_n_tile(d->c_out);
    if (!n_tile) return SC2_ERR_UNSUPPORTED;
    const bool gdn = d->mode == MODE_GDN1_SPLIT || d->mode == MODE_GDN_SPLIT;
    if (d->mode < 0 || d->mode > MODE_GDN_SPLIT) return SC2_ERR_INVALID_ARG;
    if (gdn && (!vec || !gdn_x_hi || !gdn_x_lo || d->kh != 1 || d->kw != 1 || d->stride != 1)) return SC2_ERR_INVALID_ARG;
    if (d->mode == MODE_QUANT ? !out_sym : (!out_hi || !out_lo)) return SC2_ERR_INVALID_ARG;
    if (d->mode == MODE_QUANT && (d->c_total < d->n_off + d->c_out)) return SC2_ERR_INVALID_ARG;
    if (d->in_nhwc && (d->stride != 2 || (d->h_in & 1) || (d->w_in & 1))) return SC2_ERR_UNSUPPORTED;
    if (d->act < ACT_NONE || d->act > ACT_LEAKY) return SC2_ERR_INVALID_ARG;
    const bool interleave = d->out_stride == 2;
    if (d->out_stride != 1 && d->out_stride != 2) return SC2_ERR_INVALID_ARG;
    if (interleave && (d->mode != MODE_STORE_SPLIT || d->stride != 1 || (d->out_py | d->out_px) & ~1)) return SC2_ERR_INVALID_ARG;
    // full-resolution output of an interleaved store: out_h x out_w (0: 2 h_out x 2 w_out); this launch owns the pixels of its parity
    const int full_h = d->out_h > 0 ? d->out_h : 2 * d->h_out, full_w = d->out_w > 0 ? d->out_w : 2 * d->w_out;
    if (interleave && ((full_h - d->out_py + 1) / 2 != d->h_out || (full_w - d->out_px + 1) / 2 != d->w_out)) return SC2_ERR_INVALID_ARG;
    const int pad_x = d->pad_x < 0 ? d->pad : d->pad_x;
    // channels of the output planes this launch owns: [n_off, n_off + out_ext); a launch that does not reach the end of the
    // pixel (an inner N tile) must fill its tile completely
    int out_ext = 0;
    if (d->mode != MODE_QUANT) {
        if (d->out_pitch % 8 || d->out_pitch < d->n_off + d->c_out) return SC2_ERR_INVALID_ARG;
        out_ext = d->out_pitch - d->n_off < n_tile ? d->out_pitch - d->n_off : n_tile;
        if (d->n_off + out_ext < d->out_pitch && d->c_out != n_tile) return SC2_ERR_INVALID_ARG;
    }
    Params p;
    const int planes = (d->stride == 2 && !d->in_nhwc) ? 4 : 1;
    p.planes = planes;
    p.images = d->images;
    p.h_out = d->h_out; p.w_out = d->w_out;
    int n_col_tiles = (d->w_out + 127) / 128;
    int tw = (d->w_out + n_col_tiles - 1) / n_col_tiles;
    tw = (tw + 7) / 8 * 8;
    if (tw > 128) tw = 128;
    int th = 128 / tw;
    if (th > d->h_out) th = d->h_out;
    p.tw = tw; p.th = th;
    p.tiles_x = (d->w_out + tw - 1) / tw;
    p.tiles_y = (d->h_out + th - 1) / th;
    p.n_taps = d->kh * d->kw;
    for (int dy = 0; dy < d->kh; ++dy)
        for (int dx = 0; dx < d->kw; ++dx) {
            Tap t;
            const int sy = dy - d->pad, sx = dx - pad_x;
            if (d->stride == 2) {
                const int py = ((sy % 2) + 2) % 2, px = ((sx % 2) + 2) % 2;
                t.plane = static_cast<int8_t>(py * 2 + px);
                t.dy = static_cast<int8_t>((sy - py) / 2);
                t.dx = static_cast<int8_t>((sx - px) / 2);
            } else {
                t.plane = 0;
                t.dy = static_cast<int8_t>(sy);
                t.dx = static_cast<int8_t>(sx);
            }
            t.pad_ = 0;
            p.taps[dy * d->kw + dx] = t;
        }
    p.k_chunks = (d->c_in + kBlockK - 1) / kBlockK;
    const int rem = d->c_in - (p.k_chunks - 1) * kBlockK;
    p.k_steps_last = rem / 16;
    {   // one D0 accumulator per ~24 K-steps, as many as TMEM holds (6 for N <= 64, 3 for N <= 128)
        const int steps = p.n_taps * ((p.k_chunks - 1) * 4 + p.k_steps_last);
        int groups = (steps + 23) / 24;
        const int max_groups = n_tile <= 64 ? 6 : 3;
        if (groups > max_groups) groups = max_groups;
        if (groups > p.n_taps) groups = p.n_taps;
        if (groups < 1) groups = 1;
        p.groups = groups;
    }
    p.c_out = d->c_out;
    p.beta = vec ? vec + d->n_off : nullptr;
    p.medians = medians ? medians + d->n_off : nullptr;
    __half *o_hi = static_cast<__half *>(out_hi), *o_lo = static_cast<__half *>(out_lo);
    if (o_hi) {
        const int64_t shift = d->n_off + (interleave ? (static_cast<int64_t>(d->out_py) * full_w + d->out_px) * d->out_pitch : 0);
        o_hi += shift;
        o_lo += shift;
    }
    p.out_hi = o_hi; p.out_lo = o_lo;
    p.out_c = out_ext;
    p.stage_c = (out_ext / 8) % 2 == 0 ? out_ext + 8 : out_ext;  // odd number of 16-byte units per staging row
    p.out_sym = out_sym;
    const __half *gx_hi = static_cast<const __half *>(gdn_x_hi), *gx_lo = static_cast<const __half *>(gdn_x_lo);
    if (gx_hi) { gx_hi += d->n_off; gx_lo += d->n_off; }
    p.x_hi = gx_hi; p.x_lo = gx_lo;
    p.tile_counter = tile_counter;
    p.trace = sc2::trace_sink();
    p.in5d = d->in_nhwc ? 1 : 0;
    p.c_in = d->c_in;
    p.sym_c_total = d->mode == MODE_QUANT ? d->c_total : 0;
    p.sym_c_off = d->n_off;
    p.act = d->act;
    p.slope = d->slope;
    CUtensorMap maps[8];
    CUtensorMap &mah = maps[0], &mal = maps[1], &mbh = maps[2], &mbl = maps[3];
    int rc;
    if (d->in_nhwc) {
        // input: an NHWC tensor [images, h_in, w_in, c_in] (full resolution) read through the stride-2 view
        rc = make_nhwc_s2_map(&mah, x_hi, d->c_in, d->w_in, d->h_in, d->images, kBlockK, tw, th);
        if (rc) return rc;
        rc = make_nhwc_s2_map(&mal, x_lo, d->c_in, d->w_in, d->h_in, d->images, kBlockK, tw, th);
        if (rc) return rc;
    } else {
        // input planes: [images * planes, h_in, w_in, c_in] (h_in, w_in = plane geometry)
        rc = make_nhwc_map(&mah, x_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->c_in, d->w_in, d->h_in, d->images * planes, kBlockK, tw, th);
        if (rc) return rc;
        rc = make_nhwc_map(&mal, x_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->c_in, d->w_in, d->h_in, d->images * planes, kBlockK, tw, th);
        if (rc) return rc;
    }
    // packed weights: [taps * n_tile, c_in] (rows beyond c_out are zero)
    rc = make_weight_map(&mbh, w_hi, d->c_in, p.n_taps * n_tile, n_tile);
    if (rc) return rc;
    rc = make_weight_map(&mbl, w_lo, d->c_in, p.n_taps * n_tile, n_tile);
    if (rc) return rc;
    maps[4] = maps[5] = maps[6] = maps[7] = mah;
    if (d->mode != MODE_QUANT) {
        // dense (unswizzled) boxes {stage_c, tw, th, 1}: bulk-store sources / x-tile destinations in the staging buffer; the
        // box is wider than the tensor view when the row pitch is padded (channels beyond the view: skipped / zero-filled)
        const CUtensorMapDataType f16 = CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
        if (interleave) {
            rc = make_nhwc_parity_out_map(&maps[4], o_hi, out_ext, d->out_pitch, full_w, full_h, d->out_py, d->out_px, d->images, p.stage_c, tw, th);
            if (rc) return rc;
            rc = make_nhwc_parity_out_map(&maps[5], o_lo, out_ext, d->out_pitch, full_w, full_h, d->out_py, d->out_px, d->images, p.stage_c, tw, th);
        } else {
            rc = make_nhwc_map_pitch(&maps[4], o_hi, f16, 2, out_ext, d->out_pitch, d->w_out, d->h_out, d->images, p.stage_c, tw, th,
                                     CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
            rc = make_nhwc_map_pitch(&maps[5], o_lo, f16, 2, out_ext, d->out_pitch, d->w_out, d->h_out, d->images, p.stage_c, tw, th,
                                     CU_TENSOR_MAP_SWIZZLE_NONE);
        }
        if (rc) return rc;
        if (gdn) {
            rc = make_nhwc_map_pitch(&maps[6], gx_hi, f16, 2, out_ext, d->out_pitch, d->w_out, d->h_out, d->images, p.stage_c, tw, th,
                                     CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
            rc = make_nhwc_map_pitch(&maps[7], gx_lo, f16, 2, out_ext, d->out_pitch, d->w_out, d->h_out, d->images, p.stage_c, tw, th,
                                     CU_TENSOR_MAP_SWIZZLE_NONE);
            if (rc) return rc;
        }
    }
    cudaStream_t st = sc2::as_stream(stream);
    switch (n_tile) {
        case 32: return dispatch_mode<32, 4, 5>(d->mode, maps, p, d->images, st);
        case 48: return dispatch_mode<48, 4, 4>(d->mode, maps, p, d->images, st);
        case 64: return dispatch_mode<64, 3, 4>(d->mode, maps, p, d->images, st);
        case 96: return dispatch_mode<96, 2, 3>(d->mode, maps, p, d->images, st);
        default: return dispatch_mode<128, 2, 2>(d->mode, maps, p, d->images, st);
    }
}

int sc2_tc_split_conv(const sc2_tc_split_desc *d, const void *x_hi, const void *x_lo, const void *w_hi, const void *w_lo,
                      const float *beta, const float *medians, const void *gdn_x_hi, const void *gdn_x_lo, void *out_hi,
                      void *out_lo, int32_t *out_sym, int32_t *tile_counter, sc2_stream_t stream) {
    if (!d) return SC2_ERR_INVALID_ARG;
    if (d->mode < 0 || d->mode > 2) return SC2_ERR_INVALID_ARG;
    if (d->mode != 2 && d->out_c > sc2_tc_split_n_tile(d->c_out)) return SC2_ERR_INVALID_ARG;
    sc2_tc_split_ex_desc e;
    e.images = d->images; e.h_in = d->h_in; e.w_in = d->w_in; e.c_in = d->c_in;
    e.c_out = d->c_out; e.kh = d->kh; e.kw = d->kw; e.stride = d->stride; e.pad = d->pad;
    e.mode = d->mode;
    e.h_out = d->h_out; e.w_out = d->w_out;
    e.out_pitch = d->out_c; e.n_off = 0; e.c_total = d->c_out; e.in_nhwc = 0; e.act = 0; e.slope = 0.0f;
    e.pad_x = -1; e.out_stride = 1; e.out_py = 0; e.out_px = 0; e.out_h = 0; e.out_w = 0;
    return sc2_tc_split_conv_ex(&e, x_hi, x_lo, w_hi, w_lo, beta, medians, gdn_x_hi, gdn_x_lo, out_hi, out_lo, out_sym, tile_counter,
                                stream);
}

int sc2_patchify_split(const float *x, void *out_hi, void *out_lo, int batch, int c_in, int h_in, int w_in, int kh, int kw,
                       int stride, int pad, int k_pad, sc2_stream_t stream) {
    if (!x || !out_hi || !out_lo || batch < 1 || k_pad % 16 || k_pad < c_in * kh * kw) return SC2_ERR_INVALID_ARG;
    const int h_out = (h_in + 2 * pad - kh) / stride + 1, w_out = (w_in + 2 * pad - kw) / stride + 1;
    if (h_out < 2 || w_out < 2 || (h_out & 1) || (w_out & 1)) return SC2_ERR_UNSUPPORTED;
    const int hp = h_out / 2, wp = w_out / 2;
    const int64_t total = static_cast<int64_t>(batch) * 4 * hp * wp * (k_pad / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > sc2::kNumSMs * 32) blocks = sc2::kNumSMs * 32;
    sc2::tcs::patchify_split_kernel<<<static_cast<int>(blocks), 256, 0, sc2::as_stream(stream)>>>(
        x, static_cast<__half *>(out_hi), static_cast<__half *>(out_lo), c_in, h_in, w_in, kh, kw, stride, pad, hp, wp, k_pad, total, 0);
    SC2_LAUNCH_CHECK("patchify_split_kernel");
    return SC2_OK;
}

int sc2_patchify_split_nhwc(const float *x, void *out_hi, void *out_lo, int batch, int c_in, int h_in, int w_in, int kh, int kw,
                            int stride, int pad, int k_pad, sc2_stream_t stream) {
    if (!x || !out_hi || !out_lo || batch < 1 || k_pad % 16 || k_pad < c_in * kh * kw || stride < 1) return SC2_ERR_INVALID_ARG;
    const int h_out = (h_in + 2 * pad - kh) / stride + 1, w_out = (w_in + 2 * pad - kw) / stride + 1;
    if (h_out < 1 || w_out < 1) return SC2_ERR_INVALID_ARG;
    const int64_t total = static_cast<int64_t>(batch) * h_out * w_out * (k_pad / 8);
    int64_t blocks = (total + 255) / 256;
    if (blocks > sc2::kNumSMs * 32) blocks = sc2::kNumSMs * 32;
    sc2::tcs::patchify_split_kernel<<<static_cast<int>(blocks), 256, 0, sc2::as_stream(stream)>>>(
        x, static_cast<__half *>(out_hi), static_cast<__half *>(out_lo), c_in, h_in, w_in, kh, kw, stride, pad, h_out, w_out, k_pad,
        total, 1);
    SC2_LAUNCH_CHECK("patchify_split_kernel");
    return SC2_OK;
}

}  // extern "C"
