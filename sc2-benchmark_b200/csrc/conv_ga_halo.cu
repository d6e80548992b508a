// conv_ga_halo.cu -- stride-2 convolution + GDN1, back to back in ONE kernel, fp32-grade on the tcgen05 tensor cores.
//
// This is the middle of the analysis transform g_a (sc2bench/models/layer.py:479-481: Conv2d(4b -> 2b, k5, s2, p2) followed by
// GDN1(2b)), the kernel that bounded the pipelined step in round 1 (K3: 25 shifted TMA boxes per tile re-read every input pixel
// 4.4x from L2; GDN1 was a second launch that re-read the activation from HBM).  What changed:
//
//   HALO TILES   The input is stored as four parity planes (plane py*2+px holds input pixel (2Y+py, 2X+px)), so every tap of the
//                stride-2 convolution is a unit-stride shift of ONE plane.  Per (plane, 64-channel chunk) the CTA loads ONE halo
//                tile {64 ch, P px, th+2 rows} with TMA (128-byte pixel rows, SWIZZLE_128B) and forms the A operand of tap
//                (sy, sx) by shifting the UMMA descriptor's start address by (sy * P + sx) rows: the tile is a flat array of
//                pixel rows whose pitch P (16 / 32 / 64) equals the GEMM tile's row pitch, so "tile row m -> pixel row m + shift"
//                holds for all 128 rows (the P - tw rightmost columns of each tile row are junk and are never stored).
//                Hardware fact behind it (profiles/r2a_umma_shift_experiment.log): the 128-byte swizzle is a function of the
//                absolute shared-memory address, so a descriptor may start at any multiple of 128 bytes (base_offset = 0).
//   STACKED B    fp32-grade = split fp16: a = hi + lo / 2048, dot = sum hi.hi + (sum hi.lo + sum lo.hi) / 2048.  An MMA with a
//                narrow N is bound by reading its A operand from shared memory (measured: cycles ~ (4096 + 32 N) / 128 + 10), so
//                the weights are packed [hi rows; lo rows] and ONE N = 2n MMA computes hi.hi and hi.lo from a single read of
//                A_hi; a second N = n MMA adds lo.hi.  150 -> 115 cycles per K step for n = 48.
//   GDN EPILOGUE The conv accumulators leave TMEM once: x (fp32) stays in registers, |x| is split to (hi, lo) fp16 and written to
//                shared memory as the A operand of the 1x1 "gamma" GEMM (weights resident), whose accumulator comes back as
//                norm - beta; y = x / norm is split again and leaves through a staged TMA store.  x never visits HBM.
//
// Accumulation: tcgen05 adds into fp32 accumulators with truncation; the hi.hi sum over K = 25 * 96 is dealt round-robin over
// `groups` accumulators that the epilogue adds in fp32.
#include "tc_common.cuh"

namespace sc2 {
namespace gah {

using namespace sc2::tc;

constexpr int kThreads = 384;  // warp 0: A (halo) producer + tile scheduler; 1: MMA issuer; 2: B (weights) producer; 3: idle; 4..11: epilogue
constexpr int kMaxTaps = 25;
constexpr int kMaxB = 8;
constexpr int kMaxA = 4;
constexpr float kLoScale = 2048.0f;
constexpr float kLoInv = 1.0f / 2048.0f;

struct Tap {
    int8_t plane, sx, sy, group;  // sx, sy in [0, 2]: halo shift of this tap (pixel offset + 1)
    int8_t widx, first, pad0_, pad1_;  // widx: tap index in the packed weights; first: first tap of its accumulator group
};

struct Params {
    int tiles_x, tiles_y, images;
    int tw, th, pitch_log2;  // tile = th rows x tw valid columns of output pixels; pitch = 1 << pitch_log2 pixel rows per halo row
    int n_taps;
    Tap taps[kMaxTaps];
    uint32_t tap_word[kMaxTaps];  // sx | sy << 2 | group << 4 | first << 7 | widx << 8 (what the MMA issuer reads: one constant load)
    uint8_t run_len[kMaxTaps + 3];  // run_len[t]: taps of the plane run that starts at tap t
    int k_chunks, k_steps_last;  // 64-channel chunks of the input, K steps (of 16) in the last one
    int groups;
    int c_out, out_c, stage_c;
    int stage_plane;  // bytes of one staging plane (hi or lo): th * tw * stage_c * 2 rounded up to 128
    int h_out, w_out;
    int halo_bytes;  // (th + 2) * pitch * 128
    int n_a;         // halo ring units (each: hi + lo)
    int n_b;         // B ring slots
    int off_gamma, off_ag, off_b, off_bar;  // shared-memory offsets (from the 1024-aligned base)
    const float *beta;
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float fast_rcp(float v) {  // v > 0 (a GDN norm): reciprocal to within 1 ulp in 3 instructions
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return fmaf(r, fmaf(-v, r, 1.0f), r);
}

// 8 fp32 values -> 8 (hi, lo) fp16 pairs as two 16-byte vectors
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

// N: output channels padded to a multiple of 16 (the GEMM's N; the stacked MMA is 2N wide)
template <int N>
__global__ void __launch_bounds__(kThreads, 1)
ga_halo_gdn_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_g,
                   const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                   const __grid_constant__ Params p) {
    constexpr int kBSlot = 2 * N * 128;           // stacked weights of one (tap, 64-channel chunk)
    constexpr int kGC = (N + 63) / 64;            // K chunks of the gamma GEMM
    constexpr int kGStepsLast = (N - (kGC - 1) * 64) / 16;
    constexpr int kUnits = N / 16;                // 16-channel units of the epilogue
    constexpr int kUnits0 = (kUnits + 1) / 2;     // units of epilogue half 0 (half 1 takes the rest)
    constexpr uint32_t kTmemCols = 512;
    static_assert(N % 16 == 0 && N >= 16 && N <= 96, "N");
    extern __shared__ uint8_t smem_raw[];
    // (aligned by OFFSET, not through an integer cast: the compiler keeps the shared address space -> LDS / STS, not generic LD / ST)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *s_ag = smem + p.off_ag;  // |x| (hi chunks, then lo chunks), K-major SWIZZLE_128B; aliased by the output staging tile
    uint8_t *s_b = smem + p.off_b;
    uint64_t *a_full = reinterpret_cast<uint64_t *>(smem + p.off_bar);
    uint64_t *a_empty = a_full + kMaxA;
    uint64_t *b_full = a_empty + kMaxA;
    uint64_t *b_empty = b_full + kMaxB;
    uint64_t *acc_full = b_empty + kMaxB;
    uint64_t *acc_empty = acc_full + 1;
    uint64_t *ag_full = acc_empty + 1;
    uint64_t *g_full = ag_full + 1;
    uint64_t *stage_free = g_full + 1;  // the bulk stores of the previous tile have read the staging tile (= the |x| operand's memory)
    uint64_t *y_done = stage_free + 1;  // every epilogue thread has written its part of the staging tile
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(y_done + 1);
    float *s_beta = reinterpret_cast<float *>(tmem_slot + 4);  // [N]
    TileSched sched;
    sched.bind(reinterpret_cast<uint8_t *>(s_beta + N), p.tile_counter, p.tiles_x * p.tiles_y * p.images);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_xy = p.tiles_x * p.tiles_y;
    const int pitch = 1 << p.pitch_log2;
    const unsigned long long trace_t0 = p.trace.buf ? trace_now() : 0ull;
    int trace_tiles = 0;

    if (threadIdx.x < N) s_beta[threadIdx.x] = static_cast<int>(threadIdx.x) < p.c_out ? __ldg(p.beta + threadIdx.x) : 1.0f;
    if (threadIdx.x == 0) {
        sched.init(10);  // consumers: MMA warp, B producer warp, 8 epilogue warps
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_w);
        tma_prefetch_desc(&map_g);
        tma_prefetch_desc(&map_o_hi);
        tma_prefetch_desc(&map_o_lo);
        for (int s = 0; s < kMaxA; ++s) {
            mbar_init(&a_full[s], 1);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < kMaxB; ++s) {
            mbar_init(&b_full[s], 1);
            mbar_init(&b_empty[s], 1);
        }
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 256);
        mbar_init(ag_full, 256);
        mbar_init(g_full, 1);
        mbar_init(stage_free, 1);
        mbar_init(y_done, 256);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t gamma_col = static_cast<uint32_t>(p.groups) * 2u * N;  // accumulator of the gamma GEMM

    if (warp == 0) {
        // =============================== halo producer + tile scheduler ===============================
        if (elect_one()) {
            uint32_t as = 0, a_ph = 0;
            int tile = sched.claim(0);
            for (uint32_t qn = 0;; ++qn) {
                sched.publish(qn, tile);
                if (tile < 0) break;
                const int next_tile = sched.claim(qn + 1);
                const int sp = tile % tiles_xy, img = tile / tiles_xy;
                const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
                for (int t = 0; t < p.n_taps;) {
                    const int plane = p.taps[t].plane;
                    while (t < p.n_taps && p.taps[t].plane == plane) ++t;
                    for (int kc = 0; kc < p.k_chunks; ++kc) {
                        mbar_wait(&a_empty[as], a_ph ^ 1u);
                        uint8_t *dst = smem + as * 2 * p.halo_bytes;
                        mbar_expect_tx(&a_full[as], static_cast<uint32_t>(2 * p.halo_bytes));
                        tma_load_4d(&map_a_hi, &a_full[as], dst, kc * kBlockK, x0 - 1, y0 - 1, img * 4 + plane);
                        tma_load_4d(&map_a_lo, &a_full[as], dst + p.halo_bytes, kc * kBlockK, x0 - 1, y0 - 1, img * 4 + plane);
                        if (++as == static_cast<uint32_t>(p.n_a)) { as = 0; a_ph ^= 1u; }
                    }
                }
                tile = next_tile;
            }
        }
    } else if (warp == 2) {
        // =============================== weights producer ===============================
        uint32_t bs = 0, b_ph = 0;
        for (uint32_t lt = 0; sched.next(lt, lane) >= 0; ++lt) {
            if (elect_one()) {
                for (int t0 = 0; t0 < p.n_taps;) {
                    int t1 = t0;
                    while (t1 < p.n_taps && p.taps[t1].plane == p.taps[t0].plane) ++t1;
                    for (int kc = 0; kc < p.k_chunks; ++kc)
                        for (int t = t0; t < t1; ++t) {
                            mbar_wait(&b_empty[bs], b_ph ^ 1u);
                            mbar_expect_tx(&b_full[bs], static_cast<uint32_t>(kBSlot));
                            tma_load_2d(&map_w, &b_full[bs], s_b + bs * kBSlot, kc * kBlockK, p.taps[t].widx * 2 * N);
                            if (++bs == static_cast<uint32_t>(p.n_b)) { bs = 0; b_ph ^= 1u; }
                        }
                    t0 = t1;
                }
                // ... and the gamma matrix of this tile's GDN1 rides through the same ring (kGC more slots): 2 % more L2 traffic
                // buys a whole slot of shared memory for the rings
                for (int c = 0; c < kGC; ++c) {
                    mbar_wait(&b_empty[bs], b_ph ^ 1u);
                    mbar_expect_tx(&b_full[bs], static_cast<uint32_t>(kBSlot));
                    tma_load_2d(&map_g, &b_full[bs], s_b + bs * kBSlot, c * kBlockK, 0);
                    if (++bs == static_cast<uint32_t>(p.n_b)) { bs = 0; b_ph ^= 1u; }
                }
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        // (One issuer.  A variant with one issuer warp per accumulator group was ~8 % faster in isolation -- this warp's own
        // instruction stream, ~10 SASS instructions per tcgen05.mma, is what bounds the kernel -- but it hung about once in four
        // runs when the kernel ran beside the coder kernels of other streams, with every barrier protocol checked on paper;
        // it is kept in history (commit dde4825) until that is understood.)
        // Convergent: all 32 lanes run this code with uniform values; only the tcgen05 instructions are predicated on `lead`.
        constexpr uint32_t idesc_stack = make_idesc(2 * N), idesc_n = make_idesc(N);
        const bool lead = elect_one();
        const uint32_t a_base16 = smem_u32(smem) >> 4, halo16 = static_cast<uint32_t>(p.halo_bytes) >> 4;
        const uint32_t b_base16 = smem_u32(s_b) >> 4, ag16 = smem_u32(s_ag) >> 4;
        uint32_t as = 0, a_ph = 0, bs = 0, b_ph = 0;  // ring positions and phase bits
        for (uint32_t lt = 0; sched.next(lt, lane) >= 0; ++lt) {
            mbar_wait(acc_empty, (lt & 1u) ^ 1u);  // the epilogue has drained the conv accumulators of the previous tile
            tcgen05_fence_after();
            uint32_t prev_group = 0;
            for (int t0 = 0; t0 < p.n_taps;) {
                const int t1 = t0 + p.run_len[t0];
                for (int kc = 0; kc < p.k_chunks; ++kc) {
                    mbar_wait(&a_full[as], a_ph);
                    const uint32_t a_hi16 = a_base16 + as * 2u * halo16, a_lo16 = a_hi16 + halo16;
                    const bool last_chunk = kc == p.k_chunks - 1;
                    const int k_steps = last_chunk ? p.k_steps_last : kBlockK / 16;
                    uint32_t g1 = t0 == 0 ? 0u : prev_group;
                    for (int t = t0; t < t1; ++t) {
                        mbar_wait(&b_full[bs], b_ph);
                        tcgen05_fence_after();
                        const uint32_t tw = p.tap_word[t];  // sx | sy << 2 | group << 4 | first << 7 | widx << 8
                        const uint32_t shift16 = ((((tw >> 2) & 3u) << p.pitch_log2) + (tw & 3u)) * 8u;  // rows * 128 bytes >> 4
                        const uint32_t g = (tw >> 4) & 7u;
                        // hi.hi | hi.lo -> group g (2N columns); lo.hi -> the D1 half of the PREVIOUS tap's group, so that
                        // consecutive MMAs never write the same columns (a dependent MMA costs ~16 cycles more)
                        const uint32_t d_stack = tmem_base + g * 2u * N, d_lohi = tmem_base + g1 * 2u * N + N;
                        const uint32_t da_hi = a_hi16 + shift16, da_lo = a_lo16 + shift16, db = b_base16 + bs * (kBSlot >> 4);
                        const uint32_t acc0 = ((tw >> 7) & 1u) && kc == 0 ? 0u : 1u;
                        switch (k_steps) {  // (uniform) all MMAs of the tap in one asm block
                            case 4: umma_split_tap4(lead, d_stack, d_lohi, da_hi, da_lo, db, kDescHiSw128, idesc_stack, idesc_n, acc0); break;
                            case 3: umma_split_tap3(lead, d_stack, d_lohi, da_hi, da_lo, db, kDescHiSw128, idesc_stack, idesc_n, acc0); break;
                            case 2: umma_split_tap2(lead, d_stack, d_lohi, da_hi, da_lo, db, kDescHiSw128, idesc_stack, idesc_n, acc0); break;
                            default: umma_split_tap1(lead, d_stack, d_lohi, da_hi, da_lo, db, kDescHiSw128, idesc_stack, idesc_n, acc0); break;
                        }
                        umma_commit_lead(lead, &b_empty[bs]);
                        g1 = g;
                        if (++bs == static_cast<uint32_t>(p.n_b)) { bs = 0; b_ph ^= 1u; }
                    }
                    if (last_chunk) prev_group = g1;
                    umma_commit_lead(lead, &a_empty[as]);
                    if (t1 == p.n_taps && last_chunk) umma_commit_lead(lead, acc_full);
                    if (++as == static_cast<uint32_t>(p.n_a)) { as = 0; a_ph ^= 1u; }
                }
                t0 = t1;
            }
            // ---- the gamma GEMM of this tile: |x| (written by the epilogue warps) x gamma^T ----
            mbar_wait(ag_full, lt & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int c = 0; c < kGC; ++c) {
                mbar_wait(&b_full[bs], b_ph);
                tcgen05_fence_after();
                const uint32_t da_hi = ag16 + c * (kABytes >> 4), da_lo = ag16 + (kGC + c) * (kABytes >> 4), db = b_base16 + bs * (kBSlot >> 4);
                constexpr int k_steps_last = kGStepsLast;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (c < kGC - 1 || k < k_steps_last) {
                        umma_f16_lead(lead, tmem_base + gamma_col, da_hi + 2 * k, db + 2 * k, kDescHiSw128, idesc_stack, (c > 0 || k > 0) ? 1u : 0u);
                        umma_f16_lead(lead, tmem_base + gamma_col + N, da_lo + 2 * k, db + 2 * k, kDescHiSw128, idesc_n, 1u);
                    }
                }
                umma_commit_lead(lead, &b_empty[bs]);
                if (c == kGC - 1) umma_commit_lead(lead, g_full);
                if (++bs == static_cast<uint32_t>(p.n_b)) { bs = 0; b_ph ^= 1u; }
            }
        }
    } else if (warp >= 4) {
        // =============================== epilogue warps (4..11) ===============================
        const int half = (warp - 4) >> 2;
        const int quarter = warp & 3;          // TMEM lane quarter this warp may access
        const int m = quarter * 32 + lane;     // tile row = TMEM lane
        const int ty = m >> p.pitch_log2, tx = m & (pitch - 1);
        const bool issuer = threadIdx.x == 128;
        const int u_begin = half == 0 ? 0 : kUnits0, u_end = half == 0 ? kUnits0 : kUnits;
        const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
        __half *st_hi = reinterpret_cast<__half *>(s_ag);
        __half *st_lo = reinterpret_cast<__half *>(s_ag + p.stage_plane);
        for (uint32_t lt = 0;; ++lt) {
            const int tile = sched.next(lt, lane);
            if (tile < 0) break;
            ++trace_tiles;
            const int sp = tile % tiles_xy, img = tile / tiles_xy;
            const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
            float x[kUnits0][16];
            // ---- A: conv accumulators -> x (registers), |x| split -> the gamma GEMM's A operand ----
            // (mbarriers, not block-wide bar.sync: only the store-issuing thread waits for the whole tile)
            if (issuer) {
                tma_store_wait_read();  // the previous tile's bulk stores have read the staging tile (same memory)
                mbar_arrive(stage_free);
            }
            mbar_wait(acc_full, lt & 1u);
            tcgen05_fence_after();
#pragma unroll
            for (int ui = 0; ui < kUnits0; ++ui) {
                const int u = u_begin + ui;
                if (u < u_end) {
                    uint32_t d0[16], d1[16];
                    tmem_ld16_nowait(lane_addr + u * 16, d0);
                    tmem_ld16_nowait(lane_addr + N + u * 16, d1);
                    for (int g = 1; g < p.groups; ++g) {  // chunked summation: fp32 round-to-nearest adds of the partial sums
                        uint32_t e0[16], e1[16];
                        tmem_ld16_nowait(lane_addr + g * 2 * N + u * 16, e0);
                        tmem_ld16_nowait(lane_addr + g * 2 * N + N + u * 16, e1);
                        tmem_ld_wait();  // (also covers the loads of d0 / d1 issued before)
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            d0[e] = __float_as_uint(__uint_as_float(d0[e]) + __uint_as_float(e0[e]));
                            d1[e] = __float_as_uint(__uint_as_float(d1[e]) + __uint_as_float(e1[e]));
                        }
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) x[ui][e] = fmaf(__uint_as_float(d1[e]), kLoInv, __uint_as_float(d0[e]));
                }
            }
            tcgen05_fence_before();
            mbar_arrive(acc_empty);  // 256 arrivals: the MMA warp may start the next tile's convolution
            mbar_wait(stage_free, lt & 1u);  // the |x| tile below overwrites the previous tile's staged output
#pragma unroll
            for (int ui = 0; ui < kUnits0; ++ui) {
                const int u = u_begin + ui;
                if (u < u_end) {
                    float a[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) a[e] = fabsf(x[ui][e]);
                    const int c = u * 16, kc = c >> 6, j = (c & 63) >> 3;  // 16-byte chunk j (and j + 1) of K chunk kc
                    uint8_t *row_hi = s_ag + kc * kABytes + m * 128, *row_lo = s_ag + (kGC + kc) * kABytes + m * 128;
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        uint4 h, l;
                        split8(&a[8 * q], h, l);
                        const int phys = ((j + q) ^ (m & 7)) << 4;
                        *reinterpret_cast<uint4 *>(row_hi + phys) = h;
                        *reinterpret_cast<uint4 *>(row_lo + phys) = l;
                    }
                }
            }
            fence_proxy_async();
            mbar_arrive(ag_full);
            // ---- B: gamma accumulator -> y = x / (beta + gamma.|x|) -> staging -> TMA store ----
            mbar_wait(g_full, lt & 1u);
            tcgen05_fence_after();
            const bool store_row = tx < p.tw;
            const int srow = ty * p.tw + tx;
#pragma unroll
            for (int ui = 0; ui < kUnits0; ++ui) {
                const int u = u_begin + ui;
                if (u < u_end) {
                    uint32_t d0[16], d1[16];
                    tmem_ld16_nowait(lane_addr + gamma_col + u * 16, d0);
                    tmem_ld16_nowait(lane_addr + gamma_col + N + u * 16, d1);
                    tmem_ld_wait();
                    const int c = u * 16;
                    if (store_row && c < p.out_c) {
                        float y[16], bt[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) *reinterpret_cast<float4 *>(&bt[4 * q]) = *reinterpret_cast<const float4 *>(s_beta + c + 4 * q);
                        const bool whole = c + 16 <= p.c_out;  // (uniform) every channel of the unit is real
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            const float norm = fmaf(__uint_as_float(d1[e]), kLoInv, __uint_as_float(d0[e])) + bt[e];
                            // x * (1 / norm), like the reference; 1 / norm = MUFU.RCP + one Newton step (< 1 ulp)
                            y[e] = (whole || c + e < p.c_out) ? x[ui][e] * fast_rcp(norm) : 0.0f;
                        }
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            if (c + 8 * q < p.out_c) {
                                uint4 h, l;
                                split8(&y[8 * q], h, l);
                                *reinterpret_cast<uint4 *>(st_hi + srow * p.stage_c + c + 8 * q) = h;
                                *reinterpret_cast<uint4 *>(st_lo + srow * p.stage_c + c + 8 * q) = l;
                            }
                        }
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
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (threadIdx.x == 128) trace_emit(p.trace, TRACE_CONV_SPLIT, trace_t0, trace_tiles);
}

template <int N>
static int launch(const CUtensorMap *maps, const Params &p, int smem, cudaStream_t st) {
    static std::atomic<uint64_t> configured{0};  // per device ordinal
    if (int rc = ensure_dyn_smem(ga_halo_gdn_kernel<N>, 227 * 1024, configured)) return rc;
    const int64_t total = static_cast<int64_t>(p.tiles_x) * p.tiles_y * p.images;
    if (total > 0x7fffffff) return SC2_ERR_UNSUPPORTED;
    const int grid = total < persistent_grid() ? static_cast<int>(total) : persistent_grid();
    ga_halo_gdn_kernel<N><<<grid, kThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], maps[4], maps[5], p);
    SC2_LAUNCH_CHECK("ga_halo_gdn_kernel");
    return SC2_OK;
}

}  // namespace gah
}  // namespace sc2

extern "C" {

int sc2_ga_halo_n(int c_out) {
    const int n = (c_out + 15) / 16 * 16;
    return (c_out >= 1 && n <= 96) ? n : 0;
}

int sc2_ga_halo_conv_gdn(const sc2_ga_halo_desc *d, const void *x_hi, const void *x_lo, const void *w_stack, const void *gamma_stack,
                         const float *beta, void *out_hi, void *out_lo, int32_t *tile_counter, sc2_stream_t stream) {
    using namespace sc2::gah;
    if (!d || !x_hi || !x_lo || !w_stack || !gamma_stack || !beta || !out_hi || !out_lo) return SC2_ERR_INVALID_ARG;
    if (d->images < 1 || d->c_in < 16 || d->c_in % 16 || d->c_out < 1 || d->h_in < 1 || d->w_in < 1) return SC2_ERR_INVALID_ARG;
    if (d->kh < 1 || d->kw < 1 || d->kh * d->kw > kMaxTaps) return SC2_ERR_UNSUPPORTED;
    const int n = sc2_ga_halo_n(d->c_out);
    if (!n) return SC2_ERR_UNSUPPORTED;
    if (d->out_c % 8 || d->out_c < d->c_out || d->out_c > n) return SC2_ERR_INVALID_ARG;
    Params p;
    p.images = d->images;
    p.h_out = d->h_out; p.w_out = d->w_out;
    p.c_out = d->c_out; p.out_c = d->out_c;
    p.stage_c = (d->out_c / 8) % 2 == 0 ? d->out_c + 8 : d->out_c;  // odd number of 16-byte units per staging row
    // taps: stride 2 on parity planes; halo shift of a tap = its plane offset + 1, must lie in [0, 2]
    p.n_taps = 0;
    Tap by_plane[4][kMaxTaps];
    int n_plane[4] = {0, 0, 0, 0};
    for (int dy = 0; dy < d->kh; ++dy)
        for (int dx = 0; dx < d->kw; ++dx) {
            const int sy = dy - d->pad, sx = dx - d->pad;
            const int py = ((sy % 2) + 2) % 2, px = ((sx % 2) + 2) % 2;
            const int oy = (sy - py) / 2 + 1, ox = (sx - px) / 2 + 1;
            if (oy < 0 || oy > 2 || ox < 0 || ox > 2) return SC2_ERR_UNSUPPORTED;
            Tap t;
            t.plane = static_cast<int8_t>(py * 2 + px);
            t.sy = static_cast<int8_t>(oy); t.sx = static_cast<int8_t>(ox);
            t.widx = static_cast<int8_t>(dy * d->kw + dx);
            t.group = 0; t.first = 0; t.pad0_ = t.pad1_ = 0;
            by_plane[t.plane][n_plane[t.plane]++] = t;
        }
    p.k_chunks = (d->c_in + kBlockK - 1) / kBlockK;
    p.k_steps_last = (d->c_in - (p.k_chunks - 1) * kBlockK) / 16;
    {   // accumulator groups: as many as TMEM holds next to the gamma accumulator, at most one per ~24 K steps
        const int steps = d->kh * d->kw * ((p.k_chunks - 1) * 4 + p.k_steps_last);
        int groups = (steps + 23) / 24;
        const int max_groups = 512 / (2 * n) - 1;
        if (groups > max_groups) groups = max_groups;
        if (groups > d->kh * d->kw) groups = d->kh * d->kw;
        if (groups > 4) groups = 4;
        if (groups < 1) groups = 1;
        p.groups = groups;
    }
    for (int pl = 0; pl < 4; ++pl)
        for (int i = 0; i < n_plane[pl]; ++i) {
            Tap t = by_plane[pl][i];
            t.group = static_cast<int8_t>(p.n_taps % p.groups);  // round-robin: consecutive taps use different accumulators
            t.first = p.n_taps < p.groups;
            p.taps[p.n_taps++] = t;
        }
    for (int t = 0; t < kMaxTaps; ++t) {
        p.tap_word[t] = 0;
        p.run_len[t] = 0;
    }
    for (int t = 0; t < p.n_taps; ++t) {
        const Tap &tp = p.taps[t];
        p.tap_word[t] = static_cast<uint32_t>(tp.sx) | (static_cast<uint32_t>(tp.sy) << 2) | (static_cast<uint32_t>(tp.group) << 4) |
                        (static_cast<uint32_t>(tp.first) << 7) | (static_cast<uint32_t>(tp.widx) << 8);
        if (t == 0 || p.taps[t - 1].plane != tp.plane) {
            int t1 = t;
            while (t1 < p.n_taps && p.taps[t1].plane == tp.plane) ++t1;
            p.run_len[t] = static_cast<uint8_t>(t1 - t);
        }
    }
    // tile shape: th rows x (pitch - 2) columns with th * pitch = 128; pick the pitch that loads the fewest halo pixels
    int best_log2 = 4;
    int64_t best_cost = -1;
    for (int lg = 4; lg <= 6; ++lg) {
        const int pitch = 1 << lg, th = 128 / pitch, tw_max = pitch - 2;
        const int n_col = (d->w_out + tw_max - 1) / tw_max;
        const int64_t cost = static_cast<int64_t>(n_col) * ((d->h_out + th - 1) / th) * (th + 2) * pitch;
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_log2 = lg; }
    }
    p.pitch_log2 = best_log2;
    const int pitch = 1 << best_log2;
    p.th = 128 / pitch;
    {
        const int tw_max = pitch - 2;
        const int n_col = (d->w_out + tw_max - 1) / tw_max;
        p.tw = (d->w_out + n_col - 1) / n_col;
        p.tiles_x = n_col;
    }
    p.tiles_y = (d->h_out + p.th - 1) / p.th;
    p.halo_bytes = (p.th + 2) * pitch * 128;
    // shared memory: [halo ring: n_a units of (hi, lo)] [|x| operand / output staging] [weights ring: n_b slots] [barriers, beta,
    // scheduler].  A halo unit takes ~1.5 us to arrive and a short unit (4 taps x 32 channels) is consumed in 0.5 us: with two
    // units the MMAs waited for the halo loads most of the time, so the ring gets 3 units when 4 weight slots still fit.
    const int gc = (n + 63) / 64, b_slot = 2 * n * 128;
    p.stage_plane = (p.th * p.tw * p.stage_c * 2 + 127) / 128 * 128;
    const int ag_bytes = 2 * gc * kABytes, staging_bytes = 2 * p.stage_plane;
    const int ag_region = ((ag_bytes > staging_bytes ? ag_bytes : staging_bytes) + 1023) / 1024 * 1024;
    const int tail = (2 * kMaxA + 2 * kMaxB + 6) * 8 + 16 + n * 4 + kTileSchedBytes + 64;
    const int budget = 227 * 1024 - 1024 - tail;  // (1024: slack for the manual alignment of the dynamic shared memory base)
    int n_a = kMaxA, n_b = 0;
    for (; n_a >= 2; --n_a) {
        n_b = (budget - n_a * 2 * p.halo_bytes - ag_region) / b_slot;
        if (n_b >= (n_a > 2 ? 4 : 2)) break;
    }
    if (n_a < 2) return SC2_ERR_UNSUPPORTED;
    if (n_a > 3 && n_b < 6) { n_a = 3; n_b = (budget - n_a * 2 * p.halo_bytes - ag_region) / b_slot; }
    if (n_b > kMaxB) n_b = kMaxB;
    p.n_a = n_a; p.n_b = n_b;
    p.off_gamma = 0;
    p.off_ag = n_a * 2 * p.halo_bytes;
    p.off_b = p.off_ag + ag_region;
    p.off_bar = p.off_b + n_b * b_slot;
    const int smem = p.off_bar + tail + 1024;
    p.beta = beta;
    p.tile_counter = tile_counter;
    p.trace = sc2::trace_sink();
    CUtensorMap maps[6];
    // input parity planes [images * 4, h_in, w_in, c_in]: halo boxes {64 ch, pitch px, th + 2 rows}
    int rc = make_nhwc_map(&maps[0], x_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->c_in, d->w_in, d->h_in, d->images * 4, kBlockK, pitch, p.th + 2);
    if (rc) return rc;
    rc = make_nhwc_map(&maps[1], x_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->c_in, d->w_in, d->h_in, d->images * 4, kBlockK, pitch, p.th + 2);
    if (rc) return rc;
    // stacked weights [taps * 2n, c_in] and gamma [2n, n]: boxes {64, 2n}
    rc = make_weight_map(&maps[2], w_stack, d->c_in, d->kh * d->kw * 2 * n, 2 * n);
    if (rc) return rc;
    rc = make_weight_map(&maps[3], gamma_stack, n, 2 * n, 2 * n);
    if (rc) return rc;
    rc = make_nhwc_map(&maps[4], out_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->out_c, d->w_out, d->h_out, d->images, p.stage_c, p.tw, p.th,
                       CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_nhwc_map(&maps[5], out_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->out_c, d->w_out, d->h_out, d->images, p.stage_c, p.tw, p.th,
                       CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    cudaStream_t st = sc2::as_stream(stream);
    switch (n) {
        case 16: return launch<16>(maps, p, smem, st);
        case 32: return launch<32>(maps, p, smem, st);
        case 48: return launch<48>(maps, p, smem, st);
        case 64: return launch<64>(maps, p, smem, st);
        case 80: return launch<80>(maps, p, smem, st);
        default: return launch<96>(maps, p, smem, st);
    }
}

}  // extern "C"
