// conv_tc.cu -- persistent tcgen05 / TMEM / TMA implicit-GEMM convolution for NHWC fp16 activations (sm_100a).
//
// The synthesis transform g_s of the bottleneck (sc2bench/models/layer.py:485-494: Conv 2x2 -> IGDN1 -> Conv 2x2 ->
// IGDN1 -> Conv 2x2) is 86 % of the path's FLOPs and tolerates fp16 operands (fp32 accumulation; 4e-4 relative on the
// decoded features), so it runs on the 5th-generation tensor cores:
//
//   GEMM view   M = output pixels (a TH x TW patch of one image = up to 128 rows), N = output channels (<= 256 per tile),
//               K = taps x input channels.
//   A operand   no im2col: for tap (dy, dx) the A tile is the SAME activation tensor read through a 4-D TMA box
//               {64 channels, TW, TH, 1} shifted by (dx - pad, dy - pad); TMA zero-fills outside the image = padding.
//               The box lands in shared memory as 128-byte rows (K-major, SWIZZLE_128B) = the UMMA canonical layout.
//   B operand   weights repacked once per model to [tap][c_out][c_in] fp16, 2-D TMA box {64, N_TILE}.
//   D           fp32 accumulators in TMEM, DOUBLE BUFFERED (2 x N_TILE columns): the epilogue of tile i overlaps the
//               MMAs of tile i + 1.
//   schedule    persistent: one CTA per SM walks the tile list (stride gridDim.x); barrier / TMEM / descriptor set-up
//               is paid once per CTA instead of once per tile (it dominated the small-K layers).
//   roles       warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane), warps 2..9 = epilogue (TMEM -> registers
//               -> global, thread = output pixel, the two warpgroups take alternate 32-column chunks),
//               warps 10..13 (GDN modes) = |x| transform of the landed A tiles.
//   epilogues   store fp16 | store fp32 | IGDN1: out = x * (beta + gamma.|x|) | GDN1: out = x / (beta + gamma.|x|),
//               where the GEMM is the 1x1 "gamma" contraction over |x| (sign bits cleared in shared memory right after
//               the TMA lands) and x is re-read by the epilogue.
#include "tc_common.cuh"

namespace sc2 {
namespace tc {

enum Mode { MODE_STORE_F16 = 0, MODE_STORE_F32 = 1, MODE_IGDN1_F16 = 2, MODE_GDN1_F16 = 3, MODE_STORE_ABS_F16 = 4, MODE_IGDN1_ABS_F16 = 5,
            MODE_STORE_SQ_F16 = 6, MODE_IGDN_SQ_F16 = 7, MODE_NCHW_F32_CLAMP = 8 };
// Modes 6 / 7 / 8 (round 2) serve the synthesis transforms of the CompressAI zoo codecs (bmshj2018-*: GDN proper, transposed
// convolutions, biases; sc2bench/models/registry.py:12-14): a ConvTranspose2d(k5, s2, p2, op1) is FOUR stride-1 sub-convolutions,
// one per output parity (3x3, 3x2, 2x3, 2x2 taps), each a launch of this kernel that writes every second output pixel
// (out_stride 2, out_py / out_px); mode 6 stores x (fp16) AND x^2 / 256 (fp16, squared in fp32 from the accumulator: the scale
// keeps it inside fp16's range), mode 7 is the inverse GDN on such a pair: the 1x1 gamma GEMM reads the x^2 tensor,
// out = x * sqrt(beta + 256 * acc); mode 8 is the last layer: c_out <= N_TILE real channels, bias, clamp to [0, 1], fp32 NCHW.
// Modes 4 / 5 (round 2) split "x" into |x| (fp16) and one sign bit per value: the conv in front of an IGDN1 stores |x| and the
// packed signs, and the IGDN1's 1x1 gamma GEMM reads |x| straight from the TMA-loaded tile -- no in-smem |.| pass between the
// TMA and the MMA (that extra hop is what kept IGDN1(512) at 25 % tensor-pipe utilisation with a 4-stage ring while the plain
// conv with the same tiles reached 80 %), and the epilogue rebuilds x = sign * |x| from the sign words.
// Sign word of 32 consecutive channels (16 half2 words h_0..h_15): S = OR_k ((h_k & 0x80008000) >> k), so bit 15-k is the sign
// of channel 2k and bit 31-k of channel 2k+1; the reader recovers word k's signs as (S << k) & 0x80008000.

struct Params {
    int tiles_x, tiles_y;  // output tile grid per image
    int tw, th;            // tile = th rows x tw columns of output pixels (th * tw <= 128)
    int taps_x, taps_y, pad_x, pad_y;
    int k_chunks;          // c_in_padded / 64
    int n_total;           // c_out (rows per tap of the packed weight tensor)
    int n_tiles;           // c_out / N_TILE
    int batch;
    int h_out, w_out;      // grid of output pixels this launch computes
    int out_h, out_w;      // the output TENSOR: pixel (oy, ox) of the grid lands at (oy * out_stride + out_py, ox * out_stride + out_px)
    int out_stride, out_py, out_px;
    int c_real;            // mode 8: real output channels (<= N_TILE)
    void *out2;            // mode 6: the x^2 / 256 tensor; mode 7 with gdn_x_lo: the lo half of y
    void *out3;            // mode 6, optional: the lo half of x (x = out + out3: activation rounding out of the last stage)
    const __half *gdn_x_lo;  // mode 7, optional: lo half of x
    int a_passes;          // 1, or 2: the A operand is the sum of two fp16 tensors (map_a, map_a2), same weights
    const float *beta;     // GDN modes: effective beta [n_total]; conv modes: bias [n_total] or nullptr
    const __half *gdn_x;   // GDN modes: x itself (mode 5: |x|), NHWC [batch, h_out, w_out, n_total]
    uint32_t *signs;       // mode 4: out, mode 5: in -- packed sign words [batch * h_out * w_out, n_total / 32]
    void *out;             // NHWC [batch, h_out, w_out, n_total], fp16 or fp32
    int *tile_counter;     // zeroed by the caller: dynamic tile schedule; nullptr: static
    TraceSink trace;       // diagnostics (common.cuh)
};

constexpr int kMaxBeta = 512;  // channels of a GDN layer held in shared memory

// STAGED epilogue (fp16 outputs): results leave through 64-channel staging blocks [128 rows][128 bytes] in the SWIZZLE_128B layout
// (16-byte unit j of row r sits at unit j ^ (r & 7): conflict-free for one row per thread) and TMA bulk stores; the x tile of the GDN
// epilogue arrives in the same block by TMA and y overwrites it in place.  Three blocks rotate: while block g is computed, the
// store of g - 1 drains and the x tile of g + 1 loads.  (Per-thread 16-byte global loads / stores at a 512..1024-byte stride
// between lanes were 32 sectors per request: IGDN1(512) moved 7.9 GB through L2 for 4.8 GB of operands, profiles/r4_*.)
constexpr int kStageBlocks = 3;
constexpr int kStageBlockBytes = kTileM * 128;

template <int N_TILE, int STAGES, bool STAGED = false>
struct Smem {
    static constexpr int kBBytes = N_TILE * 128;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kRingBytes = STAGES * kStageBytes;
    static constexpr int kOutOffset = kRingBytes;  // (1024-byte aligned: every ring stage is a multiple of 1024 bytes)
    static constexpr int kOutBytes = STAGED ? kStageBlocks * kStageBlockBytes : 0;
    // full[STAGES], empty[STAGES], xform[STAGES], acc_full[2], acc_empty[2], x_full[3] : 8 bytes each; then the TMEM base address
    static constexpr int kBarOffset = kOutOffset + kOutBytes;
    static constexpr int kSchedOffset = kBarOffset + (3 * STAGES + 4 + kStageBlocks) * 8 + 16;
    static constexpr int kBetaOffset = (kSchedOffset + kTileSchedBytes + 15) / 16 * 16;  // float[kMaxBeta] (GDN modes)
    static constexpr int kTotal = kBetaOffset + kMaxBeta * 4;
};

__device__ __forceinline__ void tma_store_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read_2() { asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); }

template <int N_TILE, int STAGES, int MODE, bool STAGED>
__global__ void __launch_bounds__((MODE == MODE_IGDN1_F16 || MODE == MODE_GDN1_F16) ? 448 : 320, 1)  // + 4 |x| transform warps
tc_conv_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a2, const __grid_constant__ CUtensorMap map_b,
               const __grid_constant__ CUtensorMap map_o, const __grid_constant__ CUtensorMap map_x, const __grid_constant__ Params p) {
    using L = Smem<N_TILE, STAGES, STAGED>;
    static_assert(!STAGED || MODE == MODE_STORE_F16 || MODE == MODE_STORE_ABS_F16 || MODE == MODE_IGDN1_ABS_F16 || MODE == MODE_STORE_F32,
                  "staged epilogue: dense NHWC outputs");
    static_assert(!STAGED || N_TILE % 64 == 0, "staged epilogue works on 64-channel blocks");
    constexpr bool kXform = MODE == MODE_IGDN1_F16 || MODE == MODE_GDN1_F16;  // |x| formed in shared memory by 4 extra warps
    constexpr bool kGdn = kXform || MODE == MODE_IGDN1_ABS_F16 || MODE == MODE_IGDN_SQ_F16;  // GDN epilogue
    constexpr bool kSigned = MODE == MODE_IGDN1_ABS_F16;                      // x = sign word * |x|
    constexpr bool kOutF32 = MODE == MODE_STORE_F32;
    constexpr uint32_t kTmemCols = 2 * N_TILE <= 32 ? 32 : 2 * N_TILE <= 64 ? 64 : 2 * N_TILE <= 128 ? 128 : 2 * N_TILE <= 256 ? 256 : 512;
    static_assert(2 * N_TILE <= 512, "two accumulator stages must fit TMEM");
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + L::kBarOffset);
    uint64_t *empty = full + STAGES;
    uint64_t *xform = empty + STAGES;
    uint64_t *acc_full = xform + STAGES;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *x_full = acc_empty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(x_full + kStageBlocks);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rows = p.tw * p.th;
    const int k_iters = p.a_passes * p.taps_x * p.taps_y * p.k_chunks;
    const int tiles_xy = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_xy * p.n_tiles * p.batch;
    const unsigned long long trace_t0 = p.trace.buf ? trace_now() : 0ull;
    int trace_tiles = 0;
    TileSched sched;
    sched.bind(smem + L::kSchedOffset, p.tile_counter, total_tiles);

    // GDN modes: beta in shared memory (the epilogue read it with one LDG per element: a third of its instructions)
    const float *s_beta = reinterpret_cast<const float *>(smem + L::kBetaOffset);
    const bool has_vec = p.beta != nullptr;  // beta (GDN modes) or bias (conv modes)
    if (has_vec)
        for (int i = threadIdx.x; i < p.n_total; i += blockDim.x) reinterpret_cast<float *>(smem + L::kBetaOffset)[i] = __ldg(p.beta + i);

    if (threadIdx.x == 0) {
        sched.init(kXform ? 13 : 9);  // consumers: MMA warp, 8 epilogue warps (, 4 transform warps)
        tma_prefetch_desc(&map_a);
        tma_prefetch_desc(&map_a2);
        tma_prefetch_desc(&map_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&xform[s], 128);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 256);
        }
        for (int s = 0; s < kStageBlocks; ++s) mbar_init(&x_full[s], 1);
        if (STAGED) {
            tma_prefetch_desc(&map_o);
            tma_prefetch_desc(&map_x);
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
            const uint32_t stage_tx = static_cast<uint32_t>(rows * 128 + L::kBBytes);
            uint32_t it = 0;
            int tile = sched.claim(0);
            for (uint32_t qn = 0;; ++qn) {
                sched.publish(qn, tile);
                if (tile < 0) break;
                const int next_tile = sched.claim(qn + 1);  // claimed early: the atomic's latency hides behind this tile's loads
                const int sp = tile % tiles_xy, rest = tile / tiles_xy;
                const int n0 = (rest % p.n_tiles) * N_TILE, img = rest / p.n_tiles;
                const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
                for (int pass = 0; pass < p.a_passes; ++pass)
                    for (int ty = 0; ty < p.taps_y; ++ty)
                        for (int tx = 0; tx < p.taps_x; ++tx)
                            for (int kc = 0; kc < p.k_chunks; ++kc, ++it) {
                                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                                mbar_wait(&empty[s], ph ^ 1u);
                                uint8_t *dst = smem + s * L::kStageBytes;
                                mbar_expect_tx(&full[s], stage_tx);
                                tma_load_4d(pass == 0 ? &map_a : &map_a2, &full[s], dst, kc * kBlockK, x0 + tx - p.pad_x, y0 + ty - p.pad_y, img);
                                tma_load_2d(&map_b, &full[s], dst + kABytes, kc * kBlockK, (ty * p.taps_x + tx) * p.n_total + n0);
                            }
                tile = next_tile;
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc = make_idesc(N_TILE);
        uint32_t it = 0;
        for (uint32_t lt = 0; sched.next(lt, lane) >= 0; ++lt) {
            const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
            mbar_wait(&acc_empty[as], aph ^ 1u);  // the epilogue has drained this accumulator stage
            tcgen05_fence_after();
            for (int k_it = 0; k_it < k_iters; ++k_it, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                mbar_wait(kXform ? &xform[s] : &full[s], ph);
                tcgen05_fence_after();
                if (elect_one()) {
                    const uint32_t a_addr = smem_u32(smem + s * L::kStageBytes);
                    const uint64_t a_desc = make_smem_desc(a_addr);
                    const uint64_t b_desc = make_smem_desc(a_addr + kABytes);
#pragma unroll
                    for (int k = 0; k < kBlockK / 16; ++k)
                        umma_f16(tmem_base + as * N_TILE, a_desc + 2 * k, b_desc + 2 * k, idesc, (k_it > 0 || k > 0) ? 1u : 0u);
                    umma_commit(&empty[s]);                               // frees the smem stage once these MMAs have read it
                    if (k_it == k_iters - 1) umma_commit(&acc_full[as]);  // accumulator complete
                }
                __syncwarp();
            }
        }
    } else if (warp < 10) {
        // =============================== epilogue warps (2..9) ===============================
        const int half = (warp - 2) >> 2;     // warpgroup 0 / 1: even / odd 32-column chunks
        const int quarter = warp & 3;         // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;  // tile row = TMEM lane = pixel index inside the tile
        const int ty = row / p.tw, tx = row - ty * p.tw;
        if constexpr (STAGED && MODE == MODE_STORE_F32) {
            // fp32 output (the last g_s layer): a thread holds 32 channels = one 128-byte piece of its pixel's row.  Written directly,
            // a warp store touches 32 pixel rows 1 KB apart (32 sectors in 32 lines per request); instead every warp transposes its
            // 32 rows x 128 bytes through a private 4 KB block (swizzled: conflict-free both ways, __syncwarp only) so that one
            // store instruction writes four complete 128-byte lines.
            uint4 *wblk = reinterpret_cast<uint4 *>(smem + L::kOutOffset) + (warp - 2) * 256;
            float *outf = static_cast<float *>(p.out);
            for (uint32_t lt = 0;; ++lt) {
                const int tile = sched.next(lt, lane);
                if (tile < 0) break;
                ++trace_tiles;
                const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
                const int sp = tile % tiles_xy, rest = tile / tiles_xy;
                const int n0 = (rest % p.n_tiles) * N_TILE, img = rest / p.n_tiles;
                const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
                // read-back phase: instruction k moves rows quarter * 32 + 4 k + (lane >> 3), unit lane & 7
                const int r0 = quarter * 32 + (lane >> 3);
                const int ty0 = r0 / p.tw, tx0 = r0 - ty0 * p.tw;
                mbar_wait(&acc_full[as], aph);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + as * N_TILE + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll 1
                for (int c0 = half * 32; c0 < N_TILE; c0 += 64) {
                    uint32_t v[32];
                    tmem_ld32(taddr + c0, v);
                    if (has_vec) {
#pragma unroll
                        for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + s_beta[n0 + c0 + e]);
                    }
                    __syncwarp();  // the previous piece has been read back
#pragma unroll
                    for (int u = 0; u < 8; ++u) wblk[lane * 8 + (u ^ (lane & 7))] = make_uint4(v[4 * u], v[4 * u + 1], v[4 * u + 2], v[4 * u + 3]);
                    __syncwarp();
                    int tyk = ty0, txk = tx0;
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const int rl = 4 * k + (lane >> 3);  // row inside the warp's block
                        const int oy = y0 + tyk, ox = x0 + txk;
                        if (quarter * 32 + rl < rows && oy < p.h_out && ox < p.w_out) {
                            const int64_t pix = (static_cast<int64_t>(img) * p.h_out + oy) * p.w_out + ox;
                            reinterpret_cast<uint4 *>(outf + pix * p.n_total + n0 + c0)[lane & 7] = wblk[rl * 8 + ((lane & 7) ^ (rl & 7))];
                        }
                        txk += 4;
                        if (txk >= p.tw) { txk -= p.tw; ++tyk; }
                    }
                }
                tcgen05_fence_before();
                mbar_arrive(&acc_empty[as]);
            }
        } else if constexpr (STAGED) {
            constexpr int kChunks = N_TILE / 64;
            const bool issuer = threadIdx.x == 64;  // first epilogue thread: owns the bulk-store groups and the x-tile loads
            uint8_t *blocks = smem + L::kOutOffset;
            const uint32_t x_bytes = static_cast<uint32_t>(rows * 128);
            uint32_t g = 0;  // 64-channel blocks processed so far (the same in every epilogue thread): block g uses buffer g % 3
            for (uint32_t lt = 0;; ++lt) {
                const int tile = sched.next(lt, lane);
                if (tile < 0) break;
                ++trace_tiles;
                const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
                const int sp = tile % tiles_xy, rest = tile / tiles_xy;
                const int n0 = (rest % p.n_tiles) * N_TILE, img = rest / p.n_tiles;
                const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
                const int oy = y0 + ty, ox = x0 + tx;
                const bool valid = row < rows && oy < p.h_out && ox < p.w_out;
                const int64_t pix = (static_cast<int64_t>(img) * p.h_out + oy) * p.w_out + ox;
                uint32_t spre[kSigned ? kChunks : 1];
                if (kSigned && valid) {
#pragma unroll
                    for (int ci = 0; ci < kChunks; ++ci) spre[ci] = __ldg(p.signs + pix * (p.n_total >> 5) + ((n0 + half * 32 + ci * 64) >> 5));
                }
                if (kGdn && issuer) {  // x tile of the first block: requested before the accumulator is complete
                    tma_store_wait_read_1();
                    mbar_expect_tx(&x_full[g % kStageBlocks], x_bytes);
                    tma_load_4d(&map_x, &x_full[g % kStageBlocks], blocks + (g % kStageBlocks) * kStageBlockBytes, n0, x0, y0, img);
                }
                mbar_wait(&acc_full[as], aph);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + as * N_TILE + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
                for (int ci = 0; ci < kChunks; ++ci, ++g) {
                    const uint32_t b = g % kStageBlocks;
                    uint8_t *blk = blocks + b * kStageBlockBytes;
                    if (kGdn) {
                        if (issuer && ci + 1 < kChunks) {  // x tile of the next block (its buffer was last stored two blocks ago)
                            const uint32_t nb = (g + 1) % kStageBlocks;
                            tma_store_wait_read_1();
                            mbar_expect_tx(&x_full[nb], x_bytes);
                            tma_load_4d(&map_x, &x_full[nb], blocks + nb * kStageBlockBytes, n0 + (ci + 1) * 64, x0, y0, img);
                        }
                        mbar_wait(&x_full[b], (g / kStageBlocks) & 1u);
                    } else {
                        if (issuer) tma_store_wait_read_2();  // the store that read this buffer three blocks ago is done with it
                        asm volatile("bar.sync 1, 256;" ::: "memory");
                    }
                    uint32_t v[32];
                    tmem_ld32(taddr + half * 32 + ci * 64, v);
                    if (row < rows) {
                        uint32_t sign_word = 0;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            // 16-byte unit (half * 4 + c) of this row's 128-byte line, swizzled
                            uint4 *slot = reinterpret_cast<uint4 *>(blk + row * 128 + (((half * 4 + c) ^ (row & 7)) << 4));
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * c + e]);
                            if (kGdn) {
                                uint4 xv = *slot;
                                const uint32_t sw = spre[kSigned ? ci : 0];  // x = sign * |x|: half2 word k = 4c + e takes (S << k) & 0x80008000
                                xv.x |= (sw << (4 * c)) & 0x80008000u;
                                xv.y |= (sw << (4 * c + 1)) & 0x80008000u;
                                xv.z |= (sw << (4 * c + 2)) & 0x80008000u;
                                xv.w |= (sw << (4 * c + 3)) & 0x80008000u;
                                const __half2 *xh = reinterpret_cast<const __half2 *>(&xv);
                                const float4 ba = *reinterpret_cast<const float4 *>(s_beta + n0 + half * 32 + ci * 64 + 8 * c);
                                const float4 bb = *reinterpret_cast<const float4 *>(s_beta + n0 + half * 32 + ci * 64 + 8 * c + 4);
                                const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    const float2 xf = __half22float2(xh[e]);
                                    f[2 * e] = xf.x * (f[2 * e] + bv[2 * e]);
                                    f[2 * e + 1] = xf.y * (f[2 * e + 1] + bv[2 * e + 1]);
                                }
                            } else if (has_vec) {  // bias of the convolution
#pragma unroll
                                for (int e = 0; e < 8; ++e) f[e] += s_beta[n0 + half * 32 + ci * 64 + 8 * c + e];
                            }
                            uint4 ov;
                            __half2 h;
                            h = __floats2half2_rn(f[0], f[1]); ov.x = *reinterpret_cast<uint32_t *>(&h);
                            h = __floats2half2_rn(f[2], f[3]); ov.y = *reinterpret_cast<uint32_t *>(&h);
                            h = __floats2half2_rn(f[4], f[5]); ov.z = *reinterpret_cast<uint32_t *>(&h);
                            h = __floats2half2_rn(f[6], f[7]); ov.w = *reinterpret_cast<uint32_t *>(&h);
                            if (MODE == MODE_STORE_ABS_F16) {
                                sign_word |= ((ov.x & 0x80008000u) >> (4 * c)) | ((ov.y & 0x80008000u) >> (4 * c + 1)) |
                                             ((ov.z & 0x80008000u) >> (4 * c + 2)) | ((ov.w & 0x80008000u) >> (4 * c + 3));
                                ov.x &= 0x7fff7fffu; ov.y &= 0x7fff7fffu; ov.z &= 0x7fff7fffu; ov.w &= 0x7fff7fffu;
                            }
                            *slot = ov;
                        }
                        if (MODE == MODE_STORE_ABS_F16 && valid) p.signs[pix * (p.n_total >> 5) + ((n0 + half * 32 + ci * 64) >> 5)] = sign_word;
                    }
                    fence_proxy_async();
                    asm volatile("bar.sync 2, 256;" ::: "memory");
                    if (issuer) {
                        tma_store_4d(&map_o, blk, n0 + ci * 64, x0, y0, img);
                        tma_store_commit();
                    }
                }
                tcgen05_fence_before();
                mbar_arrive(&acc_empty[as]);  // 256 arrivals release the accumulator stage to the MMA warp
            }
            if (issuer) tma_store_wait_all();
        } else
        for (uint32_t lt = 0;; ++lt) {
            const int tile = sched.next(lt, lane);
            if (tile < 0) break;
            ++trace_tiles;
            const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
            const int sp = tile % tiles_xy, rest = tile / tiles_xy;
            const int n0 = (rest % p.n_tiles) * N_TILE, img = rest / p.n_tiles;
            const int oy = (sp / p.tiles_x) * p.th + ty, ox = (sp % p.tiles_x) * p.tw + tx;
            const bool valid = row < rows && oy < p.h_out && ox < p.w_out;
            const int64_t pix = (static_cast<int64_t>(img) * p.out_h + oy * p.out_stride + p.out_py) * p.out_w + ox * p.out_stride + p.out_px;
            // GDN modes: request this thread's x values BEFORE waiting for the accumulator (their latency hides behind the MMAs)
            constexpr int kMaxChunks = (N_TILE + 63) / 64;
            uint4 xpre[kGdn ? kMaxChunks : 1][4];
            uint32_t spre[kSigned ? kMaxChunks : 1];
            uint4 xlo[MODE == MODE_IGDN_SQ_F16 ? kMaxChunks : 1][4];  // (split activations: x = hi + lo)
            if (kGdn && valid) {
#pragma unroll
                for (int ci = 0; ci < kMaxChunks; ++ci) {
                    const int c0 = half * 32 + ci * 64;
                    if (c0 < N_TILE) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) xpre[ci][c] = __ldg(reinterpret_cast<const uint4 *>(p.gdn_x + pix * p.n_total + n0 + c0) + c);
                        if (kSigned) spre[ci] = __ldg(p.signs + pix * (p.n_total >> 5) + ((n0 + c0) >> 5));
                        if (MODE == MODE_IGDN_SQ_F16 && p.gdn_x_lo) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) xlo[ci][c] = __ldg(reinterpret_cast<const uint4 *>(p.gdn_x_lo + pix * p.n_total + n0 + c0) + c);
                        }
                    }
                }
            }
            mbar_wait(&acc_full[as], aph);
            tcgen05_fence_after();
            const uint32_t taddr = tmem_base + as * N_TILE + (static_cast<uint32_t>(quarter * 32) << 16);
#pragma unroll
            for (int ci = 0; ci < kMaxChunks; ++ci) {
                const int c0 = half * 32 + ci * 64;
                if (c0 >= N_TILE) break;
                uint32_t v[32];
                tmem_ld32(taddr + c0, v);
                if (!valid) continue;
                const int64_t o = pix * p.n_total + n0 + c0;
                if (!kGdn && has_vec) {  // bias of the convolution
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + s_beta[n0 + c0 + e]);
                }
                if (MODE == MODE_NCHW_F32_CLAMP) {
                    float *dst = static_cast<float *>(p.out);
                    const int64_t plane = static_cast<int64_t>(p.out_h) * p.out_w;
                    const int64_t base = static_cast<int64_t>(img) * p.c_real * plane + (pix - static_cast<int64_t>(img) * plane);
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (c0 + e < p.c_real) dst[base + (c0 + e) * plane] = fminf(fmaxf(__uint_as_float(v[e]), 0.0f), 1.0f);
                } else if (kOutF32) {
                    float4 *dst = reinterpret_cast<float4 *>(static_cast<float *>(p.out) + o);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        dst[c] = make_float4(__uint_as_float(v[4 * c]), __uint_as_float(v[4 * c + 1]), __uint_as_float(v[4 * c + 2]),
                                             __uint_as_float(v[4 * c + 3]));
                } else {
                    uint4 *dst = reinterpret_cast<uint4 *>(static_cast<__half *>(p.out) + o);
                    uint32_t sign_word = 0;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(v[8 * c + e]);
                        if (kGdn) {
                            uint4 xv = xpre[kGdn ? ci : 0][c];
                            if (kSigned) {  // x = sign * |x|: half2 word k = 4c + e of the chunk takes (S << k) & 0x80008000
                                const uint32_t sw = spre[kSigned ? ci : 0];
                                xv.x |= (sw << (4 * c)) & 0x80008000u;
                                xv.y |= (sw << (4 * c + 1)) & 0x80008000u;
                                xv.z |= (sw << (4 * c + 2)) & 0x80008000u;
                                xv.w |= (sw << (4 * c + 3)) & 0x80008000u;
                            }
                            const __half2 *xh = reinterpret_cast<const __half2 *>(&xv);
                            const float4 ba = *reinterpret_cast<const float4 *>(s_beta + n0 + c0 + 8 * c);
                            const float4 bb = *reinterpret_cast<const float4 *>(s_beta + n0 + c0 + 8 * c + 4);
                            const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 xf = __half22float2(xh[e]);
                                const float b0 = bv[2 * e], b1 = bv[2 * e + 1];
                                if (MODE == MODE_IGDN_SQ_F16) {  // inverse GDN: x * sqrt(beta + gamma . x^2), the GEMM ran on x^2 / 256
                                    float x0f = xf.x, x1f = xf.y;
                                    if (p.gdn_x_lo) {
                                        const float2 lf = __half22float2(reinterpret_cast<const __half2 *>(&xlo[MODE == MODE_IGDN_SQ_F16 ? ci : 0][c])[e]);
                                        x0f += lf.x;
                                        x1f += lf.y;
                                    }
                                    f[2 * e] = x0f * sqrtf(fmaf(f[2 * e], 256.0f, b0));
                                    f[2 * e + 1] = x1f * sqrtf(fmaf(f[2 * e + 1], 256.0f, b1));
                                } else if (MODE == MODE_IGDN1_F16 || MODE == MODE_IGDN1_ABS_F16) {
                                    f[2 * e] = xf.x * (f[2 * e] + b0);
                                    f[2 * e + 1] = xf.y * (f[2 * e + 1] + b1);
                                } else {
                                    f[2 * e] = xf.x / (f[2 * e] + b0);
                                    f[2 * e + 1] = xf.y / (f[2 * e + 1] + b1);
                                }
                            }
                        }
                        uint4 ov;
                        __half2 h;
                        h = __floats2half2_rn(f[0], f[1]); ov.x = *reinterpret_cast<uint32_t *>(&h);
                        h = __floats2half2_rn(f[2], f[3]); ov.y = *reinterpret_cast<uint32_t *>(&h);
                        h = __floats2half2_rn(f[4], f[5]); ov.z = *reinterpret_cast<uint32_t *>(&h);
                        h = __floats2half2_rn(f[6], f[7]); ov.w = *reinterpret_cast<uint32_t *>(&h);
                        if (MODE == MODE_STORE_ABS_F16) {
                            sign_word |= ((ov.x & 0x80008000u) >> (4 * c)) | ((ov.y & 0x80008000u) >> (4 * c + 1)) |
                                         ((ov.z & 0x80008000u) >> (4 * c + 2)) | ((ov.w & 0x80008000u) >> (4 * c + 3));
                            ov.x &= 0x7fff7fffu; ov.y &= 0x7fff7fffu; ov.z &= 0x7fff7fffu; ov.w &= 0x7fff7fffu;
                        }
                        dst[c] = ov;
                        if ((MODE == MODE_STORE_SQ_F16 && p.out3) || (MODE == MODE_IGDN_SQ_F16 && p.gdn_x_lo)) {  // lo half: value - fp16(value)
                            const __half2 *hh = reinterpret_cast<const __half2 *>(&ov);
                            uint4 lo;
                            uint32_t *lw = reinterpret_cast<uint32_t *>(&lo);
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float2 back = __half22float2(hh[e]);
                                const __half2 l2 = __floats2half2_rn(f[2 * e] - back.x, f[2 * e + 1] - back.y);
                                lw[e] = *reinterpret_cast<const uint32_t *>(&l2);
                            }
                            reinterpret_cast<uint4 *>(static_cast<__half *>(MODE == MODE_STORE_SQ_F16 ? p.out3 : p.out2) + o)[c] = lo;
                        }
                        if (MODE == MODE_STORE_SQ_F16) {  // x^2 / 256, squared in fp32
                            uint4 sq;
                            h = __floats2half2_rn(f[0] * f[0] * (1.0f / 256.0f), f[1] * f[1] * (1.0f / 256.0f)); sq.x = *reinterpret_cast<uint32_t *>(&h);
                            h = __floats2half2_rn(f[2] * f[2] * (1.0f / 256.0f), f[3] * f[3] * (1.0f / 256.0f)); sq.y = *reinterpret_cast<uint32_t *>(&h);
                            h = __floats2half2_rn(f[4] * f[4] * (1.0f / 256.0f), f[5] * f[5] * (1.0f / 256.0f)); sq.z = *reinterpret_cast<uint32_t *>(&h);
                            h = __floats2half2_rn(f[6] * f[6] * (1.0f / 256.0f), f[7] * f[7] * (1.0f / 256.0f)); sq.w = *reinterpret_cast<uint32_t *>(&h);
                            reinterpret_cast<uint4 *>(static_cast<__half *>(p.out2) + o)[c] = sq;
                        }
                    }
                    if (MODE == MODE_STORE_ABS_F16) p.signs[pix * (p.n_total >> 5) + ((n0 + c0) >> 5)] = sign_word;
                }
            }
            tcgen05_fence_before();
            mbar_arrive(&acc_empty[as]);  // 256 arrivals release the accumulator stage to the MMA warp
        }
    } else if (kXform) {
        // =============================== |x| transform warps (10..13, GDN modes 2 / 3) ===============================
        const int row = (warp - 10) * 32 + lane;
        uint32_t it = 0;
        for (uint32_t lt = 0; sched.next(lt, lane) >= 0; ++lt) {
            for (int k_it = 0; k_it < k_iters; ++k_it, ++it) {
                const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
                mbar_wait(&full[s], ph);
                if (row < rows) {
                    uint4 *r = reinterpret_cast<uint4 *>(smem + s * L::kStageBytes + row * 128);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        // rotate the 16-byte chunk with the row: 8 neighbouring rows hit 8 different bank groups
                        const int cc = (c + row) & 7;
                        uint4 v = r[cc];
                        v.x &= 0x7fff7fffu; v.y &= 0x7fff7fffu; v.z &= 0x7fff7fffu; v.w &= 0x7fff7fffu;
                        r[cc] = v;
                    }
                }
                fence_proxy_async();
                mbar_arrive(&xform[s]);
            }
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
    if (threadIdx.x == 64) trace_emit(p.trace, TRACE_CONV_TC, trace_t0, trace_tiles);
}

template <int N_TILE, int STAGES, int MODE, bool STAGED = false>
static int launch(const CUtensorMap &ma, const CUtensorMap &ma2, const CUtensorMap &mb, const Params &p, cudaStream_t st,
                  const CUtensorMap *mo = nullptr, const CUtensorMap *mx = nullptr) {
    using L = Smem<N_TILE, STAGES, STAGED>;
    constexpr bool kXform = MODE == MODE_IGDN1_F16 || MODE == MODE_GDN1_F16;
    static_assert(L::kTotal + 1024 <= 227 * 1024, "shared memory budget");
    const int smem = uniform_smem(L::kTotal + 1024);  // + slack for the manual 1024-byte alignment
    static std::atomic<uint64_t> configured{0};  // per device ordinal
    if (int rc = ensure_dyn_smem(tc_conv_kernel<N_TILE, STAGES, MODE, STAGED>, smem, configured)) return rc;
    const int total = p.tiles_x * p.tiles_y * p.n_tiles * p.batch;
    const int grid = total < persistent_grid() ? total : persistent_grid();
    tc_conv_kernel<N_TILE, STAGES, MODE, STAGED><<<grid, kXform ? 448 : 320, smem, st>>>(ma, ma2, mb, mo ? *mo : ma, mx ? *mx : ma, p);
    SC2_LAUNCH_CHECK("tc_conv_kernel");
    return SC2_OK;
}

// experiments: SC2_TC_UNSTAGED=1 keeps the per-thread global loads / stores of the first version
bool unstaged_epilogue() {
    static const bool v = [] { const char *e = std::getenv("SC2_TC_UNSTAGED"); return e && e[0] == '1'; }();
    return v;
}

// NCHW fp32 -> NHWC fp16 with the channel dimension zero-padded to c_pad (feeds the first tensor-core layer).
__global__ void nchw_f32_to_nhwc_f16_kernel(const float *__restrict__ x, __half *__restrict__ y, int c, int64_t hw, int c_pad,
                                            int64_t total) {
    // one thread per (pixel, channel pair); consecutive threads walk the padded channel dimension -> coalesced writes
    const int pairs = c_pad / 2;
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int cp = static_cast<int>(i % pairs);
        const int64_t pix = i / pairs;          // b * hw + s
        const int64_t b = pix / hw, sp = pix - b * hw;
        const int c0 = 2 * cp;
        const float v0 = c0 < c ? __ldg(x + (b * c + c0) * hw + sp) : 0.0f;
        const float v1 = c0 + 1 < c ? __ldg(x + (b * c + c0 + 1) * hw + sp) : 0.0f;
        reinterpret_cast<__half2 *>(y)[i] = __floats2half2_rn(v0, v1);
    }
}

// The same for c_pad == 64 (the bottleneck's latent: 24 channels in front of the first g_s layer) as a tiled transpose: coalesced
// 128-byte reads along the pixels of one channel, 8 channels packed per thread into a 16-byte unit, units staged in shared memory
// (swizzled like the tensor-core tiles: conflict-free) and written back as whole 128-byte pixel rows.  The one-thread-per-channel-pair
// version above read 12 different cache lines per warp load: 0.135 ms for 130 MB (13 % of the copy bandwidth).
__global__ void __launch_bounds__(256) nchw_f32_to_nhwc64_f16_kernel(const float *__restrict__ x, __half *__restrict__ y, int c, int hw) {
    __shared__ __align__(128) uint4 tile[64 * 8];
    const int b = blockIdx.y, p0 = blockIdx.x * 64;
    const int px = threadIdx.x & 63, cg0 = threadIdx.x >> 6;
    const float *xb = x + static_cast<int64_t>(b) * c * hw;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int cg = cg0 + 4 * r;
        float f[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int ch = cg * 8 + e;
            f[e] = (ch < c && p0 + px < hw) ? __ldg(xb + static_cast<int64_t>(ch) * hw + p0 + px) : 0.0f;
        }
        uint4 v;
        __half2 h;
        h = __floats2half2_rn(f[0], f[1]); v.x = *reinterpret_cast<uint32_t *>(&h);
        h = __floats2half2_rn(f[2], f[3]); v.y = *reinterpret_cast<uint32_t *>(&h);
        h = __floats2half2_rn(f[4], f[5]); v.z = *reinterpret_cast<uint32_t *>(&h);
        h = __floats2half2_rn(f[6], f[7]); v.w = *reinterpret_cast<uint32_t *>(&h);
        tile[px * 8 + (cg ^ (px & 7))] = v;
    }
    __syncthreads();
    uint4 *yb = reinterpret_cast<uint4 *>(y) + (static_cast<int64_t>(b) * hw + p0) * 8;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int i = threadIdx.x + 256 * r;  // unit index inside the 64-pixel tile: pixel i / 8, unit i % 8
        const int q = i >> 3, u = i & 7;
        if (p0 + q < hw) yb[i] = tile[q * 8 + (u ^ (q & 7))];
    }
}

}  // namespace tc
}  // namespace sc2

extern "C" {

int sc2_nchw_f32_to_nhwc_f16(const float *x, void *y, int batch, int channels, int64_t spatial, int c_pad, sc2_stream_t stream) {
    if (!x || !y || batch < 0 || channels < 1 || c_pad < channels || (c_pad & 1) || spatial < 0) return SC2_ERR_INVALID_ARG;
    const int64_t total = static_cast<int64_t>(batch) * spatial * (c_pad / 2);
    if (total == 0) return SC2_OK;
    if (c_pad == 64 && spatial <= 0x7fffffff - 64 && batch <= 65535) {
        const dim3 grid(static_cast<unsigned>((spatial + 63) / 64), static_cast<unsigned>(batch));
        sc2::tc::nchw_f32_to_nhwc64_f16_kernel<<<grid, 256, 0, sc2::as_stream(stream)>>>(x, static_cast<__half *>(y), channels,
                                                                                         static_cast<int>(spatial));
        SC2_LAUNCH_CHECK("nchw_f32_to_nhwc64_f16_kernel");
        return SC2_OK;
    }
    int64_t blocks = (total + 255) / 256;
    if (blocks > sc2::kNumSMs * 32) blocks = sc2::kNumSMs * 32;
    sc2::tc::nchw_f32_to_nhwc_f16_kernel<<<static_cast<int>(blocks), 256, 0, sc2::as_stream(stream)>>>(
        x, static_cast<__half *>(y), channels, spatial, c_pad, total);
    SC2_LAUNCH_CHECK("nchw_f32_to_nhwc_f16_kernel");
    return SC2_OK;
}

int sc2_tc_conv_ex(const sc2_tc_conv_ex_desc *d, const void *x, const void *x_lo, const void *w_packed, const float *vec, const void *gdn_x,
                   const void *gdn_x_lo, void *out, void *out2, void *out3, uint32_t *signs, int32_t *tile_counter, sc2_stream_t stream) {
    using namespace sc2::tc;
    if (!d || !x || !w_packed || !out) return SC2_ERR_INVALID_ARG;
    if (d->batch < 1 || d->c_in_pad % kBlockK || d->c_in_pad < kBlockK) return SC2_ERR_INVALID_ARG;
    if (d->mode < 0 || d->mode > MODE_NCHW_F32_CLAMP) return SC2_ERR_INVALID_ARG;
    if (d->kh < 1 || d->kw < 1 || d->h_out < 1 || d->w_out < 1 || d->out_stride < 1) return SC2_ERR_INVALID_ARG;
    if (d->out_py < 0 || d->out_px < 0 || d->out_py >= d->out_stride || d->out_px >= d->out_stride) return SC2_ERR_INVALID_ARG;
    if ((d->h_out - 1) * d->out_stride + d->out_py >= d->out_h || (d->w_out - 1) * d->out_stride + d->out_px >= d->out_w) return SC2_ERR_INVALID_ARG;
    const bool gdn = d->mode == MODE_IGDN1_F16 || d->mode == MODE_GDN1_F16 || d->mode == MODE_IGDN1_ABS_F16 || d->mode == MODE_IGDN_SQ_F16;
    if ((d->mode == MODE_STORE_ABS_F16 || d->mode == MODE_IGDN1_ABS_F16) && (!signs || d->c_out % 32)) return SC2_ERR_INVALID_ARG;
    if (d->mode == MODE_STORE_SQ_F16 && !out2) return SC2_ERR_INVALID_ARG;
    if (gdn && (!vec || !gdn_x || d->kh != 1 || d->kw != 1 || d->pad_x != 0 || d->pad_y != 0 || d->c_in_pad != d->c_out || d->out_stride != 1))
        return SC2_ERR_INVALID_ARG;
    if (d->c_out > kMaxBeta && vec) return SC2_ERR_UNSUPPORTED;
    const int h_out = d->h_out, w_out = d->w_out;
    // tile shape: tw columns x th rows, tw * th <= 128
    int n_col_tiles = (w_out + 127) / 128;
    int tw = (w_out + n_col_tiles - 1) / n_col_tiles;
    tw = (tw + 7) / 8 * 8;
    if (tw > 128) tw = 128;
    int th = 128 / tw;
    if (th > h_out) th = h_out;
    if (th > 256) th = 256;
    int n_tile, n_rows = d->c_out;  // n_rows: rows per tap of the packed weights
    if (d->mode == MODE_NCHW_F32_CLAMP) {
        if (d->c_out > 32) return SC2_ERR_UNSUPPORTED;
        n_tile = 32; n_rows = 32;  // (the pack is zero-padded to 32 rows per tap)
    } else if (d->c_out % 256 == 0) n_tile = 256;
    else if (d->c_out % 192 == 0 && (d->mode == MODE_STORE_SQ_F16 || d->mode == MODE_IGDN_SQ_F16 || d->mode == MODE_STORE_F16)) n_tile = 192;
    else if (d->c_out % 128 == 0) n_tile = 128;
    else if (d->c_out % 64 == 0) n_tile = 64;
    else return SC2_ERR_UNSUPPORTED;
    Params p;
    p.tw = tw; p.th = th;
    p.tiles_x = (w_out + tw - 1) / tw;
    p.tiles_y = (h_out + th - 1) / th;
    p.taps_x = d->kw; p.taps_y = d->kh; p.pad_x = d->pad_x; p.pad_y = d->pad_y;
    p.k_chunks = d->c_in_pad / kBlockK;
    p.n_total = n_rows;
    p.n_tiles = n_rows / n_tile;
    p.batch = d->batch;
    p.h_out = h_out; p.w_out = w_out;
    p.out_h = d->out_h; p.out_w = d->out_w; p.out_stride = d->out_stride; p.out_py = d->out_py; p.out_px = d->out_px;
    p.c_real = d->c_out;
    p.out2 = out2;
    p.out3 = out3;
    p.gdn_x_lo = static_cast<const __half *>(gdn_x_lo);
    p.a_passes = x_lo ? 2 : 1;
    p.beta = vec;
    p.gdn_x = static_cast<const __half *>(gdn_x);
    p.signs = signs;
    p.out = out;
    p.tile_counter = tile_counter;
    p.trace = sc2::trace_sink();
    if (static_cast<int64_t>(p.tiles_x) * p.tiles_y * p.n_tiles * p.batch > 0x7fffffff - 1024) return SC2_ERR_UNSUPPORTED;
    if (d->mode == MODE_IGDN_SQ_F16 && gdn_x_lo && !out2) return SC2_ERR_INVALID_ARG;
    CUtensorMap ma, ma2, mb;
    int rc = make_nhwc_map(&ma, x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->c_in_pad, d->w_in, d->h_in, d->batch, kBlockK, tw, th);
    if (rc) return rc;
    ma2 = ma;
    if (x_lo) {
        rc = make_nhwc_map(&ma2, x_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d->c_in_pad, d->w_in, d->h_in, d->batch, kBlockK, tw, th);
        if (rc) return rc;
    }
    rc = make_weight_map(&mb, w_packed, d->c_in_pad, d->kh * d->kw * n_rows, n_tile);
    if (rc) return rc;
    cudaStream_t st = sc2::as_stream(stream);
#define SC2_TC_DISPATCH(NT, STG)                                                   \
    switch (d->mode) {                                                             \
        case MODE_STORE_F16: return launch<NT, STG, MODE_STORE_F16>(ma, ma2, mb, p, st);  \
        case MODE_STORE_F32: return launch<NT, STG, MODE_STORE_F32>(ma, ma2, mb, p, st);  \
        case MODE_IGDN1_F16: return launch<NT, STG, MODE_IGDN1_F16>(ma, ma2, mb, p, st);  \
        case MODE_STORE_ABS_F16: return launch<NT, STG, MODE_STORE_ABS_F16>(ma, ma2, mb, p, st);  \
        case MODE_IGDN1_ABS_F16: return launch<NT, STG, MODE_IGDN1_ABS_F16>(ma, ma2, mb, p, st);  \
        case MODE_GDN1_F16: return launch<NT, STG, MODE_GDN1_F16>(ma, ma2, mb, p, st);    \
        default: return SC2_ERR_UNSUPPORTED;                                       \
    }
    if (n_tile == 32) return launch<32, 8, MODE_NCHW_F32_CLAMP>(ma, ma2, mb, p, st);
    if (n_tile == 192) {
        switch (d->mode) {
            case MODE_STORE_SQ_F16: return launch<192, 4, MODE_STORE_SQ_F16>(ma, ma2, mb, p, st);
            case MODE_IGDN_SQ_F16: return launch<192, 4, MODE_IGDN_SQ_F16>(ma, ma2, mb, p, st);
            default: return launch<192, 4, MODE_STORE_F16>(ma, ma2, mb, p, st);
        }
    }
    if (d->mode == MODE_STORE_SQ_F16 || d->mode == MODE_IGDN_SQ_F16) {  // (other channel counts: 64-wide tiles)
        if (d->c_out % 64) return SC2_ERR_UNSUPPORTED;
        p.n_tiles = n_rows / 64;
        CUtensorMap mb64;
        rc = make_weight_map(&mb64, w_packed, d->c_in_pad, d->kh * d->kw * n_rows, 64);
        if (rc) return rc;
        return d->mode == MODE_STORE_SQ_F16 ? launch<64, 8, MODE_STORE_SQ_F16>(ma, ma2, mb64, p, st) : launch<64, 8, MODE_IGDN_SQ_F16>(ma, ma2, mb64, p, st);
    }
    if (n_tile == 256 && d->out_stride == 1 && !x_lo &&
        (d->mode == MODE_STORE_F16 || d->mode == MODE_STORE_ABS_F16 || d->mode == MODE_IGDN1_ABS_F16) && !sc2::tc::unstaged_epilogue()) {
        // fp16 outputs through shared-memory staging blocks and TMA stores (3 ring stages leave room for them)
        CUtensorMap mo, mx;
        rc = make_nhwc_map(&mo, out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, n_rows, w_out, h_out, d->batch, 64, tw, th);
        if (rc) return rc;
        mx = mo;
        if (d->mode == MODE_IGDN1_ABS_F16) {
            rc = make_nhwc_map(&mx, gdn_x, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, n_rows, w_out, h_out, d->batch, 64, tw, th);
            if (rc) return rc;
            return launch<256, 3, MODE_IGDN1_ABS_F16, true>(ma, ma2, mb, p, st, &mo, &mx);
        }
        if (d->mode == MODE_STORE_ABS_F16) return launch<256, 3, MODE_STORE_ABS_F16, true>(ma, ma2, mb, p, st, &mo, &mx);
        return launch<256, 3, MODE_STORE_F16, true>(ma, ma2, mb, p, st, &mo, &mx);
    }
    if (n_tile == 256 && d->out_stride == 1 && !x_lo && d->mode == MODE_STORE_F32 && !sc2::tc::unstaged_epilogue())
        return launch<256, 3, MODE_STORE_F32, true>(ma, ma2, mb, p, st);
    if (n_tile == 256) { SC2_TC_DISPATCH(256, 4) }
    if (n_tile == 128) { SC2_TC_DISPATCH(128, 6) }
    SC2_TC_DISPATCH(64, 8)
#undef SC2_TC_DISPATCH
}

int sc2_tc_conv_nhwc(const sc2_tc_conv_desc *d, const void *x, const void *w_packed, const float *beta, const void *gdn_x,
                     void *out, uint32_t *signs, int32_t *tile_counter, sc2_stream_t stream) {
    if (!d) return SC2_ERR_INVALID_ARG;
    if (d->mode < 0 || d->mode > 5) return SC2_ERR_INVALID_ARG;
    sc2_tc_conv_ex_desc e;
    e.batch = d->batch; e.h_in = d->h_in; e.w_in = d->w_in; e.c_in_pad = d->c_in_pad;
    e.c_out = d->c_out; e.kh = d->kh; e.kw = d->kw; e.pad_y = d->pad; e.pad_x = d->pad;
    e.mode = d->mode;
    e.h_out = d->h_in + 2 * d->pad - d->kh + 1; e.w_out = d->w_in + 2 * d->pad - d->kw + 1;
    if (e.h_out < 1 || e.w_out < 1) return SC2_ERR_INVALID_ARG;
    e.out_h = e.h_out; e.out_w = e.w_out; e.out_stride = 1; e.out_py = 0; e.out_px = 0;
    return sc2_tc_conv_ex(&e, x, nullptr, w_packed, beta, gdn_x, nullptr, out, nullptr, nullptr, signs, tile_counter, stream);
}

}  // extern "C"
