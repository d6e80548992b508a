// conv_tc_first.cu -- first layer of the analysis transform g_a on the tensor cores, im2col fused into the kernel.
//
// Conv2d(3 -> C, k5, s2, p2) on an fp32 NCHW image (sc2bench/models/layer.py:476-477) has K = 75: too few input channels for
// the shifted-TMA-box scheme of conv_tc_split.cu.  Instead four PRODUCER WARPS gather each output pixel's 5x5x3 patch straight
// from the image (row segments as 16-byte / 8-byte vector loads, neighbouring threads = neighbouring pixels -> coalesced),
// split every value into the (hi, lo) fp16 pair of the fp32-grade scheme (value = hi + lo / 2048) and write the A tile into
// shared memory in the UMMA canonical K-major SWIZZLE_128B layout.  The weights (hi, lo) stay resident in shared memory for
// the life of the persistent CTA.  Everything downstream is the split kernel's machinery: three tcgen05.mma passes per K step
// into two TMEM accumulators (double buffered across tiles), epilogue warps add D0 + D1 / 2048, split the result again and
// TMA-store it as PARITY PLANES [B * 4, H/4, W/4, C] (plane py*2+px holds output pixel (2Y+py, 2X+px)), which is what the
// GDN1 and the stride-2 second convolution read next.  No im2col tensor ever exists in HBM.
#include "tc_common.cuh"

namespace sc2 {
namespace tcf {

using namespace sc2::tc;

constexpr float kLoScale = 2048.0f;
constexpr float kLoInv = 1.0f / 2048.0f;
constexpr int kThreads = 448;  // warp 0 TMA (weights), warp 1 MMA, warps 2..9 epilogue, warps 10..13 im2col producers
constexpr int kAStages = 2;

struct Params {
    const float *image;  // [batch, c_in, h_in, w_in] fp32
    int batch, c_in, h_in, w_in, kh, kw, pad;
    int hp, wp;          // parity-plane geometry = output size / 2
    int K;               // c_in * kh * kw (<= 128)
    int k_steps;         // ceil(K / 16)
    int c_out, out_c;    // valid channels, channel pitch of the output planes
    int stage_c;         // channel pitch of the staging tile (out_c padded to an odd number of 16-byte units: no bank conflicts)
    int tiles_x, tiles_y, tw, th;
    int *tile_counter;   // zeroed by the caller: dynamic tile schedule; nullptr: static
    TraceSink trace;     // diagnostics (common.cuh)
};

template <int N_TILE>
struct Smem {
    static constexpr int kBBytes = N_TILE * 128;
    static constexpr int kResBytes = 4 * kBBytes;          // [chunk 0 hi][chunk 0 lo][chunk 1 hi][chunk 1 lo]
    static constexpr int kStageBytes = 4 * kABytes;        // same order for A
    static constexpr int kRingBytes = kAStages * kStageBytes;
    // staging rows padded by 16 bytes against bank conflicts (see conv_tc_split.cu) where the budget allows: not for N_TILE = 96
    static constexpr int kPadC = (kResBytes + kRingBytes + 2 * kTileM * (N_TILE + 8) * 2 + 2048 <= 227 * 1024) ? 8 : 0;
    static constexpr int kStagePlane = kTileM * (N_TILE + kPadC) * 2;
    static constexpr int kStagingOffset = kResBytes + kRingBytes;
    static constexpr int kBarOffset = kStagingOffset + 2 * kStagePlane;
    static constexpr int kSchedOffset = kBarOffset + (2 * kAStages + 5) * 8 + 16;
    static constexpr int kTotal = kSchedOffset + kTileSchedBytes;
    static_assert(kTotal + 1024 <= 227 * 1024, "shared memory budget");
};

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

// One row segment of KW = 5 input floats starting at column ix0 (may stick out of the image).  Interior segments use vector
// loads: ix0 = 4X + 2px - 2 is 16-byte aligned for px = 1 and 8-byte aligned for px = 0, and neighbouring threads read
// neighbouring 16-byte pieces, so a warp's loads are contiguous.
__device__ __forceinline__ void load_row5(const float *row, int ix0, int w_in, bool row_ok, bool vec_ok, float *out) {
    if (!row_ok) {
#pragma unroll
        for (int i = 0; i < 5; ++i) out[i] = 0.0f;
    } else if (vec_ok && ix0 >= 0 && ix0 + 4 < w_in) {
        if ((ix0 & 3) == 0) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(row + ix0));
            out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w;
        } else {
            const float2 a = __ldg(reinterpret_cast<const float2 *>(row + ix0));
            const float2 b = __ldg(reinterpret_cast<const float2 *>(row + ix0 + 2));
            out[0] = a.x; out[1] = a.y; out[2] = b.x; out[3] = b.y;
        }
        out[4] = __ldg(row + ix0 + 4);
    } else {
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int ix = ix0 + i;
            out[i] = (ix >= 0 && ix < w_in) ? __ldg(row + ix) : 0.0f;
        }
    }
}

template <int N_TILE>
__global__ void __launch_bounds__(kThreads, 1)
tc_first_layer_kernel(const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                      const __grid_constant__ CUtensorMap map_o_hi, const __grid_constant__ CUtensorMap map_o_lo,
                      const __grid_constant__ Params p) {
    using L = Smem<N_TILE>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem_res = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint8_t *ring = smem_res + L::kResBytes;
    uint8_t *staging = smem_res + L::kStagingOffset;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_res + L::kBarOffset);
    uint64_t *empty = full + kAStages;
    uint64_t *acc_full = empty + kAStages;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *b_full = acc_empty + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(b_full + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rows = p.tw * p.th;
    const int tiles_xy = p.tiles_x * p.tiles_y;
    const int total_tiles = tiles_xy * p.batch * 4;
    const unsigned long long trace_t0 = p.trace.buf ? trace_now() : 0ull;
    int trace_tiles = 0;
    TileSched sched;
    sched.bind(smem_res + L::kSchedOffset, p.tile_counter, total_tiles);

    if (threadIdx.x == 0) {
        sched.init(13);  // consumers: MMA warp, 8 epilogue warps, 4 im2col producer warps
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        tma_prefetch_desc(&map_o_hi);
        tma_prefetch_desc(&map_o_lo);
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(&full[s], 128);  // one arrival per producer thread
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 256);
        }
        mbar_init(b_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // =============================== weights: once per CTA ===============================
        if (elect_one()) {
            mbar_expect_tx(b_full, static_cast<uint32_t>(4 * L::kBBytes));
            tma_load_2d(&map_b_hi, b_full, smem_res + 0 * L::kBBytes, 0, 0);
            tma_load_2d(&map_b_lo, b_full, smem_res + 1 * L::kBBytes, 0, 0);
            tma_load_2d(&map_b_hi, b_full, smem_res + 2 * L::kBBytes, kBlockK, 0);
            tma_load_2d(&map_b_lo, b_full, smem_res + 3 * L::kBBytes, kBlockK, 0);
            // ... and the tile scheduler: claim tiles and hand them to the other warps
            for (uint32_t qn = 0;; ++qn) {
                const int tile = sched.claim(qn);
                sched.publish(qn, tile);
                if (tile < 0 || p.tile_counter == nullptr) break;
            }
        }
    } else if (warp == 1) {
        // =============================== MMA issuer ===============================
        constexpr uint32_t idesc = make_idesc(N_TILE);
        mbar_wait(b_full, 0);
        for (uint32_t lt = 0; sched.next(lt, lane) >= 0; ++lt) {
            const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
            const uint32_t s = lt % kAStages, ph = (lt / kAStages) & 1u;
            mbar_wait(&acc_empty[as], aph ^ 1u);
            mbar_wait(&full[s], ph);
            tcgen05_fence_after();
            if (elect_one()) {
                const uint32_t a_base = smem_u32(ring + s * L::kStageBytes), b_base = smem_u32(smem_res);
                const uint32_t d0 = tmem_base + as * 256u, d1 = d0 + 128u;
                for (int k = 0; k < p.k_steps; ++k) {
                    const int kc = k >> 2, kk = k & 3;
                    const uint64_t a_hi = make_smem_desc(a_base + (2 * kc) * kABytes) + 2 * kk;
                    const uint64_t a_lo = make_smem_desc(a_base + (2 * kc + 1) * kABytes) + 2 * kk;
                    const uint64_t b_hi = make_smem_desc(b_base + (2 * kc) * L::kBBytes) + 2 * kk;
                    const uint64_t b_lo = make_smem_desc(b_base + (2 * kc + 1) * L::kBBytes) + 2 * kk;
                    umma_f16(d0, a_hi, b_hi, idesc, k > 0 ? 1u : 0u);  // D0 += hi * hi
                    umma_f16(d1, a_hi, b_lo, idesc, k > 0 ? 1u : 0u);  // D1 += hi * lo
                    umma_f16(d1, a_lo, b_hi, idesc, 1u);               // D1 += lo * hi
                }
                umma_commit(&empty[s]);
                umma_commit(&acc_full[as]);
            }
            __syncwarp();
        }
    } else if (warp < 10) {
        // =============================== epilogue warps (2..9) ===============================
        const int half = (warp - 2) >> 2;
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const bool issuer = threadIdx.x == 64;
        __half *st_hi = reinterpret_cast<__half *>(staging), *st_lo = reinterpret_cast<__half *>(staging + L::kStagePlane);
        for (uint32_t lt = 0;; ++lt) {
            const int tile = sched.next(lt, lane);
            if (tile < 0) break;
            ++trace_tiles;
            const uint32_t as = lt & 1u, aph = (lt >> 1) & 1u;
            const int sp = tile % tiles_xy, img = tile / tiles_xy;  // img = b * 4 + parity
            const int x0 = (sp % p.tiles_x) * p.tw, y0 = (sp / p.tiles_x) * p.th;
            if (issuer) tma_store_wait_read();  // the previous tile's bulk stores have read the staging buffer
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&acc_full[as], aph);
            tcgen05_fence_after();
            const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + as * 256u;
#pragma unroll 1
            for (int c0 = half * 32; c0 < N_TILE; c0 += 64) {
                uint32_t d0[32], d1[32];
                tmem_ld32(lane_addr + c0, d0);
                tmem_ld32(lane_addr + 128 + c0, d1);
                if (row < rows) {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int c = c0 + 8 * g;
                        if (c >= p.out_c) continue;
                        float f[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e)
                            f[e] = (c + e < p.c_out) ? __uint_as_float(d0[8 * g + e]) + __uint_as_float(d1[8 * g + e]) * kLoInv : 0.0f;
                        uint4 h, l;
                        split8(f, h, l);
                        *reinterpret_cast<uint4 *>(st_hi + row * p.stage_c + c) = h;
                        *reinterpret_cast<uint4 *>(st_lo + row * p.stage_c + c) = l;
                    }
                }
            }
            tcgen05_fence_before();
            mbar_arrive(&acc_empty[as]);
            fence_proxy_async();
            asm volatile("bar.sync 2, 256;" ::: "memory");
            if (issuer) {
                tma_store_4d(&map_o_hi, st_hi, 0, x0, y0, img);
                tma_store_4d(&map_o_lo, st_lo, 0, x0, y0, img);
                tma_store_commit();
            }
        }
        if (issuer) tma_store_wait_all();
    } else {
        // =============================== im2col producers (warps 10..13): thread = tile row = output pixel ===============================
        const int row = (warp - 10) * 32 + lane;
        const int ty = row / p.tw, tx = row - ty * p.tw;
        const int KK = p.kh * p.kw;
        for (uint32_t lt = 0;; ++lt) {
            const int tile = sched.next(lt, lane);
            if (tile < 0) break;
            const uint32_t s = lt % kAStages, ph = (lt / kAStages) & 1u;
            const int sp = tile % tiles_xy, img = tile / tiles_xy;
            const int b = img >> 2, py = (img >> 1) & 1, px = img & 1;
            const int Y = (sp / p.tiles_x) * p.th + ty, X = (sp % p.tiles_x) * p.tw + tx;
            const bool valid = row < rows && Y < p.hp && X < p.wp;
            // gather the patch (K order = (c, dy, dx), like weight.reshape(c_out, -1)); 8 values at a time -> one 16-byte
            // chunk of the hi tile and one of the lo tile
            const int iy0 = 2 * (2 * Y + py) - p.pad, ix0 = 2 * (2 * X + px) - p.pad;
            const float *img_base = p.image + static_cast<int64_t>(b) * p.c_in * p.h_in * p.w_in;
            mbar_wait(&empty[s], ph ^ 1u);
            uint8_t *stage = ring + s * L::kStageBytes;
            if (row < rows && p.c_in == 3 && p.kh == 5 && p.kw == 5) {
                // specialised 3 x 5 x 5 patch: fully unrolled, vector row loads, K index known at compile time
                float f[80];
                const bool vec_ok = (p.w_in & 3) == 0 && ((ix0 & 1) == 0);
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int dy = 0; dy < 5; ++dy) {
                        const int iy = iy0 + dy;
                        load_row5(img_base + (static_cast<int64_t>(c) * p.h_in + iy) * p.w_in, ix0, p.w_in,
                                  valid && iy >= 0 && iy < p.h_in, vec_ok, &f[(c * 5 + dy) * 5]);
                    }
#pragma unroll
                for (int k = 75; k < 80; ++k) f[k] = 0.0f;
#pragma unroll
                for (int g = 0; g < 10; ++g) {
                    uint4 h, l;
                    split8(&f[g * 8], h, l);
                    const int kc = g >> 3, j = g & 7;
                    const int phys = j ^ (row & 7);
                    *reinterpret_cast<uint4 *>(stage + (2 * kc) * kABytes + row * 128 + phys * 16) = h;
                    *reinterpret_cast<uint4 *>(stage + (2 * kc + 1) * kABytes + row * 128 + phys * 16) = l;
                }
            } else if (row < rows) {
                const int n_groups = p.k_steps * 2;  // 16-byte groups of 8 K values
#pragma unroll 1
                for (int g = 0; g < n_groups; ++g) {
                    float f[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int k = g * 8 + e;
                        float v = 0.0f;
                        if (valid && k < p.K) {
                            const int c = k / KK, r = k - c * KK;
                            const int dy = r / p.kw, dx = r - dy * p.kw;
                            const int iy = iy0 + dy, ix = ix0 + dx;
                            if (iy >= 0 && iy < p.h_in && ix >= 0 && ix < p.w_in) v = __ldg(img_base + (static_cast<int64_t>(c) * p.h_in + iy) * p.w_in + ix);
                        }
                        f[e] = v;
                    }
                    uint4 h, l;
                    split8(f, h, l);
                    const int kc = g >> 3, j = g & 7;
                    const int phys = j ^ (row & 7);  // SWIZZLE_128B: 16-byte chunk index XOR (row mod 8)
                    *reinterpret_cast<uint4 *>(stage + (2 * kc) * kABytes + row * 128 + phys * 16) = h;
                    *reinterpret_cast<uint4 *>(stage + (2 * kc + 1) * kABytes + row * 128 + phys * 16) = l;
                }
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
        }
    }
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
    if (threadIdx.x == 64) trace_emit(p.trace, TRACE_CONV_FIRST, trace_t0, trace_tiles);
}

template <int N_TILE>
static int launch(const CUtensorMap *maps, const Params &p, cudaStream_t st) {
    using L = Smem<N_TILE>;
    const int smem = uniform_smem(L::kTotal + 1024);  // + slack for the manual 1024-byte alignment
    static std::atomic<uint64_t> configured{0};  // per device ordinal
    if (int rc = ensure_dyn_smem(tc_first_layer_kernel<N_TILE>, smem, configured)) return rc;
    const int64_t total = static_cast<int64_t>(p.tiles_x) * p.tiles_y * p.batch * 4;
    if (total > 0x7fffffff) return SC2_ERR_UNSUPPORTED;
    const int grid = total < persistent_grid() ? static_cast<int>(total) : persistent_grid();
    tc_first_layer_kernel<N_TILE><<<grid, kThreads, smem, st>>>(maps[0], maps[1], maps[2], maps[3], p);
    SC2_LAUNCH_CHECK("tc_first_layer_kernel");
    return SC2_OK;
}

}  // namespace tcf
}  // namespace sc2

extern "C" {

int sc2_tc_first_layer(const float *image, int batch, int c_in, int h_in, int w_in, int kh, int kw, int pad, int c_out,
                       const void *w_hi, const void *w_lo, void *out_hi, void *out_lo, int out_c, int32_t *tile_counter,
                       sc2_stream_t stream) {
    using namespace sc2::tcf;
    if (!image || !w_hi || !w_lo || !out_hi || !out_lo || batch < 1) return SC2_ERR_INVALID_ARG;
    const int K = c_in * kh * kw;
    if (K > 128 || K < 1) return SC2_ERR_UNSUPPORTED;
    const int h_out = (h_in + 2 * pad - kh) / 2 + 1, w_out = (w_in + 2 * pad - kw) / 2 + 1;
    if (h_out < 2 || w_out < 2 || (h_out & 1) || (w_out & 1)) return SC2_ERR_UNSUPPORTED;
    int n_tile;
    if (c_out <= 32) n_tile = 32;
    else if (c_out <= 48) n_tile = 48;
    else if (c_out <= 64) n_tile = 64;
    else if (c_out <= 96) n_tile = 96;
    else return SC2_ERR_UNSUPPORTED;
    if (out_c % 8 || out_c < c_out || out_c > n_tile) return SC2_ERR_INVALID_ARG;
    Params p;
    p.image = image; p.batch = batch; p.c_in = c_in; p.h_in = h_in; p.w_in = w_in; p.kh = kh; p.kw = kw; p.pad = pad;
    p.hp = h_out / 2; p.wp = w_out / 2;
    p.K = K; p.k_steps = (K + 15) / 16;
    p.c_out = c_out; p.out_c = out_c;
    {   // staging pitch: padded to an odd number of 16-byte units where the kernel for this n_tile has the room
        const int pad_c = n_tile == 32 ? Smem<32>::kPadC : n_tile == 48 ? Smem<48>::kPadC : n_tile == 64 ? Smem<64>::kPadC : Smem<96>::kPadC;
        p.stage_c = (pad_c && (out_c / 8) % 2 == 0) ? out_c + 8 : out_c;
    }
    p.tile_counter = tile_counter;
    p.trace = sc2::trace_sink();
    int n_col_tiles = (p.wp + 127) / 128;
    int tw = (p.wp + n_col_tiles - 1) / n_col_tiles;
    tw = (tw + 7) / 8 * 8;
    if (tw > 128) tw = 128;
    int th = 128 / tw;
    if (th > p.hp) th = p.hp;
    p.tw = tw; p.th = th;
    p.tiles_x = (p.wp + tw - 1) / tw;
    p.tiles_y = (p.hp + th - 1) / th;
    CUtensorMap maps[4];
    // weights packed as ONE tap: [n_tile, k_pad] with k_pad = k_steps * 16 (as pack_conv_weight_split(as_patches=True) makes them)
    const int k_pad = p.k_steps * 16;
    int rc = make_weight_map(&maps[0], w_hi, k_pad, n_tile, n_tile);
    if (rc) return rc;
    rc = make_weight_map(&maps[1], w_lo, k_pad, n_tile, n_tile);
    if (rc) return rc;
    rc = make_nhwc_map(&maps[2], out_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out_c, p.wp, p.hp, batch * 4, p.stage_c, tw, th, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    rc = make_nhwc_map(&maps[3], out_lo, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, out_c, p.wp, p.hp, batch * 4, p.stage_c, tw, th, CU_TENSOR_MAP_SWIZZLE_NONE);
    if (rc) return rc;
    cudaStream_t st = sc2::as_stream(stream);
    switch (n_tile) {
        case 32: return launch<32>(maps, p, st);
        case 48: return launch<48>(maps, p, st);
        case 64: return launch<64>(maps, p, st);
        default: return launch<96>(maps, p, st);
    }
}

}  // extern "C"
