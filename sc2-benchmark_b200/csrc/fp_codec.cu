// fp_codec.cu -- one C call per batch for the factorized-prior bottleneck: the whole of FPBasedResNetBottleneck.encode
// (g_a -> quantise -> rANS encode -> pack; sc2bench/models/layer.py:496-507) or .decode (rANS decode -> dequantise -> g_s;
// layer.py:509-521) is enqueued on caller-provided streams into caller-provided buffers.
//
// Why: the host was the unstable part of the pipelined step.  Issued kernel by kernel from Python (12 launches, 16 allocator
// calls and ~30 stream / event calls per batch) a step costs ~1 ms of host time when the process has a quiet core, and 3-8 ms
// when it does not -- one bench run in four was host-bound at 6-9 ms per step against 4.6 ms of GPU work (scripts/
// diag_host_issue.py, profiles/r2_host_issue.md).  Here a batch is TWO calls that allocate nothing: every intermediate lives in
// a workspace the caller allocated once (sc2_fp_workspace_bytes), the cross-stream ordering uses the caller's events.
//
// Stream contract (both calls): the transforms run on `transform_stream`, the coder on `coder_stream` (they may be the same
// stream; then the events may be NULL).  encode: transform_stream waits for `ev_in` (input ready, recorded by the caller), and
// `ev_mid` is recorded on transform_stream after g_a and waited for by coder_stream.  decode: `ev_in` = bitstreams ready (waited
// for by coder_stream), `ev_mid` is recorded on coder_stream after the decoder and waited for by transform_stream; `ev_out` (if not
// NULL) is recorded on the stream that produced the call's result.
#include "common.cuh"

extern "C" {

static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

struct FpGeometry {
    int h1, w1, hp, wp;        // first conv output and its parity planes
    int h2, w2;                // second conv output
    int h3, w3;                // latent
    int hd1, wd1, hd2, wd2, hd3, wd3;  // decoder outputs
    int c1p, c2p;              // channel pitches of the g_a planes (rounded up to 8)
    int64_t n_sym;             // symbols per image
    // g_a workspace offsets
    int64_t y1_hi, y1_lo, y2_hi, y2_lo, ga_total;
    // g_s workspace offsets
    int64_t nhwc, x1, s1, y1d, x2, s2, y2d, gs_total;
};

static int fp_geometry(const sc2_fp_plan *p, FpGeometry *g) {
    if (!p || p->batch < 1 || p->h_in < 8 || p->w_in < 8) return SC2_ERR_INVALID_ARG;
    if (p->k1 != 5 || p->k2 != 5) return SC2_ERR_UNSUPPORTED;
    g->h1 = (p->h_in + 4 - 5) / 2 + 1; g->w1 = (p->w_in + 4 - 5) / 2 + 1;
    if ((g->h1 & 1) || (g->w1 & 1)) return SC2_ERR_UNSUPPORTED;
    g->hp = g->h1 / 2; g->wp = g->w1 / 2;
    g->h2 = (g->h1 + 4 - 5) / 2 + 1; g->w2 = (g->w1 + 4 - 5) / 2 + 1;
    g->h3 = g->h2 + 2 * p->p3 - p->k3 + 1; g->w3 = g->w2 + 2 * p->p3 - p->k3 + 1;
    g->hd1 = g->h3 + 2 * p->pd1 - p->kd1 + 1; g->wd1 = g->w3 + 2 * p->pd1 - p->kd1 + 1;
    g->hd2 = g->hd1 + 2 * p->pd2 - p->kd2 + 1; g->wd2 = g->wd1 + 2 * p->pd2 - p->kd2 + 1;
    g->hd3 = g->hd2 + 2 * p->pd3 - p->kd3 + 1; g->wd3 = g->wd2 + 2 * p->pd3 - p->kd3 + 1;
    if (g->h3 < 1 || g->w3 < 1 || g->hd3 < 1 || g->wd3 < 1) return SC2_ERR_INVALID_ARG;
    if (p->d1 % 64 || p->d2 % 64 || p->d3 % 64 || p->d1 % 32 || p->d2 % 32) return SC2_ERR_UNSUPPORTED;
    g->c1p = (p->c1 + 7) / 8 * 8; g->c2p = (p->c2 + 7) / 8 * 8;
    g->n_sym = static_cast<int64_t>(p->c3) * g->h3 * g->w3;
    const int64_t B = p->batch;
    int64_t o = 0;
    g->y1_hi = o; o += align256(B * 4 * g->hp * g->wp * g->c1p * 2);
    g->y1_lo = o; o += align256(B * 4 * g->hp * g->wp * g->c1p * 2);
    g->y2_hi = o; o += align256(B * g->h2 * g->w2 * g->c2p * 2);
    g->y2_lo = o; o += align256(B * g->h2 * g->w2 * g->c2p * 2);
    g->ga_total = o;
    o = 0;
    const int64_t c3p = (p->c3 + 63) / 64 * 64;
    g->nhwc = o; o += align256(B * g->h3 * g->w3 * c3p * 2);
    g->x1 = o; o += align256(B * g->hd1 * g->wd1 * p->d1 * 2);
    g->s1 = o; o += align256(B * g->hd1 * g->wd1 * (p->d1 / 32) * 4);
    g->y1d = o; o += align256(B * g->hd1 * g->wd1 * p->d1 * 2);
    g->x2 = o; o += align256(B * g->hd2 * g->wd2 * p->d2 * 2);
    g->s2 = o; o += align256(B * g->hd2 * g->wd2 * (p->d2 / 32) * 4);
    g->y2d = o; o += align256(B * g->hd2 * g->wd2 * p->d2 * 2);
    g->gs_total = o;
    return SC2_OK;
}

int sc2_fp_workspace_bytes(const sc2_fp_plan *plan, int64_t *ga_bytes, int64_t *gs_bytes, int64_t *symbols_per_image,
                           int *latent_h, int *latent_w, int *out_h, int *out_w) {
    FpGeometry g;
    const int rc = fp_geometry(plan, &g);
    if (rc) return rc;
    if (ga_bytes) *ga_bytes = g.ga_total;
    if (gs_bytes) *gs_bytes = g.gs_total;
    if (symbols_per_image) *symbols_per_image = g.n_sym;
    if (latent_h) *latent_h = g.h3;
    if (latent_w) *latent_w = g.w3;
    if (out_h) *out_h = g.hd3;
    if (out_w) *out_w = g.wd3;
    return SC2_OK;
}

#define SC2_TRY(expr)            \
    do {                         \
        const int rc_ = (expr);  \
        if (rc_) return rc_;     \
    } while (0)

int sc2_fp_encode_batch(const sc2_fp_plan *p, const void *image, int image_is_u8, void *ws_ga, int32_t *symbols, uint8_t *arena,
                        int64_t slot_bytes, int32_t *lengths, uint8_t *packed, int64_t *offsets, int32_t *status,
                        int32_t *tile_counters, int coder_layout, sc2_stream_t transform_stream, sc2_stream_t coder_stream,
                        void *ev_in, void *ev_mid, void *ev_out) {
    FpGeometry g;
    SC2_TRY(fp_geometry(p, &g));
    if (!image || !ws_ga || !symbols || !arena || !lengths || !packed || !offsets || !status || !tile_counters) return SC2_ERR_INVALID_ARG;
    cudaStream_t ts = sc2::as_stream(transform_stream), cs = sc2::as_stream(coder_stream);
    if (ts != cs && !ev_mid) return SC2_ERR_INVALID_ARG;
    uint8_t *ws = static_cast<uint8_t *>(ws_ga);
    if (ev_in) SC2_CUDA_TRY(cudaStreamWaitEvent(ts, static_cast<cudaEvent_t>(ev_in), 0));
    SC2_CUDA_TRY(cudaMemsetAsync(tile_counters, 0, 3 * sizeof(int32_t), ts));
    // ---- g_a: three launches ----
    SC2_TRY(sc2_ga_first_conv_gdn(image, image_is_u8, p->lut, p->batch, p->h_in, p->w_in, p->c1, p->w1_stack, p->g1_stack, p->beta1,
                                  ws + g.y1_hi, ws + g.y1_lo, g.c1p, tile_counters + 0, transform_stream));
    sc2_ga_halo_desc hd;
    hd.images = p->batch; hd.h_in = g.hp; hd.w_in = g.wp; hd.c_in = g.c1p; hd.c_out = p->c2; hd.kh = hd.kw = p->k2; hd.pad = 2;
    hd.h_out = g.h2; hd.w_out = g.w2; hd.out_c = g.c2p;
    SC2_TRY(sc2_ga_halo_conv_gdn(&hd, ws + g.y1_hi, ws + g.y1_lo, p->w2_stack, p->g2_stack, p->beta2, ws + g.y2_hi, ws + g.y2_lo,
                                 tile_counters + 1, transform_stream));
    sc2_tc_split_desc sd;
    sd.images = p->batch; sd.h_in = g.h2; sd.w_in = g.w2; sd.c_in = g.c2p; sd.c_out = p->c3; sd.kh = sd.kw = p->k3; sd.stride = 1;
    sd.pad = p->p3; sd.mode = SC2_TCS_QUANT; sd.h_out = g.h3; sd.w_out = g.w3; sd.out_c = p->c3;
    SC2_TRY(sc2_tc_split_conv(&sd, ws + g.y2_hi, ws + g.y2_lo, p->w3_hi, p->w3_lo, nullptr, p->medians, nullptr, nullptr, nullptr, nullptr,
                              symbols, tile_counters + 2, transform_stream));
    if (ts != cs) {
        SC2_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(ev_mid), ts));
        SC2_CUDA_TRY(cudaStreamWaitEvent(cs, static_cast<cudaEvent_t>(ev_mid), 0));
    }
    // ---- coder ----
    SC2_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), cs));
    SC2_TRY(sc2_rans_encode_batch(symbols, nullptr, p->batch, g.n_sym, static_cast<int64_t>(g.h3) * g.w3, p->tables, p->n_rows, p->cdf_stride,
                                  arena, slot_bytes, lengths, status, coder_layout, coder_stream));
    SC2_TRY(sc2_rans_pack(arena, slot_bytes, lengths, p->batch, packed, offsets, coder_stream));
    if (ev_out) SC2_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(ev_out), cs));
    return SC2_OK;
}

int sc2_fp_decode_batch(const sc2_fp_plan *p, const uint8_t *packed, const int64_t *offsets, float *latent_hat, void *ws_gs,
                        float *out, int32_t *status, int32_t *tile_counters, int coder_layout, sc2_stream_t coder_stream,
                        sc2_stream_t transform_stream, void *ev_in, void *ev_mid, void *ev_out) {
    FpGeometry g;
    SC2_TRY(fp_geometry(p, &g));
    if (!latent_hat || !ws_gs || !out || !tile_counters) return SC2_ERR_INVALID_ARG;
    cudaStream_t ts = sc2::as_stream(transform_stream), cs = sc2::as_stream(coder_stream);
    uint8_t *ws = static_cast<uint8_t *>(ws_gs);
    if (packed) {
        if (!offsets || !status) return SC2_ERR_INVALID_ARG;
        if (ts != cs && !ev_mid) return SC2_ERR_INVALID_ARG;
        if (ev_in) SC2_CUDA_TRY(cudaStreamWaitEvent(cs, static_cast<cudaEvent_t>(ev_in), 0));
        SC2_TRY(sc2_rans_decode_batch(packed, offsets, p->batch, g.n_sym, nullptr, static_cast<int64_t>(g.h3) * g.w3, p->tables, p->n_rows,
                                      p->cdf_stride, nullptr, latent_hat, p->medians, status, coder_layout, coder_stream));
        if (ts != cs) {
            SC2_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(ev_mid), cs));
            SC2_CUDA_TRY(cudaStreamWaitEvent(ts, static_cast<cudaEvent_t>(ev_mid), 0));
        }
    } else if (ev_in) {  // packed == NULL: latent_hat is already there (the caller decoded); only g_s runs, after ev_in
        SC2_CUDA_TRY(cudaStreamWaitEvent(ts, static_cast<cudaEvent_t>(ev_in), 0));
    }
    SC2_CUDA_TRY(cudaMemsetAsync(tile_counters, 0, 5 * sizeof(int32_t), ts));
    // ---- g_s: layout change + five launches ----
    const int c3p = (p->c3 + 63) / 64 * 64;
    SC2_TRY(sc2_nchw_f32_to_nhwc_f16(latent_hat, ws + g.nhwc, p->batch, p->c3, static_cast<int64_t>(g.h3) * g.w3, c3p, transform_stream));
    sc2_tc_conv_desc d;
    d.batch = p->batch;
    d.h_in = g.h3; d.w_in = g.w3; d.c_in_pad = c3p; d.c_out = p->d1; d.kh = d.kw = p->kd1; d.pad = p->pd1; d.mode = SC2_TC_STORE_ABS_F16;
    SC2_TRY(sc2_tc_conv_nhwc(&d, ws + g.nhwc, p->wd1, nullptr, nullptr, ws + g.x1, reinterpret_cast<uint32_t *>(ws + g.s1), tile_counters + 0,
                             transform_stream));
    d.h_in = g.hd1; d.w_in = g.wd1; d.c_in_pad = p->d1; d.c_out = p->d1; d.kh = d.kw = 1; d.pad = 0; d.mode = SC2_TC_IGDN1_ABS_F16;
    SC2_TRY(sc2_tc_conv_nhwc(&d, ws + g.x1, p->gd1, p->betad1, ws + g.x1, ws + g.y1d, reinterpret_cast<uint32_t *>(ws + g.s1),
                             tile_counters + 1, transform_stream));
    d.c_out = p->d2; d.kh = d.kw = p->kd2; d.pad = p->pd2; d.mode = SC2_TC_STORE_ABS_F16;
    SC2_TRY(sc2_tc_conv_nhwc(&d, ws + g.y1d, p->wd2, nullptr, nullptr, ws + g.x2, reinterpret_cast<uint32_t *>(ws + g.s2), tile_counters + 2,
                             transform_stream));
    d.h_in = g.hd2; d.w_in = g.wd2; d.c_in_pad = p->d2; d.c_out = p->d2; d.kh = d.kw = 1; d.pad = 0; d.mode = SC2_TC_IGDN1_ABS_F16;
    SC2_TRY(sc2_tc_conv_nhwc(&d, ws + g.x2, p->gd2, p->betad2, ws + g.x2, ws + g.y2d, reinterpret_cast<uint32_t *>(ws + g.s2),
                             tile_counters + 3, transform_stream));
    d.c_out = p->d3; d.kh = d.kw = p->kd3; d.pad = p->pd3; d.mode = SC2_TC_STORE_F32;
    SC2_TRY(sc2_tc_conv_nhwc(&d, ws + g.y2d, p->wd3, nullptr, nullptr, out, nullptr, tile_counters + 4, transform_stream));
    if (ev_out) SC2_CUDA_TRY(cudaEventRecord(static_cast<cudaEvent_t>(ev_out), ts));
    return SC2_OK;
}

}  // extern "C"
