/*
 * sc2b200.h -- C ABI of libsc2b200.so: the B200 (sm_100a) implementation of sc2bench's
 * supervised-compression bottleneck path (g_a conv+GDN -> EntropyBottleneck quantise ->
 * rANS encode -> rANS decode -> g_s IGDN+conv).
 *
 * Boundary rules:
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - the CALLER allocates every buffer; the library allocates nothing that outlives a call,
 *     keeps no global state (apart from a lazily resolved driver entry point) and never frees
 *     caller memory;
 *   - every device entry point takes an explicit CUDA stream (cudaStream_t passed as void*),
 *     only enqueues work on it and never synchronises the device;
 *   - return value: 0 = ok, <0 = invalid argument / launch failure (SC2_ERR_*), device-side
 *     faults (arena overflow, truncated stream) are reported asynchronously through the
 *     caller-provided `status` words (SC2_FAULT_* bit flags);
 *   - nothing throws across the boundary.
 *
 * What each entry point replaces in the reference stack (the reference itself is pure Python;
 * the FFI underneath its hot path is CompressAI's two pybind modules):
 *   compressai._CXX.pmf_to_quantized_cdf                 -> sc2_pmf_to_quantized_cdf
 *       reached from sc2bench/models/layer.py:441 (BaseBottleneck.update)
 *   compressai.ans.RansEncoder.encode_with_indexes       -> sc2_rans_encode_batch (+ sc2_rans_pack)
 *       reached from sc2bench/models/layer.py:506 (entropy_bottleneck.compress), :647 (gaussian_conditional.compress)
 *   compressai.ans.RansDecoder.decode_with_indexes       -> sc2_rans_decode_batch
 *       reached from sc2bench/models/layer.py:520 (entropy_bottleneck.decompress), :665
 *   EntropyModel.quantize("symbols") / dequantize        -> sc2_quantize_symbols / fused into decode
 *       reached from sc2bench/models/layer.py:506,520,545-546
 *   GaussianConditional.build_indexes                    -> sc2_gc_build_indexes
 *       reached from sc2bench/models/layer.py:646,664
 *   torch conv2d / conv_transpose2d + compressai.layers.GDN / GDN1 (library calls)
 *                                                        -> sc2_conv2d_f32 / sc2_gdn_f32 (+ sc2_tc_* tensor-core path)
 *       instantiated at sc2bench/models/layer.py:475-494
 */
#ifndef SC2B200_H
#define SC2B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SC2_ABI_VERSION 10

#if defined(__GNUC__)
#define SC2_API __attribute__((visibility("default")))
#else
#define SC2_API
#endif

/* return codes */
#define SC2_OK 0
#define SC2_ERR_INVALID_ARG (-1)
#define SC2_ERR_UNSUPPORTED (-2)
#define SC2_ERR_CUDA (-3)
#define SC2_ERR_DOMAIN (-4) /* pmf has a negative / non-finite entry, or sums to zero */

/* device-side fault flags written (OR-ed) into a caller-provided int32 status word */
#define SC2_FAULT_ARENA_OVERFLOW 1   /* encoder ran out of its output slot */
#define SC2_FAULT_STREAM_TRUNCATED 2 /* decoder would read past the end of a stream */
#define SC2_FAULT_BAD_STREAM 4       /* stream shorter than 8 bytes or not a multiple of 4 */
#define SC2_FAULT_BAD_INDEX 8        /* a caller-supplied CDF index is outside [0, n_rows): that symbol was coded with row 0 */

typedef void *sc2_stream_t; /* cudaStream_t */

SC2_API int sc2_abi_version(void);
/* Tuning knob (process-wide): CTAs of the persistent tensor-core kernels, 1..148; 0 restores the default (SC2_TC_GRID environment
 * variable, else one per SM).  Fewer CTAs leave SMs to kernels of other streams (the coders of batches in flight). */
SC2_API int sc2_set_persistent_ctas(int ctas);
SC2_API const char *sc2_error_string(int code);
/* last CUDA error string seen by this thread inside the library (for SC2_ERR_CUDA) */
SC2_API const char *sc2_last_cuda_error(void);

/* ------------------------------------------------------------------------------------------
 * Host: CDF construction.  Bit-exact restatement of compressai._CXX.pmf_to_quantized_cdf.
 * cdf_out has n + 1 entries.  Runs once per update(), never on the per-image path. */
SC2_API int sc2_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf_out);

/* ------------------------------------------------------------------------------------------
 * Host: coder tables.  From CompressAI's three buffers (_quantized_cdf [n_rows x cdf_stride] int32,
 * _cdf_length [n_rows], _offset [n_rows]) build the blob the kernels read:
 * per-entry exact-division reciprocals for the encoder, sentinel-padded CDF rows for the decoder.
 * The caller uploads the blob to device memory (any 16-byte aligned address). */
SC2_API size_t sc2_rans_table_bytes(int n_rows, int cdf_stride);
SC2_API int sc2_rans_build_tables(const int32_t *cdfs, const int32_t *cdf_sizes, const int32_t *offsets,
                          int n_rows, int cdf_stride, void *blob_out);

/* Bytes one stream can need in the worst case (every symbol escaped with 8 nibbles). */
SC2_API int64_t sc2_rans_max_stream_bytes(int64_t n_symbols);

/* ------------------------------------------------------------------------------------------
 * Device: batched rANS encode, one CompressAI-compatible stream per sample.
 *   symbols   [batch x n_per_stream] int32, C order within a sample (c, h, w)
 *   indexes   [batch x n_per_stream] int32 CDF row per symbol, or NULL for "channel mode":
 *             row = (i / spatial) for symbol i of a sample (EntropyBottleneck._build_indexes)
 *   tables    device copy of the sc2_rans_build_tables blob built for (n_rows, cdf_stride)
 *   arena     batch slots of slot_bytes (multiple of 4); stream b is written BACKWARDS from the end
 *             of slot b and occupies its last lengths[b] bytes
 *   lengths   [batch] int32 out, bytes
 *   status    [1] int32, OR-ed fault flags (caller zeroes it)
 *   layout    channel mode only (a stream is one serial chain; the layouts produce identical bytes):
 *             SC2_RANS_WARP_PER_STREAM  lowest latency for ONE batch: a warp per stream, 25 instructions per symbol
 *             SC2_RANS_LANE_PER_STREAM  32 streams per warp: ~2x the latency of a batch but 1/16 of the issue slots and
 *                                       one SM per 256 streams -- for batches in flight next to the tensor-core kernels
 *             SC2_RANS_AUTO             warp per stream (or what the SC2_CODER=warp|lanes environment variable says) */
#define SC2_RANS_AUTO 0
#define SC2_RANS_WARP_PER_STREAM 1
#define SC2_RANS_LANE_PER_STREAM 2
SC2_API int sc2_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int batch, int64_t n_per_stream,
                          int64_t spatial, const void *tables, int n_rows, int cdf_stride, uint8_t *arena,
                          int64_t slot_bytes, int32_t *lengths, int32_t *status, int layout, sc2_stream_t stream);

/* Device: compact the streams to the front of `packed`: offsets[b] = sum(lengths[:b]) (int64, batch+1
 * entries, offsets[batch] = total).  One D2H copy of offsets[batch] bytes then carries all strings. */
SC2_API int sc2_rans_pack(const uint8_t *arena, int64_t slot_bytes, const int32_t *lengths, int batch,
                  uint8_t *packed, int64_t *offsets, sc2_stream_t stream);

/* Device: batched rANS decode (+ fused dequantise).
 *   packed/offsets  streams laid out as sc2_rans_pack produces them (offsets: batch+1 int64, device)
 *   indexes/spatial as for encode
 *   out_symbols     [batch x n_per_stream] int32 or NULL
 *   out_values      [batch x n_per_stream] float32 or NULL: float(symbol) + means[row]   (means may be NULL -> +0)
 *   means           [n_rows] float32 per-row medians (EntropyBottleneck) or NULL */
SC2_API int sc2_rans_decode_batch(const uint8_t *packed, const int64_t *offsets, int batch, int64_t n_per_stream,
                          const int32_t *indexes, int64_t spatial, const void *tables, int n_rows, int cdf_stride,
                          int32_t *out_symbols, float *out_values, const float *means,
                          int32_t *status, int layout, sc2_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Device: EntropyModel.quantize(x, "symbols", means): int32(rint(x - mean[c])) over [batch, C, spatial].
 * means may be NULL.  (Fused into the last g_a conv on the tensor-core path.) */
SC2_API int sc2_quantize_symbols(const float *x, const float *means, int32_t *symbols, int batch, int channels,
                         int64_t spatial, sc2_stream_t stream);

/* Device: EntropyModel.dequantize(symbols, means) with one mean per ELEMENT (mean-scale hyperprior,
 * sc2bench/models/layer.py:785): out[i] = float(symbols[i]) + means[i].  means may be NULL. */
SC2_API int sc2_dequantize(const int32_t *symbols, const float *means, float *out, int64_t n, sc2_stream_t stream);

/* Device: GaussianConditional.build_indexes: idx = #(table[:-1] < max(scale, bound)) per element. */
SC2_API int sc2_gc_build_indexes(const float *scales, int64_t n, const float *scale_table, int n_levels,
                         float scale_bound, int32_t *indexes, sc2_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Device: fp32 (CUDA-core) direct convolution, NCHW, with a fused epilogue.  Exact-fp32 path used by
 * g_a (symbol exactness needs fp32-grade latents) and as the general fallback. */
#define SC2_EPI_NONE 0
#define SC2_EPI_RELU 1
#define SC2_EPI_CLAMP01 2
#define SC2_EPI_QUANTIZE 3 /* y -> int32(rint(y - aux[c])) written to out as int32 (aux = medians or NULL) */
#define SC2_EPI_ABS 4
#define SC2_EPI_LEAKY_RELU 5 /* y > 0 ? y : y * epi_param (nn.LeakyReLU; sc2bench/models/layer.py:611-616,760-770) */
#define SC2_IN_NONE 0
#define SC2_IN_ABS 1 /* the convolution reads |x| (h_a(torch.abs(y)), sc2bench/models/layer.py:642) */

typedef struct sc2_conv_desc {
    int batch, c_in, h_in, w_in;
    int c_out, kh, kw;
    int stride, pad;
    int transposed;     /* 0: Conv2d (weight [c_out, c_in, kh, kw]); 1: ConvTranspose2d (weight [c_in, c_out, kh, kw]) */
    int output_padding; /* transposed only */
    int epilogue;       /* SC2_EPI_* */
    int in_transform;   /* SC2_IN_* */
    float epi_param;    /* negative slope for SC2_EPI_LEAKY_RELU */
} sc2_conv_desc;

/* output spatial size for a descriptor */
SC2_API int sc2_conv_out_size(const sc2_conv_desc *d, int *h_out, int *w_out);

SC2_API int sc2_conv2d_f32(const sc2_conv_desc *d, const float *x, const float *weight, const float *bias /*nullable*/,
                   const float *aux /*nullable*/, void *out, sc2_stream_t stream);

/* Device: GDN family on NCHW fp32.  gamma [C x C] and beta [C] are the EFFECTIVE (reparametrised)
 * values.  kind 0 = GDN1 (norm = beta + gamma.|x|; y = x / norm, inverse: x * norm)
 *          kind 1 = GDN  (norm = beta + gamma.x^2; y = x * rsqrt(norm), inverse: x * sqrt(norm)) */
SC2_API int sc2_gdn_f32(const float *x, const float *gamma, const float *beta, float *y, int batch, int channels,
                int64_t spatial, int kind, int inverse, sc2_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Device: tensor-core (tcgen05 + TMEM + TMA) implicit-GEMM convolution on NHWC fp16 activations, stride 1.
 * Used for g_s (sc2bench/models/layer.py:485-494), where fp16 operands with fp32 accumulation meet the 1e-3 bar.
 *   x         [batch, h_in, w_in, c_in_pad] fp16, c_in_pad a multiple of 64 (zero-padded channels)
 *   w_packed  [kh*kw, c_out, c_in_pad] fp16  (tap-major repack of the Conv2d weight; for the GDN modes: gamma [c, c])
 *   out       [batch, h_out, w_out, c_out]  fp16 or fp32 (mode), h_out = h_in + 2*pad - kh + 1
 *   modes     0 store fp16 | 1 store fp32 | 2 IGDN1: out = gdn_x * (beta + acc) | 3 GDN1: out = gdn_x / (beta + acc)
 *             (GDN modes: 1x1 "gamma" GEMM over |x|, gdn_x = x itself, c_in_pad == c_out)
 *             4 store |acc| as fp16 + one sign bit per value in `signs` (the conv in front of an IGDN1)
 *             5 IGDN1 on such a pair: x = |x| tensor (also the A operand, no in-kernel |.| pass), gdn_x = the same |x|,
 *               signs = the sign words; out = sign * |x| * (beta + acc)
 *   signs     modes 4 (out) / 5 (in): uint32 [batch * h_out * w_out, c_out / 32]; word layout: conv_tc.cu; else NULL */
#define SC2_TC_STORE_F16 0
#define SC2_TC_STORE_F32 1
#define SC2_TC_IGDN1_F16 2
#define SC2_TC_GDN1_F16 3
#define SC2_TC_STORE_ABS_F16 4
#define SC2_TC_IGDN1_ABS_F16 5

typedef struct sc2_tc_conv_desc {
    int batch, h_in, w_in, c_in_pad;
    int c_out, kh, kw, pad;
    int mode;
} sc2_tc_conv_desc;

/* Extended form (round 2): asymmetric kernels / paddings, an explicit output grid, strided placement of that grid in the output
 * tensor, a per-channel bias, and the modes the CompressAI zoo codecs' synthesis transforms need (GDN proper, ConvTranspose2d;
 * sc2bench/models/registry.py:12-14, wrapper.py:119-135).  A ConvTranspose2d(k5, s2, p2, op1) is four launches, one per output
 * parity (py, px): a stride-1 correlation with (3 or 2) x (3 or 2) taps, pad (1 or 0), written to every second output pixel.
 *   h_out, w_out   grid of output pixels this launch computes; pixel (oy, ox) of the grid is output-tensor pixel
 *                  (oy * out_stride + out_py, ox * out_stride + out_px) of an [batch, out_h, out_w, c_out] NHWC tensor
 *   vec            conv modes: bias [c_out] or NULL; GDN modes: effective beta [c_out]
 *   modes          0..5 as above, plus
 *                  6 store x as fp16 to `out` AND x*x/256 as fp16 to `out2` (the layer in front of an inverse GDN)
 *                  7 inverse GDN on such a pair: x = the x*x/256 tensor (A operand), gdn_x = the x tensor,
 *                    out = gdn_x * sqrt(beta + 256 * acc)
 *                  8 last layer: c_out <= 32 real channels (weights packed to 32 rows per tap), bias, clamp to [0, 1],
 *                    `out` is fp32 NCHW [batch, c_out, out_h, out_w] */
#define SC2_TC_STORE_SQ_F16 6
#define SC2_TC_IGDN_SQ_F16 7
#define SC2_TC_NCHW_F32_CLAMP 8

typedef struct sc2_tc_conv_ex_desc {
    int batch, h_in, w_in, c_in_pad;
    int c_out, kh, kw, pad_y, pad_x;
    int mode;
    int h_out, w_out;
    int out_h, out_w, out_stride, out_py, out_px;
} sc2_tc_conv_ex_desc;

/* Split activations (optional, all NULL = plain fp16): fp16 activations between the LAST layers of a synthesis transform cost
 * 1.5e-3 on x_hat (scripts/diag note in DESIGN.md), so that stage can carry x = hi + lo (two fp16 tensors):
 *   x_lo       second A tensor: the MMA runs over x and x_lo with the same weights (two passes into one accumulator)
 *   out3       mode 6: also store the lo half of x
 *   gdn_x_lo   mode 7: lo half of the multiplicand x; the result is then stored as out (hi) + out2 (lo) */
SC2_API int sc2_tc_conv_ex(const sc2_tc_conv_ex_desc *d, const void *x, const void *x_lo, const void *w_packed, const float *vec,
                           const void *gdn_x, const void *gdn_x_lo, void *out, void *out2, void *out3, uint32_t *signs,
                           int32_t *tile_counter, sc2_stream_t stream);

/* tile_counter (all sc2_tc_* kernels): the kernels are persistent (one CTA per SM).  With a caller-provided int32 that is
 * ZERO at launch (one per launch; stream-ordered reuse is fine) the CTAs claim tiles dynamically, so a CTA that is placed
 * late -- its SM busy with another stream's blocks -- does not delay the kernel; NULL selects the static schedule. */
SC2_API int sc2_tc_conv_nhwc(const sc2_tc_conv_desc *d, const void *x, const void *w_packed, const float *beta,
                             const void *gdn_x, void *out, uint32_t *signs, int32_t *tile_counter, sc2_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Device: fp32-grade tensor-core convolution ("split fp16", three tcgen05 passes; conv_tc_split.cu) for g_a
 * (sc2bench/models/layer.py:475-484).  Every tensor is a pair of fp16 planes (hi, lo): value = hi + lo / 2048.
 *   x_hi/x_lo   stride 1: [images, h_in, w_in, c_in]; stride 2: PARITY PLANES [images * 4, h_in, w_in, c_in] where
 *               plane py*2+px holds input pixel (2Y+py, 2X+px) at (Y, X)   (h_in, w_in = plane geometry)
 *   w_hi/w_lo   [kh*kw, n_tile, c_in] tap-major packed weights, n_tile = sc2_tc_split_n_tile(c_out), extra rows zero
 *   modes       0 store split planes [images, h_out, w_out, out_c] | 1 GDN1 forward (1x1: w = gamma, gdn_x = x itself)
 *               | 2 quantise: int32(rint(y - medians[c])) to out_sym in NCHW (coder) order
 *   c_in        multiple of 16 (zero-pad the patches / weights) */
#define SC2_TCS_STORE 0
#define SC2_TCS_GDN1 1
#define SC2_TCS_QUANT 2

typedef struct sc2_tc_split_desc {
    int images, h_in, w_in, c_in;
    int c_out, kh, kw, stride, pad;
    int mode;
    int h_out, w_out, out_c;
} sc2_tc_split_desc;

SC2_API int sc2_tc_split_n_tile(int c_out);
SC2_API int sc2_tc_split_conv(const sc2_tc_split_desc *d, const void *x_hi, const void *x_lo, const void *w_hi,
                              const void *w_lo, const float *beta, const float *medians, const void *gdn_x_hi,
                              const void *gdn_x_lo, void *out_hi, void *out_lo, int32_t *out_sym, int32_t *tile_counter,
                              sc2_stream_t stream);

/* sc2_tc_split_conv for wider layers and the other layer types on the path (analysis / hyper-analysis transforms of the CompressAI
 * zoo codecs, sc2bench/models/registry.py:12-14 -> compressai.models.google [mem]; g_a of the hyperprior bottlenecks, layer.py:596-603):
 *   N tiling    one launch computes output channels [n_off, n_off + c_out) (c_out <= 128) of planes whose pixels are out_pitch
 *               channels apart; w_hi/w_lo hold that tile's rows; vec / medians are the FULL vectors (indexed from n_off);
 *               mode 2 writes symbols [images, c_total, h_out, w_out].  An inner tile must be full (c_out == n_tile).
 *   in_nhwc     stride 2 only: x planes are a plain NHWC tensor [images, h_in, w_in, c_in] (h_in, w_in even = FULL resolution),
 *               read through a 5-D tensor map (pixel parity as a coordinate) -- no parity-plane re-layout between two stride-2 layers
 *   vec         modes 0 / 2: the convolution's bias (or NULL), added before act / the quantiser; modes 1 / 3: the GDN beta
 *   act         mode 0: 0 none | 1 ReLU | 2 LeakyReLU(slope)
 *   mode 3      GDN proper: y = x / sqrt(beta + gamma . x^2) (1x1: w = gamma, the squares are formed in shared memory)
 *   pad_x       horizontal padding when it differs from pad (< 0: pad_x = pad)
 *   out_stride  2 (mode 0, stride 1): the h_out x w_out results are the pixels (2Y + out_py, 2X + out_px) of planes
 *               [images, out_h, out_w, out_pitch] (out_h = out_w = 0: 2 h_out x 2 w_out; h_out, w_out must be the pixel counts of that
 *               parity) -- one of the four stride-1 sub-convolutions of a ConvTranspose2d(k5, s2): p2 / op1 in the zoo codecs' h_s,
 *               p1 / op0 (odd output sizes) in the hyperprior bottlenecks' h_s (sc2bench/models/layer.py:611-617) */
#define SC2_TCS_GDN 3
#define SC2_TCS_ACT_NONE 0
#define SC2_TCS_ACT_RELU 1
#define SC2_TCS_ACT_LEAKY 2

typedef struct sc2_tc_split_ex_desc {
    int images, h_in, w_in, c_in;
    int c_out, kh, kw, stride, pad;
    int mode;
    int h_out, w_out;
    int out_pitch, n_off, c_total;
    int in_nhwc;
    int act;
    float slope;
    int pad_x;
    int out_stride, out_py, out_px;
    int out_h, out_w;
} sc2_tc_split_ex_desc;

SC2_API int sc2_tc_split_conv_ex(const sc2_tc_split_ex_desc *d, const void *x_hi, const void *x_lo, const void *w_hi,
                                 const void *w_lo, const float *vec, const float *medians, const void *gdn_x_hi,
                                 const void *gdn_x_lo, void *out_hi, void *out_lo, int32_t *out_sym, int32_t *tile_counter,
                                 sc2_stream_t stream);

/* Device: stride-2 convolution + GDN1 back to back in one kernel (conv_ga_halo.cu), fp32-grade split fp16: the middle of g_a
 * (sc2bench/models/layer.py:479-481, Conv2d(k5, s2, p2) -> GDN1).  The conv accumulators never leave the SM: |x| becomes the
 * A operand of the 1x1 gamma GEMM in shared memory, y = x / (beta + gamma.|x|) leaves as split planes.
 *   x_hi/x_lo     PARITY PLANES [images * 4, h_in, w_in, c_in] (as for sc2_tc_split_conv, stride 2)
 *   w_stack       [kh*kw, 2n, c_in] fp16: per tap, rows [0, n) = hi and rows [n, 2n) = lo of the weights, n = sc2_ga_halo_n(c_out)
 *   gamma_stack   [2n, n] fp16: the effective gamma [c_out, c_out] packed the same way (zero padded)
 *   out_hi/out_lo [images, h_out, w_out, out_c] split planes, out_c = c_out rounded up to 8 */
typedef struct sc2_ga_halo_desc {
    int images, h_in, w_in, c_in;
    int c_out, kh, kw, pad;
    int h_out, w_out, out_c;
} sc2_ga_halo_desc;

SC2_API int sc2_ga_halo_n(int c_out);
SC2_API int sc2_ga_halo_conv_gdn(const sc2_ga_halo_desc *d, const void *x_hi, const void *x_lo, const void *w_stack,
                                 const void *gamma_stack, const float *beta, void *out_hi, void *out_lo, int32_t *tile_counter,
                                 sc2_stream_t stream);

/* Device: first layer of g_a + its GDN1 in one kernel (conv_ga_first.cu): Conv2d(3 -> c_out, k5, s2, p2) on an NCHW image followed
 * by GDN1(c_out) (sc2bench/models/layer.py:476-478), fp32-grade split fp16, output as split PARITY PLANES
 * [batch * 4, h_out/2, w_out/2, out_c] -- the input layout of sc2_ga_halo_conv_gdn.
 *   image        fp32 [batch, 3, h_in, w_in] (w_in % 4 == 0), or, with image_is_u8, uint8 of the same shape (w_in % 16 == 0)
 *   lut          uint8 input only: float [3][256], the value of byte v of channel c after the data loader's ToTensor + Normalize
 *                ((v / 255 - mean[c]) / std[c], computed by the caller with the reference's own ops); the conversion happens on
 *                the fly inside the im2col (device-side pre-transform, sc2bench/transforms: H2D traffic drops 4x)
 *   w_stack      [2n, 80] fp16: rows [0, n) hi / [n, 2n) lo of weight.reshape(c_out, 75), zero padded; n = sc2_ga_halo_n(c_out)
 *   gamma_stack  [2n, n] fp16 as for sc2_ga_halo_conv_gdn */
SC2_API int sc2_ga_first_conv_gdn(const void *image, int image_is_u8, const float *lut, int batch, int h_in, int w_in, int c_out,
                                  const void *w_stack, const void *gamma_stack, const float *beta, void *out_hi, void *out_lo,
                                  int out_c, int32_t *tile_counter, sc2_stream_t stream);

/* Device: im2col of an fp32 NCHW image for the first (c_in = 3) layer: split fp16 patches, K = (c, dy, dx) zero-padded
 * to k_pad, pixels in parity-plane order [batch * 4, h_out/2, w_out/2, k_pad] (h_out, w_out must be even). */
SC2_API int sc2_patchify_split(const float *x, void *out_hi, void *out_lo, int batch, int c_in, int h_in, int w_in,
                               int kh, int kw, int stride, int pad, int k_pad, sc2_stream_t stream);

/* The same im2col with the pixels in plain NHWC order [batch, h_out, w_out, k_pad] (any stride >= 1, any output size): the first
 * layer of the zoo codecs' g_a, whose next layer reads NHWC through sc2_tc_split_conv_ex(in_nhwc). */
SC2_API int sc2_patchify_split_nhwc(const float *x, void *out_hi, void *out_lo, int batch, int c_in, int h_in, int w_in,
                                    int kh, int kw, int stride, int pad, int k_pad, sc2_stream_t stream);

/* Device: first layer of g_a with the im2col fused into the tensor-core kernel (conv_tc_first.cu): stride-2 Conv2d on an fp32
 * NCHW image with c_in*kh*kw <= 128 and c_out <= 96, fp32-grade (split fp16), output as split parity planes
 * [batch * 4, h_out/2, w_out/2, out_c].  w_hi/w_lo: [1, n_tile, ceil16(c_in*kh*kw)] as for sc2_tc_split_conv. */
SC2_API int sc2_tc_first_layer(const float *image, int batch, int c_in, int h_in, int w_in, int kh, int kw, int pad, int c_out,
                               const void *w_hi, const void *w_lo, void *out_hi, void *out_lo, int out_c, int32_t *tile_counter,
                               sc2_stream_t stream);

/* Device: NCHW fp32 -> NHWC fp16 with channels zero-padded to c_pad (even). */
SC2_API int sc2_nchw_f32_to_nhwc_f16(const float *x, void *y, int batch, int channels, int64_t spatial, int c_pad,
                                     sc2_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Device: one call per batch for the factorized-prior bottleneck (fp_codec.cu) -- everything FPBasedResNetBottleneck.encode /
 * .decode do on the device (sc2bench/models/layer.py:496-521), enqueued on the caller's streams into the caller's buffers.
 * Nothing is allocated; the intermediates live in workspaces sized by sc2_fp_workspace_bytes.  All pointers are device pointers
 * owned by the caller; the packed weights are the ones the per-layer entry points take (same layouts). */
typedef struct sc2_fp_plan {
    int batch, h_in, w_in;               /* image batch [batch, 3, h_in, w_in] */
    int c1, c2, c3, k1, k2, k3, p3;      /* g_a: Conv(3->c1, k1=5, s2, p2) GDN1 Conv(c1->c2, k2=5, s2, p2) GDN1 Conv(c2->c3, k3, s1, p3) */
    int d1, d2, d3, kd1, pd1, kd2, pd2, kd3, pd3; /* g_s: Conv(c3->d1, kd1, pd1) IGDN1 Conv(d1->d2, kd2, pd2) IGDN1 Conv(d2->d3, kd3, pd3) */
    int n_rows, cdf_stride;              /* coder tables */
    const void *w1_stack, *g1_stack;     /* sc2_ga_first_conv_gdn packs */
    const float *beta1;
    const void *w2_stack, *g2_stack;     /* sc2_ga_halo_conv_gdn packs */
    const float *beta2;
    const void *w3_hi, *w3_lo;           /* sc2_tc_split_conv packs of the last g_a conv */
    const float *medians;                /* [c3] EntropyBottleneck medians */
    const float *lut;                    /* [3][256] normalisation table for uint8 images, or NULL */
    const void *tables;                  /* sc2_rans_build_tables blob */
    const void *wd1, *gd1, *wd2, *gd2, *wd3; /* sc2_tc_conv_nhwc packs (fp16) of the g_s convs / gammas */
    const float *betad1, *betad2;
} sc2_fp_plan;

/* bytes of the g_a / g_s workspaces (each shared by all batches whose transforms run on ONE stream), symbols per image and
 * the latent / output geometry; SC2_ERR_UNSUPPORTED when the fused kernels do not cover the shape */
SC2_API int sc2_fp_workspace_bytes(const sc2_fp_plan *plan, int64_t *ga_bytes, int64_t *gs_bytes, int64_t *symbols_per_image,
                                   int *latent_h, int *latent_w, int *out_h, int *out_w);

/* encode: image (fp32, or uint8 when image_is_u8 and plan->lut) -> symbols [batch, c3, h3, w3] (int32, coder order) -> streams.
 *   transform_stream waits for ev_in (may be NULL), runs g_a, records ev_mid; coder_stream waits for ev_mid, zeroes *status, runs
 *   the encoder and the pack (arena / slot_bytes / lengths / packed / offsets as for sc2_rans_encode_batch + sc2_rans_pack) and
 *   records ev_out (may be NULL).  tile_counters: 3 int32 (zeroed by the call).  Events are cudaEvent_t passed as void*. */
SC2_API int sc2_fp_encode_batch(const sc2_fp_plan *plan, const void *image, int image_is_u8, void *ws_ga, int32_t *symbols,
                                uint8_t *arena, int64_t slot_bytes, int32_t *lengths, uint8_t *packed, int64_t *offsets,
                                int32_t *status, int32_t *tile_counters, int coder_layout, sc2_stream_t transform_stream,
                                sc2_stream_t coder_stream, void *ev_in, void *ev_mid, void *ev_out);

/* decode: streams -> latent_hat [batch, c3, h3, w3] fp32 (symbol + median) -> g_s -> out [batch, out_h, out_w, d3] fp32 (NHWC).
 *   coder_stream waits for ev_in (may be NULL), decodes (fault flags are OR-ed into *status, which the call does not clear),
 *   records ev_mid; transform_stream waits for it, runs g_s, records ev_out (may be NULL).  tile_counters: 5 int32.
 *   packed == NULL: the caller has filled latent_hat itself; only g_s runs (transform_stream waits for ev_in if given). */
SC2_API int sc2_fp_decode_batch(const sc2_fp_plan *plan, const uint8_t *packed, const int64_t *offsets, float *latent_hat,
                                void *ws_gs, float *out, int32_t *status, int32_t *tile_counters, int coder_layout,
                                sc2_stream_t coder_stream, sc2_stream_t transform_stream, void *ev_in, void *ev_mid, void *ev_out);

/* ------------------------------------------------------------------------------------------
 * Diagnostics: per-CTA trace.  While a (caller-allocated, zeroed) device buffer is installed, every CTA of the coder and
 * tensor-core kernels appends a 32-byte record {u64 start_ns, u64 end_ns, i32 kind, i32 sm, i32 aux, i32 block} after a
 * 16-byte header whose first u32 counts records (kind: 1 conv_tc 2 conv_split 3 conv_first 4 rans_encode 5 rans_decode;
 * aux: tiles the CTA processed).  Not thread-safe against concurrent start/stop; off by default. */
SC2_API int sc2_trace_start(void *device_buffer, int64_t bytes);
SC2_API int sc2_trace_stop(void);

#ifdef __cplusplus
}
#endif
#endif /* SC2B200_H */
