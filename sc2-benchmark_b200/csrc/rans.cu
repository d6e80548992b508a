// rans.cu -- batched CompressAI-compatible rANS coder for sm_100a.
//
// One warp owns one stream (one sample).  A CompressAI stream is a single serial rANS chain
// (64-bit state, 32-bit renormalisation words, 16-bit precision, 4-bit bypass escapes),
// so parallelism is across samples; inside a stream the 32 lanes take everything OFF the chain:
//   encode: lanes look up (reciprocal, bias, freq) for 32 symbols at a time and stage them in shared
//           memory; the chain itself is compare -> mul.hi.u64 -> mad per symbol, no division;
//   decode: each lane holds one CDF entry of the current row, the symbol search is one ballot;
//           renormalisation words are prefetched into a shared-memory ring with cp.async.
// All lanes carry the (uniform) state redundantly, so there is no broadcast on the chain.
//
// Replaces compressai.ans.RansEncoder.encode_with_indexes / RansDecoder.decode_with_indexes as called
// per sample from EntropyModel.compress / decompress (sc2bench/models/layer.py:506,520,647,665).
#include "common.cuh"

namespace sc2 {

constexpr int kWarpsPerBlock = 1;
constexpr uint64_t kRansL = 1ull << 31;

struct TableView {
    const int32_t *sizes;
    const int32_t *offsets;
    const RansEncEntry *enc;
    const int32_t *dec;
    int n_rows, cdf_stride, dec_stride;
};

__device__ __forceinline__ TableView view_tables(const void *blob) {
    const auto *h = reinterpret_cast<const RansTableHeader *>(blob);
    const auto *b = reinterpret_cast<const uint8_t *>(blob);
    TableView t;
    t.n_rows = h->n_rows;
    t.cdf_stride = h->cdf_stride;
    t.dec_stride = h->dec_stride;
    t.sizes = reinterpret_cast<const int32_t *>(b + h->meta_off);
    t.offsets = t.sizes + h->n_rows;
    t.enc = reinterpret_cast<const RansEncEntry *>(b + h->enc_off);
    t.dec = reinterpret_cast<const int32_t *>(b + h->dec_off);
    return t;
}

// ------------------------------------------------------------------------------------------------
// encode
// ------------------------------------------------------------------------------------------------
struct EncState {
    uint64_t x;
    uint32_t *words;  // slot base
    int64_t p;        // next free word is words[p - 1]
    bool overflow;
};

__device__ __forceinline__ void enc_emit(EncState &s, int lane) {
    if (s.p > 0) {
        --s.p;
        if (lane == 0) s.words[s.p] = static_cast<uint32_t>(s.x);
    } else {
        s.overflow = true;
    }
    s.x >>= 32;
}

__device__ __forceinline__ void enc_put_bits(EncState &s, uint32_t val, int lane) {
    // Rans64EncPutBits with nbits = 4: freq = 1 << 12, x_max = (2^15 << 32) << 12 = 2^59
    if (s.x >= (1ull << 59)) enc_emit(s, lane);
    s.x = (s.x << kBypassPrecision) | val;
}

__device__ __forceinline__ void enc_put(EncState &s, const uint4 e, int lane) {
    const uint32_t freq = e.w;
    // x_max = ((L >> 16) << 32) * freq = freq << 47: low 32 bits are zero -> compare high words
    if (static_cast<uint32_t>(s.x >> 32) >= (freq << 15)) enc_emit(s, lane);
    const uint64_t rcp = (static_cast<uint64_t>(e.y) << 32) | e.x;
    const uint32_t shift = (e.z >> 24) & 15u;
    const uint32_t bias = e.z & 0x1ffffu;
    const uint64_t q = __umul64hi(s.x, rcp) >> shift;
    s.x = s.x + bias + q * static_cast<uint64_t>(65536u - freq);
}

template <bool kExplicitIndex>
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
rans_encode_kernel(const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, int batch,
                   int64_t n, int64_t spatial, const void *__restrict__ tables, uint8_t *__restrict__ arena,
                   int64_t slot_bytes, int32_t *__restrict__ lengths, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kWarpsPerBlock + warp;
    const TableView t = view_tables(tables);

    // dynamic smem layout: [staged encoder table (enc_entries x 16 B), if the launcher made room for it]
    //                      [per warp: 32 x uint4 prepared entries][per warp: 32 x u32 raw escape values]
    const int enc_entries = t.n_rows * t.cdf_stride;
    uint32_t dyn_smem;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
    const bool staged = dyn_smem >= static_cast<uint32_t>(enc_entries) * 16u + kWarpsPerBlock * (32 * 16 + 32 * 4);
    uint4 *s_enc = reinterpret_cast<uint4 *>(smem_raw);
    uint4 *s_chunk_base = staged ? s_enc + enc_entries : s_enc;
    uint4 *s_chunk = s_chunk_base + warp * 32;
    uint32_t *s_rawv = reinterpret_cast<uint32_t *>(s_chunk_base + kWarpsPerBlock * 32) + warp * 32;
    if (staged) {
        const uint4 *g = reinterpret_cast<const uint4 *>(t.enc);
        for (int i = threadIdx.x; i < enc_entries; i += blockDim.x) s_enc[i] = g[i];
    }
    __syncthreads();
    if (b >= batch) return;

    const int32_t *sym = symbols + static_cast<int64_t>(b) * n;
    const int32_t *idx = kExplicitIndex ? indexes + static_cast<int64_t>(b) * n : nullptr;
    EncState s;
    s.x = kRansL;
    s.words = reinterpret_cast<uint32_t *>(arena + static_cast<int64_t>(b) * slot_bytes);
    const int64_t slot_words = slot_bytes >> 2;
    s.p = slot_words;
    s.overflow = false;

    for (int64_t hi = n; hi > 0; hi -= 32) {
        const int cnt = hi >= 32 ? 32 : static_cast<int>(hi);
        // ---- off-chain: lane l prepares symbol hi-1-l (descending order = coder order) ----
        if (lane < cnt) {
            const int64_t i = hi - 1 - lane;
            const int32_t v_in = __ldg(sym + i);
            int row = kExplicitIndex ? __ldg(idx + i) : static_cast<int>(i / spatial);
            if (kExplicitIndex && static_cast<uint32_t>(row) >= static_cast<uint32_t>(t.n_rows)) {
                // caller-supplied CDF index out of range (EntropyModel.compress(inputs, indexes) forwards arbitrary tensors):
                // flag it and code the symbol with row 0 instead of reading outside the tables
                atomicOr(status, SC2_FAULT_BAD_INDEX);
                row = 0;
            }
            const int32_t max_value = __ldg(t.sizes + row) - 2;
            int32_t value = v_in - __ldg(t.offsets + row);
            uint32_t raw = 0;
            if (value < 0) {
                raw = static_cast<uint32_t>(-2 * value - 1);
                value = max_value;
            } else if (value >= max_value) {
                raw = static_cast<uint32_t>(2 * (value - max_value));
                value = max_value;
            }
            const int eidx = row * t.cdf_stride + value;
            uint4 e = staged ? s_enc[eidx] : __ldg(reinterpret_cast<const uint4 *>(t.enc) + eidx);
            if (value == max_value) e.z |= 0x80000000u;  // escape marker
            s_chunk[lane] = e;
            s_rawv[lane] = raw;
        }
        __syncwarp();
        // ---- the chain: uniform across lanes ----
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const uint4 e = s_chunk[j];
            if (e.z & 0x80000000u) {
                // entries of an escaped symbol in coder (reverse) order: raw nibbles high..low,
                // then the nibble count (a single digit: raw < 2^32 -> count <= 8 < 15), then the symbol
                const uint32_t raw = s_rawv[j];
                const int n_bypass = raw == 0 ? 0 : (35 - __clz(raw)) >> 2;  // ceil(bits / 4)
                for (int k = n_bypass - 1; k >= 0; --k) enc_put_bits(s, (raw >> (4 * k)) & kMaxBypassVal, lane);
                enc_put_bits(s, static_cast<uint32_t>(n_bypass), lane);
            }
            enc_put(s, e, lane);
        }
        __syncwarp();
    }
    // flush: two words, low then high
    if (s.p >= 2) {
        s.p -= 2;
        if (lane == 0) {
            s.words[s.p] = static_cast<uint32_t>(s.x);
            s.words[s.p + 1] = static_cast<uint32_t>(s.x >> 32);
        }
    } else {
        s.overflow = true;
    }
    if (lane == 0) {
        lengths[b] = s.overflow ? 0 : static_cast<int32_t>((slot_words - s.p) * 4);
        if (s.overflow) atomicOr(status, SC2_FAULT_ARENA_OVERFLOW);
    }
}

// ------------------------------------------------------------------------------------------------
// pack: gather the streams (each at the END of its slot) to the front of one buffer
// ------------------------------------------------------------------------------------------------
__global__ void rans_offsets_kernel(const int32_t *__restrict__ lengths, int batch, int64_t *__restrict__ offsets) {
    // single block exclusive scan; batch is small (<= a few thousand)
    __shared__ int64_t s_carry;
    __shared__ int64_t s_warp[32];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < batch; base += blockDim.x) {
        const int i = base + threadIdx.x;
        int64_t v = i < batch ? lengths[i] : 0;
        int64_t incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t o = __shfl_up_sync(0xffffffffu, incl, d);
            if ((threadIdx.x & 31) >= d) incl += o;
        }
        if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        int64_t warp_off = 0;
        for (int w = 0; w < (threadIdx.x >> 5); ++w) warp_off += s_warp[w];
        const int64_t carry = s_carry;
        if (i < batch) offsets[i] = carry + warp_off + incl - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) s_carry = carry + warp_off + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) offsets[batch] = s_carry;
}

__global__ void rans_pack_kernel(const uint8_t *__restrict__ arena, int64_t slot_bytes,
                                 const int32_t *__restrict__ lengths, const int64_t *__restrict__ offsets,
                                 uint8_t *__restrict__ packed) {
    const int b = blockIdx.y;
    const int32_t len_words = lengths[b] >> 2;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(arena + static_cast<int64_t>(b + 1) * slot_bytes) - len_words;
    uint32_t *dst = reinterpret_cast<uint32_t *>(packed + offsets[b]);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < len_words; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

// ------------------------------------------------------------------------------------------------
// decode
// ------------------------------------------------------------------------------------------------
constexpr int kRing = 64;  // words per warp, refilled by halves

struct DecState {
    uint64_t x;
    int64_t p;  // index of the next unread word
    int64_t n_words;
    const uint32_t *words;
    uint32_t *ring;
    bool truncated;
};

__device__ __forceinline__ void ring_fill_half(const DecState &s, int64_t first, int lane) {
    // loads words [first, first + 32) into their ring slots; zero-fills past the end
    const int64_t w = first + lane;
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(s.ring + (w & (kRing - 1))));
    const bool ok = w < s.n_words;
    const uint32_t *src = ok ? s.words + w : s.words;
    const int src_bytes = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}

__device__ __forceinline__ uint32_t dec_next_word(DecState &s, int lane) {
    if (s.p >= s.n_words) s.truncated = true;
    const uint32_t w = s.ring[s.p & (kRing - 1)];
    ++s.p;
    if ((s.p & 31) == 0) {
        // the half [p - 32, p) is consumed: make sure the previous refill has landed, then reuse it
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncwarp();
        ring_fill_half(s, s.p + 32, lane);
    }
    return w;
}

__device__ __forceinline__ uint32_t dec_get_bits(DecState &s, int lane) {
    const uint32_t val = static_cast<uint32_t>(s.x) & kMaxBypassVal;
    s.x >>= kBypassPrecision;
    if (s.x < kRansL) s.x = (s.x << 32) | dec_next_word(s, lane);
    return val;
}

template <bool kExplicitIndex>
__global__ void __launch_bounds__(32 * kWarpsPerBlock)
rans_decode_kernel(const uint8_t *__restrict__ packed, const int64_t *__restrict__ offsets, int batch, int64_t n,
                   const int32_t *__restrict__ indexes, int64_t spatial, const void *__restrict__ tables,
                   int32_t *__restrict__ out_symbols, float *__restrict__ out_values,
                   const float *__restrict__ means, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int b = blockIdx.x * kWarpsPerBlock + warp;
    const TableView t = view_tables(tables);

    // layout: [per-warp ring][staged decoder CDF rows if they fit]
    uint32_t *s_ring = reinterpret_cast<uint32_t *>(smem_raw) + warp * kRing;
    int32_t *s_dec = reinterpret_cast<int32_t *>(smem_raw) + kWarpsPerBlock * kRing;
    uint32_t dyn_smem;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
    const int dec_entries = t.n_rows * t.dec_stride;
    const bool staged = dyn_smem >= static_cast<uint32_t>(dec_entries + kWarpsPerBlock * kRing) * 4u;
    if (staged)
        for (int i = threadIdx.x; i < dec_entries; i += blockDim.x) s_dec[i] = t.dec[i];
    __syncthreads();
    if (b >= batch) return;
    const int32_t *dec = staged ? s_dec : t.dec;

    const int64_t off = offsets[b];
    const int64_t n_bytes = offsets[b + 1] - off;
    DecState s;
    s.words = reinterpret_cast<const uint32_t *>(packed + off);
    s.n_words = n_bytes >> 2;
    s.ring = s_ring;
    s.truncated = false;
    if (n_bytes < 8 || (n_bytes & 3) || (off & 3)) {
        if (lane == 0) atomicOr(status, SC2_FAULT_BAD_STREAM);
        return;
    }
    ring_fill_half(s, 0, lane);
    ring_fill_half(s, 32, lane);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    s.x = static_cast<uint64_t>(s.ring[0]) | (static_cast<uint64_t>(s.ring[1]) << 32);
    s.p = 2;

    const int32_t *idx = kExplicitIndex ? indexes + static_cast<int64_t>(b) * n : nullptr;
    int32_t *osym = out_symbols ? out_symbols + static_cast<int64_t>(b) * n : nullptr;
    float *oval = out_values ? out_values + static_cast<int64_t>(b) * n : nullptr;

    if (kExplicitIndex) {
        // ---- per-element rows (GaussianConditional): every symbol has its own CDF row, so there is no "previous symbol of the row"
        // to speculate on and no row slice worth keeping in registers.  Off the chain, lane l prepares symbol base + l: its row, the
        // row's CENTRE symbol c (the mode of the quantised Gaussian: value 0 around the mean) with (start, freq) of c.  The chain then
        // tests cum - start < freq first; only a symbol away from the centre pays for a search, which walks outwards from c in
        // 32-entry ballots (one or two for a Gaussian) instead of scanning the row from its first entry.
        for (int64_t base = 0; base < n; base += 32) {
            const int cnt = (n - base) >= 32 ? 32 : static_cast<int>(n - base);
            int my_row = lane < cnt ? __ldg(idx + base + lane) : 0;
            if (static_cast<uint32_t>(my_row) >= static_cast<uint32_t>(t.n_rows)) {
                atomicOr(status, SC2_FAULT_BAD_INDEX);
                my_row = 0;
            }
            const int32_t *my_drow = dec + static_cast<int64_t>(my_row) * t.dec_stride;
            const int32_t my_size = __ldg(t.sizes + my_row);
            const int32_t my_max = my_size - 2, my_off = __ldg(t.offsets + my_row);
            const int32_t my_c = my_max > 0 ? (my_max - 1) >> 1 : 0;   // pmf_length = my_max + 1 entries: centre (pmf_length - 1) / 2 ... of the regular symbols
            const uint32_t my_start = static_cast<uint32_t>(my_drow[my_c]);
            uint32_t my_freq = static_cast<uint32_t>(my_drow[my_c + 1]) - my_start;
            if (my_c >= my_max) my_freq = 0u;  // (the centre would be the escape symbol: always take the search path)
            const float my_mean = means ? __ldg(means + my_row) : 0.0f;
            int32_t my_value = 0;
            for (int j = 0; j < cnt; ++j) {
                const uint32_t start = __shfl_sync(0xffffffffu, my_start, j), freq = __shfl_sync(0xffffffffu, my_freq, j);
                const int32_t c = __shfl_sync(0xffffffffu, my_c, j);
                const uint32_t cum = static_cast<uint32_t>(s.x) & 0xffffu;
                int32_t value;
                if (cum - start < freq) {  // (uniform) the centre symbol
                    s.x = static_cast<uint64_t>(freq) * (s.x >> kRansPrecision) + (cum - start);
                    if (s.x < kRansL) s.x = (s.x << 32) | dec_next_word(s, lane);
                    value = c;
                } else {
                    const int r = __shfl_sync(0xffffffffu, my_row, j);
                    const int32_t max_value = __shfl_sync(0xffffffffu, my_max, j);
                    const int32_t *drow = dec + static_cast<int64_t>(r) * t.dec_stride;
                    int32_t sidx;  // largest entry with cdf[sidx] <= cum
                    if (cum < start) {  // to the left of the centre: entries c - 32 .. c - 1, then further left
                        for (int32_t hi = c;; hi -= 32) {
                            const int32_t e = hi - 32 + lane;
                            const uint32_t mm = __ballot_sync(0xffffffffu, e >= 0 && static_cast<uint32_t>(drow[e < 0 ? 0 : e]) <= cum);
                            if (mm) {
                                sidx = hi - 32 + (31 - __clz(mm));
                                break;
                            }
                        }
                    } else {            // to the right: first entry k > c with cdf[k] > cum (rows are sentinel-padded beyond their size)
                        for (int32_t lo = c + 1;; lo += 32) {
                            const int32_t e = lo + lane < t.dec_stride ? lo + lane : t.dec_stride - 1;  // (clamped: the pad entries compare true)
                            const uint32_t mm = __ballot_sync(0xffffffffu, static_cast<uint32_t>(drow[e]) > cum);
                            if (mm) {
                                sidx = lo + __ffs(mm) - 2;
                                break;
                            }
                        }
                    }
                    const uint32_t st2 = static_cast<uint32_t>(drow[sidx]);
                    const uint32_t fr2 = static_cast<uint32_t>(drow[sidx + 1]) - st2;
                    s.x = static_cast<uint64_t>(fr2) * (s.x >> kRansPrecision) + (cum - st2);
                    if (s.x < kRansL) s.x = (s.x << 32) | dec_next_word(s, lane);
                    value = sidx;
                    if (value == max_value) {
                        uint32_t val = dec_get_bits(s, lane);
                        int n_bypass = static_cast<int>(val);
                        while (val == kMaxBypassVal) {
                            val = dec_get_bits(s, lane);
                            n_bypass += static_cast<int>(val);
                        }
                        uint32_t raw = 0;
                        for (int q = 0; q < n_bypass; ++q) {
                            const uint32_t nib = dec_get_bits(s, lane);
                            if (q < 8) raw |= nib << (4 * q);
                        }
                        value = static_cast<int32_t>(raw >> 1);
                        value = (raw & 1u) ? -value - 1 : value + max_value;
                    }
                }
                if (j == lane) my_value = value + my_off;
            }
            if (lane < cnt) {
                if (osym) osym[base + lane] = my_value;
                if (oval) oval[base + lane] = static_cast<float>(my_value) + my_mean;
            }
        }
        if (s.truncated && lane == 0) atomicOr(status, SC2_FAULT_STREAM_TRUNCATED);
        return;
    }

    int row = -1;
    int32_t c0 = 0, c1 = 0;  // this lane's entries [lane] and [lane + 32] of the current CDF row
    int32_t max_value = 0, offset = 0, row_size = 0;
    float mean = 0.0f;
    const int32_t *drow = dec;

    for (int64_t base = 0; base < n; base += 32) {
        const int cnt = (n - base) >= 32 ? 32 : static_cast<int>(n - base);
        int my_row = 0;
        if (kExplicitIndex) my_row = lane < cnt ? __ldg(idx + base + lane) : 0;
        int32_t my_value = 0;
        float my_mean = 0.0f;
        for (int j = 0; j < cnt; ++j) {
            int r = kExplicitIndex ? __shfl_sync(0xffffffffu, my_row, j) : static_cast<int>((base + j) / spatial);
            if (kExplicitIndex && static_cast<uint32_t>(r) >= static_cast<uint32_t>(t.n_rows)) {  // (uniform) see the encoder
                if (lane == 0) atomicOr(status, SC2_FAULT_BAD_INDEX);
                r = 0;
            }
            if (r != row) {  // uniform branch: (re)load this lane's slice of the row
                row = r;
                drow = dec + static_cast<int64_t>(row) * t.dec_stride;
                c0 = drow[lane];
                c1 = t.dec_stride > 32 ? drow[lane + 32] : 0x7fffffff;
                row_size = __ldg(t.sizes + row);
                max_value = row_size - 2;
                offset = __ldg(t.offsets + row);
                mean = means ? __ldg(means + row) : 0.0f;
            }
            // ---- the chain ----
            const int32_t cum = static_cast<int32_t>(static_cast<uint32_t>(s.x) & 0xffffu);
            int k;  // first entry with cdf[k] > cum
            uint32_t m = __ballot_sync(0xffffffffu, c0 > cum);
            if (m) {
                k = __ffs(m) - 1;
            } else {
                m = __ballot_sync(0xffffffffu, c1 > cum);
                if (m) {
                    k = 32 + __ffs(m) - 1;
                } else {
                    k = 64;
                    for (;; k += 32) {
                        const uint32_t mm = __ballot_sync(0xffffffffu, drow[k + lane] > cum);
                        if (mm) {
                            k += __ffs(mm) - 1;
                            break;
                        }
                    }
                }
            }
            const int32_t sidx = k - 1;
            const uint32_t start = static_cast<uint32_t>(drow[sidx]);
            const uint32_t freq = static_cast<uint32_t>(drow[k]) - start;
            s.x = static_cast<uint64_t>(freq) * (s.x >> kRansPrecision) + static_cast<uint32_t>(cum) - start;
            if (s.x < kRansL) s.x = (s.x << 32) | dec_next_word(s, lane);
            int32_t value = sidx;
            if (value == max_value) {
                uint32_t val = dec_get_bits(s, lane);
                int n_bypass = static_cast<int>(val);
                while (val == kMaxBypassVal) {
                    val = dec_get_bits(s, lane);
                    n_bypass += static_cast<int>(val);
                }
                uint32_t raw = 0;
                for (int q = 0; q < n_bypass; ++q) {
                    const uint32_t nib = dec_get_bits(s, lane);
                    if (q < 8) raw |= nib << (4 * q);
                }
                value = static_cast<int32_t>(raw >> 1);
                value = (raw & 1u) ? -value - 1 : value + max_value;
            }
            value += offset;
            if (j == lane) {
                my_value = value;
                my_mean = mean;
            }
        }
        if (lane < cnt) {
            if (osym) osym[base + lane] = my_value;
            if (oval) oval[base + lane] = static_cast<float>(my_value) + my_mean;
        }
    }
    if (s.truncated && lane == 0) atomicOr(status, SC2_FAULT_STREAM_TRUNCATED);
}

// ------------------------------------------------------------------------------------------------
// elementwise helpers
// ------------------------------------------------------------------------------------------------
__global__ void quantize_symbols_kernel(const float *__restrict__ x, const float *__restrict__ means,
                                        int32_t *__restrict__ symbols, int channels, int64_t spatial, int64_t total) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>((i / spatial) % channels);
        const float m = means ? __ldg(means + c) : 0.0f;
        symbols[i] = __float2int_rn(rintf(x[i] - m));  // torch.round = half to even, then .int()
    }
}

__global__ void dequantize_kernel(const int32_t *__restrict__ symbols, const float *__restrict__ means,
                                  float *__restrict__ out, int64_t n) {
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        out[i] = static_cast<float>(symbols[i]) + (means ? means[i] : 0.0f);
}

__global__ void gc_build_indexes_kernel(const float *__restrict__ scales, int64_t n,
                                        const float *__restrict__ table, int n_levels, float bound,
                                        int32_t *__restrict__ indexes) {
    extern __shared__ float s_table[];
    for (int i = threadIdx.x; i < n_levels; i += blockDim.x) s_table[i] = table[i];
    __syncthreads();
    for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const float s = fmaxf(scales[i], bound);
        // idx = (n_levels - 1) - #{k < n_levels - 1 : s <= table[k]}  (table ascending -> binary search)
        int lo = 0, hi = n_levels - 1;  // first k in [0, n_levels-1) with s <= table[k]
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (s <= s_table[mid]) hi = mid; else lo = mid + 1;
        }
        indexes[i] = lo;
    }
}

}  // namespace sc2

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int sc2_rans_encode_batch(const int32_t *symbols, const int32_t *indexes, int batch, int64_t n_per_stream,
                          int64_t spatial, const void *tables, int n_rows, int cdf_stride, uint8_t *arena,
                          int64_t slot_bytes, int32_t *lengths, int32_t *status, int layout, sc2_stream_t stream) {
    if (batch == 0) return SC2_OK;
    if ((!symbols && n_per_stream > 0) || !tables || !arena || !lengths || !status) return SC2_ERR_INVALID_ARG;
    if (batch < 0 || n_per_stream < 0 || n_per_stream > 0x7fffffff || slot_bytes < 8 || (slot_bytes & 3)) return SC2_ERR_INVALID_ARG;
    if (!indexes && spatial < 1) return SC2_ERR_INVALID_ARG;
    if (batch == 0) return SC2_OK;
    cudaStream_t st = sc2::as_stream(stream);
    if (n_rows < 1 || cdf_stride < 2) return SC2_ERR_INVALID_ARG;
    if (!indexes && (n_per_stream + spatial - 1) / spatial > n_rows) return SC2_ERR_INVALID_ARG;
    if (layout < SC2_RANS_AUTO || layout > SC2_RANS_LANE_PER_STREAM) return SC2_ERR_INVALID_ARG;
    if (!indexes && spatial <= 0x7fffffff && sc2::rans_use_lanes(layout))
        return sc2::launch_rans_encode_lanes(symbols, batch, n_per_stream, spatial, tables, n_rows, cdf_stride, arena, slot_bytes,
                                             lengths, status, st);
    if (indexes || spatial <= 0x7fffffff)  // (the latency-tuned chain, for channel rows and for per-element rows alike)
        return sc2::launch_rans_encode_fast(symbols, indexes, batch, n_per_stream, spatial, tables, n_rows, cdf_stride, arena, slot_bytes,
                                            lengths, status, st);
    const size_t chunk = sc2::kWarpsPerBlock * (32 * 16 + 32 * 4);
    size_t smem = chunk;
    const size_t table_bytes = static_cast<size_t>(n_rows) * cdf_stride * 16;
    if (table_bytes + chunk <= 96 * 1024) smem += table_bytes;
    const int grid = (batch + sc2::kWarpsPerBlock - 1) / sc2::kWarpsPerBlock;
    if (indexes) {
        SC2_CUDA_TRY(cudaFuncSetAttribute(sc2::rans_encode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        sc2::rans_encode_kernel<true><<<grid, 32 * sc2::kWarpsPerBlock, smem, st>>>(
            symbols, indexes, batch, n_per_stream, spatial, tables, arena, slot_bytes, lengths, status);
    } else {
        SC2_CUDA_TRY(cudaFuncSetAttribute(sc2::rans_encode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        sc2::rans_encode_kernel<false><<<grid, 32 * sc2::kWarpsPerBlock, smem, st>>>(
            symbols, indexes, batch, n_per_stream, spatial, tables, arena, slot_bytes, lengths, status);
    }
    SC2_LAUNCH_CHECK("rans_encode_kernel");
    return SC2_OK;
}

int sc2_rans_pack(const uint8_t *arena, int64_t slot_bytes, const int32_t *lengths, int batch, uint8_t *packed,
                  int64_t *offsets, sc2_stream_t stream) {
    if (!arena || !lengths || !packed || !offsets || batch < 0 || (slot_bytes & 3)) return SC2_ERR_INVALID_ARG;
    cudaStream_t st = sc2::as_stream(stream);
    sc2::rans_offsets_kernel<<<1, 256, 0, st>>>(lengths, batch, offsets);
    SC2_LAUNCH_CHECK("rans_offsets_kernel");
    if (batch == 0) return SC2_OK;
    const int64_t words = slot_bytes >> 2;
    int gx = static_cast<int>((words + 256 * 8 - 1) / (256 * 8));
    if (gx < 1) gx = 1;
    if (gx > 64) gx = 64;
    sc2::rans_pack_kernel<<<dim3(gx, batch), 256, 0, st>>>(arena, slot_bytes, lengths, offsets, packed);
    SC2_LAUNCH_CHECK("rans_pack_kernel");
    return SC2_OK;
}

int sc2_rans_decode_batch(const uint8_t *packed, const int64_t *offsets, int batch, int64_t n_per_stream,
                          const int32_t *indexes, int64_t spatial, const void *tables, int n_rows, int cdf_stride,
                          int32_t *out_symbols, float *out_values, const float *means, int32_t *status, int layout,
                          sc2_stream_t stream) {
    if (!packed || !offsets || !tables || !status || (!out_symbols && !out_values)) return SC2_ERR_INVALID_ARG;
    if (batch < 0 || n_per_stream < 0 || n_per_stream > 0x7fffffff) return SC2_ERR_INVALID_ARG;
    if (!indexes && spatial < 1) return SC2_ERR_INVALID_ARG;
    if (batch == 0 || n_per_stream == 0) return SC2_OK;
    cudaStream_t st = sc2::as_stream(stream);
    if (n_rows < 1 || cdf_stride < 2) return SC2_ERR_INVALID_ARG;
    if (!indexes && (n_per_stream + spatial - 1) / spatial > n_rows) return SC2_ERR_INVALID_ARG;
    if (layout < SC2_RANS_AUTO || layout > SC2_RANS_LANE_PER_STREAM) return SC2_ERR_INVALID_ARG;
    if (!indexes && spatial <= 0x7fffffff && sc2::rans_use_lanes(layout))
        return sc2::launch_rans_decode_lanes(packed, offsets, batch, n_per_stream, spatial, tables, out_symbols, out_values, means,
                                             status, st);
    if (!indexes && spatial <= 0x7fffffff)
        return sc2::launch_rans_decode_fast(packed, offsets, batch, n_per_stream, spatial, tables, n_rows, cdf_stride, out_symbols,
                                            out_values, means, status, st);
    size_t smem = sc2::kWarpsPerBlock * sc2::kRing * 4;
    const size_t table_bytes = static_cast<size_t>(n_rows) * ((cdf_stride + 31) / 32 * 32) * 4;
    if (table_bytes + smem <= 96 * 1024) smem += table_bytes;
    const int grid = (batch + sc2::kWarpsPerBlock - 1) / sc2::kWarpsPerBlock;
    if (indexes) {
        SC2_CUDA_TRY(cudaFuncSetAttribute(sc2::rans_decode_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        sc2::rans_decode_kernel<true><<<grid, 32 * sc2::kWarpsPerBlock, smem, st>>>(
            packed, offsets, batch, n_per_stream, indexes, spatial, tables, out_symbols, out_values, means, status);
    } else {
        SC2_CUDA_TRY(cudaFuncSetAttribute(sc2::rans_decode_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        sc2::rans_decode_kernel<false><<<grid, 32 * sc2::kWarpsPerBlock, smem, st>>>(
            packed, offsets, batch, n_per_stream, indexes, spatial, tables, out_symbols, out_values, means, status);
    }
    SC2_LAUNCH_CHECK("rans_decode_kernel");
    return SC2_OK;
}

int sc2_quantize_symbols(const float *x, const float *means, int32_t *symbols, int batch, int channels,
                         int64_t spatial, sc2_stream_t stream) {
    if (!x || !symbols || batch < 0 || channels < 1 || spatial < 0) return SC2_ERR_INVALID_ARG;
    const int64_t total = static_cast<int64_t>(batch) * channels * spatial;
    if (total == 0) return SC2_OK;
    int64_t blocks = (total + 255) / 256;
    if (blocks > sc2::kNumSMs * 16) blocks = sc2::kNumSMs * 16;
    sc2::quantize_symbols_kernel<<<static_cast<int>(blocks), 256, 0, sc2::as_stream(stream)>>>(x, means, symbols, channels, spatial, total);
    SC2_LAUNCH_CHECK("quantize_symbols_kernel");
    return SC2_OK;
}

int sc2_dequantize(const int32_t *symbols, const float *means, float *out, int64_t n, sc2_stream_t stream) {
    if (!symbols || !out || n < 0) return SC2_ERR_INVALID_ARG;
    if (n == 0) return SC2_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > sc2::kNumSMs * 16) blocks = sc2::kNumSMs * 16;
    sc2::dequantize_kernel<<<static_cast<int>(blocks), 256, 0, sc2::as_stream(stream)>>>(symbols, means, out, n);
    SC2_LAUNCH_CHECK("dequantize_kernel");
    return SC2_OK;
}

int sc2_gc_build_indexes(const float *scales, int64_t n, const float *scale_table, int n_levels, float scale_bound,
                         int32_t *indexes, sc2_stream_t stream) {
    if (!scales || !scale_table || !indexes || n < 0 || n_levels < 1 || n_levels > 4096) return SC2_ERR_INVALID_ARG;
    if (n == 0) return SC2_OK;
    int64_t blocks = (n + 255) / 256;
    if (blocks > sc2::kNumSMs * 16) blocks = sc2::kNumSMs * 16;
    sc2::gc_build_indexes_kernel<<<static_cast<int>(blocks), 256, n_levels * sizeof(float), sc2::as_stream(stream)>>>(
        scales, n, scale_table, n_levels, scale_bound, indexes);
    SC2_LAUNCH_CHECK("gc_build_indexes_kernel");
    return SC2_OK;
}

}  // extern "C"
