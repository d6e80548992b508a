// rans_fast.cu -- latency-tuned channel-mode rANS kernels (EntropyBottleneck: CDF row = channel = i / spatial).
//
// A CompressAI stream is one serial chain, so per-stream speed = length of the dependent instruction sequence per
// symbol.  These kernels keep that sequence minimal (everything else is issued off the chain or hoisted per 32 symbols):
//
//   encode  chain = ISETP (renorm?) -> SEL -> mul.hi.u64 (exact reciprocal division) -> SHF -> IMAD.WIDE.
//           Renormalisation is branch-free (predicated store + selects); escapes leave through a cold call; the next
//           chunk's symbols are requested from HBM before the current chunk's chain starts.
//   decode  the symbol search is speculated: the (start, freq) of the previously decoded symbol of the row is kept in
//           registers and `cum - start < freq` is tested first (2 dependent ops).  Latents are sparse (runs of the
//           same symbol), so the ballot search + table lookup is only paid on a change of symbol.
//           chain (hit) = LOP -> IADD -> ISETP -> IMAD.WIDE -> IMAD -> ISETP -> SEL.
//           Renormalisation words sit in a cp.async-fed shared-memory ring and the next word is always already in a
//           register; decoded symbols go through shared memory so that global stores are coalesced 128-byte lines.
//
// Bit-exactness: identical arithmetic to rans.cu / the oracle; tests/test_gpu_parity.py compares bytes.
#include "common.cuh"

namespace sc2 {

namespace {

constexpr uint64_t kL = 1ull << 31;
constexpr int kRingWords = 64;

struct Tables {
    const int32_t *sizes, *offsets, *dec;
    const uint4 *enc;
    int n_rows, cdf_stride, dec_stride;
};

__device__ __forceinline__ Tables view(const void *blob) {
    const auto *h = reinterpret_cast<const RansTableHeader *>(blob);
    const auto *b = reinterpret_cast<const uint8_t *>(blob);
    Tables t;
    t.n_rows = h->n_rows;
    t.cdf_stride = h->cdf_stride;
    t.dec_stride = h->dec_stride;
    t.sizes = reinterpret_cast<const int32_t *>(b + h->meta_off);
    t.offsets = t.sizes + h->n_rows;
    t.enc = reinterpret_cast<const uint4 *>(b + h->enc_off);
    t.dec = reinterpret_cast<const int32_t *>(b + h->dec_off);
    return t;
}

// ---------------------------------------------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------------------------------------------
struct EncChain {
    uint64_t x;
    uint32_t *words;  // slot base (u32 words)
    uint32_t pw;      // words[pw - 1] is the next free word
    uint32_t overflow;
};

__device__ __forceinline__ void enc_emit_checked(EncChain &s) {
    if (s.pw != 0u) {
        --s.pw;
        s.words[s.pw] = static_cast<uint32_t>(s.x);  // every lane writes the same word to the same address
    } else {
        s.overflow = 1u;
    }
    s.x >>= 32;
}

// cold path: nibbles of an escaped symbol in coder order (high nibble first, then the nibble count).
// State goes in and out BY VALUE so that the hot loop keeps it in registers (a by-reference call would pin it to the stack).
__device__ __noinline__ EncChain enc_escape(EncChain s, uint32_t raw) {
    const int n_bypass = raw == 0 ? 0 : (35 - __clz(raw)) >> 2;
    for (int k = n_bypass - 1; k >= -1; --k) {
        const uint32_t val = k >= 0 ? ((raw >> (4 * k)) & kMaxBypassVal) : static_cast<uint32_t>(n_bypass);
        if (s.x >= (1ull << 59)) enc_emit_checked(s);
        s.x = (s.x << kBypassPrecision) | val;
    }
    return s;
}

__global__ void __launch_bounds__(32)
rans_encode_fast_kernel(const int32_t *__restrict__ symbols, int batch, uint32_t n, uint32_t spatial,
                        const void *__restrict__ tables, uint8_t *__restrict__ arena, int64_t slot_bytes,
                        int32_t *__restrict__ lengths, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x;
    const Tables t = view(tables);
    // smem: [32 x uint4 A][32 x uint4 B][staged encoder table, if the launcher made room]
    uint4 *sA = reinterpret_cast<uint4 *>(smem_raw);
    uint4 *sB = sA + 32;
    uint4 *s_enc = sB + 32;
    const int enc_entries = t.n_rows * t.cdf_stride;
    uint32_t dyn_smem;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
    const bool staged = dyn_smem >= static_cast<uint32_t>(enc_entries + 64) * 16u;
    if (staged)
        for (int i = lane; i < enc_entries; i += 32) s_enc[i] = __ldg(t.enc + i);
    __syncwarp();

    const int32_t *sym = symbols + static_cast<int64_t>(b) * n;
    EncChain s;
    s.x = kL;
    s.words = reinterpret_cast<uint32_t *>(arena + static_cast<int64_t>(b) * slot_bytes);
    const uint32_t slot_words = static_cast<uint32_t>(slot_bytes >> 2);
    s.pw = slot_words;
    s.overflow = 0u;

    // symbols are consumed back to front; lane l of a chunk ending at `hi` owns symbol hi - 1 - l
    int32_t nxt = (lane < static_cast<int>(n)) ? __ldg(sym + (n - 1 - lane)) : 0;
    for (uint32_t hi = n; hi > 0; hi = hi > 32 ? hi - 32 : 0) {
        const int cnt = hi >= 32 ? 32 : static_cast<int>(hi);
        const int32_t cur = nxt;
        if (hi > 32) {  // request the next chunk now: its HBM/L2 latency hides behind this chunk's chain
            const uint32_t hn = hi - 32;
            nxt = (static_cast<uint32_t>(lane) < hn) ? __ldg(sym + (hn - 1 - lane)) : 0;
        }
        if (lane < cnt) {
            const uint32_t i = hi - 1 - lane;
            const int row = static_cast<int>(i / spatial);
            const int32_t max_value = __ldg(t.sizes + row) - 2;
            int32_t value = cur - __ldg(t.offsets + row);
            uint32_t raw = 0, esc = 0;
            if (value < 0) {
                raw = static_cast<uint32_t>(-2 * value - 1);
                value = max_value;
            } else if (value >= max_value) {
                raw = static_cast<uint32_t>(2 * (value - max_value));
                value = max_value;
            }
            if (value == max_value) esc = 1;
            const int eidx = row * t.cdf_stride + value;
            const uint4 e = staged ? s_enc[eidx] : __ldg(t.enc + eidx);
            // unpack here, in parallel across lanes, so the chain only sees ready-to-use operands
            sA[lane] = make_uint4(e.x, e.y, e.w << 15, 65536u - e.w);               // rcp_lo, rcp_hi, renorm threshold, 2^16 - freq
            sB[lane] = make_uint4(e.z & 0x1ffffu, (e.z >> 24) & 15u, esc, raw);     // bias, shift, escape?, raw
        }
        __syncwarp();
        uint4 A = sA[0], B = sB[0];
#pragma unroll 2
        for (int j = 0; j < cnt; ++j) {
            const uint4 a = A, bb = B;
            if (j + 1 < cnt) {  // software prefetch of the next entry
                A = sA[j + 1];
                B = sB[j + 1];
            }
            if (__builtin_expect(bb.z != 0u, 0)) s = enc_escape(s, bb.w);
            // ---- the chain ----
            const uint32_t xl = static_cast<uint32_t>(s.x), xh = static_cast<uint32_t>(s.x >> 32);
            const bool ren = xh >= a.z;                 // x >= freq << 47
            const bool can = s.pw != 0u;
            if (ren && can) s.words[s.pw - 1] = xl;
            s.pw -= (ren && can) ? 1u : 0u;
            s.overflow |= (ren && !can) ? 1u : 0u;
            const uint64_t y = ren ? static_cast<uint64_t>(xh) : s.x;
            const uint64_t rcp = (static_cast<uint64_t>(a.y) << 32) | a.x;
            const uint64_t q = __umul64hi(y, rcp) >> bb.y;
            s.x = y + bb.x + q * static_cast<uint64_t>(a.w);
        }
        __syncwarp();
    }
    if (s.pw >= 2u) {
        s.pw -= 2;
        if (lane == 0) {
            s.words[s.pw] = static_cast<uint32_t>(s.x);
            s.words[s.pw + 1] = static_cast<uint32_t>(s.x >> 32);
        }
    } else {
        s.overflow = 1u;
    }
    if (lane == 0) {
        lengths[b] = s.overflow ? 0 : static_cast<int32_t>((slot_words - s.pw) * 4u);
        if (s.overflow) atomicOr(status, SC2_FAULT_ARENA_OVERFLOW);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------------------------
struct DecChain {
    uint64_t x;
    uint32_t p;        // index of the word held in `next_w`
    uint32_t next_w;   // words[p], already in a register
    uint32_t n_words;
    const uint32_t *words;
    uint32_t *ring;
    uint32_t truncated;
};

__device__ __forceinline__ void ring_fill(const DecChain &s, uint32_t first, int lane) {
    const uint32_t w = first + lane;
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(s.ring + (w & (kRingWords - 1))));
    const bool ok = w < s.n_words;
    const uint32_t *src = ok ? s.words + w : s.words;
    const int src_bytes = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// cold path: ring refill when the word index crosses a half
__device__ __noinline__ void dec_refill(const uint32_t *words, uint32_t n_words, uint32_t *ring, uint32_t first, int lane) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    DecChain t;
    t.words = words;
    t.n_words = n_words;
    t.ring = ring;
    ring_fill(t, first, lane);
}

// consume next_w and fetch the following word from the ring (refilling the ring half that was just left)
__device__ __forceinline__ void dec_advance_word(DecChain &s, int lane) {
    s.truncated |= (s.p >= s.n_words) ? 1u : 0u;
    ++s.p;
    if (__builtin_expect((s.p & 31u) == 0u, 0)) dec_refill(s.words, s.n_words, s.ring, s.p + 32u, lane);
    s.next_w = s.ring[s.p & (kRingWords - 1)];
}

__device__ __forceinline__ uint32_t dec_get_nibble(DecChain &s, int lane) {
    const uint32_t val = static_cast<uint32_t>(s.x) & kMaxBypassVal;
    s.x >>= kBypassPrecision;
    if (s.x < kL) {
        s.x = (s.x << 32) | s.next_w;
        dec_advance_word(s, lane);
    }
    return val;
}

struct DecEscapeResult {
    DecChain s;
    int32_t value;
};

// cold path: the bypass-coded magnitude of an escaped symbol (state by value: see enc_escape)
__device__ __noinline__ DecEscapeResult dec_escape(DecChain s, int32_t max_value, int lane) {
    uint32_t val = dec_get_nibble(s, lane);
    int n_bypass = static_cast<int>(val);
    while (val == kMaxBypassVal) {
        val = dec_get_nibble(s, lane);
        n_bypass += static_cast<int>(val);
    }
    uint32_t raw = 0;
    for (int q = 0; q < n_bypass; ++q) {
        const uint32_t nib = dec_get_nibble(s, lane);
        if (q < 8) raw |= nib << (4 * q);
    }
    const int32_t v = static_cast<int32_t>(raw >> 1);
    DecEscapeResult r;
    r.s = s;
    r.value = (raw & 1u) ? -v - 1 : v + max_value;
    return r;
}

__global__ void __launch_bounds__(32)
rans_decode_fast_kernel(const uint8_t *__restrict__ packed, const int64_t *__restrict__ offsets, int batch, uint32_t n,
                        uint32_t spatial, const void *__restrict__ tables, int32_t *__restrict__ out_symbols,
                        float *__restrict__ out_values, const float *__restrict__ means, int32_t *__restrict__ status) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x;
    const int b = blockIdx.x;
    const Tables t = view(tables);
    // smem: [ring 64 words][out staging 32 words][staged CDF rows, if the launcher made room]
    uint32_t *s_ring = reinterpret_cast<uint32_t *>(smem_raw);
    int32_t *s_out = reinterpret_cast<int32_t *>(s_ring + kRingWords);
    int32_t *s_dec = s_out + 32;
    uint32_t dyn_smem;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
    const int dec_entries = t.n_rows * t.dec_stride;
    const bool staged = dyn_smem >= static_cast<uint32_t>(dec_entries + kRingWords + 32) * 4u;
    if (staged)
        for (int i = lane; i < dec_entries; i += 32) s_dec[i] = __ldg(t.dec + i);
    __syncwarp();
    const int32_t *dec = staged ? s_dec : t.dec;

    const int64_t off = offsets[b];
    const int64_t n_bytes = offsets[b + 1] - off;
    if (n_bytes < 8 || (n_bytes & 3) || (off & 3)) {
        if (lane == 0) atomicOr(status, SC2_FAULT_BAD_STREAM);
        return;
    }
    DecChain s;
    s.words = reinterpret_cast<const uint32_t *>(packed + off);
    s.n_words = static_cast<uint32_t>(n_bytes >> 2);
    s.ring = s_ring;
    s.truncated = 0u;
    ring_fill(s, 0, lane);
    ring_fill(s, 32, lane);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    s.x = static_cast<uint64_t>(s.ring[0]) | (static_cast<uint64_t>(s.ring[1]) << 32);
    s.p = 2;
    s.next_w = s.ring[2];

    int32_t *osym = out_symbols ? out_symbols + static_cast<int64_t>(b) * n : nullptr;
    float *oval = out_values ? out_values + static_cast<int64_t>(b) * n : nullptr;

    uint32_t done = 0;
    for (int row = 0; done < n; ++row) {
        const uint32_t row_n = (n - done) < spatial ? (n - done) : spatial;
        const int32_t *drow = dec + static_cast<int64_t>(row) * t.dec_stride;
        const int32_t c0 = drow[lane];
        const int32_t c1 = t.dec_stride > 32 ? drow[lane + 32] : 0x7fffffff;
        const int32_t max_value = __ldg(t.sizes + row) - 2;
        const int32_t offset = __ldg(t.offsets + row);
        const float mean = means ? __ldg(means + row) : 0.0f;
        // speculation state: the last regular symbol decoded in this row
        uint32_t sp_start = 0, sp_freq = 0;
        int32_t sp_value = 0;
        for (uint32_t base = 0; base < row_n; base += 32) {
            const int cnt = (row_n - base) >= 32 ? 32 : static_cast<int>(row_n - base);
#pragma unroll 4
            for (int j = 0; j < cnt; ++j) {
                // ---- the chain ----
                const uint32_t cum = static_cast<uint32_t>(s.x) & 0xffffu;
                uint32_t d = cum - sp_start;
                uint32_t freq = sp_freq;
                int32_t value = sp_value;
                if (__builtin_expect(!(d < sp_freq), 0)) {
                    // miss: warp-wide search, lane l compares CDF entry l (and l + 32)
                    int k;
                    uint32_t m = __ballot_sync(0xffffffffu, c0 > static_cast<int32_t>(cum));
                    if (m) {
                        k = __ffs(m) - 1;
                    } else {
                        m = __ballot_sync(0xffffffffu, c1 > static_cast<int32_t>(cum));
                        if (m) {
                            k = 32 + __ffs(m) - 1;
                        } else {
                            k = 64;
                            for (;; k += 32) {
                                const uint32_t mm = __ballot_sync(0xffffffffu, drow[k + lane] > static_cast<int32_t>(cum));
                                if (mm) {
                                    k += __ffs(mm) - 1;
                                    break;
                                }
                            }
                        }
                    }
                    const uint32_t start = static_cast<uint32_t>(drow[k - 1]);
                    freq = static_cast<uint32_t>(drow[k]) - start;
                    d = cum - start;
                    value = k - 1;
                    if (value != max_value) {
                        sp_start = start;
                        sp_freq = freq;
                        sp_value = value;
                    }
                }
                uint64_t xn = static_cast<uint64_t>(freq) * (s.x >> kRansPrecision) + d;
                if (xn < kL) {
                    xn = (xn << 32) | s.next_w;
                    dec_advance_word(s, lane);
                }
                s.x = xn;
                if (__builtin_expect(value == max_value, 0)) {
                    const DecEscapeResult r = dec_escape(s, max_value, lane);
                    s = r.s;
                    value = r.value;
                }
                s_out[j] = value;  // uniform value, one shared-memory word
            }
            __syncwarp();
            if (lane < cnt) {
                const int32_t v = s_out[lane] + offset;
                const uint32_t o = done + base + lane;
                if (osym) osym[o] = v;
                if (oval) oval[o] = static_cast<float>(v) + mean;
            }
            __syncwarp();
        }
        done += row_n;
    }
    if (s.truncated && lane == 0) atomicOr(status, SC2_FAULT_STREAM_TRUNCATED);
}

}  // namespace

int launch_rans_encode_fast(const int32_t *symbols, int batch, int64_t n, int64_t spatial, const void *tables, int n_rows,
                            int cdf_stride, uint8_t *arena, int64_t slot_bytes, int32_t *lengths, int32_t *status,
                            cudaStream_t st) {
    size_t smem = 64 * 16;
    const size_t table_bytes = static_cast<size_t>(n_rows) * cdf_stride * 16;
    if (table_bytes + smem <= 96 * 1024) smem += table_bytes;
    static bool configured = false;
    if (!configured) {
        SC2_CUDA_TRY(cudaFuncSetAttribute(rans_encode_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured = true;
    }
    rans_encode_fast_kernel<<<batch, 32, smem, st>>>(symbols, batch, static_cast<uint32_t>(n), static_cast<uint32_t>(spatial),
                                                      tables, arena, slot_bytes, lengths, status);
    SC2_LAUNCH_CHECK("rans_encode_fast_kernel");
    return SC2_OK;
}

int launch_rans_decode_fast(const uint8_t *packed, const int64_t *offsets, int batch, int64_t n, int64_t spatial,
                            const void *tables, int n_rows, int cdf_stride, int32_t *out_symbols, float *out_values,
                            const float *means, int32_t *status, cudaStream_t st) {
    size_t smem = (kRingWords + 32) * 4;
    const size_t table_bytes = static_cast<size_t>(n_rows) * ((cdf_stride + 31) / 32 * 32) * 4;
    if (table_bytes + smem <= 96 * 1024) smem += table_bytes;
    static bool configured = false;
    if (!configured) {
        SC2_CUDA_TRY(cudaFuncSetAttribute(rans_decode_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        configured = true;
    }
    rans_decode_fast_kernel<<<batch, 32, smem, st>>>(packed, offsets, batch, static_cast<uint32_t>(n),
                                                      static_cast<uint32_t>(spatial), tables, out_symbols, out_values, means, status);
    SC2_LAUNCH_CHECK("rans_decode_fast_kernel");
    return SC2_OK;
}

}  // namespace sc2
