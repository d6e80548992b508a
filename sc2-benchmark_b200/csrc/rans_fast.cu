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
// Decoder word ring: 4 blocks of 32 words.  Crossing into block B waits for every refill issued so far (words < B + 64 are then
// in shared memory) and requests [B + 64, B + 96) into the block just left: any word up to 32 ahead of the read position is
// always valid, so the hot loop may look two words ahead and may notice a crossing a few words late.
constexpr int kRingWords = 128;
constexpr uint32_t kRefillAhead = 64;
constexpr int kSpecGroup = 8;   // symbols decoded speculatively per group (see the hot loop)

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

// one regular symbol, unchecked (caller guarantees s.pw > 0): the dependent chain is
//   ISETP -> SEL -> IMAD.WIDE x3 (mul.hi.u64) -> SHF -> IMAD.WIDE -> IADD
// The renormalisation's side effects (the emitted word, the write position) are issued AFTER the division chain: a single warp issues
// in order, so with the store first (as the source reads) its address arithmetic and the register hand-over between the stored word
// and the selected state sat in front of the chain and cost ~25 of 112 cycles per symbol (ncu source view, profiles/r4_coder_*).
__device__ __forceinline__ void enc_step(EncChain &s, const uint4 a, const uint4 b) {
    const uint32_t xl = static_cast<uint32_t>(s.x), xh = static_cast<uint32_t>(s.x >> 32);
    const bool ren = xh >= a.z;  // x >= freq << 47
    const uint32_t emit = xl;
    const uint64_t y = ren ? static_cast<uint64_t>(xh) : s.x;
    const uint64_t rcp = (static_cast<uint64_t>(a.y) << 32) | a.x;
    const uint64_t q = __umul64hi(y, rcp) >> b.y;
    s.x = y + b.x + q * static_cast<uint64_t>(a.w);
    asm volatile("" ::: "memory");  // keep the store below the chain
    if (ren) s.words[s.pw - 1] = emit;
    s.pw -= ren ? 1u : 0u;
}

// cold path: a chunk that contains escapes, is the (short) last chunk, or runs close to the end of the arena slot
__device__ __noinline__ EncChain enc_chunk_careful(EncChain s, const uint4 *sA, const uint4 *sB, int cnt) {
    for (int j = 0; j < cnt; ++j) {
        const uint4 a = sA[j], b = sB[j];
        if (b.z) {
            const uint32_t raw = b.w;
            const int n_bypass = raw == 0 ? 0 : (35 - __clz(raw)) >> 2;
            for (int k = n_bypass - 1; k >= -1; --k) {
                const uint32_t val = k >= 0 ? ((raw >> (4 * k)) & kMaxBypassVal) : static_cast<uint32_t>(n_bypass);
                if (s.x >= (1ull << 59)) enc_emit_checked(s);
                s.x = (s.x << kBypassPrecision) | val;
            }
        }
        if (static_cast<uint32_t>(s.x >> 32) >= a.z) enc_emit_checked(s);
        const uint64_t rcp = (static_cast<uint64_t>(a.y) << 32) | a.x;
        const uint64_t q = __umul64hi(s.x, rcp) >> b.y;
        s.x = s.x + b.x + q * static_cast<uint64_t>(a.w);
    }
    return s;
}

// IDX: the CDF row of every symbol comes from `indexes` (GaussianConditional: one row per element) instead of i / spatial
template <bool IDX>
__global__ void __launch_bounds__(32)
rans_encode_fast_kernel(const int32_t *__restrict__ symbols, const int32_t *__restrict__ indexes, int batch, uint32_t n, uint32_t spatial,
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
    const int32_t *idx = IDX ? indexes + static_cast<int64_t>(b) * n : nullptr;
    EncChain s;
    s.x = kL;
    s.words = reinterpret_cast<uint32_t *>(arena + static_cast<int64_t>(b) * slot_bytes);
    const uint32_t slot_words = static_cast<uint32_t>(slot_bytes >> 2);
    s.pw = slot_words;
    s.overflow = 0u;

    // symbols are consumed back to front; lane l of a chunk ending at `hi` owns symbol hi - 1 - l
    int32_t nxt = (lane < static_cast<int>(n)) ? __ldg(sym + (n - 1 - lane)) : 0;
    int32_t nxt_row = (IDX && lane < static_cast<int>(n)) ? __ldg(idx + (n - 1 - lane)) : 0;
    for (uint32_t hi = n; hi > 0; hi = hi > 32 ? hi - 32 : 0) {
        const int cnt = hi >= 32 ? 32 : static_cast<int>(hi);
        const int32_t cur = nxt;
        const int32_t cur_row = nxt_row;
        if (hi > 32) {  // request the next chunk now: its HBM/L2 latency hides behind this chunk's chain
            const uint32_t hn = hi - 32;
            nxt = (static_cast<uint32_t>(lane) < hn) ? __ldg(sym + (hn - 1 - lane)) : 0;
            if (IDX) nxt_row = (static_cast<uint32_t>(lane) < hn) ? __ldg(idx + (hn - 1 - lane)) : 0;
        }
        uint32_t my_esc = 0;
        if (lane < cnt) {
            const uint32_t i = hi - 1 - lane;
            int row = IDX ? cur_row : static_cast<int>(i / spatial);
            if (IDX && static_cast<uint32_t>(row) >= static_cast<uint32_t>(t.n_rows)) {
                // caller-supplied CDF index out of range: flag it and code the symbol with row 0 instead of reading outside the tables
                atomicOr(status, SC2_FAULT_BAD_INDEX);
                row = 0;
            }
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
            my_esc = esc;
            const int eidx = row * t.cdf_stride + value;
            const uint4 e = staged ? s_enc[eidx] : __ldg(t.enc + eidx);
            // unpack here, in parallel across lanes, so the chain only sees ready-to-use operands
            sA[lane] = make_uint4(e.x, e.y, e.w << 15, 65536u - e.w);               // rcp_lo, rcp_hi, renorm threshold, 2^16 - freq
            sB[lane] = make_uint4(e.z & 0x1ffffu, (e.z >> 24) & 15u, esc, raw);     // bias, shift, escape?, raw
        }
        // does any symbol of this chunk escape?  (uniform; computed off the chain)
        const bool any_escape = __any_sync(0xffffffffu, my_esc != 0u);
        __syncwarp();
        if (cnt == 32 && !any_escape && s.pw >= 40u) {
            // ---- fast chain: 32 regular symbols, no escapes, room for 32 words guaranteed -> no checks inside.
            // Two register sets alternate so that the next entry is already loaded when a step starts.
            uint4 a0 = sA[0], b0 = sB[0];
#pragma unroll 1
            for (int j = 0; j < 32; j += 2) {
                const uint4 a1 = sA[j + 1], b1 = sB[j + 1];
                enc_step(s, a0, b0);
                if (j + 2 < 32) {
                    a0 = sA[j + 2];
                    b0 = sB[j + 2];
                }
                enc_step(s, a1, b1);
            }
        } else {
            s = enc_chunk_careful(s, sA, sB, cnt);
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
    uint32_t p;        // index of the word held in `next_w` (= number of words consumed so far)
    uint32_t next_w;   // words[p], already in a register
};

struct DecStream {
    const uint32_t *words;
    uint32_t n_words;
    uint32_t *ring;
};

__device__ __forceinline__ void ring_fill(const DecStream &st, uint32_t first, int lane) {
    const uint32_t w = first + lane;
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(st.ring + (w & (kRingWords - 1))));
    const bool ok = w < st.n_words;
    const uint32_t *src = ok ? st.words + w : st.words;
    const int src_bytes = ok ? 4 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
}

// cold path: the word index crossed a ring half: wait for the previous refill, reuse the half that was just left
__device__ __noinline__ void dec_refill(const uint32_t *words, uint32_t n_words, uint32_t *ring, uint32_t first, int lane) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    DecStream st;
    st.words = words;
    st.n_words = n_words;
    st.ring = ring;
    ring_fill(st, first, lane);
}

// consume next_w and fetch the following word from the ring
__device__ __forceinline__ void dec_advance_word(DecChain &s, const DecStream &st, int lane) {
    ++s.p;
    if (__builtin_expect((s.p & 31u) == 0u, 0)) dec_refill(st.words, st.n_words, st.ring, s.p + kRefillAhead, lane);
    s.next_w = st.ring[s.p & (kRingWords - 1)];
}

__device__ __forceinline__ uint32_t dec_get_nibble(DecChain &s, const DecStream &st, int lane) {
    const uint32_t val = static_cast<uint32_t>(s.x) & kMaxBypassVal;
    s.x >>= kBypassPrecision;
    if (s.x < kL) {
        s.x = (s.x << 32) | s.next_w;
        dec_advance_word(s, st, lane);
    }
    return val;
}

struct DecMissResult {
    DecChain s;
    uint32_t sp_start, sp_freq;
    int32_t sp_value;
    int32_t value;
};

// cold path (by value, see enc_chunk_careful): the speculation missed.  Warp-wide search (lane l compares CDF entries
// l and l + 32, further 32-entry groups from memory), then the COMPLETE symbol step including escapes.
__device__ __noinline__ DecMissResult dec_miss(DecChain s, const uint32_t *words, uint32_t n_words, uint32_t *ring,
                                               const int32_t *drow, int32_t c0, int32_t c1, int32_t max_value,
                                               uint32_t sp_start, uint32_t sp_freq, int32_t sp_value, int lane) {
    DecStream st;
    st.words = words;
    st.n_words = n_words;
    st.ring = ring;
    const int32_t cum = static_cast<int32_t>(static_cast<uint32_t>(s.x) & 0xffffu);
    int k;
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
    const uint32_t start = static_cast<uint32_t>(drow[k - 1]);
    const uint32_t freq = static_cast<uint32_t>(drow[k]) - start;
    int32_t value = k - 1;
    s.x = static_cast<uint64_t>(freq) * (s.x >> kRansPrecision) + (static_cast<uint32_t>(cum) - start);
    if (s.x < kL) {
        s.x = (s.x << 32) | s.next_w;
        dec_advance_word(s, st, lane);
    }
    DecMissResult r;
    r.sp_start = sp_start;
    r.sp_freq = sp_freq;
    r.sp_value = sp_value;
    if (value != max_value) {
        r.sp_start = start;
        r.sp_freq = freq;
        r.sp_value = value;
    } else {
        uint32_t val = dec_get_nibble(s, st, lane);
        int n_bypass = static_cast<int>(val);
        while (val == kMaxBypassVal) {
            val = dec_get_nibble(s, st, lane);
            n_bypass += static_cast<int>(val);
        }
        uint32_t raw = 0;
        for (int q = 0; q < n_bypass; ++q) {
            const uint32_t nib = dec_get_nibble(s, st, lane);
            if (q < 8) raw |= nib << (4 * q);
        }
        const int32_t v = static_cast<int32_t>(raw >> 1);
        value = (raw & 1u) ? -v - 1 : v + max_value;
    }
    r.s = s;
    r.value = value;
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
    DecStream st;
    st.words = reinterpret_cast<const uint32_t *>(packed + off);
    st.n_words = static_cast<uint32_t>(n_bytes >> 2);
    st.ring = s_ring;
    ring_fill(st, 0, lane);
    ring_fill(st, 32, lane);
    ring_fill(st, 64, lane);
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncwarp();
    DecChain s;
    s.x = static_cast<uint64_t>(s_ring[0]) | (static_cast<uint64_t>(s_ring[1]) << 32);
    s.p = 2;
    s.next_w = s_ring[2];

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
        // speculation state: the last regular symbol decoded in this row (freq 0 = nothing yet -> first symbol misses)
        uint32_t sp_start = 0, sp_freq = 0;
        int32_t sp_value = 0;
        int streak = 0;  // symbols since the last speculation miss (run mode from kSpecGroup on)
        for (uint32_t base = 0; base < row_n; base += 32) {
            const int cnt = (row_n - base) >= 32 ? 32 : static_cast<int>(row_n - base);
            // ---- hot loop: nothing but the hit chain on 32-bit halves of the state; a rare event (speculation miss,
            // ring half consumed) jumps out, is handled, and the loop resumes.  Chain per symbol:
            //   LOP -> IADD -> ISETP(miss?) -> IMAD.WIDE -> IMAD -> SHF/LOP -> ISETP(renorm?) -> SEL
            uint32_t xl = static_cast<uint32_t>(s.x), xh = static_cast<uint32_t>(s.x >> 32);
            int j = 0;
#define SC2_DEC_STEP(K)                                                                              \
    {                                                                                                \
        const uint32_t d = (xl & 0xffffu) - sp_start;                                                \
        if (d >= sp_freq) { j += (K); goto miss_event; }                                             \
        const uint64_t prod = static_cast<uint64_t>(sp_freq) * __funnelshift_r(xl, xh, 16) + d;      \
        const uint32_t nl = static_cast<uint32_t>(prod);                                             \
        const uint32_t nh = static_cast<uint32_t>(prod >> 32) + sp_freq * (xh >> 16);                \
        const bool ren = (nh | (nl >> 31)) == 0u; /* x < 2^31 */                                     \
        xl = ren ? s.next_w : nl;                                                                    \
        xh = ren ? nl : nh;                                                                          \
        s_out[j + (K)] = sp_value;                                                                   \
        if (ren) {                                                                                   \
            ++s.p;                                                                                   \
            s.next_w = s_ring[s.p & (kRingWords - 1)];                                               \
            if ((s.p & 31u) == 0u) { j += (K) + 1; goto refill_event; }                              \
        }                                                                                            \
    }
        resume:
            while (j < cnt) {
                if (streak >= kSpecGroup && j + kSpecGroup <= cnt) {
                    // ---- run mode: the last symbols all hit the speculation.  Decode a GROUP on copies of the state with no branch at
                    // all (miss flags are OR-ed, renormalisation is select-only, the two next words sit in registers) and commit it
                    // if every symbol hit; otherwise drop the copies and take the checked steps below.  Chain per symbol:
                    //   LOP -> IADD -> IMAD.WIDE -> IMAD -> LOP/ISETP -> SEL
                    uint32_t gxl = xl, gxh = xh, gp = s.p, w0 = s.next_w, w1 = s_ring[(s.p + 1u) & (kRingWords - 1)];
                    bool bad = false;
#pragma unroll
                    for (int k = 0; k < kSpecGroup; ++k) {
                        const uint32_t d = (gxl & 0xffffu) - sp_start;
                        bad |= d >= sp_freq;
                        const uint64_t prod = static_cast<uint64_t>(sp_freq) * __funnelshift_r(gxl, gxh, 16) + d;
                        const uint32_t nl = static_cast<uint32_t>(prod);
                        const uint32_t nh = static_cast<uint32_t>(prod >> 32) + sp_freq * (gxh >> 16);
                        const bool ren = (nh | (nl >> 31)) == 0u;
                        gxl = ren ? w0 : nl;
                        gxh = ren ? nl : nh;
                        gp += ren ? 1u : 0u;
                        w0 = ren ? w1 : w0;
                        w1 = s_ring[(gp + 1u) & (kRingWords - 1)];
                    }
                    if (!bad) {
#pragma unroll
                        for (int k = 0; k < kSpecGroup; ++k) s_out[j + k] = sp_value;
                        j += kSpecGroup;
                        xl = gxl;
                        xh = gxh;
                        const bool crossed = ((gp ^ s.p) & 32u) != 0u;
                        s.p = gp;
                        s.next_w = w0;
                        if (crossed) dec_refill(st.words, st.n_words, st.ring, (gp & ~31u) + kRefillAhead, lane);
                        continue;
                    }
                    streak = 0;
                }
                if (j + 4 <= cnt) {
                    SC2_DEC_STEP(0)
                    SC2_DEC_STEP(1)
                    SC2_DEC_STEP(2)
                    SC2_DEC_STEP(3)
                    j += 4;
                    streak += 4;
                } else {
                    SC2_DEC_STEP(0)
                    j += 1;
                    streak += 1;
                }
            }
            goto chunk_done;
        refill_event:
            // the word index crossed a ring half: top up the half that was just left, then re-read next_w
            dec_refill(st.words, st.n_words, st.ring, s.p + kRefillAhead, lane);
            s.next_w = s_ring[s.p & (kRingWords - 1)];
            goto resume;
        miss_event : {
            streak = 0;
            s.x = (static_cast<uint64_t>(xh) << 32) | xl;
            const DecMissResult r = dec_miss(s, st.words, st.n_words, st.ring, drow, c0, c1, max_value, sp_start, sp_freq,
                                             sp_value, lane);
            s = r.s;
            xl = static_cast<uint32_t>(s.x);
            xh = static_cast<uint32_t>(s.x >> 32);
            sp_start = r.sp_start;
            sp_freq = r.sp_freq;
            sp_value = r.sp_value;
            s_out[j] = r.value;
            ++j;
            goto resume;
        }
        chunk_done:
#undef SC2_DEC_STEP
            s.x = (static_cast<uint64_t>(xh) << 32) | xl;
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
    // a well-formed stream is consumed exactly; reading past the end (zero-filled) means it was truncated
    if (s.p > st.n_words && lane == 0) atomicOr(status, SC2_FAULT_STREAM_TRUNCATED);
}

}  // namespace

int launch_rans_encode_fast(const int32_t *symbols, const int32_t *indexes, int batch, int64_t n, int64_t spatial, const void *tables,
                            int n_rows, int cdf_stride, uint8_t *arena, int64_t slot_bytes, int32_t *lengths, int32_t *status,
                            cudaStream_t st) {
    size_t smem = 64 * 16;
    const size_t table_bytes = static_cast<size_t>(n_rows) * cdf_stride * 16;
    if (table_bytes + smem <= 96 * 1024) smem += table_bytes;
    const uint32_t un = static_cast<uint32_t>(n), us = static_cast<uint32_t>(indexes ? 1 : spatial);
    if (indexes) {
        static std::atomic<uint64_t> configured{0};  // per device ordinal
        if (int rc = ensure_dyn_smem(rans_encode_fast_kernel<true>, 100 * 1024, configured)) return rc;
        rans_encode_fast_kernel<true><<<batch, 32, smem, st>>>(symbols, indexes, batch, un, us, tables, arena, slot_bytes, lengths, status);
    } else {
        static std::atomic<uint64_t> configured{0};  // per device ordinal
        if (int rc = ensure_dyn_smem(rans_encode_fast_kernel<false>, 100 * 1024, configured)) return rc;
        rans_encode_fast_kernel<false><<<batch, 32, smem, st>>>(symbols, nullptr, batch, un, us, tables, arena, slot_bytes, lengths, status);
    }
    SC2_LAUNCH_CHECK("rans_encode_fast_kernel");
    return SC2_OK;
}

int launch_rans_decode_fast(const uint8_t *packed, const int64_t *offsets, int batch, int64_t n, int64_t spatial,
                            const void *tables, int n_rows, int cdf_stride, int32_t *out_symbols, float *out_values,
                            const float *means, int32_t *status, cudaStream_t st) {
    size_t smem = (kRingWords + 32) * 4;
    const size_t table_bytes = static_cast<size_t>(n_rows) * ((cdf_stride + 31) / 32 * 32) * 4;
    if (table_bytes + smem <= 96 * 1024) smem += table_bytes;
    static std::atomic<uint64_t> configured{0};  // per device ordinal
    if (int rc = ensure_dyn_smem(rans_decode_fast_kernel, 100 * 1024, configured)) return rc;
    rans_decode_fast_kernel<<<batch, 32, smem, st>>>(packed, offsets, batch, static_cast<uint32_t>(n),
                                                      static_cast<uint32_t>(spatial), tables, out_symbols, out_values, means, status);
    SC2_LAUNCH_CHECK("rans_decode_fast_kernel");
    return SC2_OK;
}

}  // namespace sc2
