// rans_lanes.cu -- channel-mode rANS kernels with ONE LANE PER STREAM (32 CompressAI streams per warp).
//
// A CompressAI stream is one serial chain: its speed is the length of the dependent instruction sequence
// per symbol, and nothing a warp's other 31 lanes do can shorten it.  rans_fast.cu spends a whole warp on each stream, so
// a batch of 256 streams issues 32x more warp instructions than it has useful work -- enough to take ~1.7 ms of the whole
// GPU's issue slots per batch once several batches are in flight, which is then no longer hidden behind the tensor-core
// kernels of the other batches.  Here every lane runs the chain of its OWN stream: the same instruction stream now
// serves 32 streams, a batch of 256 is 8 warps, and the coder all but disappears next to the convolution kernels it
// overlaps with (it is sized to co-reside with them: 1 warp, <= 21 KB of shared memory per block).
//
//   encode  per lane: chain = 4 x IMAD.WIDE (partial products of the exact 64-bit reciprocal division, shared between
//           the "renormalise first" and the "no renormalisation" case, which are BOTH computed and selected at the end)
//           -> carry adds -> SEL -> SHF -> IMAD.WIDE.  Table entries are looked up 8 symbols ahead and symbols are
//           loaded 16..24 ahead (register rings), so neither shared-memory nor L2 latency is on the chain.
//   decode  per lane: chain = LOP -> LDS (4096-bucket LUT of the CURRENT CDF row: start | freq-1 for every 16-wide slice
//           of the 2^16 range that lies inside one symbol; rebuilt by the warp at each row change, all streams of a
//           warp being at the same position) -> IMAD.WIDE -> renormalise (SEL).  Slices that straddle a symbol boundary
//           are flagged and resolved by a short walk over the CDF row.
//
// Bit-exactness: identical arithmetic to rans.cu / rans_fast.cu / the oracle; tests compare bytes.
#include <cstdlib>

#include "common.cuh"

namespace sc2 {

namespace {

constexpr int kMaxWarps = 8;             // warps (x 32 streams) per block, see the launchers
constexpr uint32_t kLutBuckets = 4096;   // 2^16 / 16
constexpr uint32_t kLutFlag = 0xffffffffu;
constexpr int kSymTab = 256;             // symbols per row with a shared-memory (start, freq) entry
constexpr int kEncStageLimit = 16 * 1024;  // encoder table staged in shared memory up to this size

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
struct EncState {
    uint32_t xl, xh;
    uint32_t pw;        // words[pw - 1] is the next free word (the stream grows downwards from the end of the slot)
    uint32_t overflow;
};

// cold path (by value: a by-reference state would live in local memory): the bypass digits of one escaped symbol
__device__ __noinline__ EncState enc_escape(EncState s, uint32_t *words, uint32_t raw) {
    const int n_bypass = raw == 0 ? 0 : (35 - __clz(raw)) >> 2;
    for (int k = n_bypass - 1; k >= -1; --k) {
        const uint32_t val = k >= 0 ? ((raw >> (4 * k)) & kMaxBypassVal) : static_cast<uint32_t>(n_bypass);
        if (s.xh >= (1u << 27)) {  // x >= 2^59
            if (s.pw != 0u) {
                --s.pw;
                words[s.pw] = s.xl;
            } else {
                s.overflow = 1u;
            }
            s.xl = s.xh;
            s.xh = 0u;
        }
        s.xh = (s.xh << kBypassPrecision) | (s.xl >> (32 - kBypassPrecision));
        s.xl = (s.xl << kBypassPrecision) | val;
    }
    return s;
}

// One regular symbol.  Entry = (rcp_lo, rcp_hi, bias | shift << 24, freq).
//   q = floor(y / freq) through the exact reciprocal, y = x or (renormalised) x >> 32;  x' = y + bias + q * (2^16 - freq)
// CHECKED = false: the caller guarantees a free word (no arena bound check on the chain).
template <bool CHECKED>
__device__ __forceinline__ void enc_step(EncState &s, uint32_t *words, const uint4 e) {
    const uint32_t rl = e.x, rh = e.y, freq = e.w;
    const bool ren = s.xh >= (freq << 15);  // x >= freq << 47
    const uint64_t p0 = static_cast<uint64_t>(s.xl) * rl;
    const uint64_t p1 = static_cast<uint64_t>(s.xl) * rh;
    const uint64_t p2 = static_cast<uint64_t>(s.xh) * rl;
    const uint64_t p3 = static_cast<uint64_t>(s.xh) * rh;
    const uint64_t mid = (p0 >> 32) + static_cast<uint32_t>(p1) + static_cast<uint32_t>(p2);
    const uint64_t q_keep = p3 + (p1 >> 32) + (p2 >> 32) + (mid >> 32);  // mulhi64(x, rcp)
    const uint64_t q_ren = (p3 + (p2 >> 32)) >> 32;                       // mulhi64(x >> 32, rcp): same partial products
    const uint32_t bias = e.z & 0x1ffffu, shift = e.z >> 24;
    const uint64_t x = (static_cast<uint64_t>(s.xh) << 32) | s.xl;
    const uint64_t yb = (ren ? static_cast<uint64_t>(s.xh) : x) + bias;
    const uint64_t q = (ren ? q_ren : q_keep) >> shift;
    if (ren) {
        if (!CHECKED || s.pw != 0u) {
            --s.pw;
            words[s.pw] = s.xl;
        } else {
            s.overflow = 1u;
        }
    }
    const uint32_t cmpl = 65536u - freq;
    const uint64_t r = static_cast<uint64_t>(static_cast<uint32_t>(q)) * cmpl + yb;
    s.xl = static_cast<uint32_t>(r);
    s.xh = static_cast<uint32_t>(r >> 32) + static_cast<uint32_t>(q >> 32) * cmpl;
}

struct RowCursor {  // warp-uniform: the CDF row of a symbol position walking DOWN from n - 1
    int row;
    uint32_t rem;  // position inside the row
    uint32_t off, maxv;
    int ebase;
};

__device__ __forceinline__ void cursor_load(RowCursor &c, const Tables &t) {
    c.off = static_cast<uint32_t>(__ldg(t.offsets + c.row));
    c.maxv = static_cast<uint32_t>(__ldg(t.sizes + c.row) - 2);
    c.ebase = c.row * t.cdf_stride;
}

__device__ __forceinline__ void cursor_step(RowCursor &c, const Tables &t, uint32_t spatial) {
    if (c.rem == 0u) {
        if (c.row > 0) {
            --c.row;
            c.rem = spatial - 1u;
            cursor_load(c, t);
        }
    } else {
        --c.rem;
    }
}

// symbol -> encoder entry of its (clamped) value; `esc` accumulates "some symbol of the block escapes"
__device__ __forceinline__ uint4 enc_lookup(int32_t sym, const RowCursor &c, const uint4 *enc_tab, bool &esc) {
    uint32_t v = static_cast<uint32_t>(sym) - c.off;  // negative values wrap to huge: one unsigned compare covers both tails
    const bool out = v >= c.maxv;
    v = out ? c.maxv : v;
    esc = esc || out;
    return enc_tab[c.ebase + static_cast<int>(v)];
}

__global__ void __launch_bounds__(kMaxWarps * 32)
rans_encode_lanes_kernel(const int32_t *__restrict__ symbols, int batch, uint32_t n, uint32_t spatial,
                         const void *__restrict__ tables, uint8_t *__restrict__ arena, int64_t slot_bytes,
                         int32_t *__restrict__ lengths, int32_t *__restrict__ status, const TraceSink trace) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long trace_t0 = trace.buf ? trace_now() : 0ull;
    const Tables t = view(tables);
    uint4 *s_enc = reinterpret_cast<uint4 *>(smem_raw);
    const int enc_entries = t.n_rows * t.cdf_stride;
    uint32_t dyn_smem;
    asm volatile("mov.u32 %0, %%dynamic_smem_size;" : "=r"(dyn_smem));
    const bool staged = dyn_smem >= static_cast<uint32_t>(enc_entries) * 16u;
    if (staged)
        for (int i = threadIdx.x; i < enc_entries; i += blockDim.x) s_enc[i] = __ldg(t.enc + i);
    __syncthreads();
    if (b >= batch) return;  // no warp- or block-level operation below: spare lanes simply leave
    const uint4 *enc_tab = staged ? s_enc : t.enc;

    const int32_t *sym = symbols + static_cast<int64_t>(b) * n;
    uint32_t *words = reinterpret_cast<uint32_t *>(arena + static_cast<int64_t>(b) * slot_bytes);
    const uint32_t slot_words = static_cast<uint32_t>(slot_bytes >> 2);
    EncState s;
    s.xl = 1u << 31;  // RANS64_L
    s.xh = 0u;
    s.pw = slot_words;
    s.overflow = 0u;

    if (n > 0) {
        // Symbols are consumed back to front in blocks of 8; I = index of the symbol the chain is at.  Register rings:
        //   ent[j]   entry of symbol I - j       (looked up one block ahead of the chain)
        //   ring[j]  symbol I - j - 8            (loaded two blocks ahead of the chain)
        int32_t I = static_cast<int32_t>(n) - 1;
        const int32_t *sp = sym + I;
        RowCursor cc;  // row of the chain position
        cc.row = static_cast<int>(static_cast<uint32_t>(I) / spatial);
        cc.rem = static_cast<uint32_t>(I) % spatial;
        cursor_load(cc, t);
        RowCursor cur = cc;  // row of the lookup position
        uint4 ent[8];
        int32_t ring[8];
        bool esc_cur = false;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            ent[j] = make_uint4(0, 0, 0, 1);
            if (I - j >= 0) {
                ent[j] = enc_lookup(__ldg(sp - j), cur, enc_tab, esc_cur);
                cursor_step(cur, t, spatial);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) ring[j] = (I - j - 8 >= 0) ? __ldg(sp - j - 8) : 0;

        while (I >= 0) {
            bool esc_next = false;
            if (!esc_cur && I >= 23 && cur.rem >= 8u && cc.rem >= 8u && s.pw >= 8u) {
                // ---- fast block: 8 regular symbols of one row, no bounds to check; nothing but the chain and the rings
                // The 8 symbols a lane loads per block share one 32-byte sector (a new one every block); the compiler sinks those loads
                // towards their first use, so the sector miss (L2 / HBM, 32 different sectors per warp) was waited for at the top of
                // every block: 41 % of the kernel's stall samples on one move (ncu source view, profiles/r4_lanes_*).  The sector
                // of the block after next is therefore requested here, two blocks (~3.5 k cycles) early: 8.46 -> 7.77 ms per batch of 256.
                // (a prefetch, not a load into an unused register: that load shares a scoreboard with the ones the chain waits for)
                asm volatile("prefetch.global.L1 [%0];" ::"l"(sp - (I >= 32 ? 32 : 0)));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint4 e = ent[j];
                    ent[j] = enc_lookup(ring[j], cur, enc_tab, esc_next);
                    ring[j] = __ldg(sp - j - 16);  // (16-byte loads per 4 symbols measured slower: 6.6 vs 5.1 ms per batch)
                    enc_step<false>(s, words, e);
                }
                cur.rem -= 8u;
                cc.rem -= 8u;
            } else {
                // ---- general block: escapes, a row boundary, the head of the stream or a nearly full arena slot
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (I - j < 0) break;
                    const uint4 e = ent[j];
                    if (I - j - 8 >= 0) {
                        ent[j] = enc_lookup(ring[j], cur, enc_tab, esc_next);
                        cursor_step(cur, t, spatial);
                    }
                    ring[j] = (I - j - 16 >= 0) ? __ldg(sp - j - 16) : 0;
                    const uint32_t v = static_cast<uint32_t>(__ldg(sp - j)) - cc.off;
                    if (v >= cc.maxv) {  // escape: bypass digits first (the decoder reads them after the escape symbol)
                        const int32_t value = static_cast<int32_t>(v);
                        const uint32_t raw = value < 0 ? static_cast<uint32_t>(-2 * value - 1)
                                                       : 2u * (v - cc.maxv);
                        s = enc_escape(s, words, raw);
                    }
                    enc_step<true>(s, words, e);
                    cursor_step(cc, t, spatial);
                }
            }
            esc_cur = esc_next;
            I -= 8;
            sp -= 8;
        }
    }
    if (s.pw >= 2u) {
        s.pw -= 2;
        words[s.pw] = s.xl;
        words[s.pw + 1] = s.xh;
    } else {
        s.overflow = 1u;
    }
    lengths[b] = s.overflow ? 0 : static_cast<int32_t>((slot_words - s.pw) * 4u);
    if (s.overflow) atomicOr(status, SC2_FAULT_ARENA_OVERFLOW);
    if (lane == 0) trace_emit(trace, TRACE_RANS_ENCODE, trace_t0, batch);
}

// ---------------------------------------------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------------------------------------------
// shared-memory loads through 32-bit shared addresses (through generic pointers the compiler rebuilt the shared window base
// -- an S2UR -- inside the symbol loop, on the chain)
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}

struct DecState {
    uint32_t xl, xh;
    uint32_t p;       // index of the word held in next_w (= number of words consumed so far)
    uint32_t next_w;  // words[p], already in a register
};

// consume next_w and fetch the following word (clamped to the stream: reading past a truncated stream repeats its last
// word, and the kernel reports the truncation).  A stream spends >= 1 symbol per 16 bits, so the load has at least two
// chain steps to land.
__device__ __forceinline__ void dec_take_word(DecState &s, const uint32_t *words, uint32_t n_words) {
    ++s.p;
    s.next_w = __ldg(words + min(s.p, n_words - 1u));
}

// cold path: the bypass digits of an escaped symbol -> its value relative to the row offset
struct DecEscape {
    DecState s;
    int32_t value;
};

__device__ __noinline__ DecEscape dec_escape(DecState s, const uint32_t *words, uint32_t n_words, int32_t max_value) {
    auto nibble = [&]() {
        const uint32_t val = s.xl & kMaxBypassVal;
        s.xl = (s.xl >> kBypassPrecision) | (s.xh << (32 - kBypassPrecision));
        s.xh >>= kBypassPrecision;
        if (s.xh == 0u && (s.xl >> 31) == 0u) {  // x < 2^31
            s.xh = s.xl;
            s.xl = s.next_w;
            dec_take_word(s, words, n_words);
        }
        return val;
    };
    uint32_t val = nibble();
    int n_bypass = static_cast<int>(val);
    while (val == kMaxBypassVal) {
        val = nibble();
        n_bypass += static_cast<int>(val);
    }
    uint32_t raw = 0;
    for (int q = 0; q < n_bypass; ++q) {
        const uint32_t nib = nibble();
        if (q < 8) raw |= nib << (4 * q);
    }
    const int32_t v = static_cast<int32_t>(raw >> 1);
    DecEscape r;
    r.s = s;
    r.value = (raw & 1u) ? -v - 1 : v + max_value;
    return r;
}

// cold path, the COMPLETE step of a symbol whose bucket straddles a symbol boundary (walk the row from the bucket's first
// symbol) or that decodes to the escape symbol (read the bypass digits)
__device__ __noinline__ DecEscape dec_slow_step(DecState s, const uint32_t *words, uint32_t n_words,
                                                const int32_t *__restrict__ crow, int32_t value, int32_t max_value) {
    const uint32_t cum = s.xl & 0xffffu;
    int32_t c0 = __ldg(crow + value), c1 = __ldg(crow + value + 1);
    while (static_cast<uint32_t>(c1) <= cum) {
        ++value;
        c0 = c1;
        c1 = __ldg(crow + value + 1);
    }
    const uint32_t start = static_cast<uint32_t>(c0), freq = static_cast<uint32_t>(c1 - c0);
    const uint64_t prod = static_cast<uint64_t>(freq) * __funnelshift_r(s.xl, s.xh, 16) + (cum - start);
    s.xl = static_cast<uint32_t>(prod);
    s.xh = static_cast<uint32_t>(prod >> 32) + freq * (s.xh >> 16);
    if (s.xh == 0u && (s.xl >> 31) == 0u) {
        s.xh = s.xl;
        s.xl = s.next_w;
        dec_take_word(s, words, n_words);
    }
    if (value == max_value) return dec_escape(s, words, n_words, max_value);
    DecEscape r;
    r.s = s;
    r.value = value;
    return r;
}

// The block rebuilds the LUT of one CDF row: bucket j covers cumulative values [16 j, 16 j + 16).
//   lut32[j] = start << 16 | (freq - 1) of the symbol holding the whole bucket, kLutFlag when a symbol boundary falls inside
//   lutv[j]  = index of the symbol holding 16 j (saturated at 255): the decoded value, or where the flagged walk starts
//   symtab[k] = start << 16 | (freq - 1) of symbol k < 256: resolves a flagged bucket with two shared-memory loads
__device__ __forceinline__ void build_row_lut(const int32_t *__restrict__ crow, int n_sym, uint32_t *lut32, uint8_t *lutv,
                                              uint32_t *symtab) {
    for (int k = threadIdx.x; k < kSymTab; k += blockDim.x) {
        uint32_t e = kLutFlag;
        if (k < n_sym) {
            const int32_t a = __ldg(crow + k), b = __ldg(crow + k + 1);
            e = (static_cast<uint32_t>(a) << 16) | static_cast<uint32_t>(b - a - 1);
        }
        symtab[k] = e;
    }
    int k = 0;
    int32_t c0 = __ldg(crow), c1 = __ldg(crow + 1);
    for (uint32_t j = threadIdx.x; j < kLutBuckets; j += blockDim.x) {
        const int32_t lo = static_cast<int32_t>(j << 4);
        while (c1 <= lo) {  // the row ends at 65536 > lo: terminates
            ++k;
            c0 = c1;
            c1 = __ldg(crow + k + 1);
        }
        const bool whole = c1 >= lo + 16 && k <= 255;
        lut32[j] = whole ? ((static_cast<uint32_t>(c0) << 16) | static_cast<uint32_t>(c1 - c0 - 1)) : kLutFlag;
        lutv[j] = static_cast<uint8_t>(k < 255 ? k : 255);
    }
}

template <bool SYM, bool VAL>
__global__ void __launch_bounds__(kMaxWarps * 32)
rans_decode_lanes_kernel(const uint8_t *__restrict__ packed, const int64_t *__restrict__ offsets, int batch, uint32_t n,
                         uint32_t spatial, const void *__restrict__ tables, int32_t *__restrict__ out_symbols,
                         float *__restrict__ out_values, const float *__restrict__ means, int32_t *__restrict__ status,
                         const TraceSink trace) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int lane = threadIdx.x & 31;
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long trace_t0 = trace.buf ? trace_now() : 0ull;
    const Tables t = view(tables);
    uint32_t *lut32 = reinterpret_cast<uint32_t *>(smem_raw);
    uint8_t *lutv = reinterpret_cast<uint8_t *>(lut32 + kLutBuckets);
    // (volatile: computed ONCE; left to itself the compiler rematerialises the shared window base at every use)
    uint32_t lut32_s;
    asm volatile("mov.u32 %0, %1;" : "=r"(lut32_s) : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(lut32))));
    const uint32_t lutv_s = lut32_s + kLutBuckets * 4u;
    uint32_t *symtab = reinterpret_cast<uint32_t *>(lutv + kLutBuckets);
    const uint32_t symtab_s = lutv_s + kLutBuckets;

    bool active = b < batch;
    const uint32_t *words = nullptr;
    uint32_t n_words = 0;
    if (active) {
        const int64_t off = offsets[b];
        const int64_t n_bytes = offsets[b + 1] - off;
        if (n_bytes < 8 || (n_bytes & 3) || (off & 3)) {
            atomicOr(status, SC2_FAULT_BAD_STREAM);
            active = false;
        } else {
            words = reinterpret_cast<const uint32_t *>(packed + off);
            n_words = static_cast<uint32_t>(n_bytes >> 2);
        }
    }
    DecState s;
    s.xl = s.xh = s.next_w = 0u;
    s.p = 2;
    if (active) {
        s.xl = __ldg(words);
        s.xh = __ldg(words + 1);
        s.next_w = n_words > 2u ? __ldg(words + 2) : 0u;
    }
    int32_t *osym = SYM ? out_symbols + static_cast<int64_t>(active ? b : 0) * n : nullptr;
    float *oval = VAL ? out_values + static_cast<int64_t>(active ? b : 0) * n : nullptr;
    const bool vec_out = (!SYM || (reinterpret_cast<uintptr_t>(osym) & 15u) == 0u) && (!VAL || (reinterpret_cast<uintptr_t>(oval) & 15u) == 0u);

    uint32_t done = 0;
    for (int row = 0; done < n; ++row) {
        const uint32_t row_n = (n - done) < spatial ? (n - done) : spatial;
        const int32_t *crow = t.dec + static_cast<int64_t>(row) * t.dec_stride;
        __syncthreads();  // every warp has left the previous row's LUT (all streams of a block walk the rows together)
        build_row_lut(crow, __ldg(t.sizes + row) - 1, lut32, lutv, symtab);
        __syncthreads();
        if (active) {
            const int32_t max_value = __ldg(t.sizes + row) - 2;
            const int32_t offset = __ldg(t.offsets + row);
            const float mean = means ? __ldg(means + row) : 0.0f;
            // one symbol -> its value relative to the row offset
            auto step = [&]() -> int32_t {
                const uint32_t cum = s.xl & 0xffffu;
                uint32_t ent = lds_u32(lut32_s + ((cum >> 4) << 2));
                int32_t value = static_cast<int32_t>(lds_u8(lutv_s + (cum >> 4)));
                if (ent == kLutFlag && value < kSymTab - 1) {
                    // a symbol boundary inside the bucket (some lane of the warp hits one in a quarter of the steps): the
                    // symbol holding the bucket's first value, or the next one
                    const uint32_t e0 = lds_u32(symtab_s + 4u * value), e1 = lds_u32(symtab_s + 4u * value + 4u);
                    const bool up = cum >= (e1 >> 16);
                    const uint32_t e = up ? e1 : e0;
                    if (cum - (e >> 16) <= (e & 0xffffu)) {  // (else: two boundaries in one bucket -> the walk below)
                        ent = e;
                        value += up ? 1 : 0;
                    }
                }
                if (__builtin_expect(ent == kLutFlag || value == max_value, 0)) {
                    const DecEscape r = dec_slow_step(s, words, n_words, crow, value, max_value);
                    s = r.s;
                    value = r.value;
                } else {
                    // x = freq * (x >> 16) + cum - start, then renormalise: straight-line code
                    const uint32_t start = ent >> 16, freq = (ent & 0xffffu) + 1u;
                    const uint64_t prod = static_cast<uint64_t>(freq) * __funnelshift_r(s.xl, s.xh, 16) + (cum - start);
                    const uint32_t nl = static_cast<uint32_t>(prod);
                    const uint32_t nh = static_cast<uint32_t>(prod >> 32) + freq * (s.xh >> 16);
                    const bool ren = (nh | (nl >> 31)) == 0u;  // x < 2^31
                    s.xl = ren ? s.next_w : nl;
                    s.xh = ren ? nl : nh;
                    // predicated refill of next_w, written so that no branch (and no move that would wait for the
                    // load) lands on the chain: the loaded word is first needed at the next renormalisation
                    // (tried: taking next_w from a second register and merging the requested word one step later -- 11.7 -> 12.9 ms)
                    s.p += ren ? 1u : 0u;
                    const uint32_t *wp = words + min(s.p, n_words - 1u);
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p ld.global.nc.u32 %0, [%1];\n\t}"
                        : "+r"(s.next_w)
                        : "l"(wp), "r"(ren ? 1u : 0u));
                }
                return value + offset;
            };
            auto put = [&](uint32_t o, int32_t v) {
                if (SYM) osym[o] = v;
                if (VAL) oval[o] = static_cast<float>(v) + mean;
            };
            // scalar head up to a 16-byte boundary of the output, groups of four with ONE 16-byte store per output (the 32
            // lanes write 32 different lines: one L1 cycle per lane and store, whatever its width), scalar tail
            uint32_t i = 0;
            const uint32_t head = vec_out ? min(row_n, (4u - (done & 3u)) & 3u) : row_n;
            for (; i < head; ++i) put(done + i, step());
            for (; i + 4 <= row_n; i += 4) {
                {   // Touch the stream one 32-byte sector ahead (a load whose result is never read does not stall).  next_w
                    // shares its register with the other 31 lanes: whenever ANY lane's refill misses L1, the whole warp
                    // waits for L2 at its next step -- with this the refills hit L1.
                    uint32_t unused;
                    asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(unused) : "l"(words + min(s.p + 8u, n_words - 1u)));
                }
                const int32_t v0 = step(), v1 = step(), v2 = step(), v3 = step();
                if (SYM) *reinterpret_cast<int4 *>(osym + done + i) = make_int4(v0, v1, v2, v3);
                if (VAL)
                    *reinterpret_cast<float4 *>(oval + done + i) = make_float4(static_cast<float>(v0) + mean, static_cast<float>(v1) + mean,
                                                                              static_cast<float>(v2) + mean, static_cast<float>(v3) + mean);
            }
            for (; i < row_n; ++i) put(done + i, step());
        }
        done += row_n;
    }
    // a well-formed stream is consumed exactly; reading past the end (zero-filled) means it was truncated
    if (active && s.p > n_words) atomicOr(status, SC2_FAULT_STREAM_TRUNCATED);
    __syncwarp();
    if (lane == 0) trace_emit(trace, TRACE_RANS_DECODE, trace_t0, batch);
}

}  // namespace

// Streams of one launch are packed into as few blocks as possible (up to 8 warps = 256 streams per block): a coder block
// runs for milliseconds and, measured with the per-CTA trace, keeps the persistent convolution CTAs of the other batches
// off its SM most of that time -- one lost SM per launch instead of eight.
static int lanes_block_threads(int batch) {
    static const int max_warps = [] {
        const char *e = std::getenv("SC2_CODER_WARPS");
        const int v = e ? std::atoi(e) : 0;
        return (v >= 1 && v <= kMaxWarps) ? v : kMaxWarps;
    }();
    const int warps = (batch + 31) / 32;
    return 32 * (warps < max_warps ? warps : max_warps);
}

bool rans_use_lanes(int layout) {
    if (layout == SC2_RANS_LANE_PER_STREAM) return true;
    if (layout == SC2_RANS_WARP_PER_STREAM) return false;
    static const bool env_lanes = [] {
        const char *e = std::getenv("SC2_CODER");
        return e && e[0] == 'l';  // SC2_CODER=lanes; default (and SC2_CODER=warp): warp per stream
    }();
    return env_lanes;
}

int launch_rans_encode_lanes(const int32_t *symbols, int batch, int64_t n, int64_t spatial, const void *tables, int n_rows,
                             int cdf_stride, uint8_t *arena, int64_t slot_bytes, int32_t *lengths, int32_t *status,
                             cudaStream_t st) {
    const size_t table_bytes = static_cast<size_t>(n_rows) * cdf_stride * 16;
    const size_t smem = table_bytes <= static_cast<size_t>(kEncStageLimit) ? table_bytes : 0;
    const int threads = lanes_block_threads(batch);
    rans_encode_lanes_kernel<<<(batch + threads - 1) / threads, threads, smem, st>>>(symbols, batch, static_cast<uint32_t>(n),
                                                                  static_cast<uint32_t>(spatial), tables, arena, slot_bytes,
                                                                  lengths, status, trace_sink());
    SC2_LAUNCH_CHECK("rans_encode_lanes_kernel");
    return SC2_OK;
}

int launch_rans_decode_lanes(const uint8_t *packed, const int64_t *offsets, int batch, int64_t n, int64_t spatial,
                             const void *tables, int32_t *out_symbols, float *out_values, const float *means,
                             int32_t *status, cudaStream_t st) {
    const size_t smem = kLutBuckets * 4 + kLutBuckets + kSymTab * 4;
    const int threads = lanes_block_threads(batch);
    const int grid = (batch + threads - 1) / threads;
    const uint32_t un = static_cast<uint32_t>(n), us = static_cast<uint32_t>(spatial);
    if (out_symbols && out_values)
        rans_decode_lanes_kernel<true, true><<<grid, threads, smem, st>>>(packed, offsets, batch, un, us, tables, out_symbols, out_values, means, status, trace_sink());
    else if (out_symbols)
        rans_decode_lanes_kernel<true, false><<<grid, threads, smem, st>>>(packed, offsets, batch, un, us, tables, out_symbols, out_values, means, status, trace_sink());
    else if (out_values)
        rans_decode_lanes_kernel<false, true><<<grid, threads, smem, st>>>(packed, offsets, batch, un, us, tables, out_symbols, out_values, means, status, trace_sink());
    else
        return SC2_ERR_INVALID_ARG;
    SC2_LAUNCH_CHECK("rans_decode_lanes_kernel");
    return SC2_OK;
}

}  // namespace sc2
