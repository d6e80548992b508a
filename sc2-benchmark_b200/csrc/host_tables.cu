// host_tables.cu -- host-side pieces of the C ABI: CDF quantisation and coder-table construction.
// They run once per update() (sc2bench/models/layer.py:431-441 -> CompressionModel.update), never per image.
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace sc2 {

static thread_local std::string g_last_cuda_error;

int cuda_fail(cudaError_t e, const char *where) {
    g_last_cuda_error = std::string(where) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
    return SC2_ERR_CUDA;
}

}  // namespace sc2

namespace sc2 {
static std::atomic<int> g_persistent_ctas{0};  // 0: not set (environment variable or one CTA per SM)
int persistent_grid() {
    const int set = g_persistent_ctas.load(std::memory_order_relaxed);
    if (set > 0 && set <= kNumSMs) return set;
    static const int from_env = [] {
        const char *e = std::getenv("SC2_TC_GRID");
        const int v = e ? std::atoi(e) : 0;
        return (v > 0 && v < kNumSMs) ? v : kNumSMs;
    }();
    return from_env;
}
}  // namespace sc2

namespace sc2 {
static TraceSink g_trace = {nullptr, nullptr, 0};
TraceSink trace_sink() { return g_trace; }
}  // namespace sc2

extern "C" {

int sc2_abi_version(void) { return SC2_ABI_VERSION; }

int sc2_set_persistent_ctas(int ctas) {
    if (ctas < 0 || ctas > sc2::kNumSMs) return SC2_ERR_INVALID_ARG;
    sc2::g_persistent_ctas.store(ctas, std::memory_order_relaxed);
    return SC2_OK;
}

int sc2_trace_start(void *device_buffer, int64_t bytes) {
    if (!device_buffer || bytes < 16 + static_cast<int64_t>(sizeof(sc2::TraceRec))) return SC2_ERR_INVALID_ARG;
    sc2::g_trace.count = static_cast<unsigned *>(device_buffer);
    sc2::g_trace.buf = reinterpret_cast<sc2::TraceRec *>(static_cast<uint8_t *>(device_buffer) + 16);
    sc2::g_trace.cap = static_cast<unsigned>((bytes - 16) / static_cast<int64_t>(sizeof(sc2::TraceRec)));
    return SC2_OK;
}

int sc2_trace_stop(void) {
    sc2::g_trace = sc2::TraceSink{nullptr, nullptr, 0};
    return SC2_OK;
}

const char *sc2_error_string(int code) {
    switch (code) {
        case SC2_OK: return "ok";
        case SC2_ERR_INVALID_ARG: return "invalid argument";
        case SC2_ERR_UNSUPPORTED: return "unsupported configuration";
        case SC2_ERR_CUDA: return "CUDA error";
        case SC2_ERR_DOMAIN: return "pmf outside the domain (negative, non-finite or all-zero)";
        default: return "unknown";
    }
}

const char *sc2_last_cuda_error(void) { return sc2::g_last_cuda_error.c_str(); }

// Quantises a float pmf to a strictly increasing `precision`-bit CDF the way CompressAI's
// pmf_to_quantized_cdf does: round each mass to an integer frequency, rescale the frequencies so they
// add up to 2^precision with truncating integer division, prefix-sum, pin the last entry, then repair
// every zero-width bin by taking one count from the narrowest bin that can spare it.
int sc2_pmf_to_quantized_cdf(const float *pmf, int n, int precision, uint32_t *cdf_out) {
    if (!pmf || !cdf_out || n < 1 || precision < 1 || precision > 16) return SC2_ERR_INVALID_ARG;
    for (int i = 0; i < n; ++i)
        if (!(pmf[i] >= 0.0f) || !std::isfinite(pmf[i])) return SC2_ERR_DOMAIN;
    const float scale = static_cast<float>(1 << precision);
    std::vector<uint32_t> freq(static_cast<size_t>(n));
    int32_t total = 0;  // CompressAI accumulates in `int`
    for (int i = 0; i < n; ++i) {
        freq[i] = static_cast<uint32_t>(std::round(pmf[i] * scale));
        total += static_cast<int32_t>(freq[i]);
    }
    if (total == 0) return SC2_ERR_DOMAIN;
    const uint64_t target = 1ull << precision;
    uint32_t running = 0;
    cdf_out[0] = 0;
    for (int i = 0; i < n; ++i) {
        running += static_cast<uint32_t>((target * freq[i]) / static_cast<uint32_t>(total));
        cdf_out[i + 1] = running;
    }
    cdf_out[n] = static_cast<uint32_t>(target);
    for (int i = 0; i < n; ++i) {
        if (cdf_out[i] != cdf_out[i + 1]) continue;
        int donor = -1;
        uint32_t donor_width = ~0u;
        for (int j = 0; j < n; ++j) {
            const uint32_t width = cdf_out[j + 1] - cdf_out[j];
            if (width > 1 && width < donor_width) {
                donor_width = width;
                donor = j;
            }
        }
        if (donor < 0) return SC2_ERR_DOMAIN;
        if (donor < i) {
            for (int j = donor + 1; j <= i; ++j) --cdf_out[j];
        } else {
            for (int j = i + 1; j <= donor; ++j) ++cdf_out[j];
        }
    }
    return SC2_OK;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

size_t sc2_rans_table_bytes(int n_rows, int cdf_stride) {
    if (n_rows < 1 || cdf_stride < 2) return 0;
    const size_t meta = static_cast<size_t>(round_up(2 * n_rows * 4, 16));
    const size_t enc = static_cast<size_t>(n_rows) * cdf_stride * sizeof(sc2::RansEncEntry);
    const size_t dec = static_cast<size_t>(n_rows) * round_up(cdf_stride, 32) * 4;
    return sizeof(sc2::RansTableHeader) + meta + enc + dec;
}

int sc2_rans_build_tables(const int32_t *cdfs, const int32_t *cdf_sizes, const int32_t *offsets, int n_rows,
                          int cdf_stride, void *blob_out) {
    if (!cdfs || !cdf_sizes || !offsets || !blob_out || n_rows < 1 || cdf_stride < 2) return SC2_ERR_INVALID_ARG;
    const size_t total = sc2_rans_table_bytes(n_rows, cdf_stride);
    if (total > 0x7fffffffull) return SC2_ERR_UNSUPPORTED;
    auto *blob = static_cast<uint8_t *>(blob_out);
    std::memset(blob, 0, total);
    sc2::RansTableHeader h;
    h.magic = sc2::kRansMagic;
    h.n_rows = n_rows;
    h.cdf_stride = cdf_stride;
    h.dec_stride = round_up(cdf_stride, 32);
    h.meta_off = static_cast<int32_t>(sizeof(sc2::RansTableHeader));
    h.enc_off = h.meta_off + round_up(2 * n_rows * 4, 16);
    h.dec_off = h.enc_off + static_cast<int32_t>(static_cast<size_t>(n_rows) * cdf_stride * sizeof(sc2::RansEncEntry));
    h.total_bytes = static_cast<int32_t>(total);
    std::memcpy(blob, &h, sizeof(h));
    auto *meta = reinterpret_cast<int32_t *>(blob + h.meta_off);
    auto *enc = reinterpret_cast<sc2::RansEncEntry *>(blob + h.enc_off);
    auto *dec = reinterpret_cast<int32_t *>(blob + h.dec_off);
    for (int r = 0; r < n_rows; ++r) {
        const int size = cdf_sizes[r];
        if (size < 2 || size > cdf_stride) return SC2_ERR_INVALID_ARG;
        const int32_t *row = cdfs + static_cast<size_t>(r) * cdf_stride;
        if (row[0] != 0 || row[size - 1] != (1 << sc2::kRansPrecision)) return SC2_ERR_INVALID_ARG;
        meta[r] = size;
        meta[n_rows + r] = offsets[r];
        for (int v = 0; v < h.dec_stride; ++v) dec[static_cast<size_t>(r) * h.dec_stride + v] = v < size ? row[v] : 0x7fffffff;
        for (int v = 0; v + 1 < size; ++v) {
            const uint32_t start = static_cast<uint32_t>(row[v]);
            const int64_t width = static_cast<int64_t>(row[v + 1]) - row[v];
            if (width < 1 || width > 65535) return SC2_ERR_INVALID_ARG;  // must be strictly increasing
            const uint32_t freq = static_cast<uint32_t>(width);
            sc2::RansEncEntry e;
            e.freq = freq;
            if (freq == 1) {
                // x / 1: mulhi(x, 2^64 - 1) = x - 1 for x > 0, compensated by the bias
                e.rcp_lo = e.rcp_hi = 0xffffffffu;
                e.bias_shift = start + 65535u;
            } else {
                uint32_t shift = 0;
                while (freq > (1u << shift)) ++shift;
                const unsigned __int128 num = (static_cast<unsigned __int128>(1) << (shift + 63)) + freq - 1;
                const uint64_t rcp = static_cast<uint64_t>(num / freq);
                e.rcp_lo = static_cast<uint32_t>(rcp);
                e.rcp_hi = static_cast<uint32_t>(rcp >> 32);
                e.bias_shift = start | ((shift - 1) << 24);
            }
            enc[static_cast<size_t>(r) * cdf_stride + v] = e;
        }
    }
    return SC2_OK;
}

int64_t sc2_rans_max_stream_bytes(int64_t n_symbols) {
    // regular symbol: <= 16 bits; escape: one 4-bit count digit + up to 8 nibbles = 36 more bits.
    // The state is flushed as 64 bits and words are 32 bits; add slack for the partial word.
    const int64_t bits = n_symbols * (16 + 36);
    return 4 * ((bits + 31) / 32 + 3);
}

}  // extern "C"
