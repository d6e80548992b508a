// common.cuh -- shared helpers for libsc2b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sc2b200.h"

namespace sc2 {

// Records the last CUDA error text for sc2_last_cuda_error().
int cuda_fail(cudaError_t e, const char *where);

#define SC2_CUDA_TRY(expr)                                         \
    do {                                                           \
        cudaError_t _e = (expr);                                   \
        if (_e != cudaSuccess) return ::sc2::cuda_fail(_e, #expr); \
    } while (0)

#define SC2_LAUNCH_CHECK(name)                                       \
    do {                                                             \
        cudaError_t _e = cudaGetLastError();                         \
        if (_e != cudaSuccess) return ::sc2::cuda_fail(_e, name);    \
    } while (0)

static inline cudaStream_t as_stream(sc2_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;  // B200

// ---- coder table blob (built on the host by sc2_rans_build_tables) ------------------------------
struct RansTableHeader {
    int32_t magic;       // 'S2RT'
    int32_t n_rows;
    int32_t cdf_stride;  // entries per row in the caller's _quantized_cdf
    int32_t dec_stride;  // entries per row of the sentinel-padded decoder copy (multiple of 32)
    int32_t meta_off;    // byte offsets from the blob start
    int32_t enc_off;
    int32_t dec_off;
    int32_t total_bytes;
};
constexpr int32_t kRansMagic = 0x54523253;
constexpr int kRansPrecision = 16;
constexpr int kBypassPrecision = 4;
constexpr uint32_t kMaxBypassVal = 15;

// Encoder entry, 16 bytes: exact division by `freq` through a 64-bit reciprocal
// (q = mulhi64(x, rcp) >> shift, exact for x < 2^63).
struct __align__(16) RansEncEntry {
    uint32_t rcp_lo, rcp_hi;
    uint32_t bias_shift;  // bits 0..16 bias (start, or start + 65535 when freq == 1); bits 24..27 shift
    uint32_t freq;        // 1..65535
};

// channel-mode fast paths (rans_fast.cu)
int launch_rans_encode_fast(const int32_t *symbols, int batch, int64_t n, int64_t spatial, const void *tables, int n_rows,
                            int cdf_stride, uint8_t *arena, int64_t slot_bytes, int32_t *lengths, int32_t *status,
                            cudaStream_t st);
int launch_rans_decode_fast(const uint8_t *packed, const int64_t *offsets, int batch, int64_t n, int64_t spatial,
                            const void *tables, int n_rows, int cdf_stride, int32_t *out_symbols, float *out_values,
                            const float *means, int32_t *status, cudaStream_t st);

}  // namespace sc2
