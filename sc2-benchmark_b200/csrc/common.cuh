// common.cuh -- shared helpers for libsc2b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

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

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: a process that drives several GPUs
// (nn.DataParallel replicas, a model moved from cuda:0 to cuda:1) must raise it on each of them.  `done` (one static per
// call site) remembers the device ordinals already configured; safe to call from several host threads.
template <typename Kernel>
static inline int ensure_dyn_smem(Kernel *kernel, int bytes, std::atomic<uint64_t> &done) {
    int dev = 0;
    SC2_CUDA_TRY(cudaGetDevice(&dev));
    const uint64_t bit = 1ull << (dev & 63);
    if (dev < 64 && (done.load(std::memory_order_acquire) & bit)) return SC2_OK;
    SC2_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    if (dev < 64) done.fetch_or(bit, std::memory_order_release);
    return SC2_OK;
}

constexpr int kNumSMs = 148;  // B200

// ---- diagnostics: per-CTA trace (sc2_trace_start / sc2_trace_stop) -------------------------------------------------
// When a trace buffer is installed every CTA of the long-running kernels appends {start, end (globaltimer ns), kernel
// kind, SM id, aux (tiles processed / streams), block id}: which SM ran what, when -- the only way to see how kernels of
// different streams share the machine.  Off (null sink): one predictable branch per CTA.
struct TraceRec {
    unsigned long long t0, t1;
    int kind, smid, aux, bid;
};
struct TraceSink {
    TraceRec *buf;
    unsigned *count;
    unsigned cap;
};
TraceSink trace_sink();  // host side: the installed sink ({nullptr, nullptr, 0} when tracing is off)
enum TraceKind { TRACE_CONV_TC = 1, TRACE_CONV_SPLIT = 2, TRACE_CONV_FIRST = 3, TRACE_RANS_ENCODE = 4, TRACE_RANS_DECODE = 5 };

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trace_emit(const TraceSink &ts, int kind, unsigned long long t0, int aux) {
    if (ts.buf == nullptr) return;
    const unsigned i = atomicAdd(ts.count, 1u);
    if (i >= ts.cap) return;
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    TraceRec r;
    r.t0 = t0;
    r.t1 = trace_now();
    r.kind = kind;
    r.smid = static_cast<int>(smid);
    r.aux = aux;
    r.bid = static_cast<int>(blockIdx.x);
    ts.buf[i] = r;
}
#endif

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

// latency-tuned fast paths (rans_fast.cu): the encoder for channel rows (indexes == nullptr) and per-element rows, the decoder for
// channel rows
int launch_rans_encode_fast(const int32_t *symbols, const int32_t *indexes, int batch, int64_t n, int64_t spatial, const void *tables,
                            int n_rows, int cdf_stride, uint8_t *arena, int64_t slot_bytes, int32_t *lengths, int32_t *status,
                            cudaStream_t st);
int launch_rans_decode_fast(const uint8_t *packed, const int64_t *offsets, int batch, int64_t n, int64_t spatial,
                            const void *tables, int n_rows, int cdf_stride, int32_t *out_symbols, float *out_values,
                            const float *means, int32_t *status, cudaStream_t st);

// channel-mode, one lane per stream (rans_lanes.cu): the default; SC2_CODER=warp selects rans_fast.cu
bool rans_use_lanes(int layout);
int launch_rans_encode_lanes(const int32_t *symbols, int batch, int64_t n, int64_t spatial, const void *tables, int n_rows,
                             int cdf_stride, uint8_t *arena, int64_t slot_bytes, int32_t *lengths, int32_t *status,
                             cudaStream_t st);
int launch_rans_decode_lanes(const uint8_t *packed, const int64_t *offsets, int batch, int64_t n, int64_t spatial,
                             const void *tables, int32_t *out_symbols, float *out_values, const float *means,
                             int32_t *status, cudaStream_t st);

}  // namespace sc2
