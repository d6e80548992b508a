// tc_common.cuh -- PTX wrappers shared by the tcgen05 kernels (mbarrier, TMA, TMEM, UMMA descriptors).
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.cuh"

namespace sc2 {
namespace tc {

constexpr int kBlockK = 64;        // fp16 elements per 128-byte swizzle row
constexpr int kTileM = 128;        // UMMA_M
constexpr int kABytes = kTileM * 128;

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifdef SC2_HANG_DEBUG
// diagnostics build (SC2_NVCC_DEFINES=-DSC2_HANG_DEBUG): a wait that spins for ~seconds reports who waits on what and traps
static __device__ __noinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (unsigned long long spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if (spins > (1ull << 22)) {
            if ((threadIdx.x & 31) == 0)
                printf("HANG block %d warp %d waits on barrier at smem 0x%x parity %u\n", blockIdx.x, threadIdx.x >> 5, smem_u32(bar), parity);
            __nanosleep(1000000);
            if (spins > (1ull << 22) + 4) __trap();
        }
    }
}
#else
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
#endif
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "elect.sync _|P, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}

// ---- tile scheduler of the persistent kernels -----------------------------------------------------------------------
// Static: CTA b walks tiles b, b + gridDim.x, ...  A persistent kernel scheduled that way takes TWICE as long as soon as
// one of its CTAs cannot be placed next to blocks of another stream (the coder kernels run for milliseconds), because the
// late CTA still owns 1/gridDim of the tiles.  Dynamic (counter != nullptr): one thread of the CTA claims tiles from a
// global counter (zero at launch) and publishes them to the other warps through a small shared-memory queue; a CTA that
// starts late simply finds no work.
constexpr int kTileQ = 4;
constexpr int kTileSchedBytes = 2 * kTileQ * 8 + kTileQ * 4;  // full[], empty[] mbarriers + tile ids

struct TileSched {
    uint64_t *qfull, *qempty;
    int *qtile;
    int *counter;
    int total;

    __device__ __forceinline__ void bind(uint8_t *smem_at, int *ctr, int total_tiles) {  // smem_at: 8-byte aligned
        qfull = reinterpret_cast<uint64_t *>(smem_at);
        qempty = qfull + kTileQ;
        qtile = reinterpret_cast<int *>(qempty + kTileQ);
        counter = ctr;
        total = total_tiles;
    }
    __device__ __forceinline__ void init(uint32_t consumer_warps) const {  // one thread, before fence_barrier_init
        for (int i = 0; i < kTileQ; ++i) {
            mbar_init(&qfull[i], 1);
            mbar_init(&qempty[i], consumer_warps);
        }
    }
    // producer (ONE thread): the n-th tile of this CTA, -1 when there is none
    __device__ __forceinline__ int claim(uint32_t n) const {
        if (counter == nullptr) {
            const int64_t t = static_cast<int64_t>(blockIdx.x) + static_cast<int64_t>(n) * gridDim.x;
            return t < total ? static_cast<int>(t) : -1;
        }
        const int t = atomicAdd(counter, 1);
        return t < total ? t : -1;
    }
    __device__ __forceinline__ void publish(uint32_t n, int tile) const {
        if (counter == nullptr) return;
        const uint32_t slot = n % kTileQ, ph = (n / kTileQ) & 1u;
        mbar_wait(&qempty[slot], ph ^ 1u);
        *reinterpret_cast<volatile int *>(qtile + slot) = tile;
        mbar_arrive(&qfull[slot]);
    }
    // consumer: every lane of a consumer warp calls it (converged); -1 ends the CTA's tile loop
    __device__ __forceinline__ int next(uint32_t n, int lane) const {
        if (counter == nullptr) return claim(n);
        const uint32_t slot = n % kTileQ, ph = (n / kTileQ) & 1u;
        mbar_wait(&qfull[slot], ph);
        const int t = *reinterpret_cast<volatile int *>(qtile + slot);
        __syncwarp();
        if (lane == 0) mbar_arrive(&qempty[slot]);
        return t;
    }
};

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *m, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap *m, const void *src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all previously committed bulk stores have finished READING shared memory (the staging buffer may be reused)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_commit_and_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- the same two instructions for a CONVERGENT issuer warp (round 2) ------------------------------------------------
// The MMA-issuing warp of a kernel with many short MMAs (the split-fp16 kernels: 300+ per tile, N = 48..192) is bound by its
// own instruction stream: `if (elect_one()) { build two 64-bit descriptors; mma }` per tap cost ~950 cycles per (tap, chunk)
// (ncu source view, profiles/r2g_*).  Here every lane runs the same straight-line code with uniform values, only the
// tcgen05 instruction is predicated on the leader lane chosen ONCE, and a descriptor is (lo word = address >> 4, constant
// hi word), so advancing K or shifting rows is one 32-bit add.
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024, version 1, SWIZZLE_128B
constexpr uint32_t kDescHiSw64 = (512u >> 4) | (1u << 14) | (4u << 29);    // SBO 512, version 1, SWIZZLE_64B
__device__ __forceinline__ void umma_f16_lead(bool lead, uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.ne.b32 q, %6, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
        "}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate), "r"(static_cast<uint32_t>(lead)) : "memory");
}
__device__ __forceinline__ void umma_commit_lead(bool lead, uint64_t *bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "setp.ne.b32 q, %1, 0;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(static_cast<uint32_t>(lead)) : "memory");
}

// All MMAs of one (tap, K chunk) of the split-fp16 kernels in ONE asm block: KSTEPS x { D[stack] += A_hi . [B_hi; B_lo],
// D[lohi] += A_lo . B_hi }.  One predicate set-up and three descriptor registers for the whole block (advancing K by 16 elements
// is a 64-bit add of 2 = 32 bytes >> 4) instead of two predicate set-ups and two descriptor packs per MMA.
#define SC2_UMMA_SPLIT_TAP(NAME, BODY)                                                                                              \
    __device__ __forceinline__ void NAME(bool lead, uint32_t d_stack, uint32_t d_lohi, uint32_t a_hi, uint32_t a_lo, uint32_t b,    \
                                         uint32_t desc_hi, uint32_t idesc_stack, uint32_t idesc_n, uint32_t acc0) {                 \
        asm volatile(BODY ::"r"(d_stack), "r"(d_lohi), "r"(a_hi), "r"(a_lo), "r"(b), "r"(desc_hi), "r"(idesc_stack), "r"(idesc_n),   \
                     "r"(acc0), "r"(static_cast<uint32_t>(lead)) : "memory");                                                        \
    }
SC2_UMMA_SPLIT_TAP(umma_split_tap1, \
        "{\n\t" \
        ".reg .pred p, q, t;\n\t" \
        ".reg .b64 dah, dal, db;\n\t" \
        "setp.ne.b32 p, %8, 0;\n\t" \
        "setp.ne.b32 q, %9, 0;\n\t" \
        "setp.eq.b32 t, 0, 0;\n\t" \
        "mov.b64 dah, {%2, %5};\n\t" \
        "mov.b64 dal, {%3, %5};\n\t" \
        "mov.b64 db, {%4, %5};\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, p;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "}")
SC2_UMMA_SPLIT_TAP(umma_split_tap2, \
        "{\n\t" \
        ".reg .pred p, q, t;\n\t" \
        ".reg .b64 dah, dal, db;\n\t" \
        "setp.ne.b32 p, %8, 0;\n\t" \
        "setp.ne.b32 q, %9, 0;\n\t" \
        "setp.eq.b32 t, 0, 0;\n\t" \
        "mov.b64 dah, {%2, %5};\n\t" \
        "mov.b64 dal, {%3, %5};\n\t" \
        "mov.b64 db, {%4, %5};\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, p;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "add.s64 dah, dah, 2;\n\t" \
        "add.s64 dal, dal, 2;\n\t" \
        "add.s64 db, db, 2;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, t;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "}")
SC2_UMMA_SPLIT_TAP(umma_split_tap3, \
        "{\n\t" \
        ".reg .pred p, q, t;\n\t" \
        ".reg .b64 dah, dal, db;\n\t" \
        "setp.ne.b32 p, %8, 0;\n\t" \
        "setp.ne.b32 q, %9, 0;\n\t" \
        "setp.eq.b32 t, 0, 0;\n\t" \
        "mov.b64 dah, {%2, %5};\n\t" \
        "mov.b64 dal, {%3, %5};\n\t" \
        "mov.b64 db, {%4, %5};\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, p;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "add.s64 dah, dah, 2;\n\t" \
        "add.s64 dal, dal, 2;\n\t" \
        "add.s64 db, db, 2;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, t;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "add.s64 dah, dah, 2;\n\t" \
        "add.s64 dal, dal, 2;\n\t" \
        "add.s64 db, db, 2;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, t;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "}")
SC2_UMMA_SPLIT_TAP(umma_split_tap4, \
        "{\n\t" \
        ".reg .pred p, q, t;\n\t" \
        ".reg .b64 dah, dal, db;\n\t" \
        "setp.ne.b32 p, %8, 0;\n\t" \
        "setp.ne.b32 q, %9, 0;\n\t" \
        "setp.eq.b32 t, 0, 0;\n\t" \
        "mov.b64 dah, {%2, %5};\n\t" \
        "mov.b64 dal, {%3, %5};\n\t" \
        "mov.b64 db, {%4, %5};\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, p;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "add.s64 dah, dah, 2;\n\t" \
        "add.s64 dal, dal, 2;\n\t" \
        "add.s64 db, db, 2;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, t;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "add.s64 dah, dah, 2;\n\t" \
        "add.s64 dal, dal, 2;\n\t" \
        "add.s64 db, db, 2;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, t;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "add.s64 dah, dah, 2;\n\t" \
        "add.s64 dal, dal, 2;\n\t" \
        "add.s64 db, db, 2;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, db, %6, t;\n\t" \
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dal, db, %7, t;\n\t" \
        "}")
#undef SC2_UMMA_SPLIT_TAP

// UMMA shared-memory descriptor: K-major operand, 128-byte swizzle, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3ffff) >> 4);        // start address >> 4            bits [0, 14)
    d |= static_cast<uint64_t>(0) << 16;                       // leading byte offset (unused)  bits [16, 30)
    d |= static_cast<uint64_t>(1024 >> 4) << 32;               // stride byte offset = 1024     bits [32, 46)
    d |= static_cast<uint64_t>(1) << 46;                       // descriptor version (sm_100)   bits [46, 48)
    d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B                  bits [61, 64)
    return d;
}

// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, A and B K-major, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | (static_cast<uint32_t>(n >> 3) << 17) |
           (static_cast<uint32_t>(kTileM >> 4) << 24);
}

// ---- host: tensor maps -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

// NHWC tensor [batch, h, w, c] seen as 4-D {c, w, h, batch}; box {box_c, box_w, box_h, 1}, 128-byte swizzle.
inline int make_nhwc_map(CUtensorMap *m, const void *base, CUtensorMapDataType dt, int elem_bytes, int c, int w, int h, int batch,
                         int box_c, int box_w, int box_h, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SC2_ERR_CUDA;
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(batch)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(c) * elem_bytes, static_cast<cuuint64_t>(w) * c * elem_bytes,
                             static_cast<cuuint64_t>(h) * w * c * elem_bytes};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, dt, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SC2_OK : SC2_ERR_INVALID_ARG;
}

// channels [0, c) of an NHWC tensor whose pixels are `pitch_c` elements apart (one N tile of a wider tensor)
inline int make_nhwc_map_pitch(CUtensorMap *m, const void *base, CUtensorMapDataType dt, int elem_bytes, int c, int pitch_c, int w, int h,
                               int batch, int box_c, int box_w, int box_h, CUtensorMapSwizzle swizzle) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SC2_ERR_CUDA;
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>(w), static_cast<cuuint64_t>(h), static_cast<cuuint64_t>(batch)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(pitch_c) * elem_bytes, static_cast<cuuint64_t>(w) * pitch_c * elem_bytes,
                             static_cast<cuuint64_t>(h) * w * pitch_c * elem_bytes};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, dt, 4, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SC2_OK : SC2_ERR_INVALID_ARG;
}

// An fp16 NHWC tensor [batch, h, w, c] (h, w even) seen by a stride-2 convolution: dims (px * c + channel, X, py, Y, image) with
// pixel (2Y + py, 2X + px), so that a tap is a unit-stride box {box_c, box_w, 1, box_h, 1} of one pixel parity -- the parity
// planes of the bottleneck's g_a without the re-layout.
inline int make_nhwc_s2_map(CUtensorMap *m, const void *base, int c, int w, int h, int batch, int box_c, int box_w, int box_h) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SC2_ERR_CUDA;
    const cuuint64_t row = static_cast<cuuint64_t>(w) * c * 2;
    cuuint64_t dims[5] = {static_cast<cuuint64_t>(2 * c), static_cast<cuuint64_t>(w / 2), 2, static_cast<cuuint64_t>(h / 2),
                          static_cast<cuuint64_t>(batch)};
    cuuint64_t strides[4] = {static_cast<cuuint64_t>(2 * c) * 2, row, 2 * row, static_cast<cuuint64_t>(h) * row};
    cuuint32_t box[5] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), 1, static_cast<cuuint32_t>(box_h), 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SC2_OK : SC2_ERR_INVALID_ARG;
}

// Output pixels of ONE parity (py, px) of an NHWC tensor [batch, out_h, out_w, pitch_c] (a transposed convolution's sub-grid; out_h,
// out_w may be odd): dims (channel, X, Y, image) with pixel (2Y + py, 2X + px), extents exactly the pixels of that parity, from a base
// already offset to pixel (py, px); dense box {box_c, box_w, box_h, 1}.
inline int make_nhwc_parity_out_map(CUtensorMap *m, const void *base, int c, int pitch_c, int out_w, int out_h, int py, int px, int batch,
                                    int box_c, int box_w, int box_h) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SC2_ERR_CUDA;
    const cuuint64_t row = static_cast<cuuint64_t>(out_w) * pitch_c * 2;  // one full-resolution row, bytes
    cuuint64_t dims[4] = {static_cast<cuuint64_t>(c), static_cast<cuuint64_t>((out_w - px + 1) / 2), static_cast<cuuint64_t>((out_h - py + 1) / 2),
                          static_cast<cuuint64_t>(batch)};
    cuuint64_t strides[3] = {static_cast<cuuint64_t>(2 * pitch_c) * 2, 2 * row, static_cast<cuuint64_t>(out_h) * row};
    cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SC2_OK : SC2_ERR_INVALID_ARG;
}

// packed weights [rows_total, c_in_pad] fp16 (K contiguous); box {64, n_tile}
inline int make_weight_map(CUtensorMap *m, const void *base, int c_in_pad, int rows_total, int n_tile) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return SC2_ERR_CUDA;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(c_in_pad), static_cast<cuuint64_t>(rows_total)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(c_in_pad) * 2};
    cuuint32_t box[2] = {kBlockK, static_cast<cuuint32_t>(n_tile)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SC2_OK : SC2_ERR_INVALID_ARG;
}

}  // namespace tc
// Every persistent tensor-core kernel asks for the SAME amount of dynamic shared memory (when it needs no more than that).
// Shared memory is handed out as contiguous ranges: while a long-running coder block of another stream sits on the SM, the
// range a finished convolution CTA leaves behind can only be reused by a CTA that is not larger -- with kernels of
// different sizes following each other, the SM stays closed to them until the coder block is gone.
constexpr int kUniformSmem = 202 * 1024;
inline int uniform_smem(int needed) { return needed <= kUniformSmem ? kUniformSmem : needed; }

// CTAs of a persistent kernel: one per SM by default.  Fewer leave SMs to co-running kernels: with batches in flight the coder
// blocks of other streams (one SM each, 8-12 ms long) otherwise start only in the bubbles of the transform stream; with 8 SMs
// left to them the pipelined step is 4 % faster although every transform kernel is 4 % slower (profiles/r2w_*).
// sc2_set_persistent_ctas() (process-wide tuning knob, CodecPipeline sets it) or the SC2_TC_GRID environment variable.
int persistent_grid();
}  // namespace sc2
