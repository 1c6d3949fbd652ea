// spconv_tc.cu — K4b: sparse convolution as an output-stationary implicit GEMM on tcgen05 (sm_100a).
//
// PERSISTENT, warp-specialised kernel on CTA PAIRS (cta_group::2): a cluster of two CTAs (one TPC) walks the work
// list (256-row tile pair x n-tile) round-robin; each CTA gathers the A rows of its own 128-row tile and loads HALF
// of every weight tile, the leader CTA issues M = 256 MMAs that read both CTAs' shared memory, and each CTA's TMEM
// holds the accumulator of its own 128 rows. Halving the per-SM weight traffic is what lets the stage ring be deep
// enough (4 x 40 KB at n_tile = 384) to cover the DRAM latency of the gathered rows.
//
//   work item     2 x 128 output voxels (rows perm[256 t ..]) x n_tile output channels (n_tile <= 384 TMEM columns)
//   reduction     items = (kernel offset k with at least one neighbour in the tile) x (64-channel chunk)
//   A operand     128 gathered input rows x 64 bf16 (= one 128-byte swizzle row per voxel), cp.async 16-byte
//                 pieces straight into the SWIZZLE_128B K-major smem image, zero-fill for missing neighbours
//   B operand     W[k][chunk] pre-packed on the host side of the ABI into the exact smem image of each CTA's half,
//                 so one cp.async.bulk (UBLKCP) per item brings n_tile/2 x 128 bytes and completes on the stage
//                 mbarrier
//   MMA           one elected thread of the leader CTA issues tcgen05.mma.cta_group::2.kind::f16 (M=256, N<=256,
//                 K=16), fp32 accumulators stay in TMEM for the whole tile; tcgen05.commit multicasts the stage
//                 release / accumulator-ready arrivals to both CTAs
//   epilogue      tcgen05.ld -> folded BatchNorm scale/shift, residual add, ReLU/LeakyReLU -> bf16 rows, staged
//                 through shared memory so that residual loads and output stores are 64-byte coalesced segments
//
//   warps  0-3    gather producers (stage ring runs on across tiles, so the next tile's rows are in flight while
//                 the tensor pipe finishes the current one); prefetch.global.L2 of the next offset's rows
//          4      TMEM alloc + MMA issuer (leader) / stage-full relay to the leader's barrier (peer). The issue
//                 loop is warp-uniform: shfl-broadcast warp index / tile masks / TMEM base, one elect.sync branch
//                 per item, descriptors as 32-bit low words -> UTCHMMA operands live in uniform registers
//          5      weight (B) bulk-copy issuer
//          6-7    kernel-map prefetch: the NEXT tile's 128 x K neighbour rows go global -> registers while the
//                 current tile runs, then registers -> smem the moment the producers release the buffer
//          8-15   epilogue (warp w owns TMEM lanes 32 (w % 4) .. and the column half (w - 8) / 4); overlaps the
//                 next tile's gathers, and its MMAs when two accumulators fit TMEM (n_tile <= 256)
//
//   barriers      full[s]  (128 cp.async-completion arrivals + 1 expect_tx + the peer's relay on the leader)
//                 empty[s] (tcgen05.commit, multicast to both CTAs)
//                 nbr_full (2 prefetch warps)                        nbr_empty    (4 producer warps)
//                 tmem_full[2] (commit) / tmem_empty[2] (8 epilogue warps of each CTA, on the leader)
//
//   B2ME_TC_TMA=1 operands through the TMA unit instead: tile::gather4 copies of the gathered rows (absent neighbour =
//                 row -1 = out of bounds = zeros) and 2-D boxes of the packed weights, cta_group::2 with the LEADER's
//                 stage barrier as completion target - no relay, no proxy fence; same speed, kept opt-in
//
// The two sources (in1 | in2) implement ME.cat without materialising the concatenation.
#include "common.cuh"
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no -lcuda)
#include <stdlib.h>

#define TC_BM 128
#define TC_BK 64
#define TC_A_BYTES (TC_BM * 128)
#define TC_THREADS 512
#define TC_EPI_WARPS 8
#ifndef TC_L2_PREFETCH
#define TC_L2_PREFETCH 1  // producers prefetch the next offset's rows into L2
#endif
#ifndef TC_GI
#define TC_GI 1  // (offset, chunk) items per stage = per barrier round. 2 (fewer, larger rounds; 2 stages at n_tile 384) measured slower: 776 vs 837 TFLOP/s
#endif
static_assert(TC_GI == 1 || TC_GI == 2, "the B loader announces 1 or TC_GI items per group");
#ifndef TC_EPI_STAGED
#define TC_EPI_STAGED 1  // residual / output rows staged through smem for 64-byte coalesced segments (0: row per lane)
#endif
#ifndef TC_A_COLLECTOR
#define TC_A_COLLECTOR 0
#endif
#ifndef TC_NSPLIT0
#define TC_NSPLIT0 256  // N of the first MMA of a K step when the tile is wider than 256 columns
#endif
#define TC_MAX_STAGES 8
#define TC_MAX_SMEM 232448
#define TC_MIN_SMEM (120 * 1024)  // more than half an SM: one CTA per SM, so a 512-column TMEM alloc never blocks
#define TC_STAGE_OUT_BYTES 2048   // per epilogue warp: 32 rows x 64 bytes

struct TcParams {
    const __nv_bfloat16* in1;
    const __nv_bfloat16* in2;
    const uint8_t* wpacked;
    const int32_t* nbr;
    const int32_t* perm;
    const uint32_t* tile_masks;
    const float* scale;
    const float* shift;
    const __nv_bfloat16* residual;
    void* out;
    long long V_out;
    int Cin1, Cin2, nchunk1, nchunk2;
    int Cout, n_tile, n_ntiles, stages;
    int act, out_dtype, tmem_cols;
    int acc_bufs;  // 2 when two accumulators fit TMEM (2 n_tile <= 512): the epilogue overlaps the next tile's MMAs
    float slope;
    unsigned int b_bytes;
    int total_work;  // 256-row tile pairs x n-tiles
    int debug;       // debug build only (B2ME_TC_DEBUG): 1 skip the A gathers, 2 skip the B copies, 4 skip the MMAs
    int tma;         // 1: operands come through the TMA unit (gather4 rows / 2-D weight boxes, cta_group::2)
    // tensor maps (TMA mode): the two sources as [V_in, Cin] bf16 with a 64-channel x 1-row box (SWIZZLE_128B; rows are
    // picked by tile::gather4, absent neighbours (-1) and channels past Cin are out of bounds = zero-filled) and the
    // packed weights as 128-byte rows (box = one CTA's half of an item, no swizzle: the image is pre-swizzled)
    alignas(64) CUtensorMap tm_in1;
    alignas(64) CUtensorMap tm_in2;
    alignas(64) CUtensorMap tm_w;
};

// Optional in-kernel role timers (build with -DB2ME_TC_PROFILE; tools/conv_probe.py): cycles that lane 0 of each
// role of cluster 0 spends in each wait, accumulated over launches. Not compiled into the product library.
#ifdef B2ME_TC_PROFILE
__device__ unsigned long long g_tc_prof[2][8][8];  // [cta rank][role][counter]
#define PROF_DECL unsigned long long prof_[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_ = 0; (void)pt_;
#define PROF(slot, stmt) do { pt_ = clock64(); stmt; prof_[slot] += (unsigned long long)(clock64() - pt_); } while (0)
#define PROF_COUNT(slot) (++prof_[slot])
#define PROF_DUMP(role) do { if (blockIdx.x < 2 && lane == 0) for (int q_ = 0; q_ < 8; ++q_) \
        atomicAdd(&g_tc_prof[rank][role][q_], prof_[q_]); } while (0)
extern "C" int b2me_tc_prof_read(unsigned long long* host_out, int reset) {
    if (host_out && cudaMemcpyFromSymbol(host_out, g_tc_prof, sizeof(g_tc_prof)) != cudaSuccess) return B2ME_ELAUNCH;
    if (reset) {
        static unsigned long long zero[2 * 8 * 8];
        if (cudaMemcpyToSymbol(g_tc_prof, zero, sizeof(zero)) != cudaSuccess) return B2ME_ELAUNCH;
    }
    return B2ME_OK;
}
#else
#define PROF_DECL
#define PROF(slot, stmt) stmt
#define PROF_COUNT(slot)
#define PROF_DUMP(role)
#endif

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded wait: a protocol bug traps (launch error) instead of hanging the GPU
#ifndef TC_SPIN_WAIT
#define TC_SPIN_WAIT 0  // 1: poll with mbarrier.test_wait (never suspends) instead of try_wait
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
#if TC_SPIN_WAIT
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 28)) __trap();
    }
#else
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
#endif
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// the executing thread's arrival on `bar` fires when all of its prior cp.async copies have landed (no wait_group)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// TMA loads of a CTA pair (cta_group::2): the data lands in the issuing CTA's shared memory, the complete_tx goes to
// the mbarrier at cluster address `bar_cluster` (the LEADER's stage barrier, in either CTA), so the MMA thread waits on
// one barrier for both CTAs' operands and no relay is needed.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(addr), "r"(rank));
    return ra;
}
__device__ __forceinline__ void tma_gather4_pair(uint32_t dst, const CUtensorMap* tm, int col, int r0, int r1, int r2,
                                                 int r3, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(bar_cluster)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tm, int col, int row,
                                                 uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(col), "r"(row), "r"(bar_cluster)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// SWIZZLE_128B, K-major, 8-row groups 1024 B apart, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }


__device__ __forceinline__ void tmem_ld_x32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t rank) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
        "r"(rank)
        : "memory");
}
// wait whose acquire covers arrivals made by the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0, ok = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}
// COLL: 0 = default (A discarded after use), 1 = collector::a::fill (keep the A slice in the collector buffer),
// 2 = collector::a::lastuse (reuse the kept A slice: no second shared-memory read of A)
template <int COLL>
__device__ __forceinline__ void tc_mma_bf16_pair(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                 uint32_t accumulate) {
    if (COLL == 1) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16.collector::a::fill [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else if (COLL == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
            "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// one lane of the (converged) warp: true in exactly one lane
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}
// tcgen05.mma with descriptors given as low words (see below), executed by the calling thread
__device__ __forceinline__ void tc_mma_lo(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
        : "memory");
}
// Warp-uniform issue helpers: ALL lanes execute the statement with identical operands, elect.sync picks the issuing
// lane inside the block. Descriptors are passed as their low words (start address >> 4 | LBO field); the high word of
// the SWIZZLE_128B K-major descriptor (SBO 1024 B, version 1, swizzle mode 2) is the constant 0x40004040.
// FENCE: the elected lane first orders the generic-proxy (cp.async) writes of the stage before its async-proxy reads.
template <bool FENCE>
__device__ __forceinline__ void tc_kstep2(uint32_t d0, uint32_t d1, uint32_t a_lo, uint32_t b0_lo, uint32_t b1_lo,
                                          uint32_t idesc0, uint32_t idesc1, uint32_t accumulate) {
    if (FENCE) {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db0, db1;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %7, 0;\n\t"
            "mov.b64 da, {%2, %8};\n\tmov.b64 db0, {%3, %8};\n\tmov.b64 db1, {%4, %8};\n\t"
            "@q fence.proxy.async.shared::cta;\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db0, %5, p;\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%1], da, db1, %6, p;\n\t}" ::"r"(d0),
            "r"(d1), "r"(a_lo), "r"(b0_lo), "r"(b1_lo), "r"(idesc0), "r"(idesc1), "r"(accumulate), "r"(0x40004040u)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db0, db1;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %7, 0;\n\t"
            "mov.b64 da, {%2, %8};\n\tmov.b64 db0, {%3, %8};\n\tmov.b64 db1, {%4, %8};\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db0, %5, p;\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%1], da, db1, %6, p;\n\t}" ::"r"(d0),
            "r"(d1), "r"(a_lo), "r"(b0_lo), "r"(b1_lo), "r"(idesc0), "r"(idesc1), "r"(accumulate), "r"(0x40004040u)
            : "memory");
    }
}
template <bool FENCE>
__device__ __forceinline__ void tc_kstep1(uint32_t d0, uint32_t a_lo, uint32_t b0_lo, uint32_t idesc0,
                                          uint32_t accumulate) {
    if (FENCE) {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db0;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "mov.b64 da, {%1, %5};\n\tmov.b64 db0, {%2, %5};\n\t"
            "@q fence.proxy.async.shared::cta;\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db0, %3, p;\n\t}" ::"r"(d0),
            "r"(a_lo), "r"(b0_lo), "r"(idesc0), "r"(accumulate), "r"(0x40004040u)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db0;\n\t"
            "elect.sync _|q, 0xffffffff;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "mov.b64 da, {%1, %5};\n\tmov.b64 db0, {%2, %5};\n\t"
            "@q tcgen05.mma.cta_group::2.kind::f16 [%0], da, db0, %3, p;\n\t}" ::"r"(d0),
            "r"(a_lo), "r"(b0_lo), "r"(idesc0), "r"(accumulate), "r"(0x40004040u)
            : "memory");
    }
}
__device__ __forceinline__ void tc_commit_pair_elect(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)
                 : "memory");
    return v;
}

// ------------------------------------------------------------------------------------------ kernel
// N of the MMA instructions of one item: n_tile <= 256 -> one instruction; wider tiles -> TC_NSPLIT0 + the rest.
__host__ __device__ __forceinline__ int tc_n_first(int n_tile) { return n_tile > 256 ? TC_NSPLIT0 : n_tile; }

// KT = kernel volume of the map (27: k3 s1, 8: k2 s2 and its transpose, 1: identity / MinkowskiLinear)
template <int KT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) k_spconv_tc(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);

    const int S = p.stages;
    const uint32_t stage_bytes = TC_A_BYTES + p.b_bytes;
    const int tid = threadIdx.x;
    // shfl-broadcast: the compiler then knows the warp index (hence every role branch) is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer

    // carve: [stages][nbr_s 128*KT i32][scale Cout][shift Cout][epilogue staging 8 x 2 KB][barriers][tmem ptr]
    uint32_t off = (uint32_t)(S * TC_GI) * stage_bytes;  // S stages of TC_GI items
    int32_t* nbr_s = reinterpret_cast<int32_t*>(sm + off);
    off += TC_BM * KT * 4;
    off = (off + 15u) & ~15u;
    float* scale_s = reinterpret_cast<float*>(sm + off);
    off += p.Cout * 4;
    float* shift_s = reinterpret_cast<float*>(sm + off);
    off += p.Cout * 4;
    off = (off + 15u) & ~15u;
    const uint32_t stage_out = base + off;
    off += TC_EPI_WARPS * TC_STAGE_OUT_BYTES;
    const uint32_t bar_full = base + off;
    off += 8 * TC_MAX_STAGES;
    const uint32_t bar_empty = base + off;
    off += 8 * TC_MAX_STAGES;
    const uint32_t bar_nbr_full = base + off;
    off += 8;
    const uint32_t bar_nbr_empty = base + off;
    off += 8;
    const uint32_t bar_tmem_full = base + off;   // [2] one per accumulator buffer
    off += 16;
    const uint32_t bar_tmem_empty = base + off;  // [2]
    off += 16;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sm + off);

    // ---- one-time setup
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            // 128 producer threads (cp.async-completion arrivals) + the B expect_tx arrive (+ the peer's relay)
            // TMA mode: the leader's single arrive.expect_tx (all four transfers complete_tx on the leader's barrier)
            mbar_init(bar_full + 8 * s, p.tma ? 1 : (rank == 0 ? 130 : 129));
            mbar_init(bar_empty + 8 * s, 1);  // tcgen05.commit (multicast from the leader)
        }
        mbar_init(bar_nbr_full, 2);            // 2 prefetch warps
        mbar_init(bar_nbr_empty, 4);           // 4 producer warps
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_tmem_full + 8 * b, 1);                  // tcgen05.commit (multicast from the leader)
            mbar_init(bar_tmem_empty + 8 * b, 2 * TC_EPI_WARPS);  // epilogue warps of both CTAs (leader's barrier)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        // cta_group::2 allocation: the same warp of both CTAs issues it
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < p.Cout; i += TC_THREADS) {
        scale_s[i] = p.scale ? p.scale[i] : 1.f;
        shift_s[i] = p.shift ? p.shift[i] : 0.f;
    }
    tc_fence_before();
    cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    const int nchunk = p.nchunk1 + p.nchunk2;
    const int G = gridDim.x >> 1;          // clusters
    const int unit0 = blockIdx.x >> 1;     // this cluster's first work item
    // offsets (bit k) that at least one row of a 256-row tile pair needs: precomputed per map (b2me_tc_tile_masks),
    // so every role knows a tile's item list without waiting for the kernel-map rows
    auto tile_mask = [&](int w) -> uint32_t {
        if (!p.tile_masks) return 1u;
        const uint32_t m = __ldg(p.tile_masks + w / p.n_ntiles);
        return m ? m : 1u;
    };

    if (warp < 4 && p.tma) {
        // =============================== gather producers, TMA mode ===============================
        // warp w owns rows 32 w .. 32 w + 31 of the CTA's tile. Lane l reads the neighbour index of row 32 w + l; per
        // item the warp issues 8 tile::gather4 copies (4 rows x 128 bytes each, straight into the swizzled image;
        // absent neighbours = row -1 = out of bounds = zeros) from shfl-broadcast, i.e. warp-uniform, operands, so each
        // UTMALDG takes its operands from uniform registers without a per-lane serialisation loop.
        const uint32_t full_leader = mapa_u32(bar_full, 0u);
        int ist = 0, iph = 0, it = 0;
        uint32_t kmask_next = __shfl_sync(0xffffffffu, unit0 < p.total_work ? tile_mask(unit0) : 0u, 0);
        for (int w = unit0; w < p.total_work; w += G, ++it) {
            const uint32_t kmask = kmask_next;
            if (w + G < p.total_work) kmask_next = __shfl_sync(0xffffffffu, tile_mask(w + G), 0);
            mbar_wait(bar_nbr_full, (uint32_t)it & 1u);
#pragma unroll 1
            for (uint32_t m = kmask; m; m &= m - 1u) {
                const int k = __ffs((int)m) - 1;
                const int id = nbr_s[(32 * warp + lane) * KT + k];
                int ids[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) ids[q] = __shfl_sync(0xffffffffu, id, q);
                if ((m & (m - 1u)) == 0u) {  // last offset of this tile: the kernel-map buffer may be refilled
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_nbr_empty);
                } else if (TC_L2_PREFETCH) {
                    // pull the rows of the NEXT offset into L2 one whole offset (= nchunk items) ahead of their gather
                    const uint32_t m2 = m & (m - 1u);
                    const int k2 = __ffs((int)m2) - 1;
                    const int id2 = nbr_s[(32 * warp + lane) * KT + k2];
                    if (id2 >= 0) {
                        for (int ch = 0; ch < p.nchunk1; ++ch) prefetch_l2(p.in1 + (long long)id2 * p.Cin1 + ch * TC_BK);
                        for (int ch = 0; ch < p.nchunk2; ++ch) prefetch_l2(p.in2 + (long long)id2 * p.Cin2 + ch * TC_BK);
                    }
                }
#pragma unroll 1
                for (int c = 0; c < nchunk; ++c) {
                    mbar_wait(bar_empty + 8 * ist, (uint32_t)iph ^ 1u);
                    if (tc_elect_one()) {
                        const uint32_t a_s = base + (uint32_t)ist * stage_bytes + (uint32_t)(32 * warp) * 128u;
                        const CUtensorMap* tm = c < p.nchunk1 ? &p.tm_in1 : &p.tm_in2;
                        const int col = (c < p.nchunk1 ? c : c - p.nchunk1) * TC_BK;
#pragma unroll
                        for (int g = 0; g < 8; ++g)
                            tma_gather4_pair(a_s + (uint32_t)g * 512u, tm, col, ids[4 * g], ids[4 * g + 1],
                                             ids[4 * g + 2], ids[4 * g + 3], full_leader + 8 * ist);
                    }
                    __syncwarp();
                    if (++ist == S) { ist = 0; iph ^= 1; }
                }
            }
        }
    } else if (warp < 4) {
        // =============================== gather producers ===============================
        const int j = tid & 7;        // 16-byte piece inside the 128-byte row
        const int rbase = tid >> 3;   // rows rbase + 16*i
        int ist = 0, iph = 0;         // stage / phase of the next group of items to issue
        int isub = 0;                 // item slot inside the group (a stage holds TC_GI items)
        int it = 0;
        PROF_DECL
        const long long t_role0 = clock64();
        (void)t_role0;
        uint32_t kmask_next = unit0 < p.total_work ? tile_mask(unit0) : 0u;
        for (int w = unit0; w < p.total_work; w += G, ++it) {
            const uint32_t kmask = kmask_next;
            if (w + G < p.total_work) kmask_next = tile_mask(w + G);
            PROF(1, mbar_wait(bar_nbr_full, (uint32_t)it & 1u));
#pragma unroll 1
            for (int k = 0; k < KT; ++k) {
                if (!((kmask >> k) & 1u)) continue;
                int idx[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) idx[i] = nbr_s[(rbase + 16 * i) * KT + k];
                if ((kmask >> k) == 1u) {  // last offset of this tile: the kernel-map buffer may be refilled
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_nbr_empty);
                } else if (TC_L2_PREFETCH && j < nchunk) {
                    // the gathered rows come from all over the tensor (DRAM latency >> the ring's depth in time): pull
                    // the rows of the NEXT offset into L2 now, one whole offset (= nchunk items) ahead of their gather.
                    // lane j takes the 128-byte chunk j of each of this thread's 8 rows.
                    const int k2 = k + 1 + __ffs((int)(kmask >> (k + 1))) - 1;
                    const __nv_bfloat16* psrc;
                    int pcin, pcoff;
                    if (j < p.nchunk1) { psrc = p.in1; pcin = p.Cin1; pcoff = j * TC_BK; }
                    else { psrc = p.in2; pcin = p.Cin2; pcoff = (j - p.nchunk1) * TC_BK; }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int id = nbr_s[(rbase + 16 * i) * KT + k2];
                        if (id >= 0) prefetch_l2(psrc + (long long)id * pcin + pcoff);
                    }
                }
#pragma unroll 1
                for (int c = 0; c < nchunk; ++c) {
                    if (isub == 0) {
                        PROF(2, mbar_wait(bar_empty + 8 * ist, (uint32_t)iph ^ 1u));
                        PROF_COUNT(7);
                    }
                    const __nv_bfloat16* src;
                    int cin, coff;
                    if (c < p.nchunk1) { src = p.in1; cin = p.Cin1; coff = c * TC_BK; }
                    else { src = p.in2; cin = p.Cin2; coff = (c - p.nchunk1) * TC_BK; }
                    const int kw = min(TC_BK, cin - coff);
                    if (j * 8 < kw && !(p.debug & 1)) {
                        const uint32_t a_s = base + (uint32_t)(ist * TC_GI + isub) * stage_bytes;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = rbase + 16 * i;
                            const uint32_t dst = a_s + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
                            const int id = idx[i];
                            const __nv_bfloat16* g = src + (id >= 0 ? ((long long)id * cin + coff + j * 8) : 0);
                            cp_async_16(dst, g, id >= 0 ? 16u : 0u);
                        }
                    }
                    const bool tile_last = ((kmask >> k) == 1u) && (c == nchunk - 1);
                    if (++isub == TC_GI || tile_last) {
                        // asynchronous publication: this thread's arrival fires when its copies have landed, so the
                        // producers run ahead as far as the ring allows and never wait for their own gathers
                        cp_async_mbar_arrive_noinc(bar_full + 8 * ist);
                        isub = 0;
                        if (++ist == S) { ist = 0; iph ^= 1; }
                    }
                }
            }
        }
#ifdef B2ME_TC_PROFILE
        prof_[0] = (unsigned long long)(clock64() - t_role0);
        if (warp == 0) PROF_DUMP(0);
#endif
    } else if (warp == 4) {
        // =============================== MMA issuer (leader) / stage relay (peer) ===============================
        // The whole warp walks the item list and waits on the barriers; lane 0 alone issues the tcgen05 instructions
        // (measured: a wait executed by a lone lane of a divergent warp costs ~210 cycles even on a complete phase).
        if (rank == 0) {
            // The whole warp walks the item list with warp-uniform values; every tcgen05 instruction sits in an asm
            // block that elects the issuing lane itself (elect.sync), so the compiler emits ELECT + predicated
            // UTCHMMA instead of a per-lane serialisation loop: ~90 instead of ~300 SASS instructions per item. The
            // tensor pipe queues only a couple of MMAs, so every cycle of issue overhead beyond that slack idles it.
            const int n_a = tc_n_first(p.n_tile), n_b = p.n_tile - n_a;  // N of the one or two MMAs per K step
            // kind::f16, bf16 x bf16 -> f32, K-major A and B, M = 256 (cta_group::2); each CTA's smem holds N/2 rows
            const uint32_t idesc_a = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_a >> 3) << 17) | (16u << 24);
            const uint32_t idesc_b = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_b >> 3) << 17) | (16u << 24);
            // K steps (16 channels) of a full chunk and of the last chunk of each source
            const int ks_last1 = (p.Cin1 - (p.nchunk1 - 1) * TC_BK) / 16;
            const int ks_last2 = p.nchunk2 ? (p.Cin2 - (p.nchunk2 - 1) * TC_BK) / 16 : 0;
            // low words of the SWIZZLE_128B K-major descriptors of stage 0 (high word is constant, see tc_kstep)
            const uint32_t a_lo0 = ((base >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t stage_lo = stage_bytes >> 4;
            const uint32_t b_off_lo = TC_A_BYTES >> 4, b2_off_lo = (TC_A_BYTES + (uint32_t)(n_a >> 1) * 128u) >> 4;
            const bool need_fence = !p.tma;  // cp.async (generic proxy) writes of A -> visible to the MMA (async proxy)
            int st = 0, ph = 0, it = 0;
            PROF_DECL
            const long long t_role0 = clock64();
            (void)t_role0;
            // shfl-broadcast values are warp-uniform to the compiler: with the tile masks and the TMEM base uniform, the
            // whole loop (stage index, phase, descriptors) is computed on the uniform datapath
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            uint32_t kmask_next = __shfl_sync(0xffffffffu, unit0 < p.total_work ? tile_mask(unit0) : 0u, 0);
            for (int w = unit0; w < p.total_work; w += G, ++it) {
                const uint32_t kmask = kmask_next;
                if (w + G < p.total_work) kmask_next = __shfl_sync(0xffffffffu, tile_mask(w + G), 0);
                // accumulator buffer of this work item and how often it has been used before
                const int ab = p.acc_bufs == 2 ? (it & 1) : 0;
                const int au = p.acc_bufs == 2 ? (it >> 1) : it;
                const uint32_t tmem_acc = tmem_u + (uint32_t)(ab * p.n_tile);
                if (au > 0) {  // both CTAs' epilogues must have drained the previous tile of this buffer
                    PROF(2, mbar_wait(bar_tmem_empty + 8 * ab, (uint32_t)(au - 1) & 1u));
                    tc_fence_after();
                }
                uint32_t acc = 0u;
                for (uint32_t m = kmask; m; m &= m - 1u) {  // one pass per kernel offset the tile needs
                    for (int c = 0; c < nchunk; ++c) {
                        int ks = (c == p.nchunk1 - 1) ? ks_last1 : ((c == nchunk - 1) ? ks_last2 : TC_BK / 16);
#ifdef B2ME_TC_PROFILE
                        if (p.debug & 4) ks = 0;
#endif
                        PROF(3, mbar_wait(bar_full + 8 * st, (uint32_t)ph));
                        PROF_COUNT(7);
#ifdef B2ME_TC_PROFILE
                        pt_ = clock64();
#endif
                        tc_fence_after();
                        const uint32_t a_lo = a_lo0 + (uint32_t)st * stage_lo;
                        if (tc_elect_one()) {
#if defined(B2ME_TC_PROFILE) || defined(B2ME_TC_EXPERIMENT)
                            if (need_fence && !(p.debug & 8)) fence_proxy_async();  // 8: timing experiment only
#else
                            if (need_fence) fence_proxy_async();
#endif
                            // the two instructions of a K step share the A slice; both read it from shared memory
                            // (keeping it in the collector buffer, collector::a::fill / lastuse, was measured slower)
                            for (int kk = 0; kk < ks; ++kk) {
                                tc_mma_lo(tmem_acc, a_lo + 2u * kk, a_lo + b_off_lo + 2u * kk, idesc_a, acc);
                                if (n_b)
                                    tc_mma_lo(tmem_acc + (uint32_t)n_a, a_lo + 2u * kk, a_lo + b2_off_lo + 2u * kk,
                                              idesc_b, acc);
                                acc = 1u;
                            }
                            tc_commit_pair(bar_empty + 8 * st);  // frees the stage in both CTAs
                        }
                        acc = ks > 0 ? 1u : acc;
                        __syncwarp();
                        if (++st == S) { st = 0; ph ^= 1; }
#ifdef B2ME_TC_PROFILE
                        prof_[5] += (unsigned long long)(clock64() - pt_);
#endif
                    }
                }
                tc_commit_pair_elect(bar_tmem_full + 8 * ab);  // accumulators of both CTAs are complete
            }
#ifdef B2ME_TC_PROFILE
            prof_[0] = (unsigned long long)(clock64() - t_role0);
            PROF_DUMP(1);
#endif
        } else if (!p.tma) {
            // peer: when a stage of THIS CTA is full (A gathered, B landed), tell the leader's full barrier
            // (TMA mode: the peer's transfers complete_tx on the leader's barrier themselves, nothing to relay)
            int st = 0, ph = 0;
            PROF_DECL
            const long long t_role0 = clock64();
            (void)t_role0;
            uint32_t kmask_next = unit0 < p.total_work ? tile_mask(unit0) : 0u;
            for (int w = unit0; w < p.total_work; w += G) {
                const uint32_t kmask = kmask_next;
                if (w + G < p.total_work) kmask_next = tile_mask(w + G);
                int sub = 0;
                for (int k = 0; k < KT; ++k) {
                    if (!((kmask >> k) & 1u)) continue;
                    for (int c = 0; c < nchunk; ++c) {
                        const bool tile_last = ((kmask >> k) == 1u) && (c == nchunk - 1);
                        if (++sub < TC_GI && !tile_last) continue;
                        sub = 0;
                        PROF(3, mbar_wait(bar_full + 8 * st, (uint32_t)ph));
                        PROF_COUNT(7);
                        if (lane == 0) {
                            fence_proxy_async();
                            mbar_arrive_remote(bar_full + 8 * st, 0u);
                        }
                        __syncwarp();
                        if (++st == S) { st = 0; ph ^= 1; }
                    }
                }
            }
#ifdef B2ME_TC_PROFILE
            prof_[0] = (unsigned long long)(clock64() - t_role0);
            PROF_DUMP(5);
#endif
        }
    } else if (warp == 5) {
        // =============================== weight (B) loader ===============================
        {
            int st = 0, ph = 0;
            const uint32_t full_leader = mapa_u32(bar_full, 0u);
            uint32_t kmask_next = unit0 < p.total_work ? tile_mask(unit0) : 0u;
            for (int w = unit0; w < p.total_work; w += G) {
                const uint32_t kmask = kmask_next;
                if (w + G < p.total_work) kmask_next = tile_mask(w + G);
                const int nt = w % p.n_ntiles;
                int sub = 0;
                for (int k = 0; k < KT; ++k) {
                    if (!((kmask >> k) & 1u)) continue;
                    for (int c = 0; c < nchunk; ++c) {
                        const bool tile_last = ((kmask >> k) == 1u) && (c == nchunk - 1);
                        if (sub == 0) mbar_wait(bar_empty + 8 * st, (uint32_t)ph ^ 1u);
                        if (lane == 0) {
                            const uint32_t b_s = base + (uint32_t)(st * TC_GI + sub) * stage_bytes + TC_A_BYTES;
                            const uint8_t* g =
                                p.wpacked + ((((size_t)nt * KT + k) * nchunk + c) * 2 + rank) * (size_t)p.b_bytes;
                            if (p.tma) {
                                // leader: one arrive announcing all four transfers of the stage (A and B of both CTAs)
                                if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * st, 2u * (TC_A_BYTES + p.b_bytes));
                                const long long item = ((long long)nt * KT + k) * nchunk + c;
                                tma_load_2d_pair(b_s, &p.tm_w, 0, (int)((item * 2 + rank) * (p.n_tile / 2)),
                                                 full_leader + 8 * st);
                            } else if (p.debug & 2) {
                                if (sub == 0) mbar_arrive(bar_full + 8 * st);
                            } else {
                                // one arrive per group, announcing the bytes of all its items (1 or TC_GI)
                                if (sub == 0)
                                    mbar_arrive_expect_tx(bar_full + 8 * st, tile_last ? p.b_bytes : TC_GI * p.b_bytes);
                                bulk_copy_g2s(b_s, g, p.b_bytes, bar_full + 8 * st);
                            }
                        }
                        if (++sub == TC_GI || tile_last) {
                            __syncwarp();
                            sub = 0;
                            if (++st == S) { st = 0; ph ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp < 8) {
        // =============================== kernel-map prefetch ===============================
        // warp wl stages rows 64 wl .. 64 wl + 63 of every tile: 64 x KT entries, 2 KT per lane, held in registers
        // until the producers have released the (single) smem buffer.
        const int wl = warp - 6;
        constexpr int NJ = 2 * KT;
        int it = 0;
        for (int w = unit0; w < p.total_work; w += G, ++it) {
            const long long row0 = ((long long)(w / p.n_ntiles) * 2 + rank) * TC_BM + 64 * wl;
            int rowreg[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const long long slot = row0 + lane + 32 * h;
                rowreg[h] = (slot < p.V_out) ? (p.perm ? __ldg(p.perm + slot) : (int)slot) : -1;
            }
            int v[NJ];
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) {
                const int e = lane + 32 * jj;  // < 64 * KT
                const int rl = e / KT, k = e - rl * KT;
                const int r0 = __shfl_sync(0xffffffffu, rowreg[0], rl & 31);
                const int r1 = __shfl_sync(0xffffffffu, rowreg[1], rl & 31);
                const int row = (rl >> 5) ? r1 : r0;
                if (p.nbr) v[jj] = (row >= 0) ? __ldg(p.nbr + (long long)row * KT + k) : -1;
                else v[jj] = row;
            }
            if (it >= 1) mbar_wait(bar_nbr_empty, (uint32_t)(it - 1) & 1u);
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) nbr_s[64 * wl * KT + lane + 32 * jj] = v[jj];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_nbr_full);
        }
    } else {
        // =============================== epilogue ===============================
        // 8 warps: warp e handles TMEM lanes 32 (e & 3) .. (a warp may only touch the lane quarter warp_id % 4) and
        // the column half e >> 2 of the tile
        const int ew = warp - 8;
        const int we = ew & 3, chalf = ew >> 2;
        const uint32_t lane_addr0 = tmem_base + ((uint32_t)(we * 32) << 16);
        const uint32_t stg = stage_out + (uint32_t)ew * TC_STAGE_OUT_BYTES;
        // staging image: row r (0..31) = 64 bytes, 16-byte piece q stored at q ^ ((r >> 1) & 3): conflict-free for both
        // the row-per-lane view (lane = row) and the coalesced view (4 lanes per row, 8 rows per access)
        const uint32_t own = stg + (uint32_t)lane * 64u;
        const int own_sw = (lane >> 1) & 3;
        const int crow = lane >> 2, cq = lane & 3;  // coalesced view: rows crow + 8 m, piece cq
        // this warp's columns: 32-column chunks [c_lo, c_hi) of the tile
        const int nch = (p.n_tile + 31) >> 5;
        const int c_lo = chalf ? ((nch + 1) >> 1) * 32 : 0;
        const int c_hi = chalf ? p.n_tile : min(p.n_tile, ((nch + 1) >> 1) * 32);
        int it = 0;
        PROF_DECL
        for (int w = unit0; w < p.total_work; w += G, ++it) {
            const int tile_p = w / p.n_ntiles;
            const int n0 = (w - tile_p * p.n_ntiles) * p.n_tile;
            const int ab = p.acc_bufs == 2 ? (it & 1) : 0;
            const uint32_t au_par = (uint32_t)(p.acc_bufs == 2 ? (it >> 1) : it) & 1u;
            const uint32_t lane_addr = lane_addr0 + (uint32_t)(ab * p.n_tile);
            const uint32_t bar_tf = bar_tmem_full + 8 * ab, bar_te = bar_tmem_empty + 8 * ab;
            const long long slot = ((long long)tile_p * 2 + rank) * TC_BM + we * 32 + lane;
            const int row = (slot < p.V_out) ? (p.perm ? __ldg(p.perm + slot) : (int)slot) : -1;
            int crows[4];  // output rows of the coalesced view
#pragma unroll
            for (int m = 0; m < 4; ++m) crows[m] = __shfl_sync(0xffffffffu, row, crow + 8 * m);

            if (p.out_dtype == B2ME_BF16) {
                const __nv_bfloat16* resp = p.residual;
                __nv_bfloat16* outp = reinterpret_cast<__nv_bfloat16*>(p.out);
#if TC_EPI_STAGED
                uint4 R[4];
                auto load_res = [&](int cb) {  // coalesced: 64-byte segment of 8 rows per access
                    const int cw = min(32, p.n_tile - cb);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        R[m] = make_uint4(0u, 0u, 0u, 0u);
                        if (crows[m] >= 0 && cq * 8 < cw)
                            R[m] = __ldg(reinterpret_cast<const uint4*>(resp + (long long)crows[m] * p.Cout + n0 + cb +
                                                                        cq * 8));
                    }
                };
                if (resp && c_lo < c_hi) load_res(c_lo);
                PROF(1, mbar_wait(bar_tf, au_par));
                tc_fence_after();
#ifdef B2ME_TC_PROFILE
                const long long t_epi0 = clock64();
#endif
                if (c_lo >= c_hi) {  // narrow tile: this warp has no columns
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_te, 0u);
                }
                for (int cb = c_lo; cb < c_hi; cb += 32) {
                    const int cw = min(32, p.n_tile - cb);
                    uint32_t r[32];
                    if (cw == 32) {
                        tmem_ld_x32(lane_addr + (uint32_t)cb, r);
                    } else {
                        tmem_ld_x16(lane_addr + (uint32_t)cb, r);
#pragma unroll
                        for (int q = 16; q < 32; ++q) r[q] = 0u;
                    }
                    if (resp) {
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const int rr = crow + 8 * m;
                            st_shared_v4(stg + (uint32_t)rr * 64u + (uint32_t)((cq ^ ((rr >> 1) & 3)) << 4), R[m]);
                        }
                        __syncwarp();
                        if (cb + 32 < c_hi) load_res(cb + 32);  // in flight during this chunk's math and stores
                    }
                    tmem_ld_wait();
                    if (cb + 32 >= c_hi) {  // this warp's part of the accumulator is read: release it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(bar_te, 0u);  // the leader's barrier
                    }
                    float v[32];
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const int col = min(n0 + cb + q4 * 4, p.Cout - 4);  // Cout % 16 == 0: float4-aligned
                        const float4 sc = *reinterpret_cast<const float4*>(scale_s + col);
                        const float4 sh = *reinterpret_cast<const float4*>(shift_s + col);
                        v[q4 * 4 + 0] = __uint_as_float(r[q4 * 4 + 0]) * sc.x + sh.x;
                        v[q4 * 4 + 1] = __uint_as_float(r[q4 * 4 + 1]) * sc.y + sh.y;
                        v[q4 * 4 + 2] = __uint_as_float(r[q4 * 4 + 2]) * sc.z + sh.z;
                        v[q4 * 4 + 3] = __uint_as_float(r[q4 * 4 + 3]) * sc.w + sh.w;
                    }
                    if (resp) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const uint4 a = ld_shared_v4(own + (uint32_t)((q4 ^ own_sw) << 4));
                            const uint32_t wv[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                v[q4 * 8 + 2 * e] += __uint_as_float(wv[e] << 16);
                                v[q4 * 8 + 2 * e + 1] += __uint_as_float(wv[e] & 0xFFFF0000u);
                            }
                        }
                        __syncwarp();  // every lane has read its residual row before the buffer takes the outputs
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint32_t wv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float x0 = apply_act(v[q4 * 8 + 2 * e], p.act, p.slope);
                            const float x1 = apply_act(v[q4 * 8 + 2 * e + 1], p.act, p.slope);
                            __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                            wv[e] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        st_shared_v4(own + (uint32_t)((q4 ^ own_sw) << 4), make_uint4(wv[0], wv[1], wv[2], wv[3]));
                    }
                    __syncwarp();
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int rr = crow + 8 * m;
                        const uint4 o = ld_shared_v4(stg + (uint32_t)rr * 64u + (uint32_t)((cq ^ ((rr >> 1) & 3)) << 4));
                        if (crows[m] >= 0 && cq * 8 < cw)
                            *reinterpret_cast<uint4*>(outp + (long long)crows[m] * p.Cout + n0 + cb + cq * 8) = o;
                    }
                    __syncwarp();  // staging is reused by the next chunk
                }
#else
                uint4 R[4];
                // row-per-lane epilogue (TC_EPI_STAGED 0): every lane loads / stores the 64 contiguous bytes of ITS row per
                // 32-column chunk, no shared-memory round trip. Measured 10-60 % SLOWER than the staged variant (32
                // half-used sectors per store instruction), kept for experiments only.
                const long long rbase_o = (long long)(row >= 0 ? row : 0) * p.Cout + n0;
                auto load_res = [&](int cb) {
                    const int cw = min(32, p.n_tile - cb);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        R[m] = make_uint4(0u, 0u, 0u, 0u);
                        if (row >= 0 && m * 8 < cw)
                            R[m] = __ldg(reinterpret_cast<const uint4*>(resp + rbase_o + cb + m * 8));
                    }
                };
                if (resp && c_lo < c_hi) load_res(c_lo);
                PROF(1, mbar_wait(bar_tf, au_par));
                tc_fence_after();
#ifdef B2ME_TC_PROFILE
                const long long t_epi0 = clock64();
#endif
                if (c_lo >= c_hi) {  // narrow tile: this warp has no columns
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_te, 0u);
                }
                for (int cb = c_lo; cb < c_hi; cb += 32) {
                    const int cw = min(32, p.n_tile - cb);
                    uint32_t r[32];
                    if (cw == 32) {
                        tmem_ld_x32(lane_addr + (uint32_t)cb, r);
                    } else {
                        tmem_ld_x16(lane_addr + (uint32_t)cb, r);
#pragma unroll
                        for (int q = 16; q < 32; ++q) r[q] = 0u;
                    }
                    uint4 Rc[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) Rc[m] = R[m];
                    if (resp && cb + 32 < c_hi) load_res(cb + 32);  // in flight during this chunk's math and stores
                    tmem_ld_wait();
                    if (cb + 32 >= c_hi) {  // this warp's part of the accumulator is read: release it
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(bar_te, 0u);  // the leader's barrier
                    }
                    float v[32];
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const int col = min(n0 + cb + q4 * 4, p.Cout - 4);  // Cout % 16 == 0: float4-aligned
                        const float4 sc = *reinterpret_cast<const float4*>(scale_s + col);
                        const float4 sh = *reinterpret_cast<const float4*>(shift_s + col);
                        v[q4 * 4 + 0] = __uint_as_float(r[q4 * 4 + 0]) * sc.x + sh.x;
                        v[q4 * 4 + 1] = __uint_as_float(r[q4 * 4 + 1]) * sc.y + sh.y;
                        v[q4 * 4 + 2] = __uint_as_float(r[q4 * 4 + 2]) * sc.z + sh.z;
                        v[q4 * 4 + 3] = __uint_as_float(r[q4 * 4 + 3]) * sc.w + sh.w;
                    }
                    if (resp) {
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const uint32_t wv[4] = {Rc[q4].x, Rc[q4].y, Rc[q4].z, Rc[q4].w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                v[q4 * 8 + 2 * e] += __uint_as_float(wv[e] << 16);
                                v[q4 * 8 + 2 * e + 1] += __uint_as_float(wv[e] & 0xFFFF0000u);
                            }
                        }
                    }
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint32_t wv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float x0 = apply_act(v[q4 * 8 + 2 * e], p.act, p.slope);
                            const float x1 = apply_act(v[q4 * 8 + 2 * e + 1], p.act, p.slope);
                            __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
                            wv[e] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        if (row >= 0 && q4 * 8 < cw)
                            *reinterpret_cast<uint4*>(outp + rbase_o + cb + q4 * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
                    }
                }
#endif
#ifdef B2ME_TC_PROFILE
                prof_[2] += (unsigned long long)(clock64() - t_epi0);
                ++prof_[7];
#endif
            } else {
                // fp32 rows (parity tests only): row-per-lane stores, 16-column chunks split between the two halves
                mbar_wait(bar_tf, au_par);
                tc_fence_after();
                const int n16 = p.n_tile >> 4;
                const int q_lo = chalf ? (n16 + 1) >> 1 : 0, q_hi = chalf ? n16 : (n16 + 1) >> 1;
                if (q_lo >= q_hi) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_te, 0u);
                }
                for (int qc = q_lo; qc < q_hi; ++qc) {
                    const int cb = qc * 16;
                    uint32_t r[16];
                    tmem_ld_x16(lane_addr + (uint32_t)cb, r);
                    tmem_ld_wait();
                    if (qc + 1 >= q_hi) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(bar_te, 0u);
                    }
                    if (row < 0) continue;
                    float v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        v[q] = __uint_as_float(r[q]) * scale_s[n0 + cb + q] + shift_s[n0 + cb + q];
                    const long long o = (long long)row * p.Cout + n0 + cb;
                    if (p.residual) {
                        const uint4* rp = reinterpret_cast<const uint4*>(p.residual + o);
                        const uint4 ra = __ldg(rp), rb = __ldg(rp + 1);
                        const uint32_t wv[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            v[2 * q] += __uint_as_float(wv[q] << 16);
                            v[2 * q + 1] += __uint_as_float(wv[q] & 0xFFFF0000u);
                        }
                    }
                    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        op[q] = make_float4(apply_act(v[4 * q], p.act, p.slope), apply_act(v[4 * q + 1], p.act, p.slope),
                                            apply_act(v[4 * q + 2], p.act, p.slope),
                                            apply_act(v[4 * q + 3], p.act, p.slope));
                }
            }
        }
#ifdef B2ME_TC_PROFILE
        if (warp == 8) PROF_DUMP(4);
#endif
    }
    __syncwarp();
    tc_fence_before();
    cluster_sync_all();  // neither CTA may exit (or free TMEM) while the pair's MMAs / remote arrives are in flight
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

// one warp per 256-row tile pair: OR of the neighbour-occupancy masks of its rows. Lane l takes rows l, l + 32, ...
// (8 rows) and reads each row's K entries itself, so a lane has 8 K independent loads in flight.
__global__ void __launch_bounds__(256) k_tc_tile_masks(const int32_t* __restrict__ nbr, const int32_t* __restrict__ perm,
                                                       long long V, int K, uint32_t* __restrict__ masks,
                                                       long long npairs) {
    const long long pair = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (pair >= npairs) return;
    const int lane = threadIdx.x & 31;
    long long rows[(2 * TC_BM) / 32];
#pragma unroll
    for (int i = 0; i < (2 * TC_BM) / 32; ++i) {
        const long long slot = pair * (2 * TC_BM) + lane + 32 * i;
        rows[i] = slot < V ? (perm ? (long long)__ldg(perm + slot) : slot) : -1;
    }
    uint32_t m = 0u;
#pragma unroll
    for (int i = 0; i < (2 * TC_BM) / 32; ++i) {
        if (rows[i] < 0) continue;
        const int32_t* r = nbr + rows[i] * K;
        for (int k = 0; k < K; ++k)
            if (__ldg(r + k) >= 0) m |= 1u << k;
    }
    m = __reduce_or_sync(0xffffffffu, m);
    if (lane == 0) masks[pair] = m;
}

extern "C" int b2me_tc_tile_masks(const int32_t* nbr, const int32_t* perm, int64_t V_out, int K, uint32_t* masks,
                                  b2me_stream_t stream) {
    if (!nbr || !masks || V_out < 0 || K < 1 || K > 32) return B2ME_EINVAL;
    if (V_out == 0) return B2ME_OK;
    const long long npairs = ceil_div64(V_out, 2 * TC_BM);
    k_tc_tile_masks<<<(unsigned)ceil_div64(npairs, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        nbr, perm, V_out, K, masks, npairs);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ host side
// Output-channel tile. K = 27 (many items per tile): the widest tile TMEM takes (<= 384 columns), so that the gathered
// rows are read once (K = 8 too: measured). K = 1 (few items per tile, contiguous A): tiles of <= 256 columns so that
// two accumulators fit the 512 TMEM columns and the epilogue overlaps the next tile's MMAs.
static int tc_n_tile(int K, int Cout) {
    if (Cout <= 256) return Cout;
    if (K != 1 && Cout <= 384) return Cout;
    if (Cout % 256 == 0) return 256;
    if (Cout % 192 == 0) return 192;
    if (Cout % 128 == 0) return 128;
    return Cout <= 384 ? Cout : 0;
}

extern "C" int b2me_tc_supported(int K, int Cin1, int Cin2, int Cout) {
    if ((K != 1 && K != 8 && K != 27) || Cin1 < 16 || Cin2 < 0 || Cout < 16 || Cout > 1024) return 0;
    if (Cin1 % 16 || Cin2 % 16 || Cout % 16) return 0;
    const int nt = tc_n_tile(K, Cout);
    if (nt <= 0 || Cout % nt) return 0;
    if (nt > 256 && ((nt - TC_NSPLIT0) % 16 || nt - TC_NSPLIT0 < 32)) return 0;
    return 1;
}

extern "C" size_t b2me_tc_packed_bytes(int K, int Cin1, int Cin2, int Cout) {
    if (!b2me_tc_supported(K, Cin1, Cin2, Cout)) return 0;
    const int nchunk = (Cin1 + TC_BK - 1) / TC_BK + (Cin2 + TC_BK - 1) / TC_BK;
    return (size_t)K * nchunk * (size_t)Cout * 128;
}

// one thread per 16-byte piece of the packed image
__global__ void k_tc_pack(const float* __restrict__ W, int K, int Cin1, int Cin2, int Cout, int n_tile, int nchunk1,
                          int nchunk2, uint4* __restrict__ packed, long long total_pieces) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_pieces) return;
    const int nchunk = nchunk1 + nchunk2;
    const int j = (int)(t & 7);
    long long rest = t >> 3;
    const int n = (int)(rest % n_tile);
    rest /= n_tile;
    const int c = (int)(rest % nchunk);
    rest /= nchunk;
    const int k = (int)(rest % K);
    const int nt = (int)(rest / K);
    int cin_base, cin_end;
    if (c < nchunk1) { cin_base = c * TC_BK; cin_end = Cin1; }
    else { cin_base = Cin1 + (c - nchunk1) * TC_BK; cin_end = Cin1 + Cin2; }
    const int Cin = Cin1 + Cin2;
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        float f[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int cin = cin_base + j * 8 + e * 2 + h;
            f[h] = (cin < cin_end) ? W[((long long)k * Cin + cin) * Cout + (nt * n_tile + n)] : 0.f;
        }
        __nv_bfloat162 v = __floats2bfloat162_rn(f[0], f[1]);
        w[e] = *reinterpret_cast<uint32_t*>(&v);
    }
    // image of one item = [cta rank 0 | cta rank 1]; a CTA's part = its half of the rows of each of the (one or two)
    // MMA instructions of a K step: instruction a covers columns [0, n_a), instruction b covers [n_a, n_tile)
    const int n_a = tc_n_first(n_tile);
    const int in_b = n >= n_a;
    const int nseg = in_b ? n_tile - n_a : n_a;       // N of the instruction this column belongs to
    const int nn = in_b ? n - n_a : n;
    const int q = nseg / 2;
    const int r = nn / q;
    const int lrow = (in_b ? n_a / 2 : 0) + (nn - r * q);
    const long long item = ((long long)nt * K + k) * nchunk + c;
    const long long piece = (item * 2 + r) * ((long long)(n_tile / 2) * 8) + (long long)lrow * 8 + (j ^ (lrow & 7));
    packed[piece] = make_uint4(w[0], w[1], w[2], w[3]);
}

extern "C" int b2me_tc_pack_weights(const float* W, int K, int Cin1, int Cin2, int Cout, void* packed,
                                    b2me_stream_t stream) {
    if (!W || !packed) return B2ME_EINVAL;
    if (!b2me_tc_supported(K, Cin1, Cin2, Cout)) return B2ME_EUNSUPPORTED;
    const int nt = tc_n_tile(K, Cout);
    const int nchunk1 = (Cin1 + TC_BK - 1) / TC_BK, nchunk2 = (Cin2 + TC_BK - 1) / TC_BK;
    const long long total = (long long)(Cout / nt) * K * (nchunk1 + nchunk2) * nt * 8;
    k_tc_pack<<<(unsigned)ceil_div64(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        W, K, Cin1, Cin2, Cout, nt, nchunk1, nchunk2, reinterpret_cast<uint4*>(packed), total);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

static int tc_num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            sms = n;
        else
            sms = B2ME_NUM_SMS;
    }
    return sms;
}

// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup (no link-time dependency on libcuda)
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tc_encode_fn tc_encoder() {
    static tc_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tc_encode_fn>(sym);
    }
    return fn;
}
// 2-D bf16 tensor [rows, cols] (row pitch = cols), box = box_cols x box_rows
static bool tc_make_map(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint32_t box_cols,
                        uint32_t box_rows, CUtensorMapSwizzle swz) {
    tc_encode_fn enc = tc_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15u) || (cols * 2) % 16) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    return true;
}

template <int KT>
static int tc_launch(const TcParams& p, size_t smem, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_spconv_tc<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) !=
            cudaSuccess)
            return B2ME_ELAUNCH;
        attr_set = true;
    }
    const int clusters = p.total_work < tc_num_sms() / 2 ? p.total_work : tc_num_sms() / 2;
    k_spconv_tc<KT><<<2 * clusters, TC_THREADS, smem, stream>>>(p);  // __cluster_dims__(2,1,1): CTA pairs
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

extern "C" int b2me_spconv_fwd_tc(const void* in1, int Cin1, const void* in2, int Cin2, int64_t V_in,
                                  const void* packed_w,
                                  const int32_t* nbr, const int32_t* perm, const uint32_t* tile_masks, int K,
                                  int64_t V_out, int Cout,
                                  const float* scale, const float* shift, const void* residual, int act, float slope,
                                  void* out, int out_dtype, b2me_stream_t stream) {
    if (!in1 || !packed_w || !out || V_out < 0 || V_in < 0) return B2ME_EINVAL;
    if (Cin2 > 0 && !in2) return B2ME_EINVAL;
    if (!nbr && K != 1) return B2ME_EINVAL;
    if (nbr && !tile_masks) return B2ME_EINVAL;
    if (!b2me_tc_supported(K, Cin1, Cin2, Cout)) return B2ME_EUNSUPPORTED;
    if (out_dtype != B2ME_BF16 && out_dtype != B2ME_F32) return B2ME_EINVAL;
    if (V_out == 0) return B2ME_OK;

    TcParams p;
    p.in1 = reinterpret_cast<const __nv_bfloat16*>(in1);
    p.in2 = reinterpret_cast<const __nv_bfloat16*>(in2);
    p.wpacked = reinterpret_cast<const uint8_t*>(packed_w);
    p.nbr = nbr;
    p.perm = perm;
    p.tile_masks = nbr ? tile_masks : nullptr;
    p.scale = scale;
    p.shift = shift;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.out = out;
    p.V_out = V_out;
    p.Cin1 = Cin1;
    p.Cin2 = Cin2;
    p.nchunk1 = (Cin1 + TC_BK - 1) / TC_BK;
    p.nchunk2 = (Cin2 + TC_BK - 1) / TC_BK;
    p.Cout = Cout;
    p.n_tile = tc_n_tile(K, Cout);
    p.n_ntiles = Cout / p.n_tile;
    p.act = act;
    p.out_dtype = out_dtype;
    p.slope = slope;
    p.b_bytes = (unsigned)(p.n_tile / 2) * 128u;  // each CTA of the pair stages half of every weight tile
    p.acc_bufs = 2 * p.n_tile <= 512 ? 2 : 1;
    int cols = 32;
    while (cols < p.acc_bufs * p.n_tile) cols <<= 1;
    p.tmem_cols = cols;
    const int64_t work = ceil_div64(V_out, 2 * TC_BM) * p.n_ntiles;  // 256-row tile pairs x n-tiles
    if (work > 0x7fffffff) return B2ME_EUNSUPPORTED;
    p.total_work = (int)work;

    const size_t fixed = 1024 /*align slack*/ + (size_t)TC_BM * K * 4 + 16 + (size_t)Cout * 8 + 16 +
                         (size_t)TC_EPI_WARPS * TC_STAGE_OUT_BYTES + 16 * TC_MAX_STAGES + 6 * 8 + 16;
    const size_t stage_bytes = (size_t)TC_GI * (TC_A_BYTES + p.b_bytes);  // a stage holds TC_GI items
    int S = TC_MAX_STAGES;
    while (S >= 2 && fixed + (size_t)S * stage_bytes > TC_MAX_SMEM) --S;
    if (S < 2) return B2ME_EUNSUPPORTED;
    p.debug = 0;
#if defined(B2ME_TC_PROFILE) || defined(B2ME_TC_EXPERIMENT)
    if (const char* e = getenv("B2ME_TC_DEBUG")) p.debug = atoi(e);
    if (const char* e = getenv("B2ME_TC_STAGES")) {  // debug build only: ring-depth experiments
        const int want = atoi(e);
        if (want >= 2 && want < S) S = want;
    }
#endif
    p.stages = S;
    // B2ME_TC_TMA=1 routes the operands through the TMA unit (tile::gather4 rows + 2-D weight boxes completing on the
    // leader's barrier, no relay, no proxy fence). Measured on the same box it runs at the speed of the default
    // cp.async gather + bulk copy + relay path (K27 384->384: 5.02 vs 5.04 ms once the 8 UTMALDG per warp and item
    // issue from uniform registers; whole step 232.2 vs 231.1 ms), so it stays an opt-in alternative that the parity
    // tests also cover.
    static int want_tma = -1;
    if (want_tma < 0) {
        const char* e = getenv("B2ME_TC_TMA");
        want_tma = (e && atoi(e) == 1) ? 1 : 0;
    }
    p.tma = 0;
    if (want_tma && TC_GI == 1 && V_in > 0) {
        const uint64_t w_rows = (uint64_t)K * (p.nchunk1 + p.nchunk2) * (uint64_t)Cout;
        bool ok = tc_make_map(&p.tm_in1, in1, (uint64_t)V_in, (uint64_t)Cin1, TC_BK, 1, CU_TENSOR_MAP_SWIZZLE_128B);
        if (ok && Cin2 > 0)
            ok = tc_make_map(&p.tm_in2, in2, (uint64_t)V_in, (uint64_t)Cin2, TC_BK, 1, CU_TENSOR_MAP_SWIZZLE_128B);
        if (ok)
            ok = tc_make_map(&p.tm_w, packed_w, w_rows, 64, 64, (uint32_t)(p.n_tile / 2), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (!ok) return B2ME_ELAUNCH;
        p.tma = 1;
    }
    size_t smem = fixed + (size_t)S * stage_bytes;
    if (smem < TC_MIN_SMEM) smem = TC_MIN_SMEM;

    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (K == 27) return tc_launch<27>(p, smem, s);
    if (K == 8) return tc_launch<8>(p, smem, s);
    if (K == 1) return tc_launch<1>(p, smem, s);
    return B2ME_EUNSUPPORTED;
}
