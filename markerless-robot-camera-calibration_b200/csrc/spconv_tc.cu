// spconv_tc.cu — K4b: sparse convolution as an output-stationary implicit GEMM on tcgen05 (sm_100a).
//
// PERSISTENT, warp-specialised kernel on CTA PAIRS (cta_group::2): a cluster of two CTAs (one TPC) walks the work
// list (256-row tile pair x n-tile; the n-tiles of one tile pair run back to back on the same cluster) round-robin;
// each CTA gathers the A rows of its own 128-row tile and loads HALF of every weight tile, the leader CTA issues
// M = 256 MMAs that read both CTAs' shared memory, and each CTA's TMEM holds the accumulator of its own 128 rows.
// Halving the per-SM weight traffic leaves shared memory for a deep ring of gathered rows: at n_tile = 384 seven
// 16 KB A stages (random rows from DRAM: long, variable latency) beside three 24 KB B stages (weights from L2).
//
//   operand types ES = 2: bf16 x bf16 (kind::f16), 64 channels per 128-byte chunk, K = 16 per MMA
//                 ES = 4: tf32 x tf32 (kind::tf32) on fp32 rows, 32 channels per 128-byte chunk, K = 8 per MMA; the
//                         rows and the packed weights hold tf32-representable values (rounded to nearest by the
//                         producing epilogue / the weight pack), so the tensor core's truncation is exact
//   work item     2 x 128 output voxels (rows perm[256 t ..]) x n_tile output channels (n_tile <= 384 TMEM columns)
//   reduction     items = (kernel offset k with at least one neighbour in the tile pair) x (128-byte channel chunk)
//   A operand     128 gathered input rows x 128 bytes (= one swizzle row per voxel), cp.async 16-byte pieces
//                 straight into the SWIZZLE_128B K-major smem image, zero-fill for missing neighbours
//   B operand     W[k][chunk] pre-packed on the host side of the ABI into the exact smem image of each CTA's half,
//                 so one cp.async.bulk (UBLKCP) per item brings n_tile/2 x 128 bytes and completes on the stage
//                 mbarrier
//   MMA           one elected thread of the leader CTA issues tcgen05.mma.cta_group::2 (M=256, N<=256), fp32
//                 accumulators stay in TMEM for the whole tile; tcgen05.commit multicasts the stage release /
//                 accumulator-ready arrivals to both CTAs
//   epilogue      tcgen05.ld -> folded BatchNorm scale/shift, residual add, ReLU/LeakyReLU -> bf16 rows, staged
//                 through shared memory so that residual loads and output stores are 64-byte coalesced segments
//   fused head    (b2me_head_fused_tc) K = 1, Cin -> Cout hidden units (n-tiles of <= 256 columns) + activation, then
//                 the small second linear Cout -> C (C <= 16) INSIDE the epilogue: every epilogue lane keeps the C
//                 partial logits of its row in registers across the n-tiles, the two column halves are combined
//                 through shared memory, and only [V, C] logits + the per-row arg-max reach HBM (the hidden
//                 activation never does). model/robotnet_segmentation.py:43-49.
//
//   warps  0-3    gather producers (stage ring runs on across tiles, so the next tile's rows are in flight while
//                 the tensor pipe finishes the current one); optional L2 prefetch of the next offset's rows (off by
//                 default: the deep A ring already covers the latency, measured)
//          4      TMEM alloc + MMA issuer (leader) / stage-full relay to the leader's barrier (peer). The issue
//                 loop is warp-uniform: shfl-broadcast warp index / tile masks / TMEM base, one elect.sync branch
//                 per item, descriptors as 32-bit low words -> UTCHMMA operands live in uniform registers
//          5      weight (B) bulk-copy issuer
//          6-7    kernel-map prefetch: the NEXT item's 128 x K neighbour rows go global -> registers while the
//                 current tile runs, then registers -> smem the moment the producers release the buffer
//          8-15   epilogue (warp w owns TMEM lanes 32 (w % 4) .. and the column half (w - 8) / 4); overlaps the
//                 next tile's gathers, and its MMAs when TMEM has room (see accumulators)
//
//   accumulators  n_tile <= 256: two buffers of n_tile columns alternate, the epilogue of tile i overlaps the MMAs
//                 of tile i + 1.  n_tile = 384 (256 + 128 columns, the K = 27 / 8 layers): the 256-column part lives
//                 in TMEM columns [0, 256) and is released as soon as its four 32-column chunks per warp are read,
//                 the 128-column part alternates between columns [256, 384) and [384, 512): the next tile's MMAs
//                 start after two thirds of the drain instead of after all of it (B2ME_TC_FLAG_NO_ROT128 = single
//                 384-column accumulator, the round-1 layout).
//
//   barriers      full[s]  (128 cp.async-completion arrivals + 1 expect_tx + the peer's relay on the leader)
//                 empty[s] (tcgen05.commit, multicast to both CTAs)
//                 nbr_full (2 prefetch warps)                        nbr_empty    (4 producer warps)
//                 tmem_full[2] (commit) / tmem_empty[3] (8 epilogue warps of each CTA, on the leader)
//
//   B2ME_TC_FLAG_TMA  (what the Python package passes by default) operands through the TMA unit instead: tile::gather4
//                 copies of the gathered rows (absent neighbour = row -1 = out of bounds = zeros) and 2-D boxes of the
//                 packed weights, cta_group::2 with the LEADER's stage barrier as completion target - no relay, no
//                 proxy fence; 3-26 % faster per layer than the cp.async path in round 2 (profiles/r02_ab_medians.md)
//
// The two sources (in1 | in2) implement ME.cat without materialising the concatenation.
#include "common.cuh"
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched with cudaGetDriverEntryPoint, no -lcuda)
#include <stdlib.h>

#define TC_BM 128
#define TC_A_BYTES (TC_BM * 128)
#define TC_THREADS 512
#define TC_EPI_WARPS 8
#ifndef TC_L2_PREFETCH
#define TC_L2_PREFETCH 1  // producers prefetch the next offset's rows into L2
#endif
#ifndef TC_NSPLIT0
#define TC_NSPLIT0 256  // N of the first MMA of a K step when the tile is wider than 256 columns
#endif
#define TC_MAX_STAGES 8
#define TC_MAX_SMEM 232448
#define TC_MIN_SMEM (120 * 1024)  // more than half an SM: one CTA per SM, so a 512-column TMEM alloc never blocks
#define TC_STAGE_OUT_BYTES 2048   // per epilogue warp: 32 rows x 64 bytes
#define TC_HEAD_MAX 16            // classes of the fused second linear

struct TcParams {
    const uint8_t* in1;  // rows of Cin1 * ES bytes
    const uint8_t* in2;
    const uint8_t* wpacked;
    const int32_t* nbr;
    const int32_t* perm;
    const uint32_t* tile_masks;
    const float* scale;
    const float* shift;
    const void* residual;  // bf16 rows (ES = 2) / f32 rows (ES = 4)
    void* out;
    // fused head (null head_w = plain convolution)
    const float* head_w;   // [Cout, head_cp] f32, zero padded columns
    const float* head_b;   // [head_c] or null
    float* head_logits;    // [V_out, head_c]
    uint8_t* head_argmax;  // [V_out] or null
    long long V_out;
    int Cin1, Cin2, nchunk1, nchunk2;
    int Cout, n_tile, n_ntiles;
    int stages;    // A ring: gathered-row stages of 16 KB (= number of full / empty barriers)
    int stages_b;  // B ring: weight stages of b_bytes (<= stages). The gathers come from DRAM (random rows, long
                   // latency), the weights from L2: a short B ring leaves shared memory for a deep A ring
    int act, out_dtype, tmem_cols;
    int acc_bufs;  // 2 when two accumulators fit TMEM (2 n_tile <= 512): the epilogue overlaps the next tile's MMAs
    int rot128;    // n_tile = 384: 256-column part released early, 128-column part alternates between two regions
    int head_c, head_cp;
    float slope;
    unsigned int b_bytes;
    int n_pairs;  // 256-row tile pairs
    int debug;    // debug build only (B2ME_TC_DEBUG): 1 skip the A gathers, 2 skip the B copies, 4 skip the MMAs
    int tma;      // 1: operands come through the TMA unit (gather4 rows / 2-D weight boxes, cta_group::2)
    int pf_mode;  // L2 prefetch of the next offset's rows: 2 none (default), 0 prefetch.global.L2 per 128-byte chunk
                  // (round-1 scheme), 1 one cp.async.bulk.prefetch.L2 per row
    // tensor maps (TMA mode): the two sources as [V_in, Cin] with a (128-byte chunk) x 1-row box (SWIZZLE_128B; rows are
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
// try_wait with a suspend-time hint (ns; compiles to NANOSLEEP.SYNCS between two phase checks): the hardware parks the
// thread until the phase completes or the time is up instead of returning after its short default window. It removes
// the spin instructions of the waiting roles (ncu: issue slots busy 43 % instead of ~60 % on the K = 1 layers) but
// changes no layer's time, so the plain form stays the default (-DTC_WAIT_HINT_NS=20000 to enable).
#ifndef TC_WAIT_HINT_NS
#define TC_WAIT_HINT_NS 0   // measured (run 10, interleaved two-build medians): no gain on any layer shape, K = 27 -0..5 %
#endif
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"((uint32_t)TC_WAIT_HINT_NS)
        : "memory");
    return ok != 0;
}
// predicated 16-byte global store (no branch, no reconvergence point around the store)
__device__ __forceinline__ void st_global_v4_if(uint4* ptr, const uint4& v, bool on) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "@p st.global.v4.b32 [%0], {%1, %2, %3, %4};\n\t}"
        :
        : "l"(ptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((uint32_t)on)
        : "memory");
}
// bounded wait (by time: ~2 s): a protocol bug traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#if TC_WAIT_HINT_NS > 0
    if (mbar_try_wait_hint(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_hint(bar, parity)) {
        if (clock64() - t0 > (1ll << 32)) __trap();
    }
#else
    uint32_t spins = 0;
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
// bulk (TMA engine) prefetch of `bytes` contiguous bytes into L2; address and size multiples of 16
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tc_commit_pair(uint32_t bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"((uint16_t)3)
        : "memory");
}
// one lane of the (converged) warp: true in exactly one lane
__device__ __forceinline__ bool tc_elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\tselp.u32 %0, 1, 0, q;\n\t}" : "=r"(pred)::"memory");
    return pred != 0;
}
// tcgen05.mma of a CTA pair, executed by the calling thread. Descriptors are passed as their low words (start address
// >> 4 | LBO field); the high word of the SWIZZLE_128B K-major descriptor (SBO 1024 B, version 1, swizzle mode 2) is the
// constant 0x40004040. ES = 2: kind::f16 (bf16 operands), ES = 4: kind::tf32.
template <int ES>
__device__ __forceinline__ void tc_mma_lo(uint32_t d, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    if (ES == 2) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d),
            "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "mov.b64 da, {%1, %5};\n\tmov.b64 db, {%2, %5};\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], da, db, %3, p;\n\t}" ::"r"(d),
            "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(0x40004040u)
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
// named barrier of `count` threads (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// Activation of N accumulator values with ONE uniform branch around the loop (a runtime `switch (act)` around every
// element compiled into two uniform branches per value in the fused-head epilogue: ~1400 cycles per 32-column chunk).
// mode 0 none, 1 ReLU (max(x, 0)), 2 LeakyReLU with slope in [0, 1] as max(x, slope x), 3 any other slope.
__device__ __forceinline__ int tc_act_mode(int act, float slope) {
    if (act == B2ME_ACT_RELU) return 1;
    if (act == B2ME_ACT_LEAKY) return (slope >= 0.f && slope <= 1.f) ? 2 : 3;
    return 0;
}
template <int N>
__device__ __forceinline__ void tc_act_n(float (&v)[N], int mode, float slope) {
    if (mode == 1) {
#pragma unroll
        for (int q = 0; q < N; ++q) v[q] = fmaxf(v[q], 0.f);
    } else if (mode == 2) {
#pragma unroll
        for (int q = 0; q < N; ++q) v[q] = fmaxf(v[q], v[q] * slope);
    } else if (mode == 3) {
#pragma unroll
        for (int q = 0; q < N; ++q) v[q] = v[q] > 0.f ? v[q] : v[q] * slope;
    }
}

// fused head: acc[c] += x[q] * W2[col0 + q][c] for the 32 columns of one accumulator chunk; W2 rows of HCP floats in
// shared memory (every lane reads the same address: broadcast)
template <int HC, int NH>
__device__ __forceinline__ void tc_head_accumulate(const float (&x)[32], const float* __restrict__ w_rows, int hcp,
                                                   float (&hacc)[NH]) {
    static_assert(HC <= NH, "head accumulators");
#pragma unroll
    for (int q = 0; q < 32; ++q) {
        float w[(HC + 3) / 4 * 4];
#pragma unroll
        for (int g = 0; g < (HC + 3) / 4; ++g) {
            const float4 t = *reinterpret_cast<const float4*>(w_rows + q * hcp + 4 * g);
            w[4 * g] = t.x; w[4 * g + 1] = t.y; w[4 * g + 2] = t.z; w[4 * g + 3] = t.w;
        }
#pragma unroll
        for (int c = 0; c < HC; ++c) hacc[c] = fmaf(x[q], w[c], hacc[c]);
    }
}

// ------------------------------------------------------------------------------------------ kernel
// N of the MMA instructions of one item: n_tile <= 256 -> one instruction; wider tiles -> TC_NSPLIT0 + the rest.
__host__ __device__ __forceinline__ int tc_n_first(int n_tile) { return n_tile > 256 ? TC_NSPLIT0 : n_tile; }

// KT = kernel volume of the map (27: k3 s1, 8: k2 s2 and its transpose, 1: identity / MinkowskiLinear)
// ES = operand element size (2: bf16, 4: fp32 rows read as tf32); HEAD: fused second linear in the epilogue (K = 1)
template <int KT, int ES, bool HEAD>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_THREADS, 1) k_spconv_tc(const __grid_constant__ TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    constexpr int CPC = 128 / ES;    // channels per 128-byte chunk
    constexpr int EPP = 16 / ES;     // channels per 16-byte piece
    constexpr int KSTEP = 32 / ES;   // channels per MMA (32 bytes of K)

    const int S = p.stages, SB = p.stages_b;
    const uint32_t b_ring = base + (uint32_t)S * TC_A_BYTES;   // B ring follows the A ring
    const int tid = threadIdx.x;
    // shfl-broadcast: the compiler then knows the warp index (hence every role branch) is warp-uniform
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const uint32_t rank = cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer

    // carve: [A ring S x 16 KB][B ring SB x b_bytes][nbr_s 2 x 128*KT i32][scale Cout][shift Cout][head W2 Cout*head_cp][epilogue staging 8 x 2 KB]
    //        [barriers][tmem ptr]
    uint32_t off = (uint32_t)S * TC_A_BYTES + (uint32_t)SB * p.b_bytes;
    // kernel-map rows of the tile being gathered and of the next one (double buffer: the prefetch warps publish tile
    // t + 1 while the producers still read tile t, so the producers never wait for the hand-over)
    int32_t* nbr_s0 = reinterpret_cast<int32_t*>(sm + off);
    off += 2 * TC_BM * KT * 4;
    off = (off + 15u) & ~15u;
    // folded BatchNorm scale / shift, padded by one 32-column chunk of zeros so that the epilogue reads whole chunks
    // with immediate offsets (no per-access clamp) when the last chunk of a tile is narrower than 32 columns
    float* scale_s = reinterpret_cast<float*>(sm + off);
    off += (p.Cout + 32) * 4;
    float* shift_s = reinterpret_cast<float*>(sm + off);
    off += (p.Cout + 32) * 4;
    off = (off + 15u) & ~15u;
    float* head_w_s = reinterpret_cast<float*>(sm + off);
    off += (uint32_t)(HEAD ? p.Cout * p.head_cp * 4 : 0);
    off = (off + 15u) & ~15u;
    const uint32_t stage_out = base + off;
    off += TC_EPI_WARPS * TC_STAGE_OUT_BYTES;
    const uint32_t bar_full = base + off;
    off += 8 * TC_MAX_STAGES;
    const uint32_t bar_empty = base + off;
    off += 8 * TC_MAX_STAGES;
    const uint32_t bar_nbr_full = base + off;   // [2]
    off += 16;
    const uint32_t bar_nbr_empty = base + off;  // [2]
    off += 16;
    const uint32_t bar_tmem_full = base + off;   // [2] one per accumulator buffer
    off += 16;
    const uint32_t bar_tmem_empty = base + off;  // [3] accumulator buffers / (256-column part, 128-column regions 0, 1)
    off += 24;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sm + off);

    // ---- one-time setup
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            // 128 producer threads (cp.async-completion arrivals) + the B expect_tx arrive (+ the peer's relay)
            // TMA mode: the leader's single arrive.expect_tx (all four transfers complete_tx on the leader's barrier)
            mbar_init(bar_full + 8 * s, p.tma ? 1 : (rank == 0 ? 130 : 129));
            mbar_init(bar_empty + 8 * s, 1);  // tcgen05.commit (multicast from the leader)
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(bar_nbr_full + 8 * b, 2);   // 2 prefetch warps
            mbar_init(bar_nbr_empty + 8 * b, 4);  // 4 producer warps
        }
        for (int b = 0; b < 2; ++b) mbar_init(bar_tmem_full + 8 * b, 1);  // tcgen05.commit (multicast from the leader)
        for (int b = 0; b < 3; ++b)
            mbar_init(bar_tmem_empty + 8 * b, 2 * TC_EPI_WARPS);  // epilogue warps of both CTAs (leader's barrier)
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        // cta_group::2 allocation: the same warp of both CTAs issues it
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < p.Cout + 32; i += TC_THREADS) {
        scale_s[i] = i < p.Cout ? (p.scale ? p.scale[i] : 1.f) : 0.f;
        shift_s[i] = i < p.Cout ? (p.shift ? p.shift[i] : 0.f) : 0.f;
    }
    if (HEAD)
        for (int i = tid; i < p.Cout * p.head_cp; i += TC_THREADS) head_w_s[i] = __ldg(p.head_w + i);
    tc_fence_before();
    cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    const int nchunk = p.nchunk1 + p.nchunk2;
    const int G = gridDim.x >> 1;          // clusters
    const int unit0 = blockIdx.x >> 1;     // this cluster's first tile pair
    const int NT = p.n_ntiles;
    // work item `it` of this cluster = (tile pair unit0 + (it / NT) G, n-tile it % NT): the n-tiles of a tile pair run
    // back to back on one cluster (its A rows are re-read from L2; the fused head sums over them in registers)
    auto item_pair = [&](int it) -> int { return unit0 + (it / NT) * G; };
    // offsets (bit k) that at least one row of a 256-row tile pair needs: precomputed per map (b2me_tc_tile_masks),
    // so every role knows a tile's item list without waiting for the kernel-map rows
    auto pair_mask = [&](int tp) -> uint32_t {
        if (!p.tile_masks) return 1u;
        const uint32_t m = __ldg(p.tile_masks + tp);
        return m ? m : 1u;
    };

    if (warp < 4 && p.tma) {
        // =============================== gather producers, TMA mode ===============================
        // warp w owns rows 32 w .. 32 w + 31 of the CTA's tile. Lane l reads the neighbour index of row 32 w + l; per
        // item the warp issues 8 tile::gather4 copies (4 rows x 128 bytes each, straight into the swizzled image;
        // absent neighbours = row -1 = out of bounds = zeros) from shfl-broadcast, i.e. warp-uniform, operands, so each
        // UTMALDG takes its operands from uniform registers without a per-lane serialisation loop.
        const uint32_t full_leader = mapa_u32(bar_full, 0u);
        int ist = 0, iph = 0;
        uint32_t kmask_next = __shfl_sync(0xffffffffu, unit0 < p.n_pairs ? pair_mask(unit0) : 0u, 0);
        for (int it = 0;; ++it) {
            if (item_pair(it) >= p.n_pairs) break;
            const uint32_t kmask = kmask_next;
            const int tpn = item_pair(it + 1);
            if (tpn < p.n_pairs) kmask_next = __shfl_sync(0xffffffffu, pair_mask(tpn), 0);
            const int32_t* nbr_s = nbr_s0 + (it & 1) * (TC_BM * KT);
            mbar_wait(bar_nbr_full + 8 * (it & 1), (uint32_t)(it >> 1) & 1u);
#pragma unroll 1
            for (uint32_t m = kmask; m; m &= m - 1u) {
                const int k = __ffs((int)m) - 1;
                const int id = nbr_s[(32 * warp + lane) * KT + k];
                int ids[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) ids[q] = __shfl_sync(0xffffffffu, id, q);
                if ((m & (m - 1u)) == 0u) {  // last offset of this tile: the kernel-map buffer may be refilled
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_nbr_empty + 8 * (it & 1));
                } else if (TC_L2_PREFETCH && p.pf_mode != 2) {
                    // optional: pull the rows of the NEXT offset into L2 (see the cp.async producers)
                    const uint32_t m2 = m & (m - 1u);
                    const int k2 = __ffs((int)m2) - 1;
                    const int id2 = nbr_s[(32 * warp + lane) * KT + k2];
                    if (id2 >= 0) {
                        if (p.pf_mode == 1) {
                            prefetch_l2_bulk(p.in1 + (long long)id2 * p.Cin1 * ES, (uint32_t)(p.Cin1 * ES));
                            if (p.nchunk2) prefetch_l2_bulk(p.in2 + (long long)id2 * p.Cin2 * ES, (uint32_t)(p.Cin2 * ES));
                        } else {
                            for (int ch = 0; ch < p.nchunk1; ++ch)
                                prefetch_l2(p.in1 + ((long long)id2 * p.Cin1 + ch * CPC) * ES);
                            for (int ch = 0; ch < p.nchunk2; ++ch)
                                prefetch_l2(p.in2 + ((long long)id2 * p.Cin2 + ch * CPC) * ES);
                        }
                    }
                }
#pragma unroll 1
                for (int c = 0; c < nchunk; ++c) {
                    mbar_wait(bar_empty + 8 * ist, (uint32_t)iph ^ 1u);
                    if (tc_elect_one()) {
                        const uint32_t a_s = base + (uint32_t)ist * TC_A_BYTES + (uint32_t)(32 * warp) * 128u;
                        const CUtensorMap* tm = c < p.nchunk1 ? &p.tm_in1 : &p.tm_in2;
                        const int col = (c < p.nchunk1 ? c : c - p.nchunk1) * CPC;
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
        int ist = 0, iph = 0;         // stage / phase of the next item to issue
        PROF_DECL
        const long long t_role0 = clock64();
        (void)t_role0;
        uint32_t kmask_next = unit0 < p.n_pairs ? pair_mask(unit0) : 0u;
        for (int it = 0;; ++it) {
            if (item_pair(it) >= p.n_pairs) break;
            const uint32_t kmask = kmask_next;
            const int tpn = item_pair(it + 1);
            if (tpn < p.n_pairs) kmask_next = pair_mask(tpn);
            const int32_t* nbr_s = nbr_s0 + (it & 1) * (TC_BM * KT);
            PROF(1, mbar_wait(bar_nbr_full + 8 * (it & 1), (uint32_t)(it >> 1) & 1u));
#pragma unroll 1
            for (int k = 0; k < KT; ++k) {
                if (!((kmask >> k) & 1u)) continue;
                // rows of this offset: index clamped to 0 for an absent neighbour (src-size 0 = zero fill, nothing is
                // read), presence kept in a bit mask
                int idx[8];
                uint32_t have = 0u;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int id = nbr_s[(rbase + 16 * i) * KT + k];
                    have |= (id >= 0 ? 1u : 0u) << i;
                    idx[i] = max(id, 0);
                }
                if ((kmask >> k) == 1u) {  // last offset of this tile: the kernel-map buffer may be refilled
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_nbr_empty + 8 * (it & 1));
                } else if (TC_L2_PREFETCH && p.pf_mode == 0) {
                    // optional (B2ME_TC_FLAG_PF_NEAR): pull the rows of the NEXT offset into L2 now, one whole offset
                    // (= nchunk items) ahead of their gather. lane j takes the 128-byte chunks j, j + 8, ... of each of
                    // this thread's 8 rows. Measured (interleaved A/B, same box): no gain over no prefetch at all
                    // (10.6 vs 10.4 ms at tensor stride 1, 5.54 vs 5.27 ms at stride 2), so it is off by default.
                    const int k2 = k + 1 + __ffs((int)(kmask >> (k + 1))) - 1;
                    for (int cj = j; cj < nchunk; cj += 8) {
                        const uint8_t* psrc;
                        int pcin, pcoff;
                        if (cj < p.nchunk1) { psrc = p.in1; pcin = p.Cin1; pcoff = cj * CPC; }
                        else { psrc = p.in2; pcin = p.Cin2; pcoff = (cj - p.nchunk1) * CPC; }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int id = nbr_s[(rbase + 16 * i) * KT + k2];
                            if (id >= 0) prefetch_l2(psrc + ((long long)id * pcin + pcoff) * ES);
                        }
                    }
                } else if (TC_L2_PREFETCH && p.pf_mode == 1) {
                    // the same through the bulk-copy engine: ONE prefetch per row and source covering all its chunks
                    // (thread t of the 128 producers takes row t of the tile)
                    const int k2 = k + 1 + __ffs((int)(kmask >> (k + 1))) - 1;
                    const int id = nbr_s[tid * KT + k2];
                    if (id >= 0) {
                        prefetch_l2_bulk(p.in1 + (long long)id * p.Cin1 * ES, (uint32_t)(p.Cin1 * ES));
                        if (p.nchunk2) prefetch_l2_bulk(p.in2 + (long long)id * p.Cin2 * ES, (uint32_t)(p.Cin2 * ES));
                    }
                }
#pragma unroll 1
                for (int c = 0; c < nchunk; ++c) {
                    PROF(2, mbar_wait(bar_empty + 8 * ist, (uint32_t)iph ^ 1u));
                    PROF_COUNT(7);
                    // addresses in 16-byte pieces (32-bit: tensors up to 64 GB): row id starts at piece id * rp, this
                    // thread's piece of chunk cc of that source is cc * 8 + j
                    const uint8_t* src;
                    int cin, cc;
                    if (c < p.nchunk1) { src = p.in1; cin = p.Cin1; cc = c; }
                    else { src = p.in2; cin = p.Cin2; cc = c - p.nchunk1; }
                    const uint32_t rp = (uint32_t)(cin * ES) >> 4;   // pieces per row (Cin % 16 == 0)
                    const uint32_t pc = (uint32_t)(cc * 8 + j);
                    if (pc < rp && !(p.debug & 1)) {
                        const uint32_t a_s = base + (uint32_t)ist * TC_A_BYTES;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int r = rbase + 16 * i;
                            const uint32_t dst = a_s + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
                            const uint8_t* g = src + ((unsigned long long)((uint32_t)idx[i] * rp + pc) << 4);
                            cp_async_16(dst, g, ((have >> i) & 1u) ? 16u : 0u);
                        }
                    }
                    // asynchronous publication: this thread's arrival fires when its copies have landed, so the
                    // producers run ahead as far as the ring allows and never wait for their own gathers
                    cp_async_mbar_arrive_noinc(bar_full + 8 * ist);
                    if (++ist == S) { ist = 0; iph ^= 1; }
                }
            }
        }
#ifdef B2ME_TC_PROFILE
        prof_[0] = (unsigned long long)(clock64() - t_role0);
        if (warp == 0) PROF_DUMP(0);
#endif
    } else if (warp == 4) {
        // =============================== MMA issuer (leader) / stage relay (peer) ===============================
        // The whole warp walks the item list and waits on the barriers; one elected lane issues the tcgen05 instructions
        // (measured: a wait executed by a lone lane of a divergent warp costs ~210 cycles even on a complete phase).
        if (rank == 0) {
            // Every value below is warp-uniform (shfl-broadcast or derived from kernel parameters), so the compiler
            // keeps stage index, phase and descriptors in uniform registers and emits the UTCHMMA of an item back to
            // back. The tensor pipe queues only a couple of MMAs, so every cycle of issue overhead idles it.
            const int n_a = tc_n_first(p.n_tile), n_b = p.n_tile - n_a;  // N of the one or two MMAs per K step
            // operands K-major, fp32 accumulate, M = 256 (cta_group::2); each CTA's smem holds N/2 rows of B
            constexpr uint32_t FMT = ES == 2 ? 1u : 2u;  // 1 = bf16, 2 = tf32
            const uint32_t idesc_a = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(n_a >> 3) << 17) | (16u << 24);
            const uint32_t idesc_b = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(n_b >> 3) << 17) | (16u << 24);
            // K steps (32 bytes of K) of a full chunk and of the last chunk of each source
            const int ks_last1 = (p.Cin1 - (p.nchunk1 - 1) * CPC) / KSTEP;
            const int ks_last2 = p.nchunk2 ? (p.Cin2 - (p.nchunk2 - 1) * CPC) / KSTEP : 0;
            // low words of the SWIZZLE_128B K-major descriptors of stage 0 (high word is constant, see tc_mma_lo)
            const uint32_t a_lo0 = ((base >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t b_lo0 = ((b_ring >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t a_stage_lo = TC_A_BYTES >> 4, b_stage_lo = p.b_bytes >> 4;
            const uint32_t b2_off_lo = ((uint32_t)(n_a >> 1) * 128u) >> 4;   // second instruction's rows inside a B stage
            int sb = 0;                                                        // B stage of the current item
            const bool need_fence = !p.tma;  // cp.async (generic proxy) writes of A -> visible to the MMA (async proxy)
            int st = 0, ph = 0;
            PROF_DECL
            const long long t_role0 = clock64();
            (void)t_role0;
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            uint32_t kmask_next = __shfl_sync(0xffffffffu, unit0 < p.n_pairs ? pair_mask(unit0) : 0u, 0);
            for (int it = 0;; ++it) {
                if (item_pair(it) >= p.n_pairs) break;
                const uint32_t kmask = kmask_next;
                const int tpn = item_pair(it + 1);
                if (tpn < p.n_pairs) kmask_next = __shfl_sync(0xffffffffu, pair_mask(tpn), 0);
                uint32_t tmem_acc, tmem_acc_b;  // columns of the first / second MMA of a K step
                int fb;                         // tmem_full barrier of this item
                if (p.rot128) {
                    // 256-column part: one region, free once the previous tile's epilogue has read it; 128-column
                    // part: region (it & 1), free once the epilogue of tile it - 2 has read it
                    tmem_acc = tmem_u;
                    tmem_acc_b = tmem_u + 256u + 128u * (uint32_t)(it & 1);
                    fb = 0;
                    if (it >= 1) PROF(2, mbar_wait(bar_tmem_empty, (uint32_t)(it - 1) & 1u));
                    if (it >= 2) PROF(2, mbar_wait(bar_tmem_empty + 8 * (1 + (it & 1)), (uint32_t)((it >> 1) - 1) & 1u));
                } else {
                    // accumulator buffer of this work item and how often it has been used before
                    const int ab = p.acc_bufs == 2 ? (it & 1) : 0;
                    const int au = p.acc_bufs == 2 ? (it >> 1) : it;
                    tmem_acc = tmem_u + (uint32_t)(ab * p.n_tile);
                    tmem_acc_b = tmem_acc + (uint32_t)n_a;
                    fb = ab;
                    if (au > 0)  // both CTAs' epilogues must have drained the previous tile of this buffer
                        PROF(2, mbar_wait(bar_tmem_empty + 8 * ab, (uint32_t)(au - 1) & 1u));
                }
                tc_fence_after();
                uint32_t acc = 0u;
                for (uint32_t m = kmask; m; m &= m - 1u) {  // one pass per kernel offset the tile needs
                    for (int c = 0; c < nchunk; ++c) {
                        int ks = (c == p.nchunk1 - 1) ? ks_last1 : ((c == nchunk - 1) ? ks_last2 : 4);
#ifdef B2ME_TC_PROFILE
                        if (p.debug & 4) ks = 0;
#endif
                        PROF(3, mbar_wait(bar_full + 8 * st, (uint32_t)ph));
                        PROF_COUNT(7);
#ifdef B2ME_TC_PROFILE
                        pt_ = clock64();
#endif
                        tc_fence_after();
                        const uint32_t a_lo = a_lo0 + (uint32_t)st * a_stage_lo;
                        const uint32_t b_lo = b_lo0 + (uint32_t)sb * b_stage_lo;
                        if (tc_elect_one()) {
                            if (need_fence) fence_proxy_async();
                            // the two instructions of a K step share the A slice; both read it from shared memory
                            // (keeping it in the collector buffer, collector::a::fill / lastuse, was measured slower)
                            for (int kk = 0; kk < ks; ++kk) {
                                tc_mma_lo<ES>(tmem_acc, a_lo + 2u * kk, b_lo + 2u * kk, idesc_a, acc);
                                if (n_b)
                                    tc_mma_lo<ES>(tmem_acc_b, a_lo + 2u * kk, b_lo + b2_off_lo + 2u * kk, idesc_b, acc);
                                acc = 1u;
                            }
                            tc_commit_pair(bar_empty + 8 * st);  // frees the stage in both CTAs
                        }
                        acc = ks > 0 ? 1u : acc;
                        __syncwarp();
                        if (++st == S) { st = 0; ph ^= 1; }
                        if (++sb == SB) sb = 0;
#ifdef B2ME_TC_PROFILE
                        prof_[5] += (unsigned long long)(clock64() - pt_);
#endif
                    }
                }
                tc_commit_pair_elect(bar_tmem_full + 8 * fb);  // accumulators of both CTAs are complete
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
            uint32_t kmask_next = unit0 < p.n_pairs ? pair_mask(unit0) : 0u;
            for (int it = 0;; ++it) {
                if (item_pair(it) >= p.n_pairs) break;
                const uint32_t kmask = kmask_next;
                const int tpn = item_pair(it + 1);
                if (tpn < p.n_pairs) kmask_next = pair_mask(tpn);
                const int items = __popc(kmask) * nchunk;
                for (int q = 0; q < items; ++q) {
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
#ifdef B2ME_TC_PROFILE
            prof_[0] = (unsigned long long)(clock64() - t_role0);
            PROF_DUMP(5);
#endif
        }
    } else if (warp == 5) {
        // =============================== weight (B) loader ===============================
        // item i announces its bytes on full[i % S] (the item's A stage) and lands in B stage i % SB, which is free once
        // item i - SB has been consumed: the loader follows the empty barriers with a cursor SB items behind its own
        int st = 0, sb = 0;            // A stage (full barrier) / B stage of the item being loaded
        int wst = 0, wph = 0, lag = 0; // cursor of the item whose consumption frees this B stage
        const uint32_t full_leader = mapa_u32(bar_full, 0u);
        uint32_t kmask_next = unit0 < p.n_pairs ? pair_mask(unit0) : 0u;
        for (int it = 0;; ++it) {
            if (item_pair(it) >= p.n_pairs) break;
            const uint32_t kmask = kmask_next;
            const int tpn = item_pair(it + 1);
            if (tpn < p.n_pairs) kmask_next = pair_mask(tpn);
            const int nt = it % NT;
            for (int k = 0; k < KT; ++k) {
                if (!((kmask >> k) & 1u)) continue;
                for (int c = 0; c < nchunk; ++c) {
                    if (lag == SB) {   // item (this - SB) consumed -> its B stage (= this item's) is free
                        mbar_wait(bar_empty + 8 * wst, (uint32_t)wph);
                        if (++wst == S) { wst = 0; wph ^= 1; }
                    } else {
                        ++lag;
                    }
                    if (lane == 0) {
                        const uint32_t b_s = b_ring + (uint32_t)sb * p.b_bytes;
                        const long long item = ((long long)nt * KT + k) * nchunk + c;
                        if (p.tma) {
                            // leader: one arrive announcing all four transfers of the stage (A and B of both CTAs)
                            if (rank == 0) mbar_arrive_expect_tx(bar_full + 8 * st, 2u * (TC_A_BYTES + p.b_bytes));
                            tma_load_2d_pair(b_s, &p.tm_w, 0, (int)((item * 2 + rank) * (p.n_tile / 2)),
                                             full_leader + 8 * st);
                        } else if (p.debug & 2) {
                            mbar_arrive(bar_full + 8 * st);
                        } else {
                            mbar_arrive_expect_tx(bar_full + 8 * st, p.b_bytes);
                            bulk_copy_g2s(b_s, p.wpacked + (size_t)(item * 2 + rank) * (size_t)p.b_bytes, p.b_bytes,
                                          bar_full + 8 * st);
                        }
                    }
                    __syncwarp();
                    if (++st == S) st = 0;
                    if (++sb == SB) sb = 0;
                }
            }
        }
    } else if (warp < 8) {
        // =============================== kernel-map prefetch ===============================
        // warp wl stages rows 64 wl .. 64 wl + 63 of every item's tile: 64 x KT entries, 2 KT per lane, held in
        // registers until the producers have released the (single) smem buffer.
        const int wl = warp - 6;
        constexpr int NJ = 2 * KT;
        for (int it = 0;; ++it) {
            const int tp = item_pair(it);
            if (tp >= p.n_pairs) break;
            const long long row0 = ((long long)tp * 2 + rank) * TC_BM + 64 * wl;
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
            int32_t* nbr_s = nbr_s0 + (it & 1) * (TC_BM * KT);
            if (it >= 2) mbar_wait(bar_nbr_empty + 8 * (it & 1), (uint32_t)((it >> 1) - 1) & 1u);
#pragma unroll
            for (int jj = 0; jj < NJ; ++jj) nbr_s[64 * wl * KT + lane + 32 * jj] = v[jj];
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_nbr_full + 8 * (it & 1));
        }
    } else {
        // =============================== epilogue ===============================
        // 8 warps: warp e handles TMEM lanes 32 (e & 3) .. (a warp may only touch the lane quarter warp_id % 4) and
        // the column half e >> 2 of the tile. One tcgen05.ld of 32 columns per chunk, waited for at once: splitting a
        // chunk into two 16-column loads with two register sets (the next load in flight during the arithmetic) was
        // measured slower (head 3.68 -> 4.44 ms, K = 8 1.77 -> 1.88 ms, same box, interleaved): TMEM reads 64 B/cycle
        // with 12 cycles of latency, the chunk time is instruction issue, not TMEM latency.
        const int ew = warp - 8;
        const int we = ew & 3, chalf = ew >> 2;
        const int act_mode = tc_act_mode(p.act, p.slope);
        const uint32_t lane_addr0 = tmem_base + ((uint32_t)(we * 32) << 16);
        const uint32_t stg = stage_out + (uint32_t)ew * TC_STAGE_OUT_BYTES;
        // staging image: row r (0..31) = 64 bytes, 16-byte piece q stored at q ^ ((r >> 1) & 3): conflict-free for both
        // the row-per-lane view (lane = row) and the coalesced view (4 lanes per row, 8 rows per access)
        const uint32_t own = stg + (uint32_t)lane * 64u;
        const int own_sw = (lane >> 1) & 3;
        const int crow = lane >> 2, cq = lane & 3;  // coalesced view: rows crow + 8 m, piece cq
        // this warp's columns: up to two runs of 32-column chunks, [s0_lo, s0_hi) then [s1_lo, s1_hi)
        int s0_lo, s0_hi, s1_lo, s1_hi;
        if (p.rot128) {  // n_tile = 384: its half of the 256-column part first, then its half of the 128-column part
            s0_lo = chalf ? 128 : 0;   s0_hi = s0_lo + 128;
            s1_lo = chalf ? 320 : 256; s1_hi = s1_lo + 64;
        } else {
            const int nch = (p.n_tile + 31) >> 5;
            s0_lo = chalf ? ((nch + 1) >> 1) * 32 : 0;
            s0_hi = chalf ? p.n_tile : min(p.n_tile, ((nch + 1) >> 1) * 32);
            s1_lo = s1_hi = 0;
        }
        const int n_s0 = s0_hi > s0_lo ? (s0_hi - s0_lo + 31) >> 5 : 0;
        const int n_s1 = s1_hi > s1_lo ? (s1_hi - s1_lo + 31) >> 5 : 0;
        const int n_my = n_s0 + n_s1;                       // chunks of this warp per item
        auto chunk_col = [&](int ci) -> int { return ci < n_s0 ? s0_lo + 32 * ci : s1_lo + 32 * (ci - n_s0); };
        float hacc[HEAD ? TC_HEAD_MAX : 1];  // fused head: partial logits of this lane's row over this warp's columns
#pragma unroll
        for (int c = 0; c < (HEAD ? TC_HEAD_MAX : 1); ++c) hacc[c] = 0.f;
        PROF_DECL
        for (int it = 0;; ++it) {
            const int tile_p = item_pair(it);
            if (tile_p >= p.n_pairs) break;
            const int nt = it % NT;
            const int n0 = nt * p.n_tile;
            // accumulator columns / barriers of this item (see the MMA warp)
            uint32_t lane_addr, col_b_off;
            uint32_t bar_tf, bar_te_a, bar_te_b, tf_par;
            if (p.rot128) {
                lane_addr = lane_addr0;
                col_b_off = 128u * (uint32_t)(it & 1);   // TMEM column of tile column c >= 256: c + col_b_off
                bar_tf = bar_tmem_full;
                tf_par = (uint32_t)it & 1u;
                bar_te_a = bar_tmem_empty;
                bar_te_b = bar_tmem_empty + 8 * (1 + (it & 1));
            } else {
                const int ab = p.acc_bufs == 2 ? (it & 1) : 0;
                lane_addr = lane_addr0 + (uint32_t)(ab * p.n_tile);
                col_b_off = 0u;
                bar_tf = bar_tmem_full + 8 * ab;
                tf_par = (uint32_t)(p.acc_bufs == 2 ? (it >> 1) : it) & 1u;
                bar_te_a = bar_te_b = bar_tmem_empty + 8 * ab;
            }
            auto tmem_col = [&](int cb) -> uint32_t { return lane_addr + (uint32_t)cb + (cb >= 256 ? col_b_off : 0u); };
            // release of the accumulator part(s) after chunk ci of this warp has been read out of TMEM
            auto release_after = [&](int ci) {
                const bool last = ci == n_my - 1;
                const bool last_a = p.rot128 && ci == n_s0 - 1;
                if (last || last_a) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(last ? bar_te_b : bar_te_a, 0u);  // the leader's barrier
                }
            };
            const long long slot = ((long long)tile_p * 2 + rank) * TC_BM + we * 32 + lane;
            const int row = (slot < p.V_out) ? (p.perm ? __ldg(p.perm + slot) : (int)slot) : -1;

            if constexpr (HEAD) {
                // ---------------- fused head: hidden activation -> partial logits, nothing of the tile is stored
                PROF(1, mbar_wait(bar_tf, tf_par));
                tc_fence_after();
                if (n_my == 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_te_b, 0u);
                }
                for (int ci = 0; ci < n_my; ++ci) {
                    const int cb = chunk_col(ci);
                    uint32_t r[32];
                    tmem_ld_x32(tmem_col(cb), r);  // head mode: n_tile % 64 == 0, every chunk is 32 columns wide
                    tmem_ld_wait();
                    release_after(ci);
                    float x[32];
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const int col = n0 + cb + q4 * 4;
                        const float4 sc = *reinterpret_cast<const float4*>(scale_s + col);
                        const float4 sh = *reinterpret_cast<const float4*>(shift_s + col);
                        x[q4 * 4 + 0] = __uint_as_float(r[q4 * 4 + 0]) * sc.x + sh.x;
                        x[q4 * 4 + 1] = __uint_as_float(r[q4 * 4 + 1]) * sc.y + sh.y;
                        x[q4 * 4 + 2] = __uint_as_float(r[q4 * 4 + 2]) * sc.z + sh.z;
                        x[q4 * 4 + 3] = __uint_as_float(r[q4 * 4 + 3]) * sc.w + sh.w;
                    }
                    tc_act_n(x, act_mode, p.slope);
                    // the hidden activation is a bf16 (tf32-rounded fp32) tensor in the unfused data path: same rounding
                    if (ES == 2) {
#pragma unroll
                        for (int q = 0; q < 32; q += 2) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(x[q], x[q + 1]);
                            const uint32_t u = *reinterpret_cast<uint32_t*>(&h);
                            x[q] = __uint_as_float(u << 16);
                            x[q + 1] = __uint_as_float(u & 0xFFFF0000u);
                        }
                    } else {
#pragma unroll
                        for (int q = 0; q < 32; ++q) x[q] = round_tf32(x[q]);
                    }
                    const float* w_rows = head_w_s + (n0 + cb) * p.head_cp;
                    switch (p.head_c) {   // exact FMA count for the class counts of the reference's heads
                        case 2: tc_head_accumulate<2>(x, w_rows, p.head_cp, hacc); break;
                        case 3: tc_head_accumulate<3>(x, w_rows, p.head_cp, hacc); break;
                        case 4: tc_head_accumulate<4>(x, w_rows, p.head_cp, hacc); break;
                        case 6: tc_head_accumulate<6>(x, w_rows, p.head_cp, hacc); break;
                        case 10: tc_head_accumulate<10>(x, w_rows, p.head_cp, hacc); break;
                        default:
                            if (p.head_cp <= 8) tc_head_accumulate<8>(x, w_rows, p.head_cp, hacc);
                            else tc_head_accumulate<16>(x, w_rows, p.head_cp, hacc);
                            break;
                    }
                }
                if (nt == NT - 1) {
                    // combine the two column halves of every row: the upper-half warp hands its sums over through its
                    // staging block, the lower-half warp adds (fixed order: lower + upper + bias), stores, arg-maxes
                    if (chalf == 1) {
#pragma unroll
                        for (int g = 0; g < TC_HEAD_MAX / 4; ++g)
                            st_shared_v4(own + 16u * g, make_uint4(__float_as_uint(hacc[4 * g]), __float_as_uint(hacc[4 * g + 1]),
                                                                  __float_as_uint(hacc[4 * g + 2]), __float_as_uint(hacc[4 * g + 3])));
                    }
                    named_bar_sync(1 + we, 64);
                    if (chalf == 0) {
                        const uint32_t partner = stage_out + (uint32_t)(ew + 4) * TC_STAGE_OUT_BYTES + (uint32_t)lane * 64u;
                        float tot[TC_HEAD_MAX];
#pragma unroll
                        for (int g = 0; g < TC_HEAD_MAX / 4; ++g) {
                            const uint4 o = ld_shared_v4(partner + 16u * g);
                            tot[4 * g + 0] = hacc[4 * g + 0] + __uint_as_float(o.x);
                            tot[4 * g + 1] = hacc[4 * g + 1] + __uint_as_float(o.y);
                            tot[4 * g + 2] = hacc[4 * g + 2] + __uint_as_float(o.z);
                            tot[4 * g + 3] = hacc[4 * g + 3] + __uint_as_float(o.w);
                        }
                        if (row >= 0) {
                            float best = -INFINITY;
                            int besti = 0;
#pragma unroll
                            for (int c = 0; c < TC_HEAD_MAX; ++c) {
                                if (c < p.head_c) {
                                    const float v = tot[c] + (p.head_b ? __ldg(p.head_b + c) : 0.f);
                                    p.head_logits[(long long)row * p.head_c + c] = v;
                                    if (v > best) {  // strict: lowest index wins ties (torch.max semantics)
                                        best = v;
                                        besti = c;
                                    }
                                }
                            }
                            if (p.head_argmax) p.head_argmax[row] = (uint8_t)besti;
                        }
                    }
                    named_bar_sync(1 + we, 64);  // the staging block may be rewritten
#pragma unroll
                    for (int c = 0; c < TC_HEAD_MAX; ++c) hacc[c] = 0.f;
                }
            } else if (ES == 2 && p.out_dtype == B2ME_BF16) {
                int crows[4];  // output rows of the coalesced view
#pragma unroll
                for (int m = 0; m < 4; ++m) crows[m] = __shfl_sync(0xffffffffu, row, crow + 8 * m);
                const uint4* resp = reinterpret_cast<const uint4*>(p.residual);
                uint4* outp = reinterpret_cast<uint4*>(p.out);
                // 16-byte piece index of (row crows[m], column n0 + 8 cq) in the [V_out, Cout] bf16 tensors (32-bit:
                // the launcher refuses tensors of 64 GB and more); a chunk adds cb / 8
                uint32_t o16[4];
#pragma unroll
                for (int m = 0; m < 4; ++m)
                    o16[m] = (uint32_t)max(crows[m], 0) * (uint32_t)(p.Cout >> 3) + (uint32_t)((n0 >> 3) + cq);
                uint4 R[4];
                auto load_res = [&](int cb) {  // coalesced: 64-byte segment of 8 rows per access
                    const int cw = min(32, p.n_tile - cb);
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        R[m] = make_uint4(0u, 0u, 0u, 0u);
                        if (crows[m] >= 0 && cq * 8 < cw) R[m] = __ldg(resp + (o16[m] + (uint32_t)(cb >> 3)));
                    }
                };
                if (resp && n_my > 0) load_res(chunk_col(0));
                PROF(1, mbar_wait(bar_tf, tf_par));
                tc_fence_after();
#ifdef B2ME_TC_PROFILE
                const long long t_epi0 = clock64();
#endif
                if (n_my == 0) {  // narrow tile: this warp has no columns
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_te_b, 0u);
                }
                for (int ci = 0; ci < n_my; ++ci) {
                    const int cb = chunk_col(ci);
                    const int cw = min(32, p.n_tile - cb);
                    uint32_t r[32];
                    if (cw == 32) {
                        tmem_ld_x32(tmem_col(cb), r);
                    } else {
                        tmem_ld_x16(tmem_col(cb), r);
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
                        if (ci + 1 < n_my) load_res(chunk_col(ci + 1));  // in flight during this chunk's math and stores
                    }
                    tmem_ld_wait();
                    release_after(ci);  // this warp's part of the accumulator (region) is read: release it
                    float v[32];
                    const float* scp = scale_s + n0 + cb;  // padded arrays: columns past Cout read zeros
                    const float* shp = shift_s + n0 + cb;
#pragma unroll
                    for (int q4 = 0; q4 < 8; ++q4) {
                        const float4 sc = *reinterpret_cast<const float4*>(scp + q4 * 4);
                        const float4 sh = *reinterpret_cast<const float4*>(shp + q4 * 4);
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
                    tc_act_n(v, act_mode, p.slope);
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint32_t wv[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(v[q4 * 8 + 2 * e], v[q4 * 8 + 2 * e + 1]);
                            wv[e] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        st_shared_v4(own + (uint32_t)((q4 ^ own_sw) << 4), make_uint4(wv[0], wv[1], wv[2], wv[3]));
                    }
                    __syncwarp();
                    // all four staged pieces first, then four predicated stores: one shared-memory latency per chunk
                    // instead of four load -> store round trips through the same registers
                    uint4 o[4];
#pragma unroll
                    for (int m = 0; m < 4; ++m) {
                        const int rr = crow + 8 * m;
                        o[m] = ld_shared_v4(stg + (uint32_t)rr * 64u + (uint32_t)((cq ^ ((rr >> 1) & 3)) << 4));
                    }
#pragma unroll
                    for (int m = 0; m < 4; ++m)
                        st_global_v4_if(outp + (o16[m] + (uint32_t)(cb >> 3)), o[m], crows[m] >= 0 && cq * 8 < cw);
                    __syncwarp();  // staging is reused by the next chunk
                }
#ifdef B2ME_TC_PROFILE
                prof_[2] += (unsigned long long)(clock64() - t_epi0);
                ++prof_[7];
#endif
            } else if (ES == 4) {
                // fp32 rows of the tf32 mode, staged like the bf16 rows: 16-column sub-chunks (64 bytes per row), so
                // residual loads and output stores are 64-byte row segments, 8 rows per warp access. B2ME_TF32: the
                // stored values are rounded to tf32 (they feed the next tf32 MMA).
                int crows[4];
#pragma unroll
                for (int m = 0; m < 4; ++m) crows[m] = __shfl_sync(0xffffffffu, row, crow + 8 * m);
                const float* resp = reinterpret_cast<const float*>(p.residual);
                float* outp = reinterpret_cast<float*>(p.out);
                mbar_wait(bar_tf, tf_par);
                tc_fence_after();
                if (n_my == 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_te_b, 0u);
                }
                for (int ci = 0; ci < n_my; ++ci) {
                    const int cb32 = chunk_col(ci);
                    const int nsub = min(32, p.n_tile - cb32) >> 4;   // n_tile % 16 == 0: one or two 16-column sub-chunks
                    for (int hs = 0; hs < nsub; ++hs) {
                        const int cb = cb32 + 16 * hs;
                        uint32_t r[16];
                        tmem_ld_x16(tmem_col(cb), r);
                        uint4 R[4];
                        if (resp) {
#pragma unroll
                            for (int m = 0; m < 4; ++m) {
                                R[m] = make_uint4(0u, 0u, 0u, 0u);
                                if (crows[m] >= 0)
                                    R[m] = __ldg(reinterpret_cast<const uint4*>(resp + (long long)crows[m] * p.Cout + n0 + cb +
                                                                                cq * 4));
                            }
                        }
                        tmem_ld_wait();
                        if (hs == nsub - 1) release_after(ci);
                        float v[16];
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4) {
                            const float4 sc = *reinterpret_cast<const float4*>(scale_s + n0 + cb + q4 * 4);
                            const float4 sh = *reinterpret_cast<const float4*>(shift_s + n0 + cb + q4 * 4);
                            v[q4 * 4 + 0] = __uint_as_float(r[q4 * 4 + 0]) * sc.x + sh.x;
                            v[q4 * 4 + 1] = __uint_as_float(r[q4 * 4 + 1]) * sc.y + sh.y;
                            v[q4 * 4 + 2] = __uint_as_float(r[q4 * 4 + 2]) * sc.z + sh.z;
                            v[q4 * 4 + 3] = __uint_as_float(r[q4 * 4 + 3]) * sc.w + sh.w;
                        }
                        if (resp) {
#pragma unroll
                            for (int m = 0; m < 4; ++m) {
                                const int rr = crow + 8 * m;
                                st_shared_v4(stg + (uint32_t)rr * 64u + (uint32_t)((cq ^ ((rr >> 1) & 3)) << 4), R[m]);
                            }
                            __syncwarp();
#pragma unroll
                            for (int q4 = 0; q4 < 4; ++q4) {
                                const uint4 a = ld_shared_v4(own + (uint32_t)((q4 ^ own_sw) << 4));
                                v[q4 * 4 + 0] += __uint_as_float(a.x);
                                v[q4 * 4 + 1] += __uint_as_float(a.y);
                                v[q4 * 4 + 2] += __uint_as_float(a.z);
                                v[q4 * 4 + 3] += __uint_as_float(a.w);
                            }
                            __syncwarp();  // every lane has read its residual row before the buffer takes the outputs
                        }
                        tc_act_n(v, act_mode, p.slope);
                        if (p.out_dtype == B2ME_TF32) {
#pragma unroll
                            for (int q = 0; q < 16; ++q) v[q] = round_tf32(v[q]);
                        }
#pragma unroll
                        for (int q4 = 0; q4 < 4; ++q4)
                            st_shared_v4(own + (uint32_t)((q4 ^ own_sw) << 4),
                                         make_uint4(__float_as_uint(v[q4 * 4]), __float_as_uint(v[q4 * 4 + 1]),
                                                    __float_as_uint(v[q4 * 4 + 2]), __float_as_uint(v[q4 * 4 + 3])));
                        __syncwarp();
#pragma unroll
                        for (int m = 0; m < 4; ++m) {
                            const int rr = crow + 8 * m;
                            const uint4 o = ld_shared_v4(stg + (uint32_t)rr * 64u + (uint32_t)((cq ^ ((rr >> 1) & 3)) << 4));
                            if (crows[m] >= 0)
                                *reinterpret_cast<uint4*>(outp + (long long)crows[m] * p.Cout + n0 + cb + cq * 4) = o;
                        }
                        __syncwarp();  // staging is reused by the next sub-chunk
                    }
                }
            } else {
                // fp32 rows from bf16 operands (parity tests only): row-per-lane stores of 16-column chunks split
                // between the two column halves.
                mbar_wait(bar_tf, tf_par);
                tc_fence_after();
                const int n16 = p.n_tile >> 4;
                const int q_lo = chalf ? (n16 + 1) >> 1 : 0, q_hi = chalf ? n16 : (n16 + 1) >> 1;
                if (q_lo >= q_hi) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive_remote(bar_te_b, 0u);
                }
                for (int qc = q_lo; qc < q_hi; ++qc) {
                    const int cb = qc * 16;
                    uint32_t r[16];
                    tmem_ld_x16(lane_addr + (uint32_t)cb, r);
                    tmem_ld_wait();
                    if (qc + 1 >= q_hi) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_remote(bar_te_b, 0u);
                    }
                    if (row < 0) continue;
                    float v[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        v[q] = __uint_as_float(r[q]) * scale_s[n0 + cb + q] + shift_s[n0 + cb + q];
                    const long long o = (long long)row * p.Cout + n0 + cb;
                    if (p.residual) {
                        if (ES == 4) {
                            const float4* rp = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.residual) + o);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const float4 ra = __ldg(rp + q);
                                v[4 * q] += ra.x; v[4 * q + 1] += ra.y; v[4 * q + 2] += ra.z; v[4 * q + 3] += ra.w;
                            }
                        } else {
                            const uint4* rp = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.residual) + o);
                            const uint4 ra = __ldg(rp), rb = __ldg(rp + 1);
                            const uint32_t wv[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                v[2 * q] += __uint_as_float(wv[q] << 16);
                                v[2 * q + 1] += __uint_as_float(wv[q] & 0xFFFF0000u);
                            }
                        }
                    }
                    tc_act_n(v, act_mode, p.slope);
                    if (p.out_dtype == B2ME_TF32) {
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] = round_tf32(v[q]);
                    }
                    float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
#pragma unroll
                    for (int q = 0; q < 4; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
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
#ifdef TC_K8_SPLIT   // experiment: K = 8 like K = 1 (two 192-column accumulators instead of 256 + 128 rotating)
    if (K == 27 && Cout <= 384) return Cout;
#else
    if (K != 1 && Cout <= 384) return Cout;
#endif
    if (Cout % 256 == 0) return 256;
    if (Cout % 192 == 0) return 192;
    if (Cout % 128 == 0) return 128;
    return Cout <= 384 ? Cout : 0;
}
static int tc_elem_size(int op_dtype) { return op_dtype == B2ME_BF16 ? 2 : (op_dtype == B2ME_TF32 ? 4 : 0); }

extern "C" int b2me_tc_supported(int K, int Cin1, int Cin2, int Cout) {
    if ((K != 1 && K != 8 && K != 27) || Cin1 < 16 || Cin2 < 0 || Cout < 16 || Cout > 1024) return 0;
    if (Cin1 % 16 || Cin2 % 16 || Cout % 16) return 0;
    const int nt = tc_n_tile(K, Cout);
    if (nt <= 0 || Cout % nt) return 0;
    if (nt > 256 && ((nt - TC_NSPLIT0) % 16 || nt - TC_NSPLIT0 < 32)) return 0;
    return 1;
}

extern "C" size_t b2me_tc_packed_bytes(int K, int Cin1, int Cin2, int Cout, int op_dtype) {
    const int es = tc_elem_size(op_dtype);
    if (!es || !b2me_tc_supported(K, Cin1, Cin2, Cout)) return 0;
    const int cpc = 128 / es;
    const int nchunk = (Cin1 + cpc - 1) / cpc + (Cin2 + cpc - 1) / cpc;
    return (size_t)K * nchunk * (size_t)Cout * 128;
}

// one thread per 16-byte piece of the packed image
template <int ES>
__global__ void k_tc_pack(const float* __restrict__ W, int K, int Cin1, int Cin2, int Cout, int n_tile, int nchunk1,
                          int nchunk2, uint4* __restrict__ packed, long long total_pieces) {
    constexpr int CPC = 128 / ES, EPP = 16 / ES;
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
    if (c < nchunk1) { cin_base = c * CPC; cin_end = Cin1; }
    else { cin_base = Cin1 + (c - nchunk1) * CPC; cin_end = Cin1 + Cin2; }
    const int Cin = Cin1 + Cin2;
    float f[EPP];
#pragma unroll
    for (int e = 0; e < EPP; ++e) {
        const int cin = cin_base + j * EPP + e;
        f[e] = (cin < cin_end) ? W[((long long)k * Cin + cin) * Cout + (nt * n_tile + n)] : 0.f;
    }
    uint32_t w[4];
    if (ES == 2) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            __nv_bfloat162 v = __floats2bfloat162_rn(f[(2 * e) % EPP], f[(2 * e + 1) % EPP]);
            w[e] = *reinterpret_cast<uint32_t*>(&v);
        }
    } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = __float_as_uint(round_tf32(f[e % EPP]));
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

extern "C" int b2me_tc_pack_weights(const float* W, int K, int Cin1, int Cin2, int Cout, int op_dtype, void* packed,
                                    b2me_stream_t stream) {
    if (!W || !packed) return B2ME_EINVAL;
    const int es = tc_elem_size(op_dtype);
    if (!es || !b2me_tc_supported(K, Cin1, Cin2, Cout)) return B2ME_EUNSUPPORTED;
    const int nt = tc_n_tile(K, Cout);
    const int cpc = 128 / es;
    const int nchunk1 = (Cin1 + cpc - 1) / cpc, nchunk2 = (Cin2 + cpc - 1) / cpc;
    const long long total = (long long)(Cout / nt) * K * (nchunk1 + nchunk2) * nt * 8;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (es == 2)
        k_tc_pack<2><<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(W, K, Cin1, Cin2, Cout, nt, nchunk1, nchunk2,
                                                                     reinterpret_cast<uint4*>(packed), total);
    else
        k_tc_pack<4><<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(W, K, Cin1, Cin2, Cout, nt, nchunk1, nchunk2,
                                                                     reinterpret_cast<uint4*>(packed), total);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// SM count of the CURRENT device (one cached entry per device ordinal; a process may drive several GPUs)
static int tc_num_sms() {
    static int sms[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return B2ME_NUM_SMS;
    if (!sms[dev]) {
        int n = 0;
        sms[dev] = (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
                       ? n : B2ME_NUM_SMS;
    }
    return sms[dev];
}

// cuTensorMapEncodeTiled through the runtime's driver entry-point lookup (no link-time dependency on libcuda)
typedef CUresult (*tc_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tc_encode_fn tc_encoder() {
    static tc_encode_fn fn = nullptr;  // a process-wide function pointer of the driver, not a per-device fact
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
// 2-D tensor [rows, cols] of es-byte elements (row pitch = cols), box = box_cols x box_rows
static bool tc_make_map(CUtensorMap* tm, const void* base, int es, uint64_t rows, uint64_t cols, uint32_t box_cols,
                        uint32_t box_rows, CUtensorMapSwizzle swz) {
    tc_encode_fn enc = tc_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15u) || (cols * es) % 16) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * es};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(tm, es == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
            const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    return true;
}

template <int KT, int ES, bool HEAD>
static int tc_launch(const TcParams& p, size_t smem, cudaStream_t stream) {
    // per-device function attribute: set on every launch (cheap), so a process that drives several GPUs works
    if (cudaFuncSetAttribute(k_spconv_tc<KT, ES, HEAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) !=
        cudaSuccess)
        return B2ME_ELAUNCH;
    const int half = tc_num_sms() / 2;
    const int clusters = p.n_pairs < half ? p.n_pairs : half;
    k_spconv_tc<KT, ES, HEAD><<<2 * clusters, TC_THREADS, smem, stream>>>(p);  // __cluster_dims__(2,1,1): CTA pairs
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

struct TcHead {
    const float* w;
    const float* b;
    int c;
    float* logits;
    uint8_t* argmax;
};

static int tc_run(const void* in1, int Cin1, const void* in2, int Cin2, int64_t V_in, int op_dtype, const void* packed_w,
                  const int32_t* nbr, const int32_t* perm, const uint32_t* tile_masks, int K, int64_t V_out, int Cout,
                  const float* scale, const float* shift, const void* residual, int act, float slope, void* out,
                  int out_dtype, const TcHead* head, int flags, b2me_stream_t stream) {
    const int es = tc_elem_size(op_dtype);
    if (!es) return B2ME_EINVAL;
    if (!in1 || !packed_w || V_out < 0 || V_in < 0) return B2ME_EINVAL;
    if (!head && !out) return B2ME_EINVAL;
    if (Cin2 > 0 && !in2) return B2ME_EINVAL;
    if (!nbr && K != 1) return B2ME_EINVAL;
    if (nbr && !tile_masks) return B2ME_EINVAL;
    if (!b2me_tc_supported(K, Cin1, Cin2, Cout)) return B2ME_EUNSUPPORTED;
    // the gather producers address the sources in 16-byte pieces with 32-bit arithmetic: each source < 64 GB
    if (((unsigned long long)V_in * (unsigned long long)(Cin1 > Cin2 ? Cin1 : Cin2) * (unsigned)es) >> 36) return B2ME_EUNSUPPORTED;
    if (((unsigned long long)V_out * (unsigned long long)Cout * 4ull) >> 36) return B2ME_EUNSUPPORTED;  // epilogue: same
    if (!head) {
        if (es == 2 && out_dtype != B2ME_BF16 && out_dtype != B2ME_F32) return B2ME_EINVAL;
        if (es == 4 && out_dtype != B2ME_TF32 && out_dtype != B2ME_F32) return B2ME_EINVAL;
    }
    if (V_out == 0) return B2ME_OK;

    TcParams p;
    const int cpc = 128 / es;
    p.in1 = reinterpret_cast<const uint8_t*>(in1);
    p.in2 = reinterpret_cast<const uint8_t*>(in2);
    p.wpacked = reinterpret_cast<const uint8_t*>(packed_w);
    p.nbr = nbr;
    p.perm = perm;
    p.tile_masks = nbr ? tile_masks : nullptr;
    p.scale = scale;
    p.shift = shift;
    p.residual = residual;
    p.out = out;
    p.head_w = nullptr;
    p.head_b = nullptr;
    p.head_logits = nullptr;
    p.head_argmax = nullptr;
    p.head_c = p.head_cp = 0;
    p.V_out = V_out;
    p.Cin1 = Cin1;
    p.Cin2 = Cin2;
    p.nchunk1 = (Cin1 + cpc - 1) / cpc;
    p.nchunk2 = (Cin2 + cpc - 1) / cpc;
    p.Cout = Cout;
    p.n_tile = tc_n_tile(K, Cout);
    p.n_ntiles = Cout / p.n_tile;
    p.act = act;
    p.out_dtype = out_dtype;
    p.slope = slope;
    p.b_bytes = (unsigned)(p.n_tile / 2) * 128u;  // each CTA of the pair stages half of every weight tile
    p.acc_bufs = 2 * p.n_tile <= 512 ? 2 : 1;
    size_t head_bytes = 0;
    if (head) {
        if (!head->w || !head->logits || head->c < 1 || head->c > TC_HEAD_MAX || p.n_tile % 64 || residual)
            return B2ME_EINVAL;
        p.head_w = head->w;
        p.head_b = head->b;
        p.head_logits = head->logits;
        p.head_argmax = head->argmax;
        p.head_c = head->c;
        p.head_cp = (head->c + 3) / 4 * 4;
        head_bytes = (size_t)Cout * p.head_cp * 4;
    }
    // n_tile = 256 + 128 with bf16 rows out: early release of the 256-column part + alternating 128-column regions
    p.rot128 = (!head && p.n_tile == 384 && TC_NSPLIT0 == 256 && (out_dtype == B2ME_BF16 || es == 4) &&
                !(flags & B2ME_TC_FLAG_NO_ROT128)) ? 1 : 0;
    int cols = 32;
    while (cols < (p.rot128 ? 512 : p.acc_bufs * p.n_tile)) cols <<= 1;
    p.tmem_cols = cols;
    const int64_t pairs = ceil_div64(V_out, 2 * TC_BM);  // 256-row tile pairs
    if (pairs * p.n_ntiles > 0x7fffffff) return B2ME_EUNSUPPORTED;
    p.n_pairs = (int)pairs;

    const size_t fixed = 1024 /*align slack*/ + 2 * (size_t)TC_BM * K * 4 + 16 + (size_t)(Cout + 32) * 8 + 16 + head_bytes + 16 +
                         (size_t)TC_EPI_WARPS * TC_STAGE_OUT_BYTES + 16 * TC_MAX_STAGES + 32 + 16 + 24 + 16;
    // B ring: 3 stages of weights (L2-resident, short latency; A/B override in flags bits 8-10), A ring: as many
    // 16 KB stages of gathered rows (DRAM, long latency) as the rest of shared memory holds
    int SB = (flags >> 8) & 7;
    if (SB == 0) SB = K == 27 ? 3 : 4;   // measured (interleaved A/B): K = 8 384->384 1.55 vs 1.71 ms with 4 vs 3 stages
    if (SB < 2) SB = 2;
    int S = TC_MAX_STAGES;
    while (S >= 2 && fixed + (size_t)S * TC_A_BYTES + (size_t)(SB < S ? SB : S) * p.b_bytes > TC_MAX_SMEM) --S;
    if (S < 2) return B2ME_EUNSUPPORTED;
    if (SB > S) SB = S;
    p.debug = 0;
#if defined(B2ME_TC_PROFILE) || defined(B2ME_TC_EXPERIMENT)
    if (const char* e = getenv("B2ME_TC_DEBUG")) p.debug = atoi(e);
    if (const char* e = getenv("B2ME_TC_STAGES")) {  // debug build only: ring-depth experiments
        const int want = atoi(e);
        if (want >= 2 && want < S) S = want;
        if (SB > S) SB = S;
    }
#endif
    p.stages = S;
    p.stages_b = SB;
    // B2ME_TC_FLAG_TMA routes the operands through the TMA unit (tile::gather4 rows + 2-D weight boxes completing on
    // the leader's barrier, no relay, no proxy fence). Measured on the same box it runs at the speed of the default
    // cp.async gather + bulk copy + relay path (K27 384->384: 5.02 vs 5.04 ms), so it is an alternative the caller picks
    // per call; the parity tests run both.
    p.pf_mode = (flags & B2ME_TC_FLAG_PF_BULK) ? 1 : ((flags & B2ME_TC_FLAG_PF_NEAR) ? 0 : 2);
    p.tma = 0;
    if ((flags & B2ME_TC_FLAG_TMA) && V_in > 0) {
        const uint64_t w_rows = (uint64_t)K * (p.nchunk1 + p.nchunk2) * (uint64_t)Cout;
        bool ok = tc_make_map(&p.tm_in1, in1, es, (uint64_t)V_in, (uint64_t)Cin1, (uint32_t)cpc, 1,
                              CU_TENSOR_MAP_SWIZZLE_128B);
        if (ok && Cin2 > 0)
            ok = tc_make_map(&p.tm_in2, in2, es, (uint64_t)V_in, (uint64_t)Cin2, (uint32_t)cpc, 1,
                             CU_TENSOR_MAP_SWIZZLE_128B);
        if (ok)  // the packed image as rows of 64 two-byte words = 128 bytes, whatever the operand type
            ok = tc_make_map(&p.tm_w, packed_w, 2, w_rows, 64, 64, (uint32_t)(p.n_tile / 2), CU_TENSOR_MAP_SWIZZLE_NONE);
        if (!ok) return B2ME_ELAUNCH;
        p.tma = 1;
    }
    size_t smem = fixed + (size_t)S * TC_A_BYTES + (size_t)SB * p.b_bytes;
    if (smem < TC_MIN_SMEM) smem = TC_MIN_SMEM;

    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (head) return es == 2 ? tc_launch<1, 2, true>(p, smem, s) : tc_launch<1, 4, true>(p, smem, s);
    if (es == 2) {
        if (K == 27) return tc_launch<27, 2, false>(p, smem, s);
        if (K == 8) return tc_launch<8, 2, false>(p, smem, s);
        if (K == 1) return tc_launch<1, 2, false>(p, smem, s);
    } else {
        if (K == 27) return tc_launch<27, 4, false>(p, smem, s);
        if (K == 8) return tc_launch<8, 4, false>(p, smem, s);
        if (K == 1) return tc_launch<1, 4, false>(p, smem, s);
    }
    return B2ME_EUNSUPPORTED;
}

extern "C" int b2me_spconv_fwd_tc(const void* in1, int Cin1, const void* in2, int Cin2, int64_t V_in, int op_dtype,
                                  const void* packed_w,
                                  const int32_t* nbr, const int32_t* perm, const uint32_t* tile_masks, int K,
                                  int64_t V_out, int Cout,
                                  const float* scale, const float* shift, const void* residual, int act, float slope,
                                  void* out, int out_dtype, int flags, b2me_stream_t stream) {
    return tc_run(in1, Cin1, in2, Cin2, V_in, op_dtype, packed_w, nbr, perm, tile_masks, K, V_out, Cout, scale, shift,
                  residual, act, slope, out, out_dtype, nullptr, flags, stream);
}

extern "C" int b2me_head_fused_tc(const void* in, int Cin, int64_t V, int op_dtype, const void* packed_w1, int Chid,
                                  const float* scale1, const float* shift1, int act1, float slope1,
                                  const float* W2p, const float* b2, int C2, float* out_logits, uint8_t* out_argmax,
                                  int flags, b2me_stream_t stream) {
    if (!out_logits || !W2p) return B2ME_EINVAL;
    TcHead h{W2p, b2, C2, out_logits, out_argmax};
    return tc_run(in, Cin, nullptr, 0, V, op_dtype, packed_w1, nullptr, nullptr, nullptr, 1, V, Chid, scale1, shift1,
                  nullptr, act1, slope1, nullptr, B2ME_F32, &h, flags, stream);
}
