// mma_probe.cu — developer microbenchmark: raw issue throughput of tcgen05.mma (kind::f16, bf16 -> f32) from static
// shared-memory operands, cta_group::1 (M=128) and cta_group::2 (M=256), for several N. No loads, no epilogue.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu && tools/mma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
template <int CG>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// warp-uniform issue: every lane executes the statement with the same operands, elect.sync picks the issuing lane
template <int CG>
__device__ __forceinline__ void mma_elect(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (CG == 1)
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                     ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit(uint32_t bar) {
    if (CG == 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// per round: KSTEPS k-steps x (one MMA of N=n_a and, if n_b, one of N=n_b); commit every `per_commit` rounds
// wdelay >= 0: warps 1-3 of every CTA stream st.shared.v4 (512 bytes per warp instruction) into a scratch region while
// the MMAs run, with `wdelay` dependent ALU steps between stores: MMA time per item against competing shared-memory
// write traffic (out[2] = bytes written by CTA 0, out[3] = cycles the writers ran)
template <int CG>
__global__ void __launch_bounds__(128, 1) k_probe(int n_a, int n_b, int ksteps, int rounds, int distinct_stages,
                                                  int commit_each, unsigned long long* out, int wdelay = -1,
                                                  const uint8_t* gsrc = nullptr, int amode = 0) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint64_t bar2[8];
    __shared__ uint64_t bar3;
    __shared__ uint32_t tmem_ptr;
    __shared__ volatile int stop_flag;
    if (threadIdx.x == 0) stop_flag = 0;
    uint32_t rank = 0;
    if (CG == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // zero the operands (values do not matter, NaN patterns could)
    for (uint32_t i = threadIdx.x; i < 200u * 1024u / 16u; i += blockDim.x)
        reinterpret_cast<uint4*>(smem_raw + (base - raw))[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar), 1);
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bar2[i]), 1);
        mbar_init(smem_u32(&bar3), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        if (CG == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(512u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_ptr;
    const int M = CG == 2 ? 256 : 128;
    const uint32_t idesc_a = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_a >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const uint32_t idesc_b = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_b >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const int rows_a = CG == 2 ? n_a / 2 : n_a;  // B rows of instruction a held by this CTA
    const uint32_t stage_bytes = 16384u + (uint32_t)(CG == 2 ? (n_a + n_b) / 2 : (n_a + n_b)) * 128u;
    if ((commit_each & 4) && warp == 0 && rank == 0) {
        // whole warp, uniform operands, elect inside the asm
        const long long t0 = clock64();
        uint32_t acc = 0;
        for (int r = 0; r < rounds; ++r) {
            const uint32_t a_s = base + (uint32_t)(r % distinct_stages) * stage_bytes;
            const uint32_t b_s = a_s + 16384u;
            const uint64_t ad = desc_sw128(a_s), bda = desc_sw128(b_s), bdb = desc_sw128(b_s + (uint32_t)rows_a * 128u);
            for (int kk = 0; kk < ksteps; ++kk) {
                mma_elect<CG>(tmem, ad + (uint64_t)(kk * 2), bda + (uint64_t)(kk * 2), idesc_a, acc);
                if (n_b) mma_elect<CG>(tmem + (uint32_t)n_a, ad + (uint64_t)(kk * 2), bdb + (uint64_t)(kk * 2), idesc_b, acc);
                acc = 1;
            }
        }
        const long long t1 = clock64();
        if (lane == 0) {
            commit<CG>(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
        }
        __syncwarp();
        const long long t2 = clock64();
        if (blockIdx.x == 0 && lane == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t2 - t0); }
    } else if (!(commit_each & 4) && warp == 0 && lane == 0 && rank == 0) {
        const long long t0 = clock64();
        uint32_t acc = 0;
        for (int r = 0; r < rounds; ++r) {
            const uint32_t a_s = base + (uint32_t)(r % distinct_stages) * stage_bytes;
            const uint32_t b_s = a_s + 16384u;
            const uint64_t ad = desc_sw128(a_s), bda = desc_sw128(b_s), bdb = desc_sw128(b_s + (uint32_t)rows_a * 128u);
            for (int kk = 0; kk < ksteps; ++kk) {
                mma<CG>(tmem, ad + (uint64_t)(kk * 2), bda + (uint64_t)(kk * 2), idesc_a, acc);
                if (n_b) mma<CG>(tmem + (uint32_t)n_a, ad + (uint64_t)(kk * 2), bdb + (uint64_t)(kk * 2), idesc_b, acc);
                acc = 1;
            }
            if (commit_each & 2) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (commit_each & 1) commit<CG>(smem_u32(&bar2[r % distinct_stages]));
            if (commit_each & 8) {   // a plain shared-memory poll between items
                uint32_t v;
                asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(&tmem_ptr)) : "memory");
                if (v == 0xdeadbeefu) acc = 0;
            }
            if (commit_each & 16) {  // an mbarrier check on a phase that completed long ago (parity 1 of a fresh barrier)
                uint32_t ok;
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(smem_u32(&bar3)), "r"(1u) : "memory");
                if (!ok) acc = 0;
            }
            if (commit_each & 32) asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        const long long t1 = clock64();
        commit<CG>(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = (unsigned long long)(t1 - t0); out[1] = (unsigned long long)(t2 - t0); }
    }
    if (warp == 0 && lane == 0 && rank == 0 && (wdelay >= 0 || amode)) {
        stop_flag = 1;
        if (CG == 2) {  // stop the peer's writers too
            uint32_t ra;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32((const void*)&stop_flag)), "r"(1u));
            asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(1u) : "memory");
        }
    }
    // amode bit 1: warp 1 streams cp.async.bulk copies (24 KB each, two in flight, L2-resident source) into scratch
    // shared memory; bit 2: warps 2-3 stream 16-byte cp.async gathers (8 per thread per round). wdelay = ALU steps
    // between rounds. These are the async-proxy / LDGSTS write paths the convolution's operand loads use.
    __shared__ uint64_t wbar[2];
    if (amode && threadIdx.x == 32) {
        mbar_init(smem_u32(&wbar[0]), 1);
        mbar_init(smem_u32(&wbar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (amode && warp == 1 && (amode & 1)) {
        unsigned long long bytes = 0;
        uint32_t x = threadIdx.x, i = 0;
        const long long t0 = clock64();
        __syncwarp();
        while (!stop_flag) {
            const uint32_t b = i & 1u;
            if (i >= 2) mbar_wait(smem_u32(&wbar[b]), ((i >> 1) - 1) & 1u);
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&wbar[b])), "r"(24576u) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(base + 120u * 1024u + b * 24576u), "l"(gsrc + (size_t)((i * 7u + blockIdx.x) & 31u) * 24576u),
                               "r"(24576u), "r"(smem_u32(&wbar[b])) : "memory");
            }
            __syncwarp();
            ++i;
            bytes += 24576;
            for (int d = 0; d < wdelay; ++d) x = x * 1664525u + 1013904223u;
        }
        // drain
        if (i >= 1) mbar_wait(smem_u32(&wbar[(i - 1) & 1u]), (((i - 1) >> 1)) & 1u);
        if (i >= 2) mbar_wait(smem_u32(&wbar[(i - 2) & 1u]), (((i - 2) >> 1)) & 1u);
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) {
            atomicAdd(&out[2], bytes + (x == 0xdeadbeefu));
            out[3] = (unsigned long long)(t1 - t0);
        }
    } else if (amode && warp >= 2 && (amode & 2)) {
        const uint32_t scratch = base + 168u * 1024u + (uint32_t)(warp - 1) * 8192u;
        unsigned long long bytes = 0;
        uint32_t x = threadIdx.x * 2654435761u, i = 0;
        const long long t0 = clock64();
        while (!stop_flag) {
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                // 8 lanes fetch one 128-byte row from a pseudo-random place of a 32 MB window (L2-resident)
                const uint32_t row = (x >> 3) + (uint32_t)u * 977u + (uint32_t)(lane >> 3) * 131u;
                const uint8_t* g = gsrc + ((size_t)(row & 0x3FFFFu) * 128u) + (uint32_t)(lane & 7) * 16u;
                const uint32_t dst = scratch + (uint32_t)u * 512u + (uint32_t)lane * 16u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(g) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 2;" ::: "memory");
            ++i;
            bytes += 8 * 512;
            x = x * 1664525u + 1013904223u;
            x = __shfl_sync(0xffffffffu, x, 0);
            for (int d = 0; d < wdelay; ++d) x = x * 1664525u + 1013904223u;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) {
            atomicAdd(&out[4], bytes + (x == 0xdeadbeefu));
            if (warp == 2) out[5] = (unsigned long long)(t1 - t0);
        }
    }
    if (!amode && warp >= 1 && wdelay >= 0) {
        const uint32_t scratch = base + 168u * 1024u + (uint32_t)(warp - 1) * 8192u;
        unsigned long long bytes = 0;
        uint32_t x = threadIdx.x, i = 0;
        const long long t0 = clock64();
        while (!stop_flag) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {  // 16 independent 512-byte stores per flag check
                const uint32_t dst = scratch + (uint32_t)u * 512u + (uint32_t)lane * 16u;
                asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst), "r"(x) : "memory");
            }
            ++i;
            bytes += 16 * 512;
            for (int d = 0; d < wdelay; ++d) x = x * 1664525u + 1013904223u;
        }
        const long long t1 = clock64();
        if (blockIdx.x == 0 && lane == 0) {
            atomicAdd(&out[2], bytes + (x == 0xdeadbeefu));
            if (warp == 1) out[3] = (unsigned long long)(t1 - t0);
        }
    }
    if (CG == 2 && rank == 1 && threadIdx.x == 0) mbar_wait(smem_u32(&bar), 0);  // multicast commit reaches the peer too
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncwarp();
    if (CG == 2) {
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    if (warp == 0) {
        if (CG == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <int CG>
static void run(int n_a, int n_b, int grid, int stages, int commit_each, int wdelay = -1, int amode = 0) {
    unsigned long long* d;
    cudaMalloc(&d, 64);
    cudaMemset(d, 0, 64);
    static uint8_t* gsrc = nullptr;
    if (!gsrc) { cudaMalloc(&gsrc, 34u << 20); cudaMemset(gsrc, 0, 34u << 20); }
    const int ksteps = 4, rounds = 2000;
    const size_t smem = 201 * 1024 + 1024;
    cudaFuncSetAttribute(k_probe<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_probe<CG>, n_a, n_b, ksteps, rounds, stages, commit_each, d, wdelay,
                                       (const uint8_t*)gsrc, amode);
    cudaError_t e2 = cudaDeviceSynchronize();
    unsigned long long h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    if (amode)
        printf("  [async writers mode %d delay %d: bulk %.1f B/cyc, cp.async %.1f B/cyc per SM] ", amode, wdelay,
               h[3] ? (double)h[2] / (double)h[3] : 0.0, h[5] ? (double)h[4] / (double)h[5] : 0.0);
    else if (wdelay >= 0)
        printf("  [writers: delay %d, %.1f B/cycle of st.shared per SM] ", wdelay, h[3] ? (double)h[2] / (double)h[3] : 0.0);
    const double n_mma = (double)rounds * ksteps * (n_b ? 2 : 1);
    const double flop_per_sm = (double)rounds * ksteps * 2.0 * 128 * (n_a + n_b) * 16;
    printf("cta_group::%d grid %3d N=%3d+%3d stages %d commit/fence %d: issue %.1f cyc/MMA, complete %.1f cyc/MMA, %.0f cyc per 64-K item, "
           "%.0f FLOP/cyc/SM  (%s %s)\n", CG, grid, n_a, n_b, stages, commit_each, h[0] / n_mma, h[1] / n_mma, h[1] / (double)rounds,
           flop_per_sm / h[1], cudaGetErrorString(e), cudaGetErrorString(e2));
    cudaFree(d);
}

int main(int argc, char** argv) {
    if (argc > 1 && argv[1][0] == 'a') {
        // async-proxy contention sweep (the convolution's operand load paths) against the 256 + 128 instruction pair
        run<2>(256, 128, 148, 3, 1);
        for (int am : {1, 2, 3})
            for (int wd : {2048, 512, 128, 0}) run<2>(256, 128, 148, 3, 1, wd, am);
        run<2>(256, 128, 2, 3, 1);
        for (int wd : {512, 0}) run<2>(256, 128, 2, 3, 1, wd, 3);
        return 0;
    }
    if (argc > 1 && argv[1][0] == 's') {
        // shared-memory contention sweep: the 256 + 128 instruction pair on every SM against st.shared write streams
        for (int grid : {2, 148})
            for (int wd : {-1, 1024, 512, 256, 128, 64, 32, 16, 0}) run<2>(256, 128, grid, 3, 1, wd);
        for (int wd : {-1, 256, 64, 0}) run<2>(128, 0, 148, 3, 1, wd);
        for (int wd : {-1, 256, 64, 0}) run<2>(256, 0, 148, 3, 1, wd);
        return 0;
    }
    // bits: 1 commit per item, 2 fence.proxy.async, 8 shared-memory poll, 16 mbarrier test_wait, 32 tcgen05.fence::after
    for (int ce : {0, 1, 3, 1 + 8, 1 + 16, 1 + 32, 1 + 2 + 16 + 32}) run<2>(256, 128, 2, 3, ce);
    return 0;
}
