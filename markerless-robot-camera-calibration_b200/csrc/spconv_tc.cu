// spconv_tc.cu — K4b: sparse convolution as an output-stationary implicit GEMM on tcgen05 (sm_100a).
//
//   CTA tile      128 output voxels x n_tile output channels (n_tile <= 384 fp32 columns of TMEM)
//   reduction     items = (kernel offset k that has at least one neighbour in the tile) x (64-channel chunk)
//   A operand     128 gathered input rows x 64 bf16 (= one 128-byte swizzle row per voxel), cp.async 16-byte
//                 pieces straight into the SWIZZLE_128B K-major smem image, zero-fill for missing neighbours
//   B operand     W[k][chunk] pre-packed on the host side of the ABI into the exact smem image, so one
//                 cp.async.bulk (UBLKCP) per item brings n_tile x 128 bytes and completes on the stage mbarrier
//   MMA           one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N<=256, K=16), fp32
//                 accumulators stay in TMEM for the whole tile
//   epilogue      tcgen05.ld -> folded BatchNorm scale/shift, residual add, ReLU/LeakyReLU -> bf16/f32 rows
//   warps         0-3 gather producers, then epilogue (warp w owns TMEM lanes 32w..32w+31)
//                 4   TMEM alloc + MMA issuer        5   weight (B) bulk-copy issuer
//   pipeline      S-stage ring: full[s] (4 producer warps + 1 expect_tx arrive), empty[s] (tcgen05.commit)
//
// The two sources (in1 | in2) implement ME.cat without materialising the concatenation.
#include "common.cuh"
#include <stdlib.h>

#define TC_BM 128
#define TC_BK 64
#define TC_A_BYTES (TC_BM * 128)
#define TC_THREADS 192
#define TC_LAG 2
#define TC_MAX_SMEM 232448

struct TcParams {
    const __nv_bfloat16* in1;
    const __nv_bfloat16* in2;
    const uint8_t* wpacked;
    const int32_t* nbr;
    const int32_t* perm;
    const float* scale;
    const float* shift;
    const __nv_bfloat16* residual;
    void* out;
    long long V_out;
    int Cin1, Cin2, nchunk1, nchunk2;
    int K, Cout, n_tile, stages;
    int act, out_dtype, tmem_cols;
    float slope;
    unsigned int b_bytes;
    int debug;  // B2ME_TC_DEBUG (timing experiments only, results invalid): 1 skip gather, 2 skip weights, 4 skip MMA
};

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
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

// ------------------------------------------------------------------------------------------ kernel
__global__ void __launch_bounds__(TC_THREADS) k_spconv_tc(const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);

    const int S = p.stages;
    const uint32_t stage_bytes = TC_A_BYTES + p.b_bytes;
    const int K = p.K;
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    // carve: [stages][nbr_s 128*K i32][rows_s 128 i32][scale n_tile][shift n_tile][full S][empty S][tmem_full][tmem_ptr][mask]
    uint32_t off = (uint32_t)S * stage_bytes;
    int32_t* nbr_s = reinterpret_cast<int32_t*>(sm + off);
    off += TC_BM * K * 4;
    int32_t* rows_s = reinterpret_cast<int32_t*>(sm + off);  // output row of every tile slot (-1 = past the end)
    off += TC_BM * 4;
    float* scale_s = reinterpret_cast<float*>(sm + off);
    off += p.n_tile * 4;
    float* shift_s = reinterpret_cast<float*>(sm + off);
    off += p.n_tile * 4;
    off = (off + 7u) & ~7u;
    const uint32_t bar_full = base + off;
    off += 8 * S;
    const uint32_t bar_empty = base + off;
    off += 8 * S;
    const uint32_t bar_tmem = base + off;
    off += 8;
    uint32_t* tmem_ptr_s = reinterpret_cast<uint32_t*>(sm + off);
    off += 4;
    uint32_t* mask_s = reinterpret_cast<uint32_t*>(sm + off);

    const long long row0 = (long long)blockIdx.x * TC_BM;
    const int n0 = blockIdx.y * p.n_tile;

    // ---- one-time setup
    if (tid == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(bar_full + 8 * s, 5);   // 4 producer warps + 1 expect_tx arrive of the B loader
            mbar_init(bar_empty + 8 * s, 1);  // tcgen05.commit
        }
        mbar_init(bar_tmem, 1);
        *mask_s = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_s)),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < TC_BM) {
        const long long slot = row0 + tid;
        rows_s[tid] = (slot < p.V_out) ? (p.perm ? __ldg(p.perm + slot) : (int)slot) : -1;
    }
    __syncthreads();  // mask_s = 0 and rows_s visible
    {
        // stage the kernel-map rows of the tile's output rows and find the offsets that occur
        uint32_t local = 0;
        if (p.nbr) {
            const int total = TC_BM * K;
            for (int i = tid; i < total; i += TC_THREADS) {
                const int r = i / K;
                const int row = rows_s[r];
                const int v = (row >= 0) ? __ldg(p.nbr + (long long)row * K + (i - r * K)) : -1;
                nbr_s[i] = v;
                if (v >= 0) local |= 1u << (i - r * K);
            }
        } else {
            local = 1u;
        }
        local = __reduce_or_sync(0xffffffffu, local);
        if (lane == 0 && local) atomicOr(mask_s, local);
        for (int i = tid; i < p.n_tile; i += TC_THREADS) {
            scale_s[i] = p.scale ? p.scale[n0 + i] : 1.f;
            shift_s[i] = p.shift ? p.shift[n0 + i] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_s;
    const uint32_t kmask = *mask_s;
    const int nchunk = p.nchunk1 + p.nchunk2;

    if (warp < 4) {
        // =============================== gather producers ===============================
        const int j = tid & 7;        // 16-byte piece inside the 128-byte row
        const int rbase = tid >> 3;   // rows rbase + 16*i
        int issued = 0, arrived = 0;
        for (int k = 0; k < K; ++k) {
            if (!((kmask >> k) & 1u)) continue;
            int idx[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rbase + 16 * i;
                if (p.nbr) idx[i] = nbr_s[r * K + k];
                else idx[i] = rows_s[r];
            }
            for (int c = 0; c < nchunk; ++c) {
                const int s = issued % S;
                if (issued >= S) {
                    // the slot is free once the MMAs of item (issued - S) have completed. Before blocking on that,
                    // publish every gather already issued: the MMA warp must never wait for data whose arrival is
                    // only signalled after a LATER gather could be issued (that would serialise MMA and gather).
                    const uint32_t par = ((issued / S) & 1) ^ 1;
                    const bool ready = __all_sync(0xffffffffu, mbar_try_wait(bar_empty + 8 * s, par));
                    if (!ready) {
                        if (arrived < issued) {
                            cp_async_wait<0>();
                            fence_proxy_async();
                            __syncwarp();
                            for (; arrived < issued; ++arrived)
                                if (lane == 0) mbar_arrive(bar_full + 8 * (arrived % S));
                        }
                        mbar_wait(bar_empty + 8 * s, par);
                    }
                }
                const __nv_bfloat16* src;
                int cin, coff;
                if (c < p.nchunk1) { src = p.in1; cin = p.Cin1; coff = c * TC_BK; }
                else { src = p.in2; cin = p.Cin2; coff = (c - p.nchunk1) * TC_BK; }
                const int kw = min(TC_BK, cin - coff);
                if (j * 8 < kw && !(p.debug & 1)) {
                    const uint32_t a_s = base + (uint32_t)s * stage_bytes;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int r = rbase + 16 * i;
                        const uint32_t dst = a_s + (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
                        const int id = idx[i];
                        const __nv_bfloat16* g = src + (id >= 0 ? ((long long)id * cin + coff + j * 8) : 0);
                        cp_async_16(dst, g, id >= 0 ? 16u : 0u);
                    }
                }
                cp_async_commit();
                ++issued;
                if (issued - arrived > TC_LAG) {
                    cp_async_wait<TC_LAG>();
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(bar_full + 8 * (arrived % S));
                    ++arrived;
                }
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        __syncwarp();
        for (; arrived < issued; ++arrived)
            if (lane == 0) mbar_arrive(bar_full + 8 * (arrived % S));

        // =============================== epilogue ===============================
        mbar_wait(bar_tmem, 0);
        tc_fence_after();
        const long long row = rows_s[tid];
        const bool row_ok = row >= 0;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        for (int cb = 0; cb < p.n_tile; cb += 16) {
            uint32_t r[16];
            if (kmask) {
                tmem_ld_x16(lane_addr + (uint32_t)cb, r);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int q = 0; q < 16; ++q) r[q] = 0u;
            }
            if (!row_ok) continue;
            float v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = __uint_as_float(r[q]) * scale_s[cb + q] + shift_s[cb + q];
            const long long o = row * p.Cout + n0 + cb;
            if (p.residual) {
                const uint4* rp = reinterpret_cast<const uint4*>(p.residual + o);
                const uint4 ra = __ldg(rp), rb = __ldg(rp + 1);
                const uint32_t w[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    v[2 * q] += __uint_as_float(w[q] << 16);
                    v[2 * q + 1] += __uint_as_float(w[q] & 0xFFFF0000u);
                }
            }
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = apply_act(v[q], p.act, p.slope);
            if (p.out_dtype == B2ME_BF16) {
                uint32_t w[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * q], v[2 * q + 1]);
                    w[q] = *reinterpret_cast<uint32_t*>(&h);
                }
                uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + o);
                op[0] = make_uint4(w[0], w[1], w[2], w[3]);
                op[1] = make_uint4(w[4], w[5], w[6], w[7]);
            } else {
                float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + o);
#pragma unroll
                for (int q = 0; q < 4; ++q) op[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
        }
    } else if (warp == 4) {
        // =============================== MMA issuer ===============================
        if (lane == 0) {
            const int nhalf = p.n_tile > 256 ? 2 : 1;
            const int nh = p.n_tile / nhalf;
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(nh >> 3) << 17) | (8u << 24);
            int it = 0;
            for (int k = 0; k < K; ++k) {
                if (!((kmask >> k) & 1u)) continue;
                for (int c = 0; c < nchunk; ++c) {
                    const int s = it % S;
                    const int kw = (c < p.nchunk1) ? min(TC_BK, p.Cin1 - c * TC_BK)
                                                   : min(TC_BK, p.Cin2 - (c - p.nchunk1) * TC_BK);
                    mbar_wait(bar_full + 8 * s, (it / S) & 1);
                    tc_fence_after();
                    const uint32_t a_s = base + (uint32_t)s * stage_bytes;
                    const uint32_t b_s = a_s + TC_A_BYTES;
                    const uint64_t adesc = make_smem_desc_sw128(a_s);
                    for (int kk = 0; kk < ((p.debug & 4) ? 0 : kw / 16); ++kk) {
                        for (int h = 0; h < nhalf; ++h) {
                            const uint64_t bdesc = make_smem_desc_sw128(b_s + (uint32_t)(h * nh) * 128u);
                            tc_mma_bf16(tmem_base + (uint32_t)(h * nh), adesc + (uint64_t)(kk * 2),
                                        bdesc + (uint64_t)(kk * 2), idesc, (it > 0 || kk > 0) ? 1u : 0u);
                        }
                    }
                    tc_commit(bar_empty + 8 * s);
                    ++it;
                }
            }
            tc_commit(bar_tmem);
        }
    } else {
        // =============================== weight (B) loader ===============================
        if (lane == 0) {
            int it = 0;
            for (int k = 0; k < K; ++k) {
                if (!((kmask >> k) & 1u)) continue;
                for (int c = 0; c < nchunk; ++c) {
                    const int s = it % S;
                    if (it >= S) mbar_wait(bar_empty + 8 * s, ((it / S) & 1) ^ 1);
                    const uint32_t b_s = base + (uint32_t)s * stage_bytes + TC_A_BYTES;
                    const uint8_t* g =
                        p.wpacked + ((size_t)((size_t)blockIdx.y * K + k) * nchunk + c) * (size_t)p.b_bytes;
                    if (p.debug & 2) {
                        mbar_arrive(bar_full + 8 * s);
                    } else {
                        mbar_arrive_expect_tx(bar_full + 8 * s, p.b_bytes);
                        bulk_copy_g2s(b_s, g, p.b_bytes, bar_full + 8 * s);
                    }
                    ++it;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)p.tmem_cols)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------ host side
static int tc_n_tile(int Cout) {
    if (Cout <= 384) return Cout;
    return 256;
}

extern "C" int b2me_tc_supported(int K, int Cin1, int Cin2, int Cout) {
    if (K < 1 || K > 32 || Cin1 < 16 || Cin2 < 0 || Cout < 16) return 0;
    if (Cin1 % 16 || Cin2 % 16 || Cout % 16) return 0;
    const int nt = tc_n_tile(Cout);
    if (Cout % nt) return 0;
    if (nt > 256 && (nt % 32)) return 0;
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
    const long long item = ((long long)nt * K + k) * nchunk + c;
    const long long piece = item * ((long long)n_tile * 8) + (long long)n * 8 + (j ^ (n & 7));
    packed[piece] = make_uint4(w[0], w[1], w[2], w[3]);
}

extern "C" int b2me_tc_pack_weights(const float* W, int K, int Cin1, int Cin2, int Cout, void* packed,
                                    b2me_stream_t stream) {
    if (!W || !packed) return B2ME_EINVAL;
    if (!b2me_tc_supported(K, Cin1, Cin2, Cout)) return B2ME_EUNSUPPORTED;
    const int nt = tc_n_tile(Cout);
    const int nchunk1 = (Cin1 + TC_BK - 1) / TC_BK, nchunk2 = (Cin2 + TC_BK - 1) / TC_BK;
    const long long total = (long long)(Cout / nt) * K * (nchunk1 + nchunk2) * nt * 8;
    k_tc_pack<<<(unsigned)ceil_div64(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        W, K, Cin1, Cin2, Cout, nt, nchunk1, nchunk2, reinterpret_cast<uint4*>(packed), total);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

extern "C" int b2me_spconv_fwd_tc(const void* in1, int Cin1, const void* in2, int Cin2, const void* packed_w,
                                  const int32_t* nbr, const int32_t* perm, int K, int64_t V_out, int Cout,
                                  const float* scale, const float* shift, const void* residual, int act, float slope,
                                  void* out, int out_dtype, b2me_stream_t stream) {
    if (!in1 || !packed_w || !out || V_out < 0) return B2ME_EINVAL;
    if (Cin2 > 0 && !in2) return B2ME_EINVAL;
    if (!nbr && K != 1) return B2ME_EINVAL;
    if (!b2me_tc_supported(K, Cin1, Cin2, Cout)) return B2ME_EUNSUPPORTED;
    if (out_dtype != B2ME_BF16 && out_dtype != B2ME_F32) return B2ME_EINVAL;
    if (V_out == 0) return B2ME_OK;

    TcParams p;
    p.in1 = reinterpret_cast<const __nv_bfloat16*>(in1);
    p.in2 = reinterpret_cast<const __nv_bfloat16*>(in2);
    p.wpacked = reinterpret_cast<const uint8_t*>(packed_w);
    p.nbr = nbr;
    p.perm = perm;
    p.scale = scale;
    p.shift = shift;
    p.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
    p.out = out;
    p.V_out = V_out;
    p.Cin1 = Cin1;
    p.Cin2 = Cin2;
    p.nchunk1 = (Cin1 + TC_BK - 1) / TC_BK;
    p.nchunk2 = (Cin2 + TC_BK - 1) / TC_BK;
    p.K = K;
    p.Cout = Cout;
    p.n_tile = tc_n_tile(Cout);
    p.act = act;
    p.out_dtype = out_dtype;
    p.slope = slope;
    p.b_bytes = (unsigned)p.n_tile * 128u;
    {
        const char* dbg = getenv("B2ME_TC_DEBUG");
        p.debug = dbg ? atoi(dbg) : 0;
    }
    int cols = 32;
    while (cols < p.n_tile) cols <<= 1;
    p.tmem_cols = cols;

    const size_t fixed = 1024 /*align slack*/ + (size_t)TC_BM * (K + 1) * 4 + (size_t)p.n_tile * 8 + 8 + 16 * 8 + 64;
    const size_t stage_bytes = TC_A_BYTES + p.b_bytes;
    int S = 4;
    while (S >= 3 && fixed + (size_t)S * stage_bytes > TC_MAX_SMEM) --S;
    if (S < 3) return B2ME_EUNSUPPORTED;
    p.stages = S;
    const size_t smem = fixed + (size_t)S * stage_bytes;

    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(k_spconv_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_MAX_SMEM) != cudaSuccess)
            return B2ME_ELAUNCH;
        attr_set = true;
    }
    dim3 grid((unsigned)ceil_div64(V_out, TC_BM), (unsigned)(Cout / p.n_tile));
    k_spconv_tc<<<grid, TC_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}
