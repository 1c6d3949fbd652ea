// spconv_simt.cu — K4a: fp32-accumulate SIMT gather-GEMM (output-stationary, no scatter/atomics),
// plus the small element-wise / gather / small-N linear kernels of the ME-compatible layer.
//
// The SIMT convolution is the exact-fp32 path (parity tolerance 1e-3, SURVEY.md north star) and the
// path for shapes the tcgen05 kernel does not take (Cin = 3 stem, odd channel counts).
#include "common.cuh"

#define SM_BM 64
#define SM_BN 64
#define SM_BK 16
#define SM_THREADS 256

// out[o,:] = act((sum_kk A[o,kk] * W[kk,:]) * scale + shift + residual), kk = k*Cin + c flattened
__global__ void __launch_bounds__(SM_THREADS)
k_spconv_simt(const void* __restrict__ in1, int Cin1, const void* __restrict__ in2, int Cin2, int in_dtype,
              const float* __restrict__ W, const int32_t* __restrict__ nbr, int K, int64_t V_out, int Cout,
              const float* __restrict__ scale, const float* __restrict__ shift, const void* __restrict__ residual,
              int res_dtype, int act, float slope, void* __restrict__ out, int out_dtype) {
    __shared__ float As[SM_BK][SM_BM + 4];
    __shared__ float Bs[SM_BK][SM_BN + 4];
    const int Cin = Cin1 + Cin2;
    const int KK = K * Cin;
    const int64_t row0 = (int64_t)blockIdx.x * SM_BM;
    const int n0 = blockIdx.y * SM_BN;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int a_r = tid >> 2;         // 0..63
    const int a_k = (tid & 3) * 4;    // 0,4,8,12
    const int b_k = tid >> 4;         // 0..15
    const int b_n = (tid & 15) * 4;   // 0..60
    const int64_t a_row = row0 + a_r;

    for (int kk0 = 0; kk0 < KK; kk0 += SM_BK) {
        // ---- gather A
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = kk0 + a_k + j;
            float v = 0.f;
            if (kk < KK && a_row < V_out) {
                const int k = kk / Cin;
                const int c = kk - k * Cin;
                const int64_t idx = nbr ? (int64_t)nbr[a_row * K + k] : a_row;
                if (idx >= 0) {
                    v = (c < Cin1) ? load_as_f32(in1, in_dtype, idx * Cin1 + c)
                                   : load_as_f32(in2, in_dtype, idx * Cin2 + (c - Cin1));
                }
            }
            As[a_k + j][a_r] = v;
        }
        // ---- load B
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int kk = kk0 + b_k;
            const int n = n0 + b_n + j;
            Bs[b_k][b_n + j] = (kk < KK && n < Cout) ? __ldg(W + (int64_t)kk * Cout + n) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < SM_BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t row = row0 + ty * 4 + i;
        if (row >= V_out) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= Cout) continue;
            float v = acc[i][j];
            if (scale) v *= scale[n];
            if (shift) v += shift[n];
            if (residual) v += load_as_f32(residual, res_dtype, row * Cout + n);
            v = apply_act(v, act, slope);
            store_from_f32(out, out_dtype, row * Cout + n, v);
        }
    }
}

// Stem variant (conv0p1s1: Cin = 3, Cout = 32, K = 27, fp32 point features): one thread per output row keeps the 32
// accumulators in registers, the 27 x Cin x 32 weights sit in shared memory (broadcast reads), absent neighbours are
// skipped. Same summation order (k, then c) as the generic kernel: bit-identical results.
#define STEM_COUT 32
__global__ void __launch_bounds__(256)
k_spconv_stem(const float* __restrict__ in, int Cin, const float* __restrict__ W, const int32_t* __restrict__ nbr, int K,
              int64_t V_out, const float* __restrict__ scale, const float* __restrict__ shift, int act, float slope,
              void* __restrict__ out, int out_dtype) {
    extern __shared__ __align__(16) float stem_w[];  // [K * Cin][32]
    for (int i = threadIdx.x; i < K * Cin * STEM_COUT; i += blockDim.x) stem_w[i] = W[i];
    __syncthreads();
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < V_out;
         row += (int64_t)gridDim.x * blockDim.x) {
        float acc[STEM_COUT];
#pragma unroll
        for (int n = 0; n < STEM_COUT; ++n) acc[n] = 0.f;
        for (int k = 0; k < K; ++k) {
            const int64_t id = nbr ? (int64_t)__ldg(nbr + row * K + k) : row;
            if (id < 0) continue;
            for (int c = 0; c < Cin; ++c) {
                const float x = __ldg(in + id * Cin + c);
                const float4* w = reinterpret_cast<const float4*>(stem_w + (k * Cin + c) * STEM_COUT);
#pragma unroll
                for (int q = 0; q < STEM_COUT / 4; ++q) {
                    const float4 wv = w[q];
                    acc[4 * q + 0] = fmaf(x, wv.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(x, wv.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(x, wv.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(x, wv.w, acc[4 * q + 3]);
                }
            }
        }
#pragma unroll
        for (int n = 0; n < STEM_COUT; ++n) {
            float v = acc[n];
            if (scale) v *= scale[n];
            if (shift) v += shift[n];
            acc[n] = apply_act(v, act, slope);
        }
        if (out_dtype == B2ME_BF16) {
            uint4* op = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(out) + row * STEM_COUT);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint32_t wv[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(acc[q * 8 + 2 * e], acc[q * 8 + 2 * e + 1]);
                    wv[e] = *reinterpret_cast<uint32_t*>(&h);
                }
                op[q] = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
        } else {
            if (out_dtype == B2ME_TF32) {
#pragma unroll
                for (int n = 0; n < STEM_COUT; ++n) acc[n] = round_tf32(acc[n]);
            }
            float4* op = reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + row * STEM_COUT);
#pragma unroll
            for (int q = 0; q < 8; ++q) op[q] = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        }
    }
}

extern "C" int b2me_spconv_fwd_simt(const void* in1, int Cin1, const void* in2, int Cin2, int in_dtype,
                                    const float* W, const int32_t* nbr, int K, int64_t V_out, int Cout,
                                    const float* scale, const float* shift, const void* residual, int res_dtype,
                                    int act, float slope, void* out, int out_dtype, b2me_stream_t stream) {
    if (!in1 || !W || !out || Cin1 <= 0 || Cin2 < 0 || K <= 0 || Cout <= 0 || V_out < 0) return B2ME_EINVAL;
    if (Cin2 > 0 && !in2) return B2ME_EINVAL;
    if (!nbr && K != 1) return B2ME_EINVAL;
    if (V_out == 0) return B2ME_OK;
    if (Cin2 == 0 && Cin1 <= 4 && Cout == STEM_COUT && K <= 27 && (in_dtype == B2ME_F32 || in_dtype == B2ME_TF32) && !residual) {
        int64_t blocks = ceil_div64(V_out, 256);
        if (blocks > B2ME_NUM_SMS * 16) blocks = B2ME_NUM_SMS * 16;
        const size_t smem = (size_t)K * Cin1 * STEM_COUT * sizeof(float);
        k_spconv_stem<<<(unsigned)blocks, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
            reinterpret_cast<const float*>(in1), Cin1, W, nbr, K, V_out, scale, shift, act, slope, out, out_dtype);
        B2ME_CHECK_LAUNCH();
        return B2ME_OK;
    }
    dim3 grid((unsigned)ceil_div64(V_out, SM_BM), (unsigned)((Cout + SM_BN - 1) / SM_BN));
    k_spconv_simt<<<grid, SM_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        in1, Cin1, in2, Cin2, in_dtype, W, nbr, K, V_out, Cout, scale, shift, residual, res_dtype, act, slope, out,
        out_dtype);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ elementwise
__global__ void k_affine_act(const void* __restrict__ in, int in_dtype, int64_t total, int C,
                             const float* __restrict__ scale, const float* __restrict__ shift,
                             const void* __restrict__ residual, int res_dtype, int act, float slope,
                             void* __restrict__ out, int out_dtype) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        float v = load_as_f32(in, in_dtype, t);
        if (scale) v *= scale[c];
        if (shift) v += shift[c];
        if (residual) v += load_as_f32(residual, res_dtype, t);
        store_from_f32(out, out_dtype, t, apply_act(v, act, slope));
    }
}

extern "C" int b2me_affine_act(const void* in, int in_dtype, int64_t V, int C, const float* scale,
                               const float* shift, const void* residual, int res_dtype, int act, float slope,
                               void* out, int out_dtype, b2me_stream_t stream) {
    if (!in || !out || V < 0 || C <= 0) return B2ME_EINVAL;
    const int64_t total = V * C;
    if (total == 0) return B2ME_OK;
    int64_t blocks = ceil_div64(total, 256);
    if (blocks > B2ME_NUM_SMS * 16) blocks = B2ME_NUM_SMS * 16;
    k_affine_act<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        in, in_dtype, total, C, scale, shift, residual, res_dtype, act, slope, out, out_dtype);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

__global__ void k_convert(const void* __restrict__ in, int in_dtype, void* __restrict__ out, int out_dtype,
                          int64_t n) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        store_from_f32(out, out_dtype, t, load_as_f32(in, in_dtype, t));
}

extern "C" int b2me_convert(const void* in, int in_dtype, void* out, int out_dtype, int64_t n,
                            b2me_stream_t stream) {
    if (!in || !out || n < 0) return B2ME_EINVAL;
    if (n == 0) return B2ME_OK;
    int64_t blocks = ceil_div64(n, 256);
    if (blocks > B2ME_NUM_SMS * 16) blocks = B2ME_NUM_SMS * 16;
    k_convert<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, in_dtype, out, out_dtype, n);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ K5 small-N linear
// one warp per voxel row; lanes stride the input channels (coalesced row read), Cout <= 16 partial sums
// per lane, fixed-order butterfly reduction (deterministic), lane 0 writes logits and the arg-max.
#define LS_MAX_COUT 16
__global__ void __launch_bounds__(256)
k_linear_small(const void* __restrict__ in, int in_dtype, int64_t V, int Cin, const float* __restrict__ Wt,
               const float* __restrict__ bias, int Cout, float* __restrict__ out_logits,
               uint8_t* __restrict__ out_argmax) {
    extern __shared__ float w_s[];  // [Cout][Cin]
    for (int i = threadIdx.x; i < Cout * Cin; i += blockDim.x) w_s[i] = Wt[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    for (int64_t row = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < V;
         row += (int64_t)gridDim.x * warps_per_block) {
        float acc[LS_MAX_COUT];
#pragma unroll
        for (int o = 0; o < LS_MAX_COUT; ++o) acc[o] = 0.f;
        for (int c = lane; c < Cin; c += 32) {
            const float x = load_as_f32(in, in_dtype, row * Cin + c);
#pragma unroll
            for (int o = 0; o < LS_MAX_COUT; ++o)
                if (o < Cout) acc[o] = fmaf(x, w_s[o * Cin + c], acc[o]);
        }
#pragma unroll
        for (int o = 0; o < LS_MAX_COUT; ++o) {
            if (o < Cout) {
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) acc[o] += __shfl_xor_sync(0xffffffffu, acc[o], s);
            }
        }
        if (lane == 0) {
            float best = -INFINITY;
            int besti = 0;
#pragma unroll
            for (int o = 0; o < LS_MAX_COUT; ++o) {
                if (o < Cout) {
                    const float v = acc[o] + (bias ? bias[o] : 0.f);
                    if (out_logits) out_logits[row * Cout + o] = v;
                    if (v > best) {  // strict: lowest index wins ties
                        best = v;
                        besti = o;
                    }
                }
            }
            if (out_argmax) out_argmax[row] = (uint8_t)besti;
        }
    }
}

// Vectorised variant for the head (Cin = 1024 bf16 / f32 rows): a warp takes LS_ROWS rows at a time, every lane reads
// 16 bytes of each row per step (the warp covers 512 contiguous bytes), the weights come from shared memory once per
// step for all LS_ROWS rows. HBM-bound: V * Cin * e bytes in, V * (4 Cout + 1) out.
#ifndef LS_ROWS
#define LS_ROWS 4
#endif
#ifndef LS_BLOCKS_PER_SM
// persistent blocks per SM of the vectorised head linear. Measured on V = 8.9 M rows x 1024 bf16 -> 3 (tools/
// linear_probe.py, same box): 4 blocks 4.19 ms (4.35 TB/s), 8 blocks 3.40 ms (5.37 TB/s); 8 rows per warp: 4.8 ms
#define LS_BLOCKS_PER_SM 8
#endif
template <int CO, bool BF16>
__global__ void __launch_bounds__(256)
k_linear_small_vec(const void* __restrict__ in, int64_t V, int Cin, const float* __restrict__ Wt,
                   const float* __restrict__ bias, float* __restrict__ out_logits, uint8_t* __restrict__ out_argmax) {
    extern __shared__ float w_s[];  // [CO][Cin]
    for (int i = threadIdx.x; i < CO * Cin; i += blockDim.x) w_s[i] = Wt[i];
    __syncthreads();
    constexpr int VEC = BF16 ? 8 : 4;
    const int lane = threadIdx.x & 31;
    const int warps_per_block = blockDim.x >> 5;
    const int64_t ngroups = (V + LS_ROWS - 1) / LS_ROWS;
    for (int64_t grp = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); grp < ngroups;
         grp += (int64_t)gridDim.x * warps_per_block) {
        const int64_t row0 = grp * LS_ROWS;
        float acc[LS_ROWS][CO];
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
            for (int o = 0; o < CO; ++o) acc[r][o] = 0.f;
        for (int c = lane * VEC; c < Cin; c += 32 * VEC) {
            float x[LS_ROWS][VEC];
#pragma unroll
            for (int r = 0; r < LS_ROWS; ++r) {
                const int64_t row = row0 + r < V ? row0 + r : V - 1;  // clamped: the tail rows are not stored
                if (BF16) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(
                        reinterpret_cast<const __nv_bfloat16*>(in) + row * Cin + c));
                    const uint32_t wv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        x[r][2 * e] = __uint_as_float(wv[e] << 16);
                        x[r][2 * e + 1] = __uint_as_float(wv[e] & 0xFFFF0000u);
                    }
                } else {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) +
                                                                           row * Cin + c));
                    x[r][0] = v.x; x[r][1] = v.y; x[r][2] = v.z; x[r][3] = v.w;
                }
            }
#pragma unroll
            for (int o = 0; o < CO; ++o) {
                float w[VEC];
#pragma unroll
                for (int e = 0; e < VEC; e += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(w_s + o * Cin + c + e);
                    w[e] = t.x; w[e + 1] = t.y; w[e + 2] = t.z; w[e + 3] = t.w;
                }
#pragma unroll
                for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
                    for (int e = 0; e < VEC; ++e) acc[r][o] = fmaf(x[r][e], w[e], acc[r][o]);
            }
        }
#pragma unroll
        for (int r = 0; r < LS_ROWS; ++r)
#pragma unroll
            for (int o = 0; o < CO; ++o)
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) acc[r][o] += __shfl_xor_sync(0xffffffffu, acc[r][o], s);
        if (lane < LS_ROWS && row0 + lane < V) {
            const int64_t row = row0 + lane;
            float best = -INFINITY;
            int besti = 0;
#pragma unroll
            for (int r = 0; r < LS_ROWS; ++r) {
                if (r != lane) continue;
#pragma unroll
                for (int o = 0; o < CO; ++o) {
                    const float v = acc[r][o] + (bias ? bias[o] : 0.f);
                    if (out_logits) out_logits[row * CO + o] = v;
                    if (v > best) {  // strict: lowest index wins ties
                        best = v;
                        besti = o;
                    }
                }
            }
            if (out_argmax) out_argmax[row] = (uint8_t)besti;
        }
    }
}

template <int CO>
static void launch_linear_small_vec(const void* in, int in_dtype, int64_t V, int Cin, const float* Wt, const float* bias,
                                    float* out_logits, uint8_t* out_argmax, size_t smem, cudaStream_t s) {
    int64_t blocks = ceil_div64(ceil_div64(V, LS_ROWS), 8);
    if (blocks > B2ME_NUM_SMS * LS_BLOCKS_PER_SM) blocks = B2ME_NUM_SMS * LS_BLOCKS_PER_SM;
    if (in_dtype == B2ME_BF16) {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(k_linear_small_vec<CO, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_linear_small_vec<CO, true><<<(unsigned)blocks, 256, smem, s>>>(in, V, Cin, Wt, bias, out_logits, out_argmax);
    } else {
        if (smem > 48 * 1024)
            cudaFuncSetAttribute(k_linear_small_vec<CO, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        k_linear_small_vec<CO, false><<<(unsigned)blocks, 256, smem, s>>>(in, V, Cin, Wt, bias, out_logits, out_argmax);
    }
}

extern "C" int b2me_linear_small(const void* in, int in_dtype, int64_t V, int Cin, const float* Wt,
                                 const float* bias, int Cout, float* out_logits, uint8_t* out_argmax,
                                 b2me_stream_t stream) {
    if (!in || !Wt || V < 0 || Cin <= 0 || Cout <= 0 || Cout > LS_MAX_COUT) return B2ME_EINVAL;
    const size_t smem = (size_t)Cout * Cin * sizeof(float);
    if (smem > 200 * 1024) return B2ME_EUNSUPPORTED;
    if (V == 0) return B2ME_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int vec = in_dtype == B2ME_BF16 ? 8 : 4;
    if (Cin % vec == 0 && Cin >= 32 * vec) {
        switch (Cout) {
#define LS_CASE(N) case N: launch_linear_small_vec<N>(in, in_dtype, V, Cin, Wt, bias, out_logits, out_argmax, smem, s); break;
            LS_CASE(1) LS_CASE(2) LS_CASE(3) LS_CASE(4) LS_CASE(5) LS_CASE(6) LS_CASE(7) LS_CASE(8)
            LS_CASE(9) LS_CASE(10) LS_CASE(11) LS_CASE(12) LS_CASE(13) LS_CASE(14) LS_CASE(15) LS_CASE(16)
#undef LS_CASE
        }
        B2ME_CHECK_LAUNCH();
        return B2ME_OK;
    }
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(k_linear_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int64_t blocks = ceil_div64(V, 8);
    if (blocks > B2ME_NUM_SMS * 8) blocks = B2ME_NUM_SMS * 8;
    k_linear_small<<<(unsigned)blocks, 256, smem, s>>>(in, in_dtype, V, Cin, Wt, bias, Cout, out_logits, out_argmax);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ gathers
__global__ void k_gather_words(const uint32_t* __restrict__ in, int words_per_row, const int32_t* __restrict__ index,
                               int64_t total_words, uint32_t* __restrict__ out) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total_words;
         t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / words_per_row;
        const int w = (int)(t - i * words_per_row);
        const int32_t src = index[i];
        out[t] = src >= 0 ? in[(int64_t)src * words_per_row + w] : 0u;
    }
}

extern "C" int b2me_gather_rows(const void* in, int dtype, int C, const int32_t* index, int64_t N, void* out,
                                b2me_stream_t stream) {
    if (!in || !index || !out || C <= 0 || N < 0) return B2ME_EINVAL;
    const int row_bytes = C * (dtype == B2ME_BF16 ? 2 : 4);
    if (row_bytes % 4) return B2ME_EUNSUPPORTED;
    if (N == 0) return B2ME_OK;
    const int wpr = row_bytes / 4;
    const int64_t total = N * wpr;
    int64_t blocks = ceil_div64(total, 256);
    if (blocks > B2ME_NUM_SMS * 16) blocks = B2ME_NUM_SMS * 16;
    k_gather_words<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        reinterpret_cast<const uint32_t*>(in), wpr, index, total, reinterpret_cast<uint32_t*>(out));
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

__global__ void k_gather_labels(const uint8_t* __restrict__ lab, const int32_t* __restrict__ inverse, int64_t N,
                                uint8_t* __restrict__ out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = inverse[i];
        out[i] = v >= 0 ? lab[v] : (uint8_t)0;
    }
}

extern "C" int b2me_gather_labels(const uint8_t* voxel_labels, const int32_t* inverse, int64_t N, uint8_t* out,
                                  b2me_stream_t stream) {
    if (!voxel_labels || !inverse || !out || N < 0) return B2ME_EINVAL;
    if (N == 0) return B2ME_OK;
    int64_t blocks = ceil_div64(N, 256);
    if (blocks > B2ME_NUM_SMS * 16) blocks = B2ME_NUM_SMS * 16;
    k_gather_labels<<<(unsigned)blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(voxel_labels, inverse, N,
                                                                                         out);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}
