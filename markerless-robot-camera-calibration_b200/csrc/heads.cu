// heads.cu — K6 global pooling, K8 key-point / vote reductions, fused "magic" translation.
// All are small HBM/latency-bound reductions; each is deterministic (fixed reduction order).
#include "common.cuh"

// ------------------------------------------------------------------------------------------ K6
// grid (B, ceil(C/32)); 8 warps split the row range into 8 contiguous slices, lane = channel.
__global__ void __launch_bounds__(256)
k_global_pool(const void* __restrict__ in, int dtype, const int4* __restrict__ coords, int64_t V, int C, int mode,
              float* __restrict__ out) {
    __shared__ float part[8][32];
    __shared__ int cnt[8];
    const int b = blockIdx.x;
    const int c = blockIdx.y * 32 + (threadIdx.x & 31);
    const int w = threadIdx.x >> 5;
    const int64_t r0 = V * w / 8, r1 = V * (w + 1) / 8;
    float acc = (mode == 1) ? -INFINITY : 0.f;
    int n = 0;
    for (int64_t r = r0; r < r1; ++r) {
        if (__ldg(&coords[r].x) != b) continue;
        ++n;
        if (c < C) {
            const float v = load_as_f32(in, dtype, r * C + c);
            acc = (mode == 1) ? fmaxf(acc, v) : acc + v;
        }
    }
    part[w][threadIdx.x & 31] = acc;
    if ((threadIdx.x & 31) == 0) cnt[w] = n;
    __syncthreads();
    if (w == 0 && c < C) {
        float tot = part[0][threadIdx.x];
        int nt = cnt[0];
        for (int i = 1; i < 8; ++i) {
            tot = (mode == 1) ? fmaxf(tot, part[i][threadIdx.x]) : tot + part[i][threadIdx.x];
            nt += cnt[i];
        }
        if (mode == 0) tot = nt > 0 ? tot / (float)nt : 0.f;
        else if (nt == 0) tot = 0.f;
        out[(int64_t)b * C + c] = tot;
    }
}

extern "C" int b2me_global_pool(const void* in, int dtype, const int32_t* coords, int64_t V, int C, int B, int mode,
                                float* out, b2me_stream_t stream) {
    if (!in || !coords || !out || V < 0 || C <= 0 || B <= 0 || (mode != 0 && mode != 1)) return B2ME_EINVAL;
    dim3 grid((unsigned)B, (unsigned)((C + 31) / 32));
    k_global_pool<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        in, dtype, reinterpret_cast<const int4*>(coords), V, C, mode, out);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ K8a
#define KP_MAX_K 16
// one block per segment; each thread keeps (best prob, lowest index) per class over its rows, then a
// fixed-order block reduction. prob = exp(x_c - max) / sum_j exp(x_j - max) in fp32.
__global__ void __launch_bounds__(256)
k_keypoint_reduce(const float* __restrict__ logits, int K, const int32_t* __restrict__ seg_offsets,
                  float* __restrict__ best_prob, int32_t* __restrict__ best_idx) {
    __shared__ float sp[256];
    __shared__ int si[256];
    const int seg = blockIdx.x;
    const int r0 = seg_offsets[seg], r1 = seg_offsets[seg + 1];
    float bp[KP_MAX_K];
    int bi[KP_MAX_K];
#pragma unroll
    for (int c = 0; c < KP_MAX_K; ++c) { bp[c] = -1.f; bi[c] = 0x7FFFFFFF; }
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        float x[KP_MAX_K];
        float m = -INFINITY;
#pragma unroll
        for (int c = 0; c < KP_MAX_K; ++c) {
            x[c] = c < K ? logits[(int64_t)r * K + c] : -INFINITY;
            m = fmaxf(m, x[c]);
        }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < KP_MAX_K; ++c) {
            x[c] = c < K ? expf(x[c] - m) : 0.f;
            s += x[c];
        }
#pragma unroll
        for (int c = 0; c < KP_MAX_K; ++c) {
            const float pr = x[c] / s;
            if (c < K && pr > bp[c]) { bp[c] = pr; bi[c] = r; }  // rows visited in increasing order per thread
        }
    }
    for (int c = 0; c < K; ++c) {
        sp[threadIdx.x] = bp[c];
        si[threadIdx.x] = bi[c];
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const float p2 = sp[threadIdx.x + o];
                const int i2 = si[threadIdx.x + o];
                if (p2 > sp[threadIdx.x] || (p2 == sp[threadIdx.x] && i2 < si[threadIdx.x])) {
                    sp[threadIdx.x] = p2;
                    si[threadIdx.x] = i2;
                }
            }
            __syncthreads();
        }
        if (threadIdx.x == 0) {
            best_prob[seg * K + c] = (r1 > r0) ? sp[0] : 0.f;
            best_idx[seg * K + c] = (r1 > r0) ? si[0] : -1;
        }
        __syncthreads();
    }
}

extern "C" int b2me_keypoint_reduce(const float* logits, int K, const int32_t* seg_offsets, int S, float* best_prob,
                                    int32_t* best_idx, b2me_stream_t stream) {
    if (!logits || !seg_offsets || !best_prob || !best_idx || K <= 0 || K > KP_MAX_K || S < 0) return B2ME_EINVAL;
    if (S == 0) return B2ME_OK;
    k_keypoint_reduce<<<(unsigned)S, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, K, seg_offsets,
                                                                                        best_prob, best_idx);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ K8b
// top-k by repeated selection: round t picks the largest (value, lowest index) strictly after the
// previous pick in (value desc, index asc) order. k is 8 in the reference, segments are ~10^4 rows.
__global__ void __launch_bounds__(256)
k_vote_center(const float* __restrict__ logits, int K, int col, const float* __restrict__ pts,
              const int32_t* __restrict__ seg_offsets, int topk, float* __restrict__ out_center) {
    __shared__ float sv[256];
    __shared__ int si[256];
    __shared__ float prev_v;
    __shared__ int prev_i;
    const int seg = blockIdx.x;
    const int r0 = seg_offsets[seg], r1 = seg_offsets[seg + 1];
    if (threadIdx.x == 0) { prev_v = INFINITY; prev_i = -1; }
    __syncthreads();
    float cx = 0.f, cy = 0.f, cz = 0.f;
    int picked = 0;
    for (int t = 0; t < topk; ++t) {
        const float pv = prev_v;
        const int pi = prev_i;
        float bv = -INFINITY;
        int bi = 0x7FFFFFFF;
        for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
            const float v = logits[(int64_t)r * K + col];
            const bool after_prev = (v < pv) || (v == pv && r > pi);
            if (after_prev && (v > bv || (v == bv && r < bi))) { bv = v; bi = r; }
        }
        sv[threadIdx.x] = bv;
        si[threadIdx.x] = bi;
        __syncthreads();
        for (int o = 128; o > 0; o >>= 1) {
            if (threadIdx.x < o) {
                const float v2 = sv[threadIdx.x + o];
                const int i2 = si[threadIdx.x + o];
                if (v2 > sv[threadIdx.x] || (v2 == sv[threadIdx.x] && i2 < si[threadIdx.x])) {
                    sv[threadIdx.x] = v2;
                    si[threadIdx.x] = i2;
                }
            }
            __syncthreads();
        }
        const int sel = si[0];
        const float selv = sv[0];
        __syncthreads();
        if (sel == 0x7FFFFFFF) break;  // fewer than topk rows (uniform across the block)
        if (threadIdx.x == 0) {
            prev_v = selv;
            prev_i = sel;
            cx += pts[(int64_t)sel * 3 + 0];
            cy += pts[(int64_t)sel * 3 + 1];
            cz += pts[(int64_t)sel * 3 + 2];
        }
        ++picked;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float inv = picked > 0 ? 1.f / (float)picked : 0.f;
        out_center[seg * 3 + 0] = cx * inv;
        out_center[seg * 3 + 1] = cy * inv;
        out_center[seg * 3 + 2] = cz * inv;
    }
}

extern "C" int b2me_vote_center(const float* logits, int K, int col, const float* points_xyz,
                                const int32_t* seg_offsets, int S, int topk, float* out_center,
                                b2me_stream_t stream) {
    if (!logits || !points_xyz || !seg_offsets || !out_center || K <= 0 || col < 0 || col >= K || S < 0 || topk <= 0)
        return B2ME_EINVAL;
    if (S == 0) return B2ME_OK;
    k_vote_center<<<(unsigned)S, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(logits, K, col, points_xyz,
                                                                                    seg_offsets, topk, out_center);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ translation
__device__ __forceinline__ void quat_to_rot_f32(const float* q, float R[9]) {
    // utils/transformation.py:16-60 (w first), evaluated in fp32 like the reference's float32 numpy scalars
    const float q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
    R[0] = 2 * (q0 * q0 + q1 * q1) - 1; R[1] = 2 * (q1 * q2 - q0 * q3);     R[2] = 2 * (q1 * q3 + q0 * q2);
    R[3] = 2 * (q1 * q2 + q0 * q3);     R[4] = 2 * (q0 * q0 + q2 * q2) - 1; R[5] = 2 * (q2 * q3 - q0 * q1);
    R[6] = 2 * (q1 * q3 - q0 * q2);     R[7] = 2 * (q2 * q3 + q0 * q1);     R[8] = 2 * (q0 * q0 + q3 * q3) - 1;
}

__global__ void __launch_bounds__(256)
k_translation_magic(const float* __restrict__ pts, const int32_t* __restrict__ seg_offsets,
                    const float* __restrict__ quat, float x_offset, double* __restrict__ out_pos) {
    __shared__ float smin[3][256], smax[3][256];
    const int seg = blockIdx.x;
    const int r0 = seg_offsets[seg], r1 = seg_offsets[seg + 1];
    float R[9];
    quat_to_rot_f32(quat + seg * 4, R);
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        const float x = pts[(int64_t)r * 3], y = pts[(int64_t)r * 3 + 1], z = pts[(int64_t)r * 3 + 2];
        // p' = R^T p
        const float px = __fadd_rn(__fadd_rn(__fmul_rn(R[0], x), __fmul_rn(R[3], y)), __fmul_rn(R[6], z));
        const float py = __fadd_rn(__fadd_rn(__fmul_rn(R[1], x), __fmul_rn(R[4], y)), __fmul_rn(R[7], z));
        const float pz = __fadd_rn(__fadd_rn(__fmul_rn(R[2], x), __fmul_rn(R[5], y)), __fmul_rn(R[8], z));
        mn[0] = fminf(mn[0], px); mx[0] = fmaxf(mx[0], px);
        mn[1] = fminf(mn[1], py); mx[1] = fmaxf(mx[1], py);
        mn[2] = fminf(mn[2], pz); mx[2] = fmaxf(mx[2], pz);
    }
    for (int a = 0; a < 3; ++a) { smin[a][threadIdx.x] = mn[a]; smax[a][threadIdx.x] = mx[a]; }
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            for (int a = 0; a < 3; ++a) {
                smin[a][threadIdx.x] = fminf(smin[a][threadIdx.x], smin[a][threadIdx.x + o]);
                smax[a][threadIdx.x] = fmaxf(smax[a][threadIdx.x], smax[a][threadIdx.x + o]);
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (r1 <= r0) {
            out_pos[seg * 3] = out_pos[seg * 3 + 1] = out_pos[seg * 3 + 2] = 0.0;
            return;
        }
        float o[3];
        for (int a = 0; a < 3; ++a) o[a] = __fmul_rn(__fadd_rn(smax[a][0], smin[a][0]), 0.5f);  // (max+min)/2
        const float min_z = __fsub_rn(smin[2][0], o[2]);
        // float64 from here on, like numpy's promotion of [-0.015, 0.0, min_z] + offset (inference_engine.py:484-487)
        const double m[3] = {(double)x_offset + (double)o[0], 0.0 + (double)o[1], (double)min_z + (double)o[2]};
        for (int a = 0; a < 3; ++a)
            out_pos[seg * 3 + a] = (double)R[a * 3] * m[0] + (double)R[a * 3 + 1] * m[1] + (double)R[a * 3 + 2] * m[2];
    }
}

extern "C" int b2me_translation_magic(const float* points_xyz, const int32_t* seg_offsets, int S,
                                      const float* quat_wxyz, float x_offset, double* out_pos,
                                      b2me_stream_t stream) {
    if (!points_xyz || !seg_offsets || !quat_wxyz || !out_pos || S < 0) return B2ME_EINVAL;
    if (S == 0) return B2ME_OK;
    k_translation_magic<<<(unsigned)S, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(points_xyz, seg_offsets,
                                                                                          quat_wxyz, x_offset, out_pos);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ sanity check
// InferenceEngine.check_sanity (app/inference_engine.py:246-279) with get_6_key_points (utils/data.py:255-335,
// switch_w = False, euclidean_threshold = 0.04) and compute_kp_error (utils/metrics.py:130-136), batched: one CTA per
// EE crop, float64 like the reference's NumPy. Two passes over the crop: (1) EE-frame coordinates, the four corner
// arg-mins over the "front" points, the largest z of each gripper side; (2) the gripper points nearest to the lifted
// probes. Arg-mins keep the lowest index among equal distances (np.argmin).
struct SanityArg {
    double v;
    int i;
};
__device__ __forceinline__ SanityArg sanity_min(SanityArg a, SanityArg b) {
    return (b.v < a.v || (b.v == a.v && b.i < a.i)) ? b : a;
}
__device__ __forceinline__ SanityArg sanity_block_argmin(SanityArg x, SanityArg* sh) {
    for (int o = 16; o > 0; o >>= 1) {
        SanityArg y;
        y.v = __shfl_xor_sync(0xffffffffu, x.v, o);
        y.i = __shfl_xor_sync(0xffffffffu, x.i, o);
        x = sanity_min(x, y);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    SanityArg r = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = sanity_min(r, sh[w]);
    return r;
}
__device__ __forceinline__ double sanity_block_max(double x, double* sh) {
    for (int o = 16; o > 0; o >>= 1) x = fmax(x, __shfl_xor_sync(0xffffffffu, x, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    double r = sh[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = fmax(r, sh[w]);
    return r;
}
__device__ __forceinline__ int sanity_block_sum(int x, int* sh) {
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = x;
    __syncthreads();
    int r = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) r += sh[w];
    return r;
}

__global__ void __launch_bounds__(256)
k_sanity_check(const float* __restrict__ pts, const int32_t* __restrict__ seg_offsets, const double* __restrict__ poses,
               const float* __restrict__ kp_prob, const float* __restrict__ kp_xyz, int K, float kp_threshold,
               int min_points, double kp_margin, uint8_t* __restrict__ out_confident) {
    __shared__ SanityArg sh_arg[8];
    __shared__ double sh_d[8];
    __shared__ int sh_i[8];
    // utils/data.py:264-271 key-point template and :280-285 corner probes (EE frame)
    const double tmpl[6][3] = {{0.02, 0.09, 0.0}, {0.01, -0.1, 0.0}, {0.014, 0.095, 0.07}, {0.014, -0.095, 0.07},
                               {0.0, 0.048, 0.12}, {0.0, -0.048, 0.12}};
    const double probe[4][3] = {{0.24, 0.32, -0.2}, {0.24, -0.32, -0.2}, {0.24, 0.32, 0.2}, {0.24, -0.32, 0.2}};
    const int seg = blockIdx.x;
    const int r0 = seg_offsets[seg], r1 = seg_offsets[seg + 1];
    const int n = r1 - r0;
    const double* pose = poses + (int64_t)seg * 7;
    const double q0 = pose[3], q1 = pose[4], q2 = pose[5], q3 = pose[6];
    const double R[3][3] = {{2 * (q0 * q0 + q1 * q1) - 1, 2 * (q1 * q2 - q0 * q3), 2 * (q1 * q3 + q0 * q2)},
                            {2 * (q1 * q2 + q0 * q3), 2 * (q0 * q0 + q2 * q2) - 1, 2 * (q2 * q3 - q0 * q1)},
                            {2 * (q1 * q3 - q0 * q2), 2 * (q2 * q3 + q0 * q1), 2 * (q0 * q0 + q3 * q3) - 1}};
    double origin[3];
    for (int c = 0; c < 3; ++c) origin[c] = pose[0] * R[0][c] + pose[1] * R[1][c] + pose[2] * R[2][c];
    auto local_of = [&](int r, double* l) {
        const double x = pts[(int64_t)r * 3], y = pts[(int64_t)r * 3 + 1], z = pts[(int64_t)r * 3 + 2];
        for (int c = 0; c < 3; ++c) l[c] = (x * R[0][c] + y * R[1][c] + z * R[2][c]) - origin[c];
    };
    int nkp = 0;  // predicted key points above the confidence threshold (ResultDTO.key_points)
    if (kp_prob)
        for (int k = 0; k < K; ++k) nkp += kp_prob[(int64_t)seg * K + k] > kp_threshold ? 1 : 0;

    // ---- pass 1
    SanityArg corner[4];
    for (int c = 0; c < 4; ++c) { corner[c].v = INFINITY; corner[c].i = 0x7FFFFFFF; }
    int nfront = 0, cnt_a = 0, cnt_b = 0;
    double maxz_a = -INFINITY, maxz_b = -INFINITY;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
        double l[3];
        local_of(r, l);
        if (l[0] > -0.005 && l[2] < 0.09) {
            ++nfront;
            for (int c = 0; c < 4; ++c) {
                const double dx = probe[c][0] - l[0], dy = probe[c][1] - l[1], dz = probe[c][2] - l[2];
                SanityArg cand;
                cand.v = sqrt(dx * dx + dy * dy + dz * dz);
                cand.i = r - r0;
                corner[c] = sanity_min(corner[c], cand);
            }
        }
        if (l[2] > 0.08) {
            if (l[1] > 0) { ++cnt_a; maxz_a = fmax(maxz_a, l[2]); }
            if (l[1] < 0) { ++cnt_b; maxz_b = fmax(maxz_b, l[2]); }
        }
    }
    for (int c = 0; c < 4; ++c) corner[c] = sanity_block_argmin(corner[c], sh_arg);
    nfront = sanity_block_sum(nfront, sh_i);
    cnt_a = sanity_block_sum(cnt_a, sh_i);
    cnt_b = sanity_block_sum(cnt_b, sh_i);
    maxz_a = sanity_block_max(maxz_a, sh_d);
    maxz_b = sanity_block_max(maxz_b, sh_d);

    // ---- pass 2: gripper points nearest to the probes lifted to the largest z of their side (get_closest_point)
    SanityArg grip[2];
    grip[0].v = grip[1].v = INFINITY;
    grip[0].i = grip[1].i = 0x7FFFFFFF;
    if (nkp > 3 && (cnt_a > 0 || cnt_b > 0)) {
        for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) {
            double l[3];
            local_of(r, l);
            if (l[2] > 0.08) {
                if (l[1] > 0) {
                    const double dy = l[1] - 0.01, dz = l[2] - maxz_a;
                    SanityArg cand;
                    cand.v = sqrt(l[0] * l[0] + dy * dy + dz * dz);
                    cand.i = r - r0;
                    grip[0] = sanity_min(grip[0], cand);
                }
                if (l[1] < 0) {
                    const double dy = l[1] + 0.01, dz = l[2] - maxz_b;
                    SanityArg cand;
                    cand.v = sqrt(l[0] * l[0] + dy * dy + dz * dz);
                    cand.i = r - r0;
                    grip[1] = sanity_min(grip[1], cand);
                }
            }
        }
    }
    grip[0] = sanity_block_argmin(grip[0], sh_arg);
    grip[1] = sanity_block_argmin(grip[1], sh_arg);
    if (threadIdx.x != 0) return;

    // ---- verdict (app/inference_engine.py:252-279)
    uint8_t ok = 1;
    if (n < min_points) {
        ok = 0;                                            // "fail min # points"
    } else if (nfront < 1) {
        // get_6_key_points returns empty arrays: the corner test is vacuous, compute_kp_error returns 100
        if (nkp > 3 && 100.0 > kp_margin) ok = 0;
    } else {
        double kps[6][3];
        for (int k = 0; k < 6; ++k)
            for (int c = 0; c < 3; ++c) kps[k][c] = tmpl[k][c];
        for (int c = 0; c < 4 && ok; ++c) {
            double l[3];
            local_of(r0 + corner[c].i, l);
            const double dx = tmpl[c][0] - l[0], dy = tmpl[c][1] - l[1], dz = tmpl[c][2] - l[2];
            if (sqrt(dx * dx + dy * dy + dz * dz) < 0.04) {
                for (int a = 0; a < 3; ++a) kps[c][a] = l[a];
            } else {
                ok = 0;                                    // "fail dim check": a corner is not where the pose puts it
            }
        }
        if (ok && nkp > 3) {
            bool found[2] = {cnt_a > 0, cnt_b > 0};
            for (int sd = 0; sd < 2; ++sd)
                if (found[sd]) local_of(r0 + grip[sd].i, kps[4 + sd]);
            if (!found[0] && found[1]) { kps[4][0] = kps[5][0]; kps[4][1] = -kps[5][1]; kps[4][2] = kps[5][2]; }
            else if (found[0] && !found[1]) { kps[5][0] = kps[4][0]; kps[5][1] = -kps[4][1]; kps[5][2] = kps[4][2]; }
            const double zz = fmax(kps[4][2], kps[5][2]);
            kps[4][2] = kps[5][2] = zz;
            double err = 0.0;
            int cnt = 0;
            for (int k = 0; k < K; ++k) {
                if (!(kp_prob[(int64_t)seg * K + k] > kp_threshold)) continue;
                if (k >= 6) { err = INFINITY; ++cnt; continue; }   // no such template key point
                double cam[3];
                for (int r = 0; r < 3; ++r)
                    cam[r] = (kps[k][0] + origin[0]) * R[r][0] + (kps[k][1] + origin[1]) * R[r][1] +
                             (kps[k][2] + origin[2]) * R[r][2];
                const float* pk = kp_xyz + ((int64_t)seg * K + k) * 3;   // float32 coordinates, like the reference's array
                const double dx = cam[0] - (double)pk[0], dy = cam[1] - (double)pk[1], dz = cam[2] - (double)pk[2];
                err += sqrt(dx * dx + dy * dy + dz * dz);
                ++cnt;
            }
            if (cnt >= 2 && err / cnt > kp_margin) ok = 0;  // "fail kp error margin"
        }
    }
    out_confident[seg] = ok;
}

extern "C" int b2me_sanity_check(const float* points_xyz, const int32_t* seg_offsets, int S, const double* ee_pose,
                                 const float* kp_prob, const float* kp_xyz, int K, float kp_threshold, int min_points,
                                 double kp_margin, uint8_t* out_confident, b2me_stream_t stream) {
    if (!points_xyz || !seg_offsets || !ee_pose || !out_confident || S < 0 || K < 0) return B2ME_EINVAL;
    if (K > 0 && (!kp_prob || !kp_xyz)) return B2ME_EINVAL;
    if (S == 0) return B2ME_OK;
    k_sanity_check<<<(unsigned)S, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        points_xyz, seg_offsets, ee_pose, K > 0 ? kp_prob : nullptr, kp_xyz, K, kp_threshold, min_points, kp_margin,
        out_confident);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}
