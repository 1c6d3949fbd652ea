// pointnet.cu — GPU-native PointNet++ primitives: farthest point sampling, ball query and
// 3-nearest-neighbour interpolation weights with the semantics of model/pointnet2_utils.py:65-143, 283-299
// (the reference builds dense [B,S,N] distance matrices in PyTorch and runs FPS as npoint sequential torch steps).
// The shared MLPs of the set-abstraction / feature-propagation layers stay library GEMMs (cuDNN / cuBLAS).
#include "common.cuh"

// ------------------------------------------------------------------------------------------ FPS
// farthest_point_sample (model/pointnet2_utils.py:65-86): one CTA per cloud, every thread keeps its points and their
// running min-distance in registers; per sample: distance update, block arg-max (ties: lowest index, torch.max).
// dist = ((x-cx)^2 + (y-cy)^2) + (z-cz)^2 in fp32 without FMA contraction, like torch.sum((xyz - c) ** 2, -1).
#define FPS_THREADS 512
#define FPS_PER 16  // points per thread: N <= 8192

__global__ void __launch_bounds__(FPS_THREADS)
k_fps(const float* __restrict__ xyz, int N, int npoint, const int32_t* __restrict__ start, int32_t* __restrict__ out) {
    __shared__ float red_v[FPS_THREADS / 32];
    __shared__ int red_i[FPS_THREADS / 32];
    __shared__ float cen[3];
    __shared__ int far_s;
    const int b = blockIdx.x;
    const float* p = xyz + (size_t)b * N * 3;
    float px[FPS_PER], py[FPS_PER], pz[FPS_PER], dist[FPS_PER];
#pragma unroll
    for (int k = 0; k < FPS_PER; ++k) {
        const int i = threadIdx.x + k * FPS_THREADS;
        if (i < N) { px[k] = p[i * 3]; py[k] = p[i * 3 + 1]; pz[k] = p[i * 3 + 2]; }
        else { px[k] = py[k] = pz[k] = 0.f; }
        dist[k] = 1e10f;
    }
    int farthest = start ? start[b] : 0;
    for (int s = 0; s < npoint; ++s) {
        if (threadIdx.x == 0) {
            out[(size_t)b * npoint + s] = farthest;
            cen[0] = p[farthest * 3]; cen[1] = p[farthest * 3 + 1]; cen[2] = p[farthest * 3 + 2];
        }
        __syncthreads();
        const float cx = cen[0], cy = cen[1], cz = cen[2];
        float bv = -1.f;
        int bi = 0x7FFFFFFF;
#pragma unroll
        for (int k = 0; k < FPS_PER; ++k) {
            const int i = threadIdx.x + k * FPS_THREADS;
            if (i < N) {
                const float dx = __fsub_rn(px[k], cx), dy = __fsub_rn(py[k], cy), dz = __fsub_rn(pz[k], cz);
                const float d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                if (d < dist[k]) dist[k] = d;
                if (dist[k] > bv) { bv = dist[k]; bi = i; }  // k ascending = index ascending: first maximum kept
            }
        }
        // block arg-max, ties to the lowest index
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = bv; red_i[threadIdx.x >> 5] = bi; }
        __syncthreads();
        if (threadIdx.x < 32) {
            float v = threadIdx.x < FPS_THREADS / 32 ? red_v[threadIdx.x] : -2.f;
            int i2 = threadIdx.x < FPS_THREADS / 32 ? red_i[threadIdx.x] : 0x7FFFFFFF;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int oi = __shfl_xor_sync(0xffffffffu, i2, o);
                if (ov > v || (ov == v && oi < i2)) { v = ov; i2 = oi; }
            }
            if (threadIdx.x == 0) far_s = i2;
        }
        __syncthreads();
        farthest = far_s;
    }
}

extern "C" int b2me_fps(const float* xyz, int B, int N, int npoint, const int32_t* start, int32_t* out_idx,
                        b2me_stream_t stream) {
    if (!xyz || !out_idx || B < 0 || N <= 0 || npoint <= 0) return B2ME_EINVAL;
    if (N > FPS_THREADS * FPS_PER) return B2ME_EUNSUPPORTED;
    if (B == 0) return B2ME_OK;
    k_fps<<<(unsigned)B, FPS_THREADS, 0, reinterpret_cast<cudaStream_t>(stream)>>>(xyz, N, npoint, start, out_idx);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ ball query
// query_ball_point (model/pointnet2_utils.py:89-110): the first `nsample` point indices (ascending) whose squared
// distance to the query is <= radius^2, padded with the first one; a query with no point in range gets N (the
// reference's sentinel) in every slot. Squared distance in the reference's expanded form
// -2 <q,p> + |q|^2 + |p|^2 (square_distance, :22-44), fp32. One warp per query, ballot + prefix popcount.
__global__ void __launch_bounds__(256)
k_ball_query(const float* __restrict__ xyz, const float* __restrict__ new_xyz, int B, int N, int S, float r2, int nsample,
             int32_t* __restrict__ out) {
    const long long q = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (q >= (long long)B * S) return;
    const int lane = threadIdx.x & 31;
    const int b = (int)(q / S);
    const float* p = xyz + (size_t)b * N * 3;
    const float qx = new_xyz[q * 3], qy = new_xyz[q * 3 + 1], qz = new_xyz[q * 3 + 2];
    const float qq = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
    int32_t* o = out + q * nsample;
    int found = 0, first = N;
    for (int base = 0; base < N && found < nsample; base += 32) {
        const int i = base + lane;
        bool in = false;
        if (i < N) {
            const float x = p[i * 3], y = p[i * 3 + 1], z = p[i * 3 + 2];
            const float dot = fmaf(qz, z, fmaf(qy, y, __fmul_rn(qx, x)));
            const float pp = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
            const float d = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), qq), pp);
            in = !(d > r2);
        }
        const unsigned m = __ballot_sync(0xffffffffu, in);
        if (m) {
            if (first == N) first = base + __ffs((int)m) - 1;
            const int pos = found + __popc(m & ((1u << lane) - 1u));
            if (in && pos < nsample) o[pos] = i;
            found += __popc(m);
        }
    }
    if (found > nsample) found = nsample;
    for (int s = found + lane; s < nsample; s += 32) o[s] = first;
}

extern "C" int b2me_ball_query(const float* xyz, const float* new_xyz, int B, int N, int S, float radius, int nsample,
                               int32_t* out_idx, b2me_stream_t stream) {
    if (!xyz || !new_xyz || !out_idx || B < 0 || N <= 0 || S <= 0 || nsample <= 0 || !(radius > 0)) return B2ME_EINVAL;
    if (B == 0) return B2ME_OK;
    const long long nq = (long long)B * S;
    k_ball_query<<<(unsigned)ceil_div64(nq, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        xyz, new_xyz, B, N, S, radius * radius, nsample, out_idx);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ 3-NN
// PointNetFeaturePropagation (model/pointnet2_utils.py:283-293): for every point of xyz1 the 3 nearest points of xyz2
// (expanded-form squared distances, ascending, ties to the lowest index) and the inverse-distance weights
// w = (1 / (d + 1e-8)) / sum. One thread per xyz1 point, xyz2 of the cloud staged in shared memory.
__global__ void __launch_bounds__(256)
k_three_nn(const float* __restrict__ xyz1, const float* __restrict__ xyz2, int N, int S, int32_t* __restrict__ out_idx,
           float* __restrict__ out_w) {
    extern __shared__ float s2[];  // [S][4] = x, y, z, |p|^2
    const int b = blockIdx.y;
    const float* p2 = xyz2 + (size_t)b * S * 3;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const float x = p2[i * 3], y = p2[i * 3 + 1], z = p2[i * 3 + 2];
        s2[i * 4] = x; s2[i * 4 + 1] = y; s2[i * 4 + 2] = z;
        s2[i * 4 + 3] = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
    }
    __syncthreads();
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const float* p1 = xyz1 + ((size_t)b * N + n) * 3;
    const float qx = p1[0], qy = p1[1], qz = p1[2];
    const float qq = __fadd_rn(__fadd_rn(__fmul_rn(qx, qx), __fmul_rn(qy, qy)), __fmul_rn(qz, qz));
    float d0 = INFINITY, d1 = INFINITY, d2 = INFINITY;
    int i0 = 0, i1 = 0, i2 = 0;
    for (int i = 0; i < S; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(s2 + i * 4);
        const float dot = fmaf(qz, v.z, fmaf(qy, v.y, __fmul_rn(qx, v.x)));
        const float d = __fadd_rn(__fadd_rn(__fmul_rn(-2.f, dot), qq), v.w);
        if (d < d2) {
            if (d < d1) {
                d2 = d1; i2 = i1;
                if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = i; }
                else { d1 = d; i1 = i; }
            } else { d2 = d; i2 = i; }
        }
    }
    if (S < 3) { if (S < 2) { d1 = d0; i1 = i0; } d2 = d1; i2 = i1; }
    const float r0 = 1.f / (d0 + 1e-8f), r1 = 1.f / (d1 + 1e-8f), r2 = 1.f / (d2 + 1e-8f);
    const float nrm = __fadd_rn(__fadd_rn(r0, r1), r2);
    const size_t o = ((size_t)b * N + n) * 3;
    out_idx[o] = i0; out_idx[o + 1] = i1; out_idx[o + 2] = i2;
    out_w[o] = r0 / nrm; out_w[o + 1] = r1 / nrm; out_w[o + 2] = r2 / nrm;
}

extern "C" int b2me_three_nn(const float* xyz1, const float* xyz2, int B, int N, int S, int32_t* out_idx, float* out_w,
                             b2me_stream_t stream) {
    if (!xyz1 || !xyz2 || !out_idx || !out_w || B < 0 || N <= 0 || S <= 0) return B2ME_EINVAL;
    if (B == 0) return B2ME_OK;
    const size_t smem = (size_t)S * 16;
    if (smem > 200 * 1024) return B2ME_EUNSUPPORTED;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(k_three_nn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return B2ME_ELAUNCH;
    dim3 grid((unsigned)((N + 255) / 256), (unsigned)B);
    k_three_nn<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(xyz1, xyz2, N, S, out_idx, out_w);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}
