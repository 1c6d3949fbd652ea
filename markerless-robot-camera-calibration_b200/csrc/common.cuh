// common.cuh — shared device helpers of libb2me (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/b2me.h"

#define B2ME_NUM_SMS 148

#define B2ME_CHECK_LAUNCH()                                  \
    do {                                                     \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) return B2ME_ELAUNCH;         \
    } while (0)

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---------------------------------------------------------------- coordinate keys
// 64-bit key = b:10 | x:18 | y:18 | z:18, spatial fields biased by 2^17.
#define B2ME_KEY_EMPTY 0xFFFFFFFFFFFFFFFFull
#define B2ME_AXIS_BIAS 131072
#define B2ME_AXIS_MAX 262143

struct __align__(16) HashSlot {
    unsigned long long key;
    unsigned int val;
    unsigned int pad;
};

__device__ __forceinline__ bool coord_in_range(int b, int x, int y, int z) {
    return (unsigned)b < 1024u && (unsigned)(x + B2ME_AXIS_BIAS) <= (unsigned)B2ME_AXIS_MAX &&
           (unsigned)(y + B2ME_AXIS_BIAS) <= (unsigned)B2ME_AXIS_MAX &&
           (unsigned)(z + B2ME_AXIS_BIAS) <= (unsigned)B2ME_AXIS_MAX;
}

__device__ __forceinline__ unsigned long long pack_key(int b, int x, int y, int z) {
    return ((unsigned long long)(unsigned)b << 54) |
           ((unsigned long long)(unsigned)(x + B2ME_AXIS_BIAS) << 36) |
           ((unsigned long long)(unsigned)(y + B2ME_AXIS_BIAS) << 18) |
           ((unsigned long long)(unsigned)(z + B2ME_AXIS_BIAS));
}

__device__ __forceinline__ unsigned long long hash_key(unsigned long long k) {
    // murmur3 finaliser
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}

// Home slot of a key. The eight voxels of a 2x2x2 group (lowest bit of x, y, z) share one 128-byte line of the table
// (8 slots x 16 bytes): the group is hashed, the three low bits pick the slot inside the line. Depth-image points arrive
// in scan order, so consecutive points and the 27 neighbours of a voxel fall into few lines instead of one random
// 32-byte sector each. Coordinates that are multiples of a tensor stride >= 2 all start at slot 0 of their line and
// probe inside it (same line, no extra DRAM access); linear probing keeps the table exact whatever the layout.
__device__ __forceinline__ unsigned long long hash_slot(unsigned long long key, unsigned long long mask) {
    const unsigned long long low3 = (key & 1ull) | (((key >> 18) & 1ull) << 1) | (((key >> 36) & 1ull) << 2);
    const unsigned long long group = key & ~((1ull << 36) | (1ull << 18) | 1ull);
    return ((hash_key(group) & mask) & ~7ull) | low3;
}

// read-only probe of a finished table: returns val or 0xFFFFFFFF
__device__ __forceinline__ unsigned int table_lookup(const HashSlot* __restrict__ tab,
                                                     unsigned long long mask,
                                                     unsigned long long key) {
    unsigned long long s = hash_slot(key, mask);
    while (true) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(tab + s));
        const unsigned long long k = ((unsigned long long)v.y << 32) | v.x;
        if (k == key) return v.z;
        if (k == B2ME_KEY_EMPTY) return 0xFFFFFFFFu;
        s = (s + 1) & mask;
    }
}

// ---------------------------------------------------------------- dtype helpers
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float apply_act(float v, int act, float slope) {
    if (act == B2ME_ACT_RELU) return v > 0.f ? v : 0.f;
    if (act == B2ME_ACT_LEAKY) return v > 0.f ? v : v * slope;
    return v;
}

// round to nearest tf32 (10-bit mantissa), returned in an fp32 container
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

__device__ __forceinline__ float load_as_f32(const void* p, int dtype, int64_t i) {
    return dtype == B2ME_BF16 ? __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i])
                              : reinterpret_cast<const float*>(p)[i];
}
__device__ __forceinline__ void store_from_f32(void* p, int dtype, int64_t i, float v) {
    if (dtype == B2ME_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
    else reinterpret_cast<float*>(p)[i] = dtype == B2ME_TF32 ? round_tf32(v) : v;
}

// exclusive scan of an int32 array (in place), total written to *total. ws: >= scan_ws_bytes(n)
size_t scan_ws_bytes(int64_t n);
int exclusive_scan_i32(int32_t* data, int64_t n, int32_t* total, void* ws, cudaStream_t s);
