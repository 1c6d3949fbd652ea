// coords.cu — K1 voxelise / unique, K2 stride map, K3 kernel maps (sm_100a).
//
// All three are HBM-bound integer kernels: one thread per point / voxel / (voxel, offset), coalesced
// streaming reads and writes, random 16-byte sector accesses into an open-addressing hash table.
// Determinism: the voxel that owns a key is the LOWEST point index that maps to it (atomicMin), and a
// voxel's row is the exclusive prefix sum of the "I am the first point of my voxel" flags in point
// order, so rows come out in first-occurrence order without a sort.
#include "common.cuh"

// ------------------------------------------------------------------------------------------ scan
#define SCAN_THREADS 512
#define SCAN_ITEMS 4
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

size_t scan_ws_bytes(int64_t n) {
    int64_t nb = ceil_div64(n > 0 ? n : 1, SCAN_TILE);
    return align_up((size_t)(nb + 1) * sizeof(int32_t), 256);
}

__device__ __forceinline__ int block_exclusive_scan(int v, int* total, int* smem /*>=32*/) {
    // returns exclusive prefix of v over the block; *total = block sum
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) smem[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        int w = (lane < (blockDim.x >> 5)) ? smem[lane] : 0;
        int winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        smem[lane] = winc - w;  // exclusive warp offsets
        if (lane == 31) smem[32] = winc;
    }
    __syncthreads();
    int res = inc - v + smem[warp];
    *total = smem[32];
    __syncthreads();
    return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_reduce(const int32_t* __restrict__ data, int64_t n,
                                                              int32_t* __restrict__ block_sums) {
    __shared__ int sm[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j)
        if (base + j < n) s += data[base + j];
    int total;
    block_exclusive_scan(s, &total, sm);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_block_sums(int32_t* __restrict__ block_sums, int64_t nb,
                                                         int32_t* __restrict__ total_out) {
    __shared__ int sm[33];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int64_t base = 0; base < nb; base += blockDim.x) {
        int64_t i = base + threadIdx.x;
        int v = (i < nb) ? block_sums[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total, sm);
        int carry = carry_s;
        if (i < nb) block_sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total_out = carry_s;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_apply(int32_t* __restrict__ data, int64_t n,
                                                             const int32_t* __restrict__ block_sums) {
    __shared__ int sm[33];
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int s = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        v[j] = (base + j < n) ? data[base + j] : 0;
        s += v[j];
    }
    int total;
    int ex = block_exclusive_scan(s, &total, sm) + block_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; ++j) {
        if (base + j < n) data[base + j] = ex;
        ex += v[j];
    }
}

int exclusive_scan_i32(int32_t* data, int64_t n, int32_t* total, void* ws, cudaStream_t s) {
    if (n <= 0) {
        cudaMemsetAsync(total, 0, sizeof(int32_t), s);
        return B2ME_OK;
    }
    int32_t* block_sums = reinterpret_cast<int32_t*>(ws);
    const int64_t nb = ceil_div64(n, SCAN_TILE);
    k_scan_reduce<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(data, n, block_sums);
    k_scan_block_sums<<<1, 1024, 0, s>>>(block_sums, nb, total);
    k_scan_apply<<<(unsigned)nb, SCAN_THREADS, 0, s>>>(data, n, block_sums);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ K1
__global__ void k_quantize_float(const float4* __restrict__ coords, int64_t n, int4* __restrict__ q) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 c = __ldg(coords + i);
    // the caller has already multiplied by `scale` in fp32 (app/inference_engine.py:408); ME only floors
    q[i] = make_int4((int)c.x, (int)floorf(c.y), (int)floorf(c.z), (int)floorf(c.w));
}

__device__ __forceinline__ int floor_div(int a, int b) {  // b > 0
    int q = a / b;
    return (a % b != 0 && a < 0) ? q - 1 : q;
}

__global__ void k_quantize_stride(const int4* __restrict__ in, int64_t n, int ts_out, int4* __restrict__ q,
                                  uint8_t* __restrict__ koff) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = __ldg(in + i);
    const int ts_in = ts_out >> 1;
    const int px = floor_div(c.y, ts_out) * ts_out;
    const int py = floor_div(c.z, ts_out) * ts_out;
    const int pz = floor_div(c.w, ts_out) * ts_out;
    q[i] = make_int4(c.x, px, py, pz);
    const int dx = (c.y - px) / ts_in, dy = (c.z - py) / ts_in, dz = (c.w - pz) / ts_in;
    koff[i] = (uint8_t)(dx + 2 * dy + 4 * dz);  // x fastest
}

__global__ void k_hash_insert(const int4* __restrict__ q, int64_t n, HashSlot* __restrict__ tab,
                              unsigned long long mask, int32_t* __restrict__ slot_of,
                              int32_t* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int4 c = __ldg(q + i);
    if (!coord_in_range(c.x, c.y, c.z, c.w)) {
        atomicOr(reinterpret_cast<unsigned int*>(counts + 1), 1u);
        slot_of[i] = -1;
        return;
    }
    const unsigned long long key = pack_key(c.x, c.y, c.z, c.w);
    unsigned long long s = hash_slot(key, mask);
    while (true) {
        unsigned long long prev = tab[s].key;  // cheap pre-read: most probes hit an owned slot
        if (prev != key) {
            if (prev != B2ME_KEY_EMPTY) {
                s = (s + 1) & mask;
                continue;
            }
            prev = atomicCAS(&tab[s].key, B2ME_KEY_EMPTY, key);
            if (prev != B2ME_KEY_EMPTY && prev != key) {
                s = (s + 1) & mask;
                continue;
            }
        }
        break;
    }
    atomicMin(&tab[s].val, (unsigned int)i);
    slot_of[i] = (int32_t)s;
}

__global__ void k_first_flags(const HashSlot* __restrict__ tab, const int32_t* __restrict__ slot_of, int64_t n,
                              int32_t* __restrict__ rank) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t s = slot_of[i];
    rank[i] = (s >= 0 && tab[s].val == (unsigned int)i) ? 1 : 0;
}

// winners publish their row: table val <- row, coords, first index, (mode 0) features
__global__ void k_assign_rows(const int4* __restrict__ q, const int32_t* __restrict__ slot_of,
                              const int32_t* __restrict__ rank, const int32_t* __restrict__ counts, int64_t n,
                              HashSlot* __restrict__ tab, int4* __restrict__ out_coords,
                              int32_t* __restrict__ first_idx, const float* __restrict__ feats, int C, int mode,
                              float* __restrict__ out_feats) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t r = rank[i];
    const int32_t next = (i + 1 < n) ? rank[i + 1] : counts[0];
    if (next == r) return;  // not the first point of its voxel
    tab[slot_of[i]].val = (unsigned int)r;
    out_coords[r] = q[i];
    if (first_idx) first_idx[r] = (int32_t)i;
    if (mode == 0 && C > 0) {
        for (int c = 0; c < C; ++c) out_feats[(int64_t)r * C + c] = feats[i * C + c];
    }
}

#define FIXED_ONE 4294967296.0  // 2^32

__global__ void k_inverse_accumulate(const HashSlot* __restrict__ tab, const int32_t* __restrict__ slot_of,
                                     int64_t n, int32_t* __restrict__ inverse, const float* __restrict__ feats,
                                     int C, int mode, unsigned long long* __restrict__ fsum,
                                     int32_t* __restrict__ cnt, int32_t* __restrict__ counts) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t s = slot_of[i];
    if (s < 0) {
        inverse[i] = -1;
        return;
    }
    const int32_t row = (int32_t)tab[s].val;
    inverse[i] = row;
    if (mode == 1 && C > 0) {
        bool bad = false;
        for (int c = 0; c < C; ++c) {
            const float f = feats[i * C + c];
            if (!(fabsf(f) < 1048576.f)) bad = true;  // also catches NaN
            const long long fx = __double2ll_rn((double)f * FIXED_ONE);
            atomicAdd(fsum + (int64_t)row * C + c, (unsigned long long)fx);
        }
        atomicAdd(cnt + row, 1);
        if (bad) atomicOr(reinterpret_cast<unsigned int*>(counts + 1), 2u);
    }
}

__global__ void k_mean_features(const unsigned long long* __restrict__ fsum, const int32_t* __restrict__ cnt,
                                const int32_t* __restrict__ counts, int C, float* __restrict__ out_feats) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t total = (int64_t)counts[0] * C;
    if (t >= total) return;
    const int64_t row = t / C;
    const double s = (double)(long long)fsum[t];
    out_feats[t] = (float)((s / FIXED_ONE) / (double)cnt[row]);
}

extern "C" int64_t b2me_table_slots(int64_t n) {
    int64_t s = 1024;
    while (s < 2 * n) s <<= 1;
    return s;
}
extern "C" size_t b2me_table_bytes(int64_t n) { return (size_t)b2me_table_slots(n) * sizeof(HashSlot); }

struct UniqueWs {
    int4* q;
    int32_t* slot_of;
    int32_t* rank;
    void* scan;
    unsigned long long* fsum;
    int32_t* cnt;
    size_t total;
};

static UniqueWs carve_unique_ws(void* ws, int64_t n, int C) {
    UniqueWs w;
    size_t off = 0;
    char* base = reinterpret_cast<char*>(ws);
    const int64_t n1 = n > 0 ? n : 1;
    w.q = reinterpret_cast<int4*>(base + off);
    off += align_up((size_t)n1 * sizeof(int4), 256);
    w.slot_of = reinterpret_cast<int32_t*>(base + off);
    off += align_up((size_t)n1 * sizeof(int32_t), 256);
    w.rank = reinterpret_cast<int32_t*>(base + off);
    off += align_up((size_t)n1 * sizeof(int32_t), 256);
    w.scan = base + off;
    off += scan_ws_bytes(n1);
    w.fsum = reinterpret_cast<unsigned long long*>(base + off);
    off += align_up((size_t)n1 * (size_t)(C > 0 ? C : 1) * sizeof(unsigned long long), 256);
    w.cnt = reinterpret_cast<int32_t*>(base + off);
    off += align_up((size_t)n1 * sizeof(int32_t), 256);
    w.total = off;
    return w;
}

extern "C" size_t b2me_unique_workspace_bytes(int64_t n, int C) {
    return carve_unique_ws(nullptr, n, C).total;
}

// common tail: q (int4 rows) -> table, rows, inverse
static int unique_core(const int4* q, int64_t N, const float* feats, int C, int mode, int32_t* out_coords,
                       float* out_feats, int32_t* inverse, int32_t* first_idx, int32_t* counts, void* table,
                       size_t table_bytes, const UniqueWs& w, cudaStream_t s) {
    const int64_t slots = b2me_table_slots(N);
    if (table_bytes < (size_t)slots * sizeof(HashSlot)) return B2ME_EWORKSPACE;
    HashSlot* tab = reinterpret_cast<HashSlot*>(table);
    cudaMemsetAsync(tab, 0xFF, (size_t)slots * sizeof(HashSlot), s);
    cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), s);
    if (N == 0) return B2ME_OK;
    const int T = 256;
    const unsigned G = (unsigned)ceil_div64(N, T);
    k_hash_insert<<<G, T, 0, s>>>(q, N, tab, (unsigned long long)(slots - 1), w.slot_of, counts);
    k_first_flags<<<G, T, 0, s>>>(tab, w.slot_of, N, w.rank);
    int rc = exclusive_scan_i32(w.rank, N, counts, w.scan, s);
    if (rc != B2ME_OK) return rc;
    k_assign_rows<<<G, T, 0, s>>>(q, w.slot_of, w.rank, counts, N, tab, reinterpret_cast<int4*>(out_coords),
                                  first_idx, feats, C, mode, out_feats);
    if (mode == 1 && C > 0) {
        cudaMemsetAsync(w.fsum, 0, (size_t)N * C * sizeof(unsigned long long), s);
        cudaMemsetAsync(w.cnt, 0, (size_t)N * sizeof(int32_t), s);
    }
    k_inverse_accumulate<<<G, T, 0, s>>>(tab, w.slot_of, N, inverse, feats, C, mode, w.fsum, w.cnt, counts);
    if (mode == 1 && C > 0) {
        const int64_t total = N * C;  // upper bound on V*C; kernel bounds itself with counts[0]
        k_mean_features<<<(unsigned)ceil_div64(total, T), T, 0, s>>>(w.fsum, w.cnt, counts, C, out_feats);
    }
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

extern "C" int b2me_quantize_unique(const float* coords_f, const int32_t* coords_i, int64_t N, const float* feats,
                                    int C, int mode, int32_t* out_coords, float* out_feats, int32_t* inverse,
                                    int32_t* first_idx, int32_t* counts, void* table, size_t table_bytes, void* ws,
                                    size_t ws_bytes, b2me_stream_t stream) {
    if (N < 0 || C < 0 || (mode != 0 && mode != 1)) return B2ME_EINVAL;
    if ((coords_f == nullptr) == (coords_i == nullptr) && N > 0) return B2ME_EINVAL;
    if (!out_coords || !inverse || !counts || !table || !ws) return B2ME_EINVAL;
    if (C > 0 && (!feats || !out_feats)) return B2ME_EINVAL;
    if (N >= (int64_t)1 << 31) return B2ME_EINVAL;
    UniqueWs w = carve_unique_ws(ws, N, C);
    if (ws_bytes < w.total) return B2ME_EWORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const int4* q = reinterpret_cast<const int4*>(coords_i);
    if (coords_f && N > 0) {
        k_quantize_float<<<(unsigned)ceil_div64(N, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(coords_f), N,
                                                                      w.q);
        q = w.q;
    }
    return unique_core(q, N, feats, C, mode, out_coords, out_feats, inverse, first_idx, counts, table, table_bytes, w,
                       s);
}

__global__ void k_quantize_labels(const int32_t* __restrict__ labels, const int32_t* __restrict__ inverse,
                                  const int32_t* __restrict__ first_idx, int64_t N, int32_t ignore_label,
                                  int32_t* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int32_t row = inverse[i];
    if (row < 0) return;
    const int32_t mine = labels[i];
    const int32_t ref = labels[first_idx[row]];
    if (mine != ref) out[row] = ignore_label;  // all writers store the same value
}
__global__ void k_init_labels(const int32_t* __restrict__ labels, const int32_t* __restrict__ first_idx, int64_t V,
                              int32_t* __restrict__ out) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v < V) out[v] = labels[first_idx[v]];
}

extern "C" int b2me_quantize_labels(const int32_t* labels, const int32_t* inverse, const int32_t* first_idx,
                                    int64_t N, int64_t V, int32_t ignore_label, int32_t* out_labels,
                                    b2me_stream_t stream) {
    if (!labels || !inverse || !first_idx || !out_labels || N < 0 || V < 0) return B2ME_EINVAL;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (V > 0) k_init_labels<<<(unsigned)ceil_div64(V, 256), 256, 0, s>>>(labels, first_idx, V, out_labels);
    if (N > 0)
        k_quantize_labels<<<(unsigned)ceil_div64(N, 256), 256, 0, s>>>(labels, inverse, first_idx, N, ignore_label,
                                                                       out_labels);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ K2
extern "C" int b2me_stride_map(const int32_t* in_coords, int64_t V_in, int ts_out, int32_t* out_coords,
                               int32_t* in2out, uint8_t* koff, int32_t* counts, void* table_out, size_t table_bytes,
                               void* ws, size_t ws_bytes, b2me_stream_t stream) {
    if (V_in < 0 || ts_out < 2 || (ts_out & 1)) return B2ME_EINVAL;
    if (!in_coords || !out_coords || !in2out || !koff || !counts || !table_out || !ws) return B2ME_EINVAL;
    UniqueWs w = carve_unique_ws(ws, V_in, 0);
    if (ws_bytes < w.total) return B2ME_EWORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (V_in > 0)
        k_quantize_stride<<<(unsigned)ceil_div64(V_in, 256), 256, 0, s>>>(reinterpret_cast<const int4*>(in_coords),
                                                                          V_in, ts_out, w.q, koff);
    return unique_core(w.q, V_in, nullptr, 0, 0, out_coords, nullptr, in2out, nullptr, counts, table_out, table_bytes,
                       w, s);
}

__global__ void k_stride_kernel_maps(const int32_t* __restrict__ in2out, const uint8_t* __restrict__ koff,
                                     int64_t V_in, int32_t* __restrict__ nbr_down, int32_t* __restrict__ nbr_up) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= V_in) return;
    const int32_t p = in2out[i];
    const int k = koff[i];
    if (p < 0) return;
    if (nbr_down) nbr_down[(int64_t)p * 8 + k] = (int32_t)i;  // (parent, offset) has exactly one child
    if (nbr_up) nbr_up[i * 8 + k] = p;
}

extern "C" int b2me_stride_kernel_maps(const int32_t* in2out, const uint8_t* koff, int64_t V_in, int64_t V_out,
                                       int32_t* nbr_down, int32_t* nbr_up, b2me_stream_t stream) {
    if (!in2out || !koff || V_in < 0 || V_out < 0) return B2ME_EINVAL;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (nbr_down && V_out > 0) cudaMemsetAsync(nbr_down, 0xFF, (size_t)V_out * 8 * sizeof(int32_t), s);
    if (nbr_up && V_in > 0) cudaMemsetAsync(nbr_up, 0xFF, (size_t)V_in * 8 * sizeof(int32_t), s);
    if (V_in > 0)
        k_stride_kernel_maps<<<(unsigned)ceil_div64(V_in, 256), 256, 0, s>>>(in2out, koff, V_in, nbr_down, nbr_up);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ K3
// one thread per (voxel, offset): consecutive threads write consecutive nbr entries (full 128-B lines
// per warp); the 27 threads of a voxel share its coordinate row through L1.
__global__ void __launch_bounds__(256) k_kernel_map_k3(const int4* __restrict__ coords, int64_t V, int ts,
                                                       const HashSlot* __restrict__ tab, unsigned long long mask,
                                                       int32_t* __restrict__ nbr, uint32_t* __restrict__ tile_mask) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= V * 27) return;
    const int64_t v = t / 27;
    const int k = (int)(t - v * 27);
    int32_t res;
    if (k == 13) {
        res = (int32_t)v;
    } else {
        const int4 c = __ldg(coords + v);
        const int dx = k % 3 - 1, dy = (k / 3) % 3 - 1, dz = k / 9 - 1;
        const int x = c.y + dx * ts, y = c.z + dy * ts, z = c.w + dz * ts;
        res = -1;
        if (coord_in_range(c.x, x, y, z)) res = (int32_t)table_lookup(tab, mask, pack_key(c.x, x, y, z));
    }
    nbr[t] = res;
    if (tile_mask && res >= 0) {
        const uint32_t bit = 1u << k;
        volatile uint32_t* m = tile_mask + (v >> 7);
        if (!(*m & bit)) atomicOr(tile_mask + (v >> 7), bit);
    }
}

extern "C" int b2me_kernel_map_k3(const int32_t* coords, int64_t V, int ts, const void* table, size_t table_bytes,
                                  int32_t* nbr, uint32_t* tile_mask, b2me_stream_t stream) {
    if (!coords || !table || !nbr || V < 0 || ts < 1) return B2ME_EINVAL;
    const int64_t slots = (int64_t)(table_bytes / sizeof(HashSlot));
    if (slots < 8 || (slots & (slots - 1))) return B2ME_EINVAL;  // hash_slot: 8-slot groups
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (V == 0) return B2ME_OK;
    if (tile_mask) cudaMemsetAsync(tile_mask, 0, (size_t)ceil_div64(V, 128) * sizeof(uint32_t), s);
    const int64_t total = V * 27;
    k_kernel_map_k3<<<(unsigned)ceil_div64(total, 256), 256, 0, s>>>(reinterpret_cast<const int4*>(coords), V, ts,
                                                                     reinterpret_cast<const HashSlot*>(table),
                                                                     (unsigned long long)(slots - 1), nbr, tile_mask);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ---- K3 through 4 x 4 x 4 blocks (large maps). The coordinate hierarchy of the network already holds the map at
// tensor stride 4 ts (two stride-2 levels up): its rows ARE the occupied 4 x 4 x 4 blocks of this level and its hash
// table (a few MB: L2-resident) maps a block origin to its row. `brows[B, 64]` = the voxel of every cell of block B.
// A voxel's 27 neighbours lie in its own block and in at most one other block per axis, so a thread probes the
// (small) block table 1..8 times (3.4 on average) instead of the (large) voxel table 26 times, and reads its
// neighbours out of 256-byte block arrays that the threads next to it (scan-line order) read too: cache hits instead
// of 26 random DRAM sectors. The 256 x 27 results of a CTA are staged in shared memory and leave as one contiguous
// 27.6 KB store; the row's occupancy mask (27 bits) and the per-offset neighbour counts of the map (inputs of the
// K3b sort keys / tile masks) come out of the same pass.
__global__ void k_block_rows(const int4* __restrict__ coords, int64_t V, int shift, const int32_t* __restrict__ in2out1,
                             const int32_t* __restrict__ in2out2, int32_t* __restrict__ brows) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int4 c = __ldg(coords + v);
    const int32_t p1 = in2out1[v];
    if (p1 < 0) return;
    const int32_t blk = in2out2[p1];
    if (blk < 0) return;
    const int l = ((c.y >> shift) & 3) | (((c.z >> shift) & 3) << 2) | (((c.w >> shift) & 3) << 4);
    brows[(int64_t)blk * 64 + l] = (int32_t)v;
}

extern "C" int b2me_block_rows(const int32_t* coords, int64_t V, int ts, const int32_t* in2out1,
                               const int32_t* in2out2, int64_t V_blocks, int32_t* brows, b2me_stream_t stream) {
    if (!coords || !in2out1 || !in2out2 || !brows || V < 0 || V_blocks < 0 || ts < 1 || (ts & (ts - 1))) return B2ME_EINVAL;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (V_blocks > 0) cudaMemsetAsync(brows, 0xFF, (size_t)V_blocks * 64 * sizeof(int32_t), s);
    if (V == 0) return B2ME_OK;
    int shift = 0;
    while ((1 << shift) < ts) ++shift;
    k_block_rows<<<(unsigned)ceil_div64(V, 256), 256, 0, s>>>(reinterpret_cast<const int4*>(coords), V, shift, in2out1,
                                                             in2out2, brows);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

#define K3B_THREADS 256
__global__ void __launch_bounds__(K3B_THREADS)
k_kernel_map_k3_blocks(const int4* __restrict__ coords, int64_t V, int ts, int shift, const HashSlot* __restrict__ btab,
                       unsigned long long bmask, const int32_t* __restrict__ brows, int32_t* __restrict__ nbr,
                       uint32_t* __restrict__ row_masks, unsigned int* __restrict__ offset_counts) {
    __shared__ int32_t res_s[K3B_THREADS * 27];   // [thread][offset]: stride 27 words, conflict-free
    __shared__ uint32_t brow_s[8][K3B_THREADS];   // block row of the (x, y, z)-crossing combination, per thread
    __shared__ unsigned int cnt_s[27];
    const int tid = threadIdx.x;
    if (tid < 27) cnt_s[tid] = 0u;
    __syncthreads();
    const int64_t v0 = (int64_t)blockIdx.x * K3B_THREADS;
    const int64_t v = v0 + tid;
    uint32_t m = 0u;
    if (v < V) {
        const int4 c = __ldg(coords + v);
        const int bs = 4 * ts;
        const int o[3] = {c.y & ~(bs - 1), c.z & ~(bs - 1), c.w & ~(bs - 1)};     // block origin (floor to 4 ts)
        const int li[3] = {(c.y >> shift) & 3, (c.z >> shift) & 3, (c.w >> shift) & 3};
        int dir[3];                                                                // the one other block an axis can reach
#pragma unroll
        for (int a = 0; a < 3; ++a) dir[a] = li[a] == 0 ? -1 : (li[a] == 3 ? 1 : 0);
#pragma unroll
        for (int combo = 0; combo < 8; ++combo) {
            const int bx = combo & 1, by = (combo >> 1) & 1, bz = combo >> 2;
            uint32_t r = 0xFFFFFFFFu;
            if ((!bx || dir[0]) && (!by || dir[1]) && (!bz || dir[2])) {
                const int x = o[0] + bx * dir[0] * bs, y = o[1] + by * dir[1] * bs, z = o[2] + bz * dir[2] * bs;
                if (coord_in_range(c.x, x, y, z)) r = table_lookup(btab, bmask, pack_key(c.x, x, y, z));
            }
            brow_s[combo][tid] = r;
        }
        // per (dy, dz): the 4-cell x-row of the block that holds the neighbours' y / z (ONE 16-byte load), plus one cell
        // of the x-adjacent block when the voxel sits on an x face of its block: 9..18 loads instead of 26
        const int lx = li[0];
#pragma unroll
        for (int dz = -1; dz <= 1; ++dz) {
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy) {
                const int ny = li[1] + dy, nz = li[2] + dz;
                const int cyz = (((ny < 0 || ny > 3) ? 1 : 0) << 1) | (((nz < 0 || nz > 3) ? 1 : 0) << 2);
                const int row = ((ny & 3) << 2) | ((nz & 3) << 4);
                const uint32_t B0 = brow_s[cyz][tid];
                int4 cells = make_int4(-1, -1, -1, -1);
                if (B0 != 0xFFFFFFFFu) cells = __ldg(reinterpret_cast<const int4*>(brows + (int64_t)B0 * 64 + row));
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const int k = (dx + 1) + 3 * (dy + 1) + 9 * (dz + 1);
                    const int nx = lx + dx;
                    int32_t r;
                    if (k == 13) {
                        r = (int32_t)v;
                    } else if (nx >= 0 && nx <= 3) {
                        r = nx == 0 ? cells.x : (nx == 1 ? cells.y : (nx == 2 ? cells.z : cells.w));
                    } else {
                        const uint32_t B1 = brow_s[cyz | 1][tid];
                        r = (B1 == 0xFFFFFFFFu) ? -1 : __ldg(brows + (int64_t)B1 * 64 + row + (nx & 3));
                    }
                    res_s[tid * 27 + k] = r;
                    m |= (r >= 0 ? 1u : 0u) << k;
                }
            }
        }
        row_masks[v] = m;
    }
    // per-offset neighbour counts of the map: one ballot per offset and warp
#pragma unroll
    for (int k = 0; k < 27; ++k) {
        const unsigned int b = __ballot_sync(0xffffffffu, (m >> k) & 1u);
        if ((tid & 31) == 0 && b) atomicAdd(&cnt_s[k], (unsigned int)__popc(b));
    }
    __syncthreads();
    const int64_t nvalid = (V - v0) < K3B_THREADS ? (V - v0) : K3B_THREADS;
    int32_t* dst = nbr + v0 * 27;
    for (int i = tid; i < (int)nvalid * 27; i += K3B_THREADS) dst[i] = res_s[i];
    if (tid < 27 && cnt_s[tid]) atomicAdd(offset_counts + tid, cnt_s[tid]);
}

// row masks + per-offset counts from an existing nbr table (small maps built by the direct kernel; K = 8 maps)
__global__ void __launch_bounds__(256) k_row_masks(const int32_t* __restrict__ nbr, int64_t V, int K,
                                                   uint32_t* __restrict__ row_masks,
                                                   unsigned int* __restrict__ offset_counts) {
    __shared__ unsigned int cnt_s[32];
    if (threadIdx.x < 32) cnt_s[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t m = 0u;
    if (v < V) {
        for (int k = 0; k < K; ++k)
            if (__ldg(nbr + v * K + k) >= 0) m |= 1u << k;
        row_masks[v] = m;
    }
    for (int k = 0; k < K; ++k) {
        const unsigned int b = __ballot_sync(0xffffffffu, (m >> k) & 1u);
        if ((threadIdx.x & 31) == 0 && b) atomicAdd(&cnt_s[k], (unsigned int)__popc(b));
    }
    __syncthreads();
    if (threadIdx.x < K && cnt_s[threadIdx.x]) atomicAdd(offset_counts + threadIdx.x, cnt_s[threadIdx.x]);
}

extern "C" int b2me_row_masks(const int32_t* nbr, int64_t V, int K, uint32_t* row_masks, uint32_t* offset_counts,
                              b2me_stream_t stream) {
    if (!nbr || !row_masks || !offset_counts || V < 0 || K < 1 || K > 31) return B2ME_EINVAL;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    cudaMemsetAsync(offset_counts, 0, 32 * sizeof(uint32_t), s);
    if (V == 0) return B2ME_OK;
    k_row_masks<<<(unsigned)ceil_div64(V, 256), 256, 0, s>>>(nbr, V, K, row_masks, offset_counts);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

extern "C" int b2me_kernel_map_k3_blocks(const int32_t* coords, int64_t V, int ts, const void* block_table,
                                         size_t block_table_bytes, const int32_t* brows, int32_t* nbr,
                                         uint32_t* row_masks, uint32_t* offset_counts, b2me_stream_t stream) {
    if (!coords || !block_table || !brows || !nbr || !row_masks || !offset_counts || V < 0 || ts < 1 || (ts & (ts - 1)))
        return B2ME_EINVAL;
    if (ts > (1 << 14)) return B2ME_EINVAL;
    const int64_t slots = (int64_t)(block_table_bytes / sizeof(HashSlot));
    if (slots < 8 || (slots & (slots - 1))) return B2ME_EINVAL;  // hash_slot: 8-slot groups
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    cudaMemsetAsync(offset_counts, 0, 32 * sizeof(uint32_t), s);
    if (V == 0) return B2ME_OK;
    int shift = 0;
    while ((1 << shift) < ts) ++shift;
    k_kernel_map_k3_blocks<<<(unsigned)ceil_div64(V, K3B_THREADS), K3B_THREADS, 0, s>>>(
        reinterpret_cast<const int4*>(coords), V, ts, shift, reinterpret_cast<const HashSlot*>(block_table),
        (unsigned long long)(slots - 1), brows, nbr, row_masks, offset_counts);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ K3b
// Sort keys for the convolution's row permutation: rows with the same neighbour pattern are made adjacent so
// that a 128-row tile of the gather-GEMM kernel needs few kernel offsets (the tile skips an offset when none of
// its rows has that neighbour). key(row) = the row's K-bit occupancy mask with the bits reordered so that the
// RAREST offset of this map is the most significant bit (measured: 0.94 -> 0.37 non-empty (tile, offset) pairs
// on 5 mm Kinect clouds, fill 0.25).
// Reflected (Gray-code) order: bit i of the sort key = bit i of the mask XOR the parity of the mask bits above it, so
// that consecutive keys differ in few mask bits wherever a more significant bit flips (measured on 5 mm Kinect
// clouds: 10.76 -> 10.46 offsets per 256-row tile pair, -3 % MMA passes).
__device__ __forceinline__ unsigned int reflect_key(unsigned int m) {
#ifdef B2ME_NO_REFLECT
    return m;  // A/B switch: plain lexicographic mask order
#endif
    unsigned int p = m >> 1;  // p bit i = parity of m bits above i
    p ^= p >> 1;
    p ^= p >> 2;
    p ^= p >> 4;
    p ^= p >> 8;
    p ^= p >> 16;
    return m ^ p;
}

__global__ void __launch_bounds__(256) k_offset_counts(const int32_t* __restrict__ nbr, int64_t total, int K,
                                                       unsigned int* __restrict__ counts /*[32]*/) {
    __shared__ unsigned int h[32];
    if (threadIdx.x < 32) h[threadIdx.x] = 0u;
    __syncthreads();
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x)
        if (__ldg(nbr + t) >= 0) atomicAdd(&h[(int)(t % K)], 1u);
    __syncthreads();
    if (threadIdx.x < K && h[threadIdx.x]) atomicAdd(&counts[threadIdx.x], h[threadIdx.x]);
}

__global__ void __launch_bounds__(256) k_mask_keys(const int32_t* __restrict__ nbr, int64_t V, int K,
                                                   const unsigned int* __restrict__ counts,
                                                   int32_t* __restrict__ keys) {
    __shared__ int bitpos[32];  // offset k -> bit position in the key
    if (threadIdx.x < K) {
        // rank of offset k by (count ascending, k ascending); rarest -> most significant of the K bits
        const unsigned int mine = counts[threadIdx.x];
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const unsigned int c = counts[j];
            if (c < mine || (c == mine && j < (int)threadIdx.x)) ++rank;
        }
        bitpos[threadIdx.x] = K - 1 - rank;
    }
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    unsigned int key = 0u;
    for (int k = 0; k < K; ++k)
        if (__ldg(nbr + v * K + k) >= 0) key |= 1u << bitpos[k];
    keys[v] = (int32_t)reflect_key(key);
}

// 64-bit variant: key = (row / block_rows) << 32 | mask key. Sorting these keeps the rows of one block of
// block_rows consecutive rows (= one spatial neighbourhood in first-occurrence order) together, so the tiles that
// run concurrently gather from one L2-sized window of the input instead of the whole tensor.
__global__ void __launch_bounds__(256) k_mask_keys64(const int32_t* __restrict__ nbr, int64_t V, int K,
                                                     const unsigned int* __restrict__ counts, int block_rows,
                                                     long long* __restrict__ keys) {
    __shared__ int bitpos[32];
    if (threadIdx.x < K) {
        const unsigned int mine = counts[threadIdx.x];
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const unsigned int c = counts[j];
            if (c < mine || (c == mine && j < (int)threadIdx.x)) ++rank;
        }
        bitpos[threadIdx.x] = K - 1 - rank;
    }
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    unsigned int key = 0u;
    for (int k = 0; k < K; ++k)
        if (__ldg(nbr + v * K + k) >= 0) key |= 1u << bitpos[k];
    const long long blk = block_rows > 0 ? (long long)(v / block_rows) : 0ll;
    keys[v] = (blk << 32) | (long long)reflect_key(key);
}

extern "C" int b2me_mask_sort_keys64(const int32_t* nbr, int64_t V, int K, int block_rows, int64_t* keys, void* ws,
                                     size_t ws_bytes, b2me_stream_t stream) {
    if (!nbr || !keys || !ws || V < 0 || K < 1 || K > 31 || block_rows < 0) return B2ME_EINVAL;
    if (ws_bytes < 32 * sizeof(unsigned int)) return B2ME_EWORKSPACE;
    if (V == 0) return B2ME_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned int* counts = reinterpret_cast<unsigned int*>(ws);
    cudaMemsetAsync(counts, 0, 32 * sizeof(unsigned int), s);
    const int64_t total = V * K;
    int64_t blocks = ceil_div64(total, 256 * 8);
    if (blocks > B2ME_NUM_SMS * 16) blocks = B2ME_NUM_SMS * 16;
    k_offset_counts<<<(unsigned)blocks, 256, 0, s>>>(nbr, total, K, counts);
    k_mask_keys64<<<(unsigned)ceil_div64(V, 256), 256, 0, s>>>(nbr, V, K, counts, block_rows,
                                                               reinterpret_cast<long long*>(keys));
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

extern "C" int b2me_mask_sort_keys(const int32_t* nbr, int64_t V, int K, int32_t* keys, void* ws, size_t ws_bytes,
                                   b2me_stream_t stream) {
    if (!nbr || !keys || !ws || V < 0 || K < 1 || K > 31) return B2ME_EINVAL;
    if (ws_bytes < 32 * sizeof(unsigned int)) return B2ME_EWORKSPACE;
    if (V == 0) return B2ME_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned int* counts = reinterpret_cast<unsigned int*>(ws);
    cudaMemsetAsync(counts, 0, 32 * sizeof(unsigned int), s);
    const int64_t total = V * K;
    int64_t blocks = ceil_div64(total, 256 * 8);
    if (blocks > B2ME_NUM_SMS * 16) blocks = B2ME_NUM_SMS * 16;
    k_offset_counts<<<(unsigned)blocks, 256, 0, s>>>(nbr, total, K, counts);
    k_mask_keys<<<(unsigned)ceil_div64(V, 256), 256, 0, s>>>(nbr, V, K, counts, keys);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// the same keys from the per-row occupancy masks + per-offset counts that the kernel-map pass already produced
// (4 bytes per row instead of 4 K: the nbr table is not read again)
__global__ void __launch_bounds__(256) k_mask_keys_rows(const uint32_t* __restrict__ row_masks, int64_t V, int K,
                                                        const unsigned int* __restrict__ counts,
                                                        int32_t* __restrict__ keys) {
    __shared__ int bitpos[32];
    if (threadIdx.x < K) {
        const unsigned int mine = counts[threadIdx.x];
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const unsigned int c = counts[j];
            if (c < mine || (c == mine && j < (int)threadIdx.x)) ++rank;
        }
        bitpos[threadIdx.x] = K - 1 - rank;
    }
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const uint32_t m = __ldg(row_masks + v);
    unsigned int key = 0u;
    for (int k = 0; k < K; ++k)
        if ((m >> k) & 1u) key |= 1u << bitpos[k];
    keys[v] = (int32_t)reflect_key(key);
}

extern "C" int b2me_mask_sort_keys_rows(const uint32_t* row_masks, const uint32_t* offset_counts, int64_t V, int K,
                                        int32_t* keys, b2me_stream_t stream) {
    if (!row_masks || !offset_counts || !keys || V < 0 || K < 1 || K > 31) return B2ME_EINVAL;
    if (V == 0) return B2ME_OK;
    k_mask_keys_rows<<<(unsigned)ceil_div64(V, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        row_masks, V, K, offset_counts, keys);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// per 256-row tile pair of the permuted order: OR of the row masks (one thread per row, warp OR, 8 warps per pair)
__global__ void __launch_bounds__(256) k_tile_masks_rows(const uint32_t* __restrict__ row_masks,
                                                         const int32_t* __restrict__ perm, int64_t V,
                                                         uint32_t* __restrict__ masks) {
    __shared__ uint32_t part[8];
    const int64_t slot = (int64_t)blockIdx.x * 256 + threadIdx.x;
    uint32_t m = 0u;
    if (slot < V) m = __ldg(row_masks + (perm ? (int64_t)__ldg(perm + slot) : slot));
    m = __reduce_or_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) masks[blockIdx.x] = part[0] | part[1] | part[2] | part[3] | part[4] | part[5] | part[6] | part[7];
}

extern "C" int b2me_tile_masks_rows(const uint32_t* row_masks, const int32_t* perm, int64_t V, uint32_t* masks,
                                    b2me_stream_t stream) {
    if (!row_masks || !masks || V < 0) return B2ME_EINVAL;
    if (V == 0) return B2ME_OK;
    k_tile_masks_rows<<<(unsigned)ceil_div64(V, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(row_masks, perm,
                                                                                                       V, masks);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// Mask keys with a spatial tie-break: key = mask key << 37 | frame << 30 | Morton code of the voxel (10 bits per axis
// of coord / ts, wrapped). Rows with the same neighbour pattern then follow a space-filling curve instead of the
// first-occurrence (depth-image raster) order, so the rows of a tile - and of the tiles that run concurrently - form
// compact 3-D patches whose gathered neighbours overlap across kernel offsets: the re-gathers hit L2 instead of DRAM.
__device__ __forceinline__ unsigned int morton_spread10(unsigned int v) {  // 10 bits -> every third bit
    v &= 0x3FFu;
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void __launch_bounds__(256) k_mask_keys_morton(const int32_t* __restrict__ nbr, const int4* __restrict__ coords,
                                                          int64_t V, int K, int ts_shift,
                                                          const unsigned int* __restrict__ counts,
                                                          long long* __restrict__ keys) {
    __shared__ int bitpos[32];
    if (threadIdx.x < K) {
        const unsigned int mine = counts[threadIdx.x];
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const unsigned int c = counts[j];
            if (c < mine || (c == mine && j < (int)threadIdx.x)) ++rank;
        }
        bitpos[threadIdx.x] = K - 1 - rank;
    }
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    unsigned int key = 0u;
    for (int k = 0; k < K; ++k)
        if (__ldg(nbr + v * K + k) >= 0) key |= 1u << bitpos[k];
    const int4 c = __ldg(coords + v);  // (b, x, y, z)
    const unsigned int m = morton_spread10((unsigned int)(c.y >> ts_shift)) |
                           (morton_spread10((unsigned int)(c.z >> ts_shift)) << 1) |
                           (morton_spread10((unsigned int)(c.w >> ts_shift)) << 2);
    keys[v] = ((long long)reflect_key(key) << 37) | ((long long)(c.x & 0x7F) << 30) | (long long)m;
}

extern "C" int b2me_mask_sort_keys_morton(const int32_t* nbr, const int32_t* coords, int64_t V, int K, int ts,
                                          int64_t* keys, void* ws, size_t ws_bytes, b2me_stream_t stream) {
    if (!nbr || !coords || !keys || !ws || V < 0 || K < 1 || K > 27 || ts < 1 || (ts & (ts - 1))) return B2ME_EINVAL;
    if (ws_bytes < 32 * sizeof(unsigned int)) return B2ME_EWORKSPACE;
    if (V == 0) return B2ME_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned int* counts = reinterpret_cast<unsigned int*>(ws);
    cudaMemsetAsync(counts, 0, 32 * sizeof(unsigned int), s);
    const int64_t total = V * K;
    int64_t blocks = ceil_div64(total, 256 * 8);
    if (blocks > B2ME_NUM_SMS * 16) blocks = B2ME_NUM_SMS * 16;
    int shift = 0;
    while ((1 << shift) < ts) ++shift;
    k_offset_counts<<<(unsigned)blocks, 256, 0, s>>>(nbr, total, K, counts);
    k_mask_keys_morton<<<(unsigned)ceil_div64(V, 256), 256, 0, s>>>(nbr, reinterpret_cast<const int4*>(coords), V, K, shift,
                                                                    counts, reinterpret_cast<long long*>(keys));
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// Two-level keys: the T rarest offsets of the map form the segment (most significant part of the key, rarest first,
// as above); the remaining R = K - T offsets are then ordered PER SEGMENT by their frequency among the rows of that
// segment (rarest first; offsets that no row or every row of the segment has go last: they cost nothing). The
// conditional order resolves more offsets exactly before the groups fall below a tile (measured on 5 mm Kinect clouds,
// 6 frames: 10.01 -> 9.48 offsets per 256-row tile pair; an exact recursive tree gives 9.28).
#define MS2_T 10
#define MS2_SEGS (1 << MS2_T)

// pass 1: raw masks + per-segment counts of every offset (+ segment size in column 31). Lanes of a warp that share a
// segment add through one leader (match_any), so the hot segments see one atomic per warp and offset, not per row.
__global__ void __launch_bounds__(256) k_ms2_segment_counts(const int32_t* __restrict__ nbr, int64_t V, int K,
                                                            const unsigned int* __restrict__ counts,
                                                            unsigned int* __restrict__ masks,
                                                            unsigned int* __restrict__ seg_counts /*[SEGS][32]*/) {
    __shared__ int bitpos[32];  // offset k -> bit position in the GLOBAL rarest-first key (K-1 = rarest)
    if (threadIdx.x < K) {
        const unsigned int mine = counts[threadIdx.x];
        int rank = 0;
        for (int j = 0; j < K; ++j) {
            const unsigned int c = counts[j];
            if (c < mine || (c == mine && j < (int)threadIdx.x)) ++rank;
        }
        bitpos[threadIdx.x] = K - 1 - rank;
    }
    __syncthreads();
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = v < V;
    unsigned int key = 0u;
    if (live) {
        for (int k = 0; k < K; ++k)
            if (__ldg(nbr + v * K + k) >= 0) key |= 1u << bitpos[k];
        masks[v] = key;  // bits already in global rarest-first order
    }
    const int R = K - MS2_T;
    const unsigned int seg = live ? (key >> R) : 0xFFFFFFFFu;
    const unsigned int peers = __match_any_sync(0xffffffffu, seg);
    const int leader = __ffs((int)peers) - 1;
    const int lane = threadIdx.x & 31;
    for (int b = 0; b < R; ++b) {
        const unsigned int has = __ballot_sync(0xffffffffu, live && ((key >> b) & 1u));
        if (live && lane == leader) {
            const int c = __popc(has & peers);
            if (c) atomicAdd(&seg_counts[seg * 32 + b], (unsigned int)c);
        }
    }
    if (live && lane == leader) atomicAdd(&seg_counts[seg * 32 + 31], (unsigned int)__popc(peers));
}

// pass 2 (one thread per segment): position of each low bit in the segment's own order -> shift table
__global__ void __launch_bounds__(256) k_ms2_segment_order(const unsigned int* __restrict__ seg_counts, int R,
                                                           unsigned char* __restrict__ shifts /*[SEGS][32]*/) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= MS2_SEGS) return;
    const unsigned int tot = seg_counts[s * 32 + 31];
    unsigned int c[32];
    for (int b = 0; b < R; ++b) {
        const unsigned int x = seg_counts[s * 32 + b];
        c[b] = (x == 0u || x == tot) ? 0xFFFFFFFFu : x;  // free offsets last
    }
    for (int b = 0; b < R; ++b) {
        // rank by (count ascending, global rarity descending = higher bit first): rarest -> most significant
        int rank = 0;
        for (int j = 0; j < R; ++j)
            if (c[j] < c[b] || (c[j] == c[b] && j > b)) ++rank;
        shifts[s * 32 + b] = (unsigned char)(R - 1 - rank);
    }
}

// pass 3: final key = reflect(segment) << R | reflect(low bits in the segment's order)
__global__ void __launch_bounds__(256) k_ms2_keys(const unsigned int* __restrict__ masks, int64_t V, int K,
                                                  const unsigned char* __restrict__ shifts,
                                                  int32_t* __restrict__ keys) {
    const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= V) return;
    const int R = K - MS2_T;
    const unsigned int key = masks[v];
    const unsigned int seg = key >> R;
    const uint4* sh4 = reinterpret_cast<const uint4*>(shifts + (size_t)seg * 32);
    const uint4 q0 = __ldg(sh4), q1 = __ldg(sh4 + 1);
    const unsigned int w[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    unsigned int low = 0u;
    for (int b = 0; b < R; ++b)
        if ((key >> b) & 1u) low |= 1u << ((w[b >> 2] >> (8 * (b & 3))) & 0xFFu);
    keys[v] = (int32_t)((reflect_key(seg) << R) | reflect_key(low));
}

extern "C" size_t b2me_mask_sort_keys2_ws_bytes(int64_t V) {
    return 128 + (size_t)MS2_SEGS * 32 * sizeof(unsigned int) + (size_t)MS2_SEGS * 32 + (size_t)(V > 0 ? V : 0) * 4 + 64;
}

extern "C" int b2me_mask_sort_keys2(const int32_t* nbr, int64_t V, int K, int32_t* keys, void* ws, size_t ws_bytes,
                                    b2me_stream_t stream) {
    if (!nbr || !keys || !ws || V < 0 || K < 1 || K > 31) return B2ME_EINVAL;
    if (K <= MS2_T + 1) return b2me_mask_sort_keys(nbr, V, K, keys, ws, ws_bytes, stream);  // too few offsets to split
    if (ws_bytes < b2me_mask_sort_keys2_ws_bytes(V)) return B2ME_EWORKSPACE;
    if (V == 0) return B2ME_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    uint8_t* w8 = reinterpret_cast<uint8_t*>(ws);
    unsigned int* counts = reinterpret_cast<unsigned int*>(w8);
    unsigned int* seg_counts = reinterpret_cast<unsigned int*>(w8 + 128);
    unsigned char* shifts = w8 + 128 + (size_t)MS2_SEGS * 32 * sizeof(unsigned int);
    unsigned int* masks = reinterpret_cast<unsigned int*>(shifts + (size_t)MS2_SEGS * 32);
    cudaMemsetAsync(w8, 0, 128 + (size_t)MS2_SEGS * 32 * sizeof(unsigned int), s);
    const int64_t total = V * K;
    int64_t blocks = ceil_div64(total, 256 * 8);
    if (blocks > B2ME_NUM_SMS * 16) blocks = B2ME_NUM_SMS * 16;
    k_offset_counts<<<(unsigned)blocks, 256, 0, s>>>(nbr, total, K, counts);
    k_ms2_segment_counts<<<(unsigned)ceil_div64(V, 256), 256, 0, s>>>(nbr, V, K, counts, masks, seg_counts);
    k_ms2_segment_order<<<MS2_SEGS / 256, 256, 0, s>>>(seg_counts, K - MS2_T, shifts);
    k_ms2_keys<<<(unsigned)ceil_div64(V, 256), 256, 0, s>>>(masks, V, K, shifts, keys);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ ingest
// Organised PointCloud2 / PCD records (x, y, z f32 + PCL-packed rgb: 0x00RRGGBB in the bits of a float) of a batch
// of frames -> the tensors K1 consumes, in one pass on the device instead of the reference's NumPy hops:
//   utils/ros_utils.py:142-167  drop non-finite points, split the packed rgb
//   app/freenect_data_engine.py:81  rgb / 255           utils/preprocess.py:20-37  rgb - 0.5
//   utils/data.py:58-75  ROI mask (strict inequalities)
// Order-preserving compaction (flags -> exclusive scan -> scatter), so first-occurrence voxel order is unchanged.
__global__ void k_ingest_flags(const float4* __restrict__ rec, int64_t n, float min_x, float max_x, float min_y,
                               float max_y, float min_z, float max_z, int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 r = rec[i];
    const bool ok = isfinite(r.x) && isfinite(r.y) && isfinite(r.z) && r.x > -500.f && r.x < max_x && r.x > min_x &&
                    r.y < max_y && r.y > min_y && r.z < max_z && r.z > min_z;
    flag[i] = ok ? 1 : 0;
}

__global__ void k_ingest_scatter(const float4* __restrict__ rec, int64_t n, const int32_t* __restrict__ pos,
                                 const int32_t* __restrict__ total, const int32_t* __restrict__ frame_offsets, int F,
                                 float* __restrict__ out_xyz, float* __restrict__ out_rgb, float* __restrict__ out_bidx,
                                 int32_t* __restrict__ out_src, int32_t* __restrict__ out_offsets) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= F) {  // compacted frame offsets: position of the frame's first input record
        const int32_t fo = frame_offsets[i];
        out_offsets[i] = fo < n ? pos[fo] : *total;
    }
    if (i >= n) return;
    const int32_t p = pos[i];
    const int32_t nxt = (i + 1 < n) ? pos[i + 1] : *total;
    if (nxt == p) return;  // dropped
    const float4 r = rec[i];
    int lo = 0, hi = F;    // frame of record i: largest f with frame_offsets[f] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (frame_offsets[mid] <= i) lo = mid;
        else hi = mid;
    }
    out_xyz[(int64_t)p * 3 + 0] = r.x;
    out_xyz[(int64_t)p * 3 + 1] = r.y;
    out_xyz[(int64_t)p * 3 + 2] = r.z;
    const unsigned int c = __float_as_uint(r.w);
    out_rgb[(int64_t)p * 3 + 0] = (float)((double)((c >> 16) & 255u) / 255.0 - 0.5);
    out_rgb[(int64_t)p * 3 + 1] = (float)((double)((c >> 8) & 255u) / 255.0 - 0.5);
    out_rgb[(int64_t)p * 3 + 2] = (float)((double)(c & 255u) / 255.0 - 0.5);
    out_bidx[p] = (float)lo;
    if (out_src) out_src[p] = (int32_t)i;
}

extern "C" size_t b2me_ingest_workspace_bytes(int64_t n) {
    const int64_t n1 = n > 0 ? n : 1;
    return align_up((size_t)(n1 + 1) * 4, 256) + align_up(16, 256) + scan_ws_bytes(n1 + 1);
}

extern "C" int b2me_ingest_clouds(const float* xyzrgb, int64_t n, const int32_t* frame_offsets, int F,
                                  const float* roi6, float* out_xyz, float* out_rgb, float* out_bidx, int32_t* out_src,
                                  int32_t* out_offsets, void* ws, size_t ws_bytes, b2me_stream_t stream) {
    if (!xyzrgb || !frame_offsets || !out_xyz || !out_rgb || !out_bidx || !out_offsets || !ws || n < 0 || F < 1)
        return B2ME_EINVAL;
    if (n >= (int64_t)1 << 31) return B2ME_EINVAL;
    if (ws_bytes < b2me_ingest_workspace_bytes(n)) return B2ME_EWORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    char* base = reinterpret_cast<char*>(ws);
    int32_t* pos = reinterpret_cast<int32_t*>(base);
    int32_t* total = reinterpret_cast<int32_t*>(base + align_up((size_t)(n + 1) * 4, 256));
    void* scan = base + align_up((size_t)(n + 1) * 4, 256) + align_up(16, 256);
    float r[6] = {-500.f, 500.f, -500.f, 500.f, -500.f, 500.f};  // utils/data.py:58 defaults
    if (roi6)
        for (int i = 0; i < 6; ++i) r[i] = roi6[i];  // host array
    const unsigned G = (unsigned)ceil_div64((n > F ? n : F) + 1, 256);
    if (n > 0) k_ingest_flags<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(reinterpret_cast<const float4*>(xyzrgb), n,
                                                                              r[0], r[1], r[2], r[3], r[4], r[5], pos);
    cudaMemsetAsync(pos + n, 0, 4, s);
    const int rc = exclusive_scan_i32(pos, n + 1, total, scan, s);
    if (rc != B2ME_OK) return rc;
    k_ingest_scatter<<<G, 256, 0, s>>>(reinterpret_cast<const float4*>(xyzrgb), n, pos, total, frame_offsets, F, out_xyz,
                                       out_rgb, out_bidx, out_src, out_offsets);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ colour normalisation
// utils/preprocess.py:20-37 (normalize_colors) per FRAME of a batch, on the device, without a host round trip:
//   max > 2          -> rgb / 255                       (float32 division, like the in-place NumPy op)
//   min < 0          -> per channel x * scale + min_, scale = 1 / (max - min), min_ = 0 - min * scale
//                       (sklearn.preprocessing.minmax_scale on a float32 column: two rounded float32 operations)
//   result in [0, 1] -> rgb - 0.5
// Pass 1 reduces the per-frame, per-channel extrema (block partials -> ordered-integer atomics), pass 2 applies the
// three steps with every decision taken from the frame's own extrema (division and the affine map are monotone, so
// the extrema of the intermediate arrays are the images of the input extrema under the very same float32 operations).
__device__ __forceinline__ unsigned int f32_to_ordered(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(unsigned int u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

__global__ void k_color_stats_init(unsigned int* __restrict__ stats, int F) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < F * 6) stats[i] = (i % 6) < 3 ? 0xFFFFFFFFu : 0u;  // [min r g b | max r g b], ordered encoding
}

#define COLOR_SLICES 16
__global__ void __launch_bounds__(256) k_color_stats(const float* __restrict__ rgb, const int32_t* __restrict__ offs,
                                                     unsigned int* __restrict__ stats) {
    const int f = blockIdx.y;
    const long long o0 = offs[f], o1 = offs[f + 1];
    const long long len = o1 - o0, per = (len + COLOR_SLICES - 1) / COLOR_SLICES;
    const long long a = o0 + per * blockIdx.x, b = (a + per < o1) ? a + per : o1;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = a + threadIdx.x; i < b; i += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = rgb[i * 3 + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
    if ((threadIdx.x & 31) == 0 && a < b) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            atomicMin(stats + f * 6 + c, f32_to_ordered(mn[c]));
            atomicMax(stats + f * 6 + 3 + c, f32_to_ordered(mx[c]));
        }
    }
}

__global__ void __launch_bounds__(256) k_color_apply(const float* __restrict__ rgb, const float* __restrict__ bidx,
                                                     int64_t n, int F, const unsigned int* __restrict__ stats,
                                                     float* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int f = (int)bidx[i];
    f = f < 0 ? 0 : (f >= F ? F - 1 : f);
    float mn[3], mx[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mn[c] = ordered_to_f32(__ldg(stats + f * 6 + c));
        mx[c] = ordered_to_f32(__ldg(stats + f * 6 + 3 + c));
    }
    float v[3] = {rgb[i * 3], rgb[i * 3 + 1], rgb[i * 3 + 2]};
    if (fmaxf(mx[0], fmaxf(mx[1], mx[2])) > 2.f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            v[c] = __fdiv_rn(v[c], 255.f);
            mn[c] = __fdiv_rn(mn[c], 255.f);
            mx[c] = __fdiv_rn(mx[c], 255.f);
        }
    }
    if (fminf(mn[0], fminf(mn[1], mn[2])) < 0.f) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float range = __fsub_rn(mx[c], mn[c]);
            const float scale = __fdiv_rn(1.f, range != 0.f ? range : 1.f);
            const float min_ = __fsub_rn(0.f, __fmul_rn(mn[c], scale));
            v[c] = __fadd_rn(__fmul_rn(v[c], scale), min_);   // two rounded operations, never an FMA
            const float lo = __fadd_rn(__fmul_rn(mn[c], scale), min_), hi = __fadd_rn(__fmul_rn(mx[c], scale), min_);
            mn[c] = lo;
            mx[c] = hi;
        }
    }
    if ((double)fminf(mn[0], fminf(mn[1], mn[2])) > -1e-6 && (double)fmaxf(mx[0], fmaxf(mx[1], mx[2])) < 1.0 + 1e-6) {
#pragma unroll
        for (int c = 0; c < 3; ++c) v[c] = __fsub_rn(v[c], 0.5f);
    }
    out[i * 3] = v[0];
    out[i * 3 + 1] = v[1];
    out[i * 3 + 2] = v[2];
}

extern "C" int b2me_normalize_colors(const float* rgb, const float* bidx, int64_t n, const int32_t* frame_offsets, int F,
                                     float* out, void* ws, size_t ws_bytes, b2me_stream_t stream) {
    if (!rgb || !bidx || !frame_offsets || !out || !ws || n < 0 || F < 1 || F > 65535) return B2ME_EINVAL;
    if (ws_bytes < (size_t)F * 6 * sizeof(unsigned int)) return B2ME_EWORKSPACE;
    if (n == 0) return B2ME_OK;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    unsigned int* stats = reinterpret_cast<unsigned int*>(ws);
    k_color_stats_init<<<(unsigned)((F * 6 + 255) / 256), 256, 0, s>>>(stats, F);
    k_color_stats<<<dim3(COLOR_SLICES, (unsigned)F), 256, 0, s>>>(rgb, frame_offsets, stats);
    k_color_apply<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(rgb, bidx, n, F, stats, out);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ row selection
// Rows of a batch that carry a given label (or a non-zero mask), frame by frame, order preserving: what the reference
// does per frame with np.where(seg == 2) (app/inference_engine.py:422-433). flags -> exclusive scan -> scatter; the
// per-segment offsets of the selection come out of the same scan (out_offsets[s] = rows selected before segment s).
__global__ void k_select_flags(const uint8_t* __restrict__ key, int want, int64_t n, int32_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    flag[i] = (want < 0 ? key[i] != 0 : key[i] == (uint8_t)want) ? 1 : 0;
}

__global__ void k_select_scatter(const int32_t* __restrict__ pos, const int32_t* __restrict__ total,
                                 const int32_t* __restrict__ src_rows, int64_t n,
                                 const int32_t* __restrict__ seg_offsets, int S, int32_t* __restrict__ out_rows,
                                 int32_t* __restrict__ out_offsets) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= S) {
        const int32_t so = seg_offsets[i];
        out_offsets[i] = so < n ? pos[so] : *total;
    }
    if (i >= n) return;
    const int32_t p = pos[i];
    const int32_t nxt = (i + 1 < n) ? pos[i + 1] : *total;
    if (nxt != p) out_rows[p] = src_rows ? src_rows[i] : (int32_t)i;
}

extern "C" size_t b2me_select_workspace_bytes(int64_t n) {
    const int64_t n1 = n > 0 ? n : 1;
    return align_up((size_t)(n1 + 1) * 4, 256) + align_up(16, 256) + scan_ws_bytes(n1 + 1);
}

extern "C" int b2me_select_rows(const uint8_t* key, int want, const int32_t* src_rows, int64_t n,
                                const int32_t* seg_offsets, int S, int32_t* out_rows, int32_t* out_offsets, void* ws,
                                size_t ws_bytes, b2me_stream_t stream) {
    if (!key || !seg_offsets || !out_rows || !out_offsets || !ws || n < 0 || S < 1 || want > 255) return B2ME_EINVAL;
    if (n >= (int64_t)1 << 31) return B2ME_EINVAL;
    if (ws_bytes < b2me_select_workspace_bytes(n)) return B2ME_EWORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    char* base = reinterpret_cast<char*>(ws);
    int32_t* pos = reinterpret_cast<int32_t*>(base);
    int32_t* total = reinterpret_cast<int32_t*>(base + align_up((size_t)((n > 0 ? n : 1) + 1) * 4, 256));
    void* scan = base + align_up((size_t)((n > 0 ? n : 1) + 1) * 4, 256) + align_up(16, 256);
    if (n > 0) k_select_flags<<<(unsigned)ceil_div64(n, 256), 256, 0, s>>>(key, want, n, pos);
    cudaMemsetAsync(pos + n, 0, 4, s);
    const int rc = exclusive_scan_i32(pos, n + 1, total, scan, s);
    if (rc != B2ME_OK) return rc;
    const unsigned G = (unsigned)ceil_div64((n > S ? n : S) + 1, 256);
    k_select_scatter<<<G, 256, 0, s>>>(pos, total, src_rows, n, seg_offsets, S, out_rows, out_offsets);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// gather of xyz / rgb rows + the frame index of every selected row (crop compaction of the EE points)
__global__ void k_gather_crops(const float* __restrict__ xyz, const float* __restrict__ rgb,
                               const int32_t* __restrict__ rows, int64_t m, const int32_t* __restrict__ seg_offsets,
                               int S, float* __restrict__ out_xyz, float* __restrict__ out_rgb,
                               float* __restrict__ out_seg) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    const int64_t r = rows[i];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        out_xyz[i * 3 + c] = xyz[r * 3 + c];
        if (out_rgb) out_rgb[i * 3 + c] = rgb[r * 3 + c];
    }
    if (out_seg) {
        int lo = 0, hi = S;  // largest s with seg_offsets[s] <= i
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (seg_offsets[mid] <= i) lo = mid;
            else hi = mid;
        }
        out_seg[i] = (float)lo;
    }
}

extern "C" int b2me_gather_crops(const float* xyz, const float* rgb, const int32_t* rows, int64_t m,
                                 const int32_t* seg_offsets, int S, float* out_xyz, float* out_rgb, float* out_seg,
                                 b2me_stream_t stream) {
    if (!xyz || !rows || !out_xyz || m < 0 || (out_rgb && !rgb) || (out_seg && (!seg_offsets || S < 1)))
        return B2ME_EINVAL;
    if (m == 0) return B2ME_OK;
    k_gather_crops<<<(unsigned)ceil_div64(m, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        xyz, rgb, rows, m, seg_offsets, S, out_xyz, out_rgb, out_seg);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// center_at_origin (utils/preprocess.py:8-11) per segment: offset = (max + min) / 2 in float32, points - offset
__global__ void __launch_bounds__(256) k_center_segments(const float* __restrict__ pts,
                                                         const int32_t* __restrict__ seg_offsets,
                                                         float* __restrict__ out_center,
                                                         float* __restrict__ out_centered) {
    __shared__ float smin[3][8], smax[3][8];
    __shared__ float off_s[3];
    const int seg = blockIdx.x;
    const int r0 = seg_offsets[seg], r1 = seg_offsets[seg + 1];
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = pts[(int64_t)r * 3 + c];
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if ((threadIdx.x & 31) == 0) { smin[c][threadIdx.x >> 5] = mn[c]; smax[c][threadIdx.x >> 5] = mx[c]; }
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float a = smin[threadIdx.x][0], b = smax[threadIdx.x][0];
        for (int w = 1; w < 8; ++w) { a = fminf(a, smin[threadIdx.x][w]); b = fmaxf(b, smax[threadIdx.x][w]); }
        const float o = r1 > r0 ? __fmul_rn(__fadd_rn(b, a), 0.5f) : 0.f;
        off_s[threadIdx.x] = o;
        out_center[seg * 3 + threadIdx.x] = o;
    }
    __syncthreads();
    if (!out_centered) return;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x)
#pragma unroll
        for (int c = 0; c < 3; ++c) out_centered[(int64_t)r * 3 + c] = __fsub_rn(pts[(int64_t)r * 3 + c], off_s[c]);
}

extern "C" int b2me_center_segments(const float* points_xyz, const int32_t* seg_offsets, int S, float* out_center,
                                    float* out_centered, b2me_stream_t stream) {
    if (!points_xyz || !seg_offsets || !out_center || S < 0) return B2ME_EINVAL;
    if (S == 0) return B2ME_OK;
    k_center_segments<<<(unsigned)S, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(points_xyz, seg_offsets,
                                                                                         out_center, out_centered);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}
