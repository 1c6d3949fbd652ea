// cluster.cu — K7: largest single-linkage cluster per segment (utils/output.py:13-28) as connected
// components of {d(i,j) < dist} with a lock-free union-find over a uniform grid of cell size 0.57 dist:
//   * the diagonal of a cell is 0.987 dist < dist, so all points of one cell are in one component: one union per
//     point with its cell's representative, no distance tests;
//   * two cells can only be linked when they are at most 2 cells apart per axis (a gap of 2 cells is 1.14 dist):
//     one warp per (cell, neighbour cell) pair skips pairs that are already in one component and otherwise tests
//     point pairs (fp64, the reference's threshold test) until the first one within dist.
// Exactly the components of the full graph, in O(n) distance tests for a dense blob instead of O(n^2).
//
// Determinism: unions always hook the larger root under the smaller one, so after flattening a
// component's label is its lowest point index whatever the thread interleaving; component sizes are
// integer atomics; the winner is max(size) with ties to the lowest label.
#include "common.cuh"

#define CLUSTER_CELL_FRAC 0.57  // cell edge / dist; sqrt(3) * 0.57 = 0.987 < 1

// from coords.cu (same translation unit set; declared here)
extern "C" int b2me_quantize_unique(const float*, const int32_t*, int64_t, const float*, int, int, int32_t*, float*,
                                    int32_t*, int32_t*, int32_t*, void*, size_t, void*, size_t, b2me_stream_t);

struct ClusterWs {
    int4* cellq;
    int32_t* cell_coords;  // [n,4]
    int32_t* cell_of;      // [n]
    int32_t* counts;       // [2]
    int32_t* cell_start;   // [n+1]
    int32_t* cursor;       // [n]
    int32_t* sorted;       // [n]
    int32_t* parent;       // [n]
    int32_t* comp_size;    // [n]
    unsigned long long* best;  // [S]
    void* table;
    size_t table_bytes;
    void* uws;
    size_t uws_bytes;
    void* scan;
    size_t total;
};

static ClusterWs carve_cluster_ws(void* ws, int64_t n, int S) {
    ClusterWs w;
    char* base = reinterpret_cast<char*>(ws);
    size_t off = 0;
    const int64_t n1 = n > 0 ? n : 1;
    auto take = [&](size_t bytes) {
        void* p = base + off;
        off += align_up(bytes, 256);
        return p;
    };
    w.cellq = reinterpret_cast<int4*>(take((size_t)n1 * 16));
    w.cell_coords = reinterpret_cast<int32_t*>(take((size_t)n1 * 16));
    w.cell_of = reinterpret_cast<int32_t*>(take((size_t)n1 * 4));
    w.counts = reinterpret_cast<int32_t*>(take(16));
    w.cell_start = reinterpret_cast<int32_t*>(take((size_t)(n1 + 1) * 4));
    w.cursor = reinterpret_cast<int32_t*>(take((size_t)n1 * 4));
    w.sorted = reinterpret_cast<int32_t*>(take((size_t)n1 * 4));
    w.parent = reinterpret_cast<int32_t*>(take((size_t)n1 * 4));
    w.comp_size = reinterpret_cast<int32_t*>(take((size_t)n1 * 4));
    w.best = reinterpret_cast<unsigned long long*>(take((size_t)(S > 0 ? S : 1) * 8));
    w.table_bytes = b2me_table_bytes(n1);
    w.table = take(w.table_bytes);
    w.uws_bytes = b2me_unique_workspace_bytes(n1, 0);
    w.uws = take(w.uws_bytes);
    w.scan = take(scan_ws_bytes(n1 + 1));
    w.total = off;
    return w;
}

extern "C" size_t b2me_cluster_workspace_bytes(int64_t n, int S) { return carve_cluster_ws(nullptr, n, S).total; }

__device__ __forceinline__ int seg_of_row(const int32_t* __restrict__ seg_offsets, int S, int i) {
    int lo = 0, hi = S;  // largest s with seg_offsets[s] <= i
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (seg_offsets[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

__global__ void k_cluster_cells(const float* __restrict__ pts, const int32_t* __restrict__ seg_offsets, int S,
                                int64_t n, double dist, int4* __restrict__ cellq, int32_t* __restrict__ parent,
                                int32_t* __restrict__ comp_size) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int seg = seg_of_row(seg_offsets, S, (int)i);
    const double cs = dist * CLUSTER_CELL_FRAC;
    const double fx = floor((double)pts[i * 3 + 0] / cs), fy = floor((double)pts[i * 3 + 1] / cs),
                 fz = floor((double)pts[i * 3 + 2] / cs);
    // a non-finite or far-away point gets a cell outside the key range: b2me_quantize_unique then raises its error flag
    // (reported through out_sizes = -1) instead of the point silently landing in cell 0
    const bool ok = isfinite(fx) && isfinite(fy) && isfinite(fz) && fabs(fx) < 1e9 && fabs(fy) < 1e9 && fabs(fz) < 1e9;
    const int cx = ok ? (int)fx : 0x7fffffff, cy = ok ? (int)fy : 0, cz = ok ? (int)fz : 0;
    cellq[i] = make_int4(seg, cx, cy, cz);
    parent[i] = (int32_t)i;
    comp_size[i] = 0;
}

__global__ void k_cell_count(const int32_t* __restrict__ cell_of, int64_t n, int32_t* __restrict__ cell_start) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t c = cell_of[i];
    if (c >= 0) atomicAdd(cell_start + c, 1);
}

__global__ void k_cell_fill(const int32_t* __restrict__ cell_of, int64_t n, const int32_t* __restrict__ cell_start,
                            int32_t* __restrict__ cursor, int32_t* __restrict__ sorted) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t c = cell_of[i];
    if (c < 0) return;
    const int pos = atomicAdd(cursor + c, 1);
    sorted[cell_start[c] + pos] = (int32_t)i;
}

__device__ __forceinline__ int uf_find(int32_t* parent, int x) {
    // path halving with plain stores is safe: parents only ever decrease towards the root
    while (true) {
        const int p = ((volatile int32_t*)parent)[x];
        if (p == x) return x;
        const int gp = ((volatile int32_t*)parent)[p];
        if (gp != p) parent[x] = gp;
        x = p;
    }
}

__device__ __forceinline__ void uf_union(int32_t* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }  // a = larger root, hooked under b
        const int old = atomicMin(parent + a, b);
        if (old == a) return;  // a was still a root: hooked
        a = old;               // somebody re-parented a meanwhile; retry with what they wrote
    }
}

// every point joins the representative (first sorted member) of its cell
__global__ void k_cluster_intra(const int32_t* __restrict__ cell_of, int64_t n, const int32_t* __restrict__ cell_start,
                                const int32_t* __restrict__ sorted, int32_t* __restrict__ parent) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int c = cell_of[i];
    if (c < 0) return;
    const int rep = sorted[cell_start[c]];
    if (rep != (int)i) uf_union(parent, (int)i, rep);
}

// one warp per (cell a, neighbour offset in the positive half space, |d| <= reach): link the two cells when any
// point pair is within dist. Pairs already in one component are skipped, so the second pass (reach 2) only tests
// pairs of cells that the first pass (reach 1) left in different components.
__global__ void __launch_bounds__(256)
k_cluster_inter(const float* __restrict__ pts, const int32_t* __restrict__ cell_coords, const int32_t* __restrict__ ncells_p,
                const HashSlot* __restrict__ tab, unsigned long long mask, const int32_t* __restrict__ cell_start,
                const int32_t* __restrict__ sorted, double dist2, int reach, int32_t* __restrict__ parent) {
    const int ncells = *ncells_p;
    const int side = 2 * reach + 1;
    const int noff = (side * side * side) / 2;  // offsets after the centre in (dz, dy, dx) lexicographic order
    const long long ntask = (long long)ncells * noff;
    const int lane = threadIdx.x & 31;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < ntask; t += wstride) {
        const int a = (int)(t / noff);
        const int o = (int)(t - (long long)a * noff) + noff + 1;  // index in the side^3 cube, strictly after the centre
        const int dx = o % side - reach, dy = (o / side) % side - reach, dz = o / (side * side) - reach;
        const int4 ca = *reinterpret_cast<const int4*>(cell_coords + 4 * (long long)a);
        const int x = ca.y + dx, y = ca.z + dy, z = ca.w + dz;
        if (!coord_in_range(ca.x, x, y, z)) continue;
        const unsigned int b = table_lookup(tab, mask, pack_key(ca.x, x, y, z));
        if (b == 0xFFFFFFFFu) continue;
        const int a0 = cell_start[a], a1 = cell_start[a + 1], b0 = cell_start[b], b1 = cell_start[b + 1];
        const int repa = sorted[a0], repb = sorted[b0];
        // lane 0 alone walks (and path-halves) the forest that other warps are mutating; its verdict is broadcast, so
        // the skip below is warp-uniform by construction and every lane reaches the full-mask votes that follow
        int same = 0;
        if (lane == 0) same = uf_find(parent, repa) == uf_find(parent, repb);
        same = __shfl_sync(0xffffffffu, same, 0);
        if (same) continue;
        bool linked = false;
        for (int ia = a0; ia < a1 && !linked; ++ia) {
            const int i = sorted[ia];
            const double xi = pts[(int64_t)i * 3], yi = pts[(int64_t)i * 3 + 1], zi = pts[(int64_t)i * 3 + 2];
            for (int jb = b0; jb < b1; jb += 32) {
                bool hit = false;
                if (jb + lane < b1) {
                    const int j = sorted[jb + lane];
                    const double ddx = xi - (double)pts[(int64_t)j * 3];
                    const double ddy = yi - (double)pts[(int64_t)j * 3 + 1];
                    const double ddz = zi - (double)pts[(int64_t)j * 3 + 2];
                    hit = (ddx * ddx + ddy * ddy + ddz * ddz) < dist2;
                }
                if (__any_sync(0xffffffffu, hit)) { linked = true; break; }
            }
        }
        if (linked && lane == 0) uf_union(parent, repa, repb);
        __syncwarp();
    }
}

__global__ void k_cluster_flatten(int32_t* __restrict__ parent, int64_t n, int32_t* __restrict__ comp_size) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x = (int)i;
    while (true) {  // read-only walk: no writes race with other walkers
        const int p = parent[x];
        if (p == x) break;
        x = p;
    }
    atomicAdd(comp_size + x, 1);
}

// separate pass so that no thread reads a parent entry another thread is flattening
__global__ void k_cluster_root(const int32_t* __restrict__ parent, int64_t n, int32_t* __restrict__ root_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int x = (int)i;
    while (true) {
        const int p = parent[x];
        if (p == x) break;
        x = p;
    }
    root_out[i] = x;
}

__global__ void k_cluster_best(const int32_t* __restrict__ comp_size, const int32_t* __restrict__ seg_offsets, int S,
                               int64_t n, unsigned long long* __restrict__ best) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int sz = comp_size[i];
    if (sz <= 0) return;  // not a root
    const int seg = seg_of_row(seg_offsets, S, (int)i);
    const unsigned long long key = ((unsigned long long)(unsigned)sz << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
    atomicMax(best + seg, key);
}

__global__ void k_cluster_mask(const int32_t* __restrict__ root, const int32_t* __restrict__ seg_offsets, int S,
                               int64_t n, const unsigned long long* __restrict__ best,
                               const int32_t* __restrict__ counts, uint8_t* __restrict__ mask,
                               int32_t* __restrict__ sizes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    // counts[1] != 0: a point was non-finite / outside the cell key range -> every size reads -1 (error)
    if (i < S && sizes) sizes[i] = counts[1] ? -1 : (int32_t)(best[i] >> 32);
    if (i >= n) return;
    const int seg = seg_of_row(seg_offsets, S, (int)i);
    const unsigned long long b = best[seg];
    const int best_root = (int)(0xFFFFFFFFu - (unsigned)(b & 0xFFFFFFFFull));
    mask[i] = (b != 0ull && root[i] == best_root) ? 1 : 0;
}

extern "C" int b2me_largest_cluster(const float* points_xyz, const int32_t* seg_offsets, int S, int64_t n, double dist,
                                    uint8_t* out_mask, int32_t* out_sizes, void* ws, size_t ws_bytes,
                                    b2me_stream_t stream) {
    if (!points_xyz || !seg_offsets || !out_mask || !ws || S <= 0 || S > 1024 || n < 0 || !(dist > 0)) return B2ME_EINVAL;
    if (n >= (int64_t)1 << 31) return B2ME_EINVAL;
    ClusterWs w = carve_cluster_ws(ws, n, S);
    if (ws_bytes < w.total) return B2ME_EWORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    cudaMemsetAsync(w.best, 0, (size_t)S * 8, s);
    cudaMemsetAsync(w.counts, 0, 2 * sizeof(int32_t), s);
    const int T = 256;
    const unsigned G = (unsigned)ceil_div64(n > 0 ? n : 1, T);
    if (n > 0) {
        k_cluster_cells<<<G, T, 0, s>>>(points_xyz, seg_offsets, S, n, dist, w.cellq, w.parent, w.comp_size);
        int rc = b2me_quantize_unique(nullptr, reinterpret_cast<const int32_t*>(w.cellq), n, nullptr, 0, 0,
                                      w.cell_coords, nullptr, w.cell_of, nullptr, w.counts, w.table, w.table_bytes,
                                      w.uws, w.uws_bytes, stream);
        if (rc != B2ME_OK) return rc;
        // cell occupancy -> exclusive offsets (n+1 entries so that cell_start[cid+1] is always readable)
        cudaMemsetAsync(w.cell_start, 0, (size_t)(n + 1) * 4, s);
        cudaMemsetAsync(w.cursor, 0, (size_t)n * 4, s);
        k_cell_count<<<G, T, 0, s>>>(w.cell_of, n, w.cell_start);
        rc = exclusive_scan_i32(w.cell_start, n + 1, w.counts + 2, w.scan, s);
        if (rc != B2ME_OK) return rc;
        k_cell_fill<<<G, T, 0, s>>>(w.cell_of, n, w.cell_start, w.cursor, w.sorted);
        const unsigned long long mask = (unsigned long long)(b2me_table_slots(n) - 1);
        k_cluster_intra<<<G, T, 0, s>>>(w.cell_of, n, w.cell_start, w.sorted, w.parent);
        for (int reach = 1; reach <= 2; ++reach)
            k_cluster_inter<<<B2ME_NUM_SMS * 4, 256, 0, s>>>(points_xyz, w.cell_coords, w.counts,
                                                             reinterpret_cast<const HashSlot*>(w.table), mask,
                                                             w.cell_start, w.sorted, dist * dist, reach, w.parent);
        k_cluster_flatten<<<G, T, 0, s>>>(w.parent, n, w.comp_size);
        k_cluster_root<<<G, T, 0, s>>>(w.parent, n, w.cursor);  // cursor reused as root[]
        k_cluster_best<<<G, T, 0, s>>>(w.comp_size, seg_offsets, S, n, w.best);
    }
    const unsigned G2 = (unsigned)ceil_div64((n > S ? n : S), T);
    k_cluster_mask<<<G2, T, 0, s>>>(w.cursor, seg_offsets, S, n, w.best, w.counts, out_mask, out_sizes);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}
