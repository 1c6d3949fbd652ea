// pose.cu — K9 batched 3x3 Kabsch/SVD and K10 batched point-to-point ICP (fp64 arithmetic on fp32 clouds).
//
// ICP: one CTA per frame. The frame's target (EE) points are binned once into a dense uniform grid
// (<= 32^3 cells, cell >= 2 cm) by a preparation kernel; every iteration each thread takes source (CAD)
// points, transforms them with the current 4x4 (fp64), finds the EXACT nearest target point within
// max_corr by an expanding-ring search over the grid, and the block reduces the 17 sums Kabsch needs in
// a fixed order. The loop and stopping rule restate Open3D's RegistrationICP as called by
// utils/icp.py:65-71.
#include "common.cuh"

// ------------------------------------------------------------------------------------------ 3x3 helpers
__device__ inline void jacobi_eig3(double A[3][3], double V[3][3], double w[3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j) ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 30; ++sweep) {
        const double offd = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (offd <= 1e-300 || offd <= 1e-22 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double apq = A[p][q];
                if (fabs(apq) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    w[0] = A[0][0]; w[1] = A[1][1]; w[2] = A[2][2];
}

__device__ inline void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ inline double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// R = argmin sum |R a_i - b_i|^2 for H = sum a_i b_i^T (centred), det(R) = +1.
// Equivalent to numpy's U,S,Vt = svd(H); R = Vt.T U.T with the last row of Vt flipped when det < 0
// (utils/transformation.py:207-218): we build right-handed U and V from the two leading singular pairs.
__device__ inline void kabsch_from_H(const double H[3][3], double R[3][3]) {
    double A[3][3], V[3][3], w[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = H[0][i] * H[0][j] + H[1][i] * H[1][j] + H[2][i] * H[2][j];  // H^T H
    jacobi_eig3(A, V, w);
    int o[3] = {0, 1, 2};  // sort eigenvalues descending
    if (w[o[0]] < w[o[1]]) { int t = o[0]; o[0] = o[1]; o[1] = t; }
    if (w[o[1]] < w[o[2]]) { int t = o[1]; o[1] = o[2]; o[2] = t; }
    if (w[o[0]] < w[o[1]]) { int t = o[0]; o[0] = o[1]; o[1] = t; }
    double v0[3], v1[3], v2[3], u0[3], u1[3], u2[3];
    for (int i = 0; i < 3; ++i) { v0[i] = V[i][o[0]]; v1[i] = V[i][o[1]]; }
    cross3(v0, v1, v2);
    for (int i = 0; i < 3; ++i) {
        u0[i] = H[i][0] * v0[0] + H[i][1] * v0[1] + H[i][2] * v0[2];
        u1[i] = H[i][0] * v1[0] + H[i][1] * v1[1] + H[i][2] * v1[2];
    }
    double n0 = norm3(u0);
    if (n0 < 1e-300) {  // H == 0: identity
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
        return;
    }
    for (int i = 0; i < 3; ++i) u0[i] /= n0;
    const double d = u1[0] * u0[0] + u1[1] * u0[1] + u1[2] * u0[2];
    for (int i = 0; i < 3; ++i) u1[i] -= d * u0[i];
    double n1 = norm3(u1);
    if (n1 < 1e-14 * n0) {  // rank 1: any unit vector orthogonal to u0 (and the matching one for v)
        double e[3] = {1, 0, 0};
        if (fabs(u0[0]) > 0.9) { e[0] = 0; e[1] = 1; }
        cross3(u0, e, u1);
        n1 = norm3(u1);
    }
    for (int i = 0; i < 3; ++i) u1[i] /= n1;
    cross3(u0, u1, u2);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = v0[i] * u0[j] + v1[i] * u1[j] + v2[i] * u2[j];  // V U^T
}

// ------------------------------------------------------------------------------------------ K9
__global__ void k_kabsch_batched(const double* __restrict__ ref, const double* __restrict__ tgt,
                                 const int32_t* __restrict__ npairs, int P, int kmax, double* __restrict__ out_R,
                                 double* __restrict__ out_t) {
    const int pb = blockIdx.x * blockDim.x + threadIdx.x;
    if (pb >= P) return;
    const int n = npairs[pb];
    const double* a = ref + (int64_t)pb * kmax * 3;
    const double* b = tgt + (int64_t)pb * kmax * 3;
    double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
        for (int d = 0; d < 3; ++d) { ca[d] += a[i * 3 + d]; cb[d] += b[i * 3 + d]; }
    const double inv = n > 0 ? 1.0 / n : 0.0;
    for (int d = 0; d < 3; ++d) { ca[d] *= inv; cb[d] *= inv; }
    double H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    for (int i = 0; i < n; ++i)
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) H[r][c] += (a[i * 3 + r] - ca[r]) * (b[i * 3 + c] - cb[c]);
    double R[3][3];
    kabsch_from_H(H, R);
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 3; ++c) out_R[(int64_t)pb * 9 + r * 3 + c] = R[r][c];
        out_t[(int64_t)pb * 3 + r] = cb[r] - (R[r][0] * ca[0] + R[r][1] * ca[1] + R[r][2] * ca[2]);
    }
}

extern "C" int b2me_kabsch_batched(const double* ref, const double* tgt, const int32_t* npairs, int P, int kmax,
                                   double* out_R, double* out_t, b2me_stream_t stream) {
    if (!ref || !tgt || !npairs || !out_R || !out_t || P < 0 || kmax <= 0) return B2ME_EINVAL;
    if (P == 0) return B2ME_OK;
    k_kabsch_batched<<<(unsigned)((P + 63) / 64), 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(ref, tgt, npairs, P,
                                                                                                   kmax, out_R, out_t);
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}

// ------------------------------------------------------------------------------------------ K10
// Batched point-to-point ICP. Per frame: the target (EE) points are binned once into a dense uniform grid
// (<= 20^3 cells, cell >= 2 cm, counting sort). The whole ICP loop is ONE launch (k_icp_persistent): a thread-block
// cluster per frame keeps the frame's targets in shared memory across all evaluations; a thread transforms a CAD
// point with the frame's current 4x4 (fp64), finds its EXACT nearest target within max_corr by a sphere-bounded
// search over the grid, the CTAs reduce the 17 Kabsch sums in a fixed order into per-CTA slots, and rank 0 of the
// cluster adds the slots in rank order (deterministic), applies the convergence test and the Kabsch update of
// Open3D's RegistrationICP loop; a frame leaves the loop when it has converged.
#define ICP_GRID 20
#define ICP_CELLS (ICP_GRID * ICP_GRID * ICP_GRID)
#define ICP_MIN_CELL 0.02
#define ICP_THREADS 256
#define ICP_MAX_CHUNKS 64
#define ICP_NSUM 17  // count, sum d2, sum p(3), sum q(3), sum p q^T (9)

struct IcpFrameGrid {
    double origin[3];
    double inv_h;
    double h;
    int dims[3];
    int pad;
};

struct IcpState {
    double T[16];
    double prev_fit, prev_rmse, fit, rmse, ncorr;
    int iters;
    int done;
    unsigned int counter;
    int pad;
};

struct IcpWs {
    IcpFrameGrid* grids;   // [F]
    IcpState* state;       // [F]
    int32_t* cell_start;   // [F, ICP_CELLS + 1]
    int32_t* cell_cursor;  // [F, ICP_CELLS]
    double* partial;       // [F, ICP_MAX_CHUNKS, ICP_NSUM]
    float4* sorted;        // [T_total] xyz + original index bits
    int32_t* match;        // [F, S] position (in `sorted`, frame-relative) of the last match of every source point
    size_t total;
};

static IcpWs carve_icp_ws(void* ws, int64_t T_total, int F, int S) {
    IcpWs w;
    char* base = reinterpret_cast<char*>(ws);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* p = base + off;
        off += align_up(bytes, 256);
        return p;
    };
    const int F1 = F > 0 ? F : 1;
    w.grids = reinterpret_cast<IcpFrameGrid*>(take((size_t)F1 * sizeof(IcpFrameGrid)));
    w.state = reinterpret_cast<IcpState*>(take((size_t)F1 * sizeof(IcpState)));
    w.cell_start = reinterpret_cast<int32_t*>(take((size_t)F1 * (ICP_CELLS + 1) * 4));
    w.cell_cursor = reinterpret_cast<int32_t*>(take((size_t)F1 * ICP_CELLS * 4));
    w.partial = reinterpret_cast<double*>(take((size_t)F1 * ICP_MAX_CHUNKS * ICP_NSUM * sizeof(double)));
    w.sorted = reinterpret_cast<float4*>(take((size_t)(T_total > 0 ? T_total : 1) * sizeof(float4)));
    w.match = reinterpret_cast<int32_t*>(take((size_t)F1 * (size_t)(S > 0 ? S : 1) * 4));
    w.total = off;
    return w;
}

extern "C" size_t b2me_icp_workspace_bytes(int64_t T_total, int F, int S) {
    return carve_icp_ws(nullptr, T_total, F, S).total;
}

__device__ __forceinline__ int icp_cell_coord(double v, double origin, double inv_h, int dim) {
    int c = (int)floor((v - origin) * inv_h);
    return c < 0 ? 0 : (c >= dim ? dim - 1 : c);
}

// one block per frame: bbox -> grid, counting sort of the frame's target points by cell, state <- init
__global__ void __launch_bounds__(ICP_THREADS)
k_icp_build_grid(const float* __restrict__ tgt, const int32_t* __restrict__ tgt_offsets,
                 const double* __restrict__ init_T, IcpFrameGrid* __restrict__ grids, IcpState* __restrict__ state,
                 int32_t* __restrict__ cell_start_all, int32_t* __restrict__ cursor_all, float4* __restrict__ sorted) {
    __shared__ float smin[3][ICP_THREADS], smax[3][ICP_THREADS];
    __shared__ IcpFrameGrid g;
    __shared__ int scan_s[33];
    __shared__ int carry_s;
    const int f = blockIdx.x;
    const int t0 = tgt_offsets[f], t1 = tgt_offsets[f + 1];
    int32_t* cell_start = cell_start_all + (int64_t)f * (ICP_CELLS + 1);
    int32_t* cursor = cursor_all + (int64_t)f * ICP_CELLS;
    if (threadIdx.x < 16) state[f].T[threadIdx.x] = init_T[(int64_t)f * 16 + threadIdx.x];
    if (threadIdx.x == 0) {
        IcpState* st = state + f;
        st->prev_fit = st->prev_rmse = st->fit = st->rmse = st->ncorr = 0.0;
        st->iters = 0;
        st->done = 0;
        st->counter = 0u;
        st->pad = 0;
    }
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int i = t0 + threadIdx.x; i < t1; i += blockDim.x)
        for (int a = 0; a < 3; ++a) {
            const float v = tgt[(int64_t)i * 3 + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    for (int a = 0; a < 3; ++a) { smin[a][threadIdx.x] = mn[a]; smax[a][threadIdx.x] = mx[a]; }
    __syncthreads();
    for (int o = ICP_THREADS / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o)
            for (int a = 0; a < 3; ++a) {
                smin[a][threadIdx.x] = fminf(smin[a][threadIdx.x], smin[a][threadIdx.x + o]);
                smax[a][threadIdx.x] = fmaxf(smax[a][threadIdx.x], smax[a][threadIdx.x + o]);
            }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double ext = 0.0;
        for (int a = 0; a < 3; ++a) {
            const double e = (t1 > t0) ? (double)smax[a][0] - (double)smin[a][0] : 0.0;
            ext = e > ext ? e : ext;
        }
        double h = ext / (ICP_GRID - 1);
        if (h < ICP_MIN_CELL) h = ICP_MIN_CELL;
        g.h = h;
        g.inv_h = 1.0 / h;
        for (int a = 0; a < 3; ++a) {
            g.origin[a] = (t1 > t0) ? (double)smin[a][0] : 0.0;
            const double e = (t1 > t0) ? (double)smax[a][0] - (double)smin[a][0] : 0.0;
            int d = (int)floor(e / h) + 1;
            g.dims[a] = d > ICP_GRID ? ICP_GRID : (d < 1 ? 1 : d);
        }
        g.pad = 0;
        grids[f] = g;
    }
    __syncthreads();
    const int ncell = g.dims[0] * g.dims[1] * g.dims[2];
    for (int c = threadIdx.x; c <= ncell; c += blockDim.x) cell_start[c] = 0;
    for (int c = threadIdx.x; c < ncell; c += blockDim.x) cursor[c] = 0;
    __syncthreads();
    for (int i = t0 + threadIdx.x; i < t1; i += blockDim.x) {
        const int cx = icp_cell_coord(tgt[(int64_t)i * 3], g.origin[0], g.inv_h, g.dims[0]);
        const int cy = icp_cell_coord(tgt[(int64_t)i * 3 + 1], g.origin[1], g.inv_h, g.dims[1]);
        const int cz = icp_cell_coord(tgt[(int64_t)i * 3 + 2], g.origin[2], g.inv_h, g.dims[2]);
        atomicAdd(&cell_start[(cz * g.dims[1] + cy) * g.dims[0] + cx], 1);
    }
    __syncthreads();
    // exclusive scan of ncell+1 counters by the block (chunks of blockDim)
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base <= ncell; base += blockDim.x) {
        const int c = base + threadIdx.x;
        const int v = (c <= ncell) ? cell_start[c] : 0;
        const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
        int inc = v;
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += t;
        }
        if (lane == 31) scan_s[wp] = inc;
        __syncthreads();
        if (wp == 0) {
            const int wv = lane < (ICP_THREADS / 32) ? scan_s[lane] : 0;
            int winc = wv;
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, winc, o);
                if (lane >= o) winc += t;
            }
            scan_s[lane] = winc - wv;
            if (lane == 31) scan_s[32] = winc;
        }
        __syncthreads();
        const int carry = carry_s;
        if (c <= ncell) cell_start[c] = inc - v + scan_s[wp] + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + scan_s[32];
        __syncthreads();
    }
    for (int i = t0 + threadIdx.x; i < t1; i += blockDim.x) {
        const float x = tgt[(int64_t)i * 3], y = tgt[(int64_t)i * 3 + 1], z = tgt[(int64_t)i * 3 + 2];
        const int cx = icp_cell_coord(x, g.origin[0], g.inv_h, g.dims[0]);
        const int cy = icp_cell_coord(y, g.origin[1], g.inv_h, g.dims[1]);
        const int cz = icp_cell_coord(z, g.origin[2], g.inv_h, g.dims[2]);
        const int cell = (cz * g.dims[1] + cy) * g.dims[0] + cx;
        const int pos = atomicAdd(&cursor[cell], 1);
        sorted[t0 + cell_start[cell] + pos] = make_float4(x, y, z, __int_as_float(i - t0));
    }
}

// convergence test + Kabsch update of one frame (runs on one thread of the frame's last CTA)
__device__ __noinline__ void icp_frame_update(IcpState* st, const double* tot, const double* T_s, int f, int S, int ev,
                                              int max_iter, double rel_fitness, double rel_rmse,
                                              double* __restrict__ out_T, double* __restrict__ out_stats) {
    st->counter = 0u;
    const double ncorr = tot[0];
    const double fit = S > 0 ? ncorr / (double)S : 0.0;
    const double rmse = ncorr > 0 ? sqrt(tot[1] / ncorr) : 0.0;
    bool stop = false;
    if (ev > 0) {
        st->iters = ev;
        if (fabs(st->prev_fit - fit) < rel_fitness && fabs(st->prev_rmse - rmse) < rel_rmse) stop = true;
    }
    if (ev == max_iter) stop = true;
    st->fit = fit; st->rmse = rmse; st->ncorr = ncorr;
    if (!stop) {
        st->prev_fit = fit; st->prev_rmse = rmse;
        if (ncorr > 0) {
            // update = Kabsch(transformed source -> matched targets) (Eigen::umeyama without scaling)
            const double n = ncorr;
            const double mp[3] = {tot[2] / n, tot[3] / n, tot[4] / n};
            const double mq[3] = {tot[5] / n, tot[6] / n, tot[7] / n};
            double H[3][3], R[3][3];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) H[r][c] = tot[8 + r * 3 + c] - n * mp[r] * mq[c];
            kabsch_from_H(H, R);
            double t[3];
            for (int r = 0; r < 3; ++r) t[r] = mq[r] - (R[r][0] * mp[0] + R[r][1] * mp[1] + R[r][2] * mp[2]);
            double Tn[12];
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 4; ++c)
                    Tn[r * 4 + c] = R[r][0] * T_s[c] + R[r][1] * T_s[4 + c] + R[r][2] * T_s[8 + c];
                Tn[r * 4 + 3] += t[r];
            }
            for (int q = 0; q < 12; ++q) st->T[q] = Tn[q];
        }
    } else {
        st->done = 1;
        for (int q = 0; q < 16; ++q) out_T[(int64_t)f * 16 + q] = st->T[q];
        out_stats[(int64_t)f * 4 + 0] = fit;
        out_stats[(int64_t)f * 4 + 1] = rmse;
        out_stats[(int64_t)f * 4 + 2] = (double)st->iters;
        out_stats[(int64_t)f * 4 + 3] = ncorr;
    }
}

// Exact nearest target of (px,py,pz) within sqrt(R2): every grid cell that intersects the sphere of radius sqrt(best)
// around the query is scanned, and `best` shrinks as candidates are found. z planes and y rows are visited centre-out
// and pruned by their slab distance; inside a row the cells that can still hold a closer point form ONE contiguous x
// range, i.e. one contiguous run of the cell-sorted target array (x is the fastest cell index).
// `seed` = position of this source point's match in the previous iteration (or -1): the pose moves little between
// iterations, so the seed's distance is a tight bound and only a thin shell of candidates is touched.
#define ICP_LB_SLACK 1e-9  // metres: absorbs the rounding of the cell assignment at cell faces
__device__ __forceinline__ double icp_slab_dist(double v, double lo, double h) {
    const double a = lo - v, b = v - (lo + h);
    double d = a > b ? a : b;
    d -= ICP_LB_SLACK;
    return d > 0.0 ? d : 0.0;
}

template <bool kSmemTargets>
__device__ __forceinline__ void icp_search(const float4* __restrict__ tg, const int* cs, const IcpFrameGrid& g,
                                           double px, double py, double pz, double R2, int seed, double& best,
                                           int& best_j, int& best_s, double& bx, double& by, double& bz) {
    best = R2;  // strict '<' below: only neighbours inside the radius qualify
    best_j = 0x7FFFFFFF;
    best_s = -1;
    // fp32 pre-filter in front of the exact fp64 test (which alone decides, so the result is the fp64 result): a
    // candidate whose fp32 squared distance exceeds best + `slack` cannot win. slack bounds the fp32 error for every
    // candidate within sqrt(best) <= sqrt(R2) of the query: rounding the query to fp32 and the subtraction cost at most
    // e = 3 * 2^-23 (|p|_inf + 2 r) in the coordinates, the squared distance then moves by <= 2 r e + e^2 + 2^-22 d2;
    // four times that bound is used. Most candidates are rejected after 3 FADD + 3 FFMA instead of 3 fp32->fp64
    // conversions + 6 fp64 operations.
    const float pxf = (float)px, pyf = (float)py, pzf = (float)pz;
    const float rmax = sqrtf((float)R2);
    const float pinf = fmaxf(fabsf(pxf), fmaxf(fabsf(pyf), fabsf(pzf)));
    const float e32 = 3.6e-7f * (pinf + 2.f * rmax);
    const float slack = 4.f * (2.f * rmax * e32 + e32 * e32 + 2.4e-7f * (float)R2);
    float thr = __double2float_ru(best) + slack;
    auto load = [&](int s) -> float4 { return kSmemTargets ? tg[s] : __ldg(tg + s); };
    auto d2f = [&](const float4& q) -> float {
        const float fx = pxf - q.x, fy = pyf - q.y, fz = pzf - q.z;
        return fmaf(fx, fx, fmaf(fy, fy, fz * fz));
    };
    auto exact = [&](int s, const float4& q) {   // the deciding test, fp64 like Open3D's KD-tree distances
        const double dx = px - (double)q.x, dy = py - (double)q.y, dz = pz - (double)q.z;
        const double d2 = dx * dx + dy * dy + dz * dz;
        const int j = __float_as_int(q.w);
        if (d2 < best || (d2 == best && j < best_j && best_j != 0x7FFFFFFF)) {
            best = d2; best_j = j; best_s = s; bx = q.x; by = q.y; bz = q.z;
            thr = __double2float_ru(best) + slack;
        }
    };
    auto consider = [&](int s) {
        const float4 q = load(s);
        if (d2f(q) <= thr) exact(s, q);
    };
    // a run of candidates, four at a time: the four loads and fp32 distances are independent (the search is bound by
    // the latency of load -> distance -> compare chains, not by issue slots), one branch rejects all four in the
    // common case; survivors take the exact test in index order, so the result equals the one-by-one scan (thr only
    // shrinks: testing against the value from before the group is conservative)
    auto consider_run = [&](int s0, int s1) {
        int s = s0;
        for (; s + 4 <= s1; s += 4) {
            const float4 q0 = load(s), q1 = load(s + 1), q2 = load(s + 2), q3 = load(s + 3);
            const float e0 = d2f(q0), e1 = d2f(q1), e2 = d2f(q2), e3 = d2f(q3);
            if (fminf(fminf(e0, e1), fminf(e2, e3)) > thr) continue;
            if (e0 <= thr) exact(s, q0);
            if (e1 <= thr) exact(s + 1, q1);
            if (e2 <= thr) exact(s + 2, q2);
            if (e3 <= thr) exact(s + 3, q3);
        }
        for (; s < s1; ++s) consider(s);
    };
    if (seed >= 0) consider(seed);
    const int cy = icp_cell_coord(py, g.origin[1], g.inv_h, g.dims[1]);
    const int cz = icp_cell_coord(pz, g.origin[2], g.inv_h, g.dims[2]);
    // planes / rows farther than `reach` cells from the query's (clamped) cell cannot intersect the sphere
    const int reach = (int)(sqrt(best) * g.inv_h) + 2;
    const int zspan = min(max(cz, g.dims[2] - 1 - cz), reach), yspan = min(max(cy, g.dims[1] - 1 - cy), reach);
    for (int kz = 0; kz <= 2 * zspan; ++kz) {
        const int z = cz + ((kz & 1) ? -((kz + 1) >> 1) : (kz >> 1));
        if (z < 0 || z >= g.dims[2]) continue;
        const double dz = icp_slab_dist(pz, g.origin[2] + z * g.h, g.h);
        const double dz2 = dz * dz;
        if (dz2 >= best) continue;
        for (int ky = 0; ky <= 2 * yspan; ++ky) {
            const int y = cy + ((ky & 1) ? -((ky + 1) >> 1) : (ky >> 1));
            if (y < 0 || y >= g.dims[1]) continue;
            const double dy = icp_slab_dist(py, g.origin[1] + y * g.h, g.h);
            const double rem = best - dz2 - dy * dy;
            if (rem <= 0.0) continue;
            const double rx = sqrt(rem) + ICP_LB_SLACK;
            int x0 = (int)floor((px - rx - g.origin[0]) * g.inv_h);
            int x1 = (int)floor((px + rx - g.origin[0]) * g.inv_h);
            if (x0 < 0) x0 = 0;
            if (x1 > g.dims[0] - 1) x1 = g.dims[0] - 1;
            if (x0 > x1) continue;
            const int row = (z * g.dims[1] + y) * g.dims[0];
            const int s0 = cs[row + x0], s1 = cs[row + x1 + 1];
            consider_run(s0, s1);
        }
    }
}

#define ICP_EVAL_THREADS 512
#define ICP_EVAL_SMEM (200 * 1024)  // dynamic smem of k_icp_persistent: cell offsets + as many target points as fit
#define ICP_MAX_CLUSTER 8

__device__ __forceinline__ unsigned icp_cluster_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned icp_cluster_size() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// cluster-wide barrier with release / acquire ordering of global memory between the CTAs of the cluster
__device__ __forceinline__ void icp_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ONE launch for the whole ICP loop. A thread-block CLUSTER (1, 2, 4 or 8 CTAs: as many as the GPU has room for) owns
// one frame: every CTA stages the frame's cell-sorted targets and cell offsets in its shared memory ONCE and takes
// the source points rank * 512 + tid + q * (cluster size * 512) of every evaluation. Per evaluation: queries ->
// fixed-order block reduction of the 17 Kabsch sums -> this CTA's slot in global memory -> cluster barrier -> rank 0
// adds the slots in rank order, applies Open3D's stopping rule and the Kabsch update (icp_frame_update) -> cluster
// barrier -> every CTA re-reads T / done. Frames leave the loop as soon as they have converged (no launches for
// finished frames, no reload of the targets per evaluation, no atomics: results do not depend on scheduling).
__global__ void __launch_bounds__(ICP_EVAL_THREADS)
k_icp_persistent(const float* __restrict__ src, int S, const int32_t* __restrict__ tgt_offsets,
                 const IcpFrameGrid* __restrict__ grids, const int32_t* __restrict__ cell_start_all,
                 const float4* __restrict__ sorted, int32_t* __restrict__ match_all, IcpState* __restrict__ state,
                 double* __restrict__ partial_all, int tcap, double max_corr, int max_iter, double rel_fitness,
                 double rel_rmse, double* __restrict__ out_T, double* __restrict__ out_stats) {
    extern __shared__ __align__(16) unsigned char icp_smem[];
    float4* tg_s = reinterpret_cast<float4*>(icp_smem);
    int* cs = reinterpret_cast<int*>(icp_smem + (size_t)tcap * sizeof(float4));
    __shared__ double T_s[12];
    __shared__ double red[ICP_EVAL_THREADS / 32][ICP_NSUM];
    __shared__ double tot[ICP_NSUM];
    __shared__ IcpFrameGrid g;
    const int csize = (int)icp_cluster_size(), rank = (int)icp_cluster_rank();
    const int f = blockIdx.x / csize;
    IcpState* st = state + f;
    const int t0 = tgt_offsets[f];
    const int nT = tgt_offsets[f + 1] - t0;
    if (threadIdx.x == 0) g = grids[f];
    __syncthreads();
    const int ncell = g.dims[0] * g.dims[1] * g.dims[2];
    {
        const int32_t* cell_start = cell_start_all + (int64_t)f * (ICP_CELLS + 1);
        for (int c = threadIdx.x; c <= ncell; c += blockDim.x) cs[c] = cell_start[c];
    }
    const bool in_smem = nT <= tcap;
    if (in_smem)
        for (int i = threadIdx.x; i < nT; i += blockDim.x) tg_s[i] = __ldg(sorted + t0 + i);
    const double R2 = max_corr * max_corr;
    int32_t* match = match_all + (int64_t)f * S;
    double* partial = partial_all + (int64_t)f * ICP_MAX_CHUNKS * ICP_NSUM;
    const int stride = csize * ICP_EVAL_THREADS;
    for (int ev = 0; ev <= max_iter; ++ev) {
        // the state of this frame as rank 0 left it (ordered by the cluster barrier at the end of the previous round)
        if (threadIdx.x < 12) T_s[threadIdx.x] = __ldcg(st->T + threadIdx.x);
        __syncthreads();
        double acc[ICP_NSUM];
#pragma unroll
        for (int q = 0; q < ICP_NSUM; ++q) acc[q] = 0.0;
        if (nT > 0) {
            for (int i = rank * ICP_EVAL_THREADS + threadIdx.x; i < S; i += stride) {
                const double sx = src[i * 3], sy = src[i * 3 + 1], sz = src[i * 3 + 2];
                const double px = T_s[0] * sx + T_s[1] * sy + T_s[2] * sz + T_s[3];
                const double py = T_s[4] * sx + T_s[5] * sy + T_s[6] * sz + T_s[7];
                const double pz = T_s[8] * sx + T_s[9] * sy + T_s[10] * sz + T_s[11];
                double best, bx = 0, by = 0, bz = 0;
                int best_j, best_s;
                const int seed = ev > 0 ? match[i] : -1;  // written by this very thread in the previous evaluation
                if (in_smem) icp_search<true>(tg_s, cs, g, px, py, pz, R2, seed, best, best_j, best_s, bx, by, bz);
                else icp_search<false>(sorted + t0, cs, g, px, py, pz, R2, seed, best, best_j, best_s, bx, by, bz);
                match[i] = best_s;
                if (best_j != 0x7FFFFFFF) {
                    acc[0] += 1.0; acc[1] += best;
                    acc[2] += px; acc[3] += py; acc[4] += pz;
                    acc[5] += bx; acc[6] += by; acc[7] += bz;
                    acc[8] += px * bx;  acc[9] += px * by;  acc[10] += px * bz;
                    acc[11] += py * bx; acc[12] += py * by; acc[13] += py * bz;
                    acc[14] += pz * bx; acc[15] += pz * by; acc[16] += pz * bz;
                }
            }
        }
        // fixed-order block reduction -> this CTA's slot
#pragma unroll
        for (int q = 0; q < ICP_NSUM; ++q) {
            double v = acc[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
        }
        __syncthreads();
        if (threadIdx.x < ICP_NSUM) {
            double v = 0.0;
            for (int wv = 0; wv < ICP_EVAL_THREADS / 32; ++wv) v += red[wv][threadIdx.x];
            if (csize == 1) tot[threadIdx.x] = v;
            else __stcg(partial + rank * ICP_NSUM + threadIdx.x, v);
        }
        if (csize > 1) {
            icp_cluster_sync();  // every CTA's slot is written and visible
            if (rank == 0 && threadIdx.x < ICP_NSUM) {
                double v = 0.0;
                for (int c = 0; c < csize; ++c) v += __ldcg(partial + c * ICP_NSUM + threadIdx.x);
                tot[threadIdx.x] = v;
            }
        }
        __syncthreads();
        if (rank == 0 && threadIdx.x == 0) {
            icp_frame_update(st, tot, T_s, f, S, ev, max_iter, rel_fitness, rel_rmse, out_T, out_stats);
            __threadfence();
        }
        icp_cluster_sync();  // the new T / done of the frame are visible to every CTA of its cluster
        if (__ldcg(&st->done)) break;  // same value in every thread of the cluster
    }
}

// Multi-launch variant for FEW frames (fewer than the GPU has SMs): one launch per evaluation over (source chunks of
// 512 x frames). The (frame, chunk) CTAs of all live frames are scheduled dynamically over all SMs, so the slowly
// converging / expensive frames (bad initial pose: every query scans most of the targets) spread over the whole GPU
// instead of pinning one cluster each; the last CTA of a frame (atomic ticket) adds the chunk slots in chunk order and
// runs the update. A frame that is done returns immediately. Measured on the bench's 64 frames (32 crops x 2 initial
// poses, half of them far off): 13.5 ms against 22.7-28.9 ms for the persistent kernel; with 1 000 frames the
// persistent kernel wins (140 ms against 31 launches of 8 000 CTAs each).
__global__ void __launch_bounds__(ICP_EVAL_THREADS)
k_icp_eval(const float* __restrict__ src, int S, const int32_t* __restrict__ tgt_offsets,
           const IcpFrameGrid* __restrict__ grids, const int32_t* __restrict__ cell_start_all,
           const float4* __restrict__ sorted, int32_t* __restrict__ match_all, IcpState* __restrict__ state,
           double* __restrict__ partial_all,
           int ev, int tcap, double max_corr, int max_iter, double rel_fitness, double rel_rmse,
           double* __restrict__ out_T, double* __restrict__ out_stats) {
    extern __shared__ __align__(16) unsigned char icp_smem[];
    float4* tg_s = reinterpret_cast<float4*>(icp_smem);
    int* cs = reinterpret_cast<int*>(icp_smem + (size_t)tcap * sizeof(float4));
    __shared__ double T_s[12];
    __shared__ double red[ICP_EVAL_THREADS / 32][ICP_NSUM];
    __shared__ double tot[ICP_NSUM];
    __shared__ IcpFrameGrid g;
    __shared__ int last_s;
    const int f = blockIdx.y;
    const int nchunk = gridDim.x;
    IcpState* st = state + f;
    if (st->done) return;  // uniform for the CTA; written only by the frame's last CTA of an earlier launch
    const int t0 = tgt_offsets[f];
    const int nT = tgt_offsets[f + 1] - t0;
    if (threadIdx.x < 12) T_s[threadIdx.x] = st->T[threadIdx.x];
    if (threadIdx.x == 0) g = grids[f];
    __syncthreads();
    const int ncell = g.dims[0] * g.dims[1] * g.dims[2];
    {
        const int32_t* cell_start = cell_start_all + (int64_t)f * (ICP_CELLS + 1);
        for (int c = threadIdx.x; c <= ncell; c += blockDim.x) cs[c] = cell_start[c];
    }
    const bool in_smem = nT <= tcap;
    if (in_smem)
        for (int i = threadIdx.x; i < nT; i += blockDim.x) tg_s[i] = __ldg(sorted + t0 + i);
    __syncthreads();
    const double R2 = max_corr * max_corr;
    double acc[ICP_NSUM];
#pragma unroll
    for (int q = 0; q < ICP_NSUM; ++q) acc[q] = 0.0;
    if (nT > 0) {
        int32_t* match = match_all + (int64_t)f * S;
        for (int i = blockIdx.x * ICP_EVAL_THREADS + threadIdx.x; i < S; i += nchunk * ICP_EVAL_THREADS) {
            const double sx = src[i * 3], sy = src[i * 3 + 1], sz = src[i * 3 + 2];
            const double px = T_s[0] * sx + T_s[1] * sy + T_s[2] * sz + T_s[3];
            const double py = T_s[4] * sx + T_s[5] * sy + T_s[6] * sz + T_s[7];
            const double pz = T_s[8] * sx + T_s[9] * sy + T_s[10] * sz + T_s[11];
            double best, bx = 0, by = 0, bz = 0;
            int best_j, best_s;
            const int seed = ev > 0 ? match[i] : -1;
            if (in_smem) icp_search<true>(tg_s, cs, g, px, py, pz, R2, seed, best, best_j, best_s, bx, by, bz);
            else icp_search<false>(sorted + t0, cs, g, px, py, pz, R2, seed, best, best_j, best_s, bx, by, bz);
            match[i] = best_s;
            if (best_j != 0x7FFFFFFF) {
                acc[0] += 1.0; acc[1] += best;
                acc[2] += px; acc[3] += py; acc[4] += pz;
                acc[5] += bx; acc[6] += by; acc[7] += bz;
                acc[8] += px * bx;  acc[9] += px * by;  acc[10] += px * bz;
                acc[11] += py * bx; acc[12] += py * by; acc[13] += py * bz;
                acc[14] += pz * bx; acc[15] += pz * by; acc[16] += pz * bz;
            }
        }
    }
    // fixed-order block reduction -> this chunk's slot
#pragma unroll
    for (int q = 0; q < ICP_NSUM; ++q) {
        double v = acc[q];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][q] = v;
    }
    __syncthreads();
    double* partial = partial_all + (int64_t)f * ICP_MAX_CHUNKS * ICP_NSUM;
    if (threadIdx.x < ICP_NSUM) {
        double v = 0.0;
        for (int wv = 0; wv < ICP_EVAL_THREADS / 32; ++wv) v += red[wv][threadIdx.x];
        partial[blockIdx.x * ICP_NSUM + threadIdx.x] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int ticket = atomicAdd(&st->counter, 1u);
        last_s = (ticket == (unsigned int)(nchunk - 1)) ? 1 : 0;
    }
    __syncthreads();
    if (!last_s) return;
    // ---- last CTA of the frame: chunk slots in chunk order, convergence test, Kabsch update
    __threadfence();
    if (threadIdx.x < ICP_NSUM) {
        double v = 0.0;
        for (int c = 0; c < nchunk; ++c) v += __ldcg(partial + c * ICP_NSUM + threadIdx.x);
        tot[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0)
        icp_frame_update(st, tot, T_s, f, S, ev, max_iter, rel_fitness, rel_rmse, out_T, out_stats);
}


extern "C" int b2me_icp_p2p_batched(const float* source_xyz, int S, const float* target_xyz,
                                    const int32_t* tgt_offsets, int F, int64_t T_total, const double* init_T,
                                    double max_corr, int max_iter, double rel_fitness, double rel_rmse, double* out_T,
                                    double* out_stats, void* ws, size_t ws_bytes, int cluster_size,
                                    b2me_stream_t stream) {
    if (!source_xyz || !tgt_offsets || !init_T || !out_T || !out_stats || !ws) return B2ME_EINVAL;
    if (!target_xyz && T_total > 0) return B2ME_EINVAL;  // an empty target cloud may come with a null pointer
    if (S <= 0 || F < 0 || T_total < 0 || max_iter < 0 || !(max_corr > 0)) return B2ME_EINVAL;
    if (F == 0) return B2ME_OK;
    if (F > (1 << 20)) return B2ME_EUNSUPPORTED;
    IcpWs w = carve_icp_ws(ws, T_total, F, S);
    if (ws_bytes < w.total) return B2ME_EWORKSPACE;
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    k_icp_build_grid<<<(unsigned)F, ICP_THREADS, 0, s>>>(target_xyz, tgt_offsets, init_T, w.grids, w.state,
                                                         w.cell_start, w.cell_cursor, w.sorted);
    const size_t smem = ICP_EVAL_SMEM;
    const int tcap = (int)((smem - (size_t)(ICP_CELLS + 1) * sizeof(int)) / sizeof(float4));
    int dev = 0, sms = B2ME_NUM_SMS;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (cluster_size < 0 || cluster_size > ICP_MAX_CLUSTER || (cluster_size & (cluster_size - 1))) return B2ME_EINVAL;
    // cluster_size 0 = automatic: at least as many frames as SMs -> persistent kernel, one CTA per frame (frames are
    // scheduled dynamically, every thread takes S / 512 queries per evaluation: the heavy-tailed query costs average
    // out); fewer frames -> one launch per evaluation with (frame, chunk) CTAs balanced over all SMs.
    const bool persistent = cluster_size > 0 || F >= sms;
    if (!persistent) {
        int nchunk = (S + ICP_EVAL_THREADS - 1) / ICP_EVAL_THREADS;
        if (nchunk > ICP_MAX_CHUNKS) nchunk = ICP_MAX_CHUNKS;
        if (F > 65535) return B2ME_EUNSUPPORTED;  // gridDim.y
        if (cudaFuncSetAttribute(k_icp_eval, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return B2ME_ELAUNCH;
        const dim3 grid((unsigned)nchunk, (unsigned)F);
        // evaluation 0 uses init; then for it = 1..max_iter: update from the last evaluation, evaluate, test convergence
        for (int ev = 0; ev <= max_iter; ++ev)
            k_icp_eval<<<grid, ICP_EVAL_THREADS, smem, s>>>(source_xyz, S, tgt_offsets, w.grids, w.cell_start, w.sorted,
                                                            w.match, w.state, w.partial, ev, tcap, max_corr, max_iter,
                                                            rel_fitness, rel_rmse, out_T, out_stats);
        B2ME_CHECK_LAUNCH();
        return B2ME_OK;
    }
    // per-device attribute: set on every call (cheap) so that a process driving several GPUs works
    if (cudaFuncSetAttribute(k_icp_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return B2ME_ELAUNCH;
    const int csize = cluster_size > 0 ? cluster_size : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(F * csize), 1, 1);
    cfg.blockDim = dim3(ICP_EVAL_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, k_icp_persistent, source_xyz, S, tgt_offsets, (const IcpFrameGrid*)w.grids,
                           (const int32_t*)w.cell_start, (const float4*)w.sorted, w.match, w.state, w.partial, tcap,
                           max_corr, max_iter, rel_fitness, rel_rmse, out_T, out_stats) != cudaSuccess)
        return B2ME_ELAUNCH;
    B2ME_CHECK_LAUNCH();
    return B2ME_OK;
}
