// gather_probe.cu — developer microbenchmark: DRAM throughput of random row gathers shaped like the convolution's A
// operand loads (8 lanes read one 128-byte segment of a random 768-byte row; 4 rows per warp instruction), against a
// streaming read of the same buffer. Tells how far the K = 27 layers' 3.3-3.5 TB/s of gather traffic is from what the
// memory system gives for this access pattern.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_probe tools/gather_probe.cu && tools/gather_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// mode 0: random rows, one 128-byte segment per row visit; mode 1: random rows, all `segs` segments of the row in a
// row (what a tile does over its chunks, closely spaced in time); mode 2: streaming
template <int UNROLL>
__global__ void __launch_bounds__(256) k_gather(const uint4* __restrict__ buf, uint32_t nrows, int row_bytes, int segs,
                                                int mode, uint32_t iters, unsigned long long* sink) {
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = gtid >> 3, piece = gtid & 7;  // 8 lanes per row
    const uint32_t ngroups = (gridDim.x * blockDim.x) >> 3;
    uint32_t acc = 0;
    const uint32_t row_v4 = row_bytes / 16;
    for (uint32_t it = 0; it < iters; it += UNROLL) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const uint32_t n = (it + u) * ngroups + group;
            uint32_t row, seg;
            if (mode == 0) { const uint32_t h = hash32(n); row = h % nrows; seg = (h >> 24) % segs; }
            else if (mode == 1) { row = hash32(n / segs) % nrows; seg = n % segs; }
            else { row = (n / segs) % nrows; seg = n % segs; }
            v[u] = __ldg(buf + (size_t)row * row_v4 + seg * 8 + piece);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
    }
    if (acc == 0xdeadbeefu) atomicAdd(sink, 1ull);
}

int main() {
    const int row_bytes = 768, segs = 6;
    const size_t nrows = 8900000;  // the stride-1 map of a 32-frame batch: 6.8 GB
    uint4* buf;
    unsigned long long* sink;
    if (cudaMalloc(&buf, nrows * row_bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&sink, 8);
    cudaMemset(buf, 1, nrows * row_bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const char* names[3] = {"random row, one random 128-B segment", "random row, its six 128-B segments back to back",
                            "streaming"};
    for (int blocks_per_sm : {4, 8}) {
        for (int mode = 0; mode < 3; ++mode) {
            const int grid = 148 * blocks_per_sm;
            const uint32_t iters = 4096;
            k_gather<8><<<grid, 256>>>(buf, (uint32_t)nrows, row_bytes, segs, mode, 256, sink);  // warm-up
            cudaEventRecord(e0);
            k_gather<8><<<grid, 256>>>(buf, (uint32_t)nrows, row_bytes, segs, mode, iters, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            const double bytes = (double)grid * 256 * iters * 16;
            printf("%d blocks/SM x 256 threads, 8 loads in flight per thread, %-48s: %.2f TB/s (%s)\n", blocks_per_sm,
                   names[mode], bytes / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
