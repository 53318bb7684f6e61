// layout_bw.cu -- does the [node][state][column] message layout cost DRAM efficiency?  (r2)
// A level kernel's CTA (node, 128-column tile) reads 20 row segments of 1 KB that lie 40,000 bytes apart (and
// writes as many).  This microbenchmark moves the same bytes with no arithmetic in two layouts:
//   rows : msg[node][state][ncols]            (what the library uses)
//   tiles: msg[node][tile][state][128]        (a CTA's 20 KB are contiguous)
// each thread reads R vectors of 20 doubles and writes W, like up- / down-sweep kernels do.  shift = 2, 5: the rows
// layout with segments that start 16 / 40 bytes off a 128-byte line (what a rate category of odd size does to
// every tile behind it).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 layout_bw.cu -o layout_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr int K = 20, WT = 128;
template <int R, int W, bool TILED>
__global__ void __launch_bounds__(WT) mv(const double *__restrict__ in, double *__restrict__ out, int ncols, int n_tiles, int n_nodes, int shift)
{
    const int node = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
    const int col = tile * WT + tid + shift;  // shift != 0: row segments that start off a sector / line boundary
    if (col >= ncols) return;
    const size_t node_stride = TILED ? (size_t)n_tiles * K * WT : (size_t)K * ncols;
    const size_t xs = TILED ? WT : ncols;
    const size_t off = TILED ? (size_t)tile * K * WT + tid : col;
    double acc[K];
#pragma unroll
    for (int x = 0; x < K; x++) acc[x] = 0.0;
#pragma unroll
    for (int r = 0; r < R; r++) {
        const double *p = in + ((size_t)(node + r * n_nodes)) * node_stride + off;
#pragma unroll
        for (int x = 0; x < K; x++) acc[x] += p[x * xs];
    }
#pragma unroll
    for (int w = 0; w < W; w++) {
        double *p = out + ((size_t)(node + w * n_nodes)) * node_stride + off;
#pragma unroll
        for (int x = 0; x < K; x++) p[x * xs] = acc[x] + w;
    }
}
template <typename F> static float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaEventRecord(a); for (int i = 0; i < 5; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
    const int ncols = 5000, nodes = 1808, n_tiles = (ncols + WT - 1) / WT;
    const size_t per_node = (size_t)n_tiles * K * WT;  // the larger of the two layouts
    double *in, *out; cudaMalloc(&in, per_node * nodes * 3 * 8); cudaMalloc(&out, per_node * nodes * 3 * 8);
    cudaMemset(in, 0, per_node * nodes * 3 * 8);
    dim3 grid(n_tiles, nodes);
    const double gb = (double)nodes * ncols * K * 8 / 1e9;
    float ms;
#define RUN(R, W, T) for (int shift : {0, 2, 5}) { if (T && shift) continue; ms = timeit([&] { mv<R, W, T><<<grid, WT>>>(in, out, ncols, n_tiles, nodes, shift); }); \
    printf("read %d write %d %-5s shift %d  %.3f ms  %.0f GB/s\n", R, W, T ? "tiles" : "rows", shift, ms, gb * (R + W) / ms * 1e3); }
    RUN(3, 3, false) RUN(3, 3, true) RUN(2, 1, false) RUN(2, 1, true) RUN(1, 1, false) RUN(1, 1, true) RUN(3, 0, false) RUN(3, 0, true) RUN(0, 3, false) RUN(0, 3, true)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
