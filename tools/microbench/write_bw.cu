// write_bw.cu -- pure-store bandwidth on B200 for the store patterns of the level kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 write_bw.cu -o write_bw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// pattern A: flat coalesced 8-byte stores
__global__ void flat(double *o, size_t n) { for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o[i] = (double)i; }
// pattern B: [node][20][ncols], CTA (group of gc columns, node) writes 20 row segments, one column per thread
__global__ void rows(double *o, int ncols, int gc, int stcs)
{
    const size_t node = blockIdx.y;
    const int c0 = blockIdx.x * gc;
    for (int i = threadIdx.x; i < gc && c0 + i < ncols; i += blockDim.x) {
        double *m = o + node * 20 * ncols + c0 + i;
#pragma unroll
        for (int p = 0; p < 20; p++) { if (stcs) __stcs(m + (size_t)p * ncols, (double)p); else m[(size_t)p * ncols] = (double)p; }
    }
}
// pattern C: + byte rows
__global__ void rows_b(double *o, uint8_t *pb, int ncols, int gc)
{
    const size_t node = blockIdx.y;
    const int c0 = blockIdx.x * gc;
    for (int i = threadIdx.x; i < gc && c0 + i < ncols; i += blockDim.x) {
        double *m = o + node * 20 * ncols + c0 + i;
        uint8_t *q = pb + node * 20 * ncols + c0 + i;
#pragma unroll
        for (int p = 0; p < 20; p++) { m[(size_t)p * ncols] = (double)p; q[(size_t)p * ncols] = (uint8_t)p; }
    }
}
template <typename F> static float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); f(); cudaEventRecord(a); for (int i = 0; i < 5; i++) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}
int main() {
    const int ncols = 5000, nodes = 4096;
    const size_t n = (size_t)nodes * 20 * ncols;
    double *o; uint8_t *pb; cudaMalloc(&o, n * 8); cudaMalloc(&pb, n);
    float ms = timeit([&] { flat<<<148 * 16, 256>>>(o, n); });
    printf("flat stores          %.3f ms %.0f GB/s\n", ms, n * 8 / ms / 1e6);
    ms = timeit([&] { cudaMemsetAsync(o, 0, n * 8); });
    printf("cudaMemset           %.3f ms %.0f GB/s\n", ms, n * 8 / ms / 1e6);
    for (int gc : {128, 1250, 5000}) for (int st = 0; st < 2; st++) {
        ms = timeit([&] { rows<<<dim3((ncols + gc - 1) / gc, nodes), 128>>>(o, ncols, gc, st); });
        printf("rows gc=%4d stcs=%d   %.3f ms %.0f GB/s\n", gc, st, ms, n * 8 / ms / 1e6);
    }
    ms = timeit([&] { rows_b<<<dim3(4, nodes), 128>>>(o, pb, ncols, 1250); });
    printf("rows+bytes gc=1250   %.3f ms %.0f GB/s\n", ms, n * 9 / ms / 1e6);
    ms = timeit([&] { rows<<<dim3(4, nodes), 256>>>(o, ncols, 1250, 0); });
    printf("rows gc=1250 256thr  %.3f ms %.0f GB/s\n", ms, n * 8 / ms / 1e6);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
