// joint_scan.cu -- microbenchmark of the max-product scan variants of up_wide_kernel<20,0> (wide.cu).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo joint_scan.cu -o joint_scan
// One (node, column) update per thread (or two columns per thread): two child messages read from HBM in
// [state][column] rows, 400 candidates, message + pointer rows written.  Prints ms and updates/s per variant.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <math_constants.h>
#define K 20
#define WT 128

__device__ __forceinline__ void step_asm(double &best, int &bi, double l, double g, int x)
{
    asm("{\n\t.reg .pred p;\n\t.reg .f64 v;\n\t"
        "add.rn.f64 v, %2, %3;\n\t"
        "setp.gt.f64 p, v, %0;\n\t"
        "@p add.rn.f64 %0, v, 0d0000000000000000;\n\t"
        "@p mov.s32 %1, %4;\n\t}"
        : "+d"(best), "+r"(bi)
        : "d"(l), "d"(g), "r"(x));
}
__device__ __forceinline__ void step_c(double &best, int &bi, double l, double g, int x)
{
    const double v = l + g;
    if (v > best) { best = v; bi = x; }
}

template <int V>
__device__ __forceinline__ void scan_row(const double *Lrow, const double (&G)[K], double &best, int &bi)
{
    const double2 *row2 = reinterpret_cast<const double2 *>(Lrow);
    if (V == 2) {
        double v[K];
        int idx[K];
#pragma unroll
        for (int x2 = 0; x2 < K / 2; x2++) {
            const double2 l = row2[x2];
            v[2 * x2] = l.x + G[2 * x2];
            v[2 * x2 + 1] = l.y + G[2 * x2 + 1];
        }
#pragma unroll
        for (int x = 0; x < K; x++) idx[x] = x;
#pragma unroll
        for (int s = 1; s < K; s *= 2) {
#pragma unroll
            for (int x = 0; x + s < K; x += 2 * s) {
                const bool gt = v[x + s] > v[x];
                v[x] = gt ? v[x + s] : v[x];
                idx[x] = gt ? idx[x + s] : idx[x];
            }
        }
        best = v[0];
        bi = idx[0];
    } else {
        best = -CUDART_INF;
        bi = 0;
#pragma unroll
        for (int x2 = 0; x2 < K / 2; x2++) {
            const double2 l = row2[x2];
            if (V == 0) { step_c(best, bi, l.x, G[2 * x2], 2 * x2); step_c(best, bi, l.y, G[2 * x2 + 1], 2 * x2 + 1); }
            else { step_asm(best, bi, l.x, G[2 * x2], 2 * x2); step_asm(best, bi, l.y, G[2 * x2 + 1], 2 * x2 + 1); }
        }
    }
}

// V: 0 plain C, 1 predicated-add asm (current), 2 tournament.  NC columns per thread.  UN: unroll of p.
template <int V, int NC, int MINB>
__global__ void __launch_bounds__(WT, MINB) kern(const double *__restrict__ Lg, const double *__restrict__ m0, const double *__restrict__ m1,
                                               double *__restrict__ out, uint8_t *__restrict__ po, int n, int tiles_per_cta)
{
    __shared__ __align__(16) double Ls[K * K];
    const size_t node = blockIdx.y;
    for (int q = threadIdx.x; q < K * K; q += WT) Ls[q] = Lg[(node % 64) * K * K + q];
    __syncthreads();
    for (int t = 0; t < tiles_per_cta; t++) {
        const int col0 = (blockIdx.x * tiles_per_cta + t) * WT * NC + threadIdx.x;
        if (col0 + (NC - 1) * WT >= n) return;
        double G[NC][K];
#pragma unroll
        for (int c = 0; c < NC; c++) {
            const double *a = m0 + node * K * n + col0 + c * WT, *b = m1 + node * K * n + col0 + c * WT;
#pragma unroll
            for (int x = 0; x < K; x++) G[c][x] = a[(size_t)x * n] + b[(size_t)x * n];
        }
        double *mo = out + node * K * n + col0;
        uint8_t *pp = po + node * K * n + col0;
#pragma unroll(NC == 1 ? 4 : 2)
        for (int p = 0; p < K; p++) {
#pragma unroll
            for (int c = 0; c < NC; c++) {
                double best;
                int bi;
                scan_row<V>(Ls + p * K, G[c], best, bi);
                mo[(size_t)p * n + c * WT] = best;
                pp[(size_t)p * n + c * WT] = (uint8_t)bi;
            }
        }
    }
}

template <int V, int NC, int MINB>
static void run(const char *name, const double *L, const double *m0, const double *m1, double *out, uint8_t *po, int n, int nodes, int tpc)
{
    const int tiles = n / (WT * NC);
    dim3 grid((tiles + tpc - 1) / tpc, nodes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 2; i++) kern<V, NC, MINB><<<grid, WT>>>(L, m0, m1, out, po, n, tpc);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; i++) kern<V, NC, MINB><<<grid, WT>>>(L, m0, m1, out, po, n, tpc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= reps;
    const double upd = (double)nodes * tiles * WT * NC;
    // checksum
    std::vector<uint8_t> h(4096);
    cudaMemcpy(h.data(), po, 4096, cudaMemcpyDeviceToHost);
    unsigned cs = 0;
    for (auto b : h) cs = cs * 31 + b;
    printf("%-28s %8.3f ms  %.3e upd/s  %.1f GB/s(501B/upd)  cs=%08x err=%s\n", name, ms, upd / (ms * 1e-3), upd * 501 / (ms * 1e6), cs,
           cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv)
{
    const int n = 5120, nodes = argc > 1 ? atoi(argv[1]) : 2048;
    const size_t msz = (size_t)nodes * K * n;
    double *L, *m0, *m1, *out;
    uint8_t *po;
    cudaMalloc(&L, 64 * K * K * 8);
    cudaMalloc(&m0, msz * 8);
    cudaMalloc(&m1, msz * 8);
    cudaMalloc(&out, msz * 8);
    cudaMalloc(&po, msz);
    std::vector<double> h(msz);
    srand(1);
    for (auto &v : h) v = -20.0 * (rand() / (double)RAND_MAX);
    cudaMemcpy(m0, h.data(), msz * 8, cudaMemcpyHostToDevice);
    for (auto &v : h) v = -20.0 * (rand() / (double)RAND_MAX);
    cudaMemcpy(m1, h.data(), msz * 8, cudaMemcpyHostToDevice);
    std::vector<double> hl(64 * K * K);
    for (size_t i = 0; i < hl.size(); i++) hl[i] = ((i % (K * K)) / K == (i % K)) ? -0.1 : -3.0 - 5.0 * (rand() / (double)RAND_MAX);
    cudaMemcpy(L, hl.data(), hl.size() * 8, cudaMemcpyHostToDevice);
    for (int tpc = 1; tpc <= 8; tpc *= 8) {
        printf("tiles_per_cta %d\n", tpc);
        run<1, 1, 5>("asm pred-add (current) b5", L, m0, m1, out, po, n, nodes, tpc);
        run<0, 1, 5>("plain C b5", L, m0, m1, out, po, n, nodes, tpc);
        run<0, 1, 4>("plain C b4", L, m0, m1, out, po, n, nodes, tpc);
        run<0, 1, 6>("plain C b6", L, m0, m1, out, po, n, nodes, tpc);
        run<2, 1, 5>("tournament b5", L, m0, m1, out, po, n, nodes, tpc);
        run<2, 1, 4>("tournament b4", L, m0, m1, out, po, n, nodes, tpc);
        run<2, 1, 6>("tournament b6", L, m0, m1, out, po, n, nodes, tpc);
        run<0, 2, 3>("plain C 2col b3", L, m0, m1, out, po, n, nodes, tpc);
        run<2, 2, 3>("tournament 2col b3", L, m0, m1, out, po, n, nodes, tpc);
        run<0, 2, 4>("plain C 2col b4", L, m0, m1, out, po, n, nodes, tpc);
        run<2, 2, 4>("tournament 2col b4", L, m0, m1, out, po, n, nodes, tpc);
    }
    return 0;
}
