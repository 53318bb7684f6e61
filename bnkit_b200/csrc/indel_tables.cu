// indel_tables.cu -- per-branch tables of the gap-augmented model (bn.ctmc.GapSubstModel) for the extended
// peeling of asr.IndelPeeler, built once per (model, rate) on the device.   COMPILED WITH -fmad=false.
//
// The reference evaluates, for every column, node, parent residue and child residue,
//   Math.log(getProbGapAugmented(child, x, p)) = log( P_eps(t)[p][x] * (1 - ksi(t)) )   (src/asr/IndelPeeler.java:412-419)
//   Math.log(getProbOfGap(child))              = log( (1 - ksi(t)) * gamma(t) )          (:421-429)
//   Math.log(getProbOfInsertion(v, q))         = log( ksi(t_v) * F[q] )                  (:392-400)
// with P_eps(t) = |V exp(L t) V^-1| of the (K+1) x (K+1) indel rate matrix (SubstModel.getProbs,
// src/bn/ctmc/SubstModel.java:303-327) and t = distance * rate (PhyloBN.java:158; 0 at the root).  Only one
// table per branch exists for a given (mu, lambda, rate): one CTA per node builds the residue block of
// P_eps(t) * (1 - ksi) with the arithmetic of tables.cu (k-ascending accumulation, fdlibm exp), the deletion
// term and the insertion vector.  The sweep (indel.cu) works in linear space, so no logs are taken here.
#include "fdlibm.cuh"
#include "indel.cuh"

namespace bnk {

// grid = n_nodes; block = K1 * K1 threads rounded up to a warp multiple
template <int K1>
__global__ void __launch_bounds__(((K1 * K1 + 31) / 32) * 32)
indel_tables_kernel(GapModelDev gm, const double *__restrict__ dist, const int32_t *__restrict__ parent, double rate,
                    IndelTables tb)
{
    constexpr int K = K1 - 1;
    __shared__ double iexp[K1 * K1];
    __shared__ double ek[K1];
    __shared__ double s_noins;
    const int node = blockIdx.x, tid = threadIdx.x;
    const bool is_root = parent[node] < 0;
    const double t = is_root ? 0.0 : dist[node] * rate;   // SubstNode.time
    if (tid < K1) ek[tid] = fd_exp(t * gm.eval[tid]);
    if (tid == 0) {
        double ksi = 0.0, gamma = 0.0;
        if (!(gm.mu == 0 && gm.lambda == 0)) {   // ksiT / gammaT (:402-410, :431-441)
            const double total = gm.mu + gm.lambda;
            const double indelProp = 1 - fd_exp(-(total * t));
            ksi = (gm.lambda / total) * indelProp;
            gamma = (gm.mu / total) * indelProp;
        }
        const double noins = 1 - ksi;
        s_noins = noins;
        tb.del[node] = noins * gamma;
        for (int q = 0; q < K; q++) tb.ins[(size_t)node * K + q] = ksi * gm.F[q];
    }
    __syncthreads();
    if (is_root) return;   // no branch above the root: its table is never read
    if (tid < K1 * K1) iexp[tid] = gm.ievec[tid] * ek[tid / K1];
    __syncthreads();
    if (tid >= K1 * K1) return;
    const int i = tid / K1, j = tid % K1;
    if (i >= K || j >= K) return;
    double temp = 0.0;
#pragma unroll
    for (int k = 0; k < K1; k++) temp += gm.evec[i * K1 + k] * iexp[k * K1 + j];
    tb.A[((size_t)node * K + j) * K + i] = fabs(temp) * s_noins;   // stored child-state-major: A[x][p], p = parent residue
}

// the full (K+1)^2 matrix at one time, for tests (bnk_indel_probs)
template <int K1>
__global__ void indel_probs_kernel(GapModelDev gm, double t, double *__restrict__ P)
{
    __shared__ double iexp[K1 * K1];
    __shared__ double ek[K1];
    const int tid = threadIdx.x;
    if (tid < K1) ek[tid] = fd_exp(t * gm.eval[tid]);
    __syncthreads();
    if (tid < K1 * K1) iexp[tid] = gm.ievec[tid] * ek[tid / K1];
    __syncthreads();
    if (tid >= K1 * K1) return;
    const int i = tid / K1, j = tid % K1;
    double temp = 0.0;
    for (int k = 0; k < K1; k++) temp += gm.evec[i * K1 + k] * iexp[k * K1 + j];
    P[tid] = fabs(temp);
}

void launch_indel_tables(int K1, const GapModelDev &gm, const double *dist, const int32_t *parent, double rate, int n_nodes,
                         const IndelTables &tb, cudaStream_t st)
{
    if (K1 == 21) indel_tables_kernel<21><<<n_nodes, 448, 0, st>>>(gm, dist, parent, rate, tb);
    else indel_tables_kernel<5><<<n_nodes, 32, 0, st>>>(gm, dist, parent, rate, tb);
}

void launch_indel_probs(int K1, const GapModelDev &gm, double t, double *P, cudaStream_t st)
{
    if (K1 == 21) indel_probs_kernel<21><<<1, 448, 0, st>>>(gm, t, P);
    else indel_probs_kernel<5><<<1, 32, 0, st>>>(gm, t, P);
}

}  // namespace bnk
