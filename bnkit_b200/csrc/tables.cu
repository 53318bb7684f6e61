// tables.cu -- device construction of P(t) = |V exp(L t) V^-1| and log P(t) for every
// (branch, rate category), once per run.   COMPILED WITH -fmad=false.
//
// The reference rebuilds this 20x20 matrix for every branch of every column
// (SubstNode.java:87 -> SubstModel.getProbs, src/bn/ctmc/SubstModel.java:303-327: 20 exp +
// 8,000 multiply-adds) and then takes 400 logs per internal node per column
// (SubstNode.java:498-518).  Only (#branches x #distinct rates) matrices exist, so they are
// built here once: one CTA per (node, category), one thread per matrix entry.
//
// Bit-level contract (what makes joint reconstruction reproducible against the oracle):
//   iexp[k][j] = ievec[k][j] * exp(t * eval[k])
//   P[i][j]    = abs( sum_{k ascending} evec[i][k] * iexp[k][j] )   separate mul and add
//   L[i][j]    = log(P[i][j]),  log(0) = -inf   (AbstractFactor.java:1088-1092)
// with exp/log = fdlibm's e_exp.c / e_log.c (the algorithm java.lang.StrictMath is specified
// to be; java.lang.Math may differ from it by one ulp).  fdlibm is IEEE-754 arithmetic only,
// so with contraction off this kernel is bit-identical to any other faithful evaluation.
#include "device_common.cuh"
#include "fdlibm.cuh"

namespace bnk {

// grid = (n_nodes, n_cat); block = K*K threads (rounded up to a warp multiple).
// times[node] * rates[cat] is the branch time (PhyloBN.java:158).  Nodes with skip[node] < 0
// (tree roots: no branch above them) are left untouched.
template <int K>
__global__ void __launch_bounds__(((K * K + 31) / 32) * 32)
build_tables_kernel(ModelDev md, const double *__restrict__ times, const int32_t *__restrict__ skip,
                    const double *__restrict__ rates, int n_nodes, TableSet ts)
{
    __shared__ double iexp[K * K];
    __shared__ double ek[K];
    const int node = blockIdx.x, cat = blockIdx.y, tid = threadIdx.x;
    if (skip != nullptr && skip[node] < 0) return;
    const double t = times[node] * rates[cat];
    if (tid < K) ek[tid] = fd_exp(t * md.eval[tid]);
    __syncthreads();
    if (tid < K * K) iexp[tid] = md.ievec[tid] * ek[tid / K];
    __syncthreads();
    if (tid >= K * K) return;
    const int i = tid / K, j = tid % K;
    double temp = 0.0;
#pragma unroll
    for (int k = 0; k < K; k++) temp += md.evec[i * K + k] * iexp[k * K + j];
    const double p = fabs(temp);
    const size_t base = ((size_t)cat * n_nodes + node) * (K * K);
    if (ts.P) ts.P[base + i * K + j] = p;
    if (ts.PT) ts.PT[base + j * K + i] = p;
    if (ts.L || ts.LT) {
        const double l = fd_log(p);
        if (ts.L) ts.L[base + i * K + j] = l;
        if (ts.LT) ts.LT[base + j * K + i] = l;
    }
}

// log of the root prior F / sum(F) (SubstNode.java:592-607)
__global__ void log_prior_kernel(const double *__restrict__ prior, double *__restrict__ logprior, int K)
{
    if ((int)threadIdx.x < K) logprior[threadIdx.x] = fd_log(prior[threadIdx.x]);
}

void launch_build_tables(int K, const ModelDev &md, const double *times, const int32_t *skip,
                         const double *rates, int n_nodes, int n_cat, const TableSet &ts, cudaStream_t st)
{
    dim3 grid(n_nodes, n_cat);
    if (K == 20) build_tables_kernel<20><<<grid, 416, 0, st>>>(md, times, skip, rates, n_nodes, ts);
    else build_tables_kernel<4><<<grid, 32, 0, st>>>(md, times, skip, rates, n_nodes, ts);
}

void launch_log_prior(const double *prior, double *logprior, int K, cudaStream_t st)
{
    log_prior_kernel<<<1, 32, 0, st>>>(prior, logprior, K);
}

}  // namespace bnk
