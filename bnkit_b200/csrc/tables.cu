// tables.cu -- device construction of P(t) = |V exp(L t) V^-1| and log P(t) for every
// (branch, rate category), once per run.   COMPILED WITH -fmad=false.
//
// The reference rebuilds this 20x20 matrix for every branch of every column
// (SubstNode.java:87 -> SubstModel.getProbs, src/bn/ctmc/SubstModel.java:303-327: 20 exp +
// 8,000 multiply-adds) and then takes 400 logs per internal node per column
// (SubstNode.java:498-518).  Only (#branches x #distinct rates) matrices exist, so they are
// built here once: one thread per matrix entry, a CTA per group of tables.
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

// One CTA builds TPB consecutive tables (table index = cat * n_nodes + node, the layout's own order); one thread per
// matrix entry.  r1 had one CTA per table: 20 of its 416 threads computed the exponentials while the rest waited, each
// thread ran one 20-deep add chain and one log at a time, and the transposed copies left as 8-byte stores 160 bytes
// apart.  Here the TPB x K exponentials are computed side by side, a thread carries the TPB tables' entries together
// (TPB independent add chains, evec read once per k for all of them), and the transposed copies are staged in shared
// memory (rows padded to K + 1: conflict-free both ways) and leave coalesced.  The arithmetic per entry -- and with it
// every bit of P and L -- is what it was.  (profiles/r2_w2_*, r2_w3_*: 0.537 -> 0.484 -> 0.382 ms on C3a.)
template <int K, int TPB, bool EXT>
__global__ void __launch_bounds__(((K * K + 31) / 32) * 32)
build_tables_kernel(ModelDev md, const double *__restrict__ times, const int32_t *__restrict__ skip,
                    const double *__restrict__ rates, int n_nodes, int n_tables, TableSet ts)
{
    constexpr int KK = K * K, PAD = K + 1;
    // EXT: a row of iexp is stored cyclically extended to 32 doubles (entry x holds column x mod K), so that the 16
    // lanes of a half-warp -- consecutive j, wrapping at K -- read 16 CONSECUTIVE doubles starting at their first
    // lane's column: one wavefront per half-warp instead of two or three when the wrap folds banks onto each other
    constexpr int ROW = EXT ? 32 : K;
    static_assert(!EXT || (K % 4 == 0 && K + 12 <= ROW), "first column of a half-warp is a multiple of 4");
    static_assert(TPB * K <= ((KK + 31) / 32) * 32, "one thread per exponential");
    __shared__ double ek[TPB][K];
    __shared__ double iexp[TPB][K * ROW];
    __shared__ double sp[TPB][K * PAD], sl[TPB][K * PAD];
    __shared__ long long s_base[TPB];  // first entry of the table, -1: nothing to build
    const int tid = threadIdx.x, t0 = blockIdx.x * TPB;
    if (tid < TPB * K) {
        const int m = tid / K, k = tid % K, T = t0 + m;
        bool live = T < n_tables;
        if (live) {
            const int node = T % n_nodes, cat = T / n_nodes;
            live = skip == nullptr || skip[node] >= 0;
            if (live) ek[m][k] = fd_exp(times[node] * rates[cat] * md.eval[k]);
        }
        if (k == 0) s_base[m] = live ? (long long)T * KK : -1;
    }
    __syncthreads();
    const bool entry = tid < KK;
    const int i = entry ? tid / K : 0, j = entry ? tid % K : 0;
    if (EXT) {
        for (int idx = tid; idx < K * ROW; idx += blockDim.x) {
            const int k = idx / ROW, x = idx % ROW;
            const double iv = md.ievec[k * K + x % K];
#pragma unroll
            for (int m = 0; m < TPB; m++) iexp[m][idx] = iv * ek[m][k];
        }
    } else if (entry) {
        const double iv = md.ievec[tid];
#pragma unroll
        for (int m = 0; m < TPB; m++) iexp[m][tid] = iv * ek[m][i];  // iexp[k][j] = ievec[k][j] * exp(t eval[k]): here k = i
    }
    __syncthreads();
    const int jx = EXT ? ((tid & ~15) % K) + (tid & 15) : j;  // where this lane finds column j in a row
    const bool want_log = ts.L || ts.LT;
    if (entry) {
        double ev[K];
#pragma unroll
        for (int k = 0; k < K; k++) ev[k] = md.evec[i * K + k];
        // the TPB tables side by side: TPB independent add chains and TPB independent logs per thread (a table that is
        // not built -- past the end, a root -- is computed from whatever its slot holds and never stored)
        double temp[TPB], lg[TPB];
#pragma unroll
        for (int m = 0; m < TPB; m++) temp[m] = 0.0;
#pragma unroll
        for (int k = 0; k < K; k++) {
#pragma unroll
            for (int m = 0; m < TPB; m++) temp[m] += ev[k] * iexp[m][k * ROW + jx];
        }
#pragma unroll
        for (int m = 0; m < TPB; m++) temp[m] = fabs(temp[m]);
        if (want_log) {   // (a lockstep form of the TPB logs, stage by stage, cost registers and lost: 0.404 against 0.382 ms)
#pragma unroll
            for (int m = 0; m < TPB; m++) lg[m] = fd_log(temp[m]);
        }
#pragma unroll
        for (int m = 0; m < TPB; m++) {
            const long long base = s_base[m];
            if (base < 0) continue;
            if (ts.P) ts.P[base + tid] = temp[m];
            sp[m][j * PAD + i] = temp[m];
            if (want_log) {
                if (ts.L) ts.L[base + tid] = lg[m];
                sl[m][j * PAD + i] = lg[m];
            }
        }
    }
    if (!ts.PT && !ts.LT) return;
    __syncthreads();
    if (entry) {  // entry tid of a transposed table is [i][j] of the table itself with i = tid / K: sp[.][i * PAD + j]
#pragma unroll
        for (int m = 0; m < TPB; m++) {
            const long long base = s_base[m];
            if (base < 0) continue;
            if (ts.PT) ts.PT[base + tid] = sp[m][i * PAD + j];
            if (ts.LT) ts.LT[base + tid] = sl[m][i * PAD + j];
        }
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
    constexpr int TPB = 4;
    const long long n_tables = (long long)n_nodes * n_cat;
    if (n_tables < 1) return;
    const unsigned grid = (unsigned)((n_tables + TPB - 1) / TPB);
    if (K == 20) build_tables_kernel<20, TPB, true><<<grid, 416, 0, st>>>(md, times, skip, rates, n_nodes, (int)n_tables, ts);
    else build_tables_kernel<4, TPB, false><<<grid, 32, 0, st>>>(md, times, skip, rates, n_nodes, (int)n_tables, ts);
}

void launch_log_prior(const double *prior, double *logprior, int K, cudaStream_t st)
{
    log_prior_kernel<<<1, 32, 0, st>>>(prior, logprior, K);
}

}  // namespace bnk
