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

namespace bnk {

__device__ __forceinline__ double with_hi(double x, int hi) { return __hiloint2double(hi, __double2loint(x)); }

__device__ double fd_exp(double x)
{
    const double huge = 1.0e+300, twom1000 = 9.33263618503218878990e-302,
                 o_threshold = 7.09782712893383973096e+02, u_threshold = -7.45133219101941108420e+02,
                 ln2HI = 6.93147180369123816490e-01, ln2LO = 1.90821492927058770002e-10,
                 invln2 = 1.44269504088896338700e+00, P1 = 1.66666666666666019037e-01,
                 P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
                 P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
    double y, hi = 0.0, lo = 0.0, c, t;
    int k = 0;
    unsigned hx = (unsigned)__double2hiint(x);
    const int xsb = (int)((hx >> 31) & 1u);
    hx &= 0x7fffffffu;
    if (hx >= 0x40862E42u) {
        if (hx >= 0x7ff00000u) {
            if (((hx & 0xfffffu) | (unsigned)__double2loint(x)) != 0u) return x + x;
            return xsb == 0 ? x : 0.0;
        }
        if (x > o_threshold) return huge * huge;
        if (x < u_threshold) return twom1000 * twom1000;
    }
    if (hx > 0x3fd62e42u) {
        if (hx < 0x3FF0A2B2u) {
            hi = xsb ? x + ln2HI : x - ln2HI;
            lo = xsb ? -ln2LO : ln2LO;
            k = 1 - xsb - xsb;
        } else {
            k = (int)(invln2 * x + (xsb ? -0.5 : 0.5));
            t = (double)k;
            hi = x - t * ln2HI;
            lo = t * ln2LO;
        }
        x = hi - lo;
    } else if (hx < 0x3e300000u) {
        if (huge + x > 1.0) return 1.0 + x;
    }
    t = x * x;
    c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
    if (k == 0) return 1.0 - ((x * c) / (c - 2.0) - x);
    y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
    if (k >= -1021) return with_hi(y, __double2hiint(y) + (k << 20));
    y = with_hi(y, __double2hiint(y) + ((k + 1000) << 20));
    return y * twom1000;
}

__device__ double fd_log(double x)
{
    const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
                 two54 = 1.80143985094819840000e+16, Lg1 = 6.666666666666735130e-01,
                 Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
                 Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01,
                 Lg6 = 1.531383769920937332e-01, Lg7 = 1.479819860511658591e-01;
    double hfsq, f, s, z, R, w, t1, t2, dk;
    int k = 0, i, j;
    int hx = __double2hiint(x);
    const unsigned lx = (unsigned)__double2loint(x);
    if (hx < 0x00100000) {
        if (((hx & 0x7fffffff) | (int)lx) == 0) return -CUDART_INF;
        if (hx < 0) return CUDART_NAN;
        k -= 54;
        x *= two54;
        hx = __double2hiint(x);
    }
    if (hx >= 0x7ff00000) return x + x;
    k += (hx >> 20) - 1023;
    hx &= 0x000fffff;
    i = (hx + 0x95f64) & 0x100000;
    x = with_hi(x, hx | (i ^ 0x3ff00000));
    k += (i >> 20);
    f = x - 1.0;
    if ((0x000fffff & (2 + hx)) < 3) {
        if (f == 0.0) {
            if (k == 0) return 0.0;
            dk = (double)k;
            return dk * ln2_hi + dk * ln2_lo;
        }
        R = f * f * (0.5 - 0.33333333333333333 * f);
        if (k == 0) return f - R;
        dk = (double)k;
        return dk * ln2_hi - ((R - dk * ln2_lo) - f);
    }
    s = f / (2.0 + f);
    dk = (double)k;
    z = s * s;
    i = hx - 0x6147a;
    w = z * z;
    j = 0x6b851 - hx;
    t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
    t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
    i |= j;
    R = t2 + t1;
    if (i > 0) {
        hfsq = 0.5 * f * f;
        if (k == 0) return f - (hfsq - s * (hfsq + R));
        return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
    }
    if (k == 0) return f - s * (f - R);
    return dk * ln2_hi - ((s * (f - R) - dk * ln2_lo) - f);
}

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
