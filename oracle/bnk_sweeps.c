/*
 * bnk_sweeps.c -- CPU ORACLE, part 2: tree-specialised restatement of what
 * bnkit's bucket elimination computes on a phylogenetic BN, batched over
 * columns and threaded, plus the brute-force enumerator the reference's own
 * tests use.  Test infrastructure only (see bnk_oracle.h); -ffp-contract=off.
 *
 * Follows (reference file:line):
 *   src/dat/phylo/PhyloBN.java:143-164       one SubstNode per *connected* node, t = dist * rate,
 *                                            component roots get the prior
 *   src/dat/phylo/IdxTree.java:471-503       edges survive iff both ends survive (no distance merging)
 *   src/bn/ctmc/SubstNode.java:471-609       evidence cases 1-4 + root prior (log P, log 0 = -inf)
 *   src/bn/alg/VarElim.java:187-334          bucket order = leaf-to-root message passing on a tree
 *   src/bn/factor/Factorize.java:351-373     >2 factors: cheap single-variable factors multiplied first
 *   src/bn/factor/Factorize.java:1414-1543   max-out: ascending scan, strict '>', init (-inf, index 0)
 *   src/bn/factor/Factorize.java:1223-1401   sum-out: sequential pairwise logSumOfLogs, ascending
 *   src/bn/factor/Factorize.java:1668-1717   getNormal; src/bn/alg/CGTable.java:127-190, :844-863 query
 *   src/bn/prob/EnumDistrib.java:363-371     final re-normalisation of a posterior
 *   src/asr/MaxLhoodJoint.java:119-170       observed values pass through; unconnected nodes untouched
 *   test/asr/MaxLhoodJointTest.java:318-350  brute-force enumeration (first strict maximum)
 *
 * Association order of the log-space additions.  A bucket's factors are multiplied along the tree
 * Factorize.getProductTree builds (Factorize.java:331-375): repeatedly the FIRST cheapest pair (i < j)
 * of the pool is replaced by its product, appended at the END of the pool.  Single-variable factors
 * (cost 20) therefore combine before the 400-entry CPT factor, and among themselves by "queue pairing":
 *     Q = [f1, f2, ..., fm];  while |Q| > 1:  a = pop_front, b = pop_front, push_back(a + b)
 * The bucket list holds the factors made directly from nodes first (VarElim.java:211-236: the prior or
 * the observed-parent row of the node itself, and the evidence factors of its OBSERVED children), then
 * the messages of its unobserved children as their buckets finish (:306-316).  Within each group the
 * reference's order depends on identity hash codes (HashMap / HashSet iteration); the canonical order
 * used here -- one of its possible outcomes -- is base, observed children by index, unobserved children
 * by index.  Then:   non-root, parent unobserved:  L_v[p][x] + Q(x)      root-like:  Q(x) with base first.
 * For nodes with at most two children every order gives the same bits (fp addition commutes).
 */
#include "bnk_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---------- small parallel-for ---------- */
typedef void (*pf_body)(void *ctx, int tid, long lo, long hi);
typedef struct { pf_body f; void *ctx; int tid; long lo, hi; } pf_arg;
static void *pf_tramp(void *a_) { pf_arg *a = (pf_arg *)a_; a->f(a->ctx, a->tid, a->lo, a->hi); return NULL; }
static void parallel_for(long n, int nthreads, pf_body f, void *ctx)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (n < nthreads) nthreads = n > 0 ? (int)n : 1;
    if (nthreads == 1) { f(ctx, 0, 0, n); return; }
    pthread_t th[256];
    pf_arg arg[256];
    for (int t = 0; t < nthreads; t++) {
        arg[t].f = f; arg[t].ctx = ctx; arg[t].tid = t;
        arg[t].lo = n * t / nthreads; arg[t].hi = n * (t + 1) / nthreads;
        pthread_create(&th[t], NULL, pf_tramp, &arg[t]);
    }
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
}

/* ---------- shared batch context ---------- */
typedef struct {
    const om_model *m;
    int K, n, ncols;
    const int *parent;
    const double *dist;
    const uint8_t *present, *obs;
    int *first, *kids;        /* children CSR, index order */
    double logprior[OM_MAXK];
    double rate;              /* current rate group */
    double *logP;             /* [n][K*K] for the current rate */
    const int *cols;          /* columns of the current rate group */
    uint8_t *out_state; double *out_logmax;   /* joint */
    double *out_post, *out_logl;              /* marginal */
} sweep_ctx;

static void build_children(int n, const int *parent, int **first_, int **kids_)
{
    int *cnt = (int *)calloc(n + 1, sizeof(int));
    int *first = (int *)malloc(sizeof(int) * (n + 1));
    int *kids = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int i = 0; i < n; i++) if (parent[i] >= 0) cnt[parent[i]]++;
    first[0] = 0;
    for (int i = 0; i < n; i++) first[i + 1] = first[i] + cnt[i];
    memset(cnt, 0, sizeof(int) * (n + 1));
    for (int i = 0; i < n; i++) if (parent[i] >= 0) kids[first[parent[i]] + cnt[parent[i]]++] = i;
    free(cnt);
    *first_ = first; *kids_ = kids;
}

static void table_body(void *c_, int tid, long lo, long hi)
{
    sweep_ctx *c = (sweep_ctx *)c_;
    int K = c->K;
    double P[OM_MAXK * OM_MAXK];
    (void)tid;
    for (long v = lo; v < hi; v++) {
        if (c->parent[v] < 0) continue;
        om_getprobs(c->m, c->dist[v] * c->rate, P);          /* PhyloBN.java:158 */
        for (int i = 0; i < K * K; i++) c->logP[v * K * K + i] = om_log(P[i]); /* AbstractFactor.java:1088-1092 */
    }
}

#define PRESENT(c, v, col) ((c)->present == NULL || (c)->present[(size_t)(v) * (c)->ncols + (col)])
#define OBS(c, v, col) ((c)->obs[(size_t)(v) * (c)->ncols + (col)])

/* queue pairing of m values (m >= 1): the association Factorize.getProductTree produces for m
 * equal-cost factors */
static double queue_pair(double *q, int m)
{
    int head = 0, tail = m;
    while (tail - head > 1) {
        const double a = q[head++], b = q[head++];
        q[tail++] = a + b;
    }
    return q[head];
}

/* the single-variable factors of node v's bucket for state x of v, canonical order; returns their count */
static int gather_factors(const sweep_ctx *c, int v, int col, const double *M, int x, int rootlike, const double *base,
                          double *q)
{
    const int K = c->K, KK = K * K;
    int m = 0;
    if (rootlike) q[m++] = base[x];
    for (int pass = 0; pass < 2; pass++) {
        for (int e = c->first[v]; e < c->first[v + 1]; e++) {
            const int u = c->kids[e];
            if (!PRESENT(c, u, col)) continue;
            const uint8_t ou = OBS(c, u, col);
            if ((ou != OM_NONE) != (pass == 0)) continue;
            q[m++] = (ou != OM_NONE) ? c->logP[(size_t)u * KK + x * K + ou] : M[(size_t)u * K + x];
        }
    }
    return m;
}

/* ---------- joint (max-product) ---------- */
static void joint_body(void *c_, int tid, long lo, long hi)
{
    sweep_ctx *c = (sweep_ctx *)c_;
    const int K = c->K, n = c->n, KK = K * K;
    double *M = (double *)malloc(sizeof(double) * (size_t)n * K);
    uint8_t *ptr = (uint8_t *)malloc((size_t)n * K);
    uint8_t *rstate = (uint8_t *)malloc(n);
    uint8_t *kind = (uint8_t *)malloc(n); /* 0 skip, 1 message node, 2 root-like, 3 observed */
    int maxdeg = 0;
    for (int v = 0; v < n; v++) if (c->first[v + 1] - c->first[v] > maxdeg) maxdeg = c->first[v + 1] - c->first[v];
    double *q = (double *)malloc(sizeof(double) * (2 * (maxdeg + 1) + 1));
    (void)tid;
    for (long ci = lo; ci < hi; ci++) {
        const int col = c->cols[ci];
        double logmax = 0.0;
        for (int v = n - 1; v >= 0; v--) {
            kind[v] = 0;
            if (!PRESENT(c, v, col)) continue;
            int pe = c->parent[v];
            if (pe >= 0 && !PRESENT(c, pe, col)) pe = -1;
            int nk = 0;
            double G[OM_MAXK];
            const uint8_t ov = OBS(c, v, col);
            for (int e = c->first[v]; e < c->first[v + 1]; e++) {
                int u = c->kids[e];
                if (!PRESENT(c, u, col)) continue;
                nk++;
            }
            if (pe < 0 && nk == 0) continue;               /* not connected: no BN node (PhyloBN.java:149) */
            const double *Lv = c->logP + (size_t)v * KK;
            if (ov != OM_NONE) {                            /* observed: separates the tree */
                kind[v] = 3;
                if (pe < 0) logmax += c->logprior[ov];
                else if (OBS(c, pe, col) != OM_NONE) logmax += Lv[OBS(c, pe, col) * K + ov];
                continue;
            }
            const int rootlike = (pe < 0) || (OBS(c, pe, col) != OM_NONE);
            {
                const double *base = !rootlike ? NULL : (pe < 0) ? c->logprior : Lv + (size_t)OBS(c, pe, col) * K;
                for (int x = 0; x < K; x++) {
                    const int m = gather_factors(c, v, col, M, x, rootlike, base, q);
                    G[x] = m > 0 ? queue_pair(q, m) : 0.0;
                }
            }
            if (rootlike) {
                double max = -INFINITY; int maxidx = 0;
                for (int x = 0; x < K; x++) if (G[x] > max) { max = G[x]; maxidx = x; }
                rstate[v] = (uint8_t)maxidx;
                logmax += max;
                kind[v] = 2;
            } else {
                for (int p = 0; p < K; p++) {
                    double max = -INFINITY; int maxidx = 0;
                    for (int x = 0; x < K; x++) {
                        double val = (nk > 0) ? Lv[p * K + x] + G[x] : Lv[p * K + x];
                        if (val > max) { max = val; maxidx = x; }
                    }
                    M[(size_t)v * K + p] = max;
                    ptr[(size_t)v * K + p] = (uint8_t)maxidx;
                }
                kind[v] = 1;
            }
        }
        for (int v = 0; v < n; v++) {
            uint8_t s;
            switch (kind[v]) {
            case 1: s = ptr[(size_t)v * K + c->out_state[(size_t)c->parent[v] * c->ncols + col]]; break;
            case 2: s = rstate[v]; break;
            case 3: s = OBS(c, v, col); break;
            default: s = PRESENT(c, v, col) ? OBS(c, v, col) : OM_NONE; break;
            }
            c->out_state[(size_t)v * c->ncols + col] = s;
        }
        if (c->out_logmax) c->out_logmax[col] = logmax;
    }
    free(M); free(ptr); free(rstate); free(kind); free(q);
}

/* ---------- marginal (sum-product, log space) ---------- */
static double lse_seq(const double *v, int K)
{
    double s = v[0];
    for (int x = 1; x < K; x++) s = om_logsum(s, v[x]);
    return s;
}

static void marginal_body(void *c_, int tid, long lo, long hi)
{
    sweep_ctx *c = (sweep_ctx *)c_;
    const int K = c->K, n = c->n, KK = K * K;
    double *IN = (double *)malloc(sizeof(double) * (size_t)n * K);  /* message v -> parent, indexed by parent state */
    double *B = (double *)malloc(sizeof(double) * (size_t)n * K);   /* G_v[x]: product of child messages (log) */
    double *D = (double *)malloc(sizeof(double) * (size_t)n * K);   /* message from above into v (log), incl. prior */
    uint8_t *kind = (uint8_t *)malloc(n);
    (void)tid;
    for (long ci = lo; ci < hi; ci++) {
        const int col = c->cols[ci];
        double logl = 0.0;
        for (int v = n - 1; v >= 0; v--) {
            kind[v] = 0;
            if (!PRESENT(c, v, col)) continue;
            int pe = c->parent[v];
            if (pe >= 0 && !PRESENT(c, pe, col)) pe = -1;
            int nk = 0;
            for (int e = c->first[v]; e < c->first[v + 1]; e++)
                if (PRESENT(c, c->kids[e], col)) nk++;
            if (pe < 0 && nk == 0) continue;
            const double *Lv = c->logP + (size_t)v * KK;
            const uint8_t ov = OBS(c, v, col);
            if (ov != OM_NONE) {
                kind[v] = 3;
                if (pe < 0) logl += c->logprior[ov];
                else if (OBS(c, pe, col) != OM_NONE) logl += Lv[OBS(c, pe, col) * K + ov];
                continue;
            }
            double *G = B + (size_t)v * K;
            for (int x = 0; x < K; x++) G[x] = 0.0;
            int firstk = 1;
            for (int e = c->first[v]; e < c->first[v + 1]; e++) {
                int u = c->kids[e];
                if (!PRESENT(c, u, col)) continue;
                const uint8_t ou = OBS(c, u, col);
                const double *Lu = c->logP + (size_t)u * KK;
                for (int x = 0; x < K; x++) {
                    double g = (ou != OM_NONE) ? Lu[x * K + ou] : IN[(size_t)u * K + x];
                    G[x] = firstk ? g : G[x] + g;
                }
                firstk = 0;
            }
            const int rootlike = (pe < 0) || (OBS(c, pe, col) != OM_NONE);
            if (rootlike) {
                const double *base = (pe < 0) ? c->logprior : Lv + (size_t)OBS(c, pe, col) * K;
                double tmp[OM_MAXK];
                for (int x = 0; x < K; x++) {
                    D[(size_t)v * K + x] = base[x];
                    tmp[x] = base[x] + G[x];
                }
                logl += lse_seq(tmp, K);
                kind[v] = 2;
            } else {
                double tmp[OM_MAXK];
                for (int p = 0; p < K; p++) {
                    for (int x = 0; x < K; x++) tmp[x] = Lv[p * K + x] + G[x];
                    IN[(size_t)v * K + p] = lse_seq(tmp, K);
                }
                kind[v] = 1;
            }
        }
        if (c->out_logl) c->out_logl[col] = logl;
        if (!c->out_post) continue;
        for (int v = 0; v < n; v++) {
            double *post = c->out_post + ((size_t)v * c->ncols + col) * K;
            if (kind[v] != 1 && kind[v] != 2) {
                for (int x = 0; x < K; x++) post[x] = NAN;
                continue;
            }
            if (kind[v] == 1) {
                /* D_v[x] = LSE_xp( L_v[xp][x] + (D_p[xp] + sum_{siblings} g_w[xp]) ) */
                const int p = c->parent[v];
                const double *Lv = c->logP + (size_t)v * KK;
                double T[OM_MAXK], tmp[OM_MAXK];
                for (int xp = 0; xp < K; xp++) T[xp] = D[(size_t)p * K + xp];
                for (int e = c->first[p]; e < c->first[p + 1]; e++) {
                    int w = c->kids[e];
                    if (w == v || !PRESENT(c, w, col)) continue;
                    const uint8_t ow = OBS(c, w, col);
                    const double *Lw = c->logP + (size_t)w * KK;
                    for (int xp = 0; xp < K; xp++)
                        T[xp] += (ow != OM_NONE) ? Lw[xp * K + ow] : IN[(size_t)w * K + xp];
                }
                for (int x = 0; x < K; x++) {
                    for (int xp = 0; xp < K; xp++) tmp[xp] = Lv[xp * K + x] + T[xp];
                    D[(size_t)v * K + x] = lse_seq(tmp, K);
                }
            }
            /* getNormal (Factorize.java:1668-1717) then CGTable.query + EnumDistrib.normalise */
            double y[OM_MAXK], logmax = -INFINITY, sum = 0.0, tot = 0.0;
            for (int x = 0; x < K; x++) {
                y[x] = D[(size_t)v * K + x] + B[(size_t)v * K + x];
                if (y[x] > logmax) logmax = y[x];
            }
            for (int x = 0; x < K; x++) {
                y[x] = y[x] - logmax;
                sum = (x == 0) ? y[x] : om_logsum(sum, y[x]);
            }
            for (int x = 0; x < K; x++) { post[x] = om_exp(y[x] - sum); tot += post[x]; }
            for (int x = 0; x < K; x++) post[x] /= tot;
        }
    }
    free(IN); free(B); free(D); free(kind);
}

/* ---------- drivers: group columns by rate, build tables once per distinct rate ---------- */
typedef struct { double r; int col; } rc_t;
static int rc_cmp(const void *a, const void *b)
{
    const rc_t *x = (const rc_t *)a, *y = (const rc_t *)b;
    if (x->r < y->r) return -1;
    if (x->r > y->r) return 1;
    return (x->col > y->col) - (x->col < y->col);
}

static int run_batch(sweep_ctx *c, const double *rates, int nthreads, pf_body body)
{
    const int K = c->K, n = c->n, ncols = c->ncols;
    rc_t *rc = (rc_t *)malloc(sizeof(rc_t) * (ncols > 0 ? ncols : 1));
    int *cols = (int *)malloc(sizeof(int) * (ncols > 0 ? ncols : 1));
    for (int i = 0; i < ncols; i++) { rc[i].r = rates ? rates[i] : 1.0; rc[i].col = i; }
    qsort(rc, ncols, sizeof(rc_t), rc_cmp);
    for (int i = 0; i < ncols; i++) cols[i] = rc[i].col;
    c->logP = (double *)malloc(sizeof(double) * (size_t)n * K * K);
    build_children(n, c->parent, &c->first, &c->kids);
    {
        double pr[OM_MAXK];
        om_model_get(c->m, NULL, NULL, NULL, NULL, NULL, pr);
        for (int x = 0; x < K; x++) c->logprior[x] = om_log(pr[x]);
    }
    for (int a = 0; a < ncols;) {
        int b = a;
        while (b < ncols && rc[b].r == rc[a].r) b++;
        c->rate = rc[a].r;
        parallel_for(n, nthreads, table_body, c);
        c->cols = cols + a;
        parallel_for(b - a, nthreads, body, c);
        a = b;
    }
    free(rc); free(cols); free(c->logP); free(c->first); free(c->kids);
    return 0;
}

int om_fast_joint(const om_model *m, int n, const int *parent, const double *dist, int ncols,
                  const uint8_t *present, const uint8_t *obs, const double *rates,
                  uint8_t *out_state, double *out_logmax, int nthreads)
{
    sweep_ctx c;
    memset(&c, 0, sizeof(c));
    c.m = m; c.K = om_model_nstates(m); c.n = n; c.ncols = ncols; c.parent = parent; c.dist = dist;
    c.present = present; c.obs = obs; c.out_state = out_state; c.out_logmax = out_logmax;
    return run_batch(&c, rates, nthreads, joint_body);
}

int om_fast_marginal(const om_model *m, int n, const int *parent, const double *dist, int ncols,
                     const uint8_t *present, const uint8_t *obs, const double *rates,
                     double *out_post, double *out_logl, int nthreads)
{
    sweep_ctx c;
    memset(&c, 0, sizeof(c));
    c.m = m; c.K = om_model_nstates(m); c.n = n; c.ncols = ncols; c.parent = parent; c.dist = dist;
    c.present = present; c.obs = obs; c.out_post = out_post; c.out_logl = out_logl;
    return run_batch(&c, rates, nthreads, marginal_body);
}

/* ---------- brute force (MaxLhoodJointTest.java:318-350, MaxLhoodMarginalTest.java:262-300) ---------- */
int om_brute(const om_model *m, int n, const int *parent, const double *dist,
             const uint8_t *present, const uint8_t *obs, double rate, long max_assign,
             uint8_t *best_state, double *best_p, double *post, double *total_p)
{
    const int K = om_model_nstates(m), KK = K * K;
    double prior[OM_MAXK];
    double *P = (double *)malloc(sizeof(double) * (size_t)n * KK);
    int *pe = (int *)malloc(sizeof(int) * n);
    int *freev = (int *)malloc(sizeof(int) * n);
    uint8_t *st = (uint8_t *)malloc(n), *conn = (uint8_t *)calloc(n, 1);
    int nfree = 0, rc = 0;
    long total = 1;
    om_model_get(m, NULL, NULL, NULL, NULL, NULL, prior);
#define PR(v) (present == NULL || present[v])
    for (int v = 0; v < n; v++) {
        pe[v] = (parent[v] >= 0 && PR(v) && PR(parent[v])) ? parent[v] : -1;
        if (pe[v] >= 0) { conn[v] = 1; conn[pe[v]] = 1; }
    }
    for (int v = 0; v < n; v++) {
        st[v] = OM_NONE;
        if (!PR(v)) { conn[v] = 0; continue; }
        if (!conn[v]) { st[v] = obs[v]; continue; }
        if (pe[v] >= 0) om_getprobs(m, dist[v] * rate, P + (size_t)v * KK);
        if (obs[v] != OM_NONE) st[v] = obs[v];
        else {
            freev[nfree++] = v;
            if (total > max_assign / K) { rc = -1; goto done; }
            total *= K;
        }
    }
    if (post) for (int i = 0; i < n * K; i++) post[i] = 0.0;
    for (int i = 0; i < nfree; i++) st[freev[i]] = 0;
    {
        double best = -1.0, tot = 0.0;
        for (long a = 0; a < total; a++) {
            double p = 1.0;
            for (int v = 0; v < n; v++) {
                if (!conn[v]) continue;
                p *= (pe[v] < 0) ? prior[st[v]] : P[(size_t)v * KK + st[pe[v]] * K + st[v]];
            }
            tot += p;
            if (p > best) {
                best = p;
                for (int v = 0; v < n; v++) best_state[v] = st[v];
            }
            if (post) for (int i = 0; i < nfree; i++) post[freev[i] * K + st[freev[i]]] += p;
            for (int i = nfree - 1; i >= 0; i--) { /* odometer, last free node fastest */
                if (++st[freev[i]] < K) break;
                st[freev[i]] = 0;
            }
        }
        if (best_p) *best_p = best;
        if (total_p) *total_p = tot;
        if (post) {
            for (int i = 0; i < nfree; i++)
                for (int x = 0; x < K; x++) post[freev[i] * K + x] /= tot;
        }
    }
done:
#undef PR
    free(P); free(pe); free(freev); free(st); free(conn);
    return rc;
}
