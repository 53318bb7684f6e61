/*
 * bnk_ve.c -- CPU ORACLE, part 3: a generic restatement of bnkit's bucket elimination on
 * log-space factor tables, one column at a time.  It knows nothing about trees: it builds one
 * factor per Bayesian-network node, drops it into buckets and multiplies / maxes / sums exactly
 * the way the reference does.  Its only job is to pin the tree-specialised sweeps of bnk_sweeps.c
 * (and through them the GPU kernels): same joint states, same association of the floating-point
 * additions, posteriors and log-likelihoods equal to rounding.  Test infrastructure only.
 *
 * Follows (reference file:line):
 *   src/dat/phylo/PhyloBN.java:143-164        one variable per connected node, created in index order
 *   src/dat/Variable.java:228-235, src/bn/factor/AbstractFactor.java:155-170
 *                                             factor axes sorted by variable creation order, row-major
 *   src/bn/ctmc/SubstNode.java:471-609        makeDenseFactor: the four evidence cases and the prior
 *   src/bn/alg/VarElim.java:72-104, :121-142  makeQuery (relevance = ancestors of query and evidence,
 *                                             src/bn/BNet.java:571-596) / makeMPE (everything relevant)
 *   src/bn/alg/VarElim.java:187-334           infer: bucket per unobserved variable in topological order,
 *                                             a factor goes to the LAST bucket sharing a variable, buckets
 *                                             processed last to first, result forwarded to the nearest
 *                                             earlier bucket sharing a variable, scalars to bucket 0
 *   src/bn/factor/Factorize.java:262-308, :331-375  getComplexity / getProductTree (first cheapest pair,
 *                                             product appended at the end of the pool)
 *   src/bn/factor/Factorize.java:632-1000     getProduct = entry-wise '+' with index broadcasting
 *   src/bn/factor/Factorize.java:1414-1543    getMaxMargin: ascending scan, strict '>', (-inf, 0) start, traceback
 *   src/bn/factor/Factorize.java:1223-1401, :1199-1210  getMargin: sequential logSumOfLogs, ascending
 *   src/bn/factor/Factorize.java:1668-1717, src/bn/alg/CGTable.java:127-190, :844-863,
 *   src/bn/prob/EnumDistrib.java:363-371      getNormal, query, final re-normalisation
 *   src/bn/alg/VarElim.java:363-476           logLikelihood (see note at om_ve_loglik)
 *
 * Orders the reference leaves to identity hash codes (HashMap / HashSet iteration) are fixed here to
 * one admissible choice: topological order = depth-first with children in DESCENDING index order
 * (so that, processing buckets last to first, sibling messages reach their parent's bucket in
 * ascending index order), and node factors are created in index order.
 */
#include "bnk_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int var, state, a, b; } trace_t;  /* leaf: (var, state); merge: var = -1, children a, b */

typedef struct {
    int nv;
    int vars[3];     /* ascending variable ids */
    double *val;     /* K^nv entries, row-major over vars */
    int *tr;         /* trace id per entry or NULL */
} factor;

typedef struct {
    int K;
    trace_t *trace; int ntrace, captrace;
    int traced;
} ve_ctx;

static int new_trace(ve_ctx *c, int var, int state, int a, int b)
{
    if (c->ntrace == c->captrace) {
        c->captrace = c->captrace ? 2 * c->captrace : 1024;
        c->trace = (trace_t *)realloc(c->trace, sizeof(trace_t) * c->captrace);
    }
    c->trace[c->ntrace] = (trace_t){var, state, a, b};
    return c->ntrace++;
}
static int merge_trace(ve_ctx *c, int a, int b)
{
    if (a < 0) return b;
    if (b < 0) return a;
    return new_trace(c, -1, 0, a, b);
}
static int ipow(int K, int n) { int r = 1; while (n-- > 0) r *= K; return r; }

static factor *new_factor(ve_ctx *c, int nv, const int *vars)
{
    factor *f = (factor *)calloc(1, sizeof(factor));
    f->nv = nv;
    for (int i = 0; i < nv; i++) f->vars[i] = vars[i];
    const int sz = ipow(c->K, nv);
    f->val = (double *)malloc(sizeof(double) * sz);
    f->tr = NULL;
    if (c->traced) { f->tr = (int *)malloc(sizeof(int) * sz); for (int i = 0; i < sz; i++) f->tr[i] = -1; }
    return f;
}
static void free_factor(factor *f) { if (f) { free(f->val); free(f->tr); free(f); } }
static int has_var(const factor *f, int v) { for (int i = 0; i < f->nv; i++) if (f->vars[i] == v) return 1; return 0; }

/* Factorize.getProduct(X, Y) */
static factor *product(ve_ctx *c, const factor *X, const factor *Y)
{
    const int K = c->K;
    int vars[6], nv = 0, i = 0, j = 0;
    while (i < X->nv || j < Y->nv) {
        if (j >= Y->nv || (i < X->nv && X->vars[i] < Y->vars[j])) vars[nv++] = X->vars[i++];
        else if (i >= X->nv || Y->vars[j] < X->vars[i]) vars[nv++] = Y->vars[j++];
        else { vars[nv++] = X->vars[i]; i++; j++; }
    }
    if (nv > 3) return NULL;
    factor *Z = new_factor(c, nv, vars);
    const int sz = ipow(K, nv);
    for (int e = 0; e < sz; e++) {
        int st[3] = {0, 0, 0}, r = e;
        for (int k = nv - 1; k >= 0; k--) { st[k] = r % K; r /= K; }
        int ex = 0, ey = 0;
        for (int k = 0; k < nv; k++) {
            if (has_var(X, vars[k])) ex = ex * K + st[k];
            if (has_var(Y, vars[k])) ey = ey * K + st[k];
        }
        Z->val[e] = X->val[ex] + Y->val[ey];
        if (c->traced) Z->tr[e] = merge_trace(c, X->tr ? X->tr[ex] : -1, Y->tr ? Y->tr[ey] : -1);
    }
    return Z;
}

/* Factorize.getComplexity(Xvars, Yvars, includeJoin = true): product of the sizes of the union */
static int complexity(const ve_ctx *c, const factor *X, const factor *Y)
{
    int n = X->nv;
    for (int j = 0; j < Y->nv; j++) if (!has_var(X, Y->vars[j])) n++;
    return ipow(c->K, n);
}

/* Factorize.getProduct(AbstractFactor[]) via getProductTree */
static factor *product_all(ve_ctx *c, factor **fs, int n)
{
    if (n == 1) return fs[0];
    factor **pool = (factor **)malloc(sizeof(factor *) * n);
    char *owned = (char *)calloc(n, 1);
    for (int i = 0; i < n; i++) pool[i] = fs[i];
    int m = n;
    while (m > 1) {
        int lowest = 0x7fffffff, a = -1, b = -1;
        if (m == 2 && n == 2) { a = 0; b = 1; }
        else
            for (int i = 0; i < m; i++)
                for (int j = i + 1; j < m; j++) {
                    const int cost = complexity(c, pool[i], pool[j]);
                    if (cost < lowest) { a = i; b = j; lowest = cost; }
                }
        factor *z = product(c, pool[a], pool[b]);
        if (owned[a]) free_factor(pool[a]);
        if (owned[b]) free_factor(pool[b]);
        /* remove b then a, append z */
        for (int k = b; k < m - 1; k++) { pool[k] = pool[k + 1]; owned[k] = owned[k + 1]; }
        for (int k = a; k < m - 2; k++) { pool[k] = pool[k + 1]; owned[k] = owned[k + 1]; }
        pool[m - 2] = z; owned[m - 2] = 1;
        m--;
    }
    factor *r = pool[0];
    free(pool); free(owned);
    return r;
}

/* Factorize.getMaxMargin / getMargin over one variable */
static factor *eliminate(ve_ctx *c, const factor *F, int var, int maxout)
{
    const int K = c->K;
    int vars[3], nv = 0, pos = -1;
    for (int i = 0; i < F->nv; i++) { if (F->vars[i] == var) pos = i; else vars[nv++] = F->vars[i]; }
    factor *Y = new_factor(c, nv, vars);
    const int sz = ipow(K, nv);
    int stride = 1;
    for (int k = F->nv - 1; k > pos; k--) stride *= K;
    for (int e = 0; e < sz; e++) {
        /* index of the (e, var = 0) entry in F */
        int st[3] = {0, 0, 0}, r = e;
        for (int k = nv - 1; k >= 0; k--) { st[k] = r % K; r /= K; }
        int base = 0, q = 0;
        for (int k = 0; k < F->nv; k++) base = base * K + (k == pos ? 0 : st[q++]);
        if (maxout) {
            double max = -INFINITY; int maxidx = 0;
            for (int x = 0; x < K; x++) {
                const double xval = F->val[base + x * stride];
                if (xval > max) { max = xval; maxidx = x; }
            }
            Y->val[e] = max;
            if (c->traced) Y->tr[e] = merge_trace(c, new_trace(c, var, maxidx, -1, -1), F->tr ? F->tr[base + maxidx * stride] : -1);
        } else {
            double logsum = 0.0;
            for (int x = 0; x < K; x++) {
                const double xval = F->val[base + x * stride];
                logsum = (x == 0) ? xval : om_logsum(logsum, xval);
            }
            Y->val[e] = logsum;
        }
    }
    return Y;
}

/* ------------------------------------------------------------------ */
typedef struct { int var; factor **fs; int nf, cap; } bucket;
static void bucket_put(bucket *b, factor *f)
{
    if (b->nf == b->cap) { b->cap = b->cap ? 2 * b->cap : 4; b->fs = (factor **)realloc(b->fs, sizeof(factor *) * b->cap); }
    b->fs[b->nf++] = f;
}

typedef struct {
    const om_model *m;
    int n;
    const int *parent;
    const double *dist;
    const uint8_t *obs;
    double rate;
    int *pe;          /* effective parent or -1 */
    uint8_t *conn;    /* has a BN node */
} net;

static void build_net(net *N, const uint8_t *present)
{
    N->pe = (int *)malloc(sizeof(int) * N->n);
    N->conn = (uint8_t *)calloc(N->n, 1);
    for (int v = 0; v < N->n; v++) {
        const int pr = present == NULL || present[v];
        N->pe[v] = (pr && N->parent[v] >= 0 && (present == NULL || present[N->parent[v]])) ? N->parent[v] : -1;
        if (N->pe[v] >= 0) { N->conn[v] = 1; N->conn[N->pe[v]] = 1; }
    }
    if (present) for (int v = 0; v < N->n; v++) if (!present[v]) N->conn[v] = 0;
}

/* SubstNode.makeDenseFactor for node v; relevant[] marks the variables in scope */
static factor *node_factor(ve_ctx *c, const net *N, int v, const uint8_t *relevant)
{
    const int K = c->K;
    const uint8_t ov = N->obs[v];
    double P[OM_MAXK * OM_MAXK], prior[OM_MAXK];
    if (N->pe[v] >= 0) {
        const int p = N->pe[v];
        const uint8_t op = N->obs[p];
        om_getprobs(N->m, N->dist[v] * N->rate, P);
        (void)relevant;  /* on a tree a relevant node's parent is always relevant (ancestors are relevant) */
        if (ov == OM_NONE && op == OM_NONE) {
            int vars[2] = {p, v};  /* parent created first -> first axis */
            factor *f = new_factor(c, 2, vars);
            for (int y = 0; y < K; y++) for (int x = 0; x < K; x++) f->val[y * K + x] = om_log(P[y * K + x]);
            return f;
        } else if (ov != OM_NONE && op == OM_NONE) {
            int vars[1] = {p};
            factor *f = new_factor(c, 1, vars);
            for (int y = 0; y < K; y++) f->val[y] = om_log(P[y * K + ov]);
            return f;
        } else if (ov == OM_NONE) {
            int vars[1] = {v};
            factor *f = new_factor(c, 1, vars);
            for (int x = 0; x < K; x++) f->val[x] = om_log(P[op * K + x]);
            return f;
        }
        factor *f = new_factor(c, 0, NULL);
        f->val[0] = om_log(P[op * K + ov]);
        return f;
    }
    om_model_get(N->m, NULL, NULL, NULL, NULL, NULL, prior);
    if (ov != OM_NONE) { factor *f = new_factor(c, 0, NULL); f->val[0] = om_log(prior[ov]); return f; }
    int vars[1] = {v};
    factor *f = new_factor(c, 1, vars);
    for (int x = 0; x < K; x++) f->val[x] = om_log(prior[x]);
    return f;
}

/* topological order of the variables in scope: depth-first, children in descending index order */
static int topo_order(const net *N, const uint8_t *inscope, int *order)
{
    const int n = N->n;
    int *first = (int *)calloc(n + 1, sizeof(int)), *kids = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    int *fill = (int *)calloc(n + 1, sizeof(int)), *stack = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int v = 0; v < n; v++) if (N->pe[v] >= 0) first[N->pe[v] + 1]++;
    for (int v = 0; v < n; v++) first[v + 1] += first[v];
    for (int v = 0; v < n; v++) if (N->pe[v] >= 0) kids[first[N->pe[v]] + fill[N->pe[v]]++] = v;
    int cnt = 0, sp = 0;
    for (int r = n - 1; r >= 0; r--) if (N->conn[r] && N->pe[r] < 0) stack[sp++] = r;  /* lowest root on top */
    while (sp > 0) {
        const int v = stack[--sp];
        if (inscope[v]) order[cnt++] = v;
        for (int e = first[v]; e < first[v + 1]; e++) stack[sp++] = kids[e];  /* ascending push -> descending visit */
    }
    free(first); free(kids); free(fill); free(stack);
    return cnt;
}

/* VarElim.infer.  query < 0: MPE (max-out) or, with maxout = 0, sum everything (log-likelihood).
 * Returns the final factor of bucket 0 (scalar, or F(query)). */
static factor *infer(ve_ctx *c, const net *N, int query, int maxout, const uint8_t *relevant)
{
    const int n = N->n;
    int *order = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    uint8_t *isx = (uint8_t *)calloc(n, 1);
    for (int v = 0; v < n; v++) isx[v] = N->conn[v] && relevant[v] && N->obs[v] == OM_NONE && v != query;
    const int nx = topo_order(N, isx, order);
    const int nb = nx + 1;
    bucket *bk = (bucket *)calloc(nb, sizeof(bucket));
    int *bucket_of = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    for (int v = 0; v < n; v++) bucket_of[v] = -1;
    bk[0].var = query;
    if (query >= 0) bucket_of[query] = 0;
    for (int i = 0; i < nx; i++) { bk[i + 1].var = order[i]; bucket_of[order[i]] = i + 1; }
    /* fill: node factors in index order; a factor goes to the last bucket sharing one of its variables */
    for (int v = 0; v < n; v++) {
        if (!N->conn[v] || !relevant[v]) continue;
        factor *f = node_factor(c, N, v, relevant);
        int best = 0;
        for (int k = 0; k < f->nv; k++) if (bucket_of[f->vars[k]] > best) best = bucket_of[f->vars[k]];
        bucket_put(&bk[best], f);
    }
    factor *final = NULL;
    for (int i = nb - 1; i >= 0; i--) {
        bucket *b = &bk[i];
        if (b->nf == 0) continue;
        factor *result = product_all(c, b->fs, b->nf);
        for (int k = 0; k < b->nf; k++) if (b->fs[k] != result) free_factor(b->fs[k]);
        if (i == 0) { final = result; break; }
        factor *r2 = eliminate(c, result, b->var, maxout);
        free_factor(result);
        int dst = 0;  /* nearest earlier bucket sharing a variable; factors without variables go to bucket 0 */
        for (int jj = i - 1; jj >= 0; jj--) {
            int match = 0;
            for (int k = 0; k < r2->nv; k++) if (bk[jj].var == r2->vars[k]) match = 1;
            if (match) { dst = jj; break; }
        }
        bucket_put(&bk[dst], r2);
    }
    for (int i = 0; i < nb; i++) free(bk[i].fs);
    free(bk); free(order); free(isx); free(bucket_of);
    return final;
}

static void collect(const ve_ctx *c, int t, uint8_t *out)
{
    if (t < 0) return;
    int *stack = (int *)malloc(sizeof(int) * (c->ntrace + 1));
    int sp = 0;
    stack[sp++] = t;
    while (sp > 0) {
        const trace_t x = c->trace[stack[--sp]];
        if (x.var >= 0) out[x.var] = (uint8_t)x.state;
        if (x.a >= 0) stack[sp++] = x.a;
        if (x.b >= 0) stack[sp++] = x.b;
    }
    free(stack);
}

int om_ve_joint(const om_model *m, int n, const int *parent, const double *dist,
                const uint8_t *present, const uint8_t *obs, double rate,
                uint8_t *out_state, double *out_logmax)
{
    ve_ctx c = {om_model_nstates(m), NULL, 0, 0, 1};
    net N = {m, n, parent, dist, obs, rate, NULL, NULL};
    build_net(&N, present);
    uint8_t *all = (uint8_t *)malloc(n > 0 ? n : 1);
    memset(all, 1, n);
    for (int v = 0; v < n; v++) out_state[v] = (present == NULL || present[v]) ? obs[v] : OM_NONE;
    factor *f = infer(&c, &N, -1, 1, all);
    if (f) {
        collect(&c, f->tr ? f->tr[0] : -1, out_state);
        if (out_logmax) *out_logmax = f->val[0];
        free_factor(f);
    }
    free(all); free(N.pe); free(N.conn); free(c.trace);
    return 0;
}

int om_ve_marginal(const om_model *m, int n, const int *parent, const double *dist,
                   const uint8_t *present, const uint8_t *obs, double rate,
                   int query, double *out_post)
{
    ve_ctx c = {om_model_nstates(m), NULL, 0, 0, 0};
    const int K = c.K;
    net N = {m, n, parent, dist, obs, rate, NULL, NULL};
    build_net(&N, present);
    if (query < 0 || query >= n || !N.conn[query] || obs[query] != OM_NONE) { free(N.pe); free(N.conn); return 1; }
    /* BNet.getRelevantAndSome: ancestors of the query and of every instantiated node */
    uint8_t *rel = (uint8_t *)calloc(n, 1);
    for (int v = 0; v < n; v++)
        if (N.conn[v] && (v == query || obs[v] != OM_NONE))
            for (int u = v; u >= 0 && !rel[u]; u = N.pe[u]) rel[u] = 1;
    factor *f = infer(&c, &N, query, 0, rel);
    /* Factorize.getNormal, CGTable.query, EnumDistrib.normalise */
    double y[OM_MAXK], logmax = -INFINITY, sum = 0.0, tot = 0.0;
    for (int x = 0; x < K; x++) if (f->val[x] > logmax) logmax = f->val[x];
    for (int x = 0; x < K; x++) { y[x] = f->val[x] - logmax; sum = (x == 0) ? y[x] : om_logsum(sum, y[x]); }
    for (int x = 0; x < K; x++) { out_post[x] = om_exp(y[x] - sum); tot += out_post[x]; }
    for (int x = 0; x < K; x++) out_post[x] /= tot;
    free_factor(f); free(rel); free(N.pe); free(N.conn);
    return 0;
}

/* VarElim.logLikelihood sums out every uninstantiated variable and takes getLogSum of what reaches
 * bucket 0.  NOTE: in the reference a bucket result without variables (a whole component of a forest
 * eliminated before bucket 0) finds no bucket to go to and is dropped (VarElim.java:455 "FIXME: problem
 * if FT is atomic"); here such scalars are kept (bucket 0), i.e. the value is log P(evidence) over all
 * components -- identical to the reference on a single tree. */
int om_ve_loglik(const om_model *m, int n, const int *parent, const double *dist,
                 const uint8_t *present, const uint8_t *obs, double rate, double *out_logl)
{
    ve_ctx c = {om_model_nstates(m), NULL, 0, 0, 0};
    net N = {m, n, parent, dist, obs, rate, NULL, NULL};
    build_net(&N, present);
    uint8_t *all = (uint8_t *)malloc(n > 0 ? n : 1);
    memset(all, 1, n);
    factor *f = infer(&c, &N, -1, 0, all);
    *out_logl = f ? f->val[0] : 0.0;
    free_factor(f);
    free(all); free(N.pe); free(N.conn);
    return 0;
}
