/*
 * bnk_indel.c -- CPU ORACLE, part 4: the gap-augmented substitution model and the extended peeling
 * recursion of asr.IndelPeeler (Rivas & Eddy 2008), restated in plain C in the reference's own log-space
 * form.  Test infrastructure only (see bnk_oracle.h); -ffp-contract=off.
 *
 * Follows (reference file:line):
 *   src/bn/ctmc/GapSubstModel.java:26-44      constructor: the base model's (normalised) R, then R_eps, Exp(R_eps),
 *                                             the alphabet with '-' appended, the stationary frequencies with a gap entry
 *   src/bn/ctmc/GapSubstModel.java:61-86      constructIndelR: R - mu I | mu column, bottom row lambda * pi, -lambda
 *   src/bn/ctmc/GapSubstModel.java:104-131    addGapToStationaryFreqs
 *   src/bn/ctmc/GapSubstModel.java:166-190    gammaT, ksiT (eq. 6)
 *   src/asr/IndelPeeler.java:229-252          decorate: log P(column | rate)
 *   src/asr/IndelPeeler.java:259-356          felsensteinsExtendedPeeling: leaves, ancestral gap / residue terms
 *   src/asr/IndelPeeler.java:359-390          containsGap
 *   src/asr/IndelPeeler.java:392-446          getProbOfInsertion (time of the PARENT node bpidx, 0 at the root),
 *                                             getProbGapAugmented, getProbOfGap (time of the child), ksiT, gammaT
 *   src/asr/IndelPeeler.java:125-161          calcProbAlnGivenTree: sum of columns + probExtraCol - log(1 - P(all-gap column))
 *   src/asr/IndelPeeler.java:163-193          probExtraCol (uses tree distances without a rate)
 *   src/asr/IndelPeeler.java:93-123           computeColumnPriors: one pass per rate
 *   src/asr/IndelPeeler.java:469-510, 526-560 optimiseMuLambda / LikelihoodEvaluator: Brent over mu = lambda
 *   src/smile/math/MathEx.java:4553-4584      logsumexp (max-shifted), logm1exp
 *   src/smile/math/special/Minimise.java:8-107 brent (golden start, parabolic steps, EPS = 10e-6, MAX = 500);
 *                                             the reference re-evaluates f(v), f(w), f(x) in every recursion -- the
 *                                             function is deterministic, so evaluations are cached here
 *   src/dat/phylo/PhyloBN.java:143-164        branch time = distance * rate; the root's SubstNode has time 0
 *   src/bn/ctmc/SubstNode.java:75-93, 149-163 probs = model.getProbs(t), getProb(child, parent) = probs[parent][child]
 */
#include "bnk_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

struct om_gap_model {
    int K1;            /* residues + gap */
    double mu, lambda;
    double F[OM_MAXK]; /* fGap */
    double R[OM_MAXK * OM_MAXK];
    double evec[OM_MAXK * OM_MAXK], ievec[OM_MAXK * OM_MAXK];
    double eval[OM_MAXK], eval_i[OM_MAXK];
};

om_gap_model *om_gap_model_create(const om_model *base, double mu, double lambda)
{
    int K = om_model_nstates(base), K1 = K + 1, i, j;
    double Rb[OM_MAXK * OM_MAXK], Fb[OM_MAXK];
    om_gap_model *g;
    if (K1 > OM_MAXK || mu + lambda < 0) return NULL;
    g = (om_gap_model *)calloc(1, sizeof(om_gap_model));
    if (!g) return NULL;
    om_model_get(base, Rb, NULL, NULL, NULL, Fb, NULL);
    g->K1 = K1; g->mu = mu; g->lambda = lambda;
    for (i = 0; i < K; i++) {
        g->R[i * K1 + K] = mu;
        for (j = 0; j < K; j++) {
            g->R[i * K1 + j] = Rb[i * K + j];
            if (i == j) g->R[i * K1 + j] -= mu;
        }
    }
    for (j = 0; j < K; j++) g->R[K * K1 + j] = Fb[j] * lambda;
    g->R[K * K1 + K] = -lambda;
    if (om_eigen(K1, g->R, g->evec, g->eval, g->eval_i, g->ievec) != 0) { free(g); return NULL; }
    if (mu + lambda > 0) {
        double gapFreq = mu / (mu + lambda);
        for (i = 0; i < K; i++) g->F[i] = (1 - gapFreq) * Fb[i];
        g->F[K] = gapFreq;
    } else {
        for (i = 0; i < K; i++) g->F[i] = Fb[i];
        g->F[K] = 0.0;
    }
    return g;
}

void om_gap_model_free(om_gap_model *g) { free(g); }
int om_gap_nstates(const om_gap_model *g) { return g->K1; }

void om_gap_get(const om_gap_model *g, double *R, double *F, double *evec, double *ievec, double *eval)
{
    int K1 = g->K1;
    if (R) memcpy(R, g->R, sizeof(double) * K1 * K1);
    if (F) memcpy(F, g->F, sizeof(double) * K1);
    if (evec) memcpy(evec, g->evec, sizeof(double) * K1 * K1);
    if (ievec) memcpy(ievec, g->ievec, sizeof(double) * K1 * K1);
    if (eval) memcpy(eval, g->eval, sizeof(double) * K1);
}

/* SubstModel.getProbs (src/bn/ctmc/SubstModel.java:303-327) on the augmented eigensystem */
void om_gap_getprobs(const om_gap_model *g, double time, double *P)
{
    int i, j, k, K = g->K1;
    double temp, iexp[OM_MAXK * OM_MAXK];
    for (k = 0; k < K; k++) {
        temp = om_exp(time * g->eval[k]);
        for (j = 0; j < K; j++) iexp[k * K + j] = g->ievec[k * K + j] * temp;
    }
    for (i = 0; i < K; i++)
        for (j = 0; j < K; j++) {
            temp = 0.0;
            for (k = 0; k < K; k++) temp += g->evec[i * K + k] * iexp[k * K + j];
            P[i * K + j] = fabs(temp);
        }
}

double om_gap_ksi(const om_gap_model *g, double time)
{
    if (g->mu == 0 && g->lambda == 0) return 0.0;
    {
        double insertionRate = g->lambda / (g->mu + g->lambda);
        double indelProp = 1 - om_exp(-((g->mu + g->lambda) * time));
        return insertionRate * indelProp;
    }
}

double om_gap_gamma(const om_gap_model *g, double time)
{
    if (g->mu == 0 && g->lambda == 0) return 0.0;
    {
        double indelTotal = g->mu + g->lambda;
        double deletionProp = g->mu / indelTotal;
        double indelProp = 1 - om_exp(-(indelTotal * time));
        return deletionProp * indelProp;
    }
}

double om_logsumexp(const double *a, int n)
{
    int i;
    double xmax = a[0], sum = 0.0;
    for (i = 1; i < n; i++)
        if (a[i] > xmax) xmax = a[i];
    if (xmax == -INFINITY) return -INFINITY;
    for (i = 0; i < n; i++) sum += om_exp(a[i] - xmax);
    return xmax + om_log(sum);
}

double om_logm1exp(double x)
{
    double LOG_HALF = om_log(0.5);
    if (x > 0) return NAN; /* IllegalArgumentException in the reference */
    if (x <= LOG_HALF) return log1p(-om_exp(x));
    return om_log(-expm1(x));
}

/* ---- per-node log tables for one (model, rate): shared by all columns, as PhyloBN's SubstNodes are ---- */
typedef struct {
    const om_gap_model *g;
    int n, K;                 /* K = residues */
    const int *parent;
    int *first, *kids;        /* children in index order */
    double *logA;             /* [node][parent residue][child residue] = log(P_eps(t_node)[p][x] * (1 - ksi(t_node))) */
    double *logdel;           /* [node] = log((1 - ksi(t_node)) * gamma(t_node)) */
    double *logins;           /* [node][q] = log(ksi(t_node) * F[q]) -- used for the node's CHILDREN (IndelPeeler.java:392-400) */
    double logF[OM_MAXK];
    int ncols;
    const uint8_t *obs;
    double geom;
    double *out;
} indel_ctx;

static int indel_prepare(indel_ctx *c, const om_gap_model *g, int n, const int *parent, const double *dist, double rate)
{
    int i, p, x, K1 = g->K1, K = K1 - 1;
    int *cnt;
    double P[OM_MAXK * OM_MAXK];
    memset(c, 0, sizeof(*c));
    c->g = g; c->n = n; c->K = K; c->parent = parent;
    c->first = (int *)calloc(n + 1, sizeof(int));
    c->kids = (int *)malloc(sizeof(int) * (n > 0 ? n : 1));
    cnt = (int *)calloc(n + 1, sizeof(int));
    c->logA = (double *)malloc(sizeof(double) * (size_t)n * K * K);
    c->logdel = (double *)malloc(sizeof(double) * n);
    c->logins = (double *)malloc(sizeof(double) * (size_t)n * K);
    if (!c->first || !c->kids || !cnt || !c->logA || !c->logdel || !c->logins) return -1;
    for (i = 0; i < n; i++)
        if (parent[i] >= 0) cnt[parent[i]]++;
    for (i = 0; i < n; i++) c->first[i + 1] = c->first[i] + cnt[i];
    memset(cnt, 0, sizeof(int) * (n + 1));
    for (i = 0; i < n; i++)
        if (parent[i] >= 0) c->kids[c->first[parent[i]] + cnt[parent[i]]++] = i;
    free(cnt);
    for (i = 0; i < n; i++) {
        double t = parent[i] < 0 ? 0.0 : dist[i] * rate;  /* SubstNode.time; 0 for the root node */
        double ksi = om_gap_ksi(g, t), noins = 1 - ksi;
        for (x = 0; x < K; x++) c->logins[(size_t)i * K + x] = om_log(ksi * g->F[x]);
        if (parent[i] < 0) {  /* no branch above the root: never read */
            c->logdel[i] = 0.0;
            continue;
        }
        om_gap_getprobs(g, t, P);
        for (p = 0; p < K; p++)
            for (x = 0; x < K; x++) c->logA[((size_t)i * K + p) * K + x] = om_log(P[p * K1 + x] * noins);
        c->logdel[i] = om_log(noins * om_gap_gamma(g, t));
    }
    for (x = 0; x < K; x++) c->logF[x] = om_log(g->F[x]);
    return 0;
}

static void indel_release(indel_ctx *c)
{
    free(c->first); free(c->kids); free(c->logA); free(c->logdel); free(c->logins);
}

/* IndelPeeler.decorate for one column; res / gap / cg are scratch of n*K, n, n */
static double indel_column(const indel_ctx *c, int col, double *res, double *gap, uint8_t *cg)
{
    int n = c->n, K = c->K, v, k, p, x;
    double terms[OM_MAXK + 1];
    for (v = 0; v < n * K; v++) res[v] = -INFINITY;
    for (v = 0; v < n; v++) gap[v] = -INFINITY;
    /* containsGap (:359-390) */
    for (v = n - 1; v >= 0; v--) {
        if (c->first[v + 1] == c->first[v]) cg[v] = (c->obs[(size_t)v * c->ncols + col] == OM_NONE);
        else {
            cg[v] = 1;
            for (k = c->first[v]; k < c->first[v + 1]; k++)
                if (!cg[c->kids[k]]) { cg[v] = 0; break; }
        }
    }
    for (v = n - 1; v >= 0; v--) {
        int nch = c->first[v + 1] - c->first[v];
        if (nch == 0) { /* calcLeafPeelingProbabilities (:273-284) */
            uint8_t o = c->obs[(size_t)v * c->ncols + col];
            if (o == OM_NONE) gap[v] = -INFINITY;
            else res[v * K + o] = 0.0;
            continue;
        }
        {   /* calcLogProbChildrenGivenAncestralGap (:297-327): logsumexp over the children */
            double *childterm = (double *)malloc(sizeof(double) * nch);
            for (k = 0; k < nch; k++) {
                int ch = c->kids[c->first[v] + k];
                if (cg[ch]) {
                    for (x = 0; x < K; x++) terms[x] = res[ch * K + x] + c->logins[(size_t)v * K + x];
                    terms[K] = gap[ch];
                    childterm[k] = om_logsumexp(terms, K + 1);
                } else childterm[k] = -INFINITY;
            }
            gap[v] = om_logsumexp(childterm, nch);
            /* calcLogProbChildrenGivenAncestralResidue (:329-356): sum over the children */
            for (p = 0; p < K; p++) {
                double s = 0.0;
                for (k = 0; k < nch; k++) {
                    int ch = c->kids[c->first[v] + k];
                    const double *la = c->logA + ((size_t)ch * K + p) * K;
                    int nt = K;
                    for (x = 0; x < K; x++) terms[x] = res[ch * K + x] + la[x];
                    if (cg[ch]) terms[nt++] = c->logdel[ch];
                    s += om_logsumexp(terms, nt);
                }
                res[v * K + p] = s;
            }
            free(childterm);
        }
    }
    {   /* decorate (:229-252) */
        double fin[2];
        for (x = 0; x < K; x++) terms[x] = c->logF[x] + res[x];
        fin[0] = gap[0];
        fin[1] = om_logsumexp(terms, K) + om_log(c->geom);
        return om_logsumexp(fin, 2);
    }
}

typedef struct { pthread_t th; const indel_ctx *c; int lo, hi; } indel_job;
static void *indel_worker(void *a_)
{
    indel_job *a = (indel_job *)a_;
    const indel_ctx *c = a->c;
    double *res = (double *)malloc(sizeof(double) * (size_t)c->n * c->K);
    double *gap = (double *)malloc(sizeof(double) * c->n);
    uint8_t *cg = (uint8_t *)malloc(c->n);
    for (int col = a->lo; col < a->hi; col++) c->out[col] = indel_column(c, col, res, gap, cg);
    free(res); free(gap); free(cg);
    return NULL;
}

/* log P(column | tree, model, rate) for every column: obs [node][col], OM_NONE = gap, only childless rows are read */
int om_indel_column_logl(const om_gap_model *g, int n, const int *parent, const double *dist, int ncols,
                         const uint8_t *obs, double rate, double geom, double *out_logl, int nthreads)
{
    indel_ctx c;
    indel_job jobs[256];
    int t;
    if (indel_prepare(&c, g, n, parent, dist, rate) != 0) return -1;
    c.ncols = ncols; c.obs = obs; c.geom = geom; c.out = out_logl;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    if (nthreads > ncols) nthreads = ncols > 0 ? ncols : 1;
    for (t = 0; t < nthreads; t++) {
        jobs[t].c = &c; jobs[t].lo = (int)((long)ncols * t / nthreads); jobs[t].hi = (int)((long)ncols * (t + 1) / nthreads);
        if (nthreads == 1) indel_worker(&jobs[t]);
        else pthread_create(&jobs[t].th, NULL, indel_worker, &jobs[t]);
    }
    if (nthreads > 1)
        for (t = 0; t < nthreads; t++) pthread_join(jobs[t].th, NULL);
    indel_release(&c);
    return 0;
}

/* probExtraCol (:163-193): distances WITHOUT a rate, model.ksiT */
double om_indel_prob_extra_col(const om_gap_model *g, int n, const int *parent, const double *dist, double geom)
{
    double *pStar = (double *)malloc(sizeof(double) * n), col;
    int *nch = (int *)calloc(n, sizeof(int));
    int v, u;
    for (v = 0; v < n; v++)
        if (parent[v] >= 0) nch[parent[v]]++;
    for (v = n - 1; v >= 0; v--) {
        if (nch[v] == 0) { pStar[v] = 0.0; continue; }
        {
            double childrenLL = 0.0;
            for (u = v + 1; u < n; u++) { /* children in index order */
                if (parent[u] != v) continue;
                childrenLL += pStar[u];
                childrenLL += om_log(1 - om_gap_ksi(g, dist[u]));
            }
            pStar[v] = childrenLL;
        }
    }
    col = 0.0;
    col += om_log(1 - geom);
    col += pStar[0];
    free(pStar); free(nch);
    return col;
}

/* calcProbAlnGivenTree (:125-161) */
int om_indel_aln_logl(const om_gap_model *g, int n, const int *parent, const double *dist, int ncols,
                      const uint8_t *obs, double geom, double *out, int nthreads)
{
    double *cols = (double *)malloc(sizeof(double) * (ncols > 0 ? ncols : 1));
    uint8_t *gapobs = (uint8_t *)malloc(n);
    double logLikelihood = 0.0, gapcol, extra;
    int c;
    if (om_indel_column_logl(g, n, parent, dist, ncols, obs, 1.0, geom, cols, nthreads) != 0) return -1;
    for (c = 0; c < ncols; c++) logLikelihood += cols[c];
    extra = om_indel_prob_extra_col(g, n, parent, dist, geom);
    memset(gapobs, OM_NONE, n);
    if (om_indel_column_logl(g, n, parent, dist, 1, gapobs, 1.0, geom, &gapcol, 1) != 0) return -1;
    *out = (extra - om_logm1exp(gapcol)) + logLikelihood;
    free(cols); free(gapobs);
    return 0;
}

/* ---- Minimise.brent with its evaluations cached ---- */
typedef struct {
    const om_model *base;
    int n, ncols, nthreads;
    const int *parent;
    const double *dist;
    const uint8_t *obs;
    double geom;
    int n_evals, n_cached;
    double xs[2048], fs[2048];
    int failed;
} brent_ctx;

static double brent_f(brent_ctx *b, double x)
{
    int i;
    double v = NAN;
    om_gap_model *g;
    for (i = 0; i < b->n_cached; i++)
        if (b->xs[i] == x) return b->fs[i];
    g = om_gap_model_create(b->base, x, x);
    if (!g || om_indel_aln_logl(g, b->n, b->parent, b->dist, b->ncols, b->obs, b->geom, &v, b->nthreads) != 0) b->failed = 1;
    if (g) om_gap_model_free(g);
    v = -v;
    b->n_evals++;
    if (b->n_cached < 2048) { b->xs[b->n_cached] = x; b->fs[b->n_cached] = v; b->n_cached++; }
    return v;
}

double om_indel_optimise_mulambda(const om_model *base, int n, const int *parent, const double *dist, int ncols,
                                  const uint8_t *obs, double geom, double min_val, double max_val, int *n_evals,
                                  int nthreads)
{
    const int MAX = 500;
    const double EPS = 10e-6;
    const double INVPHI = (1.0 + sqrt(5.0)) / 2.0 - 1.0;
    const double RATIO = (3.0 - sqrt(5.0)) / 2;
    brent_ctx b;
    double a = min_val, bb = max_val, v, w, x, dold = 0, eold = 0, m = NAN;
    int i = 0;
    memset(&b, 0, sizeof(b));
    b.base = base; b.n = n; b.parent = parent; b.dist = dist; b.ncols = ncols; b.obs = obs; b.geom = geom; b.nthreads = nthreads;
    x = bb + INVPHI * (a - bb);
    v = w = x;
    for (;;) {   /* the reference recurses; the arguments of the recursive call become the next iteration's state */
        double fv = brent_f(&b, v), fw = brent_f(&b, w), fx = brent_f(&b, x);
        double r, tq, tp, tq2, p, q, deltax, e, d, u, fu;
        int new_i = i + 1, safe, parabolic;
        m = 0.5 * (a + bb);
        if (b.failed) { m = NAN; break; }
        if (bb - a <= EPS) break;
        if (i > MAX) break;
        r = (x - w) * (fx - fv);
        tq = (x - v) * (fx - fw);
        tp = (x - v) * tq - (x - w) * r;
        tq2 = 2.0 * (tq - r);
        p = tq2 > 0.0 ? -tp : tp;
        q = tq2 > 0.0 ? tq2 : -tq2;
        safe = q != 0.0;
        deltax = safe ? p / q : 0.0;
        parabolic = safe && a < x + deltax && x + deltax < bb && fabs(deltax) < 0.5 * fabs(eold);
        if (parabolic) e = dold;
        else if (x < m) e = bb - x;
        else e = a - x;
        d = parabolic ? deltax : RATIO * e;
        u = x + d;
        fu = brent_f(&b, u);
        if (fu <= fx) {
            double newa = u < x ? a : x, newb = u < x ? x : bb;
            a = newa; bb = newb; v = w; w = x; x = u;
        } else {
            double newa = u < x ? u : a, newb = u < x ? bb : u;
            a = newa; bb = newb;
            if (fu <= fw || w == x) { v = w; w = u; }
            else if (fu <= fv || v == x || v == w) { v = u; }
        }
        dold = d; eold = e; i = new_i;
    }
    if (n_evals) *n_evals = b.n_evals;
    return m;
}
