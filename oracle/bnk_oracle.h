/*
 * bnk_oracle.h -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic that bodenlab/bnkit runs, one
 * alignment column at a time, for GRASP character inference:
 *   bn.ctmc.SubstModel / bn.math.Matrix.Exp  -> om_model_*, om_getprobs
 *   bn.ctmc.SubstNode.makeDenseFactor        -> factor construction in om_ve_*
 *   bn.factor.Factorize (product / margin / max-margin / normal)
 *   bn.alg.VarElim (makeMPE / makeQuery / infer / logLikelihood)
 *   bn.alg.CGTable (getMPE / query), bn.prob.EnumDistrib
 *   dat.phylo.PhyloBN.create, dat.phylo.IdxTree pruning
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product (bnkit_b200/,
 * include/bnkgpu.h) never links, imports or calls it.
 *
 * PARITY PINNING: the reference is Java and cannot be compiled or run in the
 * build container (no JDK), so the oracle is pinned against the reference's own
 * printed known answers (tutorial_ML.html), its brute-force test method
 * (MaxLhoodJointTest / MaxLhoodMarginalTest), the IdxTreeTest prune fixture and
 * the documented Recon output in docs/json-api.md -- see tests/test_oracle.py.
 */
#ifndef BNK_ORACLE_H
#define BNK_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OM_NONE 255 /* "no observation" / "no state" marker in uint8 arrays */
#define OM_MAXK 24  /* largest alphabet the oracle handles */

typedef struct om_model om_model;

/* ---- elementary functions (fdlibm algorithm == java.lang.StrictMath) ---- */
double om_exp(double x);
double om_log(double x);
/* Factorize.logSumOfLogs (Factorize.java:1199-1210) */
double om_logsum(double logx, double logy);

/* ---- substitution models (SubstModel.java:80-110, :192-219, :303-327) ---- */
/* name in {"LG","JTT","WAG","Dayhoff","JC","Yang"}; F_override (K values) or NULL. */
om_model *om_model_create(const char *name, const double *F_override);
/* generic: K states, frequencies F, matrix M (K*K row-major).  symmetric!=0: M is the
 * PAML-style exchangeability (upper triangle used); normalise!=0: scale to 1 subst/time. */
om_model *om_model_create_custom(int K, const double *F, const double *M, int symmetric, int normalise);
void om_model_free(om_model *m);
int om_model_nstates(const om_model *m);
/* any output pointer may be NULL. R, evec, ievec: K*K row-major; eval, F, prior: K. */
void om_model_get(const om_model *m, double *R, double *evec, double *ievec, double *eval,
                  double *F, double *prior);
/* P[i*K+j] = P(child=j | parent=i, t)  (SubstModel.getProbs) */
void om_getprobs(const om_model *m, double t, double *P);
/* the raw eigen-solver, exposed for tests: A (n*n) -> evec, eval_r, eval_i, ievec */
int om_eigen(int n, const double *A, double *evec, double *wr, double *wi, double *ievec);

/* ---- forest construction (IdxTree.java:411-452, :471-503, :746-751) ---- */
/* present_in[n] (1 = position exists at node) -> present_out[n] after optional orphan
 * removal; returns number of surviving nodes; *n_roots = surviving nodes w/o surviving parent */
int om_prune(int n, const int *parent, const uint8_t *present_in, int remove_orphans,
             uint8_t *present_out, int *n_roots);

/* ---- generic bucket elimination on one column (the faithful restatement) ---- */
/* parent[i] < i, parent[root] = -1.  present may be NULL (all present).
 * obs[i] = state index or OM_NONE.  Return 0 on success. */
int om_ve_joint(const om_model *m, int n, const int *parent, const double *dist,
                const uint8_t *present, const uint8_t *obs, double rate,
                uint8_t *out_state, double *out_logmax);
/* posterior of one query node; returns 1 if the node has no BN variable */
int om_ve_marginal(const om_model *m, int n, const int *parent, const double *dist,
                   const uint8_t *present, const uint8_t *obs, double rate,
                   int query, double *out_post);
/* log P(evidence), all components summed (see DESIGN.md on VarElim.logLikelihood's FIXME) */
int om_ve_loglik(const om_model *m, int n, const int *parent, const double *dist,
                 const uint8_t *present, const uint8_t *obs, double rate, double *out_logl);

/* ---- tree-specialised sweeps with the same association order (batch, threaded) ---- */
/* arrays are [node][col] with col fastest; rates[ncols] or NULL (= 1.0) */
int om_fast_joint(const om_model *m, int n, const int *parent, const double *dist, int ncols,
                  const uint8_t *present, const uint8_t *obs, const double *rates,
                  uint8_t *out_state, double *out_logmax, int nthreads);
/* out_post [node][col][K] (NaN where the node has no posterior), out_logl[ncols]; either may be NULL */
int om_fast_marginal(const om_model *m, int n, const int *parent, const double *dist, int ncols,
                     const uint8_t *present, const uint8_t *obs, const double *rates,
                     double *out_post, double *out_logl, int nthreads);

/* ---- brute-force enumeration (the reference's own test oracle) ---- */
/* enumerates all K^(#unobserved connected nodes) assignments (<= max_assign); linear-space
 * products of P entries as in MaxLhoodJointTest.java:318-350.  best_state = first strict max
 * in lexicographic order of the unobserved nodes (ascending node index, last varies fastest).
 * post (n*K, may be NULL) = exact posteriors by summation; total_p = P(evidence). */
int om_brute(const om_model *m, int n, const int *parent, const double *dist,
             const uint8_t *present, const uint8_t *obs, double rate, long max_assign,
             uint8_t *best_state, double *best_p, double *post, double *total_p);

/* ---- gap-augmented model + extended peeling (bnk_indel.c: GapSubstModel, asr.IndelPeeler) ---- */
typedef struct om_gap_model om_gap_model;
/* new GapSubstModel(F, IRM, alphabet, mu, lambda, ...) on top of an already constructed base model */
om_gap_model *om_gap_model_create(const om_model *base, double mu, double lambda);
void om_gap_model_free(om_gap_model *g);
int om_gap_nstates(const om_gap_model *g); /* residues + 1 */
void om_gap_get(const om_gap_model *g, double *R, double *F, double *evec, double *ievec, double *eval);
void om_gap_getprobs(const om_gap_model *g, double t, double *P); /* (K+1)^2, P[parent][child] */
double om_gap_ksi(const om_gap_model *g, double t);
double om_gap_gamma(const om_gap_model *g, double t);
double om_logsumexp(const double *a, int n);
double om_logm1exp(double x);
/* IndelPeeler.decorate for every column; obs [node][col], OM_NONE = gap (childless rows only) */
int om_indel_column_logl(const om_gap_model *g, int n, const int *parent, const double *dist, int ncols,
                         const uint8_t *obs, double rate, double geom, double *out_logl, int nthreads);
double om_indel_prob_extra_col(const om_gap_model *g, int n, const int *parent, const double *dist, double geom);
/* IndelPeeler.calcProbAlnGivenTree */
int om_indel_aln_logl(const om_gap_model *g, int n, const int *parent, const double *dist, int ncols,
                      const uint8_t *obs, double geom, double *out, int nthreads);
/* IndelPeeler.optimiseMuLambda (Minimise.brent over mu = lambda); n_evals = distinct likelihood evaluations */
double om_indel_optimise_mulambda(const om_model *base, int n, const int *parent, const double *dist, int ncols,
                                  const uint8_t *obs, double geom, double min_val, double max_val, int *n_evals,
                                  int nthreads);

#ifdef __cplusplus
}
#endif
#endif
