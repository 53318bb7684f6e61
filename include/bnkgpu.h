/*
 * bnkgpu.h -- C ABI of libbnkgpu.so: B200 (sm_100a) per-alignment-column inference on the
 * phylogenetic Bayesian network behind GRASP ancestral sequence reconstruction.
 *
 * This is the drop-in boundary for ONE path of bodenlab/bnkit: what
 *   asr.Prediction.getJoint    (src/asr/Prediction.java:614-649)
 *   asr.Prediction.getMarginal (src/asr/Prediction.java:555-597)
 * run today, one column at a time, through asr.MaxLhoodJoint / asr.MaxLhoodMarginal
 * (TreeDecor<E>, src/dat/phylo/TreeDecor.java:8-11), dat.phylo.PhyloBN.create
 * (src/dat/phylo/PhyloBN.java:143-164), bn.ctmc.SubstNode / SubstModel.getProbs and
 * bn.alg.VarElim (src/bn/alg/VarElim.java:187-334, :363-476).  The reference has no FFI of
 * its own (pure Java); INTEGRATION.md shows the JNI stub a maintainer would add.
 *
 * Conventions
 *   - every function returns BNK_OK (0) or a negative bnk_status; bnk_last_error() gives
 *     the message for the calling thread.  Nothing throws across the boundary.
 *   - state indices follow the model alphabet order (ARNDCQEGHILKMFPSTWYV for the protein
 *     models, ACGT for JC/Yang): index i <-> model.getDomain().get(i).  BNK_NONE (255)
 *     means "not instantiated" on input and "no state" on output (Java null).
 *   - trees are given as dat.phylo.IdxTree stores them (src/dat/phylo/IdxTree.java:27-32):
 *     parent[i] < i, parent[root] = -1, distance[i] = branch length above node i.
 *   - per-column arrays are [node][column] with the column index fastest, i.e. one
 *     alignment row per node (dat.EnumSeq order), n_cols bytes per row.
 *   - a workspace owns its device buffers and a CUDA stream; calls on one workspace are
 *     serialised by the caller, different workspaces are independent.
 *   - lifetimes: a workspace copies the model and holds a reference on the tree, so both handles may be
 *     destroyed before the workspace (the tree's memory goes with its last workspace).
 *   - every entry point selects the workspace's device for the duration of the call and restores the
 *     caller's current device before it returns.
 *   - there is no CPU fallback: every entry point that computes fails with
 *     BNK_ERR_CUDA when no sm_100-class device is usable.
 */
#ifndef BNKGPU_H
#define BNKGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BNK_NONE 255
#define BNK_MAX_STATES 20

typedef enum {
    BNK_OK = 0,
    BNK_ERR_ARG = -1,      /* invalid argument (IllegalArgumentException in the reference) */
    BNK_ERR_MODEL = -2,    /* unknown model name / eigen-decomposition failed (ArithmeticException) */
    BNK_ERR_TREE = -3,     /* malformed tree arrays */
    BNK_ERR_CUDA = -4,     /* no usable device, launch or runtime failure */
    BNK_ERR_NOMEM = -5,    /* host or device allocation failed */
    BNK_ERR_STATE = -6,    /* call order: nothing uploaded / result not computed */
    BNK_ERR_QUERY = -7     /* invalid ancestor id (ASRRuntimeException, Prediction.java:561-562) */
} bnk_status;

typedef struct bnk_model bnk_model; /* <-> bn.ctmc.SubstModel */
typedef struct bnk_tree bnk_tree;   /* <-> dat.phylo.IdxTree  */
typedef struct bnk_ws bnk_ws;       /* device workspace for one (model, tree) pair */

const char *bnk_last_error(void);
const char *bnk_version(void);
/* number of visible CUDA devices of compute capability 10.x; < 0 on runtime failure */
int bnk_device_count(void);

/* ---- indel inference: the per-column presence mask (what the sweeps below take as `present`) -------- */
/* asr.Prediction.PredictByBidirEdgeParsimony + patchAncestorsWithBEP (src/asr/Prediction.java:1443-1590) with
 * GRASP's defaults RECODE_NULL = true, NIBBLE = true (src/asr/GRASP.java:36-40), as a mask transform:
 *   obs         [n_nodes][n_cols]  alignment rows of the childless nodes (BNK_NONE = gap); other rows are ignored
 *   present_out [n_nodes][n_cols]  extants: 1 where the sequence has a residue; ancestors: 1 where the column is a
 *                                  node of the ancestor's inferred partial-order graph (Prediction.getTree :336-349)
 * The parsimony sweeps of all 2 (n_cols + 1) edge instances (asr.Parsimony, src/asr/Parsimony.java:215-412) run on
 * `device` (< 0: current); graph assembly / nibbling / patching is host-side integer work.  Host buffers.
 * info_or_null[8]: ancestors without a start-to-end path after the first pass, patch rounds, largest number of
 * distinct edge targets in one column, kernels launched, microseconds spent in the device sweeps (with their
 * copies) / in host graph assembly / in total / in the kernels alone (CUDA events).  The tree needs a single root at index 0. */
int bnk_bep_presence(const bnk_tree *t, int32_t n_cols, const uint8_t *obs, int device, uint8_t *present_out,
                     int32_t *info_or_null);

/* The same inference, also handing out the ancestors' partial-order graphs (<-> the Map<Object, POGraph> a
 * Prediction holds, src/asr/Prediction.java:60-75): what Prediction.getConsensus / getSequence / toJSON need
 * downstream of the sweeps.  Edges of ancestor `node` (tree index) sorted by (from, to); from = -1 is the start
 * terminal, to = n_cols the end terminal; flags bit 0 / bit 1 = inferred by a forward / backward edge instance
 * (POGraph.BidirEdge, src/dat/pog/POGraph.java:1039-1053; both = "Recip").  bnk_pogs_edges returns the edge count;
 * any output array may be NULL. */
typedef struct bnk_pogs bnk_pogs;
int bnk_bep_infer(const bnk_tree *t, int32_t n_cols, const uint8_t *obs, int device, uint8_t *present_out,
                  int32_t *info_or_null, bnk_pogs **out);
int bnk_pogs_n_edges(const bnk_pogs *p, int32_t node);
int bnk_pogs_edges(const bnk_pogs *p, int32_t node, int32_t *from, int32_t *to, uint8_t *flags);
void bnk_pogs_destroy(bnk_pogs *p);

/* ---- substitution models ------------------------------------------------------------ */
/* SubstModel.createModel(name) / createModel(name, empiricalFreqs)  (SubstModel.java:334-385).
 * name in {"LG","JTT","WAG","Dayhoff","JC","Yang"}; F_or_null = 20 (or 4) frequencies. */
int bnk_model_create(const char *name, const double *F_or_null, bnk_model **out);
/* new SubstModel(F, IRM, alphabet, symmetric, normalise)  (SubstModel.java:80-110) */
int bnk_model_create_custom(int n_states, const double *F, const double *IRM, int symmetric,
                            int normalise, bnk_model **out);
/* Adopt an eigensystem computed elsewhere -- what the JNI stub passes from the JVM's own
 * SubstModel.getRexp() (getEigvec / getInvEigvec / getEigval, Matrix.java:63-77) so the
 * device tables start from exactly the reference's bits. */
int bnk_model_create_from_eigen(int n_states, const double *F, const double *evec,
                                const double *ievec, const double *eval, bnk_model **out);
void bnk_model_destroy(bnk_model *m);
int bnk_model_nstates(const bnk_model *m);
/* any output may be NULL. R, evec, ievec: K*K row-major; eval, F, prior: K. */
int bnk_model_get(const bnk_model *m, double *R, double *evec, double *ievec, double *eval,
                  double *F, double *prior);

/* ---- trees -------------------------------------------------------------------------- */
/* IdxTree arrays (IdxTree.toJSON "Parents"/"Distances", IdxTree.java:277-351).
 * dist[root] is ignored.  Distances are used as given (Newick.java:54-56 already mapped 0 -> 1e-5). */
int bnk_tree_create(int32_t n_nodes, const int32_t *parent, const double *dist, bnk_tree **out);
void bnk_tree_destroy(bnk_tree *t);
int bnk_tree_size(const bnk_tree *t);
int bnk_tree_n_internal(const bnk_tree *t); /* nodes with at least one child = ancestors */
/* IdxTree.getPrunedIndex(pruneMe, removeOrphans=true) (IdxTree.java:411-464) as a mask
 * transform: present_in/out [n_nodes][n_cols]; out may alias in.  Host-side integer work. */
int bnk_tree_prune_orphans(const bnk_tree *t, int32_t n_cols, const uint8_t *present_in,
                           uint8_t *present_out);

/* ---- workspace ---------------------------------------------------------------------- */
/* device < 0: current device.  stream: a cudaStream_t to launch on (e.g. torch's current
 * stream) or NULL for a private one.  max_cols bounds n_cols of later calls. */
int bnk_ws_create(const bnk_model *m, const bnk_tree *t, int32_t max_cols, int device,
                  void *stream, bnk_ws **out);
void bnk_ws_destroy(bnk_ws *ws);
int bnk_ws_sync(bnk_ws *ws);
/* n_nodes of the tree, its ancestors (nodes with children) and the model's states; any output may be NULL */
int bnk_ws_dims(const bnk_ws *ws, int32_t *n_nodes, int32_t *n_internal, int32_t *n_states);
/* tuning knobs (0 = automatic): columns per CTA tile, columns per thread */
int bnk_ws_configure(bnk_ws *ws, int tile_cols, int cols_per_thread);
/* Where may observed states sit?  AUTO (default) scans the observation matrix on the device in every call: it
 * selects the general kernels when a node with children carries evidence (SubstNode.makeDenseFactor cases 2-4,
 * src/bn/ctmc/SubstNode.java:538-590) and rejects bytes that are neither BNK_NONE nor < K (BNK_ERR_ARG) -- at
 * the price of one 4-byte read-back + stream synchronisation per call.  LEAVES / ANYWHERE take the caller's
 * word instead (GRASP only instantiates extants, POGTree.getNodeInstance src/dat/pog/POGTree.java:419-436):
 * nothing is scanned or validated, *_dev calls only enqueue work and can be captured in a CUDA graph. */
#define BNK_EVIDENCE_AUTO 0
#define BNK_EVIDENCE_LEAVES 1
#define BNK_EVIDENCE_ANYWHERE 2
int bnk_ws_set_evidence_mode(bnk_ws *ws, int mode);
/* Host-buffer calls (bnk_joint / bnk_marginal / bnk_loglik) on wide alignments run as column chunks through
 * two internal workspaces on their own streams, so the result copy of a chunk overlaps the kernels of the
 * next.  chunk_cols: 0 = automatic (default), < 0 = never chunk, > 0 = columns per chunk. */
int bnk_ws_set_pipeline(bnk_ws *ws, int32_t chunk_cols);
/* kernels launched by this workspace since creation (bench.py's gpu_launches) */
int64_t bnk_ws_launch_count(const bnk_ws *ws);
/* bytes of device memory currently held */
int64_t bnk_ws_device_bytes(const bnk_ws *ws);

/* Per-column evolutionary rates (Prediction.java:615-618: null => 1.0 for every column;
 * PhyloBN.java:158: t = distance * rate).  Distinct values become rate categories; the
 * P(t) / log P(t) tables of every (branch, category) are built on the device. */
int bnk_set_rates(bnk_ws *ws, int32_t n_cols, const double *rates_or_null);
/* Replace branch lengths (optimisation loops); tables are rebuilt at the next run. */
int bnk_set_distances(bnk_ws *ws, const double *dist);

/* Device-side SubstModel.getProbs (SubstModel.java:303-327) for n_t times:
 * P [n_t][K][K] (P[i][j] = P(child=j | parent=i)), logP likewise; either may be NULL. */
int bnk_probs(bnk_ws *ws, int32_t n_t, const double *t, double *P, double *logP);

/* ---- the hot path, host buffers (what the JNI stub calls) ----------------------------- */
/* obs      [n_nodes][n_cols]  TreeInstance per column (TreeInstance.java:21-57), BNK_NONE = null
 * present  [n_nodes][n_cols]  1 = node kept in this column's pruned tree (Prediction.getTree,
 *                             Prediction.java:329-402); NULL = every node in every column
 * out_state[n_nodes][n_cols]  MaxLhoodJoint.getDecoration per node (MaxLhoodJoint.java:157-167):
 *                             inferred state, the observed state passed through, or BNK_NONE */
int bnk_joint(bnk_ws *ws, int32_t n_cols, const uint8_t *obs, const uint8_t *present_or_null,
              uint8_t *out_state);
/* query_nodes: n_query node indices, or NULL with n_query = -1 for every internal node in
 * index order.  A node may be listed more than once; a childless node that is present but not instantiated
 * (possible after orphan flooding, IdxTree.java:429-432) gets its distribution like any other variable.
 * out_post [n_query][n_cols][K] (NaN where the node has no posterior in that column: absent, observed or
 * unconnected); out_logl [n_cols] = VarElim.logLikelihood() of the column; either may be NULL. */
int bnk_marginal(bnk_ws *ws, int32_t n_cols, const uint8_t *obs, const uint8_t *present_or_null,
                 const int32_t *query_nodes, int32_t n_query, double *out_post, double *out_logl);
/* bnk_joint and bnk_marginal of the same alignment in one call (Prediction.getJoint + getMarginal,
 * src/asr/Prediction.java:555-649): same results; the joint pass of every column chunk runs while the previous
 * chunk's posteriors cross PCIe, so the call costs what bnk_marginal alone costs. */
int bnk_joint_marginal(bnk_ws *ws, int32_t n_cols, const uint8_t *obs, const uint8_t *present_or_null, uint8_t *out_state,
                       const int32_t *query_nodes, int32_t n_query, double *out_post, double *out_logl);
/* out_logl [n_cols] (may be NULL), *out_sum = sum over columns (may be NULL) */
int bnk_loglik(bnk_ws *ws, int32_t n_cols, const uint8_t *obs, const uint8_t *present_or_null,
               double *out_logl, double *out_sum);

/* ---- the hot path, device-resident buffers (bench "value", multi-GPU shards) ---------- */
/* Same semantics; every pointer is a device pointer valid on the workspace's device.
 * Launches are asynchronous on the workspace stream. */
int bnk_joint_dev(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present_or_null,
                  uint8_t *d_out_state);
int bnk_marginal_dev(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present_or_null,
                     const int32_t *query_nodes_host, int32_t n_query, double *d_out_post,
                     double *d_out_logl);
/* Joint AND marginal reconstruction of the same alignment in one call -- what a run with both inference modes
 * asks for (asr.GRASP `-m Joint` then `-m Marginal`, src/asr/GRASP.java; Prediction.getJoint + getMarginal,
 * src/asr/Prediction.java:555-649).  Results are those of bnk_joint_dev followed by bnk_marginal_dev; the two
 * passes are independent and run side by side on two streams inside the library (joined before return of control
 * to the workspace stream), which is what makes the call cheaper than the two it replaces. */
int bnk_joint_marginal_dev(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present_or_null,
                           uint8_t *d_out_state, const int32_t *query_nodes_host, int32_t n_query,
                           double *d_out_post, double *d_out_logl);
/* d_out_sum: one double on the device, = sum of the column log-likelihoods (input to the
 * NCCL all-reduce of an optimisation loop) */
int bnk_loglik_dev(bnk_ws *ws, int32_t n_cols, const uint8_t *d_obs, const uint8_t *d_present_or_null,
                   double *d_out_logl, double *d_out_sum);

/* ---- several devices in one process ------------------------------------------------------- */
/* The reference spreads columns over a thread pool (ThreadedDecorators, src/asr/ThreadedDecorators.java:28-72,
 * sized by GRASP.NTHREADS, src/asr/GRASP.java:35); bnk_mgpu spreads them over devices: contiguous blocks of
 * n_cols / n_devices columns (sizes differ by at most one), one workspace + streams + host thread per device, results written straight
 * into the caller's rows (no gather).  devices_or_null: CUDA device ordinals (NULL: 0 .. n_devices-1;
 * n_devices <= 0: every visible sm_100 device).  Same argument meaning as the single-device calls. */
typedef struct bnk_mgpu bnk_mgpu;
int bnk_mgpu_create(const bnk_model *m, const bnk_tree *t, int32_t max_cols, int n_devices,
                    const int *devices_or_null, bnk_mgpu **out);
void bnk_mgpu_destroy(bnk_mgpu *g);
int bnk_mgpu_n_devices(const bnk_mgpu *g);
/* the column block device idx receives for an alignment of n_cols columns */
int bnk_mgpu_shard(const bnk_mgpu *g, int32_t n_cols, int idx, int32_t *col0, int32_t *ncols);
int bnk_mgpu_set_rates(bnk_mgpu *g, int32_t n_cols, const double *rates_or_null);
int bnk_mgpu_set_distances(bnk_mgpu *g, const double *dist);
int bnk_mgpu_joint(bnk_mgpu *g, int32_t n_cols, const uint8_t *obs, const uint8_t *present_or_null,
                   uint8_t *out_state);
int bnk_mgpu_marginal(bnk_mgpu *g, int32_t n_cols, const uint8_t *obs, const uint8_t *present_or_null,
                      const int32_t *query_nodes, int32_t n_query, double *out_post, double *out_logl);
int bnk_mgpu_loglik(bnk_mgpu *g, int32_t n_cols, const uint8_t *obs, const uint8_t *present_or_null,
                    double *out_logl, double *out_sum);
/* Optimisation loops (the shape of IndelPeeler.optimiseMuLambda, src/asr/IndelPeeler.java:469-510: evaluate all
 * columns, sum, iterate): upload the alignment once, then per evaluation change rates / distances and call
 * bnk_mgpu_loglik_resident -- kernels on every device, then ONE all-reduce of the per-device sums
 * (ncclAllReduce over NVLink when libnccl.so.2 can be loaded; bnk_mgpu_reduction() says which path ran). */
int bnk_mgpu_upload(bnk_mgpu *g, int32_t n_cols, const uint8_t *obs, const uint8_t *present_or_null);
int bnk_mgpu_loglik_resident(bnk_mgpu *g, double *out_sum);
const char *bnk_mgpu_reduction(const bnk_mgpu *g);
int64_t bnk_mgpu_launch_count(const bnk_mgpu *g);

/* ---- gap-augmented peeling: indel-rate likelihoods ----------------------------
 * What asr.IndelPeeler computes one column at a time on a thread pool (src/asr/IndelPeeler.java, Rivas & Eddy
 * 2008) under bn.ctmc.GapSubstModel (src/bn/ctmc/GapSubstModel.java): the alphabet gains a gap state, gaps at
 * extants are evidence (not pruned positions), and the column likelihood sums over residue and gap histories.
 *   bnk_indel_create          one tree rooted at node 0 + the base substitution model (20 or 4 states, built
 *                             from a rate matrix: the gap model needs R); max_cols bounds the alignment width
 *   bnk_indel_set_alignment   obs [n_nodes][n_cols], BNK_NONE = gap; only the extants' rows are read
 *                             (new POGTree(aln, tree), IndelPeeler.java:540)
 *   bnk_indel_set_rates       new GapSubstModel(F, IRM, alphabet, mu, lambda) (GapSubstModel.java:26-50)
 *   bnk_indel_column_logl     IndelPeeler.decorate() (:229-252) for every column at indel rate `rate`
 *                             (branch time = distance * rate, PhyloBN.java:158); geom = the geometric
 *                             sequence-length parameter; out_logl [n_cols]; *out_sum = their sum in column order
 *   bnk_indel_column_priors   IndelPeeler.computeColumnPriors (:93-123): out [n_cols][n_rates]
 *   bnk_indel_aln_logl        IndelPeeler.calcProbAlnGivenTree (:125-161): sum of the columns at rate 1
 *                             + probExtraCol (:163-193) - log(1 - P(all-gap column))
 *   bnk_indel_optimise_mulambda  IndelPeeler.optimiseMuLambda (:469-510): Minimise.brent over mu = lambda in
 *                             [min_val, max_val] (src/smile/math/special/Minimise.java:25-107); *n_evals = distinct
 *                             likelihood evaluations (each = new model + one sweep of all columns)
 *   bnk_indel_probs           test probe: P_eps(t) [(K+1)][(K+1)] = GapSubstModel.getProbs(t), {ksiT(t), gammaT(t)},
 *                             the K+1 stationary frequencies
 *   bnk_indel_phase_ms        CUDA-event time of the last sweep: 0 tables, 1 peeling kernels
 *   bnk_indel_base_model      the base model optimiseMuLambda builds from a model name (:474-503): F and IRM of
 *                             the named matrix, always symmetric + normalised (differs from bnk_model_create for
 *                             "JC" and "Yang"); unknown names -> BNK_ERR_MODEL (IllegalArgumentException) */
typedef struct bnk_indel bnk_indel;
int bnk_indel_base_model(const char *name, bnk_model **out);
int bnk_indel_create(const bnk_model *base, const bnk_tree *t, int32_t max_cols, int device, bnk_indel **out);
/* every device of the node behind one handle (GRASP.NTHREADS -> ThreadedPeeler, IndelPeeler.java:208-224): contiguous
 * column blocks, one per device, swept concurrently; every other bnk_indel_* call works on the handle unchanged and
 * sums over columns run in column order whatever the device count.  n_devices <= 0: all visible devices. */
int bnk_indel_create_multi(const bnk_model *base, const bnk_tree *t, int32_t max_cols, int n_devices,
                           const int *devices_or_null, bnk_indel **out);
void bnk_indel_destroy(bnk_indel *h);
int bnk_indel_set_alignment(bnk_indel *h, int32_t n_cols, const uint8_t *obs);
int bnk_indel_set_rates(bnk_indel *h, double mu, double lambda);
int bnk_indel_column_logl(bnk_indel *h, double rate, double geom, double *out_logl, double *out_sum_or_null);
int bnk_indel_column_priors(bnk_indel *h, double geom, int32_t n_rates, const double *rates, double *out);
int bnk_indel_aln_logl(bnk_indel *h, double geom, double *out);
int bnk_indel_optimise_mulambda(bnk_indel *h, double min_val, double max_val, double geom, double *out_mulambda,
                                int32_t *n_evals_or_null);
int bnk_indel_probs(bnk_indel *h, double t, double *out_P, double *out_ksi_gamma_or_null, double *out_F_or_null);
double bnk_indel_phase_ms(const bnk_indel *h, int phase);
int64_t bnk_indel_launch_count(const bnk_indel *h);
int64_t bnk_indel_device_bytes(const bnk_indel *h);

/* ---- instrumentation ------------------------------------------------------------------ */
/* CUDA-event time (ms) of the kernels of the most recent *_dev / host call, by phase:
 * 0 tables, 1 joint up-sweep, 2 traceback, 3 marginal up-sweep, 4 marginal down-sweep.
 * Requires bnk_ws_sync first. */
int bnk_ws_phase_ms(bnk_ws *ws, int phase, float *ms);

#ifdef __cplusplus
}
#endif
#endif
