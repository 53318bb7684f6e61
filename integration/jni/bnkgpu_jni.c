/* bnkgpu_jni.c -- JNI shim between asr.NativeInference (integration/java/asr/NativeInference.java) and the C ABI
 * of libbnkgpu (include/bnkgpu.h).  One function per native method; no state of its own.
 *
 *   gcc -shared -fPIC -I"$JAVA_HOME/include" -I"$JAVA_HOME/include/linux" -I../../include \
 *       bnkgpu_jni.c -o libbnkgpu_jni.so -L../../bnkit_b200 -lbnkgpu
 *
 * NOT compiled in this repository (the build image has no JDK / jni.h); syntax-checked against a stub header only.
 * Every direct buffer is checked for being direct and large enough, every array for its length, every Get*
 * for NULL, before the library is called.
 * Error mapping: BNK_ERR_QUERY -> asr.ASRRuntimeException (Prediction.java:561-562),
 * BNK_ERR_ARG / BNK_ERR_TREE -> IllegalArgumentException, BNK_ERR_MODEL -> ArithmeticException, anything else ->
 * RuntimeException, always with bnk_last_error() as the message. */
#include <jni.h>
#include <stdint.h>
#include <stddef.h>
#include "bnkgpu.h"

static void raise(JNIEnv *env, int rc)
{
    const char *cls = (rc == BNK_ERR_QUERY) ? "asr/ASRRuntimeException"
                    : (rc == BNK_ERR_ARG || rc == BNK_ERR_TREE) ? "java/lang/IllegalArgumentException"
                    : (rc == BNK_ERR_MODEL) ? "java/lang/ArithmeticException"
                    : (rc == BNK_ERR_NOMEM) ? "java/lang/OutOfMemoryError" : "java/lang/RuntimeException";
    jclass c = (*env)->FindClass(env, cls);
    if (c) (*env)->ThrowNew(env, c, bnk_last_error());
}

static void throw_arg(JNIEnv *env, const char *msg)
{
    jclass c = (*env)->FindClass(env, "java/lang/IllegalArgumentException");
    if (c) (*env)->ThrowNew(env, c, msg);
}

/* Address of a direct buffer holding at least `need` elements (bytes of a ByteBuffer, doubles of a DoubleBuffer).
 * NULL object = "not given" (allowed where the ABI takes NULL).  A heap buffer or a short one raises
 * IllegalArgumentException and sets *bad: a non-direct `present` must not silently mean "everything present". */
static void *buf(JNIEnv *env, jobject b, jlong need, int *bad)
{
    void *p;
    if (!b) return NULL;
    p = (*env)->GetDirectBufferAddress(env, b);
    if (!p) { throw_arg(env, "bnkgpu: a direct (allocateDirect) buffer is required"); *bad = 1; return NULL; }
    if ((*env)->GetDirectBufferCapacity(env, b) < need) { throw_arg(env, "bnkgpu: buffer too small for n_nodes x n_cols"); *bad = 1; return NULL; }
    return p;
}


JNIEXPORT jint JNICALL Java_asr_NativeInference_deviceCount(JNIEnv *env, jclass c)
{
    (void)env; (void)c;
    return bnk_device_count();
}

JNIEXPORT jlong JNICALL Java_asr_NativeInference_modelCreate(JNIEnv *env, jclass c, jstring name, jdoubleArray F)
{
    (void)c;
    const char *nm = (*env)->GetStringUTFChars(env, name, NULL);
    if (!nm) return 0;   /* OutOfMemoryError is pending */
    jdouble *f = F ? (*env)->GetDoubleArrayElements(env, F, NULL) : NULL;
    bnk_model *m = NULL;
    if (F && (!f || (*env)->GetArrayLength(env, F) < 4)) {
        if (f) (*env)->ReleaseDoubleArrayElements(env, F, f, JNI_ABORT);
        (*env)->ReleaseStringUTFChars(env, name, nm);
        if (f) throw_arg(env, "bnkgpu: F needs one entry per state");
        return 0;
    }
    int rc = bnk_model_create(nm, f, &m);
    if (f) (*env)->ReleaseDoubleArrayElements(env, F, f, JNI_ABORT);
    (*env)->ReleaseStringUTFChars(env, name, nm);
    if (rc != BNK_OK) raise(env, rc);
    return (jlong)(intptr_t)m;
}

JNIEXPORT jlong JNICALL Java_asr_NativeInference_modelFromEigen(JNIEnv *env, jclass c, jint K, jdoubleArray F,
                                                                jdoubleArray evec, jdoubleArray ievec, jdoubleArray eval)
{
    (void)c;
    if (K < 2 || (*env)->GetArrayLength(env, F) < K || (*env)->GetArrayLength(env, eval) < K ||
        (*env)->GetArrayLength(env, evec) < K * K || (*env)->GetArrayLength(env, ievec) < K * K) {
        throw_arg(env, "bnkgpu: F / eval need K entries, evec / ievec K x K");
        return 0;
    }
    jdouble *f = (*env)->GetDoubleArrayElements(env, F, NULL), *ev = (*env)->GetDoubleArrayElements(env, evec, NULL);
    jdouble *iv = (*env)->GetDoubleArrayElements(env, ievec, NULL), *el = (*env)->GetDoubleArrayElements(env, eval, NULL);
    bnk_model *m = NULL;
    int rc = (f && ev && iv && el) ? bnk_model_create_from_eigen(K, f, ev, iv, el, &m) : BNK_ERR_NOMEM;
    if (f) (*env)->ReleaseDoubleArrayElements(env, F, f, JNI_ABORT);
    if (ev) (*env)->ReleaseDoubleArrayElements(env, evec, ev, JNI_ABORT);
    if (iv) (*env)->ReleaseDoubleArrayElements(env, ievec, iv, JNI_ABORT);
    if (el) (*env)->ReleaseDoubleArrayElements(env, eval, el, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
    return (jlong)(intptr_t)m;
}

JNIEXPORT jlong JNICALL Java_asr_NativeInference_treeCreate(JNIEnv *env, jclass c, jintArray parents, jdoubleArray distances)
{
    (void)c;
    const jsize n = (*env)->GetArrayLength(env, parents);
    if ((*env)->GetArrayLength(env, distances) < n) { throw_arg(env, "bnkgpu: distances shorter than parents"); return 0; }
    jint *p = (*env)->GetIntArrayElements(env, parents, NULL);
    jdouble *d = (*env)->GetDoubleArrayElements(env, distances, NULL);
    bnk_tree *t = NULL;
    int rc = (p && d) ? bnk_tree_create((int32_t)n, (const int32_t *)p, d, &t) : BNK_ERR_NOMEM;
    if (p) (*env)->ReleaseIntArrayElements(env, parents, p, JNI_ABORT);
    if (d) (*env)->ReleaseDoubleArrayElements(env, distances, d, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
    return (jlong)(intptr_t)t;
}

JNIEXPORT jlong JNICALL Java_asr_NativeInference_wsCreate(JNIEnv *env, jclass c, jlong model, jlong tree, jint maxCols, jint device)
{
    (void)c;
    bnk_ws *ws = NULL;
    int rc = bnk_ws_create((const bnk_model *)(intptr_t)model, (const bnk_tree *)(intptr_t)tree, maxCols, device, NULL, &ws);
    if (rc != BNK_OK) raise(env, rc);
    return (jlong)(intptr_t)ws;
}

JNIEXPORT void JNICALL Java_asr_NativeInference_setRates(JNIEnv *env, jclass c, jlong ws, jint nCols, jdoubleArray rates)
{
    (void)c;
    if (rates && (*env)->GetArrayLength(env, rates) < nCols) { throw_arg(env, "bnkgpu: rates shorter than nCols"); return; }
    jdouble *r = rates ? (*env)->GetDoubleArrayElements(env, rates, NULL) : NULL;
    if (rates && !r) return;
    int rc = bnk_set_rates((bnk_ws *)(intptr_t)ws, nCols, r);
    if (r) (*env)->ReleaseDoubleArrayElements(env, rates, r, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
}

/* n_nodes / n_internal / n_states of a workspace, for the capacity checks */
static int dims(JNIEnv *env, jlong ws, int32_t *n, int32_t *ni, int32_t *K)
{
    int rc = bnk_ws_dims((const bnk_ws *)(intptr_t)ws, n, ni, K);
    if (rc != BNK_OK) raise(env, rc);
    return rc;
}

JNIEXPORT void JNICALL Java_asr_NativeInference_joint(JNIEnv *env, jclass c, jlong ws, jint nCols, jobject obs, jobject present,
                                                      jobject out)
{
    (void)c;
    int32_t n, ni, K;
    int bad = 0, rc;
    if (dims(env, ws, &n, &ni, &K) != BNK_OK) return;
    const uint8_t *o = (const uint8_t *)buf(env, obs, (jlong)n * nCols, &bad);
    const uint8_t *p = (const uint8_t *)buf(env, present, (jlong)n * nCols, &bad);
    uint8_t *st = (uint8_t *)buf(env, out, (jlong)n * nCols, &bad);
    if (bad) return;
    rc = bnk_joint((bnk_ws *)(intptr_t)ws, nCols, o, p, st);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_marginal(JNIEnv *env, jclass c, jlong ws, jint nCols, jobject obs, jobject present,
                                                         jintArray query, jobject outPost, jdoubleArray outLogl)
{
    (void)c;
    int32_t n, ni, K;
    int bad = 0, rc;
    if (dims(env, ws, &n, &ni, &K) != BNK_OK) return;
    const jint nq = query ? (*env)->GetArrayLength(env, query) : -1;
    const uint8_t *o = (const uint8_t *)buf(env, obs, (jlong)n * nCols, &bad);
    const uint8_t *p = (const uint8_t *)buf(env, present, (jlong)n * nCols, &bad);
    double *post = (double *)buf(env, outPost, (jlong)(nq < 0 ? ni : nq) * nCols * K, &bad);
    if (bad) return;
    if (outLogl && (*env)->GetArrayLength(env, outLogl) < nCols) { throw_arg(env, "bnkgpu: outLogl shorter than nCols"); return; }
    jint *q = query ? (*env)->GetIntArrayElements(env, query, NULL) : NULL;
    jdouble *ll = outLogl ? (*env)->GetDoubleArrayElements(env, outLogl, NULL) : NULL;
    if ((query && !q) || (outLogl && !ll)) {   /* OutOfMemoryError is pending */
        if (q) (*env)->ReleaseIntArrayElements(env, query, q, JNI_ABORT);
        if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, JNI_ABORT);
        return;
    }
    rc = bnk_marginal((bnk_ws *)(intptr_t)ws, nCols, o, p, (const int32_t *)q, nq, post, ll);
    if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, rc == BNK_OK ? 0 : JNI_ABORT);
    if (q) (*env)->ReleaseIntArrayElements(env, query, q, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
}

/* joint + marginal of the same alignment in one call (bnk_joint_marginal): the joint pass hides under the result copies */
JNIEXPORT void JNICALL Java_asr_NativeInference_jointMarginal(JNIEnv *env, jclass c, jlong ws, jint nCols, jobject obs, jobject present,
                                                              jobject outState, jintArray query, jobject outPost, jdoubleArray outLogl)
{
    (void)c;
    int32_t n, ni, K;
    int bad = 0, rc;
    if (dims(env, ws, &n, &ni, &K) != BNK_OK) return;
    const jint nq = query ? (*env)->GetArrayLength(env, query) : -1;
    const uint8_t *o = (const uint8_t *)buf(env, obs, (jlong)n * nCols, &bad);
    const uint8_t *p = (const uint8_t *)buf(env, present, (jlong)n * nCols, &bad);
    uint8_t *st = (uint8_t *)buf(env, outState, (jlong)n * nCols, &bad);
    double *post = (double *)buf(env, outPost, (jlong)(nq < 0 ? ni : nq) * nCols * K, &bad);
    if (bad) return;
    if (!st) { throw_arg(env, "bnkgpu: outState is null"); return; }
    if (outLogl && (*env)->GetArrayLength(env, outLogl) < nCols) { throw_arg(env, "bnkgpu: outLogl shorter than nCols"); return; }
    jint *q = query ? (*env)->GetIntArrayElements(env, query, NULL) : NULL;
    jdouble *ll = outLogl ? (*env)->GetDoubleArrayElements(env, outLogl, NULL) : NULL;
    if ((query && !q) || (outLogl && !ll)) {   /* OutOfMemoryError is pending */
        if (q) (*env)->ReleaseIntArrayElements(env, query, q, JNI_ABORT);
        if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, JNI_ABORT);
        return;
    }
    rc = bnk_joint_marginal((bnk_ws *)(intptr_t)ws, nCols, o, p, st, (const int32_t *)q, nq, post, ll);
    if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, rc == BNK_OK ? 0 : JNI_ABORT);
    if (q) (*env)->ReleaseIntArrayElements(env, query, q, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT jdouble JNICALL Java_asr_NativeInference_logLik(JNIEnv *env, jclass c, jlong ws, jint nCols, jobject obs, jobject present,
                                                          jdoubleArray outLogl)
{
    (void)c;
    int32_t n, ni, K;
    int bad = 0, rc;
    double sum = 0.0;
    if (dims(env, ws, &n, &ni, &K) != BNK_OK) return 0.0;
    const uint8_t *o = (const uint8_t *)buf(env, obs, (jlong)n * nCols, &bad);
    const uint8_t *p = (const uint8_t *)buf(env, present, (jlong)n * nCols, &bad);
    if (bad) return 0.0;
    if (outLogl && (*env)->GetArrayLength(env, outLogl) < nCols) { throw_arg(env, "bnkgpu: outLogl shorter than nCols"); return 0.0; }
    jdouble *ll = outLogl ? (*env)->GetDoubleArrayElements(env, outLogl, NULL) : NULL;
    if (outLogl && !ll) return 0.0;
    rc = bnk_loglik((bnk_ws *)(intptr_t)ws, nCols, o, p, ll, &sum);
    if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, rc == BNK_OK ? 0 : JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
    return sum;
}

JNIEXPORT void JNICALL Java_asr_NativeInference_bepPresence(JNIEnv *env, jclass c, jlong tree, jint nCols, jobject occupancy,
                                                            jint device, jobject present)
{
    (void)c;
    int bad = 0, rc;
    const jlong need = (jlong)bnk_tree_size((const bnk_tree *)(intptr_t)tree) * nCols;
    const uint8_t *occ = (const uint8_t *)buf(env, occupancy, need, &bad);
    uint8_t *pr = (uint8_t *)buf(env, present, need, &bad);
    if (bad) return;
    rc = bnk_bep_presence((const bnk_tree *)(intptr_t)tree, nCols, occ, device, pr, NULL);
    if (rc != BNK_OK) raise(env, rc);
}

/* Prediction.PredictByBidirEdgeParsimony incl. the ancestors' graphs: returns a bnk_pogs* (0 on error) */
JNIEXPORT jlong JNICALL Java_asr_NativeInference_bepInfer(JNIEnv *env, jclass c, jlong tree, jint nCols, jobject occupancy,
                                                          jint device, jobject present)
{
    (void)c;
    int bad = 0, rc;
    bnk_pogs *pogs = NULL;
    const jlong need = (jlong)bnk_tree_size((const bnk_tree *)(intptr_t)tree) * nCols;
    const uint8_t *occ = (const uint8_t *)buf(env, occupancy, need, &bad);
    uint8_t *pr = (uint8_t *)buf(env, present, need, &bad);
    if (bad) return 0;
    rc = bnk_bep_infer((const bnk_tree *)(intptr_t)tree, nCols, occ, device, pr, NULL, &pogs);
    if (rc != BNK_OK) { raise(env, rc); return 0; }
    return (jlong)(intptr_t)pogs;
}

JNIEXPORT jint JNICALL Java_asr_NativeInference_pogEdgeCount(JNIEnv *env, jclass c, jlong pogs, jint node)
{
    (void)c;
    int n = bnk_pogs_n_edges((const bnk_pogs *)(intptr_t)pogs, node);
    if (n < 0) raise(env, n);
    return n;
}

/* from / to: int[edge count]; flags: direct ByteBuffer (bit 0 forward, bit 1 backward) */
JNIEXPORT void JNICALL Java_asr_NativeInference_pogEdges(JNIEnv *env, jclass c, jlong pogs, jint node, jintArray from, jintArray to,
                                                         jobject flags)
{
    (void)c;
    int bad = 0;
    const int n = bnk_pogs_n_edges((const bnk_pogs *)(intptr_t)pogs, node);
    if (n < 0) { raise(env, n); return; }
    if ((*env)->GetArrayLength(env, from) < n || (*env)->GetArrayLength(env, to) < n) { throw_arg(env, "bnkgpu: edge arrays too short"); return; }
    uint8_t *fl = (uint8_t *)buf(env, flags, n, &bad);
    if (bad) return;
    jint *f = (*env)->GetIntArrayElements(env, from, NULL), *t = (*env)->GetIntArrayElements(env, to, NULL);
    if (f && t) bnk_pogs_edges((const bnk_pogs *)(intptr_t)pogs, node, (int32_t *)f, (int32_t *)t, fl);
    if (f) (*env)->ReleaseIntArrayElements(env, from, f, 0);
    if (t) (*env)->ReleaseIntArrayElements(env, to, t, 0);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_pogsDestroy(JNIEnv *env, jclass c, jlong pogs)
{
    (void)env; (void)c;
    if (pogs) bnk_pogs_destroy((bnk_pogs *)(intptr_t)pogs);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_pruneOrphans(JNIEnv *env, jclass c, jlong tree, jint nCols, jobject in, jobject out)
{
    (void)c;
    int bad = 0, rc;
    const jlong need = (jlong)bnk_tree_size((const bnk_tree *)(intptr_t)tree) * nCols;
    const uint8_t *pi = (const uint8_t *)buf(env, in, need, &bad);
    uint8_t *po = (uint8_t *)buf(env, out, need, &bad);
    if (bad) return;
    rc = bnk_tree_prune_orphans((const bnk_tree *)(intptr_t)tree, nCols, pi, po);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_destroy(JNIEnv *env, jclass c, jlong model, jlong tree, jlong ws)
{
    (void)env; (void)c;
    if (ws) bnk_ws_destroy((bnk_ws *)(intptr_t)ws);
    if (tree) bnk_tree_destroy((bnk_tree *)(intptr_t)tree);
    if (model) bnk_model_destroy((bnk_model *)(intptr_t)model);
}

/* ---- all devices of the node behind one handle: GRASP.NTHREADS -> ThreadedDecorators (src/asr/ThreadedDecorators.java:28-72) ---- */
JNIEXPORT jlong JNICALL Java_asr_NativeInference_mgpuCreate(JNIEnv *env, jclass c, jlong model, jlong tree, jint maxCols, jint nDevices)
{
    (void)c;
    bnk_mgpu *g = NULL;
    int rc = bnk_mgpu_create((const bnk_model *)(intptr_t)model, (const bnk_tree *)(intptr_t)tree, maxCols, nDevices, NULL, &g);
    if (rc != BNK_OK) raise(env, rc);
    return (jlong)(intptr_t)g;
}

JNIEXPORT void JNICALL Java_asr_NativeInference_mgpuSetRates(JNIEnv *env, jclass c, jlong g, jint nCols, jdoubleArray rates)
{
    (void)c;
    if (rates && (*env)->GetArrayLength(env, rates) < nCols) { throw_arg(env, "bnkgpu: rates shorter than nCols"); return; }
    jdouble *r = rates ? (*env)->GetDoubleArrayElements(env, rates, NULL) : NULL;
    if (rates && !r) return;
    int rc = bnk_mgpu_set_rates((bnk_mgpu *)(intptr_t)g, nCols, r);
    if (r) (*env)->ReleaseDoubleArrayElements(env, rates, r, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_mgpuJoint(JNIEnv *env, jclass c, jlong g, jint nNodes, jint nCols, jobject obs,
                                                          jobject present, jobject out)
{
    (void)c;
    int bad = 0, rc;
    const uint8_t *o = (const uint8_t *)buf(env, obs, (jlong)nNodes * nCols, &bad);
    const uint8_t *p = (const uint8_t *)buf(env, present, (jlong)nNodes * nCols, &bad);
    uint8_t *st = (uint8_t *)buf(env, out, (jlong)nNodes * nCols, &bad);
    if (bad) return;
    rc = bnk_mgpu_joint((bnk_mgpu *)(intptr_t)g, nCols, o, p, st);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_mgpuMarginal(JNIEnv *env, jclass c, jlong g, jint nNodes, jint nStates, jint nCols,
                                                             jobject obs, jobject present, jintArray query, jobject outPost,
                                                             jdoubleArray outLogl)
{
    (void)c;
    int bad = 0, rc;
    if (!query) { throw_arg(env, "bnkgpu: mgpuMarginal needs the list of query nodes"); return; }
    const jint nq = (*env)->GetArrayLength(env, query);
    const uint8_t *o = (const uint8_t *)buf(env, obs, (jlong)nNodes * nCols, &bad);
    const uint8_t *p = (const uint8_t *)buf(env, present, (jlong)nNodes * nCols, &bad);
    double *post = (double *)buf(env, outPost, (jlong)nq * nCols * nStates, &bad);
    if (bad) return;
    if (outLogl && (*env)->GetArrayLength(env, outLogl) < nCols) { throw_arg(env, "bnkgpu: outLogl shorter than nCols"); return; }
    jint *q = (*env)->GetIntArrayElements(env, query, NULL);
    jdouble *ll = outLogl ? (*env)->GetDoubleArrayElements(env, outLogl, NULL) : NULL;
    if (!q || (outLogl && !ll)) {
        if (q) (*env)->ReleaseIntArrayElements(env, query, q, JNI_ABORT);
        if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, JNI_ABORT);
        return;
    }
    rc = bnk_mgpu_marginal((bnk_mgpu *)(intptr_t)g, nCols, o, p, (const int32_t *)q, nq, post, ll);
    if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, rc == BNK_OK ? 0 : JNI_ABORT);
    (*env)->ReleaseIntArrayElements(env, query, q, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT jdouble JNICALL Java_asr_NativeInference_mgpuLogLik(JNIEnv *env, jclass c, jlong g, jint nNodes, jint nCols, jobject obs,
                                                              jobject present, jdoubleArray outLogl)
{
    (void)c;
    int bad = 0, rc;
    double sum = 0.0;
    const uint8_t *o = (const uint8_t *)buf(env, obs, (jlong)nNodes * nCols, &bad);
    const uint8_t *p = (const uint8_t *)buf(env, present, (jlong)nNodes * nCols, &bad);
    if (bad) return 0.0;
    if (outLogl && (*env)->GetArrayLength(env, outLogl) < nCols) { throw_arg(env, "bnkgpu: outLogl shorter than nCols"); return 0.0; }
    jdouble *ll = outLogl ? (*env)->GetDoubleArrayElements(env, outLogl, NULL) : NULL;
    if (outLogl && !ll) return 0.0;
    rc = bnk_mgpu_loglik((bnk_mgpu *)(intptr_t)g, nCols, o, p, ll, &sum);
    if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, rc == BNK_OK ? 0 : JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
    return sum;
}

JNIEXPORT void JNICALL Java_asr_NativeInference_mgpuDestroy(JNIEnv *env, jclass c, jlong g)
{
    (void)env; (void)c;
    if (g) bnk_mgpu_destroy((bnk_mgpu *)(intptr_t)g);
}

/* ---- gap-augmented peeling: asr.IndelPeeler (src/asr/IndelPeeler.java) ---- */
JNIEXPORT jlong JNICALL Java_asr_NativeInference_indelBaseModel(JNIEnv *env, jclass c, jstring name)
{
    (void)c;
    const char *nm = (*env)->GetStringUTFChars(env, name, NULL);
    bnk_model *m = NULL;
    if (!nm) return 0;
    int rc = bnk_indel_base_model(nm, &m);
    (*env)->ReleaseStringUTFChars(env, name, nm);
    if (rc == BNK_ERR_MODEL) { throw_arg(env, bnk_last_error()); return 0; }   /* IllegalArgumentException, IndelPeeler.java:502 */
    if (rc != BNK_OK) raise(env, rc);
    return (jlong)(intptr_t)m;
}

JNIEXPORT jlong JNICALL Java_asr_NativeInference_indelCreate(JNIEnv *env, jclass c, jlong model, jlong tree, jint nCols, jobject obs,
                                                             jint device)
{
    (void)c;
    int bad = 0, rc;
    bnk_indel *h = NULL;
    const uint8_t *o = (const uint8_t *)buf(env, obs, (jlong)bnk_tree_size((const bnk_tree *)(intptr_t)tree) * nCols, &bad);
    if (bad) return 0;
    if (!o) { throw_arg(env, "bnkgpu: the alignment buffer is required"); return 0; }
    rc = bnk_indel_create((const bnk_model *)(intptr_t)model, (const bnk_tree *)(intptr_t)tree, nCols, device, &h);
    if (rc == BNK_OK) rc = bnk_indel_set_alignment(h, nCols, o);
    if (rc != BNK_OK) { if (h) bnk_indel_destroy(h); raise(env, rc); return 0; }
    return (jlong)(intptr_t)h;
}

/* the same over nDevices devices (0 = all visible): GRASP.NTHREADS in IndelPeeler.runPeelingJobs */
JNIEXPORT jlong JNICALL Java_asr_NativeInference_indelCreateMulti(JNIEnv *env, jclass c, jlong model, jlong tree, jint nCols, jobject obs,
                                                                  jint nDevices)
{
    (void)c;
    int bad = 0, rc;
    bnk_indel *h = NULL;
    const uint8_t *o = (const uint8_t *)buf(env, obs, (jlong)bnk_tree_size((const bnk_tree *)(intptr_t)tree) * nCols, &bad);
    if (bad) return 0;
    if (!o) { throw_arg(env, "bnkgpu: the alignment buffer is required"); return 0; }
    rc = bnk_indel_create_multi((const bnk_model *)(intptr_t)model, (const bnk_tree *)(intptr_t)tree, nCols, nDevices, NULL, &h);
    if (rc == BNK_OK) rc = bnk_indel_set_alignment(h, nCols, o);
    if (rc != BNK_OK) { if (h) bnk_indel_destroy(h); raise(env, rc); return 0; }
    return (jlong)(intptr_t)h;
}

JNIEXPORT void JNICALL Java_asr_NativeInference_indelSetRates(JNIEnv *env, jclass c, jlong h, jdouble mu, jdouble lambda)
{
    (void)c;
    int rc = bnk_indel_set_rates((bnk_indel *)(intptr_t)h, mu, lambda);
    if (rc != BNK_OK) raise(env, rc);
}

/* computeColumnPriors: out = double[nCols * nRates], row per column */
JNIEXPORT void JNICALL Java_asr_NativeInference_indelColumnPriors(JNIEnv *env, jclass c, jlong h, jdouble geom, jint nCols,
                                                                  jdoubleArray rates, jdoubleArray out)
{
    (void)c;
    const jint nr = (*env)->GetArrayLength(env, rates);
    if ((jlong)(*env)->GetArrayLength(env, out) < (jlong)nCols * nr) { throw_arg(env, "bnkgpu: out shorter than nCols x nRates"); return; }
    jdouble *r = (*env)->GetDoubleArrayElements(env, rates, NULL), *o = (*env)->GetDoubleArrayElements(env, out, NULL);
    int rc = BNK_OK;
    if (r && o) rc = bnk_indel_column_priors((bnk_indel *)(intptr_t)h, geom, nr, r, o);
    if (r) (*env)->ReleaseDoubleArrayElements(env, rates, r, JNI_ABORT);
    if (o) (*env)->ReleaseDoubleArrayElements(env, out, o, rc == BNK_OK ? 0 : JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT jdouble JNICALL Java_asr_NativeInference_indelAlnLogl(JNIEnv *env, jclass c, jlong h, jdouble geom)
{
    (void)c;
    double v = 0.0;
    int rc = bnk_indel_aln_logl((bnk_indel *)(intptr_t)h, geom, &v);
    if (rc != BNK_OK) raise(env, rc);
    return v;
}

JNIEXPORT jdouble JNICALL Java_asr_NativeInference_indelOptimiseMuLambda(JNIEnv *env, jclass c, jlong h, jdouble minVal, jdouble maxVal,
                                                                         jdouble geom)
{
    (void)c;
    double x = 0.0;
    int rc = bnk_indel_optimise_mulambda((bnk_indel *)(intptr_t)h, minVal, maxVal, geom, &x, NULL);
    if (rc != BNK_OK) raise(env, rc);
    return x;
}

JNIEXPORT void JNICALL Java_asr_NativeInference_indelDestroy(JNIEnv *env, jclass c, jlong h)
{
    (void)env; (void)c;
    if (h) bnk_indel_destroy((bnk_indel *)(intptr_t)h);
}
