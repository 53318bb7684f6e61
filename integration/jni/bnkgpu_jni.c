/* bnkgpu_jni.c -- JNI shim between asr.NativeInference (integration/java/asr/NativeInference.java) and the C ABI
 * of libbnkgpu (include/bnkgpu.h).  One function per native method; no state of its own.
 *
 *   gcc -shared -fPIC -I"$JAVA_HOME/include" -I"$JAVA_HOME/include/linux" -I../../include \
 *       bnkgpu_jni.c -o libbnkgpu_jni.so -L../../bnkit_b200 -lbnkgpu
 *
 * NOT compiled in this repository (the build image has no JDK / jni.h); syntax-checked against a stub header only.
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

static void *buf(JNIEnv *env, jobject b) { return b ? (*env)->GetDirectBufferAddress(env, b) : NULL; }

JNIEXPORT jint JNICALL Java_asr_NativeInference_deviceCount(JNIEnv *env, jclass c)
{
    (void)env; (void)c;
    return bnk_device_count();
}

JNIEXPORT jlong JNICALL Java_asr_NativeInference_modelCreate(JNIEnv *env, jclass c, jstring name, jdoubleArray F)
{
    (void)c;
    const char *nm = (*env)->GetStringUTFChars(env, name, NULL);
    jdouble *f = F ? (*env)->GetDoubleArrayElements(env, F, NULL) : NULL;
    bnk_model *m = NULL;
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
    jdouble *f = (*env)->GetDoubleArrayElements(env, F, NULL), *ev = (*env)->GetDoubleArrayElements(env, evec, NULL);
    jdouble *iv = (*env)->GetDoubleArrayElements(env, ievec, NULL), *el = (*env)->GetDoubleArrayElements(env, eval, NULL);
    bnk_model *m = NULL;
    int rc = bnk_model_create_from_eigen(K, f, ev, iv, el, &m);
    (*env)->ReleaseDoubleArrayElements(env, F, f, JNI_ABORT);
    (*env)->ReleaseDoubleArrayElements(env, evec, ev, JNI_ABORT);
    (*env)->ReleaseDoubleArrayElements(env, ievec, iv, JNI_ABORT);
    (*env)->ReleaseDoubleArrayElements(env, eval, el, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
    return (jlong)(intptr_t)m;
}

JNIEXPORT jlong JNICALL Java_asr_NativeInference_treeCreate(JNIEnv *env, jclass c, jintArray parents, jdoubleArray distances)
{
    (void)c;
    const jsize n = (*env)->GetArrayLength(env, parents);
    jint *p = (*env)->GetIntArrayElements(env, parents, NULL);
    jdouble *d = (*env)->GetDoubleArrayElements(env, distances, NULL);
    bnk_tree *t = NULL;
    int rc = bnk_tree_create((int32_t)n, (const int32_t *)p, d, &t);
    (*env)->ReleaseIntArrayElements(env, parents, p, JNI_ABORT);
    (*env)->ReleaseDoubleArrayElements(env, distances, d, JNI_ABORT);
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
    jdouble *r = rates ? (*env)->GetDoubleArrayElements(env, rates, NULL) : NULL;
    int rc = bnk_set_rates((bnk_ws *)(intptr_t)ws, nCols, r);
    if (r) (*env)->ReleaseDoubleArrayElements(env, rates, r, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_joint(JNIEnv *env, jclass c, jlong ws, jint nCols, jobject obs, jobject present,
                                                      jobject out)
{
    (void)c;
    int rc = bnk_joint((bnk_ws *)(intptr_t)ws, nCols, (const uint8_t *)buf(env, obs), (const uint8_t *)buf(env, present),
                       (uint8_t *)buf(env, out));
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_marginal(JNIEnv *env, jclass c, jlong ws, jint nCols, jobject obs, jobject present,
                                                         jintArray query, jobject outPost, jdoubleArray outLogl)
{
    (void)c;
    jint *q = query ? (*env)->GetIntArrayElements(env, query, NULL) : NULL;
    const jint nq = query ? (*env)->GetArrayLength(env, query) : -1;
    jdouble *ll = outLogl ? (*env)->GetDoubleArrayElements(env, outLogl, NULL) : NULL;
    int rc = bnk_marginal((bnk_ws *)(intptr_t)ws, nCols, (const uint8_t *)buf(env, obs), (const uint8_t *)buf(env, present),
                          (const int32_t *)q, nq, (double *)buf(env, outPost), ll);
    if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, rc == BNK_OK ? 0 : JNI_ABORT);
    if (q) (*env)->ReleaseIntArrayElements(env, query, q, JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT jdouble JNICALL Java_asr_NativeInference_logLik(JNIEnv *env, jclass c, jlong ws, jint nCols, jobject obs, jobject present,
                                                          jdoubleArray outLogl)
{
    (void)c;
    double sum = 0.0;
    jdouble *ll = outLogl ? (*env)->GetDoubleArrayElements(env, outLogl, NULL) : NULL;
    int rc = bnk_loglik((bnk_ws *)(intptr_t)ws, nCols, (const uint8_t *)buf(env, obs), (const uint8_t *)buf(env, present), ll, &sum);
    if (ll) (*env)->ReleaseDoubleArrayElements(env, outLogl, ll, rc == BNK_OK ? 0 : JNI_ABORT);
    if (rc != BNK_OK) raise(env, rc);
    return sum;
}

JNIEXPORT void JNICALL Java_asr_NativeInference_bepPresence(JNIEnv *env, jclass c, jlong tree, jint nCols, jobject occupancy,
                                                            jint device, jobject present)
{
    (void)c;
    int rc = bnk_bep_presence((const bnk_tree *)(intptr_t)tree, nCols, (const uint8_t *)buf(env, occupancy), device,
                              (uint8_t *)buf(env, present), NULL);
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_pruneOrphans(JNIEnv *env, jclass c, jlong tree, jint nCols, jobject in, jobject out)
{
    (void)c;
    int rc = bnk_tree_prune_orphans((const bnk_tree *)(intptr_t)tree, nCols, (const uint8_t *)buf(env, in), (uint8_t *)buf(env, out));
    if (rc != BNK_OK) raise(env, rc);
}

JNIEXPORT void JNICALL Java_asr_NativeInference_destroy(JNIEnv *env, jclass c, jlong model, jlong tree, jlong ws)
{
    (void)env; (void)c;
    if (ws) bnk_ws_destroy((bnk_ws *)(intptr_t)ws);
    if (tree) bnk_tree_destroy((bnk_tree *)(intptr_t)tree);
    if (model) bnk_model_destroy((bnk_model *)(intptr_t)model);
}
