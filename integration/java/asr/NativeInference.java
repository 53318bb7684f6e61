// NativeInference.java -- the Java half of the binding INTEGRATION.md describes: what a bnkit maintainer would add
// under src/asr/ to let asr.Prediction.getJoint / getMarginal (src/asr/Prediction.java:555-649) and
// Prediction.PredictByBidirEdgeParsimony (:1443-1590) call libbnkgpu through JNI.
// NOT compiled in this repository (the build image has no JDK); kept in sync with include/bnkgpu.h by hand.
package asr;

import java.nio.ByteBuffer;
import java.nio.DoubleBuffer;

public final class NativeInference {
    private static final boolean LOADED;
    static {
        boolean ok;
        try {
            System.loadLibrary("bnkgpu_jni");   // links libbnkgpu.so
            ok = deviceCount() > 0;
        } catch (UnsatisfiedLinkError e) {
            ok = false;
        }
        LOADED = ok;
    }

    private NativeInference() { }

    /** true when the library loaded and an sm_100 device is visible; callers fall back to ThreadedDecorators otherwise. */
    public static boolean available() { return LOADED; }

    static native int deviceCount();
    /** SubstModel.createModel(name): "LG", "JTT", "WAG", "Dayhoff", "JC", "Yang"; F may be null. Returns a bnk_model*. */
    static native long modelCreate(String name, double[] F);
    /** evec / ievec: K*K row-major as SubstModel.getRexp() holds them (Matrix.java:63-77). Returns a bnk_model*. */
    static native long modelFromEigen(int K, double[] F, double[] evec, double[] ievec, double[] eval);
    /** IdxTree.getParents() / getDistances() (IdxTree.java:27-32). Returns a bnk_tree*. */
    static native long treeCreate(int[] parents, double[] distances);
    static native long wsCreate(long model, long tree, int maxCols, int device);
    /** rates == null: 1.0 for every column (Prediction.java:615-618). */
    static native void setRates(long ws, int nCols, double[] ratesOrNull);
    /** obs, present, out: direct ByteBuffers of n_nodes * n_cols bytes, one row per branch point; 255 = null. */
    static native void joint(long ws, int nCols, ByteBuffer obs, ByteBuffer presentOrNull, ByteBuffer outState);
    /** queryNodes == null: every ancestor in index order. outPost: direct buffer of n_query * n_cols * K doubles. */
    static native void marginal(long ws, int nCols, ByteBuffer obs, ByteBuffer presentOrNull, int[] queryNodesOrNull,
                                DoubleBuffer outPostOrNull, double[] outLoglOrNull);
    /** returns the sum over columns; outLogl may be null. */
    static native double logLik(long ws, int nCols, ByteBuffer obs, ByteBuffer presentOrNull, double[] outLoglOrNull);
    /** occupancy != 255 where an extant has a residue; present receives the BEP mask (extants pass through). */
    static native void bepPresence(long tree, int nCols, ByteBuffer occupancy, int device, ByteBuffer present);
    /** IdxTree.getPrunedIndex(pruneMe, true) as a mask transform; in and out may be the same buffer. */
    static native void pruneOrphans(long tree, int nCols, ByteBuffer presentIn, ByteBuffer presentOut);
    static native void destroy(long model, long tree, long ws);
}
