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
    /** joint + marginal of the same alignment in one call: the joint pass hides under the posteriors' way back over PCIe. */
    static native void jointMarginal(long ws, int nCols, ByteBuffer obs, ByteBuffer presentOrNull, ByteBuffer outState,
                                     int[] queryNodesOrNull, DoubleBuffer outPostOrNull, double[] outLoglOrNull);
    /** returns the sum over columns; outLogl may be null. */
    static native double logLik(long ws, int nCols, ByteBuffer obs, ByteBuffer presentOrNull, double[] outLoglOrNull);
    /** occupancy != 255 where an extant has a residue; present receives the BEP mask (extants pass through). */
    static native void bepPresence(long tree, int nCols, ByteBuffer occupancy, int device, ByteBuffer present);
    /** IdxTree.getPrunedIndex(pruneMe, true) as a mask transform; in and out may be the same buffer. */
    static native void pruneOrphans(long tree, int nCols, ByteBuffer presentIn, ByteBuffer presentOut);
    static native void destroy(long model, long tree, long ws);

    // ---- Prediction.PredictByBidirEdgeParsimony incl. the ancestors' POGs (src/asr/Prediction.java:1443-1590) ----
    /** as bepPresence, and returns a handle on every ancestor's edge list (a bnk_pogs*). */
    static native long bepInfer(long tree, int nCols, ByteBuffer occupancy, int device, ByteBuffer present);
    static native int pogEdgeCount(long pogs, int branchPointIndex);
    /** edges sorted by (from, to); from = -1 start, to = nCols end; flags bit 0 / 1 = BidirEdge forward / backward. */
    static native void pogEdges(long pogs, int branchPointIndex, int[] from, int[] to, ByteBuffer flags);
    static native void pogsDestroy(long pogs);

    // ---- every device of the node behind one handle: the GRASP.NTHREADS knob over GPUs (ThreadedDecorators.java:28-72) ----
    /** nDevices = 0: all visible devices. Columns are split into contiguous blocks, one per device. Returns a bnk_mgpu*. */
    static native long mgpuCreate(long model, long tree, int maxCols, int nDevices);
    static native void mgpuSetRates(long mgpu, int nCols, double[] ratesOrNull);
    static native void mgpuJoint(long mgpu, int nNodes, int nCols, ByteBuffer obs, ByteBuffer presentOrNull, ByteBuffer outState);
    static native void mgpuMarginal(long mgpu, int nNodes, int nStates, int nCols, ByteBuffer obs, ByteBuffer presentOrNull,
                                    int[] queryNodes, DoubleBuffer outPost, double[] outLoglOrNull);
    /** sum over all columns (per-device sums all-reduced); outLogl may be null. */
    static native double mgpuLogLik(long mgpu, int nNodes, int nCols, ByteBuffer obs, ByteBuffer presentOrNull, double[] outLoglOrNull);
    static native void mgpuDestroy(long mgpu);

    // ---- asr.IndelPeeler under bn.ctmc.GapSubstModel (src/asr/IndelPeeler.java) ----
    /** the base model optimiseMuLambda builds for a name (:474-503); IllegalArgumentException for unknown names. */
    static native long indelBaseModel(String substModelName);
    /** aln: direct ByteBuffer n_nodes * nCols, 255 = gap, extants' rows only. Returns a bnk_indel*. */
    static native long indelCreate(long baseModel, long tree, int nCols, ByteBuffer aln, int device);
    /** the same handle over nDevices GPUs (0 = all visible): column blocks swept concurrently, sums in column order. */
    static native long indelCreateMulti(long baseModel, long tree, int nCols, ByteBuffer aln, int nDevices);
    static native void indelSetRates(long indel, double mu, double lambda);
    /** computeColumnPriors: out[col * rates.length + rate]. */
    static native void indelColumnPriors(long indel, double geometricSeqLenParam, int nCols, double[] rates, double[] out);
    /** calcProbAlnGivenTree at the rates last set. */
    static native double indelAlnLogl(long indel, double geometricSeqLenParam);
    /** optimiseMuLambda: Brent over mu = lambda in [minVal, maxVal]. */
    static native double indelOptimiseMuLambda(long indel, double minVal, double maxVal, double geometricSeqLenParam);
    static native void indelDestroy(long indel);
}
