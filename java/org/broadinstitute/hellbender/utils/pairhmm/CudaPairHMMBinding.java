package org.broadinstitute.hellbender.utils.pairhmm;

import org.broadinstitute.gatk.nativebindings.pairhmm.HaplotypeDataHolder;
import org.broadinstitute.gatk.nativebindings.pairhmm.PairHMMNativeArguments;
import org.broadinstitute.gatk.nativebindings.pairhmm.PairHMMNativeBinding;
import org.broadinstitute.gatk.nativebindings.pairhmm.ReadDataHolder;
import org.broadinstitute.hellbender.exceptions.GATKException;
import org.broadinstitute.hellbender.utils.NativeUtils;

import java.io.File;

/**
 * {@link PairHMMNativeBinding} backed by libgpuphmm (CUDA, sm_100a) through the JNI shim
 * {@code libgpuphmm_jni.so} (source: gatk_b200/csrc/gpuphmm_jni.cpp, C ABI: include/gpuphmm.h).
 *
 * Same contract as the GKL bindings that {@link VectorLoglessPairHMM} drives
 * (load / initialize / computeLikelihoods / done, see VectorLoglessPairHMM.java:63,81,138,164):
 * <ul>
 *   <li>{@link #load(File)} returns {@code false} (never throws) when the library cannot be loaded or no
 *       compute-capability-10.x device is visible; the caller turns that into a
 *       {@code UserException.HardwareFeatureException}.  There is no CPU fallback.</li>
 *   <li>{@link #computeLikelihoods} fills {@code likelihoodArray[r * nHaps + h]} with log10 likelihoods.</li>
 *   <li>CUDA failures surface as {@link GATKException}; out-of-range qualities as
 *       {@link IllegalArgumentException} (as PairHMMModel.qualToTransProbs does).</li>
 * </ul>
 * Besides the synchronous call the binding exposes the library's asynchronous queue
 * ({@link #submit}/{@link #await}) so that a pipelined caller can keep several regions in flight.
 */
public final class CudaPairHMMBinding implements PairHMMNativeBinding {
    /**
     * Path inside the GATK jar of the single shared object that holds the JNI shim AND the CUDA library
     * (gpuphmm_jni.cpp + gpuphmm.cu linked together, CUDA runtime linked statically).  One file, because
     * {@link NativeUtils#loadLibraryFromClasspath} extracts to a random temp name and a second library could not be
     * resolved by soname.
     */
    private static final String JNI_LIBRARY = "/native/libgpuphmm_jni.so";
    private static boolean loaded = false;

    private long handle = 0L;

    /** Device ordinals to use; {@code null} or empty = the current CUDA device. Set before {@link #initialize}. */
    private int[] devices = null;

    public void setDevices(final int[] devices) {
        this.devices = devices == null ? null : devices.clone();
    }

    @Override
    public synchronized boolean load(final File tmpDir) {
        if (!loaded) {
            loaded = NativeUtils.loadLibraryFromClasspath(JNI_LIBRARY);
        }
        return loaded && nativeDeviceCount() > 0;
    }

    @Override
    public void initialize(final PairHMMNativeArguments args) {
        if (handle != 0L) {
            done();
        }
        final boolean fp64 = args != null && args.useDoublePrecision;
        final int threads = args == null ? 0 : args.maxNumberOfThreads;
        handle = nativeCreate(devices, fp64, threads);   // throws GATKException on failure
    }

    @Override
    public void computeLikelihoods(final ReadDataHolder[] readDataArray, final HaplotypeDataHolder[] haplotypeDataArray,
                                   final double[] likelihoodArray) {
        if (handle == 0L) {
            throw new IllegalStateException("CudaPairHMMBinding.initialize() has not been called");
        }
        nativeCompute(handle, readDataArray, haplotypeDataArray, likelihoodArray);
    }

    /** Queue one (reads x haplotypes) unit; inputs are copied before the call returns. */
    public long submit(final ReadDataHolder[] readDataArray, final HaplotypeDataHolder[] haplotypeDataArray) {
        if (handle == 0L) {
            throw new IllegalStateException("CudaPairHMMBinding.initialize() has not been called");
        }
        return nativeSubmit(handle, readDataArray, haplotypeDataArray);
    }

    /** Block until the ticket's results are ready and copy them into {@code likelihoodArray}. */
    public void await(final long ticket, final double[] likelihoodArray) {
        nativeAwait(handle, ticket, likelihoodArray);
    }

    /**
     * One (region, sample) unit with the steps either side of the kernel fused in on the GPU (gphmm_compute_regions):
     * PairHMMLikelihoodCalculationEngine.modifyReadQualities before, AlleleLikelihoods.normalizeLikelihoods and the
     * keep/drop decision of filterPoorlyModeledEvidence after.
     *
     * @param readDataArray   reads as modifyReadQualities receives them (soft clips already removed; raw base qualities;
     *                        BI/BD or flat Q45 in insertionGOP/deletionGOP; flat GCP)
     * @param mappingQualities one MAPQ per read
     * @param intParams       {flags (GPHMM_RS_*), baseQualityScoreThreshold, index of the reference haplotype or -1}
     * @param doubleParams    {PCR rate factor (0 = NONE), log10GlobalReadMismappingRate, expectedErrorRatePerBase, readDisqualificationScale}
     * @param likelihoods     out, allele-major: {@code likelihoods[h * nReads + r]}, normalised
     * @param keep            out, per read: 0 = the read is removed as poorly modeled
     * @param hmmBaseQualities out, nullable: modified base qualities of all reads back to back (HMM_BASE_QUALITIES_TAG)
     */
    public void computeRegion(final ReadDataHolder[] readDataArray, final byte[] mappingQualities,
                              final HaplotypeDataHolder[] haplotypeDataArray, final int[] intParams, final double[] doubleParams,
                              final double[] likelihoods, final byte[] keep, final byte[] hmmBaseQualities) {
        if (handle == 0L) {
            throw new IllegalStateException("CudaPairHMMBinding.initialize() has not been called");
        }
        nativeComputeRegion(handle, readDataArray, mappingQualities, haplotypeDataArray, intParams, doubleParams, likelihoods, keep, hmmBaseQualities);
    }

    /**
     * PD-HMM (DRAGEN-GATK partially determined haplotypes): the same call as {@link #computeLikelihoods} with
     * {@code HaplotypeDataHolder.haplotypePDBases} filled in, as {@code PairPDHMMNativeBinding.computeLikelihoods} is
     * called at VectorLoglessPairPDHMM.java:115.
     */
    public void computePDLikelihoods(final ReadDataHolder[] readDataArray, final HaplotypeDataHolder[] haplotypeDataArray,
                                     final double[] likelihoodArray) {
        if (handle == 0L) {
            throw new IllegalStateException("CudaPairHMMBinding.initialize() has not been called");
        }
        nativeComputePD(handle, readDataArray, haplotypeDataArray, likelihoodArray);
    }

    /**
     * Batched Smith-Waterman (gphmm_sw_align): pair k aligns {@code alternates[k]} to {@code references[k]}.
     * {@code params = {match, mismatch, gapOpen, gapExtend, strategy}} with strategy 0 SOFTCLIP, 1 INDEL, 2 LEADING_INDEL,
     * 3 IGNORE; {@code elems[k * capacity + e] = (length << 4) | op} with op 0 M, 1 I, 2 D, 3 S.
     *
     * @return false when some CIGAR needs more than {@code capacity} elements ({@code nElems[k] == -1} for those pairs)
     */
    public boolean smithWatermanBatch(final byte[][] references, final byte[][] alternates, final int[] params, final int capacity,
                                      final int[] offsets, final int[] nElems, final int[] elems) {
        if (handle == 0L) {
            throw new IllegalStateException("CudaPairHMMBinding.initialize() has not been called");
        }
        return nativeSwAlign(handle, references, alternates, params, capacity, offsets, nElems, elems);
    }

    @Override
    public void done() {
        if (handle != 0L) {
            nativeDestroy(handle);
            handle = 0L;
        }
    }

    /** kernel / transfer counters accumulated by the library: {pairs, cells, rescuedPairs, h2dBytes, d2hBytes, launches} */
    public long[] counters() {
        return handle == 0L ? new long[6] : nativeCounters(handle);
    }

    /** {fp32 kernel ms, fp64 kernel ms, device ms, host staging ms, wall ms} */
    public double[] timers() {
        return handle == 0L ? new double[5] : nativeTimers(handle);
    }

    private static native int nativeDeviceCount();
    private static native long nativeCreate(int[] devices, boolean forceFp64, int hostThreads);
    private static native void nativeCompute(long handle, ReadDataHolder[] reads, HaplotypeDataHolder[] haps, double[] out);
    private static native void nativeComputeRegion(long handle, ReadDataHolder[] reads, byte[] mapq, HaplotypeDataHolder[] haps,
                                                   int[] intParams, double[] doubleParams, double[] out, byte[] keep, byte[] hmmBaseQuals);
    private static native void nativeComputePD(long handle, ReadDataHolder[] reads, HaplotypeDataHolder[] haps, double[] out);
    private static native boolean nativeSwAlign(long handle, byte[][] refs, byte[][] alts, int[] params, int capacity,
                                                int[] offsets, int[] nElems, int[] elems);
    private static native long nativeSubmit(long handle, ReadDataHolder[] reads, HaplotypeDataHolder[] haps);
    private static native void nativeAwait(long handle, long ticket, double[] out);
    private static native void nativeDestroy(long handle);
    private static native long[] nativeCounters(long handle);
    private static native double[] nativeTimers(long handle);
}
