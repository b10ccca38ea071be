package org.broadinstitute.hellbender.utils.pairhmm;

import org.apache.logging.log4j.LogManager;
import org.apache.logging.log4j.Logger;
import org.broadinstitute.gatk.nativebindings.pairhmm.HaplotypeDataHolder;
import org.broadinstitute.gatk.nativebindings.pairhmm.PairHMMNativeArguments;
import org.broadinstitute.gatk.nativebindings.pairhmm.ReadDataHolder;
import org.broadinstitute.hellbender.exceptions.UserException;
import org.broadinstitute.hellbender.utils.genotyper.LikelihoodMatrix;
import org.broadinstitute.hellbender.utils.haplotype.Haplotype;
import org.broadinstitute.hellbender.utils.read.GATKRead;
import org.broadinstitute.hellbender.utils.read.ReadUtils;

import java.util.Arrays;
import java.util.HashMap;
import java.util.List;
import java.util.Map;

/**
 * {@code -pairHMM CUDA_LOGLESS_CACHING}: the PairHMM forward algorithm on NVIDIA B200 GPUs (libgpuphmm).
 *
 * <p>Fulfils the contract of the other native implementation, {@link VectorLoglessPairHMM}: haplotypes are handed
 * over once per region in {@link #initialize(List, Map, int, int)}; each per-sample
 * {@link #computeLog10Likelihoods} call gathers the five per-base arrays of every read, makes ONE native call and
 * copies the read-major result {@code lk[read * nHaplotypes + haplotype]} into the {@link LikelihoodMatrix}; the
 * flat array stays available through {@link #getLogLikelihoodArray()}; {@code --pair-hmm-results-file} is honoured.</p>
 *
 * <p>Likelihoods are computed in fp32 on the GPU; pairs whose fp32 sum leaves the float range are recomputed in fp64
 * on the GPU. {@code --native-pair-hmm-use-double-precision} computes everything in fp64.
 * {@code --native-pair-hmm-threads} sizes the host-side staging pool. The GPUs to use come from the environment
 * variable {@code GATK_CUDA_PAIRHMM_DEVICES} (comma separated ordinals; unset = the current device).</p>
 *
 * <p>There is no CPU fallback. Without a usable GPU or library the constructor throws
 * {@link UserException.HardwareFeatureException}, and the implementation is not part of FASTEST_AVAILABLE.</p>
 */
public final class CudaLoglessPairHMM extends LoglessPairHMM {
    private static final Logger cudaLogger = LogManager.getLogger(CudaLoglessPairHMM.class);
    private static final String DEVICES_ENV = "GATK_CUDA_PAIRHMM_DEVICES";

    private final CudaPairHMMBinding gpu = new CudaPairHMMBinding();

    /** Haplotypes of the current region in the order the native side sees them. */
    private HaplotypeDataHolder[] nativeHaplotypes = new HaplotypeDataHolder[0];
    /** Position of each haplotype (bases + reference flag, i.e. {@link Haplotype#equals}) in {@link #nativeHaplotypes}. */
    private final Map<Haplotype, Integer> nativeIndexOf = new HashMap<>();

    private long nanosInSetup = 0L;

    public CudaLoglessPairHMM(final PairHMMNativeArguments args) throws UserException.HardwareFeatureException {
        // --cuda-pair-hmm-devices (PairHMMNativeArgumentCollection hands over a CudaPairHMMArguments), else the environment
        final int[] fromCommandLine = args instanceof CudaPairHMMArguments ? ((CudaPairHMMArguments) args).devices : null;
        gpu.setDevices(fromCommandLine != null ? fromCommandLine : parseDeviceList(System.getenv(DEVICES_ENV)));
        if (!gpu.load(null)) {
            throw new UserException.HardwareFeatureException(
                    "Machine does not support the CUDA PairHMM: libgpuphmm could not be loaded or no compute-capability 10.x GPU is visible.");
        }
        gpu.initialize(args);
    }

    static int[] parseDeviceList(final String spec) {
        return CudaPairHMMArguments.parseDeviceList(spec);
    }

    /**
     * Stages the region's haplotypes for all the per-sample calls that follow. The Java DP matrices of the parent class
     * are never allocated: {@code readMaxLength} and {@code haplotypeMaxLength} are not needed on this path.
     */
    @Override
    public void initialize(final List<Haplotype> haplotypes, final Map<String, List<GATKRead>> perSampleReadList,
                           final int readMaxLength, final int haplotypeMaxLength) {
        nativeIndexOf.clear();
        nativeHaplotypes = new HaplotypeDataHolder[haplotypes.size()];
        int next = 0;
        for (final Haplotype haplotype : haplotypes) {
            final HaplotypeDataHolder holder = new HaplotypeDataHolder();
            holder.haplotypeBases = haplotype.getBases();
            nativeHaplotypes[next] = holder;
            nativeIndexOf.put(haplotype, next);
            next++;
        }
    }

    @Override
    public void computeLog10Likelihoods(final LikelihoodMatrix<GATKRead, Haplotype> logLikelihoods,
                                        final List<GATKRead> processedReads,
                                        final PairHMMInputScoreImputator inputScoreImputator) {
        if (processedReads.isEmpty()) {
            return;   // nothing to do; getLogLikelihoodArray() keeps its previous value, as in the other implementations
        }
        final long callStart = doProfiling ? System.nanoTime() : 0L;

        final ReadDataHolder[] nativeReads = gatherReads(processedReads, inputScoreImputator);
        final int nHaplotypes = nativeHaplotypes.length;
        final double[] flat = new double[nativeReads.length * nHaplotypes];
        final long setupDone = doProfiling ? System.nanoTime() : 0L;

        gpu.computeLikelihoods(nativeReads, nativeHaplotypes, flat);   // the only native call

        mLogLikelihoodArray = flat;
        scatter(flat, nativeReads, logLikelihoods);

        if (doProfiling) {
            threadLocalPairHMMComputeTimeDiff = System.nanoTime() - callStart;
            pairHMMComputeTime += threadLocalPairHMMComputeTimeDiff;
            nanosInSetup += setupDone - callStart;
        }
    }

    /** One {@link ReadDataHolder} per read: bases, base qualities and the imputed gap-open / gap-continuation penalties. */
    private static ReadDataHolder[] gatherReads(final List<GATKRead> reads, final PairHMMInputScoreImputator imputator) {
        final ReadDataHolder[] holders = new ReadDataHolder[reads.size()];
        int r = 0;
        for (final GATKRead read : reads) {
            final PairHMMInputScoreImputation scores = imputator.impute(read);
            final ReadDataHolder holder = new ReadDataHolder();
            holder.readBases = read.getBases();
            holder.readQuals = read.getBaseQualities();
            holder.insertionGOP = scores.insOpenPenalties();
            holder.deletionGOP = scores.delOpenPenalties();
            holder.overallGCP = scores.gapContinuationPenalties();
            holders[r++] = holder;
        }
        return holders;
    }

    /**
     * A per-sample likelihood computation that has been handed to the native cross-region queue and not collected yet.
     * It owns copies of everything {@link #initialize} will overwrite for the next region, so any number of regions can
     * be in flight; {@link #complete()} blocks until the GPU has the result, then fills the matrix exactly as
     * {@link #computeLog10Likelihoods} would have. Must be completed on the thread that submitted it (the tool thread).
     */
    public final class PendingLikelihoods {
        private final long ticket;
        private final ReadDataHolder[] nativeReads;
        private final int[] column;
        private final int nHaplotypes;
        private final LikelihoodMatrix<GATKRead, Haplotype> matrix;
        private boolean completed;

        private PendingLikelihoods(final long ticket, final ReadDataHolder[] nativeReads, final int[] column,
                                   final int nHaplotypes, final LikelihoodMatrix<GATKRead, Haplotype> matrix) {
            this.ticket = ticket;
            this.nativeReads = nativeReads;
            this.column = column;
            this.nHaplotypes = nHaplotypes;
            this.matrix = matrix;
        }

        /** Waits for the native result and writes it into the matrix; a second call is a no-op. */
        public void complete() {
            if (completed) {
                return;
            }
            completed = true;
            if (nativeReads.length == 0) {
                return;
            }
            final long start = doProfiling ? System.nanoTime() : 0L;
            final double[] flat = new double[nativeReads.length * nHaplotypes];
            gpu.await(ticket, flat);   // throws GATKException if the batch failed on the device
            mLogLikelihoodArray = flat;
            scatter(flat, nativeReads, matrix, column, nHaplotypes);
            if (doProfiling) {
                threadLocalPairHMMComputeTimeDiff = System.nanoTime() - start;
                pairHMMComputeTime += threadLocalPairHMMComputeTimeDiff;
            }
        }
    }

    /**
     * Asynchronous form of {@link #computeLog10Likelihoods}: gathers the reads, submits them against the haplotypes of the
     * last {@link #initialize} call to the native queue (gphmm_submit) and returns at once. Units submitted while the GPU
     * is busy are merged into one batch by the native worker, which is what fills a B200 from 60-read regions.
     */
    public PendingLikelihoods submitLog10Likelihoods(final LikelihoodMatrix<GATKRead, Haplotype> logLikelihoods,
                                                     final List<GATKRead> processedReads,
                                                     final PairHMMInputScoreImputator inputScoreImputator) {
        final int[] column = columnsOf(logLikelihoods);
        if (processedReads.isEmpty()) {
            return new PendingLikelihoods(0L, new ReadDataHolder[0], column, nativeHaplotypes.length, logLikelihoods);
        }
        final long start = doProfiling ? System.nanoTime() : 0L;
        final ReadDataHolder[] nativeReads = gatherReads(processedReads, inputScoreImputator);
        final long ticket = gpu.submit(nativeReads, nativeHaplotypes);   // copies the arrays into the pinned arena
        if (doProfiling) {
            nanosInSetup += System.nanoTime() - start;
        }
        return new PendingLikelihoods(ticket, nativeReads, column, nativeHaplotypes.length, logLikelihoods);
    }

    /** Column of each matrix allele in the native (read-major) result. */
    private int[] columnsOf(final LikelihoodMatrix<GATKRead, Haplotype> matrix) {
        final List<Haplotype> matrixAlleles = matrix.alleles();
        final int[] column = new int[matrixAlleles.size()];
        for (int a = 0; a < column.length; a++) {
            final Integer idx = nativeIndexOf.get(matrixAlleles.get(a));
            if (idx == null) {
                throw new IllegalStateException("haplotype of the likelihood matrix was not passed to initialize()");
            }
            column[a] = idx;
        }
        return column;
    }

    /**
     * Copies the read-major native result into the allele-major matrix. The matrix may list the haplotypes in another
     * order than {@link #initialize} received them, so the column of each matrix allele is looked up once per call.
     */
    private void scatter(final double[] flat, final ReadDataHolder[] nativeReads,
                         final LikelihoodMatrix<GATKRead, Haplotype> matrix) {
        scatter(flat, nativeReads, matrix, columnsOf(matrix), nativeHaplotypes.length);
    }

    private void scatter(final double[] flat, final ReadDataHolder[] nativeReads,
                         final LikelihoodMatrix<GATKRead, Haplotype> matrix, final int[] column, final int nHaplotypes) {
        final List<Haplotype> matrixAlleles = matrix.alleles();
        for (int r = 0; r < nativeReads.length; r++) {
            final ReadDataHolder read = nativeReads[r];
            final int row = r * nHaplotypes;
            for (int a = 0; a < column.length; a++) {
                final double lk = flat[row + column[a]];
                matrix.set(a, r, lk);
                writeToResultsFileIfApplicable(read.readBases, read.readQuals, read.insertionGOP, read.deletionGOP,
                        read.overallGCP, matrixAlleles.get(a).getBases(), lk);
            }
        }
    }

    /** The fields of PairHMMLikelihoodCalculationEngine that parameterise the steps fused around the kernel. */
    public static final class RegionSteps {
        public double pcrRateFactor = 3.0;                       // PCRErrorModel.getRateFactor(); 0 = NONE
        public byte baseQualityScoreThreshold = 18;
        public boolean disableCapReadQualitiesToMapQ = false;
        public double log10GlobalReadMismappingRate = -4.5;
        public boolean symmetricallyNormalizeAllelesToReference = false;
        public boolean filterPoorly = true;
        public double expectedErrorRatePerBase = 0.02;
        public boolean dynamicDisqualification = false;
        public double readDisqualificationScale = 1.0;
    }

    /**
     * One sample of a region with modifyReadQualities, normalizeLikelihoods and the filterPoorlyModeledEvidence decision
     * done on the GPU (see java/patches/PairHMMLikelihoodCalculationEngine.regionSteps.patch for the caller).
     * Standard input-score imputation only (flat GCP, BI/BD or Q45 gap-open penalties); with DRAGstr parameters the
     * caller keeps using {@link #computeLog10Likelihoods}.
     *
     * @param clippedReads the sample's reads after ReadClipper.hardClipSoftClippedBases (or as they are when
     *                     modifySoftclippedBases is set), in matrix order
     * @return indexes (in matrix order) of the reads to remove as poorly modeled; the matrix holds the normalised
     *         likelihoods of all reads
     */
    public int[] computeRegionLikelihoods(final LikelihoodMatrix<GATKRead, Haplotype> logLikelihoods,
                                          final List<GATKRead> clippedReads, final byte constantGCP, final RegionSteps steps,
                                          final String hmmBaseQualitiesTag) {
        final int nReads = clippedReads.size();
        if (nReads == 0) {
            return new int[0];
        }
        final ReadDataHolder[] nativeReads = new ReadDataHolder[nReads];
        final byte[] mapq = new byte[nReads];
        int totalBases = 0;
        for (int r = 0; r < nReads; r++) {
            final GATKRead read = clippedReads.get(r);
            final ReadDataHolder holder = new ReadDataHolder();
            holder.readBases = read.getBases();
            holder.readQuals = read.getBaseQualities();
            holder.insertionGOP = ReadUtils.getBaseInsertionQualities(read);
            holder.deletionGOP = ReadUtils.getBaseDeletionQualities(read);
            holder.overallGCP = new byte[holder.readBases.length];
            Arrays.fill(holder.overallGCP, constantGCP);
            nativeReads[r] = holder;
            mapq[r] = (byte) Math.min(255, read.getMappingQuality());
            totalBases += holder.readBases.length;
        }
        int referenceIndex = -1;
        for (final Map.Entry<Haplotype, Integer> entry : nativeIndexOf.entrySet()) {
            if (entry.getKey().isReference()) {
                referenceIndex = entry.getValue();
            }
        }
        final int flags = (steps.disableCapReadQualitiesToMapQ ? 1 : 0) | (steps.symmetricallyNormalizeAllelesToReference ? 2 : 0)
                | (steps.filterPoorly ? 4 : 0) | (steps.dynamicDisqualification ? 8 : 0);
        final int nHaplotypes = nativeHaplotypes.length;
        final double[] alleleMajor = new double[nReads * nHaplotypes];
        final byte[] keep = new byte[nReads];
        final byte[] hmmBaseQualities = new byte[totalBases];
        gpu.computeRegion(nativeReads, mapq, nativeHaplotypes, new int[]{flags, steps.baseQualityScoreThreshold, referenceIndex},
                new double[]{steps.pcrRateFactor, steps.log10GlobalReadMismappingRate, steps.expectedErrorRatePerBase, steps.readDisqualificationScale},
                alleleMajor, keep, hmmBaseQualities);

        final List<Haplotype> matrixAlleles = logLikelihoods.alleles();
        for (int a = 0; a < matrixAlleles.size(); a++) {
            final int column = nativeIndexOf.get(matrixAlleles.get(a)) * nReads;   // one contiguous row of the native result per allele
            for (int r = 0; r < nReads; r++) {
                logLikelihoods.set(a, r, alleleMajor[column + r]);
            }
        }
        int dropped = 0;
        int offset = 0;
        for (int r = 0; r < nReads; r++) {
            final int length = nativeReads[r].readBases.length;
            // PairHMMLikelihoodCalculationEngine.java:302: the qualities the HMM used, for the DRAGEN filters downstream
            logLikelihoods.evidence().get(r).setTransientAttribute(hmmBaseQualitiesTag, Arrays.copyOfRange(hmmBaseQualities, offset, offset + length));
            offset += length;
            dropped += keep[r] == 0 ? 1 : 0;
        }
        final int[] toRemove = new int[dropped];
        for (int r = 0, k = 0; r < nReads; r++) {
            if (keep[r] == 0) {
                toRemove[k++] = r;
            }
        }
        return toRemove;
    }

    @Override
    public void close() {
        if (doProfiling) {
            final long[] n = gpu.counters();
            final double[] ms = gpu.timers();
            cudaLogger.info("Time spent in setup for JNI call : " + (nanosInSetup * 1e-9));
            cudaLogger.info(String.format(
                    "CUDA PairHMM: %d pairs, %d cells, %d pairs recomputed in fp64, %d kernel launches; "
                            + "GPU fp32 %.3f s, GPU fp64 %.3f s, host staging %.3f s, H2D %d bytes, D2H %d bytes",
                    n[0], n[1], n[2], n[5], ms[0] * 1e-3, ms[1] * 1e-3, ms[3] * 1e-3, n[3], n[4]));
        }
        gpu.done();
        super.close();
    }
}
