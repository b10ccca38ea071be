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

import java.util.LinkedHashMap;
import java.util.List;
import java.util.Map;

/**
 * {@code -pairHMM CUDA_LOGLESS_CACHING}: the PairHMM forward algorithm on NVIDIA B200 GPUs.
 *
 * Drop-in sibling of {@link VectorLoglessPairHMM} (same parent, same overrides, same result layout): haplotypes are
 * staged once per region in {@link #initialize}, every per-sample call packs the reads into {@link ReadDataHolder}s,
 * makes ONE native call, and scatters the read-major {@code double[]} into the {@link LikelihoodMatrix}.
 * Likelihoods are computed in fp32 on the GPU with a GPU fp64 redo of under-flowing pairs
 * ({@code --native-pair-hmm-use-double-precision} forces fp64 for every pair).  There is no CPU fallback: when no
 * usable GPU is present the constructor throws {@link UserException.HardwareFeatureException}, exactly like the AVX
 * implementations do when AVX is missing, and the entry is NOT part of FASTEST_AVAILABLE.
 */
public final class CudaLoglessPairHMM extends LoglessPairHMM {
    private static final Logger logger = LogManager.getLogger(CudaLoglessPairHMM.class);

    private long threadLocalSetupTimeDiff = 0;
    private long pairHMMSetupTime = 0;

    private final CudaPairHMMBinding pairHmm;

    // Haplotype -> index in the list passed to initialize(); keyed by Haplotype.equals (bases + isReference)
    private final Map<Haplotype, Integer> haplotypeToHaplotypeListIdxMap = new LinkedHashMap<>();
    private HaplotypeDataHolder[] mHaplotypeDataArray;

    public CudaLoglessPairHMM(final PairHMMNativeArguments args) throws UserException.HardwareFeatureException {
        pairHmm = new CudaPairHMMBinding();
        final String deviceList = System.getenv("GATK_CUDA_PAIRHMM_DEVICES");   // e.g. "0,1,2,3"; unset = current device
        if (deviceList != null && !deviceList.trim().isEmpty()) {
            final String[] tok = deviceList.split(",");
            final int[] devices = new int[tok.length];
            for (int i = 0; i < tok.length; i++) {
                devices[i] = Integer.parseInt(tok[i].trim());
            }
            pairHmm.setDevices(devices);
        }
        if (!pairHmm.load(null)) {
            throw new UserException.HardwareFeatureException("Machine does not support the CUDA PairHMM (no compute-capability 10.x GPU or libgpuphmm could not be loaded).");
        }
        pairHmm.initialize(args);
    }

    /**
     * {@inheritDoc}
     */
    @Override
    public void initialize(final List<Haplotype> haplotypes, final Map<String, List<GATKRead>> perSampleReadList,
                           final int readMaxLength, final int haplotypeMaxLength) {
        // like VectorLoglessPairHMM: the Java matrices of the parent are never allocated
        final int numHaplotypes = haplotypes.size();
        mHaplotypeDataArray = new HaplotypeDataHolder[numHaplotypes];
        int idx = 0;
        haplotypeToHaplotypeListIdxMap.clear();
        for (final Haplotype currHaplotype : haplotypes) {
            mHaplotypeDataArray[idx] = new HaplotypeDataHolder();
            mHaplotypeDataArray[idx].haplotypeBases = currHaplotype.getBases();
            haplotypeToHaplotypeListIdxMap.put(currHaplotype, idx);
            ++idx;
        }
    }

    /**
     * {@inheritDoc}
     */
    @Override
    public void computeLog10Likelihoods(final LikelihoodMatrix<GATKRead, Haplotype> logLikelihoods,
                                        final List<GATKRead> processedReads,
                                        final PairHMMInputScoreImputator inputScoreImputator) {
        if (processedReads.isEmpty()) {
            return;
        }
        if (doProfiling) {
            startTime = System.nanoTime();
        }
        final int readListSize = processedReads.size();
        final int numHaplotypes = logLikelihoods.numberOfAlleles();
        final ReadDataHolder[] readDataArray = new ReadDataHolder[readListSize];
        int idx = 0;
        for (final GATKRead read : processedReads) {
            final PairHMMInputScoreImputation inputScoreImputation = inputScoreImputator.impute(read);
            readDataArray[idx] = new ReadDataHolder();
            readDataArray[idx].readBases = read.getBases();
            readDataArray[idx].readQuals = read.getBaseQualities();
            readDataArray[idx].insertionGOP = inputScoreImputation.insOpenPenalties();
            readDataArray[idx].deletionGOP = inputScoreImputation.delOpenPenalties();
            readDataArray[idx].overallGCP = inputScoreImputation.gapContinuationPenalties();
            ++idx;
        }

        mLogLikelihoodArray = new double[readListSize * numHaplotypes];
        if (doProfiling) {
            threadLocalSetupTimeDiff = (System.nanoTime() - startTime);
        }

        pairHmm.computeLikelihoods(readDataArray, mHaplotypeDataArray, mLogLikelihoodArray);

        int readIdx = 0;
        for (int r = 0; r < readListSize; r++) {
            int hapIdx = 0;
            for (final Haplotype haplotype : logLikelihoods.alleles()) {
                // the matrix's allele order may differ from the order given to initialize()
                final int idxInsideHaplotypeList = haplotypeToHaplotypeListIdxMap.get(haplotype);
                final double lk = mLogLikelihoodArray[readIdx + idxInsideHaplotypeList];
                logLikelihoods.set(hapIdx, r, lk);
                writeToResultsFileIfApplicable(readDataArray[r].readBases, readDataArray[r].readQuals, readDataArray[r].insertionGOP,
                        readDataArray[r].deletionGOP, readDataArray[r].overallGCP, haplotype.getBases(), lk);
                ++hapIdx;
            }
            readIdx += numHaplotypes;
        }
        if (doProfiling) {
            threadLocalPairHMMComputeTimeDiff = (System.nanoTime() - startTime);
            pairHMMComputeTime += threadLocalPairHMMComputeTimeDiff;
            pairHMMSetupTime += threadLocalSetupTimeDiff;
        }
    }

    @Override
    public void close() {
        if (doProfiling) {
            final long[] c = pairHmm.counters();
            final double[] t = pairHmm.timers();
            logger.info("Time spent in setup for JNI call : " + (pairHMMSetupTime * 1e-9));
            logger.info(String.format("CUDA PairHMM: %d pairs, %d cells, %d pairs redone in fp64, %d kernel launches; fp32 kernels %.3f s, fp64 kernels %.3f s, H2D %d bytes, D2H %d bytes",
                    c[0], c[1], c[2], c[5], t[0] * 1e-3, t[1] * 1e-3, c[3], c[4]));
        }
        pairHmm.done();
        super.close();
    }
}
