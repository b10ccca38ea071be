package org.broadinstitute.hellbender.utils.pairhmm;

import htsjdk.samtools.util.Locatable;
import org.broadinstitute.gatk.nativebindings.pairhmm.HaplotypeDataHolder;
import org.broadinstitute.gatk.nativebindings.pairhmm.PairHMMNativeArguments;
import org.broadinstitute.gatk.nativebindings.pairhmm.ReadDataHolder;
import org.broadinstitute.gatk.nativebindings.pdhmm.PDHMMNativeArguments;
import org.broadinstitute.hellbender.exceptions.UserException;
import org.broadinstitute.hellbender.tools.walkers.haplotypecaller.PDPairHMMLikelihoodCalculationEngine;
import org.broadinstitute.hellbender.utils.genotyper.LikelihoodMatrix;
import org.broadinstitute.hellbender.utils.haplotype.PartiallyDeterminedHaplotype;
import org.broadinstitute.hellbender.utils.read.GATKRead;

import java.util.List;

/**
 * The partially determined PairHMM of DRAGEN-GATK mode on NVIDIA B200 GPUs (libgpuphmm, gphmm_pd_compute): the CUDA
 * sibling of {@link VectorLoglessPairPDHMM}, to be registered as {@code PDPairHMM.Implementation.CUDA_LOGLESS_CACHING}
 * (see java/patches/PDPairHMM.Implementation.patch).  Same contract: one native call per sample, the read-major
 * result array is kept for {@link #getLogLikelihoodArray()}, reads that do not overlap the determined span of an
 * allele get negative infinity on the Java side (VectorLoglessPairPDHMM.java:129-137).  No CPU fallback.
 */
public final class CudaLoglessPairPDHMM extends LoglessPDPairHMM {
    private final CudaPairHMMBinding gpu = new CudaPairHMMBinding();

    public CudaLoglessPairPDHMM(final PDHMMNativeArguments args) throws UserException.HardwareFeatureException {
        gpu.setDevices(CudaLoglessPairHMM.parseDeviceList(System.getenv("GATK_CUDA_PAIRHMM_DEVICES")));
        if (!gpu.load(null)) {
            throw new UserException.HardwareFeatureException(
                    "Machine does not support the CUDA PDHMM: libgpuphmm could not be loaded or no compute-capability 10.x GPU is visible.");
        }
        final PairHMMNativeArguments plain = new PairHMMNativeArguments();
        plain.maxNumberOfThreads = args == null ? 0 : args.maxNumberOfThreads;
        plain.useDoublePrecision = false;
        gpu.initialize(plain);
    }

    @Override
    public void computeLog10Likelihoods(final LikelihoodMatrix<GATKRead, PartiallyDeterminedHaplotype> logLikelihoods,
                                        final List<GATKRead> processedReads,
                                        final PairHMMInputScoreImputator inputScoreImputator,
                                        final int rangeForReadOverlapToDeterminedBases) {
        if (processedReads.isEmpty()) {
            return;
        }
        final List<PartiallyDeterminedHaplotype> alleles = logLikelihoods.alleles();
        final int readCount = processedReads.size();
        final int alleleCount = alleles.size();

        final ReadDataHolder[] reads = new ReadDataHolder[readCount];
        for (int r = 0; r < readCount; r++) {
            final GATKRead read = processedReads.get(r);
            final PairHMMInputScoreImputation scores = inputScoreImputator.impute(read);
            final ReadDataHolder holder = new ReadDataHolder();
            holder.readBases = read.getBases();
            holder.readQuals = read.getBaseQualities();
            holder.insertionGOP = scores.insOpenPenalties();
            holder.deletionGOP = scores.delOpenPenalties();
            holder.overallGCP = scores.gapContinuationPenalties();
            reads[r] = holder;
        }
        final HaplotypeDataHolder[] haplotypes = new HaplotypeDataHolder[alleleCount];
        for (int a = 0; a < alleleCount; a++) {
            final HaplotypeDataHolder holder = new HaplotypeDataHolder();
            holder.haplotypeBases = alleles.get(a).getBases();
            holder.haplotypePDBases = alleles.get(a).getAlternateBases();
            haplotypes[a] = holder;
        }

        mLogLikelihoodArray = new double[readCount * alleleCount];
        gpu.computePDLikelihoods(reads, haplotypes, mLogLikelihoodArray);   // the only native call

        for (int r = 0; r < readCount; r++) {
            final GATKRead read = processedReads.get(r);
            final Locatable unclippedSpan = (Locatable) read.getTransientAttribute(PDPairHMMLikelihoodCalculationEngine.UNCLIPPED_ORIGINAL_SPAN_ATTR);
            for (int a = 0; a < alleleCount; a++) {
                final int k = r * alleleCount + a;
                final boolean scored = rangeForReadOverlapToDeterminedBases < 0
                        || alleles.get(a).getMaximumExtentOfSiteDeterminedAlleles().overlapsWithMargin(unclippedSpan, rangeForReadOverlapToDeterminedBases + 1);
                if (!scored) {
                    mLogLikelihoodArray[k] = Double.NEGATIVE_INFINITY;
                }
                logLikelihoods.set(a, r, mLogLikelihoodArray[k]);
                writeToResultsFileIfApplicable(reads[r].readBases, reads[r].readQuals, reads[r].insertionGOP, reads[r].deletionGOP,
                        reads[r].overallGCP, haplotypes[a].haplotypeBases, haplotypes[a].haplotypePDBases, mLogLikelihoodArray[k]);
            }
        }
    }

    @Override
    public void close() {
        gpu.done();
        super.close();
    }
}
