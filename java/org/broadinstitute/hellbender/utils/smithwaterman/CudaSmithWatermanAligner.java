package org.broadinstitute.hellbender.utils.smithwaterman;

import htsjdk.samtools.Cigar;
import htsjdk.samtools.CigarElement;
import htsjdk.samtools.CigarOperator;
import org.broadinstitute.gatk.nativebindings.pairhmm.PairHMMNativeArguments;
import org.broadinstitute.gatk.nativebindings.smithwaterman.SWOverhangStrategy;
import org.broadinstitute.gatk.nativebindings.smithwaterman.SWParameters;
import org.broadinstitute.hellbender.exceptions.UserException;
import org.broadinstitute.hellbender.utils.Utils;
import org.broadinstitute.hellbender.utils.pairhmm.CudaPairHMMBinding;

import java.util.ArrayList;
import java.util.List;

/**
 * Smith-Waterman on NVIDIA B200 GPUs (libgpuphmm, gphmm_sw_align): bit-identical with {@link SmithWatermanJavaAligner}
 * (offset and CIGAR), to be registered as {@code SmithWatermanAligner.Implementation.CUDA}
 * (java/patches/SmithWatermanAligner.Implementation.patch).
 *
 * <p>One alignment is far too little work for a GPU call: {@link #align} is there for interface completeness, the useful
 * entry point is {@link #alignBatch}, which the per-region loops (AssemblyBasedCallerUtils.realignReadsToTheirBestHaplotype,
 * :107-135, one alignment per read) can feed with all the reads of a region at once.</p>
 */
public final class CudaSmithWatermanAligner implements SmithWatermanAligner {
    private static final CigarOperator[] OPS = {CigarOperator.M, CigarOperator.I, CigarOperator.D, CigarOperator.S};
    private final CudaPairHMMBinding gpu = new CudaPairHMMBinding();

    public CudaSmithWatermanAligner() throws UserException.HardwareFeatureException {
        if (!gpu.load(null)) {
            throw new UserException.HardwareFeatureException(
                    "Machine does not support the CUDA Smith-Waterman: libgpuphmm could not be loaded or no compute-capability 10.x GPU is visible.");
        }
        gpu.initialize(new PairHMMNativeArguments());
    }

    @Override
    public SmithWatermanAlignment align(final byte[] reference, final byte[] alternate, final SWParameters parameters,
                                        final SWOverhangStrategy overhangStrategy) {
        final List<byte[]> refs = new ArrayList<>(1);
        final List<byte[]> alts = new ArrayList<>(1);
        refs.add(reference);
        alts.add(alternate);
        return alignBatch(refs, alts, parameters, overhangStrategy).get(0);
    }

    /** Pair k aligns {@code alternates.get(k)} to {@code references.get(k)}; all pairs share parameters and strategy. */
    public List<SmithWatermanAlignment> alignBatch(final List<byte[]> references, final List<byte[]> alternates,
                                                   final SWParameters parameters, final SWOverhangStrategy overhangStrategy) {
        Utils.nonNull(parameters);
        Utils.nonNull(overhangStrategy);
        Utils.validateArg(references.size() == alternates.size(), "one alternate per reference");
        final int n = references.size();
        final byte[][] refs = references.toArray(new byte[n][]);
        final byte[][] alts = alternates.toArray(new byte[n][]);
        final int[] params = {parameters.getMatchValue(), parameters.getMismatchPenalty(), parameters.getGapOpenPenalty(),
                parameters.getGapExtendPenalty(), strategyCode(overhangStrategy)};
        int capacity = 32;
        int[] offsets;
        int[] nElems;
        int[] elems;
        while (true) {
            offsets = new int[n];
            nElems = new int[n];
            elems = new int[n * capacity];
            if (gpu.smithWatermanBatch(refs, alts, params, capacity, offsets, nElems, elems)) {
                break;
            }
            capacity *= 8;   // some CIGAR did not fit: rare (tens of indels in one alignment), redo with more room
        }
        final List<SmithWatermanAlignment> result = new ArrayList<>(n);
        for (int k = 0; k < n; k++) {
            final List<CigarElement> cigar = new ArrayList<>(nElems[k]);
            for (int e = 0; e < nElems[k]; e++) {
                final int packed = elems[k * capacity + e];
                cigar.add(new CigarElement(packed >>> 4, OPS[packed & 15]));
            }
            final int offset = offsets[k];
            final Cigar finished = new Cigar(cigar);
            result.add(new SmithWatermanAlignment() {
                @Override
                public Cigar getCigar() {
                    return finished;
                }

                @Override
                public int getAlignmentOffset() {
                    return offset;
                }
            });
        }
        return result;
    }

    private static int strategyCode(final SWOverhangStrategy strategy) {
        switch (strategy) {
            case SOFTCLIP: return 0;
            case INDEL: return 1;
            case LEADING_INDEL: return 2;
            case IGNORE: return 3;
            default: throw new IllegalArgumentException("Unknown overhang strategy " + strategy);
        }
    }

    @Override
    public void close() {
        gpu.done();
    }
}
