package org.broadinstitute.hellbender.utils.pairhmm;

import org.broadinstitute.gatk.nativebindings.pairhmm.PairHMMNativeArguments;

/**
 * {@link PairHMMNativeArguments} (an external class with two fields: maxNumberOfThreads, useDoublePrecision) plus what only
 * the CUDA implementation needs. {@code PairHMMNativeArgumentCollection.getPairHMMArgs()} hands one of these to every
 * {@code PairHMM.Implementation} factory (PairHMM.java:100-108); {@link CudaLoglessPairHMM} looks for the subclass with
 * {@code instanceof}, every other implementation sees the two fields it knows.
 */
public final class CudaPairHMMArguments extends PairHMMNativeArguments {

    /** CUDA ordinals of the GPUs one PairHMM instance spreads its regions over; null = environment variable / current device. */
    public int[] devices = null;

    /** "0,1,2,3" -> {0, 1, 2, 3}; null or blank -> null. */
    public static int[] parseDeviceList(final String spec) {
        if (spec == null || spec.trim().isEmpty()) {
            return null;
        }
        final String[] fields = spec.split(",");
        final int[] ordinals = new int[fields.length];
        for (int k = 0; k < fields.length; k++) {
            ordinals[k] = Integer.parseInt(fields[k].trim());
        }
        return ordinals;
    }
}
