/*
 * region_steps_oracle.c -- CPU restatement of the steps either side of the PairHMM kernel in
 * PairHMMLikelihoodCalculationEngine.computeReadLikelihoods, used ONLY as a test oracle.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as pairhmm_oracle.c).
 *
 * Parity status: integer steps PINNED to the reference's known-answer tests
 *   - GATKVariantContextUtilsUnitTest.java:920-955  (findNumberOfRepetitions, 28 vectors)
 *   - PairHMMLikelihoodCalculationEngineUnitTest.java:101-136 (PCR error model on pure repeats; the expected repeat
 *     lengths are restated by hand in tests/test_region_steps.py)
 * normalisation / filtering follow AlleleLikelihoodsUnitTest.java:192-217,357-387 (property tests, restated).
 *
 * Path prefixes: HC/ = src/main/java/org/broadinstitute/hellbender/tools/walkers/haplotypecaller/
 *                U/  = src/main/java/org/broadinstitute/hellbender/utils/
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#define MAX_STR_UNIT_LENGTH 8 /* HC/ReadLikelihoodCalculationEngine.java:25 */
#define MAX_REPEAT_LENGTH 20  /* HC/ReadLikelihoodCalculationEngine.java:26 */
#define MIN_USABLE_Q_SCORE 6  /* U/QualityUtils.java:28 */

static int equal_range(const uint8_t *a, int ao, const uint8_t *b, int bo, int n) { return memcmp(a + ao, b + bo, (size_t)n) == 0; }

/* U/variant/GATKVariantContextUtils.java:976-1012 */
int region_oracle_find_repetitions(const uint8_t *unit, int unit_off, int unit_len, const uint8_t *test, int test_off,
                                   int test_len, int leading)
{
    if (test_len == 0) return 0;
    const int diff = test_len - unit_len;
    int n = 0;
    if (leading) {
        for (int start = 0; start <= diff; start += unit_len) {
            if (!equal_range(test, start + test_off, unit, unit_off, unit_len)) return n;
            ++n;
        }
    } else {
        for (int start = diff; start >= 0; start -= unit_len) {
            if (!equal_range(test, start + test_off, unit, unit_off, unit_len)) return n;
            ++n;
        }
    }
    return n;
}

/* HC/ReadLikelihoodCalculationEngine.java:193-253; returns the repeat length (the Pair's right member) */
int region_oracle_tandem_repeat_length(const uint8_t *bases, int n, int offset)
{
    int maxBW = 0;
    int bw_off = offset, bw_len = 1; /* bestBWRepeatUnit = {readBases[offset]} */
    for (int str = 1; str <= MAX_STR_UNIT_LENGTH; ++str) {
        if (offset + 1 - str < 0) break;
        maxBW = region_oracle_find_repetitions(bases, offset - str + 1, str, bases, 0, offset + 1, 0);
        if (maxBW > 1) { bw_off = offset - str + 1; bw_len = str; break; }
    }
    int maxRL = maxBW;
    if (offset < n - 1) {
        int fw_off = offset + 1, fw_len = 1, maxFW = 0;
        for (int str = 1; str <= MAX_STR_UNIT_LENGTH; ++str) {
            if (offset + str + 1 > n) break;
            maxFW = region_oracle_find_repetitions(bases, offset + 1, str, bases, offset + 1, n - offset - 1, 1);
            if (maxFW > 1) { fw_off = offset + 1; fw_len = str; break; }
        }
        if (fw_len == bw_len && equal_range(bases, fw_off, bases, bw_off, fw_len)) {
            maxRL = maxBW + maxFW;
        } else {
            /* the 3-argument overload on testString = readBases[0, offset] (:949-957) */
            maxBW = region_oracle_find_repetitions(bases, fw_off, fw_len, bases, 0, offset + 1, 0);
            maxRL = maxFW + maxBW;
        }
    }
    return maxRL > MAX_REPEAT_LENGTH ? MAX_REPEAT_LENGTH : maxRL;
}

/* HC/PairHMMLikelihoodCalculationEngine.java:32,35,343-358; U/MathUtils.java:428-430 */
void region_oracle_pcr_cache(double rate_factor, uint8_t out[MAX_REPEAT_LENGTH + 1])
{
    for (int i = 0; i <= MAX_REPEAT_LENGTH; ++i) {
        const double d = 40.0 - exp(i / (rate_factor * M_PI)) + 1.0;
        const int r = d > 0.0 ? (int)(d + 0.5) : (int)(d - 0.5);
        out[i] = (uint8_t)(int8_t)(r > 10 ? r : 10);
    }
}

static uint8_t fixed_if_too_low(uint8_t v, int min_qual, uint8_t fixed)
{
    return (int8_t)v < (int8_t)min_qual ? fixed : v; /* Java bytes are signed (:311-313) */
}

/* HC/PairHMMLikelihoodCalculationEngine.java:283-313 + :361-371, in place on one read.
 * rate_factor 0 = PCRErrorModel.NONE.  The soft-clip handling of :287 needs the CIGAR and stays with the caller. */
void region_oracle_modify_read(const uint8_t *bases, uint8_t *base_q, uint8_t *ins_q, uint8_t *del_q, int n, int mapq,
                               double rate_factor, int bq_threshold, int disable_cap_to_mapq)
{
    if (rate_factor != 0.0) {
        uint8_t cache[MAX_REPEAT_LENGTH + 1];
        region_oracle_pcr_cache(rate_factor, cache);
        for (int i = 1; i < n; ++i) {
            const int rl = region_oracle_tandem_repeat_length(bases, n, i - 1);
            if (cache[rl] < ins_q[i - 1]) ins_q[i - 1] = cache[rl];
            if (cache[rl] < del_q[i - 1]) del_q[i - 1] = cache[rl];
        }
    }
    for (int i = 0; i < n; ++i) {
        if (!disable_cap_to_mapq) base_q[i] = (uint8_t)(base_q[i] < mapq ? base_q[i] : mapq);
        base_q[i] = fixed_if_too_low(base_q[i], bq_threshold, MIN_USABLE_Q_SCORE);
        ins_q[i] = fixed_if_too_low(ins_q[i], MIN_USABLE_Q_SCORE, MIN_USABLE_Q_SCORE);
        del_q[i] = fixed_if_too_low(del_q[i], MIN_USABLE_Q_SCORE, MIN_USABLE_Q_SCORE);
    }
}

/* U/genotyper/AlleleLikelihoods.java:416-458 with searchBestAllele :505-533 (no priorities).
 * lk is read-major [r * n_haps + h] (the PairHMM's array); out is allele-major [h * n_reads + r]
 * (valuesBySampleIndex[s][a][r], :71-74).  ref_hap < 0: no reference allele. */
void region_oracle_normalize(const double *lk, int n_reads, int n_haps, int ref_hap, double max_diff_cap,
                             int symmetric, double *out)
{
    for (int r = 0; r < n_reads; ++r) {
        const double *row = lk + (size_t)r * n_haps;
        for (int h = 0; h < n_haps; ++h) out[(size_t)h * n_reads + r] = row[h];
        if (max_diff_cap == -INFINITY || n_haps <= 1) continue;
        const int can_be_ref = symmetric;
        int best = (can_be_ref || ref_hap != 0) ? 0 : 1;
        double best_lk = row[best];
        for (int a = best + 1; a < n_haps; ++a) {
            if (!can_be_ref && ref_hap == a) continue;
            if (row[a] > best_lk) { best = a; best_lk = row[a]; }
        }
        const double cap = best_lk + max_diff_cap;
        for (int h = 0; h < n_haps; ++h)
            if (row[h] < cap) out[(size_t)h * n_reads + r] = cap;
    }
}

/* HC/ReadLikelihoodCalculationEngine.java:160-191: baseQ, mean, variance */
static const double dyn_table[] = {
    1,  5.996842844, 0.196616587, 2,  5.870018422, 1.388545569, 3,  5.401558531, 5.641990128,
    4,  4.818940919, 10.33176216, 5,  4.218758304, 14.25799688, 6,  3.646319832, 17.02880749,
    7,  3.122346753, 18.64537883, 8,  2.654731979, 19.27521677, 9,  2.244479156, 19.13584613,
    10, 1.88893867,  18.43922003, 11, 1.583645342, 17.36842261, 12, 1.3233807,   16.07088712,
    13, 1.102785365, 14.65952563, 14, 0.916703025, 13.21718577, 15, 0.760361881, 11.80207947,
    16, 0.629457387, 10.45304833, 17, 0.520175654, 9.194183767, 18, 0.42918208,  8.038657241,
    19, 0.353590663, 6.991779595, 20, 0.290923699, 6.053379213, 21, 0.23906788,  5.219610436,
    22, 0.196230431, 4.484302033, 23, 0.160897421, 3.839943445, 24, 0.131795374, 3.27839108,
    25, 0.1078567,   2.791361596, 26, 0.088189063, 2.370765375, 27, 0.072048567, 2.008921719,
    28, 0.058816518, 1.698687797, 29, 0.047979438, 1.433525748, 30, 0.039111985, 1.207526336,
    31, 0.031862437, 1.015402928, 32, 0.025940415, 0.852465956, 33, 0.021106532, 0.714585285,
    34, 0.017163711, 0.598145851, 35, 0.013949904, 0.500000349, 36, 0.011332027, 0.41742159,
    37, 0.009200898, 0.348056286, 38, 0.007467036, 0.289881373, 39, 0.006057179, 0.241163527,
    40, 0.004911394, 0.200422214};

const double *region_oracle_dynamic_table(void) { return dyn_table; }

/* HC/ReadLikelihoodCalculationEngine.java:95-113 (static cap) and :118-151 (DRAGEN dynamic threshold) */
double region_oracle_min_true_likelihood(const uint8_t *hmm_base_q, int n, double max_error_per_base, int dynamic,
                                         double dynamic_scale)
{
    const double errs = ceil(n * max_error_per_base);
    if (!dynamic) return (errs < 2.0 ? errs : 2.0) * -4.0;
    const double static_thr = errs * -4.0; /* capLikelihoods = false on the dynamic route (:100) */
    double mean = 0, var = 0;
    for (int i = 0; i < n; ++i) {
        const int bq = hmm_base_q[i];
        const int entry = bq <= 1 ? 0 : (bq < 40 ? bq : 40) - 1;
        mean += dyn_table[entry * 3 + 1];
        var += dyn_table[entry * 3 + 2];
    }
    const double dyn = (mean + dynamic_scale * sqrt(var)) * -0.1; /* U/QualityUtils.qualToErrorProbLog10(double) */
    return dyn < static_thr ? dyn : static_thr;
}

/* U/genotyper/AlleleLikelihoods.java:1351-1376, 1198-1208: keep[r] = 0 when the read is removed.
 * lk_allele_major is the normalised matrix [h * n_reads + r]. */
void region_oracle_filter(const double *lk_allele_major, int n_reads, int n_haps, const uint8_t *hmm_base_q,
                          const int64_t *read_off, double max_error_per_base, int dynamic, double dynamic_scale,
                          uint8_t *keep)
{
    for (int r = 0; r < n_reads; ++r) {
        double best = -INFINITY;
        for (int h = 0; h < n_haps; ++h)
            if (lk_allele_major[(size_t)h * n_reads + r] > best) best = lk_allele_major[(size_t)h * n_reads + r];
        const int n = (int)(read_off[r + 1] - read_off[r]);
        const double thr = region_oracle_min_true_likelihood(hmm_base_q + read_off[r], n, max_error_per_base, dynamic, dynamic_scale);
        keep[r] = best < thr ? 0 : 1;
    }
}
