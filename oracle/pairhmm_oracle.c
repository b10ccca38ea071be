/*
 * pairhmm_oracle.c -- CPU restatement of GATK's Java PairHMM, used ONLY as a test oracle.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product path
 * (gatk_b200/csrc, libgpuphmm.so) never links, loads or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks this file against the reference's
 * own fixtures (converted by tests/golden/make_golden.py):
 *   - src/test/resources/pairhmm-testdata.txt                               (104 pairs, tol 1e-5,
 *     VectorPairHMMUnitTest.java:24,100)
 *   - src/test/resources/org/broadinstitute/hellbender/tools/haplotypecaller/
 *     expected.{Java,Exact,Original,AVX}.hmmresults.txt                     (284 pairs each,
 *     HaplotypeCallerIntegrationTest.java:2193-2244; the Java file is an exact-text fixture)
 *
 * Each function cites the reference lines it follows.  Path prefixes:
 *   PH/ = src/main/java/org/broadinstitute/hellbender/utils/pairhmm/
 *   U/  = src/main/java/org/broadinstitute/hellbender/utils/
 *
 * Java double arithmetic is IEEE-754 binary64 without FMA contraction: compile with
 * -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAX_QUAL 254 /* U/QualityUtils.java:43 */

/* PH/PairHMMModel.java:26-51 : positions in the per-base transition array */
enum { matchToMatch = 0, indelToMatch = 1, matchToInsertion = 2, insertionToInsertion = 3,
       matchToDeletion = 4, deletionToDeletion = 5, TRANS_PROB_ARRAY_LENGTH = 6 };

/* error codes (mirroring the Java exceptions at the call sites cited) */
#define ORACLE_OK 0
#define ORACLE_ERR_BAD_QUAL (-1)   /* PH/PairHMMModel.java:109-111 IllegalArgumentException;
                                      U/QualityUtils.java:157 cache index 255 is out of bounds */
#define ORACLE_ERR_BAD_LENGTH (-2) /* PH/PairHMM.java:162-164 -> :139 haplotypeMaxLength must be > 0 */

static double qualToErrorProbCache[MAX_QUAL + 1];  /* U/QualityUtils.java:48-57 */
static double qualToProbLog10Cache[MAX_QUAL + 1];
static double *jacobianCache;                      /* U/MathUtils.java:406-423 */
#define JACOBIAN_MAX_TOLERANCE 8.0
#define JACOBIAN_TABLE_STEP 0.0001
static const double JACOBIAN_INV_STEP = 1.0 / JACOBIAN_TABLE_STEP;
#define M2M_LEN (((MAX_QUAL + 1) * (MAX_QUAL + 2)) >> 1)
static double matchToMatchProbTab[M2M_LEN];        /* PH/PairHMMModel.java:71 */
static double matchToMatchLog10Tab[M2M_LEN];       /* PH/PairHMMModel.java:81 */
static int tables_ready = 0;

/* U/MathUtils.java:428-430 */
static int fastRound(double d) { return (d > 0.0) ? (int)(d + 0.5) : (int)(d - 0.5); }

/* U/MathUtils.java:467-480 */
static double approximateLog10SumLog10(double a, double b) {
    if (a > b) { double t = a; a = b; b = t; }
    if (a == -INFINITY) return b;
    const double diff = b - a;
    return b + (diff < JACOBIAN_MAX_TOLERANCE ? jacobianCache[fastRound(diff * JACOBIAN_INV_STEP)] : 0.0);
}

/* U/MathUtils.java:435-456 (array form, used by Log10PairHMM ORIGINAL) */
static double approximateLog10SumLog10_arr(const double *vals, int n) {
    int maxIdx = 0;
    for (int i = 1; i < n; i++) if (vals[i] > vals[maxIdx]) maxIdx = i;
    double approxSum = vals[maxIdx];
    for (int i = 0; i < n; i++) {
        if (i == maxIdx || vals[i] == -INFINITY) continue;
        const double diff = approxSum - vals[i];
        approxSum += diff < JACOBIAN_MAX_TOLERANCE ? jacobianCache[fastRound(diff * JACOBIAN_INV_STEP)] : 0.0;
    }
    return approxSum;
}

/* U/MathUtils.java:622-646 (used by Log10PairHMM EXACT) */
static double log10SumLog10_arr(const double *vals, int n) {
    int maxIdx = 0;
    for (int i = 1; i < n; i++) if (vals[i] > vals[maxIdx]) maxIdx = i;
    const double maxValue = vals[maxIdx];
    if (maxValue == -INFINITY) return maxValue;
    double sum = 1.0;
    for (int i = 0; i < n; i++) {
        if (i == maxIdx || vals[i] == -INFINITY) continue;
        sum += pow(10.0, vals[i] - maxValue);
    }
    return maxValue + (sum != 1.0 ? log10(sum) : 0.0);
}

/* Class-load-time table construction.
 * U/QualityUtils.java:51-57,138-141 ; U/MathUtils.java:414-416 ; PH/PairHMMModel.java:86-94 */
void phmm_oracle_init(void) {
    if (tables_ready) return;
    for (int i = 0; i <= MAX_QUAL; i++) {
        qualToErrorProbCache[i] = pow(10.0, (double)i / -10.0);
        qualToProbLog10Cache[i] = log10(1.0 - qualToErrorProbCache[i]);
    }
    const int n = (int)(JACOBIAN_MAX_TOLERANCE / JACOBIAN_TABLE_STEP) + 1;
    jacobianCache = (double *)malloc(sizeof(double) * (size_t)n);
    for (int k = 0; k < n; k++) jacobianCache[k] = log10(1.0 + pow(10.0, -k * JACOBIAN_TABLE_STEP));
    const double LN10 = log(10.0), INV_LN10 = 1.0 / LN10;
    for (int i = 0, offset = 0; i <= MAX_QUAL; offset += ++i)
        for (int j = 0; j <= i; j++) {
            const double log10Sum = approximateLog10SumLog10(-0.1 * i, -0.1 * j);
            matchToMatchLog10Tab[offset + j] = log1p(-fmin(1.0, pow(10.0, log10Sum))) * INV_LN10;
            matchToMatchProbTab[offset + j] = pow(10.0, matchToMatchLog10Tab[offset + j]);
        }
    tables_ready = 1;
}

/* U/QualityUtils.java:156-158 (index = qual & 0xff; 255 would throw ArrayIndexOutOfBounds) */
double phmm_oracle_qual_to_error_prob(int qual) { phmm_oracle_init(); return qualToErrorProbCache[qual & 0xff]; }
/* U/QualityUtils.java:83-85 */
static double qualToProb(int qual) { return 1.0 - qualToErrorProbCache[qual & 0xff]; }

/* PH/PairHMMModel.java:373-388 */
double phmm_oracle_match_to_match_prob(int insQual, int delQual) {
    phmm_oracle_init();
    int minQual = insQual <= delQual ? insQual : delQual;
    int maxQual = insQual <= delQual ? delQual : insQual;
    return (MAX_QUAL < maxQual) ? 1.0 - pow(10.0, approximateLog10SumLog10(-0.1 * minQual, -0.1 * maxQual))
                                : matchToMatchProbTab[((maxQual * (maxQual + 1)) >> 1) + minQual];
}
/* PH/PairHMMModel.java:403-418 */
static double matchToMatchProbLog10(int insQual, int delQual) {
    int minQual = insQual <= delQual ? insQual : delQual;
    int maxQual = insQual <= delQual ? delQual : insQual;
    return (MAX_QUAL < maxQual)
               ? log1p(-fmin(1.0, pow(10.0, approximateLog10SumLog10(-.1 * minQual, -.1 * maxQual)))) * (1.0 / log(10.0))
               : matchToMatchLog10Tab[((maxQual * (maxQual + 1)) >> 1) + minQual];
}

/* PH/PairHMMModel.java:107-117.  dest has TRANS_PROB_ARRAY_LENGTH entries.
 * Java takes signed bytes and rejects negative ones (:109-111). */
int phmm_oracle_qual_to_trans_probs(double *dest, uint8_t insQual, uint8_t delQual, uint8_t gcp) {
    phmm_oracle_init();
    if (insQual > 127 || delQual > 127 || gcp > 127) return ORACLE_ERR_BAD_QUAL;
    dest[matchToMatch] = phmm_oracle_match_to_match_prob(insQual, delQual);
    dest[matchToInsertion] = qualToErrorProbCache[insQual];
    dest[matchToDeletion] = qualToErrorProbCache[delQual];
    dest[indelToMatch] = qualToProb(gcp);
    dest[insertionToInsertion] = dest[deletionToDeletion] = qualToErrorProbCache[gcp];
    return ORACLE_OK;
}

/* PH/PairHMMModel.java:214-225 */
static int qualToTransProbsLog10(double *dest, uint8_t insQual, uint8_t delQual, uint8_t gcp) {
    if (insQual > 127 || delQual > 127 || gcp > 127) return ORACLE_ERR_BAD_QUAL;
    dest[matchToMatch] = matchToMatchProbLog10(insQual, delQual);
    dest[matchToInsertion] = insQual * -0.1; /* U/QualityUtils.java:174-191 */
    dest[matchToDeletion] = delQual * -0.1;
    dest[indelToMatch] = qualToProbLog10Cache[gcp];
    dest[insertionToInsertion] = gcp * -0.1;
    dest[deletionToDeletion] = gcp * -0.1;
    return ORACLE_OK;
}

/*
 * LoglessPairHMM.subComputeReadLikelihoodGivenHaplotypeLog10 with hapStartIndex = 0 and
 * recacheReadValues = true, which is what PairHMM.computeLog10Likelihoods always passes
 * (PH/PairHMM.java:226-229,297).
 *
 *   PH/LoglessPairHMM.java:8-9   INITIAL_CONDITION = 2^1020 and its log10
 *   PH/LoglessPairHMM.java:30-36 deletionMatrix[0][j] = INITIAL_CONDITION / hapLen for all j
 *   PH/LoglessPairHMM.java:38-43 transition table per read base (qualToTransProbs)
 *   PH/LoglessPairHMM.java:84-92 priors
 *   PH/LoglessPairHMM.java:47-57 the recurrence (operation order kept as written)
 *   PH/LoglessPairHMM.java:62-67 final sum over the last row of M and I
 *
 * The Java code stores full (R+1)x(H+1) matrices (PH/N2MemoryPairHMM.java:27-36); only rows i-1
 * and i are ever read, so two rolling rows give bit-identical arithmetic.  Row 0 of M and I and
 * column 0 of every row >= 1 are Java's default 0.0.
 *
 * tristate_off != 0 mirrors doNotUseTristateCorrection() (PH/PairHMM.java:118-119), which the
 * reference's PairHMMUnitTest turns on.
 * Returns the log10 likelihood in *out.  A zero-length read yields log10(0) = -inf like Java.
 */
int phmm_oracle_logless(const uint8_t *hap, int H, const uint8_t *read, const uint8_t *baseQ,
                        const uint8_t *insQ, const uint8_t *delQ, const uint8_t *gcp, int R,
                        int tristate_off, double *out) {
    phmm_oracle_init();
    if (H <= 0) return ORACLE_ERR_BAD_LENGTH;
    static const double TRISTATE_CORRECTION = 3.0;
    const double INITIAL_CONDITION = pow(2.0, 1020.0);
    const double INITIAL_CONDITION_LOG10 = log10(INITIAL_CONDITION);
    const int W = H + 1;
    double *buf = (double *)calloc((size_t)6 * W, sizeof(double));
    double *Mp = buf, *Ip = buf + W, *Dp = buf + 2 * W, *Mc = buf + 3 * W, *Ic = buf + 4 * W, *Dc = buf + 5 * W;
    const double initialValue = INITIAL_CONDITION / H;
    for (int j = 0; j < W; j++) Dp[j] = initialValue;
    int rc = ORACLE_OK;
    for (int i = 1; i <= R; i++) {
        double t[TRANS_PROB_ARRAY_LENGTH];
        rc = phmm_oracle_qual_to_trans_probs(t, insQ[i - 1], delQ[i - 1], gcp[i - 1]);
        if (rc != ORACLE_OK) break;
        if (baseQ[i - 1] > MAX_QUAL) { rc = ORACLE_ERR_BAD_QUAL; break; }
        const uint8_t x = read[i - 1];
        const double pMatch = qualToProb(baseQ[i - 1]);
        const double pMis = qualToErrorProbCache[baseQ[i - 1]] / (tristate_off ? 1.0 : TRISTATE_CORRECTION);
        Mc[0] = 0.0; Ic[0] = 0.0; Dc[0] = 0.0;
        for (int j = 1; j < W; j++) {
            const uint8_t y = hap[j - 1];
            const double prior = (x == y || x == (uint8_t)'N' || y == (uint8_t)'N') ? pMatch : pMis;
            Mc[j] = prior * (Mp[j - 1] * t[matchToMatch] + Ip[j - 1] * t[indelToMatch] + Dp[j - 1] * t[indelToMatch]);
            Ic[j] = Mp[j] * t[matchToInsertion] + Ip[j] * t[insertionToInsertion];
            Dc[j] = Mc[j - 1] * t[matchToDeletion] + Dc[j - 1] * t[deletionToDeletion];
        }
        double *s;
        s = Mp; Mp = Mc; Mc = s;
        s = Ip; Ip = Ic; Ic = s;
        s = Dp; Dp = Dc; Dc = s;
    }
    if (rc == ORACLE_OK) {
        double finalSumProbabilities = 0.0;
        for (int j = 1; j < W; j++) finalSumProbabilities += Mp[j] + Ip[j];
        *out = log10(finalSumProbabilities) - INITIAL_CONDITION_LOG10;
    }
    free(buf);
    return rc;
}

/*
 * LoglessPDPairHMM.subComputeReadLikelihoodGivenHaplotypeLog10 (PH/LoglessPDPairHMM.java:34-153), the
 * "partially determined" PairHMM of DRAGEN-GATK mode, with hapStartIndex = 0 (PH/PDPairHMM.java:176,241).
 * Parity status of THIS function: UNPINNED -- the reference's only fixture for it,
 * src/test/resources/large/expected.PDHMM.hmmresults.txt (PDPairHMMLikelihoodCalculationEngineUnitTest.java:22),
 * is a git-lfs stub in /root/reference.  It is anchored instead by (i) pd bases all 0 == phmm_oracle_logless
 * exactly, (ii) an independent pure-Python restatement (oracle/pyoracle.py), (iii) hand-checked tiny cases.
 *
 *   :44-50   D[0][j] = 2^1020 / hapLen;  branch matrices row 0 / column 0 stay Java's 0.0
 *   :60-147  per cell, by state: NORMAL / INSIDE_DEL copy or hold the branch values, AFTER_DEL merges them with
 *            max(); the state advances along j from the PD flags DEL_START / DEL_END of column j and -- as written --
 *            is NOT reset at the start of a row (it is declared outside the row loop, :59)
 *   :170-181 priors: a SNP-flagged column also matches the alternative bases of its mask (:184-204)
 * pd: one byte per haplotype base, utils/haplotype/PartiallyDeterminedHaplotype.java:59-65
 *     SNP=1 DEL_START=2 DEL_END=4 A=8 C=16 G=32 T=64.
 * Returns ORACLE_ERR_BAD_QUAL - 10 when a read base other than ACGTacgt meets a SNP column (Java throws, :202).
 */
int phmm_oracle_pd_logless(const uint8_t *hap, const uint8_t *pd, int H, const uint8_t *read, const uint8_t *baseQ,
                           const uint8_t *insQ, const uint8_t *delQ, const uint8_t *gcp, int R,
                           int tristate_off, double *out) {
    phmm_oracle_init();
    if (H <= 0) return ORACLE_ERR_BAD_LENGTH;
    enum { PD_SNP = 1, PD_DEL_START = 2, PD_DEL_END = 4, PD_A = 8, PD_C = 16, PD_G = 32, PD_T = 64 };
    enum { NORMAL, INSIDE_DEL, AFTER_DEL };
    const double INITIAL_CONDITION = pow(2.0, 1020.0);
    const double INITIAL_CONDITION_LOG10 = log10(INITIAL_CONDITION);
    const int W = H + 1;
    /* 6 matrices x 2 rolling rows */
    double *buf = (double *)calloc((size_t)12 * W, sizeof(double));
    double *Mp = buf, *Ip = Mp + W, *Dp = Ip + W, *bMp = Dp + W, *bIp = bMp + W, *bDp = bIp + W;
    double *Mc = bDp + W, *Ic = Mc + W, *Dc = Ic + W, *bMc = Dc + W, *bIc = bMc + W, *bDc = bIc + W;
    const double initialValue = INITIAL_CONDITION / H;
    for (int j = 0; j < W; j++) Dp[j] = initialValue;
    int rc = ORACLE_OK;
    int state = NORMAL;
    for (int i = 1; i <= R && rc == ORACLE_OK; i++) {
        double t[TRANS_PROB_ARRAY_LENGTH];
        rc = phmm_oracle_qual_to_trans_probs(t, insQ[i - 1], delQ[i - 1], gcp[i - 1]);
        if (rc != ORACLE_OK) break;
        if (baseQ[i - 1] > MAX_QUAL) { rc = ORACLE_ERR_BAD_QUAL; break; }
        const uint8_t x = read[i - 1];
        const double pMatch = qualToProb(baseQ[i - 1]);
        const double pMis = qualToErrorProbCache[baseQ[i - 1]] / (tristate_off ? 1.0 : 3.0);
        Mc[0] = Ic[0] = Dc[0] = bMc[0] = bIc[0] = bDc[0] = 0.0;
        for (int j = 1; j < W; j++) {
            const uint8_t y = hap[j - 1], f = pd[j - 1];
            int match = x == y || x == (uint8_t)'N' || y == (uint8_t)'N';
            if (!match && (f & PD_SNP)) {
                switch (x) {
                    case 'A': case 'a': match = (f & PD_A) != 0; break;
                    case 'C': case 'c': match = (f & PD_C) != 0; break;
                    case 'T': case 't': match = (f & PD_T) != 0; break;
                    case 'G': case 'g': match = (f & PD_G) != 0; break;
                    default: rc = ORACLE_ERR_BAD_QUAL - 10; break;
                }
            }
            const double prior = match ? pMatch : pMis;
            const int del_end = (f & PD_DEL_END) == PD_DEL_END;
            double dm = Mp[j - 1], di = Ip[j - 1], dd = Dp[j - 1]; /* diagonal operands */
            double lm = Mc[j - 1], ld = Dc[j - 1];                  /* left operands of D */
            if (state == NORMAL) {
                bMc[j] = Mc[j - 1]; bDc[j] = Dc[j - 1]; bIc[j] = Ic[j - 1];
            } else if (state == INSIDE_DEL) {
                bMc[j] = bMc[j - 1]; bDc[j] = bDc[j - 1]; bIc[j] = bIc[j - 1];
            } else {
                bMc[j] = fmax(bMc[j - 1], Mc[j - 1]); bDc[j] = fmax(bDc[j - 1], Dc[j - 1]); bIc[j] = fmax(bIc[j - 1], Ic[j - 1]);
                dm = fmax(bMp[j - 1], Mp[j - 1]); di = fmax(bIp[j - 1], Ip[j - 1]); dd = fmax(bDp[j - 1], Dp[j - 1]);
                lm = fmax(bMc[j - 1], Mc[j - 1]); ld = fmax(bDc[j - 1], Dc[j - 1]);
            }
            Mc[j] = prior * (dm * t[matchToMatch] + di * t[indelToMatch] + dd * t[indelToMatch]);
            Dc[j] = lm * t[matchToDeletion] + ld * t[deletionToDeletion];
            if (del_end)
                Ic[j] = fmax(bMp[j], Mp[j]) * t[matchToInsertion] + fmax(bIp[j], Ip[j]) * t[insertionToInsertion];
            else
                Ic[j] = Mp[j] * t[matchToInsertion] + Ip[j] * t[insertionToInsertion];
            if (state == AFTER_DEL) state = NORMAL;
            if ((f & PD_DEL_START) == PD_DEL_START) state = INSIDE_DEL;
            if (del_end) state = AFTER_DEL;
        }
        double *s;
        s = Mp; Mp = Mc; Mc = s;    s = Ip; Ip = Ic; Ic = s;    s = Dp; Dp = Dc; Dc = s;
        s = bMp; bMp = bMc; bMc = s; s = bIp; bIp = bIc; bIc = s; s = bDp; bDp = bDc; bDc = s;
    }
    if (rc == ORACLE_OK) {
        double finalSumProbabilities = 0.0;
        for (int j = 1; j < W; j++) finalSumProbabilities += Mp[j] + Ip[j];
        *out = log10(finalSumProbabilities) - INITIAL_CONDITION_LOG10;
    }
    free(buf);
    return rc;
}

/*
 * Log10PairHMM (EXACT when exact != 0, ORIGINAL otherwise).
 *   PH/Log10PairHMM.java:33-41   matrices start at -inf
 *   PH/Log10PairHMM.java:58-89   driver
 *   PH/Log10PairHMM.java:91-96   deletionMatrix[0][j] = log10(1/hapLen)
 *   PH/Log10PairHMM.java:98-105  final sum, folded left to right
 *   PH/Log10PairHMM.java:118-133 priors in log10 space
 *   PH/Log10PairHMM.java:175-183 updateCell
 */
int phmm_oracle_log10(const uint8_t *hap, int H, const uint8_t *read, const uint8_t *baseQ,
                      const uint8_t *insQ, const uint8_t *delQ, const uint8_t *gcp, int R,
                      int exact, int tristate_off, double *out) {
    phmm_oracle_init();
    if (H <= 0) return ORACLE_ERR_BAD_LENGTH;
    const double log10_3 = log10(3.0);
    const int W = H + 1;
    double *buf = (double *)malloc((size_t)6 * W * sizeof(double));
    for (int k = 0; k < 6 * W; k++) buf[k] = -INFINITY;
    double *Mp = buf, *Ip = buf + W, *Dp = buf + 2 * W, *Mc = buf + 3 * W, *Ic = buf + 4 * W, *Dc = buf + 5 * W;
    const double initialValue = log10(1.0 / H);
    for (int j = 0; j < W; j++) Dp[j] = initialValue;
    int rc = ORACLE_OK;
    for (int i = 1; i <= R; i++) {
        double t[TRANS_PROB_ARRAY_LENGTH];
        rc = qualToTransProbsLog10(t, insQ[i - 1], delQ[i - 1], gcp[i - 1]);
        if (rc != ORACLE_OK) break;
        if (baseQ[i - 1] > MAX_QUAL) { rc = ORACLE_ERR_BAD_QUAL; break; }
        const uint8_t x = read[i - 1];
        const double pMatch = qualToProbLog10Cache[baseQ[i - 1]];
        const double pMis = baseQ[i - 1] * -0.1 - (tristate_off ? 0.0 : log10_3);
        Mc[0] = Ic[0] = Dc[0] = -INFINITY;
        for (int j = 1; j < W; j++) {
            const uint8_t y = hap[j - 1];
            const double prior = (x == y || x == (uint8_t)'N' || y == (uint8_t)'N') ? pMatch : pMis;
            double a3[3] = { Mp[j - 1] + t[matchToMatch], Ip[j - 1] + t[indelToMatch], Dp[j - 1] + t[indelToMatch] };
            double i2[2] = { Mp[j] + t[matchToInsertion], Ip[j] + t[insertionToInsertion] };
            Mc[j] = prior + (exact ? log10SumLog10_arr(a3, 3) : approximateLog10SumLog10_arr(a3, 3));
            Ic[j] = exact ? log10SumLog10_arr(i2, 2) : approximateLog10SumLog10_arr(i2, 2);
            double d2[2] = { Mc[j - 1] + t[matchToDeletion], Dc[j - 1] + t[deletionToDeletion] };
            Dc[j] = exact ? log10SumLog10_arr(d2, 2) : approximateLog10SumLog10_arr(d2, 2);
        }
        double *s;
        s = Mp; Mp = Mc; Mc = s;
        s = Ip; Ip = Ic; Ic = s;
        s = Dp; Dp = Dc; Dc = s;
    }
    if (rc == ORACLE_OK) {
        double v2[2] = { Mp[1], Ip[1] };
        double acc = exact ? log10SumLog10_arr(v2, 2) : approximateLog10SumLog10_arr(v2, 2);
        for (int j = 2; j < W; j++) {
            double v3[3] = { acc, Mp[j], Ip[j] };
            acc = exact ? log10SumLog10_arr(v3, 3) : approximateLog10SumLog10_arr(v3, 3);
        }
        *out = acc;
    }
    free(buf);
    return rc;
}

/*
 * Driver over one (region, sample) unit, restating PairHMM.computeLog10Likelihoods
 * (PH/PairHMM.java:196-247): reads outer, alleles inner, result index r*nHaps + h (:236), which is
 * also the native layout VectorLoglessPairHMM reads back (PH/VectorLoglessPairHMM.java:148).
 * Flat SoA inputs: read_off[nReads+1] indexes the five per-base read arrays, hap_off[nHaps+1]
 * indexes hap_bases.  threads <= 1 runs serially (GATK runs this loop on one thread); threads > 1
 * splits reads over an OpenMP team only so bench.py can time the same arithmetic on all host cores.
 */
int phmm_oracle_unit(const uint8_t *read_bases, const uint8_t *base_q, const uint8_t *ins_q,
                     const uint8_t *del_q, const uint8_t *gcp, const int32_t *read_off, int n_reads,
                     const uint8_t *hap_bases, const int32_t *hap_off, int n_haps, int tristate_off,
                     int threads, double *out) {
    phmm_oracle_init();
    int rc_all = ORACLE_OK;
    (void)threads;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1) if (threads > 1)
#endif
    for (int r = 0; r < n_reads; r++) {
        const int ro = read_off[r], R = read_off[r + 1] - ro;
        for (int h = 0; h < n_haps; h++) {
            const int ho = hap_off[h], H = hap_off[h + 1] - ho;
            double v = NAN;
            int rc = phmm_oracle_logless(hap_bases + ho, H, read_bases + ro, base_q + ro, ins_q + ro,
                                         del_q + ro, gcp + ro, R, tristate_off, &v);
            out[(size_t)r * n_haps + h] = v;
            if (rc != ORACLE_OK) {
#ifdef _OPENMP
#pragma omp critical
#endif
                rc_all = rc;
            }
        }
    }
    return rc_all;
}

int phmm_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
