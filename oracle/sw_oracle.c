/*
 * sw_oracle.c -- CPU restatement of GATK's Java Smith-Waterman aligner, used ONLY as a
 * test oracle.  THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (same rules as pairhmm_oracle.c).
 *
 * Parity status: PINNED to the reference's known answers, SmithWatermanAlignerAbstractUnitTest.java:143-281
 * (13 expected (offset, CIGAR) pairs over the four overhang strategies + the flank-invariance property), restated in
 * tests/test_smith_waterman.py.
 *
 * Follows src/main/java/org/broadinstitute/hellbender/utils/smithwaterman/SmithWatermanJavaAligner.java
 *   :60-92    align(): exact-substring shortcut (Utils.lastIndexOf, utils/Utils.java:1124-1138) for SOFTCLIP / IGNORE
 *   :105-215  calculateMatrix(): scores and backtrack with the best-gap arrays, ties resolved diag >= right >= down
 *   :262-375  calculateCigar(): end point by strategy, traceback, soft clips / overhang handling
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { SW_SOFTCLIP = 0, SW_INDEL = 1, SW_LEADING_INDEL = 2, SW_IGNORE = 3 };
enum { SW_OP_M = 0, SW_OP_I = 1, SW_OP_D = 2, SW_OP_S = 3 };

static int last_index_of(const uint8_t *ref, int n, const uint8_t *q, int m) {
    for (int r = n - m; r >= 0; r--) {
        int k = 0;
        while (k < m && ref[r + k] == q[k]) k++;
        if (k == m) return r;
    }
    return -1;
}

/* elems[k] = (length << 4) | op, in CIGAR order.  Returns the number of elements, or -1 if it exceeds cap. */
int sw_oracle_align(const uint8_t *ref, int n_ref, const uint8_t *alt, int n_alt, int w_match, int w_mismatch, int w_open,
                    int w_extend, int strategy, uint32_t *elems, int cap, int *offset_out) {
    if (n_ref <= 0 || n_alt <= 0) return -2; /* :64-66 IllegalArgumentException */
    if (strategy == SW_SOFTCLIP || strategy == SW_IGNORE) {
        const int at = last_index_of(ref, n_ref, alt, n_alt);
        if (at != -1) {
            if (cap < 1) return -1;
            elems[0] = ((uint32_t)n_alt << 4) | SW_OP_M;
            *offset_out = at;
            return 1;
        }
    }
    const int n = n_ref + 1, m = n_alt + 1;
    int *sw = (int *)calloc((size_t)n * m, sizeof(int)), *bt = (int *)calloc((size_t)n * m, sizeof(int));
    int *best_gap_v = (int *)malloc((size_t)(m + 1) * sizeof(int)), *gap_size_v = (int *)calloc((size_t)(m + 1), sizeof(int));
    int *best_gap_h = (int *)malloc((size_t)(n + 1) * sizeof(int)), *gap_size_h = (int *)calloc((size_t)(n + 1), sizeof(int));
    const int MATRIX_MIN_CUTOFF = (int)-1.0e8, lowInitValue = INT32_MIN / 2;
    for (int j = 0; j <= m; j++) best_gap_v[j] = lowInitValue;
    for (int i = 0; i <= n; i++) best_gap_h[i] = lowInitValue;
#define SW(i, j) sw[(size_t)(i) * m + (j)]
#define BT(i, j) bt[(size_t)(i) * m + (j)]
    if (strategy == SW_INDEL || strategy == SW_LEADING_INDEL) {
        int cur = w_open;
        if (m > 1) SW(0, 1) = w_open;
        for (int j = 2; j < m; j++) { cur += w_extend; SW(0, j) = cur; }
        cur = w_open;
        if (n > 1) SW(1, 0) = w_open;
        for (int i = 2; i < n; i++) { cur += w_extend; SW(i, 0) = cur; }
    }
    for (int i = 1; i < n; i++) {
        const uint8_t a = ref[i - 1];
        for (int j = 1; j < m; j++) {
            const uint8_t b = alt[j - 1];
            const int step_diag = SW(i - 1, j - 1) + (a == b ? w_match : w_mismatch);
            int prev_gap = SW(i - 1, j) + w_open;
            best_gap_v[j] += w_extend;
            if (prev_gap > best_gap_v[j]) { best_gap_v[j] = prev_gap; gap_size_v[j] = 1; } else gap_size_v[j]++;
            const int step_down = best_gap_v[j], kd = gap_size_v[j];
            prev_gap = SW(i, j - 1) + w_open;
            best_gap_h[i] += w_extend;
            if (prev_gap > best_gap_h[i]) { best_gap_h[i] = prev_gap; gap_size_h[i] = 1; } else gap_size_h[i]++;
            const int step_right = best_gap_h[i], ki = gap_size_h[i];
            if (step_diag >= step_down && step_diag >= step_right) {
                SW(i, j) = step_diag > MATRIX_MIN_CUTOFF ? step_diag : MATRIX_MIN_CUTOFF; BT(i, j) = 0;
            } else if (step_right >= step_down) {
                SW(i, j) = step_right > MATRIX_MIN_CUTOFF ? step_right : MATRIX_MIN_CUTOFF; BT(i, j) = -ki;
            } else {
                SW(i, j) = step_down > MATRIX_MIN_CUTOFF ? step_down : MATRIX_MIN_CUTOFF; BT(i, j) = kd;
            }
        }
    }
    /* calculateCigar */
    int p1 = 0, p2 = 0, maxscore = INT32_MIN, segment_length = 0;
    const int refLength = n - 1, altLength = m - 1;
    if (strategy == SW_INDEL) {
        p1 = refLength; p2 = altLength;
    } else {
        p2 = altLength;
        for (int i = 1; i < n; i++) {
            const int cur = SW(i, altLength);
            if (cur >= maxscore) { p1 = i; maxscore = cur; }
        }
        if (strategy != SW_LEADING_INDEL) {
            for (int j = 1; j < m; j++) {
                const int cur = SW(refLength, j);
                if (cur > maxscore || (cur == maxscore && abs(refLength - j) < abs(p1 - p2))) {
                    p1 = refLength; p2 = j; maxscore = cur; segment_length = altLength - j;
                }
            }
        }
    }
    /* elements are produced back to front */
    uint32_t *rev = (uint32_t *)malloc((size_t)(n + m + 4) * sizeof(uint32_t));
    int nrev = 0;
#define PUSH(op, len) rev[nrev++] = ((uint32_t)(len) << 4) | (uint32_t)(op)
    if (segment_length > 0 && strategy == SW_SOFTCLIP) { PUSH(SW_OP_S, segment_length); segment_length = 0; }
    int state = SW_OP_M;
    do {
        const int btr = BT(p1, p2);
        int new_state, step_length = 1;
        if (btr > 0) { new_state = SW_OP_D; step_length = btr; }
        else if (btr < 0) { new_state = SW_OP_I; step_length = -btr; }
        else new_state = SW_OP_M;
        if (new_state == SW_OP_M) { p1--; p2--; } else if (new_state == SW_OP_I) p2 -= step_length; else p1 -= step_length;
        if (new_state == state) segment_length += step_length;
        else {
            if (segment_length > 0) PUSH(state, segment_length);
            segment_length = step_length;
            state = new_state;
        }
    } while (p1 > 0 && p2 > 0);
    int offset;
    if (strategy == SW_SOFTCLIP) {
        PUSH(state, segment_length);
        if (p2 > 0) PUSH(SW_OP_S, p2);
        offset = p1;
    } else if (strategy == SW_IGNORE) {
        PUSH(state, segment_length + p2);
        offset = p1 - p2;
    } else {
        PUSH(state, segment_length);
        if (p1 > 0) PUSH(SW_OP_D, p1); else if (p2 > 0) PUSH(SW_OP_I, p2);
        offset = 0;
    }
    int rc = nrev;
    if (nrev > cap) rc = -1;
    else for (int k = 0; k < nrev; k++) elems[k] = rev[nrev - 1 - k];
    *offset_out = offset;
    free(rev); free(sw); free(bt); free(best_gap_v); free(gap_size_v); free(best_gap_h); free(gap_size_h);
    return rc;
}
