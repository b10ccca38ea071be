// region_steps.cuh -- the per-read steps either side of the forward kernels, on the device:
//
//   phmm_modify_quals_kernel      PairHMMLikelihoodCalculationEngine.modifyReadQualities: PCR indel error model
//                                 (tandem-repeat scan per base) + quality caps, in place on the chunk's read arrays
//                                 BEFORE classification, so the forward kernels see exactly what LoglessPairHMM would
//   phmm_normalize_filter_kernel  AlleleLikelihoods.normalizeLikelihoods + the decision of filterPoorlyModeledEvidence,
//                                 AFTER the fp64 redo; also turns the read-major result into the allele-major matrix
//                                 AlleleLikelihoods stores, so the Java scatter loop (VectorLoglessPairHMM.java:141-153)
//                                 becomes a block copy
//
// Integer work is bit-exact with the reference; reference lines are cited per function
// (HC/ = src/main/java/org/broadinstitute/hellbender/tools/walkers/haplotypecaller/, U/ = .../hellbender/utils/).
#pragma once
#include <stdint.h>
#include <math.h>
#include "phmm_kernels.cuh"  // UnitDesc

namespace phmm_dev {

constexpr int RS_MAX_STR_UNIT = 8;     // HC/ReadLikelihoodCalculationEngine.java:25 MAX_STR_UNIT_LENGTH
constexpr int RS_MAX_REPEAT = 20;      // HC/ReadLikelihoodCalculationEngine.java:26 MAX_REPEAT_LENGTH
constexpr uint8_t RS_MIN_USABLE_Q = 6; // U/QualityUtils.java MIN_USABLE_Q_SCORE

struct ModifyArgs {
    const uint8_t *rd_bases;
    uint8_t *rd_q, *rd_i, *rd_d;  // modified in place
    const uint32_t *read_off;     // chunk-local
    uint32_t n_reads;
    const uint8_t *mapq;          // chunk-local, one per read
    uint8_t pcr_cache[RS_MAX_REPEAT + 4];  // pcrIndelErrorModelCache (HC/PairHMMLikelihoodCalculationEngine.java:343-354)
    int32_t has_pcr;              // pcrErrorModel != NONE
    int32_t bq_threshold;         // baseQualityScoreThreshold (a Java byte)
    int32_t disable_cap_to_mapq;
};

// does b[x, x+len) equal b[y, y+len) ?
__device__ __forceinline__ bool rs_same_block(const uint8_t *b, int x, int y, int len) {
    for (int k = 0; k < len; ++k)
        if (b[x + k] != b[y + k]) return false;
    return true;
}

// copies of the unit b[u, u+len) that END the prefix b[0, end]  (findNumberOfRepetitions(..., leadingRepeats=false),
// U/variant/GATKVariantContextUtils.java:1000-1010).  Counting stops at RS_MAX_REPEAT: callers only use min(sum, 20).
__device__ __forceinline__ int rs_trailing_copies(const uint8_t *b, int u, int len, int end) {
    int n = 0;
    for (int s = end + 1 - len; s >= 0 && n < RS_MAX_REPEAT; s -= len) {
        if (!rs_same_block(b, s, u, len)) break;
        ++n;
    }
    return n;
}

// copies of the unit b[u, u+len) that START the suffix b[from, n)  (leadingRepeats=true, :988-998)
__device__ __forceinline__ int rs_leading_copies(const uint8_t *b, int u, int len, int from, int n) {
    int c = 0;
    for (int s = from; s + len <= n && c < RS_MAX_REPEAT; s += len) {
        if (!rs_same_block(b, s, u, len)) break;
        ++c;
    }
    return c;
}

// Repeat length at read offset o (HC/ReadLikelihoodCalculationEngine.findTandemRepeatUnits :193-253, right member).
// Backward: the shortest unit (1..8 bases) ending at o that occurs at least twice back to back; forward likewise from
// o+1.  When nothing repeats the unit is the single base and the count 1.  Same unit on both sides: the counts add;
// otherwise the forward unit is also counted backwards from o.
__device__ int rs_tandem_repeat_length(const uint8_t *b, int n, int o) {
    int bw_len = 1, bw_cnt = 1;
    for (int len = 1; len <= RS_MAX_STR_UNIT && len <= o + 1; ++len) {
        const int c = rs_trailing_copies(b, o - len + 1, len, o);
        if (c > 1) { bw_len = len; bw_cnt = c; break; }
    }
    int total = bw_cnt;
    if (o < n - 1) {
        int fw_len = 1, fw_cnt = 1;
        for (int len = 1; len <= RS_MAX_STR_UNIT && o + len + 1 <= n; ++len) {
            const int c = rs_leading_copies(b, o + 1, len, o + 1, n);
            if (c > 1) { fw_len = len; fw_cnt = c; break; }
        }
        if (fw_len == bw_len && rs_same_block(b, o + 1, o - bw_len + 1, fw_len))
            total = bw_cnt + fw_cnt;
        else
            total = fw_cnt + rs_trailing_copies(b, o + 1, fw_len, o);
    }
    return min(total, RS_MAX_REPEAT);
}

// Java compares these as signed bytes (HC/PairHMMLikelihoodCalculationEngine.java:311-313)
__device__ __forceinline__ uint8_t rs_floor_qual(uint8_t v, int min_qual) {
    return (int)(int8_t)v < (int)(int8_t)min_qual ? RS_MIN_USABLE_Q : v;
}

// one warp per read (HC/PairHMMLikelihoodCalculationEngine.java:283-306, 361-371)
__global__ void __launch_bounds__(128) phmm_modify_quals_kernel(const ModifyArgs a)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < a.n_reads; r += n_warps) {
        const uint32_t ro = a.read_off[r];
        const int n = (int)(a.read_off[r + 1] - ro);
        const uint8_t *b = a.rd_bases + ro;
        const int mapq = a.mapq[r];
        for (int i = (int)lane; i < n; i += 32) {
            uint8_t q = a.rd_q[ro + i], qi = a.rd_i[ro + i], qd = a.rd_d[ro + i];
            if (a.has_pcr && i < n - 1) {  // applyPCRErrorModel touches bases 0 .. n-2
                const uint8_t cap = a.pcr_cache[rs_tandem_repeat_length(b, n, i)];
                qi = min(qi, cap);
                qd = min(qd, cap);
            }
            if (!a.disable_cap_to_mapq) q = (uint8_t)min((int)q, mapq);
            a.rd_q[ro + i] = rs_floor_qual(q, a.bq_threshold);
            a.rd_i[ro + i] = rs_floor_qual(qi, RS_MIN_USABLE_Q);
            a.rd_d[ro + i] = rs_floor_qual(qd, RS_MIN_USABLE_Q);
        }
    }
}

// HC/ReadLikelihoodCalculationEngine.java:160-191 dynamicReadQualThreshLookupTable, mean and variance per base quality 1..40
__constant__ double c_dyn_mean[40] = {
    5.996842844, 5.870018422, 5.401558531, 4.818940919, 4.218758304, 3.646319832, 3.122346753, 2.654731979,
    2.244479156, 1.88893867,  1.583645342, 1.3233807,   1.102785365, 0.916703025, 0.760361881, 0.629457387,
    0.520175654, 0.42918208,  0.353590663, 0.290923699, 0.23906788,  0.196230431, 0.160897421, 0.131795374,
    0.1078567,   0.088189063, 0.072048567, 0.058816518, 0.047979438, 0.039111985, 0.031862437, 0.025940415,
    0.021106532, 0.017163711, 0.013949904, 0.011332027, 0.009200898, 0.007467036, 0.006057179, 0.004911394};
__constant__ double c_dyn_var[40] = {
    0.196616587, 1.388545569, 5.641990128, 10.33176216, 14.25799688, 17.02880749, 18.64537883, 19.27521677,
    19.13584613, 18.43922003, 17.36842261, 16.07088712, 14.65952563, 13.21718577, 11.80207947, 10.45304833,
    9.194183767, 8.038657241, 6.991779595, 6.053379213, 5.219610436, 4.484302033, 3.839943445, 3.27839108,
    2.791361596, 2.370765375, 2.008921719, 1.698687797, 1.433525748, 1.207526336, 1.015402928, 0.852465956,
    0.714585285, 0.598145851, 0.500000349, 0.41742159,  0.348056286, 0.289881373, 0.241163527, 0.200422214};

struct PostArgs {
    const UnitDesc *units;
    uint32_t n_units;
    const double *lk;         // read-major per unit: lk[out_base + r * n_haps + h]
    double *out;              // allele-major per unit: out[out_base + h * n_reads + r]
    uint8_t *keep;            // keep[keep_base + r]: 0 = filterPoorlyModeledEvidence removes the read
    const uint8_t *rd_q;      // modified base qualities (HMM_BASE_QUALITIES_TAG)
    const uint32_t *read_off;
    double max_diff_cap;      // log10globalReadMismappingRate (< 0; -inf: no capping)
    double max_error_per_base;
    double dynamic_scale;     // readDisqualificationScale
    int32_t symmetric;        // symmetricallyNormalizeAllelesToReference
    int32_t filter;           // computeReadLikelihoods(..., filterPoorly)
    int32_t dynamic;          // dynamicDisqualification (DRAGEN-GATK)
};

// minimum log10 likelihood the best allele of a read must reach (HC/ReadLikelihoodCalculationEngine.java:95-113 and,
// for the DRAGEN dynamic model, :66-81,118-151)
__device__ double rs_min_true_likelihood(const uint8_t *q, int n, const PostArgs &a) {
    const double errs = ceil((double)n * a.max_error_per_base);
    if (!a.dynamic) return fmin(2.0, errs) * -4.0;
    double mean = 0.0, var = 0.0;
    for (int i = 0; i < n; ++i) {
        const int bq = q[i];
        const int entry = bq <= 1 ? 0 : min(40, bq) - 1;
        mean += c_dyn_mean[entry];
        var += c_dyn_var[entry];
    }
    const double dyn = (mean + a.dynamic_scale * sqrt(var)) * -0.1;  // QualityUtils.qualToErrorProbLog10(double)
    const double fixed = errs * -4.0;
    return dyn < fixed ? dyn : fixed;
}

// one CTA per unit at a time, one thread per read.  Per read (U/genotyper/AlleleLikelihoods.java:416-458):
// best = max over alleles (the reference haplotype only competes when `symmetric`, searchBestAllele :505-533), every
// allele below best + cap is raised to it; then (:1351-1376) the read is dropped when its maximum over ALL alleles is
// below the threshold.  Writes are allele-major, i.e. consecutive threads write consecutive doubles.
__global__ void __launch_bounds__(128) phmm_normalize_filter_kernel(const PostArgs a)
{
    for (uint32_t u = blockIdx.x; u < a.n_units; u += gridDim.x) {
        const UnitDesc d = a.units[u];
        const int nh = (int)d.n_haps, ref = d.ref_hap;
        for (uint32_t r = threadIdx.x; r < d.n_reads; r += blockDim.x) {
            const double *row = a.lk + d.out_base + (size_t)r * nh;
            double *col = a.out + d.out_base + r;
            double top = -INFINITY;  // maximumLikelihoodOverAllAlleles (:1198-1208)
            for (int h = 0; h < nh; ++h) top = row[h] > top ? row[h] : top;
            double cap = -INFINITY;
            if (nh > 1 && a.max_diff_cap != -INFINITY) {
                const int first = (a.symmetric || ref != 0) ? 0 : 1;
                double best = row[first];
                for (int h = first + 1; h < nh; ++h) {
                    if (!a.symmetric && h == ref) continue;
                    best = row[h] > best ? row[h] : best;
                }
                cap = best + a.max_diff_cap;
            }
            for (int h = 0; h < nh; ++h) {
                const double v = row[h];
                col[(size_t)h * d.n_reads] = v < cap ? cap : v;
            }
            uint8_t keep = 1;
            if (a.filter && nh > 0) {
                const uint32_t ro = a.read_off[d.read_first + r];
                const int n = (int)(a.read_off[d.read_first + r + 1] - ro);
                keep = top < rs_min_true_likelihood(a.rd_q + ro, n, a) ? 0 : 1;
            }
            a.keep[d.keep_base + r] = keep;
        }
    }
}

}  // namespace phmm_dev
