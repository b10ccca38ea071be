/*
 * gpuphmm.h -- C ABI of libgpuphmm.so, the B200 (sm_100a) PairHMM forward engine behind GATK's
 *              `-pairHMM CUDA_LOGLESS_CACHING`.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, no globals, no
 * callbacks.  It is what a JNI shim for GATK's native PairHMM binding calls; the shim is
 * gatk_b200/csrc/gpuphmm_jni.cpp and the Java side is under java/ (see INTEGRATION.md).
 *
 * Reference interfaces replaced (paths under /root/reference/src/main/java/org/broadinstitute/hellbender/):
 *   utils/pairhmm/VectorLoglessPairHMM.java:63      PairHMMNativeBinding.load(File)            -> gphmm_device_count / gphmm_create
 *   utils/pairhmm/VectorLoglessPairHMM.java:81      PairHMMNativeBinding.initialize(args)       -> gphmm_create(gphmm_config)
 *   utils/pairhmm/VectorLoglessPairHMM.java:138     PairHMMNativeBinding.computeLikelihoods(ReadDataHolder[], HaplotypeDataHolder[], double[])
 *                                                                                                -> gphmm_compute (one unit) / gphmm_submit + gphmm_wait (many units)
 *   utils/pairhmm/VectorLoglessPairHMM.java:164     PairHMMNativeBinding.done()                 -> gphmm_destroy
 *   tools/walkers/haplotypecaller/PairHMMNativeArgumentCollection.java:11-23 (maxNumberOfThreads, useDoublePrecision)
 *                                                                                                -> gphmm_config.host_threads / force_fp64
 *   utils/pairhmm/PairHMM.java:196-247 / :236       result layout out[r * nHaps + h] (log10)    -> gphmm_unit.out_off + r*nHaps + h
 *
 * Semantics: for every unit u (one sample's reads against one region's haplotypes) and every read r and
 * haplotype h of the unit, out[u.out_off + r*nHaps + h] = log10 P(read r | haplotype h) as defined by
 * utils/pairhmm/LoglessPairHMM.java:20-68.  Computation is fp32 on the GPU with an fp64 redo on the GPU of
 * every pair whose fp32 sum under-/overflows (or everything in fp64 when force_fp64 is set).  There is no CPU
 * fallback: without a usable CUDA device gphmm_create fails.
 *
 * Thread safety: a gphmm_t may be used from one thread at a time; several gphmm_t may coexist in a process.
 */
#ifndef GPUPHMM_H
#define GPUPHMM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPHMM_ABI_VERSION 1

/* error codes (0 = success, negative = failure) */
#define GPHMM_OK 0
#define GPHMM_ERR_INVALID_ARG (-1) /* null pointer, bad offsets, zero-length haplotype (PairHMM.java:139,284-292) */
#define GPHMM_ERR_NO_DEVICE (-2)   /* no CUDA device / not sm_100 (-> UserException.HardwareFeatureException) */
#define GPHMM_ERR_CUDA (-3)        /* a CUDA call failed; gphmm_last_error() has the text (-> GATKException) */
#define GPHMM_ERR_BAD_QUAL (-4)    /* ins/del/gcp qual > 127 or base qual 255 (PairHMMModel.java:109-111, QualityUtils.java:157) */
#define GPHMM_ERR_ALPHABET (-5)    /* more distinct haplotype byte values than the prior table supports */
#define GPHMM_ERR_NOMEM (-6)
#define GPHMM_ERR_BAD_TICKET (-7)
#define GPHMM_ERR_TOO_LARGE (-8)   /* a single unit exceeds the per-chunk device budget */

typedef struct gphmm gphmm_t;
typedef struct gphmm_prepared gphmm_prepared_t;

typedef struct gphmm_config {
    int32_t struct_size;      /* sizeof(gphmm_config); lets the struct grow compatibly */
    int32_t n_devices;        /* 0: use the current CUDA device only; >0: devices[0..n) */
    const int32_t *devices;   /* CUDA ordinals, may be NULL when n_devices == 0 */
    int32_t force_fp64;       /* PairHMMNativeArguments.useDoublePrecision */
    int32_t host_threads;     /* PairHMMNativeArguments.maxNumberOfThreads: planner threads, 0 = 4; a handle over several
                                 devices uses at least 3 per device */
    int32_t tristate_off;     /* PairHMM.doNotUseTristateCorrection() (tests only) */
    int32_t no_prefix_sharing; /* 1: recompute every haplotype from column 1 (A/B switch; results are bit-identical) */
    int64_t chunk_cells;      /* target DP cells per device chunk, 0 = default */
    int64_t chunk_bytes;      /* max staged input bytes per device chunk, 0 = default */
} gphmm_config;

/* One (region, sample) unit: reads [read_begin, read_end) against haplotypes [hap_begin, hap_end). */
typedef struct gphmm_unit {
    int64_t read_begin, read_end; /* indices into read_off */
    int64_t hap_begin, hap_end;   /* indices into hap_off */
    int64_t out_off;              /* first output slot of this unit */
} gphmm_unit;

/* Flat structure-of-arrays batch.  All pointers are host memory (pageable or pinned). */
typedef struct gphmm_batch {
    const uint8_t *read_bases; /* ReadDataHolder.readBases    (VectorLoglessPairHMM.java:123) */
    const uint8_t *base_q;     /* ReadDataHolder.readQuals    (:124) raw phred, not ASCII */
    const uint8_t *ins_q;      /* ReadDataHolder.insertionGOP (:125) */
    const uint8_t *del_q;      /* ReadDataHolder.deletionGOP  (:126) */
    const uint8_t *gcp;        /* ReadDataHolder.overallGCP   (:127) */
    const int64_t *read_off;   /* n_reads + 1 offsets shared by the five arrays above */
    int64_t n_reads;
    const uint8_t *hap_bases;  /* HaplotypeDataHolder.haplotypeBases (:98), concatenated */
    const int64_t *hap_off;    /* n_haps + 1 */
    int64_t n_haps;
    const gphmm_unit *units;
    int64_t n_units;
} gphmm_batch;

typedef struct gphmm_stats {
    int64_t pairs;           /* (read, haplotype) pairs computed */
    int64_t cells;           /* sum of R*H over pairs (LoglessPairHMM.java:47-49, no padding) */
    int64_t rescued_pairs;   /* pairs redone in fp64 */
    int64_t skipped_cells;   /* cells NOT executed thanks to haplotype-prefix sharing (still counted in `cells`) */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int64_t kernel_launches; /* CUDA kernels launched by this library */
    double fp32_kernel_ms;   /* CUDA-event time of the fp32 forward kernels */
    double fp64_kernel_ms;   /* CUDA-event time of the fp64 forward kernels */
    double device_ms;        /* CUDA-event time of everything queued on the compute streams */
    double host_stage_ms;    /* wall time spent packing staging buffers */
    double wall_ms;          /* wall time inside compute/submit/wait/run_prepared */
} gphmm_stats;

int gphmm_abi_version(void);
/* Number of usable devices (compute capability 10.x); 0 when there is none or the driver is absent. */
int gphmm_device_count(void);
const char *gphmm_strerror(int code);

int gphmm_create(const gphmm_config *cfg, gphmm_t **out);
void gphmm_destroy(gphmm_t *h);
/* Text of the last failure on this handle (never NULL).
 * Threads: a handle may be used from several threads (calls that touch the device are serialised inside; several
 * handles run independently -- one PairHMM instance per Spark task, J/tools/HaplotypeCallerSpark.java:175), but the
 * error text is one slot per handle: read it on the thread whose call failed, before that handle fails again. */
const char *gphmm_last_error(const gphmm_t *h);

/* Synchronous: returns when every out[] slot of the batch is written. */
int gphmm_compute(gphmm_t *h, const gphmm_batch *batch, double *out);

/* Asynchronous cross-region batching queue.  submit copies the batch into staging buffers before it
 * returns (the caller may reuse its input arrays at once, as JNI requires) and queues it; `out` must stay
 * valid until gphmm_wait(ticket) returns.  Tickets complete in submission order. */
int gphmm_submit(gphmm_t *h, const gphmm_batch *batch, double *out, uint64_t *ticket);
int gphmm_wait(gphmm_t *h, uint64_t ticket);

/* Device-resident path (benchmarks, repeated evaluation): prepare uploads and plans a batch once,
 * run_prepared executes only the GPU work (kernels + result download when out != NULL). */
int gphmm_prepare(gphmm_t *h, const gphmm_batch *batch, gphmm_prepared_t **out);
int gphmm_run_prepared(gphmm_t *h, gphmm_prepared_t *p, double *out);
void gphmm_release_prepared(gphmm_t *h, gphmm_prepared_t *p);

/* ---- Region steps: the per-read steps either side of the kernel, on the device ----------------
 * Replaces, for one call over many (region, sample) units,
 *   tools/walkers/haplotypecaller/PairHMMLikelihoodCalculationEngine.java:283-316,361-371  modifyReadQualities
 *       (applyPCRErrorModel with ReadLikelihoodCalculationEngine.findTandemRepeatUnits :193-253, capMinimumReadQualities)
 *   utils/genotyper/AlleleLikelihoods.java:416-458   normalizeLikelihoods(log10globalReadMismappingRate, symmetric)
 *   utils/genotyper/AlleleLikelihoods.java:1351-1376 filterPoorlyModeledEvidence (the keep/drop decision;
 *       thresholds ReadLikelihoodCalculationEngine.java:66-151)
 * The batch carries the reads as modifyReadQualities receives them (soft clips already hard-clipped by the caller,
 * PairHMMLikelihoodCalculationEngine.java:287; ins/del = BI/BD tags or flat Q45, ReadUtils.java:838-862). */
#define GPHMM_RS_DISABLE_CAP_TO_MAPQ 1  /* disableCapReadQualitiesToMapQ */
#define GPHMM_RS_SYMMETRIC_NORMALIZE 2  /* symmetricallyNormalizeAllelesToReference */
#define GPHMM_RS_FILTER_POORLY 4        /* computeReadLikelihoods(..., filterPoorly = true) */
#define GPHMM_RS_DYNAMIC_DISQ 8         /* dynamicDisqualification (DRAGEN-GATK) */

typedef struct gphmm_region_steps {
    int32_t struct_size;                      /* sizeof(gphmm_region_steps) */
    int32_t flags;                            /* GPHMM_RS_* */
    double pcr_rate_factor;                   /* PCRErrorModel.getRateFactor(): 0 NONE, 1 HOSTILE, 2 AGGRESSIVE, 3 CONSERVATIVE */
    int32_t base_quality_score_threshold;     /* baseQualityScoreThreshold, default 18 */
    int32_t reserved;
    double log10_global_read_mismapping_rate; /* default -4.5; must be < 0; -inf: no capping */
    double expected_error_rate_per_base;      /* default 0.02 */
    double read_disqualification_scale;       /* dynamic model only, default 1.0 */
    const uint8_t *mapq;                      /* n_reads mapping qualities (GATKRead.getMappingQuality) */
    const int32_t *ref_hap;                   /* per unit: index of the reference haplotype within the unit, -1 none; NULL = none */
    uint8_t *keep;                            /* out, n_reads: 1 kept, 0 removed as poorly modeled (units must not share reads) */
    uint8_t *hmm_base_q;                      /* out, optional: modified base qualities (HMM_BASE_QUALITIES_TAG), layout of base_q */
    uint8_t *hmm_ins_q, *hmm_del_q;           /* out, optional: the insertion / deletion qualities the kernel used */
    double *raw_lk;                           /* out, optional: the un-normalised likelihoods, layout of gphmm_compute's out
                                                 (PairHMM.getLogLikelihoodArray / --pair-hmm-results-file) */
} gphmm_region_steps;

/* Like gphmm_compute with the steps above fused in.  out[u.out_off + h*nReads + r] is the NORMALISED log10 likelihood,
 * allele-major like AlleleLikelihoods.valuesBySampleIndex[s][a][r] (AlleleLikelihoods.java:71-74). */
int gphmm_compute_regions(gphmm_t *h, const gphmm_batch *batch, const gphmm_region_steps *steps, double *out);
/* Asynchronous form, same queue and ticket space as gphmm_submit: inputs (mapq and ref_hap included) are copied before
 * the call returns; out and the output arrays named in `steps` must stay valid until gphmm_wait(ticket) returns.
 * Queued requests with equal parameters are merged into one GPU batch. */
int gphmm_submit_regions(gphmm_t *h, const gphmm_batch *batch, const gphmm_region_steps *steps, double *out, uint64_t *ticket);

/* ---- PD-HMM: the "partially determined" PairHMM of DRAGEN-GATK mode ------------------------------
 * Replaces PairPDHMMNativeBinding.computeLikelihoods(ReadDataHolder[], HaplotypeDataHolder[] with haplotypePDBases,
 * double[]) as called at utils/pairhmm/VectorLoglessPairPDHMM.java:115; semantics of
 * utils/pairhmm/LoglessPDPairHMM.java:34-153.  hap_pd_bases is parallel to batch->hap_bases (one flag byte per
 * haplotype base, PartiallyDeterminedHaplotype.getAlternateBases(): SNP=1 DEL_START=2 DEL_END=4 A=8 C=16 G=32 T=64).
 * Output layout as gphmm_compute.  The read-span filter of VectorLoglessPairPDHMM.java:129-137 stays with the caller. */
int gphmm_pd_compute(gphmm_t *h, const gphmm_batch *batch, const uint8_t *hap_pd_bases, double *out);

/* ---- Smith-Waterman ---------------------------------------------------------------------------------
 * Batched form of SmithWatermanAligner.align(reference, alternate, SWParameters, SWOverhangStrategy)
 * (utils/smithwaterman/SmithWatermanJavaAligner.java:60-92; native counterpart SWNativeAlignerWrapper.java:33-60 over
 * GKL's SWAlignerNativeBinding).  Pair k aligns alt k to ref k; all pairs share one parameter set.  Results are
 * bit-identical with the Java aligner: offsets[k] = getAlignmentOffset(), elems[k*cigar_capacity .. +n_elems[k]) =
 * the CIGAR, each element (length << 4) | op with op 0 M, 1 I, 2 D, 3 S.  A CIGAR longer than cigar_capacity sets
 * n_elems[k] = -1 and the call returns GPHMM_ERR_TOO_LARGE (all other pairs are valid). */
#define GPHMM_SW_SOFTCLIP 0
#define GPHMM_SW_INDEL 1
#define GPHMM_SW_LEADING_INDEL 2
#define GPHMM_SW_IGNORE 3

typedef struct gphmm_sw_params {
    int32_t struct_size;        /* sizeof(gphmm_sw_params) */
    int32_t match_value;        /* SWParameters.getMatchValue() */
    int32_t mismatch_penalty;   /* getMismatchPenalty(), negative */
    int32_t gap_open_penalty;   /* getGapOpenPenalty(), negative */
    int32_t gap_extend_penalty; /* getGapExtendPenalty(), negative */
    int32_t overhang_strategy;  /* GPHMM_SW_* */
} gphmm_sw_params;

typedef struct gphmm_sw_batch {
    const uint8_t *ref_bases;   /* references, concatenated */
    const int64_t *ref_off;     /* n_pairs + 1 */
    const uint8_t *alt_bases;   /* alternates (reads / haplotypes), concatenated */
    const int64_t *alt_off;     /* n_pairs + 1 */
    int64_t n_pairs;
} gphmm_sw_batch;

int gphmm_sw_align(gphmm_t *h, const gphmm_sw_batch *batch, const gphmm_sw_params *params, int32_t cigar_capacity,
                   int32_t *offsets, int32_t *n_elems, uint32_t *elems);

/* Statistics accumulate over calls until reset. */
int gphmm_get_stats(const gphmm_t *h, gphmm_stats *out);
void gphmm_reset_stats(gphmm_t *h);

/* Host-only: plans the batch exactly as gphmm_compute would (chunking, haplotype sorting, prefix-sharing snapshots,
 * step schedule) without touching a GPU and reports the totals:
 *   out[0] units, out[1] haplotype passes, out[2] schedule segments, out[3] branch-free steps, out[4] checked steps,
 *   out[5] snapshots, out[6] haplotype columns skipped by prefix sharing, out[7] haplotype columns in total,
 *   out[8] chunks, out[9] tasks.  Used by tests and for sizing; needs no device. */
int gphmm_plan_stats(const gphmm_batch *batch, int prefix_sharing, int64_t out[10]);

/* Measurement aid (bench.py "roofline.peak_measured"): runs a dense FFMA kernel with constant-bank operands -- the
 * operand form of the flat-quality forward kernel -- on the handle's first device for about `millis` milliseconds and
 * returns the sustained FP32 rate in TFLOP/s (2 flop per FFMA lane).  Not part of the likelihood path. */
int gphmm_measure_fp32_peak(gphmm_t *h, double millis, double *tflops);

/* Pinned host memory for callers that want zero-copy staging. */
void *gphmm_host_alloc(size_t bytes);
void gphmm_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* GPUPHMM_H */
