/*
 * pairhmm_simd_baseline.c -- a vectorised fp32 CPU PairHMM used ONLY as a reported baseline in bench.py.
 *
 * TEST / BENCHMARK INFRASTRUCTURE, NOT PRODUCT CODE, and NOT the parity oracle (that is pairhmm_oracle.c).
 *
 * Why it exists: the reference's fast CPU path is Intel GKL's AVX-512 float kernel with a double redo of
 * under-flowing pairs (external jar com.intel.gkl:gkl:0.9.0, not in /root/reference, not buildable here).  The
 * scalar fp64 oracle is a poor stand-in for that path's SPEED, so this file provides an honest SIMD number from the
 * same host cores: the same recurrence (LoglessPairHMM.java:47-67) in fp32 with the 2^120 scaling GKL is known to
 * use, 16 reads per vector (one read per lane, all against the same haplotype), OpenMP over (read group,
 * haplotype), and a double-precision redo (phmm_oracle_logless) when the fp32 sum falls under 1e-28.
 * It is checked against the oracle in tests/test_oracle_golden.py::test_simd_baseline_matches_oracle.
 *
 * GCC vector extensions.  The hot function is compiled three times from this file (-DSIMD_KERNEL=..._avx512 with
 * -mavx512f, ..._avx2 with -mavx2, ..._generic) and picked at run time with __builtin_cpu_supports, so one .so runs
 * anywhere (GCC's target_clones does not re-lower 64-byte generic vectors per clone).  See oracle/Makefile.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#if defined(__x86_64__)
#include <xmmintrin.h>
#endif
#ifdef _OPENMP
#include <omp.h>
#endif

int phmm_oracle_logless(const uint8_t *hap, int H, const uint8_t *read, const uint8_t *baseQ, const uint8_t *insQ,
                        const uint8_t *delQ, const uint8_t *gcp, int R, int tristate_off, double *out);
double phmm_oracle_qual_to_error_prob(int qual);
double phmm_oracle_match_to_match_prob(int insQual, int delQual);

#define LANES 16
typedef float v16f __attribute__((vector_size(64), aligned(64)));
typedef int32_t v16i __attribute__((vector_size(64), aligned(64)));

/* One group of up to 16 reads against one haplotype.  Per-row, per-lane coefficient tables are prepared by the
 * caller: tab[row][k][lane], k = tMM tIM tMI tII tMD tDD pMatch pMis, and the read bases rb[row][lane] (0 beyond a
 * lane's read).  len[lane] = read length (0 = unused lane).  sums[lane] receives the raw last-row sum. */
#ifdef SIMD_KERNEL
void SIMD_KERNEL(const uint8_t *hap, int H, int Rmax, const float *tab, const int32_t *rb, const int32_t *len,
                         float *work, float *sums)
{
    v16f *Mp = (v16f *)work, *Ip = Mp + (H + 1), *Dp = Ip + (H + 1), *Mc = Dp + (H + 1), *Ic = Mc + (H + 1), *Dc = Ic + (H + 1);
    const float init = ldexpf(1.0f, 120) / (float)H;
    v16f zero, vinit;
    zero = (v16f){0.f} * 0.f; vinit = zero + init;
    for (int j = 0; j <= H; j++) { Mp[j] = zero; Ip[j] = zero; Dp[j] = vinit; }
    v16i vlen, nbase;
    for (int l = 0; l < LANES; l++) vlen[l] = len[l];
    nbase = (v16i){0} * 0 + (int32_t)'N';
    v16f result = zero;
    for (int i = 1; i <= Rmax; i++) {
        const v16f *t = (const v16f *)(tab + (size_t)(i - 1) * 8 * LANES);
        const v16f tMM = t[0], tIM = t[1], tMI = t[2], tII = t[3], tMD = t[4], tDD = t[5], pM = t[6], pX = t[7];
        const v16i x = *(const v16i *)(rb + (size_t)(i - 1) * LANES);
        const v16i x_is_n = x == nbase;
        Mc[0] = zero; Ic[0] = zero; Dc[0] = zero;
        v16f mleft = zero, dleft = zero, rowsum = zero;
        for (int j = 1; j <= H; j++) {
            const v16i y = (v16i){0} * 0 + (int32_t)hap[j - 1];
            const v16i match = (x == y) | x_is_n | (y == nbase);
            const v16f prior = (v16f)(((v16i)pM & match) | ((v16i)pX & ~match));
            const v16f m = prior * (Mp[j - 1] * tMM + (Ip[j - 1] + Dp[j - 1]) * tIM);
            const v16f ins = Mp[j] * tMI + Ip[j] * tII;
            const v16f del = mleft * tMD + dleft * tDD;
            Mc[j] = m; Ic[j] = ins; Dc[j] = del;
            mleft = m; dleft = del;
            rowsum += m + ins;
        }
        {   /* lanes whose read ends on this row take their sum */
            const v16i done = vlen == ((v16i){0} * 0 + i);
            result = (v16f)(((v16i)rowsum & done) | ((v16i)result & ~done));
        }
        v16f *s;
        s = Mp; Mp = Mc; Mc = s; s = Ip; Ip = Ic; Ic = s; s = Dp; Dp = Dc; Dc = s;
    }
    for (int l = 0; l < LANES; l++) sums[l] = result[l];
}
#else  /* dispatcher + drivers */

typedef void (*group_fn)(const uint8_t *, int, int, const float *, const int32_t *, const int32_t *, float *, float *);
void phmm_simd_group_avx512(const uint8_t *, int, int, const float *, const int32_t *, const int32_t *, float *, float *);
void phmm_simd_group_avx2(const uint8_t *, int, int, const float *, const int32_t *, const int32_t *, float *, float *);
void phmm_simd_group_generic(const uint8_t *, int, int, const float *, const int32_t *, const int32_t *, float *, float *);

static group_fn pick_kernel(void) {
#if defined(__x86_64__)
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512dq")) return phmm_simd_group_avx512;
    if (__builtin_cpu_supports("avx2") && __builtin_cpu_supports("fma")) return phmm_simd_group_avx2;
#endif
    return phmm_simd_group_generic;
}

/* 512 / 256 / 0: which vector ISA the baseline runs with on this host */
int phmm_simd_isa(void) {
    group_fn f = pick_kernel();
    return f == phmm_simd_group_avx512 ? 512 : f == phmm_simd_group_avx2 ? 256 : 0;
}

/* One (region, sample) unit, flat SoA like phmm_oracle_unit; out[r*nHaps + h] = log10 likelihood.
 * Returns the number of pairs redone in double. */
long phmm_simd_unit(const uint8_t *read_bases, const uint8_t *base_q, const uint8_t *ins_q, const uint8_t *del_q,
                    const uint8_t *gcp, const int32_t *read_off, int n_reads, const uint8_t *hap_bases,
                    const int32_t *hap_off, int n_haps, int threads, double *out)
{
    if (n_reads == 0 || n_haps == 0) return 0;
#if defined(__x86_64__)
    /* flush-to-zero + denormals-are-zero, as GKL does: most of the DP matrix under-flows in fp32 and denormal
     * arithmetic would cost ~100 cycles per operation */
    const unsigned int saved_csr = _mm_getcsr();
    _mm_setcsr(saved_csr | 0x8040);
#endif
    /* reads in groups of 16 by descending length (similar lengths share a vector) */
    int *order = (int *)malloc(sizeof(int) * n_reads);
    for (int r = 0; r < n_reads; r++) order[r] = r;
    for (int a = 1; a < n_reads; a++) {  /* insertion sort: units hold ~100 reads */
        int v = order[a], la = read_off[v + 1] - read_off[v], b = a - 1;
        while (b >= 0 && read_off[order[b] + 1] - read_off[order[b]] < la) { order[b + 1] = order[b]; b--; }
        order[b + 1] = v;
    }
    const int n_groups = (n_reads + LANES - 1) / LANES;
    int maxH = 1;
    for (int h = 0; h < n_haps; h++) if (hap_off[h + 1] - hap_off[h] > maxH) maxH = hap_off[h + 1] - hap_off[h];
    long rescued = 0;
    const group_fn group_vs_hap = pick_kernel();
    const double log10_init = 120.0 * log10(2.0);
    (void)threads;
    {
        float *tab = NULL, *work = NULL;
        int32_t *rb = NULL;
        int tab_rows = 0;
        work = (float *)aligned_alloc(64, sizeof(float) * LANES * 6 * (size_t)(maxH + 1));
        for (int g = 0; g < n_groups; g++) {
            int32_t len[LANES];
            int Rmax = 0;
            for (int l = 0; l < LANES; l++) {
                const int idx = g * LANES + l;
                len[l] = idx < n_reads ? read_off[order[idx] + 1] - read_off[order[idx]] : 0;
                if (len[l] > Rmax) Rmax = len[l];
            }
            if (Rmax > tab_rows) {
                free(tab); free(rb);
                tab_rows = Rmax;
                tab = (float *)aligned_alloc(64, sizeof(float) * 8 * LANES * (size_t)tab_rows);
                rb = (int32_t *)aligned_alloc(64, sizeof(int32_t) * LANES * (size_t)tab_rows);
            }
            for (int i = 0; i < Rmax; i++)
                for (int l = 0; l < LANES; l++) {
                    float *t = tab + (size_t)i * 8 * LANES;
                    const int idx = g * LANES + l;
                    if (i < len[l]) {
                        const int o = read_off[order[idx]] + i;
                        const double ei = phmm_oracle_qual_to_error_prob(ins_q[o]), ed = phmm_oracle_qual_to_error_prob(del_q[o]);
                        const double ec = phmm_oracle_qual_to_error_prob(gcp[o]), e = phmm_oracle_qual_to_error_prob(base_q[o]);
                        t[0 * LANES + l] = (float)phmm_oracle_match_to_match_prob(ins_q[o], del_q[o]);
                        t[1 * LANES + l] = (float)(1.0 - ec); t[2 * LANES + l] = (float)ei; t[3 * LANES + l] = (float)ec;
                        t[4 * LANES + l] = (float)ed; t[5 * LANES + l] = (float)ec;
                        t[6 * LANES + l] = (float)(1.0 - e); t[7 * LANES + l] = (float)(e / 3.0);
                        rb[(size_t)i * LANES + l] = read_bases[o];
                    } else {
                        for (int k = 0; k < 8; k++) t[k * LANES + l] = 0.f;
                        rb[(size_t)i * LANES + l] = 0;
                    }
                }
            for (int h = 0; h < n_haps; h++) {
                const int H = hap_off[h + 1] - hap_off[h];
                float sums[LANES];
                group_vs_hap(hap_bases + hap_off[h], H, Rmax, tab, rb, len, work, sums);
                for (int l = 0; l < LANES; l++) {
                    const int idx = g * LANES + l;
                    if (idx >= n_reads) continue;
                    const int r = order[idx];
                    double v;
                    if (!(sums[l] >= 1e-28f) || len[l] == 0) {  /* GKL-style float -> double redo */
                        const int o = read_off[r];
                        phmm_oracle_logless(hap_bases + hap_off[h], H, read_bases + o, base_q + o, ins_q + o, del_q + o, gcp + o,
                                            len[l], 0, &v);
                        rescued++;
                    } else {
                        v = log10((double)sums[l]) - log10_init;
                    }
                    out[(size_t)r * n_haps + h] = v;
                }
            }
        }
        free(tab); free(rb); free(work);
    }
    free(order);
#if defined(__x86_64__)
    _mm_setcsr(saved_csr);
#endif
    return rescued;
}

/* A whole batch in the layout of include/gpuphmm.h (int64 offsets, units = {read_begin, read_end, hap_begin, hap_end,
 * out_off}); OpenMP over units.  Returns the number of pairs redone in double. */
long phmm_simd_batch(const uint8_t *read_bases, const uint8_t *base_q, const uint8_t *ins_q, const uint8_t *del_q,
                     const uint8_t *gcp, const int64_t *read_off, const uint8_t *hap_bases, const int64_t *hap_off,
                     const int64_t *units, long n_units, int threads, double *out)
{
    long rescued = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads > 1 ? threads : 1) reduction(+ : rescued)
#endif
    for (long u = 0; u < n_units; u++) {
        const int64_t r0 = units[5 * u], r1 = units[5 * u + 1], h0 = units[5 * u + 2], h1 = units[5 * u + 3], o = units[5 * u + 4];
        const int nr = (int)(r1 - r0), nh = (int)(h1 - h0);
        if (nr == 0 || nh == 0) continue;
        int32_t *ro = (int32_t *)malloc(sizeof(int32_t) * (nr + 1)), *ho = (int32_t *)malloc(sizeof(int32_t) * (nh + 1));
        const int64_t rb = read_off[r0], hb = hap_off[h0];
        for (int k = 0; k <= nr; k++) ro[k] = (int32_t)(read_off[r0 + k] - rb);
        for (int k = 0; k <= nh; k++) ho[k] = (int32_t)(hap_off[h0 + k] - hb);
        rescued += phmm_simd_unit(read_bases + rb, base_q + rb, ins_q + rb, del_q + rb, gcp + rb, ro, nr, hap_bases + hb, ho, nh, 1, out + o);
        free(ro); free(ho);
    }
    return rescued;
}
#endif /* SIMD_KERNEL */
