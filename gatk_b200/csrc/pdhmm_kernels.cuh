// pdhmm_kernels.cuh -- the "partially determined" PairHMM of DRAGEN-GATK mode on the device.
//
// Reference: utils/pairhmm/LoglessPDPairHMM.java:34-153 (recurrence and state machine), :170-204 (priors with the
// SNP mask of a column), utils/haplotype/PartiallyDeterminedHaplotype.java:59-65 (flag bits); native counterpart of
// the reference: VectorLoglessPairPDHMM.java:71-147 (GKL PDHMM binding).
//
// Same wavefront as phmm_forward_kernel (one warp per (read, haplotype) pair, lane l owns K consecutive read rows,
// lane l works on column step-l, last row handed down by shuffles, longer reads in strips through a boundary buffer),
// but with the reference's UNSCALED M/I/D plus the three "branch" values per row, so that the max() merges of the
// AFTER_DEL state act on exactly the quantities the Java code compares.  max(a x, a y) = a max(x, y) for a > 0, so the
// result is still homogeneous in the initial value and the fp32 pass may start from 2^(125 - ceil log2 H);
// pairs whose fp32 sum falls below 1e-28 are redone by the same kernel in double from Java's own 2^1020 / H.
#pragma once
#include <stdint.h>
#include "phmm_kernels.cuh"

namespace phmm_dev {

// per haplotype column: bits 0-3 = alternative bases allowed by a SNP flag (A, C, G, T), bit 4 = DEL_END on this
// column, bits 5-6 = state in which row 1 processes the column (0 NORMAL, 1 INSIDE_DEL, 2 AFTER_DEL), bit 7 = SNP flag
constexpr uint32_t PD_MASK_BITS = 0x0f, PD_DEL_END_BIT = 0x10, PD_TYPE_SHIFT = 5, PD_TYPE_BITS = 3, PD_SNP_BIT = 0x80;
constexpr uint32_t PD_NORMAL = 0, PD_INSIDE_DEL = 1, PD_AFTER_DEL = 2;

struct PdTask {
    uint32_t read;         // chunk-local read
    uint32_t hap_off, H;   // chunk-local haplotype columns
    uint32_t out_slot;     // chunk-local output slot
    // LoglessPDPairHMM keeps its state variable across rows (:59): rows >= 2 start in the state the previous row ended
    // in (`carry`) and keep it up to and including the first flagged column (`first_event`, H + 1 if none)
    uint32_t first_event, carry;
    int32_t c0_exp;        // fp32 pass: exponent of the initial value
    uint32_t pad;
};

struct PdArgs {
    const uint8_t *rd_bases, *rd_q, *rd_i, *rd_d, *rd_c;
    const uint32_t *read_off;
    const uint8_t *hap_bases, *hap_flags;
    const PdTask *tasks;
    const uint32_t *task_index;   // fp64 redo: indices into tasks (nullptr: tasks [first, first + n) directly)
    const uint32_t *n_tasks_ptr;  // device-side count (redo list) or nullptr
    uint32_t first, n_tasks;
    uint32_t *counter;
    void *sums;                   // float (per task) or double (per redo slot)
    void *bnd;
    uint32_t bnd_stride;
    const double *m2m;
    int *err;                     // 1: quality out of range, 2: read base other than ACGT on a SNP column (:202)
    int32_t tristate_off;
};

template <typename T> struct __align__(16) BndPD { T m, i, d, bm, bi, bd, pad0, pad1; };

template <typename T> __device__ __forceinline__ T pd_max(T a, T b) { return a > b ? a : b; }

template <typename T, int K>
__global__ void __launch_bounds__(32) phmm_pd_kernel(const PdArgs g)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int ROWS = 32 * K;
    const int lane = threadIdx.x;
    T *const sums = reinterpret_cast<T *>(g.sums);
    BndPD<T> *const bnd = reinterpret_cast<BndPD<T> *>(g.bnd) + (size_t)blockIdx.x * g.bnd_stride;
    const uint32_t n_tasks = g.n_tasks_ptr ? *g.n_tasks_ptr : g.n_tasks;

    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(g.counter, 1u);
        ti = __shfl_sync(FULL, ti, 0);
        if (ti >= n_tasks) break;
        const PdTask t = g.tasks[g.task_index ? g.task_index[ti] : g.first + ti];
        const uint32_t ro = g.read_off[t.read];
        const int R = (int)(g.read_off[t.read + 1] - ro), H = (int)t.H;
        const uint8_t *__restrict__ hap = g.hap_bases + t.hap_off;
        const uint8_t *__restrict__ flg = g.hap_flags + t.hap_off;
        // fp64: the reference's own initial value 2^1020 / H (LoglessPDPairHMM.java:11,45); fp32: a power of two
        const T c0 = sizeof(T) == 8 ? (T)(0x1p1020 / (double)H) : (T)scalbn(1.0, t.c0_exp);
        const int n_strips = (R + ROWS - 1) / ROWS;
        const int row_lane = R > 0 ? ((R - 1) % ROWS) / K : -1, row_k = R > 0 ? (R - 1) % K : -1;  // where row R lives in the last strip
        T sum = (T)0;

        for (int strip = 0; strip < n_strips; ++strip) {
            const bool first_strip = strip == 0, last_strip = strip == n_strips - 1;
            // ---- per-row constants (PairHMMModel.java:107-117, LoglessPDPairHMM.java:170-181) ----
            T tMM[K], tIM[K], tMI[K], tII[K], tMD[K], pm[K], px[K];
            uint32_t xb[K], abit[K];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = strip * ROWS + lane * K + k + 1;
                tMM[k] = tIM[k] = tMI[k] = tII[k] = tMD[k] = pm[k] = px[k] = (T)0;
                xb[k] = 0x100u; abit[k] = 0;  // 0x100 never equals a haplotype byte
                if (i <= R) {
                    uint32_t q = g.rd_q[ro + i - 1], qi = g.rd_i[ro + i - 1], qd = g.rd_d[ro + i - 1], qc = g.rd_c[ro + i - 1];
                    if (q > (uint32_t)MAX_QUAL || qi > 127u || qd > 127u || qc > 127u) {
                        atomicExch(g.err, 1);
                        q = min(q, (uint32_t)MAX_QUAL); qi = min(qi, 127u); qd = min(qd, 127u); qc = min(qc, 127u);
                    }
                    const uint32_t mn = min(qi, qd), mx = max(qi, qd);
                    const double ec = c_eps[qc], e = c_eps[q];
                    tMM[k] = (T)__ldg(g.m2m + ((mx * (mx + 1)) >> 1) + mn);
                    tIM[k] = (T)(1.0 - ec); tMI[k] = (T)c_eps[qi]; tII[k] = (T)ec; tMD[k] = (T)c_eps[qd];
                    pm[k] = (T)(1.0 - e); px[k] = (T)(g.tristate_off ? e : e / 3.0);
                    const uint32_t x = g.rd_bases[ro + i - 1];
                    xb[k] = x;
                    switch (x) {  // LoglessPDPairHMM.isBasePDMatching :184-204
                        case 'A': case 'a': abit[k] = 1; break;
                        case 'C': case 'c': abit[k] = 2; break;
                        case 'G': case 'g': abit[k] = 4; break;
                        case 'T': case 't': abit[k] = 8; break;
                        case 'N': abit[k] = 0; break;   // matches before the mask is looked at
                        default: abit[k] = 0x80000000u; // the reference throws when such a base meets a SNP column
                    }
                }
            }
            T M[K], I[K], D[K], bM[K], bI[K], bD[K];
#pragma unroll
            for (int k = 0; k < K; ++k) M[k] = I[k] = D[k] = bM[k] = bI[k] = bD[k] = (T)0;
            // row above at the previous column; D[0][0] is the initial value, branch row 0 is Java's 0.0
            T dgm = (T)0, dgi = (T)0, dgd = (first_strip && lane == 0) ? c0 : (T)0, dgbm = (T)0, dgbi = (T)0, dgbd = (T)0;
            int p = 1 - lane;  // 1-based column of this lane in the current step
            const int n_steps = H + 31;
            for (int s = 1; s <= n_steps; ++s, ++p) {
                const bool valid = p >= 1 && p <= H;
                uint32_t hb = 0x200u, fl = 0;
                if (valid) { hb = hap[p - 1]; fl = flg[p - 1]; }
                // row above at this column: last row of the lane above (its state before this step), strip boundary, or row 0
                T mu = __shfl_up_sync(FULL, M[K - 1], 1), iu = __shfl_up_sync(FULL, I[K - 1], 1), du = __shfl_up_sync(FULL, D[K - 1], 1);
                T bmu = __shfl_up_sync(FULL, bM[K - 1], 1), biu = __shfl_up_sync(FULL, bI[K - 1], 1), bdu = __shfl_up_sync(FULL, bD[K - 1], 1);
                if (lane == 0) {
                    mu = iu = bmu = biu = bdu = (T)0;
                    du = first_strip ? c0 : (T)0;
                    if (!first_strip && valid) {
                        const BndPD<T> b = bnd[p];
                        mu = b.m; iu = b.i; du = b.d; bmu = b.bm; biu = b.bi; bdu = b.bd;
                    }
                }
                const uint32_t mask = fl & PD_MASK_BITS;
                const bool del_end = (fl & PD_DEL_END_BIT) != 0;
                const uint32_t t_row1 = (fl >> PD_TYPE_SHIFT) & PD_TYPE_BITS;
                uint32_t t_rest = t_row1;
                if (valid && (uint32_t)p <= t.first_event)
                    t_rest = t.carry == PD_INSIDE_DEL ? PD_INSIDE_DEL : (t.carry == PD_AFTER_DEL && p == 1 ? PD_AFTER_DEL : PD_NORMAL);

                T Mn[K], Dn[K], In[K], nbM[K], nbI[K], nbD[K];
                const bool row1_lane = first_strip && lane == 0;
                // Most columns carry no flag: when every lane of the warp is on such a column (and row 1 agrees with the
                // other rows), the update is the plain recurrence plus "branch = left neighbour".
                const bool plain_col = !del_end && t_rest == PD_NORMAL && (!row1_lane || t_row1 == PD_NORMAL);
                if (__all_sync(FULL, plain_col)) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const T am = k ? M[k - 1] : dgm, ai = k ? I[k - 1] : dgi, ad = k ? D[k - 1] : dgd;
                        bool match = xb[k] == hb || xb[k] == (uint32_t)'N' || hb == (uint32_t)'N';
                        if (!match && (fl & PD_SNP_BIT)) {
                            if (abit[k] & 0x80000000u) atomicExch(g.err, 2);
                            match = (mask & abit[k]) != 0;
                        }
                        const T prior = valid ? (match ? pm[k] : px[k]) : (T)0;
                        nbM[k] = M[k]; nbI[k] = I[k]; nbD[k] = D[k];
                        Mn[k] = prior * (am * tMM[k] + ai * tIM[k] + ad * tIM[k]);
                        Dn[k] = M[k] * tMD[k] + D[k] * tII[k];
                    }
                    In[0] = mu * tMI[0] + iu * tII[0];
#pragma unroll
                    for (int k = 1; k < K; ++k) In[k] = Mn[k - 1] * tMI[k] + In[k - 1] * tII[k];
                } else {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const uint32_t ty = (row1_lane && k == 0) ? t_row1 : t_rest;
                        const bool inside = ty == PD_INSIDE_DEL, after = ty == PD_AFTER_DEL;
                        // row above, previous column
                        T am = k ? M[k - 1] : dgm, ai = k ? I[k - 1] : dgi, ad = k ? D[k - 1] : dgd;
                        const T abm = k ? bM[k - 1] : dgbm, abi = k ? bI[k - 1] : dgbi, abd = k ? bD[k - 1] : dgbd;
                        // branch values: copy from the left (NORMAL), hold (INSIDE_DEL) or merge (AFTER_DEL); selects, no branches:
                        // the lanes of a warp sit on different columns
                        const T xm = pd_max(bM[k], M[k]), xi = pd_max(bI[k], I[k]), xd = pd_max(bD[k], D[k]);
                        nbM[k] = inside ? bM[k] : (after ? xm : M[k]);
                        nbI[k] = inside ? bI[k] : (after ? xi : I[k]);
                        nbD[k] = inside ? bD[k] : (after ? xd : D[k]);
                        am = after ? pd_max(abm, am) : am; ai = after ? pd_max(abi, ai) : ai; ad = after ? pd_max(abd, ad) : ad;
                        const T lm = after ? xm : M[k], ld = after ? xd : D[k];  // this row, previous column
                        bool match = xb[k] == hb || xb[k] == (uint32_t)'N' || hb == (uint32_t)'N';
                        if (!match && (fl & PD_SNP_BIT)) {
                            if (abit[k] & 0x80000000u) atomicExch(g.err, 2);
                            match = (mask & abit[k]) != 0;
                        }
                        const T prior = valid ? (match ? pm[k] : px[k]) : (T)0;
                        Mn[k] = prior * (am * tMM[k] + ai * tIM[k] + ad * tIM[k]);
                        Dn[k] = lm * tMD[k] + ld * tII[k];  // deletionToDeletion = insertionToInsertion = eps(gcp)
                    }
                    // insertion: row above at THIS column -- chain down the lane's rows
                    T um = mu, ui = iu, ubm = bmu, ubi = biu;
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const T xm = del_end ? pd_max(ubm, um) : um, xi = del_end ? pd_max(ubi, ui) : ui;
                        In[k] = xm * tMI[k] + xi * tII[k];
                        um = Mn[k]; ui = In[k]; ubm = nbM[k]; ubi = nbI[k];
                    }
                }
#pragma unroll
                for (int k = 0; k < K; ++k) { M[k] = Mn[k]; I[k] = In[k]; D[k] = Dn[k]; bM[k] = nbM[k]; bI[k] = nbI[k]; bD[k] = nbD[k]; }
                dgm = mu; dgi = iu; dgd = du; dgbm = bmu; dgbi = biu; dgbd = bdu;
                if (last_strip && lane == row_lane && valid) {
#pragma unroll
                    for (int k = 0; k < K; ++k)
                        if (k == row_k) sum += M[k] + I[k];  // LoglessPDPairHMM.java:149-152
                }
                if (!last_strip && lane == 31 && valid) {
                    BndPD<T> o;
                    o.m = M[K - 1]; o.i = I[K - 1]; o.d = D[K - 1]; o.bm = bM[K - 1]; o.bi = bI[K - 1]; o.bd = bD[K - 1];
                    o.pad0 = o.pad1 = (T)0;
                    bnd[p] = o;
                }
            }
        }
        sum = __shfl_sync(FULL, sum, row_lane < 0 ? 0 : row_lane);
        if (lane == 0) sums[ti] = sum;
    }
}

// ---------------------------------------------------------------------------------------------
// Fast fp32 PD-HMM kernel for reads of up to 254 bases (one strip).  Flags are sparse: a partially determined
// haplotype carries a handful of SNP / deletion events, so for most steps every lane of the warp sits on a column
// whose update is the plain LoglessPairHMM recurrence.  Those steps run fast_step<K, false> of phmm_kernels.cuh --
// the scaled 5-instruction recurrence with rotation hand-off, accumulator row and shared-memory prior table -- and
// do not touch the branch values at all.  The host marks, per haplotype, the steps in which some lane is on (or one
// column before) a column that takes part in the NORMAL / INSIDE_DEL / AFTER_DEL state machine; only those run the
// slow step below, which keeps the branch values and merges them with max() exactly where the reference does
// (LoglessPDPairHMM.java:62-141).  The states are scaled per row by positive constants (I~ = I / tMI_i,
// D~ = D / tMD_i), which commutes with max(): every comparison has the outcome it has in the reference.
//   * SNP columns are plain columns with a wider match test: the column CODE (haplotype byte + alternative-base
//     mask) indexes the prior table, so isBasePDMatching (:184-204) costs nothing per cell.
//   * the row below the read accumulates sum_j (M + I)[R][j] (:149-152): no per-step work for the result.
// ---------------------------------------------------------------------------------------------
constexpr int PD_MAX_CODES = 64;   // prior-table rows per chunk

constexpr int PD_MAX_SNP_CODES = 8;  // distinct (haplotype byte, alternative-base mask) pairs on the SNP columns of ONE haplotype (fast kernels)
struct PdHap {
    uint32_t code_off;             // first column of the haplotype in the padded code / flag streams
    uint32_t seg_first, n_segs;    // schedule: into PdFastArgs::segs, (fast steps, slow steps) pairs covering H + 33 steps
    uint32_t ev_first, n_events;   // SIMPLE kernels: deletion events (a, b) of the haplotype, into PdFastArgs::events
    uint32_t n_snp;                // SNP column codes of this haplotype: prior-table rows n_plain + 3 + s
    uint16_t snp[PD_MAX_SNP_CODES];  // haplotype byte | mask << 8 (mask: alternative-base bits A, C, G, T = 1, 2, 4, 8)
    uint32_t pad0, pad1;
};

// one fast-kernel task: a read against consecutive pairs of its bucket (the haplotypes of its unit that run this kernel)
struct PdGroup {
    uint32_t read;        // chunk-local read
    uint32_t task_first;  // first pair, relative to PdFastArgs::first
    uint32_t n;           // pairs
    uint32_t pad;
};

struct PdFastArgs {
    const uint8_t *rd_bases, *rd_q, *rd_i, *rd_d, *rd_c;
    const uint32_t *read_off;
    const uint8_t *codes, *flags;  // per column, STREAM_PAD zero columns in front of every haplotype and 2 * STREAM_PAD behind
    const PdTask *tasks;           // PdTask::pad = index into haps
    const PdGroup *groups;
    uint32_t n_groups;
    const PdHap *haps;
    const uint2 *segs;
    const uint2 *events;           // SIMPLE kernels: (a, b) = first and last column of a deletion, 1-based, sorted by column
    uint32_t first, n_tasks;
    uint32_t *counter;
    float *sums;                   // per task (same index as tasks)
    const double *m2m;
    int *err;
    int32_t tristate_off;
    int32_t n_plain;                   // prior-table rows: 0 = outside the haplotype, 1 .. n_plain = haplotype bytes without a SNP flag,
    int32_t n_rows;                    // n_plain + 1 / + 2 = match / mismatch prior of every row (source of the SNP rows), then the SNP rows
    uint8_t code_byte[PD_MAX_CODES];   // plain code -> haplotype byte
};

//
// SIMPLE = true (round 2, second form of the slow window).  For a haplotype whose deletions are well formed and apart
// (state NORMAL before the first flagged column a of an event, DEL_END only on its last column b, the next event starting
// at b + 3 or later, state NORMAL again at the end of the haplotype so that every row sees the same column states), the
// state machine reduces to three moments per lane and event:
//   column a   (before the update): the branch matrices hold column a - 1 from here to the end of the deletion
//              (:69-71 copy from the left in NORMAL, :90-92 hold in INSIDE_DEL)  ->  the lane SAVES its M, I~, D~ registers
//              in a shared-memory record (zero for rows that are not read rows, so that a max() with them changes nothing);
//   column b   (DEL_END, after the update): the insertion chain takes max(branch, value) of the row above (:80-82) ->
//              the lane recomputes its K insertion values from the saved record;
//   column b+1 (AFTER_DEL, before the update): every use of column b is max(branch, value) (:107-118) -> the lane takes
//              the max IN PLACE and runs the plain step.  Only the accumulator row must not see the merged values (the sum
//              :149-152 adds the stored M + I of column b): its input is computed before the merge and put back after it.
// The row above a lane's first row belongs to the previous lane, which is one column ahead: its saved last row is read
// from its record (row 0 above lane 0 has branch values 0).  Every other lane runs the plain fast step in the same
// instruction stream; the three moments are divergent branches with one active lane each.
template <int K> __host__ __device__ constexpr int pd_save_vecs() { return (3 * K + 3) / 4 + 4; }  // record of a lane: 3K floats; by event parity: (M, I~, D~) of its last row, its unmerged (M, I~) at column b
template <int K> constexpr size_t pd_fast_smem(int n_codes, bool simple) {
    return (size_t)n_codes * ((K + 3) / 4) * 512 + (simple ? (size_t)pd_save_vecs<K>() * 512 : 0);
}

__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 lds128m(uint32_t addr) {  // like lds128, ordered against the stores above
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

template <int K, bool SIMPLE>
__global__ void __launch_bounds__(32, K > 5 ? 16 : 20) phmm_pd_fast_kernel(const PdFastArgs g)
{
    constexpr int NV = (K + 3) / 4;
    constexpr int NS = (3 * K + 3) / 4;  // float4s of a saved record
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tab_s = reinterpret_cast<float *>(smem_raw);
    int lane, src_lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("{ .reg .u32 t; add.u32 t, %1, 31; and.b32 %0, t, 31; }" : "=r"(src_lane) : "r"(lane));
    const uint32_t tab_lane = (uint32_t)__cvta_generic_to_shared(smem_raw) + lane * 16;
    const ptrdiff_t flag_delta = g.flags - g.codes;

    for (;;) {
        uint32_t gi = 0;
        if (lane == 0) gi = atomicAdd(g.counter, 1u);
        gi = __shfl_sync(FULL, gi, 0);
        if (gi >= g.n_groups) break;
        const PdGroup gr = g.groups[gi];
        const uint32_t ro = g.read_off[gr.read];
        const int R = (int)(g.read_off[gr.read + 1] - ro);  // host guarantees R + 2 <= 32 * K
        const int acc_lane = R / K, acc_slot = R % K;  // accumulator row = 0-based row R

        // ---- once per read: transition coefficients, the static rows of the prior table ----
        float cb[K], cc[K], cg[K], cd[K];
        uint32_t real_mask = 0;  // bit k: slot k holds a read row (only those take part in the state machine)
        uint32_t xs[NV];         // the read bases of this lane's rows, one byte each (SNP rows are built per haplotype)
#pragma unroll
        for (int v = 0; v < NV; ++v) xs[v] = 0;
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = lane * K + k + 1;  // 1-based read row
            double A = 0.0, B = 0.0, C = 0.0, G = 0.0, DD = 0.0, pm = 0.0, px = 0.0;
            uint32_t x = 0;
            const bool real = i <= R;
            if (real) {
                real_mask |= 1u << k;
                uint32_t q = g.rd_q[ro + i - 1], qi = g.rd_i[ro + i - 1], qd = g.rd_d[ro + i - 1], qc = g.rd_c[ro + i - 1];
                x = g.rd_bases[ro + i - 1];
                xs[k / 4] |= x << (8 * (k % 4));
                if (q > (uint32_t)MAX_QUAL || qi > 127u || qd > 127u || qc > 127u) {
                    atomicExch(g.err, 1);
                    q = min(q, (uint32_t)MAX_QUAL); qi = min(qi, 127u); qd = min(qd, 127u); qc = min(qc, 127u);
                }
                const double ei = c_eps[qi], ec = c_eps[qc];
                const uint32_t mn = min(qi, qd), mx = max(qi, qd);
                const double tIM = 1.0 - ec;
                A = __ldg(g.m2m + ((mx * (mx + 1)) >> 1) + mn);
                if (i > 1) {
                    const double tmi_prev = c_eps[min((uint32_t)g.rd_i[ro + i - 2], 127u)];
                    const double tmd_prev = c_eps[min((uint32_t)g.rd_d[ro + i - 2], 127u)];
                    B = tIM * tmi_prev;
                    C = tIM * tmd_prev;
                    G = ec * tmi_prev / ei;
                } else {
                    C = tIM;  // row 0: D~ = c0 (tMD_0 = 1); I~ of row 0 is 0, so b = g = 0 and the rotated-in value is ignored
                }
                DD = ec;
                const double e = c_eps[q];
                // tMM goes into the priors and divides b and c; tMM = 0 cannot be factored out: NaN -> fp64 redo
                const double inv = A > 0.0 ? 1.0 / A : __longlong_as_double(0x7ff8000000000000LL);
                pm = (1.0 - e) * A;
                px = (g.tristate_off ? e : e / 3.0) * A;
                B *= inv; C *= inv;
            } else if (i == R + 1) {  // accumulator row: M_acc = 1 * (M_R + tMI_R * I~_R) = (M + I)[R]
                B = R >= 1 ? c_eps[min((uint32_t)g.rd_i[ro + R - 1], 127u)] : 0.0;
                DD = 1.0;
            }
            if (lane == 31 && k == K - 1) DD = 1.0;  // carrier of the virtual row 0: keeps D~ = c0
            cb[k] = (float)B; cc[k] = (float)C; cg[k] = (float)G; cd[k] = (float)DD;
            const float pmf = (float)pm, pxf = (float)px;
            const float other = (!real && i == R + 1) ? 1.f : 0.f;  // the accumulator row passes everything, pads nothing
            tab_s[((0 * NV + k / 4) * 32 + lane) * 4 + (k % 4)] = other;  // outside the haplotype (its M input is 0 there)
            for (int y = 1; y <= g.n_plain; ++y) {
                const uint32_t hb = g.code_byte[y];
                const bool match = x == hb || x == (uint32_t)'N' || hb == (uint32_t)'N';  // LoglessPDPairHMM.java:176-180
                tab_s[((y * NV + k / 4) * 32 + lane) * 4 + (k % 4)] = real ? (match ? pmf : pxf) : other;
            }
            tab_s[(((g.n_plain + 1) * NV + k / 4) * 32 + lane) * 4 + (k % 4)] = real ? pmf : other;
            tab_s[(((g.n_plain + 2) * NV + k / 4) * 32 + lane) * 4 + (k % 4)] = real ? pxf : other;
        }

        for (uint32_t hk = 0; hk < gr.n; ++hk) {
        const uint32_t ti = gr.task_first + hk;
        const PdTask t = g.tasks[g.first + ti];
        const PdHap *const hpp = g.haps + t.pad;
        const PdHap hp = *hpp;
        const int H = (int)t.H;
        const float c0 = (float)scalbn(1.0, t.c0_exp);
        // ---- once per haplotype: the prior rows of its SNP columns (isBasePDMatching :184-204) ----
        __syncwarp();  // the previous sweep is through with them
        if (hp.n_snp) {
            float pmv[NV * 4], pxv[NV * 4];
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const float4 a4 = lds128m(tab_lane + ((g.n_plain + 1) * NV + v) * 512), b4 = lds128m(tab_lane + ((g.n_plain + 2) * NV + v) * 512);
                pmv[4 * v] = a4.x; pmv[4 * v + 1] = a4.y; pmv[4 * v + 2] = a4.z; pmv[4 * v + 3] = a4.w;
                pxv[4 * v] = b4.x; pxv[4 * v + 1] = b4.y; pxv[4 * v + 2] = b4.z; pxv[4 * v + 3] = b4.w;
            }
            for (uint32_t sn = 0; sn < hp.n_snp; ++sn) {
                const uint32_t sc = hpp->snp[sn], hb = sc & 0xffu, cm = sc >> 8;
                float v4[NV * 4];
#pragma unroll
                for (int k = 0; k < NV * 4; ++k) {
                    v4[k] = 0.f;
                    if (k < K) {
                        const uint32_t x = (xs[k / 4] >> (8 * (k % 4))) & 0xffu;
                        bool match = x == hb || x == (uint32_t)'N' || hb == (uint32_t)'N';
                        if (!match && ((real_mask >> k) & 1u)) {
                            uint32_t abit;
                            switch (x) {
                                case 'A': case 'a': abit = 1; break;
                                case 'C': case 'c': abit = 2; break;
                                case 'G': case 'g': abit = 4; break;
                                case 'T': case 't': abit = 8; break;
                                default: abit = 0; atomicExch(g.err, 2);  // the reference throws (:202)
                            }
                            match = (cm & abit) != 0;
                        }
                        v4[k] = match ? pmv[k] : pxv[k];  // rows that are not read rows: both hold the same value
                    }
                }
#pragma unroll
                for (int v = 0; v < NV; ++v)
                    sts128(tab_lane + ((g.n_plain + 3 + sn) * NV + v) * 512, v4[4 * v], v4[4 * v + 1], v4[4 * v + 2], v4[4 * v + 3]);
            }
        }
        __syncwarp();

        FastState<K> st;
        float bM[K], bI[K], bD[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { st.M[k] = 0.f; st.I[k] = 0.f; st.D[k] = 0.f; bM[k] = 0.f; bI[k] = 0.f; bD[k] = 0.f; }
        if (lane == 31) st.D[K - 1] = c0;
        st.dgm = 0.f; st.dgi = 0.f; st.dgd = lane == 0 ? c0 : 0.f;
        float dgbm = 0.f, dgbi = 0.f, dgbd = 0.f;
        st.acc = 0.f;
        st.p = 0;
        st.sp = g.codes + hp.code_off - lane;  // lane l works on column (step - l); columns <= 0 and > H read the zero padding
        st.y = ldg_u8(st.sp);
        // SIMPLE: this lane's next deletion event, the column after the previous one, and its record behind the prior table
        uint32_t ev = hp.ev_first;
        const uint32_t ev_end = hp.ev_first + hp.n_events;
        constexpr int NO_COL = -(1 << 30);  // no column matches (p >= -31)
        int ea = NO_COL, eb = NO_COL, after_col = NO_COL;
        uint32_t after_par = 0;
        if (SIMPLE && ev < ev_end) { const uint2 e = g.events[ev]; ea = (int)e.x; eb = (int)e.y; }
        const uint32_t rec = tab_lane + (uint32_t)g.n_rows * (NV * 512);         // float4 v of this lane: rec + v * 512
        // what the next lane reads -- (M, I~, D~) of this lane's last row as saved, its unmerged (M, I~) at column b -- is kept
        // twice, by event parity: events may touch, and the next lane is one column behind
        const uint32_t hand0 = rec + NS * 512, unm0 = rec + (NS + 2) * 512;
        const int prev_delta = lane ? -16 : 31 * 16;

        int step = 1;
        for (uint32_t sg = 0; sg < hp.n_segs; ++sg) {
            const uint2 seg = g.segs[hp.seg_first + sg];
#pragma unroll 2
            for (uint32_t s = 0; s < seg.x; ++s)
                fast_step<K, false>(st, cb, cc, cg, cd, tab_lane, src_lane, lane, 0, 0, c0, nullptr, nullptr, 0, 0, 0, 0);
            step += (int)seg.x;
            int p = step - lane;  // 1-based column of this lane
            if constexpr (SIMPLE) {
#pragma unroll 1
                for (uint32_t s = 0; s < seg.y; ++s, ++p) {
                    float acc_in = 0.f;
                    const bool is_after = p == after_col;
                    if (is_after) {  // column b + 1 of the previous event (AFTER_DEL): merge column b with the branch in place
                        // the lane below takes this lane's last row at column b from the shuffle of THIS step, i.e. merged --
                        // right for every use the reference makes of it (:80-82, :107-118) except the accumulator row
                        sts128(unm0 + after_par * 512, st.M[K - 1], st.I[K - 1], 0.f, 0.f);
                        if (lane == acc_lane) {  // the accumulator row adds the unmerged (M + I)[R][b]
                            acc_in = __fmaf_rn(cb[0], st.dgi, st.dgm);  // (an empty read: row 0)
                            if (acc_slot == 0 && lane != 0) {  // row R is the previous lane's last row: its unmerged values were left one step ago
                                const float4 t4 = lds128m(unm0 + after_par * 512 + prev_delta);
                                acc_in = __fmaf_rn(cb[0], t4.y, t4.x);
                            }
#pragma unroll
                            for (int k = 1; k < K; ++k)
                                if (k == acc_slot) acc_in = __fmaf_rn(cb[k], st.I[k - 1], st.M[k - 1]);
                        }
                        float v[NS * 4];
#pragma unroll
                        for (int q = 0; q < NS; ++q) {
                            const float4 t4 = lds128m(rec + q * 512);
                            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
                        }
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            st.M[k] = pd_max(v[k], st.M[k]); st.I[k] = pd_max(v[K + k], st.I[k]); st.D[k] = pd_max(v[2 * K + k], st.D[k]);
                        }
                        if (lane != 0) {  // row above this lane's first row; row 0 has no branch
                            const float4 t4 = lds128m(hand0 + after_par * 512 + prev_delta);
                            st.dgm = pd_max(t4.x, st.dgm); st.dgi = pd_max(t4.y, st.dgi); st.dgd = pd_max(t4.z, st.dgd);
                        }
                    }
                    __syncwarp();  // (the lane above may start its next event in this step: its save comes after these reads)
                    if (p == ea) {  // column a: save column a - 1 -- merged a moment ago if the events touch -- (rows that are not read rows save 0)
                        float v[NS * 4];
#pragma unroll
                        for (int k = 0; k < NS * 4; ++k) v[k] = 0.f;
                        if (real_mask == (1u << K) - 1u) {
#pragma unroll
                            for (int k = 0; k < K; ++k) { v[k] = st.M[k]; v[K + k] = st.I[k]; v[2 * K + k] = st.D[k]; }
                        } else {
#pragma unroll
                            for (int k = 0; k < K; ++k)
                                if ((real_mask >> k) & 1u) { v[k] = st.M[k]; v[K + k] = st.I[k]; v[2 * K + k] = st.D[k]; }
                        }
#pragma unroll
                        for (int q = 0; q < NS; ++q) sts128(rec + q * 512, v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                        sts128(hand0 + (ev & 1u) * 512, v[K - 1], v[2 * K - 1], v[3 * K - 1], 0.f);
                    }
                    __syncwarp();
                    fast_step<K, false>(st, cb, cc, cg, cd, tab_lane, src_lane, lane, 0, 0, c0, nullptr, nullptr, 0, 0, 0, 0);
                    if (is_after) {
                        if (lane == acc_lane) {
#pragma unroll
                            for (int k = 0; k < K; ++k)
                                if (k == acc_slot) st.M[k] = acc_in;
                        }
                        after_col = NO_COL;
                    }
                    if (p == eb) {  // column b (DEL_END): the insertion chain sees max(branch, value) of the row above
                        float v[NS * 4];
#pragma unroll
                        for (int q = 0; q < NS; ++q) {
                            const float4 t4 = lds128m(rec + q * 512);
                            v[4 * q] = t4.x; v[4 * q + 1] = t4.y; v[4 * q + 2] = t4.z; v[4 * q + 3] = t4.w;
                        }
                        float um = st.dgm, ui = st.dgi;  // row above at this column (fast_step leaves them here)
                        if (lane != 0) {
                            const float4 t4 = lds128m(hand0 + (ev & 1u) * 512 + prev_delta);
                            um = pd_max(t4.x, um); ui = pd_max(t4.y, ui);
                        }
                        st.I[0] = __fmaf_rn(cg[0], ui, um);
#pragma unroll
                        for (int k = 1; k < K; ++k) st.I[k] = __fmaf_rn(cg[k], pd_max(v[K + k - 1], st.I[k - 1]), pd_max(v[k - 1], st.M[k - 1]));
                        // the next column is the event's AFTER_DEL column; the lane moves on to the next event (which may start there)
                        after_col = eb + 1; after_par = ev & 1u;
                        ++ev;
                        ea = NO_COL; eb = NO_COL;
                        if (ev < ev_end) { const uint2 e = g.events[ev]; ea = (int)e.x; eb = (int)e.y; }
                    }
                    __syncwarp();  // the records read in this step may be overwritten in the next one
                }
            } else {
            // the branch values are dead between two slow windows: every lane refreshes them (NORMAL: branch = the value one
            // column back) in the step before it reaches a column of the state machine, and that step is inside the window.
            // Saying so here keeps them out of the registers the fast loop holds
#pragma unroll
            for (int k = 0; k < K; ++k) { bM[k] = 0.f; bI[k] = 0.f; bD[k] = 0.f; }
            dgbm = 0.f; dgbi = 0.f; dgbd = 0.f;
#pragma unroll 1
            for (uint32_t s = 0; s < seg.y; ++s, ++p) {
                // ---- slow step: the full state machine, in the scaled representation ----
                constexpr uint32_t CODE_STRIDE = NV * 32 * 16;
                const uint32_t fl = ldg_u8(st.sp + flag_delta);
                ++st.sp;
                const uint32_t y_next = ldg_u8(st.sp);
                const bool valid = p >= 1 && p <= H;
                const float mu = __shfl_sync(FULL, st.M[K - 1], src_lane), iu = __shfl_sync(FULL, st.I[K - 1], src_lane);
                const float du = __shfl_sync(FULL, st.D[K - 1], src_lane);
                const float bmu = __shfl_sync(FULL, bM[K - 1], src_lane), biu = __shfl_sync(FULL, bI[K - 1], src_lane);
                const float bdu = __shfl_sync(FULL, bD[K - 1], src_lane);
                float pr[NV * 4];
                {
                    const uint32_t addr = st.y * CODE_STRIDE + tab_lane;
#pragma unroll
                    for (int v = 0; v < NV; ++v) {
                        const float4 q = lds128(addr + v * 512);
                        pr[v * 4 + 0] = q.x; pr[v * 4 + 1] = q.y; pr[v * 4 + 2] = q.z; pr[v * 4 + 3] = q.w;
                    }
                }
                const bool del_end = (fl & PD_DEL_END_BIT) != 0;
                const uint32_t t_row1 = (fl >> PD_TYPE_SHIFT) & PD_TYPE_BITS;
                uint32_t t_rest = t_row1;
                if (valid && (uint32_t)p <= t.first_event)
                    t_rest = t.carry == PD_INSIDE_DEL ? PD_INSIDE_DEL : (t.carry == PD_AFTER_DEL && p == 1 ? PD_AFTER_DEL : PD_NORMAL);
                float Mn[K], Dn[K], In[K], nbM[K], nbI[K], nbD[K];
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const bool is_real = (real_mask >> k) & 1u;
                    const uint32_t ty = !is_real ? PD_NORMAL : ((lane == 0 && k == 0) ? t_row1 : t_rest);
                    const bool inside = ty == PD_INSIDE_DEL, after = ty == PD_AFTER_DEL;
                    float am = k ? st.M[k - 1] : st.dgm, ai = k ? st.I[k - 1] : st.dgi, ad = k ? st.D[k - 1] : st.dgd;
                    const float abm = k ? bM[k - 1] : dgbm, abi = k ? bI[k - 1] : dgbi, abd = k ? bD[k - 1] : dgbd;
                    const float xm = pd_max(bM[k], st.M[k]), xi = pd_max(bI[k], st.I[k]), xd = pd_max(bD[k], st.D[k]);
                    nbM[k] = inside ? bM[k] : (after ? xm : st.M[k]);
                    nbI[k] = inside ? bI[k] : (after ? xi : st.I[k]);
                    nbD[k] = inside ? bD[k] : (after ? xd : st.D[k]);
                    am = after ? pd_max(abm, am) : am; ai = after ? pd_max(abi, ai) : ai; ad = after ? pd_max(abd, ad) : ad;
                    const float lm = after ? xm : st.M[k], ld = after ? xd : st.D[k];  // this row, previous column
                    float u = __fmaf_rn(cc[k], ad, am);
                    u = __fmaf_rn(cb[k], ai, u);
                    Mn[k] = pr[k] * u;
                    Dn[k] = __fmaf_rn(cd[k], ld, lm);
                }
                // insertion: row above at THIS column -- chain down the lane's rows (DEL_END merges the branch in)
                float um = mu, ui = iu, ubm = bmu, ubi = biu;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const bool is_real = (real_mask >> k) & 1u;
                    const bool merge = del_end && is_real;
                    const float xm = merge ? pd_max(ubm, um) : um, xi = merge ? pd_max(ubi, ui) : ui;
                    In[k] = __fmaf_rn(cg[k], xi, xm);
                    um = Mn[k]; ui = In[k]; ubm = nbM[k]; ubi = nbI[k];
                }
#pragma unroll
                for (int k = 0; k < K; ++k) { st.M[k] = Mn[k]; st.I[k] = In[k]; st.D[k] = Dn[k]; bM[k] = nbM[k]; bI[k] = nbI[k]; bD[k] = nbD[k]; }
                st.dgm = mu; st.dgi = iu; st.dgd = du; dgbm = bmu; dgbi = biu; dgbd = bdu;
                st.y = y_next;
            }
            }  // !SIMPLE
            step += (int)seg.y;
        }
        // every lane is past column H + 1: the accumulator row holds sum_j (M + I)[R][j] as M + D~
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k)
            if (k == acc_slot) v = st.M[k] + st.D[k];
        v = __shfl_sync(FULL, v, acc_lane);
        if (lane == 0) g.sums[g.first + ti] = v;
        }  // haplotypes of the group
    }
}

constexpr float PD_RESCUE_THRESHOLD_F32 = 1e-28f;

// fp32 sums -> log10 likelihoods, or onto the redo list
__global__ void __launch_bounds__(128) phmm_pd_epilogue_f32(const PdTask *tasks, uint32_t n, const float *sums, double *out,
                                                            uint32_t *redo, uint32_t *n_redo)
{
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const PdTask t = tasks[k];
        const float s = sums[k];
        if (!(s >= PD_RESCUE_THRESHOLD_F32) || s > 3.0e38f) {
            redo[atomicAdd(n_redo, 1u)] = k;
            continue;
        }
        out[t.out_slot] = log10((double)s) - log10_c0H(t.c0_exp, t.H);
    }
}

__global__ void __launch_bounds__(128) phmm_pd_epilogue_f64(const PdTask *tasks, const uint32_t *redo, const uint32_t *n_redo,
                                                            const double *sums, double *out)
{
    const uint32_t n = *n_redo;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const PdTask t = tasks[redo[k]];
        out[t.out_slot] = log10(sums[k]) - 307.0505955772608;  // Math.log10(Math.pow(2, 1020)), LoglessPDPairHMM.java:12,153
    }
}

}  // namespace phmm_dev
