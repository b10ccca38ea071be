// phmm_kernels.cuh -- PairHMM forward kernels for sm_100a (B200).
//
// Computes what LoglessPairHMM.subComputeReadLikelihoodGivenHaplotypeLog10 computes
// (reference: src/main/java/org/broadinstitute/hellbender/utils/pairhmm/LoglessPairHMM.java:20-68,
// priors :79-93, transitions PairHMMModel.java:107-117) but laid out for a GPU warp:
//
//  * One warp = one task = one read against a back-to-back stream of haplotypes of its region.
//    Lane l owns K consecutive read rows (l*K+1 .. l*K+K); their M/I/D state lives in registers.
//    The warp sweeps the stream one column per step; lane l works on column (step - l), so every
//    step hands the last row of each lane to the next lane with three shuffles -- the anti-diagonal
//    wavefront.
//  * The recurrence is re-scaled so a cell costs 6 FP instructions (DESIGN.md "Kernel recurrence"):
//        I~ = I / tMI_i ,  D~ = D / tMD_i
//        M[i][j]  = prior(i,j) * ( a_i*M[i-1][j-1] + b_i*I~[i-1][j-1] + c_i*D~[i-1][j-1] )
//        D~[i][j] = M[i][j-1] + tDD_i * D~[i][j-1]
//        I~[i][j] = M[i-1][j] + g_i  * I~[i-1][j]
//    with a=tMM_i, b=tIM_i*tMI_{i-1}, c=tIM_i*tMD_{i-1}, g=tII_i*tMI_{i-1}/tMI_i, tMI_0=tMD_0=1.
//  * prior(i,j) comes from a per-task shared-memory table indexed by the haplotype code of the
//    column: one 16-byte LDS returns the priors of 4 (fp32) / 2 (fp64) rows of the lane.
//  * Haplotypes are separated by an END column (prior 0): M and I~ vanish there by themselves.  There is no
//    pipeline drain between haplotypes.
//
// Kernels in this file:
//   phmm_forward_kernel<T,K,STRIPED>  general reference kernel: explicit lane-0 selects, running sum carried down pad
//                                     rows, reads of any length in strips.  Used for reads of 255+ bases (fp32) and for
//                                     the fp64 redo of reads that are not flat-quality.
//   phmm_fast_f32_kernel<K>           per-base qualities, reads <= 254 bases: rotation hand-off, accumulator row,
//                                     branch-free step loop, host-planned schedule with haplotype-prefix sharing.
//   phmm_flat_f32_kernel<K,MODE,LANES> the same for reads with flat insertion/deletion/GCP qualities (MODE_FLAT): transition
//                                     coefficients are kernel parameters (constant / uniform-register operands).
//                                     MODE_SYM: ins == del per base with a flat GCP (the PCR indel model): two per-row
//                                     coefficients left, the rest still constant operands.  MODE_GEN: four per-row
//                                     coefficients (any per-base qualities; LANES = 16 only).  LANES = 32 / 16 / 8: one,
//                                     two or four reads of a unit per warp (full- / half- / quarter-warp form).
//   phmm_flat_f64_kernel              fp64 redo of flat-quality reads (one read x one haplotype per task).
//   phmm_classify_kernel              per read: flat class, symmetric class or general.
//   phmm_epilogue_f32 / _rescue       raw sums -> log10 likelihoods; builds / consumes the fp64 redo list; the rescue
//                                     epilogue puts sums it cannot trust on the deep list.
//   phmm_exact_f64_kernel             last tier: the reference's own unscaled arithmetic for the deep list.
// The kernels of the rows built around this path live in region_steps.cuh, pdhmm_kernels.cuh and sw_kernels.cuh.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace phmm_dev {

// Unroll factors of the flat kernels' step loops (branch-free / checked).  Measured on B200 (A/B builds on one box): with 16
// rows per lane 2 / 1 is best (4 / 1 loses 22 %, 2 / 2 loses 12 %: the 16 sweep instances of a kernel outgrow the instruction
// cache), with 10 rows per lane (half-warp form, 150-base reads) 4 / 2 gains 2-3 %; the full-warp kernels spill at 4.
__host__ __device__ constexpr int flat_unroll(int K) { return K == 10 ? 4 : 2; }
__host__ __device__ constexpr int flat_chk_unroll(int K) { return K == 10 ? 2 : 1; }
// (the full-warp kernels use K <= 8 and spill at 4; the half-warp kernels of 6 and 8 rows per lane are separate instantiations)
constexpr uint32_t CODE_END = 0;   // column after the last base of a haplotype
constexpr uint32_t CODE_NULL = 1;  // outside the stream (pipeline fill / drain)
constexpr uint32_t CODE_FIRST_BASE = 2;  // A C G T = 2..5, further byte values (N included) 6..
constexpr int MAX_CODES = 64;
constexpr int MAX_QUAL = 254;      // QualityUtils.java:43
constexpr float RESCUE_THRESHOLD_F32 = 1e-28f;
constexpr int STREAM_PAD = 32;            // NULL codes around every haplotype stream (pipeline fill / drain)
constexpr uint8_t CLASS_GENERAL = 0xff;   // read_class of reads without a flat-quality class
constexpr int MAX_FLAT_CLASSES = 4;
constexpr int MAX_SYM_CLASSES = 4;        // read_class MAX_FLAT_CLASSES + k: ins == del per base, flat gcp (PCR indel model)
constexpr uint32_t SYM_MAX_GAP_QUAL = 70; // gap-open quals above this go to the general kernel (M^ = M * 2^10 * eps would lose range)

// D[0][j] of the reference is 2^1020/H (LoglessPairHMM.java:8,31).  The kernels use a power of two
// 2^(BASE - ceil(log2 H)) <= that leaves headroom for the scaled states I~ and D~.
constexpr int C0_BASE_EXP_F32 = 116;
constexpr int C0_BASE_EXP_F64 = 960;

struct __align__(16) Task {
    uint32_t read;        // chunk-local read index
    uint32_t stream_off;  // first column of the haplotype stream in `streams`
    uint32_t stream_len;  // columns in the stream, one END per haplotype included
    uint32_t out_base;    // slot of the stream's first haplotype in `sums`
    int32_t c0_exp;       // D[0][j] = 2^c0_exp
    uint32_t n_haps;      // haplotypes in the stream
    uint32_t hap_first;   // chunk-local index of the stream's first haplotype (hap_len[]); rescue: final output slot
    uint32_t unit;        // chunk-local unit index (unit_sched[]); rescue: haplotype length
};

// ---- prefix sharing between the haplotypes of a unit ("LOGLESS_CACHING": PairHMM.java:296-313,341-352) ----
// The host sorts a unit's haplotypes lexicographically; consecutive haplotypes share a prefix whose DP columns
// are identical (same read, same initial condition), so a pass over haplotype i+1 starts from a register
// SNAPSHOT taken at the last shared column during an earlier pass instead of column 1.  The fast kernels run a
// host-planned schedule of segments: n_free branch-free steps, then n_chk checked steps during which each lane
// (a) snapshots its state when it has just finished stream position snap_pos, (b) handles END columns: write
// the haplotype's sum, then restore the next pass's snapshot (or start from zero).  The planner makes every pass
// at least 32 stream positions long (NULL columns before END if needed), so the 32-step windows of two END columns
// never overlap and everything a lane needs at an END is uniform per segment.
struct PassInfo {
    uint16_t out_idx;      // haplotype index inside the unit = output column of this pass
    int16_t restore_slot;  // snapshot restored BEFORE this pass starts; -1 = fresh start
};
struct Segment {
    uint32_t n_free, n_chk;
    int32_t snap_pos;      // stream position (1-based) whose completion triggers a snapshot in this segment, INT32_MIN = none
    uint8_t snap_slot;
    int8_t end_restore;    // the (single) END column met in this segment: snapshot slot the next pass restores, -1 = fresh start
    uint16_t end_out;      // ... and the output column of the pass that ends there
};
struct UnitSched {
    uint32_t pass_first, n_passes;  // into pass_info
    uint32_t seg_first, n_segs;     // into segments
    uint32_t sstream_off;           // first column of the unit's shared (prefix-compressed) stream
    uint32_t seg16_first, n_segs16; // the same schedule with 16-step windows (half-warp kernels); n_segs16 = 0: not planned
    uint32_t seg8_first, n_segs8;   // ... with 8-step windows (quarter-warp kernels)
    uint32_t pad2;
};
constexpr int MAX_SNAP_SLOTS = 8;      // + 1 slot per CTA for the pass-start state
constexpr int SNAP_REGS = 3 * 8 + 4;  // M, I~, D~ of up to 8 rows + hand-off triple + accumulator
// rows per lane above 8 (half-warp kernels): the record grows with K (multiple of 4 floats: 128-bit accesses)
__host__ __device__ constexpr int snap_regs(int K) { return K <= 8 ? SNAP_REGS : ((3 * K + 4 + 3) / 4) * 4; }
__host__ __device__ constexpr size_t snap_slab_bytes(int K) { return (size_t)(MAX_SNAP_SLOTS + 1) * snap_regs(K) * 32 * sizeof(float); }

struct KernelArgs {
    const uint8_t *rd_bases, *rd_q, *rd_i, *rd_d, *rd_c;  // raw per-base read arrays of the chunk
    const uint32_t *read_off;                             // chunk-local, n_reads+1
    const uint8_t *streams;                               // haplotype code streams
    const uint32_t *hap_len;                              // chunk-local haplotype lengths
    const Task *tasks;
    const uint32_t *n_tasks_ptr;  // device-side task count (rescue lists) or nullptr
    uint32_t n_tasks;             // host-side task count when n_tasks_ptr == nullptr
    uint32_t *counter;            // work-queue cursor, zeroed before the launch
    void *sums;                   // float* or double*: raw last-row sums per (task, haplotype)
    void *bnd;                    // STRIPED: boundary buffer, bnd_stride columns per CTA
    uint32_t bnd_stride;
    const double *m2m;            // triangular matchToMatch table (PairHMMModel.java:71,86-94)
    int *err;                     // device error flag (GPHMM_ERR_BAD_QUAL)
    const uint8_t *read_class;    // per read: flat-quality class id (phmm_classify_kernel) or CLASS_GENERAL
    const uint8_t *sstreams;      // shared (prefix-compressed) haplotype streams of the fast kernels
    const UnitSched *unit_sched;
    const PassInfo *pass_info;
    const Segment *segments;
    float *snap;                  // snapshot slab: per CTA MAX_SNAP_SLOTS x SNAP_REGS x 32 floats
    int32_t n_codes;
    int32_t tristate_off;
    int32_t pair_tasks;            // 1: tasks of a half-warp bucket: a second read in stream_off (NO_READ = none), its out_base in stream_len
                                   // 2: tasks of a quarter-warp bucket: reads in read, stream_off, n_haps, hap_first (NO_READ = none), the sums of
                                   //    read r start at out_base + r * stream_len (stream_len = haplotypes of the unit, arithmetic mod 2^32)
    uint8_t code_byte[MAX_CODES];  // code -> haplotype byte value
};
constexpr uint32_t NO_READ = 0xffffffffu;

__constant__ double c_eps[256];  // QualityUtils.qualToErrorProb cache: 10^(-q/10), q = 0..254

template <typename T> struct Vec16;
template <> struct Vec16<float> { using type = float4; static constexpr int W = 4; };
template <> struct Vec16<double> { using type = double2; static constexpr int W = 2; };

template <typename T> struct __align__(16) Bnd { T m, i, d, pad; };

__device__ __forceinline__ float fma_(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ double fma_(double a, double b, double c) { return __fma_rn(a, b, c); }

template <typename T, int W> __device__ __forceinline__ void unpack(const typename Vec16<T>::type &v, T *dst);
template <> __device__ __forceinline__ void unpack<float, 4>(const float4 &v, float *dst) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w; }
template <> __device__ __forceinline__ void unpack<double, 2>(const double2 &v, double *dst) { dst[0] = v.x; dst[1] = v.y; }

// Dynamic shared memory a CTA (one warp) needs for its prior table.
template <typename T, int K> constexpr size_t prior_table_bytes(int n_codes) {
    return (size_t)n_codes * ((K + Vec16<T>::W - 1) / Vec16<T>::W) * 32 * 16;
}

template <typename T, int K, bool STRIPED>
__global__ void __launch_bounds__(32) phmm_forward_kernel(const KernelArgs g)
{
    using V = typename Vec16<T>::type;
    constexpr int W = Vec16<T>::W;
    constexpr int NV = (K + W - 1) / W;
    constexpr int KT = NV * W;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int ROWS_PER_STRIP = 32 * K;
    // fp32: tMM folded into the prior table (5 FP instructions per cell, tMM = 0 poisons the pair); fp64: the redo tier keeps
    // the multiply so that every legal quality, tMM = 0 included, is handled here and not by the one-thread-per-pair tier
    constexpr bool FOLD = sizeof(T) == 4;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    V *tab = reinterpret_cast<V *>(smem_raw);
    T *tab_s = reinterpret_cast<T *>(smem_raw);

    const int lane = threadIdx.x;
    T *const sums = reinterpret_cast<T *>(g.sums);
    Bnd<T> *const bnd = STRIPED ? reinterpret_cast<Bnd<T> *>(g.bnd) + (size_t)blockIdx.x * g.bnd_stride : nullptr;
    const uint32_t n_tasks = g.n_tasks_ptr ? *g.n_tasks_ptr : g.n_tasks;
    const int n_codes = g.n_codes;

    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(g.counter, 1u);
        ti = __shfl_sync(FULL, ti, 0);
        if (ti >= n_tasks) break;
        const Task t = g.tasks[ti];
        const uint32_t ro = g.read_off[t.read];
        const int R = (int)(g.read_off[t.read + 1] - ro);
        // fp64 rescue lists: flat-quality reads of up to 254 bases belong to phmm_flat_f64_kernel
        if (g.read_class != nullptr && g.read_class[t.read] < MAX_FLAT_CLASSES && R <= 254) continue;
        const int P = (int)t.stream_len;
        const uint8_t *__restrict__ stream = g.streams + t.stream_off;
        const T c0 = (T)scalbn(1.0, t.c0_exp);
        // The last strip must hold at least one pad row, hence R / ROWS + 1 strips.
        const int n_strips = STRIPED ? R / ROWS_PER_STRIP + 1 : 1;

        for (int strip = 0; strip < n_strips; ++strip) {
            const bool first_strip = !STRIPED || strip == 0;
            const bool last_strip = !STRIPED || strip == n_strips - 1;
            // ---- per-strip setup: transition coefficients into registers, priors into shared memory ----
            T ca[K], cb[K], cc[K], cg[K], cd[K];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int i = strip * ROWS_PER_STRIP + lane * K + k + 1;  // 1-based read row
                double A = 0.0, B = 0.0, C = 0.0, G = 1.0, DD = 0.0, pm = 0.0, px = 0.0;
                uint32_t x = 0;
                const bool real = i <= R;
                if (real) {
                    uint32_t q = g.rd_q[ro + i - 1], qi = g.rd_i[ro + i - 1], qd = g.rd_d[ro + i - 1], qc = g.rd_c[ro + i - 1];
                    x = g.rd_bases[ro + i - 1];
                    if (q > (uint32_t)MAX_QUAL || qi > 127u || qd > 127u || qc > 127u) {
                        atomicExch(g.err, 1);
                        q = min(q, (uint32_t)MAX_QUAL); qi = min(qi, 127u); qd = min(qd, 127u); qc = min(qc, 127u);
                    }
                    double tmi_prev = 1.0, tmd_prev = 1.0;
                    if (i > 1) {
                        tmi_prev = c_eps[min((uint32_t)g.rd_i[ro + i - 2], 127u)];
                        tmd_prev = c_eps[min((uint32_t)g.rd_d[ro + i - 2], 127u)];
                    }
                    const double ei = c_eps[qi], ec = c_eps[qc];
                    const uint32_t mn = min(qi, qd), mx = max(qi, qd);
                    const double tIM = 1.0 - ec;
                    A = __ldg(g.m2m + ((mx * (mx + 1)) >> 1) + mn);
                    B = tIM * tmi_prev;
                    C = tIM * tmd_prev;
                    G = ec * tmi_prev / ei;
                    DD = ec;
                    const double e = c_eps[q];
                    pm = 1.0 - e;
                    px = g.tristate_off ? e : e / 3.0;
                    if (FOLD) {
                        // tMM = 0 cannot be factored out: NaN coefficients, the pair is redone in double, where the multiply stays
                        const double inv = A > 0.0 ? 1.0 / A : __longlong_as_double(0x7ff8000000000000LL);
                        pm *= A; px *= A;
                        B *= inv; C *= inv;
                    }
                } else if (i == R + 1) {
                    G = R >= 1 ? c_eps[min((uint32_t)g.rd_i[ro + R - 1], 127u)] : 1.0;
                }
                ca[k] = (T)A; cb[k] = (T)B; cc[k] = (T)C; cg[k] = (T)G; cd[k] = (T)DD;
                const T pmT = (T)pm, pxT = (T)px;
                for (int y = 0; y < n_codes; ++y) {
                    T v = (T)0;
                    if (real && y >= (int)CODE_FIRST_BASE) {
                        const uint32_t hb = g.code_byte[y];
                        v = (x == hb || x == (uint32_t)'N' || hb == (uint32_t)'N') ? pmT : pxT;  // LoglessPairHMM.java:89
                    }
                    tab_s[((y * NV + k / W) * 32 + lane) * W + (k % W)] = v;
                }
            }
            __syncwarp();

            // ---- the sweep ----
            T M[K], I[K], D[K];
#pragma unroll
            for (int k = 0; k < K; ++k) { M[k] = (T)0; I[k] = (T)0; D[k] = (T)0; }
            T dgm = (T)0, dgi = (T)0, dgd = first_strip ? c0 : (T)0;  // row above, previous column
            T sum = (T)0;
            uint32_t hap_idx = 0;
            int p = 1 - lane;  // this lane's column in the current step (1-based)
            uint32_t y = ((unsigned)(p - 1) < (unsigned)P) ? (uint32_t)__ldg(stream + p - 1) : CODE_NULL;
            Bnd<T> up_next;  // lane 0 of strips > 0: row above at column p, prefetched
            up_next.m = up_next.i = up_next.d = (T)0;
            if (STRIPED && !first_strip && lane == 0 && P > 0) up_next = bnd[0];

            const int n_steps = P + 31;
            for (int s = 1; s <= n_steps; ++s) {
                // prefetch the next column's code (and boundary row)
                const uint32_t y_next = ((unsigned)p < (unsigned)P) ? (uint32_t)__ldg(stream + p) : CODE_NULL;
                T mu = __shfl_up_sync(FULL, M[K - 1], 1);
                T iu = __shfl_up_sync(FULL, I[K - 1], 1);
                T du = __shfl_up_sync(FULL, D[K - 1], 1);
                if (lane == 0) {
                    if (first_strip) { mu = (T)0; iu = (T)0; du = c0; }
                    else { mu = up_next.m; iu = up_next.i; du = up_next.d; }
                }
                if (STRIPED && !first_strip && lane == 0) {
                    if ((unsigned)p < (unsigned)P) up_next = bnd[p];
                    else { up_next.m = (T)0; up_next.i = (T)0; up_next.d = (T)0; }
                }
                // priors of this lane's K rows for haplotype code y
                T pr[KT];
                {
                    const V *tp = tab + (size_t)y * (NV * 32) + lane;
#pragma unroll
                    for (int v = 0; v < NV; ++v) unpack<T, W>(tp[v * 32], pr + v * W);
                }
                // M of column p from column p-1 (old M/I/D of the row above)
                T Mn[K];
                if (FOLD) {  // tMM is folded into the prior table and divides b, c (see fast_step)
                    T u = fma_(cc[0], dgd, dgm);
                    u = fma_(cb[0], dgi, u);
                    Mn[0] = pr[0] * u;
#pragma unroll
                    for (int k = 1; k < K; ++k) {
                        u = fma_(cc[k], D[k - 1], M[k - 1]);
                        u = fma_(cb[k], I[k - 1], u);
                        Mn[k] = pr[k] * u;
                    }
                } else {
                    T u = cc[0] * dgd;
                    u = fma_(cb[0], dgi, u);
                    u = fma_(ca[0], dgm, u);
                    Mn[0] = pr[0] * u;
#pragma unroll
                    for (int k = 1; k < K; ++k) {
                        u = cc[k] * D[k - 1];
                        u = fma_(cb[k], I[k - 1], u);
                        u = fma_(ca[k], M[k - 1], u);
                        Mn[k] = pr[k] * u;
                    }
                }
                // D~ of column p from column p-1 of the same row
#pragma unroll
                for (int k = 0; k < K; ++k) D[k] = fma_(cd[k], D[k], M[k]);
                // I~ of column p: chain down the column
                I[0] = fma_(cg[0], iu, mu);
#pragma unroll
                for (int k = 1; k < K; ++k) I[k] = fma_(cg[k], I[k - 1], Mn[k - 1]);
#pragma unroll
                for (int k = 0; k < K; ++k) M[k] = Mn[k];
                dgm = mu; dgi = iu; dgd = du;
                sum += I[K - 1];  // meaningful on lane 31 of the last strip: (M+I)[R][p]

                if (y == CODE_END) {
#pragma unroll
                    for (int k = 0; k < K; ++k) D[k] = (T)0;
                    if (last_strip && lane == 31) sums[t.out_base + hap_idx] = sum;
                    ++hap_idx;
                    sum = (T)0;
                }
                if (STRIPED && !last_strip && lane == 31 && (unsigned)(p - 1) < (unsigned)P) {
                    Bnd<T> o;
                    o.m = M[K - 1]; o.i = I[K - 1]; o.d = D[K - 1]; o.pad = (T)0;
                    bnd[p - 1] = o;
                }
                y = y_next;
                ++p;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Fast fp32 kernel for reads of up to 32*K-2 bases (one strip).  Same recurrence as above, with the
// per-step overhead stripped:
//  * haplotype streams are physically padded with STREAM_PAD NULL codes on both sides, so the
//    column code is one unguarded byte load;
//  * the hand-off between lanes is a rotation (lane 0 reads lane 31): the last row of lane 31 is a
//    pad row that holds (M, I~, D~) = (0, *, c0), which is exactly the virtual row 0 lane 0 needs
//    (its b and g coefficients are 0 because I~ of row 0 is 0) -- no per-step lane-0 selects;
//  * the row below the read (row R+1) is an accumulator row: a=1, b=tMI_R, prior=1, tDD=1 make
//        M[R+1][j]  = (M+I)[R][j-1]          D~[R+1][j] = sum_{j' < j-1} (M+I)[R][j']
//    so the haplotype's likelihood sum is M+D~ of that row at the END column -- no per-step add;
//  * the prior table is addressed with one IMAD from a precomputed 32-bit shared address.
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

__device__ __forceinline__ uint32_t ldg_u8(const uint8_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Per-warp register state of the fast kernel.
template <int K> struct FastState {
    float M[K], I[K], D[K];
    float dgm, dgi, dgd;  // row above this lane's first row, previous column
    float acc;            // flat kernel only: running sum of (M+I)[R][.]
    uint32_t y;           // haplotype code of this lane's current column
    const uint8_t *sp;    // points at that code
    int p;                // checked steps only: stream position of this lane's current column
};

// snapshot slab layout: [slot][lane][SNAP_REGS]: only one lane saves / restores in any given step (lanes reach a
// stream position one step apart), so each lane moves its own 112-byte record with 128-bit accesses.  `rec` is the
// lane's record of slot 0; slot MAX_SNAP_SLOTS holds the pass-start state (all zero except the row-0 carrier), so
// that "start from scratch" and "resume from a snapshot" are the same seven loads.
constexpr int ZERO_SLOT = MAX_SNAP_SLOTS;
constexpr int SLOT_STRIDE = 32 * SNAP_REGS;  // floats between two slots of the same lane (K <= 8)
template <int K> __device__ __forceinline__ void snap_save(const FastState<K> &st, float *rec, int slot) {
    constexpr int REGS = snap_regs(K);
    constexpr int SLOT_STRIDE = 32 * REGS;
    float v[REGS];
#pragma unroll
    for (int k = 0; k < REGS; ++k) v[k] = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) { v[3 * k + 0] = st.M[k]; v[3 * k + 1] = st.I[k]; v[3 * k + 2] = st.D[k]; }
    v[3 * K + 0] = st.dgm; v[3 * K + 1] = st.dgi; v[3 * K + 2] = st.dgd; v[3 * K + 3] = st.acc;
    float4 *q = reinterpret_cast<float4 *>(rec + (size_t)slot * SLOT_STRIDE);
#pragma unroll
    for (int k = 0; k < (3 * K + 4 + 3) / 4; ++k) q[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}
template <int K> __device__ __forceinline__ void snap_restore(FastState<K> &st, const float *rec, int slot) {
    constexpr int REGS = snap_regs(K);
    constexpr int SLOT_STRIDE = 32 * REGS;
    float v[REGS];
    const float4 *q = reinterpret_cast<const float4 *>(rec + (size_t)slot * SLOT_STRIDE);
#pragma unroll
    for (int k = 0; k < (3 * K + 4 + 3) / 4; ++k) {
        const float4 t = q[k];
        v[4 * k] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) { st.M[k] = v[3 * k + 0]; st.I[k] = v[3 * k + 1]; st.D[k] = v[3 * k + 2]; }
    st.dgm = v[3 * K + 0]; st.dgi = v[3 * K + 1]; st.dgd = v[3 * K + 2]; st.acc = v[3 * K + 3];
}

// One wavefront step.  CHECKED steps handle the END column (haplotype boundary); unchecked steps may only
// run while no lane of the warp sits on an END column, which the caller guarantees from the haplotype
// lengths -- that loop is branch-free.
template <int K, bool CHECKED>
__device__ __forceinline__ void fast_step(FastState<K> &st, const float (&cb)[K], const float (&cc)[K],
                                          const float (&cg)[K], const float (&cd)[K], uint32_t tab_lane, int src_lane, int lane,
                                          int acc_lane, int acc_slot, float c0, float *sums_task, float *slab, int snap_pos,
                                          int snap_slot, int end_restore, uint32_t end_out)
{
    constexpr int NV = (K + 3) / 4;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t CODE_STRIDE = NV * 32 * 16;
    ++st.sp;
    const uint32_t y_next = ldg_u8(st.sp);
    const float mu = __shfl_sync(FULL, st.M[K - 1], src_lane);
    const float iu = __shfl_sync(FULL, st.I[K - 1], src_lane);
    const float du = __shfl_sync(FULL, st.D[K - 1], src_lane);
    float pr[NV * 4];
    {
        const uint32_t addr = st.y * CODE_STRIDE + tab_lane;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float4 q = lds128(addr + v * 512);
            pr[v * 4 + 0] = q.x; pr[v * 4 + 1] = q.y; pr[v * 4 + 2] = q.z; pr[v * 4 + 3] = q.w;
        }
    }
    // M = (prior * a) * (M_diag + (b/a) * I~_diag + (c/a) * D~_diag): tMM is folded into the prior table and into b, c,
    // so the match update is 2 FFMA + 1 FMUL (5 FP instructions per cell in all)
    float Mn[K];
    {
        float u = __fmaf_rn(cc[0], st.dgd, st.dgm);
        u = __fmaf_rn(cb[0], st.dgi, u);
        Mn[0] = pr[0] * u;
    }
#pragma unroll
    for (int k = 1; k < K; ++k) {
        float u = __fmaf_rn(cc[k], st.D[k - 1], st.M[k - 1]);
        u = __fmaf_rn(cb[k], st.I[k - 1], u);
        Mn[k] = pr[k] * u;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) st.D[k] = __fmaf_rn(cd[k], st.D[k], st.M[k]);
    st.I[0] = __fmaf_rn(cg[0], iu, mu);
#pragma unroll
    for (int k = 1; k < K; ++k) st.I[k] = __fmaf_rn(cg[k], st.I[k - 1], Mn[k - 1]);
#pragma unroll
    for (int k = 0; k < K; ++k) st.M[k] = Mn[k];
    st.dgm = mu; st.dgi = iu; st.dgd = du;
    if (CHECKED) {
        if (__builtin_expect(st.p == snap_pos, 0)) snap_save<K>(st, slab, snap_slot);
        if (__builtin_expect(st.y == CODE_END, 0)) {
            if (lane == acc_lane) {
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (k == acc_slot) v = st.M[k] + st.D[k];
                sums_task[end_out] = v;
            }
            snap_restore<K>(st, slab, end_restore);  // ZERO_SLOT = fresh start
        }
        ++st.p;
    }
    st.y = y_next;
}

template <int K>
__global__ void __launch_bounds__(32) phmm_fast_f32_kernel(const KernelArgs g)
{
    constexpr int NV = (K + 3) / 4;
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tab_s = reinterpret_cast<float *>(smem_raw);
    int lane, src_lane;
    // opaque to the optimiser: keeps lane / rotation source in registers instead of re-deriving them every step
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("{ .reg .u32 t; add.u32 t, %1, 31; and.b32 %0, t, 31; }" : "=r"(src_lane) : "r"(lane));
    const uint32_t tab_lane = (uint32_t)__cvta_generic_to_shared(smem_raw) + lane * 16;
    float *const sums = reinterpret_cast<float *>(g.sums);
    const uint32_t n_tasks = g.n_tasks;
    const int n_codes = g.n_codes;

    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(g.counter, 1u);
        ti = __shfl_sync(FULL, ti, 0);
        if (ti >= n_tasks) break;
        const Task t = g.tasks[ti];
        // a task of a half-warp bucket carries two reads: this (full-warp) kernel takes them one after the other
        for (int sub = 0; sub < (g.pair_tasks ? 2 : 1); ++sub) {
        const uint32_t rd = sub ? t.stream_off : t.read;
        if (rd == NO_READ) continue;
        if (g.read_class[rd] != CLASS_GENERAL) continue;  // a flat-quality kernel owns this read
        const uint32_t out_base = sub ? t.stream_len : t.out_base;
        const uint32_t ro = g.read_off[rd];
        const int R = (int)(g.read_off[rd + 1] - ro);  // host guarantees R + 2 <= 32 * K
        const float c0 = (float)scalbn(1.0, t.c0_exp);
        const int acc_lane = R / K, acc_slot = R % K;  // accumulator row = 0-based row R

        float cb[K], cc[K], cg[K], cd[K];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = lane * K + k + 1;  // 1-based read row
            double A = 0.0, B = 0.0, C = 0.0, G = 0.0, DD = 0.0, pm = 0.0, px = 0.0;
            uint32_t x = 0;
            const bool real = i <= R;
            if (real) {
                uint32_t q = g.rd_q[ro + i - 1], qi = g.rd_i[ro + i - 1], qd = g.rd_d[ro + i - 1], qc = g.rd_c[ro + i - 1];
                x = g.rd_bases[ro + i - 1];
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
                // tMM goes into the priors and divides b and c.  tMM = 0 (insertion + deletion error >= 1, qualities <= 3)
                // cannot be factored out: NaN coefficients poison the sums and the pairs are redone by the fp64 kernels
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
            for (int y = 0; y < n_codes; ++y) {
                float v = 0.f;
                if (real) {
                    if (y >= (int)CODE_FIRST_BASE) {
                        const uint32_t hb = g.code_byte[y];
                        v = (x == hb || x == (uint32_t)'N' || hb == (uint32_t)'N') ? pmf : pxf;  // LoglessPairHMM.java:89
                    }
                } else if (i == R + 1) {
                    v = 1.f;
                }
                tab_s[((y * NV + k / 4) * 32 + lane) * 4 + (k % 4)] = v;
            }
        }
        __syncwarp();

        FastState<K> st;
#pragma unroll
        for (int k = 0; k < K; ++k) { st.M[k] = 0.f; st.I[k] = 0.f; st.D[k] = 0.f; }
        if (lane == 31) st.D[K - 1] = c0;
        st.dgm = 0.f; st.dgi = 0.f; st.dgd = lane == 0 ? c0 : 0.f;
        st.acc = 0.f;
        st.p = 0;
        const UnitSched us = g.unit_sched[t.unit];
        // lane l works on stream position (step - l); positions <= 0 and > P read the NULL padding around the stream
        st.sp = g.sstreams + us.sstream_off - lane;
        st.y = ldg_u8(st.sp);
        float *const sums_task = sums + out_base;
        float *const slab = g.snap + (size_t)blockIdx.x * ((MAX_SNAP_SLOTS + 1) * SLOT_STRIDE) + lane * SNAP_REGS;
        snap_save<K>(st, slab, ZERO_SLOT);  // the pass-start state, restored at every END that begins a fresh pass

        int step = 1;
        for (uint32_t sg = 0; sg < us.n_segs; ++sg) {
            const Segment seg = g.segments[us.seg_first + sg];
#pragma unroll 2
            for (uint32_t s = 0; s < seg.n_free; ++s)
                fast_step<K, false>(st, cb, cc, cg, cd, tab_lane, src_lane, lane, acc_lane, acc_slot, c0, sums_task, slab, 0, 0, 0, 0);
            step += (int)seg.n_free;
            st.p = step - lane;
#pragma unroll 1
            for (uint32_t s = 0; s < seg.n_chk; ++s)
                fast_step<K, true>(st, cb, cc, cg, cd, tab_lane, src_lane, lane, acc_lane, acc_slot, c0, sums_task, slab, seg.snap_pos,
                                   seg.snap_slot, seg.end_restore, seg.end_out);
            step += (int)seg.n_chk;
        }
        }  // sub
    }
}

// ---------------------------------------------------------------------------------------------
// Flat-quality fp32 kernel.  When a read has the same insertion, deletion and gap-continuation quality on
// every base (HaplotypeCaller's default: flat Q45/Q45 unless BI/BD tags exist, flat GCP 10 --
// StandardPairHMMInputScoreImputator.java:27-47) the transition coefficients a,b,c,g,d are the same for every
// row, so they are passed as KERNEL PARAMETERS and reach the FMA pipe as constant-bank operands.  That cuts the
// register-file reads per cell from 16 to 11, which is what limits the general kernel (ncu: dispatch_stall;
// profiles/r01_microbench_issue_rates.txt: the same instruction mix issues at 3.8/clk/SM with constant
// coefficients against 2.7 with register coefficients).
//  * row 1 is slot 0 of lane 0: slot 0 keeps per-lane registers B0, G0 (0 on lane 0: I~ of row 0 is 0) and an
//    addend E0 = tIM*c0 on lane 0 (the c*D~[0] term), so the rotated-in garbage of lane 31 is ignored for free;
//  * rows below the read are plain pads (prior 0);  lane 31's last row is always a pad, so the rotation feeds
//    lane 0 with M = 0;
//  * the likelihood sum is taken from row R directly: acc += M[slot] + tMI*I~[slot] with the slot a template
//    parameter (the task's R picks the loop instance; warp-uniform switch outside the loops).
// ---------------------------------------------------------------------------------------------
struct FlatCoef {
    float a, b, c, g, d;  // tMM (folded into the priors), tIM*tMI/tMM, tIM*tMD/tMM, tII (= tII*tMI/tMI), tDD
    float tmi;            // tMI: I = tMI * I~
    float tim;            // tIM/tMM: row 1 sees c = tIM (tMD_0 = 1)
    uint32_t class_id;    // reads whose read_class equals this are ours
    uint32_t qi, qd, qc;
};

struct ClassifyArgs {
    const uint8_t *rd_i, *rd_d, *rd_c;
    const uint32_t *read_off;
    uint32_t n_reads;
    uint8_t *read_class;
    uint32_t n_classes;
    uint8_t qi[MAX_FLAT_CLASSES], qd[MAX_FLAT_CLASSES], qc[MAX_FLAT_CLASSES];
    uint32_t n_sym;                  // symmetric classes: ins == del on every base, flat gcp == sym_qc[k]
    uint8_t sym_qc[MAX_SYM_CLASSES];
};

// one warp per read: flat iff every base carries the class's (ins, del, gcp) triple
__global__ void __launch_bounds__(128) phmm_classify_kernel(const ClassifyArgs a)
{
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = warp; r < a.n_reads; r += n_warps) {
        const uint32_t ro = a.read_off[r], R = a.read_off[r + 1] - ro;
        uint8_t cls = CLASS_GENERAL;
        if (R > 0) {
            const uint32_t qi0 = a.rd_i[ro], qd0 = a.rd_d[ro], qc0 = a.rd_c[ro];
            bool same = true, sym = true;
            for (uint32_t i = lane; i < R; i += 32) {
                const uint32_t qi = a.rd_i[ro + i], qd = a.rd_d[ro + i], qc = a.rd_c[ro + i];
                same = same && qi == qi0 && qd == qd0 && qc == qc0;
                sym = sym && qi == qd && qc == qc0 && qi <= SYM_MAX_GAP_QUAL;
            }
            if (__all_sync(0xffffffffu, same)) {
                for (uint32_t k = 0; k < a.n_classes; ++k)
                    if (a.qi[k] == qi0 && a.qd[k] == qd0 && a.qc[k] == qc0) cls = (uint8_t)k;
            }
            if (cls == CLASS_GENERAL && __all_sync(0xffffffffu, sym)) {
                for (uint32_t k = 0; k < a.n_sym; ++k)
                    if (a.sym_qc[k] == qc0) cls = (uint8_t)(MAX_FLAT_CLASSES + k);
            }
        }
        if (lane == 0) a.read_class[r] = cls;
    }
}

// MODE of the flat-kernel family: how many of the five transition coefficients are per-row registers instead of constant operands
constexpr int MODE_FLAT = 0;  // none: flat insertion / deletion / GCP qualities
constexpr int MODE_SYM = 1;   // two (A, C): ins == del per base with a flat GCP (the PCR indel model)
constexpr int MODE_GEN = 2;   // four (A, C, G, Dd): any per-base qualities (DRAGstr, BI/BD tags) -- the half-warp form of the general kernel

template <int K, int SLOT, bool CHECKED, int MODE>
__device__ __forceinline__ void flat_step(FastState<K> &st, const FlatCoef &f, const float (&A)[K], const float (&C)[K], const float (&G)[K],
                                          const float (&Dd)[K], float B0, float G0, float E0, float tmi, uint32_t tab_lane, int src_lane, int lane,
                                          int acc_lane, float *sums_task, float *slab, int snap_pos, int snap_slot, int end_restore, uint32_t end_out)
{
    constexpr bool SYM = MODE != MODE_FLAT;   // per-row coefficients of I^ and D^ in the match update
    constexpr bool GEN = MODE == MODE_GEN;    // ... and of the gap recurrences
    constexpr int NV = (K + 3) / 4;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t CODE_STRIDE = NV * 32 * 16;
    ++st.sp;
    const uint32_t y_next = ldg_u8(st.sp);  // (issued before the table loads: after them it measured 1 % slower)
    const float mu = __shfl_sync(FULL, st.M[K - 1], src_lane);
    const float iu = __shfl_sync(FULL, st.I[K - 1], src_lane);
    const float du = __shfl_sync(FULL, st.D[K - 1], src_lane);
    float pr[NV * 4];
    {
        const uint32_t addr = st.y * CODE_STRIDE + tab_lane;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const float4 q = lds128(addr + v * 512);
            pr[v * 4 + 0] = q.x; pr[v * 4 + 1] = q.y; pr[v * 4 + 2] = q.z; pr[v * 4 + 3] = q.w;
        }
    }
    // The coefficient of M_diag is folded into the prior table and divides the other two (see fast_step): the match
    // update is 2 FFMA + 1 FMUL.  SYM: the coefficients of I^ and D^ are per-row registers (A, C), the rest constant operands.
    float Mn[K];
    {
        // same operation order as the rows below (E0 is 0 except on the lane that holds row 1, where dgm is 0): a row's
        // arithmetic does not depend on the slot it lands on, so every lane layout gives bit-identical results
        float u = __fmaf_rn(SYM ? C[0] : f.c, st.dgd, st.dgm + E0);
        u = __fmaf_rn(B0, st.dgi, u);
        Mn[0] = pr[0] * u;
    }
#pragma unroll
    for (int k = 1; k < K; ++k) {
        float u = __fmaf_rn(SYM ? C[k] : f.c, st.D[k - 1], st.M[k - 1]);
        u = __fmaf_rn(SYM ? A[k] : f.b, st.I[k - 1], u);
        Mn[k] = pr[k] * u;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) st.D[k] = __fmaf_rn(GEN ? Dd[k] : f.d, st.D[k], st.M[k]);
    st.I[0] = __fmaf_rn(G0, iu, mu);
#pragma unroll
    for (int k = 1; k < K; ++k) st.I[k] = __fmaf_rn(GEN ? G[k] : f.g, st.I[k - 1], Mn[k - 1]);
#pragma unroll
    for (int k = 0; k < K; ++k) st.M[k] = Mn[k];
    st.dgm = mu; st.dgi = iu; st.dgd = du;
    // The END column contributes nothing to the sum by construction (prior 0), but when the next pass restores a
    // snapshot the lane above has ALREADY restored it one step ago, so the I~ chain of this END column sees foreign
    // values: the haplotype's sum is therefore the accumulator BEFORE this step.
    const float acc_before = st.acc;
    st.acc += st.M[SLOT];
    st.acc = __fmaf_rn(GEN ? tmi : f.tmi, st.I[SLOT], st.acc);
    if (CHECKED) {
        if (__builtin_expect(st.p == snap_pos, 0)) snap_save<K>(st, slab, snap_slot);
        if (__builtin_expect(st.y == CODE_END, 0)) {
            if (lane == acc_lane) sums_task[end_out] = acc_before;
            snap_restore<K>(st, slab, end_restore);  // ZERO_SLOT = fresh start
        }
        ++st.p;
    }
    st.y = y_next;
}

template <int K, int SLOT, int MODE>
__device__ __forceinline__ void flat_sweep(FastState<K> &st, const FlatCoef &f, const float (&A)[K], const float (&C)[K], const float (&G)[K],
                                           const float (&Dd)[K], float B0, float G0, float E0, float tmi, uint32_t tab_lane, int src_lane, int lane,
                                           int acc_lane, float *sums_task, float *slab, const Segment *segs, uint32_t n_segs)
{
    // the general form carries four coefficient registers per row: a shorter unrolled body keeps it at 16 resident warps
    constexpr int UNROLL = MODE == MODE_GEN ? 2 : flat_unroll(K), CHK_UNROLL = MODE == MODE_GEN ? 1 : flat_chk_unroll(K);
    int step = 1;
    for (uint32_t sg = 0; sg < n_segs; ++sg) {
        const Segment seg = segs[sg];
#pragma unroll (UNROLL)
        for (uint32_t s = 0; s < seg.n_free; ++s)
            flat_step<K, SLOT, false, MODE>(st, f, A, C, G, Dd, B0, G0, E0, tmi, tab_lane, src_lane, lane, acc_lane, sums_task, slab, 0, 0, 0, 0);
        step += (int)seg.n_free;
        st.p = step - lane;
#pragma unroll (CHK_UNROLL)
        for (uint32_t s = 0; s < seg.n_chk; ++s)
            flat_step<K, SLOT, true, MODE>(st, f, A, C, G, Dd, B0, G0, E0, tmi, tab_lane, src_lane, lane, acc_lane, sums_task, slab, seg.snap_pos,
                                           seg.snap_slot, seg.end_restore, seg.end_out);
        step += (int)seg.n_chk;
    }
}

template <int K, int SLOT, int MODE>
__device__ __forceinline__ void flat_dispatch(int slot, FastState<K> &st, const FlatCoef &f, const float (&A)[K], const float (&C)[K],
                                              const float (&G)[K], const float (&Dd)[K], float B0, float G0, float E0, float tmi,
                                              uint32_t tab_lane, int src_lane, int lane, int acc_lane, float *sums_task, float *slab,
                                              const Segment *segs, uint32_t n_segs)
{
    if (slot == SLOT) flat_sweep<K, SLOT, MODE>(st, f, A, C, G, Dd, B0, G0, E0, tmi, tab_lane, src_lane, lane, acc_lane, sums_task, slab, segs, n_segs);
    else if constexpr (SLOT + 1 < K)
        flat_dispatch<K, SLOT + 1, MODE>(slot, st, f, A, C, G, Dd, B0, G0, E0, tmi, tab_lane, src_lane, lane, acc_lane, sums_task, slab, segs, n_segs);
}

// 28 one-warp CTAs per SM (the shared-memory limit of the K=8 prior table) need <= 73 registers per thread.
// LANES = 16 (half-warp form, reads of 64..254 bases): the warp runs TWO reads of the same unit side by side, one per
// 16-lane half, K <= 16 rows per lane.  Both halves sweep the same haplotype stream with the same schedule, so every
// per-step cost that does not depend on the number of rows (hand-off shuffles, column-code load, address arithmetic,
// loop control) is paid once per 2 x 16 x K cells instead of once per 32 x K' cells with K' = K/2: 150-base reads run 10
// rows per lane instead of 5, 250-base reads 16 instead of 8.  The planner pairs reads whose last row falls on the same
// register slot (R - 1) mod K, which keeps the slot of the likelihood sum a template parameter; a read without a partner
// leaves its half idle (all rows pads).
__host__ __device__ constexpr int flat_min_ctas(int K, int MODE, int LANES) {
    return LANES == 32 ? (MODE != MODE_FLAT ? 22 : 28)
         : LANES == 8 ? (K <= 10 ? (MODE != MODE_FLAT ? 16 : 18) : K <= 16 ? 14 : (MODE != MODE_FLAT ? 12 : 14))  // quarter-warp form (20 rows: 128 / 168 registers; with 12 CTAs the flat kernel took 158 and ran 5 % slower)
         : K <= 8 ? (MODE == MODE_GEN ? 18 : MODE == MODE_SYM ? 22 : 28)  // half-warp form for reads of 64..127 bases
         : (MODE == MODE_GEN ? (K > 12 ? 10 : 14) : (K > 12 ? 14 : (MODE != MODE_FLAT ? 16 : 18)));
}

template <int K, int MODE, int LANES = 32>
__global__ void __launch_bounds__(32, flat_min_ctas(K, MODE, LANES)) phmm_flat_f32_kernel(const KernelArgs g, const FlatCoef f)
{
    constexpr bool SYM = MODE == MODE_SYM, GEN = MODE == MODE_GEN;
    constexpr int NV = (K + 3) / 4;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int REGS = snap_regs(K);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float *tab_s = reinterpret_cast<float *>(smem_raw);
    int lane, pl, src_lane;  // lane id; pipeline lane (this lane works on column step - pl); rotation source
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    if (LANES == 16) {
        asm volatile("and.b32 %0, %1, 15;" : "=r"(pl) : "r"(lane));
        asm volatile("{ .reg .u32 t, u; add.u32 t, %1, 15; and.b32 t, t, 15; and.b32 u, %1, 16; or.b32 %0, t, u; }" : "=r"(src_lane) : "r"(lane));
    } else if (LANES == 8) {
        asm volatile("and.b32 %0, %1, 7;" : "=r"(pl) : "r"(lane));
        asm volatile("{ .reg .u32 t, u; add.u32 t, %1, 7; and.b32 t, t, 7; and.b32 u, %1, 24; or.b32 %0, t, u; }" : "=r"(src_lane) : "r"(lane));
    } else {
        asm volatile("mov.u32 %0, %1;" : "=r"(pl) : "r"(lane));
        asm volatile("{ .reg .u32 t; add.u32 t, %1, 31; and.b32 %0, t, 31; }" : "=r"(src_lane) : "r"(lane));
    }
    const uint32_t tab_lane = (uint32_t)__cvta_generic_to_shared(smem_raw) + lane * 16;
    float *const sums = reinterpret_cast<float *>(g.sums);
    const uint32_t n_tasks = g.n_tasks;
    const int n_codes = g.n_codes;

    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(g.counter, 1u);
        ti = __shfl_sync(FULL, ti, 0);
        if (ti >= n_tasks) break;
        const Task t = g.tasks[ti];
        // this lane's read: the task's read, (upper half of a half-warp task) its partner, or (quarter-warp task) one of four.
        // A half-warp kernel that meets a quarter-warp task takes its reads two by two.
        // (quarter-warp tasks reach the quarter-warp kernels and, of the half-warp kernels, only the general form)
        const int part = LANES == 32 ? 0 : lane / LANES;
        const bool quad = LANES == 8 || (LANES == 16 && MODE == MODE_GEN && g.pair_tasks == 2);
        for (int sub = 0; sub < (LANES == 16 && MODE == MODE_GEN && quad ? 2 : 1); ++sub) {
        const int idx = LANES == 16 && MODE == MODE_GEN && quad ? 2 * sub + part : part;
        uint32_t rd, out_base;
        if (quad) {
            rd = idx == 0 ? t.read : idx == 1 ? t.stream_off : idx == 2 ? t.n_haps : t.hap_first;
            out_base = t.out_base + rd * t.stream_len;
        } else {
            rd = idx ? t.stream_off : t.read;
            out_base = idx ? t.stream_len : t.out_base;
        }
        const bool mine = rd != NO_READ && g.read_class[rd] == (uint8_t)f.class_id;
        if (!__any_sync(FULL, mine)) continue;
        const uint32_t ro = mine ? g.read_off[rd] : 0u;
        const int R = mine ? (int)(g.read_off[rd + 1] - ro) : 0;  // 1 <= R <= LANES * K - 1 (flat reads are never empty); 0 = idle part
        const float c0 = (float)scalbn(1.0, t.c0_exp);

        float A[K], C[K];    // SYM, GEN: per-row coefficients of I^ and D^ in the match update
        float G[K], Dd[K];   // GEN: per-row coefficients of the insertion and deletion recurrences
        float tmi_r = 0.f, tim_1 = 0.f;  // GEN: tMI of the read's last row (the sum adds tMI * I~), tIM / tMM of row 1
        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = pl * K + k + 1;
            const bool real = i <= R;
            float pmf = 0.f, pxf = 0.f;
            uint32_t x = 0;
            A[k] = 0.f; C[k] = 0.f; G[k] = 0.f; Dd[k] = 0.f;
            if (real) {
                uint32_t q = g.rd_q[ro + i - 1];
                x = g.rd_bases[ro + i - 1];
                if (q > (uint32_t)MAX_QUAL) { atomicExch(g.err, 1); q = MAX_QUAL; }
                const double e = c_eps[q];
                double pm = 1.0 - e, px = g.tristate_off ? e : e / 3.0;
                if (GEN) {
                    // any per-base qualities: the recurrence of phmm_fast_f32_kernel (I~ = I / tMI_i, D~ = D / tMD_i, tMM folded into
                    // the priors) in this kernel's layout -- row 1 on slot 0 of pipeline lane 0, pads below the read, the sum taken
                    // from row R directly
                    uint32_t qi = g.rd_i[ro + i - 1], qd = g.rd_d[ro + i - 1], qc = g.rd_c[ro + i - 1];
                    if (qi > 127u || qd > 127u || qc > 127u) { atomicExch(g.err, 1); qi = min(qi, 127u); qd = min(qd, 127u); qc = min(qc, 127u); }
                    const double ei = c_eps[qi], ec = c_eps[qc];
                    const uint32_t mn = min(qi, qd), mx = max(qi, qd);
                    const double tIM = 1.0 - ec;
                    const double Am = __ldg(g.m2m + ((mx * (mx + 1)) >> 1) + mn);
                    // tMM = 0 cannot be factored out: NaN coefficients poison the sums and the pairs are redone by the fp64 kernels
                    const double inv = Am > 0.0 ? 1.0 / Am : __longlong_as_double(0x7ff8000000000000LL);
                    double Bc = 0.0, Cc = tIM, Gc = 0.0;  // row 1: D~ of row 0 is c0 (tMD_0 = 1), I~ of row 0 is 0
                    if (i > 1) {
                        const double tmi_prev = c_eps[min((uint32_t)g.rd_i[ro + i - 2], 127u)];
                        const double tmd_prev = c_eps[min((uint32_t)g.rd_d[ro + i - 2], 127u)];
                        Bc = tIM * tmi_prev; Cc = tIM * tmd_prev; Gc = ec * tmi_prev / ei;
                    }
                    A[k] = (float)(Bc * inv); C[k] = (float)(Cc * inv); G[k] = (float)Gc; Dd[k] = (float)ec;
                    pm *= Am; px *= Am;
                } else if (SYM) {
                    // ins == del = e_i on every base, flat gcp (DESIGN.md "Symmetric-quality kernel"):
                    //   M^_i = M_i eps_{i+1}/kappa,  I^_i = I_i/kappa,  D^_i = D_i eps_{i+1}/(eps_i kappa),  eps_{R+1} := kappa
                    //   M^_i = (p_ij eps_{i+1}) (A_i M^_{i-1} + T I^_{i-1} + C_i D^_{i-1}),  A_i = tMM_i/eps_i,  C_i = T eps_{i-1}/eps_i
                    //   I^_i = M^_up + Gc I^_up,  D^_i = M^_left + Gc D^_left,  sum = M^_R + kappa I^_R.
                    // kappa = f.tmi = 2^-10 is a pure scale: the bracket above is (true value)/kappa <= 2^10 c0 < FLT_MAX.
                    const uint32_t qe = min((uint32_t)g.rd_i[ro + i - 1], 127u);
                    const double kappa = (double)f.tmi, T = (double)f.b;
                    const double eps_i = c_eps[qe];
                    const double eps_prev = i > 1 ? c_eps[min((uint32_t)g.rd_i[ro + i - 2], 127u)] : 1.0;
                    const double eps_next = i < R ? c_eps[min((uint32_t)g.rd_i[ro + i], 127u)] : kappa;
                    // with A_i factored out of the bracket:  M^_i = (p_ij eps_{i+1} A_i) (M^_diag + (T/A_i) I^_diag + (C_i/A_i) D^_diag)
                    const double Ai = __ldg(g.m2m + ((qe * (qe + 1)) >> 1) + qe) / eps_i;
                    const double inv = Ai > 0.0 ? 1.0 / Ai : __longlong_as_double(0x7ff8000000000000LL);  // tMM = 0: poison -> fp64 redo
                    A[k] = (float)(T * inv);
                    C[k] = (float)(T * eps_prev / eps_i * inv);
                    double scale = eps_next * Ai;
                    if (i == 1) scale = eps_next * T / kappa;  // row 1: the bracket is the injected constant E0 = c0
                    pm *= scale; px *= scale;
                } else {
                    pm *= (double)f.a; px *= (double)f.a;  // flat: tMM folded into the priors; f.b, f.c, f.tim arrive divided by it
                }
                pmf = (float)pm;
                pxf = (float)px;
            }
            for (int y = 0; y < n_codes; ++y) {
                float v = 0.f;
                if (real && y >= (int)CODE_FIRST_BASE) {
                    const uint32_t hb = g.code_byte[y];
                    v = (x == hb || x == (uint32_t)'N' || hb == (uint32_t)'N') ? pmf : pxf;  // LoglessPairHMM.java:89
                }
                tab_s[((y * NV + k / 4) * 32 + lane) * 4 + (k % 4)] = v;
            }
        }
        if (GEN && mine) {
            tmi_r = (float)c_eps[min((uint32_t)g.rd_i[ro + R - 1], 127u)];
            const uint32_t qi = min((uint32_t)g.rd_i[ro], 127u), qd = min((uint32_t)g.rd_d[ro], 127u), qc = min((uint32_t)g.rd_c[ro], 127u);
            const double Am = __ldg(g.m2m + ((max(qi, qd) * (max(qi, qd) + 1)) >> 1) + min(qi, qd));
            tim_1 = (float)((1.0 - c_eps[qc]) * (Am > 0.0 ? 1.0 / Am : __longlong_as_double(0x7ff8000000000000LL)));
        }
        if (!GEN && (f.qi > 127u || f.qd > 127u || f.qc > 127u)) atomicExch(g.err, 1);
        __syncwarp();

        FastState<K> st;
#pragma unroll
        for (int k = 0; k < K; ++k) { st.M[k] = 0.f; st.I[k] = 0.f; st.D[k] = 0.f; }
        st.dgm = 0.f; st.dgi = 0.f; st.dgd = 0.f;
        st.acc = 0.f;
        st.p = 0;
        const UnitSched us = g.unit_sched[t.unit];
        st.sp = g.sstreams + us.sstream_off - pl;
        st.y = ldg_u8(st.sp);
        // slot 0: pipeline lane 0 holds row 1 (virtual row 0 above it: M = I~ = 0, c*D~ = tIM*c0; SYM: f.tim = 1, the rest is in row 1's table)
        const float B0 = pl == 0 ? 0.f : (SYM || GEN ? A[0] : f.b);
        const float G0 = pl == 0 ? 0.f : (GEN ? G[0] : f.g);
        const float E0 = pl == 0 ? (GEN ? tim_1 : f.tim) * c0 : 0.f;
        const int acc_lane = mine ? (R - 1) / K : -1;
        // the slot of the likelihood sum is warp-uniform: both reads of a pair share it (planner); an idle half defers to the other
        int acc_slot = mine ? (R - 1) % K : -1;
        if (LANES == 16) {
            const int lo = __shfl_sync(FULL, acc_slot, 0), hi = __shfl_sync(FULL, acc_slot, 16);
            acc_slot = lo >= 0 ? lo : hi;
        } else if (LANES == 8) {
            int agreed = -1;
#pragma unroll
            for (int q = 0; q < 32; q += 8) {
                const int v = __shfl_sync(FULL, acc_slot, q);
                if (agreed < 0) agreed = v;
            }
            acc_slot = agreed;
        }
        float *const slab = g.snap + (size_t)blockIdx.x * ((MAX_SNAP_SLOTS + 1) * 32 * REGS) + lane * REGS;
        snap_save<K>(st, slab, ZERO_SLOT);  // the pass-start state (all zero), restored at every END that begins a fresh pass
        // 16 / 8 lanes per read: END / snapshot windows of 16 / 8 steps (the 32-step schedule is correct for any width)
        uint32_t seg_first = us.seg_first, n_segs = us.n_segs;
        if (LANES == 16 && us.n_segs16 != 0) { seg_first = us.seg16_first; n_segs = us.n_segs16; }
        if (LANES == 8 && us.n_segs8 != 0) { seg_first = us.seg8_first; n_segs = us.n_segs8; }
        flat_dispatch<K, 0, MODE>(acc_slot, st, f, A, C, G, Dd, B0, G0, E0, tmi_r, tab_lane, src_lane, pl, acc_lane, sums + out_base, slab,
                                  g.segments + seg_first, n_segs);
        }  // sub
    }
}

// ---------------------------------------------------------------------------------------------
// fp64 redo of flat-quality reads (the rescue list of a chunk).  Same formulation as phmm_flat_f32_kernel in
// double: 8 rows per lane, transition coefficients as kernel parameters (constant-bank operands of DFMA), prior
// table in shared memory (one LDS.128 = 2 rows).  A rescue task is ONE read against ONE haplotype, so there is no
// prefix sharing and no schedule: H branch-free steps, then 32 steps in which the lane holding row R writes the sum
// when it reaches the END position.  Reads of 255+ bases and non-flat reads stay with phmm_forward_kernel<double>.
// ---------------------------------------------------------------------------------------------
struct FlatCoefD {
    double a, b, c, g, d, tmi, tim;
    uint32_t class_id, qi, qd, qc;
};

__device__ __forceinline__ double2 lds128d(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}

struct F64State {
    double M[8], I[8], D[8];
    double dgm, dgi, dgd, acc;
    uint32_t y;
    const uint8_t *sp;
};

template <int SLOT, bool CHECKED>
__device__ __forceinline__ void flat64_step(F64State &st, const FlatCoefD &f, double B0, double G0, double E0, uint32_t tab_lane,
                                            int src_lane, bool is_acc_lane, int &p, int end_pos, double *out)
{
    constexpr int K = 8, NV = 4;
    constexpr unsigned FULL = 0xffffffffu;
    constexpr uint32_t CODE_STRIDE = NV * 32 * 16;
    ++st.sp;
    const uint32_t y_next = ldg_u8(st.sp);
    const double mu = __shfl_sync(FULL, st.M[K - 1], src_lane);
    const double iu = __shfl_sync(FULL, st.I[K - 1], src_lane);
    const double du = __shfl_sync(FULL, st.D[K - 1], src_lane);
    double pr[K];
    {
        const uint32_t addr = st.y * CODE_STRIDE + tab_lane;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const double2 q = lds128d(addr + v * 512);
            pr[2 * v] = q.x; pr[2 * v + 1] = q.y;
        }
    }
    double Mn[K];  // tMM is folded into the prior table and divides b, c (see fast_step): 5 DFMA-pipe instructions per cell
    {
        double u = __fma_rn(f.c, st.dgd, E0);
        u = __fma_rn(B0, st.dgi, u);
        u += st.dgm;
        Mn[0] = pr[0] * u;
    }
#pragma unroll
    for (int k = 1; k < K; ++k) {
        double u = __fma_rn(f.c, st.D[k - 1], st.M[k - 1]);
        u = __fma_rn(f.b, st.I[k - 1], u);
        Mn[k] = pr[k] * u;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) st.D[k] = __fma_rn(f.d, st.D[k], st.M[k]);
    st.I[0] = __fma_rn(G0, iu, mu);
#pragma unroll
    for (int k = 1; k < K; ++k) st.I[k] = __fma_rn(f.g, st.I[k - 1], Mn[k - 1]);
#pragma unroll
    for (int k = 0; k < K; ++k) st.M[k] = Mn[k];
    st.dgm = mu; st.dgi = iu; st.dgd = du;
    const double acc_before = st.acc;
    st.acc += st.M[SLOT];
    st.acc = __fma_rn(f.tmi, st.I[SLOT], st.acc);
    if (CHECKED) {
        if (p == end_pos && is_acc_lane) *out = acc_before;  // position based: codes after this END belong to other haplotypes
        ++p;
    }
    st.y = y_next;
}

template <int SLOT>
__device__ __forceinline__ void flat64_sweep(int slot, F64State &st, const FlatCoefD &f, double B0, double G0, double E0,
                                             uint32_t tab_lane, int src_lane, int lane, bool is_acc_lane, int H, double *out)
{
    if (slot == SLOT) {
        int p = 0;
#pragma unroll 2
        for (int s = 0; s < H; ++s) flat64_step<SLOT, false>(st, f, B0, G0, E0, tab_lane, src_lane, is_acc_lane, p, 0, out);
        p = H + 1 - lane;
#pragma unroll 1
        for (int s = 0; s < 32; ++s) flat64_step<SLOT, true>(st, f, B0, G0, E0, tab_lane, src_lane, is_acc_lane, p, H + 1, out);
    } else if constexpr (SLOT + 1 < 8) {
        flat64_sweep<SLOT + 1>(slot, st, f, B0, G0, E0, tab_lane, src_lane, lane, is_acc_lane, H, out);
    }
}

__global__ void __launch_bounds__(32) phmm_flat_f64_kernel(const KernelArgs g, const FlatCoefD f)
{
    constexpr int K = 8, NV = 4;
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *tab_s = reinterpret_cast<double *>(smem_raw);
    int lane, src_lane;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(lane));
    asm volatile("{ .reg .u32 t; add.u32 t, %1, 31; and.b32 %0, t, 31; }" : "=r"(src_lane) : "r"(lane));
    const uint32_t tab_lane = (uint32_t)__cvta_generic_to_shared(smem_raw) + lane * 16;
    double *const sums = reinterpret_cast<double *>(g.sums);
    const uint32_t n_tasks = g.n_tasks_ptr ? *g.n_tasks_ptr : g.n_tasks;
    const int n_codes = g.n_codes;

    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(g.counter, 1u);
        ti = __shfl_sync(FULL, ti, 0);
        if (ti >= n_tasks) break;
        const Task t = g.tasks[ti];
        if (g.read_class[t.read] != (uint8_t)f.class_id) continue;
        const uint32_t ro = g.read_off[t.read];
        const int R = (int)(g.read_off[t.read + 1] - ro);
        if (R > 254) continue;  // phmm_forward_kernel<double, 4, striped> takes these
        const int H = (int)t.stream_len - 1;
        const double c0 = scalbn(1.0, t.c0_exp);

        __syncwarp();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int i = lane * K + k + 1;
            const bool real = i <= R;
            double pm = 0.0, px = 0.0;
            uint32_t x = 0;
            if (real) {
                uint32_t q = g.rd_q[ro + i - 1];
                if (q > (uint32_t)MAX_QUAL) { atomicExch(g.err, 1); q = MAX_QUAL; }  // QualityUtils.java:157 (forced-fp64 mode has no fp32 pass to flag it)
                x = g.rd_bases[ro + i - 1];
                const double e = c_eps[q];
                pm = (1.0 - e) * f.a;
                px = (g.tristate_off ? e : e / 3.0) * f.a;
            }
            for (int y = 0; y < n_codes; ++y) {
                double v = 0.0;
                if (real && y >= (int)CODE_FIRST_BASE) {
                    const uint32_t hb = g.code_byte[y];
                    v = (x == hb || x == (uint32_t)'N' || hb == (uint32_t)'N') ? pm : px;  // LoglessPairHMM.java:89
                }
                tab_s[((y * NV + k / 2) * 32 + lane) * 2 + (k % 2)] = v;
            }
        }
        __syncwarp();

        F64State st;
#pragma unroll
        for (int k = 0; k < K; ++k) { st.M[k] = 0.0; st.I[k] = 0.0; st.D[k] = 0.0; }
        st.dgm = 0.0; st.dgi = 0.0; st.dgd = 0.0; st.acc = 0.0;
        st.sp = g.streams + t.stream_off - lane;
        st.y = ldg_u8(st.sp);
        const double B0 = lane == 0 ? 0.0 : f.b;
        const double G0 = lane == 0 ? 0.0 : f.g;
        const double E0 = lane == 0 ? f.tim * c0 : 0.0;
        const int acc_lane = (R - 1) / K, acc_slot = (R - 1) % K;
        flat64_sweep<0>(acc_slot, st, f, B0, G0, E0, tab_lane, src_lane, lane, lane == acc_lane, H, sums + t.out_base);
    }
}

// ---------------------------------------------------------------------------------------------
// Epilogue: raw sums -> log10 likelihoods (LoglessPairHMM.java:67), rescue-list construction.
// ---------------------------------------------------------------------------------------------
struct UnitDesc {
    uint32_t read_first, n_reads;  // chunk-local
    uint32_t hap_first, n_haps;    // chunk-local
    uint32_t out_base;             // chunk-local output slot of (read 0, hap 0)
    int32_t c0_exp;                // fp32 (or forced-fp64) initial-condition exponent of this unit
    int32_t ref_hap;               // region steps: unit-local index of the reference haplotype, -1 = none
    uint32_t keep_base;            // region steps: chunk-local slot of this unit's first keep flag
};

struct EpilogueArgs {
    const UnitDesc *units;
    uint32_t n_units;
    const uint32_t *hap_len;         // chunk-local haplotype lengths
    const uint32_t *hap_stream_off;  // chunk-local: first column of each haplotype in `streams`
    const void *sums;                // float* (fp32 pass) or double* (forced fp64)
    double *out;                     // chunk-local log10 likelihoods
    Task *rescue_tasks;              // filled when the fp32 sum is unusable
    uint32_t *n_rescue;
    uint32_t rescue_capacity;
};

__device__ __forceinline__ double log10_c0H(int c0_exp, uint32_t H) {
    return (double)c0_exp * 0.30102999566398119521 + log10((double)H);
}

// One CTA per unit; threads stride over the unit's (read, haplotype) pairs.
__global__ void __launch_bounds__(128) phmm_epilogue_f32(const EpilogueArgs e)
{
    const float *sums = reinterpret_cast<const float *>(e.sums);
    for (uint32_t u = blockIdx.x; u < e.n_units; u += gridDim.x) {
        const UnitDesc d = e.units[u];
        const uint32_t n = d.n_reads * d.n_haps;
        // gridDim.y CTAs share a unit (chunks with few units: a per-region call has one)
        for (uint32_t k = blockIdx.y * blockDim.x + threadIdx.x; k < n; k += blockDim.x * gridDim.y) {
            const uint32_t r = k / d.n_haps, h = k - r * d.n_haps;
            const float s = sums[d.out_base + k];
            const uint32_t H = e.hap_len[d.hap_first + h];
            // !(s >= thr) also catches NaN; +inf means the scaled states overflowed
            if (!(s >= RESCUE_THRESHOLD_F32) || s > 3.0e38f) {
                const uint32_t slot = atomicAdd(e.n_rescue, 1u);
                if (slot < e.rescue_capacity) {
                    Task t;
                    t.read = d.read_first + r;
                    t.stream_off = e.hap_stream_off[d.hap_first + h];
                    t.stream_len = H + 1;
                    t.out_base = slot;  // fp64 sums are indexed by rescue slot
                    int lg = 0;
                    while ((1u << lg) < H) ++lg;
                    t.c0_exp = C0_BASE_EXP_F64 - lg;
                    t.n_haps = 1;
                    t.hap_first = d.out_base + k;  // final output slot
                    t.unit = H;
                    e.rescue_tasks[slot] = t;
                }
                e.out[d.out_base + k] = __longlong_as_double(0x7ff8000000000000LL);  // placeholder NaN
            } else {
                e.out[d.out_base + k] = log10((double)s) - log10_c0H(d.c0_exp, H);
            }
        }
    }
}

// rescue pass: one double sum per rescue slot -> its final output slot.  A sum that is zero, not finite or closer than
// 2^142 to the denormal range cannot be trusted (the scaled fp64 states start 2^60 lower than Java's 2^1020): those
// pairs go on the deep list for phmm_exact_f64_kernel.
constexpr double DEEP_THRESHOLD_F64 = 0x1p-880;

__global__ void __launch_bounds__(128) phmm_epilogue_rescue(const Task *tasks, const uint32_t *n_rescue, uint32_t capacity,
                                                            const double *sums, double *out, uint32_t *n_deep, uint32_t *deep_list)
{
    const uint32_t n = min(*n_rescue, capacity);
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const Task t = tasks[k];
        const double s = sums[k];
        if (!(s >= DEEP_THRESHOLD_F64) || s > 1.0e308) {
            deep_list[atomicAdd(n_deep, 1u)] = k;
            continue;
        }
        out[t.hap_first] = log10(s) - log10_c0H(t.c0_exp, t.unit);
    }
}

// Last tier: the reference's arithmetic itself (LoglessPairHMM.java:20-68, unscaled M/I/D, initial value 2^1020/H,
// same operation order, no fused multiply-add), one thread per pair with its two DP rows in global scratch.  Only
// pairs whose likelihood is below ~1e-550 get here (Java still returns finite values down to ~1e-631); speed is
// irrelevant, agreement with Java double in its own underflow regime is the point.
struct ExactArgs {
    const Task *tasks;          // the chunk's rescue list
    const uint32_t *deep_list;  // indices into it
    const uint32_t *n_deep;
    uint32_t *cursor;
    double *scratch;            // 6 rows x row_len columns x (gridDim.x * blockDim.x) workers, column-major per worker
    uint32_t row_len;           // max haplotype length + 1
    double *out;
};

__global__ void __launch_bounds__(32) phmm_exact_f64_kernel(const KernelArgs g, const ExactArgs e)
{
    const uint32_t W = gridDim.x * blockDim.x, w = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = *e.n_deep;
    const size_t plane = (size_t)e.row_len * W;
    double *Mp = e.scratch + w, *Ip = Mp + plane, *Dp = Ip + plane, *Mc = Dp + plane, *Ic = Mc + plane, *Dc = Ic + plane;
    for (;;) {
        const uint32_t idx = atomicAdd(e.cursor, 1u);
        if (idx >= n) break;
        const Task t = e.tasks[e.deep_list[idx]];
        const uint32_t ro = g.read_off[t.read];
        const int R = (int)(g.read_off[t.read + 1] - ro), H = (int)t.unit;
        const uint8_t *hap = g.streams + t.stream_off;
        const double init = 0x1p1020 / (double)H;  // LoglessPairHMM.java:8,31
        for (int j = 0; j <= H; ++j) { Mp[(size_t)j * W] = 0.0; Ip[(size_t)j * W] = 0.0; Dp[(size_t)j * W] = init; }
        for (int i = 1; i <= R; ++i) {
            const uint32_t q = min((uint32_t)g.rd_q[ro + i - 1], (uint32_t)MAX_QUAL), qi = min((uint32_t)g.rd_i[ro + i - 1], 127u);
            const uint32_t qd = min((uint32_t)g.rd_d[ro + i - 1], 127u), qc = min((uint32_t)g.rd_c[ro + i - 1], 127u);
            const uint32_t mn = min(qi, qd), mx = max(qi, qd);
            // PairHMMModel.java:107-117
            const double tMM = g.m2m[((mx * (mx + 1)) >> 1) + mn], tIM = 1.0 - c_eps[qc], tMI = c_eps[qi], tII = c_eps[qc];
            const double tMD = c_eps[qd], tDD = c_eps[qc];
            const double err = c_eps[q], pm = 1.0 - err, px = g.tristate_off ? err : err / 3.0;
            const uint32_t x = g.rd_bases[ro + i - 1];
            Mc[0] = 0.0; Ic[0] = 0.0; Dc[0] = 0.0;
            double m_left = 0.0, d_left = 0.0;
            for (int j = 1; j <= H; ++j) {
                const uint32_t y = g.code_byte[hap[j - 1]];
                const double prior = (x == y || x == (uint32_t)'N' || y == (uint32_t)'N') ? pm : px;  // :89-91
                const size_t a = (size_t)(j - 1) * W, b = (size_t)j * W;
                // :51-55, evaluated left to right like Java, products and sums rounded separately
                const double u = __dadd_rn(__dadd_rn(__dmul_rn(Mp[a], tMM), __dmul_rn(Ip[a], tIM)), __dmul_rn(Dp[a], tIM));
                const double m = __dmul_rn(prior, u);
                const double ins = __dadd_rn(__dmul_rn(Mp[b], tMI), __dmul_rn(Ip[b], tII));
                const double del = __dadd_rn(__dmul_rn(m_left, tMD), __dmul_rn(d_left, tDD));
                Mc[b] = m; Ic[b] = ins; Dc[b] = del;
                m_left = m; d_left = del;
            }
            double *tmp;
            tmp = Mp; Mp = Mc; Mc = tmp;
            tmp = Ip; Ip = Ic; Ic = tmp;
            tmp = Dp; Dp = Dc; Dc = tmp;
        }
        double sum = 0.0;
        for (int j = 1; j <= H; ++j) sum = __dadd_rn(sum, __dadd_rn(Mp[(size_t)j * W], Ip[(size_t)j * W]));  // :62-65
        e.out[t.hap_first] = log10(sum) - 307.0505955772608;  // Math.log10(Math.pow(2, 1020)), :9
    }
}

// Measurement aid: independent FFMA chains with constant-bank multiplier and addend (kernel parameters), the operand form
// of the flat-quality forward kernel.  gphmm_measure_fp32_peak times it; bench.py reports it as roofline.peak_measured.
template <int ILP, int ITER>
__global__ void __launch_bounds__(256) phmm_ffma_peak_kernel(float *sink, const float a, const float b)
{
    float x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x * 1e-3f + i;
#pragma unroll 1
    for (int it = 0; it < ITER; it += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int i = 0; i < ILP; ++i) x[i] = __fmaf_rn(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace phmm_dev
