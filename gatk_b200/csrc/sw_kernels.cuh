// sw_kernels.cuh -- GATK's Smith-Waterman aligner on the device.
//
// Reference: utils/smithwaterman/SmithWatermanJavaAligner.java  :60-92 align() with the exact-substring shortcut,
// :105-215 calculateMatrix() (affine gaps through best-gap arrays, ties diag >= right >= down, floor -1e8),
// :262-375 calculateCigar() (end point per overhang strategy, traceback, clips).  Native counterpart in the reference:
// SmithWatermanIntelAligner (GKL) behind SWNativeAlignerWrapper.java:33-60.  Integer work: results are bit-exact.
//
// One warp per (reference, alternate) pair.  Lane l owns 8 consecutive reference rows of a 256-row strip and sweeps the
// alternate's columns one per step, lane l on column step-l (the same anti-diagonal wavefront as the PairHMM kernels):
// the row state (cell to the left, best horizontal gap and its length) stays in registers, the column state (cell
// above, best vertical gap and its length) is handed down the lanes by three shuffles, strips hand it over through a
// per-pair boundary row in global memory.  Backtrack entries are written in wavefront order (fully coalesced int16)
// and lane 0 walks them back afterwards exactly like calculateCigar.
#pragma once
#include <stdint.h>
#include <type_traits>

namespace phmm_dev {

constexpr int SW_K = 8, SW_ROWS = 32 * SW_K;   // 8 rows per lane: the per-step overhead (hand-off, boundary, stores, loop) is paid per 8 cells
template <int K> struct __align__(2 * K) SwBt { int16_t v[K]; };  // the backtrack entries of a lane's step: one 16- / 8- / 4-byte store
// The last strip of a reference runs with as few rows per lane as hold its rows (a 300-row reference leaves 44 rows for its
// second strip): 32 x rows >= remaining rows, rows in {2, 4, 8}.
__host__ __device__ constexpr int sw_last_rows(int remaining) { return remaining <= 64 ? 2 : (remaining <= 128 ? 4 : SW_K); }
constexpr int SW_SOFTCLIP = 0, SW_INDEL = 1, SW_LEADING_INDEL = 2, SW_IGNORE = 3;  // SWOverhangStrategy
constexpr uint32_t SW_OP_M = 0, SW_OP_I = 1, SW_OP_D = 2, SW_OP_S = 3;
constexpr int SW_MATRIX_MIN_CUTOFF = -100000000;   // SmithWatermanJavaAligner.java:114
constexpr int SW_LOW_INIT = INT32_MIN / 2;          // :115

struct SwTask {
    uint32_t ref_off, n_ref, alt_off, n_alt;
    uint64_t bt_off;     // first int16 of this pair's backtrack area
    uint32_t aux_off;    // first int of: last column [n_ref + 1] | bottom row [n_alt + 1] | boundary [3 * (n_alt + 1)]
    uint32_t out_off;    // first CIGAR element slot
};

struct SwArgs {
    const uint8_t *ref_bases, *alt_bases;
    const SwTask *tasks;
    uint32_t n_tasks;
    uint32_t *counter;
    int16_t *bt;
    int32_t *aux;
    uint32_t *elems;     // out: (length << 4) | op, CIGAR order
    int32_t *n_elems;    // out per pair; -1: more than `capacity` elements
    int32_t *offsets;    // out per pair: alignment offset
    uint32_t capacity;
    int32_t w_match, w_mismatch, w_open, w_extend, strategy;
};

__device__ __forceinline__ int sw_edge(int idx, bool indel, int w_open, int w_extend) {
    // sw[0][j] and sw[i][0]: 0, or the leading-gap penalties of the INDEL strategies (:125-140)
    return (!indel || idx == 0) ? 0 : w_open + (idx - 1) * w_extend;
}

// One wavefront step of a lane: its K cells of the current column (SmithWatermanJavaAligner.java:160-215).  ALL_VALID
// steps (every lane of the warp is on a column of the alternate) carry no per-cell predication; the pipeline fill and
// drain run the guarded form.  The three-way choice keeps the reference's tie rules: diagonal if it is >= both gaps,
// else the horizontal gap (insertion, -length) if it is >= the vertical one, else the vertical gap (deletion, +length).
template <int K> struct SwLane {
    int a[K], left[K], bgh[K], gsh[K];
    int last_sw, last_bgv, last_gsv;  // this lane's bottom row at its current column (what the lane below needs)
    int diag0;                        // row above at the previous column
};

template <bool ALL_VALID, int K>
__device__ __forceinline__ void sw_step(SwLane<K> &L, const bool valid, const int b, int up, int bgv, int gsv, const int w_match,
                                        const int w_mismatch, const int w_open, const int w_extend, SwBt<K> &btv)
{
    const int up_in = up;
    int diag = L.diag0;
    int16_t *btk = btv.v;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int step_diag = diag + (L.a[k] == b ? w_match : w_mismatch);
        // vertical gap ending here: open a new one from the cell above, or extend the best one of this column
        // (DPX add-max: max(a + b, c) in one instruction; "opened a new gap" = the opening won strictly)
        const int ext_v = bgv + w_extend;
        bgv = __viaddmax_s32(up, w_open, ext_v);
        const bool new_v = bgv > ext_v;
        gsv = new_v ? 1 : gsv + 1;
        // horizontal gap ending here
        const int ext_h = L.bgh[k] + w_extend;
        const int nbgh = __viaddmax_s32(L.left[k], w_open, ext_h);
        const bool new_h = nbgh > ext_h;
        const int ngsh = new_h ? 1 : L.gsh[k] + 1;
        const int gap = max(bgv, nbgh);
        const bool take_diag = step_diag >= gap;
        const int btr_gap = nbgh >= bgv ? -ngsh : gsv;
        const int cur = __vimax3_s32(step_diag, gap, SW_MATRIX_MIN_CUTOFF);
        btk[k] = (int16_t)(take_diag ? 0 : btr_gap);
        diag = L.left[k];  // sw[i][j-1] is the diagonal of the row below
        if (ALL_VALID || valid) { L.left[k] = cur; L.bgh[k] = nbgh; L.gsh[k] = ngsh; }
        up = cur;          // and this cell is "above" for the row below
    }
    if (ALL_VALID || valid) {
        L.last_sw = up; L.last_bgv = bgv; L.last_gsv = gsv;
        L.diag0 = up_in;   // row above at this column = diagonal of k = 0 at the next column
    }
}

// One strip of the matrix: 32 x K reference rows against every column of the alternate.  bt_strip = this strip's backtrack area.
template <int K>
__device__ __forceinline__ void sw_strip(const SwArgs &g, const uint8_t *__restrict__ ref, const uint8_t *__restrict__ alt, const int n_ref,
                                         const int n_alt, const int strip_row0, const bool first_strip, const bool last_strip, const bool indel,
                                         const int lane, const int n_steps, int32_t *lastcol, int32_t *bottom, int32_t *bnd, int16_t *bt_strip)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int row0 = strip_row0 + lane * K;  // row of k = 0 is row0 + 1
    SwLane<K> L;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int i = row0 + k + 1;
        L.a[k] = i <= n_ref ? (int)ref[i - 1] : 0x100;
        L.left[k] = sw_edge(i, indel, g.w_open, g.w_extend);  // sw[i][0]
        L.bgh[k] = SW_LOW_INIT; L.gsh[k] = 0;
    }
    L.last_sw = 0; L.last_bgv = SW_LOW_INIT; L.last_gsv = 0;
    L.diag0 = sw_edge(row0, indel, g.w_open, g.w_extend);  // sw[row0][0]: row above at column 0
    // the row of the last reference base (bottom row of the matrix) lives in one slot of one lane of the last strip
    const int k_last = last_strip && (n_ref - 1 - row0) >= 0 && (n_ref - 1 - row0) < K ? n_ref - 1 - row0 : -1;
    __syncwarp();
    int p = 1 - lane;
    auto step = [&](auto all_valid) {
        constexpr bool ALL = decltype(all_valid)::value;
        const bool valid = ALL || (p >= 1 && p <= n_alt);
        // row above at this column
        int up = __shfl_up_sync(FULL, L.last_sw, 1), bgv = __shfl_up_sync(FULL, L.last_bgv, 1), gsv = __shfl_up_sync(FULL, L.last_gsv, 1);
        if (lane == 0) {
            if (first_strip) { up = sw_edge(p, indel, g.w_open, g.w_extend); bgv = SW_LOW_INIT; gsv = 0; }
            else if (valid) { up = bnd[3 * p]; bgv = bnd[3 * p + 1]; gsv = bnd[3 * p + 2]; }
        }
        const int b = valid ? (int)alt[p - 1] : 0x200;
        SwBt<K> btv;
        sw_step<ALL, K>(L, valid, b, up, bgv, gsv, g.w_match, g.w_mismatch, g.w_open, g.w_extend, btv);
        if (valid) {
            const int s = p + lane;  // step number (1-based): backtrack entries are stored in wavefront order
            *reinterpret_cast<SwBt<K> *>(bt_strip + ((size_t)(s - 1) * 32 + lane) * K) = btv;
            if (!last_strip && lane == 31) { bnd[3 * p] = L.last_sw; bnd[3 * p + 1] = L.last_bgv; bnd[3 * p + 2] = L.last_gsv; }
            if (k_last >= 0) {  // bottom row, every column
                int v = L.left[0];
#pragma unroll
                for (int k = 1; k < K; ++k) v = k == k_last ? L.left[k] : v;
                bottom[p] = v;
            }
            if (p == n_alt) {   // last column, every row of this lane
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (row0 + k + 1 <= n_ref) lastcol[row0 + k + 1] = L.left[k];
            }
        }
        ++p;
    };
    // fill (some lanes are still in front of column 1), steady state, drain (some lanes are past the last column)
    int s = 1;
    for (; s <= min(31, n_steps); ++s) step(std::false_type{});
    for (; s <= n_alt; ++s) step(std::true_type{});
    for (; s <= n_steps; ++s) step(std::false_type{});
    __syncwarp();
}

__global__ void __launch_bounds__(32) phmm_sw_kernel(const SwArgs g)
{
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int K = SW_K;
    const int lane = threadIdx.x;
    const bool indel = g.strategy == SW_INDEL || g.strategy == SW_LEADING_INDEL;
    for (;;) {
        uint32_t ti = 0;
        if (lane == 0) ti = atomicAdd(g.counter, 1u);
        ti = __shfl_sync(FULL, ti, 0);
        if (ti >= g.n_tasks) break;
        const SwTask t = g.tasks[ti];
        const uint8_t *__restrict__ ref = g.ref_bases + t.ref_off;
        const uint8_t *__restrict__ alt = g.alt_bases + t.alt_off;
        const int n_ref = (int)t.n_ref, n_alt = (int)t.n_alt;
        uint32_t *out = g.elems + t.out_off;

        // ---- exact-substring shortcut (SmithWatermanJavaAligner.java:72-80, Utils.lastIndexOf) ----
        if (g.strategy == SW_SOFTCLIP || g.strategy == SW_IGNORE) {
            int found = -1;
            for (int base = n_ref - n_alt; base >= 0 && found < 0; base -= 32) {
                const int r = base - lane;
                bool ok = r >= 0;
                for (int q = 0; ok && q < n_alt; ++q) ok = ref[r + q] == alt[q];
                const unsigned hit = __ballot_sync(FULL, ok);
                if (hit) found = base - (__ffs(hit) - 1);  // the highest start wins
            }
            if (found >= 0) {
                if (lane == 0) {
                    if (g.capacity >= 1) { out[0] = ((uint32_t)n_alt << 4) | SW_OP_M; g.n_elems[ti] = 1; } else g.n_elems[ti] = -1;
                    g.offsets[ti] = found;
                }
                continue;
            }
        }

        // ---- the matrix ----
        int32_t *lastcol = g.aux + t.aux_off, *bottom = lastcol + (n_ref + 1), *bnd = bottom + (n_alt + 1);
        int16_t *bt = g.bt + t.bt_off;
        const int n_strips = (n_ref + SW_ROWS - 1) / SW_ROWS, n_steps = n_alt + 31;
        const int k_tail = sw_last_rows(n_ref - (n_strips - 1) * SW_ROWS);  // rows per lane of the last strip
        for (int strip = 0; strip < n_strips; ++strip) {
            const bool first_strip = strip == 0, last_strip = strip == n_strips - 1;
            int16_t *bt_strip = bt + (size_t)strip * n_steps * SW_ROWS;
            if (!last_strip || k_tail == SW_K)
                sw_strip<SW_K>(g, ref, alt, n_ref, n_alt, strip * SW_ROWS, first_strip, last_strip, indel, lane, n_steps, lastcol, bottom, bnd, bt_strip);
            else if (k_tail == 4)
                sw_strip<4>(g, ref, alt, n_ref, n_alt, strip * SW_ROWS, first_strip, last_strip, indel, lane, n_steps, lastcol, bottom, bnd, bt_strip);
            else
                sw_strip<2>(g, ref, alt, n_ref, n_alt, strip * SW_ROWS, first_strip, last_strip, indel, lane, n_steps, lastcol, bottom, bnd, bt_strip);
        }
        __threadfence_block();
        __syncwarp();

        // ---- calculateCigar (:262-375) ----
        // Every lane carries the same traceback state; lane 0 writes.  The walk is a chain of dependent loads (one
        // backtrack entry tells where the next one is), so the warp reads 32 entries along the diagonal at once and
        // consumes the leading run of zeros (diagonal moves) in one go: an alignment is mostly long match runs.
        {
            auto BT = [&](int i, int j) -> int {
                const int strip = (i - 1) / SW_ROWS, r = (i - 1) % SW_ROWS, ks = strip == n_strips - 1 ? k_tail : SW_K;
                const int ln = r / ks, k = r % ks;
                return (int)bt[(size_t)strip * n_steps * SW_ROWS + ((size_t)(j + ln - 1) * 32 + ln) * ks + k];
            };
            int p1 = 0, p2 = 0, maxscore = INT32_MIN, segment_length = 0;
            if (g.strategy == SW_INDEL) {
                p1 = n_ref; p2 = n_alt;
            } else {
                p2 = n_alt;
                // last column: the LAST row holding the maximum (`cur >= maxscore`, :283-289)
                int best = INT32_MIN, bi = 0;
                for (int i = 1 + lane; i <= n_ref; i += 32) {
                    const int cur = lastcol[i];
                    if (cur >= best) { best = cur; bi = i; }
                }
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    const int ob = __shfl_xor_sync(FULL, best, off), oi = __shfl_xor_sync(FULL, bi, off);
                    if (ob > best || (ob == best && oi > bi)) { best = ob; bi = oi; }
                }
                p1 = bi; maxscore = best;
                if (g.strategy != SW_LEADING_INDEL) {
                    // bottom row (:291-300): a column replaces the incumbent when it scores higher, or equally with a
                    // smaller |n_ref - j| -- i.e. the result is the FIRST candidate (incumbent, then j = 1, 2, ...) that
                    // nothing after it beats strictly: per lane the first best of its columns, then the same rule across lanes
                    int bs = INT32_MIN, bd = INT32_MAX, bj = 0;
                    for (int j = 1 + lane; j <= n_alt; j += 32) {
                        const int cur = bottom[j], d = abs(n_ref - j);
                        if (cur > bs || (cur == bs && d < bd)) { bs = cur; bd = d; bj = j; }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const int os = __shfl_xor_sync(FULL, bs, off), od = __shfl_xor_sync(FULL, bd, off), oj = __shfl_xor_sync(FULL, bj, off);
                        if (oj != 0 && (bj == 0 || os > bs || (os == bs && (od < bd || (od == bd && oj < bj))))) { bs = os; bd = od; bj = oj; }
                    }
                    if (bj != 0 && (bs > maxscore || (bs == maxscore && bd < abs(p1 - p2)))) {
                        p1 = n_ref; p2 = bj; maxscore = bs; segment_length = n_alt - bj;
                    }
                }
            }
            uint32_t n = 0;
            bool overflow = false;
            auto push = [&](uint32_t op, int len) {
                if (n < g.capacity) { if (lane == 0) out[n] = ((uint32_t)len << 4) | op; } else overflow = true;
                ++n;
            };
            if (segment_length > 0 && g.strategy == SW_SOFTCLIP) { push(SW_OP_S, segment_length); segment_length = 0; }
            uint32_t state = SW_OP_M;
            do {
                // entries along the diagonal from (p1, p2); outside the matrix: a non-zero sentinel that ends the run
                const int q1 = p1 - lane, q2 = p2 - lane;
                int my = 1;
                if (q1 > 0 && q2 > 0) my = BT(q1, q2);
                const unsigned nz = __ballot_sync(FULL, my != 0);
                const int run = nz ? __ffs(nz) - 1 : 32;
                if (run > 0) {  // `run` diagonal steps (state MATCH, length 1 each)
                    if (state == SW_OP_M) segment_length += run;
                    else {
                        if (segment_length > 0) push(state, segment_length);
                        segment_length = run;
                        state = SW_OP_M;
                    }
                    p1 -= run; p2 -= run;
                }
                if (run < 32 && p1 > 0 && p2 > 0) {  // the entry that ended the run is a gap
                    const int btr = __shfl_sync(FULL, my, run);
                    const uint32_t new_state = btr > 0 ? SW_OP_D : SW_OP_I;
                    const int step_length = btr > 0 ? btr : -btr;
                    if (new_state == SW_OP_I) p2 -= step_length; else p1 -= step_length;
                    if (new_state == state) segment_length += step_length;
                    else {
                        if (segment_length > 0) push(state, segment_length);
                        segment_length = step_length;
                        state = new_state;
                    }
                }
            } while (p1 > 0 && p2 > 0);
            int offset;
            if (g.strategy == SW_SOFTCLIP) {
                push(state, segment_length);
                if (p2 > 0) push(SW_OP_S, p2);
                offset = p1;
            } else if (g.strategy == SW_IGNORE) {
                push(state, segment_length + p2);
                offset = p1 - p2;
            } else {
                push(state, segment_length);
                if (p1 > 0) push(SW_OP_D, p1); else if (p2 > 0) push(SW_OP_I, p2);
                offset = 0;
            }
            if (lane == 0) {
                if (overflow) {
                    g.n_elems[ti] = -1;
                } else {
                    for (uint32_t x = 0, y = n - 1; x < y; ++x, --y) { const uint32_t tmp = out[x]; out[x] = out[y]; out[y] = tmp; }  // Lists.reverse (:374)
                    g.n_elems[ti] = (int32_t)n;
                }
                g.offsets[ti] = offset;
            }
        }
        __syncwarp();
    }
}

}  // namespace phmm_dev
