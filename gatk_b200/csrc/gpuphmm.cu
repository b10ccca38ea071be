// gpuphmm.cu -- host side of libgpuphmm.so: C ABI (include/gpuphmm.h), chunk planner, pinned staging,
// per-device stream slots, multi-GPU work queue, async submit/wait queue.
//
// Mirrors the role of the native library behind PairHMMNativeBinding in the reference
// (call sites: src/main/java/org/broadinstitute/hellbender/utils/pairhmm/VectorLoglessPairHMM.java:63,81,138,164).
// No CPU compute path exists here: every likelihood comes from the CUDA kernels in phmm_kernels.cuh.
#include "../../include/gpuphmm.h"
#include "phmm_kernels.cuh"
#include "region_steps.cuh"
#include "pdhmm_kernels.cuh"
#include "sw_kernels.cuh"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

using namespace phmm_dev;

namespace {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            char buf_[512];                                                                           \
            snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,          \
                     cudaGetErrorString(e_));                                                         \
            throw Error(e_ == cudaErrorMemoryAllocation ? GPHMM_ERR_NOMEM : GPHMM_ERR_CUDA, buf_);    \
        }                                                                                             \
    } while (0)

double now_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

// ---- tables (PairHMMModel.java:86-94, QualityUtils.java:51-57, MathUtils.java:406-423,467-480) ----
struct Tables {
    std::vector<double> eps;  // 256 entries, [255] unused
    std::vector<double> m2m;  // triangular, (MAX_QUAL+1)(MAX_QUAL+2)/2
    Tables() : eps(256, 0.0), m2m(((MAX_QUAL + 1) * (MAX_QUAL + 2)) >> 1) {
        for (int q = 0; q <= MAX_QUAL; ++q) eps[q] = std::pow(10.0, (double)q / -10.0);
        // Jacobian-logarithm table: step 1e-4, cut-off 8.0 log10 units
        const double step = 0.0001, inv_step = 1.0 / step, tol = 8.0;
        const int n = (int)(tol / step) + 1;
        std::vector<double> jac(n);
        for (int k = 0; k < n; ++k) jac[k] = std::log10(1.0 + std::pow(10.0, -k * step));
        auto approx_sum = [&](double a, double b) {
            if (a > b) std::swap(a, b);
            if (a == -INFINITY) return b;
            const double diff = b - a;
            if (!(diff < tol)) return b;
            const double d = diff * inv_step;
            const int idx = d > 0.0 ? (int)(d + 0.5) : (int)(d - 0.5);
            return b + jac[idx];
        };
        const double inv_ln10 = 1.0 / std::log(10.0);
        for (int i = 0, off = 0; i <= MAX_QUAL; off += ++i)
            for (int j = 0; j <= i; ++j) {
                const double ls = approx_sum(-0.1 * i, -0.1 * j);
                const double l10 = std::log1p(-std::min(1.0, std::pow(10.0, ls))) * inv_ln10;
                m2m[off + j] = std::pow(10.0, l10);
            }
    }
};
const Tables &tables() {
    static Tables t;
    return t;
}

// ---- a growable device / pinned buffer ----
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) CK(cudaFree(p));
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)1 << 16);
        want += want / 4;
        CK(cudaMalloc(&p, want));
        cap = want;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    void reserve(size_t n) {
        if (n <= cap) return;
        if (p) CK(cudaFreeHost(p));
        p = nullptr; cap = 0;
        size_t want = std::max(n, (size_t)1 << 16);
        want += want / 4;
        CK(cudaMallocHost(&p, want));
        cap = want;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

constexpr int N_SLOTS = 3;         // chunks in flight per device: one computing, one finishing (epilogue/D2H), one being staged
constexpr int N_FP32_BUCKETS = 9;  // K = 1..8 plain, bucket 8 = striped K=8 (reads of 256+ bases)
constexpr int N_AUX = N_FP32_BUCKETS + 8 * (MAX_FLAT_CLASSES + MAX_SYM_CLASSES);  // side streams: general buckets + flat (class, bucket)
constexpr int N_COUNTERS = 128;    // [0..8] general fp32 buckets, [9] fp64 queue, [10] n_rescue, [11] n_deep, [120] deep cursor,
                                   // [16 + 8*c + k] flat-quality kernels: class c, bucket k

// Host-side plan of one device chunk: which units, and every metadata array the kernels need.
struct ChunkPlan {
    int64_t u0 = 0, u1 = 0;          // unit range in the batch
    int64_t r_lo = 0, r_hi = 0;      // read span in the batch
    int64_t base_lo = 0, base_hi = 0;  // byte span of the five per-base arrays
    uint32_t n_pairs = 0;
    int64_t cells = 0;
    uint32_t max_stream_len = 0, max_hap_len = 0;
    int n_codes = 7;
    uint8_t code_byte[MAX_CODES];
    std::vector<uint32_t> read_off;        // span-local, n_reads_span + 1
    std::vector<uint8_t> streams;
    std::vector<uint32_t> hap_len, hap_stream_off;
    std::vector<UnitDesc> units;
    std::vector<Task> tasks;               // bucket-sorted
    uint32_t bucket_begin[N_FP32_BUCKETS + 1];
    std::vector<uint8_t> sstreams;         // prefix-compressed streams of the fast kernels
    std::vector<PassInfo> pass_info;
    std::vector<Segment> segments;
    std::vector<UnitSched> unit_sched;
    int64_t skipped_cells = 0;
    int64_t computed_columns = 0;          // haplotype columns the fast kernels really sweep (after prefix sharing)
    int n_classes = 0;                     // flat-quality classes sampled from the chunk's reads
    uint8_t class_qi[MAX_FLAT_CLASSES], class_qd[MAX_FLAT_CLASSES], class_qc[MAX_FLAT_CLASSES];
    uint32_t n_keep = 0;                   // region steps: keep flags of the chunk (sum of n_reads over its units)
    int n_sym = 0;                         // symmetric-quality classes (ins == del per base, flat gcp): their gcp values
    uint8_t sym_qc[MAX_SYM_CLASSES];
};

int ceil_log2(uint32_t v) {
    int lg = 0;
    while ((1u << lg) < v) ++lg;
    return lg;
}

void validate_batch(const gphmm_batch *b) {
    if (!b) throw Error(GPHMM_ERR_INVALID_ARG, "batch is null");
    if (b->n_units < 0 || b->n_reads < 0 || b->n_haps < 0) throw Error(GPHMM_ERR_INVALID_ARG, "negative count");
    if (b->n_units == 0) return;
    if (!b->units || !b->read_off || !b->hap_off) throw Error(GPHMM_ERR_INVALID_ARG, "null offsets/units");
    for (int64_t r = 0; r < b->n_reads; ++r)
        if (b->read_off[r + 1] < b->read_off[r]) throw Error(GPHMM_ERR_INVALID_ARG, "read_off not monotone");
    for (int64_t h = 0; h < b->n_haps; ++h)
        if (b->hap_off[h + 1] <= b->hap_off[h])
            throw Error(GPHMM_ERR_INVALID_ARG, "zero-length haplotype (PairHMM.initialize requires haplotypeMaxLength > 0)");
    for (int64_t u = 0; u < b->n_units; ++u) {
        const gphmm_unit &un = b->units[u];
        if (un.read_begin < 0 || un.read_end < un.read_begin || un.read_end > b->n_reads || un.hap_begin < 0 ||
            un.hap_end < un.hap_begin || un.hap_end > b->n_haps || un.out_off < 0)
            throw Error(GPHMM_ERR_INVALID_ARG, "unit range out of bounds");
        if (un.hap_end - un.hap_begin > 65535) throw Error(GPHMM_ERR_TOO_LARGE, "more than 65535 haplotypes in one unit");
    }
    if (b->n_reads > 0 && b->read_off[b->n_reads] > 0 &&
        (!b->read_bases || !b->base_q || !b->ins_q || !b->del_q || !b->gcp))
        throw Error(GPHMM_ERR_INVALID_ARG, "null read array");
    if (b->n_haps > 0 && !b->hap_bases) throw Error(GPHMM_ERR_INVALID_ARG, "null hap_bases");
}

// Greedy split of the unit list into chunks bounded by cells and staged bytes.
std::vector<std::pair<int64_t, int64_t>> split_units(const gphmm_batch *b, int64_t chunk_cells, int64_t chunk_bytes, bool ramp_up) {
    std::vector<std::pair<int64_t, int64_t>> out;
    int64_t u = 0;
    const int64_t full_cells = chunk_cells;
    while (u < b->n_units) {
        // ramp up: the first chunks are small so that the GPU starts early while the host is still staging
        const size_t ci = out.size();
        // (2.5e8 cells = a handful of regions, then doubling: the planner threads stay ahead of the GPU from there on)
        chunk_cells = (!ramp_up || ci >= 16) ? full_cells : std::min<int64_t>(full_cells, (int64_t)250000000 << ci);
        int64_t cells = 0, bytes = 0, pairs = 0, u_end = u;
        int64_t r_lo = INT64_MAX, r_hi = 0;
        while (u_end < b->n_units) {
            const gphmm_unit &un = b->units[u_end];
            const int64_t nr = un.read_end - un.read_begin, nh = un.hap_end - un.hap_begin;
            const int64_t rb = nr ? b->read_off[un.read_end] - b->read_off[un.read_begin] : 0;
            const int64_t hb = nh ? b->hap_off[un.hap_end] - b->hap_off[un.hap_begin] : 0;
            const int64_t n_lo = std::min(r_lo, nr ? un.read_begin : r_lo), n_hi = std::max(r_hi, nr ? un.read_end : r_hi);
            const int64_t span = n_hi > n_lo ? b->read_off[n_hi] - b->read_off[n_lo] : 0;
            const int64_t c = rb * hb;
            if (u_end > u && (cells + c > chunk_cells || span * 5 + bytes + hb > chunk_bytes || pairs + nr * nh > (int64_t)1 << 25))
                break;
            cells += c; bytes += hb + nh; pairs += nr * nh;
            r_lo = n_lo; r_hi = n_hi;
            ++u_end;
        }
        const int64_t span = r_hi > r_lo ? b->read_off[r_hi] - b->read_off[r_lo] : 0;
        if (span >= ((int64_t)1 << 31) || bytes >= ((int64_t)1 << 31) || pairs >= ((int64_t)1 << 28))
            throw Error(GPHMM_ERR_TOO_LARGE, "a single unit exceeds the per-chunk device budget");
        out.emplace_back(u, u_end);
        u = u_end;
    }
    return out;
}

// Plans the shared (prefix-compressed) stream of one unit: haplotypes sorted lexicographically, pass i+1 resumes
// from a snapshot taken at the last column it shares with its predecessors.  Appends to c.sstreams / pass_info /
// segments and returns the unit's schedule.  With share == false every pass starts from column 1 in input order.
// Lexicographic order of a unit's haplotypes (input order when sharing is off).
std::vector<int> sorted_hap_order(const gphmm_batch *b, const gphmm_unit &un, bool share) {
    const int n = (int)(un.hap_end - un.hap_begin);
    auto hap_ptr = [&](int k) { return b->hap_bases + b->hap_off[un.hap_begin + k]; };
    auto hap_len = [&](int k) { return (uint32_t)(b->hap_off[un.hap_begin + k + 1] - b->hap_off[un.hap_begin + k]); };
    std::vector<int> order(n);
    for (int k = 0; k < n; ++k) order[k] = k;
    if (share)
        std::sort(order.begin(), order.end(), [&](int x, int y) {
            const uint32_t lx = hap_len(x), ly = hap_len(y);
            const int cmp = memcmp(hap_ptr(x), hap_ptr(y), std::min(lx, ly));
            if (cmp != 0) return cmp < 0;
            if (lx != ly) return lx < ly;
            return x < y;
        });
    return order;
}

UnitSched plan_unit_sharing(const gphmm_batch *b, const gphmm_unit &un, uint32_t hap_first_local, bool share, ChunkPlan &c,
                            int64_t sum_read_len, const std::vector<int> &full_order, int g_first, int g_count) {
    constexpr uint32_t SPACING = 32;
    static const uint32_t MIN_DEPTH = getenv("GPHMM_MIN_DEPTH") ? (uint32_t)std::max(32, atoi(getenv("GPHMM_MIN_DEPTH"))) : 32u;  // tuning knob
    UnitSched us;
    memset(&us, 0, sizeof us);
    const int n = g_count;  // haplotypes of this group: full_order[g_first .. g_first + g_count)
    us.pass_first = (uint32_t)c.pass_info.size();
    us.n_passes = (uint32_t)n;
    us.seg_first = (uint32_t)c.segments.size();
    c.sstreams.insert(c.sstreams.end(), STREAM_PAD, (uint8_t)CODE_NULL);
    us.sstream_off = (uint32_t)c.sstreams.size();
    if (n == 0) return us;
    auto hap_ptr = [&](int k) { return b->hap_bases + b->hap_off[un.hap_begin + k]; };
    auto hap_len = [&](int k) { return (uint32_t)(b->hap_off[un.hap_begin + k + 1] - b->hap_off[un.hap_begin + k]); };
    std::vector<int> order(full_order.begin() + g_first, full_order.begin() + g_first + g_count);
    struct Snap { int pass; uint32_t depth, pos; int slot; uint32_t free_after; };
    std::vector<Snap> snaps;
    int slot_owner[MAX_SNAP_SLOTS];
    for (int k = 0; k < MAX_SNAP_SLOTS; ++k) slot_owner[k] = -1;
    std::vector<uint32_t> r(n, 0), pass_start(n + 1, 1), end_pos(n, 0);
    std::vector<int> snap_of_pass(n, -1);
    std::vector<uint32_t> lcp(n, 0), n_pad(n, 0);
    for (int i = 0; i < n; ++i) {
        const uint32_t H = hap_len(order[i]);
        // every pass spans at least 32 stream positions (NULL columns before its END if it is shorter), so that the
        // 32-step END windows of consecutive passes never overlap
        n_pad[i] = (H - r[i]) < 32u ? 32u - (H - r[i]) : 0u;
        end_pos[i] = pass_start[i] + (H - r[i]) + n_pad[i];
        pass_start[i + 1] = end_pos[i] + 1;
        if (i + 1 >= n || !share) continue;
        // longest common prefix with the next haplotype in sorted order
        const uint8_t *x = hap_ptr(order[i]), *y = hap_ptr(order[i + 1]);
        const uint32_t m = std::min(H, hap_len(order[i + 1]));
        uint32_t d = 0;
        while (d < m && x[d] == y[d]) ++d;
        lcp[i] = d;
        if (d < MIN_DEPTH) continue;
        int k = -1;
        {
            // preferred: a snapshot at exactly the shared depth, taken by the latest pass that computed column d
            int j = i;
            while (r[j] >= d) --j;  // r[0] = 0 < d
            const uint32_t pos = pass_start[j] + (d - r[j]) - 1;
            for (size_t q = 0; q < snaps.size(); ++q)
                if (snaps[q].pass == j && snaps[q].depth == d && slot_owner[snaps[q].slot] == (int)q) k = (int)q;
            if (k < 0) {
                bool ok = true;
                for (const Snap &sn : snaps) ok = ok && (sn.pos + SPACING <= pos || pos + SPACING <= sn.pos);
                int slot = -1;
                for (int q = 0; q < MAX_SNAP_SLOTS && ok && slot < 0; ++q)
                    if (slot_owner[q] < 0 || snaps[slot_owner[q]].free_after < pos) slot = q;
                if (ok && slot >= 0) {
                    snaps.push_back({j, d, pos, slot, 0});
                    k = (int)snaps.size() - 1;
                    slot_owner[slot] = k;
                }
            }
        }
        if (k < 0) {
            // fallback: the deepest live snapshot whose prefix the next haplotype still shares
            uint32_t best = 0;
            for (size_t q = 0; q < snaps.size(); ++q) {
                if (slot_owner[snaps[q].slot] != (int)q || snaps[q].depth < MIN_DEPTH || snaps[q].depth <= best) continue;
                uint32_t shared_len = UINT32_MAX;  // LCP(haplotype of snaps[q].pass, haplotype i+1) = min lcp[pass..i]
                for (int t = snaps[q].pass; t <= i; ++t) shared_len = std::min(shared_len, lcp[t]);
                if (shared_len >= snaps[q].depth) { best = snaps[q].depth; k = (int)q; }
            }
            if (k < 0) continue;
        }
        snaps[k].free_after = end_pos[i];  // restored at the END column of pass i
        r[i + 1] = snaps[k].depth;
        snap_of_pass[i + 1] = k;
    }
    // stream + pass table
    for (int i = 0; i < n; ++i) {
        const uint32_t H = hap_len(order[i]);
        const size_t w = c.sstreams.size();
        c.sstreams.resize(w + (H - r[i]) + n_pad[i] + 1);
        uint8_t *dst = c.sstreams.data() + w;
        // the full stream of this unit was encoded a moment ago: copy the columns behind the shared prefix
        memcpy(dst, c.streams.data() + c.hap_stream_off[hap_first_local + order[i]] + r[i], H - r[i]);
        for (uint32_t q = 0; q < n_pad[i]; ++q) dst[H - r[i] + q] = (uint8_t)CODE_NULL;  // prior 0: no effect on the sum
        dst[H - r[i] + n_pad[i]] = (uint8_t)CODE_END;
        PassInfo pi;
        pi.out_idx = (uint16_t)order[i];
        pi.restore_slot = (int16_t)(snap_of_pass[i] >= 0 ? snaps[snap_of_pass[i]].slot : -1);
        c.pass_info.push_back(pi);
        c.skipped_cells += (int64_t)r[i] * sum_read_len;
        c.computed_columns += H - r[i];
    }
    // schedule.  Lane l meets stream position q at step q + l, so an END column at e keeps some lane busy with it
    // during steps [e, e+32) and a snapshot position s during [s, s+32).  END windows never overlap each other
    // (passes span >= 32 positions), snapshot windows never overlap each other (SPACING), so at any step at most one
    // of each is active.  A segment = branch-free steps, then checked steps with one constant (END, snapshot) pair.
    std::vector<uint32_t> pts;
    pts.push_back(1);
    for (int i = 0; i < n; ++i) { pts.push_back(end_pos[i]); pts.push_back(end_pos[i] + 32); }
    for (const Snap &sn : snaps) { pts.push_back(sn.pos); pts.push_back(sn.pos + 32); }
    std::sort(pts.begin(), pts.end());
    pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
    std::vector<int> snap_by_pos(snaps.size());
    for (size_t q = 0; q < snaps.size(); ++q) snap_by_pos[q] = (int)q;
    std::sort(snap_by_pos.begin(), snap_by_pos.end(), [&](int x, int y) { return snaps[x].pos < snaps[y].pos; });
    Segment seg;
    auto clear_seg = [&]() { seg.n_free = 0; seg.n_chk = 0; seg.snap_pos = INT32_MIN; seg.snap_slot = 0; seg.end_restore = MAX_SNAP_SLOTS; seg.end_out = 0; };
    clear_seg();
    size_t ie = 0, is = 0;  // first END / snapshot whose window has not expired yet
    for (size_t k = 0; k + 1 < pts.size(); ++k) {
        const uint32_t x = pts[k], y = pts[k + 1];
        while (ie < (size_t)n && end_pos[ie] + 32 <= x) ++ie;
        while (is < snaps.size() && snaps[snap_by_pos[is]].pos + 32 <= x) ++is;
        const bool end_on = ie < (size_t)n && end_pos[ie] <= x;
        const bool snap_on = is < snaps.size() && snaps[snap_by_pos[is]].pos <= x;
        static const bool all_checked = getenv("GPHMM_ALL_CHECKED") != nullptr;  // experiment: cost of the checked loop
        if (!end_on && !snap_on && !all_checked) {
            if (seg.n_chk) { c.segments.push_back(seg); clear_seg(); }
            seg.n_free += y - x;
            continue;
        }
        if (seg.n_chk) { c.segments.push_back(seg); clear_seg(); }
        seg.n_chk = y - x;
        if (snap_on) { seg.snap_pos = (int32_t)snaps[snap_by_pos[is]].pos; seg.snap_slot = (uint8_t)snaps[snap_by_pos[is]].slot; }
        if (end_on) {
            seg.end_out = (uint16_t)order[ie];
            seg.end_restore = (int8_t)(ie + 1 < (size_t)n && snap_of_pass[ie + 1] >= 0 ? snaps[snap_of_pass[ie + 1]].slot : MAX_SNAP_SLOTS);  // MAX_SNAP_SLOTS = pass-start state
        }
    }
    if (seg.n_free || seg.n_chk) c.segments.push_back(seg);
    us.n_segs = (uint32_t)c.segments.size() - us.seg_first;
    return us;
}

// When a chunk has too few reads to fill the GPU with one warp per read (a single HaplotypeCaller region is ~100 reads),
// each unit's haplotypes are split into groups and every (read, group) pair becomes a task of its own.
constexpr int64_t TARGET_TASKS = 148 * 28 * 2;

void plan_chunk(const gphmm_batch *b, int64_t u0, int64_t u1, bool force_fp64, bool share, ChunkPlan &c, bool pcr_hint = false) {
    c.u0 = u0; c.u1 = u1;
    c.r_lo = INT64_MAX; c.r_hi = 0;
    for (int64_t u = u0; u < u1; ++u) {
        const gphmm_unit &un = b->units[u];
        if (un.read_end > un.read_begin) { c.r_lo = std::min(c.r_lo, un.read_begin); c.r_hi = std::max(c.r_hi, un.read_end); }
    }
    if (c.r_hi <= c.r_lo) { c.r_lo = c.r_hi = 0; }
    c.base_lo = b->n_reads ? b->read_off[c.r_lo] : 0;
    c.base_hi = b->n_reads ? b->read_off[c.r_hi] : 0;
    const int64_t n_span = c.r_hi - c.r_lo;
    c.read_off.resize(n_span + 1);
    for (int64_t r = 0; r <= n_span; ++r) c.read_off[r] = (uint32_t)(b->read_off[c.r_lo + r] - c.base_lo);

    // quality classes from a sample of the chunk's reads: flat (one (ins, del, gcp) triple on every base) and
    // symmetric (ins == del per base, flat gcp).  The device decides per read which class it really belongs to
    // (phmm_classify_kernel); a class that is missed here only means those reads take the general kernel.
    c.n_classes = 0;
    c.n_sym = 0;
    if (!force_fp64 && n_span > 0) {
        const int64_t stride = std::max<int64_t>(1, n_span / 256);
        for (int64_t r = 0; r < n_span; r += stride) {
            const int64_t o = b->read_off[c.r_lo + r], e = b->read_off[c.r_lo + r + 1];
            if (e == o) continue;
            const uint8_t qi = b->ins_q[o], qd = b->del_q[o], qc = b->gcp[o];
            if (qi > 127 || qd > 127 || qc > 127) continue;
            // an array is constant iff it equals itself shifted by one (memcmp is vectorised)
            const size_t n1 = (size_t)(e - o - 1);
            const bool flat_c = memcmp(b->gcp + o, b->gcp + o + 1, n1) == 0;
            const bool flat = flat_c && memcmp(b->ins_q + o, b->ins_q + o + 1, n1) == 0 && memcmp(b->del_q + o, b->del_q + o + 1, n1) == 0;
            bool sym = flat_c && memcmp(b->ins_q + o, b->del_q + o, n1 + 1) == 0;
            if (sym && !flat) {
                uint8_t mx = 0;
                for (int64_t i = o; i < e; ++i) mx = std::max(mx, b->ins_q[i]);
                sym = mx <= SYM_MAX_GAP_QUAL;
            } else if (sym) {
                sym = qi <= SYM_MAX_GAP_QUAL;
            }
            if (flat && !(pcr_hint && qi == qd)) {  // the PCR indel model (region steps) will lower ins and del together
                bool seen = false;
                for (int k = 0; k < c.n_classes; ++k) seen = seen || (c.class_qi[k] == qi && c.class_qd[k] == qd && c.class_qc[k] == qc);
                if (seen) continue;
                if (c.n_classes < MAX_FLAT_CLASSES) {
                    c.class_qi[c.n_classes] = qi; c.class_qd[c.n_classes] = qd; c.class_qc[c.n_classes] = qc; ++c.n_classes;
                    continue;
                }
            }
            if (sym) {  // includes flat reads that found no free flat class
                bool seen = false;
                for (int k = 0; k < c.n_sym; ++k) seen = seen || c.sym_qc[k] == qc;
                if (!seen && c.n_sym < MAX_SYM_CLASSES) c.sym_qc[c.n_sym++] = qc;
            }
        }
    }

    // haplotype alphabet of the chunk: A C G T N are fixed codes, any other byte value gets the next free code
    int16_t lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = -1;
    memset(c.code_byte, 0, sizeof c.code_byte);
    const char fixed[5] = {'A', 'C', 'G', 'T', 'N'};
    for (int i = 0; i < 5; ++i) { lut[(uint8_t)fixed[i]] = (int16_t)(CODE_FIRST_BASE + i); c.code_byte[CODE_FIRST_BASE + i] = (uint8_t)fixed[i]; }
    c.n_codes = CODE_FIRST_BASE + 5;

    c.streams.clear(); c.hap_len.clear(); c.hap_stream_off.clear(); c.units.clear(); c.tasks.clear();
    c.sstreams.clear(); c.pass_info.clear(); c.segments.clear(); c.unit_sched.clear(); c.skipped_cells = 0; c.computed_columns = 0; c.n_keep = 0;
    c.n_pairs = 0; c.cells = 0; c.max_stream_len = 0; c.max_hap_len = 0;
    int64_t n_reads_with_work = 0;
    for (int64_t u = u0; u < u1; ++u)
        if (b->units[u].hap_end > b->units[u].hap_begin) n_reads_with_work += b->units[u].read_end - b->units[u].read_begin;
    const int64_t want_groups = n_reads_with_work > 0 ? (TARGET_TASKS + n_reads_with_work - 1) / n_reads_with_work : 1;
    std::vector<Task> raw;
    std::vector<uint8_t> bucket_of;
    raw.reserve((size_t)(c.r_hi - c.r_lo));
    bucket_of.reserve((size_t)(c.r_hi - c.r_lo));
    c.streams.reserve((size_t)(u1 - u0) * 64);
    uint32_t bucket_count[N_FP32_BUCKETS] = {0};
    for (int64_t u = u0; u < u1; ++u) {
        const gphmm_unit &un = b->units[u];
        const uint32_t nr = (uint32_t)(un.read_end - un.read_begin), nh = (uint32_t)(un.hap_end - un.hap_begin);
        UnitDesc d;
        d.read_first = nr ? (uint32_t)(un.read_begin - c.r_lo) : 0;
        d.n_reads = nr;
        d.hap_first = (uint32_t)c.hap_len.size();
        d.n_haps = nh;
        d.out_base = c.n_pairs;
        d.ref_hap = -1;
        d.keep_base = c.n_keep;
        c.n_keep += nr;
        c.streams.insert(c.streams.end(), STREAM_PAD, (uint8_t)CODE_NULL);  // fill/drain codes of the fast kernels
        const uint32_t stream_off = (uint32_t)c.streams.size();
        uint32_t max_h = 1;
        int64_t sum_h = 0;
        {
            const int64_t hap_bytes = nh ? b->hap_off[un.hap_end] - b->hap_off[un.hap_begin] : 0;
            size_t w = c.streams.size();
            c.streams.resize(w + (size_t)hap_bytes + nh);
            uint8_t *dst = c.streams.data();
            for (int64_t h = un.hap_begin; h < un.hap_end; ++h) {
                const int64_t ho = b->hap_off[h];
                const uint32_t H = (uint32_t)(b->hap_off[h + 1] - ho);
                c.hap_len.push_back(H);
                c.hap_stream_off.push_back((uint32_t)w);
                const uint8_t *src = b->hap_bases + ho;
                for (uint32_t j = 0; j < H; ++j) {
                    int16_t code = lut[src[j]];
                    if (code < 0) {
                        if (c.n_codes >= MAX_CODES) throw Error(GPHMM_ERR_ALPHABET, "too many distinct haplotype byte values");
                        code = lut[src[j]] = (int16_t)c.n_codes;
                        c.code_byte[c.n_codes++] = src[j];
                    }
                    dst[w + j] = (uint8_t)code;
                }
                w += H;
                dst[w++] = (uint8_t)CODE_END;
                max_h = std::max(max_h, H);
                sum_h += H;
            }
        }
        const uint32_t stream_len = (uint32_t)c.streams.size() - stream_off;
        c.max_stream_len = std::max(c.max_stream_len, stream_len);
        c.max_hap_len = std::max(c.max_hap_len, max_h);
        d.c0_exp = (force_fp64 ? C0_BASE_EXP_F64 : C0_BASE_EXP_F32) - ceil_log2(max_h);
        c.units.push_back(d);
        // reads of 255+ bases run the striped kernel on the full stream: only shorter reads use the shared streams
        int64_t fast_read_len = 0;
        for (uint32_t r = 0; r < nr && nh; ++r) {
            const uint32_t R = c.read_off[d.read_first + r + 1] - c.read_off[d.read_first + r];
            if ((R + 1) / 32 + 1 <= 8) fast_read_len += R;
        }
        const int n_groups = force_fp64 ? 1 : (int)std::max<int64_t>(1, std::min<int64_t>(want_groups, nh));
        const uint32_t sched_first = (uint32_t)c.unit_sched.size();
        {
            const std::vector<int> order = sorted_hap_order(b, un, share && !force_fp64);
            for (int gi = 0; gi < n_groups; ++gi) {
                const int g0 = (int)((int64_t)nh * gi / n_groups), g1 = (int)((int64_t)nh * (gi + 1) / n_groups);
                c.unit_sched.push_back(plan_unit_sharing(b, un, d.hap_first, share && !force_fp64, c, fast_read_len, order, g0, g1 - g0));
            }
        }
        if (nh == 0) continue;
        for (uint32_t r = 0; r < nr; ++r) {
            const uint32_t rl = d.read_first + r;
            const uint32_t R = c.read_off[rl + 1] - c.read_off[rl];
            Task t;
            t.read = rl; t.stream_off = stream_off; t.stream_len = stream_len;
            t.out_base = d.out_base + r * nh; t.c0_exp = d.c0_exp; t.n_haps = nh; t.hap_first = d.hap_first; t.unit = sched_first;
            // fast kernels need two spare rows below the read (accumulator row + row-0 carrier): R + 2 <= 32 K
            const uint32_t k = (R + 1) / 32 + 1;
            const uint8_t bucket = force_fp64 ? 0 : (k <= 8 ? (uint8_t)(k - 1) : (uint8_t)8);
            // one task per haplotype group for the fast kernels; the striped / fp64 kernels sweep the full stream once
            const int n_t = bucket < 8 && !force_fp64 ? n_groups : 1;
            for (int gi = 0; gi < n_t; ++gi) {
                t.unit = sched_first + (uint32_t)gi;
                raw.push_back(t);
                bucket_of.push_back(bucket);
                ++bucket_count[bucket];
            }
            c.cells += (int64_t)R * sum_h;
        }
        c.n_pairs += nr * nh;
    }
    // counting sort by bucket (stable: unit order is kept inside a bucket)
    c.bucket_begin[0] = 0;
    for (int k = 0; k < N_FP32_BUCKETS; ++k) c.bucket_begin[k + 1] = c.bucket_begin[k] + bucket_count[k];
    c.tasks.resize(raw.size());
    uint32_t cursor[N_FP32_BUCKETS];
    for (int k = 0; k < N_FP32_BUCKETS; ++k) cursor[k] = c.bucket_begin[k];
    for (size_t i = 0; i < raw.size(); ++i) c.tasks[cursor[bucket_of[i]]++] = raw[i];
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Device-side image of a chunk: one metadata blob + the five read arrays + work buffers.
struct DeviceChunk {
    DevBuf reads;   // 5 * span bytes (each array padded to 16 B)
    DevBuf meta;    // read_off | streams | hap_len | hap_stream_off | units | tasks
    DevBuf work;    // sums(float or double) | out(double) | rescue tasks | rescue sums | counters | err
    DevBuf bnd;     // boundary rows of striped kernels
    PinBuf h_meta;  // pinned image of meta
    PinBuf h_reads; // pinned bounce buffer when the caller's arrays are pageable
    PinBuf h_out;   // pinned result buffer (out doubles + [keep flags] + counters + err)
    PinBuf h_modq;  // region steps: pinned image of the modified base/ins/del qualities (only when the caller wants them)
    PinBuf h_raw;   // region steps: pinned image of the un-normalised likelihoods (only when the caller wants them)
    size_t read_stride = 0;
    size_t off_read_off = 0, off_streams = 0, off_hap_len = 0, off_hap_stream_off = 0, off_units = 0, off_tasks = 0, meta_bytes = 0;
    size_t off_sstreams = 0, off_pass = 0, off_segs = 0, off_sched = 0, off_mapq = 0;
    DevBuf snap;    // snapshot slabs of the fast kernels (prefix sharing)
    DevBuf deep;    // DP rows of phmm_exact_f64_kernel (last rescue tier)
    size_t off_sums = 0, off_out = 0, off_rtasks = 0, off_rsums = 0, off_counters = 0, off_err = 0, off_class = 0, off_raw = 0, off_keep = 0, off_deep = 0, work_bytes = 0;
    cudaEvent_t ev_start = nullptr, ev_f32 = nullptr, ev_f64 = nullptr, ev_done = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join[N_AUX] = {nullptr};  // bucket kernels run on side streams
    // small chunks defer the fp64 redo: the rescue kernels are only launched (by finish_chunk) when the downloaded
    // counter says some pair needs them -- three launches less on the latency path of a per-region call
    bool lazy_rescue = false;
    KernelArgs lazy_ka;
    EpilogueArgs lazy_ea;
    PostArgs lazy_pa;
    bool lazy_post = false;
    cudaStream_t lazy_tail = nullptr;
    struct Device *lazy_dev = nullptr;
    bool busy = false;
    void release() {
        reads.release(); meta.release(); work.release(); bnd.release(); snap.release(); deep.release(); h_meta.release(); h_reads.release(); h_out.release(); h_modq.release(); h_raw.release();
        if (ev_start) cudaEventDestroy(ev_start);
        if (ev_f32) cudaEventDestroy(ev_f32);
        if (ev_f64) cudaEventDestroy(ev_f64);
        if (ev_done) cudaEventDestroy(ev_done);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto &e : ev_join) { if (e) cudaEventDestroy(e); e = nullptr; }
        ev_start = ev_f32 = ev_f64 = ev_done = ev_fork = nullptr;
    }
};

struct KernelInfo {
    const void *fn = nullptr;
    int ctas_per_sm = 0;
    size_t smem = 0;
};

// Forward kernels are persistent and several of them (the K buckets of a chunk, the next chunk) are in flight at once;
// the hardware would pack them until no SM has room for the small kernels that CLOSE a chunk (epilogue, rescue), so
// chunks would only complete in groups and the host pipeline would stall.  Every forward CTA therefore asks for
// 1/(occ-2) of the SM's shared memory: any mix of forward kernels then leaves >= 2 CTAs' worth of registers, warps
// and CTA slots free on every SM.
void reserve_headroom(KernelInfo &ki, const void *fn) {
    const int occ = ki.ctas_per_sm;
    static const int headroom = getenv("GPHMM_HEADROOM") ? std::max(1, atoi(getenv("GPHMM_HEADROOM"))) : 2;  // tuning knob
    if (occ < headroom + 2) return;
    const size_t want = (size_t)(227 * 1024) / (size_t)(occ - headroom) - 1024;  // 1 KB per CTA is reserved by the system
    if (want > ki.smem) {
        ki.smem = want & ~(size_t)15;
        if (ki.smem > 48 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem));
    }
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) ki.ctas_per_sm = 1;
}

template <typename T, int K, bool S> KernelInfo kernel_info(int n_codes) {
    KernelInfo ki;
    auto fn = phmm_forward_kernel<T, K, S>;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<T, K>(n_codes);
    if (ki.smem > 48 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) throw Error(GPHMM_ERR_ALPHABET, "prior table does not fit in shared memory");
    reserve_headroom(ki, (const void *)fn);
    return ki;
}

template <int K> KernelInfo fast_kernel_info(int n_codes) {
    KernelInfo ki;
    auto fn = phmm_fast_f32_kernel<K>;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<float, K>(n_codes);
    if (ki.smem > 48 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) throw Error(GPHMM_ERR_ALPHABET, "prior table does not fit in shared memory");
    reserve_headroom(ki, (const void *)fn);
    return ki;
}

KernelInfo fp32_kernel(int bucket, int n_codes) {
    switch (bucket) {
        case 0: return fast_kernel_info<1>(n_codes);
        case 1: return fast_kernel_info<2>(n_codes);
        case 2: return fast_kernel_info<3>(n_codes);
        case 3: return fast_kernel_info<4>(n_codes);
        case 4: return fast_kernel_info<5>(n_codes);
        case 5: return fast_kernel_info<6>(n_codes);
        case 6: return fast_kernel_info<7>(n_codes);
        case 7: return fast_kernel_info<8>(n_codes);
        default: return kernel_info<float, 8, true>(n_codes);  // reads of 255+ bases: striped
    }
}
template <int K, bool SYM> KernelInfo flat_kernel_info(int n_codes) {
    KernelInfo ki;
    auto fn = phmm_flat_f32_kernel<K, SYM>;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<float, K>(n_codes);
    if (ki.smem > 48 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) throw Error(GPHMM_ERR_ALPHABET, "prior table does not fit in shared memory");
    reserve_headroom(ki, (const void *)fn);
    return ki;
}

KernelInfo flat_kernel(int bucket, int n_codes) {
    switch (bucket) {
        case 0: return flat_kernel_info<1, false>(n_codes);
        case 1: return flat_kernel_info<2, false>(n_codes);
        case 2: return flat_kernel_info<3, false>(n_codes);
        case 3: return flat_kernel_info<4, false>(n_codes);
        case 4: return flat_kernel_info<5, false>(n_codes);
        case 5: return flat_kernel_info<6, false>(n_codes);
        case 6: return flat_kernel_info<7, false>(n_codes);
        default: return flat_kernel_info<8, false>(n_codes);
    }
}

KernelInfo sym_kernel(int bucket, int n_codes) {
    switch (bucket) {
        case 0: return flat_kernel_info<1, true>(n_codes);
        case 1: return flat_kernel_info<2, true>(n_codes);
        case 2: return flat_kernel_info<3, true>(n_codes);
        case 3: return flat_kernel_info<4, true>(n_codes);
        case 4: return flat_kernel_info<5, true>(n_codes);
        case 5: return flat_kernel_info<6, true>(n_codes);
        case 6: return flat_kernel_info<7, true>(n_codes);
        default: return flat_kernel_info<8, true>(n_codes);
    }
}

KernelInfo fp64_kernel(int n_codes) { return kernel_info<double, 4, true>(n_codes); }

KernelInfo flat_fp64_kernel(int n_codes) {
    KernelInfo ki;
    auto fn = phmm_flat_f64_kernel;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<double, 8>(n_codes);
    if (ki.smem > 48 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ki.smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) throw Error(GPHMM_ERR_ALPHABET, "prior table does not fit in shared memory");
    return ki;
}

struct Stats {
    std::mutex mu;
    gphmm_stats s{};
};

constexpr int FLAT_KEY = 16;      // Device::info key of the flat-quality kernel of bucket k is FLAT_KEY + k
constexpr int FLAT_F64_KEY = 32;  // ... and of phmm_flat_f64_kernel
constexpr int SYM_KEY = 48;       // symmetric-quality kernel of bucket k is SYM_KEY + k

// One CTA per resident slot (the occupancy already includes the headroom of reserve_headroom()).
inline uint32_t persistent_grid(uint32_t n_tasks, int n_sms, int ctas_per_sm) {
    return std::min<uint32_t>(n_tasks, (uint32_t)(n_sms * ctas_per_sm));
}

struct Device {
    int ordinal = 0;
    int n_sms = 0;
    std::map<int, KernelInfo> kinfo;  // (bucket, n_codes) -> occupancy / smem, queried once
    const KernelInfo &info(int bucket, int n_codes) {
        const int key = bucket * 1024 + n_codes;
        auto it = kinfo.find(key);
        if (it == kinfo.end())
            it = kinfo.emplace(key, bucket < N_FP32_BUCKETS ? fp32_kernel(bucket, n_codes)
                                    : bucket == N_FP32_BUCKETS ? fp64_kernel(n_codes)
                                    : bucket == FLAT_F64_KEY ? flat_fp64_kernel(n_codes)
                                    : bucket >= SYM_KEY ? sym_kernel(bucket - SYM_KEY, n_codes) : flat_kernel(bucket - FLAT_KEY, n_codes)).first;
        return it->second;
    }
    cudaStream_t streams[N_SLOTS] = {nullptr};
    cudaStream_t tails[N_SLOTS] = {nullptr};  // high priority: the small kernels + download that close a chunk
    cudaStream_t aux[N_SLOTS][N_AUX] = {{nullptr}};  // side streams: the kernels of a chunk overlap their tails
    DeviceChunk slots[N_SLOTS];
    DevBuf m2m;
    DevBuf pd_reads, pd_meta, pd_work, pd_bnd;  // PD-HMM path (gphmm_pd_compute): one chunk at a time on streams[0]
    PinBuf pd_h_meta, pd_h_out;
    DevBuf sw_in, sw_bt, sw_aux, sw_out;        // Smith-Waterman path (gphmm_sw_align)
    PinBuf sw_h_in, sw_h_out;
    cudaEvent_t ev_step0 = nullptr, ev_step1 = nullptr;  // bracket a whole run_prepared step on streams[0]
    cudaEvent_t ev_slot_done[N_SLOTS] = {nullptr};
    void init(int ord) {
        ordinal = ord;
        CK(cudaSetDevice(ord));
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, ord));
        if (prop.major != 10) throw Error(GPHMM_ERR_NO_DEVICE, "device is not compute capability 10.x (sm_100a kernels only)");
        n_sms = prop.multiProcessorCount;
        const Tables &t = tables();
        CK(cudaMemcpyToSymbol(c_eps, t.eps.data(), 256 * sizeof(double)));
        m2m.reserve(t.m2m.size() * sizeof(double));
        CK(cudaMemcpy(m2m.p, t.m2m.data(), t.m2m.size() * sizeof(double), cudaMemcpyHostToDevice));
        int prio_lo = 0, prio_hi = 0;  // numerically lower = higher priority
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CK(cudaEventCreate(&ev_step0));
        CK(cudaEventCreate(&ev_step1));
        for (int i = 0; i < N_SLOTS; ++i) CK(cudaEventCreateWithFlags(&ev_slot_done[i], cudaEventDisableTiming));
        for (int i = 0; i < N_SLOTS; ++i) {
            CK(cudaStreamCreateWithPriority(&streams[i], cudaStreamNonBlocking, prio_lo));
            CK(cudaStreamCreateWithPriority(&tails[i], cudaStreamNonBlocking, prio_hi));
            for (int k = 0; k < N_AUX; ++k) CK(cudaStreamCreateWithPriority(&aux[i][k], cudaStreamNonBlocking, prio_lo));
            CK(cudaEventCreate(&slots[i].ev_start));
            CK(cudaEventCreate(&slots[i].ev_f32));
            CK(cudaEventCreate(&slots[i].ev_f64));
            CK(cudaEventCreate(&slots[i].ev_done));
        }
    }
    void release() {
        cudaSetDevice(ordinal);
        for (int i = 0; i < N_SLOTS; ++i) {
            slots[i].release();
            if (streams[i]) cudaStreamDestroy(streams[i]);
            if (tails[i]) cudaStreamDestroy(tails[i]);
            streams[i] = nullptr; tails[i] = nullptr;
            for (int k = 0; k < N_AUX; ++k) { if (aux[i][k]) cudaStreamDestroy(aux[i][k]); aux[i][k] = nullptr; }
        }
        m2m.release();
        if (ev_step0) cudaEventDestroy(ev_step0);
        if (ev_step1) cudaEventDestroy(ev_step1);
        for (auto &e : ev_slot_done) { if (e) cudaEventDestroy(e); e = nullptr; }
        ev_step0 = ev_step1 = nullptr;
    }
};

bool is_pinned(const void *p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

struct RunOptions {
    bool force_fp64 = false;
    bool tristate_off = false;
    const gphmm_region_steps *rs = nullptr;  // gphmm_compute_regions: the steps either side of the kernel
};

// Lay the chunk out on the device and copy its inputs (async on `st`).
void upload_chunk(Device &dev, DeviceChunk &dc, const gphmm_batch *b, const ChunkPlan &c, cudaStream_t st, bool force_fp64,
                  Stats &stats, const gphmm_region_steps *rs = nullptr) {
    const double t0 = now_ms();
    const size_t span = (size_t)(c.base_hi - c.base_lo);
    dc.read_stride = align_up(span, 16);
    dc.reads.reserve(std::max<size_t>(dc.read_stride * 5, 16));
    // metadata blob
    size_t o = 0;
    dc.off_read_off = o; o = align_up(o + c.read_off.size() * 4, 16);
    dc.off_streams = o; o = align_up(o + c.streams.size() + 64, 16);  // + trailing NULL pad
    dc.off_hap_len = o; o = align_up(o + c.hap_len.size() * 4, 16);
    dc.off_hap_stream_off = o; o = align_up(o + c.hap_stream_off.size() * 4, 16);
    dc.off_units = o; o = align_up(o + c.units.size() * sizeof(UnitDesc), 16);
    dc.off_tasks = o; o = align_up(o + c.tasks.size() * sizeof(Task), 16);
    dc.off_sstreams = o; o = align_up(o + c.sstreams.size() + 64, 16);
    dc.off_pass = o; o = align_up(o + c.pass_info.size() * sizeof(PassInfo), 16);
    dc.off_segs = o; o = align_up(o + c.segments.size() * sizeof(Segment), 16);
    dc.off_sched = o; o = align_up(o + c.unit_sched.size() * sizeof(UnitSched), 16);
    dc.off_mapq = o; o = align_up(o + (rs ? c.read_off.size() : 0), 16);
    dc.meta_bytes = std::max<size_t>(o, 16);
    dc.meta.reserve(dc.meta_bytes);
    dc.h_meta.reserve(dc.meta_bytes);
    uint8_t *hm = (uint8_t *)dc.h_meta.p;
    memcpy(hm + dc.off_read_off, c.read_off.data(), c.read_off.size() * 4);
    memcpy(hm + dc.off_streams, c.streams.data(), c.streams.size());
    memset(hm + dc.off_streams + c.streams.size(), CODE_NULL, 64);
    memcpy(hm + dc.off_hap_len, c.hap_len.data(), c.hap_len.size() * 4);
    memcpy(hm + dc.off_hap_stream_off, c.hap_stream_off.data(), c.hap_stream_off.size() * 4);
    memcpy(hm + dc.off_units, c.units.data(), c.units.size() * sizeof(UnitDesc));
    memcpy(hm + dc.off_tasks, c.tasks.data(), c.tasks.size() * sizeof(Task));
    memcpy(hm + dc.off_sstreams, c.sstreams.data(), c.sstreams.size());
    memset(hm + dc.off_sstreams + c.sstreams.size(), CODE_NULL, 64);
    memcpy(hm + dc.off_pass, c.pass_info.data(), c.pass_info.size() * sizeof(PassInfo));
    memcpy(hm + dc.off_segs, c.segments.data(), c.segments.size() * sizeof(Segment));
    memcpy(hm + dc.off_sched, c.unit_sched.data(), c.unit_sched.size() * sizeof(UnitSched));
    if (rs) {
        if (c.r_hi > c.r_lo) memcpy(hm + dc.off_mapq, rs->mapq + c.r_lo, (size_t)(c.r_hi - c.r_lo));
        if (rs->ref_hap) {
            UnitDesc *ud = (UnitDesc *)(hm + dc.off_units);
            for (int64_t u = c.u0; u < c.u1; ++u) ud[u - c.u0].ref_hap = rs->ref_hap[u];
        }
    }
    // work buffers
    const size_t np = std::max<uint32_t>(c.n_pairs, 1);
    o = 0;
    dc.off_sums = o; o = align_up(o + np * (force_fp64 ? 8 : 4), 16);
    dc.off_raw = o; if (rs) o = align_up(o + np * 8, 16);  // region steps: raw read-major likelihoods; `out` is then the normalised matrix
    dc.off_out = o; o = align_up(o + np * 8, 16);
    dc.off_keep = o; o = align_up(o + (rs ? std::max<uint32_t>(c.n_keep, 1) : 0), 16);
    dc.off_counters = o; o = align_up(o + N_COUNTERS * 4, 16);
    dc.off_err = o; o = align_up(o + 16, 16);
    const size_t d2h_bytes = o - dc.off_out;  // out | [keep] | counters | err are contiguous and downloaded together
    dc.off_rtasks = o; o = align_up(o + (force_fp64 ? 0 : np * sizeof(Task)), 16);
    dc.off_rsums = o; o = align_up(o + (force_fp64 ? 0 : np * 8), 16);
    dc.off_class = o; o = align_up(o + c.read_off.size(), 16);
    dc.off_deep = o; o = align_up(o + np * 4, 16);  // deep list: rescue slots whose fp64 sum is not trustworthy either
    dc.work_bytes = o;
    dc.work.reserve(dc.work_bytes);
    dc.h_out.reserve(d2h_bytes);

    const uint8_t *src[5] = {b->read_bases, b->base_q, b->ins_q, b->del_q, b->gcp};
    int64_t h2d = 0;
    if (span > 0) {
        bool pinned = true;
        for (int a = 0; a < 5; ++a) pinned = pinned && is_pinned(src[a] + c.base_lo);
        if (!pinned) dc.h_reads.reserve(dc.read_stride * 5);
        for (int a = 0; a < 5; ++a) {
            const void *from = src[a] + c.base_lo;
            if (!pinned) {
                memcpy((uint8_t *)dc.h_reads.p + a * dc.read_stride, from, span);
                from = (uint8_t *)dc.h_reads.p + a * dc.read_stride;
            }
            CK(cudaMemcpyAsync((uint8_t *)dc.reads.p + a * dc.read_stride, from, span, cudaMemcpyHostToDevice, st));
            h2d += (int64_t)span;
        }
    }
    CK(cudaMemcpyAsync(dc.meta.p, dc.h_meta.p, dc.meta_bytes, cudaMemcpyHostToDevice, st));
    h2d += (int64_t)dc.meta_bytes;
    {
        std::lock_guard<std::mutex> lk(stats.mu);
        stats.s.h2d_bytes += h2d;
        stats.s.host_stage_ms += now_ms() - t0;
    }
    (void)dev;
}

constexpr int64_t LAZY_RESCUE_CELLS = 2000000000;  // chunks below this defer the fp64 redo until the counter is known

// fp64 redo of the chunk's rescue list on `tail`: persistent grids, task count read on the device.
int launch_rescue(Device &dev, DeviceChunk &dc, const ChunkPlan &c, KernelArgs ka, const EpilogueArgs &ea, cudaStream_t tail) {
    int launches = 0;
    uint8_t *work = (uint8_t *)dc.work.p;
    uint32_t *counters = (uint32_t *)(work + dc.off_counters);
    constexpr int FP64_BUCKET = N_FP32_BUCKETS;
    {
        // fp64 redo of the rescue list: persistent grid, task count read on the device
            const KernelInfo &ki = dev.info(FP64_BUCKET, c.n_codes);
            const uint32_t grid = std::min<uint32_t>(c.n_pairs, (uint32_t)(dev.n_sms * ki.ctas_per_sm));
            ka.tasks = (const Task *)(work + dc.off_rtasks);
            ka.n_tasks = 0;
            ka.n_tasks_ptr = counters + 10;
            ka.sums = work + dc.off_rsums;
            // flat-quality reads: constant-coefficient fp64 kernel, one launch per class
            for (int cl = 0; cl < c.n_classes; ++cl) {
                const Tables &tb2 = tables();
                FlatCoefD fd;
                const int qi = c.class_qi[cl], qd = c.class_qd[cl], qc = c.class_qc[cl];
                const double ei = tb2.eps[qi], ed = tb2.eps[qd], ec = tb2.eps[qc], tIM = 1.0 - ec;
                const int mn = std::min(qi, qd), mx = std::max(qi, qd);
                fd.a = tb2.m2m[((mx * (mx + 1)) >> 1) + mn];
                fd.b = tIM * ei; fd.c = tIM * ed; fd.g = ec; fd.d = ec; fd.tmi = ei; fd.tim = tIM;
                fd.class_id = (uint32_t)cl; fd.qi = qi; fd.qd = qd; fd.qc = qc;
                const KernelInfo &kf = dev.info(FLAT_F64_KEY, c.n_codes);
                ka.counter = counters + 12 + cl;
                void *fargs[] = {&ka, &fd};
                CK(cudaLaunchKernel(kf.fn, dim3(std::min<uint32_t>(c.n_pairs, (uint32_t)(dev.n_sms * kf.ctas_per_sm))), dim3(32), fargs, kf.smem, tail));
                ++launches;
            }
            ka.counter = counters + 9;
            ka.bnd = dc.bnd.p;
            ka.bnd_stride = c.max_hap_len + 1;
            void *args[] = {&ka};
            CK(cudaLaunchKernel(ki.fn, dim3(grid), dim3(32), args, ki.smem, tail));
            ++launches;
            phmm_epilogue_rescue<<<std::min<uint32_t>((c.n_pairs + 127) / 128, 1024), 128, 0, tail>>>(
                (const Task *)(work + dc.off_rtasks), counters + 10, ea.rescue_capacity, (const double *)(work + dc.off_rsums), ea.out,
                counters + 11, (uint32_t *)(work + dc.off_deep));
            CK(cudaGetLastError());
            ++launches;
            // last tier: the reference's own arithmetic for the pairs on the deep list (usually none)
            constexpr uint32_t EXACT_CTAS = 8;
            ExactArgs xa;
            xa.tasks = (const Task *)(work + dc.off_rtasks);
            xa.deep_list = (const uint32_t *)(work + dc.off_deep);
            xa.n_deep = counters + 11;
            xa.cursor = counters + 120;
            xa.row_len = c.max_hap_len + 1;
            dc.deep.reserve((size_t)6 * xa.row_len * EXACT_CTAS * 32 * sizeof(double));
            xa.scratch = (double *)dc.deep.p;
            xa.out = ea.out;
            phmm_exact_f64_kernel<<<EXACT_CTAS, 32, 0, tail>>>(ka, xa);
            CK(cudaGetLastError());
            ++launches;
        }
    return launches;
}

// Queue every kernel of the chunk on `st` (no host sync).  Returns the number of launches.
int launch_chunk(Device &dev, DeviceChunk &dc, const ChunkPlan &c, cudaStream_t st, cudaStream_t tail, cudaStream_t *aux, const RunOptions &opt, bool download) {
    int launches = 0;
    uint8_t *meta = (uint8_t *)dc.meta.p, *work = (uint8_t *)dc.work.p;
    uint32_t *counters = (uint32_t *)(work + dc.off_counters);
    CK(cudaMemsetAsync(work + dc.off_counters, 0, (dc.off_err + 16) - dc.off_counters, st));
    CK(cudaEventRecord(dc.ev_start, st));

    KernelArgs ka;
    memset(&ka, 0, sizeof ka);
    ka.rd_bases = (const uint8_t *)dc.reads.p;
    ka.rd_q = ka.rd_bases + dc.read_stride;
    ka.rd_i = ka.rd_q + dc.read_stride;
    ka.rd_d = ka.rd_i + dc.read_stride;
    ka.rd_c = ka.rd_d + dc.read_stride;
    ka.read_off = (const uint32_t *)(meta + dc.off_read_off);
    ka.streams = meta + dc.off_streams;
    ka.hap_len = (const uint32_t *)(meta + dc.off_hap_len);
    ka.sstreams = meta + dc.off_sstreams;
    ka.unit_sched = (const UnitSched *)(meta + dc.off_sched);
    ka.pass_info = (const PassInfo *)(meta + dc.off_pass);
    ka.segments = (const Segment *)(meta + dc.off_segs);
    ka.m2m = (const double *)dev.m2m.p;
    ka.err = (int *)(work + dc.off_err);
    ka.n_codes = c.n_codes;
    ka.tristate_off = opt.tristate_off ? 1 : 0;
    memcpy(ka.code_byte, c.code_byte, sizeof ka.code_byte);

    EpilogueArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.units = (const UnitDesc *)(meta + dc.off_units);
    ea.n_units = (uint32_t)c.units.size();
    ea.hap_len = (const uint32_t *)(meta + dc.off_hap_len);
    ea.hap_stream_off = (const uint32_t *)(meta + dc.off_hap_stream_off);
    ea.sums = work + dc.off_sums;
    ea.out = (double *)(work + (opt.rs ? dc.off_raw : dc.off_out));
    ea.rescue_tasks = (Task *)(work + dc.off_rtasks);
    ea.n_rescue = counters + 10;
    ea.rescue_capacity = std::max<uint32_t>(c.n_pairs, 1);

    const uint32_t n_tasks_total = (uint32_t)c.tasks.size();
    constexpr int FP64_BUCKET = N_FP32_BUCKETS;  // key of the <double, 4, striped> kernel in Device::info

    // Size the boundary buffer for every striped launch of this chunk before anything is queued.
    {
        size_t need = 16;
        const uint32_t n8 = c.bucket_begin[9] - c.bucket_begin[8];
        if (n8) {
            const KernelInfo &ki = dev.info(8, c.n_codes);
            need = std::max(need, (size_t)std::min<uint32_t>(n8, (uint32_t)(dev.n_sms * ki.ctas_per_sm)) * c.max_stream_len * sizeof(Bnd<float>));
        }
        if (n_tasks_total) {
            const KernelInfo &ki = dev.info(FP64_BUCKET, c.n_codes);
            need = std::max(need, (size_t)std::min<uint32_t>(c.n_pairs, (uint32_t)(dev.n_sms * ki.ctas_per_sm)) * (c.max_hap_len + 1) * sizeof(Bnd<double>));
        }
        dc.bnd.reserve(need);
    }

    {
        const uint32_t n_span_reads = (uint32_t)c.read_off.size() - 1;
        // 0. region steps: modifyReadQualities in place on the chunk's read arrays
        if (opt.rs && n_span_reads) {
            ModifyArgs ma;
            memset(&ma, 0, sizeof ma);
            ma.rd_bases = ka.rd_bases;
            ma.rd_q = (uint8_t *)ka.rd_q; ma.rd_i = (uint8_t *)ka.rd_i; ma.rd_d = (uint8_t *)ka.rd_d;
            ma.read_off = ka.read_off;
            ma.n_reads = n_span_reads;
            ma.mapq = meta + dc.off_mapq;
            ma.has_pcr = opt.rs->pcr_rate_factor != 0.0;
            if (ma.has_pcr)
                for (int i = 0; i <= RS_MAX_REPEAT; ++i) {
                    // getErrorModelAdjustedQual (HC/PairHMMLikelihoodCalculationEngine.java:356-358; MathUtils.fastRound)
                    const double d = 40.0 - std::exp(i / (opt.rs->pcr_rate_factor * M_PI)) + 1.0;
                    const int r = d > 0.0 ? (int)(d + 0.5) : (int)(d - 0.5);
                    ma.pcr_cache[i] = (uint8_t)(int8_t)std::max(10, r);
                }
            ma.bq_threshold = opt.rs->base_quality_score_threshold;
            ma.disable_cap_to_mapq = (opt.rs->flags & GPHMM_RS_DISABLE_CAP_TO_MAPQ) != 0;
            phmm_modify_quals_kernel<<<std::min<uint32_t>((n_span_reads + 3) / 4, 148 * 16), 128, 0, st>>>(ma);
            CK(cudaGetLastError());
            ++launches;
            if (opt.rs->hmm_base_q || opt.rs->hmm_ins_q || opt.rs->hmm_del_q) {
                dc.h_modq.reserve(dc.read_stride * 3);
                CK(cudaMemcpyAsync(dc.h_modq.p, (uint8_t *)dc.reads.p + dc.read_stride, dc.read_stride * 3, cudaMemcpyDeviceToHost, st));
            }
        }
        // 1. per-read flat-quality classification (device side; the host only sampled candidate classes)
        {
            ClassifyArgs ca;
            memset(&ca, 0, sizeof ca);
            ca.rd_i = ka.rd_i; ca.rd_d = ka.rd_d; ca.rd_c = ka.rd_c;
            ca.read_off = ka.read_off;
            ca.n_reads = n_span_reads;
            ca.read_class = (uint8_t *)(work + dc.off_class);
            ca.n_classes = (uint32_t)c.n_classes;
            for (int k = 0; k < c.n_classes; ++k) { ca.qi[k] = c.class_qi[k]; ca.qd[k] = c.class_qd[k]; ca.qc[k] = c.class_qc[k]; }
            ca.n_sym = (uint32_t)c.n_sym;
            for (int k = 0; k < c.n_sym; ++k) ca.sym_qc[k] = c.sym_qc[k];
            if (n_span_reads) {
                phmm_classify_kernel<<<std::min<uint32_t>((n_span_reads + 3) / 4, 148 * 16), 128, 0, st>>>(ca);
                CK(cudaGetLastError());
                ++launches;
            }
        }
        ka.read_class = (const uint8_t *)(work + dc.off_class);

        if (opt.force_fp64) {
            // --native-pair-hmm-use-double-precision: no fp32 pass at all.  NaN sums make the epilogue put EVERY pair on
            // the fp64 list, which the two fp64 kernels below then compute.
            CK(cudaMemsetAsync(work + dc.off_sums, 0xff, std::max<size_t>((size_t)c.n_pairs, 1) * sizeof(float), st));
        } else {
        // 2. forward kernels.  Every (bucket) task list is visited by the flat kernel of each class and by the general
        // kernel; each kernel only runs the tasks whose read it owns.  The first launch stays on the chunk's stream,
        // the others fork onto side streams so that short grids overlap the tail of the big one.
        if (!dc.ev_fork) CK(cudaEventCreateWithFlags(&dc.ev_fork, cudaEventDisableTiming));
        CK(cudaEventRecord(dc.ev_fork, st));
        const Tables &tb = tables();
        int order[N_FP32_BUCKETS];
        for (int k = 0; k < N_FP32_BUCKETS; ++k) order[k] = k;
        std::sort(order, order + N_FP32_BUCKETS, [&](int x, int y) {
            return c.bucket_begin[x + 1] - c.bucket_begin[x] > c.bucket_begin[y + 1] - c.bucket_begin[y];
        });
        // every concurrently running fast/flat launch gets its own snapshot slab (one region per CTA)
        constexpr size_t SLAB_PER_CTA = (size_t)(MAX_SNAP_SLOTS + 1) * SNAP_REGS * 32 * sizeof(float);
        {
            size_t need = 0;
            for (int k = 0; k < 8; ++k) {
                const uint32_t n = c.bucket_begin[k + 1] - c.bucket_begin[k];
                if (!n) continue;
                need += (size_t)persistent_grid(n, dev.n_sms, dev.info(k, c.n_codes).ctas_per_sm) * SLAB_PER_CTA;
                for (int cl = 0; cl < c.n_classes; ++cl)
                    need += (size_t)persistent_grid(n, dev.n_sms, dev.info(FLAT_KEY + k, c.n_codes).ctas_per_sm) * SLAB_PER_CTA;
                for (int cl = 0; cl < c.n_sym; ++cl)
                    need += (size_t)persistent_grid(n, dev.n_sms, dev.info(SYM_KEY + k, c.n_codes).ctas_per_sm) * SLAB_PER_CTA;
            }
            dc.snap.reserve(std::max<size_t>(need, 16));
        }
        size_t slab_cursor = 0;
        bool first_launch = true;
        auto launch_on = [&](const KernelInfo &ki, uint32_t n, int aux_idx, void **args) {
            const uint32_t grid = persistent_grid(n, dev.n_sms, ki.ctas_per_sm);
            cudaStream_t ks = st;
            if (!first_launch) {
                ks = aux[aux_idx];
                CK(cudaStreamWaitEvent(ks, dc.ev_fork, 0));
            }
            CK(cudaLaunchKernel(ki.fn, dim3(grid), dim3(32), args, ki.smem, ks));
            ++launches;
            if (!first_launch) {
                if (!dc.ev_join[aux_idx]) CK(cudaEventCreateWithFlags(&dc.ev_join[aux_idx], cudaEventDisableTiming));
                CK(cudaEventRecord(dc.ev_join[aux_idx], ks));
                CK(cudaStreamWaitEvent(st, dc.ev_join[aux_idx], 0));
            }
            first_launch = false;
        };
        for (int oi = 0; oi < N_FP32_BUCKETS; ++oi) {
            const int k = order[oi];
            const uint32_t n = c.bucket_begin[k + 1] - c.bucket_begin[k];
            if (!n) continue;
            ka.tasks = (const Task *)(meta + dc.off_tasks) + c.bucket_begin[k];
            ka.n_tasks = n;
            ka.n_tasks_ptr = nullptr;
            ka.sums = work + dc.off_sums;
            ka.bnd = k == 8 ? dc.bnd.p : nullptr;
            ka.bnd_stride = k == 8 ? c.max_stream_len : 0;
            if (k < 8) {
                for (int cl = 0; cl < c.n_classes; ++cl) {
                    FlatCoef fc;
                    const int qi = c.class_qi[cl], qd = c.class_qd[cl], qc = c.class_qc[cl];
                    const double ei = tb.eps[qi], ed = tb.eps[qd], ec = tb.eps[qc], tIM = 1.0 - ec;
                    const int mn = std::min(qi, qd), mx = std::max(qi, qd);
                    fc.a = (float)tb.m2m[((mx * (mx + 1)) >> 1) + mn];
                    fc.b = (float)(tIM * ei);
                    fc.c = (float)(tIM * ed);
                    fc.g = (float)ec;
                    fc.d = (float)ec;
                    fc.tmi = (float)ei;
                    fc.tim = (float)tIM;
                    fc.class_id = (uint32_t)cl;
                    fc.qi = qi; fc.qd = qd; fc.qc = qc;
                    ka.counter = counters + 16 + 8 * cl + k;
                    ka.snap = (float *)((uint8_t *)dc.snap.p + slab_cursor);
                    slab_cursor += (size_t)persistent_grid(n, dev.n_sms, dev.info(FLAT_KEY + k, c.n_codes).ctas_per_sm) * SLAB_PER_CTA;
                    void *args[] = {&ka, &fc};
                    launch_on(dev.info(FLAT_KEY + k, c.n_codes), n, N_FP32_BUCKETS + 8 * cl + k, args);
                }
                for (int cl = 0; cl < c.n_sym; ++cl) {
                    // symmetric-quality reads (ins == del per base, flat gcp): b = tIM, g = d = eps(gcp), tmi = kappa
                    FlatCoef fc;
                    const int qc = c.sym_qc[cl];
                    const double ec = tb.eps[qc];
                    fc.a = 0.f; fc.c = 0.f;
                    fc.b = (float)(1.0 - ec);
                    fc.g = (float)ec;
                    fc.d = (float)ec;
                    fc.tmi = 1.f / 1024.f;  // kappa: exact power of two; M^ = M eps/kappa ~ M/10 at Q40 and (true sum)/kappa never overflows
                    fc.tim = 1.f;
                    fc.class_id = (uint32_t)(MAX_FLAT_CLASSES + cl);
                    fc.qi = 0; fc.qd = 0; fc.qc = qc;
                    ka.counter = counters + 16 + 8 * (MAX_FLAT_CLASSES + cl) + k;
                    ka.snap = (float *)((uint8_t *)dc.snap.p + slab_cursor);
                    slab_cursor += (size_t)persistent_grid(n, dev.n_sms, dev.info(SYM_KEY + k, c.n_codes).ctas_per_sm) * SLAB_PER_CTA;
                    void *args[] = {&ka, &fc};
                    launch_on(dev.info(SYM_KEY + k, c.n_codes), n, N_FP32_BUCKETS + 8 * (MAX_FLAT_CLASSES + cl) + k, args);
                }
            }
            ka.counter = counters + k;
            if (k < 8) {
                ka.snap = (float *)((uint8_t *)dc.snap.p + slab_cursor);
                slab_cursor += (size_t)persistent_grid(n, dev.n_sms, dev.info(k, c.n_codes).ctas_per_sm) * SLAB_PER_CTA;
            }
            void *args[] = {&ka};
            launch_on(dev.info(k, c.n_codes), n, k, args);
        }
        }
        CK(cudaEventRecord(dc.ev_f32, st));
        CK(cudaStreamWaitEvent(tail, dc.ev_f32, 0));  // the closing kernels run at high priority in the reserved headroom
        if (ea.n_units) {
            const uint32_t per_unit = (c.n_pairs / ea.n_units + 127) / 128;
            const dim3 egrid(std::min<uint32_t>(ea.n_units, 4096), ea.n_units < 296 ? std::max<uint32_t>(1, std::min<uint32_t>(per_unit, 16)) : 1);
            phmm_epilogue_f32<<<egrid, 128, 0, tail>>>(ea);
            CK(cudaGetLastError());
            ++launches;
        }
        // fp64 redo of the rescue list.  Large chunks queue it unconditionally (the list length is read on the device, no
        // host round trip); small chunks look at the downloaded counter first (finish_chunk).
        dc.lazy_rescue = download && !opt.force_fp64 && c.cells < LAZY_RESCUE_CELLS && n_tasks_total > 0;
        dc.lazy_ka = ka; dc.lazy_ea = ea; dc.lazy_tail = tail; dc.lazy_dev = &dev;
        if (n_tasks_total && !dc.lazy_rescue) launches += launch_rescue(dev, dc, c, ka, ea, tail);
        CK(cudaEventRecord(dc.ev_f64, tail));
        // region steps: normalizeLikelihoods + filterPoorlyModeledEvidence, read-major -> allele-major
        if (opt.rs && ea.n_units) {
            PostArgs pa;
            memset(&pa, 0, sizeof pa);
            pa.units = ea.units;
            pa.n_units = ea.n_units;
            pa.lk = ea.out;
            pa.out = (double *)(work + dc.off_out);
            pa.keep = work + dc.off_keep;
            pa.rd_q = ka.rd_q;
            pa.read_off = ka.read_off;
            pa.max_diff_cap = opt.rs->log10_global_read_mismapping_rate;
            pa.max_error_per_base = opt.rs->expected_error_rate_per_base;
            pa.dynamic_scale = opt.rs->read_disqualification_scale;
            pa.symmetric = (opt.rs->flags & GPHMM_RS_SYMMETRIC_NORMALIZE) != 0;
            pa.filter = (opt.rs->flags & GPHMM_RS_FILTER_POORLY) != 0;
            pa.dynamic = (opt.rs->flags & GPHMM_RS_DYNAMIC_DISQ) != 0;
            phmm_normalize_filter_kernel<<<std::min<uint32_t>(ea.n_units, 148 * 8), 128, 0, tail>>>(pa);
            CK(cudaGetLastError());
            ++launches;
            dc.lazy_pa = pa;
        }
        dc.lazy_post = opt.rs && ea.n_units;
    }
    if (download && opt.rs && opt.rs->raw_lk && c.n_pairs) {
        dc.h_raw.reserve((size_t)c.n_pairs * 8);
        CK(cudaMemcpyAsync(dc.h_raw.p, work + dc.off_raw, (size_t)c.n_pairs * 8, cudaMemcpyDeviceToHost, tail));
    }
    if (download) {
        const size_t bytes = (dc.off_err + 16) - dc.off_out;
        CK(cudaMemcpyAsync(dc.h_out.p, work + dc.off_out, bytes, cudaMemcpyDeviceToHost, tail));
    }
    CK(cudaEventRecord(dc.ev_done, tail));
    return launches;
}

// After the chunk's stream work completed: check the error flag, scatter results, account stats.
void finish_chunk(DeviceChunk &dc, const gphmm_batch *b, const ChunkPlan &c, double *out, Stats &stats, bool downloaded,
                  bool count_device_ms = true, const gphmm_region_steps *rs = nullptr) {
    CK(cudaEventSynchronize(dc.ev_done));
    float ms32 = 0, ms64 = 0, msall = 0;
    CK(cudaEventElapsedTime(&ms32, dc.ev_start, dc.ev_f32));
    CK(cudaEventElapsedTime(&ms64, dc.ev_f32, dc.ev_f64));
    CK(cudaEventElapsedTime(&msall, dc.ev_start, dc.ev_done));
    int64_t rescued = 0;
    if (downloaded) {
        const uint8_t *ho = (const uint8_t *)dc.h_out.p;
        const uint32_t *counters = (const uint32_t *)(ho + (dc.off_counters - dc.off_out));
        const int err = *(const int *)(ho + (dc.off_err - dc.off_out));
        if (err) throw Error(GPHMM_ERR_BAD_QUAL, "quality score out of range: ins/del/gcp > 127 or base qual 255");
        rescued = counters[10];
        if (dc.lazy_rescue && rescued > 0) {
            // deferred fp64 redo of a small chunk: now that the list is known to be non-empty, run it and download again
            int n = launch_rescue(*dc.lazy_dev, dc, c, dc.lazy_ka, dc.lazy_ea, dc.lazy_tail);
            if (dc.lazy_post) {
                phmm_normalize_filter_kernel<<<std::min<uint32_t>(dc.lazy_pa.n_units, 148 * 8), 128, 0, dc.lazy_tail>>>(dc.lazy_pa);
                CK(cudaGetLastError());
                ++n;
            }
            if (rs && rs->raw_lk && c.n_pairs)
                CK(cudaMemcpyAsync(dc.h_raw.p, (uint8_t *)dc.work.p + dc.off_raw, (size_t)c.n_pairs * 8, cudaMemcpyDeviceToHost, dc.lazy_tail));
            CK(cudaMemcpyAsync(dc.h_out.p, (uint8_t *)dc.work.p + dc.off_out, (dc.off_err + 16) - dc.off_out, cudaMemcpyDeviceToHost, dc.lazy_tail));
            CK(cudaStreamSynchronize(dc.lazy_tail));
            std::lock_guard<std::mutex> lk(stats.mu);
            stats.s.kernel_launches += n;
            stats.s.d2h_bytes += (int64_t)((dc.off_err + 16) - dc.off_out);
        }
        if (out) {
            const double *res = (const double *)ho;
            for (int64_t u = c.u0; u < c.u1; ++u) {
                const UnitDesc &d = c.units[u - c.u0];
                const size_t n = (size_t)d.n_reads * d.n_haps;
                if (n) memcpy(out + b->units[u].out_off, res + d.out_base, n * sizeof(double));
            }
        }
        if (rs) {
            if (rs->raw_lk) {
                const double *raw = (const double *)dc.h_raw.p;
                for (int64_t u = c.u0; u < c.u1; ++u) {
                    const UnitDesc &d = c.units[u - c.u0];
                    const size_t n = (size_t)d.n_reads * d.n_haps;
                    if (n) memcpy(rs->raw_lk + b->units[u].out_off, raw + d.out_base, n * sizeof(double));
                }
            }
            if (rs->keep) {
                const uint8_t *dk = ho + (dc.off_keep - dc.off_out);
                for (int64_t u = c.u0; u < c.u1; ++u) {
                    const UnitDesc &d = c.units[u - c.u0];
                    if (d.n_reads) memcpy(rs->keep + b->units[u].read_begin, dk + d.keep_base, d.n_reads);
                }
            }
            uint8_t *dst[3] = {rs->hmm_base_q, rs->hmm_ins_q, rs->hmm_del_q};
            const size_t span = (size_t)(c.base_hi - c.base_lo);
            for (int a = 0; a < 3; ++a)
                if (dst[a] && span) memcpy(dst[a] + c.base_lo, (const uint8_t *)dc.h_modq.p + a * dc.read_stride, span);
        }
    }
    std::lock_guard<std::mutex> lk(stats.mu);
    stats.s.pairs += c.n_pairs;
    stats.s.cells += c.cells;
    stats.s.skipped_cells += c.skipped_cells;
    stats.s.rescued_pairs += rescued;
    stats.s.fp32_kernel_ms += ms32;
    stats.s.fp64_kernel_ms += ms64;
    if (count_device_ms) stats.s.device_ms += msall;
    if (downloaded) stats.s.d2h_bytes += (int64_t)((dc.off_err + 16) - dc.off_out);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
struct gphmm_prepared {
    struct Part {
        int device_index;
        ChunkPlan plan;
        DeviceChunk dc;
    };
    std::vector<std::unique_ptr<Part>> parts;
    // out_off per unit is the only thing run_prepared needs from the original batch
    std::vector<gphmm_unit> units;
};

struct gphmm {
    gphmm_config cfg{};
    std::vector<int> ordinals;
    std::vector<std::unique_ptr<Device>> devices;
    std::string last_error = "";
    Stats stats;
    std::mutex run_mu;  // one batch at a time per handle (compute vs. the async worker)
    // async queue.  gphmm_submit appends the caller's arrays to the open *arena* (pinned, reused, SoA like gphmm_batch);
    // the worker takes a whole arena as ONE batch: one host copy per byte, no merge step, DMA straight from the arena.
    struct PinVec {  // growable pinned byte array that keeps its contents
        uint8_t *p = nullptr;
        size_t size = 0, cap = 0;
        void append(const uint8_t *src, size_t n) {
            if (size + n > cap) {
                const size_t want = std::max<size_t>((size + n) * 2, (size_t)4 << 20);
                void *q = nullptr;
                CK(cudaHostAlloc(&q, want, cudaHostAllocPortable));
                if (p) { memcpy(q, p, size); cudaFreeHost(p); }
                p = (uint8_t *)q; cap = want;
            }
            if (n) memcpy(p + size, src, n);
            size += n;
        }
        ~PinVec() { if (p) cudaFreeHost(p); }
    };
    struct JobRef {
        uint64_t ticket;
        double *out;
        int64_t out_base, out_len, read0, n_reads, base0, n_bases, unit0, n_units;
        gphmm_region_steps rs;  // region steps: the caller's output arrays (keep, hmm_*)
    };
    struct Arena {
        PinVec rb, bq, iq, dq, gq, hb;
        std::vector<int64_t> ro, ho;
        std::vector<gphmm_unit> units;
        std::vector<uint8_t> mapq;
        std::vector<int32_t> ref_hap;
        bool has_rs = false;
        gphmm_region_steps rs{};  // parameters shared by every job of the arena
        int64_t out_len = 0;
        std::vector<JobRef> jobs;
        void clear() {
            rb.size = bq.size = iq.size = dq.size = gq.size = hb.size = 0;
            ro.assign(1, 0); ho.assign(1, 0);
            units.clear(); mapq.clear(); ref_hap.clear(); jobs.clear();
            has_rs = false; out_len = 0;
        }
    };
    struct Done { uint64_t ticket; int rc; std::string err; };
    std::mutex q_mu;
    std::condition_variable q_cv, done_cv;
    std::deque<std::unique_ptr<Arena>> pending;      // FIFO; only the back one accepts more jobs
    std::vector<std::unique_ptr<Arena>> free_arenas;
    std::vector<Done> finished;                      // completed, not yet waited for
    uint64_t completed_upto = 0;                     // tickets complete in order
    uint64_t next_ticket = 1;
    std::thread worker;
    bool stop = false;

    int64_t chunk_cells() const { return cfg.chunk_cells > 0 ? cfg.chunk_cells : (int64_t)30000000000LL; }
    int64_t chunk_bytes() const { return cfg.chunk_bytes > 0 ? cfg.chunk_bytes : (int64_t)256 << 20; }
};

namespace {

// Process chunks [ci, ...) of a batch on one device with two stream slots; chunks are claimed from a
// shared cursor so that several devices drain the same batch (the host-side work queue of SURVEY 8e).
// Host-side planning pool: chunk plans are built ahead of the GPU by a few threads (PairHMMNativeArguments.
// maxNumberOfThreads -> gphmm_config.host_threads) and handed to the device loops in chunk order.
struct PlanPool {
    const gphmm_batch *b;
    const std::vector<std::pair<int64_t, int64_t>> &chunks;
    bool f64, share, pcr_hint;
    Stats &stats;
    std::vector<std::unique_ptr<ChunkPlan>> ready;
    std::vector<char> done;
    std::vector<int> err_code;
    std::vector<std::string> err_text;
    std::mutex mu;
    std::condition_variable cv;
    size_t next_plan = 0, consumed = 0, lookahead;
    bool cancel = false;
    std::vector<std::thread> threads;

    PlanPool(const gphmm_batch *b_, const std::vector<std::pair<int64_t, int64_t>> &c, bool f64_, bool share_, int n_threads, Stats &st,
             bool pcr_hint_ = false)
        : b(b_), chunks(c), f64(f64_), share(share_), pcr_hint(pcr_hint_), stats(st), ready(c.size()), done(c.size(), 0), err_code(c.size(), 0),
          err_text(c.size()) {
        n_threads = c.size() <= 1 ? 0 : std::max(1, std::min<int>(n_threads, (int)c.size()));  // one chunk: plan inline, no threads
        lookahead = (size_t)n_threads + N_SLOTS;
        for (int t = 0; t < n_threads; ++t) threads.emplace_back([this] { work(); });
    }
    ~PlanPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            cancel = true;
        }
        cv.notify_all();
        for (auto &t : threads) t.join();
    }
    void work() {
        for (;;) {
            size_t ci;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return cancel || next_plan >= chunks.size() || next_plan < consumed + lookahead; });
                if (cancel || next_plan >= chunks.size()) return;
                ci = next_plan++;
            }
            std::unique_ptr<ChunkPlan> p(new ChunkPlan());
            int code = 0;
            std::string text;
            const double t0 = now_ms();
            try {
                plan_chunk(b, chunks[ci].first, chunks[ci].second, f64, share, *p, pcr_hint);
            } catch (const Error &e) {
                code = e.code; text = e.what();
            } catch (const std::exception &e) {
                code = GPHMM_ERR_NOMEM; text = e.what();
            }
            {
                std::lock_guard<std::mutex> lk(stats.mu);
                stats.s.host_stage_ms += now_ms() - t0;
            }
            {
                std::lock_guard<std::mutex> lk(mu);
                ready[ci] = std::move(p);
                err_code[ci] = code; err_text[ci] = text;
                done[ci] = 1;
            }
            cv.notify_all();
        }
    }
    std::unique_ptr<ChunkPlan> take(size_t ci) {
        if (threads.empty()) {  // synchronous planning (single-chunk batches: the per-region JNI call)
            std::unique_ptr<ChunkPlan> p(new ChunkPlan());
            const double t0 = now_ms();
            plan_chunk(b, chunks[ci].first, chunks[ci].second, f64, share, *p, pcr_hint);
            std::lock_guard<std::mutex> lk(stats.mu);
            stats.s.host_stage_ms += now_ms() - t0;
            return p;
        }
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return done[ci] != 0; });
        consumed = std::max(consumed, ci + 1);
        cv.notify_all();
        if (err_code[ci]) throw Error(err_code[ci], err_text[ci]);
        return std::move(ready[ci]);
    }
};

// Process chunks of a batch on one device with N_SLOTS stream slots; chunks are claimed from a shared cursor so
// that several devices drain the same batch (the host-side work queue of SURVEY 8e).
void device_loop(gphmm *h, Device &dev, const gphmm_batch *b, const std::vector<std::pair<int64_t, int64_t>> &chunks,
                 PlanPool &pool, std::atomic<size_t> &cursor, double *out, std::string &err, int &rc,
                 const gphmm_region_steps *rs = nullptr) {
    try {
        CK(cudaSetDevice(dev.ordinal));
        RunOptions opt;
        opt.force_fp64 = h->cfg.force_fp64 != 0;
        opt.tristate_off = h->cfg.tristate_off != 0;
        opt.rs = rs;
        std::unique_ptr<ChunkPlan> plans[N_SLOTS];
        bool inflight[N_SLOTS] = {false};
        const bool trace = getenv("GPHMM_TRACE") != nullptr;  // per-chunk host timeline on stderr
        const double t_start = now_ms();
        int slot = 0;
        int launches = 0;
        for (;;) {
            const size_t ci = cursor.fetch_add(1);
            if (ci >= chunks.size()) break;
            const double t_f = now_ms();
            if (inflight[slot]) {
                finish_chunk(dev.slots[slot], b, *plans[slot], out, h->stats, true, true, rs);
                inflight[slot] = false;
            }
            const double t_a = now_ms();
            plans[slot] = pool.take(ci);
            const double t_b = now_ms();
            upload_chunk(dev, dev.slots[slot], b, *plans[slot], dev.streams[slot], false, h->stats, rs);
            const double t_c = now_ms();
            launches += launch_chunk(dev, dev.slots[slot], *plans[slot], dev.streams[slot], dev.tails[slot], dev.aux[slot], opt, true);
            if (trace)
                fprintf(stderr, "[gpuphmm] chunk %zu dev %d slot %d: t=%.2f ms  wait-finish %.2f  wait-plan %.2f  upload %.2f  launch %.2f  (cells %.3g)\n",
                        ci, dev.ordinal, slot, t_a - t_start, t_a - t_f, t_b - t_a, t_c - t_b, now_ms() - t_c, (double)plans[slot]->cells);
            inflight[slot] = true;
            slot = (slot + 1) % N_SLOTS;
        }
        for (int s = 0; s < N_SLOTS; ++s) {
            if (inflight[slot]) {
                finish_chunk(dev.slots[slot], b, *plans[slot], out, h->stats, true, true, rs);
                inflight[slot] = false;
            }
            slot = (slot + 1) % N_SLOTS;
        }
        if (trace) fprintf(stderr, "[gpuphmm] dev %d drained at t=%.2f ms\n", dev.ordinal, now_ms() - t_start);
        std::lock_guard<std::mutex> lk(h->stats.mu);
        h->stats.s.kernel_launches += launches;
    } catch (const Error &e) {
        rc = e.code;
        err = e.what();
        // leave the device in a clean state for the next call
        cudaDeviceSynchronize();
        cudaGetLastError();
    }
}

int run_batch(gphmm *h, const gphmm_batch *b, double *out, const gphmm_region_steps *rs = nullptr) {
    std::lock_guard<std::mutex> run_lk(h->run_mu);
    const double t0 = now_ms();
    validate_batch(b);
    if (b->n_units == 0) return GPHMM_OK;
    if (!out) throw Error(GPHMM_ERR_INVALID_ARG, "out is null");
    auto chunks = split_units(b, h->chunk_cells(), h->chunk_bytes(), true);
    std::atomic<size_t> cursor{0};
    const size_t nd = h->devices.size();
    std::vector<std::string> errs(nd);
    std::vector<int> rcs(nd, GPHMM_OK);
    {
        const int n_threads = h->cfg.host_threads > 0 ? h->cfg.host_threads : 4;  // GATK's --native-pair-hmm-threads default
        PlanPool pool(b, chunks, false, h->cfg.no_prefix_sharing == 0, n_threads, h->stats, rs && rs->pcr_rate_factor != 0.0);
        if (nd == 1 || chunks.size() == 1) {
            device_loop(h, *h->devices[0], b, chunks, pool, cursor, out, errs[0], rcs[0], rs);
        } else {
            std::vector<std::thread> th;
            for (size_t d = 0; d < nd; ++d)
                th.emplace_back(device_loop, h, std::ref(*h->devices[d]), b, std::cref(chunks), std::ref(pool), std::ref(cursor), out,
                                std::ref(errs[d]), std::ref(rcs[d]), rs);
            for (auto &t : th) t.join();
        }
    }
    {
        std::lock_guard<std::mutex> lk(h->stats.mu);
        h->stats.s.wall_ms += now_ms() - t0;
    }
    for (size_t d = 0; d < nd; ++d)
        if (rcs[d] != GPHMM_OK) throw Error(rcs[d], errs[d]);
    return GPHMM_OK;
}

// The asynchronous cross-region batching queue: every arena (= everything submitted with equal parameters while the
// worker was busy) is ONE batch, so that many small (region, sample) units fill the GPU together.
void worker_main(gphmm *h) {
    static const bool trace = getenv("GPHMM_TRACE") != nullptr;
    std::vector<double> merged_out, m_raw;
    std::vector<uint8_t> m_keep, m_hq, m_hi, m_hd;
    for (;;) {
        std::unique_ptr<gphmm::Arena> ar;
        {
            std::unique_lock<std::mutex> lk(h->q_mu);
            h->q_cv.wait(lk, [&] { return h->stop || !h->pending.empty(); });
            if (h->pending.empty()) return;
            ar = std::move(h->pending.front());
            h->pending.pop_front();
        }
        const double t_wake = now_ms();
        const size_t nj = ar->jobs.size();
        gphmm_batch b;
        memset(&b, 0, sizeof b);
        b.read_bases = ar->rb.p; b.base_q = ar->bq.p; b.ins_q = ar->iq.p; b.del_q = ar->dq.p; b.gcp = ar->gq.p;
        b.read_off = ar->ro.data(); b.n_reads = (int64_t)ar->ro.size() - 1;
        b.hap_bases = ar->hb.p; b.hap_off = ar->ho.data(); b.n_haps = (int64_t)ar->ho.size() - 1;
        b.units = ar->units.data(); b.n_units = (int64_t)ar->units.size();
        const bool direct = nj == 1;  // a single job writes straight into the caller's arrays
        double *out = direct ? ar->jobs[0].out : nullptr;
        gphmm_region_steps rs = ar->rs;
        if (!direct) {
            merged_out.resize((size_t)ar->out_len + 1);
            out = merged_out.data();
        }
        if (ar->has_rs) {
            ar->mapq.push_back(0); ar->ref_hap.push_back(-1);  // never empty: data() is a valid pointer
            rs.mapq = ar->mapq.data();
            rs.ref_hap = ar->ref_hap.data();
            if (direct) {
                rs.keep = ar->jobs[0].rs.keep; rs.hmm_base_q = ar->jobs[0].rs.hmm_base_q;
                rs.hmm_ins_q = ar->jobs[0].rs.hmm_ins_q; rs.hmm_del_q = ar->jobs[0].rs.hmm_del_q;
                rs.raw_lk = ar->jobs[0].rs.raw_lk;
            } else {
                bool want_keep = false, want_q = false, want_i = false, want_d = false, want_raw = false;
                for (const auto &j : ar->jobs) {
                    want_raw = want_raw || j.rs.raw_lk;
                    want_keep = want_keep || j.rs.keep; want_q = want_q || j.rs.hmm_base_q;
                    want_i = want_i || j.rs.hmm_ins_q; want_d = want_d || j.rs.hmm_del_q;
                }
                if (want_keep) m_keep.resize((size_t)b.n_reads + 1);
                if (want_q) m_hq.resize(ar->rb.size + 1);
                if (want_i) m_hi.resize(ar->rb.size + 1);
                if (want_d) m_hd.resize(ar->rb.size + 1);
                rs.keep = want_keep ? m_keep.data() : nullptr;
                rs.hmm_base_q = want_q ? m_hq.data() : nullptr;
                rs.hmm_ins_q = want_i ? m_hi.data() : nullptr;
                rs.hmm_del_q = want_d ? m_hd.data() : nullptr;
                if (want_raw) m_raw.resize((size_t)ar->out_len + 1);
                rs.raw_lk = want_raw ? m_raw.data() : nullptr;
            }
        }
        auto run = [&](const gphmm_batch &bb, const gphmm_region_steps &rr, std::string &err) -> int {
            try {
                return run_batch(h, &bb, out, ar->has_rs ? &rr : nullptr);
            } catch (const Error &e) {
                err = e.what();
                return e.code;
            } catch (const std::exception &e) {
                err = e.what();
                return GPHMM_ERR_CUDA;
            }
        };
        std::string err;
        const int rc = nj ? run(b, rs, err) : GPHMM_OK;
        std::vector<int> rcs(nj, rc);
        std::vector<std::string> errs(nj, err);
        if (!direct && rc != GPHMM_OK) {
            // something in the merged batch is bad (e.g. a quality out of range): rerun the jobs one by one (each is a
            // sub-range of the arena's units) so that only the offending ticket reports the error
            for (size_t q = 0; q < nj; ++q) {
                const gphmm::JobRef &j = ar->jobs[q];
                gphmm_batch one = b;
                one.units = ar->units.data() + j.unit0;
                one.n_units = j.n_units;
                gphmm_region_steps one_rs = rs;
                one_rs.ref_hap = ar->ref_hap.data() + j.unit0;
                errs[q].clear();
                rcs[q] = run(one, one_rs, errs[q]);
            }
        }
        if (!direct)
            for (size_t q = 0; q < nj; ++q) {
                const gphmm::JobRef &j = ar->jobs[q];
                if (rcs[q] != GPHMM_OK) continue;
                if (j.out_len) memcpy(j.out, merged_out.data() + j.out_base, (size_t)j.out_len * sizeof(double));
                if (!ar->has_rs) continue;
                if (j.rs.raw_lk && j.out_len) memcpy(j.rs.raw_lk, m_raw.data() + j.out_base, (size_t)j.out_len * sizeof(double));
                if (j.rs.keep && j.n_reads) memcpy(j.rs.keep, m_keep.data() + j.read0, (size_t)j.n_reads);
                if (j.rs.hmm_base_q && j.n_bases) memcpy(j.rs.hmm_base_q, m_hq.data() + j.base0, (size_t)j.n_bases);
                if (j.rs.hmm_ins_q && j.n_bases) memcpy(j.rs.hmm_ins_q, m_hi.data() + j.base0, (size_t)j.n_bases);
                if (j.rs.hmm_del_q && j.n_bases) memcpy(j.rs.hmm_del_q, m_hd.data() + j.base0, (size_t)j.n_bases);
            }
        if (trace) fprintf(stderr, "[gpuphmm] queue: batch of %zu jobs took %.3f ms\n", nj, now_ms() - t_wake);
        {
            std::lock_guard<std::mutex> lk(h->q_mu);
            for (size_t q = 0; q < nj; ++q) {
                h->finished.push_back({ar->jobs[q].ticket, rcs[q], errs[q]});
                h->completed_upto = ar->jobs[q].ticket;
            }
            ar->clear();
            h->free_arenas.push_back(std::move(ar));
        }
        h->done_cv.notify_all();
    }
}

// ---- PD-HMM (LoglessPDPairHMM, DRAGEN-GATK mode) ---------------------------------------------------------------------
// Column flags of one partially determined haplotype: alternative-base mask, DEL_END, the state in which row 1
// processes the column, plus how rows >= 2 start (LoglessPDPairHMM.java:59 keeps the state across rows).
void encode_pd_columns(const uint8_t *pd, uint32_t H, uint8_t *flags, uint32_t &first_event, uint32_t &carry) {
    enum { SNP = 1, DEL_START = 2, DEL_END = 4 };  // PartiallyDeterminedHaplotype.java:59-61; A, C, G, T = 8, 16, 32, 64
    uint32_t state = PD_NORMAL;
    first_event = H + 1;
    for (uint32_t j = 1; j <= H; ++j) {
        const uint8_t f = pd[j - 1];
        uint8_t v = (uint8_t)(state << PD_TYPE_SHIFT);
        if (f & SNP) v |= (uint8_t)(PD_SNP_BIT | ((f >> 3) & PD_MASK_BITS));
        if (f & DEL_END) v |= (uint8_t)PD_DEL_END_BIT;
        flags[j - 1] = v;
        if (state == PD_AFTER_DEL) state = PD_NORMAL;
        if (f & DEL_START) state = PD_INSIDE_DEL;
        if (f & DEL_END) state = PD_AFTER_DEL;
        if ((f & (DEL_START | DEL_END)) && first_event == H + 1) first_event = j;
    }
    carry = state;
}

template <typename T, int K> KernelInfo pd_kernel_info() {
    KernelInfo ki;
    auto fn = phmm_pd_kernel<T, K>;
    ki.fn = (const void *)fn;
    ki.smem = 0;
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, 32, 0));
    ki.ctas_per_sm = std::max(1, occ);
    return ki;
}

int run_pd_batch(gphmm *h, const gphmm_batch *b, const uint8_t *hap_pd, double *out) {
    std::lock_guard<std::mutex> run_lk(h->run_mu);
    const double t0 = now_ms();
    validate_batch(b);
    if (b->n_units == 0) return GPHMM_OK;
    if (!out) throw Error(GPHMM_ERR_INVALID_ARG, "out is null");
    if (!hap_pd) throw Error(GPHMM_ERR_INVALID_ARG, "hap_pd_bases is null");
    Device &dev = *h->devices[0];
    CK(cudaSetDevice(dev.ordinal));
    cudaStream_t st = dev.streams[0];
    // reads of 128+ bases: 4 rows per lane in strips of 128 rows keeps 18 warps per SM resident (113 registers) where the
    // 8-row variant (181 registers) keeps 11; GPHMM_PD_K8=1 selects the latter for comparison
    static const bool k8 = getenv("GPHMM_PD_K8") != nullptr;
    static const KernelInfo kf[3] = {pd_kernel_info<float, 2>(), pd_kernel_info<float, 4>(), k8 ? pd_kernel_info<float, 8>() : pd_kernel_info<float, 4>()};
    static const KernelInfo kd = pd_kernel_info<double, 4>();
    const auto chunks = split_units(b, h->chunk_cells() / 4, h->chunk_bytes(), false);
    int64_t launches = 0, total_pairs = 0, total_cells = 0, total_redo = 0, h2d = 0, d2h = 0;
    double device_ms = 0;
    std::vector<uint32_t> read_off;
    std::vector<uint8_t> hap_bytes, hap_flags;
    std::vector<PdTask> tasks[3], all;
    std::vector<uint32_t> unit_out_base;
    for (const auto &ch : chunks) {
        int64_t r_lo = INT64_MAX, r_hi = 0;
        for (int64_t u = ch.first; u < ch.second; ++u) {
            const gphmm_unit &un = b->units[u];
            if (un.read_end > un.read_begin) { r_lo = std::min(r_lo, un.read_begin); r_hi = std::max(r_hi, un.read_end); }
        }
        if (r_hi <= r_lo) r_lo = r_hi = 0;
        const int64_t base_lo = b->n_reads ? b->read_off[r_lo] : 0, base_hi = b->n_reads ? b->read_off[r_hi] : 0;
        const size_t span = (size_t)(base_hi - base_lo), stride = align_up(span, 16);
        read_off.resize((size_t)(r_hi - r_lo) + 1);
        for (int64_t r = 0; r <= r_hi - r_lo; ++r) read_off[r] = (uint32_t)(b->read_off[r_lo + r] - base_lo);
        hap_bytes.clear(); hap_flags.clear(); unit_out_base.clear();
        for (auto &v : tasks) v.clear();
        uint32_t n_pairs = 0, max_h = 1;
        int64_t cells = 0;
        for (int64_t u = ch.first; u < ch.second; ++u) {
            const gphmm_unit &un = b->units[u];
            const uint32_t nr = (uint32_t)(un.read_end - un.read_begin), nh = (uint32_t)(un.hap_end - un.hap_begin);
            unit_out_base.push_back(n_pairs);
            struct HapInfo { uint32_t off, H, first_event, carry; };
            std::vector<HapInfo> hi(nh);
            for (uint32_t k = 0; k < nh; ++k) {
                const int64_t ho = b->hap_off[un.hap_begin + k];
                const uint32_t H = (uint32_t)(b->hap_off[un.hap_begin + k + 1] - ho);
                hi[k].off = (uint32_t)hap_bytes.size(); hi[k].H = H;
                hap_bytes.insert(hap_bytes.end(), b->hap_bases + ho, b->hap_bases + ho + H);
                hap_flags.resize(hap_bytes.size());
                encode_pd_columns(hap_pd + ho, H, hap_flags.data() + hi[k].off, hi[k].first_event, hi[k].carry);
                max_h = std::max(max_h, H);
            }
            for (uint32_t r = 0; r < nr; ++r) {
                const uint32_t rl = (uint32_t)(un.read_begin - r_lo) + r, R = read_off[rl + 1] - read_off[rl];
                const int bucket = R <= 63 ? 0 : (R <= 127 ? 1 : 2);
                for (uint32_t k = 0; k < nh; ++k) {
                    PdTask t;
                    t.read = rl; t.hap_off = hi[k].off; t.H = hi[k].H; t.out_slot = n_pairs + r * nh + k;
                    t.first_event = hi[k].first_event; t.carry = hi[k].carry;
                    t.c0_exp = 125 - ceil_log2(hi[k].H); t.pad = 0;
                    tasks[bucket].push_back(t);
                    cells += (int64_t)R * hi[k].H;
                }
            }
            n_pairs += nr * nh;
        }
        if (n_pairs == 0) continue;
        all.clear();
        uint32_t first[4] = {0, 0, 0, 0};
        for (int k = 0; k < 3; ++k) { first[k + 1] = first[k] + (uint32_t)tasks[k].size(); all.insert(all.end(), tasks[k].begin(), tasks[k].end()); }
        // device image
        size_t o = 0;
        const size_t off_ro = o; o = align_up(o + read_off.size() * 4, 16);
        const size_t off_hb = o; o = align_up(o + hap_bytes.size(), 16);
        const size_t off_hf = o; o = align_up(o + hap_flags.size(), 16);
        const size_t off_tk = o; o = align_up(o + all.size() * sizeof(PdTask), 16);
        const size_t meta_bytes = o;
        dev.pd_meta.reserve(meta_bytes); dev.pd_h_meta.reserve(meta_bytes);
        uint8_t *hm = (uint8_t *)dev.pd_h_meta.p;
        memcpy(hm + off_ro, read_off.data(), read_off.size() * 4);
        memcpy(hm + off_hb, hap_bytes.data(), hap_bytes.size());
        memcpy(hm + off_hf, hap_flags.data(), hap_flags.size());
        memcpy(hm + off_tk, all.data(), all.size() * sizeof(PdTask));
        o = 0;
        const size_t off_out = o; o = align_up(o + (size_t)n_pairs * 8, 16);
        const size_t off_cnt = o; o = align_up(o + 16 * 4, 16);
        const size_t off_err = o; o = align_up(o + 16, 16);
        const size_t dl_bytes = o;
        const size_t off_s32 = o; o = align_up(o + (size_t)n_pairs * 4, 16);
        const size_t off_redo = o; o = align_up(o + (size_t)n_pairs * 4, 16);
        const size_t off_s64 = o; o = align_up(o + (size_t)n_pairs * 8, 16);
        dev.pd_work.reserve(o); dev.pd_h_out.reserve(dl_bytes);
        dev.pd_reads.reserve(std::max<size_t>(stride * 5, 16));
        const uint32_t max_grid = (uint32_t)dev.n_sms * 32;
        dev.pd_bnd.reserve((size_t)max_grid * (max_h + 1) * sizeof(BndPD<double>));
        const uint8_t *src[5] = {b->read_bases, b->base_q, b->ins_q, b->del_q, b->gcp};
        for (int a = 0; a < 5 && span; ++a)
            CK(cudaMemcpyAsync((uint8_t *)dev.pd_reads.p + a * stride, src[a] + base_lo, span, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dev.pd_meta.p, dev.pd_h_meta.p, meta_bytes, cudaMemcpyHostToDevice, st));
        h2d += (int64_t)(span * 5 + meta_bytes);
        uint8_t *meta = (uint8_t *)dev.pd_meta.p, *work = (uint8_t *)dev.pd_work.p;
        uint32_t *counters = (uint32_t *)(work + off_cnt);
        CK(cudaMemsetAsync(work + off_cnt, 0, (off_err + 16) - off_cnt, st));
        CK(cudaEventRecord(dev.ev_step0, st));
        PdArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.rd_bases = (const uint8_t *)dev.pd_reads.p;
        pa.rd_q = pa.rd_bases + stride; pa.rd_i = pa.rd_q + stride; pa.rd_d = pa.rd_i + stride; pa.rd_c = pa.rd_d + stride;
        pa.read_off = (const uint32_t *)(meta + off_ro);
        pa.hap_bases = meta + off_hb; pa.hap_flags = meta + off_hf;
        pa.tasks = (const PdTask *)(meta + off_tk);
        pa.bnd = dev.pd_bnd.p; pa.bnd_stride = max_h + 1;
        pa.m2m = (const double *)dev.m2m.p;
        pa.err = (int *)(work + off_err);
        pa.tristate_off = h->cfg.tristate_off != 0;
        if (!h->cfg.force_fp64) {
            for (int k = 0; k < 3; ++k) {
                const uint32_t n = first[k + 1] - first[k];
                if (!n) continue;
                pa.first = first[k]; pa.n_tasks = n; pa.counter = counters + k;
                pa.sums = work + off_s32 + (size_t)first[k] * 4;
                const uint32_t grid = std::min<uint32_t>(n, std::min<uint32_t>(max_grid, (uint32_t)(dev.n_sms * kf[k].ctas_per_sm)));
                void *args[] = {&pa};
                CK(cudaLaunchKernel(kf[k].fn, dim3(grid), dim3(32), args, 0, st));
                ++launches;
            }
        } else {
            CK(cudaMemsetAsync(work + off_s32, 0xff, (size_t)n_pairs * 4, st));  // NaN: every pair goes to the fp64 list
        }
        phmm_pd_epilogue_f32<<<std::min<uint32_t>((n_pairs + 127) / 128, 2048), 128, 0, st>>>(
            pa.tasks, n_pairs, (const float *)(work + off_s32), (double *)(work + off_out), (uint32_t *)(work + off_redo), counters + 8);
        CK(cudaGetLastError());
        {
            pa.task_index = (const uint32_t *)(work + off_redo);
            pa.n_tasks_ptr = counters + 8; pa.n_tasks = 0; pa.first = 0; pa.counter = counters + 4;
            pa.sums = work + off_s64;
            const uint32_t grid = std::min<uint32_t>(n_pairs, std::min<uint32_t>(max_grid, (uint32_t)(dev.n_sms * kd.ctas_per_sm)));
            void *args[] = {&pa};
            CK(cudaLaunchKernel(kd.fn, dim3(grid), dim3(32), args, 0, st));
        }
        phmm_pd_epilogue_f64<<<std::min<uint32_t>((n_pairs + 127) / 128, 2048), 128, 0, st>>>(
            (const PdTask *)(meta + off_tk), (const uint32_t *)(work + off_redo), counters + 8, (const double *)(work + off_s64), (double *)(work + off_out));
        CK(cudaGetLastError());
        launches += 3;
        CK(cudaEventRecord(dev.ev_step1, st));
        CK(cudaMemcpyAsync(dev.pd_h_out.p, work + off_out, dl_bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        d2h += (int64_t)dl_bytes;
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, dev.ev_step0, dev.ev_step1));
        device_ms += ms;
        const uint8_t *ho = (const uint8_t *)dev.pd_h_out.p;
        const int err = *(const int *)(ho + off_err);
        if (err == 1) throw Error(GPHMM_ERR_BAD_QUAL, "quality score out of range: ins/del/gcp > 127 or base qual 255");
        if (err == 2) throw Error(GPHMM_ERR_INVALID_ARG, "read base other than ACGT on a SNP column of a partially determined haplotype (LoglessPDPairHMM.java:202)");
        total_redo += ((const uint32_t *)(ho + off_cnt))[8];
        const double *res = (const double *)ho;
        for (int64_t u = ch.first; u < ch.second; ++u) {
            const gphmm_unit &un = b->units[u];
            const size_t n = (size_t)(un.read_end - un.read_begin) * (size_t)(un.hap_end - un.hap_begin);
            if (n) memcpy(out + un.out_off, res + unit_out_base[u - ch.first], n * sizeof(double));
        }
        total_pairs += n_pairs; total_cells += cells;
    }
    std::lock_guard<std::mutex> lk(h->stats.mu);
    h->stats.s.pairs += total_pairs; h->stats.s.cells += total_cells; h->stats.s.rescued_pairs += total_redo;
    h->stats.s.kernel_launches += launches; h->stats.s.h2d_bytes += h2d; h->stats.s.d2h_bytes += d2h;
    h->stats.s.device_ms += device_ms; h->stats.s.wall_ms += now_ms() - t0;
    return GPHMM_OK;
}

// ---- Smith-Waterman (SmithWatermanJavaAligner) -------------------------------------------------------------------------
int run_sw_batch(gphmm *h, const gphmm_sw_batch *b, const gphmm_sw_params *prm, int32_t capacity, int32_t *offsets, int32_t *n_elems,
                 uint32_t *elems) {
    std::lock_guard<std::mutex> run_lk(h->run_mu);
    const double t0 = now_ms();
    if (!b || !prm || prm->struct_size != (int32_t)sizeof(gphmm_sw_params)) throw Error(GPHMM_ERR_INVALID_ARG, "sw batch/params null or wrong struct_size");
    if (b->n_pairs < 0 || capacity < 1) throw Error(GPHMM_ERR_INVALID_ARG, "negative pair count or capacity < 1");
    if (b->n_pairs == 0) return GPHMM_OK;
    if (!b->ref_bases || !b->ref_off || !b->alt_bases || !b->alt_off || !offsets || !n_elems || !elems) throw Error(GPHMM_ERR_INVALID_ARG, "null array");
    if (prm->overhang_strategy < 0 || prm->overhang_strategy > 3) throw Error(GPHMM_ERR_INVALID_ARG, "unknown overhang strategy");
    for (int64_t k = 0; k < b->n_pairs; ++k) {
        const int64_t nr = b->ref_off[k + 1] - b->ref_off[k], na = b->alt_off[k + 1] - b->alt_off[k];
        // SmithWatermanJavaAligner.java:64-66: non-null, non-empty sequences are required
        if (nr <= 0 || na <= 0) throw Error(GPHMM_ERR_INVALID_ARG, "Non-null, non-empty sequences are required for the Smith-Waterman calculation");
        if (nr > 32000 || na > 32000) throw Error(GPHMM_ERR_TOO_LARGE, "sequence longer than 32000 bases (backtrack entries are 16 bit)");
    }
    Device &dev = *h->devices[0];
    CK(cudaSetDevice(dev.ordinal));
    cudaStream_t st = dev.streams[0];
    static int ctas_per_sm = 0;
    if (!ctas_per_sm) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, phmm_sw_kernel, 32, 0));
    constexpr uint64_t MAX_BT = (uint64_t)1 << 30;  // int16 entries per chunk (2 GB)
    int64_t launches = 0, cells = 0;
    double device_ms = 0;
    bool overflow = false;
    std::vector<SwTask> tasks;
    for (int64_t k0 = 0; k0 < b->n_pairs;) {
        tasks.clear();
        uint64_t bt = 0, aux = 0;
        int64_t k1 = k0;
        const int64_t rb0 = b->ref_off[k0], ab0 = b->alt_off[k0];
        while (k1 < b->n_pairs && (int64_t)tasks.size() < ((int64_t)1 << 20)) {
            const uint32_t nr = (uint32_t)(b->ref_off[k1 + 1] - b->ref_off[k1]), na = (uint32_t)(b->alt_off[k1 + 1] - b->alt_off[k1]);
            const uint64_t need = (uint64_t)((nr + SW_ROWS - 1) / SW_ROWS) * (na + 31) * SW_ROWS;
            if (!tasks.empty() && (bt + need > MAX_BT || aux + (nr + 1) + 4 * (uint64_t)(na + 1) > ((uint64_t)1 << 30))) break;
            SwTask t;
            t.ref_off = (uint32_t)(b->ref_off[k1] - rb0); t.n_ref = nr;
            t.alt_off = (uint32_t)(b->alt_off[k1] - ab0); t.n_alt = na;
            t.bt_off = bt; t.aux_off = (uint32_t)aux; t.out_off = (uint32_t)tasks.size() * (uint32_t)capacity;
            bt += need; aux += (nr + 1) + 4 * (uint64_t)(na + 1);
            cells += (int64_t)nr * na;
            tasks.push_back(t);
            ++k1;
        }
        const size_t n = tasks.size();
        const size_t ref_bytes = (size_t)(b->ref_off[k1] - rb0), alt_bytes = (size_t)(b->alt_off[k1] - ab0);
        size_t o = 0;
        const size_t off_ref = o; o = align_up(o + ref_bytes, 16);
        const size_t off_alt = o; o = align_up(o + alt_bytes, 16);
        const size_t off_tk = o; o = align_up(o + n * sizeof(SwTask), 16);
        const size_t in_bytes = o;
        dev.sw_in.reserve(in_bytes); dev.sw_h_in.reserve(in_bytes);
        uint8_t *hi = (uint8_t *)dev.sw_h_in.p;
        memcpy(hi + off_ref, b->ref_bases + rb0, ref_bytes);
        memcpy(hi + off_alt, b->alt_bases + ab0, alt_bytes);
        memcpy(hi + off_tk, tasks.data(), n * sizeof(SwTask));
        o = 0;
        const size_t off_el = o; o = align_up(o + n * (size_t)capacity * 4, 16);
        const size_t off_ne = o; o = align_up(o + n * 4, 16);
        const size_t off_of = o; o = align_up(o + n * 4, 16);
        const size_t out_bytes = o;
        const size_t off_cnt = o; o = align_up(o + 16, 16);
        dev.sw_out.reserve(o); dev.sw_h_out.reserve(out_bytes);
        dev.sw_bt.reserve(std::max<size_t>((size_t)bt * 2, 16));
        dev.sw_aux.reserve(std::max<size_t>((size_t)aux * 4, 16));
        CK(cudaMemcpyAsync(dev.sw_in.p, dev.sw_h_in.p, in_bytes, cudaMemcpyHostToDevice, st));
        CK(cudaMemsetAsync((uint8_t *)dev.sw_out.p + off_cnt, 0, 16, st));
        CK(cudaEventRecord(dev.ev_step0, st));
        SwArgs a;
        memset(&a, 0, sizeof a);
        a.ref_bases = (const uint8_t *)dev.sw_in.p + off_ref; a.alt_bases = (const uint8_t *)dev.sw_in.p + off_alt;
        a.tasks = (const SwTask *)((const uint8_t *)dev.sw_in.p + off_tk); a.n_tasks = (uint32_t)n;
        a.counter = (uint32_t *)((uint8_t *)dev.sw_out.p + off_cnt);
        a.bt = (int16_t *)dev.sw_bt.p; a.aux = (int32_t *)dev.sw_aux.p;
        a.elems = (uint32_t *)((uint8_t *)dev.sw_out.p + off_el);
        a.n_elems = (int32_t *)((uint8_t *)dev.sw_out.p + off_ne); a.offsets = (int32_t *)((uint8_t *)dev.sw_out.p + off_of);
        a.capacity = (uint32_t)capacity;
        a.w_match = prm->match_value; a.w_mismatch = prm->mismatch_penalty; a.w_open = prm->gap_open_penalty; a.w_extend = prm->gap_extend_penalty;
        a.strategy = prm->overhang_strategy;
        phmm_sw_kernel<<<(uint32_t)std::min<size_t>(n, (size_t)dev.n_sms * std::max(1, ctas_per_sm)), 32, 0, st>>>(a);
        CK(cudaGetLastError());
        ++launches;
        CK(cudaEventRecord(dev.ev_step1, st));
        CK(cudaMemcpyAsync(dev.sw_h_out.p, dev.sw_out.p, out_bytes, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, dev.ev_step0, dev.ev_step1));
        device_ms += ms;
        const uint8_t *ho = (const uint8_t *)dev.sw_h_out.p;
        memcpy(elems + (size_t)k0 * capacity, ho + off_el, n * (size_t)capacity * 4);
        memcpy(n_elems + k0, ho + off_ne, n * 4);
        memcpy(offsets + k0, ho + off_of, n * 4);
        for (size_t k = 0; k < n; ++k) overflow = overflow || n_elems[k0 + (int64_t)k] < 0;
        {
            std::lock_guard<std::mutex> lk(h->stats.mu);
            h->stats.s.h2d_bytes += (int64_t)in_bytes; h->stats.s.d2h_bytes += (int64_t)out_bytes;
        }
        k0 = k1;
    }
    {
        std::lock_guard<std::mutex> lk(h->stats.mu);
        h->stats.s.pairs += b->n_pairs; h->stats.s.cells += cells; h->stats.s.kernel_launches += launches;
        h->stats.s.device_ms += device_ms; h->stats.s.wall_ms += now_ms() - t0;
    }
    if (overflow) throw Error(GPHMM_ERR_TOO_LARGE, "a CIGAR has more elements than cigar_capacity (n_elems = -1 for those pairs)");
    return GPHMM_OK;
}

void validate_region_steps(const gphmm_batch *batch, const gphmm_region_steps *steps) {
    if (!steps || steps->struct_size != (int32_t)sizeof(gphmm_region_steps)) throw Error(GPHMM_ERR_INVALID_ARG, "steps is null or has the wrong struct_size");
    if (batch->n_reads > 0 && !steps->mapq) throw Error(GPHMM_ERR_INVALID_ARG, "steps.mapq is null");
    // AlleleLikelihoods.java:417-418: the cap must be negative and not NaN
    if (!(steps->log10_global_read_mismapping_rate < 0.0)) throw Error(GPHMM_ERR_INVALID_ARG, "log10_global_read_mismapping_rate must be negative");
    if (!(steps->pcr_rate_factor >= 0.0) || !(steps->expected_error_rate_per_base >= 0.0)) throw Error(GPHMM_ERR_INVALID_ARG, "negative rate in steps");
    if (steps->base_quality_score_threshold < -128 || steps->base_quality_score_threshold > 127) throw Error(GPHMM_ERR_INVALID_ARG, "base_quality_score_threshold is a Java byte");
    if (steps->ref_hap)
        for (int64_t u = 0; u < batch->n_units; ++u) {
            const int64_t nh = batch->units[u].hap_end - batch->units[u].hap_begin;
            if (steps->ref_hap[u] < -1 || steps->ref_hap[u] >= nh) throw Error(GPHMM_ERR_INVALID_ARG, "ref_hap out of range");
        }
}

template <typename F> int guarded(gphmm *h, F &&f) {
    try {
        return f();
    } catch (const Error &e) {
        if (h) h->last_error = e.what();
        return e.code;
    } catch (const std::bad_alloc &) {
        if (h) h->last_error = "host allocation failed";
        return GPHMM_ERR_NOMEM;
    } catch (const std::exception &e) {
        if (h) h->last_error = e.what();
        return GPHMM_ERR_CUDA;
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

int gphmm_abi_version(void) { return GPHMM_ABI_VERSION; }

int gphmm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    int ok = 0;
    for (int i = 0; i < n; ++i) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, i) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

const char *gphmm_strerror(int code) {
    switch (code) {
        case GPHMM_OK: return "ok";
        case GPHMM_ERR_INVALID_ARG: return "invalid argument";
        case GPHMM_ERR_NO_DEVICE: return "no usable CUDA device (need compute capability 10.x)";
        case GPHMM_ERR_CUDA: return "CUDA error";
        case GPHMM_ERR_BAD_QUAL: return "quality score out of range";
        case GPHMM_ERR_ALPHABET: return "haplotype alphabet too large";
        case GPHMM_ERR_NOMEM: return "out of memory";
        case GPHMM_ERR_BAD_TICKET: return "unknown ticket";
        case GPHMM_ERR_TOO_LARGE: return "unit too large for one device chunk";
        default: return "unknown error";
    }
}

int gphmm_create(const gphmm_config *cfg, gphmm_t **out) {
    if (!out) return GPHMM_ERR_INVALID_ARG;
    *out = nullptr;
    gphmm *h = nullptr;
    try {
        h = new gphmm();
    } catch (...) {
        return GPHMM_ERR_NOMEM;
    }
    int rc = guarded(h, [&]() -> int {
        if (cfg) {
            size_t n = std::min<size_t>(sizeof(gphmm_config), cfg->struct_size > 0 ? (size_t)cfg->struct_size : sizeof(gphmm_config));
            memcpy(&h->cfg, cfg, n);
        }
        int n_dev = 0;
        if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
            cudaGetLastError();
            throw Error(GPHMM_ERR_NO_DEVICE, "no CUDA device visible");
        }
        if (h->cfg.n_devices > 0) {
            if (!h->cfg.devices) throw Error(GPHMM_ERR_INVALID_ARG, "devices is null");
            for (int i = 0; i < h->cfg.n_devices; ++i) {
                if (h->cfg.devices[i] < 0 || h->cfg.devices[i] >= n_dev) throw Error(GPHMM_ERR_NO_DEVICE, "device ordinal out of range");
                h->ordinals.push_back(h->cfg.devices[i]);
            }
        } else {
            int cur = 0;
            CK(cudaGetDevice(&cur));
            h->ordinals.push_back(cur);
        }
        h->cfg.devices = nullptr;
        for (int ord : h->ordinals) {
            h->devices.emplace_back(new Device());
            h->devices.back()->init(ord);
        }
        CK(cudaSetDevice(h->ordinals[0]));
        h->worker = std::thread(worker_main, h);
        return GPHMM_OK;
    });
    if (rc != GPHMM_OK) {
        for (auto &d : h->devices) d->release();
        delete h;
        return rc;
    }
    *out = h;
    return GPHMM_OK;
}

void gphmm_destroy(gphmm_t *h) {
    if (!h) return;
    {
        std::lock_guard<std::mutex> lk(h->q_mu);
        h->stop = true;
    }
    h->q_cv.notify_all();
    if (h->worker.joinable()) h->worker.join();
    for (auto &d : h->devices) d->release();
    delete h;
}

const char *gphmm_last_error(const gphmm_t *h) { return h ? h->last_error.c_str() : "null handle"; }

int gphmm_compute(gphmm_t *h, const gphmm_batch *batch, double *out) {
    if (!h) return GPHMM_ERR_INVALID_ARG;
    return guarded(h, [&]() -> int { return run_batch(h, batch, out); });
}

int gphmm_compute_regions(gphmm_t *h, const gphmm_batch *batch, const gphmm_region_steps *steps, double *out) {
    if (!h) return GPHMM_ERR_INVALID_ARG;
    return guarded(h, [&]() -> int {
        if (!batch) throw Error(GPHMM_ERR_INVALID_ARG, "batch is null");
        validate_region_steps(batch, steps);
        return run_batch(h, batch, out, steps);
    });
}

int gphmm_pd_compute(gphmm_t *h, const gphmm_batch *batch, const uint8_t *hap_pd_bases, double *out) {
    if (!h) return GPHMM_ERR_INVALID_ARG;
    return guarded(h, [&]() -> int { return run_pd_batch(h, batch, hap_pd_bases, out); });
}

int gphmm_sw_align(gphmm_t *h, const gphmm_sw_batch *batch, const gphmm_sw_params *params, int32_t cigar_capacity, int32_t *offsets,
                   int32_t *n_elems, uint32_t *elems) {
    if (!h) return GPHMM_ERR_INVALID_ARG;
    return guarded(h, [&]() -> int { return run_sw_batch(h, batch, params, cigar_capacity, offsets, n_elems, elems); });
}

static int submit_job(gphmm_t *h, const gphmm_batch *b, const gphmm_region_steps *steps, double *out, uint64_t *ticket) {
    if (!h || !ticket) return GPHMM_ERR_INVALID_ARG;
    constexpr size_t MAX_JOBS_PER_ARENA = 4096;
    return guarded(h, [&]() -> int {
        validate_batch(b);
        if (steps) validate_region_steps(b, steps);
        if (b->n_units > 0 && !out) throw Error(GPHMM_ERR_INVALID_ARG, "out is null");
        const bool empty = b->n_units == 0;
        const int64_t n_reads = empty ? 0 : b->n_reads, n_haps = empty ? 0 : b->n_haps;
        const int64_t nb = n_reads ? b->read_off[n_reads] : 0, hb = n_haps ? b->hap_off[n_haps] : 0;
        auto same_request = [&](const gphmm::Arena &a) {
            if (a.has_rs != (steps != nullptr)) return false;
            if (!steps) return true;
            return a.rs.flags == steps->flags && a.rs.pcr_rate_factor == steps->pcr_rate_factor &&
                   a.rs.base_quality_score_threshold == steps->base_quality_score_threshold &&
                   a.rs.log10_global_read_mismapping_rate == steps->log10_global_read_mismapping_rate &&
                   a.rs.expected_error_rate_per_base == steps->expected_error_rate_per_base &&
                   a.rs.read_disqualification_scale == steps->read_disqualification_scale;
        };
        {
            std::lock_guard<std::mutex> lk(h->q_mu);
            CK(cudaSetDevice(h->devices[0]->ordinal));  // the arenas are pinned (portable) allocations
            gphmm::Arena *ar = h->pending.empty() ? nullptr : h->pending.back().get();
            if (!ar || ar->jobs.size() >= MAX_JOBS_PER_ARENA || (!ar->jobs.empty() && !same_request(*ar))) {
                std::unique_ptr<gphmm::Arena> fresh;
                if (!h->free_arenas.empty()) { fresh = std::move(h->free_arenas.back()); h->free_arenas.pop_back(); }
                else fresh.reset(new gphmm::Arena());
                fresh->clear();
                h->pending.push_back(std::move(fresh));
                ar = h->pending.back().get();
            }
            if (ar->jobs.empty()) {
                ar->has_rs = steps != nullptr;
                if (steps) ar->rs = *steps;
            }
            gphmm::JobRef j;
            memset(&j.rs, 0, sizeof j.rs);
            if (steps) j.rs = *steps;
            j.out = out;
            j.read0 = (int64_t)ar->ro.size() - 1; j.n_reads = n_reads;
            j.base0 = ar->ro.back(); j.n_bases = nb;
            j.unit0 = (int64_t)ar->units.size(); j.n_units = b->n_units;
            j.out_base = ar->out_len; j.out_len = 0;
            const int64_t h0 = (int64_t)ar->ho.size() - 1, hbase0 = ar->ho.back();
            ar->rb.append(b->read_bases, (size_t)nb); ar->bq.append(b->base_q, (size_t)nb); ar->iq.append(b->ins_q, (size_t)nb);
            ar->dq.append(b->del_q, (size_t)nb); ar->gq.append(b->gcp, (size_t)nb); ar->hb.append(b->hap_bases, (size_t)hb);
            for (int64_t k = 1; k <= n_reads; ++k) ar->ro.push_back(j.base0 + b->read_off[k]);
            for (int64_t k = 1; k <= n_haps; ++k) ar->ho.push_back(hbase0 + b->hap_off[k]);
            for (int64_t k = 0; k < b->n_units; ++k) {
                gphmm_unit u = b->units[k];
                j.out_len = std::max(j.out_len, u.out_off + (u.read_end - u.read_begin) * (u.hap_end - u.hap_begin));
                u.read_begin += j.read0; u.read_end += j.read0; u.hap_begin += h0; u.hap_end += h0; u.out_off += j.out_base;
                ar->units.push_back(u);
            }
            if (steps) {
                if (n_reads) ar->mapq.insert(ar->mapq.end(), steps->mapq, steps->mapq + n_reads);
                for (int64_t k = 0; k < b->n_units; ++k) ar->ref_hap.push_back(steps->ref_hap ? steps->ref_hap[k] : -1);
            }
            ar->out_len += j.out_len;
            j.ticket = h->next_ticket++;
            ar->jobs.push_back(j);
            *ticket = j.ticket;
        }
        h->q_cv.notify_all();
        return GPHMM_OK;
    });
}

int gphmm_submit(gphmm_t *h, const gphmm_batch *b, double *out, uint64_t *ticket) { return submit_job(h, b, nullptr, out, ticket); }

int gphmm_submit_regions(gphmm_t *h, const gphmm_batch *b, const gphmm_region_steps *steps, double *out, uint64_t *ticket) {
    if (!steps) return GPHMM_ERR_INVALID_ARG;
    return submit_job(h, b, steps, out, ticket);
}

int gphmm_wait(gphmm_t *h, uint64_t ticket) {
    if (!h) return GPHMM_ERR_INVALID_ARG;
    std::unique_lock<std::mutex> lk(h->q_mu);
    if (ticket == 0 || ticket >= h->next_ticket) { h->last_error = "unknown ticket"; return GPHMM_ERR_BAD_TICKET; }
    for (;;) {
        for (size_t i = 0; i < h->finished.size(); ++i)
            if (h->finished[i].ticket == ticket) {
                const int rc = h->finished[i].rc;
                if (rc != GPHMM_OK) h->last_error = h->finished[i].err;
                h->finished.erase(h->finished.begin() + i);
                return rc;
            }
        if (ticket <= h->completed_upto) { h->last_error = "ticket already waited for"; return GPHMM_ERR_BAD_TICKET; }
        h->done_cv.wait(lk);
    }
}

int gphmm_prepare(gphmm_t *h, const gphmm_batch *b, gphmm_prepared_t **out) {
    if (!h || !out) return GPHMM_ERR_INVALID_ARG;
    *out = nullptr;
    return guarded(h, [&]() -> int {
        validate_batch(b);
        std::unique_ptr<gphmm_prepared> p(new gphmm_prepared());
        p->units.assign(b->units, b->units + b->n_units);
        auto chunks = split_units(b, h->chunk_cells(), h->chunk_bytes(), false);  // inputs are resident: no staging to hide
        const bool f64 = false;  // forced fp64 only changes which kernels launch_chunk queues
        for (size_t ci = 0; ci < chunks.size(); ++ci) {
            std::unique_ptr<gphmm_prepared::Part> part(new gphmm_prepared::Part());
            part->device_index = (int)(ci % h->devices.size());
            Device &dev = *h->devices[part->device_index];
            CK(cudaSetDevice(dev.ordinal));
            plan_chunk(b, chunks[ci].first, chunks[ci].second, f64, h->cfg.no_prefix_sharing == 0, part->plan);
            CK(cudaEventCreate(&part->dc.ev_start));
            CK(cudaEventCreate(&part->dc.ev_f32));
            CK(cudaEventCreate(&part->dc.ev_f64));
            CK(cudaEventCreate(&part->dc.ev_done));
            upload_chunk(dev, part->dc, b, part->plan, dev.streams[0], f64, h->stats);
            CK(cudaStreamSynchronize(dev.streams[0]));
            part->dc.h_reads.release();
            p->parts.push_back(std::move(part));
        }
        CK(cudaSetDevice(h->ordinals[0]));
        *out = p.release();
        return GPHMM_OK;
    });
}

int gphmm_run_prepared(gphmm_t *h, gphmm_prepared_t *p, double *out) {
    if (!h || !p) return GPHMM_ERR_INVALID_ARG;
    return guarded(h, [&]() -> int {
        const double t0 = now_ms();
        RunOptions opt;
        opt.force_fp64 = h->cfg.force_fp64 != 0;
        opt.tristate_off = h->cfg.tristate_off != 0;
        gphmm_batch b;
        memset(&b, 0, sizeof b);
        b.units = p->units.data();
        b.n_units = (int64_t)p->units.size();
        int launches = 0;
        std::vector<char> used(h->devices.size(), 0);
        std::vector<int> n_on_dev(h->devices.size(), 0);
        for (auto &part : p->parts) {
            Device &dev = *h->devices[part->device_index];
            CK(cudaSetDevice(dev.ordinal));
            if (!used[part->device_index]) {
                CK(cudaEventRecord(dev.ev_step0, dev.streams[0]));
                for (int sl = 1; sl < N_SLOTS; ++sl) CK(cudaStreamWaitEvent(dev.streams[sl], dev.ev_step0, 0));
                used[part->device_index] = 1;
            }
            // consecutive chunks go to different streams so that the tail of one overlaps the head of the next
            const int sl = n_on_dev[part->device_index]++ % N_SLOTS;
            launches += launch_chunk(dev, part->dc, part->plan, dev.streams[sl], dev.tails[sl], dev.aux[sl], opt, out != nullptr);
        }
        for (size_t d = 0; d < h->devices.size(); ++d)
            if (used[d]) {
                Device &dev = *h->devices[d];
                CK(cudaSetDevice(dev.ordinal));
                for (int sl = 1; sl < N_SLOTS; ++sl) {
                    CK(cudaEventRecord(dev.ev_slot_done[sl], dev.streams[sl]));
                    CK(cudaStreamWaitEvent(dev.streams[0], dev.ev_slot_done[sl], 0));
                }
                for (auto &part : p->parts)
                    if ((size_t)part->device_index == d) CK(cudaStreamWaitEvent(dev.streams[0], part->dc.ev_done, 0));
                CK(cudaEventRecord(dev.ev_step1, dev.streams[0]));
            }
        for (auto &part : p->parts) {
            Device &dev = *h->devices[part->device_index];
            CK(cudaSetDevice(dev.ordinal));
            finish_chunk(part->dc, &b, part->plan, out, h->stats, out != nullptr, false);
        }
        // device time of the step = slowest device, first launch to last download
        float step_ms = 0.f;
        for (size_t d = 0; d < h->devices.size(); ++d)
            if (used[d]) {
                CK(cudaSetDevice(h->devices[d]->ordinal));
                CK(cudaEventSynchronize(h->devices[d]->ev_step1));
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, h->devices[d]->ev_step0, h->devices[d]->ev_step1));
                step_ms = std::max(step_ms, ms);
            }
        {
            std::lock_guard<std::mutex> lk(h->stats.mu);
            h->stats.s.device_ms += step_ms;
        }
        CK(cudaSetDevice(h->ordinals[0]));
        std::lock_guard<std::mutex> lk(h->stats.mu);
        h->stats.s.kernel_launches += launches;
        h->stats.s.wall_ms += now_ms() - t0;
        return GPHMM_OK;
    });
}

void gphmm_release_prepared(gphmm_t *h, gphmm_prepared_t *p) {
    if (!p) return;
    for (auto &part : p->parts) {
        if (h) cudaSetDevice(h->devices[part->device_index]->ordinal);
        part->dc.release();
    }
    if (h) cudaSetDevice(h->ordinals[0]);
    delete p;
}

int gphmm_get_stats(const gphmm_t *h, gphmm_stats *out) {
    if (!h || !out) return GPHMM_ERR_INVALID_ARG;
    gphmm *hh = const_cast<gphmm *>(h);
    std::lock_guard<std::mutex> lk(hh->stats.mu);
    *out = hh->stats.s;
    return GPHMM_OK;
}

void gphmm_reset_stats(gphmm_t *h) {
    if (!h) return;
    std::lock_guard<std::mutex> lk(h->stats.mu);
    memset(&h->stats.s, 0, sizeof h->stats.s);
}

int gphmm_plan_stats(const gphmm_batch *b, int prefix_sharing, int64_t out[10]) {
    if (!out) return GPHMM_ERR_INVALID_ARG;
    return guarded(nullptr, [&]() -> int {
        for (int k = 0; k < 10; ++k) out[k] = 0;
        validate_batch(b);
        auto chunks = split_units(b, (int64_t)100000000000LL, (int64_t)256 << 20, false);
        ChunkPlan c;
        for (auto &ch : chunks) {
            plan_chunk(b, ch.first, ch.second, false, prefix_sharing != 0, c);
            out[0] += (int64_t)c.units.size();
            out[1] += (int64_t)c.pass_info.size();
            out[2] += (int64_t)c.segments.size();
            for (const Segment &sg : c.segments) {
                out[3] += sg.n_free;
                out[4] += sg.n_chk;
                out[5] += sg.snap_pos != INT32_MIN;
            }
            for (size_t u = 0; u < c.units.size(); ++u) {
                int64_t cols = 0;
                for (uint32_t k = 0; k < c.units[u].n_haps; ++k) cols += c.hap_len[c.units[u].hap_first + k];
                out[7] += cols;
            }
            out[6] -= c.computed_columns;
            out[8] += 1;
            out[9] += (int64_t)c.tasks.size();
        }
        out[6] += out[7];
        return GPHMM_OK;
    });
}

void *gphmm_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void gphmm_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

}  // extern "C"
