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
#include <functional>
#include <future>
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
// Task buckets: 0..7 = K = 1..8 rows per lane (32 lanes per read), 8 = striped K=8 (reads of 255+ bases), then the buckets
// that put several reads of a unit on one warp: 9..16 = half-warp buckets (two reads, 16 lanes each) for reads of 64..79 /
// ..95 / ..111 / ..127 / ..159 / ..191 / ..223 / ..254 bases, 17..20 = quarter-warp buckets (four reads of EQUAL length, 8 lanes
// each) for reads of up to 79 / 103 / 127 / 159 bases.  MULTI_ROWS = rows per lane (lanes x rows hold R + 1).
constexpr int FIRST_PAIR_BUCKET = 9;
constexpr int N_PAIR_BUCKETS = 8;
constexpr int N_QUAD_BUCKETS = 4;
constexpr int N_MULTI_BUCKETS = N_PAIR_BUCKETS + N_QUAD_BUCKETS;
constexpr int FIRST_QUAD_BUCKET = FIRST_PAIR_BUCKET + N_PAIR_BUCKETS;
constexpr int MULTI_ROWS[N_MULTI_BUCKETS] = {5, 6, 7, 8, 10, 12, 14, 16, /* quarter-warp */ 10, 13, 16, 20};
// half-warp bucket whose general-kernel form (MODE_GEN) takes the non-flat reads of a quarter-warp bucket, two by two
constexpr int QUAD_GEN_PAIR[N_QUAD_BUCKETS] = {0, 2, 3, 4};  // 5, 7, 8, 10 rows per lane: 16 x rows hold the bucket's longest read + 1
constexpr int N_FP32_BUCKETS = FIRST_PAIR_BUCKET + N_MULTI_BUCKETS;
constexpr int N_CLASSES_MAX = MAX_FLAT_CLASSES + MAX_SYM_CLASSES;
constexpr int N_AUX = N_FP32_BUCKETS + (8 + N_MULTI_BUCKETS) * N_CLASSES_MAX;  // side streams: general buckets + flat (class, bucket)
inline bool is_pair_bucket(int k) { return k >= FIRST_PAIR_BUCKET; }   // any bucket with several reads per warp
inline bool is_quad_bucket(int k) { return k >= FIRST_QUAD_BUCKET; }
inline int pair_bucket_rows(int k) { return MULTI_ROWS[k - FIRST_PAIR_BUCKET]; }
// full-warp kernel (index K - 1) for the general reads of a half-warp bucket when MODE_GEN is switched off: 32 K rows hold R + 2
inline int general_bucket_of(int k) { return !is_pair_bucket(k) ? k : std::min(7, ((is_quad_bucket(k) ? 8 : 16) * pair_bucket_rows(k) - 1 + 2 + 31) / 32 - 1); }
// Longest read that takes the half-warp form (A/B knob).  Measured on B200 (profiles/r02_halfwarp_sweep.txt), batches of
// equal-length reads: +29 % at 130 bases, +29..38 % at 150..159, +25 % at 175..190, +17 % at 207..222, +14 % at 235..250.
inline uint32_t half_warp_max_read() {
    static const uint32_t v = getenv("GPHMM_HALFWARP_MAX_READ") ? (uint32_t)atoi(getenv("GPHMM_HALFWARP_MAX_READ")) : 254u;
    return v;
}
inline int pair_bucket_of_read(uint32_t R) {
    if (R < 64 || R > 254 || R > half_warp_max_read()) return -1;
    int p = 0;
    while (16u * (uint32_t)MULTI_ROWS[p] < R + 1) ++p;
    return FIRST_PAIR_BUCKET + p;
}
// Quarter-warp bucket of a read length (four reads of this length on one warp), or -1.  GPHMM_NO_QUAD=1: A/B switch.
inline int quad_bucket_of_read(uint32_t R) {
    static const bool off = getenv("GPHMM_NO_QUAD") != nullptr || getenv("GPHMM_NO_GEN16") != nullptr;
    if (off || R < 64 || R > 159 || R > half_warp_max_read()) return -1;
    int q = 0;
    while (8u * (uint32_t)MULTI_ROWS[N_PAIR_BUCKETS + q] < R + 1) ++q;
    return FIRST_QUAD_BUCKET + q;
}
constexpr int N_COUNTERS = 512;    // [0..8] general fp32 buckets, [9] fp64 queue, [10] n_rescue, [11] n_deep, [120] deep cursor,
                                   // [128 + p] general kernel on half- / quarter-warp bucket p, [160 + 16*c + p] their flat kernels: class c
                                   // [16 + 8*c + k] flat-quality kernels: class c, bucket k

// Growable byte buffer without value-initialisation (std::vector<uint8_t>::resize would memset what is overwritten next).
struct ByteBuf {
    uint8_t *p = nullptr;
    size_t n = 0, cap = 0;
    ByteBuf() = default;
    ByteBuf(const ByteBuf &) = delete;
    ByteBuf &operator=(const ByteBuf &) = delete;
    ~ByteBuf() { free(p); }
    size_t size() const { return n; }
    const uint8_t *data() const { return p; }
    uint8_t *data() { return p; }
    void clear() { n = 0; }
    void reserve(size_t want) {
        if (want <= cap) return;
        want = std::max(want, cap + cap / 2 + 4096);
        void *q = realloc(p, want);
        if (!q) throw std::bad_alloc();
        p = (uint8_t *)q; cap = want;
    }
    uint8_t *append_raw(size_t k) { reserve(n + k); uint8_t *at = p + n; n += k; return at; }  // caller fills [at, at + k)
    void append_fill(size_t k, uint8_t v) { memset(append_raw(k), v, k); }
};

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
    ByteBuf streams;
    std::vector<uint32_t> hap_len, hap_stream_off;
    std::vector<UnitDesc> units;
    std::vector<Task> tasks;               // bucket-sorted
    uint32_t bucket_begin[N_FP32_BUCKETS + 1];
    ByteBuf sstreams;                      // prefix-compressed streams of the fast kernels
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
    // small chunks (a per-region call): the host classifies every read itself, so the classify kernel and the forward
    // launches that would find no read of theirs are skipped (launch latency is what such a call consists of)
    std::vector<uint8_t> host_class;       // per span read: class id as phmm_classify_kernel would write it; empty = classify on the device
    uint32_t class_count[N_FP32_BUCKETS][MAX_FLAT_CLASSES + MAX_SYM_CLASSES + 1];  // tasks per (bucket, class); last = general
};

#include "host_planner.inl"

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Device-side image of a chunk: one metadata blob + the five read arrays + work buffers.
struct DeviceChunk {
    DevBuf reads;   // 5 * span bytes (each array padded to 16 B)
    DevBuf meta;    // read_off | streams | hap_len | hap_stream_off | units | tasks
    DevBuf work;    // sums(float or double) | out(double) | rescue tasks | rescue sums | counters | err
    DevBuf bnd;     // boundary rows of striped kernels
    PinBuf h_meta;  // pinned image of meta
    PinBuf h_out;   // pinned result buffer (out doubles + [keep flags] + counters + err)
    PinBuf h_modq;  // region steps: pinned image of the modified base/ins/del qualities (only when the caller wants them)
    PinBuf h_raw;   // region steps: pinned image of the un-normalised likelihoods (only when the caller wants them)
    size_t read_stride = 0;
    uint8_t *reads_dev = nullptr;  // the five read arrays on the device: `reads`, or inside `meta` for small chunks (one DMA for everything)
    size_t off_reads = 0, off_hclass = 0;
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
        reads.release(); meta.release(); work.release(); bnd.release(); snap.release(); deep.release(); h_meta.release(); h_out.release(); h_modq.release(); h_raw.release();
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

// cudaFuncAttributeMaxDynamicSharedMemorySize is ONE value per (function, device): a later, smaller request must not
// lower it under a cached KernelInfo that still launches with the larger size.  Keep a running maximum.
void raise_dyn_smem(const void *fn, size_t bytes) {
    if (bytes <= 48 * 1024) return;
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> high;
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(mu);
    size_t &cur = high[{dev, fn}];
    if (bytes <= cur) return;
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    cur = bytes;
}

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
        raise_dyn_smem((const void *)fn, ki.smem);
    }
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) ki.ctas_per_sm = 1;
}

template <typename T, int K, bool S> KernelInfo kernel_info(int n_codes) {
    KernelInfo ki;
    auto fn = phmm_forward_kernel<T, K, S>;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<T, K>(n_codes);
    raise_dyn_smem((const void *)fn, ki.smem);
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
    raise_dyn_smem((const void *)fn, ki.smem);
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
    auto fn = phmm_flat_f32_kernel<K, SYM ? MODE_SYM : MODE_FLAT>;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<float, K>(n_codes);
    raise_dyn_smem((const void *)fn, ki.smem);
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

template <int K, bool SYM> KernelInfo flat16_kernel_info(int n_codes) {
    KernelInfo ki;
    auto fn = phmm_flat_f32_kernel<K, SYM ? MODE_SYM : MODE_FLAT, 16>;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<float, K>(n_codes);
    raise_dyn_smem((const void *)fn, ki.smem);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) throw Error(GPHMM_ERR_ALPHABET, "prior table does not fit in shared memory");
    reserve_headroom(ki, (const void *)fn);
    return ki;
}

template <int K, bool SYM> KernelInfo flat8_kernel_info(int n_codes) {  // quarter-warp form: four reads per warp
    KernelInfo ki;
    auto fn = phmm_flat_f32_kernel<K, SYM ? MODE_SYM : MODE_FLAT, 8>;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<float, K>(n_codes);
    raise_dyn_smem((const void *)fn, ki.smem);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) throw Error(GPHMM_ERR_ALPHABET, "prior table does not fit in shared memory");
    reserve_headroom(ki, (const void *)fn);
    return ki;
}

KernelInfo flat16_kernel(int p, bool sym, int n_codes) {  // p: index among the half- (0..7) and quarter-warp (8..11) buckets
    switch (p) {
        case 8: return sym ? flat8_kernel_info<10, true>(n_codes) : flat8_kernel_info<10, false>(n_codes);
        case 9: return sym ? flat8_kernel_info<13, true>(n_codes) : flat8_kernel_info<13, false>(n_codes);
        case 10: return sym ? flat8_kernel_info<16, true>(n_codes) : flat8_kernel_info<16, false>(n_codes);
        case 11: return sym ? flat8_kernel_info<20, true>(n_codes) : flat8_kernel_info<20, false>(n_codes);
        case 0: return sym ? flat16_kernel_info<5, true>(n_codes) : flat16_kernel_info<5, false>(n_codes);
        case 1: return sym ? flat16_kernel_info<6, true>(n_codes) : flat16_kernel_info<6, false>(n_codes);
        case 2: return sym ? flat16_kernel_info<7, true>(n_codes) : flat16_kernel_info<7, false>(n_codes);
        case 3: return sym ? flat16_kernel_info<8, true>(n_codes) : flat16_kernel_info<8, false>(n_codes);
        case 4: return sym ? flat16_kernel_info<10, true>(n_codes) : flat16_kernel_info<10, false>(n_codes);
        case 5: return sym ? flat16_kernel_info<12, true>(n_codes) : flat16_kernel_info<12, false>(n_codes);
        case 6: return sym ? flat16_kernel_info<14, true>(n_codes) : flat16_kernel_info<14, false>(n_codes);
        default: return sym ? flat16_kernel_info<16, true>(n_codes) : flat16_kernel_info<16, false>(n_codes);
    }
}

// half-warp form of the general (per-base quality) kernel: 6-16 rows per lane like the flat half-warp kernels
template <int K> KernelInfo gen16_kernel_info(int n_codes) {
    KernelInfo ki;
    auto fn = phmm_flat_f32_kernel<K, MODE_GEN, 16>;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<float, K>(n_codes);
    raise_dyn_smem((const void *)fn, ki.smem);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) throw Error(GPHMM_ERR_ALPHABET, "prior table does not fit in shared memory");
    reserve_headroom(ki, (const void *)fn);
    return ki;
}
KernelInfo gen16_kernel(int p, int n_codes) {
    if (p >= N_PAIR_BUCKETS) p = QUAD_GEN_PAIR[p - N_PAIR_BUCKETS];  // quarter-warp tasks: the half-warp kernel takes their reads two by two
    switch (p) {
        case 0: return gen16_kernel_info<5>(n_codes);
        case 1: return gen16_kernel_info<6>(n_codes);
        case 2: return gen16_kernel_info<7>(n_codes);
        case 3: return gen16_kernel_info<8>(n_codes);
        case 4: return gen16_kernel_info<10>(n_codes);
        case 5: return gen16_kernel_info<12>(n_codes);
        case 6: return gen16_kernel_info<14>(n_codes);
        default: return gen16_kernel_info<16>(n_codes);
    }
}

KernelInfo fp64_kernel(int n_codes) { return kernel_info<double, 4, true>(n_codes); }

KernelInfo flat_fp64_kernel(int n_codes) {
    KernelInfo ki;
    auto fn = phmm_flat_f64_kernel;
    ki.fn = (const void *)fn;
    ki.smem = prior_table_bytes<double, 8>(n_codes);
    raise_dyn_smem((const void *)fn, ki.smem);
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ki.ctas_per_sm, fn, 32, ki.smem));
    if (ki.ctas_per_sm < 1) throw Error(GPHMM_ERR_ALPHABET, "prior table does not fit in shared memory");
    return ki;
}

struct Stats {
    std::mutex mu;
    gphmm_stats s{};
};

// Device::info keys: 0..8 = general fp32 kernel of bucket k, then
// Device::info keys: 0..8 = general fp32 kernel of bucket k, then
constexpr int FP64_KEY = 9;       // phmm_forward_kernel<double, 4, striped>
constexpr int FLAT_F64_KEY = 32;  // phmm_flat_f64_kernel
constexpr int FLAT_KEY = 100;     // flat-quality kernel of bucket k < 8 is FLAT_KEY + k, of half- / quarter-warp bucket p FLAT_KEY + 8 + p
constexpr int SYM_KEY = 130;      // symmetric-quality kernel: SYM_KEY + k, SYM_KEY + 8 + p
constexpr int GEN16_KEY = 160;    // half-warp form of the general kernel for half- / quarter-warp bucket p is GEN16_KEY + p
inline bool gen16_bucket(int bucket) {
    static const bool off = getenv("GPHMM_NO_GEN16") != nullptr;  // A/B switch: general reads of half-warp buckets on full warps
    return !off && is_pair_bucket(bucket);
}
// key of the kernel that takes the general (per-base quality) reads of a bucket, and the rows per lane its snapshot slab is sized for
inline int general_key(int bucket) { return gen16_bucket(bucket) ? GEN16_KEY + bucket - FIRST_PAIR_BUCKET : general_bucket_of(bucket); }
inline int general_rows(int bucket) {
    if (!gen16_bucket(bucket)) return 8;
    return is_quad_bucket(bucket) ? MULTI_ROWS[QUAD_GEN_PAIR[bucket - FIRST_QUAD_BUCKET]] : pair_bucket_rows(bucket);
}
inline int flat_key(int bucket, bool sym) { return (sym ? SYM_KEY : FLAT_KEY) + (is_pair_bucket(bucket) ? 8 + bucket - FIRST_PAIR_BUCKET : bucket); }
inline size_t slab_per_cta(int bucket) { return snap_slab_bytes(is_pair_bucket(bucket) ? pair_bucket_rows(bucket) : 8); }

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
            it = kinfo.emplace(key, bucket >= GEN16_KEY ? gen16_kernel(bucket - GEN16_KEY, n_codes)
                                    : bucket >= SYM_KEY + 8 ? flat16_kernel(bucket - SYM_KEY - 8, true, n_codes)
                                    : bucket >= SYM_KEY ? sym_kernel(bucket - SYM_KEY, n_codes)
                                    : bucket >= FLAT_KEY + 8 ? flat16_kernel(bucket - FLAT_KEY - 8, false, n_codes)
                                    : bucket >= FLAT_KEY ? flat_kernel(bucket - FLAT_KEY, n_codes)
                                    : bucket == FLAT_F64_KEY ? flat_fp64_kernel(n_codes)
                                    : bucket == FP64_KEY ? fp64_kernel(n_codes)
                                    : fp32_kernel(bucket, n_codes)).first;
        return it->second;
    }
    cudaStream_t streams[N_SLOTS] = {nullptr};
    cudaStream_t tails[N_SLOTS] = {nullptr};  // high priority: the small kernels + download that close a chunk
    cudaStream_t aux[N_SLOTS][N_AUX] = {{nullptr}};  // side streams: the kernels of a chunk overlap their tails
    DeviceChunk slots[N_SLOTS];
    DevBuf m2m;
    // PD-HMM path (gphmm_pd_compute): two chunks in flight, slot k on streams[k]
    DevBuf pd_reads[2], pd_meta[2], pd_work[2], pd_bnd[2];
    PinBuf pd_h_meta[2], pd_h_out[2];
    cudaEvent_t pd_ev0[2] = {nullptr, nullptr}, pd_ev1[2] = {nullptr, nullptr};
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
        for (int i = 0; i < 2; ++i) { CK(cudaEventCreate(&pd_ev0[i])); CK(cudaEventCreate(&pd_ev1[i])); }
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
        for (int i = 0; i < 2; ++i) {
            pd_reads[i].release(); pd_meta[i].release(); pd_work[i].release(); pd_bnd[i].release();
            pd_h_meta[i].release(); pd_h_out[i].release();
            if (pd_ev0[i]) cudaEventDestroy(pd_ev0[i]);
            if (pd_ev1[i]) cudaEventDestroy(pd_ev1[i]);
            pd_ev0[i] = pd_ev1[i] = nullptr;
        }
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
    bool single_chunk = false;               // the whole batch is this chunk (a per-region call): nothing to overlap with
};

// Lay the chunk out on the device and copy its inputs (async on `st`).
void upload_chunk(Device &dev, DeviceChunk &dc, const gphmm_batch *b, const ChunkPlan &c, cudaStream_t st, bool force_fp64,
                  Stats &stats, const gphmm_region_steps *rs = nullptr) {
    const double t0 = now_ms();
    const size_t span = (size_t)(c.base_hi - c.base_lo);
    dc.read_stride = align_up(span, 16);
    const uint8_t *src[5] = {b->read_bases, b->base_q, b->ins_q, b->del_q, b->gcp};
    bool pinned = true;
    if (span > 0)
        for (int a = 0; a < 5; ++a) pinned = pinned && is_pinned(src[a] + c.base_lo);
    // pageable arrays need a pinned bounce copy anyway, and for a small chunk one copy + ONE DMA beats five DMAs:
    // the read arrays then travel inside the metadata blob
    const bool inline_reads = span > 0 && (!pinned || span * 5 <= ((size_t)256 << 10));
    if (!inline_reads) dc.reads.reserve(std::max<size_t>(dc.read_stride * 5, 16));
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
    dc.off_hclass = o; o = align_up(o + c.host_class.size(), 16);
    dc.off_reads = o; o = align_up(o + (inline_reads ? dc.read_stride * 5 : 0), 16);
    dc.meta_bytes = std::max<size_t>(o, 16);
    dc.meta.reserve(dc.meta_bytes);
    dc.h_meta.reserve(dc.meta_bytes);
    uint8_t *hm = (uint8_t *)dc.h_meta.p;
    // a large pageable chunk (what a JVM caller's pack() hands over): four of the five read arrays are copied by helper threads
    // while this thread -- which also launches the device's kernels -- copies the metadata and the fifth
    std::future<void> side[4];
    const bool parallel = inline_reads && span * 5 > ((size_t)4 << 20);
    for (int a = 1; a < 5 && parallel; ++a) {
        uint8_t *dst = hm + dc.off_reads + a * dc.read_stride;
        const uint8_t *from = src[a] + c.base_lo;
        side[a - 1] = std::async(std::launch::async, [dst, from, span]() { memcpy(dst, from, span); });
    }
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
    if (!c.host_class.empty()) memcpy(hm + dc.off_hclass, c.host_class.data(), c.host_class.size());
    if (inline_reads)
        for (int a = 0; a < (parallel ? 1 : 5); ++a) memcpy(hm + dc.off_reads + a * dc.read_stride, src[a] + c.base_lo, span);
    for (auto &f : side) if (f.valid()) f.get();
    dc.reads_dev = inline_reads ? (uint8_t *)dc.meta.p + dc.off_reads : (uint8_t *)dc.reads.p;
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

    int64_t h2d = 0;
    if (span > 0 && !inline_reads) {
        for (int a = 0; a < 5; ++a) {
            CK(cudaMemcpyAsync((uint8_t *)dc.reads.p + a * dc.read_stride, src[a] + c.base_lo, span, cudaMemcpyHostToDevice, st));
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
    constexpr int FP64_BUCKET = FP64_KEY;
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
                fd.b = tIM * ei / fd.a; fd.c = tIM * ed / fd.a; fd.g = ec; fd.d = ec; fd.tmi = ei; fd.tim = tIM / fd.a;  // tMM folded into the priors
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
    // counters and error flag start at 0; the keep flags and the alignment gaps of the downloaded block are cleared with them
    // (compute-sanitizer initcheck then sees a fully initialised D2H source)
    {
        const size_t out_end = dc.off_out + (size_t)std::max<uint32_t>(c.n_pairs, 1) * 8;  // <= off_keep: one memset covers the gap too
        CK(cudaMemsetAsync(work + out_end, 0, (dc.off_err + 16) - out_end, st));
    }
    // a per-region call is a chain of a few short kernels: keep it on ONE stream (a cross-stream event costs more than it
    // hides; with several chunks in flight the closing kernels need the high-priority stream, see reserve_headroom)
    if (opt.single_chunk && c.cells < LAZY_RESCUE_CELLS) tail = st;
    CK(cudaEventRecord(dc.ev_start, st));

    KernelArgs ka;
    memset(&ka, 0, sizeof ka);
    ka.rd_bases = (const uint8_t *)dc.reads_dev;
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
    constexpr int FP64_BUCKET = FP64_KEY;  // key of the <double, 4, striped> kernel in Device::info

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
                const size_t span = (size_t)(c.base_hi - c.base_lo);  // without the alignment tail of each array
                for (int a = 0; a < 3 && span; ++a)
                    CK(cudaMemcpyAsync((uint8_t *)dc.h_modq.p + a * dc.read_stride, dc.reads_dev + (a + 1) * dc.read_stride, span,
                                       cudaMemcpyDeviceToHost, st));
            }
        }
        // 1. per-read flat-quality classification: on the device (the host only sampled candidate classes), except for
        // small chunks, which the planner classified itself
        const bool host_classes = !c.host_class.empty() && !opt.rs;
        if (host_classes) {
            ka.read_class = meta + dc.off_hclass;
        } else {
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
            ka.read_class = (const uint8_t *)(work + dc.off_class);
        }
        // with host-side classes the launches that would find no read of theirs are skipped
        auto has_work = [&](int bucket, int cls) { return !host_classes || bucket == 8 || c.class_count[bucket][cls] > 0; };

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
        {
            size_t need = 0;
            for (int k = 0; k < N_FP32_BUCKETS; ++k) {
                const uint32_t n = c.bucket_begin[k + 1] - c.bucket_begin[k];
                if (!n || k == 8) continue;
                need += (size_t)persistent_grid(n, dev.n_sms, dev.info(general_key(k), c.n_codes).ctas_per_sm) * snap_slab_bytes(general_rows(k));
                for (int cl = 0; cl < c.n_classes; ++cl)
                    need += (size_t)persistent_grid(n, dev.n_sms, dev.info(flat_key(k, false), c.n_codes).ctas_per_sm) * slab_per_cta(k);
                for (int cl = 0; cl < c.n_sym; ++cl)
                    need += (size_t)persistent_grid(n, dev.n_sms, dev.info(flat_key(k, true), c.n_codes).ctas_per_sm) * slab_per_cta(k);
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
            const bool pair = is_pair_bucket(k);
            const int kp = pair ? k - FIRST_PAIR_BUCKET : k;  // index of the bucket among its kind
            ka.tasks = (const Task *)(meta + dc.off_tasks) + c.bucket_begin[k];
            ka.n_tasks = n;
            ka.n_tasks_ptr = nullptr;
            ka.sums = work + dc.off_sums;
            ka.bnd = k == 8 ? dc.bnd.p : nullptr;
            ka.bnd_stride = k == 8 ? c.max_stream_len : 0;
            ka.pair_tasks = is_quad_bucket(k) ? 2 : (pair ? 1 : 0);
            if (k != 8) {
                for (int cl = 0; cl < c.n_classes; ++cl) {
                    if (!has_work(k, cl)) continue;
                    FlatCoef fc;
                    const int qi = c.class_qi[cl], qd = c.class_qd[cl], qc = c.class_qc[cl];
                    const double ei = tb.eps[qi], ed = tb.eps[qd], ec = tb.eps[qc], tIM = 1.0 - ec;
                    const int mn = std::min(qi, qd), mx = std::max(qi, qd);
                    // tMM is folded into the kernel's prior table and divides the other coefficients of the match update
                    // (plan_chunk registers no flat class with tMM = 0)
                    const double a = tb.m2m[((mx * (mx + 1)) >> 1) + mn];
                    fc.a = (float)a;
                    fc.b = (float)(tIM * ei / a);
                    fc.c = (float)(tIM * ed / a);
                    fc.g = (float)ec;
                    fc.d = (float)ec;
                    fc.tmi = (float)ei;
                    fc.tim = (float)(tIM / a);
                    fc.class_id = (uint32_t)cl;
                    fc.qi = qi; fc.qd = qd; fc.qc = qc;
                    ka.counter = pair ? counters + 160 + 16 * cl + kp : counters + 16 + 8 * cl + k;
                    ka.snap = (float *)((uint8_t *)dc.snap.p + slab_cursor);
                    const KernelInfo &ki = dev.info(flat_key(k, false), c.n_codes);
                    slab_cursor += (size_t)persistent_grid(n, dev.n_sms, ki.ctas_per_sm) * slab_per_cta(k);
                    void *args[] = {&ka, &fc};
                    launch_on(ki, n, pair ? N_FP32_BUCKETS + 8 * N_CLASSES_MAX + N_MULTI_BUCKETS * cl + kp : N_FP32_BUCKETS + 8 * cl + k, args);
                }
                for (int cl = 0; cl < c.n_sym; ++cl) {
                    if (!has_work(k, MAX_FLAT_CLASSES + cl)) continue;
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
                    const int ci = MAX_FLAT_CLASSES + cl;
                    ka.counter = pair ? counters + 160 + 16 * ci + kp : counters + 16 + 8 * ci + k;
                    ka.snap = (float *)((uint8_t *)dc.snap.p + slab_cursor);
                    const KernelInfo &ki = dev.info(flat_key(k, true), c.n_codes);
                    slab_cursor += (size_t)persistent_grid(n, dev.n_sms, ki.ctas_per_sm) * slab_per_cta(k);
                    void *args[] = {&ka, &fc};
                    launch_on(ki, n, pair ? N_FP32_BUCKETS + 8 * N_CLASSES_MAX + N_MULTI_BUCKETS * ci + kp : N_FP32_BUCKETS + 8 * ci + k, args);
                }
            }
            if (!has_work(k, MAX_FLAT_CLASSES + MAX_SYM_CLASSES)) continue;
            ka.counter = pair ? counters + 128 + kp : counters + k;
            const KernelInfo &kg = dev.info(general_key(k), c.n_codes);
            if (k != 8) {
                ka.snap = (float *)((uint8_t *)dc.snap.p + slab_cursor);
                slab_cursor += (size_t)persistent_grid(n, dev.n_sms, kg.ctas_per_sm) * snap_slab_bytes(general_rows(k));
            }
            FlatCoef gc;  // half-warp form of the general kernel: the flat-kernel family with every coefficient per row
            memset(&gc, 0, sizeof gc);
            gc.class_id = CLASS_GENERAL;
            void *args[] = {&ka, &gc};  // (the full-warp kernels take the first argument only)
            launch_on(kg, n, k, args);
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
    std::unique_ptr<ChunkPlan> spare_plan;  // the plan of the last single-chunk call: its buffers serve the next one (under run_mu)
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

// Host-side planning pool: chunk plans are built ahead of the GPU by a few threads (PairHMMNativeArguments.
// maxNumberOfThreads -> gphmm_config.host_threads) and handed to the device loops in chunk order.
struct PlanPool {
    const gphmm_batch *b;
    const std::vector<std::pair<int64_t, int64_t>> &chunks;
    bool f64, share;
    int steps_mode;  // 0 = plain likelihoods, 1 = region steps, 2 = region steps with the PCR indel model (plan_chunk)
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
             int steps_mode_ = 0)
        : b(b_), chunks(c), f64(f64_), share(share_), steps_mode(steps_mode_), stats(st), ready(c.size()), done(c.size(), 0), err_code(c.size(), 0),
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
                plan_chunk(b, chunks[ci].first, chunks[ci].second, f64, share, *p, steps_mode);
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
    std::unique_ptr<ChunkPlan> recycled;  // a plan of an earlier call whose buffers can be reused (per-region calls)
    std::unique_ptr<ChunkPlan> take(size_t ci) {
        if (threads.empty()) {  // synchronous planning (single-chunk batches: the per-region JNI call)
            std::unique_ptr<ChunkPlan> p = recycled ? std::move(recycled) : std::unique_ptr<ChunkPlan>(new ChunkPlan());
            const double t0 = now_ms();
            plan_chunk(b, chunks[ci].first, chunks[ci].second, f64, share, *p, steps_mode);
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
        opt.single_chunk = chunks.size() == 1;
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
        if (chunks.size() == 1 && plans[0]) h->spare_plan = std::move(plans[0]);
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
    // several devices drain one list of chunks: full-size chunks through the middle of the batch (their fixed costs weigh
    // least), geometrically smaller ones at the end so that the devices finish together (unless the caller fixed the size)
    const bool fixed_size = h->cfg.chunk_cells > 0;
    auto chunks = split_units(b, h->chunk_cells(), h->chunk_bytes(), true, fixed_size ? 0 : (int)std::max<size_t>(h->devices.size(), 1));
    if (getenv("GPHMM_TRACE")) fprintf(stderr, "[gpuphmm] batch of %lld units: validated and split into %zu chunks in %.2f ms\n", (long long)b->n_units, chunks.size(), now_ms() - t0);
    std::atomic<size_t> cursor{0};
    const size_t nd = h->devices.size();
    std::vector<std::string> errs(nd);
    std::vector<int> rcs(nd, GPHMM_OK);
    {
        // GATK's --native-pair-hmm-threads default is 4; one planner thread keeps about one B200 busy (12 us per configs[1]
        // region on either side), so a handle over several devices uses at least 3 per device
        // (GATK passes its default of 4 explicitly: with several devices the floor applies to explicit values too)
        const int asked = h->cfg.host_threads > 0 ? h->cfg.host_threads : 4;
        const int n_threads = h->devices.size() > 1 ? std::max<int>(asked, 3 * (int)h->devices.size()) : asked;
        PlanPool pool(b, chunks, false, h->cfg.no_prefix_sharing == 0, n_threads, h->stats, rs ? (rs->pcr_rate_factor != 0.0 ? 2 : 1) : 0);
        if (chunks.size() == 1) pool.recycled = std::move(h->spare_plan);
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
    if (getenv("GPHMM_TRACE")) fprintf(stderr, "[gpuphmm] batch done after %.2f ms\n", now_ms() - t0);
    for (size_t d = 0; d < nd; ++d)
        if (rcs[d] != GPHMM_OK) throw Error(rcs[d], errs[d]);
    return GPHMM_OK;
}

#include "host_queue.inl"

#include "host_pdhmm.inl"

#include "host_sw.inl"

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
        validate_batch(batch);  // before anything indexes batch->units
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
            // a failed append (pinned or heap allocation) must leave the arena as it was: later jobs share its arrays
            const size_t keep_rb = ar->rb.size, keep_hb = ar->hb.size, keep_ro = ar->ro.size(), keep_ho = ar->ho.size(),
                         keep_units = ar->units.size(), keep_mapq = ar->mapq.size(), keep_ref = ar->ref_hap.size();
            struct Rollback {
                std::function<void()> undo;
                bool armed = true;
                ~Rollback() { if (armed) undo(); }
            } rollback{[&] {
                ar->rb.size = ar->bq.size = ar->iq.size = ar->dq.size = ar->gq.size = keep_rb;
                ar->hb.size = keep_hb;
                ar->ro.resize(keep_ro); ar->ho.resize(keep_ho); ar->units.resize(keep_units);
                ar->mapq.resize(keep_mapq); ar->ref_hap.resize(keep_ref);
            }};
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
            ar->jobs.reserve(ar->jobs.size() + 1);
            rollback.armed = false;  // nothing below throws
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
        std::lock_guard<std::mutex> run_lk(h->run_mu);  // device-touching calls on one handle are serialised (include/gpuphmm.h)
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
        std::lock_guard<std::mutex> run_lk(h->run_mu);
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
    std::unique_lock<std::mutex> run_lk;
    if (h) run_lk = std::unique_lock<std::mutex>(h->run_mu);
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
            for (const UnitSched &us : c.unit_sched) {  // the full-warp schedule (half-warp tasks run a second one with 16-step windows)
                out[2] += (int64_t)us.n_segs;
                for (uint32_t q = us.seg_first; q < us.seg_first + us.n_segs; ++q) {
                    const Segment &sg = c.segments[q];
                    out[3] += sg.n_free;
                    out[4] += sg.n_chk;
                    out[5] += sg.snap_pos != INT32_MIN;
                }
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

int gphmm_measure_fp32_peak(gphmm_t *h, double millis, double *tflops) {
    if (!h || !tflops) return GPHMM_ERR_INVALID_ARG;
    return guarded(h, [&]() -> int {
        std::lock_guard<std::mutex> run_lk(h->run_mu);
        Device &dev = *h->devices[0];
        CK(cudaSetDevice(dev.ordinal));
        constexpr int THREADS = 256, ILP = 8, ITER = 4096;
        const int grid = dev.n_sms * 8;  // 64 warps per SM: every scheduler always has an eligible warp
        DevBuf sink;
        sink.reserve((size_t)grid * THREADS * sizeof(float));
        cudaStream_t st = dev.streams[0];
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        auto run = [&](int reps) {
            for (int r = 0; r < reps; ++r) phmm_ffma_peak_kernel<ILP, ITER><<<grid, THREADS, 0, st>>>((float *)sink.p, 0.999f, 1e-3f);
        };
        run(2);
        CK(cudaEventRecord(e0, st));
        run(1);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms1 = 0.f;
        CK(cudaEventElapsedTime(&ms1, e0, e1));
        const int reps = std::max(1, (int)(millis / std::max(ms1, 1e-3f)));
        CK(cudaEventRecord(e0, st));
        run(reps);
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        CK(cudaGetLastError());
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        sink.release();
        *tflops = 2.0 * (double)grid * THREADS * ILP * ITER * reps / (ms * 1e-3) / 1e12;
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
