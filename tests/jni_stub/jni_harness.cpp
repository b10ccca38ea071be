// jni_harness.cpp -- executes gatk_b200/csrc/gpuphmm_jni.cpp without a JVM (tests only).
//
// The build image has no JDK, so the JNI shim -- the actual drop-in boundary for GATK -- could only be syntax-checked.
// This file implements the JNIEnv of tests/jni_stub/jni.h over a toy object model (byte[] / int[] / double[] / Object[]
// and holder objects with named fields, pending-exception slot) and drives the shim's Java_..._CudaPairHMMBinding_native*
// entry points the way CudaPairHMMBinding.java does, comparing every result with the same request made directly through
// the C ABI (include/gpuphmm.h).  No oracle is involved: shim and ABI must agree bit for bit.
//
//   jni_harness cpu   no GPU expected: nativeDeviceCount() == 0 and nativeCreate raises HardwareFeatureException
//   jni_harness gpu   full flow on cuda:0
#include <jni.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/gpuphmm.h"

struct _jfieldID {
    std::string name, sig;
};

namespace {

enum Kind { K_CLASS, K_HOLDER, K_BYTES, K_INTS, K_LONGS, K_DOUBLES, K_OBJECTS };

struct Fake {
    Kind kind;
    std::string class_name;                  // K_CLASS, K_HOLDER
    std::map<std::string, jobject> fields;   // K_HOLDER
    std::vector<jbyte> bytes;
    std::vector<jint> ints;
    std::vector<jlong> longs;
    std::vector<jdouble> doubles;
    std::vector<jobject> objects;
    _jobject handle;                         // the address handed to the shim
};

std::unordered_map<const void *, Fake *> g_heap;
std::vector<std::unique_ptr<Fake>> g_owned;
std::map<std::string, std::unique_ptr<_jfieldID>> g_field_ids;
std::string g_pending_class, g_pending_msg;
long g_local_refs_deleted = 0;

// classes the fake JVM knows, with their byte[] fields
const std::map<std::string, std::vector<std::string>> KNOWN_CLASSES = {
    {"org/broadinstitute/gatk/nativebindings/pairhmm/ReadDataHolder", {"readBases", "readQuals", "insertionGOP", "deletionGOP", "overallGCP"}},
    {"org/broadinstitute/gatk/nativebindings/pairhmm/HaplotypeDataHolder", {"haplotypeBases", "haplotypePDBases"}},
    {"java/lang/IllegalArgumentException", {}},
    {"java/lang/IllegalStateException", {}},
    {"java/lang/OutOfMemoryError", {}},
    {"org/broadinstitute/hellbender/exceptions/GATKException", {}},
    {"org/broadinstitute/hellbender/exceptions/UserException$HardwareFeatureException", {}},
};

Fake *make(Kind k) {
    g_owned.emplace_back(new Fake());
    Fake *f = g_owned.back().get();
    f->kind = k;
    g_heap[&f->handle] = f;
    return f;
}
Fake *deref(const void *p, Kind k, const char *what) {
    auto it = g_heap.find(p);
    if (it == g_heap.end() || it->second->kind != k) {
        fprintf(stderr, "harness: %s received a reference of the wrong type\n", what);
        exit(3);
    }
    return it->second;
}
Fake *deref_array(const void *p, const char *what) {
    auto it = g_heap.find(p);
    if (it == g_heap.end() || it->second->kind < K_BYTES) {
        fprintf(stderr, "harness: %s received a non-array reference\n", what);
        exit(3);
    }
    return it->second;
}
template <typename T> T as(Fake *f) { return reinterpret_cast<T>(&f->handle); }

jbyteArray new_bytes(const std::vector<uint8_t> &v) {
    Fake *f = make(K_BYTES);
    f->bytes.assign(v.begin(), v.end());
    return as<jbyteArray>(f);
}
jintArray new_ints(const std::vector<int32_t> &v) {
    Fake *f = make(K_INTS);
    f->ints.assign(v.begin(), v.end());
    return as<jintArray>(f);
}
jdoubleArray new_doubles(const std::vector<double> &v) {
    Fake *f = make(K_DOUBLES);
    f->doubles = v;
    return as<jdoubleArray>(f);
}
jobjectArray new_objects(const std::vector<jobject> &v) {
    Fake *f = make(K_OBJECTS);
    f->objects = v;
    return as<jobjectArray>(f);
}
jobject new_holder(const std::string &cls, const std::map<std::string, jobject> &fields) {
    Fake *f = make(K_HOLDER);
    f->class_name = cls;
    f->fields = fields;
    return as<jobject>(f);
}
std::vector<double> &doubles_of(jdoubleArray a) { return deref(a, K_DOUBLES, "doubles_of")->doubles; }
std::vector<jint> &ints_of(jintArray a) { return deref(a, K_INTS, "ints_of")->ints; }
std::vector<jbyte> &bytes_of(jbyteArray a) { return deref(a, K_BYTES, "bytes_of")->bytes; }

template <typename V, typename T> void get_region(V &v, jsize start, jsize len, T *dst, const char *what) {
    if (start < 0 || len < 0 || static_cast<size_t>(start) + static_cast<size_t>(len) > v.size()) {
        fprintf(stderr, "harness: %s out of bounds (ArrayIndexOutOfBoundsException in a JVM)\n", what);
        exit(3);
    }
    if (len) memcpy(dst, v.data() + start, sizeof(T) * static_cast<size_t>(len));
}
template <typename V, typename T> void set_region(V &v, jsize start, jsize len, const T *src, const char *what) {
    if (start < 0 || len < 0 || static_cast<size_t>(start) + static_cast<size_t>(len) > v.size()) {
        fprintf(stderr, "harness: %s out of bounds (ArrayIndexOutOfBoundsException in a JVM)\n", what);
        exit(3);
    }
    if (len) memcpy(v.data() + start, src, sizeof(T) * static_cast<size_t>(len));
}

}  // namespace

// ---- the JNIEnv of tests/jni_stub/jni.h -------------------------------------------------------------------------------
jclass JNIEnv::FindClass(const char *name) {
    if (!KNOWN_CLASSES.count(name)) {
        g_pending_class = "java/lang/NoClassDefFoundError";
        g_pending_msg = name;
        return nullptr;
    }
    for (auto &kv : g_heap)
        if (kv.second->kind == K_CLASS && kv.second->class_name == name) return as<jclass>(kv.second);
    Fake *f = make(K_CLASS);
    f->class_name = name;
    return as<jclass>(f);
}
jint JNIEnv::ThrowNew(jclass c, const char *msg) {
    g_pending_class = deref(c, K_CLASS, "ThrowNew")->class_name;
    g_pending_msg = msg ? msg : "";
    return 0;
}
jboolean JNIEnv::ExceptionCheck() { return g_pending_class.empty() ? JNI_FALSE : JNI_TRUE; }
jint JNIEnv::EnsureLocalCapacity(jint n) { return n >= 0 ? 0 : -1; }
jfieldID JNIEnv::GetFieldID(jclass c, const char *name, const char *sig) {
    Fake *cls = deref(c, K_CLASS, "GetFieldID");
    const auto &known = KNOWN_CLASSES.at(cls->class_name);
    bool found = false;
    for (const auto &f : known) found = found || f == name;
    if (!found || std::string(sig) != "[B") {
        g_pending_class = "java/lang/NoSuchFieldError";
        g_pending_msg = name;
        return nullptr;
    }
    const std::string key = cls->class_name + "." + name;
    auto &slot = g_field_ids[key];
    if (!slot) slot.reset(new _jfieldID{name, sig});
    return slot.get();
}
jobject JNIEnv::GetObjectField(jobject o, jfieldID id) {
    Fake *h = deref(o, K_HOLDER, "GetObjectField");
    auto it = h->fields.find(id->name);
    return it == h->fields.end() ? nullptr : it->second;
}
jsize JNIEnv::GetArrayLength(jarray a) {
    Fake *f = deref_array(a, "GetArrayLength");
    switch (f->kind) {
        case K_BYTES: return static_cast<jsize>(f->bytes.size());
        case K_INTS: return static_cast<jsize>(f->ints.size());
        case K_LONGS: return static_cast<jsize>(f->longs.size());
        case K_DOUBLES: return static_cast<jsize>(f->doubles.size());
        default: return static_cast<jsize>(f->objects.size());
    }
}
jobject JNIEnv::GetObjectArrayElement(jobjectArray a, jsize i) {
    Fake *f = deref(a, K_OBJECTS, "GetObjectArrayElement");
    if (i < 0 || static_cast<size_t>(i) >= f->objects.size()) {
        fprintf(stderr, "harness: GetObjectArrayElement out of bounds\n");
        exit(3);
    }
    return f->objects[static_cast<size_t>(i)];
}
void JNIEnv::DeleteLocalRef(jobject) { ++g_local_refs_deleted; }
void JNIEnv::GetByteArrayRegion(jbyteArray a, jsize s, jsize n, jbyte *dst) { get_region(bytes_of(a), s, n, dst, "GetByteArrayRegion"); }
void JNIEnv::SetByteArrayRegion(jbyteArray a, jsize s, jsize n, const jbyte *src) { set_region(bytes_of(a), s, n, src, "SetByteArrayRegion"); }
void JNIEnv::GetIntArrayRegion(jintArray a, jsize s, jsize n, jint *dst) { get_region(ints_of(a), s, n, dst, "GetIntArrayRegion"); }
void JNIEnv::SetIntArrayRegion(jintArray a, jsize s, jsize n, const jint *src) { set_region(ints_of(a), s, n, src, "SetIntArrayRegion"); }
void JNIEnv::GetDoubleArrayRegion(jdoubleArray a, jsize s, jsize n, jdouble *dst) { get_region(doubles_of(a), s, n, dst, "GetDoubleArrayRegion"); }
void JNIEnv::SetDoubleArrayRegion(jdoubleArray a, jsize s, jsize n, const jdouble *src) { set_region(doubles_of(a), s, n, src, "SetDoubleArrayRegion"); }
void JNIEnv::SetLongArrayRegion(jlongArray a, jsize s, jsize n, const jlong *src) { set_region(deref(a, K_LONGS, "SetLongArrayRegion")->longs, s, n, src, "SetLongArrayRegion"); }
jlongArray JNIEnv::NewLongArray(jsize n) {
    Fake *f = make(K_LONGS);
    f->longs.assign(static_cast<size_t>(n), 0);
    return as<jlongArray>(f);
}
jdoubleArray JNIEnv::NewDoubleArray(jsize n) {
    Fake *f = make(K_DOUBLES);
    f->doubles.assign(static_cast<size_t>(n), 0.0);
    return as<jdoubleArray>(f);
}

// ---- the shim's entry points (as javah would declare them for CudaPairHMMBinding's native methods) ----------------------
#define JNIFN(name) Java_org_broadinstitute_hellbender_utils_pairhmm_CudaPairHMMBinding_##name
extern "C" {
jint JNIFN(nativeDeviceCount)(JNIEnv *, jclass);
jlong JNIFN(nativeCreate)(JNIEnv *, jclass, jintArray, jboolean, jint);
void JNIFN(nativeCompute)(JNIEnv *, jclass, jlong, jobjectArray, jobjectArray, jdoubleArray);
void JNIFN(nativeComputePD)(JNIEnv *, jclass, jlong, jobjectArray, jobjectArray, jdoubleArray);
void JNIFN(nativeComputeRegion)(JNIEnv *, jclass, jlong, jobjectArray, jbyteArray, jobjectArray, jintArray, jdoubleArray, jdoubleArray, jbyteArray, jbyteArray);
jboolean JNIFN(nativeSwAlign)(JNIEnv *, jclass, jlong, jobjectArray, jobjectArray, jintArray, jint, jintArray, jintArray, jintArray);
jlong JNIFN(nativeSubmit)(JNIEnv *, jclass, jlong, jobjectArray, jobjectArray);
void JNIFN(nativeAwait)(JNIEnv *, jclass, jlong, jlong, jdoubleArray);
void JNIFN(nativeDestroy)(JNIEnv *, jclass, jlong);
jlongArray JNIFN(nativeCounters)(JNIEnv *, jclass, jlong);
}

namespace {

int g_failures = 0;
#define EXPECT(cond, ...)                                  \
    do {                                                   \
        if (!(cond)) {                                     \
            fprintf(stderr, "FAIL %s:%d: ", __FILE__, __LINE__); \
            fprintf(stderr, __VA_ARGS__);                  \
            fprintf(stderr, "\n");                         \
            ++g_failures;                                  \
        }                                                  \
    } while (0)

bool take_exception(const char *cls, const char *needle = nullptr) {
    const bool ok = g_pending_class == cls && (!needle || g_pending_msg.find(needle) != std::string::npos);
    if (!ok) fprintf(stderr, "  pending exception: '%s' (%s)\n", g_pending_class.c_str(), g_pending_msg.c_str());
    g_pending_class.clear();
    g_pending_msg.clear();
    return ok;
}

struct Lcg {
    uint64_t s;
    uint32_t next() { s = s * 6364136223846793005ULL + 1442695040888963407ULL; return static_cast<uint32_t>(s >> 33); }
    uint32_t below(uint32_t n) { return next() % n; }
};

struct Region {
    std::vector<std::vector<uint8_t>> read_bases, base_q, ins_q, del_q, gcp, haps;
};

Region make_region(uint64_t seed, int n_reads, int n_haps) {
    Lcg rng{seed};
    Region g;
    const char acgt[] = "ACGT";
    std::vector<uint8_t> root(220 + rng.below(120));
    for (auto &c : root) c = static_cast<uint8_t>(acgt[rng.below(4)]);
    for (int h = 0; h < n_haps; ++h) {
        std::vector<uint8_t> hap = root;
        for (int k = 0; k < h; ++k) hap[rng.below(static_cast<uint32_t>(hap.size()))] = static_cast<uint8_t>(acgt[rng.below(4)]);
        if (h % 3 == 2) hap.erase(hap.begin() + 50, hap.begin() + 53);
        g.haps.push_back(hap);
    }
    for (int r = 0; r < n_reads; ++r) {
        const std::vector<uint8_t> &hap = g.haps[rng.below(static_cast<uint32_t>(n_haps))];
        const uint32_t len = 30 + rng.below(121), start = rng.below(static_cast<uint32_t>(hap.size()) - len);
        std::vector<uint8_t> b(hap.begin() + start, hap.begin() + start + len), q(len), i(len, 45), d(len, 45), c(len, 10);
        for (uint32_t k = 0; k < len; ++k) {
            q[k] = static_cast<uint8_t>(6 + rng.below(36));
            if (rng.below(50) == 0) b[k] = static_cast<uint8_t>(acgt[rng.below(4)]);
            if (r % 4 == 1) { i[k] = static_cast<uint8_t>(20 + rng.below(26)); d[k] = static_cast<uint8_t>(20 + rng.below(26)); }
        }
        g.read_bases.push_back(b); g.base_q.push_back(q); g.ins_q.push_back(i); g.del_q.push_back(d); g.gcp.push_back(c);
    }
    return g;
}

jobjectArray java_reads(const Region &g) {
    std::vector<jobject> v;
    for (size_t r = 0; r < g.read_bases.size(); ++r)
        v.push_back(new_holder("org/broadinstitute/gatk/nativebindings/pairhmm/ReadDataHolder",
                               {{"readBases", new_bytes(g.read_bases[r])}, {"readQuals", new_bytes(g.base_q[r])},
                                {"insertionGOP", new_bytes(g.ins_q[r])}, {"deletionGOP", new_bytes(g.del_q[r])},
                                {"overallGCP", new_bytes(g.gcp[r])}}));
    return new_objects(v);
}
jobjectArray java_haps(const Region &g) {
    std::vector<jobject> v;
    for (const auto &h : g.haps)
        v.push_back(new_holder("org/broadinstitute/gatk/nativebindings/pairhmm/HaplotypeDataHolder", {{"haplotypeBases", new_bytes(h)}}));
    return new_objects(v);
}

// the same region as one flat unit for the C ABI
struct Flat {
    std::vector<uint8_t> rb, bq, iq, dq, gq, hb;
    std::vector<int64_t> ro{0}, ho{0};
    gphmm_unit unit;
    gphmm_batch b;
    explicit Flat(const Region &g) {
        for (size_t r = 0; r < g.read_bases.size(); ++r) {
            rb.insert(rb.end(), g.read_bases[r].begin(), g.read_bases[r].end());
            bq.insert(bq.end(), g.base_q[r].begin(), g.base_q[r].end());
            iq.insert(iq.end(), g.ins_q[r].begin(), g.ins_q[r].end());
            dq.insert(dq.end(), g.del_q[r].begin(), g.del_q[r].end());
            gq.insert(gq.end(), g.gcp[r].begin(), g.gcp[r].end());
            ro.push_back(static_cast<int64_t>(rb.size()));
        }
        for (const auto &h : g.haps) {
            hb.insert(hb.end(), h.begin(), h.end());
            ho.push_back(static_cast<int64_t>(hb.size()));
        }
        unit = {0, static_cast<int64_t>(g.read_bases.size()), 0, static_cast<int64_t>(g.haps.size()), 0};
        memset(&b, 0, sizeof b);
        b.read_bases = rb.data(); b.base_q = bq.data(); b.ins_q = iq.data(); b.del_q = dq.data(); b.gcp = gq.data();
        b.read_off = ro.data(); b.n_reads = unit.read_end;
        b.hap_bases = hb.data(); b.hap_off = ho.data(); b.n_haps = unit.hap_end;
        b.units = &unit; b.n_units = 1;
    }
};

int run_cpu(JNIEnv *env) {
    EXPECT(JNIFN(nativeDeviceCount)(env, nullptr) == 0, "a GPU is visible: run the gpu mode");
    const jlong h = JNIFN(nativeCreate)(env, nullptr, nullptr, JNI_FALSE, 4);
    EXPECT(h == 0, "nativeCreate returned a handle without a GPU");
    EXPECT(take_exception("org/broadinstitute/hellbender/exceptions/UserException$HardwareFeatureException", "libgpuphmm"),
           "nativeCreate without a GPU must raise HardwareFeatureException");
    JNIFN(nativeDestroy)(env, nullptr, 0);   // done() before initialize(): a no-op
    EXPECT(!env->ExceptionCheck(), "nativeDestroy(0) raised");
    return g_failures;
}

int run_gpu(JNIEnv *env) {
    EXPECT(JNIFN(nativeDeviceCount)(env, nullptr) >= 1, "no usable GPU");
    const jlong h = JNIFN(nativeCreate)(env, nullptr, new_ints({0}), JNI_FALSE, 4);
    EXPECT(h != 0 && !env->ExceptionCheck(), "nativeCreate failed");
    if (!h) return ++g_failures;
    gphmm_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = static_cast<int32_t>(sizeof cfg);
    gphmm_t *direct = nullptr;
    EXPECT(gphmm_create(&cfg, &direct) == GPHMM_OK, "gphmm_create failed");
    if (!direct) return ++g_failures;

    // computeLikelihoods: shim == C ABI, bit for bit; results are log10 likelihoods (<= 0)
    const Region g = make_region(1, 24, 5);
    Flat f(g);
    const size_t n = g.read_bases.size() * g.haps.size();
    std::vector<double> want(n, 1.0);
    EXPECT(gphmm_compute(direct, &f.b, want.data()) == GPHMM_OK, "gphmm_compute: %s", gphmm_last_error(direct));
    jdoubleArray out = new_doubles(std::vector<double>(n, 1.0));
    JNIFN(nativeCompute)(env, nullptr, h, java_reads(g), java_haps(g), out);
    EXPECT(!env->ExceptionCheck(), "nativeCompute raised %s: %s", g_pending_class.c_str(), g_pending_msg.c_str());
    EXPECT(memcmp(doubles_of(out).data(), want.data(), n * sizeof(double)) == 0, "nativeCompute differs from gphmm_compute");
    for (double v : want) EXPECT(v <= 0.0 && v > -400.0, "implausible log10 likelihood %g", v);

    // submit / await: three regions in flight, collected in order
    std::vector<Region> regions = {make_region(2, 9, 3), make_region(3, 17, 7), make_region(4, 1, 1)};
    std::vector<jlong> tickets;
    for (const auto &r : regions) tickets.push_back(JNIFN(nativeSubmit)(env, nullptr, h, java_reads(r), java_haps(r)));
    for (size_t k = 0; k < regions.size(); ++k) {
        Flat fk(regions[k]);
        const size_t nk = regions[k].read_bases.size() * regions[k].haps.size();
        std::vector<double> wk(nk);
        EXPECT(gphmm_compute(direct, &fk.b, wk.data()) == GPHMM_OK, "gphmm_compute");
        jdoubleArray ok = new_doubles(std::vector<double>(nk, 1.0));
        EXPECT(tickets[k] != 0, "nativeSubmit returned 0");
        JNIFN(nativeAwait)(env, nullptr, h, tickets[k], ok);
        EXPECT(!env->ExceptionCheck(), "nativeAwait raised");
        double worst = 0;
        for (size_t i = 0; i < nk; ++i) worst = std::max(worst, std::abs(doubles_of(ok)[i] - wk[i]));
        EXPECT(worst < 1e-5, "queued region %zu differs from the synchronous result by %g", k, worst);   // merged batches: float noise only
    }
    JNIFN(nativeAwait)(env, nullptr, h, tickets[0], new_doubles({0.0}));
    EXPECT(take_exception("java/lang/IllegalArgumentException", "unknown ticket"), "second await of a ticket must raise");

    // error conventions: ragged holder -> IllegalArgumentException (PairHMM.java:286-292), quality > 127 -> IllegalArgumentException
    Region bad = make_region(5, 3, 2);
    bad.base_q[1].pop_back();
    JNIFN(nativeCompute)(env, nullptr, h, java_reads(bad), java_haps(bad), new_doubles(std::vector<double>(6)));
    EXPECT(take_exception("java/lang/IllegalArgumentException", "same size"), "ragged read arrays must raise IllegalArgumentException");
    Region badq = make_region(6, 3, 2);
    badq.ins_q[0][0] = 200;
    JNIFN(nativeCompute)(env, nullptr, h, java_reads(badq), java_haps(badq), new_doubles(std::vector<double>(6)));
    EXPECT(take_exception("java/lang/IllegalArgumentException"), "insertion quality 200 must raise IllegalArgumentException");
    JNIFN(nativeCompute)(env, nullptr, h, java_reads(g), java_haps(g), new_doubles({0.0}));
    EXPECT(take_exception("java/lang/IllegalArgumentException", "too small"), "short output array must raise");
    // empty read list: returns without touching the output (VectorLoglessPairHMM.java:113-115)
    Region none = make_region(7, 0, 2);
    jdoubleArray untouched = new_doubles({7.0});
    JNIFN(nativeCompute)(env, nullptr, h, java_reads(none), java_haps(none), untouched);
    EXPECT(!env->ExceptionCheck() && doubles_of(untouched)[0] == 7.0, "empty read list must be a no-op");

    // region steps: shim == gphmm_compute_regions
    {
        const size_t nr = g.read_bases.size(), nh = g.haps.size();
        std::vector<uint8_t> mapq(nr, 60), keep_want(nr, 1), hq_want(f.rb.size());
        mapq[3] = 20;
        const int32_t ref = 0;
        gphmm_region_steps rs;
        memset(&rs, 0, sizeof rs);
        rs.struct_size = static_cast<int32_t>(sizeof rs);
        rs.flags = GPHMM_RS_FILTER_POORLY;
        rs.pcr_rate_factor = 3.0; rs.base_quality_score_threshold = 18;
        rs.log10_global_read_mismapping_rate = -4.5; rs.expected_error_rate_per_base = 0.02; rs.read_disqualification_scale = 1.0;
        rs.mapq = mapq.data(); rs.ref_hap = &ref; rs.keep = keep_want.data(); rs.hmm_base_q = hq_want.data();
        std::vector<double> lk_want(nr * nh);
        EXPECT(gphmm_compute_regions(direct, &f.b, &rs, lk_want.data()) == GPHMM_OK, "gphmm_compute_regions: %s", gphmm_last_error(direct));
        jdoubleArray lk = new_doubles(std::vector<double>(nr * nh, 1.0));
        jbyteArray keep = new_bytes(std::vector<uint8_t>(nr, 1)), hq = new_bytes(std::vector<uint8_t>(f.rb.size(), 0));
        JNIFN(nativeComputeRegion)(env, nullptr, h, java_reads(g), new_bytes(mapq), java_haps(g), new_ints({GPHMM_RS_FILTER_POORLY, 18, 0}),
                                   new_doubles({3.0, -4.5, 0.02, 1.0}), lk, keep, hq);
        EXPECT(!env->ExceptionCheck(), "nativeComputeRegion raised %s: %s", g_pending_class.c_str(), g_pending_msg.c_str());
        EXPECT(memcmp(doubles_of(lk).data(), lk_want.data(), lk_want.size() * sizeof(double)) == 0, "region-step likelihoods differ");
        EXPECT(memcmp(bytes_of(keep).data(), keep_want.data(), nr) == 0, "keep flags differ");
        EXPECT(memcmp(bytes_of(hq).data(), hq_want.data(), hq_want.size()) == 0, "HMM base qualities differ");
    }

    // Smith-Waterman: shim == gphmm_sw_align
    {
        const int np = 6, cap = 32;
        std::vector<jobject> refs, alts;
        std::vector<uint8_t> rb, ab;
        std::vector<int64_t> ro{0}, ao{0};
        for (int k = 0; k < np; ++k) {
            const auto &ref = g.haps[static_cast<size_t>(k) % g.haps.size()];
            const auto &alt = g.read_bases[static_cast<size_t>(k)];
            refs.push_back(new_bytes(ref)); alts.push_back(new_bytes(alt));
            rb.insert(rb.end(), ref.begin(), ref.end()); ro.push_back(static_cast<int64_t>(rb.size()));
            ab.insert(ab.end(), alt.begin(), alt.end()); ao.push_back(static_cast<int64_t>(ab.size()));
        }
        gphmm_sw_params prm = {static_cast<int32_t>(sizeof(gphmm_sw_params)), 10, -15, -30, -5, GPHMM_SW_SOFTCLIP};
        gphmm_sw_batch sb = {rb.data(), ro.data(), ab.data(), ao.data(), np};
        std::vector<int32_t> off_want(np), ne_want(np);
        std::vector<uint32_t> el_want(static_cast<size_t>(np) * cap);
        EXPECT(gphmm_sw_align(direct, &sb, &prm, cap, off_want.data(), ne_want.data(), el_want.data()) == GPHMM_OK, "gphmm_sw_align");
        jintArray off = new_ints(std::vector<int32_t>(np, -7)), ne = new_ints(std::vector<int32_t>(np, -7)), el = new_ints(std::vector<int32_t>(static_cast<size_t>(np) * cap, 0));
        const jboolean fit = JNIFN(nativeSwAlign)(env, nullptr, h, new_objects(refs), new_objects(alts), new_ints({10, -15, -30, -5, GPHMM_SW_SOFTCLIP}), cap, off, ne, el);
        EXPECT(fit == JNI_TRUE && !env->ExceptionCheck(), "nativeSwAlign failed");
        EXPECT(memcmp(ints_of(off).data(), off_want.data(), np * sizeof(int32_t)) == 0, "alignment offsets differ");
        EXPECT(memcmp(ints_of(ne).data(), ne_want.data(), np * sizeof(int32_t)) == 0, "CIGAR lengths differ");
        for (int k = 0; k < np; ++k)
            EXPECT(ne_want[k] >= 1 && memcmp(ints_of(el).data() + k * cap, el_want.data() + static_cast<size_t>(k) * cap, static_cast<size_t>(ne_want[k]) * 4) == 0, "CIGAR %d differs", k);
    }

    // PD-HMM: HaplotypeDataHolder.haplotypePDBases travels as the flag array; shim == gphmm_pd_compute
    {
        const Region pg = make_region(8, 6, 2);
        Flat pf(pg);
        std::vector<uint8_t> pd(pf.hb.size(), 0);
        pd[40] = 1 | 8;            // SNP site, alternative A
        pd[80] = 2; pd[85] = 4;    // DEL_START .. DEL_END
        std::vector<jobject> holders;
        for (size_t k = 0; k < pg.haps.size(); ++k)
            holders.push_back(new_holder("org/broadinstitute/gatk/nativebindings/pairhmm/HaplotypeDataHolder",
                                         {{"haplotypeBases", new_bytes(pg.haps[k])},
                                          {"haplotypePDBases", new_bytes(std::vector<uint8_t>(pd.begin() + pf.ho[k], pd.begin() + pf.ho[k + 1]))}}));
        const size_t np = pg.read_bases.size() * pg.haps.size();
        std::vector<double> pd_want(np, 1.0);
        EXPECT(gphmm_pd_compute(direct, &pf.b, pd.data(), pd_want.data()) == GPHMM_OK, "gphmm_pd_compute: %s", gphmm_last_error(direct));
        jdoubleArray pd_out = new_doubles(std::vector<double>(np, 1.0));
        JNIFN(nativeComputePD)(env, nullptr, h, java_reads(pg), new_objects(holders), pd_out);
        EXPECT(!env->ExceptionCheck(), "nativeComputePD raised %s: %s", g_pending_class.c_str(), g_pending_msg.c_str());
        EXPECT(memcmp(doubles_of(pd_out).data(), pd_want.data(), np * sizeof(double)) == 0, "nativeComputePD differs from gphmm_pd_compute");
        // a holder without the flag array is an argument error, not a crash
        JNIFN(nativeComputePD)(env, nullptr, h, java_reads(pg), java_haps(pg), new_doubles(std::vector<double>(np)));
        EXPECT(take_exception("java/lang/IllegalArgumentException", "haplotypePDBases"), "missing haplotypePDBases must raise");
    }

    // counters: pairs and launches were counted
    jlongArray counters = JNIFN(nativeCounters)(env, nullptr, h);
    Fake *c = deref(counters, K_LONGS, "counters");
    EXPECT(c->longs.size() == 6 && c->longs[0] >= static_cast<jlong>(n) && c->longs[5] > 0, "counters: pairs %ld launches %ld", (long)c->longs[0], (long)c->longs[5]);
    EXPECT(g_local_refs_deleted > 0, "the shim never released a local reference");

    JNIFN(nativeDestroy)(env, nullptr, h);
    gphmm_destroy(direct);
    return g_failures;
}

}  // namespace

int main(int argc, char **argv) {
    JNIEnv env;
    const std::string mode = argc > 1 ? argv[1] : "cpu";
    const int failures = mode == "gpu" ? run_gpu(&env) : run_cpu(&env);
    if (failures == 0) printf("jni_harness %s: ok\n", mode.c_str());
    return failures == 0 ? 0 : 1;
}
