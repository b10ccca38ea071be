// gpuphmm_jni.cpp -- JNI shim between org.broadinstitute.hellbender.utils.pairhmm.CudaPairHMMBinding (java/) and the
// C ABI of libgpuphmm (include/gpuphmm.h).  Pure marshalling: Java holder arrays -> one flat gphmm_batch unit in
// pinned staging memory -> gphmm_compute / gphmm_submit -> double[] back.  No arithmetic here.
//
// Replaces the native half of the GKL binding the reference calls at
// src/main/java/org/broadinstitute/hellbender/utils/pairhmm/VectorLoglessPairHMM.java:138
// (PairHMMNativeBinding.computeLikelihoods(ReadDataHolder[], HaplotypeDataHolder[], double[])).
//
// Built only where a JDK is present (`make jni JAVA_HOME=...`); this image has none, so here it is checked for
// syntax against tests/jni_stub/jni.h (tests/test_abi_cpu.py).
#include <jni.h>

#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gpuphmm.h"

namespace {

struct FieldIds {
    bool ready = false;
    jfieldID readBases, readQuals, insertionGOP, deletionGOP, overallGCP, haplotypeBases;
    jfieldID haplotypePDBases = nullptr;  // present in the holders the PD-HMM binding uses (VectorLoglessPairPDHMM.java:104)
};
FieldIds g_ids;
std::mutex g_mu;

// growable pinned buffer (gphmm_host_alloc), one set per handle
struct Pinned {
    uint8_t *p = nullptr;
    size_t cap = 0;
    bool reserve(size_t n) {
        if (n <= cap) return true;
        size_t want = n + n / 2 + 4096;
        uint8_t *q = static_cast<uint8_t *>(gphmm_host_alloc(want));
        if (!q) return false;
        if (p) gphmm_host_free(p);
        p = q;
        cap = want;
        return true;
    }
    ~Pinned() { if (p) gphmm_host_free(p); }
};

struct Pending {
    std::vector<double> out;  // never empty (gphmm_submit rejects a null out); `n` is what the caller gets back
    size_t n = 0;
};

struct Session {
    gphmm_t *h = nullptr;
    Pinned bases, q, iq, dq, gcp, haps, pd;
    std::vector<int64_t> read_off, hap_off;
    std::map<uint64_t, Pending> pending;
};

void throw_java(JNIEnv *env, const char *cls, const std::string &msg) {
    jclass c = env->FindClass(cls);
    if (c) env->ThrowNew(c, msg.c_str());
}

void throw_for(JNIEnv *env, Session *s, int rc) {
    const std::string msg = std::string("libgpuphmm: ") + (s && s->h ? gphmm_last_error(s->h) : gphmm_strerror(rc));
    switch (rc) {
        case GPHMM_ERR_INVALID_ARG:
        case GPHMM_ERR_BAD_QUAL:  // PairHMMModel.java:109-111 / PairHMM.java:286-292 throw IllegalArgumentException
            throw_java(env, "java/lang/IllegalArgumentException", msg);
            break;
        case GPHMM_ERR_NOMEM:
            throw_java(env, "java/lang/OutOfMemoryError", msg);
            break;
        default:
            throw_java(env, "org/broadinstitute/hellbender/exceptions/GATKException", msg);
    }
}

bool resolve_fields(JNIEnv *env, jobjectArray reads, jobjectArray haps) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_ids.ready) return true;
    jclass rc = env->FindClass("org/broadinstitute/gatk/nativebindings/pairhmm/ReadDataHolder");
    jclass hc = env->FindClass("org/broadinstitute/gatk/nativebindings/pairhmm/HaplotypeDataHolder");
    (void)reads; (void)haps;
    if (!rc || !hc) return false;
    g_ids.readBases = env->GetFieldID(rc, "readBases", "[B");
    g_ids.readQuals = env->GetFieldID(rc, "readQuals", "[B");
    g_ids.insertionGOP = env->GetFieldID(rc, "insertionGOP", "[B");
    g_ids.deletionGOP = env->GetFieldID(rc, "deletionGOP", "[B");
    g_ids.overallGCP = env->GetFieldID(rc, "overallGCP", "[B");
    g_ids.haplotypeBases = env->GetFieldID(hc, "haplotypeBases", "[B");
    g_ids.ready = g_ids.readBases && g_ids.readQuals && g_ids.insertionGOP && g_ids.deletionGOP && g_ids.overallGCP && g_ids.haplotypeBases;
    return g_ids.ready;
}

// Flatten the holder arrays into the session's pinned SoA buffers.  Returns false with a Java exception pending.
bool pack(JNIEnv *env, Session *s, jobjectArray reads, jobjectArray haps, gphmm_batch *b, gphmm_unit *unit) {
    if (!resolve_fields(env, reads, haps)) {
        if (!env->ExceptionCheck()) throw_java(env, "java/lang/IllegalStateException", "cannot resolve ReadDataHolder/HaplotypeDataHolder fields");
        return false;
    }
    const jsize n_reads = env->GetArrayLength(reads), n_haps = env->GetArrayLength(haps);
    // the array references of every holder stay alive until pass 2: far more than the 16 local references JNI guarantees
    if (env->EnsureLocalCapacity(5 * n_reads + n_haps + 16) != 0) return false;  // OutOfMemoryError is pending
    s->read_off.assign(1, 0);
    s->hap_off.assign(1, 0);
    // pass 1: lengths
    std::vector<jbyteArray> rb(n_reads), rq(n_reads), ri(n_reads), rd(n_reads), rg(n_reads), hb(n_haps);
    for (jsize r = 0; r < n_reads; ++r) {
        jobject o = env->GetObjectArrayElement(reads, r);
        rb[r] = static_cast<jbyteArray>(env->GetObjectField(o, g_ids.readBases));
        rq[r] = static_cast<jbyteArray>(env->GetObjectField(o, g_ids.readQuals));
        ri[r] = static_cast<jbyteArray>(env->GetObjectField(o, g_ids.insertionGOP));
        rd[r] = static_cast<jbyteArray>(env->GetObjectField(o, g_ids.deletionGOP));
        rg[r] = static_cast<jbyteArray>(env->GetObjectField(o, g_ids.overallGCP));
        env->DeleteLocalRef(o);
        if (!rb[r] || !rq[r] || !ri[r] || !rd[r] || !rg[r]) {
            throw_java(env, "java/lang/IllegalArgumentException", "null array in ReadDataHolder");
            return false;
        }
        const jsize len = env->GetArrayLength(rb[r]);
        if (env->GetArrayLength(rq[r]) != len || env->GetArrayLength(ri[r]) != len || env->GetArrayLength(rd[r]) != len ||
            env->GetArrayLength(rg[r]) != len) {
            // PairHMM.java:286-292
            throw_java(env, "java/lang/IllegalArgumentException", "Read bases and quals aren't the same size");
            return false;
        }
        s->read_off.push_back(s->read_off.back() + len);
    }
    for (jsize h = 0; h < n_haps; ++h) {
        jobject o = env->GetObjectArrayElement(haps, h);
        hb[h] = static_cast<jbyteArray>(env->GetObjectField(o, g_ids.haplotypeBases));
        env->DeleteLocalRef(o);
        if (!hb[h]) {
            throw_java(env, "java/lang/IllegalArgumentException", "haplotypeBases may not be null");
            return false;
        }
        s->hap_off.push_back(s->hap_off.back() + env->GetArrayLength(hb[h]));
    }
    const size_t nb = static_cast<size_t>(s->read_off.back()), nh = static_cast<size_t>(s->hap_off.back());
    if (!s->bases.reserve(nb) || !s->q.reserve(nb) || !s->iq.reserve(nb) || !s->dq.reserve(nb) || !s->gcp.reserve(nb) || !s->haps.reserve(nh)) {
        throw_java(env, "java/lang/OutOfMemoryError", "pinned staging allocation failed");
        return false;
    }
    // pass 2: one copy per array, straight into pinned memory
    for (jsize r = 0; r < n_reads; ++r) {
        const jsize off = static_cast<jsize>(s->read_off[r]), len = static_cast<jsize>(s->read_off[r + 1] - s->read_off[r]);
        env->GetByteArrayRegion(rb[r], 0, len, reinterpret_cast<jbyte *>(s->bases.p + off));
        env->GetByteArrayRegion(rq[r], 0, len, reinterpret_cast<jbyte *>(s->q.p + off));
        env->GetByteArrayRegion(ri[r], 0, len, reinterpret_cast<jbyte *>(s->iq.p + off));
        env->GetByteArrayRegion(rd[r], 0, len, reinterpret_cast<jbyte *>(s->dq.p + off));
        env->GetByteArrayRegion(rg[r], 0, len, reinterpret_cast<jbyte *>(s->gcp.p + off));
        env->DeleteLocalRef(rb[r]); env->DeleteLocalRef(rq[r]); env->DeleteLocalRef(ri[r]); env->DeleteLocalRef(rd[r]); env->DeleteLocalRef(rg[r]);
    }
    for (jsize h = 0; h < n_haps; ++h) {
        env->GetByteArrayRegion(hb[h], 0, static_cast<jsize>(s->hap_off[h + 1] - s->hap_off[h]), reinterpret_cast<jbyte *>(s->haps.p + s->hap_off[h]));
        env->DeleteLocalRef(hb[h]);
    }
    std::memset(b, 0, sizeof *b);
    b->read_bases = s->bases.p; b->base_q = s->q.p; b->ins_q = s->iq.p; b->del_q = s->dq.p; b->gcp = s->gcp.p;
    b->read_off = s->read_off.data(); b->n_reads = n_reads;
    b->hap_bases = s->haps.p; b->hap_off = s->hap_off.data(); b->n_haps = n_haps;
    unit->read_begin = 0; unit->read_end = n_reads; unit->hap_begin = 0; unit->hap_end = n_haps; unit->out_off = 0;
    b->units = unit; b->n_units = 1;
    return true;
}

}  // namespace

#define JNIFN(name) Java_org_broadinstitute_hellbender_utils_pairhmm_CudaPairHMMBinding_##name

extern "C" {

JNIEXPORT jint JNICALL JNIFN(nativeDeviceCount)(JNIEnv *, jclass) { return gphmm_device_count(); }

JNIEXPORT jlong JNICALL JNIFN(nativeCreate)(JNIEnv *env, jclass, jintArray devices, jboolean force_fp64, jint host_threads) {
    gphmm_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = static_cast<int32_t>(sizeof cfg);
    std::vector<int32_t> dev;
    if (devices) {
        const jsize n = env->GetArrayLength(devices);
        dev.resize(n);
        if (n) env->GetIntArrayRegion(devices, 0, n, reinterpret_cast<jint *>(dev.data()));
    }
    cfg.n_devices = static_cast<int32_t>(dev.size());
    cfg.devices = dev.empty() ? nullptr : dev.data();
    cfg.force_fp64 = force_fp64 ? 1 : 0;
    cfg.host_threads = host_threads;
    Session *s = new Session();
    const int rc = gphmm_create(&cfg, &s->h);
    if (rc != GPHMM_OK) {
        delete s;
        throw_java(env, rc == GPHMM_ERR_NO_DEVICE ? "org/broadinstitute/hellbender/exceptions/UserException$HardwareFeatureException"
                                                  : "org/broadinstitute/hellbender/exceptions/GATKException",
                   std::string("libgpuphmm: ") + gphmm_strerror(rc));
        return 0;
    }
    return reinterpret_cast<jlong>(s);
}

JNIEXPORT void JNICALL JNIFN(nativeCompute)(JNIEnv *env, jclass, jlong handle, jobjectArray reads, jobjectArray haps, jdoubleArray out) {
    Session *s = reinterpret_cast<Session *>(handle);
    gphmm_batch b;
    gphmm_unit unit;
    if (!pack(env, s, reads, haps, &b, &unit)) return;
    const jsize need = static_cast<jsize>(b.n_reads * b.n_haps);
    if (env->GetArrayLength(out) < need) {
        throw_java(env, "java/lang/IllegalArgumentException", "likelihood array is too small");
        return;
    }
    if (need == 0) return;
    std::vector<double> res(static_cast<size_t>(need));
    const int rc = gphmm_compute(s->h, &b, res.data());
    if (rc != GPHMM_OK) {
        throw_for(env, s, rc);
        return;
    }
    env->SetDoubleArrayRegion(out, 0, need, res.data());
}

// Region steps (include/gpuphmm.h gphmm_compute_regions): the reads arrive as modifyReadQualities receives them.
// iparams = {flags, baseQualityScoreThreshold, refHaplotypeIndex}, dparams = {pcrRateFactor,
// log10GlobalReadMismappingRate, expectedErrorRatePerBase, readDisqualificationScale}.  out is allele-major
// (out[h * nReads + r], normalised), keep[r] = 0 for reads filterPoorlyModeledEvidence removes, hmmBaseQuals (nullable)
// receives the modified base qualities of all reads back to back (HMM_BASE_QUALITIES_TAG).
JNIEXPORT void JNICALL JNIFN(nativeComputeRegion)(JNIEnv *env, jclass, jlong handle, jobjectArray reads, jbyteArray mapq,
                                                  jobjectArray haps, jintArray iparams, jdoubleArray dparams, jdoubleArray out,
                                                  jbyteArray keep, jbyteArray hmmBaseQuals) {
    Session *s = reinterpret_cast<Session *>(handle);
    gphmm_batch b;
    gphmm_unit unit;
    if (!pack(env, s, reads, haps, &b, &unit)) return;
    const jsize n_reads = static_cast<jsize>(b.n_reads), need = static_cast<jsize>(b.n_reads * b.n_haps);
    if (env->GetArrayLength(iparams) < 3 || env->GetArrayLength(dparams) < 4 || env->GetArrayLength(mapq) < n_reads ||
        env->GetArrayLength(keep) < n_reads || env->GetArrayLength(out) < need ||
        (hmmBaseQuals && env->GetArrayLength(hmmBaseQuals) < static_cast<jsize>(s->read_off.back()))) {
        throw_java(env, "java/lang/IllegalArgumentException", "region-step arrays are too small");
        return;
    }
    jint ip[3];
    jdouble dp[4];
    env->GetIntArrayRegion(iparams, 0, 3, ip);
    env->GetDoubleArrayRegion(dparams, 0, 4, dp);
    std::vector<uint8_t> mq(static_cast<size_t>(n_reads) + 1), kp(static_cast<size_t>(n_reads) + 1, 1);
    std::vector<uint8_t> hq(hmmBaseQuals ? static_cast<size_t>(s->read_off.back()) + 1 : 0);
    if (n_reads) env->GetByteArrayRegion(mapq, 0, n_reads, reinterpret_cast<jbyte *>(mq.data()));
    const int32_t ref = ip[2];
    gphmm_region_steps rs;
    std::memset(&rs, 0, sizeof rs);
    rs.struct_size = static_cast<int32_t>(sizeof rs);
    rs.flags = ip[0];
    rs.base_quality_score_threshold = ip[1];
    rs.pcr_rate_factor = dp[0];
    rs.log10_global_read_mismapping_rate = dp[1];
    rs.expected_error_rate_per_base = dp[2];
    rs.read_disqualification_scale = dp[3];
    rs.mapq = mq.data();
    rs.ref_hap = &ref;
    rs.keep = kp.data();
    rs.hmm_base_q = hmmBaseQuals ? hq.data() : nullptr;
    rs.raw_lk = nullptr;  // the plugin writes --pair-hmm-results-file from the plain call only
    std::vector<double> res(static_cast<size_t>(need) + 1);
    const int rc = gphmm_compute_regions(s->h, &b, &rs, res.data());
    if (rc != GPHMM_OK) {
        throw_for(env, s, rc);
        return;
    }
    if (need) env->SetDoubleArrayRegion(out, 0, need, res.data());
    if (n_reads) env->SetByteArrayRegion(keep, 0, n_reads, reinterpret_cast<const jbyte *>(kp.data()));
    if (hmmBaseQuals && s->read_off.back()) env->SetByteArrayRegion(hmmBaseQuals, 0, static_cast<jsize>(s->read_off.back()), reinterpret_cast<const jbyte *>(hq.data()));
}

// PD-HMM (include/gpuphmm.h gphmm_pd_compute): like nativeCompute, with HaplotypeDataHolder.haplotypePDBases
// (VectorLoglessPairPDHMM.java:100-105) handed over as the flag array parallel to the haplotype bases.
JNIEXPORT void JNICALL JNIFN(nativeComputePD)(JNIEnv *env, jclass, jlong handle, jobjectArray reads, jobjectArray haps, jdoubleArray out) {
    Session *s = reinterpret_cast<Session *>(handle);
    gphmm_batch b;
    gphmm_unit unit;
    if (!pack(env, s, reads, haps, &b, &unit)) return;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        if (!g_ids.haplotypePDBases) {
            jclass hc = env->FindClass("org/broadinstitute/gatk/nativebindings/pairhmm/HaplotypeDataHolder");
            if (hc) g_ids.haplotypePDBases = env->GetFieldID(hc, "haplotypePDBases", "[B");
        }
    }
    if (!g_ids.haplotypePDBases) {
        if (!env->ExceptionCheck()) throw_java(env, "java/lang/IllegalStateException", "HaplotypeDataHolder has no haplotypePDBases field");
        return;
    }
    if (!s->pd.reserve(static_cast<size_t>(s->hap_off.back()) + 1)) {
        throw_java(env, "java/lang/OutOfMemoryError", "pinned staging allocation failed");
        return;
    }
    for (jsize h = 0; h < static_cast<jsize>(b.n_haps); ++h) {
        jobject o = env->GetObjectArrayElement(haps, h);
        jbyteArray pd = static_cast<jbyteArray>(env->GetObjectField(o, g_ids.haplotypePDBases));
        env->DeleteLocalRef(o);
        const jsize len = static_cast<jsize>(s->hap_off[h + 1] - s->hap_off[h]);
        if (!pd || env->GetArrayLength(pd) != len) {
            throw_java(env, "java/lang/IllegalArgumentException", "haplotypePDBases must have one byte per haplotype base");
            return;
        }
        env->GetByteArrayRegion(pd, 0, len, reinterpret_cast<jbyte *>(s->pd.p + s->hap_off[h]));
        env->DeleteLocalRef(pd);
    }
    const jsize need = static_cast<jsize>(b.n_reads * b.n_haps);
    if (env->GetArrayLength(out) < need) {
        throw_java(env, "java/lang/IllegalArgumentException", "likelihood array is too small");
        return;
    }
    if (need == 0) return;
    std::vector<double> res(static_cast<size_t>(need));
    const int rc = gphmm_pd_compute(s->h, &b, s->pd.p, res.data());
    if (rc != GPHMM_OK) {
        throw_for(env, s, rc);
        return;
    }
    env->SetDoubleArrayRegion(out, 0, need, res.data());
}

// Smith-Waterman (include/gpuphmm.h gphmm_sw_align): pair k aligns alts[k] to refs[k].  params = {match, mismatch, gapOpen,
// gapExtend, overhangStrategy ordinal as GPHMM_SW_*}.  offsets / nElems: one int per pair; elems: capacity ints per
// pair, (length << 4) | op.  Returns false when some CIGAR needs more than `capacity` elements (nElems = -1 there).
JNIEXPORT jboolean JNICALL JNIFN(nativeSwAlign)(JNIEnv *env, jclass, jlong handle, jobjectArray refs, jobjectArray alts, jintArray params,
                                                jint capacity, jintArray offsets, jintArray nElems, jintArray elems) {
    Session *s = reinterpret_cast<Session *>(handle);
    const jsize n = env->GetArrayLength(refs);
    if (env->GetArrayLength(alts) != n || env->GetArrayLength(params) < 5 || env->GetArrayLength(offsets) < n || env->GetArrayLength(nElems) < n ||
        capacity < 1 || env->GetArrayLength(elems) < n * capacity) {
        throw_java(env, "java/lang/IllegalArgumentException", "Smith-Waterman batch arrays have inconsistent sizes");
        return JNI_FALSE;
    }
    std::vector<uint8_t> rb, ab;
    std::vector<int64_t> ro(1, 0), ao(1, 0);
    for (int side = 0; side < 2; ++side) {
        std::vector<uint8_t> &bytes = side ? ab : rb;
        std::vector<int64_t> &off = side ? ao : ro;
        for (jsize k = 0; k < n; ++k) {
            jbyteArray a = static_cast<jbyteArray>(env->GetObjectArrayElement(side ? alts : refs, k));
            if (!a) {
                throw_java(env, "java/lang/IllegalArgumentException", "Non-null, non-empty sequences are required for the Smith-Waterman calculation");
                return JNI_FALSE;
            }
            const jsize len = env->GetArrayLength(a);
            bytes.resize(bytes.size() + static_cast<size_t>(len));
            if (len) env->GetByteArrayRegion(a, 0, len, reinterpret_cast<jbyte *>(bytes.data() + off.back()));
            off.push_back(off.back() + len);
            env->DeleteLocalRef(a);
        }
    }
    jint p[5];
    env->GetIntArrayRegion(params, 0, 5, p);
    gphmm_sw_params prm;
    prm.struct_size = static_cast<int32_t>(sizeof prm);
    prm.match_value = p[0]; prm.mismatch_penalty = p[1]; prm.gap_open_penalty = p[2]; prm.gap_extend_penalty = p[3]; prm.overhang_strategy = p[4];
    gphmm_sw_batch b;
    b.ref_bases = rb.data(); b.ref_off = ro.data(); b.alt_bases = ab.data(); b.alt_off = ao.data(); b.n_pairs = n;
    std::vector<int32_t> off_out(static_cast<size_t>(n) + 1), ne(static_cast<size_t>(n) + 1);
    std::vector<uint32_t> el(static_cast<size_t>(n) * capacity + 1);
    const int rc = gphmm_sw_align(s->h, &b, &prm, capacity, off_out.data(), ne.data(), el.data());
    if (rc != GPHMM_OK && rc != GPHMM_ERR_TOO_LARGE) {
        throw_for(env, s, rc);
        return JNI_FALSE;
    }
    if (n) {
        env->SetIntArrayRegion(offsets, 0, n, reinterpret_cast<const jint *>(off_out.data()));
        env->SetIntArrayRegion(nElems, 0, n, reinterpret_cast<const jint *>(ne.data()));
        env->SetIntArrayRegion(elems, 0, n * capacity, reinterpret_cast<const jint *>(el.data()));
    }
    return rc == GPHMM_OK ? JNI_TRUE : JNI_FALSE;
}

JNIEXPORT jlong JNICALL JNIFN(nativeSubmit)(JNIEnv *env, jclass, jlong handle, jobjectArray reads, jobjectArray haps) {
    Session *s = reinterpret_cast<Session *>(handle);
    gphmm_batch b;
    gphmm_unit unit;
    if (!pack(env, s, reads, haps, &b, &unit)) return 0;
    Pending pend;
    // an empty read or haplotype list is a no-op like in nativeCompute: the ticket completes with nothing to copy
    pend.n = static_cast<size_t>(b.n_reads * b.n_haps);
    pend.out.assign(std::max<size_t>(pend.n, 1), 0.0);
    uint64_t ticket = 0;
    // the vector's heap block does not move when the Pending is moved into the map
    const int rc = gphmm_submit(s->h, &b, pend.out.data(), &ticket);
    if (rc != GPHMM_OK) {
        throw_for(env, s, rc);
        return 0;
    }
    s->pending.emplace(ticket, std::move(pend));
    return static_cast<jlong>(ticket);
}

JNIEXPORT void JNICALL JNIFN(nativeAwait)(JNIEnv *env, jclass, jlong handle, jlong ticket, jdoubleArray out) {
    Session *s = reinterpret_cast<Session *>(handle);
    auto it = s->pending.find(static_cast<uint64_t>(ticket));
    if (it == s->pending.end()) {
        throw_java(env, "java/lang/IllegalArgumentException", "unknown ticket");
        return;
    }
    const int rc = gphmm_wait(s->h, static_cast<uint64_t>(ticket));
    if (rc != GPHMM_OK) {
        s->pending.erase(it);
        throw_for(env, s, rc);
        return;
    }
    const jsize n = static_cast<jsize>(it->second.n);
    if (env->GetArrayLength(out) < n) {
        s->pending.erase(it);
        throw_java(env, "java/lang/IllegalArgumentException", "likelihood array is too small");
        return;
    }
    if (n) env->SetDoubleArrayRegion(out, 0, n, it->second.out.data());
    s->pending.erase(it);
}

JNIEXPORT void JNICALL JNIFN(nativeDestroy)(JNIEnv *, jclass, jlong handle) {
    Session *s = reinterpret_cast<Session *>(handle);
    if (!s) return;
    if (s->h) gphmm_destroy(s->h);
    delete s;
}

JNIEXPORT jlongArray JNICALL JNIFN(nativeCounters)(JNIEnv *env, jclass, jlong handle) {
    Session *s = reinterpret_cast<Session *>(handle);
    gphmm_stats st;
    std::memset(&st, 0, sizeof st);
    if (s && s->h) gphmm_get_stats(s->h, &st);
    const jlong v[6] = {st.pairs, st.cells, st.rescued_pairs, st.h2d_bytes, st.d2h_bytes, st.kernel_launches};
    jlongArray a = env->NewLongArray(6);
    if (a) env->SetLongArrayRegion(a, 0, 6, v);
    return a;
}

JNIEXPORT jdoubleArray JNICALL JNIFN(nativeTimers)(JNIEnv *env, jclass, jlong handle) {
    Session *s = reinterpret_cast<Session *>(handle);
    gphmm_stats st;
    std::memset(&st, 0, sizeof st);
    if (s && s->h) gphmm_get_stats(s->h, &st);
    const jdouble v[5] = {st.fp32_kernel_ms, st.fp64_kernel_ms, st.device_ms, st.host_stage_ms, st.wall_ms};
    jdoubleArray a = env->NewDoubleArray(5);
    if (a) env->SetDoubleArrayRegion(a, 0, 5, v);
    return a;
}

}  // extern "C"
