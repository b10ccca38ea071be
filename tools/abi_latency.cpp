// abi_latency.cpp -- per-call latency of the C ABI (include/gpuphmm.h) measured from C++, without Python in the loop:
// what a JNI caller sees for one small (region, sample) unit, synchronously and through the asynchronous queue.
//   g++ -O2 -std=c++17 tools/abi_latency.cpp -Iinclude -Lgatk_b200/lib -lgpuphmm -Wl,-rpath,$PWD/gatk_b200/lib -o tools/abi_latency
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "gpuphmm.h"

struct Region {
    std::vector<uint8_t> bases, q, iq, dq, gcp, haps;
    std::vector<int64_t> read_off, hap_off;
    gphmm_unit unit;
    gphmm_batch batch;
    std::vector<double> out;
    double cells = 0;
};

// BASELINE.json config 1: 128 reads x 150 bp, 8 haplotypes of 200-300 bp derived from one random sequence
static void make_region(Region &r, uint32_t seed, int n_reads = 128, int read_len = 150, int n_haps = 8) {
    std::mt19937 g(seed);
    const char L[4] = {'A', 'C', 'G', 'T'};
    const int H0 = 200 + (int)(g() % 101);
    std::vector<uint8_t> h0(H0);
    for (auto &c : h0) c = (uint8_t)L[g() & 3];
    r.hap_off.assign(1, 0);
    for (int h = 0; h < n_haps; ++h) {
        std::vector<uint8_t> x = h0;
        for (int k = 0; k < (h ? 1 + (int)(g() % 3) : 0); ++k) x[g() % x.size()] = (uint8_t)L[g() & 3];
        r.haps.insert(r.haps.end(), x.begin(), x.end());
        r.hap_off.push_back((int64_t)r.haps.size());
    }
    r.read_off.assign(1, 0);
    std::normal_distribution<double> qd(32.0, 6.0);
    for (int k = 0; k < n_reads; ++k) {
        const int off = (int)(g() % (H0 - read_len + 1));
        for (int i = 0; i < read_len; ++i) {
            int q = (int)qd(g);
            q = q < 6 ? 6 : (q > 41 ? 41 : q);
            if (q < 18) q = 6;
            r.bases.push_back(h0[off + i]);
            r.q.push_back((uint8_t)q);
            r.iq.push_back(45); r.dq.push_back(45); r.gcp.push_back(10);
        }
        r.read_off.push_back((int64_t)r.bases.size());
    }
    r.unit = {0, n_reads, 0, n_haps, 0};
    std::memset(&r.batch, 0, sizeof r.batch);
    r.batch.read_bases = r.bases.data(); r.batch.base_q = r.q.data(); r.batch.ins_q = r.iq.data();
    r.batch.del_q = r.dq.data(); r.batch.gcp = r.gcp.data(); r.batch.read_off = r.read_off.data(); r.batch.n_reads = n_reads;
    r.batch.hap_bases = r.haps.data(); r.batch.hap_off = r.hap_off.data(); r.batch.n_haps = n_haps;
    r.batch.units = &r.unit; r.batch.n_units = 1;
    r.out.assign((size_t)n_reads * n_haps, 0.0);
    for (int h = 0; h < n_haps; ++h) r.cells += (double)n_reads * read_len * (double)(r.hap_off[h + 1] - r.hap_off[h]);
}

static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    const int in_flight = argc > 1 ? atoi(argv[1]) : 64;
    gphmm_config cfg;
    std::memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = (int32_t)sizeof cfg;
    gphmm_t *h = nullptr;
    int rc = gphmm_create(&cfg, &h);
    if (rc != GPHMM_OK) { fprintf(stderr, "gphmm_create: %s\n", gphmm_strerror(rc)); return 1; }
    std::vector<Region> regions(in_flight);
    for (int k = 0; k < in_flight; ++k) make_region(regions[k], 47382911u + k);

    for (int k = 0; k < 50; ++k) gphmm_compute(h, &regions[0].batch, regions[0].out.data());
    gphmm_reset_stats(h);
    const int n = 500;
    double t0 = now_us();
    for (int k = 0; k < n; ++k)
        if ((rc = gphmm_compute(h, &regions[k % in_flight].batch, regions[k % in_flight].out.data())) != GPHMM_OK) { fprintf(stderr, "compute: %s\n", gphmm_last_error(h)); return 1; }
    double dt = (now_us() - t0) / n;
    gphmm_stats st;
    gphmm_get_stats(h, &st);
    printf("sync gphmm_compute, 1 region (128x150 reads, 8 haps): %.1f us/call = %.0f GCUPS; device %.1f us, host staging %.1f us, %.1f launches/call\n",
           dt, regions[0].cells / dt / 1e3, st.device_ms * 1e3 / n, st.host_stage_ms * 1e3 / n, (double)st.kernel_launches / n);

    std::vector<uint64_t> tickets(in_flight);
    double best = 1e30, best_sub = 0;
    for (int rep = 0; rep < 10; ++rep) {
        t0 = now_us();
        for (int k = 0; k < in_flight; ++k) gphmm_submit(h, &regions[k].batch, regions[k].out.data(), &tickets[k]);
        const double t_sub = now_us() - t0;
        for (int k = 0; k < in_flight; ++k)
            if ((rc = gphmm_wait(h, tickets[k])) != GPHMM_OK) { fprintf(stderr, "wait: %s\n", gphmm_last_error(h)); return 1; }
        dt = now_us() - t0;
        if (rep >= 3 && dt < best) { best = dt; best_sub = t_sub; }
        if (rep == 9) printf("async queue, %d regions in flight: %.1f us/region = %.0f GCUPS (submit loop %.1f us/region)\n", in_flight,
                             best / in_flight, regions[0].cells * in_flight / best / 1e3, best_sub / in_flight);
    }
    gphmm_destroy(h);
    return 0;
}
