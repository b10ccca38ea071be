// host_queue.inl: worker of the asynchronous cross-region queue -- part of gpuphmm.cu (included inside its anonymous namespace; not a translation unit of its own).
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

