// host_sw.inl: host side of gphmm_sw_align -- part of gpuphmm.cu (included inside its anonymous namespace; not a translation unit of its own).
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
            if (!tasks.empty() && (bt + need > MAX_BT || aux + (nr + 1) + 4 * (uint64_t)(na + 1) > ((uint64_t)1 << 30) ||
                                   (uint64_t)(tasks.size() + 1) * (uint64_t)capacity > (uint64_t)UINT32_MAX)) break;  // out_off is 32-bit
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
        CK(cudaMemsetAsync(dev.sw_out.p, 0, off_cnt + 16, st));  // cursor; CIGAR slots beyond n_elems read back as 0
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

