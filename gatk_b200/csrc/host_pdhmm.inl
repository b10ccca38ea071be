// host_pdhmm.inl: host side of gphmm_pd_compute -- part of gpuphmm.cu (included inside its anonymous namespace; not a translation unit of its own).
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

template <int K, bool SIMPLE> KernelInfo pd_fast_kernel_info() {
    KernelInfo ki;
    auto fn = phmm_pd_fast_kernel<K, SIMPLE>;
    ki.fn = (const void *)fn;
    ki.smem = pd_fast_smem<K>(PD_MAX_CODES, SIMPLE);  // sized per launch from the chunk's alphabet; this is the ceiling
    raise_dyn_smem((const void *)fn, ki.smem);
    return ki;
}

// Deletion events of a haplotype for the SIMPLE kernels (pdhmm_kernels.cuh): (a, b) = first and last column of a deletion.
// Events may touch: a flag met in the AFTER_DEL state starts the next event on that very column.  False when a deletion is
// still open (or just closed) at the last column, which LoglessPDPairHMM.java:59 carries into the next row.
bool pd_simple_events(const uint8_t *pd, uint32_t H, std::vector<uint2> &events) {
    enum { DEL_START = 2, DEL_END = 4 };
    const size_t n0 = events.size();
    uint32_t state = PD_NORMAL, a = 0;
    for (uint32_t j = 1; j <= H; ++j) {
        const bool ds = pd[j - 1] & DEL_START, de = pd[j - 1] & DEL_END;
        if (state != PD_INSIDE_DEL) {  // NORMAL or AFTER_DEL
            if (ds || de) {
                a = j;
                if (de) { events.push_back(make_uint2(a, j)); state = PD_AFTER_DEL; } else state = PD_INSIDE_DEL;
            } else {
                state = PD_NORMAL;
            }
        } else if (de) {  // INSIDE_DEL: a further DEL_START changes nothing
            events.push_back(make_uint2(a, j));
            state = PD_AFTER_DEL;
        }
    }
    if (state != PD_NORMAL) { events.resize(n0); return false; }
    return true;
}

// Schedule of a SIMPLE haplotype: lane l is on column j at step j + l, and only the columns a, b, b + 1 of an event need the
// window code.
void plan_pd_steps_simple(const uint2 *events, size_t n_events, uint32_t H, std::vector<uint2> &segs, std::vector<uint8_t> &slow) {
    const uint32_t T = H + 33;
    slow.assign(T + 2, 0);
    for (size_t e = 0; e < n_events; ++e)
        for (uint32_t j : {events[e].x, events[e].y, events[e].y + 1})
            for (uint32_t st = j; st <= std::min(T, j + 31); ++st) slow[st] = 1;
    for (uint32_t st = 1; st <= T;) {
        uint2 seg = make_uint2(0, 0);
        while (st <= T && !slow[st]) { ++seg.x; ++st; }
        while (st <= T && slow[st]) { ++seg.y; ++st; }
        segs.push_back(seg);
    }
}

// The steps of a haplotype's sweep (H + 33 of them: every lane gets past column H + 1) in which some lane of the warp is
// on, or one column before, a column of the deletion state machine run the slow step; all others the fast one.
// Appends (fast, slow) pairs to `segs`.
void plan_pd_steps(const uint8_t *flags, uint32_t H, uint32_t first_event, uint32_t carry, std::vector<uint2> &segs,
                   std::vector<uint8_t> &slow) {
    const uint32_t T = H + 33;
    slow.assign(T + 2, 0);
    for (uint32_t j = 1; j <= H; ++j) {  // 1-based column
        const uint8_t f = flags[j - 1];
        const bool special = (f & PD_DEL_END_BIT) || ((f >> PD_TYPE_SHIFT) & PD_TYPE_BITS) || (carry == PD_INSIDE_DEL && j <= first_event) ||
                             (carry == PD_AFTER_DEL && j == 1);
        if (!special) continue;
        // lane l meets column j at step j + l and needs its branch values refreshed one step earlier
        for (uint32_t st = j > 1 ? j - 1 : 1; st <= std::min(T, j + 31); ++st) slow[st] = 1;
    }
    for (uint32_t st = 1; st <= T;) {
        uint2 seg = make_uint2(0, 0);
        while (st <= T && !slow[st]) { ++seg.x; ++st; }
        while (st <= T && slow[st]) { ++seg.y; ++st; }
        segs.push_back(seg);
    }
}

// Host-side plan of one PD-HMM chunk (everything that needs no device): built ahead of the GPU by helper threads.
struct PdChunkPlan {
    std::pair<int64_t, int64_t> ch;
    int64_t r_lo = 0, r_hi = 0, base_lo = 0, cells = 0;
    size_t span = 0, stride = 0;
    uint32_t n_pairs = 0, max_h = 1;
    int n_plain = 0, max_snp = 0;
    bool fast_ok = false;
    uint8_t code_byte[PD_MAX_CODES];
    uint32_t first[10], group_first[7];
    std::vector<PdGroup> groups;          // fast kernels: one task per read = its consecutive pairs of a bucket
    std::vector<uint32_t> read_off, unit_out_base;
    std::vector<uint8_t> hap_bytes, hap_flags, code_stream, flag_stream, slow_scratch;
    std::vector<PdTask> tasks[9], all;   // 0..2: first-version kernels by read length, 3..5: fast kernels, 6..8: fast kernels, SIMPLE form
    std::vector<PdHap> haps;
    std::vector<uint2> segs, events;

    void build(const gphmm_batch *b, const uint8_t *hap_pd, bool allow_fast, bool allow_simple) {
        r_lo = INT64_MAX; r_hi = 0;
        for (int64_t u = ch.first; u < ch.second; ++u) {
            const gphmm_unit &un = b->units[u];
            if (un.read_end > un.read_begin) { r_lo = std::min(r_lo, un.read_begin); r_hi = std::max(r_hi, un.read_end); }
        }
        if (r_hi <= r_lo) r_lo = r_hi = 0;
        base_lo = b->n_reads ? b->read_off[r_lo] : 0;
        const int64_t base_hi = b->n_reads ? b->read_off[r_hi] : 0;
        span = (size_t)(base_hi - base_lo); stride = align_up(span, 16);
        read_off.resize((size_t)(r_hi - r_lo) + 1);
        for (int64_t r = 0; r <= r_hi - r_lo; ++r) read_off[r] = (uint32_t)(b->read_off[r_lo + r] - base_lo);
        hap_bytes.clear(); hap_flags.clear(); unit_out_base.clear();
        for (auto &v : tasks) v.clear();
        code_stream.clear(); flag_stream.clear(); haps.clear(); segs.clear(); events.clear();
        // prior-table rows of the chunk: 0 = outside a haplotype, 1 .. n_plain = one per haplotype byte on a column without a SNP
        // flag, two source rows, then the SNP rows of the haplotype being swept (rebuilt per haplotype: at most PD_MAX_SNP_CODES
        // distinct (byte, mask) pairs).  SNP columns enter the stream as 0xc0 + s and are renumbered once n_plain is known.
        memset(code_byte, 0, sizeof code_byte);
        n_plain = 0; max_snp = 0;
        fast_ok = allow_fast;
        uint8_t plain_of[256];
        memset(plain_of, 0, sizeof plain_of);
        n_pairs = 0; max_h = 1; cells = 0;
        for (int64_t u = ch.first; u < ch.second; ++u) {
            const gphmm_unit &un = b->units[u];
            const uint32_t nr = (uint32_t)(un.read_end - un.read_begin), nh = (uint32_t)(un.hap_end - un.hap_begin);
            unit_out_base.push_back(n_pairs);
            struct HapInfo { uint32_t off, H, first_event, carry, index; bool sparse, simple; };
            std::vector<HapInfo> hi(nh);
            for (uint32_t k = 0; k < nh; ++k) {
                const int64_t ho = b->hap_off[un.hap_begin + k];
                const uint32_t H = (uint32_t)(b->hap_off[un.hap_begin + k + 1] - ho);
                hi[k].off = (uint32_t)hap_bytes.size(); hi[k].H = H;
                hap_bytes.insert(hap_bytes.end(), b->hap_bases + ho, b->hap_bases + ho + H);
                hap_flags.resize(hap_bytes.size());
                encode_pd_columns(hap_pd + ho, H, hap_flags.data() + hi[k].off, hi[k].first_event, hi[k].carry);
                max_h = std::max(max_h, H);
                // fast kernels: padded code and flag streams, the step schedule, the SNP codes of the haplotype
                hi[k].index = (uint32_t)haps.size();
                PdHap ph;
                memset(&ph, 0, sizeof ph);
                code_stream.insert(code_stream.end(), STREAM_PAD, 0);
                flag_stream.insert(flag_stream.end(), STREAM_PAD, 0);
                ph.code_off = (uint32_t)code_stream.size();
                bool hap_ok = fast_ok;
                for (uint32_t j = 0; j < H && hap_ok; ++j) {
                    const uint8_t f = hap_flags[hi[k].off + j], hb = b->hap_bases[ho + j];
                    if (f & PD_SNP_BIT) {
                        const uint16_t key = (uint16_t)(hb | (uint16_t)(f & PD_MASK_BITS) << 8);
                        uint32_t sn = 0;
                        while (sn < ph.n_snp && ph.snp[sn] != key) ++sn;
                        if (sn == ph.n_snp) {
                            if (ph.n_snp == PD_MAX_SNP_CODES) { hap_ok = false; break; }  // the first-version kernels take this haplotype
                            ph.snp[ph.n_snp++] = key;
                        }
                        code_stream.push_back((uint8_t)(0xc0 + sn));
                    } else {
                        if (!plain_of[hb]) {
                            if (n_plain + 3 + PD_MAX_SNP_CODES >= PD_MAX_CODES) { fast_ok = hap_ok = false; break; }  // exotic alphabet
                            plain_of[hb] = (uint8_t)++n_plain;
                            code_byte[n_plain] = hb;
                        }
                        code_stream.push_back(plain_of[hb]);
                    }
                }
                max_snp = std::max(max_snp, (int)ph.n_snp);
                code_stream.resize(ph.code_off + H, 0);
                flag_stream.insert(flag_stream.end(), hap_flags.begin() + hi[k].off, hap_flags.begin() + hi[k].off + H);
                code_stream.insert(code_stream.end(), 2 * STREAM_PAD, 0);
                flag_stream.insert(flag_stream.end(), 2 * STREAM_PAD, 0);
                ph.seg_first = (uint32_t)segs.size();
                ph.ev_first = (uint32_t)events.size();
                hi[k].simple = allow_simple && pd_simple_events(hap_pd + ho, H, events);
                ph.n_events = (uint32_t)events.size() - ph.ev_first;
                if (hi[k].simple) plan_pd_steps_simple(events.data() + ph.ev_first, ph.n_events, H, segs, slow_scratch);
                else plan_pd_steps(hap_flags.data() + hi[k].off, H, hi[k].first_event, hi[k].carry, segs, slow_scratch);
                ph.n_segs = (uint32_t)segs.size() - ph.seg_first;
                // densely flagged haplotypes (most steps inside a deletion window) stay with the first-version kernels,
                // whose per-step fast path works column by column
                uint32_t n_slow = 0;
                for (uint32_t q = ph.seg_first; q < (uint32_t)segs.size(); ++q) n_slow += segs[q].y;
                // (the SIMPLE window form costs ~3 plain steps per window step: still ahead of the first-version kernels when every step is one)
                hi[k].sparse = hap_ok && (hi[k].simple || 2 * n_slow <= H + 33);
                haps.push_back(ph);
            }
            for (uint32_t r = 0; r < nr; ++r) {
                const uint32_t rl = (uint32_t)(un.read_begin - r_lo) + r, R = read_off[rl + 1] - read_off[rl];
                // (the choice between the kernel families is made once the whole chunk is known: see fast_ok below)
                const int bucket = R <= 63 ? 0 : (R <= 127 ? 1 : 2);
                const int fbucket = R <= 94 ? 3 : (R <= 158 ? 4 : (R <= 254 ? 5 : -1));
                for (uint32_t k = 0; k < nh; ++k) {
                    PdTask t;
                    t.read = rl; t.hap_off = hi[k].off; t.H = hi[k].H; t.out_slot = n_pairs + r * nh + k;
                    t.first_event = hi[k].first_event; t.carry = hi[k].carry;
                    t.c0_exp = 125 - ceil_log2(hi[k].H); t.pad = hi[k].index;
                    if (fbucket >= 0 && hi[k].sparse) {
                        // scaled states (I / tMI, D / tMD) need the head-room of the plain fp32 kernels
                        PdTask tf = t;
                        tf.c0_exp = C0_BASE_EXP_F32 - ceil_log2(hi[k].H);
                        tasks[fbucket + (hi[k].simple ? 3 : 0)].push_back(tf);
                        cells += (int64_t)R * hi[k].H;
                        continue;
                    }
                    tasks[bucket].push_back(t);
                    cells += (int64_t)R * hi[k].H;
                }
            }
            n_pairs += nr * nh;
        }
        if (n_pairs == 0) return;
        if (!fast_ok)  // every read on the first-version kernels (initial value 2^(125 - ceil log2 H) again)
            for (int k = 3; k < 9; ++k) {
                for (PdTask t : tasks[k]) {
                    const uint32_t R = read_off[t.read + 1] - read_off[t.read];
                    t.c0_exp = 125 - ceil_log2(t.H);
                    tasks[R <= 63 ? 0 : (R <= 127 ? 1 : 2)].push_back(t);
                }
                tasks[k].clear();
            }
        all.clear();
        memset(first, 0, sizeof first);
        for (int k = 0; k < 9; ++k) { first[k + 1] = first[k] + (uint32_t)tasks[k].size(); all.insert(all.end(), tasks[k].begin(), tasks[k].end()); }
        groups.clear();
        memset(group_first, 0, sizeof group_first);
        for (int k = 3; k < 9; ++k) {
            const std::vector<PdTask> &tk = tasks[k];
            for (size_t i = 0; i < tk.size();) {
                size_t j = i + 1;
                while (j < tk.size() && tk[j].read == tk[i].read) ++j;
                PdGroup gr;
                gr.read = tk[i].read; gr.task_first = (uint32_t)i; gr.n = (uint32_t)(j - i); gr.pad = 0;
                groups.push_back(gr);
                i = j;
            }
            group_first[k - 3 + 1] = (uint32_t)groups.size();
        }
        for (uint8_t &c : code_stream)
            if (c >= 0xc0) c = (uint8_t)(n_plain + 3 + (c - 0xc0));
    }
};

int run_pd_batch(gphmm *h, const gphmm_batch *b, const uint8_t *hap_pd, double *out) {
    std::lock_guard<std::mutex> run_lk(h->run_mu);
    const double t0 = now_ms();
    validate_batch(b);
    if (b->n_units == 0) return GPHMM_OK;
    if (!out) throw Error(GPHMM_ERR_INVALID_ARG, "out is null");
    if (!hap_pd) throw Error(GPHMM_ERR_INVALID_ARG, "hap_pd_bases is null");
    Device &dev = *h->devices[0];
    CK(cudaSetDevice(dev.ordinal));
    // reads of 128+ bases: 4 rows per lane in strips of 128 rows keeps 18 warps per SM resident (113 registers) where the
    // 8-row variant (181 registers) keeps 11; GPHMM_PD_K8=1 selects the latter for comparison
    static const bool k8 = getenv("GPHMM_PD_K8") != nullptr;
    static const KernelInfo kf[3] = {pd_kernel_info<float, 2>(), pd_kernel_info<float, 4>(), k8 ? pd_kernel_info<float, 8>() : pd_kernel_info<float, 4>()};
    static const KernelInfo kd = pd_kernel_info<double, 4>();
    // fast fp32 kernels for reads of up to 94 / 158 / 254 bases (3 / 5 / 8 rows per lane); GPHMM_PD_SLOW=1 keeps every read on
    // the first-version kernels (A/B switch)
    // SIMPLE form (kfast[3..5]): haplotypes with well-formed, separate deletions; GPHMM_PD_NO_SIMPLE=1 keeps them on the first form
    static const KernelInfo kfast[6] = {pd_fast_kernel_info<3, false>(), pd_fast_kernel_info<5, false>(), pd_fast_kernel_info<8, false>(),
                                        pd_fast_kernel_info<3, true>(), pd_fast_kernel_info<5, true>(), pd_fast_kernel_info<8, true>()};
    static const bool no_fast = getenv("GPHMM_PD_SLOW") != nullptr;
    static const bool allow_simple = getenv("GPHMM_PD_NO_SIMPLE") == nullptr;
    // two chunks in flight (slot k on streams[k]): the upload and the kernels of one overlap the download and the host-side
    // scatter of the other; GPHMM_PD_SERIAL=1 runs them one after the other (device_ms is then the kernels' time alone)
    static const bool serial = getenv("GPHMM_PD_SERIAL") != nullptr;
    const int n_slots = serial ? 1 : 2;
    // the first chunks are small (doubling from 2.5e8 cells): the GPU starts while the later plans are still being built
    const auto chunks = split_units(b, h->chunk_cells() / 4, h->chunk_bytes(), !serial);
    int64_t launches = 0, total_pairs = 0, total_cells = 0, total_redo = 0, h2d = 0, d2h = 0;
    double device_ms = 0;
    // chunk plans are built by helper threads a few chunks ahead of the GPU (the per-pair task list is the expensive part)
    const bool allow_fast = !no_fast && !h->cfg.force_fp64;
    const size_t lookahead = std::max<size_t>(2, std::min<size_t>(4, (size_t)(h->cfg.host_threads > 0 ? h->cfg.host_threads : 4)));
    std::deque<std::future<std::unique_ptr<PdChunkPlan>>> ahead;
    size_t next_chunk = 0;
    auto start_plan = [&]() {
        const auto ch = chunks[next_chunk++];
        ahead.push_back(std::async(std::launch::async, [b, hap_pd, allow_fast, ch]() {
            std::unique_ptr<PdChunkPlan> P(new PdChunkPlan());
            P->ch = ch;
            P->build(b, hap_pd, allow_fast, allow_simple);
            return P;
        }));
    };
    struct InFlight { std::unique_ptr<PdChunkPlan> P; size_t off_err = 0, off_cnt = 0; };
    InFlight fl[2];
    bool any_launched = false;
    int last_slot = 0;

    // everything of one chunk onto the slot's stream: upload, kernels, epilogues, download
    auto launch = [&](int sl, std::unique_ptr<PdChunkPlan> Pp) {
        PdChunkPlan &P = *Pp;
        cudaStream_t st = dev.streams[sl];
        const uint32_t n_pairs = P.n_pairs, max_h = P.max_h;
        const int64_t base_lo = P.base_lo;
        const size_t span = P.span, stride = P.stride;
        const int n_rows = P.n_plain + 3 + P.max_snp;
        const uint32_t *first = P.first;
        auto &read_off = P.read_off; auto &hap_bytes = P.hap_bytes; auto &hap_flags = P.hap_flags; auto &all = P.all;
        auto &code_stream = P.code_stream; auto &flag_stream = P.flag_stream; auto &haps = P.haps; auto &segs = P.segs; auto &events = P.events;
        auto &groups = P.groups;
        // device image
        size_t o = 0;
        const size_t off_ro = o; o = align_up(o + read_off.size() * 4, 16);
        const size_t off_hb = o; o = align_up(o + hap_bytes.size(), 16);
        const size_t off_hf = o; o = align_up(o + hap_flags.size(), 16);
        const size_t off_tk = o; o = align_up(o + all.size() * sizeof(PdTask), 16);
        const size_t off_cs = o; o = align_up(o + code_stream.size(), 16);
        const size_t off_fs = o; o = align_up(o + flag_stream.size(), 16);
        const size_t off_hp = o; o = align_up(o + haps.size() * sizeof(PdHap), 16);
        const size_t off_sg = o; o = align_up(o + segs.size() * sizeof(uint2), 16);
        const size_t off_ev = o; o = align_up(o + events.size() * sizeof(uint2), 16);
        const size_t off_gr = o; o = align_up(o + groups.size() * sizeof(PdGroup), 16);
        const size_t meta_bytes = o;
        dev.pd_meta[sl].reserve(meta_bytes); dev.pd_h_meta[sl].reserve(meta_bytes);
        uint8_t *hm = (uint8_t *)dev.pd_h_meta[sl].p;
        memcpy(hm + off_ro, read_off.data(), read_off.size() * 4);
        memcpy(hm + off_hb, hap_bytes.data(), hap_bytes.size());
        memcpy(hm + off_hf, hap_flags.data(), hap_flags.size());
        memcpy(hm + off_tk, all.data(), all.size() * sizeof(PdTask));
        memcpy(hm + off_cs, code_stream.data(), code_stream.size());
        memcpy(hm + off_fs, flag_stream.data(), flag_stream.size());
        memcpy(hm + off_hp, haps.data(), haps.size() * sizeof(PdHap));
        memcpy(hm + off_sg, segs.data(), segs.size() * sizeof(uint2));
        memcpy(hm + off_ev, events.data(), events.size() * sizeof(uint2));
        memcpy(hm + off_gr, groups.data(), groups.size() * sizeof(PdGroup));
        o = 0;
        const size_t off_out = o; o = align_up(o + (size_t)n_pairs * 8, 16);
        const size_t off_cnt = o; o = align_up(o + 16 * 4, 16);
        const size_t off_err = o; o = align_up(o + 16, 16);
        const size_t dl_bytes = o;
        const size_t off_s32 = o; o = align_up(o + (size_t)n_pairs * 4, 16);
        const size_t off_redo = o; o = align_up(o + (size_t)n_pairs * 4, 16);
        const size_t off_s64 = o; o = align_up(o + (size_t)n_pairs * 8, 16);
        dev.pd_work[sl].reserve(o); dev.pd_h_out[sl].reserve(dl_bytes);
        dev.pd_reads[sl].reserve(std::max<size_t>(stride * 5, 16));
        const uint32_t max_grid = (uint32_t)dev.n_sms * 32;
        dev.pd_bnd[sl].reserve((size_t)max_grid * (max_h + 1) * sizeof(BndPD<double>));
        const uint8_t *src[5] = {b->read_bases, b->base_q, b->ins_q, b->del_q, b->gcp};
        for (int a = 0; a < 5 && span; ++a)
            CK(cudaMemcpyAsync((uint8_t *)dev.pd_reads[sl].p + a * stride, src[a] + base_lo, span, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(dev.pd_meta[sl].p, dev.pd_h_meta[sl].p, meta_bytes, cudaMemcpyHostToDevice, st));
        h2d += (int64_t)(span * 5 + meta_bytes);
        uint8_t *meta = (uint8_t *)dev.pd_meta[sl].p, *work = (uint8_t *)dev.pd_work[sl].p;
        uint32_t *counters = (uint32_t *)(work + off_cnt);
        CK(cudaMemsetAsync(work + off_out + (size_t)n_pairs * 8, 0, (off_err + 16) - (off_out + (size_t)n_pairs * 8), st));  // alignment gap, counters, error flag
        CK(cudaEventRecord(dev.pd_ev0[sl], st));
        if (!any_launched) { CK(cudaEventRecord(dev.ev_step0, st)); any_launched = true; }
        last_slot = sl;
        PdArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.rd_bases = (const uint8_t *)dev.pd_reads[sl].p;
        pa.rd_q = pa.rd_bases + stride; pa.rd_i = pa.rd_q + stride; pa.rd_d = pa.rd_i + stride; pa.rd_c = pa.rd_d + stride;
        pa.read_off = (const uint32_t *)(meta + off_ro);
        pa.hap_bases = meta + off_hb; pa.hap_flags = meta + off_hf;
        pa.tasks = (const PdTask *)(meta + off_tk);
        pa.bnd = dev.pd_bnd[sl].p; pa.bnd_stride = max_h + 1;
        pa.m2m = (const double *)dev.m2m.p;
        pa.err = (int *)(work + off_err);
        pa.tristate_off = h->cfg.tristate_off != 0;
        if (!h->cfg.force_fp64) {
            for (int k = 3; k < 9; ++k) {
                const uint32_t n = first[k + 1] - first[k];
                if (!n) continue;
                PdFastArgs fa;
                memset(&fa, 0, sizeof fa);
                fa.rd_bases = pa.rd_bases; fa.rd_q = pa.rd_q; fa.rd_i = pa.rd_i; fa.rd_d = pa.rd_d; fa.rd_c = pa.rd_c;
                fa.read_off = pa.read_off;
                fa.codes = meta + off_cs; fa.flags = meta + off_fs;
                fa.tasks = pa.tasks; fa.haps = (const PdHap *)(meta + off_hp); fa.segs = (const uint2 *)(meta + off_sg);
                fa.events = (const uint2 *)(meta + off_ev);
                fa.groups = (const PdGroup *)(meta + off_gr) + P.group_first[k - 3]; fa.n_groups = P.group_first[k - 3 + 1] - P.group_first[k - 3];
                fa.first = first[k]; fa.n_tasks = n; fa.counter = counters + 10 + (k - 3);
                fa.sums = (float *)(work + off_s32);
                fa.m2m = pa.m2m; fa.err = pa.err; fa.tristate_off = pa.tristate_off; fa.n_plain = P.n_plain; fa.n_rows = n_rows;
                memcpy(fa.code_byte, P.code_byte, sizeof fa.code_byte);
                const int rows = (k - 3) % 3 == 0 ? 3 : ((k - 3) % 3 == 1 ? 5 : 8);
                const bool simple = k >= 6;
                const size_t smem = rows == 3 ? pd_fast_smem<3>(n_rows, simple) : (rows == 5 ? pd_fast_smem<5>(n_rows, simple) : pd_fast_smem<8>(n_rows, simple));
                int occ = 0;
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfast[k - 3].fn, 32, smem));
                const uint32_t grid = std::min<uint32_t>(fa.n_groups, (uint32_t)(dev.n_sms * std::max(1, occ)));
                void *args[] = {&fa};
                CK(cudaLaunchKernel(kfast[k - 3].fn, dim3(grid), dim3(32), args, smem, st));
                ++launches;
            }
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
        CK(cudaEventRecord(dev.pd_ev1[sl], st));
        CK(cudaMemcpyAsync(dev.pd_h_out[sl].p, work + off_out, dl_bytes, cudaMemcpyDeviceToHost, st));
        d2h += (int64_t)dl_bytes;
        fl[sl].P = std::move(Pp); fl[sl].off_err = off_err; fl[sl].off_cnt = off_cnt;
    };
    // wait for the slot's chunk, check its error flag, scatter its results
    auto finish = [&](int sl) {
        if (!fl[sl].P) return;
        std::unique_ptr<PdChunkPlan> Pp = std::move(fl[sl].P);
        PdChunkPlan &P = *Pp;
        CK(cudaStreamSynchronize(dev.streams[sl]));
        if (serial) {
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, dev.pd_ev0[sl], dev.pd_ev1[sl]));
            device_ms += ms;
        }
        const uint8_t *ho = (const uint8_t *)dev.pd_h_out[sl].p;
        const int err = *(const int *)(ho + fl[sl].off_err);
        if (err == 1) throw Error(GPHMM_ERR_BAD_QUAL, "quality score out of range: ins/del/gcp > 127 or base qual 255");
        if (err == 2) throw Error(GPHMM_ERR_INVALID_ARG, "read base other than ACGT on a SNP column of a partially determined haplotype (LoglessPDPairHMM.java:202)");
        total_redo += ((const uint32_t *)(ho + fl[sl].off_cnt))[8];
        const double *res = (const double *)ho;
        for (int64_t u = P.ch.first; u < P.ch.second; ++u) {
            const gphmm_unit &un = b->units[u];
            const size_t n = (size_t)(un.read_end - un.read_begin) * (size_t)(un.hap_end - un.hap_begin);
            if (n) memcpy(out + un.out_off, res + P.unit_out_base[u - P.ch.first], n * sizeof(double));
        }
        total_pairs += P.n_pairs; total_cells += P.cells;
    };
    try {
        while (next_chunk < chunks.size() && ahead.size() < lookahead) start_plan();
        int sl = 0;
        while (!ahead.empty()) {
            std::unique_ptr<PdChunkPlan> Pp = ahead.front().get();
            ahead.pop_front();
            if (next_chunk < chunks.size()) start_plan();
            if (Pp->n_pairs == 0) continue;
            finish(sl);  // the chunk that used this slot two chunks ago
            launch(sl, std::move(Pp));
            sl = (sl + 1) % n_slots;
        }
        for (int k = 0; k < n_slots; ++k) { finish(sl); sl = (sl + 1) % n_slots; }
        if (!serial && any_launched) {  // overlapping chunks: the span from the first kernel to the last epilogue
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, dev.ev_step0, dev.pd_ev1[last_slot]));
            device_ms = ms;
        }
    } catch (...) {
        // nothing of this call may still be running (or planning) when the error reaches the caller
        for (auto &f : ahead) if (f.valid()) f.wait();
        for (int k = 0; k < 2; ++k) { cudaStreamSynchronize(dev.streams[k]); fl[k].P.reset(); }
        throw;
    }
    std::lock_guard<std::mutex> lk(h->stats.mu);
    h->stats.s.pairs += total_pairs; h->stats.s.cells += total_cells; h->stats.s.rescued_pairs += total_redo;
    h->stats.s.kernel_launches += launches; h->stats.s.h2d_bytes += h2d; h->stats.s.d2h_bytes += d2h;
    h->stats.s.device_ms += device_ms; h->stats.s.wall_ms += now_ms() - t0;
    return GPHMM_OK;
}

