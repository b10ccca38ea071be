// host_planner.inl: batch validation, chunking, haplotype-prefix sharing plan, chunk plan -- part of gpuphmm.cu (included inside its anonymous namespace; not a translation unit of its own).
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
    // 0 = fine; the scans are linear in reads + haplotypes + units, which is milliseconds of serial time in front of a
    // multi-GPU batch (millions of reads): large batches are checked by a few threads
    auto scan = [b](int64_t r0, int64_t r1, int64_t h0, int64_t h1, int64_t u0, int64_t u1) -> int {
        for (int64_t r = r0; r < r1; ++r)
            if (b->read_off[r + 1] < b->read_off[r]) return 1;
        for (int64_t h = h0; h < h1; ++h)
            if (b->hap_off[h + 1] <= b->hap_off[h]) return 2;
        for (int64_t u = u0; u < u1; ++u) {
            const gphmm_unit &un = b->units[u];
            if (un.read_begin < 0 || un.read_end < un.read_begin || un.read_end > b->n_reads || un.hap_begin < 0 ||
                un.hap_end < un.hap_begin || un.hap_end > b->n_haps || un.out_off < 0)
                return 3;
            if (un.hap_end - un.hap_begin > 65535) return 4;
        }
        return 0;
    };
    int worst = 0;
    const int n_threads = b->n_reads + b->n_haps + b->n_units < 400000 ? 1 : (int)std::min<int64_t>(8, std::max(1u, std::thread::hardware_concurrency()));
    if (n_threads == 1) {
        worst = scan(0, b->n_reads, 0, b->n_haps, 0, b->n_units);
    } else {
        std::vector<int> rc((size_t)n_threads, 0);
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; ++t)
            th.emplace_back([&, t] {
                rc[(size_t)t] = scan(b->n_reads * t / n_threads, b->n_reads * (t + 1) / n_threads, b->n_haps * t / n_threads,
                                     b->n_haps * (t + 1) / n_threads, b->n_units * t / n_threads, b->n_units * (t + 1) / n_threads);
            });
        for (auto &t : th) t.join();
        for (int v : rc) worst = worst ? worst : v;
    }
    switch (worst) {
        case 1: throw Error(GPHMM_ERR_INVALID_ARG, "read_off not monotone");
        case 2: throw Error(GPHMM_ERR_INVALID_ARG, "zero-length haplotype (PairHMM.initialize requires haplotypeMaxLength > 0)");
        case 3: throw Error(GPHMM_ERR_INVALID_ARG, "unit range out of bounds");
        case 4: throw Error(GPHMM_ERR_TOO_LARGE, "more than 65535 haplotypes in one unit");
        default: break;
    }
    if (b->n_reads > 0 && b->read_off[b->n_reads] > 0 &&
        (!b->read_bases || !b->base_q || !b->ins_q || !b->del_q || !b->gcp))
        throw Error(GPHMM_ERR_INVALID_ARG, "null read array");
    if (b->n_haps > 0 && !b->hap_bases) throw Error(GPHMM_ERR_INVALID_ARG, "null hap_bases");
}

// Shape of a read's gap qualities in one pass over the three arrays (8 bytes at a time): bit 0 = insertion quality
// constant, bit 1 = deletion quality constant, bit 2 = gap-continuation penalty constant, bit 3 = ins == del on every base.
inline unsigned qual_shape(const uint8_t *qi, const uint8_t *qd, const uint8_t *qc, size_t n) {
    const uint64_t ones = 0x0101010101010101ULL;
    const uint64_t bi = ones * qi[0], bd = ones * qd[0], bc = ones * qc[0];
    uint64_t ai = 0, ad = 0, ac = 0, as = 0;
    size_t k = 0;
    for (; k + 8 <= n; k += 8) {
        uint64_t wi, wd, wc;
        memcpy(&wi, qi + k, 8); memcpy(&wd, qd + k, 8); memcpy(&wc, qc + k, 8);
        ai |= wi ^ bi; ad |= wd ^ bd; ac |= wc ^ bc; as |= wi ^ wd;
    }
    for (; k < n; ++k) {
        ai |= (uint64_t)(qi[k] ^ qi[0]); ad |= (uint64_t)(qd[k] ^ qd[0]); ac |= (uint64_t)(qc[k] ^ qc[0]); as |= (uint64_t)(qi[k] ^ qd[k]);
    }
    return (ai == 0 ? 1u : 0u) | (ad == 0 ? 2u : 0u) | (ac == 0 ? 4u : 0u) | (as == 0 ? 8u : 0u);
}

// Greedy split of the unit list into chunks bounded by cells and staged bytes.  With `ramp_up` (staged batches) the first
// chunks are small, so that the GPU starts while the host is still planning; with `taper_devices` > 0 the last chunks shrink
// geometrically (remaining cells / (2 x devices), not below 2.5e9), so that the devices draining one chunk list finish
// together and little is left after the last kernel -- the middle of the batch keeps full-size chunks, whose fixed costs
// (drain of the persistent grids, closing kernels, download) weigh least.
std::vector<std::pair<int64_t, int64_t>> split_units(const gphmm_batch *b, int64_t chunk_cells, int64_t chunk_bytes, bool ramp_up,
                                                     int taper_devices = 0) {
    std::vector<std::pair<int64_t, int64_t>> out;
    int64_t u = 0;
    const int64_t full_cells = chunk_cells;
    auto unit_cells = [b](const gphmm_unit &un) {
        const int64_t nr = un.read_end - un.read_begin, nh = un.hap_end - un.hap_begin;
        const int64_t rb = nr ? b->read_off[un.read_end] - b->read_off[un.read_begin] : 0;
        const int64_t hb = nh ? b->hap_off[un.hap_end] - b->hap_off[un.hap_begin] : 0;
        return rb * hb;
    };
    int64_t remaining = 0;
    if (taper_devices > 0)
        for (int64_t k = 0; k < b->n_units; ++k) remaining += unit_cells(b->units[k]);
    while (u < b->n_units) {
        // ramp up: the first chunks are small so that the GPU starts early while the host is still staging
        const size_t ci = out.size();
        // (2.5e8 cells = a handful of regions, then doubling: the planner threads stay ahead of the GPU from there on)
        // (with several devices every device gets a small first chunk, a larger second one, ...)
        const size_t ramp_step = ci / (size_t)std::max(1, taper_devices);
        chunk_cells = (!ramp_up || ramp_step >= 16) ? full_cells : std::min<int64_t>(full_cells, (int64_t)250000000 << ramp_step);
        if (taper_devices > 0) chunk_cells = std::min(chunk_cells, std::max<int64_t>(2500000000LL, remaining / (2 * (int64_t)taper_devices)));
        int64_t cells = 0, bytes = 0, pairs = 0, u_end = u;
        int64_t r_lo = INT64_MAX, r_hi = 0;
        while (u_end < b->n_units) {
            const gphmm_unit &un = b->units[u_end];
            const int64_t nr = un.read_end - un.read_begin, nh = un.hap_end - un.hap_begin;
            const int64_t hb = nh ? b->hap_off[un.hap_end] - b->hap_off[un.hap_begin] : 0;
            const int64_t n_lo = std::min(r_lo, nr ? un.read_begin : r_lo), n_hi = std::max(r_hi, nr ? un.read_end : r_hi);
            const int64_t span = n_hi > n_lo ? b->read_off[n_hi] - b->read_off[n_lo] : 0;
            const int64_t c = unit_cells(un);
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
        remaining -= cells;
        u = u_end;
    }
    return out;
}

// Plans the shared (prefix-compressed) stream of one unit: haplotypes sorted lexicographically, pass i+1 resumes
// from a snapshot taken at the last column it shares with its predecessors.  Appends to c.sstreams / pass_info /
// segments and returns the unit's schedule.  With share == false every pass starts from column 1 in input order.
// Lexicographic order of a unit's haplotypes (input order when sharing is off).
void sorted_hap_order(const gphmm_batch *b, const gphmm_unit &un, bool share, std::vector<int> &order) {
    const int n = (int)(un.hap_end - un.hap_begin);
    auto hap_ptr = [&](int k) { return b->hap_bases + b->hap_off[un.hap_begin + k]; };
    auto hap_len = [&](int k) { return (uint32_t)(b->hap_off[un.hap_begin + k + 1] - b->hap_off[un.hap_begin + k]); };
    order.resize(n);
    for (int k = 0; k < n; ++k) order[k] = k;
    if (share)
        std::sort(order.begin(), order.end(), [&](int x, int y) {
            const uint32_t lx = hap_len(x), ly = hap_len(y);
            const int cmp = memcmp(hap_ptr(x), hap_ptr(y), std::min(lx, ly));
            if (cmp != 0) return cmp < 0;
            if (lx != ly) return lx < ly;
            return x < y;
        });
}

// Scratch of plan_unit_sharing, reused across the units of a chunk (no heap traffic per unit once warm).
struct SharingScratch {
    struct Snap { int pass; uint32_t depth, pos; int slot; uint32_t free_after; };
    std::vector<Snap> snaps;
    std::vector<uint32_t> r, pass_start, end_pos, lcp, n_pad, pts;
    std::vector<int> snap_of_pass, snap_by_pos, order;
};

UnitSched plan_unit_sharing(const gphmm_batch *b, const gphmm_unit &un, uint32_t hap_first_local, bool share, ChunkPlan &c,
                            int64_t sum_read_len, const std::vector<int> &full_order, int g_first, int g_count, SharingScratch &ws,
                            int narrow_schedules = 0) {  // bit 0: also the 16-step-window schedule, bit 1: the 8-step-window one
    constexpr uint32_t SPACING = 32;
    static const uint32_t NEAR_DEPTHS = getenv("GPHMM_NEAR_DEPTHS") ? (uint32_t)std::max(1, atoi(getenv("GPHMM_NEAR_DEPTHS"))) : 96u;  // how far below the shared depth to look
    static const uint32_t MIN_DEPTH = getenv("GPHMM_MIN_DEPTH") ? (uint32_t)std::max(32, atoi(getenv("GPHMM_MIN_DEPTH"))) : 32u;  // tuning knob
    UnitSched us;
    memset(&us, 0, sizeof us);
    const int n = g_count;  // haplotypes of this group: full_order[g_first .. g_first + g_count)
    us.pass_first = (uint32_t)c.pass_info.size();
    us.n_passes = (uint32_t)n;
    us.seg_first = (uint32_t)c.segments.size();
    c.sstreams.append_fill(STREAM_PAD, (uint8_t)CODE_NULL);
    us.sstream_off = (uint32_t)c.sstreams.size();
    if (n == 0) return us;
    auto hap_ptr = [&](int k) { return b->hap_bases + b->hap_off[un.hap_begin + k]; };
    auto hap_len = [&](int k) { return (uint32_t)(b->hap_off[un.hap_begin + k + 1] - b->hap_off[un.hap_begin + k]); };
    using Snap = SharingScratch::Snap;
    std::vector<int> &order = ws.order;
    order.assign(full_order.begin() + g_first, full_order.begin() + g_first + g_count);
    std::vector<Snap> &snaps = ws.snaps;
    snaps.clear();
    int slot_owner[MAX_SNAP_SLOTS];
    for (int k = 0; k < MAX_SNAP_SLOTS; ++k) slot_owner[k] = -1;
    std::vector<uint32_t> &r = ws.r, &pass_start = ws.pass_start, &end_pos = ws.end_pos, &lcp = ws.lcp, &n_pad = ws.n_pad;
    std::vector<int> &snap_of_pass = ws.snap_of_pass;
    r.assign(n, 0); pass_start.assign(n + 1, 1); end_pos.assign(n, 0); lcp.assign(n, 0); n_pad.assign(n, 0);
    snap_of_pass.assign(n, -1);
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
        while (d + 8 <= m) {  // eight bytes at a time
            uint64_t wx, wy;
            memcpy(&wx, x + d, 8); memcpy(&wy, y + d, 8);
            if (wx != wy) { d += (uint32_t)(__builtin_ctzll(wx ^ wy) >> 3); break; }
            d += 8;
        }
        if (d + 8 > m || x[d] == y[d])  // tail (or the loop ended without a difference)
            while (d < m && x[d] == y[d]) ++d;
        lcp[i] = d;
        if (d < MIN_DEPTH) continue;
        int k = -1;
        // preferred: a snapshot at exactly the shared depth, taken by the latest pass that computed that column; when its
        // window would overlap another snapshot window (two variants less than 32 columns apart), the deepest shallower
        // column that fits -- giving up a few columns beats falling back to a much earlier snapshot
        for (uint32_t dd = d; k < 0 && dd >= MIN_DEPTH && dd + NEAR_DEPTHS > d; --dd) {
            int j = i;
            while (r[j] >= dd) --j;  // r[0] = 0 < dd
            const uint32_t pos = pass_start[j] + (dd - r[j]) - 1;
            for (size_t q = 0; q < snaps.size(); ++q)
                if (snaps[q].pass == j && snaps[q].depth == dd && slot_owner[snaps[q].slot] == (int)q) k = (int)q;
            if (k >= 0) break;
            bool ok = true;
            for (const Snap &sn : snaps) ok = ok && (sn.pos + SPACING <= pos || pos + SPACING <= sn.pos);
            int slot = -1;
            for (int q = 0; q < MAX_SNAP_SLOTS && ok && slot < 0; ++q)
                if (slot_owner[q] < 0 || snaps[slot_owner[q]].free_after < pos) slot = q;
            if (ok && slot >= 0) {
                snaps.push_back({j, dd, pos, slot, 0});
                k = (int)snaps.size() - 1;
                slot_owner[slot] = k;
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
    static const bool slot_histogram = getenv("GPHMM_SLOT_HISTOGRAM") != nullptr;
    if (slot_histogram) {  // debugging aid: how many snapshot slots do units use?
        static std::atomic<long> hist[MAX_SNAP_SLOTS + 1];
        int mx = 0;
        for (const Snap &sn : snaps) mx = std::max(mx, sn.slot + 1);
        if (hist[mx].fetch_add(1) % 5000 == 4999 || mx == MAX_SNAP_SLOTS) {
            fprintf(stderr, "[gpuphmm] snapshot slots used per unit:");
            for (int k = 0; k <= MAX_SNAP_SLOTS; ++k) fprintf(stderr, " %d:%ld", k, hist[k].load());
            fprintf(stderr, "\n");
        }
    }
    // stream + pass table
    for (int i = 0; i < n; ++i) {
        const uint32_t H = hap_len(order[i]);
        uint8_t *dst = c.sstreams.append_raw((H - r[i]) + n_pad[i] + 1);
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
    // during steps [e, e+W) and a snapshot position s during [s, s+W), W = lanes per read (32, or 16 for the half-warp
    // kernels).  END windows never overlap each other (passes span >= 32 positions), snapshot windows never overlap each
    // other (SPACING), so at any step at most one of each is active.  A segment = branch-free steps, then checked steps
    // with one constant (END, snapshot) pair.
    std::vector<int> &snap_by_pos = ws.snap_by_pos;
    snap_by_pos.resize(snaps.size());
    for (size_t q = 0; q < snaps.size(); ++q) snap_by_pos[q] = (int)q;
    std::sort(snap_by_pos.begin(), snap_by_pos.end(), [&](int x, int y) { return snaps[x].pos < snaps[y].pos; });
    auto build_schedule = [&](uint32_t W) {
        std::vector<uint32_t> &pts = ws.pts;
        pts.clear();
        pts.push_back(1);
        for (int i = 0; i < n; ++i) { pts.push_back(end_pos[i]); pts.push_back(end_pos[i] + W); }
        for (const Snap &sn : snaps) { pts.push_back(sn.pos); pts.push_back(sn.pos + W); }
        std::sort(pts.begin(), pts.end());
        pts.erase(std::unique(pts.begin(), pts.end()), pts.end());
        Segment seg;
        auto clear_seg = [&]() { seg.n_free = 0; seg.n_chk = 0; seg.snap_pos = INT32_MIN; seg.snap_slot = 0; seg.end_restore = MAX_SNAP_SLOTS; seg.end_out = 0; };
        clear_seg();
        size_t ie = 0, is = 0;  // first END / snapshot whose window has not expired yet
        for (size_t k = 0; k + 1 < pts.size(); ++k) {
            const uint32_t x = pts[k], y = pts[k + 1];
            while (ie < (size_t)n && end_pos[ie] + W <= x) ++ie;
            while (is < snaps.size() && snaps[snap_by_pos[is]].pos + W <= x) ++is;
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
    };
    build_schedule(32);
    us.n_segs = (uint32_t)c.segments.size() - us.seg_first;
    if (narrow_schedules & 1) {  // the same passes and snapshots with 16-step windows (the last window ends 16 steps earlier)
        us.seg16_first = (uint32_t)c.segments.size();
        build_schedule(16);
        us.n_segs16 = (uint32_t)c.segments.size() - us.seg16_first;
    }
    if (narrow_schedules & 2) {  // quarter-warp tasks: 8-step windows
        us.seg8_first = (uint32_t)c.segments.size();
        build_schedule(8);
        us.n_segs8 = (uint32_t)c.segments.size() - us.seg8_first;
    }
    return us;
}

// Haplotype bytes -> column codes through a 256-entry table.  Returns false when a byte without a code (0xff) was met;
// valid codes are < MAX_CODES = 64, so the OR of everything written has bit 7 set iff some byte was unknown.
inline bool encode_hap(uint8_t *dst, const uint8_t *src, uint32_t H, const uint8_t *lut8) {
    uint32_t acc = 0, j = 0;
    for (; j + 8 <= H; j += 8) {
        const uint8_t c0 = lut8[src[j]], c1 = lut8[src[j + 1]], c2 = lut8[src[j + 2]], c3 = lut8[src[j + 3]];
        const uint8_t c4 = lut8[src[j + 4]], c5 = lut8[src[j + 5]], c6 = lut8[src[j + 6]], c7 = lut8[src[j + 7]];
        acc |= c0 | c1 | c2 | c3 | c4 | c5 | c6 | c7;
        const uint64_t w = (uint64_t)c0 | (uint64_t)c1 << 8 | (uint64_t)c2 << 16 | (uint64_t)c3 << 24 | (uint64_t)c4 << 32 |
                           (uint64_t)c5 << 40 | (uint64_t)c6 << 48 | (uint64_t)c7 << 56;
        memcpy(dst + j, &w, 8);
    }
    for (; j < H; ++j) { const uint8_t cd = lut8[src[j]]; acc |= cd; dst[j] = cd; }
    return (acc & 0x80u) == 0;
}

// When a chunk has too few reads to fill the GPU with one warp per read (a single HaplotypeCaller region is ~100 reads),
// each unit's haplotypes are split into groups and every (read, group) pair becomes a task of its own.
constexpr int64_t TARGET_TASKS = 148 * 28 * 2;
constexpr int64_t HOST_CLASSIFY_MAX_READS = 512;  // chunks up to this many reads are classified on the host (per-region calls)

// steps_mode: 0 = plain likelihoods; 1 = the region steps run on the device first (they change the qualities, so only the
// device can classify the reads); 2 = as 1 with the PCR indel model, which lowers ins and del together.
void plan_chunk(const gphmm_batch *b, int64_t u0, int64_t u1, bool force_fp64, bool share, ChunkPlan &c, int steps_mode = 0) {
    const bool pcr_hint = steps_mode == 2;
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
    // a per-region call (few reads) looks at EVERY read and keeps what it saw: the reads are then classified here, with
    // the rules of phmm_classify_kernel, so that no classify launch and no forward launch without work is needed
    const bool classify_here = !force_fp64 && steps_mode == 0 && n_span > 0 && n_span <= HOST_CLASSIFY_MAX_READS;
    std::vector<uint8_t> &shape = c.host_class;  // first the shape bits (| 0x10: symmetric and within range), then the class ids
    shape.clear();
    if (classify_here) shape.assign((size_t)n_span, 0);
    if (!force_fp64 && n_span > 0) {
        const int64_t stride = classify_here ? 1 : std::max<int64_t>(1, n_span / 256);
        for (int64_t r = 0; r < n_span; r += stride) {
            const int64_t o = b->read_off[c.r_lo + r], e = b->read_off[c.r_lo + r + 1];
            if (e == o) continue;
            const uint8_t qi = b->ins_q[o], qd = b->del_q[o], qc = b->gcp[o];
            const unsigned sh = qual_shape(b->ins_q + o, b->del_q + o, b->gcp + o, (size_t)(e - o));
            const bool flat = (sh & 7u) == 7u;
            bool sym = (sh & 12u) == 12u;
            if (sym && !flat) {
                uint8_t mx = 0;
                for (int64_t i = o; i < e; ++i) mx = std::max(mx, b->ins_q[i]);
                sym = mx <= SYM_MAX_GAP_QUAL;
            } else if (sym) {
                sym = qi <= SYM_MAX_GAP_QUAL;
            }
            if (classify_here) shape[(size_t)r] = (uint8_t)(sh | (sym ? 0x10u : 0u) | 0x20u);  // 0x20: not empty
            if (qi > 127 || qd > 127 || qc > 127) continue;
            // (a flat class needs tMM > 0: the kernels factor it out of the match update)
            const bool tmm_positive = tables().m2m[((std::max(qi, qd) * (std::max(qi, qd) + 1)) >> 1) + std::min(qi, qd)] > 0.0;
            if (flat && tmm_positive && !(pcr_hint && qi == qd)) {  // the PCR indel model (region steps) will lower ins and del together
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
    if (classify_here) {
        for (int64_t r = 0; r < n_span; ++r) {
            const uint8_t sh = shape[(size_t)r];
            uint8_t cls = CLASS_GENERAL;
            if (sh & 0x20u) {
                const int64_t o = b->read_off[c.r_lo + r];
                const uint8_t qi = b->ins_q[o], qd = b->del_q[o], qc = b->gcp[o];
                if ((sh & 7u) == 7u)
                    for (int k = 0; k < c.n_classes; ++k)
                        if (c.class_qi[k] == qi && c.class_qd[k] == qd && c.class_qc[k] == qc) cls = (uint8_t)k;
                if (cls == CLASS_GENERAL && (sh & 0x10u))
                    for (int k = 0; k < c.n_sym; ++k)
                        if (c.sym_qc[k] == qc) cls = (uint8_t)(MAX_FLAT_CLASSES + k);
            }
            shape[(size_t)r] = cls;
        }
    }
    memset(c.class_count, 0, sizeof c.class_count);

    // haplotype alphabet of the chunk: A C G T are fixed codes, any other byte value (N included) gets the next free code
    uint8_t lut8[256];
    memset(lut8, 0xff, sizeof lut8);
    memset(c.code_byte, 0, sizeof c.code_byte);
    // (N is not a fixed code: assembled haplotypes rarely contain it, and every code costs a row of each warp's prior table,
    // i.e. shared memory and therefore resident warps)
    const char fixed[4] = {'A', 'C', 'G', 'T'};
    for (int i = 0; i < 4; ++i) { lut8[(uint8_t)fixed[i]] = (uint8_t)(CODE_FIRST_BASE + i); c.code_byte[CODE_FIRST_BASE + i] = (uint8_t)fixed[i]; }
    c.n_codes = CODE_FIRST_BASE + 4;

    c.streams.clear(); c.hap_len.clear(); c.hap_stream_off.clear(); c.units.clear(); c.tasks.clear();
    c.sstreams.clear(); c.pass_info.clear(); c.segments.clear(); c.unit_sched.clear(); c.skipped_cells = 0; c.computed_columns = 0; c.n_keep = 0;
    c.n_pairs = 0; c.cells = 0; c.max_stream_len = 0; c.max_hap_len = 0;
    int64_t n_reads_with_work = 0;
    for (int64_t u = u0; u < u1; ++u)
        if (b->units[u].hap_end > b->units[u].hap_begin) n_reads_with_work += b->units[u].read_end - b->units[u].read_begin;
    // two reads per warp (half-warp kernels) when the chunk still fills the GPU that way: one wave of 148 SMs x 16 resident
    // warps; smaller chunks keep one read per warp and split the haplotypes of a unit into groups instead
    static const bool half_warp = getenv("GPHMM_NO_HALFWARP") == nullptr;  // A/B switch: every read on a full warp
    const bool pair_reads_ok = half_warp && !force_fp64 && n_reads_with_work >= 148 * 16 * 2;
    const int64_t want_groups = pair_reads_ok ? 1 : (n_reads_with_work > 0 ? (TARGET_TASKS + n_reads_with_work - 1) / n_reads_with_work : 1);
    std::vector<Task> raw;
    std::vector<uint8_t> bucket_of;
    raw.reserve((size_t)(c.r_hi - c.r_lo));
    bucket_of.reserve((size_t)(c.r_hi - c.r_lo));
    {
        // exact size of the full streams, upper bound of the shared ones (every pass <= its haplotype + END + 32 pad columns)
        size_t total = 0, n_h = 0;
        for (int64_t u = u0; u < u1; ++u) {
            const gphmm_unit &un = b->units[u];
            const size_t nh = (size_t)(un.hap_end - un.hap_begin);
            total += (nh ? (size_t)(b->hap_off[un.hap_end] - b->hap_off[un.hap_begin]) : 0) + nh + STREAM_PAD;
            n_h += nh;
        }
        c.streams.reserve(total + 64);
        c.sstreams.reserve(total + n_h * 32 + (size_t)(u1 - u0) * STREAM_PAD * (size_t)std::max<int64_t>(1, want_groups) + 64);
        c.hap_len.reserve(n_h); c.hap_stream_off.reserve(n_h); c.pass_info.reserve(n_h);
        c.units.reserve((size_t)(u1 - u0)); c.unit_sched.reserve((size_t)(u1 - u0)); c.segments.reserve(n_h * 5);
    }
    SharingScratch scratch;
    std::vector<int> order;
    std::vector<uint64_t> pair_reads[N_PAIR_BUCKETS];  // per half-warp bucket: (slot of the last row << 32) | unit-local read index
    std::vector<uint64_t> quad_cand;                   // (read length << 32) | unit-local read index
    std::vector<uint8_t> in_quad;
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
        c.streams.append_fill(STREAM_PAD, (uint8_t)CODE_NULL);  // fill/drain codes of the fast kernels
        const uint32_t stream_off = (uint32_t)c.streams.size();
        uint32_t max_h = 1;
        int64_t sum_h = 0;
        {
            const int64_t hap_bytes = nh ? b->hap_off[un.hap_end] - b->hap_off[un.hap_begin] : 0;
            size_t w = c.streams.size();
            c.streams.append_raw((size_t)hap_bytes + nh);
            uint8_t *dst = c.streams.data();
            for (int64_t h = un.hap_begin; h < un.hap_end; ++h) {
                const int64_t ho = b->hap_off[h];
                const uint32_t H = (uint32_t)(b->hap_off[h + 1] - ho);
                c.hap_len.push_back(H);
                c.hap_stream_off.push_back((uint32_t)w);
                const uint8_t *src = b->hap_bases + ho;
                // a byte value seen for the first time is rare: give it a code and redo the haplotype
                while (!encode_hap(dst + w, src, H, lut8)) {
                    for (uint32_t j = 0; j < H; ++j)
                        if (lut8[src[j]] == 0xff) {
                            if (c.n_codes >= MAX_CODES) throw Error(GPHMM_ERR_ALPHABET, "too many distinct haplotype byte values");
                            lut8[src[j]] = (uint8_t)c.n_codes;
                            c.code_byte[c.n_codes++] = src[j];
                        }
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
        // quarter-warp tasks: four (or three) reads of the same length, up to 159 bases, on one warp
        quad_cand.clear();
        bool any_quad = false;
        // (only where the sampled reads showed a flat or symmetric quality class: general reads gain nothing from it)
        if (pair_reads_ok && nh && c.n_classes + c.n_sym > 0) {
            bool in_order = true;  // reads of one length (the usual case) arrive sorted
            for (uint32_t r = 0; r < nr; ++r) {
                const uint32_t R = c.read_off[d.read_first + r + 1] - c.read_off[d.read_first + r];
                if (quad_bucket_of_read(R) < 0) continue;
                const uint64_t key = ((uint64_t)R << 32) | r;
                in_order = in_order && (quad_cand.empty() || quad_cand.back() < key);
                quad_cand.push_back(key);
            }
            if (!in_order) std::sort(quad_cand.begin(), quad_cand.end());
            for (size_t i = 0; i + 2 < quad_cand.size() && !any_quad; ++i) any_quad = (quad_cand[i] >> 32) == (quad_cand[i + 2] >> 32);
        }
        // which of the narrow-window schedules the unit's tasks will use: 16-step windows for half-warp tasks (two or more
        // reads are left for them), 8-step windows for quarter-warp tasks.  (The 32-step schedule is always built; a kernel
        // that finds no narrower one falls back to it.)
        int narrow = 0;
        if (pair_reads_ok && nh) {
            uint32_t n_pair_cand = 0;
            size_t qi = 0;  // quad_cand is sorted by (length, read): runs of equal length leave len % 4 reads (0 if that is 3)
            uint32_t in_quads = 0;
            while (any_quad && qi < quad_cand.size()) {
                size_t qj = qi;
                while (qj < quad_cand.size() && (quad_cand[qj] >> 32) == (quad_cand[qi] >> 32)) ++qj;
                const size_t len = qj - qi, left = len % 4 == 3 ? 0 : (len >= 3 ? len % 4 : len);
                in_quads += (uint32_t)(len - left);
                qi = qj;
            }
            for (uint32_t r = 0; r < nr; ++r) {
                const uint32_t R = c.read_off[d.read_first + r + 1] - c.read_off[d.read_first + r];
                if (pair_bucket_of_read(R) >= 0) ++n_pair_cand;
            }
            narrow = (n_pair_cand - std::min(n_pair_cand, in_quads) >= 2 ? 1 : 0) | (any_quad ? 2 : 0);
        }
        {
            sorted_hap_order(b, un, share && !force_fp64, order);
            for (int gi = 0; gi < n_groups; ++gi) {
                const int g0 = (int)((int64_t)nh * gi / n_groups), g1 = (int)((int64_t)nh * (gi + 1) / n_groups);
                c.unit_sched.push_back(plan_unit_sharing(b, un, d.hap_first, share && !force_fp64, c, fast_read_len, order, g0, g1 - g0, scratch, narrow));
            }
        }
        if (nh == 0) continue;
        for (auto &v : pair_reads) v.clear();
        in_quad.assign(nr, 0);
        for (size_t i = 0; any_quad && i < quad_cand.size();) {
            size_t j = i;
            while (j < quad_cand.size() && (quad_cand[j] >> 32) == (quad_cand[i] >> 32)) ++j;
            const uint32_t R = (uint32_t)(quad_cand[i] >> 32);
            const int qb = quad_bucket_of_read(R);
            // groups of four; a remainder of three still beats a pair plus a single (two or one left over: the half-warp path)
            for (; j - i >= 3; i += std::min<size_t>(4, j - i)) {
                const size_t m = std::min<size_t>(4, j - i);
                uint32_t rr[4] = {NO_READ, NO_READ, NO_READ, NO_READ};
                for (size_t q = 0; q < m; ++q) { rr[q] = d.read_first + (uint32_t)quad_cand[i + q]; in_quad[(uint32_t)quad_cand[i + q]] = 1; }
                Task t;
                t.read = rr[0]; t.stream_off = rr[1]; t.n_haps = rr[2]; t.hap_first = rr[3];
                t.out_base = d.out_base - d.read_first * nh;  // sums of read r (chunk-local) start at out_base + r * nh (mod 2^32)
                t.stream_len = nh;
                t.c0_exp = d.c0_exp; t.unit = sched_first;
                raw.push_back(t);
                bucket_of.push_back((uint8_t)qb);
                ++bucket_count[qb];
                for (size_t q = 0; q < m; ++q) c.cells += (int64_t)R * sum_h;
                if (!c.host_class.empty()) {  // reads classified on the host: which kernels find work in this bucket
                    int seen[4], n_seen = 0;
                    for (size_t q = 0; q < m; ++q) {
                        const uint8_t cq = c.host_class[rr[q]];
                        const int ci = cq == CLASS_GENERAL ? MAX_FLAT_CLASSES + MAX_SYM_CLASSES : cq;
                        bool dup = false;
                        for (int w = 0; w < n_seen; ++w) dup = dup || seen[w] == ci;
                        if (!dup) { seen[n_seen++] = ci; ++c.class_count[qb][ci]; }
                    }
                }
            }
            i = j;
        }
        auto emit = [&](Task &t, uint8_t bucket, int n_t, uint32_t rl_a, uint32_t rl_b) {
            for (int gi = 0; gi < n_t; ++gi) {
                t.unit = sched_first + (uint32_t)gi;
                raw.push_back(t);
                bucket_of.push_back(bucket);
                ++bucket_count[bucket];
            }
            if (!c.host_class.empty()) {
                const uint8_t ca = c.host_class[rl_a];
                c.class_count[bucket][ca == CLASS_GENERAL ? MAX_FLAT_CLASSES + MAX_SYM_CLASSES : ca] += (uint32_t)n_t;
                if (rl_b != NO_READ && c.host_class[rl_b] != ca) {
                    const uint8_t cb = c.host_class[rl_b];
                    c.class_count[bucket][cb == CLASS_GENERAL ? MAX_FLAT_CLASSES + MAX_SYM_CLASSES : cb] += (uint32_t)n_t;
                }
            }
        };
        auto emit_single = [&](uint32_t r, uint32_t R) {  // one read per warp
            Task t;
            t.read = d.read_first + r; t.stream_off = stream_off; t.stream_len = stream_len;
            t.out_base = d.out_base + r * nh; t.c0_exp = d.c0_exp; t.n_haps = nh; t.hap_first = d.hap_first; t.unit = sched_first;
            // fast kernels need two spare rows below the read (accumulator row + row-0 carrier): R + 2 <= 32 K
            const uint32_t k = (R + 1) / 32 + 1;
            const uint8_t bucket = force_fp64 ? 0 : (k <= 8 ? (uint8_t)(k - 1) : (uint8_t)8);
            // one task per haplotype group for the fast kernels; the striped / fp64 kernels sweep the full stream once
            emit(t, bucket, bucket < 8 && !force_fp64 ? n_groups : 1, d.read_first + r, NO_READ);
        };
        for (uint32_t r = 0; r < nr; ++r) {
            if (in_quad[r]) continue;  // on a quarter-warp task
            const uint32_t rl = d.read_first + r;
            const uint32_t R = c.read_off[rl + 1] - c.read_off[rl];
            c.cells += (int64_t)R * sum_h;
            const int pb = pair_reads_ok ? pair_bucket_of_read(R) : -1;
            if (pb >= 0) {  // paired below, once every read of the unit is known
                pair_reads[pb - FIRST_PAIR_BUCKET].push_back(((uint64_t)((R - 1) % (uint32_t)pair_bucket_rows(pb)) << 32) | r);
                continue;
            }
            emit_single(r, R);
        }
        // half-warp buckets: two reads per task, both with their last row on the same register slot (R - 1) mod K
        // (phmm_flat_f32_kernel<K, SYM, 16>); a read that finds no partner takes a full warp like the shorter ones
        for (int p = 0; p < N_PAIR_BUCKETS; ++p) {
            std::vector<uint64_t> &v = pair_reads[p];
            if (v.empty()) continue;
            std::sort(v.begin(), v.end());
            for (size_t i = 0; i < v.size();) {
                const uint32_t ra = (uint32_t)v[i];
                const bool both = i + 1 < v.size() && (v[i + 1] >> 32) == (v[i] >> 32);
                if (!both) {
                    emit_single(ra, c.read_off[d.read_first + ra + 1] - c.read_off[d.read_first + ra]);
                    ++i;
                    continue;
                }
                const uint32_t rb = (uint32_t)v[i + 1];
                Task t;
                t.read = d.read_first + ra; t.out_base = d.out_base + ra * nh;
                t.stream_off = d.read_first + rb; t.stream_len = d.out_base + rb * nh;
                t.c0_exp = d.c0_exp; t.n_haps = nh; t.hap_first = d.hap_first; t.unit = sched_first;
                emit(t, (uint8_t)(FIRST_PAIR_BUCKET + p), n_groups, d.read_first + ra, d.read_first + rb);
                i += 2;
            }
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

