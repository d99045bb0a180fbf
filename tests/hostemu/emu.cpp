/* emu.cpp -- host-compiled build of the engine's mode E transaction source (TEST INFRASTRUCTURE ONLY,
 * see cuda_shim.h).  Exposes one agent step of the exact mode on HOST arrays laid out like the device
 * SoA (include/ssb.h); environment step, compaction and the reductions of a test run come from the
 * CPU oracle, so what this library adds to the `-m "not gpu"` suite is exactly the device rule code. */
#include "cuda_shim.h"

#include "../../sugarscape_b200/csrc/ssb_exact.cuh"
#include "../../sugarscape_b200/csrc/ssb_flow_core.cuh"
#include "../../sugarscape_b200/csrc/ssb_sweep_core.cuh"

#include <algorithm>
#include <vector>

using namespace ssb;

extern "C" {

int emu_abi_version(void) { return SSB_ABI_VERSION; }

/* mt: 624 state words + the index (the engine's d_mt layout) */
int emu_agent_step_exact(const ssb_params_t *p, const ssb_grid_t *g, const ssb_agents_t *a, uint32_t *mt, int32_t t,
                         double prev_mean_wealth, const uint32_t *endow_bits, const uint32_t *tag_bits, int32_t n_tables,
                         const ssb_ethics_t *ethics, const ssb_ext_t *ext) {
    if (!p || !g || !a || !mt || p->rng_mode != SSB_RNG_EXACT) return -1;
    DevCtx c;
    memset(&c, 0, sizeof(c));
    c.p = *p; c.g = *g; c.a = *a;
    const int rmax = p->max_agent_range;          /* as ssb_bind */
    c.max_dx = rmax < p->width / 2 ? rmax : p->width / 2;
    c.max_dy = rmax < p->height / 2 ? rmax : p->height / 2;
    if (!(p->flags & SSB_F_RADIAL) && 4 * (c.max_dx > c.max_dy ? c.max_dx : c.max_dy) > MAXC_EXACT) return -2;
    c.t = t;
    c.prev_mean_wealth = prev_mean_wealth;
    c.endow_bits = endow_bits; c.tag_bits = tag_bits; c.n_step_tables = n_tables;
    alignas(8) unsigned char scratch[warp_scratch_bytes(MAXC_EXACT)];
    if (ethics || ext) {   /* as ssb_agent_step: the extended instantiation only when an extension is bound */
        if (ethics) c.eth = ethics;
        if (ext) c.x = ext;
        agent_step_exact_body<1, true>(c, mt, make_warp_scratch(scratch, MAXC_EXACT));
    } else {
        agent_step_exact_body<1, false>(c, mt, make_warp_scratch(scratch, MAXC_EXACT));
    }
    return 0;
}

int emu_ext_stats(const ssb_agents_t *a, const ssb_ext_t *ext, int32_t t, ssb_ext_stats_t *out) {
    if (!a || !ext || !out) return -1;
    DevCtx c;
    memset(&c, 0, sizeof(c));
    c.a = *a; c.x = ext; c.t = t;
    ext_stats_body(c, out);
    return 0;
}

int emu_group_stats(const ssb_agents_t *a, const ssb_ethics_t *ethics, const ssb_ext_t *ext, int32_t t, ssb_group_stats_t out[3]) {
    if (!a || !ext || !out) return -1;
    DevCtx c;
    memset(&c, 0, sizeof(c));
    c.a = *a; c.x = ext; c.t = t;
    if (ethics) c.eth = ethics;
    memset(out, 0, 3 * sizeof(ssb_group_stats_t));
    group_stats_body(c, out);
    return 0;
}

int emu_ethics_stats(const ssb_agents_t *a, const ssb_ethics_t *ethics, ssb_ethics_stats_t *out) {
    if (!a || !ethics || !out) return -1;
    DevCtx c;
    memset(&c, 0, sizeof(c));
    c.a = *a; c.eth = ethics;
    ethics_stats_body(c, out);
    return 0;
}


/* Mode S agent step through the engine's static conflict DAG (csrc/ssb_flow_core.cuh: the SAME conflict predicate
 * and the SAME staged-tile scan the CUDA build kernel runs), executed in an ADVERSARIAL topological order:
 * order_mode 0 = FIFO (about level order), 1 = LIFO (depth first, as far from the priority order as the DAG
 * allows), 2 = pseudo-random pick among the ready agents.  The result must equal the oracle's sequential
 * priority order bit for bit -- that is the claim the dataflow scheduler rests on.
 * out_stats (may be null): {agents, edges, longest path, tiles visited}. */
int emu_agent_step_flow(const ssb_params_t *p, const ssb_grid_t *g, const ssb_agents_t *a, int32_t t, double prev_mean_wealth,
                        const uint32_t *endow_bits, const uint32_t *tag_bits, int32_t n_tables, int32_t *born_list,
                        int32_t order_mode, int64_t *out_stats) {
    if (!p || !g || !a || p->rng_mode != SSB_RNG_SCALABLE) return -1;
    DevCtx c;
    memset(&c, 0, sizeof(c));
    c.p = *p; c.g = *g; c.a = *a;
    const int rmax = p->max_agent_range;
    c.max_dx = rmax < p->width / 2 ? rmax : p->width / 2;
    c.max_dy = rmax < p->height / 2 ? rmax : p->height / 2;
    const int RB = c.max_dx > c.max_dy ? c.max_dx : c.max_dy, KB = flow_max_dilation(p->flags), R = flow_halo(RB, KB);
    if (4 * RB > 32 || R > FLOW_MAX_HALO) return -2;
    c.t = t; c.prev_mean_wealth = prev_mean_wealth;
    c.endow_bits = endow_bits; c.tag_bits = tag_bits; c.n_step_tables = n_tables;
    c.born_list = born_list;
    const uint64_t key = step_key(p->seed, t);
    const int W = p->width, H = p->height;
    const int n_list = (int)a->counters[SSB_C_N_LIST];
    a->counters[SSB_C_N_BORN] = 0;
    /* prepare (k_flow_prepare) */
    struct Entry { int slot, pos; uint32_t prio; int rk; };
    std::vector<Entry> entries(n_list);
    std::vector<unsigned long long> desc((size_t)W * H, 0ull);
    std::vector<int> dep(n_list, 0), ready;
    std::vector<std::vector<int>> succ(n_list);
    for (int i = 0; i < n_list; i++) {
        const int s = a->order[i];
        const uint32_t pr = priority_of(key, (uint32_t)a->id[s], p->id_bits);
        const int r = min(min(max(a->vision[s], 0), max(a->movement[s], 0)), max(c.max_dx, c.max_dy));
        const int k = footprint_dilation(c, s);
        entries[i] = {s, a->pos[s], pr, r | (k << 8)};
        if (a->pos[s] >= 0) desc[a->pos[s]] = flow_pack((uint32_t)i, pr, r, k); else ready.push_back(i);
    }
    /* build (k_flow_build): tile by tile, staged with the torus wrap, masks, one scan per agent */
    const int sw = FLOW_TX + 2 * R, sh = FLOW_TY + 2 * R;
    std::vector<unsigned long long> cell((size_t)sw * sh), mask((size_t)sw * 2);
    long long edges = 0, tiles = 0;
    for (int tx = 0; tx * FLOW_TX < W; tx++) for (int ty = 0; ty * FLOW_TY < H; ty++) {
        const int x0 = tx * FLOW_TX - R, y0 = ty * FLOW_TY - R;
        for (int col = 0; col < sw; col++) for (int row = 0; row < sh; row++)
            cell[(size_t)col * sh + row] = desc[(size_t)wrapi(x0 + col, W) * H + wrapi(y0 + row, H)];
        for (int col = 0; col < sw; col++) {
            mask[col * 2] = mask[col * 2 + 1] = 0;
            for (int row = 0; row < sh; row++) if (cell[(size_t)col * sh + row]) mask[col * 2 + (row >> 6)] |= 1ull << (row & 63);
        }
        FlowTile tile;
        tile.cell = cell.data(); tile.mask = mask.data(); tile.sw = sw; tile.sh = sh; tile.halo = R;
        tiles++;
        for (int ix = 0; ix < FLOW_TX && tx * FLOW_TX + ix < W; ix++) for (int iy = 0; iy < FLOW_TY && ty * FLOW_TY + iy < H; iy++) {
            const int sc = (R + ix) * sh + R + iy;
            const unsigned long long da = cell[sc];
            if (!da) continue;
            const int ia = (int)flow_index(da);
            const uint32_t pa = flow_prio(da);
            flow_scan_agent(tile, R + ix, R + iy, da, RB, KB, [&](int, unsigned long long db) {
                if (flow_prio(db) < pa) dep[ia]++;
                else if (flow_prio(db) > pa) { succ[ia].push_back((int)flow_index(db)); edges++; }
            });
            if (dep[ia] == 0) ready.push_back(ia);
        }
    }
    /* execute (k_flow_execute) in the requested topological order */
    alignas(8) unsigned char scratch[warp_scratch_bytes(32)];
    const WarpScratch ws = make_warp_scratch(scratch, 32);
    std::vector<int> level(n_list, 0);
    long long done = 0, depth = 0;
    uint64_t lcg = 0x9E3779B97F4A7C15ull ^ (uint64_t)t;
    size_t fifo_head = 0;
    while (order_mode == 0 ? fifo_head < ready.size() : !ready.empty()) {
        int i;
        if (order_mode == 0) i = ready[fifo_head++];
        else if (order_mode == 1) { i = ready.back(); ready.pop_back(); }
        else {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            const size_t k = (size_t)((lcg >> 33) % ready.size());
            i = ready[k]; ready[k] = ready.back(); ready.pop_back();
        }
        const int s = entries[i].slot;
        AgentRec rec;
        rec.load(c, s);
        RngStream rng{key, (uint32_t)rec.id, 0u, 0u, nullptr};
        if (transaction_move_phase<1, 4>(c, s, rec, rng, ws, 32)) transaction_interact_phase<RngStream, false>(c, s, rec, rng);
        done++;
        if (level[i] + 1 > depth) depth = level[i] + 1;
        for (int b : succ[i]) {
            if (level[i] + 1 > level[b]) level[b] = level[i] + 1;
            if (--dep[b] == 0) ready.push_back(b);
        }
    }
    if (out_stats) { out_stats[0] = n_list; out_stats[1] = edges; out_stats[2] = depth; out_stats[3] = tiles; }
    return done == n_list ? 0 : -3;     /* -3: the graph was not acyclic / a counter never reached zero */
}


/* Mode S agent step through the ORDERED SWEEP (csrc/ssb_sweep_core.cuh: the SAME predecessor scan, done test and --
 * with lean != 0 -- the SAME lean transaction on packed records that the CUDA kernels run).  The list indices are
 * handed out in an arbitrary order and the agents run in passes over the pending entries, in an ADVERSARIAL order:
 * order_mode 0 = ascending list index, 1 = DESCENDING priority (as far from the sequential order as the predecessor
 * masks allow), 2 = pseudo-random.  An entry runs only when sweep_preds_done says so, exactly like a sweep thread.
 * order_mode | 0x100: one priority bucket (every predecessor is a near one) instead of 128 buckets in 32 slabs.
 * out_stats (may be null): {agents, near predecessors, passes, tiles visited, mask lines}. */
int emu_agent_step_sweep(const ssb_params_t *p, const ssb_grid_t *g, const ssb_agents_t *a, int32_t t, double prev_mean_wealth,
                         const uint32_t *endow_bits, const uint32_t *tag_bits, int32_t n_tables, int32_t *born_list,
                         int32_t order_mode, int32_t lean, int64_t *out_stats) {
    if (!p || !g || !a || p->rng_mode != SSB_RNG_SCALABLE) return -1;
    DevCtx c;
    memset(&c, 0, sizeof(c));
    c.p = *p; c.g = *g; c.a = *a;
    const int rmax = p->max_agent_range;
    c.max_dx = rmax < p->width / 2 ? rmax : p->width / 2;
    c.max_dy = rmax < p->height / 2 ? rmax : p->height / 2;
    const int RB = c.max_dx > c.max_dy ? c.max_dx : c.max_dy, KB = flow_max_dilation(p->flags), R = flow_halo(RB, KB);
    const int W = p->width, H = p->height;
    if (4 * RB > 32 || R > SWEEP_MAX_HALO || R >= W || R >= H) return -2;
    if (lean && !(lean_supported(*p) && 4 * RB <= 32)) return -4;
    c.t = t; c.prev_mean_wealth = prev_mean_wealth;
    c.endow_bits = endow_bits; c.tag_bits = tag_bits; c.n_step_tables = n_tables;
    c.born_list = born_list;
    const uint64_t key = step_key(p->seed, t);
    const int n_list = (int)a->counters[SSB_C_N_LIST], n_slots = (int)a->counters[SSB_C_N_SLOTS];
    a->counters[SSB_C_N_BORN] = 0;
    /* prepare (k_sweep_prepare): list indices in a scrambled order */
    std::vector<LeanA> recA(lean ? (size_t)a->capacity : 0);
    std::vector<LeanB> recB(lean ? (size_t)a->capacity : 0);
    std::vector<int2> pcell(lean ? (size_t)lean_tiled_cells(W, H) : 0);
    LeanCtx L;
    L.a = recA.data(); L.b = recB.data(); L.maxc = 4 * RB;
    L.pcell = pcell.data(); L.ty4 = (H + 3) / 4;
    if (lean)   /* k_lean_cells<true>: {cell.agent, cell.sugar} into the tiled cell array */
        for (int x = 0; x < W; x++) for (int y = 0; y < H; y++) pcell[lean_tiled(x, y, L.ty4)] = make_int2(g->occ[(size_t)x * H + y], g->sugar[(size_t)x * H + y]);
    struct Entry { int slot, pos; uint32_t prio; };
    std::vector<Entry> entries;
    std::vector<unsigned long long> desc((size_t)W * H, 0ull);
    std::vector<int> live;
    if (lean) { for (int s = 0; s < n_slots; s++) if (lean_pack_a(c, L, s)) live.push_back(s); }
    else { for (int i = 0; i < n_list; i++) if (a->pos[a->order[i]] >= 0) live.push_back(a->order[i]); }
    if (lean && (int)live.size() != n_list) return -5;
    const int n = (int)live.size();
    {   /* scramble: index = position in a pseudo-random permutation ... */
        uint64_t lcg = 0x2545F4914F6CDD1Dull ^ (uint64_t)t;
        for (int i = n - 1; i > 0; i--) {
            lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
            const int j = (int)((lcg >> 33) % (uint64_t)(i + 1));
            const int tmp = live[i]; live[i] = live[j]; live[j] = tmp;
        }
    }
    /* ... grouped by priority bucket like k_sweep_prepare (128 buckets = 32 slabs of 4, or one bucket: order_mode & 0x100) */
    const int bucket_bits = (order_mode & 0x100) ? 0 : (p->id_bits < 7 ? p->id_bits : 7);
    order_mode &= 0xff;
    const int E = 1 << bucket_bits, prio_shift = p->id_bits - bucket_bits, slab_shift = sweep_slab_shift(bucket_bits, SWEEP_MAX_SLAB_BITS);
    const int n_slabs = sweep_n_slabs(bucket_bits, SWEEP_MAX_SLAB_BITS), slab_prio_shift = prio_shift + slab_shift;
    auto bucket_of = [&](int s) { return (int)(priority_of(key, (uint32_t)a->id[s], p->id_bits) >> prio_shift); };
    std::stable_sort(live.begin(), live.end(), [&](int x, int y) { return bucket_of(x) < bucket_of(y); });
    std::vector<int> bucket_size(E, 0), bucket_fin(E, 0);
    entries.resize(n);
    for (int index = 0; index < n; index++) {
        const int s = live[index], pos = a->pos[s];
        const uint32_t pr = priority_of(key, (uint32_t)a->id[s], p->id_bits);
        const int r = lean_range(c, s), k = lean ? 0 : footprint_dilation(c, s);
        desc[pos] = sweep_desc((uint32_t)index, sweep_lo(pr, r, k));
        entries[index] = {s, pos, pr};
        bucket_size[pr >> prio_shift]++;
        if (lean) lean_pack_b(c, L, s, index, r);
    }
    /* build (k_sweep_build): tile by tile, staged with the torus wrap, one bit-plane per slab, one scan per agent */
    const int sw = FLOW_TX + 2 * R, sh = FLOW_TY + 2 * R;
    std::vector<uint32_t> cell((size_t)sw * sh), preds((size_t)n * SWEEP_COLS, 0u);
    std::vector<unsigned long long> gates((size_t)n, 0ull);
    std::vector<unsigned long long> smask((size_t)n_slabs * sw * 2);
    std::vector<int8_t> tables((size_t)sweep_tables_bytes(RB, KB, R), 0);
    sweep_fill_tables(RB, KB, R, tables.data(), tables.data() + sweep_any_size(RB, KB, R));
    long long bits = 0, tiles = 0, lines = 0;
    for (int tx = 0; tx * FLOW_TX < W; tx++) for (int ty = 0; ty * FLOW_TY < H; ty++) {
        const int x0 = tx * FLOW_TX - R, y0 = ty * FLOW_TY - R;
        std::fill(smask.begin(), smask.end(), 0ull);
        for (int col = 0; col < sw; col++) for (int row = 0; row < sh; row++) {
            const uint32_t lo = (uint32_t)desc[(size_t)wrapi(x0 + col, W) * H + wrapi(y0 + row, H)];
            cell[(size_t)col * sh + row] = lo;
            if (lo) smask[((size_t)(sweep_prio(lo) >> slab_prio_shift) * sw + col) * 2 + (row >> 6)] |= 1ull << (row & 63);
        }
        for (int word = 0; word < sw * 2; word++) sweep_fold_planes(smask.data(), n_slabs, sw, word);
        SweepTile tile;
        tile.cell = cell.data(); tile.smask = smask.data(); tile.sw = sw; tile.sh = sh; tile.halo = R; tile.slab_prio_shift = slab_prio_shift;
        tile.any = tables.data(); tile.pair = tables.data() + sweep_any_size(RB, KB, R);
        tiles++;
        for (int ix = 0; ix < FLOW_TX && tx * FLOW_TX + ix < W; ix++) for (int iy = 0; iy < FLOW_TY && ty * FLOW_TY + iy < H; iy++) {
            const uint32_t la = cell[(size_t)(R + ix) * sh + R + iy];
            if (!la) continue;
            const uint32_t gcell = (uint32_t)((tx * FLOW_TX + ix) * H + ty * FLOW_TY + iy);
            const uint32_t index = (uint32_t)(desc[gcell] >> 32);
            unsigned long long gate = 0ull;
            int n_near = 0;
            sweep_near_preds(tile, R + ix, R + iy, la, RB, KB, [&](int dx, int dy) {
                if (n_near < SWEEP_GATE_PREDS) gate |= sweep_gate_code(dx, dy, R) << (10 * n_near);
                n_near++;
            });
            bits += n_near;
            if (n_near <= SWEEP_GATE_PREDS) { gates[index] = gate; continue; }
            uint32_t *line = preds.data() + (size_t)index * SWEEP_COLS;
            sweep_near_preds(tile, R + ix, R + iy, la, RB, KB, [&](int dx, int dy) { line[dx + R] |= 1u << (dy + R); });
            gates[index] = SWEEP_GATE_LINE;
            lines++;
        }
    }
    /* sweep (k_sweep_lean / k_sweep_generic): passes over the pending entries; an entry runs only when the watermark has
     * passed the slabs before its near slabs and its gate is open, exactly like a sweep thread */
    std::vector<unsigned long long> done(sweep_done_words(W, H), 0ull);
    auto ld = [&](size_t wd) { return done[wd]; };
    const int wpc = sweep_done_wpc(H);
    int watermark = 0;
    while (watermark < E && bucket_size[watermark] == 0) watermark++;
    std::vector<int> pending(n);
    for (int i = 0; i < n; i++) pending[i] = i;
    if (order_mode == 1) std::sort(pending.begin(), pending.end(), [&](int x, int y) { return entries[x].prio > entries[y].prio; });
    alignas(8) unsigned char scratch[warp_scratch_bytes(32)];
    const WarpScratch ws = make_warp_scratch(scratch, 32);
    int32_t st_occ[33];
    int16_t st_sugar[34];
    const LeanStage st = {st_occ, st_sugar, 1};
    long long passes = 0, ran_total = 0;
    uint64_t lcg = 0x9E3779B97F4A7C15ull ^ (uint64_t)t;
    while (!pending.empty()) {
        passes++;
        if (order_mode == 2)
            for (int i = (int)pending.size() - 1; i > 0; i--) {
                lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
                const int j = (int)((lcg >> 33) % (uint64_t)(i + 1));
                const int tmp = pending[i]; pending[i] = pending[j]; pending[j] = tmp;
            }
        std::vector<int> left;
        int ran = 0;
        for (int i : pending) {
            const int pos = entries[i].pos, x = pos / H, y = pos % H;
            const int bucket = (int)(entries[i].prio >> prio_shift);
            if (lean && L.b[i].pos != pos) return -6;
            bool ready = watermark >= sweep_watermark_needed(bucket, slab_shift);
            if (ready) ready = gates[i] == SWEEP_GATE_LINE ? sweep_line_done(preds.data() + (size_t)i * SWEEP_COLS, ld, x, y, W, H, R)
                                                            : sweep_gate_done(gates[i], ld, x, y, W, H, R);
            if (!ready) { left.push_back(i); continue; }
            if (lean) { if (lean_mode(*p) == 0) lean_transaction<0>(c, L, L.b[i], key, st, 1u); else lean_transaction<LEAN_M_ALL>(c, L, L.b[i], key, st, 1u); }
            else {
                const int s = entries[i].slot;
                AgentRec rec;
                rec.load(c, s);
                RngStream rng{key, (uint32_t)rec.id, 0u, 0u, nullptr};
                if (transaction_move_phase<1, 4>(c, s, rec, rng, ws, 32)) transaction_interact_phase<RngStream, false>(c, s, rec, rng);
            }
            done[sweep_done_word(x, y, wpc)] |= 1ull << (y & 63);
            bucket_fin[bucket]++;
            while (watermark < E && bucket_fin[watermark] == bucket_size[watermark]) watermark++;
            ran++;
        }
        if (ran == 0) return -3;     /* nobody could run: the gates deadlock */
        ran_total += ran;
        pending.swap(left);
    }
    if (lean) {
        for (int x = 0; x < W; x++) for (int y = 0; y < H; y++) {      /* k_lean_cells<false> */
            const int2 v = pcell[lean_tiled(x, y, L.ty4)];
            g->occ[(size_t)x * H + y] = v.x; g->sugar[(size_t)x * H + y] = v.y;
        }
        for (int s = 0; s < n_slots; s++) lean_unpack_slot(c, L, s);
    }
    if (out_stats) { out_stats[0] = n; out_stats[1] = bits; out_stats[2] = passes; out_stats[3] = tiles; out_stats[4] = lines; }
    return ran_total == n ? 0 : -3;
}


int emu_flow_conflict(int32_t dx, int32_t dy, int32_t ra, int32_t rb, int32_t ka, int32_t kb) {
    return flow_conflict(dx < 0 ? -dx : dx, dy < 0 ? -dy : dy, ra, rb, ka + kb) ? 1 : 0;
}
int emu_flow_column_reach(int32_t adx, int32_t ra, int32_t ka, int32_t RB, int32_t KB) { return flow_column_reach(adx, ra, ka, RB, KB); }

}
