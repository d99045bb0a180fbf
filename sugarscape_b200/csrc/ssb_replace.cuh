// ssb_replace.cuh -- replaceDeadAgents -> configureAgents (reference sugarscape.py:855-859,
// 171-267) on the device, scalable RNG mode.
//
// The reference shuffles the list of empty cells of every active quadrant, shuffles the quadrant
// order and pops cells round-robin.  Mode S defines both orders by hashes (cells ascending by
// hash32(cell), quadrants ascending by hash64(position); pop() takes the LARGEST key first), so the
// cells a step needs are simply the k_q largest hashes among the empty cells of quadrant q:
//   count    empty cells per quadrant (one pass over the grid)
//   plan     need = agentReplacements - population, split round-robin over the hashed quadrant
//            order; per quadrant a hash threshold that keeps ~1.25 k_q + slack candidates
//   collect  second pass: cells whose hash clears the threshold -> candidate list per quadrant
//   sort     radix sort of each candidate list (ssb_sort.cuh)
//   create   replacement i takes the (i / Q)-th largest key of quadrant order[i % Q], its endowment
//            (cursor + i) mod len, id next_id + i; tribe tags per quadrant come from the substream
//            of the new id (generateTribeTags, sugarscape.py:548-559)
//   finish   counters, free-slot stack (the slots of this step's dead stay untouched until the
//            stats reduction has read them)
// Sharded over x-slabs (ssb_shard.cuh): every rank counts and collects on its own slab; populations,
// empty-cell counts and candidate counts are all-gathered with the box-wide barrier, every rank
// gathers all candidates over NVLink, sorts them redundantly (they are few) and creates exactly the
// replacements whose cell lies on its slab -- ids and endowments depend only on the global index i.
#pragma once
#include "ssb_rules.cuh"

namespace ssb {

constexpr unsigned long long SALT_CELL = 0x52455043454C4Cull;
constexpr unsigned long long SALT_QUAD = 0x5245505155414Dull;
enum { SITE_NEWTAGS = 4 };

struct ReplacePlan {
    unsigned long long empty[4];       // empty cells per active quadrant position
    long long cand_count[4];           // candidates collected (int64: also the radix sort's n)
    unsigned long long varying[2];     // OR = ~0, AND = 0: every radix digit counts
    unsigned int threshold[4];
    int quota[4];                      // replacements taken from each quadrant position
    int qorder[4];
    int Q;
    long long need, older;
    unsigned long long created;        // replacements created by this rank
    unsigned long long msg[SHARD_PAYLOAD];   // payload of the next box-wide exchange
};

// position (0-based) of the cell's quadrant among the ACTIVE quadrants (sugarscape.py:475-497), or -1
__device__ __forceinline__ int quadrant_of_cell(const ssb_params_t &p, int x, int y) {
    const int W = p.width, H = p.height, qw = p.quad_w, qh = p.quad_h;
    int q = 0;
    if (x < qw && y < qh) q = 1;
    else if (x >= W - qw && y < qh) q = 2;
    else if (x >= W - qw && y >= H - qh) q = 3;
    else if (x < qw && y >= H - qh) q = 4;
    if (q == 0 || !((p.quad_mask >> (q - 1)) & 1)) return -1;
    return __popc(p.quad_mask & ((1 << (q - 1)) - 1));
}

__global__ void k_repl_reset(ReplacePlan *rp) {
    if (threadIdx.x < 4) { rp->empty[threadIdx.x] = 0; rp->cand_count[threadIdx.x] = 0; rp->quota[threadIdx.x] = 0; }
    if (threadIdx.x == 0) { rp->varying[0] = ~0ull; rp->varying[1] = 0ull; rp->need = 0; rp->older = 0; rp->Q = 0; rp->created = 0; }
}

__global__ void __launch_bounds__(256) k_repl_count(DevCtx c, ReplacePlan *rp) {
    const bool sharded = c.sh.world > 1;
    if (!sharded && c.a.counters[SSB_C_N_LIST] >= c.p.agent_replacements) return;
    __shared__ unsigned int cnt[4];
    if (threadIdx.x < 4) cnt[threadIdx.x] = 0;
    __syncthreads();
    const int H = c.p.height;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long cell_lo = sharded ? c.sh.cell_lo : 0, cell_hi = sharded ? c.sh.cell_hi : c.g.n_cells;
    unsigned int m0 = 0, m1 = 0, m2 = 0, m3 = 0;   // per thread, folded per warp: no same-address contention
    for (long long i = cell_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cell_hi; i += stride) {
        if (c.g.occ[i] >= 0) continue;
        const int x = (int)((unsigned int)i / (unsigned int)H), y = (int)i - x * H;      // n_cells < 2^31
        const int q = quadrant_of_cell(c.p, x, y);
        m0 += q == 0; m1 += q == 1; m2 += q == 2; m3 += q == 3;
    }
    const unsigned int mine[4] = {m0, m1, m2, m3};
#pragma unroll
    for (int q = 0; q < 4; q++) {
        unsigned int v = mine[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(&cnt[q], v);
    }
    __syncthreads();
    if (threadIdx.x < 4 && cnt[threadIdx.x]) atomicAdd(&rp->empty[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
}

// sharded: {population, empty cells per quadrant} of this rank, for the all-gather before the plan
__global__ void k_repl_pack_counts(DevCtx c, ReplacePlan *rp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    rp->msg[0] = (unsigned long long)c.a.counters[SSB_C_N_LIST];
    for (int q = 0; q < 4; q++) rp->msg[1 + q] = rp->empty[q];
    rp->msg[5] = rp->msg[6] = 0;
}

__global__ void k_repl_plan(DevCtx c, ReplacePlan *rp, long long endow_count, long long cand_capacity) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    c.a.counters[SSB_C_N_REPLACED] = 0;
    long long population = c.a.counters[SSB_C_N_LIST];
    if (c.sh.world > 1) {            // box-wide totals; rp->empty becomes the global count on every rank
        population = 0;
        for (int q = 0; q < 4; q++) rp->empty[q] = 0;
        for (int r = 0; r < c.sh.world; r++) {
            population += (long long)c.sh.peer_vals[r * SHARD_PAYLOAD];
            for (int q = 0; q < 4; q++) rp->empty[q] += c.sh.peer_vals[r * SHARD_PAYLOAD + 1 + q];
        }
    }
    long long need = (long long)c.p.agent_replacements - population;
    const int Q = __popc(c.p.quad_mask & 15);
    long long total = 0;
    for (int q = 0; q < Q; q++) total += (long long)rp->empty[q];
    if (need <= 0 || endow_count <= 0 || Q == 0 || total == 0) { rp->need = 0; return; }
    if (need > total) need = total;
    const uint64_t key = step_key(c.p.seed, c.t);
    int qorder[4];
    uint64_t qk[4];
    for (int q = 0; q < Q; q++) { qorder[q] = q; qk[q] = mix64(key ^ SALT_QUAD ^ (uint64_t)q); }
    for (int i = 1; i < Q; i++)
        for (int j = i; j > 0 && (qk[qorder[j - 1]] > qk[qorder[j]] || (qk[qorder[j - 1]] == qk[qorder[j]] && qorder[j - 1] > qorder[j])); j--) {
            const int tq = qorder[j]; qorder[j] = qorder[j - 1]; qorder[j - 1] = tq;
        }
    for (int j = 0; j < Q; j++) {
        const int q = qorder[j];
        rp->qorder[j] = q;
        const long long k = need / Q + (j < need % Q ? 1 : 0);
        rp->quota[q] = (int)k;
        const double E = (double)rp->empty[q];
        double keep = 1.25 * (double)k + 4.0 * sqrt((double)k) + 64.0;      // expected candidates
        if (keep > (double)cand_capacity) keep = (double)cand_capacity;
        const double f = E > 0 ? keep / E : 1.0;
        rp->threshold[q] = f >= 1.0 ? 0u : (unsigned int)(4294967296.0 * (1.0 - f));
        if (k > (long long)rp->empty[q]) c.a.counters[SSB_C_OVERFLOW] = 4;    // the reference would raise IndexError
    }
    rp->Q = Q;
    rp->need = need;
    rp->older = c.a.counters[SSB_C_N_FREE] - c.a.counters[SSB_C_N_DEAD];
}

__global__ void __launch_bounds__(256) k_repl_collect(DevCtx c, ReplacePlan *rp, uint64_t *cand, long long cand_capacity) {
    if (rp->need <= 0) return;
    const int H = c.p.height;
    const uint64_t key = step_key(c.p.seed, c.t);
    const long long stride = (long long)gridDim.x * blockDim.x;
    const bool sharded = c.sh.world > 1;
    const long long cell_lo = sharded ? c.sh.cell_lo : 0, cell_hi = sharded ? c.sh.cell_hi : c.g.n_cells;
    for (long long i = cell_lo + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cell_hi; i += stride) {
        if (c.g.occ[i] >= 0) continue;
        const int x = (int)(i / H), y = (int)(i - (long long)x * H);
        const int q = quadrant_of_cell(c.p, x, y);
        if (q < 0) continue;
        const unsigned int h = (unsigned int)(mix64(key ^ SALT_CELL ^ (uint64_t)i) >> 32);
        if (h < rp->threshold[q]) continue;
        const long long at = (long long)atomicAdd((unsigned long long *)&rp->cand_count[q], 1ull);
        if (at < cand_capacity) cand[(long long)q * cand_capacity + at] = ((uint64_t)h << 32) | (uint32_t)i;
        else c.a.counters[SSB_C_OVERFLOW] = 5;
    }
}

__global__ void k_repl_clamp(ReplacePlan *rp, long long cand_capacity) {
    if (threadIdx.x < 4 && rp->cand_count[threadIdx.x] > cand_capacity) rp->cand_count[threadIdx.x] = cand_capacity;
    __syncwarp();
    if (threadIdx.x == 0) {          // sharded: candidate counts of this rank for the all-gather
        for (int q = 0; q < 4; q++) rp->msg[q] = (unsigned long long)rp->cand_count[q];
        rp->msg[4] = rp->msg[5] = rp->msg[6] = 0;
    }
}

// sharded: copy every rank's candidates (rank r's list of quadrant q starts at word
// r * gather_stride + q * local_capacity of the stitched scratch) into the local buffer, rank
// after rank; the sort that follows makes the order of arrival irrelevant
__global__ void __launch_bounds__(256) k_repl_gather(DevCtx c, ReplacePlan *rp, uint64_t *cand, long long cand_capacity,
                                                     long long local_capacity) {
    if (rp->need <= 0) return;
    const long long stride = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (int q = 0; q < 4; q++) {
        long long base = 0;
        for (int r = 0; r < c.sh.world; r++) {
            const long long n = (long long)c.sh.peer_vals[r * SHARD_PAYLOAD + q];
            const unsigned long long *src = c.sh.gather + (long long)r * c.sh.gather_stride + (long long)q * local_capacity;
            for (long long i = tid; i < n; i += stride)
                if (base + i < cand_capacity) cand[(long long)q * cand_capacity + base + i] = src[i];
            base += n;
        }
    }
}

__global__ void k_repl_gather_counts(DevCtx c, ReplacePlan *rp, long long cand_capacity) {
    if (threadIdx.x >= 4) return;
    long long n = 0;
    for (int r = 0; r < c.sh.world; r++) n += (long long)c.sh.peer_vals[r * SHARD_PAYLOAD + threadIdx.x];
    if (n > cand_capacity) { n = cand_capacity; c.a.counters[SSB_C_OVERFLOW] = 5; }
    rp->cand_count[threadIdx.x] = n;
}

__global__ void __launch_bounds__(256) k_repl_create(DevCtx c, ReplacePlan *rp, const uint64_t *cand, long long cand_capacity,
                                                     ssb_endowments_t en) {
    const long long need = rp->need;
    if (need <= 0) return;
    const ssb_agents_t &a = c.a;
    const int Q = rp->Q;
    const long long n_list = a.counters[SSB_C_N_LIST], next_id = a.counters[SSB_C_NEXT_ID];
    const long long cursor = a.counters[SSB_C_ENDOW_INDEX], n_slots = a.counters[SSB_C_N_SLOTS], older = rp->older;
    const uint64_t key = step_key(c.p.seed, c.t);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < need; i += stride) {
        const int q = rp->qorder[i % Q];
        const long long rank = i / Q, have = rp->cand_count[q];
        if (rank >= have) { a.counters[SSB_C_OVERFLOW] = 6; continue; }      // threshold too tight (not expected)
        const int cell = (int)(cand[(long long)q * cand_capacity + (have - 1 - rank)] & 0xffffffffull);
        long long j = i;                   // index among the replacements THIS rank creates
        if (c.sh.world > 1) {
            if (cell < c.sh.cell_lo || cell >= c.sh.cell_hi) continue;       // another rank's slab
            j = (long long)atomicAdd(&rp->created, 1ull);
        }
        const bool bump = j >= older;          // (a recycled slot may lie in any rank's range; the bump allocation may not)
        long long slot = bump ? n_slots + (j - older) : (long long)a.free_slots[older - 1 - j];
        const long long slot_hi = c.sh.world > 1 ? c.sh.slot_hi : a.capacity;
        if ((bump && slot >= slot_hi) || n_list + j >= a.capacity) { a.counters[SSB_C_OVERFLOW] = 1; continue; }
        const int s = (int)slot;
        const long long ei = (cursor + i) % en.count;
        const int id = (int)(next_id + i);
        a.id[s] = id; a.pos[s] = cell; a.born[s] = c.t; a.cod[s] = SSB_ALIVE;
        a.sugar[s] = a.start_sugar[s] = en.sugar[ei];
        a.spice[s] = a.start_spice[s] = en.spice[ei];
        a.sugar_metab[s] = en.sugar_metab[ei]; a.spice_metab[s] = en.spice_metab[ei];
        a.vision[s] = en.vision[ei]; a.movement[s] = en.movement[ei];
        a.age[s] = 0; a.max_age[s] = en.max_age[ei]; a.sex[s] = en.sex[ei];
        a.fert_age[s] = en.fert_age[ei]; a.infert_age[s] = en.infert_age[ei];
        a.fert_factor[s] = en.fert_factor[ei]; a.aggression[s] = en.aggression[ei];
        a.lookahead[s] = en.lookahead[ei]; a.trade_factor[s] = en.trade_factor[ei];
        a.mrs[s] = 1;
        uint32_t tags = en.tags[ei];
        if (c.p.tribe_per_quadrant && c.p.tag_len > 0 && c.p.max_tribes > 0) {
            // generateTribeTags(tribe = quadrant position), sugarscape.py:548-559
            const int L = c.p.tag_len;
            const double tribe_size = (double)(L + 1) / (double)c.p.max_tribes;
            const int min_z = (int)floor(q * tribe_size);
            const int max_z = min((int)floor((q + 1) * tribe_size) - 1, L);
            RngStream rng{key, (uint32_t)id, 0u, 0u, nullptr};
            rng.at(SITE_NEWTAGS);
            const int zeroes = min_z + (int)rand_below(rng, (uint32_t)(max_z - min_z + 1));
            uint8_t bits[32];
            for (int k = 0; k < L; k++) bits[k] = k < zeroes ? 0 : 1;
            shuffle(rng, bits, L);
            tags = 0;
            for (int k = 0; k < L; k++) tags |= (uint32_t)bits[k] << k;
        }
        a.tags[s] = tags;
        a.tribe[s] = find_tribe(c, tags);
        a.last_moved[s] = -1; a.last_sugar[s] = 0; a.last_spice[s] = 0;
        a.n_neighborhood[s] = 0; a.n_valid_moves[s] = 0;
        a.happiness[s] = 0; a.wealth_happiness[s] = 0; a.family_happiness[s] = 0;
        a.trade_volume[s] = 0; a.trade_price[s] = 0; a.last_repro[s] = -1;
        a.children_head[s] = -1; a.mates_head[s] = -1; a.father[s] = -1; a.mother[s] = -1;
        c.g.occ[cell] = s;
        a.order[n_list + j] = s;
    }
}

// the free stack keeps [older slots not consumed][this step's dead]; the dead block is rebuilt from
// dead_list (same order as compaction pushed it)
__global__ void __launch_bounds__(256) k_repl_free_stack(DevCtx c, const ReplacePlan *rp) {
    if (rp->need <= 0) return;
    const long long need = c.sh.world > 1 ? (long long)rp->created : rp->need;     // created by this rank
    const long long n_dead = c.a.counters[SSB_C_N_DEAD];
    const long long older_left = rp->older > need ? rp->older - need : 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n_dead; i += stride)
        c.a.free_slots[older_left + i] = c.a.dead_list[i];
}

__global__ void k_repl_finish(DevCtx c, const ReplacePlan *rp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const long long need = rp->need;
    if (need <= 0) return;
    const long long mine = c.sh.world > 1 ? (long long)rp->created : need;         // created by this rank
    const long long older = rp->older;
    const long long bumped = mine > older ? mine - older : 0;
    c.a.counters[SSB_C_N_SLOTS] += bumped;
    c.a.counters[SSB_C_N_FREE] = (older > mine ? older - mine : 0) + c.a.counters[SSB_C_N_DEAD];
    c.a.counters[SSB_C_N_LIST] += mine;
    c.a.counters[SSB_C_NEXT_ID] += need;          // ids and the endowment cursor are box-wide
    c.a.counters[SSB_C_ENDOW_INDEX] += need;
    c.a.counters[SSB_C_N_REPLACED] = mine;
}

}  // namespace ssb
