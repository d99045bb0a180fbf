// ssb_migrate.cuh -- agent migration between x-slabs, last phase of a sharded step.
//
// An agent that walked onto another rank's slab stays correct where it is (all state is reachable
// over NVLink) but every later access to its record would cross the link, so at the end of the step
// the record moves to the rank that owns its cell:
//   out      the source drops the emigrants from its list (tail entries fill the holes: list order
//            carries no meaning in the scalable mode) and writes their slots into one outbox per
//            destination, in its part of the stitched scratch;
//   barrier  all-gather of the outbox lengths;
//   in       the destination takes a slot of its own range for every immigrant, copies the record
//            over the link, repoints occ[] of the cell (on its slab) and appends it to its list;
//   barrier  the copies are done;
//   recycle  the source pushes the old slots onto its free stack.
#pragma once
#include "ssb_rules.cuh"

namespace ssb {

struct MigratePlan {
    unsigned long long out_count[SSB_MAX_RANKS];   // emigrants per destination
    unsigned long long n_out, n_holes, n_keep;
    unsigned long long msg[SHARD_PAYLOAD];
};

__device__ __forceinline__ int shard_owner_of_cell(const ShardCtx &s, long long cell) {
    int r = 0;
    while (r < s.world - 1 && cell >= s.cell_end[r]) r++;
    return r;
}

// payload word of the count "source -> dest" in the source's message (the own rank is skipped)
__device__ __forceinline__ int migrate_word(int source, int dest) { return dest < source ? dest : dest - 1; }

__global__ void k_migrate_reset(MigratePlan *mp) {
    if (threadIdx.x < SSB_MAX_RANKS) mp->out_count[threadIdx.x] = 0;
    if (threadIdx.x == 0) { mp->n_out = 0; mp->n_holes = 0; mp->n_keep = 0; }
}

// emigrants -> outboxes (outbox of destination d: words [d * box_cap, ...) of this rank's scratch part)
__global__ void __launch_bounds__(256) k_migrate_out(DevCtx c, MigratePlan *mp, long long box_cap) {
    const long long n = c.a.counters[SSB_C_N_LIST];
    unsigned long long *mine = c.sh.gather + (long long)c.sh.rank * c.sh.gather_stride;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int s = c.a.order[i];
        const long long cell = c.a.pos[s];
        if (cell >= c.sh.cell_lo && cell < c.sh.cell_hi) continue;
        const int d = shard_owner_of_cell(c.sh, cell);
        const unsigned long long at = atomicAdd(&mp->out_count[d], 1ull);
        if ((long long)at < box_cap) mine[(long long)d * box_cap + at] = (unsigned long long)s;
        else c.a.counters[SSB_C_OVERFLOW] = 8;
        atomicAdd(&mp->n_out, 1ull);
    }
}

__global__ void k_migrate_pack(DevCtx c, MigratePlan *mp, long long box_cap) {
    if (threadIdx.x != 0) return;
    for (int w = 0; w < SHARD_PAYLOAD; w++) mp->msg[w] = 0;
    for (int d = 0; d < c.sh.world; d++) {
        if (d == c.sh.rank) continue;
        if ((long long)mp->out_count[d] > box_cap) mp->out_count[d] = (unsigned long long)box_cap;
        mp->msg[migrate_word(c.sh.rank, d)] = mp->out_count[d];
    }
}

// holes = list positions below the new length held by emigrants; keep = survivors beyond it
__global__ void __launch_bounds__(256) k_migrate_holes(DevCtx c, MigratePlan *mp, int32_t *holes, int32_t *keep) {
    const long long n = c.a.counters[SSB_C_N_LIST], n_new = n - (long long)mp->n_out;
    if (mp->n_out == 0) return;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int s = c.a.order[i];
        const long long cell = c.a.pos[s];
        const bool stays = cell >= c.sh.cell_lo && cell < c.sh.cell_hi;
        if (!stays && i < n_new) holes[atomicAdd(&mp->n_holes, 1ull)] = (int32_t)i;
        if (stays && i >= n_new) keep[atomicAdd(&mp->n_keep, 1ull)] = s;
    }
}

__global__ void __launch_bounds__(256) k_migrate_fill(DevCtx c, MigratePlan *mp, const int32_t *holes, const int32_t *keep) {
    const long long n_holes = (long long)mp->n_holes;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n_holes; k += stride) c.a.order[holes[k]] = keep[k];
}

__global__ void k_migrate_out_finish(DevCtx c, MigratePlan *mp) {
    if (threadIdx.x == 0 && blockIdx.x == 0) c.a.counters[SSB_C_N_LIST] -= (long long)mp->n_out;
}

#define SSB_COPY_FIELD(f) a.f[ns] = a.f[os]

// immigrants: after the all-gather, peer_vals[r][migrate_word(r, me)] records wait in rank r's outbox for this rank
__global__ void __launch_bounds__(256) k_migrate_in(DevCtx c, long long box_cap) {
    const ssb_agents_t &a = c.a;
    const int me = c.sh.rank;
    const long long n_list = a.counters[SSB_C_N_LIST], n_free = a.counters[SSB_C_N_FREE], n_slots = a.counters[SSB_C_N_SLOTS];
    const long long stride = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long base = 0;
    for (int r = 0; r < c.sh.world; r++) {
        if (r == me) continue;
        const long long n = (long long)c.sh.peer_vals[r * SHARD_PAYLOAD + migrate_word(r, me)];
        const unsigned long long *box = c.sh.gather + (long long)r * c.sh.gather_stride + (long long)me * box_cap;
        for (long long k = tid; k < n; k += stride) {
            const long long j = base + k;
            // a recycled slot may lie in ANY rank's range (a child's record is allocated by the acting parent's
            // rank); only the bump allocation is bounded by this rank's range
            const bool bump = j >= n_free;
            const long long slot = bump ? n_slots + (j - n_free) : (long long)a.free_slots[n_free - 1 - j];
            if ((bump && slot >= c.sh.slot_hi) || n_list + j >= a.capacity) { a.counters[SSB_C_OVERFLOW] = 1; continue; }
            const int os = (int)box[k], ns = (int)slot;
            SSB_COPY_FIELD(id); SSB_COPY_FIELD(pos); SSB_COPY_FIELD(born); SSB_COPY_FIELD(cod);
            SSB_COPY_FIELD(sugar); SSB_COPY_FIELD(spice); SSB_COPY_FIELD(sugar_metab); SSB_COPY_FIELD(spice_metab);
            SSB_COPY_FIELD(vision); SSB_COPY_FIELD(movement); SSB_COPY_FIELD(age); SSB_COPY_FIELD(max_age);
            SSB_COPY_FIELD(sex); SSB_COPY_FIELD(fert_age); SSB_COPY_FIELD(infert_age); SSB_COPY_FIELD(fert_factor);
            SSB_COPY_FIELD(start_sugar); SSB_COPY_FIELD(start_spice); SSB_COPY_FIELD(aggression); SSB_COPY_FIELD(lookahead);
            SSB_COPY_FIELD(trade_factor); SSB_COPY_FIELD(mrs); SSB_COPY_FIELD(tags); SSB_COPY_FIELD(tribe);
            SSB_COPY_FIELD(last_moved); SSB_COPY_FIELD(last_sugar); SSB_COPY_FIELD(last_spice);
            SSB_COPY_FIELD(n_neighborhood); SSB_COPY_FIELD(n_valid_moves); SSB_COPY_FIELD(happiness);
            SSB_COPY_FIELD(wealth_happiness); SSB_COPY_FIELD(family_happiness); SSB_COPY_FIELD(trade_volume);
            SSB_COPY_FIELD(trade_price); SSB_COPY_FIELD(last_repro); SSB_COPY_FIELD(children_head); SSB_COPY_FIELD(mates_head);
            SSB_COPY_FIELD(father); SSB_COPY_FIELD(mother);
            c.g.occ[a.pos[ns]] = ns;
            a.order[n_list + j] = ns;
        }
        base += n;
    }
}

__global__ void k_migrate_in_finish(DevCtx c) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long total = 0;
    for (int r = 0; r < c.sh.world; r++)
        if (r != c.sh.rank) total += (long long)c.sh.peer_vals[r * SHARD_PAYLOAD + migrate_word(r, c.sh.rank)];
    const long long n_free = c.a.counters[SSB_C_N_FREE];
    const long long from_stack = total < n_free ? total : n_free;
    c.a.counters[SSB_C_N_FREE] = n_free - from_stack;
    c.a.counters[SSB_C_N_SLOTS] += total - from_stack;
    c.a.counters[SSB_C_N_LIST] += total;
}

// after the second barrier: the emigrants' old slots are free
__global__ void __launch_bounds__(256) k_migrate_recycle(DevCtx c, MigratePlan *mp, long long box_cap) {
    const unsigned long long *mine = c.sh.gather + (long long)c.sh.rank * c.sh.gather_stride;
    const long long n_free = c.a.counters[SSB_C_N_FREE];
    const long long stride = (long long)gridDim.x * blockDim.x, tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    long long base = 0;
    for (int d = 0; d < c.sh.world; d++) {
        if (d == c.sh.rank) continue;
        const long long n = (long long)mp->out_count[d];
        for (long long k = tid; k < n; k += stride) {
            const int32_t s = (int32_t)mine[(long long)d * box_cap + k];
            c.a.free_slots[n_free + base + k] = s;
            c.a.pos[s] = -1;             // the record lives on at its new rank: this slot holds no agent (the sweep's prepare walks the slots)
        }
        base += n;
    }
}

__global__ void k_migrate_recycle_finish(DevCtx c, MigratePlan *mp) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    long long total = 0;
    for (int d = 0; d < c.sh.world; d++) if (d != c.sh.rank) total += (long long)mp->out_count[d];
    c.a.counters[SSB_C_N_FREE] += total;
    c.a.counters[SSB_C_N_MIGRATED] = total;
}

}  // namespace ssb
