// ssb_sweep_core.cuh -- the ORDERED SWEEP scheduler of the mode S agent step (host- and device-compilable part).
//
// Mode S defines a step as "agents act one at a time in ascending priority" (ssb_rng.cuh; the reference's
// random.shuffle(self.agents) + agent loop, sugarscape.py:400-407).  ssb_flow_core.cuh shows that any execution
// which orders every pair of agents with intersecting footprints by priority equals that sequential run.  The
// sweep realises such an execution without queues, dependency counters or commit rounds:
//
//   prepare  the agents are bucketed by the top bits of their priority (two counting passes, no sort) into the
//            list `entries`; every agent leaves a descriptor {priority, range, dilation | list index} on its cell;
//            consecutive groups of buckets (at most 32) are the SLABS of the priority order;
//   sweep    a persistent grid walks the bucketed list by tickets.  An agent runs its whole transaction once all
//            its PREDECESSORS -- the agents with a smaller priority whose footprint intersects its own -- have
//            finished, then sets the DONE bit of its start cell and counts itself in its bucket.  The agents in
//            flight are a thin slice of the priority order (a few buckets), so almost every predecessor of an
//            agent finished long ago.  The WATERMARK -- the number of leading buckets whose agents have ALL
//            finished -- certifies that wholesale: an agent of slab g waits until the watermark has passed every
//            bucket of the slabs < g - 1, and checks individually only its NEAR predecessors, those of the slabs
//            g - 1 and g (about one per agent).  Both waits are for agents of smaller priority only, so the
//            unfinished agent with the globally smallest priority can always run: the sweep cannot deadlock.
//   build    one CTA per 64 x 64 tile stages the descriptors of the tile and its halo in shared memory, with one
//            occupancy bit-plane per slab; one thread per agent scans the two bit-planes of its near slabs over
//            its conflict region and stores the near predecessors it finds as the agent's GATE: up to six
//            (dx, dy) offsets in one 64-bit word.  (More than six: the gate says so and the predecessors go to a
//            128-byte line of column bit masks, as bit dy + halo of word dx + halo.)
//
// Everything in this file is plain C++ so that tests/hostemu compiles the SAME predecessor scan, the SAME gate
// test and the SAME lean transaction with g++ and replays the reference fixtures through adversarial orders.
#pragma once
#include "ssb_flow_core.cuh"

namespace ssb {

// ---- descriptor of the agent standing on a cell at the start of the step (0 = empty) ---------------------
//   high word: list index;  low word: priority (26 bits) << 6 | (range + 1) << 2 | dilation
constexpr int SWEEP_MAX_RANGE = 14;
__host__ __device__ __forceinline__ uint32_t sweep_lo(uint32_t prio, int r, int k) { return (prio << 6) | ((uint32_t)(r + 1) << 2) | (uint32_t)k; }
__host__ __device__ __forceinline__ uint32_t sweep_prio(uint32_t lo) { return lo >> 6; }
__host__ __device__ __forceinline__ int sweep_r(uint32_t lo) { return (int)((lo >> 2) & 15u) - 1; }
__host__ __device__ __forceinline__ int sweep_k(uint32_t lo) { return (int)(lo & 3u); }
__host__ __device__ __forceinline__ unsigned long long sweep_desc(uint32_t index, uint32_t lo) { return ((unsigned long long)index << 32) | lo; }

constexpr int SWEEP_COLS = 32;                      // words of a predecessor mask line (one 128-byte line)
constexpr int SWEEP_MAX_HALO = 15;                  // 2 * halo + 1 <= 31 rows and columns

// ---- slabs: groups of consecutive priority buckets --------------------------------------------------------
// The number of slabs is chosen per engine (sweep_slab_bits): a slab must be several times longer than the slice of the
// priority order that is in flight, or the watermark holds agents back; shorter slabs mean fewer near predecessors.
constexpr int SWEEP_MAX_SLAB_BITS = 5;
constexpr int SWEEP_MAX_SLABS = 1 << SWEEP_MAX_SLAB_BITS;
// slabs for `n` list entries walked by T concurrent transactions: a slab holds at least 8 T entries (measured on S2: with
// 4 T a fifth of all readiness checks fail on the watermark, with 8 T one in a hundred); one slab: every predecessor is near
__host__ __device__ __forceinline__ int sweep_slab_bits(long long n, long long T) {
    int bits = 0;
    while (bits < SWEEP_MAX_SLAB_BITS && (n >> (bits + 1)) >= 8 * T) bits++;
    return bits;
}
// buckets per slab = 1 << sweep_slab_shift; slabs = 1 << min(bucket_bits, slab_bits)
__host__ __device__ __forceinline__ int sweep_slab_shift(int bucket_bits, int slab_bits) { return bucket_bits > slab_bits ? bucket_bits - slab_bits : 0; }
__host__ __device__ __forceinline__ int sweep_n_slabs(int bucket_bits, int slab_bits) { return 1 << (bucket_bits < slab_bits ? bucket_bits : slab_bits); }
// the watermark an agent of bucket b waits for: the first bucket of the slab before its own
__host__ __device__ __forceinline__ int sweep_watermark_needed(int bucket, int slab_shift) {
    const int slab = bucket >> slab_shift;
    return slab > 0 ? (slab - 1) << slab_shift : 0;
}

// ---- gates -------------------------------------------------------------------------------------------------
// up to SWEEP_GATE_PREDS near predecessors, 10 bits each: (dx + halo + 1) | (dy + halo) << 5; 0 ends the list
constexpr int SWEEP_GATE_PREDS = 6;
constexpr unsigned long long SWEEP_GATE_LINE = ~0ull;      // too many: see the agent's mask line
__host__ __device__ __forceinline__ unsigned long long sweep_gate_code(int dx, int dy, int halo) { return (unsigned long long)((dx + halo + 1) | ((dy + halo) << 5)); }

// Conflict tables (filled once per engine on the host, sweep_fill_tables; staged in shared memory by the build kernel):
//   any [ka][ra][adx]          largest |dy| at column offset adx at which an agent of range ra / dilation ka can conflict
//                              with ANY partner (flow_column_reach, clipped to the halo; -1: none)
//   pair[K][ra][rb][adx]       largest |dy| at which ranges ra and rb conflict at column offset adx with K = ka + kb
//                              (flow_conflict is monotone in |dy|; -1: never)
__host__ __device__ __forceinline__ int sweep_any_size(int RB, int KB, int halo) { return (KB + 1) * (RB + 1) * (halo + 1); }
__host__ __device__ __forceinline__ int sweep_pair_size(int RB, int KB, int halo) { return (2 * KB + 1) * (RB + 1) * (RB + 1) * (halo + 1); }
__host__ __device__ __forceinline__ int sweep_tables_bytes(int RB, int KB, int halo) { return (sweep_any_size(RB, KB, halo) + sweep_pair_size(RB, KB, halo) + 15) & ~15; }
inline void sweep_fill_tables(int RB, int KB, int halo, int8_t *any, int8_t *pair) {
    for (int ka = 0; ka <= KB; ka++) for (int ra = 0; ra <= RB; ra++) for (int adx = 0; adx <= halo; adx++) {
        int h = flow_column_reach(adx, ra, ka, RB, KB);
        if (h > halo) h = halo;
        any[(ka * (RB + 1) + ra) * (halo + 1) + adx] = (int8_t)h;
    }
    for (int K = 0; K <= 2 * KB; K++) for (int ra = 0; ra <= RB; ra++) for (int rb = 0; rb <= RB; rb++) for (int adx = 0; adx <= halo; adx++) {
        int h = -1;
        for (int ady = halo; ady >= 0; ady--) if (flow_conflict(adx, ady, ra, rb, K)) { h = ady; break; }
        pair[((K * (RB + 1) + ra) * (RB + 1) + rb) * (halo + 1) + adx] = (int8_t)h;
    }
}

struct SweepTile {
    const uint32_t *cell;                // [sw * sh] low descriptor words, column (x) major like the grid
    const unsigned long long *smask;     // [slabs][sw * 2] plane g: occupancy bits, per staged column (sh <= 128), of the agents of the
                                         // slabs g - 1 and g -- the near slabs of an agent of slab g (sweep_fold_planes)
    const int8_t *any, *pair;            // conflict tables
    int sw, sh, halo;
    int slab_prio_shift;                 // slab = priority >> slab_prio_shift
};

// planes filled per slab -> planes per PAIR of slabs, for word `word` (< sw * 2) of every plane
__host__ __device__ __forceinline__ void sweep_fold_planes(unsigned long long *smask, int n_slabs, int sw, int word) {
    for (int g = n_slabs - 1; g >= 1; g--) smask[(size_t)g * sw * 2 + word] |= smask[(size_t)(g - 1) * sw * 2 + word];
}

// The NEAR predecessors of the agent with low descriptor word `la` on staged cell (lx, ly): the agents of its own slab
// and of the slab before it that have a smaller priority and an intersecting footprint, column by column:
// emit_col(column index dx + halo, mask with bit dy + halo per predecessor) for every column of the conflict region
// (also the empty ones).  RB / KB: the largest range / dilation of any agent in this step.
template <class F>
__host__ __device__ __forceinline__ void sweep_near_masks(const SweepTile &tile, int lx, int ly, uint32_t la, int RB, int KB, F emit_col) {
    const int ra = sweep_r(la), ka = sweep_k(la), halo = tile.halo, hs = halo + 1;
    const uint32_t pa = sweep_prio(la);
    const int g = (int)(pa >> tile.slab_prio_shift);
    const unsigned long long *near = tile.smask + (size_t)g * tile.sw * 2;
    const int8_t *any = tile.any + (ka * (RB + 1) + ra) * hs;
    const int8_t *pair = tile.pair + ((ka * (RB + 1) + ra) * (RB + 1)) * hs;       // + (kb (RB + 1) (RB + 1) + rb) hs + adx
    const int k_stride = (RB + 1) * (RB + 1) * hs;
    for (int dx = -halo; dx <= halo; dx++) {
        const int adx = dx < 0 ? -dx : dx;
        const int h = any[adx];
        uint32_t m = 0;
        if (h >= 0) {
            const int col = lx + dx, y0 = ly - h;
            unsigned long long bits = flow_mask_window(near + col * 2, y0, 2 * h + 1);
            if (dx == 0) bits &= ~(1ull << h);
            const uint32_t *column = tile.cell + col * tile.sh + y0;
            while (bits) {
                const int j = flow_ctz64(bits);
                bits &= bits - 1;
                const uint32_t lb = column[j];
                if (sweep_prio(lb) >= pa) continue;
                const int dy = j - h, ady = dy < 0 ? -dy : dy;
                if (ady <= pair[sweep_k(lb) * k_stride + sweep_r(lb) * hs + adx]) m |= 1u << (dy + halo);
            }
        }
        emit_col(dx + halo, m);
    }
}
// ... one by one: emit(dx, dy).  (The same scan written out once more: routing it through the column masks costs the
// build kernel a millisecond at S2.)
template <class F>
__host__ __device__ __forceinline__ void sweep_near_preds(const SweepTile &tile, int lx, int ly, uint32_t la, int RB, int KB, F emit) {
    const int ra = sweep_r(la), ka = sweep_k(la), halo = tile.halo, hs = halo + 1;
    const uint32_t pa = sweep_prio(la);
    const int g = (int)(pa >> tile.slab_prio_shift);
    const unsigned long long *near = tile.smask + (size_t)g * tile.sw * 2;
    const int8_t *any = tile.any + (ka * (RB + 1) + ra) * hs;
    const int8_t *pair = tile.pair + ((ka * (RB + 1) + ra) * (RB + 1)) * hs;
    const int k_stride = (RB + 1) * (RB + 1) * hs;
    for (int dx = -halo; dx <= halo; dx++) {
        const int adx = dx < 0 ? -dx : dx;
        const int h = any[adx];
        if (h < 0) continue;
        const int col = lx + dx, y0 = ly - h;
        unsigned long long bits = flow_mask_window(near + col * 2, y0, 2 * h + 1);
        if (dx == 0) bits &= ~(1ull << h);
        const uint32_t *column = tile.cell + col * tile.sh + y0;
        while (bits) {
            const int j = flow_ctz64(bits);
            bits &= bits - 1;
            const uint32_t lb = column[j];
            if (sweep_prio(lb) >= pa) continue;
            const int dy = j - h, ady = dy < 0 ? -dy : dy;
            if (ady <= pair[sweep_k(lb) * k_stride + sweep_r(lb) * hs + adx]) emit(dx, dy);
        }
    }
}

// The DONE bitmap: one bit per cell -- "the agent that started the step on this cell has finished" -- in 64-bit words
// along y:  done[x * wpc + (y >> 6)], bit y & 63.
__host__ __device__ __forceinline__ int sweep_done_wpc(int H) { return (H + 63) >> 6; }
__host__ __device__ __forceinline__ size_t sweep_done_words(int W, int H) { return (size_t)W * sweep_done_wpc(H); }
__host__ __device__ __forceinline__ size_t sweep_done_word(int x, int y, int wpc) { return (size_t)x * wpc + (y >> 6); }

// the cell at offset (dx, dy) from (x, y) on the torus (|dx|, |dy| < W, H)
__host__ __device__ __forceinline__ void sweep_offset_cell(int x, int y, int dx, int dy, int W, int H, int *cx, int *cy) {
    int u = x + dx, v = y + dy;
    if (u < 0) u += W; else if (u >= W) u -= W;
    if (v < 0) v += H; else if (v >= H) v -= H;
    *cx = u; *cy = v;
}

// Have the near predecessors listed in `gate` finished?  (x, y): the agent's start cell; ld(w) reads word w of the done
// bitmap past every cache that could hold a stale copy.
template <class LoadDone>
__host__ __device__ __forceinline__ bool sweep_gate_done(unsigned long long gate, LoadDone ld, int x, int y, int W, int H, int halo) {
    const int wpc = sweep_done_wpc(H);
    bool ok = true;
    for (int k = 0; k < SWEEP_GATE_PREDS; k++) {
        const int code = (int)(gate >> (10 * k)) & 1023;
        if ((code & 31) == 0) break;
        int cx, cy;
        sweep_offset_cell(x, y, (code & 31) - 1 - halo, (code >> 5) - halo, W, H, &cx, &cy);
        if (!((ld(sweep_done_word(cx, cy, wpc)) >> (cy & 63)) & 1ull)) ok = false;
    }
    return ok;
}
// ... and those of a mask line (the rare agent with more near predecessors than a gate holds)
template <class LoadDone>
__host__ __device__ inline bool sweep_line_done(const uint32_t *masks, LoadDone ld, int x, int y, int W, int H, int halo) {
    const int wpc = sweep_done_wpc(H);
    bool ok = true;
    for (int col = 0; col <= 2 * halo; col++) {
        uint32_t m = masks[col];
        while (m) {
            const int j = flow_ctz64((unsigned long long)m);
            m &= m - 1;
            int cx, cy;
            sweep_offset_cell(x, y, col - halo, j - halo, W, H, &cx, &cy);
            if (!((ld(sweep_done_word(cx, cy, wpc)) >> (cy & 63)) & 1ull)) ok = false;
        }
    }
    return ok;
}

// number of priority buckets for n agents walked by T concurrent transactions: every bucket must stay well below T
// entries (a thread meets entries T apart, which must lie in different buckets)
__host__ __device__ __forceinline__ int sweep_bucket_bits(long long n, long long T, int id_bits, int max_bits) {
    int bits = 0;
    while (bits < max_bits && bits < id_bits && (n >> bits) * 4 > T) bits++;
    return bits;
}

// =============================================================================================================
// The lean transaction: Agent.doTimestep for rule sets whose transaction stays inside the cross (no tagging /
// trading / reproduction: footprint dilation 0 -- S2's combat_limited rules, the growback / pollution / combat /
// replacement examples), on two sector-sized records, ONE THREAD PER AGENT:
//
//   LeanA [slot]        read + written   holdings, position, tribe, cause of death, by-products for the log
//                                        (what neighbours read through cell.agent)
//   LeanB [list index]  read only        slot, id, range, metabolisms, aggression, lookahead  (coalesced in sweep order)
//
// k_sweep_prepare fills both from the ABI's arrays in slot order, k_lean_unpack writes the mutable fields back
// (and evaluates happiness, agent.py:1522-1530) -- the ABI arrays stay the canonical state between steps.
// =============================================================================================================
enum { LEAN_COD_MASK = 7, LEAN_ACTED = 8, LEAN_RANGED = 16, LEAN_AGED = 32, LEAN_PACKED = 64 };
enum : uint32_t { LEANB_ID_MASK = (1u << 26) - 1u, LEANB_RANGE_SHIFT = 26, LEANB_DIES = 1u << 30, LEANB_SKIP = 1u << 31 };

struct __align__(32) LeanA {
    double sugar, spice;
    int32_t pos;
    int8_t tribe;
    uint8_t hood, valid, state;      // len(movementNeighborhood), len(validMoves), cause of death | LEAN_* flags
    int32_t pad[2];
};
struct __align__(32) LeanB {
    int32_t slot;
    uint32_t idf;                    // id | range << 26 | LEANB_DIES (reaches its maximum age in this step) | LEANB_SKIP
    uint32_t metab;                  // sugar metabolism | spice metabolism << 16 (both clipped at 0)
    int32_t pos;                     // the cell the agent starts the step on
    double aggression, lookahead;
};
static_assert(sizeof(LeanA) == 32 && sizeof(LeanB) == 32, "one 32-byte sector per record");

struct LeanCtx {
    LeanA *a;                  // [capacity] by slot
    LeanB *b;                  // [capacity] by list index
    int2 *pcell;               // [lean_tiled_cells] {occupant slot, sugar} of every cell in 4 x 4 tiles (lean_tiled)
    int32_t ty4;               // tiles per column of tiles: (H + 3) / 4
    int32_t maxc;              // 4 * largest range: candidate cells staged per agent
};

// The cells as the transactions see them: {cell.agent, cell.sugar} in one 8-byte word, 4 x 4 cells per 128-byte line
// and 2 x 2 cells per 32-byte sector, so that a cross of radius r touches about 2 r + 1 sectors instead of one sector
// per candidate and array.  Packed from / unpacked to the ABI's x-major arrays around the sweep (populated tiles only).
__host__ __device__ __forceinline__ uint32_t lean_tiled(int x, int y, int ty4) {
    return ((((uint32_t)(x >> 2) * (uint32_t)ty4 + (uint32_t)(y >> 2)) << 4) | (uint32_t)(((x & 2) << 2) | ((y & 2) << 1) | ((x & 1) << 1) | (y & 1)));
}
__host__ __device__ __forceinline__ long long lean_tiled_cells(int W, int H) { return (long long)((W + 3) / 4) * ((H + 3) / 4) * 16; }

__device__ __forceinline__ LeanA lean_load_a(const LeanA *p) {
#if defined(__CUDA_ARCH__)
    union { LeanA a; uint4 q[2]; } u;
    u.q[0] = __ldcg(reinterpret_cast<const uint4 *>(p));
    u.q[1] = __ldcg(reinterpret_cast<const uint4 *>(p) + 1);
    return u.a;
#else
    return *p;
#endif
}

// The rule sets the lean transaction covers (anything else runs the generic transaction): scalable RNG mode,
// footprint dilation 0, no tag preferences, at most 25 tribes, and a map on which every distance of the cross has
// four distinct cells (environment.py:111-122 merges them when 2 d equals the map side).
__host__ __device__ inline bool lean_supported(const ssb_params_t &p) {
    if (p.rng_mode != SSB_RNG_SCALABLE) return false;
    if (p.flags & (SSB_F_TAGGING | SSB_F_TRADE | SSB_F_REPRO | SSB_F_TAGPREF)) return false;
    if (p.max_tribes > 25 || p.max_sugar_capacity > 32767) return false;
    const int side = p.width < p.height ? p.width : p.height;
    return p.max_agent_range >= 0 && p.max_agent_range <= 7 && 2 * p.max_agent_range + 1 < side;
}

// MODE: which rule families the instantiation can handle (LEAN_M_*); the plain instantiation (0) of rule sets without
// combat, spice and pollution (S2) carries no fp64 pow, no occupant records and no per-tribe table.
enum { LEAN_M_FIGHT = 1, LEAN_M_SPICE = 2, LEAN_M_POLLUTION = 4, LEAN_M_ALL = 7 };
__host__ __device__ inline int lean_mode(const ssb_params_t &p) {
    return (p.flags & (SSB_F_COMBAT | SSB_F_SPICE | SSB_F_POLLUTION)) ? LEAN_M_ALL : 0;
}

// Per-thread staging arrays (shared memory on the device, element i of this thread at [i * stride]).
struct LeanStage {
    int32_t *occ;              // [maxc]: occupants of the candidates in enumeration order
    int16_t *sugar;            // [maxc + 1]: their sugar, then the own cell's (lean_supported: capacities fit 15 bits)
    int stride;
};

// Everything a transaction reads that another transaction of the same step may have written (records A, cell
// contents) is loaded PAST the L1 cache (ld.global.cg = LDG.STRONG.GPU): a predecessor on another SM wrote it to L2
// before it set its done bit, and a stale L1 line is the only thing that could hide that.  This replaces the acquire
// fence on the consumer side, whose CCTL.IVALL (invalidate the whole L1, once per transaction and thread) dominated
// the first version of the sweep kernel.
template <class T> __device__ __forceinline__ T lean_ld(const T *p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
__device__ __forceinline__ void lean_prefetch(const void *p) {
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// Candidate i of Agent.findCellsInRange in the insertion order of Environment.findCardinalCellRanges
// (environment.py:111-122): distance d = i / 4 + 1; within a distance the order is
// W N E S in the interior, N E S W if the west cell wraps (x < d), W E S N if the north cell wraps (y < d),
// E S N W if both.  (2 d < map side: the four cells are distinct, lean_supported.)
__device__ __forceinline__ int lean_candidate(int i, int x, int y, int W, int H) {
    const int d = (i >> 2) + 1, j = i & 3;
    const unsigned order = (x < d) ? ((y < d) ? 0x1Eu : 0x39u) : ((y < d) ? 0x78u : 0xE4u);
    const int dir = (order >> (2 * j)) & 3;                       // 0 W, 1 N, 2 E, 3 S
    int cx = x, cy = y;
    if (dir == 0) { cx = x - d; if (cx < 0) cx += W; }
    else if (dir == 1) { cy = y - d; if (cy < 0) cy += H; }
    else if (dir == 2) { cx = x + d; if (cx >= W) cx -= W; }
    else { cy = y + d; if (cy >= H) cy -= H; }
    return cx * H + cy;
}

// the four candidates at distance d (candidates 4 (d - 1) .. 4 (d - 1) + 3) in that order, as x-major cell indices
// (TILED = false) or as indices of the tiled cell array (TILED = true; hy = ty4 instead of H)
template <bool TILED>
__device__ __forceinline__ void lean_ring(int d, int x, int y, int W, int H, int hy, int out[4]) {
    int xw = x - d, xe = x + d, yn = y - d, ys = y + d;
    if (xw < 0) xw += W;
    if (xe >= W) xe -= W;
    if (yn < 0) yn += H;
    if (ys >= H) ys -= H;
    const int cw = TILED ? (int)lean_tiled(xw, y, hy) : xw * H + y, ce = TILED ? (int)lean_tiled(xe, y, hy) : xe * H + y;
    const int cn = TILED ? (int)lean_tiled(x, yn, hy) : x * H + yn, cs = TILED ? (int)lean_tiled(x, ys, hy) : x * H + ys;
    if (x < d) {
        if (y < d) { out[0] = ce; out[1] = cs; out[2] = cn; out[3] = cw; }
        else { out[0] = cn; out[1] = ce; out[2] = cs; out[3] = cw; }
    } else {
        if (y < d) { out[0] = cw; out[1] = ce; out[2] = cs; out[3] = cn; }
        else { out[0] = cw; out[1] = cn; out[2] = ce; out[3] = cs; }
    }
}

// Ahead of the readiness check: the lines a freshly claimed agent will read -- its record A, its own cell and its cross
// -- are requested from DRAM into L2 (the coherence point: nothing can go stale there), so that the gate check, the
// record and the cross overlap instead of queueing up as dependent round trips.
__device__ __forceinline__ void lean_prefetch_agent(const LeanCtx &L, int slot, int pos, int r, int W, int H) {
    lean_prefetch(L.a + slot);
    const int x = pos / H, y = pos - x * H;
    lean_prefetch(L.pcell + lean_tiled(x, y, L.ty4));
    for (int d = 1; d <= r; d++) {
        int cell[4];
        lean_ring<true>(d, x, y, W, H, L.ty4, cell);
#pragma unroll
        for (int u = 0; u < 4; u++) lean_prefetch(L.pcell + cell[u]);
    }
}

// the range of Agent.findCellsInRange for slot s: min(vision, movement), clipped to the per-axis memoised range
__device__ __forceinline__ int lean_range(const DevCtx &c, int s) {
    return min(min(max(c.a.vision[s], 0), max(c.a.movement[s], 0)), max(c.max_dx, c.max_dy));
}

// k_sweep_prepare for one slot: record A.  live = the slot holds an agent of the list (every agent that is alive
// and on the map is on the list after a compaction).
__device__ __forceinline__ bool lean_pack_a(const DevCtx &c, const LeanCtx &L, int s) {
    const ssb_agents_t &a = c.a;
    const int pos = a.pos[s], cod = a.cod[s];
    const bool live = cod == SSB_ALIVE && pos >= 0;
    LeanA A;
    A.sugar = a.sugar[s]; A.spice = a.spice[s];
    A.pos = pos; A.tribe = (int8_t)a.tribe[s];
    A.hood = 0; A.valid = 0;
    A.state = live ? (uint8_t)LEAN_PACKED : (uint8_t)0;
    A.pad[0] = 0; A.pad[1] = 0;
    L.a[s] = A;
    return live;
}
// ... and record B at its list index
__device__ __forceinline__ void lean_pack_b(const DevCtx &c, const LeanCtx &L, int s, int index, int r) {
    const ssb_agents_t &a = c.a;
    LeanB B;
    B.slot = s;
    const int age = a.age[s], max_age = a.max_age[s];
    B.idf = ((uint32_t)a.id[s] & LEANB_ID_MASK) | ((uint32_t)r << LEANB_RANGE_SHIFT) |
            ((max_age != -1 && age + 1 >= max_age) ? (uint32_t)LEANB_DIES : 0u) | (a.last_moved[s] == c.t ? (uint32_t)LEANB_SKIP : 0u);
    const int sugar_metab = max(a.sugar_metab[s], 0), spice_metab = max(a.spice_metab[s], 0);
    B.metab = (uint32_t)min(sugar_metab, 65535) | ((uint32_t)min(spice_metab, 65535) << 16);
    B.pos = a.pos[s];
    B.aggression = a.aggression[s]; B.lookahead = a.lookahead[s];
    L.b[index] = B;
    // the plain instantiation relies on the rule flags, the record on 16-bit metabolisms: an agent that contradicts them
    // stops the run (overflow code 13)
    if ((lean_mode(c.p) == 0 && (B.aggression > 0 || spice_metab > 0 || a.spice[s] != 0)) || sugar_metab > 65535 || spice_metab > 65535)
        atomicExch((unsigned long long *)&c.a.counters[SSB_C_OVERFLOW], 13ull);
}

// k_lean_unpack for one slot: the mutable fields go back to the ABI's arrays; happiness of the agents that are
// still alive after aging (updateHappiness, agent.py:1522-1530: family 0 in mode S, health 1, wealth erf).
__device__ __forceinline__ void lean_unpack_slot(const DevCtx &c, const LeanCtx &L, int s) {
    const LeanA A = L.a[s];
    if (!(A.state & LEAN_PACKED)) return;
    const ssb_agents_t &a = c.a;
    const int cod = A.state & LEAN_COD_MASK;
    const double sugar0 = a.sugar[s], spice0 = a.spice[s];     // holdings at the start of the step
    a.sugar[s] = A.sugar; a.spice[s] = A.spice; a.pos[s] = cod == SSB_ALIVE ? A.pos : -1; a.cod[s] = cod;
    if (!(A.state & LEAN_ACTED)) return;
    // nobody touches the holdings of a living agent before its own transaction in these rule sets (combat kills):
    // lastSugar / lastSpice (agent.py:578-579) are the holdings the step started with
    a.last_moved[s] = c.t;
    a.last_sugar[s] = sugar0; a.last_spice[s] = spice0;
    if (A.state & LEAN_RANGED) { a.n_neighborhood[s] = A.hood; a.n_valid_moves[s] = A.valid; }
    if (!(A.state & LEAN_AGED)) return;
    a.age[s] = a.age[s] + 1;
    if (cod != SSB_ALIVE) return;
    const double wealth_h = erf((A.sugar + A.spice) - c.prev_mean_wealth);
    const double family_h = erf(0.0);
    a.wealth_happiness[s] = wealth_h;
    a.family_happiness[s] = family_h;
    a.happiness[s] = family_h + 1.0 + wealth_h;
}

// Agent.doTimestep (agent.py:568-598) for a rule set with footprint dilation 0: findBestCell with the shuffled
// tie-break (1395-1443, 1479-1490), doCombat (274-294), gotoCell, collectResourcesAtCell (249-259), doMetabolism
// (485-498) with the pollution terms (cell.py:27-45), starvation, doAging (266-272).  Mirrors
// transaction_move_phase / transaction_interact_phase of ssb_rules.cuh statement by statement.
// warpm: the lanes of the warp that enter together (they must be converged at the call).  The lanes are independent
// agents, but they run the same loops with the same trip counts (the sweep groups agents by range), so the function
// re-converges them after every data-dependent stretch -- otherwise a warp drifts apart into a handful of sub-warps
// that each pay the full instruction stream.
// preloaded: the agent's record A, loaded by the caller ahead of the readiness check -- only for instantiations in which
// nobody but the agent itself writes its record (MODE without LEAN_M_FIGHT); otherwise null.
// The plain instantiation (MODE 0) only needs to know WHETHER a candidate is occupied: it stages one 16-bit word per
// candidate (sugar | occupied << 15) and never touches st.occ.
template <int MODE>
__device__ inline void lean_transaction(const DevCtx &c, const LeanCtx &L, const LeanB &B, uint64_t key, const LeanStage &st, unsigned warpm,
                                        const LeanA *preloaded = nullptr) {
    constexpr bool SLIM = MODE == 0;
    const int s = B.slot;
    LeanA A = (preloaded && !(MODE & LEAN_M_FIGHT)) ? *preloaded : lean_load_a(L.a + s);
    const bool acting = (A.state & LEAN_COD_MASK) == SSB_ALIVE && A.sugar >= 0 && A.spice >= 0 && !(B.idf & LEANB_SKIP);
    const unsigned actm = __ballot_sync(warpm, acting);
    if (!acting) return;
    const int W = c.p.width, H = c.p.height, pos = A.pos, x = pos / H, y = pos - x * H;
    const int r = (int)((B.idf >> LEANB_RANGE_SHIFT) & 15u), n = 4 * r, S = st.stride;
    const uint32_t id = B.idf & LEANB_ID_MASK;
    const bool polluted = (MODE & LEAN_M_POLLUTION) && (c.p.flags & SSB_F_POLLUTION) != 0, spicy = (MODE & LEAN_M_SPICE) && (c.p.flags & SSB_F_SPICE) != 0;
    const double aggression = (MODE & LEAN_M_FIGHT) ? fmax(B.aggression, 0.0) : 0.0;
    const bool fighter = (MODE & LEAN_M_FIGHT) && aggression > 0;
    // the cross (and the own cell, in case the agent stays): two distances = eight independent 8-byte loads in flight
    const int2 *pcell = L.pcell;
    const uint32_t tpos = lean_tiled(x, y, L.ty4);
    const int2 own = lean_ld(pcell + tpos);                     // (in flight together with the first rings)
    for (int d = 1; d <= r; d += 2) {
        int cell[8];
        int2 v[8];
        lean_ring<true>(d, x, y, W, H, L.ty4, cell);
        const bool two = d + 1 <= r;
        if (two) lean_ring<true>(d + 1, x, y, W, H, L.ty4, cell + 4);
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (u < 4 || two) v[u] = lean_ld(pcell + cell[u]);
#pragma unroll
        for (int u = 0; u < 8; u++)
            if (u < 4 || two) {
                if (SLIM) st.sugar[(4 * (d - 1) + u) * S] = (int16_t)(v[u].y | (v[u].x >= 0 ? 0x8000 : 0));
                else { st.occ[(4 * (d - 1) + u) * S] = v[u].x; st.sugar[(4 * (d - 1) + u) * S] = (int16_t)v[u].y; }
            }
    }
    st.sugar[n * S] = (int16_t)own.y;
    __syncwarp(actm);
    // findWelfare's per-agent part (agent.py:1170-1201, no tag preferences)
    Welfare w;
    const double sm = (double)(B.metab & 65535u), pm = (double)(B.metab >> 16);
    {
        w.sugar = A.sugar; w.spice = A.spice;
        w.s_look = sm * B.lookahead; w.p_look = pm * B.lookahead;
        const double total = sm + pm;
        w.e_s = 0; w.e_p = 0;
        if (pm == 0) { if (sm != 0) w.e_s = 1.0; }          // x / x == 1 and 0 / x == 0 exactly: no division (its zero-numerator
        else if (sm == 0) w.e_p = 1.0;                      // case takes the slow path of the software fp64 division)
        else { w.e_s = sm / total; w.e_p = pm / total; }
    }
    const double my_wealth = A.sugar + A.spice, loot = c.p.max_combat_loot;
    int b_cell = pos, b_dist = 0, b_cs = 0, b_cp = 0, b_occ = -1, hood = 0, m = 0;
    unsigned b_ties = 0;                       // candidates j of ring b_dist that share the best (welfare, distance)
    double b_welfare = -1.0, b_pl = 0.0;
    bool b_valid = false;
    {   // (no `if (n > 0)` around this block: every lane must reach the __syncwarp()s; with n == 0 the loops do nothing)
        // findRetaliatorsInVision (agent.py:1088-1098): the richest visible agent of every tribe
        double retaliator[(MODE & LEAN_M_FIGHT) ? 27 : 1];
        if (fighter) {
            for (int i = 0; i < n; i++) { const int o = st.occ[i * S]; if (o >= 0) lean_prefetch(L.a + o); }
            for (int i = 0; i < 27; i++) retaliator[i] = 0.0;
            for (int i = 0; i < n; i++) {
                const int o = st.occ[i * S];
                if (o < 0) continue;
                const LeanA O = lean_load_a(L.a + o);
                const double ow = fmax(O.sugar + O.spice, 0.0);
                const int slot = min((int)O.tribe + 1, 26);
                if (ow > retaliator[slot]) retaliator[slot] = ow;
            }
        }
        // rankCellsInRange (agent.py:1395-1443): best welfare, then the smallest distance, then the position in the
        // shuffled candidate list.  Candidates that tie in (welfare, distance) lie on ONE ring of four cells, so the
        // shuffle is only consulted for those (below) -- and not at all when the best candidate is unique.
        for (int d = 1; d <= r; d++) {
            int ring[4];
            lean_ring<false>(d, x, y, W, H, H, ring);
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int i = 4 * (d - 1) + j, cell = ring[j];
                const int staged = st.sugar[i * S];
                const int o = SLIM ? -1 : st.occ[i * S], cs = SLIM ? (staged & 0x7fff) : staged;
                const bool occupied = SLIM ? staged < 0 : o >= 0;
                int otribe = -1;
                double ps = 0, pp = 0;
                if (occupied) {
                    // findNeighborhood (agent.py:1052-1065): an agent found through a cell's occupant pointer is alive
                    // (every death of this rule family clears the pointer in the same transaction)
                    hood++;
                    if (!fighter) continue;
                    const LeanA O = lean_load_a(L.a + o);
                    otribe = O.tribe;
                    // isNeighborValidPrey (agent.py:1309-1314)
                    if (!(A.tribe != otribe && my_wealth >= O.sugar + O.spice)) continue;
                    ps = O.sugar; pp = O.spice;
                }
                const int cp = spicy ? lean_ld(c.g.spice + cell) : 0;
                const double pl = polluted ? lean_ld(c.g.pollution + cell) : 0.0;
                double sugar_reward = (double)cs, spice_reward = (double)cp;
                if (fighter) { sugar_reward += aggression * fmin(loot, ps); spice_reward += aggression * fmin(loot, pp); }   // (+ 0.0 otherwise)
                if (polluted) { sugar_reward /= 1 + pl; spice_reward /= 1 + pl; }                                            // (/ 1.0 otherwise)
                double welfare;
                if (MODE & LEAN_M_SPICE) welfare = w.eval(sugar_reward, spice_reward);
                else {      // no spice anywhere: the exponents are (1, 0) -- or (0, 0) without a sugar metabolism -- and x ** 1 == x, x ** 0 == 1
                    double ts = (w.sugar + sugar_reward) - w.s_look;
                    if (ts < 0) ts = 0;
                    welfare = w.e_s != 0 ? ts : 1.0;
                }
                if ((MODE & LEAN_M_FIGHT) && occupied && retaliator[min(otribe + 1, 26)] > my_wealth + welfare) continue;
                m++;
                if (!b_valid || welfare > b_welfare) { b_welfare = welfare; b_dist = d; b_ties = 1u << j; b_valid = true; }    // (rings come in ascending distance)
                else if (welfare == b_welfare && d == b_dist) b_ties |= 1u << j;
            }
        }
        if (n > 0 && m == 0) m = 1;
    }
    __syncwarp(actm);
    if (b_valid) {
        int jw = __ffs(b_ties) - 1;
        if (b_ties & (b_ties - 1u)) {
            // random.shuffle(cellsInRange), agent.py:1400-1401, in mode S the keyed shuffle: candidate i stands where its
            // substream word i stands among all words (ties by i) -- of the tied candidates of this ring the one with the
            // smallest word comes first
            RngStream rng{key, id, 0u, 0u, nullptr};
            rng.at(SITE_MOVE);
            const int base = 4 * (b_dist - 1);
            uint32_t best_key = 0;
            bool have = false;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                if (!((b_ties >> j) & 1u)) continue;
                const uint32_t kj = rng.word((uint32_t)(base + j));
                if (!have || kj < best_key) { best_key = kj; jw = j; have = true; }
            }
        }
        int ring[4];
        lean_ring<false>(b_dist, x, y, W, H, H, ring);
        const int i = 4 * (b_dist - 1) + jw;
        b_cell = ring[jw];
        b_cs = st.sugar[i * S] & 0x7fff; b_occ = SLIM ? -1 : st.occ[i * S];        // (a chosen cell is empty in the plain instantiation)
        b_cp = spicy ? lean_ld(c.g.spice + b_cell) : 0;
        b_pl = polluted ? lean_ld(c.g.pollution + b_cell) : 0.0;
    }
    __syncwarp(actm);
    uint8_t state = (uint8_t)(LEAN_PACKED | LEAN_ACTED | (n > 0 ? LEAN_RANGED : 0));
    double sugar = A.sugar, spice = A.spice;
    const int best = b_valid ? b_cell : pos;
    if (!b_valid) {          // stays on its own cell: its contents were not among the candidates
        b_cs = st.sugar[n * S];
        b_cp = spicy ? lean_ld(c.g.spice + pos) : 0;
        b_pl = polluted ? lean_ld(c.g.pollution + pos) : 0.0;
    }
    // doCombat (agent.py:274-294) then gotoCell (agent.py:1213-1219)
    if (fighter && b_valid && b_occ >= 0 && b_occ != s) {
        LeanA P = lean_load_a(L.a + b_occ);
        const double sl = fmin(loot, P.sugar), pl = fmin(loot, P.spice);
        sugar += sl; spice += pl;
        P.sugar -= sl; P.spice -= pl;
        P.state = (uint8_t)((P.state & ~LEAN_COD_MASK) | SSB_COMBAT);       // doDeath("combat"): the killer takes the cell below
        L.a[b_occ] = P;
    }
    const int bx = best / H;
    int2 *pbest = L.pcell + lean_tiled(bx, best - bx * H, L.ty4);
    L.pcell[tpos].x = -1;
    *pbest = make_int2(s, 0);                                    // the agent stands there; the cell's sugar is collected below
    // collectResourcesAtCell (agent.py:249-259) + production pollution (cell.py:32-45)
    sugar += b_cs;
    spice += b_cp;
    const bool polluting = c.p.pollution_start <= c.t && c.t <= c.p.pollution_end;
    double cell_poll = b_pl;
    if (polluting) {
        cell_poll += c.p.sugar_prod_pollution * b_cs;
        cell_poll += c.p.spice_prod_pollution * b_cp;
    }
    if (spicy) c.g.spice[best] = 0;
    // doMetabolism (agent.py:485-498) + consumption pollution (cell.py:27-40)
    sugar -= sm;
    spice -= pm;
    if (polluting) {
        cell_poll += c.p.sugar_cons_pollution * sm;
        cell_poll += c.p.spice_cons_pollution * pm;
        c.g.pollution[best] = cell_poll;
    }
    A.sugar = sugar; A.spice = spice; A.pos = best;
    A.hood = (uint8_t)(hood + 1); A.valid = (uint8_t)m;
    if (sugar < 0 || spice < 0 || (sugar <= 0 && sm > 0) || (spice <= 0 && pm > 0)) {
        state |= SSB_STARVATION;                                              // doDeath("starvation") (agent.py:296-313)
        pbest->x = -1;
    } else {
        // doAging (agent.py:266-272); isAlive() holds here: both holdings are positive or their metabolism is 0
        if (sugar >= 0 && spice >= 0) {
            state |= LEAN_AGED;
            if (B.idf & LEANB_DIES) {
                state |= SSB_AGING;
                pbest->x = -1;
            }
        }
    }
    A.state = state;
    L.a[s] = A;
}

}  // namespace ssb
