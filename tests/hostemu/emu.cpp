/* emu.cpp -- host-compiled build of the engine's mode E transaction source (TEST INFRASTRUCTURE ONLY,
 * see cuda_shim.h).  Exposes one agent step of the exact mode on HOST arrays laid out like the device
 * SoA (include/ssb.h); environment step, compaction and the reductions of a test run come from the
 * CPU oracle, so what this library adds to the `-m "not gpu"` suite is exactly the device rule code. */
#include "cuda_shim.h"

#include "../../sugarscape_b200/csrc/ssb_exact.cuh"

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
    if (4 * (c.max_dx > c.max_dy ? c.max_dx : c.max_dy) > MAXC_EXACT) return -2;
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

}
