// C ABI of the batched drone environment (include/aac_env.h): handle management, map / scenario
// upload and kernel launches.  No torch types, no exceptions across the boundary, no synchronisation
// except where the header says so.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "aac_kernels.cuh"

using namespace aac;

struct AacEnv {
    AacConfig cfg;
    int device = 0;
    int group = 0, threads = 0;   // envs per warp, threads per CTA
    WarpLayout wl{};
    CtaLayout cl{};
    MapDev *d_maps = nullptr;
    int n_maps = 0, n_staged = 1;   // maps in the table / staged per CTA in shared memory
    std::vector<int> h_gx, h_gy;   // grid sizes of the installed maps (table validation)
    double h_ox = 0.0, h_oy = 0.0; // local-frame origin of map 0 (clouds are configured in global metres)
    float4 *d_ray = nullptr;
    std::vector<float4> h_ray;   // host copy: the walk constants are derived from it once the cell size is known (aac_set_maps)
    DdaRay *d_dda = nullptr;
    uint4 *d_walk = nullptr;     // the walk table of the grid radar (aac_radar.cuh)
    uint16_t *d_bank_cells = nullptr;
    uint8_t *d_bank_w = nullptr;
    int32_t *d_bank_map = nullptr;
    int n_scen = 0;
    const float *rtab = nullptr, *rtab_min = nullptr;   // caller-owned radar table (aac_set_radar_table)
    const int16_t *rtab_hit = nullptr;
    const uint32_t *rtab_minr = nullptr;
    OdDev *d_od = nullptr;       // origin / destination tables (one per map) and the buffers they point into
    std::vector<void *> od_bufs;
    float *d_actions = nullptr;  // staging for aac_step_host
    double *d_stats = nullptr;
    int *d_work = nullptr;       // ping-pong group counters of the persistent kernel
    bool od_plan = false;        // an origin / destination table has pools only: paths are searched per episode on the device
    uint8_t *d_plan_scratch = nullptr;   // per-warp scratch of that search
    unsigned long long *d_flags = nullptr;   // [groups] phased launch: (epoch << 32) | terminated envs of the group
    int epoch = 0;               // phased launches so far
    int sms = 0;
    int grid = 0;                // resident CTAs per SM of the kernel, queried at the first launch
    float cell = 0.0f;           // cell size of the maps (all maps of a handle share it)
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};   // aac_step_host: chunks rotate over these streams
    cudaEvent_t pipe_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t pair_launches[1 + 16] = {0};                 // launches per group-counter pair (0 = whole range)
    AacState st{};
    bool bound = false;
    int64_t launches = 0;
};

static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
static int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return AAC_ERR_CUDA;
}
#define CU(call)                                             \
    do {                                                     \
        cudaError_t e_ = (call);                             \
        if (e_ != cudaSuccess) return cuda_fail(e_, #call);  \
    } while (0)

extern "C" const char *aac_last_error(void) { return g_err; }

extern "C" int aac_own_dim(int32_t variant, int32_t n_agents) { return own_dim(variant, n_agents); }

extern "C" int64_t aac_launch_count(const AacEnv *env) { return env ? env->launches : 0; }

extern "C" int aac_create(const AacConfig *cfg, AacEnv **out) {
    if (!cfg || !out) return fail(AAC_ERR_ARG, "aac_create: null argument");
    if (cfg->abi_version != AAC_ABI_VERSION) return fail(AAC_ERR_ARG, "aac_create: abi_version mismatch");
    if (cfg->variant != AAC_VARIANT_ATT && cfg->variant != AAC_VARIANT_V2 && cfg->variant != AAC_VARIANT_MM)
        return fail(AAC_ERR_ARG, "aac_create: unknown variant");
    if (cfg->variant == AAC_VARIANT_MM && (cfg->out_flags & (AAC_OUT_NBR6 | AAC_OUT_TCPA_PAIR)))
        return fail(AAC_ERR_ARG, "aac_create: the multipleMap variant has no neighbour outputs");
    if (cfg->eval_by_step && cfg->variant != AAC_VARIANT_V2) return fail(AAC_ERR_ARG, "aac_create: eval_by_step is a mode of the tdCPA_forV2 variant");
    if (cfg->autoreset_launches < 0 || cfg->autoreset_launches > 3) return fail(AAC_ERR_ARG, "aac_create: autoreset_launches must be 0, 1, 2 or 3");
    if (cfg->radar_targets || cfg->n_nbr_obs || cfg->n_clouds) {   // the later fork's sensor classes
        if (cfg->variant != AAC_VARIANT_V2 || cfg->eval_by_step) return fail(AAC_ERR_ARG, "aac_create: the sensor classes extend the tdCPA_forV2 variant (training mode)");
        if (cfg->radar_targets & ~0xF) return fail(AAC_ERR_ARG, "aac_create: unknown radar target class");
        if (cfg->n_clouds < 0 || cfg->n_clouds > AAC_MAX_CLOUDS) return fail(AAC_ERR_ARG, "aac_create: at most AAC_MAX_CLOUDS clouds");
        if ((cfg->radar_targets & AAC_TARGET_CLOUDS) && cfg->n_clouds < 1) return fail(AAC_ERR_ARG, "aac_create: AAC_TARGET_CLOUDS without clouds");
        if (cfg->n_nbr_obs < 0 || cfg->n_nbr_obs > cfg->n_agents - 1) return fail(AAC_ERR_ARG, "aac_create: n_nbr_obs must lie in 0 .. n_agents - 1");
        if (cfg->n_nbr_obs && (cfg->out_flags & (AAC_OUT_NBR6 | AAC_OUT_TCPA_PAIR))) return fail(AAC_ERR_ARG, "aac_create: the per-pair optional outputs cover every neighbour: not with n_nbr_obs");
        for (int c = 0; c < cfg->n_clouds; ++c)
            if (!(cfg->clouds[c][4] > 0) || !(cfg->clouds[c][5] >= 0)) return fail(AAC_ERR_ARG, "aac_create: a cloud needs a positive radius and a non-negative speed");
    }
    if (cfg->n_envs < 1) return fail(AAC_ERR_ARG, "aac_create: n_envs < 1");
    if (cfg->n_agents < 1 || cfg->n_agents > AAC_MAX_AGENTS) return fail(AAC_ERR_ARG, "aac_create: n_agents out of range");
    if (cfg->n_rays < 1 || cfg->n_rays > AAC_MAX_RAYS || 360 % cfg->n_rays) return fail(AAC_ERR_ARG, "aac_create: n_rays must divide 360");
    if (cfg->w_max < 2 || cfg->w_max > AAC_MAX_W || cfg->w_max % 8) return fail(AAC_ERR_ARG, "aac_create: w_max must be a multiple of 8 in [8, 64]");
    if (!(cfg->dt > 0) || !(cfg->vmax > 0) || !(cfg->prot > 0) || !(cfg->ray_len > 0)) return fail(AAC_ERR_ARG, "aac_create: non-positive physical constant");
    AacEnv *env = new (std::nothrow) AacEnv();
    if (!env) return fail(AAC_ERR_STATE, "aac_create: out of host memory");
    env->cfg = *cfg;
    CU(cudaGetDevice(&env->device));
    // a warp owns G = 32 / N whole envs; a CTA is a handful of independent warps sharing the map
    CU(cudaDeviceGetAttribute(&env->sms, cudaDevAttrMultiProcessorCount, env->device));
    env->group = cfg->tile_envs > 0 ? cfg->tile_envs : (32 / cfg->n_agents > 0 ? 32 / cfg->n_agents : 1);
    if (cfg->tile_envs <= 0) {
        // small batches are latency-bound (one group's pipeline is a ~60 us dependency chain): give each warp fewer envs
        // so that there are about 12 groups per SM, even though lanes idle in the drone-per-lane phases (C2, 4096 envs x 3
        // drones on 148 SMs: 2 envs per warp measured best: 0.039 ms against 0.042 with 1 or 3, 0.046 with 5)
        const int want = cfg->n_envs / (12 * (env->sms > 0 ? env->sms : 1));
        if (want < env->group) env->group = want < 1 ? 1 : want;
    }
    if (env->group * cfg->n_agents > 32) { delete env; return fail(AAC_ERR_ARG, "aac_create: tile_envs * n_agents must not exceed 32"); }
    if (cfg->block_threads > 0) {
        env->threads = cfg->block_threads;
    } else {
        // small batches: fewer warps per CTA so that every SM gets work (a 4096-env x 3-drone batch is only 410 groups)
        const int groups = (cfg->n_envs + env->group - 1) / env->group;
        int wpc = (groups + env->sms - 1) / (env->sms > 0 ? env->sms : 1);
        wpc = wpc < 1 ? 1 : (wpc > MAX_THREADS / 32 ? MAX_THREADS / 32 : wpc);
        env->threads = 32 * wpc;
    }
    if (env->threads > MAX_THREADS || env->threads % 32) { delete env; return fail(AAC_ERR_ARG, "aac_create: block_threads must be a multiple of 32, <= 256"); }
    env->wl = make_warp_layout(cfg->variant, cfg->n_agents, cfg->out_flags);
    const int optin = max_smem_optin();
    while (true) {
        env->cl = make_cta_layout(env->wl, cfg->n_rays, env->threads / 32);
        if (optin <= 0 || (int)env->cl.total <= optin || env->threads == 32) break;
        env->threads -= 32;
    }
    if (optin > 0 && (int)env->cl.total > optin) { delete env; return fail(AAC_ERR_ARG, "aac_create: one warp's envs do not fit in shared memory"); }
    cudaError_t e = upload_constants();
    if (e != cudaSuccess) { delete env; return cuda_fail(e, "upload_constants"); }
    // ray table: direction k*(360/R) degrees (ATT:1058-1066).  The reference adds 15*cos to a
    // coordinate of a few hundred metres, so components below ~1e-14 vanish: axis rays are exact.
    std::vector<float4> rays(cfg->n_rays);
    const int step_deg = 360 / cfg->n_rays;
    for (int k = 0; k < cfg->n_rays; ++k) {
        const double rad = (double)(k * step_deg) * (M_PI / 180.0);
        double c = cos(rad), s = sin(rad);
        if (fabs(c) < 1e-12) c = 0.0;
        if (fabs(s) < 1e-12) s = 0.0;
        const float dx = (float)(cfg->ray_len * c), dy = (float)(cfg->ray_len * s);
        rays[k] = make_float4(dx, dy, dx != 0.0f ? 1.0f / dx : INFINITY, dy != 0.0f ? 1.0f / dy : INFINITY);
    }
    // an even fan is made exactly antisymmetric: the kernel casts ray k and ray k + R/2 from one set of
    // slab products (cos(x + pi) and -cos(x) agree to 1e-16 in the reference's float64)
    if (cfg->n_rays % 2 == 0)
        for (int k = 0; k < cfg->n_rays / 2; ++k) {
            const float4 r = rays[k];
            rays[k + cfg->n_rays / 2] = make_float4(-r.x, -r.y, r.x != 0.0f ? -r.z : r.z, r.y != 0.0f ? -r.w : r.w);
        }
    CU(cudaMalloc(&env->d_ray, sizeof(float4) * cfg->n_rays));
    CU(cudaMemcpy(env->d_ray, rays.data(), sizeof(float4) * cfg->n_rays, cudaMemcpyHostToDevice));
    env->h_ray = rays;
    CU(cudaMalloc(&env->d_dda, sizeof(DdaRay) * cfg->n_rays));
    CU(cudaMemset(env->d_dda, 0, sizeof(DdaRay) * cfg->n_rays));
    std::vector<uint4> walk(WALK_BYTES / 16);
    make_walk_table(walk.data());
    CU(cudaMalloc(&env->d_walk, WALK_BYTES));
    CU(cudaMemcpy(env->d_walk, walk.data(), WALK_BYTES, cudaMemcpyHostToDevice));
    CU(cudaMalloc(&env->d_work, 4 * 17 * sizeof(int)));   // ping-pong pairs of (step, reset) counters: one for whole-range launches, 16 for pipeline chunks
    CU(cudaMemset(env->d_work, 0, 4 * 17 * sizeof(int)));
    CU(cudaMalloc(&env->d_stats, sizeof(double) * AAC_N_STATS));
    CU(cudaMemset(env->d_stats, 0, sizeof(double) * AAC_N_STATS));
    *out = env;
    return 0;
}

extern "C" void aac_destroy(AacEnv *env) {
    if (!env) return;
    cudaFree(env->d_maps);
    cudaFree(env->d_ray);
    cudaFree(env->d_dda);
    cudaFree(env->d_walk);
    cudaFree(env->d_bank_cells);
    cudaFree(env->d_bank_w);
    cudaFree(env->d_bank_map);
    cudaFree(env->d_actions);
    cudaFree(env->d_stats);
    cudaFree(env->d_work);
    cudaFree(env->d_flags);
    cudaFree(env->d_plan_scratch);
    cudaFree(env->d_od);
    for (void *b : env->od_bufs) cudaFree(b);
    for (auto &s : env->pipe) if (s) cudaStreamDestroy(s);
    for (auto &e : env->pipe_ev) if (e) cudaEventDestroy(e);
    delete env;
}

extern "C" int aac_set_maps(AacEnv *env, const AacMapDesc *maps, const uint8_t *occ, int32_t n_maps) {
    if (env) { env->rtab = env->rtab_min = nullptr; env->rtab_hit = nullptr; env->rtab_minr = nullptr; }   // a radar table belongs to the maps it was built for
    if (!env || !maps || !occ || n_maps < 1) return fail(AAC_ERR_ARG, "aac_set_maps: bad argument");
    if (env->cfg.variant != AAC_VARIANT_MM && n_maps != 1) return fail(AAC_ERR_ARG, "aac_set_maps: this variant uses exactly one map");
    if (n_maps > 255) return fail(AAC_ERR_ARG, "aac_set_maps: at most 255 maps");
    std::vector<MapDev> host(n_maps);
    for (int m = 0; m < n_maps; ++m) {
        const AacMapDesc &d = maps[m];
        MapDev &o = host[m];
        memset(&o, 0, sizeof(o));
        if (d.gx < 1 || d.gy < 1 || d.gx > 255 || d.gy > 255 || d.gx * d.gy > AAC_MAP_STRIDE ||
            (d.gx + 2 * MAP_PAD) * (d.gy + 2 * MAP_PAD) + 32 > MAP_WORDS * 32)
            return fail(AAC_ERR_ARG, "aac_set_maps: grid too large");
        if (!(d.cell > 0) || env->cfg.ray_len > 1.5f * d.cell || 2.0f * env->cfg.prot > d.cell)
            return fail(AAC_ERR_ARG, "aac_set_maps: need ray_len <= 1.5*cell and 2*prot <= cell");
        o.gx = d.gx; o.gy = d.gy; o.pgx = d.gx + 2 * MAP_PAD; o.pgy = d.gy + 2 * MAP_PAD;
        o.hx = 0.5f * (d.bound[1] - d.bound[0]); o.hy = 0.5f * (d.bound[3] - d.bound[2]);
        o.ox = d.origin_x; o.oy = d.origin_y;
        o.ex0 = (d.x0c - 0.5f * d.cell) - d.origin_x; o.ey0 = (d.y0c - 0.5f * d.cell) - d.origin_y;
        o.xmin_g = d.bound[0]; o.ymin_g = d.bound[2];
        if (m > 0 && d.cell != maps[0].cell) return fail(AAC_ERR_ARG, "aac_set_maps: every map must use the same cell size");
        o.cell = d.cell; o.inv_cell = 1.0f / d.cell;
        o.ihx = 1.0f / o.hx; o.ihy = 1.0f / o.hy;
        for (int ix = 0; ix < d.gx; ++ix)
            for (int iy = 0; iy < d.gy; ++iy)
                if (occ[(size_t)m * AAC_MAP_STRIDE + ix * d.gy + iy]) {
                    const int b = (ix + MAP_PAD) * o.pgy + iy + MAP_PAD;
                    o.bits[b >> 5] |= 1u << (b & 31);
                }
    }
    if (env->d_maps) { CU(cudaFree(env->d_maps)); env->d_maps = nullptr; }
    CU(cudaMalloc(&env->d_maps, sizeof(MapDev) * n_maps));
    CU(cudaMemcpy(env->d_maps, host.data(), sizeof(MapDev) * n_maps, cudaMemcpyHostToDevice));
    env->n_maps = n_maps;
    env->cell = maps[0].cell;
    // multipleMap: the whole table goes to shared memory with the prologue's bulk copy when it is small (the reference's 14 maps: 4.5 KB)
    env->n_staged = (env->cfg.variant == AAC_VARIANT_MM && n_maps <= MAX_STAGED_MAPS) ? n_maps : 1;
    {
        const int optin = max_smem_optin();
        CtaLayout cl = make_cta_layout(env->wl, env->cfg.n_rays, env->threads / 32, env->n_staged);
        if (optin > 0 && (int)cl.total > optin) { env->n_staged = 1; cl = make_cta_layout(env->wl, env->cfg.n_rays, env->threads / 32, 1); }
        env->cl = cl;
        env->grid = 0;   // the residency is queried again with the new shared-memory size
    }
    env->h_ox = maps[0].origin_x; env->h_oy = maps[0].origin_y;
    env->h_gx.resize(n_maps); env->h_gy.resize(n_maps);
    for (int m = 0; m < n_maps; ++m) { env->h_gx[m] = maps[m].gx; env->h_gy[m] = maps[m].gy; }
    // the rays as constants of the cell walk (aac_radar.cuh): they depend on the cell size
    std::vector<DdaRay> dda(env->h_ray.size());
    for (size_t k = 0; k < dda.size(); ++k) dda[k] = make_dda_ray(env->h_ray[k].x, env->h_ray[k].y, env->cell);
    CU(cudaMemcpy(env->d_dda, dda.data(), sizeof(DdaRay) * dda.size(), cudaMemcpyHostToDevice));
    return 0;
}

extern "C" int aac_set_bank(AacEnv *env, const AacBank *bank) {
    if (!env || !bank || bank->n_scenarios < 1 || !bank->cells || !bank->w) return fail(AAC_ERR_ARG, "aac_set_bank: bad argument");
    const size_t S = bank->n_scenarios, N = env->cfg.n_agents, W = env->cfg.w_max;
    const size_t w_cap = env->cfg.variant == AAC_VARIANT_MM ? 31 : W;   // multipleMap keeps the remaining waypoints in a 32-bit mask ((1 << w) - 2)
    for (size_t k = 0; k < S * N; ++k)
        if (bank->w[k] < 2 || bank->w[k] > w_cap) return fail(AAC_ERR_ARG, "aac_set_bank: reference line needs 2..w_max (multipleMap: 31) vertices");
    if (bank->map_id)
        for (size_t k = 0; k < S; ++k)
            if (bank->map_id[k] < 0 || bank->map_id[k] >= (env->n_maps ? env->n_maps : 1)) return fail(AAC_ERR_ARG, "aac_set_bank: map_id out of range");
    cudaFree(env->d_bank_cells); cudaFree(env->d_bank_w); cudaFree(env->d_bank_map);
    env->d_bank_cells = nullptr; env->d_bank_w = nullptr; env->d_bank_map = nullptr;
    CU(cudaMalloc(&env->d_bank_cells, S * N * W * sizeof(uint16_t)));
    CU(cudaMemcpy(env->d_bank_cells, bank->cells, S * N * W * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&env->d_bank_w, S * N));
    CU(cudaMemcpy(env->d_bank_w, bank->w, S * N, cudaMemcpyHostToDevice));
    if (bank->map_id) {
        CU(cudaMalloc(&env->d_bank_map, S * sizeof(int32_t)));
        CU(cudaMemcpy(env->d_bank_map, bank->map_id, S * sizeof(int32_t), cudaMemcpyHostToDevice));
    }
    env->n_scen = (int)S;
    return 0;
}

extern "C" int aac_set_od_tables(AacEnv *env, const AacOdTable *tables, int32_t n_maps) {
    if (!env || !tables || n_maps < 1) return fail(AAC_ERR_ARG, "aac_set_od_tables: bad argument");
    if (n_maps != (env->n_maps ? env->n_maps : 1)) return fail(AAC_ERR_ARG, "aac_set_od_tables: one table per map (call aac_set_maps first)");
    std::vector<OdDev> host(n_maps);
    std::vector<void *> bufs;
    auto upload = [&](const void *src, size_t bytes, const void **dst) -> cudaError_t {
        void *d = nullptr;
        cudaError_t e = cudaMalloc(&d, bytes ? bytes : 1);
        if (e != cudaSuccess) return e;
        bufs.push_back(d);
        *dst = d;
        return cudaMemcpy(d, src, bytes, cudaMemcpyHostToDevice);
    };
    // every check runs before the first upload of a table, and an error frees what earlier tables uploaded
    auto reject = [&](const char *msg) {
        for (void *b : bufs) cudaFree(b);
        return fail(AAC_ERR_ARG, msg);
    };
    const int w_cap = env->cfg.variant == AAC_VARIANT_MM ? 31 : env->cfg.w_max;   // multipleMap keeps the remaining waypoints in a 32-bit mask
    bool any_plan = false;
    for (int m = 0; m < n_maps; ++m) {
        const AacOdTable &t = tables[m];
        // pools without paths: the reference line of every episode is searched on the device when the episode starts
        // (reset_world's per-episode jps_find_path, ATT:317), by the warp that re-initialises the env
        const bool pools_only = !t.path_off && !t.path_len && !t.path_cells;
        if (pools_only && (env->cfg.radar_targets || env->cfg.n_nbr_obs || env->cfg.eval_by_step))
            return reject("aac_set_od_tables: the sensor / evaluation configurations need tables with paths");
        if (t.n_cells < 2 || !t.cell_code || (!pools_only && (!t.path_off || !t.path_len || !t.path_cells))) return reject("aac_set_od_tables: incomplete table");
        for (int q = 0; q < 4; ++q)
            if (t.pool_off[q + 1] <= t.pool_off[q]) return reject("aac_set_od_tables: every quadrant pool needs at least one cell");
        if (t.pool_off[0] != 0 || t.pool_off[4] != t.n_cells) return reject("aac_set_od_tables: pool offsets do not cover the cells");
        if (!pools_only && (t.n_path_cells < 0 || (t.n_path_cells & 7))) return reject("aac_set_od_tables: path_cells must be padded to a multiple of 8");
        const size_t P = t.n_cells;
        const int gx = env->d_maps ? env->h_gx[m] : 255, gy = env->d_maps ? env->h_gy[m] : 255;
        for (size_t k = 0; k < P; ++k)
            if ((t.cell_code[k] >> 8) >= gx || (t.cell_code[k] & 255) >= gy) return reject("aac_set_od_tables: a pool cell lies outside the map's grid");
        for (size_t k = 0; k < P * P && !pools_only; ++k) {
            if (!t.path_len[k]) continue;
            if (t.path_len[k] < 2 || t.path_len[k] > w_cap || (t.path_off[k] & 7u)) return reject("aac_set_od_tables: a path has fewer than 2 or more than w_max (multipleMap: 31) vertices or is not 8-cell aligned");
            if ((int64_t)t.path_off[k] + ((t.path_len[k] + 7) & ~7) > t.n_path_cells) return reject("aac_set_od_tables: a path runs past the end of path_cells");
            for (int v = 0; v < t.path_len[k]; ++v) {
                const uint16_t c = t.path_cells[t.path_off[k] + v];
                if ((c >> 8) >= gx || (c & 255) >= gy) return reject("aac_set_od_tables: a path vertex lies outside the map's grid");
            }
        }
        OdDev &o = host[m];
        o.n_cells = t.n_cells;
        for (int q = 0; q < 5; ++q) o.pool_off[q] = t.pool_off[q];
        cudaError_t e;
        o.path_off = nullptr; o.path_len = nullptr; o.path_cells = nullptr;
        any_plan = any_plan || pools_only;
        if ((e = upload(t.cell_code, P * 2, (const void **)&o.cell_code)) != cudaSuccess ||
            (!pools_only && ((e = upload(t.path_off, P * P * 4, (const void **)&o.path_off)) != cudaSuccess ||
                             (e = upload(t.path_len, P * P, (const void **)&o.path_len)) != cudaSuccess ||
                             (e = upload(t.path_cells, (size_t)t.n_path_cells * 2, (const void **)&o.path_cells)) != cudaSuccess))) {
            for (void *b : bufs) cudaFree(b);
            return cuda_fail(e, "aac_set_od_tables upload");
        }
    }
    cudaFree(env->d_od);
    for (void *b : env->od_bufs) cudaFree(b);
    env->od_bufs = bufs;
    env->od_plan = any_plan;
    env->d_od = nullptr;
    CU(cudaMalloc(&env->d_od, sizeof(OdDev) * n_maps));
    CU(cudaMemcpy(env->d_od, host.data(), sizeof(OdDev) * n_maps, cudaMemcpyHostToDevice));
    return 0;
}

// the reference's grid search and collinear pruning, host only (ATT/jps_straight.py:17-70, ATT:321-331)
extern "C" int aac_plan_path(const uint8_t *occ, int32_t gx, int32_t gy, int32_t sx, int32_t sy, int32_t tx, int32_t ty, uint16_t *out_cells,
                             int32_t max_cells) {
    if (!occ || !out_cells || gx < 1 || gy < 1 || gx > 255 || gy > 255) return fail(AAC_ERR_ARG, "aac_plan_path: bad argument");
    if (sx < 0 || sy < 0 || tx < 0 || ty < 0 || sx >= gx || tx >= gx || sy >= gy || ty >= gy) return fail(AAC_ERR_ARG, "aac_plan_path: cell outside the grid");
    const int n = gx * gy;
    std::vector<uint8_t> status(n, 0);   // 0 unseen, 1 open, 2 closed
    std::vector<int> gcost(n, 0), fcost(n, 0), parent(n, -1), frontier;
    const int s = sx * gy + sy, t = tx * gy + ty;
    frontier.push_back(s);
    status[s] = 1;
    static const int DX[4] = {0, 0, -1, 1}, DY[4] = {-1, 1, 0, 0};
    int found = -1;
    while (!frontier.empty()) {
        size_t kb = 0;   // first entry, in discovery order, with the smallest f
        for (size_t k = 1; k < frontier.size(); ++k)
            if (fcost[frontier[k]] < fcost[frontier[kb]]) kb = k;
        const int cur = frontier[kb];
        frontier.erase(frontier.begin() + kb);
        status[cur] = 2;
        if (cur == t) { found = cur; break; }
        const int cx = cur / gy, cy = cur % gy;
        for (int d = 0; d < 4; ++d) {
            const int nx = cx + DX[d], ny = cy + DY[d];
            if (nx < 0 || ny < 0 || nx >= gx || ny >= gy || occ[nx * gy + ny]) continue;
            const int c = nx * gy + ny;
            if (status[c]) continue;   // a discovered cell is never re-queued or re-costed
            gcost[c] = gcost[cur] + 1;
            fcost[c] = gcost[c] + abs(nx - tx) + abs(ny - ty);
            parent[c] = cur;
            status[c] = 1;
            frontier.push_back(c);
        }
    }
    if (found < 0) return 0;
    std::vector<int> path;
    for (int c = found; c != -1; c = parent[c]) path.push_back(c);
    const int L = (int)path.size();   // path[L-1] = start ... path[0] = goal
    auto cell = [&](int k) { return path[L - 1 - k]; };
    std::vector<int> keep;
    keep.push_back(cell(0));
    if (L >= 2) {
        int dx = cell(1) / gy - cell(0) / gy, dy = cell(1) % gy - cell(0) % gy;
        for (int k = 2; k < L; ++k) {
            const int ex = cell(k) / gy - cell(k - 1) / gy, ey = cell(k) % gy - cell(k - 1) % gy;
            if (ex != dx || ey != dy) { keep.push_back(cell(k - 1)); dx = ex; dy = ey; }
        }
        keep.push_back(cell(L - 1));
    }
    if ((int)keep.size() > max_cells) return -1;
    for (size_t k = 0; k < keep.size(); ++k) out_cells[k] = (uint16_t)(((keep[k] / gy) << 8) | (keep[k] % gy));
    return (int)keep.size();
}

// ---- the same search on the device: one warp per origin / destination pair (aac_plan.cuh) ------------------------------
__global__ void __launch_bounds__(256) plan_paths_kernel(const uint8_t *__restrict__ occ, const int gx, const int gy, const uint16_t *__restrict__ pairs,
                                                         const long long n_pairs, uint16_t *__restrict__ out_cells, int *__restrict__ out_len,
                                                         const int max_cells, uint8_t *__restrict__ scratch) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    uint8_t *mine = scratch + (size_t)warp * plan_scratch_bytes((size_t)gx * gy);
    for (long long pi = warp; pi < n_pairs; pi += n_warps) {
        const int sc = pairs[2 * pi], tc = pairs[2 * pi + 1];
        const int s = (sc >> 8) * gy + (sc & 255), t = (tc >> 8) * gy + (tc & 255);
        const int cnt = plan_path_warp([&](const int c) { return occ[c] != 0; }, gx, gy, s, t, mine, out_cells + pi * max_cells, max_cells);
        if (lane == 0) out_len[pi] = cnt;
    }
}

extern "C" int aac_plan_paths_device(const uint8_t *occ, int32_t gx, int32_t gy, const uint16_t *pairs, int64_t n_pairs, uint16_t *out_cells,
                                     int32_t *out_len, int32_t max_cells, void *stream_) {
    if (!occ || !pairs || !out_cells || !out_len || gx < 1 || gy < 1 || gx > 255 || gy > 255 || max_cells < 1 || n_pairs < 0)
        return fail(AAC_ERR_ARG, "aac_plan_paths_device: bad argument");
    if (n_pairs == 0) return 0;
    for (int64_t k = 0; k < 2 * n_pairs; ++k)
        if ((pairs[k] >> 8) >= gx || (pairs[k] & 255) >= gy) return fail(AAC_ERR_ARG, "aac_plan_paths_device: cell outside the grid");
    cudaStream_t stream = (cudaStream_t)stream_;
    int dev = 0, sms = 0;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int threads = 256, wpc = threads / 32;
    int64_t grid = (n_pairs + wpc - 1) / wpc;
    if (grid > (int64_t)sms * 8) grid = (int64_t)sms * 8;   // 8 resident CTAs of 256 threads per SM: persistent warps
    const size_t n = (size_t)gx * gy, n_warps = (size_t)grid * wpc;
    uint8_t *d_occ = nullptr, *d_scratch = nullptr;
    uint16_t *d_pairs = nullptr, *d_cells = nullptr;
    int *d_len = nullptr;
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return e == cudaSuccess; };
    if (ok(cudaMalloc(&d_occ, n)) && ok(cudaMalloc(&d_pairs, (size_t)n_pairs * 4)) && ok(cudaMalloc(&d_cells, (size_t)n_pairs * max_cells * 2)) &&
        ok(cudaMalloc(&d_len, (size_t)n_pairs * 4)) && ok(cudaMalloc(&d_scratch, n_warps * plan_scratch_bytes(n))) &&
        ok(cudaMemcpyAsync(d_occ, occ, n, cudaMemcpyHostToDevice, stream)) &&
        ok(cudaMemcpyAsync(d_pairs, pairs, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, stream)) &&
        ok(cudaMemsetAsync(d_cells, 0, (size_t)n_pairs * max_cells * 2, stream))) {
        plan_paths_kernel<<<(unsigned)grid, threads, 0, stream>>>(d_occ, gx, gy, d_pairs, n_pairs, d_cells, d_len, max_cells, d_scratch);
        ok(cudaGetLastError());
        ok(cudaMemcpyAsync(out_cells, d_cells, (size_t)n_pairs * max_cells * 2, cudaMemcpyDeviceToHost, stream));
        ok(cudaMemcpyAsync(out_len, d_len, (size_t)n_pairs * 4, cudaMemcpyDeviceToHost, stream));
        ok(cudaStreamSynchronize(stream));
    }
    cudaFree(d_occ); cudaFree(d_pairs); cudaFree(d_cells); cudaFree(d_len); cudaFree(d_scratch);
    if (e != cudaSuccess) return cuda_fail(e, "aac_plan_paths_device");
    return 0;
}

extern "C" int aac_bind_state(AacEnv *env, const AacState *s) {
    if (!env || !s) return fail(AAC_ERR_ARG, "aac_bind_state: null argument");
    if (!s->px || !s->py || !s->vx || !s->vy || !s->heading || !s->meta || !s->ref_cells || !s->ref_w || !s->ep_step || !s->ep_index || !s->ep_return)
        return fail(AAC_ERR_ARG, "aac_bind_state: a required state array is NULL");
    if (env->cfg.variant == AAC_VARIANT_MM && (!s->map_id || !s->wp_mask)) return fail(AAC_ERR_ARG, "aac_bind_state: the multipleMap variant needs map_id and wp_mask");
    env->st = *s;
    env->bound = true;
    return 0;
}

static int check_out(const AacEnv *env, const AacOut *o, int mode) {
    const int f = env->cfg.out_flags;
    const bool v2 = env->cfg.variant == AAC_VARIANT_V2;
    const bool pairs = env->cfg.n_agents > 1;   // a single drone has no neighbour blocks (zero-size tensors)
    if (!o) return fail(AAC_ERR_ARG, "output block is NULL");
    if (!o->norm_own || !o->radar || (v2 && pairs && !o->norm_nbr)) return fail(AAC_ERR_ARG, "norm_own / norm_nbr / radar must be provided");
    if (mode == MODE_STEP && (!o->reward || !o->done || !o->check_goal || !o->bbc || !o->terminated || !o->tcpa_min))
        return fail(AAC_ERR_ARG, "reward / done / check_goal / bbc / terminated / tcpa_min must be provided");
    if ((f & AAC_OUT_NBR6) && pairs && !o->norm_nbr6) return fail(AAC_ERR_ARG, "AAC_OUT_NBR6 set but norm_nbr6 is NULL");
    if ((f & AAC_OUT_RAW) && (!o->raw_own || (v2 && pairs && !o->raw_nbr) || ((f & AAC_OUT_NBR6) && pairs && !o->raw_nbr6))) return fail(AAC_ERR_ARG, "AAC_OUT_RAW set but a raw_* pointer is NULL");
    if ((f & AAC_OUT_TCPA_PAIR) && pairs && (!o->tcpa_pair || !o->nbr_order)) return fail(AAC_ERR_ARG, "AAC_OUT_TCPA_PAIR set but tcpa_pair / nbr_order is NULL");
    if ((f & AAC_OUT_RADAR_AUX) && (!o->radar_min || !o->radar_hit)) return fail(AAC_ERR_ARG, "AAC_OUT_RADAR_AUX set but radar_min / radar_hit is NULL");
    if ((f & AAC_OUT_PARTS) && mode == MODE_STEP && (!o->parts || !o->branch)) return fail(AAC_ERR_ARG, "AAC_OUT_PARTS set but parts / branch is NULL");
    return 0;
}

// advance every per-env / per-drone pointer of the state and output blocks to env `e_lo`
static void offset_rows(AacState &s, AacOut &o, size_t e_lo, size_t N, size_t R, size_t W, size_t D, size_t Mo) {
    const size_t a = e_lo * N, M = N - 1;   // Mo: neighbours per drone in norm_nbr / raw_nbr (nearest-N selection)
#define ADV(ptr, n) if (ptr) ptr += (n)
    ADV(s.px, a); ADV(s.py, a); ADV(s.vx, a); ADV(s.vy, a); ADV(s.heading, a); ADV(s.meta, a); ADV(s.ref_cells, a * W); ADV(s.ref_w, a);
    ADV(s.wall_count, a); ADV(s.ep_step, e_lo); ADV(s.ep_index, e_lo); ADV(s.ep_return, e_lo); ADV(s.map_id, e_lo); ADV(s.wp_mask, a);
    ADV(o.norm_own, a * D); ADV(o.norm_nbr, a * 5 * Mo); ADV(o.radar, a * R); ADV(o.norm_nbr6, a * 6 * M);
    ADV(o.raw_own, a * D); ADV(o.raw_nbr, a * 5 * Mo); ADV(o.raw_nbr6, a * 6 * M); ADV(o.cloud_contact, a);
    ADV(o.reward, a); ADV(o.done, a); ADV(o.check_goal, a); ADV(o.bbc, e_lo * 4); ADV(o.terminated, e_lo); ADV(o.tcpa_min, a * 4);
    ADV(o.tcpa_pair, a * 4 * M); ADV(o.nbr_order, a * M); ADV(o.radar_min, a * R); ADV(o.radar_hit, a * R); ADV(o.parts, a * 8); ADV(o.branch, a);
#undef ADV
}

// launch over the envs [e_lo, e_lo + e_cnt) of the handle (e_cnt <= 0: all of them).  `pair` selects the
// ping-pong group-counter pair: launches that may run concurrently must use different pairs.
static int launch(AacEnv *env, int mode, const uint8_t *mask, const float *actions, const AacOut *out, void *stream, int autoreset = 0,
                  int e_lo = 0, int e_cnt = 0, int pair = 0) {
    if (!env) return fail(AAC_ERR_ARG, "null handle");
    if (!env->bound) return fail(AAC_ERR_STATE, "aac_bind_state has not been called");
    if (!env->d_maps) return fail(AAC_ERR_STATE, "aac_set_maps has not been called");
    if ((mode == MODE_RESET || autoreset) && !env->d_bank_cells && !env->d_od) return fail(AAC_ERR_STATE, "neither aac_set_bank nor aac_set_od_tables has been called");
    if (mode == MODE_STEP && !actions) return fail(AAC_ERR_ARG, "actions is NULL");
    const int rc = check_out(env, out, mode);
    if (rc) return rc;
    KParams p;
    memset(&p, 0, sizeof(p));
    const AacConfig &c = env->cfg;
    p.E = c.n_envs; p.N = c.n_agents; p.R = c.n_rays; p.W = c.w_max; p.G = env->group;
    p.radar_mode = c.radar_mode; p.sum_reward = c.sum_reward; p.ep_len = c.episode_length; p.out_flags = c.out_flags; p.eval_by_step = c.eval_by_step;
    p.radar_targets = c.radar_targets; p.n_nbr_obs = c.n_nbr_obs; p.n_clouds = c.n_clouds;
    for (int ci = 0; ci < c.n_clouds; ++ci) {
        // local frame; the stop step is found by walking the cloud as the fork does (calculate_next_position, float64)
        const double ox = env->h_ox, oy = env->h_oy, travel = (double)c.clouds[ci][5] * c.dt;
        double x = c.clouds[ci][0], y = c.clouds[ci][1];
        const double gx = c.clouds[ci][2], gy = c.clouds[ci][3];
        const double d0 = sqrt((gx - x) * (gx - x) + (gy - y) * (gy - y));
        int n_stop = 0;
        while (n_stop < (1 << 20) && travel > 0) {
            const double d = sqrt((gx - x) * (gx - x) + (gy - y) * (gy - y));
            if (d < 1.0) break;
            x += (gx - x) / d * travel; y += (gy - y) / d * travel;
            ++n_stop;
        }
        CloudDev &cd = p.clouds[ci];
        cd.sx = (float)(c.clouds[ci][0] - ox); cd.sy = (float)(c.clouds[ci][1] - oy);
        cd.dx = d0 >= 1.0 ? (float)((gx - c.clouds[ci][0]) / d0) : 0.0f; cd.dy = d0 >= 1.0 ? (float)((gy - c.clouds[ci][1]) / d0) : 0.0f;
        cd.travel = (float)travel; cd.radius = c.clouds[ci][4]; cd.n_stop = n_stop; cd.pad_ = 0;
    }
    p.cell = env->cell; p.dt = c.dt; p.vmax = c.vmax; p.acc_max = c.acc_max; p.prot = c.prot; p.ray_len = c.ray_len; p.goal_r = c.goal_r;
    p.env_id_base = c.env_id_base; p.seed = c.seed;
    p.maps = env->d_maps; p.n_maps = env->n_maps; p.n_staged = env->n_staged; p.ray_tab = env->d_ray; p.dda_tab = env->d_dda; p.walk_tab = env->d_walk; p.autoreset = autoreset;
    p.bank_cells = env->d_bank_cells; p.bank_w = env->d_bank_w; p.bank_map = env->d_bank_map; p.n_scen = env->n_scen; p.od = env->d_od;
    p.mask = mask; p.actions = actions; p.stats = env->d_stats;
    p.rtab = env->rtab; p.rtab_min = env->rtab_min; p.rtab_hit = env->rtab_hit; p.rtab_minr = env->rtab_minr;
    p.work = env->d_work + 4 * pair; p.parity = (int)(env->pair_launches[pair] & 1);
    p.st = env->st; p.out = *out; p.CL = env->cl; p.WL = env->wl;
    if (env->od_plan && (mode == MODE_RESET || autoreset)) {   // per-episode path search: scratch for every warp that can be resident
        size_t cells = 0;
        for (int m = 0; m < (env->n_maps ? env->n_maps : 1); ++m) cells = std::max(cells, (size_t)env->h_gx[m] * env->h_gy[m]);
        p.plan_stride = (unsigned)plan_scratch_bytes(cells);
        if (!env->d_plan_scratch) CU(cudaMalloc(&env->d_plan_scratch, (size_t)env->sms * 64 * p.plan_stride   /* 64 resident warps per SM at most */));
        p.plan_scratch = env->d_plan_scratch;
    }
    if (autoreset == 2) {   // phased launch: per-group completion flags, tagged with the launch's epoch
        if (e_lo % env->group) return fail(AAC_ERR_ARG, "phased launch: the env range must start on a group boundary");
        if (!env->d_flags) {
            const size_t n_groups = ((size_t)c.n_envs + env->group - 1) / env->group;
            CU(cudaMalloc(&env->d_flags, n_groups * sizeof(unsigned long long)));
            CU(cudaMemset(env->d_flags, 0, n_groups * sizeof(unsigned long long)));
        }
        if (env->epoch == 0x7FFFFFFF) {   // the epoch wraps: start over with clean flags
            CU(cudaDeviceSynchronize());
            CU(cudaMemset(env->d_flags, 0, (((size_t)c.n_envs + env->group - 1) / env->group) * sizeof(unsigned long long)));
            env->epoch = 0;
        }
        p.flags = env->d_flags + e_lo / env->group;
        p.epoch = ++env->epoch;
    }
    if (e_cnt > 0) {
        p.E = e_cnt;
        p.env_id_base += e_lo;
        if (p.mask) p.mask += e_lo;
        if (p.actions) p.actions += (size_t)e_lo * c.n_agents * 2;
        offset_rows(p.st, p.out, e_lo, c.n_agents, c.n_rays, c.w_max, own_dim(c.variant, c.n_agents), c.n_nbr_obs > 0 ? c.n_nbr_obs : c.n_agents - 1);
    }
    cudaError_t e = launch_env_kernel(c.variant, p, mode, env->threads, env->sms, &env->grid, (cudaStream_t)stream);
    if (e != cudaSuccess) return cuda_fail(e, "env_kernel launch");
    env->launches += 1;
    env->pair_launches[pair] += 1;
    return 0;
}

extern "C" int aac_set_radar_table(AacEnv *env, const float *radar, const float *radar_min, const int16_t *radar_hit, const uint32_t *min_bits) {
    if (!env) return fail(AAC_ERR_ARG, "null handle");
    if (!radar && !radar_min && !radar_hit && !min_bits) {   // remove the table
        env->rtab = env->rtab_min = nullptr; env->rtab_hit = nullptr; env->rtab_minr = nullptr;
        return 0;
    }
    if (env->cfg.variant == AAC_VARIANT_ATT) return fail(AAC_ERR_ARG, "aac_set_radar_table: the one_model_att radar senses the other drones, not the map");
    if (!radar || !min_bits) return fail(AAC_ERR_ARG, "aac_set_radar_table: radar and min_bits are required");
    if ((env->cfg.out_flags & AAC_OUT_RADAR_AUX) && (!radar_min || !radar_hit))
        return fail(AAC_ERR_ARG, "aac_set_radar_table: radar_min and radar_hit are required with AAC_OUT_RADAR_AUX");
    env->rtab = radar; env->rtab_min = radar_min; env->rtab_hit = radar_hit; env->rtab_minr = min_bits;
    return 0;
}

extern "C" int aac_reset(AacEnv *env, const uint8_t *mask_dev, const AacOut *out, void *stream) {
    return launch(env, MODE_RESET, mask_dev, nullptr, out, stream);
}
extern "C" int aac_observe(AacEnv *env, const AacOut *out, void *stream) { return launch(env, MODE_OBSERVE, nullptr, nullptr, out, stream); }
extern "C" int aac_step(AacEnv *env, const float *actions_dev, const AacOut *out, void *stream) {
    return launch(env, MODE_STEP, nullptr, actions_dev, out, stream);
}
// The step of every env, then the re-initialisation of the envs it terminated - as two launches on the stream, or as the two
// loops of one phased launch - leave exactly the state and outputs of the single fused launch (tests/test_gpu_parity.py::
// test_fused_autoreset_equals_step_then_autoreset, ::test_mode_specialised_launches_equal_fused_launch) and are faster: the
// code a warp executes at any one time is a smaller part of the kernel, which is bound by instruction issue and
// instruction-cache misses (C3: 0.330 ms against 0.362 ms fused, round 1).
// Step loop and reset loop in ONE launch (env_kernel MT_PHASED) instead of two launches?  The reset work then fills the tail
// of the step loop - the last group of every warp, run at falling occupancy - and one launch ramp goes away.  Measured on
// B200, same box: C3 (21 846 groups, 4.6 per resident warp) 0.2531 -> 0.2453 ms; C5's shard (131 072 groups, 37 per warp:
// the tail is 1 / 37 of the loop) 1.715 -> 1.748 ms.  So: where the instantiation exists and a resident warp gets at most
// 8 groups, unless the configuration says which.
static bool phased_autoreset(const AacEnv *env, int e_cnt = 0) {
    const AacConfig &c = env->cfg;
    if (c.autoreset_launches != 3 && c.autoreset_launches != 0) return false;
    if (env->od_plan) return false;   // per-episode path search lives in the run-time-mode kernels
    if (c.autoreset_launches == 0) {
        if (c.variant != AAC_VARIANT_V2) return false;   // multipleMap at C4: two launches 0.0713, phased 0.0734, fused 0.0747 ms
        const long long groups = ((long long)(e_cnt > 0 ? e_cnt : c.n_envs) + env->group - 1) / env->group;
        if (groups > 8LL * env->sms * 32) return false;
    }
    KParams q;
    memset(&q, 0, sizeof(q));
    q.N = c.n_agents; q.R = c.n_rays; q.out_flags = c.out_flags; q.radar_targets = c.radar_targets; q.n_nbr_obs = c.n_nbr_obs; q.eval_by_step = c.eval_by_step;
    return phased_launch_available(c.variant, q);
}
static int step_then_reset(AacEnv *env, const float *actions_dev, const AacOut *out, void *stream, int e_lo = 0, int e_cnt = 0, int pair = 0) {
    if (out && !out->terminated) return fail(AAC_ERR_ARG, "reward / done / check_goal / bbc / terminated / tcpa_min must be provided");
    // both loops in one launch where the shape has the instantiation: the reset work fills the step loop's tail
    if (phased_autoreset(env, e_cnt)) return launch(env, MODE_STEP, nullptr, actions_dev, out, stream, 2, e_lo, e_cnt, pair);
    const int rc = launch(env, MODE_STEP, nullptr, actions_dev, out, stream, 0, e_lo, e_cnt, pair);
    if (rc) return rc;
    return launch(env, MODE_RESET, out->terminated, nullptr, out, stream, 0, e_lo, e_cnt, pair);
}
// One fused launch or two?  Measured on B200 (tests/tools/unfused_time.py): the second launch costs about 0.011 ms,
// the smaller executed code per launch saves 9-12 % of a tdCPA_forV2 step: C3 (65 536 x 10 x 36) 0.362 -> 0.330 ms,
// C5 (131 072 x 20 x 72) 2.88 -> 2.52 ms, but C2 (one_model_att 4096 x 3) 0.031 -> 0.041 and C4 (multipleMap 65 536 x 3)
// 0.088 -> 0.100 (round 1).  Round 2, with every finished env of a warp's group re-initialised in one pass, C4 turned: fused
// 0.0747, two launches 0.0713 ms (C2 stays: 0.0347 against 0.0447).  Two launches (or the phased launch, phased_autoreset)
// for tdCPA_forV2 batches of at least 8M (ray + pair) items and multipleMap batches of at least 3M, one fused launch
// otherwise, unless the configuration says which.
static bool split_autoreset(const AacEnv *env, int e_cnt = 0) {
    const AacConfig &c = env->cfg;
    if (c.autoreset_launches == 1) return false;
    if (c.autoreset_launches == 2) return true;
    if (c.autoreset_launches == 3) return phased_autoreset(env, e_cnt);
    const double items = (double)(e_cnt > 0 ? e_cnt : c.n_envs) * c.n_agents * (c.n_agents - 1 + c.n_rays);
    return (c.variant == AAC_VARIANT_V2 && items >= 8e6) || (c.variant == AAC_VARIANT_MM && items >= 3e6);
}
extern "C" int aac_step_autoreset(AacEnv *env, const float *actions_dev, const AacOut *out, void *stream) {
    if (!env) return fail(AAC_ERR_ARG, "null handle");
    if (split_autoreset(env)) return step_then_reset(env, actions_dev, out, stream);
    return launch(env, MODE_STEP, nullptr, actions_dev, out, stream, 1);
}
extern "C" int aac_step_fused(AacEnv *env, const float *actions_dev, const AacOut *out, void *stream) {
    return launch(env, MODE_STEP, nullptr, actions_dev, out, stream, 1);
}
extern "C" int aac_autoreset(AacEnv *env, const AacOut *out, void *stream) {
    if (!out || !out->terminated) return fail(AAC_ERR_ARG, "aac_autoreset: out->terminated is NULL");
    return launch(env, MODE_RESET, out->terminated, nullptr, out, stream);
}

// one device -> host copy of the env rows [e_lo, e_lo + cnt) of an output array (skipped when the host side is NULL)
static int copy_rows(void *host, const void *dev, size_t per_env, size_t e_lo, size_t cnt, cudaStream_t s, const char *name) {
    if (!host) return 0;
    if (!dev) return fail(AAC_ERR_ARG, "aac_step_host: host buffer without a device buffer: %s", name);
    CU(cudaMemcpyAsync((char *)host + e_lo * per_env, (const char *)dev + e_lo * per_env, cnt * per_env, cudaMemcpyDeviceToHost, s));
    return 0;
}

extern "C" int aac_step_host(AacEnv *env, const float *actions_host, const AacOut *od, const AacOut *oh, int32_t autoreset, void *stream_) {
    if (!env || !actions_host || !od || !oh) return fail(AAC_ERR_ARG, "aac_step_host: null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    const size_t E = env->cfg.n_envs, N = env->cfg.n_agents, M = N - 1, R = env->cfg.n_rays;
    const size_t D = own_dim(env->cfg.variant, (int)N);
    const size_t Mo = env->cfg.n_nbr_obs > 0 ? (size_t)env->cfg.n_nbr_obs : M;
    if (!env->d_actions) CU(cudaMalloc(&env->d_actions, E * N * 2 * sizeof(float)));
    if (!env->pipe[0]) {
        for (auto &s : env->pipe) CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
        for (auto &e : env->pipe_ev) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    // The envs are cut into chunks that rotate over three streams: while chunk c's observations cross PCIe, chunk
    // c+1 steps and chunk c+2's actions arrive.  Small batches go as one chunk.  With autoreset the terminal
    // transition's reward / done / flags are kept and the observation rows of the finished envs carry their reset
    // observation (one fused launch per chunk).
    const size_t G = env->group;
    size_t n_chunks = E * N >= 65536 ? 8 : 1;
    const size_t per = ((E + n_chunks - 1) / n_chunks + G - 1) / G * G;
    n_chunks = (E + per - 1) / per;
#define ROWS(field, per_env, lo, cnt, s)                                                           \
    do {                                                                                           \
        const int rc_ = copy_rows(oh->field, od->field, (per_env), (lo), (cnt), (s), #field);      \
        if (rc_) return rc_;                                                                       \
    } while (0)
    CU(cudaEventRecord(env->pipe_ev[3], stream));
    for (size_t ch = 0; ch < n_chunks; ++ch) {
        const size_t e_lo = ch * per, cnt = (e_lo + per <= E ? per : E - e_lo), a_lo = e_lo * N;
        cudaStream_t s = env->pipe[ch % 3];
        if (ch < 3) CU(cudaStreamWaitEvent(s, env->pipe_ev[3], 0));
        CU(cudaMemcpyAsync(env->d_actions + a_lo * 2, actions_host + a_lo * 2, cnt * N * 2 * sizeof(float), cudaMemcpyHostToDevice, s));
        const int rc = (autoreset && split_autoreset(env, (int)cnt)) ? step_then_reset(env, env->d_actions, od, s, (int)e_lo, (int)cnt, 1 + (int)ch)
                                                                     : launch(env, MODE_STEP, nullptr, env->d_actions, od, s, autoreset ? 1 : 0, (int)e_lo, (int)cnt, 1 + (int)ch);
        if (rc) return rc;
        // the wide observation blocks leave with their chunk ...
        ROWS(norm_own, N * D * 4, e_lo, cnt, s); ROWS(norm_nbr, N * 5 * Mo * 4, e_lo, cnt, s); ROWS(radar, N * R * 4, e_lo, cnt, s);
        ROWS(norm_nbr6, N * M * 24, e_lo, cnt, s); ROWS(raw_own, N * D * 4, e_lo, cnt, s); ROWS(raw_nbr, N * 5 * Mo * 4, e_lo, cnt, s);
        ROWS(raw_nbr6, N * M * 24, e_lo, cnt, s); ROWS(tcpa_pair, N * M * 16, e_lo, cnt, s); ROWS(radar_min, N * R * 4, e_lo, cnt, s);
        ROWS(radar_hit, N * R * 2, e_lo, cnt, s);
    }
    // ... the narrow arrays (a few bytes per drone) go once for the whole batch, after every chunk has stepped:
    // fewer, larger transfers
    cudaStream_t last = env->pipe[(n_chunks - 1) % 3];
    for (int k = 0; k < 3 && (size_t)k < n_chunks; ++k) {
        if (env->pipe[k] == last) continue;
        CU(cudaEventRecord(env->pipe_ev[k], env->pipe[k]));
        CU(cudaStreamWaitEvent(last, env->pipe_ev[k], 0));
    }
    ROWS(reward, N * 4, 0, E, last); ROWS(done, N, 0, E, last); ROWS(check_goal, N, 0, E, last); ROWS(bbc, 4, 0, E, last);
    ROWS(terminated, 1, 0, E, last); ROWS(tcpa_min, N * 16, 0, E, last); ROWS(nbr_order, N * M, 0, E, last);
    ROWS(parts, N * 32, 0, E, last); ROWS(branch, N, 0, E, last); ROWS(cloud_contact, N, 0, E, last);
#undef ROWS
    CU(cudaEventRecord(env->pipe_ev[0], last));
    CU(cudaStreamWaitEvent(stream, env->pipe_ev[0], 0));
    CU(cudaStreamSynchronize(stream));
    return 0;
}

extern "C" int aac_read_stats(AacEnv *env, double *stats_host, int32_t reset, void *stream_) {
    if (!env || !stats_host) return fail(AAC_ERR_ARG, "aac_read_stats: null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    CU(cudaMemcpyAsync(stats_host, env->d_stats, sizeof(double) * AAC_N_STATS, cudaMemcpyDeviceToHost, stream));
    if (reset) CU(cudaMemsetAsync(env->d_stats, 0, sizeof(double) * AAC_N_STATS, stream));
    CU(cudaStreamSynchronize(stream));
    return 0;
}
