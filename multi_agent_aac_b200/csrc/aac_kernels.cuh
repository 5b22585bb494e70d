// Device-side parameter block shared by aac_kernels.cu (kernels) and aac_capi.cu (C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aac_env.h"
#include "aac_radar.cuh"
#include "aac_plan.cuh"

namespace aac {

constexpr int MAP_STRIDE_CELLS = 1024;   // rows per map in the radar table (= AAC_MAP_STRIDE: gx * gy <= 1024)
constexpr int MAX_THREADS = 256;

// origin / destination table of one map on the device (AacOdTable with device pointers)
struct OdDev {
    int n_cells;
    int pool_off[5];
    const uint16_t *cell_code;
    const uint32_t *path_off;
    const uint8_t *path_len;
    const uint16_t *path_cells;
};

// one cloud on the device (local frame): position after k steps = start + dir * (min(k, n_stop) * travel)
struct CloudDev {
    float sx, sy, dx, dy;   // start, unit direction towards the goal (0 when the start already lies within 1 m of it)
    float travel, radius;   // speed * dt, radius of the 64-gon
    int n_stop, pad_;       // steps after which the cloud has come closer than 1 m to its goal and stays
};

enum Mode : int { MODE_STEP = 0, MODE_OBSERVE = 1, MODE_RESET = 2 };

// shared-memory carve-up (byte offsets), computed once on the host: CTA-wide data, then one slice per warp
struct CtaLayout {
    unsigned map, ray, dda, walk, bar, warps, total;
};
// Per-warp slice (32 drone slots).  The fixed-size arrays sit at compile-time offsets so the kernel addresses them
// as `slice + immediate`; the arrays whose size depends on the drone count follow at WS_VAR.
constexpr unsigned WS_CUR = 0;                      // px py vx vy heading        5 x 32 floats
constexpr unsigned WS_PRE = WS_CUR + 5 * 128;       // pre_pos, pre_vel           4 x 32 floats
constexpr unsigned WS_META = WS_PRE + 4 * 128;      // meta, meta2, min radar bits, result flags, waypoint mask
constexpr unsigned WS_AGR = WS_META + 5 * 128;      // reward
constexpr unsigned WS_BYTES = WS_AGR + 128;         // at-goal flag, ref-line vertex count, per-env scratch, map row
constexpr unsigned WS_WIN = WS_BYTES + 4 * 32;      // 5x5 occupancy window | line / slow flags (aac_radar.cuh), own cell ix | iy << 16
constexpr unsigned WS_WREL = WS_WIN + 256;          // float4: low corner of the drone's cell and the boundary lines in reach, relative to the drone
constexpr unsigned WS_STG = WS_WREL + 512;          // transient staging of one warp iteration's pair blocks
constexpr unsigned WS_C8 = WS_STG + 32 * 6 * 4;     // the first 8 vertices of every drone's reference line (one 16-byte load per drone)
constexpr unsigned WS_VAR = WS_C8 + 512;            // d2 [32][M|1] floats, order [32][M] bytes, own rows, raw own rows
struct WarpLayout {
    unsigned d2, order, own, raw_own, total;        // byte offsets of the variable part inside the slice
};

struct KParams {
    int E, N, R, W, G;      // G = envs per warp (G * N <= 32)
    int radar_mode, sum_reward, ep_len, out_flags, eval_by_step;
    int radar_targets, n_nbr_obs, n_clouds;   // the later fork's sensor classes (aac_env.h); all zero = off
    CloudDev clouds[AAC_MAX_CLOUDS];
    // radar of a drone standing on a cell centre, per (map, cell, ray): what every freshly reset drone observes (aac_set_radar_table)
    const float *rtab, *rtab_min;
    const int16_t *rtab_hit;
    const unsigned *rtab_minr;
    float dt, vmax, acc_max, prot, ray_len, goal_r;
    float cell;             // cell size shared by every map of the handle
    long long env_id_base;
    unsigned long long seed;
    const MapDev *maps;
    int n_maps;
    int n_staged;           // maps staged per CTA in shared memory by the prologue's bulk copy (1, or all of them: multipleMap)
    const float4 *ray_tab;  // [R] (dx, dy, 1/dx, 1/dy) of the ray at k*360/R degrees; exact zeros on the axes, 1/0 = +inf
    const DdaRay *dda_tab;  // [R] the same rays as the constants of the cell walk (aac_radar.cuh)
    const uint4 *walk_tab;  // [WALK_BYTES / 16] the walk table
    int autoreset;          // MODE_STEP: re-initialise the envs that terminate and emit their reset observation
    const uint16_t *bank_cells;
    const uint8_t *bank_w;
    const int32_t *bank_map;
    int n_scen;
    const OdDev *od;        // [n_maps] origin / destination tables, or NULL (resets then use the scenario bank)
    const uint8_t *mask;  // MODE_RESET: per-env byte, NULL = every env
    const float *actions;
    double *stats;        // [AAC_N_STATS]
    int *work;            // [2] group counters of the persistent warps, ping-pong between launches
    int parity;           // which counter pair this launch consumes: work[2 * parity + {0, 1}] (it zeroes the other pair)
    uint8_t *plan_scratch;       // per-warp scratch of the per-episode path search (pools-only origin / destination tables), else NULL
    unsigned plan_stride;        // bytes per warp
    unsigned long long *flags;   // [groups] phased launch: (epoch of the step loop that last completed the group) << 32 | its terminated envs
    int epoch;            // this launch's epoch (phased launch)
    AacState st;
    AacOut out;
    CtaLayout CL;
    WarpLayout WL;
};

__host__ __device__ inline unsigned align16(unsigned x) { return (x + 15u) & ~15u; }

__host__ __device__ inline int own_dim(int variant, int N) {
    return variant == AAC_VARIANT_ATT ? 6 + 4 * (N - 1) : (variant == AAC_VARIANT_V2 ? 7 : 6);
}

__host__ __device__ inline WarpLayout make_warp_layout(int variant, int N, int flags) {
    WarpLayout L;
    const unsigned M = N - 1, Mp = M | 1, M1 = M ? M : 1;
    const unsigned D = own_dim(variant, N);
    unsigned o = WS_VAR;
    L.d2 = o; o = align16(o + 32 * Mp * 4);        // neighbour distances^2 in iteration order, odd row stride
    L.order = o; o = align16(o + 32 * M1);
    L.own = o; o = align16(o + 32 * D * 4);        // own rows of the observation
    L.raw_own = o; o = align16(o + ((flags & AAC_OUT_RAW) ? 32 * D * 4 : 0));
    L.total = o;
    return L;
}

constexpr int MAX_STAGED_MAPS = 16;   // multipleMap: the whole table travels to shared memory when it has at most this many maps (5 KB)

inline CtaLayout make_cta_layout(const WarpLayout &WL, int R, int warps, int n_staged = 1) {
    CtaLayout L;
    unsigned o = 0;
    auto take = [&](unsigned bytes) { unsigned r = o; o = align16(o + bytes); return r; };
    L.map = take(n_staged * sizeof(MapDev));
    L.ray = take(R * 16);
    L.dda = take(R * sizeof(DdaRay));
    L.walk = take(WALK_BYTES);
    L.bar = take(16);
    L.warps = take(0);
    L.total = L.warps + warps * WL.total;
    return L;
}

// fills the __constant__ polygon tables of the current device (call once per device)
cudaError_t upload_constants();
// a phased launch (autoreset == 2: step loop, then reset loop, one launch) exists for this variant / shape / output set
bool phased_launch_available(int variant, const KParams &p);
cudaError_t launch_env_kernel(int variant, const KParams &p, int mode, int threads, int sms, int *grid_cache, cudaStream_t stream);
int max_smem_optin();

}  // namespace aac
