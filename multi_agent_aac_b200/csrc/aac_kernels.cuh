// Device-side parameter block shared by aac_kernels.cu (kernels) and aac_capi.cu (C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/aac_env.h"

namespace aac {

constexpr int MAP_PAD = 4;      // free cells added on every side of the occupancy bitmap
constexpr int MAP_WORDS = 64;   // 2048 bits: (gx + 8) * (gy + 8) + 32 must fit
constexpr int MAX_THREADS = 256;

// one map in the LOCAL frame (origin = bound centre): the bound is x in [-hx, hx], y in [-hy, hy].
// Cell (ix, iy) covers [ex0 + ix*cell, ex0 + (ix+1)*cell] x [ey0 + iy*cell, ...]; occupancy is the
// padded bitmap bit (ix + MAP_PAD) * pgy + (iy + MAP_PAD).  320 bytes, staged per CTA by one bulk copy.
struct __align__(16) MapDev {
    int gx, gy, pgx, pgy;
    float hx, hy;          // half spans of the bound
    float ex0, ey0;        // local lower edge of cell (0,0)
    float ox, oy;          // global coordinates of the local origin
    float xmin_g, ymin_g;  // global bound minima (ATT applies scale_pos to a delta)
    float cell, inv_cell;
    float pad0, pad1;
    uint32_t bits[MAP_WORDS];
};
static_assert(sizeof(MapDev) % 16 == 0, "MapDev is moved with 16-byte bulk copies");

enum Mode : int { MODE_STEP = 0, MODE_OBSERVE = 1, MODE_RESET = 2 };

// shared-memory carve-up (byte offsets), computed once on the host
struct SmemLayout {
    unsigned maps, ray, envi, cur, pre, meta, meta2, cells, refw, d2, order, win, tc, pflag, agf, agr,
        own, nbr, radar, nbr6, raw_own, raw_nbr, raw_nbr6, tmin, tpair, rmin, rhit, parts, total;
};

struct KParams {
    int E, N, R, W, TE;
    int radar_mode, sum_reward, ep_len, out_flags;
    float dt, vmax, acc_max, prot, ray_len, goal_r;
    long long env_id_base;
    unsigned long long seed;
    const MapDev *maps;
    int n_maps;
    const float2 *ray_dir;  // [R] (cos, sin) of k*360/R degrees, exact zeros on the axes
    const uint16_t *bank_cells;
    const uint8_t *bank_w;
    const int32_t *bank_map;
    int n_scen;
    const uint8_t *mask;  // MODE_RESET: per-env byte, NULL = every env
    const float *actions;
    double *stats;        // [AAC_N_STATS]
    AacState st;
    AacOut out;
    SmemLayout L;
};

__host__ __device__ inline unsigned align16(unsigned x) { return (x + 15u) & ~15u; }

__host__ __device__ inline int own_dim(int variant, int N) {
    return variant == AAC_VARIANT_ATT ? 6 + 4 * (N - 1) : (variant == AAC_VARIANT_V2 ? 7 : 6);
}

inline SmemLayout make_layout(int variant, int TE, int N, int R, int W, int flags) {
    SmemLayout L;
    const unsigned A = TE * N, M = N - 1;
    const unsigned D = own_dim(variant, N);
    const bool v2 = variant == AAC_VARIANT_V2;
    unsigned o = 0;
    auto take = [&](unsigned bytes) { unsigned r = o; o = align16(o + bytes); return r; };
    const unsigned nmap = variant == AAC_VARIANT_MM ? TE : 1;
    L.maps = take(nmap * sizeof(MapDev));
    L.ray = take(R * 8);
    L.envi = take(TE * 16);          // per env: active, scenario, map row, flags
    L.cur = take(5 * A * 4);         // px py vx vy heading
    L.pre = take(4 * A * 4);         // pre_pos, pre_vel
    L.meta = take(A * 4);
    L.meta2 = take(A * 4);
    L.cells = take(A * W * 2);
    L.refw = take(A);
    L.d2 = take(A * N * 4);
    L.order = take(A * (M ? M : 1));
    L.win = take(A * 8);             // 4x4 occupancy window: mask, ix0 | iy0 << 16
    L.tc = take(A * M * 8);          // current tcpa, d_tcpa per ordered pair
    L.pflag = take(A * (M ? M : 1)); // per pair: bit0 cur conflict, bit1 pre conflict
    L.agf = take(A * 4);             // per agent result flags
    L.agr = take(A * 4);             // per agent reward
    L.own = take(A * D * 4);
    L.nbr = take(v2 ? A * 5 * M * 4 : 0);
    L.radar = take(A * R * 4);
    L.nbr6 = take((flags & AAC_OUT_NBR6) ? A * M * 6 * 4 : 0);
    L.raw_own = take((flags & AAC_OUT_RAW) ? A * D * 4 : 0);
    L.raw_nbr = take((flags & AAC_OUT_RAW) && v2 ? A * 5 * M * 4 : 0);
    L.raw_nbr6 = take((flags & AAC_OUT_RAW) && (flags & AAC_OUT_NBR6) ? A * M * 6 * 4 : 0);
    L.tmin = take(A * 16);
    L.tpair = take((flags & AAC_OUT_TCPA_PAIR) ? A * M * 16 : 0);
    L.rmin = take((flags & AAC_OUT_RADAR_AUX) ? A * R * 4 : 0);
    L.rhit = take((flags & AAC_OUT_RADAR_AUX) ? A * R * 2 : 0);
    L.parts = take((flags & AAC_OUT_PARTS) ? A * 8 * 4 : 0);
    L.total = o;
    return L;
}

// fills the __constant__ polygon tables of the current device (call once per device)
cudaError_t upload_constants();
cudaError_t launch_env_kernel(int variant, const KParams &p, int mode, int threads, cudaStream_t stream);
int max_smem_optin();

}  // namespace aac
