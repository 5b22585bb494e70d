// Batched multi-drone environment step for sm_100a: kinematics, neighbour ordering, radar, tdCPA,
// observation assembly, reward / done.  One CTA owns a tile of TE whole environments; every phase is
// a strided loop over that tile's work items (agents, ordered pairs, rays), results are staged in
// shared memory and leave the SM as coalesced row stores.
//
// Reference behaviour (file:line; ATT / V2 as in include/aac_env.h, UA / UV2 = Utilities_own_*.py):
//   kinematics            ATT:2639-2713, V2:3729-3787
//   neighbour order       ATT:758-773 (index order), V2:769-801 (distance-sorted, stable)
//   radar                 ATT:1052-1170 (other drones' 64-gons), V2:1210-1300 (grid cells + 4 bounds)
//   tdCPA                 UA:308-329 == UV2:337-358
//   observation layout    ATT:1285-1296,1357-1493; V2:1417-1429,1490-1713; NormalizeData UA:554-607
//   reward / done         ATT:2105-2618 (ss_reward), V2:2995-3684 (ss_reward_Mar, train mode)
// Geometry: every shapely "circle" is the regular 64-gon GEOS builds for Point.buffer(r);
// because all of them share vertex angles, polygon-polygon and polygon-square emptiness tests reduce
// to a support-function test over the 16 first-quadrant edge normals (Minkowski sum of two such
// polygons is again one, with the radii added).
#include <math_constants.h>

#include "aac_kernels.cuh"

namespace aac {

__constant__ float2 c_n16[16];  // edge normals of the 64-gon at (i + 0.5) * 5.625 deg, i = 0..15
__constant__ float2 c_n64[64];  // all 64 edge normals
__constant__ float c_apo;       // cos(pi / 64): apothem of the unit 64-gon

cudaError_t upload_constants() {
    float2 n16[16], n64[64];
    const double w = 2.0 * 3.14159265358979323846 / 64.0;
    for (int i = 0; i < 64; ++i) {
        n64[i].x = (float)cos(-(i + 0.5) * w);
        n64[i].y = (float)sin(-(i + 0.5) * w);
    }
    for (int i = 0; i < 16; ++i) {
        n16[i].x = (float)cos((i + 0.5) * w);
        n16[i].y = (float)sin((i + 0.5) * w);
    }
    const float apo = (float)cos(w / 2.0);
    cudaError_t e = cudaMemcpyToSymbol(c_n16, n16, sizeof(n16));
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_n64, n64, sizeof(n64));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_apo, &apo, sizeof(apo));
}

int max_smem_optin() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return v;
}

// ------------------------------------------------------------------------------------ helpers

constexpr unsigned F_DONE = 1u, F_GOAL = 2u, F_BRANCH_SHIFT = 2u, F_BBC3 = 32u, F_DOUBLE = 64u, F_ATGOAL = 128u;
constexpr unsigned M_REACH = 1u << 8, M_VBOUND = 1u << 9, M_VBLDG = 1u << 10, M_VDRONE = 1u << 11;

__device__ __forceinline__ float cell_cx(const MapDev &m, int ix) { return m.ex0 + (ix + 0.5f) * m.cell; }
__device__ __forceinline__ float cell_cy(const MapDev &m, int iy) { return m.ey0 + (iy + 0.5f) * m.cell; }

__device__ __forceinline__ bool occupied(const MapDev &m, int ix, int iy) {
    const int px = ix + MAP_PAD, py = iy + MAP_PAD;
    if (px < 0 || py < 0 || px >= m.pgx || py >= m.pgy) return false;
    const int b = px * m.pgy + py;
    return (m.bits[b >> 5] >> (b & 31)) & 1u;
}

// max over the 64 edge normals of n . (ax, ay) for ax, ay >= 0 (attained in the first quadrant)
__device__ __forceinline__ float support64_q1(float ax, float ay) {
    float m = -CUDART_INF_F;
#pragma unroll
    for (int i = 0; i < 16; ++i) m = fmaxf(m, fmaf(c_n16[i].x, ax, c_n16[i].y * ay));
    return m;
}

// closed 64-gon(centre p, r1) n 64-gon(centre g, r2) non-empty  (goal test, ATT:2266-2269)
__device__ __forceinline__ bool gons_touch(float qx, float qy, float rsum) {
    const float d2 = qx * qx + qy * qy;
    if (d2 > rsum * rsum) return false;
    const float apo = rsum * c_apo;
    if (d2 <= apo * apo) return true;
    return support64_q1(fabsf(qx), fabsf(qy)) - apo <= 0.0f;
}

// closed 64-gon(centre p, r) n axis-aligned square(centre q, half h) non-empty  (ATT:2243-2250)
__device__ __forceinline__ bool gon_square_touch(float dxc, float dyc, float h, float r) {
    const float ex = fabsf(dxc) - h, ey = fabsf(dyc) - h;
    if (ex > r || ey > r) return false;
    if (ex <= 0.0f || ey <= 0.0f) return true;
    return support64_q1(ex, ey) - r * c_apo <= 0.0f;
}

// max over i = 0..32 of cos(t0 - i * pi/32): x-extent of one end cap of the GEOS round buffer
__device__ __forceinline__ float cap_extent(float t0) {
    const float two_pi = 6.283185307179586f, pi = 3.14159265358979f, q = 0.09817477042468103f;
    float u = t0 - two_pi * floorf(t0 * (1.0f / two_pi));
    if (u <= pi) {
        const float rem = u - q * floorf(u * (1.0f / q));
        return cosf(fminf(rem, q - rem));
    }
    return fabsf(cosf(u));
}

// LineString([p0, p1]).buffer(r) vertex bounding box against the 4 boundary lines (ATT:2172-2173,
// :2507).  Local frame: the lines are x = -hx, hx and y = -hy, hy.
__device__ bool capsule_hits_bound(float x0, float y0, float x1, float y1, float r, float hx, float hy) {
    const float lox = fminf(x0, x1) - r, hix = fmaxf(x0, x1) + r, loy = fminf(y0, y1) - r, hiy = fmaxf(y0, y1) + r;
    if (lox > -hx && hix < hx && loy > -hy && hiy < hy) return false;  // cannot reach any line
    float mnx, mxx, mny, mxy;
    if (x0 == x1 && y0 == y1) {  // GEOS drops the repeated point: plain 64-gon, vertices on the axes
        mnx = x0 - r; mxx = x0 + r; mny = y0 - r; mxy = y0 + r;
    } else {
        const float ang = atan2f(y1 - y0, x1 - x0), hp = 1.5707963267948966f;
        const float e0 = cap_extent(ang), e1 = cap_extent(ang + hp), e2 = cap_extent(ang + 2.0f * hp), e3 = cap_extent(ang + 3.0f * hp);
        mxx = fmaxf(x1 + r * e1, x0 + r * e3);
        mnx = fminf(x1 - r * e3, x0 - r * e1);
        mxy = fmaxf(y1 + r * e0, y0 + r * e2);
        mny = fminf(y1 - r * e2, y0 - r * e0);
    }
    return (mnx <= -hx && -hx <= mxx) || (mnx <= hx && hx <= mxx) || (mny <= -hy && -hy <= mxy) || (mny <= hy && hy <= mxy);
}

// UA:308-329: returns tcpa, d_tcpa and whether the pair counts as a potential conflict
__device__ __forceinline__ void tcpa_dcpa(float hpx, float hpy, float hvx, float hvy, float opx, float opy, float ovx, float ovy,
                                          float bound_sum, float &tcpa, float &d, bool &conf) {
    const float rx = hpx - opx, ry = hpy - opy;
    const float wx = ovx - hvx, wy = ovy - hvy;
    const float w2 = wx * wx + wy * wy;
    conf = false;
    if (w2 == 0.0f) {
        tcpa = -10.0f;
        d = sqrtf(rx * rx + ry * ry);  // both advance by the same velocity: separation is unchanged
        conf = d < bound_sum;
    } else {
        tcpa = (rx * wx + ry * wy) / w2;
        const float ex = fmaf(wx, tcpa, -rx), ey = fmaf(wy, tcpa, -ry);
        d = sqrtf(ex * ex + ey * ey);
        conf = tcpa <= 1.0f && tcpa >= 0.0f && d < bound_sum;
    }
}

// UV2:31-44
__device__ __forceinline__ float bearing_deg(float xh, float yh, float xi, float yi) {
    const float th = atan2f(yi - yh, xi - xh) * 57.29577951308232f;
    return th < 0.0f ? -th : 360.0f - th;
}

__device__ __forceinline__ unsigned pick_scenario(long long gid, int episode, unsigned long long seed, int n) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(gid + 1) + 0xD1B54A32D192ED03ull * (unsigned long long)(episode + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (unsigned)(x % (unsigned long long)n);
}

// ------------------------------------------------------------------------------------ radar

// one ray against the occupied cells of the agent's 4x4 window and the 4 boundary lines
// (V2:1210-1300).  `sensed` = last hit in ascending cell order then L,R,B,T.
__device__ __forceinline__ void radar_grid_ray(const MapDev &mp, float px, float py, float2 dir, float ray_len, unsigned win,
                                               int wix0, int wiy0, bool last_hit, float &out, float &out_min, int &out_id) {
    const float dx = ray_len * dir.x, dy = ray_len * dir.y, len = ray_len;
    const float ex = px + dx, ey = py + dy;
    const float idx = dx != 0.0f ? 1.0f / dx : 0.0f, idy = dy != 0.0f ? 1.0f / dy : 0.0f;
    int r0 = (int)floorf((fminf(px, ex) - mp.ex0) * mp.inv_cell) - wix0, r1 = (int)floorf((fmaxf(px, ex) - mp.ex0) * mp.inv_cell) - wix0;
    int c0 = (int)floorf((fminf(py, ey) - mp.ey0) * mp.inv_cell) - wiy0, c1 = (int)floorf((fmaxf(py, ey) - mp.ey0) * mp.inv_cell) - wiy0;
    r0 = max(r0, 0); r1 = min(r1, 3); c0 = max(c0, 0); c1 = min(c1, 3);
    unsigned m = 0;
    if (r0 <= r1 && c0 <= c1) {
        const unsigned rows = (0xFFFFu >> (4 * (3 - r1))) & (0xFFFFu << (4 * r0));
        const unsigned cols = ((0xFu >> (3 - c1)) & (0xFu << c0)) * 0x1111u;
        m = win & rows & cols;
    }
    float shortest = CUDART_INF_F, sensed = len;
    int shortest_id = -1, sensed_id = -1;
    while (m) {
        const int b = __ffs(m) - 1;
        m &= m - 1;
        const int ix = wix0 + (b >> 2), iy = wiy0 + (b & 3);
        const float x0 = mp.ex0 + ix * mp.cell, x1 = x0 + mp.cell, y0 = mp.ey0 + iy * mp.cell, y1 = y0 + mp.cell;
        float ent = -CUDART_INF_F, ext = CUDART_INF_F;
        if (dx != 0.0f) {
            const float t0 = (x0 - px) * idx, t1 = (x1 - px) * idx;
            ent = fminf(t0, t1); ext = fmaxf(t0, t1);
        } else if (px < x0 || px > x1) continue;
        if (dy != 0.0f) {
            const float t0 = (y0 - py) * idy, t1 = (y1 - py) * idy;
            ent = fmaxf(ent, fminf(t0, t1)); ext = fminf(ext, fmaxf(t0, t1));
        } else if (py < y0 || py > y1) continue;
        const float lo = fmaxf(ent, 0.0f), hi = fminf(ext, 1.0f);
        if (lo > hi) continue;
        // nearest point of segment n cell BOUNDARY (V2:1258-1265): the entry point, or the exit point
        // when the drone centre is strictly inside the cell, or nothing (nan) when the whole ray is
        const bool inside = px > x0 && px < x1 && py > y0 && py < y1;
        const float d = !inside ? lo * len : (ext <= 1.0f ? ext * len : CUDART_NAN_F);
        const int id = ix * mp.gy + iy;
        sensed = d; sensed_id = id;
        if (d < shortest) { shortest = d; shortest_id = id; }
    }
    const int nb = mp.gx * mp.gy;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        const float dd = b < 2 ? dx : dy, pp = b < 2 ? px : py, inv = b < 2 ? idx : idy;
        const float lim = b < 2 ? mp.hx : mp.hy, line = (b & 1) ? lim : -lim;
        if (dd != 0.0f) {
            const float t = (line - pp) * inv;
            if (t >= 0.0f && t <= 1.0f) {
                const float d = t * len;
                sensed = d; sensed_id = nb + b;
                if (d < shortest) { shortest = d; shortest_id = nb + b; }
            }
        } else if (pp == line) {  // ray runs along the boundary: GEOS returns the whole overlap
            sensed = len; sensed_id = nb + b;
            if (0.0f < shortest) { shortest = 0.0f; shortest_id = nb + b; }
        }
    }
    out_min = shortest == CUDART_INF_F ? len : shortest;
    if (last_hit) { out = sensed; out_id = sensed_id; }
    else { out = out_min; out_id = shortest_id; }
}

// one ray against the other drones' protective 64-gons (ATT:1052-1170): entry distance, 0 inside
__device__ __forceinline__ void radar_drones_ray(const float *s_px, const float *s_py, int env_base, int N, int i, float2 dir,
                                                 float ray_len, float r, int id_base, float &out, int &out_id) {
    const float dx = ray_len * dir.x, dy = ray_len * dir.y, len = ray_len;
    const float px = s_px[env_base + i], py = s_py[env_base + i];
    const float apo = r * c_apo;
    float best = len, shortest = CUDART_INF_F;
    int best_id = -1;
    for (int j = 0; j < N; ++j) {
        if (j == i) continue;
        const float qx = px - s_px[env_base + j], qy = py - s_py[env_base + j];
        // distance from the polygon centre to the segment; beyond r the ray cannot touch it
        float tt = -(qx * dx + qy * dy) / (len * len);
        tt = fminf(fmaxf(tt, 0.0f), 1.0f);
        const float cx = fmaf(tt, dx, qx), cy = fmaf(tt, dy, qy);
        if (cx * cx + cy * cy > r * r * 1.00001f) continue;
        float lo = 0.0f, hi = 1.0f;
        bool ok = true;
        for (int e = 0; e < 64; ++e) {  // clip q + t*d to every half-plane n_e . x <= apothem
            const float f0 = apo - fmaf(c_n64[e].x, qx, c_n64[e].y * qy);
            const float f1 = -fmaf(c_n64[e].x, dx, c_n64[e].y * dy);
            if (f1 == 0.0f) { if (f0 < 0.0f) { ok = false; break; } continue; }
            const float t = -f0 / f1;
            if (f1 > 0.0f) lo = fmaxf(lo, t); else hi = fminf(hi, t);
            if (lo > hi) { ok = false; break; }
        }
        if (!ok) continue;
        const float d = lo * len;
        if (d < shortest) { shortest = d; best = d; best_id = id_base + j; }
    }
    out = best;
    out_id = best_id;
}

// ------------------------------------------------------------------------------------ kernel

template <int VAR>
__global__ void __launch_bounds__(MAX_THREADS) env_kernel(const __grid_constant__ KParams p, const int mode) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout &L = p.L;
    const int tid = threadIdx.x, nt = blockDim.x;
    const int N = p.N, M = N - 1, R = p.R, W = p.W;
    const int D = own_dim(VAR, N);
    const int e0 = blockIdx.x * p.TE;
    const int ne = min(p.TE, p.E - e0);
    const int A = ne * N, a0 = e0 * N;
    const int flags = p.out_flags;

    MapDev *s_maps = reinterpret_cast<MapDev *>(smem + L.maps);
    float2 *s_ray = reinterpret_cast<float2 *>(smem + L.ray);
    int4 *s_envi = reinterpret_cast<int4 *>(smem + L.envi);  // x active, y scenario, z map row, w result bits
    float *s_px = reinterpret_cast<float *>(smem + L.cur), *s_py = s_px + p.TE * N, *s_vx = s_py + p.TE * N, *s_vy = s_vx + p.TE * N,
          *s_hd = s_vy + p.TE * N;
    float *s_ppx = reinterpret_cast<float *>(smem + L.pre), *s_ppy = s_ppx + p.TE * N, *s_pvx = s_ppy + p.TE * N, *s_pvy = s_pvx + p.TE * N;
    unsigned *s_meta = reinterpret_cast<unsigned *>(smem + L.meta), *s_meta2 = reinterpret_cast<unsigned *>(smem + L.meta2);
    uint16_t *s_cells = reinterpret_cast<uint16_t *>(smem + L.cells);
    uint8_t *s_refw = smem + L.refw;
    float *s_d2 = reinterpret_cast<float *>(smem + L.d2);
    uint8_t *s_order = smem + L.order;
    uint2 *s_win = reinterpret_cast<uint2 *>(smem + L.win);
    float2 *s_tc = reinterpret_cast<float2 *>(smem + L.tc);
    uint8_t *s_pflag = smem + L.pflag;
    unsigned *s_agf = reinterpret_cast<unsigned *>(smem + L.agf);
    float *s_agr = reinterpret_cast<float *>(smem + L.agr);
    float *s_own = reinterpret_cast<float *>(smem + L.own), *s_nbr = reinterpret_cast<float *>(smem + L.nbr);
    float *s_radar = reinterpret_cast<float *>(smem + L.radar), *s_nbr6 = reinterpret_cast<float *>(smem + L.nbr6);
    float *s_raw_own = reinterpret_cast<float *>(smem + L.raw_own), *s_raw_nbr = reinterpret_cast<float *>(smem + L.raw_nbr);
    float *s_raw_nbr6 = reinterpret_cast<float *>(smem + L.raw_nbr6);
    float4 *s_tmin = reinterpret_cast<float4 *>(smem + L.tmin), *s_tpair = reinterpret_cast<float4 *>(smem + L.tpair);
    float *s_rmin = reinterpret_cast<float *>(smem + L.rmin);
    int16_t *s_rhit = reinterpret_cast<int16_t *>(smem + L.rhit);
    float *s_parts = reinterpret_cast<float *>(smem + L.parts);

    // ---- phase 0: which envs of the tile take part; scenario / map row per env; ray table
    bool any_active = false;
    for (int e = tid; e < ne; e += nt) {
        const int ge = e0 + e;
        int active = 1, scen = 0, map_row = 0;
        if (mode == MODE_RESET) {
            active = p.mask ? (p.mask[ge] != 0) : 1;
            if (active) {
                scen = (int)pick_scenario(p.env_id_base + ge, p.st.ep_index[ge], p.seed, p.n_scen);
                map_row = p.bank_map ? p.bank_map[scen] : 0;
                if (p.st.map_id) p.st.map_id[ge] = map_row;
            }
        } else if (p.st.map_id) {
            map_row = p.st.map_id[ge];
        }
        s_envi[e] = make_int4(active, scen, map_row, 0);
        any_active |= active != 0;
    }
    if (mode == MODE_RESET && !__syncthreads_or(any_active)) return;
    for (int k = tid; k < R; k += nt) s_ray[k] = p.ray_dir[k];
    __syncthreads();
    {   // stage the map(s) this tile uses (one shared map unless VAR == MM)
        const int nmap = VAR == AAC_VARIANT_MM ? ne : 1;
        constexpr int WORDS = sizeof(MapDev) / 4;
        for (int k = tid; k < nmap * WORDS; k += nt) {
            const int mi = k / WORDS, w = k - mi * WORDS;
            const int row = VAR == AAC_VARIANT_MM ? s_envi[mi].z : 0;
            reinterpret_cast<unsigned *>(s_maps)[k] = reinterpret_cast<const unsigned *>(p.maps + row)[w];
        }
    }
    __syncthreads();

    // ---- phase A: load (or re-initialise) the per-drone records, integrate the action
    for (int a = tid; a < A; a += nt) {
        const int ga = a0 + a, e = a / N, i = a - e * N;
        const MapDev &mp = s_maps[VAR == AAC_VARIANT_MM ? e : 0];
        float px, py, vx, vy, hd;
        unsigned meta;
        if (mode == MODE_RESET && s_envi[e].x) {
            const size_t src = ((size_t)s_envi[e].y * N + i) * W;
            for (int k = 0; k < W; ++k) {
                const uint16_t c = p.bank_cells[src + k];
                p.st.ref_cells[(size_t)ga * W + k] = c;
                s_cells[a * W + k] = c;
            }
            const int w = p.bank_w[(size_t)s_envi[e].y * N + i];
            p.st.ref_w[ga] = (uint8_t)w;
            s_refw[a] = (uint8_t)w;
            const uint16_t c0 = s_cells[a * W], c1 = s_cells[a * W + 1];
            px = cell_cx(mp, c0 >> 8); py = cell_cy(mp, c0 & 255);
            vx = 0.0f; vy = 0.0f;
            hd = atan2f(cell_cy(mp, c1 & 255) - py, cell_cx(mp, c1 >> 8) - px);  // ATT:359
            meta = 0xFFFF0000u;
            if (p.st.wall_count) p.st.wall_count[ga] = 0;
        } else {
            px = p.st.px[ga]; py = p.st.py[ga]; vx = p.st.vx[ga]; vy = p.st.vy[ga]; hd = p.st.heading[ga];
            meta = p.st.meta[ga];
            const int w = p.st.ref_w[ga];
            s_refw[a] = (uint8_t)w;
            for (int k = 0; k < w; ++k) s_cells[a * W + k] = p.st.ref_cells[(size_t)ga * W + k];
        }
        s_ppx[a] = px; s_ppy[a] = py; s_pvx[a] = vx; s_pvy[a] = vy;
        if (mode == MODE_STEP) {  // ATT:2655-2713
            const float2 act = reinterpret_cast<const float2 *>(p.actions)[ga];
            const float cvx = fmaf(act.x * p.acc_max, p.dt, vx), cvy = fmaf(act.y * p.acc_max, p.dt, vy);
            const float sp = sqrtf(cvx * cvx + cvy * cvy);
            if (sp >= p.vmax) { const float s = p.vmax / sp; vx = cvx * s; vy = cvy * s; }
            else { vx = cvx; vy = cvy; }
            float ddx = vx * p.dt, ddy = vy * p.dt;
            if (VAR == AAC_VARIANT_V2) {
                if (meta & M_REACH) { ddx = 0.0f; ddy = 0.0f; }  // V2:3770-3775
                hd = atan2f(ddy, ddx);                            // V2:3783
            }
            px += ddx; py += ddy;
        }
        s_px[a] = px; s_py[a] = py; s_vx[a] = vx; s_vy[a] = vy; s_hd[a] = hd;
        s_meta[a] = meta;
    }
    __syncthreads();

    // ---- phase B1: squared distance matrix, occupancy window per drone
    for (int w = tid; w < A * N; w += nt) {
        const int a = w / N, j = w - a * N, e = a / N, i = a - e * N;
        const float dx = s_px[e * N + j] - s_px[a], dy = s_py[e * N + j] - s_py[a];
        s_d2[w] = j == i ? CUDART_INF_F : dx * dx + dy * dy;
    }
    for (int a = tid; a < A; a += nt) {
        const int e = a / N;
        const MapDev &mp = s_maps[VAR == AAC_VARIANT_MM ? e : 0];
        int ix0 = (int)floorf((s_px[a] - p.ray_len - mp.ex0) * mp.inv_cell);
        int iy0 = (int)floorf((s_py[a] - p.ray_len - mp.ey0) * mp.inv_cell);
        ix0 = min(max(ix0, -MAP_PAD), mp.gx + MAP_PAD - 4);
        iy0 = min(max(iy0, -MAP_PAD), mp.gy + MAP_PAD - 4);
        unsigned mask = 0;
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int b = (ix0 + r + MAP_PAD) * mp.pgy + iy0 + MAP_PAD;
            const unsigned lo = mp.bits[b >> 5], hi = mp.bits[min((b >> 5) + 1, MAP_WORDS - 1)];
            mask |= (__funnelshift_r(lo, hi, b & 31) & 0xFu) << (4 * r);
        }
        s_win[a] = make_uint2(mask, (unsigned)(ix0 & 0xFFFF) | ((unsigned)iy0 << 16));
    }
    __syncthreads();

    // ---- phase B2: neighbour iteration order (ATT: index order; V2: stable sort by distance)
    for (int w = tid; w < A * N; w += nt) {
        const int a = w / N, j = w - a * N, e = a / N, i = a - e * N;
        if (j == i) continue;
        int rank;
        if (VAR == AAC_VARIANT_V2) {
            const float dj = s_d2[w];
            rank = 0;
            for (int q = 0; q < N; ++q) {
                const float dq = s_d2[a * N + q];
                rank += (dq < dj) || (dq == dj && q < j);  // the self entry is +inf and never counts
            }
        } else {
            rank = j < i ? j : j - 1;
        }
        s_order[a * M + rank] = (uint8_t)j;
    }
    __syncthreads();

    // ---- phase C: ordered pairs -> tdCPA (cur, pre) and the neighbour blocks of the observation
    const float inv_vmax = 1.0f / p.vmax;
    for (int w = tid; w < A * M; w += nt) {
        const int a = w / M, k = w - a * M, e = a / N;
        const MapDev &mp = s_maps[VAR == AAC_VARIANT_MM ? e : 0];
        const int b = e * N + s_order[w];
        const float px = s_px[a], py = s_py[a], ox = s_px[b], oy = s_py[b], ovx = s_vx[b], ovy = s_vy[b];
        const float dx = ox - px, dy = oy - py;
        float t1, d1, t2, d2;
        bool c1, c2;
        tcpa_dcpa(px, py, s_vx[a], s_vy[a], ox, oy, ovx, ovy, 2.0f * p.prot, t1, d1, c1);
        tcpa_dcpa(s_ppx[a], s_ppy[a], s_pvx[a], s_pvy[a], s_ppx[b], s_ppy[b], s_pvx[b], s_pvy[b], 2.0f * p.prot, t2, d2, c2);
        s_tc[w] = make_float2(t1, d1);
        s_pflag[w] = (uint8_t)((c1 ? 1 : 0) | (c2 ? 2 : 0));
        if (flags & AAC_OUT_TCPA_PAIR) s_tpair[w] = make_float4(t1, d1, t2, d2);
        const float ihx = 1.0f / mp.hx, ihy = 1.0f / mp.hy;
        if (VAR == AAC_VARIANT_ATT) {
            float *r = s_own + a * D + 6 + 4 * k;  // scale_pos applied to a delta (ATT:1374)
            r[0] = (dx - mp.xmin_g) * ihx - 1.0f; r[1] = (dy - mp.ymin_g) * ihy - 1.0f; r[2] = ovx * inv_vmax; r[3] = ovy * inv_vmax;
            if (flags & AAC_OUT_RAW) { float *q = s_raw_own + a * D + 6 + 4 * k; q[0] = dx; q[1] = dy; q[2] = ovx; q[3] = ovy; }
        } else if (VAR == AAC_VARIANT_V2) {
            float *r = s_nbr + a * 5 * M + 5 * k;  // host - neighbour, host heading (V2:1519,1571)
            r[0] = -dx * ihx; r[1] = -dy * ihy; r[2] = ovx * inv_vmax; r[3] = ovy * inv_vmax; r[4] = s_hd[a];
            if (flags & AAC_OUT_RAW) { float *q = s_raw_nbr + a * 5 * M + 5 * k; q[0] = dx; q[1] = dy; q[2] = ovx; q[3] = ovy; q[4] = s_hd[b]; }
        }
        if (flags & AAC_OUT_NBR6) {  // legacy block built from [px,py,vx,vy,2.5] (ATT:1396-1410)
            const float oxg = ox + mp.ox, oyg = oy + mp.oy;
            float *r = s_nbr6 + w * 6;
            r[0] = dx * 0.5f * ihx; r[1] = dy * 0.5f * ihy; r[2] = (ovy - oxg) * 0.5f * ihx; r[3] = (p.prot - oyg) * 0.5f * ihy;
            r[4] = ovx * inv_vmax; r[5] = ovy * inv_vmax;
            if (flags & AAC_OUT_RAW) { float *q = s_raw_nbr6 + w * 6; q[0] = dx; q[1] = dy; q[2] = ovy - oxg; q[3] = p.prot - oyg; q[4] = ovx; q[5] = ovy; }
        }
    }

    // ---- phase D: radar, one work item per (drone, ray)
    for (int w = tid; w < A * R; w += nt) {
        const int a = w / R, k = w - a * R, e = a / N, i = a - e * N;
        const MapDev &mp = s_maps[VAR == AAC_VARIANT_MM ? e : 0];
        float out, out_min;
        int id;
        if (VAR == AAC_VARIANT_ATT) {
            radar_drones_ray(s_px, s_py, e * N, N, i, s_ray[k], p.ray_len, p.prot, mp.gx * mp.gy + 4, out, id);
            out_min = out;
        } else {
            const uint2 wn = s_win[a];
            radar_grid_ray(mp, s_px[a], s_py[a], s_ray[k], p.ray_len, wn.x, (int)(short)(wn.y & 0xFFFF), (int)(short)(wn.y >> 16),
                           VAR == AAC_VARIANT_V2 && p.radar_mode == AAC_RADAR_LAST_HIT, out, out_min, id);
        }
        s_radar[w] = out;
        if (flags & AAC_OUT_RADAR_AUX) { s_rmin[w] = out_min; s_rhit[w] = (int16_t)id; }
    }

    // ---- phase E1: own block of the observation, goal contact
    for (int a = tid; a < A; a += nt) {
        const int e = a / N;
        const MapDev &mp = s_maps[VAR == AAC_VARIANT_MM ? e : 0];
        const int w = s_refw[a];
        const uint16_t cg = s_cells[a * W + w - 1];
        const float gx = cell_cx(mp, cg >> 8), gy = cell_cy(mp, cg & 255);
        const float px = s_px[a], py = s_py[a];
        const float ihx = 1.0f / mp.hx, ihy = 1.0f / mp.hy;
        float nvx = s_vx[a] * inv_vmax, nvy = s_vy[a] * inv_vmax;
        if (VAR == AAC_VARIANT_MM) { nvx = s_vx[a] * ihx; nvy = s_vy[a] * ihy; }
        else if (M > 0) {
            // `norm_vel` is re-bound inside the neighbour loop (ATT:1408, V2:1586) before the own block
            // is assembled (ATT:1463, V2:1672): the slot carries the LAST neighbour's velocity
            const int b = e * N + s_order[a * M + M - 1];
            nvx = s_vx[b] * inv_vmax; nvy = s_vy[b] * inv_vmax;
        }
        float *r = s_own + a * D;
        r[0] = px * ihx; r[1] = py * ihy; r[2] = nvx; r[3] = nvy; r[4] = gx * ihx - r[0]; r[5] = gy * ihy - r[1];
        if (VAR == AAC_VARIANT_V2) r[6] = s_hd[a];
        if (flags & AAC_OUT_RAW) {
            float *q = s_raw_own + a * D;
            q[0] = px + mp.ox; q[1] = py + mp.oy; q[2] = s_vx[a]; q[3] = s_vy[a]; q[4] = gx - px; q[5] = gy - py;
            if (VAR == AAC_VARIANT_V2) q[6] = s_hd[a];
        }
        s_agf[a] = gons_touch(gx - px, gy - py, p.prot + p.goal_r) ? F_ATGOAL : 0u;
        s_agr[a] = 0.0f;
    }
    __syncthreads();

    // ---- phase E2: reward / collision / goal per drone
    if (mode == MODE_STEP) {
        for (int a = tid; a < A; a += nt) {
            const int e = a / N, i = a - e * N, eb = e * N;
            const MapDev &mp = s_maps[VAR == AAC_VARIANT_MM ? e : 0];
            const float px = s_px[a], py = s_py[a];
            const int w = s_refw[a];
            unsigned meta = s_meta[a];
            const bool at_goal = s_agf[a] & F_ATGOAL;
            if (VAR == AAC_VARIANT_V2 && at_goal) meta |= M_REACH;  // top of the drone's iteration (V2:3025-3033)
            const bool reach_i = meta & M_REACH;
            // neighbour scan (ATT:2187-2236, V2:3092-3168)
            int nearest = -1, n_coll = 0, last_coll = -1, imm_key = -1, conf_cur = 0, conf_pre = 0;
            bool prev2 = false;
            float shortest = CUDART_INF_F, imm_tcpa = CUDART_INF_F, imm_d = CUDART_INF_F;
            const int pn0 = (meta >> 16) & 0xFF, pn1 = (meta >> 24) & 0xFF;
            for (int k = 0; k < M; ++k) {
                const int j = s_order[a * M + k];
                const float2 tc = s_tc[a * M + k];
                const unsigned pf = s_pflag[a * M + k];
                conf_cur += pf & 1; conf_pre += (pf >> 1) & 1;
                if (tc.x >= 0.0f && tc.x < imm_tcpa) { imm_tcpa = tc.x; imm_d = tc.y; imm_key = j; }
                else if (tc.x == -10.0f && tc.y < imm_tcpa) { imm_tcpa = tc.x; imm_d = tc.y; imm_key = j; }
                const float d = sqrtf(s_d2[a * N + j]);
                if (d < shortest) { shortest = d; nearest = j; }
                if (d <= 2.0f * p.prot) {
                    if (VAR == AAC_VARIANT_V2) {
                        // reach_target of drone j as drone i sees it: set in earlier steps, or earlier in
                        // this step's loop when j < i (V2:3160)
                        const bool reach_j = (s_meta[eb + j] & M_REACH) || (j < i && (s_agf[eb + j] & F_ATGOAL));
                        if (reach_j || reach_i) continue;
                        prev2 |= (j == pn0) || (M > 1 && j == pn1);  // V2:3170-3179
                    }
                    ++n_coll; last_coll = j;
                }
            }
            s_tmin[a] = make_float4(imm_tcpa, imm_d, (float)imm_key, (float)(conf_cur + 256 * conf_pre));
            if (VAR == AAC_VARIANT_V2 && n_coll > 0) meta |= M_VDRONE;
            // building contact: the protective 64-gon against the (at most 2x2) cells it can reach
            bool collide_building = false;
            {
                const int ixa = (int)floorf((px - p.prot - mp.ex0) * mp.inv_cell), ixb = (int)floorf((px + p.prot - mp.ex0) * mp.inv_cell);
                const int iya = (int)floorf((py - p.prot - mp.ey0) * mp.inv_cell), iyb = (int)floorf((py + p.prot - mp.ey0) * mp.inv_cell);
                for (int ix = ixa; ix <= ixb; ++ix)
                    for (int iy = iya; iy <= iyb; ++iy)
                        if (occupied(mp, ix, iy) && gon_square_touch(px - cell_cx(mp, ix), py - cell_cy(mp, iy), 0.5f * mp.cell, p.prot))
                            collide_building = true;
            }
            if (collide_building && VAR == AAC_VARIANT_V2) meta |= M_VBLDG;
            // waypoint (ATT:2297-2303)
            const int cur = meta & 0xFF;
            const uint16_t cw = s_cells[a * W + 1 + cur], cg = s_cells[a * W + w - 1];
            const float gx = cell_cx(mp, cg >> 8), gy = cell_cy(mp, cg & 255);
            const float wdx = px - cell_cx(mp, cw >> 8), wdy = py - cell_cy(mp, cw & 255);
            const bool wp_flag = sqrtf(wdx * wdx + wdy * wdy) < 5.0f;
            const float ppx = s_ppx[a], ppy = s_ppy[a];
            const bool hit_bound = capsule_hits_bound(ppx, ppy, px, py, p.prot, mp.hx, mp.hy);
            const float spd = sqrtf(s_vx[a] * s_vx[a] + s_vy[a] * s_vy[a]);
            const float after_hg = sqrtf((px - gx) * (px - gx) + (py - gy) * (py - gy));
            // nearest point on the reference line: first segment attaining the minimum (ATT:3203-3214)
            float cross_err = CUDART_INF_F, arc = 0.0f, run = 0.0f;
            {
                float ax = cell_cx(mp, s_cells[a * W] >> 8), ay = cell_cy(mp, s_cells[a * W] & 255);
                for (int k = 1; k < w; ++k) {
                    const uint16_t c = s_cells[a * W + k];
                    const float bx = cell_cx(mp, c >> 8), by = cell_cy(mp, c & 255);
                    const float sx = bx - ax, sy = by - ay, len2 = sx * sx + sy * sy;
                    float rr = len2 > 0.0f ? ((px - ax) * sx + (py - ay) * sy) / len2 : 0.0f;
                    rr = fminf(fmaxf(rr, 0.0f), 1.0f);
                    const float qx = fmaf(rr, sx, ax), qy = fmaf(rr, sy, ay);
                    const float d = sqrtf((px - qx) * (px - qx) + (py - qy) * (py - qy));
                    const float sl = sqrtf(len2);
                    if (d < cross_err) { cross_err = d; arc = fmaf(rr, sl, run); }
                    run += sl;
                    ax = bx; ay = by;
                }
            }
            float min_radar = CUDART_INF_F;
            for (int k = 0; k < R; ++k) { const float v = s_radar[a * R + k]; if (v < min_radar) min_radar = v; }
            float dist_to_goal, near_drone = 0.0f, near_bldg = 0.0f, small_step, rew = 0.0f;
            unsigned res = 0, branch;
            if (VAR == AAC_VARIANT_ATT) {
                // (|pre-g| - |pos-g|) / vmax without the cancellation: (a-b).(a+b) / (|a|+|b|)  (ATT:2319-2325)
                const float bx = ppx - gx, by = ppy - gy, cx = px - gx, cy = py - gy;
                const float before_hg = sqrtf(bx * bx + by * by);
                const float den = before_hg + after_hg;
                dist_to_goal = den > 0.0f ? ((bx - cx) * (bx + cx) + (by - cy) * (by + cy)) / den * inv_vmax : 0.0f;
                const float c_dr = 1.0f + (2.5f / (10.0f - 2.5f)), m_dr = (0.0f - 1.0f) / (10.0f - 2.5f);  // ATT:2420-2426
                for (int k = 0; k < M; ++k) {  // every in-band neighbour adds the NEAREST one's penalty (ATT:2430-2432)
                    const float d = sqrtf(s_d2[a * N + s_order[a * M + k]]);
                    if (d >= 2.5f && d <= 10.0f) near_drone += fmaf(m_dr, shortest, c_dr);
                }
                small_step = 0.0f;  // coefficient 0 (ATT:2438)
                if (hit_bound) { rew = -20.0f - small_step - near_bldg; res |= F_DONE; branch = 0; }
                else if (n_coll > 0) {
                    rew = -20.0f - small_step - near_drone; res |= F_DONE; branch = 2;
                    if (last_coll == nearest) res |= F_BBC3;
                } else if (at_goal) { res |= F_GOAL; meta |= M_REACH; rew = 20.0f; branch = 3; }
                else {
                    if (wp_flag && (w - 1 - cur) > 1) meta = (meta & ~0xFFu) | (unsigned)(cur + 1);  // ATT:2565-2566
                    rew = dist_to_goal - small_step - near_bldg - near_drone;                           // ATT:2576-2578
                    branch = 4;
                }
            } else {
                const float dist_left = cross_err + (run - arc);  // UV2:413-441
                dist_to_goal = 6.0f * (1.0f - dist_left / run);    // V2:3257-3268
                if (nearest >= 0 && shortest >= 2.5f && shortest <= 6.0f) {  // V2:3365-3386
                    const float c_dr = 1.0f + (2.5f / (6.0f - 2.5f)), m_dr = (0.0f - 1.0f) / (6.0f - 2.5f);
                    const float brg = bearing_deg(px, py, s_px[eb + nearest], s_py[eb + nearest]);
                    const float coef = (brg >= 90.0f && brg <= 180.0f) ? 20.0f : 10.0f;
                    near_drone = coef * fmaf(m_dr, shortest, c_dr);
                }
                const float thr = 0.5f * p.vmax;  // V2:3446-3453
                small_step = 5.0f * ((thr - fminf(fmaxf(spd, 0.0f), thr)) * (1.0f / thr));
                if (min_radar >= p.prot && min_radar <= 5.0f) near_bldg = 3.0f * fmaf((0.0f - 1.0f) / (5.0f - p.prot), min_radar, 2.0f);  // V2:3522-3539
                // crash rewards are filled in by the per-env pass below: the penalty doubles along the
                // drone loop (V2:3590-3594)
                if (hit_bound) { meta |= M_VBOUND; res |= F_DONE; branch = 0; }
                else if (collide_building) { res |= F_DONE; branch = 1; }
                else if (n_coll > 0) {
                    res |= F_DONE; branch = 2;
                    const float brg = bearing_deg(px, py, s_px[eb + last_coll], s_py[eb + last_coll]);
                    if (brg >= 90.0f && brg <= 180.0f) res |= F_DOUBLE;
                    if (prev2) res |= F_BBC3;
                } else if (at_goal) { res |= F_GOAL; rew = 20.0f; branch = 3; }
                else {
                    if (wp_flag && (w - 1 - cur) > 1) meta = (meta & ~0xFFu) | (unsigned)(cur + 1);
                    rew = dist_to_goal - small_step - near_bldg - near_drone;  // V2:3631-3633
                    branch = 4;
                }
            }
            if (collide_building && p.st.wall_count) p.st.wall_count[a0 + a] += 1;
            // s_agf is read by the other drones of this env for F_ATGOAL only; keep that bit
            atomicOr(&s_agf[a], res | (branch << F_BRANCH_SHIFT));
            s_agr[a] = rew;
            if (flags & AAC_OUT_PARTS) {
                float *q = s_parts + a * 8;
                q[0] = dist_to_goal; q[1] = near_drone; q[2] = near_bldg; q[3] = small_step; q[4] = cross_err; q[5] = after_hg;
                q[6] = min_radar; q[7] = shortest;
            }
            // other drones read s_meta[j] & M_REACH (the record as it stood before this step) in the loop
            // above: publish the new record only after every drone of the tile is through
            s_meta2[a] = meta;
        }
        __syncthreads();
        for (int a = tid; a < A; a += nt) s_meta[a] = s_meta2[a];
        __syncthreads();
    }

    // ---- phase E3: per env -- crash penalties, summed reward, bound_building_check, episode end
    for (int e = tid; e < ne; e += nt) {
        if (mode == MODE_RESET && !s_envi[e].x) continue;
        const int ge = e0 + e, eb = e * N;
        if (mode == MODE_STEP) {
            float cp = 20.0f, sum = 0.0f;
            unsigned bbc = 0, any_done = 0, all_reach = 1, n_reach = 0;
            for (int i = 0; i < N; ++i) {
                const unsigned f = s_agf[eb + i], br = (f >> F_BRANCH_SHIFT) & 7u;
                if (VAR == AAC_VARIANT_V2 && br <= 2) {
                    if (f & F_DOUBLE) cp *= 2.0f;
                    s_agr[eb + i] = -cp;
                }
                if (br <= 2) bbc |= 1u << br;
                if (f & F_BBC3) bbc |= 8u;
                any_done |= f & F_DONE;
                const unsigned reached = (s_meta[eb + i] & M_REACH) ? 1u : 0u;
                all_reach &= reached;
                n_reach += reached;
                sum += s_agr[eb + i];
            }
            if (p.sum_reward) {  // reward = [sum(reward)] * N (ATT:2602-2603)
                for (int i = 0; i < N; ++i) s_agr[eb + i] = sum;
                sum *= (float)N;
            }
            const int step = p.st.ep_step[ge] + 1;
            p.st.ep_step[ge] = step;
            const float ret = p.st.ep_return[ge] + sum;
            p.st.ep_return[ge] = ret;
            const unsigned term = (step > p.ep_len ? 1u : 0u) | (any_done ? 2u : 0u) | (all_reach ? 4u : 0u);
            reinterpret_cast<uchar4 *>(p.out.bbc)[ge] = make_uchar4(bbc & 1, (bbc >> 1) & 1, (bbc >> 2) & 1, (bbc >> 3) & 1);
            p.out.terminated[ge] = (uint8_t)term;
            if (term && p.stats) {  // episode statistics (ATT/ma_main:581-637)
                atomicAdd(p.stats + 0, 1.0);
                atomicAdd(p.stats + 1, (double)step);
                atomicAdd(p.stats + 2, (double)ret);
                if (bbc & 1) atomicAdd(p.stats + 3, 1.0);
                if (bbc & 2) atomicAdd(p.stats + 4, 1.0);
                if (bbc & 4) atomicAdd(p.stats + 5, 1.0);
                if (bbc & 8) atomicAdd(p.stats + 6, 1.0);
                if (all_reach) atomicAdd(p.stats + 7, 1.0);
                if (n_reach) atomicAdd(p.stats + 8, (double)n_reach);
                if (term == 1u) atomicAdd(p.stats + 9, 1.0);
            }
        } else {
            if (mode == MODE_RESET) {
                p.st.ep_step[ge] = 0;
                p.st.ep_return[ge] = 0.0f;
                p.st.ep_index[ge] += 1;
            }
        }
    }
    __syncthreads();

    // ---- phase F: state write-back and coalesced row stores of the staged outputs
    const bool partial = mode == MODE_RESET;
    for (int a = tid; a < A; a += nt) {
        const int e = a / N;
        if (partial && !s_envi[e].x) continue;
        const int ga = a0 + a;
        unsigned meta = s_meta[a];
        if (M > 0) meta = (meta & 0x0000FFFFu) | ((unsigned)s_order[a * M] << 16) | ((unsigned)(M > 1 ? s_order[a * M + 1] : 0xFF) << 24);
        p.st.px[ga] = s_px[a]; p.st.py[ga] = s_py[a]; p.st.vx[ga] = s_vx[a]; p.st.vy[ga] = s_vy[a]; p.st.heading[ga] = s_hd[a];
        p.st.meta[ga] = meta;
        if (mode == MODE_STEP) {
            const unsigned f = s_agf[a];
            p.out.reward[ga] = s_agr[a];
            p.out.done[ga] = (uint8_t)(f & F_DONE ? 1 : 0);
            p.out.check_goal[ga] = (uint8_t)(f & F_GOAL ? 1 : 0);
            reinterpret_cast<float4 *>(p.out.tcpa_min)[ga] = s_tmin[a];
            if (flags & AAC_OUT_PARTS) p.out.branch[ga] = (int8_t)((f >> F_BRANCH_SHIFT) & 7u);
        }
    }
    auto store_rows = [&](float *dst, const float *src, int per_agent) {
        if (!dst || per_agent == 0) return;
        float *g = dst + (size_t)a0 * per_agent;
        for (int k = tid; k < A * per_agent; k += nt) {
            if (partial && !s_envi[(k / per_agent) / N].x) continue;
            g[k] = src[k];
        }
    };
    store_rows(p.out.norm_own, s_own, D);
    if (VAR == AAC_VARIANT_V2) store_rows(p.out.norm_nbr, s_nbr, 5 * M);
    store_rows(p.out.radar, s_radar, R);
    if (flags & AAC_OUT_NBR6) store_rows(p.out.norm_nbr6, s_nbr6, 6 * M);
    if (flags & AAC_OUT_RAW) {
        store_rows(p.out.raw_own, s_raw_own, D);
        if (VAR == AAC_VARIANT_V2) store_rows(p.out.raw_nbr, s_raw_nbr, 5 * M);
        if (flags & AAC_OUT_NBR6) store_rows(p.out.raw_nbr6, s_raw_nbr6, 6 * M);
    }
    if (flags & AAC_OUT_TCPA_PAIR) {
        store_rows(p.out.tcpa_pair, reinterpret_cast<const float *>(s_tpair), 4 * M);
        int8_t *g = p.out.nbr_order + (size_t)a0 * M;
        for (int k = tid; k < A * M; k += nt) {
            if (partial && !s_envi[(k / M) / N].x) continue;
            g[k] = (int8_t)s_order[k];
        }
    }
    if (flags & AAC_OUT_RADAR_AUX) {
        store_rows(p.out.radar_min, s_rmin, R);
        int16_t *g = p.out.radar_hit + (size_t)a0 * R;
        for (int k = tid; k < A * R; k += nt) {
            if (partial && !s_envi[(k / R) / N].x) continue;
            g[k] = s_rhit[k];
        }
    }
    if ((flags & AAC_OUT_PARTS) && mode == MODE_STEP) store_rows(p.out.parts, s_parts, 8);
}

cudaError_t launch_env_kernel(int variant, const KParams &p, int mode, int threads, cudaStream_t stream) {
    const int grid = (p.E + p.TE - 1) / p.TE;
    if (grid <= 0) return cudaSuccess;
    void (*fn)(KParams, int) = nullptr;
    switch (variant) {
        case AAC_VARIANT_ATT: fn = env_kernel<AAC_VARIANT_ATT>; break;
        case AAC_VARIANT_V2: fn = env_kernel<AAC_VARIANT_V2>; break;
        default: return cudaErrorInvalidValue;
    }
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.L.total);
    if (e != cudaSuccess) return e;
    fn<<<grid, threads, p.L.total, stream>>>(p, mode);
    return cudaGetLastError();
}

}  // namespace aac
