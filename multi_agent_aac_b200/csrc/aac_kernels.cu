// Batched multi-drone environment step for sm_100a: kinematics, neighbour ordering, radar, tdCPA,
// observation assembly, reward / done, auto-reset.
//
// Execution model: WARP-AUTONOMOUS.  Each warp owns a group of G = 32 / N whole environments (G * N <= 32
// drones) and runs the entire pipeline for them with __syncwarp only; there is no CTA barrier on the
// step path, so warps of one CTA drift apart and the SM overlaps their latencies.
//   lane = drone     load + action integration, neighbour sort + 4x4 occupancy window, own block,
//                    reward / collision / goal
//   lane = item      ordered pairs (tdCPA + neighbour blocks) and rays (radar), flattened over the group
//   lane = env       crash-penalty chain, summed reward, flags, episode end
// Outputs leave as coalesced stores: radar / tdCPA items are item-major already; 5- and 6-float pair
// blocks and the own rows pass through a small per-warp staging buffer.  The only CTA-wide data are the
// map (occupancy bitmap + constants, one TMA bulk copy per CTA) and the ray table.
// Envs that terminate are re-initialised from the scenario bank by the same warp and the observation
// pipeline runs again over those envs only (fused auto-reset).
//
// Reference behaviour (file:line; ATT / V2 as in include/aac_env.h, UA / UV2 = Utilities_own_*.py):
//   kinematics            ATT:2639-2713, V2:3729-3787
//   neighbour order       ATT:758-773 (index order), V2:769-801 (distance-sorted, stable)
//   radar                 ATT:1052-1170 (other drones' 64-gons), V2:1210-1300 (grid cells + 4 bounds)
//   tdCPA                 UA:308-329 == UV2:337-358
//   observation layout    ATT:1285-1296,1357-1493; V2:1417-1429,1490-1713; NormalizeData UA:554-607
//   reward / done         ATT:2105-2618 (ss_reward), V2:2995-3684 (ss_reward_Mar, train mode)
//   reset                 ATT:301-372 (state after reset_world), episode rule ATT/ma_main:448-462
// Geometry: every shapely "circle" is the regular 64-gon GEOS builds for Point.buffer(r);
// because all of them share vertex angles, polygon-polygon and polygon-square emptiness tests reduce
// to a support-function test over the 16 first-quadrant edge normals (Minkowski sum of two such
// polygons is again one, with the radii added).
#include <type_traits>
#include <math_constants.h>

#include "aac_kernels.cuh"

namespace aac {

__constant__ float2 c_n16[16];  // edge normals of the 64-gon at (i + 0.5) * 5.625 deg, i = 0..15
__constant__ float2 c_n64[64];  // all 64 edge normals
__constant__ float c_apo;       // cos(pi / 64): apothem of the unit 64-gon
__device__ float2 g_n64[64];    // the same 64 normals in global memory: read with one edge per lane (a constant-bank read
                                // with 32 different addresses would be serialised)

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
    e = cudaMemcpyToSymbol(g_n64, n64, sizeof(n64));
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_apo, &apo, sizeof(apo));
}

int max_smem_optin() {
    int dev = 0, v = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return v;
}

// ------------------------------------------------------------------------------------ TMA / mbarrier

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy, completion counted in bytes on the mbarrier (bytes % 16 == 0, 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------ helpers

constexpr unsigned F_DONE = 1u, F_GOAL = 2u, F_BRANCH_SHIFT = 2u, F_BBC3 = 32u, F_DOUBLE = 64u;
constexpr unsigned M_REACH = 1u << 8, M_VBOUND = 1u << 9, M_VBLDG = 1u << 10, M_VDRONE = 1u << 11;
// (hi, lo) = divmod(index, n) advanced by a fixed stride without dividing again
struct Walk {
    int hi, lo, dhi, dlo, n;
    __device__ __forceinline__ Walk(int start, int stride, int n_) : n(n_) {
        hi = start / n_; lo = start - hi * n_;
        dhi = stride / n_; dlo = stride - dhi * n_;
    }
    __device__ __forceinline__ void next() {
        lo += dlo; hi += dhi;
        if (lo >= n) { lo -= n; ++hi; }
    }
};

__device__ __forceinline__ float cell_cx(const MapDev &m, int ix) { return m.ex0 + (ix + 0.5f) * m.cell; }
__device__ __forceinline__ float cell_cy(const MapDev &m, int iy) { return m.ey0 + (iy + 0.5f) * m.cell; }

// max over the 64 edge normals of n . (ax, ay) for ax, ay >= 0 (attained in the first quadrant)
__device__ __forceinline__ float support64_q1(float ax, float ay) {
    float m = -CUDART_INF_F;
#pragma unroll 2
    for (int i = 0; i < 16; ++i) m = fmaxf(m, fmaf(c_n16[i].x, ax, c_n16[i].y * ay));
    return m;
}

// closed 64-gon(centre p, r1) n 64-gon(centre g, r2) non-empty  (goal test, ATT:2266-2269)
__device__ __forceinline__ bool gons_touch(float qx, float qy, float rsum) {
    const float d2 = qx * qx + qy * qy;
    if (d2 > rsum * rsum) return false;
    const float apo = __fmul_rn(rsum, c_apo);   // never fused with the subtraction below: all instantiations decide alike
    if (d2 <= apo * apo) return true;
    return support64_q1(fabsf(qx), fabsf(qy)) - apo <= 0.0f;
}

// closed 64-gon(centre p, r) n axis-aligned square(centre q, half h) non-empty  (ATT:2243-2250)
__device__ __forceinline__ bool gon_square_touch(float dxc, float dyc, float h, float r) {
    const float ex = fabsf(dxc) - h, ey = fabsf(dyc) - h;
    if (ex > r || ey > r) return false;
    if (ex <= 0.0f || ey <= 0.0f) return true;
    return support64_q1(ex, ey) - __fmul_rn(r, c_apo) <= 0.0f;
}

// max over i = 0..32 of cos(t0 - i * pi/32): x-extent of one end cap of the GEOS round buffer
__device__ __forceinline__ float cap_extent(float t0) {
    const float two_pi = 6.283185307179586f, pi = 3.14159265358979f, q = 0.09817477042468103f;
    float u = t0 - two_pi * floorf(t0 * (1.0f / two_pi));
    // (the library cosf carries a large-argument reduction of several hundred instructions per call site: both
    // arguments here are small, a short series / the hardware cosine are exact to a few 1e-7)
    if (u <= pi) {
        const float rem = u - q * floorf(u * (1.0f / q));
        const float x = fminf(rem, q - rem), x2 = x * x;   // |x| <= pi / 64
        return fmaf(x2, fmaf(x2, 1.0f / 24.0f, -0.5f), 1.0f);
    }
    return fabsf(__cosf(u - pi));
}

// LineString([p0, p1]).buffer(r) vertex bounding box against the 4 boundary lines (ATT:2172-2173,
// :2507).  Local frame: the lines are x = -hx, hx and y = -hy, hy.
__device__ __noinline__ bool capsule_hits_bound_exact(float x0, float y0, float x1, float y1, float r, float hx, float hy) {
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
__device__ __forceinline__ bool capsule_hits_bound(float x0, float y0, float x1, float y1, float r, float hx, float hy) {
    const float lox = fminf(x0, x1) - r, hix = fmaxf(x0, x1) + r, loy = fminf(y0, y1) - r, hiy = fmaxf(y0, y1) + r;
    if (lox > -hx && hix < hx && loy > -hy && hiy < hy) return false;  // cannot reach any line
    return capsule_hits_bound_exact(x0, y0, x1, y1, r, hx, hy);
}

// UA:308-329: returns tcpa, d_tcpa and whether the pair counts as a potential conflict
__device__ __forceinline__ void tcpa_dcpa(float hpx, float hpy, float hvx, float hvy, float opx, float opy, float ovx, float ovy,
                                          float bound_sum, float &tcpa, float &d, bool &conf) {
    const float rx = hpx - opx, ry = hpy - opy;
    const float wx = ovx - hvx, wy = ovy - hvy;
    const float w2 = wx * wx + wy * wy;
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
__device__ __noinline__ float bearing_deg(float xh, float yh, float xi, float yi) {
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

// order-preserving map float <-> unsigned (for redux.sync min / max over float values)
__device__ __forceinline__ unsigned f2ord(float f) {
    const unsigned b = __float_as_uint(f);
    return b ^ ((unsigned)((int)b >> 31) | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned k) { return __uint_as_float(k ^ ((k & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu)); }

// one ray against the other drones' protective 64-gons (ATT:1052-1170): entry distance, 0 inside.  Called by all 32
// lanes together (one ray each).  Few rays pass near another drone, so when at most 12 lanes need the 64 half-plane
// clip of a neighbour the warp does them one after the other with the edges spread over the lanes (two per lane, the
// running max / min by redux.sync - max and min do not depend on the order, the result is the serial loop's); with more
// lanes in need every lane clips its own ray against the 64 edges in turn.
__device__ __forceinline__ void radar_drones_ray(const float *s_px, const float *s_py, int env_base, int N, int i, float4 ray, float len,
                                                 float r, int id_base, float &out, int &out_id) {
    constexpr unsigned ALL = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const float dx = ray.x, dy = ray.y;
    const float px = s_px[env_base + i], py = s_py[env_base + i];
    // (__fmul_rn: the product must not be re-formed next to the subtraction below and fused with it in one instantiation and
    //  not in another - every instantiation returns the same bits)
    const float apo = __fmul_rn(r, c_apo), inv_l2 = 1.0f / (len * len);
    float best = len, shortest = CUDART_INF_F;
    int best_id = -1;
    for (int j = 0; j < N; ++j) {
        const float qx = px - s_px[env_base + j], qy = py - s_py[env_base + j];
        // distance from the polygon centre to the segment; beyond r the ray cannot touch it
        float tt = -(qx * dx + qy * dy) * inv_l2;
        tt = fminf(fmaxf(tt, 0.0f), 1.0f);
        const float cx = fmaf(tt, dx, qx), cy = fmaf(tt, dy, qy);
        const bool near = j != i && !(cx * cx + cy * cy > r * r * 1.00001f);
        const unsigned need = __ballot_sync(ALL, near);
        if (!need) continue;
        float lo = 0.0f, hi = 1.0f;
        bool ok = true;
        if (__popc(need) > 12) {
            if (near)
                for (int e = 0; e < 64; ++e) {  // clip q + t*d to every half-plane n_e . x <= apothem
                    const float f0 = apo - fmaf(c_n64[e].x, qx, c_n64[e].y * qy);
                    const float f1 = -fmaf(c_n64[e].x, dx, c_n64[e].y * dy);
                    if (f1 == 0.0f) { if (f0 < 0.0f) { ok = false; break; } continue; }
                    const float t = -f0 / f1;
                    if (f1 > 0.0f) lo = fmaxf(lo, t); else hi = fminf(hi, t);
                    if (lo > hi) { ok = false; break; }
                }
        } else {
            const float2 n0 = g_n64[lane], n1 = g_n64[lane + 32];
#pragma unroll 1
            for (unsigned rest = need; rest; rest &= rest - 1) {
                const int src = __ffs(rest) - 1;
                const float bqx = __shfl_sync(ALL, qx, src), bqy = __shfl_sync(ALL, qy, src);
                const float bdx = __shfl_sync(ALL, dx, src), bdy = __shfl_sync(ALL, dy, src);
                float l_lo = 0.0f, l_hi = 1.0f;
                bool bad = false;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const float2 n = h ? n1 : n0;
                    const float f0 = apo - fmaf(n.x, bqx, n.y * bqy);
                    const float f1 = -fmaf(n.x, bdx, n.y * bdy);
                    if (f1 == 0.0f) bad |= f0 < 0.0f;
                    else {
                        const float t = -f0 / f1;
                        if (f1 > 0.0f) l_lo = fmaxf(l_lo, t); else l_hi = fminf(l_hi, t);
                    }
                }
                const float w_lo = ord2f(__reduce_max_sync(ALL, f2ord(l_lo))), w_hi = ord2f(__reduce_min_sync(ALL, f2ord(l_hi)));
                const bool w_bad = __ballot_sync(ALL, bad) != 0u;
                if (lane == src) { lo = w_lo; hi = w_hi; ok = !w_bad && !(w_lo > w_hi); }
            }
        }
        if (!near || !ok) continue;
        const float d = lo * len;
        if (d < shortest) { shortest = d; best = d; best_id = id_base + j; }
    }
    out = best;
    out_id = best_id;
}

// ------------------------------------------------------------------------------------ kernel

constexpr unsigned FULL = 0xFFFFFFFFu;
// The reset loop's read of a group's flag: an acquire.  (ld.acquire.gpu compiles to the load plus CCTL.IVALL, which empties the
// SM's L1 once per group; a relaxed load - defensible here, the reset loop loads nothing the step loop stored - was measured
// and is not faster: C3 0.2302 against 0.2301 ms.)
__device__ __forceinline__ unsigned long long ld_flag(const unsigned long long *ptr) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(unsigned long long *ptr, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(ptr), "l"(v) : "memory");
}

// ---- the later fork's sensor classes (CS:1379-1506; aac_env.h AAC_TARGET_*) ----------------------------------------
__device__ __forceinline__ float2 cloud_pos(const CloudDev &c, int k) {
    const float s = (float)min(k, c.n_stop) * c.travel;
    return make_float2(fmaf(c.dx, s, c.sx), fmaf(c.dy, s, c.sy));
}
// ray q + t d (t in [0, 1]) against the OUTLINE of the regular 64-gon of radius r centred at the origin: parameter of the
// nearest intersection point (entry when the ray starts outside, exit when it starts inside), +inf when there is none
// (CS:1436-1448: min over the points of line n polygon.boundary)
__device__ __noinline__ float ray_gon_outline(float qx, float qy, float dx, float dy, float r) {
    // the polygon lies inside its circle: a ray that stays farther than r from the centre cannot touch it
    float tt = -(qx * dx + qy * dy) / (dx * dx + dy * dy);
    tt = fminf(fmaxf(tt, 0.0f), 1.0f);
    const float cx = fmaf(tt, dx, qx), cy = fmaf(tt, dy, qy);
    if (cx * cx + cy * cy > r * r * 1.00001f) return CUDART_INF_F;
    const float apo = __fmul_rn(r, c_apo);
    float lo = 0.0f, hi = 1.0f;
    bool inside = true;
#pragma unroll 1
    for (int e = 0; e < 64; ++e) {   // clip to every half-plane n_e . x <= apothem
        const float f0 = apo - fmaf(c_n64[e].x, qx, c_n64[e].y * qy);
        const float f1 = -fmaf(c_n64[e].x, dx, c_n64[e].y * dy);
        inside = inside && f0 >= 0.0f;
        if (f1 == 0.0f) { if (f0 < 0.0f) return CUDART_INF_F; continue; }
        const float t = -f0 / f1;
        if (f1 > 0.0f) lo = fmaxf(lo, t); else hi = fminf(hi, t);
        if (lo > hi) return CUDART_INF_F;
    }
    if (!inside) return lo;
    return hi < 1.0f ? hi : CUDART_INF_F;
}

// per-warp shared-memory slice (32 drone slots) and the handles a warp needs
struct Warp {
    const MapDev *map;
    const float4 *ray;
    const DdaRay *dda;  // [R] walk constants of the rays (aac_radar.cuh)
    WalkRef walk;       // the walk table
    int lane;
    int e_lo, ng, a0, nA;   // first env of the group, envs / drones in it, global index of its first drone
    float *px, *py, *vx, *vy, *hd, *ppx, *ppy, *pvx, *pvy;
    unsigned *meta, *meta2, *minr, *agf, *wpm;   // wpm: multipleMap waypoint mask
    float *agr, *d2, *stg, *own, *raw_own;
    uint8_t *order, *atgoal, *refw, *rs, *amap;   // amap: map row per drone (multipleMap)
    uint2 *win;
    float4 *wrel;
    uint16_t *c8;       // [32][8] first vertices of the reference lines
};

// lanes hold PI floats each for the items [0, n_valid) of one warp iteration; the block leaves as
// n_valid * PI consecutive floats at dst
template <int PI>
__device__ __forceinline__ void flush_items(float *dst, int n_valid, const float (&v)[PI], float *stg, int lane) {
    if (lane < n_valid) {
#pragma unroll
        for (int c = 0; c < PI; ++c) stg[lane * PI + c] = v[c];
    }
    __syncwarp();
    if (n_valid == 32) {
#pragma unroll
        for (int c = 0; c < PI; ++c) dst[c * 32 + lane] = stg[c * 32 + lane];
    } else {
        for (int f = lane; f < n_valid * PI; f += 32) dst[f] = stg[f];
    }
    __syncwarp();
}

// observation pipeline for the drones [a_lo, a_lo + n_ag) of the warp's group (whole envs):
// neighbour order + window -> { pairs | rays | own block }.  `cells` = the lane's reference-line row.
// MASKED (a literal at the call sites of the mode-specialised launches): only the envs whose bit is set in `active` (bit g = env g of the group; a_lo is 0 then) are observed and stored - the
// re-initialised envs of a reset.  The drone-per-lane phases take all of them at once; the item-per-lane phases, whose stores
// are contiguous per env, run env by env.
template <int VAR, bool AUX, bool LEAN, int NT, int RT, int RM, bool CS>
__device__ __forceinline__ void observe_range(const KParams &p, const Warp &w, const int a_lo, const int n_ag, const uint16_t *cells, const bool tab,
                                              const bool step_mode, const bool MASKED, const unsigned active) {
    // step_mode: the observation closes a step (sensor configurations: the clouds have moved once more than ep_step says)
    const int radar_mode = RM < 0 ? p.radar_mode : RM;   // RM >= 0: the radar mode is a compile-time constant of the instantiation
    // tab: the drones of the range have just been reset, i.e. stand on cell centres, and the handle has a radar table:
    // their ranges are looked up (they are what this very code computes for that cell, see aac_set_radar_table)
    const int lane = w.lane;
    // NT / RT > 0: drone count / ray count known at compile time
    const int N = NT ? NT : p.N, M = N - 1, R = RT ? RT : p.R, Mp = M | 1;
    const int D = own_dim(VAR, N);
    const int flags = LEAN ? 0 : p.out_flags;   // LEAN: no optional output was requested, their code is compiled out
    // one map staged in shared memory, or (multipleMap) the env's own map out of the staged table (w.map = the table in
    // shared memory when it fits, else the table in global memory)
    auto map_of = [&](const int aa) -> const MapDev & { return VAR == AAC_VARIANT_MM ? w.map[w.amap[aa]] : *w.map; };
    const float inv_vmax = 1.0f / p.vmax;
    const int a = a_lo + lane;           // the lane's drone in the drone-per-lane phases
    const bool mine = lane < n_ag && (!MASKED || ((active >> (a / N)) & 1u));
    // the contiguous sub-ranges the item-per-lane phases run over: the whole range, or one env at a time
    auto for_ranges = [&](auto body) {
        const int n_r = MASKED ? n_ag / N : 1;
#pragma unroll 1
        for (int g = 0; g < n_r; ++g) {
            if (MASKED && !((active >> g) & 1u)) continue;
            body(MASKED ? g * N : a_lo, MASKED ? N : n_ag);
        }
    };
    const int eb = mine ? (a / N) * N : 0;  // first drone of its env
    const MapDev &mp = map_of(mine ? a : a_lo);

    // ---- neighbour iteration order (ATT: index order; V2: stable insertion sort by distance), window
    if (mine) {
        const int i = a - eb;
        const float px = w.px[a], py = w.py[a];
        float *dist = w.d2 + a * Mp;
        uint8_t *ord = w.order + a * M;
        if (VAR == AAC_VARIANT_V2 && NT > 1) {
            // Stable sort by distance (V2:769-801) with the drone count known at compile time: the keys live in
            // registers and go through a fixed compare-exchange network (no divergence between lanes).  A key is
            // the distance's bit pattern with the neighbour slot in the 5 low mantissa bits, so one unsigned
            // compare orders by (distance, index); distances closer than 32 ulp order by index.
            constexpr int MC = NT > 1 ? NT - 1 : 1;
            unsigned key[MC];
#pragma unroll
            for (int m = 0; m < MC; ++m) {
                const int j = m + (m >= i ? 1 : 0);
                const float dx = w.px[eb + j] - px, dy = w.py[eb + j] - py;
                key[m] = (__float_as_uint(dx * dx + dy * dy) & ~31u) | (unsigned)m;
            }
#pragma unroll
            for (int s_ = 1; s_ < MC; ++s_) {
#pragma unroll
                for (int q = s_; q > 0; --q) {
                    const unsigned lo = min(key[q - 1], key[q]), hi = max(key[q - 1], key[q]);
                    key[q - 1] = lo; key[q] = hi;
                }
            }
#pragma unroll
            for (int k = 0; k < MC; ++k) {
                const int m = key[k] & 31u;
                dist[k] = __uint_as_float(key[k] & ~31u);
                ord[k] = (uint8_t)(m + (m >= i ? 1 : 0));
            }
        } else {
            int m = 0;
            for (int j = 0; j < N && VAR != AAC_VARIANT_MM; ++j) {   // multipleMap has no neighbour terms on the path
                if (j == i) continue;
                const float dx = w.px[eb + j] - px, dy = w.py[eb + j] - py;
                const float d2 = dx * dx + dy * dy;
                int q = m;
                if (VAR == AAC_VARIANT_V2)
                    while (q > 0 && dist[q - 1] > d2) { dist[q] = dist[q - 1]; ord[q] = ord[q - 1]; --q; }
                dist[q] = d2; ord[q] = (uint8_t)j;
                ++m;
            }
        }
        if (VAR != AAC_VARIANT_ATT && tab) {
            const int row = VAR == AAC_VARIANT_MM ? w.amap[a] : 0;
            const int idx = row * MAP_STRIDE_CELLS + (int)floorf((px - mp.ex0) * mp.inv_cell) * mp.gy + (int)floorf((py - mp.ey0) * mp.inv_cell);
            w.win[a] = make_uint2(0u, (unsigned)idx);
            w.minr[a] = p.rtab_minr[idx];
        } else if (VAR != AAC_VARIANT_ATT) {   // (the one_model_att radar senses the other drones only: no occupancy window)
        // 5x5 occupancy window around the drone's own cell: every cell a ray of length ray_len can enter
        float ax, ay, dlx, dly;
        int ixc, iyc;
        const unsigned win = build_window5(mp, px, py, p.ray_len, ax, ay, dlx, dly, ixc, iyc);
        w.win[a] = make_uint2(win, (unsigned)(ixc & 0xFFFF) | ((unsigned)iyc << 16));
        w.wrel[a] = make_float4(ax, ay, dlx, dly);
        }
    }
    __syncwarp();

    // ---- ordered pairs -> tdCPA (cur, pre) and the neighbour blocks of the observation
    if (M > 0 && VAR != AAC_VARIANT_MM) for_ranges([&](const int a_lo, const int n_ag) {
        // Mo = neighbours per drone that enter the blocks: all of them, or (sensor configurations) the nearest n_nbr_obs
        const int Mo = (CS && p.n_nbr_obs > 0) ? min(p.n_nbr_obs, M) : M;
        const int n_items = n_ag * Mo;
        const size_t pg0 = (size_t)(w.a0 + a_lo) * Mo;  // global index of the range's first pair
        for (Walk it(lane, 32, Mo); it.hi * Mo + it.lo - lane < n_items; it.next()) {
            const int idx = it.hi * Mo + it.lo, base = idx - lane;   // base = first item of this warp iteration
            const int n_valid = min(32, n_items - base);
            const bool ok = idx < n_items;
            const int aa = a_lo + (ok ? it.hi : 0), k = ok ? it.lo : 0, pi = aa * M + k;
            const int ebb = (aa / N) * N, b = ebb + w.order[pi];
            const float px = w.px[aa], py = w.py[aa], ox = w.px[b], oy = w.py[b], ovx = w.vx[b], ovy = w.vy[b];
            const float dx = ox - px, dy = oy - py;
            if (ok && (flags & AAC_OUT_TCPA_PAIR)) {
                float t1, d1, t2, d2;
                bool c1, c2;
                tcpa_dcpa(px, py, w.vx[aa], w.vy[aa], ox, oy, ovx, ovy, 2.0f * p.prot, t1, d1, c1);
                tcpa_dcpa(w.ppx[aa], w.ppy[aa], w.pvx[aa], w.pvy[aa], w.ppx[b], w.ppy[b], w.pvx[b], w.pvy[b], 2.0f * p.prot, t2, d2, c2);
                reinterpret_cast<float4 *>(p.out.tcpa_pair)[pg0 + idx] = make_float4(t1, d1, t2, d2);
                p.out.nbr_order[pg0 + idx] = (int8_t)w.order[pi];
            }
            if (VAR == AAC_VARIANT_ATT) {
                if (ok) {
                    float *r = w.own + aa * D + 6 + 4 * k;  // scale_pos applied to a delta (ATT:1374)
                    r[0] = (dx - mp.xmin_g) * mp.ihx - 1.0f; r[1] = (dy - mp.ymin_g) * mp.ihy - 1.0f; r[2] = ovx * inv_vmax; r[3] = ovy * inv_vmax;
                    if (flags & AAC_OUT_RAW) { float *q = w.raw_own + aa * D + 6 + 4 * k; q[0] = dx; q[1] = dy; q[2] = ovx; q[3] = ovy; }
                }
            } else if (VAR == AAC_VARIANT_V2) {
                // host - neighbour, host heading (V2:1519,1571)
                const float v[5] = {-dx * mp.ihx, -dy * mp.ihy, ovx * inv_vmax, ovy * inv_vmax, w.hd[aa]};
                flush_items<5>(p.out.norm_nbr + (pg0 + base) * 5, n_valid, v, w.stg, lane);
                if (flags & AAC_OUT_RAW) {
                    const float q[5] = {dx, dy, ovx, ovy, w.hd[b]};
                    flush_items<5>(p.out.raw_nbr + (pg0 + base) * 5, n_valid, q, w.stg, lane);
                }
            }
            if (flags & AAC_OUT_NBR6) {  // legacy block built from [px,py,vx,vy,2.5] (ATT:1396-1410)
                const float oxg = ox + mp.ox, oyg = oy + mp.oy;
                const float v[6] = {dx * 0.5f * mp.ihx, dy * 0.5f * mp.ihy, (ovy - oxg) * 0.5f * mp.ihx, (p.prot - oyg) * 0.5f * mp.ihy,
                                    ovx * inv_vmax, ovy * inv_vmax};
                flush_items<6>(p.out.norm_nbr6 + (pg0 + base) * 6, n_valid, v, w.stg, lane);
                if (flags & AAC_OUT_RAW) {
                    const float q[6] = {dx, dy, ovy - oxg, p.prot - oyg, ovx, ovy};
                    flush_items<6>(p.out.raw_nbr6 + (pg0 + base) * 6, n_valid, q, w.stg, lane);
                }
            }
        }
    });

    // ---- radar.  A warp iteration covers 32 rays of ONE drone (window mask, bounds flag and position are
    //      then warp-uniform: no divergence in the cell loop); the R % 32 leftover rays of several drones
    //      are packed into shared iterations.  Ranges are >= 0, so their bit patterns order like unsigned
    //      integers and nan (0x7FC00000) sorts above every number: the per-drone minimum the reward needs
    //      (min_radar) is one redux.sync per drone and iteration.
    for_ranges([&](const int a_lo, const int n_ag) {
    if (VAR != AAC_VARIANT_ATT && tab) {
        const size_t rg0 = (size_t)(w.a0 + a_lo) * R;
        // twelve table reads in flight per lane (they come from L2): one round trip for the 360 ranges of a 10-drone env
        constexpr int TB = 12;
        for (int f0 = lane; f0 < n_ag * R; f0 += 32 * TB) {
            float v[TB];
            unsigned src[TB];   // (a table has fewer than 2^31 entries: 14 maps x 1024 cells x R)
#pragma unroll
            for (int u = 0; u < TB; ++u) {
                const int f = f0 + 32 * u, q = f / R, k = f - q * R;
                src[u] = f < n_ag * R ? w.win[a_lo + q].y * (unsigned)R + (unsigned)k : 0u;
                v[u] = __ldg(p.rtab + src[u]);
            }
#pragma unroll
            for (int u = 0; u < TB; ++u) {
                const int f = f0 + 32 * u;
                if (f < n_ag * R) {
                    p.out.radar[rg0 + f] = v[u];
                    if (AUX) { p.out.radar_min[rg0 + f] = p.rtab_min[src[u]]; p.out.radar_hit[rg0 + f] = p.rtab_hit[src[u]]; }
                }
            }
        }
    } else {
        const float len = p.ray_len;
        // One pass over the rays of the range for one radar flavour (WANT: aac_radar.cuh).  A warp iteration covers 32 rays
        // of ONE drone (full chunks) or `rem` leftover rays of 32 / rem drones; a lane casts the same ray for every drone,
        // so the ray's walk constants stay in registers across the drone loop.  Ranges are >= 0: their bit patterns order
        // like unsigned integers and nan (0x7FC00000) sorts above every number, so the per-drone minimum the reward needs
        // (min_radar) is a redux per iteration, kept by the lane whose index is the drone's and stored once at the end.
        const size_t rg0 = (size_t)(w.a0 + a_lo) * R;
        auto radar_pass = [&](auto want_c) {
            constexpr int WANT = decltype(want_c)::value;
            const bool last_hit = VAR == AAC_VARIANT_V2 && radar_mode == AAC_RADAR_LAST_HIT;
            auto cast = [&](const int aa, const int k, const DdaRay &dr, float &out_min, int &id) -> float {
                float out;
                if (VAR == AAC_VARIANT_ATT) {
                    const MapDev &mr = map_of(aa);
                    const int ebb = (aa / N) * N;
                    radar_drones_ray(w.px, w.py, ebb, N, aa - ebb, w.ray[k], len, p.prot, mr.gx * mr.gy + 4, out, id);
                    out_min = out;
                    return out;
                }
                const uint2 wn = w.win[aa];
                if (CS) {
                    // the later fork's radar: true minimum over the configured target classes, in ray-parameter space
                    const MapDev &mr = map_of(aa);
                    const float4 ray = w.ray[k];
                    const float px = w.px[aa], py = w.py[aa];
                    const int nb = mr.gx * mr.gy, ebb = (aa / N) * N;
                    float best = CUDART_INF_F;
                    id = -1;
                    if (p.radar_targets & AAC_TARGET_CELLS) {
                        bool ok_walk = false;
                        float o1 = len, m1 = len;
                        int id1 = -1;
                        if (!(wn.x & WIN5_SLOW)) {   // the walk with no boundary line in reach (the lines are a class of their own here)
                            const float4 rel = w.wrel[aa];
                            ok_walk = cast_grid_fast<AUX ? 3 : 1, AUX>(dr, WalkRef{0u, nullptr}, p.cell, make_float4(rel.x, rel.y, CUDART_INF_F, CUDART_INF_F),
                                                                     wn.x & ~WIN5_NEAR_BOUND, (int)(short)(wn.y & 0xFFFF), (int)(short)(wn.y >> 16), mr.gx, mr.gy, len, false, o1, m1, id1);
                        }
                        if (!ok_walk) {
                            const SlowCast sc = cast_grid_slow<AUX>(mr, ray, px, py, wn.x & ~WIN5_NEAR_BOUND, len, 0);
                            m1 = sc.out_min; id1 = sc.id;
                        }
                        if (m1 < len) { best = m1 / len; id = id1; }
                    }
                    if (p.radar_targets & AAC_TARGET_BOUNDS) {   // left, right, top, bottom segments (CS:676-680)
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            const bool vert = b < 2;
                            const float line = b == 0 ? -mr.hx : b == 1 ? mr.hx : b == 2 ? mr.hy : -mr.hy;
                            const float dd = vert ? ray.x : ray.y, pp = vert ? px : py, inv = vert ? ray.z : ray.w;
                            if (dd != 0.0f) {
                                const float t = (line - pp) * inv;
                                const float other = vert ? fmaf(t, ray.y, py) : fmaf(t, ray.x, px);
                                if (t >= 0.0f && t <= 1.0f && fabsf(other) <= (vert ? mr.hy : mr.hx) && t < best) { best = t; id = nb + b; }
                            }
                        }
                    }
                    if (p.radar_targets & AAC_TARGET_CLOUDS) {
                        const int kc = p.st.ep_step[w.e_lo + aa / N] + (step_mode ? 1 : 0);   // the clouds move at the top of step (CS:4667-4681)
                        for (int ci = 0; ci < p.n_clouds; ++ci) {
                            const float2 c = cloud_pos(p.clouds[ci], kc);
                            const float t = ray_gon_outline(px - c.x, py - c.y, ray.x, ray.y, p.clouds[ci].radius);
                            if (t < best) { best = t; id = nb + 4 + N + ci; }
                        }
                    }
                    if (p.radar_targets & AAC_TARGET_AIRCRAFT)
                        for (int j = 0; j < N; ++j) {
                            if (ebb + j == aa) continue;
                            const float t = ray_gon_outline(px - w.px[ebb + j], py - w.py[ebb + j], ray.x, ray.y, p.prot);
                            if (t < best) { best = t; id = nb + 4 + j; }
                        }
                    out = best == CUDART_INF_F ? len : best * len;
                    out_min = out;
                    return out;
                }
                bool done = false;
                if (!(wn.x & WIN5_SLOW)) {
                    int gx = 0, gy = 0;
                    if (AUX) { const MapDev &mr = map_of(aa); gx = mr.gx; gy = mr.gy; }
                    done = cast_grid_fast<WANT, AUX>(dr, WalkRef{0u, nullptr}, p.cell, w.wrel[aa], wn.x, (int)(short)(wn.y & 0xFFFF), (int)(short)(wn.y >> 16), gx, gy, len, last_hit, out, out_min, id);
                }
                if (!done) {
                    const SlowCast sc = cast_grid_slow<AUX>(map_of(aa), w.ray[k], w.px[aa], w.py[aa], wn.x, len, WANT == 3 ? (int)last_hit : (WANT == 2));
                    out = sc.out; out_min = sc.out_min; id = sc.id;
                }
                return out;
            };
            const int full = R >> 5, rem = R & 31;
            unsigned my_min = 0x7F800000u;   // lane q: the minimum of drone a_lo + q so far
#pragma unroll 1
            for (int c = 0; c < full; ++c) {
                const int k = (c << 5) + lane;
                DdaRay dr = w.dda[k];
                dr.quad += w.walk.s;   // the walk table's shared-memory address rides in the ray's quadrant offset
                float *g_row = p.out.radar + rg0 + k;
#pragma unroll 1
                for (int q = 0; q < n_ag; ++q, g_row += R) {
                    float out_min;
                    int id = -1;
                    const float out = cast(a_lo + q, k, dr, out_min, id);
                    *g_row = out;
                    if (AUX) { p.out.radar_min[rg0 + q * R + k] = out_min; p.out.radar_hit[rg0 + q * R + k] = (int16_t)id; }
                    const unsigned m = __reduce_min_sync(FULL, __float_as_uint(out));
                    if (lane == q) my_min = min(my_min, m);
                }
            }
            // leftover rays: `per` drones share an iteration, `rem` lanes each
            if (rem) {
                const int per = 32 / rem;
                const int sub = lane / rem, k = (full << 5) + lane - sub * rem;
                const int kr = k < R ? k : 0;
                DdaRay drr = w.dda[kr];
                drr.quad += w.walk.s;
#pragma unroll 1
                for (int q0 = 0; q0 < n_ag; q0 += per) {
                    const int nsub = min(per, n_ag - q0);
                    const bool ok = sub < nsub;
                    const int q = q0 + (ok ? sub : 0);
                    float out_min;
                    int id = -1;
                    const float out = cast(a_lo + q, kr, drr, out_min, id);
                    if (ok) {
                        p.out.radar[rg0 + q * R + k] = out;
                        if (AUX) { p.out.radar_min[rg0 + q * R + k] = out_min; p.out.radar_hit[rg0 + q * R + k] = (int16_t)id; }
                    }
                    const unsigned key = ok ? __float_as_uint(out) : 0xFFFFFFFFu;
                    if ((rem & (rem - 1)) == 0) {   // aligned power-of-two lane groups: butterfly inside each drone's group
                        unsigned m = key;
                        for (int d = 1; d < rem; d <<= 1) m = min(m, __shfl_xor_sync(FULL, m, d));
                        // lane q0 + s fetches the minimum of sub-group s
                        const int s_of = lane - q0;
                        const unsigned mine_m = __shfl_sync(FULL, m, (s_of >= 0 && s_of < per ? s_of : 0) * rem);
                        if (s_of >= 0 && s_of < nsub) my_min = min(my_min, mine_m);
                    } else {
#pragma unroll 1
                        for (int s2 = 0; s2 < nsub; ++s2) {
                            const unsigned m = __reduce_min_sync(FULL, sub == s2 ? key : 0xFFFFFFFFu);
                            if (lane == q0 + s2) my_min = min(my_min, m);
                        }
                    }
                }
            }
            if (lane < n_ag) w.minr[a_lo + lane] = my_min;
        };
        if (AUX) radar_pass(std::integral_constant<int, 3>{});
        else if (VAR == AAC_VARIANT_V2 && radar_mode == AAC_RADAR_LAST_HIT) radar_pass(std::integral_constant<int, 2>{});
        else radar_pass(std::integral_constant<int, 1>{});
    }
    });
    __syncwarp();

    // ---- own block of the observation, goal contact
    if (mine) {
        const int nw = w.refw[a];
        // agent.goal[-1]: the last vertex, or (multipleMap, where waypoints are popped from anywhere in the list,
        // MM:1757) the highest vertex still in the list
        const int glast = VAR == AAC_VARIANT_MM ? 31 - __clz(w.wpm[a] | 1u) : nw - 1;
        const unsigned cg = glast < 8 ? w.c8[a * 8 + glast] : cells[glast];
        const float gx = cell_cx(mp, cg >> 8), gy = cell_cy(mp, cg & 255);
        const float px = w.px[a], py = w.py[a];
        float nvx = w.vx[a] * inv_vmax, nvy = w.vy[a] * inv_vmax;
        if (VAR == AAC_VARIANT_MM) { nvx = w.vx[a] * mp.ihx; nvy = w.vy[a] * mp.ihy; }
        else if (M > 0) {
            // `norm_vel` is re-bound inside the neighbour loop (ATT:1408, V2:1586) before the own block
            // is assembled (ATT:1463, V2:1672): the slot carries the LAST neighbour's velocity
            const int b = eb + w.order[a * M + M - 1];
            nvx = w.vx[b] * inv_vmax; nvy = w.vy[b] * inv_vmax;
        }
        float *r = w.own + a * D;
        r[0] = px * mp.ihx; r[1] = py * mp.ihy; r[2] = nvx; r[3] = nvy; r[4] = gx * mp.ihx - r[0]; r[5] = gy * mp.ihy - r[1];
        if (VAR == AAC_VARIANT_V2) r[6] = w.hd[a];
        if (flags & AAC_OUT_RAW) {
            float *q = w.raw_own + a * D;
            q[0] = px + mp.ox; q[1] = py + mp.oy; q[2] = w.vx[a]; q[3] = w.vy[a]; q[4] = gx - px; q[5] = gy - py;
            if (VAR == AAC_VARIANT_V2) q[6] = w.hd[a];
        }
        w.atgoal[a] = gons_touch(gx - px, gy - py, p.prot + p.goal_r) ? 1 : 0;
        if (CS && p.out.cloud_contact) {   // the protective 64-gon overlaps a cloud's (polygons_single_cloud_conflict, CS:4099-4110)
            const int kc = p.st.ep_step[w.e_lo + a / N] + (step_mode ? 1 : 0);
            bool touch = false;
            for (int ci = 0; ci < p.n_clouds && (p.radar_targets & AAC_TARGET_CLOUDS); ++ci) {
                const float2 c = cloud_pos(p.clouds[ci], kc);
                touch = touch || gons_touch(c.x - px, c.y - py, p.prot + p.clouds[ci].radius);
            }
            p.out.cloud_contact[w.a0 + a] = touch ? 1 : 0;
        }
    }
    __syncwarp();
    for_ranges([&](const int a_lo, const int n_ag) {   // own rows of a range are contiguous: coalesced copy
        float *dst = p.out.norm_own + (size_t)(w.a0 + a_lo) * D;
        const float *src = w.own + a_lo * D;
        for (int f = lane; f < n_ag * D; f += 32) dst[f] = src[f];
        if (flags & AAC_OUT_RAW) {
            float *dr = p.out.raw_own + (size_t)(w.a0 + a_lo) * D;
            const float *sr = w.raw_own + a_lo * D;
            for (int f = lane; f < n_ag * D; f += 32) dr[f] = sr[f];
        }
    });
    __syncwarp();
}

// counter-based random draws for (env, episode, draw index): independent of how envs are sharded over GPUs.
// `episode_key` folds seed, global env id and episode index once; a draw is one multiply-xorshift finaliser on top.
__device__ __forceinline__ unsigned episode_key(unsigned long long seed, long long gid, int episode) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(gid + 1) + 0xD1B54A32D192ED03ull * (unsigned long long)(episode + 1);
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return (unsigned)(x >> 32) ^ (unsigned)x;
}
__device__ __forceinline__ unsigned draw(unsigned key, unsigned ctr) {
    unsigned h = key ^ (ctr * 0x9E3779B1u);
    h ^= h >> 16; h *= 0x7FEB352Du;
    h ^= h >> 15; h *= 0x846CA68Bu;
    h ^= h >> 16;
    return h;
}

// re-initialise the envs of the warp's group whose bit is set in `mask`, all at once (lane a = drone a of the group): what
// reset_world leaves behind (ATT:251-372).  Origins, destinations and reference lines come from the map's origin /
// destination table when one is installed (the device draws them with reset_world's rule, ATT:254-276), else from a
// pre-planned scenario of the bank.  Returns the lane's reference-line row (nullptr for a lane outside the masked envs).
// PL: the instantiation carries the per-episode path search for maps whose table has pools but no paths (aac_plan.cuh).
template <int VAR, int NT, bool PL>
__device__ __forceinline__ const uint16_t *init_envs(const KParams &p, const Warp &w, const unsigned mask) {
    const int lane = w.lane, N = NT ? NT : p.N, W = p.W;
    const int g = lane / N, i = lane - g * N;             // the lane's env of the group and its drone in it
    const bool act = lane < w.nA && ((mask >> g) & 1u);
    const int ge = w.e_lo + (act ? g : 0);
    const long long gid = p.env_id_base + ge;
    const int ep = p.st.ep_index[ge];
    const bool use_od = p.od != nullptr;
    const unsigned scen = use_od ? 0u : pick_scenario(gid, ep, p.seed, p.n_scen);
    // a map is drawn per episode (MM/ma_main:464)
    int map_row = 0;
    const unsigned key = use_od ? episode_key(p.seed, gid, ep) : 0u;
    if (VAR == AAC_VARIANT_MM) map_row = use_od ? (int)(draw(key, 0) % (unsigned)p.n_maps) : (p.bank_map ? p.bank_map[scen] : 0);
    const MapDev &mp = VAR == AAC_VARIANT_MM ? w.map[map_row] : *w.map;
    __syncwarp();                                          // every lane has read its env's episode index
    if (act && i == 0) {
        p.st.ep_index[ge] = ep + 1;
        p.st.ep_step[ge] = 0;
        p.st.ep_return[ge] = 0.0f;
        if (VAR == AAC_VARIANT_MM) p.st.map_id[ge] = map_row;
    }
    const uint16_t *row = nullptr;
    int nw = 0;
    uint4 head = make_uint4(0u, 0u, 0u, 0u);   // the first 8 vertices of the lane's reference line
    if (use_od) {
        const OdDev &od = p.od[map_row];
        // Lane (g, i) draws for drone i of env g.  reset_world draws drone by drone and redraws a start until it is more
        // than 2 * protectiveBound from every EARLIER drone's start (ATT:254-270): the candidates are drawn in parallel
        // (each drone has its own counter stream), the accept / redraw decisions run drone by drone - for all the masked
        // envs in lockstep - with the earlier drones' lanes of the same env voting.
        const float sep2 = 4.0f * p.prot * p.prot * mp.inv_cell * mp.inv_cell;   // (2 * protectiveBound)^2 in cells^2
        auto candidate = [&](const int di, const int attempt, int &s_idx, int &tq) -> int {
            const unsigned base = 1u + 256u * (unsigned)di + 4u * (unsigned)attempt;
            const int sq = draw(key, base) >> 30;
            tq = (int)(draw(key, base + 1) % 3u);
            if (tq >= sq) ++tq;                               // a different quadrant for the goal (ATT:256-258)
            const int n_s = od.pool_off[sq + 1] - od.pool_off[sq];
            s_idx = od.pool_off[sq] + (int)(draw(key, base + 2) % (unsigned)n_s);
            return od.cell_code[s_idx];
        };
        int s_idx = 0, tq = 0, code = 0;
        if (act) code = candidate(i, 0, s_idx, tq);
        // Starts sit on cell centres: while the separation is below one cell (2 * protectiveBound < grid length, the
        // reference's 5 m on a 10 m grid) only two drones on the SAME cell are too close, and when no two first draws of
        // an env coincide - most episodes - every drone keeps its draw and the drone-by-drone loop below has nothing to do
        const unsigned same = __match_any_sync(FULL, act ? ((g << 16) | code) : (0x40000000 | lane));
        const bool distinct = __all_sync(FULL, !act || (sep2 < 1.0f && same == (1u << lane)));
        const unsigned emask = (N >= 32 ? FULL : ((1u << N) - 1u)) << (g * N & 31);   // the lanes of the lane's env
        for (int ii = 1; ii < N && !distinct; ++ii) {         // drone 0 keeps its first draw
            for (int attempt = 1; attempt < 48; ++attempt) {
                const int ci = __shfl_sync(FULL, code, (g * N + ii) & 31);
                const float dx = (float)((ci >> 8) - (code >> 8)), dy = (float)((ci & 255) - (code & 255));
                const unsigned conf = __ballot_sync(FULL, act && i < ii && dx * dx + dy * dy <= sep2);
                const bool redo = act && (conf & emask);      // the env's drone ii is too close to an earlier one
                if (!__any_sync(FULL, redo)) break;
                if (redo && i == ii) code = candidate(ii, attempt, s_idx, tq);
            }
        }
        int t_idx = 0;
        if (act) {
            const int n_t = od.pool_off[tq + 1] - od.pool_off[tq];
            t_idx = od.pool_off[tq] + (int)(draw(key, 1u + 256u * (unsigned)i + 3u) % (unsigned)n_t);
        }
        if (PL && __any_sync(FULL, act && od.path_cells == nullptr)) {
            // no table of paths for this map: reset_world's per-episode search (jps_find_path + pruning, ATT:317-331), one drone
            // after the other with the whole warp on each, straight into the drone's reference-line row of the state
            uint8_t *scratch = p.plan_scratch + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * p.plan_stride;
            const int t_code = act ? (int)od.cell_code[t_idx] : 0;
            const int wcap = VAR == AAC_VARIANT_MM ? min(W, 31) : W;
            for (unsigned left = __ballot_sync(FULL, act && od.path_cells == nullptr); left; left &= left - 1u) {
                const int src = __ffs(left) - 1;
                const int sc = __shfl_sync(FULL, code, src), tc = __shfl_sync(FULL, t_code, src);
                const int mrow = __shfl_sync(FULL, map_row, src), gee = __shfl_sync(FULL, ge, src), ii = __shfl_sync(FULL, i, src);
                const MapDev &mq = VAR == AAC_VARIANT_MM ? w.map[mrow] : *w.map;
                const int gyq = mq.gy;
                uint16_t *dst = p.st.ref_cells + ((size_t)gee * N + ii) * W;
                int cnt = plan_path_warp([&](const int c) { return occupied(mq, c / gyq, c - (c / gyq) * gyq); }, mq.gx, gyq, (sc >> 8) * gyq + (sc & 255),
                                         (tc >> 8) * gyq + (tc & 255), scratch, dst, wcap);
                if (cnt < 2) {   // unreachable goal, or more vertices than the row holds: a straight line, and the counter says so
                    if (lane == 0) { dst[0] = (uint16_t)sc; dst[1] = (uint16_t)tc; if (p.stats) atomicAdd(p.stats + 10, 1.0); }
                    cnt = 2;
                }
                if (lane == 0)
                    for (int k = cnt; k < ((cnt + 7) & ~7); ++k) dst[k] = 0;   // whole 16-byte chunks, zero padded as the table's paths are
                __syncwarp();
                if (lane == src) { nw = cnt; row = dst; }
            }
            if (act && od.path_cells == nullptr) head = *reinterpret_cast<const uint4 *>(row);
        }
        if (act && !(PL && od.path_cells == nullptr)) {
            const int pr = s_idx * od.n_cells + t_idx;
            nw = od.path_len[pr];
            const uint16_t *src = od.path_cells + od.path_off[pr];   // 16-byte aligned: paths are padded to 8 cells
            const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
            uint4 *d4 = reinterpret_cast<uint4 *>(p.st.ref_cells + ((size_t)ge * N + i) * W);
            const int nc = (nw + 7) >> 3;
            for (int k0 = 0; k0 < nc; k0 += 4) {   // four 16-byte chunks (a whole 32-vertex line) in flight
                uint4 t[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (k0 + u < nc) t[u] = s4[k0 + u];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (k0 + u < nc) d4[k0 + u] = t[u];
                if (k0 == 0) head = t[0];
            }
            row = src;
        }
        __syncwarp();
    } else {
        // reference lines: 16-byte chunks bank -> global state (W is a multiple of 8), env by env with the whole warp
        for (unsigned left = mask; left; left &= left - 1u) {
            const int gg = __ffs(left) - 1;
            const unsigned sc = __shfl_sync(FULL, scen, gg * N);
            const uint4 *src = reinterpret_cast<const uint4 *>(p.bank_cells + (size_t)sc * N * W);
            uint4 *dst = reinterpret_cast<uint4 *>(p.st.ref_cells + (size_t)(w.e_lo + gg) * N * W);
            for (int c = lane; c < N * W / 8; c += 32) dst[c] = src[c];
        }
        if (act) {
            row = p.bank_cells + ((size_t)scen * N + i) * W;
            nw = p.bank_w[(size_t)scen * N + i];
            head = *reinterpret_cast<const uint4 *>(row);   // bank rows start on 16-byte boundaries
        }
    }
    if (act) {
        const int a = lane;
        w.refw[a] = (uint8_t)nw;
        p.st.ref_w[(size_t)ge * N + i] = (uint8_t)nw;
        reinterpret_cast<uint4 *>(w.c8)[a] = head;
        const unsigned c0 = head.x & 0xFFFFu, c1 = head.x >> 16;
        const float px = cell_cx(mp, c0 >> 8), py = cell_cy(mp, c0 & 255);
        w.px[a] = px; w.py[a] = py; w.vx[a] = 0.0f; w.vy[a] = 0.0f;
        w.hd[a] = atan2f(cell_cy(mp, c1 & 255) - py, cell_cx(mp, c1 >> 8) - px);  // ATT:359
        w.ppx[a] = px; w.ppy[a] = py; w.pvx[a] = 0.0f; w.pvy[a] = 0.0f;
        w.meta[a] = 0xFFFF0000u;
        if (VAR == AAC_VARIANT_MM) { w.amap[a] = (uint8_t)map_row; w.wpm[a] = (1u << nw) - 2u; }   // every vertex after the start (MM:345)
        if (p.st.wall_count) p.st.wall_count[(size_t)ge * N + i] = 0;
    }
    __syncwarp();
    return row;
}

// vertex k of a drone's reference line: the first 8 sit in shared memory (w.c8), longer lines continue in the global row
struct CellRow {
    const uint16_t *s, *g;
    __device__ __forceinline__ unsigned operator[](int k) const { return k < 8 ? s[k] : g[k]; }
};

// distance to the reference polyline and arc length of the nearest point: first segment attaining the
// minimum wins (ATT:3203-3214, V2:4286-4297, UV2:413-441).  Vertices come 8 per 16-byte load.
__device__ __forceinline__ void polyline_nearest(const MapDev &mp, const CellRow cells, int nw, float px, float py, float &best2, float &arc,
                                                 float &total) {
    best2 = CUDART_INF_F; arc = 0.0f;
    float run = 0.0f;
    unsigned c = cells[0];
    float ax = cell_cx(mp, c >> 8), ay = cell_cy(mp, c & 255);
#pragma unroll 1
    for (int k = 1; k < nw; ++k) {
        c = cells[k];   // (beyond the 8 staged vertices: one 2-byte load per vertex; a 16-byte load per 8 with the chunk shifted
                        //  along in registers was measured 0.9 % slower)
        const float bx = cell_cx(mp, c >> 8), by = cell_cy(mp, c & 255);
        const float sx = bx - ax, sy = by - ay, len2 = sx * sx + sy * sy;
        const float sl = sqrtf(len2);
        float rr = len2 > 0.0f ? __fdividef((px - ax) * sx + (py - ay) * sy, len2) : 0.0f;
        rr = fminf(fmaxf(rr, 0.0f), 1.0f);
        const float qx = fmaf(rr, sx, ax) - px, qy = fmaf(rr, sy, ay) - py;
        const float dq = qx * qx + qy * qy;
        if (dq < best2) { best2 = dq; arc = fmaf(rr, sl, run); }
        run += sl;
        ax = bx; ay = by;
    }
    total = run;
}

// EVS, per-env lane, drone i of the env in loop order: the collision part of ss_reward_Mar's neighbour scan with the
// crash flags as they stand at that point of the reference's loop (V2:3143-3158) - previous steps' flags, this step's
// boundary / building flags of the drones already visited, and the drone_collision flags raised so far in this pass
// (`dcol`).  A collision overrides the goal / normal outcome computed in the per-drone phase.
__device__ __noinline__ void evs_collisions(const KParams &p, const Warp &w, int eb, int i, int N, int M, int Mp, unsigned &dcol) {
    const int a = eb + i;
    const float px = w.px[a], py = w.py[a], coll2 = 4.0f * p.prot * p.prot;
    const unsigned old = w.meta[a], crash = M_VBOUND | M_VBLDG | M_VDRONE;
    const int pn0 = (old >> 16) & 0xFF, pn1 = (old >> 24) & 0xFF;
    int n_coll = 0;
    bool prev2 = false;
    float brg = -1.0f;
    for (int k = 0; k < M; ++k) {
        const int j = w.order[a * M + k], b = eb + j;
        if (w.atgoal[b]) continue;                  // V2:3128-3130
        if (w.d2[a * Mp + k] > coll2) break;        // neighbours come in ascending distance
        brg = bearing_deg(px, py, w.px[b], w.py[b]);  // taken before the terminal-state test (V2:3146)
        const unsigned fj = (j < i ? w.meta2[b] : w.meta[b]) & crash;
        if (fj || ((dcol >> j) & 1u)) continue;     // V2:3149-3152
        ++n_coll;
        dcol |= (1u << i) | (1u << j);              // V2:3157-3158
        w.meta2[b] |= M_VDRONE;
        prev2 |= (j == pn0) || (M > 1 && j == pn1);
    }
    if ((dcol >> i) & 1u) w.meta2[a] |= M_VDRONE;
    const unsigned f = w.agf[a], br = (f >> F_BRANCH_SHIFT) & 7u;
    if (n_coll > 0 && br >= 3) {  // boundary and building come first in the branch chain (V2:3546-3567)
        w.agf[a] = (2u << F_BRANCH_SHIFT) | ((brg >= 90.0f && brg <= 180.0f) ? F_DOUBLE : 0u) | (prev2 ? F_BBC3 : 0u);
        w.meta2[a] = (w.meta2[a] & ~0xFFu) | (old & 0xFFu);  // no waypoint pop on the crash branch
        if (p.out.branch) p.out.branch[w.a0 + a] = 2;
    }
}

// EVS: forV2's evaluation "by sorties" (args.mode == 'eval' and evaluation_by_episode == False): terminal drones stay put
// (V2:3729-3734), neighbours at their goal are invisible and crash flags are live across the drone loop (V2:3128-3158),
// crashes do not end the episode (V2:3551-3587).  A separate instantiation: the training kernels carry none of it.
// MT: the mode as a compile-time constant for the launches of the benchmark shapes (-1 = the `mode` argument decides;
// MT_STEP_ONLY = MODE_STEP without the fused auto-reset; MODE_RESET): the step launch then carries no reset code and the
// reset launch no reward code - less code per launch is what the instruction cache rewards.
constexpr int MT_STEP_ONLY = 3;
constexpr int MT_PHASED = 4;   // one launch: MT_STEP_ONLY's loop over all groups, then MODE_RESET's over the groups with terminated envs
#ifndef AAC_MIN_BLOCKS
#define AAC_MIN_BLOCKS 4   // resident CTAs of 256 threads per SM the register allocation aims at (64 registers)
#endif
// The persistent warp's loop over groups for one mode (MT as in env_kernel).  PHASE: 0 = the launch runs this loop alone;
// 1 / 2 = first / second loop of a phased launch (MT_PHASED: every group is stepped, then - by whichever warp gets to it -
// the envs it terminated are re-initialised): loop 1 publishes a group's completion in p.flags (release), loop 2 takes
// the groups from a second counter and waits for the group's flag (see ld_flag).  A flag is (epoch << 32) | the group's
// terminated-env bits: loop 2 learns from the one word both that the group is through and which envs to re-initialise.
template <int VAR, bool AUX, bool LEAN, int NT, int RT, bool EVS, int MT, int RM, bool CS, int PHASE, bool PL>
__device__ __forceinline__ void group_loop(const KParams &p, Warp &w, const int mode_arg, const MapDev *s_map, int &st_i, float &st_f) {
    constexpr bool STEP_ONLY = MT == MT_STEP_ONLY;
    const int mode = MT < 0 ? mode_arg : (STEP_ONLY ? (int)MODE_STEP : MT);
    const int lane = w.lane;
    const int N = NT ? NT : p.N, M = N - 1, W = p.W, G = p.G;
    const int Mp = M | 1;
    const int flags = LEAN ? 0 : p.out_flags;   // LEAN: no optional output was requested, their code is compiled out
    int *const counter = p.work + 2 * p.parity + (PHASE == 2 ? 1 : 0);

    // ---- persistent warp: fetch a group of G whole envs, run the pipeline, fetch the next
    const int n_groups = (p.E + G - 1) / G;
    int done_gi = -1;   // PHASE 1: the group whose completion flag is still to be published, and its terminated envs
    unsigned done_mask = 0;
    auto publish = [&]() { st_release(p.flags + done_gi, ((unsigned long long)(unsigned)p.epoch << 32) | done_mask); };
    for (;;) {
        // (claiming the next group ahead of time hides the atomic's round trip but was measured 4 % SLOWER: with four to five
        //  groups per warp a group claimed early by a busy warp is a group an idle warp cannot take at the tail.  The same in
        //  the reset loop of a phased launch: claim ahead 1.4 % slower, claim ahead + prefetched flag 5 % slower)
        int gi = 0;
        if (lane == 0) gi = atomicAdd(counter, 1);
        gi = __shfl_sync(FULL, gi, 0);
        if (gi >= n_groups) break;
        w.e_lo = gi * G;
        w.ng = max(0, min(G, p.E - w.e_lo));
        unsigned group_mask = 0;   // PHASE 2: the envs of the group its step terminated
        if (PHASE == 2) {   // wait until the group's step has left its results (whichever warp ran it)
            if (lane == 0) {
                unsigned long long f;
                while ((unsigned)((f = ld_flag(p.flags + gi)) >> 32) != (unsigned)p.epoch) __nanosleep(64);
                group_mask = (unsigned)f;
            }
            group_mask = __shfl_sync(FULL, group_mask, 0);
            if (!group_mask) continue;
        }
        // reset launch: a group none of whose envs is masked has nothing to do (with a trained policy most groups)
        if (PHASE != 2 && mode == MODE_RESET && p.mask && !__ballot_sync(FULL, lane < w.ng && p.mask[w.e_lo + lane])) continue;
        w.a0 = w.e_lo * N;
        w.nA = w.ng * N;
        const int nA = w.nA, a0 = w.a0;
        const bool mine = lane < nA;
        const int a = lane, ga = a0 + lane;
        const int my_env = mine ? a / N : 0;
        const uint16_t *cells = p.st.ref_cells + (size_t)(mine ? ga : a0) * W;

        // ---- load the per-drone records; integrate the action (ATT:2655-2713).  A reset launch re-initialises every env it
        //      touches (init_envs fills all of a drone's slots) and stores only those: it never reads the old records
        if (mine && mode != MODE_RESET) {
            float px = p.st.px[ga], py = p.st.py[ga], vx = p.st.vx[ga], vy = p.st.vy[ga], hd = p.st.heading[ga];
            const unsigned meta = p.st.meta[ga];
            w.ppx[a] = px; w.ppy[a] = py; w.pvx[a] = vx; w.pvy[a] = vy;
            const bool frozen = EVS && (meta & (M_REACH | M_VBOUND | M_VBLDG | M_VDRONE));  // V2:3729-3734
            if (mode == MODE_STEP && !frozen) {
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
            w.px[a] = px; w.py[a] = py; w.vx[a] = vx; w.vy[a] = vy; w.hd[a] = hd;
            w.meta[a] = meta;
            w.refw[a] = p.st.ref_w[ga];
            reinterpret_cast<uint4 *>(w.c8)[a] = *reinterpret_cast<const uint4 *>(cells);   // rows are 16-byte aligned (w_max % 8 == 0)
            if (VAR == AAC_VARIANT_MM) { w.amap[a] = (uint8_t)p.st.map_id[w.e_lo + my_env]; w.wpm[a] = p.st.wp_mask[ga]; }
        }
        __syncwarp();
        // PHASE 1: the previous group's flag leaves here - its stores (ordered before this point for every lane by the
        // __syncwarp()s since) have drained while this group's records were on their way, so the release waits for nothing
        if (PHASE == 1 && done_gi >= 0 && lane == 0) publish();

        unsigned reset_mask = 0, store_mask = (w.ng >= 32) ? FULL : ((1u << w.ng) - 1u), term_mask = 0;
        if (mode == MODE_RESET) {
            bool m = lane < w.ng && (!p.mask || p.mask[w.e_lo + lane]);
            reset_mask = PHASE == 2 ? group_mask : __ballot_sync(FULL, m);
            store_mask = reset_mask;
        }
        // Jobs of a group: job 0 = the whole group (step / observe), job 1 = the envs to re-initialise, all at once (reset
        // mode: the masked envs; step mode: the envs that just terminated).  One copy of the pipeline serves both.
        for (int job = (mode == MODE_RESET ? 1 : 0); job <= (STEP_ONLY ? 0 : 1); ++job) {
            const uint16_t *cl = cells;
            if (job > 0) {
                if (!reset_mask) break;
                const uint16_t *row = init_envs<VAR, NT, PL>(p, w, reset_mask);
                cl = row ? row : p.st.ref_cells;
            }
            observe_range<VAR, AUX, LEAN, NT, RT, RM, CS>(p, w, 0, nA, cl, job > 0 && p.rtab != nullptr, job == 0 && mode == MODE_STEP,
                                                          STEP_ONLY ? false : (MT == MODE_RESET ? true : job > 0), job > 0 ? reset_mask : FULL);
            if (job > 0 || mode != MODE_STEP) continue;

            // ---- reward / collision / goal per drone
            if (mine) {
                const MapDev &mp = VAR == AAC_VARIANT_MM ? w.map[w.amap[a]] : *s_map;
                const int eb = my_env * N, i = a - eb;
                const float px = w.px[a], py = w.py[a];
                const int nw = w.refw[a];
                unsigned meta = w.meta[a];
                const bool at_goal = w.atgoal[a];
                if (VAR == AAC_VARIANT_V2 && at_goal) meta |= M_REACH;  // top of the drone's iteration (V2:3025-3033)
                const bool reach_i = meta & M_REACH;
                // neighbour scan (ATT:2187-2236, V2:3092-3168)
                int nearest = -1, n_coll = 0, last_coll = -1, imm_key = -1, conf_cur = 0, conf_pre = 0;
                bool prev2 = false;
                float shortest2 = CUDART_INF_F, imm_tcpa = CUDART_INF_F, imm_d = CUDART_INF_F;
                const int pn0 = (meta >> 16) & 0xFF, pn1 = (meta >> 24) & 0xFF;
                const float coll2 = 4.0f * p.prot * p.prot;
                const float hvx = w.vx[a], hvy = w.vy[a], hppx = w.ppx[a], hppy = w.ppy[a], hpvx = w.pvx[a], hpvy = w.pvy[a];
#pragma unroll 1
                for (int k = 0; k < M && VAR != AAC_VARIANT_MM; ++k) {
                    const int j = w.order[a * M + k], b = eb + j;
                    // tdCPA against every neighbour for the current and the previous state (ATT:2189-2196)
                    float t1, d1, t2, dd2;
                    bool c1, c2;
                    tcpa_dcpa(px, py, hvx, hvy, w.px[b], w.py[b], w.vx[b], w.vy[b], 2.0f * p.prot, t1, d1, c1);
                    tcpa_dcpa(hppx, hppy, hpvx, hpvy, w.ppx[b], w.ppy[b], w.pvx[b], w.pvy[b], 2.0f * p.prot, t2, dd2, c2);
                    conf_cur += c1; conf_pre += c2;
                    if (t1 >= 0.0f && t1 < imm_tcpa) { imm_tcpa = t1; imm_d = d1; imm_key = j; }
                    else if (t1 == -10.0f && d1 < imm_tcpa) { imm_tcpa = t1; imm_d = d1; imm_key = j; }
                    if (EVS && w.atgoal[b]) continue;  // a neighbour touching its goal is invisible (V2:3114-3130)
                    const float d2 = w.d2[a * Mp + k];
                    if (d2 < shortest2) { shortest2 = d2; nearest = j; }
                    if (!EVS && d2 <= coll2) {  // EVS: collisions depend on flags set along the drone loop: per-env pass below
                        if (VAR == AAC_VARIANT_V2) {
                            // reach_target of drone j as drone i sees it: set in earlier steps, or earlier in
                            // this step's loop when j < i (V2:3160)
                            const bool reach_j = (w.meta[eb + j] & M_REACH) || (j < i && w.atgoal[eb + j]);
                            if (reach_j || reach_i) continue;
                            prev2 |= (j == pn0) || (M > 1 && j == pn1);  // V2:3170-3179
                        }
                        ++n_coll; last_coll = j;
                    }
                }
                float shortest = sqrtf(shortest2);
                reinterpret_cast<float4 *>(p.out.tcpa_min)[ga] = make_float4(imm_tcpa, imm_d, (float)imm_key, (float)(conf_cur + 256 * conf_pre));
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
                const CellRow crow{w.c8 + a * 8, cells};
                const unsigned cw = crow[1 + cur], cg = crow[nw - 1];
                const float gx = cell_cx(mp, cg >> 8), gy = cell_cy(mp, cg & 255);
                const float wdx = px - cell_cx(mp, cw >> 8), wdy = py - cell_cy(mp, cw & 255);
                const bool wp_flag = wdx * wdx + wdy * wdy < 25.0f;
                const float ppx = w.ppx[a], ppy = w.ppy[a];
                const bool hit_bound = capsule_hits_bound(ppx, ppy, px, py, p.prot, mp.hx, mp.hy);
                float after_hg = sqrtf((px - gx) * (px - gx) + (py - gy) * (py - gy));
                const unsigned mr_bits = w.minr[a];
                const float min_radar = mr_bits > 0x7F800000u ? CUDART_INF_F : __uint_as_float(mr_bits);
                float dist_to_goal, near_drone = 0.0f, near_bldg = 0.0f, small_step = 0.0f, rew = 0.0f, cross_err = 0.0f;
                unsigned res = 0, branch;
                if (VAR == AAC_VARIANT_MM) {
                    // ss_reward of the multipleMap variant (MM:1674-2007): no drone-collision branch
                    unsigned mask = w.wpm[a];
                    const unsigned cl = crow[31 - __clz(mask | 1u)];   // goal[-1] before this step's pop (MM:1735)
                    const float glx = cell_cx(mp, cl >> 8), gly = cell_cy(mp, cl & 255);
                    after_hg = sqrtf((px - glx) * (px - glx) + (py - gly) * (py - gly));
                    // waypoint scan (MM:1742-1762): walk the remaining waypoints in order; every new running minimum
                    // becomes next_wp; the first running minimum closer than 5 m is popped (unless it is the only
                    // one left) and next_wp becomes the nearest remaining waypoint
                    int n_goal = __popc(mask);
                    float smallest2 = CUDART_INF_F, nwx = 0.0f, nwy = 0.0f;
                    bool wpf = false;
                    unsigned scan = mask;
#pragma unroll 1
                    while (scan) {
                        const int k = __ffs(scan) - 1;
                        scan &= scan - 1;
                        const unsigned c = crow[k];
                        const float wx = cell_cx(mp, c >> 8), wy = cell_cy(mp, c & 255);
                        const float d2 = (px - wx) * (px - wx) + (py - wy) * (py - wy);
                        if (d2 < smallest2) {
                            smallest2 = d2; nwx = wx; nwy = wy;
                            if (d2 < 25.0f) {
                                wpf = true;
                                if (n_goal > 1) {
                                    mask &= ~(1u << k); --n_goal;
                                    float best = CUDART_INF_F;
                                    unsigned s2 = mask;
#pragma unroll 1
                                    while (s2) {
                                        const int q = __ffs(s2) - 1;
                                        s2 &= s2 - 1;
                                        const unsigned cq = crow[q];
                                        const float qx = cell_cx(mp, cq >> 8), qy = cell_cy(mp, cq & 255);
                                        const float dq = (px - qx) * (px - qx) + (py - qy) * (py - qy);
                                        if (dq < best) { best = dq; nwx = qx; nwy = qy; }
                                    }
                                }
                                break;
                            }
                        }
                    }
                    w.wpm[a] = mask;
                    shortest = sqrtf(smallest2);
                    // |pre - wp| - |pos - wp| without the cancellation (MM:1779-1795)
                    const float bx = ppx - nwx, by = ppy - nwy, cx = px - nwx, cy = py - nwy;
                    const float den = sqrtf(bx * bx + by * by) + sqrtf(cx * cx + cy * cy);
                    dist_to_goal = den > 0.0f ? ((bx - cx) * (bx + cx) + (by - cy) * (by + cy)) / den : 0.0f;
                    float best2, arc, total;
                    polyline_nearest(mp, crow, nw, px, py, best2, arc, total);
                    cross_err = sqrtf(best2);
                    const float dist_to_ref = cross_err <= p.prot ? 3.0f * (1.0f - cross_err / p.prot) : -3.0f;   // MM:1812-1817
                    near_drone = dist_to_ref;   // reported in the `near_drone` slot of the parts record
                    if (min_radar >= p.prot && min_radar <= 10.0f) near_bldg = 3.0f * fmaf((0.0f - 1.0f) / (10.0f - p.prot), min_radar, 10.0f / 7.5f);  // MM:1845-1860
                    if (hit_bound) { rew = dist_to_ref - 5.0f + dist_to_goal - near_bldg; res |= F_DONE; branch = 0; }          // MM:1962-1969
                    else if (collide_building) { rew = dist_to_ref - 5.0f + dist_to_goal - near_bldg; res |= F_DONE; branch = 1; }
                    else if (at_goal) { res |= F_GOAL; meta |= M_REACH; rew = 5.0f; branch = 3; }
                    else { rew = ((wpf && n_goal > 1) ? 3.0f : 0.0f) + dist_to_ref + dist_to_goal - near_bldg; branch = 4; }    // MM:1987-1993
                } else if (VAR == AAC_VARIANT_ATT) {
                    // (|pre-g| - |pos-g|) / vmax without the cancellation: (a-b).(a+b) / (|a|+|b|)  (ATT:2319-2325)
                    const float bx = ppx - gx, by = ppy - gy, cx = px - gx, cy = py - gy;
                    const float den = sqrtf(bx * bx + by * by) + after_hg;
                    dist_to_goal = den > 0.0f ? ((bx - cx) * (bx + cx) + (by - cy) * (by + cy)) / den / p.vmax : 0.0f;
                    const float c_dr = 1.0f + (2.5f / (10.0f - 2.5f)), m_dr = (0.0f - 1.0f) / (10.0f - 2.5f);  // ATT:2420-2426
                    for (int k = 0; k < M; ++k) {  // every in-band neighbour adds the NEAREST one's penalty (ATT:2430-2432)
                        const float d2 = w.d2[a * Mp + k];
                        if (d2 >= 6.25f && d2 <= 100.0f) near_drone += fmaf(m_dr, shortest, c_dr);
                    }
                    if (hit_bound) { rew = -20.0f - small_step - near_bldg; res |= F_DONE; branch = 0; }
                    else if (n_coll > 0) {
                        rew = -20.0f - small_step - near_drone; res |= F_DONE; branch = 2;
                        if (last_coll == nearest) res |= F_BBC3;
                    } else if (at_goal) { res |= F_GOAL; meta |= M_REACH; rew = 20.0f; branch = 3; }
                    else {
                        if (wp_flag && (nw - 1 - cur) > 1) meta = (meta & ~0xFFu) | (unsigned)(cur + 1);  // ATT:2565-2566
                        rew = __fsub_rn(__fsub_rn(__fsub_rn(dist_to_goal, small_step), near_bldg), near_drone);  // ATT:2576-2578
                        branch = 4;
                    }
                    if (flags & AAC_OUT_PARTS) {  // cross-track error is reported, not rewarded (ATT:2368 coefficient 0)
                        float best2, arc, total;
                        polyline_nearest(mp, crow, nw, px, py, best2, arc, total);
                        cross_err = sqrtf(best2);
                    }
                } else {
                    float best2, arc, total;
                    polyline_nearest(mp, crow, nw, px, py, best2, arc, total);
                    cross_err = sqrtf(best2);
                    const float dist_left = cross_err + (total - arc);  // UV2:413-441
                    dist_to_goal = 6.0f * (1.0f - dist_left / total);    // V2:3257-3268
                    if (nearest >= 0 && shortest >= 2.5f && shortest <= 6.0f) {  // V2:3365-3386
                        const float c_dr = 1.0f + (2.5f / (6.0f - 2.5f)), m_dr = (0.0f - 1.0f) / (6.0f - 2.5f);
                        const float brg = bearing_deg(px, py, w.px[eb + nearest], w.py[eb + nearest]);
                        const float coef = (brg >= 90.0f && brg <= 180.0f) ? 20.0f : 10.0f;
                        near_drone = coef * fmaf(m_dr, shortest, c_dr);
                    }
                    const float spd = sqrtf(w.vx[a] * w.vx[a] + w.vy[a] * w.vy[a]);
                    const float thr = 0.5f * p.vmax;  // V2:3446-3453
                    small_step = 5.0f * ((thr - fminf(fmaxf(spd, 0.0f), thr)) * (1.0f / thr));
                    if (min_radar >= p.prot && min_radar <= 5.0f) near_bldg = 3.0f * fmaf((0.0f - 1.0f) / (5.0f - p.prot), min_radar, 2.0f);  // V2:3522-3539
                    // crash rewards are filled in by the per-env lanes below: the penalty doubles along the
                    // drone loop (V2:3590-3594)
                    if (hit_bound) { meta |= M_VBOUND; res |= EVS ? 0u : F_DONE; branch = 0; }
                    else if (collide_building) { res |= EVS ? 0u : F_DONE; branch = 1; }
                    else if (n_coll > 0) {
                        res |= F_DONE; branch = 2;
                        const float brg = bearing_deg(px, py, w.px[eb + last_coll], w.py[eb + last_coll]);
                        if (brg >= 90.0f && brg <= 180.0f) res |= F_DOUBLE;
                        if (prev2) res |= F_BBC3;
                    } else if (at_goal) { res |= F_GOAL; rew = 20.0f; branch = 3; }
                    else {
                        if (wp_flag && (nw - 1 - cur) > 1) meta = (meta & ~0xFFu) | (unsigned)(cur + 1);
                        // explicit roundings: the terms are also an output (parts), and a product contracted into
                        // this sum in one instantiation but not another would make them differ in the last bit
                        rew = __fsub_rn(__fsub_rn(__fsub_rn(dist_to_goal, small_step), near_bldg), near_drone);  // V2:3631-3633
                        branch = 4;
                    }
                }
                if (collide_building && p.st.wall_count) atomicAdd(p.st.wall_count + ga, 1);   // result unused: a reduction, nothing to wait for
                w.agf[a] = res | (branch << F_BRANCH_SHIFT);
                w.agr[a] = rew;
                if (flags & AAC_OUT_PARTS) {
                    float4 *q = reinterpret_cast<float4 *>(p.out.parts + (size_t)ga * 8);
                    q[0] = make_float4(dist_to_goal, near_drone, near_bldg, small_step);
                    q[1] = make_float4(cross_err, after_hg, min_radar, shortest);
                    p.out.branch[ga] = (int8_t)branch;
                }
                // other drones read w.meta[j] & M_REACH (the record as it stood before this step) in the loop
                // above: publish the new record only after every drone of the group is through
                w.meta2[a] = meta;
            }
            __syncwarp();

            // ---- per env: crash penalties, summed reward, bound_building_check, episode end
            bool reset_me = false;
            unsigned bbc = 0, any_done = 0, all_reach = 1, n_reach = 0, any_goal = 0, term = 0;
            int step = 0;
            float ret = 0.0f;
            if (lane < w.ng) { step = p.st.ep_step[w.e_lo + lane]; ret = p.st.ep_return[w.e_lo + lane]; }   // in flight across the ballots below
            if (!EVS) {
                // every drone's outcome flags meet in a handful of ballots; an env lane reads its env's bits out of them
                const unsigned f = mine ? w.agf[a] : (4u << F_BRANCH_SHIFT), br = (f >> F_BRANCH_SHIFT) & 7u;
                const unsigned b_dbl = __ballot_sync(FULL, br <= 2 && (f & F_DOUBLE));
                if (VAR == AAC_VARIANT_V2 && br <= 2) {
                    // the crash penalty doubles along the drone loop with every doubled crash up to and including this one
                    // (V2:3590-3594): 20 * 2^count, exact
                    const unsigned upto = b_dbl & (((N >= 32 ? FULL : ((1u << N) - 1u)) << (my_env * N)) & (0xFFFFFFFFu >> (31 - lane)));
                    w.agr[a] = -__int_as_float(__float_as_int(20.0f) + (__popc(upto) << 23));
                }
                const unsigned b0 = __ballot_sync(FULL, br == 0), b1 = __ballot_sync(FULL, br == 1), b2 = __ballot_sync(FULL, br == 2);
                const unsigned b3 = __ballot_sync(FULL, (f & F_BBC3) != 0), bd = __ballot_sync(FULL, (f & F_DONE) != 0);
                const unsigned bg = __ballot_sync(FULL, (f & F_GOAL) != 0), brc = __ballot_sync(FULL, mine && (w.meta2[a] & M_REACH));
                __syncwarp();
                if (lane < w.ng) {
                    const unsigned m = (N >= 32 ? FULL : ((1u << N) - 1u)) << (lane * N);
                    bbc = ((b0 & m) ? 1u : 0u) | ((b1 & m) ? 2u : 0u) | ((b2 & m) ? 4u : 0u) | ((b3 & m) ? 8u : 0u);
                    any_done = bd & m; any_goal = bg & m;
                    all_reach = (brc & m) == m; n_reach = __popc(brc & m);
                }
            }
            if (lane < w.ng) {
                const int ge = w.e_lo + lane, eb = lane * N;
                float sum = 0.0f;
                if (EVS) {   // crash flags are live along the drone loop (V2:3128-3158): serial
                    float cp = 20.0f;
                    unsigned evs_dcol = 0;
#pragma unroll 1
                    for (int i = 0; i < N; ++i) {
                        evs_collisions(p, w, eb, i, N, M, Mp, evs_dcol);
                        const unsigned f = w.agf[eb + i], br = (f >> F_BRANCH_SHIFT) & 7u;
                        if (br <= 2) {
                            if (f & F_DOUBLE) cp *= 2.0f;
                            w.agr[eb + i] = -cp;
                            bbc |= 1u << br;
                        }
                        if (f & F_BBC3) bbc |= 8u;
                        any_done |= f & F_DONE;
                        any_goal |= f & F_GOAL;
                        const unsigned reached = (w.meta2[eb + i] & M_REACH) ? 1u : 0u;
                        all_reach &= reached;
                        n_reach += reached;
                    }
                }
#pragma unroll 1
                for (int i = 0; i < N; ++i) sum += w.agr[eb + i];
                if (p.sum_reward) {  // reward = [sum(reward)] * N (ATT:2602-2603)
#pragma unroll 1
                    for (int i = 0; i < N; ++i) w.agr[eb + i] = sum;
                    sum *= (float)N;
                }
                step += 1;
                ret += sum;
                term = (step > p.ep_len ? 1u : 0u) | (any_done ? 2u : 0u) | (all_reach ? 4u : 0u);
                reinterpret_cast<uchar4 *>(p.out.bbc)[ge] = make_uchar4(bbc & 1, (bbc >> 1) & 1, (bbc >> 2) & 1, (bbc >> 3) & 1);
                p.out.terminated[ge] = (uint8_t)term;
                w.rs[lane] = any_goal ? 1 : 0;
                reset_me = !STEP_ONLY && term && p.autoreset;
                if (!reset_me) { p.st.ep_step[ge] = step; p.st.ep_return[ge] = ret; }
            }
            // episode counters (ATT/ma_main:581-637) of the envs that finished, summed over the group
            const unsigned tb = __ballot_sync(FULL, term != 0);
            term_mask = tb;
            if (tb && p.stats) {
                const int steps_sum = __reduce_add_sync(FULL, term ? step : 0), reach_sum = __reduce_add_sync(FULL, term ? (int)n_reach : 0);
                float r = term ? ret : 0.0f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(FULL, r, o);
                const int c0 = __popc(__ballot_sync(FULL, term && (bbc & 1u))), c1 = __popc(__ballot_sync(FULL, term && (bbc & 2u)));
                const int c2 = __popc(__ballot_sync(FULL, term && (bbc & 4u))), c3 = __popc(__ballot_sync(FULL, term && (bbc & 8u)));
                const int c4 = __popc(__ballot_sync(FULL, term && all_reach)), c6 = __popc(__ballot_sync(FULL, term == 1u));
                st_i += lane == 0 ? __popc(tb) : lane == 1 ? steps_sum : lane == 3 ? c0 : lane == 4 ? c1 : lane == 5 ? c2 : lane == 6 ? c3 : lane == 7 ? c4
                        : lane == 8 ? reach_sum : lane == 9 ? c6 : 0;
                st_f += r;
            }
            reset_mask = STEP_ONLY ? 0u : __ballot_sync(FULL, reset_me);
            __syncwarp();
            // the terminal transition leaves before the reset touches the records
            if (mine) {
                const unsigned f = w.agf[a];
                w.meta[a] = w.meta2[a];
                p.out.reward[ga] = w.agr[a];
                p.out.done[ga] = (uint8_t)(f & F_DONE ? 1 : 0);
                // multipleMap raises check_goal[reward_record_idx] with an index that is never incremented
                // (MM:1680, :1958): slot 0 of the env stands for "some drone touched its goal"
                if (VAR == AAC_VARIANT_MM) p.out.check_goal[ga] = (uint8_t)(a == my_env * N ? w.rs[my_env] : 0);
                else p.out.check_goal[ga] = (uint8_t)(f & F_GOAL ? 1 : 0);
            }
            __syncwarp();
        }

        // ---- write the per-drone records back; the first two neighbour keys become the next step's
        //      pre_surroundingNeighbor (V2:3170-3179)
        if (mine && ((store_mask >> my_env) & 1u)) {
            unsigned meta = w.meta[a];
            if (M > 0 && VAR != AAC_VARIANT_MM) meta = (meta & 0x0000FFFFu) | ((unsigned)w.order[a * M] << 16) | ((unsigned)(M > 1 ? w.order[a * M + 1] : 0xFF) << 24);
            p.st.px[ga] = w.px[a]; p.st.py[ga] = w.py[a]; p.st.vx[ga] = w.vx[a]; p.st.vy[ga] = w.vy[a]; p.st.heading[ga] = w.hd[a];
            p.st.meta[ga] = meta;
            if (VAR == AAC_VARIANT_MM) p.st.wp_mask[ga] = w.wpm[a];
        }
        __syncwarp();
        done_gi = gi; done_mask = term_mask;
    }
    if (PHASE == 1 && done_gi >= 0 && lane == 0) publish();
}


template <int VAR, bool AUX, bool LEAN, int NT, int RT, bool EVS = false, int MT = -1, int RM = -1, bool CS = false, bool PL = false>
__global__ void __launch_bounds__(MAX_THREADS, NT >= 16 ? 3 : AAC_MIN_BLOCKS) env_kernel(   // 20-drone envs: 85 registers, no spills, measured 5 % faster on C5
const __grid_constant__ KParams p, const int mode_arg) {
    const int mode = MT < 0 ? mode_arg : ((MT == MT_STEP_ONLY || MT == MT_PHASED) ? (int)MODE_STEP : MT);
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const CtaLayout &CL = p.CL;
    const WarpLayout &WL = p.WL;

    MapDev *s_map = reinterpret_cast<MapDev *>(smem + CL.map);
    float4 *s_ray = reinterpret_cast<float4 *>(smem + CL.ray);
    unsigned long long *s_bar = reinterpret_cast<unsigned long long *>(smem + CL.bar);

    // ---- CTA prologue: the map arrives by one TMA bulk copy, the ray table by plain loads.  This is the
    //      only CTA-wide synchronisation of the kernel.
    if (tid == 0) {
        mbar_init(s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // the map - for multipleMap the whole table of maps when it fits the budget (14 maps = 4.5 KB) - by ONE bulk copy
        mbar_expect_tx(s_bar, p.n_staged * (unsigned)sizeof(MapDev));
        bulk_g2s(s_map, p.maps, p.n_staged * (unsigned)sizeof(MapDev), s_bar);
        if (blockIdx.x == 0) { p.work[2 * (p.parity ^ 1)] = 0; p.work[2 * (p.parity ^ 1) + 1] = 0; }   // the next launch's group counters
    }
    for (int k = tid; k < p.R; k += blockDim.x) s_ray[k] = p.ray_tab[k];
    DdaRay *s_dda = reinterpret_cast<DdaRay *>(smem + CL.dda);
    for (int k = tid; k < p.R; k += blockDim.x) s_dda[k] = p.dda_tab[k];
    uint4 *s_walk = reinterpret_cast<uint4 *>(smem + CL.walk);
    for (int k = tid; k < (int)(WALK_BYTES / 16); k += blockDim.x) s_walk[k] = p.walk_tab[k];
    __syncthreads();
    mbar_wait(s_bar, 0);

    unsigned char *ws = smem + CL.warps + (size_t)warp * WL.total;
    Warp w;
    w.map = (VAR == AAC_VARIANT_MM && p.n_staged < p.n_maps) ? p.maps : s_map; w.ray = s_ray; w.dda = s_dda; w.walk.s = smem_u32(s_walk); w.walk.p = nullptr; w.lane = lane;
    w.px = reinterpret_cast<float *>(ws + WS_CUR); w.py = w.px + 32; w.vx = w.py + 32; w.vy = w.vx + 32; w.hd = w.vy + 32;
    w.ppx = reinterpret_cast<float *>(ws + WS_PRE); w.ppy = w.ppx + 32; w.pvx = w.ppy + 32; w.pvy = w.pvx + 32;
    w.meta = reinterpret_cast<unsigned *>(ws + WS_META); w.meta2 = w.meta + 32; w.minr = w.meta2 + 32; w.agf = w.minr + 32; w.wpm = w.agf + 32;
    w.agr = reinterpret_cast<float *>(ws + WS_AGR);
    w.atgoal = ws + WS_BYTES; w.refw = w.atgoal + 32; w.rs = w.refw + 32; w.amap = w.rs + 32;
    w.win = reinterpret_cast<uint2 *>(ws + WS_WIN);
    w.wrel = reinterpret_cast<float4 *>(ws + WS_WREL);
    w.stg = reinterpret_cast<float *>(ws + WS_STG);
    w.c8 = reinterpret_cast<uint16_t *>(ws + WS_C8);
    // the variable part: compile-time offsets too when the drone count is a template parameter
    const WarpLayout VL = NT ? make_warp_layout(VAR, NT, 0) : WL;
    w.d2 = reinterpret_cast<float *>(ws + VL.d2);
    w.order = ws + VL.order;
    w.own = reinterpret_cast<float *>(ws + VL.own);
    w.raw_own = reinterpret_cast<float *>(ws + (NT ? WL.raw_own : VL.raw_own));

    // episode counters of this lane's envs (ATT/ma_main:581-637), flushed once per warp
    // lane j of the warp carries statistic j (aac_read_stats order; lane 2 = the sum of returns, a float): two registers
    // per lane instead of ten
    int st_i = 0;
    float st_f = 0.0f;

    if (MT == MT_PHASED) {
        group_loop<VAR, AUX, LEAN, NT, RT, EVS, MT_STEP_ONLY, RM, CS, 1, PL>(p, w, mode_arg, s_map, st_i, st_f);
        group_loop<VAR, AUX, LEAN, NT, RT, EVS, MODE_RESET, RM, CS, 2, PL>(p, w, mode_arg, s_map, st_i, st_f);
    } else {
        group_loop<VAR, AUX, LEAN, NT, RT, EVS, MT, RM, CS, 0, PL>(p, w, mode_arg, s_map, st_i, st_f);
    }

    // ---- warp epilogue: the counters leave, one atomic per statistic and warp
    if (mode == MODE_STEP && p.stats && lane < 10) {
        const double v = lane == 2 ? (double)st_f : (double)st_i;
        if (v != 0.0) atomicAdd(p.stats + lane, v);
    }
}

template <int VAR, bool AUX, bool LEAN, int NT, int RT, bool EVS = false, int MT = -1, int RM = -1, bool CS = false, bool PL = false>
static cudaError_t launch_one(const KParams &p, int mode, int threads, int sms, int *grid_cache, cudaStream_t stream) {
    const int wpc = threads / 32;
    const int groups = (p.E + p.G - 1) / p.G;
    auto fn = env_kernel<VAR, AUX, LEAN, NT, RT, EVS, MT, RM, CS, PL>;
    static int opted_in[64] = {0};   // dynamic shared memory this instantiation has opted in to, per device
    int dev = 0;
    cudaGetDevice(&dev);
    if ((int)p.CL.total > opted_in[dev & 63]) {   // handles of different shapes share the function attribute
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.CL.total);
        if (e != cudaSuccess) return e;
        opted_in[dev & 63] = (int)p.CL.total;
    }
    if (*grid_cache <= 0) {   // first launch of this handle: query the residency
        cudaError_t e;
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, p.CL.total);
        if (e != cudaSuccess) return e;
        *grid_cache = per_sm < 1 ? 1 : per_sm;
    }
    // persistent warps: as many CTAs as fit on the device at once (a multiple of the SM count), never more than
    // there are groups to hand out
    int grid = sms * *grid_cache;
    const int need = (groups + wpc - 1) / wpc;
    if (grid > need) grid = need;
    fn<<<grid, threads, p.CL.total, stream>>>(p, mode);
    return cudaGetLastError();
}

template <int VAR, int NT, int RT>
static cudaError_t launch_aux(const KParams &p, int mode, int threads, int sms, int *grid_cache, cudaStream_t stream) {
    if (p.out_flags == 0) {
        if (NT > 0 && RT > 0) {   // the two launches of a step (and the plain step / reset calls): one kernel per mode and radar mode
            constexpr bool V2 = VAR == AAC_VARIANT_V2;
            const bool lh = V2 && p.radar_mode == AAC_RADAR_LAST_HIT;
            if (mode == MODE_STEP && p.autoreset == 2)
                return lh ? launch_one<VAR, false, true, NT, RT, false, MT_PHASED, V2 ? AAC_RADAR_LAST_HIT : -1>(p, mode, threads, sms, grid_cache, stream)
                          : launch_one<VAR, false, true, NT, RT, false, MT_PHASED, V2 ? AAC_RADAR_MIN : -1>(p, mode, threads, sms, grid_cache, stream);
            if (mode == MODE_STEP && !p.autoreset)
                return lh ? launch_one<VAR, false, true, NT, RT, false, MT_STEP_ONLY, V2 ? AAC_RADAR_LAST_HIT : -1>(p, mode, threads, sms, grid_cache, stream)
                          : launch_one<VAR, false, true, NT, RT, false, MT_STEP_ONLY, V2 ? AAC_RADAR_MIN : -1>(p, mode, threads, sms, grid_cache, stream);
            if (mode == MODE_RESET)
                return lh ? launch_one<VAR, false, true, NT, RT, false, MODE_RESET, V2 ? AAC_RADAR_LAST_HIT : -1>(p, mode, threads, sms, grid_cache, stream)
                          : launch_one<VAR, false, true, NT, RT, false, MODE_RESET, V2 ? AAC_RADAR_MIN : -1>(p, mode, threads, sms, grid_cache, stream);
        }
        return launch_one<VAR, false, true, NT, RT>(p, mode, threads, sms, grid_cache, stream);
    }
    return (p.out_flags & AAC_OUT_RADAR_AUX) ? launch_one<VAR, true, false, NT, RT>(p, mode, threads, sms, grid_cache, stream)
                                             : launch_one<VAR, false, false, NT, RT>(p, mode, threads, sms, grid_cache, stream);
}

bool phased_launch_available(int variant, const KParams &p) {
    if (p.out_flags != 0) return false;
    switch (variant) {   // the shapes launch_env_kernel specialises on both counts
        case AAC_VARIANT_ATT: return p.N == 3 && (p.R == 36 || p.R == 18);
        case AAC_VARIANT_MM: return p.N == 3 && p.R == 18;
        case AAC_VARIANT_V2: return !p.radar_targets && !p.n_nbr_obs && !p.eval_by_step && ((p.N == 10 && p.R == 36) || (p.N == 20 && p.R == 72));
        default: return false;
    }
}

// the run-time-mode kernels with the per-episode path search compiled in (a map's origin / destination table has pools only)
template <int VAR>
static cudaError_t launch_plan(const KParams &p, int mode, int threads, int sms, int *grid_cache, cudaStream_t stream) {
    if (p.out_flags == 0) return launch_one<VAR, false, true, 0, 0, false, -1, -1, false, true>(p, mode, threads, sms, grid_cache, stream);
    return (p.out_flags & AAC_OUT_RADAR_AUX) ? launch_one<VAR, true, false, 0, 0, false, -1, -1, false, true>(p, mode, threads, sms, grid_cache, stream)
                                             : launch_one<VAR, false, false, 0, 0, false, -1, -1, false, true>(p, mode, threads, sms, grid_cache, stream);
}

cudaError_t launch_env_kernel(int variant, const KParams &p, int mode, int threads, int sms, int *grid_cache, cudaStream_t stream) {
    if (p.E <= 0) return cudaSuccess;
    if (p.autoreset == 2 && (mode != MODE_STEP || !phased_launch_available(variant, p) || !p.flags || p.plan_scratch)) return cudaErrorNotSupported;
    if (p.plan_scratch) {
        if (p.radar_targets || p.n_nbr_obs || p.eval_by_step) return cudaErrorNotSupported;   // (sensor / evaluation configurations: tables with paths)
        switch (variant) {
            case AAC_VARIANT_ATT: return launch_plan<AAC_VARIANT_ATT>(p, mode, threads, sms, grid_cache, stream);
            case AAC_VARIANT_MM: return launch_plan<AAC_VARIANT_MM>(p, mode, threads, sms, grid_cache, stream);
            case AAC_VARIANT_V2: return launch_plan<AAC_VARIANT_V2>(p, mode, threads, sms, grid_cache, stream);
            default: return cudaErrorInvalidValue;
        }
    }
    switch (variant) {
        // the reference's own shapes (3 drones; 18 rays, or the 36 of the batched configuration) are specialised too
        case AAC_VARIANT_ATT:
            if (p.N == 3 && p.R == 36) return launch_aux<AAC_VARIANT_ATT, 3, 36>(p, mode, threads, sms, grid_cache, stream);
            if (p.N == 3 && p.R == 18) return launch_aux<AAC_VARIANT_ATT, 3, 18>(p, mode, threads, sms, grid_cache, stream);
            return launch_aux<AAC_VARIANT_ATT, 0, 0>(p, mode, threads, sms, grid_cache, stream);
        case AAC_VARIANT_MM:
            if (p.N == 3 && p.R == 18) return launch_aux<AAC_VARIANT_MM, 3, 18>(p, mode, threads, sms, grid_cache, stream);
            return launch_aux<AAC_VARIANT_MM, 0, 0>(p, mode, threads, sms, grid_cache, stream);
        case AAC_VARIANT_V2:
            if (p.radar_targets || p.n_nbr_obs)   // the later fork's sensor classes: one generic instantiation per output set
                return (p.out_flags & AAC_OUT_RADAR_AUX) ? launch_one<AAC_VARIANT_V2, true, false, 0, 0, false, -1, -1, true>(p, mode, threads, sms, grid_cache, stream)
                                                         : launch_one<AAC_VARIANT_V2, false, false, 0, 0, false, -1, -1, true>(p, mode, threads, sms, grid_cache, stream);
            if (p.eval_by_step)  // evaluation mode: one generic instantiation per output set
                return (p.out_flags & AAC_OUT_RADAR_AUX) ? launch_one<AAC_VARIANT_V2, true, false, 0, 0, true>(p, mode, threads, sms, grid_cache, stream)
                                                         : launch_one<AAC_VARIANT_V2, false, false, 0, 0, true>(p, mode, threads, sms, grid_cache, stream);
            // the benchmark configurations get kernels specialised on the drone and ray counts (register-resident
            // neighbour sort, constant loop bounds and addressing); everything else runs the generic kernel
            if (p.N == 10 && p.R == 36) return launch_aux<AAC_VARIANT_V2, 10, 36>(p, mode, threads, sms, grid_cache, stream);
            if (p.N == 20 && p.R == 72) return launch_aux<AAC_VARIANT_V2, 20, 72>(p, mode, threads, sms, grid_cache, stream);
            if (p.N == 10) return launch_aux<AAC_VARIANT_V2, 10, 0>(p, mode, threads, sms, grid_cache, stream);
            if (p.N == 20) return launch_aux<AAC_VARIANT_V2, 20, 0>(p, mode, threads, sms, grid_cache, stream);
            return launch_aux<AAC_VARIANT_V2, 0, 0>(p, mode, threads, sms, grid_cache, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace aac
