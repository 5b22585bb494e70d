// Grid radar (V2:1210-1300, MM:877-971; V2 = MADDPG_ownENV_randomOD_radar_N_model_use_tdCPA_forV2/env_simulator_...,
// MM = ..._multipleMap/env_simulator_...): occupancy window, the cell walk of one ray, the generic closed-interval
// routine and the boundary lines.  Everything here compiles for the host as well (the CPU-side checks under tests/
// run the same code without a GPU).
//
// The reference asks, per ray, every occupied 10 m cell whose box overlaps the ray's box whether the 15 m segment
// intersects it, and keeps the distance to the nearest point of segment n cell boundary.  A 15 m ray from inside a
// cell crosses at most two grid lines per axis, so the cells it can touch are the (at most four) cells it ENTERS, each
// through one crossing: the crossing parameters are t_x1 < t_x2 (vertical grid lines) and t_y1 < t_y2 (horizontal
// ones); the cell entered at an x crossing lies as many rows along y as there are y crossings before it, and vice
// versa.  The occupancy of the 5 x 5 cells around the drone's own cell travels as one 25-bit word (bit 5 * r + c,
// r = column offset + 2 along x, c = row offset + 2 along y: ascending bit = ascending cell index ix * gy + iy, the
// order in which the reference's STRtree was filled), so a candidate is one bit test and "the last hit in
// query order" is the highest set bit among the entered cells.  No loop, no divergence, the same few dozen
// instructions for every ray whatever the map looks like.
//
// Exactness: the walk gives the closed-set answer of the slab test except when the ray passes exactly through a grid
// corner (two crossing parameters equal: the cells around that corner are touched in a point) or starts exactly on
// a grid line; radar_dda reports the first (returns false), build_window5 flags the second, and the caller casts such
// rays with radar_generic (closed intervals, the reference's inside-a-cell semantics).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define AAC_HD __host__ __device__ __forceinline__
#define AAC_HD_NOINLINE __host__ __device__ __noinline__
#else
#define AAC_HD inline
#define AAC_HD_NOINLINE inline
#endif
#if defined(__CUDA_ARCH__)
#define AAC_FADD(a, b) __fadd_rn((a), (b))   // never contracted into a multiply-add: every instantiation rounds alike
#define AAC_FMUL(a, b) __fmul_rn((a), (b))
#else
#define AAC_FADD(a, b) ((a) + (b))
#define AAC_FMUL(a, b) ((a) * (b))
#endif

namespace aac {

constexpr int MAP_PAD = 4;      // free cells added on every side of the occupancy bitmap
constexpr int MAP_WORDS = 64;   // 2048 bits: (gx + 8) * (gy + 8) + 32 must fit

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
    float ihx, ihy;        // 1 / hx, 1 / hy
    uint32_t bits[MAP_WORDS];
};
static_assert(sizeof(MapDev) % 16 == 0, "MapDev is moved with 16-byte bulk copies");

AAC_HD bool occupied(const MapDev &m, int ix, int iy) {
    const int px = ix + MAP_PAD, py = iy + MAP_PAD;
    if (px < 0 || py < 0 || px >= m.pgx || py >= m.pgy) return false;
    const int b = px * m.pgy + py;
    return (m.bits[b >> 5] >> (b & 31)) & 1u;
}

constexpr int WIN5_CENTRE = 12;             // bit of the drone's own cell
constexpr unsigned WIN5_BITS = 0x1FFFFFFu;  // the 25 occupancy bits of a window word
constexpr unsigned WIN5_LINE_SHIFT = 25;    // bits 25..28: boundary line L, R, B, T is within ray reach
constexpr unsigned WIN5_NEAR_BOUND = 0xFu << WIN5_LINE_SHIFT;
constexpr unsigned WIN5_SLOW = 1u << 31;    // own cell occupied / on a grid line / outside the padded grid: generic path

// per-ray constants of the walk (host-built, aac_set_maps): two 16-byte shared-memory loads per ray
constexpr float DDA_NEVER = 1e30f;   // stands in for 1 / 0: an axis-parallel ray crosses no grid line of that axis (t >> 1), and no
                                     // product with it is a NaN or an infinity
struct __align__(16) DdaRay {
    float idx, idy;      // 1 / dx, 1 / dy of the ray vector (length = ray_len); DDA_NEVER for an axis-parallel component
    float cx, cy;        // first crossing: t1 = a * id + c with c = (cell if d > 0 else 0) * id
    float stepx, stepy;  // cell * |1 / dx|, cell * |1 / dy|: the second crossing lies that much further
    unsigned quad;       // byte offset of the ray's direction quadrant in the walk table
    unsigned pad_;
};

// The walk table: which window bit each of the four crossings enters.  The crossings tx1 < tx2, ty1 < ty2 merge in one of
// six orders, named by key = 2 [ty1 < tx1] + [ty2 < tx1] + [ty1 < tx2] + [ty2 < tx2] (0 XXYY, 1 XYXY, 2 XYYX, 3 YXXY,
// 4 YXYX, 5 YYXX); the crossings that lie on the ray (t <= 1) are the first nv of the merged order.  Entry
// [quadrant][key][nv] = the window masks of the cells entered at (x1, x2, y1, y2), 0 for a crossing beyond the ray's end.
constexpr int WALK_KEYS = 6, WALK_NV = 5;
constexpr unsigned WALK_QUAD_BYTES = WALK_KEYS * WALK_NV * 16, WALK_BYTES = 4 * WALK_QUAD_BYTES;
constexpr unsigned WALK_KEY_BYTES = WALK_NV * 16;

inline void make_walk_table(uint4 *tab /* [4 * WALK_KEYS * WALK_NV] */) {
    static const char *orders[WALK_KEYS] = {"XXYY", "XYXY", "XYYX", "YXXY", "YXYX", "YYXX"};
    for (int q = 0; q < 4; ++q) {
        const int sxi = (q & 1) ? -5 : 5, syi = (q & 2) ? -1 : 1;
        for (int key = 0; key < WALK_KEYS; ++key)
            for (int nv = 0; nv < WALK_NV; ++nv) {
                unsigned m[4] = {0, 0, 0, 0};   // x1, x2, y1, y2
                int nx = 0, ny = 0;
                for (int step = 0; step < nv; ++step) {
                    const bool is_x = orders[key][step] == 'X';
                    if (is_x) ++nx; else ++ny;
                    const int pos = 12 + sxi * nx + syi * ny;
                    m[is_x ? nx - 1 : 2 + ny - 1] = 1u << pos;
                }
                tab[(q * WALK_KEYS + key) * WALK_NV + nv] = make_uint4(m[0], m[1], m[2], m[3]);
            }
    }
}

inline DdaRay make_dda_ray(float dx, float dy, float cell) {
    DdaRay r;
    r.idx = dx != 0.0f ? 1.0f / dx : DDA_NEVER;
    r.idy = dy != 0.0f ? 1.0f / dy : DDA_NEVER;
    r.cx = dx >= 0.0f ? cell * r.idx : 0.0f;
    r.cy = dy >= 0.0f ? cell * r.idy : 0.0f;
    r.stepx = cell * fabsf(r.idx);
    r.stepy = cell * fabsf(r.idy);
    r.quad = ((dx < 0.0f ? 1u : 0u) | (dy < 0.0f ? 2u : 0u)) * WALK_QUAD_BYTES;
    r.pad_ = 0;
    return r;
}

// The 5 x 5 occupancy window around the drone's cell, the reach flags of the four boundary lines and the generic-path
// flag, in one word; (ax, ay) = low corner of the drone's cell minus the drone's position; (dlx, dly) = the boundary
// line of each axis that lies within ray reach, minus the drone's position (+inf: none).
AAC_HD unsigned build_window5(const MapDev &mp, const float px, const float py, const float ray_len, float &ax, float &ay, float &dlx, float &dly,
                              int &ixc, int &iyc) {
    ixc = (int)floorf((px - mp.ex0) * mp.inv_cell);
    iyc = (int)floorf((py - mp.ey0) * mp.inv_cell);
    ax = (mp.ex0 + (float)ixc * mp.cell) - px;
    ay = (mp.ey0 + (float)iyc * mp.cell) - py;
    unsigned win = 0;
    // the walk needs the drone strictly inside its cell (a float32 quotient next to a grid line may round across it)
    bool slow = !(ax < 0.0f && ax > -mp.cell && ay < 0.0f && ay > -mp.cell);
    if (ixc - 2 < -MAP_PAD || iyc - 2 < -MAP_PAD || ixc + 2 >= mp.gx + MAP_PAD || iyc + 2 >= mp.gy + MAP_PAD) slow = true;   // left the padded grid
    else {
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            const int b = (ixc - 2 + r + MAP_PAD) * mp.pgy + iyc - 2 + MAP_PAD;
            const unsigned lo = mp.bits[b >> 5], hi = mp.bits[min((b >> 5) + 1, MAP_WORDS - 1)];
#if defined(__CUDA_ARCH__)
            const unsigned row = __funnelshift_r(lo, hi, b & 31);
#else
            const unsigned row = (unsigned)((((unsigned long long)hi << 32) | lo) >> (b & 31));
#endif
            win |= (row & 31u) << (5 * r);
        }
    }
    if ((win >> WIN5_CENTRE) & 1u) slow = true;   // inside an occupied cell: the reference measures to the EXIT point
    // a line can only be crossed if it lies within ray reach along its axis
    dlx = INFINITY; dly = INFINITY;
    if (fabsf(px + mp.hx) <= ray_len) { win |= 1u << WIN5_LINE_SHIFT; dlx = -mp.hx - px; }
    if (fabsf(px - mp.hx) <= ray_len) { win |= 2u << WIN5_LINE_SHIFT; dlx = mp.hx - px; }
    if (fabsf(py + mp.hy) <= ray_len) { win |= 4u << WIN5_LINE_SHIFT; dly = -mp.hy - py; }
    if (fabsf(py - mp.hy) <= ray_len) { win |= 8u << WIN5_LINE_SHIFT; dly = mp.hy - py; }
    // both lines of an axis in reach (a map narrower than two ray lengths) or standing exactly on a line (the ray
    // along it overlaps it): the generic routine
    const unsigned lines = win >> WIN5_LINE_SHIFT;
    if ((lines & 3u) == 3u || (lines & 12u) == 12u || dlx == 0.0f || dly == 0.0f) slow = true;
    return slow ? (win | WIN5_SLOW) : win;
}

// One ray against the occupied cells of the 5 x 5 window.  (ax, ay) strictly inside (-cell, 0).  WANT: 1 = nearest
// hit, 2 = last hit in cell-index order, 3 = both.  t_min / t_last are ray parameters (range = t * ray_len); no hit
// leaves t_min = +inf, t_last = 1.  b_* = window bit of the hit cell or -1 (AUX only).  Returns false when the ray
// passes through a grid corner (the outputs are then meaningless).
#if defined(__CUDA_ARCH__)
__device__ __forceinline__ unsigned dda_bit(int pos) {      // 1 << pos, 0 for pos >= 32 or pos < 0 (shl clamps)
    unsigned d;
    asm("shl.b32 %0, 1, %1;" : "=r"(d) : "r"(pos));
    return d;
}
__device__ __forceinline__ int dda_top(unsigned x) {        // index of the highest set bit, -1 for 0
    int d;
    asm("bfind.u32 %0, %1;" : "=r"(d) : "r"(x));
    return d;
}
#define AAC_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#else
inline unsigned dda_bit(int pos) { return (pos >= 32 || pos < 0) ? 0u : 1u << pos; }
inline int dda_top(unsigned x) { return x ? 31 - __builtin_clz(x) : -1; }
#define AAC_FMA(a, b, c) fmaf((a), (b), (c))
#endif
// x / y pairs in the packed single-precision pipe of sm_100 (FFMA2 / FADD2 / FMUL2: one instruction for both axes); component-wise
// with the same roundings anywhere else (the host build of the CPU-side checks)
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
#define AAC_FMA2(a, b, c) __ffma2_rn((a), (b), (c))
#define AAC_FADD2(a, b) __fadd2_rn((a), (b))
#define AAC_FMUL2(a, b) __fmul2_rn((a), (b))
#else
#define AAC_FMA2(a, b, c) make_float2(fmaf((a).x, (b).x, (c).x), fmaf((a).y, (b).y, (c).y))
#define AAC_FADD2(a, b) make_float2((a).x + (b).x, (a).y + (b).y)
#define AAC_FMUL2(a, b) make_float2((a).x * (b).x, (a).y * (b).y)
#endif

// the walk table as the kernels see it: a shared-memory address (one ld.shared.v4 per ray, no generic-pointer arithmetic);
// a plain pointer on the host
struct WalkRef {
    unsigned s;               // device: shared-memory address of the table
    const unsigned char *p;   // host: the table
    AAC_HD uint4 row(unsigned off) const {
#if defined(__CUDA_ARCH__)
        uint4 m;
        asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(m.x), "=r"(m.y), "=r"(m.z), "=r"(m.w) : "r"(s + off));
        return m;
#else
        return *reinterpret_cast<const uint4 *>(p + off);
#endif
    }
};

template <int WANT, bool AUX>
AAC_HD bool radar_dda(const DdaRay &r, const WalkRef walk, const float cell, const float ax, const float ay, const unsigned win, float &t_min,
                      float &t_last, int &b_min, int &b_last) {
    (void)cell;
    const float2 c1 = AAC_FMA2(make_float2(ax, ay), make_float2(r.idx, r.idy), make_float2(r.cx, r.cy));   // first crossings, both axes at once
    const float2 c2 = AAC_FADD2(c1, make_float2(r.stepx, r.stepy));
    const float tx1 = c1.x, ty1 = c1.y, tx2 = c2.x, ty2 = c2.y;
    const bool clean = (tx1 != ty1) & (tx1 != ty2) & (tx2 != ty1) & (tx2 != ty2);
    // the merge order of the crossings and how many of them lie on the ray select the row of the walk table
    unsigned off = r.quad;
    if (ty1 < tx1) off += 2 * WALK_KEY_BYTES;
    if (ty2 < tx1) off += WALK_KEY_BYTES;
    if (ty1 < tx2) off += WALK_KEY_BYTES;
    if (ty2 < tx2) off += WALK_KEY_BYTES;
    if (tx1 <= 1.0f) off += 16;
    if (tx2 <= 1.0f) off += 16;
    if (ty1 <= 1.0f) off += 16;
    if (ty2 <= 1.0f) off += 16;
    const uint4 m = walk.row(off);   // masks of the cells entered at x1, x2, y1, y2 (0: beyond the ray)
    const unsigned hits = (m.x | m.y | m.z | m.w) & win;            // positions are < 25: the flag bits of `win` are never selected
    if (WANT & 1) {
        const float inf = INFINITY;
        const float u1 = (hits & m.x) ? tx1 : inf, u2 = (hits & m.y) ? tx2 : inf, u3 = (hits & m.z) ? ty1 : inf, u4 = (hits & m.w) ? ty2 : inf;
        const float t = fminf(fminf(u1, u2), fminf(u3, u4));
        t_min = t;
        if (AUX) b_min = t == inf ? -1 : dda_top(t == u1 ? m.x : (t == u2 ? m.y : (t == u3 ? m.z : m.w)));
    }
    if (WANT & 2) {
        // highest set bit = last cell of the reference's query order that the ray hits
        const int top = dda_top(hits);
        const unsigned tb = dda_bit(top);
        float t = 1.0f;
        t = (tb & m.x) ? tx1 : t;
        t = (tb & m.y) ? tx2 : t;
        t = (tb & m.z) ? ty1 : t;
        t = (tb & m.w) ? ty2 : t;
        t_last = t;
        if (AUX) b_last = top;
    }
    return clean;
}

// Generic path (corner / grid-line ties, drone centre inside an occupied cell or outside the padded grid): tests
// every occupied cell whose box overlaps the ray's box with closed intervals, and with the reference's
// inside-the-cell semantics: the nearest point of segment n cell BOUNDARY is the exit point, or nothing (nan) when
// the whole ray is inside (V2:1258-1265).  ray = (dx, dy, 1/dx, 1/dy).  Results travel by value: reference
// parameters of a non-inlined function would pin the caller's accumulators in local memory on the hot path.
struct GenericHit {
    float shortest, sensed;
    int shortest_id, sensed_id;
};
template <bool AUX>
AAC_HD_NOINLINE GenericHit radar_generic(const MapDev &mp, float px, float py, float4 ray, float len) {
    float shortest = INFINITY, sensed = len;
    int shortest_id = -1, sensed_id = -1;
    const float ex = px + ray.x, ey = py + ray.y;
    const int ixa = (int)floorf((fminf(px, ex) - mp.ex0) * mp.inv_cell), ixb = (int)floorf((fmaxf(px, ex) - mp.ex0) * mp.inv_cell);
    const int iya = (int)floorf((fminf(py, ey) - mp.ey0) * mp.inv_cell), iyb = (int)floorf((fmaxf(py, ey) - mp.ey0) * mp.inv_cell);
    // one cell of slack: a float32 quotient next to a grid line may round across it (the slab test itself decides)
    for (int ix = max(ixa - 1, 0); ix <= min(ixb + 1, mp.gx - 1); ++ix)
        for (int iy = max(iya - 1, 0); iy <= min(iyb + 1, mp.gy - 1); ++iy) {
            if (!occupied(mp, ix, iy)) continue;
            const float x0 = mp.ex0 + ix * mp.cell, x1 = x0 + mp.cell, y0 = mp.ey0 + iy * mp.cell, y1 = y0 + mp.cell;
            float ent = -INFINITY, ext = INFINITY;
            if (ray.x != 0.0f) {
                const float t0 = (x0 - px) * ray.z, t1 = (x1 - px) * ray.z;
                ent = fminf(t0, t1); ext = fmaxf(t0, t1);
            } else if (px < x0 || px > x1) continue;
            if (ray.y != 0.0f) {
                const float t0 = (y0 - py) * ray.w, t1 = (y1 - py) * ray.w;
                ent = fmaxf(ent, fminf(t0, t1)); ext = fminf(ext, fmaxf(t0, t1));
            } else if (py < y0 || py > y1) continue;
            const float lo = fmaxf(ent, 0.0f), hi = fminf(ext, 1.0f);
            if (lo > hi) continue;
            const bool inside = px > x0 && px < x1 && py > y0 && py < y1;
            const float d = !inside ? lo * len : (ext <= 1.0f ? ext * len : NAN);
            sensed = d;
            if (AUX) sensed_id = ix * mp.gy + iy;
            if (d < shortest) { shortest = d; if (AUX) shortest_id = ix * mp.gy + iy; }
        }
    return GenericHit{shortest, sensed, shortest_id, sensed_id};
}

// the boundary lines in the reference's order L, R, B, T (V2:145-152); ray = (dx, dy, 1/dx, 1/dy).  `lines` has
// one bit per line that can be reached at all from the drone's position.
struct RadarAcc {
    float shortest, sensed;
    int shortest_id, sensed_id;
};
template <bool AUX>
AAC_HD void bound_line(const int b, const int nb, const float dd, const float pp, const float inv, const float line, const float len,
                       float &shortest, float &sensed, int &shortest_id, int &sensed_id) {
    if (dd != 0.0f) {
        const float t = (line - pp) * inv;
        if (t >= 0.0f && t <= 1.0f) {
            const float d = t * len;
            sensed = d;
            if (AUX) sensed_id = nb + b;
            if (d < shortest) { shortest = d; if (AUX) shortest_id = nb + b; }
        }
    } else if (pp == line) {  // ray runs along the boundary: GEOS returns the whole overlap
        sensed = len;
        if (AUX) sensed_id = nb + b;
        if (0.0f < shortest) { shortest = 0.0f; if (AUX) shortest_id = nb + b; }
    }
}
// any set of lines (a map narrower than two ray lengths can have both lines of an axis in reach); off the hot
// path, so the accumulators travel by value
template <bool AUX>
AAC_HD_NOINLINE RadarAcc radar_bounds_any(const MapDev &mp, float px, float py, float4 ray, float len, unsigned lines, RadarAcc acc) {
    const int nb = mp.gx * mp.gy;
#pragma unroll 1
    while (lines) {
        int b = 0;
        while (!((lines >> b) & 1u)) ++b;
        lines &= lines - 1;
        const float lim = b < 2 ? mp.hx : mp.hy;
        bound_line<AUX>(b, nb, b < 2 ? ray.x : ray.y, b < 2 ? px : py, b < 2 ? ray.z : ray.w, (b & 1) ? lim : -lim, len, acc.shortest, acc.sensed,
                        acc.shortest_id, acc.sensed_id);
    }
    return acc;
}
template <bool AUX>
AAC_HD void radar_bounds(const MapDev &mp, float px, float py, float4 ray, float len, unsigned lines, float &shortest, float &sensed,
                         int &shortest_id, int &sensed_id) {
    if ((lines & 3u) == 3u || (lines & 12u) == 12u) {
        const RadarAcc r = radar_bounds_any<AUX>(mp, px, py, ray, len, lines, RadarAcc{shortest, sensed, shortest_id, sensed_id});
        shortest = r.shortest; sensed = r.sensed; shortest_id = r.shortest_id; sensed_id = r.sensed_id;
        return;
    }
    // usual case: at most one line per axis, x first
    const int nb = AUX ? mp.gx * mp.gy : 0;
    if (lines & 3u) bound_line<AUX>((lines & 1u) ? 0 : 1, nb, ray.x, px, ray.z, (lines & 1u) ? -mp.hx : mp.hx, len, shortest, sensed, shortest_id, sensed_id);
    if (lines & 12u) bound_line<AUX>((lines & 4u) ? 2 : 3, nb, ray.y, py, ray.w, (lines & 4u) ? -mp.hy : mp.hy, len, shortest, sensed, shortest_id, sensed_id);
}

// The grid radar of one ray on the common path: the window walk, then the (at most one per axis) boundary line within
// reach, appended after the cells in the reference's order L, R, B, T (V2:145-152), all in ray-parameter space and
// without a branch.  rec = (ax, ay, dlx, dly) of build_window5.  Returns false when the ray needs the generic routine
// (corner tie); the caller must not have called it for a window flagged WIN5_SLOW.  out = the value the radar mode
// stores (last hit: V2's source behaviour; else the true minimum), out_min = the true minimum (WANT & 1),
// id = hit id of `out` (AUX): cell ix * gy + iy, gx * gy + {0..3} for a line, -1 none.
template <int WANT, bool AUX>
AAC_HD bool cast_grid_fast(const DdaRay &r, const WalkRef walk, const float cell, const float4 rec, const unsigned win, const int ixc, const int iyc, const int gx,
                           const int gy, const float len, const bool last_hit, float &out, float &out_min, int &id) {
    float t_min = INFINITY, t_last = 1.0f;
    int b_min = -1, b_last = -1;
    const bool clean = radar_dda<WANT, AUX>(r, walk, cell, rec.x, rec.y, win, t_min, t_last, b_min, b_last);
    int id_min = -1, id_last = -1;
    if (AUX) {
        id_min = b_min < 0 ? -1 : (ixc - 2 + b_min / 5) * gy + iyc - 2 + b_min % 5;
        id_last = b_last < 0 ? -1 : (ixc - 2 + b_last / 5) * gy + iyc - 2 + b_last % 5;
    }
    // boundary lines: t = (line - p) / d; an axis-parallel ray (1 / d = DDA_NEVER) or a line out of reach (+inf) is never valid
    const float2 tb = AAC_FMUL2(make_float2(rec.z, rec.w), make_float2(r.idx, r.idy));
    const float tbx = tb.x, tby = tb.y;
    const bool vx = tbx >= 0.0f && tbx <= 1.0f, vy = tby >= 0.0f && tby <= 1.0f;
    if (WANT & 2) {
        t_last = vx ? tbx : t_last;
        t_last = vy ? tby : t_last;
        if (AUX) {
            const int nb = gx * gy;
            id_last = vx ? nb + ((win >> WIN5_LINE_SHIFT) & 1u ? 0 : 1) : id_last;
            id_last = vy ? nb + ((win >> WIN5_LINE_SHIFT) & 4u ? 2 : 3) : id_last;
        }
    }
    if (WANT & 1) {
        if (AUX) {
            const int nb = gx * gy;
            if (vx && tbx < t_min) id_min = nb + ((win >> WIN5_LINE_SHIFT) & 1u ? 0 : 1);
            if (vy && tby < fminf(t_min, vx ? tbx : INFINITY)) id_min = nb + ((win >> WIN5_LINE_SHIFT) & 4u ? 2 : 3);
        }
        t_min = fminf(t_min, vx ? tbx : INFINITY);
        t_min = fminf(t_min, vy ? tby : INFINITY);
    }
    const float sensed = AAC_FMUL(t_last, len);
    if (WANT == 2) { out = sensed; out_min = sensed; return clean; }
    const float shortest = t_min == INFINITY ? len : AAC_FMUL(t_min, len);
    out_min = shortest;
    if (WANT == 1) { out = shortest; return clean; }
    out = last_hit ? sensed : shortest;
    if (AUX) id = last_hit ? id_last : id_min;
    return clean;
}

// The same ray through the generic routines (flagged window or corner tie): off the hot path, one call.
struct SlowCast {
    float out, out_min;
    int id;
};
template <bool AUX>
AAC_HD_NOINLINE SlowCast cast_grid_slow(const MapDev &mp, const float4 ray, const float px, const float py, const unsigned win, const float len,
                                        const int last_hit) {
    const GenericHit h = radar_generic<AUX>(mp, px, py, ray, len);
    float shortest = h.shortest, sensed = h.sensed;
    int shortest_id = h.shortest_id, sensed_id = h.sensed_id;
    if (win & WIN5_NEAR_BOUND) radar_bounds<AUX>(mp, px, py, ray, len, (win >> WIN5_LINE_SHIFT) & 0xFu, shortest, sensed, shortest_id, sensed_id);
    SlowCast o;
    o.out_min = shortest == INFINITY ? len : shortest;
    o.out = last_hit ? sensed : o.out_min;
    o.id = last_hit ? sensed_id : shortest_id;
    return o;
}

}  // namespace aac
