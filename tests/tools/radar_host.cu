// Host build of the grid radar the kernels run (multi_agent_aac_b200/csrc/aac_radar.cuh: occupancy window, cell walk,
// generic routine, boundary lines), for tests/test_radar_host.py: the same float32 code checked against the float64
// checker without a GPU.  TEST INFRASTRUCTURE: nothing in the product loads this.
#include <vector>
#include "../../multi_agent_aac_b200/csrc/aac_radar.cuh"
#include "../../include/aac_env.h"

using namespace aac;

// the MapDev aac_set_maps builds (aac_capi.cu), for one map
static MapDev make_map(const AacMapDesc &d, const uint8_t *occ) {
    MapDev o;
    memset(&o, 0, sizeof(o));
    o.gx = d.gx; o.gy = d.gy; o.pgx = d.gx + 2 * MAP_PAD; o.pgy = d.gy + 2 * MAP_PAD;
    o.hx = 0.5f * (d.bound[1] - d.bound[0]); o.hy = 0.5f * (d.bound[3] - d.bound[2]);
    o.ox = d.origin_x; o.oy = d.origin_y;
    o.ex0 = (d.x0c - 0.5f * d.cell) - d.origin_x; o.ey0 = (d.y0c - 0.5f * d.cell) - d.origin_y;
    o.xmin_g = d.bound[0]; o.ymin_g = d.bound[2];
    o.cell = d.cell; o.inv_cell = 1.0f / d.cell;
    o.ihx = 1.0f / o.hx; o.ihy = 1.0f / o.hy;
    for (int ix = 0; ix < d.gx; ++ix)
        for (int iy = 0; iy < d.gy; ++iy)
            if (occ[ix * d.gy + iy]) {
                const int b = (ix + MAP_PAD) * o.pgy + iy + MAP_PAD;
                o.bits[b >> 5] |= 1u << (b & 31);
            }
    return o;
}

// radar of n drones at local positions (px, py): out / out_min / hit [n][n_rays]; path[n][n_rays] = 0 walk, 1 generic
// routine because of a corner tie, 2 generic routine because the window is flagged.  Returns 0.
extern "C" int radar_host(const AacMapDesc *desc, const uint8_t *occ, int n_rays, float ray_len, int last_hit, int n, const float *px,
                          const float *py, float *out, float *out_min, int *hit, int *path) {
    const MapDev mp = make_map(*desc, occ);
    std::vector<float4> rays(n_rays);
    std::vector<DdaRay> dda(n_rays);
    const int step_deg = 360 / n_rays;
    for (int k = 0; k < n_rays; ++k) {   // as aac_create builds the fan (aac_capi.cu)
        const double rad = (double)(k * step_deg) * (M_PI / 180.0);
        double c = cos(rad), s = sin(rad);
        if (fabs(c) < 1e-12) c = 0.0;
        if (fabs(s) < 1e-12) s = 0.0;
        const float dx = (float)(ray_len * c), dy = (float)(ray_len * s);
        rays[k] = make_float4(dx, dy, dx != 0.0f ? 1.0f / dx : INFINITY, dy != 0.0f ? 1.0f / dy : INFINITY);
    }
    if (n_rays % 2 == 0)
        for (int k = 0; k < n_rays / 2; ++k) {
            const float4 r = rays[k];
            rays[k + n_rays / 2] = make_float4(-r.x, -r.y, r.x != 0.0f ? -r.z : r.z, r.y != 0.0f ? -r.w : r.w);
        }
    for (int k = 0; k < n_rays; ++k) dda[k] = make_dda_ray(rays[k].x, rays[k].y, mp.cell);
    std::vector<uint4> walk(WALK_BYTES / 16);
    make_walk_table(walk.data());
    const WalkRef wref{0u, reinterpret_cast<const unsigned char *>(walk.data())};
    for (int i = 0; i < n; ++i) {
        float ax, ay, dlx, dly;
        int ixc, iyc;
        const unsigned win = build_window5(mp, px[i], py[i], ray_len, ax, ay, dlx, dly, ixc, iyc);
        for (int k = 0; k < n_rays; ++k) {
            float o = 0.0f, om = 0.0f;
            int id = -1, how = 2;
            bool done = false;
            if (!(win & WIN5_SLOW)) {
                done = cast_grid_fast<3, true>(dda[k], wref, mp.cell, make_float4(ax, ay, dlx, dly), win, ixc, iyc, mp.gx, mp.gy, ray_len, last_hit != 0, o, om, id);
                how = done ? 0 : 1;
            }
            if (!done) {
                const SlowCast sc = cast_grid_slow<true>(mp, rays[k], px[i], py[i], win, ray_len, last_hit);
                o = sc.out; om = sc.out_min; id = sc.id;
            }
            out[i * n_rays + k] = o; out_min[i * n_rays + k] = om; hit[i * n_rays + k] = id; path[i * n_rays + k] = how;
        }
    }
    return 0;
}
