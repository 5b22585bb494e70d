// The reference's grid path search (ATT/jps_straight.py:17-70) and collinear pruning (ATT:321-331) for ONE origin / destination
// pair, run by one warp.  Shared by the table builder (plan_paths_kernel, aac_capi.cu: one warp per pair) and by the env
// kernel's re-initialisation when a map's origin / destination table carries pools but no paths (per-episode planning,
// aac_kernels.cu init_envs).
//
// The frontier keeps discovery order; a popped entry is tombstoned in place (f = 0xFFFF) instead of erased, so "the
// first minimum in discovery order" is the minimum of (f << 16 | position) over the live entries: the lanes scan the
// list in strides and meet in one redux.  A cell is queued at most once, so every array is bounded by the cell count.
// Scratch per warp (global memory, L1 / L2 resident for the reference's 23 x 13 .. 31 x 21 grids): status u8[n],
// g u16[n], parent u16[n], frontier cell u16[n], frontier f u16[n].
#pragma once
#include <cstdint>

namespace aac {

constexpr int PLAN_SCRATCH_PER_CELL = 9;
constexpr unsigned PLAN_DEAD = 0xFFFFu;

// scratch bytes one warp needs for a grid of n cells
__host__ __device__ inline size_t plan_scratch_bytes(size_t n) { return ((n + 7) & ~(size_t)7) * PLAN_SCRATCH_PER_CELL; }

// `occ(c)`: cell c = ix * gy + iy is occupied.  s, t: start and goal cell indices.  The pruned path (cells where the direction
// changes, start and goal included) is written to out[0 .. cnt) as (ix << 8 | iy) by lane 0 when cnt <= max_cells.
// Returns (to every lane) cnt, 0 when the goal is unreachable, -1 when the path has more than max_cells vertices.
template <class Occ>
__device__ __forceinline__ int plan_path_warp(const Occ occ, const int gx, const int gy, const int s, const int t, uint8_t *scratch, uint16_t *out,
                                              const int max_cells) {
    const unsigned FULLM = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int n = gx * gy;
    const size_t n_al = ((size_t)n + 7) & ~(size_t)7;
    const int tx = t / gy, ty = t - tx * gy;
    uint8_t *status = scratch;
    uint16_t *gcost = reinterpret_cast<uint16_t *>(status + n_al), *parent = gcost + n_al, *fcell = parent + n_al, *ff = fcell + n_al;
    for (int c = lane; c < n; c += 32) status[c] = 0;
    __syncwarp();
    if (lane == 0) { status[s] = 1; gcost[s] = 0; parent[s] = PLAN_DEAD; fcell[0] = (uint16_t)s; ff[0] = 0; }
    __syncwarp();
    int head = 0, tail = 1, live = 1;   // live entries lie in [head, tail)
    bool found = false;
    for (;;) {
        unsigned key = 0xFFFFFFFFu;
        for (int k = head + lane; k < tail; k += 32) {
            const unsigned f = ff[k];
            if (f != PLAN_DEAD) key = min(key, (f << 16) | (unsigned)k);
        }
        key = __reduce_min_sync(FULLM, key);
        if (key == 0xFFFFFFFFu) break;   // frontier empty: the goal is unreachable
        const int kb = (int)(key & 0xFFFFu), cur = fcell[kb];
        if (cur == t) { found = true; break; }
        const int cx = cur / gy, cy = cur - cx * gy, g1 = gcost[cur] + 1;
        __syncwarp();
        if (lane == 0) ff[kb] = (uint16_t)PLAN_DEAD;
        while (head < tail && (head == kb || ff[head] == PLAN_DEAD)) ++head;   // warp-uniform: every lane reads the same entries
        // neighbours in the reference's order (0,-1), (0,1), (-1,0), (1,0): lane d takes the d-th
        const int nx = cx + (lane == 2 ? -1 : lane == 3 ? 1 : 0), ny = cy + (lane == 0 ? -1 : lane == 1 ? 1 : 0);
        bool push = lane < 4 && nx >= 0 && ny >= 0 && nx < gx && ny < gy;
        const int c = nx * gy + ny;
        if (push) push = !occ(c) && !status[c];
        const unsigned m = __ballot_sync(FULLM, push);
        if (push) {
            const int k = tail + __popc(m & ((1u << lane) - 1u));
            status[c] = 1;
            gcost[c] = (uint16_t)g1;
            parent[c] = (uint16_t)cur;
            fcell[k] = (uint16_t)c;
            ff[k] = (uint16_t)(g1 + abs(nx - tx) + abs(ny - ty));
        }
        tail += __popc(m);
        live += __popc(m) - 1;
        __syncwarp();
        // a long search leaves tombstones between the live entries: squeeze them out, order kept, once they outnumber
        // the live ones (write position <= read position, a chunk is read before it is written)
        if (tail - head > 64 && 2 * live < tail - head) {
            int wpos = 0;
            for (int k0 = head; k0 < tail; k0 += 32) {
                const int k = k0 + lane;
                const unsigned f = k < tail ? ff[k] : PLAN_DEAD;
                const unsigned cc = k < tail ? fcell[k] : 0u;
                const unsigned keep = __ballot_sync(FULLM, f != PLAN_DEAD);
                __syncwarp();
                if (f != PLAN_DEAD) {
                    const int dst = wpos + __popc(keep & ((1u << lane) - 1u));
                    ff[dst] = (uint16_t)f;
                    fcell[dst] = (uint16_t)cc;
                }
                wpos += __popc(keep);
                __syncwarp();
            }
            head = 0;
            tail = wpos;
        }
    }
    // walk back from the goal; the cells where the direction changes are the same in either direction (ATT:321-331)
    int cnt = 0;
    if (lane == 0 && found) {
        for (int pass = 0; pass < 2; ++pass) {   // count, then write from the back
            int k = 0, c = t, pc = parent[c];
            auto emit = [&](int cell) {
                if (pass == 1 && cnt <= max_cells) out[cnt - 1 - k] = (uint16_t)(((cell / gy) << 8) | (cell % gy));
                ++k;
            };
            emit(c);
            if (pc != (int)PLAN_DEAD) {
                int d = pc - c;   // steps are +-1 or +-gy: the index difference names the direction
                for (;;) {
                    const int nc = parent[pc];
                    if (nc == (int)PLAN_DEAD) break;
                    const int e = nc - pc;
                    if (e != d) { emit(pc); d = e; }
                    pc = nc;
                }
                emit(pc);
            }
            if (pass == 0) cnt = k;
        }
    }
    cnt = __shfl_sync(FULLM, cnt, 0);
    __syncwarp();
    return cnt > max_cells ? -1 : cnt;
}

}  // namespace aac
