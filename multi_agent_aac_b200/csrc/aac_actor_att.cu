// Batched attention actor of the one_model_att variant (include/aac_actor.h, aac_actor_att_*): the whole
// ActorNetwork_ATT_TwoPortion (ATT/Nnetworks:177-213), fp32 on the CUDA cores.
//
// This network is 66 k multiply-adds per drone and the variant's batches are small (4096 envs x 3 drones), so it runs
// in plain fp32 (agreement with the float64 reference ~1e-6) instead of on the tensor cores.  A CTA of 256 threads
// takes 128 drones through the layers as a chain of register-tiled 128 x 64 GEMM tiles: activations stay in shared
// memory feature-major ([feature][drone], so both the operand reads and the result stores are conflict-free), the
// transposed weights ([in][out], built by aac_actor_att_load) are staged 32 rows at a time, and a thread accumulates
// 4 outputs x 8 drones (32 FMAs per three 16-byte shared loads).  The attention is folded algebraically so that
// nothing per-neighbour is 64 x 64:
//     score_m = k(x_m) . q = x_m . (Wk^T q),      v_att = sum_m alpha_m Wv x_m = Wv (sum_m alpha_m x_m)
// (x_m = relu(neigh_fc(nei_m)) is recomputed for the second pass instead of stored: a K = 6 tile).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/aac_actor.h"

namespace {

constexpr int HID = 64, MERGE = 256, CAT = 3 * HID, THREADS = 256, TB = 128, KC = 32;
constexpr int MAX_NEI = 31, MAX_DNEI = 32;
// shared memory (floats): cat [192][TB] | buf [128][TB] | nei [32][TB] | alpha [31][TB] | ws [KC][64]
constexpr int S_CAT = 0, S_BUF = S_CAT + CAT * TB, S_NEI = S_BUF + 2 * HID * TB, S_ALPHA = S_NEI + MAX_DNEI * TB, S_WS = S_ALPHA + MAX_NEI * TB,
              S_TOTAL = S_WS + KC * HID;

struct AttArgs {
    const float *own, *grid, *nei;
    float *actions;
    int n_rows, d_own, d_grid, d_nei, n_nei;
    // transposed weights [in][out] and biases
    const float *w_own, *b_own, *w_grid, *b_grid, *w_nei, *b_nei, *w_q, *w_k, *w_v, *w_merge, *b_merge, *w_out, *b_out;
    float noise_scale;
    unsigned long long noise_seed;
};

// ys[0..64)[0..TB) = act(W^T x + b) for the 64 output columns c0.. of wt ([n_in][ldw]); xs / ys are feature-major
// [row][TB].  All 256 threads; ends with every thread past its stores' barrier.
template <bool RELU>
__device__ __forceinline__ void gemm64(const float *__restrict__ wt, int ldw, int c0, const float *__restrict__ bias, int n_in, const float *xs,
                                       float *ys, float *ws) {
    // thread = outputs 8 og .. +7 (og = warp) x drones 4 lane .. +3: per k a warp reads one activation row as 32 distinct
    // float4 (all 128 drones) and two weight float4 that every lane shares (broadcast): 6 shared-memory wavefronts for
    // 32 x 32 FMAs
    const int tid = threadIdx.x, og = tid >> 5, lane = tid & 31;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float b = bias ? __ldg(bias + c0 + 8 * og + i) : 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = b;
    }
    // the next 32 x 64 weight slab travels global -> registers while the current one is being used
    float4 pre[2];
    auto fetch = [&](int k0) {
        const int kc = min(KC, n_in - k0);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int f = tid + u * THREADS, kk = f >> 4, c4 = f & 15;
            if (kk < kc) pre[u] = __ldg(reinterpret_cast<const float4 *>(wt + (size_t)(k0 + kk) * ldw + c0) + c4);
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < n_in; k0 += KC) {
        const int kc = min(KC, n_in - k0);
        __syncthreads();  // the previous slab (or the caller's writes to xs) is done with
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int f = tid + u * THREADS;
            if ((f >> 4) < kc) reinterpret_cast<float4 *>(ws)[f] = pre[u];
        }
        __syncthreads();
        if (k0 + KC < n_in) fetch(k0 + KC);
#pragma unroll 4
        for (int kk = 0; kk < kc; ++kk) {
            const float4 w0 = reinterpret_cast<const float4 *>(ws)[kk * 16 + 2 * og], w1 = reinterpret_cast<const float4 *>(ws)[kk * 16 + 2 * og + 1];
            const float4 x = reinterpret_cast<const float4 *>(xs + (size_t)(k0 + kk) * TB)[lane];
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[i], xv[j], acc[i][j]);
        }
    }
    __syncthreads();  // ys may alias xs
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        float4 o;
        o.x = RELU ? fmaxf(acc[i][0], 0.f) : acc[i][0]; o.y = RELU ? fmaxf(acc[i][1], 0.f) : acc[i][1];
        o.z = RELU ? fmaxf(acc[i][2], 0.f) : acc[i][2]; o.w = RELU ? fmaxf(acc[i][3], 0.f) : acc[i][3];
        reinterpret_cast<float4 *>(ys + (size_t)(8 * og + i) * TB)[lane] = o;
    }
    __syncthreads();
}

// rows [r0, r0 + TB) x d of a row-major global block -> feature-major shared rows (zero beyond n_rows)
__device__ __forceinline__ void stage_T(const float *src, long long r0, int n_rows, int d, int stride, float *dst) {
    for (int f = threadIdx.x; f < TB * d; f += THREADS) {
        const int r = f / d, k = f - r * d;
        dst[k * TB + r] = (r0 + r < n_rows) ? __ldg(src + (r0 + r) * stride + k) : 0.f;
    }
}

__device__ __forceinline__ unsigned mix32(unsigned x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ float2 normal_pair(unsigned long long seed, unsigned row) {
    const unsigned k = mix32((unsigned)seed ^ mix32((unsigned)(seed >> 32) + 0x9e3779b9u) ^ mix32(row * 2u + 1u));
    const float u1 = ((mix32(k ^ 0x68bc21ebu) >> 8) + 1) * (1.0f / 16777216.0f), u2 = (mix32(k ^ 0x02e5be93u) >> 8) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307f * u2, &s, &c);
    return make_float2(r * c, r * s);
}

__global__ void __launch_bounds__(THREADS, 1) actor_att_kernel(const __grid_constant__ AttArgs p) {
    extern __shared__ __align__(16) float smem[];
    float *cat = smem + S_CAT, *buf = smem + S_BUF, *neis = smem + S_NEI, *alpha = smem + S_ALPHA, *ws = smem + S_WS;
    const int tid = threadIdx.x, d = tid & (TB - 1), hf = tid >> 7;   // drone of the tile, half of a per-drone loop
    const long long r0 = (long long)blockIdx.x * TB;
    // own_fc, own_grid (ATT/Nnetworks:193-194) -> cat rows 0..63, 64..127
    stage_T(p.own, r0, p.n_rows, p.d_own, p.d_own, buf);
    gemm64<true>(p.w_own, HID, 0, p.b_own, p.d_own, buf, cat, ws);
    stage_T(p.grid, r0, p.n_rows, p.d_grid, p.d_grid, buf);
    gemm64<true>(p.w_grid, HID, 0, p.b_grid, p.d_grid, buf, cat + HID * TB, ws);
    // q = Wq own_obs -> buf rows 0..63; qk = Wk^T q -> buf rows 64..127 (w_k is kept untransposed = the [in][out] form of Wk^T)
    gemm64<false>(p.w_q, HID, 0, nullptr, HID, cat, buf, ws);
    gemm64<false>(p.w_k, HID, 0, nullptr, HID, buf, buf + HID * TB, ws);
    // scores over the neighbours (ATT/Nnetworks:195-203): masked where the neighbour row averages to zero
    const int nstride = p.n_nei * p.d_nei;
    for (int m = 0; m < p.n_nei; ++m) {
        stage_T(p.nei + m * p.d_nei, r0, p.n_rows, p.d_nei, nstride, neis);
        gemm64<true>(p.w_nei, HID, 0, p.b_nei, p.d_nei, neis, buf, ws);          // x_m -> buf rows 0..63
        float s = 0.f;
        for (int r = hf * 32; r < hf * 32 + 32; ++r) s = fmaf(buf[(HID + r) * TB + d], buf[r * TB + d], s);
        if (hf) ws[d] = s;
        __syncthreads();
        if (!hf) {
            float sum = 0.f;
            for (int k = 0; k < p.d_nei; ++k) sum += neis[k * TB + d];
            alpha[m * TB + d] = (sum / (float)p.d_nei != 0.f) ? (s + ws[d]) * 0.125f : -INFINITY;   // / sqrt(64)
        }
        __syncthreads();
    }
    // softmax, masked entries zero (ATT/Nnetworks:205-207); every neighbour masked: v_att = 0
    if (!hf) {
        float smax = -INFINITY, denom = 0.f;
        for (int m = 0; m < p.n_nei; ++m) smax = fmaxf(smax, alpha[m * TB + d]);
        for (int m = 0; m < p.n_nei; ++m) {
            const float sc = alpha[m * TB + d], e = sc == -INFINITY ? 0.f : __expf(sc - smax);
            alpha[m * TB + d] = e;
            denom += e;
        }
        const float inv = denom > 0.f ? 1.0f / denom : 0.f;
        for (int m = 0; m < p.n_nei; ++m) alpha[m * TB + d] *= inv;
    }
    for (int r = hf * 32; r < hf * 32 + 32; ++r) buf[(HID + r) * TB + d] = 0.f;   // qk is free now: xbar = sum alpha_m x_m
    __syncthreads();
    for (int m = 0; m < p.n_nei; ++m) {
        stage_T(p.nei + m * p.d_nei, r0, p.n_rows, p.d_nei, nstride, neis);
        gemm64<true>(p.w_nei, HID, 0, p.b_nei, p.d_nei, neis, buf, ws);
        const float al = alpha[m * TB + d];
        for (int r = hf * 32; r < hf * 32 + 32; ++r) buf[(HID + r) * TB + d] = fmaf(al, buf[r * TB + d], buf[(HID + r) * TB + d]);
    }
    gemm64<false>(p.w_v, HID, 0, nullptr, HID, buf + HID * TB, cat + 2 * HID * TB, ws);   // v_att -> cat rows 128..191
    // merge_feature in four tiles of 64 neurons, each folded into the two outputs at once
    float a0 = 0.f, a1 = 0.f;
    for (int qd = 0; qd < MERGE / HID; ++qd) {
        gemm64<true>(p.w_merge, MERGE, qd * HID, p.b_merge, CAT, cat, buf, ws);
        for (int r = hf * 32; r < hf * 32 + 32; ++r) {
            const float h = buf[r * TB + d];
            a0 = fmaf(h, __ldg(p.w_out + qd * HID + r), a0);
            a1 = fmaf(h, __ldg(p.w_out + MERGE + qd * HID + r), a1);
        }
    }
    __syncthreads();
    if (hf) { ws[d] = a0; ws[TB + d] = a1; }
    __syncthreads();
    const long long row = r0 + d;
    if (!hf && row < p.n_rows) {
        a0 = tanhf(a0 + ws[d] + __ldg(p.b_out));
        a1 = tanhf(a1 + ws[TB + d] + __ldg(p.b_out + 1));
        if (p.noise_scale != 0.0f) {  // choose_action: act += randn(2) * var, clamp (ATT/maddpg_agent:497-501)
            const float2 z = normal_pair(p.noise_seed, (unsigned)row);
            a0 = fminf(fmaxf(fmaf(p.noise_scale, z.x, a0), -1.0f), 1.0f);
            a1 = fminf(fmaxf(fmaf(p.noise_scale, z.y, a1), -1.0f), 1.0f);
        }
        reinterpret_cast<float2 *>(p.actions)[row] = make_float2(a0, a1);
    }
}

thread_local char g_err_att[512] = "";
int fail(int code, const char *msg) { snprintf(g_err_att, sizeof(g_err_att), "%s", msg); return code; }
int cuda_fail(cudaError_t e, const char *what) { snprintf(g_err_att, sizeof(g_err_att), "%s: %s", what, cudaGetErrorString(e)); return AAC_ACTOR_ERR_CUDA; }

}  // namespace

struct AacActorAtt {
    AacActorAttConfig cfg;
    float *d_params = nullptr;
    size_t off[13] = {0};
    bool loaded = false;
    int64_t launches = 0;
};

extern "C" const char *aac_actor_att_last_error(void) { return g_err_att; }

extern "C" int aac_actor_att_create(const AacActorAttConfig *cfg, AacActorAtt **out) {
    if (!cfg || !out) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_att_create: null argument");
    if (cfg->abi_version != AAC_ACTOR_ABI_VERSION) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_att_create: abi_version mismatch");
    if (cfg->d_own < 1 || cfg->d_own > 96 || cfg->d_grid < 1 || cfg->d_grid > 96 || cfg->d_nei < 1 || cfg->d_nei > 32)
        return fail(AAC_ACTOR_ERR_ARG, "aac_actor_att_create: input widths out of range (own / grid 1..96, nei 1..32)");
    if (cfg->n_nei < 0 || cfg->n_nei > MAX_NEI) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_att_create: n_nei outside 0..31");
    AacActorAtt *a = new (std::nothrow) AacActorAtt();
    if (!a) return fail(AAC_ACTOR_ERR_STATE, "aac_actor_att_create: out of memory");
    a->cfg = *cfg;
    const cudaError_t e = cudaFuncSetAttribute(actor_att_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_TOTAL * 4);
    if (e != cudaSuccess) { delete a; return cuda_fail(e, "aac_actor_att_create: cudaFuncSetAttribute"); }
    *out = a;
    return 0;
}

extern "C" void aac_actor_att_destroy(AacActorAtt *a) {
    if (!a) return;
    cudaFree(a->d_params);
    delete a;
}

extern "C" int aac_actor_att_load(AacActorAtt *a, const AacActorAttParams *hp) {
    if (!a || !hp) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_att_load: null argument");
    const float *ptrs[13] = {hp->w_own, hp->b_own, hp->w_grid, hp->b_grid, hp->w_nei, hp->b_nei, hp->w_q, hp->w_k, hp->w_v, hp->w_merge, hp->b_merge, hp->w_out, hp->b_out};
    for (const float *q : ptrs)
        if (!q) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_att_load: null parameter pointer");
    std::vector<float> buf;
    auto put = [&](int slot, const float *w, int n_out, int n_in, bool transpose) {
        while (buf.size() % 4) buf.push_back(0.f);   // 16-byte alignment of every block
        a->off[slot] = buf.size();
        if (!transpose) { buf.insert(buf.end(), w, w + (size_t)n_out * n_in); return; }
        for (int k = 0; k < n_in; ++k)
            for (int r = 0; r < n_out; ++r) buf.push_back(w[(size_t)r * n_in + k]);
    };
    const AacActorAttConfig &c = a->cfg;
    put(0, hp->w_own, HID, c.d_own, true);   put(1, hp->b_own, 1, HID, false);
    put(2, hp->w_grid, HID, c.d_grid, true); put(3, hp->b_grid, 1, HID, false);
    put(4, hp->w_nei, HID, c.d_nei, true);   put(5, hp->b_nei, 1, HID, false);
    put(6, hp->w_q, HID, HID, true);
    put(7, hp->w_k, HID, HID, false);        // Wk^T in [in][out] form is Wk itself
    put(8, hp->w_v, HID, HID, true);
    put(9, hp->w_merge, MERGE, CAT, true);   put(10, hp->b_merge, 1, MERGE, false);
    put(11, hp->w_out, 1, 2 * MERGE, false); put(12, hp->b_out, 1, 2, false);
    cudaFree(a->d_params);
    a->d_params = nullptr;
    a->loaded = false;
    cudaError_t e = cudaMalloc(&a->d_params, buf.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(a->d_params, buf.data(), buf.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "aac_actor_att_load: upload");
    a->loaded = true;
    return 0;
}

extern "C" int aac_actor_att_forward(AacActorAtt *a, const float *own, const float *grid, const float *nei, int32_t n_rows, float noise_scale,
                                     uint64_t noise_seed, float *actions, void *stream) {
    if (!a) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_att_forward: null handle");
    if (!a->loaded) return fail(AAC_ACTOR_ERR_STATE, "aac_actor_att_forward: no parameters loaded (aac_actor_att_load)");
    if (!own || !grid || !actions || (!nei && a->cfg.n_nei > 0)) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_att_forward: null pointer");
    if (n_rows < 0) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_att_forward: n_rows < 0");
    if (n_rows == 0) return 0;
    AttArgs k;
    k.own = own; k.grid = grid; k.nei = nei; k.actions = actions; k.n_rows = n_rows;
    k.d_own = a->cfg.d_own; k.d_grid = a->cfg.d_grid; k.d_nei = a->cfg.d_nei; k.n_nei = a->cfg.n_nei;
    const float *b = a->d_params;
    k.w_own = b + a->off[0]; k.b_own = b + a->off[1]; k.w_grid = b + a->off[2]; k.b_grid = b + a->off[3];
    k.w_nei = b + a->off[4]; k.b_nei = b + a->off[5]; k.w_q = b + a->off[6]; k.w_k = b + a->off[7]; k.w_v = b + a->off[8];
    k.w_merge = b + a->off[9]; k.b_merge = b + a->off[10]; k.w_out = b + a->off[11]; k.b_out = b + a->off[12];
    k.noise_scale = noise_scale; k.noise_seed = noise_seed;
    actor_att_kernel<<<(n_rows + TB - 1) / TB, THREADS, S_TOTAL * 4, (cudaStream_t)stream>>>(k);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "aac_actor_att_forward: launch");
    ++a->launches;
    return 0;
}

extern "C" int64_t aac_actor_att_launch_count(const AacActorAtt *a) { return a ? a->launches : 0; }
