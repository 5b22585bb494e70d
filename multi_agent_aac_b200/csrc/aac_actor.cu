// Batched actor forward (include/aac_actor.h): one persistent, warp-specialised sm_100a kernel that runs the
// whole ActorNetwork_allnei_wRadar (V2/Nnetworks:273-340) for 128 drones per tile on the 5th-generation
// tensor cores.
//
//   warps 0-15 stage the tile's observations as the bf16 A operand, then act as the epilogue of every layer
//              (warp w owns rows 32 (w % 4) .. and column group w / 4; four warps per scheduler hide the latencies):
//              tcgen05.ld the fp32 accumulators out of tensor memory, add the bias, LeakyReLU, round to bf16
//              and write the result back to shared memory AS THE NEXT LAYER'S A OPERAND (activations never
//              leave the SM); the last layer (256 -> 2), tanh, exploration noise and clamp run in fp32 here
//   warp 16    streams the pre-tiled bf16 weights through a 3-slot shared-memory ring with bulk async copies
//              (cp.async.bulk + mbarrier complete_tx); also owns the tensor-memory allocation
//   warp 17    one elected thread issues tcgen05.mma (M = 128, N = 128 / 256, K = 16, kind::f16 with bf16
//              operands, fp32 accumulators in TMEM) and tcgen05.commit's ring slots / layer completion
//
// Operand layout: the canonical K-major 128-byte-swizzled UMMA layout.  An operand of `rows` rows is cut into
// K blocks of 64 elements; a block is `rows` rows of 128 bytes, and the 16-byte chunk c of row r sits at chunk
// position c ^ (r % 8):  byte = block * rows * 128 + r * 128 + ((k / 8 % 8) ^ (r % 8)) * 16 + (k % 8) * 2
// (stride byte offset = 1024 = 8 rows; a K = 16 step advances the descriptor's start address by 32 bytes).
// An epilogue thread (= one row) writes whole 16-byte chunks, and the swizzle spreads the eight rows of a
// store phase over all banks: conflict-free 128-bit shared stores; the host tiles the weights into the same
// form once (aac_actor_load).  (A first version used the no-swizzle interleaved layout: correct, but the
// tensor core fetched those operands at ~1/3 of the rate: 325 instead of ~130 cycles per 128x256x16 MMA.)
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/aac_actor.h"

namespace {

constexpr int TILE_M = 128;
constexpr int H1 = AAC_ACTOR_H1, H1C = 3 * AAC_ACTOR_H1, H2 = AAC_ACTOR_H2, H3 = AAC_ACTOR_H3;
constexpr int BLK_BYTES = TILE_M * 128;          // one 64-wide K block of a 128-row A operand
constexpr int ACT_BYTES = (H2 / 64) * BLK_BYTES;  // widest A operand: 128 x 512 bf16 = 128 KB
constexpr int SLOT_BYTES = 32768;                // one 64-wide K block of a 256-row B operand
constexpr int N_SLOTS = 3;
constexpr int NPART = 4;  // column groups of the epilogue: 4 warps (the four TMEM lane windows) each
constexpr int EPI_THREADS = 128 * NPART, PRODUCER_WARP = 4 * NPART, MMA_WARP = 4 * NPART + 1, THREADS = EPI_THREADS + 64;
constexpr int SCRATCH_OFF = 7 * BLK_BYTES;  // last K block of the A buffer: free between the layer-3 MMAs and the next layer-2 epilogue
constexpr int MAX_CHUNKS = 64;
constexpr unsigned TMEM_COLS = 512;
constexpr unsigned CH_LAYER_END = 2u, CH_FRESH = 4u, CH_HALF_END = 8u;
// `ready` barriers the MMA thread waits on before a chunk: the A operand blocks it reads are written and the
// accumulator columns it overwrites are drained (one completion of each per tile)
constexpr int RDY_STAGE = 0, RDY_E1 = 1 /* .. 3: own | nbr | grid third of layer 1's output */, RDY_E2 = 4 /* first half of layer 2's output, 5 .. 6: the two quarters of its second half */, N_READY = 7;
constexpr int L1_COL0 = 128;  // layer 1 accumulates in TMEM columns 128 .. 511 so that columns 0 .. 255 are free for layer 2's first half as soon as the `own` third is drained

// one ring slot's worth of weights and the MMAs that consume it
struct Chunk {
    uint32_t src_off;   // byte offset into the packed weights
    uint32_t bytes;     // multiple of 16
    uint32_t n_mma;     // K = 16 steps in this chunk
    uint32_t idesc;     // tcgen05 instruction descriptor (N of this chunk)
    uint32_t wait_rdy;  // 1 + index of the ready barrier to pass first, 0 = none
    uint32_t a_off16;   // (byte offset of the A operand's K block in the A buffer) >> 4
    uint32_t tmem_col;  // accumulator column
    uint32_t flags;
};

struct KArgs {
    const float *own, *nbr, *grid;
    float *actions, *dbg;
    int dbg_layer, n_rows, n_tiles;
    int d_own, d_nbr, d_grid;
    int blk_nbr0, blk_grid0, blk_end;  // first K block of the nbr / grid part of the input operand, total blocks
    const uint8_t *wpack;
    const Chunk *sched;
    int n_chunks;
    const float *b1, *b2, *b3, *w4, *b4;
    float noise_scale;
    unsigned long long noise_seed;
    int raw_off;      // byte offset in the A buffer for bulk-fetched fp32 observations, 0 = fetch rows with plain loads
    long long *prof;  // optional [gridDim.x][8] phase clocks of epilogue thread 0 (tuning aid)
};

constexpr int SMEM_BARS = ACT_BYTES + N_SLOTS * SLOT_BYTES;
constexpr int SMEM_SCHED = SMEM_BARS + 128;
constexpr int SMEM_TOTAL = SMEM_SCHED + MAX_CHUNKS * (int)sizeof(Chunk);

// ------------------------------------------------------------------------------------ PTX wrappers

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ bool mbar_try(unsigned bar_addr, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __noinline__ void mbar_wait_slow(unsigned bar_addr, unsigned parity) {
    const long long t0 = clock64();
    while (!mbar_try(bar_addr, parity))
        if (clock64() - t0 > 4000000000ll) __trap();
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    const unsigned a = smem_u32(bar);
    if (!mbar_try(a, parity)) mbar_wait_slow(a, parity);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ bool elect_one() {   // one lane of the (converged) warp
    unsigned pred;
    asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(unsigned *slot, unsigned cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(unsigned addr, unsigned cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// shared-memory matrix descriptor, K-major SWIZZLE_128B (cute::UMMA::SmemDescriptor: start >> 4 at bit 0,
// leading byte offset (unused for this layout, 1) at bit 16, stride byte offset >> 4 = 1024 >> 4 at bit 32,
// version 1 at bit 46, layout type 2 at bit 61)
__device__ __forceinline__ unsigned long long umma_desc(unsigned addr) {
    return (unsigned long long)((addr & 0x3FFFFu) >> 4) | (1ull << 16) | ((unsigned long long)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor, kind::f16: D = f32 (bit 4), A = B = bf16 (bits 7, 10), both K-major, N >> 3 at bit 17,
// M >> 4 at bit 24 (cute::UMMA::InstrDescriptor)
__host__ __device__ inline unsigned umma_idesc(unsigned n) { return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((TILE_M >> 4) << 24); }
__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long a, unsigned long long b, unsigned idesc, unsigned accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(a), "l"(b),
        "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier once every MMA issued so far by this thread has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void umma_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of warp w gets row 32 * (w % 4) + t.  Issue only: the
// registers are valid after tmem_ld_wait(), so the next block's load can fly while this one is processed.
__device__ __forceinline__ void tmem_ld32_issue(unsigned taddr, unsigned (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, "
        "%21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }  // the epilogue warps only

__device__ __forceinline__ unsigned pack_bf16(float lo, float hi) {
    unsigned r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// (x0, x1) <- leaky((x0, x1) + (b0, b1)) with the packed fp32 pair instructions of sm_100 (FADD2 / FMUL2): the same
// roundings as the scalar form, two thirds of the instructions
__device__ __forceinline__ void bias_leaky2(unsigned &x0, unsigned &x1, float b0, float b1) {
    unsigned long long a, b, y, t;
    const unsigned long long c = 0x3c23d70a3c23d70aull;  // (0.01f, 0.01f)
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(x0), "r"(x1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(y) : "l"(a), "l"(b));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(y), "l"(c));
    float y0, y1, t0, t1;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(y));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(t));
    x0 = __float_as_uint(fmaxf(y0, t0));
    x1 = __float_as_uint(fmaxf(y1, t1));
}

// counter-based standard normal pair for (seed, row): Box-Muller on two hashed uniforms
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

// ------------------------------------------------------------------------------------ epilogue pieces

// byte offset of the 16-byte chunk holding columns 8c .. 8c+7 of row m in a 128-row K-major SWIZZLE_128B operand
__device__ __forceinline__ int a_chunk(int c, int m) { return (c >> 3) * BLK_BYTES + m * 128 + (((c & 7) ^ (m & 7)) << 4); }

// 32 columns c0 * 8 .. of one row of one observation part -> bf16 chunks of the A operand, K blocks blk0 .. (zero
// padded to a multiple of 16 columns); the batch's loads are in flight together
template <bool FROM_SMEM>
__device__ __forceinline__ void stage_batch(uint8_t *act, const float *src, int d, int blk0, int c0, int n_chunks, int m, bool valid) {
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        const int k = c0 * 8 + i;
        v[i] = (valid && k < d) ? (FROM_SMEM ? src[k] : __ldg(src + k)) : 0.0f;
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
        if (c0 + g < n_chunks)
            *reinterpret_cast<uint4 *>(act + a_chunk(blk0 * 8 + c0 + g, m)) = make_uint4(
                pack_bf16(v[8 * g], v[8 * g + 1]), pack_bf16(v[8 * g + 2], v[8 * g + 3]), pack_bf16(v[8 * g + 4], v[8 * g + 5]), pack_bf16(v[8 * g + 6], v[8 * g + 7]));
}
// the three parts of a row (own | nbr | grid) as batches of 32 columns, dealt round-robin to the NPART warp groups
template <bool FROM_SMEM>
__device__ __forceinline__ void stage_row(uint8_t *act, const float *own, const float *nbr, const float *grid, const KArgs &p, int part, int m, bool valid) {
    const float *src[3] = {own, nbr, grid};
    const int d[3] = {p.d_own, p.d_nbr, p.d_grid}, blk0[3] = {0, p.blk_nbr0, p.blk_grid0};
    int item = 0;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const int n_chunks = ((d[q] + 15) >> 4) << 1;
        for (int c0 = 0; c0 < n_chunks; c0 += 4, ++item)
            if (item % NPART == part) stage_batch<FROM_SMEM>(act, src[q], d[q], blk0[q], c0, n_chunks, m, valid);
    }
}
__device__ __forceinline__ void prefetch_row(const float *src, int d) {
    const char *q = reinterpret_cast<const char *>(src);
    for (int off = 0; off < d * 4 + 127; off += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + (off < d * 4 ? off : d * 4 - 1)));
}

// accumulators [row, col0 .. col0 + 32 NBLK) -> leaky(acc + bias) -> bf16 A operand of the next layer (columns of
// the layer's output; they sit at TMEM column col0 + tshift).  The biases of a 32-column block are fetched while
// its TMEM load is in flight; latency is otherwise hidden by the four epilogue warps per scheduler.
template <int NBLK, bool DBG>
__device__ __forceinline__ void hidden_epilogue(uint8_t *act, unsigned tmem_row, int tshift, int col0, const float *bias, int m, float *dbg_row) {
    tmem_row += tshift;
#pragma unroll
    for (int u = 0; u < NBLK; ++u) {
        const int c = col0 + 32 * u;
        unsigned x[32];
        tmem_ld32_issue(tmem_row + c, x);
        float4 bv[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) bv[g] = __ldg(reinterpret_cast<const float4 *>(bias + c + 4 * g));
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 b0 = bv[2 * g], b1 = bv[2 * g + 1];
            bias_leaky2(x[8 * g], x[8 * g + 1], b0.x, b0.y);
            bias_leaky2(x[8 * g + 2], x[8 * g + 3], b0.z, b0.w);
            bias_leaky2(x[8 * g + 4], x[8 * g + 5], b1.x, b1.y);
            bias_leaky2(x[8 * g + 6], x[8 * g + 7], b1.z, b1.w);
            auto f = [&](int i) { return __uint_as_float(x[8 * g + i]); };
            *reinterpret_cast<uint4 *>(act + a_chunk((c >> 3) + g, m)) =
                make_uint4(pack_bf16(f(0), f(1)), pack_bf16(f(2), f(3)), pack_bf16(f(4), f(5)), pack_bf16(f(6), f(7)));
            if (DBG && dbg_row)
                for (int i = 0; i < 8; ++i) dbg_row[c + 8 * g + i] = f(i);
        }
    }
}

// The same, in two steps: the accumulators are converted into registers (`held`) while the MMAs of the layer's other
// half still read the A buffer, and stored once they are done.
template <int NBLK, bool DBG>
__device__ __forceinline__ void hidden_epilogue_hold(unsigned tmem_row, int col0, const float *bias, uint4 (&held)[4 * NBLK], float *dbg_row) {
#pragma unroll
    for (int u = 0; u < NBLK; ++u) {
        const int c = col0 + 32 * u;
        unsigned x[32];
        tmem_ld32_issue(tmem_row + c, x);
        float4 bv[8];
#pragma unroll
        for (int g = 0; g < 8; ++g) bv[g] = __ldg(reinterpret_cast<const float4 *>(bias + c + 4 * g));
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const float4 b0 = bv[2 * g], b1 = bv[2 * g + 1];
            bias_leaky2(x[8 * g], x[8 * g + 1], b0.x, b0.y);
            bias_leaky2(x[8 * g + 2], x[8 * g + 3], b0.z, b0.w);
            bias_leaky2(x[8 * g + 4], x[8 * g + 5], b1.x, b1.y);
            bias_leaky2(x[8 * g + 6], x[8 * g + 7], b1.z, b1.w);
            auto f = [&](int i) { return __uint_as_float(x[8 * g + i]); };
            held[4 * u + g] = make_uint4(pack_bf16(f(0), f(1)), pack_bf16(f(2), f(3)), pack_bf16(f(4), f(5)), pack_bf16(f(6), f(7)));
            if (DBG && dbg_row)
                for (int i = 0; i < 8; ++i) dbg_row[c + 8 * g + i] = f(i);
        }
    }
}
template <int NBLK>
__device__ __forceinline__ void hidden_epilogue_store(uint8_t *act, int col0, int m, const uint4 (&held)[4 * NBLK]) {
#pragma unroll
    for (int q = 0; q < 4 * NBLK; ++q) *reinterpret_cast<uint4 *>(act + a_chunk((col0 >> 3) + q, m)) = held[q];
}

// ------------------------------------------------------------------------------------ kernel

template <bool DBG>
__global__ void __launch_bounds__(THREADS, 1) actor_kernel(const __grid_constant__ KArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t *act = smem, *ring = smem + ACT_BYTES;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + SMEM_BARS);
    unsigned long long *full = bars, *empty = bars + N_SLOTS, *ready = bars + 2 * N_SLOTS, *layer_done = ready + N_READY, *in_ready = layer_done + 1,
                       *half_done = in_ready + 1;
    unsigned *tmem_slot = reinterpret_cast<unsigned *>(half_done + 1);
    Chunk *sched = reinterpret_cast<Chunk *>(smem + SMEM_SCHED);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < p.n_chunks * (int)(sizeof(Chunk) / 4); i += THREADS) reinterpret_cast<uint32_t *>(sched)[i] = reinterpret_cast<const uint32_t *>(p.sched)[i];
    if (tid == 0) {
        for (int s = 0; s < N_SLOTS; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        for (int i = 0; i < N_READY; ++i) mbar_init(ready + i, EPI_THREADS / 32);  // one arrival per epilogue warp
        mbar_init(layer_done, 1);
        mbar_init(in_ready, 1);
        mbar_init(half_done, 1);
        fence_barrier_init();
    }
    if (warp == PRODUCER_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem_base = *tmem_slot;

    if (warp == PRODUCER_WARP) {
        if (lane == 0) {
            unsigned s = 0, ph = 0;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x)
                for (int c = 0; c < p.n_chunks; ++c) {
                    mbar_wait(empty + s, ph ^ 1u);  // slot free (passes at once the first time round)
                    mbar_expect_tx(full + s, sched[c].bytes);
                    bulk_g2s(ring + s * SLOT_BYTES, p.wpack + sched[c].src_off, sched[c].bytes, full + s);
                    if (++s == N_SLOTS) { s = 0; ph ^= 1u; }
                }
        }
    } else if (warp == MMA_WARP) {
        {
            // One thread feeds the tensor core, so its instruction stream is kept short: descriptors are a constant
            // plus (address >> 4), the per-chunk fields come precomputed from the host's schedule, ring slot and
            // parity are running counters.  The whole warp walks the loop (warp-uniform values live in uniform
            // registers, which is where tcgen05.mma takes its operands from); one elected lane issues.
            const bool issuer = elect_one();
            unsigned s = 0, ph = 0, tile_par = 0;
            const unsigned long long desc_hi = umma_desc(0) & 0xFFFFFFFF00000000ull;
            const unsigned desc_lo0 = (unsigned)umma_desc(0);  // leading-byte-offset field; the start address is added below
            const unsigned a_base = (smem_u32(act) & 0x3FFFFu) >> 4, ring_base = (smem_u32(ring) & 0x3FFFFu) >> 4;
            for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x, tile_par ^= 1u)
                for (int c = 0; c < p.n_chunks; ++c) {
                    const uint4 c0 = reinterpret_cast<const uint4 *>(sched + c)[0], c1 = reinterpret_cast<const uint4 *>(sched + c)[1];
                    const unsigned n_mma = c0.z, idesc = c0.w, wait_rdy = c1.x, a_off16 = c1.y, tmem_col = c1.z, flags = c1.w;
                    if (wait_rdy) mbar_wait(ready + wait_rdy - 1, tile_par);  // A operand blocks written, accumulator columns drained
                    mbar_wait(full + s, ph);
                    tc_fence_after();
                    const unsigned a_lo = desc_lo0 + a_base + a_off16, b_lo = desc_lo0 + ring_base + s * (SLOT_BYTES >> 4);
                    const unsigned d = tmem_base + tmem_col;
                    if (issuer) {
                        if (p.prof && blockIdx.x == 0) p.prof[gridDim.x * 8 + 64 + c] = clock64();  // issue time of chunk c (the last tile's survives)
                        umma_bf16(d, desc_hi | a_lo, desc_hi | b_lo, idesc, (flags & CH_FRESH) ? 0u : 1u);
                        for (unsigned j = 1; j < n_mma; ++j) umma_bf16(d, desc_hi | (a_lo + 2 * j), desc_hi | (b_lo + 2 * j), idesc, 1u);
                        umma_commit(empty + s);
                        if (flags & CH_LAYER_END) umma_commit(layer_done);
                        if (flags & CH_HALF_END) umma_commit(half_done);   // layer 2's first output half is complete
                    }
                    __syncwarp();
                    if (++s == N_SLOTS) { s = 0; ph ^= 1u; }
                }
        }
    } else {
        // epilogue warps: warp w works on TMEM lanes 32 * (w % 4) .. (the hardware's lane window of a warp), i.e. row
        // m = tid % 128, and on column group tid / 128 of every piece of every layer
        const int m = tid & 127, part = tid >> 7;
        const unsigned tmem_row = tmem_base + ((unsigned)((warp & 3) * 32) << 16);
        unsigned done_phase = 0, in_phase = 0, tile_par = 0;
        long long pc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tp = clock64();
        auto lap = [&](int i) {
            if (p.prof) {
                const long long now = clock64();
                pc[i] += now - tp;
                tp = now;
                if (tid == 0 && blockIdx.x == 0) p.prof[gridDim.x * 8 + i] = now;  // absolute time of the phase boundary (the last tile's survives)
            }
        };
        // A full tile's observations are three contiguous fp32 blocks in global memory: when they fit beside the
        // input operand (p.raw_off > 0), one thread pulls them into the idle K blocks of the A buffer with bulk async
        // copies, issued as soon as the previous tile's last MMA has read that memory, so the fetch hides behind the
        // layer-3 epilogue; the rows are then converted shared -> shared.  Otherwise: direct global loads.
        float *raw = reinterpret_cast<float *>(act + p.raw_off);
        const float *raw_own = raw, *raw_nbr = raw + TILE_M * p.d_own, *raw_grid = raw_nbr + TILE_M * p.d_nbr;
        auto bulk_tile = [&](int t) -> bool { return p.raw_off > 0 && (long long)(t + 1) * TILE_M <= p.n_rows; };
        auto fetch_tile = [&](int t) {
            if (tid == 0 && t < p.n_tiles && bulk_tile(t)) {
                const long long r0 = (long long)t * TILE_M;
                mbar_expect_tx(in_ready, (unsigned)(TILE_M * 4 * (p.d_own + p.d_nbr + p.d_grid)));
                bulk_g2s(raw, p.own + r0 * p.d_own, TILE_M * 4 * p.d_own, in_ready);
                bulk_g2s(raw + TILE_M * p.d_own, p.nbr + r0 * p.d_nbr, TILE_M * 4 * p.d_nbr, in_ready);
                bulk_g2s(raw + TILE_M * (p.d_own + p.d_nbr), p.grid + r0 * p.d_grid, TILE_M * 4 * p.d_grid, in_ready);
            }
        };
        // this warp's writes to the A buffer (generic proxy) and reads of the accumulators are done: tell the MMA thread
        auto publish = [&](int rdy) {
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(ready + rdy);
        };
        fetch_tile(blockIdx.x);
        for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
            const long long row = (long long)t * TILE_M + m;
            const bool valid = row < p.n_rows;
            // observations -> A operand (own | nbr | grid parts, each zero padded to a multiple of 16 columns, each
            // starting a K block); the two halves share the work
            if (bulk_tile(t)) {
                mbar_wait(in_ready, in_phase & 1u);
                ++in_phase;
                stage_row<true>(act, raw_own + m * p.d_own, raw_nbr + m * p.d_nbr, raw_grid + m * p.d_grid, p, part, m, true);
            } else stage_row<false>(act, p.own + row * p.d_own, p.nbr + row * p.d_nbr, p.grid + row * p.d_grid, p, part, m, valid);
            publish(RDY_STAGE);
            if (p.raw_off == 0) {  // pull the next tile's rows towards L2 while this tile computes
                const long long nrow = row + (long long)gridDim.x * TILE_M;
                if (nrow < p.n_rows) {
                    if (part == 0) prefetch_row(p.nbr + nrow * p.d_nbr, p.d_nbr);
                    else if (part == 1) { prefetch_row(p.own + nrow * p.d_own, p.d_own); prefetch_row(p.grid + nrow * p.d_grid, p.d_grid); }
                }
            }
            lap(0);
            // hidden epilogues, published piecewise so that the next layer's MMAs start on the finished K blocks while
            // the rest is still being converted: layer 1 by branch (own | nbr | grid), layer 2 by output half; in each
            // piece the NPART warp groups take an equal share of the columns
            for (int layer = 1; layer <= 2; ++layer) {
                constexpr int NB2 = H2 / 2 / NPART / 32;
                uint4 held[4 * NB2];
                float *dbg2 = (DBG && p.dbg_layer == 2 && valid) ? p.dbg + row * H2 : nullptr;
                if (layer == 2) {
                    // layer 2's first output half is finished while the MMAs of its second half still read the A buffer:
                    // convert it into registers now, store it when they are done
                    mbar_wait(half_done, tile_par);
                    tc_fence_after();
                    hidden_epilogue_hold<NB2, DBG>(tmem_row, part * (H2 / 2 / NPART), p.b2, held, dbg2);
                }
                mbar_wait(layer_done, done_phase & 1u);
                ++done_phase;
                tc_fence_after();
                lap(layer == 1 ? 1 : 3);
                const int ncols = layer == 1 ? H1C : H2;
                float *dbg_row = (DBG && p.dbg_layer == layer && valid) ? p.dbg + row * ncols : nullptr;
                if (layer == 1) {
#pragma unroll 1
                    for (int pc_ = 0; pc_ < 3; ++pc_) {  // 128 columns per branch
                        hidden_epilogue<H1 / NPART / 32, DBG>(act, tmem_row, L1_COL0, pc_ * H1 + part * (H1 / NPART), p.b1, m, dbg_row);
                        publish(RDY_E1 + pc_);
                    }
                } else {
                    hidden_epilogue_store<NB2>(act, part * (H2 / 2 / NPART), m, held);
                    publish(RDY_E2);
                    // second half in two pieces of 128 columns (two K blocks of layer 3 each), 32 columns per warp group
#pragma unroll 1
                    for (int u = 0; u < 2; ++u) {
                        hidden_epilogue<1, DBG>(act, tmem_row, 0, H2 / 2 + 128 * u + 32 * part, p.b2, m, dbg_row);
                        publish(RDY_E2 + 1 + u);
                    }
                }
                lap(layer == 1 ? 2 : 4);
            }
            // layer 3 epilogue fused with the output layer (256 -> 2), tanh, noise, clamp
            mbar_wait(layer_done, done_phase & 1u);
            ++done_phase;
            tc_fence_after();
            lap(5);
            fetch_tile(t + gridDim.x);  // the layer-3 MMAs were the last readers of the A buffer
            float a0 = 0.0f, a1 = 0.0f;
            float *dbg_row = (DBG && p.dbg_layer == 3 && valid) ? p.dbg + row * H3 : nullptr;
            {
                const int col0 = part * (H3 / NPART);
#pragma unroll
                for (int u = 0; u < H3 / NPART / 32; ++u) {
                    const int c = col0 + 32 * u;
                    unsigned x[32];
                    tmem_ld32_issue(tmem_row + c, x);
#pragma unroll
                    for (int hb = 0; hb < 2; ++hb) {  // 16 columns at a time: their bias / output weights are fetched first
                        float4 bq[4], w0[4], w1[4];
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            bq[g] = __ldg(reinterpret_cast<const float4 *>(p.b3 + c + 16 * hb + 4 * g));
                            w0[g] = __ldg(reinterpret_cast<const float4 *>(p.w4 + c + 16 * hb + 4 * g));
                            w1[g] = __ldg(reinterpret_cast<const float4 *>(p.w4 + H3 + c + 16 * hb + 4 * g));
                        }
                        if (hb == 0) tmem_ld_wait();
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            unsigned *y = x + 16 * hb + 4 * g;
                            bias_leaky2(y[0], y[1], bq[g].x, bq[g].y);
                            bias_leaky2(y[2], y[3], bq[g].z, bq[g].w);
                            const float h0 = __uint_as_float(y[0]), h1 = __uint_as_float(y[1]), h2 = __uint_as_float(y[2]), h3 = __uint_as_float(y[3]);
                            a0 = fmaf(h0, w0[g].x, a0); a0 = fmaf(h1, w0[g].y, a0); a0 = fmaf(h2, w0[g].z, a0); a0 = fmaf(h3, w0[g].w, a0);
                            a1 = fmaf(h0, w1[g].x, a1); a1 = fmaf(h1, w1[g].y, a1); a1 = fmaf(h2, w1[g].z, a1); a1 = fmaf(h3, w1[g].w, a1);
                            if (DBG && dbg_row) { float *o = dbg_row + c + 16 * hb + 4 * g; o[0] = h0; o[1] = h1; o[2] = h2; o[3] = h3; }
                        }
                    }
                }
            }
            tc_fence_before();  // the next tile's staging arrive tells the MMA thread the accumulators are drained
            // the column groups of a row meet through scratch lines in the (now idle) A buffer
            float2 *scratch = reinterpret_cast<float2 *>(act + SCRATCH_OFF);
            if (part > 0) scratch[(part - 1) * TILE_M + m] = make_float2(a0, a1);
            epi_bar();
            if (part == 0 && valid && p.actions) {
#pragma unroll
                for (int q = 0; q < NPART - 1; ++q) { const float2 o = scratch[q * TILE_M + m]; a0 += o.x; a1 += o.y; }
                a0 = tanhf(a0 + __ldg(p.b4));
                a1 = tanhf(a1 + __ldg(p.b4 + 1));
                if (p.noise_scale != 0.0f) {  // choose_action: act + randn(2) * var, clamp (V2/maddpg_agent:1290-1294)
                    const float2 z = normal_pair(p.noise_seed, (unsigned)row);
                    a0 = fminf(fmaxf(fmaf(p.noise_scale, z.x, a0), -1.0f), 1.0f);
                    a1 = fminf(fmaxf(fmaf(p.noise_scale, z.y, a1), -1.0f), 1.0f);
                }
                reinterpret_cast<float2 *>(p.actions)[row] = make_float2(a0, a1);
            }
            lap(6);
            tile_par ^= 1u;
        }
        if (p.prof && tid == 0)
            for (int i = 0; i < 8; ++i) p.prof[blockIdx.x * 8 + i] = pc[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PRODUCER_WARP) {
        __syncwarp();
        tmem_free(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------ host side

thread_local char g_err[512] = "";
int fail(int code, const char *fmt, const char *detail = "") {
    snprintf(g_err, sizeof(g_err), fmt, detail);
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e));
    return AAC_ACTOR_ERR_CUDA;
}

uint16_t bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return (uint16_t)((u >> 16) | 0x40u);
    return (uint16_t)((u + 0x7FFFu + ((u >> 16) & 1u)) >> 16);
}

int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

struct AacActor {
    AacActorConfig cfg;
    int device = 0, sms = 0;
    bool loaded = false;
    int64_t launches = 0;
    uint8_t *d_wpack = nullptr;
    Chunk *d_sched = nullptr;
    float *d_consts = nullptr;  // b1 [384] | b2 [512] | b3 [256] | w4 [512] | b4 [2]
    int n_chunks = 0, blk_nbr0 = 0, blk_grid0 = 0, blk_end = 0;
    long long *d_prof = nullptr;  // set by AAC_ACTOR_PROF=1 (tuning aid): per-CTA phase clocks
};

extern "C" const char *aac_actor_last_error(void) { return g_err; }

extern "C" int aac_actor_create(const AacActorConfig *cfg, AacActor **out) {
    if (!cfg || !out) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_create: null argument");
    if (cfg->abi_version != AAC_ACTOR_ABI_VERSION) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_create: abi_version mismatch");
    if (cfg->d_own < 1 || cfg->d_nbr < 1 || cfg->d_grid < 1 || cfg->d_own > 64 || cfg->d_nbr > 256 || cfg->d_grid > 256)
        return fail(AAC_ACTOR_ERR_ARG, "aac_actor_create: input widths out of range (own 1..64, nbr / grid 1..256)");
    if (cfg->max_rows < 1) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_create: max_rows < 1");
    AacActor *a = new (std::nothrow) AacActor();
    if (!a) return fail(AAC_ACTOR_ERR_STATE, "aac_actor_create: out of memory");
    a->cfg = *cfg;
    cudaError_t e = cudaGetDevice(&a->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&a->sms, cudaDevAttrMultiProcessorCount, a->device);
    int cc_major = 0, smem_max = 0;
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, a->device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, a->device);
    if (e != cudaSuccess) { delete a; return cuda_fail(e, "aac_actor_create: device query"); }
    if (cc_major != 10 || smem_max < SMEM_TOTAL) { delete a; return fail(AAC_ACTOR_ERR_STATE, "aac_actor_create: needs an sm_100 device (tcgen05, 227 KB shared memory)"); }
    e = cudaFuncSetAttribute(actor_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(actor_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL);
    if (e != cudaSuccess) { delete a; return cuda_fail(e, "aac_actor_create: cudaFuncSetAttribute"); }
    if (getenv("AAC_ACTOR_PROF")) {
        cudaMalloc(&a->d_prof, ((size_t)a->sms * 8 + 128) * sizeof(long long));
        cudaMemset(a->d_prof, 0, ((size_t)a->sms * 8 + 128) * sizeof(long long));
    }
    *out = a;
    return 0;
}

// tuning aid: copies the [sms][8] phase clocks (stage, wait L1, E1, wait L2, E2, wait L3, E3) to host memory
extern "C" int aac_actor_prof(AacActor *a, long long *host_out) {
    if (!a || !a->d_prof || !host_out) return fail(AAC_ACTOR_ERR_STATE, "aac_actor_prof: run with AAC_ACTOR_PROF=1");
    const cudaError_t e = cudaMemcpy(host_out, a->d_prof, ((size_t)a->sms * 8 + 128) * sizeof(long long), cudaMemcpyDeviceToHost);
    return e == cudaSuccess ? a->sms : cuda_fail(e, "aac_actor_prof");
}

extern "C" void aac_actor_destroy(AacActor *a) {
    if (!a) return;
    cudaFree(a->d_wpack);
    cudaFree(a->d_sched);
    cudaFree(a->d_consts);
    cudaFree(a->d_prof);
    delete a;
}

extern "C" int aac_actor_load(AacActor *a, const AacActorParams *hp) {
    if (!a || !hp) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_load: null argument");
    const float *ptrs[12] = {hp->w_own, hp->b_own, hp->w_nbr, hp->b_nbr, hp->w_grid, hp->b_grid, hp->w_merge, hp->b_merge, hp->w_hid, hp->b_hid, hp->w_out, hp->b_out};
    for (const float *q : ptrs)
        if (!q) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_load: null parameter pointer");
    std::vector<uint8_t> pack;
    std::vector<Chunk> sched;
    // B operand chunk: rows n0 .. n0 + nc of W [out, in], columns k0 .. k0 + kw (kw <= 64; zero beyond `in`), in the
    // K-major SWIZZLE_128B form: row n = 128 bytes, chunk (k / 8) stored at position (k / 8) ^ (n % 8)
    auto add_chunk = [&](const float *w, int in_dim, int n0, int nc, int k0, int kw, int a_blk, int col, unsigned flags) {
        Chunk ch;
        ch.src_off = (uint32_t)pack.size();
        ch.bytes = (uint32_t)(nc * 128);
        ch.n_mma = (uint32_t)(kw / 16);
        ch.idesc = umma_idesc((unsigned)nc);
        ch.wait_rdy = 0;
        ch.a_off16 = (uint32_t)(a_blk * BLK_BYTES) >> 4;
        ch.tmem_col = (uint32_t)col;
        ch.flags = flags;
        pack.resize(pack.size() + ch.bytes, 0);
        uint16_t *dst = reinterpret_cast<uint16_t *>(pack.data() + ch.src_off);
        for (int n = 0; n < nc; ++n)
            for (int kk = 0; kk < kw; ++kk) {
                const int k = k0 + kk;
                dst[n * 64 + (((kk >> 3) ^ (n & 7)) << 3) + (kk & 7)] = k < in_dim ? bf16_rne(w[(size_t)(n0 + n) * in_dim + k]) : (uint16_t)0;
            }
        sched.push_back(ch);
    };
    // layer 1: three branches into accumulator columns 0 / 128 / 256, each input part in its own K block(s)
    const int dims[3] = {a->cfg.d_own, a->cfg.d_nbr, a->cfg.d_grid};
    const float *w1[3] = {hp->w_own, hp->w_nbr, hp->w_grid};
    int blk = 0;
    for (int b = 0; b < 3; ++b) {
        const int kpad = round_up(dims[b], 16);
        if (b == 1) a->blk_nbr0 = blk;
        if (b == 2) a->blk_grid0 = blk;
        for (int k0 = 0; k0 < kpad; k0 += 64, ++blk) add_chunk(w1[b], dims[b], 0, H1, k0, kpad - k0 < 64 ? kpad - k0 : 64, blk, L1_COL0 + H1 * b, k0 == 0 ? CH_FRESH : 0u);
    }
    a->blk_end = blk;
    if (blk * BLK_BYTES > SCRATCH_OFF) return fail(AAC_ACTOR_ERR_STATE, "aac_actor_load: input operand does not fit");
    sched.front().wait_rdy = 1 + RDY_STAGE;
    sched.back().flags |= CH_LAYER_END;
    // layer 2: two halves of 256 output columns, K = 384 in blocks of 64
    // (its first half may start on the K blocks of a layer-1 branch as soon as that branch's epilogue is published)
    for (int h = 0; h < 2; ++h)
        for (int k0 = 0; k0 < H1C; k0 += 64) {
            add_chunk(hp->w_merge, H1C, 256 * h, 256, k0, 64, k0 / 64, 256 * h, k0 == 0 ? CH_FRESH : 0u);
            sched.back().wait_rdy = 1 + RDY_E1 + (h == 0 ? k0 / H1 : 2);
            if (h == 0 && k0 + 64 >= H1C) sched.back().flags |= CH_HALF_END;
        }
    sched.back().flags |= CH_LAYER_END;
    // layer 3: 256 output columns, K = 512
    for (int k0 = 0; k0 < H2; k0 += 64) {
        add_chunk(hp->w_hid, H2, 0, 256, k0, 64, k0 / 64, 0, k0 == 0 ? CH_FRESH : 0u);
        sched.back().wait_rdy = 1 + RDY_E2 + (k0 < 256 ? 0 : 1 + (k0 - 256) / 128);
    }
    sched.back().flags |= CH_LAYER_END;
    if ((int)sched.size() > MAX_CHUNKS) return fail(AAC_ACTOR_ERR_STATE, "aac_actor_load: chunk schedule too long");
    for (const Chunk &c : sched)
        if (c.bytes > (uint32_t)SLOT_BYTES || c.bytes % 16) return fail(AAC_ACTOR_ERR_STATE, "aac_actor_load: bad chunk size");

    std::vector<float> consts(H1C + H2 + H3 + 2 * H3 + 2);
    memcpy(consts.data(), hp->b_own, H1 * 4);
    memcpy(consts.data() + H1, hp->b_nbr, H1 * 4);
    memcpy(consts.data() + 2 * H1, hp->b_grid, H1 * 4);
    memcpy(consts.data() + H1C, hp->b_merge, H2 * 4);
    memcpy(consts.data() + H1C + H2, hp->b_hid, H3 * 4);
    memcpy(consts.data() + H1C + H2 + H3, hp->w_out, 2 * H3 * 4);
    memcpy(consts.data() + H1C + H2 + H3 + 2 * H3, hp->b_out, 2 * 4);

    cudaFree(a->d_wpack); cudaFree(a->d_sched); cudaFree(a->d_consts);
    a->d_wpack = nullptr; a->d_sched = nullptr; a->d_consts = nullptr;
    a->loaded = false;
    cudaError_t e = cudaMalloc(&a->d_wpack, pack.size());
    if (e == cudaSuccess) e = cudaMalloc(&a->d_sched, sched.size() * sizeof(Chunk));
    if (e == cudaSuccess) e = cudaMalloc(&a->d_consts, consts.size() * 4);
    if (e == cudaSuccess) e = cudaMemcpy(a->d_wpack, pack.data(), pack.size(), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(a->d_sched, sched.data(), sched.size() * sizeof(Chunk), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(a->d_consts, consts.data(), consts.size() * 4, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cuda_fail(e, "aac_actor_load: upload");
    a->n_chunks = (int)sched.size();
    a->loaded = true;
    return 0;
}

static int actor_launch(AacActor *a, const float *own, const float *nbr, const float *grid, int n_rows, float noise_scale, uint64_t seed,
                        float *actions, float *dbg, int dbg_layer, void *stream) {
    if (!a) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_forward: null handle");
    if (!a->loaded) return fail(AAC_ACTOR_ERR_STATE, "aac_actor_forward: no parameters loaded (aac_actor_load)");
    if (!own || !nbr || !grid) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_forward: null observation pointer");
    if (n_rows < 0 || n_rows > a->cfg.max_rows) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_forward: n_rows outside 0..max_rows");
    if (n_rows == 0) return 0;
    KArgs k;
    k.own = own; k.nbr = nbr; k.grid = grid; k.actions = actions; k.dbg = dbg; k.dbg_layer = dbg_layer;
    k.n_rows = n_rows; k.n_tiles = (n_rows + TILE_M - 1) / TILE_M;
    k.d_own = a->cfg.d_own; k.d_nbr = a->cfg.d_nbr; k.d_grid = a->cfg.d_grid;
    k.blk_nbr0 = a->blk_nbr0; k.blk_grid0 = a->blk_grid0; k.blk_end = a->blk_end;
    k.wpack = a->d_wpack; k.sched = a->d_sched; k.n_chunks = a->n_chunks;
    k.b1 = a->d_consts; k.b2 = k.b1 + H1C; k.b3 = k.b2 + H2; k.w4 = k.b3 + H3; k.b4 = k.w4 + 2 * H3;
    k.noise_scale = noise_scale; k.noise_seed = seed;
    k.prof = a->d_prof;
    // bulk fetch needs 16-byte aligned blocks and room for the raw tile between the input operand and the scratch line
    const int raw_bytes = TILE_M * 4 * (k.d_own + k.d_nbr + k.d_grid);
    const bool aligned = ((uintptr_t)own % 16 == 0) && ((uintptr_t)nbr % 16 == 0) && ((uintptr_t)grid % 16 == 0);
    k.raw_off = (aligned && a->blk_end * BLK_BYTES + raw_bytes <= SCRATCH_OFF) ? a->blk_end * BLK_BYTES : 0;
    const int grid_dim = k.n_tiles < a->sms ? k.n_tiles : a->sms;
    if (dbg) actor_kernel<true><<<grid_dim, THREADS, SMEM_TOTAL, (cudaStream_t)stream>>>(k);
    else actor_kernel<false><<<grid_dim, THREADS, SMEM_TOTAL, (cudaStream_t)stream>>>(k);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "aac_actor_forward: launch");
    ++a->launches;
    return 0;
}

extern "C" int aac_actor_forward(AacActor *a, const float *own, const float *nbr, const float *grid, int32_t n_rows, float noise_scale,
                                 uint64_t noise_seed, float *actions, void *stream) {
    if (!actions) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_forward: null actions pointer");
    return actor_launch(a, own, nbr, grid, n_rows, noise_scale, noise_seed, actions, nullptr, 0, stream);
}

extern "C" int aac_actor_hidden(AacActor *a, const float *own, const float *nbr, const float *grid, int32_t n_rows, int32_t layer, float *hidden,
                                void *stream) {
    if (!hidden || layer < 1 || layer > 3) return fail(AAC_ACTOR_ERR_ARG, "aac_actor_hidden: layer must be 1..3 and hidden non-null");
    return actor_launch(a, own, nbr, grid, n_rows, 0.0f, 0, nullptr, hidden, layer, stream);
}

extern "C" int64_t aac_actor_launch_count(const AacActor *a) { return a ? a->launches : 0; }
