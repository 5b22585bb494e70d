/* C ABI of the batched actor (policy) forward pass: SURVEY.md 8f rank 2, the caller side of the env step.
 *
 * Replaces, for a whole batch of drones at once, what the reference does one drone at a time in
 *   maddpg_agent.choose_action            V2/maddpg_agent:1241-1310   (N sequential batch-1 forwards per env step)
 *   ActorNetwork_allnei_wRadar.forward    V2/Nnetworks:273-340        (the network those forwards run)
 * for the network the tdCPA_forV2 configuration trains (use_allNeigh_wRadar = True, one shared model,
 * V2/ma_main:81,95,100):
 *
 *   own  [d_own ] -> Linear(128) -> LeakyReLU(0.01) \
 *   nbr  [d_nbr ] -> Linear(128) -> LeakyReLU(0.01)  > concat [384] -> Linear(512) -> LeakyReLU(0.01)
 *   grid [d_grid] -> Linear(128) -> LeakyReLU(0.01) /           -> Linear(256) -> LeakyReLU(0.01) -> Linear(2) -> tanh
 *
 * The three inputs are exactly the env step's outputs norm_own / norm_nbr / radar (include/aac_env.h), read
 * where the env kernel left them in device memory; the actions are written in the layout aac_step takes.
 * Arithmetic: bf16 operands on the tensor cores (tcgen05, accumulators in tensor memory), fp32 accumulation,
 * bias / activation / last layer / tanh in fp32.  All pointers below are DEVICE pointers unless named host_*.
 * Calls enqueue on the given stream and return 0 or a negative AAC_ACTOR_ERR_*; aac_actor_last_error() has
 * the message.  No global state; one handle per device; one host thread per handle.
 */
#ifndef AAC_ACTOR_H
#define AAC_ACTOR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AAC_ACTOR_ABI_VERSION 1
#define AAC_ACTOR_ERR_ARG (-1)
#define AAC_ACTOR_ERR_CUDA (-2)
#define AAC_ACTOR_ERR_STATE (-3)

#define AAC_ACTOR_H1 128 /* width of each input branch       (V2/Nnetworks:292-294) */
#define AAC_ACTOR_H2 512 /* merge_feature                    (V2/Nnetworks:295)     */
#define AAC_ACTOR_H3 256 /* act_out hidden                   (V2/Nnetworks:297)     */
#define AAC_ACTOR_NACT 2 /* n_actions                        (V2/ma_main: dim_act)  */

typedef struct AacActorConfig {
    int32_t abi_version; /* AAC_ACTOR_ABI_VERSION */
    int32_t d_own;       /* actor_dim[0]: 7 for tdCPA_forV2                       (V2/ma_main:132) */
    int32_t d_nbr;       /* actor_dim[1]: 5 * (n_agents - 1)                                        */
    int32_t d_grid;      /* actor_dim[2]: number of radar rays                                      */
    int32_t max_rows;    /* largest batch (n_envs * n_agents) a forward call will see               */
} AacActorConfig;

/* Parameters in torch.nn.Linear layout (weight [out, in] row-major, bias [out]), float32, HOST memory:
 * own_fc.0, own_full_nei.0, own_grid.0, merge_feature.0, act_out.0, act_out.2 (V2/Nnetworks:292-298). */
typedef struct AacActorParams {
    const float *w_own, *b_own;   /* [128, d_own],  [128] */
    const float *w_nbr, *b_nbr;   /* [128, d_nbr],  [128] */
    const float *w_grid, *b_grid; /* [128, d_grid], [128] */
    const float *w_merge, *b_merge; /* [512, 384],  [512] */
    const float *w_hid, *b_hid;   /* [256, 512],    [256] */
    const float *w_out, *b_out;   /* [2, 256],      [2]   */
} AacActorParams;

typedef struct AacActor AacActor;

int aac_actor_create(const AacActorConfig *cfg, AacActor **out);
void aac_actor_destroy(AacActor *actor);

/* Rounds the weights to bf16, tiles them into the tensor-core operand layout and uploads them (synchronous). */
int aac_actor_load(AacActor *actor, const AacActorParams *host_params);

/* actions[r, 0..1] = clamp(actor(own[r], nbr[r], grid[r]) + noise_scale * n(0, 1), -1, 1) for r < n_rows
 * (choose_action, V2/maddpg_agent:1284-1296; noise_scale = the reference's self.var[i], 0 = noisy=False).
 * The normal draws are counter-based on (noise_seed, r): repeatable, independent of launch shape.
 * own / nbr / grid / actions: float32, row-major, contiguous [n_rows, d]. */
int aac_actor_forward(AacActor *actor, const float *own, const float *nbr, const float *grid, int32_t n_rows, float noise_scale,
                      uint64_t noise_seed, float *actions, void *stream);

/* Debug / parity aid: post-activation output of hidden layer `layer` (1: [n_rows, 384], 2: [n_rows, 512],
 * 3: [n_rows, 256]) as float32, for the rows of the same inputs. */
int aac_actor_hidden(AacActor *actor, const float *own, const float *nbr, const float *grid, int32_t n_rows, int32_t layer, float *hidden,
                     void *stream);

/* Tuning aid: with AAC_ACTOR_PROF=1 in the environment at aac_actor_create, the kernel accumulates per-CTA phase clocks
 * (staging, wait / epilogue of layers 1-3); copies [n_ctas][8] + 128 int64 to host_out and returns n_ctas. */
int aac_actor_prof(AacActor *actor, long long *host_out);

int64_t aac_actor_launch_count(const AacActor *actor); /* kernels launched by this handle so far */
const char *aac_actor_last_error(void);

/* ---------------------------------------------------------------------------------------------------------------------
 * The attention actor of the one_model_att variant: ActorNetwork_ATT_TwoPortion (ATT/Nnetworks:177-213), called one drone
 * at a time by choose_action (ATT/maddpg_agent:455-503) on [obs, obs_grid, obs_nei] = the env step's norm_own / radar /
 * norm_nbr6 outputs:
 *   own [d_own] -> Linear(64) + ReLU = own_obs;  grid [d_grid] -> Linear(64) + ReLU;  nei [n_nei, d_nei] -> Linear(64) + ReLU = x_e
 *   score_m = k(x_e[m]) . q(own_obs) / 8, softmax over the neighbours whose row does not average to zero, v_att = sum alpha_m v(x_e[m])
 *   concat [192] -> Linear(256) + ReLU -> Linear(2) -> tanh
 * fp32 on the CUDA cores (66 k multiply-adds per drone; this variant's batches are small), same calling conventions as
 * above; actions get the same exploration noise + clamp (ATT/maddpg_agent:497-501). */
typedef struct AacActorAttConfig {
    int32_t abi_version; /* AAC_ACTOR_ABI_VERSION */
    int32_t d_own;       /* 6 + 4 (n_agents - 1): what the env emits */
    int32_t d_grid;      /* radar rays */
    int32_t d_nei;       /* 6 */
    int32_t n_nei;       /* n_agents - 1 neighbour rows per drone */
} AacActorAttConfig;

/* torch layout ([out, in] row-major), float32, HOST memory: own_fc.0, own_grid.0, neigh_fc.0, q, k, v (no bias),
 * merge_feature.0, act_out.0 (ATT/Nnetworks:181-190) */
typedef struct AacActorAttParams {
    const float *w_own, *b_own;     /* [64, d_own],  [64] */
    const float *w_grid, *b_grid;   /* [64, d_grid], [64] */
    const float *w_nei, *b_nei;     /* [64, d_nei],  [64] */
    const float *w_q, *w_k, *w_v;   /* [64, 64] each */
    const float *w_merge, *b_merge; /* [256, 192],   [256] */
    const float *w_out, *b_out;     /* [2, 256],     [2] */
} AacActorAttParams;

typedef struct AacActorAtt AacActorAtt;

int aac_actor_att_create(const AacActorAttConfig *cfg, AacActorAtt **out);
void aac_actor_att_destroy(AacActorAtt *actor);
int aac_actor_att_load(AacActorAtt *actor, const AacActorAttParams *host_params);
/* own [n_rows, d_own], grid [n_rows, d_grid], nei [n_rows, n_nei, d_nei], actions [n_rows, 2]: device pointers */
int aac_actor_att_forward(AacActorAtt *actor, const float *own, const float *grid, const float *nei, int32_t n_rows, float noise_scale,
                          uint64_t noise_seed, float *actions, void *stream);
int64_t aac_actor_att_launch_count(const AacActorAtt *actor);
const char *aac_actor_att_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
