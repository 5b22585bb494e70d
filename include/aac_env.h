/*
 * aac_env.h -- C ABI of the B200 batched multi-drone environment step.
 *
 * Drop-in boundary.  The reference (zhangmingcheng28/Multi_agent_AAC) has no FFI: its hot path is
 * the Python method surface of `env_simulator`
 *   ATT = MADDPG_ownENV_randomOD_radar_one_model_att/env_simulator_randomOD_radar_sur_drones_oneModel_att.py
 *   V2  = MADDPG_ownENV_randomOD_radar_N_model_use_tdCPA_forV2/env_simulator_randomOD_radar_sur_drones_N_Model_use_tdCPA_forV2.py
 *   MM  = MADDPG_ownENV_randomOD_radar_multipleMap/env_simulator_randomOD_radar_multipleMap.py
 * Each entry point below names the reference method it replaces.  multi_agent_aac_b200/ref_compat.py
 * re-exposes those methods with the reference's signatures on top of this ABI; INTEGRATION.md shows
 * the ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer inside AacState / AacOut / actions is a DEVICE pointer to caller-owned memory
 *     (torch tensors in the shipped host code), contiguous, layouts as documented per field;
 *   - map tables, ray tables and scenario banks are HOST pointers, copied at the call;
 *   - calls enqueue work on the given CUDA stream and return without synchronising;
 *   - return value 0 = ok, negative = AAC_ERR_*; aac_last_error() gives the message (thread local);
 *   - one AacEnv per device shard; a handle is not thread-safe, different handles are independent;
 *   - all calls on ONE handle must be stream-ordered: issue them on one stream (or order the streams yourself): the
 *     persistent kernels of a handle share its group counters, two of its launches must never run concurrently;
 *   - positions in AacState are LOCAL coordinates: global metres minus the map's bound centre
 *     (AacMapDesc.origin_*), float32.  Observations are emitted in the reference's global frame.
 */
#ifndef AAC_ENV_H
#define AAC_ENV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AAC_ABI_VERSION 3

/* variants */
#define AAC_VARIANT_ATT 0 /* one_model_att: radar senses other drones' 64-gons, summed reward      */
#define AAC_VARIANT_V2 1  /* tdCPA_forV2: radar senses grid + bounds, distance-sorted neighbours */
#define AAC_VARIANT_MM 2  /* radar_multipleMap: per-env maps, true-min radar                     */

/* V2 radar value stored in the observation: the reference keeps the LAST hit in
 * STRtree query order (V2:1265-1288); this build defines that order as ascending cell index,
 * then the four boundary lines L,R,B,T.  AAC_RADAR_MIN stores the true minimum instead. */
#define AAC_RADAR_MIN 0
#define AAC_RADAR_LAST_HIT 1

/* optional outputs: bit set => the matching AacOut pointer must be non-NULL */
#define AAC_OUT_RAW 0x01       /* raw_own / raw_nbr / raw_nbr6 (the reference's un-normalised state) */
#define AAC_OUT_NBR6 0x02      /* norm_nbr6 legacy 6-vector block (ATT actor input obs_nei)          */
#define AAC_OUT_TCPA_PAIR 0x04 /* tcpa_pair, nbr_order                                               */
#define AAC_OUT_RADAR_AUX 0x08 /* radar_min, radar_hit                                               */
#define AAC_OUT_PARTS 0x10     /* reward parts, branch                                               */

/* Radar target classes of the later fork (CS = MADDPG_ownENV_randomOD_radar_N_model_use_tdCPA_forV2_changeskin/
 * env_simulator_..._changeskin.py:1379-1506).  AacConfig.radar_targets = 0 keeps the variant's own
 * radar; any other value (tdCPA_forV2 only) switches to the fork's TRUE-MINIMUM radar over the classes named: */
#define AAC_TARGET_CELLS 0x1    /* occupied grid cells (the older variants' buildings)                               */
#define AAC_TARGET_BOUNDS 0x2   /* the four boundary SEGMENTS of the bound rectangle            (CS:676-680, :1418-1428) */
#define AAC_TARGET_CLOUDS 0x4   /* moving clouds: outline of Point(pos).buffer(radius)          (CS:1430-1456)        */
#define AAC_TARGET_AIRCRAFT 0x8 /* the other drones' protective outlines (include_other_AC)     (CS:1458-1490)        */
#define AAC_MAX_CLOUDS 8

#define AAC_ERR_ARG -1
#define AAC_ERR_CUDA -2
#define AAC_ERR_STATE -3

#define AAC_MAX_AGENTS 32
#define AAC_MAX_RAYS 128
#define AAC_MAX_W 64
#define AAC_MAP_STRIDE 1024 /* bytes reserved per map in the occupancy table (gx*gy <= 1024) */

typedef struct AacEnv AacEnv;

typedef struct {
    int32_t abi_version;    /* AAC_ABI_VERSION */
    int32_t variant;        /* AAC_VARIANT_* */
    int32_t n_envs;         /* E on this device */
    int32_t n_agents;       /* N, 1..AAC_MAX_AGENTS           (ATT/ma_main:100) */
    int32_t n_rays;         /* R, divides 360                 (ATT:1058 range(0,360,20) => 18) */
    int32_t w_max;          /* ref-line vertex stride, multiple of 8, <= AAC_MAX_W */
    int32_t radar_mode;     /* AAC_RADAR_* (V2 only) */
    int32_t sum_reward;     /* full_observable_critic_flag    (ATT:2602; ATT/ma_main:77) */
    int32_t episode_length; /* step cap                       (ATT/ma_main:914 => 50, V2: 100) */
    int32_t out_flags;      /* AAC_OUT_* */
    int32_t tile_envs;      /* envs per warp (tile_envs * n_agents <= 32); 0 = pick 32 / n_agents */
    int32_t block_threads;  /* threads per CTA; 0 = pick */
    int64_t env_id_base;    /* global id of env 0 of this shard (scenario hashing) */
    uint64_t seed;
    float dt;               /* 0.5   (ATT:60)  */
    float vmax;             /* 5     (ATT/ma_main:150) */
    float acc_max;          /* 8     (ATT/ma_main:136); MM integrates with 20 (MM:2025) */
    float prot;             /* 2.5   (ATT/agent:43) */
    float ray_len;          /* 15    (ATT:1066 detectionRange/2) */
    float goal_r;           /* 1     (ATT:2266) */
    int32_t eval_by_step;   /* V2 only: args.mode == 'eval' with evaluation_by_episode == False: crashed / arrived drones stay
                               where they are, crashes do not end the episode (V2:3729-3734, :3128-3156, :3551-3587) */
    int32_t autoreset_launches; /* aac_step_autoreset / aac_step_host: 1 = one fused launch (each group is stepped and its finished
                               envs re-initialised at once), 2 = step launch + reset launch, 3 = one phased launch (every group
                               is stepped, then the finished envs are re-initialised by whichever warp is free: the benchmark
                               shapes with out_flags == 0 only, else as 2), 0 = choose by batch size */
    /* the later fork's sensor classes (tdCPA_forV2 only; all zero = off) */
    int32_t radar_targets;  /* AAC_TARGET_* */
    int32_t n_nbr_obs;      /* > 0: only the nearest n neighbours enter norm_nbr / raw_nbr, rows of 5 * n floats
                               (use_nearestN_neigh_wRadar / N_neigh, CS:1799-1802) */
    int32_t n_clouds;       /* <= AAC_MAX_CLOUDS */
    float clouds[AAC_MAX_CLOUDS][6]; /* start x, y, goal x, y (global metres), radius, speed: a cloud starts every episode at its start
                               and moves speed * dt towards its goal per step until closer than 1 m (cloud.py; CS:598-664, :4667-4681;
                               calculate_next_position) */
} AacConfig;

/* one 10 m occupancy grid (ATT/grid_env_generation:140-185) */
typedef struct {
    int32_t gx, gy;          /* cells; occupancy is uint8[gx*gy], ix-major */
    float bound[4];          /* xmin xmax ymin ymax, global metres (ATT/parameters:32-36) */
    float x0c, y0c;          /* global centre of cell (0,0) */
    float cell;              /* 10 */
    float origin_x, origin_y; /* local-frame origin = bound centre */
} AacMapDesc;

/* per-agent state, SoA, index a = env*N + agent.  Mirrors the `Agent` record (ATT/agent:14-53). */
typedef struct {
    float *px, *py;       /* [E*N] local position              (agent.pos) */
    float *vx, *vy;       /* [E*N]                             (agent.vel) */
    float *heading;       /* [E*N] rad                         (agent.heading) */
    uint32_t *meta;       /* [E*N] b0-7 waypoints popped, b8 reach_target, b9 bound_collision,
                             b10 building_collision, b11 drone_collision, b16-23 / b24-31 first two
                             keys of pre_surroundingNeighbor (0xFF none) */
    uint16_t *ref_cells;  /* [E*N*w_max] ref_line vertices as ix<<8|iy (vertex 0 = ini_pos) */
    uint8_t *ref_w;       /* [E*N] vertex count */
    int32_t *wall_count;  /* [E*N] collide_wall_count, may be NULL */
    int32_t *ep_step;     /* [E] steps taken in the running episode */
    int32_t *ep_index;    /* [E] episodes finished */
    float *ep_return;     /* [E] sum of all drones' rewards in the running episode */
    int32_t *map_id;      /* [E] row of the map table (required for AAC_VARIANT_MM, else may be NULL => 0) */
    uint32_t *wp_mask;    /* [E*N] AAC_VARIANT_MM: bit k = ref-line vertex k is still in agent.goal (MM:1747-1762) */
} AacState;

/* outputs of one step / reset; D_own = 6+4(N-1) ATT, 7 V2, 6 MM.  AAC_VARIANT_MM emits norm_own, radar (and raw_own):
 * its legacy neighbour block is ragged and not read by its actors (MM:754-770, MM/maddpg_agent:361-399) */
typedef struct {
    float *norm_own;      /* [E,N,D_own]     norm state p1          (ATT:1463-1479, V2:1672-1694) */
    float *norm_nbr;      /* [E,N,5(N-1)]    V2 norm p2, else NULL  (V2:1570-1571,1697) */
    float *radar;         /* [E,N,R] metres  p2 / p2_radar          (ATT:1170, V2:1300) */
    float *norm_nbr6;     /* [E,N,N-1,6]     norm p3   AAC_OUT_NBR6 (ATT:1402-1410) */
    float *raw_own;       /* [E,N,D_own]     AAC_OUT_RAW */
    float *raw_nbr;       /* [E,N,5(N-1)]    AAC_OUT_RAW, V2 */
    float *raw_nbr6;      /* [E,N,N-1,6]     AAC_OUT_RAW|AAC_OUT_NBR6 */
    float *reward;        /* [E,N]           (ATT:2618) */
    uint8_t *done;        /* [E,N] */
    uint8_t *check_goal;  /* [E,N] */
    uint8_t *bbc;         /* [E,4]  bound_building_check */
    uint8_t *terminated;  /* [E] bit0 step cap, bit1 any done, bit2 all reached (ATT/ma_main:448-462) */
    float *tcpa_min;      /* [E,N,4] immediate tcpa (+inf none), its d_tcpa, neighbour key (-1 none),
                             conflict counters cur+256*pre   (ATT:2189-2209, V2:3094-3117) */
    float *tcpa_pair;     /* [E,N,N-1,4] tcpa,d,pre_tcpa,pre_d in neighbour order  AAC_OUT_TCPA_PAIR */
    int8_t *nbr_order;    /* [E,N,N-1]                                             AAC_OUT_TCPA_PAIR */
    float *radar_min;     /* [E,N,R]                                               AAC_OUT_RADAR_AUX */
    int16_t *radar_hit;   /* [E,N,R] cell ix*gy+iy | gx*gy+{0..3} bound | gx*gy+4+j drone | gx*gy+4+N+c cloud | -1 */
    float *parts;         /* [E,N,8] dist_to_goal, near_drone, near_bldg, small_step, cross_err,
                             after_dist_hg, min_radar, nearest_dist                AAC_OUT_PARTS */
    int8_t *branch;       /* [E,N] 0 bound 1 building 2 drone 3 goal 4 normal      AAC_OUT_PARTS */
    uint8_t *cloud_contact; /* [E,N] the drone's protective circle overlaps a cloud (CS:4099-4110); may be NULL */
} AacOut;

/* pre-planned episodes the device resets from (host arrays; see reset.py ScenarioBank) */
typedef struct {
    int32_t n_scenarios;
    const uint16_t *cells; /* [S,N,w_max] */
    const uint8_t *w;      /* [S,N] */
    const int32_t *map_id; /* [S] or NULL (all on map 0) */
} AacBank;

/* Origin / destination table of one map: every free cell that can be a start or a goal (the four quadrant pools of
 * create_world, ATT:154-197) and the pruned grid path between every pair of them (reset_world's jps_find_path +
 * collinear pruning, ATT:317-347), planned once on the host.  With a table installed the device draws origins and
 * destinations itself, following reset_world's rule (ATT:254-276): a start quadrant, a different target quadrant,
 * a start cell redrawn until it is more than 2 * protectiveBound from every earlier drone's start, a goal cell. */
typedef struct {
    int32_t n_cells;            /* P: pool cells, sorted by quadrant */
    int32_t pool_off[5];        /* quadrant q owns cells [pool_off[q], pool_off[q+1]) */
    const uint16_t *cell_code;  /* [P] ix<<8|iy */
    const uint32_t *path_off;   /* [P*P] offset of path (start, goal) in path_cells, a multiple of 8 (paths are padded to
                                   8-cell = 16-byte chunks); pairs inside one quadrant unused */
    const uint8_t *path_len;    /* [P*P] vertices of that path (2..w_max), 0 = unused pair */
    const uint16_t *path_cells; /* vertex pool */
    int64_t n_path_cells;
} AacOdTable;

/* env_simulator.__init__ + create_world (ATT:41,84) */
int aac_create(const AacConfig *cfg, AacEnv **out);
void aac_destroy(AacEnv *env);
/* world_map / bound / allGridPoly constructor arguments (ATT:41; MM:42 takes collections) */
int aac_set_maps(AacEnv *env, const AacMapDesc *maps, const uint8_t *occ /* [M,AAC_MAP_STRIDE] */, int32_t n_maps);

/* Optional (V2 / multipleMap): the radar of a drone standing on a cell centre, for every cell of every map - device arrays
 * [n_maps][AAC_MAP_STRIDE][n_rays], cell index ix * gy + iy: `radar` as aac_observe writes it for this handle's radar_mode,
 * `radar_min` / `radar_hit` likewise (required with AAC_OUT_RADAR_AUX), `min_bits` [n_maps][AAC_MAP_STRIDE] = the smallest
 * IEEE bit pattern of a cell's `radar` row.  reset_world puts every drone on a cell centre (ATT:301-372), so the observation
 * of a freshly reset env reads its ranges here instead of casting the rays again; build it with THIS library (observe on
 * drones placed at the cell centres, as `BatchedDroneEnv.build_radar_table` does) so that the values are the ones the
 * kernel itself computes.  The arrays stay caller-owned; all NULL removes the table; aac_set_maps removes it too. */
int aac_set_radar_table(AacEnv *env, const float *radar, const float *radar_min, const int16_t *radar_hit, const uint32_t *min_bits);
int aac_set_bank(AacEnv *env, const AacBank *bank);
/* one table per map (host pointers, copied); resets draw from the tables instead of the scenario bank.
 * A table with path_off = path_len = path_cells = NULL carries the pools only: the reference line of every episode is then
 * searched when the episode starts, on the device, by the warp that re-initialises the env - reset_world's per-episode
 * jps_find_path + pruning (ATT:317-331), the same planner as aac_plan_path, so the episodes are those of a table with paths,
 * bit for bit, without its P^2 paths.  An unreachable goal or a line of more than w_max vertices cannot be reported from
 * there: the line falls back to start -> goal and statistic [10] counts it.  Not with the sensor / evaluation
 * configurations (radar_targets, n_nbr_obs, eval_by_step), which keep tables with paths. */
int aac_set_od_tables(AacEnv *env, const AacOdTable *tables, int32_t n_maps);
/* Host-only helper: the reference's grid search (ATT/jps_straight.py:17-70: best-first on f = g + Manhattan, first
 * minimum in discovery order, neighbours visited in the order (0,-1), (0,1), (-1,0), (1,0), cells never re-opened)
 * followed by collinear pruning (ATT:321-331).  occ is uint8[gx*gy], ix-major.  Writes up to max_cells codes
 * ix<<8|iy and returns the vertex count, 0 if the goal is unreachable, -1 if the path needs more than max_cells. */
int aac_plan_path(const uint8_t *occ, int32_t gx, int32_t gy, int32_t sx, int32_t sy, int32_t tx, int32_t ty, uint16_t *out_cells,
                  int32_t max_cells);
/* The same search and pruning for many origin / destination pairs at once ON THE DEVICE (one warp per pair; used to
 * build a map's origin / destination table).  Host buffers: occ uint8[gx*gy] ix-major; pairs
 * uint16[n_pairs][2] = (start, goal) cell codes ix<<8|iy; out_cells uint16[n_pairs][max_cells] (zero padded);
 * out_len int32[n_pairs] with aac_plan_path's return convention per pair (vertex count, 0 unreachable, -1 more than
 * max_cells vertices).  Results equal aac_plan_path's exactly.  Synchronises the stream. */
int aac_plan_paths_device(const uint8_t *occ, int32_t gx, int32_t gy, const uint16_t *pairs, int64_t n_pairs, uint16_t *out_cells,
                          int32_t *out_len, int32_t max_cells, void *cuda_stream);
int aac_bind_state(AacEnv *env, const AacState *state);
/* reset_world (ATT:199-511): re-initialise the envs whose mask byte is non-zero (NULL = all) from
 * the scenario bank and emit their first observation */
int aac_reset(AacEnv *env, const uint8_t *mask_dev, const AacOut *out, void *cuda_stream);
/* cur_state_norm_state_v3 on the bound state as it stands, pre == cur (ATT:404-405, :837) */
int aac_observe(AacEnv *env, const AacOut *out, void *cuda_stream);
/* env.step + ss_reward / ss_reward_Mar (ATT:2627 + :2105, V2:3703 + :2995, MM:2016 + :1674) */
int aac_step(AacEnv *env, const float *actions_dev /* [E,N,2] */, const AacOut *out, void *cuda_stream);
/* aac_step followed by the caller's episode rule (ATT/ma_main:448-462 -> reset_world) in one call: the
 * envs that terminate in this step keep their terminal reward / done / check_goal / bbc / terminated
 * and are re-initialised (scenario bank or origin / destination tables); their observation rows carry the
 * reset observation.  Two launches on the stream (step, then reset of the terminated envs): measured faster
 * than the single fused launch below, whose results are bit-identical */
int aac_step_autoreset(AacEnv *env, const float *actions_dev /* [E,N,2] */, const AacOut *out, void *cuda_stream);
/* the same in ONE launch (the step kernel re-initialises and re-observes the envs it terminates) */
int aac_step_fused(AacEnv *env, const float *actions_dev /* [E,N,2] */, const AacOut *out, void *cuda_stream);
/* the caller's episode rule (ATT/ma_main:448-462 -> reset_world): reset every env whose
 * out->terminated byte is non-zero, overwrite its observation rows with the reset observation */
int aac_autoreset(AacEnv *env, const AacOut *out, void *cuda_stream);
/* env.step + ss_reward with HOST buffers: copies `actions_host` to the device, steps, and copies every
 * non-NULL field of `out_host` back (same layouts as AacOut, host memory, ideally pinned); the
 * device-side AacOut bound by the previous aac_step/aac_observe/aac_reset call is the staging area.
 * Synchronises the stream before returning.  This is the call the e2e benchmark times. */
int aac_step_host(AacEnv *env, const float *actions_host, const AacOut *out_dev, const AacOut *out_host,
                  int32_t autoreset, void *cuda_stream);
/* episode statistics accumulated on the device since the last call with reset != 0 (mirrors the
 * per-100-episode counters of ATT/ma_main:581-637): [0] episodes, [1] steps, [2] sum of returns,
 * [3] bound crash episodes, [4] building, [5] drone, [6] drone-crash-with-nearest, [7] episodes in
 * which every drone reached its goal, [8] drones that reached, [9] step-cap endings, [10] reference lines that fell
 * back to start -> goal in the per-episode search (pools-only tables), [11..15] 0.
 * Synchronises the stream. */
#define AAC_N_STATS 16
int aac_read_stats(AacEnv *env, double *stats_host /* [AAC_N_STATS] */, int32_t reset, void *cuda_stream);
/* device kernels launched by this handle so far */
int64_t aac_launch_count(const AacEnv *env);
int aac_own_dim(int32_t variant, int32_t n_agents);
const char *aac_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* AAC_ENV_H */
