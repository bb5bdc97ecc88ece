/*
 * crowdnav.h -- C ABI of libcrowdnav.so, the B200 batched crowd-navigation
 * environment step.
 *
 * The reference (ailabspace/drl-based-mapless-crowd-navigation-with-perceived-risk)
 * has no FFI of its own: its boundary is the Python duck-type `Env`
 * (turtlebot3_rl_sim/src/environment_stage_1_nobonus.py:42-43,1164,1227,
 * 1265-1283) called by the training drivers (start_td3_training.py:106-148).
 * This header is the interface a maintainer binds underneath that class
 * (ctypes stub in INTEGRATION.md).  Each entry point cites what it replaces.
 *
 * Conventions
 *   - plain C, no torch / C++ types; every pointer suffixed _dev is a CUDA
 *     device pointer on the handle's device, _host is host memory.
 *   - all work is enqueued on the caller's stream (a cudaStream_t passed as
 *     void*); no hidden synchronisation, graph-capturable.
 *   - return 0 on success, a negative cn_status otherwise; message via
 *     cn_last_error().  No C++ exception crosses the boundary.
 *   - a handle is bound to one device: every entry point makes that device
 *     current for the duration of the call and puts the caller's device back,
 *     so one process may drive handles on several GPUs.  Calls on one handle
 *     are not re-entrant.
 */
#ifndef CROWDNAV_H
#define CROWDNAV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CN_MAX_PEDS       64
#define CN_MAX_BEHAVIORS  8
#define CN_ABI_VERSION    2

/* words (4 B) per env in the three state planes */
#define CN_ROBOT_WORDS    16
#define CN_PED_WORDS      8   /* 4 in plane A (x, y, vx, vy) + 4 in plane B */

typedef enum cn_status {
    CN_OK = 0,
    CN_ERR_INVALID = -1,     /* bad argument / config */
    CN_ERR_CUDA = -2,        /* a CUDA runtime call failed */
    CN_ERR_NOMEM = -3,
    CN_ERR_UNSUPPORTED = -4  /* config outside compiled limits */
} cn_status;

/* cn_config.flags */
#define CN_FLAG_AUTO_RESET      1u  /* a world that ended on step t restarts DURING step t+1: that step ignores
                                       its action and returns the new episode's first observation, reward 0,
                                       done = 2 ("next-step" auto-reset: every world does one get_state per launch) */
#define CN_FLAG_TOPK_HIGHEST    2u  /* keep the K highest-CP objects instead of the
                                       reference's `[-K:]` (= K lowest), ENV:883 */

#define CN_FLAG_ENV_ORIGINAL    4u  /* the reference's ORIGINAL environment (environment_stage_1_original.py, used by its
                                       DQN / tabular drivers and the 363-wide `trajectory_test` checkpoints): row =
                                       [R-1 ranges | heading, distance to the GOAL | x, y], no waypoints, no K block
                                       (k_obstacles must be 0), reward = progress terms + terminal, computed -- like
                                       the reference does (original:324-326) -- from the row's last two entries.
                                       Same simulator, same state; served by the default kernel only. */

#define CN_FLAG_RISK_FAITHFUL   8u  /* the K block and the safety counters come from the reference's OWN perception
                                       chain -- gradient typing, scan segmentation, segment confirmation, the
                                       uuid-dict tracker, collision cone (ENV:270-1005), computed in float64 like
                                       CPython does -- instead of the ideal-association restatement.  A second
                                       kernel runs behind the step kernel on the same stream; the tracker record
                                       becomes a 4th plane of the state blob.  Not combinable with
                                       CN_FLAG_ENV_ORIGINAL or the fused gather of cn_step_gather. */

#define CN_FLAG_GATHER_STAGE    16u /* reserve a staging tile per CTA for cn_step_gather_async (the pipelined fused
                                       all-gather).  Costs [tile, D] floats of shared memory per CTA; no effect on results. */

#define CN_FLAG_GATHER_WIRE16   32u /* with CN_FLAG_GATHER_STAGE: the pipelined gather will use the 16-bit wire format; the
                                       tile size is then chosen so that a tile of int16 rows is a whole number of 16-byte
                                       units (bulk copies).  No effect on results. */

/* behaviour kinds (crowd_behaviors/simulate_*.py, SURVEY table P') */
#define CN_BEHAVIOR_RANDOM 0   /* U(-speed, speed)^2 redrawn every period */
#define CN_BEHAVIOR_TABLE  1   /* fixed per-pedestrian direction table * speed */

/*
 * World + sensor + episode configuration.  One struct, plain data; the Python
 * side mirrors it with ctypes.Structure.  Defaults (cn_config_default) are the
 * reference's constants, each cited in DESIGN.md "constants".
 */
typedef struct cn_config {
    uint32_t struct_size;        /* = sizeof(cn_config), ABI check */
    uint32_t flags;
    int32_t  n_envs;             /* E: worlds stepped per call */
    int32_t  n_peds;             /* N <= CN_MAX_PEDS */
    int32_t  n_samples;          /* R: LiDAR samples (scan_ranges, CFG:6); R-1 rays observed */
    int32_t  k_obstacles;        /* K: slots in the perceived-risk block (ENV:55) */
    int32_t  max_steps;          /* episode cap (td3.yaml:7) */
    int32_t  env_id_offset;      /* global id of local env 0 (multi-GPU sharding) */
    uint64_t seed;

    float dt;                    /* control period, ENV:1201 */
    float room_xmin, room_xmax, room_ymin, room_ymax;   /* inner wall faces */
    float start_x, start_y, start_yaw;  /* put_robot_in_world_*.launch:3-8 */
    float goal_x, goal_y;               /* desired_pose, CFG:10-13 */
    float heading_off_x, heading_off_y; /* starting_pose added in ENV:223-224 */

    float max_range;             /* max_scan_range, CFG:7 */
    float collision_range;       /* min_scan_range, CFG:8 */
    float sensor_min_range;      /* XACRO:164 */
    float sensor_sweep;          /* 6.28 rad, XACRO:160 */
    float mount_x;               /* scan frame offset, URDF:137 */
    float hit_angle_inc_deg;     /* UTL:113 angle increment (Py2 int division) */

    float ped_radius;            /* WORLD:109 */
    float robot_radius;          /* pedestrian<->robot contact stand-in */
    float cp_radius;             /* collision-cone circle, ENV:823 */
    float waypoint_radius;       /* ENV:250 */
    float goal_box;              /* ENV:1285,1303 */

    float rep_strength;          /* contact stand-in: A  [m/s]   */
    float rep_range;             /*                   B  [m]     */
    float rep_cutoff;            /* extra gap beyond r_i + r_j where the term is 0 */
    float layout_jitter;         /* U(-j, j) added to each pedestrian start pose */
    float wheel_accel;           /* wheel-speed ramp of libgazebo_ros_diff_drive [m/s^2] (XACRO:70); 0 = instantaneous */
    int32_t n_substeps;          /* kinematic sub-steps per control period (XACRO:65: 100 Hz -> 15); >= 1 */

    int32_t n_behaviors;         /* env behaviour = global_env_id % n_behaviors */
    int32_t behavior_kind[CN_MAX_BEHAVIORS];
    float   behavior_speed[CN_MAX_BEHAVIORS];
    int32_t behavior_period_ticks[CN_MAX_BEHAVIORS];  /* tick = dt / 3 */
    int32_t behavior_stagger_ticks[CN_MAX_BEHAVIORS]; /* per-pedestrian offset */
    float   behavior_table[CN_MAX_BEHAVIORS][CN_MAX_PEDS][2];
    float   ped_layout[CN_MAX_PEDS][2];               /* initial poses (world file) */
} cn_config;

typedef struct cn_handle cn_handle;

/* Fill *cfg with the reference's training-world constants (3 m room, 14
 * pedestrians, 360 samples, K = 8).  Replaces the rosparam load of
 * configs/turtlebot3_world.yaml + the world/xacro constants. */
int cn_config_default(cn_config* cfg);

/* Observation width (R-1) + 7 + 4K  (start_td3_training.py:88); (R-1) + 4 with CN_FLAG_ENV_ORIGINAL. */
int cn_obs_dim(const cn_config* cfg);

/* Bytes of the opaque state blob for this config (cn_get_blob / cn_set_blob). */
size_t cn_blob_bytes(const cn_config* cfg);

/* Allocate one arena of device memory for cfg->n_envs worlds on `device`.
 * Replaces Env.__init__ (ENV:43-168) + the Gazebo world load. */
int cn_create(const cn_config* cfg, int device, cn_handle** out);
int cn_destroy(cn_handle* h);

/* Reset the envs whose mask byte is non-zero (mask_dev == NULL: all) and write
 * their first observation rows into obs_dev [E, D] (rows of unmasked envs are
 * untouched).  Replaces Env.reset (ENV:1227-1263) + gazebo/reset_simulation. */
int cn_reset(cn_handle* h, const uint8_t* mask_dev, float* obs_dev, void* stream);

/* One control period for every env.  Replaces Env.step (ENV:1164-1225), the
 * 0.15 s of Gazebo physics + crowd mover behind it, get_state (ENV:245-1044)
 * and compute_reward (ENV:1046-1162).
 *   action_dev [E, 2] (v, w); obs_dev [E, D] row-major; reward_dev [E];
 *   done_dev [E]: 0 running, 1 = episode ended on this step (obs row is the terminal observation),
 *   2 = this step was an auto-reset (CN_FLAG_AUTO_RESET only; transition to be skipped).
 * When obs_dev is device memory the kernel writes the rows in place while it runs ("direct rows": the no-return fill
 * first, then the rays that hit something, the pose columns and the K block): the rows are complete when the launch
 * has completed, as stream order guarantees for any consumer; nothing else about the call changes.  Rows in
 * host-mapped memory (a pinned buffer passed as obs_dev) are staged in shared memory and leave by one bulk store
 * per tile, as they do for cn_reset and for every fused-gather entry point below. */
int cn_step(cn_handle* h, const float* action_dev, float* obs_dev,
            float* reward_dev, uint8_t* done_dev, void* stream);

/* n_steps control periods back to back (n_steps launches enqueued by ONE call: open-loop action batches, action
 * repeat / frame skip, benchmarks).  Step i reads action_dev + i * action_stride (floats; 0 = the same batch every
 * step, 2 * E = a [n_steps, E, 2] array) and writes reward_dev + i * out_stride, done_dev + i * out_stride
 * (elements; 0 = overwrite, E = [n_steps, E] arrays); obs_dev [E, D] holds the rows of the LAST step. */
int cn_step_n(cn_handle* h, int n_steps, const float* action_dev, size_t action_stride, float* obs_dev,
              float* reward_dev, uint8_t* done_dev, size_t out_stride, void* stream);

/* The same n_steps launches captured once into a CUDA graph owned by the library: cn_graph_launch replays them with
 * one driver call (no per-launch host cost, no torch needed).  The buffers are baked into the graph and must stay
 * valid until cn_graph_destroy; the handle must outlive the graph. */
typedef struct cn_graph cn_graph;
int cn_graph_create(cn_handle* h, int n_steps, const float* action_dev, size_t action_stride, float* obs_dev,
                    float* reward_dev, uint8_t* done_dev, size_t out_stride, cn_graph** out);
int cn_graph_launch(cn_graph* g, void* stream);
int cn_graph_destroy(cn_graph* g);

/* cn_step with the observation all-gather FUSED into the kernel (multi-GPU, one process per GPU):
 * besides obs_dev (this rank's [E, D] row block inside its own [E_total, D] gather buffer) the kernel
 * stores every tile of rows into the same row block of each peer's gather buffer -- peer-mapped device
 * pointers (CUDA IPC / symmetric memory), already offset to this rank's first row -- with 16-byte stores
 * over NVLink, in the order the caller lists the peers (list them starting at rank + 1 so that the ranks do not
 * all write to the same GPU at once).  The local rows leave by bulk TMA store.  The caller orders consumers behind
 * the launch with a cross-rank barrier.  Replaces the ncclAllGather SURVEY.md 8(e) puts after the step.
 * n_peers <= 8 (0 = plain cn_step). */
int cn_step_gather(cn_handle* h, const float* action_dev, float* obs_dev, float* const* peer_obs_dev, int n_peers,
                   float* reward_dev, uint8_t* done_dev, void* stream);

/* cn_step_gather with the collective's SYNCHRONISATION fused in as well: nothing but the step kernel runs per step.
 * Every rank owns an array of n_ranks 64-bit arrival counters (zeroed once), one slot per SOURCE rank, in memory its
 * peers can address.  All step counting is done on the device, so a captured CUDA graph of such steps stays correct
 * on every replay.
 *   peer_arrive_dev[p]  peer-mapped address of THIS rank's slot in peer p's array.  Every CTA makes its stores
 *                       performed (GPU-scope fence) and counts itself on a device-local word; the LAST CTA of the
 *                       launch then adds cn_kernel_ctas() to that slot on every peer with one system-scope release --
 *                       one signal per launch, not per CTA -- and to this rank's own slot of its own array, which
 *                       therefore counts the rank's own progress.  A rank holds all rows of its latest step once every
 *                       other slot has caught up with its own slot (equal shards); cn_gather_wait holds a stream
 *                       until then.
 *   arrive_local_dev, n_ranks, rank, wait_back   before its first store into the peers a CTA of step t (t = own slot
 *                       / cn_kernel_ctas()) waits until every other rank's slot shows all arrivals of that rank's step
 *                       t - wait_back (0 = no wait).  With three gather buffers in rotation pass 2: "everyone has
 *                       launched step t-2, so everyone is done reading the buffer step t overwrites".
 *   obs_mc_dev, arrive_mc_dev      optional NVSwitch MULTICAST addresses of this rank's row block / of this rank's
 *                       slot (CUDA multicast objects, e.g. torch symmetric memory's multicast_ptr): when non-NULL the
 *                       rows leave as multimem.st and the signal as multimem.red -- one store, the switch replicates
 *                       it into every rank's buffer, this rank's included -- and the unicast lists are ignored.
 * Waits are bounded (2 s): a lost peer is counted (cn_gather_timeouts) instead of hanging the GPU.
 * Default kernel only; not with CN_FLAG_RISK_FAITHFUL. */
int cn_step_gather_signal(cn_handle* h, const float* action_dev, float* obs_dev, float* const* peer_obs_dev,
                          unsigned long long* const* peer_arrive_dev, int n_peers, float* obs_mc_dev,
                          unsigned long long* arrive_mc_dev, unsigned long long* arrive_local_dev, int n_ranks,
                          int rank, int wait_back, float* reward_dev, uint8_t* done_dev, void* stream);
/* The PIPELINED fused all-gather: the kernel of step t+1 forwards the rows of step t.  At its start every CTA bulk-loads
 * its tile of the previous step's rows (push_src_dev: this rank's row block of the PREVIOUS gather buffer) into a
 * staging tile and bulk-stores it into every peer (push_peer_dev[p]: the same row block of peer p's copy of that
 * buffer), so the NVLink transfer runs under the step's compute instead of behind it; the arrival counters (as in
 * cn_step_gather_signal, slot per source rank, device-side step counting, wait_back) are signalled at the end of the
 * kernel.  Every rank therefore holds all rows of step t once the kernel of step t+1 -- or cn_gather_flush, the
 * push-only launch for the rows of the last step -- has run on every rank (cn_gather_wait).  n_peers = 0: nothing is
 * forwarded (the first step after a reset).  Handle created with CN_FLAG_GATHER_STAGE; default kernel only; not with
 * CN_FLAG_RISK_FAITHFUL.
 *
 * 16-bit wire format (wire_out_dev != NULL / wire16 != 0): every value of an observation row is a whole number of
 * thousandths -- that is how the row is rounded (ENV:1042) -- so the rows can travel as int16 thousandths, HALF the
 * NVLink bytes, and be rebuilt bit for bit (-0.0 travels as -32768; |value| must stay below 32.768: a value that does
 * not fit saturates and is counted, see cn_gather_timeouts).  The step kernel then ALSO writes its finished rows as
 * int16 into wire_out_dev (this rank's row block of its own [E_total, D] int16 wire buffer of the step), and
 * push_src_dev / push_peer_dev address int16 row blocks of the previous step's wire buffers.  The receiver turns the
 * peers' rows back into fp32 either with cn_gather_decode16 (all rows of [0, rows_total) outside its own [row_lo,
 * row_hi); behind cn_gather_wait) or INSIDE the next step kernel: dec_wire_dev / dec_obs_dev = this rank's whole int16
 * wire buffer and fp32 gather buffer of the step the peers' PREVIOUS kernels delivered; at the END of its own step every
 * CTA checks that each peer has completed as many pushing launches as this rank had before this one (that certifies
 * the delivery; by then it is long true) and rebuilds the rows of its tile in every other rank's block while its own
 * pushes drain.  One launch per step then computes step t+1, forwards step t and finishes the gather of step t-1;
 * rotate FOUR buffers (a step's buffer must survive until two launches later). */
int cn_step_gather_async(cn_handle* h, const float* action_dev, float* obs_dev, int16_t* wire_out_dev, const void* push_src_dev,
                         void* const* push_peer_dev, unsigned long long* const* peer_arrive_dev, int n_peers,
                         unsigned long long* arrive_local_dev, int n_ranks, int rank, int wait_back,
                         const int16_t* dec_wire_dev, float* dec_obs_dev,
                         float* reward_dev, uint8_t* done_dev, void* stream);
int cn_gather_flush(cn_handle* h, int wire16, const void* push_src_dev, void* const* push_peer_dev,
                    unsigned long long* const* peer_arrive_dev, int n_peers, unsigned long long* arrive_local_dev,
                    int n_ranks, int rank, int wait_back, void* stream);
int cn_gather_decode16(cn_handle* h, const int16_t* wire_dev, float* obs_all_dev, long long row_lo, long long row_hi,
                       long long rows_total, void* stream);
/* Hold `stream` until every other rank's slot of arrive_local_dev[n_ranks] has caught up with this rank's own slot
 * (one small kernel; bounded like the in-kernel wait). */
int cn_gather_wait(cn_handle* h, const unsigned long long* arrive_local_dev, int n_ranks, int rank, void* stream);
/* Number of bounded gather waits that gave up + of observation values that did not fit the 16-bit wire format since
 * cn_create (0 in a healthy run).  Synchronous on `stream`. */
int cn_gather_timeouts(cn_handle* h, unsigned int* out_host, void* stream);
/* CTAs one step launch of this handle consists of (= arrivals per peer per step). */
int cn_kernel_ctas(const cn_handle* h);

/* Per-env counters [E, 4] int32: success, ego violations, social violations,
 * obstacle-present steps.  Replaces get_episode_status / get_*_violation_status
 * (ENV:1265-1283). */
int cn_get_counters(cn_handle* h, int32_t* out_dev, void* stream);

/* Clear the sticky done flag (the drivers' `env.done = False`, TD3DRV:116). */
int cn_clear_done(cn_handle* h, const uint8_t* mask_dev, void* stream);

/* Whole-state snapshot to / from host memory (exact resume, parity fixtures).
 * Synchronous on `stream`. */
int cn_get_blob(cn_handle* h, void* host, size_t bytes, void* stream);
int cn_set_blob(cn_handle* h, const void* host, size_t bytes, void* stream);

/* Debug taps of the last cn_step / cn_reset: pre-rounding ranges [E, R-1] f32
 * and hit ids [E, R-1] u8 (0xFF none, 0xFE wall, else pedestrian).  Either may
 * be NULL.  Enabling them adds global stores to the step kernel. */
int cn_set_debug_taps(cn_handle* h, float* ranges_dev, uint8_t* hit_ids_dev);

/* Number of kernels launched by this handle since creation. */
int64_t cn_launch_count(const cn_handle* h);
/* Name of the step-kernel variant this handle launches ("cn_flat_kernel": compacted work lists, the default;
 * "cn_env_kernel": one warp per world, selected with the environment variable CN_KERNEL=warp at cn_create) and
 * the number of worlds per CTA a plain cn_step uses (the fused-gather entry points use the staged layout, whose CTA count
 * cn_kernel_ctas reports).  For benchmarks and profiles; results are bit-identical. */
const char* cn_kernel_name(const cn_handle* h);
int cn_kernel_tile(const cn_handle* h);
/* Host-only planning query (no GPU needed): the tile (worlds per CTA), CTA size and dynamic shared memory the staged
 * instance of the step kernel (reset, fused gather, rows in host-mapped memory) would use for this config on a device with n_sms SMs and smem_per_sm bytes of shared memory per SM. */
int cn_plan_tile(const cn_config* cfg, int n_sms, size_t smem_per_sm, int* tile, int* threads, size_t* smem_bytes);
/* The same for the direct-rows instance a plain cn_step into device memory launches (rows written in place, no staging
 * tile): 256 threads x 6 resident CTAs per SM, or 384 x 4 for batches that need several waves anyway. */
int cn_plan_tile_direct(const cn_config* cfg, int n_sms, size_t smem_per_sm, int* tile, int* threads, size_t* smem_bytes);

const char* cn_last_error(void);
int cn_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CROWDNAV_H */
