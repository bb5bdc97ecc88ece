/* cn_kernel.h -- launch interface between the C ABI (cn_abi.cu) and the kernels (cn_step.cu). */
#ifndef CN_KERNEL_H
#define CN_KERNEL_H

#include <cuda_runtime.h>
#include "cn_state.h"

#define CN_TILE 14       /* worlds per CTA: one warp each */
#define CN_POSE_WARPS 2   /* extra warps that do the pose-only work of the whole tile (lane = world) */
#define CN_CTA_THREADS (32 * (CN_TILE + CN_POSE_WARPS))

/* Everything the step kernel needs, passed by value (constant bank); the
 * per-behaviour tables and the layout stay in the device copy of cn_config. */
struct cn_kparams {
    uint32_t* robot;
    uint32_t* ped_a;
    uint32_t* ped_b;
    const float* action;
    float* obs;
    float* reward;
    uint8_t* done;
    const uint8_t* mask;
    float* dbg_ranges;
    uint8_t* dbg_hid;
    float* obs_peers[8];        /* fused all-gather: this rank's row block inside each PEER's [E_total, D] buffer */
    int n_obs_peers;
    /* fused all-gather with in-kernel signalling (cn_step_gather_signal; flat kernel only).  Every rank owns an array of
     * 64-bit arrival counters, one slot per SOURCE rank.  After its rows have been stored into a peer, every CTA adds 1
     * to its own slot of that peer's array (release, system scope); before the first store into the peers the CTA waits
     * until every other rank's slot of THIS rank's array has reached arrive_wait (the peers are done with the buffer
     * about to be overwritten).  obs_mc / arrive_mc: NVSwitch multicast addresses (multimem.st / .red), one store
     * reaches every rank's buffer -- used instead of the unicast lists when non-NULL. */
    unsigned long long* arrive_peers[8];   /* peer p's array, already offset to this rank's slot */
    unsigned long long* arrive_local;      /* this rank's array [arrive_slots]; its OWN slot counts this rank's CTAs */
    int arrive_slots, arrive_self;
    int arrive_back;                       /* wait for the peers' step (t - arrive_back); 0 = no wait */
    unsigned int ctas_per_step;
    float* obs_mc;
    unsigned long long* arrive_mc;
    /* pipelined variant (cn_step_gather_async): the kernel of step t+1 forwards the rows of step t -- push_src, this
     * rank's row block of the previous gather buffer -- to the peers (push_peers) at its START: thread 0 of every CTA
     * bulk-loads its tile of old rows into a staging tile and bulk-stores it into every peer, so the transfer runs
     * under the step's compute; arrivals are signalled at the end of the kernel.  Needs CN_FLAG_GATHER_STAGE. */
    const void* push_src;
    void* push_peers[8];
    int n_push_peers;
    int push_bulk_ok;               /* push_src and every push peer are 16-B aligned */
    /* 16-bit wire format of the pipelined gather: every observation value is a whole number of thousandths (that is
     * how the row is rounded), so a row travels as int16 thousandths -- half the NVLink bytes -- and the receiver
     * rebuilds the identical fp32 row (cn_gather_decode16).  wire_out: this rank's row block of its OWN wire buffer
     * for the step being computed (the kernel encodes its finished rows into it; the next step's kernel forwards it);
     * push_src / push_peers then address int16 rows.  -0.0 travels as -32768; a value outside +-32.767 saturates and
     * is counted in gather_timeouts[2]. */
    int16_t* wire_out;
    int push_wire16;
    /* ... and the receiving side inside the SAME kernel: dec_wire / dec_obs are this rank's whole [E_total, D] int16 wire
     * buffer and fp32 gather buffer of the step whose rows the peers' PREVIOUS kernels delivered (certified by the guard,
     * which therefore runs with wait_back = 1 before anything else); CTA b rebuilds rows [b W, b W + nE) of every other
     * rank's block.  NULL: nothing to decode in this launch. */
    const int16_t* dec_wire;
    float* dec_obs;
    unsigned int* gather_done;      /* device word: CTAs of the running launch whose rows have reached the peers (the last one signals) */
    unsigned int* gather_timeouts;  /* device counter: waits given up after CN_GATHER_WAIT_NS (diagnostics, never hangs the GPU) */
    const cn_config* cfg;       /* device copy */
    cn_derived d;
    int n_envs, n_peds, n_samples, k_obstacles, max_steps, env_id_offset, n_behaviors, n_substeps;
    uint32_t flags;
    int obs_bulk_ok;            /* obs base (and every peer's) is 16-B aligned: tile rows may leave by bulk store if the tile size allows */
    int act_bulk_ok;            /* action base is 16-B aligned: the tile's actions arrive by bulk load */
    int beh_kind[CN_MAX_BEHAVIORS];
    float beh_speed[CN_MAX_BEHAVIORS];
    int beh_period[CN_MAX_BEHAVIORS];
    int beh_stagger[CN_MAX_BEHAVIORS];
    float dt;
    float room_xmin, room_xmax, room_ymin, room_ymax;
    float goal_x, goal_y, heading_off_x, heading_off_y;
    float max_range, collision_range, sensor_min_range, mount_x;
    float ped_radius, robot_radius, goal_box;
    float rep_strength, rep_range, rep_cutoff, layout_jitter;
};

/* cn_flat.cu: shared-memory layout of one tile of W worlds (byte offsets), computed once per handle on the host */
struct cn_flat_layout {
    int W;                      /* worlds per CTA */
    int threads;                /* threads per CTA: 128, 256 or 512 */
    int plain_store;            /* 1: write the tile back with cooperative 16-byte stores, 0: bulk TMA stores */
    int pdl;                    /* 1: launch with programmatic stream serialization (back-to-back steps overlap their start-up) */
    uint32_t magic_n;           /* ceil(2^32 / N): world index of a flat pedestrian index by __umulhi */
    uint32_t cap_wg, cap_pg;    /* capacity of the wall / pedestrian ray-group lists */
    uint32_t off_pa, off_pb, off_pa2, off_act, off_obs, off_sc, off_rec, off_pk, off_peers,
             off_clist, off_clw, off_rlist, off_olist, off_mark, off_wg, off_pg, off_cnt, off_bar;
    int gather_debug;           /* diagnostics (CN_GATHER_DEBUG): 1 no guard wait, 2 no arrival signal, 4 no row transfer */
    uint32_t off_strips;        /* contact prefilter: per world 2 axes x 32 strips x (1 or 2) words of pedestrian bits */
    uint32_t strip_words;       /* words per strip mask: 1 (N <= 32) or 2 */
    uint32_t off_stage;         /* CN_FLAG_GATHER_STAGE: [W, D] floats for the previous step's rows on their way to the peers; else 0 */
    uint32_t total;             /* dynamic shared memory per CTA */
    /* "direct rows" variant (plain single-GPU steps): the observation rows are not staged in shared memory -- the
     * no-return fill and the few owned rays / pose columns / K slots go straight to the caller's [E, D] buffer (L2
     * absorbs them) -- which takes 4 D bytes per world out of the tile and lets an SM hold all its worlds at once */
    int obs_direct;
    uint32_t strip_mask;        /* strips per axis - 1 */
    uint32_t off_strips_y;      /* the y-axis strip masks (off_strips: the x-axis ones) */
    uint32_t off_fillc;         /* direct layout: constant tile the bulk fill stores read ... */
    uint32_t fillc_bytes;       /* ... and its size (512 bytes up to one row of ray columns, as the tile leaves room) */
};
/* stage: 0 none, 1 fp32 staging tile, 2 int16 staging tile; direct: 1 = rows straight to global memory (no stage) */
int cn_flat_make_layout(int n_peds, int n_samples, int obs_dim, int tile, int threads, int stage, int direct, cn_flat_layout* L);
int cn_flat_pick_tile(int n_peds, int n_samples, int obs_dim, int n_envs, int n_sms, size_t smem_per_sm, int stage, int direct, cn_flat_layout* L);
/* push-only launch of the pipelined gather (the rows of the LAST step, which no later step kernel will forward) */
cudaError_t cn_launch_push_kernel(const cn_kparams& P, const cn_flat_layout& L, cudaStream_t stream);
/* int16 thousandths -> fp32 rows for every row of [0, rows_total) outside [row_lo, row_hi) (this rank's own rows are
 * already there in fp32) */
cudaError_t cn_launch_wire_decode(const int16_t* wire, float* obs_all, long long row_lo, long long row_hi,
                                  long long rows_total, int obs_dim, cudaStream_t stream);
cudaError_t cn_launch_flat_kernel(const cn_kparams& P, const cn_flat_layout& L, int mode, cudaStream_t stream);

/* cn_abi.cu: raise cudaFuncAttributeMaxDynamicSharedMemorySize of `func` on the CURRENT device to at least `smem`
 * (the attribute is per device; the table is keyed by (slot, device) and guarded by a mutex).  Slots: 0-9 flat
 * kernel (mode * 5 + CTA-size index), 10-13 warp kernel (NPL, mode), 14 risk_faithful kernel, 16-20 flat kernel
 * with direct rows (CTA-size index). */
#define CN_ATTR_SLOTS 24
#define CN_ATTR_MAX_DEVICES 64
cudaError_t cn_ensure_smem_attr(const void* func, int slot, size_t smem);

size_t cn_kernel_smem_bytes(int n_peds, int n_samples, int obs_dim);
cudaError_t cn_launch_env_kernel(const cn_kparams& P, int mode /*0 step, 1 reset*/, cudaStream_t stream);
cudaError_t cn_launch_gather_wait(const unsigned long long* counters, int n_slots, int self_slot,
                                  unsigned int* timeouts, cudaStream_t stream);
#define CN_GATHER_WAIT_NS 2000000000ull   /* a peer that is 2 s late is treated as lost: count it and go on */
cudaError_t cn_launch_clear_done(uint32_t* robot, const uint8_t* mask, int E, cudaStream_t stream);
cudaError_t cn_launch_counters(const uint32_t* robot, int32_t* out, int E, cudaStream_t stream);

/* cn_faithful.cu: the risk_faithful perception block behind the step kernel (CN_FLAG_RISK_FAITHFUL) */
size_t cn_faithful_smem_bytes(int n_rays);
cudaError_t cn_launch_faithful(const cn_config* cfg, const uint32_t* robot, uint32_t* trk, const float* ranges,
                               float* obs, const uint8_t* mask, int obs_dim, cudaStream_t stream);
cudaError_t cn_launch_faithful_counters(const uint32_t* robot, const uint32_t* trk, int32_t* out, int E, cudaStream_t stream);

#endif
