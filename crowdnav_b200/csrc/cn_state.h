/*
 * cn_state.h -- HBM state layout of a batch of worlds, and the constants
 * derived once per config on the host.  Shared by the CUDA library and by the
 * CPU oracle so a state blob can be moved between them bit for bit.
 *
 * Three planes, each an array of per-env records that are contiguous in
 * memory (env-major), so a CTA that owns a tile of consecutive envs moves each
 * plane's tile with ONE 1-D bulk (TMA) copy and a warp that owns one env reads
 * its record with fully used 128-B lines:
 *
 *   robot  [E][16] words   64 B / env
 *   ped_a  [E][N][4] words 16 B / pedestrian: x, y (int32 grid), vx, vy (f32)
 *   ped_b  [E][N][4] words 16 B / pedestrian: last hit point x, y (f32, 3 dp),
 *                                             resample timer (int32 ticks), flags
 *
 * Blob = 16-word header, then the three planes back to back (+ the tracker plane
 * trk [E][CNF_WORLD_WORDS] of cn_faithful_state.h with CN_FLAG_RISK_FAITHFUL).
 */
#ifndef CN_STATE_H
#define CN_STATE_H

#include "../../include/crowdnav.h"
#include "cn_math.h"
#include "cn_math64.h"
#include "cn_faithful_state.h"

/* robot / episode record, word indices */
enum {
    CN_R_X = 0,      /* int32 grid */
    CN_R_Y = 1,      /* int32 grid */
    CN_R_TH = 2,     /* uint32 binary angle */
    CN_R_V = 3,      /* f32 achieved linear velocity  (odom twist.linear.x) */
    CN_R_W = 4,      /* f32 achieved angular velocity (odom twist.angular.z) */
    CN_R_WPX = 5,    /* f32 waypoint (ENV:80-83,252-265) */
    CN_R_WPY = 6,
    CN_R_PDIST = 7,  /* f32 previous_distance (ENV:1133,1243) */
    CN_R_PHEAD = 8,  /* f32 previous_heading  (ENV:1134,1244) */
    CN_R_PPX = 9,    /* f32 agent_pose_deque[0], rounded 3 dp (ENV:294,1208) */
    CN_R_PPY = 10,
    CN_R_STEP = 11,  /* int32 steps taken this episode */
    CN_R_EPISODE = 12, /* uint32 episodes started */
    CN_R_FLAGS = 13,
    CN_R_CNT0 = 14,  /* ego violations | social violations << 16 (ENV:998-1005) */
    CN_R_CNT1 = 15   /* obstacle-present steps | sanitised actions << 16 */
};
#define CN_RF_DONE     1u
#define CN_RF_SUCCESS  2u
#define CN_RF_FAILURE  4u

/* ped_b flags */
#define CN_PF_TRACKED  1u   /* last hit point valid (object was confirmed last step) */

/* hit ids */
#define CN_HIT_NONE 0xFFu
#define CN_HIT_WALL 0xFEu

#define CN_BLOB_MAGIC 0x56414E43u /* "CNAV" */
#define CN_BLOB_HEADER_WORDS 16

/* action sanitising bounds (agents emit v in [0, 0.22], w in [-2, 2]) */
#define CN_ACT_V_LIMIT 1.0f
#define CN_ACT_W_LIMIT 6.0f
#define CN_WHEEL_SEP   0.160f   /* XACRO:68 */
#define CN_INV_WHEEL_SEP 6.25f  /* exactly 1 / 0.16 */
#define CN_TICKS_PER_STEP 3

typedef struct cn_derived {
    uint32_t inc_bin;        /* binary angle between adjacent LiDAR samples */
    uint32_t hit_inc_bin;    /* UTL:113-123 angle increment as binary angle */
    uint32_t start_th;
    int32_t  start_xi, start_yi;
    int32_t  ped_xmin, ped_xmax, ped_ymin, ped_ymax; /* centre clamp, grid units */
    float    ped_r2;         /* ped_radius^2 */
    float    cp_r2;          /* cp_radius^2 */
    float    apothem;        /* waypoint_radius * cos(pi/64): shapely 64-gon (UTL:301-302) */
    float    cand_d2;        /* (max_range + ped_radius + 1e-3)^2 : LiDAR candidate cull */
    float    goal_lo_x, goal_hi_x, goal_lo_y, goal_hi_y; /* ENV:1303-1319 */
    float    inv_inc_bin;    /* 1 / inc_bin (span rasterisation only) */
    float    inv_dt;         /* RN(1 / dt): velocities are displacement * inv_dt */
    float    inv_cp_span;    /* RN(1 / (max_range - collision_range)), UTL:343 */
    float    max_range_r3;   /* np.around(max_range, 3): the value of a ray with no return */
    float    dt_sub;         /* dt / n_substeps */
    float    wheel_step;     /* wheel_accel * dt_sub: largest wheel-speed change per sub-step (0 = unlimited) */
    int32_t  obs_dim;
    int32_t  pair_cell_shift; /* log2 (grid units) of the contact-prefilter cell: cell / 2 >= contact range */
    int32_t  robot_contact;   /* 1 when collision_range < robot_radius: the LiDAR threshold no longer ends an episode before
                                 the body touches something, so the robot's centre is kept robot_radius off the walls and
                                 a sub-step that would move it INTO a pedestrian's disc is not taken */
    int32_t  rob_xmin, rob_xmax, rob_ymin, rob_ymax; /* robot centre clamp, grid units */
    float    rob_ped_r2;      /* (robot_radius + ped_radius)^2 */
    int32_t  strip_shift;     /* log2 (grid units) of the contact-prefilter STRIP width (cn_flat.cu): >= the contact box */
    uint32_t seed_lo, seed_hi;
} cn_derived;

static inline int cn_derive(const cn_config* c, cn_derived* d) {
    if (c->n_envs < 1 || c->n_peds < 0 || c->n_peds > CN_MAX_PEDS) return -1;
    if (c->n_samples < 3 || c->n_samples > 1025) return -1;   /* 32 chunks of 32 rays */
    if (c->k_obstacles < 0 || c->k_obstacles > CN_MAX_PEDS) return -1;
    if (c->n_behaviors < 1 || c->n_behaviors > CN_MAX_BEHAVIORS) return -1;
    if (!(c->dt > 0.0f) || !(c->max_range > 0.0f)) return -1;
    /* the packed contact prefilter keeps 14-bit coordinates at 2^-8 m: worlds within +-30 m, contact range < 0.24 m */
    if (c->room_xmin < -30.0f || c->room_xmax > 30.0f || c->room_ymin < -30.0f || c->room_ymax > 30.0f) return -1;
    if (c->ped_radius + (c->ped_radius > c->robot_radius ? c->ped_radius : c->robot_radius) + c->rep_cutoff > 0.24f) return -1;
    const double two_pi = 6.283185307179586476925286766559;
    double inc = (double)c->sensor_sweep / (double)(c->n_samples - 1);
    d->inc_bin = (uint32_t)llrint(inc / two_pi * 4294967296.0);
    d->hit_inc_bin = (uint32_t)llrint((double)c->hit_angle_inc_deg / 360.0 * 4294967296.0);
    d->start_th = (uint32_t)(int64_t)llrint((double)c->start_yaw / two_pi * 4294967296.0);
    d->start_xi = (int32_t)llrint((double)c->start_x * 16777216.0);
    d->start_yi = (int32_t)llrint((double)c->start_y * 16777216.0);
    d->ped_xmin = (int32_t)llrint(((double)c->room_xmin + c->ped_radius) * 16777216.0);
    d->ped_xmax = (int32_t)llrint(((double)c->room_xmax - c->ped_radius) * 16777216.0);
    d->ped_ymin = (int32_t)llrint(((double)c->room_ymin + c->ped_radius) * 16777216.0);
    d->ped_ymax = (int32_t)llrint(((double)c->room_ymax - c->ped_radius) * 16777216.0);
    d->robot_contact = (c->collision_range < c->robot_radius) ? 1 : 0;
    d->rob_xmin = (int32_t)llrint(((double)c->room_xmin + c->robot_radius) * 16777216.0);
    d->rob_xmax = (int32_t)llrint(((double)c->room_xmax - c->robot_radius) * 16777216.0);
    d->rob_ymin = (int32_t)llrint(((double)c->room_ymin + c->robot_radius) * 16777216.0);
    d->rob_ymax = (int32_t)llrint(((double)c->room_ymax - c->robot_radius) * 16777216.0);
    { float rr = c->robot_radius + c->ped_radius; d->rob_ped_r2 = rr * rr; }
    d->ped_r2 = c->ped_radius * c->ped_radius;
    d->cp_r2 = c->cp_radius * c->cp_radius;
    d->apothem = (float)((double)c->waypoint_radius * cos(two_pi / 128.0));
    float cd = c->max_range + c->ped_radius + 1e-3f;
    d->cand_d2 = cd * cd;
    d->goal_lo_x = c->goal_x - c->goal_box;
    d->goal_hi_x = c->goal_x + c->goal_box;
    d->goal_lo_y = c->goal_y - c->goal_box;
    d->goal_hi_y = c->goal_y + c->goal_box;
    d->inv_inc_bin = (float)(1.0 / (double)d->inc_bin);
    if (c->n_substeps < 1 || c->n_substeps > 64 || !(c->wheel_accel >= 0.0f)) return -1;
    d->dt_sub = (c->n_substeps == 1) ? c->dt : (float)((double)c->dt / (double)c->n_substeps);
    d->wheel_step = c->wheel_accel * d->dt_sub;
    d->inv_dt = (float)(1.0 / (double)c->dt);
    d->inv_cp_span = (float)(1.0 / ((double)c->max_range - (double)c->collision_range));
    d->max_range_r3 = cn_np_round3(c->max_range);
    if (d->max_range_r3 != c->max_range) return -1;   /* the no-return value must be a whole number of millimetres */
    if ((c->flags & CN_FLAG_ENV_ORIGINAL) && c->k_obstacles != 0) return -1;
    if ((c->flags & CN_FLAG_RISK_FAITHFUL) && (c->flags & CN_FLAG_ENV_ORIGINAL)) return -1;
    d->obs_dim = (c->flags & CN_FLAG_ENV_ORIGINAL) ? (c->n_samples - 1) + 4 : (c->n_samples - 1) + 7 + 4 * c->k_obstacles;
    {
        double lim = (double)(c->ped_radius + (c->ped_radius > c->robot_radius ? c->ped_radius : c->robot_radius))
                     + c->rep_cutoff + 1e-4;
        int sh = 1;
        while (sh < 29 && ldexp(1.0, sh - 1 - 24) < lim) ++sh;
        d->pair_cell_shift = sh;
        d->strip_shift = sh - 1;             /* 2^(sh-1) grid units >= lim + 1e-4 m > the kernel's contact box half-width */
    }
    d->seed_lo = (uint32_t)(c->seed & 0xFFFFFFFFu);
    d->seed_hi = (uint32_t)(c->seed >> 32);
    return 0;
}

static inline size_t cn_robot_words(const cn_config* c) { return (size_t)c->n_envs * CN_ROBOT_WORDS; }
static inline size_t cn_ped_plane_words(const cn_config* c) { return (size_t)c->n_envs * (size_t)c->n_peds * 4; }
static inline size_t cn_trk_words(const cn_config* c) {
    return (c->flags & CN_FLAG_RISK_FAITHFUL) ? (size_t)c->n_envs * CNF_WORLD_WORDS : 0;
}
static inline size_t cn_blob_words(const cn_config* c) {
    return CN_BLOB_HEADER_WORDS + cn_robot_words(c) + 2 * cn_ped_plane_words(c) + cn_trk_words(c);
}
/* 16-word blob header: shape, layout version and the mode bits that change what the planes mean.  cn_set_blob
 * accepts a blob only when all of it matches the handle. */
static inline void cn_blob_header(const cn_config* c, uint32_t* w) {
    for (int i = 0; i < CN_BLOB_HEADER_WORDS; ++i) w[i] = 0u;
    w[0] = CN_BLOB_MAGIC; w[1] = CN_ABI_VERSION;
    w[2] = (uint32_t)c->n_envs; w[3] = (uint32_t)c->n_peds;
    w[4] = (uint32_t)c->n_samples; w[5] = (uint32_t)c->k_obstacles;
    w[6] = c->flags & (CN_FLAG_RISK_FAITHFUL | CN_FLAG_ENV_ORIGINAL);
    w[7] = (c->flags & CN_FLAG_RISK_FAITHFUL) ? (uint32_t)CNF_WORLD_WORDS : 0u;
}
/* float64 constants of the risk_faithful block from the (float) config */
static inline void cnf_params_from_config(const cn_config* c, cnf_params* p) {
    p->n_rays = c->n_samples - 1;
    p->k_obstacles = c->k_obstacles;
    p->topk_highest = (c->flags & CN_FLAG_TOPK_HIGHEST) ? 1 : 0;
    p->pad_ = 0;
    p->inc_deg = (double)c->hit_angle_inc_deg;
    p->max_range = cn_dec64(c->max_range);
    p->min_range = cn_dec64(c->collision_range);
    p->dt = cn_dec64(c->dt);
    p->cp_radius = cn_dec64(c->cp_radius);
    p->track_half = 0.0505;            /* hard-coded at ENV:689 */
}

#endif /* CN_STATE_H */
