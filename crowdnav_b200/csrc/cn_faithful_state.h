/*
 * cn_faithful_state.h -- per-world tracker record of the `risk_faithful` perception block
 * (CN_FLAG_RISK_FAITHFUL) and its constants.  Shared by the CUDA library and the CPU oracle
 * (like cn_state.h) so the record can be compared bit for bit; it is the 4th plane of the blob.
 *
 *   trk [E][CNF_WORLD_WORDS] words:
 *     header (12 words): tracked count, agent_pose_deque state, counters, bounding_box_size, ego score
 *     CNF_TRK_CAP entries of 12 words = the reference's tracked_obstacles dict in insertion order
 *       (ENV:662-671): deque[0], deque[-1] = pose, range (all thousandths, int32), deque length,
 *       speed, vx, vy (float64)
 */
#ifndef CN_FAITHFUL_STATE_H
#define CN_FAITHFUL_STATE_H

#include <stdint.h>

#define CNF_TRK_CAP   32     /* tracked objects kept per world (the reference's dict is unbounded) */
#define CNF_CONF_CAP  48     /* confirmed objects per scan */

enum {
    CNF_H_N = 0,          /* entries in use */
    CNF_H_HAVE_PREV = 1,  /* len(agent_pose_deque) >= 1 before this step's append (ENV:294,1208) */
    CNF_H_PPX = 2,        /* agent_pose_deque[0], thousandths */
    CNF_H_PPY = 3,
    CNF_H_EGO = 4,        /* ego_safety_violation_count (ENV:998-1002) */
    CNF_H_SOCIAL = 5,     /* social_safety_violation_count (ENV:1004-1005) */
    CNF_H_PRESENT = 6,    /* obstacle_present_step_counts (ENV:653-654) */
    CNF_H_OVERFLOW = 7,   /* objects dropped because a capacity above was reached */
    CNF_H_BBOX = 8,       /* float64 bounding_box_size (ENV:287-290), words 8-9 */
    CNF_H_EGOSCORE = 10,  /* float64 ego_score_collision_prob (ENV:864,879), words 10-11 */
    CNF_HDR_WORDS = 12
};
enum {
    CNF_E_PX = 0, CNF_E_PY = 1,   /* deque[0] */
    CNF_E_LX = 2, CNF_E_LY = 3,   /* deque[-1] == entry[1] */
    CNF_E_DIST = 4,               /* entry[2] */
    CNF_E_NDEQ = 5,               /* len(deque): 1 or 2 */
    CNF_E_SPEED = 6,              /* float64 entry[5] (-1 until two poses are known) */
    CNF_E_VX = 8, CNF_E_VY = 10,  /* float64 entry[6] */
    CNF_ENTRY_WORDS = 12
};
#define CNF_WORLD_WORDS (CNF_HDR_WORDS + CNF_TRK_CAP * CNF_ENTRY_WORDS)   /* 396 words = 1584 B */

/* constants of the block as float64 (the reference computes in Python floats) */
typedef struct cnf_params {
    int32_t n_rays;        /* R - 1 */
    int32_t k_obstacles;
    int32_t topk_highest;  /* CN_FLAG_TOPK_HIGHEST */
    int32_t pad_;
    double inc_deg;        /* UTL:113 angle increment in degrees */
    double max_range;      /* 0.6 */
    double min_range;      /* min_scan_range, CFG:8 */
    double dt;             /* agent_vel_timestep / tracker timelapse = the control period */
    double cp_radius;      /* 0.178, ENV:823 */
    double track_half;     /* 0.0505, ENV:689 */
} cnf_params;

#endif
