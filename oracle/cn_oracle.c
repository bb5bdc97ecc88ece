/*
 * cn_oracle.c -- CPU ORACLE for the crowd-navigation env-step.  TEST
 * INFRASTRUCTURE ONLY: nothing under crowdnav_b200/ may import, link or call
 * this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker or the timed CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" for the simulator half -- the reference
 * delegates robot kinematics, pedestrian physics and the LiDAR to Gazebo/ODE,
 * which is not in /root/reference and cannot run here.
 * The env half (scan cleaning, hit points, waypoint, heading / distance, reward,
 * done, CP formulas, observation assembly) IS pinned: tests/golden/ holds
 * vectors and whole-episode traces produced by running the reference's own
 * utils.py and Env.reset/step/get_state/compute_reward in this container
 * (tests/ref_harness.py, tests/gen_golden.py) on physics injected from this
 * simulator; tests/test_oracle_golden.py replays them against this file.
 *
 * One world per call, scalar code, brute-force LiDAR (every ray against every
 * primitive).  Each function cites the reference lines it restates:
 *   ENV   = turtlebot3_rl_sim/src/environment_stage_1_nobonus.py
 *   UTL   = turtlebot3_rl_sim/src/utils.py
 *   CROWD = turtlebot3_rl_sim/src/crowd_behaviors/simulate_crowd.py
 *   FAKE  = turtlebot3_simulations/turtlebot3_fake/src/turtlebot3_fake.cpp
 *   XACRO = turtlebot3_description/urdf/turtlebot3_burger.gazebo.xacro
 *
 * Arithmetic is fp32 built from the primitives in cn_math.h (the numeric
 * spec), compiled with -ffp-contract=off, so the CUDA kernel can be compared
 * bit for bit.
 */
#include <stdlib.h>
#include <stdio.h>
#include "../crowdnav_b200/csrc/cn_state.h"

typedef struct orc_ctx {
    cn_config cfg;
    cn_derived d;
} orc_ctx;

/* ------------------------------------------------------------------------ */
orc_ctx* orc_create(const cn_config* cfg) {
    if (!cfg || cfg->struct_size != sizeof(cn_config)) return NULL;
    orc_ctx* c = (orc_ctx*)calloc(1, sizeof(orc_ctx));
    if (!c) return NULL;
    c->cfg = *cfg;
    if (cn_derive(cfg, &c->d) != 0) { free(c); return NULL; }
    return c;
}
void orc_destroy(orc_ctx* c) { free(c); }
size_t orc_blob_bytes(const orc_ctx* c) { return cn_blob_words(&c->cfg) * 4; }
int orc_obs_dim(const orc_ctx* c) { return c->d.obs_dim; }

static uint32_t* blob_robot(const orc_ctx* c, uint32_t* blob) { return blob + CN_BLOB_HEADER_WORDS; }
static uint32_t* blob_ped_a(const orc_ctx* c, uint32_t* blob) { return blob_robot(c, blob) + cn_robot_words(&c->cfg); }
static uint32_t* blob_ped_b(const orc_ctx* c, uint32_t* blob) { return blob_ped_a(c, blob) + cn_ped_plane_words(&c->cfg); }
static uint32_t* blob_trk(const orc_ctx* c, uint32_t* blob) {
    return (c->cfg.flags & CN_FLAG_RISK_FAITHFUL) ? blob_ped_b(c, blob) + cn_ped_plane_words(&c->cfg) : NULL;
}

void orc_init_blob(const orc_ctx* c, uint32_t* blob) {
    memset(blob, 0, orc_blob_bytes(c));
    cn_blob_header(&c->cfg, blob);
}

static float f_of(uint32_t u) { return cn_bits2f(u); }
static uint32_t u_of(float f) { return cn_f2bits(f); }

/* ---- A: waypoint ---------------------------------------------------------
 * utils.get_local_goal_waypoints (UTL:296-314): intersection of the segment
 * agent->goal with the boundary of Point(agent).buffer(0.3).  shapely's default
 * buffer is a 64-gon with vertices at k*pi/32, so the crossing lies at
 * apothem / cos(offset from the sector's mid-angle).  No crossing (goal inside
 * the ring) -> the (-goal_x, goal_y) fallback of UTL:310-312.
 */
static void waypoint(const orc_ctx* c, float xf, float yf, float* wx, float* wy) {
    float gxr = c->cfg.goal_x - xf, gyr = c->cfg.goal_y - yf;
    float L = sqrtf(fmaf(gxr, gxr, gyr * gyr));
    if (L > 0.0f) {
        float phi = cn_atan2(gyr, gxr);
        float t = phi * 10.1859163578813f;            /* 32 / pi */
        float m = floorf(t);
        float delta = ((t - m) - 0.5f) * 0.0981747704246810f; /* pi / 32 */
        float z = delta * delta;
        float cd = fmaf(fmaf(4.1666668e-2f, z, -0.5f), z, 1.0f);
        /* crossing at distance rho = apothem / cd; it exists iff L >= rho  <=>  L * cd >= apothem */
        if (L * cd >= c->d.apothem) {
            float sc = c->d.apothem / (cd * L);          /* rho / L */
            *wx = fmaf(sc, gxr, xf);
            *wy = fmaf(sc, gyr, yf);
            return;
        }
    }
    *wx = -c->cfg.goal_x;
    *wy = c->cfg.goal_y;
}

/* ENV:1285-1319 half-open boxes */
static int in_box(float x, float y, float lox, float hix, float loy, float hiy) {
    return (x <= hix) && (x > lox) && (y <= hiy) && (y > loy);
}
static int in_goal_box(const orc_ctx* c, float x, float y) {
    return in_box(x, y, c->d.goal_lo_x, c->d.goal_hi_x, c->d.goal_lo_y, c->d.goal_hi_y);
}

/* Env.get_distance_to_goal (ENV:191-203), unrounded */
static float dist_to_wp(float xf, float yf, float wx, float wy) {
    float dx = xf - wx, dy = yf - wy;
    return sqrtf(fmaf(dx, dx, dy * dy));
}
/* Env.get_heading_to_goal (ENV:222-237), unrounded; starting_pose is ADDED to
 * the odom position here (and only here) */
static float heading_to_wp(const orc_ctx* c, float xf, float yf, float yaw, float wx, float wy) {
    float px = xf + c->cfg.heading_off_x, py = yf + c->cfg.heading_off_y;
    float h = cn_atan2(wy - py, wx - px) - yaw;
    if (h > CN_PI) h -= CN_TWO_PI;
    else if (h < -CN_PI) h += CN_TWO_PI;
    return h;
}

/* ---- C: utils.get_scan_ranges (UTL:375-392) ------------------------------
 * inf -> max; NaN -> 0; 0.0 -> max; > max -> max; reverse; drop the last.
 * raw[R] is in sensor order, out[R-1] in observation order: out[j] = raw[R-1-j].
 */
static void clean_scan(const orc_ctx* c, const float* raw, float* out) {
    const int R = c->cfg.n_samples;
    const float maxr = c->cfg.max_range;
    for (int j = 0; j < R - 1; ++j) {
        float v = raw[(R - 1) - j];
        float r;
        if (isinf(v)) r = maxr;
        else if (isnan(v)) r = 0.0f;
        else if (v == 0.0f) r = maxr;
        else if (v > maxr) r = maxr;
        else r = v;
        out[j] = r;
    }
}

/* ---- L: LiDAR (Gazebo ray sensor configured at XACRO:148-179) -----------
 * sample i at yaw + i * sweep/(R-1), origin at the scan frame (URDF:134-138);
 * nearest hit over the four inner wall faces, then the pedestrian discs in
 * index order (strict <); no return strictly inside max_range -> +inf; returns nearer
 * than the sensor minimum read as the minimum.  Sample 0 is never observed
 * (UTL:390 drops it) and is not cast.  hid is filled in OBSERVATION order.
 */
static void lidar_raw(const orc_ctx* c, const uint32_t* rob, const uint32_t* pa,
                      float* raw /*[R]*/, uint8_t* hid /*[R-1]*/) {
    const cn_config* g = &c->cfg;
    const int R = g->n_samples, N = g->n_peds;
    int32_t xi = (int32_t)rob[CN_R_X], yi = (int32_t)rob[CN_R_Y];
    uint32_t th = rob[CN_R_TH];
    float sy, cy; cn_sincos_bin(th, &sy, &cy);
    float offx = g->mount_x * cy, offy = g->mount_x * sy;     /* sensor - robot */
    float ox = (float)xi * CN_GRID + offx, oy = (float)yi * CN_GRID + offy;
    /* pedestrian centres relative to the sensor */
    float qx[CN_MAX_PEDS], qy[CN_MAX_PEDS]; int cand[CN_MAX_PEDS];
    for (int n = 0; n < N; ++n) {
        int32_t pxi = (int32_t)pa[4 * n + 0], pyi = (int32_t)pa[4 * n + 1];
        qx[n] = (float)(pxi - xi) * CN_GRID - offx;
        qy[n] = (float)(pyi - yi) * CN_GRID - offy;
        /* cheap cull: a disc whose centre is beyond max_range + radius cannot be hit */
        cand[n] = fmaf(qx[n], qx[n], qy[n] * qy[n]) < c->d.cand_d2;
    }
    raw[0] = INFINITY;
    for (int i = 1; i < R; ++i) {
        float s, co; cn_sincos_bin(th + (uint32_t)i * c->d.inc_bin, &s, &co);
        float best = INFINITY; uint8_t id = CN_HIT_NONE;
        /* walls: x faces then y faces */
        if (co != 0.0f) {
            float t = ((co > 0.0f ? g->room_xmax : g->room_xmin) - ox) / co;
            if (t > 0.0f && t < g->max_range && t < best) { best = t; id = CN_HIT_WALL; }
        }
        if (s != 0.0f) {
            float t = ((s > 0.0f ? g->room_ymax : g->room_ymin) - oy) / s;
            if (t > 0.0f && t < g->max_range && t < best) { best = t; id = CN_HIT_WALL; }
        }
        for (int n = 0; n < N; ++n) {
            if (!cand[n]) continue;
            float b = fmaf(qx[n], co, qy[n] * s);
            float h = fmaf(qx[n], s, -(qy[n] * co));
            float disc = fmaf(-h, h, c->d.ped_r2);
            if (disc < 0.0f) continue;
            float sq = sqrtf(disc);
            if (!(b + sq > 0.0f)) continue;           /* disc entirely behind */
            float t = b - sq;
            if (t < 0.0f) t = 0.0f;                   /* sensor inside the disc */
            if (t < g->max_range && t < best) { best = t; id = (uint8_t)n; }
        }
        if (id != CN_HIT_NONE && best < g->sensor_min_range) best = g->sensor_min_range;
        raw[i] = best;
        hid[(R - 1) - i] = id;
    }
}
static void lidar(const orc_ctx* c, const uint32_t* rob, const uint32_t* pa,
                  float* ranges /*[R-1]*/, uint8_t* hid /*[R-1]*/) {
    float raw[4096];
    lidar_raw(c, rob, pa, raw, hid);
    clean_scan(c, raw, ranges);
}

/* C2: utils.convert_laserscan_to_coordinate (UTL:110-126) for observation ray j:
 * from the ROBOT CENTRE (not the scan frame), angle j * inc_deg - yaw with the
 * Python-2 integer-degree increment, y negated, both rounded to 3 dp. */
static void hit_point(const orc_ctx* c, float xf, float yf, uint32_t th, int j, float d, float* hx, float* hy) {
    float sa, ca; cn_sincos_bin((uint32_t)j * c->d.hit_inc_bin - th, &sa, &ca);
    *hx = cn_py_round3(xf + d * ca);
    *hy = cn_py_round3(yf + (d * sa) * -1.0f);
}

/* UTL:317-323 composed with ENV:835: min(1, 0.15 / (dtc / resultant)); None -> 0 */
static float cp_ttc_of(int have_dtc, float dtc, float resultant) {
    if (!have_dtc) return 0.0f;
    float q = (0.15f * resultant) / dtc;
    return (q < 1.0f) ? q : 1.0f;
}

/* UTL:326-345 */
static float cp_dto(const orc_ctx* c, float d) {
    if (d > c->cfg.max_range) return 0.0f;
    return (c->cfg.max_range - d) * c->d.inv_cp_span;
}

typedef struct { float cp, x, y, vx, vy; int n; } risk_obj;

/* cn_oracle_faithful.c: the reference's own perception block (CN_FLAG_RISK_FAITHFUL) */
void orf_observe(const cnf_params* p, uint32_t* trk, double x, double y, double yaw, const double* scan,
                 int step_counter, double* kblock);

/*
 * observe: Env.get_state (ENV:245-1044) for one world, `risk_intended` flavour
 * (SURVEY 8a "deterministic restatement of J/K").  Writes the rounded
 * observation row and returns this step's done conditions (ENV:1011-1023).
 */
static int observe(const orc_ctx* c, uint32_t* rob, const uint32_t* pa, uint32_t* pb, uint32_t* trk,
                   int step_counter, int have_prev_pose, float* obs,
                   float* dbg_ranges, uint8_t* dbg_hid) {
    const cn_config* g = &c->cfg;
    const int R = g->n_samples, N = g->n_peds, K = g->k_obstacles;
    const int NR = R - 1;
    float xf = (float)(int32_t)rob[CN_R_X] * CN_GRID;
    float yf = (float)(int32_t)rob[CN_R_Y] * CN_GRID;
    uint32_t th = rob[CN_R_TH];
    float yaw = cn_bin2rad(th);
    float wx = f_of(rob[CN_R_WPX]), wy = f_of(rob[CN_R_WPY]);

    const int original = (g->flags & CN_FLAG_ENV_ORIGINAL) != 0;
    float dist, head;
    if (original) {
        /* environment_stage_1_original.py:280-281: distance / heading to the goal itself, no waypoints
         * (desired_point; the heading has no starting_pose term there: the config carries a zero offset) */
        dist = cn_py_round2(dist_to_wp(xf, yf, g->goal_x, g->goal_y));
        head = cn_py_round2(heading_to_wp(c, xf, yf, yaw, g->goal_x, g->goal_y));
    } else {
        /* A: ENV:246-265 */
        if (step_counter == 1) waypoint(c, xf, yf, &wx, &wy);
        dist = cn_py_round2(dist_to_wp(xf, yf, wx, wy));
        head = cn_py_round2(heading_to_wp(c, xf, yf, yaw, wx, wy));
        if (step_counter % 5 == 0 || dist < f_of(rob[CN_R_PDIST])) waypoint(c, xf, yf, &wx, &wy);
        rob[CN_R_WPX] = u_of(wx); rob[CN_R_WPY] = u_of(wy);
    }

    /* B: ENV:267-268 -- yaw RATE used as an angle (sic) */
    float v = f_of(rob[CN_R_V]), w = f_of(rob[CN_R_W]);
    float sw, cw; cn_sincos_rad(w, &sw, &cw);
    float avx = -1.0f * (v * cw), avy = v * sw;

    /* L + C */
    float* ranges = obs;                      /* first NR slots, rounded at the end */
    uint8_t hid_buf[4096]; uint8_t* hid = dbg_hid ? dbg_hid : hid_buf;
    lidar(c, rob, pa, ranges, hid);
    if (dbg_ranges) memcpy(dbg_ranges, ranges, sizeof(float) * NR);
    float min_scan = INFINITY;
    for (int j = 0; j < NR; ++j) if (ranges[j] < min_scan) min_scan = ranges[j];

    /* CN_FLAG_RISK_FAITHFUL: rows E-K and M as the reference computes them (float64, its own segmentation and
     * tracker) from the same odometry and cleaned scan; the block below still runs (its state stays in the blob)
     * but its K block is replaced at the end. */
    double kfaith[4 * CN_MAX_PEDS];
    if (trk) {
        cnf_params fp; cnf_params_from_config(g, &fp);
        double scan64[1024];
        for (int j = 0; j < NR; ++j) scan64[j] = (ranges[j] >= g->max_range) ? fp.max_range : (double)ranges[j];
        orf_observe(&fp, trk, (double)xf, (double)yf, (double)yaw, scan64, step_counter, kfaith);
    }

    /* E-I collapsed by ideal association: an object = a pedestrian hit by >= 4
     * rays (ENV:573); its point = the hit ray nearest its centre line. */
    float p_cur_x = cn_py_round3(xf), p_cur_y = cn_py_round3(yf);   /* ENV:1208 */
    float p_prev_x = f_of(rob[CN_R_PPX]), p_prev_y = f_of(rob[CN_R_PPY]);
    float agent_vel = 0.0f;
    if (have_prev_pose) {                      /* UTL:227-236 */
        float vx = (p_cur_x - p_prev_x) * c->d.inv_dt, vy = (p_cur_y - p_prev_y) * c->d.inv_dt;
        agent_vel = sqrtf(fmaf(vx, vx, vy * vy));
    }
    float sy_, cy_; cn_sincos_bin(th, &sy_, &cy_);
    float offx = g->mount_x * cy_, offy = g->mount_x * sy_;
    risk_obj objs[CN_MAX_PEDS]; int n_obj = 0;
    float ego_score = 0.0f; int ego_violation = 0;
    for (int n = 0; n < N; ++n) {
        int cnt = 0;
        for (int j = 0; j < NR; ++j) cnt += (hid[j] == n);
        uint32_t* pbn = pb + 4 * n;
        if (cnt < 4) { pbn[3] &= ~CN_PF_TRACKED; continue; }
        /* centre ray: min |ray angle - bearing of the pedestrian centre| */
        float qx = (float)((int32_t)pa[4 * n] - (int32_t)rob[CN_R_X]) * CN_GRID - offx;
        float qy = (float)((int32_t)pa[4 * n + 1] - (int32_t)rob[CN_R_Y]) * CN_GRID - offy;
        uint32_t bearing = cn_rad2bin(cn_atan2(qy, qx));
        uint32_t best_key = 0xFFFFFFFFu; int jstar = -1;
        for (int j = 0; j < NR; ++j) {
            if (hid[j] != n) continue;
            int i = (R - 1) - j;
            int32_t delta = (int32_t)(th + (uint32_t)i * c->d.inc_bin - bearing);
            uint32_t ad = (delta < 0) ? (uint32_t)(-(int64_t)delta) : (uint32_t)delta;
            uint32_t key = (ad & ~1u) | (delta < 0 ? 1u : 0u);
            if (key < best_key || (key == best_key && j < jstar)) { best_key = key; jstar = j; }
        }
        float d_raw = ranges[jstar];
        float d3 = cn_py_round3(d_raw);                    /* ENV:324,384 */
        /* C2: utils.convert_laserscan_to_coordinate (UTL:110-126), from the
         * robot centre with the integer-degree increment */
        float hx, hy; hit_point(c, xf, yf, th, jstar, d_raw, &hx, &hy);
        /* H/I: tracker with ideal association (ENV:656-760) */
        float chx = 0.0f, chy = 0.0f, speed = -1.0f, ovx = 0.0f, ovy = 0.0f;
        if (pbn[3] & CN_PF_TRACKED) {
            chx = f_of(pbn[0]) - hx;                      /* last - curr (sic), ENV:806-807 */
            chy = f_of(pbn[1]) - hy;
            speed = sqrtf(fmaf(chy, chy, chx * chx)) * c->d.inv_dt;   /* ENV:754-757 */
            ovx = chx * c->d.inv_dt; ovy = chy * c->d.inv_dt;   /* ENV:808-809 */
        }
        pbn[0] = u_of(hx); pbn[1] = u_of(hy); pbn[3] |= CN_PF_TRACKED;
        if (d3 < 0.140f) ego_violation = 1;               /* ENV:1000 */
        if (!have_prev_pose) continue;                     /* ENV:769 */
        /* J: collision cone (ENV:765-860, UTL:251-293 as a true ray-circle test) */
        float tx = p_cur_x + chx, ty = p_cur_y + chy;     /* ENV:814-815 */
        float ux = tx - p_prev_x, uy = ty - p_prev_y;
        float L = sqrtf(fmaf(ux, ux, uy * uy));
        int have_dtc = 0; float dtc = 0.0f;
        if (L > 0.0f) {
            float invL = 1.0f / L;
            ux = ux * invL; uy = uy * invL;
            float wx_ = hx - p_prev_x, wy_ = hy - p_prev_y;
            float b = fmaf(wx_, ux, wy_ * uy);
            float h = fmaf(wx_, uy, -(wy_ * ux));
            float disc = fmaf(-h, h, c->d.cp_r2);
            if (disc > 0.0f) {
                float t = b - sqrtf(disc);
                if (t > 0.0f) { have_dtc = 1; dtc = t; }
            }
        }
        float resultant = agent_vel - speed;              /* ENV:825 */
        float cp_ttc = 0.0f, cp;
        float dto = cp_dto(c, d3);
        if (have_dtc && resultant == 0.0f) {
            cp = dto;                                     /* ENV:828-833 */
        } else {
            cp_ttc = cp_ttc_of(have_dtc, dtc, resultant);
            cp = 0.5f * cp_ttc + 0.5f * dto;              /* ENV:838,851 */
        }
        if (n_obj == 0 || cp_ttc > ego_score) ego_score = cp_ttc;   /* ENV:879 */
        objs[n_obj].cp = cp; objs[n_obj].x = hx; objs[n_obj].y = hy;
        objs[n_obj].vx = ovx; objs[n_obj].vy = ovy; objs[n_obj].n = n; ++n_obj;
    }
    /* count objects seen this step (also at step 0, where J is skipped) */
    int n_seen = 0;
    for (int n = 0; n < N; ++n) n_seen += (pb[4 * n + 3] & CN_PF_TRACKED) ? 1 : 0;

    /* K: ENV:862-907.  stable sort by CP descending, keep [-K:], pad */
    float* blk = obs + NR + 7;
    for (int s = 0; s < K; ++s) { blk[4 * s] = xf; blk[4 * s + 1] = yf; blk[4 * s + 2] = 0.0f; blk[4 * s + 3] = 0.0f; }
    for (int a = 0; a < (K > 0 ? n_obj : 0); ++a) {
        int rank = 0;
        for (int b = 0; b < n_obj; ++b) {
            if (b == a) continue;
            if (objs[b].cp > objs[a].cp || (objs[b].cp == objs[a].cp && b < a)) ++rank;
        }
        int slot;
        if (g->flags & CN_FLAG_TOPK_HIGHEST) slot = rank;
        else slot = rank - (n_obj > K ? n_obj - K : 0);
        if (slot < 0 || slot >= K) continue;
        blk[4 * slot] = objs[a].x; blk[4 * slot + 1] = objs[a].y;
        blk[4 * slot + 2] = objs[a].vx; blk[4 * slot + 3] = objs[a].vy;
    }

    /* M: ENV:653-654, 998-1005 */
    uint32_t ego = rob[CN_R_CNT0] & 0xFFFFu, soc = rob[CN_R_CNT0] >> 16;
    uint32_t pres = rob[CN_R_CNT1] & 0xFFFFu, bad = rob[CN_R_CNT1] >> 16;
    if (n_seen > 0 && pres < 0xFFFFu) ++pres;
    if (ego_violation && ego < 0xFFFFu) ++ego;
    if (ego_score > 0.4f && soc < 0xFFFFu) ++soc;
    rob[CN_R_CNT0] = ego | (soc << 16);
    rob[CN_R_CNT1] = pres | (bad << 16);

    /* N: ENV:1011-1023 */
    int done = 0;
    if (min_scan < g->collision_range) done = 1;
    if (in_goal_box(c, xf, yf)) done = 1;
    if (step_counter >= g->max_steps) done = 1;

    if (original) {
        /* original:309-318: [round(range, 3) ... | heading, distance | round(x, 3), round(y, 3)], Python rounding */
        for (int j = 0; j < NR; ++j) obs[j] = cn_py_round3(obs[j]);
        obs[NR + 0] = head; obs[NR + 1] = dist;
        obs[NR + 2] = p_cur_x; obs[NR + 3] = p_cur_y;
    } else {
        /* O: ENV:1025-1042 */
        obs[NR + 0] = head; obs[NR + 1] = dist;
        obs[NR + 2] = p_cur_x; obs[NR + 3] = p_cur_y;
        obs[NR + 4] = cn_py_round3(yaw);
        obs[NR + 5] = cn_py_round3(avx); obs[NR + 6] = cn_py_round3(avy);
        for (int k = 0; k < c->d.obs_dim; ++k) obs[k] = cn_np_round3(obs[k]);
        if (trk) for (int k = 0; k < 4 * K; ++k) obs[NR + 7 + k] = (float)kfaith[k];
    }

    /* ENV:1208 / 991-992: the deque keeps the current rounded pose */
    rob[CN_R_PPX] = u_of(p_cur_x); rob[CN_R_PPY] = u_of(p_cur_y);
    return done;
}

/* W (non-terminal part): step -2, distance +1 iff decreased, heading truth table (ENV:1050-1106) */
static int shaping_reward(float cur_head, float cur_dist, float prev_head, float prev_dist) {
    float dd = cur_dist - prev_dist, dh = cur_head - prev_head;
    int reward = -2;
    if (dd < 0.0f) reward += 1;                            /* ENV:1076-1078 */
    int htg = 0;                                           /* ENV:1081-1106 */
    if (dh > 0.0f) {
        if (cur_head > 0.0f && prev_head < 0.0f) htg = 1;
        if (cur_head < 0.0f && prev_head < 0.0f) htg = 1;
        if (cur_head < 0.0f && prev_head > 0.0f) htg = 1;
        if (cur_head > 0.0f && prev_head > 0.0f) htg = 0;
    }
    if (dh < 0.0f) {
        if (cur_head < 0.0f && prev_head > 0.0f) htg = 1;
        if (cur_head > 0.0f && prev_head > 0.0f) htg = 1;
        if (cur_head > 0.0f && prev_head < 0.0f) htg = 1;
        if (cur_head < 0.0f && prev_head < 0.0f) htg = 0;
    }
    return reward + htg;
}

/* ---- Z: Env.reset (ENV:1227-1263) + gazebo/reset_simulation -------------- */
static void reset_env(const orc_ctx* c, int e, uint32_t* rob, uint32_t* pa, uint32_t* pb, uint32_t* trk, float* obs,
                      float* dbg_ranges, uint8_t* dbg_hid) {
    const cn_config* g = &c->cfg;
    const int N = g->n_peds;
    uint32_t gid = (uint32_t)(g->env_id_offset + e);
    uint32_t episode = rob[CN_R_EPISODE] + 1u;
    int b = (int)(gid % (uint32_t)g->n_behaviors);
    memset(rob, 0, sizeof(uint32_t) * CN_ROBOT_WORDS);
    rob[CN_R_EPISODE] = episode;
    rob[CN_R_X] = (uint32_t)c->d.start_xi; rob[CN_R_Y] = (uint32_t)c->d.start_yi;
    rob[CN_R_TH] = c->d.start_th;
    for (int n = 0; n < N; ++n) {
        cn_u32x2 r = cn_env_rand(c->d.seed_lo, c->d.seed_hi, gid, episode, 0u, (uint32_t)n, 1u);
        float px = g->ped_layout[n][0] + cn_usym(r.v[0], g->layout_jitter);
        float py = g->ped_layout[n][1] + cn_usym(r.v[1], g->layout_jitter);
        int32_t pxi = cn_f2i(px * CN_INV_GRID), pyi = cn_f2i(py * CN_INV_GRID);
        if (pxi < c->d.ped_xmin) pxi = c->d.ped_xmin;
        if (pxi > c->d.ped_xmax) pxi = c->d.ped_xmax;
        if (pyi < c->d.ped_ymin) pyi = c->d.ped_ymin;
        if (pyi > c->d.ped_ymax) pyi = c->d.ped_ymax;
        pa[4 * n + 0] = (uint32_t)pxi; pa[4 * n + 1] = (uint32_t)pyi;
        pa[4 * n + 2] = u_of(0.0f); pa[4 * n + 3] = u_of(0.0f);
        pb[4 * n + 0] = 0; pb[4 * n + 1] = 0;
        /* first velocity command reaches pedestrian n after (n+1) staggers (CROWD:128-144) */
        pb[4 * n + 2] = (uint32_t)(int32_t)((n + 1) * g->behavior_stagger_ticks[b]);
        pb[4 * n + 3] = 0;
    }
    float xf = (float)c->d.start_xi * CN_GRID, yf = (float)c->d.start_yi * CN_GRID;
    float wx = g->goal_x, wy = g->goal_y;                 /* ENV:80-83 */
    rob[CN_R_WPX] = u_of(wx); rob[CN_R_WPY] = u_of(wy);
    /* ENV:1243-1244: unrounded */
    rob[CN_R_PDIST] = u_of(dist_to_wp(xf, yf, wx, wy));
    rob[CN_R_PHEAD] = u_of(heading_to_wp(c, xf, yf, cn_bin2rad(c->d.start_th), wx, wy));
    (void)observe(c, rob, pa, pb, trk, 0, 0, obs, dbg_ranges, dbg_hid);   /* ENV:1246 */
    rob[CN_R_CNT0] = 0; rob[CN_R_CNT1] = 0;               /* ENV:1260-1262 */
}

/* contact stand-in for ODE: short-range repulsion, exactly 0 beyond the cutoff */
static void add_rep(const orc_ctx* c, int32_t xi, int32_t yi, int32_t xj, int32_t yj,
                    float rsum, float* vex, float* vey) {
    float dx = (float)(xi - xj) * CN_GRID, dy = (float)(yi - yj) * CN_GRID;
    float d2 = fmaf(dx, dx, dy * dy);
    float lim = rsum + c->cfg.rep_cutoff;
    if (d2 < lim * lim && d2 > 0.0f) {
        float d = sqrtf(d2);
        float f = (c->cfg.rep_strength * cn_exp((rsum - d) / c->cfg.rep_range)) / d;
        *vex += f * dx;
        *vey += f * dy;
    }
}

/* ---- S: Env.step (ENV:1164-1225) for one world --------------------------- */
static void step_env(const orc_ctx* c, int e, uint32_t* rob, uint32_t* pa, uint32_t* pb, uint32_t* trk,
                     const float* action, float* obs, float* reward_out, uint8_t* done_out,
                     float* dbg_ranges, uint8_t* dbg_hid) {
    const cn_config* g = &c->cfg;
    const int N = g->n_peds;
    uint32_t gid = (uint32_t)(g->env_id_offset + e);
    int b = (int)(gid % (uint32_t)g->n_behaviors);
    int step_counter = (int)rob[CN_R_STEP] + 1;
    uint32_t episode = rob[CN_R_EPISODE];

    /* Auto-reset ("next-step" mode): a world whose episode ended on the previous step spends this step
     * restarting -- the action is ignored, the row is the first observation of the new episode, reward 0,
     * done = 2 (a transition the trainer must not store).  Every world does exactly one get_state per step. */
    if ((rob[CN_R_FLAGS] & CN_RF_DONE) && (g->flags & CN_FLAG_AUTO_RESET)) {
        uint32_t keep = rob[CN_R_FLAGS] & (CN_RF_SUCCESS | CN_RF_FAILURE);
        reset_env(c, e, rob, pa, pb, trk, obs, dbg_ranges, dbg_hid);
        rob[CN_R_FLAGS] |= keep;   /* last episode's status stays readable (ENV:1265-1267) */
        *reward_out = 0.0f;
        *done_out = 2;
        return;
    }

    /* T2: action, applied verbatim (ENV:1190-1192) after sanitising */
    float av = action[0], aw = action[1];
    if (!(fabsf(av) <= 3.0e38f) || !(fabsf(aw) <= 3.0e38f)) {
        av = 0.0f; aw = 0.0f;
        uint32_t bad = rob[CN_R_CNT1] >> 16;
        if (bad < 0xFFFFu) ++bad;
        rob[CN_R_CNT1] = (rob[CN_R_CNT1] & 0xFFFFu) | (bad << 16);
    }
    if (av > CN_ACT_V_LIMIT) av = CN_ACT_V_LIMIT;
    if (av < -CN_ACT_V_LIMIT) av = -CN_ACT_V_LIMIT;
    if (aw > CN_ACT_W_LIMIT) aw = CN_ACT_W_LIMIT;
    if (aw < -CN_ACT_W_LIMIT) aw = -CN_ACT_W_LIMIT;

    /* P: pedestrians (CROWD:98-144 + contact stand-in), Jacobi on old positions */
    int32_t rxi = (int32_t)rob[CN_R_X], ryi = (int32_t)rob[CN_R_Y];
    int32_t ox[CN_MAX_PEDS], oy[CN_MAX_PEDS];
    for (int n = 0; n < N; ++n) { ox[n] = (int32_t)pa[4 * n]; oy[n] = (int32_t)pa[4 * n + 1]; }
    for (int n = 0; n < N; ++n) {
        float vx = f_of(pa[4 * n + 2]), vy = f_of(pa[4 * n + 3]);
        int32_t timer = (int32_t)pb[4 * n + 2] - CN_TICKS_PER_STEP;
        if (timer <= 0) {
            if (g->behavior_kind[b] == CN_BEHAVIOR_RANDOM) {
                cn_u32x2 r = cn_env_rand(c->d.seed_lo, c->d.seed_hi, gid, episode, (uint32_t)step_counter,
                                         (uint32_t)n, 0u);
                vx = cn_usym(r.v[0], g->behavior_speed[b]);
                vy = cn_usym(r.v[1], g->behavior_speed[b]);
            } else {
                vx = g->behavior_table[b][n][0] * g->behavior_speed[b];
                vy = g->behavior_table[b][n][1] * g->behavior_speed[b];
            }
            timer += g->behavior_period_ticks[b];
        }
        float vex = vx, vey = vy;
        for (int m = 0; m < N; ++m)
            if (m != n) add_rep(c, ox[n], oy[n], ox[m], oy[m], g->ped_radius + g->ped_radius, &vex, &vey);
        add_rep(c, ox[n], oy[n], rxi, ryi, g->ped_radius + g->robot_radius, &vex, &vey);
        int32_t nx = ox[n] + cn_f2i((vex * g->dt) * CN_INV_GRID);
        int32_t ny = oy[n] + cn_f2i((vey * g->dt) * CN_INV_GRID);
        /* frictionless wall contact: clamp the centre, drop the inward normal velocity */
        if (nx < c->d.ped_xmin) { nx = c->d.ped_xmin; if (vx < 0.0f) vx = 0.0f; }
        if (nx > c->d.ped_xmax) { nx = c->d.ped_xmax; if (vx > 0.0f) vx = 0.0f; }
        if (ny < c->d.ped_ymin) { ny = c->d.ped_ymin; if (vy < 0.0f) vy = 0.0f; }
        if (ny > c->d.ped_ymax) { ny = c->d.ped_ymax; if (vy > 0.0f) vy = 0.0f; }
        pa[4 * n] = (uint32_t)nx; pa[4 * n + 1] = (uint32_t)ny;
        pa[4 * n + 2] = u_of(vx); pa[4 * n + 3] = u_of(vy);
        pb[4 * n + 2] = (uint32_t)timer;
    }

    /* R: unicycle, midpoint rule (FAKE:109-118, 156-167), in n_substeps sub-steps; with wheel_accel > 0 the wheel
     * speeds ramp toward their targets like libgazebo_ros_diff_drive (XACRO:65,70) instead of jumping */
    {
        float half = (aw * CN_WHEEL_SEP) * 0.5f;
        float tl = av - half, tr = av + half;                 /* FAKE:116-117: wheel speed targets */
        float cl = tl, cr = tr;
        if (c->d.wheel_step > 0.0f) {                         /* current wheel speeds from the achieved body twist */
            float cv = f_of(rob[CN_R_V]), cw = f_of(rob[CN_R_W]);
            float ch = (cw * CN_WHEEL_SEP) * 0.5f;
            cl = cv - ch; cr = cv + ch;
        }
        int32_t xi = rxi, yi = ryi;
        uint32_t th = rob[CN_R_TH];
        float v_body = 0.0f, w_body = 0.0f;
        for (int k = 0; k < g->n_substeps; ++k) {
            if (c->d.wheel_step > 0.0f) {
                float st = c->d.wheel_step;
                cl += fminf(fmaxf(tl - cl, -st), st);
                cr += fminf(fmaxf(tr - cr, -st), st);
            }
            v_body = (cr + cl) * 0.5f;                        /* FAKE:156,165: delta_s / dt     */
            w_body = (cr - cl) * CN_INV_WHEEL_SEP;            /* FAKE:157,167: delta_theta / dt */
            float ds = v_body * c->d.dt_sub;
            float dth = w_body * c->d.dt_sub;
            int32_t dth_bin = cn_f2i(dth * CN_RAD2BIN);
            uint32_t mid = th + (uint32_t)(dth_bin >> 1);
            float sm, cm; cn_sincos_bin(mid, &sm, &cm);
            int32_t nx = xi + cn_f2i((ds * cm) * CN_INV_GRID);
            int32_t ny = yi + cn_f2i((ds * sm) * CN_INV_GRID);
            if (c->d.robot_contact) {
                /* Body contact, only when the LiDAR threshold cannot end the episode first (collision_range <
                 * robot_radius: README.md:60-62 `min_scan_range 0.0`).  Gazebo blocks the body; here the centre stays
                 * robot_radius off the wall faces and a sub-step INTO a pedestrian's disc (old positions) is dropped. */
                if (nx < c->d.rob_xmin) nx = c->d.rob_xmin;
                if (nx > c->d.rob_xmax) nx = c->d.rob_xmax;
                if (ny < c->d.rob_ymin) ny = c->d.rob_ymin;
                if (ny > c->d.rob_ymax) ny = c->d.rob_ymax;
                int blocked = 0;
                for (int n = 0; n < N; ++n) {
                    float dxn = (float)(nx - ox[n]) * CN_GRID, dyn = (float)(ny - oy[n]) * CN_GRID;
                    float d2n = fmaf(dxn, dxn, dyn * dyn);
                    if (d2n < c->d.rob_ped_r2) {
                        float dxo = (float)(xi - ox[n]) * CN_GRID, dyo = (float)(yi - oy[n]) * CN_GRID;
                        if (d2n < fmaf(dxo, dxo, dyo * dyo)) blocked = 1;
                    }
                }
                if (blocked) { nx = xi; ny = yi; }
            }
            xi = nx; yi = ny;
            th += (uint32_t)dth_bin;
        }
        rob[CN_R_X] = (uint32_t)xi; rob[CN_R_Y] = (uint32_t)yi; rob[CN_R_TH] = th;
        rob[CN_R_V] = u_of(v_body);
        rob[CN_R_W] = u_of(w_body);
    }

    /* get_state */
    int done_now = observe(c, rob, pa, pb, trk, step_counter, 1, obs, dbg_ranges, dbg_hid);
    int done = ((rob[CN_R_FLAGS] & CN_RF_DONE) != 0) || done_now;   /* sticky, ENV:1011 */

    /* W: compute_reward (ENV:1046-1162) on the ROUNDED heading / distance */
    const int NR = g->n_samples - 1;
    const int original = (g->flags & CN_FLAG_ENV_ORIGINAL) != 0;
    float cur_head = obs[NR + 0], cur_dist = obs[NR + 1];
    float prev_head = f_of(rob[CN_R_PHEAD]), prev_dist = f_of(rob[CN_R_PDIST]);
    int reward = shaping_reward(cur_head, cur_dist, prev_head, prev_dist);
    float xf = (float)(int32_t)rob[CN_R_X] * CN_GRID, yf = (float)(int32_t)rob[CN_R_Y] * CN_GRID;
    float wx = f_of(rob[CN_R_WPX]), wy = f_of(rob[CN_R_WPY]);
    if (original) {
        /* original:324-326: `current_distance = state[-1]`, `current_heading = state[-2]` -- with the 363-wide row
         * these are the robot's y and x (sic); step_reward = 0 (original:334), no waypoint bonus */
        cur_head = obs[NR + 2]; cur_dist = obs[NR + 3];
        reward = shaping_reward(cur_head, cur_dist, prev_head, prev_dist) + 2;
    } else
    if (in_box(xf, yf, wx - g->goal_box, wx + g->goal_box, wy - g->goal_box, wy + g->goal_box)) {
        waypoint(c, xf, yf, &wx, &wy);                     /* ENV:1109-1116 */
        reward += 200;
        if (in_goal_box(c, wx, wy)) { wx = g->goal_x; wy = g->goal_y; }   /* ENV:1121-1123 */
        rob[CN_R_WPX] = u_of(wx); rob[CN_R_WPY] = u_of(wy);
    }
    rob[CN_R_PDIST] = u_of(cur_dist); rob[CN_R_PHEAD] = u_of(cur_head);   /* ENV:1133-1134 */
    uint32_t flags = rob[CN_R_FLAGS];
    if (done) {                                            /* ENV:1136-1159; timeout is -200 too */
        flags |= CN_RF_DONE;
        if (in_goal_box(c, xf, yf)) { flags |= CN_RF_SUCCESS; flags &= ~CN_RF_FAILURE; reward += 200; }
        else { flags |= CN_RF_FAILURE; flags &= ~CN_RF_SUCCESS; reward -= 200; }
    }
    rob[CN_R_FLAGS] = flags;
    rob[CN_R_STEP] = (uint32_t)step_counter;
    *reward_out = (float)reward;
    *done_out = (uint8_t)(done ? 1 : 0);
}

/* ---- batch entry points -------------------------------------------------- */
#ifdef _OPENMP
#include <omp.h>
#endif
/* threads of the batch loops below (bench.py: the same sample on all host cores and on ONE core); <= 0: leave as is */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void orc_reset(const orc_ctx* c, uint32_t* blob, const uint8_t* mask, float* obs,
               float* dbg_ranges, uint8_t* dbg_hid) {
    const int E = c->cfg.n_envs, N = c->cfg.n_peds, D = c->d.obs_dim, NR = c->cfg.n_samples - 1;
    uint32_t* rob = blob_robot(c, blob); uint32_t* pa = blob_ped_a(c, blob); uint32_t* pb = blob_ped_b(c, blob);
    uint32_t* trk = blob_trk(c, blob);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < E; ++e) {
        if (mask && !mask[e]) continue;
        reset_env(c, e, rob + (size_t)e * CN_ROBOT_WORDS, pa + (size_t)e * N * 4, pb + (size_t)e * N * 4,
                  trk ? trk + (size_t)e * CNF_WORLD_WORDS : NULL, obs + (size_t)e * D,
                  dbg_ranges ? dbg_ranges + (size_t)e * NR : NULL, dbg_hid ? dbg_hid + (size_t)e * NR : NULL);
        rob[(size_t)e * CN_ROBOT_WORDS + CN_R_FLAGS] = 0;
    }
}

void orc_step(const orc_ctx* c, uint32_t* blob, const float* actions, float* obs, float* reward,
              uint8_t* done, float* dbg_ranges, uint8_t* dbg_hid) {
    const int E = c->cfg.n_envs, N = c->cfg.n_peds, D = c->d.obs_dim, NR = c->cfg.n_samples - 1;
    uint32_t* rob = blob_robot(c, blob); uint32_t* pa = blob_ped_a(c, blob); uint32_t* pb = blob_ped_b(c, blob);
    uint32_t* trk = blob_trk(c, blob);
#pragma omp parallel for schedule(static)
    for (int e = 0; e < E; ++e) {
        step_env(c, e, rob + (size_t)e * CN_ROBOT_WORDS, pa + (size_t)e * N * 4, pb + (size_t)e * N * 4,
                 trk ? trk + (size_t)e * CNF_WORLD_WORDS : NULL, actions + 2 * (size_t)e, obs + (size_t)e * D, reward + e, done + e,
                 dbg_ranges ? dbg_ranges + (size_t)e * NR : NULL, dbg_hid ? dbg_hid + (size_t)e * NR : NULL);
    }
}

void orc_clear_done(const orc_ctx* c, uint32_t* blob, const uint8_t* mask) {
    uint32_t* rob = blob_robot(c, blob);
    for (int e = 0; e < c->cfg.n_envs; ++e)
        if (!mask || mask[e]) rob[(size_t)e * CN_ROBOT_WORDS + CN_R_FLAGS] &= ~CN_RF_DONE;
}

/* [E, 4] int32: success, ego violations, social violations, obstacle-present steps */
void orc_counters(const orc_ctx* c, const uint32_t* blob, int32_t* out) {
    const uint32_t* rob = blob + CN_BLOB_HEADER_WORDS;
    const uint32_t* trk = blob_trk(c, (uint32_t*)blob);
    for (int e = 0; e < c->cfg.n_envs; ++e) {
        const uint32_t* r = rob + (size_t)e * CN_ROBOT_WORDS;
        out[4 * e + 0] = (r[CN_R_FLAGS] & CN_RF_SUCCESS) ? 1 : 0;
        if (trk) {                                          /* CN_FLAG_RISK_FAITHFUL: the tracker record's counters */
            const uint32_t* t = trk + (size_t)e * CNF_WORLD_WORDS;
            out[4 * e + 1] = (int32_t)t[CNF_H_EGO]; out[4 * e + 2] = (int32_t)t[CNF_H_SOCIAL]; out[4 * e + 3] = (int32_t)t[CNF_H_PRESENT];
            continue;
        }
        out[4 * e + 1] = (int32_t)(r[CN_R_CNT0] & 0xFFFFu);
        out[4 * e + 2] = (int32_t)(r[CN_R_CNT0] >> 16);
        out[4 * e + 3] = (int32_t)(r[CN_R_CNT1] & 0xFFFFu);
    }
}

/* ---- primitive taps for tests/test_math_primitives.py --------------------- */
void orc_sincos_bin(const uint32_t* a, float* s, float* c, int n) { for (int i = 0; i < n; ++i) cn_sincos_bin(a[i], s + i, c + i); }
void orc_sincos_rad(const float* x, float* s, float* c, int n) { for (int i = 0; i < n; ++i) cn_sincos_rad(x[i], s + i, c + i); }
void orc_atan2(const float* y, const float* x, float* o, int n) { for (int i = 0; i < n; ++i) o[i] = cn_atan2(y[i], x[i]); }
void orc_exp(const float* x, float* o, int n) { for (int i = 0; i < n; ++i) o[i] = cn_exp(x[i]); }
void orc_round(const float* x, float* np3, float* py3, float* py2, int n) {
    for (int i = 0; i < n; ++i) { np3[i] = cn_np_round3(x[i]); py3[i] = cn_py_round3(x[i]); py2[i] = cn_py_round2(x[i]); }
}
void orc_philox2x32(uint32_t c0, uint32_t c1, uint32_t k0, uint32_t* out) {
    cn_u32x2 r = cn_philox2x32(c0, c1, k0);
    memcpy(out, r.v, 8);
}
/* exhaustive check helper: count k in [lo, hi] where the Markstein forms differ from IEEE division */
long orc_div_const_mismatches(int lo, int hi) {
    long bad = 0;
    for (int k = lo; k <= hi; ++k) {
        float f = (float)k;
        if (cn_div1000(f) != f / 1000.0f) ++bad;
        if (cn_div100(f) != f / 100.0f) ++bad;
    }
    return bad;
}
void orc_clean_scan(const orc_ctx* c, const float* raw, float* out) { clean_scan(c, raw, out); }
void orc_hit_points(const orc_ctx* c, float x, float y, uint32_t th, const float* scans, float* out_xy) {
    for (int j = 0; j < c->cfg.n_samples - 1; ++j) hit_point(c, x, y, th, j, scans[j], out_xy + 2 * j, out_xy + 2 * j + 1);
}
float orc_cp_ttc(int have_dtc, float dtc, float resultant) { return cp_ttc_of(have_dtc, dtc, resultant); }
float orc_cp_dto(const orc_ctx* c, float d) { return cp_dto(c, d); }
int orc_shaping_reward(float ch, float cd, float ph, float pd) { return shaping_reward(ch, cd, ph, pd); }
int orc_in_goal_box(const orc_ctx* c, float x, float y) { return in_goal_box(c, x, y); }
float orc_distance(float x, float y, float wx, float wy) { return dist_to_wp(x, y, wx, wy); }
void orc_waypoint(const orc_ctx* c, float x, float y, float* wx, float* wy) { waypoint(c, x, y, wx, wy); }
float orc_heading(const orc_ctx* c, float x, float y, float yaw, float wx, float wy) { return heading_to_wp(c, x, y, yaw, wx, wy); }
