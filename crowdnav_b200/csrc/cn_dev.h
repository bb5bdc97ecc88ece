/*
 * cn_dev.h -- device-side helpers shared by the env-step kernels (cn_step.cu: warp-per-world,
 * cn_flat.cu: compacted work lists).  PTX wrappers for mbarrier / bulk TMA, the scalar formulas of
 * Env.get_state / compute_reward for ONE world (same operation sequences as oracle/cn_oracle.c; the
 * reference lines are cited there and at each function), span construction for the LiDAR rasteriser.
 */
#ifndef CN_DEV_H
#define CN_DEV_H
#include <cuda_runtime.h>
#include <stdint.h>
#include "cn_state.h"
#include "cn_kernel.h"

#define FULL 0xFFFFFFFFu

#ifdef CN_TIMELINE
static __device__ unsigned long long* g_timeline;      // [n_warps][8] globaltimer stamps (debug builds only)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define STAMP(k) do { if ((threadIdx.x & 31) == 0 && g_timeline) g_timeline[((size_t)blockIdx.x * (CN_TILE + CN_POSE_WARPS) + (threadIdx.x >> 5)) * 16 + (k)] = gtime(); } while (0)
#else
#define STAMP(k) do { } while (0)
#endif

namespace {

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, 1-D), completion on an mbarrier
__device__ __forceinline__ void tma_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk copy (TMA, 1-D)
__device__ __forceinline__ void tma_store(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_commit_and_wait() {
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------ per-world scalar record
// Written by phase A (or, on the auto-reset path, by the world's own warp),
// read by the world's warp.  One float4-aligned 128-byte record per world.
enum {
    S_XI = 0, S_YI, S_TH, S_V, S_W,        // new pose / body twist (bit patterns)
    S_WPX, S_WPY,                          // waypoint after get_state + compute_reward
    S_HEAD, S_DIST,                        // rounded heading / distance (= new previous_*)
    S_PCX, S_PCY,                          // round(pose, 3) (= new agent_pose_deque[0])
    S_PPX, S_PPY,                          // previous rounded pose (collision cone)
    S_XF, S_YF, S_OFFX, S_OFFY,            // float pose, sensor offset
    S_AVEL,                                // agent speed from the rounded poses
    S_REWARD,                              // int: shaping + waypoint bonus (terminal part added later)
    S_PRE,                                 // bit0 in goal box, bit1 timeout
    S_BAD,                                 // int: 1 if the action was sanitised
    S_WDIRTY,                              // chunks of the row the wall spans touch
    S_NPDIST, S_NPHEAD,                    // next previous_distance / previous_heading (unrounded after a reset)
    S_WSPAN = 24,                          // 4 wall faces x (a0, a1, b0, b1): rays that can see the face
    S_WORDS = 40
};

__device__ __forceinline__ float f_of(uint32_t u) { return __uint_as_float(u); }
__device__ __forceinline__ uint32_t u_of(float f) { return __float_as_uint(f); }

// ------------------------------------------------------------ scalar formulas
// (same operation sequences as oracle/cn_oracle.c; see the citations there)

__device__ __forceinline__ void waypoint(const cn_kparams& P, float xf, float yf, float& wx, float& wy) {
    float gxr = P.goal_x - xf, gyr = P.goal_y - yf;
    float L = sqrtf(fmaf(gxr, gxr, gyr * gyr));
    if (L > 0.0f) {
        float phi = cn_atan2(gyr, gxr);
        float t = phi * 10.1859163578813f;
        float m = floorf(t);
        float delta = ((t - m) - 0.5f) * 0.0981747704246810f;
        float z = delta * delta;
        float cd = fmaf(fmaf(4.1666668e-2f, z, -0.5f), z, 1.0f);
        if (L * cd >= P.d.apothem) {
            float sc = P.d.apothem / (cd * L);
            wx = fmaf(sc, gxr, xf);
            wy = fmaf(sc, gyr, yf);
            return;
        }
    }
    wx = -P.goal_x;
    wy = P.goal_y;
}
__device__ __forceinline__ bool in_box(float x, float y, float lox, float hix, float loy, float hiy) {
    return (x <= hix) && (x > lox) && (y <= hiy) && (y > loy);
}
__device__ __forceinline__ bool in_goal_box(const cn_kparams& P, float x, float y) {
    return in_box(x, y, P.d.goal_lo_x, P.d.goal_hi_x, P.d.goal_lo_y, P.d.goal_hi_y);
}
__device__ __forceinline__ float dist_to_wp(float xf, float yf, float wx, float wy) {
    float dx = xf - wx, dy = yf - wy;
    return sqrtf(fmaf(dx, dx, dy * dy));
}
__device__ __forceinline__ float heading_to_wp(const cn_kparams& P, float xf, float yf, float yaw, float wx, float wy) {
    float px = xf + P.heading_off_x, py = yf + P.heading_off_y;
    float h = cn_atan2(wy - py, wx - px) - yaw;
    if (h > CN_PI) h -= CN_TWO_PI;
    else if (h < -CN_PI) h += CN_TWO_PI;
    return h;
}
__device__ __forceinline__ float cp_dto(const cn_kparams& P, float d) {
    if (d > P.max_range) return 0.0f;
    return (P.max_range - d) * P.d.inv_cp_span;
}
__device__ __forceinline__ int shaping_reward(float cur_head, float cur_dist, float prev_head, float prev_dist) {
    const float dd = cur_dist - prev_dist, dh = cur_head - prev_head;
    int reward = -2;
    if (dd < 0.0f) reward += 1;
    int htg = 0;
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

// ------------------------------------------------------------ span walking
// A span is up to two ranges of scan indices [a0, a1] U [b0, b1] within
// [1, NR]; walking it calls f(i, valid) for ALL lanes with warp-uniform loop
// bounds, so f may contain warp-synchronous code.
struct Span { int a0, a1, b0, b1; };

// rays whose angle i*inc lies within +-alpha of the relative bearing brel (padded, conservative)
__device__ __forceinline__ Span make_span(const cn_kparams& P, uint32_t brel, float alpha_rad) {
    const int NR = P.n_samples - 1;
    Span s; s.b0 = 1; s.b1 = 0;
    if (!(alpha_rad < 3.0f)) { s.a0 = 1; s.a1 = NR; return s; }
    const float two32 = 4294967296.0f;
    const float a = alpha_rad * CN_RAD2BIN;
    const float c = (float)brel;
    const float lo = c - a, hi = c + a;
    const float inv = P.d.inv_inc_bin;
    const int i0 = max((int)floorf(fmaxf(lo, 0.0f) * inv) - 1, 1);
    const int i1 = min((int)(fminf(hi, two32) * inv) + 2, NR);
    s.a0 = i0; s.a1 = i1;
    if (lo < 0.0f) { s.b0 = max(max((int)floorf((lo + two32) * inv) - 1, 1), i1 + 1); s.b1 = NR; }
    else if (hi >= two32) { s.b0 = 1; s.b1 = min(min((int)((hi - two32) * inv) + 2, NR), i0 - 1); }
    return s;
}
template <class F>
__device__ __forceinline__ void walk(const Span& s, int lane, F& f) {
    // one loop over both ranges (range a padded to whole 32-ray rounds) so the body is instantiated once
    const int la = (s.a1 >= s.a0) ? ((s.a1 - s.a0 + 32) & ~31) : 0;
    const int lb = (s.b1 >= s.b0) ? (s.b1 - s.b0 + 1) : 0;
    for (int base = 0; base < la + lb; base += 32) {      // warp-uniform trip count: f may vote
        const int k = base + lane;
        const bool in_a = k < la;
        const int i = in_a ? s.a0 + k : s.b0 + (k - la);
        f(i, in_a ? (i <= s.a1) : (i <= s.b1));
    }
}
// 32-ray chunks of the observation row (index j = NR - i) a span can touch
__device__ __forceinline__ uint32_t chunk_bits(int j_lo, int j_hi) {
    const int lo = j_lo >> 5, hi = min(j_hi >> 5, 31);
    if (hi < lo) return 0u;
    return (0xFFFFFFFFu >> (31 - hi)) & (0xFFFFFFFFu << lo);
}
__device__ __forceinline__ uint32_t span_chunks(const Span& s, int NR) {
    uint32_t m = 0;
    if (s.a1 >= s.a0) m |= chunk_bits(NR - s.a1, NR - s.a0);
    if (s.b1 >= s.b0) m |= chunk_bits(NR - s.b1, NR - s.b0);
    return m;
}

// rays that can see wall face `face` (0: +x, 1: -x, 2: +y, 3: -y) from the sensor origin (ox, oy) at yaw th, when the
// face is within range: acos(u) <= (pi/2) sqrt(1-u); the chunks of the row they touch are OR-ed into `chunks`
__device__ __forceinline__ Span face_span(const cn_kparams& P, float ox, float oy, uint32_t th, int face, uint32_t& chunks) {
    const bool xface = face < 2;
    const bool pos = (face & 1) == 0;
    const float wall = xface ? (pos ? P.room_xmax : P.room_xmin) : (pos ? P.room_ymax : P.room_ymin);
    const float o = xface ? ox : oy;
    const float Dw = pos ? (wall - o) : (o - wall);
    const float maxr = P.max_range;
    Span sp; sp.a0 = 1; sp.a1 = 0; sp.b0 = 1; sp.b1 = 0;
    if (Dw > 0.0f && Dw <= maxr * 1.0001f) {
        const uint32_t normal = xface ? (pos ? 0u : 0x80000000u) : (pos ? 0x40000000u : 0xC0000000u);
        const float alpha = CN_PIO2 * sqrtf(fmaxf(1.0f - Dw / maxr, 0.0f)) + 0.02f;
        sp = make_span(P, normal - th, alpha);
        chunks |= span_chunks(sp, P.n_samples - 1);
    }
    return sp;
}

// ------------------------------------------------------------------ phase A
// The part of Env.step / get_state / compute_reward that depends on the pose
// alone, for ONE world held by the calling thread.  `part` selects a slice of
// the work so the CTA's two pose warps can share it (0: waypoint chain + reward
// shaping, 1: velocities, rounded pose, goal boxes, wall spans, K-block padding)
// (auto-reset path, executed warp-uniformly by the world's own warp).
//   rob   : the world's robot record (state BEFORE this step; not modified)
//   sc    : the world's scalar record      row : the world's observation row
struct PoseIn { int32_t xi, yi; uint32_t th; float v, w; };

// Body contact of the robot (only with d.robot_contact: collision_range < robot_radius, the README test protocol's
// min_scan_range 0.0 -- otherwise the LiDAR threshold ends the episode before the body touches anything and the pose
// is integrated freely, as in the reference-in-the-loop traces).  One kinematic sub-step from (xi, yi) to the
// candidate (nx, ny): the centre stays robot_radius off the inner wall faces, and a sub-step that would bring the
// centre closer to a pedestrian (position at the START of the control period, like the pedestrians see the robot)
// while inside robot_radius + ped_radius of it is not taken; the heading still turns.
__device__ __forceinline__ void robot_contact_step(const cn_kparams& P, const uint32_t* peds, int n_peds,
                                                   int32_t xi, int32_t yi, int32_t& nx, int32_t& ny) {
    nx = min(max(nx, P.d.rob_xmin), P.d.rob_xmax);
    ny = min(max(ny, P.d.rob_ymin), P.d.rob_ymax);
    bool blocked = false;
#pragma unroll 1
    for (int n = 0; n < n_peds; ++n) {
        const int32_t px = (int32_t)peds[4 * n], py = (int32_t)peds[4 * n + 1];
        const float dxn = (float)(nx - px) * CN_GRID, dyn = (float)(ny - py) * CN_GRID;
        const float d2n = fmaf(dxn, dxn, dyn * dyn);
        if (d2n < P.d.rob_ped_r2) {
            const float dxo = (float)(xi - px) * CN_GRID, dyo = (float)(yi - py) * CN_GRID;
            if (d2n < fmaf(dxo, dxo, dyo * dyo)) blocked = true;
        }
    }
    if (blocked) { nx = xi; ny = yi; }
}

// peds_old: the world's pedestrian plane A as it was at the start of the step ([n_peds][4] words: x, y, vx, vy)
__device__ __forceinline__ PoseIn advance_robot(const cn_kparams& P, const uint32_t* rob, const float* action, int& bad,
                                                const uint32_t* peds_old, int n_peds) {
    // T2 (ENV:1190-1192), sanitised; R: unicycle, midpoint rule (FAKE:109-118, 156-167)
    float av = action[0], aw = action[1];
    bad = 0;
    if (!(fabsf(av) <= 3.0e38f) || !(fabsf(aw) <= 3.0e38f)) { av = 0.0f; aw = 0.0f; bad = 1; }
    av = fminf(fmaxf(av, -CN_ACT_V_LIMIT), CN_ACT_V_LIMIT);
    aw = fminf(fmaxf(aw, -CN_ACT_W_LIMIT), CN_ACT_W_LIMIT);
    const float half = (aw * CN_WHEEL_SEP) * 0.5f;
    const float tl = av - half, tr = av + half;           // wheel speed targets
    float cl = tl, cr = tr;
    const float st = P.d.wheel_step;
    if (st > 0.0f) {                                       // libgazebo_ros_diff_drive ramp (XACRO:65,70)
        const float cv = f_of(rob[CN_R_V]), cw = f_of(rob[CN_R_W]);
        const float ch = (cw * CN_WHEEL_SEP) * 0.5f;
        cl = cv - ch; cr = cv + ch;
    }
    PoseIn p;
    p.xi = (int32_t)rob[CN_R_X]; p.yi = (int32_t)rob[CN_R_Y]; p.th = rob[CN_R_TH];
    float v_body = 0.0f, w_body = 0.0f;
#pragma unroll 1
    for (int k = 0; k < P.n_substeps; ++k) {
        if (st > 0.0f) {
            cl += fminf(fmaxf(tl - cl, -st), st);
            cr += fminf(fmaxf(tr - cr, -st), st);
        }
        v_body = (cr + cl) * 0.5f;
        w_body = (cr - cl) * CN_INV_WHEEL_SEP;
        const float ds = v_body * P.d.dt_sub;
        const float dth = w_body * P.d.dt_sub;
        const int32_t dth_bin = cn_f2i(dth * CN_RAD2BIN);
        const uint32_t mid = p.th + (uint32_t)(dth_bin >> 1);
        float sm, cm; cn_sincos_bin(mid, &sm, &cm);
        int32_t nx = p.xi + cn_f2i((ds * cm) * CN_INV_GRID);
        int32_t ny = p.yi + cn_f2i((ds * sm) * CN_INV_GRID);
        if (P.d.robot_contact) robot_contact_step(P, peds_old, n_peds, p.xi, p.yi, nx, ny);
        p.xi = nx; p.yi = ny;
        p.th += (uint32_t)dth_bin;
    }
    p.v = v_body; p.w = w_body;
    return p;
}

__device__ __forceinline__ void pose_scalars(const cn_kparams& P, const PoseIn& p, int part,
                                             float wpx, float wpy, float prev_dist, float prev_head,
                                             float ppx, float ppy, int step_counter, bool have_prev, bool is_step,
                                             int bad, uint32_t* sc, float* row) {
    const int NR = P.n_samples - 1, K = P.k_obstacles;
    const float xf = (float)p.xi * CN_GRID, yf = (float)p.yi * CN_GRID;
    const float yaw = cn_bin2rad(p.th);
    const bool original = (P.flags & CN_FLAG_ENV_ORIGINAL) != 0u;
    if (part == 0 && original) {
        // the ORIGINAL environment (environment_stage_1_original.py:278-322): distance / heading to the goal itself;
        // compute_reward (original:324-410) reads state[-1], state[-2] -- the robot's y and x (sic) -- as "distance"
        // and "heading", has no step penalty and no waypoint bonus
        const float dist = cn_py_round2(dist_to_wp(xf, yf, P.goal_x, P.goal_y));
        const float head = cn_py_round2(heading_to_wp(P, xf, yf, yaw, P.goal_x, P.goal_y));
        const float px = cn_py_round3(xf), py = cn_py_round3(yf);
        sc[S_WPX] = u_of(P.goal_x); sc[S_WPY] = u_of(P.goal_y);
        sc[S_HEAD] = u_of(head); sc[S_DIST] = u_of(dist);
        sc[S_NPDIST] = u_of(is_step ? py : prev_dist); sc[S_NPHEAD] = u_of(is_step ? px : prev_head);
        sc[S_REWARD] = (uint32_t)(is_step ? shaping_reward(px, py, prev_head, prev_dist) + 2 : 0);
        row[NR + 0] = head; row[NR + 1] = dist;
    }
    if (part == 0 && !original) {
        // A: waypoint / distance / heading (ENV:246-265); the refresh target depends only on (pose, goal)
        float nwx, nwy;
        waypoint(P, xf, yf, nwx, nwy);
        float wx = wpx, wy = wpy;
        if (step_counter == 1) { wx = nwx; wy = nwy; }
        const float dist = cn_py_round2(dist_to_wp(xf, yf, wx, wy));
        const float head = cn_py_round2(heading_to_wp(P, xf, yf, yaw, wx, wy));
        if (step_counter % 5 == 0 || dist < prev_dist) { wx = nwx; wy = nwy; }
        int reward = 0;
        if (is_step) {
            // W: compute_reward (ENV:1046-1125); np.around(., 3) of a 2-dp value is the identity
            reward = shaping_reward(head, dist, prev_head, prev_dist);
            if (in_box(xf, yf, wx - P.goal_box, wx + P.goal_box, wy - P.goal_box, wy + P.goal_box)) {
                wx = nwx; wy = nwy;                               // ENV:1109-1116
                reward += 200;
                if (in_goal_box(P, wx, wy)) { wx = P.goal_x; wy = P.goal_y; }   // ENV:1121-1123
            }
        }
        sc[S_WPX] = u_of(wx); sc[S_WPY] = u_of(wy);
        sc[S_HEAD] = u_of(head); sc[S_DIST] = u_of(dist);
        // ENV:1133-1134 after a step; ENV:1243-1244 (unrounded, w.r.t. the goal) after a reset
        sc[S_NPDIST] = u_of(is_step ? dist : prev_dist); sc[S_NPHEAD] = u_of(is_step ? head : prev_head);
        sc[S_REWARD] = (uint32_t)reward;
        row[NR + 0] = head; row[NR + 1] = dist;
    }
    if (part == 1) {
        // B (ENV:267-268, yaw RATE used as an angle), rounded pose (ENV:1025-1027), agent speed (UTL:227-236)
        float sw, cw; cn_sincos_rad(p.w, &sw, &cw);
        const float avx = -1.0f * (p.v * cw), avy = p.v * sw;
        const float pcx = cn_py_round3(xf), pcy = cn_py_round3(yf);
        float agent_vel = 0.0f;
        if (have_prev) {
            const float vx = (pcx - ppx) * P.d.inv_dt, vy = (pcy - ppy) * P.d.inv_dt;
            agent_vel = sqrtf(fmaf(vx, vx, vy * vy));
        }
        float sy, cy; cn_sincos_bin(p.th, &sy, &cy);
        sc[S_XI] = (uint32_t)p.xi; sc[S_YI] = (uint32_t)p.yi; sc[S_TH] = p.th;
        sc[S_V] = u_of(p.v); sc[S_W] = u_of(p.w);
        sc[S_PCX] = u_of(pcx); sc[S_PCY] = u_of(pcy);
        sc[S_PPX] = u_of(ppx); sc[S_PPY] = u_of(ppy);
        sc[S_XF] = u_of(xf); sc[S_YF] = u_of(yf);
        sc[S_OFFX] = u_of(P.mount_x * cy); sc[S_OFFY] = u_of(P.mount_x * sy);
        sc[S_AVEL] = u_of(agent_vel);
        sc[S_BAD] = (uint32_t)bad;
        row[NR + 2] = pcx; row[NR + 3] = pcy;
        if (!original) {
            row[NR + 4] = cn_py_round3(yaw);
            row[NR + 5] = cn_py_round3(avx); row[NR + 6] = cn_py_round3(avy);
        }
    }
    if (part == 1) {
        uint32_t pre = 0;
        if (in_goal_box(P, xf, yf)) pre |= 1u;                    // ENV:1017
        if (step_counter >= P.max_steps) pre |= 2u;               // ENV:1021
        sc[S_PRE] = pre;
#if !defined(CN_WALLS_BY_LANE)
        // walls in range of the sensor: the rays that can see each face
        {
            float sy, cy; cn_sincos_bin(p.th, &sy, &cy);
            const float ox = xf + P.mount_x * cy, oy = yf + P.mount_x * sy;
            uint32_t wdirty = 0;
#if defined(CN_COMPACT_CODE)
#pragma unroll 1
#else
#pragma unroll
#endif
            for (int face = 0; face < 4; ++face) {      // 0: +x, 1: -x, 2: +y, 3: -y
                const Span sp = face_span(P, ox, oy, p.th, face, wdirty);
                sc[S_WSPAN + 4 * face + 0] = (uint32_t)sp.a0; sc[S_WSPAN + 4 * face + 1] = (uint32_t)sp.a1;
                sc[S_WSPAN + 4 * face + 2] = (uint32_t)sp.b0; sc[S_WSPAN + 4 * face + 3] = (uint32_t)sp.b1;
            }
            sc[S_WDIRTY] = wdirty;
        }
#endif
        // K-block padding (ENV:866-876, 895-898): [x, y, 0, 0] with the UNROUNDED pose, then np.around
        const float padx = cn_np_round3(xf), pady = cn_np_round3(yf);
        float* b = row + NR + 7;                                  // 16-B alignment is not guaranteed: scalar stores
        for (int s = 0; s < (original ? 0 : K); ++s) { b[4 * s] = padx; b[4 * s + 1] = pady; b[4 * s + 2] = 0.0f; b[4 * s + 3] = 0.0f; }
    }
}

__device__ __forceinline__ void add_rep(const cn_kparams& P, int32_t xi, int32_t yi, int32_t xj, int32_t yj,
                                        float rsum, float& vex, float& vey) {
    const float dx = (float)(xi - xj) * CN_GRID, dy = (float)(yi - yj) * CN_GRID;
    const float d2 = fmaf(dx, dx, dy * dy);
    const float lim = rsum + P.rep_cutoff;
    if (d2 < lim * lim && d2 > 0.0f) {
        const float d = sqrtf(d2);
        const float f = (P.rep_strength * cn_exp((rsum - d) / P.rep_range)) / d;
        vex += f * dx;
        vey += f * dy;
    }
}

}  // namespace

#endif /* CN_DEV_H */
