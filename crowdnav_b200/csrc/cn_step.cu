/*
 * cn_step.cu -- the fused env-step kernel for sm_100a.
 *
 * One launch advances E independent 2-D worlds by one control period:
 *   pedestrian integration (CROWD:98-144 + contact stand-in)
 *   -> unicycle robot (FAKE:109-167)
 *   -> LiDAR cast against pedestrian discs + the four wall faces (XACRO:148-179)
 *   -> scan cleaning / reversal (UTL:375-392)
 *   -> perceived-risk block: object points, tracker, collision cone, CP, top-K
 *      (ENV:568-907, `risk_intended` restatement)
 *   -> waypoint / heading / distance (ENV:246-265), done (ENV:1011-1023),
 *      reward (ENV:1046-1162), optional in-kernel auto-reset (ENV:1227-1263).
 *
 * Mapping.  A CTA owns a tile of TILE consecutive worlds.  The three state
 * planes of the tile (cn_state.h) are contiguous in HBM, so thread 0 pulls them
 * into shared memory with three 1-D bulk TMA copies (cp.async.bulk ->
 * UBLKCP) completing on one mbarrier, and pushes them -- plus the tile's
 * [TILE, D] block of observation rows, which is contiguous too -- back with
 * bulk stores.  In between, ONE WARP OWNS ONE WORLD: lanes are pedestrians for
 * the integration / risk phases and rays for the LiDAR / observation phases,
 * and all reductions (min range, centre-ray argmin, top-K rank) are warp
 * shuffles / redux / ballots; there is no inter-warp communication.
 *
 * LiDAR is span rasterisation, not brute force: the scan row starts at +inf and
 * each primitive (wall face within range, pedestrian within range + radius)
 * min-updates only the rays inside a conservative angular interval around it.
 * The interval is a superset by construction and the per-ray intersection
 * arithmetic is the oracle's, so results equal the brute-force cast bit for bit.
 *
 * Numerics: every value that reaches an output goes through cn_math.h
 * primitives; compile with -fmad=false (no contraction).  The only approximate
 * arithmetic is in choosing span bounds, which are padded.
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include "cn_state.h"
#include "cn_kernel.h"

#define FULL 0xFFFFFFFFu

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

// ------------------------------------------------------------- env registers
struct Robot {          // warp-uniform copy of one robot / episode record
    int32_t xi, yi;
    uint32_t th;
    float v, w, wpx, wpy, pdist, phead, ppx, ppy;
    int32_t step;
    uint32_t episode, flags, cnt0, cnt1;
};

__device__ __forceinline__ void robot_load(Robot& r, const uint32_t* s) {
    r.xi = (int32_t)s[CN_R_X]; r.yi = (int32_t)s[CN_R_Y]; r.th = s[CN_R_TH];
    r.v = __uint_as_float(s[CN_R_V]); r.w = __uint_as_float(s[CN_R_W]);
    r.wpx = __uint_as_float(s[CN_R_WPX]); r.wpy = __uint_as_float(s[CN_R_WPY]);
    r.pdist = __uint_as_float(s[CN_R_PDIST]); r.phead = __uint_as_float(s[CN_R_PHEAD]);
    r.ppx = __uint_as_float(s[CN_R_PPX]); r.ppy = __uint_as_float(s[CN_R_PPY]);
    r.step = (int32_t)s[CN_R_STEP]; r.episode = s[CN_R_EPISODE]; r.flags = s[CN_R_FLAGS];
    r.cnt0 = s[CN_R_CNT0]; r.cnt1 = s[CN_R_CNT1];
}
__device__ __forceinline__ void robot_store(const Robot& r, uint32_t* s) {
    uint4* q = reinterpret_cast<uint4*>(s);
    q[0] = make_uint4((uint32_t)r.xi, (uint32_t)r.yi, r.th, __float_as_uint(r.v));
    q[1] = make_uint4(__float_as_uint(r.w), __float_as_uint(r.wpx), __float_as_uint(r.wpy), __float_as_uint(r.pdist));
    q[2] = make_uint4(__float_as_uint(r.phead), __float_as_uint(r.ppx), __float_as_uint(r.ppy), (uint32_t)r.step);
    q[3] = make_uint4(r.episode, r.flags, r.cnt0, r.cnt1);
}

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

// ------------------------------------------------------------ span walking
// Visit every scan index i in [1, NR] whose ray angle i*inc lies within
// +-alpha of the relative bearing `brel` (binary angle, robot frame), padded.
// `f(i, valid)` is called by ALL lanes (valid = lane has an index) so it may
// contain warp-synchronous code; loop bounds are warp-uniform.
template <class F>
__device__ __forceinline__ void seg_walk(int s0, int s1, int lane, F& f) {
    for (int base = s0; base <= s1; base += 32) {
        int i = base + lane;
        f(i, i <= s1);
    }
}
template <class F>
__device__ __forceinline__ void span_walk(const cn_kparams& P, uint32_t brel, float alpha_rad, int lane, F& f) {
    const int NR = P.n_samples - 1;
    if (!(alpha_rad < 3.0f)) { seg_walk(1, NR, lane, f); return; }
    const float two32 = 4294967296.0f;
    float a = alpha_rad * CN_RAD2BIN;
    float c = (float)brel;
    float lo = c - a, hi = c + a;
    float inv = P.d.inv_inc_bin;
    int i0 = (int)floorf(fmaxf(lo, 0.0f) * inv) - 1;
    int i1 = (int)(fminf(hi, two32) * inv) + 2;
    seg_walk(max(i0, 1), min(i1, NR), lane, f);
    if (lo < 0.0f) {
        int w0 = (int)floorf((lo + two32) * inv) - 1;
        seg_walk(max(max(w0, 1), min(i1, NR) + 1), NR, lane, f);
    }
    if (hi >= two32) {
        int w1 = (int)((hi - two32) * inv) + 2;
        seg_walk(1, min(min(w1, NR), max(i0, 1) - 1), lane, f);
    }
}

// ------------------------------------------------------------------ observe
// Env.get_state for the warp's world.  Per-lane pedestrian state in registers
// (slot s of lane l is pedestrian l + 32 s).  Returns this step's done flag.
template <int NPL>
__device__ __forceinline__ bool observe(const cn_kparams& P, Robot& r, int lane, int step_counter, bool have_prev,
                                        const int32_t (&pxi)[NPL], const int32_t (&pyi)[NPL],
                                        float (&hitx)[NPL], float (&hity)[NPL], uint32_t (&pfl)[NPL],
                                        float* row, uint8_t* hid, size_t dbg_row) {
    const int N = P.n_peds, R = P.n_samples, K = P.k_obstacles, NR = R - 1;
    const float xf = (float)r.xi * CN_GRID, yf = (float)r.yi * CN_GRID;
    const uint32_t th = r.th;
    const float yaw = cn_bin2rad(th);

    // ---- A: waypoint / distance / heading (ENV:246-265).  The refresh target
    // depends only on (pose, goal), so it is evaluated once and reused.
    float nwx, nwy;
    waypoint(P, xf, yf, nwx, nwy);
    float wx = r.wpx, wy = r.wpy;
    if (step_counter == 1) { wx = nwx; wy = nwy; }
    const float dist = cn_py_round2(dist_to_wp(xf, yf, wx, wy));
    const float head = cn_py_round2(heading_to_wp(P, xf, yf, yaw, wx, wy));
    if (step_counter % 5 == 0 || dist < r.pdist) { wx = nwx; wy = nwy; }
    r.wpx = wx; r.wpy = wy;

    // ---- B: ENV:267-268
    float sw, cw; cn_sincos_rad(r.w, &sw, &cw);
    const float avx = -1.0f * (r.v * cw), avy = r.v * sw;

    // ---- L: LiDAR by span rasterisation
    for (int j = lane; j < NR; j += 32) row[j] = INFINITY;
    {
        uint32_t* h32 = reinterpret_cast<uint32_t*>(hid);
        for (int j = lane; j < (NR + 3) / 4; j += 32) h32[j] = 0xFFFFFFFFu;
    }
    __syncwarp();
    float sy, cy; cn_sincos_bin(th, &sy, &cy);
    const float offx = P.mount_x * cy, offy = P.mount_x * sy;
    const float ox = xf + offx, oy = yf + offy;
    const float maxr = P.max_range;

    // walls: x faces, then y faces (oracle order)
#pragma unroll 1
    for (int face = 0; face < 4; ++face) {
        // face 0: +x, 1: -x, 2: +y, 3: -y
        const bool xface = face < 2;
        const bool pos = (face & 1) == 0;
        const float wall = xface ? (pos ? P.room_xmax : P.room_xmin) : (pos ? P.room_ymax : P.room_ymin);
        const float o = xface ? ox : oy;
        const float D = pos ? (wall - o) : (o - wall);
        if (!(D > 0.0f) || D > maxr * 1.0001f) continue;
        const uint32_t normal = xface ? (pos ? 0u : 0x80000000u) : (pos ? 0x40000000u : 0xC0000000u);
        // rays with cos(angle to normal) >= D / maxr ; acos(u) <= (pi/2) sqrt(1-u)
        const float alpha = CN_PIO2 * sqrtf(fmaxf(1.0f - D / maxr, 0.0f)) + 0.02f;
        const float num = wall - o;
        auto f = [&](int i, bool valid) {
            if (!valid) return;
            float s, co; cn_sincos_bin(th + (uint32_t)i * P.d.inc_bin, &s, &co);
            const float den = xface ? co : s;
            if (pos ? !(den > 0.0f) : !(den < 0.0f)) return;
            const float t = num / den;
            const int j = NR - i;
            if (t > 0.0f && t <= maxr && t < row[j]) { row[j] = t; hid[j] = CN_HIT_WALL; }
        };
        span_walk(P, normal - th, alpha, lane, f);
        __syncwarp();
    }

    // pedestrians: per-lane candidate test, then a warp-uniform loop over candidates
    float qx[NPL], qy[NPL], alpha[NPL];
    uint32_t bearing[NPL];
    bool cand[NPL];
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        const int n = lane + 32 * s;
        cand[s] = false; qx[s] = 0.0f; qy[s] = 0.0f; alpha[s] = 0.0f; bearing[s] = 0u;
        if (n < N) {
            qx[s] = (float)(pxi[s] - r.xi) * CN_GRID - offx;
            qy[s] = (float)(pyi[s] - r.yi) * CN_GRID - offy;
            const float d2 = fmaf(qx[s], qx[s], qy[s] * qy[s]);
            cand[s] = d2 < P.d.cand_d2;
            if (cand[s]) {
                bearing[s] = cn_rad2bin(cn_atan2(qy[s], qx[s]));
                const float d = sqrtf(d2);
                if (d <= P.ped_radius * 1.001f) alpha[s] = 4.0f;       // sensor inside / touching: all rays
                else {
                    const float u = P.ped_radius / d;                   // asin(u) <= u + (pi/2 - 1) u^3
                    alpha[s] = u * fmaf(0.5708f * u, u, 1.0f) + 0.01f;
                }
            }
        }
    }
    uint32_t cmask[NPL];
#pragma unroll
    for (int s = 0; s < NPL; ++s) cmask[s] = __ballot_sync(FULL, cand[s]);

#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        uint32_t m = cmask[s];
        while (m) {
            const int src = __ffs(m) - 1; m &= m - 1;
            const float cqx = __shfl_sync(FULL, qx[s], src), cqy = __shfl_sync(FULL, qy[s], src);
            const float cal = __shfl_sync(FULL, alpha[s], src);
            const uint32_t cb = __shfl_sync(FULL, bearing[s], src);
            const uint8_t id = (uint8_t)(src + 32 * s);
            auto f = [&](int i, bool valid) {
                if (!valid) return;
                float sn, co; cn_sincos_bin(th + (uint32_t)i * P.d.inc_bin, &sn, &co);
                const float b = fmaf(cqx, co, cqy * sn);
                const float h = fmaf(cqx, sn, -(cqy * co));
                const float disc = fmaf(-h, h, P.d.ped_r2);
                if (disc < 0.0f) return;
                const float sq = sqrtf(disc);
                if (!(b + sq > 0.0f)) return;
                float t = b - sq;
                if (t < 0.0f) t = 0.0f;
                const int j = NR - i;
                if (t <= maxr && t < row[j]) { row[j] = t; hid[j] = id; }
            };
            span_walk(P, cb - th, cal, lane, f);
            __syncwarp();
        }
    }

    // ---- E-I: per candidate, count the rays it owns and find its centre ray
    int cnt[NPL], jstar[NPL];
#pragma unroll
    for (int s = 0; s < NPL; ++s) { cnt[s] = 0; jstar[s] = 0; }
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        uint32_t m = cmask[s];
        while (m) {
            const int src = __ffs(m) - 1; m &= m - 1;
            const float cal = __shfl_sync(FULL, alpha[s], src);
            const uint32_t cb = __shfl_sync(FULL, bearing[s], src);
            const uint8_t id = (uint8_t)(src + 32 * s);
            int c = 0; uint32_t bkey = 0xFFFFFFFFu; int bj = 0x7FFFFFFF;
            auto f = [&](int i, bool valid) {
                bool hit = false;
                if (valid) {
                    const int j = NR - i;
                    hit = hid[j] == id;
                    if (hit) {
                        const int32_t delta = (int32_t)(th + (uint32_t)i * P.d.inc_bin - cb);
                        const uint32_t ad = (delta < 0) ? (0u - (uint32_t)delta) : (uint32_t)delta;
                        const uint32_t key = (ad & ~1u) | (delta < 0 ? 1u : 0u);
                        if (key < bkey || (key == bkey && j < bj)) { bkey = key; bj = j; }
                    }
                }
                c += __popc(__ballot_sync(FULL, hit));
            };
            span_walk(P, cb - th, cal, lane, f);
            const uint32_t kmin = __reduce_min_sync(FULL, bkey);
            const int jm = (int)__reduce_min_sync(FULL, (uint32_t)(bkey == kmin ? bj : 0x7FFFFFFF));
            if (lane == src) { cnt[s] = c; jstar[s] = jm; }
        }
    }

    // ---- clean + min (UTL:375-392, ENV:1012)
    float mn = INFINITY;
    for (int j = lane; j < NR; j += 32) {
        float t = row[j];
        float rr = (hid[j] == CN_HIT_NONE) ? maxr : ((t < P.sensor_min_range) ? P.sensor_min_range : t);
        row[j] = rr;
        mn = fminf(mn, rr);
        if (P.dbg_ranges) P.dbg_ranges[dbg_row * NR + j] = rr;
        if (P.dbg_hid) P.dbg_hid[dbg_row * NR + j] = hid[j];
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));

    // ---- H-J: per confirmed pedestrian (lane-parallel)
    const float pcx = cn_py_round3(xf), pcy = cn_py_round3(yf);
    float agent_vel = 0.0f;
    if (have_prev) {
        const float vx = (pcx - r.ppx) * P.d.inv_dt, vy = (pcy - r.ppy) * P.d.inv_dt;
        agent_vel = sqrtf(fmaf(vx, vx, vy * vy));
    }
    bool conf[NPL], inblk[NPL];
    float o_cp[NPL], o_x[NPL], o_y[NPL], o_vx[NPL], o_vy[NPL], o_ttc[NPL];
    bool ego_v = false;
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        const int n = lane + 32 * s;
        conf[s] = (n < N) && cand[s] && cnt[s] >= 4;
        inblk[s] = false; o_cp[s] = 0.0f; o_x[s] = 0.0f; o_y[s] = 0.0f; o_vx[s] = 0.0f; o_vy[s] = 0.0f; o_ttc[s] = 0.0f;
        if (!conf[s]) { pfl[s] &= ~CN_PF_TRACKED; continue; }
        const float d_raw = row[jstar[s]];
        const float d3 = cn_py_round3(d_raw);
        float sa, ca; cn_sincos_bin((uint32_t)jstar[s] * P.d.hit_inc_bin - th, &sa, &ca);
        const float hx = cn_py_round3(xf + d_raw * ca);
        const float hy = cn_py_round3(yf + (d_raw * sa) * -1.0f);
        float chx = 0.0f, chy = 0.0f, speed = -1.0f, ovx = 0.0f, ovy = 0.0f;
        if (pfl[s] & CN_PF_TRACKED) {
            chx = hitx[s] - hx; chy = hity[s] - hy;
            speed = sqrtf(fmaf(chy, chy, chx * chx)) * P.d.inv_dt;
            ovx = chx * P.d.inv_dt; ovy = chy * P.d.inv_dt;
        }
        hitx[s] = hx; hity[s] = hy; pfl[s] |= CN_PF_TRACKED;
        if (d3 < 0.140f) ego_v = true;
        if (!have_prev) continue;
        const float tx = pcx + chx, ty = pcy + chy;
        float ux = tx - r.ppx, uy = ty - r.ppy;
        const float L = sqrtf(fmaf(ux, ux, uy * uy));
        bool have_dtc = false; float dtc = 0.0f;
        if (L > 0.0f) {
            const float invL = 1.0f / L;
            ux = ux * invL; uy = uy * invL;
            const float wx_ = hx - r.ppx, wy_ = hy - r.ppy;
            const float b = fmaf(wx_, ux, wy_ * uy);
            const float h = fmaf(wx_, uy, -(wy_ * ux));
            const float disc = fmaf(-h, h, P.d.cp_r2);
            if (disc > 0.0f) {
                const float t = b - sqrtf(disc);
                if (t > 0.0f) { have_dtc = true; dtc = t; }
            }
        }
        const float resultant = agent_vel - speed;
        float cp_ttc = 0.0f, cp;
        const float dto = cp_dto(P, d3);
        if (have_dtc && resultant == 0.0f) {
            cp = dto;
        } else {
            if (have_dtc) {
                const float q = (0.15f * resultant) / dtc;
                cp_ttc = (q < 1.0f) ? q : 1.0f;
            }
            cp = 0.5f * cp_ttc + 0.5f * dto;
        }
        inblk[s] = true; o_cp[s] = cp; o_x[s] = hx; o_y[s] = hy; o_vx[s] = ovx; o_vy[s] = ovy; o_ttc[s] = cp_ttc;
    }
    uint32_t confm[NPL], blkm[NPL];
    int n_seen = 0, n_obj = 0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        confm[s] = __ballot_sync(FULL, conf[s]); blkm[s] = __ballot_sync(FULL, inblk[s]);
        n_seen += __popc(confm[s]); n_obj += __popc(blkm[s]);
    }
    const bool ego_violation = __any_sync(FULL, ego_v);
    float ego_score = 0.0f;
    if (n_obj > 0) {
        float e = -INFINITY;
#pragma unroll
        for (int s = 0; s < NPL; ++s) if (inblk[s]) e = fmaxf(e, o_ttc[s]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) e = fmaxf(e, __shfl_xor_sync(FULL, e, o));
        ego_score = e;
    }

    // ---- K: top-K block (ENV:862-907): pad, then rank-select
    float* blk = row + NR + 7;
    for (int k = lane; k < 4 * K; k += 32) {
        const int c4 = k & 3;
        blk[k] = (c4 == 0) ? xf : ((c4 == 1) ? yf : 0.0f);
    }
    __syncwarp();
    if (n_obj > 0) {
        int rank[NPL];
#pragma unroll
        for (int s = 0; s < NPL; ++s) rank[s] = 0;
#pragma unroll
        for (int sb = 0; sb < NPL; ++sb) {
            uint32_t m = blkm[sb];
            while (m) {
                const int src = __ffs(m) - 1; m &= m - 1;
                const float cpb = __shfl_sync(FULL, o_cp[sb], src);
                const int nb = src + 32 * sb;
#pragma unroll
                for (int s = 0; s < NPL; ++s) {
                    const int n = lane + 32 * s;
                    if (nb != n && (cpb > o_cp[s] || (cpb == o_cp[s] && nb < n))) ++rank[s];
                }
            }
        }
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            if (!inblk[s]) continue;
            int slot = (P.flags & CN_FLAG_TOPK_HIGHEST) ? rank[s] : rank[s] - (n_obj > K ? n_obj - K : 0);
            if (slot < 0 || slot >= K) continue;
            blk[4 * slot + 0] = o_x[s]; blk[4 * slot + 1] = o_y[s];
            blk[4 * slot + 2] = o_vx[s]; blk[4 * slot + 3] = o_vy[s];
        }
    }

    // ---- M: counters (ENV:653-654, 998-1005)
    {
        uint32_t ego = r.cnt0 & 0xFFFFu, soc = r.cnt0 >> 16;
        uint32_t pres = r.cnt1 & 0xFFFFu, bad = r.cnt1 >> 16;
        if (n_seen > 0 && pres < 0xFFFFu) ++pres;
        if (ego_violation && ego < 0xFFFFu) ++ego;
        if (ego_score > 0.4f && soc < 0xFFFFu) ++soc;
        r.cnt0 = ego | (soc << 16);
        r.cnt1 = pres | (bad << 16);
    }

    // ---- N: done (ENV:1011-1023)
    bool done = false;
    if (mn < P.collision_range) done = true;
    if (in_goal_box(P, xf, yf)) done = true;
    if (step_counter >= P.max_steps) done = true;

    // ---- O: assemble + round (ENV:1025-1042)
    if (lane == 0) {
        row[NR + 0] = head; row[NR + 1] = dist;
        row[NR + 2] = pcx; row[NR + 3] = pcy;
        row[NR + 4] = cn_py_round3(yaw);
        row[NR + 5] = cn_py_round3(avx); row[NR + 6] = cn_py_round3(avy);
    }
    __syncwarp();
    for (int k = lane; k < P.d.obs_dim; k += 32) row[k] = cn_np_round3(row[k]);
    __syncwarp();

    r.ppx = pcx; r.ppy = pcy;
    return done;
}

// ----------------------------------------------------------------- reset_env
// Env.reset (ENV:1227-1263) + gazebo/reset_simulation for the warp's world.
template <int NPL>
__device__ __forceinline__ void reset_env(const cn_kparams& P, Robot& r, int lane, uint32_t gid,
                                          int32_t (&pxi)[NPL], int32_t (&pyi)[NPL], float (&pvx)[NPL], float (&pvy)[NPL],
                                          float (&hitx)[NPL], float (&hity)[NPL], int32_t (&timer)[NPL], uint32_t (&pfl)[NPL],
                                          float* row, uint8_t* hid, size_t dbg_row) {
    const int N = P.n_peds;
    const uint32_t episode = r.episode + 1u;
    const int b = (int)(gid % (uint32_t)P.n_behaviors);
    r.xi = P.d.start_xi; r.yi = P.d.start_yi; r.th = P.d.start_th;
    r.v = 0.0f; r.w = 0.0f; r.step = 0; r.episode = episode; r.flags = 0; r.cnt0 = 0; r.cnt1 = 0;
    r.ppx = 0.0f; r.ppy = 0.0f;
    const int stagger = __ldg(&P.cfg->behavior_stagger_ticks[b]);
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        const int n = lane + 32 * s;
        if (n < N) {
            const cn_u32x2 rnd = cn_env_rand(P.d.seed_lo, P.d.seed_hi, gid, episode, 0u, (uint32_t)n, 1u);
            const float px = __ldg(&P.cfg->ped_layout[n][0]) + cn_usym(rnd.v[0], P.layout_jitter);
            const float py = __ldg(&P.cfg->ped_layout[n][1]) + cn_usym(rnd.v[1], P.layout_jitter);
            int32_t xi = cn_f2i(px * CN_INV_GRID), yi = cn_f2i(py * CN_INV_GRID);
            xi = max(xi, P.d.ped_xmin); xi = min(xi, P.d.ped_xmax);
            yi = max(yi, P.d.ped_ymin); yi = min(yi, P.d.ped_ymax);
            pxi[s] = xi; pyi[s] = yi; pvx[s] = 0.0f; pvy[s] = 0.0f;
            hitx[s] = 0.0f; hity[s] = 0.0f; timer[s] = (n + 1) * stagger; pfl[s] = 0u;
        }
    }
    const float xf = (float)r.xi * CN_GRID, yf = (float)r.yi * CN_GRID;
    r.wpx = P.goal_x; r.wpy = P.goal_y;
    r.pdist = dist_to_wp(xf, yf, r.wpx, r.wpy);
    r.phead = heading_to_wp(P, xf, yf, cn_bin2rad(r.th), r.wpx, r.wpy);
    (void)observe<NPL>(P, r, lane, 0, false, pxi, pyi, hitx, hity, pfl, row, hid, dbg_row);
    r.cnt0 = 0; r.cnt1 = 0;
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

// ------------------------------------------------------------------- kernel
// MODE 0: step, MODE 1: reset (masked)
template <int NPL, int MODE>
__global__ void __launch_bounds__(32 * CN_TILE, (NPL == 1) ? 2 : 1)
cn_env_kernel(const __grid_constant__ cn_kparams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int N = P.n_peds, NR = P.n_samples - 1, D = P.d.obs_dim;
    const int e0 = blockIdx.x * CN_TILE;
    const int nE = min(CN_TILE, P.n_envs - e0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // shared-memory carve-up (tile planes first: they are TMA targets)
    uint32_t* s_robot = reinterpret_cast<uint32_t*>(smem);                     // [TILE][16]
    uint32_t* s_pa = s_robot + CN_TILE * CN_ROBOT_WORDS;                        // [TILE][N][4]
    uint32_t* s_pb = s_pa + CN_TILE * N * 4;                                    // [TILE][N][4]
    float* s_obs = reinterpret_cast<float*>(s_pb + CN_TILE * N * 4);            // [TILE][D]
    const int hid_stride = (NR + 15) & ~15;
    uint8_t* s_hid = reinterpret_cast<uint8_t*>(s_obs + (((size_t)CN_TILE * D + 3) & ~(size_t)3));  // [TILE][hid_stride]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_hid + (size_t)CN_TILE * hid_stride);

    const uint32_t rob_bytes = (uint32_t)nE * CN_ROBOT_WORDS * 4u;
    const uint32_t ped_bytes = (uint32_t)nE * (uint32_t)N * 16u;
    if (threadIdx.x == 0) {
        mbar_init(s_bar, 1);
        fence_mbar_init();
        mbar_expect_tx(s_bar, rob_bytes + 2u * ped_bytes);
        tma_load(s_robot, P.robot + (size_t)e0 * CN_ROBOT_WORDS, rob_bytes, s_bar);
        if (ped_bytes) {
            tma_load(s_pa, P.ped_a + (size_t)e0 * N * 4, ped_bytes, s_bar);
            tma_load(s_pb, P.ped_b + (size_t)e0 * N * 4, ped_bytes, s_bar);
        }
    }
    __syncthreads();            // barrier init visible to every waiter
    mbar_wait(s_bar, 0);

    const int e = e0 + warp;
    bool active = warp < nE;
    if (MODE == 1 && active && P.mask) active = P.mask[e] != 0;
    float* row = s_obs + (size_t)warp * D;
    if (active) {
        uint8_t* hid = s_hid + (size_t)warp * hid_stride;
        uint32_t* srob = s_robot + warp * CN_ROBOT_WORDS;
        uint4* spa = reinterpret_cast<uint4*>(s_pa) + (size_t)warp * N;
        uint4* spb = reinterpret_cast<uint4*>(s_pb) + (size_t)warp * N;
        const uint32_t gid = (uint32_t)(P.env_id_offset + e);
        const int b = (int)(gid % (uint32_t)P.n_behaviors);

        Robot r; robot_load(r, srob);
        int32_t pxi[NPL], pyi[NPL], timer[NPL];
        float pvx[NPL], pvy[NPL], hitx[NPL], hity[NPL];
        uint32_t pfl[NPL];
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            const int n = lane + 32 * s;
            pxi[s] = 0; pyi[s] = 0; timer[s] = 0; pvx[s] = 0.0f; pvy[s] = 0.0f; hitx[s] = 0.0f; hity[s] = 0.0f; pfl[s] = 0u;
            if (n < N) {
                const uint4 a = spa[n], bb = spb[n];
                pxi[s] = (int32_t)a.x; pyi[s] = (int32_t)a.y; pvx[s] = __uint_as_float(a.z); pvy[s] = __uint_as_float(a.w);
                hitx[s] = __uint_as_float(bb.x); hity[s] = __uint_as_float(bb.y); timer[s] = (int32_t)bb.z; pfl[s] = bb.w;
            }
        }

        if (MODE == 1) {
            reset_env<NPL>(P, r, lane, gid, pxi, pyi, pvx, pvy, hitx, hity, timer, pfl, row, hid, (size_t)e);
            r.flags = 0;
        } else {
            const int step_counter = r.step + 1;
            // ---- T2: action (ENV:1190-1192), sanitised
            float av = P.action[2 * (size_t)e], aw = P.action[2 * (size_t)e + 1];
            if (!(fabsf(av) <= 3.0e38f) || !(fabsf(aw) <= 3.0e38f)) {
                av = 0.0f; aw = 0.0f;
                uint32_t bad = r.cnt1 >> 16;
                if (bad < 0xFFFFu) ++bad;
                r.cnt1 = (r.cnt1 & 0xFFFFu) | (bad << 16);
            }
            av = fminf(fmaxf(av, -CN_ACT_V_LIMIT), CN_ACT_V_LIMIT);
            aw = fminf(fmaxf(aw, -CN_ACT_W_LIMIT), CN_ACT_W_LIMIT);

            // ---- P: pedestrians, Jacobi on the old positions still in s_pa
            const int kind = __ldg(&P.cfg->behavior_kind[b]);
            const float speed = __ldg(&P.cfg->behavior_speed[b]);
            const int period = __ldg(&P.cfg->behavior_period_ticks[b]);
            const float rr2 = P.ped_radius + P.ped_radius, rrob = P.ped_radius + P.robot_radius;
            // conservative integer prefilter for the contact test
            const int32_t lim_i = (int32_t)((fmaxf(rr2, rrob) + P.rep_cutoff) * CN_INV_GRID) + 64;
            int32_t nxi[NPL], nyi[NPL];
#pragma unroll
            for (int s = 0; s < NPL; ++s) {
                const int n = lane + 32 * s;
                nxi[s] = pxi[s]; nyi[s] = pyi[s];
                if (n < N) {
                    float vx = pvx[s], vy = pvy[s];
                    int32_t tm = timer[s] - CN_TICKS_PER_STEP;
                    if (tm <= 0) {
                        if (kind == CN_BEHAVIOR_RANDOM) {
                            const cn_u32x2 rnd = cn_env_rand(P.d.seed_lo, P.d.seed_hi, gid, r.episode,
                                                             (uint32_t)step_counter, (uint32_t)n, 0u);
                            vx = cn_usym(rnd.v[0], speed);
                            vy = cn_usym(rnd.v[1], speed);
                        } else {
                            vx = __ldg(&P.cfg->behavior_table[b][n][0]) * speed;
                            vy = __ldg(&P.cfg->behavior_table[b][n][1]) * speed;
                        }
                        tm += period;
                    }
                    float vex = vx, vey = vy;
                    for (int m = 0; m < N; ++m) {
                        const uint2 o = *reinterpret_cast<const uint2*>(&spa[m]);
                        const int32_t dxi = pxi[s] - (int32_t)o.x, dyi = pyi[s] - (int32_t)o.y;
                        if (m != n && abs(dxi) < lim_i && abs(dyi) < lim_i)
                            add_rep(P, pxi[s], pyi[s], (int32_t)o.x, (int32_t)o.y, rr2, vex, vey);
                    }
                    add_rep(P, pxi[s], pyi[s], r.xi, r.yi, rrob, vex, vey);
                    int32_t nx = pxi[s] + cn_f2i((vex * P.dt) * CN_INV_GRID);
                    int32_t ny = pyi[s] + cn_f2i((vey * P.dt) * CN_INV_GRID);
                    if (nx < P.d.ped_xmin) { nx = P.d.ped_xmin; if (vx < 0.0f) vx = 0.0f; }
                    if (nx > P.d.ped_xmax) { nx = P.d.ped_xmax; if (vx > 0.0f) vx = 0.0f; }
                    if (ny < P.d.ped_ymin) { ny = P.d.ped_ymin; if (vy < 0.0f) vy = 0.0f; }
                    if (ny > P.d.ped_ymax) { ny = P.d.ped_ymax; if (vy > 0.0f) vy = 0.0f; }
                    nxi[s] = nx; nyi[s] = ny; pvx[s] = vx; pvy[s] = vy; timer[s] = tm;
                }
            }
#pragma unroll
            for (int s = 0; s < NPL; ++s) { pxi[s] = nxi[s]; pyi[s] = nyi[s]; }

            // ---- R: unicycle, midpoint rule (FAKE:109-118, 156-167)
            {
                const float half = (aw * CN_WHEEL_SEP) * 0.5f;
                const float vl = av - half, vr = av + half;
                const float v_body = (vr + vl) * 0.5f;
                const float w_body = (vr - vl) * CN_INV_WHEEL_SEP;
                const float ds = v_body * P.dt;
                const float dth = w_body * P.dt;
                const int32_t dth_bin = cn_f2i(dth * CN_RAD2BIN);
                const uint32_t mid = r.th + (uint32_t)(dth_bin >> 1);
                float sm, cm; cn_sincos_bin(mid, &sm, &cm);
                r.xi += cn_f2i((ds * cm) * CN_INV_GRID);
                r.yi += cn_f2i((ds * sm) * CN_INV_GRID);
                r.th += (uint32_t)dth_bin;
                r.v = v_body;
                r.w = w_body;
            }

            // ---- get_state
            const bool done_now = observe<NPL>(P, r, lane, step_counter, true, pxi, pyi, hitx, hity, pfl, row, hid, (size_t)e);
            const bool done = ((r.flags & CN_RF_DONE) != 0) || done_now;

            // ---- W: compute_reward (ENV:1046-1162) on the rounded heading / distance
            const float cur_head = row[NR + 0], cur_dist = row[NR + 1];
            const float prev_head = r.phead, prev_dist = r.pdist;
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
            reward += htg;
            const float xf = (float)r.xi * CN_GRID, yf = (float)r.yi * CN_GRID;
            if (in_box(xf, yf, r.wpx - P.goal_box, r.wpx + P.goal_box, r.wpy - P.goal_box, r.wpy + P.goal_box)) {
                float wx, wy; waypoint(P, xf, yf, wx, wy);
                reward += 200;
                if (in_goal_box(P, wx, wy)) { wx = P.goal_x; wy = P.goal_y; }
                r.wpx = wx; r.wpy = wy;
            }
            r.pdist = cur_dist; r.phead = cur_head;
            if (done) {
                r.flags |= CN_RF_DONE;
                if (in_goal_box(P, xf, yf)) { r.flags |= CN_RF_SUCCESS; r.flags &= ~CN_RF_FAILURE; reward += 200; }
                else { r.flags |= CN_RF_FAILURE; r.flags &= ~CN_RF_SUCCESS; reward -= 200; }
            }
            r.step = step_counter;
            if (lane == 0) {
                P.reward[e] = (float)reward;
                P.done[e] = done ? 1 : 0;
            }
            if (done && (P.flags & CN_FLAG_AUTO_RESET)) {
                const uint32_t keep = r.flags & (CN_RF_SUCCESS | CN_RF_FAILURE);
                __syncwarp();
                reset_env<NPL>(P, r, lane, gid, pxi, pyi, pvx, pvy, hitx, hity, timer, pfl, row, hid, (size_t)e);
                r.flags |= keep;
            }
        }

        // ---- write the world back into the shared tile
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            const int n = lane + 32 * s;
            if (n < N) {
                spa[n] = make_uint4((uint32_t)pxi[s], (uint32_t)pyi[s], __float_as_uint(pvx[s]), __float_as_uint(pvy[s]));
                spb[n] = make_uint4(__float_as_uint(hitx[s]), __float_as_uint(hity[s]), (uint32_t)timer[s], pfl[s]);
            }
        }
        if (lane == 0) robot_store(r, srob);
        // observation row: plain coalesced stores when the tile cannot go out as one bulk copy
        const bool bulk_obs = (MODE == 0) && (((size_t)nE * D) % 4 == 0) && P.obs_bulk_ok;
        if (!bulk_obs) {
            __syncwarp();
            float* g = P.obs + (size_t)e * D;
            for (int k = lane; k < D; k += 32) g[k] = row[k];
        }
    }

    // ---- tile write-back: bulk TMA stores from shared memory
    fence_async_smem();          // generic-proxy writes -> visible to the async proxy
    __syncthreads();
    if (threadIdx.x == 0) {
        tma_store(P.robot + (size_t)e0 * CN_ROBOT_WORDS, s_robot, rob_bytes);
        if (ped_bytes) {
            tma_store(P.ped_a + (size_t)e0 * N * 4, s_pa, ped_bytes);
            tma_store(P.ped_b + (size_t)e0 * N * 4, s_pb, ped_bytes);
        }
        const bool bulk_obs = (MODE == 0) && (((size_t)nE * D) % 4 == 0) && P.obs_bulk_ok;
        if (bulk_obs) tma_store(P.obs + (size_t)e0 * D, s_obs, (uint32_t)((size_t)nE * D * 4));
        tma_store_commit_and_wait();
    }
}

// sticky-done clear / counters: trivial elementwise kernels
__global__ void cn_clear_done_kernel(uint32_t* robot, const uint8_t* mask, int E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E && (!mask || mask[e])) robot[(size_t)e * CN_ROBOT_WORDS + CN_R_FLAGS] &= ~CN_RF_DONE;
}
__global__ void cn_counters_kernel(const uint32_t* robot, int32_t* out, int E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint32_t* r = robot + (size_t)e * CN_ROBOT_WORDS;
    int4 o;
    o.x = (r[CN_R_FLAGS] & CN_RF_SUCCESS) ? 1 : 0;
    o.y = (int32_t)(r[CN_R_CNT0] & 0xFFFFu);
    o.z = (int32_t)(r[CN_R_CNT0] >> 16);
    o.w = (int32_t)(r[CN_R_CNT1] & 0xFFFFu);
    reinterpret_cast<int4*>(out)[e] = o;
}

}  // namespace

// ------------------------------------------------------------- host launchers
size_t cn_kernel_smem_bytes(int n_peds, int n_samples, int obs_dim) {
    const int NR = n_samples - 1;
    size_t b = (size_t)CN_TILE * CN_ROBOT_WORDS * 4;
    b += 2 * (size_t)CN_TILE * n_peds * 16;
    b += (((size_t)CN_TILE * obs_dim + 3) & ~(size_t)3) * 4;
    b += (size_t)CN_TILE * ((NR + 15) & ~15);
    b += 16;
    return b;
}

template <int NPL, int MODE>
static cudaError_t launch_t(const cn_kparams& P, size_t smem, cudaStream_t stream) {
    auto k = cn_env_kernel<NPL, MODE>;
    static bool attr_set = false;
    static size_t attr_smem = 0;
    if (!attr_set || smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true; attr_smem = smem;
    }
    const int grid = (P.n_envs + CN_TILE - 1) / CN_TILE;
    k<<<grid, 32 * CN_TILE, smem, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t cn_launch_env_kernel(const cn_kparams& P, int mode, cudaStream_t stream) {
    const size_t smem = cn_kernel_smem_bytes(P.n_peds, P.n_samples, P.d.obs_dim);
    if (P.n_peds <= 32) return mode == 0 ? launch_t<1, 0>(P, smem, stream) : launch_t<1, 1>(P, smem, stream);
    return mode == 0 ? launch_t<2, 0>(P, smem, stream) : launch_t<2, 1>(P, smem, stream);
}

cudaError_t cn_launch_clear_done(uint32_t* robot, const uint8_t* mask, int E, cudaStream_t stream) {
    cn_clear_done_kernel<<<(E + 255) / 256, 256, 0, stream>>>(robot, mask, E);
    return cudaGetLastError();
}
cudaError_t cn_launch_counters(const uint32_t* robot, int32_t* out, int E, cudaStream_t stream) {
    cn_counters_kernel<<<(E + 255) / 256, 256, 0, stream>>>(robot, out, E);
    return cudaGetLastError();
}
