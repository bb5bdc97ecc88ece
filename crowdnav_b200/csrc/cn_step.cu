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
 * Mapping.  A CTA owns a tile of CN_TILE consecutive worlds.  The three state
 * planes of the tile (cn_state.h) are contiguous in HBM, so thread 0 pulls them
 * into shared memory with three 1-D bulk TMA copies (cp.async.bulk -> UBLKCP)
 * completing on one mbarrier, and pushes them -- plus the tile's [TILE, D] block
 * of observation rows, which is contiguous too -- back with bulk stores.
 *
 * Work inside the CTA is laid out so that no instruction is spent 32 times on
 * a value that exists once per world:
 *   phase A  (lane = WORLD, warps 0..2)  everything get_state / compute_reward
 *            derive from the pose alone: robot kinematics, waypoint, heading,
 *            distance, agent velocity, rounded pose, reward shaping, goal boxes.
 *            Results go to a per-world scalar record in shared memory and
 *            straight into the observation row.
 *   phase P  (lane = PEDESTRIAN, one warp per world, concurrent with A)
 *            resample / contact / integrate.
 *   phase L  (lane = RAY, one warp per world) LiDAR by SPAN RASTERISATION: the
 *            row starts at its no-return value and each primitive (wall face in
 *            range, pedestrian within range + radius) min-updates only the rays
 *            inside a conservative angular interval; 32-ray chunks that no span
 *            touched are never visited again.  The interval is a superset by
 *            construction and the per-ray arithmetic is the oracle's, so the
 *            result equals the brute-force cast bit for bit.
 *   phase R  (lane = PEDESTRIAN) ray ownership count + centre ray (REDUX),
 *            collision cone, CP, top-K by ballot-ranked selection.
 *
 * Numerics: every value that reaches an output goes through cn_math.h
 * primitives; compile with -fmad=false (no contraction).  The only approximate
 * arithmetic is in choosing span bounds, which are padded.
 */
#include "cn_dev.h"

namespace {


// ------------------------------------------------------------- ray phases
// LiDAR + perceived-risk block for the warp's world (Env.get_state from
// ENV:277 on).  Pedestrian state is per lane (slot s of lane l = pedestrian
// l + 32 s).  Returns min(scan) < collision_range.
template <int NPL>
__device__ __forceinline__ bool scan_and_risk(const cn_kparams& P, const uint32_t* sc, int lane, bool have_prev,
                                              const int32_t (&pxi)[NPL], const int32_t (&pyi)[NPL],
                                              float (&hitx)[NPL], float (&hity)[NPL], uint32_t (&pfl)[NPL],
                                              float* row, uint8_t* hid, size_t dbg_row,
                                              uint32_t& cnt0, uint32_t& cnt1) {
    const int N = P.n_peds, R = P.n_samples, K = P.k_obstacles, NR = R - 1;
    const int32_t rxi = (int32_t)sc[S_XI], ryi = (int32_t)sc[S_YI];
    const uint32_t th = sc[S_TH];
    const float xf = f_of(sc[S_XF]), yf = f_of(sc[S_YF]);
    const float offx = f_of(sc[S_OFFX]), offy = f_of(sc[S_OFFY]);
    const float ox = xf + offx, oy = yf + offy;
    const float maxr = P.max_range;
    const int D = P.d.obs_dim;

    uint32_t dirty = sc[S_WDIRTY];

    // ---- pedestrians: per-lane candidate test + span, then a warp-uniform loop over candidates
    float qx[NPL], qy[NPL];
    uint32_t bearing[NPL];
    Span span[NPL];
    bool cand[NPL];
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        const int n = lane + 32 * s;
        cand[s] = false; qx[s] = 0.0f; qy[s] = 0.0f; bearing[s] = 0u;
        span[s].a0 = 1; span[s].a1 = 0; span[s].b0 = 1; span[s].b1 = 0;
        if (n < N) {
            qx[s] = (float)(pxi[s] - rxi) * CN_GRID - offx;
            qy[s] = (float)(pyi[s] - ryi) * CN_GRID - offy;
            const float d2 = fmaf(qx[s], qx[s], qy[s] * qy[s]);
            cand[s] = d2 < P.d.cand_d2;
            if (cand[s]) {
                bearing[s] = cn_rad2bin(cn_atan2(qy[s], qx[s]));
                float alpha = 4.0f;                                    // sensor inside / touching the disc: all rays
                const float rlim = P.ped_radius * 1.001f;
                if (d2 > rlim * rlim) {
                    const float u = P.ped_radius * rsqrtf(d2) * 1.0001f;   // asin(u) <= u + (pi/2 - 1) u^3
                    alpha = u * fmaf(0.5708f * u, u, 1.0f) + 0.01f;
                }
                span[s] = make_span(P, bearing[s] - th, alpha);
            }
        }
    }
    // chunks each candidate can touch; `dup` = chunks claimed by more than one primitive (walls included).  A
    // candidate whose chunks are all its own is final as soon as it is rasterised: its ray count and centre ray are
    // taken in the same pass.  Only overlapping candidates need the separate ownership pass below.
    uint32_t cmask[NPL], cchunks[NPL];
    uint32_t dup = 0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        cmask[s] = __ballot_sync(FULL, cand[s]);
        cchunks[s] = cand[s] ? span_chunks(span[s], NR) : 0u;
    }
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        for (uint32_t m = cmask[s]; m; m &= m - 1) {
            const uint32_t cm = __shfl_sync(FULL, cchunks[s], __ffs(m) - 1);
            dup |= dirty & cm;
            dirty |= cm;
        }
    }
    int cnt[NPL], jstar[NPL];
#pragma unroll
    for (int s = 0; s < NPL; ++s) { cnt[s] = 0; jstar[s] = 0; }

    // ---- row <- "no return" (already in its final rounded form) where nothing can hit; the chunks a span
    // touches get the same value and are the only ones visited again.  hit ids <- none.
    {
        const float fill = P.d.max_range_r3;
        float2* r2 = reinterpret_cast<float2*>(row);          // rows are 8-byte aligned (D even) or handled below
        if ((D & 1) == 0) {
            for (int j = lane; j < (NR >> 1); j += 32) r2[j] = make_float2(fill, fill);
            if ((NR & 1) && lane == 0) row[NR - 1] = fill;
        } else {
            for (int j = lane; j < NR; j += 32) row[j] = fill;
        }
        uint4* h4 = reinterpret_cast<uint4*>(hid);
        for (int j = lane; j < (NR + 15) / 16; j += 32) h4[j] = make_uint4(~0u, ~0u, ~0u, ~0u);
    }
    __syncwarp();
    STAMP(9);

    // ---- walls: x faces, then y faces (oracle order)
    if (dirty)
#pragma unroll 1
    for (int face = 0; face < 4; ++face) {
        Span sp;
        sp.a0 = (int)sc[S_WSPAN + 4 * face + 0]; sp.a1 = (int)sc[S_WSPAN + 4 * face + 1];
        sp.b0 = (int)sc[S_WSPAN + 4 * face + 2]; sp.b1 = (int)sc[S_WSPAN + 4 * face + 3];
        if (sp.a1 < sp.a0) continue;
        const bool xface = face < 2;
        const bool pos = (face & 1) == 0;
        const float wall = xface ? (pos ? P.room_xmax : P.room_xmin) : (pos ? P.room_ymax : P.room_ymin);
        const float num = wall - (xface ? ox : oy);
        auto f = [&](int i, bool valid) {
            if (!valid) return;
            float s, co; cn_sincos_bin(th + (uint32_t)i * P.d.inc_bin, &s, &co);
            const float den = xface ? co : s;
            if (pos ? !(den > 0.0f) : !(den < 0.0f)) return;
            const float t = num / den;
            const int j = NR - i;
            if (t > 0.0f && t < maxr && t < row[j]) { row[j] = t; hid[j] = CN_HIT_WALL; }
        };
        walk(sp, lane, f);
        __syncwarp();
    }

#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        uint32_t m = cmask[s];
        while (m) {
            const int src = __ffs(m) - 1; m &= m - 1;
            const float cqx = __shfl_sync(FULL, qx[s], src), cqy = __shfl_sync(FULL, qy[s], src);
            const uint32_t cb = __shfl_sync(FULL, bearing[s], src);
            const bool iso = (__shfl_sync(FULL, cchunks[s], src) & dup) == 0u;
            Span sp;
            sp.a0 = __shfl_sync(FULL, span[s].a0, src); sp.a1 = __shfl_sync(FULL, span[s].a1, src);
            sp.b0 = __shfl_sync(FULL, span[s].b0, src); sp.b1 = __shfl_sync(FULL, span[s].b1, src);
            const uint8_t id = (uint8_t)(src + 32 * s);
            int c = 0; uint32_t bkey = 0xFFFFFFFFu; int bj = 0x7FFFFFFF;
            auto f = [&](int i, bool valid) {
                bool hit = false;
                if (valid) {
                    float sn, co; cn_sincos_bin(th + (uint32_t)i * P.d.inc_bin, &sn, &co);
                    const float b = fmaf(cqx, co, cqy * sn);
                    const float h = fmaf(cqx, sn, -(cqy * co));
                    const float disc = fmaf(-h, h, P.d.ped_r2);
                    if (disc >= 0.0f) {
                        const float sq = sqrtf(disc);
                        if (b + sq > 0.0f) {
                            float t = b - sq;
                            if (t < 0.0f) t = 0.0f;
                            const int j = NR - i;
                            if (t < maxr && t < row[j]) {
                                row[j] = t; hid[j] = id; hit = true;
                                if (iso) {          // E-I in the same pass: nothing else can touch these rays
                                    const int32_t delta = (int32_t)(th + (uint32_t)i * P.d.inc_bin - cb);
                                    const uint32_t ad = (delta < 0) ? (0u - (uint32_t)delta) : (uint32_t)delta;
                                    const uint32_t key = (ad & ~1u) | (delta < 0 ? 1u : 0u);
                                    if (key < bkey || (key == bkey && j < bj)) { bkey = key; bj = j; }
                                }
                            }
                        }
                    }
                }
                if (iso) c += __popc(__ballot_sync(FULL, hit));
            };
            walk(sp, lane, f);
            __syncwarp();
            if (iso && c >= 4) {
                const uint32_t kmin = __reduce_min_sync(FULL, bkey);
                const int jm = (int)__reduce_min_sync(FULL, (uint32_t)(bkey == kmin ? bj : 0x7FFFFFFF));
                if (lane == src) { cnt[s] = c; jstar[s] = jm; }
            }
        }
    }

    STAMP(10);
    // ---- E-I for candidates that share chunks with another primitive: count the rays each still owns and
    // find its centre ray once everything is rasterised
    if (dup) {
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            uint32_t m = __ballot_sync(FULL, cand[s] && (cchunks[s] & dup) != 0u);
            while (m) {
                const int src = __ffs(m) - 1; m &= m - 1;
                const uint32_t cb = __shfl_sync(FULL, bearing[s], src);
                Span sp;
                sp.a0 = __shfl_sync(FULL, span[s].a0, src); sp.a1 = __shfl_sync(FULL, span[s].a1, src);
                sp.b0 = __shfl_sync(FULL, span[s].b0, src); sp.b1 = __shfl_sync(FULL, span[s].b1, src);
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
                walk(sp, lane, f);
                if (c >= 4) {
                    const uint32_t kmin = __reduce_min_sync(FULL, bkey);
                    const int jm = (int)__reduce_min_sync(FULL, (uint32_t)(bkey == kmin ? bj : 0x7FFFFFFF));
                    if (lane == src) { cnt[s] = c; jstar[s] = jm; }
                }
            }
        }
    }

    STAMP(11);
    // ---- debug taps (tests only): raw cleaned ranges + hit ids for every ray
    if (P.dbg_ranges || P.dbg_hid) {
        for (int j = lane; j < NR; j += 32) {
            const uint8_t h = hid[j];
            const float t = row[j];
            const float rr = (h == CN_HIT_NONE) ? maxr : ((t < P.sensor_min_range) ? P.sensor_min_range : t);
            if (P.dbg_ranges) P.dbg_ranges[dbg_row * NR + j] = rr;
            if (P.dbg_hid) P.dbg_hid[dbg_row * NR + j] = h;
        }
    }

    // ---- clean + min (UTL:375-392, ENV:1012) + np.around, only where a span went
    float mn = maxr;
    for (uint32_t dm = dirty; dm; dm &= dm - 1) {
        const int j = ((__ffs(dm) - 1) << 5) + lane;
        if (j < NR && hid[j] != CN_HIT_NONE) {
            const float t = row[j];
            const float rr = (t < P.sensor_min_range) ? P.sensor_min_range : t;
            mn = fminf(mn, rr);
            row[j] = rr;            // raw for the centre-ray lookups below; rounded afterwards
        }
    }
    __syncwarp();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = fminf(mn, __shfl_xor_sync(FULL, mn, o));

    STAMP(12);
    // ---- H-J: per confirmed pedestrian (lane-parallel)
    const float pcx = f_of(sc[S_PCX]), pcy = f_of(sc[S_PCY]);
    const float ppx = f_of(sc[S_PPX]), ppy = f_of(sc[S_PPY]);
    const float agent_vel = f_of(sc[S_AVEL]);
    bool conf[NPL], inblk[NPL];
    float o_cp[NPL], o_x[NPL], o_y[NPL], o_vx[NPL], o_vy[NPL], o_ttc[NPL];
    bool ego_v = false;
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
        conf[s] = cnt[s] >= 4;
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
        float ux = tx - ppx, uy = ty - ppy;
        const float L = sqrtf(fmaf(ux, ux, uy * uy));
        bool have_dtc = false; float dtc = 0.0f;
        if (L > 0.0f) {
            const float invL = 1.0f / L;
            ux = ux * invL; uy = uy * invL;
            const float wx_ = hx - ppx, wy_ = hy - ppy;
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
    __syncwarp();

    STAMP(13);
    // ---- np.around of the rays a span touched (everything else already holds the rounded no-return value)
    for (uint32_t dm = dirty; dm; dm &= dm - 1) {
        const int j = ((__ffs(dm) - 1) << 5) + lane;
        if (j < NR && hid[j] != CN_HIT_NONE) row[j] = cn_np_round3(row[j]);
    }

    if (n_seen > 0) {
        const bool ego_violation = __any_sync(FULL, ego_v);
        float ego_score = 0.0f;
        if (n_obj > 0) {
            float e = -INFINITY;
#pragma unroll
            for (int s = 0; s < NPL; ++s) if (inblk[s]) e = fmaxf(e, o_ttc[s]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) e = fmaxf(e, __shfl_xor_sync(FULL, e, o));
            ego_score = e;

            // ---- K: top-K block (ENV:862-907): stable rank by CP, keep [-K:]; padding is already in the row
            float* blk = row + NR + 7;
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
                const int slot = (P.flags & CN_FLAG_TOPK_HIGHEST) ? rank[s] : rank[s] - (n_obj > K ? n_obj - K : 0);
                if (slot < 0 || slot >= K) continue;
                blk[4 * slot + 0] = o_x[s]; blk[4 * slot + 1] = o_y[s];          // already multiples of 0.001
                blk[4 * slot + 2] = cn_np_round3(o_vx[s]); blk[4 * slot + 3] = cn_np_round3(o_vy[s]);
            }
        }
        // ---- M: counters (ENV:653-654, 998-1005)
        uint32_t ego = cnt0 & 0xFFFFu, soc = cnt0 >> 16;
        uint32_t pres = cnt1 & 0xFFFFu;
        if (pres < 0xFFFFu) ++pres;
        if (ego_violation && ego < 0xFFFFu) ++ego;
        if (ego_score > 0.4f && soc < 0xFFFFu) ++soc;
        cnt0 = ego | (soc << 16);
        cnt1 = pres | (cnt1 & 0xFFFF0000u);
    }
    __syncwarp();
    return mn < P.collision_range;
}


// reset a world's pedestrians (lane = pedestrian): gazebo/reset_simulation
template <int NPL>
__device__ __forceinline__ void reset_peds(const cn_kparams& P, int lane, uint32_t gid, uint32_t episode,
                                           int32_t (&pxi)[NPL], int32_t (&pyi)[NPL], float (&pvx)[NPL], float (&pvy)[NPL],
                                           float (&hitx)[NPL], float (&hity)[NPL], int32_t (&timer)[NPL], uint32_t (&pfl)[NPL]) {
    const int N = P.n_peds;
    const int b = (int)(gid % (uint32_t)P.n_behaviors);
    const int stagger = P.beh_stagger[b];
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
}

// ------------------------------------------------------------------- kernel
// MODE 0: step (with next-step auto-reset), MODE 1: reset of the masked worlds.
template <int NPL, int MODE>
__global__ void __launch_bounds__(CN_CTA_THREADS, (NPL == 1) ? 2 : 1)
cn_env_kernel(const __grid_constant__ cn_kparams P) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int N = P.n_peds, NR = P.n_samples - 1, D = P.d.obs_dim;
    const int e0 = blockIdx.x * CN_TILE;
    const int nE = min(CN_TILE, P.n_envs - e0);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    STAMP(0);

    // shared-memory carve-up (tile planes first: they are TMA targets)
    uint32_t* s_robot = reinterpret_cast<uint32_t*>(smem);                     // [TILE][16]
    uint32_t* s_pa = s_robot + CN_TILE * CN_ROBOT_WORDS;                        // [TILE][N][4]
    uint32_t* s_pb = s_pa + CN_TILE * N * 4;                                    // [TILE][N][4]
    float* s_obs = reinterpret_cast<float*>(s_pb + CN_TILE * N * 4);            // [TILE][D]
    const int hid_stride = (NR + 15) & ~15;
    uint8_t* s_hid = reinterpret_cast<uint8_t*>(s_obs + (((size_t)CN_TILE * D + 3) & ~(size_t)3));  // [TILE][hid_stride]
    uint32_t* s_sc = reinterpret_cast<uint32_t*>(s_hid + (size_t)CN_TILE * hid_stride);              // [TILE][S_WORDS]
    uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_sc + CN_TILE * S_WORDS);
    uint64_t* s_barA = s_bar + 1;       // phase A done: one arrival per THREAD of the phase-A warps (every writer releases its own stores)
    float* s_act = reinterpret_cast<float*>(s_bar + 2);                                              // [TILE][2]

    const uint32_t rob_bytes = (uint32_t)nE * CN_ROBOT_WORDS * 4u;
    const uint32_t ped_bytes = (uint32_t)nE * (uint32_t)N * 16u;
    const bool act_smem = (MODE == 0) && P.act_bulk_ok && (nE % 2) == 0;
    if (threadIdx.x == 0) {
        mbar_init(s_bar, 1);
        mbar_init(s_barA, 32 * CN_POSE_WARPS);
        fence_mbar_init();
        const uint32_t act_bytes = act_smem ? (uint32_t)nE * 8u : 0u;
        mbar_expect_tx(s_bar, rob_bytes + 2u * ped_bytes + act_bytes);
        if (act_bytes) tma_load(s_act, P.action + 2 * (size_t)e0, act_bytes, s_bar);
        tma_load(s_robot, P.robot + (size_t)e0 * CN_ROBOT_WORDS, rob_bytes, s_bar);
        if (ped_bytes) {
            tma_load(s_pa, P.ped_a + (size_t)e0 * N * 4, ped_bytes, s_bar);
            tma_load(s_pb, P.ped_b + (size_t)e0 * N * 4, ped_bytes, s_bar);
        }
    }
    __syncthreads();            // barrier init visible to every waiter
    mbar_wait(s_bar, 0);
    STAMP(1);

    // ---- phase A: lane = world, the CTA's two extra warps share the pose-only work of the whole tile
    // (they own no world, so nobody waits for them longer than for the pedestrian phase)
    if (warp >= CN_TILE) {
        const int part = warp - CN_TILE;
        if (lane < nE) {
            const uint32_t* rob = s_robot + lane * CN_ROBOT_WORDS;
            bool run = true, reset_now = (MODE == 1);
            if (MODE == 1) run = !P.mask || P.mask[e0 + lane] != 0;
            else reset_now = (rob[CN_R_FLAGS] & CN_RF_DONE) && (P.flags & CN_FLAG_AUTO_RESET);
            if (run) {
                uint32_t* scl = s_sc + lane * S_WORDS;
                float* rowl = s_obs + (size_t)lane * D;
                if (reset_now) {
                    // Z: Env.reset (ENV:1227-1263): spawn pose, waypoint = goal, unrounded previous_* (ENV:1243-1244)
                    PoseIn p; p.xi = P.d.start_xi; p.yi = P.d.start_yi; p.th = P.d.start_th; p.v = 0.0f; p.w = 0.0f;
                    const float xf = (float)p.xi * CN_GRID, yf = (float)p.yi * CN_GRID;
                    const float pd = dist_to_wp(xf, yf, P.goal_x, P.goal_y);
                    const float ph = heading_to_wp(P, xf, yf, cn_bin2rad(p.th), P.goal_x, P.goal_y);
                    pose_scalars(P, p, part, P.goal_x, P.goal_y, pd, ph, 0.0f, 0.0f, 0, false, false, 0, scl, rowl);
                } else {
                    int bad;
                    const PoseIn p = advance_robot(P, rob, act_smem ? s_act + 2 * lane : P.action + 2 * (size_t)(e0 + lane), bad,
                                                   s_pa + (size_t)lane * N * 4, N);
                    pose_scalars(P, p, part, f_of(rob[CN_R_WPX]), f_of(rob[CN_R_WPY]), f_of(rob[CN_R_PDIST]),
                                 f_of(rob[CN_R_PHEAD]), f_of(rob[CN_R_PPX]), f_of(rob[CN_R_PPY]), (int)rob[CN_R_STEP] + 1,
                                 true, true, bad, scl, rowl);
                }
            }
        }
        __syncwarp();
        mbar_arrive(s_barA);
    }
    STAMP(2);

    const int e = e0 + warp;
    bool active = warp < nE;        // pose warps (warp >= CN_TILE >= nE) own no world
    if (MODE == 1 && active && P.mask) active = P.mask[e] != 0;
    float* row = s_obs + (size_t)warp * D;
    uint8_t* hid = s_hid + (size_t)warp * hid_stride;
    uint32_t* srob = s_robot + warp * CN_ROBOT_WORDS;
    uint32_t* sc = s_sc + warp * S_WORDS;
    uint4* spa = reinterpret_cast<uint4*>(s_pa) + (size_t)warp * N;
    uint4* spb = reinterpret_cast<uint4*>(s_pb) + (size_t)warp * N;
    const uint32_t gid = (uint32_t)(P.env_id_offset + e);

    if (active) {
        uint32_t flags = srob[CN_R_FLAGS], cnt0 = srob[CN_R_CNT0], cnt1 = srob[CN_R_CNT1];
        uint32_t episode = srob[CN_R_EPISODE];
        int step = (int)srob[CN_R_STEP];
        const bool reset_now = (MODE == 1) || ((flags & CN_RF_DONE) && (P.flags & CN_FLAG_AUTO_RESET));

        int32_t pxi[NPL], pyi[NPL], timer[NPL];
        float pvx[NPL], pvy[NPL], hitx[NPL], hity[NPL];
        uint32_t pfl[NPL];
#pragma unroll
        for (int s = 0; s < NPL; ++s) { pxi[s] = 0; pyi[s] = 0; timer[s] = 0; pvx[s] = 0.0f; pvy[s] = 0.0f; hitx[s] = 0.0f; hity[s] = 0.0f; pfl[s] = 0u; }

        // ---- phase P: lane = pedestrian
        if (reset_now) {
            episode += 1u;
            reset_peds<NPL>(P, lane, gid, episode, pxi, pyi, pvx, pvy, hitx, hity, timer, pfl);
        } else {
            // P: CROWD:98-144 + contact stand-in, Jacobi on the old positions in s_pa / old robot pose in s_robot
            const int b = (P.n_behaviors == 1) ? 0 : (int)(gid % (uint32_t)P.n_behaviors);
            const int kind = P.beh_kind[b];
            const float speed = P.beh_speed[b];
            const int period = P.beh_period[b];
            const float rr2 = P.ped_radius + P.ped_radius, rrob = P.ped_radius + P.robot_radius;
            const int32_t lim_i = (int32_t)((fmaxf(rr2, rrob) + P.rep_cutoff) * CN_INV_GRID) + 64;   // conservative prefilter
            const uint32_t lim2 = 2u * (uint32_t)lim_i;
            const int32_t rxi = (int32_t)srob[CN_R_X], ryi = (int32_t)srob[CN_R_Y];
            const int step_counter = step + 1;
            uint32_t peers = 0;
            if (NPL == 1) {
                // contacts are rare: find the pairs inside a conservative contact box by rotating the pedestrian list
                // against itself (N/2 shuffle rounds).  Both coordinates travel in ONE word: 14-bit fields at
                // 2^-8 m with a guard bit each, so "|dx| < 64 q and |dy| < 64 q" is one subtract and one mask test.
                uint32_t pk = 0u;
                if (lane < N) {
                    const uint2 a = *reinterpret_cast<const uint2*>(&spa[lane]);
                    pk = (((a.x + 0x20000000u) >> 16) & 0x3FFFu) | ((((a.y + 0x20000000u) >> 16) & 0x3FFFu) << 16);
                }
                const uint32_t pk_biased = (pk | 0x80008000u) + 0x00400040u;
                const int half = N >> 1;
                int partner = lane;
                for (int r = 1; r <= half; ++r) {
                    partner = (partner + 1 >= N) ? partner + 1 - N : partner + 1;       // (lane + r) mod N
                    const uint32_t other = __shfl_sync(FULL, pk, (lane < N) ? partner : lane);
                    const bool hit = lane < N && ((pk_biased - other) & 0xFF80FF80u) == 0x80008000u;
                    const uint32_t hm = __ballot_sync(FULL, hit);
                    if (hm) {                                                           // rare
                        if (hit) peers |= 1u << partner;
                        const int back = (lane - r < 0) ? lane - r + N : lane - r;      // the lane whose partner I am
                        if (lane < N && ((hm >> back) & 1u)) peers |= 1u << back;
                    }
                }
                peers &= ~(1u << lane);
            }
            if (peers == 0x12345678u) STAMP(15);
            STAMP(8);
#pragma unroll
            for (int s = 0; s < NPL; ++s) {
                const int n = lane + 32 * s;
                if (n < N) {
                    const uint4 a = spa[n], bb = spb[n];
                    const int32_t x0 = (int32_t)a.x, y0 = (int32_t)a.y;
                    float vx = f_of(a.z), vy = f_of(a.w);
                    hitx[s] = f_of(bb.x); hity[s] = f_of(bb.y); pfl[s] = bb.w;
                    int32_t tm = (int32_t)bb.z - CN_TICKS_PER_STEP;
                    if (tm <= 0) {
                        if (kind == CN_BEHAVIOR_RANDOM) {
                            const cn_u32x2 rnd = cn_env_rand(P.d.seed_lo, P.d.seed_hi, gid, episode,
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
                    if (NPL == 1) {
                        for (uint32_t pm = peers; pm; pm &= pm - 1) {          // index order, like the oracle
                            const int m = __ffs(pm) - 1;
                            const uint2 o = *reinterpret_cast<const uint2*>(&spa[m]);
                            add_rep(P, x0, y0, (int32_t)o.x, (int32_t)o.y, rr2, vex, vey);
                        }
                    } else {
                        const uint32_t bx = (uint32_t)(x0 + lim_i), by = (uint32_t)(y0 + lim_i);
#pragma unroll 4
                        for (int m = 0; m < N; ++m) {
                            const uint2 o = *reinterpret_cast<const uint2*>(&spa[m]);
                            if ((bx - o.x) < lim2 && (by - o.y) < lim2 && m != n)
                                add_rep(P, x0, y0, (int32_t)o.x, (int32_t)o.y, rr2, vex, vey);
                        }
                    }
                    if ((uint32_t)(x0 - rxi + lim_i) < lim2 && (uint32_t)(y0 - ryi + lim_i) < lim2)
                        add_rep(P, x0, y0, rxi, ryi, rrob, vex, vey);
                    int32_t nx = x0 + cn_f2i((vex * P.dt) * CN_INV_GRID);
                    int32_t ny = y0 + cn_f2i((vey * P.dt) * CN_INV_GRID);
                    if (nx < P.d.ped_xmin) { nx = P.d.ped_xmin; if (vx < 0.0f) vx = 0.0f; }
                    if (nx > P.d.ped_xmax) { nx = P.d.ped_xmax; if (vx > 0.0f) vx = 0.0f; }
                    if (ny < P.d.ped_ymin) { ny = P.d.ped_ymin; if (vy < 0.0f) vy = 0.0f; }
                    if (ny > P.d.ped_ymax) { ny = P.d.ped_ymax; if (vy > 0.0f) vy = 0.0f; }
                    pxi[s] = nx; pyi[s] = ny; pvx[s] = vx; pvy[s] = vy; timer[s] = tm;
                }
            }
        }
        STAMP(3);
        mbar_wait(s_barA, 0);   // this world's scalar record + observation-row scalars are in place
        STAMP(4);

        // ---- phases L + R (Env.get_state from ENV:277 on)
        if (!reset_now && sc[S_BAD]) { uint32_t bad = cnt1 >> 16; if (bad < 0xFFFFu) ++bad; cnt1 = (cnt1 & 0xFFFFu) | (bad << 16); }
        const bool collided = scan_and_risk<NPL>(P, sc, lane, !reset_now, pxi, pyi, hitx, hity, pfl, row, hid, (size_t)e, cnt0, cnt1);
        if (reset_now) {
            cnt0 = 0; cnt1 = 0;                                                 // ENV:1260-1262
            flags = (MODE == 0) ? (flags & (CN_RF_SUCCESS | CN_RF_FAILURE)) : 0u;   // last episode's status stays readable
            step = 0;
            if (MODE == 0 && lane == 0) { P.reward[e] = 0.0f; P.done[e] = 2; }
        } else {
            // ---- N + W: done (ENV:1011-1023), terminal reward (ENV:1136-1159; a time-out is -200 too)
            const uint32_t pre = sc[S_PRE];
            const bool done = ((flags & CN_RF_DONE) != 0) || collided || pre != 0u;
            int reward = (int)sc[S_REWARD];
            if (done) {
                flags |= CN_RF_DONE;
                if (pre & 1u) { flags |= CN_RF_SUCCESS; flags &= ~CN_RF_FAILURE; reward += 200; }
                else { flags |= CN_RF_FAILURE; flags &= ~CN_RF_SUCCESS; reward -= 200; }
            }
            step += 1;
            if (lane == 0) { P.reward[e] = (float)reward; P.done[e] = done ? 1 : 0; }
        }

        // ---- write the world back into the shared tile
#pragma unroll
        for (int s = 0; s < NPL; ++s) {
            const int n = lane + 32 * s;
            if (n < N) {
                spa[n] = make_uint4((uint32_t)pxi[s], (uint32_t)pyi[s], u_of(pvx[s]), u_of(pvy[s]));
                spb[n] = make_uint4(u_of(hitx[s]), u_of(hity[s]), (uint32_t)timer[s], pfl[s]);
            }
        }
        if (lane == 0) {
            uint4* q = reinterpret_cast<uint4*>(srob);
            q[0] = make_uint4(sc[S_XI], sc[S_YI], sc[S_TH], sc[S_V]);
            q[1] = make_uint4(sc[S_W], sc[S_WPX], sc[S_WPY], sc[S_NPDIST]);
            q[2] = make_uint4(sc[S_NPHEAD], sc[S_PCX], sc[S_PCY], (uint32_t)step);
            q[3] = make_uint4(episode, flags, cnt0, cnt1);
        }
        // observation row: plain coalesced stores when the tile cannot go out as one bulk copy
        const bool bulk_obs = (MODE == 0) && (((size_t)nE * D) % 4 == 0) && P.obs_bulk_ok;
        if (!bulk_obs) {
            __syncwarp();
            float* g = P.obs + (size_t)e * D;
            for (int k = lane; k < D; k += 32) g[k] = row[k];
            for (int p = 0; p < P.n_obs_peers; ++p) {          // fused all-gather, ragged-tile path
                float* gp = P.obs_peers[p] + (size_t)e * D;
                for (int k = lane; k < D; k += 32) gp[k] = row[k];
            }
        }
    } else {
        mbar_wait(s_barA, 0);
    }

    // ---- tile write-back: bulk TMA stores from shared memory
    STAMP(5);
    fence_async_smem();          // generic-proxy writes -> visible to the async proxy
    __syncthreads();
    STAMP(6);
    if (threadIdx.x == 0) {
        tma_store(P.robot + (size_t)e0 * CN_ROBOT_WORDS, s_robot, rob_bytes);
        if (ped_bytes) {
            tma_store(P.ped_a + (size_t)e0 * N * 4, s_pa, ped_bytes);
            tma_store(P.ped_b + (size_t)e0 * N * 4, s_pb, ped_bytes);
        }
        const bool bulk_obs = (MODE == 0) && (((size_t)nE * D) % 4 == 0) && P.obs_bulk_ok;
        if (bulk_obs) {
            tma_store(P.obs + (size_t)e0 * D, s_obs, (uint32_t)((size_t)nE * D * 4));
            // fused all-gather: the same tile goes straight into every peer's gather buffer over NVLink
            for (int p = 0; p < P.n_obs_peers; ++p)
                tma_store(P.obs_peers[p] + (size_t)e0 * D, s_obs, (uint32_t)((size_t)nE * D * 4));
        }
        tma_store_commit_and_wait();
    }
    STAMP(7);
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
    b += (size_t)CN_TILE * S_WORDS * 4;
    b += 16 + (size_t)CN_TILE * 8;
    return b;
}

template <int NPL, int MODE>
static cudaError_t launch_t(const cn_kparams& P, size_t smem, cudaStream_t stream) {
    auto k = cn_env_kernel<NPL, MODE>;
    {
        cudaError_t e = cn_ensure_smem_attr(reinterpret_cast<const void*>(k), 10 + (NPL - 1) * 2 + MODE, smem);
        if (e != cudaSuccess) return e;
    }
    const int grid = (P.n_envs + CN_TILE - 1) / CN_TILE;
    k<<<grid, CN_CTA_THREADS, smem, stream>>>(P);
    return cudaGetLastError();
}

cudaError_t cn_launch_env_kernel(const cn_kparams& P, int mode, cudaStream_t stream) {
    const size_t smem = cn_kernel_smem_bytes(P.n_peds, P.n_samples, P.d.obs_dim);
    if (P.n_peds <= 32) return mode == 0 ? launch_t<1, 0>(P, smem, stream) : launch_t<1, 1>(P, smem, stream);
    return mode == 0 ? launch_t<2, 0>(P, smem, stream) : launch_t<2, 1>(P, smem, stream);
}

#ifdef CN_TIMELINE
extern "C" int cn_debug_set_timeline(unsigned long long* dev_ptr) {
    return (int)cudaMemcpyToSymbol(g_timeline, &dev_ptr, sizeof(dev_ptr));
}
#endif

cudaError_t cn_launch_clear_done(uint32_t* robot, const uint8_t* mask, int E, cudaStream_t stream) {
    cn_clear_done_kernel<<<(E + 255) / 256, 256, 0, stream>>>(robot, mask, E);
    return cudaGetLastError();
}
cudaError_t cn_launch_counters(const uint32_t* robot, int32_t* out, int E, cudaStream_t stream) {
    cn_counters_kernel<<<(E + 255) / 256, 256, 0, stream>>>(robot, out, E);
    return cudaGetLastError();
}
