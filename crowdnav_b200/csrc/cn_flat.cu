/*
 * cn_flat.cu -- the fused env-step kernel for sm_100a, "compacted work list" design.
 *
 * One launch advances E independent 2-D worlds by one control period (the same
 * path as cn_step.cu: CROWD:98-144 pedestrians -> FAKE:109-167 unicycle ->
 * XACRO:148-179 LiDAR -> UTL:375-392 scan cleaning -> ENV:568-907 perceived-risk
 * block -> ENV:246-265 waypoint, ENV:1011-1023 done, ENV:1046-1162 reward,
 * ENV:1227-1263 reset).  Results are bit-identical to cn_step.cu and to the
 * CPU oracle; only the mapping of work to threads differs.
 *
 * Why another mapping.  With one warp per world (cn_step.cu) most instructions
 * are issued for one or two useful lanes: two of twenty pedestrians resample
 * per step, one pedestrian in twenty is inside LiDAR range, one confirmed
 * object per world goes through the collision cone.  The ncu capture of that
 * kernel shows 1 620 warp-instructions per world at 20 of 32 lanes active, and
 * the kernel is issue-bound at 14 % of the HBM roofline.  Here a CTA owns a tile
 * of W consecutive worlds and every phase runs over a FLAT list of work items
 * gathered across the whole tile, so rare work is compacted into full warps:
 *
 *   phase 0   thread 0 issues the bulk TMA loads of the tile's three state
 *             planes + actions; meanwhile the threads clear the contact strips and
 *             (staged rows) pre-fill the observation rows with the no-return value.
 *   phase 1   warps 0-1: lane = WORLD, the pose-only part of get_state /
 *             compute_reward (unicycle, waypoint, heading, distance, reward
 *             shaping; warp 1 publishes the new pose early for the pedestrian
 *             side and then takes the wall faces, one lane per (world, face)).
 *             Other warps, concurrently: item = PEDESTRIAN of the tile: timers;
 *             strip-mask contact prefilter; Philox only for the compacted list
 *             of pedestrians that resample (or are re-spawned) this step;
 *             integrate + wall clamp into a second copy of the position plane
 *             (Jacobi on the old one), the few pedestrians in contact again as
 *             a compacted list; LiDAR candidate test on the spot; then
 *   phase 2   (same warps) item = candidate: bearing, angular span; every span
 *             (wall faces too) is cut into groups of <= 8 consecutive rays
 *             appended to a group list.
 *   phase 3   8 lanes per ray group: ray-disc / ray-face intersection with the
 *             oracle's per-ray arithmetic.  Whether the primitive OWNS the ray
 *             (oracle: walls first, then pedestrians in index order, strict <)
 *             is decided on the spot by intersecting the same ray with the few
 *             other primitives of that world whose span contains it -- no
 *             per-ray key array, no atomics on rays.  An owned ray gets its
 *             final cleaned, rounded value in the observation row; min(scan)
 *             and rays owned per pedestrian are accumulated.  The pedestrians'
 *             groups go first; the wall faces' groups are cast by warps 2.. while
 *   phase 5   warps 0-1, item = candidate: centre ray, hit point, tracker,
 *             collision cone, CP (ENV:656-860), tracker flag for the next step.
 *   phase 6   item = object: top-K rank + slot write; lane = world: counters,
 *             done, reward, robot record.
 *   phase 7   bulk TMA stores of the state planes and (staged rows) the [W, D]
 *             block of rows (and, for cn_step_gather, of the same block into
 *             every peer GPU).
 *
 * Two instances.  STAGED rows (DIRECT = 0): the tile's rows are assembled in
 * shared memory and leave by one bulk store -- reset launches, every fused-gather
 * entry point (the rows are pushed / encoded from shared memory), rows in
 * host-mapped memory.  DIRECT rows (DIRECT = 1, plain steps into device memory):
 * S.obs points at the caller's buffer; the no-return fill of the ray columns is
 * a few bulk stores per row from a small constant tile (warp 0, lane = row, as
 * soon as the state tile has landed; the TMA engine drains it under the
 * pedestrian phase), the owned rays / pose columns / K slots are ordinary global
 * stores ordered by the CTA barriers between the phases.  Without the row block
 * a world needs 2.0 instead of 3.8 KB of shared memory, so BASELINE configs[2]
 * (16 384 worlds) runs as ONE wave of 19-world CTAs instead of two residency
 * rounds: 23.8 instead of 30.1 us per step on a B200 (profiles/r02b).
 *
 * Numerics: everything that reaches an output goes through cn_math.h
 * primitives in the oracle's operation order; -fmad=false.  Approximate
 * arithmetic appears only in choosing (padded, conservative) span bounds.
 */
// This kernel runs every piece of code once per tile: cold loops stay rolled and every ray helper has one call site,
// which took the SASS from 144 KB to about 90 KB.  Turning the math helpers into real calls as well (CN_NOINLINE_BIG,
// 76 KB) measured 5 % slower (c2 18.3 vs 17.2 us, c3 37.7 vs 35.9 us) and is left off.
#define CN_COMPACT_CODE 1
#define CN_WALLS_BY_LANE 1      // pose_scalars leaves the wall faces to the kernel: one lane per (world, face)
#include "cn_dev.h"
#include <stdlib.h>
#include <string.h>

#define CF_POSE_WARPS 2
#ifndef CF_FILL_LATE
// 0: pose warp 0 issues the bulk fill of the ray columns as soon as the state tile has landed.  1: after it has advanced
// the robots (about 1 us later, so that the first CTAs' fill traffic does not compete with the bulk loads of the CTAs whose
// tile has not landed yet) -- measured slower at c3 (24.86 vs 23.87 us, same box), equal at c2 / c5.
#define CF_FILL_LATE 0
#endif
#define CF_RISK_WARPS 2      // warps that run E-J (phase 5) while the others cast the wall faces
#ifndef CN_FLAT_CTAS_PER_SM
// resident 256-thread CTAs per SM the kernel is compiled (register cap: 45 registers, no spills) and tiled for.
// Measured on B200 (profiles/r02/ctas_per_sm_ab.txt): 4 -> 5 is neutral at c2 / c3 and 10 % faster at c5; 6 (40 registers)
// leaves the 20-pedestrian configs with tiles too small for the lane = world warps.
#define CN_FLAT_CTAS_PER_SM 5
#endif

#ifdef CN_TIMELINE
// debug builds only: %globaltimer stamps per (CTA, warp), 16 slots each (profiles/tools/timeline_flat.py)
#define FSTAMP(k) do { if ((threadIdx.x & 31) == 0 && g_timeline) g_timeline[((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 16 + (k)] = gtime(); } while (0)
#else
#define FSTAMP(k) do { } while (0)
#endif

namespace {

// scalar-record words beyond the pose-phase record of cn_dev.h (S_* < S_WORDS = 40)
enum {
    F_MINBITS = 40,     // min over the cleaned scan, float bits (ranges are > 0)
    F_CONF0, F_CONF1,   // pedestrians confirmed as objects this step (bit n)
    F_XFLAGS,           // XF_*
    F_OVF_FACES,        // wall faces whose ray groups did not fit the group list
    F_NOBJ,             // objects of this world in the K block (entries of its object list)
    F_NCAND,            // LiDAR candidates of this world (entries of its candidate list)
    F_ETH,                                      // new robot pose + sensor offset, published early by pose warp 1 (= S_TH,
    F_EXI, F_EYI, F_EOFFX, F_EOFFY,             //  S_XI, S_YI, S_OFFX, S_OFFY); the last four are one aligned 16-byte quad
    F_WORDS = 52
};
#define XF_ACTIVE 1u    // the world is processed by this launch
#define XF_RESET  2u    // ... as a reset (MODE 1, or next-step auto-reset)
#define XF_EGO    4u    // an object's centre range < 0.140 (ENV:1000)

enum { C_NCAND = 0, C_NRES, C_NWG, C_NPG, C_OVF, C_NCON, C_NOBJ, C_WORDS = 8 };
static_assert((F_EXI % 4) == 0 && (F_WORDS % 4) == 0, "the published-pose quad must be 16-byte aligned");

// candidate record (8 words per pedestrian slot).  Phases 2-3: q (sensor-relative centre), bearing, span,
// owned-ray count, centre ray.  Phase 5 puts the object's CP row for phase 6 into the words nobody else reads
// (q and the span stay: another candidate's occluded-centre search may still need them).
enum { Q_QX = 0, Q_QY, Q_BEAR, Q_SPA, Q_SPB, Q_CNT, Q_MISC, Q_CKEY };
enum { O_CP = Q_BEAR, O_VX = Q_CNT, O_VY = Q_MISC, O_TTC = Q_CKEY };   // Q_CKEY: min centre-ray key over the owned rays
#define MARK_OVF 3           // S.mark[slot] of a candidate whose ray groups did not fit the group list
#define MARK_KIND 0x7Fu      // S.mark: low bits = what the pedestrian is this step (0 plain, 1 on the draw list, 2 untouched, MARK_OVF)
#define MARK_WAS_TRACKED 0x80u   // ... bit 7 = it was a tracked object when the step began (phase A moves the flag here)

// ray-group entry: up to 8 consecutive scan indices of one primitive
//   bits 0-2 count - 1, bits 3-13 first index, bits 14-31 primitive (pedestrian slot, or world * 4 + face)
#define GRP_NONE 0xFFFFFFFFu

struct Ptrs {
    uint32_t* robot; uint32_t* pa; uint32_t* pb; uint32_t* pa2; float* act; float* obs;
    uint32_t* sc; uint32_t* rec; uint32_t* pk; uint32_t* peers; uint16_t* clist; uint8_t* clw; uint16_t* rlist;
    uint16_t* olist; uint8_t* mark; uint32_t* wg; uint32_t* pg; uint32_t* cnt; uint64_t* bar; float* stage; uint32_t* strips;
    float* fillc; uint32_t* strips_y;
};
#define CF_FILLC_MIN 512u       // direct rows: smallest constant tile of "no return" values the bulk fill stores read

__device__ __forceinline__ void named_bar_sync(int id, int count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ int world_of(int idx, int N, uint32_t magic) {
    return (N == 1) ? idx : (int)__umulhi((uint32_t)idx, magic);
}
__device__ __forceinline__ void unpack_span(uint32_t a, uint32_t b, Span& s) {
    s.a0 = (int)(a & 0xFFFFu); s.a1 = (int)(a >> 16); s.b0 = (int)(b & 0xFFFFu); s.b1 = (int)(b >> 16);
}
__device__ __forceinline__ void wall_span(const uint32_t* sc, int face, Span& s) {
    s.a0 = (int)sc[S_WSPAN + 4 * face + 0]; s.a1 = (int)sc[S_WSPAN + 4 * face + 1];
    s.b0 = (int)sc[S_WSPAN + 4 * face + 2]; s.b1 = (int)sc[S_WSPAN + 4 * face + 3];
}
__device__ __forceinline__ int span_len_a(const Span& s) { return (s.a1 >= s.a0) ? (s.a1 - s.a0 + 1) : 0; }
__device__ __forceinline__ int span_len_b(const Span& s) { return (s.b1 >= s.b0) ? (s.b1 - s.b0 + 1) : 0; }
__device__ __forceinline__ bool in_span(const Span& s, int i) {
    return (i >= s.a0 && i <= s.a1) || (i >= s.b0 && i <= s.b1);
}
// ray `pos` of the concatenation [a0, a1] ++ [b0, b1]; -1 when past the end
__device__ __forceinline__ int span_ray(const Span& s, int pos) {
    const int la = span_len_a(s), lb = span_len_b(s);
    if (pos >= la + lb) return -1;
    return (pos < la) ? s.a0 + pos : s.b0 + (pos - la);
}

// append the ray groups of one primitive to a group list; returns false when they did not fit
__device__ __forceinline__ bool push_groups(uint32_t* list, uint32_t* counter, int cap, uint32_t tag, const Span& sp) {
    const int la = span_len_a(sp), lb = span_len_b(sp);
    const int ga = (la + 7) >> 3, gb = (lb + 7) >> 3;
    if (ga + gb == 0) return true;
    const int base = (int)atomicAdd(counter, (uint32_t)(ga + gb));
    if (base + ga + gb <= cap) {
        uint32_t* out = list + base;
        for (int k = 0; k < ga; ++k) {
            const int first = sp.a0 + 8 * k, cntm1 = min(8, sp.a1 - first + 1) - 1;
            out[k] = (tag << 14) | ((uint32_t)first << 3) | (uint32_t)cntm1;
        }
        out += ga;
        for (int k = 0; k < gb; ++k) {
            const int first = sp.b0 + 8 * k, cntm1 = min(8, sp.b1 - first + 1) - 1;
            out[k] = (tag << 14) | ((uint32_t)first << 3) | (uint32_t)cntm1;
        }
        return true;
    }
    for (int k = base; k < cap; ++k) list[k] = GRP_NONE;        // neutral entries; the primitive takes the slow path
    return false;
}
// scan index handled by this lane of a group, or -1
__device__ __forceinline__ int group_ray(uint32_t ent, int lane8) {
    if (ent == GRP_NONE || lane8 > (int)(ent & 7u)) return -1;
    return (int)((ent >> 3) & 0x7FFu) + lane8;
}

// ---- L: one ray against one primitive (XACRO:148-179; the oracle's lidar_raw arithmetic).  < 0: no return.
__device__ __forceinline__ float ped_t(const cn_kparams& P, float cqx, float cqy, float sn, float co) {
    const float b = fmaf(cqx, co, cqy * sn);
    const float h = fmaf(cqx, sn, -(cqy * co));
    const float disc = fmaf(-h, h, P.d.ped_r2);
    if (disc < 0.0f) return -1.0f;
    const float sq = sqrtf(disc);
    if (!(b + sq > 0.0f)) return -1.0f;                       // disc entirely behind
    float t = b - sq;
    if (t < 0.0f) t = 0.0f;                                   // sensor inside the disc
    return (t < P.max_range) ? t : -1.0f;
}
__device__ __forceinline__ float wall_t(const cn_kparams& P, const uint32_t* sc, int face, float sn, float co) {
    const bool xface = face < 2;
    const bool posf = (face & 1) == 0;                        // 0: +x, 1: -x, 2: +y, 3: -y
    const float den = xface ? co : sn;
    if (posf ? !(den > 0.0f) : !(den < 0.0f)) return -1.0f;
    const float wall = xface ? (posf ? P.room_xmax : P.room_xmin) : (posf ? P.room_ymax : P.room_ymin);
    const float o = xface ? (f_of(sc[S_XF]) + f_of(sc[S_OFFX])) : (f_of(sc[S_YF]) + f_of(sc[S_OFFY]));
    const float t = (wall - o) / den;
    return (t > 0.0f && t < P.max_range) ? t : -1.0f;
}

// A ray the primitive owns gets its final value: UTL:375-392 (sensor minimum) + np.around (ENV:1042); ENV:1012 min
// (returns the cleaned range's bit pattern: the caller folds the minimum over its ray group before touching the
//  world's min(scan) word -- one shared-memory atomic per group of 8 rays instead of one per ray)
__device__ __forceinline__ uint32_t finish_ray(const cn_kparams& P, const Ptrs& S, int w, int e, int j, float t, uint8_t hid) {
    const float rr = (t < P.sensor_min_range) ? P.sensor_min_range : t;
    // nobonus env: np.around of the whole row (ENV:1042); original env: Python round per ray (original:313)
    S.obs[(size_t)w * P.d.obs_dim + j] = (P.flags & CN_FLAG_ENV_ORIGINAL) ? cn_py_round3(rr) : cn_np_round3(rr);
    if (P.dbg_ranges) P.dbg_ranges[(size_t)e * (P.n_samples - 1) + j] = rr;
    if (P.dbg_hid) P.dbg_hid[(size_t)e * (P.n_samples - 1) + j] = hid;
    return u_of(rr);
}
// minimum over the 8 lanes of a ray group (every lane of the warp calls it)
__device__ __forceinline__ uint32_t group_min8(uint32_t v) {
    v = min(v, __shfl_xor_sync(FULL, v, 1));
    v = min(v, __shfl_xor_sync(FULL, v, 2));
    return min(v, __shfl_xor_sync(FULL, v, 4));
}

// Does pedestrian n of world w return ray i, and is it the primitive the oracle would report (walls first, then
// pedestrians in index order, strict <)?  The other primitives are tried only where their span contains the ray.
__device__ __forceinline__ bool ped_ray_eval(const cn_kparams& P, const Ptrs& S, int w, int n, int slot, int i, float& t_out,
                                             uint32_t& ang_out) {
    const int N = P.n_peds, NR = P.n_samples - 1;
    const uint32_t* rec = S.rec + slot * 8;
    const uint32_t* sc = S.sc + w * F_WORDS;
    const uint32_t ang = sc[S_TH] + (uint32_t)i * P.d.inc_bin;
    ang_out = ang;
    float sn, co; cn_sincos_bin(ang, &sn, &co);
    const float t = ped_t(P, f_of(rec[Q_QX]), f_of(rec[Q_QY]), sn, co);
    t_out = t;
    if (t < 0.0f) return false;
    bool own = true;
    const int j = NR - i;
    if ((sc[S_WDIRTY] >> min(j >> 5, 31)) & 1u) {
#pragma unroll 1
        for (int face = 0; face < 4; ++face) {
            Span ws; wall_span(sc, face, ws);
            if (!in_span(ws, i)) continue;
            const float tw = wall_t(P, sc, face, sn, co);
            if (tw >= 0.0f && tw <= t) own = false;             // the wall came first and the pedestrian is not nearer
        }
    }
    const int nc = (int)sc[F_NCAND];
    if (nc > 1) {
        const uint8_t* cl = S.clw + w * N;
#pragma unroll 1
        for (int k = 0; k < nc; ++k) {
            const int n2 = (int)cl[k];
            if (n2 == n) continue;
            const uint32_t* r2 = S.rec + (w * N + n2) * 8;
            Span s2; unpack_span(r2[Q_SPA], r2[Q_SPB], s2);
            if (!in_span(s2, i)) continue;
            const float t2 = ped_t(P, f_of(r2[Q_QX]), f_of(r2[Q_QY]), sn, co);
            if (t2 >= 0.0f && (t2 < t || (t2 == t && n2 < n))) own = false;
        }
    }
    return own;
}
// phase 3, one ray of a pedestrian's span; returns true when the pedestrian owns it
__device__ __forceinline__ bool cast_ped(const cn_kparams& P, const Ptrs& S, int e0, int w, int n, int slot, int i,
                                         uint32_t& rbits, uint32_t& ckey) {
    float t; uint32_t ang;
    if (!ped_ray_eval(P, S, w, n, slot, i, t, ang)) return false;
    rbits = finish_ray(P, S, w, e0 + w, (P.n_samples - 1) - i, t, (uint8_t)n);
    // centre ray (ENV:577 collapsed by ideal association): the owned ray nearest the pedestrian's centre line, in the
    // oracle's order -- smaller |ray angle - bearing| first (sign in the low bit).  Distinct rays have distinct keys
    // (adjacent rays are inc_bin >> 2 apart), so the minimum key identifies the ray; risk_candidate decodes it.
    const int32_t delta = (int32_t)(ang - S.rec[slot * 8 + Q_BEAR]);
    const uint32_t ad = (delta < 0) ? (0u - (uint32_t)delta) : (uint32_t)delta;
    ckey = (ad & ~1u) | (delta < 0 ? 1u : 0u);
    return true;
}
// phase 3, one ray of a wall face's span
__device__ __forceinline__ uint32_t cast_wall(const cn_kparams& P, const Ptrs& S, int e0, int q, int i) {
    const int N = P.n_peds;
    const int w = q >> 2, face = q & 3;
    const uint32_t* sc = S.sc + w * F_WORDS;
    float sn, co; cn_sincos_bin(sc[S_TH] + (uint32_t)i * P.d.inc_bin, &sn, &co);
    const float tw = wall_t(P, sc, face, sn, co);
    if (tw < 0.0f) return 0xFFFFFFFFu;
    const bool xface = face < 2;
    bool own = true;
#pragma unroll 1
    for (int k = 0; k < 2; ++k) {                               // the faces of the other axis (x faces come first)
        const int f2 = (xface ? 2 : 0) + k;
        Span ws; wall_span(sc, f2, ws);
        if (!in_span(ws, i)) continue;
        const float t2 = wall_t(P, sc, f2, sn, co);
        if (t2 >= 0.0f && (xface ? (t2 < tw) : (t2 <= tw))) own = false;
    }
    const int nc = (int)sc[F_NCAND];
    const uint8_t* cl = S.clw + w * N;
#pragma unroll 1
    for (int k = 0; k < nc; ++k) {
        const uint32_t* r2 = S.rec + (w * N + (int)cl[k]) * 8;
        Span s2; unpack_span(r2[Q_SPA], r2[Q_SPB], s2);
        if (!in_span(s2, i)) continue;
        const float t2 = ped_t(P, f_of(r2[Q_QX]), f_of(r2[Q_QY]), sn, co);
        if (t2 >= 0.0f && t2 < tw) own = false;
    }
    return own ? finish_ray(P, S, w, e0 + w, (P.n_samples - 1) - i, tw, CN_HIT_WALL) : 0xFFFFFFFFu;
}

// n 16-byte elements of shared memory, strided over the CTA: four predicated stores per trip (one trip at c2)
template <int T>
__device__ __forceinline__ void fill16(void* base, int n, uint32_t word, int tid) {
    uint32_t a = smem_u32(base) + (uint32_t)tid * 16u;
#pragma unroll 1
    for (int i = tid; i < n; i += 4 * T, a += 64u * T) {
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a), "r"(word) : "memory");
        if (i + T < n) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a + 16u * T), "r"(word) : "memory");
        if (i + 2 * T < n) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a + 32u * T), "r"(word) : "memory");
        if (i + 3 * T < n) asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(a + 48u * T), "r"(word) : "memory");
    }
}

// Direct rows, lane = row: "no return" into the row's ray columns [0, NR) -- the 16-byte-aligned interior by bulk stores
// from the constant tile (the TMA engine streams it, no thread waits), the at most three + three floats in front of and
// behind it by plain stores.  The pose columns and the K block are not touched (the pose warps write them, in any order
// with this).  The caller waits for the group (cp.async.bulk.wait_group 0) before barrier #A.
__device__ __forceinline__ void bulk_fill_row(float* row, int NR, float fill, const float* fillc, uint32_t fillc_bytes) {
    const uintptr_t a0 = reinterpret_cast<uintptr_t>(row), a1 = a0 + (uintptr_t)NR * 4u;
    uintptr_t b0 = (a0 + 15u) & ~(uintptr_t)15u, b1 = a1 & ~(uintptr_t)15u;
    if (b1 < b0) { b0 = a1; b1 = a1; }
#pragma unroll 1
    for (uintptr_t q = a0; q < b0; q += 4u) *reinterpret_cast<float*>(q) = fill;
#pragma unroll 1
    for (uintptr_t q = b1; q < a1; q += 4u) *reinterpret_cast<float*>(q) = fill;
#pragma unroll 1
    for (uintptr_t q = b0; q < b1; q += fillc_bytes)
        tma_store(reinterpret_cast<void*>(q), fillc, (uint32_t)min((uintptr_t)fillc_bytes, b1 - q));
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// n 16-byte elements from shared to global memory, strided over `nthreads` threads (fire-and-forget stores)
__device__ __forceinline__ void copy16_out(void* gdst, const void* ssrc, int n, int t, int nthreads) {
    uint4* g = reinterpret_cast<uint4*>(gdst);
    const uint4* s = reinterpret_cast<const uint4*>(ssrc);
#pragma unroll 1
    for (int i = t; i < n; i += nthreads) g[i] = s[i];
}

// ---- fused all-gather: system-scope signalling and NVSwitch multicast stores
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// (relaxed: the caller has just executed ONE system-scope fence for all of its signals -- fence + relaxed atomic is
//  the release pattern; a .release on every atomic would repeat the fence per peer)
__device__ __forceinline__ void red_relaxed_sys_add(unsigned long long* p, unsigned long long v) {
    asm volatile("red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void multimem_red_relaxed_sys_add(unsigned long long* p, unsigned long long v) {
    asm volatile("multimem.red.relaxed.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t;
}
// wait (bounded) until *ctr >= target; false = gave up
__device__ __forceinline__ bool wait_arrivals(const unsigned long long* ctr, unsigned long long target) {
    if (ld_acquire_sys(ctr) >= target) return true;
    const unsigned long long t0 = globaltimer_ns();
    while (ld_acquire_sys(ctr) < target) {
        __nanosleep(64);
        if (globaltimer_ns() - t0 > CN_GATHER_WAIT_NS) return false;
    }
    return true;
}
// Before a CTA overwrites a gather buffer on the peers: which step is this?  This rank's OWN slot counts the CTAs of
// its completed launches -- the last CTA of a launch adds the launch's CTA count at the very end (signal_peers) -- so
// t = own / CTAs-per-launch.  The buffer about to be overwritten was read by the peers before they launched step
// t - arrive_back, whose arrivals (slot s of OUR array; the calling thread polls source rank `slot`) we wait for;
// normally they are there long before.  Device-side counting keeps a captured graph of steps correct on every replay.
__device__ __forceinline__ void guard_peer_buffers(const cn_kparams& P, int slot) {
    if (P.arrive_back <= 0 || slot >= P.arrive_slots || slot == P.arrive_self) return;
    const unsigned long long own = ld_acquire_sys(P.arrive_local + P.arrive_self);
    const unsigned long long t = own / P.ctas_per_step;
    if (t < (unsigned long long)P.arrive_back) return;
    const unsigned long long target = (t - (unsigned long long)P.arrive_back + 1ull) * P.ctas_per_step;
    if (!wait_arrivals(P.arrive_local + slot, target) && P.gather_timeouts) atomicAdd(P.gather_timeouts, 1u);
}
// ... and after its rows have reached the peers.  ONE system-scope release per launch, not one per CTA (512 CTAs each
// fencing at system scope and hitting the peer's counter cost 17 us per step on two B200s -- profiles/r02): thread 0 of
// every CTA makes the CTA's stores (ordered before this call by a CTA barrier, or completed bulk stores of its own)
// performed with a GPU-scope fence and counts the CTA on a device-local word; the CTA that finds itself last fences
// at system scope and adds the whole launch's CTA count to this rank's slot on every peer and to its own slot.
// Causality: CTA i's stores -> its gpu-scope fence + atomic -> the last CTA's atomic (same scope) -> its sys-scope
// release -> the peer's acquire load.
__device__ __forceinline__ void signal_peers(const cn_kparams& P, int n_peers, int tid) {
    if (tid != 0) return;
    __threadfence();
    const unsigned int ctas = P.ctas_per_step;
    if (atomicAdd(P.gather_done, 1u) + 1u != ctas) return;
    *P.gather_done = 0u;                                 // the next launch starts after this one has completed
    __threadfence_system();
    if (P.arrive_mc) { multimem_red_relaxed_sys_add(P.arrive_mc, (unsigned long long)ctas); return; }   // own slot included
#pragma unroll 1
    for (int p = 0; p < n_peers; ++p) red_relaxed_sys_add(P.arrive_peers[p], (unsigned long long)ctas);
    red_relaxed_sys_add(P.arrive_local + P.arrive_self, (unsigned long long)ctas);
}

// 16-bit wire format of an observation value: thousandths as int16 (every row entry is a whole number of thousandths
// or hundredths by construction: cn_np_round3 / cn_py_round3 / cn_py_round2), -0.0 as -32768.  |k| <= 32767 survives
// fl(fl(k / 1000) * 1000) with an error far below 1/2, and cn_div1000 rebuilds the correctly rounded quotient, i.e. the
// original bits.
__device__ __forceinline__ uint32_t wire16_encode(float v, bool& saturated) {
    if (u_of(v) == 0x80000000u) return 0x8000u;
    float k = rintf(v * 1000.0f);
    if (!(fabsf(k) <= 32767.0f)) { saturated = true; k = (k < 0.0f) ? -32767.0f : 32767.0f; }
    return (uint32_t)(int32_t)k & 0xFFFFu;
}
__device__ __forceinline__ float wire16_decode(int32_t k) {
    return (k == -32768) ? f_of(0x80000000u) : cn_div1000((float)k);
}

// n 16-byte elements from shared memory to a multicast address: ONE store instruction per element, the switch
// replicates it into every rank's buffer
__device__ __forceinline__ void copy16_out_mc(void* mcdst, const void* ssrc, int n, int t, int nthreads) {
    const float4* s = reinterpret_cast<const float4*>(ssrc);
    float4* g = reinterpret_cast<float4*>(mcdst);
#pragma unroll 1
    for (int i = t; i < n; i += nthreads) {
        const float4 v = s[i];
        asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};"
                     ::"l"(g + i), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
    }
}
// the tile's [nE, D] block of rows into every peer's gather buffer.  CTA b starts with peer b mod n: at any moment
// the CTAs of one GPU are spread over all its NVLink destinations, and (with the host listing the peers from
// rank + 1 on) the GPUs do not gang up on one receiver.
__device__ __forceinline__ void push_rows_to_peers(const cn_kparams& P, const float* srows, size_t row0, int n16, int tid, int T) {
    if (P.obs_mc) { copy16_out_mc(P.obs_mc + row0, srows, n16, tid, T); return; }
    const int np = P.n_obs_peers;
    int p = (np > 1) ? (int)(blockIdx.x % (unsigned)np) : 0;
#pragma unroll 1
    for (int k = 0; k < np; ++k) {
        copy16_out(P.obs_peers[p] + row0, srows, n16, tid, T);
        p = (p + 1 == np) ? 0 : p + 1;
    }
}

// ---- phase 5, E-J for ONE candidate whose rays have all been cast (ENV:568-860): centre ray, hit point, tracker,
// collision cone, CP; the object joins its world's and the tile's object lists.
__device__ __forceinline__ void risk_candidate(const cn_kparams& P, const Ptrs& S, uint32_t magic, int slot) {
    const int N = P.n_peds, NR = P.n_samples - 1;
    uint32_t* rec = S.rec + slot * 8;
    if (rec[Q_CNT] < 4u) return;                                      // ENV:573: fewer than 4 rays is no object
    const int w = world_of(slot, N, magic), n = slot - w * N;
    uint32_t* sc = S.sc + w * F_WORDS;
    const uint32_t th = sc[S_TH];
    // decode the centre ray from its key: ray angle = bearing +- |delta|, scan index = angle / inc (exact to rounding)
    const uint32_t ckey = rec[Q_CKEY];
    const uint32_t adk = ckey & ~1u;
    const uint32_t rel2 = (rec[Q_BEAR] - th) + ((ckey & 1u) ? (0u - adk) : adk);
    const int istar = (int)fmaf((float)rel2, P.d.inv_inc_bin, 0.5f);
    const int jstar = NR - istar;
    float t_raw;
    {
        float sn, co; cn_sincos_bin(th + (uint32_t)istar * P.d.inc_bin, &sn, &co);
        t_raw = ped_t(P, f_of(rec[Q_QX]), f_of(rec[Q_QY]), sn, co);     // the same arithmetic as in phase 3
    }
    const float xf = f_of(sc[S_XF]), yf = f_of(sc[S_YF]);
    const float d_raw = (t_raw < P.sensor_min_range) ? P.sensor_min_range : t_raw;
    const float d3 = cn_py_round3(d_raw);                               // ENV:324,384
    float sa, ca; cn_sincos_bin((uint32_t)jstar * P.d.hit_inc_bin - th, &sa, &ca);     // C2: UTL:110-126
    const float hx = cn_py_round3(xf + d_raw * ca);
    const float hy = cn_py_round3(yf + (d_raw * sa) * -1.0f);
    // H/I: tracker with ideal association (ENV:656-760)
    float chx = 0.0f, chy = 0.0f, speed = -1.0f, ovx = 0.0f, ovy = 0.0f;
    if (S.mark[slot] & MARK_WAS_TRACKED) {                             // (the flag word itself was cleared by phase A)
        chx = f_of(S.pb[4 * slot + 0]) - hx; chy = f_of(S.pb[4 * slot + 1]) - hy;      // last - curr (sic), ENV:806-807
        speed = sqrtf(fmaf(chy, chy, chx * chx)) * P.d.inv_dt;
        ovx = chx * P.d.inv_dt; ovy = chy * P.d.inv_dt;
    }
    S.pb[4 * slot + 0] = u_of(hx); S.pb[4 * slot + 1] = u_of(hy);
    atomicOr(&sc[F_CONF0 + (n >> 5)], 1u << (n & 31));
    S.pb[4 * slot + 3] |= CN_PF_TRACKED;                                // tracked next step iff confirmed now (ENV:656-743, ideal association)
    if (d3 < 0.140f) atomicOr(&sc[F_XFLAGS], XF_EGO);                   // ENV:1000
    if (sc[F_XFLAGS] & XF_RESET) return;                              // ENV:769: no previous pose at step 0
    // J: collision cone (ENV:765-860, UTL:251-293 as a true ray-circle test)
    const float pcx = f_of(sc[S_PCX]), pcy = f_of(sc[S_PCY]);
    const float ppx = f_of(sc[S_PPX]), ppy = f_of(sc[S_PPY]);
    const float agent_vel = f_of(sc[S_AVEL]);
    const float tx = pcx + chx, ty = pcy + chy;
    float ux = tx - ppx, uy = ty - ppy;
    const float Ln = sqrtf(fmaf(ux, ux, uy * uy));
    bool have_dtc = false; float dtc = 0.0f;
    if (Ln > 0.0f) {
        const float invL = 1.0f / Ln;
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
            const float qq = (0.15f * resultant) / dtc;
            cp_ttc = (qq < 1.0f) ? qq : 1.0f;
        }
        cp = 0.5f * cp_ttc + 0.5f * dto;
    }
    rec[O_CP] = u_of(cp); rec[O_VX] = u_of(ovx); rec[O_VY] = u_of(ovy); rec[O_TTC] = u_of(cp_ttc);   // x, y: ped_b
    const uint32_t k = atomicAdd(&sc[F_NOBJ], 1u);                      // the world's object list (any order)
    S.olist[w * N + k] = (uint16_t)slot;
    S.rlist[atomicAdd(&S.cnt[C_NOBJ], 1u)] = (uint16_t)slot;            // ... and the tile's (the draw list is long gone)
}

// ------------------------------------------------------------------- kernel
// DIRECT = 1: the "direct rows" instance for plain single-GPU steps (cn_step / cn_step_n / library graphs).  The rows are
// not staged in shared memory: S.obs points at the caller's buffer, the no-return fill and the owned rays / pose columns /
// K slots are ordinary global stores (ordered by the CTA barriers between the phases; L2 merges them, DRAM sees each
// line once), nothing of the fused gather is compiled in, and the tile needs 4 D fewer bytes per world.
template <int MODE, int T, int DIRECT>
__global__ void __launch_bounds__(T, DIRECT ? ((T >= 512) ? 3 : (T >= 384 ? 4 : (T >= 256 ? 6 : (T >= 192 ? 8 : 12))))
                                            : ((T >= 512) ? 2 : (T >= 384 ? 3 : (T >= 256 ? CN_FLAT_CTAS_PER_SM : (T >= 192 ? 6 : 8)))))
cn_flat_kernel(const __grid_constant__ cn_kparams P, const __grid_constant__ cn_flat_layout L) {
    static_assert(!(DIRECT && MODE != 0), "direct rows: step launches only");
    constexpr int PED_THREADS = T - 32 * CF_POSE_WARPS;
    extern __shared__ __align__(128) uint8_t smem[];
    FSTAMP(14);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {                                     // first thing: the bulk loads below cannot be issued before this
        uint64_t* bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);
        mbar_init(bar, 1);
        if (L.off_stage != 0u) mbar_init(bar + 1, 1);
        fence_mbar_init();
    }
    const int W = L.W, N = P.n_peds, NR = P.n_samples - 1, D = P.d.obs_dim, K = P.k_obstacles;
    const int e0 = blockIdx.x * W;
    const int nE = min(W, P.n_envs - e0);

    Ptrs S;
    S.robot = reinterpret_cast<uint32_t*>(smem);
    S.pa = reinterpret_cast<uint32_t*>(smem + L.off_pa);
    S.pb = reinterpret_cast<uint32_t*>(smem + L.off_pb);
    S.pa2 = reinterpret_cast<uint32_t*>(smem + L.off_pa2);
    S.act = reinterpret_cast<float*>(smem + L.off_act);
    S.obs = DIRECT ? (P.obs + (size_t)e0 * D) : reinterpret_cast<float*>(smem + L.off_obs);
    S.sc = reinterpret_cast<uint32_t*>(smem + L.off_sc);
    S.rec = reinterpret_cast<uint32_t*>(smem + L.off_rec);
    S.pk = reinterpret_cast<uint32_t*>(smem + L.off_pk);
    S.peers = reinterpret_cast<uint32_t*>(smem + L.off_peers);
    S.clist = reinterpret_cast<uint16_t*>(smem + L.off_clist);
    S.clw = reinterpret_cast<uint8_t*>(smem + L.off_clw);
    S.rlist = reinterpret_cast<uint16_t*>(smem + L.off_rlist);
    S.olist = reinterpret_cast<uint16_t*>(smem + L.off_olist);
    S.mark = smem + L.off_mark;
    S.wg = reinterpret_cast<uint32_t*>(smem + L.off_wg);
    S.pg = reinterpret_cast<uint32_t*>(smem + L.off_pg);
    S.cnt = reinterpret_cast<uint32_t*>(smem + L.off_cnt);
    S.bar = reinterpret_cast<uint64_t*>(smem + L.off_bar);
    S.stage = reinterpret_cast<float*>(smem + L.off_stage);
    S.strips = reinterpret_cast<uint32_t*>(smem + L.off_strips);
    S.strips_y = reinterpret_cast<uint32_t*>(smem + L.off_strips_y);
    S.fillc = reinterpret_cast<float*>(smem + L.off_fillc);

    const int n_items = nE * N;                       // pedestrians of the tile
    const uint32_t rob_bytes = (uint32_t)nE * CN_ROBOT_WORDS * 4u;
    const uint32_t ped_bytes = (uint32_t)n_items * 16u;
    const bool act_smem = (MODE == 0) && P.act_bulk_ok && (W % 2) == 0 && (nE % 2) == 0;

    // ---------------------------------------------------------------- phase 0
    // Programmatic dependent launch (only when the launch carries the attribute, CN_PDL=1; otherwise both instructions
    // are no-ops): the NEXT launch on the stream (the next step) may start scheduling its CTAs as
    // soon as every CTA of this one has got here -- they take the SM slots this grid's CTAs free one by one -- while
    // this launch itself touches no global memory before the grid in front of it has completed and flushed
    // (griddepcontrol.wait).  What is hidden is the launch gap and the CTA start-up between back-to-back steps.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    // pipelined fused gather: the rows of the PREVIOUS step (in this rank's gather buffer) go to the peers under this
    // step's compute -- bulk load into the staging tile now, bulk stores to every peer as soon as it has landed
    const bool push = (MODE == 0) && !DIRECT && P.n_push_peers > 0;
    const uint32_t push_elem = P.push_wire16 ? 2u : 4u;                 // int16 thousandths or fp32
    const uint32_t push_bytes = (uint32_t)nE * (uint32_t)D * push_elem;
    const size_t push_off = (size_t)e0 * D * push_elem;                 // byte offset of the tile in a row block
    const bool push_bulk = push && L.off_stage != 0u && P.push_bulk_ok && (((size_t)W * D * push_elem) % 16 == 0) && (push_bytes % 16u == 0u);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) {
        if (push_bulk) {
            mbar_expect_tx(S.bar + 1, push_bytes);
            tma_load(S.stage, reinterpret_cast<const uint8_t*>(P.push_src) + push_off, push_bytes, S.bar + 1);
        }
        const uint32_t act_bytes = act_smem ? (uint32_t)nE * 8u : 0u;
        mbar_expect_tx(S.bar, rob_bytes + 2u * ped_bytes + act_bytes);
        if (act_bytes) tma_load(S.act, P.action + 2 * (size_t)e0, act_bytes, S.bar);
        tma_load(S.robot, P.robot + (size_t)e0 * CN_ROBOT_WORDS, rob_bytes, S.bar);
        if (ped_bytes) {
            tma_load(S.pa, P.ped_a + (size_t)e0 * N * 4, ped_bytes, S.bar);
            tma_load(S.pb, P.ped_b + (size_t)e0 * N * 4, ped_bytes, S.bar);
        }
        FSTAMP(15);
    }
    {
        {
            // every ray starts as "no return" (already rounded); the fill runs under the latency of the bulk loads.
            // (Tried: a bulk copy of a constant tile instead -- 15 % fewer instructions at c2 but the tile lands 0.3 us
            //  later and the step is latency-bound: 11.5 vs 11.2 us.)
            const float fill = P.d.max_range_r3;
            const int tot = nE * D;
            if (DIRECT) {
                // Direct rows: the "no return" fill of the ray columns is left to the TMA engine -- bulk stores from a
                // 512-byte constant tile, issued by warp 0 (lane = row) once the state tile has landed, draining under
                // the pedestrian phase.  (Measured at c3, one wave: the fill as plain stores from every thread here
                // saturates the L2 write path, the tile's bulk loads queue behind 26 MB of stores and land after 5.8 us
                // instead of 1.7; streamed by one warp it takes that warp 11 us -- per-warp store issue, not bandwidth.)
                if (warp == 0) {
#pragma unroll 1
                    for (int i = lane; i < (int)(L.fillc_bytes >> 4); i += 32)
                        reinterpret_cast<float4*>(S.fillc)[i] = make_float4(fill, fill, fill, fill);
                    fence_async_smem();
                }
            } else {
                const int n4 = tot >> 2;
                fill16<T>(S.obs, n4, u_of(fill), tid);
                if (tid < (tot & 3)) S.obs[(n4 << 2) + tid] = fill;
            }
        }
        {                                                                   // contact-prefilter strips: all empty
            const int n16 = (nE * ((int)L.strip_mask + 1) * (int)L.strip_words) >> 2;     // per axis
            uint4* zx = reinterpret_cast<uint4*>(S.strips);
            uint4* zy = reinterpret_cast<uint4*>(S.strips_y);
#pragma unroll 1
            for (int i = tid; i < 2 * n16; i += T) {
                if (i < n16) zx[i] = make_uint4(0u, 0u, 0u, 0u); else zy[i - n16] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        if (tid < C_WORDS) S.cnt[tid] = 0u;
        if (tid < nE) {
            uint32_t* sc = S.sc + tid * F_WORDS;
            sc[F_MINBITS] = u_of(P.max_range); sc[F_CONF0] = 0u; sc[F_CONF1] = 0u; sc[F_XFLAGS] = 0u; sc[F_OVF_FACES] = 0u;
            sc[F_NOBJ] = 0u; sc[F_NCAND] = 0u;
        }
        if (P.dbg_ranges || P.dbg_hid) {                                    // debug taps (tests only): "no return" everywhere
#pragma unroll 1
            for (int w = warp; w < nE; w += T / 32) {
                if (MODE == 1 && P.mask && P.mask[e0 + w] == 0) continue;
                const size_t base = (size_t)(e0 + w) * NR;
#pragma unroll 1
                for (int j = lane; j < NR; j += 32) {
                    if (P.dbg_ranges) P.dbg_ranges[base + j] = P.max_range;
                    if (P.dbg_hid) P.dbg_hid[base + j] = CN_HIT_NONE;
                }
            }
        }
    }
    // fused gather, pipelined: before the peers' buffers are touched by our pushes lane s of warp 0 checks source rank
    // s's progress (normally one L2 hit, under the fills)
    if (push && warp == 0 && !(L.gather_debug & 1)) guard_peer_buffers(P, lane);
    __syncthreads();            // fills done, barrier init visible
    mbar_wait(S.bar, 0);        // state tile + actions have landed
    FSTAMP(1);
    if (DIRECT && warp == 0 && !CF_FILL_LATE && lane < nE) bulk_fill_row(S.obs + (size_t)lane * D, NR, P.d.max_range_r3, S.fillc, L.fillc_bytes);
    if (push_bulk && warp == 0) {
        if (tid == 0 && !(L.gather_debug & 4)) {
            mbar_wait(S.bar + 1, 0);
            int p = (P.n_push_peers > 1) ? (int)(blockIdx.x % (unsigned)P.n_push_peers) : 0;   // CTAs start at different peers
#pragma unroll 1
            for (int k = 0; k < P.n_push_peers; ++k) {
                tma_store(reinterpret_cast<uint8_t*>(P.push_peers[p]) + push_off, S.stage, push_bytes);
                p = (p + 1 == P.n_push_peers) ? 0 : p + 1;
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }

    // ---------------------------------------------------------------- phase 1 (+ 2)
    if (warp < CF_POSE_WARPS) {
        // lane = world: everything get_state / compute_reward derive from the pose alone (cn_dev.h), shared by two
        // warps; warp 1 publishes the robot's new pose early for the pedestrian side and goes on
        const int part = warp;
        const int w = lane;
        const uint32_t* rob = S.robot + w * CN_ROBOT_WORDS;
        uint32_t* scl = S.sc + w * F_WORDS;
        float* rowl = S.obs + (size_t)w * D;
        bool run = false, reset_now = (MODE == 1);
        PoseIn p; p.xi = 0; p.yi = 0; p.th = 0u; p.v = 0.0f; p.w = 0.0f;
        int bad = 0;
        if (w < nE) {
            run = true;
            if (MODE == 1) run = !P.mask || P.mask[e0 + w] != 0;
            else reset_now = (rob[CN_R_FLAGS] & CN_RF_DONE) && (P.flags & CN_FLAG_AUTO_RESET);
            if (part == 0) scl[F_XFLAGS] = (run ? XF_ACTIVE : 0u) | ((run && reset_now) ? XF_RESET : 0u);
            if (run) {
                if (reset_now) { p.xi = P.d.start_xi; p.yi = P.d.start_yi; p.th = P.d.start_th; }
                else p = advance_robot(P, rob, act_smem ? S.act + 2 * w : P.action + 2 * (size_t)(e0 + w), bad,
                                       S.pa + (size_t)w * N * 4, N);
            }
        }
        if (DIRECT && part == 0 && CF_FILL_LATE && lane < nE) bulk_fill_row(rowl, NR, P.d.max_range_r3, S.fillc, L.fillc_bytes);
        if (part == 1) {
            if (run) {
                float sy, cy; cn_sincos_bin(p.th, &sy, &cy);
                scl[F_EXI] = (uint32_t)p.xi; scl[F_EYI] = (uint32_t)p.yi; scl[F_ETH] = p.th;
                scl[F_EOFFX] = u_of(P.mount_x * cy); scl[F_EOFFY] = u_of(P.mount_x * sy);
            }
            __threadfence_block();
            named_bar_arrive(2, 32 + PED_THREADS);
        }
        if (run) {
            // Z: Env.reset (ENV:1227-1263): spawn pose, waypoint = goal, unrounded previous_* (ENV:1243-1244); otherwise
            // the episode state carried in the robot record
            float wpx = P.goal_x, wpy = P.goal_y, pd, ph, ppx = 0.0f, ppy = 0.0f;
            int stepc = 0;
            if (reset_now) {
                const float xf = (float)p.xi * CN_GRID, yf = (float)p.yi * CN_GRID;
                pd = dist_to_wp(xf, yf, P.goal_x, P.goal_y);
                ph = heading_to_wp(P, xf, yf, cn_bin2rad(p.th), P.goal_x, P.goal_y);
            } else {
                wpx = f_of(rob[CN_R_WPX]); wpy = f_of(rob[CN_R_WPY]);
                pd = f_of(rob[CN_R_PDIST]); ph = f_of(rob[CN_R_PHEAD]);
                ppx = f_of(rob[CN_R_PPX]); ppy = f_of(rob[CN_R_PPY]);
                stepc = (int)rob[CN_R_STEP] + 1;
            }
            pose_scalars(P, p, part, wpx, wpy, pd, ph, ppx, ppy, stepc, !reset_now, !reset_now, bad, scl, rowl);
            if (part == 1) scl[S_WDIRTY] = 0u;
        }

        if (part == 1) {
            // the walls in range of the sensor, one lane per (world, face) instead of four faces in a row per world:
            // the rays that can see the face (span), the chunks of the row they touch, the span's ray groups
            const uint32_t run_mask = __ballot_sync(FULL, run);
            __syncwarp();                                                   // the lanes' scalar records
#pragma unroll 1
            for (int q = lane; q < 4 * nE; q += 32) {
                const int wq = q >> 2, face = q & 3;
                if (!((run_mask >> wq) & 1u)) continue;
                uint32_t* sq = S.sc + wq * F_WORDS;
                uint32_t chunks = 0u;
                const Span sp = face_span(P, f_of(sq[S_XF]) + f_of(sq[S_OFFX]), f_of(sq[S_YF]) + f_of(sq[S_OFFY]), sq[S_TH], face, chunks);
                sq[S_WSPAN + 4 * face + 0] = (uint32_t)sp.a0; sq[S_WSPAN + 4 * face + 1] = (uint32_t)sp.a1;
                sq[S_WSPAN + 4 * face + 2] = (uint32_t)sp.b0; sq[S_WSPAN + 4 * face + 3] = (uint32_t)sp.b1;
                if (chunks) atomicOr(&sq[S_WDIRTY], chunks);
                if (!push_groups(S.wg, &S.cnt[C_NWG], (int)L.cap_wg, (uint32_t)q, sp)) {
                    atomicOr(&sq[F_OVF_FACES], 1u << face);
                    S.cnt[C_OVF] = 1u;
                }
            }
        }
        // direct rows: the fill stores have long completed; make that formal before #A, behind which the owned rays
        // are written over them
        if (DIRECT && part == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    } else {
        // item = pedestrian of the tile (P: CROWD:98-144 + contact stand-in, Jacobi: old positions in pa, new in pa2)
        const int ptid = tid - 32 * CF_POSE_WARPS;
        uint4* spa4 = reinterpret_cast<uint4*>(S.pa);
        uint4* spa2_4 = reinterpret_cast<uint4*>(S.pa2);
        uint4* spb4 = reinterpret_cast<uint4*>(S.pb);
        const bool auto_reset = (P.flags & CN_FLAG_AUTO_RESET) != 0u;
        const float rr2 = P.ped_radius + P.ped_radius, rrob = P.ped_radius + P.robot_radius;
        const int32_t lim_i = (int32_t)((fmaxf(rr2, rrob) + P.rep_cutoff) * CN_INV_GRID) + 64;   // conservative contact box
        const uint32_t lim2 = 2u * (uint32_t)lim_i;
        uint16_t* slowlist = S.olist;                                       // the object list is not in use before phase 5

        // new position from the effective velocity, frictionless wall clamp, LiDAR candidate test (phase 2a)
        auto finish_ped = [&](int it, int w, int32_t x0, int32_t y0, float vx, float vy, float vex, float vey) {
            int32_t nx = x0 + cn_f2i((vex * P.dt) * CN_INV_GRID);
            int32_t ny = y0 + cn_f2i((vey * P.dt) * CN_INV_GRID);
            if (nx < P.d.ped_xmin) { nx = P.d.ped_xmin; if (vx < 0.0f) vx = 0.0f; }
            if (nx > P.d.ped_xmax) { nx = P.d.ped_xmax; if (vx > 0.0f) vx = 0.0f; }
            if (ny < P.d.ped_ymin) { ny = P.d.ped_ymin; if (vy < 0.0f) vy = 0.0f; }
            if (ny > P.d.ped_ymax) { ny = P.d.ped_ymax; if (vy > 0.0f) vy = 0.0f; }
            spa2_4[it] = make_uint4((uint32_t)nx, (uint32_t)ny, u_of(vx), u_of(vy));
            return make_int2(nx, ny);
        };
        auto cand_test = [&](int it, int w, int32_t nx, int32_t ny) {
            uint32_t* sc = S.sc + w * F_WORDS;
            const uint4 ep = *reinterpret_cast<const uint4*>(sc + F_EXI);   // EXI, EYI, EOFFX, EOFFY in one load
            const float qx = (float)(nx - (int32_t)ep.x) * CN_GRID - f_of(ep.z);
            const float qy = (float)(ny - (int32_t)ep.y) * CN_GRID - f_of(ep.w);
            if (fmaf(qx, qx, qy * qy) < P.d.cand_d2) {
                const uint32_t pos = atomicAdd(&S.cnt[C_NCAND], 1u);
                S.clist[pos] = (uint16_t)it;
                const uint32_t k = atomicAdd(&sc[F_NCAND], 1u);
                S.clw[w * N + k] = (uint8_t)(it - w * N);
            }
        };
        // Contact prefilter.  Each axis is cut into 32 strips (modulo) at least as wide as the contact box; strip masks
        // hold one bit per pedestrian of the world.  Whoever can be inside pedestrian n's box sits in one of the three
        // strips around n's on BOTH axes: (x strips) & (y strips) leaves the few pedestrians to test exactly -- instead
        // of N / 2 box tests per pedestrian.  Any superset of the true contacts gives the same result (a partner
        // outside the cut-off adds exactly nothing), and the masks come out symmetric like the pairwise test's.
        const int sw = (int)L.strip_words;                                  // words per strip mask: 1 (N <= 32) or 2
        const int sshift = P.d.strip_shift;
        const int smask = (int)L.strip_mask, nstrip = smask + 1;               // strips per axis (a power of two)
        auto strip_of = [&](int32_t c, int32_t origin) { return (int)(((uint32_t)(c - origin)) >> sshift) & smask; };
        auto contact_masks = [&](int w, int n, int32_t x0, int32_t y0, uint32_t& m0, uint32_t& m1) {
            const uint32_t* xm = S.strips + (size_t)w * nstrip * sw;
            const uint32_t* ym = S.strips_y + (size_t)w * nstrip * sw;
            const int sx = strip_of(x0, P.d.ped_xmin), sy = strip_of(y0, P.d.ped_ymin);
            const int xa = ((sx + smask) & smask) * sw, xb = sx * sw, xc = ((sx + 1) & smask) * sw;
            const int ya = ((sy + smask) & smask) * sw, yb = sy * sw, yc = ((sy + 1) & smask) * sw;
            uint32_t c0 = (xm[xa] | xm[xb] | xm[xc]) & (ym[ya] | ym[yb] | ym[yc]);
            uint32_t c1 = 0u;
            if (sw == 2) c1 = (xm[xa + 1] | xm[xb + 1] | xm[xc + 1]) & (ym[ya + 1] | ym[yb + 1] | ym[yc + 1]);
            if (n < 32) c0 &= ~(1u << n); else c1 &= ~(1u << (n - 32));
            m0 = 0u; m1 = 0u;
            const uint32_t bx = (uint32_t)x0 + (uint32_t)lim_i, by = (uint32_t)y0 + (uint32_t)lim_i;
            while (c0 | c1) {                                               // usually nobody
                int m;
                if (c0) { m = __ffs(c0) - 1; c0 &= c0 - 1; } else { m = __ffs(c1) + 31; c1 &= c1 - 1; }
                const uint2 o = *reinterpret_cast<const uint2*>(S.pa + 4 * (w * N + m));
                if ((bx - o.x) < lim2 && (by - o.y) < lim2) { if (m < 32) m0 |= 1u << m; else m1 |= 1u << (m - 32); }
            }
        };
        auto near_robot = [&](int w, int32_t x0, int32_t y0) {
            const uint32_t* rob = S.robot + w * CN_ROBOT_WORDS;
            const int32_t rxi = (int32_t)rob[CN_R_X], ryi = (int32_t)rob[CN_R_Y];
            return (uint32_t)(x0 - rxi + lim_i) < lim2 && (uint32_t)(y0 - ryi + lim_i) < lim2;
        };
        // pedestrian `it` has somebody inside its contact box: remember who, queue it for the repulsion pass
        auto to_slow_list = [&](int it, uint32_t m0, uint32_t m1) {
            S.peers[2 * it] = m0; S.peers[2 * it + 1] = m1;
            slowlist[atomicAdd(&S.cnt[C_NCON], 1u)] = (uint16_t)it;
        };

        // -- A: timers; who draws random numbers; every pedestrian enters its strips (old positions)
        for (int it = ptid; it < n_items; it += PED_THREADS) {
            const int w = world_of(it, N, L.magic_n), n = it - w * N;
            const uint32_t* rob = S.robot + w * CN_ROBOT_WORDS;
            bool active = true, respawn = (MODE == 1);
            if (MODE == 1) active = !P.mask || P.mask[e0 + w] != 0;
            else respawn = (rob[CN_R_FLAGS] & CN_RF_DONE) && auto_reset;
            if (!active) { spa2_4[it] = spa4[it]; S.mark[it] = 2; continue; }  // untouched world: state goes back as it came
            if (respawn) {
                const uint32_t pos = atomicAdd(&S.cnt[C_NRES], 1u);
                S.rlist[pos] = (uint16_t)(it | 0x8000);
                S.mark[it] = 1;
                continue;
            }
            uint8_t mk = 0;
            // (timer, flags) in one load.  A pedestrian is tracked next step iff it is confirmed as an object in THIS one:
            // the flag moves into the mark byte here, where phase 5 reads it, and phase 5 sets it again for what it
            // confirms -- instead of a pass over every pedestrian at the end that rewrites a flag which almost never changes
            const uint2 tf = *reinterpret_cast<const uint2*>(S.pb + 4 * it + 2);
            if (tf.y & CN_PF_TRACKED) { S.pb[4 * it + 3] = tf.y & ~CN_PF_TRACKED; mk = MARK_WAS_TRACKED; }
            int32_t tm = (int32_t)tf.x - CN_TICKS_PER_STEP;
            if (tm <= 0) {
                const uint32_t gid = (uint32_t)(P.env_id_offset + e0 + w);
                const int b = (P.n_behaviors == 1) ? 0 : (int)(gid % (uint32_t)P.n_behaviors);
                tm += P.beh_period[b];
                if (P.beh_kind[b] == CN_BEHAVIOR_RANDOM) {
                    const uint32_t pos = atomicAdd(&S.cnt[C_NRES], 1u);
                    S.rlist[pos] = (uint16_t)it;
                    mk |= 1;
                } else {
                    const float speed = P.beh_speed[b];
                    S.pa[4 * it + 2] = u_of(__ldg(&P.cfg->behavior_table[b][n][0]) * speed);
                    S.pa[4 * it + 3] = u_of(__ldg(&P.cfg->behavior_table[b][n][1]) * speed);
                }
            }
            S.pb[4 * it + 2] = (uint32_t)tm;
            S.mark[it] = mk;
            if (MODE == 0) {
                const uint2 a = *reinterpret_cast<const uint2*>(S.pa + 4 * it);
                const uint32_t bit = 1u << (n & 31);
                atomicOr(S.strips + (size_t)w * nstrip * sw + strip_of((int32_t)a.x, P.d.ped_xmin) * sw + (n >> 5), bit);
                atomicOr(S.strips_y + (size_t)w * nstrip * sw + strip_of((int32_t)a.y, P.d.ped_ymin) * sw + (n >> 5), bit);
            }
        }
        // contact masks and the draw list are complete; pose warp 1 has published the robot's new pose.
        // (Tried: a pedestrian-only barrier here and the candidate tests as a separate pass after the second barrier,
        //  so that nobody waits for pose warp 1 -- 0.9 us per warp at c3: slower, c2 11.04 -> 11.58 us, c3 24.44 ->
        //  24.94 us.  The phase is bound by the instructions issued per SM, not by a warp's critical path: a waiting warp
        //  costs nothing, the extra pass does.)
        FSTAMP(4);
        named_bar_sync(2, 32 + PED_THREADS);
        FSTAMP(0);

        // -- B: pedestrians with nothing special: integrate, candidate test.  Contacts go to the slow list.
        if (MODE == 0) {
            for (int it = ptid; it < n_items; it += PED_THREADS) {
                if ((S.mark[it] & MARK_KIND) != 0) continue;
                const int w = world_of(it, N, L.magic_n);
                const uint4 a = spa4[it];
                uint32_t m0, m1;
                contact_masks(w, it - w * N, (int32_t)a.x, (int32_t)a.y, m0, m1);
                if ((m0 | m1) != 0u || near_robot(w, (int32_t)a.x, (int32_t)a.y)) { to_slow_list(it, m0, m1); continue; }
                const int2 np = finish_ped(it, w, (int32_t)a.x, (int32_t)a.y, f_of(a.z), f_of(a.w), f_of(a.z), f_of(a.w));
                cand_test(it, w, np.x, np.y);
            }
        }
        FSTAMP(3);
        //    ... and the compacted list of pedestrians that draw this step: Philox, then the same
        //    (entry q goes to thread PED_THREADS - 1 - q: the high threads have no pedestrian of their own when the
        //     tile has fewer pedestrians than threads, so the draws start at once instead of after a plain item)
        {
            const int n_res = (int)S.cnt[C_NRES];
            for (int q = PED_THREADS - 1 - ptid; q < n_res; q += PED_THREADS) {
                const uint32_t ent = S.rlist[q];
                const int idx = (int)(ent & 0x7FFFu);
                const bool respawn = (ent & 0x8000u) != 0u;
                const int ww = world_of(idx, N, L.magic_n), nn = idx - ww * N;
                const uint32_t* rob = S.robot + ww * CN_ROBOT_WORDS;
                const uint32_t gid = (uint32_t)(P.env_id_offset + e0 + ww);
                const int bb = (P.n_behaviors == 1) ? 0 : (int)(gid % (uint32_t)P.n_behaviors);
                const uint32_t episode = rob[CN_R_EPISODE] + (respawn ? 1u : 0u);
                const uint32_t stepc = respawn ? 0u : rob[CN_R_STEP] + 1u;
                const cn_u32x2 rnd = cn_env_rand(P.d.seed_lo, P.d.seed_hi, gid, episode, stepc, (uint32_t)nn, respawn ? 1u : 0u);
                if (respawn) {
                    // gazebo/reset_simulation: world-file layout (+ seeded jitter), first command after (n+1) staggers
                    const float px = __ldg(&P.cfg->ped_layout[nn][0]) + cn_usym(rnd.v[0], P.layout_jitter);
                    const float py = __ldg(&P.cfg->ped_layout[nn][1]) + cn_usym(rnd.v[1], P.layout_jitter);
                    int32_t xi = cn_f2i(px * CN_INV_GRID), yi = cn_f2i(py * CN_INV_GRID);
                    xi = max(xi, P.d.ped_xmin); xi = min(xi, P.d.ped_xmax);
                    yi = max(yi, P.d.ped_ymin); yi = min(yi, P.d.ped_ymax);
                    spa2_4[idx] = make_uint4((uint32_t)xi, (uint32_t)yi, u_of(0.0f), u_of(0.0f));
                    spb4[idx] = make_uint4(0u, 0u, (uint32_t)((nn + 1) * P.beh_stagger[bb]), 0u);
                    cand_test(idx, ww, xi, yi);
                } else {
                    const float speed = P.beh_speed[bb];
                    const float vx = cn_usym(rnd.v[0], speed), vy = cn_usym(rnd.v[1], speed);
                    const uint2 a = *reinterpret_cast<const uint2*>(S.pa + 4 * idx);
                    uint32_t m0, m1;
                    contact_masks(ww, nn, (int32_t)a.x, (int32_t)a.y, m0, m1);
                    if ((m0 | m1) != 0u || near_robot(ww, (int32_t)a.x, (int32_t)a.y)) {
                        S.pa[4 * idx + 2] = u_of(vx); S.pa[4 * idx + 3] = u_of(vy);
                        to_slow_list(idx, m0, m1);
                    } else {
                        const int2 np = finish_ped(idx, ww, (int32_t)a.x, (int32_t)a.y, vx, vy, vx, vy);
                        cand_test(idx, ww, np.x, np.y);
                    }
                }
            }
        }
        FSTAMP(11);
        named_bar_sync(1, PED_THREADS);

        // -- C: the few pedestrians with somebody inside the contact box: repulsion (index order, like the oracle)
        if (MODE == 0) {
            const int n_con = (int)S.cnt[C_NCON];
            for (int q = ptid; q < n_con; q += PED_THREADS) {
                const int it = (int)slowlist[q];
                const int w = world_of(it, N, L.magic_n);
                const uint32_t* rob = S.robot + w * CN_ROBOT_WORDS;
                const uint4 a = spa4[it];
                const int32_t x0 = (int32_t)a.x, y0 = (int32_t)a.y;
                float vex = f_of(a.z), vey = f_of(a.w);
                const int32_t rxi = (int32_t)rob[CN_R_X], ryi = (int32_t)rob[CN_R_Y];
                uint32_t p0 = S.peers[2 * it], p1 = S.peers[2 * it + 1];
                bool robot_pending = (uint32_t)(x0 - rxi + lim_i) < lim2 && (uint32_t)(y0 - ryi + lim_i) < lim2;
                while (p0 | p1 | (robot_pending ? 1u : 0u)) {               // pedestrians in index order, then the robot
                    int32_t ox = rxi, oy = ryi;
                    float rsum = rrob;
                    if (p0 | p1) {
                        int m;
                        if (p0) { m = __ffs(p0) - 1; p0 &= p0 - 1; } else { m = __ffs(p1) + 31; p1 &= p1 - 1; }
                        const uint2 o = *reinterpret_cast<const uint2*>(S.pa + 4 * (w * N + m));
                        ox = (int32_t)o.x; oy = (int32_t)o.y; rsum = rr2;
                    } else {
                        robot_pending = false;
                    }
                    add_rep(P, x0, y0, ox, oy, rsum, vex, vey);
                }
                const int2 np = finish_ped(it, w, x0, y0, f_of(a.z), f_of(a.w), vex, vey);
                cand_test(it, w, np.x, np.y);
            }
            // every new position is in pa2, the candidate lists are complete, and nobody needs pa / peers any more
            // (the candidate records reuse that memory)
            named_bar_sync(1, PED_THREADS);
        }
        FSTAMP(13);

        // ------------------------------------------------------------ phase 2b: bearing, span, ray groups
        const int n_cand_p = (int)S.cnt[C_NCAND];
        for (int q = ptid; q < n_cand_p; q += PED_THREADS) {
            const int slot = (int)S.clist[q];
            const int w = world_of(slot, N, L.magic_n);
            const uint32_t* sc = S.sc + w * F_WORDS;
            uint32_t* rec = S.rec + slot * 8;
            const uint2 a = *reinterpret_cast<const uint2*>(S.pa2 + 4 * slot);
            const uint4 ep = *reinterpret_cast<const uint4*>(sc + F_EXI);
            const float qx = (float)((int32_t)a.x - (int32_t)ep.x) * CN_GRID - f_of(ep.z);
            const float qy = (float)((int32_t)a.y - (int32_t)ep.y) * CN_GRID - f_of(ep.w);
            const float d2 = fmaf(qx, qx, qy * qy);
            const uint32_t bearing = cn_rad2bin(cn_atan2(qy, qx));
            float alpha = 4.0f;                                    // sensor inside / touching the disc: all rays
            const float rlim = P.ped_radius * 1.001f;
            if (d2 > rlim * rlim) {
                const float u = P.ped_radius * rsqrtf(d2) * 1.0001f;   // asin(u) <= u + (pi/2 - 1) u^3
                alpha = u * fmaf(0.5708f * u, u, 1.0f) + 0.01f;
            }
            const Span sp = make_span(P, bearing - sc[F_ETH], alpha);
            rec[Q_QX] = u_of(qx); rec[Q_QY] = u_of(qy);
            rec[Q_BEAR] = bearing;
            rec[Q_SPA] = (uint32_t)sp.a0 | ((uint32_t)sp.a1 << 16);
            rec[Q_SPB] = (uint32_t)sp.b0 | ((uint32_t)sp.b1 << 16);
            rec[Q_CNT] = 0u;
            rec[Q_CKEY] = 0xFFFFFFFFu;
            rec[Q_MISC] = 0u;
            if (!push_groups(S.pg, &S.cnt[C_NPG], (int)L.cap_pg, (uint32_t)slot, sp)) {
                S.mark[slot] = (uint8_t)((S.mark[slot] & MARK_WAS_TRACKED) | MARK_OVF);   // walked directly in phase 3
                S.cnt[C_OVF] = 1u;
            }
        }
    }
    FSTAMP(9);
    __syncthreads();            // #A: scalar records, new pedestrian positions, candidate lists
    FSTAMP(2);
    const int n_cand = (int)S.cnt[C_NCAND];
    const int n_wg = min((int)S.cnt[C_NWG], (int)L.cap_wg);
    const int n_pg = min((int)S.cnt[C_NPG], (int)L.cap_pg);
    const bool overflow = S.cnt[C_OVF] != 0u;
    const int lane8 = tid & 7;

    // ---------------------------------------------------------------- phase 3: cast + ownership (L, C: XACRO:148-179, UTL:375-392)
    // source -1 is the group list; a source >= 0 is a primitive whose groups did not fit and is walked directly
    // (one call site for both keeps the code small)
    // The pedestrians' ray groups first, by everybody; then, behind barrier #E, the first CF_RISK_WARPS warps run E-J for
    // the candidates (one short list, one long dependent chain) WHILE the other warps cast the wall faces' groups: the
    // chain of phase 5 -- 1.4 us with one busy warp per CTA -- disappears under the wall rays.  (Walls do not need the
    // candidates' E-J results and E-J does not need the walls' rays: ownership tests intersect the other primitives on
    // the spot, and phase 5 leaves q and the spans of the candidate records alone.)
#pragma unroll 1
    for (int src = -1; src < (overflow ? n_cand : 0); ++src) {
        int n_groups = n_pg, slot_src = 0;
        Span sp; sp.a0 = 1; sp.a1 = 0; sp.b0 = 1; sp.b1 = 0;
        if (src >= 0) {
            slot_src = (int)S.clist[src];
            if ((S.mark[slot_src] & MARK_KIND) != MARK_OVF) continue;
            const uint32_t* rec = S.rec + slot_src * 8;
            unpack_span(rec[Q_SPA], rec[Q_SPB], sp);
            n_groups = (span_len_a(sp) + span_len_b(sp) + 7) >> 3;
        }
#pragma unroll 1
        for (int g0 = warp * 4; g0 < n_groups; g0 += (T / 32) * 4) {        // warp-uniform trip count: shuffles inside
            const int g = g0 + (lane >> 3);
            bool owned = false, counted = false;
            int slot = slot_src, w = 0;
            uint32_t rbits = 0xFFFFFFFFu, ckey = 0xFFFFFFFFu;
            if (g < n_groups) {
                int i;
                if (src < 0) { const uint32_t ent = S.pg[g]; i = group_ray(ent, lane8); slot = (int)(ent >> 14); counted = ent != GRP_NONE; }
                else { i = span_ray(sp, g * 8 + lane8); counted = true; }
                if (counted) {
                    w = world_of(slot, N, L.magic_n);
                    if (i >= 0) owned = cast_ped(P, S, e0, w, slot - w * N, slot, i, rbits, ckey);
                }
            }
            const uint32_t bm = __ballot_sync(FULL, owned);
            const int c = __popc((bm >> (lane & 24)) & 0xFFu);
            rbits = group_min8(rbits);
            ckey = group_min8(ckey);
            if (lane8 == 0 && c) {
                // one thread per group folds the group's results into the candidate record: three shared-memory
                // atomics per group of 8 rays instead of two per ray
                uint32_t* rec = S.rec + slot * 8;
                atomicAdd(&rec[Q_CNT], (uint32_t)c);
                atomicMin(&rec[Q_CKEY], ckey);
                atomicMin(S.sc + w * F_WORDS + F_MINBITS, rbits);
            }
        }
    }
    FSTAMP(10);
    __syncthreads();            // #E: rows' ray part final, owned-ray counts final
    FSTAMP(5);

    // ---------------------------------------------------------------- phase 5: E-J per candidate (ENV:568-860)
    // (compacted: item = candidate.  Tried: the thread that casts a candidate's last ray group goes straight on with
    //  E-J and barrier #F goes away -- the chain then runs once per candidate with ONE active lane instead of once per
    //  tile with one lane per candidate: c3 29.7 -> 33.6 us, c2 no better; profiles/r02/bench_*_v10.json)
    if (warp < CF_RISK_WARPS) {
        for (int q = tid; q < n_cand; q += 32 * CF_RISK_WARPS) risk_candidate(P, S, L.magic_n, (int)S.clist[q]);
    } else {
#pragma unroll 1
        for (int src = -1; src < (overflow ? nE * 4 : 0); ++src) {
            int n_groups = n_wg;
            Span sp; sp.a0 = 1; sp.a1 = 0; sp.b0 = 1; sp.b1 = 0;
            if (src >= 0) {
                const uint32_t* sc = S.sc + (src >> 2) * F_WORDS;
                if (!((sc[F_OVF_FACES] >> (src & 3)) & 1u)) continue;
                wall_span(sc, src & 3, sp);
                n_groups = (span_len_a(sp) + span_len_b(sp) + 7) >> 3;
            }
#pragma unroll 1
            for (int g0 = (warp - CF_RISK_WARPS) * 4; g0 < n_groups; g0 += (T / 32 - CF_RISK_WARPS) * 4) {   // warp-uniform trip count: shuffles inside
                const int g = g0 + (lane >> 3);
                int q = src;
                uint32_t rbits = 0xFFFFFFFFu;
                if (g < n_groups) {
                    int i;
                    if (src < 0) { const uint32_t ent = S.wg[g]; i = group_ray(ent, lane8); q = (int)(ent >> 14); }
                    else i = span_ray(sp, g * 8 + lane8);
                    if (i >= 0) rbits = cast_wall(P, S, e0, q, i);
                }
                rbits = group_min8(rbits);                                      // ENV:1012: min over the scan, one atomic per group
                if (lane8 == 0 && rbits != 0xFFFFFFFFu) atomicMin(S.sc + (q >> 2) * F_WORDS + F_MINBITS, rbits);
            }
        }
    }
    __syncthreads();            // #F
    FSTAMP(6);

    // ---------------------------------------------------------------- phase 6
    // 6a  K: top-K block (ENV:862-907): stable rank by CP among the world's objects, keep [-K:]; padding is in the row
    //     (item = object of the tile's compacted object list)
    constexpr int T6 = T - 32;                                              // the last warp does 6c meanwhile
    const int n_obj_tile = (int)S.cnt[C_NOBJ];
    for (int q = tid; q < n_obj_tile && tid < T6; q += T6) {
        const int slot = (int)S.rlist[q];
        const int w = world_of(slot, N, L.magic_n), n = slot - w * N;
        const int n_obj = (int)S.sc[w * F_WORDS + F_NOBJ];
        const uint16_t* ol = S.olist + w * N;
        const uint32_t* rec = S.rec + slot * 8;
        const float my_cp = f_of(rec[O_CP]);
        int rank = 0;
        for (int k2 = 0; k2 < n_obj; ++k2) {
            const int s2 = (int)ol[k2];
            const int n2 = s2 - w * N;
            const float cpb = f_of(S.rec[s2 * 8 + O_CP]);
            if (n2 != n && (cpb > my_cp || (cpb == my_cp && n2 < n))) ++rank;
        }
        const int slot_k = (P.flags & CN_FLAG_TOPK_HIGHEST) ? rank : rank - (n_obj > K ? n_obj - K : 0);
        if (slot_k < 0 || slot_k >= K) continue;
        float* blk = S.obs + (size_t)w * D + NR + 7 + 4 * slot_k;
        blk[0] = f_of(S.pb[4 * slot + 0]); blk[1] = f_of(S.pb[4 * slot + 1]);      // hit point: already multiples of 0.001
        blk[2] = cn_np_round3(f_of(rec[O_VX])); blk[3] = cn_np_round3(f_of(rec[O_VY]));
    }
    // (6b, the tracker flags of every pedestrian, is gone: phase A clears the flag of what was tracked, phase 5 sets it for
    //  what it confirms)
    // 6c  lane = world (last warp): M counters, N done, W terminal reward, robot record
    if (warp == T / 32 - 1) {
        for (int w = lane; w < nE; w += 32) {
            uint32_t* rob = S.robot + w * CN_ROBOT_WORDS;
            const uint32_t* sc = S.sc + w * F_WORDS;
            const uint32_t xfl = sc[F_XFLAGS];
            if (!(xfl & XF_ACTIVE)) continue;
            const int e = e0 + w;
            uint32_t flags = rob[CN_R_FLAGS], cnt0 = rob[CN_R_CNT0], cnt1 = rob[CN_R_CNT1];
            uint32_t episode = rob[CN_R_EPISODE];
            int step = (int)rob[CN_R_STEP];
            const bool reset_now = (xfl & XF_RESET) != 0u;
            if (reset_now) episode += 1u;
            if (!reset_now && sc[S_BAD]) { uint32_t bad = cnt1 >> 16; if (bad < 0xFFFFu) ++bad; cnt1 = (cnt1 & 0xFFFFu) | (bad << 16); }
            const int n_seen = __popc(sc[F_CONF0]) + __popc(sc[F_CONF1]);
            if (n_seen > 0) {                                               // M: ENV:653-654, 998-1005
                float ego_score = 0.0f, emax = -INFINITY;
                const int n_obj = (int)sc[F_NOBJ];
#pragma unroll 1
                for (int k2 = 0; k2 < n_obj; ++k2) emax = fmaxf(emax, f_of(S.rec[(int)S.olist[w * N + k2] * 8 + O_TTC]));
                if (n_obj > 0) ego_score = emax;                            // ENV:879
                uint32_t ego = cnt0 & 0xFFFFu, soc = cnt0 >> 16;
                uint32_t pres = cnt1 & 0xFFFFu;
                if (pres < 0xFFFFu) ++pres;
                if ((xfl & XF_EGO) && ego < 0xFFFFu) ++ego;
                if (ego_score > 0.4f && soc < 0xFFFFu) ++soc;
                cnt0 = ego | (soc << 16);
                cnt1 = pres | (cnt1 & 0xFFFF0000u);
            }
            const bool collided = f_of(sc[F_MINBITS]) < P.collision_range;  // ENV:1012
            if (reset_now) {
                cnt0 = 0; cnt1 = 0;                                                 // ENV:1260-1262
                flags = (MODE == 0) ? (flags & (CN_RF_SUCCESS | CN_RF_FAILURE)) : 0u;   // last episode's status stays readable
                step = 0;
                if (MODE == 0) { P.reward[e] = 0.0f; P.done[e] = 2; }
            } else {
                // N + W: done (ENV:1011-1023), terminal reward (ENV:1136-1159; a time-out is -200 too)
                const uint32_t pre = sc[S_PRE];
                const bool done = ((flags & CN_RF_DONE) != 0) || collided || pre != 0u;
                int reward = (int)sc[S_REWARD];
                if (done) {
                    flags |= CN_RF_DONE;
                    if (pre & 1u) { flags |= CN_RF_SUCCESS; flags &= ~CN_RF_FAILURE; reward += 200; }
                    else { flags |= CN_RF_FAILURE; flags &= ~CN_RF_SUCCESS; reward -= 200; }
                }
                step += 1;
                P.reward[e] = (float)reward; P.done[e] = done ? 1 : 0;
            }
            uint4* qd = reinterpret_cast<uint4*>(rob);
            qd[0] = make_uint4(sc[S_XI], sc[S_YI], sc[S_TH], sc[S_V]);
            qd[1] = make_uint4(sc[S_W], sc[S_WPX], sc[S_WPY], sc[S_NPDIST]);
            qd[2] = make_uint4(sc[S_NPHEAD], sc[S_PCX], sc[S_PCY], (uint32_t)step);
            qd[3] = make_uint4(episode, flags, cnt0, cnt1);
        }
    }

    // ---------------------------------------------------------------- phase 7: write-back
    fence_async_smem();          // generic-proxy writes -> visible to the async proxy
    const bool gather = (MODE == 0) && !DIRECT && (P.n_obs_peers > 0 || P.obs_mc != nullptr);
    if (gather && tid < 32 && !(L.gather_debug & 1)) guard_peer_buffers(P, tid);       // (a pushing launch checked at its start)
    FSTAMP(12);
    __syncthreads();             // #G
    FSTAMP(7);
    // (direct rows: the rows are already where they belong -- "bulk_obs" only switches the row copies below off)
    const bool bulk_obs = DIRECT || ((MODE == 0) && P.obs_bulk_ok && (((size_t)W * D) % 4 == 0) && (((size_t)nE * D) % 4 == 0));
    if (MODE == 0 && !DIRECT && P.wire_out != nullptr) {
        // 16-bit wire copy of the tile's finished rows into this rank's own wire buffer (the next step's kernel forwards
        // it to the peers): two values per 4-byte store, fire and forget
        const int n = nE * D;
        int16_t* wout = P.wire_out + (size_t)e0 * D;
        bool sat = false;
        if (((n | (e0 * D)) & 1) == 0) {
            uint32_t* w32 = reinterpret_cast<uint32_t*>(wout);
#pragma unroll 2
            for (int i = tid; i < (n >> 1); i += T) {
                const float2 v = *reinterpret_cast<const float2*>(S.obs + 2 * i);
                w32[i] = wire16_encode(v.x, sat) | (wire16_encode(v.y, sat) << 16);
            }
        } else {
#pragma unroll 1
            for (int i = tid; i < n; i += T) wout[i] = (int16_t)wire16_encode(S.obs[i], sat);
        }
        if (sat && P.gather_timeouts) atomicAdd(P.gather_timeouts + 2, 1u);
    }
    if (L.plain_store) {
        // cooperative 16-byte stores: nothing to wait for, the CTA's slot is free as soon as they are issued
        if (bulk_obs && gather && !(L.gather_debug & 4)) push_rows_to_peers(P, S.obs, (size_t)e0 * D, (nE * D) >> 2, tid, T);   // NVLink first
        copy16_out(P.robot + (size_t)e0 * CN_ROBOT_WORDS, S.robot, (int)(rob_bytes >> 4), tid, T);
        if (ped_bytes) {
            copy16_out(P.ped_a + (size_t)e0 * N * 4, S.pa2, n_items, tid, T);
            copy16_out(P.ped_b + (size_t)e0 * N * 4, S.pb, n_items, tid, T);
        }
        if (bulk_obs && !DIRECT) copy16_out(P.obs + (size_t)e0 * D, S.obs, (nE * D) >> 2, tid, T);
    } else {
        if (tid == 0) {
            tma_store(P.robot + (size_t)e0 * CN_ROBOT_WORDS, S.robot, rob_bytes);
            if (ped_bytes) {
                tma_store(P.ped_a + (size_t)e0 * N * 4, S.pa2, ped_bytes);
                tma_store(P.ped_b + (size_t)e0 * N * 4, S.pb, ped_bytes);
            }
            if (bulk_obs && !DIRECT) tma_store(P.obs + (size_t)e0 * D, S.obs, (uint32_t)((size_t)nE * D * 4));
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        // fused all-gather: the same block of rows goes straight into every peer's gather buffer over NVLink, as plain
        // 16-byte stores from all threads (measured at 2 GPUs: 32.4 us per step against 36.7 us with bulk stores, whose
        // completion the CTA would have to wait for), or as multimem stores the switch replicates
        if (bulk_obs && gather && !(L.gather_debug & 4)) push_rows_to_peers(P, S.obs, (size_t)e0 * D, (nE * D) >> 2, tid, T);
        if (tid == 0) {
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            FSTAMP(8);
        }
    }
    if (!bulk_obs) {             // plain coalesced stores (reset launches, unaligned or ragged tiles)
#pragma unroll 1
        for (int w = warp; w < nE; w += T / 32) {
            if (!(S.sc[w * F_WORDS + F_XFLAGS] & XF_ACTIVE)) continue;
            const float* row = S.obs + (size_t)w * D;
#pragma unroll 1
            for (int p = -1; p < (P.obs_mc ? 0 : P.n_obs_peers); ++p) {     // -1: this rank's buffer, then the peers'
                float* g = (p < 0 ? P.obs : P.obs_peers[p]) + (size_t)(e0 + w) * D;
#pragma unroll 1
                for (int k = lane; k < D; k += 32) g[k] = row[k];
            }
            if (MODE == 0 && P.obs_mc) {
                float* g = P.obs_mc + (size_t)(e0 + w) * D;
#pragma unroll 1
                for (int k = lane; k < D; k += 32)
                    asm volatile("multimem.st.weak.global.f32 [%0], %1;" ::"l"(g + k), "f"(row[k]) : "memory");
            }
        }
    }
    if (gather && (P.arrive_mc != nullptr || P.arrive_peers[0] != nullptr)) {
        // signal: every thread's stores into the peers are ordered before the barrier, the barrier before thread p's
        // system-scope release; the peer's acquire load of its counter then sees the rows (the pattern of a grid sync)
        __syncthreads();
        if (!(L.gather_debug & 2)) signal_peers(P, P.n_obs_peers, tid);
    }
    if (MODE == 0 && !DIRECT && P.dec_wire != nullptr && push) {
        // 16-bit wire format, receiving side: the peers' PREVIOUS kernels delivered a step's rows as int16 thousandths
        // into our wire buffer -- certified here, at the end of our own step, when they have long finished: lane s checks
        // that source rank s has completed as many pushing launches as this rank had before this one -- and this CTA
        // rebuilds its tile's rows of every other rank's block in the fp32 gather buffer (8 values per 16-byte load)
        // while its own pushes drain
        if (warp == 0 && lane < P.arrive_slots && lane != P.arrive_self && !(L.gather_debug & 1)) {
            const unsigned long long own = ld_acquire_sys(P.arrive_local + P.arrive_self);
            if (!wait_arrivals(P.arrive_local + lane, own) && P.gather_timeouts) atomicAdd(P.gather_timeouts, 1u);
        }
        __syncthreads();
        const int n = nE * D, n8 = n >> 3;
#pragma unroll 1
        for (int r = 0; r < P.arrive_slots; ++r) {
            if (r == P.arrive_self) continue;
            const size_t base = ((size_t)r * (size_t)P.n_envs + (size_t)e0) * (size_t)D;
            if ((base & 7u) != 0u || (n & 7) != 0 || ((((uintptr_t)P.dec_wire) | ((uintptr_t)P.dec_obs)) & 15u) != 0u) {
#pragma unroll 1
                for (int i = tid; i < n; i += T) P.dec_obs[base + i] = wire16_decode((int32_t)P.dec_wire[base + i]);   // ragged tile
                continue;
            }
            const uint4* src = reinterpret_cast<const uint4*>(P.dec_wire + base);
            float4* dst = reinterpret_cast<float4*>(P.dec_obs + base);
#pragma unroll 2
            for (int c = tid; c < n8; c += T) {
                const uint4 wv = src[c];
                const uint32_t ww[4] = {wv.x, wv.y, wv.z, wv.w};
                float o[8];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    o[2 * k] = wire16_decode((int32_t)(int16_t)(ww[k] & 0xFFFFu));
                    o[2 * k + 1] = wire16_decode((int32_t)(int16_t)(ww[k] >> 16));
                }
                dst[2 * c] = make_float4(o[0], o[1], o[2], o[3]);
                dst[2 * c + 1] = make_float4(o[4], o[5], o[6], o[7]);
            }
        }
    }
    if (push) {
        if (push_bulk) {
            // thread 0 issued the bulk stores of the old rows at the start of the kernel: they have had the whole step
            // to drain; wait for their completion (not just for the reads of the staging tile), then signal
            if (tid == 0) {
                asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                asm volatile("fence.proxy.async;" ::: "memory");
            }
        } else {
            // ragged or unaligned tile: forward the old rows with plain loads / stores from all threads
#pragma unroll 1
            for (int p = 0; p < P.n_push_peers; ++p) {
                const uint16_t* src = reinterpret_cast<const uint16_t*>(reinterpret_cast<const uint8_t*>(P.push_src) + push_off);
                uint16_t* dst = reinterpret_cast<uint16_t*>(reinterpret_cast<uint8_t*>(P.push_peers[p]) + push_off);
#pragma unroll 1
                for (int k = tid; k < (int)(push_bytes >> 1); k += T) dst[k] = src[k];
            }
        }
        __syncthreads();
        if (!(L.gather_debug & 2)) signal_peers(P, P.n_push_peers, tid);
    }
}

}  // namespace

// ------------------------------------------------------------- host side
static size_t up16(size_t x) { return (x + 15) & ~(size_t)15; }

int cn_flat_make_layout(int n_peds, int n_samples, int obs_dim, int tile, int threads, int stage, int direct, cn_flat_layout* L) {
    const int N = n_peds, D = obs_dim, W = tile;
    if (direct && stage) return -1;
    if (threads != 128 && threads != 192 && threads != 256 && threads != 384 && threads != 512) return -1;
    if (W < 1 || W > 32) return -1;
    if ((size_t)W * N > 0x3FFF) return -1;                                  // group-entry / list-entry fields
    L->W = W;
    L->threads = threads;
    L->plain_store = 0;
    L->magic_n = (N > 1) ? (uint32_t)(((1ull << 32) + (uint64_t)N - 1) / (uint64_t)N) : 0u;
    L->obs_direct = direct ? 1 : 0;
    /* ray-group lists: a primitive whose groups do not fit is walked directly (same results), so the capacity is a
     * matter of speed only; the direct layout, which is after the smallest tile, sizes them for the average world */
    {
        /* Ray-group lists: 14 wall groups and 10 pedestrian groups of 8 rays per world at 359 rays in the direct layout
         * (which is after the smallest tile), 32 and 24 in the staged one -- in proportion for other scans.  Measured at c5
         * (719 rays): with the 359-ray capacities the lists overflow all the time and the primitives are walked on the
         * slow path, 95 us per step; in proportion 67 us.  CN_FLAT_CAPS="gw,gp" overrides the per-world figures (experiments). */
        const uint32_t nr = (uint32_t)(n_samples > 1 ? n_samples - 1 : 1);
        uint32_t gw = ((direct ? 14u : 32u) * nr + 358u) / 359u, gp = ((direct ? 10u : 24u) * nr + 358u) / 359u;
        if (const char* caps = getenv("CN_FLAT_CAPS")) {
            const int a = atoi(caps); const char* comma = strchr(caps, ',');
            if (a > 0) gw = (uint32_t)a;
            if (comma && atoi(comma + 1) > 0) gp = (uint32_t)atoi(comma + 1);
        }
        L->cap_wg = (uint32_t)W * (gw < 8u ? 8u : gw);
        L->cap_pg = (uint32_t)W * (gp < 8u ? 8u : gp);
    }
    L->strip_mask = 31u;
    size_t o = 0;
    o += (size_t)W * CN_ROBOT_WORDS * 4;            L->off_pa = (uint32_t)o;
    /* 32 B per pedestrian: the old-position plane (16 B each) in the first half, the contact masks of the slow list
     * (8 B each) in the second; the candidate records (32 B per pedestrian, phases 2-6) reuse all of it, dead by then */
    L->off_rec = (uint32_t)o;
    L->off_pk = 0;
    L->off_peers = (uint32_t)(o + (size_t)W * N * 16);
    o += (size_t)W * N * 32;                        L->off_pb = (uint32_t)o;
    o += (size_t)W * N * 16;                        L->off_pa2 = (uint32_t)o;
    o += (size_t)W * N * 16;                        L->off_act = (uint32_t)o;
    o = up16(o + (size_t)W * 8);                    L->off_obs = (uint32_t)o;
    o = up16(o + (direct ? 0 : (size_t)W * D * 4)); L->off_sc = (uint32_t)o;
    L->strip_words = (N > 32) ? 2u : 1u;
    o = up16(o + (size_t)W * F_WORDS * 4);          L->off_strips = (uint32_t)o;
    {
        /* contact-prefilter strip masks: [W][32][strip_words] per axis.  The staged layout keeps both axes here; the
         * direct layout puts the y masks into the quarter of the pa / peers / rec region nobody uses while they are
         * alive (32 B per pedestrian: 16 B old position, 8 B contact masks, 8 B free until the candidate records take
         * the region over in phase 2b) when they fit there */
        const size_t axis = (size_t)W * (L->strip_mask + 1) * L->strip_words * 4;
        const size_t free_lo = up16((size_t)L->off_peers + (size_t)W * N * 8), free_hi = (size_t)L->off_pb;
        o += axis;
        if (direct && free_lo + axis <= free_hi) L->off_strips_y = (uint32_t)free_lo;
        else { L->off_strips_y = (uint32_t)o; o += axis; }
    }
    L->off_clist = (uint32_t)o;
    o = up16(o + (size_t)W * N * 2);                L->off_clw = (uint32_t)o;
    o = up16(o + (size_t)W * N);                    L->off_rlist = (uint32_t)o;
    o = up16(o + (size_t)W * N * 2);                L->off_olist = (uint32_t)o;
    o = up16(o + (size_t)W * N * 2);                L->off_mark = (uint32_t)o;
    o = up16(o + (size_t)W * N);                    L->off_wg = (uint32_t)o;
    o += (size_t)L->cap_wg * 4;                     L->off_pg = (uint32_t)o;
    o = up16(o + (size_t)L->cap_pg * 4);            L->off_cnt = (uint32_t)o;
    o += C_WORDS * 4;                               L->off_bar = (uint32_t)o;
    o += 16;
    L->off_fillc = 0;
    L->fillc_bytes = 0;
    if (direct) {
        /* direct > 1: the size of the constant tile in bytes (one bulk store covers that much of a row's ray columns) */
        L->fillc_bytes = (direct > 1) ? (uint32_t)(direct & ~15) : CF_FILLC_MIN;
        o = up16(o); L->off_fillc = (uint32_t)o; o += L->fillc_bytes;
    }
    L->off_stage = 0;
    if (stage) { o = up16(o); L->off_stage = (uint32_t)o; o = up16(o + (size_t)W * D * (stage == 2 ? 2 : 4)); }   /* int16 or fp32 rows */
    L->total = (uint32_t)o;
    return 0;
}

int cn_flat_pick_tile(int n_peds, int n_samples, int obs_dim, int n_envs, int n_sms, size_t smem_per_sm, int stage, int direct, cn_flat_layout* L) {
    // Staged rows: 256-thread CTAs, five per SM.  Among the tiles (even, <= 16 worlds, row block able to leave by bulk
    // store) that fit the CTA's share of the SM's shared memory: a batch that fits one wave gets the smallest tile that
    // still does (most CTAs in flight, shortest critical path); a larger batch the tile that fills its waves best.
    // Direct rows: tiles up to 32 worlds (the lane = world warps), any width; 256 threads x 6 CTAs per SM (40 registers)
    // or 384 threads x 4.  Measured on B200 (profiles/r02b/direct_rows_ab.txt): c3 fits ONE wave with 19 worlds x 256
    // threads (24.0 us against 30.1 us staged, two residency rounds; 28 worlds x 384 threads is the same within 1 %);
    // c5, which needs three waves either way, is 5 % faster with 12 worlds x 384 threads than with 8 x 256 -- the
    // per-CTA fixed cost is shared by more worlds.  So: the smallest one-wave tile at 256 threads if there is one, else
    // the best-scoring of both CTA sizes.
    const int n_opt = direct ? 2 : 1;
    const int opt_threads[2] = {256, 384};
    const int opt_ctas[2] = {direct ? 6 : CN_FLAT_CTAS_PER_SM, 4};
    int best = 0, best_threads = 256; double best_score = -1.0; size_t best_budget = 0;
    for (int k = 0; k < n_opt; ++k) {
        const int threads = opt_threads[k], ctas = opt_ctas[k];
        const size_t budget = smem_per_sm / ctas - 1024;
        const long slots = (long)ctas * (n_sms > 0 ? n_sms : 148);
        for (int W = direct ? 32 : 16; W >= 1; --W) {
            cn_flat_layout t;
            if (cn_flat_make_layout(n_peds, n_samples, obs_dim, W, threads, stage, direct, &t) != 0 || t.total > budget) continue;
            const bool bulk = direct || (((size_t)W * obs_dim) % 4 == 0 && W % 2 == 0 && (stage < 2 || ((size_t)W * obs_dim) % 8 == 0));
            const long n_cta = ((long)n_envs + W - 1) / W;
            const long waves = (n_cta + slots - 1) / slots;
            double score;
            if (waves == 1) score = 100.0 - W - (k ? 50.0 : 0.0);   /* one wave: the smaller the tile the better, 256 threads first ... */
            else score = (double)n_cta / (double)(waves * slots) * ((double)W / (double)(W + 4));   /* ... else wave fill x the
                                                                       share of a CTA's life that is not per-CTA fixed cost */
            if (W < 4) score -= 50.0;                               /* tiny tiles waste the lane = world warps */
            if (!bulk) score -= 10.0;
            if (direct && (W & 1)) score -= (waves == 1) ? 1.5 : 0.1;   /* odd tiles: actions by plain loads (measured c2: W = 5 11.95 us, 6 11.45) */
            if (score > best_score) { best_score = score; best = W; best_threads = threads; best_budget = budget; }
        }
    }
    if (!best) return -1;
    if (direct) {
        /* whatever the tile leaves of the CTA's share of shared memory goes to the constant tile of the bulk fill,
         * up to one whole row of ray columns (one bulk store per row instead of three) */
        cn_flat_layout t;
        if (cn_flat_make_layout(n_peds, n_samples, obs_dim, best, best_threads, stage, 1, &t) != 0) return -1;
        const size_t want = up16((size_t)(n_samples - 1) * 4);
        size_t fc = CF_FILLC_MIN + (best_budget > t.total ? best_budget - t.total : 0);
        if (fc > want) fc = want;
        direct = (int)(fc & ~(size_t)15);
        if (direct < (int)CF_FILLC_MIN) direct = (int)CF_FILLC_MIN;
    }
    return cn_flat_make_layout(n_peds, n_samples, obs_dim, best, best_threads, stage, direct, L);
}

template <int MODE, int T, int DIRECT>
static cudaError_t launch_flat_t(const cn_kparams& P, const cn_flat_layout& L, cudaStream_t stream) {
    auto k = cn_flat_kernel<MODE, T, DIRECT>;
    {
        constexpr int ti = (T == 128) ? 0 : (T == 192) ? 1 : (T == 256) ? 2 : (T == 384) ? 3 : 4;
        cudaError_t e = cn_ensure_smem_attr(reinterpret_cast<const void*>(k), DIRECT ? 16 + ti : MODE * 5 + ti, L.total);
        if (e != cudaSuccess) return e;
    }
    const int grid = (P.n_envs + L.W - 1) / L.W;
    if (!L.pdl) {
        k<<<grid, T, L.total, stream>>>(P, L);
        return cudaGetLastError();
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)T); cfg.dynamicSmemBytes = L.total; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, P, L);
}

#ifdef CN_TIMELINE
extern "C" int cn_debug_set_timeline_flat(unsigned long long* dev_ptr) {
    return (int)cudaMemcpyToSymbol(g_timeline, &dev_ptr, sizeof(dev_ptr));
}
#endif

// thread s watches source rank s: hold the stream until every other rank's slot of this rank's arrival counters has
// caught up with this rank's own step count (every CTA of that rank's latest step has stored its rows here and signalled); bounded, a lost peer is counted instead of hanging the GPU
__global__ void cn_gather_wait_kernel(const unsigned long long* counters, int n_slots, int self_slot, unsigned int* timeouts) {
    const int s = (int)threadIdx.x;
    // this rank's own slot = CTAs of all its steps so far (they have completed: stream order); equal shards, so the
    // same number of arrivals is due from every other rank
    const unsigned long long target = ld_acquire_sys(counters + self_slot);
    if (s < n_slots && s != self_slot && !wait_arrivals(counters + s, target) && timeouts) atomicAdd(timeouts, 1u);
}
cudaError_t cn_launch_gather_wait(const unsigned long long* counters, int n_slots, int self_slot,
                                  unsigned int* timeouts, cudaStream_t stream) {
    cn_gather_wait_kernel<<<1, 32, 0, stream>>>(counters, n_slots, self_slot, timeouts);
    return cudaGetLastError();
}

// Push-only launch of the pipelined gather: the rows of the LATEST step, which no later step kernel will forward.
// Same tiling, same guard and signal as the step kernel, so the arrival counting does not care which of the two
// delivered a step's rows.
__global__ void __launch_bounds__(256) cn_push_kernel(const __grid_constant__ cn_kparams P, int W) {
    const int D = P.d.obs_dim, tid = (int)threadIdx.x;
    const int e0 = (int)blockIdx.x * W, nE = min(W, P.n_envs - e0);
    if (tid < 32) guard_peer_buffers(P, tid);
    __syncthreads();
    const size_t elem = P.push_wire16 ? 2u : 4u;
    const size_t off = (size_t)e0 * D * elem, bytes = (size_t)nE * D * elem;
    const bool vec = P.push_bulk_ok && (off % 16 == 0) && (bytes % 16 == 0);
    int p = (P.n_push_peers > 1) ? (int)(blockIdx.x % (unsigned)P.n_push_peers) : 0;
#pragma unroll 1
    for (int k = 0; k < P.n_push_peers; ++k) {
        const uint8_t* src = reinterpret_cast<const uint8_t*>(P.push_src) + off;
        uint8_t* dst = reinterpret_cast<uint8_t*>(P.push_peers[p]) + off;
        if (vec) {
            const int n16 = (int)(bytes >> 4);
#pragma unroll 4
            for (int i = tid; i < n16; i += 256) reinterpret_cast<uint4*>(dst)[i] = reinterpret_cast<const uint4*>(src)[i];
        } else {
#pragma unroll 1
            for (int i = tid; i < (int)(bytes >> 1); i += 256) reinterpret_cast<uint16_t*>(dst)[i] = reinterpret_cast<const uint16_t*>(src)[i];
        }
        p = (p + 1 == P.n_push_peers) ? 0 : p + 1;
    }
    __syncthreads();
    signal_peers(P, P.n_push_peers, tid);
}
// The receiving side of the 16-bit wire format: rebuild the fp32 rows of every OTHER rank ([0, rows_total) without this
// rank's own [row_lo, row_hi), which its step kernel wrote in fp32) from the int16 thousandths the peers delivered.
// Pure streaming: 2 bytes in, 4 bytes out per value, 8 values per thread and trip where the chunk lies outside the gap.
__global__ void __launch_bounds__(256) cn_wire_decode_kernel(const int16_t* __restrict__ wire, float* __restrict__ obs_all,
                                                             long long skip_lo, long long skip_hi, long long total) {
    const long long n8 = total >> 3;
    const bool aligned = ((((uintptr_t)wire) & 15u) == 0) && ((((uintptr_t)obs_all) & 15u) == 0);
    for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < n8; c += (long long)gridDim.x * blockDim.x) {
        const long long i0 = c << 3;
        if (i0 >= skip_lo && i0 + 8 <= skip_hi) continue;                      // wholly inside this rank's own rows
        if (aligned && (i0 + 8 <= skip_lo || i0 >= skip_hi)) {
            const uint4 w = *reinterpret_cast<const uint4*>(wire + i0);
            const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
            float o[8];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                o[2 * k] = wire16_decode((int32_t)(int16_t)(ww[k] & 0xFFFFu));
                o[2 * k + 1] = wire16_decode((int32_t)(int16_t)(ww[k] >> 16));
            }
            reinterpret_cast<float4*>(obs_all + i0)[0] = make_float4(o[0], o[1], o[2], o[3]);
            reinterpret_cast<float4*>(obs_all + i0)[1] = make_float4(o[4], o[5], o[6], o[7]);
        } else {
            for (long long i = i0; i < i0 + 8; ++i)
                if (i < skip_lo || i >= skip_hi) obs_all[i] = wire16_decode((int32_t)wire[i]);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x < (total & 7)) {
        const long long i = (n8 << 3) + threadIdx.x;
        if (i < skip_lo || i >= skip_hi) obs_all[i] = wire16_decode((int32_t)wire[i]);
    }
}
cudaError_t cn_launch_wire_decode(const int16_t* wire, float* obs_all, long long row_lo, long long row_hi,
                                  long long rows_total, int obs_dim, cudaStream_t stream) {
    const long long total = rows_total * obs_dim;
    long long blocks = ((total >> 3) + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    cn_wire_decode_kernel<<<(unsigned)blocks, 256, 0, stream>>>(wire, obs_all, row_lo * obs_dim, row_hi * obs_dim, total);
    return cudaGetLastError();
}
cudaError_t cn_launch_push_kernel(const cn_kparams& P, const cn_flat_layout& L, cudaStream_t stream) {
    const int grid = (P.n_envs + L.W - 1) / L.W;
    cn_push_kernel<<<grid, 256, 0, stream>>>(P, L.W);
    return cudaGetLastError();
}

cudaError_t cn_launch_flat_kernel(const cn_kparams& P, const cn_flat_layout& L, int mode, cudaStream_t stream) {
    if (L.obs_direct) {
        if (mode != 0) return cudaErrorInvalidValue;                        // the direct-rows layout serves step launches only
        if (L.threads == 128) return launch_flat_t<0, 128, 1>(P, L, stream);
        if (L.threads == 192) return launch_flat_t<0, 192, 1>(P, L, stream);
        if (L.threads == 384) return launch_flat_t<0, 384, 1>(P, L, stream);
        if (L.threads == 256) return launch_flat_t<0, 256, 1>(P, L, stream);
        return launch_flat_t<0, 512, 1>(P, L, stream);
    }
    if (L.threads == 128) return mode == 0 ? launch_flat_t<0, 128, 0>(P, L, stream) : launch_flat_t<1, 128, 0>(P, L, stream);
    if (L.threads == 192) return mode == 0 ? launch_flat_t<0, 192, 0>(P, L, stream) : launch_flat_t<1, 192, 0>(P, L, stream);
    if (L.threads == 384) return mode == 0 ? launch_flat_t<0, 384, 0>(P, L, stream) : launch_flat_t<1, 384, 0>(P, L, stream);
    if (L.threads == 256) return mode == 0 ? launch_flat_t<0, 256, 0>(P, L, stream) : launch_flat_t<1, 256, 0>(P, L, stream);
    return mode == 0 ? launch_flat_t<0, 512, 0>(P, L, stream) : launch_flat_t<1, 512, 0>(P, L, stream);
}
