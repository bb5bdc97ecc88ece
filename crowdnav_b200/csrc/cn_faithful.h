/*
 * cn_faithful.h -- the `risk_faithful` perception block (CN_FLAG_RISK_FAITHFUL) for ONE world, as the
 * device code of cn_faithful.cu: the reference's own LiDAR segmentation, wall / obstacle typing, uuid-dict
 * tracker, collision cone and top-K block (environment_stage_1_nobonus.py:270-1005) in float64.
 *
 * Flat-array design: a world's rays live in a scratch area (shared memory on the device); every per-ray map
 * (hit points UTL:110-126, gradients ENV:329-347, gradient changes ENV:349-368, neighbour association
 * UTL:435-448) is strided over `nl` lanes, the order-dependent scans (typing state machine ENV:370-404,
 * segment walk ENV:443-620, tracker ENV:656-743, ranking ENV:862-907) run on lane 0, the tracked x confirmed
 * IoU search runs one lane per tracked object and the 64-gon ring of the collision cone (UTL:251-293) one
 * lane per pair of ring edges.  CNF_SYNC() separates the stages.
 *
 * The same source compiles for the host (g++, one "lane": nl = 1, CNF_SYNC a no-op): tests/faithful_host.cpp
 * runs it on the CPU against the independent CPU restatement kept under oracle/ so that the arithmetic is
 * checked bit for bit before it ever reaches a GPU.  That harness is a test tool: the product only launches
 * the kernel.
 */
#ifndef CN_FAITHFUL_H
#define CN_FAITHFUL_H

#include "cn_math64.h"
#include "cn_faithful_state.h"

#define CNF_LANES 64          /* device: threads (two warps) that share one world */
#if defined(__CUDACC__)
#define CNF_FN __device__ __forceinline__
/* the float64 helpers are real calls on the device: inlined everywhere the kernel was 290 KB of SASS and its warps
 * stalled on instruction fetch more than on anything else (ncu: no_instruction 3.0 per issue) */
#define CNF_FN_BIG static __device__ __noinline__
/* the world's CNF_LANES threads meet at their own named barrier (bar = 1 + world index inside the CTA) */
#define CNF_SYNC() asm volatile("bar.sync %0, %1;" :: "r"(bar), "n"(CNF_LANES) : "memory")
#else
#define CNF_FN static inline
#define CNF_FN_BIG static inline
#ifndef CNF_SYNC            /* a host harness may supply a thread barrier to run the lanes as threads */
#define CNF_SYNC() ((void)0)
#endif
#endif

/* debug builds only (-DCN_TIMELINE): %globaltimer stamps per world, 16 slots (profiles/tools/timeline_faithful.py) */
#if defined(__CUDACC__) && defined(CN_TIMELINE)
__device__ unsigned long long* cnf_timeline = nullptr;
#define CNF_STAMP(k) do { if (lane == 0 && cnf_timeline) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); \
    cnf_timeline[((size_t)blockIdx.x * (blockDim.x / CNF_LANES) + threadIdx.x / CNF_LANES) * 16 + (k)] = t_; } } while (0)
#else
#define CNF_STAMP(k) do { } while (0)
#endif

enum { CNF_T_NONE = 0, CNF_T_W = 1, CNF_T_O = 2 };

/* scratch of one world (12.5 KB at 359 rays); later stages reuse arrays the earlier ones have consumed */
typedef struct cnf_scratch {
    int32_t*  gk;       /* [n + 1] round(gradient, 3) in thousandths; then cum_o, cum_w: int16 [n + 1] each */
    int32_t*  hx;       /* [n] hit point, thousandths          */
    int32_t*  hy;
    int32_t*  rmm;      /* [n] round(range, 3), thousandths    */
    int16_t*  src;      /* [n] ray whose record ray i carries  */
    int16_t*  sub;      /* [n + 2] candidate rays of the typing walk; then sub-segment offsets */
    uint8_t*  gok;      /* [nb] gradient defined; then: the ray's record carries a hit */
    uint8_t*  close;    /* [nb] a segment closes after ray i   */
    uint8_t*  subend;   /* [nb] a sub-segment ends at position k; with type[]: the verdicts, int16 [n] */
    uint8_t*  type;     /* [nb] record type                    */
    uint32_t* trk;      /* [CNF_WORLD_WORDS] tracker record    */
    int32_t*  conf;     /* [CNF_CONF_CAP][4] type, x, y, range */
    double*   am_val;   /* [CNF_TRK_CAP] best IoU per tracked; then distance to collision; then CP */
    int32_t*  am_idx;   /* [CNF_TRK_CAP]                       */
    double*   hit;      /* [64][2] ring hits of one probe line */
    int32_t*  red;      /* [256] per-lane partial sums: the same 1 KB as hit[], used before the collision cone */
    uint8_t*  hitf;     /* [64]                                */
    int32_t*  misc;     /* [8] scalars shared between lanes    */
} cnf_scratch;

/* bytes of scratch for n rays (carved in this order: 8-byte items first) */
CN_HD size_t cnf_scratch_bytes(int n) {
    const size_t nb = ((size_t)n + 1) & ~(size_t)1;
    size_t b = 0;
    b += sizeof(double) * CNF_TRK_CAP + sizeof(double) * 128; /* am_val, hit / red */
    b += sizeof(uint32_t) * CNF_WORLD_WORDS;                  /* trk: 1584 B, keeps 8-byte alignment */
    b += 64;                                                  /* hitf */
    b += sizeof(int32_t) * ((size_t)n + 1);                   /* gk */
    b += 3 * sizeof(int32_t) * (size_t)n;                     /* hx, hy, rmm */
    b += sizeof(int32_t) * (CNF_CONF_CAP * 4 + CNF_TRK_CAP + 8);
    b += sizeof(int16_t) * (2 * (size_t)n + 2);               /* src, sub */
    b += 4 * nb;
    return (b + 15) & ~(size_t)15;
}
CN_HD void cnf_scratch_carve(unsigned char* base, int n, cnf_scratch* S) {
    const size_t nb = ((size_t)n + 1) & ~(size_t)1;
    unsigned char* p = base;
    S->am_val = (double*)p; p += sizeof(double) * CNF_TRK_CAP;
    S->hit = (double*)p; S->red = (int32_t*)p; p += sizeof(double) * 128;
    S->trk = (uint32_t*)p; p += sizeof(uint32_t) * CNF_WORLD_WORDS;
    S->hitf = p; p += 64;                                     /* 8-byte aligned: read as 8 words */
    S->gk = (int32_t*)p; p += sizeof(int32_t) * ((size_t)n + 1);
    S->hx = (int32_t*)p; p += sizeof(int32_t) * (size_t)n;
    S->hy = (int32_t*)p; p += sizeof(int32_t) * (size_t)n;
    S->rmm = (int32_t*)p; p += sizeof(int32_t) * (size_t)n;
    S->conf = (int32_t*)p; p += sizeof(int32_t) * CNF_CONF_CAP * 4;
    S->am_idx = (int32_t*)p; p += sizeof(int32_t) * CNF_TRK_CAP;
    S->misc = (int32_t*)p; p += sizeof(int32_t) * 8;
    S->src = (int16_t*)p; p += sizeof(int16_t) * (size_t)n;
    S->sub = (int16_t*)p; p += sizeof(int16_t) * ((size_t)n + 2);
    S->gok = p; p += nb; S->close = p; p += nb; S->subend = p; p += nb; S->type = p;   /* subend + type: 2-byte aligned */
}

/* k / 1000 for the thousandths this block stores (positions within +-31 m, ranges <= max_range): |k| < 2^25 */
#define CNF_MILLI(k) cn_milli64_small((int32_t)(k))

CNF_FN double cnf_ld64(const uint32_t* w) {
    unsigned long long u = (unsigned long long)w[0] | ((unsigned long long)w[1] << 32);
    double d; memcpy(&d, &u, 8); return d;
}
CNF_FN void cnf_st64(uint32_t* w, double d) {
    unsigned long long u; memcpy(&u, &d, 8);
    w[0] = (uint32_t)u; w[1] = (uint32_t)(u >> 32);
}

/* UTL:421-448: round(IoU, 3) of two squares of half-size h centred on points given in thousandths */
CNF_FN_BIG long long cnf_round3k(double x) { return cn_py_round3_k64(x); }
CNF_FN_BIG double cnf_div(double a, double b) { return a / b; }            /* one copy of the ~25-instruction IEEE division */
CNF_FN_BIG double cnf_hypot(double a, double b) { return cn_hypot64(a, b); }
CNF_FN_BIG double cnf_np_round3(double x) { return cn_np_round3_64(x); }
#if defined(__CUDACC__)
#define CNF_ROLLED _Pragma("unroll 1")       /* keep the loops rolled: the kernel is bound by instruction fetch */
#else
#define CNF_ROLLED
#endif

CNF_FN_BIG double cnf_iou(int32_t axk, int32_t ayk, int32_t bxk, int32_t byk, double h) {
    const double ax = CNF_MILLI(axk), ay = CNF_MILLI(ayk), bx = CNF_MILLI(bxk), by = CNF_MILLI(byk);
    const double ax0 = ax - h, ax1 = ax + h, ay0 = ay - h, ay1 = ay + h;
    const double bx0 = bx - h, bx1 = bx + h, by0 = by - h, by1 = by + h;
    const double w = (ax1 < bx1 ? ax1 : bx1) - (ax0 > bx0 ? ax0 : bx0);
    const double hh = (ay1 < by1 ? ay1 : by1) - (ay0 > by0 ? ay0 : by0);
    const double inter = (w > 0.0 && hh > 0.0) ? w * hh : 0.0;
    if (inter == 0.0) return 0.0;                               /* 0 / union = +0.0, round(+0.0, 3) = 0.0: the same bits, no division */
    const double area_a = (ax1 - ax0) * (ay1 - ay0), area_b = (bx1 - bx0) * (by1 - by0);
    const double uni = area_a + area_b - inter;
    return cn_milli64(cn_py_round3_k64(cnf_div(inter, uni)));
}
/* round(IoU, 3) > 0 -- all the association stage asks -- i.e. IoU >= 0.0005 after the division's and the rounding's own
 * round-off.  Far from that threshold the answer needs neither: inter > union / 1000 is a certain yes (ratio > 0.000999),
 * inter < union / 4000 a certain no; only what lies between goes through the division and the exact rounding. */
CNF_FN_BIG int cnf_iou_pos(int32_t axk, int32_t ayk, int32_t bxk, int32_t byk, double h) {
    const double ax = CNF_MILLI(axk), ay = CNF_MILLI(ayk), bx = CNF_MILLI(bxk), by = CNF_MILLI(byk);
    const double ax0 = ax - h, ax1 = ax + h, ay0 = ay - h, ay1 = ay + h;
    const double bx0 = bx - h, bx1 = bx + h, by0 = by - h, by1 = by + h;
    const double w = (ax1 < bx1 ? ax1 : bx1) - (ax0 > bx0 ? ax0 : bx0);
    const double hh = (ay1 < by1 ? ay1 : by1) - (ay0 > by0 ? ay0 : by0);
    if (!(w > 0.0 && hh > 0.0)) return 0;
    const double inter = w * hh;
    const double area_a = (ax1 - ax0) * (ay1 - ay0), area_b = (bx1 - bx0) * (by1 - by0);
    const double uni = area_a + area_b - inter;
    if (inter * 1000.0 > uni) return 1;
    if (inter * 4000.0 < uni) return 0;
    return cn_py_round3_k64(cnf_div(inter, uni)) > 0;
}

/* UTL:110-126 for observation ray i: both coordinates in thousandths */
typedef struct { int32_t x, y; } cnf_pt;
CNF_FN_BIG cnf_pt cnf_hit_point(double inc_deg, double x, double y, double yaw, int i, double r) {
    const double ang = ((double)i * inc_deg) * CN64_DEG2RAD - yaw;
    double s, c; cn_sincos64(ang, &s, &c);
    cnf_pt o;
    o.x = (int32_t)cn_py_round3_k64(x + r * c);
    o.y = (int32_t)cn_py_round3_k64(y + (r * s) * -1.0);
    return o;
}

/* ---- collectives over the world's lanes.  Every lane of the world calls them from converged code.  On the device
 * (CNF_LANES = 64 = two warps) they are shuffle scans / reductions plus ONE exchange between the two warps through
 * S.red[slot ...] and one CNF_SYNC(); on the host (nl lanes as threads, or nl = 1) they are the plain loops over
 * per-lane partials in S.red the kernel itself used before (9.7 + 6.4 + 10.4 % of its executed instructions were those
 * loops: profiles/r02/ncu_c2_faithful_v12.txt).  Call sites use distinct slots, so a lane that is still reading the
 * previous exchange cannot meet the next one's writes. ---- */
#if defined(__CUDACC__)
#define CNF_COLL_ARGS const cnf_scratch& S, int lane, int nl, int bar
#define CNF_COLL_PASS S, lane, nl, bar
CNF_FN int cnf_warp_incl(int v, int l32) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int y = __shfl_up_sync(0xFFFFFFFFu, v, d); if (l32 >= d) v += y; }
    return v;
}
/* exclusive prefix sum of v in lane order; *total = the sum over all lanes */
CNF_FN int cnf_scan_excl(CNF_COLL_ARGS, int slot, int v, int* total) {
    (void)nl;
    const int l32 = lane & 31;
    const int x = cnf_warp_incl(v, l32);
    if (l32 == 31) S.red[slot + (lane >> 5)] = x;
    CNF_SYNC();
    const int t0 = S.red[slot], t1 = S.red[slot + 1];
    *total = t0 + t1;
    return x - v + (lane >= 32 ? t0 : 0);
}
/* three exclusive prefix sums at once (a | b << 10 | c << 20 would overflow: counts go up to n each), one exchange */
CNF_FN void cnf_scan_excl3(CNF_COLL_ARGS, int slot, int* a, int* b, int* c, int* total_a) {
    (void)nl;
    const int l32 = lane & 31;
    const int xa = cnf_warp_incl(*a, l32), xb = cnf_warp_incl(*b, l32), xc = cnf_warp_incl(*c, l32);
    if (l32 == 31) { int32_t* r = S.red + slot + 3 * (lane >> 5); r[0] = xa; r[1] = xb; r[2] = xc; }
    CNF_SYNC();
    const int hi = lane >= 32;
    *total_a = S.red[slot] + S.red[slot + 3];
    *a = xa - *a + (hi ? S.red[slot] : 0);
    *b = xb - *b + (hi ? S.red[slot + 1] : 0);
    *c = xc - *c + (hi ? S.red[slot + 2] : 0);
}
/* min of mn, max of mx, sum of sm over all lanes, to every lane */
CNF_FN void cnf_min_max_sum(CNF_COLL_ARGS, int slot, int* mn, int* mx, int* sm) {
    (void)nl;
    int a = *mn, b = *mx, c = *sm;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const int a2 = __shfl_xor_sync(0xFFFFFFFFu, a, d), b2 = __shfl_xor_sync(0xFFFFFFFFu, b, d), c2 = __shfl_xor_sync(0xFFFFFFFFu, c, d);
        a = a2 < a ? a2 : a; b = b2 > b ? b2 : b; c += c2;
    }
    if ((lane & 31) == 0) { int32_t* r = S.red + slot + 3 * (lane >> 5); r[0] = a; r[1] = b; r[2] = c; }
    CNF_SYNC();
    const int32_t* r = S.red + slot;
    *mn = r[0] < r[3] ? r[0] : r[3]; *mx = r[1] > r[4] ? r[1] : r[4]; *sm = r[2] + r[5];
}
#else
#define CNF_COLL_ARGS const cnf_scratch& S, int lane, int nl, int bar
#define CNF_COLL_PASS S, lane, nl, bar
CNF_FN int cnf_scan_excl(CNF_COLL_ARGS, int slot, int v, int* total) {
    (void)bar;
    CNF_SYNC();                 /* the partials of the previous collective (same words) have been read by everybody */
    S.red[slot + lane] = v;
    CNF_SYNC();
    int base = 0, t = 0;
    for (int l = 0; l < nl; ++l) { if (l < lane) base += S.red[slot + l]; t += S.red[slot + l]; }
    *total = t;
    return base;
}
CNF_FN void cnf_scan_excl3(CNF_COLL_ARGS, int slot, int* a, int* b, int* c, int* total_a) {
    (void)bar;
    CNF_SYNC();                 /* the partials of the previous collective (same words) have been read by everybody */
    int32_t* r = S.red + slot;
    r[3 * lane] = *a; r[3 * lane + 1] = *b; r[3 * lane + 2] = *c;
    CNF_SYNC();
    int ba = 0, bb = 0, bc = 0, t = 0;
    for (int l = 0; l < nl; ++l) {
        if (l < lane) { ba += r[3 * l]; bb += r[3 * l + 1]; bc += r[3 * l + 2]; }
        t += r[3 * l];
    }
    *a = ba; *b = bb; *c = bc; *total_a = t;
}
CNF_FN void cnf_min_max_sum(CNF_COLL_ARGS, int slot, int* mn, int* mx, int* sm) {
    (void)bar;
    CNF_SYNC();                 /* the partials of the previous collective (same words) have been read by everybody */
    int32_t* r = S.red + slot;
    r[3 * lane] = *mn; r[3 * lane + 1] = *mx; r[3 * lane + 2] = *sm;
    CNF_SYNC();
    int a = r[0], b = r[1], c = 0;
    for (int l = 0; l < nl; ++l) { if (r[3 * l] < a) a = r[3 * l]; if (r[3 * l + 1] > b) b = r[3 * l + 1]; c += r[3 * l + 2]; }
    *mn = a; *mx = b; *sm = c;
}
#endif

/*
 * One get_state of the perception block for one world.
 *   scan32 : cleaned ranges in observation order (UTL:375-392), fp32, `no_return32` where nothing was hit
 *   step_counter : 0 inside reset (ENV:1245), else the 1-based step
 *   kblock : 4K floats of the observation row (written by lane 0)
 *   lane, nl : this thread's index among the nl threads that share the world; bar: their barrier (device only)
 *   S.trk must hold the world's tracker record on entry (all lanes see it) and holds the new one on exit.
 */
CNF_FN void cnf_world(const cnf_params* P, const cnf_scratch S, double x, double y, double yaw,
                      const float* scan32, float no_return32, int step_counter, float* kblock,
                      int lane, int nl, int bar) {
    (void)bar;
    CNF_STAMP(0);
    const int n = P->n_rays, K = P->k_obstacles;
    const int32_t max_mm = (int32_t)cnf_round3k(P->max_range);

    /* ---- reset: fresh tracker; bounding_box_size from the ground-truth ring (ENV:286-290, UTL:405-419) ---- */
    if (step_counter == 0) {
        CNF_ROLLED for (int k = lane; k < CNF_WORLD_WORDS; k += nl) S.trk[k] = 0u;
        CNF_ROLLED for (int i = lane; i < n; i += nl) { const cnf_pt h = cnf_hit_point(P->inc_deg, x, y, yaw, i, P->max_range); S.hx[i] = h.x; S.hy[i] = h.y; }
        CNF_SYNC();
        if (lane == 0) {                                       /* summed in ray order, like sum() does */
            double sum = 0.0;
            CNF_ROLLED for (int i = 0; i < n; ++i) {
                const int j = (i == n - 1) ? 0 : i + 1;
                sum += cnf_hypot(CNF_MILLI(S.hx[i]) - CNF_MILLI(S.hx[j]), CNF_MILLI(S.hy[i]) - CNF_MILLI(S.hy[j]));
            }
            cnf_st64(S.trk + CNF_H_BBOX, cnf_div(sum, (double)n));
        }
        CNF_SYNC();
    }
    const double bbox = cnf_ld64(S.trk + CNF_H_BBOX);

    /* ---- per-ray maps ----
     * Nine rays in ten return nothing, and nothing downstream ever looks at such a ray's hit point unless a ray that
     * did return something is its neighbour: gradients (ENV:329-347) are taken at returning rays and reach one ray
     * ahead, segments without a return are dropped before their poses are read (ENV:573), the first / last merge reads
     * rays 0 and n - 1 -- and the association of two CONSECUTIVE no-return rays (UTL:435-448) is known without their
     * points: both lie on the max-range circle one ray spacing apart, the squares around them have half-size
     * bounding_box_size = that spacing (ENV:286-290), so after rounding to millimetres they still overlap by a third
     * of their area (IoU 0.13 ... 0.33 against the 0.0005 that decides) as long as the spacing is well above the
     * millimetre -- `lean` below.  So the float64 sin / cos / roundings of a hit point run only for the rays that
     * returned something, their two neighbours and rays 0 and n - 1. */
    const int lean = bbox >= 0.003;
    CNF_ROLLED for (int i = lane; i < n; i += nl) {
        const float r32 = scan32[i];
        const int ret = !(r32 >= no_return32);
        const double r = ret ? (double)r32 : P->max_range;
        int need = !lean || ret || i == 0 || i == n - 1;
        if (!need) need = !(scan32[i - 1] >= no_return32) || !(scan32[i + 1] >= no_return32);
        cnf_pt h; h.x = 0; h.y = 0;
        if (need) h = cnf_hit_point(P->inc_deg, x, y, yaw, i, r);
        S.hx[i] = h.x; S.hy[i] = h.y;
        S.rmm[i] = ret ? (int32_t)cnf_round3k(r) : max_mm;
    }
    CNF_SYNC();
    CNF_STAMP(1);
    CNF_ROLLED for (int i = lane; i < n; i += nl) {                       /* ENV:329-347 */
        if (S.rmm[i] == max_mm) { S.gok[i] = 0; S.gk[i] = 0; continue; }
        const int j = (i == n - 1) ? 0 : i + 1;
        const double dy = CNF_MILLI(S.hy[i]) - CNF_MILLI(S.hy[j]);
        double g = 0.0;
        if (dy != 0.0) g = cnf_div(CNF_MILLI(S.hx[i]) - CNF_MILLI(S.hx[j]), dy);
        S.gk[i] = (int32_t)cnf_round3k(g); S.gok[i] = 1;   /* |g| <= 1.2 m / 1 mm: fits easily */
    }
    CNF_SYNC();
    CNF_STAMP(2);
    /* the change of gradient of ray i < n-1 (ENV:349-368) is defined when rays i and i+1 both have a gradient; it is
     * recomputed where needed from the two rounded gradients (the same doubles round() returned) */
#define CNF_G(i) cn_milli64_small(S.gk[i])          /* |gradient| <= 1.2 m / 1 mm = 1200 */
#define CNF_CHG_OK(i) (S.gok[i] && S.gok[(i) + 1])
#define CNF_CHG(i) fabs(CNF_G(i) - CNF_G((i) + 1))

    /* ---- typing (ENV:370-404): only rays with a defined gradient change take part in the state machine, so their
     * indices are compacted first (lane-chunked, order preserving) and lane 0 walks the short list ---- */
    const int chunk = (n + nl - 1) / nl;
    const int c_lo = (lane * chunk < n) ? lane * chunk : n;
    const int c_hi = (c_lo + chunk < n) ? c_lo + chunk : n;
    int16_t* cand = S.sub;                                      /* free until the sub-segment offsets are built */
    int my_cand = 0;
    {
        int cnt = 0;
        CNF_ROLLED for (int i = c_lo; i < c_hi; ++i) {
            S.type[i] = (uint8_t)CNF_T_NONE; S.src[i] = (int16_t)i;
            cnt += (i != n - 1 && CNF_CHG_OK(i));
        }
        my_cand = cnt;
    }
    CNF_STAMP(3);
    {
        int total = 0;
        int base = cnf_scan_excl(CNF_COLL_PASS, 0, my_cand, &total);
        CNF_ROLLED for (int i = c_lo; i < c_hi; ++i) if (i != n - 1 && CNF_CHG_OK(i)) cand[base++] = (int16_t)i;
        /* the last ray inherits `last_grad`: the change of the latest earlier ray that has a gradient (lane 0 finds it) */
        if (lane == 0) {
            int last_ok = 0, last_j = 0;
            if (S.gok[n - 1]) {
                CNF_ROLLED for (int i = n - 2; i >= 0; --i)
                    if (S.gok[i]) { last_ok = CNF_CHG_OK(i); last_j = i; break; }
            }
            S.misc[3] = last_ok; S.misc[4] = last_j;
        }
        CNF_SYNC();
        /* What the state machine asks of a candidate are four yes / no questions about float64 values -- is its change of
         * gradient zero, has the next ray a change, is that one zero, are the two equal -- and each depends on the
         * candidate alone: the lanes answer them side by side (the same doubles, the same comparisons), and lane 0
         * then walks the candidates over one byte each instead of evaluating ~50 dependent float64 instructions per step
         * of a serial walk (9 us of a 44 us world lifetime at c2: profiles/r02b/timeline_faithful_v20.txt). */
        uint8_t* cflag = S.close;                               /* free until the association stage writes it */
        {
            const int last_ok = S.misc[3], last_j = S.misc[4];
            CNF_ROLLED for (int q = lane; q < total; q += nl) {
                const int i = cand[q];
                const double ci = CNF_CHG(i);
                const int nok = (i + 1 == n - 1) ? last_ok : CNF_CHG_OK(i + 1);
                double cn = 0.0;
                if (nok) cn = (i + 1 == n - 1) ? CNF_CHG(last_j) : CNF_CHG(i + 1);
                cflag[q] = (uint8_t)((ci == 0.0 ? 1 : 0) | (nok ? 2 : 0) | ((nok && cn == 0.0) ? 4 : 0) |
                                     ((nok && fabs(ci - cn) == 0.0) ? 8 : 0));
            }
        }
        CNF_SYNC();
        if (lane == 0) {
            /* a record is (type, source ray): `_scans_object_type[i] = last_type` hands ray i an EARLIER ray's range and pose */
            int last_type = CNF_T_NONE, last_src = 0, du = 0;
            CNF_ROLLED for (int q = 0; q < total; ++q) {
                const int i = cand[q];
                const int f = cflag[q];
                const int ci_zero = f & 1, nok = f & 2, cn_zero = f & 4, same = f & 8;
                int t, s = i;
                if (ci_zero) { t = CNF_T_W; last_type = CNF_T_W; last_src = i; }
                else if (du != 1) {
                    t = CNF_T_O;
                    if (cn_zero) { t = CNF_T_W; last_type = CNF_T_W; last_src = i; du = 0; }
                    if (nok) {
                        if (same) { t = CNF_T_W; s = i; last_type = CNF_T_W; last_src = i; du = 0; }
                        else { t = last_type; s = (last_type == CNF_T_NONE) ? i : last_src; du += 1; }
                    }
                } else {
                    t = CNF_T_O; last_type = CNF_T_O; last_src = i;
                    if (cn_zero) du = 0;
                }
                S.type[i] = (uint8_t)t; S.src[i] = (int16_t)s;
            }
        }
    }
#undef CNF_G
#undef CNF_CHG_OK
#undef CNF_CHG
    CNF_SYNC();
    CNF_STAMP(4);
    /* neighbour association on the records' poses (ENV:443-486); does the record carry a hit? */
    uint8_t* hflag = S.gok;                                     /* gradients are consumed */
    int seg_first = n, seg_last = -1, seg_cnt = 0;
    {
        int first = n, last = -1, cnt = 0;
        CNF_ROLLED for (int i = lane; i < n; i += nl) {
            int cl = 1;
            if (i != n - 1) {
                const int a = S.src[i], b = S.src[i + 1];
                if (lean && S.rmm[a] == max_mm && S.rmm[b] == max_mm) cl = 0;       /* two no-return neighbours: see the per-ray maps */
                else cl = !cnf_iou_pos(S.hx[a], S.hy[a], S.hx[b], S.hy[b], bbox);
            }
            S.close[i] = (uint8_t)cl;
            hflag[i] = (uint8_t)(S.rmm[S.src[i]] != max_mm);
            if (cl) { ++cnt; if (i < first) first = i; if (i != n - 1 && i > last) last = i; }
        }
        seg_first = first; seg_last = last; seg_cnt = cnt;
    }
    CNF_STAMP(5);
    cnf_min_max_sum(CNF_COLL_PASS, 16, &seg_first, &seg_last, &seg_cnt);   /* (its barrier also publishes close[] / hflag[]) */
    CNF_STAMP(6);
    const int e0 = seg_first, zb = seg_last + 1, nseg = seg_cnt;
    int merged = 0;
    if (nseg > 1) {
        const int a = S.src[0], b = S.src[n - 1];
        merged = cnf_iou_pos(S.hx[a], S.hy[a], S.hx[b], S.hy[b], bbox * 2.0);
    }
    const int len_a = e0 + 1, len_z = merged ? n - zb : 0;
#define CNF_FLAT_OF(k) ((k) < len_a ? (k) : ((k) < len_a + len_z ? zb + ((k) - len_a) : (k) - len_z))
    /* a sub-segment ends where its segment ends or where hit and no-hit records meet (ENV:510-571: a segment
     * without any hit has no such place, so it stays whole); running counts of 'o' / 'w' records along the order */
    int16_t* cum_o = (int16_t*)S.gk;                             /* [n + 1] each; the gradients are consumed */
    int16_t* cum_w = cum_o + (n + 1);
    uint8_t* subend = S.subend;
    int p_end = 0, p_o = 0, p_w = 0;
    {
        int n_end = 0, n_o = 0, n_w = 0;
        CNF_ROLLED for (int k = c_lo; k < c_hi; ++k) {
            const int r = CNF_FLAT_OF(k);
            int end;
            if (merged && k < len_a + len_z) end = (k == len_a + len_z - 1);
            else end = S.close[r];
            if (!end) { const int r1 = CNF_FLAT_OF(k + 1); end = hflag[r] != hflag[r1]; }
            subend[k] = (uint8_t)end;
            n_end += end; n_o += (S.type[r] == CNF_T_O); n_w += (S.type[r] == CNF_T_W);
        }
        p_end = n_end; p_o = n_o; p_w = n_w;
    }
    CNF_STAMP(7);
    int nsub_total = 0;
    {
        int b_end = p_end, b_o = p_o, b_w = p_w;
        cnf_scan_excl3(CNF_COLL_PASS, 32, &b_end, &b_o, &b_w, &nsub_total);   /* (its barrier also publishes subend[]) */
        if (lane == 0) { S.sub[0] = 0; cum_o[0] = 0; cum_w[0] = 0; }
        CNF_ROLLED for (int k = c_lo; k < c_hi; ++k) {
            const int r = CNF_FLAT_OF(k);
            b_o += (S.type[r] == CNF_T_O); b_w += (S.type[r] == CNF_T_W);
            cum_o[k + 1] = (int16_t)b_o; cum_w[k + 1] = (int16_t)b_w;
            if (subend[k]) S.sub[++b_end] = (int16_t)(k + 1);
        }
    }
    CNF_SYNC();
    CNF_STAMP(8);
    const int nsub = nsub_total;
    /* confirmation (ENV:573-620, UTL:395-402), one lane per sub-segment; every record of a sub-segment is a hit or
     * none is, so its first one decides */
    int16_t* verdict = (int16_t*)S.subend;                       /* [nsub] <= n over subend[] + type[], both consumed */
    {
        const double span = P->max_range - P->min_range;
        CNF_ROLLED for (int s = lane; s < nsub; s += nl) {
            const int sb = S.sub[s], se = S.sub[s + 1], sl = se - sb;
            int v = -1;
            if (sl >= 4 && hflag[CNF_FLAT_OF(sb)]) {
                const int n_o = cum_o[se] - cum_o[sb], n_w = cum_w[se] - cum_w[sb], n_none = sl - n_o - n_w;
                const int kc = sb + sl / 2;
                const int rc = S.src[CNF_FLAT_OF(kc)];
                const double dctr = CNF_MILLI(S.rmm[rc]);
                const double estd = 3.0 + floor(cnf_div(29.0 * (P->max_range - dctr), span));
                const double denom = ((double)sl < estd) ? (double)sl : estd;
                const double score = cnf_div((double)n_o, denom);
                const int distinct = (n_o > 0) + (n_w > 0) + (n_none > 0);
                int t = -1;
                if (distinct > 1) {
                    if (score >= 0.5) t = (n_o > n_w) ? CNF_T_O : CNF_T_W;
                    else if ((double)sl <= estd) t = (n_o > n_w) ? CNF_T_O : CNF_T_W;
                    else t = CNF_T_W;
                } else {
                    const double lim = ((double)nsub < estd) ? (double)nsub : estd;
                    if (!((double)sl <= lim)) t = (n_w > 0) ? CNF_T_W : CNF_T_O;
                }
                if (t >= 0) v = t | (rc << 2);                   /* rc < 1024 */
            }
            verdict[s] = (int16_t)v;
        }
    }
#undef CNF_FLAT_OF
    CNF_SYNC();
    CNF_STAMP(9);
    if (lane == 0) {
        int nconf = 0, n_obst = 0, ego_hit = 0;
        CNF_ROLLED for (int s = 0; s < nsub; ++s) {
            const int v = verdict[s];
            if (v < 0) continue;
            if (nconf < CNF_CONF_CAP) {
                const int rc = v >> 2;
                int32_t* c = S.conf + 4 * nconf;
                c[0] = v & 3; c[1] = S.hx[rc]; c[2] = S.hy[rc]; c[3] = S.rmm[rc];
                ++nconf;
            } else S.trk[CNF_H_OVERFLOW] += 1;
        }
        CNF_ROLLED for (int c = 0; c < nconf; ++c)
            if (S.conf[4 * c] == CNF_T_O) { ++n_obst; if (CNF_MILLI(S.conf[4 * c + 3]) < 0.140) ego_hit = 1; }
        if (n_obst > 0) S.trk[CNF_H_PRESENT] += 1;              /* ENV:653-654 */
        S.misc[0] = nconf; S.misc[1] = ego_hit;
        /* popleft of every tracked deque (ENV:679-682) happens before the IoUs are taken */
        const int n0 = (int)S.trk[CNF_H_N];
        if (nconf > 0)
            CNF_ROLLED for (int i = 0; i < n0; ++i) {
                uint32_t* q = S.trk + CNF_HDR_WORDS + i * CNF_ENTRY_WORDS;
                if (q[CNF_E_NDEQ] > 1u) { q[CNF_E_PX] = q[CNF_E_LX]; q[CNF_E_PY] = q[CNF_E_LY]; q[CNF_E_NDEQ] = 1u; }
            }
    }
    CNF_SYNC();
    CNF_STAMP(10);


    /* ---- tracker: best confirmed object per tracked one (ENV:691-703), one lane each ---- */
    const int nconf = S.misc[0];
    const int n0 = (int)S.trk[CNF_H_N];
    if (nconf > 0)
        CNF_ROLLED for (int i = lane; i < n0; i += nl) {
            const uint32_t* q = S.trk + CNF_HDR_WORDS + i * CNF_ENTRY_WORDS;
            int m = 0; double best = 0.0;
            CNF_ROLLED for (int c = 0; c < nconf; ++c) {
                const double v = cnf_iou((int32_t)q[CNF_E_LX], (int32_t)q[CNF_E_LY], S.conf[4 * c + 1], S.conf[4 * c + 2], P->track_half);
                if (c == 0 || v > best) { best = v; m = c; }
            }
            S.am_val[i] = best; S.am_idx[i] = m;
        }
    CNF_SYNC();
    CNF_STAMP(11);

    if (lane == 0) {
        uint32_t* E = S.trk + CNF_HDR_WORDS;
        int n_ent = n0;
        uint64_t checked = 0;                                    /* CNF_CONF_CAP <= 64 */
        if (n0 > 0 && nconf == 0) n_ent = 0;                     /* ENV:686-689 */
        else if (n0 > 0) {
            /* update / drop in dict order; `len(dict) > i` (ENV:718) lets late unmatched entries survive */
            uint32_t dead = 0; int n_live = n0;
            CNF_ROLLED for (int i = 0; i < n0; ++i) {
                uint32_t* q = E + i * CNF_ENTRY_WORDS;
                if (S.am_val[i] > 0.0) {
                    const int32_t* c = S.conf + 4 * S.am_idx[i];
                    q[CNF_E_PX] = q[CNF_E_LX]; q[CNF_E_PY] = q[CNF_E_LY];
                    q[CNF_E_LX] = (uint32_t)c[1]; q[CNF_E_LY] = (uint32_t)c[2]; q[CNF_E_DIST] = (uint32_t)c[3];
                    q[CNF_E_NDEQ] = 2u;
                    checked |= (uint64_t)1 << S.am_idx[i];
                } else if (n_live > i) { dead |= 1u << i; --n_live; }
            }
            int m = 0;
            CNF_ROLLED for (int i = 0; i < n0; ++i) {
                if (dead & (1u << i)) continue;
                if (m != i) {
                    CNF_ROLLED for (int k = 0; k < CNF_ENTRY_WORDS; ++k) E[m * CNF_ENTRY_WORDS + k] = E[i * CNF_ENTRY_WORDS + k];
                }
                ++m;
            }
            n_ent = m;
        }
        if (!(n0 > 0 && nconf == 0))
            CNF_ROLLED for (int c = 0; c < nconf; ++c) {                    /* new tracked objects (ENV:663-671, 725-741) */
                if (S.conf[4 * c] != CNF_T_O || ((checked >> c) & 1)) continue;
                if (n_ent >= CNF_TRK_CAP) { S.trk[CNF_H_OVERFLOW] += 1; continue; }
                uint32_t* q = E + n_ent * CNF_ENTRY_WORDS;
                q[CNF_E_PX] = q[CNF_E_LX] = (uint32_t)S.conf[4 * c + 1];
                q[CNF_E_PY] = q[CNF_E_LY] = (uint32_t)S.conf[4 * c + 2];
                q[CNF_E_DIST] = (uint32_t)S.conf[4 * c + 3]; q[CNF_E_NDEQ] = 1u;
                cnf_st64(q + CNF_E_SPEED, -1.0); cnf_st64(q + CNF_E_VX, 0.0); cnf_st64(q + CNF_E_VY, 0.0);
                ++n_ent;
            }
        S.trk[CNF_H_N] = (uint32_t)n_ent;
        /* speed (ENV:745-760), obstacle velocity and the probe target of the collision cone (ENV:799-815) */
        const long long pose_kx = cnf_round3k(x), pose_ky = cnf_round3k(y);   /* round(x, 3), round(y, 3) in thousandths ... */
        S.misc[5] = (int32_t)pose_kx; S.misc[6] = (int32_t)pose_ky;           /* ... needed again by the last stage (lane 0) */
        const double curx = cn_milli64(pose_kx), cury = cn_milli64(pose_ky);
        double vox = curx, voy = cury;
        CNF_ROLLED for (int i = 0; i < n_ent; ++i) {
            uint32_t* q = E + i * CNF_ENTRY_WORDS;
            double cx = 0.0, cy = 0.0;
            if (q[CNF_E_NDEQ] > 1u) {
                const double px = CNF_MILLI(q[CNF_E_PX]), py = CNF_MILLI(q[CNF_E_PY]);
                const double lx = CNF_MILLI(q[CNF_E_LX]), ly = CNF_MILLI(q[CNF_E_LY]);
                cnf_st64(q + CNF_E_SPEED, cnf_div(cnf_hypot(py - ly, px - lx), P->dt));
                if (S.trk[CNF_H_HAVE_PREV]) {
                    cx = px - lx; cy = py - ly;                  /* last - curr (sic, ENV:806-807) */
                    cnf_st64(q + CNF_E_VX, cnf_div(cx, P->dt)); cnf_st64(q + CNF_E_VY, cnf_div(cy, P->dt));
                }
            }
            vox = curx + cx; voy = cury + cy;                    /* leaks out of the loop (ENV:814-815) */
        }
        S.hit[0] = vox; S.hit[1] = voy;                          /* handed to all lanes through memory */
        S.misc[2] = n_ent;
    }
    CNF_SYNC();
    CNF_STAMP(12);

    /* ---- collision cone: distance to the r = 0.178 ring along the probe lines (UTL:251-293) ---- */
    const int n_ent = S.misc[2];
    CNF_ROLLED for (int k = CNF_HDR_WORDS + n_ent * CNF_ENTRY_WORDS + lane; k < CNF_WORLD_WORDS; k += nl) S.trk[k] = 0u;   /* unused entries read as zero */
    const int have_prev = (int)S.trk[CNF_H_HAVE_PREV];
    const double a0x = CNF_MILLI(S.trk[CNF_H_PPX]), a0y = CNF_MILLI(S.trk[CNF_H_PPY]);
    if (have_prev && n_ent > 0) {
        const double a1x = S.hit[0], a1y = S.hit[1];
        CNF_SYNC();
        double gradient = 0.0;
        if (a1y != 0.0) gradient = cnf_div(a1x - a0x, a1y) - a0y;      /* precedence as written (UTL:261) */
        const double cb = a0x - (gradient * a0y);
        const long x_hi = (long)ceil(a0x + 3.5), x_lo = (long)floor(a0x - 3.5);
        CNF_ROLLED for (int i = 0; i < n_ent; ++i) {
            const uint32_t* q = S.trk + CNF_HDR_WORDS + i * CNF_ENTRY_WORDS;
            const double ox = CNF_MILLI(q[CNF_E_LX]), oy = CNF_MILLI(q[CNF_E_LY]);
            int have = 0; double dtc = 0.0;
            CNF_ROLLED for (long x2 = x_hi; x2 > x_lo; --x2) {
                const double qx = (double)x2, qy = ((double)x2 * gradient) + cb;
                const double rx = qx - a0x, ry = qy - a0y;
                {
                    /* every vertex of the ring lies on the circle: a probe segment that stays farther than the radius
                     * (+1e-6, far above round-off) from its centre cannot meet it -- 'LINESTRING EMPTY' without
                     * testing the 64 edges.  All lanes compute the same bits, so the decision is uniform. */
                    const double len2 = rx * rx + ry * ry;
                    if (len2 > 0.0) {
                        double tp = cnf_div((ox - a0x) * rx + (oy - a0y) * ry, len2);
                        tp = tp < 0.0 ? 0.0 : (tp > 1.0 ? 1.0 : tp);
                        const double ex = ox - (a0x + tp * rx), ey = oy - (a0y + tp * ry);
                        const double rm = P->cp_radius * 1.000001;
                        if (ex * ex + ey * ey > rm * rm) continue;
                    }
                }
                CNF_ROLLED for (int k = lane; k < 64; k += nl) {
                    const int k1 = (k + 1) & 63;
                    const double ax = ox + P->cp_radius * CNF_RING_COS[k], ay = oy + P->cp_radius * CNF_RING_SIN[k];
                    const double bx = ox + P->cp_radius * CNF_RING_COS[k1], by = oy + P->cp_radius * CNF_RING_SIN[k1];
                    const double sx = bx - ax, sy = by - ay;
                    const double den = rx * sy - ry * sx;
                    int f = 0;
                    if (den != 0.0) {
                        const double nt = (ax - a0x) * sy - (ay - a0y) * sx, nu = (ax - a0x) * ry - (ay - a0y) * rx;
                        /* t = nt / den and u = nu / den must both lie in [0, 1].  62 of the 64 edges miss by a wide margin:
                         * a numerator outside [0, |den|] by more than 1e-9 |den| (the quotient's round-off is 1e-16) is a
                         * certain miss and needs no division */
                        const double ad = fabs(den), st = den > 0.0 ? nt : -nt, su = den > 0.0 ? nu : -nu;
                        const double lo = -1.0e-9 * ad, hi = ad * (1.0 + 1.0e-9);
                        if (!(st < lo || st > hi || su < lo || su > hi)) {
                            const double t = cnf_div(nt, den);
                            const double u = cnf_div(nu, den);
                            if (0.0 <= t && t <= 1.0 && 0.0 <= u && u <= 1.0) { f = 1; S.hit[2 * k] = a0x + t * rx; S.hit[2 * k + 1] = a0y + t * ry; }
                        }
                    }
                    S.hitf[k] = (uint8_t)f;
                }
                CNF_SYNC();
                /* every lane resolves the (few) hits identically: no broadcast needed.  Most probes miss the ring:
                 * the 64 flags are first looked at as 8 words. */
                int nh = 0, i0 = -1, i1 = -1, first[4] = {0, 0, 0, 0}; double k0 = 0.0, k1v = 0.0;
                CNF_ROLLED for (int w8 = 0; w8 < 8; ++w8) {
                    unsigned long long word; memcpy(&word, S.hitf + 8 * w8, 8);
                    if (word == 0ull) continue;
                    CNF_ROLLED for (int k = 8 * w8; k < 8 * w8 + 8; ++k) {
                        if (!S.hitf[k]) continue;
                        int dup = 0;
                        CNF_ROLLED for (int o = 0; o < nh && o < 4; ++o)
                            if (fabs(S.hit[2 * k] - S.hit[2 * first[o]]) < 1e-12 && fabs(S.hit[2 * k + 1] - S.hit[2 * first[o] + 1]) < 1e-12) { dup = 1; break; }
                        if (dup) continue;
                        if (nh < 4) first[nh] = k;
                        ++nh;
                        const double dx = S.hit[2 * k] - a0x, dy = S.hit[2 * k + 1] - a0y;
                        const double key = dx * dx + dy * dy;
                        if (i0 < 0 || key < k0) { i1 = i0; k1v = k0; i0 = k; k0 = key; }
                        else if (i1 < 0 || key < k1v) { i1 = k; k1v = key; }
                    }
                }
                int stop = 0;
                if (nh == 1) { have = 0; stop = 1; }              /* a Point has no .geoms -> None */
                else if (nh >= 2) {
                    const double d0 = cnf_hypot(a0x - S.hit[2 * i0], a0y - S.hit[2 * i0 + 1]);
                    const double d1 = cnf_hypot(a0x - S.hit[2 * i1], a0y - S.hit[2 * i1 + 1]);
                    dtc = d0 < d1 ? d0 : d1; have = 1; stop = 1;
                }
                CNF_SYNC();                                       /* hit[] is rewritten by the next probe */
                if (stop) break;
            }
            if (lane == 0) { S.am_val[i] = dtc; S.am_idx[i] = have; }
        }
    }
    CNF_SYNC();
    CNF_STAMP(13);

    /* ---- lane 0: collision probability, ranking, K block, counters (ENV:765-907, 998-1005) ---- */
    if (lane == 0) {
        uint32_t* E = S.trk + CNF_HDR_WORDS;
        double ego_score = cnf_ld64(S.trk + CNF_H_EGOSCORE);
        const long long kx = S.misc[5], ky = S.misc[6];
        const double curx = cn_milli64(kx), cury = cn_milli64(ky);
        const float padx = (float)cnf_np_round3(x), pady = (float)cnf_np_round3(y);     /* once, not once per slot */
        CNF_ROLLED for (int s = 0; s < K; ++s) {
            kblock[4 * s] = padx; kblock[4 * s + 1] = pady;
            kblock[4 * s + 2] = 0.0f; kblock[4 * s + 3] = 0.0f;
        }
        if (have_prev) {
            const double vx = cnf_div(curx - a0x, P->dt), vy = cnf_div(cury - a0y, P->dt);     /* UTL:227-236 */
            const double agent_vel = sqrt(vx * vx + vy * vy);
            const double obstacle_vel = (n_ent == 0) ? 0.0 : cnf_ld64(E + CNF_E_SPEED);   /* ENV:789-797 */
            const double resultant = agent_vel - obstacle_vel;
            const double span = P->max_range - P->min_range;
            double ego_cur = 0.0, ego_max = 0.0;
            /* the collision probabilities overwrite am_val in place (the distance is consumed first) */
            CNF_ROLLED for (int i = 0; i < n_ent; ++i) {
                const double dist = CNF_MILLI(E[i * CNF_ENTRY_WORDS + CNF_E_DIST]);
                const double dto = (dist > P->max_range) ? 0.0 : cnf_div(P->max_range - dist, span);
                double cp;
                if (S.am_idx[i]) {
                    if (resultant == 0.0) cp = 1.0 * dto;
                    else {
                        const double ttc = cnf_div(S.am_val[i], resultant);
                        const double qq = cnf_div(0.15, ttc);
                        ego_cur = (1.0 < qq) ? 1.0 : qq;
                        cp = 0.5 * ego_cur + 0.5 * dto;
                    }
                } else { ego_cur = 0.0; cp = 0.5 * 0.0 + 0.5 * dto; }
                S.am_val[i] = cp;
                if (i == 0 || ego_cur > ego_max) ego_max = ego_cur;
            }
            ego_score = (n_ent == 0) ? 0.0 : ego_max;
            CNF_ROLLED for (int a = 0; a < (K > 0 ? n_ent : 0); ++a) {       /* stable descending sort, keep [-K:] (ENV:882-883) */
                int rank = 0;
                CNF_ROLLED for (int b = 0; b < n_ent; ++b)
                    if (b != a && (S.am_val[b] > S.am_val[a] || (S.am_val[b] == S.am_val[a] && b < a))) ++rank;
                const int slot = P->topk_highest ? rank : rank - (n_ent > K ? n_ent - K : 0);
                if (slot < 0 || slot >= K) continue;
                const uint32_t* q = E + a * CNF_ENTRY_WORDS;
                kblock[4 * slot] = (float)cnf_np_round3(CNF_MILLI(q[CNF_E_LX]));
                kblock[4 * slot + 1] = (float)cnf_np_round3(CNF_MILLI(q[CNF_E_LY]));
                kblock[4 * slot + 2] = (float)cnf_np_round3(cnf_ld64(q + CNF_E_VX));
                kblock[4 * slot + 3] = (float)cnf_np_round3(cnf_ld64(q + CNF_E_VY));
            }
        }
        cnf_st64(S.trk + CNF_H_EGOSCORE, ego_score);
        S.trk[CNF_H_HAVE_PREV] = 1u;
        S.trk[CNF_H_PPX] = (uint32_t)(int32_t)kx;
        S.trk[CNF_H_PPY] = (uint32_t)(int32_t)ky;
        if (S.misc[1]) S.trk[CNF_H_EGO] += 1;
        if (ego_score > 0.4) S.trk[CNF_H_SOCIAL] += 1;
        if (step_counter == 0) { S.trk[CNF_H_EGO] = 0; S.trk[CNF_H_SOCIAL] = 0; S.trk[CNF_H_PRESENT] = 0; }   /* ENV:1258-1262 */
    }
    CNF_SYNC();
    CNF_STAMP(14);
}

#endif /* CN_FAITHFUL_H */
