/*
 * cn_oracle_faithful.c -- CPU restatement of the reference's OWN perception block
 * (`risk_faithful`): gradient typing, scan segmentation,
 * segment confirmation, the uuid-dict tracker, collision cone, collision
 * probability, top-K block and the safety counters of
 * turtlebot3_rl_sim/src/environment_stage_1_nobonus.py:270-1005, with
 * utils.py:110-126, 227-236, 251-293, 317-345, 395-460.
 *
 * TEST INFRASTRUCTURE (see oracle/oracle.py).  Written as a literal, list-by-list
 * walk through the Python, in float64 like CPython; it shares only the primitive
 * headers cn_math.h / cn_math64.h and the per-world tracker record layout with the
 * product (crowdnav_b200/csrc/cn_faithful.h is an independent, flat-array design).
 *
 * Pinned: tests/test_faithful.py replays the reference-in-the-loop traces
 * (tests/golden/trace_*.npz: odometry + raw scans in, the reference's state row out)
 * and compares the K block element by element.
 *
 * What is NOT literal (SURVEY quirks ledger): wall-clock dt -> the fixed control period;
 * uuid keys -> insertion order (CPython >= 3.7 dict order, which is what the harness ran);
 * tracker cleared at reset; at most CNF_TRK_CAP tracked / CNF_CONF_CAP confirmed objects.
 */
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include "../crowdnav_b200/csrc/cn_math64.h"
#include "../crowdnav_b200/csrc/cn_faithful_state.h"

enum { T_NONE = 0, T_W = 1, T_O = 2 };

typedef struct { double x, y; } pt;
typedef struct { int type; double dist; pt pose; } item;          /* [type, range, pose] of ENV:382-404 */
typedef struct { int type; pt pose; double dist; } conf_obj;      /* confirmed_scan_object entry, ENV:590-617 */

/* UTL:421-448: IoU of two axis-aligned squares of half-size b, rounded to 3 dp */
static double iou_boxes(pt a, pt b, double h) {
    double ax0 = a.x - h, ax1 = a.x + h, ay0 = a.y - h, ay1 = a.y + h;
    double bx0 = b.x - h, bx1 = b.x + h, by0 = b.y - h, by1 = b.y + h;
    double w = (ax1 < bx1 ? ax1 : bx1) - (ax0 > bx0 ? ax0 : bx0);
    double hh = (ay1 < by1 ? ay1 : by1) - (ay0 > by0 ? ay0 : by0);
    double inter = (w > 0.0 && hh > 0.0) ? w * hh : 0.0;
    double area_a = (ax1 - ax0) * (ay1 - ay0), area_b = (bx1 - bx0) * (by1 - by0);
    double uni = area_a + area_b - inter;
    return cn_py_round3_64(inter / uni);
}
static int is_associated(pt a, pt b, double h) { return iou_boxes(a, b, h) > 0.0; }   /* UTL:435-448 */

/* UTL:110-126 for observation ray i */
static pt hit_point(const cnf_params* p, double x, double y, double yaw, int i, double r) {
    double ang = ((double)i * p->inc_deg) * CN64_DEG2RAD - yaw;
    double s, c; cn_sincos64(ang, &s, &c);
    pt o;
    o.x = cn_py_round3_64(x + r * c);
    o.y = cn_py_round3_64(y + (r * s) * -1.0);
    return o;
}

/* UTL:405-419 */
static double bounding_box_size(const cnf_params* p, double x, double y, double yaw) {
    int n = p->n_rays;
    pt* g = (pt*)malloc(sizeof(pt) * (size_t)n);
    for (int i = 0; i < n; ++i) g[i] = hit_point(p, x, y, yaw, i, p->max_range);
    double sum = 0.0;
    for (int i = 0; i < n; ++i) {
        int j = (i == n - 1) ? 0 : i + 1;
        sum += cn_hypot64(g[i].x - g[j].x, g[i].y - g[j].y);
    }
    free(g);
    return sum / (double)n;
}

/* segment / segment intersection as the shapely stand-in states it (tests/ref_harness.py _seg_intersect) */
static int seg_intersect(pt p, pt q, pt a, pt b, pt* out) {
    double rx = q.x - p.x, ry = q.y - p.y, sx = b.x - a.x, sy = b.y - a.y;
    double den = rx * sy - ry * sx;
    if (den == 0.0) return 0;
    double t = ((a.x - p.x) * sy - (a.y - p.y) * sx) / den;
    double u = ((a.x - p.x) * ry - (a.y - p.y) * rx) / den;
    if (0.0 <= t && t <= 1.0 && 0.0 <= u && u <= 1.0) { out->x = p.x + t * rx; out->y = p.y + t * ry; return 1; }
    return 0;
}

/* UTL:251-293 get_collision_point: returns 1 and *dist, or 0 for None */
static int collision_point(pt a0, pt a1, pt obs, double radius, double* dist) {
    pt ring[64];
    for (int k = 0; k < 64; ++k) { ring[k].x = obs.x + radius * CNF_RING_COS[k]; ring[k].y = obs.y + radius * CNF_RING_SIN[k]; }
    double gradient;
    if (a1.y == 0.0) gradient = 0.0;                              /* ZeroDivisionError -> 0 (UTL:260-263) */
    else gradient = (a1.x - a0.x) / a1.y - a0.y;                  /* precedence as written (UTL:261) */
    double cb = a0.x - (gradient * a0.y);
    long x_hi = (long)ceil(a0.x + 3.5), x_lo = (long)floor(a0.x - 3.5);
    for (long x2 = x_hi; x2 > x_lo; --x2) {
        pt q; q.x = (double)x2; q.y = ((double)x2 * gradient) + cb;
        pt hits[64]; int nh = 0;
        for (int k = 0; k < 64; ++k) {
            pt h;
            if (!seg_intersect(a0, q, ring[k], ring[(k + 1) % 64], &h)) continue;
            int dup = 0;
            for (int o = 0; o < nh; ++o)
                if (fabs(h.x - hits[o].x) < 1e-12 && fabs(h.y - hits[o].y) < 1e-12) { dup = 1; break; }
            if (!dup) hits[nh++] = h;
        }
        if (nh == 0) continue;                                     /* 'LINESTRING EMPTY': next x2 */
        if (nh == 1) return 0;                                     /* a Point has no .geoms -> except -> None */
        /* MultiPoint sorted by squared distance from the line start; the two nearest (stable) */
        int i0 = -1, i1 = -1; double k0 = 0.0, k1 = 0.0;
        for (int o = 0; o < nh; ++o) {
            double dx = hits[o].x - a0.x, dy = hits[o].y - a0.y;
            double key = dx * dx + dy * dy;
            if (i0 < 0 || key < k0) { i1 = i0; k1 = k0; i0 = o; k0 = key; }
            else if (i1 < 0 || key < k1) { i1 = o; k1 = key; }
        }
        double d0 = cn_hypot64(a0.x - hits[i0].x, a0.y - hits[i0].y);
        double d1 = cn_hypot64(a0.x - hits[i1].x, a0.y - hits[i1].y);
        *dist = d0 < d1 ? d0 : d1;
        return 1;
    }
    return 0;
}

/* tracked_obstacles entry (ENV:662-671): [type, pose, dist, deque(<=2), t, speed, [vx, vy]] */
typedef struct { pt prev, last; double dist; int ndeq; double speed, vx, vy; } trk_entry;

static void load_entries(const uint32_t* w, trk_entry* e, int n) {
    for (int i = 0; i < n; ++i) {
        const uint32_t* q = w + CNF_HDR_WORDS + i * CNF_ENTRY_WORDS;
        e[i].prev.x = cn_milli64((int32_t)q[CNF_E_PX]); e[i].prev.y = cn_milli64((int32_t)q[CNF_E_PY]);
        e[i].last.x = cn_milli64((int32_t)q[CNF_E_LX]); e[i].last.y = cn_milli64((int32_t)q[CNF_E_LY]);
        e[i].dist = cn_milli64((int32_t)q[CNF_E_DIST]); e[i].ndeq = (int)q[CNF_E_NDEQ];
        memcpy(&e[i].speed, q + CNF_E_SPEED, 8); memcpy(&e[i].vx, q + CNF_E_VX, 8); memcpy(&e[i].vy, q + CNF_E_VY, 8);
    }
}
static int32_t milli_of(double v) { return (int32_t)llrint(v * 1000.0); }
static void store_entries(uint32_t* w, const trk_entry* e, int n) {
    memset(w + CNF_HDR_WORDS, 0, sizeof(uint32_t) * CNF_ENTRY_WORDS * CNF_TRK_CAP);
    for (int i = 0; i < n; ++i) {
        uint32_t* q = w + CNF_HDR_WORDS + i * CNF_ENTRY_WORDS;
        q[CNF_E_PX] = (uint32_t)milli_of(e[i].prev.x); q[CNF_E_PY] = (uint32_t)milli_of(e[i].prev.y);
        q[CNF_E_LX] = (uint32_t)milli_of(e[i].last.x); q[CNF_E_LY] = (uint32_t)milli_of(e[i].last.y);
        q[CNF_E_DIST] = (uint32_t)milli_of(e[i].dist); q[CNF_E_NDEQ] = (uint32_t)e[i].ndeq;
        memcpy(q + CNF_E_SPEED, &e[i].speed, 8); memcpy(q + CNF_E_VX, &e[i].vx, 8); memcpy(q + CNF_E_VY, &e[i].vy, 8);
    }
    w[CNF_H_N] = (uint32_t)n;
}

/*
 * One get_state (ENV:245-1044) worth of the perception block for one world.
 *   trk    : the world's tracker record (cn_faithful_state.h), updated in place
 *   x,y,yaw: odometry (ENV:239-243)       scan: cleaned ranges, observation order (UTL:375-392), 0.6 = no return
 *   step_counter: 0 inside reset() (ENV:1245), else the driver's 1-based step
 *   kblock : 4K doubles, np.around(., 3) applied (ENV:1042)
 */
void orf_observe(const cnf_params* p, uint32_t* trk, double x, double y, double yaw, const double* scan,
                 int step_counter, double* kblock) {
    const int n = p->n_rays, K = p->k_obstacles;
    pt* pose = (pt*)malloc(sizeof(pt) * (size_t)n);
    double* fr = (double*)malloc(sizeof(double) * (size_t)n);       /* filtered_scan_ranges: round(scan, 3) */
    double* grad = (double*)malloc(sizeof(double) * (size_t)n); char* grad_ok = (char*)malloc((size_t)n);
    double* chg = (double*)malloc(sizeof(double) * (size_t)n); char* chg_ok = (char*)malloc((size_t)n);
    int* rec_type = (int*)malloc(sizeof(int) * (size_t)n); int* rec_src = (int*)malloc(sizeof(int) * (size_t)n);
    item* est = (item*)malloc(sizeof(item) * (size_t)n);
    int* flat = (int*)malloc(sizeof(int) * (size_t)n * 2);
    int* seg_off = (int*)malloc(sizeof(int) * ((size_t)n + 2));
    int* sub_off = (int*)malloc(sizeof(int) * ((size_t)n + 2));

    if (step_counter == 0) {
        /* a fresh Env per episode (quirks ledger: tracker / deques cleared at reset) */
        memset(trk, 0, sizeof(uint32_t) * CNF_WORLD_WORDS);
    }
    for (int i = 0; i < n; ++i) { pose[i] = hit_point(p, x, y, yaw, i, scan[i]); fr[i] = cn_py_round3_64(scan[i]); }
    double bbox;
    pt cur; cur.x = cn_py_round3_64(x); cur.y = cn_py_round3_64(y);
    if (step_counter == 0) {                                       /* ENV:286-294 */
        bbox = bounding_box_size(p, x, y, yaw);
        memcpy(trk + CNF_H_BBOX, &bbox, 8);
    } else {
        memcpy(&bbox, trk + CNF_H_BBOX, 8);
    }

    /* E: gradients (ENV:329-347) */
    for (int i = 0; i < n; ++i) {
        if (fr[i] == p->max_range) { grad_ok[i] = 0; grad[i] = 0.0; continue; }
        int j = (i == n - 1) ? 0 : i + 1;
        double g;
        if ((pose[i].y - pose[j].y) == 0.0) g = 0.0;
        else g = (pose[i].x - pose[j].x) / (pose[i].y - pose[j].y);
        grad[i] = cn_py_round3_64(g); grad_ok[i] = 1;
    }
    /* change of gradient (ENV:349-368) */
    {
        double last = 0.0; int last_ok = 0;
        for (int i = 0; i < n; ++i) {
            if (!grad_ok[i]) { chg_ok[i] = 0; chg[i] = 0.0; }
            else if (n == 1 || i == n - 1) { chg[i] = last; chg_ok[i] = (char)last_ok; }
            else if (grad_ok[i + 1]) { last = fabs(grad[i] - grad[i + 1]); last_ok = 1; chg[i] = last; chg_ok[i] = 1; }
            else { last_ok = 0; last = 0.0; chg[i] = 0.0; chg_ok[i] = 0; }
        }
    }
    /* wall / obstacle typing with the delayed-update counter (ENV:370-404).  A record is (type, source ray):
     * `_scans_object_type[i] = last_type` hands ray i the range and pose of an EARLIER ray. */
    {
        int last_type = T_NONE, last_src = -1, du = 0;
        for (int i = 0; i < n; ++i) {
            rec_type[i] = T_NONE; rec_src[i] = i;
            if (!chg_ok[i]) continue;
            if (i == n - 1) continue;
            if (chg[i] == 0.0) { rec_type[i] = T_W; last_type = T_W; last_src = i; continue; }
            rec_type[i] = T_O;
            if (du != 1) {
                if (chg_ok[i + 1] && chg[i + 1] == 0.0) { rec_type[i] = T_W; last_type = T_W; last_src = i; du = 0; }
                if (!chg_ok[i + 1]) {
                    /* pass */
                } else if (fabs(chg[i] - chg[i + 1]) == 0.0) {
                    rec_type[i] = T_W; rec_src[i] = i; last_type = T_W; last_src = i; du = 0;
                } else {
                    rec_type[i] = last_type; rec_src[i] = (last_type == T_NONE) ? i : last_src; du += 1;
                }
            } else {
                rec_type[i] = T_O; last_type = T_O; last_src = i;
                if (chg_ok[i + 1] && chg[i + 1] == 0.0) du = 0;
            }
        }
    }
    /* ENV:428-441 */
    for (int i = 0; i < n; ++i) {
        if (rec_type[i] == T_NONE) { est[i].type = T_NONE; est[i].dist = fr[i]; est[i].pose = pose[i]; }
        else { int s = rec_src[i]; est[i].type = rec_type[i]; est[i].dist = cn_py_round3_64(fr[s]); est[i].pose = pose[s]; }
    }
    /* F: segmentation (ENV:443-486): a segment closes after ray i when i and i+1 are not associated */
    int nseg = 0; seg_off[0] = 0;
    {
        int len = 0;
        for (int i = 0; i < n; ++i) {
            flat[len++] = i;
            int close = (i == n - 1) ? 1 : !is_associated(est[i].pose, est[i + 1].pose, bbox);
            if (close) { seg_off[++nseg] = len; }
        }
    }
    /* first / last merge across the blind spot (ENV:488-504): seg[0] = seg[0] + seg[-1] */
    if (nseg > 1) {
        int f0 = flat[seg_off[0]], l1 = flat[seg_off[nseg] - 1];
        if (is_associated(est[f0].pose, est[l1].pose, bbox * 2.0)) {
            int a_len = seg_off[1] - seg_off[0], z_beg = seg_off[nseg - 1], z_len = seg_off[nseg] - z_beg;
            int* tmp = (int*)malloc(sizeof(int) * (size_t)n);
            int len = 0;
            for (int k = 0; k < a_len; ++k) tmp[len++] = flat[k];
            for (int k = 0; k < z_len; ++k) tmp[len++] = flat[z_beg + k];
            int first_end = len;
            for (int k = seg_off[1]; k < z_beg; ++k) tmp[len++] = flat[k];
            int shift = z_len;
            for (int s = 1; s < nseg - 1; ++s) seg_off[s + 1] = seg_off[s + 1] + shift;   /* ends of middle segments */
            seg_off[1] = first_end;
            nseg -= 1;
            memcpy(flat, tmp, sizeof(int) * (size_t)len);
            free(tmp);
        }
    }
    /* split at 0.6 <-> hit transitions (ENV:510-556) and flatten (ENV:558-571) */
    int nsub = 0; sub_off[0] = 0;
    for (int s = 0; s < nseg; ++s) {
        int b = seg_off[s], e = seg_off[s + 1];
        int any_hit = 0;
        for (int k = b; k < e; ++k) if (est[flat[k]].dist != p->max_range) any_hit = 1;
        if (!any_hit) { sub_off[++nsub] = e; continue; }
        for (int k = b; k < e; ++k) {
            int close;
            if (k == e - 1) close = 1;
            else {
                int a6 = est[flat[k]].dist == p->max_range, b6 = est[flat[k + 1]].dist == p->max_range;
                close = (a6 != b6);
            }
            if (close) sub_off[++nsub] = k + 1;
        }
    }
    /* G: confirmation (ENV:573-620) */
    conf_obj conf[CNF_CONF_CAP]; int nconf = 0;
    for (int s = 0; s < nsub; ++s) {
        int b = sub_off[s], e = sub_off[s + 1], len = e - b;
        int any_hit = 0, n_o = 0, n_w = 0, n_none = 0;
        for (int k = b; k < e; ++k) {
            const item* it = &est[flat[k]];
            if (it->dist != p->max_range) any_hit = 1;
            if (it->type == T_O) ++n_o; else if (it->type == T_W) ++n_w; else ++n_none;
        }
        if (!any_hit) continue;
        if (len < 4) continue;
        const item* ctr = &est[flat[b + len / 2]];
        double estd = 3.0 + floor(29.0 * (p->max_range - ctr->dist) / (p->max_range - p->min_range));   /* UTL:395-402 */
        double denom = ((double)len < estd) ? (double)len : estd;
        double score = (double)n_o / denom;
        int distinct = (n_o > 0) + (n_w > 0) + (n_none > 0);
        int type = -1;
        if (distinct > 1) {
            if (score >= 0.5) type = (n_o > n_w) ? T_O : T_W;
            else if ((double)len <= estd) type = (n_o > n_w) ? T_O : T_W;
            else type = T_W;
        } else {
            double lim = ((double)nsub < estd) ? (double)nsub : estd;
            if (n_w > 0) { if (!((double)len <= lim)) type = T_W; }
            else { if (!((double)len <= lim)) type = T_O; }
        }
        if (type < 0) continue;
        if (nconf < CNF_CONF_CAP) { conf[nconf].type = type; conf[nconf].pose = ctr->pose; conf[nconf].dist = ctr->dist; ++nconf; }
        else trk[CNF_H_OVERFLOW] += 1;
    }
    int n_obst = 0; int ego_hit = 0;
    for (int c = 0; c < nconf; ++c) if (conf[c].type == T_O) { ++n_obst; if (conf[c].dist < 0.140) ego_hit = 1; }
    if (n_obst > 0) trk[CNF_H_PRESENT] += 1;                      /* ENV:653-654 */

    /* H: tracker (ENV:656-743) */
    trk_entry ent[CNF_TRK_CAP + CNF_CONF_CAP]; int alive[CNF_TRK_CAP + CNF_CONF_CAP];
    int n0 = (int)trk[CNF_H_N];
    load_entries(trk, ent, n0);
    int n_ent = n0;
    for (int i = 0; i < n0; ++i) alive[i] = 1;
    if (n0 == 0) {
        for (int c = 0; c < nconf; ++c) {
            if (conf[c].type != T_O) continue;
            trk_entry* t = &ent[n_ent]; alive[n_ent] = 1; ++n_ent;
            t->prev = conf[c].pose; t->last = conf[c].pose; t->dist = conf[c].dist; t->ndeq = 1;
            t->speed = -1.0; t->vx = 0.0; t->vy = 0.0;
        }
    } else if (nconf == 0) {
        for (int i = 0; i < n0; ++i) alive[i] = 0;                 /* ENV:686-689: every tracked object dropped */
    } else {
        char checked[CNF_CONF_CAP]; memset(checked, 0, sizeof(checked));
        int n_live = n0;
        for (int i = 0; i < n0; ++i) if (ent[i].ndeq > 1) { ent[i].prev = ent[i].last; ent[i].ndeq = 1; }   /* popleft */
        /* the IoUs are taken against the tracked poses BEFORE any update (ENV:684,691) */
        double iou[CNF_TRK_CAP][CNF_CONF_CAP];
        for (int i = 0; i < n0; ++i) for (int c = 0; c < nconf; ++c) iou[i][c] = iou_boxes(ent[i].last, conf[c].pose, p->track_half);
        for (int i = 0; i < n0; ++i) {
            int m = 0;
            for (int c = 1; c < nconf; ++c) if (iou[i][c] > iou[i][m]) m = c;
            if (iou[i][m] > 0.0) {
                ent[i].prev = ent[i].last; ent[i].last = conf[m].pose; ent[i].dist = conf[m].dist; ent[i].ndeq = 2;
                checked[m] = 1;
            } else if (n_live > i) {                               /* ENV:718-721 */
                alive[i] = 0; --n_live;
            }
        }
        for (int c = 0; c < nconf; ++c) {
            if (checked[c] || conf[c].type != T_O) continue;
            trk_entry* t = &ent[n_ent]; alive[n_ent] = 1; ++n_ent;
            t->prev = conf[c].pose; t->last = conf[c].pose; t->dist = conf[c].dist; t->ndeq = 1;
            t->speed = -1.0; t->vx = 0.0; t->vy = 0.0;
        }
    }
    /* compact to dict order */
    {
        int m = 0;
        for (int i = 0; i < n_ent; ++i) if (alive[i]) { if (m < CNF_TRK_CAP) ent[m++] = ent[i]; else trk[CNF_H_OVERFLOW] += 1; }
        n_ent = m;
    }
    /* I: speed (ENV:745-760) */
    for (int i = 0; i < n_ent; ++i)
        if (ent[i].ndeq > 1) ent[i].speed = cn_hypot64(ent[i].prev.y - ent[i].last.y, ent[i].prev.x - ent[i].last.x) / p->dt;

    /* J / K (ENV:765-907) */
    for (int s = 0; s < K; ++s) { kblock[4 * s] = x; kblock[4 * s + 1] = y; kblock[4 * s + 2] = 0.0; kblock[4 * s + 3] = 0.0; }
    double ego_score; memcpy(&ego_score, trk + CNF_H_EGOSCORE, 8);
    if (trk[CNF_H_HAVE_PREV]) {
        pt prev; prev.x = cn_milli64((int32_t)trk[CNF_H_PPX]); prev.y = cn_milli64((int32_t)trk[CNF_H_PPY]);
        double vx = (cur.x - prev.x) / p->dt, vy = (cur.y - prev.y) / p->dt;      /* UTL:227-236 */
        double agent_vel = sqrt(vx * vx + vy * vy);
        double obstacle_vel = (n_ent == 0) ? 0.0 : ent[0].speed;     /* ENV:789-797 */
        pt vo = cur;
        for (int i = 0; i < n_ent; ++i) {
            double cx = 0.0, cy = 0.0;
            if (ent[i].ndeq > 1) {
                cx = ent[i].prev.x - ent[i].last.x; cy = ent[i].prev.y - ent[i].last.y;   /* last - curr (sic) */
                ent[i].vx = cx / p->dt; ent[i].vy = cy / p->dt;
            }
            vo.x = cur.x + cx; vo.y = cur.y + cy;                  /* leaks out of the loop (ENV:814-815) */
        }
        double cp[CNF_TRK_CAP], ego[CNF_TRK_CAP];
        double ego_cur = 0.0;
        for (int i = 0; i < n_ent; ++i) {
            double dtc; int have = collision_point(prev, vo, ent[i].last, p->cp_radius, &dtc);
            double resultant = agent_vel - obstacle_vel;
            double dto = (ent[i].dist > p->max_range) ? 0.0 : (p->max_range - ent[i].dist) / (p->max_range - p->min_range);
            if (have) {
                if (resultant == 0.0) cp[i] = 1.0 * dto;
                else {
                    double ttc = dtc / resultant;
                    double q = 0.15 / ttc;
                    ego_cur = (1.0 < q) ? 1.0 : q;                 /* min(1, 0.15 / ttc), UTL:319 */
                    cp[i] = 0.5 * ego_cur + 0.5 * dto;
                }
            } else {
                ego_cur = 0.0;
                cp[i] = 0.5 * 0.0 + 0.5 * dto;
            }
            ego[i] = ego_cur;
        }
        if (n_ent == 0) ego_score = 0.0;
        else {
            ego_score = ego[0];
            for (int i = 1; i < n_ent; ++i) if (ego[i] > ego_score) ego_score = ego[i];
            /* sorted(..., reverse=True) is stable; keep [-K:] */
            for (int a = 0; a < (K > 0 ? n_ent : 0); ++a) {
                int rank = 0;
                for (int b = 0; b < n_ent; ++b) if (b != a && (cp[b] > cp[a] || (cp[b] == cp[a] && b < a))) ++rank;
                int slot = p->topk_highest ? rank : rank - (n_ent > K ? n_ent - K : 0);
                if (slot < 0 || slot >= K) continue;
                kblock[4 * slot] = ent[a].last.x; kblock[4 * slot + 1] = ent[a].last.y;
                kblock[4 * slot + 2] = ent[a].vx; kblock[4 * slot + 3] = ent[a].vy;
            }
        }
    }
    memcpy(trk + CNF_H_EGOSCORE, &ego_score, 8);
    /* FIFO (ENV:991-992); at step 0 the deque just received its first pose (ENV:294) */
    trk[CNF_H_HAVE_PREV] = 1; trk[CNF_H_PPX] = (uint32_t)milli_of(cur.x); trk[CNF_H_PPY] = (uint32_t)milli_of(cur.y);

    /* M (ENV:998-1005) */
    if (ego_hit) trk[CNF_H_EGO] += 1;
    if (ego_score > 0.4) trk[CNF_H_SOCIAL] += 1;
    if (step_counter == 0) { trk[CNF_H_EGO] = 0; trk[CNF_H_SOCIAL] = 0; trk[CNF_H_PRESENT] = 0; }   /* ENV:1258-1260 */

    store_entries(trk, ent, n_ent);
    for (int k = 0; k < 4 * K; ++k) kblock[k] = cn_np_round3_64(kblock[k]);    /* ENV:1042 */

    free(pose); free(fr); free(grad); free(grad_ok); free(chg); free(chg_ok); free(rec_type); free(rec_src);
    free(est); free(flat); free(seg_off); free(sub_off);
}

/* test taps */
void orf_sincos64(const double* a, double* s, double* c, int n) { for (int i = 0; i < n; ++i) cn_sincos64(a[i], s + i, c + i); }
void orf_round3(const double* x, double* out, int n) { for (int i = 0; i < n; ++i) out[i] = cn_py_round3_64(x[i]); }
int orf_world_words(void) { return CNF_WORLD_WORDS; }
long orf_milli_mismatches(long lo, long hi) {
    long bad = 0;
    for (long k = lo; k <= hi; ++k) bad += (cn_milli64(k) != (double)k / 1000.0);
    return bad;
}
