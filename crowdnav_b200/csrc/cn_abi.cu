/*
 * cn_abi.cu -- the extern "C" boundary declared in include/crowdnav.h.
 * Host-side only: argument checking, one device arena per handle, kernel
 * parameter packing.  All device work goes to the caller's stream.
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <new>
#include "cn_kernel.h"

struct cn_handle {
    cn_config cfg;
    cn_derived d;
    int device;
    void* arena;            /* one cudaMalloc: [cn_config][robot][ped_a][ped_b] */
    size_t arena_bytes;
    cn_config* cfg_dev;
    uint32_t* robot;
    uint32_t* ped_a;
    uint32_t* ped_b;
    uint32_t* trk;          /* CN_FLAG_RISK_FAITHFUL: tracker plane [E][CNF_WORLD_WORDS], else NULL */
    float* own_ranges;      /* CN_FLAG_RISK_FAITHFUL: pre-rounding ranges [E][R-1] the step kernel leaves for cn_faithful_kernel */
    float* dbg_ranges;
    uint8_t* dbg_hid;
    unsigned int* gather_timeouts;  /* device word: bounded waits of the fused gather that gave up */
    int64_t launches;
    int use_flat;           /* 1: cn_flat.cu (compacted work lists, default), 0: cn_step.cu (warp per world; CN_KERNEL=warp) */
    cn_flat_layout flat;    /* rows staged in shared memory: reset launches, every fused-gather entry point, host-mapped rows */
    cn_flat_layout flat_direct;   /* "direct rows" layout of plain steps into device memory (cn_flat.cu, DIRECT = 1) */
    int have_direct;
    const void* obs_seen[4];   /* the last few obs pointers classified by cudaPointerGetAttributes (rollouts alternate buffers) ... */
    int obs_seen_device[4];    /* ... 1: device (or managed) memory, 0: host-mapped or unknown */
    int obs_seen_next;
    cn_kparams base;        /* everything of cn_kparams that does not change between calls, packed once (repack()) */
};

/* Every entry point runs with the handle's device current and puts the caller's device back on return:
 * one process may hold handles on several GPUs. */
struct cn_device_guard {
    int prev, dev;
    explicit cn_device_guard(int d) : prev(-1), dev(d) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~cn_device_guard() { if (prev >= 0 && prev != dev) cudaSetDevice(prev); }
};

/* cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember the largest value
 * configured for (kernel slot, device) -- not one process-wide static per kernel. */
#include <mutex>
static std::mutex g_attr_mu;
static size_t g_attr[CN_ATTR_SLOTS][CN_ATTR_MAX_DEVICES];
cudaError_t cn_ensure_smem_attr(const void* func, int slot, size_t smem) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (slot < 0 || slot >= CN_ATTR_SLOTS || dev < 0 || dev >= CN_ATTR_MAX_DEVICES) return cudaErrorInvalidValue;
    std::lock_guard<std::mutex> lock(g_attr_mu);
    if (smem <= g_attr[slot][dev] && g_attr[slot][dev] != 0) return cudaSuccess;
    e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess) g_attr[slot][dev] = smem;
    return e;
}

static cudaError_t launch_env(const cn_handle* h, cn_kparams& P, int mode, cudaStream_t s, bool direct = false) {
    if (h->use_flat) return cn_launch_flat_kernel(P, direct ? h->flat_direct : h->flat, mode, s);
    P.obs_bulk_ok = P.obs_bulk_ok && ((size_t)CN_TILE * h->d.obs_dim) % 4 == 0;   /* every tile starts 16-B aligned */
    return cn_launch_env_kernel(P, mode, s);
}

/* the risk_faithful block runs behind the step / reset kernel on the same stream */
static cudaError_t launch_faithful(cn_handle* h, float* obs, const uint8_t* mask, cudaStream_t s) {
    if (!h->trk) return cudaSuccess;
    cudaError_t e = cn_launch_faithful(&h->cfg, h->robot, h->trk, h->dbg_ranges ? h->dbg_ranges : h->own_ranges,
                                       obs, mask, h->d.obs_dim, s);
    if (e == cudaSuccess) h->launches += 1;
    return e;
}

static void repack(cn_handle* h);

static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail ? detail : "");
    return code;
}
#define CN_CUDA(call)                                                        \
    do {                                                                     \
        cudaError_t e_ = (call);                                             \
        if (e_ != cudaSuccess) return fail(CN_ERR_CUDA, #call ": %s", cudaGetErrorString(e_)); \
    } while (0)

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
/* staging tile of the pipelined gather: 0 none, 1 fp32 rows, 2 int16 rows (tile sized for 16-byte units of int16) */
static int stage_mode(const cn_config* cfg) {
    if (!(cfg->flags & CN_FLAG_GATHER_STAGE)) return 0;
    return (cfg->flags & CN_FLAG_GATHER_WIRE16) ? 2 : 1;
}

extern "C" {

const char* cn_last_error(void) { return g_err; }
int cn_abi_version(void) { return CN_ABI_VERSION; }

int cn_config_default(cn_config* c) {
    if (!c) return fail(CN_ERR_INVALID, "cn_config_default: null config%s", NULL);
    memset(c, 0, sizeof(*c));
    c->struct_size = sizeof(cn_config);
    c->n_envs = 1; c->n_peds = 14; c->n_samples = 360; c->k_obstacles = 8; c->max_steps = 1000;
    c->seed = 1234;
    c->dt = 0.15f;
    c->room_xmin = -1.411074f; c->room_xmax = 1.398926f; c->room_ymin = -1.394624f; c->room_ymax = 1.405376f;
    c->start_x = 1.0f; c->start_y = -1.0f; c->start_yaw = 3.14f;
    c->goal_x = -1.0f; c->goal_y = 1.0f;
    c->heading_off_x = 0.75f; c->heading_off_y = -0.75f;
    c->max_range = 0.6f; c->collision_range = 0.12f; c->sensor_min_range = 0.08f;
    c->sensor_sweep = 6.28f; c->mount_x = -0.032f; c->hit_angle_inc_deg = 1.0f;
    c->ped_radius = 0.0505f; c->robot_radius = 0.105f; c->cp_radius = 0.178f;
    c->waypoint_radius = 0.3f; c->goal_box = 0.20f;
    c->rep_strength = 0.5f; c->rep_range = 0.05f; c->rep_cutoff = 0.05f; c->layout_jitter = 0.0f;
    c->wheel_accel = 0.0f; c->n_substeps = 1;
    c->n_behaviors = 1;
    c->behavior_kind[0] = CN_BEHAVIOR_RANDOM; c->behavior_speed[0] = 0.2f;
    c->behavior_period_ticks[0] = 30; c->behavior_stagger_ticks[0] = 2;
    static const float first6[6][2] = {{-0.01f, -1.0f}, {-1.15f, -0.3f}, {-0.32f, -0.12f},
                                       {-0.85f, 0.92f}, {0.94f, 0.99f}, {0.65f, 0.2f}};
    static const float ring[8][2] = {{1.0f, 0.0f}, {0.70710678f, 0.70710678f}, {0.0f, 1.0f}, {-0.70710678f, 0.70710678f},
                                     {-1.0f, 0.0f}, {-0.70710678f, -0.70710678f}, {0.0f, -1.0f}, {0.70710678f, -0.70710678f}};
    for (int i = 0; i < 6; ++i) { c->ped_layout[i][0] = first6[i][0]; c->ped_layout[i][1] = first6[i][1]; }
    for (int i = 0; i < 8; ++i) { c->ped_layout[6 + i][0] = 0.22f + 0.16f * ring[i][0]; c->ped_layout[6 + i][1] = 0.54f + 0.16f * ring[i][1]; }
    return CN_OK;
}

int cn_obs_dim(const cn_config* c) {
    if (!c) return fail(CN_ERR_INVALID, "cn_obs_dim: null config%s", NULL);
    if (c->flags & CN_FLAG_ENV_ORIGINAL) return (c->n_samples - 1) + 4;
    return (c->n_samples - 1) + 7 + 4 * c->k_obstacles;
}

size_t cn_blob_bytes(const cn_config* c) {
    if (!c) return 0;
    return cn_blob_words(c) * 4;
}

int cn_create(const cn_config* cfg, int device, cn_handle** out) {
    if (!cfg || !out) return fail(CN_ERR_INVALID, "cn_create: null argument%s", NULL);
    *out = NULL;
    if (cfg->struct_size != sizeof(cn_config))
        return fail(CN_ERR_INVALID, "cn_create: cn_config.struct_size mismatch (ABI)%s", NULL);
    cn_derived d;
    if (cn_derive(cfg, &d) != 0) return fail(CN_ERR_INVALID, "cn_create: config out of range%s", NULL);
    int ndev = 0;
    CN_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(CN_ERR_INVALID, "cn_create: no such device%s", NULL);
    cn_device_guard guard(device);           /* the caller's current device is put back on return */
    int max_smem = 0;
    CN_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    const char* kern = getenv("CN_KERNEL");
    const int use_flat = !(kern && strcmp(kern, "warp") == 0);
    cn_flat_layout flat; memset(&flat, 0, sizeof(flat));
    cn_flat_layout flat_direct; memset(&flat_direct, 0, sizeof(flat_direct));
    int have_direct = 0;
    if (!use_flat && (cfg->flags & CN_FLAG_ENV_ORIGINAL))
        return fail(CN_ERR_UNSUPPORTED, "cn_create: CN_FLAG_ENV_ORIGINAL needs the default kernel (unset CN_KERNEL)%s", NULL);
    if (use_flat) {
        /* CN_FLAT_TILE=W[,threads] overrides the automatic choice (experiments) */
        const char* tile = getenv("CN_FLAT_TILE");
        int max_sm = 0;
        CN_CUDA(cudaDeviceGetAttribute(&max_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device));
        int rc;
        if (tile) {
            int tw = atoi(tile), tt = 256;
            const char* comma = strchr(tile, ',');
            if (comma) tt = atoi(comma + 1);
            rc = cn_flat_make_layout(cfg->n_peds, cfg->n_samples, d.obs_dim, tw, tt, stage_mode(cfg), 0, &flat);
        } else {
            int n_sms = 0;
            CN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
            rc = cn_flat_pick_tile(cfg->n_peds, cfg->n_samples, d.obs_dim, cfg->n_envs, n_sms, (size_t)max_sm, stage_mode(cfg), 0, &flat);
        }
        { const char* st = getenv("CN_FLAT_STORE"); flat.plain_store = (st && strcmp(st, "plain") == 0) ? 1 : 0; }
        /* programmatic dependent launch is opt-in (CN_PDL=1): measured on B200 it helps back-to-back stream launches
         * (c2 23.2 -> 20.6 us per step) but not the graph-replayed step (11.8 -> 12.8 us), see DESIGN.md 4.4 */
        { const char* pd = getenv("CN_PDL"); flat.pdl = (pd && strcmp(pd, "1") == 0) ? 1 : 0; }
        { const char* gd = getenv("CN_GATHER_DEBUG"); flat.gather_debug = gd ? atoi(gd) : 0; }   /* timing diagnostics only */
        if (rc != 0 || flat.total > (size_t)max_smem)
            return fail(CN_ERR_UNSUPPORTED, "cn_create: tile does not fit shared memory (reduce n_samples / n_peds)%s", NULL);
        /* the direct-rows layout of plain steps (CN_FLAT_DIRECT=0 turns it off; CN_FLAT_TILE applies to it too) */
        const char* dr = getenv("CN_FLAT_DIRECT");
        if (!(dr && strcmp(dr, "0") == 0)) {
            int drc;
            if (tile) {
                int tw = atoi(tile), tt = 256;
                const char* comma = strchr(tile, ',');
                if (comma) tt = atoi(comma + 1);
                drc = cn_flat_make_layout(cfg->n_peds, cfg->n_samples, d.obs_dim, tw, tt, 0, 1, &flat_direct);
                if (drc == 0) {     /* what the CTA's share of shared memory leaves goes to the fill tile (as cn_flat_pick_tile does) */
                    const int ctas = tt >= 512 ? 3 : tt >= 384 ? 4 : tt >= 256 ? 6 : tt >= 192 ? 8 : 12;
                    const size_t budget = (size_t)max_sm / ctas - 1024, want = (((size_t)cfg->n_samples - 1) * 4 + 15) & ~(size_t)15;
                    size_t fc = 512 + (budget > flat_direct.total ? budget - flat_direct.total : 0);
                    if (fc > want) fc = want;
                    fc &= ~(size_t)15;
                    if (fc > 512) drc = cn_flat_make_layout(cfg->n_peds, cfg->n_samples, d.obs_dim, tw, tt, 0, (int)fc, &flat_direct);
                }
            } else {
                int n_sms = 0;
                CN_CUDA(cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, device));
                drc = cn_flat_pick_tile(cfg->n_peds, cfg->n_samples, d.obs_dim, cfg->n_envs, n_sms, (size_t)max_sm, 0, 1, &flat_direct);
            }
            have_direct = (drc == 0 && flat_direct.total <= (size_t)max_smem) ? 1 : 0;
            flat_direct.plain_store = flat.plain_store; flat_direct.pdl = flat.pdl; flat_direct.gather_debug = 0;
        }
    } else {
        const size_t smem = cn_kernel_smem_bytes(cfg->n_peds, cfg->n_samples, d.obs_dim);
        if (smem > (size_t)max_smem)
            return fail(CN_ERR_UNSUPPORTED, "cn_create: tile does not fit shared memory (reduce n_samples / n_peds)%s", NULL);
    }

    if ((cfg->flags & CN_FLAG_RISK_FAITHFUL) && cn_faithful_smem_bytes(cfg->n_samples - 1) > (size_t)max_smem)
        return fail(CN_ERR_UNSUPPORTED, "cn_create: risk_faithful scratch does not fit shared memory (reduce n_samples)%s", NULL);
    cn_handle* h = new (std::nothrow) cn_handle();
    if (!h) return fail(CN_ERR_NOMEM, "cn_create: host allocation failed%s", NULL);
    h->cfg = *cfg; h->d = d; h->device = device; h->launches = 0;
    h->use_flat = use_flat; h->flat = flat;
    h->flat_direct = flat_direct; h->have_direct = have_direct; h->obs_seen_next = 0;
    for (int i = 0; i < 4; ++i) { h->obs_seen[i] = NULL; h->obs_seen_device[i] = 0; }
    h->dbg_ranges = NULL; h->dbg_hid = NULL;
    const size_t cfg_b = align_up(sizeof(cn_config), 256);
    const size_t rob_b = align_up(cn_robot_words(cfg) * 4, 256);
    const size_t ped_b = align_up(cn_ped_plane_words(cfg) * 4 + 16, 256);
    const bool faithful = (cfg->flags & CN_FLAG_RISK_FAITHFUL) != 0;
    const size_t trk_b = faithful ? align_up(cn_trk_words(cfg) * 4, 256) : 0;
    const size_t rng_b = faithful ? align_up((size_t)cfg->n_envs * (size_t)(cfg->n_samples - 1) * 4, 256) : 0;
    h->arena_bytes = cfg_b + rob_b + 2 * ped_b + trk_b + rng_b + 256;
    cudaError_t e = cudaMalloc(&h->arena, h->arena_bytes);
    if (e != cudaSuccess) { delete h; return fail(CN_ERR_NOMEM, "cn_create: cudaMalloc: %s", cudaGetErrorString(e)); }
    uint8_t* p = (uint8_t*)h->arena;
    h->cfg_dev = (cn_config*)p; p += cfg_b;
    h->robot = (uint32_t*)p; p += rob_b;
    h->ped_a = (uint32_t*)p; p += ped_b;
    h->ped_b = (uint32_t*)p; p += ped_b;
    h->trk = faithful ? (uint32_t*)p : NULL; p += trk_b;
    h->own_ranges = faithful ? (float*)p : NULL; p += rng_b;
    h->gather_timeouts = (unsigned int*)p;
    e = cudaMemset(h->arena, 0, h->arena_bytes);
    if (e == cudaSuccess) e = cudaMemcpy(h->cfg_dev, cfg, sizeof(cn_config), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(h->arena); delete h; return fail(CN_ERR_CUDA, "cn_create: init: %s", cudaGetErrorString(e)); }
    repack(h);
    *out = h;
    return CN_OK;
}

int cn_destroy(cn_handle* h) {
    if (!h) return CN_OK;
    cn_device_guard guard(h->device);
    cudaFree(h->arena);
    delete h;
    return CN_OK;
}

}  /* extern "C" */

/* the call-invariant part of cn_kparams, built once per handle (and again when the debug taps change) */
static void repack(cn_handle* h) {
    const cn_config* c = &h->cfg;
    cn_kparams* P = &h->base;
    memset(P, 0, sizeof(*P));
    P->robot = h->robot; P->ped_a = h->ped_a; P->ped_b = h->ped_b;
    P->dbg_ranges = h->dbg_ranges ? h->dbg_ranges : h->own_ranges; P->dbg_hid = h->dbg_hid;
    P->cfg = h->cfg_dev; P->d = h->d;
    P->gather_timeouts = h->gather_timeouts; P->gather_done = h->gather_timeouts + 1;
    P->n_envs = c->n_envs; P->n_peds = c->n_peds; P->n_samples = c->n_samples; P->k_obstacles = c->k_obstacles;
    P->max_steps = c->max_steps; P->env_id_offset = c->env_id_offset; P->n_behaviors = c->n_behaviors;
    P->n_substeps = c->n_substeps;
    P->flags = c->flags; P->dt = c->dt;
    P->room_xmin = c->room_xmin; P->room_xmax = c->room_xmax; P->room_ymin = c->room_ymin; P->room_ymax = c->room_ymax;
    P->goal_x = c->goal_x; P->goal_y = c->goal_y; P->heading_off_x = c->heading_off_x; P->heading_off_y = c->heading_off_y;
    P->max_range = c->max_range; P->collision_range = c->collision_range; P->sensor_min_range = c->sensor_min_range;
    P->mount_x = c->mount_x; P->ped_radius = c->ped_radius; P->robot_radius = c->robot_radius; P->goal_box = c->goal_box;
    P->rep_strength = c->rep_strength; P->rep_range = c->rep_range; P->rep_cutoff = c->rep_cutoff;
    P->layout_jitter = c->layout_jitter;
    for (int b = 0; b < CN_MAX_BEHAVIORS; ++b) {
        P->beh_kind[b] = c->behavior_kind[b]; P->beh_speed[b] = c->behavior_speed[b];
        P->beh_period[b] = c->behavior_period_ticks[b]; P->beh_stagger[b] = c->behavior_stagger_ticks[b];
    }
}

extern "C" {

int cn_reset(cn_handle* h, const uint8_t* mask_dev, float* obs_dev, void* stream) {
    if (!h || !obs_dev) return fail(CN_ERR_INVALID, "cn_reset: null argument%s", NULL);
    cn_device_guard guard(h->device);
    cn_kparams P = h->base;
    P.mask = mask_dev; P.obs = obs_dev; P.obs_bulk_ok = 0;
    CN_CUDA(launch_env(h, P, 1, (cudaStream_t)stream));
    h->launches += 1;
    CN_CUDA(launch_faithful(h, obs_dev, mask_dev, (cudaStream_t)stream));
    return CN_OK;
}

/* one step launch (+ the risk_faithful kernel) with the handle's device already current */
/* Direct rows only into device memory: into a host-mapped buffer (CrowdNavVecEnv.step_host(mode="mapped")) the fill and
 * the scattered stores would cross PCIe one by one, where the staged tile leaves as one bulk store.  The pointer is
 * classified once (no stream operation: fine under graph capture) and remembered. */
static bool rows_in_device_memory(cn_handle* h, const void* obs_dev) {
    for (int i = 0; i < 4; ++i)
        if (h->obs_seen[i] == obs_dev) return h->obs_seen_device[i] != 0;
    cudaPointerAttributes a;
    const cudaError_t e = cudaPointerGetAttributes(&a, obs_dev);
    if (e != cudaSuccess) cudaGetLastError();
    const int k = h->obs_seen_next;
    h->obs_seen_next = (k + 1) & 3;
    h->obs_seen[k] = obs_dev;
    h->obs_seen_device[k] = (e == cudaSuccess && (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged)) ? 1 : 0;
    return h->obs_seen_device[k] != 0;
}

static int step_once(cn_handle* h, const float* action_dev, float* obs_dev, float* const* peers, int n_peers,
                     float* reward_dev, uint8_t* done_dev, cudaStream_t stream) {
    cn_kparams P = h->base;
    const bool direct = h->use_flat && h->have_direct && n_peers == 0 && rows_in_device_memory(h, obs_dev);
    P.action = action_dev; P.obs = obs_dev; P.reward = reward_dev; P.done = done_dev;
    bool aligned = (((uintptr_t)obs_dev) & 15u) == 0;
    for (int p = 0; p < n_peers; ++p) {
        P.obs_peers[p] = peers[p];
        aligned = aligned && (((uintptr_t)peers[p]) & 15u) == 0;
    }
    P.n_obs_peers = n_peers;
    P.obs_bulk_ok = aligned;
    P.act_bulk_ok = (((uintptr_t)action_dev) & 15u) == 0;
    CN_CUDA(launch_env(h, P, 0, stream, direct));
    h->launches += 1;
    CN_CUDA(launch_faithful(h, obs_dev, NULL, stream));
    return CN_OK;
}

int cn_step(cn_handle* h, const float* action_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev, void* stream) {
    if (!h || !action_dev || !obs_dev || !reward_dev || !done_dev)
        return fail(CN_ERR_INVALID, "cn_step: null argument%s", NULL);
    cn_device_guard guard(h->device);
    return step_once(h, action_dev, obs_dev, NULL, 0, reward_dev, done_dev, (cudaStream_t)stream);
}

int cn_step_n(cn_handle* h, int n_steps, const float* action_dev, size_t action_stride, float* obs_dev,
              float* reward_dev, uint8_t* done_dev, size_t out_stride, void* stream) {
    if (!h || !action_dev || !obs_dev || !reward_dev || !done_dev)
        return fail(CN_ERR_INVALID, "cn_step_n: null argument%s", NULL);
    if (n_steps < 0) return fail(CN_ERR_INVALID, "cn_step_n: n_steps < 0%s", NULL);
    cn_device_guard guard(h->device);
    for (int i = 0; i < n_steps; ++i) {
        int rc = step_once(h, action_dev + (size_t)i * action_stride, obs_dev, NULL, 0,
                           reward_dev + (size_t)i * out_stride, done_dev + (size_t)i * out_stride, (cudaStream_t)stream);
        if (rc != CN_OK) return rc;
    }
    return CN_OK;
}

int cn_step_gather(cn_handle* h, const float* action_dev, float* obs_dev, float* const* peer_obs_dev, int n_peers,
                   float* reward_dev, uint8_t* done_dev, void* stream) {
    if (!h || !action_dev || !obs_dev || !reward_dev || !done_dev)
        return fail(CN_ERR_INVALID, "cn_step_gather: null argument%s", NULL);
    if (n_peers < 0 || n_peers > 8 || (n_peers > 0 && !peer_obs_dev))
        return fail(CN_ERR_INVALID, "cn_step_gather: 0..8 peer buffers%s", NULL);
    if (h->trk && n_peers > 0)
        return fail(CN_ERR_UNSUPPORTED, "cn_step_gather: CN_FLAG_RISK_FAITHFUL rewrites the K block after the step kernel; "
                                        "gather with ncclAllGather instead%s", NULL);
    for (int p = 0; p < n_peers; ++p)
        if (!peer_obs_dev[p]) return fail(CN_ERR_INVALID, "cn_step_gather: null peer buffer%s", NULL);
    cn_device_guard guard(h->device);
    return step_once(h, action_dev, obs_dev, peer_obs_dev, n_peers, reward_dev, done_dev, (cudaStream_t)stream);
}

int cn_step_gather_signal(cn_handle* h, const float* action_dev, float* obs_dev, float* const* peer_obs_dev,
                          unsigned long long* const* peer_arrive_dev, int n_peers, float* obs_mc_dev,
                          unsigned long long* arrive_mc_dev, unsigned long long* arrive_local_dev, int n_ranks,
                          int rank, int wait_back, float* reward_dev, uint8_t* done_dev, void* stream) {
    if (!h || !action_dev || !obs_dev || !reward_dev || !done_dev)
        return fail(CN_ERR_INVALID, "cn_step_gather_signal: null argument%s", NULL);
    if (!h->use_flat) return fail(CN_ERR_UNSUPPORTED, "cn_step_gather_signal: default (flat) kernel only%s", NULL);
    if (h->trk) return fail(CN_ERR_UNSUPPORTED, "cn_step_gather_signal: CN_FLAG_RISK_FAITHFUL rewrites the K block after "
                                                "the step kernel; gather with ncclAllGather instead%s", NULL);
    const bool mc = obs_mc_dev != NULL;
    if (mc != (arrive_mc_dev != NULL)) return fail(CN_ERR_INVALID, "cn_step_gather_signal: obs_mc_dev and arrive_mc_dev go together%s", NULL);
    if (!mc && (n_peers < 1 || n_peers > 8 || !peer_obs_dev || !peer_arrive_dev))
        return fail(CN_ERR_INVALID, "cn_step_gather_signal: 1..8 peers (or multicast addresses)%s", NULL);
    if (!arrive_local_dev || wait_back < 0) return fail(CN_ERR_INVALID, "cn_step_gather_signal: arrive_local_dev / wait_back%s", NULL);
    if (n_ranks < 2 || n_ranks > 32 || rank < 0 || rank >= n_ranks) return fail(CN_ERR_INVALID, "cn_step_gather_signal: 2..32 ranks, 0 <= rank < n_ranks%s", NULL);
    if (mc && (((uintptr_t)obs_mc_dev) & 15u)) return fail(CN_ERR_INVALID, "cn_step_gather_signal: obs_mc_dev must be 16-byte aligned%s", NULL);
    cn_device_guard guard(h->device);
    cn_kparams P = h->base;
    P.action = action_dev; P.obs = obs_dev; P.reward = reward_dev; P.done = done_dev;
    bool aligned = (((uintptr_t)obs_dev) & 15u) == 0;
    if (mc) {
        P.obs_mc = obs_mc_dev; P.arrive_mc = arrive_mc_dev;
    } else {
        for (int p = 0; p < n_peers; ++p) {
            if (!peer_obs_dev[p] || !peer_arrive_dev[p]) return fail(CN_ERR_INVALID, "cn_step_gather_signal: null peer pointer%s", NULL);
            P.obs_peers[p] = peer_obs_dev[p]; P.arrive_peers[p] = peer_arrive_dev[p];
            aligned = aligned && (((uintptr_t)peer_obs_dev[p]) & 15u) == 0;
        }
        P.n_obs_peers = n_peers;
    }
    P.arrive_local = arrive_local_dev; P.arrive_back = wait_back; P.arrive_slots = n_ranks; P.arrive_self = rank;
    P.ctas_per_step = (unsigned int)cn_kernel_ctas(h);
    P.obs_bulk_ok = aligned;
    P.act_bulk_ok = (((uintptr_t)action_dev) & 15u) == 0;
    CN_CUDA(launch_env(h, P, 0, (cudaStream_t)stream));
    h->launches += 1;
    return CN_OK;
}

/* common argument checks + parameter block of the two pipelined-gather entry points */
static int pack_push(cn_handle* h, const char* who, cn_kparams* P, int wire16, const void* push_src_dev, void* const* push_peer_dev,
                     unsigned long long* const* peer_arrive_dev, int n_peers, unsigned long long* arrive_local_dev,
                     int n_ranks, int rank, int wait_back) {
    if (!h->use_flat) return fail(CN_ERR_UNSUPPORTED, "%s: default (flat) kernel only", who);
    if (h->trk) return fail(CN_ERR_UNSUPPORTED, "%s: not with CN_FLAG_RISK_FAITHFUL (gather with ncclAllGather instead)", who);
    if (n_peers < 0 || n_peers > 8 || wait_back < 0) return fail(CN_ERR_INVALID, "%s: 0..8 peers", who);
    *P = h->base;
    if (n_peers == 0) return CN_OK;                                     /* nothing to forward (first step after a reset) */
    if (!push_src_dev || !push_peer_dev || !peer_arrive_dev || !arrive_local_dev) return fail(CN_ERR_INVALID, "%s: null argument", who);
    if (n_ranks < 2 || n_ranks > 32 || rank < 0 || rank >= n_ranks) return fail(CN_ERR_INVALID, "%s: 2..32 ranks, 0 <= rank < n_ranks", who);
    bool aligned = (((uintptr_t)push_src_dev) & 15u) == 0;
    for (int p = 0; p < n_peers; ++p) {
        if (!push_peer_dev[p] || !peer_arrive_dev[p]) return fail(CN_ERR_INVALID, "%s: null peer pointer", who);
        P->push_peers[p] = push_peer_dev[p]; P->arrive_peers[p] = peer_arrive_dev[p];
        aligned = aligned && (((uintptr_t)push_peer_dev[p]) & 15u) == 0;
    }
    P->push_src = push_src_dev; P->n_push_peers = n_peers; P->push_bulk_ok = aligned ? 1 : 0; P->push_wire16 = wire16 ? 1 : 0;
    P->arrive_local = arrive_local_dev; P->arrive_back = wait_back; P->arrive_slots = n_ranks; P->arrive_self = rank;
    P->ctas_per_step = (unsigned int)cn_kernel_ctas(h);
    return CN_OK;
}

int cn_step_gather_async(cn_handle* h, const float* action_dev, float* obs_dev, int16_t* wire_out_dev, const void* push_src_dev,
                         void* const* push_peer_dev, unsigned long long* const* peer_arrive_dev, int n_peers,
                         unsigned long long* arrive_local_dev, int n_ranks, int rank, int wait_back,
                         const int16_t* dec_wire_dev, float* dec_obs_dev,
                         float* reward_dev, uint8_t* done_dev, void* stream) {
    if (!h || !action_dev || !obs_dev || !reward_dev || !done_dev)
        return fail(CN_ERR_INVALID, "cn_step_gather_async: null argument%s", NULL);
    if (!(h->cfg.flags & CN_FLAG_GATHER_STAGE))
        return fail(CN_ERR_INVALID, "cn_step_gather_async: create the handle with CN_FLAG_GATHER_STAGE%s", NULL);
    if (wire_out_dev && (((uintptr_t)wire_out_dev) & 3u)) return fail(CN_ERR_INVALID, "cn_step_gather_async: wire_out_dev must be 4-byte aligned%s", NULL);
    cn_kparams P;
    int rc = pack_push(h, "cn_step_gather_async", &P, wire_out_dev != NULL, push_src_dev, push_peer_dev, peer_arrive_dev, n_peers,
                       arrive_local_dev, n_ranks, rank, wait_back);
    if (rc != CN_OK) return rc;
    cn_device_guard guard(h->device);
    P.action = action_dev; P.obs = obs_dev; P.reward = reward_dev; P.done = done_dev;
    P.wire_out = wire_out_dev;
    if ((dec_wire_dev != NULL) != (dec_obs_dev != NULL)) return fail(CN_ERR_INVALID, "cn_step_gather_async: dec_wire_dev and dec_obs_dev go together%s", NULL);
    if (dec_wire_dev && n_peers < 1)
        return fail(CN_ERR_INVALID, "cn_step_gather_async: decoding inside the kernel needs a pushing launch (its arrival "
                                    "counters are what certify the delivery)%s", NULL);
    P.dec_wire = dec_wire_dev; P.dec_obs = dec_obs_dev;
    P.obs_bulk_ok = (((uintptr_t)obs_dev) & 15u) == 0;
    P.act_bulk_ok = (((uintptr_t)action_dev) & 15u) == 0;
    CN_CUDA(launch_env(h, P, 0, (cudaStream_t)stream));
    h->launches += 1;
    return CN_OK;
}

int cn_gather_flush(cn_handle* h, int wire16, const void* push_src_dev, void* const* push_peer_dev,
                    unsigned long long* const* peer_arrive_dev, int n_peers, unsigned long long* arrive_local_dev,
                    int n_ranks, int rank, int wait_back, void* stream) {
    if (!h) return fail(CN_ERR_INVALID, "cn_gather_flush: null handle%s", NULL);
    if (n_peers < 1) return fail(CN_ERR_INVALID, "cn_gather_flush: 1..8 peers%s", NULL);
    cn_kparams P;
    int rc = pack_push(h, "cn_gather_flush", &P, wire16, push_src_dev, push_peer_dev, peer_arrive_dev, n_peers, arrive_local_dev,
                       n_ranks, rank, wait_back);
    if (rc != CN_OK) return rc;
    cn_device_guard guard(h->device);
    CN_CUDA(cn_launch_push_kernel(P, h->flat, (cudaStream_t)stream));
    h->launches += 1;
    return CN_OK;
}

int cn_gather_decode16(cn_handle* h, const int16_t* wire_dev, float* obs_all_dev, long long row_lo, long long row_hi,
                       long long rows_total, void* stream) {
    if (!h || !wire_dev || !obs_all_dev) return fail(CN_ERR_INVALID, "cn_gather_decode16: null argument%s", NULL);
    if (row_lo < 0 || row_hi < row_lo || rows_total < row_hi) return fail(CN_ERR_INVALID, "cn_gather_decode16: 0 <= row_lo <= row_hi <= rows_total%s", NULL);
    cn_device_guard guard(h->device);
    CN_CUDA(cn_launch_wire_decode(wire_dev, obs_all_dev, row_lo, row_hi, rows_total, h->d.obs_dim, (cudaStream_t)stream));
    h->launches += 1;
    return CN_OK;
}

int cn_gather_wait(cn_handle* h, const unsigned long long* arrive_local_dev, int n_ranks, int rank, void* stream) {
    if (!h || !arrive_local_dev) return fail(CN_ERR_INVALID, "cn_gather_wait: null argument%s", NULL);
    if (n_ranks < 2 || n_ranks > 32 || rank < 0 || rank >= n_ranks) return fail(CN_ERR_INVALID, "cn_gather_wait: 2..32 ranks, 0 <= rank < n_ranks%s", NULL);
    cn_device_guard guard(h->device);
    CN_CUDA(cn_launch_gather_wait(arrive_local_dev, n_ranks, rank, h->gather_timeouts, (cudaStream_t)stream));
    h->launches += 1;
    return CN_OK;
}

int cn_gather_timeouts(cn_handle* h, unsigned int* out_host, void* stream) {
    if (!h || !out_host) return fail(CN_ERR_INVALID, "cn_gather_timeouts: null argument%s", NULL);
    cn_device_guard guard(h->device);
    unsigned int w[3] = {0u, 0u, 0u};                 /* [0] waits given up, [1] CTA counter of the running launch, [2] wire saturations */
    CN_CUDA(cudaMemcpyAsync(w, h->gather_timeouts, sizeof(w), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *out_host = w[0] + w[2];
    return CN_OK;
}

/* ---- n steps as one CUDA graph (no torch needed for graph-replayed rollouts of open-loop action batches) */
struct cn_graph {
    cn_handle* h;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    int n_steps;
    int launches_per_replay;
};

int cn_graph_create(cn_handle* h, int n_steps, const float* action_dev, size_t action_stride, float* obs_dev,
                    float* reward_dev, uint8_t* done_dev, size_t out_stride, cn_graph** out) {
    if (!h || !out || !action_dev || !obs_dev || !reward_dev || !done_dev)
        return fail(CN_ERR_INVALID, "cn_graph_create: null argument%s", NULL);
    *out = NULL;
    if (n_steps < 1) return fail(CN_ERR_INVALID, "cn_graph_create: n_steps < 1%s", NULL);
    cn_device_guard guard(h->device);
    (void)rows_in_device_memory(h, obs_dev);            /* classify the row buffer before the capture starts */
    cudaStream_t cap;
    CN_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    const int64_t before = h->launches;
    cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) { cudaStreamDestroy(cap); return fail(CN_ERR_CUDA, "cn_graph_create: begin capture: %s", cudaGetErrorString(e)); }
    int rc = cn_step_n(h, n_steps, action_dev, action_stride, obs_dev, reward_dev, done_dev, out_stride, cap);
    cudaGraph_t g = NULL;
    e = cudaStreamEndCapture(cap, &g);
    const int per_replay = (int)(h->launches - before);
    h->launches = before;                               /* captured, not launched */
    cudaStreamDestroy(cap);
    if (rc != CN_OK) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess || !g) return fail(CN_ERR_CUDA, "cn_graph_create: end capture: %s", cudaGetErrorString(e));
    cudaGraphExec_t ex = NULL;
    e = cudaGraphInstantiate(&ex, g, 0);
    if (e != cudaSuccess) { cudaGraphDestroy(g); return fail(CN_ERR_CUDA, "cn_graph_create: instantiate: %s", cudaGetErrorString(e)); }
    cn_graph* G = new (std::nothrow) cn_graph();
    if (!G) { cudaGraphExecDestroy(ex); cudaGraphDestroy(g); return fail(CN_ERR_NOMEM, "cn_graph_create: host allocation failed%s", NULL); }
    G->h = h; G->graph = g; G->exec = ex; G->n_steps = n_steps; G->launches_per_replay = per_replay;
    *out = G;
    return CN_OK;
}

int cn_graph_launch(cn_graph* g, void* stream) {
    if (!g) return fail(CN_ERR_INVALID, "cn_graph_launch: null graph%s", NULL);
    cn_device_guard guard(g->h->device);
    CN_CUDA(cudaGraphLaunch(g->exec, (cudaStream_t)stream));
    g->h->launches += g->launches_per_replay;
    return CN_OK;
}

int cn_graph_destroy(cn_graph* g) {
    if (!g) return CN_OK;
    cn_device_guard guard(g->h->device);
    cudaGraphExecDestroy(g->exec);
    cudaGraphDestroy(g->graph);
    delete g;
    return CN_OK;
}

int cn_get_counters(cn_handle* h, int32_t* out_dev, void* stream) {
    if (!h || !out_dev) return fail(CN_ERR_INVALID, "cn_get_counters: null argument%s", NULL);
    if (((uintptr_t)out_dev) & 15u) return fail(CN_ERR_INVALID, "cn_get_counters: out_dev must be 16-byte aligned%s", NULL);
    cn_device_guard guard(h->device);
    if (h->trk) CN_CUDA(cn_launch_faithful_counters(h->robot, h->trk, out_dev, h->cfg.n_envs, (cudaStream_t)stream));
    else CN_CUDA(cn_launch_counters(h->robot, out_dev, h->cfg.n_envs, (cudaStream_t)stream));
    h->launches += 1;
    return CN_OK;
}

int cn_clear_done(cn_handle* h, const uint8_t* mask_dev, void* stream) {
    if (!h) return fail(CN_ERR_INVALID, "cn_clear_done: null handle%s", NULL);
    cn_device_guard guard(h->device);
    CN_CUDA(cn_launch_clear_done(h->robot, mask_dev, h->cfg.n_envs, (cudaStream_t)stream));
    h->launches += 1;
    return CN_OK;
}

int cn_get_blob(cn_handle* h, void* host, size_t bytes, void* stream) {
    if (!h || !host) return fail(CN_ERR_INVALID, "cn_get_blob: null argument%s", NULL);
    if (bytes != cn_blob_words(&h->cfg) * 4) return fail(CN_ERR_INVALID, "cn_get_blob: size mismatch%s", NULL);
    cn_device_guard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    uint32_t* w = (uint32_t*)host;
    cn_blob_header(&h->cfg, w);
    w += CN_BLOB_HEADER_WORDS;
    const size_t rw = cn_robot_words(&h->cfg), pw = cn_ped_plane_words(&h->cfg);
    CN_CUDA(cudaMemcpyAsync(w, h->robot, rw * 4, cudaMemcpyDeviceToHost, s));
    if (pw) {
        CN_CUDA(cudaMemcpyAsync(w + rw, h->ped_a, pw * 4, cudaMemcpyDeviceToHost, s));
        CN_CUDA(cudaMemcpyAsync(w + rw + pw, h->ped_b, pw * 4, cudaMemcpyDeviceToHost, s));
    }
    if (h->trk) CN_CUDA(cudaMemcpyAsync(w + rw + 2 * pw, h->trk, cn_trk_words(&h->cfg) * 4, cudaMemcpyDeviceToHost, s));
    CN_CUDA(cudaStreamSynchronize(s));
    return CN_OK;
}

int cn_set_blob(cn_handle* h, const void* host, size_t bytes, void* stream) {
    if (!h || !host) return fail(CN_ERR_INVALID, "cn_set_blob: null argument%s", NULL);
    if (bytes != cn_blob_words(&h->cfg) * 4) return fail(CN_ERR_INVALID, "cn_set_blob: size mismatch%s", NULL);
    const uint32_t* w = (const uint32_t*)host;
    {
        /* a blob is only accepted by a handle of the same shape, layout version and risk-block mode */
        uint32_t want[CN_BLOB_HEADER_WORDS];
        cn_blob_header(&h->cfg, want);
        if (memcmp(w, want, sizeof(want)) != 0)
            return fail(CN_ERR_INVALID, "cn_set_blob: blob header does not match this handle (magic, layout version, "
                                        "n_envs, n_peds, n_samples, k_obstacles, risk-block mode, tracker record size)%s", NULL);
    }
    cn_device_guard guard(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    w += CN_BLOB_HEADER_WORDS;
    const size_t rw = cn_robot_words(&h->cfg), pw = cn_ped_plane_words(&h->cfg);
    CN_CUDA(cudaMemcpyAsync(h->robot, w, rw * 4, cudaMemcpyHostToDevice, s));
    if (pw) {
        CN_CUDA(cudaMemcpyAsync(h->ped_a, w + rw, pw * 4, cudaMemcpyHostToDevice, s));
        CN_CUDA(cudaMemcpyAsync(h->ped_b, w + rw + pw, pw * 4, cudaMemcpyHostToDevice, s));
    }
    if (h->trk) CN_CUDA(cudaMemcpyAsync(h->trk, w + rw + 2 * pw, cn_trk_words(&h->cfg) * 4, cudaMemcpyHostToDevice, s));
    CN_CUDA(cudaStreamSynchronize(s));
    return CN_OK;
}

int cn_set_debug_taps(cn_handle* h, float* ranges_dev, uint8_t* hit_ids_dev) {
    if (!h) return fail(CN_ERR_INVALID, "cn_set_debug_taps: null handle%s", NULL);
    h->dbg_ranges = ranges_dev; h->dbg_hid = hit_ids_dev;
    repack(h);
    return CN_OK;
}

int64_t cn_launch_count(const cn_handle* h) { return h ? h->launches : 0; }
const char* cn_kernel_name(const cn_handle* h) { return (h && !h->use_flat) ? "cn_env_kernel" : "cn_flat_kernel"; }
/* worlds per CTA of a plain cn_step (the direct-rows layout when the handle has one; the fused-gather entry points use
 * the staged layout, whose CTA count cn_kernel_ctas reports) */
int cn_kernel_tile(const cn_handle* h) { return !h ? 0 : (h->use_flat ? (h->have_direct ? h->flat_direct.W : h->flat.W) : CN_TILE); }
int cn_kernel_ctas(const cn_handle* h) {
    if (!h) return 0;
    const int W = h->use_flat ? h->flat.W : CN_TILE;
    return (h->cfg.n_envs + W - 1) / W;
}

int cn_plan_tile(const cn_config* cfg, int n_sms, size_t smem_per_sm, int* tile, int* threads, size_t* smem_bytes) {
    if (!cfg || !tile || !threads || !smem_bytes) return fail(CN_ERR_INVALID, "cn_plan_tile: null argument%s", NULL);
    cn_derived d;
    if (cn_derive(cfg, &d) != 0) return fail(CN_ERR_INVALID, "cn_plan_tile: config out of range%s", NULL);
    cn_flat_layout L; memset(&L, 0, sizeof(L));
    if (cn_flat_pick_tile(cfg->n_peds, cfg->n_samples, d.obs_dim, cfg->n_envs, n_sms, smem_per_sm, stage_mode(cfg), 0, &L) != 0)
        return fail(CN_ERR_UNSUPPORTED, "cn_plan_tile: no tile fits%s", NULL);
    *tile = L.W; *threads = L.threads; *smem_bytes = L.total;
    return CN_OK;
}

int cn_plan_tile_direct(const cn_config* cfg, int n_sms, size_t smem_per_sm, int* tile, int* threads, size_t* smem_bytes) {
    if (!cfg || !tile || !threads || !smem_bytes) return fail(CN_ERR_INVALID, "cn_plan_tile_direct: null argument%s", NULL);
    cn_derived d;
    if (cn_derive(cfg, &d) != 0) return fail(CN_ERR_INVALID, "cn_plan_tile_direct: config out of range%s", NULL);
    cn_flat_layout L; memset(&L, 0, sizeof(L));
    if (cn_flat_pick_tile(cfg->n_peds, cfg->n_samples, d.obs_dim, cfg->n_envs, n_sms, smem_per_sm, 0, 1, &L) != 0)
        return fail(CN_ERR_UNSUPPORTED, "cn_plan_tile_direct: no tile fits%s", NULL);
    *tile = L.W; *threads = L.threads; *smem_bytes = L.total;
    return CN_OK;
}

}  /* extern "C" */
