/*
 * cn_faithful.cu -- kernel of the `risk_faithful` perception block (CN_FLAG_RISK_FAITHFUL): two warps (64 threads, their own named barrier) per world
 * run cnf_world() (cn_faithful.h) behind the step kernel, on the same stream.  It reads what the step kernel
 * left in HBM -- the robot record (pose on the integer grid, steps taken: 0 = the world was (re)started by this
 * launch) and the cleaned pre-rounding ranges -- keeps the world's tracker record in shared memory, and
 * overwrites the K block of the observation row (ENV:862-907) and the safety counters (ENV:653-654, 998-1005).
 *
 * Bound: latency of the order-dependent walks on lane 0 (typing state machine, segment walk), not bandwidth:
 * per world it moves 4 (R-1) B of ranges, 2 x 1584 B of tracker record and 16 K B of row.
 */
#include <cuda_runtime.h>
#include "cn_kernel.h"
#include "cn_faithful.h"

#define CNF_WORLDS 4            /* worlds per CTA, CNF_LANES (64) threads each */

__global__ void __launch_bounds__(CNF_LANES * CNF_WORLDS, 4)
cn_faithful_kernel(cnf_params P, const uint32_t* __restrict__ robot, uint32_t* __restrict__ trk,
                   const float* __restrict__ ranges, float* __restrict__ obs, const uint8_t* __restrict__ mask,
                   int E, int obs_dim, float no_return32, unsigned scratch_bytes) {
    extern __shared__ __align__(16) unsigned char cnf_smem[];
    const int wi = threadIdx.x / CNF_LANES, lane = threadIdx.x % CNF_LANES;
    const int e = blockIdx.x * CNF_WORLDS + wi;
    const int bar = 1 + wi;                                    /* named barrier of this world's threads */
    if (e >= E) return;                                        /* whole groups leave together */
    if (mask != nullptr && mask[e] == 0) return;              /* masked reset: untouched worlds keep their tracker */
    cnf_scratch S;
    cnf_scratch_carve(cnf_smem + (size_t)wi * scratch_bytes, P.n_rays, &S);
    const uint32_t* rob = robot + (size_t)e * CN_ROBOT_WORDS;
    uint32_t* g = trk + (size_t)e * CNF_WORLD_WORDS;
    static_assert(CNF_WORLD_WORDS % 4 == 0, "the tracker record moves as 16-byte words");
    // (16-byte words: 99 per world, two trips per lane instead of seven; the plane and the scratch are 16-byte aligned)
    for (int k = lane; k < CNF_WORLD_WORDS / 4; k += CNF_LANES)
        reinterpret_cast<uint4*>(S.trk)[k] = reinterpret_cast<const uint4*>(g)[k];
    CNF_SYNC();
    const double x = (double)((float)(int32_t)rob[CN_R_X] * CN_GRID);
    const double y = (double)((float)(int32_t)rob[CN_R_Y] * CN_GRID);
    const double yaw = (double)cn_bin2rad(rob[CN_R_TH]);
    const int step_counter = (int)rob[CN_R_STEP];
    cnf_world(&P, S, x, y, yaw, ranges + (size_t)e * P.n_rays, no_return32, step_counter,
              obs + (size_t)e * obs_dim + P.n_rays + 7, lane, CNF_LANES, bar);
    for (int k = lane; k < CNF_WORLD_WORDS / 4; k += CNF_LANES)
        reinterpret_cast<uint4*>(g)[k] = reinterpret_cast<const uint4*>(S.trk)[k];
}

/* [E, 4] int32: success (robot record), ego / social violations, obstacle-present steps (tracker record) */
__global__ void cn_faithful_counters_kernel(const uint32_t* __restrict__ robot, const uint32_t* __restrict__ trk,
                                            int32_t* __restrict__ out, int E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const uint32_t* t = trk + (size_t)e * CNF_WORLD_WORDS;
    int4 v;
    v.x = (robot[(size_t)e * CN_ROBOT_WORDS + CN_R_FLAGS] & CN_RF_SUCCESS) ? 1 : 0;
    v.y = (int32_t)t[CNF_H_EGO]; v.z = (int32_t)t[CNF_H_SOCIAL]; v.w = (int32_t)t[CNF_H_PRESENT];
    reinterpret_cast<int4*>(out)[e] = v;
}

#ifdef CN_TIMELINE
extern "C" int cn_debug_set_timeline_faithful(unsigned long long* dev_ptr) {
    return (int)cudaMemcpyToSymbol(cnf_timeline, &dev_ptr, sizeof(dev_ptr));
}
#endif

size_t cn_faithful_smem_bytes(int n_rays) { return CNF_WORLDS * cnf_scratch_bytes(n_rays); }

cudaError_t cn_launch_faithful(const cn_config* cfg, const uint32_t* robot, uint32_t* trk, const float* ranges,
                               float* obs, const uint8_t* mask, int obs_dim, cudaStream_t stream) {
    cnf_params P; cnf_params_from_config(cfg, &P);
    const size_t per = cnf_scratch_bytes(P.n_rays), smem = CNF_WORLDS * per;
    {
        cudaError_t e = cn_ensure_smem_attr(reinterpret_cast<const void*>(cn_faithful_kernel), 14, smem);
        if (e != cudaSuccess) return e;
    }
    const int E = cfg->n_envs;
    cn_faithful_kernel<<<(E + CNF_WORLDS - 1) / CNF_WORLDS, CNF_LANES * CNF_WORLDS, smem, stream>>>(
        P, robot, trk, ranges, obs, mask, E, obs_dim, cfg->max_range, (unsigned)per);
    return cudaGetLastError();
}

cudaError_t cn_launch_faithful_counters(const uint32_t* robot, const uint32_t* trk, int32_t* out, int E, cudaStream_t stream) {
    cn_faithful_counters_kernel<<<(E + 255) / 256, 256, 0, stream>>>(robot, trk, out, E);
    return cudaGetLastError();
}
