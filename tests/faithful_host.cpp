// Host build of the DEVICE code of the risk_faithful block (crowdnav_b200/csrc/cn_faithful.h) with one "lane",
// so tests/test_faithful.py can check its arithmetic bit for bit against the independent oracle on a machine
// without a GPU.  A test tool: nothing in crowdnav_b200/ loads this.
#include <cstdlib>
#include <cstring>
#include <pthread.h>
#include <thread>
#include <vector>
// the "warp" can also be run as nl host threads: CNF_SYNC() becomes a barrier over them, so the stage structure
// (who writes what between which syncs) is exercised with real concurrency (and under -fsanitize=thread)
static thread_local pthread_barrier_t* cnfh_bar = nullptr;
static inline void cnfh_sync() { if (cnfh_bar) pthread_barrier_wait(cnfh_bar); }
#define CNF_SYNC() cnfh_sync()
#include "../crowdnav_b200/csrc/cn_faithful.h"

extern "C" {

int cnfh_world_words(void) { return CNF_WORLD_WORDS; }
size_t cnfh_scratch_bytes(int n) { return cnf_scratch_bytes(n); }

// one get_state of the block for one world; trk is updated in place, kblock receives 4K floats
void cnfh_observe(const cnf_params* P, uint32_t* trk, double x, double y, double yaw, const float* scan32,
                  float no_return32, int step_counter, float* kblock) {
    const size_t bytes = cnf_scratch_bytes(P->n_rays);
    unsigned char* base = (unsigned char*)aligned_alloc(16, bytes);
    memset(base, 0xA5, bytes);                    // scratch is never assumed to be zero
    cnf_scratch S;
    cnf_scratch_carve(base, P->n_rays, &S);
    memcpy(S.trk, trk, sizeof(uint32_t) * CNF_WORLD_WORDS);
    cnf_world(P, S, x, y, yaw, scan32, no_return32, step_counter, kblock, 0, 1, 0);
    memcpy(trk, S.trk, sizeof(uint32_t) * CNF_WORLD_WORDS);
    free(base);
}

// the same with nl "lanes" run as nl threads over one shared scratch
void cnfh_observe_lanes(const cnf_params* P, uint32_t* trk, double x, double y, double yaw, const float* scan32,
                        float no_return32, int step_counter, float* kblock, int nl) {
    const size_t bytes = cnf_scratch_bytes(P->n_rays);
    unsigned char* base = (unsigned char*)aligned_alloc(16, bytes);
    memset(base, 0xA5, bytes);
    cnf_scratch S;
    cnf_scratch_carve(base, P->n_rays, &S);
    memcpy(S.trk, trk, sizeof(uint32_t) * CNF_WORLD_WORDS);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, (unsigned)nl);
    std::vector<std::thread> th;
    for (int lane = 0; lane < nl; ++lane)
        th.emplace_back([&, lane]() {
            cnfh_bar = &bar;
            cnf_world(P, S, x, y, yaw, scan32, no_return32, step_counter, kblock, lane, nl, 0);
            cnfh_bar = nullptr;
        });
    for (auto& t : th) t.join();
    pthread_barrier_destroy(&bar);
    memcpy(trk, S.trk, sizeof(uint32_t) * CNF_WORLD_WORDS);
    free(base);
}

}
