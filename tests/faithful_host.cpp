// Host build of the DEVICE code of the risk_faithful block (crowdnav_b200/csrc/cn_faithful.h) with one "lane",
// so tests/test_faithful.py can check its arithmetic bit for bit against the independent oracle on a machine
// without a GPU.  A test tool: nothing in crowdnav_b200/ loads this.
#include <cstdlib>
#include <cstring>
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
    cnf_world(P, &S, x, y, yaw, scan32, no_return32, step_counter, kblock, 0, 1);
    memcpy(trk, S.trk, sizeof(uint32_t) * CNF_WORLD_WORDS);
    free(base);
}

}
