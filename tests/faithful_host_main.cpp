// Race check of the risk_faithful device code: replays recorded worlds (written by tests/test_faithful.py) with 64
// "lanes" as 64 threads and compares with the recorded oracle results.  Built with -fsanitize=thread, any access
// to the shared scratch that is not ordered by a CNF_SYNC() shows up as a data race.
//   usage: faithful_host_main <records.bin>      exit code = mismatching records (ThreadSanitizer reports on stderr)
#include <cstdio>
#include "faithful_host.cpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 2;
    int32_t hdr[3];
    cnf_params P;
    if (fread(hdr, 4, 3, f) != 3 || fread(&P, sizeof(P), 1, f) != 1) return 2;
    const int n_rec = hdr[0], n = hdr[1], K = hdr[2];
    std::vector<uint32_t> trk(CNF_WORLD_WORDS), want_trk(CNF_WORLD_WORDS);
    std::vector<float> scan(n), kb(4 * K), want_kb(4 * K);
    int bad = 0;
    for (int r = 0; r < n_rec; ++r) {
        double pose[3]; int32_t step;
        if (fread(trk.data(), 4, CNF_WORLD_WORDS, f) != (size_t)CNF_WORLD_WORDS || fread(pose, 8, 3, f) != 3 ||
            fread(scan.data(), 4, n, f) != (size_t)n || fread(&step, 4, 1, f) != 1 ||
            fread(want_kb.data(), 4, 4 * K, f) != (size_t)(4 * K) ||
            fread(want_trk.data(), 4, CNF_WORLD_WORDS, f) != (size_t)CNF_WORLD_WORDS) return 2;
        cnfh_observe_lanes(&P, trk.data(), pose[0], pose[1], pose[2], scan.data(), (float)0.6f, step, kb.data(), 64);
        if (memcmp(kb.data(), want_kb.data(), 16 * K) != 0 || memcmp(trk.data(), want_trk.data(), 4 * CNF_WORLD_WORDS) != 0) ++bad;
    }
    fclose(f);
    printf("%d records, %d mismatching\n", n_rec, bad);
    return bad;
}
