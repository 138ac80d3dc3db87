// Differential fuzz of the host clustering engine (galah_b200/csrc/host/cluster_engine.cpp), meant to be built with
// sanitizers (tests/test_engine_sanitizers.py):
//   g++ -std=c++17 -O1 -g -fsanitize=address,undefined -pthread -I galah_b200/csrc tools/engine_fuzz.cpp \
//       galah_b200/csrc/host/cluster_engine.cpp -o engine_fuzz && ./engine_fuzz 0|1|2
// Every round runs the same hit list through the serial callback engine, the table engine (threaded sweeps from 4,096
// preclusters), the wave engine and the wave engine with a budget of 1-3 waves (one-batch finish), with Nones and
// orientation-dependent values, and compares clusters, order and calculate_ani counts.
// mode 0: 300 small random graphs (also shuffled input); 1: 60,000 genomes in families of 8 (threaded sweeps);
// mode 2: one clade of 1,100 genomes, 604,450 hits (threaded adjacency fill).
#include "host/cluster_engine.hpp"
#include <cstdio>
#include <random>
#include <map>
using namespace gb200;
int main(int argc, char **argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;  // 0: small random graphs, 1: many preclusters (threaded sweeps), 2: dense (threaded fill)
    std::mt19937_64 rng(argc > 2 ? strtoull(argv[2], nullptr, 10) : 7);  // optional second argument: the seed
    int rounds = mode == 0 ? 300 : 3;
    for (int round = 0; round < rounds; round++) {
        size_t n = mode == 0 ? 1 + rng() % 120 : mode == 1 ? 60000 : 1100;
        std::vector<PreclusterHit> hits;
        if (mode == 2) { for (uint32_t a = 0; a < n; a++) for (uint32_t b = a + 1; b < n; b++) hits.push_back({a, b, 0.95f}); }
        else {
            size_t fam = mode == 0 ? 1 + rng() % 12 : 8;
            for (uint32_t base = 0; base < n; base += fam)
                for (uint32_t a = base; a < std::min<size_t>(n, base + fam); a++)
                    for (uint32_t b = a + 1; b < std::min<size_t>(n, base + fam); b++)
                        if (rng() % 10 < 7) hits.push_back({a, b, 0.9f + (rng() % 100) / 1000.f});
            if (mode == 0) for (size_t x = 0; x < n / 6; x++) { uint32_t a = rng() % n, b = rng() % n; if (a < b) hits.push_back({a, b, 0.91f}); }
        }
        if (mode == 0 && round % 3 == 0) std::shuffle(hits.begin(), hits.end(), rng);
        auto val = [&](uint32_t rep, uint32_t g, bool &some) -> float {
            uint64_t h = (rep * 0x9E3779B97F4A7C15ull) ^ (g * 0xC2B2AE3D27D4EB4Full); h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
            some = mode != 0 || (h % 17) != 0;
            return mode == 2 ? (((rep ^ g) & 1) ? 80.f : 98.f) : 90.f + (h % 1000) / 100.f;
        };
        AniFn fn = [&](uint32_t rep, uint32_t g, float *ani) { bool s; *ani = val(rep, g, s); return s; };
        AniByHitFn by_hit = [&](uint32_t rep, uint32_t g, size_t, float *ani) { bool s; *ani = val(rep, g, s); return s; };
        AniBatchFn batch = [&](const std::vector<AniRequest> &reqs, uint8_t *some, float *ani) {
            for (size_t q = 0; q < reqs.size(); q++) { bool s; ani[q] = val(reqs[q].rep, reqs[q].genome, s); some[q] = s; }
            return 0;
        };
        ClusterResult a, b, c, d;
        std::string e1, e2, e3, e4;
        int r1 = cluster_from_hits(n, hits.data(), hits.size(), false, 95.f, fn, a, e1);
        int r2 = cluster_from_hits(n, hits.data(), hits.size(), false, 95.f, AniFn(), b, e2, &by_hit);
        int r3 = cluster_from_hits(n, hits.data(), hits.size(), false, 95.f, AniFn(), c, e3, nullptr, nullptr, &batch, 16);
        int r4 = cluster_from_hits(n, hits.data(), hits.size(), false, 95.f, AniFn(), d, e4, nullptr, nullptr, &batch, 1 + round % 3);
        if (r1 != r2 || r1 != r3 || r1 != r4) { printf("rc mismatch %d %d %d %d round %d\n", r1, r2, r3, r4, round); return 1; }
        if (r1) continue;
        if (a.members != b.members || a.offsets != b.offsets || a.members != c.members || a.offsets != c.offsets ||
            a.members != d.members || a.offsets != d.offsets || a.ani_calls != b.ani_calls || a.ani_calls != c.ani_calls) {
            printf("MISMATCH round %d n %zu hits %zu calls %lu %lu %lu\n", round, n, hits.size(), a.ani_calls, b.ani_calls, c.ani_calls);
            return 1;
        }
        ClusterResult s1; std::string e5;
        if (cluster_from_hits(n, hits.data(), hits.size(), true, 0.95f, AniFn(), s1, e5)) { printf("skip failed\n"); return 1; }
    }
    printf("mode %d ok\n", mode);
    return 0;
}
