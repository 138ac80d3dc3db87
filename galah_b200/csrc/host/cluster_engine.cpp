// Greedy representative selection + membership assignment of galah::clusterer::cluster()
// (/root/reference/src/clusterer.rs:14-152), restated over a sparse pair list.
//
// What is kept bit-for-bit from the reference (at --threads 1, where its output order is a
// function of the input):
//   * partition_sketches (:452-487): single-linkage components of the precluster hits;
//     `sets()` ordered by smallest member, members ascending (:67-76); preclusters ordered by
//     size, largest first (:79; ties keep smallest-member order, as Rust's sort_unstable does
//     for the <= 20-element insertion-sort case -- beyond that the reference's own order is
//     unspecified).
//   * find_precluster_cluster_representatives (:182-259): genomes in ascending index order;
//     candidates = existing representatives with a precluster hit, sorted by precluster ANI
//     ASCENDING (:200); calculate_ani in that order until one value reaches the threshold
//     (find_any on one thread, :276-296); genome i is a representative iff no computed ANI
//     satisfies `ani >= threshold` in f32 (:241).  Only Some(ani) values enter the cache (:236-239).
//   * find_precluster_cluster_memberships (:350-449): for every non-representative, ANI is
//     computed against every representative with a precluster hit that is not cached yet (no
//     early stop, None results are cached too, :398-405); the genome joins the representative
//     with the highest ANI, strict `>` so ties go to the lowest representative index (:436-442);
//     no representative with an ANI -> the reference panics on `best_rep.unwrap()` (:444).
//   * skip_clusterer (:32-44, :209-214, :253-254): ANI values are read from the precluster
//     cache and the whole precluster cache is handed to the membership stage.
//
// What is different by design: the reference probes a BTreeMap for all N(N-1)/2 pairs in
// partition_sketches and all m(m-1)/2 pairs per precluster in transform_ids
// (src/sorted_pair_genome_distance_cache.rs:47-58); here both are driven by the hit list
// (CSR adjacency + union-find), O(hits) instead of O(N^2) -- SURVEY.md 8f item 1.
#include "cluster_engine.hpp"

#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <new>
#include <numeric>
#include <thread>

namespace gb200 {

namespace {

struct Dsu {
    std::vector<uint32_t> parent;
    explicit Dsu(size_t n) : parent(n) { std::iota(parent.begin(), parent.end(), 0u); }
    uint32_t find(uint32_t x) {
        while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; }
        return x;
    }
    void join(uint32_t a, uint32_t b) {
        a = find(a); b = find(b);
        if (a != b) parent[std::max(a, b)] = std::min(a, b);
    }
};

// Edge arrays are written exactly once right after they are sized: plain buffers that are not cleared
// first (a vector's resize touches every page a second time; at millions of hits that is the cost).
template <typename T>
struct RawBuf {
    T *p = nullptr;
    size_t n = 0;
    RawBuf() = default;
    RawBuf(const RawBuf &) = delete;
    RawBuf &operator=(const RawBuf &) = delete;
    ~RawBuf() { free(p); }
    void resize(size_t count, bool zero = false) {
        free(p);
        n = count;
        p = (T *)(zero ? calloc(std::max<size_t>(count, 1), sizeof(T)) : malloc(std::max<size_t>(count, 1) * sizeof(T)));
        if (!p) throw std::bad_alloc();
    }
    size_t size() const { return n; }
    T *data() { return p; }
    const T *data() const { return p; }
    T &operator[](size_t x) { return p[x]; }
    const T &operator[](size_t x) const { return p[x]; }
};

struct Adjacency {  // CSR over both directions, neighbours ascending
    std::vector<uint64_t> off;
    RawBuf<uint32_t> nbr;
    RawBuf<float> ani;
    RawBuf<uint32_t> hit;  // index of the precluster hit behind the slot
    // index of (a, b) in nbr/ani, or -1
    int64_t find(uint32_t a, uint32_t b) const {
        const uint32_t *lo = nbr.data() + off[a], *hi = nbr.data() + off[a + 1];
        const uint32_t *it = std::lower_bound(lo, hi, b);
        return (it != hi && *it == b) ? (int64_t)(it - nbr.data()) : -1;
    }
};

inline uint64_t pair_key(uint32_t a, uint32_t b) {
    return a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a;
}

struct OptAni { bool some; float ani; };

// CSR adjacency of the hit list, both directions, neighbours ascending.  Later duplicates of a key overwrite
// earlier ones, as BTreeMap::insert does (src/sorted_pair_genome_distance_cache.rs:22-28).
void build_adjacency(size_t n, const PreclusterHit *hits, size_t n_hits, Adjacency &adj) {
    adj.off.assign(n + 1, 0);
    // unique keys, last write wins.  The usual input -- the preclusterer's own output -- is already
    // strictly ascending by key: then the hits are used as they are; otherwise stable sort by key and
    // keep the last record of every run.
    bool sorted = true;
    for (size_t h = 1; h < n_hits && sorted; h++)
        sorted = pair_key(hits[h - 1].i, hits[h - 1].j) < pair_key(hits[h].i, hits[h].j);
    std::vector<uint32_t> uniq;  // indices of the hits that count, ascending by key (unsorted input only)
    if (!sorted) {
        uniq.reserve(n_hits);
        std::vector<std::pair<uint64_t, uint32_t>> keyed(n_hits);
        for (size_t h = 0; h < n_hits; h++) keyed[h] = {pair_key(hits[h].i, hits[h].j), (uint32_t)h};
        std::stable_sort(keyed.begin(), keyed.end(),
                         [](const std::pair<uint64_t, uint32_t> &a, const std::pair<uint64_t, uint32_t> &b) { return a.first < b.first; });
        for (size_t h = 0; h < n_hits; h++)
            if (h + 1 == n_hits || keyed[h + 1].first != keyed[h].first) uniq.push_back(keyed[h].second);
    }
    const size_t n_uniq = sorted ? n_hits : uniq.size();
    auto hit_at = [&](size_t u) -> uint32_t { return sorted ? (uint32_t)u : uniq[u]; };
    for (size_t u = 0; u < n_uniq; u++) {
        const PreclusterHit &h = hits[hit_at(u)];
        adj.off[std::min(h.i, h.j) + 1]++;
        adj.off[std::max(h.i, h.j) + 1]++;
    }
    for (size_t g = 0; g < n; g++) adj.off[g + 1] += adj.off[g];
    adj.nbr.resize(adj.off[n]); adj.ani.resize(adj.off[n]); adj.hit.resize(adj.off[n]);
    std::vector<uint64_t> fill(adj.off.begin(), adj.off.end() - 1);
    // keys ascend, so every row receives its neighbours in ascending order: first the smaller
    // partners (as the second genome of earlier keys), then the larger ones
    // rows [r0, r1): every thread of a large list walks all hits in key order and writes the slots of its own rows
    auto fill_rows = [&](uint32_t r0, uint32_t r1) {
        for (size_t u = 0; u < n_uniq; u++) {
            const uint32_t x = hit_at(u);
            const PreclusterHit &h = hits[x];
            const uint32_t a = std::min(h.i, h.j), b = std::max(h.i, h.j);
            if (a >= r0 && a < r1) { const uint64_t pa = fill[a]++; adj.nbr[pa] = b; adj.ani[pa] = h.ani; adj.hit[pa] = x; }
            if (b >= r0 && b < r1) { const uint64_t pb = fill[b]++; adj.nbr[pb] = a; adj.ani[pb] = h.ani; adj.hit[pb] = x; }
        }
    };
    const size_t hw_fill = std::max<size_t>(1, std::thread::hardware_concurrency());
    const size_t nt_fill = n_uniq >= ((size_t)1 << 19) ? std::min<size_t>(hw_fill, 16) : 1;
    if (nt_fill == 1) {
        fill_rows(0, (uint32_t)n);
    } else {
        // row ranges holding about the same number of slots (the scattered writes are the cost)
        std::vector<uint32_t> cut(nt_fill + 1, (uint32_t)n);
        cut[0] = 0;
        for (size_t t = 1; t < nt_fill; t++)
            cut[t] = (uint32_t)(std::lower_bound(adj.off.begin(), adj.off.end(), adj.off[n] / nt_fill * t) - adj.off.begin());
        std::vector<std::thread> th;
        for (size_t t = 0; t < nt_fill; t++)
            if (cut[t] < cut[t + 1]) th.emplace_back(fill_rows, cut[t], cut[t + 1]);
        for (auto &t : th) t.join();
    }
}

// partition_sketches (src/clusterer.rs:452-487): single-linkage components of the hit graph, as one flat array
// (CSR): sets in order of their smallest member, members ascending, then ordered by size, largest first
// (src/clusterer.rs:67-79; stable, so equal sizes keep the order of their smallest members).
void partition_preclusters(size_t n, const Adjacency &adj, std::vector<uint32_t> &pc_members, std::vector<uint64_t> &pc_off,
                           uint32_t &n_preclusters, uint32_t &largest_precluster) {
    Dsu dsu(n);
    for (size_t g = 0; g < n; g++)
        for (uint64_t x = adj.off[g]; x < adj.off[g + 1]; x++)
            if (adj.nbr[x] < g) dsu.join((uint32_t)g, adj.nbr[x]);
    // preclusters as one flat array (CSR): sets in order of their smallest member (the root),
    // members ascending, then ordered by size, largest first (stable)
    pc_members.assign(n, 0);
    {
        std::vector<uint32_t> set_id(n), set_size;
        std::vector<int64_t> set_of_root(n, -1);
        for (uint32_t g = 0; g < n; g++) {
            const uint32_t r = dsu.find(g);
            if (set_of_root[r] < 0) { set_of_root[r] = (int64_t)set_size.size(); set_size.push_back(0); }
            set_id[g] = (uint32_t)set_of_root[r];
            set_size[set_id[g]]++;
        }
        // sets by size, largest first, ties in set order: a counting sort on the size (stable), O(sets + largest)
        std::vector<uint32_t> order(set_size.size());
        {
            uint32_t largest = 0;
            for (const uint32_t z : set_size) largest = std::max(largest, z);
            std::vector<uint64_t> at((size_t)largest + 2, 0);  // at[z] = first slot of size z, sizes descending
            for (const uint32_t z : set_size) at[largest - z + 1]++;
            for (size_t z = 0; z + 1 < at.size(); z++) at[z + 1] += at[z];
            for (uint32_t sid = 0; sid < set_size.size(); sid++) order[at[largest - set_size[sid]]++] = sid;
        }
        std::vector<uint64_t> start(set_size.size());
        pc_off.assign(1, 0);
        for (const uint32_t sid : order) { start[sid] = pc_off.back(); pc_off.push_back(pc_off.back() + set_size[sid]); }
        for (uint32_t g = 0; g < n; g++) pc_members[start[set_id[g]]++] = g;
        n_preclusters = (uint32_t)set_size.size();
        largest_precluster = set_size[order[0]];
    }
}

// Representatives found lazily, in waves (AniBatchFn), all preclusters at once.
// state: 0 open, 1 representative, 2 member.  pending[g] = hit partners below g that are not settled
// towards g yet: a lower partner settles when it becomes a member, or -- a representative -- when its
// ANI with g has been applied.  An open genome whose pending count reaches zero without having been
// claimed (ANI >= threshold with a representative below it) is a representative: exactly
// src/clusterer.rs:216-259, where genome i is tested against the representatives found before it.
// Every request is a pair the reference evaluates too: (representative, partner) for every partner
// that is not an earlier representative (tested in the representative pass if the partner comes
// later and was still undecided, asked for by the membership pass otherwise).
// Fills state (1 representative, 2 member), the edge cache with every value asked for, `calls` with the number
// of pairs asked, `waves_asked` with the number of batches; any_none: some answer was None.
int representatives_in_waves(size_t n, const Adjacency &adj, float ani_threshold, const AniBatchFn &batch_fn, uint32_t max_waves,
                             RawBuf<uint8_t> &edge_state, RawBuf<float> &edge_ani, std::vector<uint8_t> &state, uint64_t &calls,
                             uint32_t &waves_asked, bool &any_none, std::string &err) {
    state.assign(n, 0);
    std::vector<uint32_t> pending(n, 0), fresh, ready, claimed;
    for (uint32_t g = 0; g < n; g++) {
        for (uint64_t x = adj.off[g]; x < adj.off[g + 1] && adj.nbr[x] < g; x++) pending[g]++;
        if (!pending[g]) { state[g] = 1; fresh.push_back(g); }
    }
    std::vector<AniRequest> reqs;
    std::vector<uint64_t> slot;  // per request: the edge slot in the GENOME's row (the cache key)
    std::vector<uint8_t> some;
    std::vector<float> ani;
    auto ask = [&]() -> int {
        if (reqs.empty()) return 0;
        some.assign(reqs.size(), 0); ani.assign(reqs.size(), 0.f);
        if (batch_fn(reqs, some.data(), ani.data())) { err = "ANI batch failed"; return 2; }
        waves_asked++;
        calls += reqs.size();
        for (size_t q = 0; q < reqs.size(); q++) {
            edge_state[slot[q]] = some[q] ? 1 : 2; edge_ani[slot[q]] = ani[q];
            any_none = any_none || !some[q];
        }
        return 0;
    };
    uint32_t waves = 0;
    while (!fresh.empty() && waves < max_waves) {
        waves++;
        reqs.clear(); slot.clear();
        for (const uint32_t r : fresh)
            for (uint64_t y = adj.off[r]; y < adj.off[r + 1]; y++) {
                const uint32_t i = adj.nbr[y];
                if (state[i] == 1) continue;  // an earlier representative: (i, r) was asked for when i was confirmed
                reqs.push_back(AniRequest{r, i, adj.hit[y]});
                slot.push_back((uint64_t)adj.find(i, r));  // the same edge in i's row
            }
        if (int rc = ask()) return rc;
        ready.clear(); claimed.clear();
        for (size_t q = 0; q < reqs.size(); q++) {
            const uint32_t r = reqs[q].rep, i = reqs[q].genome;
            if (i < r) continue;  // a member below the representative: only its membership needs the value
            if (state[i] == 0 && some[q] && ani[q] >= ani_threshold) { state[i] = 2; claimed.push_back(i); }
            if (--pending[i] == 0) ready.push_back(i);
        }
        for (const uint32_t m : claimed)
            for (uint64_t y = adj.off[m + 1]; y-- > adj.off[m] && adj.nbr[y] > m;)
                if (--pending[adj.nbr[y]] == 0) ready.push_back(adj.nbr[y]);
        fresh.clear();
        for (const uint32_t g : ready)
            if (state[g] == 0) { state[g] = 1; fresh.push_back(g); }
    }
    if (!fresh.empty()) {
        // wave budget spent (long chains of mutually distant genomes): everything the undecided genomes
        // can still ask about in ONE batch -- the pairs of the representatives confirmed last, and for
        // every open genome its open partners below it --, then the undecided genomes are settled in index order
        reqs.clear(); slot.clear();
        for (const uint32_t r : fresh)
            for (uint64_t y = adj.off[r]; y < adj.off[r + 1]; y++) {
                const uint32_t i = adj.nbr[y];
                if (state[i] == 1) continue;
                reqs.push_back(AniRequest{r, i, adj.hit[y]});
                slot.push_back((uint64_t)adj.find(i, r));  // the same edge in i's row
            }
        for (uint32_t i = 0; i < n; i++) {
            if (state[i] != 0) continue;
            for (uint64_t x = adj.off[i]; x < adj.off[i + 1] && adj.nbr[x] < i; x++)
                if (state[adj.nbr[x]] == 0) { reqs.push_back(AniRequest{adj.nbr[x], i, adj.hit[x]}); slot.push_back(x); }
        }
        if (int rc = ask()) return rc;
        for (uint32_t i = 0; i < n; i++) {
            if (state[i] != 0) continue;
            state[i] = 1;
            for (uint64_t x = adj.off[i]; x < adj.off[i + 1] && adj.nbr[x] < i; x++)
                if (state[adj.nbr[x]] == 1 && edge_state[x] == 1 && edge_ani[x] >= ani_threshold) { state[i] = 2; break; }
        }
    }
    // what the membership sweep reads and nobody asked for yet (only after the one-batch finish)
    reqs.clear(); slot.clear();
    for (uint32_t i = 0; i < n; i++) {
        if (state[i] != 2) continue;
        for (uint64_t x = adj.off[i]; x < adj.off[i + 1]; x++)
            if (state[adj.nbr[x]] == 1 && !edge_state[x]) { reqs.push_back(AniRequest{adj.nbr[x], i, adj.hit[x]}); slot.push_back(x); }
    }
    if (int rc = ask()) return rc;
    return 0;
}

}  // namespace

int cluster_from_hits(size_t n, const PreclusterHit *hits, size_t n_hits, bool skip_clusterer,
                      float ani_threshold, const AniFn &calculate_ani_fn, ClusterResult &out,
                      std::string &err, const AniByHitFn *by_hit, const ReversePrefetchFn *prefetch_reverse,
                      const AniBatchFn *batch, uint32_t max_waves) {
    out = ClusterResult();
    out.offsets.push_back(0);
    const bool dbg = getenv("GALAH_B200_DEBUG") != nullptr;
    auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = now();
    if (n == 0) {
        // the reference indexes preclusters[0] unconditionally (src/clusterer.rs:84)
        err = "index out of bounds: the len is 0 but the index is 0";
        return 1;
    }
    for (size_t h = 0; h < n_hits; h++) {
        if (hits[h].i >= n || hits[h].j >= n || hits[h].i == hits[h].j) {
            err = "precluster hit with genome index out of range";
            return 1;
        }
    }
    if (!skip_clusterer && !calculate_ani_fn && !by_hit && !batch) { err = "calculate_ani callback required"; return 1; }
    const bool lazy = batch && !skip_clusterer;

    // ---- adjacency
    Adjacency adj;
    build_adjacency(n, hits, n_hits, adj);

    const double t1 = now();
    // ---- partition_sketches
    std::vector<uint32_t> pc_members;
    std::vector<uint64_t> pc_off;
    partition_preclusters(n, adj, pc_members, pc_off, out.n_preclusters, out.largest_precluster);

    const double t2 = now();
    std::vector<uint8_t> is_rep(n, 0);
    std::vector<uint32_t> cluster_of_rep(n, 0);
    // clusterer cache (src/clusterer.rs:236-239, 398-405), one slot per adjacency edge: a pair is
    // only ever looked up from the row of the genome being placed, so the edge index is its key
    RawBuf<uint8_t> edge_state;  // 0 = not computed, 1 = Some(ani), 2 = None
    RawBuf<float> edge_ani;      // read only where edge_state says computed
    edge_state.resize(adj.nbr.size(), true);
    edge_ani.resize(adj.nbr.size());
    struct Cand { float pre; uint32_t j; uint64_t edge; };
    const size_t n_pc = pc_off.size() - 1;

    // Two sweeps over the preclusters: all representatives first, then all memberships -- between
    // the two a table-driven caller learns which reverse-orientation values the membership sweep is
    // going to ask for and can produce them in one batch.  Preclusters are independent of each other
    // (the reference runs them on rayon threads, src/clusterer.rs:190-214): with ANI values served
    // from tables (or skip_clusterer) the sweeps are split over the host threads; a caller-supplied
    // calculate_ani callback is only ever invoked from the calling thread, in precluster order.
    // A precluster's clusters partition its members, so precluster pc owns members[pc_off[pc] ..
    // pc_off[pc + 1]) of the output and its representatives sit in the same range of rep_slot.
    std::vector<uint32_t> rep_slot(n), n_reps_pc(n_pc, 0);
    const size_t hw = std::max<size_t>(1, std::thread::hardware_concurrency());
    const size_t n_threads = ((by_hit || skip_clusterer || lazy) && n_pc >= 4096) ? std::min<size_t>(hw, 16) : 1;
    std::atomic<uint64_t> ani_calls{0};
    auto sweep = [&](const std::function<void(size_t, std::vector<Cand> &, std::vector<std::pair<uint32_t, uint32_t>> &,
                                              std::vector<uint64_t> &, uint64_t &)> &body) {
        std::atomic<size_t> next{0};
        auto run = [&] {
            std::vector<Cand> cands;
            std::vector<std::pair<uint32_t, uint32_t>> joins;
            std::vector<uint64_t> cl_fill;
            uint64_t calls = 0;
            for (;;) {
                const size_t p0 = next.fetch_add(256);
                if (p0 >= n_pc) break;
                for (size_t pc = p0; pc < std::min(n_pc, p0 + 256); pc++) body(pc, cands, joins, cl_fill, calls);
            }
            ani_calls.fetch_add(calls);
        };
        if (n_threads == 1) { run(); return; }
        std::vector<std::thread> th;
        for (size_t t = 0; t < n_threads; t++) th.emplace_back(run);
        for (auto &t : th) t.join();
    };

    // ---- representatives: lazily in waves (representatives_in_waves), or by the sweep below
    bool any_none = false;
    if (lazy) {
        std::vector<uint8_t> state;
        uint64_t calls = 0;
        if (int rc = representatives_in_waves(n, adj, ani_threshold, *batch, max_waves, edge_state, edge_ani, state, calls,
                                              out.ani_waves, any_none, err))
            return rc;
        for (size_t pc = 0; pc < n_pc; pc++) {
            uint32_t n_reps = 0;
            for (uint64_t m = pc_off[pc]; m < pc_off[pc + 1]; m++) {
                const uint32_t g = pc_members[m];
                if (state[g] == 0) { err = "wave engine left genome " + std::to_string(g) + " undecided"; return 3; }
                if (state[g] == 1) { is_rep[g] = 1; rep_slot[pc_off[pc] + n_reps++] = g; }
            }
            n_reps_pc[pc] = n_reps;
        }
        if (any_none) {
            // the reference caches only Some(ani) in its representative pass, so a None met there before the
            // first passing representative is computed again by the membership pass (src/clusterer.rs:236-239, 398-405)
            std::vector<Cand> cands;
            for (uint32_t i = 0; i < n; i++) {
                if (is_rep[i]) continue;
                cands.clear();
                for (uint64_t x = adj.off[i]; x < adj.off[i + 1] && adj.nbr[x] < i; x++)
                    if (is_rep[adj.nbr[x]]) cands.push_back(Cand{adj.ani[x], adj.nbr[x], x});
                std::stable_sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) { return a.pre < b.pre; });
                for (const auto &c : cands) {
                    if (edge_state[c.edge] == 2) calls++;
                    else if (edge_ani[c.edge] >= ani_threshold) break;
                }
            }
        }
        ani_calls.fetch_add(calls);
    } else
    sweep([&](size_t pc, std::vector<Cand> &cands, std::vector<std::pair<uint32_t, uint32_t>> &, std::vector<uint64_t> &,
              uint64_t &calls) {
        const uint32_t *mb = pc_members.data() + pc_off[pc], *me = pc_members.data() + pc_off[pc + 1];
        uint32_t n_reps = 0;
        for (const uint32_t *ip = mb; ip != me; ip++) {
            const uint32_t i = *ip;
            cands.clear();
            for (uint64_t x = adj.off[i]; x < adj.off[i + 1]; x++) {
                const uint32_t j = adj.nbr[x];
                if (is_rep[j]) cands.push_back(Cand{adj.ani[x], j, x});  // reps found so far all have j < i
            }
            std::stable_sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) { return a.pre < b.pre; });
            bool rep = true;
            for (const auto &c : cands) {
                float ani = c.pre;
                bool some = true;
                if (!skip_clusterer) {
                    some = by_hit ? (*by_hit)(c.j, i, adj.hit[c.edge], &ani) : calculate_ani_fn(c.j, i, &ani);
                    calls++;
                    if (some) { edge_state[c.edge] = 1; edge_ani[c.edge] = ani; }
                }
                if (some && ani >= ani_threshold) {
                    rep = false;
                    if (!skip_clusterer) break;  // find_any stops at the first hit
                }
            }
            if (rep) { is_rep[i] = 1; rep_slot[pc_off[pc] + n_reps++] = i; }
        }
        n_reps_pc[pc] = n_reps;
    });
    std::vector<uint64_t> cluster_base(n_pc + 1, 0);  // clusters before precluster pc
    for (size_t pc = 0; pc < n_pc; pc++) cluster_base[pc + 1] = cluster_base[pc] + n_reps_pc[pc];

    if (prefetch_reverse && !skip_clusterer && !lazy) {
        // what the membership sweep will ask for with the representative BEHIND the genome
        std::vector<size_t> want;
        for (uint32_t i = 0; i < n; i++) {
            if (is_rep[i]) continue;
            for (uint64_t x = adj.off[i]; x < adj.off[i + 1]; x++) {
                const uint32_t r = adj.nbr[x];
                if (r > i && is_rep[r] && !edge_state[x]) want.push_back(adj.hit[x]);
            }
        }
        if (!want.empty() && (*prefetch_reverse)(want)) { err = "reverse-orientation ANI batch failed"; return 2; }
    }

    // ---- memberships, written straight into the precluster's range of the output
    out.members.assign(n, 0);
    out.offsets.assign(cluster_base[n_pc] + 1, 0);
    std::atomic<int64_t> orphan{-1};
    sweep([&](size_t pc, std::vector<Cand> &, std::vector<std::pair<uint32_t, uint32_t>> &joins, std::vector<uint64_t> &cl_fill,
              uint64_t &calls) {
        const uint32_t *mb = pc_members.data() + pc_off[pc], *me = pc_members.data() + pc_off[pc + 1];
        const uint32_t *reps = rep_slot.data() + pc_off[pc];
        const size_t n_reps = n_reps_pc[pc];
        joins.clear();
        for (size_t c = 0; c < n_reps; c++) cluster_of_rep[reps[c]] = (uint32_t)c;
        for (const uint32_t *ip = mb; ip != me; ip++) {
            const uint32_t i = *ip;
            if (is_rep[i]) continue;
            bool have_best = false;
            float best = 0.f;
            uint32_t best_rep = 0;
            for (uint64_t x = adj.off[i]; x < adj.off[i + 1]; x++) {  // ascending representative index
                const uint32_t r = adj.nbr[x];
                if (!is_rep[r]) continue;
                OptAni v;
                if (skip_clusterer) {
                    v = OptAni{true, adj.ani[x]};
                } else {
                    if (edge_state[x]) v = OptAni{edge_state[x] == 1, edge_ani[x]};
                    else if (lazy) { v = OptAni{false, 0.f}; }  // cannot happen: the waves asked for every such pair
                    else {
                        float ani = 0.f;
                        const bool some = by_hit ? (*by_hit)(r, i, adj.hit[x], &ani) : calculate_ani_fn(r, i, &ani);
                        calls++;
                        v = OptAni{some, ani};
                        edge_state[x] = some ? 1 : 2; edge_ani[x] = ani;
                    }
                }
                if (v.some && (!have_best || v.ani > best)) { have_best = true; best = v.ani; best_rep = r; }
            }
            if (!have_best) {  // the reference panics here; remember the first such genome
                int64_t none = -1;
                orphan.compare_exchange_strong(none, (int64_t)i);
                return;
            }
            joins.emplace_back(cluster_of_rep[best_rep], i);
        }
        // clusters of this precluster in representative order: representative first, then its
        // members in ascending genome order (the order they were assigned in)
        cl_fill.assign(n_reps + 1, 0);
        for (const auto &jn : joins) cl_fill[jn.first + 1]++;
        const uint64_t base = pc_off[pc];
        for (size_t c = 0; c < n_reps; c++) cl_fill[c + 1] += cl_fill[c] + 1;  // +1: the representative
        for (size_t c = 0; c < n_reps; c++) {
            const uint64_t at = base + (c ? cl_fill[c] : 0);
            out.members[at] = reps[c];
            out.offsets[cluster_base[pc] + c + 1] = base + cl_fill[c + 1];
        }
        for (size_t c = n_reps; c-- > 0;) cl_fill[c + 1] = c ? cl_fill[c] + 1 : 1;  // next free slot after each representative
        for (const auto &jn : joins) out.members[base + cl_fill[jn.first + 1]++] = jn.second;
    });
    out.ani_calls = ani_calls.load();
    if (orphan.load() >= 0) {
        err = "called `Option::unwrap()` on a `None` value (genome " + std::to_string(orphan.load()) +
              " has no representative with an ANI; src/clusterer.rs:444)";
        return 1;
    }
    if (dbg) fprintf(stderr, "[engine] adjacency %.2f ms, preclusters %.2f ms, greedy %.2f ms\n", t1 - t0, t2 - t1, now() - t2);
    return 0;
}

}  // namespace gb200
