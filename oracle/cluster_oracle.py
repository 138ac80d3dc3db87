"""oracle/cluster_oracle.py -- pure-Python restatement of galah::clusterer::cluster().

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Follows /root/reference/src/clusterer.rs line
by line, with the same data structures the reference uses (an order-normalised pair map for
SortedPairGenomeDistanceCache, src/sorted_pair_genome_distance_cache.rs:4-58; a BTreeSet of
representatives), at --threads 1 where every rayon construct runs in sequence order.  It is
quadratic exactly where the reference is quadratic, so use it on small cases only.

Parity pins available in the reference for this logic: src/sorted_pair_genome_distance_cache.rs
:69-114 (transform_ids) is reproduced in tests/test_cluster_engine.py; the cluster-output tests
(src/clusterer.rs:537-824) need a real skani/fastANI binary and cannot run here, so the engine is
pinned against THIS restatement on randomised inputs plus hand-worked cases.
"""
import numpy as np


class SortedPairGenomeDistanceCache:
    """src/sorted_pair_genome_distance_cache.rs:4-58"""

    def __init__(self):
        self.internal = {}

    def insert(self, ids, distance):
        a, b = ids
        self.internal[(a, b) if a < b else (b, a)] = distance

    def get(self, ids):
        """Returns (present, value) -- Option<&Option<f32>>."""
        a, b = ids
        k = (a, b) if a < b else (b, a)
        return (True, self.internal[k]) if k in self.internal else (False, None)

    def contains_key(self, ids):
        return self.get(ids)[0]

    def transform_ids(self, input_ids):
        out = SortedPairGenomeDistanceCache()
        for i, g1 in enumerate(input_ids):
            for j in range(i + 1, len(input_ids)):
                present, v = self.get((g1, input_ids[j]))
                if present:
                    out.insert((i, j), v)
        return out

    def clone(self):
        c = SortedPairGenomeDistanceCache()
        c.internal = dict(self.internal)
        return c


def partition_sketches(n, cache):
    """src/clusterer.rs:452-487 + `.indices().sets()` (sets by smallest member, members ascending)."""
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for i in range(n):
        for j in range(i):
            if cache.contains_key((i, j)):
                a, b = find(i), find(j)
                if a != b:
                    parent[max(a, b)] = min(a, b)
    sets, where = [], {}
    for g in range(n):
        r = find(g)
        if r not in where:
            where[r] = len(sets)
            sets.append([])
        sets[where[r]].append(g)
    return sets


def _f32(x):
    return None if x is None else np.float32(x)


def find_representatives(calculate_ani, threshold, pre, n, skip_clusterer, counter):
    """src/clusterer.rs:182-259"""
    reps = []  # ascending (BTreeSet)
    clusterer_cache = SortedPairGenomeDistanceCache()
    for i in range(n):
        cands = []
        for j in reps:
            present, v = pre.get((i, j))
            if present:
                cands.append((j, v))
        # sort_unstable_by partial_cmp on Option<f32>: None < Some; stable for short inputs
        cands.sort(key=lambda t: (t[1] is not None, t[1] if t[1] is not None else 0.0))
        potential_refs = [j for j, _ in cands]
        if skip_clusterer:
            anis = [pre.get((j, i))[1] for j in potential_refs]
            anis = [a for a in anis if a is not None]  # .flatten()
        else:
            anis = [None] * len(potential_refs)
            for x, j in enumerate(potential_refs):  # find_any, one thread
                a = _f32(calculate_ani(j, i))
                counter[0] += 1
                anis[x] = a
                if a is not None and a >= threshold:
                    break
        is_rep = True
        for j, a in zip(potential_refs, anis):
            if a is not None:
                if not skip_clusterer:
                    clusterer_cache.insert((j, i), a)
                if a >= threshold:
                    is_rep = False
        if is_rep:
            reps.append(i)
    return reps, (pre.clone() if skip_clusterer else clusterer_cache)


def find_memberships(calculate_ani, reps, pre, n, cache, counter):
    """src/clusterer.rs:350-449"""
    rep_to_index = {r: x for x, r in enumerate(reps)}
    out = [[r] for r in reps]
    for i in range(n):
        if i in rep_to_index:
            continue
        potential = [r for r in reps if not cache.contains_key((i, r)) and pre.contains_key((i, r))]
        for r in potential:
            cache.insert((i, r), _f32(calculate_ani(r, i)))
            counter[0] += 1
        best, best_rep = None, None
        for r in reps:
            present, v = cache.get((i, r))
            a = v if present else None
            if a is not None and (best is None or a > best):
                best, best_rep = a, r
        if best_rep is None:
            raise RuntimeError("called `Option::unwrap()` on a `None` value")
        out[rep_to_index[best_rep]].append(i)
    return out


def cluster(n, hits, threshold, calculate_ani=None, skip_clusterer=False):
    """src/clusterer.rs:14-152 after the preclusterer ran.  hits: iterable of (i, j, ani)."""
    threshold = np.float32(threshold)
    pre = SortedPairGenomeDistanceCache()
    for i, j, ani in hits:
        pre.insert((int(i), int(j)), np.float32(ani))
    preclusters = partition_sketches(n, pre)
    preclusters.sort(key=lambda c: -len(c))  # stable, see cluster_engine.cpp
    counter = [0]
    all_clusters = []
    for original in preclusters:
        sub = pre.transform_ids(original)
        reps, cache = find_representatives(calculate_ani and (lambda a, b: calculate_ani(original[a], original[b])),
                                           threshold, sub, len(original), skip_clusterer, counter)
        clusters = find_memberships(calculate_ani and (lambda a, b: calculate_ani(original[a], original[b])),
                                    reps, sub, len(original), cache, counter)
        for c in clusters:
            all_clusters.append([original[x] for x in c])
    return all_clusters, {"ani_calls": counter[0], "n_preclusters": len(preclusters),
                          "largest_precluster": len(preclusters[0])}
