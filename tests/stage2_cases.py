"""The stage-2 outcomes the reference's own tests pin, as data: which committed
fixture files, which mode, which thresholds, which clusters.  Read by tests/golden/make_golden.py
(golden values), tests/test_stage2_fixtures.py (oracle, CPU) and tests/test_stage2_fixtures_gpu.py
(CUDA path through the C ABI).  Cluster lists are sorted members, clusters sorted -- the reference's
library tests sort the same way (src/clusterer.rs:716-722); `order` (optional) is the exact
representative-first output of the CLI tests."""
import os

AB = ["abisko4/73.20120800_S1X.13.fna.gz", "abisko4/73.20120600_S2D.19.fna.gz",
      "abisko4/73.20120700_S3X.12.fna.gz", "abisko4/73.20110800_S2D.13.fna.gz"]
MAG52, MAG189 = "antonio_mags/BE_RX_R2_MAG52.fna.gz", "antonio_mags/BE_RX_R3_MAG189.fna.gz"

# parse_percentage (src/cluster_argument_parsing.rs:1491-1512) divides by 100 in f32 and the skani
# callers multiply by 100 again (src/skani.rs:153, 742): 15 -> 15.000001, 60 -> 60.000004, 20 -> 20
CLI_AF_DEFAULT = 15.000001

CASES = {
    # ---- src/clusterer.rs library tests (preclusterer threshold 90, clusterer 99, min_af 0.2)
    "skani_skani_two_clusters_same_ani": dict(  # src/clusterer.rs:692-723
        files=AB, contigs=False, small=False, pre="skani", pre_thr=90.0, ani=99.0, min_af=20.0,
        clusters=[[0, 1, 3], [2]]),
    "skani_skani_two_preclusters": dict(  # src/clusterer.rs:725-757
        files=AB + [MAG52], contigs=False, small=False, pre="skani", pre_thr=90.0, ani=99.0, min_af=20.0,
        clusters=[[0, 1, 3], [2], [4]]),
    "skani_skani_low_memory": dict(  # src/clusterer.rs:759-791 (low_memory: true)
        files=AB + [MAG52], contigs=False, small=False, pre="skani", pre_thr=90.0, ani=99.0, min_af=20.0,
        low_memory=True, clusters=[[0, 1, 3], [2], [4]]),
    "lib_contig_cluster": dict(  # src/clusterer.rs:793-823
        files=["contigs/contigs.fna.gz"], contigs=True, small=False, pre="skani", pre_thr=90.0, ani=99.0,
        min_af=20.0, clusters=[[0, 1], [2], [3]]),
    "finch_skani_abisko4_95": dict(  # src/clusterer.rs:631-660 (finch 0.9 + skani 95)
        files=AB, contigs=False, small=False, pre="finch", pre_thr=0.9, ani=95.0, min_af=20.0,
        clusters=[[0, 1, 2, 3]]),
    "finch_skani_abisko4_99": dict(  # src/clusterer.rs:662-690
        files=AB, contigs=False, small=False, pre="finch", pre_thr=0.9, ani=99.0, min_af=20.0,
        clusters=[[0, 1, 3], [2]]),
    # ---- tests/test_cmdline.rs (CLI defaults: skani + skani, ANI 95, min-AF 15)
    "cli_min_aligned_fraction_0.2": dict(  # tests/test_cmdline.rs:262-280 (finch precluster)
        files=["set1_1mbp.fna.gz", "set2_1mbp.half_aligned.fna.gz"], contigs=False, small=False, pre="finch",
        pre_thr=0.9, ani=95.0, min_af=20.0, clusters=[[0, 1]]),
    "cli_min_aligned_fraction_0.6": dict(  # tests/test_cmdline.rs:282-302
        files=["set1_1mbp.fna.gz", "set2_1mbp.half_aligned.fna.gz"], contigs=False, small=False, pre="finch",
        pre_thr=0.9, ani=95.0, min_af=60.000004, clusters=[[0], [1]]),
    "cli_antonio_af60": dict(  # tests/test_cmdline.rs:417-440: one representative, MAG52
        files=[MAG52, MAG189], contigs=False, small=False, pre="skani", pre_thr=95.0, ani=95.0,
        min_af=60.000004, clusters=[[0, 1]], order=[[0, 1]]),
    "cli_contig_cluster_large": dict(  # tests/test_cmdline.rs:460-480
        files=["contigs/contigs.fna.gz"], contigs=True, small=False, pre="skani", pre_thr=95.0, ani=95.0,
        min_af=CLI_AF_DEFAULT, clusters=[[0, 1], [2], [3]], order=[[0, 1], [2], [3]]),
    "cli_contig_cluster_specific_small": dict(  # tests/test_cmdline.rs:482-507
        files=["contigs/contigs_specific.fna.gz"], contigs=True, small=True, pre="skani", pre_thr=95.0, ani=95.0,
        min_af=CLI_AF_DEFAULT, clusters=[[0, 1, 2, 3, 4, 5], [6], [7], [8]],
        order=[[0, 1, 2, 3, 4, 5], [6], [7], [8]]),
    "cli_contig_cluster_multiple_files_small": dict(  # tests/test_cmdline.rs:546-567
        files=["contigs/contigs.fna.gz", "contigs/contigs_extra.fna.gz"], contigs=True, small=True, pre="skani",
        pre_thr=95.0, ani=95.0, min_af=CLI_AF_DEFAULT, clusters=[[0, 1, 4], [2], [3]],
        order=[[0, 1, 4], [2], [3]]),
    "cli_contig_rep_bug_large": dict(  # tests/test_cmdline.rs:569-588
        files=["contigs/contigs_rep_bug.fna.gz"], contigs=True, small=False, pre="skani", pre_thr=95.0, ani=95.0,
        min_af=CLI_AF_DEFAULT, clusters=[[0, 1, 2]], order=[[0, 1, 2]]),
    "cli_contig_rep_bug_small": dict(  # tests/test_cmdline.rs:590-609
        files=["contigs/contigs_rep_bug.fna.gz"], contigs=True, small=True, pre="skani", pre_thr=95.0, ani=95.0,
        min_af=CLI_AF_DEFAULT, clusters=[[0, 1], [2]], order=[[0, 1], [2]]),
    "cli_small_genomes_pair": dict(  # tests/test_cmdline.rs:442-458 (stdout contains S1X.13: it is a rep)
        files=AB[:2], contigs=False, small=True, pre="skani", pre_thr=95.0, ani=95.0, min_af=CLI_AF_DEFAULT,
        clusters=None),
}


def paths_of(golden_dir, case):
    return [os.path.join(golden_dir, f) for f in case["files"]]


def units_of(golden_dir, case):
    """Oracle units (codes, rec_start, rec_end): one per file, or one per record in contig mode."""
    import oracle
    units = []
    for p in paths_of(golden_dir, case):
        codes, rs, re_ = oracle.load_codes(p)
        if case["contigs"]:
            units += [(codes, rs[i:i + 1], re_[i:i + 1]) for i in range(len(rs))]
        else:
            units.append((codes, rs, re_))
    return units
