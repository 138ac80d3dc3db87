"""Generates the committed golden vectors from the reference's own fixtures.

Run in the build container (needs /root/reference, which does NOT exist on the GPU box):
    python tests/golden/make_golden.py

Outputs (all small, committed):
  fixture_sketches.npz   oracle sketches (k=21, s=1000, seed 0) of the reference's test genomes
                         + the oracle's stage-1 pair list at min_ani 0.9 over all of them
  set1_1mbp.fna.gz, set1_500kb.fna.gz
                         the two inputs of the reference's only numeric known-answer test
                         (/root/reference/src/finch.rs:107-129 -> Some(0.9808188)), recompressed
                         so the GPU box can run that KAT end to end through the CUDA path
  contigs/*.fna.gz, set2_1mbp.half_aligned.fna.gz, antonio_mags/*.fna.gz (+ abisko4/*.fna.gz)
                         every fixture genome behind a stage-2 outcome the reference's tests pin
                         (SURVEY.md 8c), recompressed, so those outcomes run on the GPU box
                         (set2/1mbp.fna is byte-identical to set1/1mbp.fna: set1_1mbp.fna.gz)
  stage2_golden.json     the stage-2 oracle's integers and ANI for every pair those outcomes read
                         (oracle/skani_oracle.c; PARITY UNPINNED vs skani itself, see its header)
The oracle itself is pinned against that KAT in tests/test_oracle_golden.py.
"""
import gzip
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
import oracle  # noqa: E402

REF = "/root/reference/tests/data"
FIXTURES = [
    "set1/1mbp.fna", "set1/500kb.fna", "set2/1mbp.fna", "set2/1mbp.half_aligned.fna",
    "set1_name_clash/500kb.fna",
    "antonio_mags/BE_RX_R2_MAG52.fna", "antonio_mags/BE_RX_R3_MAG189.fna",
    "abisko4/73.20120800_S1X.13.fna", "abisko4/73.20120600_S2D.19.fna",
    "abisko4/73.20120700_S3X.12.fna", "abisko4/73.20110800_S2D.13.fna",
    "abisko4/73.20120800_S1D.21.fna", "abisko4/73.20110800_S2M.16.fna",
    "abisko4/73.20110800_S2M.16.fna.gz", "abisko4/73.20120800_S1D.21.fna.gz",
    "abisko4/73.20110600_S2D.10.fna", "abisko4/73.20110600_S3M.17.fna",
    "abisko4/73.20110700_S2D.12.fna", "abisko4/73.20110700_S2M.14.fna",
    "abisko4/73.20110800_S1D.9.fna", "abisko4/73.20110800_S3D.14.fna",
    "abisko4/73.20120600_E3D.30.fna", "abisko4/73.20120700_S1D.20.fna",
    "abisko4/73.20120700_S1X.9.fna", "abisko4/73.20120700_S2X.9.fna",
    "abisko4/73.20120700_S3D.12.fna", "abisko4/73.20120800_S2X.9.fna",
    "abisko_tabs/73.20120800_S1D.21.fna", "abisko_tabs/73.20110800_S2M.16.fna",
    "contigs/contigs.fna",
]


STAGE2_COPIES = [
    ("contigs/contigs.fna", "contigs/contigs.fna.gz"),
    ("contigs/contigs_extra.fna", "contigs/contigs_extra.fna.gz"),
    ("contigs/contigs_specific.fna", "contigs/contigs_specific.fna.gz"),
    ("contigs/contigs_rep_bug.fna", "contigs/contigs_rep_bug.fna.gz"),
    ("set2/1mbp.half_aligned.fna", "set2_1mbp.half_aligned.fna.gz"),
    ("antonio_mags/BE_RX_R2_MAG52.fna", "antonio_mags/BE_RX_R2_MAG52.fna.gz"),
    ("antonio_mags/BE_RX_R3_MAG189.fna", "antonio_mags/BE_RX_R3_MAG189.fna.gz"),
]


def stage2_golden():
    """Oracle outputs for the pairs behind the reference-pinned stage-2 outcomes."""
    import json
    sys.path.insert(0, os.path.join(HERE, ".."))
    from stage2_cases import CASES, units_of  # the same case list the tests read
    out = {}
    for name, case in CASES.items():
        units = units_of(HERE, case)
        c = 30 if case["small"] else 125
        gen = [oracle.AniGenome(*u, c=c) for u in units]
        rows = []
        for i in range(len(units)):
            for j in range(i + 1, len(units)):
                ints = oracle.ani_pair_integers(gen[i], gen[j])
                ani, afq, afr, un, est = oracle.ani_finish(ints, case["min_af"], c, case["contigs"])
                rows.append({"i": i, "j": j, "sum_fx": ints[0], "n_chunks": ints[1], "cov_q": ints[2],
                             "cov_r": ints[3], "sum_m": ints[6], "span_m": ints[7], "span_n": ints[8],
                             "n_chains": ints[9], "estimator": est, "ani": float(ani),
                             "ani_unrounded": round(un, 6), "af_q": round(afq, 6), "af_r": round(afr, 6)})
        out[name] = rows
    with open(os.path.join(HERE, "stage2_golden.json"), "w") as f:
        json.dump(out, f, indent=1)
    return sum(len(v) for v in out.values())


def main():
    for src, dst in STAGE2_COPIES:
        os.makedirs(os.path.dirname(os.path.join(HERE, dst)), exist_ok=True)
        with open(os.path.join(REF, src), "rb") as f, gzip.GzipFile(
                os.path.join(HERE, dst), "wb", compresslevel=9, mtime=0) as g:
            g.write(f.read())
    print("stage-2 golden rows:", stage2_golden())
    sketches = [oracle.sketch_fasta(os.path.join(REF, f)) for f in FIXTURES]
    table, counts = oracle.pack_table(sketches, 1000)
    pairs = oracle.prefilter(table, counts, 21, 0.9)
    np.savez_compressed(os.path.join(HERE, "fixture_sketches.npz"), names=np.array(FIXTURES),
                        table=table, counts=counts, pairs_min_ani_0p9=pairs)
    for src, dst in [("set1/1mbp.fna", "set1_1mbp.fna.gz"), ("set1/500kb.fna", "set1_500kb.fna.gz")]:
        with open(os.path.join(REF, src), "rb") as f, gzip.GzipFile(
                os.path.join(HERE, dst), "wb", compresslevel=9, mtime=0) as g:
            g.write(f.read())
    print(f"{len(FIXTURES)} sketches, {len(pairs)} passing pairs at 0.9; counts min {counts.min()}")


if __name__ == "__main__":
    main()
