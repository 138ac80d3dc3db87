"""galah-b200, the `galah cluster` command line over the C ABI (galah_b200/csrc/cli/main.cpp): the reference's
argument rules (src/cluster_argument_parsing.rs:570-673, 1660-1757; panics of src/finch.rs:15,40 and
src/clusterer.rs:39-41) without a device, and -- on the GPU -- the reference's own command-line tests
(tests/test_cmdline.rs) replayed on the committed copies of its fixtures: same flags, same stdout."""
import os
import subprocess

import pytest

from conftest import GOLDEN, ROOT

CLI = os.path.join(ROOT, "galah_b200", "galah-b200")


def run(*args):
    from galah_b200 import build
    build.build()
    return subprocess.run([CLI, *args], capture_output=True, text=True, timeout=600)


def g(name):
    return os.path.join(GOLDEN, name)


# ---------------------------------------------------------------- argument rules (no device needed)
def test_an_output_flag_is_required():
    r = run("cluster", "-f", "a.fna")  # cluster_argument_parsing.rs:1715-1751
    assert r.returncode == 2 and "required arguments were not provided" in r.stderr


def test_contig_flags():
    r = run("cluster", "-f", "a.fna", "--cluster-contigs", "-o", "/dev/null")  # tests/test_cmdline.rs:509-526
    assert r.returncode == 1 and "either --small-contigs or --large-contigs must be specified" in r.stderr
    r = run("cluster", "-f", "a.fna", "--small-contigs", "-o", "/dev/null")  # clap `requires`
    assert r.returncode == 2 and "--cluster-contigs" in r.stderr
    r = run("cluster", "-f", "a.fna", "--cluster-contigs", "--small-contigs", "--large-contigs", "-o", "/dev/null")
    assert r.returncode == 2 and "cannot be used with" in r.stderr
    r = run("cluster", "-f", "a.fna", "--cluster-contigs", "--small-contigs", "--output-representative-fasta-directory", "/tmp/x")
    assert r.returncode == 1 and "Cannot specify --cluster-contigs with --output-representative-fasta-directory" in r.stderr


def test_reference_panics_keep_their_text():
    r = run("cluster", "-f", "a.fna", "--precluster-method", "finch", "--cluster-contigs", "--small-contigs", "-o", "/dev/null")
    assert r.returncode == 1 and "finch does not support contig comparisons." in r.stderr  # src/clusterer.rs:39-41
    r = run("cluster", "-f", "a.fna", "--precluster-method", "finch", "--low-memory", "-o", "/dev/null")
    assert r.returncode == 1 and "Low-memory clustering currently only supported with skani preclusterer" in r.stderr
    r = run("cluster", "-f", "a.fna", "--precluster-method", "finch", "--reference-genomes", "r.fna", "-o", "/dev/null")
    assert r.returncode == 1 and "Reference genome clustering currently only supported with skani preclusterer" in r.stderr
    r = run("cluster", "-f", "a.fna", "--reference-genomes", "r.fna", "--cluster-contigs", "--small-contigs", "-o", "/dev/null")
    assert r.returncode == 1 and "Reference genome clustering is not currently supported with --cluster-contigs" in r.stderr
    r = run("cluster", "-f", "a.fna", "--low-memory", "--reference-genomes", "r.fna", "-o", "/dev/null")
    assert r.returncode == 2 and "cannot be used with" in r.stderr


def test_percentages_and_methods():
    r = run("cluster", "-f", "a.fna", "--ani", "150", "-o", "/dev/null")  # parse_percentage, :1491-1512
    assert r.returncode == 1 and "Invalid percentage specified for --ani: '150'" in r.stderr
    r = run("cluster", "-f", "a.fna", "--precluster-method", "dashing", "-o", "/dev/null")
    assert r.returncode == 2 and "possible values: skani, finch" in r.stderr
    r = run("cluster", "-f", "a.fna", "--checkm-tab-table", "x", "-o", "/dev/null")
    assert r.returncode == 2 and "quality ordering" in r.stderr
    r = run("--version")
    assert r.returncode == 0 and "galah-b200" in r.stdout


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = run("cluster", "-f", g("set1_1mbp.fna.gz"), g("set1_500kb.fna.gz"), "-o", "/dev/null")
    assert r.returncode == 1 and "no CPU fallback" in r.stderr


# ---------------------------------------------------------------- the reference's CLI tests on the GPU
C13024, C50844, C37820 = ("73.20110600_S2D.10_contig_13024", "73.20110600_S2D.10_contig_50844",
                          "73.20110600_S2D.10_contig_37820")
K313, K401, NODE = ("k141_313035 flag=1 multi=13.9893 len=27966", "k141_401621 flag=1 multi=12.7497 len=42088",
                    "NODE_1070_length_34582_cov_11.872969")


def lines(*pairs):
    return "".join(f"{a}\t{b}\n" for a, b in pairs)


@pytest.mark.gpu
@pytest.mark.parametrize("files,flags,expected", [
    # tests/test_cmdline.rs:460-480
    (["contigs/contigs.fna.gz"], ["--cluster-contigs", "--large-contigs"],
     lines((C13024, C13024), (C13024, C13024 + "_2"), (C50844, C50844), (C37820, C37820))),
    # :482-507
    (["contigs/contigs_specific.fna.gz"], ["--cluster-contigs", "--small-contigs"],
     lines((C13024, C13024), (C13024, "100ANI_100AF"), (C13024, "100ANI_100refAF_90queryAF"), (C13024, "100ANI_90refAF_90queryAF"),
           (C13024, "100ANI_80refAF_80queryAF"), (C13024, "96ANI_80refAF_80queryAF"),
           ("94ANI_80refAF_80queryAF", "94ANI_80refAF_80queryAF"), (C50844, C50844), (C37820, C37820))),
    # :546-567
    (["contigs/contigs.fna.gz", "contigs/contigs_extra.fna.gz"], ["--cluster-contigs", "--small-contigs"],
     lines((C13024, C13024), (C13024, C13024 + "_2"), (C13024, C13024 + "_3"), (C50844, C50844), (C37820, C37820))),
    # :569-588, :590-609
    (["contigs/contigs_rep_bug.fna.gz"], ["--cluster-contigs", "--large-contigs"], lines((K313, K313), (K313, K401), (K313, NODE))),
    (["contigs/contigs_rep_bug.fna.gz"], ["--cluster-contigs", "--small-contigs"], lines((K313, K313), (K313, K401), (NODE, NODE))),
])
def test_contig_cluster_definitions(gb, files, flags, expected):
    r = run("cluster", "--genome-fasta-files", *[g(f) for f in files], *flags, "--output-cluster-definition", "/dev/stdout", "-q")
    assert r.returncode == 0, r.stderr
    assert r.stdout == expected


@pytest.mark.gpu
def test_min_aligned_fraction(gb):
    """tests/test_cmdline.rs:262-302 (set1/1mbp.fna is byte-identical to set2/1mbp.fna in its sequence)."""
    a, b = g("set1_1mbp.fna.gz"), g("set2_1mbp.half_aligned.fna.gz")
    common = ["cluster", "--genome-fasta-files", a, b, "--precluster-method", "finch", "--output-representative-list", "/dev/stdout", "-q"]
    r = run(*common, "--min-aligned-fraction", "0.2")
    assert r.returncode == 0 and r.stdout == a + "\n", r.stderr
    r = run(*common, "--min-aligned-fraction", "0.6")
    assert r.returncode == 0 and r.stdout == a + "\n" + b + "\n", r.stderr


@pytest.mark.gpu
def test_github7(gb):
    """tests/test_cmdline.rs:417-440."""
    a, b = g("antonio_mags/BE_RX_R2_MAG52.fna.gz"), g("antonio_mags/BE_RX_R3_MAG189.fna.gz")
    r = run("cluster", "--genome-fasta-files", a, b, "--precluster-method", "finch", "--precluster-ani", "90", "--ani", "95",
            "--min-aligned-fraction", "60", "--output-representative-list", "/dev/stdout", "-q")
    assert r.returncode == 0 and r.stdout == a + "\n", r.stderr


@pytest.mark.gpu
def test_cluster_definition_rep_directories_and_low_memory(gb, tmp_path):
    """abisko4 at 99 % (src/clusterer.rs:662-690 through the CLI flags): cluster file, representative list, symlinked and
    copied representatives; the default skani + skani path, --low-memory and a genome list file give the same clusters."""
    from stage2_cases import AB
    paths = [g(f) for f in AB]
    out = tmp_path / "clusters.tsv"
    reps = tmp_path / "reps.txt"
    r = run("cluster", "-f", *paths, "--precluster-method", "finch", "--ani", "99", "--min-aligned-fraction", "20",
            "-o", str(out), "--output-representative-list", str(reps),
            "--output-representative-fasta-directory", str(tmp_path / "links"),
            "--output-representative-fasta-directory-copy", str(tmp_path / "copies"))
    assert r.returncode == 0, r.stderr
    assert "Found 1 preclusters. The largest contained 4 genomes" in r.stderr and "Found 2 genome clusters" in r.stderr
    want = lines((paths[0], paths[0]), (paths[0], paths[1]), (paths[0], paths[3]), (paths[2], paths[2]))
    assert out.read_text() == want
    assert reps.read_text() == paths[0] + "\n" + paths[2] + "\n"
    for d in ("links", "copies"):
        assert sorted(os.listdir(tmp_path / d)) == sorted(os.path.basename(paths[x]) for x in (0, 2))
    assert os.path.islink(tmp_path / "links" / os.path.basename(paths[0]))
    assert (tmp_path / "copies" / os.path.basename(paths[2])).read_bytes() == open(paths[2], "rb").read()
    # an existing, non-empty output directory is refused (cluster_argument_parsing.rs:789-797)
    r = run("cluster", "-f", *paths, "--output-representative-fasta-directory", str(tmp_path / "links"))
    assert r.returncode == 1 and "exists and is not empty" in r.stderr
    listing = tmp_path / "genomes.txt"
    listing.write_text("".join(p + "\tignored column\n" for p in paths))
    outs = []
    for extra in ([], ["--low-memory"]):
        r = run("cluster", "--genome-fasta-list", str(listing), "--ani", "99", "--min-aligned-fraction", "20", "-o", "/dev/stdout", "-q",
                *extra)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert outs[0] == outs[1] == want  # src/clusterer.rs:692-723, 759-791


@pytest.mark.gpu
def test_reference_genomes_and_directory_input(gb, tmp_path):
    """--reference-genomes (src/skani.rs:502-687 through the CLI: references first in the combined list, clusters only
    across the two groups) and -d / -x directory input (sorted listing)."""
    import shutil
    from stage2_cases import AB
    p = [g(f) for f in AB]
    r = run("cluster", "-f", p[1], p[3], p[2], "--reference-genomes", p[0], "--ani", "99", "--min-aligned-fraction", "20",
            "-o", "/dev/stdout", "-q")
    assert r.returncode == 0, r.stderr
    assert r.stdout == lines((p[0], p[0]), (p[0], p[1]), (p[0], p[3]), (p[2], p[2]))
    listing = tmp_path / "refs.txt"
    listing.write_text(p[0] + "\n\n")
    r2 = run("cluster", "-f", p[1], p[3], p[2], "--reference-genomes-list", str(listing), "--ani", "99", "--min-aligned-fraction", "20",
             "-o", "/dev/stdout", "-q")
    assert r2.returncode == 0 and r2.stdout == r.stdout, r2.stderr
    d = tmp_path / "genomes"
    d.mkdir()
    for f in (p[0], p[1]):
        shutil.copy(f, d / os.path.basename(f))
    (d / "notes.txt").write_text("not a genome")
    r = run("cluster", "-d", str(d), "-x", "gz", "--ani", "99", "--min-aligned-fraction", "20", "-o", "/dev/stdout", "-q")
    assert r.returncode == 0, r.stderr
    first, second = sorted(str(d / os.path.basename(f)) for f in (p[0], p[1]))
    assert r.stdout == lines((first, first), (first, second))
    r = run("cluster", "-d", str(d), "-x", "fasta", "-o", "/dev/stdout", "-q")
    assert r.returncode == 1 and "Found 0 genomes" in r.stderr


def test_percentages_follow_the_references_f32_arithmetic():
    """parse_percentage (src/cluster_argument_parsing.rs:1491-1512) divides by 100 in f32 and the backends multiply by 100
    again (:1300-1352, 1478-1485; src/skani.rs:153, 742): --min-aligned-fraction 15 reaches skani as 15.000001, 60 as
    60.000004, 0.2 as 20; --ani 95 as 95; finch's min_ani 90 as 0.9f.  skip_clusterer: same method names, or contigs."""
    import numpy as np

    def config(*args):
        r = run("cluster", "-f", "a.fna", "b.fna", "--print-config", *args)
        assert r.returncode == 0, r.stderr
        return dict(line.split("\t") for line in r.stdout.strip().split("\n"))
    c = config()
    assert np.float32(c["min_af_percent"]) == np.float32(0.15) * np.float32(100.0) == np.float32(15.000001)
    assert float(c["ani_threshold"]) == 95.0 and float(c["skani_precluster_threshold"]) == 95.0  # skip_clusterer: --ani
    assert c["skip_clusterer"] == "1" and c["precluster_method"] == c["cluster_method"] == "skani"
    c = config("--min-aligned-fraction", "60", "--precluster-method", "finch", "--ani", "99")
    assert np.float32(c["min_af_percent"]) == np.float32(0.6) * np.float32(100.0) == np.float32(60.000004)
    assert np.float32(c["finch_min_ani"]) == np.float32(0.9) and c["skip_clusterer"] == "0"
    assert float(c["ani_threshold"]) == float(np.float32(0.99) * np.float32(100.0)) and float(c["skani_precluster_threshold"]) == 90.0
    c = config("--min-aligned-fraction", "0.2", "--precluster-ani", "0.85")
    assert float(c["min_af_percent"]) == 20.0
    c = config("--cluster-contigs", "--small-contigs", "--small-genomes")
    assert c["small_genomes"] == "1" and c["cluster_contigs"] == "1" and c["skip_clusterer"] == "1"
    c = config("--cluster-contigs", "--large-contigs", "--small-genomes")  # --small-genomes is ignored for contigs (:1760-1782)
    assert c["small_genomes"] == "0"
    c = config("--reference-genomes", "r1.fna", "r2.fna")
    assert c["genomes"] == "4" and c["references"] == "2"


def test_genome_input_flags_without_a_device(tmp_path):
    """-f, --genome-fasta-list (a path is the text before the first TAB, blank lines skipped) and -d with -x
    (docs/tools/cluster.md:40-62) add up; an empty selection is refused before any device is touched."""
    listing = tmp_path / "list.txt"
    listing.write_text("x/one.fna\tquality 98\n\n  \nx/two.fna\n")
    d = tmp_path / "dir"
    d.mkdir()
    for name in ("b.fa", "a.fa", "c.fna", "notes.txt"):
        (d / name).write_text(">c\nACGT\n")
    r = run("cluster", "-f", "z.fna", "--genome-fasta-list", str(listing), "-d", str(d), "-x", "fa", "--print-config")
    assert r.returncode == 0, r.stderr
    assert dict(line.split("\t") for line in r.stdout.strip().split("\n"))["genomes"] == "5"
    r = run("cluster", "-d", str(d), "-x", "fasta", "-o", "/dev/null")
    assert r.returncode == 1 and "Found 0 genomes" in r.stderr
    r = run("cluster", "--genome-fasta-list", str(tmp_path / "missing.txt"), "-o", "/dev/null")
    assert r.returncode == 1 and "Failed to read genome fasta list" in r.stderr
    r = run("cluster", "-o", "/dev/null")
    assert r.returncode == 1 and "No genome fasta files found" in r.stderr
    r = run("derep", "-f", "a.fna")
    assert r.returncode == 2 and "unrecognized subcommand" in r.stderr
