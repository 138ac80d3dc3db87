"""galah_b200_contig_names (host only): the record names `galah cluster --cluster-contigs` collects
(src/cluster_argument_parsing.rs:596-629) -- header text up to the first TAB, file order then record order, plain and
gzip, FASTQ headers too; a duplicate name fails with the reference's panic text -- on the committed copies of the
reference's contig fixtures and on synthetic files."""
import gzip
import os

import pytest

import galah_b200 as gb
from conftest import GOLDEN


def test_fixture_names_are_the_reference_cli_tests_names():
    """tests/test_cmdline.rs:460-507, 546-609 print exactly these names."""
    c = os.path.join(GOLDEN, "contigs")
    assert gb.contig_names([os.path.join(c, "contigs.fna.gz")]) == [
        "73.20110600_S2D.10_contig_13024", "73.20110600_S2D.10_contig_13024_2", "73.20110600_S2D.10_contig_50844",
        "73.20110600_S2D.10_contig_37820"]
    both = gb.contig_names([os.path.join(c, "contigs.fna.gz"), os.path.join(c, "contigs_extra.fna.gz")])
    assert both[4] == "73.20110600_S2D.10_contig_13024_3" and len(both) == 5
    assert gb.contig_names([os.path.join(c, "contigs_rep_bug.fna.gz")]) == [
        "k141_313035 flag=1 multi=13.9893 len=27966", "k141_401621 flag=1 multi=12.7497 len=42088",
        "NODE_1070_length_34582_cov_11.872969"]  # spaces stay: only a TAB ends the name
    assert gb.contig_names([os.path.join(c, "contigs_specific.fna.gz")])[5:7] == ["96ANI_80refAF_80queryAF", "94ANI_80refAF_80queryAF"]


def test_tab_cut_crlf_gzip_fastq_and_duplicates(tmp_path):
    a = tmp_path / "a.fna"
    a.write_bytes(b">c1 first\tdropped column\r\nACGT\r\n>c2\nAC\nGT\n>\nAC\n")
    b = tmp_path / "b.fna.gz"
    with gzip.open(b, "wb") as f:
        f.write(b">c3\nACGT\n")
    q = tmp_path / "r.fq"
    q.write_bytes(b"@read1 x\tcol\nACGT\n+\n>>>>\n@read2\nAC\n+\n@I\n")
    assert gb.contig_names([str(a), str(b), str(q)]) == ["c1 first", "c2", "", "c3", "read1 x", "read2"]
    assert gb.contig_names([]) == []
    dup = tmp_path / "dup.fna"
    dup.write_bytes(b">same\nAC\n>other\nAC\n>same\tagain\nAC\n")
    with pytest.raises(gb.GalahB200Error) as e:
        gb.contig_names([str(a), str(dup)])
    assert "Duplicate contig name found in file" in str(e.value) and "same" in str(e.value)
    with pytest.raises(gb.GalahB200Error) as e:  # the same name in two files is a duplicate too
        gb.contig_names([str(b), str(b)])
    assert "Duplicate contig name" in str(e.value)
    with pytest.raises(gb.GalahB200Error):
        gb.contig_names([str(tmp_path / "missing.fna")])
