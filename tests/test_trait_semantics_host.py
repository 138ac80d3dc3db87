"""The parts of the reference's plugin interface that need no device: method names,
FinchPreclusterer's panics and its empty cache for contigs (src/finch.rs:12-46), `finch` + contigs in cluster()
(src/clusterer.rs:39-41), SkaniClusterer::initialise's assertion (src/skani.rs:696-698) -- and that everything that
computes refuses to run without a bound B200 instead of falling back to the CPU."""
import pytest

import galah_b200 as gb


def _no_device():
    import torch
    return not torch.cuda.is_available()


def test_method_names_and_thresholds():
    s = gb.Session()
    assert gb.FinchPreclusterer(0.9, session=s).method_name() == "finch"       # src/finch.rs:43-45
    assert gb.SkaniPreclusterer(90.0, session=s).method_name() == "skani"      # src/skani.rs:71-73
    cl = gb.SkaniClusterer(95.0, session=s)
    assert cl.method_name() == "skani" and cl.get_ani_threshold() == 95.0      # src/skani.rs:700-706
    cl.initialise()
    with pytest.raises(gb.GalahB200Error, match="threshold > 1.0"):            # a fraction is refused, src/skani.rs:696-698
        gb.SkaniClusterer(0.95, session=s).initialise()
    s.close()


def test_finch_preclusterer_panics_and_empty_contig_cache():
    s = gb.Session()
    with pytest.raises(gb.GalahB200Error, match="Low-memory clustering currently only supported with skani preclusterer"):
        gb.FinchPreclusterer(0.9, low_memory=True, session=s).distances(["a.fna", "b.fna"])          # src/finch.rs:14-15
    with pytest.raises(gb.GalahB200Error, match="Reference genome clustering currently only supported with skani preclusterer"):
        gb.FinchPreclusterer(0.9, session=s).distances_with_references(["a.fna", "b.fna"], ["a.fna"])  # src/finch.rs:40
    assert len(gb.FinchPreclusterer(0.9, session=s).distances_contigs(["a.fna"], ["c1", "c2"])) == 0  # src/finch.rs:26-33
    with pytest.raises(gb.GalahB200Error, match="finch does not support contig comparisons"):         # src/clusterer.rs:39-41
        gb.cluster_with(["a.fna"], gb.FinchPreclusterer(0.9, session=s), gb.SkaniClusterer(95.0, session=s),
                        cluster_contigs=True, contig_names=["c1"])
    s.close()


def test_compute_entries_refuse_to_run_without_a_device(tmp_path):
    if not _no_device():
        pytest.skip("a device is present")
    import numpy as np
    p = tmp_path / "g.fna"
    p.write_bytes(b">c\n" + b"ACGT" * 2000 + b"\n")
    s = gb.Session()
    calls = [lambda: gb.sketch_files([str(p)], 21, 1000),
             lambda: gb.finch_distances([str(p), str(p)], 0.9, 1000, 21),
             lambda: gb.cluster([str(p), str(p)]),
             lambda: gb.cluster_skani([str(p), str(p)]),
             lambda: gb.skani_distances([str(p), str(p)], 95.0),
             lambda: gb.SkaniPreclusterer(95.0, session=s).distances([str(p), str(p)]),
             lambda: gb.prefilter(np.zeros((2, 1000), np.uint64), np.zeros(2, np.uint32), 21, 0.9)]
    for call in calls:
        with pytest.raises(gb.GalahB200Error, match="no CPU fallback|no device|NO_DEVICE|init"):
            call()
    s.close()
