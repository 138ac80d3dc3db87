"""The host clustering engine under AddressSanitizer + UBSan and under ThreadSanitizer: tools/engine_fuzz.cpp runs the
same hit lists through the serial, table (threaded sweeps), wave and short-budget wave engines and compares clusters,
order and calculate_ani counts; the sanitizers watch the threaded adjacency fill and the threaded sweeps."""
import os
import shutil
import subprocess

import pytest

from conftest import ROOT

SRC = [os.path.join(ROOT, "tools", "engine_fuzz.cpp"), os.path.join(ROOT, "galah_b200", "csrc", "host", "cluster_engine.cpp")]


def _build(tmp_path, flags, name):
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path / name)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", *flags, "-pthread", "-I", os.path.join(ROOT, "galah_b200", "csrc"), *SRC,
                        "-o", exe], capture_output=True, text=True)
    if r.returncode and "sanitize" in r.stderr:
        pytest.skip("this g++ has no sanitizer runtime: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    return exe


def test_engine_under_asan_and_ubsan(tmp_path):
    exe = _build(tmp_path, ["-fsanitize=address,undefined", "-fno-omit-frame-pointer"], "fuzz_asan")
    for mode in ("0", "1", "2"):
        r = subprocess.run([exe, mode], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and f"mode {mode} ok" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])


def test_threaded_paths_under_tsan(tmp_path):
    exe = _build(tmp_path, ["-fsanitize=thread"], "fuzz_tsan")
    for mode in ("1", "2"):
        r = subprocess.run([exe, mode], capture_output=True, text=True, timeout=900)
        if "FATAL: ThreadSanitizer" in r.stderr and "unexpected memory mapping" in r.stderr:
            pytest.skip("ThreadSanitizer cannot run under this kernel's address-space layout")
        assert r.returncode == 0 and f"mode {mode} ok" in r.stdout and "WARNING: ThreadSanitizer" not in r.stderr, \
            (r.stdout[-500:], r.stderr[-3000:])


def test_host_fasta_packer_fuzz_under_asan(tmp_path):
    """tools/fasta_fuzz.cpp: 4,000 random FASTA / FASTQ / noise / truncated inputs through the host packer
    (csrc/host/fasta.cpp) under ASan + UBSan; accepted inputs must give self-consistent packed genomes."""
    if not shutil.which("g++"):
        pytest.skip("no g++")
    exe = str(tmp_path / "fasta_fuzz")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer", "-I",
                        os.path.join(ROOT, "galah_b200", "csrc"), os.path.join(ROOT, "tools", "fasta_fuzz.cpp"),
                        os.path.join(ROOT, "galah_b200", "csrc", "host", "fasta.cpp"), "-l:libz.so.1", "-o", exe],
                       capture_output=True, text=True)
    if r.returncode and ("sanitize" in r.stderr or "libz" in r.stderr):
        pytest.skip("no sanitizer runtime / zlib for a standalone build: " + r.stderr[-200:])
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "fasta fuzz ok" in r.stdout, (r.stdout[-500:], r.stderr[-3000:])
