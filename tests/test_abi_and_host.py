"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol declared in
include/gossamer_b200.h, fails loudly without a GPU (no CPU fallback), and the `goss` host CLI
keeps the reference's option / error / exit-code behaviour (src/App.cc:253-276,328-418)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import gossamer_b200 as G

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOSS = os.path.join(ROOT, "gossamer_b200", "goss")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_are_all_exported():
    hdr = open(os.path.join(ROOT, "include", "gossamer_b200.h")).read()
    declared = set(re.findall(r"\b(gsb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"gsb_log_fn"}
    assert len(declared) >= 20
    lib = G.lib()
    missing = [n for n in sorted(declared) if not hasattr(lib, n)]
    assert not missing, missing
    assert set(G.EXPORTS) <= declared


def test_library_has_no_torch_or_oracle_dependency():
    out = subprocess.run(["ldd", G.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out and "libcudart" in out


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "gossamer_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hh", ".cc")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_py" not in src and "goss_oracle" not in src and "liboracle" not in src, f


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(G.GossamerError) as e:
        G.Builder(G.GRAPH, 25)
    assert e.value.status == -5 and "no CPU path" in e.value.message


def test_k_range_is_checked_before_touching_the_device():
    with pytest.raises(G.GossamerError) as e:
        G.Builder(G.GRAPH, 63)                       # Graph::MaxK = 62, src/Graph.hh:89
    assert e.value.status == -7 and e.value.message == "unable to build a graph with k=63"
    with pytest.raises(G.GossamerError) as e:
        G.Builder(G.KMERSET, 64)                     # KmerSet::MaxK = 63, src/KmerSet.hh:30
    assert e.value.status == -7


def test_plan_splitters_quantiles():
    import numpy as np
    lo = np.arange(1000, dtype=np.uint64)[::-1].copy()
    slo, shi = G.plan_splitters(lo, np.zeros_like(lo), 4)
    assert list(slo) == [250, 500, 750] and not shi.any()
    hi = np.repeat(np.arange(4, dtype=np.uint64), 250)
    slo, shi = G.plan_splitters(np.zeros(1000, np.uint64), hi, 2)
    assert list(shi) == [2]


def _goss(*args):
    return subprocess.run([GOSS, *args], capture_output=True, text=True)


def test_cli_usage_errors_exit_1(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_text(">a\nACGTACGT\n")
    r = _goss("frobnicate")
    assert r.returncode == 1 and "unrecognised command" in r.stderr
    r = _goss("build-graph", "-O", str(tmp_path / "g"), "-I", str(fa))
    assert r.returncode == 1 and "--kmer-size" in r.stderr and "for more usage information" in r.stderr
    r = _goss("build-graph", "-k", "25", "-I", str(fa))
    assert r.returncode == 1 and "--graph-out" in r.stderr
    r = _goss("build-graph", "-k", "25", "-O", str(tmp_path / "g"), "-I", str(fa), "--bogus")
    assert r.returncode == 1 and "unrecognised option '--bogus'" in r.stderr
    r = _goss("build-graph", "-k", "63", "-O", str(tmp_path / "g"), "-I", str(fa))
    assert r.returncode == 1 and "at most 62" in r.stderr
    r = _goss("build-kmer-set", "-k", "63", "-O", str(tmp_path / "g"), "-I", str(tmp_path / "missing.fa"))
    assert r.returncode == 1 and "missing.fa" in r.stderr
    r = _goss("build-graph", "-k", "25", "-O", str(tmp_path / "nodir" / "g"), "-I", str(fa))
    assert r.returncode == 1 and "cannot create filenames with prefix" in r.stderr
    r = _goss("build-graph", "-h")
    assert r.returncode == 0 and "--kmer-size" in r.stdout
    assert _goss("help").returncode == 0


def test_cli_file_set_commands_usage(tmp_path):
    """Option surface of the commands that read existing file sets (host/goss_rewrite.cc): every one of these fails (or prints
    its help) before a device is touched."""
    r = _goss("trim-graph", "-G", "in", "-O", str(tmp_path / "out"))
    assert r.returncode == 1 and "--cutoff" in r.stderr
    r = _goss("trim-graph", "-O", str(tmp_path / "out"), "-C", "1")
    assert r.returncode == 1 and "--graph-in" in r.stderr
    r = _goss("trim-graph", "-G", "a", "-G", "b", "-O", str(tmp_path / "out"), "-C", "1")
    assert r.returncode == 1 and "only be given once" in r.stderr
    r = _goss("trim-graph", "-G", "a", "-O", str(tmp_path / "out"), "-C", "minus-one")
    assert r.returncode == 1 and "is invalid" in r.stderr
    r = _goss("merge-graphs", "-O", str(tmp_path / "out"))
    assert r.returncode == 1 and "At least one input graph" in r.stderr
    r = _goss("merge-kmer-sets", "-G", "a", "-G", "b")
    assert r.returncode == 1 and "--graph-out" in r.stderr
    r = _goss("merge-and-annotate-kmer-sets", "-G", "a", "-O", str(tmp_path / "out"))
    assert r.returncode == 1 and "exactly twice" in r.stderr
    r = _goss("merge-and-annotate-kmer-sets", "-G", "a", "-G", "b", "-G", "c", "-O", str(tmp_path / "out"))
    assert r.returncode == 1 and "exactly twice" in r.stderr
    r = _goss("compute-near-kmers")
    assert r.returncode == 1 and "--graph-in" in r.stderr
    r = _goss("dump-graph", "-G", "a", "--max-merge", "3")
    assert r.returncode == 1 and "unrecognised option" in r.stderr
    r = _goss("restore-graph", "-f", str(tmp_path / "missing.txt"), "-O", str(tmp_path / "nodir" / "g"))
    assert r.returncode == 1 and "cannot create filenames with prefix" in r.stderr
    for cmd in ("trim-graph", "merge-graphs", "merge-kmer-sets", "dump-graph", "restore-graph", "merge-and-annotate-kmer-sets", "compute-near-kmers"):
        r = _goss(cmd, "-h")
        assert r.returncode == 0 and r.stdout.startswith("usage: goss " + cmd), cmd
    # a file set that does not exist: the message names it, exit code 1 (no device needed to find that out)
    r = _goss("trim-graph", "-G", str(tmp_path / "nothing"), "-O", str(tmp_path / "out"), "-C", "1")
    assert r.returncode == 1 and "error performing trim-graph" in r.stderr and "nothing" in r.stderr


@pytest.mark.skipif(_have_gpu(), reason="checks the no-GPU failure mode")
def test_cli_without_gpu_fails_loudly(tmp_path):
    fa = tmp_path / "a.fa"
    fa.write_text(">a\nACGTACGT\n")
    r = _goss("build-graph", "-k", "3", "-O", str(tmp_path / "g"), "-I", str(fa))
    assert r.returncode == 1 and "error performing build-graph" in r.stderr and "no CPU path" in r.stderr
    assert not list(tmp_path.glob("g*"))


def test_key_mix_is_a_bijection_and_separates_substitution_variants(tmp_path):
    """gossamer_b200/csrc/keys.h: key_unmix(key_mix(x)) == x (64- and 128-bit keys) and single-base substitutions never share
    the low 32 bits of the mixed key beyond chance -- compiled with the host compiler alone (tests/cpp/mix_check.cc)."""
    import shutil
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    inc = "/usr/local/cuda/include"
    if not cxx or not os.path.exists(os.path.join(inc, "cuda_runtime.h")):
        pytest.skip("needs g++ and the CUDA headers")
    exe = str(tmp_path / "mix_check")
    r = subprocess.run([cxx, "-O2", "-std=c++17", "-w", "-I", inc, "-o", exe, os.path.join(ROOT, "tests", "cpp", "mix_check.cc")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def test_compressed_input_readers(tmp_path):
    """host/file_io.cc InputFile: plain, .gz and .bz2 input decode to the same bytes (src/PhysicalFileFactory.cc:261-280);
    concatenated bzip2 / gzip streams are read through; truncated bzip2 data is an error, not a short file."""
    import bz2
    import gzip
    import shutil
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
    lib_dir = os.path.join(ROOT, "gossamer_b200")
    if not cxx or not os.path.exists(os.path.join(lib_dir, "libgossamer_b200.so")):
        pytest.skip("needs g++ and the built library")
    exe = str(tmp_path / "inputfile_check")
    r = subprocess.run([cxx, "-O1", "-std=c++17", "-w", "-I", "/usr/local/cuda/include", "-o", exe, os.path.join(ROOT, "tests", "cpp", "inputfile_check.cc"),
                        os.path.join(lib_dir, "host", "file_io.cc"), "-L", lib_dir, "-lgossamer_b200", "-Wl,-rpath," + lib_dir, "-lz", "-ldl"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import simreads_py as S
    text = bytes(S.reads_fastq(S.genome(30_000, 3), 100, 9_000, err=0.01, seed=4))          # ~2 MB
    half = text.index(b"\n@r", len(text) // 2) + 1
    cases = {
        "plain.fq": text,
        "one.fq.gz": gzip.compress(text, 6),
        "two.fq.gz": gzip.compress(text[:half], 1) + gzip.compress(text[half:], 9),
        "one.fq.bz2": bz2.compress(text, 9),
        "two.fq.bz2": bz2.compress(text[:half], 1) + bz2.compress(text[half:], 9),
        "empty.fq.bz2": bz2.compress(b""),
    }
    for name, data in cases.items():
        (tmp_path / name).write_bytes(data)
        r = subprocess.run([exe, str(tmp_path / name)], capture_output=True)
        assert r.returncode == 0, (name, r.stderr)
        assert r.stdout == (b"" if name.startswith("empty") else text), name
    (tmp_path / "cut.fq.bz2").write_bytes(cases["one.fq.bz2"][:len(cases["one.fq.bz2"]) // 2])
    r = subprocess.run([exe, str(tmp_path / "cut.fq.bz2")], capture_output=True)
    assert r.returncode == 1 and b"bzip2" in r.stderr
    (tmp_path / "junk.fq.bz2").write_bytes(b"this is not bzip2 data at all, not even close" * 10)
    r = subprocess.run([exe, str(tmp_path / "junk.fq.bz2")], capture_output=True)
    assert r.returncode == 1 and b"bzip2" in r.stderr


def test_pass_planning_invariants():
    """Host logic of partition.cu (no device): the pass widths of the counting, of a streamed build (first pass fixed at 8 bits,
    run per block) and of the pair sort -- every pass at most 10 bits wide, the widths add up, buckets fit their tables with
    head room, a streamed plan always has a gathering pass."""
    import gossamer_b200 as G
    for kb, key_bits in ((8, 64), (16, 112)):
        for n in (0, 1, 100, 5_000, 300_000, 30_000_000, 198_333_373, 1_900_000_000, 40_000_000_000):
            levels, total, slots, bits = G.debug_plan(0, kb, key_bits, n)
            assert sum(bits) == total and len(bits) == levels and all(1 <= b <= 10 for b in bits)
            cap = slots - slots // 8 - 2
            assert n == 0 or (n >> total) <= cap // 2 or total == 40             # mean bucket at most half of what a table holds
            slevels, stotal, sslots, sbits = G.debug_plan(1, kb, key_bits, n, 8)
            assert sbits[0] == 8 and slevels >= 2 and sum(sbits) == stotal == max(total, 8) and all(0 <= b <= 10 for b in sbits[1:])
            assert sslots == slots
            for first in (0, 10):
                plevels, ptotal, pcap, pb = G.debug_plan(2, kb, key_bits, n, first)
                assert sum(pb) == ptotal and all(0 <= b <= 10 for b in pb) and ptotal <= key_bits
                if first:
                    assert pb[0] == 10 and plevels >= 1
                assert n == 0 or ptotal == key_bits or (n >> ptotal) <= pcap // 2
