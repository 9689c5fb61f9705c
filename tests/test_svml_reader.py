"""The host SVMLight reader (host/src/quickrank_host.cc, Svml::read_horizontal; reference src/io/svml.cc:38-161):
same grammar, parsed by all host threads.  CPU only."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECK = os.path.join(ROOT, "host", "bin", "svml_check")

pytestmark = pytest.mark.skipif(not os.path.exists(CHECK), reason="host/bin/svml_check not built")


def fnv(b):
    h = 1469598103934665603
    for c in b:
        h = ((h ^ c) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def run(path, threads):
    env = dict(os.environ, QR_SVML_THREADS=str(threads))
    out = subprocess.run([CHECK, path], capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stderr
    return out.stdout.split()


def test_parallel_parse_equals_serial_and_an_independent_parse(tmp_path):
    rng = np.random.default_rng(2)
    n, f = 30000, 23
    x = np.round(rng.random((n, f)), 4).astype(np.float32)
    x[rng.random((n, f)) < 0.3] = 0.0                     # sparse rows: zero features are omitted
    labels = rng.integers(0, 5, size=n)
    qlen = rng.integers(1, 120, size=2000)
    off = np.concatenate([[0], np.cumsum(qlen)])
    off = off[off < n].tolist() + [n]
    path = str(tmp_path / "data.txt")
    with open(path, "w") as fh:
        fh.write("# a comment line\n\n")
        for q in range(len(off) - 1):
            for i in range(off[q], off[q + 1]):
                feats = " ".join("%d:%.9g" % (j + 1, x[i, j]) for j in range(f) if x[i, j] != 0 or j == f - 1)
                eol = "\r\n" if i % 97 == 0 else "\n"
                fh.write("%d qid:%d %s%s%s" % (labels[i], q + 1, feats, " # doc %d" % i if i % 13 == 0 else "", eol))
        fh.write("1 qid:%d 1:0.5 %d:0.25" % (len(off), f))   # last line without a newline
    serial = run(path, 1)
    for t in (2, 7, 16):
        assert run(path, t) == serial
    xx = np.vstack([x, np.zeros((1, f), np.float32)])
    xx[-1, 0], xx[-1, f - 1] = 0.5, 0.25
    ll = np.concatenate([labels, [1]]).astype(np.float32)
    oo = np.array(off + [n + 1], np.uint64)
    assert serial[:3] == [str(n + 1), str(f), str(len(off))]
    assert int(serial[3], 16) == fnv(ll.tobytes())
    assert int(serial[4], 16) == fnv(oo.tobytes())
    assert int(serial[5], 16) == fnv(np.ascontiguousarray(xx).tobytes())


@pytest.mark.parametrize("line,code", [("x qid:1 1:0.5\n", 2), ("1 1:0.5\n", 2), ("1 qid:1 a:0.5\n", 4), ("1 qid:1 0:0.5\n", 4)])
def test_malformed_lines_exit_like_the_reference(tmp_path, line, code):
    """svml.cc:91,112: exit(2) for a bad label / qid, exit(4) for a bad feature token."""
    path = str(tmp_path / "bad.txt")
    open(path, "w").write("2 qid:1 1:0.1 2:0.2\n" + line)
    out = subprocess.run([CHECK, path], capture_output=True, text=True)
    assert out.returncode == code


def test_binary_cache_round_trip_and_invalidation(tmp_path):
    """QR_SVML_CACHE=1: <file>.qrb is written after the parse and read instead of the text while the text's size and
    mtime are unchanged; a changed text is parsed again (and the cache rewritten)."""
    path = str(tmp_path / "data.txt")
    rows = ["%d qid:%d 1:%.3f 3:%.3f 7:%.3f" % (i % 5, i // 6 + 1, i * 0.001, i * 0.5, 1.0 / (i + 1)) for i in range(500)]
    open(path, "w").write("\n".join(rows) + "\n")
    plain = run(path, 3)
    assert not os.path.exists(path + ".qrb")
    env = dict(os.environ, QR_SVML_CACHE="1")
    first = subprocess.run([CHECK, path], capture_output=True, text=True, env=env)
    assert first.returncode == 0 and first.stdout.split() == plain and os.path.exists(path + ".qrb")
    # the second run must come from the cache: corrupt the text IN PLACE (same size, same mtime) and see it ignored
    st = os.stat(path)
    with open(path, "r+b") as fh:
        fh.write(b"x" * 64)
    os.utime(path, ns=(st.st_atime_ns, st.st_mtime_ns))
    second = subprocess.run([CHECK, path], capture_output=True, text=True, env=env)
    assert second.returncode == 0 and second.stdout.split() == plain
    assert subprocess.run([CHECK, path], capture_output=True, text=True).returncode == 2   # without the cache: malformed
    # a new text (different size) invalidates the cache
    open(path, "w").write("\n".join(rows[:300]) + "\n")
    third = subprocess.run([CHECK, path], capture_output=True, text=True, env=env)
    assert third.returncode == 0 and third.stdout.split()[0] == "300"
    # a truncated cache is ignored, not trusted
    blob = open(path + ".qrb", "rb").read()
    open(path + ".qrb", "wb").write(blob[:len(blob) // 2])
    os.utime(path + ".qrb")
    fourth = subprocess.run([CHECK, path], capture_output=True, text=True, env=env)
    assert fourth.returncode == 0 and fourth.stdout.split() == third.stdout.split()


def test_host_reader_equals_the_reference_reader(tmp_path):
    """The same file through the unmodified reference's io::Svml::read_horizontal (svml.cc:38-161, compiled into
    oracle/_ref): shape, labels, query boundaries and every feature value equal (checksums of the raw arrays)."""
    from oracle import pyref
    if not pyref.available():
        pytest.skip("oracle/_ref/libqr_ref.so not built")
    rng = np.random.default_rng(4)
    n, f = 8000, 31
    x = rng.normal(size=(n, f)).astype(np.float32)
    x[rng.random((n, f)) < 0.4] = 0.0
    labels = rng.integers(0, 5, size=n)
    qlen = rng.integers(1, 90, size=400)
    off = np.concatenate([[0], np.cumsum(qlen)])
    off = off[off < n].tolist() + [n]
    path = str(tmp_path / "data.txt")
    with open(path, "w") as fh:
        for q in range(len(off) - 1):
            for i in range(off[q], off[q + 1]):
                feats = " ".join("%d:%.9g" % (j + 1, x[i, j]) for j in range(f) if x[i, j] != 0 or j == f - 1)
                fh.write("%d qid:%d %s%s\n" % (labels[i], q + 7, feats, " # d%d" % i if i % 11 == 0 else ""))
    shape, sums, _sec = pyref.read_svml(path)
    got = run(path, 5)
    assert [int(v) for v in got[:3]] == list(shape)
    assert tuple(int(v, 16) for v in got[3:6]) == sums


def test_fast_float_parse_equals_strtof():
    """host/include/qr_fast_float.h (the reader's decimal -> float fast path) against strtof on 2 million random and
    adversarial strings — exact float rounding boundaries and their neighbours, long digit strings, exponents,
    range limits, specials: same bits, same end pointer."""
    tool = os.path.join(ROOT, "host", "bin", "float_check")
    assert os.path.exists(tool), "build host/ first"
    out = subprocess.run([tool, "2", "7"], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip().endswith(" 0 mismatches"), out.stdout + out.stderr
