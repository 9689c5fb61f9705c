"""CPU: the C-ABI library loads and exports every symbol include/quickrank_b200.h declares, and
fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "quickrank_b200.h")
LIB = os.path.join(ROOT, "quickrank_b200", "libquickrank_b200.so")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(qr_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    assert os.path.exists(LIB), "build the CUDA library first (python -c 'import __graft_entry__ as g; g.build()')"
    lib = ctypes.CDLL(LIB)
    names = declared_functions()
    assert len(names) >= 25
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_header_cites_the_reference_interfaces():
    src = open(HEADER).read()
    for cite in ("mart.cc:117-176", "lambdamart.cc:62-152", "mart.cc:459-468", "ltr_algorithm.cc:44-52",
                 "rtnode_histogram.cc", "metric.h:93-106", "ranker.cc:23-25", "dart.cc:634-687",
                 "lambdamartselective.cc:185-206", "lambdamart.cc:84-105", "line_search.cc:252-281",
                 "quality_loss_adv_pruning.cc:88-92", "score_loss_pruning.cc:58-63"):
        assert cite in src, cite


def test_no_device_is_a_loud_error():
    from quickrank_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    x = np.zeros((10, 2), np.float32)
    with pytest.raises(api.QrError) as e:
        api.Trainer(x, np.zeros(10, np.float32), np.array([0, 10], np.uint64))
    assert "no CPU fallback" in str(e.value)
    tree = dict(feature=np.array([-1], np.int32), threshold=np.zeros(1, np.float32), left=np.array([-1], np.int32),
                right=np.array([-1], np.int32), value=np.zeros(1), threshold_idx=np.zeros(1, np.uint32))
    with pytest.raises(api.QrError):
        api.Scorer([tree], [1.0], 2)


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under quickrank_b200/ or host/ may reference it."""
    bad = []
    for base in ("quickrank_b200", "host"):
        for dirpath, _dirs, files in os.walk(os.path.join(ROOT, base)):
            if "build" in dirpath or "__pycache__" in dirpath:
                continue
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".cc", ".cpp", ".h", ".hpp")):
                    txt = open(os.path.join(dirpath, fn), errors="replace").read()
                    if re.search(r"qr_oracle|pyoracle|pyref|libqr_ref|from oracle|import oracle", txt):
                        bad.append(os.path.join(dirpath, fn))
    assert not bad, bad
