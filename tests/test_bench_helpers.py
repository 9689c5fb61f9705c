"""bench.py's pure helpers (CPU): the numbers the JSON line derives from measured times."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_bench():
    spec = importlib.util.spec_from_file_location("qr_bench", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_shared_atomics_view():
    b = load_bench()
    # one root launch: 1M documents x 136 features x 2 limb atomics in 91 us ~ 2.99 T lane-atomics/s
    v = b.shared_atomics_view(1e6, 1e6, 0.091, 136)
    assert abs(v["achieved"] - 2.989) < 0.01 and v["ceiling"] == 3.1 and abs(v["frac"] - 0.964) < 0.01
    # child launches add the count atomic
    v = b.shared_atomics_view(2e6, 1e6, 0.2, 100)
    assert abs(v["achieved"] - (1e6 * 2 + 1e6 * 3) * 100 / 0.2e-3 / 1e12) < 1e-3
    # degenerate inputs never raise
    assert b.shared_atomics_view(0.0, 0.0, 0.0, 136)["achieved"] == 0.0


def test_algorithmic_bytes_per_tree_follow_the_survey_formula():
    b = load_bench()
    n, f, leaves = 1_000_000, 136, 64
    # SURVEY.md section 8d with rho = 3, sigma = 6: about 0.88 GB per tree
    total = b.tree_bytes(n, f, 3.0, 6.0, leaves, f * 257)
    assert abs(total - 0.88e9) < 0.01e9
