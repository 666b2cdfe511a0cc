"""CPU checks of the int8 pre-filter's DEFINITION (oracle/flat_ip.py: int8_shadow, int8_query_digits,
int8_guard_band, int8_fast_scores) -- the restatement the GPU tests compare the device-built shadow
and the guard band with.  The inequality the exactness argument of csrc/dense.cu rests on,
|exact - fast| <= band, is checked here on random, clustered and adversarial rows."""
import numpy as np
import pytest

from oracle import flat_ip
from legal_rag_engine_b200 import synth


def test_shadow_reconstructs_rows_within_half_a_step():
    x = synth.host_vectors(3000, seed=1, dup_frac=0.0)
    x[5] = 0
    xi, scale, E, X = flat_ip.int8_shadow(x)
    assert xi.dtype == np.int8 and scale.dtype == np.float32
    assert np.abs(xi).max() <= 127 and not xi[5].any() and scale[5] == 0
    err = np.abs(x.astype(np.float64) - scale.astype(np.float64)[:, None] * xi)
    assert (err <= scale.astype(np.float64)[:, None] * 0.5001 + 1e-12).all()
    assert 0 < E < 0.02 and 0.98 < X < 1.02
    # every non-zero row uses the full range: its largest element maps to +-127
    nz = scale > 0
    assert (np.abs(xi[nz]).max(axis=1) == 127).all()


def test_query_digits_are_256_times_finer_than_a_row():
    q = synth.host_queries(8, seed=2)
    for b in range(8):
        hi, lo, cq = flat_ip.int8_query_digits(q[b])
        assert np.abs(hi).max() <= 127 and np.abs(lo).max() <= 127
        qhat = float(cq) * (256 * hi + lo)
        err = np.abs(q[b].astype(np.float64) - qhat)
        assert err.max() <= float(cq) * 1.6 + 1e-9                     # half a lo step; 1.5 where lo = +-128 is clamped
    hi, lo, cq = flat_ip.int8_query_digits(np.zeros(384, np.float16))
    assert not hi.any() and not lo.any() and cq == 0


@pytest.mark.parametrize("kind", ["random", "clustered", "spiky"])
def test_guard_band_bounds_the_fast_score_error(kind):
    rng = np.random.default_rng(7)
    n = 4000
    if kind == "random":
        x = synth.host_vectors(n, seed=3, dup_frac=0.0)
    elif kind == "clustered":
        c = rng.standard_normal((8, 384)).astype(np.float32)
        x = c[rng.integers(0, 8, n)] + 0.05 * rng.standard_normal((n, 384)).astype(np.float32)
        x = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float16)
    else:                                                   # one huge coordinate per row: the worst case for
        x = 0.02 * rng.standard_normal((n, 384)).astype(np.float32)   # a per-row scale
        x[np.arange(n), rng.integers(0, 384, n)] = 1.0
        x = (x / np.linalg.norm(x, axis=1, keepdims=True)).astype(np.float16)
    q = np.concatenate([synth.host_queries(3, seed=4), synth.host_planted_queries(x, [11], seed=5)])
    _, _, E, X = flat_ip.int8_shadow(x)
    s = flat_ip.exact_scores(x, q)
    for b in range(q.shape[0]):
        band = flat_ip.int8_guard_band(q[b], E, X)
        fast = flat_ip.int8_fast_scores(x, q[b]).astype(np.float64)
        assert np.abs(s[b] - fast).max() <= band, kind
