"""GPU parity of postprocess.link (csrc/link.cu): link groups BIT-EXACT against the real reference
(golden) and the oracle, also on dense data where several chains compete for the same
localization; the combined table bit-identical to the reference's DataFrame."""
import os

import numpy as np
import pandas as pd
import pytest

from picasso_b200 import postprocess, testing

pytestmark = pytest.mark.gpu

CASES = (("plain", {}), ("group", {"with_group": True}), ("f64", {"f64_xy": True, "seed": 8}))


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "link.npz"))


@pytest.mark.parametrize("tag,kw", CASES)
def test_link_bit_identical_to_reference(g, tag, kw):
    locs, info = testing.synthetic_link_locs(**kw)
    sl = locs.sort_values(kind="quicksort", by="frame")
    group = sl["group"].to_numpy() if "group" in sl.columns else np.zeros(len(sl), np.int32)
    for dark in (3, 1):
        lg = postprocess.get_link_groups(sl["frame"].to_numpy(), sl["x"].to_numpy(), sl["y"].to_numpy(), 0.05,
                                         dark, group)
        assert lg.dtype == np.int32
        np.testing.assert_array_equal(lg, g[f"{tag}_lg_dark{dark}"])
    linked = postprocess.link(locs, info, r_max=0.05, max_dark_time=3)
    cols = [k[len(tag) + 8:] for k in g.files if k.startswith(f"{tag}_linked_") and k != f"{tag}_linked_index"]
    assert list(linked.columns) == cols
    np.testing.assert_array_equal(linked.index.to_numpy(), g[f"{tag}_linked_index"])
    for c in cols:
        ref = g[f"{tag}_linked_{c}"]
        assert linked[c].dtype == ref.dtype, c
        assert linked[c].to_numpy().tobytes() == ref.tobytes(), c


@pytest.mark.parametrize("n,side,frames,r,dark,dt", [(30000, 12.0, 300, 0.15, 3, np.float32),
                                                     (30000, 12.0, 300, 0.15, 0, np.float64),
                                                     (20000, 3.0, 60, 0.2, 5, np.float32),
                                                     (5000, 1.0, 10, 5.0, 2, np.float32)])
def test_dense_data_equals_oracle(oracle, n, side, frames, r, dark, dt):
    """Random dense localizations: many candidates per step, competing chains, several groups, big
    connected components (the last case is one component) -- still the sequential result."""
    rng = np.random.default_rng(n + dark)
    frame = np.sort(rng.integers(0, frames, n)).astype(np.int64)
    x = rng.uniform(0, side, n).astype(dt)
    y = rng.uniform(0, side, n).astype(dt)
    group = rng.integers(0, 3, n).astype(np.int32)
    want = oracle.get_link_groups(frame, x, y, r, dark, group)
    got = postprocess.get_link_groups(frame, x, y, r, dark, group)
    np.testing.assert_array_equal(got, want)
    assert want.max() + 1 < n                      # something was actually linked


def test_end_of_data_quirk_and_edge_cases(oracle):
    frame = np.array([0, 1, 5, 5, 5], np.int64)
    x = np.array([1.0, 1.0, 3.0, 9.0, 3.0], np.float32)
    y = np.zeros(5, np.float32)
    lg = postprocess.get_link_groups(frame, x, y, 0.05, 3, np.zeros(5, np.int32))
    np.testing.assert_array_equal(lg, [0, 0, 1, 2, 1])
    one = postprocess.get_link_groups(frame[:1], x[:1], y[:1], 0.05, 3, np.zeros(1, np.int32))
    np.testing.assert_array_equal(one, [0])
    with pytest.raises(Exception, match="sorted"):
        postprocess.get_link_groups(frame[::-1].copy(), x, y, 0.05, 3, np.zeros(5, np.int32))
    locs, info = testing.synthetic_link_locs(n_frames=100, n_sites=10)
    empty = postprocess.link(locs.iloc[0:0], info)                        # reference test_postprocess.py:574-579
    assert len(empty) == 0 and {"len", "n", "photon_rate"} <= set(empty.columns)
    with pytest.raises(NotImplementedError):
        postprocess.link(locs, info, combine_mode="refit")
    kept_all = postprocess.link(locs, info, remove_ambiguous_lengths=False)
    kept = postprocess.link(locs, info)
    assert len(kept_all) >= len(kept) and (kept["frame"] > 0).all()


def test_group_reductions_equal_numpy(oracle):
    rng = np.random.default_rng(1)
    n, G = 50000, 7000
    lg = rng.integers(0, G, n).astype(np.int32)
    lg[:G] = np.arange(G)                                                  # every group used
    cols = [rng.normal(0, 1, n).astype(np.float32), rng.normal(0, 1, n), rng.integers(0, 1000, n).astype(np.uint32),
            rng.integers(-5, 5, n).astype(np.int32)]
    outs = postprocess._group_reduce(lg, G, cols + [cols[2], cols[2], cols[3]], [0, 0, 0, 0, 1, 2, 3])
    for c, o in zip(cols, outs[:4]):
        want = np.zeros(G, c.dtype)
        for i in range(n):
            want[lg[i]] += c[i]
        assert o.tobytes() == want.tobytes()
    order = np.argsort(lg, kind="stable")
    st = np.searchsorted(lg[order], np.arange(G))
    np.testing.assert_array_equal(outs[4], np.minimum.reduceat(cols[2][order], st))
    np.testing.assert_array_equal(outs[5], np.maximum.reduceat(cols[2][order], st))
    last = np.zeros(G, np.int32); last[lg] = cols[3]
    np.testing.assert_array_equal(outs[6], last)
