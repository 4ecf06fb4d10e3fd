"""CPU: the link oracle (oracle/link_oracle.c + numpy column arithmetic) is pinned BIT FOR BIT to
the real reference (tests/golden/link.npz from picasso.postprocess.link / _get_link_groups)."""
import os

import numpy as np
import pytest

from picasso_b200 import testing

CASES = (("plain", {}), ("group", {"with_group": True}), ("f64", {"f64_xy": True, "seed": 8}))


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "link.npz"))


@pytest.mark.parametrize("tag,kw", CASES)
def test_link_groups_bit_identical(oracle, g, tag, kw):
    locs, info = testing.synthetic_link_locs(**kw)
    sl = locs.sort_values(kind="quicksort", by="frame")
    group = sl["group"].to_numpy() if "group" in sl.columns else np.zeros(len(sl), np.int32)
    for dark in (3, 1):
        lg = oracle.get_link_groups(sl["frame"].to_numpy(), sl["x"].to_numpy(), sl["y"].to_numpy(), 0.05, dark, group)
        np.testing.assert_array_equal(lg, g[f"{tag}_lg_dark{dark}"])


@pytest.mark.parametrize("tag,kw", CASES)
def test_linked_table_bit_identical(oracle, g, tag, kw):
    locs, info = testing.synthetic_link_locs(**kw)
    linked, lg = oracle.link(locs, info)
    cols = [k[len(tag) + 8:] for k in g.files if k.startswith(f"{tag}_linked_") and k != f"{tag}_linked_index"]
    assert list(linked.columns) == cols
    np.testing.assert_array_equal(linked.index.to_numpy(), g[f"{tag}_linked_index"])
    for c in cols:
        ref = g[f"{tag}_linked_{c}"]
        assert linked[c].dtype == ref.dtype, c
        assert linked[c].to_numpy().tobytes() == ref.tobytes(), c


def test_end_of_data_quirk(oracle):
    """When no later frame exists the reference's `min_index` stays at N-1: the very last
    localization can be linked within the same frame (postprocess.py:2528-2540)."""
    frame = np.array([0, 1, 5, 5, 5], np.int64)
    x = np.array([1.0, 1.0, 3.0, 9.0, 3.0], np.float32)
    y = np.zeros(5, np.float32)
    lg = oracle.get_link_groups(frame, x, y, 0.05, 3, np.zeros(5, np.int32))
    np.testing.assert_array_equal(lg, [0, 0, 1, 2, 1])
