"""GPU parity of AIM drift correction (csrc/aim.cu through pb_aim_*): the intersection counts are
BIT-EXACT against the real reference (golden) and the oracle; with identical counts the host
peak / spline arithmetic is the reference's, so drift and undrifted coordinates are
bit-identical too."""
import os

import numpy as np
import pandas as pd
import pytest

from picasso_b200 import aim, testing

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "aim.npz"))


@pytest.mark.parametrize("tag,with_z", [("2d", False), ("3d", True)])
def test_aim_bit_identical_to_reference(g, tag, with_z):
    locs, info, _ = testing.synthetic_aim_locs(with_z=with_z)
    und, new_info, drift = aim.aim(locs, info, segmentation=100)
    for c in drift.columns:
        assert drift[c].dtype == np.float32
        assert drift[c].to_numpy().tobytes() == g[f"{tag}_drift_{c}"].tobytes(), c
        assert und[c].dtype == np.float64
        assert und[c].to_numpy().tobytes() == g[f"{tag}_und_{c}"].tobytes(), c
    assert list(drift.columns) == (["x", "y", "z"] if with_z else ["x", "y"])
    assert new_info[-1]["Segmentation"] == 100 and abs(new_info[-1]["Intersect distance (nm)"] - 20.0) < 1e-9
    assert len(new_info) == len(info) + 1


def test_round1_intersection_counts_equal_reference(g):
    """Every per-segment 7x7 count array of round 1 (float32 index arithmetic) and of round 2
    (float64) equals the reference's."""
    locs, info, _ = testing.synthetic_aim_locs()
    frame = locs["frame"] + 1 - locs["frame"].min()
    bounds = np.concatenate((np.arange(0, 2000, 100), [2000]))
    ref = frame <= 100
    rec = []
    x1, y1, _, _ = aim.intersection_max(locs["x"], locs["y"], locs["x"][ref], locs["y"][ref], frame, bounds,
                                        20 / 130, 60 / 130, 64, aim_round=1, _record=rec)
    aim.intersection_max(x1, y1, x1, y1, frame, bounds, 20 / 130, 60 / 130, 64, aim_round=2, _record=rec)
    got = np.stack(rec)
    assert got.shape == g["2d_roi_cc"].shape == (39, 7, 7)
    np.testing.assert_array_equal(got, g["2d_roi_cc"])


def test_unsorted_frames_offset_and_other_parameters(g):
    locs = pd.DataFrame({"frame": g["alt_perm_frame"], "x": g["alt_perm_x"], "y": g["alt_perm_y"]})
    info = [{"Height": 48, "Width": 80, "Frames": 900, "Pixelsize": 130}]
    und, _, drift = aim.aim(locs, info, segmentation=150, intersect_d=0.2, roi_r=0.55)
    assert drift["x"].to_numpy().tobytes() == g["alt_drift_x"].tobytes()
    assert drift["y"].to_numpy().tobytes() == g["alt_drift_y"].tobytes()
    assert und["x"].to_numpy().tobytes() == g["alt_und_x"].tobytes()
    assert und["y"].to_numpy().tobytes() == g["alt_und_y"].tobytes()


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_counts_equal_oracle_on_random_multisets(dt):
    """Raw pb_aim_count vs the oracle's count_grid: heavy multiplicities, coordinates beyond 2^24
    units (float32 index rounding), NaNs (np.int32(nan) = INT_MIN on both sides)."""
    from oracle import aim_oracle

    rng = np.random.default_rng(3)
    d, width = 0.01, 700.0                      # 70 000 units per row -> indices up to 4.9e9 wrap / round
    wu = width / d
    n0, n1 = 60_000, 9_000
    rx = rng.uniform(0, 30, n0).astype(dt); ry = rng.uniform(0, 4, n0).astype(dt)
    tx = rx[rng.integers(0, n0, n1)] + rng.normal(0, 0.004, n1).astype(dt)
    ty = ry[rng.integers(0, n0, n1)] + rng.normal(0, 0.004, n1).astype(dt)
    rx[:5] = np.nan
    tx[:3] = np.nan
    steps = np.arange(-3, 4)
    shifts = np.array([sx + sy * wu for sx in steps for sy in steps]).astype(np.int32)
    with np.errstate(invalid="ignore"):
        l0 = np.int32(np.round(rx / d) + np.round(ry / d) * wu)
        rel = (0.0123, -0.0345)
        a = tx.copy(); a += rel[0]
        b = ty.copy(); b += rel[1]
        l1 = np.int32(np.round(a / d) + np.round(b / d) * wu)
    c0, k0 = np.unique(l0, return_counts=True)
    c1, k1 = np.unique(l1, return_counts=True)
    want = aim_oracle.count_grid(c0, k0, c1, k1, shifts)
    with aim._Counter() as gpu:
        gpu.set_targets(tx, ty)
        gpu.set_reference(rx, ry, None, d, wu)
        got = gpu.count(0, n1, rel[0], rel[1], 0.0, shifts)
        assert (gpu.count(0, 0, 0, 0, 0, shifts) == 0).all()
    np.testing.assert_array_equal(got, want)
    assert want.sum() > 1000


def test_empty_segments_and_contract():
    locs, info, _ = testing.synthetic_aim_locs(n_frames=600, seed=6)
    locs = locs[(locs["frame"] < 200) | (locs["frame"] >= 300)].reset_index(drop=True)   # segment 2 is empty
    from oracle import aim_oracle

    und, _, drift = aim.aim(locs, info, segmentation=100)
    ound, odrift = aim_oracle.aim(locs, info, 100)
    assert drift["x"].to_numpy().tobytes() == odrift["x"].to_numpy().tobytes()
    assert und["y"].to_numpy().tobytes() == ound["y"].to_numpy().tobytes()
    with pytest.raises(KeyError):
        aim.aim(locs, [{"Width": 64}], segmentation=100)
    with pytest.raises(AssertionError):
        aim.aim(locs, info, progress=3)
