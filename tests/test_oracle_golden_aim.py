"""CPU: the AIM oracle (oracle/aim_oracle.py) is pinned BIT FOR BIT to the real reference
(tests/golden/aim.npz from picasso.aim.aim: per-frame drift, undrifted coordinates and every
per-segment intersection-count array, 2-D and 3-D), plus GPU-free host logic of picasso_b200.aim."""
import os

import numpy as np
import pytest

from picasso_b200 import aim, testing


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "aim.npz"))


@pytest.mark.parametrize("tag,with_z", [("2d", False), ("3d", True)])
def test_oracle_bit_identical_to_reference(g, tag, with_z):
    from oracle import aim_oracle

    locs, info, truth = testing.synthetic_aim_locs(with_z=with_z)
    rec = {}
    und, drift = aim_oracle.aim(locs, info, 100, record=rec)
    np.testing.assert_array_equal(np.stack(rec["roi_cc"]), g[f"{tag}_roi_cc"])
    if with_z:
        np.testing.assert_array_equal(np.stack(rec["roi_cc_z"]), g["3d_roi_cc_z"])
    for c in drift.columns:
        assert drift[c].dtype == np.float32
        assert drift[c].to_numpy().tobytes() == g[f"{tag}_drift_{c}"].tobytes(), c
        assert und[c].to_numpy().tobytes() == g[f"{tag}_und_{c}"].tobytes(), c
    d = drift["x"].to_numpy()
    t = truth[:, 0]
    assert np.abs((d - d.mean()) - (t - t.mean())).max() < 0.03      # the injected drift is recovered


def test_oracle_unsorted_frames_other_parameters(g):
    from oracle import aim_oracle
    import pandas as pd

    locs = pd.DataFrame({"frame": g["alt_perm_frame"], "x": g["alt_perm_x"], "y": g["alt_perm_y"]})
    info = [{"Height": 48, "Width": 80, "Frames": 900, "Pixelsize": 130}]
    und, drift = aim_oracle.aim(locs, info, 150, 0.2, 0.55)
    assert drift["x"].to_numpy().tobytes() == g["alt_drift_x"].tobytes()
    assert drift["y"].to_numpy().tobytes() == g["alt_drift_y"].tobytes()
    assert und["x"].to_numpy().tobytes() == g["alt_und_x"].tobytes()


def test_segment_ranges_equal_reference_masks():
    rng = np.random.default_rng(0)
    frame = rng.integers(1, 501, 3000)
    bounds = np.concatenate((np.arange(0, 500, 70), [500]))
    for fr in (frame, np.sort(frame)):                      # unsorted: argsort path; sorted: slice path
        order, start, end = aim._segments_in_frame_order(fr, bounds)
        idx = np.arange(len(fr))[order]
        for s in range(len(bounds) - 1):
            mask = (fr > bounds[s]) & (fr <= bounds[s + 1])
            assert sorted(idx[start[s]:end[s]]) == list(np.flatnonzero(mask))


def test_fft_peaks():
    roi = np.zeros((7, 7), np.int32)
    roi[3, 3] = 100
    px, py = aim.get_fft_peak(roi, 6.0)
    assert abs(px) < 1e-12 and abs(py) < 1e-12
    roi = np.zeros((7, 7), np.int32)
    roi[5, 2] = 50                                   # x index 5 -> +2 units, y index 2 -> -1 unit
    px, py = aim.get_fft_peak(roi, 7.0)
    assert abs(px - 2.0) < 1e-9 and abs(py + 1.0) < 1e-9
    z = np.zeros(7, np.int32)
    z[4] = 9
    assert abs(aim.get_fft_peak_z(z, 7.0) - 1.0) < 1e-9
