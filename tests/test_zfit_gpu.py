"""GPU parity of the astigmatic z fit (csrc/zfit.cu through pb_zfit): the Brent trajectory is
BIT-IDENTICAL to scipy's minimize_scalar on the real reference's target (z, residual and number
of function evaluations, golden + oracle); lpz (float32 column arithmetic, the reference uses
libm powf for z**k) within 2e-5 relative."""
import os

import numpy as np
import pytest

from picasso_b200 import testing, zfit

pytestmark = pytest.mark.gpu

CASES = (("lq_f0", "gausslq", True, 0), ("lq_f2", "gausslq", True, 2),
         ("mle_f0", "gaussmle", True, 0), ("mleunc_f2", "gaussmle", False, 2))
LPZ_RTOL = 2e-5


@pytest.fixture(scope="module")
def g(golden_dir):
    return np.load(os.path.join(golden_dir, "zfit.npz"))


def test_minimiser_bit_identical_to_reference(g):
    locs, info, calib = testing.synthetic_zfit_locs(3000, 11)
    z, dz, lpz, nfev = zfit._run(locs, g["cx"], g["cy"], 1.0, 130, "gausslq", want_nfev=True)
    assert z.tobytes() == g["raw_z"].astype(np.float32).tobytes()
    assert dz.tobytes() == np.sqrt(g["raw_fun"].astype(np.float32)).tobytes()
    np.testing.assert_array_equal(nfev, g["raw_nfev"])


@pytest.mark.parametrize("tag,method,drop_unc,flt", CASES)
def test_zfit_tables_match_reference(g, tag, method, drop_unc, flt):
    locs, info, calib = testing.synthetic_zfit_locs(3000, 11)
    if drop_unc:
        locs = locs.drop(columns=["sx_unc", "sy_unc"])
    n_info = len(info)
    res, new_info = zfit.zfit(locs, info, calibration=dict(calib), fitting_method=method, filter=flt)
    np.testing.assert_array_equal(res.index.to_numpy(), g[f"{tag}_index"])
    assert list(res.columns[-3:]) == ["z", "d_zcalib", "lpz"]
    for c in ("z", "d_zcalib", "lpz"):
        assert res[c].dtype == np.float32
    assert res["z"].to_numpy().tobytes() == g[f"{tag}_z"].tobytes()
    assert res["d_zcalib"].to_numpy().tobytes() == g[f"{tag}_d_zcalib"].tobytes()
    np.testing.assert_allclose(res["lpz"].to_numpy(), g[f"{tag}_lpz"], rtol=LPZ_RTOL)
    assert len(new_info) == n_info + 1 and new_info[-1]["Filter range"] == flt
    assert new_info[-1]["X Coefficients"] == calib["X Coefficients"]


def test_large_batch_matches_oracle(oracle):
    """200 k localizations: z, residual and nfev bit-identical to the C oracle; recovered z close
    to the simulated one inside the calibrated range."""
    locs, info, calib = testing.synthetic_zfit_locs(200_000, 5)
    cx, cy = np.array(calib["X Coefficients"]), np.array(calib["Y Coefficients"])
    z, dz, lpz, nfev = zfit._run(locs, cx, cy, 1.0, 130, "gaussmle", want_nfev=True)
    oz, osq, _, _, onf = oracle.zfit_minimise(locs["sx"].to_numpy(), locs["sy"].to_numpy(), cx, cy)
    assert z.tobytes() == oz.tobytes()
    assert dz.tobytes() == np.sqrt(osq).tobytes()
    np.testing.assert_array_equal(nfev, onf)
    assert np.isfinite(lpz[2000:]).all() and (lpz[2000:] > 0).all()


def test_api_contract(g):
    locs, info, calib = testing.synthetic_zfit_locs(500, 3)
    seen = []
    res, inf2 = zfit.zfit(locs, [{"Width": 64, "Height": 64, "Frames": 1000}], calibration=dict(calib),
                          pixelsize=100, magnification_factor=1.0, filter=0, progress_callback=seen.append)
    assert seen == list(range(500))
    assert inf2[-2] == {"Pixelsize": 100.0} and inf2[-1]["Magnification factor"] == 1.0
    res2, _ = zfit.zfit(locs, info, calibration=dict(calib), filter=0)
    # z scales with the magnification factor (reference test_zfit.py:222-234)
    both = res.index.intersection(res2.index)
    np.testing.assert_allclose(res2.loc[both, "z"], res.loc[both, "z"] * np.float32(0.79), rtol=1e-6)
    par = zfit.fit_z_parallel(locs, info, calib, 0.79, 130, filter=0)
    assert par["z"].to_numpy().tobytes() == res2["z"].to_numpy().tobytes()
    fut = zfit.fit_z_parallel(locs, info, calib, 0.79, 130, asynch=True)
    assert len(zfit.locs_from_futures(fut, filter=0)) == len(res2)
    lp = zfit.axial_localization_precision(res2, info, calib, "gausslq")
    np.testing.assert_allclose(lp.to_numpy(), res2["lpz"].to_numpy(), rtol=LPZ_RTOL)


@pytest.mark.filterwarnings("ignore::DeprecationWarning")
@pytest.mark.parametrize("method", ["gausslq", "gaussmle"])
def test_localize_3d_equals_localize_plus_zfit(method):
    """Reference localize.py:1818-2034: localize_3D == localize(...) followed by zfit(filter=0)."""
    from picasso_b200 import localize

    movie = testing.synthetic_movie(8, 64, 64, emitters_per_frame=6, seed=13)
    cam = {"Baseline": 100, "Sensitivity": 1.0, "Gain": 1, "Pixelsize": 130}
    minfo = [{"Width": 64, "Height": 64, "Frames": 8}]
    _, _, calib = testing.synthetic_zfit_locs(4, 1)
    locs3, info3 = localize.localize_3D(movie, movie_info=list(minfo), camera_info=dict(cam), box=7,
                                        minimum_ng=5000, calibration_3d=dict(calib), fitting_method=method)
    locs2, info2 = localize.localize(movie, dict(cam), {"Min. Net Gradient": 5000, "Box Size": 7},
                                     movie_info=list(minfo), fitting_method=method, return_info=True)
    ref, _ = zfit.zfit(locs2, info2, calibration=dict(calib), fitting_method=method, filter=0)
    assert len(locs3) == len(ref) > 20 and list(locs3.columns[-3:]) == ["z", "d_zcalib", "lpz"]
    for c in ref.columns:
        assert locs3[c].to_numpy().tobytes() == ref[c].to_numpy().tobytes(), c
    assert info3[-1]["Filter range"] == 0 and info3[0] == minfo[0]
    with pytest.raises(AssertionError):
        localize.localize_3D(movie, movie_info=minfo, camera_info=cam, box=6, minimum_ng=5000,
                             calibration_3d=calib)


def _messy_table(n=200_000, seed=2):
    import pandas as pd

    rng = np.random.default_rng(seed)
    t = pd.DataFrame({
        "frame": rng.integers(0, 1000, n).astype(np.uint32),
        "x": rng.uniform(-2, 70, n).astype(np.float32), "y": rng.uniform(-2, 50, n).astype(np.float32),
        "photons": rng.uniform(-10, 5000, n).astype(np.float32), "sx": rng.uniform(-0.1, 2, n).astype(np.float32),
        "sy": rng.uniform(0.5, 2, n).astype(np.float32), "bg": rng.uniform(0, 50, n).astype(np.float32),
        "lpx": rng.uniform(-0.01, 0.2, n).astype(np.float32), "lpy": rng.uniform(0.0, 0.2, n).astype(np.float32),
        "ellipticity": rng.uniform(-0.01, 0.5, n).astype(np.float32),
        "net_gradient": rng.uniform(-5e3, 5e4, n).astype(np.float32),       # may be negative: not filtered
        "group": rng.integers(-3, 3, n).astype(np.int32),
    })
    for c, v in (("bg", np.nan), ("net_gradient", np.inf), ("lpy", -np.inf), ("x", np.nan)):
        t.loc[rng.integers(0, n, 50), c] = np.float32(v)
    return t, [{"Width": 64, "Height": 48, "Frames": 1000, "Pixelsize": 130}]


def test_ensure_sanity_device_equals_host_chain():
    """lib.ensure_sanity on the GPU (fused mask + stream compaction, csrc/table.cu) keeps exactly the
    rows, index labels, dtypes and values of the reference's filter chain (lib.py:1786-1832; host
    single-mask form pinned in tests/test_host_logic_cpu.py)."""
    import pandas as pd

    from picasso_b200 import lib

    t, info = _messy_table()
    assert lib._device_table_ok(t)
    got = lib.ensure_sanity(t, info)
    ref = lib._ensure_sanity_host(t, info)
    assert 0 < len(ref) < len(t)
    pd.testing.assert_frame_equal(got, ref)
    # the reference's own chain, literally
    r = t.copy()
    r.replace([np.inf, -np.inf], np.nan, inplace=True)
    r.dropna(axis=0, how="any", inplace=True)
    r = r[r["x"] < 64]; r = r[r["y"] < 48]
    for attr in ["x", "y", "lpx", "lpy", "photons", "ellipticity", "sx", "sy"]:
        r = r[r[attr] >= 0]
    pd.testing.assert_frame_equal(got, r)
    # a non-default index is carried through
    t2 = t.set_index(np.arange(len(t))[::-1] * 3)
    pd.testing.assert_frame_equal(lib.ensure_sanity(t2, info), lib._ensure_sanity_host(t2, info))
    # tables with 8-byte columns take the host path, same rows
    t3 = t.copy(); t3["n_id"] = np.arange(len(t), dtype=np.int64)
    assert not lib._device_table_ok(t3)
    assert lib.ensure_sanity(t3, info).index.equals(ref.index)
    with pytest.raises(KeyError):
        lib.ensure_sanity(t, [{"Width": 64, "Height": 48}])


def test_locs_to_records_matches_save_locs_packing():
    """io.save_locs (reference io.py:2089-2110): lib.ensure_sanity(locs, info).to_records(index=False);
    sanity filter, compaction and the column -> record transposition on the device."""
    from picasso_b200 import lib

    t, info = _messy_table(50_000, seed=4)
    rec = lib.locs_to_records(t, info)
    ref = lib._ensure_sanity_host(t, info).to_records(index=False)
    assert rec.dtype == ref.dtype and rec.shape == ref.shape
    assert rec.tobytes() == ref.tobytes()


def test_fused_zfit_equals_stepwise_path(g):
    """The fused device call (upload once -> z fit -> ensure_sanity -> RMSD filter with numpy's float32
    pairwise sum -> one download) against the step-by-step host path, 150 k localizations, filter 0 / 2
    / 3, incl. rows the sanity pass removes."""
    import pandas as pd

    from picasso_b200 import lib

    locs, info, calib = testing.synthetic_zfit_locs(150_000, 7)
    rng = np.random.default_rng(0)
    locs.loc[rng.integers(0, len(locs), 300), "photons"] = np.float32(-1.0)
    locs.loc[rng.integers(0, len(locs), 300), "sx"] = np.float32(np.nan)
    cx, cy = np.array(calib["X Coefficients"]), np.array(calib["Y Coefficients"])
    for flt in (0, 2, 3):
        got = zfit._fit_z(locs, info, calib, calib["Magnification factor"], 130, "gaussmle", flt)
        step = locs.copy()
        z, dz, lpz = zfit._run(step, cx, cy, calib["Magnification factor"], 130, "gaussmle")
        step["z"], step["d_zcalib"], step["lpz"] = z, dz, lpz
        step = zfit.filter_z_fits(lib._ensure_sanity_host(step, info), flt)
        assert 0 < len(step) < len(locs)
        pd.testing.assert_frame_equal(got, step)
