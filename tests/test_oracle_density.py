"""Pins oracle/density_oracle.py to the reference: every function is
checked against tests/golden/density.npz and gsio.npz, which hold the outputs of the
reference's own gsplat/gsmodel.py and gsplat/gau_io.py (tests/golden/make_golden_density.py)."""
import os

import numpy as np
import pytest

from oracle import density_oracle as do

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RTOL = 2e-6


def close(got, want, name, rtol=RTOL, atol=1e-7):
    assert got.shape == want.shape, (name, got.shape, want.shape)
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin), name
    assert np.array_equal(got[~fin], want[~fin]), name        # +-inf in the same places
    err = np.abs(got[fin].astype(np.float64) - want[fin]) - (rtol * np.abs(want[fin]) + atol)
    assert (err <= 0).all(), "%s: max excess %.3e" % (name, err.max())


@pytest.fixture(scope="module")
def dens():
    return dict(np.load(os.path.join(G, "density.npz")))


@pytest.fixture(scope="module")
def gsio():
    return dict(np.load(os.path.join(G, "gsio.npz")))


@pytest.mark.parametrize("tag", ["s_", "n_"])
def test_density_accumulate_matches_reference(dens, tag):
    acc, cnt = None, None
    for it in range(3):
        acc, cnt = do.accumulate(dens[tag + "acc%d_dloss_dus" % it], dens[tag + "acc%d_mask" % it], acc, cnt)
        close(acc, dens[tag + "acc%d_grad_accum" % it], "grad_accum step %d" % it, rtol=1e-6, atol=1e-13)
        assert np.array_equal(cnt, dens[tag + "acc%d_cunt" % it])


@pytest.mark.parametrize("tag", ["s_", "n_"])
def test_density_update_matches_reference(dens, tag):
    th = do.thresholds(float(dens[tag + "sense_size"]))
    P = {k: dens[tag + "in_" + k] for k in do.NAMES}
    has_state = (tag + "in_m_pws") in dens
    M = {k: dens[tag + "in_m_" + k] for k in do.NAMES} if has_state else None
    V = {k: dens[tag + "in_v_" + k] for k in do.NAMES} if has_state else None
    cls = do.classify(P["alphas_raw"], P["scales_raw"], dens[tag + "acc2_grad_accum"], dens[tag + "acc2_cunt"], th)
    for c in (do.KEEP, do.CLONE, do.SPLIT, do.PRUNE):
        assert (cls == c).sum() > 20, "fixture does not exercise class %d" % c
    assert (dens[tag + "acc2_cunt"] == 0).sum() >= 20                  # the 0/0 -> 0 rule is exercised
    oP, oM, oV, (K, C, S) = do.densify(P, M, V, cls, dens[tag + "z"])
    assert S == dens[tag + "z"].shape[0] and K + C + S == dens[tag + "out_pws"].shape[0]
    for k in do.NAMES:
        # survivors are moved, not recomputed: bit exact
        assert np.array_equal(oP[k][:K], dens[tag + "out_" + k][:K]), k
        close(oP[k], dens[tag + "out_" + k], "out " + k, rtol=3e-6, atol=2e-6)
        if has_state:
            assert np.array_equal(oM[k], dens[tag + "out_m_" + k]), k
            assert np.array_equal(oV[k], dens[tag + "out_v_" + k]), k
        else:
            assert (tag + "out_m_" + k) not in dens
    if has_state:
        a, m, v = do.reset_alpha(oP["alphas_raw"])
        close(a, dens[tag + "reset_alphas_raw"], "reset alphas", rtol=1e-6)
        assert np.array_equal(m, dens[tag + "reset_m_alphas_raw"]) and np.array_equal(v, dens[tag + "reset_v_alphas_raw"])
        assert (dens[tag + "out_alphas_raw"] > a).sum() > 50


@pytest.mark.parametrize("tag", ["deg3", "deg1"])
def test_gsio_matches_reference(gsio, tag):
    gs = do.decode_ply(gsio[tag + "_ply"].tobytes())
    for k in ("pw", "rot", "scale", "alpha", "sh"):
        close(np.asarray(gs[k]), gsio[tag + "_" + k], k)
    assert np.array_equal(np.asarray(gs["pw"]), gsio[tag + "_pw"]) and np.array_equal(np.asarray(gs["sh"]), gsio[tag + "_sh"])
    tp = do.training_params(gs)
    for k in do.NAMES:
        close(tp[k], gsio[tag + "_tp_" + k], "training " + k, rtol=3e-6, atol=2e-6)
    back = do.params_to_gs(tp)
    assert back.dtype["sh"].shape == (48,)
    for k in ("pw", "rot", "scale", "alpha", "sh"):
        close(np.asarray(back[k]), gsio[tag + "_back_" + k], "saved " + k, rtol=3e-6, atol=1e-7)
