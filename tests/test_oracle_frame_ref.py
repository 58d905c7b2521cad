"""CPU: pins the restatements of Frame::ComputeStereoMatches (oracle/orb_port_stereo.cpp), of the feature grid and of
Frame::isInFrustum (oracle/orb_port_match.cpp) — and through them the CUDA kernels — to the REFERENCE SOURCE:
/root/reference/src/Frame.cc compiled verbatim against the reference's real include/Frame.h (oracle/_ref/libframeref.so,
oracle/Makefile target `ref`, oracle/frameref_wrap.cpp, oracle/frameshim/pre.hpp)."""
import numpy as np
import pytest

from orb_slam2_b200 import synth
from tests import match_fixtures as mf


@pytest.fixture(scope="module")
def O(oracle):
    if not oracle.have_frameref():
        pytest.skip("oracle/_ref/libframeref.so not built (reference tree absent)")
    return oracle


@pytest.mark.parametrize("shape,nf,bf,fx,seed", [((640, 360), 1000, 386.1448, 718.856, 1), ((752, 480), 1200, 47.9, 435.2, 2),
                                                 (synth.KITTI, 2000, 386.1448, 718.856, 3)])
def test_compute_stereo_matches_equals_reference_source(O, shape, nf, bf, fx, seed):
    w, h = shape
    L, R, _ = synth.stereo_pair(seed, 0, 0, w, h)
    E1, E2 = O.PortExtractor(nf), O.PortExtractor(nf)
    kl, dl = E1(L)
    kr, dr = E2(R)
    pl, pr = [E1.level(i) for i in range(8)], [E2.level(i) for i in range(8)]
    ur_p, dp_p, _ = O.port_stereo(kl, dl, kr, dr, pl, pr, E1.scale, E1.inv_scale, bf, fx)
    ur_r, dp_r = O.ref_stereo(kl, dl, kr, dr, pl, pr, E1.scale, E1.inv_scale, bf, fx)
    assert (ur_r >= 0).sum() > 0.3 * len(kl)
    assert np.array_equal(ur_r, ur_p) and np.array_equal(dp_r, dp_p), int((ur_r != ur_p).sum())


def test_stereo_degenerate_inputs(O):
    """Unrelated left / right images (few, poor matches: the median cull and the rejection branches dominate)."""
    L = synth.mono_frame(5, 0, 0, 640, 360)
    R = synth.mono_frame(6, 0, 0, 640, 360)
    E1, E2 = O.PortExtractor(800), O.PortExtractor(800)
    kl, dl = E1(L)
    kr, dr = E2(R)
    pl, pr = [E1.level(i) for i in range(8)], [E2.level(i) for i in range(8)]
    ur_p, dp_p, _ = O.port_stereo(kl, dl, kr, dr, pl, pr, E1.scale, E1.inv_scale, 386.1448, 718.856)
    ur_r, dp_r = O.ref_stereo(kl, dl, kr, dr, pl, pr, E1.scale, E1.inv_scale, 386.1448, 718.856)
    assert np.array_equal(ur_r, ur_p) and np.array_equal(dp_r, dp_p)


def test_feature_grid_equals_reference_source(O):
    v = mf.two_views(O, 7)
    k = v["kl"]
    rng = np.random.default_rng(1)
    for bounds in [(0.0, 0.0, 640.0, 480.0), (-12.5, -7.25, 652.0, 491.5)]:
        for _ in range(150):
            x, y, r = rng.uniform(-30, 670), rng.uniform(-30, 510), rng.uniform(2, 80)
            lo = int(rng.integers(-1, 7)); hi = int(rng.integers(-1, 8))
            got = O.port_features_in_area(k, bounds, x, y, r, lo, hi)
            want = O.ref_features_in_area(k, bounds, x, y, r, lo, hi)
            assert np.array_equal(got, want), (bounds, x, y, r, lo, hi)          # same features in the same order


@pytest.mark.parametrize("seed", [7, 8])
@pytest.mark.parametrize("limit", [0.5, 0.8])
def test_is_in_frustum_equals_reference_source(O, seed, limit):
    v = mf.two_views(O, seed)
    F, P, Tcw, _, K = mf.world_points_case(v, seed + 70)
    ref = O.ref_is_in_frustum(F, P, Tcw, K, 40.0, limit)
    port = O.port_is_in_frustum(F, P, Tcw, ref["Ow"], K, 40.0, limit)              # mOw as Frame::UpdatePoseMatrices computes it
    assert ref["count"] == port["count"] > 100
    for f in ("in_view", "proj_x", "proj_y", "proj_xr", "level", "view_cos"):
        assert np.array_equal(ref[f], port[f]), f
