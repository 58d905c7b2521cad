"""GPU parity for the ORBmatcher paths and the BoW feeder through the C ABI, against the restatements in
oracle/orb_port_match.cpp: bit-exact indices and counts (all integer / order-dependent work)."""
import os

import numpy as np
import pytest

from tests import match_fixtures as mf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    from orb_slam2_b200 import matcher
    return matcher


@pytest.fixture(scope="module")
def views(oracle):
    return {s: mf.two_views(oracle, s) for s in (7, 8)}


@pytest.mark.parametrize("seed", [7, 8])
@pytest.mark.parametrize("th,ratio", [(1.0, 0.8), (3.0, 0.8), (5.0, 0.9)])
def test_search_by_projection(M, oracle, views, seed, th, ratio):
    F, mps = mf.projection_case(views[seed], seed + 10, n_mp=400)
    n_o, m_o = oracle.port_search_by_projection(F, mps, th, ratio)
    n_g, m_g = M.ORBmatcher(ratio, True).SearchByProjection(F, mps, th)
    assert n_o > 30
    assert n_g == n_o and np.array_equal(m_g, m_o), int((m_g != m_o).sum())


def test_search_by_projection_edge_cases(M, oracle, views):
    v = views[7]
    F, mps = mf.projection_case(v, 99, n_mp=50)
    mt = M.ORBmatcher(0.8, True)
    # monocular frame (no stereo check), no occupancy, all map points valid and observed
    F2 = M.FrameView(F.mvKeysUn, F.mDescriptors, F.mvScaleFactors, F.bounds)
    mp2 = M.MapPointsView(mps.mTrackProjX, mps.mTrackProjY, mps.mTrackProjXR, mps.mnTrackScaleLevel, mps.mTrackViewCos, mps.descriptors)
    assert_same = lambda a, b: (a[0] == b[0] and np.array_equal(a[1], b[1]))
    assert assert_same(mt.SearchByProjection(F2, mp2, 3.0), oracle.port_search_by_projection(F2, mp2, 3.0, 0.8))
    # every map point projects to the same place: the order-dependent claiming decides
    mp3 = M.MapPointsView(np.full(50, 320.0, np.float32), np.full(50, 240.0, np.float32), np.full(50, 300.0, np.float32),
                          np.zeros(50, np.int32), np.full(50, 0.9, np.float32), mps.descriptors)
    assert assert_same(mt.SearchByProjection(F2, mp3, 15.0), oracle.port_search_by_projection(F2, mp3, 15.0, 0.8))
    # projections outside the image / nothing in range
    mp4 = M.MapPointsView(np.full(50, -500.0, np.float32), np.full(50, 9000.0, np.float32), np.zeros(50, np.float32),
                          np.full(50, 7, np.int32), np.ones(50, np.float32), mps.descriptors)
    n, m = mt.SearchByProjection(F2, mp4, 3.0)
    assert n == 0 and np.all(m == -1)
    # zero map points
    mp5 = M.MapPointsView(*[np.zeros(0, np.float32)] * 3, np.zeros(0, np.int32), np.zeros(0, np.float32), np.zeros((0, 32), np.uint8))
    n, m = mt.SearchByProjection(F2, mp5, 3.0)
    assert n == 0 and len(m) == 0


@pytest.mark.parametrize("seed", [7, 8])
@pytest.mark.parametrize("ratio,ori", [(0.7, True), (0.9, True), (0.75, False)])
def test_search_by_bow_both_variants(M, oracle, views, seed, ratio, ori):
    voc = oracle.PortVocabulary.random(10, 4, 5)
    kf1, kf2 = mf.keyframe_views(views[seed], voc, seed)
    mt = M.ORBmatcher(ratio, ori)
    n_o, m_o = oracle.port_search_by_bow(kf1, kf2, ratio, ori)
    n_g, m_g = mt.SearchByBoW(kf1, kf2)
    assert n_o > 20 and n_g == n_o and np.array_equal(m_g, m_o)
    n_o, m_o = oracle.port_search_by_bow_kf(kf1, kf2, ratio, ori)
    n_g, m_g = mt.SearchByBoW_KF(kf1, kf2)
    assert n_g == n_o and np.array_equal(m_g, m_o)


def test_search_by_bow_batched_keyframes(M, oracle, views):
    """Config-5 shape in miniature: one query frame against several keyframes in one launch."""
    voc = oracle.PortVocabulary.random(10, 4, 5)
    a1, a2 = mf.keyframe_views(views[7], voc, 1)
    b1, b2 = mf.keyframe_views(views[8], voc, 2)
    kfs = [a1, b1, b2, a1]
    nm, match = M.ORBmatcher(0.75, True).SearchByBoW(kfs, a2)
    for i, kf in enumerate(kfs):
        n_o, m_o = oracle.port_search_by_bow(kf, a2, 0.75, True)
        assert nm[i] == n_o and np.array_equal(match[i], m_o), i
    assert np.array_equal(match[0], match[3])


@pytest.mark.parametrize("only_stereo", [False, True])
@pytest.mark.parametrize("ori", [True, False])
def test_search_for_triangulation(M, oracle, views, only_stereo, ori):
    voc = oracle.PortVocabulary.random(10, 4, 5)
    for seed in (7, 8):
        kf1, kf2 = mf.keyframe_views(views[seed], voc, seed + 3, mp_frac=0.3)
        F12 = mf.rectified_F12(seed)
        ep = (-1000.0, 200.0) if seed == 7 else (300.0, 240.0)      # the second epipole sits inside the image: exercises :743-749
        want = oracle.port_search_for_triangulation(kf1, kf2, F12, ep, only_stereo, ori)
        got = M.ORBmatcher(0.6, ori).SearchForTriangulation(kf1, kf2, F12, ep, only_stereo)
        assert len(want) > 5 and np.array_equal(got, want)


def test_vocabulary_transform_and_blob_round_trip(M, oracle, views, tmp_path):
    pv = oracle.PortVocabulary.random(10, 5, 21)             # 111,111 nodes
    e = pv.export()
    voc = M.ORBVocabulary.from_arrays(e["parent"], e["is_leaf"], e["desc"], e["weight"], e["k"], e["L"])
    d = np.concatenate([views[7]["dl"], views[8]["dr"]])
    for levelsup in (4, 3, 0, 7):
        wg, tg, ng = voc.transform_raw(d, levelsup)
        wo, to, no = pv.transform_raw(d, levelsup)
        assert np.array_equal(wg, wo) and np.array_equal(tg, to) and np.array_equal(ng, no), levelsup
    # text loader (ORBvoc.txt format) gives the same tree
    path = os.path.join(tmp_path, "voc.txt")
    small = oracle.PortVocabulary.random(10, 3, 4)
    small.save_text(path)
    v2 = M.ORBVocabulary.loadFromTextFile(path)
    assert all(np.array_equal(a, b) for a, b in zip(v2.transform_raw(d[:500], 2), small.transform_raw(d[:500], 2)))
    # packed blob adoption (what the NCCL broadcast receiver does)
    ptr, nbytes = voc.blob()
    v3 = M.ORBVocabulary.from_blob(ptr, nbytes)
    assert all(np.array_equal(a, b) for a, b in zip(v3.transform_raw(d[:300], 4), pv.transform_raw(d[:300], 4)))
    bow, fv = voc.transform(views[7]["dl"], 4)
    assert abs(sum(bow.values()) - 1.0) < 1e-9 and list(bow) == sorted(bow)
    assert np.all(np.diff(fv.node_id.astype(np.int64)) > 0) and fv.start[-1] == len(views[7]["dl"])


@pytest.mark.parametrize("seed", [7, 8])
@pytest.mark.parametrize("mode", [(False, False), (True, False), (False, True)])
@pytest.mark.parametrize("th,ori", [(7.0, True), (15.0, True), (15.0, False)])
def test_search_by_projection_from_last_frame(M, oracle, views, seed, mode, th, ori):
    """SearchByProjection(CurrentFrame, LastFrame, th, bMono), src/ORBmatcher.cc:1328-1470 (TrackWithMotionModel)."""
    Cur, Last, Tcw, K = mf.last_frame_case(views[seed], seed + 20)
    fw, bw = mode
    n_o, s_o = oracle.port_search_by_projection_last(Cur, Last, Tcw, K, 40.0, th, fw, bw, ori)
    n_g, s_g = M.ORBmatcher(0.9, ori).SearchByProjectionLast(Cur, Last, Tcw, K, 40.0, th, fw, bw)
    assert n_o > 20
    assert n_g == n_o and np.array_equal(s_g, s_o), int((s_g != s_o).sum())
    m = s_g[s_g >= 0]
    assert np.all(Last.valid[m] == 1) and np.all(Cur.occupied[np.nonzero(s_g >= 0)[0]] == 0)


@pytest.mark.parametrize("seed", [7, 8])
@pytest.mark.parametrize("th,orb_dist,ori", [(10.0, 100, True), (3.0, 64, True), (10.0, 100, False), (25.0, 50, True)])
def test_search_by_projection_from_keyframe(M, oracle, views, seed, th, orb_dist, ori):
    """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist), src/ORBmatcher.cc:1472-1599 (Relocalization)."""
    Cur, P, Tcw, Ow, K = mf.world_points_case(views[seed], seed + 30)
    n_o, s_o = oracle.port_search_by_projection_kf(Cur, P, Tcw, Ow, K, th, orb_dist, ori)
    n_g, s_g = M.ORBmatcher(0.9, ori).SearchByProjectionKF(Cur, P, Tcw, Ow, K, th, orb_dist)
    assert n_o > 20
    assert n_g == n_o and np.array_equal(s_g, s_o), int((s_g != s_o).sum())
    hit = np.nonzero(s_g >= 0)[0]
    assert np.all(P.valid[s_g[hit]] == 1) and np.all(Cur.occupied[hit] == 0)


@pytest.mark.parametrize("seed", [7, 8])
@pytest.mark.parametrize("th", [3, 10, 25])
def test_search_by_projection_sim3(M, oracle, views, seed, th):
    """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th), src/ORBmatcher.cc:290-403 (LoopClosing::ComputeSim3)."""
    KF, P, Tcw, Ow, K = mf.world_points_case(views[seed], seed + 40)
    n_o, s_o = oracle.port_search_by_projection_sim3(KF, P, Tcw, Ow, K, th)
    n_g, s_g = M.ORBmatcher(0.75, True).SearchByProjectionSim3(KF, P, Tcw, Ow, K, th)
    assert n_o > 20
    assert n_g == n_o and np.array_equal(s_g, s_o), int((s_g != s_o).sum())
    hit = np.nonzero(s_g >= 0)[0]
    assert np.all(P.valid[s_g[hit]] == 1) and np.all(KF.occupied[hit] == 0)


def test_pose_projection_overloads_edge_cases(M, oracle, views):
    v = views[7]
    F, P, Tcw, Ow, K = mf.world_points_case(v, 77)
    mt = M.ORBmatcher(0.9, True)
    same = lambda a, b: (a[0] == b[0] and np.array_equal(a[1], b[1]))
    # no occupancy, no validity mask; identity pose puts most points outside the image
    F2 = M.FrameView(F.mvKeysUn, F.mDescriptors, F.mvScaleFactors, F.bounds)
    P2 = M.WorldPointsView(P.world_pos, P.descriptors, P.max_distance, P.min_distance, P.normal, P.angle)
    assert same(mt.SearchByProjectionKF(F2, P2, Tcw, Ow, K, 10.0, 100), oracle.port_search_by_projection_kf(F2, P2, Tcw, Ow, K, 10.0, 100, True))
    assert same(mt.SearchByProjectionSim3(F2, P2, Tcw, Ow, K, 10), oracle.port_search_by_projection_sim3(F2, P2, Tcw, Ow, K, 10))
    I = np.eye(4, dtype=np.float32)[:3]
    z3 = np.zeros(3, np.float32)
    assert same(mt.SearchByProjectionKF(F2, P2, I, z3, K, 10.0, 100), oracle.port_search_by_projection_kf(F2, P2, I, z3, K, 10.0, 100, True))
    assert same(mt.SearchByProjectionSim3(F2, P2, I, z3, K, 10), oracle.port_search_by_projection_sim3(F2, P2, I, z3, K, 10))
    # points exactly at the camera centre (zero distance, division by zero depth) and at infinity must not crash or match
    bad = P.world_pos.copy(); bad[:40] = Ow; bad[40:80] = 1e30
    P3 = M.WorldPointsView(bad, P.descriptors, P.max_distance, P.min_distance, P.normal, P.angle)
    assert same(mt.SearchByProjectionKF(F2, P3, Tcw, Ow, K, 10.0, 100), oracle.port_search_by_projection_kf(F2, P3, Tcw, Ow, K, 10.0, 100, True))
    assert same(mt.SearchByProjectionSim3(F2, P3, Tcw, Ow, K, 10), oracle.port_search_by_projection_sim3(F2, P3, Tcw, Ow, K, 10))
    # zero query points
    P0 = M.WorldPointsView(np.zeros((0, 3), np.float32), np.zeros((0, 32), np.uint8), np.zeros(0, np.float32), np.zeros(0, np.float32),
                           np.zeros((0, 3), np.float32), np.zeros(0, np.float32))
    n, s = mt.SearchByProjectionSim3(F2, P0, Tcw, Ow, K, 10)
    assert n == 0 and np.all(s == -1)
