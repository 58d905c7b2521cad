"""Seeded matcher cases shared by tests/golden/make_golden_match.py (which records what the REFERENCE's verbatim-compiled
src/ORBmatcher.cc / src/Frame.cc produce), tests/test_golden_oracle.py (CPU: restatement == golden) and
tests/test_gpu_match.py (GPU: CUDA library == golden).  Each case: name -> (build(oracle) -> ctx, port(oracle, ctx),
gpu(M, ctx), ref(oracle, ctx)); all three calls return a tuple of integer arrays / ints in the same convention."""
import numpy as np

from tests import match_fixtures as mf

_views = {}


def _v(O, seed):
    if seed not in _views:
        _views[seed] = mf.two_views(O, seed)
    return _views[seed]


def _i(*xs):
    return tuple(np.asarray(x).astype(np.int64) for x in xs)


def camera_center_f32(Tcw):
    """Ow = -Rcw.t()*tcw as the reference's float32 cv::Mat arithmetic evaluates it (src/ORBmatcher.cc:1478): products and sums in
    float32, left to right, no FMA — numpy's matmul is free to differ in the last bit, this is not."""
    f32 = np.float32
    T = np.asarray(Tcw, f32)
    R, t = T[:3, :3], T[:3, 3]
    return np.array([f32(f32(f32(-R[0, i] * t[0]) + f32(-R[1, i] * t[1])) + f32(-R[2, i] * t[2])) for i in range(3)], f32)


def _proj_build(O):
    F, mps = mf.projection_case(_v(O, 7), 17, n_mp=400)
    return dict(F=F, mps=mps)


def _last_build(O):
    Cur, Last, Tcw, K = mf.last_frame_case(_v(O, 8), 28)
    return dict(Cur=Cur, Last=Last, Tcw=Tcw, K=K)


def _world_build(O, seed, off):
    F, P, Tcw, _, K = mf.world_points_case(_v(O, seed), seed + off)
    return dict(F=F, P=P, Tcw=Tcw, Ow=camera_center_f32(Tcw), K=K)


def _bow_build(O):
    voc = O.PortVocabulary.random(10, 4, 5)
    kf1, kf2 = mf.keyframe_views(_v(O, 7), voc, 8)
    return dict(kf1=kf1, kf2=kf2)


def _tri_build(O):
    voc = O.PortVocabulary.random(10, 4, 5)
    kf1, kf2 = mf.keyframe_views(_v(O, 8), voc, 10, mp_frac=0.4)
    return dict(kf1=kf1, kf2=kf2, F12=mf.rectified_F12(8), epi=(-1000.0, 200.0))


def _init_build(O):
    from orb_slam2_b200.matcher import FrameView
    v = _v(O, 7)
    b = (0.0, 0.0, float(v["w"]), float(v["h"]))
    return dict(F1=FrameView(v["kl"], v["dl"], v["scale"], b), F2=FrameView(v["kr"], v["dr"], v["scale"], b),
                prev=np.stack([v["kl"]["x"], v["kl"]["y"]], 1).astype(np.float32))


def _sim3_build(O):
    KF1, KF2, P1, P2, T1w, T2w, S12, S21, K = mf.sim3_case(_v(O, 7), 67)
    return dict(a=(KF1, KF2, P1, P2, T1w, T2w, S12, S21, K))


def _fuse_build(O):
    KF, P, Tcw, _, K, bf = mf.fuse_case(_v(O, 8), 58)
    return dict(KF=KF, P=P, Tcw=Tcw, Ow=camera_center_f32(Tcw), K=K, bf=bf)


CASES = {
    "projection_local_map": (
        _proj_build,
        lambda O, c: _i(*O.port_search_by_projection(c["F"], c["mps"], 3.0, 0.8)),
        lambda M, c: _i(*M.ORBmatcher(0.8, True).SearchByProjection(c["F"], c["mps"], 3.0))),
    "projection_last_frame": (
        _last_build,
        lambda O, c: _i(*O.port_search_by_projection_last(c["Cur"], c["Last"], c["Tcw"], c["K"], 40.0, 15.0, False, False, True)),
        lambda M, c: _i(*M.ORBmatcher(0.9, True).SearchByProjectionLast(c["Cur"], c["Last"], c["Tcw"], c["K"], 40.0, 15.0, False, False))),
    "projection_keyframe": (
        lambda O: _world_build(O, 7, 30),
        lambda O, c: _i(*O.port_search_by_projection_kf(c["F"], c["P"], c["Tcw"], c["Ow"], c["K"], 10.0, 100, True)),
        lambda M, c: _i(*M.ORBmatcher(0.9, True).SearchByProjectionKF(c["F"], c["P"], c["Tcw"], c["Ow"], c["K"], 10.0, 100))),
    "projection_sim3": (
        lambda O: _world_build(O, 8, 40),
        lambda O, c: _i(*O.port_search_by_projection_sim3(c["F"], c["P"], c["Tcw"], c["Ow"], c["K"], 10)),
        lambda M, c: _i(*M.ORBmatcher(0.75, True).SearchByProjectionSim3(c["F"], c["P"], c["Tcw"], c["Ow"], c["K"], 10))),
    "bow_keyframe_frame": (
        _bow_build,
        lambda O, c: _i(*O.port_search_by_bow(c["kf1"], c["kf2"], 0.7, True)),
        lambda M, c: _i(*M.ORBmatcher(0.7, True).SearchByBoW(c["kf1"], c["kf2"]))),
    "bow_keyframe_keyframe": (
        _bow_build,
        lambda O, c: _i(*O.port_search_by_bow_kf(c["kf1"], c["kf2"], 0.75, True)),
        lambda M, c: _i(*M.ORBmatcher(0.75, True).SearchByBoW_KF(c["kf1"], c["kf2"]))),
    "triangulation": (
        _tri_build,
        lambda O, c: _i(O.port_search_for_triangulation(c["kf1"], c["kf2"], c["F12"], c["epi"], False, True)),
        lambda M, c: _i(M.ORBmatcher(0.6, True).SearchForTriangulation(c["kf1"], c["kf2"], c["F12"], c["epi"]))),
    "initialization": (
        _init_build,
        lambda O, c: _i(*O.port_search_for_initialization(c["F1"], c["F2"], c["prev"], 100, 0.9, True)[:2]),
        lambda M, c: _i(*M.ORBmatcher(0.9, True).SearchForInitialization(c["F1"], c["F2"], c["prev"], 100)[:2])),
    "sim3": (
        _sim3_build,
        lambda O, c: _i(*O.port_search_by_sim3(*c["a"], 7.5)),
        lambda M, c: _i(*M.ORBmatcher(0.75, True).SearchBySim3(*c["a"], 7.5))),
    "fuse_keyframe": (
        _fuse_build,
        lambda O, c: _i(*O.port_fuse(c["KF"], c["P"], c["Tcw"], c["Ow"], c["K"], c["bf"], 3.0, False)),
        lambda M, c: _i(*M.ORBmatcher().Fuse(c["KF"], c["P"], c["Tcw"], c["Ow"], c["K"], c["bf"], 3.0, Scw=False))),
    "fuse_scw": (
        _fuse_build,
        lambda O, c: _i(*O.port_fuse(c["KF"], c["P"], c["Tcw"], c["Ow"], c["K"], c["bf"], 3.0, True)),
        lambda M, c: _i(*M.ORBmatcher().Fuse(c["KF"], c["P"], c["Tcw"], c["Ow"], c["K"], c["bf"], 3.0, Scw=True))),
}


def flatten(res):
    """tuple of ints / arrays -> one int64 vector with length prefixes (what the golden file stores per case)"""
    out = []
    for r in res:
        r = np.atleast_1d(r).astype(np.int64).reshape(-1)
        out.append(np.array([len(r)], np.int64)); out.append(r)
    return np.concatenate(out)
