"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck) — dev tooling."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import oracle_lib as O
from orb_slam2_b200 import synth
from orb_slam2_b200.extractor import ORBextractor
from orb_slam2_b200 import matcher as M
from tests import match_fixtures as mf

L, R, _ = synth.stereo_pair(3, 0, 0, 640, 480)
G = ORBextractor(1000)
out = G.stereo_frames([L, L], [R, R], 40.0, 525.0)
print("stereo", len(out[0]["mvKeys"]), int((out[0]["mvuRight"] >= 0).sum()))
G2 = ORBextractor(500)
k, d = G2(synth.white_noise(1, 400, 300))
print("noise", len(k))
v = mf.two_views(O, 7)
F, mps = mf.projection_case(v, 1)
mt = M.ORBmatcher(0.8, True)
print("proj", mt.SearchByProjection(F, mps, 3.0)[0])
Cur, Last, Tcw, K = mf.last_frame_case(v, 2)
print("last", mt.SearchByProjectionLast(Cur, Last, Tcw, K, 40.0, 7.0)[0])
Fw, Pw, Tw, Ow, Kw = mf.world_points_case(v, 4)
print("kf/sim3", mt.SearchByProjectionKF(Fw, Pw, Tw, Ow, Kw, 10.0, 100)[0], mt.SearchByProjectionSim3(Fw, Pw, Tw, Ow, Kw, 10)[0])
KFf, Pf, Tf, Owf, Kf, bff = mf.fuse_case(v, 5)
print("fuse", mt.Fuse(KFf, Pf, Tf, Owf, Kf, bff, 3.0)[0], mt.Fuse(KFf, Pf, Tf, Owf, Kf, bff, 3.0, Scw=True)[0])
print("sim3", mt.SearchBySim3(*mf.sim3_case(v, 6), 7.5)[0])
bb = (0.0, 0.0, float(v["w"]), float(v["h"]))
print("init", mt.SearchForInitialization(M.FrameView(v["kl"], v["dl"], v["scale"], bb), M.FrameView(v["kr"], v["dr"], v["scale"], bb),
                                         np.stack([v["kl"]["x"], v["kl"]["y"]], 1), 100)[0])
pv = O.PortVocabulary.random(10, 3, 5)
e = pv.export()
voc = M.ORBVocabulary.from_arrays(e["parent"], e["is_leaf"], e["desc"], e["weight"], e["k"], e["L"])
kf1, kf2 = mf.keyframe_views(v, pv, 3, levelsup=1)
print("bow", mt.SearchByBoW(kf1, kf2)[0], mt.SearchByBoW_KF(kf1, kf2)[0], len(mt.SearchForTriangulation(kf1, kf2, mf.rectified_F12(1), (-1000.0, 200.0))))
print("voc", voc.transform_raw(v["dl"], 1)[0][:4])
db = M.KeyFrameDatabase(mt)
bow1, _ = voc.transform(v["dl"], 1)
bow2, _ = voc.transform(v["dr"], 1)
db.add(kf1, bow1); db.add(kf2, bow2)
print("kfdb", db.query(bow1)[0], db.SearchByBoW([0, 1], kf2)[0])
print("distinctive", mt.ComputeDistinctiveDescriptors([v["dl"][:9], v["dl"][:1], v["dl"][:0], v["dl"][:70]]))
