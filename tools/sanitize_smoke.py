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

# ---- round 2 kernels: device-side Frame tail (RGB-D), resident frames + in-place (pinned) inputs/outputs of the small calls,
# CTA-wide claim resolution, database SearchByBoW (node-major items, compact pairs), ComputeBoW, persistent scoring kernel
import dataclasses
X = ORBextractor(1000)
imgs = [synth.mono_frame(40 + i, 0, 0, 640, 480) for i in range(2)]
outs = X.extract_batch(imgs)
rng = np.random.default_rng(0)
raw = (5000.0 * (1.5 + 0.5 * rng.random((480, 640)))).astype(np.uint16)
TUM1_K = (517.306408, 516.469215, 318.643040, 255.313989)
TUM1_DIST = (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)
frames, host = M.frames_from_extractor(mt, X, [1, 0], [len(outs[1][0]), len(outs[0][0])], TUM1_K, TUM1_DIST, bf=40.0, mode=2,
                                       depth=[raw, raw], depth_factor=np.float32(1.0 / 5000.0))
keys_un, desc = host["keys_un"][0], outs[1][1]
vv = dict(w=640, h=480, kl=keys_un, dl=desc, kr=keys_un, dr=desc, ur=host["u_right"][0], disp=np.zeros((480, 640), np.float32),
          scale=X.GetScaleFactors(), sigma2=X.GetScaleSigmaSquares())
Fh, mps2 = mf.projection_case(vv, 9, n_mp=300)
FR = dataclasses.replace(frames[0], occupied=Fh.occupied)
print("resident proj", mt.SearchByProjection(FR, mps2, 3.0)[0])
Cur2, Last2, Tcw2, K2 = mf.last_frame_case(vv, 3)
CurR = dataclasses.replace(frames[0], occupied=Cur2.occupied)
print("resident last", mt.SearchByProjectionLast(CurR, Last2, Tcw2, K2, 40.0, 7.0)[0])
voc6 = M.ORBVocabulary.from_arrays(*__import__("orb_slam2_b200.sharding", fromlist=["x"]).random_vocabulary_arrays(10, 4, 7), 10, 4)
db2 = M.KeyFrameDatabase(mt)
for j in range(12):
    k_, d_ = outs[j % 2]
    flip = (rng.random((len(d_), 32, 8)) < 0.03)
    d_ = d_ ^ np.packbits(flip, axis=2, bitorder="little").reshape(len(d_), 32)
    bow_, fv_ = voc6.ComputeBoW(d_, 2)
    hm = (rng.random(len(k_)) < 0.6).astype(np.uint8)
    db2.add(M.KeyFrameView(mvKeysUn=k_, mDescriptors=d_, mFeatVec=fv_, has_mp=hm), bow_)
qb, qf = voc6.ComputeBoW(outs[0][1], 2)
Fq = M.KeyFrameView(mvKeysUn=outs[0][0], mDescriptors=outs[0][1], mFeatVec=qf)
nm, off, pairs = db2.SearchByBoWPairs(None, Fq)
print("bowdb pairs", int(nm.sum()), "query", db2.query(qb)[0][:4], "dense", db2.SearchByBoW(np.arange(12, dtype=np.int32), Fq)[0][:4])
