"""Synthetic inputs for the matcher tests: two views of one scene (a synthetic stereo pair), a seeded vocabulary tree,
and the snapshots (FrameView / MapPointsView / KeyFrameView) the matchers consume.  Test tooling."""
import numpy as np

from orb_slam2_b200 import synth
from orb_slam2_b200.matcher import FeatureVector, FrameView, KeyFrameView, MapPointsView

BF, FX = 386.1448, 718.856


def two_views(oracle, seed, shape=(640, 480), nf=1000):
    w, h = shape
    L, R, disp = synth.stereo_pair(seed, 0, 0, w, h)
    EL, ER = oracle.PortExtractor(nf), oracle.PortExtractor(nf)
    kl, dl = EL(L)
    kr, dr = ER(R)
    ur, dp, _ = oracle.port_stereo(kl, dl, kr, dr, [EL.level(i) for i in range(8)], [ER.level(i) for i in range(8)],
                                   EL.scale, EL.inv_scale, BF, FX)
    return dict(w=w, h=h, kl=kl, dl=dl, kr=kr, dr=dr, ur=ur, disp=disp, scale=EL.scale.copy(), sigma2=EL.sigma2.copy())


def projection_case(v, seed, n_mp=300, occupied_frac=0.1):
    rng = np.random.default_rng(seed)
    w, h = v["w"], v["h"]
    F = FrameView(mvKeysUn=v["kl"], mDescriptors=v["dl"], mvScaleFactors=v["scale"], bounds=(0.0, 0.0, float(w), float(h)),
                  mvuRight=v["ur"], occupied=(rng.random(len(v["kl"])) < occupied_frac).astype(np.uint8))
    sel = rng.choice(len(v["kr"]), size=min(n_mp, len(v["kr"])), replace=False)
    kr = v["kr"][sel]
    d = v["disp"][np.clip(kr["y"].astype(int), 0, h - 1), np.clip(kr["x"].astype(int), 0, w - 1)]
    px = (kr["x"] + d + rng.normal(0, 1.5, len(sel))).astype(np.float32)
    py = (kr["y"] + rng.normal(0, 1.5, len(sel))).astype(np.float32)
    lvl = np.clip(kr["octave"] + rng.integers(-1, 2, len(sel)), 0, 7).astype(np.int32)
    mps = MapPointsView(mTrackProjX=px, mTrackProjY=py, mTrackProjXR=(px - d.astype(np.float32) + rng.normal(0, 2, len(sel)).astype(np.float32)),
                        mnTrackScaleLevel=lvl, mTrackViewCos=rng.uniform(0.99, 1.0, len(sel)).astype(np.float32),
                        descriptors=v["dr"][sel], valid=(rng.random(len(sel)) < 0.95).astype(np.uint8),
                        has_obs=(rng.random(len(sel)) < 0.9).astype(np.uint8))
    return F, mps


def keyframe_views(v, voc, seed, levelsup=2, mp_frac=0.7):
    """KF1 = left view, KF2 = right view, FeatureVectors from the (oracle) vocabulary."""
    rng = np.random.default_rng(seed)
    out = []
    for k, d, ur in ((v["kl"], v["dl"], v["ur"]), (v["kr"], v["dr"], None)):
        _, weight, node = voc.transform_raw(d, levelsup)
        fv = FeatureVector.from_nodes(node, weight > 0)
        u = ur if ur is not None else np.where(rng.random(len(k)) < 0.5, k["x"] - 20.0, -1.0).astype(np.float32)
        out.append(KeyFrameView(mvKeysUn=k, mDescriptors=d, mFeatVec=fv, has_mp=(rng.random(len(k)) < mp_frac).astype(np.uint8),
                                mvuRight=u.astype(np.float32), mvScaleFactors=v["scale"], mvLevelSigma2=v["sigma2"]))
    return out


def rectified_F12(seed):
    rng = np.random.default_rng(seed)
    F = np.array([[0, 0, 0], [0, 0, 1], [0, -1, 0]], np.float32)
    return (F + rng.normal(0, 2e-6, (3, 3))).astype(np.float32)


def last_frame_case(v, seed, K=(525.0, 525.0, 319.5, 239.5), jitter=2.0):
    """Current frame = left view; LastFrame = right view whose MapPoints project (through a non-trivial pose) close to
    their true correspondences in the left view."""
    from orb_slam2_b200.matcher import LastFrameView
    rng = np.random.default_rng(seed)
    w, h = v["w"], v["h"]
    fx, fy, cx, cy = K
    Cur = FrameView(mvKeysUn=v["kl"], mDescriptors=v["dl"], mvScaleFactors=v["scale"], bounds=(0.0, 0.0, float(w), float(h)),
                    mvuRight=v["ur"], occupied=(rng.random(len(v["kl"])) < 0.05).astype(np.uint8))
    kr = v["kr"]
    d = v["disp"][np.clip(kr["y"].astype(int), 0, h - 1), np.clip(kr["x"].astype(int), 0, w - 1)]
    uA = kr["x"] + d + rng.normal(0, jitter, len(kr))
    vA = kr["y"] + rng.normal(0, jitter, len(kr))
    z = rng.uniform(2.0, 40.0, len(kr))
    Pc = np.stack([(uA - cx) * z / fx, (vA - cy) * z / fy, z], 1)
    z[rng.random(len(kr)) < 0.03] *= -1.0                      # a few points behind the camera (invzc < 0)
    Pc[:, 2] = z
    ang = 0.05
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]) @ \
        np.array([[1, 0, 0], [0, np.cos(0.02), -np.sin(0.02)], [0, np.sin(0.02), np.cos(0.02)]])
    t = np.array([0.3, -0.1, 0.5])
    Pw = (Pc - t) @ R                                             # R^T (Pc - t)
    Tcw = np.zeros((3, 4), np.float32); Tcw[:, :3] = R; Tcw[:, 3] = t
    Last = LastFrameView(mvKeysUn=kr, world_pos=Pw.astype(np.float32), descriptors=v["dr"],
                         valid=(rng.random(len(kr)) < 0.9).astype(np.uint8), has_obs=(rng.random(len(kr)) < 0.9).astype(np.uint8))
    return Cur, Last, Tcw, K


def world_points_case(v, seed, K=(525.0, 525.0, 319.5, 239.5), jitter=2.0):
    """Frame (or keyframe) = left view; query MapPoints = the right view's features placed in the world so that, through a
    non-trivial pose, they project near their true correspondences in the left view.  Distances, normals and validity
    are spread so that every rejection branch of the two pose-projection overloads is exercised."""
    from orb_slam2_b200.matcher import WorldPointsView
    rng = np.random.default_rng(seed)
    w, h = v["w"], v["h"]
    fx, fy, cx, cy = K
    F = FrameView(mvKeysUn=v["kl"], mDescriptors=v["dl"], mvScaleFactors=v["scale"], bounds=(0.0, 0.0, float(w), float(h)),
                  occupied=(rng.random(len(v["kl"])) < 0.05).astype(np.uint8))
    kr = v["kr"]
    n = len(kr)
    d = v["disp"][np.clip(kr["y"].astype(int), 0, h - 1), np.clip(kr["x"].astype(int), 0, w - 1)]
    uA = kr["x"] + d + rng.normal(0, jitter, n)
    vA = kr["y"] + rng.normal(0, jitter, n)
    z = rng.uniform(2.0, 40.0, n)
    Pc = np.stack([(uA - cx) * z / fx, (vA - cy) * z / fy, z], 1)
    flip = rng.random(n) < 0.03
    Pc[flip] *= -1.0                                              # behind the camera: projects to the same pixel with z < 0
    ang = 0.05
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]]) @ \
        np.array([[1, 0, 0], [0, np.cos(0.02), -np.sin(0.02)], [0, np.sin(0.02), np.cos(0.02)]])
    t = np.array([0.3, -0.1, 0.5])
    Pw = ((Pc - t) @ R).astype(np.float32)
    Tcw = np.zeros((3, 4), np.float32); Tcw[:, :3] = R; Tcw[:, 3] = t
    R32, t32 = Tcw[:, :3], Tcw[:, 3]
    Ow = (-(R32.T @ t32)).astype(np.float32)                      # -Rcw.t()*tcw
    dist = np.linalg.norm(Pw.astype(np.float64) - Ow, axis=1)
    scale = np.asarray(v["scale"], np.float64)
    maxd = dist * scale[np.clip(kr["octave"], 0, len(scale) - 1)] * rng.uniform(0.9, 1.1, n)
    maxd[rng.random(n) < 0.05] *= 0.3                            # too far for the point's scale-invariance range
    mind = maxd / scale[-1]
    mind[rng.random(n) < 0.05] *= 30.0                           # too close
    view = (Pw.astype(np.float64) - Ow) / np.maximum(dist, 1e-9)[:, None]
    tilt = rng.uniform(0.0, np.deg2rad(80.0), n)                 # > 60 deg fails the viewing-angle test
    axis = np.cross(view, rng.normal(size=(n, 3))); axis /= np.linalg.norm(axis, axis=1)[:, None]
    normal = view * np.cos(tilt)[:, None] + np.cross(axis, view) * np.sin(tilt)[:, None]
    P = WorldPointsView(world_pos=Pw, descriptors=v["dr"], max_distance=maxd.astype(np.float32), min_distance=mind.astype(np.float32),
                        normal=normal.astype(np.float32), angle=kr["angle"].astype(np.float32),
                        valid=(rng.random(n) < 0.9).astype(np.uint8))
    return F, P, Tcw, Ow, K


def fuse_case(v, seed, K=(525.0, 525.0, 319.5, 239.5), bf=200.0, jitter=1.0):
    """Keyframe = left view with stereo coordinates; candidate MapPoints = right-view features at the depth their
    disparity implies (so the stereo reprojection gate of Fuse sees consistent and inconsistent candidates)."""
    from orb_slam2_b200.matcher import WorldPointsView
    rng = np.random.default_rng(seed)
    w, h = v["w"], v["h"]
    fx, fy, cx, cy = K
    scale = np.asarray(v["scale"], np.float32)
    ur = v["ur"].copy()
    ur[rng.random(len(ur)) < 0.3] = -1.0                          # some keyframe features are monocular
    KF = FrameView(mvKeysUn=v["kl"], mDescriptors=v["dl"], mvScaleFactors=scale, bounds=(0.0, 0.0, float(w), float(h)),
                   mvuRight=ur, mvInvLevelSigma2=(np.float32(1.0) / (scale * scale)).astype(np.float32))
    kr = v["kr"]
    n = len(kr)
    d = np.maximum(v["disp"][np.clip(kr["y"].astype(int), 0, h - 1), np.clip(kr["x"].astype(int), 0, w - 1)], 1.0)
    uA = kr["x"] + d + rng.normal(0, jitter, n)
    vA = kr["y"] + rng.normal(0, jitter, n)
    z = bf / (d + rng.normal(0, 0.7, n))
    Pc = np.stack([(uA - cx) * z / fx, (vA - cy) * z / fy, z], 1)
    Pc[rng.random(n) < 0.03] *= -1.0
    ang = -0.03
    R = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t = np.array([-0.2, 0.05, 0.1])
    Pw = ((Pc - t) @ R).astype(np.float32)
    Tcw = np.zeros((3, 4), np.float32); Tcw[:, :3] = R; Tcw[:, 3] = t
    Ow = (-(Tcw[:, :3].T @ Tcw[:, 3])).astype(np.float32)
    dist = np.linalg.norm(Pw.astype(np.float64) - Ow, axis=1)
    s64 = scale.astype(np.float64)
    maxd = dist * s64[np.clip(kr["octave"], 0, len(s64) - 1)] * rng.uniform(0.9, 1.1, n)
    maxd[rng.random(n) < 0.05] *= 0.3
    mind = maxd / s64[-1]
    view = (Pw.astype(np.float64) - Ow) / np.maximum(dist, 1e-9)[:, None]
    tilt = rng.uniform(0.0, np.deg2rad(75.0), n)
    axis = np.cross(view, rng.normal(size=(n, 3))); axis /= np.linalg.norm(axis, axis=1)[:, None]
    normal = view * np.cos(tilt)[:, None] + np.cross(axis, view) * np.sin(tilt)[:, None]
    P = WorldPointsView(world_pos=Pw, descriptors=v["dr"], max_distance=maxd.astype(np.float32), min_distance=mind.astype(np.float32),
                        normal=normal.astype(np.float32), valid=(rng.random(n) < 0.9).astype(np.uint8))
    return KF, P, Tcw, Ow, K, np.float32(bf)


def sim3_case(v, seed, K=(525.0, 525.0, 319.5, 239.5), baseline=0.4):
    """KF1 = left view, KF2 = right view of a rectified pair (disparity = fx*baseline/z), each feature carrying a MapPoint at
    the depth its disparity implies; the Sim3 handed to the matcher is the true relative pose perturbed in scale and angle."""
    from orb_slam2_b200.matcher import WorldPointsView
    rng = np.random.default_rng(seed)
    w, h = v["w"], v["h"]
    fx, fy, cx, cy = K
    scale = np.asarray(v["scale"], np.float32)
    s64 = scale.astype(np.float64)
    bounds = (0.0, 0.0, float(w), float(h))
    KF1 = FrameView(mvKeysUn=v["kl"], mDescriptors=v["dl"], mvScaleFactors=scale, bounds=bounds)
    KF2 = FrameView(mvKeysUn=v["kr"], mDescriptors=v["dr"], mvScaleFactors=scale, bounds=bounds)
    ang = 0.04
    R1 = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    t1 = np.array([0.1, -0.2, 0.3])
    T1w = np.zeros((3, 4), np.float32); T1w[:, :3] = R1; T1w[:, 3] = t1
    T2w = T1w.copy(); T2w[0, 3] -= baseline                        # camera 2 sits `baseline` to the right of camera 1

    def points(k, desc, cam_T, shift):
        n = len(k)
        d = np.maximum(v["disp"][np.clip(k["y"].astype(int), 0, h - 1), np.clip((k["x"] + shift * 0).astype(int), 0, w - 1)], 1.0)
        z = fx * baseline / (d + rng.normal(0, 0.5, n))
        Pc = np.stack([(k["x"] - cx) * z / fx, (k["y"] - cy) * z / fy, z], 1)
        Rm, tm = cam_T[:, :3].astype(np.float64), cam_T[:, 3].astype(np.float64)
        Pw = ((Pc - tm) @ Rm).astype(np.float32)
        dist = np.linalg.norm(Pc, axis=1)
        maxd = dist * s64[np.clip(k["octave"], 0, len(s64) - 1)] * rng.uniform(0.9, 1.1, n)
        maxd[rng.random(n) < 0.05] *= 0.3
        mind = maxd / s64[-1]
        return WorldPointsView(world_pos=Pw, descriptors=desc, max_distance=maxd.astype(np.float32), min_distance=mind.astype(np.float32),
                               valid=(rng.random(n) < 0.85).astype(np.uint8))
    P1 = points(v["kl"], v["dl"], T1w, 0)
    P2 = points(v["kr"], v["dr"], T2w, 0)
    s12 = np.float32(1.03)
    a = 0.004
    R12 = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    t12 = np.array([baseline, 0.01, -0.02], np.float32)              # p_c1 = s12*R12*p_c2 + t12
    sR12 = (s12 * R12).astype(np.float32)
    sR21 = ((np.float32(1.0) / s12) * R12.T).astype(np.float32)
    t21 = (-(sR21 @ t12)).astype(np.float32)
    S12 = np.concatenate([sR12, t12[:, None]], 1).astype(np.float32)
    S21 = np.concatenate([sR21, t21[:, None]], 1).astype(np.float32)
    return KF1, KF2, P1, P2, T1w, T2w, S12, S21, K
