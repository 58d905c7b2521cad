"""Host-side mirror of ORB_SLAM2::ORBmatcher (reference include/ORBmatcher.h:37-102) and ORBVocabulary over libborb.

The reference methods walk pointer graphs (Frame, KeyFrame, MapPoint); here their inputs are plain snapshots
(`FrameView`, `MapPointsView`, `KeyFrameView`) — exactly what the C++ adapter builds on the calling thread before it
calls the C ABI — and results are indices instead of MapPoint pointers.  Same method names, argument meaning and
thresholds as the reference; all compute happens in the CUDA library.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check

TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30        # src/ORBmatcher.cc:37-39


class _FrameViewC(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys_un", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p), ("occupied", C.c_void_p),
                ("min_x", C.c_float), ("min_y", C.c_float), ("max_x", C.c_float), ("max_y", C.c_float), ("n_levels", C.c_int32),
                ("scale_factors", C.c_void_p), ("resident", C.c_void_p)]


class _MapPointViewC(C.Structure):
    _fields_ = [("n", C.c_int32), ("proj_x", C.c_void_p), ("proj_y", C.c_void_p), ("proj_xr", C.c_void_p), ("level", C.c_void_p),
                ("view_cos", C.c_void_p), ("desc", C.c_void_p), ("valid", C.c_void_p), ("has_obs", C.c_void_p)]


class _LastFrameViewC(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys_un", C.c_void_p), ("world_pos", C.c_void_p), ("desc", C.c_void_p), ("valid", C.c_void_p),
                ("has_obs", C.c_void_p)]


class _WorldPointsViewC(C.Structure):
    _fields_ = [("n", C.c_int32), ("world_pos", C.c_void_p), ("desc", C.c_void_p), ("max_distance", C.c_void_p),
                ("min_distance", C.c_void_p), ("normal", C.c_void_p), ("angle", C.c_void_p), ("valid", C.c_void_p)]


class _FeatVecC(C.Structure):
    _fields_ = [("n_nodes", C.c_int32), ("node_id", C.c_void_p), ("start", C.c_void_p), ("feat_idx", C.c_void_p)]


class _KeyFrameViewC(C.Structure):
    _fields_ = [("n", C.c_int32), ("keys_un", C.c_void_p), ("desc", C.c_void_p), ("has_mp", C.c_void_p), ("u_right", C.c_void_p),
                ("fv", _FeatVecC), ("n_levels", C.c_int32), ("scale_factors", C.c_void_p), ("level_sigma2", C.c_void_p)]


def _p(a):
    return a.ctypes.data if a is not None else None


@dataclass
class FeatureVector:
    """DBoW2::FeatureVector as CSR: node ids ascending, feature indices ascending inside a node."""
    node_id: np.ndarray          # uint32 [n_nodes]
    start: np.ndarray            # int32  [n_nodes+1]
    feat_idx: np.ndarray         # uint32

    @staticmethod
    def from_nodes(node_of_feature: np.ndarray, keep: Optional[np.ndarray] = None) -> "FeatureVector":
        """fv.addFeature(nid, i_feature) for i_feature = 0..N-1 (TemplatedVocabulary.h:1160-1163)."""
        idx = np.arange(len(node_of_feature), dtype=np.uint32)
        nodes = np.asarray(node_of_feature, np.int64)
        if keep is not None:
            idx, nodes = idx[keep], nodes[keep]
        order = np.lexsort((idx, nodes))
        nodes, idx = nodes[order], idx[order]
        uniq, first = np.unique(nodes, return_index=True)
        start = np.append(first, len(nodes)).astype(np.int32)
        return FeatureVector(uniq.astype(np.uint32), start, np.ascontiguousarray(idx, np.uint32))

    def as_dict(self) -> Dict[int, List[int]]:
        return {int(n): self.feat_idx[self.start[i]:self.start[i + 1]].tolist() for i, n in enumerate(self.node_id)}


@dataclass
class FrameView:
    """What SearchByProjection reads of a Frame (include/Frame.h)."""
    mvKeysUn: np.ndarray
    mDescriptors: np.ndarray
    mvScaleFactors: np.ndarray
    bounds: Tuple[float, float, float, float]            # mnMinX, mnMinY, mnMaxX, mnMaxY
    mvuRight: Optional[np.ndarray] = None
    occupied: Optional[np.ndarray] = None                 # mvpMapPoints[i] && Observations()>0 (or != NULL, per overload)
    mfLogScaleFactor: Optional[float] = None              # Frame::mfLogScaleFactor; default logf(mvScaleFactors[1])
    mvInvLevelSigma2: Optional[np.ndarray] = None         # only read by Fuse(pKF, vpMapPoints, th)
    resident: Optional["ResidentFrame"] = None            # device-resident copy (borb_frame): only `occupied` travels per call

    def _view(self, with_ur: bool = True):
        """(ctypes view, arrays to keep alive).  With a resident frame only `occupied` is read from the host."""
        oc = np.ascontiguousarray(self.occupied, np.uint8) if self.occupied is not None else None
        sf = np.ascontiguousarray(self.mvScaleFactors, np.float32)
        if self.resident is not None:
            return _FrameViewC(0, None, None, None, _p(oc), 0.0, 0.0, 0.0, 0.0, len(sf), None, self.resident._h), [oc, sf]
        k = np.ascontiguousarray(self.mvKeysUn, KP_DTYPE); d = np.ascontiguousarray(self.mDescriptors, np.uint8)
        ur = np.ascontiguousarray(self.mvuRight, np.float32) if (with_ur and self.mvuRight is not None) else None
        return _FrameViewC(len(k), _p(k), _p(d), _p(ur), _p(oc), *[float(x) for x in self.bounds], len(sf), _p(sf), None), [k, d, ur, oc, sf]

    def make_resident(self, matcher: "ORBmatcher") -> "FrameView":
        """Uploads the frame once (borb_frame_create: keypoints, descriptors, mvuRight, feature grid) and returns a view that
        refers to the device copy."""
        import dataclasses
        return dataclasses.replace(self, resident=ResidentFrame(matcher, self))


class ResidentFrame:
    """borb_frame: a Frame's features and 64x48 grid kept in HBM across the matcher calls of one Track()."""

    def __init__(self, matcher: "ORBmatcher", F: "FrameView"):
        self._lib = _lib.load()
        fv, keep = FrameView._view(dataclass_replace_resident(F), True)
        h = C.c_void_p()
        check(self._lib.borb_frame_create(matcher._h, C.byref(fv), C.byref(h)), "borb_frame_create")
        self._h = h
        self.n = len(F.mvKeysUn)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.borb_frame_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _CameraC(C.Structure):
    _fields_ = [(n, C.c_float) for n in ("fx", "fy", "cx", "cy", "k1", "k2", "p1", "p2", "k3", "bf")]


def frames_from_extractor(matcher: "ORBmatcher", extractor, images, n_keys, K, dist=(0, 0, 0, 0, 0), bf: float = 0.0, mode: int = 0,
                          depth=None, depth_factor: float = 1.0, want_host: bool = True):
    """borb_frames_from_extractor: the Frame constructor tail (UndistortKeyPoints, ComputeStereoFromRGBD, AssignFeaturesToGrid,
    src/Frame.cc:404-434,643-664,230-245) on the device for images of the extractor's last batch.  K = (fx, fy, cx, cy),
    dist = (k1, k2, p1, p2, k3); mode 0 mono / 1 stereo / 2 RGB-D with depth = list of (h,w) float32 (metres) or uint16 (raw) maps.
    Returns (frames, host) where frames[i] is a FrameView bound to the resident frame and host = dict(keys_un, u_right, depth, bounds)."""
    lib = _lib.load()
    images = np.ascontiguousarray(images, np.int32); nk = np.ascontiguousarray(n_keys, np.int32)
    nf = len(images)
    d5 = list(dist) + [0.0] * (5 - len(dist))
    cam = _CameraC(*[float(x) for x in K], *[float(x) for x in d5], float(bf))
    cap = int(nk.max()) if nf else 0
    ku = np.zeros((nf, max(cap, 1)), KP_DTYPE); ur = np.full((nf, max(cap, 1)), -1, np.float32); dp = np.full((nf, max(cap, 1)), -1, np.float32)
    b4 = np.zeros(4, np.float32)
    handles = (C.c_void_p * max(nf, 1))()
    dptr, dtype_flag, stride, keep = None, 0, 0, []
    if mode == 2:
        keep = [np.ascontiguousarray(d) for d in depth]
        dtype_flag = 1 if keep[0].dtype == np.uint16 else 0
        keep = [np.ascontiguousarray(d, np.uint16 if dtype_flag else np.float32) for d in keep]
        stride = keep[0].strides[0]
        dptr = (C.c_void_p * nf)(*[d.ctypes.data for d in keep])
    check(lib.borb_frames_from_extractor(matcher._h, extractor._h, _p(images), nf, _p(nk), C.byref(cam), int(mode), dptr, dtype_flag,
                                         float(np.float32(depth_factor)), int(stride), _p(ku) if want_host else None,
                                         _p(ur) if want_host else None, _p(dp) if want_host else None, cap, _p(b4), handles),
          "borb_frames_from_extractor")
    sf = extractor.GetScaleFactors()
    out = []
    for i in range(nf):
        rf = ResidentFrame.__new__(ResidentFrame)
        rf._lib, rf._h, rf.n = lib, C.c_void_p(handles[i]), int(nk[i])
        n = int(nk[i])
        out.append(FrameView(mvKeysUn=ku[i, :n], mDescriptors=np.zeros((n, 32), np.uint8), mvScaleFactors=sf, bounds=tuple(float(x) for x in b4),
                             mvuRight=ur[i, :n] if mode else None, resident=rf))
    return out, dict(keys_un=[ku[i, :nk[i]] for i in range(nf)], u_right=[ur[i, :nk[i]] for i in range(nf)],
                     depth=[dp[i, :nk[i]] for i in range(nf)], bounds=b4)


def dataclass_replace_resident(F):
    import dataclasses
    return dataclasses.replace(F, resident=None)


@dataclass
class MapPointsView:
    """Local map points after Frame::isInFrustum (src/Frame.cc:269-325), in vpMapPoints order."""
    mTrackProjX: np.ndarray
    mTrackProjY: np.ndarray
    mTrackProjXR: np.ndarray
    mnTrackScaleLevel: np.ndarray
    mTrackViewCos: np.ndarray
    descriptors: np.ndarray
    valid: Optional[np.ndarray] = None                    # mbTrackInView && !isBad()
    has_obs: Optional[np.ndarray] = None                  # Observations()>0


@dataclass
class WorldPointsView:
    """MapPoints with world-frame data for SearchByProjection(CurrentFrame, KeyFrame, ...) and (KeyFrame, Scw, ...)."""
    world_pos: np.ndarray                                  # (n,3) float32, GetWorldPos()
    descriptors: np.ndarray                                # (n,32), GetDescriptor()
    max_distance: np.ndarray                               # mfMaxDistance
    min_distance: np.ndarray                               # mfMinDistance
    normal: Optional[np.ndarray] = None                    # (n,3) GetNormal() — Sim3 overload
    angle: Optional[np.ndarray] = None                     # pKF->mvKeysUn[i].angle — keyframe overload, orientation check
    valid: Optional[np.ndarray] = None


def _libm_logf(x: float) -> float:
    """glibc logf — what Frame::mfLogScaleFactor = log(mfScaleFactor) evaluates to (src/Frame.cc:71)."""
    libm = C.CDLL("libm.so.6")
    libm.logf.restype = C.c_float
    libm.logf.argtypes = [C.c_float]
    return float(libm.logf(float(np.float32(x))))


@dataclass
class LastFrameView:
    """What SearchByProjection(CurrentFrame, LastFrame, ...) reads of LastFrame and of its MapPoints."""
    mvKeysUn: np.ndarray
    world_pos: np.ndarray                                  # (N,3) float32, pMP->GetWorldPos()
    descriptors: np.ndarray                                # (N,32), pMP->GetDescriptor()
    valid: Optional[np.ndarray] = None                     # mvpMapPoints[i] && !mvbOutlier[i]
    has_obs: Optional[np.ndarray] = None                   # Observations()>0


@dataclass
class KeyFrameView:
    """What the BoW-guided searches read of a KeyFrame / Frame."""
    mvKeysUn: np.ndarray
    mDescriptors: np.ndarray
    mFeatVec: FeatureVector
    has_mp: Optional[np.ndarray] = None                   # MapPoint present && !isBad(), per feature
    mvuRight: Optional[np.ndarray] = None
    mvScaleFactors: Optional[np.ndarray] = None
    mvLevelSigma2: Optional[np.ndarray] = None
    _keep: list = field(default_factory=list, repr=False)

    def _c(self) -> _KeyFrameViewC:
        k = np.ascontiguousarray(self.mvKeysUn, KP_DTYPE)
        d = np.ascontiguousarray(self.mDescriptors, np.uint8)
        hm = np.ascontiguousarray(self.has_mp, np.uint8) if self.has_mp is not None else None
        ur = np.ascontiguousarray(self.mvuRight, np.float32) if self.mvuRight is not None else None
        nd = np.ascontiguousarray(self.mFeatVec.node_id, np.uint32)
        st = np.ascontiguousarray(self.mFeatVec.start, np.int32)
        fi = np.ascontiguousarray(self.mFeatVec.feat_idx, np.uint32)
        sf = np.ascontiguousarray(self.mvScaleFactors, np.float32) if self.mvScaleFactors is not None else None
        sg = np.ascontiguousarray(self.mvLevelSigma2, np.float32) if self.mvLevelSigma2 is not None else None
        self._keep = [k, d, hm, ur, nd, st, fi, sf, sg]
        nl = len(sf) if sf is not None else (len(sg) if sg is not None else 0)
        return _KeyFrameViewC(len(k), _p(k), _p(d), _p(hm), _p(ur), _FeatVecC(len(nd), _p(nd), _p(st), _p(fi)), nl, _p(sf), _p(sg))


class ORBmatcher:
    """ORBmatcher(nnratio=0.6, checkOri=True) — include/ORBmatcher.h:41."""

    TH_LOW, TH_HIGH, HISTO_LENGTH = TH_LOW, TH_HIGH, HISTO_LENGTH

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, device: int = 0):
        self._lib = _lib.load()
        self.mfNNratio = float(np.float32(nnratio))
        self.mbCheckOrientation = bool(checkOri)
        h = C.c_void_p()
        check(self._lib.borb_matcher_create(device, C.byref(h)), "borb_matcher_create")
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.borb_matcher_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def DescriptorDistance(a: np.ndarray, b: np.ndarray) -> int:
        """256-bit Hamming distance (src/ORBmatcher.cc:1647-1663) — host convenience for tests."""
        return int(np.unpackbits(np.bitwise_xor(np.asarray(a, np.uint8), np.asarray(b, np.uint8))).sum())

    def SearchByProjection(self, F: FrameView, mps: MapPointsView, th: float = 3.0) -> Tuple[int, np.ndarray]:
        """src/ORBmatcher.cc:45-129.  Returns (nmatches, match_feat[n_mp]): frame feature that received map point i, or -1."""
        fv, keep = F._view()
        px = np.ascontiguousarray(mps.mTrackProjX, np.float32); py = np.ascontiguousarray(mps.mTrackProjY, np.float32)
        pxr = np.ascontiguousarray(mps.mTrackProjXR, np.float32); lv = np.ascontiguousarray(mps.mnTrackScaleLevel, np.int32)
        vc = np.ascontiguousarray(mps.mTrackViewCos, np.float32); md = np.ascontiguousarray(mps.descriptors, np.uint8)
        va = np.ascontiguousarray(mps.valid, np.uint8) if mps.valid is not None else None
        ho = np.ascontiguousarray(mps.has_obs, np.uint8) if mps.has_obs is not None else None
        mv = _MapPointViewC(len(px), _p(px), _p(py), _p(pxr), _p(lv), _p(vc), _p(md), _p(va), _p(ho))
        match = np.full(max(len(px), 1), -1, np.int32)
        n = C.c_int32(0)
        check(self._lib.borb_search_by_projection(self._h, C.byref(fv), C.byref(mv), float(th), self.mfNNratio, _p(match), C.byref(n)),
              "borb_search_by_projection")
        return n.value, match[:len(px)]

    def SearchByProjectionBatch(self, frames: Sequence[FrameView], mps_list: Sequence[MapPointsView], th: float = 3.0):
        """borb_search_by_projection_batch: SearchByProjection(F, vpMapPoints, th) of many independent device-resident frames in one
        launch pair.  Returns [(nmatches, match_feat)] per job, each equal to the single call's result."""
        n = len(frames)
        assert n == len(mps_list)
        FV = (_FrameViewC * n)()
        MV = (_MapPointViewC * n)()
        keep, outs = [], []
        for j, (F, mps) in enumerate(zip(frames, mps_list)):
            fv, k = F._view()
            FV[j] = fv
            px = np.ascontiguousarray(mps.mTrackProjX, np.float32); py = np.ascontiguousarray(mps.mTrackProjY, np.float32)
            pxr = np.ascontiguousarray(mps.mTrackProjXR, np.float32); lv = np.ascontiguousarray(mps.mnTrackScaleLevel, np.int32)
            vc = np.ascontiguousarray(mps.mTrackViewCos, np.float32); md = np.ascontiguousarray(mps.descriptors, np.uint8)
            va = np.ascontiguousarray(mps.valid, np.uint8) if mps.valid is not None else None
            ho = np.ascontiguousarray(mps.has_obs, np.uint8) if mps.has_obs is not None else None
            MV[j] = _MapPointViewC(len(px), _p(px), _p(py), _p(pxr), _p(lv), _p(vc), _p(md), _p(va), _p(ho))
            out = np.full(max(len(px), 1), -1, np.int32)
            keep.append((k, px, py, pxr, lv, vc, md, va, ho)); outs.append(out)
        ptrs = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        nm = np.zeros(max(n, 1), np.int32)
        check(self._lib.borb_search_by_projection_batch(self._h, FV, MV, n, float(th), self.mfNNratio, ptrs, _p(nm)), "borb_search_by_projection_batch")
        return [(int(nm[j]), outs[j][:MV[j].n]) for j in range(n)]

    def SearchByProjectionLast(self, Cur: FrameView, Last: LastFrameView, Tcw: np.ndarray, K: Tuple[float, float, float, float], bf: float,
                               th: float, bForward: bool = False, bBackward: bool = False) -> Tuple[int, np.ndarray]:
        """SearchByProjection(CurrentFrame, LastFrame, th, bMono) — src/ORBmatcher.cc:1328-1470.  Tcw: (3,4) or (4,4) current pose;
        K = (fx, fy, cx, cy).  Returns (nmatches, state[cur.N]): >=0 last-frame index now matched, -1 untouched, -2 culled."""
        fv, keep = Cur._view()
        k = Cur.mvKeysUn
        lk = np.ascontiguousarray(Last.mvKeysUn, KP_DTYPE); wp = np.ascontiguousarray(Last.world_pos, np.float32)
        ld = np.ascontiguousarray(Last.descriptors, np.uint8)
        va = np.ascontiguousarray(Last.valid, np.uint8) if Last.valid is not None else None
        ho = np.ascontiguousarray(Last.has_obs, np.uint8) if Last.has_obs is not None else None
        lv = _LastFrameViewC(len(lk), _p(lk), _p(wp), _p(ld), _p(va), _p(ho))
        T = np.ascontiguousarray(np.asarray(Tcw, np.float32)[:3, :4]).reshape(12)
        state = np.full(max(len(k), 1), -1, np.int32)
        n = C.c_int32(0)
        check(self._lib.borb_search_by_projection_last(self._h, C.byref(fv), C.byref(lv), _p(T), float(K[0]), float(K[1]), float(K[2]), float(K[3]),
                                                       float(bf), float(th), int(bForward), int(bBackward), int(self.mbCheckOrientation),
                                                       _p(state), C.byref(n)), "borb_search_by_projection_last")
        return n.value, state[:len(k)]

    def _points_call(self, F: FrameView, P: WorldPointsView, with_stereo: bool = False):
        fv, keep = F._view(with_stereo)
        k = F.mvKeysUn
        sf = np.ascontiguousarray(F.mvScaleFactors, np.float32)
        arrs = []
        for a, dt in ((P.world_pos, np.float32), (P.descriptors, np.uint8), (P.max_distance, np.float32), (P.min_distance, np.float32),
                      (P.normal, np.float32), (P.angle, np.float32), (P.valid, np.uint8)):
            arrs.append(np.ascontiguousarray(a, dt) if a is not None else None)
        pv = _WorldPointsViewC(len(arrs[0]), *[_p(a) for a in arrs])
        logs = F.mfLogScaleFactor if F.mfLogScaleFactor is not None else _libm_logf(sf[1] if len(sf) > 1 else 1.2)
        return fv, pv, float(logs), len(k), keep + arrs

    def SearchByProjectionKF(self, Cur: FrameView, P: WorldPointsView, Tcw: np.ndarray, Ow: np.ndarray, K: Tuple[float, float, float, float],
                             th: float, ORBdist: int) -> Tuple[int, np.ndarray]:
        """SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) — src/ORBmatcher.cc:1472-1599 (relocalisation).
        P = the keyframe's MapPoints (valid = present, not bad, not already found); Cur.occupied = mvpMapPoints[i] != NULL;
        Ow = -Rcw.t()*tcw.  Returns (nmatches, state[cur.N]): >=0 index into P, -1 untouched, -2 culled by orientation."""
        fv, pv, logs, n, keep = self._points_call(Cur, P)
        T = np.ascontiguousarray(np.asarray(Tcw, np.float32)[:3, :4]).reshape(12)
        ow = np.ascontiguousarray(np.asarray(Ow, np.float32).reshape(3))
        state = np.full(max(n, 1), -1, np.int32)
        nm = C.c_int32(0)
        check(self._lib.borb_search_by_projection_kf(self._h, C.byref(fv), C.byref(pv), _p(T), _p(ow), float(K[0]), float(K[1]), float(K[2]),
                                                     float(K[3]), logs, float(th), int(ORBdist), int(self.mbCheckOrientation), _p(state),
                                                     C.byref(nm)), "borb_search_by_projection_kf")
        return nm.value, state[:n]

    def SearchByProjectionSim3(self, pKF: FrameView, P: WorldPointsView, Tcw: np.ndarray, Ow: np.ndarray,
                               K: Tuple[float, float, float, float], th: int) -> Tuple[int, np.ndarray]:
        """SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) — src/ORBmatcher.cc:290-403 (loop closing).  Tcw = [Rcw|tcw] with
        the Sim3 scale divided out, Ow = -Rcw.t()*tcw; pKF.occupied = vpMatched[idx] != NULL.  Returns (nmatches, state[kf.N])."""
        fv, pv, logs, n, keep = self._points_call(pKF, P)
        T = np.ascontiguousarray(np.asarray(Tcw, np.float32)[:3, :4]).reshape(12)
        ow = np.ascontiguousarray(np.asarray(Ow, np.float32).reshape(3))
        state = np.full(max(n, 1), -1, np.int32)
        nm = C.c_int32(0)
        check(self._lib.borb_search_by_projection_sim3(self._h, C.byref(fv), C.byref(pv), _p(T), _p(ow), float(K[0]), float(K[1]), float(K[2]),
                                                       float(K[3]), logs, int(th), _p(state), C.byref(nm)), "borb_search_by_projection_sim3")
        return nm.value, state[:n]

    def SearchLocalPoints(self, F: FrameView, P: WorldPointsView, Tcw: np.ndarray, Ow: np.ndarray, K: Tuple[float, float, float, float],
                          mbf: float, th: float = 1.0, has_obs: Optional[np.ndarray] = None, viewingCosLimit: float = 0.5):
        """Tracking::SearchLocalPoints (src/Tracking.cc:1148-1194): Frame::isInFrustum for every point of P, then
        SearchByProjection(F, vpMapPoints, th) on those in view, in one call.  Returns a dict with in_view, the MapPoint track
        fields (proj_x, proj_y, proj_xr, level, view_cos), match (feature per point or -1) and nmatches."""
        fv, pv, logs, n, keep = self._points_call(F, P, with_stereo=True)
        T = np.ascontiguousarray(np.asarray(Tcw, np.float32)[:3, :4]).reshape(12)
        ow = np.ascontiguousarray(np.asarray(Ow, np.float32).reshape(3))
        ho = np.ascontiguousarray(has_obs, np.uint8) if has_obs is not None else None
        nq = len(P.world_pos)
        m1 = max(nq, 1)
        out = dict(in_view=np.zeros(m1, np.uint8), proj_x=np.zeros(m1, np.float32), proj_y=np.zeros(m1, np.float32),
                   proj_xr=np.zeros(m1, np.float32), level=np.zeros(m1, np.int32), view_cos=np.zeros(m1, np.float32),
                   match=np.full(m1, -1, np.int32))
        nm = C.c_int32(0)
        check(self._lib.borb_search_local_points(self._h, C.byref(fv), C.byref(pv), _p(ho), _p(T), _p(ow), float(K[0]), float(K[1]), float(K[2]),
                                                 float(K[3]), float(mbf), float(viewingCosLimit), logs, float(th), self.mfNNratio,
                                                 _p(out["in_view"]), _p(out["proj_x"]), _p(out["proj_y"]), _p(out["proj_xr"]), _p(out["level"]),
                                                 _p(out["view_cos"]), _p(out["match"]), C.byref(nm)), "borb_search_local_points")
        out = {k: v[:nq] for k, v in out.items()}
        out["nmatches"] = nm.value
        return out

    def SearchForInitialization(self, F1: FrameView, F2: FrameView, vbPrevMatched: np.ndarray, windowSize: int = 10):
        """SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize) — src/ORBmatcher.cc:405-520.
        Returns (nmatches, vnMatches12[F1.N], updated vbPrevMatched (N,2))."""
        views, keep = [], []
        for F in (F1, F2):
            k = np.ascontiguousarray(F.mvKeysUn, KP_DTYPE); d = np.ascontiguousarray(F.mDescriptors, np.uint8)
            sf = np.ascontiguousarray(F.mvScaleFactors, np.float32)
            views.append(_FrameViewC(len(k), _p(k), _p(d), None, None, *[float(x) for x in F.bounds], len(sf), _p(sf)))
            keep += [k, d, sf]
        n1 = views[0].n
        prev = np.ascontiguousarray(np.asarray(vbPrevMatched, np.float32).reshape(-1, 2)).copy()
        if len(prev) == 0:
            prev = np.zeros((1, 2), np.float32)
        m12 = np.full(max(n1, 1), -1, np.int32)
        nm = C.c_int32(0)
        check(self._lib.borb_search_for_initialization(self._h, C.byref(views[0]), C.byref(views[1]), _p(prev), int(windowSize), self.mfNNratio,
                                                       int(self.mbCheckOrientation), _p(m12), C.byref(nm)), "borb_search_for_initialization")
        return nm.value, m12[:n1], prev[:n1]

    def ComputeDistinctiveDescriptors(self, groups) -> np.ndarray:
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:242-307) for a batch of MapPoints: groups[p] = (N_p,32)
        uint8 descriptors of the point's observations.  Returns best[p] = index of the medoid-by-median descriptor (-1 if empty)."""
        groups = [np.ascontiguousarray(g, np.uint8).reshape(-1, 32) for g in groups]
        off = np.zeros(len(groups) + 1, np.int32)
        off[1:] = np.cumsum([len(g) for g in groups])
        desc = np.ascontiguousarray(np.concatenate(groups, 0)) if len(groups) and off[-1] > 0 else np.zeros((1, 32), np.uint8)
        best = np.full(max(len(groups), 1), -1, np.int32)
        check(self._lib.borb_distinctive_descriptors(self._h, _p(desc), _p(off), len(groups), _p(best)), "borb_distinctive_descriptors")
        return best[:len(groups)]

    def Fuse(self, pKF: FrameView, P: WorldPointsView, Tcw: np.ndarray, Ow: np.ndarray, K: Tuple[float, float, float, float], bf: float,
             th: float = 3.0, Scw: bool = False) -> Tuple[int, np.ndarray]:
        """Search part of Fuse(pKF, vpMapPoints, th) — src/ORBmatcher.cc:825-970 — or, with Scw=True, of
        Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) — :972-1100.  Returns (n_found, best_idx[len(P)]); the MapPoint
        bookkeeping (Replace / AddObservation / vpReplacePoint) is the caller's, applied in order."""
        fv, pv, logs, n, keep = self._points_call(pKF, P, with_stereo=not Scw)
        inv = np.ascontiguousarray(pKF.mvInvLevelSigma2, np.float32) if pKF.mvInvLevelSigma2 is not None else None
        T = np.ascontiguousarray(np.asarray(Tcw, np.float32)[:3, :4]).reshape(12)
        ow = np.ascontiguousarray(np.asarray(Ow, np.float32).reshape(3))
        nq = len(P.world_pos)
        best = np.full(max(nq, 1), -1, np.int32)
        nf = C.c_int32(0)
        check(self._lib.borb_fuse(self._h, C.byref(fv), _p(inv), C.byref(pv), _p(T), _p(ow), float(K[0]), float(K[1]), float(K[2]), float(K[3]),
                                  float(bf), logs, float(th), int(Scw), _p(best), C.byref(nf)), "borb_fuse")
        return nf.value, best[:nq]

    def SearchBySim3(self, pKF1: FrameView, pKF2: FrameView, P1: WorldPointsView, P2: WorldPointsView, T1w: np.ndarray, T2w: np.ndarray,
                     S12: np.ndarray, S21: np.ndarray, K: Tuple[float, float, float, float], th: float) -> Tuple[int, np.ndarray]:
        """SearchBySim3(pKF1, pKF2, vpMatches12, s12, R12, t12, th) — src/ORBmatcher.cc:1102-1326.  S12 = [s12*R12 | t12],
        S21 = [(1/s12)*R12^T | -sR21*t12] (3x4).  Returns (nFound, match12[kf1.N]): index in KF2 or -1."""
        fv1, pv1, logs1, n1, keep1 = self._points_call(pKF1, P1)
        fv2, pv2, logs2, n2, keep2 = self._points_call(pKF2, P2)
        mats = [np.ascontiguousarray(np.asarray(Mx, np.float32)[:3, :4]).reshape(12) for Mx in (T1w, T2w, S12, S21)]
        match = np.full(max(n1, 1), -1, np.int32)
        nf = C.c_int32(0)
        check(self._lib.borb_search_by_sim3(self._h, C.byref(fv1), C.byref(fv2), C.byref(pv1), C.byref(pv2), _p(mats[0]), _p(mats[1]),
                                            _p(mats[2]), _p(mats[3]), float(K[0]), float(K[1]), float(K[2]), float(K[3]), logs1, logs2,
                                            float(th), _p(match), C.byref(nf)), "borb_search_by_sim3")
        return nf.value, match[:n1]

    def SearchByBoW(self, pKF, F: KeyFrameView):
        """SearchByBoW(KeyFrame*, Frame&, vpMapPointMatches) — src/ORBmatcher.cc:159-288.  pKF may be one KeyFrameView or a
        sequence (batched candidates).  Returns (nmatches, match[F.N]) or lists of them: match[j] = keyframe feature whose
        MapPoint frame feature j received, or -1."""
        single = isinstance(pKF, KeyFrameView)
        kfs = [pKF] if single else list(pKF)
        arr = (_KeyFrameViewC * len(kfs))(*[kf._c() for kf in kfs])
        fc = F._c()
        nF = len(F.mvKeysUn)
        match = np.full((len(kfs), max(nF, 1)), -1, np.int32)
        nm = np.zeros(len(kfs), np.int32)
        check(self._lib.borb_search_by_bow(self._h, arr, len(kfs), C.byref(fc), self.mfNNratio, int(self.mbCheckOrientation), _p(match), _p(nm)),
              "borb_search_by_bow")
        match = match[:, :nF]
        return (int(nm[0]), match[0]) if single else (nm, match)

    def SearchByBoW_KF(self, pKF1: KeyFrameView, pKF2: KeyFrameView) -> Tuple[int, np.ndarray]:
        """SearchByBoW(KeyFrame*, KeyFrame*, vpMatches12) — src/ORBmatcher.cc:522-655.  match12[i] = index in KF2 or -1."""
        c1, c2 = pKF1._c(), pKF2._c()
        n1 = len(pKF1.mvKeysUn)
        match = np.full(max(n1, 1), -1, np.int32)
        nm = C.c_int32(0)
        check(self._lib.borb_search_by_bow_kf(self._h, C.byref(c1), C.byref(c2), self.mfNNratio, int(self.mbCheckOrientation), _p(match), C.byref(nm)),
              "borb_search_by_bow_kf")
        return nm.value, match[:n1]

    def SearchForTriangulation(self, pKF1: KeyFrameView, pKF2: KeyFrameView, F12: np.ndarray, epipole: Tuple[float, float],
                               bOnlyStereo: bool = False) -> np.ndarray:
        """src/ORBmatcher.cc:657-823.  Returns vMatchedPairs as an (m,2) int array (idx1, idx2), ascending idx1."""
        c1, c2 = pKF1._c(), pKF2._c()
        f = np.ascontiguousarray(F12, np.float32).reshape(9)
        cap = max(len(pKF1.mvKeysUn), 1)
        pairs = np.zeros((cap, 2), np.int32)
        n = C.c_int32(0)
        check(self._lib.borb_search_for_triangulation(self._h, C.byref(c1), C.byref(c2), _p(f), float(epipole[0]), float(epipole[1]),
                                                      int(bOnlyStereo), int(self.mbCheckOrientation), _p(pairs), cap, C.byref(n)),
              "borb_search_for_triangulation")
        return pairs[:n.value]


def _sharing_order(cw, fw, seq, skip=()):
    """lKFsSharingWords: keyframes sharing a word with the query, in the order the inverted-file walk meets them — by first
    shared word id, then by insertion into that word's list (src/KeyFrameDatabase.cc:86-108, :211-224)."""
    s = [int(i) for i in np.nonzero(np.asarray(cw) > 0)[0] if int(i) not in skip]
    return sorted(s, key=lambda i: (int(fw[i]), seq[i]))


def relocalization_candidates(cw, sc, fw, seq, covisibility, reloc_score: Optional[dict] = None) -> list:
    """The host part of KeyFrameDatabase::DetectRelocalizationCandidates (src/KeyFrameDatabase.cc:226-310) from the per-keyframe
    shared-word counts `cw`, float L1 scores `sc` and first shared words `fw` (borb_kfdb_query).
    `reloc_score` is the database's persistent {slot: KeyFrame::mRelocScore}: the reference assigns mRelocScore only to keyframes
    above minCommonWords (:236-243) and the covisibility accumulation (:262-275) reads the field of EVERY neighbour that shares a
    word with the query — for a neighbour below the threshold that is the value an EARLIER query left there.  Passing the same
    dict to successive queries reproduces that; None (or a fresh dict) is a fresh database, where the field is 0."""
    if reloc_score is None:
        reloc_score = {}
    sharing = _sharing_order(cw, fw, seq)
    if not sharing:
        return []
    maxCommonWords = max(int(cw[s]) for s in sharing)
    minCommonWords = int(np.float32(maxCommonWords) * np.float32(0.8))
    scored = [(np.float32(sc[s]), s) for s in sharing if cw[s] > minCommonWords]
    for si, s in scored:
        reloc_score[s] = si                               # pKFi->mRelocScore = si (:241)
    if not scored:
        return []
    acc, bestAcc = [], np.float32(0)
    for si, s in scored:
        bestScore, accScore, best = si, si, s
        for s2 in covisibility(s):
            if cw[s2] <= 0:
                continue                                  # mnRelocQuery != F->mnId: shares no word with the query
            r = np.float32(reloc_score.get(s2, 0.0))      # pKF2->mRelocScore: this query's score, or the stale one (see above)
            accScore = np.float32(accScore + r)
            if r > bestScore:
                best, bestScore = s2, r
        acc.append((accScore, best))
        if accScore > bestAcc:
            bestAcc = accScore
    minScoreToRetain = np.float32(0.75) * bestAcc
    out, seen = [], set()
    for a, s in acc:
        if a > minScoreToRetain and s not in seen:
            out.append(s); seen.add(s)
    return out


def loop_candidates(cw, sc, fw, seq, connected, covisibility, minScore) -> list:
    """The host part of KeyFrameDatabase::DetectLoopCandidates (src/KeyFrameDatabase.cc:110-197); `connected` = slots of the
    query keyframe's connected keyframes (never candidates, :96-101)."""
    minScore = np.float32(minScore)
    sharing = _sharing_order(cw, fw, seq, skip=connected)
    if not sharing:
        return []
    maxCommonWords = max(int(cw[s]) for s in sharing)
    minCommonWords = int(np.float32(maxCommonWords) * np.float32(0.8))
    in_list = set(sharing)
    scored = [(np.float32(sc[s]), s) for s in sharing if cw[s] > minCommonWords and np.float32(sc[s]) >= minScore]
    if not scored:
        return []
    acc, bestAcc = [], minScore
    for si, s in scored:
        bestScore, accScore, best = si, si, s
        for s2 in covisibility(s):
            if s2 in in_list and cw[s2] > minCommonWords:                 # mnLoopQuery == pKF->mnId && mnLoopWords > minCommonWords (:157)
                r = np.float32(sc[s2])
                accScore = np.float32(accScore + r)
                if r > bestScore:
                    best, bestScore = s2, r
        acc.append((accScore, best))
        if accScore > bestAcc:
            bestAcc = accScore
    minScoreToRetain = np.float32(0.75) * bestAcc
    out, seen = [], set()
    for a, s in acc:
        if a > minScoreToRetain and s not in seen:
            out.append(s); seen.add(s)
    return out


def bow_and_featvec(word, weight, node):
    """The bookkeeping half of TemplatedVocabulary::transform (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1150-1194) from the
    per-feature results of the tree descent (word id, word weight, node id at level L-levelsup): BowVector::addWeight in
    feature order for every feature whose word weight is > 0 ("not stopped"), L1 normalisation with the norm summed in map
    (word id) order, FeatureVector::addFeature in feature order."""
    word = np.asarray(word); weight = np.asarray(weight, np.float64); node = np.asarray(node)
    bow: Dict[int, float] = {}
    keep = weight > 0
    for i in np.nonzero(keep)[0]:
        w = int(word[i])
        bow[w] = bow.get(w, 0.0) + float(weight[i])
    norm = 0.0                                   # plain left-to-right accumulation: Python >= 3.12's sum() is compensated, the
    for _, v in sorted(bow.items()):             # reference's `norm += fabs(it->second)` loop (BowVector.cpp:67-70) is not
        norm += abs(v)
    if norm > 0.0:
        bow = {k: v / norm for k, v in bow.items()}
    return dict(sorted(bow.items())), FeatureVector.from_nodes(node, keep)


class KeyFrameDatabase:
    """KeyFrameDatabase (include/KeyFrameDatabase.h) with the keyframes resident in HBM.  add/erase/clear mirror
    src/KeyFrameDatabase.cc:41-73; query() is the data-parallel part of DetectLoopCandidates / DetectRelocalizationCandidates
    (shared-word count + L1 score for every keyframe in one launch); DetectRelocalizationCandidates() finishes the
    reference's procedure on the host from those arrays and the caller's covisibility lists."""

    def __init__(self, matcher: "ORBmatcher", device: int = 0):
        self._lib = _lib.load()
        self._m = matcher
        h = C.c_void_p()
        check(self._lib.borb_kfdb_create(device, C.byref(h)), "borb_kfdb_create")
        self._h = h
        self._seq = []                  # insertion sequence number per slot (inverted-file list order)
        self._reloc_score = {}          # KeyFrame::mRelocScore per slot, persistent across queries as in the reference

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.borb_kfdb_destroy(self._h)
            self._h = None

    @staticmethod
    def _bow_arrays(bow):
        w = np.fromiter(bow.keys(), np.uint32, len(bow)); v = np.fromiter(bow.values(), np.float64, len(bow))
        o = np.argsort(w, kind="stable")
        return np.ascontiguousarray(w[o]), np.ascontiguousarray(v[o])

    def add(self, pKF: KeyFrameView, mBowVec: Dict[int, float]) -> int:
        w, v = self._bow_arrays(mBowVec)
        slot = C.c_int32(-1)
        kc = pKF._c()
        check(self._lib.borb_kfdb_add(self._h, C.byref(kc), _p(w), _p(v), len(w), C.byref(slot)), "borb_kfdb_add")
        self._seq.append(len(self._seq))
        return slot.value

    def erase(self, slot: int) -> None:
        check(self._lib.borb_kfdb_erase(self._h, int(slot)), "borb_kfdb_erase")
        self._reloc_score.pop(int(slot), None)

    def clear(self) -> None:
        check(self._lib.borb_kfdb_clear(self._h), "borb_kfdb_clear")
        self._seq = []
        self._reloc_score = {}

    def set_has_mp(self, slot: int, has_mp: np.ndarray) -> None:
        hm = np.ascontiguousarray(has_mp, np.uint8)
        check(self._lib.borb_kfdb_set_has_mp(self._h, int(slot), _p(hm)), "borb_kfdb_set_has_mp")

    def size(self) -> Tuple[int, int]:
        n = C.c_int32(0); b = C.c_uint64(0)
        check(self._lib.borb_kfdb_size(self._h, C.byref(n), C.byref(b)), "borb_kfdb_size")
        return n.value, b.value

    def query(self, mBowVec: Dict[int, float]):
        """Returns (common_words[int32], score[float32], first_word[uint32]) with one entry per slot."""
        w, v = self._bow_arrays(mBowVec)
        n = self.size()[0]
        cw = np.zeros(max(n, 1), np.int32); sc = np.zeros(max(n, 1), np.float32); fw = np.zeros(max(n, 1), np.uint32)
        ns = C.c_int32(0)
        check(self._lib.borb_kfdb_query(self._m._h, self._h, _p(w), _p(v), len(w), _p(cw), _p(sc), _p(fw), len(cw), C.byref(ns)),
              "borb_kfdb_query")
        return cw[:n], sc[:n], fw[:n]

    def DetectRelocalizationCandidates(self, mBowVec: Dict[int, float], covisibility) -> list:
        """src/KeyFrameDatabase.cc:199-310.  covisibility(slot) -> up to 10 slots (GetBestCovisibilityKeyFrames(10))."""
        cw, sc, fw = self.query(mBowVec)
        return relocalization_candidates(cw, sc, fw, self._seq, covisibility, self._reloc_score)

    def DetectLoopCandidates(self, mBowVec: Dict[int, float], connected, covisibility, minScore: float) -> list:
        """src/KeyFrameDatabase.cc:76-197.  connected: slots of pKF->GetConnectedKeyFrames(); covisibility as above."""
        cw, sc, fw = self.query(mBowVec)
        return loop_candidates(cw, sc, fw, self._seq, set(int(c) for c in connected), covisibility, minScore)

    def SearchByBoW(self, slots, F: KeyFrameView):
        """SearchByBoW(pKF, F, vpMapPointMatches) (src/ORBmatcher.cc:159-288) for database keyframes `slots` against frame F."""
        sl = np.ascontiguousarray(slots, np.int32)
        fc = F._c()
        nF = len(F.mvKeysUn)
        match = np.full((len(sl), max(nF, 1)), -1, np.int32)
        nm = np.zeros(max(len(sl), 1), np.int32)
        m = self._m
        check(self._lib.borb_search_by_bow_db(m._h, self._h, _p(sl), len(sl), C.byref(fc), m.mfNNratio, int(m.mbCheckOrientation),
                                              _p(match), _p(nm)), "borb_search_by_bow_db")
        return nm[:len(sl)], match[:, :nF]


    def SearchByBoWPairs(self, slots, F: KeyFrameView, pairs_cap: Optional[int] = None, want_pairs: bool = True):
        """SearchByBoW of frame F against database keyframes `slots` (None = every slot) with compact results:
        returns (nmatches[n_kf], pair_offset[n_kf], pairs) where pairs[off[k]:off[k]+nm[k]] = (frame feature | keyframe feature << 16)."""
        m = self._m
        fc = F._c()
        if slots is None:
            n_kf, sl = self.size()[0], None
        else:
            sl = np.ascontiguousarray(slots, np.int32); n_kf = len(sl)
        nF = len(F.mvKeysUn)
        cap = int(pairs_cap) if pairs_cap is not None else max(n_kf * nF, 1)
        nm = np.zeros(max(n_kf, 1), np.int32); off = np.zeros(max(n_kf, 1), np.int32)
        pairs = np.zeros(cap if want_pairs else 1, np.uint32)
        tot = C.c_int32(0)
        check(self._lib.borb_search_by_bow_db_pairs(m._h, self._h, _p(sl), n_kf, C.byref(fc), m.mfNNratio, int(m.mbCheckOrientation), _p(nm),
                                                    _p(off), _p(pairs) if want_pairs else None, cap if want_pairs else 0, C.byref(tot)),
              "borb_search_by_bow_db_pairs")
        return nm[:n_kf], off[:n_kf], pairs[:tot.value] if want_pairs else None


class ORBVocabulary:
    """ORBVocabulary = DBoW2::TemplatedVocabulary<FORB> (include/ORBVocabulary.h), device resident."""

    def __init__(self, handle, lib):
        self._h, self._lib = handle, lib

    @staticmethod
    def from_arrays(parent, is_leaf, desc, weight, k: int, L: int, device: int = 0) -> "ORBVocabulary":
        lib = _lib.load()
        parent = np.ascontiguousarray(parent, np.int32); is_leaf = np.ascontiguousarray(is_leaf, np.uint8)
        desc = np.ascontiguousarray(desc, np.uint8); weight = np.ascontiguousarray(weight, np.float64)
        h = C.c_void_p()
        check(lib.borb_voc_create(_p(parent), _p(is_leaf), _p(desc), _p(weight), len(parent), k, L, device, C.byref(h)), "borb_voc_create")
        return ORBVocabulary(h, lib)

    @staticmethod
    def loadFromTextFile(path: str, device: int = 0) -> "ORBVocabulary":
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.borb_voc_load_text(path.encode(), device, C.byref(h)), "borb_voc_load_text")
        return ORBVocabulary(h, lib)

    @staticmethod
    def from_blob(d_ptr: int, nbytes: int, device: int = 0) -> "ORBVocabulary":
        lib = _lib.load()
        h = C.c_void_p()
        check(lib.borb_voc_from_blob(C.c_void_p(d_ptr), nbytes, device, C.byref(h)), "borb_voc_from_blob")
        return ORBVocabulary(h, lib)

    def blob(self) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_size_t()
        check(self._lib.borb_voc_blob(self._h, C.byref(p), C.byref(n)), "borb_voc_blob")
        return p.value, n.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.borb_voc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def transform_raw(self, descriptors: np.ndarray, levelsup: int = 4):
        """Per feature: word id, word weight, node id at level L-levelsup (TemplatedVocabulary.h:1218-1259)."""
        d = np.ascontiguousarray(descriptors, np.uint8)
        n = len(d)
        word = np.zeros(max(n, 1), np.int32); weight = np.zeros(max(n, 1), np.float64); node = np.zeros(max(n, 1), np.int32)
        check(self._lib.borb_bow_transform(self._h, _p(d), n, levelsup, _p(word), _p(weight), _p(node)), "borb_bow_transform")
        return word[:n], weight[:n], node[:n]

    def ComputeBoW(self, descriptors: np.ndarray, levelsup: int = 4):
        """Frame::ComputeBoW (src/Frame.cc:395-402) through borb_compute_bow: (mBowVec as {word: value}, mFeatVec) with the
        ordered-map bookkeeping done in C++ — same results as transform() below."""
        d = np.ascontiguousarray(descriptors, np.uint8)
        n = len(d)
        bw = np.zeros(max(n, 1), np.uint32); bv = np.zeros(max(n, 1), np.float64)
        fn = np.zeros(max(n, 1), np.uint32); fs = np.zeros(n + 1, np.int32); fi = np.zeros(max(n, 1), np.uint32)
        nb, nn = C.c_int32(0), C.c_int32(0)
        check(self._lib.borb_compute_bow(self._h, _p(d), n, levelsup, _p(bw), _p(bv), C.byref(nb), _p(fn), _p(fs), _p(fi), C.byref(nn)), "borb_compute_bow")
        nb, nn = nb.value, nn.value
        return dict(zip(bw[:nb].tolist(), bv[:nb].tolist())), FeatureVector(fn[:nn].copy(), fs[:nn + 1].copy(), fi[:fs[nn]].copy())

    def transform(self, descriptors: np.ndarray, levelsup: int = 4):
        """transform(features, BowVector, FeatureVector, levelsup) (:1127-1194) for TF-IDF / L1 (ORBvoc.txt "10 6 0 0"):
        the tree descent runs on the GPU; the ordered-map bookkeeping is done on the host in feature order, as the reference does."""
        return bow_and_featvec(*self.transform_raw(descriptors, levelsup))
