"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes bindings for oracle/liborbport.so (CPU restatement) and oracle/_ref/liborbref.so (the
reference's own ORBextractor.cc compiled verbatim against oracle/cvshim).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package (orb_slam2_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liborbport.so")
REF_SO = os.path.join(HERE, "_ref", "liborbref.so")
REFERENCE_ROOT = "/root/reference"

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28

_u8p = C.POINTER(C.c_uint8)
_f32p = C.POINTER(C.c_float)
_i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> None:
    """Compile the oracle (port always; _ref only where /root/reference exists)."""
    targets = ["port"]
    if os.path.exists(os.path.join(REFERENCE_ROOT, "src", "ORBextractor.cc")):
        targets.append("ref")
    if force:
        subprocess.check_call(["make", "-C", HERE, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", HERE] + targets, stdout=subprocess.DEVNULL)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def _p(a, t):
    return a.ctypes.data_as(t)


class _ExtractorBase:
    """Shared python face of the two CPU extractors (same call signature as the product's)."""

    def __init__(self, lib, prefix, nfeatures=2000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self._lib, self._px = lib, prefix
        self.nfeatures, self.nlevels = nfeatures, nlevels
        create = getattr(lib, prefix + "_create")
        create.restype = C.c_void_p
        create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        self._h = C.c_void_p(create(nfeatures, scale_factor, nlevels, ini_th, min_th))
        ext = getattr(lib, prefix + "_extract")
        ext.restype = C.c_int
        ext.argtypes = [C.c_void_p, _u8p, C.c_int, C.c_int, C.c_int, C.c_void_p, _u8p, C.c_int]
        self._extract = ext
        tab = getattr(lib, prefix + "_tables")
        tab.restype = None
        tab.argtypes = [C.c_void_p, _f32p, _f32p, _f32p, _f32p, _i32p, _i32p]
        L = nlevels
        self.scale = np.zeros(L, np.float32); self.inv_scale = np.zeros(L, np.float32)
        self.sigma2 = np.zeros(L, np.float32); self.inv_sigma2 = np.zeros(L, np.float32)
        self.per_level = np.zeros(L, np.int32); self.umax = np.zeros(16, np.int32)
        tab(self._h, _p(self.scale, _f32p), _p(self.inv_scale, _f32p), _p(self.sigma2, _f32p),
            _p(self.inv_sigma2, _f32p), _p(self.per_level, _i32p), _p(self.umax, _i32p))

    def __del__(self):
        try:
            d = getattr(self._lib, self._px + "_destroy")
            d.restype = None
            d.argtypes = [C.c_void_p]
            d(self._h)
        except Exception:
            pass

    def __call__(self, img: np.ndarray):
        img = np.ascontiguousarray(img, dtype=np.uint8)
        h, w = img.shape
        cap = self.nfeatures + 64
        while True:
            kps = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, 32), np.uint8)
            n = self._extract(self._h, _p(img, _u8p), w, h, img.strides[0], kps.ctypes.data_as(C.c_void_p), _p(desc, _u8p), cap)
            if n <= cap:
                return kps[:n].copy(), desc[:n].copy()
            cap = n


class PortExtractor(_ExtractorBase):
    def __init__(self, *a, **k):
        lib = C.CDLL(PORT_SO)
        super().__init__(lib, "orbport", *a, **k)
        lib.orbport_level_size.argtypes = [C.c_void_p, C.c_int, _i32p, _i32p]
        lib.orbport_level_ptr.restype = C.c_void_p
        lib.orbport_level_ptr.argtypes = [C.c_void_p, C.c_int]
        lib.orbport_blur_ptr.restype = C.c_void_p
        lib.orbport_blur_ptr.argtypes = [C.c_void_p, C.c_int]
        lib.orbport_candidates.argtypes = [C.c_void_p, C.c_int, _i32p, C.c_int]
        lib.orbport_level_count.argtypes = [C.c_void_p, C.c_int]

    def _level(self, fn, level):
        w, h = C.c_int32(), C.c_int32()
        self._lib.orbport_level_size(self._h, level, C.byref(w), C.byref(h))
        p = fn(self._h, level)
        if not p:
            return None
        return np.ctypeslib.as_array(C.cast(p, _u8p), shape=(h.value, w.value)).copy()

    def level(self, level):
        return self._level(self._lib.orbport_level_ptr, level)

    def blurred(self, level):
        return self._level(self._lib.orbport_blur_ptr, level)

    def candidates(self, level):
        n = self._lib.orbport_candidates(self._h, level, None, 0)
        out = np.zeros((max(n, 1), 3), np.int32)
        self._lib.orbport_candidates(self._h, level, _p(out, _i32p), n)
        return out[:n]

    def level_count(self, level):
        return self._lib.orbport_level_count(self._h, level)


class RefExtractor(_ExtractorBase):
    """The reference's ORBextractor::operator() itself (verbatim source + shim + monotonic allocator)."""

    def __init__(self, *a, **k):
        lib = C.CDLL(REF_SO)
        super().__init__(lib, "orbref", *a, **k)
        lib.orbref_pyramid.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), _i32p, _i32p, _i32p]
        lib.orbref_distribute.argtypes = [C.c_void_p, _f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, _f32p, C.c_int]

    def level(self, level):
        p = C.c_void_p()
        w, h, s = C.c_int32(), C.c_int32(), C.c_int32()
        if self._lib.orbref_pyramid(self._h, level, C.byref(p), C.byref(w), C.byref(h), C.byref(s)) != 0:
            return None
        a = np.ctypeslib.as_array(C.cast(p, _u8p), shape=(h.value, s.value))
        return a[:, :w.value].copy()

    def distribute(self, xyr: np.ndarray, width: int, height: int, N: int):
        """DistributeOctTree on (x,y,response) rows relative to minBorder; returns selected rows in list order."""
        xyr = np.ascontiguousarray(xyr, np.float32)
        cap = len(xyr) + 8
        out = np.zeros((cap, 3), np.float32)
        n = self._lib.orbref_distribute(self._h, _p(xyr, _f32p), len(xyr), 16, 16 + width, 16, 16 + height, N, 0,
                                        _p(out, _f32p), cap)
        return out[:n]


def port_lib():
    lib = C.CDLL(PORT_SO)
    lib.orbport_distribute.argtypes = [_i32p, C.c_int, C.c_int, C.c_int, C.c_int, _i32p, C.c_int]
    lib.orbport_hamming.argtypes = [_u8p, _u8p]
    return lib


def port_distribute(xys: np.ndarray, width: int, height: int, N: int) -> np.ndarray:
    lib = port_lib()
    xys = np.ascontiguousarray(xys, np.int32)
    out = np.zeros((len(xys) + 8, 3), np.int32)
    n = lib.orbport_distribute(_p(xys, _i32p), len(xys), width, height, N, _p(out, _i32p), len(out))
    return out[:n]


def port_stereo(kL, dL, kR, dR, pyrL, pyrR, scale, inv_scale, bf, fx):
    """Frame::ComputeStereoMatches restatement. pyrL/pyrR: lists of tight uint8 level images.
    Returns (uRight, depth, sad) with -1 for 'no match'."""
    lib = C.CDLL(PORT_SO)
    nlev = len(pyrL)
    pyrL = [np.ascontiguousarray(p, np.uint8) for p in pyrL]
    pyrR = [np.ascontiguousarray(p, np.uint8) for p in pyrR]
    lw = np.array([p.shape[1] for p in pyrL], np.int32)
    lh = np.array([p.shape[0] for p in pyrL], np.int32)
    PL = (C.c_void_p * nlev)(*[p.ctypes.data for p in pyrL])
    PR = (C.c_void_p * nlev)(*[p.ctypes.data for p in pyrR])
    kL = np.ascontiguousarray(kL); kR = np.ascontiguousarray(kR)
    dL = np.ascontiguousarray(dL, np.uint8); dR = np.ascontiguousarray(dR, np.uint8)
    n = len(kL)
    ur = np.zeros(max(n, 1), np.float32); dp = np.zeros(max(n, 1), np.float32); sad = np.zeros(max(n, 1), np.int32)
    scale = np.ascontiguousarray(scale, np.float32); inv_scale = np.ascontiguousarray(inv_scale, np.float32)
    b = np.float32(bf) / np.float32(fx)
    lib.orbport_stereo.restype = C.c_int
    lib.orbport_stereo.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_float,
                                   C.c_void_p, C.c_void_p, C.c_void_p]
    lib.orbport_stereo(kL.ctypes.data, dL.ctypes.data, n, kR.ctypes.data, dR.ctypes.data, len(kR), PL, PR,
                       lw.ctypes.data, lh.ctypes.data, nlev, scale.ctypes.data, inv_scale.ctypes.data,
                       float(bf), float(b), ur.ctypes.data, dp.ctypes.data, sad.ctypes.data)
    return ur[:n], dp[:n], sad[:n]
