"""CPU: sanity of the matcher / vocabulary restatements (oracle/orb_port_match.cpp) against brute-force numpy
re-derivations and invariants.  (The reference has no tests for these functions; see the file header there.)"""
import os

import numpy as np
import pytest

from tests import match_fixtures as mf


@pytest.fixture(scope="module")
def views(oracle):
    return mf.two_views(oracle, 7)


def test_features_in_area_equals_bruteforce_in_grid_order(oracle, views):
    k = views["kl"]
    bounds = (0.0, 0.0, 640.0, 480.0)
    rng = np.random.default_rng(0)
    invw, invh = np.float32(64) / np.float32(640), np.float32(48) / np.float32(480)
    for _ in range(50):
        x, y, r = rng.uniform(-20, 660), rng.uniform(-20, 500), rng.uniform(3, 60)
        lo = int(rng.integers(-1, 7)); hi = lo + int(rng.integers(0, 2))
        got = oracle.port_features_in_area(k, bounds, x, y, r, lo, hi)
        m = (np.abs(k["x"] - np.float32(x)) < np.float32(r)) & (np.abs(k["y"] - np.float32(y)) < np.float32(r))
        if lo > 0 or hi >= 0:
            m &= k["octave"] >= lo
            if hi >= 0:
                m &= k["octave"] <= hi
        want = np.nonzero(m)[0]
        assert sorted(got.tolist()) == want.tolist()
        # order: (cell x, cell y, insertion)
        cx = np.floor(k["x"][got] * invw + np.float32(0.5)).astype(int); cy = np.floor(k["y"][got] * invh + np.float32(0.5)).astype(int)   # C round(): half away from zero
        key = list(zip(cx.tolist(), cy.tolist(), got.tolist()))
        assert key == sorted(key)


def test_search_by_projection_invariants(oracle, views):
    F, mps = mf.projection_case(views, 3)
    n, match = oracle.port_search_by_projection(F, mps, 3.0, 0.8)
    assert n == int((match >= 0).sum()) and n > 50
    m = match[match >= 0]
    assert np.all(F.occupied[m] == 0)                                   # never lands on an occupied feature
    # a feature is claimed twice only when the first claimant had no observations
    for f in set(m.tolist()):
        owners = np.nonzero(match == f)[0]
        assert np.all(mps.has_obs[owners[:-1]] == 0)
    for i in np.nonzero(match >= 0)[0][:100]:
        d = oracle.port_lib().orbport_hamming(mps.descriptors[i].ctypes.data_as(oracle._u8p), F.mDescriptors[match[i]].ctypes.data_as(oracle._u8p))
        assert d <= 100
    assert np.all(match[mps.valid == 0] == -1)


def test_bow_searches_invariants(oracle, views):
    voc = oracle.PortVocabulary.random(10, 4, 5)
    kf1, kf2 = mf.keyframe_views(views, voc, 9)
    n, match = oracle.port_search_by_bow(kf1, kf2, 0.7, True)
    assert n == int((match >= 0).sum()) and n > 20
    assert np.all(kf1.has_mp[match[match >= 0]] == 1)
    assert len(set(match[match >= 0].tolist())) <= n
    n2, m12 = oracle.port_search_by_bow_kf(kf1, kf2, 0.75, True)
    assert n2 == int((m12 >= 0).sum())
    j = m12[m12 >= 0]
    assert len(set(j.tolist())) == len(j) and np.all(kf2.has_mp[j] == 1) and np.all(kf1.has_mp[np.nonzero(m12 >= 0)[0]] == 1)
    n_no, _ = oracle.port_search_by_bow(kf1, kf2, 0.7, False)
    assert n_no >= n                                                     # the rotation cull only removes matches
    pairs = oracle.port_search_for_triangulation(kf1, kf2, mf.rectified_F12(1), (-1000.0, 200.0), False, True)
    assert len(pairs) > 10 and np.all(np.diff(pairs[:, 0]) > 0)
    assert np.all(kf1.has_mp[pairs[:, 0]] == 0) and np.all(kf2.has_mp[pairs[:, 1]] == 0)
    so = oracle.port_search_for_triangulation(kf1, kf2, mf.rectified_F12(1), (-1000.0, 200.0), True, True)
    assert np.all(kf1.mvuRight[so[:, 0]] >= 0) and np.all(kf2.mvuRight[so[:, 1]] >= 0)


def test_vocabulary_text_round_trip_and_descent(oracle, tmp_path, views):
    voc = oracle.PortVocabulary.random(10, 3, 11)
    path = os.path.join(tmp_path, "voc.txt")
    voc.save_text(path)
    voc2 = oracle.PortVocabulary.load_text(path)
    e1, e2 = voc.export(), voc2.export()
    for key in ("parent", "is_leaf", "word_id", "desc", "weight"):
        assert np.array_equal(e1[key], e2[key]), key
    assert e1["k"] == 10 and e1["L"] == 3 and len(e1["parent"]) == 1 + 10 + 100 + 1000
    d = views["dl"]
    w1, wt1, n1 = voc.transform_raw(d, 2)
    w2, wt2, n2 = voc2.transform_raw(d, 2)
    assert np.array_equal(w1, w2) and np.array_equal(wt1, wt2) and np.array_equal(n1, n2)
    # brute-force descent in numpy
    desc = e1["desc"]; parent = e1["parent"]
    children = {}
    for i in range(1, len(parent)):
        children.setdefault(int(parent[i]), []).append(i)
    bits = np.unpackbits(desc, axis=1)
    for f in range(0, len(d), 37):
        fb = np.unpackbits(d[f])
        node, lvl, at1 = 0, 0, 0
        while node in children:
            ch = children[node]
            dist = [(int((bits[c] != fb).sum())) for c in ch]
            node = ch[int(np.argmin(dist))]          # argmin: first minimum
            lvl += 1
            if lvl == 1:
                at1 = node
        assert e1["word_id"][node] == w1[f] and at1 == n1[f]


def _hamming(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def test_projection_sim3_equals_numpy_rederivation(oracle, views):
    """SearchByProjection(pKF, Scw, ...) (ORBmatcher.cc:290-403) re-derived in float32 numpy, without the grid: the area
    query is a brute-force window, candidates visited in the grid's (cell x, cell y, insertion) order."""
    import ctypes as C
    libm = C.CDLL("libm.so.6"); libm.logf.restype = C.c_float; libm.logf.argtypes = [C.c_float]
    f32 = np.float32
    KF, P, Tcw, Ow, K = mf.world_points_case(views, 41)
    fx, fy, cx, cy = map(f32, K)
    th = 10
    n_o, s_o = oracle.port_search_by_projection_sim3(KF, P, Tcw, Ow, K, th)
    k = KF.mvKeysUn
    minX, minY, maxX, maxY = map(f32, KF.bounds)
    invw, invh = f32(64) / (maxX - minX), f32(48) / (maxY - minY)
    cellx = np.floor((k["x"] - minX) * invw + f32(0.5)).astype(int); celly = np.floor((k["y"] - minY) * invh + f32(0.5)).astype(int)
    order = sorted(range(len(k)), key=lambda i: (cellx[i], celly[i], i))
    logs = f32(libm.logf(float(f32(KF.mvScaleFactors[1]))))
    held = KF.occupied.astype(bool).copy()
    state = np.full(len(k), -1, np.int32)
    nm = 0
    T = Tcw.astype(f32)
    for i in range(len(P.world_pos)):
        if not P.valid[i]:
            continue
        p = P.world_pos[i]
        pc = [f32(f32(f32(T[r, 0] * p[0]) + f32(T[r, 1] * p[1])) + f32(T[r, 2] * p[2])) + T[r, 3] for r in range(3)]
        if pc[2] < 0:
            continue
        with np.errstate(all="ignore"):
            invz = f32(1) / pc[2]
            u = f32(fx * f32(pc[0] * invz)) + cx
            v = f32(fy * f32(pc[1] * invz)) + cy
        if not (u >= minX and u < maxX and v >= minY and v < maxY):
            continue
        PO = (p - Ow.astype(f32)).astype(f32)
        dist = f32(np.sqrt(np.sum(PO.astype(np.float64) ** 2)))
        if dist < f32(0.8) * P.min_distance[i] or dist > f32(1.2) * P.max_distance[i]:
            continue
        if float(np.dot(PO.astype(np.float64), P.normal[i].astype(np.float64))) < 0.5 * float(dist):
            continue
        ratio = P.max_distance[i] / dist
        lvl = int(np.ceil(f32(libm.logf(float(ratio))) / logs))
        lvl = min(max(lvl, 0), len(KF.mvScaleFactors) - 1)
        r = f32(th) * f32(KF.mvScaleFactors[lvl])
        best, bidx = 256, -1
        for j in order:
            if not (abs(k["x"][j] - u) < r and abs(k["y"][j] - v) < r) or held[j]:
                continue
            if k["octave"][j] < lvl - 1 or k["octave"][j] > lvl:
                continue
            d = _hamming(P.descriptors[i], KF.mDescriptors[j])
            if d < best:
                best, bidx = d, j
        if best <= 50:
            state[bidx] = i; held[bidx] = True; nm += 1
    assert nm == n_o and nm > 50
    assert np.array_equal(state, s_o)


def test_projection_kf_invariants(oracle, views):
    Cur, P, Tcw, Ow, K = mf.world_points_case(views, 42)
    n, s = oracle.port_search_by_projection_kf(Cur, P, Tcw, Ow, K, 10.0, 100, True)
    n2, s2 = oracle.port_search_by_projection_kf(Cur, P, Tcw, Ow, K, 10.0, 100, False)
    assert n == int((s >= 0).sum()) and n > 50
    assert n2 >= n and np.all((s2 >= 0) | (s < 0))                       # the orientation cull only removes matches
    assert np.array_equal(s2[s >= 0], s[s >= 0])
    hit = np.nonzero(s2 >= 0)[0]
    assert np.all(Cur.occupied[hit] == 0) and np.all(P.valid[s2[hit]] == 1)
    assert len(set(s2[hit].tolist())) == len(hit)                        # a query claims at most one feature
    for j in hit[:100]:
        assert _hamming(P.descriptors[s2[j]], Cur.mDescriptors[j]) <= 100


def test_fuse_sim3_initialization_invariants(oracle, views):
    """Invariants of the restatements of Fuse (search part), SearchBySim3 and SearchForInitialization."""
    KF, P, Tcw, Ow, K, bf = mf.fuse_case(views, 43)
    n, best = oracle.port_fuse(KF, P, Tcw, Ow, K, bf, 3.0, False)
    n_s, best_s = oracle.port_fuse(KF, P, Tcw, Ow, K, bf, 3.0, True)
    assert n == int((best >= 0).sum()) and n > 50 and n_s >= n              # the Scw overload has no reprojection gates
    assert np.all(best[P.valid == 0] == -1)
    for i in np.nonzero(best >= 0)[0][:100]:
        assert _hamming(P.descriptors[i], KF.mDescriptors[best[i]]) <= 50
    a = mf.sim3_case(views, 44)
    n, m12 = oracle.port_search_by_sim3(*a, 7.5)
    hit = np.nonzero(m12 >= 0)[0]
    assert n == len(hit) > 50 and len(set(m12[hit].tolist())) == len(hit)
    # swapping the roles of the two keyframes must return the inverse matching (the agreement test is symmetric)
    KF1, KF2, P1, P2, T1w, T2w, S12, S21, Kc = a
    n2, m21 = oracle.port_search_by_sim3(KF2, KF1, P2, P1, T2w, T1w, S21, S12, Kc, 7.5)
    assert n2 == n and all(m21[m12[i]] == i for i in hit)
    b = (0.0, 0.0, float(views["w"]), float(views["h"]))
    from orb_slam2_b200.matcher import FrameView
    F1, F2 = FrameView(views["kl"], views["dl"], views["scale"], b), FrameView(views["kr"], views["dr"], views["scale"], b)
    prev = np.stack([views["kl"]["x"], views["kl"]["y"]], 1).astype(np.float32)
    n, m, p = oracle.port_search_for_initialization(F1, F2, prev, 100, 0.9, True)
    hit = np.nonzero(m >= 0)[0]
    assert n == len(hit) > 20 and len(set(m[hit].tolist())) == len(hit)
    assert np.all(views["kl"]["octave"][hit] == 0)
    assert np.array_equal(p[hit], np.stack([views["kr"]["x"][m[hit]], views["kr"]["y"][m[hit]]], 1))
    untouched = np.setdiff1d(np.arange(len(prev)), hit)
    assert np.array_equal(p[untouched], prev[untouched])


def test_bow_score_and_reloc_candidates_restatements(oracle):
    """L1 score against a dense numpy evaluation; DetectRelocalizationCandidates against hand-checkable cases."""
    rng = np.random.default_rng(6)
    def rand_bow(n_words, n):
        w = np.sort(rng.choice(n_words, n, replace=False))
        v = rng.random(n); v /= v.sum()
        return dict(zip(w.tolist(), v.tolist()))
    for _ in range(20):
        a, b = rand_bow(500, 120), rand_bow(500, 150)
        s, c, f = oracle.port_bow_score(a, b)
        da = np.zeros(500); db = np.zeros(500)
        da[list(a)] = list(a.values()); db[list(b)] = list(b.values())
        shared = sorted(set(a) & set(b))
        assert c == len(shared) and (f == shared[0] if shared else f == 0xFFFFFFFF)
        assert abs(s - (1.0 - 0.5 * np.abs(da - db).sum())) < 1e-12          # ||v-w||_1 identity (Nister 2006), both L1-normalised
    assert oracle.port_bow_score(a, a)[0] == pytest.approx(1.0, abs=1e-15)
    # three keyframes: 0 and 1 share the query's words, 2 shares none; 1 is 0's covisible neighbour and scores higher
    q = {1: 0.5, 2: 0.5}
    bows = [{1: 0.6, 2: 0.2, 9: 0.2}, {1: 0.5, 2: 0.5}, {7: 1.0}]
    neigh = np.full((3, 10), -1, np.int32); neigh[0, 0] = 1; neigh[1, 0] = 0; neigh[2, 0] = 0
    assert oracle.port_detect_reloc_candidates(bows, 16, q, neigh).tolist() == [1]      # both groups elect keyframe 1, listed once
    assert oracle.port_detect_reloc_candidates(bows, 16, {7: 1.0}, neigh).tolist() == [2]
    assert oracle.port_detect_reloc_candidates(bows, 16, {5: 1.0}, neigh).tolist() == []


def test_keyframe_database_host_logic_equals_restated_reference(oracle):
    """orb_slam2_b200.matcher.relocalization_candidates / loop_candidates (the host half of the device-resident keyframe
    database: ordering, thresholds, covisibility accumulation) against the restatements of KeyFrameDatabase.cc that walk a real
    inverted file.  The per-keyframe inputs the GPU query provides (shared words, float L1 score, first shared word) come
    from the oracle here, so the whole chain is checked without a GPU."""
    from orb_slam2_b200 import matcher as M
    rng = np.random.default_rng(12)
    n_words, n_kf = 800, 60
    centers = [rng.choice(n_words, 90, replace=False) for _ in range(6)]           # six "places"

    def bow_near(c):
        keep = centers[c][rng.random(90) < 0.8]
        extra = rng.choice(n_words, 25, replace=False)
        w = np.unique(np.concatenate([keep, extra]))
        v = rng.random(len(w)); v /= v.sum()
        return dict(zip(w.tolist(), v.tolist()))
    place = rng.integers(0, 6, n_kf)
    bows = [bow_near(int(p)) for p in place]
    neigh = np.full((n_kf, 10), -1, np.int32)
    for s in range(n_kf):
        same = [int(x) for x in np.nonzero(place == place[s])[0] if x != s]
        nb = list(dict.fromkeys(same[:4] + rng.integers(0, n_kf, 3).tolist()))
        nb = [x for x in nb if x != s][:10]
        neigh[s, :len(nb)] = nb
    covis = lambda s: [int(x) for x in neigh[s] if x >= 0]
    seq = list(range(n_kf))
    checked = 0
    for q in [bow_near(0), bow_near(3), bows[17], {5: 1.0}, {}]:
        per = [oracle.port_bow_score(q, b) for b in bows]
        sc = np.array([np.float32(p[0]) for p in per], np.float32); cw = np.array([p[1] for p in per], np.int32); fw = np.array([p[2] for p in per], np.uint32)
        got = M.relocalization_candidates(cw, sc, fw, seq, covis)
        want = oracle.port_detect_reloc_candidates(bows, n_words, q, neigh).tolist()
        assert got == want, (got, want)
        for min_score, conn in [(0.0, []), (0.05, [3, 17, 20]), (0.3, list(range(0, n_kf, 2)))]:
            connected = np.zeros(n_kf, np.uint8); connected[conn] = 1
            got = M.loop_candidates(cw, sc, fw, seq, set(conn), covis, min_score)
            want = oracle.port_detect_loop_candidates(bows, n_words, q, connected, neigh, min_score).tolist()
            assert got == want, (min_score, conn, got, want)
            assert not set(got) & set(conn)
            checked += len(want)
    assert checked > 10
