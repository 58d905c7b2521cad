"""GPU parity of the device-resident keyframe database (KeyFrameDatabase + SearchByBoW against resident keyframes)
with the restatements in oracle/orb_port_match.cpp: shared-word counts, L1 scores (bit-exact as float), candidate
lists of DetectRelocalizationCandidates, and SearchByBoW results."""
import numpy as np
import pytest

from orb_slam2_b200 import synth

pytestmark = pytest.mark.gpu

N_KF = 48


@pytest.fixture(scope="module")
def world(oracle):
    from orb_slam2_b200 import matcher as M
    from orb_slam2_b200.extractor import ORBextractor
    pv = oracle.PortVocabulary.random(10, 4, 5)
    e = pv.export()
    voc = M.ORBVocabulary.from_arrays(e["parent"], e["is_leaf"], e["desc"], e["weight"], e["k"], e["L"])
    X = ORBextractor(800)
    rng = np.random.default_rng(3)
    base = [synth.mono_frame(200 + i // 3, 0, 0, 640, 480) for i in range(N_KF)]      # triples of frames share a scene
    imgs = [np.clip(b.astype(np.int32) + rng.integers(-6, 7, b.shape), 0, 255).astype(np.uint8) for b in base]
    outs = X.extract_batch(imgs)
    kfs, bows = [], []
    for k, d in outs:
        bow, fv = voc.transform(d, 2)
        kfs.append(M.KeyFrameView(mvKeysUn=k, mDescriptors=d, mFeatVec=fv, has_mp=(rng.random(len(k)) < 0.7).astype(np.uint8)))
        bows.append(bow)
    qimg = np.clip(base[7].astype(np.int32) + rng.integers(-8, 9, base[7].shape), 0, 255).astype(np.uint8)
    qk, qd = X(qimg)
    qbow, qfv = voc.transform(qd, 2)
    F = M.KeyFrameView(mvKeysUn=qk, mDescriptors=qd, mFeatVec=qfv)
    return dict(M=M, voc=voc, kfs=kfs, bows=bows, F=F, qbow=qbow, n_words=int(e["is_leaf"].sum()) + len(e["is_leaf"]))


def test_query_counts_and_scores_match_oracle(world, oracle):
    M = world["M"]
    mt = M.ORBmatcher(0.75, True)
    db = M.KeyFrameDatabase(mt)
    slots = [db.add(kf, bow) for kf, bow in zip(world["kfs"], world["bows"])]
    assert slots == list(range(N_KF)) and db.size()[0] == N_KF and db.size()[1] > 0
    cw, sc, fw = db.query(world["qbow"])
    for s in range(N_KF):
        so, co, fo = oracle.port_bow_score(world["qbow"], world["bows"][s])
        assert cw[s] == co and fw[s] == fo, s
        assert sc[s] == np.float32(so), (s, sc[s], so)                    # bit-exact: terms are added in word order
    assert cw.max() > 20 and int(np.argmax(sc)) in (6, 7, 8)                # the query's own scene scores highest
    # erase: the slot stops matching, the others are unchanged
    db.erase(7)
    cw2, sc2, fw2 = db.query(world["qbow"])
    assert cw2[7] == 0 and fw2[7] == 0xFFFFFFFF
    keep = np.arange(N_KF) != 7
    assert np.array_equal(cw2[keep], cw[keep]) and np.array_equal(sc2[keep], sc[keep])
    with pytest.raises(Exception):
        db.erase(7)
    # empty query, empty database
    cw3, sc3, _ = db.query({})
    assert np.all(cw3 == 0) and np.all(sc3 == 0)
    db.clear()
    assert db.size()[0] == 0 and len(db.query(world["qbow"])[0]) == 0


def test_relocalization_candidates_match_oracle(world, oracle):
    M = world["M"]
    mt = M.ORBmatcher(0.75, True)
    db = M.KeyFrameDatabase(mt)
    for kf, bow in zip(world["kfs"], world["bows"]):
        db.add(kf, bow)
    rng = np.random.default_rng(11)
    neigh = np.full((N_KF, 10), -1, np.int32)
    for s in range(N_KF):
        nb = [x for x in (s - 2, s - 1, s + 1, s + 2) if 0 <= x < N_KF] + rng.integers(0, N_KF, 3).tolist()
        nb = [x for x in dict.fromkeys(nb) if x != s][:10]
        neigh[s, :len(nb)] = nb
    covis = lambda s: [int(x) for x in neigh[s] if x >= 0]
    for q in (world["qbow"], world["bows"][20], world["bows"][41]):
        got = db.DetectRelocalizationCandidates(q, covis)
        want = oracle.port_detect_reloc_candidates(world["bows"], world["n_words"], q, neigh)
        assert got == want.tolist() and len(got) >= 1, (got, want)


def test_loop_candidates_match_oracle(world, oracle):
    """KeyFrameDatabase::DetectLoopCandidates (src/KeyFrameDatabase.cc:76-197) from the GPU query + the host logic."""
    M = world["M"]
    mt = M.ORBmatcher(0.75, True)
    db = M.KeyFrameDatabase(mt)
    for kf, bow in zip(world["kfs"], world["bows"]):
        db.add(kf, bow)
    rng = np.random.default_rng(13)
    neigh = np.full((N_KF, 10), -1, np.int32)
    for s in range(N_KF):
        nb = [x for x in (s - 1, s + 1, s + 2) if 0 <= x < N_KF] + rng.integers(0, N_KF, 2).tolist()
        nb = [x for x in dict.fromkeys(nb) if x != s][:10]
        neigh[s, :len(nb)] = nb
    covis = lambda s: [int(x) for x in neigh[s] if x >= 0]
    for q, conn, min_score in ((world["qbow"], [6, 30], 0.0), (world["bows"][20], [19, 20, 21], 0.02), (world["bows"][41], [], 0.15)):
        connected = np.zeros(N_KF, np.uint8); connected[conn] = 1
        got = db.DetectLoopCandidates(q, conn, covis, min_score)
        want = oracle.port_detect_loop_candidates(world["bows"], world["n_words"], q, connected, neigh, min_score).tolist()
        assert got == want, (got, want)


def test_search_by_bow_against_resident_keyframes(world, oracle):
    M = world["M"]
    mt = M.ORBmatcher(0.75, True)
    db = M.KeyFrameDatabase(mt)
    for kf, bow in zip(world["kfs"], world["bows"]):
        db.add(kf, bow)
    F = world["F"]
    slots = [6, 7, 8, 20, 7, 47, 0]
    nm, match = db.SearchByBoW(slots, F)
    nm_h, match_h = mt.SearchByBoW([world["kfs"][s] for s in slots], F)          # same kernel, keyframes staged from the host
    assert np.array_equal(nm, nm_h) and np.array_equal(match, match_h)
    for i, s in enumerate(slots):
        n_o, m_o = oracle.port_search_by_bow(world["kfs"][s], F, 0.75, True)
        assert nm[i] == n_o and np.array_equal(match[i], m_o), s
    assert nm[1] > 30
    # MapPoint mask updated in place
    hm = np.zeros(len(world["kfs"][7].mvKeysUn), np.uint8); hm[::2] = 1
    db.set_has_mp(7, hm)
    kf7 = M.KeyFrameView(world["kfs"][7].mvKeysUn, world["kfs"][7].mDescriptors, world["kfs"][7].mFeatVec, has_mp=hm)
    nm2, match2 = db.SearchByBoW([7], F)
    n_o, m_o = oracle.port_search_by_bow(kf7, F, 0.75, True)
    assert nm2[0] == n_o and np.array_equal(match2[0], m_o)
