#!/usr/bin/env python
"""bench.py — stereo frames/s for ORB extract(L)+extract(R)+ComputeStereoMatches on synthetic
KITTI-shaped 1242x375 pairs at 2000 keypoints (BASELINE.json configs[1]); see DESIGN.md §Measurement.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs B] [--impl b200|reference]
  N>1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of B synthetic stereo pairs per GPU.
  value     frames/s, whole job, inputs already resident in HBM (borb_stereo_frames_device_enqueue, two handles =
            two batches in flight on two CUDA streams)
  e2e       same metric through the C-ABI call with HOST (pinned) buffers: H2D of the images and D2H of
            keypoints/descriptors/uRight/depth inside the timed region (borb_stereo_frames_enqueue,
            two handles double-buffered)
  roofline  FAST/NMS kernel: algorithmic bytes (sum of level pixels x images per launch) / its mean
            launch time from CUDA events on the library's stream, over the timed region
  cpu_baseline  the reference's own ORBextractor.cc (oracle/_ref, compiled verbatim) + the stereo
            restatement, timed on this box's host cores on a bounded sample (rank 0, N=1 only)
--impl reference: that CPU implementation alone, all host threads, same metric/config (no GPU work).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

W_IMG, H_IMG, NFEAT = 1242, 375, 2000            # KITTI00-02.yaml shape / ORBextractor.nFeatures
BF, FX = 386.1448, 718.856                        # Camera.bf, Camera.fx (KITTI00-02.yaml)
LEVEL_PIXELS = 1441432                            # sum_l w_l*h_l for 1242x375, 8 levels, x1.2
SEED = 2024
METRIC = "stereo frames/sec (ORB extract L+R + ComputeStereoMatches, KITTI-shaped 1242x375 @2000 kpts)"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_pairs(stream_id: int, n: int):
    from orb_slam2_b200 import synth
    L, R = [], []
    for i in range(n):
        l, r, _ = synth.stereo_pair(SEED, stream_id, i, W_IMG, H_IMG)
        L.append(l); R.append(r)
    return L, R


# ----------------------------------------------------------------------------------------------
# CPU reference arm (oracle/_ref = the reference's ORBextractor.cc verbatim; stereo = restatement)
# ----------------------------------------------------------------------------------------------
def cpu_worker_factory():
    from oracle import oracle_lib as O
    O.build()
    kind = "reference" if O.have_ref() else "port"
    Ext = O.RefExtractor if kind == "reference" else O.PortExtractor

    def make():
        EL, ER = Ext(NFEAT), Ext(NFEAT)

        def run(l, r):
            # reference-faithful work per frame: extract L, extract R, ComputeStereoMatches
            kl, dl = EL(l)
            kr, dr = ER(r)
            ur, dp, _ = O.port_stereo(kl, dl, kr, dr, [EL.level(i) for i in range(8)], [ER.level(i) for i in range(8)],
                                      EL.scale, EL.inv_scale, BF, FX)
            return len(kl), int((ur >= 0).sum())
        return run
    return kind, make


class CpuPool:
    """One independent camera stream per host thread (ctypes releases the GIL inside the oracle calls)."""

    def __init__(self, L, R, threads: int):
        self.kind, make = cpu_worker_factory()
        self.L, self.R, self.threads = L, R, threads
        self.runs = [make() for _ in range(threads)]
        self.run(1)                      # warm-up: one frame per worker (arena page faults, caches)

    def run(self, pairs_per_thread: int):
        """Processes threads*pairs_per_thread frames; returns (frames/s, seconds)."""
        def body(t):
            for i in range(pairs_per_thread):
                j = (t * pairs_per_thread + i) % len(self.L)
                self.runs[t](self.L[j], self.R[j])
        ths = [threading.Thread(target=body, args=(t,)) for t in range(self.threads)]
        t0 = time.perf_counter()
        for th in ths: th.start()
        for th in ths: th.join()
        dt = time.perf_counter() - t0
        return self.threads * pairs_per_thread / dt, dt


def run_reference(args, rank: int):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    L, R = make_pairs(0, 16)
    per = max(1, args.ref_pairs_per_thread)
    pool = CpuPool(L, R, threads)
    for _ in range(args.warmup):
        pool.run(1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.run(per)
    total = time.perf_counter() - t0
    val = threads * per * args.steps / total
    kind = pool.kind
    sample = f"{threads} threads x {per} synthetic KITTI-shaped stereo pairs per step, {args.steps} steps"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "configs[1]: stereo KITTI-00-shaped 1242x375, 2000 feats, extract + ComputeStereoMatches"},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampler
# ----------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw")

    def __init__(self, gpu: int):
        self.gpu, self.p, self.lines = gpu, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------
def run_b200(args, rank: int, world: int, local_rank: int):
    import torch
    import torch.distributed as dist
    from orb_slam2_b200 import _lib
    from orb_slam2_b200.extractor import ORBextractor

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device visible - the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    B, K, Wm = args.pairs, args.steps, args.warmup
    NBUF = 4

    # ---- NCCL plumbing that the path really has: the packed vocabulary (k=10, L=6 tree of ORBvoc's shape,
    # ~48 MB) is built on rank 0 only, broadcast ONCE over NCCL into every GPU's HBM and adopted there; counters are
    # all-gathered at the end.  No collective touches the per-frame data path.
    from orb_slam2_b200 import sharding
    from orb_slam2_b200.matcher import ORBVocabulary
    voc_ms, voc_bytes, voc = None, None, None
    if world > 1:
        if rank == 0:
            voc = ORBVocabulary.from_arrays(*sharding.random_vocabulary_arrays(10, 6, 7), 10, 6, device=local_rank)
            ptr, nbytes = voc.blob()
            src_blob = torch.as_tensor(sharding.DeviceBlobView(ptr, nbytes), device=dev)
        else:
            src_blob = None
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        blob = sharding.broadcast_blob(src_blob, src=0, device=dev)
        torch.cuda.synchronize()
        voc_ms = (time.perf_counter() - t0) * 1e3
        voc_bytes = int(blob.numel())
        if rank != 0:
            voc = ORBVocabulary.from_blob(blob.data_ptr(), voc_bytes, device=local_rank)

    # ---- synthetic inputs: B distinct pairs of this rank's camera stream; NBUF rotating batches (row-rolled
    # copies keep the stereo geometry) so consecutive steps never re-read the same pixels from L2
    t0 = time.perf_counter()
    Ls, Rs = make_pairs(rank, B)
    host = np.empty((NBUF, 2 * B, H_IMG, W_IMG), np.uint8)
    for j in range(NBUF):
        for p in range(B):
            host[j, 2 * p] = np.roll(Ls[p], 37 * j, axis=0)
            host[j, 2 * p + 1] = np.roll(Rs[p], 37 * j, axis=0)
    log(f"[rank {rank}] generated {B} pairs x {NBUF} buffers in {time.perf_counter() - t0:.1f}s")
    d_in = torch.from_numpy(host).to(dev)
    pitch, img_stride = W_IMG, W_IMG * H_IMG

    # Two handles (= two camera-stream batches in flight, each with its own CUDA stream) keep the GPU busy across the
    # dependent kernels of one batch; both timed regions use the same two handles.
    NH = args.handles
    exts = [ORBextractor(NFEAT, 1.2, 8, 20, 7, device=local_rank) for _ in range(NH)]
    ext = exts[0]
    cap = ext.capacity(W_IMG, H_IMG)
    for x in exts:
        x.reserve(W_IMG, H_IMG, 2 * B)
    b = float(np.float32(BF) / np.float32(FX))
    streams = []
    for x in exts:
        sp = C.c_void_p()
        _lib.check(lib.borb_extractor_stream(x._h, C.byref(sp)), "borb_extractor_stream")
        streams.append(torch.cuda.ExternalStream(sp.value, device=dev))
    n_lr = [(torch.zeros(B, dtype=torch.int32).pin_memory(), torch.zeros(B, dtype=torch.int32).pin_memory()) for _ in range(NH)]
    n_left, n_right = n_lr[0]

    def step_resident(k):
        x, (nl, nr) = exts[k % NH], n_lr[k % NH]
        buf = d_in[k % NBUF]
        _lib.check(lib.borb_stereo_frames_device_enqueue(x._h, buf.data_ptr(), B, W_IMG, H_IMG, pitch, img_stride, BF, b,
                                                         nl.data_ptr(), nr.data_ptr(), None, None, cap), "stereo_frames_device_enqueue")

    def drain():
        for x in exts:
            _lib.check(lib.borb_sync(x._h), "borb_sync")

    clocks = Clocks(local_rank if os.environ.get("CUDA_VISIBLE_DEVICES") is None else int(os.environ["CUDA_VISIBLE_DEVICES"].split(",")[local_rank]))
    clocks.start()          # sampled across warm-up and BOTH timed regions (continuous load)
    for k in range(max(Wm, 3) * NH):
        step_resident(k)
    drain()
    assert int(n_left.min()) >= NFEAT, "warm-up produced too few keypoints"

    # ---- timed region 1: HBM-resident throughput, CUDA events on the library's streams
    launches0 = sum(x.launch_count() for x in exts)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1s = [torch.cuda.Event(enable_timing=True) for _ in exts]
    e0.record(streams[0])
    for st_ in streams[1:]:
        st_.wait_event(e0)                 # every stream starts after the common start mark
    for k in range(K):
        step_resident(k)
    for ev, st_ in zip(e1s, streams):
        ev.record(st_)
    drain()
    torch.cuda.synchronize()
    ms = max(e0.elapsed_time(ev) for ev in e1s)
    launches = sum(x.launch_count() for x in exts) - launches0
    # per-kernel device times: a short single-handle pass right after the timed region (with two batches in flight the
    # events of one stream would also count the other stream's kernels), CUDA events on the launching stream
    exts[0].set_timing(True)
    for k in range(0, NH * min(K, 16), NH):
        step_resident(k)
    drain()
    tot = (C.c_double * 8)()
    nst = C.c_uint64()
    _lib.check(lib.borb_stage_times_total(exts[0]._h, tot, C.byref(nst)), "borb_stage_times_total")
    exts[0].set_timing(False)
    stage_ms = {n: float(tot[i] / max(nst.value, 1)) for i, n in enumerate(("upload", "pyramid", "fast_nms", "quadtree", "blur", "orient_brief", "stereo", "download"))}
    t_ms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_max = float(t_ms.item())
    value = world * B * K / (ms_max * 1e-3)

    # ---- timed region 2: end to end through the C ABI with HOST buffers (two handles, double-buffered)
    h_in = torch.from_numpy(host).pin_memory()                      # pinned staging of the camera frames
    outs = []
    for _ in range(NH):
        o = dict(kl=torch.empty((B, cap, 28), dtype=torch.uint8).pin_memory(), kr=torch.empty((B, cap, 28), dtype=torch.uint8).pin_memory(),
                 dl=torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory(), dr=torch.empty((B, cap, 32), dtype=torch.uint8).pin_memory(),
                 nl=torch.zeros(B, dtype=torch.int32).pin_memory(), nr=torch.zeros(B, dtype=torch.int32).pin_memory(),
                 ur=torch.empty((B, cap), dtype=torch.float32).pin_memory(), dp=torch.empty((B, cap), dtype=torch.float32).pin_memory())
        outs.append(o)
    ptr_tabs = []
    for j in range(NBUF):
        base = h_in[j].data_ptr()
        pl = (C.c_void_p * B)(*[base + (2 * p) * img_stride for p in range(B)])
        pr = (C.c_void_p * B)(*[base + (2 * p + 1) * img_stride for p in range(B)])
        ptr_tabs.append((pl, pr))

    def step_e2e(k):
        x, o = exts[k % NH], outs[k % NH]
        _lib.check(lib.borb_sync(x._h), "borb_sync")               # previous use of this handle / its host buffers
        pl, pr = ptr_tabs[k % NBUF]
        _lib.check(lib.borb_stereo_frames_enqueue(x._h, pl, pr, B, W_IMG, H_IMG, W_IMG, BF, b, o["kl"].data_ptr(), o["dl"].data_ptr(),
                                                  o["nl"].data_ptr(), o["kr"].data_ptr(), o["dr"].data_ptr(), o["nr"].data_ptr(),
                                                  o["ur"].data_ptr(), o["dp"].data_ptr(), cap), "stereo_frames_enqueue")

    for k in range(max(Wm, 3) * NH):
        step_e2e(k)
    drain()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(K):
        step_e2e(k)
    drain()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    t2 = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * B * K / (float(t2.item()) * 1e-3)
    clk = clocks.stop()
    assert int(outs[0]["nl"].min()) >= NFEAT
    h2d = 2 * B * W_IMG * H_IMG
    d2h = B * (2 * cap * (28 + 32) + 2 * 4 + 2 * cap * 4)

    # ---- per-stream counters gathered over NCCL; the vocabulary checksum proves every rank walks the same tree
    gathered = None
    if world > 1:
        kps0 = outs[0]["dl"][0, :NFEAT].numpy()
        words, _, _ = voc.transform_raw(kps0, 4)
        gathered = sharding.gather_counters([B * K, int(n_left.sum()), int(n_right.sum()), int((outs[0]["ur"] >= 0).sum()),
                                             int(words.astype(np.int64).sum() % (1 << 31))], device=dev).tolist()

    if rank == 0:
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        fast_bytes = LEVEL_PIXELS * 2 * B
        fast_ms = stage_ms["fast_nms"]
        achieved = fast_bytes / (fast_ms * 1e-3) / 1e9 if fast_ms > 0 else 0.0
        traffic = None
        tp = os.path.join(ROOT, "profiles", "fast_traffic.json")
        if os.path.exists(tp):
            tj = json.load(open(tp))
            if tj.get("pairs_per_launch"):
                traffic = tj["dram_bytes_per_launch"] * B / tj["pairs_per_launch"]
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            per = args.cpu_pairs_per_thread
            pool = CpuPool(Ls[:16], Rs[:16], threads)
            fps, dt = pool.run(per)
            kind = pool.kind
            cpu = {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind,
                   "sample": f"{threads} host threads x {per} of the same synthetic KITTI-shaped stereo pairs ({dt:.1f}s wall); "
                             "extract L+R with the reference's ORBextractor.cc compiled verbatim (oracle/_ref) + ComputeStereoMatches restatement"}
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": max(Wm, 3),
                "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": {"workload": "configs[1]: stereo KITTI-00-shaped 1242x375, 2000 feats, extract + ComputeStereoMatches",
                           "pairs_per_step_per_gpu": B, "batches_in_flight": NH, "parallelism": f"{world} independent camera streams, one per GPU (no data-path collective)",
                           "cache": f"inputs larger than L2: {NBUF} rotating batches x {2 * B * W_IMG * H_IMG / 1e6:.0f} MB input + {2 * B * 2 * 1.75:.0f} MB pyramids per step vs 126 MB L2"},
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
                "roofline": {"kernel": "fast_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                             "algorithmic_bytes_per_launch": fast_bytes, "mean_launch_ms": fast_ms},
                "stage_ms_per_step": stage_ms,
                "cpu_baseline": cpu}
        if voc_ms is not None:
            line["nccl"] = {"vocabulary_broadcast_ms": voc_ms, "vocabulary_bytes": voc_bytes,
                            "counter_fields": list(sharding.COUNTER_FIELDS[:4]) + ["vocabulary_word_checksum"],
                            "counters_all_gather": gathered}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=32, help="stereo pairs per step per GPU")
    ap.add_argument("--handles", type=int, default=4, help="batches in flight per GPU (one CUDA stream each)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-pairs-per-thread", type=int, default=8)
    ap.add_argument("--ref-pairs-per-thread", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
