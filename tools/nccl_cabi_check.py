"""2+ GPU check of the torch-free NCCL entry points of the C ABI (borb_nccl_unique_id / borb_nccl_comm_create / borb_voc_broadcast).
Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_cabi_check.py
torch.distributed (gloo) is used ONLY to hand the 128-byte ncclUniqueId to the other ranks and to compare checksums — what a C++
host does over its own control channel.  Rank 0 builds the vocabulary; every rank receives it through ncclBroadcast issued by
libborb and must compute the same words for the same descriptors."""
import ctypes as C
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch                                      # noqa: E402
import torch.distributed as dist                  # noqa: E402
from orb_slam2_b200 import _lib, matcher as M, sharding    # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    lib = _lib.load()
    uid = np.zeros(128, np.uint8)
    if rank == 0:
        _lib.check(lib.borb_nccl_unique_id(uid.ctypes.data), "borb_nccl_unique_id")
    t = torch.from_numpy(uid)
    dist.broadcast(t, 0)
    comm = C.c_void_p()
    _lib.check(lib.borb_nccl_comm_create(uid.ctypes.data, world, rank, local, C.byref(comm)), "borb_nccl_comm_create")
    voc = None
    if rank == 0:
        voc = M.ORBVocabulary.from_arrays(*sharding.random_vocabulary_arrays(10, 6, 7), 10, 6, device=local)
    out = C.c_void_p()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(lib.borb_voc_broadcast(voc._h if voc else None, comm, 0, rank, local, C.byref(out)), "borb_voc_broadcast")
    ms = (time.perf_counter() - t0) * 1e3
    v = voc if rank == 0 else M.ORBVocabulary(out, lib)
    rng = np.random.default_rng(3)
    d = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    w, wt, nd = v.transform_raw(d, 4)
    chk = torch.tensor([int(w.astype(np.int64).sum()), int(nd.astype(np.int64).sum()), int(v.blob()[1])], dtype=torch.int64)
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    ok = all(torch.equal(allc[0], c) for c in allc)
    tm = torch.tensor([ms], dtype=torch.float64); allt = [torch.zeros_like(tm) for _ in range(world)]
    dist.all_gather(allt, tm)
    if rank == 0:
        print({"world": world, "ok": bool(ok), "blob_MB": int(chk[2]) / 1e6, "broadcast_ms_max": max(float(x) for x in allt), "checksums": [c.tolist() for c in allc]}, flush=True)
    _lib.check(lib.borb_nccl_comm_destroy(comm), "borb_nccl_comm_destroy")
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
