"""Intrinsic latency of the NVLink peer-store ParRepartition: back-to-back exchanges of the truncated C4 tensor (20,16,16,nz,8) ComplexF32,
z-slabs -> t-mode slabs, nothing else on the stream (ranks stay in lockstep, so rank skew is not in the number).
Usage (GPU box): python scripts/bench_exchange.py WORLD"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def worker(rank, world, port, peer):
    os.environ["PO_NO_PEER"] = "0" if peer else "1"
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, init_method="tcp://127.0.0.1:%d" % port, device_id=torch.device("cuda", rank))
    import parops
    eng = parops.default_engine()
    parops.init_comm(eng)
    nz = 128 if world == 8 else 64
    gs = (20, 16, 16, nz, 8)
    R = parops.ParRepartition(parops.ComplexF32, parops.CartComm([1, 1, 1, world, 1]), parops.CartComm([1, 1, 1, 1, world]), gs)
    x = torch.randn(R.Domain, 2, device="cuda").view(torch.complex64) if False else torch.view_as_complex(torch.randn(R.Domain, 2, device="cuda")).reshape(1, R.Domain).t()
    for _ in range(10):
        y = R * x
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 200
    e0.record()
    for _ in range(n):
        y = R * x
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    t = torch.tensor([us], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        remote = R.Domain * 8 * (world - 1) / world
        print("world %d %s: %.1f us per exchange (max over ranks), %.2f MB sent per rank -> %.0f GB/s per direction" % (
            world, "peer-store kernel" if peer else "pack + grouped NCCL send/recv + unpack", float(t), remote / 1e6, remote / float(t) / 1e3))
    dist.barrier(); dist.destroy_process_group()


if __name__ == "__main__":
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    import socket
    for peer in (1, 0):
        s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
        mp.spawn(worker, args=(world, port, peer), nprocs=world, join=True)
