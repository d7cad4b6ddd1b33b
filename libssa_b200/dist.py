"""Multi-GPU plumbing, one process per GPU (torchrun): the C library shards the database itself and merges the shard
scores itself - ncclAllGather on its own CUDA stream, on a communicator it creates inside the library
(ssa_b200_nccl_unique_id / ssa_b200_nccl_init, include/libssa_b200.h).  The only thing the host has to do is carry rank 0's
128-byte NCCL id to the other ranks; here torch.distributed does that (and provides the barrier / reductions bench.py
needs).  `collective="torch"` keeps the older route for hosts that bring their own collective: the library calls back
into dist.all_gather_into_tensor with raw device pointers and its stream (ssa_b200_set_allgather).
"""
import os


class _DevicePtr:
    """Zero-copy view of `nbytes` bytes of device memory for torch.as_tensor."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def make_torch_allgather(device):
    import torch
    import torch.distributed as dist

    def allgather(send_ptr, recv_ptr, nbytes, stream_ptr):
        world = dist.get_world_size()
        send = torch.as_tensor(_DevicePtr(send_ptr, nbytes), device=device)
        recv = torch.as_tensor(_DevicePtr(recv_ptr, nbytes * world), device=device)
        ext = torch.cuda.ExternalStream(int(stream_ptr), device=device)
        with torch.cuda.stream(ext):
            dist.all_gather_into_tensor(recv, send)

    return allgather


def init_from_env(lib, backend="nccl", collective=None):
    """Reads RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* (torchrun), initialises torch.distributed and shards `lib`.
    collective: "nccl" (default; $LIBSSA_B200_COLLECTIVE overrides) = the library's own NCCL communicator,
    "torch" = callback into torch.distributed."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    collective = collective or os.environ.get("LIBSSA_B200_COLLECTIVE", "nccl")
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
    lib.set_device(local)
    if world > 1 and collective == "nccl":
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            ident = torch.tensor(list(lib.nccl_unique_id()), dtype=torch.uint8)
        ident = ident.to(torch.device("cuda", local))
        dist.broadcast(ident, 0)
        lib.nccl_init(bytes(ident.cpu().tolist()), rank, world)
    else:
        lib.set_shard(rank, world)
        if world > 1:
            lib.set_allgather(make_torch_allgather(torch.device("cuda", local)))
    return rank, world, local
