"""Multi-GPU plumbing: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch) as the collective.

The C library shards the database itself (ssa_b200_set_shard) and needs exactly one collective per search:
an all-gather of each rank's shard scores (4 bytes per local record).  It calls back into the function
registered here with raw device pointers and its CUDA stream; we wrap the pointers as torch tensors
(zero copy, __cuda_array_interface__) and issue dist.all_gather_into_tensor on that stream.
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


def init_from_env(lib, backend="nccl"):
    """Reads RANK/LOCAL_RANK/WORLD_SIZE/MASTER_* (torchrun), initialises torch.distributed and shards `lib`."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
    lib.set_device(local)
    lib.set_shard(rank, world)
    if world > 1:
        lib.set_allgather(make_torch_allgather(torch.device("cuda", local)))
    return rank, world, local
