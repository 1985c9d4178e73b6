"""torch.distributed plumbing for multi-GPU runs: one process per GPU, no collective on the compute path.

The FIR path shards without any exchange (each output depends only on tapCount consecutive inputs,
ref: src/fir.cu:64-70), so ranks only meet for (a) timing (max over ranks) and (b) the optional gather of the
decimated outputs to rank 0 (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os
from typing import List, Optional

import torch
import torch.distributed as dist

from . import api


def init_from_env(backend: Optional[str] = None) -> tuple[int, int, int]:
    """Initialises the default process group from RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun's env).
    Returns (rank, world_size, local_rank); a no-op single-rank answer when WORLD_SIZE is unset or 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa_node(device_index: int) -> Optional[int]:
    """Pins this process to the CPUs of the NUMA node the GPU hangs off, so that pinned host buffers allocated
    afterwards are local to the GPU's PCIe root (host-buffer runs with several ranks are bound by host memory
    placement, not by the GPUs).  Returns the node, or None when the topology cannot be read (nothing is changed)."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus: set[int] = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None


def time_shard(num_outputs: int, decimation: int, tap_count: int, first_sample_index: int, world: int, rank: int):
    """This rank's contiguous block of outputs of one long capture (gsdrShardPlanTime)."""
    return api.shard_plan_time(num_outputs, decimation, tap_count, first_sample_index, world, rank)


def channel_shard(num_channels: int, world: int, rank: int) -> tuple[int, int]:
    return api.shard_plan_channels(num_channels, world, rank)


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_outputs(local: torch.Tensor, counts: List[int], dst: int = 0) -> Optional[torch.Tensor]:
    """Gathers the ranks' output blocks (lengths `counts`, known to every rank from the shard plan) to rank `dst`
    in shard order.  Blocks are padded to the longest so a single gather collective moves everything."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    assert len(counts) == world and local.shape[0] == counts[rank]
    longest = max(counts)
    is_complex = local.is_complex()
    flat = torch.view_as_real(local) if is_complex else local
    pad_shape = (longest,) + tuple(flat.shape[1:])
    send = torch.zeros(pad_shape, dtype=flat.dtype, device=flat.device)
    send[: counts[rank]] = flat
    recv = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
    dist.gather(send, recv, dst=dst)
    if rank != dst:
        return None
    out = torch.cat([recv[r][: counts[r]] for r in range(world)], dim=0)
    return torch.view_as_complex(out.contiguous()) if is_complex else out


def gather_outputs_p2p(local: torch.Tensor, counts: List[int], firsts: List[int], full: Optional[torch.Tensor],
                       dst: int = 0) -> Optional[torch.Tensor]:
    """Collects the ranks' output blocks on rank `dst` with grouped point-to-point operations (NCCL: ncclGroupStart,
    one ncclSend per sender / one ncclRecv per sender on `dst`, ncclGroupEnd): every block is received straight at its
    final offset `firsts[r]` of `full` (preallocated on `dst`, None elsewhere) — no padding, no concatenation.  The
    rank that owns `dst` copies its own block on the device.  Works with gloo too (CPU tests)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        if full is not None:
            full[firsts[0]: firsts[0] + counts[0]].copy_(local)
        return full
    world, rank = dist.get_world_size(), dist.get_rank()
    assert len(counts) == world and len(firsts) == world and local.shape[0] == counts[rank]
    ops = []
    if rank == dst:
        assert full is not None
        for r in range(world):
            if r != dst and counts[r]:
                ops.append(dist.P2POp(dist.irecv, full[firsts[r]: firsts[r] + counts[r]], r))
        full[firsts[dst]: firsts[dst] + counts[dst]].copy_(local, non_blocking=True)
    elif counts[rank]:
        ops.append(dist.P2POp(dist.isend, local, dst))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return full


class PeerOutput:
    """An output buffer on rank 0's GPU that every rank's kernels can store into (gsdrSharedBuffer*: cudaIpc mapping
    over NVLink).  rank 0 creates it and broadcasts the 64-byte handle; the others open it.  `ptr` is the address
    valid in THIS process — pass `ptr + firstOutput * 8` as `output` of gsdrFirFC to gather while computing."""

    def __init__(self, nbytes: int, rank: int, local_device: int, device):
        self.rank, self.local = rank, local_device
        self.nbytes = nbytes
        handle = torch.zeros(64, dtype=torch.uint8, device=device)
        if rank == 0:
            self.ptr, raw = api.shared_buffer_create(nbytes, local_device)
            handle.copy_(torch.frombuffer(bytearray(raw), dtype=torch.uint8))
        if dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(handle, src=0)
        if rank != 0:
            self.ptr = api.shared_buffer_open(bytes(handle.cpu().numpy().tobytes()), local_device)

    def as_tensor(self, count: int) -> torch.Tensor:
        """rank 0 only: the buffer as a complex64 tensor (copied out through a raw device-to-device memcpy)."""
        import ctypes

        out = torch.empty(count, dtype=torch.complex64, device=torch.device("cuda", self.local))
        rt = ctypes.CDLL("libcudart.so.12")
        rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
        rc = rt.cudaMemcpy(out.data_ptr(), self.ptr, count * 8, 3)  # cudaMemcpyDeviceToDevice
        if rc:
            raise api.CudaError(rc, "cudaMemcpy")
        return out

    def close(self) -> None:
        if self.ptr:
            if self.rank == 0:
                api.shared_buffer_destroy(self.ptr, self.local)
            else:
                api.shared_buffer_close(self.ptr, self.local)
            self.ptr = 0
