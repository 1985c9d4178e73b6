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
