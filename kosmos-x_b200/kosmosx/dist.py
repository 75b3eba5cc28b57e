"""Data-parallel plumbing for the forward path.

The forward shards over the batch dimension with no data-path collective: rank r owns a
contiguous slice of the global batch, weights are replicated.  torch.distributed (NCCL on the
GPU box, gloo in CPU tests) is used only for rendezvous, barriers and max-over-ranks timing.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """[lo, hi) of the global batch owned by `rank`; contiguous, sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(global_batch, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(text_tokens: torch.Tensor, images: torch.Tensor, rank: int, world: int):
    lo, hi = shard_range(text_tokens.shape[0], rank, world)
    return text_tokens[lo:hi], images[lo:hi]


def init_from_env(backend: str | None = None):
    """One process per GPU, launched by torchrun (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        dist.init_process_group(backend or ("nccl" if torch.cuda.is_available() else "gloo"), rank=rank,
                                world_size=world)
    return rank, local, world


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_logits(local_logits: torch.Tensor, global_batch: int):
    """Reassemble the global (B, T, V) logits on every rank (test / debugging helper — the forward
    itself never needs it)."""
    if not dist.is_initialized():
        return local_logits
    world, rank = dist.get_world_size(), dist.get_rank()
    outs = []
    for r in range(world):
        lo, hi = shard_range(global_batch, r, world)
        buf = local_logits if r == rank else torch.empty((hi - lo,) + tuple(local_logits.shape[1:]),
                                                         dtype=local_logits.dtype, device=local_logits.device)
        dist.broadcast(buf, src=r)
        outs.append(buf)
    return torch.cat(outs, 0)


def bind_host_thread_to_gpu(device_index: int) -> str:
    """Pin the calling thread (and the threads it creates later) to the CPUs NVML reports as closest to the GPU, so that
    pinned host buffers allocated afterwards — and the copies into / out of them — stay on the GPU's own NUMA node.
    On an 8-GPU box the ranks otherwise land on arbitrary cores and half of the device<->host traffic crosses the
    socket interconnect.  Returns a short description; never raises (no NVML = no binding)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            props = torch.cuda.get_device_properties(device_index)
            bus = "%08X:%02X:%02X.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return "NVML reports no CPU affinity for this GPU inside the allowed set; not bound"
        os.sched_setaffinity(0, cpus)
        return f"bound to {len(cpus)} CPUs near GPU {device_index} (NVML affinity)"
    except Exception as e:                                   # noqa: BLE001 — strictly best effort
        return f"not bound ({type(e).__name__}: {e})"
