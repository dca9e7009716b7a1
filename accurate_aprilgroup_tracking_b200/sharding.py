"""Multi-GPU partitioning of the path: independent frames are split into
contiguous blocks, camera streams are pinned to GPU ``stream mod world`` (frames inside a
stream are sequential: the predictor needs poses t-1, t-2 and LK needs frame t-1).  There is
no collective on the path; the only exchange is an all-gather of the final [n,6] poses."""
from __future__ import annotations

import contextlib
import os
from typing import List, Optional, Tuple


def frame_block(n_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """[start, stop) of the frames rank owns; sizes differ by at most one."""
    if world < 1 or not 0 <= rank < world or n_frames < 0:
        raise ValueError("bad partition request")
    base, extra = divmod(n_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def stream_owner(stream_id: int, world: int) -> int:
    return stream_id % world


def local_streams(n_streams: int, rank: int, world: int) -> List[int]:
    return [s for s in range(n_streams) if stream_owner(s, world) == rank]


def gather_poses(local, n_total: int, group=None):
    """All-gather block-partitioned poses: ``local`` is this rank's [n_local,6] tensor (block of
    frame_block()); returns the [n_total,6] tensor in global frame order on every rank.
    Works on any torch.distributed backend (NCCL over NVLink on the GPU box, gloo in CPU tests)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    sizes = [frame_block(n_total, r, world) for r in range(world)]
    width = max(b - a for a, b in sizes)
    pad = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[: b - a] for o, (a, b) in zip(out, sizes)], dim=0)


def gather_stream_poses(local, n_streams: int, group=None):
    """All-gather per-stream poses ([n_local_streams,6], streams s with s % world == rank, ascending)
    into [n_streams,6] in stream order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world = dist.get_world_size(group)
    width = (n_streams + world - 1) // world
    pad = torch.zeros((width, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    full = torch.zeros((n_streams, local.shape[1]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ids = local_streams(n_streams, r, world)
        full[ids] = out[r][: len(ids)]
    return full


# ---- host side of one-process-per-GPU: keep a rank's pinned frame buffers on its GPU's NUMA node -------------------
def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (sysfs cpulist syntax) -> [0, 1, 2, 3, 8, 10, 11]."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int, sysfs: str = "/sys") -> Optional[int]:
    """NUMA node the GPU's PCIe root port hangs off (None when the platform does not say)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        addr = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"{sysfs}/bus/pci/devices/{addr}/numa_node") as fh:
            node = int(fh.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int, sysfs: str = "/sys") -> Optional[int]:
    """Restrict this process to the CPUs of the NUMA node its GPU is attached to, so that host frame buffers pinned
    afterwards are allocated there (first touch) and the GPU reads them without crossing the socket interconnect -
    with eight ranks gathering regions of interest out of host memory at once, buffers that all sit on one socket
    saturate that socket's memory and the inter-socket link.  Returns the node, or None if nothing was changed
    (single-node host, unknown topology, CPUs not available to this process)."""
    node = gpu_numa_node(device_index, sysfs)
    if node is None:
        return None
    try:
        with open(f"{sysfs}/devices/system/node/node{node}/cpulist") as fh:
            cpus = set(parse_cpulist(fh.read())) & set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


@contextlib.contextmanager
def on_gpu_numa_node(device_index: int, sysfs: str = "/sys"):
    """``with on_gpu_numa_node(local_rank) as node: buf = torch.empty(..., pin_memory=True)``: allocate host buffers on
    the GPU's NUMA node, then give the thread its previous CPU set back (pinned pages do not migrate)."""
    before = os.sched_getaffinity(0)
    node = bind_to_gpu_numa_node(device_index, sysfs)
    try:
        yield node
    finally:
        if node is not None:
            os.sched_setaffinity(0, before)
