"""Multi-GPU sharding helpers for the benchmark and for callers that feed several GPUs of one box.

The hot path has no exchange step: a code block depends only on its own LLRs and soft buffer, a subframe only on its own
samples.  So ranks just take disjoint shards; the only collective is the max/sum of a few scalars
for reporting, which works on any torch.distributed backend (nccl on the GPU box, gloo in the CPU tests)."""
from __future__ import annotations

import os


def rank_info() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when launched plainly."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_units: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous [first, last) slice of n_units (cells, subframes, code blocks) owned by `rank`; sizes differ by <= 1."""
    first = (n_units * rank) // world
    last = (n_units * (rank + 1)) // world
    return first, last


def cell_to_rank(cell: int, world: int) -> int:
    """Config 5 of BASELINE.json: cell c lives on GPU c mod G, so its HARQ soft buffers stay resident there."""
    return cell % world


def shard_seed(base_seed: int, rank: int) -> int:
    return base_seed + 0x9E3779B1 * rank


def reduce_scalars(values: list[float], op: str, dist=None, device=None) -> list[float]:
    """max / sum of per-rank scalars over all ranks (identity without a process group)."""
    if dist is None or not dist.is_initialized():
        return list(values)
    import torch

    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX if op == "max" else dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def aggregate_throughput(units_per_rank: float, ms_this_rank: float, dist=None, device=None) -> tuple[float, float]:
    """Whole-job throughput the way the bench contract defines it: all ranks' units / max-over-ranks time.
    Returns (units per second, max ms)."""
    (ms_max,) = reduce_scalars([ms_this_rank], "max", dist, device)
    (units,) = reduce_scalars([units_per_rank], "sum", dist, device)
    return units / (ms_max * 1e-3), ms_max


def bind_to_gpu_numa_node(local_rank: int) -> dict:
    """Pins this process to the CPUs next to its GPU (sysfs: the PCI device's local_cpulist) BEFORE it allocates page-locked
    host memory, so that the buffers every rank feeds its GPU from live on the memory of the GPU's own NUMA node instead of all
    ranks sharing node 0.  Returns what was found (reported in the bench line); never fails."""
    info = {"bound": False}
    try:
        import torch

        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        info["pci"] = bdf
        node = open(os.path.join(base, "numa_node")).read().strip()
        cpus = open(os.path.join(base, "local_cpulist")).read().strip()
        info["numa_node"], info["local_cpulist"] = int(node), cpus
        want = set()
        for part in cpus.split(","):
            if "-" in part:
                a, b = part.split("-")
                want.update(range(int(a), int(b) + 1))
            elif part:
                want.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = want & allowed
        info["cpus_allowed"] = len(allowed)
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            info["bound"] = True
            info["cpus_used"] = len(use)
    except Exception as ex:  # containers often hide sysfs; the benchmark then runs unbound
        info["note"] = f"not bound: {type(ex).__name__}"
    return info
