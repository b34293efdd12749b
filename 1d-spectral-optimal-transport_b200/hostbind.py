"""Host-side placement for the host-buffer entry point (`sot_loss_grad_host`): one process per GPU streams ~1 GB per
step each way through pinned host memory, so the thread that drives the copies and the pages it pins should sit on
the NUMA node the GPU hangs off.  `torchrun` leaves every rank on all cores of the box.

Linux only, best effort: everything comes from sysfs / procfs; nothing here touches the GPU.
"""
from __future__ import annotations

import os


def _read(path: str):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of CUDA device `device_index` (from its PCI address), or None when the platform does not say
    (-1 in sysfs: single-node machine or a VM without topology)."""
    import torch
    bus = None
    try:
        props = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    except Exception:
        return None
    node = _read(f"/sys/bus/pci/devices/{bus}/numa_node")
    if node is None or int(node) < 0:
        return None
    return int(node)


def bind_to_gpu_node(device_index: int, spread_if_unknown: bool = True) -> dict:
    """Restrict this process to the CPUs of the GPU's NUMA node (new pinned allocations then land on that node by
    first touch).  When the platform exposes one node only, optionally give each rank its own slice of the cores so
    that eight ranks' copy threads do not migrate over each other.  Returns what was done (for the bench line)."""
    allowed = sorted(os.sched_getaffinity(0))
    node = gpu_numa_node(device_index)
    nodes = [d for d in os.listdir("/sys/devices/system/node")] if os.path.isdir("/sys/devices/system/node") else []
    n_nodes = sum(1 for d in nodes if d.startswith("node") and d[4:].isdigit())
    info = {"bound": False, "gpu_numa_node": node, "numa_nodes": n_nodes, "cpus_before": len(allowed)}
    if node is not None:
        text = _read(f"/sys/devices/system/node/node{node}/cpulist")
        cpus = sorted(_parse_cpulist(text) & set(allowed)) if text else []
        if cpus:
            os.sched_setaffinity(0, cpus)
            info.update(bound=True, cpus_after=len(cpus), how="cpus of the GPU's NUMA node")
            return info
    world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    if spread_if_unknown and world > 1 and len(allowed) >= 2 * world:
        per = len(allowed) // world
        r = int(os.environ.get("LOCAL_RANK", "0")) % world
        cpus = allowed[r * per:(r + 1) * per]
        os.sched_setaffinity(0, cpus)
        info.update(bound=True, cpus_after=len(cpus), how="no NUMA information: an own slice of the cores per rank")
    return info
