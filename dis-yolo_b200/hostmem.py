"""Host-side placement for the host-buffer calls (dy_forward_host*): pinned staging memory should live on
the NUMA node the GPU's PCIe root hangs off.  With one process per GPU (torchrun) and no binding, every
rank's pinned buffers are first-touched wherever the scheduler happened to run the process, and at 8 GPUs
most H2D / D2H traffic then crosses the socket interconnect through one node's memory controllers
(round-1 SCALE: end-to-end efficiency 0.29 at N=8 while the device-timed number scaled at 0.99).

bind_to_gpu_numa(device) pins the calling process to the CPUs of the GPU's NUMA node (Linux sysfs);
memory allocated afterwards (cudaHostAlloc / torch pin_memory: first touch) is then node-local.
It is best effort: on a single-node host, or when sysfs does not expose the topology, it does nothing.
"""
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(text):
    cpus = set()
    for part in (text or '').split(','):
        part = part.strip()
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_pci_bus_id(device):
    """'0000:1b:00.0' of CUDA device ordinal `device` (honours CUDA_VISIBLE_DEVICES), or None."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        return '%04x:%02x:%02x.0' % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None


def gpu_numa_node(device):
    bus = gpu_pci_bus_id(device)
    if bus is None:
        return None
    node = _read('/sys/bus/pci/devices/%s/numa_node' % bus)
    try:
        node = int(node)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def bind_to_gpu_numa(device):
    """Returns a dict describing what was done (reported by bench.py)."""
    info = dict(bound=False, node=None, cpus=0)
    node = gpu_numa_node(device)
    if node is None:
        return info
    cpus = _parse_cpulist(_read('/sys/devices/system/node/node%d/cpulist' % node))
    try:
        allowed = os.sched_getaffinity(0)
    except (AttributeError, OSError):
        return info
    cpus &= allowed
    if not cpus or cpus == allowed:
        info.update(node=node, cpus=len(cpus))
        return info
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return info
    info.update(bound=True, node=node, cpus=len(cpus))
    return info
