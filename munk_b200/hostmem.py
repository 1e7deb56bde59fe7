"""Where page-locked host buffers land.

The results of the pipeline leave the GPU by DMA into pinned host memory (HostSink, ShardSink: up to
8.3 GB per pipeline at 20k/16k nodes).  The pages of a pinned allocation are taken from the NUMA
node of the CPU that allocates them; a rank whose process happens to run on the other socket gets
buffers its GPU can only reach across the socket interconnect, and on an 8-GPU box all such ranks
share that link.  `numa_local(index)` binds the calling thread to the CPUs next to GPU `index` for
the duration of the allocation and restores the previous affinity afterwards (threads started
meanwhile would keep the binding - nothing here starts any).

Host-side placement only: no arithmetic, nothing the reference has a counterpart for.  Every
lookup failure (no sysfs, one node, a cpuset that excludes the node, MUNK_NUMA_BIND=0) means "do
not bind", never an error.
"""
import contextlib
import os


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the format of /sys/devices/system/node/*/cpulist)."""
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_local_cpus(index):
    """CPUs of the NUMA node GPU `index` hangs off (sysfs by PCI address, else NVML), or None."""
    import torch
    props = torch.cuda.get_device_properties(index)
    bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
    try:
        with open("/sys/bus/pci/devices/%s/numa_node" % bdf) as fh:
            node = int(fh.read())
        if node >= 0:
            with open("/sys/devices/system/node/node%d/cpulist" % node) as fh:
                return parse_cpulist(fh.read()) or None
    except (OSError, ValueError):
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bdf)
        except TypeError:
            handle = pynvml.nvmlDeviceGetHandleByPciBusId(bdf.encode())
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        return cpus or None
    except Exception:
        return None


@contextlib.contextmanager
def bound_to(cpus):
    """Calling thread restricted to `cpus` (intersected with what it may use now) inside the block;
    previous affinity restored on exit.  No-op for None, an empty intersection, or no change."""
    previous = None
    try:
        if cpus and hasattr(os, "sched_getaffinity"):
            now = os.sched_getaffinity(0)
            target = set(cpus) & now
            if target and target != now:
                os.sched_setaffinity(0, target)
                previous = now
    except OSError:
        previous = None
    try:
        yield previous is not None
    finally:
        if previous is not None:
            try:
                os.sched_setaffinity(0, previous)
            except OSError:
                pass


def numa_local(index):
    """Context manager: allocate pinned host memory next to GPU `index` (see module docstring)."""
    if os.environ.get("MUNK_NUMA_BIND", "1") == "0":
        return bound_to(None)
    try:
        cpus = gpu_local_cpus(index)
    except Exception:
        cpus = None
    return bound_to(cpus)
