"""Host-side NUMA placement: one process per GPU works best when the process (and therefore the pinned
host buffers it allocates by first touch, and the threads that fill them) lives on the CPU socket the
GPU's PCIe root hangs off — an MPI launcher's `--bind-to` does this for the reference; `torchrun` does
not. On a 2-socket B200 box GPUs 0-3 / 4-7 sit on NUMA nodes 0 / 1: a buffer on the wrong node crosses
the socket interconnect on every host<->device copy.

`bind_to_gpu(deviceid)` restricts the calling process to the GPU's local CPUs (`local_cpulist` of its PCI
device in sysfs). Opt out with B2_NUMA_BIND=0. It is applied by `init_distributed()`, by `bench.py` and
before the first pinned host allocation of a field.
"""
import ctypes
import os

_done = {}


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_local_cpus(deviceid):
    """CPUs local to the GPU's PCIe root, or None when sysfs does not say."""
    from ._lib import load_library
    L = load_library()
    buf = ctypes.create_string_buffer(64)
    if L.b2_device_pci_bus_id(int(deviceid), buf, 64) != 0:
        return None
    bus = buf.value.decode().lower()
    for cand in (bus, bus[-12:]):
        path = f'/sys/bus/pci/devices/{cand}/local_cpulist'
        if os.path.exists(path):
            try:
                with open(path) as f:
                    cpus = _parse_cpulist(f.read())
                return cpus or None
            except OSError:
                return None
    return None


def bind_to_gpu(deviceid):
    """Pin this process to the CPUs next to GPU `deviceid`. Returns a dict describing what was done."""
    deviceid = int(deviceid)
    if deviceid in _done:
        return _done[deviceid]
    info = {'device': deviceid, 'bound': False, 'cpus': None}
    if os.environ.get('B2_NUMA_BIND', '1') == '0' or not hasattr(os, 'sched_setaffinity'):
        _done[deviceid] = info
        return info
    try:
        local = gpu_local_cpus(deviceid)
        if local:
            allowed = os.sched_getaffinity(0)
            target = local & allowed
            if target and target != allowed:
                os.sched_setaffinity(0, target)
                info['bound'] = True
            info['cpus'] = len(target or allowed)
    except Exception as e:                      # placement is an optimisation, never an error
        info['error'] = str(e)
    _done[deviceid] = info
    return info
