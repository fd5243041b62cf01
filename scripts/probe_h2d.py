"""Probe: host topology seen by this process and pinned H2D / D2H copy bandwidth (one GPU)."""
import glob
import os
import time

import torch

print("cpus allowed:", sorted(os.sched_getaffinity(0)))
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    try:
        print(os.path.basename(n), open(n + "/cpulist").read().strip())
    except OSError:
        pass
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    p = "/sys/bus/pci/devices/" + bus.lower()[-12:] + "/numa_node"
    print("gpu0", bus, "numa_node", open(p).read().strip() if os.path.exists(p) else "?")
    print("pcie gen/width", pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), pynvml.nvmlDeviceGetCurrPcieLinkWidth(h))
except Exception as e:
    print("nvml:", e)

torch.cuda.init()
nbytes = 1 << 30
d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")


def bw(host, label):
    s = torch.cuda.Stream()
    for direction in ("h2d", "d2h"):
        with torch.cuda.stream(s):
            for _ in range(2):
                d.copy_(host, non_blocking=True) if direction == "h2d" else host.copy_(d, non_blocking=True)
            s.synchronize()
            t = time.perf_counter()
            for _ in range(5):
                d.copy_(host, non_blocking=True) if direction == "h2d" else host.copy_(d, non_blocking=True)
            s.synchronize()
            print(label, direction, "%.1f GB/s" % (5 * nbytes / (time.perf_counter() - t) / 1e9))


host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
host.fill_(1)
bw(host, "default placement")
allowed = sorted(os.sched_getaffinity(0))
for n in sorted(glob.glob("/sys/devices/system/node/node*")):
    try:
        cl = open(n + "/cpulist").read().strip()
    except OSError:
        continue
    cpus = set()
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        elif part:
            cpus.add(int(part))
    use = cpus & set(allowed)
    if not use:
        print(os.path.basename(n), "no allowed cpus")
        continue
    os.sched_setaffinity(0, use)
    h2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h2.fill_(1)                                   # first touch on this node
    bw(h2, "first-touch on " + os.path.basename(n))
    del h2
    os.sched_setaffinity(0, allowed)
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
host2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        d.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2):
        host2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print("bidirectional: %.1f GB/s each way" % (5 * nbytes / dt / 1e9))
