#!/usr/bin/env python3
"""Where do sa_host_alloc's pages land?  Runs the calling thread on the far NUMA node (plus a single CPU
of the GPU's node, so the allocator is allowed to go there), allocates, and reads /proc/self/numa_maps
for the buffer (N<node>=<pages>); a torch pin_memory() buffer allocated the same way is the control."""
import ctypes, glob, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "stark-anatomy_b200"), ROOT]
import torch
import sa_engine
eng = sa_engine.get_engine()
lib = eng.lib
pr = torch.cuda.get_device_properties(0)
bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
gpu_node = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
def cpus(text):
    out = set()
    for part in text.strip().split(","):
        a, _, b = part.partition("-")
        out.update(range(int(a), int(b or a) + 1))
    return out
nodes = {os.path.basename(n)[4:]: cpus(open(n + "/cpulist").read()) for n in glob.glob("/sys/devices/system/node/node[0-9]*")}
far = [k for k in nodes if k != gpu_node]
print("gpu", bdf, "numa node", gpu_node, "nodes", sorted(nodes))
def where(ptr):
    for line in open("/proc/self/numa_maps"):
        start = int(line.split()[0], 16)
        if start <= ptr < start + (64 << 20) and ("N0=" in line or "N1=" in line) and start == ptr & ~0xFFF:
            return line.strip()[:200]
    best = None
    for line in open("/proc/self/numa_maps"):
        start = int(line.split()[0], 16)
        if start <= ptr and (best is None or start > best[0]):
            best = (start, line.strip()[:200])
    return best[1] if best else None
if far:
    allowed = os.sched_getaffinity(0)
    mask = (nodes[far[0]] & allowed) | {min(nodes[gpu_node] & allowed)}
    os.sched_setaffinity(0, mask)
    print("calling thread now on node", far[0], "+ cpu", min(nodes[gpu_node] & allowed), "of node", gpu_node)
p = lib.sa_host_alloc(64 << 20)
print("sa_host_alloc :", where(p))
t = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
print("torch pinned  :", where(t.data_ptr()))
print("affinity restored:", os.sched_getaffinity(0) == mask if far else None)
lib.sa_host_free(p)
