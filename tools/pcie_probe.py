#!/usr/bin/env python3
"""Host<->device copy bandwidth of pinned buffers by the NUMA node they were allocated on (first touch
under a CPU affinity mask): tells whether the e2e path is limited by the link or by buffer placement."""
import glob, json, os, time
import torch

dev = torch.device("cuda:0")
props = torch.cuda.get_device_properties(dev)
bdf = "%04x:%02x:%02x.0" % (props.pci_domain_id, props.pci_bus_id, props.pci_device_id)
def rd(p):
    try:
        return open(p).read().strip()
    except OSError:
        return None
info = {"gpu": props.name, "bdf": bdf, "gpu_numa_node": rd("/sys/bus/pci/devices/%s/numa_node" % bdf),
        "gpu_local_cpulist": rd("/sys/bus/pci/devices/%s/local_cpulist" % bdf),
        "link_speed": rd("/sys/bus/pci/devices/%s/current_link_speed" % bdf),
        "link_width": rd("/sys/bus/pci/devices/%s/current_link_width" % bdf),
        "nodes": {os.path.basename(n): rd(n + "/cpulist") for n in sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))}}
print(json.dumps(info))

def parse(cl):
    out = set()
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-")
            out.update(range(int(a), int(b) + 1))
        elif part:
            out.add(int(part))
    return out

N = 256 << 20
d_in = torch.empty(N, dtype=torch.uint8, device=dev)
d_out = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
all_cpus = os.sched_getaffinity(0)
for name, cl in list(info["nodes"].items()) + [("any", None)]:
    if cl is not None:
        cpus = parse(cl) & all_cpus
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
    else:
        os.sched_setaffinity(0, all_cpus)
    h_in = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(N, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(2)
    res = {"alloc_on": name}
    def timed(fn, reps=6):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    t = timed(lambda: d_in.copy_(h_in, non_blocking=True))
    res["h2d_GBps"] = round(N / t / 1e9, 1)
    t = timed(lambda: h_out.copy_(d_out, non_blocking=True))
    res["d2h_GBps"] = round(N / t / 1e9, 1)
    def both():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    t = timed(both)
    res["bidir_each_GBps"] = round(N / t / 1e9, 1)
    print(json.dumps(res))
    del h_in, h_out
