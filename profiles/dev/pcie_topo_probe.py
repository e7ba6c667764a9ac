"""Why does the host-buffer (e2e) rate stop scaling past ~2 GPUs?  Prints the box's PCIe/NUMA topology and measures
pinned D2H bandwidth for single GPUs, pairs and all GPUs at once, with the pinned buffers bound to each NUMA node.
    gpurun --gpus 8 -- python profiles/dev/pcie_topo_probe.py"""
import ctypes, glob, os, subprocess, time
import torch

def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=20).stdout.strip()
    except Exception as e:
        return "ERR %s" % e

print("== nproc", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:64])
print("== numa nodes:", sh("ls -d /sys/devices/system/node/node* 2>/dev/null | tr '\\n' ' '"))
print(sh("lscpu | grep -i -E 'model name|socket|numa|^CPU\\(s\\)'"))
print("== mems allowed:", sh("grep -i mems_allowed_list /proc/self/status"))
print(sh("nvidia-smi topo -m"))
n = torch.cuda.device_count()
for i in range(n):
    bdf = torch.cuda.get_device_properties(i).pci_bus_id if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else None
    print("gpu", i, bdf)
print(sh("for d in /sys/bus/pci/devices/*; do if [ \"$(cat $d/vendor)\" = 0x10de ]; then echo $d $(cat $d/class) numa=$(cat $d/numa_node) $(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null); fi; done"))

libc = ctypes.CDLL(None, use_errno=True)
MPOL_DEFAULT, MPOL_BIND, MPOL_PREFERRED = 0, 2, 1
def set_mempolicy(mode, node):
    if node is None:
        return libc.syscall(238, MPOL_DEFAULT, None, 0)
    mask = ctypes.c_ulong(1 << node)
    return libc.syscall(238, mode, ctypes.byref(mask), 65)  # x86_64 __NR_set_mempolicy = 238

nodes = sorted(int(p.rsplit("node", 1)[1]) for p in glob.glob("/sys/devices/system/node/node[0-9]*"))
SZ = 1 << 30
dbuf = [torch.empty(SZ, dtype=torch.uint8, device="cuda:%d" % i) for i in range(n)]
streams = [torch.cuda.Stream(device=i) for i in range(n)]
def pinned_on(node):
    rc = set_mempolicy(MPOL_BIND, node)
    t = torch.empty(SZ, dtype=torch.uint8).pin_memory()
    t.fill_(1)
    set_mempolicy(MPOL_DEFAULT, None)
    return t, rc

def run(gpus, hbufs, reps=6, h2d=False):
    for i in gpus:
        torch.cuda.synchronize(i)
    t0 = time.perf_counter()
    for _ in range(reps):
        for i in gpus:
            with torch.cuda.stream(streams[i]):
                if h2d: dbuf[i].copy_(hbufs[i], non_blocking=True)
                else: hbufs[i].copy_(dbuf[i], non_blocking=True)
    for i in gpus:
        streams[i].synchronize()
    dt = time.perf_counter() - t0
    return len(gpus) * reps * SZ / dt / 1e9

print("== default-policy pinned buffers")
hb, _ = zip(*[pinned_on(None) for _ in range(n)])
for g in range(n):
    print("D2H gpu%d alone: %.1f GB/s   H2D: %.1f" % (g, run([g], hb), run([g], hb, h2d=True)))
for pair in [(0, 1), (0, 2), (0, 3), (0, 4), (0, 7), (2, 3), (4, 5), (6, 7)]:
    if max(pair) < n:
        print("D2H pair %s: %.1f GB/s total" % (pair, run(list(pair), hb)))
for k in (2, 4, 8):
    if k <= n:
        print("D2H first %d gpus: %.1f GB/s total;  H2D %.1f;  evens-only(%s) %.1f" % (k, run(list(range(k)), hb), run(list(range(k)), hb, h2d=True), list(range(0, n, 2))[:k], run(list(range(0, n, 2))[:k], hb)))
del hb
for node in nodes:
    bufs = []
    rcs = []
    for _ in range(n):
        t, rc = pinned_on(node)
        bufs.append(t); rcs.append(rc)
    print("== pinned buffers bound to NUMA node %d (set_mempolicy rc %s)" % (node, set(rcs)))
    print("   per-gpu D2H:", " ".join("%.1f" % run([g], bufs, reps=4) for g in range(n)))
    print("   all %d gpus D2H: %.1f GB/s total" % (n, run(list(range(n)), bufs)))
    del bufs
if len(nodes) >= 2 and n >= 2:
    # each GPU's buffer on the node its PCIe device reports
    gn = []
    for i in range(n):
        bdf = torch.cuda.get_device_properties(i).pci_bus_id.lower() if hasattr(torch.cuda.get_device_properties(i), "pci_bus_id") else ""
        p = glob.glob("/sys/bus/pci/devices/*%s/numa_node" % bdf[-10:]) if bdf else []
        gn.append(int(open(p[0]).read()) if p else -1)
    print("== gpu -> numa node:", gn)
    bufs = [pinned_on(x if x >= 0 else None)[0] for x in gn]
    print("   local-node buffers, all gpus D2H: %.1f GB/s total" % run(list(range(n)), bufs))
