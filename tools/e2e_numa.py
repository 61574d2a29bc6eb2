"""Host-buffer path: NUMA placement of the pinned staging memory and the time of every piece of one end-to-end call.
usage: e2e_numa.py [N]"""
import os, sys, time, ctypes as C
sys.path.insert(0, '.')
import torch
import spike_petsc_b200 as sp

def gpu_numa(index=0):
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(':')[0]) == 8:
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        cpus = open(f"/sys/devices/system/node/node{max(node,0)}/cpulist").read().strip()
        return bus, node, cpus
    except Exception as e:  # noqa: BLE001
        return None, -1, str(e)

def parse_cpulist(s):
    out = set()
    for part in s.split(','):
        if '-' in part:
            a, b = part.split('-'); out.update(range(int(a), int(b) + 1))
        elif part:
            out.add(int(part))
    return out

N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
K = 100
print("nodes online:", open("/sys/devices/system/node/online").read().strip(), "| affinity:", sorted(os.sched_getaffinity(0)))
bus, node, cpus = gpu_numa(0)
print("GPU 0 bus", bus, "numa node", node, "cpus", cpus)
orig = os.sched_getaffinity(0)

def h2d_bw(tag):
    h = torch.empty(1 << 28, dtype=torch.float64).pin_memory()   # 2 GB
    h.fill_(1.0)
    d = torch.empty_like(h, device='cuda')
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(4):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{tag}: H2D {4 * h.numel() * 8 / dt / 1e9:.1f} GB/s", flush=True)
    del h, d

h2d_bw("default placement")
if node >= 0:
    try:
        os.sched_setaffinity(0, parse_cpulist(cpus) & orig or orig)
        print("bound to", sorted(os.sched_getaffinity(0)))
        h2d_bw("allocated from the GPU's node")
    except Exception as e:  # noqa: BLE001
        print("cannot bind:", e)

L = sp.lib()
L.spk_get_band_rows.argtypes = [C.c_void_p, C.c_void_p]
g = sp.Spike(partitions=296, tip_tiles=78, mem=sp.MEM_DEVICE)
g.set_band_synthetic(N, K)
rows = torch.empty((N, 2 * K + 1), dtype=torch.float64).pin_memory()
L.spk_get_band_rows(g._h, rows.data_ptr())
u = torch.ones(N, dtype=torch.float64, device='cuda'); b = torch.empty_like(u)
g.mult(u.data_ptr(), b.data_ptr()); torch.cuda.synchronize()
bh = b.cpu().pin_memory(); xh = torch.empty_like(bh).pin_memory()
g.close(); del u, b; torch.cuda.empty_cache()
for it in range(3):
    t0 = time.perf_counter()
    h = sp.Spike(partitions=296, tip_tiles=78, mem=sp.MEM_HOST)
    t1 = time.perf_counter()
    L.spk_set_band_dense(h._h, N, K, rows.data_ptr(), sp.LAYOUT_ROWS, sp.MEM_HOST)
    t2 = time.perf_counter()
    h.n, h.k = N, K
    h.factor(); torch.cuda.synchronize()
    t3 = time.perf_counter()
    L.spk_solve(h._h, bh.data_ptr(), xh.data_ptr(), 1); torch.cuda.synchronize()
    t4 = time.perf_counter()
    h.close()
    t5 = time.perf_counter()
    print(f"create {1e3*(t1-t0):.1f} | set_band_dense {1e3*(t2-t1):.1f} ({N*(2*K+1)*8/(t2-t1)/1e9:.1f} GB/s) | factor {1e3*(t3-t2):.1f} | solve(host) {1e3*(t4-t3):.1f} | close {1e3*(t5-t4):.1f} | total {1e3*(t4-t0):.1f} ms  err {float((xh-1).norm()/N**0.5):.1e}", flush=True)
