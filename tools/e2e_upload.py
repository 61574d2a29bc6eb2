"""Host-buffer path timing: spk_set_band_dense(host rows) + factor + solve, for several upload chunk sizes.
usage: e2e_upload.py [chunk_rows ...]   (0 = default 256 MB chunks)"""
import sys, time, ctypes as C
sys.path.insert(0, '.')
import torch
import spike_petsc_b200 as sp
N, K = 10_000_000, 100
L = sp.lib()
L.spk_debug_set_pack_chunk_rows.argtypes = [C.c_int64]
g = sp.Spike(partitions=296, tip_tiles=78, mem=sp.MEM_DEVICE)
g.set_band_synthetic(N, K)
rows = torch.empty((N, 2 * K + 1), dtype=torch.float64).pin_memory()
L.spk_get_band_rows.argtypes = [C.c_void_p, C.c_void_p]
L.spk_get_band_rows(g._h, rows.data_ptr())
u = torch.ones(N, dtype=torch.float64, device='cuda'); b = torch.empty_like(u)
g.mult(u.data_ptr(), b.data_ptr()); torch.cuda.synchronize()
bh = b.cpu().pin_memory(); xh = torch.empty_like(bh).pin_memory()
g.close(); del u, b; torch.cuda.empty_cache()
for chunk in [int(v) for v in sys.argv[1:]] or [0]:
    L.spk_debug_set_pack_chunk_rows(chunk)
    ts = []
    for it in range(3):
        t0 = time.perf_counter()
        h = sp.Spike(partitions=296, tip_tiles=78, mem=sp.MEM_HOST)
        L.spk_set_band_dense(h._h, N, K, rows.data_ptr(), sp.LAYOUT_ROWS, sp.MEM_HOST)
        t1 = time.perf_counter()
        h.n, h.k = N, K
        h.factor()
        L.spk_solve(h._h, bh.data_ptr(), xh.data_ptr(), 1)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        h.close()
        ts.append(((t2 - t0) * 1e3, (t1 - t0) * 1e3))
    print(f"chunk_rows={chunk}: total/upload ms {[(round(a,1), round(b,1)) for a,b in ts]} err {float((xh-1).norm()/N**0.5):.1e}", flush=True)
