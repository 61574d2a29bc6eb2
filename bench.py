#!/usr/bin/env python
"""bench.py -- SPIKE factor+solve of the synthetic diagonally dominant band N=10M, K=100, fp64
(BASELINE.json metric / configs[2]) on N B200s of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  (N > 1: torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...)

One "step" = one in-place factorisation (bottom-up tip windows, band LU, spike tips, reduced
system) plus one solve of b = A*1 of the whole system, all ranks together (strong scaling: the same
10M-row system is row-block sharded over the ranks; only spike tips cross NVLink).  The band is
restored from a pristine device copy between steps, outside the timed brackets, because the
factorisation is in place.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ROWS = 10_000_000
K_HALF = 100
SEED, DELTA = 20140601, 1.2
METRIC = "spike_factor_plus_solve_ms_N10M_K100_fp64"


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0, interval_ms=20):
        self.rows, self.proc, self.index, self.interval_ms = [], None, index, interval_ms

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", str(self.interval_ms),
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(2)
        except Exception:
            pass
        mhz = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        # median under load: upper half of the samples (idle samples before/after the region drop out)
        load = mhz[len(mhz) // 2:] if mhz else []
        return {"sm_mhz": load[len(load) // 2] if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(mhz)}


def traffic_from_profiles():
    """dram bytes (read+write) per launch of the dominant kernel from the committed ncu capture."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    try:
        return json.load(open(p)).get("k_band_lu_dram_bytes_per_launch")
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------
def run_reference(args):
    """--impl reference: the reference CPU path restated (oracle): exact banded factor + solve as the
    reference's PCBANDED + `-banded_pc_type lu` computes it, run partition-parallel (OpenMP SPIKE port)
    on all host cores, on a bounded sample of the workload, scaled linearly in N."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the reference arm uses every host core
    if "TORCHELASTIC_RUN_ID" in os.environ or "LOCAL_RANK" in os.environ:
        os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    from oracle import oracle as O
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1)))
    except OSError:
        pass
    cores = O.num_threads()
    n_s = 1_000_000          # 1/10 of the rows; banded factor/solve cost is linear in N at fixed K
    scale = N_ROWS / n_s
    a = O.gen_band(n_s, K_HALF, SEED, DELTA)
    b = O.band_mult(a, np.ones(n_s))
    times = []
    for it in range(args.warmup + args.steps):
        S = O.Spike(n_s, K_HALF, max(cores * 4, 8), align=8, tip_rows=0)
        work = a.copy()
        t0 = time.perf_counter()
        S.factor(work, inplace=True)
        x = S.solve(b)
        dt = (time.perf_counter() - t0) * 1e3
        if it >= args.warmup:
            times.append(dt)
        err = float(np.abs(x - 1.0).max())
    ms = sum(times) / len(times) * scale
    line = {"impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "synthetic diagonally dominant band N=10M K=100 fp64, factor+solve (u=1, b=A*u)",
                       "seed": SEED, "delta": DELTA},
            "cpu_baseline": {"value": ms, "unit": "ms", "cores": cores, "kind": "port",
                             "sample": f"N={n_s} rows (1/{int(scale)} of the workload) of the K=100 band, CPU truncated-SPIKE port "
                                       f"with exact windows on {cores} threads, time scaled x{int(scale)} (cost linear in N); max|x-1|={err:.1e}"},
            "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample():
    import numpy as np
    from oracle import oracle as O
    cores = O.num_threads()
    n_s = 500_000
    scale = N_ROWS / n_s
    a = O.gen_band(n_s, K_HALF, SEED, DELTA)
    b = O.band_mult(a, np.ones(n_s))
    S = O.Spike(n_s, K_HALF, max(cores * 4, 8), align=8, tip_rows=0)
    t0 = time.perf_counter()
    S.factor(a, inplace=True)
    x = S.solve(b)
    ms = (time.perf_counter() - t0) * 1e3 * scale
    # serial exact band LU (the reference runs `-n 1`, src/makefile:18) on a smaller slice
    n1 = 100_000
    a1 = O.gen_band(n1, K_HALF, SEED, DELTA)
    b1 = O.band_mult(a1, np.ones(n1))
    t0 = time.perf_counter()
    lu, _ = O.band_lu(a1)
    O.band_solve(lu, b1)
    serial_ms = (time.perf_counter() - t0) * 1e3 * (N_ROWS / n1)
    return {"value": ms, "unit": "ms", "cores": cores, "kind": "port",
            "sample": f"N={n_s} rows (1/{int(scale)}) of the K=100 workload, OpenMP SPIKE port, scaled x{int(scale)}; "
                      f"max|x-1|={float(np.abs(x - 1).max()):.1e}; serial no-pivot band LU+solve on 1 core (N={n1}, scaled): {serial_ms:.0f} ms"}


# ---------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import spike_petsc_b200 as sp
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the SPIKE engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    bounds = sp.shard_rows(N_ROWS, world)
    n_loc = bounds[rank + 1] - bounds[rank]
    parts = args.partitions if args.partitions > 0 else 296   # 2 CTAs/SM per GPU (the LU kernel interleaves two partitions per SM)
    eng = sp.Spike(device=local, partitions=parts, tip_tiles=args.tip_tiles, mem=sp.MEM_DEVICE, rank=rank, nranks=world,
                   row_offset=bounds[rank], n_global=N_ROWS)
    eng.keep_original(True)                        # pristine copy: the factorisation is in place
    eng.set_band_synthetic(n_loc, K_HALF, SEED, DELTA)
    S = sp.ShardedSpike(eng, rank, world)
    u = torch.ones(n_loc, dtype=torch.float64, device=dev)
    b = torch.empty_like(u)
    x = torch.empty_like(u)
    S.mult(u, b)                                   # b = A*1 (src/testbed2.c:120-122), halos over NVLink
    # the band is restored from the kept original between steps, outside the timed brackets
    L = sp.lib()
    import ctypes as C
    L.spk_debug_restore_band.argtypes = [C.c_void_p]

    def restore():
        # Nothing to restore: with the unfactored band kept, spk_factor reads it (16 GB, far beyond L2) and writes the
        # factors into the working band, so every step is a complete factorisation of the same matrix with the same
        # 2B of traffic as an in-place one and no copy in between.  (SPIKE_B200_BENCH_RESTORE=1: old behaviour.)
        if os.environ.get("SPIKE_B200_BENCH_RESTORE", "0") == "1":
            rc = L.spk_debug_restore_band(eng._h)
            if rc:
                raise SystemExit(f"restore failed ({rc})")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local, args.clock_interval_ms)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    step_ms, lu_ms, stage = [], [], None
    for it in range(args.warmup + args.steps):
        restore()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        S.factor(u)
        S.solve(b, x)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        local_ms = ms.item()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if it >= args.warmup:
            step_ms.append(ms.item())
            stage = eng.view()["stage_ms"]
            lu_ms.append(stage[1])
    clocks = sampler.stop() if rank == 0 else None
    per_rank = None
    if world > 1:   # last step, every rank: its own event time, factor / solve split and stage times (where the ranks wait)
        iv = eng.view()
        mine = torch.tensor([local_ms, iv["factor_ms"], iv["solve_ms"]] + list(stage[:6]), dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [[round(v, 4) for v in t.tolist()] for t in allr]
    S.check()                                      # a timed-out mailbox spin would have left garbage in x
    exchange = "NVLink peer mailboxes (kernel stores + flags, csrc/peer.cu)" if getattr(S, "_peer", False) else "NCCL p2p"
    err = ((x - u).norm() ** 2)
    cnt = torch.tensor([float(n_loc)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(err); dist.all_reduce(cnt)
    relerr = (err.sqrt() / cnt.sqrt()).item()
    info = eng.view()

    # ---- end-to-end through the C ABI with HOST buffers (PCSetUp + PCApply as the glue calls them):
    #      pinned host band -> device (pack), factor, solve with host b / x.  Single GPU only.
    e2e = None
    if world == 1 and not args.no_e2e:
        import numpy as np
        if L.spk_debug_restore_band(eng._h):      # the working band holds factors: fetch the unfactored rows
            raise SystemExit("restore failed")
        torch.cuda.synchronize()
        rows = torch.empty((N_ROWS, 2 * K_HALF + 1), dtype=torch.float64).pin_memory()
        L.spk_get_band_rows(eng._h, rows.data_ptr())
        bh = b.cpu().pin_memory()
        xh = torch.empty_like(bh).pin_memory()
        eng.close()
        torch.cuda.empty_cache()
        times = []
        for it in range(2):
            t0 = time.perf_counter()
            h = sp.Spike(device=local, partitions=parts, tip_tiles=args.tip_tiles, mem=sp.MEM_HOST)
            L.spk_set_band_dense(h._h, N_ROWS, K_HALF, rows.data_ptr(), sp.LAYOUT_ROWS, sp.MEM_HOST)
            h.n, h.k = N_ROWS, K_HALF
            h.factor()
            L.spk_solve(h._h, bh.data_ptr(), xh.data_ptr(), 1)
            torch.cuda.synchronize()
            times.append((time.perf_counter() - t0) * 1e3)
            h.close()
        e2e_err = float((xh - 1.0).norm() / (N_ROWS ** 0.5))
        e2e = {"value": min(times), "unit": "ms", "h2d_bytes_per_step": int(rows.numel() * 8 + bh.numel() * 8),
               "d2h_bytes_per_step": int(xh.numel() * 8), "rel_err": e2e_err,
               "note": "spk_set_band_dense(host rows band, pinned) + spk_factor + spk_solve(host b -> host x); PCIe H2D of the 16 GB band dominates"}

    if rank == 0:
        ms = sum(step_ms) / len(step_ms)
        lu = sum(lu_ms) / len(lu_ms)
        peak, which = measured_peak()
        band_alg = 8.0 * N_ROWS * (2 * K_HALF + 1)
        lu_bytes = 2.0 * band_alg / world          # LU kernel: read + write every band entry of this rank's rows once
        achieved = lu_bytes / (lu * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "synthetic diagonally dominant band N=10M K=100 fp64, in-place SPIKE factor + solve of b=A*1",
                       "seed": SEED, "delta": DELTA, "partitions_per_gpu": info["partitions"], "tip_tiles": info["tip_tiles"],
                       "parallelism": f"row-block x{world}, spike-tip exchange over {exchange}" if world > 1 else "row-block x1",
                       "l2": "inputs (16 GB band) exceed the 126 MB L2; every step factors the kept unfactored band again (out of place: read original, write factors)"},
            "rel_err_vs_exact_u": relerr,
            "step_ms_all": [round(v, 4) for v in step_ms],
            "stage_ms": {"tip_windows": stage[0], "band_lu": stage[1], "spike_tips": stage[2], "sweeps": stage[3],
                         "reduced": stage[4], "corrections": stage[5]},
            "roofline": {"bound": "hbm", "kernel": "k_band_lu", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": which, "traffic": traffic_from_profiles() if world == 1 else None,
                         "algorithmic_bytes_per_launch": lu_bytes,
                         "whole_step_frac": (2 * band_alg + band_alg + 32.0 * N_ROWS) / world / (ms * 1e-3) / 1e9 / peak,
                         "fp64_tflops": N_ROWS * (2.0 * K_HALF * K_HALF + K_HALF) / world / (lu * 1e-3) / 1e12},
            "gpu_launches": info["kernel_launches"],
            "per_rank_ms": {"columns": ["step", "factor", "solve", "tip_windows", "band_lu", "spike_tips", "sweeps", "reduced", "corrections"],
                            "rows": per_rank} if per_rank else None,
            "clocks": clocks,
            "e2e": e2e if e2e is not None else {"value": None, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                                                "note": "host-buffer path measured at N=1 only"},
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_sample()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--partitions", type=int, default=0)
    ap.add_argument("--tip-tiles", type=int, default=78)   # 6 bandwidths: 1e-13 (profiles/r01_truncation_window.md)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--clock-interval-ms", type=int, default=20)   # nvidia-smi sampling period during the timed region
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
