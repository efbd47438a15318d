#!/usr/bin/env python
"""bench.py — RMI build keys/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's CUDA path
  python bench.py --impl reference --steps 3 --warmup 1     # the reference's CPU algorithm (oracle port)
  torchrun ... bench.py --gpus N ...                        # one rank per GPU

A "step" is one complete two-layer build (rmi_lib::train) of the workload
`linear,linear 1048576 on 200M synthetic uint64` (BASELINE.json configs[1]) per GPU:
top-model fit, leaf boundaries, per-leaf fits, forward/error pass, lower-bound widening,
statistics, and the copy of all leaf parameters and error bounds back to the host.

value       keys/s with the key array already resident in HBM (all ranks' keys / max-over-ranks time)
e2e         the same build through the C ABI from a PINNED HOST buffer: H2D copy of the keys +
            build + results on the host, every step
roofline    the dominant kernel (fused leaf fit + forward/error pass): algorithmic bytes per
            launch / its CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth
cpu_baseline the CPU oracle (a C++ port of the reference algorithm; the Rust reference cannot be
            built offline) timed on a bounded sample of the same workload on this box's cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "RMI build keys/sec (200M uint64, linear,linear 2^20)"
HBM_FALLBACK_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--keys", type=float, default=200e6, help="keys per GPU")
    ap.add_argument("--leaves", type=int, default=1 << 20)
    ap.add_argument("--spec", default="linear,linear")
    ap.add_argument("--exact-top", action="store_true", help="RMI_FLAG_TOP_FIT_EXACT (serial top fit)")
    ap.add_argument("--cpu-sample-div", type=int, default=8, help="cpu baseline runs on n/div keys, N/div leaves")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed regions run."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch():
    """dram read+write bytes of the dominant kernel from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


def sample_for_cpu(keys_np, div):
    return keys_np[::div].copy()


def time_oracle(keys_np, spec, leaves, threads=2):
    import oracle
    t0 = time.perf_counter()
    r = oracle.train(keys_np, spec, leaves, threads=threads)
    dt = time.perf_counter() - t0
    r.close()
    return dt


def time_oracle_concurrent(keys_np, spec, leaves, workers):
    """`workers` independent builds at once, 2 threads each — how the reference occupies a many-core host
    (optimizer.rs:224 par_iter over configurations); returns aggregate keys/s.  ctypes releases the GIL."""
    from concurrent.futures import ThreadPoolExecutor
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=workers) as ex:
        list(ex.map(lambda _: time_oracle(keys_np, spec, leaves), range(workers)))
    return workers * keys_np.size / (time.perf_counter() - t0)


def host_keys_numpy(n, seed):
    """Sorted uniform uint64 keys (< 2^63) on the host, without needing a GPU."""
    import numpy as np
    rng = np.random.Generator(np.random.MT19937(seed))
    k = rng.integers(0, (1 << 63) - 1, size=n, dtype=np.int64).astype(np.uint64)
    k.sort()
    return k


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU algorithm (oracle port, <= 2 threads as
    rayon::join gives the reference, two_layer.rs:161-169) on a bounded sample."""
    if rank != 0:
        return
    import oracle
    oracle.build()
    n_full = int(args.keys)
    n_s = max(n_full // args.cpu_sample_div, 1000)
    leaves_s = max(args.leaves // args.cpu_sample_div, 2)
    keys = host_keys_numpy(n_s, 42)
    for _ in range(args.warmup):
        time_oracle(keys, args.spec, leaves_s)
    t = []
    for _ in range(args.steps):
        t.append(time_oracle(keys, args.spec, leaves_s))
    tot = sum(t)
    val = n_s * args.steps / tot
    sample = f"{n_s} uniform uint64 keys, {args.spec} {leaves_s} (1/{args.cpu_sample_div} of the workload, same keys per leaf)"
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "keys/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{args.spec} {args.leaves} on {n_full} synthetic sorted uint64 per GPU",
                      "note": "Rust reference cannot be built offline (no cargo); this is the C++ oracle port of "
                              "rmi_lib::train, single build uses <= 2 threads like the reference"},
           "cpu_baseline": {"value": val, "unit": "keys/s", "cores": 2, "kind": "port", "sample": sample,
                            "host_cores": os.cpu_count()},
           "e2e": {"value": val, "unit": "keys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    emit(out)


def emit(obj):
    """The ONE JSON line of the contract, on the process's real stdout."""
    line = (json.dumps(obj) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    os.write(fd, line)


_REAL_STDOUT = None


def quiet_stdout():
    """Libraries print banners to stdout (NCCL's version line, torchrun notes): send everything
    except the final JSON line to stderr so that stdout carries exactly one line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def main():
    quiet_stdout()
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import rmi_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    rmi_b200.load_library()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = int(args.keys)
    N = args.leaves
    flags = rmi_b200.FLAG_TOP_FIT_EXACT if args.exact_top else 0
    # ---- synthetic workload: sorted uniform uint64 keys; rank r draws from the r-th slice of
    # the key space so that the concatenation over ranks is globally sorted -------------------
    g = torch.Generator(device=dev)
    g.manual_seed(42 + rank)
    width = ((1 << 63) - 1) // world
    k = torch.randint(0, width, (n,), dtype=torch.int64, device=dev, generator=g) + rank * width
    k, _ = torch.sort(k)
    torch.cuda.synchronize()
    ppm = 2
    key_bytes = 8
    if world == 1:
        ds = rmi_b200.RMITrainingData.from_device(k.data_ptr(), n, rmi_b200.KEY_U64, local_rank, keep_alive=k)

        def build():
            return rmi_b200.train(ds, args.spec, N, flags, counts=False)
    else:
        # range-partitioned build: ONE global RMI with N leaves over all ranks' keys
        # (rank r holds the r-th slab of the globally sorted array); weak scaling in keys.
        from rmi_b200 import sharded
        sdata = sharded.ShardedTrainingData(k, n, rmi_b200.KEY_U64, halo_capacity=1 << 20)

        def build():
            return sharded.train_sharded(sdata, args.spec, N, flags, counts=False)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    # ---- warm-up --------------------------------------------------------------------------
    res = None
    for _ in range(max(args.warmup, 3)):
        res = build()
    if rank == 0:
        clocks.start()
    # ---- timed region: K resident builds ---------------------------------------------------
    launches0 = rmi_b200.kernel_launch_count()
    phase = np.zeros(4)
    dev_ns = 0.0
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res = build()      # synchronous: returns with results on the host
        phase += np.array(res.phase_device_ns, dtype=np.float64)
        dev_ns += res.device_time_ns
    e1.record()
    barrier()
    wall = time.perf_counter() - t0
    ev_ms = e0.elapsed_time(e1)
    launches = rmi_b200.kernel_launch_count() - launches0
    t_rank = torch.tensor([ev_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_rank, op=dist.ReduceOp.MAX)
    ms_total = float(t_rank.item())
    ms_per_step = ms_total / args.steps
    value = (n * world) / (ms_per_step / 1e3)

    # ---- e2e: pinned host keys -> H2D -> build -> results on host, every step ----------------
    host = torch.empty(n, dtype=torch.int64, pin_memory=True)
    host.copy_(k)
    torch.cuda.synchronize()
    host_np = host.numpy().view(np.uint64)
    def e2e_step():
        if world == 1:
            d2 = rmi_b200.RMITrainingData(host_np, device=local_rank)   # cudaMemcpy H2D from pinned memory
            r2 = rmi_b200.train(d2, args.spec, N, flags, counts=False)
            d2.close()
        else:
            from rmi_b200 import sharded
            kd = torch.empty(n + (1 << 20), dtype=torch.int64, device=dev)
            kd[:n].copy_(host, non_blocking=False)                       # H2D from pinned memory
            sd = sharded.ShardedTrainingData(kd, n, rmi_b200.KEY_U64, halo_capacity=1 << 20)
            r2 = sharded.train_sharded(sd, args.spec, N, flags, counts=False)
        return r2

    e2e_step()
    barrier()
    e2 = torch.cuda.Event(enable_timing=True)
    e3 = torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.e2e_steps):
        e2e_step()
    e3.record()
    barrier()
    t_e2e = torch.tensor([e2.elapsed_time(e3)], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_ms = float(t_e2e.item()) / args.e2e_steps
    e2e_val = (n * world) / (e2e_ms / 1e3)
    clk = clocks.stop() if rank == 0 else None

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------
    phase_ms = phase / args.steps / 1e6
    names = ["top_fit", "leaf_bounds", "leaf_fit_error(k_leaf)", "statistics"]
    dom = int(np.argmax(phase_ms))
    peak, peak_src = measured_hbm_peak()
    out_bytes = N * (8 * ppm + 8)     # leaf parameters + error bounds copied to the host every step
    kern_bytes = {0: n * key_bytes, 1: n * key_bytes + (N + 1) * 8, 2: n * key_bytes + (N + 1) * 8 + out_bytes,
                  3: N * 16}[dom]
    achieved = kern_bytes / (phase_ms[dom] / 1e3) / 1e9
    build_bytes = 2 * n * key_bytes + N * (8 * ppm + 8)
    tr = ncu_traffic_per_launch()
    roofline = {"bound": "hbm", "kernel": names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": (tr or {}).get("dram_bytes_per_launch"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": kern_bytes,
                "kernel_ms": float(phase_ms[dom]),
                "phases_ms": {nm: float(v) for nm, v in zip(names, phase_ms)},
                "whole_build": {"algorithmic_bytes": build_bytes, "device_ms": dev_ns / args.steps / 1e6,
                                "achieved": build_bytes / (dev_ns / args.steps / 1e9) / 1e9,
                                "frac": build_bytes / (dev_ns / args.steps / 1e9) / 1e9 / peak}}

    # ---- CPU baseline on a bounded sample ------------------------------------------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        import oracle
        oracle.build()
        ks = sample_for_cpu(host_np, args.cpu_sample_div)
        leaves_s = max(N // args.cpu_sample_div, 2)
        dt = min(time_oracle(ks, args.spec, leaves_s) for _ in range(2))
        cpu = {"value": ks.size / dt, "unit": "keys/s", "cores": 2, "kind": "port",
               "sample": f"every {args.cpu_sample_div}th key: {ks.size} keys, {args.spec} {leaves_s} "
                         f"(same keys per leaf), best of 2", "host_cores": os.cpu_count(), "seconds": dt}
        # one build cannot use more than 2 threads (two_layer.rs:161-169); a many-core host is only filled by
        # independent builds, as in the reference's --optimize sweep: aggregate throughput of W such builds
        workers = max(1, min((os.cpu_count() or 2) // 2, 32))
        cpu["many_builds_at_once"] = {"value": time_oracle_concurrent(ks, args.spec, leaves_s, workers), "unit": "keys/s",
                                      "builds": workers, "cores": 2 * workers}

    out = {"metric": METRIC, "value": value, "unit": "keys/s", "n_gpus": world, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": f"{args.spec} {N} on {n} synthetic sorted uint64 per GPU",
                      "top_fit": "exact-serial" if args.exact_top else "parallel (coefficients within 1e-9 of the reference)",
                      "keys_per_gpu": n, "leaves": N, "key_type": "uint64",
                      "l2": "inputs (1.6 GB) larger than L2, no flush needed",
                      "parallelism": "1 GPU" if world == 1 else
                      f"range-partitioned over {world} GPUs: one global RMI with {N} leaves over {n * world} keys; "
                      "per build: all-reduce of the top-model sums (64 B), of the leaf boundaries ((N+1)*8 B) and of the leaf "
                      "records (N*24 B); halo keys between neighbours are fetched once per data set",
                      "timing": "CUDA events around K synchronous builds, max over ranks",
                      "roofline_note": "dominant kernel = the fused leaf fit + forward pass (k_leaf); on one GPU it runs as 4 launch "
                                       "slices per build (their results cross PCIe while the next slice computes): achieved = "
                                       "algorithmic bytes of all slices / the leaf phase's device time, traffic = ncu DRAM bytes "
                                       "summed over the slices of one build (profiles/dominant_kernel_traffic.json)",
                      "wall_ms_per_step": 1e3 * wall / args.steps},
           "clocks": clk,
           "e2e": {"value": e2e_val, "unit": "keys/s", "h2d_bytes_per_step": n * key_bytes,
                   "d2h_bytes_per_step": out_bytes, "ms_per_step": e2e_ms},
           "gpu_launches": int(launches),
           "roofline": roofline, "cpu_baseline": cpu}
    emit(out)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
