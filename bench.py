#!/usr/bin/env python
"""bench.py -- lattice cells/s of the CTC best-path hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload segments|chapters|gon]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's CPU path (numpy port), host cores

A "step" is one pass of the hot path over one batch of synthetic lattices.  The default
workload is BASELINE config 2 (10 000 silence-split segments of 1-10 s: T_b ~ U{86..861},
L_b = round(0.14 T_b), V = 39, beam_size 1000, max_move 4) -- for these shapes
the window never clips, so evaluated cells == nominal T*S cells.

value  : cells/s with log_probs already resident in HBM (CUDA events on the launch stream).
e2e    : same metric through the host-buffer C-ABI call (pinned host memory, H2D of the
         log-probs and D2H of the three output arrays inside the timed region).
Multi-GPU: one process per GPU, every rank aligns its own batch (no collective on the data
path; weak scaling); time = max over ranks; value = cells of all ranks / that time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from kokoro_align_b200 import synth  # noqa: E402

METRIC = "lattice cells/sec (TxS)"
UNIT = "cells/s"


def band_kernel_name(n_band, beam_size=1000):
    """Which band kernel kab_plan_create picks (kab_api.cu): the pipelined cluster kernel when every
    band lattice of the plan gets its own cluster of ceil(ceil((W + 32) / 104) / 4) CTAs at once."""
    import torch
    nc = -(-(-(-(beam_size + 32) // 104)) // 4)
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    forced = os.environ.get("KAB_BAND_CLUSTER")
    cluster = n_band <= sms // nc if forced is None else int(forced) >= 1
    return "kab_bandp_kernel" if cluster else "kab_band_kernel"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_shapes(name, seed, n_lattices=None):
    """(T[], L[], description) of a workload; lattice n uses its own RNG stream."""
    if name == "segments":      # BASELINE config 2
        B = n_lattices or 10000
        T, L = synth.segment_lengths(B, seed)
        return T, L, f"config2: {B} silence-split segments, T~U{{86..861}}, L=round(0.14T), V=39, W=1000, M=4"
    if name == "chapters":      # BASELINE config 3
        C = n_lattices or 36
        T, L = synth.chapter_lengths(C, 2721800, seed)
        return T, L, f"config3: Kokoro book, {C} chapter lattices, sum T=2721800, V=39, W=1000, M=4"
    if name == "gon":           # BASELINE config 1
        return (np.array([81135]), np.array([11359]),
                "config1: Gon gitsune single lattice T=81135 L=11359 V=39 W=1000 M=4")
    raise SystemExit(f"unknown workload {name}")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 3 + k and r[3 + k] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def numpy_port_rate(lp, t_off, labels, l_off, budget_s, picks):
    """Time the numpy port (reference cost structure) on lattices `picks` until budget_s."""
    from oracle import ctc_oracle_np
    cells, t0, used = 0, time.perf_counter(), 0
    for b in picks:
        a, e = int(t_off[b]), int(t_off[b + 1])
        la, le = int(l_off[b]), int(l_off[b + 1])
        ctc_oracle_np.ctc_best_path(lp[a:e], labels[la:le])
        cells += ctc_oracle_np.cells_eval(e - a, le - la)
        used += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return cells / dt, used, dt


def _pool_job(args):
    from oracle import ctc_oracle_np
    lp, labels = args
    ctc_oracle_np.ctc_best_path(lp, labels)
    return ctc_oracle_np.cells_eval(lp.shape[0], labels.shape[0])


def run_reference(args):
    """--impl reference: the reference's CPU path (numpy port: the reference is pure Python, so
    there is no compiled oracle/_ref) on all host cores, bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = os.cpu_count() or 1
    T, L, desc = workload_shapes(args.workload, args.seed)
    per_step = max(cores * 32, 64) if args.workload == "segments" else min(len(T), cores)
    n = min(len(T), per_step)
    order = np.random.default_rng(args.seed + 1).permutation(len(T))[:n]
    if args.workload != "segments":       # bound long lattices to ~20k frames each
        T, L = T.copy(), L.copy()
        for b in order:
            if T[b] > 20000:
                L[b] = int(round(L[b] * 20000 / T[b])); T[b] = 20000
    jobs = [synth.make_lattice(int(T[b]), int(L[b]), 39, args.seed + 10 + int(b)) for b in order]
    with mp.get_context("fork").Pool(min(cores, n)) as pool:
        for _ in range(args.warmup):
            pool.map(_pool_job, jobs, chunksize=1)
        t0 = time.perf_counter()
        cells = 0
        for _ in range(args.steps):
            cells += sum(pool.map(_pool_job, jobs, chunksize=1))
        dt = time.perf_counter() - t0
    value = cells / dt
    sample = f"{n} lattices of the workload per step (sum T={int(T[order].sum())}), numpy port, Pool({min(cores, n)})"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": desc},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": min(cores, n), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="segments", choices=["segments", "chapters", "gon"])
    ap.add_argument("--lattices", type=int, default=None)
    ap.add_argument("--seed", type=int, default=2000)
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    # libraries (NCCL prints its version) write to fd 1: keep the real stdout for the one JSON line
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from kokoro_align_b200 import align, parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    cpus = parallel.bind_host_to_gpu(local_rank) if world > 1 else None   # NUMA-local pinned buffers
    # ---- this rank's batch (weak scaling: every rank its own 10k segments / book)
    T, L, desc = workload_shapes(args.workload, args.seed + 7919 * rank, args.lattices)
    lp, t_off, labels, l_off = synth.make_batch_fast(T, L, seed=args.seed + 1 + 7919 * rank)
    plan = align.AlignPlan(t_off, labels, l_off, 39, device=local_rank)
    info = plan.info
    cells_eval, cells_nominal = int(info.cells_eval), int(info.cells_nominal)
    n_frames = int(t_off[-1])

    # pinned host buffers for the e2e path
    h_lp = torch.from_numpy(lp).pin_memory()
    h_out = (torch.empty(n_frames, dtype=torch.int32).pin_memory(), torch.empty(n_frames, dtype=torch.int32).pin_memory(),
             torch.empty(n_frames, dtype=torch.float32).pin_memory(), torch.empty(plan.B, dtype=torch.float32).pin_memory(),
             torch.empty(plan.B, dtype=torch.int32).pin_memory())
    h_out_np = tuple(t.numpy() for t in h_out)
    d_lp = h_lp.to(dev, non_blocking=True)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def sum_over_ranks(x):
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return float(x)

    # ---- device-resident: value + roofline (CUDA events on the launching stream)
    outs = None
    for _ in range(args.warmup):
        outs = plan.run_torch(d_lp)
    torch.cuda.synchronize()
    status = outs[4].cpu().numpy()
    assert (status == 0).all(), f"non-zero lattice status in bench: {np.unique(status)}"
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for k in range(args.steps):
        outs = plan.run_torch(d_lp)
        ev[k + 1].record()
    barrier()
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = max_over_ranks(ev[0].elapsed_time(ev[args.steps]))
    cells_all = sum_over_ranks(cells_eval)
    nominal_all = sum_over_ranks(cells_nominal)
    value = cells_all * args.steps / (total_ms * 1e-3)
    kernel_ms = float(np.mean(step_ms))

    # ---- e2e: host buffers through the C-ABI (H2D + kernels + D2H per step)
    for _ in range(2):
        plan.run_host(h_lp.numpy(), out=h_out_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.run_host(h_lp.numpy(), out=h_out_np)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = cells_all * args.steps / e2e_s
    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions
    np.testing.assert_array_equal(h_out_np[0], outs[0].cpu().numpy())

    # ---- secondary: the same call on RAW LOGITS, align.py:116-117 on the device (kab_softmax.cuh)
    for _ in range(2):
        plan.run_host(h_lp.numpy(), out=h_out_np, logits=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        plan.run_host(h_lp.numpy(), out=h_out_np, logits=True)
    barrier()
    e2e_logits_s = max_over_ranks(time.perf_counter() - t0)
    sm_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    d_tmp = torch.empty_like(d_lp)
    align.log_softmax_torch(d_lp, out=d_tmp)
    sm_ev[0].record()
    for _ in range(args.steps):
        align.log_softmax_torch(d_lp, out=d_tmp)
    sm_ev[1].record()
    torch.cuda.synchronize()
    softmax_ms = sm_ev[0].elapsed_time(sm_ev[1]) / args.steps
    del d_tmp
    plan.run_host(h_lp.numpy(), out=h_out_np)   # restore the log-prob results for the parity spot check

    # ---- CPU baseline (rank 0, bounded sample of the same workload, numpy port, 1 core)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        picks = np.random.default_rng(args.seed + 1).permutation(plan.B)
        rate, used, dt = numpy_port_rate(lp, t_off, labels, l_off, args.cpu_seconds, picks)
        cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{used} randomly chosen lattices of the workload, {dt:.1f} s, numpy port of align.py:43-109 (oracle/ctc_oracle_np.py)"}
        # parity spot check of the bench inputs against the C oracle
        from oracle import ctc_oracle
        for b in picks[:16]:
            a, e = int(t_off[b]), int(t_off[b + 1])
            rp = ctc_oracle.ctc_best_path(lp[a:e], labels[int(l_off[b]):int(l_off[b + 1])])[0]
            np.testing.assert_array_equal(h_out_np[0][a:e], rp)

    if rank == 0:
        peak, peak_src = load_peaks()
        alg_bytes = int(info.algorithmic_bytes)
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        prof = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        traffic = None
        if os.path.exists(prof):
            with open(prof) as f:
                traffic = json.load(f).get(args.workload)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "lattices_per_gpu": plan.B, "frames_per_gpu": n_frames,
                       "cells_eval_per_gpu": cells_eval, "cells_nominal_per_gpu": cells_nominal,
                       "cells_nominal_per_s": nominal_all * args.steps / (total_ms * 1e-3),
                       "kernel_classes": {"warp": int(info.n_class[0]), "band": int(info.n_class[1]),
                                          "generic": int(info.n_class[2])},
                       "l2": f"inputs ({n_frames * 39 * 4 / 1e6:.0f} MB log-probs + {int(info.backptr_bytes) / 1e6:.0f} MB backpointers per step) exceed the 126 MB L2",
                       "parallelism": f"{world} independent ranks, no collective on the data path",
                       "host_binding": (f"rank 0 bound to {len(cpus)} GPU-local cores (NVML affinity)" if cpus else "none")},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n_frames * 39 * 4,
                    "d2h_bytes_per_step": n_frames * 12 + plan.B * 8, "ms_per_step": e2e_s / args.steps * 1e3},
            "e2e_raw_logits": {"value": cells_all * args.steps / e2e_logits_s, "unit": UNIT,
                               "ms_per_step": e2e_logits_s / args.steps * 1e3,
                               "note": "same call on raw logits, log-softmax (align.py:116-117) on the device",
                               "log_softmax_kernel_ms": softmax_ms,
                               "log_softmax_gbs": 8.0 * n_frames * 39 / (softmax_ms * 1e-3) / 1e9},
            "gpu_launches": int(info.kernel_launches) * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "kab_warp_kernel" if info.n_class[0] >= info.n_class[1] else band_kernel_name(int(info.n_class[1])),
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms},
            "cpu_baseline": cpu, "clocks": clocks,
        }
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    plan.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
