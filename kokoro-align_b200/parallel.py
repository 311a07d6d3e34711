"""Multi-GPU host logic: independent lattices (chapters / segments) are partitioned across the
ranks of one node, one process per GPU, NO collective on the data path -- the reference's
outer loop over files (run_example.py:247-254) is embarrassingly parallel.  Only the result
gather and the timing reduction use torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
import numpy as np


def cells_eval(T, L, beam_size=1000):
    """Evaluated cells sum_i (hi_i - lo_i) of one lattice (align.py:64-65), vectorised."""
    T, L = int(T), int(L)
    S = 2 * L + 1
    if beam_size >= S and (S * (T - 1)) // T <= beam_size // 2:
        return T * S
    i = np.arange(T, dtype=np.int64)
    lo = np.maximum(0, S * i // T - beam_size // 2)
    hi = np.maximum(np.minimum(lo + beam_size, S), lo)
    return int((hi - lo).sum())


def lpt_partition(costs, n_parts):
    """Longest-processing-time-first bin packing.  Returns n_parts sorted index arrays that
    cover range(len(costs)) exactly once; the heaviest part is within max(costs) of the mean."""
    costs = np.asarray(costs, dtype=np.float64)
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(n_parts)
    parts = [[] for _ in range(n_parts)]
    for idx in order:
        k = int(np.argmin(load))
        parts[k].append(int(idx))
        load[k] += costs[idx]
    return [np.array(sorted(p), dtype=np.int64) for p in parts]


def shard_batch(T_list, L_list, rank, world, beam_size=1000):
    """Indices of the lattices rank `rank` of `world` aligns (same answer on every rank)."""
    costs = [cells_eval(t, l, beam_size) for t, l in zip(T_list, L_list)]
    return lpt_partition(costs, world)[rank]


def flat_batch(log_probs_list, labels_list):
    """Pack per-lattice arrays into the C-ABI flat batch layout."""
    t_off = np.concatenate([[0], np.cumsum([len(x) for x in log_probs_list])]).astype(np.int64)
    l_off = np.concatenate([[0], np.cumsum([len(x) for x in labels_list])]).astype(np.int64)
    V = log_probs_list[0].shape[1] if log_probs_list else 1
    lp = np.concatenate(log_probs_list).astype(np.float32, copy=False) if log_probs_list else np.zeros((0, V), np.float32)
    labels = (np.concatenate([np.asarray(x).astype(np.int32) for x in labels_list])
              if labels_list else np.zeros(0, np.int32))
    return np.ascontiguousarray(lp), t_off, labels, l_off


def _default_align(lp, t_off, labels, l_off, V, beam_size, max_move, device):
    from . import align
    with align.AlignPlan(t_off, labels, l_off, V, beam_size, max_move, device=device) as plan:
        return plan.run_host(lp)


def align_sharded(log_probs_list, labels_list, beam_size=1000, max_move=4, group=None, device=0,
                  align_fn=None, dst=0):
    """Align a list of lattices across the ranks of `group`.

    Every rank passes the same lists (or at least the same shapes), aligns its LPT shard with
    `align_fn` (default: the CUDA AlignPlan on `device`) and the per-lattice results
    (best_path, best_labels, best_scores, final_score, status) are gathered on rank `dst` in
    the original order.  Other ranks return None.
    """
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    T_list = [len(x) for x in log_probs_list]
    L_list = [len(x) for x in labels_list]
    mine = shard_batch(T_list, L_list, rank, world, beam_size)
    V = log_probs_list[0].shape[1]
    lp, t_off, labels, l_off = flat_batch([log_probs_list[i] for i in mine], [labels_list[i] for i in mine])
    fn = align_fn or _default_align
    if len(mine):
        path, labs, scores, final, status = fn(lp, t_off, labels, l_off, V, beam_size, max_move, device)
    local = {}
    for n, i in enumerate(mine):
        a, b = int(t_off[n]), int(t_off[n + 1])
        local[int(i)] = (path[a:b].copy(), labs[a:b].copy(), scores[a:b].copy(), float(final[n]), int(status[n]))
    if world == 1:
        return [local[i] for i in range(len(T_list))]
    gathered = [None] * world if rank == dst else None
    dist.gather_object(local, gathered, dst=dst, group=group)
    if rank != dst:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    assert len(merged) == len(T_list), "a lattice was aligned twice or not at all"
    return [merged[i] for i in range(len(T_list))]


def max_over_ranks(x, group=None, device=None):
    """max of a python float over ranks (the multi-GPU timing rule of bench.py)."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def sum_over_ranks(x, group=None, device=None):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(x)
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())


def bind_host_to_gpu(device, world=1):
    """Pin this process to the CPU cores NVML reports as local to GPU `device` (same NUMA node /
    PCIe root), BEFORE its pinned staging buffers are allocated: first-touch then places them in
    the memory next to the GPU, so the H2D streams of the ranks of one box do not all cross the
    same socket link.  When NVML gives every GPU the same mask (one NUMA node, or a container
    that hides the topology) the `world` ranks would all sit on the same cores: the mask is then
    cut into `world` contiguous slices and rank `device` takes its own.  Returns the cpu set, or
    None when NVML / affinity is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            props_uuid = None
            if torch.cuda.is_available():
                props_uuid = str(torch.cuda.get_device_properties(device).uuid)
            handle = None
            if props_uuid:   # CUDA_VISIBLE_DEVICES may renumber the devices: match by UUID
                for k in range(pynvml.nvmlDeviceGetCount()):
                    h = pynvml.nvmlDeviceGetHandleByIndex(k)
                    u = pynvml.nvmlDeviceGetUUID(h)
                    u = u.decode() if isinstance(u, bytes) else u
                    if props_uuid in u or u.replace("GPU-", "") == props_uuid:
                        handle = h
                        break
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(device)
            n_words = (os.cpu_count() + 63) // 64
            words = pynvml.nvmlDeviceGetCpuAffinity(handle, n_words)
        finally:
            pynvml.nvmlShutdown()
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        if world > 1 and len(cpus) >= 2 * world:
            # do the other GPUs report the same mask?  then share it out instead of piling up
            same = True
            try:
                pynvml.nvmlInit()
                try:
                    for k in range(min(world, pynvml.nvmlDeviceGetCount())):
                        wk = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(k), n_words)
                        if list(wk) != list(words):
                            same = False
                finally:
                    pynvml.nvmlShutdown()
            except Exception:  # noqa: BLE001
                same = False
            if same:
                order = sorted(cpus)
                per = len(order) // world
                cpus = set(order[(device % world) * per:(device % world + 1) * per])
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:  # noqa: BLE001  (no NVML, no permission, non-Linux: run unbound)
        return None
