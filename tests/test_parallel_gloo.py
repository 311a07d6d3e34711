"""N > 1 host path on CPU: world_size-2 gloo.  The CUDA aligner is replaced by the CPU oracle
(tests only) so that partitioning, sharded execution, the result gather and the timing
reductions are exercised without a GPU."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _oracle_align(lp, t_off, labels, l_off, V, beam_size, max_move, device):
    from oracle import ctc_oracle
    return ctc_oracle.ctc_best_path_batch(lp, t_off, labels, l_off, beam_size, max_move, n_threads=1)


def _make(n=9):
    from kokoro_align_b200 import synth
    T = np.array([40, 300, 120, 75, 900, 33, 210, 64, 500][:n])
    L = np.maximum(1, np.round(0.14 * T)).astype(np.int64)
    lps, labs = [], []
    for k, (t, l) in enumerate(zip(T, L)):
        a, b = synth.make_lattice(int(t), int(l), 39, seed=900 + k)
        lps.append(a)
        labs.append(b)
    return lps, labs


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from kokoro_align_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lps, labs = _make()
    res = parallel.align_sharded(lps, labs, align_fn=_oracle_align)
    tmax = parallel.max_over_ranks(1.0 + rank)
    tsum = parallel.sum_over_ranks(10.0 * (rank + 1))
    mine = parallel.shard_batch([len(x) for x in lps], [len(x) for x in labs], rank, world)
    if rank == 0:
        q.put(([(r[0].tolist(), r[3], r[4]) for r in res], tmax, tsum, mine.tolist()))
    else:
        assert res is None
        q.put((None, tmax, tsum, mine.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_alignment_world2():
    from oracle import ctc_oracle
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res = next(o[0] for o in outs if o[0] is not None)
    lps, labs = _make()
    assert len(res) == len(lps)
    for (path, final, status), lp, lab in zip(res, lps, labs):
        rp, _, _, rf = ctc_oracle.ctc_best_path(lp, lab, return_final_score=True)
        assert status == 0 and path == rp.tolist() and np.float32(final) == np.float32(rf)
    assert all(o[1] == 2.0 for o in outs) and all(o[2] == 30.0 for o in outs)
    shards = sorted(i for o in outs for i in o[3])
    assert shards == list(range(len(lps)))          # every lattice exactly once


def _files_worker(rank, world, port, q, d):
    import torch.distributed as dist
    from kokoro_align_b200 import align
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 7
    lf = [os.path.join(d, f"c{k}.logits.npz") for k in range(n)]
    vf = [os.path.join(d, f"c{k}.voca.txt") for k in range(n)]
    bf = [os.path.join(d, f"c{k}.best_path.npz") for k in range(n)]
    t = {}
    written = align.best_path_files(lf, vf, bf, verbose=False, align_fn=_oracle_align, timings=t)
    q.put((rank, written, t["chapters"]))
    dist.barrier()
    dist.destroy_process_group()


def test_best_path_files_world2(tmp_path):
    """The per-book entry under world_size 2: every rank reads, aligns (oracle stand-in) and
    writes its own LPT shard of the chapters; rank 0 returns all written paths in order; the
    files equal the single-lattice oracle on the host-normalised logits."""
    from kokoro_align_b200 import align
    from oracle import ctc_oracle
    rng = np.random.default_rng(44)
    d = str(tmp_path)
    Ts = [400, 90, 1500, 33, 800, 260, 1]
    np.savez(os.path.join(d, "c3.best_path.npz"), best_path=np.zeros(1, np.int32))   # exists: skipped
    for k, T in enumerate(Ts):
        np.savez(os.path.join(d, f"c{k}.logits.npz"), data=(rng.standard_normal((T, 39)) * 3).astype(np.float32),
                 indices=np.array([T], np.int32))
        with open(os.path.join(d, f"c{k}.voca.txt"), "w") as f:
            for _ in range(max(1, T // 60) if T > 1 else 0):
                f.write("t|k o k o r o\n")
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_files_worker, args=(r, world, port, q, d)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict((o[0], o[1:]) for o in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = [os.path.join(d, f"c{k}.best_path.npz") for k in range(7) if k != 3]
    assert outs[0][0] == expect and outs[1][0] == []
    assert outs[0][1] + outs[1][1] == 6 and outs[0][1] > 0 and outs[1][1] > 0
    for k in (0, 1, 2, 4, 5, 6):
        with np.load(os.path.join(d, f"c{k}.logits.npz")) as f:
            lp = align.log_softmax(f["data"])
        lab = align.read_transcript_labels(os.path.join(d, f"c{k}.voca.txt"))
        rp, rl, rs = ctc_oracle.ctc_best_path(lp, lab)
        with np.load(os.path.join(d, f"c{k}.best_path.npz")) as f:
            assert f["best_path"].dtype == np.int32 and f["best_scores"].dtype == np.float32
            assert f["best_path"].tolist() == rp.tolist() and f["best_labels"].tolist() == rl.tolist()
            assert f["best_scores"].tobytes() == rs.tobytes()


def test_lpt_partition_properties():
    from kokoro_align_b200 import parallel
    rng = np.random.default_rng(0)
    costs = rng.integers(1, 1000, 200)
    for n in (1, 2, 4, 8):
        parts = parallel.lpt_partition(costs, n)
        allidx = np.sort(np.concatenate(parts))
        assert allidx.tolist() == list(range(200))
        loads = np.array([costs[p].sum() for p in parts])
        assert loads.max() - costs.sum() / n <= costs.max()
    assert parallel.lpt_partition([], 3)[0].size == 0


def test_cells_eval_matches_oracle():
    from kokoro_align_b200 import parallel
    from oracle import ctc_oracle
    for T, L, W in ((81135, 11359, 1000), (861, 121, 1000), (100, 300, 20), (1500, 450, 1000), (1, 0, 5)):
        assert parallel.cells_eval(T, L, W) == ctc_oracle.cells_eval(T, L, W)


def _late_rank_worker(rank, world, port, q, d, mode):
    """mode 'late': rank 1 enters best_path_files after rank 0 has finished its own shard (its
    outputs already exist when rank 1 would scan); mode 'error': rank 1's alignment raises;
    mode 'nothing': every output exists."""
    import time
    import torch.distributed as dist
    from kokoro_align_b200 import align
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 6
    lf = [os.path.join(d, f"c{k}.logits.npz") for k in range(n)]
    vf = [os.path.join(d, f"c{k}.voca.txt") for k in range(n)]
    bf = [os.path.join(d, f"c{k}.best_path.npz") for k in range(n)]

    def fn(*a):
        if mode == "late" and rank == 1:
            time.sleep(1.0)      # rank 0's files are on disk long before this rank is done
        if mode == "error" and rank == 1:
            raise align.KabError("injected failure on rank 1")
        return _oracle_align(*a)
    if mode == "late" and rank == 1:
        time.sleep(0.5)
    try:
        written = align.best_path_files(lf, vf, bf, verbose=False, align_fn=fn)
        q.put((rank, "ok", written))
    except Exception as e:   # noqa: BLE001
        q.put((rank, type(e).__name__, str(e)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["late", "error", "nothing"])
def test_best_path_files_world2_consistent_todo_and_errors(tmp_path, mode):
    """The list of chapters to align is decided once (rank 0) and broadcast, so a rank that starts
    late cannot see the other rank's fresh outputs and compute a different partition; an error on
    one rank is raised on every rank AFTER the collective (nobody hangs); nothing to do returns []
    on every rank."""
    rng = np.random.default_rng(45)
    d = str(tmp_path)
    Ts = [400, 90, 700, 33, 500, 260]
    for k, T in enumerate(Ts):
        np.savez(os.path.join(d, f"c{k}.logits.npz"), data=(rng.standard_normal((T, 39)) * 3).astype(np.float32),
                 indices=np.array([T], np.int32))
        with open(os.path.join(d, f"c{k}.voca.txt"), "w") as f:
            for _ in range(max(1, T // 60)):
                f.write("t|k o k o r o\n")
        if mode == "nothing":
            np.savez(os.path.join(d, f"c{k}.best_path.npz"), best_path=np.zeros(1, np.int32))
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_late_rank_worker, args=(r, world, port, q, d, mode)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict((o[0], o[1:]) for o in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if mode == "late":
        expect = [os.path.join(d, f"c{k}.best_path.npz") for k in range(6)]
        assert outs[0] == ("ok", expect) and outs[1] == ("ok", [])
        for k in range(6):   # every chapter aligned exactly once, by its own rank
            with np.load(os.path.join(d, f"c{k}.best_path.npz")) as f:
                assert len(f["best_path"]) == Ts[k]
    elif mode == "error":
        assert outs[0][0] == "KabError" and outs[1][0] == "KabError"
        assert "injected failure" in outs[0][1]
    else:
        assert outs[0] == ("ok", []) and outs[1] == ("ok", [])
