"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank computes its local statistics with the oracle
(standing in for the CUDA kernels, which need a GPU), packs them with the product's layout, runs the
product's single packed all-reduce, applies the EMA, and must reproduce (a) the unmodified reference run
under 2-rank gloo (golden fixture) and (b) the single-process result on the concatenated batch."""
import os

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from conftest import load_golden


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import torch.distributed as dist
    from medtok_b200.distributed import all_reduce_stats, pack_stats, shard_rows, unpack_stats
    from oracle import vq_oracle
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import load_golden as lg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    meta, a = lg("hard_gloo2_n320_d64_k512")
    w0 = a["weight0"]
    k, d = w0.shape
    z = a[f"z_{rank}"].reshape(-1, d)
    zf, idx, bins, esum, zq, sse = vq_oracle.hard_local(z, w0)
    stats = pack_stats(torch.from_numpy(bins), torch.from_numpy(esum), sse, float(z.shape[0]))
    all_reduce_stats(stats, dist.all_reduce)                       # the product's exchange step
    gb, ge, gsse, grows = unpack_stats(stats, d, k)
    c = vq_oracle.C()
    cs = c.ema_cluster_size(np.zeros(k, np.float32), gb.numpy(), meta["decay"])
    w1 = c.ema_codebook(w0, gb.numpy(), ge.numpy(), meta["decay"])
    lo, hi = shard_rows(640, world, rank)
    q.put((rank, idx, cs, w1, float(gsse), float(grows), (lo, hi), zq))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_matches_reference_and_single_process(oracle):
    meta, a = load_golden("hard_gloo2_n320_d64_k512")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29741, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    k, d = a["weight0"].shape
    zcat = np.concatenate([a["z_0"].reshape(-1, d), a["z_1"].reshape(-1, d)])
    single = oracle.hard_forward(zcat, a["weight0"], np.zeros(k, np.float32), meta["beta"], meta["decay"])
    for rank, idx, cs, w1, gsse, grows, (lo, hi), zq in res:
        # (a) the unmodified reference under 2-rank gloo
        assert np.array_equal(idx, a[f"idx_{rank}"])
        assert np.array_equal(cs, a[f"cluster_size_{rank}"])
        assert np.array_equal(zq.reshape(a[f"z_q_{rank}"].shape), a[f"z_q_{rank}"])
        assert np.max(np.abs(w1 - a[f"weight_{rank}"])) <= 1e-6
        # (b) single process on the concatenated batch: statistics are fp64, so identical after rounding
        assert np.array_equal(cs, single.cluster_size)
        assert np.array_equal(w1, single.weight)
        assert abs(gsse - single.sse) <= 1e-12 * single.sse and grows == 640
        assert (lo, hi) == (rank * 320, (rank + 1) * 320)
        assert np.array_equal(idx, single.idx[lo:hi])
    # replicas stay bit-identical
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])


def test_shard_rows_covers_everything():
    from medtok_b200.distributed import shard_rows
    for n in (0, 1, 7, 640, 1 << 20, (1 << 20) + 3):
        for w in (1, 2, 3, 8):
            spans = [shard_rows(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))


def _sweep_worker(rank, world, port, tmp, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from medtok_b200.tokenize_sweep import TokenSweep, compose_eval_outputs
    n, d, k = 1003, 8, 5
    g = torch.Generator().manual_seed(3)
    perm = torch.randperm(n, generator=g)                      # the loader's (shuffled) order of the vocabulary
    full = {name: torch.randn(n, d, generator=g) for name in
            ("specific_embedding_text", "specific_embedding_graph", "shared_text_embedding", "shared_graph_embedding")}
    for name in ("text", "graph", "shared_text", "shared_graph"):
        full[name + "_tokens"] = torch.randint(0, 999, (n, k), generator=g)
        full[name + "_tokens_weights"] = torch.rand(n, k, generator=g)
    sweep = TokenSweep(n, 4 * d, k, pin=False)
    mine = perm[rank::world]                                    # DistributedSampler-style interleaving
    for lo in range(0, mine.numel(), 97):                       # ragged last batch
        ci = mine[lo:lo + 97]
        e, t, w = compose_eval_outputs({key: v[ci] for key, v in full.items()})
        sweep.add(ci, e, t, w)
    done = sweep.finish()
    if done:
        sweep.save(tmp)
        e_all, t_all, w_all = compose_eval_outputs(full)
        ok = (np.array_equal(np.load(os.path.join(tmp, "embeddings_all.npy")), e_all.numpy())
              and np.array_equal(np.load(os.path.join(tmp, "tokens_all.npy")), t_all.numpy())
              and np.array_equal(np.load(os.path.join(tmp, "weights_all.npy")), w_all.numpy()))
        q.put(("rank0", ok, tuple(np.load(os.path.join(tmp, "tokens_all.npy")).shape)))
    dist.barrier()
    dist.destroy_process_group()


def test_token_sweep_two_ranks_gloo(tmp_path):
    """Tokenization sweep writer (inference.py:97-134): two ranks own interleaved code indices, rank 0 writes the three
    files once, rows in vocabulary order (the reference orders by the last batch's indices only)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sweep_worker, args=(r, 2, 29747, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    tag, ok, shape = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
    assert tag == "rank0" and ok and shape == (1003, 4, 5)
