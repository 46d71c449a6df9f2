import os

import numpy as np
import pytest

import helpers
from herro_b200 import shard
from herro_b200 import weights as hbw


def test_weights_blob_roundtrip(tmp_path):
    cfg = hbw.NetConfig(stem_k=9, channels=64, heads=2, layers=1, ffn=128, collapse=64)
    T = hbw.random_weights(cfg, seed=5)
    p = str(tmp_path / "m.hbw")
    hbw.save_blob(p, cfg, T)
    cfg2, T2 = hbw.load_blob(p)
    assert cfg2 == cfg and set(T2) == set(T)
    for k in T:
        assert np.array_equal(T[k], T2[k])
    assert np.all(T["emb"][11] == 0)  # padding_idx


def test_export_from_torch_state_dict_roundtrip(tmp_path):
    import torch
    from oracle import forward_ref
    from tools import export_weights
    cfg = hbw.NetConfig(stem_k=9, channels=64, heads=2, layers=1, ffn=128, collapse=64)
    T = hbw.random_weights(cfg, seed=6)
    net = forward_ref.from_weights(cfg, T)
    dims, T2 = export_weights.state_dict_to_tensors(net.state_dict())
    assert dims == dict(stem_k=9, channels=64, layers=1, ffn=128, collapse=64)
    for k in T:
        assert np.allclose(T[k], T2[k]), k


def test_forward_ref_gather_first_equals_full_forward():
    """The property the CUDA forward relies on: logits at `indices` depend only on rows within
    the stem halo, with batch padding rows (token 11 / qual 126) up to Lmax and zeros beyond."""
    import torch
    from oracle import forward_ref
    cfg = hbw.NetConfig(stem_k=9, channels=64, heads=2, layers=1, ffn=128, collapse=64)
    net = forward_ref.from_weights(cfg, hbw.random_weights(cfg, seed=7))
    rng = np.random.default_rng(0)
    L1, L2 = 60, 45
    bases = np.full((2, L1, 31), 11, np.uint8)
    quals = np.full((2, L1, 31), 126, np.uint8)
    bases[0] = rng.integers(0, 11, (L1, 31)); quals[0] = rng.integers(33, 90, (L1, 31))
    bases[1, :L2] = rng.integers(0, 11, (L2, 31)); quals[1, :L2] = rng.integers(33, 90, (L2, 31))
    idx = [np.array([0, 5, 59], np.int32), np.array([2, 44], np.int32)]
    info, bl = forward_ref.run_batch(net, bases, quals, np.array([3, 2], np.int32), idx)
    # window 1 alone but padded to the same Lmax gives the same logits; cropping the pad rows does not
    info1, bl1 = forward_ref.run_batch(net, bases[1:2], quals[1:2], np.array([2], np.int32), idx[1:])
    assert np.allclose(bl[1], bl1[0], atol=1e-6)
    info2, bl2 = forward_ref.run_batch(net, bases[1:2, :L2], quals[1:2, :L2], np.array([2], np.int32), idx[1:])
    assert np.allclose(bl[1][0], bl2[0][0], atol=1e-6)       # row 2 is far from the end: unaffected
    assert not np.allclose(bl[1][1], bl2[0][1], atol=1e-4)   # row 44 sees the pad rows (H10)


def test_shard_targets_partition():
    rng = np.random.default_rng(1)
    lens = rng.integers(4096, 60000, 1000)
    for world in (1, 2, 4, 8):
        parts = [shard.shard_targets(lens, 4096, r, world) for r in range(world)]
        allr = np.concatenate(parts)
        assert np.array_equal(allr, np.arange(1000))  # disjoint, complete, contiguous, ordered
        wins = [int(((lens[p] + 4095) // 4096).sum()) for p in parts]
        assert max(wins) - min(wins) <= 2 * 15 + 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lens = np.full(100, 9000)
    mine = shard.shard_targets(lens, 4096, rank, world)
    t, u = shard.reduce_throughput(dist, seconds=1.0 + rank, units=float(len(mine)))
    q.put((rank, t, u, len(mine)))
    dist.destroy_process_group()


def test_gloo_world2_reduction():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=120) for _ in ps)
    [p.join(timeout=60) for p in ps]
    assert res[0][1] == res[1][1] == 2.0          # max over ranks
    assert res[0][2] == res[1][2] == 100.0        # all units accounted for exactly once
    assert res[0][3] + res[1][3] == 100


def _shard_worker(rank, world, port, q):
    """What a rank of `bench.py --gpus N` does on the host: same read set everywhere (deterministic generator), alignments only
    for its own read-id shard, one all_gather of the per-rank totals (gloo here, nccl on the box)."""
    import torch
    import torch.distributed as dist
    from tools import synth
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = synth.Generator(120, 6000, seed=5, coverage=15.0, min_ovl=1200, threads=2)
    lens = g.read_lens()
    n_job = 100
    mine = shard.shard_targets(lens[:n_job], 4096, rank, world)
    lo, hi = int(mine[0]), int(mine[-1]) + 1
    rs = g.readset(targets=(lo, hi))
    g.close()
    import hashlib
    store = hashlib.sha1(rs.seqs.tobytes() + rs.quals.tobytes()).hexdigest()
    vals = torch.tensor([float(len(rs.ovl9)), float(len(rs.cigars)), float(hi - lo)], dtype=torch.float64)
    gathered = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(gathered, vals)
    q.put((rank, store, lo, hi, [g_.tolist() for g_ in gathered]))
    dist.destroy_process_group()


def test_gloo_world2_sharded_generation_covers_the_job_once():
    import torch.multiprocessing as mp
    from tools import synth
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + ((os.getpid() + 7) % 2000)
    ps = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted(q.get(timeout=180) for _ in ps)
    [p.join(timeout=60) for p in ps]
    (r0, s0, lo0, hi0, g0), (r1, s1, lo1, hi1, g1) = res
    assert s0 == s1                                   # every rank holds the same read store
    assert lo0 == 0 and hi0 == lo1 and hi1 == 100     # contiguous shards covering the job exactly once
    assert g0 == g1                                   # the gathered totals agree on both ranks
    whole = synth.generate(120, 6000, seed=5, coverage=15.0, min_ovl=1200, targets=(0, 100))
    assert sum(x[0] for x in g0) == len(whole.ovl9) and sum(x[1] for x in g0) == len(whole.cigars)
