"""CPU, world_size 2, gloo: the host-side data-parallel logic -- chunked all-reduce of the flat gradient buffer by
backward stage equals one all-reduce of the whole buffer, and ranks end with identical averaged gradients."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, total, ranges, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    flat = torch.randn(total, generator=g)
    whole = flat.clone()
    dist.all_reduce(whole)
    for lo, hi in ranges:                      # same order as FusedStep.micro_step issues them
        dist.all_reduce(flat[lo:hi])
    ok = torch.equal(flat, whole)
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok = ok and all(torch.equal(gathered[0], t) for t in gathered)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_chunked_allreduce_equals_whole():
    from ttts_b200.gpt import engine as E
    cfg = E.GptConfig(layers=6, model_dim=128, heads=2, max_text_tokens=20, max_mel_tokens=30, n_text_vocab=257, n_mel_vocab=1026,
                      start_text_token=255, stop_text_token=0, start_mel_token=1024, stop_mel_token=1025, mel_length_compression=1024)
    lay = E.Layout(cfg)
    L = cfg.layers
    n = 3
    bounds = [0, 1] + [1 + (L * (i + 1)) // n for i in range(n)]
    chunks = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)] + [(L + 1, L + 2)]
    ranges = []
    for s0, s1 in chunks:
        rs = [lay.stage_range(s) for s in range(s0, s1)]
        ranges.append((min(r[0] for r in rs), max(r[1] for r in rs)))
    covered = sorted(ranges)
    assert covered[0][0] == 0 and covered[-1][1] == lay.total
    for (a0, a1), (b0, b1) in zip(covered, covered[1:]):
        assert a1 == b0
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29000 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, lay.total, ranges, out), nprocs=2, join=True)
    assert out[0] and out[1]
