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


def _vq_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ttts_b200.vqvae.quantize import EuclideanCodebook
    g = torch.Generator().manual_seed(7 + rank)
    hist = torch.randint(0, 5, (1024,), generator=g).float()
    esum = torch.randn(1024, 192, generator=g)
    h_all = [torch.zeros_like(hist) for _ in range(world)]
    e_all = [torch.zeros_like(esum) for _ in range(world)]
    dist.all_gather(h_all, hist); dist.all_gather(e_all, esum)
    h1, e1 = EuclideanCodebook.sync_stats(hist.clone(), esum.clone(), "allreduce")
    ok = torch.equal(h1, sum(h_all)) and torch.allclose(e1, sum(e_all), atol=1e-6)          # global-batch statistics on every rank
    h2, e2 = EuclideanCodebook.sync_stats(hist.clone(), esum.clone(), "rank0")
    ok = ok and torch.equal(h2, h_all[0]) and torch.equal(e2, e_all[0])                       # the reference's broadcast_buffers semantics
    h3, e3 = EuclideanCodebook.sync_stats(hist.clone(), esum.clone(), None)
    ok = ok and torch.equal(h3, hist) and torch.equal(e3, esum)
    try:
        EuclideanCodebook.sync_stats(hist.clone(), esum.clone(), "bogus")
        ok = False
    except ValueError:
        pass
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_codebook_statistics_sync_modes():
    """the VQ codebook EMA under data parallelism -- all-reduce of (histogram, embedding sums) or rank-0 broadcast."""
    mgr = mp.Manager()
    out = mgr.dict()
    port = 31000 + (os.getpid() % 2000)
    mp.spawn(_vq_worker, args=(2, port, out), nprocs=2, join=True)
    assert out[0] and out[1]


def _vqvae_worker(rank, world, port, out):
    """the VQ-VAE step's data parallelism (train_step.gather_and_reduce): the per-rank gradient dict is gathered into one flat buffer in the
    optimizer's name order, summed with ONE all-reduce, and the returned factor turns the sum into DDP's mean"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ttts_b200.vqvae.train_step import gather_and_reduce
    shapes = {"dec.conv_pre.weight": (8, 4, 7), "flow.flows.0.pre.bias": (5,), "enc_q.proj.weight": (6, 3, 1), "enc_p_2.text_embedding.weight": (9, 4)}
    names = list(shapes.keys())

    def grads_for(r):
        g = torch.Generator().manual_seed(500 + r)
        return {k: torch.randn(*shapes[k], generator=g) for k in names}
    flat = torch.zeros(sum(int(torch.tensor(s).prod()) for s in shapes.values()))
    scale = gather_and_reduce(grads_for(rank), names, flat)
    want = sum(torch.cat([grads_for(r)[k].reshape(-1) for k in names]) for r in range(world)) / world
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    out[rank] = bool(scale == 1.0 / world and torch.allclose(flat * scale, want, atol=1e-6) and all(torch.equal(gathered[0], t) for t in gathered))
    dist.destroy_process_group()


def test_vqvae_step_gradients_are_averaged_with_one_allreduce():
    world = 2
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_vqvae_worker, args=(world, 29611, out), nprocs=world, join=True)
        assert all(out[r] for r in range(world))
