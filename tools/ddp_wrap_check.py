"""The drop-in module under the wrapper the reference's trainer puts around it (`accelerator.prepare` = DistributedDataParallel):
UnifiedVoice wrapped in torch's DDP, driven by the loop body of ttts/gpt/train.py:99-121 restated (torch AdamW, clip_grad_norm_, zero_grad),
2+ ranks, versus ONE process stepping on the concatenated batch.  Also the module's own single flat all-reduce
(`UnifiedVoice.enable_flat_allreduce`, no wrapper): the same update from ONE NCCL call per step instead of DDP's 25 MB buckets.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/ddp_wrap_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from torch.nn.parallel import DistributedDataParallel as DDP
from ttts_b200.gpt.model import UnifiedVoice
from ttts_b200.gpt import synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
kw = dict(layers=4, model_dim=256, heads=4, max_text_tokens=60, max_mel_tokens=200, number_text_tokens=256, start_text_token=255,
          number_mel_codes=1026, start_mel_token=1024, stop_mel_token=1025)
b, TL, CL, STEPS = 2, 24, 100, 3


def make():
    torch.manual_seed(0)
    return UnifiedVoice(**kw).to(dev).eval()       # eval: dropout off, so data parallel == big batch up to fp32 summation order


def loop(model, params, batch, steps=STEPS):
    """ttts/gpt/train.py:99-121: forward, 0.01 * loss_text + loss_mel, backward, clip_grad_norm_(1.0), AdamW step, zero_grad"""
    opt = torch.optim.AdamW(params, lr=1e-3, betas=(0.9, 0.96), weight_decay=0.01)
    for _ in range(steps):
        loss_text, loss_mel, mel_logits = model(batch[0], batch[1], batch[2].clone(), batch[3])
        loss = loss_text * 0.01 + loss_mel * 1.0
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad()
    torch.cuda.synchronize()


batches = [synth.synthetic_batch(b, TL, CL, seed=100 + r) for r in range(world)]
mine = [t.to(dev) for t in batches[rank]]
results = {}
# (1) torch DistributedDataParallel around the module, as accelerate does
m = make()
ddp = DDP(m, device_ids=[local])
loop(ddp, list(ddp.parameters()), mine)
results["DistributedDataParallel"] = m._flat.clone()
# (2) no wrapper: the module all-reduces its flat gradient buffer itself at the end of backward (ONE collective per step)
m2 = make()
m2.enable_flat_allreduce()
loop(m2, list(m2.parameters()), mine)
results["flat all-reduce"] = m2._flat.clone()
for name, flat in results.items():
    gathered = [torch.zeros_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        ref = make()
        big = [torch.cat([batches[r][i] for r in range(world)], dim=0).to(dev) for i in range(4)]
        loop(ref, list(ref.parameters()), big)
        p0 = make()._flat
        num = (flat - ref._flat).norm().item(); den = (ref._flat - p0).norm().item()
        print("%-24s replicas identical: %s   |data parallel - big batch| / |update| = %.3e" % (name, same, num / den), flush=True)
        assert same and num / den < 2e-2, name
if rank == 0:
    print("DDP WRAP CHECK PASS", flush=True)
dist.barrier()
dist.destroy_process_group()
