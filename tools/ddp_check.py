"""2+ GPU check of the data-parallel step (run under torchrun): N ranks x per-rank batch b, one FusedStep each, versus ONE rank
stepping on the concatenated batch N*b.  With dropout off the averaged gradient -- hence the AdamW update -- must agree.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from ttts_b200.gpt.model import UnifiedVoice
from ttts_b200.gpt.train import FusedStep
from ttts_b200.gpt import synth

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
kw = dict(layers=4, model_dim=256, heads=4, max_text_tokens=60, max_mel_tokens=200, number_text_tokens=256, start_text_token=255,
          number_mel_codes=1026, start_mel_token=1024, stop_mel_token=1025)
torch.manual_seed(0)
b, TL, CL = 2, 24, 100


def make():
    torch.manual_seed(0)
    m = UnifiedVoice(**kw).to(dev).eval()          # eval: dropout off, so DDP == big-batch exactly (up to fp32 summation order)
    return m


m = make()
dist.broadcast(m._flat, src=0)
fused = FusedStep(m, lr=1e-3)
fused.sched_step = 600
batches = [synth.synthetic_batch(b, TL, CL, seed=100 + r) for r in range(world)]
mine = [t.to(dev) for t in batches[rank]]
for _ in range(2):
    fused(*mine, clip_inputs=False)
torch.cuda.synchronize()
flat = m._flat.clone()
gathered = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
ok = all(torch.equal(gathered[0], g) for g in gathered)          # replicas stay bit-identical
if rank == 0:
    ref = make()
    solo = FusedStep(ref, lr=1e-3, process_group=None)
    solo.world = 1; solo.comm_stream = None
    solo.sched_step = 600
    big = [torch.cat([batches[r][i] for r in range(world)], dim=0).to(dev) for i in range(4)]
    for _ in range(2):
        solo(*big, clip_inputs=False)
    torch.cuda.synchronize()
    p0 = make()._flat
    num = (flat - ref._flat).norm().item(); den = (ref._flat - p0).norm().item()
    print("replicas identical:", ok, " |ddp - bigbatch| / |update| = %.3e" % (num / den), flush=True)
    assert ok and num / den < 2e-2
    print("DDP CHECK PASS", flush=True)
dist.barrier()
dist.destroy_process_group()
