"""`Trainer` -- the reference's GPT train loop (ttts/gpt/train.py:41-139) on the B200-native engine.

Same public surface (`Trainer(cfg_path)`, `.train()`, `.save(milestone)`, `.load(path)`; config JSON layout of
ttts/gpt/config.json; checkpoint dict {'step', 'model': state_dict}), same step definition:

    loss = text_weight*loss_text + mel_weight*loss_mel ; backward ; get_grad_norm ; clip_grad_norm_(1.0) ;
    AdamW(lr*warmup(step), betas (0.9, 0.96), wd 0.01) ; zero_grad ; scheduler.step()

but the step body is `FusedStep`: forward + backward in the CUDA engine, ONE all-reduce (NCCL over NVLink/NVSwitch via
torch.distributed) of the flat gradient buffer per optimizer step -- chunked by layer group on a side stream so it overlaps
the rest of backward -- and a fused global-norm + clip + AdamW kernel that also refreshes the bf16 weight shadow.  The
reference's 12*L+12 blocking `.item()` calls per step (train.py:22-31) disappear; the norm stays on the device and is
read only when logging.  `accelerate` is not required (and not installed here); one process per GPU is launched with
torchrun / `accelerate launch` alike (RANK / LOCAL_RANK / WORLD_SIZE env).
"""
import json
import os
from datetime import datetime
from pathlib import Path

import torch
import torch.distributed as dist

from .model import UnifiedVoice


def warmup(step):
    """ttts/gpt/train.py:36-40"""
    return float(step / 500) if step < 500 else 1


class FusedStep:
    """One optimizer step of UnifiedVoice on one GPU (one rank of a data-parallel job)."""

    def __init__(self, model, lr=1e-4, betas=(0.9, 0.96), weight_decay=0.01, eps=1e-8, max_norm=1.0, text_weight=0.01, mel_weight=1.0,
                 accumulate=1, comm_chunks=6, process_group=None):
        self.model = model
        self.eng = model._engine()
        self.eng.trust_version = True
        self.lr, self.betas, self.weight_decay, self.eps, self.max_norm = lr, betas, weight_decay, eps, max_norm
        self.text_weight, self.mel_weight = text_weight, mel_weight
        self.accumulate = accumulate
        self.opt_step = 0           # number of optimizer steps taken
        self.sched_step = 0         # LambdaLR step index
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.comm_stream = torch.cuda.Stream(device=self.eng.device) if self.world > 1 else None
        L = model.layers
        # TTTS_COMM_CHUNKS overrides the number of overlapped all-reduce chunks; 0 = ONE all-reduce of the whole buffer after backward
        comm_chunks = int(os.environ.get("TTTS_COMM_CHUNKS", comm_chunks))
        self.overlap = comm_chunks > 0
        n = max(1, min(comm_chunks, L))
        # stage 0 = heads/final norms, stages 1..L = layers L-1..0, stage L+1 = embeddings
        bounds = [0, 1] + [1 + (L * (i + 1)) // n for i in range(n)]
        bounds[-1] = L + 1
        self.chunks = [(bounds[i], bounds[i + 1]) for i in range(len(bounds) - 1)] + [(L + 1, L + 2)]
        self._micro = 0
        self._seed = 0x1234ABCD

    def _grad_range(self, s0, s1):
        lay = self.eng.layout
        lo, hi = None, None
        for s in range(s0, s1):
            b, e = lay.stage_range(s)
            lo = b if lo is None else min(lo, b)
            hi = e if hi is None else max(hi, e)
        return lo, hi

    def micro_step(self, text, text_lengths, codes, wav_lengths, clip_inputs=True):
        """forward + backward of one micro-batch (device tensors).  Returns the device tensor [loss_text, loss_mel]."""
        with torch.cuda.nvtx.range("ttts.step.micro"):
            return self._micro_step(text, text_lengths, codes, wav_lengths, clip_inputs)

    def _micro_step(self, text, text_lengths, codes, wav_lengths, clip_inputs=True):
        m, eng = self.model, self.eng
        TL, CL = text.shape[1], codes.shape[1]
        if clip_inputs:
            TL = min(TL, int(text_lengths.max()))
            CL = min(CL, int(wav_lengths.max()) // m.mel_length_compression)
        drop_p = m.dropout_p if m.training else 0.0
        self._seed = (self._seed * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        eng.refresh_shadow()
        eng.forward(text, codes, wav_lengths, TL, CL, save=True, drop_p=drop_p, seed=self._seed)
        if self._micro == 0:
            eng.grads.zero_()
        last = (self._micro == self.accumulate - 1)
        wt, wm = self.text_weight / self.accumulate, self.mel_weight / self.accumulate
        if self.world > 1 and last and not self.overlap:
            eng.backward(weight_text=wt, weight_mel=wm)
            dist.all_reduce(eng.grads, op=dist.ReduceOp.SUM, group=self.pg)
        elif self.world > 1 and last:
            main = torch.cuda.current_stream()
            for (s0, s1) in self.chunks:
                eng.backward(weight_text=wt, weight_mel=wm, stage_begin=s0, stage_end=s1)
                lo, hi = self._grad_range(s0, s1)
                ev = torch.cuda.Event()
                ev.record(main)
                self.comm_stream.wait_event(ev)
                with torch.cuda.stream(self.comm_stream), torch.cuda.nvtx.range("ttts.step.allreduce_chunk"):
                    dist.all_reduce(eng.grads[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)
            main.wait_stream(self.comm_stream)
        else:
            eng.backward(weight_text=wt, weight_mel=wm)
        self._micro = (self._micro + 1) % self.accumulate
        return eng.losses

    def skip_micro_step(self):
        """A micro-step without a batch (the collater dropped every sample): contributes a zero gradient, keeps the micro-step counter and
        the collective schedule in phase with the other ranks."""
        eng = self.eng
        if self._micro == 0:
            eng.grads.zero_()
        last = (self._micro == self.accumulate - 1)
        if self.world > 1 and last:
            dist.all_reduce(eng.grads, op=dist.ReduceOp.SUM, group=self.pg)
        self._micro = (self._micro + 1) % self.accumulate

    def optimizer_step(self):
        """get_grad_norm + clip_grad_norm_(max_norm) + AdamW + scheduler.step (ttts/gpt/train.py:114-120)."""
        eng = self.eng
        torch.cuda.nvtx.range_push("ttts.step.clip_adamw")
        eng.grad_norm()                      # norm of the SUMMED gradient; the kernel rescales by 1/world
        self.opt_step += 1
        lr = self.lr * warmup(self.sched_step)
        eng.adamw(lr, self.opt_step, betas=self.betas, eps=self.eps, weight_decay=self.weight_decay, max_norm=self.max_norm,
                  grad_scale=1.0 / self.world)
        torch.cuda.nvtx.range_pop()
        self.sched_step += 1
        return eng.norm

    def __call__(self, text, text_lengths, codes, wav_lengths, clip_inputs=True):
        losses = self.micro_step(text, text_lengths, codes, wav_lengths, clip_inputs)
        if self._micro == 0:
            self.optimizer_step()
        return losses


def cycle(dl):
    """the reference's `cycle` (ttts/gpt/train.py:32-35), plus `sampler.set_epoch` so that a DistributedSampler reshuffles every epoch"""
    epoch = 0
    while True:
        sampler = getattr(dl, "sampler", None)
        if hasattr(sampler, "set_epoch"):
            sampler.set_epoch(epoch)
        for data in dl:
            yield data
        epoch += 1


class Trainer(object):
    def __init__(self, cfg_path="ttts/gpt/config.json", cfg=None, dataloader=None, device=None, logs=True):
        self.cfg = cfg if cfg is not None else json.load(open(cfg_path))
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        local = int(os.environ.get("LOCAL_RANK", "0"))
        self.device = torch.device(device) if device is not None else torch.device("cuda", local)
        torch.cuda.set_device(self.device)
        if self.world > 1 and not dist.is_initialized():
            dist.init_process_group("nccl", device_id=self.device)
        self.is_main_process = self.rank == 0
        self.gpt = UnifiedVoice(**self.cfg["gpt"]).to(self.device)
        if self.world > 1:      # identical replicas: rank 0 is authoritative (what DDP's constructor does)
            dist.broadcast(self.gpt._flat, src=0)
        tr = self.cfg["train"]
        self.train_steps = tr["train_steps"]
        self.val_freq = tr["val_freq"]
        self.gradient_accumulate_every = tr["accumulate_num"]
        self.mel_loss_weight = tr["mel_weight"]
        self.text_loss_weight = tr["text_weight"]
        self.fused = FusedStep(self.gpt, lr=tr["lr"], text_weight=self.text_loss_weight, mel_weight=self.mel_loss_weight,
                               accumulate=self.gradient_accumulate_every)
        if dataloader is None:
            from .dataset import build_dataloader       # optional: needs the reference's on-disk dataset
            dataloader = build_dataloader(self.cfg, self.rank, self.world)
        self.dataloader = cycle(dataloader)
        self.step = 0
        self.logs_folder = None
        self.writer = None
        if self.is_main_process and logs:
            now = datetime.now()
            self.logs_folder = Path(tr["logs_folder"] + "/" + now.strftime("%Y-%m-%d-%H-%M-%S"))
            self.logs_folder.mkdir(exist_ok=True, parents=True)
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.writer = SummaryWriter(log_dir=self.logs_folder)
            except Exception:
                self.writer = None

    # ---- checkpoints: same dict as the reference (weights + step only, train.py:70-88) ----
    def save(self, milestone):
        if not self.is_main_process or self.logs_folder is None:
            return
        data = {"step": self.step, "model": self.gpt.state_dict()}
        torch.save(data, str(self.logs_folder / f"model-{milestone}.pt"))

    def load(self, model_path):
        data = torch.load(model_path, map_location=self.device)
        self.step = data["step"]
        self.gpt.load_state_dict(data["model"])

    def train_step(self, data):
        """One optimizer step from a collated host batch dict (the reference's loop body, train.py:99-121).
        Returns (total_loss: float, loss_text, loss_mel, grad_norm) -- the last three are device tensors."""
        total_loss = 0.0
        losses = norm = None
        for i in range(self.gradient_accumulate_every):
            if i > 0:
                data = next(self.dataloader)                 # one batch per micro-step, fetched at the top like the reference (train.py:100)
            if data is None:
                # every sample of the batch was dropped (train.py:101-102 `continue`).  The micro-step still counts: the all-reduce of the
                # last micro-step must be issued on every rank, or the ranks that did get a batch wait for this one forever.
                self.fused.skip_micro_step()
                continue
            inp = [data["padded_text"], data["text_lengths"], data["padded_qmel"], data["wav_lens"]]
            inp = [d.to(self.device, non_blocking=True) for d in inp]
            losses = self.fused.micro_step(*inp)
            lt, lm = losses.tolist()                         # the reference's loss.item() (train.py:111)
            total_loss += (lt * self.text_loss_weight + lm * self.mel_loss_weight) / self.gradient_accumulate_every
        if losses is not None or self.fused.world > 1:
            norm = self.fused.optimizer_step()
        return total_loss, losses, norm

    def train(self):
        self.gpt.train()
        while self.step < self.train_steps:
            data = next(self.dataloader)
            total_loss, losses, norm = self.train_step(data)
            if self.is_main_process and self.step % self.val_freq == 0 and losses is not None:
                lt, lm = losses.tolist()
                scalars = {"loss": total_loss, "loss_mel": lm, "loss_text": lt, "loss/grad": float(norm.item()) / self.fused.world,
                           "lr": self.fused.lr * warmup(max(self.fused.sched_step - 1, 0))}
                if self.writer is not None:
                    for k, v in scalars.items():
                        self.writer.add_scalar(k, v, self.step)
            if self.is_main_process and self.step % self.cfg["train"]["save_freq"] == 0:
                self.save(self.step // 1000)
            self.step += 1
        if self.is_main_process:
            print("training complete")


if __name__ == "__main__":
    Trainer().train()
