"""One optimisation step of the diffusion trainer (ttts/diffusion/train.py:156-203): `accumulate_num` micro-steps of
q_sample -> AA_diffusion -> training_losses(...)["loss"].mean() / accumulate_num -> backward, then get_grad_norm, clip_grad_norm_(1.0),
AdamW(lr 1e-4, betas (0.9, 0.999), weight decay 0.01), LambdaLR warm-up over 1000 steps (train.py:72-76, 116-117).

The frozen GPT's latent (train.py:161-165, `return_latent=True`) is an INPUT of the step: it is produced by `ttts_b200.gpt.model.UnifiedVoice`
(already built).  The random draws of a micro-step -- t, the noise, the unconditioned samples, the dropped layers --
happen here, on the host generator handed in, so that a test can pin them."""
import random

import torch

from ..vqvae.train_step import FlatAdamW, gather_and_reduce
from .train_graph import DiffusionGraph

TACOTRON_MEL_MAX = 5.5451774444795624753378569716654


def normalize_tacotron_mel(mel):
    """aa_model.py:19-21 (input preparation of the trainer, train.py:168-169)"""
    return torch.clamp(mel, min=-TACOTRON_MEL_MAX) * 0.18215


def warmup(step):
    """train.py:72-76"""
    return float(step / 1000) if step < 1000 else 1.0


class FlatAdamWClip(FlatAdamW):
    """FlatAdamW + the trainer's clip_grad_norm_(parameters, 1.0) (train.py:190-191): ttts_grad_norm writes the norm of the flat gradient
    buffer to device memory and the fused AdamW kernel applies min(1, max_norm / norm) while it reads the gradient -- no host sync."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01, max_norm=1.0):
        super().__init__(params, lr, betas, eps, weight_decay)
        self.max_norm = max_norm
        self.norm = torch.zeros(1, dtype=torch.float32, device=self.device)
        self._scratch = torch.empty(1024, dtype=torch.float32, device=self.device)

    def step(self, grads, lr_scale=1.0):
        self.t += 1
        scale = gather_and_reduce(grads, self.names, self.grad)
        L = self.L
        st = L.stream_ptr().value
        if scale != 1.0:
            self.grad.mul_(scale)             # the clip threshold applies to the AVERAGED gradient; one pass over a buffer the kernel re-reads
        L.check(L.lib().ttts_grad_norm(self.grad.data_ptr(), self.grad.numel(), self._scratch.data_ptr(), self.norm.data_ptr(), st), "ttts_grad_norm")
        L.check(L.lib().ttts_adamw_step(self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), None, self.flat.numel(),
                                        self.norm.data_ptr(), float(self.max_norm), 1.0, float(self.lr * lr_scale), float(self.betas[0]),
                                        float(self.betas[1]), float(self.eps), float(self.wd), int(self.t), st), "ttts_adamw_step")
        return self.norm


class DiffusionStep:
    def __init__(self, K, params, cfg, lr=1e-4, accumulate_num=1, unconditioned_percentage=0.1, layer_drop=0.1, optimizer=None, seed=0):
        """`params`: the reference's AA_diffusion state_dict (name -> tensor); `cfg`: model_channels, num_layers, num_heads;
        `optimizer`: class (params, lr) -> .params() / .step(grads, lr_scale); product default FlatAdamWClip."""
        self.K, self.cfg = K, cfg
        self.opt = (optimizer or FlatAdamWClip)(params, lr)
        self.accumulate_num = accumulate_num
        self.p_uncond, self.p_drop = unconditioned_percentage, layer_drop
        self.step_no = 0
        self.rng = random.Random(seed)
        self.gen = torch.Generator().manual_seed(seed)

    def draw(self, B):
        """the random decisions of one micro-step: t (train.py:170), unconditioned samples (aa_model.py:247-249), dropped layers (:269-271)"""
        n_layers = self.cfg["num_layers"] + 3
        t = torch.randint(0, 1000, (B,), generator=self.gen)
        uncond = torch.rand(B, generator=self.gen) < self.p_uncond if self.p_uncond > 0 else None
        dropped = tuple(i for i in range(1, n_layers - 1) if self.p_drop > 0 and self.rng.random() < self.p_drop)
        return t, uncond, dropped

    def micro_step(self, x_start, latent, refer, noise, t=None, uncond=None, dropped=None):
        """forward + backward of one micro-batch; x_start / refer already normalised (normalize_tacotron_mel).  Returns (loss tensor [1], grads)."""
        if t is None:
            t, uncond, dropped = self.draw(x_start.shape[0])
        graph = DiffusionGraph(self.K, self.opt.params(), self.cfg)
        lossv, terms = graph.loss(x_start, t, noise, latent, refer, uncond, dropped)
        if self.accumulate_num > 1:
            lossv = graph.ops.scale(lossv, 1.0 / self.accumulate_num)
        return lossv.v, graph.backward(lossv), terms

    def step(self, batches):
        """`batches`: accumulate_num dicts with x_start, latent, refer, noise (+ optionally t, uncond, dropped).  Returns loss (sum of the
        scaled micro losses, like `total_loss`) and the gradient norm tensor."""
        assert len(batches) == self.accumulate_num
        total, grads = None, None
        for b in batches:
            loss, g, _ = self.micro_step(b["x_start"], b["latent"], b["refer"], b["noise"], b.get("t"), b.get("uncond"), b.get("dropped"))
            total = loss if total is None else self.K.add(total, loss)
            grads = g if grads is None else {k: self.K.add(grads[k], g[k]) for k in grads}
        norm = self.opt.step(grads, warmup(self.step_no))
        self.step_no += 1
        return dict(loss=total, grad_norm=norm)
