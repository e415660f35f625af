"""`Trainer` of the diffusion mel-refiner with the surface of the reference's (ttts/diffusion/train.py:78-256): `.train()`, `.save(milestone)`,
`.load(path)`; checkpoints are `{'step', 'model'}` with the reference's AA_diffusion state_dict keys, so they interchange both ways.

Differences from the reference's constructor, all about things outside the hot path (SURVEY.md section 2 OOS): the dataloader (batches in the
layout of `DiffusionCollater`, ttts/diffusion/dataset.py:76-110: padded_text, padded_mel_code, padded_mel, padded_mel_refer) and the frozen
GPT (`ttts_b200.gpt.model.UnifiedVoice`, already on the GPU) are handed in instead of being built from paths in a YAML; the evaluation
sampler / vocoder / tensorboard side of the loop is not reproduced."""
import torch

from .kernels import DiffusionCudaKernels
from .params import default_config, init_params
from .train_step import DiffusionStep, normalize_tacotron_mel


class Trainer:
    def __init__(self, gpt, dataloader, cfg=None, lr=1e-4, accumulate_num=1, unconditioned_percentage=0.1, layer_drop=0.1, train_steps=1000000,
                 params=None, device="cuda", seed=0):
        self.cfg = cfg or default_config()
        self.gpt, self.dataloader, self.train_steps, self.device = gpt, dataloader, train_steps, torch.device(device)
        params = params if params is not None else init_params(self.cfg, seed=seed, device=self.device)
        self.stepper = DiffusionStep(DiffusionCudaKernels(), {k: v.to(self.device) for k, v in params.items()}, self.cfg, lr=lr,
                                     accumulate_num=accumulate_num, unconditioned_percentage=unconditioned_percentage, layer_drop=layer_drop, seed=seed)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.step = 0

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.stepper.opt.params().items()}

    def save(self, path):
        torch.save({"step": self.step, "model": self.state_dict()}, path)

    def load(self, path):
        data = torch.load(path, map_location=self.device)
        views = self.stepper.opt.params()
        for k, v in data["model"].items():                            # load_state_dict(strict=False) of the reference (train.py:147)
            if k in views:
                views[k].copy_(v)
        self.step = self.stepper.step_no = data["step"]

    def micro_batch(self, data):
        """train.py:159-170: the frozen GPT's latents (no_grad, return_latent), normalised mels, fresh noise"""
        dev = self.device
        text, codes = data["padded_text"].to(dev), data["padded_mel_code"].to(dev)
        with torch.no_grad():
            latent = self.gpt(text, torch.tensor([text.shape[-1]], device=dev), codes,
                              torch.tensor([codes.shape[-1] * self.gpt.mel_length_compression], device=dev),
                              return_latent=True, clip_inputs=False).transpose(1, 2).float().contiguous()
        x_start = normalize_tacotron_mel(data["padded_mel"].to(dev).float()).contiguous()
        refer = normalize_tacotron_mel(data["padded_mel_refer"].to(dev).float()).contiguous()
        noise = torch.randn(x_start.shape, device=dev, generator=self.gen)
        return dict(x_start=x_start, latent=latent, refer=refer, noise=noise)

    def train(self, log=None):
        it = iter(self.dataloader)
        while self.step < self.train_steps:
            batches = []
            while len(batches) < self.stepper.accumulate_num:
                try:
                    data = next(it)
                except StopIteration:
                    it = iter(self.dataloader)
                    data = next(it)
                if data is None:                                      # every sample of the batch failed to load (dataset.py:52)
                    continue
                batches.append(self.micro_batch(data))
            out = self.stepper.step(batches)
            self.step += 1
            if log is not None:
                log(self.step, out)
