"""The generator half of one VQ-VAE-GAN train step on ONE tape (next scope row): SynthesizerTrn.forward
(ttts/vqvae/vq2.py:843-871) -> MultiPeriodDiscriminator -> loss_gen_all = loss_gen + loss_fm + loss_mel + kl_ssl + loss_kl
(ttts/vqvae/train.py:336-395), assembled from the module graphs (train_encoder / train_text_encoder / train_flow / train_decoder / train_disc).

DRAFT, NOT YET RUN ON HARDWARE.  Over the torch restatement of the kernel contract it reproduces the five losses and the gradients of all
1455 net_g parameter tensors of the REAL reference step (tests/test_train_step_cpu.py vs tests/golden/vqvae_step.npz).  The posterior noises
and the segment starts are inputs (the reference draws them with torch.randn_like / torch.rand); the TextEncoder's p = 0.1 dropouts are not
drawn yet (eval semantics); the optimizers' step is not part of this file."""
import torch

from .train_decoder import DecoderGraph
from .train_disc import DiscriminatorGraph
from .train_encoder import EncoderGraph, Ops, Tape, Var
from .train_flow import FlowGraph
from .train_text_encoder import TextEncoderGraph

HOP, C_MEL, C_KL = 640, 45.0, 1.0


class GeneratorStep:
    def __init__(self, K, params_g, params_d):
        self.K = K
        self.tape = Tape()
        self.ops = Ops(K, self.tape)
        enc_keys = ("enc_p.", "enc_q.", "ref_enc.", "proj.")
        self.enc = EncoderGraph(K, {k: v for k, v in params_g.items() if k.startswith(enc_keys) and not k.startswith("enc_p_2.")}, self.tape)
        self.te = TextEncoderGraph(K, params_g, self.tape, "enc_p_2.")
        self.flow = FlowGraph(K, params_g, self.tape, "flow.")
        self.dec = DecoderGraph(K, params_g, self.tape, "dec.")
        self.disc = DiscriminatorGraph(K, params_d, self.tape)

    def synthesize(self, wav, spec, lengths, text, text_lengths, codebook, eps_p, eps_q, ids_slice, segment_frames, wav_aug=None, spec_aug=None):
        """SynthesizerTrn.forward (vq2.py:843-871) and the losses that do not involve the discriminators.
        wav [B,L], spec [B,1025,T] (= spectrogram_torch(wav)); wav_aug / spec_aug: the augmented waveform and its spectrogram that `enc_p`
        sees when the quantizer is not frozen (train.py:330-345, vq2.py:849; default = wav / spec, the `freeze_quantizer` branch and the
        benchmark's aug = identity; the Praat augmentation itself is CPU third-party code).  The TextEncoder's
        p = 0.1 dropouts are not drawn (eval semantics).  lengths [B] frames, text [B,Tt] int64, codebook [1024,192]
        (or, with the CUDA backend, the EuclideanCodebook module: its EMA buffers are then updated like in the reference's training forward),
        eps_p / eps_q [B,192,T] posterior noises, ids_slice [B] segment starts (frames)."""
        o, enc = self.ops, self.enc
        B, _, T = spec.shape
        dev = spec.device
        mask2 = (torch.arange(T, device=dev)[None, :] < lengths[:, None]).float().contiguous()
        specv, wavv = Var(spec.contiguous()), Var(wav.unsqueeze(1).contiguous())
        spec_p = specv if spec_aug is None else Var(spec_aug.contiguous())
        wav_p = wavv if wav_aug is None else Var(wav_aug.unsqueeze(1).contiguous())
        ge = enc.mel_style_encoder(Var(self.K.mul_mask(spec.contiguous(), mask2)), mask2, lengths)            # vq2.py:847
        x, _ = enc.posterior_audio_encoder(spec_p, wav_p, mask2, ge, eps_p, "enc_p.")                         # :849 (the noisy sample feeds proj)
        x = o.conv(x, enc.P["proj.weight"], enc.P["proj.bias"], stride=2)                                     # :851
        quantized, commit, codes = o.vq(x, codebook)                                                          # :852-853
        q_up = o.upsample2(quantized)                                                                         # :854-856
        _, stats_p = self.te.forward(q_up, lengths, text, text_lengths, ge)                                   # :857
        m_p, logs_p = o.slice_c(stats_p, 0, 192), o.slice_c(stats_p, 192, 384)
        z, stats_q = enc.posterior_audio_encoder(specv, wavv, mask2, ge, eps_q, "enc_q.")                     # :858
        logs_q = o.slice_c(stats_q, 192, 384)
        z_p = self.flow.forward(z, mask2, ge)                                                                 # :859
        y_hat = self.dec.forward(o.slice_t(z, ids_slice, segment_frames), ge)                                 # :861-864
        # ---- mel / commitment / KL losses (train.py:357-366, 389-390) ----
        L = segment_frames * HOP
        y_mel = torch.stack([self.K.logmel_fwd(wav)[b, :, int(s):int(s) + segment_frames] for b, s in enumerate(ids_slice)]).contiguous()
        loss_mel = o.scale(o.l1_mean(y_mel, o.logmel(o.reshape(y_hat, (B, L)))), C_MEL)
        loss_kl = o.scale(o.kl(z_p, logs_q, m_p, logs_p, mask2), C_KL)
        y_seg = torch.stack([wav[b, int(s) * HOP:int(s) * HOP + L] for b, s in enumerate(ids_slice)]).unsqueeze(1).contiguous()
        self.out = dict(loss_mel=loss_mel, kl_ssl=commit, loss_kl=loss_kl, y_hat=y_hat, y_seg=y_seg, z=z, codes=codes)
        return self.out

    def adversarial(self):
        """generator_loss + feature_loss through the discriminators as they are NOW (the trainer steps optim_d between the synthesis and this
        call, train.py:372-388), then the total"""
        o, out = self.ops, self.out
        _, fmap_r = self.disc.forward(out["y_seg"])
        gen, fmap_g = self.disc.forward(out["y_hat"])
        loss_gen, loss_fm = self.disc.generator_losses(gen, fmap_r, fmap_g)
        total = o.add(o.add(o.add(loss_gen, loss_fm), o.add(out["loss_mel"], out["kl_ssl"])), out["loss_kl"])
        out.update(loss_gen=loss_gen, loss_fm=loss_fm, total=total)
        return out

    def forward(self, wav, spec, lengths, text, text_lengths, codebook, eps_p, eps_q, ids_slice, segment_frames, wav_aug=None, spec_aug=None):
        """synthesize + adversarial with the discriminators unchanged in between.  Returns the dict of loss Vars."""
        self.synthesize(wav, spec, lengths, text, text_lengths, codebook, eps_p, eps_q, ids_slice, segment_frames, wav_aug, spec_aug)
        return self.adversarial()

    def backward(self):
        """d loss_gen_all / d every net_g parameter, by the reference's state_dict names"""
        self.out["total"].g = torch.ones_like(self.out["total"].v)
        self.tape.backward()
        grads = {}
        for graph, prefix in ((self.enc, ""), (self.te, "enc_p_2."), (self.flow, "flow."), (self.dec, "dec.")):
            for k, v in graph.P.items():
                shape = graph.shapes[k]
                grads[prefix + k] = v.g.reshape(shape) if v.g is not None else torch.zeros(shape, device=v.v.device)
        return grads


def gather_and_reduce(grads, names, flat):
    """Gather the named gradients into `flat` (the order of `names`) and, under torch.distributed, SUM them over the ranks with one
    all-reduce (the reference wraps net_g / net_d in DDP, train.py:206-208: mean over ranks).  Returns the factor that turns the sum into the
    mean (1 / world size), which the fused AdamW kernel applies while it reads the gradient."""
    import torch.distributed as dist
    torch.cat([grads[k].reshape(-1) for k in names], out=flat)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        return 1.0 / dist.get_world_size()
    return 1.0


class FlatAdamW:
    """torch.optim.AdamW(params, lr, betas, eps) of the trainer (train.py:292-303: lr 1e-4, betas (0.8, 0.99), eps 1e-9, weight decay 0.01) over
    ONE flat fp32 buffer: the named tensors become views of it, the gradients are gathered into a second flat buffer and ONE launch of the
    fused ttts_adamw_step (the GPT trainer's kernel) updates everything.  `clip_grad_value_(params, None)` of the reference only measures."""

    def __init__(self, params, lr=1e-4, betas=(0.8, 0.99), eps=1e-9, weight_decay=0.01):
        from .. import _lib as L
        from ..gpt import engine as E
        self.L = L
        E._setup_prototypes(L.lib())
        self.names = list(params.keys())
        sizes = [params[k].numel() for k in self.names]
        dev = params[self.names[0]].device
        L.require_cuda(params[self.names[0]])
        self.flat = torch.cat([params[k].detach().reshape(-1).float() for k in self.names]).contiguous()
        self.grad = torch.zeros_like(self.flat)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.views, o = {}, 0
        for k, n in zip(self.names, sizes):
            self.views[k] = self.flat[o:o + n].view(params[k].shape)
            o += n
        self.lr, self.betas, self.eps, self.wd, self.t = lr, betas, eps, weight_decay, 0
        self.device = dev

    def params(self):
        """name -> view of the flat buffer (pass these to the graphs: the update is then visible to the next step without copies)"""
        return self.views

    def step(self, grads):
        self.t += 1
        scale = gather_and_reduce(grads, self.names, self.grad)       # data parallel: ONE all-reduce of the flat gradient buffer per optimizer
        L = self.L
        L.check(L.lib().ttts_adamw_step(self.flat.data_ptr(), self.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), None, self.flat.numel(),
                                        None, 0.0, float(scale), float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                        float(self.wd), int(self.t), L.stream_ptr().value), "ttts_adamw_step")


class TrainStep:
    """One optimisation step of ttts/vqvae/train.py:330-406 in the reference's order: synthesis -> discriminator loss on (y, y_hat.detach())
    -> optim_d.step() -> adversarial + feature losses through the UPDATED discriminators -> optim_g.step().  Pass the EuclideanCodebook module as
    `codebook` to have the quantizer's EMA update (core_vq.py:217-228) inside the step; posterior noises and segment starts are inputs."""

    def __init__(self, K, params_g, params_d, lr=1e-4, optimizer=None):
        """`optimizer`: class with (params, lr) -> .params() / .step(grads); the product default is FlatAdamW (CUDA).  Tests pass a torch one
        to check the ORDER of the step against the reference on CPU."""
        self.K = K
        opt = optimizer if optimizer is not None else FlatAdamW
        self.opt_g, self.opt_d = opt(params_g, lr), opt(params_d, lr)

    def step(self, wav, spec, lengths, text, text_lengths, codebook, eps_p, eps_q, ids_slice, segment_frames, wav_aug=None, spec_aug=None):
        Pg, Pd = self.opt_g.params(), self.opt_d.params()
        gen = GeneratorStep(self.K, Pg, Pd)
        out = gen.synthesize(wav, spec, lengths, text, text_lengths, codebook, eps_p, eps_q, ids_slice, segment_frames, wav_aug, spec_aug)
        dgraph = DiscriminatorGraph(self.K, Pd)
        real, _ = dgraph.forward(out["y_seg"])
        fake, _ = dgraph.forward(out["y_hat"].v)                      # .detach(): a constant for the discriminator step
        loss_d = dgraph.discriminator_loss(real, fake)
        self.opt_d.step(dgraph.backward(loss_d))
        # the generator graph read the discriminator weights as views of the flat buffer: rebuild its discriminator leaves after the update
        gen.disc = DiscriminatorGraph(self.K, Pd, gen.tape)
        out = gen.adversarial()
        self.opt_g.step(gen.backward())
        return dict(loss_disc=loss_d.v, **{k: out[k].v for k in ("loss_gen", "loss_fm", "loss_mel", "kl_ssl", "loss_kl", "total")})
