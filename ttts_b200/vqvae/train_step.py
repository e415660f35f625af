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

    def forward(self, wav, spec, lengths, text, text_lengths, codebook, eps_p, eps_q, ids_slice, segment_frames):
        """wav [B,L], spec [B,1025,T] (= spectrogram_torch(wav); wav_aug = wav), lengths [B] frames, text [B,Tt] int64, codebook [1024,192],
        eps_p / eps_q [B,192,T] posterior noises, ids_slice [B] segment starts (frames).  Returns the dict of loss Vars."""
        o, enc = self.ops, self.enc
        B, _, T = spec.shape
        dev = spec.device
        mask2 = (torch.arange(T, device=dev)[None, :] < lengths[:, None]).float().contiguous()
        specv, wavv = Var(spec.contiguous()), Var(wav.unsqueeze(1).contiguous())
        ge = enc.mel_style_encoder(Var(self.K.mul_mask(spec.contiguous(), mask2)), mask2, lengths)            # vq2.py:847
        x, _ = enc.posterior_audio_encoder(specv, wavv, mask2, ge, eps_p, "enc_p.")                           # :849 (the noisy sample feeds proj)
        x = o.conv(x, enc.P["proj.weight"], enc.P["proj.bias"], stride=2)                                     # :851
        quantized, commit, codes = o.vq(x, codebook)                                                          # :852-853
        q_up = o.upsample2(quantized)                                                                         # :854-856
        _, stats_p = self.te.forward(q_up, lengths, text, text_lengths, ge)                                   # :857
        m_p, logs_p = o.slice_c(stats_p, 0, 192), o.slice_c(stats_p, 192, 384)
        z, stats_q = enc.posterior_audio_encoder(specv, wavv, mask2, ge, eps_q, "enc_q.")                     # :858
        logs_q = o.slice_c(stats_q, 192, 384)
        z_p = self.flow.forward(z, mask2, ge)                                                                 # :859
        y_hat = self.dec.forward(o.slice_t(z, ids_slice, segment_frames), ge)                                 # :861-864
        # ---- losses (train.py:357-395) ----
        L = segment_frames * HOP
        y_mel = torch.stack([self.K.logmel_fwd(wav)[b, :, int(s):int(s) + segment_frames] for b, s in enumerate(ids_slice)]).contiguous()
        loss_mel = o.scale(o.l1_mean(y_mel, o.logmel(o.reshape(y_hat, (B, L)))), C_MEL)
        y_seg = torch.stack([wav[b, int(s) * HOP:int(s) * HOP + L] for b, s in enumerate(ids_slice)]).unsqueeze(1).contiguous()
        _, fmap_r = self.disc.forward(y_seg)
        gen, fmap_g = self.disc.forward(y_hat)
        loss_gen, loss_fm = self.disc.generator_losses(gen, fmap_r, fmap_g)
        loss_kl = o.scale(o.kl(z_p, logs_q, m_p, logs_p, mask2), C_KL)
        total = o.add(o.add(o.add(loss_gen, loss_fm), o.add(loss_mel, commit)), loss_kl)
        self.out = dict(loss_gen=loss_gen, loss_fm=loss_fm, loss_mel=loss_mel, kl_ssl=commit, loss_kl=loss_kl, total=total, y_hat=y_hat, z=z, codes=codes)
        return self.out

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
