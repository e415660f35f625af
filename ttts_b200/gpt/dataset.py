"""The caller side of the GPT train step: jsonl manifest -> (text ids, VQ codes, wav length) -> padded batch dict.

Mirrors `ttts/gpt/dataset.py` (`GptTtsDataset` :29-62, `GptTtsCollater` :65-98, `read_jsonl` / `write_jsonl` :16-27) so that
`Trainer(cfg_path)` works against the reference's on-disk data unchanged: manifest lines `{"path": ..., "text": ...}`, codes in
`<path>.vq.pth` (`torch.save(list[int])`, written by `ttts_b200.prepare.extract_vq`), text -> pinyin (TONE3, neutral tone as 5) -> BPE ids
with the reference's tokenizer file (`ttts/gpt/gpt_tts_tokenizer.json`, a data file of the reference checkout).

Differences, all host-side: (1) the reference decodes and resamples every clip to 24 kHz just to read its length (:52-54); the default
`wav_length_fn` here reads the header (`torchaudio.info`) and computes the resampled length `ceil(n * 24000 / sr)` -- the same number;
(2) `pypinyin` and the tokenizer are imported lazily and can be replaced (`text_fn`, `tokenizer`), so manifests that already hold
pinyin or token ids load without them; (3) under data parallelism the loader is sharded with a `DistributedSampler` (the reference leaves
that to `accelerator.prepare`, train.py:58).
"""
import json
import math
import re

import torch
import torch.nn.functional as F
import torch.utils.data

MAX_TEXT_TOKENS = 400          # dataset.py:55: longer samples are dropped
MAX_MEL_CODES = 600


def read_jsonl(path):
    with open(path, "r") as f:
        return [json.loads(line) for line in f.read().splitlines()]


def write_jsonl(path, all_paths):
    with open(path, "w", encoding="utf-8") as f:
        for item in all_paths:
            json.dump(item, f, ensure_ascii=False)
            f.write("\n")


_PUNCT = {"{": "(", "}": ")", "[": "(", "]": ")", "`": "'", "—": "-", "ʼ": "'"}
_PUNCT_RE = re.compile("|".join(re.escape(k) for k in sorted(_PUNCT, key=len, reverse=True)), flags=re.DOTALL)
_EXTRANEOUS_RE = re.compile(r"^[@#%_=\$\^&\*\+\\]$")


class BpeTextTokenizer:
    """`VoiceBpeTokenizer` (ttts/gpt/voice_tokenizer.py:32-58): punctuation normalisation, ' ' -> '[SPACE]', HF `tokenizers` BPE."""

    def __init__(self, vocab_file="ttts/gpt/gpt_tts_tokenizer.json"):
        from tokenizers import Tokenizer
        self.tokenizer = Tokenizer.from_file(vocab_file)

    @staticmethod
    def preprocess_text(txt):
        txt = _PUNCT_RE.sub(lambda m: _PUNCT[m.group(0)], txt)
        return _EXTRANEOUS_RE.sub("", txt)

    def encode(self, txt):
        return self.tokenizer.encode(self.preprocess_text(txt).replace(" ", "[SPACE]")).ids

    def decode(self, seq):
        if isinstance(seq, torch.Tensor):
            seq = seq.cpu().numpy()
        txt = self.tokenizer.decode(seq, skip_special_tokens=False).replace(" ", "")
        return txt.replace("[SPACE]", " ").replace("[STOP]", "").replace("[UNK]", "")


def to_pinyin(text):
    """dataset.py:42: ' '.join(lazy_pinyin(text, style=Style.TONE3, neutral_tone_with_five=True))"""
    try:
        from pypinyin import Style, lazy_pinyin
    except ImportError as e:                      # pragma: no cover - depends on the environment
        raise ImportError("pypinyin is needed to turn Chinese text into pinyin; pass text_fn= to GptTtsDataset for pre-romanised manifests") from e
    return " ".join(lazy_pinyin(text, style=Style.TONE3, neutral_tone_with_five=True))


def resampled_length(path, rate=24000):
    """length of torchaudio.functional.resample(wav, sr, rate) without decoding the file: ceil(rate * n / sr)"""
    import torchaudio
    info = torchaudio.info(path)
    return int(math.ceil(rate * info.num_frames / info.sample_rate))


class GptTtsDataset(torch.utils.data.Dataset):
    def __init__(self, opt, tokenizer=None, text_fn=to_pinyin, wav_length_fn=resampled_length):
        self.tok = tokenizer if tokenizer is not None else BpeTextTokenizer()
        self.text_fn = text_fn
        self.wav_length_fn = wav_length_fn
        self.jsonl_path = opt["dataset"]["path"]
        self.audiopaths_and_text = read_jsonl(self.jsonl_path)

    def __getitem__(self, index):
        try:
            item = self.audiopaths_and_text[index]
            audiopath, text = item["path"], item["text"]
            text = torch.LongTensor(self.tok.encode(self.text_fn(text) if self.text_fn is not None else text))
            qmel = torch.LongTensor(torch.load(audiopath + ".vq.pth"))
        except Exception as e:                    # the reference prints and drops the sample (dataset.py:49-51)
            print(e)
            return None
        wav_length = self.wav_length_fn(audiopath)
        if text.shape[0] > MAX_TEXT_TOKENS or qmel.shape[0] > MAX_MEL_CODES:
            return None
        return text, qmel, wav_length

    def __len__(self):
        return len(self.audiopaths_and_text)


class GptTtsCollater:
    """dataset.py:65-98: drop None samples (None batch if nothing is left), right-pad text and codes with 0, lengths as LongTensors."""

    def __init__(self, cfg):
        self.cfg = cfg

    def __call__(self, batch):
        batch = [x for x in batch if x is not None]
        if len(batch) == 0:
            return None
        text_lens = [len(x[0]) for x in batch]
        qmel_lens = [len(x[1]) for x in batch]
        max_text_len, max_qmel_len = max(text_lens), max(qmel_lens)
        texts = [F.pad(x[0], (0, max_text_len - len(x[0])), value=0) for x in batch]
        qmels = [F.pad(x[1], (0, max_qmel_len - len(x[1])), value=0) for x in batch]
        return {
            "padded_text": torch.stack(texts),
            "text_lengths": torch.LongTensor(text_lens),
            "padded_qmel": torch.stack(qmels),
            "qmel_lengths": torch.LongTensor(qmel_lens),
            "wav_lens": torch.LongTensor([x[2] for x in batch]),
        }


def build_dataloader(cfg, rank=0, world=1, dataset=None):
    """DataLoader(dataset, **cfg['dataloader'], collate_fn=GptTtsCollater(cfg)) (train.py:46-47), sharded across ranks when world > 1."""
    ds = dataset if dataset is not None else GptTtsDataset(cfg)
    kw = dict(cfg.get("dataloader", {}))
    if world > 1:
        sampler = torch.utils.data.distributed.DistributedSampler(ds, num_replicas=world, rank=rank, shuffle=bool(kw.pop("shuffle", False)),
                                                                  drop_last=bool(kw.get("drop_last", False)))
        kw["sampler"] = sampler
    return torch.utils.data.DataLoader(ds, collate_fn=GptTtsCollater(cfg), **kw)
