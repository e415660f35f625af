"""Locate and import the REAL reference (adelacvg/ttts) for the CPU arm of bench.py / the golden generator: `baseline/_ref/` (installed by
baseline/install_ref.py, travels to the GPU box) or `/root/reference` (build container).  The import shims are SURVEY.md Appendix D: three
`sys.modules` stubs for things the installed transformers 5.5 / absent librosa / encodec no longer provide -- nothing in the reference is edited."""
import importlib.machinery as im
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def find_reference():
    for p in (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if os.path.exists(os.path.join(p, "ttts", "gpt", "model.py")):
            return p
    return None


def import_reference(path=None):
    """Returns the module ttts.gpt.model of the reference (and leaves ttts.vqvae.* importable), or None when no reference tree is present."""
    path = path or find_reference()
    if path is None:
        return None
    if path not in sys.path:
        sys.path.insert(0, path)
    sys.dont_write_bytecode = True
    s = types.ModuleType("transformers.utils.model_parallel_utils")
    s.get_device_map = s.assert_device_map = lambda *a, **k: None
    sys.modules["transformers.utils.model_parallel_utils"] = s
    t = types.ModuleType("ttts.utils.typical_sampling")
    t.TypicalLogitsWarper = object
    sys.modules["ttts.utils.typical_sampling"] = t
    import ttts.gpt.model as gm  # noqa: must precede the librosa stub (transformers probes find_spec('librosa'))
    import torchaudio
    lb = types.ModuleType("librosa"); lb.__spec__ = im.ModuleSpec("librosa", None)
    lu = types.ModuleType("librosa.util"); lu.normalize = lu.pad_center = lu.tiny = None
    lf = types.ModuleType("librosa.filters")
    lf.mel = lambda sr, n_fft, n_mels, fmin, fmax: torchaudio.functional.melscale_fbanks(
        n_fft // 2 + 1, fmin, fmax or sr / 2, n_mels, sr, norm="slaney", mel_scale="slaney").T.numpy()
    lb.util, lb.filters = lu, lf
    sys.modules.update({"librosa": lb, "librosa.util": lu, "librosa.filters": lf})
    e = types.ModuleType("encodec"); e.EncodecModel = object; sys.modules["encodec"] = e
    import logging
    logging.disable(logging.CRITICAL)
    return gm
