"""Install the UNMODIFIED reference (adelacvg/ttts, /root/reference) into baseline/_ref/ so that `bench.py --impl reference` can run the real
`ttts.gpt.model.UnifiedVoice` / `ttts.vqvae.vq2.SynthesizerTrn` on the GPU box's host cores.  baseline/_ref/ is
git-ignored (the reference's sources never enter this repo's history) but NOT gpurun-ignored, so it travels to the GPU box.

    python baseline/install_ref.py            # build container only: /root/reference does not exist on the GPU box

Recipe = the task's: `pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref <src>`.
Two facts about the reference's packaging (recorded in DESIGN.md):
  * /root/reference is read-only and `setup.py bdist_wheel` writes build/ + egg-info into the source tree -> install from a copy under /tmp;
  * its setup.py is `setup(packages=find_packages())` but `ttts/` and the sub-packages on this path (`gpt`, `vqvae`, `utils`, `vocoder`, ...)
    have NO `__init__.py` (the authors run from a checkout with PYTHONPATH), so the wheel built from the pristine tree is EMPTY (6 KB, only
    dist-info).  The copy under /tmp therefore gets empty `__init__.py` files in the package directories that lack one -- packaging only,
    no source line of the reference is touched -- and non-code payload (52 MB checkpoint, corpora, notebooks, wavs) is left out.
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")
TMP = "/tmp/ttts_ref_src"


def install():
    if not os.path.isdir(SRC):
        print("no %s here: nothing to install (the GPU box uses the prebuilt baseline/_ref)" % SRC)
        return os.path.isdir(os.path.join(DST, "ttts", "gpt"))
    shutil.rmtree(TMP, ignore_errors=True)
    skip = shutil.ignore_patterns("pretrained_models", "data", "*.ipynb", "*.wav", "*.png", "build", "*.egg-info", "__pycache__", "spider")
    shutil.copytree(SRC, TMP, ignore=skip)
    for d, subdirs, files in os.walk(os.path.join(TMP, "ttts")):
        if any(f.endswith(".py") for f in files) and "__init__.py" not in files:
            open(os.path.join(d, "__init__.py"), "w").close()
    shutil.rmtree(DST, ignore_errors=True)
    cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links", "/opt/wheelhouse",
           "--target", DST, TMP]
    r = subprocess.run(cmd, capture_output=True, text=True)
    print(r.stdout[-600:], r.stderr[-600:])
    # config files the reference reads next to its sources (vqvae/config.json is read at import of vqvae/train.py; gpt/config.json by Trainer)
    for rel in ("ttts/gpt/config.json", "ttts/vqvae/config.json"):
        if os.path.exists(os.path.join(SRC, rel)) and os.path.isdir(os.path.dirname(os.path.join(DST, rel))):
            shutil.copy(os.path.join(SRC, rel), os.path.join(DST, rel))
    ok = r.returncode == 0 and os.path.exists(os.path.join(DST, "ttts", "gpt", "model.py"))
    print("baseline/_ref:", "ok" if ok else "FAILED")
    return ok


if __name__ == "__main__":
    sys.exit(0 if install() else 1)
