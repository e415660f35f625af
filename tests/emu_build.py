"""TEST INFRASTRUCTURE: one place that builds the host emulation libraries (tests/emu/*_emu.cpp = the CUDA sources of ttts_b200/csrc compiled by
g++ on cuda_emu.h).  Builds are cached under build/emu_cache by a hash of everything that goes into them (compiler flags, every file of
ttts_b200/csrc, include/ and tests/emu), so a test session compiles each library at most once and a second session none at all."""
import hashlib
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_CACHE = os.path.join(ROOT, "build", "emu_cache")
_digest = None


def _inputs_digest():
    global _digest
    if _digest is None:
        h = hashlib.sha256()
        for d in (os.path.join(ROOT, "ttts_b200", "csrc"), os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "emu")):
            for name in sorted(os.listdir(d)):
                path = os.path.join(d, name)
                if os.path.isfile(path):
                    h.update(name.encode())
                    with open(path, "rb") as f:
                        h.update(f.read())
        _digest = h.hexdigest()
    return _digest


def compile_emu(src, so):
    """g++ -shared build of tests/emu/<src> into <so>.  TTTS_EMU_CXXFLAGS="-g -fsanitize=address" (with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
    ASAN_OPTIONS=detect_leaks=0) turns every out-of-bounds shared / global access of the emulated kernels into a hard error."""
    flags = ["-O1", "-std=c++20", "-pthread", "-shared", "-fPIC"] + os.environ.get("TTTS_EMU_CXXFLAGS", "").split()
    key = hashlib.sha256((_inputs_digest() + "|" + src + "|" + " ".join(flags)).encode()).hexdigest()[:24]
    cached = os.path.join(_CACHE, "%s_%s.so" % (os.path.splitext(src)[0], key))
    if not os.path.exists(cached):
        os.makedirs(_CACHE, exist_ok=True)
        tmp = cached + ".%d.tmp" % os.getpid()
        cmd = ["g++"] + flags + ["-x", "c++", "-I", os.path.join(ROOT, "tests", "emu"), os.path.join(ROOT, "tests", "emu", src), "-o", tmp]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        os.replace(tmp, cached)
    shutil.copyfile(cached, so)
    return so
