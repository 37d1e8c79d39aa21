"""In-tree build of libzipvoice_b200.so (nvcc, sm_100a only).

Freshness is decided by CONTENT, not by mtime: the sha256 of the sources is compiled into the library
(`ZVB_SRC_HASH=<hex>` marker, also returned by `zvb_source_hash()`); `build()` re-compiles whenever the
marker found in the existing binary differs from the hash of the sources on disk, so a shipped `.so` that is
newer than edited sources is never reused."""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "engine.cu")
OUT = os.path.join(HERE, "libzipvoice_b200.so")
DEPS = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))
              if f.endswith((".cu", ".cuh", ".h")))
DEPS.append(os.path.join(os.path.dirname(HERE), "include", "zipvoice_b200.h"))
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC"]


def nvcc_path() -> str:
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc"):
        if p and os.path.exists(p):
            return p
    return "nvcc"


def source_hash() -> str:
    h = hashlib.sha256()
    h.update(" ".join(FLAGS).encode())
    for d in DEPS:
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    return h.hexdigest()


def built_hash(path: str = OUT):
    """The source hash compiled into an existing library (None if absent / unmarked)."""
    if not os.path.exists(path):
        return None
    m = re.search(rb"ZVB_SRC_HASH=([0-9a-f]{64})", open(path, "rb").read())
    return m.group(1).decode() if m else None


def is_fresh() -> bool:
    return built_hash() == source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    want = source_hash()
    if not force and built_hash() == want:
        return OUT
    cmd = [nvcc_path()] + FLAGS + [f'-DZVB_SOURCE_HASH="{want}"', "-o", OUT, SRC]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libzipvoice_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    assert built_hash() == want, "source-hash marker missing from the built library"
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
