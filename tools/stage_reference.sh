#!/bin/sh
# Stages the UNMODIFIED reference python package under baseline/_ref (git-ignored, NOT gpurun-ignored: it
# travels to the GPU box) for the reference arm of bench.py, tests/test_accelerate_gpu.py and the
# reference-initialised parity fixtures.  The reference's own pyproject fails setuptools' flat-layout discovery
# ("Multiple top-level packages: egs, runtime, zipvoice"), so the install runs from a pruned copy that holds only
# the `zipvoice` package and the project metadata; --no-deps because lhotse / vocos / pydub / tokenizer
# dependencies are not available offline (the model-level API does not import them).
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$SRC/zipvoice" "$SRC/pyproject.toml" "$SRC/README.md" "$SRC/LICENSE" "$TMP/"
rm -rf "$ROOT/baseline/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP"
rm -rf "$TMP"
diff -rq -x __pycache__ "$SRC/zipvoice" "$ROOT/baseline/_ref/zipvoice" && echo "baseline/_ref/zipvoice is identical to $SRC/zipvoice"
