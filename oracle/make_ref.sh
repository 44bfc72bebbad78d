#!/bin/sh
# Install the UNMODIFIED reference (pure Python, GoekeLab/xpore v2.1) where it can travel to the GPU box:
# a plain copy of its package directory into oracle/_ref/ (git-ignored, not gpurun-ignored), nothing edited.
# oracle/run_ref.py runs it with the two import shims of SURVEY.md 8c (ujson -> json, an empty h5py module:
# neither is installed in this image, neither touches the arithmetic of diffmod).
# Test / benchmark infrastructure only: bench.py --impl reference and bench.py's cli_e2e leg time it.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
[ -d "$SRC/xpore" ] || { echo "no reference at $SRC"; exit 1; }
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$SRC/xpore" "$HERE/_ref/xpore"
find "$HERE/_ref" -name "__pycache__" -type d -prune -exec rm -rf {} +
echo "reference copied to $HERE/_ref ($(find "$HERE/_ref" -name '*.py' | wc -l) python files)"
