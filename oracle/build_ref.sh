#!/bin/bash
# Recipe for oracle/_ref: the reference's own nearest-neighbour engine.
# ann/ann_sample.cpp needs the ANN 1.1.2 headers and library, which the reference does not vendor, so it
# cannot be compiled here; the reference ships the executable it built from that file (ann/ann_sample,
# x86-64 ELF, libANN linked statically).  This script places that executable under oracle/_ref/ (git-ignored,
# shipped to the GPU box) so that tests can pin the oracle's nearest-neighbour search against the true engine.
# No reference SOURCE is copied.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${VPOWER_REFERENCE:-/root/reference}"
mkdir -p "$HERE/_ref"
if [ -x "$REF/ann/ann_sample" ]; then
  cp "$REF/ann/ann_sample" "$HERE/_ref/ann_sample"
  chmod +x "$HERE/_ref/ann_sample"
  echo "oracle/_ref/ann_sample installed"
else
  echo "reference tree not found at $REF; oracle/_ref not built" >&2
fi
