#!/usr/bin/env bash
# Build the UNMODIFIED reference (seung-lab/fastremap) from the sources where they lie
# under /root/reference into oracle/_ref/ (git-ignored, travels to the GPU box).
#
# TEST / BASELINE INFRASTRUCTURE ONLY: nothing under fastremap_b200/ may import this.
#
# No reference source is copied: cython reads /root/reference/fastremap/fastremap.pyx
# (+ fastremap.pxd) and writes its generated C++ into oracle/_ref/build/; g++ includes the
# vendored headers (ska_flat_hash_map.hpp, ipt.hpp) straight from /root/reference/fastremap.
# Same compiler flags as the reference's setup.py:17-19 (-std=c++11 -O3).
# The only file written by this recipe that is not compiler output is the two-line package
# __init__ that re-exports the compiled module (the reference's own __init__.py:1-24 does the
# same with an explicit name list).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${FASTREMAP_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"

if [ ! -f "$REF/fastremap/fastremap.pyx" ]; then
  echo "build_ref.sh: reference sources not found at $REF (fine on the GPU box: prebuilt oracle/_ref is used)" >&2
  exit 0
fi

EXT="$($PY -c 'import sysconfig; print(sysconfig.get_config_var("EXT_SUFFIX"))')"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
NPINC="$($PY -c 'import numpy; print(numpy.get_include())')"
SO="$OUT/fastremap/fastremap$EXT"

if [ -f "$SO" ] && [ "$SO" -nt "$REF/fastremap/fastremap.pyx" ] && [ "${1:-}" != "--force" ]; then
  echo "build_ref.sh: $SO up to date"
  exit 0
fi

mkdir -p "$OUT/build" "$OUT/fastremap"
$PY -m cython --cplus -3 -I "$REF/fastremap" -o "$OUT/build/fastremap.cpp" "$REF/fastremap/fastremap.pyx"
g++ -std=c++11 -O3 -shared -fPIC -w \
    -I"$REF/fastremap" -I"$NPINC" -I"$PYINC" \
    "$OUT/build/fastremap.cpp" -o "$SO"
printf 'from .fastremap import *  # noqa: F401,F403  (compiled, unmodified reference)\nfrom . import fastremap  # noqa: F401\n' > "$OUT/fastremap/__init__.py"
rm -rf "$OUT/build"
echo "build_ref.sh: built $SO"
