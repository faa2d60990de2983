#!/bin/bash
# Installs the UNMODIFIED reference (graphtools v2.1.0) into baseline/_ref for the CPU arm of bench.py.
# /root/reference is read-only and setuptools writes egg-info into the source tree, so the install runs from
# a copy under /tmp.  --no-deps: the reference's `tasklogger`, `pygsp`, `future` dependencies are not in the
# offline wheelhouse; bench.py puts the three stand-ins of oracle/shims in front of sys.path instead (everything
# else -- numpy, scipy, scikit-learn, numba -- is the image's own).  baseline/_ref is git-ignored but travels to
# the GPU box with the gpurun snapshot.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${GT_REF_PATH:-/root/reference}"
[ -d "$SRC/graphtools" ] || { echo "reference not found at $SRC (nothing to do)"; exit 0; }
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
diff -rq -x __pycache__ "$HERE/_ref/graphtools" "$SRC/graphtools" && echo "baseline/_ref == $SRC/graphtools (unmodified)"
