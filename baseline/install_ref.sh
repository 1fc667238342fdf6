#!/bin/sh
# Installs the UNMODIFIED reference Python package (tskit + its _tskit extension) from
# $TSKIT_REFERENCE/python into baseline/_ref (git-ignored, travels to the GPU box).  Used only
# by tests of the drop-in seam (tskit_b200/dropin.py) as the host product and as the checker.
# /root/reference is read-only, so the build runs on a copy under /tmp.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${TSKIT_REFERENCE:-/root/reference}"
if [ ! -d "$REF/python" ]; then
    echo "install_ref.sh: $REF/python not found; keeping any existing baseline/_ref" >&2
    exit 0
fi
if [ -f "$HERE/_ref/tskit/__init__.py" ]; then exit 0; fi
TMP="$(mktemp -d /tmp/tskit_ref_build.XXXXXX)"
cp -rL "$REF/python" "$TMP/python" 2>/dev/null || true
cd "$TMP/python"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" . > "$TMP/install.log" 2>&1 || { tail -20 "$TMP/install.log"; exit 1; }
rm -rf "$TMP"
echo "installed reference tskit into $HERE/_ref"
