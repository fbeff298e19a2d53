#!/bin/bash
# Test / baseline infrastructure: copy the UNMODIFIED reference modules of the hot path into oracle/_ref/ (git-ignored,
# but shipped to the GPU box by gpurun) so that `bench.py --impl reference` can time the reference itself on the box's
# host cores instead of the oracle port.  Run in the authoring container only (needs $TORCH_CFD_REF, default
# /root/reference); nothing under oracle/_ref/ is ever committed, imported by the package, or edited.
set -e
REF=${TORCH_CFD_REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
[ -d "$REF/torch_cfd" ] || { echo "make_ref.sh: $REF/torch_cfd not found (reference absent: the reference arm will use the oracle port)"; exit 0; }
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$REF/torch_cfd" "$HERE/_ref/torch_cfd"
mkdir -p "$HERE/_ref/fno"
cp "$REF"/fno/__init__.py "$REF"/fno/base.py "$REF"/fno/sfno.py "$REF"/fno/fno3d.py "$HERE/_ref/fno/" 2>/dev/null || true
find "$HERE/_ref" -name "__pycache__" -type d -exec rm -rf {} + 2>/dev/null || true
( cd "$REF" && git rev-parse HEAD 2>/dev/null || echo "unknown" ) > "$HERE/_ref/REVISION"
echo "oracle/_ref: $(find "$HERE/_ref" -name '*.py' | wc -l) reference files copied from $REF"
