#!/bin/bash
# Copies the reference's hot-path modules (modules/Transformer/{model,layers}.py, pure Python) into the git-ignored
# baseline/_ref/ so that bench.py --impl reference and the cpu_baseline leg can time the UNMODIFIED reference on the GPU
# box (gpurun ships git-ignored files; /root/reference itself does not travel).  Nothing under baseline/_ref is tracked.
# usage: tools/install_ref.sh [reference root, default /root/reference]
set -e
REF=${1:-/root/reference}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
DST="$ROOT/baseline/_ref/modules/Transformer"
if [ ! -f "$REF/modules/Transformer/model.py" ]; then
  echo "install_ref: $REF/modules/Transformer/model.py not found (nothing installed)"; exit 0
fi
mkdir -p "$DST"
cp "$REF/modules/Transformer/model.py" "$REF/modules/Transformer/layers.py" "$DST/"
( cd "$REF" && sha256sum modules/Transformer/model.py modules/Transformer/layers.py ) > "$ROOT/baseline/_ref/SHA256SUMS"
echo "install_ref: reference hot-path modules installed under baseline/_ref (sha256 in baseline/_ref/SHA256SUMS)"
