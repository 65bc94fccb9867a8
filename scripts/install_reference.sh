#!/usr/bin/env bash
# Copies the reference's own Python sources for the hot path into baseline/_ref (git-ignored, NOT gpurun-ignored), so
# that `bench.py --impl reference` on the GPU box runs the reference's own modules (cpu_baseline.kind = "reference")
# instead of the oracle port.  Run in the build container, where /root/reference exists:
#     bash scripts/install_reference.sh
# The reference has no setup.py / pyproject.toml, so `pip install --target baseline/_ref` does not
# apply; its third-party dependency Uni-Core is absent and is resolved through oracle/shims at import time
# (oracle/ref_loader.py).  Nothing under baseline/_ref is imported by the product path, the -m gpu tests or smoke().
set -euo pipefail
SRC="${1:-/root/reference}"
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
DST="$ROOT/baseline/_ref"
[ -d "$SRC/models" ] || { echo "no reference tree at $SRC" >&2; exit 1; }
rm -rf "$DST"
mkdir -p "$DST"
for d in models utils config; do
  [ -d "$SRC/$d" ] && cp -r "$SRC/$d" "$DST/$d"
done
find "$DST" -name '__pycache__' -type d -prune -exec rm -rf {} +
echo "installed $(find "$DST" -name '*.py' | wc -l) reference files into $DST"
