#!/usr/bin/env bash
# Stage the UNMODIFIED reference under the git-ignored baseline/_ref/ (src/, scripts/, configs/ -- 0.3 MB).
# /root/reference does not exist on the GPU box; baseline/_ref is git-ignored but NOT gpurun-ignored, so it travels with the
# snapshot.  Used by: bench.py --impl reference (times the reference's own src.* on the host cores, kind "reference"),
# bench.py's cpu_baseline leg, and tests/test_gpu_reference_scripts.py (drives the reference's scripts/evaluate.py and Trainer
# through msd_b200.run).  Nothing here is product code and nothing is copied into tracked paths.
set -euo pipefail
SRC="${1:-/root/reference}"
DST="$(cd "$(dirname "$0")/.." && pwd)/baseline/_ref"
if [ ! -d "$SRC/src" ]; then
  echo "stage_ref: no reference tree at $SRC (nothing staged)"
  exit 0
fi
rm -rf "$DST"
mkdir -p "$DST"
cp -r "$SRC/src" "$SRC/scripts" "$SRC/configs" "$DST/"
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "stage_ref: staged $(find "$DST" -type f | wc -l) files from $SRC into $DST"
