#!/usr/bin/env bash
# Install the UNMODIFIED reference next to this repo's product so that tests and bench.py can run it on
# the GPU box (where /root/reference does not exist):
#   * baseline/_ref/core/      <- /root/reference/core/        (pure Python: models, core/corr.py, utils)
#   * oracle/_ref/corr_sampler_ref.so  <- the reference's CUDA sampler, compiled from its own sources
#     (oracle/build_ref_sampler.py; 2-token torch-2.x compatibility patch, see there)
# Both directories are git-ignored (reference sources never enter the history) but NOT gpurun-ignored,
# so they travel to the GPU box with the snapshot, like our own built libsmoecorr.so.
#   tools/install_ref.sh [/path/to/reference]
set -euo pipefail
REF="${1:-/root/reference}"
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
if [ ! -d "$REF/core" ]; then
  echo "install_ref: $REF/core not found (nothing installed)" >&2
  exit 0
fi
mkdir -p "$ROOT/baseline/_ref"
rm -rf "$ROOT/baseline/_ref/core"
cp -r "$REF/core" "$ROOT/baseline/_ref/core"
find "$ROOT/baseline/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
(cd "$REF" && git rev-parse HEAD 2>/dev/null || echo "unknown") > "$ROOT/baseline/_ref/REFERENCE_COMMIT"
echo "install_ref: copied $REF/core -> baseline/_ref/core ($(find "$ROOT/baseline/_ref/core" -name '*.py' | wc -l) python files)"
if [ "${SMOE_SKIP_REF_SAMPLER:-0}" != "1" ]; then
  python "$ROOT/oracle/build_ref_sampler.py" --reference "$REF"
fi
