#!/bin/sh
# Populate oracle/_ref/ with the UNMODIFIED reference modules of the hot path so that `bench.py --impl reference`
# and the cpu_baseline leg can time the reference ITSELF on the GPU box (where /root/reference does not exist).
#
# TEST / BASELINE INFRASTRUCTURE ONLY.  oracle/_ref/ is git-ignored (no reference source enters the history) but not
# gpurun-ignored, so it travels with the snapshot like the built .so files.  The modules are copied byte for byte
# (sha256 recorded in MANIFEST) and imported under oracle/ref_shim.py (torch 1.1 -> 2.x API stubs); nothing in the
# product path (asvspoof2021_air_b200/, main_train.py, generate_score.py) ever imports them.
#
#   sh oracle/build_ref.sh            (run by __graft_entry__.build() when the reference tree is mounted)
set -e
SRC="${AIR_REFERENCE_SRC:-/root/reference}"
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
FILES="feature_extraction.py utils_dsp.py resnet.py ecapa_tdnn.py loss.py"
if [ ! -f "$SRC/feature_extraction.py" ]; then
  echo "build_ref: no reference tree at $SRC (nothing to do)"; exit 0
fi
mkdir -p "$DST"
: > "$DST/MANIFEST"
for f in $FILES; do
  cp -f "$SRC/$f" "$DST/$f"
  chmod 644 "$DST/$f"
  (cd "$SRC" && sha256sum "$f") >> "$DST/MANIFEST"
done
(cd "$SRC" && git rev-parse HEAD 2>/dev/null || echo "no-git") > "$DST/COMMIT"
echo "build_ref: $(echo $FILES | wc -w) reference modules -> $DST"
