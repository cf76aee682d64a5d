#!/bin/bash
# Installs the UNMODIFIED reference (GEOS-ESM/GEOSmie) into the git-ignored baseline/_ref/ so that it travels to the GPU box with
# the snapshot (the box has no /root/reference).  Used only by `bench.py --impl reference`, bench's cpu_baseline leg and the
# live-reference tests (GEOSMIE_REFERENCE=baseline/_ref); nothing in geosmie_b200/ imports it.
#   1. pymiecoated (the one pip-installable part: src/pymiecoated/setup.py) -> baseline/_ref/site  (pip --target, offline)
#   2. the rest of the path is plain scripts, not a package: src/{geosmie,config,utils,gsf,pymiecoated} are copied verbatim to
#      baseline/_ref/src/ (same relative layout as the reference tree, so tests/refharness.py works on either root).
set -u
REF=${1:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
[ -d "$REF/src/pymiecoated/pymiecoated" ] || { echo "no reference tree at $REF: keeping the existing baseline/_ref"; exit 0; }
rm -rf "$HERE/_ref" && mkdir -p "$HERE/_ref/src"
TMP=$(mktemp -d)
cp -r "$REF/src/pymiecoated" "$TMP/pymiecoated"          # the build writes into the source tree; /root/reference is read-only
if python -m pip install -q --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref/site" "$TMP/pymiecoated" > "$HERE/_ref/pip.log" 2>&1; then
  echo "pip install pymiecoated -> baseline/_ref/site: ok"
else
  echo "pip install pymiecoated failed (see baseline/_ref/pip.log): using the verbatim copy under baseline/_ref/src/pymiecoated"
fi
rm -rf "$TMP"
for d in geosmie config utils gsf pymiecoated; do cp -r "$REF/src/$d" "$HERE/_ref/src/$d"; done
(cd "$REF" && find src -type f | sort | xargs sha1sum) > "$HERE/_ref/MANIFEST.sha1" 2>/dev/null
echo "reference copied to baseline/_ref/src ($(find "$HERE/_ref/src" -type f | wc -l) files)"
