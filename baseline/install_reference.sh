#!/bin/sh
# Reference arm of bench.py (--impl reference): the UNMODIFIED reference package, importable on the GPU
# box (which has no /root/reference).  `pip install --target baseline/_ref /root/reference` fails here:
# the reference's build backend (hatchling) is not in the offline wheelhouse.  The package is pure
# Python (src/kinematics, dependencies numpy / scipy / pydantic / pyyaml are in the image), so the
# install step pip would perform -- copying the package directory -- is done directly.
# baseline/_ref is git-ignored (never committed) and travels to the GPU box with the gpurun snapshot.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref"
cp -r "$SRC/src/kinematics" "$HERE/_ref/kinematics"
find "$HERE/_ref" -name __pycache__ -type d -prune -exec rm -rf {} +
python - <<PY
import sys
sys.path.insert(0, "$HERE/_ref")
import kinematics
print("reference importable from", kinematics.__file__)
PY
