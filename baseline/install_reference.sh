#!/bin/bash
# Installs the UNMODIFIED reference (drcut/CPD, /root/reference) into baseline/_ref (git-ignored).
# The reference ships no setup.py, so a packaging-only setup.py is added to a COPY under /tmp;
# no reference source file is changed.  Its examples (not a package) are copied verbatim.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC=/tmp/cpd_ref_src
rm -rf "$SRC" && cp -r /root/reference "$SRC"
cat > "$SRC/setup.py" <<'PY'
from setuptools import setup, find_packages
setup(name="CPDtorch", version="0.0.0+ddd9a67", packages=find_packages(include=["CPDtorch", "CPDtorch.*"]),
      package_data={"CPDtorch.quant": ["quant_cuda/*"]}, include_package_data=True, zip_safe=False)
PY
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$HERE/_ref" "$SRC" 2>&1 | tail -3
cp -r /root/reference/example "$HERE/_ref/example"
ls "$HERE/_ref" "$HERE/_ref/CPDtorch/quant/quant_cuda"
