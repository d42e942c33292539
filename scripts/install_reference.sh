#!/bin/bash
# Offline install of the UNMODIFIED reference into baseline/_ref (git-ignored; it still travels to
# the GPU box with the gpurun snapshot).  /root/reference is read-only, so the build runs from a
# copy under /tmp.  The reference's setup.py lists packages=['ipsolver'] only -- a plain pip install
# leaves out ipsolver._large_scale_constrained and the result does not import -- so the COPY's
# setup.py gets the sub-packages added to that list; no module of the reference is touched.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=${1:-/root/reference}
TMP=$(mktemp -d /tmp/ipsolver_ref.XXXXXX)
cp -r "$SRC"/. "$TMP"/
python - "$TMP/setup.py" <<'P'
import sys
p = sys.argv[1]
s = open(p).read()
new = ("packages=['ipsolver', 'ipsolver.tests', 'ipsolver._large_scale_constrained', "
       "'ipsolver._large_scale_constrained.tests']")
assert "packages=['ipsolver']" in s
open(p, "w").write(s.replace("packages=['ipsolver']", new))
P
rm -rf "$ROOT/baseline/_ref"
mkdir -p "$ROOT/baseline"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$ROOT/baseline/_ref" "$TMP"
rm -rf "$TMP"
PYTHONPATH="$ROOT/baseline/_ref" python -c "import ipsolver, ipsolver._large_scale_constrained.qp_subproblem as q; print('reference installed:', ipsolver.__file__)"
