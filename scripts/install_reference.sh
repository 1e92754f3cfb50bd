#!/bin/bash
# Puts the UNMODIFIED reference (package + its own tests/) under baseline/_ref (git-ignored, but shipped to the GPU box by
# gpurun) so that the GPU box can (a) run the reference's own acceptance suite against the CUDA handle
# (tests/test_gpu_zzz_reference_suite.py) and (b) time the reference's Python calculators on the bench host (bench.py).
# Build container only: needs /root/reference.  The un-vendored dependency `guancodes` stays the stand-in of oracle/refshim/.
set -e
cd "$(dirname "$0")/.."
REF=${THEBOSS_REFERENCE_SRC:-/root/reference}
rm -rf baseline/_ref /tmp/theboss_ref_copy
cp -r "$REF" /tmp/theboss_ref_copy            # the build writes into the source tree; /root/reference is read-only
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/theboss_ref_copy
cp -r "$REF/tests" baseline/_ref/tests        # pip does not install the test modules
rm -rf /tmp/theboss_ref_copy
ls baseline/_ref
