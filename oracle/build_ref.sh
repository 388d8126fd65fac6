#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED-in-behaviour reference (EMS-TU-Ilmenau/fastmat,
# read-only at /root/reference) into oracle/_ref/ so that tests/ and bench.py's CPU-baseline legs can
# import the real thing next to the numpy restatement in oracle/fastmat_oracle.py.
#
# The reference is Cython; it is cythonized + compiled from a scratch copy under /tmp (the source tree is
# read-only and needs a two-line numpy>=2 compatibility rename).  Only BUILD OUTPUTS
# (the compiled extension modules and the package's runtime .py files, exactly what `pip install --target`
# would lay down) land in oracle/_ref/, which is git-ignored: no reference source enters the history.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${FASTMAT_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
SCRATCH="${TMPDIR:-/tmp}/fastmat_ref_build.$$"
if [ ! -d "$REF/fastmat" ]; then echo "reference not present at $REF - nothing to build"; exit 0; fi
if [ -f "$OUT/fastmat/__init__.py" ] && ls "$OUT"/fastmat/*.so >/dev/null 2>&1 && [ -z "${FORCE:-}" ]; then
  echo "oracle/_ref already built"; exit 0; fi
rm -rf "$SCRATCH"; cp -r "$REF" "$SCRATCH"; chmod -R u+w "$SCRATCH"; cd "$SCRATCH"
rm -f "fastmat/Matrix.pyx,cover"
# numpy >= 2 renamed these C-API names; behaviour is unchanged
sed -i 's/np\.NPY_NTYPES\b/np.NPY_NTYPES_LEGACY/g' fastmat/core/types.pxd fastmat/core/types.pyx
sed -i -E 's/np\.NPY_(F_CONTIGUOUS|C_CONTIGUOUS|OWNDATA|ENSUREARRAY|ENSURECOPY)\b/np.NPY_ARRAY_\1/g' fastmat/core/cmath.pyx
python setup.py build_ext --inplace > "$SCRATCH/build.log" 2>&1 || { tail -40 "$SCRATCH/build.log"; exit 1; }
rm -rf "$OUT"; mkdir -p "$OUT"
# install = compiled modules + runtime python files (no .pyx/.pxd/.c sources)
(cd "$SCRATCH" && find fastmat -type f \( -name '*.so' -o -name '*.py' \) -print0 | xargs -0 -I{} cp --parents {} "$OUT/")
PYTHONPATH="$OUT" python -c "import fastmat, numpy as np; F=fastmat.Fourier(8); print('oracle/_ref ok: fastmat', fastmat.__version__, np.abs(F.forward(np.ones(8))).max())"
rm -rf "$SCRATCH"
