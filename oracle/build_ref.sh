#!/bin/bash
# oracle/build_ref.sh -- build the REAL reference (quantom-collab/tiktaalik: Fortran 90 + OpenMP core, f2py wrappers) into
# oracle/_ref/ when a Fortran compiler exists.  TEST INFRASTRUCTURE ONLY (see oracle/tko_fp.h); nothing under
# tiktaalik_b200/ uses its output.  No reference source is copied: the files are compiled where they lie.
#
# Status in this project's images (probed in round 2, profiles/r2_gpu_box_toolchain_probe.txt): neither the authoring
# container nor the GPU box has gfortran / f951 / flang / nvfortran / ifx, `import tiktaalik` fails and baseline/_ref does
# not exist, so this script exits 3 there and parity stays "unpinned" (DESIGN.md section 6).  On a machine WITH gfortran:
#     REF=/path/to/tiktaalik bash oracle/build_ref.sh && python tests/make_golden_ref.py
# builds the library and the three f2py modules exactly as the reference's own CMake files do
# (f90src/CMakeLists.txt:11-24: -Ofast -march=native -fopenmp -mcmodel=large; tiktaalik/f90wrap/CMakeLists.txt:2-27:
# one f2py module per wrapper linked against libtiktaalik), then dumps golden vectors and picks the -Ofast variant.
set -euo pipefail
REF=${REF:-/root/reference}
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
FC=${FC:-gfortran}
if ! command -v "$FC" >/dev/null 2>&1; then
    echo "build_ref.sh: no Fortran compiler ($FC) on this machine: the reference cannot be built here" >&2
    exit 3
fi
python -c "import numpy.f2py" 2>/dev/null || { echo "build_ref.sh: numpy.f2py is missing" >&2; exit 3; }
mkdir -p "$OUT/tiktaalik/f90wrap" "$OUT/build"
FFLAGS="-Ofast -march=native -fopenmp -mcmodel=large -fPIC -J $OUT/build"
# module dependency order of f90src (qcd constants -> helpers -> kernels -> evolution -> model)
SRC="qcd/constants.f90 qcd/alpha_qcd.f90 helpers/gridspace.f90 helpers/specfun.f90 helpers/pixelation.f90 helpers/integration.f90
     kernels/kernels_common.f90 kernels/kernels_lo.f90 kernels/kernels_nlo.f90 evolution/convolution.f90 evolution/matevo.f90 model/gk.f90"
OBJS=""
for f in $SRC; do
    o="$OUT/build/$(basename "${f%.f90}").o"
    "$FC" $FFLAGS -c "$REF/f90src/$f" -o "$o"
    OBJS="$OBJS $o"
done
"$FC" -shared -fopenmp -o "$OUT/tiktaalik/libtiktaalik.so" $OBJS
# the f2py wrapper modules (module `dummy` in each: tiktaalik/f90wrap/{matrices,qcd,model}.f90)
for w in matrices qcd model; do
    # tiktaalik/f90wrap/CMakeLists.txt:8: f2py -m <name> -c <file> -I<include> -L<lib> -ltiktaalik (no extra flags)
    ( cd "$OUT/build" && python -m numpy.f2py -m "$w" -c "$REF/tiktaalik/f90wrap/$w.f90" -I"$OUT/build" \
        -L"$OUT/tiktaalik" -ltiktaalik >/dev/null )
    cp "$OUT"/build/$w*.so "$OUT/tiktaalik/f90wrap/"
done
# the Python layer stays in the reference tree: the package directory here only links to it (no source is copied)
for f in "$REF"/tiktaalik/*.py; do ln -sf "$f" "$OUT/tiktaalik/$(basename "$f")"; done
[ -f "$REF/tiktaalik/f90wrap/__init__.py" ] && ln -sf "$REF/tiktaalik/f90wrap/__init__.py" "$OUT/tiktaalik/f90wrap/__init__.py" || touch "$OUT/tiktaalik/f90wrap/__init__.py"
cat > "$OUT/README.txt" <<EOT
Built by oracle/build_ref.sh from $REF with $("$FC" --version | head -1).
Use:  LD_LIBRARY_PATH=$OUT/tiktaalik python tests/make_golden_ref.py
EOT
echo "build_ref.sh: reference built into $OUT"
