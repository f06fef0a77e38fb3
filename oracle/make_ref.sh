#!/bin/sh
# oracle/make_ref.sh -- TEST INFRASTRUCTURE.  Oracle #0: the reference's OWN Fortran for the hot path, compiled from the
# sources where they lie under /root/reference into oracle/_ref/libref.so (git-ignored, travels to the GPU box).
#
# NOT RUNNABLE IN THIS IMAGE: there is no Fortran compiler, and the module as a whole needs MKL's
# mkl_rci.f90 (NonlinearOptimization.f90:15) which is not in the tree.  The hot path itself reaches neither MKL nor
# LinearAlgebra, so this recipe excerpts exactly the procedures the oracle restates
#     SteepestDescent 55-188, ConjugateGradient 193-394, LBFGS 398-625, the line searchers 1272-1699,
#     ConjugateGradient_basic 2249-2346
# into an MKL-free module with the reference's module name (so the symbols are the reference's:
# __nonlinearoptimization_MOD_lbfgs, ...), without copying any source into the repository.  __graft_entry__.build()
# calls it whenever `gfortran` is on PATH; tests/test_ref_pin.py then pins oracle.c against it bit for bit and the
# "PARITY UNPINNED" notices can go.
set -e
SRC=${1:-/root/reference/source/NonlinearOptimization.f90}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/_ref"
FC=${FC:-gfortran}
command -v "$FC" >/dev/null 2>&1 || { echo "make_ref.sh: no Fortran compiler ($FC) -- oracle #0 not built" >&2; exit 3; }
[ -r "$SRC" ] || { echo "make_ref.sh: $SRC not readable -- oracle #0 not built" >&2; exit 4; }
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
{
    echo "module NonlinearOptimization"
    echo "    implicit none"
    echo "contains"
    sed -n '55,188p;193,394p;398,625p;1272,1699p;2249,2346p' "$SRC"
    echo "end module NonlinearOptimization"
} > "$TMP/nlopt_excerpt.f90"
# the reference's own gfortran flags (makefile:12,33-36: -O3 -ffree-line-length-0 -fno-range-check; its -fopenmp finds no
# directive on this path); no fast-math, no -march: sums stay sequential and unfused
"$FC" -O3 -ffree-line-length-0 -fno-range-check -fPIC -shared -J "$TMP" -o "$OUT/libref.so" "$TMP/nlopt_excerpt.f90"
echo "make_ref.sh: built $OUT/libref.so from $SRC"
