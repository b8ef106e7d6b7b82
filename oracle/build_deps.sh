#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Builds the two third-party libraries the reference links against, from the
# tarballs the reference ships under deps/ (SURVEY.md 8c recipe; no network needed, both carry a
# pre-generated configure):
#   NLopt 2.4.2    -> $PREFIX/lib/libnlopt.a      (consumer of the path: optimize_nlopt.cpp)
#   ADOL-C 2.6.3   -> $PREFIX/lib64/libadolc.so   (arithmetic ON the path: calc_*_function_gradient_adolc)
# Sources are unpacked and compiled under a scratch directory; only the installed libraries and headers
# land in oracle/_ref/deps (git-ignored, travels to the GPU box with the snapshot).
set -e
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
PREFIX=${1:-$HERE/_ref/deps}
WORK=${2:-${TMPDIR:-/tmp}/treepl_ref_deps}
mkdir -p "$PREFIX" "$WORK"
cd "$WORK"
# the image's $CC / $CXX wrappers link libstdc++ statically; use the system compilers
export CC=/usr/bin/gcc CXX=/usr/bin/g++
if [ ! -f "$PREFIX/lib/libnlopt.a" ]; then
  rm -rf nlopt-2.4.2
  tar xzf "$REF/deps/nlopt-2.4.2.tar.gz"
  (cd nlopt-2.4.2 && ./configure --prefix="$PREFIX" --with-pic --without-python --without-octave \
       --without-matlab --without-guile > configure.log 2>&1 && make -j8 install > make.log 2>&1)
fi
if [ ! -f "$PREFIX/lib64/libadolc.so" ]; then
  rm -rf ADOL-C-2.6.3
  tar xzf "$REF/deps/ADOL-C-2.6.3.tgz"
  # LIBRARY_PATH: libtool links with -nostdlib and otherwise cannot read libgomp.spec
  (cd ADOL-C-2.6.3 && ./configure --prefix="$PREFIX" --with-openmp-flag=-fopenmp --disable-sparse > configure.log 2>&1 \
       && LIBRARY_PATH=$(dirname "$(gcc -print-file-name=libgomp.spec)") make -j8 install > make.log 2>&1)
fi
echo "deps ok: $PREFIX"
