#!/usr/bin/env bash
# oracle/build_ref_tests.sh -- TEST INFRASTRUCTURE.
# Compiles the reference's OWN test executables, unmodified and where they lie under
# /root/reference, against THIS repository's headers (include/) and libraries
# (libdvren.so + libdvren_hp.so): the drop-in acceptance test.  Binaries go to oracle/_ref/bin
# (git-ignored, they travel to the GPU box); no reference source is copied into the repo.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REPO="$(dirname "$HERE")"
REF="${REF:-/root/reference}"
CUDA="${CUDA:-/usr/local/cuda}"
OUT="$HERE/_ref/bin"
PKG="$REPO/diff-volume-renderer_b200"
mkdir -p "$OUT"
FLAGS=(-O2 -w -std=c++23 -DHP_WITH_CUDA=1 -I"$REPO/include" -I"$CUDA/include")
LIBS=(-L"$PKG" -ldvren -ldvren_hp -L"$CUDA/lib64" -lcudart_static -ldl -lrt -lpthread
      "-Wl,-rpath,\$ORIGIN/../../../diff-volume-renderer_b200")
build() {  # name, source, extra include dirs...
    local name="$1" src="$2"; shift 2
    if [[ ! -f "$OUT/$name" || "$src" -nt "$OUT/$name" || "$PKG/libdvren.so" -nt "$OUT/$name" ]]; then
        echo "[ref-tests] $name"
        g++ "${FLAGS[@]}" "$@" "$src" -o "$OUT/$name" "${LIBS[@]}"
    fi
}
build hp_runner              "$REF/hotpath/tests/runner/hp_runner.cpp"
build dvren_core_tests       "$REF/tests/core/test_core.cpp"
build dvren_smoke_forward    "$REF/tests/render/test_smoke_forward.cpp"           -I"$REF/tests/render"
build dvren_smoke_highres    "$REF/tests/render/test_smoke_forward_highres.cpp"   -I"$REF/tests/render"
build dvren_sdf_sphere       "$REF/tests/render/test_sdf_sphere.cpp"              -I"$REF/tests/render"
build dvren_smoke_animation  "$REF/tests/render/test_smoke_animation.cpp"         -I"$REF/tests/render"
echo "[ref-tests] done: $(ls "$OUT" | tr '\n' ' ')"
