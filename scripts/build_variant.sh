#!/bin/bash
# Build an A/B variant of the library from an alternative head.cu:  scripts/build_variant.sh <head_variant.cu> <name> [extra nvcc flags]
# -> variants/lib<name>.so (git-ignored); select it with SIMT_B200_LIB=variants/lib<name>.so.  The variant replaces head.cu AND
# head_ident.cu (a whole-file head.cu of the round-1 layout); it must export every symbol simt_b200/_lib.py binds.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$1; NAME=$2; shift 2
mkdir -p "$ROOT/variants" "$ROOT/build/obj"
cp "$SRC" "$ROOT/simt_b200/csrc/_variant_$NAME.cu"
trap 'rm -f "$ROOT/simt_b200/csrc/_variant_$NAME.cu"' EXIT
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" \
     -c "$ROOT/simt_b200/csrc/_variant_$NAME.cu" -o "$ROOT/build/obj/head_$NAME.o"
OBJS=""
for f in "$ROOT"/simt_b200/csrc/*.cu; do b=$(basename "$f" .cu); case "$b" in head|head_ident|_variant_*) ;; *) OBJS="$OBJS $ROOT/build/obj/$b.o";; esac; done
nvcc -shared -o "$ROOT/variants/lib$NAME.so" "$ROOT/build/obj/head_$NAME.o" $OBJS \
     -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
echo "$ROOT/variants/lib$NAME.so"
