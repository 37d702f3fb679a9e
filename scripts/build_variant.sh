#!/bin/sh
# Build a tuning variant of the library next to the product one:
#   scripts/build_variant.sh NAME "-DFLAG=1 ..."   ->  build/variants/libcvpack_b200_NAME.so
# Select it at run time with CVPACK_B200_LIB=build/variants/libcvpack_b200_NAME.so (never a fallback:
# the same sources with other compile-time constants).
set -e
name="$1"; shift
mkdir -p build/variants
make -j8 OBJ="build/obj_$name" LIB="build/variants/libcvpack_b200_$name.so" EXTRA="$*" "build/variants/libcvpack_b200_$name.so"
