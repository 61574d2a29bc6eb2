#!/bin/sh
# Build an alternative libspike_b200.so with extra nvcc flags for A/B kernel experiments:
#   tools/build_variant.sh NAME "-DNS_XT_SHFL=1 ..."   ->  build/var/NAME/libspike_b200.so
# Select it with SPIKE_B200_LIB=build/var/NAME/libspike_b200.so (tools only; tests and bench.py use the product library).
set -e
name="$1"; flags="$2"
make -s LIBDIR="build/var/$name" NVEXTRA="$flags" "build/var/$name/libspike_b200.so"
echo "built build/var/$name/libspike_b200.so ($flags)"
