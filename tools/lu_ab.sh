#!/bin/sh
# A/B of band-LU build variants (tools/build_variant.sh): stage times for the product library and every variant.
# usage: tools/lu_ab.sh [n,k,P,tip ...]   (default: C3)
cases="${@:-10000000,100,296,78}"
for lib in spike_petsc_b200/lib/libspike_b200.so build/var/*/libspike_b200.so; do
  echo "== $lib"
  SPIKE_B200_LIB=$lib python tools/config_sweep.py $cases
done
