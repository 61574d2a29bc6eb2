#!/bin/sh
# A/B of band-LU build variants (tools/build_variant.sh): stage times at C3 for the product library and every variant
for lib in spike_petsc_b200/lib/libspike_b200.so build/var/*/libspike_b200.so; do
  echo "== $lib"
  SPIKE_B200_LIB=$lib python tools/config_sweep.py 10000000,100,296,78 "$@"
done
