#!/bin/sh
# A/B of wide-path build variants (tools/build_variant.sh) at C5; SPIKE_WS_NCT=1|2 forces 8 / 16 columns per sweep CTA
for lib in spike_petsc_b200/lib/libspike_b200.so build/var/*/libspike_b200.so; do
  for nct in 1 2; do
    echo "== $lib NCT=$nct"
    SPIKE_WS_NCT=$nct SPIKE_B200_LIB=$lib python tools/wide_perf.py 1000000,512,32,288,32 125000,512,16,320,32 "$@"
  done
done
