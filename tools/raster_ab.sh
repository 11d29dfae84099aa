#!/bin/bash
# same-box A/B of raster builds: tools/raster_ab.sh libA.so libB.so ...  (TG_LIB_OVERRIDE), two interleaved rounds
for round in 1 2; do
  for lib in "$@"; do
    echo "== $lib (round $round)"
    if [ "$lib" = "default" ]; then python tools/raster_time.py 2>&1 | tail -3; else TG_LIB_OVERRIDE=$lib python tools/raster_time.py 2>&1 | tail -3; fi
  done
done
