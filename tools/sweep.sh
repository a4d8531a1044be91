#!/bin/bash
# Developer sweep (GPU box): kernel build variants.
cd "$(dirname "$0")/.."
for lib in "" as_cops_and_thieves_b200/variants/libcat_w8_c4.so as_cops_and_thieves_b200/variants/libcat_w16_c2.so as_cops_and_thieves_b200/variants/libcat_w32_c1.so; do
  echo "=== lib=${lib:-default}"
  export CAT_B200_LIB=$lib
  [ -z "$lib" ] && unset CAT_B200_LIB
  python tools/prof_step.py --map squarinth --worlds 4096
  python tools/prof_step.py --map squarinth --worlds 16384
  python tools/prof_step.py --map agh-map --free 1 --worlds 16384
  python tools/prof_step.py --map labyrinth --free 1 --worlds 8192
done
