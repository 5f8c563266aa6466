#!/bin/bash
# A/B of the quad kernel's generator grouping at the C2 shape (normal type)
for g in 4 2 1; do
  echo "groups=$g"
  GADM_QUAD_GEN_GROUPS=$g timeout 300 python tools/bench_projection.py --type normal --iters 3 --check 2 | cut -c1-420
done
