#!/bin/bash
# tools/sweep.sh OUT ENV1 ENV2 ... : run the kernel-only bench once per environment setting
out=$1; shift
: > "$out"
for envs in "$@"; do
  echo "## $envs" >> "$out"
  env $envs timeout 300 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 >> "$out" 2>&1
done
