#!/bin/bash
# perf_flags.sh <tag> <cfg:variant> <flags...>: the same bench line under different AB2_DEBUG_FLAGS
tag=$1; cv=$2; shift; shift
for f in "$@"; do
  echo "== AB2_DEBUG_FLAGS=$f"
  AB2_DEBUG_FLAGS=$f bash tools/gpu/perf_sweep.sh ${tag}_f$f $cv
done
