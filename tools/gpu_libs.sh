#!/bin/bash
# A/B of library builds: tools/gpu_libs.sh "<gpu_quick args>" lib1.so lib2.so ...   (each run in its own process, FERMI_PT_B200_LIB)
args="$1"; shift
for lib in "$@"; do
  echo "== $lib"
  FERMI_PT_B200_LIB=$PWD/$lib python tools/gpu_quick.py $args 2>&1 | grep '^{"o"'
done
