#!/bin/bash
# A/B of two builds of the library on ONE box (boxes differ by a few per cent): tools/ab_ifit.sh [n] [d] [kind]
# expects rag-cobweb_b200/libcobweb_b200.base.so beside the current build
n=${1:-30000}; d=${2:-768}; kind=${3:-unit}
L=rag-cobweb_b200/libcobweb_b200.so
cp $L /tmp/new.so
for rep in 1 2; do
  echo "== new"; cp /tmp/new.so $L; python tools/ifit_phases.py $n $d $kind | head -2
  echo "== base"; cp rag-cobweb_b200/libcobweb_b200.base.so $L; python tools/ifit_phases.py $n $d $kind | head -2
done
cp /tmp/new.so $L
