#!/bin/bash
# A/B of two builds of the library on ONE box: tools/ab_run.sh <command...>; expects rag-cobweb_b200/libcobweb_b200.base.so
L=rag-cobweb_b200/libcobweb_b200.so
cp $L /tmp/new.so
for rep in 1 2; do
  echo "== new"; cp /tmp/new.so $L; "$@"
  echo "== base"; cp rag-cobweb_b200/libcobweb_b200.base.so $L; "$@"
done
cp /tmp/new.so $L
