#!/bin/bash
# usage: tools/build_variant.sh <name> <nvcc defines...>  — libtsb_<name>.so with the k_merge_bits_f translation units rebuilt under the defines
set -e
name=$1; shift
cd "$(dirname "$0")/../tensorsat_b200/csrc"
objs=""
for f in *.cu; do
  b=${f%.cu}
  case $b in
    tsb_merge_bits_g32|tsb_merge_bits_g42|tsb_merge_bits_gx1|tsb_merge_bits_h32|tsb_merge_bits_h42|tsb_merge_bits_i42|tsb_merge_bits_f32|tsb_merge_bits_f42|tsb_merge_bits_f)
      nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c $f -o /tmp/${b}_$name.o &
      objs="$objs /tmp/${b}_$name.o" ;;
    *) objs="$objs $b.o" ;;
  esac
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o libtsb_$name.so $objs -lcuda
echo built libtsb_$name.so
