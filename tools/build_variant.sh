#!/bin/bash
# Build a kernel variant of libsrm.so for A/B measurements: build/variants/libsrm_<name>.so = the default objects with the
# listed sources re-compiled with extra flags.  Usage: tools/build_variant.sh <name> "<nvcc -D flags>" [src.cu ...]
# (default source: srm_band.cu).  Select at run time with SRM_LIB=... or tools/ab_inproc.py <name>.
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2; shift 2 || true
srcs=${@:-srm_band.cu}
CS=surface-remesher_b200/csrc
make -C $CS -j4 >/dev/null
mkdir -p build/variants/obj_$name
objs=""
for s in srm_api.cu srm_label.cu srm_band.cu srm_jfa.cu srm_centroid.cu srm_lloyd.cu srm_raster.cu srm_recover.cu srm_host.cu; do
  if [[ " $srcs " == *" $s "* ]]; then
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xptxas -v $flags \
        -c $CS/$s -o build/variants/obj_$name/${s%.cu}.o 2> build/variants/obj_$name/${s%.cu}.ptxas.log
    grep -E "registers|spill" build/variants/obj_$name/${s%.cu}.ptxas.log | sort | uniq -c | head -4
    objs="$objs build/variants/obj_$name/${s%.cu}.o"
  else
    objs="$objs $CS/${s%.cu}.o"
  fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/libsrm_$name.so $objs -lcudart
echo "built build/variants/libsrm_$name.so"
