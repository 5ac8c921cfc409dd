#!/bin/sh
# A/B builds: tools/build_variant.sh NAME file.cu "-DFOO=1 ..."  ->  min_llm_inference_b200/libmli_b200_NAME.so
# (the named source recompiled with the extra defines, every other object taken from the normal build;
# select it with MLI_B200_LIB=<path>).  Diagnostic helper, not part of the product build.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
CS=$ROOT/min_llm_inference_b200/csrc
NAME=$1; SRC=$2; DEFS=$3
make -C "$CS" -j8 >/dev/null
mkdir -p "$CS/_obj/var_$NAME"
OBJS=""
for o in "$CS"/_obj/*.o; do
  b=$(basename "$o" .o)
  if [ "$b.cu" = "$SRC" ]; then
    /usr/local/cuda/bin/nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC \
      -I"$ROOT/include" -I"$CS" --expt-relaxed-constexpr $DEFS -c "$CS/$SRC" -o "$CS/_obj/var_$NAME/$b.o"
    OBJS="$OBJS $CS/_obj/var_$NAME/$b.o"
  else
    OBJS="$OBJS $o"
  fi
done
/usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o "$ROOT/min_llm_inference_b200/libmli_b200_$NAME.so" $OBJS
echo "$ROOT/min_llm_inference_b200/libmli_b200_$NAME.so"
