#!/bin/bash
# Side libraries of ONE source file built with extra -D flags (CPU box), for A/B timing on the GPU box:
#   tools/variants.sh build <file.cu> name1:"-DX=1 -DY=2" name2:"..."   ->  <pkg>/libig_<name>.so
#   tools/variants.sh run "<command>" name1 name2 ...                    (command run once per variant + the shipped lib)
PKG=instageo-e2e-geospatial-ml_b200
cd "$(dirname "$0")/.."
if [ "$1" = build ]; then
  src=$2; shift 2
  python $PKG/build.py > /dev/null || exit 1
  base=$(basename $src .cu)
  for spec in "$@"; do
    name=${spec%%:*}; flags=${spec#*:}
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
      $flags -c $PKG/csrc/$src -o $PKG/build/${base}_var_$name.o || exit 1
    objs=$(ls $PKG/build/*.o | grep -v "_var_\|_abl\|_prof" | grep -v "/$base.o")
    nvcc -shared -o $PKG/libig_$name.so $objs $PKG/build/${base}_var_$name.o -gencode arch=compute_100a,code=sm_100a \
      -Xcompiler -fPIC -cudart static || exit 1
    echo "built libig_$name.so ($flags)"
  done
else
  cmd=$2; shift 2
  echo "== shipped"; eval "$cmd"
  for name in "$@"; do
    echo "== $name"; INSTAGEO_B200_LIB=$PWD/$PKG/libig_$name.so eval "$cmd"
  done
fi
