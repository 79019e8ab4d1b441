#!/bin/bash
# Build attention.cu with -DATTN_ABLATE=n into side libraries (CPU box), or time them (GPU box: `run`).
PKG=instageo-e2e-geospatial-ml_b200
if [ "$1" = prof ]; then
  python $PKG/build.py > /dev/null
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
    -DATTN_PROFILE=${PROFMODE:-1} ${EXTRA} -c $PKG/csrc/attention.cu -o $PKG/build/attention_prof.o || exit 1
  objs=$(ls $PKG/build/*.o | grep -v attention)
  nvcc -shared -o $PKG/libig_prof.so $objs $PKG/build/attention_prof.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static || exit 1
  rm -f $PKG/libig_abl*.so
elif [ "$1" = trace ]; then
  INSTAGEO_B200_LIB=$PWD/$PKG/libig_prof.so python tools/attn_trace.py $2 $3 2>&1 | tee gpurun_out/attn_trace.txt
elif [ "$1" = profrun ]; then
  INSTAGEO_B200_LIB=$PWD/$PKG/libig_prof.so python tools/attn_profile.py 2>&1 | tee gpurun_out/attn_profile.txt
elif [ "$1" = build ]; then
  python $PKG/build.py > /dev/null
  for n in ${ABL:-1 2 3 4}; do
    nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
      -DATTN_ABLATE=${ABLV:-$n} -DIG_WAIT_MODE=${WAITMODE:-0} ${EXTRA} -c $PKG/csrc/attention.cu -o $PKG/build/attention_abl$n.o || exit 1
    objs=$(ls $PKG/build/*.o | grep -v attention)
    nvcc -shared -o $PKG/libig_abl$n.so $objs $PKG/build/attention_abl$n.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static || exit 1
  done
  ls -la $PKG/*.so
else
  mkdir -p gpurun_out
  for n in 0 ${ABL:-1 2 3 4}; do
    lib=""; [ $n != 0 ] && lib=$PWD/$PKG/libig_abl$n.so
    echo "== ablate $n"; INSTAGEO_B200_LIB=$lib timeout 120 python tools/gpu_probe.py perf 2>&1 | grep "attention"
  done | tee gpurun_out/attn_ablate.txt
fi
