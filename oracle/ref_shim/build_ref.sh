#!/bin/bash
# Build oracle/_ref/libvv_ref.so: the reference's OWN layer sources, compiled where they lie under
# /root/reference (never copied), against the shim headers in oracle/ref_shim/include and accessor classes
# generated from the reference's .proto files.  Outputs only into oracle/_ref/ (git-ignored).
# The reference's build system is not used (it needs glog/gflags/boost/protobuf/lmdb/leveldb/hdf5, all absent).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
REF="${VV_REFERENCE:-/root/reference}"
OUT="$ROOT/oracle/_ref"
[ -d "$REF/src/caffe" ] || { echo "no reference tree at $REF: keeping any prebuilt oracle/_ref"; exit 0; }
mkdir -p "$OUT/obj"
# generated accessor headers: replaced only when their content changes, so that object files can depend on them by mtime
rm -rf "$OUT/gen.new"; python "$HERE/gen_pb_shim.py" "$REF" "$OUT/gen.new" > /dev/null
mkdir -p "$OUT/gen/caffe/proto"
for f in "$OUT"/gen.new/caffe/proto/*.h; do
  cmp -s "$f" "$OUT/gen/caffe/proto/$(basename "$f")" || cp "$f" "$OUT/gen/caffe/proto/$(basename "$f")"
done
rm -rf "$OUT/gen.new"
NEWEST_HDR="$(ls -t "$OUT"/gen/caffe/proto/*.h $(find "$HERE/include" -type f) | head -1)"
BLAS="$(python - <<'PY'
import glob, os, scipy
c = glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)), "scipy.libs", "libscipy_openblas*.so"))
print(c[0] if c else "")
PY
)"
[ -n "$BLAS" ] || { echo "no OpenBLAS found"; exit 1; }
# hidden visibility: the reference's caffe:: symbols must not interpose with (or be interposed by) the product's
# caffe_compat classes of the same names when both libraries are loaded into one process (bench.py)
CXXFLAGS="-std=c++14 -O2 -fPIC -fvisibility=hidden -fvisibility-inlines-hidden -DCPU_ONLY -w -include cstring -include climits -include unistd.h -include cstdlib -include numeric -I$HERE/include -I$OUT/gen -I$REF/include"
SRCS="blob syncedmem common util/math_functions layers/inner_product_layer layers/relu_layer layers/dropout_layer layers/eltwise_layer
      layers/normalization_layer layers/sum_layer layers/split_layer layers/slice_layer layers/concat_layer layers/flatten_layer
      layers/max_margin_loss_layer layers/loss_layer layers/neuron_layer layers/retrieval_stats_layer layers/id_to_weight_mapping_layer
      layers/video_sampled_shots_data_layer layers/video_shot_window_test_data_layer layers/base_data_layer internal_thread data_transformer
      net solver util/insert_splits"
OBJS=""
JOBS="$(nproc 2>/dev/null || echo 4)"
pids=""
for s in $SRCS; do
  o="$OUT/obj/$(basename $s).o"
  if [ ! -f "$o" ] || [ "$REF/src/caffe/$s.cpp" -nt "$o" ] || [ "$NEWEST_HDR" -nt "$o" ]; then
    g++ $CXXFLAGS -c "$REF/src/caffe/$s.cpp" -o "$o" &
    pids="$pids $!"
    while [ "$(jobs -rp | wc -l)" -ge "$JOBS" ]; do sleep 0.2; done
  fi
  OBJS="$OBJS $o"
done
for p in $pids; do wait "$p"; done
g++ $CXXFLAGS -c "$HERE/ref_driver.cpp" -o "$OUT/obj/ref_driver.o"
# link the SciPy wheel's OpenBLAS in place (same image, hence same path, on the GPU box)
g++ -shared -o "$OUT/libvv_ref.so" $OBJS "$OUT/obj/ref_driver.o" "$BLAS" -Wl,-Bsymbolic -Wl,-rpath,"$(dirname "$BLAS")" -lpthread
echo "built $OUT/libvv_ref.so"

# ---- the drop-in proof: the same reference sources in GPU mode (no CPU_ONLY), their device side supplied by
# dropin_gpu.cpp = calls into the product's C-ABI.  Needs the product library (make all) and the CUDA headers.
VVLIB="$ROOT/videovector_b200/lib/libvv_b200.so"
CUDA="${CUDA_HOME:-/usr/local/cuda}"
if [ -f "$VVLIB" ] && [ -f "$CUDA/include/cuda_runtime_api.h" ]; then
  mkdir -p "$OUT/obj_gpu"
  GPUFLAGS="${CXXFLAGS/-DCPU_ONLY/} -DVV_DROPIN_GPU -I$CUDA/include -I$ROOT/include"
  GOBJS=""
  pids=""
  for s in $SRCS; do
    o="$OUT/obj_gpu/$(basename $s).o"
    if [ ! -f "$o" ] || [ "$REF/src/caffe/$s.cpp" -nt "$o" ] || [ "$NEWEST_HDR" -nt "$o" ]; then
      g++ $GPUFLAGS -c "$REF/src/caffe/$s.cpp" -o "$o" &
      pids="$pids $!"
      while [ "$(jobs -rp | wc -l)" -ge "$JOBS" ]; do sleep 0.2; done
    fi
    GOBJS="$GOBJS $o"
  done
  for p in $pids; do wait "$p"; done
  g++ $GPUFLAGS -c "$HERE/ref_driver.cpp" -o "$OUT/obj_gpu/ref_driver.o"
  g++ $GPUFLAGS -c "$HERE/dropin_gpu.cpp" -o "$OUT/obj_gpu/dropin_gpu.o"
  g++ -shared -o "$OUT/libvv_dropin.so" $GOBJS "$OUT/obj_gpu/ref_driver.o" "$OUT/obj_gpu/dropin_gpu.o" "$BLAS" "$VVLIB" \
      -L"$CUDA/lib64" -lcudart -Wl,-Bsymbolic -Wl,--no-undefined \
      -Wl,-rpath,"$(dirname "$BLAS")" -Wl,-rpath,'$ORIGIN/../../videovector_b200/lib' -Wl,-rpath,"$CUDA/lib64" -lpthread
  echo "built $OUT/libvv_dropin.so"
else
  echo "product library or CUDA headers absent: skipping the drop-in build"
fi
