# Builds a variant of libecho_b200 with extra compiler flags into variants/lib_<name>.so (a build product, not kept):
#   bash variants/build_variant.sh stream -DECHO_STREAM_STATE=1
# Load it with ECHO_B200_LIBRARY=$PWD/variants/lib_<name>.so (the Python side) — how every A/B of a compile-time switch is run.
set -e
name=$1; shift
cd "$(dirname "$0")/../echorenderer_b200/csrc"
out=/tmp/echo_variant_$name
rm -rf $out && mkdir -p $out
for f in api trace instanced build render debug peaks; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC "$@" -c $f.cu -o $out/$f.o &
done
wait
nvcc -shared -o ../../variants/lib_$name.so $out/*.o -gencode arch=compute_100a,code=sm_100a
ls -la ../../variants/lib_$name.so
