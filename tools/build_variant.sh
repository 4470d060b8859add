#!/bin/bash
# Dev tool: build an experiment variant of libdmesh_b200.so next to the real one, without touching it.
#   tools/build_variant.sh rcp -DDMR_TRI_BWD_RCP_ALPHA=1            -> tools/_bin/libdmesh_b200_rcp.so
#   tools/build_variant.sh gl4 -DDMR_TRI_BWD_GROUP_LANES=4
# and compare on the GPU box (one gpurun call):
#   python tools/time_compare.py C2 C5; DMESH_B200_LIB=tools/_bin/libdmesh_b200_rcp.so python tools/time_compare.py C2 C5
#   DMESH_B200_LIB=$PWD/tools/_bin/libdmesh_b200_rcp.so python -m pytest tests -m gpu -x -q
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/dmesh_renderer_b200/csrc
out=$root/tools/_bin
mkdir -p $out/obj_$name
objs=()
for f in capi capi_tet preprocess binning radix_sort tri_render tet_kernels collective inverse; do
  nvcc -c $src/$f.cu -o $out/obj_$name/$f.o -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC "$@" &
  objs+=($out/obj_$name/$f.o)
done
wait
nvcc -shared -o $out/libdmesh_b200_$name.so "${objs[@]}" -gencode arch=compute_100a,code=sm_100a -lcudart
echo $out/libdmesh_b200_$name.so
