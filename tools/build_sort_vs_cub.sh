#!/bin/bash
# Dev tool: builds tools/_bin/sort_vs_cub (needs the in-tree libdmesh_b200.so; CUB from the CUDA toolkit).
set -e
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $root/tools/_bin
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o $root/tools/_bin/sort_vs_cub $root/tools/sort_vs_cub.cu \
  -L$root/dmesh_renderer_b200 -ldmesh_b200 -Xlinker -rpath -Xlinker '$ORIGIN/../../dmesh_renderer_b200'
echo $root/tools/_bin/sort_vs_cub
