#!/bin/sh
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libvsrt_ref.so: the reference's OWN
# implementation of the hot path (createTreelets / traceRay / traceRayWithTreelets and
# the decode + math helpers), compiled from the sources where they lie under
# /root/reference with the reference's flags (-O3 -fpermissive, src/cuda-sim/Makefile:74).
# Nothing from the reference is copied into the repository: the assembled translation
# unit and the .so exist only under oracle/_ref/ (git-ignored, shipped to the GPU box).
# Line ranges: SURVEY.md Appendix B.  The one dialect patch (pointer '> 0' -> '!= 0') is
# needed because g++-13 rejects an ordering comparison g++-9 accepted.  The compiler is
# the PATH g++ on purpose: this image exports CXX=/opt/gcc/bin/g++, a wrapper whose
# libstdc++ gets linked statically and whose std::cout then crashes inside a dlopen()ed .so.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
R=${VSRT_REFERENCE:-/root/reference}/src
OUT=$HERE/_ref
[ -d "$R" ] || { echo "build_ref.sh: reference sources not found at $R (using prebuilt $OUT if present)"; exit 0; }
mkdir -p "$OUT"
{
  cat "$HERE/ref_shim/shim_pre.h"
  sed -n '201,216p;315,329p;333,370p;375,380p' "$R/abstract_hardware_model.h"
  sed -n '44,56p;66,508p' "$R/cuda-sim/vulkan_acceleration_structure_util.h"
  sed -n '53,56p;180,208p' "$R/cuda-sim/vulkan_ray_tracing.h"
  sed -n '29,58p' "$R/cuda-sim/vulkan_rt_thread_data.h"
  sed -n '5,23p' "$R/gpgpu-sim/vector-math.h"
  grep -v '#include' "$R/gpgpu-sim/vector-math.cc"
  cat "$HERE/ref_shim/shim_post.h"
  # the warp intersection tables (abstract class, Coalescing and Baseline tables); members opened for the test driver (class -> struct)
  echo '#define class struct'
  sed -n '42p;52,67p;70,107p;109,139p' "$R/cuda-sim/intersection_table.h"
  echo '#undef class'
  sed -n '123,129p;136,140p' "$R/cuda-sim/vulkan_ray_tracing.cc"
  sed -n '148,257p;456,510p;823,1520p;1522,2307p;2309,3076p;3089,3130p' "$R/cuda-sim/vulkan_ray_tracing.cc"
  sed -n '155,192p;198,229p' "$R/cuda-sim/intersection_table.cc"
  # Coalescing table: constructor + add_intersection and the getters; its clear() (:101-111) is NOT taken -- the inner loop
  # increments i instead of j and never terminates on a non-empty table -- ref_api.cc defines the evident intent instead
  sed -n '36,99p;113,148p' "$R/cuda-sim/intersection_table.cc"
  # the AS dumper: dump_descriptor_set_for_AS, pass_child_addr, findOffsetBounds
  sed -n '4455,4558p;4886,4889p;4901,4945p' "$R/cuda-sim/vulkan_ray_tracing.cc"
  cat "$HERE/ref_shim/ref_api.cc"
  # ---- RT-unit replay helpers: RTMemoryTransactionRecord, rt_unit::sort_mem_accesses and the
  # treelet-prefetch vote block of rt_unit::cycle, the latter spliced in as the body of a member function
  sed -n '1372,1401p' "$R/abstract_hardware_model.h"
  cat "$HERE/ref_shim/shim_rtunit.h"
  sed -n '3012,3164p;4307,4392p' "$R/gpgpu-sim/shader.cc"
  echo 'void rt_unit::prefetch_vote_block() {'
  sed -n '3419,3684p' "$R/gpgpu-sim/shader.cc"
  echo '  out_root = prefetched_treelet_root; out_num_nodes = num_nodes_to_prefetch; out_seen = true; } }'
  cat "$HERE/ref_shim/ref_api_rtunit.cc"
} | sed 's/next_node_addr > 0/next_node_addr != 0/g' > "$OUT/ref_tu.cc"
g++ -O3 -fpermissive -w -std=c++14 -fPIC -shared -I/usr/local/cuda/include \
    "$OUT/ref_tu.cc" -o "$OUT/libvsrt_ref.so"
echo "built $OUT/libvsrt_ref.so"
