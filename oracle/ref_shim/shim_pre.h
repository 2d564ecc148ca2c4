// TEST INFRASTRUCTURE ONLY -- part of the oracle, never linked into the product.
// Prologue for the translation unit that build_ref.sh assembles from the reference
// sources where they lie under /root/reference (outputs only in oracle/_ref/).
// Declares the handful of external types the extracted reference bodies expect
// (Vulkan handle, SPIR-V ray-flag constants, debug macro), nothing else.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cassert>
#include <algorithm>
#include <vector>
#include <map>
#include <deque>
#include <list>
#include <bitset>
#include <set>
#include <string>
#include <fstream>
#include <iostream>
#include <sys/types.h>
#include <vector_types.h>   // CUDA float3/float4/dim3 (header-only use)

typedef void* VkAccelerationStructureKHR;
enum VkGeometryTypeKHR { VK_GEOMETRY_TYPE_TRIANGLES_KHR = 0, VK_GEOMETRY_TYPE_AABBS_KHR = 1 };
enum VkDescriptorType { VK_DESCRIPTOR_TYPE_ACCELERATION_STRUCTURE_KHR = 1000150000, VK_DESCRIPTOR_TYPE_ACCELERATION_STRUCTURE_NV = 1000165000 };   /* the two values dump_descriptor_set_for_AS switches on */
enum {
  SpvRayFlagsOpaqueKHRMask = 0x1,
  SpvRayFlagsTerminateOnFirstHitKHRMask = 0x4,
  SpvRayFlagsSkipClosestHitShaderKHRMask = 0x8
};
#define VSIM_DPRINTF(...)
#define MAX_DESCRIPTOR_SETS 1
#define MAX_DESCRIPTOR_SET_BINDINGS 32

class ptx_instruction;
class ptx_thread_info;
typedef unsigned long long new_addr_type;
typedef unsigned long long address_type;
