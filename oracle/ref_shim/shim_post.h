// TEST INFRASTRUCTURE ONLY -- stubs for the simulator objects the extracted reference
// bodies touch (GPGPU_Context() counters/config, ptx_thread_info setters, the warp
// intersection table) plus a trimmed VulkanRayTracing declaration.  Own code.
#pragma once

// ---- simulator-global counters (cuda-sim.h:155-166 in the reference) ----
struct ref_func_sim {
  unsigned g_rt_mem_access_type[9];
  unsigned g_rt_num_hits, g_rt_num_any_hits;
  bool g_rt_world_set;
  float3 g_rt_world_min, g_rt_world_max;
  unsigned g_n_anyhit_rays, g_n_closesthit_rays;
  unsigned g_max_nodes_per_ray;
  unsigned long long g_tot_nodes_per_ray;
  unsigned g_max_tree_depth;
};
struct ref_core_config { bool remap_to_treelet_layout; bool load_treelet_metadata; unsigned treelet_remap_stride; };
struct ref_core { ref_core_config cfg; const ref_core_config* get_config() const { return &cfg; } };
struct ref_cluster { ref_core core; ref_core* cores[1]; ref_cluster() { cores[0] = &core; } ref_core** get_m_core() { return cores; } };
struct ref_gpu_config { int max_treelet_size; };
struct ref_gpu {
  ref_cluster cluster; ref_cluster* clusters[1]; ref_gpu_config cfg; unsigned long long gpu_sim_cycle;
  ref_gpu() : gpu_sim_cycle(0) { clusters[0] = &cluster; }
  ref_cluster** get_m_cluster() { return clusters; }
  ref_gpu_config& get_config() { return cfg; }
};
struct ref_gpgpusim { ref_gpu gpu; ref_gpu* g_the_gpu; ref_gpgpusim() { g_the_gpu = &gpu; } };
struct gpgpu_context {
  ref_func_sim fs; ref_func_sim* func_sim; ref_gpgpusim sim; ref_gpgpusim* the_gpgpusim;
  gpgpu_context() { memset(&fs, 0, sizeof(fs)); func_sim = &fs; the_gpgpusim = &sim; }
};
static gpgpu_context* GPGPU_Context() { static gpgpu_context c; return &c; }

struct memory_space {
  void write(void* addr, size_t n, const void* data, ptx_thread_info*, const ptx_instruction*) { memcpy(addr, data, n); }
  void read(const void* addr, size_t n, void* out) { memcpy(out, addr, n); }
};

struct ref_rt_thread_data {
  std::vector<Traversal_data*> traversal_data;
  std::vector<Hit_data*> all_hit_data;
  float3 hit_attribute; bool has_attribute;
  ref_rt_thread_data() : has_attribute(false) {}
  void set_hitAttribute(float3 b, const ptx_instruction*, ptx_thread_info*) { hit_attribute = b; has_attribute = true; }
};

class ptx_thread_info {
 public:
  ref_rt_thread_data data; ref_rt_thread_data* RT_thread_data;
  memory_space mem;
  std::vector<Ray> rays; unsigned n_intersect;
  std::vector<MemoryTransactionRecord> txns;
  std::vector<MemoryStoreTransactionRecord> store_txns;
  unsigned tid_x;   // lane inside its warp-sized CTA row: what the warp intersection tables are indexed by
  ptx_thread_info() : n_intersect(0), tid_x(0) { RT_thread_data = &data; }
  void add_ray_properties(Ray r) { rays.push_back(r); }
  void add_ray_intersect() { n_intersect++; }
  void set_rt_transactions(std::vector<MemoryTransactionRecord> t) { txns = t; }
  void set_rt_store_transactions(std::vector<MemoryStoreTransactionRecord> t) { store_txns = t; }
  memory_space* get_global_memory() { return &mem; }
  dim3 get_ctaid() const { return dim3(0, 0, 0); }
  dim3 get_tid() const { return dim3(tid_x, 0, 0); }
  unsigned get_uid() const { return 0; }
  unsigned get_hw_tid() const { return 0; }
};

enum class IntersectionTableType { Baseline, Function_Call_Coalescing };
class warp_intersection_table;   // the reference's own class (intersection_table.h) is spliced in right after this file

struct DESCRIPTOR_SET_STRUCT;

class VulkanRayTracing {
 public:
  static struct DESCRIPTOR_SET_STRUCT* descriptorSet;
  static void* launcher_descriptorSets[MAX_DESCRIPTOR_SETS][MAX_DESCRIPTOR_SET_BINDINGS];
  static void* launcher_deviceDescriptorSets[MAX_DESCRIPTOR_SETS][MAX_DESCRIPTOR_SET_BINDINGS];
  static std::map<void*, void*> blas_addr_map;
  static void* tlas_addr;
  static bool dumped;
  static warp_intersection_table*** intersection_table;
  static warp_intersection_table*** anyhit_table;

  static std::map<StackEntry, std::vector<StackEntry> > treelet_roots;
  static std::map<uint8_t*, std::vector<StackEntry> > treelet_roots_addr_only;
  static std::map<StackEntry, std::vector<StackEntry> > treelet_child_map;
  static std::map<uint8_t*, std::vector<StackEntry> > treelet_addr_only_child_map;
  static std::map<uint8_t*, uint8_t*> node_map_addr_only;
  static void* treelet_metadata;
  static std::map<uint8_t*, unsigned> treelet_addr_to_metadata_idx;
  static unsigned per_treelet_metadata_size;
  static uint8_t* treelet_layout_bvh;
  static std::map<uint8_t*, uint8_t*> original_bvh_to_treelet_bvh_mapping;
  static unsigned accessedDataSize;

  static bool mt_ray_triangle_test(float3 p0, float3 p1, float3 p2, Ray ray_properties, float* thit);
  static float3 Barycentric(float3 p, float3 a, float3 b, float3 c);
  static void traceRay(VkAccelerationStructureKHR, uint, uint, uint, uint, uint, float3, float, float3, float, int,
                       const ptx_instruction*, ptx_thread_info*);
  static void traceRayWithTreelets(VkAccelerationStructureKHR, uint, uint, uint, uint, uint, float3, float, float3, float, int,
                                   const ptx_instruction*, ptx_thread_info*);
  static void createTreelets(VkAccelerationStructureKHR _topLevelAS, int64_t device_offset, int maxBytesPerTreelet);
  static void remapBVHToTreeletLayout();
  static float calculateSAH(float3 lo, float3 hi);
  static bool isTreeletRoot(StackEntry node);
  static bool isTreeletRoot(uint8_t* addr);
  static uint8_t* addrToTreeletID(uint8_t* addr);
  static std::vector<StackEntry> treeletIDToChildren(StackEntry treelet_root);
  static std::vector<StackEntry> treeletIDToChildren(uint8_t* treelet_root);
  static void buildNodeToRootMap();
  static void dump_AS(struct DESCRIPTOR_SET_STRUCT*, VkAccelerationStructureKHR) {}
  /* the AS dumper: bodies taken from vulkan_ray_tracing.cc:4455-4558, :4886-4889, :4901-4945 */
  static std::vector<void*> child_addrs_from_driver;
  static void pass_child_addr(void* address);
  static void findOffsetBounds(int64_t& max_backwards, int64_t& min_backwards, int64_t& min_forwards, int64_t& max_forwards, VkAccelerationStructureKHR _topLevelAS);
  static void dump_descriptor_set_for_AS(uint32_t setID, uint32_t descID, void* address, uint32_t desc_size, VkDescriptorType type, uint32_t backwards_range,
                                         uint32_t forward_range, bool split_files, VkAccelerationStructureKHR _topLevelAS);
  static void* gpgpusim_alloc(uint32_t size) { return calloc(1, size); }
  static void* gpgpusim_malloc(uint32_t size) { return calloc(1, size); }
};

struct DESCRIPTOR_SET_STRUCT* VulkanRayTracing::descriptorSet = NULL;
void* VulkanRayTracing::launcher_descriptorSets[MAX_DESCRIPTOR_SETS][MAX_DESCRIPTOR_SET_BINDINGS] = {{NULL}};
void* VulkanRayTracing::launcher_deviceDescriptorSets[MAX_DESCRIPTOR_SETS][MAX_DESCRIPTOR_SET_BINDINGS] = {{NULL}};
std::map<void*, void*> VulkanRayTracing::blas_addr_map;
void* VulkanRayTracing::tlas_addr = NULL;
bool VulkanRayTracing::dumped = false;
std::vector<void*> VulkanRayTracing::child_addrs_from_driver;
warp_intersection_table*** VulkanRayTracing::intersection_table = NULL;
warp_intersection_table*** VulkanRayTracing::anyhit_table = NULL;
bool use_external_launcher = false;
const bool dump_trace = false;
