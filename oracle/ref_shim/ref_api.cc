// TEST INFRASTRUCTURE ONLY -- extern "C" driver around the reference's own function
// bodies (assembled before this file by build_ref.sh).  Own code; it only calls the
// reference entry points createTreelets / traceRay / traceRayWithTreelets and copies
// their results out into flat arrays that tests and bench.py's cpu_baseline can read.
#include <unistd.h>
#include <fcntl.h>

extern "C" {

struct ref_ray { float origin[3]; float tmin; float dir[3]; float tmax; uint32_t flags, cull_mask, sbt_offset, sbt_stride, miss_index; };
struct ref_hit { uint32_t hit; float t; uint32_t prim, geom, instance_id; float bary[3]; float point[3]; uint32_t n_all_hits; };
struct ref_txn { uint64_t address; uint32_t size; uint32_t type; };
struct ref_counters {
  uint64_t mem_access_type[9];
  uint64_t num_hits, num_any_hits, n_anyhit_rays, n_closesthit_rays;
  uint64_t max_nodes_per_ray, tot_nodes_per_ray, max_tree_depth, accessed_data_size, ray_count;
};

static int g_saved_stdout = -1;
static void silence(bool on) {
  fflush(stdout);
  if (on) {
    if (g_saved_stdout >= 0) return;
    g_saved_stdout = dup(1);
    int nul = open("/dev/null", O_WRONLY);
    dup2(nul, 1); close(nul);
  } else {
    if (g_saved_stdout < 0) return;
    dup2(g_saved_stdout, 1); close(g_saved_stdout); g_saved_stdout = -1;
  }
}

// ref_trace() keeps the shader tables out of the way (a table holds 100 rows per thread and asserts beyond, and ref_trace runs
// every ray as thread 0): a table that records nothing.  ref_trace_tables() switches the reference's own Baseline tables in.
struct noop_intersection_table : public warp_intersection_table {
  std::pair<std::vector<MemoryTransactionRecord>, std::vector<MemoryStoreTransactionRecord> >
  add_intersection(uint32_t, uint32_t, uint32_t, uint32_t, const ptx_instruction*, ptx_thread_info*) {
    return std::pair<std::vector<MemoryTransactionRecord>, std::vector<MemoryStoreTransactionRecord> >();
  }
  void clear(const ptx_instruction*, ptx_thread_info*) {}
  bool shader_exists(uint32_t, uint32_t, const ptx_instruction*, ptx_thread_info*) { return false; }
  bool exit_shaders(uint32_t, uint32_t) { return true; }
  uint32_t get_primitiveID(uint32_t, uint32_t, const ptx_instruction*, ptx_thread_info*) { return 0; }
  uint32_t get_instanceID(uint32_t, uint32_t, const ptx_instruction*, ptx_thread_info*) { return 0; }
  uint32_t get_hitGroupIndex(uint32_t, uint32_t, const ptx_instruction*, ptx_thread_info*) { return 0; }
  void* get_shader_data_address(uint32_t, uint32_t) { return NULL; }
};
static noop_intersection_table g_noop_table;
static warp_intersection_table* g_row[1] = { &g_noop_table };
static warp_intersection_table** g_tab[1] = { g_row };
// one CTA's pair of Baseline tables (vulkan_ray_tracing.cc:424-447): intersection_table[0][0] and anyhit_table[0][0]
static Baseline_warp_intersection_table* g_itab = NULL;
static Baseline_warp_intersection_table* g_atab = NULL;
static warp_intersection_table* g_irow[1]; static warp_intersection_table** g_itab3[1] = { g_irow };
static warp_intersection_table* g_arow[1]; static warp_intersection_table** g_atab3[1] = { g_arow };

void ref_reset(void) {
  VulkanRayTracing::treelet_roots.clear();
  VulkanRayTracing::treelet_roots_addr_only.clear();
  VulkanRayTracing::treelet_child_map.clear();
  VulkanRayTracing::treelet_addr_only_child_map.clear();
  VulkanRayTracing::node_map_addr_only.clear();
  VulkanRayTracing::treelet_addr_to_metadata_idx.clear();
  VulkanRayTracing::original_bvh_to_treelet_bvh_mapping.clear();
  VulkanRayTracing::blas_addr_map.clear();
  VulkanRayTracing::tlas_addr = NULL;
  VulkanRayTracing::accessedDataSize = 0;
  if (!g_itab) { g_itab = new Baseline_warp_intersection_table(); g_atab = new Baseline_warp_intersection_table(); }
  g_itab->clear(NULL, NULL); g_atab->clear(NULL, NULL);
  g_irow[0] = g_itab; g_arow[0] = g_atab;
  VulkanRayTracing::intersection_table = g_tab;
  VulkanRayTracing::anyhit_table = g_tab;
  rayCount = 0;
  treeletsFormed = false;
  gpgpu_context* c = GPGPU_Context();
  memset(&c->fs, 0, sizeof(c->fs));
}

void ref_config(int max_treelet_size, int remap_to_treelet_layout, unsigned treelet_remap_stride, int load_treelet_metadata) {
  gpgpu_context* c = GPGPU_Context();
  c->sim.gpu.cfg.max_treelet_size = max_treelet_size;
  c->sim.gpu.cluster.core.cfg.remap_to_treelet_layout = remap_to_treelet_layout != 0;
  c->sim.gpu.cluster.core.cfg.treelet_remap_stride = treelet_remap_stride;
  c->sim.gpu.cluster.core.cfg.load_treelet_metadata = load_treelet_metadata != 0;
}

// mirrors VulkanRayTracing::allocTLAS / allocBLAS (vulkan_ray_tracing.cc:4891-4899) without the printf
void ref_alloc_tlas(void* root, uint64_t, void* dev) { VulkanRayTracing::tlas_addr = dev; }
void ref_alloc_blas(void* root, uint64_t, void* dev) { VulkanRayTracing::blas_addr_map[root] = dev; }

void ref_form_treelets(void* tlas) {
  if (treeletsFormed) return;
  silence(true);
  int64_t off = (uint64_t)VulkanRayTracing::tlas_addr - (uint64_t)tlas;
  VulkanRayTracing::createTreelets(tlas, off, GPGPU_Context()->sim.gpu.cfg.max_treelet_size);
  treeletsFormed = true;
  silence(false);
}

uint64_t ref_treelet_count(void) { return VulkanRayTracing::treelet_roots_addr_only.size(); }
uint64_t ref_treelet_total_nodes(void) {
  uint64_t n = 0;
  for (auto& kv : VulkanRayTracing::treelet_roots_addr_only) n += kv.second.size();
  return n;
}
// roots ascending; per root: node count, metadata idx; node lists concatenated in root order
void ref_treelet_table(uint64_t* roots, uint32_t* counts, uint32_t* meta_idx, uint64_t* node_addr, uint32_t* node_size) {
  uint64_t i = 0, k = 0;
  for (auto& kv : VulkanRayTracing::treelet_roots_addr_only) {
    roots[i] = (uint64_t)kv.first;
    counts[i] = (uint32_t)kv.second.size();
    meta_idx[i] = VulkanRayTracing::treelet_addr_to_metadata_idx.count(kv.first)
                      ? VulkanRayTracing::treelet_addr_to_metadata_idx[kv.first] : 0xffffffffu;
    for (auto& e : kv.second) { node_addr[k] = (uint64_t)e.addr; node_size[k] = (uint32_t)e.size; k++; }
    i++;
  }
}
uint64_t ref_node_map_size(void) { return VulkanRayTracing::node_map_addr_only.size(); }
void ref_node_map(uint64_t* nodes, uint64_t* roots) {
  uint64_t i = 0;
  for (auto& kv : VulkanRayTracing::node_map_addr_only) { nodes[i] = (uint64_t)kv.first; roots[i] = (uint64_t)kv.second; i++; }
}
uint64_t ref_remap_size(void) { return VulkanRayTracing::original_bvh_to_treelet_bvh_mapping.size(); }
uint64_t ref_remap_base(void) { return (uint64_t)VulkanRayTracing::treelet_layout_bvh; }
void ref_remap(uint64_t* orig, uint64_t* mapped) {
  uint64_t i = 0;
  for (auto& kv : VulkanRayTracing::original_bvh_to_treelet_bvh_mapping) { orig[i] = (uint64_t)kv.first; mapped[i] = (uint64_t)kv.second; i++; }
}
int ref_addr_to_treelet(uint64_t addr, uint64_t* root) {
  auto it = VulkanRayTracing::node_map_addr_only.find((uint8_t*)addr);
  if (it == VulkanRayTracing::node_map_addr_only.end()) return -1;
  *root = (uint64_t)it->second; return 0;
}
int ref_is_treelet_root(uint64_t addr) { return VulkanRayTracing::isTreeletRoot((uint8_t*)addr) ? 1 : 0; }

// mode 0 = traceRay (DFS), 1 = traceRayWithTreelets.  Returns total #transactions, or -(needed) if cap too small.
// treelet_ids may be NULL; otherwise addrToTreeletID(txn.address) where the address is known, else ~0.
int64_t ref_trace(void* tlas, int mode, uint32_t n, const ref_ray* rays, ref_hit* hits, uint32_t* counts,
                  ref_txn* txns, uint64_t cap, uint64_t* treelet_ids, int keep_stdout) {
  if (!keep_stdout) silence(true);
  uint64_t total = 0; bool overflow = false;
  for (uint32_t i = 0; i < n; i++) {
    const ref_ray& r = rays[i];
    ptx_thread_info th;
    float3 o = { r.origin[0], r.origin[1], r.origin[2] }, d = { r.dir[0], r.dir[1], r.dir[2] };
    if (mode == 1)
      VulkanRayTracing::traceRayWithTreelets(tlas, r.flags, r.cull_mask, r.sbt_offset, r.sbt_stride, r.miss_index, o, r.tmin, d, r.tmax, 0, NULL, &th);
    else
      VulkanRayTracing::traceRay(tlas, r.flags, r.cull_mask, r.sbt_offset, r.sbt_stride, r.miss_index, o, r.tmin, d, r.tmax, 0, NULL, &th);
    Traversal_data* td = th.data.traversal_data.back();
    if (hits) {
      ref_hit& h = hits[i];
      memset(&h, 0, sizeof(h));
      h.hit = td->hit_geometry ? 1u : 0u;
      h.n_all_hits = (mode == 0) ? td->n_all_hits : 0u;
      if (td->hit_geometry) {
        h.t = td->closest_hit.world_min_thit;
        h.prim = td->closest_hit.primitive_index; h.geom = td->closest_hit.geometry_index; h.instance_id = td->closest_hit.instance_index;
        h.bary[0] = td->closest_hit.barycentric_coordinates.x; h.bary[1] = td->closest_hit.barycentric_coordinates.y; h.bary[2] = td->closest_hit.barycentric_coordinates.z;
        h.point[0] = td->closest_hit.intersection_point.x; h.point[1] = td->closest_hit.intersection_point.y; h.point[2] = td->closest_hit.intersection_point.z;
      }
    }
    free(td);
    for (auto* p : th.data.all_hit_data) free(p);
    if (counts) counts[i] = (uint32_t)th.txns.size();
    for (auto& t : th.txns) {
      if (total < cap && txns) {
        txns[total].address = (uint64_t)t.address; txns[total].size = t.size; txns[total].type = (uint32_t)t.type;
        if (treelet_ids) {
          auto it = VulkanRayTracing::node_map_addr_only.find((uint8_t*)t.address);
          treelet_ids[total] = it == VulkanRayTracing::node_map_addr_only.end() ? ~0ull : (uint64_t)it->second;
        }
      } else if (txns) overflow = true;
      total++;
    }
  }
  if (!keep_stdout) silence(false);
  return overflow ? -(int64_t)total : (int64_t)total;
}

// ---- shader-table side effects (Baseline tables): rays [32g, 32g+32) are the threads tid.x = i % 32 of one
// CTA whose two tables start empty.  Per add_intersection call: which table, the row (index[tid] before the call), the
// values written, and the two MemoryStoreTransactionRecords; per any-hit call the Hit_data pushed to all_hit_data.
struct ref_table_event { uint32_t table, shader_counter, hit_group_index, primitive_id, instance_id, tid; uint64_t store_addr[2]; uint32_t store_size[2]; };
void ref_table_bases(uint64_t* itab, uint64_t* atab) { *itab = (uint64_t)g_itab->table; *atab = (uint64_t)g_atab->table; }
int64_t ref_trace_tables(void* tlas, int mode, uint32_t n, const ref_ray* rays, uint32_t* ev_counts, ref_table_event* ev, ref_hit* anyhit, uint64_t cap) {
  silence(true);
  VulkanRayTracing::intersection_table = g_itab3; VulkanRayTracing::anyhit_table = g_atab3;
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (i % 32 == 0) { g_itab->clear(NULL, NULL); g_atab->clear(NULL, NULL); }
    const ref_ray& r = rays[i];
    ptx_thread_info th; th.tid_x = i % 32;
    float3 o = { r.origin[0], r.origin[1], r.origin[2] }, d = { r.dir[0], r.dir[1], r.dir[2] };
    if (mode == 1)
      VulkanRayTracing::traceRayWithTreelets(tlas, r.flags, r.cull_mask, r.sbt_offset, r.sbt_stride, r.miss_index, o, r.tmin, d, r.tmax, 0, NULL, &th);
    else
      VulkanRayTracing::traceRay(tlas, r.flags, r.cull_mask, r.sbt_offset, r.sbt_stride, r.miss_index, o, r.tmin, d, r.tmax, 0, NULL, &th);
    free(th.data.traversal_data.back());
    // the Baseline table emits exactly two stores per call (intersection_table.cc:180-181): hitGroupIndex[tid], shader_data[tid]
    const uint32_t n_ev = (uint32_t)(th.store_txns.size() / 2);
    uint32_t k_any = 0;
    for (uint32_t k = 0; k < n_ev; k++) {
      const MemoryStoreTransactionRecord& s0 = th.store_txns[2 * k]; const MemoryStoreTransactionRecord& s1 = th.store_txns[2 * k + 1];
      const uint64_t a0 = (uint64_t)s0.address;
      const bool is_any = a0 >= (uint64_t)g_atab->table && a0 < (uint64_t)(g_atab->table + INTERSECTION_TABLE_MAX_LENGTH);
      Baseline_warp_intersection_table* t = is_any ? g_atab : g_itab;
      const uint64_t off = a0 - (uint64_t)t->table;
      if (total < cap && ev) {
        ref_table_event& e = ev[total];
        e.table = is_any ? 1u : 0u; e.shader_counter = (uint32_t)(off / sizeof(Baseline_Entry)); e.tid = (uint32_t)((off % sizeof(Baseline_Entry)) / 4);
        e.hit_group_index = t->table[e.shader_counter].hitGroupIndex[e.tid];
        e.primitive_id = t->table[e.shader_counter].shader_data[e.tid].primitiveID; e.instance_id = t->table[e.shader_counter].shader_data[e.tid].instanceID;
        e.store_addr[0] = a0; e.store_size[0] = s0.size; e.store_addr[1] = (uint64_t)s1.address; e.store_size[1] = s1.size;
        if (anyhit) {
          ref_hit& h = anyhit[total]; memset(&h, 0, sizeof(h));
          if (is_any && k_any < th.data.all_hit_data.size()) {
            const Hit_data* hd = th.data.all_hit_data[k_any];
            h.hit = 1; h.t = hd->world_min_thit; h.prim = hd->primitive_index; h.geom = hd->geometry_index; h.instance_id = hd->instance_index;
            h.bary[0] = hd->barycentric_coordinates.x; h.bary[1] = hd->barycentric_coordinates.y; h.bary[2] = hd->barycentric_coordinates.z;
            h.point[0] = hd->intersection_point.x; h.point[1] = hd->intersection_point.y; h.point[2] = hd->intersection_point.z;
          }
        }
      }
      if (is_any) k_any++;
      total++;
    }
    for (auto* p : th.data.all_hit_data) free(p);
    if (ev_counts) ev_counts[i] = n_ev;
  }
  VulkanRayTracing::intersection_table = g_tab; VulkanRayTracing::anyhit_table = g_tab;
  silence(false);
  return (int64_t)total;
}

// ---- Function_Call_Coalescing intersection table (-gpgpu_rt_intersection_table_type 1, intersection_table.cc:36-99): rays
// [32g, 32g+32) are the threads tid.x = i % 32 of one CTA, traced in order like execute_warp_inst_t does, against a FRESH
// reference table per CTA.  Returned per ray: its whole transaction list (the Intersection_Table_Load records merged in by
// :2186-2200 / :2966-2980 included) and its store list; table addresses are reported relative to the table base
// (| 1 << 63), everything else as is.  mode 0 (traceRay) only: traceRayWithTreelets passes every record of the finished list
// to addrToTreeletID (:2256-2262), whose assert (:470) fires on the first table address.
void Coalescing_warp_intersection_table::clear(const ptx_instruction*, ptx_thread_info*) {
  for (uint32_t i = 0; i < tableSize; i++) for (int j = 0; j < 32; j++) table[i].thread_mask[j] = false;
  tableSize = 0;
}
struct ref_store_rec { uint64_t address; uint32_t size; uint32_t type; };
int64_t ref_trace_coalescing(void* tlas, int mode, uint32_t n, const ref_ray* rays, uint64_t* txn_offsets, ref_txn* txns, uint64_t txn_cap,
                             uint64_t* store_offsets, ref_store_rec* stores, uint64_t store_cap, uint32_t* entry_size) {
  silence(true);
  static warp_intersection_table* crow[1]; static warp_intersection_table** ctab[1] = { crow };
  Coalescing_warp_intersection_table* tab = NULL;
  if (entry_size) *entry_size = (uint32_t)sizeof(Coalescing_Entry);
  uint64_t nt = 0, ns = 0;
  for (uint32_t i = 0; i < n; i++) {
    if (i % 32 == 0) { if (tab) { free(tab->table); tab->table = NULL; operator delete(tab); } tab = new Coalescing_warp_intersection_table(); crow[0] = tab; }
    VulkanRayTracing::intersection_table = ctab; VulkanRayTracing::anyhit_table = g_tab;
    const ref_ray& r = rays[i];
    ptx_thread_info th; th.tid_x = i % 32;
    float3 o = { r.origin[0], r.origin[1], r.origin[2] }, d = { r.dir[0], r.dir[1], r.dir[2] };
    if (mode == 1)
      VulkanRayTracing::traceRayWithTreelets(tlas, r.flags, r.cull_mask, r.sbt_offset, r.sbt_stride, r.miss_index, o, r.tmin, d, r.tmax, 0, NULL, &th);
    else
      VulkanRayTracing::traceRay(tlas, r.flags, r.cull_mask, r.sbt_offset, r.sbt_stride, r.miss_index, o, r.tmin, d, r.tmax, 0, NULL, &th);
    free(th.data.traversal_data.back());
    for (auto* p : th.data.all_hit_data) free(p);
    const uint64_t b0 = (uint64_t)tab->table, b1 = b0 + sizeof(Coalescing_Entry) * INTERSECTION_TABLE_MAX_LENGTH;
    txn_offsets[i] = nt; store_offsets[i] = ns;
    for (const MemoryTransactionRecord& t : th.txns) {
      if (txns && nt < txn_cap) { uint64_t a = (uint64_t)t.address; if (a >= b0 && a < b1) a = (a - b0) | (1ull << 63); txns[nt].address = a; txns[nt].size = t.size; txns[nt].type = (uint32_t)t.type; }
      nt++;
    }
    for (const MemoryStoreTransactionRecord& t : th.store_txns) {
      if (stores && ns < store_cap) { uint64_t a = (uint64_t)t.address; if (a >= b0 && a < b1) a = (a - b0) | (1ull << 63); stores[ns].address = a; stores[ns].size = t.size; stores[ns].type = (uint32_t)t.type; }
      ns++;
    }
  }
  txn_offsets[n] = nt; store_offsets[n] = ns;
  if (tab) { free(tab->table); tab->table = NULL; operator delete(tab); }
  VulkanRayTracing::intersection_table = g_tab; VulkanRayTracing::anyhit_table = g_tab;
  silence(false);
  return (int64_t)nt;
}

// The reference's own AS dumper: writes <mesa_root>gpgpusimShaders/0_0.asmain / .asback / .asfront / .asmetadata for a TLAS of
// desc_size bytes whose BLAS headers are kids[] (what Mesa passes through gpgpusim_pass_child_addr).  mesa_root must end in '/'
// and the directory gpgpusimShaders/ must exist under it.
void ref_dump_as(const char* mesa_root, void* tlas, uint32_t desc_size, void** kids, uint32_t n_kids) {
  setenv("MESA_ROOT", mesa_root, 1);
  VulkanRayTracing::child_addrs_from_driver.clear();
  for (uint32_t i = 0; i < n_kids; i++) VulkanRayTracing::pass_child_addr(kids[i]);
  VulkanRayTracing::dump_descriptor_set_for_AS(0, 0, tlas, desc_size, VK_DESCRIPTOR_TYPE_ACCELERATION_STRUCTURE_KHR, 0, 0, true, tlas);
}

void ref_get_counters(ref_counters* out) {
  ref_func_sim& f = GPGPU_Context()->fs;
  for (int i = 0; i < 9; i++) out->mem_access_type[i] = f.g_rt_mem_access_type[i];
  out->num_hits = f.g_rt_num_hits; out->num_any_hits = f.g_rt_num_any_hits;
  out->n_anyhit_rays = f.g_n_anyhit_rays; out->n_closesthit_rays = f.g_n_closesthit_rays;
  out->max_nodes_per_ray = f.g_max_nodes_per_ray; out->tot_nodes_per_ray = f.g_tot_nodes_per_ray;
  out->max_tree_depth = f.g_max_tree_depth; out->accessed_data_size = VulkanRayTracing::accessedDataSize;
  out->ray_count = rayCount;
}

}  // extern "C"
