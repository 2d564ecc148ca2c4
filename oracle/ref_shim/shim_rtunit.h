// TEST INFRASTRUCTURE ONLY -- stand-in for the parts of rt_unit (gpgpu-sim/shader.h) that the two extracted bodies of
// shader.cc touch: rt_unit::sort_mem_accesses (:3012-3164), rt_unit::schedule_next_warp (:4307-4392) and the
// treelet-prefetch vote block of rt_unit::cycle (:3419-3685), which build_ref.sh splices in as the body of
// rt_unit::prefetch_vote_block().  Own code: only names and
// types the extracted lines need, no reference text.
#pragma once
#include <bitset>
#include <set>
#define THREAD_SORT_DPRINTF(...)
#define TOMMY_DPRINTF(...)
#define SECOND_PREFETCH_DPRINTF(...)
#define RT_SCHEDULER_DPRINTF(...)

struct ref_rt_thread_info { std::deque<RTMemoryTransactionRecord> RT_mem_accesses; };
struct ref_warp_inst {
  ref_rt_thread_info th[32];
  unsigned uid; bool stalled, m_empty;
  ref_warp_inst() : uid(0), stalled(false), m_empty(true) {}
  ref_rt_thread_info& get_thread_info(unsigned i) { return th[i]; }
  bool is_stalled() const { return stalled; }
  bool empty() const { return m_empty; }
  unsigned get_uid() const { return uid; }
};
typedef ref_warp_inst warp_inst_t;
struct ref_rt_config {
  bool m_treelet_prefetch; unsigned prefetch_delay; unsigned m_treelet_prefetch_heuristic; double m_treelet_prefetch_threshold;
  unsigned m_max_prefetch_queue_size; bool m_flush_prefetch_queue_on_new_treelet; bool load_treelet_metadata; unsigned m_treelet_scheduler;
  bool prefetch_next_treelet_when_queue_empty; unsigned m_sort_method;
};
struct ref_rt_gpu { unsigned long long gpu_sim_cycle, gpu_tot_sim_cycle; };
struct ref_rt_core { ref_rt_gpu gpu; ref_rt_gpu* get_gpu() { return &gpu; } };

class rt_unit {
 public:
  ref_rt_config cfg; ref_rt_config* m_config;
  ref_rt_core core; ref_rt_core* m_core;
  unsigned m_sid;
  std::map<unsigned, ref_warp_inst> m_current_warps;
  std::deque<std::pair<new_addr_type, new_addr_type> > prefetch_mem_access_q;
  std::deque<std::pair<unsigned long long, new_addr_type> > prefetch_generation_cycles;
  uint8_t* last_prefetched_treelet; uint8_t* last_rejected_treelet; uint8_t* last_prefetched_second_treelet;
  unsigned long long matches, comparisons, prefetch_treelet_switches, total_cycles_between_prefetch_treelet_switch,
      timestamp_of_last_treelet, prefetch_metadata_added, prefetches_added_to_queue;
  // what the vote block decided (copied out of its locals by the two marker lines build_ref.sh appends inside the block)
  uint8_t* out_root; int out_num_nodes; bool out_seen;
  rt_unit() { memset(&cfg, 0, sizeof(cfg)); m_config = &cfg; memset(&core, 0, sizeof(core)); m_core = &core; m_sid = 0; reset_state(); }
  void reset_state() {
    prefetch_mem_access_q.clear(); prefetch_generation_cycles.clear();
    last_prefetched_treelet = last_rejected_treelet = last_prefetched_second_treelet = NULL;
    matches = comparisons = prefetch_treelet_switches = total_cycles_between_prefetch_treelet_switch = 0;
    timestamp_of_last_treelet = prefetch_metadata_added = prefetches_added_to_queue = 0;
    out_root = NULL; out_num_nodes = 0; out_seen = false;
  }
  void sort_mem_accesses(std::deque<RTMemoryTransactionRecord>& mem_accesses, std::map<uint8_t*, int> node_access_counts_per_treelet = std::map<uint8_t*, int>());
  void prefetch_vote_block();
  void schedule_next_warp(warp_inst_t& inst);
};
