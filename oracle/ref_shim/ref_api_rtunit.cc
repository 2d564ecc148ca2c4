// TEST INFRASTRUCTURE ONLY -- extern "C" driver around the two rt_unit bodies extracted from gpgpu-sim/shader.cc
// (assembled before this file by build_ref.sh).  Own code.
extern "C" {

// rt_unit::sort_mem_accesses on every ray of a CSR trace, in place.  method = -sort_method (0 strict, 1 loose).
void ref_sort_trace(int method, uint64_t n_rays, const uint64_t* offsets, ref_txn* txns) {
  rt_unit u; u.cfg.m_sort_method = (unsigned)method;
  for (uint64_t r = 0; r < n_rays; r++) {
    std::deque<RTMemoryTransactionRecord> q;
    for (uint64_t k = offsets[r]; k < offsets[r + 1]; k++) q.push_back(RTMemoryTransactionRecord(txns[k].address, txns[k].size, (TransactionType)txns[k].type));
    u.sort_mem_accesses(q);
    uint64_t k = offsets[r];
    for (auto& t : q) { txns[k].address = t.address; txns[k].size = t.size; txns[k].type = (uint32_t)t.type; k++; }
  }
}

struct ref_prefetch_decision { uint64_t root; uint32_t votes, total, submit, n_nodes, first_node, num_nodes; };

// The treelet-prefetch vote of rt_unit::cycle for ONE group of rays (the threads of the warps resident in the unit):
// thread t votes with the treelet of txns[offsets[ray] + front[ray]] unless its list is exhausted.  A fresh unit per
// call (no last_prefetched_treelet history), so the block's decision and the chunks it queues are a pure function of
// the inputs.  Returns the number of (chunk address, owner address) pairs pushed to prefetch_mem_access_q.
int64_t ref_prefetch_vote(int heuristic, double threshold, int load_metadata, uint64_t n_rays, const uint64_t* ray_ids,
                          const uint64_t* offsets, const uint32_t* front, const ref_txn* txns,
                          ref_prefetch_decision* dec, uint64_t* chunk_addr, uint64_t* chunk_owner, uint64_t cap) {
  rt_unit u;
  u.cfg.m_treelet_prefetch = true; u.cfg.prefetch_delay = 1; u.cfg.m_treelet_prefetch_heuristic = (unsigned)heuristic;
  u.cfg.m_treelet_prefetch_threshold = threshold; u.cfg.m_max_prefetch_queue_size = 0xffffffffu;
  u.cfg.load_treelet_metadata = load_metadata != 0;
  uint32_t votes = 0; std::map<uint64_t, uint32_t> tally;
  for (uint64_t i = 0; i < n_rays; i++) {
    const uint64_t r = ray_ids ? ray_ids[i] : i;
    ref_warp_inst& w = u.m_current_warps[(unsigned)(i / 32)];
    const uint64_t f = front ? front[r] : 0;
    for (uint64_t k = offsets[r] + f; k < offsets[r + 1]; k++)
      w.th[i % 32].RT_mem_accesses.push_back(RTMemoryTransactionRecord(txns[k].address, txns[k].size, (TransactionType)txns[k].type));
    if (offsets[r] + f < offsets[r + 1]) { votes++; tally[(uint64_t)VulkanRayTracing::addrToTreeletID((uint8_t*)txns[offsets[r] + f].address)]++; }
  }
  u.prefetch_vote_block();
  if (dec) {
    memset(dec, 0, sizeof(*dec));
    dec->root = (uint64_t)u.out_root; dec->total = votes; dec->votes = u.out_root ? tally[(uint64_t)u.out_root] : 0;
    dec->n_nodes = u.out_root ? (uint32_t)VulkanRayTracing::treelet_roots_addr_only[u.out_root].size() : 0;
    // :3548-3553: heuristics 2 and 3 prefetch num_nodes_to_prefetch nodes (3: the LAST ones), 0 and 1 the whole list
    // (no voter at all: the block computes (int)(0 * NaN + 0.5), which is not a value -- report zeros for "no root")
    dec->num_nodes = !u.out_root ? 0u : (heuristic == 2 || heuristic == 3) ? (uint32_t)u.out_num_nodes : dec->n_nodes;
    dec->first_node = heuristic == 3 ? dec->n_nodes - dec->num_nodes : 0u;
    dec->submit = u.prefetch_treelet_switches ? 1u : 0u;     // submit_prefetch && root != nullptr, seen through the counter at :3541
  }
  uint64_t n = 0;
  for (auto& p : u.prefetch_mem_access_q) { if (n < cap && chunk_addr) { chunk_addr[n] = p.first; chunk_owner[n] = p.second; } n++; }
  return (int64_t)n;
}

// rt_unit::schedule_next_warp (:4307-4392) for one RT unit: warps in m_current_warps order, lane l of warp w is ray
// ray_ids[32 * w + l] (~0 = no thread), pending access = txns[offsets[ray] + front[ray]].  Returns the index of the
// picked warp or -1.
int64_t ref_schedule_pick(int scheduler, uint64_t last_prefetched, uint64_t n_warps, const uint64_t* ray_ids, const uint8_t* stalled,
                          const uint64_t* offsets, const uint32_t* front, const ref_txn* txns) {
  rt_unit u;
  u.cfg.m_treelet_scheduler = (unsigned)scheduler; u.cfg.m_treelet_prefetch = true;
  u.last_prefetched_treelet = (uint8_t*)last_prefetched;
  for (uint64_t w = 0; w < n_warps; w++) {
    ref_warp_inst& wi = u.m_current_warps[(unsigned)w];
    wi.uid = (unsigned)w; wi.m_empty = false; wi.stalled = stalled && stalled[w];
    for (int l = 0; l < 32; l++) {
      const uint64_t r = ray_ids[32 * w + l];
      if (r == ~0ull) continue;
      for (uint64_t k = offsets[r] + (front ? front[r] : 0); k < offsets[r + 1]; k++)
        wi.th[l].RT_mem_accesses.push_back(RTMemoryTransactionRecord(txns[k].address, txns[k].size, (TransactionType)txns[k].type));
    }
  }
  warp_inst_t picked;
  u.schedule_next_warp(picked);
  return picked.empty() ? -1 : (int64_t)picked.get_uid();
}

void ref_set_treelet_metadata(uint64_t base, unsigned per_treelet_size) {
  VulkanRayTracing::treelet_metadata = (void*)base; VulkanRayTracing::per_treelet_metadata_size = per_treelet_size;
}

}  // extern "C"
