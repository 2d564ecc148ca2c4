"""numpy/ctypes mirrors of the C-ABI structs in include/vsrt.h (layouts must match byte for byte)."""
import ctypes
import numpy as np

RAY = np.dtype([("origin", "<f4", 3), ("tmin", "<f4"), ("direction", "<f4", 3), ("tmax", "<f4"),
                ("ray_flags", "<u4"), ("cull_mask", "<u4"), ("sbt_record_offset", "<u4"),
                ("sbt_record_stride", "<u4"), ("miss_index", "<u4")])
HIT = np.dtype([("hit_geometry", "<u4"), ("world_min_thit", "<f4"), ("primitive_index", "<u4"),
                ("geometry_index", "<u4"), ("instance_index", "<u4"), ("barycentric", "<f4", 3),
                ("intersection_point", "<f4", 3), ("n_all_hits", "<u4"), ("instance_leaf_address", "<u8")])
TXN = np.dtype([("address", "<u8"), ("size", "<u4"), ("type", "<u4")])
assert RAY.itemsize == 52 and HIT.itemsize == 56 and TXN.itemsize == 16

COUNTER_FIELDS = (["mem_access_type_%d" % i for i in range(9)] +
                  ["num_hits", "num_any_hits", "n_anyhit_rays", "n_closesthit_rays", "tot_nodes_per_ray",
                   "accessed_data_size", "ray_count", "max_nodes_per_ray", "max_tree_depth"])
N_SUM, N_MAX = 16, 2
assert len(COUNTER_FIELDS) == N_SUM + N_MAX

FLAG_OPAQUE = 0x1
FLAG_TERMINATE_ON_FIRST_HIT = 0x4
MODE_DFS, MODE_TREELET = 0, 1
RAY_ORDER_AUTO, RAY_ORDER_INPUT, RAY_ORDER_SORTED = 0, 1, 2

ERRORS = {0: "OK", -1: "INVALID", -2: "NO_DEVICE", -3: "CUDA", -4: "CAPACITY", -5: "UNKNOWN_AS",
          -6: "BAD_BVH", -7: "STACK_OVERFLOW", -8: "BUDGET", -9: "UNSUPPORTED", -10: "COMM"}
COMM_ID_BYTES = 128


class Config(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int32), ("max_treelet_size", ctypes.c_uint32),
                ("treelet_based_traversal", ctypes.c_uint32), ("remap_to_treelet_layout", ctypes.c_uint32),
                ("treelet_remap_stride", ctypes.c_uint32), ("load_treelet_metadata", ctypes.c_uint32),
                ("stack_entries", ctypes.c_uint32), ("ray_order", ctypes.c_uint32)]


class TreeletInfo(ctypes.Structure):
    _fields_ = [("n_treelets", ctypes.c_uint64), ("n_list_entries", ctypes.c_uint64),
                ("n_mapped_nodes", ctypes.c_uint64), ("total_bvh_size", ctypes.c_uint64),
                ("form_ms", ctypes.c_double), ("scratch_bytes", ctypes.c_uint64)]


class DeviceResults(ctypes.Structure):
    _fields_ = [("hits", ctypes.c_void_p), ("trace_offsets", ctypes.c_void_p), ("txns", ctypes.c_void_p),
                ("treelet_ids", ctypes.c_void_p), ("n_rays", ctypes.c_uint64), ("n_txn", ctypes.c_uint64),
                ("algorithmic_bytes", ctypes.c_uint64), ("traverse_ms", ctypes.c_float),
                ("scan_ms", ctypes.c_float), ("compact_ms", ctypes.c_float),
                ("kernel_launches", ctypes.c_uint32), ("order_ms", ctypes.c_float)]


class PrefetchConfig(ctypes.Structure):
    _fields_ = [("heuristic", ctypes.c_uint32), ("load_treelet_metadata", ctypes.c_uint32), ("threshold", ctypes.c_double),
                ("treelet_metadata_base", ctypes.c_uint64)]


# vsrt_prefetch_decision
PDEC = np.dtype([("treelet_root", np.uint64), ("votes", np.uint32), ("total", np.uint32), ("submit", np.uint32), ("n_nodes", np.uint32),
                 ("first_node", np.uint32), ("num_nodes", np.uint32)])


# vsrt_table_event, vsrt_store_txn
TEV = np.dtype([("table", np.uint32), ("shader_counter", np.uint32), ("hit_group_index", np.uint32), ("primitive_id", np.uint32),
                ("instance_id", np.uint32), ("tid", np.uint32), ("record", np.uint32), ("reserved", np.uint32)])
STORE = np.dtype([("address", np.uint64), ("size", np.uint32), ("type", np.uint32)])
# vsrt_coalescing_event
CEV = np.dtype([("row", np.uint32), ("appended", np.uint32), ("n_loads", np.uint32), ("first_new_load", np.uint32)])


def ptr(a):
    """ctypes void* of a numpy array (or None)."""
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class _PackedSpan(ctypes.Structure):
    _fields_ = [("host", ctypes.c_uint64), ("slot0", ctypes.c_uint32), ("n_slots", ctypes.c_uint32)]


class PackedLayout(ctypes.Structure):   # vsrt_packed_layout
    _fields_ = [("device_delta", ctypes.c_int64), ("n_spans", ctypes.c_uint32), ("reserved", ctypes.c_uint32), ("spans", _PackedSpan * 8)]

