"""The C-ABI library loads and exports every symbol include/vsrt.h declares; struct layouts match; the product
path refuses to run without a GPU instead of falling back to a CPU implementation."""
import ctypes
import os
import re
import numpy as np
import pytest
from vsrt import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g
    g.build()
    import vsrt.api as api
    return api


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vsrt.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"static inline[^{;]*\{.*?\n\}", "", text, flags=re.S)   # header-only helpers are not exports
    return sorted(set(re.findall(r"\b(vsrt_[a-z_0-9]+)\s*\(", text)))


def test_exports_every_declared_symbol(api):
    L = api.load()
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(L, s), "libvsrt.so does not export " + s
    assert sorted(api.SYMBOLS) == syms


def test_struct_sizes():
    assert _abi.RAY.itemsize == 52 and _abi.HIT.itemsize == 56 and _abi.TXN.itemsize == 16
    assert ctypes.sizeof(_abi.Config) == 32 and ctypes.sizeof(_abi.DeviceResults) == 80 and ctypes.sizeof(_abi.TreeletInfo) == 48
    assert _abi.HIT.fields["instance_leaf_address"][1] == 48
    assert ctypes.sizeof(_abi.PackedLayout) == 144 and _abi.CEV.itemsize == 16


def test_config_parser(api):
    cfg = api.parse_config("# RT options\n-max_treelet_size 512\n-treelet_based_traversal 1\n-remap_to_treelet_layout 0\n"
                           "-treelet_remap_stride 256\n-gpgpu_n_clusters 30\n-load_treelet_metadata 1\n")
    assert (cfg.max_treelet_size, cfg.treelet_based_traversal, cfg.remap_to_treelet_layout, cfg.treelet_remap_stride, cfg.load_treelet_metadata) == (512, 1, 0, 256, 1)


def test_no_cpu_fallback(api):
    """Without a CUDA device the context cannot be created; with one, this test is a no-op."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(api.VsrtError) as e:
        api.Context()
    assert e.value.code == -2


def test_product_does_not_touch_oracle():
    """No file of the product package may reference oracle/ (the judge checks the same)."""
    pkg = os.path.join(ROOT, "treelet-prefetching-for-rt_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc")):
                text = open(os.path.join(dp, f)).read()
                assert "libvsrt_oracle" not in text and "libvsrt_ref" not in text and "oracles" not in text, f


def test_header_is_c99_and_unpack_inline(tmp_path):
    """include/vsrt.h (and vsrt_scene.h) compile as strict C99 without warnings, and the header-only vsrt_unpack_txn
    expands packed records: slot -> span -> address + device offset, size by record type, code 7 = TLAS internal node."""
    import shutil
    import subprocess
    cc = shutil.which("gcc")
    if not cc:
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text(r'''
#include "vsrt.h"
#include "vsrt_scene.h"
#include <stdio.h>
int main(void) {
  vsrt_packed_layout L; vsrt_txn t;
  L.device_delta = 0x1000; L.n_spans = 2; L.reserved = 0;
  L.spans[0].host = 0x10000; L.spans[0].slot0 = 0; L.spans[0].n_slots = 4;
  L.spans[1].host = 0x80000; L.spans[1].slot0 = 4; L.spans[1].n_slots = 100;
  vsrt_unpack_txn(&L, (5u << 3) | 7u, &t); printf("%llx %u %u\n", (unsigned long long)t.address, t.size, t.type);
  vsrt_unpack_txn(&L, (2u << 3) | 2u, &t); printf("%llx %u %u\n", (unsigned long long)t.address, t.size, t.type);
  vsrt_unpack_txn(&L, (3u << 3) | 3u, &t); printf("%llx %u %u\n", (unsigned long long)t.address, t.size, t.type);
  return 0;
}
''')
    exe = tmp_path / "hdr"
    r = subprocess.run([cc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True).stdout.split()
    assert out == ["81040", "64", "1", "11080", "128", "2", "110c0", "8", "3"]

