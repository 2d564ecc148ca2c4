"""Acceleration-structure dump files (VulkanRayTracing::dump_AS, vulkan_ray_tracing.cc:4455-4558): writer, reader and the
image registration that walks the TLAS.  The reference ships the dumper only (the loader lives in the Mesa-side external
launcher) and no sample files, so the checks are round trips: same bytes back, same traces from the reloaded image."""
import os
import struct
import numpy as np
import pytest
from vsrt import scene as sc, _abi
import helpers
import oracles


@pytest.fixture(scope="module")
def api():
    import __graft_entry__ as g
    g.build()
    import vsrt.api as api
    return api


def tlas_in_the_middle(s):
    """Rearranged copy of a 2-BLAS scene: [BLAS0 | TLAS | BLAS1] with every instance leaf's BVHAddress re-based, so that the
    dump has both an .asback and an .asfront part."""
    (o0, z0), (o1, z1) = s.blas
    tl = bytes(s.bytes[:o0]); b0 = bytes(s.bytes[o0:o0 + z0]); b1 = bytes(s.bytes[o1:o1 + z1])
    new_t, new_b1 = z0, z0 + len(tl)
    data = bytearray(b0 + tl + b1)
    moved = {o0: 0, o1: new_b1}
    root = struct.unpack_from("<Q", tl, 0)[0]
    stack = [root]
    while stack:
        off = stack.pop()
        child = off + struct.unpack_from("<i", tl, off + 12)[0] * 64
        for i in range(6):
            info = tl[off + 22 + i] & 0x3f
            sz, ty = info & 3, info >> 2
            if not sz:
                continue
            if ty == 0:
                stack.append(child)
            else:
                rel = struct.unpack_from("<Q", tl, child + 64)[0]
                hdr_old = (child + rel) & 0xFFFFFFFFFFFFFFFF
                struct.pack_into("<Q", data, new_t + child + 64, (moved[hdr_old] - (new_t + child)) & 0xFFFFFFFFFFFFFFFF)
            child += 64 * sz
    return sc.Arena(bytes(data), new_t, [(0, z0), (new_b1, z1)])


def test_writer_matches_the_reference_dumper(tmp_path):
    """vsrt_as_dump_write against the reference's own dump_descriptor_set_for_AS + findOffsetBounds (compiled into oracle/_ref
    from vulkan_ray_tracing.cc:4455-4558, :4901-4945): the four files byte for byte, for a TLAS below its BLASes, between them
    and above them -- the reference's 0 / 20 KiB slack after the last BLAS START included (it truncates a larger last BLAS and
    reads past a smaller one; the arenas here are padded so that the 20 KiB exist).  CPU only: the writer is host code."""
    if not oracles.have_ref():
        pytest.skip("oracle/_ref not built")
    import vsrt.api as api
    orc = oracles.RefOracle()
    s = sc.Scene(1200, seed=21, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    mid = tlas_in_the_middle(s)
    pad = np.zeros(24 * 1024, np.uint8)

    def padded(a):      # same arena with 24 KiB of zeros behind it, so that "last BLAS start + 20 KiB" stays inside the buffer
        return sc.Arena(np.concatenate([a.bytes, pad]), a.tlas_offset, list(a.blas))
    (o0, z0), (o1, z1) = s.blas
    top = sc.Arena(np.concatenate([s.bytes[o0:], s.bytes[:o0]]), s.size - o0, [])      # [BLAS0 | BLAS1 | TLAS]: only backward offsets
    top.blas = [(0, z0), (o1 - o0, z1)]
    for name, arena in (("fwd", padded(s)), ("mid", padded(mid)), ("back", top)):
        above = sorted(off for off, _ in arena.blas if off > arena.tlas_offset)
        desc = (above[0] if above else arena.size) - arena.tlas_offset
        want = orc.dump_as(str(tmp_path / ("ref_" + name)), arena, desc)
        got = str(tmp_path / ("0_0_" + name))
        api.write_as_dump(got, arena, desc_size=desc, back_buffer=0, front_buffer=20 * 1024)
        for ext in (".asmain", ".asback", ".asfront", ".asmetadata"):
            assert os.path.exists(want + ext) == os.path.exists(got + ext), (name, ext)
            if os.path.exists(want + ext):
                assert open(want + ext, "rb").read() == open(got + ext, "rb").read(), (name, ext)
        kinds = (os.path.exists(got + ".asback"), os.path.exists(got + ".asfront"))
        assert kinds == {"fwd": (False, True), "mid": (True, True), "back": (True, False)}[name]


def test_dump_round_trip(api, tmp_path):
    s = sc.Scene(1200, seed=21, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    for name, arena in (("fwd", s), ("mid", tlas_in_the_middle(s))):
        prefix = str(tmp_path / ("0_1_" + name))
        last = max(arena.blas)                               # slack must cover the last BLAS on each side (the reference: 0 / 20 KiB)
        below = [b for b in arena.blas if b[0] < arena.tlas_offset]
        api.write_as_dump(prefix, arena, back_buffer=max(b[1] for b in below) if below else 0, front_buffer=last[1])
        meta = open(prefix + ".asmetadata").read().split(",")
        assert len(meta) == 10 and int(meta[1]) == 1000150000
        assert (os.path.exists(prefix + ".asback"), os.path.exists(prefix + ".asfront")) == ((name == "mid"), True)
        img = api.AsImage(prefix)
        assert img.tlas_offset == arena.tlas_offset and img.size >= arena.size
        assert np.array_equal(img.bytes[:arena.size], arena.bytes)
    with pytest.raises(api.VsrtError):
        api.AsImage(str(tmp_path / "missing"))
    open(str(tmp_path / "bad.asmetadata"), "w").write("64,1,2,3")
    with pytest.raises(api.VsrtError):
        api.AsImage(str(tmp_path / "bad"))


@pytest.mark.gpu
def test_reloaded_image_traces_like_the_original(api, tmp_path):
    s = sc.Scene(1200, seed=21, n_blas=2, n_instances=3, flags=sc.F_TRANSFORMS)
    arena = tlas_in_the_middle(s)
    prefix = str(tmp_path / "0_1")
    api.write_as_dump(prefix, arena, back_buffer=arena.blas[0][1], front_buffer=arena.blas[1][1])
    img = api.AsImage(prefix)
    rays = helpers.mixed_rays(1500, 31)
    orc = oracles.RefOracle() if oracles.have_ref() else oracles.PortOracle()
    orc.register(arena); orc.form(512)
    ctx = api.Context(max_treelet_size=512, device=0)
    assert ctx.register_image(img) == 2
    ctx.form_treelets()
    for mode in (0, 1):
        o, g = orc.trace(mode, rays), ctx.trace(mode, rays)
        assert np.array_equal(o["offsets"], g["offsets"])
        assert np.array_equal(o["txns"]["address"] - np.uint64(arena.tlas), g["txns"]["address"] - np.uint64(img.tlas))
        assert np.array_equal(o["txns"]["type"], g["txns"]["type"]) and np.array_equal(o["txns"]["size"], g["txns"]["size"])
        assert np.array_equal(o["treelet_ids"] - np.uint64(arena.tlas), g["treelet_ids"] - np.uint64(img.tlas))
        assert np.array_equal(o["hits"]["prim"], g["hits"]["primitive_index"])
    ctx.close()
