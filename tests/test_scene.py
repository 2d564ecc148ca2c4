"""Scene writer + validator (CPU)."""
import numpy as np
from vsrt import scene as sc, _abi
import helpers


def test_scene_is_valid_and_deterministic():
    a = sc.Scene(5000, seed=1, n_blas=2, n_instances=5, flags=sc.F_TRANSFORMS | sc.F_HOLES)
    b = sc.Scene(5000, seed=1, n_blas=2, n_instances=5, flags=sc.F_TRANSFORMS | sc.F_HOLES)
    assert a.validate() == (0, "") and np.array_equal(a.bytes, b.bytes)
    assert a.n_leaves == 5000 + 5 and a.size % 64 == 0 and a.base % 64 == 0


def test_validator_catches_reference_asserts():
    a = helpers.kat_arena()
    assert a.validate()[0] == 0
    a.bytes[448 + 12] = 1                       # PrimitiveIndex1Delta != 0 (reference assert :2108)
    rc, msg = a.validate()
    assert rc == -6 and "PrimitiveIndex1Delta" in msg
    a = helpers.kat_arena()
    a.bytes[64 + 22] = 2 | (4 << 2)             # a quad directly under a TLAS node (reference assert :1820)
    assert a.validate()[0] == -6
    a = helpers.kat_arena()
    a.bytes[192 + 64:192 + 72] = 0              # BVHAddress == 0 (reference assert :1900)
    assert a.validate()[0] == -6


def test_quantised_boxes_are_conservative():
    """Every triangle vertex lies inside the decoded box of the leaf slot that holds it."""
    s = sc.Scene(3000, seed=5)
    buf = s.bytes
    off, _ = s.blas[0]
    root = off + int(buf[off:off + 8].view(np.uint64)[0])
    stack = [root]; checked = 0
    while stack:
        n = stack.pop()
        w = buf[n:n + 64]
        org = w[0:12].view(np.float32); co = int(w[12:16].view(np.int32)[0])
        ex = w[18:21].view(np.int8).astype(np.int32)
        child = n + co * 64
        for i in range(6):
            info = int(w[22 + i]) & 0x3f
            if info & 3:
                lo = np.array([org[a] + np.ldexp(np.float32(w[28 + 12 * a + i]), int(ex[a]) - 8) for a in range(3)], np.float32)
                hi = np.array([org[a] + np.ldexp(np.float32(w[28 + 12 * a + 6 + i]), int(ex[a]) - 8) for a in range(3)], np.float32)
                if info >> 2 == 0:
                    stack.append(child)
                else:
                    v = buf[child + 16:child + 52].view(np.float32).reshape(3, 3)
                    assert np.all(v >= lo) and np.all(v <= hi)
                    checked += 1
            child += (info & 3) * 64
    assert checked == 3000


def test_ray_generators():
    r = sc.rays_primary(32, 16)
    assert len(r) == 512 and np.allclose(np.linalg.norm(r["direction"], axis=1), 1, atol=1e-6)
    a = sc.rays_primary(32, 16, spp=2, seed=3, first=100, count=50)
    b = sc.rays_primary(32, 16, spp=2, seed=3)[100:150]
    assert np.array_equal(a, b)       # sharding a frame by ray index gives the same rays
    q = sc.rays_random(100, seed=2, first=10); p = sc.rays_random(110, seed=2)[10:]
    assert np.array_equal(q, p)


def test_primary_rays_in_raygen_tile_order():
    """tile=(8, 4): the frame's ray ids walk the image in 8 x 4 pixel tiles, the order in which the reference's raygen launch reaches
    traceRay (one-warp CTAs of 8 x 4 pixels, warp_pixel_mapping WARP_8X4, vulkan_ray_tracing.cc:3505).  Same rays as the scanline
    order, permuted (jitter keyed by the pixel); windows of the id space agree with the whole; a frame the tile does not divide
    falls back to scanlines."""
    W, H, spp = 64, 32, 2
    scan = sc.rays_primary(W, H, spp=spp, seed=3)
    tiled = sc.rays_primary(W, H, spp=spp, seed=3, tile=(8, 4))
    ids = np.arange(W * H * spp)
    sm, p = ids // (W * H), ids % (W * H)
    t, i = p // 32, p % 32
    x, y = (t % (W // 8)) * 8 + i % 8, (t // (W // 8)) * 4 + i // 8
    assert np.array_equal(tiled, scan[sm * W * H + y * W + x])
    assert not np.array_equal(tiled, scan)
    # a warp's 32 consecutive ids are one 8 x 4 tile: their directions span 8 pixels in x and 4 in y
    assert np.array_equal(sc.rays_primary(W, H, spp=spp, seed=3, tile=(8, 4), first=96, count=200), tiled[96:296])
    assert np.array_equal(sc.rays_primary(60, 30, tile=(8, 4)), sc.rays_primary(60, 30))
    assert np.array_equal(sc.rays_primary(W, H, tile=(W, 1)), sc.rays_primary(W, H))      # a tile as wide as the frame IS the scanline order
