"""ctypes binding of libvsrt_scene.so: synthetic scenes in the GEN_RT_BVH wire format and ray generators
(include/vsrt_scene.h).  CPU-only tooling shared by tests and bench.py."""
import ctypes
import os
import numpy as np
from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(os.path.dirname(_HERE), "libvsrt_scene.so")


class SceneDesc(ctypes.Structure):
    _fields_ = [("seed", ctypes.c_uint64), ("n_triangles", ctypes.c_uint64), ("n_blas", ctypes.c_uint32),
                ("n_instances", ctypes.c_uint32), ("kind", ctypes.c_uint32), ("flags", ctypes.c_uint32),
                ("max_fanout", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]


UNIFORM, CLUSTERED = 0, 1
F_HOLES, F_TRANSFORMS, F_PROCEDURAL = 1, 2, 4
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise RuntimeError("libvsrt_scene.so not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = ctypes.CDLL(_LIB)
        L.vsrt_scene_build.argtypes = [ctypes.POINTER(SceneDesc), ctypes.POINTER(ctypes.c_void_p)]
        L.vsrt_scene_free.argtypes = [ctypes.c_void_p]
        L.vsrt_scene_arena.restype = ctypes.c_void_p
        L.vsrt_scene_arena.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.vsrt_scene_n_blas.restype = ctypes.c_uint32
        L.vsrt_scene_n_blas.argtypes = [ctypes.c_void_p]
        L.vsrt_scene_blas_offset.restype = ctypes.c_uint64
        L.vsrt_scene_blas_offset.argtypes = [ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64)]
        L.vsrt_scene_n_nodes.restype = ctypes.c_uint64
        L.vsrt_scene_n_nodes.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64),
                                         ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint32)]
        L.vsrt_scene_triangles.restype = ctypes.c_void_p
        L.vsrt_scene_triangles.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint64)]
        L.vsrt_arena_validate.argtypes = [ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_char_p,
                                          ctypes.c_uint32]
        L.vsrt_rays_primary.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64,
                                        ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
        L.vsrt_rays_primary_tiled.argtypes = [ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_uint64,
                                              ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                                              ctypes.c_uint32, ctypes.c_void_p]
        L.vsrt_rays_random.argtypes = [ctypes.c_uint64, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_uint64,
                                       ctypes.c_void_p]
        L.vsrt_rays_bounce_scene.argtypes = [ctypes.c_void_p]
        L.vsrt_rays_bounce.restype = ctypes.c_uint64
        L.vsrt_rays_bounce.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint64,
                                       ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
        _lib = L
    return _lib


class Arena:
    """A host arena: bytes at a 64-byte aligned address, TLAS header offset and BLAS (offset, size) list."""

    def __init__(self, data, tlas_offset=0, blas=()):
        n = len(data)
        self._buf = np.zeros(n + 64, dtype=np.uint8)
        shift = (-self._buf.ctypes.data) % 64
        self.bytes = self._buf[shift:shift + n]
        self.bytes[:] = np.frombuffer(bytes(data), dtype=np.uint8) if not isinstance(data, np.ndarray) else data
        self.base = self.bytes.ctypes.data
        self.size = n
        self.tlas_offset = tlas_offset
        self.blas = list(blas)

    @property
    def tlas(self):
        return self.base + self.tlas_offset

    def validate(self):
        msg = ctypes.create_string_buffer(256)
        rc = lib().vsrt_arena_validate(self.base, self.size, self.tlas_offset, msg, 256)
        return rc, msg.value.decode()


class Scene(Arena):
    def __init__(self, n_triangles, seed=0x5EED0001, n_blas=1, n_instances=0, kind=UNIFORM, flags=0, max_fanout=6):
        L = lib()
        d = SceneDesc(seed, n_triangles, n_blas, max(n_instances, n_blas), kind, flags, max_fanout, 0)
        h = ctypes.c_void_p()
        rc = L.vsrt_scene_build(ctypes.byref(d), ctypes.byref(h))
        if rc != 0:
            raise RuntimeError("vsrt_scene_build failed: %d" % rc)
        self._h = h
        sz = ctypes.c_uint64()
        p = L.vsrt_scene_arena(h, ctypes.byref(sz))
        self.size = sz.value
        self.base = p
        self.bytes = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(self.size,))
        self.tlas_offset = 0
        self.blas = []
        for b in range(L.vsrt_scene_n_blas(h)):
            bs = ctypes.c_uint64()
            off = L.vsrt_scene_blas_offset(h, b, ctypes.byref(bs))
            self.blas.append((off, bs.value))
        ni, nl, dp = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_uint32()
        L.vsrt_scene_n_nodes(h, ctypes.byref(ni), ctypes.byref(nl), ctypes.byref(dp))
        self.n_internal, self.n_leaves, self.depth = ni.value, nl.value, dp.value
        self.n_triangles = n_triangles

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().vsrt_scene_free(self._h)
                self._h = None
        except Exception:
            pass

    def bounce(self, rays, hits, seed, bounce, flags):
        L = lib()
        L.vsrt_rays_bounce_scene(self._h)
        out = np.zeros(len(rays), dtype=_abi.RAY)
        n = L.vsrt_rays_bounce(_abi.ptr(rays), _abi.ptr(np.ascontiguousarray(hits)), len(rays), seed, bounce, flags,
                               _abi.ptr(out))
        L.vsrt_rays_bounce_scene(None)
        return out[:n].copy()


def rays_primary(width, height, spp=1, seed=1, flags=0, first=0, count=None, tile=None):
    """Camera rays [first, first + count) of a width x height x spp frame.  tile=(w, h): ids walk the frame in w x h pixel
    tiles, the order of the reference's raygen launch (one-warp CTAs of 8 x 4 pixels, vulkan_ray_tracing.cc:3505);
    None = scanlines.  Either way the same set of rays."""
    total = width * height * spp
    count = total - first if count is None else count
    out = np.zeros(count, dtype=_abi.RAY)
    tw, th = tile if tile else (0, 0)
    lib().vsrt_rays_primary_tiled(width, height, spp, seed, flags, first, count, tw, th, _abi.ptr(out))
    return out


def rays_random(n, seed=1, flags=0, first=0):
    out = np.zeros(n, dtype=_abi.RAY)
    lib().vsrt_rays_random(seed, flags, first, n, _abi.ptr(out))
    return out
