"""ctypes face of oracle/liboracle.so — the CPU restatement of Echo's hot path. TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the package."""
import ctypes
import os
import subprocess

import numpy as np

from echorenderer_b200 import structs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
_lib = None

BXDF_LAMBERTIAN_REFLECTION, BXDF_LAMBERTIAN, BXDF_OREN_NAYAR, BXDF_SPECULAR_REFLECTION_REAL, BXDF_SPECULAR_REFLECTION_COMPLEX, \
    BXDF_SPECULAR_TRANSMISSION, BXDF_SPECULAR_FRESNEL, BXDF_GLOSSY_REFLECTION_REAL, BXDF_GLOSSY_REFLECTION_COMPLEX, \
    BXDF_GLOSSY_TRANSMISSION, BXDF_COATED_LAMBERTIAN = range(11)

FM_MAX0, FM_CLAMP01, FM_CLAMP11, FM_CLAMP_EPSILON, FM_ABS, FM_SQRT0, FM_SQRTR0, FM_ONE_MINUS2, FM_IDENTITY, FM_FMA, \
    FM_POSITIVE, FM_ALMOST_ZERO, FM_MIN, FM_MAX = range(14)


def build():
    subprocess.run(["make", "-s", "-C", ORACLE_DIR], check=True)


def library():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(ORACLE_DIR, "liboracle.so")
    if not os.path.exists(path):
        build()
    lib = ctypes.CDLL(path)
    p, u32, u64, i32, f32 = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int32, ctypes.c_float
    lib.oracle_scene_create.restype = p
    lib.oracle_scene_destroy.argtypes = [p]
    lib.oracle_scene_set_qbvh.argtypes = [p, p, u32, u32]
    lib.oracle_scene_set_triangles.argtypes = [p, p, u32]
    lib.oracle_scene_set_spheres.argtypes = [p, p, u32]
    lib.oracle_scene_set_materials.argtypes = [p, p, u32]
    lib.oracle_scene_set_light_tree.argtypes = [p, p, u32, p, p, u32, p, u32]
    lib.oracle_scene_set_infinite.argtypes = [p, p, u32, f32, f32]
    lib.oracle_scene_set_camera.argtypes = [p, p]
    lib.oracle_scene_set_packs.argtypes = [p, p, u32, p, u32]
    lib.oracle_scene_set_textures.argtypes = [p, p, u32, p, u64, p, u32]
    lib.oracle_scene_set_distributions.argtypes = [p, p, u64]
    lib.oracle_infinite_light.argtypes = [p, u32, p, p, p]
    lib.oracle_texture_sample.argtypes = [p, u32, p, u64, p]
    lib.oracle_atan2.argtypes = [f32, f32]
    lib.oracle_atan2.restype = f32
    lib.oracle_asin.argtypes = [f32]
    lib.oracle_asin.restype = f32
    lib.oracle_acos.argtypes = [f32]
    lib.oracle_acos.restype = f32
    lib.oracle_trace_batch_hierarchy.argtypes = [p, p, p, u64, p, p, i32, i32]
    lib.oracle_occlude_batch_hierarchy.argtypes = [p, p, p, u64, p, i32, i32]
    lib.oracle_trace_batch.argtypes = [p, p, u64, p, p, i32]
    lib.oracle_occlude_batch.argtypes = [p, p, u64, p, p, i32]
    lib.oracle_trace_linear_batch.argtypes = [p, p, u64, p, i32]
    lib.oracle_occlude_linear_batch.argtypes = [p, p, u64, p, i32]
    lib.oracle_render_tiles.argtypes = [p, p, p, u32, p, p, i32]
    lib.oracle_evaluate_samples.argtypes = [p, p, p, p, u64, p, i32]
    lib.oracle_evaluate_samples4.argtypes = [p, p, p, p, u64, p, i32, i32]
    lib.oracle_scene_set_bound_radius.argtypes = [p, f32]
    lib.oracle_spawn_rays.argtypes = [p, p, p, p, u64, p]
    lib.oracle_fastmath.argtypes = [i32, f32, f32, f32]
    lib.oracle_fastmath.restype = f32
    lib.oracle_sincos.argtypes = [f32, p, p]
    lib.oracle_kahan_sum.argtypes = [p, u64]
    lib.oracle_kahan_sum.restype = f32
    lib.oracle_orthonormal.argtypes = [p, p, p]
    lib.oracle_sphere_intersect.argtypes = [p, p, p, i32, p]
    lib.oracle_sphere_intersect.restype = f32
    lib.oracle_sphere_occlude.argtypes = [p, p, p, f32, i32]
    lib.oracle_sphere_occlude.restype = i32
    lib.oracle_triangle_intersect.argtypes = [p, p, p, p]
    lib.oracle_triangle_intersect.restype = f32
    lib.oracle_triangle_occlude.argtypes = [p, p, p, f32]
    lib.oracle_triangle_occlude.restype = i32
    lib.oracle_geometry_sample.argtypes = [p, u32, p, p, p]
    lib.oracle_geometry_sample.restype = i32
    lib.oracle_geometry_pdf.argtypes = [p, u32, p, p]
    lib.oracle_geometry_pdf.restype = f32
    lib.oracle_distribution1d.argtypes = [p, i32, f32, p]
    lib.oracle_cylindrical_to_uv.argtypes = [p, p]
    lib.oracle_cylindrical_to_direction.argtypes = [p, p]
    lib.oracle_set_failed_pick_keeps_mis.argtypes = [i32]
    lib.oracle_set_failed_pick_keeps_mis.restype = None
    lib.oracle_light_pick.argtypes = [p, p, p, f32, p]
    lib.oracle_light_pick.restype = u32
    lib.oracle_light_mass.argtypes = [p, u32, p, p]
    lib.oracle_light_mass.restype = f32
    lib.oracle_bxdf_batch.argtypes = [i32, p, p, p, u64, p, p, p]
    lib.oracle_accumulate.argtypes = [p, u64, p]
    lib.oracle_sample_value.argtypes = [u32, u32, u32, u32]
    lib.oracle_sample_value.restype = f32
    _lib = lib
    return lib


def ptr(array):
    return ctypes.c_void_p(array.ctypes.data) if array is not None and array.size else ctypes.c_void_p(0)


def f32(values):
    return np.ascontiguousarray(values, dtype=np.float32)


class OracleScene:
    """The oracle's PreparedScene, fed with the same flattened arrays as the device library."""

    def __init__(self, prepared):
        lib = library()
        self.lib = lib
        self.prepared = prepared
        self.handle = ctypes.c_void_p(lib.oracle_scene_create())
        d = prepared.description
        lib.oracle_scene_set_qbvh(self.handle, ptr(prepared.nodes), len(prepared.nodes), prepared.max_depth)
        lib.oracle_scene_set_triangles(self.handle, ptr(prepared.triangles), len(prepared.triangles))
        lib.oracle_scene_set_spheres(self.handle, ptr(prepared.spheres), len(prepared.spheres))
        lib.oracle_scene_set_materials(self.handle, ptr(prepared.materials), len(prepared.materials))
        if prepared.textures is not None:
            lib.oracle_scene_set_textures(self.handle, ptr(prepared.textures), len(prepared.textures), ptr(prepared.texels), len(prepared.texels),
                                          ptr(prepared.material_textures), len(prepared.material_textures))
        if prepared.distributions is not None:
            lib.oracle_scene_set_distributions(self.handle, ptr(prepared.distributions), len(prepared.distributions))
        if prepared.packs is not None:
            lib.oracle_scene_set_packs(self.handle, ptr(prepared.packs), len(prepared.packs), ptr(prepared.instances), len(prepared.instances))
        lib.oracle_scene_set_light_tree(self.handle, ptr(prepared.light_nodes), len(prepared.light_nodes), ptr(prepared.emitter_tokens),
                                        ptr(prepared.emitter_bitpaths), len(prepared.emitter_tokens), ptr(prepared.point_lights), len(prepared.point_lights))
        lib.oracle_scene_set_infinite(self.handle, ptr(d.infinite_lights), len(d.infinite_lights), prepared.infinite_threshold, prepared.infinite_pdf)
        lib.oracle_scene_set_camera(self.handle, ptr(d.camera))
        lib.oracle_scene_set_bound_radius(self.handle, prepared.bound_radius)

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.oracle_scene_destroy(self.handle)
            self.handle = None

    def trace(self, rays, threads=0, count_visits=False):
        rays = np.ascontiguousarray(rays, dtype=structs.RAY)
        hits = np.zeros(len(rays), dtype=structs.HIT)
        counters = np.zeros(3, dtype=np.uint64)
        self.lib.oracle_trace_batch(self.handle, ptr(rays), len(rays), ptr(hits), ptr(counters) if count_visits else None, threads)
        return (hits, counters) if count_visits else hits

    def occlude(self, rays, threads=0, count_visits=False):
        rays = np.ascontiguousarray(rays, dtype=structs.RAY)
        occluded = np.zeros(len(rays), dtype=np.uint8)
        counters = np.zeros(3, dtype=np.uint64)
        self.lib.oracle_occlude_batch(self.handle, ptr(rays), len(rays), ptr(occluded), ptr(counters) if count_visits else None, threads)
        return (occluded, counters) if count_visits else occluded

    def trace_hierarchy(self, rays, ignore_layers=None, linear=False, threads=0):
        rays = np.ascontiguousarray(rays, dtype=structs.RAY)
        ignore = None if ignore_layers is None else np.ascontiguousarray(ignore_layers, dtype=structs.TOKEN_HIERARCHY)
        hits = np.zeros(len(rays), dtype=structs.HIT)
        layers = np.zeros(len(rays), dtype=structs.TOKEN_HIERARCHY)
        self.lib.oracle_trace_batch_hierarchy(self.handle, ptr(rays), ptr(ignore), len(rays), ptr(hits), ptr(layers), int(linear), threads)
        return hits, layers

    def occlude_hierarchy(self, rays, ignore_layers=None, linear=False, threads=0):
        rays = np.ascontiguousarray(rays, dtype=structs.RAY)
        ignore = None if ignore_layers is None else np.ascontiguousarray(ignore_layers, dtype=structs.TOKEN_HIERARCHY)
        occluded = np.zeros(len(rays), dtype=np.uint8)
        self.lib.oracle_occlude_batch_hierarchy(self.handle, ptr(rays), ptr(ignore), len(rays), ptr(occluded), int(linear), threads)
        return occluded

    def trace_linear(self, rays, threads=0):
        rays = np.ascontiguousarray(rays, dtype=structs.RAY)
        hits = np.zeros(len(rays), dtype=structs.HIT)
        self.lib.oracle_trace_linear_batch(self.handle, ptr(rays), len(rays), ptr(hits), threads)
        return hits

    def occlude_linear(self, rays, threads=0):
        rays = np.ascontiguousarray(rays, dtype=structs.RAY)
        occluded = np.zeros(len(rays), dtype=np.uint8)
        self.lib.oracle_occlude_linear_batch(self.handle, ptr(rays), len(rays), ptr(occluded), threads)
        return occluded

    def render_tiles(self, params, tile_xy, threads=0):
        tile_xy = np.ascontiguousarray(tile_xy, dtype=np.int32)
        tile_size = int(params["tileSize"][0])
        out = np.zeros((len(tile_xy), tile_size, tile_size, 4), dtype=np.float32)
        stats = np.zeros(1, dtype=structs.STATS)
        self.lib.oracle_render_tiles(self.handle, ptr(params), ptr(tile_xy), len(tile_xy), ptr(out), ptr(stats), threads)
        return out, stats

    def evaluate_samples(self, params, pixel_xy, sample_index, threads=0):
        pixel_xy = np.ascontiguousarray(pixel_xy, dtype=np.int32)
        sample_index = np.ascontiguousarray(sample_index, dtype=np.uint32)
        out = np.zeros((len(sample_index), 3), dtype=np.float32)
        self.lib.oracle_evaluate_samples(self.handle, ptr(params), ptr(pixel_xy), ptr(sample_index), len(sample_index), ptr(out), threads)
        return out

    def spawn_rays(self, params, pixel_xy, sample_index):
        pixel_xy = np.ascontiguousarray(pixel_xy, dtype=np.int32)
        sample_index = np.ascontiguousarray(sample_index, dtype=np.uint32)
        out = np.zeros(len(sample_index), dtype=structs.RAY)
        self.lib.oracle_spawn_rays(self.handle, ptr(params), ptr(pixel_xy), ptr(sample_index), len(sample_index), ptr(out))
        return out

    def light_pick(self, position, normal, sample):
        pdf = ctypes.c_float()
        token = self.lib.oracle_light_pick(self.handle, ptr(f32(position)), ptr(f32(normal)), float(sample), ctypes.byref(pdf))
        return int(token), float(pdf.value)

    def light_mass(self, token, position, normal):
        return float(self.lib.oracle_light_mass(self.handle, int(token), ptr(f32(position)), ptr(f32(normal))))

    def geometry_sample(self, token, origin, sample):
        out = np.zeros(7, dtype=np.float32)
        ok = self.lib.oracle_geometry_sample(self.handle, int(token), ptr(f32(origin)), ptr(f32(sample)), ptr(out))
        return bool(ok), out[:3].copy(), out[3:6].copy(), float(out[6])

    def geometry_pdf(self, token, origin, incident):
        return float(self.lib.oracle_geometry_pdf(self.handle, int(token), ptr(f32(origin)), ptr(f32(incident))))


def distribution1d(cdf, sample):
    """(Sample value, Sample pdf, ProbabilityDensity(value), Pick index, Pick pdf) of a DiscreteDistribution1D given by its cdf."""
    cdf = np.ascontiguousarray(cdf, dtype=np.float32)
    out = np.zeros(5, dtype=np.float32)
    library().oracle_distribution1d(ptr(cdf), len(cdf), float(sample), ptr(out))
    return float(out[0]), float(out[1]), float(out[2]), int(out[3]), float(out[4])


def cylindrical_to_uv(direction):
    out = np.zeros(2, dtype=np.float32)
    library().oracle_cylindrical_to_uv(ptr(f32(direction)), ptr(out))
    return out


def cylindrical_to_direction(uv):
    out = np.zeros(3, dtype=np.float32)
    library().oracle_cylindrical_to_direction(ptr(f32(uv)), ptr(out))
    return out


def fastmath(op, a, b=0.0, c=0.0):
    return np.float32(library().oracle_fastmath(op, float(a), float(b), float(c)))


def sincos(radians):
    s, c = ctypes.c_float(), ctypes.c_float()
    library().oracle_sincos(float(radians), ctypes.byref(s), ctypes.byref(c))
    return np.float32(s.value), np.float32(c.value)


def bxdf_params(alpha=(1.0, 1.0), real=None, complex_=None, roughness=None, coated=None):
    """coated = (albedo rgb, reflectance) for CoatedLambertianReflection (together with real=)."""
    params = np.zeros(11, dtype=np.float32)
    params[0:2] = alpha
    if roughness is not None:
        params[0] = roughness
    if real is not None:
        params[2:4] = real
    if complex_ is not None:
        params[2:11] = np.asarray(complex_, dtype=np.float32).reshape(-1)
    if coated is not None:
        params[5:8] = coated[0]
        params[4] = coated[1]
    return params


def bxdf_batch(kind, params, outgoing, samples):
    outgoing, samples = f32(outgoing).reshape(-1, 3), f32(samples).reshape(-1, 2)
    n = len(outgoing)
    sampled, evaluated, inverse = np.zeros((n, 8), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
    library().oracle_bxdf_batch(kind, ptr(f32(params)), ptr(outgoing), ptr(samples), n, ptr(sampled), ptr(evaluated), ptr(inverse))
    return sampled, evaluated, inverse


def accumulate(samples):
    samples = f32(samples).reshape(-1, 4)
    out = np.zeros(6, dtype=np.float32)
    library().oracle_accumulate(ptr(samples), len(samples), ptr(out))
    return out[:4].copy(), float(out[4]), int(out[5])


def kahan_sum(values):
    values = f32(values)
    return np.float32(library().oracle_kahan_sum(ptr(values), len(values)))


def orthonormal(axis_z, direction):
    out = np.zeros(15, dtype=np.float32)
    library().oracle_orthonormal(ptr(f32(axis_z)), ptr(f32(direction)), ptr(out))
    return out.reshape(5, 3)
