"""Known answers derived BY HAND from the reference's source, independent of oracle/*.hpp and of the kernels (the
oracle, the golden fixtures and the device code come from one reading of the C#; a shared misreading would pass every
device-vs-oracle test).

Every expected value below is worked out in the comments from the cited C# lines with inputs chosen so that all intermediate
values are exactly representable in binary32 (small integers, halves, quarters, infinities), so the arithmetic can be
checked with pencil and paper and no rounding is involved. They pin the parts of the hot path the reference's own unit tests
leave unpinned: PreparedTriangle.IntersectImpl (both variants), BoxBound4.Intersect's SSE NaN behaviour as seen
through a traversal, the closest-hit / ignore / strict-comparison rules of GeometryCollection + QuadBoundingVolumeHierarchy,
PreparedSphere.Intersect and the Accumulator. `TRACE_CASES` / `OCCLUDE_CASES` are run against the oracle here and against the
device through the C ABI in tests/test_gpu_boundary.py::test_hand_derived_known_answers.

A second, weaker kind of independence: `numpy_accumulator` is a transcription of Summation.cs / Accumulator.cs into numpy
float32 scalar operations, written from the C# alone; the oracle's Accumulator must agree with it bit for bit on random samples.
"""
import ctypes

import numpy as np
import pytest

from echorenderer_b200 import host, scenes, structs
from tests import oracle_lib

INF = float("inf")
F32 = np.float32


def f3(values):
    return np.asarray(values, dtype=np.float32)


# ---------------------------------------------------------------------------------------------------------------------
# PreparedTriangle.IntersectImpl(origin, direction, out uv) — TriangleEntity.cs:204-235
#   cross2 = direction x edge2; det = edge1 . cross2; det == 0 -> miss; detR = 1 / det
#   offset = origin - vertex0; u = (offset . cross2) detR; (u < 0) | (u > 1) -> miss
#   cross1 = offset x edge1;   v = (direction . cross1) detR; (v < 0) | (u + v > 1) -> miss
#   distance = (edge2 . cross1) detR; distance < 0 -> miss
# with a x b = (ay bz - az by, az bx - ax bz, ax by - ay bx).
# Triangle A: vertex0 (0,0,0), edge1 (1,0,0), edge2 (0,1,0)  [the unit right triangle in the plane z = 0]
# ---------------------------------------------------------------------------------------------------------------------
TRIANGLE_A = dict(vertex0=(0, 0, 0), edge1=(1, 0, 0), edge2=(0, 1, 0))

TRIANGLE_CASES = [
    # origin (1/4, 1/4, 1), direction (0,0,-1):
    #   cross2 = (0*0 - (-1)*1, (-1)*0 - 0*0, 0*1 - 0*0) = (1, 0, 0); det = (1,0,0).(1,0,0) = 1; detR = 1
    #   offset = (1/4, 1/4, 1); u = 1/4
    #   cross1 = offset x (1,0,0) = (1/4*0 - 1*0, 1*1 - 1/4*0, 1/4*0 - 1/4*1) = (0, 1, -1/4); v = (0,0,-1).(0,1,-1/4) = 1/4
    #   distance = (0,1,0).(0,1,-1/4) = 1
    ("front face", TRIANGLE_A, (0.25, 0.25, 1), (0, 0, -1), 1.0, (0.25, 0.25)),
    # the same from behind, origin (1/4, 1/4, -2), direction (0,0,1): no back-face culling
    #   cross2 = (0*0 - 1*1, 1*0 - 0*0, 0) = (-1, 0, 0); det = -1; detR = -1; offset = (1/4, 1/4, -2); u = (-1/4)(-1) = 1/4
    #   cross1 = (1/4*0 - (-2)*0, (-2)*1 - 1/4*0, 1/4*0 - 1/4*1) = (0, -2, -1/4); v = ((0,0,1).(0,-2,-1/4))(-1) = 1/4
    #   distance = ((0,1,0).(0,-2,-1/4))(-1) = 2
    ("back face", TRIANGLE_A, (0.25, 0.25, -2), (0, 0, 1), 2.0, (0.25, 0.25)),
    # on the hypotenuse, u + v == 1 exactly: `u + v > 1` is false -> accepted
    ("hypotenuse is inside", TRIANGLE_A, (0.5, 0.5, 1), (0, 0, -1), 1.0, (0.5, 0.5)),
    # on the edge u == 0: `(u < 0) | (u > 1)` is false -> accepted by the triangle itself (the box in front of it is another matter, below)
    ("edge u = 0 is inside", TRIANGLE_A, (0, 0.5, 1), (0, 0, -1), 1.0, (0.0, 0.5)),
    # u = 3/4, v = 1/2: u + v = 5/4 > 1
    ("outside", TRIANGLE_A, (0.75, 0.5, 1), (0, 0, -1), INF, None),
    # in the triangle's plane, direction (1,0,0): cross2 = (0*0 - 0*1, 0*0 - 1*0, 1*1 - 0*0) = (0,0,1); det = (1,0,0).(0,0,1) = 0
    ("parallel", TRIANGLE_A, (-1, 0.25, 0), (1, 0, 0), INF, None),
    # pointing away: u = v = 1/4 as in the first case with detR = -1 ... distance = -1 < 0
    #   cross2 = (0*0 - 1*1, ...) = (-1,0,0); det = -1; u = (1/4)(-1)(-1) = 1/4; cross1 = (0, 1, -1/4); v = ((0,0,1).(0,1,-1/4))(-1) = 1/4
    #   distance = ((0,1,0).(0,1,-1/4))(-1) = -1
    ("behind the origin", TRIANGLE_A, (0.25, 0.25, 1), (0, 0, 1), INF, None),
    # a scaled, tilted triangle: vertex0 (2,0,0), edge1 (0,4,0), edge2 (0,0,4), origin (0,1,1), direction (1,0,0)
    #   cross2 = (0*4 - 0*0, 0*0 - 1*4, 1*0 - 0*0) = (0,-4,0); det = (0,4,0).(0,-4,0) = -16; detR = -1/16
    #   offset = (-2,1,1); u = (-4)(-1/16) = 1/4
    #   cross1 = offset x (0,4,0) = (1*0 - 1*4, 1*0 - (-2)*0, (-2)*4 - 1*0) = (-4, 0, -8); v = ((1,0,0).(-4,0,-8))(-1/16) = 1/4
    #   distance = ((0,0,4).(-4,0,-8))(-1/16) = (-32)(-1/16) = 2
    ("scaled and tilted", dict(vertex0=(2, 0, 0), edge1=(0, 4, 0), edge2=(0, 0, 4)), (0, 1, 1), (1, 0, 0), 2.0, (0.25, 0.25)),
]

# PreparedTriangle.IntersectImpl(origin, direction, travel) — TriangleEntity.cs:237-263: the same products multiplied by sign(det) and
# compared against |det|; the last line is (distance >= 0) & (distance < travel |det|): STRICT in travel
TRIANGLE_OCCLUDE_CASES = [
    ("nearer than travel", TRIANGLE_A, (0.25, 0.25, 1), (0, 0, -1), 1.5, True),
    ("travel ends exactly on the surface", TRIANGLE_A, (0.25, 0.25, 1), (0, 0, -1), 1.0, False),  # 1 < 1 * 1 is false
    ("travel ends before", TRIANGLE_A, (0.25, 0.25, 1), (0, 0, -1), 0.5, False),
    ("scaled: distance 2 |det| 16, travel 2 is not beyond", dict(vertex0=(2, 0, 0), edge1=(0, 4, 0), edge2=(0, 0, 4)), (0, 1, 1), (1, 0, 0), 2.0, False),  # 32 < 2 * 16 false
    ("scaled: travel 2.5", dict(vertex0=(2, 0, 0), edge1=(0, 4, 0), edge2=(0, 0, 4)), (0, 1, 1), (1, 0, 0), 2.5, True),  # 32 < 40
    ("outside", TRIANGLE_A, (0.75, 0.5, 1), (0, 0, -1), 10.0, False),
]


def as_triangle(fields):
    triangle = np.zeros(1, dtype=structs.TRIANGLE)
    for key, value in fields.items():
        triangle[key] = value
    return triangle


@pytest.mark.parametrize("name,fields,origin,direction,distance,uv", TRIANGLE_CASES, ids=[case[0] for case in TRIANGLE_CASES])
def test_triangle_intersect_by_hand(name, fields, origin, direction, distance, uv):
    lib = oracle_lib.library()
    lib.oracle_triangle_intersect.restype = ctypes.c_float
    out = np.zeros(2, dtype=np.float32)
    triangle, origin, direction = as_triangle(fields), f3(origin), f3(direction)  # named: the pointers below must outlive the call
    result = lib.oracle_triangle_intersect(oracle_lib.ptr(triangle), oracle_lib.ptr(origin), oracle_lib.ptr(direction), oracle_lib.ptr(out))
    assert result == distance
    if uv is not None:
        assert tuple(out) == uv


@pytest.mark.parametrize("name,fields,origin,direction,travel,expected", TRIANGLE_OCCLUDE_CASES, ids=[case[0] for case in TRIANGLE_OCCLUDE_CASES])
def test_triangle_occlude_by_hand(name, fields, origin, direction, travel, expected):
    lib = oracle_lib.library()
    lib.oracle_triangle_occlude.restype = ctypes.c_int32
    lib.oracle_triangle_occlude.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float]
    triangle, origin, direction = as_triangle(fields), f3(origin), f3(direction)
    assert bool(lib.oracle_triangle_occlude(oracle_lib.ptr(triangle), oracle_lib.ptr(origin), oracle_lib.ptr(direction), travel)) == expected


# ---------------------------------------------------------------------------------------------------------------------
# Whole queries: PreparedScene.Trace / Occlude -> QuadBoundingVolumeHierarchy -> BoxBound4.Intersect -> GeometryCollection
#
# Scene: triangle 0 = A (z = 0), triangle 1 = A moved to z = -1, triangle 2 = B with vertices (4,0,0) (5,0,0) (5,1,0)
# [vertex0 (4,0,0), edge1 (1,0,0), edge2 (1,1,0)], sphere 0 at (10,0,0) radius 1.
# Tokens: TokenType.Triangle = 1, Sphere = 2 in the top 4 bits (EntityToken.cs:22-39), so triangle i = 0x10000000 + i.
#
# BoxBound4.Intersect (BoxBound4.cs:64-112), per axis: length0 = (min - o) dR, length1 = (max - o) dR,
#   far = min(far, maxps(length0, length1)), near = max(near, minps(length0, length1)) — Float4.Max(a, b) is Sse.Max(a, b)
#   (Float4.cs:376-377): maxps / minps return their SECOND operand when either is NaN. far *= 1.00000024; hit iff far >= near & far >= 0.
# A ray with a zero direction component has dR = +inf on that axis (Ray.cs:23: 1 / 0). If its origin sits exactly ON the box's
# MIN plane of that axis, length0 = 0 * inf = NaN and length1 = +inf: maxps(NaN, +inf) = +inf but minps(NaN, +inf) = +inf too, so
# near becomes +inf and the box is MISSED. On the MAX plane, length0 = -inf and length1 = NaN: the first axis gives far = near = NaN,
# and the next axis repairs both (minps(NaN, x) = x, maxps(NaN, x) = x): the box is HIT. Grazing rays are asymmetric in the reference.
# ---------------------------------------------------------------------------------------------------------------------
T0, T1, T2, S0 = 0x10000000, 0x10000001, 0x10000002, 0x20000000
EMPTY = 0xFFFFFFFF


def kat_scene():
    triangles = np.concatenate([
        scenes.make_triangles([(0, 0, 0)], [(1, 0, 0)], [(0, 1, 0)], 0),
        scenes.make_triangles([(0, 0, -1)], [(1, 0, -1)], [(0, 1, -1)], 0),
        scenes.make_triangles([(4, 0, 0)], [(5, 0, 0)], [(5, 1, 0)], 0),
    ])
    spheres = np.zeros(1, dtype=structs.SPHERE)
    spheres["position"], spheres["radius"], spheres["material"] = (10, 0, 0), 1.0, 0
    description = host.SceneDescription(triangles=triangles, spheres=spheres, materials=scenes.material(structs.MATERIAL_DIFFUSE, (0.5, 0.5, 0.5)),
                                        camera=scenes.perspective_camera((0, 0, -5)))
    return host.prepare(description)


# (name, origin, direction, distance in, ignore, expected token, expected distance, expected uv or None)
TRACE_CASES = [
    ("closest of two stacked triangles", (0.25, 0.25, 1), (0, 0, -1), INF, EMPTY, T0, 1.0, (0.25, 0.25)),
    # TraceQuery.ignore = the triangle the ray leaves (GeometryCollection.cs:93-94): the one behind it is found instead, at z = -1
    ("ignore skips the first triangle", (0.25, 0.25, 1), (0, 0, -1), INF, T0, T1, 2.0, (0.25, 0.25)),
    # TraceQuery.distance caps the search and acceptance is strict, `distance < query.distance` (GeometryCollection.cs:99): a cap
    # exactly at the surface finds nothing, token stays empty and distance stays the input (PreparedScene.cs:66-75)
    ("cap exactly on the surface", (0.25, 0.25, 1), (0, 0, -1), 1.0, EMPTY, EMPTY, 1.0, None),
    ("cap just beyond the first surface", (0.25, 0.25, 1), (0, 0, -1), 1.5, EMPTY, T0, 1.0, (0.25, 0.25)),
    ("between the two, looking away from both", (0.25, 0.25, 1), (0, 0, 1), INF, EMPTY, EMPTY, INF, None),
    ("from between, downwards", (0.25, 0.25, -0.5), (0, 0, -1), INF, EMPTY, T1, 0.5, (0.25, 0.25)),
    # grazing the MIN plane x = 0 of A's box with direction.x = 0: the triangle itself would accept u = 0 (above), its box does not
    ("grazing a box's min plane misses", (0, 0.5, 1), (0, 0, -1), INF, EMPTY, EMPTY, INF, None),
    # grazing the MAX plane x = 5 of B's box: the box is hit, and so is B:
    #   cross2 = (0,0,-1) x (1,1,0) = (0*0 - (-1)*1, (-1)*1 - 0*0, 0) = (1,-1,0); det = (1,0,0).(1,-1,0) = 1
    #   offset = (5,1/2,1) - (4,0,0) = (1,1/2,1); u = 1 - 1/2 = 1/2; cross1 = (1/2*0 - 1*0, 1*1 - 1*0, 1*0 - 1/2*1) = (0,1,-1/2)
    #   v = (0,0,-1).(0,1,-1/2) = 1/2 (u + v = 1: accepted); distance = (1,1,0).(0,1,-1/2) = 1
    ("grazing a box's max plane hits", (5, 0.5, 1), (0, 0, -1), INF, EMPTY, T2, 1.0, (0.5, 0.5)),
    # PreparedSphere.Intersect (SphereEntity.cs:88-124): offset = o - c = (0,0,3); center = -offset . d = 3;
    #   extend2 = fma(3, 3, 1 - 9) = 1; distance = 3 - 1 = 2
    ("sphere from outside", (10, 0, 3), (0, 0, -1), INF, EMPTY, S0, 2.0, None),
    # from the centre: offset = 0, center = 0, extend2 = fma(0, 0, 1 - 0) = 1; distance = 0 - 1 < 0 -> center + extend = 1
    ("sphere from its centre", (10, 0, 0), (0, 0, -1), INF, EMPTY, S0, 1.0, None),
    # PreparedScene.Trace's guard (PreparedScene.cs:69): distance must be positive in FastMath's sense (>= 8e-7)
    ("non-positive distance is not traced", (0.25, 0.25, 1), (0, 0, -1), 0.0, EMPTY, EMPTY, 0.0, None),
]

# (name, origin, direction, travel, ignore, expected)
OCCLUDE_CASES = [
    ("occluder inside travel", (0.25, 0.25, 1), (0, 0, -1), 1.5, EMPTY, True),
    ("travel ends exactly on the occluder", (0.25, 0.25, 1), (0, 0, -1), 1.0, EMPTY, False),  # strict, TriangleEntity.cs:262
    ("ignored occluder, travel short of the next", (0.25, 0.25, 1), (0, 0, -1), 1.5, T0, False),
    ("ignored occluder, travel reaches the next", (0.25, 0.25, 1), (0, 0, -1), 2.5, T0, True),
    ("grazing a box's min plane is not occluded", (0, 0.5, 1), (0, 0, -1), 10.0, EMPTY, False),
    ("grazing a box's max plane is occluded", (5, 0.5, 1), (0, 0, -1), 10.0, EMPTY, True),
    # PreparedSphere.Intersect(ray, travel) (SphereEntity.cs:126-148): the near root 2 is inside travel 2.5
    ("sphere inside travel", (10, 0, 3), (0, 0, -1), 2.5, EMPTY, True),
    ("sphere beyond travel", (10, 0, 3), (0, 0, -1), 1.5, EMPTY, False),
]


def trace_rays(cases):
    rays = np.zeros(len(cases), dtype=structs.RAY)
    for index, case in enumerate(cases):
        rays["origin"][index], rays["direction"][index], rays["distance"][index], rays["ignore"][index] = case[1], case[2], case[3], case[4]
    return rays


def check_trace(hits):
    for case, hit in zip(TRACE_CASES, hits):
        name, _, _, _, _, token, distance, uv = case
        assert int(hit["token"]) == token, name
        assert float(hit["distance"]) == distance, name
        if uv is not None:
            assert tuple(float(v) for v in hit["uv"]) == uv, name


def check_occlude(flags):
    for case, flag in zip(OCCLUDE_CASES, flags):
        assert bool(flag) == case[5], case[0]


def test_whole_queries_by_hand():
    oracle = oracle_lib.OracleScene(kat_scene())
    check_trace(oracle.trace(trace_rays(TRACE_CASES)))
    check_occlude(oracle.occlude(trace_rays(OCCLUDE_CASES)))


# ---------------------------------------------------------------------------------------------------------------------
# Accumulator.Add / Value / Noise — Accumulator.cs:11-71 over Summation.cs:8-58
#   ++count; delta = average - sample; average -= delta / count; squared += delta * (average - sample)
# Samples 1, 2, 3, 4 in every lane (all exactly representable, Kahan errors stay zero):
#   n=1: delta = -1;   average = 0 + 1 = 1;        squared = 0 + (-1)(1 - 1) = 0
#   n=2: delta = -1;   average = 1 + 1/2 = 3/2;    squared = 0 + (-1)(3/2 - 2) = 1/2
#   n=3: delta = -3/2; delta * (1f / 3f) = -1.5 * 0.333333343 = -0.50000001 -> rounds to -1/2; average = 2; squared = 1/2 + (-3/2)(2 - 3) = 2
#   n=4: delta = -2;   average = 2 + 2/4 = 5/2;    squared = 2 + (-2)(5/2 - 4) = 5
# Value = 5/2. Noise (count >= 2): oneLess = 3, cubed 27; numerator = Value^2 * 27 = 168.75; 1 / sqrt(168.75 * (1 / 5)) = 1 / sqrt(33.75)
# = 0.17213259 (the library evaluates the reciprocal and the square root exactly, INTEGRATION.md "Where results can differ").
# ---------------------------------------------------------------------------------------------------------------------
def test_accumulator_by_hand():
    samples = np.repeat(np.array([1, 2, 3, 4], dtype=np.float32)[:, None], 4, axis=1)
    value, noise, count = oracle_lib.accumulate(samples)
    assert count == 4
    assert np.all(value == F32(2.5))
    assert noise == pytest.approx(1.0 / np.sqrt(33.75), rel=2e-6)

    # a non-finite sample is rejected whole (Accumulator.cs:62) and does not advance the count
    samples = np.concatenate([samples, np.array([[INF, 0, 0, 0]], dtype=np.float32)])
    value, _, count = oracle_lib.accumulate(samples)
    assert count == 4 and np.all(value == F32(2.5))


def numpy_accumulator(samples):
    """Summation.cs + Accumulator.Add transcribed into numpy float32 arithmetic (every operation rounds to binary32 as the C# does)."""
    zero = np.zeros(4, dtype=np.float32)

    def add(total, error, value):  # Summation + Float4, Summation.cs:35-42
        delta = value - error
        new_total = total + delta
        return new_total, (new_total - total) - delta

    def add_summation(total, error, other_total, other_error):  # Summation + Summation, :46-55
        combined = error + other_error
        delta = other_total - combined
        new_total = total + delta
        return new_total, (new_total - total) - delta

    average, average_error, squared, squared_error, count = zero.copy(), zero.copy(), zero.copy(), zero.copy(), 0

    for sample in np.asarray(samples, dtype=np.float32):
        if not np.isfinite(np.float32(np.float32(sample[0] + sample[1]) + np.float32(sample[2] + sample[3]))):  # Float4.Sum pairs lanes (Float4.cs:73-81)
            continue
        count += 1
        delta_total, delta_error = add(average, average_error, -sample)               # average - sample
        reciprocal = np.float32(1.0) / np.full(4, count, dtype=np.float32)            # Summation / Float4 = * (1f / value)
        scaled_total, scaled_error = delta_total * reciprocal, delta_error * reciprocal
        average, average_error = add_summation(average, average_error, -scaled_total, -scaled_error)  # average -= ...
        difference, _ = add(average, average_error, -sample)                           # (average - sample).Result
        product_total, product_error = delta_total * difference, delta_error * difference  # Summation * Float4
        squared, squared_error = add_summation(squared, squared_error, product_total, product_error)

    return average, count


def test_accumulator_matches_an_independent_transcription():
    random = np.random.default_rng(5)
    for scale in (1.0, 1e-3, 1e4):
        samples = (random.random((257, 4)).astype(np.float32) * np.float32(scale)).astype(np.float32)
        samples[100, 2] = np.nan  # rejected
        value, _, count = oracle_lib.accumulate(samples)
        expected, expected_count = numpy_accumulator(samples)
        assert count == expected_count == 256
        assert np.array_equal(value.view(np.uint32), expected.view(np.uint32))


# ---------------------------------------------------------------------------------------------------------------------
# LightBound.Importance (LightBound.cs:30-80), LightTree.Pick / ProbabilityMass (LightTree.cs:53-57,115-154) transcribed into plain
# Python double arithmetic from the C# alone — no oracle header open. The oracle computes the same in binary32 in the reference's
# operation order, so it must agree to rounding (a few 1e-6 relative), on every emitter of a many-lights scene from many shading
# points: split probabilities, the descent, and pick-pdf == probability-mass.
# ---------------------------------------------------------------------------------------------------------------------
def _identity(value):  # FastMath.Identity = Sqrt0(1 - value^2), FastMath.cs:157-172
    return max(0.0, 1.0 - value * value) ** 0.5


def _clamp_subtract_cos(sin0, cos0, sin1, cos1):  # LightBound.cs:79
    return 1.0 if cos0 > cos1 else cos0 * cos1 + sin0 * sin1


def _clamp_subtract_sin(sin0, cos0, sin1, cos1):  # LightBound.cs:80
    return 0.0 if cos0 > cos1 else sin0 * cos1 - cos0 * sin1


def python_importance(node, position, normal):
    box_min, box_max = np.asarray(node["boxMin"], np.float64), np.asarray(node["boxMax"], np.float64)
    incident = np.asarray(position, np.float64) - (box_max + box_min) / 2.0  # BoxBound.Center
    length2 = float(incident @ incident)
    incident = np.zeros(3) if abs(length2) < 8e-7 else incident / length2 ** 0.5  # FastMath.AlmostZero

    cos_axis = float(np.asarray(node["coneAxis"], np.float64) @ incident)
    sin_axis = _identity(cos_axis)
    cos_offset = float(node["cosOffset"])
    sin_offset = _identity(cos_offset)

    extent = box_max - box_min  # FindSubtendedAngles, :63-77
    radius2 = float(extent @ extent) / 4.0
    if length2 < radius2:
        sin_radius, cos_radius = 0.0, -1.0
    else:
        sin_radius = (radius2 / length2) ** 0.5
        cos_radius = max(0.0, 1.0 - radius2 / length2) ** 0.5

    cos_remain = _clamp_subtract_cos(sin_axis, cos_axis, sin_offset, cos_offset)
    sin_remain = _clamp_subtract_sin(sin_axis, cos_axis, sin_offset, cos_offset)
    cos_final = _clamp_subtract_cos(sin_remain, cos_remain, sin_radius, cos_radius)
    if cos_final <= float(node["cosExtend"]):
        return 0.0

    cos_incident = abs(float(np.asarray(normal, np.float64) @ incident))
    cos_reflect = _clamp_subtract_cos(_identity(cos_incident), cos_incident, sin_radius, cos_radius)
    length2 = max(length2, float(extent @ extent) ** 0.5 / 2.0)
    return max(0.0, float(node["power"]) / length2 * cos_final * cos_reflect)


def python_pick(nodes, position, normal, sample):
    """LightTree.Pick: returns (token, pdf); pdf 0 = Probable.Impossible."""
    index, pdf = 0, 1.0
    while True:
        node = nodes[index]
        if node["child0"] == EMPTY:
            return int(node["child1"]), pdf
        first, second = python_importance(nodes[node["child0"]], position, normal), python_importance(nodes[node["child1"]], position, normal)
        if first < 8e-7 and second < 8e-7:  # !Positive && !Positive, LightTree.cs:122
            return EMPTY, 0.0
        split = first / (first + second)
        if sample < split:
            sample, index, pdf = sample / split, int(node["child0"]), pdf * split            # Sample1D.Stretch(0, split)
        else:
            sample, index, pdf = (sample - split) / (1.0 - split), int(node["child1"]), pdf * (1.0 - split)


def python_mass(nodes, path, position, normal):
    """LightTree.ProbabilityMass: the product of the split factors along the emitter's bit path (0 = first child)."""
    index, mass = 0, 1.0
    while nodes[index]["child0"] != EMPTY:
        node = nodes[index]
        first, second = python_importance(nodes[node["child0"]], position, normal), python_importance(nodes[node["child1"]], position, normal)
        split = first / (first + second) if first + second > 0 else float("nan")
        mass *= split if (path & 1) == 0 else 1.0 - split
        index = int(node["child0"] if (path & 1) == 0 else node["child1"])
        path >>= 1
    return mass


def test_light_tree_matches_an_independent_transcription():
    prepared = host.prepare(scenes.many_lights_scene(light_count=200, rings=8, segments=8))
    oracle = oracle_lib.OracleScene(prepared)
    nodes = prepared.light_nodes
    paths = dict(zip((int(t) for t in prepared.emitter_tokens), (int(p) for p in prepared.emitter_bitpaths)))
    random = np.random.default_rng(11)
    checked = impossible = 0

    for _ in range(300):
        position = random.uniform(-25, 25, 3).astype(np.float32)
        position[1] = abs(position[1]) * 0.4
        normal = random.normal(size=3)
        normal = (normal / np.linalg.norm(normal)).astype(np.float32)
        sample = float(np.float32(random.random()))

        token, pdf = oracle.light_pick(position, normal, sample)
        expected_token, expected_pdf = python_pick(nodes, position, normal, sample)

        if expected_pdf == 0.0 or pdf == 0.0:
            assert expected_pdf < 1e-6 and pdf < 1e-6
            impossible += 1
            continue
        if token != expected_token:  # the sample sat within rounding of a split: the neighbouring branch is as right as this one
            continue
        assert pdf == pytest.approx(expected_pdf, rel=2e-4)
        assert oracle.light_mass(token, position, normal) == pytest.approx(python_mass(nodes, paths[token], position, normal), rel=2e-4)
        assert oracle.light_mass(token, position, normal) == pytest.approx(pdf, rel=2e-4)  # Pick's pdf IS the probability mass
        checked += 1

    assert checked > 200, (checked, impossible)


# ---------------------------------------------------------------------------------------------------------------------
# LightTree.Build (LightTree.cs:62-113) over LightBound.Encapsulate / RelativeArea (LightBound.cs:22-28), ConeBound.Union /
# Encapsulate / RelativeArea (ConeBound.cs:28-101), BoxBound.HalfArea / MajorAxis / Center / Encapsulate (BoxBound.cs:80-132),
# Float3.Angle (Float3.cs:277-288, DEGREES), Versor(axis, angle) * v (Versor.cs, angle in DEGREES) and LightCollection.CreateBounds
# (LightCollection.cs:91-137) with PreparedTriangle.BoxBound / ConeBound / Area (TriangleEntity.cs:142-148) and Emissive.Power
# (Emissive.cs:53), transcribed into plain Python double arithmetic from the C# alone. The host mirror's recursive build and the
# device build share ONE binary32 restatement (csrc/echo_light_build.h): a misreading there would pass every byte comparison between
# them. This transcription shares nothing with it; the two must produce the same tree — the same cuts, hence the same emitter order and
# bit paths — and the same bounds to rounding. Ties are avoided by the scene (random positions), not by the code.
# ---------------------------------------------------------------------------------------------------------------------
import math  # noqa: E402


def _py_cone_union(value0, value1):  # ConeBound.Union, ConeBound.cs:76-101; a cone = (axis, cosOffset, cosExtend)
    clamp = lambda v: min(1.0, max(-1.0, v))
    offset0, offset1 = math.acos(clamp(value0[1])), math.acos(clamp(value1[1]))
    cos_extend = min(value0[2], value1[2])
    axis = value0[0]

    squared = float(np.dot(value0[0], value0[0]) * np.dot(value1[0], value1[0]))  # Float3.Angle: degrees
    angle = 0.0 if squared == 0.0 else math.degrees(math.acos(clamp(float(np.dot(value0[0], value1[0])) / math.sqrt(squared))))
    maximum = angle + offset1  # degrees + radians, as written in the reference

    if min(maximum, math.pi) <= offset0:
        return axis, value0[1], cos_extend
    offset = (offset0 + maximum) / 2.0
    if offset >= math.pi:
        return np.array([0.0, 1.0, 0.0]), -1.0, cos_extend  # CreateFullSphere

    cross = np.cross(axis, value1[0])
    cross = cross / np.linalg.norm(cross)
    half = math.radians(offset - offset0) / 2.0  # new Versor(cross, rotation): rotation is taken as DEGREES
    q = np.concatenate([cross * math.sin(half), [math.cos(half)]])  # a unit quaternion; rotate v by it: v + 2 w (u x v) + 2 u x (u x v)
    u, w = q[:3], q[3]
    rotated = axis + 2.0 * w * np.cross(u, axis) + 2.0 * np.cross(u, np.cross(u, axis))
    return rotated, math.cos(offset), cos_extend


def _py_encapsulate(a, b):  # LightBound.Encapsulate; a bound = (lo, hi, cone, power)
    cone = _py_cone_union(a[2], b[2]) if b[2][1] > a[2][1] else _py_cone_union(b[2], a[2])  # ConeBound.Encapsulate, :50-56
    return np.minimum(a[0], b[0]), np.maximum(a[1], b[1]), cone, a[3] + b[3]


def _py_relative_area(bound):  # LightBound.RelativeArea = box.HalfArea * cone.RelativeArea * power
    size = bound[1] - bound[0]
    half_area = size[0] * (size[1] + size[2]) + size[1] * size[2]
    cos_offset, cos_extend = bound[2][1], bound[2][2]
    offset, extend = math.acos(min(1.0, max(-1.0, cos_offset))), math.acos(min(1.0, max(-1.0, cos_extend)))
    angle = min(offset + extend, math.pi) * 2.0
    sin_offset = math.sqrt(max(0.0, 1.0 - cos_offset * cos_offset))
    cone_area = 2.0 * math.pi * (1.0 - cos_offset) + math.pi / 2.0 * (angle * sin_offset - math.cos(offset - angle) - 2.0 * offset * sin_offset + cos_offset)
    return half_area * cone_area * bound[3]


def _py_build(items, emitted, depth=0, path=0):
    """items: [(token, bound)]. Appends (token or None, bound, child0, child1) rows to `emitted` in the reference's construction order
    (`new Node(Build(tail), Build(head))`: the parent, then the tail subtree, then the head subtree) and yields the leaves' bit paths."""
    if len(items) == 1:
        emitted.append([items[0][0], items[0][1], None, None, path])
        return len(emitted) - 1

    lo, hi = items[0][1][0], items[0][1][1]
    for _, bound in items:
        lo, hi = np.minimum(lo, bound[0]), np.maximum(hi, bound[1])
    size = hi - lo
    major = (0 if size[0] > size[2] else 2) if size[0] > size[1] else (1 if size[1] > size[2] else 2)  # Float3.MaxIndex, Float3.cs:130-138
    items = sorted(items, key=lambda item: (item[1][1][major] + item[1][0][major]) / 2.0)  # by box centre (stable, like the mirror)

    count = len(items)
    costs = [0.0] * count
    bound = items[-1][1]
    for i in range(count - 2, -1, -1):
        costs[i + 1] = _py_relative_area(bound)
        bound = _py_encapsulate(bound, items[i][1])

    best, cut = math.inf, -1
    bound = items[0][1]
    for i in range(1, count):
        cost = costs[i] + _py_relative_area(bound)
        if cost < best:
            best, cut = cost, i
        bound = _py_encapsulate(bound, items[i][1])

    emitted.append(None)
    index = len(emitted) - 1
    child0 = _py_build(items[cut:], emitted, depth + 1, path)
    child1 = _py_build(items[:cut], emitted, depth + 1, path | 1 << depth)
    emitted[index] = [None, _py_encapsulate(emitted[child0][1], emitted[child1][1]), child0, child1, None]
    return index


def _py_emitters(description):  # LightCollection.CreateBounds: emissive triangles only (no point lights, spheres or placements in this scene)
    items = []
    for i, t in enumerate(description.triangles):
        m = description.materials[t["material"]]
        if m["type"] != structs.MATERIAL_EMISSIVE:
            continue
        e1, e2, v0 = t["edge1"].astype(np.float64), t["edge2"].astype(np.float64), t["vertex0"].astype(np.float64)
        normal = np.cross(e1, e2)
        area = np.linalg.norm(normal) / 2.0
        r, g, b = (float(c) for c in m["albedo"][:3])
        power = (r * 0.212671 + g * 0.715160 + b * 0.072169) * math.pi * area  # Emissive.Power = emission.Luminance * pi (RGB128.cs), times Area
        if not power > 0.0:
            continue
        points = np.stack([v0, v0 + e1, v0 + e2])
        items.append(((structs.TOKEN_TYPE_TRIANGLE << 28) | i, (points.min(axis=0), points.max(axis=0), (normal / np.linalg.norm(normal), 1.0, 0.0), power)))
    return items


def _tilted_ceiling(description, degrees, seed=3):
    """every emissive triangle of the scene turned to face down, tilted by at most `degrees`: unions of such cones stay proper cones
    (ConeBound.Union's rotation path) instead of jumping to the whole sphere, which any two axes more than 2 pi DEGREES apart do"""
    rng = np.random.default_rng(seed)
    lit = np.flatnonzero(description.materials["type"][description.triangles["material"]] == structs.MATERIAL_EMISSIVE)
    side = np.linalg.norm(description.triangles["edge1"][lit], axis=1, keepdims=True).astype(np.float64)
    tilt, spin = np.radians(rng.uniform(0, degrees, len(lit))), rng.uniform(0, 2 * np.pi, len(lit))
    normal = np.stack([np.sin(tilt) * np.cos(spin), -np.cos(tilt), np.sin(tilt) * np.sin(spin)], axis=-1)
    e1 = np.cross(normal, [0.0, 0.0, 1.0])
    e1 /= np.linalg.norm(e1, axis=1, keepdims=True)
    e2 = np.cross(e1, normal)  # e1 x e2 = normal
    description.triangles["edge1"][lit] = (e1 * side).astype(np.float32)
    description.triangles["edge2"][lit] = (e2 * side).astype(np.float32)
    return description


@pytest.mark.parametrize("tilt", [None, 1.5])
def test_light_tree_build_matches_an_independent_transcription(tilt):
    description = scenes.many_lights_scene(light_count=400, rings=8, segments=8)
    if tilt is not None:
        description = _tilted_ceiling(description, tilt)
    nodes, tokens, paths, power = host.build_light_tree(description)
    emitted = []
    assert _py_build(_py_emitters(description), emitted) == 0
    assert len(emitted) == len(nodes) == 799

    leaves = [(row[0], row[4]) for row in emitted if row[2] is None]
    assert [int(t) for t in tokens] == [t for t, _ in leaves]  # the same cuts everywhere: the same pre-order of emitters ...
    assert [int(p) for p in paths] == [p for _, p in leaves]   # ... reached by the same branches

    proper = 0
    for row, node in zip(emitted, nodes):
        lo, hi, cone, node_power = row[1]
        if row[2] is None:
            assert node["child0"] == 0xFFFFFFFF and node["child1"] == row[0]
        else:
            assert (int(node["child0"]), int(node["child1"])) == (row[2], row[3])
            proper += -1.0 < cone[1] < 1.0
        assert np.allclose(node["boxMin"], lo, rtol=1e-6, atol=1e-6) and np.allclose(node["boxMax"], hi, rtol=1e-6, atol=1e-6)
        assert node["power"] == pytest.approx(node_power, rel=1e-4)
        assert node["cosOffset"] == pytest.approx(cone[1], abs=2e-5) and node["cosExtend"] == pytest.approx(cone[2], abs=1e-6)
        if cone[1] > -1.0:  # the whole sphere's axis is a convention (Float3.Up), a proper cone's is geometry
            assert np.allclose(node["coneAxis"], cone[0], atol=2e-4)
    assert power == pytest.approx(emitted[0][1][3], rel=1e-4)
    if tilt is not None:
        assert proper > 300  # branches whose cone is neither a direction nor the whole sphere: the rotation path of ConeBound.Union was taken


# ---------------------------------------------------------------------------------------------------------------------
# QuadBoundingVolumeHierarchy.TraceImpl (QuadBoundingVolumeHierarchy.cs:123-216) over BoxBound4.Intersect (BoxBound4.cs:64-112),
# GeometryCollection.Trace (GeometryCollection.cs:85-104) and PreparedTriangle.IntersectImpl (TriangleEntity.cs:204-235), transcribed
# into Python binary32 scalar arithmetic from the C# alone: a stack of (node, entry distance) pairs popped last-in first-out,
# entries at or beyond the closest hit dropped at the pop; per node one four-box slab test, then the four `Push` calls in the order
# the ray's direction signs select (the second pair first when orders[axisMajor]), a Push dropping children at or beyond the
# closest hit, stacking branches and testing leaves ON THE SPOT. Hits do not depend on that order (up to ties) — the brute-force test
# of test_host_and_traversal.py pins them — but the number of node visits and primitive tests does, and those counts are what
# bench.py's roofline multiplies by 128 and 36 bytes. This pins the oracle's visit order: per ray the same hit bits,
# over the batch the same number of slab tests and triangle tests.
# ---------------------------------------------------------------------------------------------------------------------
F = np.float32


def _py_slab4(node, origin, direction_r):  # BoxBound4.Intersect; Float4.Max / Min = maxps / minps: the SECOND operand unless the first wins
    def axis(lo, hi, o, r):
        length0, length1 = (lo - o) * r, (hi - o) * r
        return np.where(length0 > length1, length0, length1), np.where(length0 < length1, length0, length1)

    with np.errstate(invalid="ignore", over="ignore"):
        far, near = axis(node["minX"], node["maxX"], origin[0], direction_r[0])
        for lo, hi, k in ((node["minY"], node["maxY"], 1), (node["minZ"], node["maxZ"], 2)):
            high, low = axis(lo, hi, origin[k], direction_r[k])
            far = np.where(far < high, far, high)
            near = np.where(near > low, near, low)
        far = far * F(1.00000024)  # BoxBound.FarMultiplier
        return np.where((far >= near) & (far >= F(0)), near, F(np.inf))


def _py_cross(a, b):  # Float3.Cross: products and difference in binary64, one rounding to binary32
    a, b = a.astype(np.float64), b.astype(np.float64)
    return np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]).astype(np.float32)


def _py_dot(a, b):  # Float3.Dot: (x x' + y y') + z z' in binary32
    return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]


def _py_triangle(triangle, origin, direction):  # PreparedTriangle.IntersectImpl; returns (distance, u, v)
    edge1, edge2, vertex0 = triangle["edge1"], triangle["edge2"], triangle["vertex0"]
    cross2 = _py_cross(direction, edge2)
    determinant = _py_dot(edge1, cross2)
    if determinant == 0:
        return F(np.inf), F(0), F(0)
    determinant_r = F(1) / determinant
    offset = origin - vertex0
    u = _py_dot(offset, cross2) * determinant_r
    if (u < 0) | (u > 1):
        return F(np.inf), u, F(0)
    cross1 = _py_cross(offset, edge1)
    v = _py_dot(direction, cross1) * determinant_r
    if (v < 0) | (u + v > 1):
        return F(np.inf), u, v
    distance = _py_dot(edge2, cross1) * determinant_r
    return (F(np.inf) if distance < 0 else distance), u, v


def _py_trace(nodes, triangles, ray, counts):
    origin, direction = ray["origin"], ray["direction"]
    with np.errstate(divide="ignore"):
        direction_r = F(1) / direction  # Ray.cs:23
    orders = [bool(direction_r[0] > 0), bool(direction_r[1] > 0), bool(direction_r[2] > 0), True]
    best, best_token, best_uv = ray["distance"], structs.TOKEN_EMPTY, (F(0), F(0))
    stack = [(0, F(0))]

    while stack:
        index, entry = stack.pop()
        if entry >= best:
            continue
        node = nodes[index]
        counts[0] += 1
        intersections = _py_slab4(node, origin, direction_r)

        first = (1, 0) if orders[node["axisMinor0"]] else (0, 1)
        second = (3, 2) if orders[node["axisMinor1"]] else (2, 3)
        for slot in (second + first if orders[node["axisMajor"]] else first + second):
            hit = intersections[slot]
            if hit >= best:
                continue
            token = int(node["token4"][slot])
            if structs.token_type(token) == structs.TOKEN_TYPE_NODE:
                stack.append((structs.token_index(token), hit))
                continue
            if token == int(ray["ignore"]):  # GeometryCollection.Trace: the ignored triangle is skipped before its test
                continue
            counts[1] += 1
            distance, u, v = _py_triangle(triangles[structs.token_index(token)], origin, direction)
            if distance >= best:
                continue
            best, best_token, best_uv = distance, token, (u, v)

    return best_token, best, best_uv


def test_trace_impl_matches_an_independent_transcription():
    prepared = host.prepare(scenes.terrain_scene(48, 24, 0))  # 2 304 triangles, no spheres: 1 109 quad nodes, depth 7
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 1500, seed=31)
    hits = oracle.trace(rays)
    rays["ignore"][::3] = hits["token"][::3]  # a third of the rays ignore what they would hit: the ignore rule takes part too
    rays["distance"][1::5] = np.where(np.isfinite(hits["distance"][1::5]), hits["distance"][1::5] * F(1.5), F(30.0))  # and a fifth start with a finite limit
    hits, counters = oracle.trace(rays, count_visits=True)

    counts = [0, 0]
    for ray, hit in zip(rays, hits):
        token, distance, uv = _py_trace(prepared.nodes, prepared.triangles, ray, counts)
        found = token != structs.TOKEN_EMPTY
        assert int(hit["token"]) == token
        if found:
            assert hit["distance"].view(np.uint32) == np.float32(distance).view(np.uint32)
            assert hit["uv"][0].view(np.uint32) == np.float32(uv[0]).view(np.uint32) and hit["uv"][1].view(np.uint32) == np.float32(uv[1]).view(np.uint32)

    assert np.count_nonzero(hits["token"] != structs.TOKEN_EMPTY) > 150
    assert [int(counters[0]), int(counters[1]), int(counters[2])] == [counts[0], counts[1], 0]  # the same slab tests, the same triangle tests


# OccludeImpl (QuadBoundingVolumeHierarchy.cs:223-315): the same walk with no entry distances on the stack (nothing is dropped at the pop),
# children dropped at or beyond `travel`, the first occluding primitive ends the query; GeometryCollection.Occlude (GeometryCollection.cs:134-172)
# and PreparedTriangle.IntersectImpl(origin, direction, travel) (TriangleEntity.cs:237-263: sign(det) instead of 1 / det, STRICT in travel).
def _py_triangle_occludes(triangle, origin, direction, travel):
    edge1, edge2, vertex0 = triangle["edge1"], triangle["edge2"], triangle["vertex0"]
    cross2 = _py_cross(direction, edge2)
    determinant = _py_dot(edge1, cross2)
    if determinant == 0:
        return False
    sign = F(1) if determinant > 0 else F(-1)
    determinant = determinant * sign
    offset = origin - vertex0
    u = _py_dot(offset, cross2) * sign
    if (u < 0) | (u > determinant):
        return False
    cross1 = _py_cross(offset, edge1)
    v = _py_dot(direction, cross1) * sign
    if (v < 0) | (u + v > determinant):
        return False
    distance = _py_dot(edge2, cross1) * sign
    return bool((distance >= 0) & (distance < travel * determinant))


def _py_occlude(nodes, triangles, ray, counts):
    origin, direction, travel = ray["origin"], ray["direction"], ray["distance"]
    with np.errstate(divide="ignore"):
        direction_r = F(1) / direction
    orders = [bool(direction_r[0] > 0), bool(direction_r[1] > 0), bool(direction_r[2] > 0), True]
    stack = [0]

    while stack:
        node = nodes[stack.pop()]
        counts[0] += 1
        intersections = _py_slab4(node, origin, direction_r)
        first = (1, 0) if orders[node["axisMinor0"]] else (0, 1)
        second = (3, 2) if orders[node["axisMinor1"]] else (2, 3)
        for slot in (second + first if orders[node["axisMajor"]] else first + second):
            if intersections[slot] >= travel:
                continue
            token = int(node["token4"][slot])
            if structs.token_type(token) == structs.TOKEN_TYPE_NODE:
                stack.append(structs.token_index(token))
                continue
            if token == int(ray["ignore"]):
                continue
            counts[1] += 1
            if _py_triangle_occludes(triangles[structs.token_index(token)], origin, direction, travel):
                return True
    return False


def test_occlude_impl_matches_an_independent_transcription():
    prepared = host.prepare(scenes.terrain_scene(48, 24, 0))
    oracle = oracle_lib.OracleScene(prepared)
    rays = scenes.random_rays(prepared.bounds, 1500, seed=37, occlusion=True)
    rays["ignore"][::3] = oracle.trace(rays)["token"][::3]
    occluded, counters = oracle.occlude(rays, count_visits=True)

    counts = [0, 0]
    for ray, expected in zip(rays, occluded):
        assert _py_occlude(prepared.nodes, prepared.triangles, ray, counts) == bool(expected)
    assert 100 < np.count_nonzero(occluded) < 1400
    assert [int(counters[0]), int(counters[1]), int(counters[2])] == [counts[0], counts[1], 0]


# ---------------------------------------------------------------------------------------------------------------------
# SweepBuilder.Build / BuildLayer / BuildChild / Sorter (SweepBuilder.cs:24-251) and the binary -> quad collapse of
# QuadBoundingVolumeHierarchy (constructor :24-36, CreateNode :363-404, Node :409-430, BuildNode :470-565), transcribed into Python from
# the C# alone: box arithmetic in binary32, the sort a stable sort on Sorter.Transform's key, the children of a quad node a linked list
# built back to front. The repository already holds two restatements that agree byte for byte — the host mirror's recursion
# (csrc/host/echo_host.cpp) and the level-synchronous device build (csrc/echo_sweep.h) — but both come from one reading of the source.
# This third one shares nothing with them and must emit the same 128-byte nodes: boxes, axes, tokens, pre-order indices, depth.
# ---------------------------------------------------------------------------------------------------------------------
class _Bin:  # HierarchyBuilder.Node
    def __init__(self, lo, hi, child0=None, child1=None, axis=0, token=None):
        self.lo, self.hi, self.child0, self.child1, self.axis, self.token = lo, hi, child0, child1, axis, token

    @property
    def leaf(self):
        return self.child0 is None


def _half_area(lo, hi):  # BoxBound.HalfArea
    size = hi - lo
    return size[0] * (size[1] + size[2]) + size[1] * size[2]


def _max_index(v):  # Float3.MaxIndex
    if v[0] > v[1]:
        return 0 if v[0] > v[2] else 2
    return 1 if v[1] > v[2] else 2


def _transform(value):  # Sorter.Transform: flip the sign bit of a positive value, all bits of a negative one
    bits = int(np.float32(value).view(np.uint32))
    return bits ^ (0xFFFFFFFF if bits >> 31 else 0x80000000)


def _sort_by_axis(data, axis):  # Sorter.Sort: insertion sort / LSD radix sort on the keys, both stable
    return sorted(data, key=lambda item: _transform(item[1][axis] + item[2][axis]))


def _build_layer(data):
    length = len(data)
    tails = [None] * length
    lo, hi = data[-1][1], data[-1][2]
    for i in range(length - 2, -1, -1):  # PrepareCutTailVolumes
        tails[i + 1] = (lo, hi)
        lo, hi = np.minimum(lo, data[i][1]), np.maximum(hi, data[i][2])

    head_lo, head_hi = data[0][1], data[0][2]
    min_cost, min_index, head, tail = F(np.finfo(np.float32).max), -1, None, None
    for i in range(1, length):  # SearchSurfaceAreaHeuristics
        cost = _half_area(head_lo, head_hi) * F(i) + _half_area(*tails[i]) * F(length - i)
        if cost < min_cost:
            min_cost, min_index, head, tail = cost, i, (head_lo, head_hi), tails[i]
        head_lo, head_hi = np.minimum(head_lo, data[i][1]), np.maximum(head_hi, data[i][2])

    lo, hi = np.minimum(head[0], tail[0]), np.maximum(head[1], tail[1])
    axis = _max_index(hi - lo)
    if min_index > length // 2:  # headData is always the larger half
        head_data, tail_data = data[:min_index], data[min_index:]
    else:
        head_data, tail_data = data[min_index:], data[:min_index]
        head, tail = tail, head

    child0, child1 = _build_child(head_data, head, axis), _build_child(tail_data, tail, axis)
    if _half_area(*head) < _half_area(*tail):  # the child with the larger surface area first
        child0, child1 = child1, child0
    return _Bin(lo, hi, child0, child1, axis)


def _build_child(data, parent, parent_axis):
    if len(data) == 1:
        return _Bin(data[0][1], data[0][2], token=data[0][0])
    axis = _max_index(parent[1] - parent[0])
    if axis != parent_axis:
        data = _sort_by_axis(data, axis)
    return _build_layer(data)


def _sweep_build(items):  # SweepBuilder.Build
    if len(items) == 1:
        return _Bin(items[0][1], items[0][2], token=items[0][0])
    lo, hi = items[0][1], items[0][2]
    for _, item_lo, item_hi in items:
        lo, hi = np.minimum(lo, item_lo), np.maximum(hi, item_hi)
    return _build_layer(_sort_by_axis(items, _max_index(hi - lo)))


def _children_sorted(node):  # BuildNode.GetChildrenSorted: the child with the smaller min along the node's axis first
    child0, child1 = node.child0, node.child1
    if child0.lo[node.axis] > child1.lo[node.axis]:
        child0, child1 = child1, child0
    return node.axis, child0, child1


def _quad_children(node):  # BuildNode's constructor: four children (None = empty) and the three axes
    def pair(child):  # AddChildren
        if child.leaf:
            return 3, [child, None]
        axis, first, second = _children_sorted(child)
        return axis, [first, second]

    axis_major, child0, child1 = _children_sorted(node)
    axis_minor0, first = pair(child0)
    axis_minor1, second = pair(child1)
    return (axis_major, axis_minor0, axis_minor1), first + second


def _quad_nodes(root):
    """CreateNode: a child that is a branch claims the next index BEFORE its own subtree is emitted (pre-order); depth as CreateNode counts it"""
    nodes = [None]

    def create(node, index):
        axes, children = _quad_children(node)
        record = np.zeros(1, dtype=structs.QBVH_NODE)[0]
        record["axisMajor"], record["axisMinor0"], record["axisMinor1"] = axes
        depth = 0
        for i, child in enumerate(children):
            lo, hi = (np.full(3, np.inf, dtype=np.float32),) * 2 if child is None else (child.lo, child.hi)  # BoxBound.None
            record["minX"][i], record["minY"][i], record["minZ"][i] = lo
            record["maxX"][i], record["maxY"][i], record["maxZ"][i] = hi
            if child is None:
                record["token4"][i], child_depth = structs.TOKEN_EMPTY, 0
            elif child.leaf:
                record["token4"][i], child_depth = child.token, 1
            else:
                child_index = len(nodes)
                nodes.append(None)
                child_depth = create(child, child_index)
                record["token4"][i] = child_index  # NewNodeToken: TokenType.Node = 0
            depth = max(depth, child_depth)
        nodes[index] = record
        return depth + 1

    depth = create(root, 0)
    return np.array(nodes, dtype=structs.QBVH_NODE), depth


@pytest.mark.parametrize("scene", ["terrain", "soup", "cornell"])
def test_sweep_builder_and_quad_collapse_match_an_independent_transcription(scene):
    if scene == "terrain":
        description = scenes.terrain_scene(24, 12, 60)  # 576 triangles + 60 spheres
        triangles, spheres = description.triangles, description.spheres
    elif scene == "cornell":
        description = scenes.cornell_box()
        triangles, spheres = description.triangles, description.spheres
    else:
        from tests.test_sweep_build import random_soup
        triangles, spheres = random_soup(14, 700, 90, 25.0)

    items = []  # GeometryCollection.CreateBounds: triangles, then spheres
    for i, t in enumerate(triangles):
        points = np.stack([t["vertex0"], t["vertex0"] + t["edge1"], t["vertex0"] + t["edge2"]])  # PreparedTriangle.BoxBound over vertex0, Vertex1, Vertex2
        items.append(((structs.TOKEN_TYPE_TRIANGLE << 28) | i, points.min(axis=0), points.max(axis=0)))
    for i, s in enumerate(spheres):
        items.append(((structs.TOKEN_TYPE_SPHERE << 28) | i, s["position"] - s["radius"], s["position"] + s["radius"]))

    expected, expected_depth = host.build_qbvh(triangles, spheres)
    nodes, depth = _quad_nodes(_sweep_build(items))
    assert len(nodes) == len(expected) and depth == expected_depth
    for field in ("token4", "axisMajor", "axisMinor0", "axisMinor1", "minX", "minY", "minZ", "maxX", "maxY", "maxZ"):
        assert nodes[field].tobytes() == expected[field].tobytes(), field


# ---------------------------------------------------------------------------------------------------------------------
# PerspectiveCamera.SpawnRay without depth of field (PerspectiveCamera.cs:93-98) over RaySpawner.SpawnX (RaySpawner.cs:37-46): the ray leaves
# the camera's position along InverseTransform * (uv.x, uv.y, forwardLength), normalised, with uv = ((x + jx) / W - 1/2, (y + jy) / W - (H / W) / 2)
# for a jitter (jx, jy) in [0, 1)^2 — both coordinates scaled by the WIDTH, rows growing upward. Worked out from the C#; the jitter comes from
# the library's own sample sequence, so the statement checked is the one that holds for any jitter: every ray of pixel (x, y), taken back into
# camera space with the transposed rotation, crosses the plane z = forwardLength inside that pixel's cell, and the cells of a frame tile the
# window [-1/2, 1/2] x [-(H / W) / 2, (H / W) / 2] — a wide frame (W > H), a rotated camera, several samples per pixel.
# ---------------------------------------------------------------------------------------------------------------------
def test_perspective_camera_rays_cross_their_pixel_cell():
    width, height, field_of_view = 48, 20, 50.0
    position, rotation = (3.0, 6.0, -14.0), (12.0, -25.0, 0.0)
    description = scenes.cornell_box()
    description.camera = scenes.perspective_camera(position, rotation, field_of_view=field_of_view, lens_radius=0.0)
    oracle = oracle_lib.OracleScene(host.prepare(description))

    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    pixels = np.repeat(np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1), 3, axis=0).astype(np.int32)
    index = np.tile(np.arange(3, dtype=np.uint32), width * height)
    rays = oracle.spawn_rays(structs.render_params(width, height, 16, extend=3, seed=4), pixels, index)

    assert np.all(rays["origin"] == np.array(position, dtype=np.float32))
    assert np.allclose(np.linalg.norm(rays["direction"].astype(np.float64), axis=1), 1.0, atol=1e-6)

    to_camera = scenes.rotation_matrix(*rotation).T  # a rotation: its inverse is its transpose
    local = rays["direction"].astype(np.float64) @ to_camera.T
    forward_length = 0.5 / math.tan(math.radians(field_of_view) / 2.0)
    assert np.all(local[:, 2] > 0)
    u, v = local[:, 0] / local[:, 2] * forward_length, local[:, 1] / local[:, 2] * forward_length

    slack = 1e-6
    assert np.all(u >= pixels[:, 0] / width - 0.5 - slack) and np.all(u <= (pixels[:, 0] + 1) / width - 0.5 + slack)
    assert np.all(v >= pixels[:, 1] / width - height / width / 2 - slack) and np.all(v <= (pixels[:, 1] + 1) / width - height / width / 2 + slack)
    jitter_x, jitter_y = (u + 0.5) * width - pixels[:, 0], (v + height / width / 2) * width - pixels[:, 1]
    assert 0.35 < jitter_x.mean() < 0.65 and 0.35 < jitter_y.mean() < 0.65 and jitter_x.std() > 0.2 and jitter_y.std() > 0.2  # the whole cell is used


def test_thin_lens_rays_leave_the_lens_and_cross_their_pixel_cell_on_the_focal_plane():
    """PerspectiveCamera.SpawnRay with depth of field (PerspectiveCamera.cs:61-77): the origin is a point of the lens disk (radius LensRadius, in the
    camera's z = 0 plane), the ray aims at uv * focusScale on the plane z = FocalDistance, focusScale = FocalDistance / forwardLength — so whatever the
    lens sample, a pixel's rays meet the focal plane inside that pixel's cell, magnified by focusScale."""
    width, height, field_of_view, lens_radius, focal_distance = 40, 24, 55.0, 0.35, 9.0
    position, rotation = (1.0, 2.0, -11.0), (-8.0, 15.0, 0.0)
    description = scenes.cornell_box()
    description.camera = scenes.perspective_camera(position, rotation, field_of_view=field_of_view, lens_radius=lens_radius, focal_distance=focal_distance)
    oracle = oracle_lib.OracleScene(host.prepare(description))

    ys, xs = np.meshgrid(np.arange(height), np.arange(width), indexing="ij")
    pixels = np.repeat(np.stack([xs.reshape(-1), ys.reshape(-1)], axis=-1), 4, axis=0).astype(np.int32)
    index = np.tile(np.arange(4, dtype=np.uint32), width * height)
    rays = oracle.spawn_rays(structs.render_params(width, height, 16, extend=4, seed=5), pixels, index)

    to_camera = scenes.rotation_matrix(*rotation).T
    origin = (rays["origin"].astype(np.float64) - np.array(position)) @ to_camera.T
    direction = rays["direction"].astype(np.float64) @ to_camera.T
    assert np.allclose(origin[:, 2], 0.0, atol=1e-5) and np.all(np.hypot(origin[:, 0], origin[:, 1]) <= lens_radius * (1 + 1e-5))
    assert np.hypot(origin[:, 0], origin[:, 1]).max() > 0.9 * lens_radius and np.abs(origin[:, :2].mean(axis=0)).max() < 0.05 * lens_radius  # the whole disk

    focus = origin + direction * ((focal_distance - origin[:, 2]) / direction[:, 2])[:, None]
    forward_length = 0.5 / math.tan(math.radians(field_of_view) / 2.0)
    u, v = focus[:, 0] / focal_distance * forward_length, focus[:, 1] / focal_distance * forward_length
    slack = 2e-6
    assert np.all(u >= pixels[:, 0] / width - 0.5 - slack) and np.all(u <= (pixels[:, 0] + 1) / width - 0.5 + slack)
    assert np.all(v >= pixels[:, 1] / width - height / width / 2 - slack) and np.all(v <= (pixels[:, 1] + 1) / width - height / width / 2 + slack)
