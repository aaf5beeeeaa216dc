"""Device BxDFs / FastMath shims against the oracle, bit for bit, plus the reference's BxDF property table replayed on the GPU."""
import numpy as np
import pytest

from echorenderer_b200 import _native
from tests import oracle_lib as ol
from tests.test_oracle_kats import BXDF_TABLE, FLOAT_VALUES, bxdf_inputs, run_bxdf_checks

pytestmark = pytest.mark.gpu


def device_bxdf_batch(kind, params, outgoing, samples):
    lib = _native.library()
    outgoing, samples = ol.f32(outgoing).reshape(-1, 3), ol.f32(samples).reshape(-1, 2)
    n = len(outgoing)
    sampled, evaluated, inverse = np.zeros((n, 8), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
    _native.check(lib.echo_b200_debug_bxdf_batch(0, kind, _native.pointer(ol.f32(params)), _native.pointer(outgoing), _native.pointer(samples), n,
                                                 _native.pointer(sampled), _native.pointer(evaluated), _native.pointer(inverse)))
    return sampled, evaluated, inverse


def device_math(op, a, b=None, c=None):
    lib = _native.library()
    a = ol.f32(a)
    b = ol.f32(b) if b is not None else np.zeros_like(a)
    c = ol.f32(c) if c is not None else np.zeros_like(a)
    out = np.zeros_like(a)
    _native.check(lib.echo_b200_debug_math(0, op, _native.pointer(a), _native.pointer(b), _native.pointer(c), len(a), _native.pointer(out)))
    return out


def bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)


@pytest.mark.parametrize("check_type,kind,params", BXDF_TABLE)
def test_device_bxdf_bit_identical_to_oracle(check_type, kind, params):
    outgoings, samples = bxdf_inputs()
    outgoing = np.repeat(outgoings, 64, axis=0)
    sample = np.tile(samples[:64], (len(outgoings), 1))
    got = device_bxdf_batch(kind, params, outgoing, sample)
    want = ol.bxdf_batch(kind, params, outgoing, sample)
    for g, w in zip(got, want):
        same = (bits(g) == bits(w)) | (np.isnan(g) & np.isnan(w))
        assert same.all(), f"{(~same).sum()} of {same.size} values differ"


@pytest.mark.parametrize("check_type,kind,params", BXDF_TABLE[::3])
def test_device_bxdf_properties(check_type, kind, params):
    """BxDFTests.Sample's properties (sample == evaluate / pdf, energy <= 1.02, reciprocity) on the device lobes."""
    assert run_bxdf_checks(device_bxdf_batch, check_type, kind, params) > 0


def test_device_fastmath_bit_identical_to_oracle():
    values = np.array(FLOAT_VALUES + list(np.linspace(-3, 3, 4001)), dtype=np.float32)
    for op in range(9):
        got = device_math(op, values)
        want = np.array([ol.fastmath(op, v) for v in values], dtype=np.float32)
        assert ((bits(got) == bits(want)) | (np.isnan(got) & np.isnan(want))).all(), op
    for op in (10, 11):
        got = device_math(op, values)
        want = np.array([ol.fastmath(op, v) for v in values], dtype=np.float32)
        assert np.array_equal(got, want)

    rng = np.random.default_rng(1)
    a, b, c = (rng.normal(size=5000).astype(np.float32) for _ in range(3))
    want = np.array([ol.fastmath(ol.FM_FMA, x, y, z) for x, y, z in zip(a, b, c)], dtype=np.float32)
    assert np.array_equal(bits(device_math(9, a, b, c)), bits(want))


def test_device_sincos_and_sample_sequence_bit_identical_to_oracle():
    angles = np.concatenate([np.linspace(-2 * np.pi, 2 * np.pi, 50001), np.linspace(-1024, 1024, 5001)]).astype(np.float32)
    want = np.array([ol.sincos(x) for x in angles], dtype=np.float32)
    assert np.array_equal(bits(device_math(100, angles)), bits(want[:, 0]))
    assert np.array_equal(bits(device_math(101, angles)), bits(want[:, 1]))

    seeds = np.arange(1, 2001, dtype=np.uint32)
    pixels = (seeds * 977) % 5003
    samples = (seeds * 31) % 257
    got = device_math(102, seeds.view(np.float32), pixels.astype(np.uint32).view(np.float32), samples.astype(np.uint32).view(np.float32))
    lib = ol.library()
    want = np.array([lib.oracle_sample_value(int(s), int(p), int(k), 0) for s, p, k in zip(seeds, pixels, samples)], dtype=np.float32)
    assert np.array_equal(bits(got), bits(want))
