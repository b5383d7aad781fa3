"""CPU: the closed-form approximations the kernels use in place of library transcendentals, restated in float32 numpy
with the kernels' constants, against float64 references.  The bounds are the ones DESIGN.md / the kernel comments state;
all sit below the fp16 rounding (2^-11 = 4.9e-4 relative) of the tensors they produce."""
import math

import numpy as np


def _f32(x):
    return np.asarray(x, np.float32)


def exp2_poly(x):
    """attn.cu::exp2_poly / exp2_poly2: clamp, Cody-Waite split with the 1.5 * 2^23 constant, cubic on [-0.5, 0.5],
    exponent re-inserted with an integer multiply-add."""
    x = np.maximum(_f32(x), _f32(-30.0))
    magic = _f32(12582912.0)
    t = _f32(x + magic)
    f = _f32(x - _f32(t - magic))
    p = _f32(_f32(0.0550081) * f + _f32(0.24220917))
    p = _f32(p * f + _f32(0.69328282))
    p = _f32(p * f + _f32(1.0))
    bits = (t.view(np.int32).astype(np.int64) * 8388608 + p.view(np.int32).astype(np.int64)) & 0xFFFFFFFF
    return bits.astype(np.uint32).view(np.float32)


def test_polynomial_exp2_is_within_1e4_relative():
    x = np.linspace(-29.5, 10.0, 400001, dtype=np.float32)
    got = exp2_poly(x).astype(np.float64)
    ref = np.exp2(x.astype(np.float64))
    rel = np.abs(got - ref) / ref
    assert rel.max() < 1.2e-4, rel.max()
    assert exp2_poly(np.array([-np.inf, -1e30], np.float32)).max() <= 2.0 ** -29   # masked columns -> 0 in fp16
    assert np.float16(exp2_poly(np.array([-np.inf], np.float32))[0]) == 0


def test_sigmoid_form_of_tanh_gelu_is_the_same_function():
    """gemm.cu::gelu_tanh_f: 0.5 x (1 + tanh(u)) == x / (1 + 2^(-2 u log2 e)), u = sqrt(2/pi) (x + 0.044715 x^3)."""
    x = np.linspace(-12, 12, 200001)
    u = math.sqrt(2 / math.pi) * (x + 0.044715 * x ** 3)
    ref = 0.5 * x * (1 + np.tanh(u))
    a = -2.0 * 0.7978845608028654 * 1.4426950408889634
    z = _f32(x) * _f32(_f32(a * 0.044715) * _f32(_f32(x) * _f32(x)) + _f32(a))
    with np.errstate(over="ignore"):   # 2^z overflows to inf for very negative x: x / inf = -0, as in the kernel
        got = _f32(x) / _f32(1.0 + np.exp2(z.astype(np.float64)))
    assert np.abs(got - ref).max() < 2e-6 * max(1.0, np.abs(ref).max())


def test_abramowitz_stegun_erf_gelu_error():
    """gemm.cu::gelu_erf_f: erf by A&S 7.1.26 (|error| <= 1.5e-7) inside 0.5 x (1 + erf(x / sqrt 2))."""
    x = np.linspace(-10, 10, 200001)
    ref = 0.5 * x * (1 + np.vectorize(math.erf)(x / math.sqrt(2)))
    az = np.abs(x) * 0.7071067811865476
    t = 1.0 / (0.3275911 * az + 1.0)
    pl = ((((1.061405429 * t - 1.453152027) * t + 1.421413741) * t - 0.284496736) * t + 0.254829592) * t
    erf_abs = 1.0 - pl * np.exp(-az * az)
    got = 0.5 * x + 0.5 * np.abs(x) * erf_abs
    assert np.abs(got - ref).max() < 1.0e-6          # 0.5 |x| * 1.5e-7 at |x| <= 10


def test_half_operations_equal_round_of_fp32_operations():
    """rowops.cu::layernorm_kernel does the adaLN epilogue with half2 instructions where the reference rounds an fp32
    result to fp16 after every operation: for half operands the two coincide (sum / product of two halfs is exact in
    fp32 up to one rounding)."""
    rng = np.random.default_rng(0)
    a = rng.normal(size=200000).astype(np.float16)
    b = (rng.normal(size=200000) * 0.3).astype(np.float16)
    one = np.float16(1.0)
    assert np.array_equal((one + b), (np.float32(1.0) + b.astype(np.float32)).astype(np.float16))
    assert np.array_equal((a * b), (a.astype(np.float32) * b.astype(np.float32)).astype(np.float16))
    assert np.array_equal((a + b), (a.astype(np.float32) + b.astype(np.float32)).astype(np.float16))
