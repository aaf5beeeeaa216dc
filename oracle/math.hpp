// oracle/math.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md). CPU restatement of Echo's scalar/vector math.
//
// Every function cites the reference file:line it follows (paths relative to /root/reference/src/Echo.Core/).
// Build with -ffp-contract=off: RyuJIT never contracts a*b+c, fused ops appear only where the reference calls
// FastMath.FMA / OneMinus2 (Fma.IsSupported hosts), which are explicit std::fma here.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>

namespace oracle
{

constexpr float kInfinity = std::numeric_limits<float>::infinity();

// Common/Mathematics/Scalars.cs:15-45 (decimal constants rounded to float32)
constexpr float kPi = 3.14159265358979323846f;
constexpr float kPiR = 0.31830988618379067154f;
constexpr float kTau = 6.28318530717958647692f;
constexpr float kTauR = 0.15915494309189533577f;
constexpr float kPhi = 1.61803398874989484820f;
constexpr float kRoot2 = 1.41421356237309504880f;

// Common/Mathematics/FastMath.cs:32,37
constexpr float kEpsilon = 8E-7f;
constexpr float kOneMinusEpsilon = 0.99999994f;

inline uint32_t float_bits(float value)
{
	uint32_t bits;
	std::memcpy(&bits, &value, 4);
	return bits;
}

inline float bits_float(uint32_t bits)
{
	float value;
	std::memcpy(&value, &bits, 4);
	return value;
}

// ---- FastMath.cs: SSE scalar min/max return the SECOND operand when either is NaN (minss/maxss) ----
inline float sse_min(float a, float b) { return a < b ? a : b; } // FastMath.cs:57-62, Float4.cs:376
inline float sse_max(float a, float b) { return a > b ? a : b; } // FastMath.cs:69-74, Float4.cs:377
inline float max0(float v) { return sse_max(0.0f, v); }          // FastMath.cs:45-49
inline float clamp01(float v) { return sse_min(1.0f, sse_max(0.0f, v)); }   // FastMath.cs:81-87
inline float clamp11(float v) { return sse_min(1.0f, sse_max(-1.0f, v)); }  // FastMath.cs:94-101
inline float clamp_epsilon(float v) { return sse_min(kOneMinusEpsilon, sse_max(0.0f, v)); } // FastMath.cs:108-114
inline float fabs_bits(float v) { return bits_float(float_bits(v) & 0x7FFFFFFFu); }         // FastMath.cs:121-127

// FastMath.cs:134-140: (value <= 0) ? +0 : sqrt(value); NaN passes through sqrt
inline float sqrt0(float v) { return v <= 0.0f ? 0.0f : std::sqrt(v); }
inline float sqrt_r0(float v) { return 1.0f / sqrt0(v); }                 // FastMath.cs:148
inline float fma_f(float a, float b, float c) { return std::fma(a, b, c); } // FastMath.cs:180-184 (FMA3 host)
inline float one_minus2(float v) { return std::fma(-v, v, 1.0f); }        // FastMath.cs:155-161 (vfnmadd: -(v*v)+1)
inline float identity(float v) { return sqrt0(one_minus2(v)); }           // FastMath.cs:170
inline bool positive(float v, float epsilon = kEpsilon) { return epsilon <= v; } // FastMath.cs:204
inline bool almost_zero(float v, float epsilon = kEpsilon)                // FastMath.cs:210-217
{
	return (float_bits(v) << 1) < (float_bits(epsilon) << 1);
}

// Math.Min/Math.Max (System.Math, NaN-propagating, -0 < +0); used by Float3.Min/Max (Float3.cs:301-302)
inline float math_min(float a, float b)
{
	if (a != a) return a;
	if (b != b) return b;
	if (a == b) return std::signbit(a) ? a : b;
	return a < b ? a : b;
}

inline float math_max(float a, float b)
{
	if (a != a) return a;
	if (b != b) return b;
	if (a == b) return std::signbit(a) ? b : a;
	return a > b ? a : b;
}

// Deterministic sine/cosine shared bit-for-bit with the CUDA path (echorenderer_b200/csrc/echo_device_math.cuh).
// The reference calls MathF.SinCos (FastMath.cs:190-198), whose bits depend on the platform CRT; this restatement
// pins one implementation (Cody-Waite reduction by pi/2 + degree 7/8 minimax polynomials, all in explicit fmaf)
// so that oracle and device agree exactly. |radians| <= 1024 on every hot-path call site.
inline void sincos_det(float radians, float& sin_out, float& cos_out)
{
	float q = std::nearbyint(radians * 0.6366197466850281f); // round to nearest even, same as rintf on device
	int quadrant = (int)q;

	float r = std::fma(q, -1.5707963705062866f, radians);   // pi/2 split in three floats
	r = std::fma(q, 4.371138828673793e-08f, r);
	r = std::fma(q, 1.7151245100058819e-15f, r);

	float r2 = r * r;

	float s = std::fma(r2, -1.9515295891e-4f, 8.3321608736e-3f);
	s = std::fma(s, r2, -1.6666654611e-1f);
	s = std::fma(s * r2, r, r);

	float c = std::fma(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
	c = std::fma(c, r2, 4.166664568298827e-2f);
	c = std::fma(c, r2, -0.5f);
	c = std::fma(c, r2, 1.0f);

	if (quadrant & 1)
	{
		float t = s;
		s = c;
		c = t;
	}

	if (quadrant & 2) s = -s;
	if ((quadrant + 1) & 2) c = -c;

	sin_out = s;
	cos_out = c;
}

// MathF.Atan2 / Asin / Acos are CRT functions like SinCos: pinned to the Cephes single-precision algorithms (atanf, asinf;
// ~2e-7 relative error), every operation a separately rounded IEEE op so that oracle and device agree exactly. They are
// only reached through textures (sphere texture coordinates, SphereEntity.cs:236-245; CylindricalTexture.ToUV).
inline float atan_det(float value)
{
	float sign = value < 0.0f ? -1.0f : 1.0f;
	float x = value < 0.0f ? -value : value;
	float y;

	if (x > 2.414213562373095f) // tan(3 pi / 8)
	{
		y = 1.5707963267948966f;
		x = -(1.0f / x);
	}
	else if (x > 0.4142135623730950f) // tan(pi / 8)
	{
		y = 0.7853981633974483f;
		x = (x - 1.0f) / (x + 1.0f);
	}
	else y = 0.0f;

	float z = x * x;
	float p = 8.05374449538e-2f * z - 1.38776856032e-1f;
	p = p * z + 1.99777106478e-1f;
	p = p * z - 3.33329491539e-1f;
	y = y + (p * z * x + x);
	return sign * y;
}

inline float atan2_det(float y, float x)
{
	if (x == 0.0f)
	{
		if (y == 0.0f) return 0.0f;
		return y > 0.0f ? 1.5707963267948966f : -1.5707963267948966f;
	}

	if (y == 0.0f) return x < 0.0f ? 3.14159265358979323846f : 0.0f;

	float w = x > 0.0f ? 0.0f : (y < 0.0f ? -3.14159265358979323846f : 3.14159265358979323846f);
	return w + atan_det(y / x);
}

inline float asin_det(float value) // |value| <= 1 (callers clamp)
{
	float sign = value < 0.0f ? -1.0f : 1.0f;
	float a = value < 0.0f ? -value : value;
	if (a < 1.0e-4f) return value;

	bool large = a > 0.5f;
	float z, x;

	if (large)
	{
		z = 0.5f * (1.0f - a);
		x = std::sqrt(z);
	}
	else
	{
		x = a;
		z = x * x;
	}

	float p = 4.2163199048e-2f * z + 2.4181311049e-2f;
	p = p * z + 4.5470025998e-2f;
	p = p * z + 7.4953002686e-2f;
	p = p * z + 1.6666752422e-1f;
	float result = p * z * x + x;

	if (large) result = 1.5707963267948966f - (result + result);
	return sign * result;
}

inline float acos_det(float value) { return 1.5707963267948966f - asin_det(value); } // |value| <= 1

struct Float2
{
	float x, y;
};

// Common/Packed/Float3.cs
struct Float3
{
	float x, y, z;

	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

inline Float3 operator+(Float3 a, Float3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; } // Float3.cs:506
inline Float3 operator-(Float3 a, Float3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; } // Float3.cs:507
inline Float3 operator*(Float3 a, float b) { return { a.x * b, a.y * b, a.z * b }; }        // Float3.cs:512
inline Float3 operator*(float a, Float3 b) { return { a * b.x, a * b.y, a * b.z }; }        // Float3.cs:515
inline Float3 operator/(Float3 a, float b) { return { a.x / b, a.y / b, a.z / b }; }        // Float3.cs:513
inline Float3 operator-(Float3 a) { return { -a.x, -a.y, -a.z }; }                          // Float3.cs:519

inline float dot(Float3 a, Float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }          // Float3.cs:275
inline float squared_magnitude(Float3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }      // Float3.cs:42-46

inline double squared_magnitude_double(Float3 a)                                            // Float3.cs:48-52
{
	return (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z;
}

inline float magnitude(Float3 a) { return (float)std::sqrt(squared_magnitude_double(a)); }  // Float3.cs:30-40

// Float3.cs:268-273: each component in fp64, rounded once to fp32
inline Float3 cross(Float3 a, Float3 b)
{
	return {
		(float)((double)a.y * b.z - (double)a.z * b.y),
		(float)((double)a.z * b.x - (double)a.x * b.z),
		(float)((double)a.x * b.y - (double)a.y * b.x)
	};
}

// Common/Mathematics/Scalars.cs:172-186 with value=x, other=0, epsilon=1e-10
inline bool almost_equals_zero(double value)
{
	if (value == 0.0) return true;
	const double Normal = 2.2250738585072014e-308;
	return std::fabs(value) < 1E-10 * Normal; // Scalars.cs:180 (other == 0 branch)
}

// Float3.cs:171-181
inline Float3 normalized(Float3 a)
{
	double squared = squared_magnitude_double(a);
	if (almost_equals_zero(squared)) return { 0.0f, 0.0f, 0.0f };
	return (1.0f / (float)std::sqrt(squared)) * a;
}

inline Float3 float3_min(Float3 a, Float3 b) { return { math_min(a.x, b.x), math_min(a.y, b.y), math_min(a.z, b.z) }; }
inline Float3 float3_max(Float3 a, Float3 b) { return { math_max(a.x, b.x), math_max(a.y, b.y), math_max(a.z, b.z) }; }

inline Float3 reflect(Float3 value, Float3 normal) { return 2.0f * dot(value, normal) * normal - value; } // Float3.cs:340

// Float4x4.MultiplyDirection with the identity matrix (PreparedScene.cs:100-101 on the root instance;
// Float4x4.cs:267-272): kept literal because 1*x + 0*y + 0*z can change the sign of a zero component.
inline Float3 identity_multiply_direction(Float3 d)
{
	return {
		1.0f * d.x + 0.0f * d.y + 0.0f * d.z,
		0.0f * d.x + 1.0f * d.y + 0.0f * d.z,
		0.0f * d.x + 0.0f * d.y + 1.0f * d.z
	};
}

// Float4x4.MultiplyPoint / MultiplyDirection (Float4x4.cs:260-272) on rows 0..2 of an affine matrix, row-major m[12];
// plain left-to-right products and sums, as the C# JIT emits them (no contraction)
inline Float3 multiply_point(const float* m, Float3 p)
{
	return {
		m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3],
		m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
		m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]
	};
}

inline Float3 multiply_direction(const float* m, Float3 d)
{
	return {
		m[0] * d.x + m[1] * d.y + m[2] * d.z,
		m[4] * d.x + m[5] * d.y + m[6] * d.z,
		m[8] * d.x + m[9] * d.y + m[10] * d.z
	};
}

// Utility.GetScale (Utility.cs:82): Float4 Magnitude of row 0 with W = 0 — SqrtScalar of (x*x + y*y) + (z*z + 0) (Float4.cs:51-61,73-81)
inline float get_scale(const float* m) { return std::sqrt((m[0] * m[0] + m[1] * m[1]) + (m[2] * m[2] + 0.0f)); }

// Textures/Colors/RGB128.cs (a Float4 with W == 0)
struct RGB
{
	float r, g, b;
};

constexpr float kWeightR = 0.212671f, kWeightG = 0.715160f, kWeightB = 0.072169f; // RGB128.cs:30-32

inline RGB operator+(RGB a, RGB b) { return { a.r + b.r, a.g + b.g, a.b + b.b }; }
inline RGB operator*(RGB a, RGB b) { return { a.r * b.r, a.g * b.g, a.b * b.b }; }
inline RGB operator*(RGB a, float b) { return { a.r * b, a.g * b, a.b * b }; }
inline RGB operator/(RGB a, float b) { return { a.r / b, a.g / b, a.b / b }; }     // RGB128.cs:109 (Sse.Divide by broadcast)
inline RGB operator/(RGB a, RGB b) { return { a.r / b.r, a.g / b.g, a.b / b.b }; } // RGB128.cs:110-117 (W lane divides by 1)

// RGB128.cs:38 + Float4.cs:73-81,359: (r*wr + g*wg) + (b*wb + 0*0)
inline float luminance(RGB c) { return (c.r * kWeightR + c.g * kWeightG) + (c.b * kWeightB + 0.0f * 0.0f); }

// RGB128.cs:40-51
inline bool is_zero(RGB c) { return c.r < kEpsilon / kWeightR && c.g < kEpsilon / kWeightG && c.b < kEpsilon / kWeightB; }

// RGB128.cs:79-86
inline RGB max_epsilon(RGB c)
{
	return { sse_max(c.r, kEpsilon / kWeightR), sse_max(c.g, kEpsilon / kWeightG), sse_max(c.b, kEpsilon / kWeightB) };
}

constexpr RGB kBlack = { 0.0f, 0.0f, 0.0f };
constexpr RGB kWhite = { 1.0f, 1.0f, 1.0f };

// Common/Mathematics/Primitives/OrthonormalTransform.cs:12-66
struct OrthonormalTransform
{
	Float3 axisX, axisY, axisZ;

	explicit OrthonormalTransform(Float3 z) : axisZ(z)
	{
		if (almost_zero(z.x) && almost_zero(z.y))
		{
			axisX = { 1.0f, 0.0f, 0.0f };
			axisY = z.z > 0.0f ? Float3{ 0.0f, 1.0f, 0.0f } : Float3{ 0.0f, -1.0f, 0.0f };
		}
		else
		{
			axisX = normalized(Float3{ z.y, -z.x, 0.0f });
			axisY = cross(z, axisX);
		}
	}

	Float3 apply_forward(Float3 d) const
	{
		return {
			axisX.x * d.x + axisY.x * d.y + axisZ.x * d.z,
			axisX.y * d.x + axisY.y * d.y + axisZ.y * d.z,
			axisX.z * d.x + axisY.z * d.y + axisZ.z * d.z
		};
	}

	Float3 apply_inverse(Float3 d) const
	{
		return {
			axisX.x * d.x + axisX.y * d.y + axisX.z * d.z,
			axisY.x * d.x + axisY.y * d.y + axisY.z * d.z,
			axisZ.x * d.x + axisZ.y * d.y + axisZ.z * d.z
		};
	}
};

// Evaluation/Sampling/Sample1D.cs:13-17: every Sample1D is clamped to [0, 1 - ulp]
inline float sample1d(float u) { return clamp_epsilon(u); }

// Sample1D.cs:31-35,52-56
inline int sample_range(float u, int max) { return (int)(u * (float)max); }

inline float sample_range(float u, int max, int& index)
{
	index = sample_range(u, max);
	return sample1d(fma_f(u, (float)max, -(float)index));
}

// Sample1D.cs:74-80
inline float sample_stretch(float u, float lower, float upper) { return sample1d((u - lower) / (upper - lower)); }

// Evaluation/Sampling/Sample2D.cs:160-166
inline Float2 project_disk(float radius, float angle)
{
	float s, c;
	sincos_det(angle, s, c);
	return { c * radius, s * radius };
}

// Sample2D.cs:153-158
inline Float3 project_sphere(float z, float u)
{
	float radius = identity(z);
	float angle = kTau * u;
	Float2 disk = project_disk(radius, angle);
	return { disk.x, disk.y, z };
}

// Sample2D.cs:35
inline Float3 uniform_sphere(Float2 s) { return project_sphere(fma_f(s.x, -2.0f, 1.0f), s.y); }

// Sample2D.cs:123-127
inline Float3 uniform_cone(Float2 s, float cosMaxP) { return project_sphere(fma_f(cosMaxP - 1.0f, s.x, 1.0f), s.y); }

// Sample2D.cs:54-62
inline Float2 uniform_triangle(Float2 s)
{
	float v = sqrt0(s.x);
	return { 1.0f - v, s.y * v };
}

// Sample2D.cs:66-91
inline Float2 concentric_disk(Float2 s)
{
	float xValue = fma_f(s.x, 2.0f, -1.0f);
	float yValue = fma_f(s.y, 2.0f, -1.0f);

	if (almost_zero(xValue) && almost_zero(yValue)) return { 0.0f, 0.0f };

	float radius, angle;

	if (fabs_bits(xValue) > fabs_bits(yValue))
	{
		radius = xValue;
		angle = kPi / 4.0f * yValue / xValue;
	}
	else
	{
		radius = yValue;
		angle = fma_f(xValue / yValue, kPi / -4.0f, kPi / 2.0f);
	}

	return project_disk(radius, angle);
}

// Sample2D.cs:98-106
inline Float3 cosine_hemisphere(Float2 s)
{
	Float2 disk = concentric_disk(s);
	float z = disk.x * disk.x + disk.y * disk.y;
	return { disk.x, disk.y, sqrt0(1.0f - z) };
}

constexpr float kUniformSpherePdf = kTauR / 2.0f;                       // Sample2D.cs:108-110
inline float uniform_cone_pdf(float cosMaxP) { return kTauR / (1.0f - cosMaxP); } // Sample2D.cs:146

// The counter-based sample sequence shared with the device (DESIGN.md "Sample sequence"). The reference draws from
// System.Random through ContinuousDistribution (Evaluation/Sampling/ContinuousDistribution.cs:46), which cannot be
// reproduced outside .NET; both this oracle and the CUDA path use this hash instead, in the reference's draw order.
inline uint32_t hash32(uint32_t x)
{
	x ^= x >> 16;
	x *= 0x7FEB352Du;
	x ^= x >> 15;
	x *= 0x846CA68Bu;
	x ^= x >> 16;
	return x;
}

inline uint32_t sample_key(uint32_t seed, uint32_t pixel, uint32_t sample)
{
	uint32_t h = hash32(seed ^ 0x9E3779B9u);
	h = hash32(h + pixel * 0x85EBCA6Bu + 0x165667B1u);
	h = hash32(h ^ (sample * 0xC2B2AE35u + 0x27D4EB2Fu));
	return h;
}

inline float sample_value(uint32_t key, uint32_t dimension)
{
	uint32_t h = hash32(key + dimension * 0x9E3779B1u);
	h = hash32(h ^ 0x68E31DA4u);
	return sample1d((float)(h >> 8) * 5.9604644775390625e-8f); // 24 bits -> [0, 1)
}

} // namespace oracle
