// echo_shading.cuh — device BSDFs: the in-scope subset of Echo.Core/Evaluation/Scattering + Evaluation/Materials as a
// register-resident tagged record instead of arena-allocated lobe objects (the reference allocates a BSDF and its BxDFs
// per hit, Material.cs:117-122). A Bsdf has at most two lobes; `kind` selects the lobe set, `KINDS` (a compile-time
// bit mask) lets each material-sorted shade kernel drop the code of lobes its queue can never contain.
// Arithmetic order follows the reference line by line; citations are relative to src/Echo.Core/Evaluation/.
#pragma once
#include "echo_scene.cuh"

namespace echo
{

// Scattering/FunctionType.cs
enum : int { FT_REFLECTIVE = 1, FT_TRANSMISSIVE = 2, FT_DIFFUSE = 4, FT_GLOSSY = 8, FT_SPECULAR = 16 };

enum BsdfKind : int
{
	BSDF_EMPTY = 0,                // Emissive: no lobes, tint 0 (Materials/Emissive.cs:56)
	BSDF_LAMBERT_REFLECTION = 1,   // Diffuse, roughness ~ 0 (Materials/Diffuse.cs:41)
	BSDF_LAMBERT_TWO_SIDED = 2,    // Diffuse.Transmissive (Materials/Diffuse.cs:44)
	BSDF_OREN_NAYAR = 3,           // Diffuse, rough (Materials/Diffuse.cs:42)
	BSDF_DIELECTRIC_GLOSSY = 4,    // GlossyReflection<TR, RealFresnel> + GlossyTransmission<TR> (Materials/Dielectric.cs:39-44)
	BSDF_DIELECTRIC_SPECULAR = 5,  // SpecularFresnel (Materials/Dielectric.cs:46)
	BSDF_CONDUCTOR_GLOSSY = 6,     // GlossyReflection<TR, ComplexFresnel> (Materials/Conductor.cs:117-119)
	BSDF_CONDUCTOR_SPECULAR = 7,   // SpecularReflection<ComplexFresnel> (Materials/Conductor.cs:121)
	BSDF_INVISIBLE = 8,            // SpecularTransmission(1 / 1) (Materials/Invisible.cs:22-26)
	BSDF_COATED_DIFFUSE = 9,       // CoatedLambertianReflection + GlossyReflection<TR, RealFresnel> (Materials/CoatedDiffuse.cs:37-55)
};

#define kind_bit(kind) (1u << (kind))
constexpr uint32_t KINDS_ALL = 0x3FFu;
constexpr uint32_t KINDS_DIFFUSE = kind_bit(BSDF_LAMBERT_REFLECTION) | kind_bit(BSDF_LAMBERT_TWO_SIDED) | kind_bit(BSDF_OREN_NAYAR) | kind_bit(BSDF_INVISIBLE);
constexpr uint32_t KINDS_DIELECTRIC = kind_bit(BSDF_DIELECTRIC_GLOSSY) | kind_bit(BSDF_DIELECTRIC_SPECULAR) | kind_bit(BSDF_COATED_DIFFUSE) | kind_bit(BSDF_INVISIBLE);
constexpr uint32_t KINDS_SMOOTH = kind_bit(BSDF_DIELECTRIC_SPECULAR) | kind_bit(BSDF_INVISIBLE); // Dielectric with both alphas specular: SpecularFresnel only
constexpr uint32_t KINDS_CONDUCTOR = kind_bit(BSDF_CONDUCTOR_GLOSSY) | kind_bit(BSDF_CONDUCTOR_SPECULAR) | kind_bit(BSDF_INVISIBLE);
constexpr uint32_t KINDS_TERMINAL = kind_bit(BSDF_EMPTY) | kind_bit(BSDF_INVISIBLE);

struct Bsdf
{
	int kind;
	rgb tint;
	frame transform;     // OrthonormalTransform of the shading normal (Scattering/BSDF.cs:34)
	vec3 geometricNormal;

	float alphaX, alphaY;         // TrowbridgeReitzMicrofacet.alpha
	float etaAbove, etaBelow;     // RealFresnel
	rgb eta2, etaK2;              // ComplexFresnel
	float orenA, orenB;           // OrenNayar a, b
	rgb coatedMultiplier;         // CoatedLambertianReflection.multiplier
};

struct Sampled // Probable<RGB128>
{
	rgb content;
	float pdf;
};

ECHO_DEVICE Sampled impossible() { return { { 0.0f, 0.0f, 0.0f }, 0.0f }; }

// ---- Scattering/BxDF.cs:44-135 ----
ECHO_DEVICE float cosine_p(vec3 d) { return d.z; }
ECHO_DEVICE float cosine_p2(vec3 d) { return d.z * d.z; }
ECHO_DEVICE float sine_p2(vec3 d) { return one_minus2(d.z); }

ECHO_DEVICE float cosine_t2(vec3 d)
{
	float sin2 = sine_p2(d);
	if (almost_zero(sin2)) return 1.0f;
	return clamp01(div(d.x * d.x, sin2));
}

ECHO_DEVICE float sine_t2(vec3 d)
{
	float sin2 = sine_p2(d);
	if (almost_zero(sin2)) return 0.0f;
	return clamp01(div(d.y * d.y, sin2));
}

ECHO_DEVICE bool flat_or_same_hemisphere(vec3 a, vec3 b) { return !positive(-cosine_p(a) * cosine_p(b)); }
ECHO_DEVICE bool flat_or_opposite_hemisphere(vec3 a, vec3 b) { return !positive(cosine_p(a) * cosine_p(b)); }
ECHO_DEVICE vec3 negate_z(vec3 v) { return { v.x, v.y, -v.z }; }

// ---- Scattering/Fresnel.cs:13-150 (RealFresnel + Packet) ----
struct FresnelPacket
{
	float etaOutgoing, etaIncident, cosOutgoing, cosIncident;
};

ECHO_DEVICE FresnelPacket fresnel_incomplete(float etaAbove, float etaBelow, float cosOutgoing) // :37-40,50-61
{
	FresnelPacket p;
	bool above = cosOutgoing > 0.0f;
	p.etaOutgoing = above ? etaAbove : etaBelow;
	p.etaIncident = above ? etaBelow : etaAbove;
	p.cosOutgoing = clamp11(cosOutgoing);
	p.cosIncident = 0.0f;
	return p;
}

ECHO_DEVICE void fresnel_complete(FresnelPacket& p) // CalculateCosineIncident :121-130, then the Packet ctor's Clamp11
{
	float eta = div(p.etaOutgoing, p.etaIncident);
	float sinO2 = one_minus2(p.cosOutgoing);
	float sinI2 = eta * eta * sinO2;

	float result;

	if (sinI2 >= 1.0f) result = 0.0f;
	else
	{
		result = sqrt0(1.0f - sinI2);
		result = p.cosOutgoing > 0.0f ? -result : result;
	}

	p.cosIncident = clamp11(result);
}

ECHO_DEVICE bool fresnel_total_internal(const FresnelPacket& p) { return almost_zero(p.cosIncident); }

ECHO_DEVICE float fresnel_value(const FresnelPacket& p) // :86-107
{
	if (fresnel_total_internal(p)) return 1.0f;

	float cosO = abs_bits(p.cosOutgoing);
	float cosI = abs_bits(p.cosIncident);

	float para0 = p.etaIncident * cosO;
	float para1 = p.etaOutgoing * cosI;
	float perp0 = p.etaOutgoing * cosO;
	float perp1 = p.etaIncident * cosI;

	float para = div(para0 - para1, para0 + para1);
	float perp = div(perp0 - perp1, perp0 + perp1);
	return div(para * para + perp * perp, 2.0f);
}

ECHO_DEVICE vec3 fresnel_refract(const FresnelPacket& p, vec3 outgoing, vec3 normal) // :111-119
{
	float eta = div(p.etaOutgoing, p.etaIncident);
	return normalized(normal * (eta * p.cosOutgoing + p.cosIncident) - eta * outgoing);
}

ECHO_SHARED_CODE float real_fresnel(float etaAbove, float etaBelow, float cosO) // :28-35
{
	FresnelPacket p = fresnel_incomplete(etaAbove, etaBelow, cosO);
	fresnel_complete(p);
	return fresnel_value(p);
}

// ---- Scattering/Fresnel.cs:152-197 (ComplexFresnel), one colour lane ----
ECHO_DEVICE float complex_fresnel_lane(float eta2, float etaK2, float cosO, float cosO2, float sinO2)
{
	float term = eta2 - etaK2 - sinO2;
	float a2b2 = sqrt0(term * term + 4.0f * eta2 * etaK2);

	float para0 = a2b2 + cosO2;
	float para1 = cosO * kRoot2 * sqrt0(a2b2 + term);

	float perp0 = cosO2 * a2b2 + sinO2 * sinO2;
	float perp1 = para1 * sinO2;

	float para = div(para0 - para1, para0 + para1);
	float perp = div(perp0 - perp1, perp0 + perp1);

	return div(para * perp + para, 2.0f);
}

ECHO_DEVICE rgb complex_fresnel(rgb eta2, rgb etaK2, float cosO)
{
	cosO = clamp01(abs_bits(cosO));
	float cosO2 = cosO * cosO;
	float sinO2 = 1.0f - cosO2;

	return { complex_fresnel_lane(eta2.r, etaK2.r, cosO, cosO2, sinO2),
	         complex_fresnel_lane(eta2.g, etaK2.g, cosO, cosO2, sinO2),
	         complex_fresnel_lane(eta2.b, etaK2.b, cosO, cosO2, sinO2) };
}

ECHO_DEVICE void complex_fresnel_setup(rgb etaAbove, rgb etaBelow, rgb extinction, rgb& eta2, rgb& etaK2) // :154-166
{
	rgb etaAboveR = { rcp(etaAbove.r), rcp(etaAbove.g), rcp(etaAbove.b) };
	eta2 = etaBelow * etaAboveR;
	etaK2 = extinction * etaAboveR;
	eta2 = eta2 * eta2;
	etaK2 = etaK2 * etaK2;
}

// ---- Scattering/IMicrofacet.cs ----
ECHO_DEVICE float microfacet_alpha(float roughness, bool& specular) // :43-51
{
	roughness = clamp01(roughness * 0.75f);
	const float Threshold = 0.0001f;
	float alpha = roughness * roughness;
	specular = alpha < Threshold;
	return specular ? Threshold : alpha;
}

ECHO_SHARED_CODE float tr_projected_area(float alphaX, float alphaY, vec3 normal) // :101-120
{
	float cos2 = cosine_p2(normal);
	if (!positive(cos2)) return 0.0f;

	float sum = cos2;

	if (positive(1.0f - cos2, 1E-5f))
	{
		float x = div(normal.x, alphaX), y = div(normal.y, alphaY);
		sum += x * x + y * y;
	}

	return rcp(sum * sum * (alphaX * alphaY) * kPi);
}

ECHO_SHARED_CODE float tr_shadowing_ratio(float alphaX, float alphaY, vec3 direction) // :123-132
{
	float cos2 = cosine_p2(direction);
	if (!positive(cos2)) return 0.0f;
	float tan2 = div(sine_p2(direction), cos2);

	float thetaX = cosine_t2(direction), thetaY = sine_t2(direction);
	float alpha2Tan2 = (alphaX * alphaX * thetaX + alphaY * alphaY * thetaY) * tan2;
	return div(sqrt0(1.0f + alpha2Tan2), 2.0f) - 0.5f;
}

ECHO_DEVICE float tr_visibility(float ax, float ay, vec3 direction) { return rcp(1.0f + tr_shadowing_ratio(ax, ay, direction)); } // :69-70

ECHO_DEVICE float tr_visibility(float ax, float ay, vec3 outgoing, vec3 incident) // :72-73
{
	return rcp(1.0f + tr_shadowing_ratio(ax, ay, outgoing) + tr_shadowing_ratio(ax, ay, incident));
}

ECHO_DEVICE float tr_probability_density(float ax, float ay, vec3 outgoing, vec3 normal) // :75-79
{
	float fraction = tr_projected_area(ax, ay, normal) * tr_visibility(ax, ay, outgoing);
	return fraction * abs_bits(div(dot(outgoing, normal), cosine_p(outgoing)));
}

ECHO_SHARED_CODE vec3 tr_sample(float alphaX, float alphaY, vec3 outgoing, vec2 sample) // :137-173, Heitz 2017 VNDF
{
	vec3 scaled = normalized(vec3{ outgoing.x * alphaX, outgoing.y * alphaY, outgoing.z });
	if (scaled.z < 0.0f) scaled = -scaled;

	float threshold = rcp(1.0f + scaled.z);
	float radius = sqrt0(sample.x);
	float theta = sample.y < threshold ? div(sample.y, threshold) : 1.0f + div(sample.y - threshold, 1.0f - threshold);

	float sin, cos;
	sincos_det(theta * -kPi, sin, cos);

	float pointX = radius * cos;
	float pointY = radius * sin;

	if (sample.y >= threshold) pointY *= scaled.z;
	float pointZ = sqrt0(1.0f - (pointX * pointX + pointY * pointY));

	frame transform = make_frame(scaled);
	vec3 transformed = apply_forward(transform, vec3{ pointX, pointY, pointZ });

	return normalized(vec3{ transformed.x * alphaX, transformed.y * alphaY, max_sse(transformed.z, kEpsilon) });
}

ECHO_SHARED_CODE vec3 glossy_find_normal(vec3 outgoing, vec3 incident) // Scattering/Glossy.cs:61-71
{
	vec3 normal = outgoing + incident;
	float length2 = squared_magnitude(normal);

	if (!positive(length2)) return { 0.0f, 0.0f, 1.0f };

	normal = normal * sqrt_r0(length2);
	return normal.z < 0.0f ? -normal : normal;
}

// =====================================================================================================================
// single lobes, local (shading) space. `REAL` selects RealFresnel vs ComplexFresnel for the reflection lobe.
// =====================================================================================================================

template<bool REAL>
ECHO_DEVICE rgb lobe_fresnel(const Bsdf& b, float cosO)
{
	if (REAL) return make_rgb(real_fresnel(b.etaAbove, b.etaBelow, cosO));
	return complex_fresnel(b.eta2, b.etaK2, cosO);
}

// Glossy.cs:23-41
// shared-code forms of the two reflection lobes' Evaluate and of ProbabilityDensity: scalar arguments only (see ECHO_SHARED_CODE)
ECHO_SHARED_CODE_2 float glossy_reflection_evaluate_real(float alphaX, float alphaY, float etaAbove, float etaBelow, vec3 outgoing, vec3 incident)
{
	if (flat_or_opposite_hemisphere(outgoing, incident)) return 0.0f;
	vec3 normal = glossy_find_normal(outgoing, incident);

	float ratio = tr_projected_area(alphaX, alphaY, normal) * tr_visibility(alphaX, alphaY, outgoing, incident) * 0.25f;
	float evaluated = div(real_fresnel(etaAbove, etaBelow, dot(outgoing, normal)), cosine_p(outgoing) * cosine_p(incident));
	return evaluated * ratio;
}

ECHO_SHARED_CODE_2 rgb glossy_reflection_evaluate_complex(float alphaX, float alphaY, rgb eta2, rgb etaK2, vec3 outgoing, vec3 incident)
{
	if (flat_or_opposite_hemisphere(outgoing, incident)) return make_rgb(0.0f);
	vec3 normal = glossy_find_normal(outgoing, incident);

	float ratio = tr_projected_area(alphaX, alphaY, normal) * tr_visibility(alphaX, alphaY, outgoing, incident) * 0.25f;
	rgb evaluated = complex_fresnel(eta2, etaK2, dot(outgoing, normal)) / (cosine_p(outgoing) * cosine_p(incident));
	return evaluated * ratio;
}

template<bool REAL>
ECHO_DEVICE rgb glossy_reflection_evaluate(const Bsdf& b, vec3 outgoing, vec3 incident)
{
	if (REAL) return make_rgb(glossy_reflection_evaluate_real(b.alphaX, b.alphaY, b.etaAbove, b.etaBelow, outgoing, incident));
	return glossy_reflection_evaluate_complex(b.alphaX, b.alphaY, b.eta2, b.etaK2, outgoing, incident);
}

ECHO_SHARED_CODE_2 float glossy_reflection_pdf_scalar(float alphaX, float alphaY, vec3 outgoing, vec3 incident)
{
	if (flat_or_opposite_hemisphere(outgoing, incident)) return 0.0f;
	vec3 normal = glossy_find_normal(outgoing, incident);
	return div(tr_probability_density(alphaX, alphaY, outgoing, normal), abs_bits(dot(outgoing, normal) * 4.0f));
}

template<bool REAL>
ECHO_DEVICE float glossy_reflection_pdf(const Bsdf& b, vec3 outgoing, vec3 incident)
{
	return glossy_reflection_pdf_scalar(b.alphaX, b.alphaY, outgoing, incident);
}

// Glossy.cs:43-59
template<bool REAL>
ECHO_DEVICE Sampled glossy_reflection_sample(const Bsdf& b, vec2 sample, vec3 outgoing, vec3& incident)
{
	vec3 normal = tr_sample(b.alphaX, b.alphaY, outgoing, sample);
	incident = reflect(outgoing, normal);

	if (flat_or_opposite_hemisphere(outgoing, incident)) return impossible();

	float ratio = tr_projected_area(b.alphaX, b.alphaY, normal) * tr_visibility(b.alphaX, b.alphaY, outgoing, incident) * 0.25f;
	rgb evaluated = lobe_fresnel<REAL>(b, dot(outgoing, normal)) / (cosine_p(outgoing) * cosine_p(incident));
	float pdf = div(tr_probability_density(b.alphaX, b.alphaY, outgoing, normal), abs_bits(dot(outgoing, normal) * 4.0f));

	return { evaluated * ratio, pdf };
}

// Glossy.cs:88-113
ECHO_SHARED_CODE_2 float glossy_transmission_evaluate_scalar(float alphaX, float alphaY, float etaAbove, float etaBelow, vec3 outgoing, vec3 incident)
{
	if (flat_or_same_hemisphere(outgoing, incident)) return 0.0f;

	FresnelPacket packet = fresnel_incomplete(etaAbove, etaBelow, cosine_p(outgoing));
	float etaR = div(packet.etaIncident, packet.etaOutgoing);
	vec3 normal = glossy_find_normal(outgoing, incident * etaR);

	float dotO = dot(outgoing, normal);
	float dotI = dot(incident, normal);
	if (positive(dotO * dotI)) return 0.0f;

	float evaluated = 1.0f - real_fresnel(etaAbove, etaBelow, dotO);
	if (!positive(evaluated)) return 0.0f;

	float numerator = etaR * etaR * dotO * dotI;
	float denominator = fma_f(etaR, dotI, dotO);
	denominator *= denominator;

	if (!positive(denominator)) denominator = 1.0f;
	denominator *= cosine_p(outgoing) * cosine_p(incident);

	float ratio = tr_projected_area(alphaX, alphaY, normal) * tr_visibility(alphaX, alphaY, outgoing, incident);
	return evaluated * ratio * abs_bits(div(numerator, denominator));
}

ECHO_DEVICE rgb glossy_transmission_evaluate(const Bsdf& b, vec3 outgoing, vec3 incident)
{
	return make_rgb(glossy_transmission_evaluate_scalar(b.alphaX, b.alphaY, b.etaAbove, b.etaBelow, outgoing, incident));
}

// Glossy.cs:115-133
ECHO_SHARED_CODE_2 float glossy_transmission_pdf_scalar(float alphaX, float alphaY, float etaAbove, float etaBelow, vec3 outgoing, vec3 incident)
{
	if (flat_or_same_hemisphere(outgoing, incident)) return 0.0f;

	FresnelPacket packet = fresnel_incomplete(etaAbove, etaBelow, cosine_p(outgoing));
	float etaR = div(packet.etaIncident, packet.etaOutgoing);
	vec3 normal = glossy_find_normal(outgoing, incident * etaR);

	float dotO = dot(outgoing, normal);
	float dotI = dot(incident, normal);

	if (positive(dotO * dotI)) return 0.0f;
	float numerator = abs_bits(etaR * etaR * dotI);
	float denominator = fma_f(etaR, dotI, dotO);
	denominator *= denominator;

	if (!positive(denominator)) denominator = 1.0f;
	return tr_probability_density(alphaX, alphaY, outgoing, normal) * div(numerator, denominator);
}

ECHO_DEVICE float glossy_transmission_pdf(const Bsdf& b, vec3 outgoing, vec3 incident)
{
	return glossy_transmission_pdf_scalar(b.alphaX, b.alphaY, b.etaAbove, b.etaBelow, outgoing, incident);
}

// Glossy.cs:135-160
ECHO_DEVICE Sampled glossy_transmission_sample(const Bsdf& b, vec2 sample, vec3 outgoing, vec3& incident)
{
	vec3 normal = tr_sample(b.alphaX, b.alphaY, outgoing, sample);
	float dotO = dot(outgoing, normal);

	FresnelPacket packet = fresnel_incomplete(b.etaAbove, b.etaBelow, dotO);
	fresnel_complete(packet);

	if (fresnel_total_internal(packet))
	{
		incident = { 0.0f, 0.0f, 0.0f };
		return impossible();
	}

	incident = fresnel_refract(packet, outgoing, normal);
	float dotI = dot(incident, normal);

	if (flat_or_same_hemisphere(outgoing, incident) || positive(dotO * dotI)) return impossible();

	float etaR = div(packet.etaIncident, packet.etaOutgoing);
	float numerator = abs_bits(etaR * etaR * dotI);
	float denominator = fma_f(etaR, dotI, dotO);
	denominator *= denominator;

	if (!positive(denominator)) denominator = 1.0f;

	float ratio = tr_projected_area(b.alphaX, b.alphaY, normal) * tr_visibility(b.alphaX, b.alphaY, outgoing, incident);
	float evaluated = div(numerator * dotO, denominator * cosine_p(outgoing) * cosine_p(incident));
	float pdf = tr_probability_density(b.alphaX, b.alphaY, outgoing, normal) * div(numerator, denominator);

	return { make_rgb(1.0f - fresnel_value(packet)) * abs_bits(evaluated) * ratio, pdf };
}

// Lambertian.cs:19-40 (+ OrenNayar.Evaluate :113-124 when OREN)
template<bool OREN>
ECHO_DEVICE rgb lambert_reflection_evaluate(const Bsdf& b, vec3 outgoing, vec3 incident)
{
	if (flat_or_opposite_hemisphere(outgoing, incident)) return make_rgb(0.0f);
	if (!OREN) return make_rgb(kPiR);

	float cosO = abs_bits(cosine_p(outgoing));
	float cosI = abs_bits(cosine_p(incident));

	float s = dot(outgoing, incident) - cosO * cosI;
	if (positive(s)) s = div(s, max_sse(cosO, cosI));
	return make_rgb(b.orenA + b.orenB * s);
}

ECHO_DEVICE float lambert_reflection_pdf(vec3 outgoing, vec3 incident)
{
	if (flat_or_opposite_hemisphere(outgoing, incident)) return 0.0f;
	return abs_bits(cosine_p(incident)) * kPiR;
}

template<bool OREN>
ECHO_DEVICE Sampled lambert_reflection_sample(const Bsdf& b, vec2 sample, vec3 outgoing, vec3& incident)
{
	incident = cosine_hemisphere(sample);
	float pdf = cosine_p(incident) * kPiR;

	if (outgoing.z < 0.0f) incident = negate_z(incident);
	return { lambert_reflection_evaluate<OREN>(b, outgoing, incident), pdf };
}

// CoatedLambertianReflection.Evaluate, Lambertian.cs:152-161 (pdf and sampling are LambertianReflection's)
ECHO_DEVICE rgb coated_lambert_evaluate(const Bsdf& b, vec3 outgoing, vec3 incident)
{
	if (flat_or_opposite_hemisphere(outgoing, incident)) return make_rgb(0.0f);

	float evaluatedOutgoing = real_fresnel(b.etaAbove, b.etaBelow, abs_bits(cosine_p(outgoing)));
	float evaluatedIncident = real_fresnel(b.etaAbove, b.etaBelow, abs_bits(cosine_p(incident)));

	return b.coatedMultiplier * (1.0f - evaluatedOutgoing) * (1.0f - evaluatedIncident);
}

ECHO_DEVICE Sampled coated_lambert_sample(const Bsdf& b, vec2 sample, vec3 outgoing, vec3& incident)
{
	incident = cosine_hemisphere(sample);
	float pdf = cosine_p(incident) * kPiR;

	if (outgoing.z < 0.0f) incident = negate_z(incident);
	return { coated_lambert_evaluate(b, outgoing, incident), pdf };
}

ECHO_DEVICE void coated_lambert_setup(Bsdf& b, rgb albedo, float reflectance) // CoatedLambertianReflection.Reset, Lambertian.cs:139-147
{
	float eta = div(b.etaAbove, b.etaBelow);
	float numerator = eta * eta * kPiR;
	b.coatedMultiplier = { div(numerator, 1.0f - albedo.r * reflectance), div(numerator, 1.0f - albedo.g * reflectance), div(numerator, 1.0f - albedo.b * reflectance) };
}

// Lambertian.cs:74-98
ECHO_DEVICE Sampled lambert_two_sided_sample(vec2 sample, vec3 outgoing, vec3& incident)
{
	bool reflectSide = sample.x > 0.5f;
	sample = { sample1d(abs_bits(sample.x * 2.0f - 1.0f)), sample.y };

	incident = cosine_hemisphere(sample);
	float pdf = cosine_p(incident) * kTauR;
	bool flip = (outgoing.z > 0.0f) ^ reflectSide;

	if (flip) incident = negate_z(incident);
	return { make_rgb(kTauR), pdf };
}

ECHO_DEVICE vec3 specular_reflect(vec3 outgoing) { return { -outgoing.x, -outgoing.y, outgoing.z }; } // Specular.cs:30

// Specular.cs:20-28
template<bool REAL>
ECHO_DEVICE Sampled specular_reflection_sample(const Bsdf& b, vec3 outgoing, vec3& incident)
{
	incident = specular_reflect(outgoing);
	float cosO = cosine_p(outgoing);
	float cosI = cosine_p(incident);
	return { lobe_fresnel<REAL>(b, cosO) / abs_bits(cosI), 1.0f };
}

// Specular.cs:44-59
ECHO_DEVICE Sampled specular_transmission_sample(float etaAbove, float etaBelow, vec3 outgoing, vec3& incident)
{
	FresnelPacket packet = fresnel_incomplete(etaAbove, etaBelow, cosine_p(outgoing));
	fresnel_complete(packet);

	if (fresnel_total_internal(packet))
	{
		incident = { 0.0f, 0.0f, 0.0f };
		return impossible();
	}

	float evaluated = 1.0f - fresnel_value(packet);
	incident = fresnel_refract(packet, outgoing, vec3{ 0.0f, 0.0f, 1.0f });
	evaluated = div(evaluated, abs_bits(cosine_p(incident)));
	return { make_rgb(evaluated), 1.0f };
}

// Specular.cs:73-90
ECHO_DEVICE Sampled specular_fresnel_sample(const Bsdf& b, vec2 sample, vec3 outgoing, vec3& incident)
{
	FresnelPacket packet = fresnel_incomplete(b.etaAbove, b.etaBelow, cosine_p(outgoing));
	fresnel_complete(packet);
	float evaluated = fresnel_value(packet);

	if (sample.x < evaluated) incident = specular_reflect(outgoing);
	else
	{
		evaluated = 1.0f - evaluated;
		incident = fresnel_refract(packet, outgoing, vec3{ 0.0f, 0.0f, 1.0f });
	}

	return { make_rgb(evaluated) / abs_bits(cosine_p(incident)), evaluated };
}

// =====================================================================================================================
// the BSDF container (Scattering/BSDF.cs)
// =====================================================================================================================

ECHO_DEVICE int bsdf_lobe_count(int kind) { return kind == BSDF_EMPTY ? 0 : ((kind == BSDF_DIELECTRIC_GLOSSY || kind == BSDF_COATED_DIFFUSE) ? 2 : 1); }

// FunctionType of lobe `index`
ECHO_DEVICE int bsdf_lobe_type(int kind, int index)
{
	switch (kind)
	{
		case BSDF_LAMBERT_REFLECTION:
		case BSDF_OREN_NAYAR: return FT_REFLECTIVE | FT_DIFFUSE;
		case BSDF_LAMBERT_TWO_SIDED: return FT_DIFFUSE | FT_REFLECTIVE | FT_TRANSMISSIVE;
		case BSDF_DIELECTRIC_GLOSSY: return index == 0 ? (FT_GLOSSY | FT_REFLECTIVE) : (FT_GLOSSY | FT_TRANSMISSIVE);
		case BSDF_DIELECTRIC_SPECULAR: return FT_SPECULAR | FT_REFLECTIVE | FT_TRANSMISSIVE;
		case BSDF_CONDUCTOR_GLOSSY: return FT_GLOSSY | FT_REFLECTIVE;
		case BSDF_CONDUCTOR_SPECULAR: return FT_SPECULAR | FT_REFLECTIVE;
		case BSDF_INVISIBLE: return FT_SPECULAR | FT_TRANSMISSIVE;
		case BSDF_COATED_DIFFUSE: return index == 0 ? (FT_REFLECTIVE | FT_DIFFUSE) : (FT_GLOSSY | FT_REFLECTIVE);
		default: return 0;
	}
}

ECHO_DEVICE int bsdf_reflect_type(const Bsdf& b, vec3 outgoingWorld, vec3 incidentWorld) // BSDF.cs:210-217
{
	float dot0 = dot(outgoingWorld, b.geometricNormal);
	float dot1 = dot(incidentWorld, b.geometricNormal);
	return dot0 * dot1 > 0.0f ? FT_REFLECTIVE : FT_TRANSMISSIVE;
}

#define ECHO_HAS(kindValue) ((KINDS & kind_bit(kindValue)) != 0u && b.kind == (kindValue))

template<uint32_t KINDS>
ECHO_DEVICE rgb lobe_evaluate(const Bsdf& b, int index, vec3 outgoing, vec3 incident)
{
	if (ECHO_HAS(BSDF_LAMBERT_REFLECTION)) return lambert_reflection_evaluate<false>(b, outgoing, incident);
	if (ECHO_HAS(BSDF_OREN_NAYAR)) return lambert_reflection_evaluate<true>(b, outgoing, incident);
	if (ECHO_HAS(BSDF_LAMBERT_TWO_SIDED)) return make_rgb(kTauR); // Lambertian.cs:78
	if (ECHO_HAS(BSDF_DIELECTRIC_GLOSSY)) return index == 0 ? glossy_reflection_evaluate<true>(b, outgoing, incident) : glossy_transmission_evaluate(b, outgoing, incident);
	if (ECHO_HAS(BSDF_CONDUCTOR_GLOSSY)) return glossy_reflection_evaluate<false>(b, outgoing, incident);
	if (ECHO_HAS(BSDF_COATED_DIFFUSE)) return index == 0 ? coated_lambert_evaluate(b, outgoing, incident) : glossy_reflection_evaluate<true>(b, outgoing, incident);
	return make_rgb(0.0f); // specular lobes evaluate to black (Specular.cs:17,41,70)
}

template<uint32_t KINDS>
ECHO_DEVICE float lobe_pdf(const Bsdf& b, int index, vec3 outgoing, vec3 incident)
{
	if (ECHO_HAS(BSDF_LAMBERT_REFLECTION) || ECHO_HAS(BSDF_OREN_NAYAR)) return lambert_reflection_pdf(outgoing, incident);
	if (ECHO_HAS(BSDF_LAMBERT_TWO_SIDED)) return abs_bits(cosine_p(incident)) * kTauR; // Lambertian.cs:80
	if (ECHO_HAS(BSDF_DIELECTRIC_GLOSSY)) return index == 0 ? glossy_reflection_pdf<true>(b, outgoing, incident) : glossy_transmission_pdf(b, outgoing, incident);
	if (ECHO_HAS(BSDF_CONDUCTOR_GLOSSY)) return glossy_reflection_pdf<false>(b, outgoing, incident);
	if (ECHO_HAS(BSDF_COATED_DIFFUSE)) return index == 0 ? lambert_reflection_pdf(outgoing, incident) : glossy_reflection_pdf<true>(b, outgoing, incident);
	return 0.0f;
}

template<uint32_t KINDS>
ECHO_DEVICE Sampled lobe_sample(const Bsdf& b, int index, vec2 sample, vec3 outgoing, vec3& incident)
{
	if (ECHO_HAS(BSDF_LAMBERT_REFLECTION)) return lambert_reflection_sample<false>(b, sample, outgoing, incident);
	if (ECHO_HAS(BSDF_OREN_NAYAR)) return lambert_reflection_sample<true>(b, sample, outgoing, incident);
	if (ECHO_HAS(BSDF_LAMBERT_TWO_SIDED)) return lambert_two_sided_sample(sample, outgoing, incident);
	if (ECHO_HAS(BSDF_DIELECTRIC_GLOSSY)) return index == 0 ? glossy_reflection_sample<true>(b, sample, outgoing, incident) : glossy_transmission_sample(b, sample, outgoing, incident);
	if (ECHO_HAS(BSDF_DIELECTRIC_SPECULAR)) return specular_fresnel_sample(b, sample, outgoing, incident);
	if (ECHO_HAS(BSDF_CONDUCTOR_GLOSSY)) return glossy_reflection_sample<false>(b, sample, outgoing, incident);
	if (ECHO_HAS(BSDF_CONDUCTOR_SPECULAR)) return specular_reflection_sample<false>(b, outgoing, incident);
	if (ECHO_HAS(BSDF_INVISIBLE)) return specular_transmission_sample(1.0f, 1.0f, outgoing, incident);
	if (ECHO_HAS(BSDF_COATED_DIFFUSE)) return index == 0 ? coated_lambert_sample(b, sample, outgoing, incident) : glossy_reflection_sample<true>(b, sample, outgoing, incident);
	incident = { 0.0f, 0.0f, 0.0f };
	return impossible();
}

// BSDF.Evaluate, BSDF.cs:97-116 (type = All)
template<uint32_t KINDS>
ECHO_DEVICE rgb bsdf_evaluate(const Bsdf& b, vec3 outgoingWorld, vec3 incidentWorld)
{
	vec3 outgoing = apply_inverse(b.transform, outgoingWorld);
	vec3 incident = apply_inverse(b.transform, incidentWorld);
	int reflectType = bsdf_reflect_type(b, outgoingWorld, incidentWorld);

	rgb total = make_rgb(0.0f);
	int count = bsdf_lobe_count(b.kind);

	for (int i = 0; i < count; i++)
	{
		if ((bsdf_lobe_type(b.kind, i) & reflectType) == 0) continue;
		total = total + lobe_evaluate<KINDS>(b, i, outgoing, incident);
	}

	return b.tint * total;
}

// BSDF.ProbabilityDensity, BSDF.cs:122-143
template<uint32_t KINDS>
ECHO_DEVICE float bsdf_pdf(const Bsdf& b, vec3 outgoingWorld, vec3 incidentWorld)
{
	vec3 outgoing = apply_inverse(b.transform, outgoingWorld);
	vec3 incident = apply_inverse(b.transform, incidentWorld);

	int count = bsdf_lobe_count(b.kind);
	float pdf = 0.0f;

	for (int i = 0; i < count; i++) pdf += lobe_pdf<KINDS>(b, i, outgoing, incident);

	return count < 2 ? pdf : div(pdf, (float)count);
}

// BSDF.Sample, BSDF.cs:150-208. selectedType = FunctionType of the sampled lobe (0 when no lobe exists).
template<uint32_t KINDS>
ECHO_DEVICE Sampled bsdf_sample(const Bsdf& b, vec3 outgoingWorld, vec2 sample, vec3& incidentWorld, int& selectedType)
{
	incidentWorld = { 0.0f, 0.0f, 0.0f };
	selectedType = 0;

	int matched = bsdf_lobe_count(b.kind); // FindFunction with FunctionType.All: every lobe matches (:219-232)
	if (matched == 0) return impossible();

	int index;
	sample.x = sample_range(sample.x, matched, index);
	selectedType = bsdf_lobe_type(b.kind, index);

	vec3 outgoing = apply_inverse(b.transform, outgoingWorld);
	vec3 incident;
	Sampled sampled = lobe_sample<KINDS>(b, index, sample, outgoing, incident);

	if (!positive(sampled.pdf) || is_zero(sampled.content)) return impossible();

	incidentWorld = apply_forward(b.transform, incident);
	int reflectType = bsdf_reflect_type(b, outgoingWorld, incidentWorld);

	if (matched == 1 || (selectedType & FT_SPECULAR))
	{
		bool wrongSide = (selectedType & reflectType) == 0;
		if (wrongSide) return impossible();
		return { b.tint * sampled.content, div(sampled.pdf, (float)matched) };
	}

	rgb total = sampled.content;
	float pdf = sampled.pdf;

	for (int i = 0; i < matched; i++)
	{
		if (i == index) continue;
		if ((bsdf_lobe_type(b.kind, i) & reflectType) == 0) continue;

		total = total + lobe_evaluate<KINDS>(b, i, outgoing, incident);
		pdf += lobe_pdf<KINDS>(b, i, outgoing, incident);
	}

	return { b.tint * total, div(pdf, (float)matched) };
}

#undef ECHO_HAS

// =====================================================================================================================
// Material.Scatter (Materials/Material.cs:63-75 and the concrete materials): 64-byte record -> Bsdf
// =====================================================================================================================

struct MaterialRecord
{
	uint32_t type, flags;
	float albedo[4];
	float roughness[2];
	float ior;
	float paramA[3], paramB[3];
	uint32_t base;
};

ECHO_DEVICE MaterialRecord load_material(const DeviceScene& scene, uint32_t index)
{
	ECHO_CHECK(scene, index < scene.materialCount, CHECK_MATERIAL);
	const float4* p = scene.materials + (size_t)index * 4;
	float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3);
	MaterialRecord m;
	m.type = __float_as_uint(a.x); m.flags = __float_as_uint(a.y);
	m.albedo[0] = a.z; m.albedo[1] = a.w; m.albedo[2] = b.x; m.albedo[3] = b.y;
	m.roughness[0] = b.z; m.roughness[1] = b.w;
	m.ior = c.x;
	m.paramA[0] = c.y; m.paramA[1] = c.z; m.paramA[2] = c.w;
	m.paramB[0] = d.x; m.paramB[1] = d.y; m.paramB[2] = d.z;
	m.base = __float_as_uint(d.w);
	return m;
}

ECHO_DEVICE rgb material_emission(const MaterialRecord& m) { return { m.albedo[0], m.albedo[1], m.albedo[2] }; }
ECHO_DEVICE float emissive_power(const MaterialRecord& m) { return luminance(material_emission(m)) * kPi; } // Emissive.cs:52-53

// Gulbrandsen 2014 artist-friendly -> physical, one lane (Conductor.cs:84-108; Float4.Lerp on an FMA host, Float4.cs:396-406)
ECHO_DEVICE void conductor_artistic(float mainColor, float edge, float& eta, float& k)
{
	mainColor = min_sse(mainColor, kOneMinusEpsilon);
	float root = __fsqrt_rn(mainColor);

	float low = div(1.0f + root, 1.0f - root);
	float high = div(1.0f - mainColor, 1.0f + mainColor);
	eta = __fmaf_rn(edge, high, __fmaf_rn(-edge, low, low));

	float value = mainColor * ((eta + 1.0f) * (eta + 1.0f)) - (eta - 1.0f) * (eta - 1.0f);
	k = __fsqrt_rn(max_sse(div(value, 1.0f - mainColor), 0.0f));
}

// =====================================================================================================================
// Image textures: TextureGrid.this[Float2] -> IFilter.Evaluate (Textures/Grids/IFilter.cs:17-68) with IWrapper (IWrapper.cs:18-96)
// =====================================================================================================================

ECHO_DEVICE int repeat_int(int value, int length) // Scalars.Repeat(int, int), Scalars.cs:205-210
{
	if ((0 <= value) & (value < length)) return value;
	int mod = value % length;
	return mod < 0 ? mod + length : mod;
}

ECHO_DEVICE void texture_wrap(uint32_t wrapper, int width, int height, int& x, int& y)
{
	if (wrapper == ECHO_WRAPPER_CLAMP)
	{
		x = x < 0 ? 0 : (x > width - 1 ? width - 1 : x);
		y = y < 0 ? 0 : (y > height - 1 ? height - 1 : y);
	}
	else if (wrapper == ECHO_WRAPPER_REPEAT)
	{
		x = repeat_int(x, width);
		y = repeat_int(y, height);
	}
	else
	{
		x = repeat_int(x, width * 2);
		y = repeat_int(y, height * 2);
		x = x < width * 2 - 1 - x ? x : width * 2 - 1 - x;
		y = y < height * 2 - 1 - y ? y : height * 2 - 1 - y;
	}
}

ECHO_DEVICE float lerp_fma(float first, float second, float value) { return __fmaf_rn(value, second, __fmaf_rn(-value, first, first)); } // Float4.Lerp, Float4.cs:396-406

ECHO_DEVICE float4 texture_sample(const DeviceScene& scene, uint32_t index, vec2 uv)
{
	ECHO_CHECK(scene, index < scene.textureCount, CHECK_TEXTURE);
	uint4 header = __ldg(scene.textures + (size_t)index * 2); // width height texelOffset filter
	ECHO_CHECK(scene, (uint64_t)header.z + (uint64_t)header.x * header.y <= scene.texelCount, CHECK_TEXEL);
	uint32_t wrapper = __ldg(scene.textures + (size_t)index * 2 + 1).x;
	int width = (int)header.x, height = (int)header.y;
	const float4* texels = scene.texels + header.z;
	float sizeX = (float)width, sizeY = (float)height;

	if (header.w == ECHO_FILTER_POINT)
	{
		int x = (int)floor((double)(uv.x * sizeX)), y = (int)floor((double)(uv.y * sizeY)); // ToPosition: (uv * size).Floored
		texture_wrap(wrapper, width, height, x, y);
		ECHO_CHECK(scene, x >= 0 && x < width && y >= 0 && y < height, CHECK_TEXEL);
		return __ldg(texels + (size_t)y * width + x);
	}

	float scaledX = uv.x * sizeX, scaledY = uv.y * sizeY;
	int roundX = __float2int_rn(scaledX), roundY = __float2int_rn(scaledY); // cvtps2dq: round to nearest even
	int x0 = roundX - 1, x1 = roundX, y0 = roundY - 1, y1 = roundY;
	int minX = x0, minY = y0;

	int ax = x0, ay = y0, bx = x1, by = y0, cx = x0, cy = y1, dx = x1, dy = y1;
	texture_wrap(wrapper, width, height, ax, ay);
	texture_wrap(wrapper, width, height, bx, by);
	texture_wrap(wrapper, width, height, cx, cy);
	texture_wrap(wrapper, width, height, dx, dy);

	ECHO_CHECK(scene, ax >= 0 && ax < width && bx >= 0 && bx < width && cx >= 0 && cx < width && dx >= 0 && dx < width, CHECK_TEXEL);
	ECHO_CHECK(scene, ay >= 0 && ay < height && by >= 0 && by < height && cy >= 0 && cy < height && dy >= 0 && dy < height, CHECK_TEXEL);
	float4 y0x0 = __ldg(texels + (size_t)ay * width + ax), y0x1 = __ldg(texels + (size_t)by * width + bx);
	float4 y1x0 = __ldg(texels + (size_t)cy * width + cx), y1x1 = __ldg(texels + (size_t)dy * width + dx);

	float timeX = scaledX - 0.5f - (float)minX;
	float timeY = scaledY - 0.5f - (float)minY;

	return make_float4(lerp_fma(lerp_fma(y0x0.x, y0x1.x, timeX), lerp_fma(y1x0.x, y1x1.x, timeX), timeY),
	                   lerp_fma(lerp_fma(y0x0.y, y0x1.y, timeX), lerp_fma(y1x0.y, y1x1.y, timeX), timeY),
	                   lerp_fma(lerp_fma(y0x0.z, y0x1.z, timeX), lerp_fma(y1x0.z, y1x1.z, timeX), timeY),
	                   lerp_fma(lerp_fma(y0x0.w, y0x1.w, timeX), lerp_fma(y1x0.w, y1x1.w, timeX), timeY));
}

// every textured slot of material `index` sampled at the contact's texture coordinate (Material.SampleAlbedo / Material.Sample)
ECHO_DEVICE void resolve_material_textures(const DeviceScene& scene, uint32_t index, vec2 texcoord, MaterialRecord& m)
{
	ECHO_CHECK(scene, index < scene.materialCount, CHECK_MATERIAL);
	uint4 slots = __ldg(scene.materialTextures + (size_t)index * 2);     // albedo normal roughness paramA
	uint32_t paramB = __ldg(scene.materialTextures + (size_t)index * 2 + 1).x;

	if (slots.x != ECHO_TEXTURE_NONE)
	{
		float4 value = texture_sample(scene, slots.x, texcoord);
		m.albedo[0] = value.x; m.albedo[1] = value.y; m.albedo[2] = value.z; m.albedo[3] = value.w;
	}

	if (slots.z != ECHO_TEXTURE_NONE)
	{
		float4 value = texture_sample(scene, slots.z, texcoord);
		m.roughness[0] = value.x; m.roughness[1] = value.y;
	}

	if (slots.w != ECHO_TEXTURE_NONE)
	{
		float4 value = texture_sample(scene, slots.w, texcoord);
		m.paramA[0] = value.x; m.paramA[1] = value.y; m.paramA[2] = value.z;
	}

	if (paramB != ECHO_TEXTURE_NONE)
	{
		float4 value = texture_sample(scene, paramB, texcoord);
		m.paramB[0] = value.x; m.paramB[1] = value.y; m.paramB[2] = value.z;
	}
}

// Material.ApplyNormalMapping (Material.cs:77-98), run by GeometryShade's constructor on the hit's own material
ECHO_DEVICE void apply_normal_mapping(const DeviceScene& scene, uint32_t index, vec2 texcoord, vec3& normal)
{
	uint32_t slot = __ldg(scene.materialTextures + (size_t)index * 2).y;
	float intensity = __uint_as_float(__ldg(scene.materialTextures + (size_t)index * 2 + 1).y);
	if (slot == ECHO_TEXTURE_NONE || almost_zero(intensity)) return; // zeroNormal, Material.cs:58

	float4 value = texture_sample(scene, slot, texcoord);
	float x = (max_sse(0.0f, min_sse(1.0f, value.x)) * -2.0f + 1.0f) * intensity; // Float4.Clamp: min.Max(max.Min(this)), Float4.cs:379
	float y = (max_sse(0.0f, min_sse(1.0f, value.y)) * 2.0f + -1.0f) * intensity;
	float z = (max_sse(0.0f, min_sse(1.0f, value.z)) * 2.0f + -2.0f) * intensity;

	frame transform = make_frame(normal);
	vec3 delta = apply_forward(transform, vec3{ x, y, z });
	normal = normalized(normal - delta);
}

// Resolves OneSided (OneSided.cs:50-58) and the alpha test (Material.cs:63-75), then fills `b`.
// `m` is the record of the hit's material; on return it still describes that top-level material (for emission tests).
// TEX: the scene has image textures; `topIndex` / `texcoord` then select and place the samples of the textured slots.
template<bool TEX = false>
ECHO_DEVICE void material_scatter(const DeviceScene& scene, const MaterialRecord& top, vec3 outgoing, vec3 geometricNormal, vec3 shadingNormal, Bsdf& b,
                                  uint32_t topIndex = 0u, vec2 texcoord = vec2{ 0.0f, 0.0f })
{
	b.transform = make_frame(shadingNormal); // BSDF.Reset, BSDF.cs:26-36
	b.geometricNormal = geometricNormal;
	b.alphaX = b.alphaY = 1.0f;
	b.etaAbove = b.etaBelow = 1.0f;
	b.eta2 = b.etaK2 = make_rgb(0.0f);
	b.orenA = b.orenB = 0.0f;
	b.coatedMultiplier = make_rgb(0.0f);

	MaterialRecord m = top;
	uint32_t materialIndex = topIndex;
	bool invisible = false;

	for (int level = 0; level < 4 && !invisible && m.type == ECHO_MATERIAL_ONESIDED; level++)
	{
		bool backface = (m.flags & ECHO_MATERIAL_FLAG_BACKFACE) != 0u;
		bool cull = positive(dot(outgoing, geometricNormal)) != backface;
		if (cull) invisible = true;
		else
		{
			materialIndex = m.base;
			m = load_material(scene, materialIndex);
		}
	}

	if (TEX && !invisible && scene.textureCount != 0u) resolve_material_textures(scene, materialIndex, texcoord, m);

	if (!invisible && m.type == ECHO_MATERIAL_EMISSIVE)
	{
		b.kind = BSDF_EMPTY;
		b.tint = make_rgb(0.0f);
		return;
	}

	if (m.type == ECHO_MATERIAL_INVISIBLE || m.type == ECHO_MATERIAL_ONESIDED) invisible = true; // OneSided chains deeper than 4 are cut
	if (!invisible && m.albedo[3] < 0.5f) invisible = true;

	if (invisible)
	{
		b.kind = BSDF_INVISIBLE;
		b.tint = make_rgb(1.0f);
		return;
	}

	b.tint = { m.albedo[0], m.albedo[1], m.albedo[2] };

	if (m.type == ECHO_MATERIAL_DIFFUSE) // Diffuse.cs:33-47
	{
		if (m.flags & ECHO_MATERIAL_FLAG_TRANSMISSIVE) b.kind = BSDF_LAMBERT_TWO_SIDED;
		else
		{
			float roughness = clamp01(m.roughness[0]);

			if (almost_zero(roughness)) b.kind = BSDF_LAMBERT_REFLECTION;
			else
			{
				b.kind = BSDF_OREN_NAYAR;
				b.orenA = rcp(fma_f(kPi / 2.0f - 2.0f / 3.0f, roughness, kPi)); // Lambertian.cs:107-108
				b.orenB = b.orenA * roughness;
			}
		}

		return;
	}

	bool specularX, specularY;
	b.alphaX = microfacet_alpha(m.roughness[0], specularX);
	b.alphaY = microfacet_alpha(m.roughness[1], specularY);
	bool glossy = !specularX || !specularY;

	if (m.type == ECHO_MATERIAL_COATED_DIFFUSE) // CoatedDiffuse.cs:37-55: always glossy, whatever the roughness
	{
		b.etaAbove = 1.0f;
		b.etaBelow = m.ior;
		coated_lambert_setup(b, b.tint, m.paramA[0]);
		b.kind = BSDF_COATED_DIFFUSE;
		return;
	}

	if (m.type == ECHO_MATERIAL_DIELECTRIC) // Dielectric.cs:29-47
	{
		b.etaAbove = 1.0f;
		b.etaBelow = m.ior;
		b.kind = glossy ? BSDF_DIELECTRIC_GLOSSY : BSDF_DIELECTRIC_SPECULAR;
		return;
	}

	// Conductor.cs:72-124
	rgb index, extinction;

	if (m.flags & ECHO_MATERIAL_FLAG_ARTISTIC)
	{
		conductor_artistic(m.paramA[0], m.paramB[0], index.r, extinction.r);
		conductor_artistic(m.paramA[1], m.paramB[1], index.g, extinction.g);
		conductor_artistic(m.paramA[2], m.paramB[2], index.b, extinction.b);
	}
	else
	{
		index = { m.paramA[0], m.paramA[1], m.paramA[2] };
		extinction = { m.paramB[0], m.paramB[1], m.paramB[2] };
	}

	complex_fresnel_setup(make_rgb(1.0f), max_epsilon(index), extinction, b.eta2, b.etaK2);
	b.kind = glossy ? BSDF_CONDUCTOR_GLOSSY : BSDF_CONDUCTOR_SPECULAR;
}

} // namespace echo
