// oracle/scattering.hpp — TEST INFRASTRUCTURE ONLY. CPU restatement of Echo.Core/Evaluation/Scattering and
// Evaluation/Materials: BxDFs, Trowbridge-Reitz microfacet with VNDF sampling, real/complex Fresnel, the BSDF
// container and the Material -> BSDF factories. Object-per-lobe with virtual calls, like the reference.
// Citations are relative to /root/reference/src/Echo.Core/.
#pragma once
#include "aggregation.hpp"

namespace oracle
{

// Evaluation/Scattering/FunctionType.cs
enum FunctionType : int
{
	Reflective = 1, Transmissive = 2, Diffuse = 4, Glossy = 8, Specular = 16, All = 31
};

inline bool type_fits(int type, int other) { return (type & other) == type; }
inline bool type_any(int type, int other) { return (type & other) != 0; }

// Common/Mathematics/Primitives/Probable.cs
struct ProbableRGB
{
	RGB content = kBlack;
	float pdf = 0.0f;

	bool not_possible() const { return !positive(pdf); }
};

// Evaluation/Scattering/BxDF.cs:44-135
inline float cosine_p(Float3 d) { return d.z; }
inline float cosine_p2(Float3 d) { return d.z * d.z; }
inline float sine_p2(Float3 d) { return one_minus2(d.z); }

inline float cosine_t2(Float3 d)
{
	float sin2 = sine_p2(d);
	if (almost_zero(sin2)) return 1.0f;
	return clamp01(d.x * d.x / sin2);
}

inline float sine_t2(Float3 d)
{
	float sin2 = sine_p2(d);
	if (almost_zero(sin2)) return 0.0f;
	return clamp01(d.y * d.y / sin2);
}

inline bool flat_or_same_hemisphere(Float3 a, Float3 b) { return !positive(-cosine_p(a) * cosine_p(b)); }
inline bool flat_or_opposite_hemisphere(Float3 a, Float3 b) { return !positive(cosine_p(a) * cosine_p(b)); }
inline Float3 negate_z(Float3 v) { return { v.x, v.y, -v.z }; } // Common/Utility.cs:57

struct BxDF
{
	explicit BxDF(int type) : type(type) {}
	virtual ~BxDF() = default;

	const int type;

	virtual RGB evaluate(Float3 outgoing, Float3 incident) const = 0;
	virtual float probability_density(Float3 outgoing, Float3 incident) const = 0;
	virtual ProbableRGB sample(Float2 sample, Float3 outgoing, Float3& incident) const = 0;
};

// ---- Evaluation/Scattering/Lambertian.cs:15-41 ----
struct LambertianReflection : BxDF
{
	LambertianReflection() : BxDF(Reflective | Diffuse) {}

	RGB evaluate(Float3 outgoing, Float3 incident) const override
	{
		if (flat_or_opposite_hemisphere(outgoing, incident)) return kBlack;
		return { kPiR, kPiR, kPiR };
	}

	float probability_density(Float3 outgoing, Float3 incident) const final
	{
		if (flat_or_opposite_hemisphere(outgoing, incident)) return 0.0f;
		return fabs_bits(cosine_p(incident)) * kPiR;
	}

	ProbableRGB sample(Float2 sample, Float3 outgoing, Float3& incident) const final
	{
		incident = cosine_hemisphere(sample);
		float pdf = cosine_p(incident) * kPiR;

		if (outgoing.z < 0.0f) incident = negate_z(incident);
		return { evaluate(outgoing, incident), pdf };
	}
};

// ---- Lambertian.cs:74-98 (two-sided; Diffuse.Transmissive) ----
struct Lambertian : BxDF
{
	Lambertian() : BxDF(Diffuse | Reflective | Transmissive) {}

	RGB evaluate(Float3, Float3) const override { return { kTauR, kTauR, kTauR }; }
	float probability_density(Float3, Float3 incident) const override { return fabs_bits(cosine_p(incident)) * kTauR; }

	ProbableRGB sample(Float2 sample, Float3 outgoing, Float3& incident) const override
	{
		bool reflect = sample.x > 0.5f;
		sample = { sample1d(fabs_bits(sample.x * 2.0f - 1.0f)), sample.y };

		incident = cosine_hemisphere(sample);
		float pdf = cosine_p(incident) * kTauR;
		bool flip = (outgoing.z > 0.0f) ^ reflect;

		if (flip) incident = negate_z(incident);
		return { { kTauR, kTauR, kTauR }, pdf };
	}
};

// ---- Lambertian.cs:101-128 ----
struct OrenNayar : LambertianReflection
{
	void reset(float roughness)
	{
		a = 1.0f / fma_f(kPi / 2.0f - 2.0f / 3.0f, roughness, kPi);
		b = a * roughness;
	}

	float a = 0.0f, b = 0.0f;

	RGB evaluate(Float3 outgoing, Float3 incident) const override
	{
		if (flat_or_opposite_hemisphere(outgoing, incident)) return kBlack;

		float cosO = fabs_bits(cosine_p(outgoing));
		float cosI = fabs_bits(cosine_p(incident));

		float s = dot(outgoing, incident) - cosO * cosI;
		if (positive(s)) s /= sse_max(cosO, cosI);
		float value = a + b * s;
		return { value, value, value };
	}
};

// ---- Evaluation/Scattering/Fresnel.cs:13-150 ----
struct RealFresnel
{
	float etaAbove = 1.0f, etaBelow = 1.0f;

	struct Packet
	{
		float etaOutgoing, etaIncident, cosOutgoing, cosIncident;

		Packet(float etaO, float etaI, float cosO, float cosI) // Fresnel.cs:50-61
			: etaOutgoing(etaO), etaIncident(etaI), cosOutgoing(clamp11(cosO)), cosIncident(clamp11(cosI)) {}

		Packet complete() const { return Packet(etaOutgoing, etaIncident, cosOutgoing, calculate_cosine_incident()); }
		bool total_internal_reflection() const { return almost_zero(cosIncident); }

		float value() const // Fresnel.cs:86-107
		{
			if (total_internal_reflection()) return 1.0f;

			float cosO = fabs_bits(cosOutgoing);
			float cosI = fabs_bits(cosIncident);

			float para0 = etaIncident * cosO;
			float para1 = etaOutgoing * cosI;
			float perp0 = etaOutgoing * cosO;
			float perp1 = etaIncident * cosI;

			float para = (para0 - para1) / (para0 + para1);
			float perp = (perp0 - perp1) / (perp0 + perp1);
			return (para * para + perp * perp) / 2.0f;
		}

		Float3 refract(Float3 outgoing, Float3 normal) const // Fresnel.cs:111-119
		{
			float eta = etaOutgoing / etaIncident;
			return normalized(normal * (eta * cosOutgoing + cosIncident) - eta * outgoing);
		}

		float calculate_cosine_incident() const // Fresnel.cs:121-130
		{
			float eta = etaOutgoing / etaIncident;
			float sinO2 = one_minus2(cosOutgoing);
			float sinI2 = eta * eta * sinO2;

			if (sinI2 >= 1.0f) return 0.0f;
			float result = sqrt0(1.0f - sinI2);
			return cosOutgoing > 0.0f ? -result : result;
		}
	};

	Packet create_incomplete(float cosOutgoing) const // Fresnel.cs:37-40; NaN cosIncident marks "incomplete"
	{
		return cosOutgoing > 0.0f ? Packet(etaAbove, etaBelow, cosOutgoing, NAN) : Packet(etaBelow, etaAbove, cosOutgoing, NAN);
	}

	float evaluate_scalar(float cosO) const { return create_incomplete(cosO).complete().value(); } // Fresnel.cs:28-35

	RGB evaluate(float cosO) const
	{
		float v = evaluate_scalar(cosO);
		return { v, v, v };
	}
};

// ---- Fresnel.cs:152-197 ----
struct ComplexFresnel
{
	RGB eta2 = kWhite, etaK2 = kBlack;

	ComplexFresnel() = default;

	ComplexFresnel(RGB etaAbove, RGB etaBelow, RGB extinction)
	{
		RGB etaAboveR = { 1.0f / etaAbove.r, 1.0f / etaAbove.g, 1.0f / etaAbove.b };

		eta2 = etaBelow * etaAboveR;
		etaK2 = extinction * etaAboveR;

		eta2 = eta2 * eta2;
		etaK2 = etaK2 * etaK2;
	}

	static float lane(float eta2, float etaK2, float cosO, float cosO2, float sinO2)
	{
		float term = eta2 - etaK2 - sinO2;
		float a2b2 = sqrt0(term * term + 4.0f * eta2 * etaK2);

		float para0 = a2b2 + cosO2;
		float para1 = cosO * kRoot2 * sqrt0(a2b2 + term);

		float perp0 = cosO2 * a2b2 + sinO2 * sinO2;
		float perp1 = para1 * sinO2;

		float para = (para0 - para1) / (para0 + para1);
		float perp = (perp0 - perp1) / (perp0 + perp1);

		return (para * perp + para) / 2.0f;
	}

	RGB evaluate(float cosO) const
	{
		cosO = clamp01(fabs_bits(cosO));

		float cosO2 = cosO * cosO;
		float sinO2 = 1.0f - cosO2;

		return { lane(eta2.r, etaK2.r, cosO, cosO2, sinO2), lane(eta2.g, etaK2.g, cosO, cosO2, sinO2), lane(eta2.b, etaK2.b, cosO, cosO2, sinO2) };
	}
};

// ---- Lambertian.cs:131-161 (CoatedDiffuse) ----
struct CoatedLambertianReflection : LambertianReflection
{
	void reset(RGB albedo, RealFresnel newFresnel, float reflectance)
	{
		fresnel = newFresnel;

		float eta = newFresnel.etaAbove / newFresnel.etaBelow;
		RGB denominator = { 1.0f - albedo.r * reflectance, 1.0f - albedo.g * reflectance, 1.0f - albedo.b * reflectance }; // RGB128.White - albedo * reflectance
		float numerator = eta * eta * kPiR;
		multiplier = { numerator / denominator.r, numerator / denominator.g, numerator / denominator.b };
	}

	RealFresnel fresnel;
	RGB multiplier = kBlack;

	RGB evaluate(Float3 outgoing, Float3 incident) const override
	{
		if (flat_or_opposite_hemisphere(outgoing, incident)) return kBlack;

		float evaluatedOutgoing = fresnel.evaluate_scalar(fabs_bits(cosine_p(outgoing)));
		float evaluatedIncident = fresnel.evaluate_scalar(fabs_bits(cosine_p(incident)));

		return multiplier * (1.0f - evaluatedOutgoing) * (1.0f - evaluatedIncident);
	}
};

// ---- Evaluation/Scattering/IMicrofacet.cs:43-51 ----
inline float microfacet_alpha(float roughness, bool& specular)
{
	roughness = clamp01(roughness * 0.75f);

	const float Threshold = 0.0001f;
	float alpha = roughness * roughness;
	specular = alpha < Threshold;
	return specular ? Threshold : alpha;
}

// ---- IMicrofacet.cs:94-174 ----
struct TrowbridgeReitz
{
	float alphaX = 1.0f, alphaY = 1.0f;

	float projected_area(Float3 normal) const
	{
		float cos2 = cosine_p2(normal);
		if (!positive(cos2)) return 0.0f;

		float sum = cos2;

		if (positive(1.0f - cos2, 1E-5f))
		{
			float x = normal.x / alphaX, y = normal.y / alphaY;
			sum += x * x + y * y;
		}

		return 1.0f / (sum * sum * (alphaX * alphaY) * kPi);
	}

	float shadowing_ratio(Float3 direction) const
	{
		float cos2 = cosine_p2(direction);
		if (!positive(cos2)) return 0.0f;
		float tan2 = sine_p2(direction) / cos2;

		float thetaX = cosine_t2(direction), thetaY = sine_t2(direction);
		float alpha2Tan2 = (alphaX * alphaX * thetaX + alphaY * alphaY * thetaY) * tan2;
		return sqrt0(1.0f + alpha2Tan2) / 2.0f - 0.5f;
	}

	float visibility(Float3 direction) const { return 1.0f / (1.0f + shadowing_ratio(direction)); } // IMicrofacet.cs:69-70

	float visibility(Float3 outgoing, Float3 incident) const // IMicrofacet.cs:72-73
	{
		return 1.0f / (1.0f + shadowing_ratio(outgoing) + shadowing_ratio(incident));
	}

	float probability_density(Float3 outgoing, Float3 normal) const // IMicrofacet.cs:75-79
	{
		float fraction = projected_area(normal) * visibility(outgoing);
		return fraction * fabs_bits(dot(outgoing, normal) / cosine_p(outgoing));
	}

	Float3 sample(Float3 outgoing, Float2 sample) const // IMicrofacet.cs:137-173 (Heitz 2017 VNDF)
	{
		Float3 scaled = normalized(Float3{ outgoing.x * alphaX, outgoing.y * alphaY, outgoing.z });
		if (scaled.z < 0.0f) scaled = -scaled;

		float threshold = 1.0f / (1.0f + scaled.z);
		float radius = sqrt0(sample.x);
		float theta = sample.y < threshold ? sample.y / threshold : 1.0f + (sample.y - threshold) / (1.0f - threshold);

		float sin, cos;
		sincos_det(theta * -kPi, sin, cos);

		float pointX = radius * cos;
		float pointY = radius * sin;

		if (sample.y >= threshold) pointY *= scaled.z;
		float pointZ = sqrt0(1.0f - (pointX * pointX + pointY * pointY));

		OrthonormalTransform transform(scaled);
		Float3 transformed = transform.apply_forward(Float3{ pointX, pointY, pointZ });

		return normalized(Float3{ transformed.x * alphaX, transformed.y * alphaY, sse_max(transformed.z, kEpsilon) });
	}
};

// ---- Evaluation/Scattering/Glossy.cs:61-71 ----
inline Float3 glossy_find_normal(Float3 outgoing, Float3 incident)
{
	Float3 normal = outgoing + incident;
	float length2 = squared_magnitude(normal);

	if (!positive(length2)) return { 0.0f, 0.0f, 1.0f };

	normal = normal * sqrt_r0(length2);
	return normal.z < 0.0f ? -normal : normal;
}

// ---- Glossy.cs:10-59 ----
template<class TFresnel>
struct GlossyReflection : BxDF
{
	GlossyReflection() : BxDF(Glossy | Reflective) {}

	TrowbridgeReitz microfacet;
	TFresnel fresnel;

	RGB evaluate(Float3 outgoing, Float3 incident) const override
	{
		if (flat_or_opposite_hemisphere(outgoing, incident)) return kBlack;
		Float3 normal = glossy_find_normal(outgoing, incident);

		float ratio = microfacet.projected_area(normal) * microfacet.visibility(outgoing, incident) * 0.25f;
		RGB evaluated = fresnel.evaluate(dot(outgoing, normal)) / (cosine_p(outgoing) * cosine_p(incident));
		return evaluated * ratio;
	}

	float probability_density(Float3 outgoing, Float3 incident) const override
	{
		if (flat_or_opposite_hemisphere(outgoing, incident)) return 0.0f;
		Float3 normal = glossy_find_normal(outgoing, incident);
		return microfacet.probability_density(outgoing, normal) / fabs_bits(dot(outgoing, normal) * 4.0f);
	}

	ProbableRGB sample(Float2 sample, Float3 outgoing, Float3& incident) const override
	{
		Float3 normal = microfacet.sample(outgoing, sample);
		incident = reflect(outgoing, normal);

		if (flat_or_opposite_hemisphere(outgoing, incident)) return {};

		float ratio = microfacet.projected_area(normal) * microfacet.visibility(outgoing, incident) * 0.25f;
		RGB evaluated = fresnel.evaluate(dot(outgoing, normal)) / (cosine_p(outgoing) * cosine_p(incident));
		float pdf = microfacet.probability_density(outgoing, normal) / fabs_bits(dot(outgoing, normal) * 4.0f);

		return { evaluated * ratio, pdf };
	}
};

// ---- Glossy.cs:74-162 ----
struct GlossyTransmission : BxDF
{
	GlossyTransmission() : BxDF(Glossy | Transmissive) {}

	TrowbridgeReitz microfacet;
	RealFresnel fresnel;

	static Float3 find_normal(Float3 outgoing, Float3 incident, float etaR) { return glossy_find_normal(outgoing, incident * etaR); }

	RGB evaluate(Float3 outgoing, Float3 incident) const override
	{
		if (flat_or_same_hemisphere(outgoing, incident)) return kBlack;

		auto packet = fresnel.create_incomplete(cosine_p(outgoing));
		float etaR = packet.etaIncident / packet.etaOutgoing;
		Float3 normal = find_normal(outgoing, incident, etaR);

		float dotO = dot(outgoing, normal);
		float dotI = dot(incident, normal);
		if (positive(dotO * dotI)) return kBlack;

		float evaluated = 1.0f - fresnel.evaluate_scalar(dotO);
		if (!positive(evaluated)) return kBlack;

		float numerator = etaR * etaR * dotO * dotI;
		float denominator = fma_f(etaR, dotI, dotO);
		denominator *= denominator;

		if (!positive(denominator)) denominator = 1.0f;
		denominator *= cosine_p(outgoing) * cosine_p(incident);

		float ratio = microfacet.projected_area(normal) * microfacet.visibility(outgoing, incident);
		float value = evaluated * ratio * fabs_bits(numerator / denominator);
		return { value, value, value };
	}

	float probability_density(Float3 outgoing, Float3 incident) const override
	{
		if (flat_or_same_hemisphere(outgoing, incident)) return 0.0f;

		auto packet = fresnel.create_incomplete(cosine_p(outgoing));
		float etaR = packet.etaIncident / packet.etaOutgoing;
		Float3 normal = find_normal(outgoing, incident, etaR);

		float dotO = dot(outgoing, normal);
		float dotI = dot(incident, normal);

		if (positive(dotO * dotI)) return 0.0f;
		float numerator = fabs_bits(etaR * etaR * dotI);
		float denominator = fma_f(etaR, dotI, dotO);
		denominator *= denominator;

		if (!positive(denominator)) denominator = 1.0f;
		return microfacet.probability_density(outgoing, normal) * (numerator / denominator);
	}

	ProbableRGB sample(Float2 sample, Float3 outgoing, Float3& incident) const override
	{
		Float3 normal = microfacet.sample(outgoing, sample);
		float dotO = dot(outgoing, normal);

		auto packet = fresnel.create_incomplete(dotO).complete();

		if (packet.total_internal_reflection())
		{
			incident = { 0.0f, 0.0f, 0.0f };
			return {};
		}

		incident = packet.refract(outgoing, normal);
		float dotI = dot(incident, normal);

		if (flat_or_same_hemisphere(outgoing, incident) || positive(dotO * dotI)) return {};

		float etaR = packet.etaIncident / packet.etaOutgoing;
		float numerator = fabs_bits(etaR * etaR * dotI);
		float denominator = fma_f(etaR, dotI, dotO);
		denominator *= denominator;

		if (!positive(denominator)) denominator = 1.0f;

		float ratio = microfacet.projected_area(normal) * microfacet.visibility(outgoing, incident);
		float evaluated = numerator * dotO / (denominator * cosine_p(outgoing) * cosine_p(incident));
		float pdf = microfacet.probability_density(outgoing, normal) * (numerator / denominator);

		float transmitted = 1.0f - packet.value();
		RGB result = RGB{ transmitted, transmitted, transmitted } * fabs_bits(evaluated) * ratio;
		return { result, pdf };
	}
};

// ---- Evaluation/Scattering/Specular.cs:9-31 ----
inline Float3 specular_reflect(Float3 outgoing) { return { -outgoing.x, -outgoing.y, outgoing.z }; }

template<class TFresnel>
struct SpecularReflection : BxDF
{
	SpecularReflection() : BxDF(Specular | Reflective) {}

	TFresnel fresnel;

	RGB evaluate(Float3, Float3) const override { return kBlack; }
	float probability_density(Float3, Float3) const override { return 0.0f; }

	ProbableRGB sample(Float2, Float3 outgoing, Float3& incident) const override
	{
		incident = specular_reflect(outgoing);
		float cosO = cosine_p(outgoing);
		float cosI = cosine_p(incident);

		RGB evaluated = fresnel.evaluate(cosO);
		return { evaluated / fabs_bits(cosI), 1.0f };
	}
};

// ---- Specular.cs:33-60 ----
struct SpecularTransmission : BxDF
{
	SpecularTransmission() : BxDF(Specular | Transmissive) {}

	RealFresnel fresnel;

	RGB evaluate(Float3, Float3) const override { return kBlack; }
	float probability_density(Float3, Float3) const override { return 0.0f; }

	ProbableRGB sample(Float2, Float3 outgoing, Float3& incident) const override
	{
		auto packet = fresnel.create_incomplete(cosine_p(outgoing)).complete();

		if (packet.total_internal_reflection())
		{
			incident = { 0.0f, 0.0f, 0.0f };
			return {};
		}

		float evaluated = 1.0f - packet.value();
		incident = packet.refract(outgoing, Float3{ 0.0f, 0.0f, 1.0f });
		evaluated /= fabs_bits(cosine_p(incident));

		return { { evaluated, evaluated, evaluated }, 1.0f };
	}
};

// ---- Specular.cs:62-91 ----
struct SpecularFresnel : BxDF
{
	SpecularFresnel() : BxDF(Specular | Reflective | Transmissive) {}

	RealFresnel fresnel;

	RGB evaluate(Float3, Float3) const override { return kBlack; }
	float probability_density(Float3, Float3) const override { return 0.0f; }

	ProbableRGB sample(Float2 sample, Float3 outgoing, Float3& incident) const override
	{
		auto packet = fresnel.create_incomplete(cosine_p(outgoing)).complete();
		float evaluated = packet.value();

		if (sample.x < evaluated) incident = specular_reflect(outgoing);
		else
		{
			evaluated = 1.0f - evaluated;
			incident = packet.refract(outgoing, Float3{ 0.0f, 0.0f, 1.0f });
		}

		RGB value = RGB{ evaluated, evaluated, evaluated } / fabs_bits(cosine_p(incident));
		return { value, evaluated };
	}
};

// ---- Evaluation/Scattering/BSDF.cs ----
struct BSDF
{
	int count = 0;
	RGB tint = kBlack;
	OrthonormalTransform transform{ Float3{ 0.0f, 0.0f, 1.0f } };
	Float3 geometricNormal = { 0.0f, 0.0f, 1.0f };
	const BxDF* functions[2] = { nullptr, nullptr };

	// storage for the lobes the in-scope materials can create (the reference allocates them from an arena, Material.cs:117-122)
	LambertianReflection lambertianReflection;
	Lambertian lambertian;
	OrenNayar orenNayar;
	CoatedLambertianReflection coatedLambertianReflection;
	GlossyReflection<RealFresnel> glossyReflectionReal;
	GlossyReflection<ComplexFresnel> glossyReflectionComplex;
	GlossyTransmission glossyTransmission;
	SpecularReflection<ComplexFresnel> specularReflectionComplex;
	SpecularTransmission specularTransmission;
	SpecularFresnel specularFresnel;

	void reset(Float3 shadingNormal, Float3 normal, RGB newTint) // BSDF.cs:26-36
	{
		count = 0;
		tint = newTint;
		transform = OrthonormalTransform(shadingNormal);
		geometricNormal = normal;
	}

	void add(const BxDF* function) { functions[count++] = function; }

	int reflect_type(Float3 outgoingWorld, Float3 incidentWorld) const // BSDF.cs:210-217
	{
		float dot0 = dot(outgoingWorld, geometricNormal);
		float dot1 = dot(incidentWorld, geometricNormal);
		return dot0 * dot1 > 0.0f ? Reflective : Transmissive;
	}

	RGB evaluate(Float3 outgoingWorld, Float3 incidentWorld) const // BSDF.cs:97-116 (type = All)
	{
		Float3 outgoing = transform.apply_inverse(outgoingWorld);
		Float3 incident = transform.apply_inverse(incidentWorld);
		int reflect = reflect_type(outgoingWorld, incidentWorld);

		RGB total = kBlack;

		for (int i = 0; i < count; i++)
		{
			const BxDF* function = functions[i];
			if (!type_fits(function->type, All) || !type_any(function->type, reflect)) continue;
			total = total + function->evaluate(outgoing, incident);
		}

		return tint * total;
	}

	float probability_density(Float3 outgoingWorld, Float3 incidentWorld) const // BSDF.cs:122-143
	{
		Float3 outgoing = transform.apply_inverse(outgoingWorld);
		Float3 incident = transform.apply_inverse(incidentWorld);

		int matched = 0;
		float pdf = 0.0f;

		for (int i = 0; i < count; i++)
		{
			pdf += functions[i]->probability_density(outgoing, incident);
			++matched;
		}

		return matched < 2 ? pdf : pdf / (float)matched;
	}

	// BSDF.cs:150-208; `selected` receives the sampled lobe (nullptr when none matched)
	ProbableRGB sample(Float3 outgoingWorld, Float2 sample, Float3& incidentWorld, const BxDF*& selected) const
	{
		selected = nullptr;
		incidentWorld = { 0.0f, 0.0f, 0.0f };

		// FindFunction, BSDF.cs:219-255 (FunctionType.All always fits -> all stored functions match)
		if (count == 0) return {};
		int matched = count;
		int index;
		sample.x = sample_range(sample.x, count, index);

		selected = functions[index];

		Float3 outgoing = transform.apply_inverse(outgoingWorld);
		Float3 incident;
		ProbableRGB sampled = selected->sample(sample, outgoing, incident);

		if (sampled.not_possible() || is_zero(sampled.content)) return {};

		incidentWorld = transform.apply_forward(incident);
		int reflect = reflect_type(outgoingWorld, incidentWorld);

		if (matched == 1 || type_any(selected->type, Specular))
		{
			bool wrongSide = !type_any(selected->type, reflect);
			if (wrongSide) return {};
			return { tint * sampled.content, sampled.pdf / (float)matched };
		}

		RGB total = sampled.content;
		float pdf = sampled.pdf;

		for (int i = 0; i < count; i++)
		{
			const BxDF* function = functions[i];
			if (function == selected) continue;
			if (!type_any(function->type, reflect)) continue;

			total = total + function->evaluate(outgoing, incident);
			pdf += function->probability_density(outgoing, incident);
		}

		return { tint * total, pdf / (float)matched };
	}
};

// ---- Aggregation/Primitives/Contact.cs ----
struct Contact
{
	uint32_t token = ECHO_TOKEN_EMPTY;
	Layers layers; // the instance layers of the hit's TokenHierarchy
	Float3 outgoing = {};
	GeometryPoint point = {};
	Float3 shadeNormal = {};
	Float2 texcoord = { 0.0f, 0.0f }; // GeometryShade.Texcoord
	uint32_t material = 0;
	BSDF bsdf;

	float normal_dot(Float3 direction) const { return fabs_bits(dot(direction, shadeNormal)); } // Contact.cs:97
};

// ---- Evaluation/Materials: Material.Scatter + concrete materials ----
inline void scatter_invisible(Contact& contact) // Invisible.cs:22-26
{
	contact.bsdf.reset(contact.shadeNormal, contact.point.normal, kWhite);
	contact.bsdf.specularTransmission.fresnel = RealFresnel{ 1.0f, 1.0f };
	contact.bsdf.add(&contact.bsdf.specularTransmission);
}

// The material with every textured slot sampled at the contact's texture coordinate (Material.SampleAlbedo / Material.Sample,
// Material.cs:100-102): the rest of Scatter then reads constants, as it does for Pure textures.
inline EchoMaterial resolve_material(const Scene& scene, uint32_t materialIndex, Float2 texcoord)
{
	EchoMaterial material = scene.materials[materialIndex];
	if (scene.materialTextures.empty()) return material;

	EchoMaterialTextures slots = scene.material_textures(materialIndex);

	if (slots.albedo != ECHO_TEXTURE_NONE)
	{
		Scene::Rgba value = scene.texture_sample(slots.albedo, texcoord);
		for (int c = 0; c < 4; c++) material.albedo[c] = value.v[c];
	}

	if (slots.roughness != ECHO_TEXTURE_NONE)
	{
		Scene::Rgba value = scene.texture_sample(slots.roughness, texcoord);
		material.roughness[0] = value.v[0];
		material.roughness[1] = value.v[1];
	}

	if (slots.paramA != ECHO_TEXTURE_NONE)
	{
		Scene::Rgba value = scene.texture_sample(slots.paramA, texcoord);
		for (int c = 0; c < 3; c++) material.paramA[c] = value.v[c];
	}

	if (slots.paramB != ECHO_TEXTURE_NONE)
	{
		Scene::Rgba value = scene.texture_sample(slots.paramB, texcoord);
		for (int c = 0; c < 3; c++) material.paramB[c] = value.v[c];
	}

	return material;
}

// Material.ApplyNormalMapping (Material.cs:77-98), run by GeometryShade's constructor on the hit's own material
inline void apply_normal_mapping(const Scene& scene, uint32_t materialIndex, Float2 texcoord, Float3& normal)
{
	if (scene.materialTextures.empty()) return;
	EchoMaterialTextures slots = scene.material_textures(materialIndex);
	if (slots.normal == ECHO_TEXTURE_NONE || almost_zero(slots.normalIntensity)) return; // zeroNormal, Material.cs:58

	Scene::Rgba value = scene.texture_sample(slots.normal, texcoord);
	float local[3];
	const float scale[3] = { -2.0f, 2.0f, 2.0f }, offset[3] = { 1.0f, -1.0f, -2.0f };

	for (int c = 0; c < 3; c++)
	{
		float clamped = sse_max(0.0f, sse_min(1.0f, value.v[c])); // Float4.Clamp: min.Max(max.Min(this)), Float4.cs:379
		local[c] = (clamped * scale[c] + offset[c]) * slots.normalIntensity;
	}

	OrthonormalTransform transform(normal);
	Float3 delta = transform.apply_forward(Float3{ local[0], local[1], local[2] });
	normal = normalized(normal - delta);
}

inline void scatter_material(const Scene& scene, uint32_t materialIndex, Contact& contact)
{
	const EchoMaterial material = resolve_material(scene, materialIndex, contact.texcoord);
	BSDF& bsdf = contact.bsdf;

	switch (material.type)
	{
		case ECHO_MATERIAL_EMISSIVE: // Emissive.cs:56
		{
			bsdf.reset(contact.shadeNormal, contact.point.normal, kBlack);
			return;
		}
		case ECHO_MATERIAL_ONESIDED: // OneSided.cs:50-58
		{
			bool backface = (material.flags & ECHO_MATERIAL_FLAG_BACKFACE) != 0;
			bool cull = positive(dot(contact.outgoing, contact.point.normal)) != backface;
			if (!cull) scatter_material(scene, material.base, contact);
			else scatter_invisible(contact);
			return;
		}
		case ECHO_MATERIAL_INVISIBLE:
		{
			scatter_invisible(contact);
			return;
		}
		default: break;
	}

	// Material.cs:63-75
	if (material.albedo[3] < 0.5f)
	{
		scatter_invisible(contact);
		return;
	}

	RGB albedo = { material.albedo[0], material.albedo[1], material.albedo[2] };
	bsdf.reset(contact.shadeNormal, contact.point.normal, albedo); // Material.cs:117-122

	switch (material.type)
	{
		case ECHO_MATERIAL_DIFFUSE: // Diffuse.cs:33-47
		{
			if (!(material.flags & ECHO_MATERIAL_FLAG_TRANSMISSIVE))
			{
				float roughness = clamp01(material.roughness[0]);

				if (almost_zero(roughness)) bsdf.add(&bsdf.lambertianReflection);
				else
				{
					bsdf.orenNayar.reset(roughness);
					bsdf.add(&bsdf.orenNayar);
				}
			}
			else bsdf.add(&bsdf.lambertian);

			break;
		}
		case ECHO_MATERIAL_DIELECTRIC: // Dielectric.cs:29-47
		{
			RealFresnel fresnel{ 1.0f, material.ior };

			bool specularX, specularY;
			float alphaX = microfacet_alpha(material.roughness[0], specularX);
			float alphaY = microfacet_alpha(material.roughness[1], specularY);

			if (!specularX || !specularY)
			{
				TrowbridgeReitz microfacet{ alphaX, alphaY };

				bsdf.glossyReflectionReal.microfacet = microfacet;
				bsdf.glossyReflectionReal.fresnel = fresnel;
				bsdf.add(&bsdf.glossyReflectionReal);

				bsdf.glossyTransmission.microfacet = microfacet;
				bsdf.glossyTransmission.fresnel = fresnel;
				bsdf.add(&bsdf.glossyTransmission);
			}
			else
			{
				bsdf.specularFresnel.fresnel = fresnel;
				bsdf.add(&bsdf.specularFresnel);
			}

			break;
		}
		case ECHO_MATERIAL_COATED_DIFFUSE: // CoatedDiffuse.cs:37-55
		{
			RealFresnel fresnel{ 1.0f, material.ior };

			bool unused;
			float alphaX = microfacet_alpha(material.roughness[0], unused);
			float alphaY = microfacet_alpha(material.roughness[1], unused);

			bsdf.coatedLambertianReflection.reset(albedo, fresnel, material.paramA[0]);
			bsdf.add(&bsdf.coatedLambertianReflection);

			bsdf.glossyReflectionReal.microfacet = TrowbridgeReitz{ alphaX, alphaY };
			bsdf.glossyReflectionReal.fresnel = fresnel;
			bsdf.add(&bsdf.glossyReflectionReal);
			break;
		}
		case ECHO_MATERIAL_CONDUCTOR: // Conductor.cs:72-124
		{
			bool specularX, specularY;
			float alphaX = microfacet_alpha(material.roughness[0], specularX);
			float alphaY = microfacet_alpha(material.roughness[1], specularY);

			RGB index, extinction;

			if (material.flags & ECHO_MATERIAL_FLAG_ARTISTIC)
			{
				// Artist Friendly Metallic Fresnel [Gulbrandsen 2014], Conductor.cs:84-108, per RGB lane
				auto convert = [](float main, float edge, float& eta, float& k)
				{
					main = sse_min(main, kOneMinusEpsilon);
					float root = std::sqrt(main);

					// Float4.Lerp on an FMA host (Float4.cs:396-406): fma(value, other, fnma(value, this, this))
					float low = (1.0f + root) / (1.0f - root);
					float high = (1.0f - main) / (1.0f + main);
					eta = std::fma(edge, high, std::fma(-edge, low, low));

					float value = main * ((eta + 1.0f) * (eta + 1.0f)) - (eta - 1.0f) * (eta - 1.0f);
					k = std::sqrt(sse_max(value / (1.0f - main), 0.0f));
				};

				convert(material.paramA[0], material.paramB[0], index.r, extinction.r);
				convert(material.paramA[1], material.paramB[1], index.g, extinction.g);
				convert(material.paramA[2], material.paramB[2], index.b, extinction.b);
			}
			else
			{
				index = { material.paramA[0], material.paramA[1], material.paramA[2] };
				extinction = { material.paramB[0], material.paramB[1], material.paramB[2] };
			}

			ComplexFresnel fresnel(kWhite, max_epsilon(index), extinction);

			if (!specularX || !specularY)
			{
				bsdf.glossyReflectionComplex.microfacet = TrowbridgeReitz{ alphaX, alphaY };
				bsdf.glossyReflectionComplex.fresnel = fresnel;
				bsdf.add(&bsdf.glossyReflectionComplex);
			}
			else
			{
				bsdf.specularReflectionComplex.fresnel = fresnel;
				bsdf.add(&bsdf.specularReflectionComplex);
			}

			break;
		}
		default: break;
	}
}

} // namespace oracle
