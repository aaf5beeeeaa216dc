// oracle/evaluation.hpp — TEST INFRASTRUCTURE ONLY. CPU restatement of the per-sample integrator and its driver:
// PreparedScene.Interact / Sample / ProbabilityDensity / EvaluateInfinite, LightCollection, PathTracedEvaluator,
// PerspectiveCamera.SpawnRay, Accumulator and the EvaluationOperation pixel/epoch/sample loop.
// Citations are relative to /root/reference/src/Echo.Core/.
#pragma once
#include "scattering.hpp"

namespace oracle
{

// ---- Evaluation/Materials/Emissive.cs:52-53,64 ----
inline RGB material_emission(const EchoMaterial& m) { return { m.albedo[0], m.albedo[1], m.albedo[2] }; }
inline float emissive_power(const EchoMaterial& m) { return luminance(material_emission(m)) * kPi; }

inline RGB emissive_emit(const EchoMaterial& m, const GeometryPoint& point, Float3 outgoing)
{
	return dot(outgoing, point.normal) > 0.0f ? material_emission(m) : kBlack;
}

// ---- PreparedScene.Interact (PreparedScene.cs:95-105) + GeometryCollection.GetContactInfo (:200-232) + Material.Scatter ----
inline void interact(const Scene& scene, const TraceQuery& query, Contact& contact)
{
	Float3 infoNormal, infoShading;
	Float2 texcoord;
	uint32_t material;

	Scene::Layer layer = scene.find_layer(query.tokenLayers); // FindLayer, :97
	EchoPack view = scene.pack_view(layer.pack);

	if (token_type(query.token) == ECHO_TOKEN_TYPE_TRIANGLE)
	{
		const EchoTriangle& triangle = scene.triangles[view.triangleOffset + token_index(query.token)];
		material = triangle.material;
		infoNormal = triangle_normal(triangle);
		infoShading = triangle_shading_normal(triangle, query.uv);

		// PreparedTriangle.GetTexcoord, TriangleEntity.cs:188
		float w = 1.0f - query.uv.x - query.uv.y;
		texcoord = { w * triangle.texcoord0[0] + query.uv.x * triangle.texcoord1[0] + query.uv.y * triangle.texcoord2[0],
		             w * triangle.texcoord0[1] + query.uv.x * triangle.texcoord1[1] + query.uv.y * triangle.texcoord2[1] };
	}
	else
	{
		material = scene.spheres[view.sphereOffset + token_index(query.token)].material;
		infoNormal = infoShading = sphere_normal(query.uv);

		// PreparedSphere.GetTexcoord, SphereEntity.cs:236-245 (Atan2 / Asin pinned, oracle/math.hpp)
		float sinT, sinP, cosT;
		sphere_theta_phi(query.uv, sinT, sinP, cosT);
		texcoord = { fma_f(atan2_det(sinT, cosT), kTauR, 0.5f), fma_f(asin_det(clamp11(sinP)), kPiR, 0.5f) };
	}

	// without layers inverseTransform is the identity, but MultiplyDirection + Normalized still run (:100-101)
	contact.token = query.token;
	contact.layers = query.tokenLayers;
	contact.outgoing = -query.ray.direction;                                    // Contact.cs:39
	contact.point.position = query.position();                                  // Contact.cs:25
	contact.point.normal = normalized(multiply_direction(layer.inverse, infoNormal));
	contact.shadeNormal = normalized(multiply_direction(layer.inverse, infoShading));
	contact.texcoord = texcoord;
	contact.material = layer.materialOffset + material;                        // instance.swatch[info.material], :102
	apply_normal_mapping(scene, contact.material, texcoord, contact.shadeNormal); // GeometryShade's constructor, GeometryShade.cs:17

	scatter_material(scene, contact.material, contact);
}

// ---- PreparedTriangle.Sample (TriangleEntity.cs:166-174) / PreparedSphere.Sample (SphereEntity.cs:151-191) ----
inline bool geometry_sample(const Scene& scene, uint32_t token, Float3 origin, Float2 sample, GeometryPoint& point, float& pdf, uint32_t pack = 0)
{
	EchoPack view = scene.pack_view(pack);

	if (token_type(token) == ECHO_TOKEN_TYPE_TRIANGLE)
	{
		const EchoTriangle& triangle = scene.triangles[view.triangleOffset + token_index(token)];
		Float2 uv = uniform_triangle(sample);
		point.position = triangle_point(triangle, uv);
		point.normal = triangle_shading_normal(triangle, uv);
		pdf = geometry_point_pdf(point, origin, triangle_area(triangle));
		return true;
	}

	const EchoSphere& sphere = scene.spheres[view.sphereOffset + token_index(token)];
	Float3 position = f3(sphere.position);
	float radius = sphere.radius;

	auto get_point = [&](Float3 normal) { return GeometryPoint{ normal * radius + position, normal }; }; // SphereEntity.cs:227

	Float3 offset = origin - position;
	float radius2 = radius * radius;
	float length2 = squared_magnitude(offset);

	if (length2 < radius2)
	{
		point = get_point(uniform_sphere(sample));
		pdf = geometry_point_pdf(point, origin, sphere_area(sphere));
		return true;
	}

	float sinMaxT2 = radius2 / length2;
	float cosMaxT = sqrt0(1.0f - sinMaxT2);

	if (almost_zero(1.0f - cosMaxT))
	{
		pdf = 0.0f; // Probable.Impossible
		return false;
	}

	float cosT = fma_f(cosMaxT - 1.0f, sample.x, 1.0f);
	float sinT = identity(cosT);
	float phi = sample.y * kTau;

	float length = sqrt0(length2);
	float project = length * cosT - sqrt0(radius2 - length2 * sinT * sinT);
	float cosA = (length2 + radius2 - project * project) / (2.0f * length * radius);
	float sinA = identity(cosA);

	float sinP, cosP;
	sincos_det(phi, sinP, cosP);
	Float3 normal = normalized(Float3{ sinA * cosP, sinA * sinP, cosA });
	pdf = uniform_cone_pdf(cosMaxT);

	OrthonormalTransform transform(offset / length);
	point = get_point(transform.apply_forward(normal));
	return true;
}

// ---- PreparedTriangle.ProbabilityDensity (TriangleEntity.cs:177-185) / PreparedSphere.ProbabilityDensity (SphereEntity.cs:194-225) ----
inline float geometry_pdf(const Scene& scene, uint32_t token, Float3 origin, Float3 incident, uint32_t pack = 0)
{
	EchoPack view = scene.pack_view(pack);

	if (token_type(token) == ECHO_TOKEN_TYPE_TRIANGLE)
	{
		const EchoTriangle& triangle = scene.triangles[view.triangleOffset + token_index(token)];
		Float2 uv = { 0.0f, 0.0f };
		float distance = triangle_intersect(triangle, origin, incident, uv);
		if (distance == kInfinity) return 0.0f;
		return distance * distance / fabs_bits(dot(triangle_shading_normal(triangle, uv), incident) * triangle_area(triangle));
	}

	const EchoSphere& sphere = scene.spheres[view.sphereOffset + token_index(token)];
	float radius = sphere.radius;
	Float3 offset = origin - f3(sphere.position);
	float radius2 = radius * radius;
	float length2 = squared_magnitude(offset);

	if (length2 <= radius2)
	{
		float projected = dot(offset, incident);
		float extend2 = fma_f(projected, projected, radius2 - length2);

		float distance = sqrt0(extend2) - projected;
		Float3 normal = offset + incident * distance;

		float cosWeight = dot(incident, normal) * radius;
		if (almost_zero(cosWeight)) return 0.0f;

		return distance * distance / fabs_bits(cosWeight) * kUniformSpherePdf;
	}

	float sinMaxT2 = radius2 / length2;
	float cosMaxT = sqrt0(1.0f - sinMaxT2);
	return almost_zero(1.0f - cosMaxT) ? 0.0f : uniform_cone_pdf(cosMaxT);
}

// ---- Evaluation/Sampling/DiscreteDistribution1D.cs:58-166 over a stored cdf ----
struct Distribution1D
{
	const float* cdf;
	int count;

	void bounds(int index, float& lower, float& upper) const // GetBounds, :157-164
	{
		lower = index == 0 ? 0.0f : cdf[index - 1];
		upper = cdf[index];
	}

	int find_index(float sample) const // FindIndex / FixIndex / BinarySearch, :111-155,166-190
	{
		uint32_t head = 0u, tail = (uint32_t)count;
		int index = -1;

		while (head < tail)
		{
			uint32_t middle = (tail + head) >> 1;
			float current = cdf[middle];
			if (current == sample) { index = (int)middle; break; }
			if (current > sample) tail = middle;
			else head = middle + 1u;
		}

		if (index < 0) return (int)head;

		float lower, upper; // landed exactly on an anchor
		bounds(index, lower, upper);
		if (upper - lower > 0.0f) return index + 1;

		do
		{
			lower = upper;
			upper = cdf[++index];
		}
		while (lower == upper);

		return index;
	}

	float sample(float u, float& pdf) const // Sample, :58-71
	{
		int index = find_index(u);
		float lower, upper;
		bounds(index, lower, upper);

		float gap = upper - lower;
		float countR = 1.0f / (float)count;
		float shift = (u - lower) / gap + (float)index;
		pdf = gap * (float)count;
		return sample1d(shift * countR);
	}

	float probability_density(float result) const // :89-93
	{
		float lower, upper;
		bounds(sample_range(result, count), lower, upper);
		return (upper - lower) * (float)count;
	}
};

// ---- Textures/Directional/CylindricalTexture.cs:98-165 ----
constexpr float kCylindricalJacobian = kTauR / kPi; // :19

inline Float2 cylindrical_to_uv(Float3 direction) // ToUV, :142-151 (Atan2 / Acos pinned, oracle/math.hpp)
{
	return { fma_f(atan2_det(direction.x, direction.z), kTauR, 0.5f), fma_f(acos_det(clamp11(direction.y)), -kPiR, 1.0f) };
}

inline Float3 cylindrical_to_direction(Float2 uv) // ToDirection, :153-164 (only the unit tests call it; Sample inlines the same lines)
{
	float sinT, cosT, sinP, cosP;
	sincos_det(uv.x * kTau, sinT, cosT);
	sincos_det(uv.y * kPi, sinP, cosP);
	return normalized(Float3{ -sinP * sinT, -cosP, -sinP * cosT });
}

inline RGB environment_texel(const Scene& scene, const EchoInfiniteLight& light, Float2 uv)
{
	Scene::Rgba value = scene.texture_sample(light.texture, uv);
	return { value.v[0], value.v[1], value.v[2] };
}

inline float cylindrical_pdf(const Scene& scene, const EchoInfiniteLight& light, Float3 incident) // ProbabilityDensity, :100-110
{
	Float2 uv = cylindrical_to_uv(incident);
	float cosP = -incident.y;
	float sinP = identity(cosP);
	if (!positive(sinP)) return 0.0f;

	const EchoTexture& texture = scene.textures[light.texture];
	const float* values = scene.distributions.data() + light.distribution;
	int width = (int)texture.width, height = (int)texture.height;
	float x = sample1d(uv.x), y = sample1d(uv.y); // (Sample2D)uv

	Distribution1D vertical{ values, height };
	Distribution1D slice{ values + height + (size_t)sample_range(y, height) * width, width }; // DiscreteDistribution2D.ProbabilityDensity, :69-77
	float pdfX = slice.probability_density(x);
	float pdfY = vertical.probability_density(y);
	return pdfX * pdfY * kCylindricalJacobian / sinP;
}

inline ProbableRGB cylindrical_sample(const Scene& scene, const EchoInfiniteLight& light, Float2 sample, Float3& incident) // Sample, :112-140
{
	const EchoTexture& texture = scene.textures[light.texture];
	const float* values = scene.distributions.data() + light.distribution;
	int width = (int)texture.width, height = (int)texture.height;

	// DiscreteDistribution2D.Sample, DiscreteDistribution2D.cs:40-46
	float pdfY, pdfX;
	float y = Distribution1D{ values, height }.sample(sample.y, pdfY);
	float x = Distribution1D{ values + height + (size_t)sample_range(y, height) * width, width }.sample(sample.x, pdfX);
	float pdf = pdfX * pdfY;

	incident = { 0.0f, 0.0f, 0.0f };
	if (!positive(pdf)) return {};

	float theta = x * kTau;
	float phi = y * kPi;

	float sinT, cosT, sinP, cosP;
	sincos_det(theta, sinT, cosT);
	sincos_det(phi, sinP, cosP);
	if (!positive(sinP)) return {};

	incident = { -sinP * sinT, -cosP, -sinP * cosT };
	return { environment_texel(scene, light, Float2{ x, y }), pdf * kCylindricalJacobian / sinP };
}

inline Float3 rotate3x3(const float* m, Float3 v) // Float3x3 * Float3, Float3x3.cs:264-269
{
	return { m[0] * v.x + m[1] * v.y + m[2] * v.z, m[3] * v.x + m[4] * v.y + m[5] * v.z, m[6] * v.x + m[7] * v.y + m[8] * v.z };
}

// ---- Textures/Directional/Cubemap.cs:62-82 with (Direction)incident (Direction.cs:325-333, Float3.MaxIndex Float3.cs:130-138) ----
inline RGB cubemap_evaluate(const Scene& scene, const EchoInfiniteLight& light, Float3 incident)
{
	float ax = std::fabs(incident.x), ay = std::fabs(incident.y), az = std::fabs(incident.z);
	int axis = ax > ay ? (ax > az ? 0 : 2) : (ay > az ? 1 : 2);
	float part = axis == 0 ? incident.x : (axis == 1 ? incident.y : incident.z);
	bool negative = part < 0.0f;
	int index = axis * 2 + (negative ? 1 : 0);

	Float2 uv;
	switch (index)
	{
		case 0: uv = { -incident.z, incident.y }; break;
		case 1: uv = { incident.z, incident.y }; break;
		case 2: uv = { incident.x, -incident.z }; break;
		case 3: uv = { incident.x, incident.z }; break;
		case 4: uv = { incident.x, incident.y }; break;
		default: uv = { -incident.x, incident.y }; break;
	}

	float scale = 0.5f / (negative ? -part : part); // target.ExtractComponent(incident)
	uv = { uv.x * scale + 0.5f, uv.y * scale + 0.5f };
	Scene::Rgba value = scene.texture_sample(light.texture + (uint32_t)index, uv);
	return { value.v[0], value.v[1], value.v[2] };
}

// ---- Scenic/Lights: AmbientLight over a Pure texture (AmbientLight.cs:53-67 with IDirectionalTexture's defaults) and
//      DirectionalLight (DirectionalLight.cs:78-108) ----
inline RGB infinite_radiance(const EchoInfiniteLight& light) { return { light.radiance[0], light.radiance[1], light.radiance[2] }; }

inline RGB infinite_evaluate(const Scene& scene, const EchoInfiniteLight& light, Float3 incident)
{
	if (light.type == ECHO_INFINITE_ENVIRONMENT) // AmbientLight.Evaluate, AmbientLight.cs:53-54
		return infinite_radiance(light) * environment_texel(scene, light, cylindrical_to_uv(rotate3x3(light.inverseRotation, incident)));

	if (light.type == ECHO_INFINITE_CUBEMAP) return infinite_radiance(light) * cubemap_evaluate(scene, light, rotate3x3(light.inverseRotation, incident));
	if (light.type != ECHO_INFINITE_DIRECTIONAL) return infinite_radiance(light);
	if (light.isDelta) return kBlack;

	float cosIncident = dot(f3(light.direction), incident);
	if (cosIncident <= light.cosAngle) return kBlack;
	return infinite_radiance(light); // scaledIntensity
}

inline float infinite_pdf(const Scene& scene, const EchoInfiniteLight& light, Float3 incident)
{
	if (light.type == ECHO_INFINITE_ENVIRONMENT) return cylindrical_pdf(scene, light, rotate3x3(light.inverseRotation, incident)); // AmbientLight.cs:56-57
	if (light.type != ECHO_INFINITE_DIRECTIONAL) return kUniformSpherePdf;
	if (light.isDelta) return 0.0f;
	return uniform_cone_pdf(light.cosAngle);
}

inline ProbableRGB infinite_sample(const Scene& scene, const EchoInfiniteLight& light, Float2 sample, Float3& incident, float& travel)
{
	travel = kInfinity;

	if (light.type == ECHO_INFINITE_ENVIRONMENT) // AmbientLight.Sample, AmbientLight.cs:59-66
	{
		ProbableRGB sampled = cylindrical_sample(scene, light, sample, incident);
		incident = rotate3x3(light.rotation, incident);
		return { sampled.content * infinite_radiance(light), sampled.pdf };
	}

	if (light.type == ECHO_INFINITE_CUBEMAP) // IDirectionalTexture.Sample's default (IDirectionalTexture.cs:22-26) inside AmbientLight.Sample
	{
		Float3 local = uniform_sphere(sample);
		incident = rotate3x3(light.rotation, local);
		return { cubemap_evaluate(scene, light, local) * infinite_radiance(light), kUniformSpherePdf };
	}

	if (light.type != ECHO_INFINITE_DIRECTIONAL)
	{
		// AmbientLight.cs:60-67 with Pure as IDirectionalTexture: the default interface Sample draws a uniform sphere
		// direction with pdf 1/(4 pi) (Textures/Directional/IDirectionalTexture.cs); rotation is the identity in scope.
		incident = uniform_sphere(sample);
		return { infinite_radiance(light), kUniformSpherePdf };
	}

	if (light.isDelta)
	{
		incident = f3(light.direction);
		return { RGB{ light.intensity[0], light.intensity[1], light.intensity[2] }, 1.0f };
	}

	Float3 local = uniform_cone(sample, light.cosAngle);
	local = { local.x, local.y, -local.z }; // Utility.NegateZ
	const float* m = light.rotation;        // Float3x3 * Float3, Float3x3.cs:264-269
	incident = { m[0] * local.x + m[1] * local.y + m[2] * local.z, m[3] * local.x + m[4] * local.y + m[5] * local.z, m[6] * local.x + m[7] * local.y + m[8] * local.z };
	return { infinite_radiance(light), uniform_cone_pdf(light.cosAngle) };
}

// ---- PreparedScene.Sample (PreparedScene.cs:182-204) -> LightCollection.Sample (LightCollection.cs:141-193),
//      PreparedPointLight.Sample (Scenic/Lights/PointLight.cs:48-66), AmbientLight.Sample over a Pure texture ----
inline ProbableRGB scene_sample_light(const Scene& scene, uint32_t light, const Layers& layers, const GeometryPoint& origin, Float2 sample, Float3& incident, float& travel)
{
	incident = { 0.0f, 0.0f, 0.0f };
	travel = 0.0f;

	if (token_is_infinite_light(light))
	{
		return infinite_sample(scene, scene.infiniteLights[token_light_index(light)], sample, incident, travel);
	}

	// FindLayer + `forwardTransform * origin` (PreparedScene.cs:192-198, GeometryPoint.cs:41-45): the shading point in the
	// space of the pack that owns the light
	Scene::Layer layer = scene.find_layer(layers);
	EchoPack view = scene.pack_view(layer.pack);
	GeometryPoint local = { multiply_point(layer.forward, origin.position), normalized(multiply_direction(layer.forward, origin.normal)) };
	ProbableRGB result = {};

	if (token_type(light) == ECHO_TOKEN_TYPE_LIGHT) // point light, PointLight.cs:48-66
	{
		const EchoPointLight& point = scene.pointLights[view.pointLightOffset + token_light_index(light)];
		Float3 offset = f3(point.position) - local.position;
		float travel2 = squared_magnitude(offset);

		if (!positive(travel2)) return {};

		travel = sqrt0(travel2);
		float travelR = 1.0f / travel;
		incident = offset * travelR;

		RGB intensity = { point.intensity[0], point.intensity[1], point.intensity[2] };
		result = { intensity * travelR * travelR, 1.0f };
	}
	else
	{
		// emissive geometry: LightCollection.HandleGeometry, LightCollection.cs:166-183; the material comes from the PACK's
		// swatch (geometries.swatch, :145,151), not from the placement's
		uint32_t materialIndex = view.materialOffset + scene.geometry_material(light, layer.pack);
		const EchoMaterial& material = scene.materials[materialIndex];
		if (material.type != ECHO_MATERIAL_EMISSIVE) return {};

		GeometryPoint point;
		float pdf;
		if (!geometry_sample(scene, light, local.position, sample, point, pdf, layer.pack)) return {};
		if (!positive(pdf)) return {};

		Float3 delta = point.position - local.position;
		float travel2 = squared_magnitude(delta);
		if (!positive(travel2)) return {};

		travel = sqrt0(travel2);
		incident = delta * (1.0f / travel);

		travel *= 1.0f - 2E-5f; // TravelMultiplier, LightCollection.cs:89
		result = { emissive_emit(material, point, -incident), pdf };
	}

	// back to world space, PreparedScene.cs:200-201
	incident = normalized(multiply_direction(layer.inverse, incident));
	travel *= get_scale(layer.inverse); // Utility.GetScale, Utility.cs:82
	return result;
}

// ---- PreparedScene.ProbabilityDensity (PreparedScene.cs:207-225) -> LightCollection.ProbabilityDensity (:196-219) ----
inline float scene_light_pdf(const Scene& scene, uint32_t light, const Layers& layers, const GeometryPoint& origin, Float3 incident)
{
	if (token_is_infinite_light(light)) return infinite_pdf(scene, scene.infiniteLights[token_light_index(light)], incident);

	Scene::Layer layer = scene.find_layer(layers);
	if (token_type(light) == ECHO_TOKEN_TYPE_LIGHT) return 1.0f; // LightCollection.cs:206 (point)

	// forwardTransform * origin, forwardTransform.MultiplyDirection(incident).Normalized (:219-223)
	Float3 position = multiply_point(layer.forward, origin.position);
	Float3 direction = normalized(multiply_direction(layer.forward, incident));
	return geometry_pdf(scene, light, position, direction, layer.pack);
}

// ---- PreparedScene.EvaluateInfinite (PreparedScene.cs:233-253) ----
inline RGB evaluate_infinite(const Scene& scene, Float3 direction, bool direct)
{
	RGB total = kBlack;

	for (const EchoInfiniteLight& light : scene.infiniteLights)
	{
		if (direct && !light.directlyVisible) continue;
		total = total + infinite_evaluate(scene, light, direction);
	}

	return total;
}

// per-sample random numbers in the reference's draw order (CameraSample.cs:19-23; PathTracedEvaluator.cs:60-66)
struct SampleStream
{
	uint32_t key;
	uint32_t dimension = 0;

	float next1d() { return sample_value(key, dimension++); }

	Float2 next2d()
	{
		float x = next1d();
		float y = next1d();
		return { x, y };
	}
};

struct EvaluatorStats // the labels of EvaluatorStatistics used on this path, echo_b200.h EchoStats order
{
	uint64_t bounceCreated = 0, bounceSpecular = 0, bounceMis = 0;
	uint64_t lightSampled = 0, lightOcclusionChecked = 0, lightOcclusionPassed = 0, lightEvaluatedInfinite = 0;
	uint64_t traceQueries = 0, occludeQueries = 0;
	VisitCounters visits;         // node / triangle / sphere visits of those queries (echo_b200.h EchoStats nodeVisits ...)
	uint64_t lightNodeVisits = 0; // LightBound.Importance evaluations of scene.pick / scene.probability_mass

	void add(const EvaluatorStats& o)
	{
		visits.nodes += o.visits.nodes; visits.triangles += o.visits.triangles; visits.spheres += o.visits.spheres;
		lightNodeVisits += o.lightNodeVisits;
		bounceCreated += o.bounceCreated; bounceSpecular += o.bounceSpecular; bounceMis += o.bounceMis;
		lightSampled += o.lightSampled; lightOcclusionChecked += o.lightOcclusionChecked;
		lightOcclusionPassed += o.lightOcclusionPassed; lightEvaluatedInfinite += o.lightEvaluatedInfinite;
		traceQueries += o.traceQueries; occludeQueries += o.occludeQueries;
	}
};

// ---- Evaluation/Evaluators/PathTracedEvaluator.cs ----
struct PathTracedEvaluator
{
	int bounceLimit = 128;      // :33
	float survivability = 2.5f; // :40

	static float power_heuristic(float pdf0, float pdf1) // :213-217
	{
		float squared = pdf0 * pdf0;
		return squared / (squared + pdf1 * pdf1);
	}

	struct Path // :223-321
	{
		RGB result = kBlack;
		RGB energy = kWhite;
		Contact contact;
		TraceQuery query;

		Float3 current_direction() const { return query.ray.direction; }

		bool advance(const Scene& scene, EvaluatorStats& stats)
		{
			++stats.traceQueries;
			if (!scene.trace(query, &stats.visits)) return false;
			interact(scene, query, contact);
			return true;
		}

		void contribute(RGB value) { result = result + energy * value; }

		void contribute_emissive(const Scene& scene, float weight = 1.0f)
		{
			const EchoMaterial& material = scene.materials[contact.material];
			if (material.type != ECHO_MATERIAL_EMISSIVE) return;
			if (!positive(emissive_power(material))) return;
			contribute(emissive_emit(material, contact.point, contact.outgoing) * weight);
		}

		bool russian_roulette(float survivability, float sample) // :313-320
		{
			float rate = clamp01(survivability * luminance(energy));
			if (sample >= rate) return false;
			energy = energy / rate;
			return true;
		}

		bool proceed(RGB scatter, float scatterPdf, Float3 incident, float survivability, float sample) // Continue, :282-293
		{
			if (!positive(scatterPdf)) return false;

			energy = energy * (scatter / scatterPdf);

			bool survived = russian_roulette(survivability, sample);

			if (survived)
			{
				// TraceQuery.SpawnTrace, TraceQuery.cs:88 — new origin is the hit position, ignore = hit token
				TraceQuery spawned;
				spawned.ray = Ray(query.position(), incident);
				spawned.ignore = query.token;
				spawned.ignoreLayers = query.tokenLayers;
				query = spawned;
			}

			return survived;
		}
	};

	// Test-only switch (oracle_set_failed_pick_keeps_mis, never set by parity tests): the reference turns MIS off when the
	// light pick fails (:165-169 `mis = false`), so the BSDF-sampled ray then adds the emission it finds with full weight
	// although the lights it can find were also reachable by Pick on other draws — an excess of P(pick fails) * w_light,
	// 3-5 % on scenes with many oriented emitters. With the switch on, a failed pick keeps MIS and the estimator is unbiased,
	// which is what lets StandardNaiveEvaluator validate everything else in this evaluator (tests/test_auxiliary_evaluators.py).
	static inline bool failedPickKeepsMis = false;

	// :162-207
	RGB importance_sample_radiant(const Scene& scene, const Contact& contact, EvaluatorStats& stats, float lightSample, Float2 radiantSample, bool& mis) const
	{
		float lightPdf;
		Layers lightLayers;
		uint32_t light = scene.pick(contact.point, lightSample, lightPdf, &lightLayers);

		if (!positive(lightPdf))
		{
			mis = failedPickKeepsMis;
			return kBlack;
		}

		Float3 incident;
		float travel;
		ProbableRGB radiantSampled = scene_sample_light(scene, light, lightLayers, contact.point, radiantSample, incident, travel);
		RGB radiant = radiantSampled.content;

		float pdf = lightPdf * radiantSampled.pdf;
		mis = token_is_area_light(light);

		if (!positive(pdf) || is_zero(radiant)) return kBlack;
		++stats.lightSampled;

		RGB scatter = contact.bsdf.evaluate(contact.outgoing, incident);
		scatter = scatter * contact.normal_dot(incident);

		if (is_zero(scatter)) return kBlack;
		++stats.lightOcclusionChecked;

		OccludeQuery query; // Contact.SpawnOcclude, Contact.cs:86
		query.ray = Ray(contact.point.position, incident);
		query.travel = travel;
		query.ignore = contact.token;
		query.ignoreLayers = contact.layers;
		++stats.occludeQueries;
		if (scene.occlude(query, &stats.visits)) return kBlack;

		++stats.lightOcclusionPassed;

		radiant = radiant * (scatter / pdf);
		if (!mis) return radiant;

		float scatterPdf = contact.bsdf.probability_density(contact.outgoing, incident);
		return radiant * power_heuristic(pdf, scatterPdf);
	}

	// :43-150
	RGB evaluate(const Scene& scene, const Ray& ray, SampleStream& distribution, EvaluatorStats& stats) const
	{
		Path path;
		path.query.ray = ray;

		if (!path.advance(scene, stats))
		{
			++stats.lightEvaluatedInfinite;
			return evaluate_infinite(scene, path.current_direction(), true);
		}

		path.contribute_emissive(scene);

		for (int depth = 0; depth < bounceLimit; depth++)
		{
			// Bounce, :326-354
			Float3 bounceIncident;
			const BxDF* function;
			ProbableRGB bounced = path.contact.bsdf.sample(path.contact.outgoing, distribution.next2d(), bounceIncident, function);
			RGB bounceScatter = bounced.content * path.contact.normal_dot(bounceIncident);
			float bounceScatterPdf = bounced.pdf;
			++stats.bounceCreated;

			float survivalSample = distribution.next1d();
			float lightSample = distribution.next1d();
			Float2 radiantSample = distribution.next2d();

			if (!positive(bounceScatterPdf) || type_any(function->type, Specular))
			{
				++stats.bounceSpecular;
				if (!path.proceed(bounceScatter, bounceScatterPdf, bounceIncident, survivability, survivalSample)) break;
			}
			else
			{
				bool mis;
				path.contribute(importance_sample_radiant(scene, path.contact, stats, lightSample, radiantSample, mis));

				if (!path.proceed(bounceScatter, bounceScatterPdf, bounceIncident, survivability, survivalSample)) break;

				if (mis)
				{
					GeometryPoint oldPoint = path.contact.point;
					++stats.bounceMis;

					if (path.advance(scene, stats))
					{
						uint32_t light = path.contact.token;
						float pmf = scene.probability_mass(light, oldPoint, path.contact.layers);
						if (!positive(pmf)) continue;

						float pdf = scene_light_pdf(scene, light, path.contact.layers, oldPoint, path.current_direction());
						if (!positive(pdf)) continue;

						path.contribute_emissive(scene, power_heuristic(bounceScatterPdf, pmf * pdf));
						continue;
					}

					Float3 direction = path.current_direction();

					for (size_t i = 0; i < scene.infiniteLights.size(); i++)
					{
						const EchoInfiniteLight& light = scene.infiniteLights[i];
						if (light.isDelta) continue; // "Skip delta lights; they do not like MIS", :118

						uint32_t token = ECHO_LIGHT_TOKEN_MAKE(ECHO_LIGHT_TYPE_INFINITE, (uint32_t)i);
						float pdf = scene.probability_mass(token, oldPoint) * infinite_pdf(scene, light, direction);
						if (!positive(pdf)) continue;

						float weight = power_heuristic(bounceScatterPdf, pdf);
						path.contribute(infinite_evaluate(scene, light, direction) * weight);
					}

					++stats.lightEvaluatedInfinite;
					break;
				}
			}

			if (!path.advance(scene, stats))
			{
				++stats.lightEvaluatedInfinite;
				path.contribute(evaluate_infinite(scene, path.current_direction(), false));
				break;
			}

			path.contribute_emissive(scene);
		}

		return path.result;
	}
};

struct Float4 // Common/Packed/Float4.cs, as the evaluators return it and the Accumulator sums it
{
	float v[4];
};

// ---- Common/Mathematics/Scalars.cs:153-168 (float.AlmostEquals) and Float3.Equals (Common/Packed/Float3.cs:369) ----
inline bool almost_equals(float value, float other, float epsilon = 1E-5f)
{
	if (value == other) return true;
	const float normal = 1.17549435E-38f; // (1L << 23) * float.Epsilon

	float difference = std::fabs(value - other);
	if (value == 0.0f || other == 0.0f || difference < normal) return difference < epsilon * normal;

	float sum = std::fabs(value) + std::fabs(other);
	return difference < epsilon * math_min(sum, 3.40282347E+38f);
}

inline bool float3_equals(Float3 a, Float3 b) { return almost_equals(a.x, b.x) && almost_equals(a.y, b.y) && almost_equals(a.z, b.z); }

// ---- Evaluation/Evaluators/AlbedoEvaluator.cs:18-55 and NormalDepthEvaluator.cs:20-60: follow purely specular bounces from
// the camera, report the first other surface. They share everything but Exit() and the escaped-ray value. ----
struct AuxiliaryEvaluator
{
	int kind = ECHO_EVALUATOR_ALBEDO;
	bool divergeOnce = true;

	Float4 evaluate(const Scene& scene, const Ray& ray, SampleStream& distribution) const
	{
		TraceQuery query;
		query.ray = ray;
		bool direct = true; // whether the path is still going in its original direction
		float depth = 0.0f;
		Contact contact;

		auto exit = [&]() -> Float4
		{
			if (kind == ECHO_EVALUATOR_ALBEDO)
			{
				EchoMaterial material = resolve_material(scene, contact.material, contact.texcoord); // (RGB128)material.SampleAlbedo(contact)
				return { { material.albedo[0], material.albedo[1], material.albedo[2], 0.0f } };
			}

			return { { contact.shadeNormal.x, contact.shadeNormal.y, contact.shadeNormal.z, depth } }; // NormalDepth128.ToFloat4
		};

		// `while (scene.Trace(ref query))`; the cap only guards against a hall of mirrors (the device has the same one)
		for (int bounce = 0; bounce < 1024 && scene.trace(query); bounce++)
		{
			interact(scene, query, contact); // Interact + material.Scatter (Scatter has no effect on Exit())

			if (!direct) return exit();
			if (kind == ECHO_EVALUATOR_NORMAL_DEPTH) depth += query.distance;

			int specular = 0;
			for (int i = 0; i < contact.bsdf.count; i++) specular += type_any(contact.bsdf.functions[i]->type, Specular) ? 1 : 0;
			if (contact.bsdf.count != specular) return exit(); // not fully specular

			Float3 incident;
			const BxDF* function;
			ProbableRGB sample = contact.bsdf.sample(contact.outgoing, distribution.next2d(), incident, function);
			if (sample.not_possible()) return exit();

			if (!float3_equals(incident, ray.direction)) direct = false; // compared with the CAMERA ray's direction
			if (!divergeOnce && !direct) return exit();

			TraceQuery spawned; // query.SpawnTrace(incident), TraceQuery.cs:88
			spawned.ray = Ray(query.position(), incident);
			spawned.ignore = query.token;
			spawned.ignoreLayers = query.tokenLayers;
			query = spawned;
		}

		if (kind == ECHO_EVALUATOR_ALBEDO)
		{
			RGB infinite = evaluate_infinite(scene, query.ray.direction, direct);
			return { { infinite.r, infinite.g, infinite.b, 0.0f } };
		}

		if (direct) depth = scene.boundRadius * 2.0f; // negative direction and scene diameter for escaped rays
		Float3 outward = -ray.direction;
		return { { outward.x, outward.y, outward.z, depth } };
	}
};

// ---- Evaluation/Evaluators/StandardNaiveEvaluator.cs:16-55: the recursion `scatter * Evaluate(depth + 1) + emission` ----
struct NaiveEvaluator
{
	int bounceLimit = 128;

	RGB evaluate(const Scene& scene, TraceQuery& query, SampleStream& distribution, EvaluatorStats& stats, int depth) const
	{
		if (depth == bounceLimit) return evaluate_infinite(scene, query.ray.direction, depth == 0);
		++stats.traceQueries;
		if (!scene.trace(query)) return evaluate_infinite(scene, query.ray.direction, depth == 0);

		Contact contact;
		interact(scene, query, contact);

		const EchoMaterial& material = scene.materials[contact.material]; // `material is Emissive emissive ? emissive.Emit(...) : Black`
		RGB emission = material.type == ECHO_MATERIAL_EMISSIVE ? emissive_emit(material, contact.point, contact.outgoing) : kBlack;

		Float3 incident;
		const BxDF* function;
		ProbableRGB sample = contact.bsdf.sample(contact.outgoing, distribution.next2d(), incident, function);
		++stats.bounceCreated;
		if (sample.not_possible() || is_zero(sample.content)) return emission;

		RGB scatter = sample.content / sample.pdf;
		scatter = scatter * contact.normal_dot(incident);

		TraceQuery spawned; // query.SpawnTrace(incident)
		spawned.ray = Ray(query.position(), incident);
		spawned.ignore = query.token;
		spawned.ignoreLayers = query.tokenLayers;

		return scatter * evaluate(scene, spawned, distribution, stats, depth + 1) + emission;
	}
};

// the evaluator EchoRenderParams selects, as Float4 (RGB128 radiance has W = 0)
inline Float4 evaluate_sample(const Scene& scene, const EchoRenderParams& params, const Ray& ray, SampleStream& distribution, EvaluatorStats& stats)
{
	int kind = params.evaluator & ECHO_EVALUATOR_KIND_MASK;

	if (kind == ECHO_EVALUATOR_PATH_TRACED)
	{
		PathTracedEvaluator evaluator;
		evaluator.bounceLimit = params.bounceLimit;
		evaluator.survivability = params.survivability;
		uint64_t lightVisitsBefore = lightNodeVisits;
		RGB value = evaluator.evaluate(scene, ray, distribution, stats);
		stats.lightNodeVisits += lightNodeVisits - lightVisitsBefore;
		return { { value.r, value.g, value.b, 0.0f } };
	}

	if (kind == ECHO_EVALUATOR_NAIVE)
	{
		NaiveEvaluator evaluator;
		evaluator.bounceLimit = params.bounceLimit;
		TraceQuery query;
		query.ray = ray;
		RGB value = evaluator.evaluate(scene, query, distribution, stats, 0);
		return { { value.r, value.g, value.b, 0.0f } };
	}

	AuxiliaryEvaluator evaluator;
	evaluator.kind = kind;
	evaluator.divergeOnce = (params.evaluator & ECHO_EVALUATOR_DIVERGE_ONCE) != 0;
	return evaluator.evaluate(scene, ray, distribution);
}

// ---- Scenic/Cameras/RaySpawner.cs:19-46 + PerspectiveCamera.cs:51-98 ----
struct CameraSpawner
{
	float sizeRX, sizeRY, offsetY;

	CameraSpawner(int width, int height)
	{
		sizeRY = 1.0f / (float)height;
		sizeRX = 1.0f / (float)width;                        // TextureGrid.cs:22
		float aspectY = (float)height / (float)width;        // TextureGrid.cs:25-29
		offsetY = aspectY / -2.0f;                           // RaySpawner.cs:22
	}

	Float2 spawn_x(int px, int py, Float2 shift) const // RaySpawner.cs:37-46
	{
		float sx = shift.x + (float)px;
		float sy = shift.y + (float)py;
		return { fma_f(sx, sizeRX, 1.0f / -2.0f), fma_f(sy, sizeRX, offsetY) };
	}
};

inline Float3 camera_multiply_direction(const EchoCamera& c, Float3 d) // Float4x4.cs:267-272
{
	const float* m = c.transform;
	return {
		m[0] * d.x + m[1] * d.y + m[2] * d.z,
		m[4] * d.x + m[5] * d.y + m[6] * d.z,
		m[8] * d.x + m[9] * d.y + m[10] * d.z
	};
}

inline Float3 camera_multiply_point(const EchoCamera& c, Float3 p) // Float4x4.cs:260-265
{
	const float* m = c.transform;
	return {
		m[0] * p.x + m[1] * p.y + m[2] * p.z + m[3],
		m[4] * p.x + m[5] * p.y + m[6] * p.z + m[7],
		m[8] * p.x + m[9] * p.y + m[10] * p.z + m[11]
	};
}

inline Ray camera_spawn_ray(const EchoCamera& camera, const CameraSpawner& spawner, int px, int py, Float2 shift, Float2 lens)
{
	if (camera.type == ECHO_CAMERA_ORTHOGRAPHIC) // OrthographicCamera.cs:33-38
	{
		Float2 uv = spawner.spawn_x(px, py, shift);
		Float3 origin = { uv.x * camera.width, uv.y * camera.width, 0.0f };
		return Ray(camera_multiply_point(camera, origin), f3(camera.direction));
	}

	if (camera.type == ECHO_CAMERA_CYLINDRICAL) // CylindricalCamera.cs:27-33 + CylindricalTexture.ToDirection, CylindricalTexture.cs:153-164
	{
		Float2 uv = { (shift.x + (float)px) * spawner.sizeRX, (shift.y + (float)py) * spawner.sizeRY };
		float sinT, cosT, sinP, cosP;
		sincos_det(uv.x * kTau, sinT, cosT);
		sincos_det(uv.y * kPi, sinP, cosP);
		Float3 direction = normalized(Float3{ -sinP * sinT, -cosP, -sinP * cosT });
		direction = normalized(camera_multiply_direction(camera, direction)); // (Float3x3)RootedRotation * direction
		return Ray(Float3{ camera.transform[3], camera.transform[7], camera.transform[11] }, direction);
	}

	bool hasDepthOfField = positive(camera.lensRadius) && positive(camera.focalDistance); // PerspectiveCamera.cs:46

	if (!hasDepthOfField) // :93-98
	{
		Float2 uv = spawner.spawn_x(px, py, shift);
		Float3 direction = camera_multiply_direction(camera, Float3{ uv.x, uv.y, camera.forwardLength });
		Float3 position = { camera.transform[3], camera.transform[7], camera.transform[11] }; // RootedPosition
		return Ray(position, normalized(direction));
	}

	// :55-68
	float focusScale = camera.focalDistance / camera.forwardLength;
	Float2 disk = concentric_disk(lens);
	Float2 lensPoint = { disk.x * camera.lensRadius, disk.y * camera.lensRadius };
	Float2 uv = spawner.spawn_x(px, py, shift);
	Float3 focus = { uv.x * focusScale, uv.y * focusScale, camera.focalDistance };

	Float3 origin = { lensPoint.x, lensPoint.y, 0.0f };
	Float3 direction = focus - origin;

	return Ray(camera_multiply_point(camera, origin), normalized(camera_multiply_direction(camera, direction)));
}

// ---- Common/Mathematics/Primitives/Summation.cs (Kahan) on one Float4 lane-set; W lane carried like the reference ----
inline Float4 f4_op(Float4 a, Float4 b, char op)
{
	Float4 r;
	for (int i = 0; i < 4; i++) r.v[i] = op == '+' ? a.v[i] + b.v[i] : (op == '-' ? a.v[i] - b.v[i] : a.v[i] * b.v[i]);
	return r;
}

struct Summation
{
	Float4 total = { { 0, 0, 0, 0 } }, error = { { 0, 0, 0, 0 } };

	Summation add(Float4 value) const // Summation.cs:31-38
	{
		Float4 delta = f4_op(value, error, '-');
		Float4 newTotal = f4_op(total, delta, '+');
		Float4 newError = f4_op(f4_op(newTotal, total, '-'), delta, '-');
		return { newTotal, newError };
	}

	Summation add(const Summation& value) const // Summation.cs:42-51
	{
		Float4 newError = f4_op(error, value.error, '+');
		Float4 delta = f4_op(value.total, newError, '-');
		Float4 newTotal = f4_op(total, delta, '+');
		newError = f4_op(f4_op(newTotal, total, '-'), delta, '-');
		return { newTotal, newError };
	}

	Summation scale(Float4 value) const { return { f4_op(total, value, '*'), f4_op(error, value, '*') }; } // Summation.cs:40

	Summation negated() const
	{
		Summation r;
		for (int i = 0; i < 4; i++) { r.total.v[i] = -total.v[i]; r.error.v[i] = -error.v[i]; }
		return r;
	}
};

// ---- Processes/Evaluation/Accumulator.cs:11-71 ----
struct Accumulator
{
	Summation average, squared;
	uint32_t count = 0;

	bool add(Float4 sample)
	{
		float sum = (sample.v[0] + sample.v[1]) + (sample.v[2] + sample.v[3]); // Float4.Sum, Float4.cs:73-81
		if (!std::isfinite(sum)) return false;

		++count;

		Float4 negative = { { -sample.v[0], -sample.v[1], -sample.v[2], -sample.v[3] } };
		Summation delta = average.add(negative); // average - sample

		float countR = 1.0f / (float)count;      // Summation operator / : multiply by 1f / value
		Float4 countRV = { { countR, countR, countR, countR } };
		average = average.add(delta.scale(countRV).negated()); // average -= delta / count

		Summation after = average.add(negative);                // (average - sample)
		squared = squared.add(delta.scale(after.total));        // squared += delta * (average - sample).Result
		return true;
	}

	Float4 value() const { return average.total; }

	// Accumulator.Noise (:28-51) with exact 1/x and 1/sqrt(x) in place of rcpps/rsqrtps (12-bit hardware
	// approximations whose bits differ between CPU vendors); only the adaptive-epoch stop test reads it.
	float noise_max() const
	{
		if (count < 2) return 0.0f;

		float oneLess = (float)(count - 1u);
		oneLess *= oneLess * oneLess;

		float result = 0.0f;

		for (int i = 0; i < 4; i++)
		{
			float mean = average.total.v[i];
			float numerator = mean * mean * oneLess;
			float denominator = 1.0f / squared.total.v[i];
			float noise = numerator != 0.0f ? 1.0f / std::sqrt(numerator * denominator) : 0.0f;
			if (i == 0 || noise > result) result = noise; // Float4.MaxComponent over the masked lanes
		}

		return result;
	}
};

// ---- Processes/Evaluation/EvaluationOperation.cs:100-141, one pixel ----
inline Float4 evaluate_pixel(const Scene& scene, const EchoRenderParams& params, int px, int py, EvaluatorStats& stats, uint64_t& samples, uint64_t& rejected)
{
	CameraSpawner spawner(params.width, params.height);
	Accumulator accumulator;
	uint32_t pixel = (uint32_t)py * (uint32_t)params.width + (uint32_t)px;

	int epoch = 0;

	do
	{
		++epoch;

		for (int i = 0; i < params.extend; i++)
		{
			uint32_t sampleIndex = (uint32_t)(params.epochOffset + epoch - 1) * (uint32_t)params.extend + (uint32_t)i;
			SampleStream distribution{ sample_key(params.seed, pixel, sampleIndex) };

			Float2 shift = distribution.next2d(); // CameraSample.Create, CameraSample.cs:19-23
			Float2 lens = distribution.next2d();
			Ray ray = camera_spawn_ray(scene.camera, spawner, px, py, shift, lens);

			Float4 evaluated = evaluate_sample(scene, params, ray, distribution, stats);
			++samples;

			if (!accumulator.add(evaluated)) ++rejected;
		}
	}
	while (epoch < params.maxEpoch && (epoch < params.minEpoch || accumulator.noise_max() > params.noiseThreshold));

	return accumulator.value();
}

} // namespace oracle
