// oracle/aggregation.hpp — TEST INFRASTRUCTURE ONLY. CPU restatement of Echo.Core/Aggregation for the hot path:
// rays/queries, the 4-wide slab test, QBVH closest-hit/occlusion traversal, triangle and sphere intersection, the
// scene-level guards, the light tree and light sampling. One query at a time, recursive/sequential like the reference.
// Citations are relative to /root/reference/src/Echo.Core/.
#pragma once
#include <alloca.h>
#include <unordered_map>
#include <vector>

#include "../include/echo_b200.h"
#include "math.hpp"

namespace oracle
{

// ---- Aggregation/Primitives/EntityToken.cs:22-73, TokenType.cs:12-38 ----
inline uint32_t token_type(uint32_t token) { return token >> ECHO_TOKEN_INDEX_BITS; }
inline uint32_t token_index(uint32_t token) { return token & ((1u << ECHO_TOKEN_INDEX_BITS) - 1u); }
inline uint32_t token_light_type(uint32_t token) { return token_index(token) >> ECHO_LIGHT_INDEX_BITS; }
inline uint32_t token_light_index(uint32_t token) { return token & ((1u << ECHO_LIGHT_INDEX_BITS) - 1u); }
inline bool token_is_geometry(uint32_t token) { uint32_t t = token_type(token); return t >= 1u && t <= 3u; } // TokenType.cs IsGeometry
inline bool token_is_raw_geometry(uint32_t token) { uint32_t t = token_type(token); return t == 1u || t == 2u; }

inline bool token_is_area_light(uint32_t token) // EntityToken.cs:187-192
{
	if (token_is_raw_geometry(token)) return true;
	if (token_type(token) != ECHO_TOKEN_TYPE_LIGHT) return false;
	return token_light_type(token) == ECHO_LIGHT_TYPE_INFINITE;
}

inline bool token_is_infinite_light(uint32_t token) // EntityToken.cs:194-202
{
	if (token_type(token) != ECHO_TOKEN_TYPE_LIGHT) return false;
	uint32_t lightType = token_light_type(token);
	return lightType == ECHO_LIGHT_TYPE_INFINITE || lightType == ECHO_LIGHT_TYPE_INFINITE_DELTA;
}

// ---- Aggregation/Primitives/Ray.cs:17-28 ----
struct Ray
{
	Float3 origin, direction, directionR;

	Ray() = default;

	Ray(Float3 o, Float3 d) : origin(o), direction(d), directionR{ 1.0f / d.x, 1.0f / d.y, 1.0f / d.z } {}

	Float3 get_point(float distance) const { return direction * distance + origin; } // Ray.cs:33
};

// ---- Aggregation/Primitives/TokenHierarchy.cs:19-129: the instance layers; the TopToken is kept beside it ----
struct Layers
{
	uint32_t count = 0;
	uint32_t instances[ECHO_MAX_INSTANCE_LAYERS] = {};

	void push(uint32_t token) { instances[count++] = token; } // :62-70
	void pop() { --count; }                                   // :72-81

	bool operator==(const Layers& other) const // :117-129 (the hash is only a shortcut there)
	{
		if (count != other.count) return false;
		for (uint32_t i = 0; i < count; i++) if (instances[i] != other.instances[i]) return false;
		return true;
	}
};

// ---- Aggregation/Primitives/TraceQuery.cs:16-70; a TokenHierarchy is {layers, top token} ----
struct TraceQuery
{
	Ray ray;
	uint32_t ignore = ECHO_TOKEN_EMPTY;
	uint32_t token = ECHO_TOKEN_EMPTY;
	float distance = kInfinity;
	Float2 uv = { 0.0f, 0.0f };
	Layers ignoreLayers, tokenLayers, current;

	Float3 position() const { return ray.get_point(math_max(distance, kEpsilon)); } // TraceQuery.cs:76-82
};

// ---- Aggregation/Primitives/OccludeQuery.cs:10-44 ----
struct OccludeQuery
{
	Ray ray;
	uint32_t ignore = ECHO_TOKEN_EMPTY;
	float travel = kInfinity;
	Layers ignoreLayers, current;
};

// per-query visit counters: the algorithmic-bytes inputs of SURVEY.md §8(d) (a node visit = one BoxBound4.Intersect,
// a leaf visit = one primitive test). They mirror what Accelerator.TraceCost estimates (QuadBoundingVolumeHierarchy.cs:317-361).
struct VisitCounters
{
	uint64_t nodes = 0, triangles = 0, spheres = 0;
};

// LightBound.Importance evaluations of the calling thread (one 64-byte light-tree node each): the D_lt term of SURVEY.md §8(d).
// The path tracer's statistics take the difference around a sample.
inline thread_local uint64_t lightNodeVisits = 0;

struct GeometryPoint
{
	Float3 position, normal;
};

// Aggregation/Primitives/GeometryPoint.cs:28-39
inline float geometry_point_pdf(const GeometryPoint& point, Float3 origin, float area)
{
	Float3 delta = point.position - origin;
	float length2 = squared_magnitude(delta);
	float length = sqrt0(length2);

	float d = fabs_bits(dot(point.normal, delta));
	if (!positive(d)) return 0.0f;
	return length2 * length / (d * area);
}

// ---- Scenic/Geometries/TriangleEntity.cs (PreparedTriangle) ----
inline Float3 f3(const float* p) { return { p[0], p[1], p[2] }; }

// TriangleEntity.cs:204-235
inline float triangle_intersect(const EchoTriangle& tri, Float3 origin, Float3 direction, Float2& uv)
{
	Float3 vertex0 = f3(tri.vertex0), edge1 = f3(tri.edge1), edge2 = f3(tri.edge2);

	Float3 cross2 = cross(direction, edge2);
	float determinant = dot(edge1, cross2);

	if (determinant == 0.0f) return kInfinity;
	float determinantR = 1.0f / determinant;

	Float3 offset = origin - vertex0;
	float u = dot(offset, cross2) * determinantR;
	uv.x = u; // written through the ref before the early-outs (TriangleEntity.cs:209-210,221)

	if ((u < 0.0f) | (u > 1.0f)) return kInfinity;

	Float3 cross1 = cross(offset, edge1);
	float v = dot(direction, cross1) * determinantR;
	uv.y = v;

	if ((v < 0.0f) | (u + v > 1.0f)) return kInfinity;

	float distance = dot(edge2, cross1) * determinantR;
	return distance < 0.0f ? kInfinity : distance;
}

// TriangleEntity.cs:237-263
inline bool triangle_occlude(const EchoTriangle& tri, Float3 origin, Float3 direction, float travel)
{
	Float3 vertex0 = f3(tri.vertex0), edge1 = f3(tri.edge1), edge2 = f3(tri.edge2);

	Float3 cross2 = cross(direction, edge2);
	float determinant = dot(edge1, cross2);

	if (determinant == 0.0f) return false;
	float sign = determinant < 0.0f ? -1.0f : 1.0f; // MathF.Sign
	determinant *= sign;

	Float3 offset = origin - vertex0;
	float u = dot(offset, cross2) * sign;

	if ((u < 0.0f) | (u > determinant)) return false;

	Float3 cross1 = cross(offset, edge1);
	float v = dot(direction, cross1) * sign;

	if ((v < 0.0f) | (u + v > determinant)) return false;

	float distance = dot(edge2, cross1) * sign;
	return (distance >= 0.0f) & (distance < travel * determinant);
}

inline Float3 triangle_normal(const EchoTriangle& tri) { return normalized(cross(f3(tri.edge1), f3(tri.edge2))); } // TriangleEntity.cs:136
inline float triangle_area(const EchoTriangle& tri) { return magnitude(cross(f3(tri.edge1), f3(tri.edge2))) / 2.0f; } // TriangleEntity.cs:148

inline Float3 triangle_shading_normal(const EchoTriangle& tri, Float2 uv) // TriangleEntity.cs:187
{
	return normalized((1.0f - uv.x - uv.y) * f3(tri.normal0) + uv.x * f3(tri.normal1) + uv.y * f3(tri.normal2));
}

inline Float3 triangle_point(const EchoTriangle& tri, Float2 uv) // TriangleEntity.cs:265
{
	return f3(tri.vertex0) + uv.x * f3(tri.edge1) + uv.y * f3(tri.edge2);
}

// ---- Scenic/Geometries/SphereEntity.cs (PreparedSphere) ----
constexpr float kSphereDistanceThreshold = 6E-4f; // SphereEntity.cs:79

// SphereEntity.cs:88-126
inline float sphere_intersect(const EchoSphere& sphere, const Ray& ray, Float2& uv, bool findFar)
{
	float radius = sphere.radius;
	Float3 offset = ray.origin - f3(sphere.position);
	float radius2 = radius * radius;
	float center = -dot(offset, ray.direction);

	float extend2 = fma_f(center, center, radius2 - squared_magnitude(offset));
	if (extend2 < 0.0f) return kInfinity;

	float extend = sqrt0(extend2);
	float distance = center - extend;

	float threshold = findFar ? kSphereDistanceThreshold : 0.0f;

	if (distance < threshold) distance = center + extend;
	if (distance < threshold) return kInfinity;

	Float3 point = offset + ray.direction * distance;
	float sinP = clamp11(point.y / radius);
	float sinT = 0.0f;

	float smallRadius = fma_f(-point.y, point.y, radius2);
	if (smallRadius > 0.0f) sinT = point.x * sqrt_r0(smallRadius);
	if (point.z < 0.0f) sinT += 3.0f;

	uv = { sinT, sinP };
	return distance;
}

// SphereEntity.cs:129-148
inline bool sphere_occlude(const EchoSphere& sphere, const Ray& ray, float travel, bool findFar)
{
	float radius = sphere.radius;
	Float3 offset = ray.origin - f3(sphere.position);
	float center = -dot(offset, ray.direction);

	float squared = fma_f(radius, radius, -squared_magnitude(offset));
	float extend2 = fma_f(center, center, squared);
	if (extend2 < 0.0f) return false;

	float extend = sqrt0(extend2);
	float distance = center - extend;

	float threshold = findFar ? kSphereDistanceThreshold : 0.0f;

	if (distance < threshold) distance = center + extend;
	return distance >= threshold && distance < travel;
}

inline float sphere_area(const EchoSphere& sphere) { return 4.0f * kPi * sphere.radius * sphere.radius; } // SphereEntity.cs:72

// SphereEntity.cs:252-266
inline void sphere_theta_phi(Float2 uv, float& sinT, float& sinP, float& cosT)
{
	sinT = uv.x;
	sinP = uv.y;
	float sign = 1.0f;

	if (sinT > 1.5f)
	{
		sinT -= 3.0f;
		sign = -1.0f;
	}

	cosT = identity(sinT) * sign;
}

// SphereEntity.cs:229-234
inline Float3 sphere_normal(Float2 uv)
{
	float sinT, sinP, cosT;
	sphere_theta_phi(uv, sinT, sinP, cosT);
	float cosP = identity(sinP);
	return normalized(Float3{ sinT * cosP, sinP, cosT * cosP });
}

// ---- Aggregation/Bounds/BoxBound4.cs:64-112 ----
inline void box4_intersect(const EchoQbvhNode& node, const Ray& ray, float out[4])
{
	for (int lane = 0; lane < 4; lane++)
	{
		float length0 = (node.minX[lane] - ray.origin.x) * ray.directionR.x;
		float length1 = (node.maxX[lane] - ray.origin.x) * ray.directionR.x;

		float far = sse_max(length0, length1);
		float near = sse_min(length0, length1);

		length0 = (node.minY[lane] - ray.origin.y) * ray.directionR.y;
		length1 = (node.maxY[lane] - ray.origin.y) * ray.directionR.y;

		far = sse_min(far, sse_max(length0, length1));
		near = sse_max(near, sse_min(length0, length1));

		length0 = (node.minZ[lane] - ray.origin.z) * ray.directionR.z;
		length1 = (node.maxZ[lane] - ray.origin.z) * ray.directionR.z;

		far = sse_min(far, sse_max(length0, length1));
		near = sse_max(near, sse_min(length0, length1));

		far *= 1.00000024f; // BoxBound.FarMultiplier, BoxBound.cs:67

		out[lane] = ((far >= near) & (far >= 0.0f)) ? near : kInfinity;
	}
}

// ---- the prepared scene as it crosses the C ABI (Aggregation/Preparation/PreparedScene.cs, PreparedPack.cs) ----
struct Scene
{
	std::vector<EchoQbvhNode> nodes;
	uint32_t maxDepth = 0;
	std::vector<EchoTriangle> triangles;
	std::vector<EchoSphere> spheres;
	std::vector<EchoMaterial> materials;

	std::vector<EchoLightNode> lightNodes;
	std::vector<uint32_t> emitterTokens; // LightTree.map of every pack, back to back (EchoPack.emitterOffset / emitterCount)
	std::vector<uint64_t> emitterPaths;
	std::vector<std::unordered_map<uint32_t, uint64_t>> lightMaps; // LightTree.map per pack, LightTree.cs:41
	std::vector<EchoPointLight> pointLights;

	std::vector<EchoInfiniteLight> infiniteLights;
	float infiniteLightsThreshold = 0.0f; // PreparedScene.cs:38
	float infiniteLightsPdf = 0.0f;       // PreparedScene.cs:39

	// image textures (include/echo_b200.h EchoTexture): RGBA128 texels of every grid back to back, and the texture slots of
	// every material (empty: every slot is a constant)
	std::vector<EchoTexture> textures;
	std::vector<float> texels;
	std::vector<EchoMaterialTextures> materialTextures;

	struct Rgba
	{
		float v[4];
	};

	Rgba texel(const EchoTexture& texture, int x, int y) const // TextureGrid.this[Int2], ArrayGrid.cs
	{
		const float* p = texels.data() + ((size_t)texture.texelOffset + (size_t)y * texture.width + (size_t)x) * 4;
		return { { p[0], p[1], p[2], p[3] } };
	}

	static int repeat_int(int value, int length) // Scalars.Repeat(int, int), Scalars.cs:205-210
	{
		if ((0 <= value) & (value < length)) return value;
		int mod = value % length;
		return mod < 0 ? mod + length : mod;
	}

	static void wrap(const EchoTexture& texture, int& x, int& y) // IWrapper.cs:18-96 (the SSE and the scalar forms agree)
	{
		int width = (int)texture.width, height = (int)texture.height;

		if (texture.wrapper == ECHO_WRAPPER_CLAMP)
		{
			x = x < 0 ? 0 : (x > width - 1 ? width - 1 : x);
			y = y < 0 ? 0 : (y > height - 1 ? height - 1 : y);
		}
		else if (texture.wrapper == ECHO_WRAPPER_REPEAT)
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

	// TextureGrid.this[Float2] -> IFilter.Evaluate, IFilter.cs:17-68
	Rgba texture_sample(uint32_t index, Float2 uv) const
	{
		const EchoTexture& texture = textures[index];
		float sizeX = (float)texture.width, sizeY = (float)texture.height;

		if (texture.filter == ECHO_FILTER_POINT)
		{
			int x = (int)std::floor((double)(uv.x * sizeX)), y = (int)std::floor((double)(uv.y * sizeY)); // ToPosition: (uv * size).Floored
			wrap(texture, x, y);
			return texel(texture, x, y);
		}

		float scaledX = uv.x * sizeX, scaledY = uv.y * sizeY;
		int roundX = (int)std::nearbyint(scaledX), roundY = (int)std::nearbyint(scaledY); // cvtps2dq: round to nearest even
		int x0 = roundX - 1, x1 = roundX, y0 = roundY - 1, y1 = roundY;
		int minX = x0, minY = y0;

		int ax = x0, ay = y0, bx = x1, by = y0, cx = x0, cy = y1, dx = x1, dy = y1;
		wrap(texture, ax, ay);
		wrap(texture, bx, by);
		wrap(texture, cx, cy);
		wrap(texture, dx, dy);

		Rgba y0x0 = texel(texture, ax, ay), y0x1 = texel(texture, bx, by), y1x0 = texel(texture, cx, cy), y1x1 = texel(texture, dx, dy);

		float timeX = scaledX - 0.5f - (float)minX;
		float timeY = scaledY - 0.5f - (float)minY;

		auto lerp = [](float first, float second, float value) { return std::fma(value, second, std::fma(-value, first, first)); }; // Float4.Lerp on an FMA host, Float4.cs:396-406
		Rgba result;

		for (int c = 0; c < 4; c++)
		{
			float low = lerp(y0x0.v[c], y0x1.v[c], timeX);
			float high = lerp(y1x0.v[c], y1x1.v[c], timeX);
			result.v[c] = lerp(low, high, timeY);
		}

		return result;
	}

	std::vector<float> distributions; // DiscreteDistribution2D cdf values of the environment lights (echo_b200_scene_set_distributions)

	EchoMaterialTextures material_textures(uint32_t material) const
	{
		if (material < materialTextures.size()) return materialTextures[material];
		EchoMaterialTextures none = {};
		none.albedo = none.normal = none.roughness = none.paramA = none.paramB = ECHO_TEXTURE_NONE;
		return none;
	}

	EchoCamera camera = {};
	float boundRadius = 0.0f; // Accelerator.SphereBound.radius (Accelerator.cs:43-63), read by NormalDepthEvaluator

	// instancing: the arrays above hold every pack back to back (include/echo_b200.h EchoPack); empty = one pack
	std::vector<EchoPack> packs;
	std::vector<EchoInstance> instances;

	EchoPack pack_view(uint32_t pack) const
	{
		if (!packs.empty()) return packs[pack];
		EchoPack view = {};
		view.nodeCount = (uint32_t)nodes.size();
		view.maxDepth = maxDepth;
		view.triangleCount = (uint32_t)triangles.size();
		view.sphereCount = (uint32_t)spheres.size();
		view.lightNodeCount = (uint32_t)lightNodes.size();
		view.emitterCount = (uint32_t)emitterTokens.size();
		view.pointLightCount = (uint32_t)pointLights.size();
		return view;
	}

	void rebuild_light_maps()
	{
		size_t count = packs.empty() ? 1 : packs.size();
		lightMaps.assign(count, {});

		for (size_t p = 0; p < count; p++)
		{
			EchoPack view = pack_view((uint32_t)p);
			for (uint32_t i = 0; i < view.emitterCount && view.emitterOffset + i < emitterTokens.size(); i++)
				lightMaps[p][emitterTokens[view.emitterOffset + i]] = emitterPaths[view.emitterOffset + i];
		}
	}

	// ---- PreparedScene.FindLayer (PreparedScene.cs:255-277): walks the instance layers; forward / inverse are rows 0..2 of the
	// composed Float4x4s (the bottom row stays 0 0 0 1). Both products put the new instance on the LEFT, as the reference
	// does (:272-273) — for the inverse transform of nested placements that is not the geometric order, and it is kept.
	struct Layer
	{
		float forward[12], inverse[12];
		uint32_t pack;           // the pack the hierarchy ends in
		uint32_t materialOffset; // its placement's swatch (PreparedInstance.swatch); the scene's own for no layers
	};

	static void multiply_rows(const float* a, const float* b, float* out) // Float4x4 operator *, Float4x4.cs:352-358, affine operands
	{
		float result[12];

		for (int i = 0; i < 3; i++)
		{
			for (int j = 0; j < 4; j++)
			{
				float bottom = j == 3 ? 1.0f : 0.0f; // second.f3j
				result[i * 4 + j] = a[i * 4 + 0] * b[0 * 4 + j] + a[i * 4 + 1] * b[1 * 4 + j] + a[i * 4 + 2] * b[2 * 4 + j] + a[i * 4 + 3] * bottom;
			}
		}

		for (int k = 0; k < 12; k++) out[k] = result[k];
	}

	Layer find_layer(const Layers& layers) const
	{
		static const float identity[12] = { 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0 };
		Layer layer;
		for (int k = 0; k < 12; k++) layer.forward[k] = layer.inverse[k] = identity[k];
		layer.pack = 0;
		layer.materialOffset = pack_view(0).materialOffset;

		for (uint32_t k = 0; k < layers.count; k++)
		{
			const EchoInstance& instance = instances[pack_view(layer.pack).instanceOffset + token_index(layers.instances[k])];
			multiply_rows(instance.forward, layer.forward, layer.forward);
			multiply_rows(instance.inverse, layer.inverse, layer.inverse);
			layer.pack = instance.pack;
			layer.materialOffset = instance.materialOffset;
		}

		return layer;
	}

	// ---- Aggregation/Preparation/PreparedInstance.cs:105-111 ----
	static void transform_forward(const EchoInstance& instance, Ray& ray)
	{
		Float3 origin = multiply_point(instance.forward, ray.origin);
		Float3 direction = multiply_direction(instance.forward, ray.direction);
		ray = Ray(origin, direction * instance.inverseScale);
	}

	// ---- Aggregation/Preparation/GeometryCollection.cs:85-134 ----
	void geometry_trace(uint32_t token, TraceQuery& query, VisitCounters* counters, uint32_t pack = 0) const
	{
		EchoPack view = pack_view(pack);

		switch (token_type(token))
		{
			case ECHO_TOKEN_TYPE_TRIANGLE:
			{
				if (query.ignore == token && query.ignoreLayers == query.current) return; // query.ignore == query.current, :93
				if (counters) ++counters->triangles;

				Float2 uv = query.uv; // an unaccepted test leaves query.uv untouched (uv is a local there, :96-103)
				float distance = triangle_intersect(triangles[view.triangleOffset + token_index(token)], query.ray.origin, query.ray.direction, uv);
				if (distance >= query.distance) return;

				query.token = token;
				query.tokenLayers = query.current;
				query.distance = distance;
				query.uv = uv;
				break;
			}
			case ECHO_TOKEN_TYPE_SPHERE:
			{
				bool findFar = query.ignore == token && query.ignoreLayers == query.current;
				if (counters) ++counters->spheres;

				Float2 uv = query.uv;
				float distance = sphere_intersect(spheres[view.sphereOffset + token_index(token)], query.ray, uv, findFar);
				if (distance >= query.distance) return;

				query.token = token;
				query.tokenLayers = query.current;
				query.distance = distance;
				query.uv = uv;
				break;
			}
			case ECHO_TOKEN_TYPE_INSTANCE: // :123-131 + PreparedInstance.Trace, PreparedInstance.cs:47-61
			{
				const EchoInstance& instance = instances[view.instanceOffset + token_index(token)];
				query.current.push(token);

				Ray oldRay = query.ray;
				transform_forward(instance, query.ray);
				query.distance *= instance.forwardScale;

				accelerator_trace(query, counters, instance.pack);

				query.ray = oldRay;
				query.distance *= instance.inverseScale;
				query.current.pop();
				break;
			}
			default: break;
		}
	}

	// ---- GeometryCollection.cs:140-171 ----
	bool geometry_occlude(uint32_t token, OccludeQuery& query, VisitCounters* counters, uint32_t pack = 0) const
	{
		EchoPack view = pack_view(pack);

		switch (token_type(token))
		{
			case ECHO_TOKEN_TYPE_TRIANGLE:
			{
				if (query.ignore == token && query.ignoreLayers == query.current) return false;
				if (counters) ++counters->triangles;
				return triangle_occlude(triangles[view.triangleOffset + token_index(token)], query.ray.origin, query.ray.direction, query.travel);
			}
			case ECHO_TOKEN_TYPE_SPHERE:
			{
				bool findFar = query.ignore == token && query.ignoreLayers == query.current;
				if (counters) ++counters->spheres;
				return sphere_occlude(spheres[view.sphereOffset + token_index(token)], query.ray, query.travel, findFar);
			}
			case ECHO_TOKEN_TYPE_INSTANCE: // :160-168 + PreparedInstance.Occlude, PreparedInstance.cs:63-83
			{
				const EchoInstance& instance = instances[view.instanceOffset + token_index(token)];
				query.current.push(token);

				Ray oldRay = query.ray;
				float oldTravel = query.travel;
				transform_forward(instance, query.ray);
				query.travel *= instance.forwardScale;

				if (accelerator_occlude(query, counters, instance.pack)) return true;

				query.ray = oldRay;
				query.travel = oldTravel;
				query.current.pop();
				return false;
			}
			default: return false;
		}
	}

	// ---- Aggregation/Acceleration/QuadBoundingVolumeHierarchy.cs:123-219 ----
	void accelerator_trace(TraceQuery& query, VisitCounters* counters, uint32_t pack = 0) const
	{
		EchoPack view = pack_view(pack);
		const EchoQbvhNode* nodes = this->nodes.data() + view.nodeOffset;
		uint32_t maxDepth = view.maxDepth;
		int stackSize = (int)maxDepth * 3 + 1; // :34
		uint32_t* stack = (uint32_t*)alloca(sizeof(uint32_t) * stackSize); // stackalloc, :125-126
		float* hitsBase = (float*)alloca(sizeof(float) * stackSize);

		uint32_t* next = stack;
		float* hits = hitsBase;
		*next++ = ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_NODE, 0);
		*hits++ = 0.0f;

		bool orders[4] = { query.ray.directionR.x > 0, query.ray.directionR.y > 0, query.ray.directionR.z > 0, true };

		auto push = [&](const float* intersections, const EchoQbvhNode& node, int offset)
		{
			float hit = intersections[offset];
			if (hit >= query.distance) return;

			uint32_t token = node.token4[offset];

			// An empty child has bounds at +infinity and never passes the test above — unless query.distance has become NaN
			// (a NaN hit distance is accepted, GeometryCollection.cs:99, e.g. after an overflow on absurdly large coordinates).
			// The reference would then index nodes[] out of range (:211); here the slot is skipped, like on the device.
			if (token == ECHO_TOKEN_EMPTY) return;

			if (!token_is_geometry(token))
			{
				*next++ = token;
				*hits++ = hit;
			}
			else geometry_trace(token, query, counters, pack);
		};

		do
		{
			uint32_t index = token_index(*--next);
			if (*--hits >= query.distance) continue;

			const EchoQbvhNode& node = nodes[index];
			float intersections[4];
			box4_intersect(node, query.ray, intersections);
			if (counters) ++counters->nodes;

			if (orders[node.axisMajor])
			{
				if (orders[node.axisMinor1])
				{
					push(intersections, node, 3);
					push(intersections, node, 2);
				}
				else
				{
					push(intersections, node, 2);
					push(intersections, node, 3);
				}

				if (orders[node.axisMinor0])
				{
					push(intersections, node, 1);
					push(intersections, node, 0);
				}
				else
				{
					push(intersections, node, 0);
					push(intersections, node, 1);
				}
			}
			else
			{
				if (orders[node.axisMinor0])
				{
					push(intersections, node, 1);
					push(intersections, node, 0);
				}
				else
				{
					push(intersections, node, 0);
					push(intersections, node, 1);
				}

				if (orders[node.axisMinor1])
				{
					push(intersections, node, 3);
					push(intersections, node, 2);
				}
				else
				{
					push(intersections, node, 2);
					push(intersections, node, 3);
				}
			}
		}
		while (next != stack);
	}

	// ---- QuadBoundingVolumeHierarchy.cs:223-315 ----
	bool accelerator_occlude(OccludeQuery& query, VisitCounters* counters, uint32_t pack = 0) const
	{
		EchoPack view = pack_view(pack);
		const EchoQbvhNode* nodes = this->nodes.data() + view.nodeOffset;
		uint32_t maxDepth = view.maxDepth;
		int stackSize = (int)maxDepth * 3 + 1;
		uint32_t* stack = (uint32_t*)alloca(sizeof(uint32_t) * stackSize); // stackalloc, :227

		uint32_t* next = stack;
		*next++ = ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_NODE, 0);

		bool orders[4] = { query.ray.directionR.x > 0, query.ray.directionR.y > 0, query.ray.directionR.z > 0, true };

		auto push = [&](const float* intersections, const EchoQbvhNode& node, int offset) -> bool
		{
			float hit = intersections[offset];
			if (hit >= query.travel) return false;

			uint32_t token = node.token4[offset];
			if (token == ECHO_TOKEN_EMPTY) return false; // only reachable with a NaN travel, see accelerator_trace
			if (token_is_geometry(token)) return geometry_occlude(token, query, counters, pack);

			*next++ = token;
			return false;
		};

		do
		{
			uint32_t index = token_index(*--next);
			const EchoQbvhNode& node = nodes[index];
			float intersections[4];
			box4_intersect(node, query.ray, intersections);
			if (counters) ++counters->nodes;

			if (orders[node.axisMajor])
			{
				if (orders[node.axisMinor1])
				{
					if (push(intersections, node, 3)) return true;
					if (push(intersections, node, 2)) return true;
				}
				else
				{
					if (push(intersections, node, 2)) return true;
					if (push(intersections, node, 3)) return true;
				}

				if (orders[node.axisMinor0])
				{
					if (push(intersections, node, 1)) return true;
					if (push(intersections, node, 0)) return true;
				}
				else
				{
					if (push(intersections, node, 0)) return true;
					if (push(intersections, node, 1)) return true;
				}
			}
			else
			{
				if (orders[node.axisMinor0])
				{
					if (push(intersections, node, 1)) return true;
					if (push(intersections, node, 0)) return true;
				}
				else
				{
					if (push(intersections, node, 0)) return true;
					if (push(intersections, node, 1)) return true;
				}

				if (orders[node.axisMinor1])
				{
					if (push(intersections, node, 3)) return true;
					if (push(intersections, node, 2)) return true;
				}
				else
				{
					if (push(intersections, node, 2)) return true;
					if (push(intersections, node, 3)) return true;
				}
			}
		}
		while (next != stack);

		return false;
	}

	// ---- Aggregation/Preparation/PreparedScene.cs:66-86 ----
	bool trace(TraceQuery& query, VisitCounters* counters = nullptr) const
	{
		if (!positive(query.distance)) return false;
		float original = query.distance;
		accelerator_trace(query, counters);
		return query.distance < original;
	}

	bool occlude(const OccludeQuery& original, VisitCounters* counters = nullptr) const
	{
		if (!positive(original.travel)) return false;
		OccludeQuery query = original; // instances rewrite ray and travel while they are entered
		return accelerator_occlude(query, counters);
	}

	// brute force over every primitive in token order: the cross-check SURVEY.md §4 asks for (the reference's
	// LinearAccelerator does the same loop, Aggregation/Acceleration/LinearAccelerator.cs)
	bool trace_linear(TraceQuery& query) const
	{
		if (!positive(query.distance)) return false;
		float original = query.distance;
		EchoPack view = pack_view(0);
		for (uint32_t i = 0; i < view.triangleCount; i++) geometry_trace(ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_TRIANGLE, i), query, nullptr);
		for (uint32_t i = 0; i < view.sphereCount; i++) geometry_trace(ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_SPHERE, i), query, nullptr);
		for (uint32_t i = 0; i < view.instanceCount; i++) geometry_trace(ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_INSTANCE, i), query, nullptr);
		return query.distance < original;
	}

	bool occlude_linear(const OccludeQuery& original) const
	{
		if (!positive(original.travel)) return false;
		OccludeQuery query = original;
		EchoPack view = pack_view(0);
		for (uint32_t i = 0; i < view.triangleCount; i++)
			if (geometry_occlude(ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_TRIANGLE, i), query, nullptr)) return true;
		for (uint32_t i = 0; i < view.sphereCount; i++)
			if (geometry_occlude(ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_SPHERE, i), query, nullptr)) return true;
		for (uint32_t i = 0; i < view.instanceCount; i++)
			if (geometry_occlude(ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_INSTANCE, i), query, nullptr)) return true;
		return false;
	}

	uint32_t geometry_material(uint32_t token, uint32_t pack = 0) const // GeometryCollection.cs:236-246, index inside the swatch
	{
		EchoPack view = pack_view(pack);
		return token_type(token) == ECHO_TOKEN_TYPE_TRIANGLE ? triangles[view.triangleOffset + token_index(token)].material
		                                                     : spheres[view.sphereOffset + token_index(token)].material;
	}

	// ---- Aggregation/Bounds/LightBound.cs:30-80 ----
	static float clamp_subtract_cos(float sin0, float cos0, float sin1, float cos1) { return cos0 > cos1 ? 1.0f : cos0 * cos1 + sin0 * sin1; }
	static float clamp_subtract_sin(float sin0, float cos0, float sin1, float cos1) { return cos0 > cos1 ? 0.0f : sin0 * cos1 - cos0 * sin1; }

	static float light_importance(const EchoLightNode& bound, const GeometryPoint& origin)
	{
		++lightNodeVisits;
		Float3 boxMin = f3(bound.boxMin), boxMax = f3(bound.boxMax);
		Float3 center = (boxMax + boxMin) / 2.0f; // BoxBound.cs:69
		Float3 incident = origin.position - center;

		float length2 = squared_magnitude(incident);

		if (almost_zero(length2)) incident = { 0.0f, 0.0f, 0.0f };
		else incident = incident * sqrt_r0(length2);

		float cosAxis = dot(f3(bound.coneAxis), incident);
		float sinAxis = identity(cosAxis);

		float cosOffset = bound.cosOffset;
		float sinOffset = identity(cosOffset);

		// FindSubtendedAngles, LightBound.cs:62-77
		float sinRadius, cosRadius;
		float radius2 = squared_magnitude(boxMax - boxMin) / 4.0f;

		if (length2 < radius2)
		{
			sinRadius = 0.0f;
			cosRadius = -1.0f;
		}
		else
		{
			float sinRadius2 = radius2 / length2;
			sinRadius = sqrt0(sinRadius2);
			cosRadius = sqrt0(1.0f - sinRadius2);
		}

		float cosRemain = clamp_subtract_cos(sinAxis, cosAxis, sinOffset, cosOffset);
		float sinRemain = clamp_subtract_sin(sinAxis, cosAxis, sinOffset, cosOffset);
		float cosFinal = clamp_subtract_cos(sinRemain, cosRemain, sinRadius, cosRadius);
		if (cosFinal <= bound.cosExtend) return 0.0f;

		float cosIncident = fabs_bits(dot(origin.normal, incident));
		float sinIncident = identity(cosIncident);
		float cosReflect = clamp_subtract_cos(sinIncident, cosIncident, sinRadius, cosRadius);

		length2 = math_max(length2, magnitude(boxMax - boxMin) / 2.0f);
		return max0(bound.power / length2 * cosFinal * cosReflect);
	}

	// ---- Aggregation/Selection/LightTree.cs:115-134 (recursion unrolled into a loop; same arithmetic order) ----
	// returns the picked token and its probability mass; mass == 0 means Probable.Impossible
	uint32_t light_tree_pick(const GeometryPoint& origin, float& sample, float& outPdf, uint32_t pack = 0) const
	{
		outPdf = 0.0f;
		EchoPack view = pack_view(pack);
		if (view.lightNodeCount == 0) return ECHO_TOKEN_EMPTY;
		const EchoLightNode* lightNodes = this->lightNodes.data() + view.lightNodeOffset;

		uint32_t index = 0;
		float pdf = 1.0f;

		while (true)
		{
			const EchoLightNode& node = lightNodes[index];

			if (node.child0 == ECHO_TOKEN_EMPTY)
			{
				outPdf = pdf;
				return node.child1;
			}

			float importance0 = light_importance(lightNodes[node.child0], origin);
			float importance1 = light_importance(lightNodes[node.child1], origin);

			if (!positive(importance0) && !positive(importance1)) return ECHO_TOKEN_EMPTY;

			float split = importance0 / (importance0 + importance1);

			if (sample < split)
			{
				sample = sample_stretch(sample, 0.0f, split);
				index = node.child0;
				pdf = pdf * split;
			}
			else
			{
				sample = sample_stretch(sample, split, 1.0f);
				index = node.child1;
				pdf = pdf * (1.0f - split);
			}
		}
	}

	// ---- LightTree.cs:53-57,136-154: split0 * (split1 * (... * 1)) — the recursion multiplies from the leaf upward ----
	float light_tree_mass(uint32_t token, const GeometryPoint& origin, uint32_t pack = 0) const
	{
		if (pack >= lightMaps.size()) return 0.0f;
		auto found = lightMaps[pack].find(token);
		if (found == lightMaps[pack].end()) return 0.0f;
		return light_tree_mass_recursive(origin, 0, found->second, this->lightNodes.data() + pack_view(pack).lightNodeOffset);
	}

	float light_tree_mass_recursive(const GeometryPoint& origin, uint32_t index, uint64_t branches, const EchoLightNode* lightNodes) const
	{
		const EchoLightNode& node = lightNodes[index];
		if (node.child0 == ECHO_TOKEN_EMPTY) return 1.0f;

		float importance0 = light_importance(lightNodes[node.child0], origin);
		float importance1 = light_importance(lightNodes[node.child1], origin);
		float split = importance0 / (importance0 + importance1);

		if ((branches & 1) == 0) return split * light_tree_mass_recursive(origin, node.child0, branches >> 1, lightNodes);
		return (1.0f - split) * light_tree_mass_recursive(origin, node.child1, branches >> 1, lightNodes);
	}

	// ---- PreparedScene.cs:113-150 ----
	uint32_t pick(const GeometryPoint& origin, float sample, float& outPdf, Layers* outLayers = nullptr) const
	{
		Layers layers;
		if (outLayers) *outLayers = layers;

		if (sample < infiniteLightsThreshold)
		{
			sample = sample_stretch(sample, 0.0f, infiniteLightsThreshold);
			int index = sample_range(sample, (int)infiniteLights.size());
			outPdf = infiniteLightsPdf;
			return ECHO_LIGHT_TOKEN_MAKE(infiniteLights[index].isDelta ? ECHO_LIGHT_TYPE_INFINITE_DELTA : ECHO_LIGHT_TYPE_INFINITE, index); // :120
		}

		sample = sample_stretch(sample, infiniteLightsThreshold, 1.0f);
		float pdf = 1.0f - infiniteLightsThreshold;
		uint32_t pack = 0;

		while (true) // :132-147: the same `origin` serves every layer (it is not moved into the placement's space)
		{
			float tokenPdf;
			uint32_t token = light_tree_pick(origin, sample, tokenPdf, pack);

			if (almost_zero(tokenPdf))
			{
				outPdf = 0.0f;
				return ECHO_TOKEN_EMPTY;
			}

			pdf *= tokenPdf;

			if (token_type(token) != ECHO_TOKEN_TYPE_INSTANCE || layers.count >= ECHO_MAX_INSTANCE_LAYERS)
			{
				outPdf = pdf;
				if (outLayers) *outLayers = layers;
				return token;
			}

			const EchoInstance& instance = instances[pack_view(pack).instanceOffset + token_index(token)];
			layers.push(token);
			pack = instance.pack;
		}
	}

	// ---- PreparedScene.cs:158-179 ----
	float probability_mass(uint32_t light, const GeometryPoint& origin, const Layers& layers = Layers()) const
	{
		if (token_is_infinite_light(light)) return infiniteLightsPdf;
		float pdf = 1.0f - infiniteLightsThreshold;
		uint32_t pack = 0;

		for (uint32_t k = 0; k < layers.count; k++)
		{
			pdf *= light_tree_mass(layers.instances[k], origin, pack);
			pack = instances[pack_view(pack).instanceOffset + token_index(layers.instances[k])].pack;
			if (almost_zero(pdf)) return 0.0f;
		}

		return pdf * light_tree_mass(light, origin, pack);
	}
};

} // namespace oracle
