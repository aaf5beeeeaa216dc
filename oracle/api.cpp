// oracle/api.cpp — TEST INFRASTRUCTURE ONLY. C entry points of liboracle.so, the CPU restatement of Echo's hot path.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library;
// the product (echorenderer_b200 + libecho_b200.so) never does. Parity status: the reference is C#/.NET and cannot
// run in this environment, and its own tests hold no vectors for traversal / triangle intersection / the light tree /
// the integrator, so those parts are "parity unpinned" (see oracle/README.md); the BxDF, sphere, FastMath, Kahan and
// orthonormal-transform parts are pinned by the reference's unit-test tables, restated in tests/test_oracle_kats.py.
#include <atomic>
#include <thread>

#include "evaluation.hpp"
#include "oracle.h"

using namespace oracle;

struct OracleScene
{
	Scene scene;
};

template<class F>
static void parallel_chunks(uint64_t count, uint64_t grain, int threads, F&& body)
{
	if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
	if (threads < 1) threads = 1;

	std::atomic<uint64_t> nextChunk{ 0 };
	uint64_t chunks = (count + grain - 1) / grain;

	auto work = [&](int worker)
	{
		// mirrors Operation.Execute's atomic procedure counter (Common/Compute/Operation.cs:164-177)
		while (true)
		{
			uint64_t chunk = nextChunk.fetch_add(1);
			if (chunk >= chunks) break;
			uint64_t begin = chunk * grain;
			uint64_t end = begin + grain < count ? begin + grain : count;
			body(worker, begin, end);
		}
	};

	if (threads == 1)
	{
		work(0);
		return;
	}

	std::vector<std::thread> pool;
	for (int i = 0; i < threads; i++) pool.emplace_back(work, i);
	for (auto& thread : pool) thread.join();
}

extern "C"
{

OracleScene* oracle_scene_create(void) { return new OracleScene(); }
void oracle_scene_destroy(OracleScene* s) { delete s; }

void oracle_scene_set_qbvh(OracleScene* s, const EchoQbvhNode* nodes, uint32_t count, uint32_t maxDepth)
{
	s->scene.nodes.assign(nodes, nodes + count);
	s->scene.maxDepth = maxDepth;
}

void oracle_scene_set_triangles(OracleScene* s, const EchoTriangle* triangles, uint32_t count) { s->scene.triangles.assign(triangles, triangles + count); }
void oracle_scene_set_spheres(OracleScene* s, const EchoSphere* spheres, uint32_t count) { s->scene.spheres.assign(spheres, spheres + count); }
void oracle_scene_set_materials(OracleScene* s, const EchoMaterial* materials, uint32_t count) { s->scene.materials.assign(materials, materials + count); }

void oracle_scene_set_light_tree(OracleScene* s, const EchoLightNode* nodes, uint32_t nodeCount,
                                 const uint32_t* tokens, const uint64_t* bitpaths, uint32_t emitterCount,
                                 const EchoPointLight* points, uint32_t pointCount)
{
	s->scene.lightNodes.assign(nodes, nodes + nodeCount);
	s->scene.emitterTokens.assign(tokens, tokens + emitterCount);
	s->scene.emitterPaths.assign(bitpaths, bitpaths + emitterCount);
	s->scene.pointLights.assign(points, points + pointCount);
	s->scene.rebuild_light_maps();
}

void oracle_scene_set_textures(OracleScene* s, const EchoTexture* textures, uint32_t textureCount, const float* texels, uint64_t texelCount,
                               const EchoMaterialTextures* materialTextures, uint32_t materialCount)
{
	s->scene.textures.assign(textures, textures + textureCount);
	s->scene.texels.assign(texels, texels + texelCount * 4);
	s->scene.materialTextures.assign(materialTextures, materialTextures + (textureCount ? materialCount : 0));
}

// KAT hook: TextureGrid.this[Float2] for n texture coordinates
void oracle_texture_sample(const OracleScene* s, uint32_t texture, const float* uv, uint64_t n, float* outRGBA)
{
	for (uint64_t i = 0; i < n; i++)
	{
		Scene::Rgba value = s->scene.texture_sample(texture, Float2{ uv[i * 2], uv[i * 2 + 1] });
		for (int c = 0; c < 4; c++) outRGBA[i * 4 + c] = value.v[c];
	}
}

float oracle_atan2(float y, float x) { return atan2_det(y, x); }
float oracle_asin(float x) { return asin_det(x); }
float oracle_acos(float x) { return acos_det(x); }

void oracle_scene_set_distributions(OracleScene* s, const float* values, uint64_t count) { s->scene.distributions.assign(values, values + count); }

// KAT hook for infinite lights: Sample(sample2) -> out[0..2] radiance, out[3] pdf, out[4..6] incident; then Evaluate(direction3)
// -> out[7..9] and ProbabilityDensity(direction3) -> out[10]
void oracle_infinite_light(const OracleScene* s, uint32_t index, const float* sample2, const float* direction3, float* out11)
{
	const EchoInfiniteLight& light = s->scene.infiniteLights[index];
	Float3 incident;
	float travel;
	ProbableRGB sampled = infinite_sample(s->scene, light, Float2{ sample1d(sample2[0]), sample1d(sample2[1]) }, incident, travel);
	out11[0] = sampled.content.r; out11[1] = sampled.content.g; out11[2] = sampled.content.b; out11[3] = sampled.pdf;
	out11[4] = incident.x; out11[5] = incident.y; out11[6] = incident.z;
	RGB value = infinite_evaluate(s->scene, light, f3(direction3));
	out11[7] = value.r; out11[8] = value.g; out11[9] = value.b;
	out11[10] = infinite_pdf(s->scene, light, f3(direction3));
}

// KAT hooks for src/Echo.UnitTests/Evaluation/DiscreteDistribution1Tests.cs over a cdf built by the host mirror of the
// constructor: Sample -> out[0] value, out[1] pdf; ProbabilityDensity(out[0]) -> out[2]; Pick -> out[3] index, out[4] pdf
// (Pick / ProbabilityMass are FindIndex + GetBounds, DiscreteDistribution1D.cs:79-101)
void oracle_distribution1d(const float* cdf, int count, float sample, float* out5)
{
	Distribution1D distribution{ cdf, count };
	float pdf;
	out5[0] = distribution.sample(sample1d(sample), pdf);
	out5[1] = pdf;
	out5[2] = distribution.probability_density(out5[0]);
	int index = distribution.find_index(sample1d(sample));
	float lower, upper;
	distribution.bounds(index, lower, upper);
	out5[3] = (float)index;
	out5[4] = upper - lower;
}

// KAT hooks for src/Echo.UnitTests/Textures/DirectionalTextureTests.cs:43-58 (CylindricalTexture.ToUV / ToDirection)
void oracle_cylindrical_to_uv(const float* direction3, float* uv2)
{
	Float2 uv = cylindrical_to_uv(f3(direction3));
	uv2[0] = uv.x; uv2[1] = uv.y;
}

void oracle_cylindrical_to_direction(const float* uv2, float* direction3)
{
	Float3 direction = cylindrical_to_direction(Float2{ uv2[0], uv2[1] });
	direction3[0] = direction.x; direction3[1] = direction.y; direction3[2] = direction.z;
}

void oracle_scene_set_bound_radius(OracleScene* s, float radius) { s->scene.boundRadius = radius; }

void oracle_scene_set_infinite(OracleScene* s, const EchoInfiniteLight* lights, uint32_t count, float threshold, float pdf)
{
	s->scene.infiniteLights.assign(lights, lights + count);
	s->scene.infiniteLightsThreshold = threshold;
	s->scene.infiniteLightsPdf = pdf;
}

void oracle_scene_set_camera(OracleScene* s, const EchoCamera* camera) { s->scene.camera = *camera; }

static TraceQuery make_trace_query(const EchoRay& in)
{
	TraceQuery query;
	query.ray = Ray(f3(in.origin), f3(in.direction));
	query.ignore = in.ignore;
	query.distance = in.distance;
	return query;
}

static void store_hit(const TraceQuery& query, bool hit, const EchoRay& in, EchoHit& out)
{
	out.token = hit ? query.token : ECHO_TOKEN_EMPTY;
	out.distance = hit ? query.distance : in.distance;
	out.uv[0] = hit ? query.uv.x : 0.0f;
	out.uv[1] = hit ? query.uv.y : 0.0f;
}

// counters[3] (optional): total node / triangle / sphere visits over the batch
void oracle_trace_batch(const OracleScene* s, const EchoRay* rays, uint64_t n, EchoHit* hits, uint64_t* counters, int threads)
{
	std::atomic<uint64_t> totals[3] = { {0}, {0}, {0} };

	parallel_chunks(n, 4096, threads, [&](int, uint64_t begin, uint64_t end)
	{
		VisitCounters local;

		for (uint64_t i = begin; i < end; i++)
		{
			TraceQuery query = make_trace_query(rays[i]);
			bool hit = s->scene.trace(query, counters ? &local : nullptr);
			store_hit(query, hit, rays[i], hits[i]);
		}

		totals[0] += local.nodes;
		totals[1] += local.triangles;
		totals[2] += local.spheres;
	});

	if (counters) for (int i = 0; i < 3; i++) counters[i] = totals[i];
}

void oracle_occlude_batch(const OracleScene* s, const EchoRay* rays, uint64_t n, uint8_t* occluded, uint64_t* counters, int threads)
{
	std::atomic<uint64_t> totals[3] = { {0}, {0}, {0} };

	parallel_chunks(n, 4096, threads, [&](int, uint64_t begin, uint64_t end)
	{
		VisitCounters local;

		for (uint64_t i = begin; i < end; i++)
		{
			OccludeQuery query;
			query.ray = Ray(f3(rays[i].origin), f3(rays[i].direction));
			query.ignore = rays[i].ignore;
			query.travel = rays[i].distance;
			occluded[i] = s->scene.occlude(query, counters ? &local : nullptr) ? 1 : 0;
		}

		totals[0] += local.nodes;
		totals[1] += local.triangles;
		totals[2] += local.spheres;
	});

	if (counters) for (int i = 0; i < 3; i++) counters[i] = totals[i];
}

void oracle_scene_set_packs(OracleScene* s, const EchoPack* packs, uint32_t packCount, const EchoInstance* instances, uint32_t instanceCount)
{
	s->scene.packs.assign(packs, packs + packCount);
	s->scene.instances.assign(instances, instances + instanceCount);
	s->scene.rebuild_light_maps();
}

static Layers make_layers(const EchoTokenHierarchy* in)
{
	Layers layers;
	if (!in) return layers;
	layers.count = in->instanceCount;
	for (uint32_t k = 0; k < layers.count && k < ECHO_MAX_INSTANCE_LAYERS; k++) layers.instances[k] = in->instances[k];
	return layers;
}

// the batch queries with full TokenHierarchy in and out (include/echo_b200.h echo_b200_trace_batch_hierarchy); linear != 0
// replaces the root accelerator by the brute-force loop over the root pack's geometry
void oracle_trace_batch_hierarchy(const OracleScene* s, const EchoRay* rays, const EchoTokenHierarchy* ignore, uint64_t n,
                                  EchoHit* hits, EchoTokenHierarchy* hitLayers, int linear, int threads)
{
	parallel_chunks(n, linear ? 16 : 4096, threads, [&](int, uint64_t begin, uint64_t end)
	{
		for (uint64_t i = begin; i < end; i++)
		{
			TraceQuery query = make_trace_query(rays[i]);
			query.ignoreLayers = make_layers(ignore ? ignore + i : nullptr);
			bool hit = linear ? s->scene.trace_linear(query) : s->scene.trace(query, nullptr);
			store_hit(query, hit, rays[i], hits[i]);

			if (hitLayers)
			{
				EchoTokenHierarchy out = {};
				if (hit)
				{
					out.instanceCount = query.tokenLayers.count;
					for (uint32_t k = 0; k < out.instanceCount; k++) out.instances[k] = query.tokenLayers.instances[k];
				}
				hitLayers[i] = out;
			}
		}
	});
}

void oracle_occlude_batch_hierarchy(const OracleScene* s, const EchoRay* rays, const EchoTokenHierarchy* ignore, uint64_t n,
                                    uint8_t* occluded, int linear, int threads)
{
	parallel_chunks(n, linear ? 16 : 4096, threads, [&](int, uint64_t begin, uint64_t end)
	{
		for (uint64_t i = begin; i < end; i++)
		{
			OccludeQuery query;
			query.ray = Ray(f3(rays[i].origin), f3(rays[i].direction));
			query.ignore = rays[i].ignore;
			query.travel = rays[i].distance;
			query.ignoreLayers = make_layers(ignore ? ignore + i : nullptr);
			occluded[i] = (linear ? s->scene.occlude_linear(query) : s->scene.occlude(query, nullptr)) ? 1 : 0;
		}
	});
}

void oracle_trace_linear_batch(const OracleScene* s, const EchoRay* rays, uint64_t n, EchoHit* hits, int threads)
{
	parallel_chunks(n, 16, threads, [&](int, uint64_t begin, uint64_t end)
	{
		for (uint64_t i = begin; i < end; i++)
		{
			TraceQuery query = make_trace_query(rays[i]);
			bool hit = s->scene.trace_linear(query);
			store_hit(query, hit, rays[i], hits[i]);
		}
	});
}

void oracle_occlude_linear_batch(const OracleScene* s, const EchoRay* rays, uint64_t n, uint8_t* occluded, int threads)
{
	parallel_chunks(n, 16, threads, [&](int, uint64_t begin, uint64_t end)
	{
		for (uint64_t i = begin; i < end; i++)
		{
			OccludeQuery query;
			query.ray = Ray(f3(rays[i].origin), f3(rays[i].direction));
			query.ignore = rays[i].ignore;
			query.travel = rays[i].distance;
			occluded[i] = s->scene.occlude_linear(query) ? 1 : 0;
		}
	});
}

static void merge_stats(EchoStats* out, const EvaluatorStats& stats, uint64_t samples, uint64_t rejected, uint64_t pixels)
{
	if (!out) return;
	*out = EchoStats{};
	out->sampleEvaluated = samples;
	out->sampleRejected = rejected;
	out->pixelEvaluated = pixels;
	out->bounceCreated = stats.bounceCreated;
	out->bounceSpecular = stats.bounceSpecular;
	out->bounceMis = stats.bounceMis;
	out->lightSampled = stats.lightSampled;
	out->lightOcclusionChecked = stats.lightOcclusionChecked;
	out->lightOcclusionPassed = stats.lightOcclusionPassed;
	out->lightEvaluatedInfinite = stats.lightEvaluatedInfinite;
	out->traceQueries = stats.traceQueries;
	out->occludeQueries = stats.occludeQueries;
	out->nodeVisits = stats.visits.nodes;
	out->triangleVisits = stats.visits.triangles;
	out->sphereVisits = stats.visits.spheres;
	out->lightNodeVisits = stats.lightNodeVisits;
}

// EvaluationOperation.Execute over a list of tiles (Processes/Evaluation/EvaluationOperation.cs:83-148); one tile per
// procedure, tiles claimed by worker threads. Output layout is echo_b200_render_tiles's.
void oracle_render_tiles(const OracleScene* s, const EchoRenderParams* params, const int32_t* tileXY, uint32_t tileCount,
                         float* outRGBA, EchoStats* outStats, int threads)
{
	int tileSize = params->tileSize;
	std::atomic<uint64_t> samples{ 0 }, rejected{ 0 }, pixels{ 0 };

	int workerCount = threads <= 0 ? (int)std::thread::hardware_concurrency() : threads;
	if (workerCount < 1) workerCount = 1;
	std::vector<EvaluatorStats> perWorker(workerCount);

	parallel_chunks(tileCount, 1, workerCount, [&](int worker, uint64_t begin, uint64_t end)
	{
		for (uint64_t t = begin; t < end; t++)
		{
			int minX = tileXY[t * 2] * tileSize, minY = tileXY[t * 2 + 1] * tileSize;
			float* tile = outRGBA + t * (uint64_t)tileSize * tileSize * 4;
			uint64_t localSamples = 0, localRejected = 0, localPixels = 0;

			for (int y = 0; y < tileSize; y++)
			for (int x = 0; x < tileSize; x++)
			{
				float* pixel = tile + ((uint64_t)y * tileSize + x) * 4;
				int px = minX + x, py = minY + y;

				if (px >= params->width || py >= params->height)
				{
					pixel[0] = pixel[1] = pixel[2] = pixel[3] = 0.0f;
					continue;
				}

				Float4 value = evaluate_pixel(s->scene, *params, px, py, perWorker[worker], localSamples, localRejected);
				for (int c = 0; c < 4; c++) pixel[c] = value.v[c];
				++localPixels;
			}

			samples += localSamples;
			rejected += localRejected;
			pixels += localPixels;
		}
	});

	EvaluatorStats total;
	for (const auto& stats : perWorker) total.add(stats);
	merge_stats(outStats, total, samples, rejected, pixels);
}

// one Evaluator.Evaluate per listed (pixel, sample index): the sample-exact view used to compare device paths
void oracle_evaluate_samples4(const OracleScene* s, const EchoRenderParams* params, const int32_t* pixelXY, const uint32_t* sampleIndex,
                              uint64_t n, float* out, int channels, int threads);

void oracle_evaluate_samples(const OracleScene* s, const EchoRenderParams* params, const int32_t* pixelXY, const uint32_t* sampleIndex,
                             uint64_t n, float* outRGB, int threads)
{
	oracle_evaluate_samples4(s, params, pixelXY, sampleIndex, n, outRGB, 3, threads);
}

// the same with `channels` floats per sample (4: the W lane of the auxiliary evaluators, e.g. NormalDepth128's depth)
void oracle_evaluate_samples4(const OracleScene* s, const EchoRenderParams* params, const int32_t* pixelXY, const uint32_t* sampleIndex,
                              uint64_t n, float* out, int channels, int threads)
{
	parallel_chunks(n, 256, threads, [&](int, uint64_t begin, uint64_t end)
	{
		CameraSpawner spawner(params->width, params->height);
		EvaluatorStats stats;

		for (uint64_t i = begin; i < end; i++)
		{
			int px = pixelXY[i * 2], py = pixelXY[i * 2 + 1];
			uint32_t pixel = (uint32_t)py * (uint32_t)params->width + (uint32_t)px;
			SampleStream distribution{ sample_key(params->seed, pixel, sampleIndex[i]) };

			Float2 shift = distribution.next2d();
			Float2 lens = distribution.next2d();
			Ray ray = camera_spawn_ray(s->scene.camera, spawner, px, py, shift, lens);
			Float4 value = evaluate_sample(s->scene, *params, ray, distribution, stats);
			for (int c = 0; c < channels; c++) out[i * channels + c] = value.v[c];
		}
	});
}

void oracle_spawn_rays(const OracleScene* s, const EchoRenderParams* params, const int32_t* pixelXY, const uint32_t* sampleIndex, uint64_t n, EchoRay* out)
{
	CameraSpawner spawner(params->width, params->height);

	for (uint64_t i = 0; i < n; i++)
	{
		int px = pixelXY[i * 2], py = pixelXY[i * 2 + 1];
		uint32_t pixel = (uint32_t)py * (uint32_t)params->width + (uint32_t)px;
		SampleStream distribution{ sample_key(params->seed, pixel, sampleIndex[i]) };
		Float2 shift = distribution.next2d();
		Float2 lens = distribution.next2d();
		Ray ray = camera_spawn_ray(s->scene.camera, spawner, px, py, shift, lens);
		out[i] = EchoRay{ { ray.origin.x, ray.origin.y, ray.origin.z }, { ray.direction.x, ray.direction.y, ray.direction.z }, kInfinity, ECHO_TOKEN_EMPTY };
	}
}

// ---------------- known-answer hooks for the reference's own unit-test tables ----------------

// FastMath (src/Echo.UnitTests/Common/FastMathTests.cs)
float oracle_fastmath(int32_t op, float a, float b, float c)
{
	switch (op)
	{
		case 0: return max0(a);
		case 1: return clamp01(a);
		case 2: return clamp11(a);
		case 3: return clamp_epsilon(a);
		case 4: return fabs_bits(a);
		case 5: return sqrt0(a);
		case 6: return sqrt_r0(a);
		case 7: return one_minus2(a);
		case 8: return identity(a);
		case 9: return fma_f(a, b, c);
		case 10: return positive(a) ? 1.0f : 0.0f;
		case 11: return almost_zero(a) ? 1.0f : 0.0f;
		case 12: return sse_min(a, b);
		case 13: return sse_max(a, b);
		default: return NAN;
	}
}

void oracle_sincos(float radians, float* sin, float* cos) { sincos_det(radians, *sin, *cos); }

// Summation (src/Echo.UnitTests/Common/SummationTests.cs): Kahan sum of n Float4 lanes-equal values
float oracle_kahan_sum(const float* values, uint64_t n)
{
	Summation sum;
	for (uint64_t i = 0; i < n; i++) sum = sum.add(Float4{ { values[i], values[i], values[i], values[i] } });
	return sum.total.v[0];
}

// OrthonormalTransform (src/Echo.UnitTests/Common/OrthonormalTransformTests.cs): out = axisX, axisY, axisZ, forward(d), inverse(d)
void oracle_orthonormal(const float* axisZ, const float* direction, float* out15)
{
	OrthonormalTransform transform(f3(axisZ));
	Float3 forward = transform.apply_forward(f3(direction));
	Float3 inverse = transform.apply_inverse(f3(direction));
	Float3 all[5] = { transform.axisX, transform.axisY, transform.axisZ, forward, inverse };
	for (int i = 0; i < 5; i++) { out15[i * 3] = all[i].x; out15[i * 3 + 1] = all[i].y; out15[i * 3 + 2] = all[i].z; }
}

// PreparedSphere (src/Echo.UnitTests/Scenic/PreparedSphereTests.cs)
float oracle_sphere_intersect(const EchoSphere* sphere, const float* origin, const float* direction, int32_t findFar, float* uv)
{
	Float2 result = { 0.0f, 0.0f };
	float distance = sphere_intersect(*sphere, Ray(f3(origin), f3(direction)), result, findFar != 0);
	uv[0] = result.x;
	uv[1] = result.y;
	return distance;
}

int32_t oracle_sphere_occlude(const EchoSphere* sphere, const float* origin, const float* direction, float travel, int32_t findFar)
{
	return sphere_occlude(*sphere, Ray(f3(origin), f3(direction)), travel, findFar != 0) ? 1 : 0;
}

float oracle_triangle_intersect(const EchoTriangle* triangle, const float* origin, const float* direction, float* uv)
{
	Float2 result = { 0.0f, 0.0f };
	float distance = triangle_intersect(*triangle, f3(origin), f3(direction), result);
	uv[0] = result.x;
	uv[1] = result.y;
	return distance;
}

int32_t oracle_triangle_occlude(const EchoTriangle* triangle, const float* origin, const float* direction, float travel)
{
	return triangle_occlude(*triangle, f3(origin), f3(direction), travel) ? 1 : 0;
}

// sphere / triangle area sampling: out = position(3), normal(3), pdf; returns 0 when impossible
int32_t oracle_geometry_sample(const OracleScene* s, uint32_t token, const float* origin, const float* sample, float* out7)
{
	GeometryPoint point = {};
	float pdf = 0.0f;
	bool ok = geometry_sample(s->scene, token, f3(origin), Float2{ sample1d(sample[0]), sample1d(sample[1]) }, point, pdf);
	out7[0] = point.position.x; out7[1] = point.position.y; out7[2] = point.position.z;
	out7[3] = point.normal.x; out7[4] = point.normal.y; out7[5] = point.normal.z;
	out7[6] = pdf;
	return ok ? 1 : 0;
}

float oracle_geometry_pdf(const OracleScene* s, uint32_t token, const float* origin, const float* incident)
{
	return geometry_pdf(s->scene, token, f3(origin), f3(incident));
}

// test-only: see PathTracedEvaluator::failedPickKeepsMis (evaluation.hpp). Process-wide; tests reset it to 0.
void oracle_set_failed_pick_keeps_mis(int enabled) { PathTracedEvaluator::failedPickKeepsMis = enabled != 0; }

// light tree: pick (returns token, writes pdf) and probability mass
uint32_t oracle_light_pick(const OracleScene* s, const float* position, const float* normal, float sample, float* pdf)
{
	GeometryPoint origin = { f3(position), f3(normal) };
	return s->scene.pick(origin, sample1d(sample), *pdf);
}

float oracle_light_mass(const OracleScene* s, uint32_t token, const float* position, const float* normal)
{
	GeometryPoint origin = { f3(position), f3(normal) };
	return s->scene.probability_mass(token, origin);
}

// BxDFs (src/Echo.UnitTests/Evaluation/BxDFTests.cs:49-83). kinds and parameter packing: see oracle.h
struct BxDFBox
{
	LambertianReflection lambertianReflection;
	Lambertian lambertian;
	OrenNayar orenNayar;
	SpecularReflection<RealFresnel> specularReflectionReal;
	SpecularReflection<ComplexFresnel> specularReflectionComplex;
	SpecularTransmission specularTransmission;
	SpecularFresnel specularFresnel;
	GlossyReflection<RealFresnel> glossyReflectionReal;
	GlossyReflection<ComplexFresnel> glossyReflectionComplex;
	GlossyTransmission glossyTransmission;
	CoatedLambertianReflection coatedLambertianReflection;

	const BxDF* configure(int32_t kind, const float* p)
	{
		auto rgb = [&](int at) { return RGB{ p[at], p[at + 1], p[at + 2] }; };

		switch (kind)
		{
			case ORACLE_BXDF_LAMBERTIAN_REFLECTION: return &lambertianReflection;
			case ORACLE_BXDF_LAMBERTIAN: return &lambertian;
			case ORACLE_BXDF_OREN_NAYAR: orenNayar.reset(p[0]); return &orenNayar;
			case ORACLE_BXDF_SPECULAR_REFLECTION_REAL: specularReflectionReal.fresnel = RealFresnel{ p[2], p[3] }; return &specularReflectionReal;
			case ORACLE_BXDF_SPECULAR_REFLECTION_COMPLEX: specularReflectionComplex.fresnel = ComplexFresnel(rgb(2), rgb(5), rgb(8)); return &specularReflectionComplex;
			case ORACLE_BXDF_SPECULAR_TRANSMISSION: specularTransmission.fresnel = RealFresnel{ p[2], p[3] }; return &specularTransmission;
			case ORACLE_BXDF_SPECULAR_FRESNEL: specularFresnel.fresnel = RealFresnel{ p[2], p[3] }; return &specularFresnel;
			case ORACLE_BXDF_GLOSSY_REFLECTION_REAL:
				glossyReflectionReal.microfacet = TrowbridgeReitz{ p[0], p[1] };
				glossyReflectionReal.fresnel = RealFresnel{ p[2], p[3] };
				return &glossyReflectionReal;
			case ORACLE_BXDF_GLOSSY_REFLECTION_COMPLEX:
				glossyReflectionComplex.microfacet = TrowbridgeReitz{ p[0], p[1] };
				glossyReflectionComplex.fresnel = ComplexFresnel(rgb(2), rgb(5), rgb(8));
				return &glossyReflectionComplex;
			case ORACLE_BXDF_GLOSSY_TRANSMISSION:
				glossyTransmission.microfacet = TrowbridgeReitz{ p[0], p[1] };
				glossyTransmission.fresnel = RealFresnel{ p[2], p[3] };
				return &glossyTransmission;
			case ORACLE_BXDF_COATED_LAMBERTIAN:
				coatedLambertianReflection.reset(rgb(5), RealFresnel{ p[2], p[3] }, p[4]);
				return &coatedLambertianReflection;
			default: return nullptr;
		}
	}
};

// for each i: sample(sample[i], outgoing[i]) -> sampled[i] = rgb(3), pdf, incident(3), type; then
// evaluate/pdf at (outgoing, incident) -> evaluated[i] = rgb(3), pdf; and the reciprocal pair (incident, outgoing) -> inverse[i] = rgb(3), pdf
void oracle_bxdf_batch(int32_t kind, const float* params, const float* outgoing, const float* samples, uint64_t n,
                       float* sampled8, float* evaluated4, float* inverse4)
{
	BxDFBox box;
	const BxDF* function = box.configure(kind, params);
	if (!function) return;

	for (uint64_t i = 0; i < n; i++)
	{
		Float3 out = f3(outgoing + i * 3);
		Float2 sample = { sample1d(samples[i * 2]), sample1d(samples[i * 2 + 1]) };

		Float3 incident = { 0.0f, 0.0f, 0.0f };
		ProbableRGB result = function->sample(sample, out, incident);

		float* s8 = sampled8 + i * 8;
		s8[0] = result.content.r; s8[1] = result.content.g; s8[2] = result.content.b; s8[3] = result.pdf;
		s8[4] = incident.x; s8[5] = incident.y; s8[6] = incident.z; s8[7] = (float)function->type;

		RGB value = function->evaluate(out, incident);
		float pdf = function->probability_density(out, incident);
		float* e4 = evaluated4 + i * 4;
		e4[0] = value.r; e4[1] = value.g; e4[2] = value.b; e4[3] = pdf;

		value = function->evaluate(incident, out);
		pdf = function->probability_density(incident, out);
		float* i4 = inverse4 + i * 4;
		i4[0] = value.r; i4[1] = value.g; i4[2] = value.b; i4[3] = pdf;
	}
}

// Accumulator (Processes/Evaluation/Accumulator.cs): feed n RGBA samples, out = value(4), noise max, count
void oracle_accumulate(const float* samples, uint64_t n, float* out6)
{
	Accumulator accumulator;
	for (uint64_t i = 0; i < n; i++) accumulator.add(Float4{ { samples[i * 4], samples[i * 4 + 1], samples[i * 4 + 2], samples[i * 4 + 3] } });
	Float4 value = accumulator.value();
	for (int i = 0; i < 4; i++) out6[i] = value.v[i];
	out6[4] = accumulator.noise_max();
	out6[5] = (float)accumulator.count;
}

float oracle_sample_value(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t dimension) { return sample_value(sample_key(seed, pixel, sample), dimension); }

} // extern "C"
