/*
 * echo_client.c — a host that is neither Python nor C++: plain C11 that dlopen()s libecho_b200.so, binds the entry points by
 * name (exactly what .NET's [DllImport] does), uploads a scene from raw arrays, and calls the two seams of the hot path:
 * echo_b200_trace_batch / echo_b200_occlude_batch (Accelerator.Trace / Occlude) and echo_b200_render_tiles
 * (EvaluationOperation.Execute). It only includes include/echo_b200.h; tests/test_gpu_c_client.py writes the arrays, runs
 * this binary and compares what it wrote with the ctypes path and the oracle.
 *
 *   echo_client <libecho_b200.so> <directory>
 *
 * <directory> holds nodes.bin triangles.bin spheres.bin materials.bin light_nodes.bin emitter_tokens.bin emitter_paths.bin
 * point_lights.bin infinite.bin camera.bin scalars.bin (u32 max_depth, f32 threshold, f32 pdf, f32 bound_radius) rays.bin
 * shadow.bin tiles.bin params.bin; the client writes hits.bin occluded.bin image.bin stats.bin there.
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "echo_b200.h"

static void* library;
static const char* (*last_error)(void);

static void* bind(const char* name)
{
	void* symbol = dlsym(library, name);
	if (!symbol) { fprintf(stderr, "missing symbol %s\n", name); exit(2); }
	return symbol;
}

static void check(int32_t status, const char* what)
{
	if (status == ECHO_B200_OK) return;
	fprintf(stderr, "%s failed with %d: %s\n", what, status, last_error()); /* errors are pulled, OidnDenoise.cs:201-206 */
	exit(3);
}

static void* read_file(const char* directory, const char* name, size_t record, size_t* count)
{
	char path[4096];
	snprintf(path, sizeof(path), "%s/%s", directory, name);
	FILE* file = fopen(path, "rb");
	if (!file) { fprintf(stderr, "cannot open %s\n", path); exit(4); }
	fseek(file, 0, SEEK_END);
	long bytes = ftell(file);
	fseek(file, 0, SEEK_SET);
	void* data = malloc(bytes > 0 ? (size_t)bytes : 1);
	if (bytes > 0 && fread(data, 1, (size_t)bytes, file) != (size_t)bytes) { fprintf(stderr, "short read of %s\n", path); exit(4); }
	fclose(file);
	if ((size_t)bytes % record != 0) { fprintf(stderr, "%s is not a whole number of %zu-byte records\n", path, record); exit(4); }
	*count = (size_t)bytes / record;
	return data;
}

static void write_file(const char* directory, const char* name, const void* data, size_t bytes)
{
	char path[4096];
	snprintf(path, sizeof(path), "%s/%s", directory, name);
	FILE* file = fopen(path, "wb");
	if (!file || fwrite(data, 1, bytes, file) != bytes) { fprintf(stderr, "cannot write %s\n", path); exit(4); }
	fclose(file);
}

int main(int argc, char** argv)
{
	if (argc != 3) { fprintf(stderr, "usage: %s <libecho_b200.so> <directory>\n", argv[0]); return 1; }
	const char* directory = argv[2];

	library = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
	if (!library) { fprintf(stderr, "dlopen: %s\n", dlerror()); return 2; } /* DllNotFoundException, OidnDenoise.cs:69 */
	last_error = (const char* (*)(void))bind("echo_b200_last_error");

	int32_t (*device_count)(int32_t*) = bind("echo_b200_device_count");
	int32_t (*scene_create)(EchoScene**, int32_t) = bind("echo_b200_scene_create");
	int32_t (*set_qbvh)(EchoScene*, const EchoQbvhNode*, uint32_t, uint32_t) = bind("echo_b200_scene_set_qbvh");
	int32_t (*set_triangles)(EchoScene*, const EchoTriangle*, uint32_t) = bind("echo_b200_scene_set_triangles");
	int32_t (*set_spheres)(EchoScene*, const EchoSphere*, uint32_t) = bind("echo_b200_scene_set_spheres");
	int32_t (*set_materials)(EchoScene*, const EchoMaterial*, uint32_t) = bind("echo_b200_scene_set_materials");
	int32_t (*set_light_tree)(EchoScene*, const EchoLightNode*, uint32_t, const uint32_t*, const uint64_t*, uint32_t, const EchoPointLight*, uint32_t) = bind("echo_b200_scene_set_light_tree");
	int32_t (*set_infinite)(EchoScene*, const EchoInfiniteLight*, uint32_t, float, float) = bind("echo_b200_scene_set_infinite");
	int32_t (*set_camera)(EchoScene*, const EchoCamera*) = bind("echo_b200_scene_set_camera");
	int32_t (*set_bound_radius)(EchoScene*, float) = bind("echo_b200_scene_set_bound_radius");
	int32_t (*commit)(EchoScene*) = bind("echo_b200_scene_commit");
	int32_t (*destroy)(EchoScene*) = bind("echo_b200_scene_destroy");
	int32_t (*trace_batch)(EchoScene*, const EchoRay*, uint64_t, EchoHit*) = bind("echo_b200_trace_batch");
	int32_t (*occlude_batch)(EchoScene*, const EchoRay*, uint64_t, uint8_t*) = bind("echo_b200_occlude_batch");
	int32_t (*render_tiles)(EchoScene*, const EchoRenderParams*, const int32_t*, uint32_t, float*, EchoStats*) = bind("echo_b200_render_tiles");

	int32_t devices = 0;
	check(device_count(&devices), "echo_b200_device_count");

	size_t nodeCount, triangleCount, sphereCount, materialCount, lightNodeCount, emitterCount, pathCount, pointCount, infiniteCount, one, rayCount, shadowCount, tileCount;
	EchoQbvhNode* nodes = read_file(directory, "nodes.bin", sizeof(EchoQbvhNode), &nodeCount);
	EchoTriangle* triangles = read_file(directory, "triangles.bin", sizeof(EchoTriangle), &triangleCount);
	EchoSphere* spheres = read_file(directory, "spheres.bin", sizeof(EchoSphere), &sphereCount);
	EchoMaterial* materials = read_file(directory, "materials.bin", sizeof(EchoMaterial), &materialCount);
	EchoLightNode* lightNodes = read_file(directory, "light_nodes.bin", sizeof(EchoLightNode), &lightNodeCount);
	uint32_t* emitterTokens = read_file(directory, "emitter_tokens.bin", sizeof(uint32_t), &emitterCount);
	uint64_t* emitterPaths = read_file(directory, "emitter_paths.bin", sizeof(uint64_t), &pathCount);
	EchoPointLight* pointLights = read_file(directory, "point_lights.bin", sizeof(EchoPointLight), &pointCount);
	EchoInfiniteLight* infinite = read_file(directory, "infinite.bin", sizeof(EchoInfiniteLight), &infiniteCount);
	EchoCamera* camera = read_file(directory, "camera.bin", sizeof(EchoCamera), &one);
	uint32_t* scalars = read_file(directory, "scalars.bin", 16, &one);
	EchoRay* rays = read_file(directory, "rays.bin", sizeof(EchoRay), &rayCount);
	EchoRay* shadow = read_file(directory, "shadow.bin", sizeof(EchoRay), &shadowCount);
	int32_t* tiles = read_file(directory, "tiles.bin", sizeof(int32_t) * 2, &tileCount);
	EchoRenderParams* params = read_file(directory, "params.bin", sizeof(EchoRenderParams), &one);
	if (emitterCount != pathCount) { fprintf(stderr, "emitter tokens and paths differ in count\n"); return 4; }

	float threshold, pdf, radius;
	memcpy(&threshold, scalars + 1, 4);
	memcpy(&pdf, scalars + 2, 4);
	memcpy(&radius, scalars + 3, 4);

	EchoScene* scene = NULL;
	check(scene_create(&scene, 0), "echo_b200_scene_create");
	check(set_qbvh(scene, nodes, (uint32_t)nodeCount, scalars[0]), "echo_b200_scene_set_qbvh");
	check(set_triangles(scene, triangles, (uint32_t)triangleCount), "echo_b200_scene_set_triangles");
	check(set_spheres(scene, spheres, (uint32_t)sphereCount), "echo_b200_scene_set_spheres");
	check(set_materials(scene, materials, (uint32_t)materialCount), "echo_b200_scene_set_materials");
	check(set_light_tree(scene, lightNodes, (uint32_t)lightNodeCount, emitterTokens, emitterPaths, (uint32_t)emitterCount, pointLights, (uint32_t)pointCount), "echo_b200_scene_set_light_tree");
	check(set_infinite(scene, infinite, (uint32_t)infiniteCount, threshold, pdf), "echo_b200_scene_set_infinite");
	check(set_camera(scene, camera), "echo_b200_scene_set_camera");
	check(set_bound_radius(scene, radius), "echo_b200_scene_set_bound_radius");
	check(commit(scene), "echo_b200_scene_commit");

	/* the library copied everything on upload: the host arrays can go (the caller owns them, OidnDenoise.cs:109-133) */
	free(nodes); free(triangles); free(spheres); free(materials); free(lightNodes); free(emitterTokens); free(emitterPaths); free(pointLights); free(infinite);

	EchoHit* hits = malloc(sizeof(EchoHit) * (rayCount ? rayCount : 1));
	uint8_t* occluded = malloc(shadowCount ? shadowCount : 1);
	check(trace_batch(scene, rays, rayCount, hits), "echo_b200_trace_batch");
	check(occlude_batch(scene, shadow, shadowCount, occluded), "echo_b200_occlude_batch");

	size_t pixels = tileCount * (size_t)params->tileSize * (size_t)params->tileSize;
	float* image = calloc(pixels ? pixels * 4 : 4, sizeof(float));
	EchoStats stats;
	check(render_tiles(scene, params, tiles, (uint32_t)tileCount, image, &stats), "echo_b200_render_tiles");

	/* a misuse must come back as an error code with a message, never as a crash */
	if (render_tiles(scene, NULL, tiles, (uint32_t)tileCount, image, &stats) != ECHO_B200_ERR_INVALID || strlen(last_error()) == 0)
	{
		fprintf(stderr, "a null EchoRenderParams was not rejected\n");
		return 5;
	}

	write_file(directory, "hits.bin", hits, sizeof(EchoHit) * rayCount);
	write_file(directory, "occluded.bin", occluded, shadowCount);
	write_file(directory, "image.bin", image, sizeof(float) * 4 * pixels);
	check(render_tiles(scene, params, tiles, (uint32_t)tileCount, image, &stats), "echo_b200_render_tiles (again)");
	write_file(directory, "stats.bin", &stats, sizeof(stats));

	check(destroy(scene), "echo_b200_scene_destroy");
	printf("echo_client ok: %d device(s), %zu rays, %zu tiles, %llu samples\n", devices, rayCount, tileCount, (unsigned long long)stats.sampleEvaluated);
	dlclose(library);
	return 0;
}
