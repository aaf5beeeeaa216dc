/*
 * echo_host.h — host-side scene preparation (libecho_host.so, plain C++, no CUDA).
 *
 * In a real deployment this work stays in Echo's C# host: ScenePreparer -> PreparedPack builds the SweepBuilder
 * hierarchy, collapses it to the QuadBoundingVolumeHierarchy node array and builds the LightTree, and the C# shim
 * hands those arrays to libecho_b200.so. This container has no .NET, so this library is the
 * host-side mirror of exactly those build steps (same algorithms, same array formats), used by the Python harness,
 * the tests and bench.py to produce the inputs of echo_b200.h. References (src/Echo.Core/):
 *   Aggregation/Acceleration/SweepBuilder.cs:24-170            full-sweep SAH, stable radix sort per axis change
 *   Aggregation/Acceleration/QuadBoundingVolumeHierarchy.cs:24-36,363-565  binary -> quad collapse, depth-first node array
 *   Aggregation/Selection/LightTree.cs:21-38,62-113             light tree build + bit-path map
 *   Aggregation/Bounds/{LightBound,ConeBound}.cs                light/cone bounds
 *   Aggregation/Preparation/LightCollection.cs:83-137           light bounds creation order
 *   Aggregation/Preparation/PreparedScene.cs:279-325            infinite light threshold
 */
#ifndef ECHO_HOST_H
#define ECHO_HOST_H

#include "echo_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Builds the QBVH over triangles then spheres (token order of GeometryCollection.CreateBounds, GeometryCollection.cs:52-81).
 * Output arrays are malloc'd; release with echo_host_free. `threads` <= 0 uses all hardware threads. */
int32_t echo_host_build_qbvh(const EchoTriangle* triangles, uint32_t triangle_count,
                             const EchoSphere* spheres, uint32_t sphere_count, int32_t threads,
                             EchoQbvhNode** out_nodes, uint32_t* out_node_count, uint32_t* out_max_depth);

/* The same with the boxes of the pack's instances appended (PreparedInstance.BoxBound, PreparedInstance.cs:29):
 * instance_bounds holds 6 floats per instance, min xyz then max xyz, in parent space. */
int32_t echo_host_build_qbvh_instanced(const EchoTriangle* triangles, uint32_t triangle_count,
                                       const EchoSphere* spheres, uint32_t sphere_count,
                                       const float* instance_bounds, uint32_t instance_count, int32_t threads,
                                       EchoQbvhNode** out_nodes, uint32_t* out_node_count, uint32_t* out_max_depth);

/* Builds the light tree over point lights, emissive triangles, emissive spheres (LightCollection.CreateBounds order).
 * out_power receives the root LightBound power (0 when there is no light; then node_count == 0). The recursion is the reference's
 * (LightTree.cs:62-113); the LightBound / ConeBound arithmetic is the one restatement it shares with the device build
 * (echo_b200_build_light_tree, csrc/echo_light_build.h: stable sort, MathF.Acos / Cos, Math.Acos and the Versor's Math.Sin / Cos pinned), so both emit the same bytes. */
int32_t echo_host_build_light_tree(const EchoTriangle* triangles, uint32_t triangle_count,
                                   const EchoSphere* spheres, uint32_t sphere_count,
                                   const EchoMaterial* materials, uint32_t material_count,
                                   const EchoPointLight* points, uint32_t point_count,
                                   EchoLightNode** out_nodes, uint32_t* out_node_count,
                                   uint32_t** out_emitter_tokens, uint64_t** out_emitter_bitpaths, uint32_t* out_emitter_count,
                                   float* out_power);

/* The same with the light bounds of the pack's instances appended (PreparedInstance.LightBound, PreparedInstance.cs:31-38;
 * LightCollection.AddInstances, LightCollection.cs:123-135). instance_lights holds 12 floats per instance: box min xyz,
 * box max xyz, cone axis xyz, cosOffset, cosExtend, power (parent space); instances without power are skipped. */
int32_t echo_host_build_light_tree_instanced(const EchoTriangle* triangles, uint32_t triangle_count,
                                             const EchoSphere* spheres, uint32_t sphere_count,
                                             const EchoMaterial* materials, uint32_t material_count,
                                             const EchoPointLight* points, uint32_t point_count,
                                             const float* instance_lights, uint32_t instance_count,
                                             EchoLightNode** out_nodes, uint32_t* out_node_count,
                                             uint32_t** out_emitter_tokens, uint64_t** out_emitter_bitpaths, uint32_t* out_emitter_count,
                                             float* out_power);

/* PreparedScene.CalculateThreshold (PreparedScene.cs:317-325) and the ambient-light power of AmbientLight.Prepare
 * (AmbientLight.cs:42-51) with the scene radius taken as the half diagonal of the root bound. */
float echo_host_infinite_threshold(float infinite_power, float scene_power);
float echo_host_ambient_power(const float radiance[3], const EchoQbvhNode* root);

/* Emissive.Power for a constant emission colour (Emissive.cs:52-53). */
float echo_host_emissive_power(const float emission[3]);

/* CoatedLambertianReflection.FresnelDiffuseReflectance / ...Fast (Lambertian.cs:168-230): what CoatedDiffuse.Prepare caches. */
float echo_host_fresnel_diffuse_reflectance(float eta);
float echo_host_fresnel_diffuse_reflectance_fast(float eta);

void echo_host_free(void* pointer);

#ifdef __cplusplus
}
#endif
#endif
