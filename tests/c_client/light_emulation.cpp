// light_emulation.cpp — test infrastructure: a sequential CPU backend for echorenderer_b200/csrc/echo_light_build.h, so that the CPU
// suite (-m "not gpu") can run the very passes and driver the device light-tree build runs (lightbuild.cu) and compare the emitted
// tree and emitter map with the host mirror's recursive build byte for byte. Not part of the product: libecho_b200.so never links
// this file, and nothing here is timed.
//
//   g++ -std=c++17 -O2 -ffp-contract=off -shared -fPIC tests/c_client/light_emulation.cpp -o liblight_emulation.so
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../echorenderer_b200/csrc/echo_light_build.h"

namespace
{

using namespace echo::lightbuild;

struct CpuBackend
{
	std::vector<char*> blocks;
	bool reverse = false; // run every pass over descending indices: no pass may depend on the order of its indices

	~CpuBackend() { for (char* block : blocks) std::free(block); }

	char* allocate(size_t bytes)
	{
		blocks.push_back((char*)std::calloc(bytes, 1));
		return blocks.back();
	}

	template<class F>
	bool for_each(uint32_t n, const F& f)
	{
		if (reverse) for (uint32_t i = n; i-- > 0u;) f(i);
		else for (uint32_t i = 0; i < n; i++) f(i);
		return true;
	}

	bool sort_pairs(const unsigned long long* keysIn, unsigned long long* keysOut, const uint32_t* valuesIn, uint32_t* valuesOut, uint32_t n, int)
	{
		std::vector<uint32_t> order(n);
		std::iota(order.begin(), order.end(), 0u);
		std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keysIn[a] < keysIn[b]; });
		for (uint32_t i = 0; i < n; i++) { keysOut[i] = keysIn[order[i]]; valuesOut[i] = valuesIn[order[i]]; }
		return true;
	}

	bool exclusive_sum(const uint32_t* in, uint32_t* out, uint32_t n)
	{
		uint32_t sum = 0u;
		for (uint32_t i = 0; i < n; i++) { uint32_t value = in[i]; out[i] = sum; sum += value; }
		return true;
	}

	template<class T> bool read(const T* source, T* destination, uint32_t n) { std::memcpy(destination, source, sizeof(T) * n); return true; }
	bool fill_zero(void* pointer, size_t bytes) { std::memset(pointer, 0, bytes); return true; }
};

} // namespace

// 0 = built; 1 = failed; 2 = refused (deeper than a 64-bit path). outNodes holds 2 * candidates nodes at most, the emitter arrays `candidates`.
extern "C" int32_t light_emulation_build(const EchoTriangle* triangles, uint32_t triangleCount, const EchoSphere* spheres, uint32_t sphereCount,
                                         const EchoMaterial* materials, uint32_t materialCount, const EchoPointLight* points, uint32_t pointCount,
                                         const float* instanceLights, uint32_t instanceCount, int32_t reverse,
                                         EchoLightNode* outNodes, uint32_t* outNodeCount, uint32_t* outTokens, uint64_t* outPaths, uint32_t* outEmitterCount, uint32_t* outLevels)
{
	CpuBackend backend;
	backend.reverse = reverse != 0;
	Sources sources = { triangles, triangleCount, spheres, sphereCount, materials, materialCount, points, pointCount, instanceLights, instanceCount };
	Result result = build(backend, sources);
	if (!result.ok) return 1;
	if (result.unsupported) return 2;

	std::memcpy(outNodes, result.nodes, sizeof(EchoLightNode) * result.nodeCount);
	std::memcpy(outTokens, result.emitterTokens, sizeof(uint32_t) * result.emitterCount);
	std::memcpy(outPaths, result.emitterPaths, sizeof(uint64_t) * result.emitterCount);
	*outNodeCount = result.nodeCount;
	*outEmitterCount = result.emitterCount;
	*outLevels = result.levels;
	return 0;
}

// the pinned transcendentals, for the accuracy checks of the suite
extern "C" double light_emulation_acos_double(double x) { return acos_double_pin(x); }
extern "C" float light_emulation_acos(float x) { return acos_pin(x); }
extern "C" float light_emulation_cos(float x) { return cos_pin(x); }
