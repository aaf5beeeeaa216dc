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

// ConeBound.Union without the whole-sphere shortcut of echo_light_build.h's cone_union (ConeBound.cs:76-101 as written), to check that the
// shortcut changes nothing: out5 = axis xyz, cosOffset, cosExtend of both forms for value0 = in[0..4], value1 = in[5..9]
extern "C" void light_emulation_cone_unions(const float* in, float* abridged5, float* unabridged5)
{
	Cone value0 = { { in[0], in[1], in[2] }, in[3], in[4] }, value1 = { { in[5], in[6], in[7] }, in[8], in[9] };
	Cone a = cone_union(value0, value1);

	Cone b;
	{
		float offset0 = acos_pin(clamp11(value0.cosOffset));
		float offset1 = acos_pin(clamp11(value1.cosOffset));
		float cosExtend = sse_min(value0.cosExtend, value1.cosExtend);
		Vec3 axis = value0.axis;
		float max = angle_degrees(value0.axis, value1.axis) + offset1;

		if (sse_min(max, kPi) <= offset0) b = { axis, value0.cosOffset, cosExtend };
		else
		{
			float offset = (offset0 + max) / 2.0f;
			if (offset >= kPi) b = { { 0.0f, 1.0f, 0.0f }, -1.0f, cosExtend };
			else
			{
				Vec3 c = normalized(cross(axis, value1.axis));
				axis = rotate_axis_angle(c, offset - offset0, axis);
				b = { axis, cos_pin(offset), cosExtend };
			}
		}
	}

	float* outs[2] = { abridged5, unabridged5 };
	const Cone* cones[2] = { &a, &b };
	for (int k = 0; k < 2; k++) { outs[k][0] = cones[k]->axis.x; outs[k][1] = cones[k]->axis.y; outs[k][2] = cones[k]->axis.z; outs[k][3] = cones[k]->cosOffset; outs[k][4] = cones[k]->cosExtend; }
}

// cone_union tests `cosOffset <= -1` instead of `acos(clamp(cosOffset)) >= pi`: counts the floats in (-1, -0.5] — every one of them — and a
// stride through (-0.5, 1] for which acos_pin reaches pi after all (expected: 0), and reports acos_pin(-1) == pi as bit 31
extern "C" uint32_t light_emulation_acos_reaches_pi()
{
	uint32_t reached = 0u;
	for (uint32_t bits = 0xBF000000u; bits < 0xBF800000u; bits++) if (acos_pin(bits_float(bits)) >= kPi) ++reached;     // -0.5 .. just above -1
	for (uint32_t bits = 0x80000000u; bits < 0xBF000000u; bits += 97u) if (acos_pin(bits_float(bits)) >= kPi) ++reached;  // -0 .. -0.5
	for (uint32_t bits = 0x00000000u; bits <= 0x3F800000u; bits += 97u) if (acos_pin(bits_float(bits)) >= kPi) ++reached; // +0 .. 1
	if (acos_pin(-1.0f) >= kPi && acos_pin(clamp11(-1.5f)) >= kPi) reached |= 0x80000000u;
	return reached;
}
