// sweep_emulation.cpp — test infrastructure: a sequential CPU backend for echorenderer_b200/csrc/echo_sweep.h, so that the CPU suite
// (-m "not gpu") can run the very passes and driver the device build runs (sweep.cu) and compare the emitted QBVH with the host
// mirror's byte for byte. Not part of the product: libecho_b200.so never links this file, and nothing here is timed.
//
//   g++ -std=c++17 -O2 -ffp-contract=off -shared -fPIC tests/c_client/sweep_emulation.cpp -o libsweep_emulation.so
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <numeric>

#include "../../echorenderer_b200/csrc/echo_sweep.h"

namespace
{

using namespace echo::sweep;

struct CpuBackend
{
	char* block = nullptr;
	bool reverse = false; // run every pass over descending indices: the passes must not depend on the order

	~CpuBackend() { std::free(block); }

	char* allocate(size_t bytes) { return block = (char*)std::calloc(bytes, 1); }

	template<class F>
	bool for_each(uint32_t n, const F& f)
	{
		if (reverse) for (uint32_t i = n; i-- > 0u;) f(i);
		else for (uint32_t i = 0; i < n; i++) f(i);
		return true;
	}

	bool sort_pairs(const unsigned long long* keysIn, unsigned long long* keysOut, const uint32_t* valuesIn, uint32_t* valuesOut, uint32_t n, int endBit)
	{
		const unsigned long long mask = endBit >= 64 ? ~0ull : (1ull << endBit) - 1ull;
		std::vector<uint32_t> order(n);
		std::iota(order.begin(), order.end(), 0u);
		std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return (keysIn[a] & mask) < (keysIn[b] & mask); });
		for (uint32_t i = 0; i < n; i++) { keysOut[i] = keysIn[order[i]]; valuesOut[i] = valuesIn[order[i]]; }
		return true;
	}

	bool scan_items(const ScanItem* in, ScanItem* out, uint32_t n)
	{
		ScanOp op;
		for (uint32_t i = 0; i < n; i++) out[i] = i == 0u ? in[0] : op(out[i - 1u], in[i]);
		return true;
	}

	bool exclusive_sum(const uint32_t* in, uint32_t* out, uint32_t n)
	{
		uint32_t sum = 0u;
		for (uint32_t i = 0; i < n; i++) { uint32_t value = in[i]; out[i] = sum; sum += value; }
		return true;
	}

	template<class T> bool read(const T* source, T* destination, uint32_t n) { std::memcpy(destination, source, sizeof(T) * n); return true; }
	template<class T> bool write(T* destination, const T* source, uint32_t n) { std::memcpy(destination, source, sizeof(T) * n); return true; }
	bool fill_zero(void* pointer, size_t bytes) { std::memset(pointer, 0, bytes); return true; }
};

} // namespace

// 0 = built; 1 = failed; 2 = gave up (deeper than kMaxLevels). `out` holds (primitives - 1) nodes at most.
extern "C" int32_t sweep_emulation_build(const EchoTriangle* triangles, uint32_t triangleCount, const EchoSphere* spheres, uint32_t sphereCount,
                                         const float* instanceBounds, uint32_t instanceCount, int32_t reverse,
                                         EchoQbvhNode* out, uint32_t* outNodeCount, uint32_t* outMaxDepth, uint32_t* outLevels)
{
	CpuBackend backend;
	backend.reverse = reverse != 0;
	Result result = build(backend, triangles, triangleCount, spheres, sphereCount, instanceBounds, instanceCount);
	if (!result.ok) return 1;
	if (result.gaveUp) return 2;
	std::memcpy(out, result.quads, sizeof(EchoQbvhNode) * result.nodeCount);
	*outNodeCount = result.nodeCount;
	*outMaxDepth = result.maxDepth;
	if (quad_depth(out, result.nodeCount) != result.maxDepth) return 3; // the depth the passes computed is not the emitted array's
	*outLevels = result.levels;
	return 0;
}
