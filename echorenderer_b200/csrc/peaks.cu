// peaks.cu — the on-chip ceilings the traversal kernels run against, measured on the device the bench runs on.
//
// The HBM roofline (MEASURED_PEAKS.json: 6.5 TB/s copy bandwidth) is the wrong denominator for a kernel whose working set lives
// in L2: the C2 closest-hit kernel moves 881 algorithmic bytes per query, of which only the 48 bytes of ray + hit come from DRAM;
// the node and triangle bytes are 32-byte sectors fetched by divergent lanes through the L1 data pipe, mostly from L2 (ncu:
// L1 hit 40 %, L2 hit 73 %, DRAM 7 %). These microbenchmarks measure what the memory system delivers for exactly those access
// shapes, so bench.py can report `frac_l2` and `frac_l1_sectors` beside the HBM fraction:
//   l2_read      coalesced 128-bit loads (ld.global.cg: L1 bypassed) over a 32 MB buffer resident in L2
//   l2_sector    one random 32-byte sector per lane per load (ld.global.nc.v8.f32, the node-fetch instruction) over 64 MB in L2:
//                every load misses L1 and costs one L2 sector — the shape of an incoherent node visit
//   l1_sector    the same instruction over an 8 KB window per CTA (64 KB per SM) that stays in L1: the LSU / L1 data-pipe ceiling
// Each kernel runs one resident wave of CTAs for a few milliseconds; rates are algorithmic bytes / event time.
#include <algorithm>

#include "echo_internal.h"
#include "echo_traverse.cuh"

namespace echo
{

constexpr int kPeakBlock = 256;

__global__ void __launch_bounds__(kPeakBlock) l2_read_kernel(const float4* __restrict__ data, uint64_t count, int passes, float* __restrict__ sink)
{
	const uint64_t stride = (uint64_t)gridDim.x * kPeakBlock;
	float4 sum = make_float4(0, 0, 0, 0);

	for (int pass = 0; pass < passes; pass++)
	{
		// every pass starts at another offset so that consecutive passes of one thread never touch the same line
		for (uint64_t i = ((uint64_t)blockIdx.x * kPeakBlock + threadIdx.x + (uint64_t)pass * 977u * kPeakBlock) % stride; i < count; i += stride)
		{
			float4 v = __ldcg(data + i);
			sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
		}
	}

	if (sum.x + sum.y + sum.z + sum.w == 12345.678f) *sink = sum.x; // keeps the loads alive
}

// one 32-byte sector per lane per load, sector index from a per-lane LCG: `window` sectors starting at `base`
template<bool PER_BLOCK_WINDOW>
__global__ void __launch_bounds__(kPeakBlock) sector_kernel(const float* __restrict__ data, uint32_t sectors, uint32_t window, int loads, float* __restrict__ sink)
{
	uint32_t state = (blockIdx.x * kPeakBlock + threadIdx.x) * 2654435761u + 12345u;
	const uint32_t base = PER_BLOCK_WINDOW ? (uint32_t)(((uint64_t)blockIdx.x * window) % (sectors - window)) : 0u;
	float sum = 0.0f;

	#pragma unroll 4
	for (int i = 0; i < loads; i++)
	{
		state = state * 1664525u + 1013904223u;
		uint32_t sector = base + (state >> 8) % window;
		float8 v = ldg256(data + (size_t)sector * 8);
		sum += v.v[0] + v.v[7];
	}

	if (sum == 12345.678f) *sink = sum;
}

template<class Launch>
static bool timed(Launch launch, float& milliseconds)
{
	cudaEvent_t begin = nullptr, end = nullptr;
	if (!check_cuda(cudaEventCreate(&begin), "cudaEventCreate") || !check_cuda(cudaEventCreate(&end), "cudaEventCreate")) return false;
	launch(); // warm-up: brings the buffer into L2 / L1
	float best = 1e30f;
	bool ok = true;

	for (int repeat = 0; ok && repeat < 3; repeat++)
	{
		cudaEventRecord(begin, nullptr);
		launch();
		cudaEventRecord(end, nullptr);
		ok = check_cuda(cudaEventSynchronize(end), "peak kernel");
		float ms = 0.0f;
		if (ok) cudaEventElapsedTime(&ms, begin, end);
		best = std::min(best, ms);
	}

	cudaEventDestroy(begin);
	cudaEventDestroy(end);
	milliseconds = best;
	return ok && check_cuda(cudaGetLastError(), "peak kernel launch");
}

// out[0] = L2 coalesced read GB/s, out[1] = L2 random-sector GB/s, out[2] = L1 random-sector GB/s
bool measure_peaks(float* out)
{
	int device = 0, sms = 0;
	cudaGetDevice(&device);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
	const int grid = std::max(sms, 1) * 8; // 2048 threads per SM: one resident wave

	const uint64_t bytes = 64ull << 20;
	float* data = nullptr;
	float* sink = nullptr;
	if (!check_cuda(cudaMalloc((void**)&data, bytes), "cudaMalloc(peaks)") || !check_cuda(cudaMalloc((void**)&sink, sizeof(float)), "cudaMalloc(peaks)")) { cudaFree(data); return false; }
	bool ok = check_cuda(cudaMemset(data, 0, bytes), "cudaMemset(peaks)");
	float ms = 0.0f;

	if (ok)
	{
		const uint64_t count = (32ull << 20) / sizeof(float4); // 32 MB: comfortably inside the 126 MB L2
		const int passes = 64;
		ok = timed([&] { l2_read_kernel<<<grid, kPeakBlock>>>(reinterpret_cast<const float4*>(data), count, passes, sink); }, ms);
		out[0] = ok ? (float)((double)count * sizeof(float4) * passes / (ms * 1e-3) / 1e9) : 0.0f;
	}

	if (ok)
	{
		const uint32_t sectors = (uint32_t)(bytes / 32);
		const int loads = 512;
		ok = timed([&] { sector_kernel<false><<<grid, kPeakBlock>>>(data, sectors, sectors, loads, sink); }, ms);
		out[1] = ok ? (float)((double)grid * kPeakBlock * loads * 32.0 / (ms * 1e-3) / 1e9) : 0.0f;
	}

	if (ok)
	{
		const uint32_t sectors = (uint32_t)(bytes / 32);
		const int loads = 4096;
		ok = timed([&] { sector_kernel<true><<<grid, kPeakBlock>>>(data, sectors, 256u, loads, sink); }, ms); // 256 sectors = 8 KB per CTA, 8 CTAs per SM
		out[2] = ok ? (float)((double)grid * kPeakBlock * loads * 32.0 / (ms * 1e-3) / 1e9) : 0.0f;
	}

	cudaFree(data);
	cudaFree(sink);
	return ok;
}

} // namespace echo
