// echo_host_logic.cpp — the parts of include/echo_b200.hpp that need no device, driven from the CPU suite (tests/test_abi.py):
// tile patterns (checked against the Python mirror, which is checked against the reference's TilePatternTests), RenderTexture.Apply
// on ragged edge tiles, EvaluationProfile.Validate's rejections, and an operation whose scene is null (aborted preparation).
//
//   echo_host_logic pattern <hilbert 0|1> <width> <height>     prints "x y" per tile
//   echo_host_logic checks                                      prints "echo_host_logic ok"
#include <cstdio>
#include <cstring>
#include <memory>

#include "echo_b200.hpp"

using namespace echo_b200;

struct Worker : IWorker
{
	uint32_t Index() const override { return 0; }
	void CheckSchedule() override { ++checks; }
	int checks = 0;
};

static bool rejects(EvaluationProfile profile)
{
	try { profile.Validate(); }
	catch (const std::invalid_argument&) { return true; }
	return false;
}

int main(int argc, char** argv)
{
	if (argc == 5 && std::strcmp(argv[1], "pattern") == 0)
	{
		Int2 size = { std::atoi(argv[3]), std::atoi(argv[4]) };
		for (Int2 p : std::atoi(argv[2]) ? HilbertCurvePattern(size) : OrderedPattern(size)) std::printf("%d %d\n", p.X, p.Y);
		return 0;
	}

	if (argc == 2 && std::strcmp(argv[1], "checks") == 0)
	{
		EvaluationProfile good;
		EvaluationProfile noEvaluator = good, noExtend = good, minEpoch = good, maxEpoch = good, noise = good;
		noEvaluator.Evaluator = -1; noExtend.Extend = 0; minEpoch.MinEpoch = 0; maxEpoch.MaxEpoch = good.MinEpoch - 1; noise.NoiseThreshold = -0.5f;
		if (rejects(good) || !rejects(noEvaluator) || !rejects(noExtend) || !rejects(minEpoch) || !rejects(maxEpoch) || !rejects(noise)) return 2;

		RenderTexture texture({ 20, 10 }, 16); // 2 x 1 tiles, the right one 4 pixels wide, both 10 high
		if (texture.TileCount().X != 2 || texture.TileCount().Y != 1) return 3;
		std::vector<Float4> tile(16 * 16);
		for (int i = 0; i < 256; i++) tile[i] = { (float)(i % 16), (float)(i / 16), 7.0f, 0.0f };
		texture.Apply({ 1, 0 }, tile.data());
		for (int y = 0; y < 10; y++)
			for (int x = 0; x < 20; x++)
			{
				Float4 pixel = texture[{ x, y }];
				bool inside = x >= 16;
				if (inside ? (pixel.X != (float)(x - 16) || pixel.Y != (float)y || pixel.Z != 7.0f) : (pixel.X != 0.0f || pixel.Z != 0.0f)) return 4;
			}

		// 600 tiles -> 3 procedures; with a null scene every procedure completes without touching the library or the destination
		RenderTexture big({ 30 * 16, 20 * 16 }, 16);
		EvaluationOperation::Factory factory(nullptr, big, good, true);
		bool noDevice = false;
		std::unique_ptr<EvaluationOperation> operation;
		try { operation.reset(factory.CreateOperation(1)); }
		catch (const NativeException& exception) { noDevice = exception.status == ECHO_B200_ERR_NO_DEVICE; } // page-locked memory needs a device
		if (operation)
		{
			if (operation->tilePositions.size() != 600 || operation->TotalProcedureCount() != 3) return 5;
			Worker worker;
			int calls = 0;
			while (operation->Execute(worker)) ++calls;
			if (calls != 2 || !operation->IsCompleted() || operation->TotalSamples() != 0 || worker.checks != 0) return 6;
			if (operation->Execute(worker)) return 7; // nothing left to claim
		}
		else if (!noDevice) return 8;

		std::printf("echo_host_logic ok%s\n", noDevice ? " (no device: operation not created)" : "");
		return 0;
	}

	std::fprintf(stderr, "usage: echo_host_logic pattern <hilbert> <width> <height> | checks\n");
	return 1;
}
