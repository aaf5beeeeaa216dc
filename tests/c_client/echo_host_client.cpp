// echo_host_client.cpp — a C++ host written against include/echo_b200.hpp, the compiled-language mirror of the reference's interface
// for this path: it prepares a scene from raw arrays (PreparedScene), traces and occludes a batch, and renders a frame the way Echo's
// Device does — `workers` threads all enter EvaluationOperation::Execute(worker) until the operation reports no work left, each claiming
// procedures atomically (Common/Compute/Operation.cs:164-177, Worker.cs). tests/test_gpu_boundary.py writes the arrays, runs this binary
// and compares what it wrote with the ctypes path and the oracle. Linked directly against libecho_b200.so.
//
//   echo_host_client <directory> <workers> <hilbert 0|1> <build accelerator on device 0|1>
//
// <directory>: the files of tests/c_client/echo_client.c, plus profile.bin (9 x int32/float: evaluator, extend, minEpoch, maxEpoch,
// noiseThreshold, bounceLimit, survivability, seed, abortAfterTiles) and size.bin (3 x int32: width, height, tileSize); writes hits.bin,
// occluded.bin, frame.bin (height x width Float4, rows upward), sequence.bin (the tile sequence used), stats.bin, summary.txt.
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <thread>

#include "echo_b200.hpp"

using namespace echo_b200;

template<class T>
static std::vector<T> read_file(const std::string& directory, const char* name)
{
	std::ifstream file(directory + "/" + name, std::ios::binary | std::ios::ate);
	if (!file) throw std::runtime_error(std::string("cannot open ") + name);
	std::streamsize bytes = file.tellg();
	if (bytes % (std::streamsize)sizeof(T) != 0) throw std::runtime_error(std::string(name) + " is not a whole number of records");
	std::vector<T> data((size_t)bytes / sizeof(T));
	file.seekg(0);
	if (bytes > 0) file.read(reinterpret_cast<char*>(data.data()), bytes);
	return data;
}

template<class T>
static void write_file(const std::string& directory, const char* name, const T* data, size_t count)
{
	std::ofstream file(directory + "/" + name, std::ios::binary);
	file.write(reinterpret_cast<const char*>(data), (std::streamsize)(sizeof(T) * count));
	if (!file) throw std::runtime_error(std::string("cannot write ") + name);
}

template<class T>
static View<T> view(const std::vector<T>& data) { return { data.data(), data.size() }; }

struct AbortRequested {}; // what Worker.CheckSchedule throws when the operation is aborted

struct Worker : IWorker
{
	Worker(uint32_t index, std::atomic<int64_t>& tilesLeft) : index(index), tilesLeft(tilesLeft) {}
	uint32_t Index() const override { return index; }
	void CheckSchedule() override { if (tilesLeft.fetch_sub(1) <= 0) throw AbortRequested(); }
	uint32_t index;
	std::atomic<int64_t>& tilesLeft;
};

int main(int argc, char** argv)
{
	if (argc != 5) { std::fprintf(stderr, "usage: echo_host_client <directory> <workers> <hilbert> <build on device>\n"); return 1; }
	const std::string directory = argv[1];
	const uint32_t workers = (uint32_t)std::atoi(argv[2]);
	const bool hilbert = std::atoi(argv[3]) != 0, buildOnDevice = std::atoi(argv[4]) != 0;

	try
	{
		auto nodes = read_file<EchoQbvhNode>(directory, "nodes.bin");
		auto triangles = read_file<EchoTriangle>(directory, "triangles.bin");
		auto spheres = read_file<EchoSphere>(directory, "spheres.bin");
		auto materials = read_file<EchoMaterial>(directory, "materials.bin");
		auto lightNodes = read_file<EchoLightNode>(directory, "light_nodes.bin");
		auto emitterTokens = read_file<uint32_t>(directory, "emitter_tokens.bin");
		auto emitterPaths = read_file<uint64_t>(directory, "emitter_paths.bin");
		auto pointLights = read_file<EchoPointLight>(directory, "point_lights.bin");
		auto infinite = read_file<EchoInfiniteLight>(directory, "infinite.bin");
		auto camera = read_file<EchoCamera>(directory, "camera.bin");
		auto scalars = read_file<uint32_t>(directory, "scalars.bin");
		auto rays = read_file<EchoRay>(directory, "rays.bin");
		auto shadow = read_file<EchoRay>(directory, "shadow.bin");
		auto profileWords = read_file<uint32_t>(directory, "profile.bin");
		auto size = read_file<int32_t>(directory, "size.bin");
		if (camera.size() != 1 || scalars.size() != 4 || profileWords.size() != 9 || size.size() != 3) throw std::runtime_error("malformed input files");

		auto as_float = [](uint32_t bits) { float value; std::memcpy(&value, &bits, 4); return value; };

		SceneSource source;
		source.nodes = view(nodes);
		source.maxDepth = scalars[0];
		source.triangles = view(triangles);
		source.spheres = view(spheres);
		source.materials = view(materials);
		source.lightNodes = view(lightNodes);
		source.emitterTokens = view(emitterTokens);
		source.emitterPaths = view(emitterPaths);
		source.pointLights = view(pointLights);
		source.infiniteLights = view(infinite);
		source.infiniteThreshold = as_float(scalars[1]);
		source.infinitePdf = as_float(scalars[2]);
		source.boundRadius = as_float(scalars[3]);
		source.camera = camera[0];
		source.buildAcceleratorOnDevice = buildOnDevice;
		source.buildLightTreeOnDevice = buildOnDevice; // both trees by the library: the frame must not change by a bit

		PreparedScene scene(source);

		std::vector<EchoHit> hits(rays.size());
		std::vector<uint8_t> occluded(shadow.size());
		scene.Trace(rays.data(), rays.size(), hits.data());
		scene.Occlude(shadow.data(), shadow.size(), occluded.data());
		write_file(directory, "hits.bin", hits.data(), hits.size());
		write_file(directory, "occluded.bin", occluded.data(), occluded.size());

		EvaluationProfile profile;
		profile.Evaluator = (int32_t)profileWords[0];
		profile.Extend = (int32_t)profileWords[1];
		profile.MinEpoch = (int32_t)profileWords[2];
		profile.MaxEpoch = (int32_t)profileWords[3];
		profile.NoiseThreshold = as_float(profileWords[4]);
		profile.PathTraced.BounceLimit = (int32_t)profileWords[5];
		profile.PathTraced.Survivability = as_float(profileWords[6]);
		profile.Seed = profileWords[7];
		const int64_t abortAfterTiles = (int32_t)profileWords[8]; // < 0: never

		// a profile the reference rejects must be rejected here too, before anything is scheduled (EvaluationOperation.Factory's constructor)
		bool rejected = false;
		try { EvaluationProfile broken = profile; broken.MaxEpoch = broken.MinEpoch - 1; RenderTexture none({ 16, 16 }); EvaluationOperation::Factory(&scene, none, broken); }
		catch (const std::invalid_argument&) { rejected = true; }
		if (!rejected) throw std::runtime_error("an invalid profile was accepted");

		RenderTexture texture({ size[0], size[1] }, size[2]);
		EvaluationOperation::Factory factory(&scene, texture, profile, hilbert);
		std::unique_ptr<EvaluationOperation> operation(factory.CreateOperation(workers));

		std::atomic<int64_t> tilesLeft{ abortAfterTiles < 0 ? INT64_MAX : abortAfterTiles };
		std::atomic<uint32_t> aborted{ 0 };
		std::vector<std::thread> threads;

		for (uint32_t index = 0; index < workers; index++)
			threads.emplace_back([&, index]
			{
				Worker worker(index, tilesLeft);
				try { while (operation->Execute(worker)) {} }
				catch (const AbortRequested&) { ++aborted; } // the worker leaves the operation; tiles applied so far stay applied
			});

		for (std::thread& thread : threads) thread.join();

		EchoStats statistics = operation->Statistics();
		write_file(directory, "frame.bin", texture.Pixels().data(), texture.Pixels().size());
		write_file(directory, "sequence.bin", operation->tilePositions.data(), operation->tilePositions.size());
		write_file(directory, "stats.bin", &statistics, 1);

		std::ofstream summary(directory + "/summary.txt");
		summary << "procedures " << operation->CompletedProcedureCount() << " of " << operation->TotalProcedureCount() << "\n"
		        << "total_samples " << operation->TotalSamples() << "\n"
		        << "aborted_workers " << aborted.load() << "\n"
		        << "gpu_count " << scene.GpuCount() << "\n";

		// errors are pulled and thrown: an out-of-range device must surface as a NativeException carrying the library's message
		bool thrown = false;
		try { PreparedScene nowhere(source, 1ull << 63); }
		catch (const NativeException& exception) { thrown = exception.status != ECHO_B200_OK && std::strlen(exception.what()) > 0; }
		if (!thrown) throw std::runtime_error("a scene on a device that does not exist was accepted");

		std::cout << "echo_host_client ok" << std::endl;
		return 0;
	}
	catch (const std::exception& exception)
	{
		std::cerr << "echo_host_client: " << exception.what() << std::endl;
		return 3;
	}
}
