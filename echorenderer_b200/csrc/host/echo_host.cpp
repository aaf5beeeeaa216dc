// echo_host.cpp — libecho_host.so: host-side scene preparation (plain C++, no CUDA). See include/echo_host.h.
//
// In a real deployment these steps stay in Echo's C# host (ScenePreparer -> PreparedPack); this is the host-side mirror
// that produces the same arrays for the Python harness, the tests and bench.py. It is not on the GPU hot path.
// Algorithms follow (paths relative to the reference's src/Echo.Core/):
//   Aggregation/Acceleration/SweepBuilder.cs:24-170            full-sweep SAH over stably sorted bounds
//   Aggregation/Acceleration/QuadBoundingVolumeHierarchy.cs:24-36,363-565  binary -> quad collapse, pre-order node array
//   Aggregation/Selection/LightTree.cs:21-113                  light tree + per-emitter branch bit paths
//   Aggregation/Bounds/{BoxBound,LightBound,ConeBound}.cs
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <thread>
#include <vector>

#include "../../../include/echo_host.h"
#include "../echo_light_build.h"

namespace
{

namespace lightbuild = echo::lightbuild;

constexpr float kInf = std::numeric_limits<float>::infinity();
constexpr float kPi = 3.14159265358979323846f;
constexpr float kTau = 6.28318530717958647692f;
constexpr float kEpsilon = 8E-7f;

struct Vec3
{
	float x, y, z;
	float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

inline Vec3 operator+(Vec3 a, Vec3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline Vec3 operator-(Vec3 a, Vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline Vec3 operator*(Vec3 a, float b) { return { a.x * b, a.y * b, a.z * b }; }
inline Vec3 v3(const float* p) { return { p[0], p[1], p[2] }; }

inline Vec3 cross(Vec3 a, Vec3 b) // Common/Packed/Float3.cs:268-273
{
	return {
		(float)((double)a.y * b.z - (double)a.z * b.y),
		(float)((double)a.z * b.x - (double)a.x * b.z),
		(float)((double)a.x * b.y - (double)a.y * b.x)
	};
}

inline double squared_double(Vec3 a) { return (double)a.x * a.x + (double)a.y * a.y + (double)a.z * a.z; }
inline float magnitude(Vec3 a) { return (float)std::sqrt(squared_double(a)); }

inline Vec3 normalized(Vec3 a) // Float3.cs:171-181
{
	double squared = squared_double(a);
	if (squared == 0.0 || std::fabs(squared) < 1E-10 * 2.2250738585072014e-308) return { 0, 0, 0 };
	return a * (1.0f / (float)std::sqrt(squared));
}

inline float min_net(float a, float b) { return (a != a) ? a : (b != b) ? b : (a == b ? (std::signbit(a) ? a : b) : (a < b ? a : b)); } // Math.Min
inline float max_net(float a, float b) { return (a != a) ? a : (b != b) ? b : (a == b ? (std::signbit(a) ? b : a) : (a > b ? a : b)); } // Math.Max

struct Box // Aggregation/Bounds/BoxBound.cs
{
	Vec3 min, max;

	float half_area() const // :80-87
	{
		Vec3 size = max - min;
		return size.x * (size.y + size.z) + size.y * size.z;
	}

	int major_axis() const // :92 + Float3.MaxIndex (Float3.cs:130-138)
	{
		Vec3 s = max - min;
		if (s.x > s.y) return s.x > s.z ? 0 : 2;
		return s.y > s.z ? 1 : 2;
	}

	Box encapsulate(const Box& o) const // :128-132
	{
		return { { min_net(min.x, o.min.x), min_net(min.y, o.min.y), min_net(min.z, o.min.z) },
		         { max_net(max.x, o.max.x), max_net(max.y, o.max.y), max_net(max.z, o.max.z) } };
	}
};

struct Tokenized
{
	uint32_t token;
	Box box;
};

// ---------------- SweepBuilder ----------------

struct BinaryNode // HierarchyBuilder.Node (Aggregation/Acceleration/HierarchyBuilder.cs:17-94)
{
	Box box;
	int32_t child0 = -1, child1 = -1;
	int32_t axis = 0;
	uint32_t token = ECHO_TOKEN_EMPTY;

	bool is_leaf() const { return child0 < 0; }
};

struct NodePool
{
	std::vector<BinaryNode> nodes;
	std::atomic<int32_t> next{ 0 };

	int32_t allocate() { return next.fetch_add(1); }
};

std::atomic<int> gActiveBuilders{ 0 };
int gMaxBuilders = 1;

inline uint32_t sort_key(float value) // Sorter.Transform, SweepBuilder.cs:241-251
{
	uint32_t converted;
	std::memcpy(&converted, &value, 4);
	uint32_t flip = (converted >> 31) * (0xFFFFFFFFu - 0x80000000u);
	return converted ^ (flip + 0x80000000u);
}

struct SweepBuilder
{
	NodePool& pool;
	std::vector<uint32_t> keys0, keys1;
	std::vector<Tokenized> buffer;
	std::vector<Box> cutTailVolumes;

	explicit SweepBuilder(NodePool& pool) : pool(pool) {}

	int32_t make_leaf(const Tokenized& item)
	{
		int32_t index = pool.allocate();
		BinaryNode& node = pool.nodes[index];
		node.box = item.box;
		node.token = item.token;
		return index;
	}

	// Sorter.Sort, SweepBuilder.cs:199-236: stable by key (insertion sort <= 32, LSD radix above)
	void sort_indices(Tokenized* data, int length, int axis)
	{
		if ((int)keys0.size() < length) keys0.resize(length);
		for (int i = 0; i < length; i++) keys0[i] = sort_key(data[i].box.min[axis] + data[i].box.max[axis]);

		if (length <= 32)
		{
			for (int i = 1; i < length; i++)
			{
				uint32_t key = keys0[i];
				Tokenized value = data[i];
				int scan = i;

				while (scan > 0 && key < keys0[scan - 1])
				{
					keys0[scan] = keys0[scan - 1];
					data[scan] = data[scan - 1];
					--scan;
				}

				keys0[scan] = key;
				data[scan] = value;
			}

			return;
		}

		if ((int)keys1.size() < length) keys1.resize(length);
		if ((int)buffer.size() < length) buffer.resize(length);

		uint32_t* k0 = keys0.data();
		uint32_t* k1 = keys1.data();
		Tokenized* v0 = data;
		Tokenized* v1 = buffer.data();

		for (int round = 0; round < 4; round++)
		{
			int shift = round * 8;
			int counts[256] = {};

			for (int i = 0; i < length; i++) ++counts[(k0[i] >> shift) & 0xFF];

			int sum = 0;
			for (int& count : counts) { count += sum; sum = count; }

			for (int i = length - 1; i >= 0; i--)
			{
				uint32_t key = k0[i];
				int index = --counts[(key >> shift) & 0xFF];
				k1[index] = key;
				v1[index] = v0[i];
			}

			std::swap(k0, k1);
			std::swap(v0, v1);
		}

		// four swaps: the sorted values are back in `data`
	}

	int32_t build(Tokenized* data, int length) // Build, SweepBuilder.cs:24-36
	{
		if (length == 1) return make_leaf(data[0]);

		Box all = { { kInf, kInf, kInf }, { -kInf, -kInf, -kInf } };
		for (int i = 0; i < length; i++) all = all.encapsulate(data[i].box);

		sort_indices(data, length, all.major_axis());
		return build_layer(data, length);
	}

	int32_t build_layer(Tokenized* data, int length) // BuildLayer, SweepBuilder.cs:38-88
	{
		// PrepareCutTailVolumes, :115-127
		if ((int)cutTailVolumes.size() < length) cutTailVolumes.resize(length);
		Box cutTailVolume = data[length - 1].box;

		for (int i = length - 2; i >= 0; i--)
		{
			cutTailVolumes[i + 1] = cutTailVolume;
			cutTailVolume = cutTailVolume.encapsulate(data[i].box);
		}

		// SearchSurfaceAreaHeuristics, :132-160
		Box cutHeadVolume = data[0].box;
		float minCost = std::numeric_limits<float>::max();
		int minIndex = -1;
		Box headVolume = {}, tailVolume = {};

		for (int i = 1; i < length; i++)
		{
			const Box& tail = cutTailVolumes[i];
			float cost = cutHeadVolume.half_area() * (float)i + tail.half_area() * (float)(length - i);

			if (cost < minCost)
			{
				minCost = cost;
				minIndex = i;
				headVolume = cutHeadVolume;
				tailVolume = tail;
			}

			cutHeadVolume = cutHeadVolume.encapsulate(data[i].box);
		}

		if (minIndex < 0) // every cost was NaN/inf (degenerate bounds): fall back to a median split
		{
			minIndex = length / 2;
			headVolume = data[0].box;
			for (int i = 1; i < minIndex; i++) headVolume = headVolume.encapsulate(data[i].box);
			tailVolume = cutTailVolumes[minIndex];
		}

		Box bound = headVolume.encapsulate(tailVolume);
		int axis = bound.major_axis();

		Tokenized* headData; int headLength;
		Tokenized* tailData; int tailLength;

		if (minIndex > length / 2)
		{
			headData = data; headLength = minIndex;
			tailData = data + minIndex; tailLength = length - minIndex;
		}
		else
		{
			headData = data + minIndex; headLength = length - minIndex;
			tailData = data; tailLength = minIndex;
			std::swap(headVolume, tailVolume);
		}

		int32_t child0, child1;
		bool parallel = headLength >= 4096 && gActiveBuilders.load() < gMaxBuilders; // ParallelBuildThreshold, :22

		if (!parallel)
		{
			child0 = build_child(headData, headLength, headVolume, axis);
			child1 = build_child(tailData, tailLength, tailVolume, axis);
		}
		else
		{
			// LayerBuilder, :162-188: a fresh builder (own scratch) sorts if needed and builds the head subtree
			++gActiveBuilders;
			int sortAxis = headVolume.major_axis();
			if (sortAxis == axis) sortAxis = -1;

			std::thread worker([&, sortAxis]()
			{
				SweepBuilder builder(pool);
				if (sortAxis >= 0) builder.sort_indices(headData, headLength, sortAxis);
				child0 = builder.build_layer(headData, headLength);
				--gActiveBuilders;
			});

			child1 = build_child(tailData, tailLength, tailVolume, axis);
			worker.join();
		}

		if (headVolume.half_area() < tailVolume.half_area()) std::swap(child0, child1);

		int32_t index = pool.allocate();
		BinaryNode& node = pool.nodes[index];
		node.box = bound;
		node.child0 = child0;
		node.child1 = child1;
		node.axis = axis;
		return index;
	}

	int32_t build_child(Tokenized* data, int length, const Box& parent, int parentAxis) // BuildChild, :90-97
	{
		if (length == 1) return make_leaf(data[0]);

		int axis = parent.major_axis();
		if (axis != parentAxis) sort_indices(data, length, axis);
		return build_layer(data, length);
	}
};

// ---------------- binary -> quad collapse ----------------

struct QuadSlot // one entry of BuildNode's four-child linked list (QuadBoundingVolumeHierarchy.cs:471-565)
{
	int32_t source; // binary node index, -1 = empty
};

struct QuadBuilder
{
	const std::vector<BinaryNode>& binary;
	std::vector<EchoQbvhNode> nodes;

	explicit QuadBuilder(const std::vector<BinaryNode>& binary) : binary(binary) {}

	// GetChildrenSorted, :551-563
	int children_sorted(int32_t node, int32_t& child0, int32_t& child1) const
	{
		const BinaryNode& n = binary[node];
		int axis = n.axis;
		child0 = n.child0;
		child1 = n.child1;
		if (binary[child0].box.min[axis] > binary[child1].box.min[axis]) std::swap(child0, child1);
		return axis;
	}

	// AddChildren, :517-543: a leaf becomes [leaf, empty] with axis 3
	int add_children(int32_t node, int32_t out[2]) const
	{
		if (binary[node].is_leaf())
		{
			out[0] = node;
			out[1] = -1;
			return 3;
		}

		return children_sorted(node, out[0], out[1]);
	}

	// CreateNode, :363-404 (pre-order: a child's index is claimed before its subtree is emitted)
	void create_node(int32_t source, uint32_t index, int& depth)
	{
		int32_t child0, child1;
		int axisMajor = children_sorted(source, child0, child1);

		int32_t slots[4];
		int axisMinor0 = add_children(child0, slots);
		int axisMinor1 = add_children(child1, slots + 2);

		EchoQbvhNode node;
		std::memset(&node, 0, sizeof(node));
		node.axisMajor = axisMajor;
		node.axisMinor0 = axisMinor0;
		node.axisMinor1 = axisMinor1;

		depth = 0;

		for (int i = 0; i < 4; i++)
		{
			int32_t current = slots[i];
			uint32_t token;
			int nodeDepth;

			Box box = { { kInf, kInf, kInf }, { kInf, kInf, kInf } }; // BoxBound.None, BoxBound.cs:94

			if (current < 0)
			{
				token = ECHO_TOKEN_EMPTY;
				nodeDepth = 0;
			}
			else
			{
				box = binary[current].box;

				if (binary[current].is_leaf())
				{
					token = binary[current].token;
					nodeDepth = 1;
				}
				else
				{
					uint32_t childIndex = (uint32_t)nodes.size();
					nodes.emplace_back();
					create_node(current, childIndex, nodeDepth);
					token = ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_NODE, childIndex);
				}
			}

			node.minX[i] = box.min.x; node.minY[i] = box.min.y; node.minZ[i] = box.min.z;
			node.maxX[i] = box.max.x; node.maxY[i] = box.max.y; node.maxZ[i] = box.max.z;
			node.token4[i] = token;
			depth = std::max(depth, nodeDepth);
		}

		++depth;
		nodes[index] = node;
	}
};

// ---------------- LightTree ----------------

// LightBound / ConeBound arithmetic and LightCollection.CreateBounds live in echo_light_build.h, the one restatement this recursive build,
// the device build (lightbuild.cu) and its CPU emulation share: MathF.Acos / Cos / SinCos and Math.Acos are pinned there, so all three
// emit the same bits.
using LightBound = lightbuild::Bound;

struct TokenizedLight
{
	uint32_t token;
	LightBound bound;
};

struct LightTreeBuilder
{
	std::vector<EchoLightNode> nodes;
	std::vector<LightBound> bounds; // parallel to nodes

	// LightTree.Build, LightTree.cs:62-113. Returns the node index.
	uint32_t build(TokenizedLight* data, int length)
	{
		if (length == 1)
		{
			uint32_t index = push(data[0].bound);
			nodes[index].child0 = ECHO_TOKEN_EMPTY;
			nodes[index].child1 = data[0].token;
			return index;
		}

		float lo[3], hi[3]; // parentBound = bounds[0].content.box, then Encapsulate over every emitter
		for (int k = 0; k < 3; k++) { lo[k] = data[0].bound.lo[k]; hi[k] = data[0].bound.hi[k]; }
		for (int i = 0; i < length; i++)
			for (int k = 0; k < 3; k++) { lo[k] = lightbuild::math_min(lo[k], data[i].bound.lo[k]); hi[k] = lightbuild::math_max(hi[k], data[i].bound.hi[k]); }

		uint32_t majorAxis = lightbuild::major_axis(lo, hi);

		// the reference sorts with Span.Sort (unstable introsort); ties are broken stably here
		std::stable_sort(data, data + length, [majorAxis](const TokenizedLight& a, const TokenizedLight& b)
		{
			return lightbuild::centre(a.bound, majorAxis) < lightbuild::centre(b.bound, majorAxis);
		});

		std::vector<float> costs(length);
		LightBound lightBound = data[length - 1].bound;

		for (int i = length - 2; i >= 0; i--)
		{
			costs[i + 1] = lightbuild::relative_area(lightBound);
			lightBound = lightbuild::encapsulate(lightBound, data[i].bound);
		}

		float minCost = kInf;
		int minIndex = -1;

		lightBound = data[0].bound;

		for (int i = 1; i < length; i++)
		{
			float cost = costs[i] + lightbuild::relative_area(lightBound);

			if (cost < minCost)
			{
				minCost = cost;
				minIndex = i;
			}

			lightBound = lightbuild::encapsulate(lightBound, data[i].bound);
		}

		if (minIndex < 0) minIndex = length / 2; // all costs NaN/inf: the reference would throw; split in the middle instead

		// new Node(Build(bounds[minIndex..]), Build(bounds[..minIndex])): child0 = tail, child1 = head
		uint32_t index = push(LightBound{});
		uint32_t child0 = build(data + minIndex, length - minIndex);
		uint32_t child1 = build(data, minIndex);

		LightBound bound = lightbuild::encapsulate(bounds[child0], bounds[child1]);
		bounds[index] = bound;
		fill(index, bound);
		nodes[index].child0 = child0;
		nodes[index].child1 = child1;
		return index;
	}

	uint32_t push(const LightBound& bound)
	{
		uint32_t index = (uint32_t)nodes.size();
		nodes.emplace_back();
		bounds.push_back(bound);
		fill(index, bound);
		return index;
	}

	void fill(uint32_t index, const LightBound& bound)
	{
		lightbuild::fill_node(nodes[index], bound);
		nodes[index].pad[0] = nodes[index].pad[1] = 0;
	}

	// LightTree ctor AddToMap, LightTree.cs:26-37
	void add_to_map(uint32_t index, int depth, uint64_t branches, std::vector<uint32_t>& tokens, std::vector<uint64_t>& paths) const
	{
		const EchoLightNode& node = nodes[index];

		if (node.child0 == ECHO_TOKEN_EMPTY)
		{
			tokens.push_back(node.child1);
			paths.push_back(branches);
			return;
		}

		add_to_map(node.child0, depth + 1, branches, tokens, paths);
		add_to_map(node.child1, depth + 1, branches | (1ull << depth), tokens, paths);
	}

	int max_depth(uint32_t index) const
	{
		const EchoLightNode& node = nodes[index];
		if (node.child0 == ECHO_TOKEN_EMPTY) return 0;
		return 1 + std::max(max_depth(node.child0), max_depth(node.child1));
	}
};

inline float luminance(const float* c) { return (c[0] * 0.212671f + c[1] * 0.715160f) + (c[2] * 0.072169f + 0.0f * 0.0f); } // RGB128.cs:30-38

template<class T>
T* copy_out(const std::vector<T>& source)
{
	T* result = (T*)std::malloc(sizeof(T) * std::max<size_t>(source.size(), 1));
	if (result && !source.empty()) std::memcpy(result, source.data(), sizeof(T) * source.size());
	return result;
}

inline float sse_min(float a, float b) { return a < b ? a : b; }
inline float sse_max(float a, float b) { return a > b ? a : b; }

Box triangle_box(const EchoTriangle& t) // PreparedTriangle.BoxBound, TriangleEntity.cs:142 (Float4 SSE min/max, no NaN in scope)
{
	Vec3 v0 = v3(t.vertex0), v1 = v0 + v3(t.edge1), v2 = v0 + v3(t.edge2);
	return { { sse_min(sse_min(v0.x, v1.x), v2.x), sse_min(sse_min(v0.y, v1.y), v2.y), sse_min(sse_min(v0.z, v1.z), v2.z) },
	         { sse_max(sse_max(v0.x, v1.x), v2.x), sse_max(sse_max(v0.y, v1.y), v2.y), sse_max(sse_max(v0.z, v1.z), v2.z) } };
}

Box sphere_box(const EchoSphere& s) // PreparedSphere.BoxBound, SphereEntity.cs:66
{
	Vec3 p = v3(s.position);
	return { { p.x - s.radius, p.y - s.radius, p.z - s.radius }, { p.x + s.radius, p.y + s.radius, p.z + s.radius } };
}

} // namespace

extern "C"
{

int32_t echo_host_build_qbvh_instanced(const EchoTriangle* triangles, uint32_t triangleCount, const EchoSphere* spheres, uint32_t sphereCount,
                                       const float* instanceBounds, uint32_t instanceCount, int32_t threads,
                                       EchoQbvhNode** outNodes, uint32_t* outNodeCount, uint32_t* outMaxDepth)
{
	uint64_t total = (uint64_t)triangleCount + sphereCount + instanceCount;
	if (total < 2 || total >= (1u << ECHO_TOKEN_INDEX_BITS) || !outNodes || !outNodeCount || !outMaxDepth) return ECHO_B200_ERR_INVALID;

	// GeometryCollection.CreateBounds, GeometryCollection.cs:52-81: triangles, spheres, then instances
	std::vector<Tokenized> bounds(total);
	for (uint32_t i = 0; i < triangleCount; i++) bounds[i] = { ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_TRIANGLE, i), triangle_box(triangles[i]) };
	for (uint32_t i = 0; i < sphereCount; i++) bounds[triangleCount + i] = { ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_SPHERE, i), sphere_box(spheres[i]) };

	for (uint32_t i = 0; i < instanceCount; i++)
	{
		const float* b = instanceBounds + (size_t)i * 6; // PreparedInstance.BoxBound: min xyz, max xyz
		Box box;
		box.min = { b[0], b[1], b[2] };
		box.max = { b[3], b[4], b[5] };
		bounds[(size_t)triangleCount + sphereCount + i] = { ECHO_TOKEN_MAKE(ECHO_TOKEN_TYPE_INSTANCE, i), box };
	}

	gMaxBuilders = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
	if (gMaxBuilders < 1) gMaxBuilders = 1;
	gActiveBuilders = 1;

	NodePool pool;
	pool.nodes.resize(total * 2);

	SweepBuilder builder(pool);
	int32_t root = builder.build(bounds.data(), (int)total);

	QuadBuilder quad(pool.nodes);
	quad.nodes.reserve(total);
	quad.nodes.emplace_back();
	int depth = 0;
	quad.create_node(root, 0, depth);

	*outNodes = copy_out(quad.nodes);
	*outNodeCount = (uint32_t)quad.nodes.size();
	*outMaxDepth = (uint32_t)depth;
	return *outNodes ? ECHO_B200_OK : ECHO_B200_ERR_INVALID;
}

int32_t echo_host_build_qbvh(const EchoTriangle* triangles, uint32_t triangleCount, const EchoSphere* spheres, uint32_t sphereCount,
                             int32_t threads, EchoQbvhNode** outNodes, uint32_t* outNodeCount, uint32_t* outMaxDepth)
{
	return echo_host_build_qbvh_instanced(triangles, triangleCount, spheres, sphereCount, nullptr, 0, threads, outNodes, outNodeCount, outMaxDepth);
}

float echo_host_emissive_power(const float emission[3]) { return lightbuild::luminance(emission) * kPi; } // Emissive.cs:52-53

int32_t echo_host_build_light_tree(const EchoTriangle* triangles, uint32_t triangleCount, const EchoSphere* spheres, uint32_t sphereCount,
                                   const EchoMaterial* materials, uint32_t materialCount, const EchoPointLight* points, uint32_t pointCount,
                                   EchoLightNode** outNodes, uint32_t* outNodeCount,
                                   uint32_t** outTokens, uint64_t** outPaths, uint32_t* outEmitterCount, float* outPower)
{
	return echo_host_build_light_tree_instanced(triangles, triangleCount, spheres, sphereCount, materials, materialCount, points, pointCount, nullptr, 0,
	                                            outNodes, outNodeCount, outTokens, outPaths, outEmitterCount, outPower);
}

int32_t echo_host_build_light_tree_instanced(const EchoTriangle* triangles, uint32_t triangleCount, const EchoSphere* spheres, uint32_t sphereCount,
                                             const EchoMaterial* materials, uint32_t materialCount, const EchoPointLight* points, uint32_t pointCount,
                                             const float* instanceLights, uint32_t instanceCount,
                                             EchoLightNode** outNodes, uint32_t* outNodeCount,
                                             uint32_t** outTokens, uint64_t** outPaths, uint32_t* outEmitterCount, float* outPower)
{
	if (!outNodes || !outNodeCount || !outTokens || !outPaths || !outEmitterCount || !outPower) return ECHO_B200_ERR_INVALID;

	// LightCollection.CreateBounds, LightCollection.cs:91-137: point lights, emissive triangles, emissive spheres, then the placements
	// (AddInstances, :123-135: PreparedInstance.LightBound; those without power are skipped)
	const lightbuild::Sources sources = { triangles, triangleCount, spheres, sphereCount, materials, materialCount, points, pointCount, instanceLights, instanceCount };
	std::vector<TokenizedLight> lights;

	for (uint32_t c = 0; c < sources.candidates(); c++)
	{
		TokenizedLight light;
		if (lightbuild::emitter(sources, c, light.bound, light.token)) lights.push_back(light);
	}

	LightTreeBuilder builder;
	std::vector<uint32_t> tokens;
	std::vector<uint64_t> paths;
	*outPower = 0.0f;

	if (!lights.empty())
	{
		builder.nodes.reserve(lights.size() * 2);
		builder.bounds.reserve(lights.size() * 2);
		uint32_t root = builder.build(lights.data(), (int)lights.size());
		if (root != 0 || builder.max_depth(0) >= 64) return ECHO_B200_ERR_UNSUPPORTED; // LightTree.cs:29
		builder.add_to_map(0, 0, 0ull, tokens, paths);
		*outPower = builder.nodes[0].power;
	}

	*outNodes = copy_out(builder.nodes);
	*outNodeCount = (uint32_t)builder.nodes.size();
	*outTokens = copy_out(tokens);
	*outPaths = copy_out(paths);
	*outEmitterCount = (uint32_t)tokens.size();
	return ECHO_B200_OK;
}

float echo_host_infinite_threshold(float infinitePower, float scenePower) // PreparedScene.CalculateThreshold, PreparedScene.cs:317-325
{
	const float phi = 1.61803398874989484820f;
	const float bias = phi * phi;
	float sum = infinitePower + scenePower * bias;
	return kEpsilon <= sum ? infinitePower / sum : 1.0f;
}

float echo_host_ambient_power(const float radiance[3], const EchoQbvhNode* root) // AmbientLight.Prepare, AmbientLight.cs:42-51
{
	// The reference takes the radius of a near-minimal bounding sphere (Accelerator.SphereBound); the half diagonal of the
	// root bound is used here instead. Only the infinite-vs-scene light selection probability depends on it.
	Box all = { { kInf, kInf, kInf }, { -kInf, -kInf, -kInf } };

	for (int i = 0; i < 4; i++)
	{
		if (root->token4[i] == ECHO_TOKEN_EMPTY) continue;
		all = all.encapsulate({ { root->minX[i], root->minY[i], root->minZ[i] }, { root->maxX[i], root->maxY[i], root->maxZ[i] } });
	}

	float radius = std::max(magnitude(all.max - all.min) / 2.0f, 1.0f);
	return kPi * radius * radius * luminance(radiance);
}

static float entrance_reflectance(float eta) // Lambertian.cs:175-199
{
	float eta2 = eta * eta;
	float eta4 = eta2 * eta2;

	float eta1a1 = eta + 1.0f, eta2a1 = eta2 + 1.0f, eta4a1 = eta4 + 1.0f;
	float eta1s1 = eta - 1.0f, eta2s1 = eta2 - 1.0f, eta4s1 = eta4 - 1.0f;

	float quotient0 = eta1s1 * std::fma(3.0f, eta, 1.0f) / (6.0f * eta1a1 * eta1a1);
	float quotient1 = eta2 * eta2s1 * eta2s1 / (eta2a1 * eta2a1 * eta2a1);
	float quotient2 = -2.0f * eta2 * eta * (eta2 + eta + eta1s1) / (eta2a1 * eta4s1);
	float quotient3 = 8.0f * eta4 * eta4a1 / (eta2a1 * eta4s1 * eta4s1);

	float fma0 = std::fma(std::log(eta1s1 / eta1a1), quotient1, quotient0);
	float fma1 = std::fma(std::log(eta), quotient3, quotient2);

	return 0.5f + fma0 + fma1;
}

float echo_host_fresnel_diffuse_reflectance(float eta) // Lambertian.cs:168-173
{
	// eta.AlmostEquals(1f), Scalars.cs:153-169
	float difference = std::fabs(eta - 1.0f);
	if (eta == 1.0f || difference < 1E-5f * std::fmin(std::fabs(eta) + 1.0f, std::numeric_limits<float>::max())) return 0.0f;
	if (eta >= 1.0f) return entrance_reflectance(eta);
	float reflectance = entrance_reflectance(1.0f / eta);
	return 1.0f - eta * eta * (1.0f - reflectance);
}

float echo_host_fresnel_diffuse_reflectance_fast(float eta) // Lambertian.cs:202-230
{
	auto entrance = [](float etaR)
	{
		float etaR2 = etaR * etaR, etaR3 = etaR2 * etaR, etaR4 = etaR2 * etaR2, etaR5 = etaR4 * etaR;
		float sum = +0.91932f;
		sum = std::fma(-3.47930f, etaR, sum);
		sum = std::fma(+6.75335f, etaR2, sum);
		sum = std::fma(-7.80989f, etaR3, sum);
		sum = std::fma(+4.98554f, etaR4, sum);
		sum = std::fma(-1.36881f, etaR5, sum);
		return sum;
	};

	if (eta >= 1.0f) return entrance(1.0f / eta);
	float reflectance = entrance(eta);
	return 1.0f - eta * eta * (1.0f - reflectance);
}

void echo_host_free(void* pointer) { std::free(pointer); }

} // extern "C"
