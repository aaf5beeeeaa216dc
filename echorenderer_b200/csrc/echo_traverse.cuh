// echo_traverse.cuh — the persistent, work-replacing QBVH traversal core used by every closest-hit / occlusion kernel.
//
// Why not one thread per ray: in an incoherent batch most rays leave the tree after 1-3 nodes while a few need 30+, so a
// warp that walks 32 fixed rays runs at ~5 active lanes (measured: profiles/r1a, 4.88 threads per instruction). Here one
// resident wave of warps pulls rays from a global counter: each warp reserves a pool of consecutive rays with one global
// atomic, a lane whose ray ended takes the next ray of the pool (work replacement), rays are staged global -> shared
// with cp.async one ray ahead of their use, and primitive tests are postponed until several lanes have one pending, so
// the expensive fp64-cross Möller–Trumbore code is issued for many lanes at once.
//
// Per ray, the sequence of operations is exactly the reference's (QuadBoundingVolumeHierarchy.cs:123-315,
// GeometryCollection.cs:85-171): same visit order, same culling comparisons, leaves intersected in push order. Only the
// interleaving ACROSS rays changes, so results stay bit-identical to the one-thread-per-ray kernels and to the oracle.
#pragma once
#include "echo_instanced.cuh"
#include "echo_scene.cuh"

namespace echo
{

#ifndef ECHO_TRAVERSE_BLOCK
#define ECHO_TRAVERSE_BLOCK 128 // A/B r2af on C2, ECHO_MIN_BLOCKS scaled with it: 64 threads x 14 CTAs equal, 256 x 3 (768 threads per SM) -4 %
#endif
#ifndef ECHO_POOL
#define ECHO_POOL 128 // A/B r2af / r2ag, two runs each: 512 -2 %, 256 (until r2ae) 5 374, 128 5 430 (+1.0 % on C2 and on the secondary batch, C3 +0.5 %), 64 5 423, 32 5 380 Mrays/s
#endif
constexpr int kTraverseBlock = ECHO_TRAVERSE_BLOCK;  // 4 warps per CTA
constexpr unsigned long long kPool = ECHO_POOL;       // rays reserved per global atomic
#ifndef ECHO_LEAF_VOTE
#define ECHO_LEAF_VOTE 8
#endif
#ifndef ECHO_MIN_BLOCKS
#define ECHO_MIN_BLOCKS 7
#endif
#ifndef ECHO_PREFETCH
#define ECHO_PREFETCH 0 // A/B: 1 = prefetch the next node into L1 after the slot scan, 2 = the pending triangle, 3 = both
#endif
#ifndef ECHO_INST_MIN_BLOCKS
#define ECHO_INST_MIN_BLOCKS 6 // resident CTAs per SM asked of the INST instantiations (80 registers), see trace.cu
#endif
#ifndef ECHO_SHARED_STACK
#define ECHO_SHARED_STACK 8 // traversal-stack entries per thread kept in shared memory (0 = all in local memory), see `stack` below
#endif
#ifndef ECHO_ANY_UNORDERED
#define ECHO_ANY_UNORDERED 1 // any-hit traversal takes a node's children as stored (see the visit-order swaps below); 0 = the reference's order
#endif
#ifndef ECHO_LEAF_MAX_WAIT
#define ECHO_LEAF_MAX_WAIT 2 // a lane waits at most two iterations for its primitive test (A/B on C2/C3/C4: +1-2 %)
#endif
// L2 residency hints (A/B in profiles/README.md): the ray stream is read once — marking its cp.async evict-first keeps it from
// pushing the tree out of L2 (C2: 1.9 GB of DRAM traffic per launch against 0.93 GB compulsory) — and node loads can ask to stay.
#ifndef ECHO_RAY_POLICY
#define ECHO_RAY_POLICY 0
#endif
#ifndef ECHO_NODE_POLICY
#define ECHO_NODE_POLICY 0
#endif
// Taking a new ray costs ~35 warp instructions (three IEEE reciprocals, the order bits, the finiteness test) whatever the number of lanes
// that take one, and rays end one or two lanes at a time: on C2 that block was 12 % of all issued instructions at 4.3 active lanes (ncu source
// page, profiles/README.md). ECHO_FETCH_VOTE > 0 holds the idle lanes back until that many are waiting (or ECHO_FETCH_MAX_WAIT iterations passed,
// or nobody in the warp has a ray): fewer, fuller executions of the block against lanes idling a little longer. A/B in profiles/README.md.
#ifndef ECHO_FETCH_VOTE
#define ECHO_FETCH_VOTE 0
#endif
#ifndef ECHO_FETCH_MAX_WAIT
#define ECHO_FETCH_MAX_WAIT 1
#endif
// ECHO_PREPARE_RAYS: closest-hit kernels compute the per-ray constants (the three IEEE reciprocals of Ray.cs:23, the order bits, the
// finiteness flag) for a whole pool of rays at once, 32 lanes at a time, when the warp reserves the pool — into the ray's slot of the HIT
// buffer, which belongs to this launch and is not written until the ray ends — and a lane that takes a ray copies them in with the ray. Taking
// a ray then costs a few loads instead of ~35 instructions issued for the one to four lanes that happen to need a ray in that iteration.
// Measured (r2v): parity green, and NO gain — C2 closest hit 4815 vs 4812 Mrays/s, secondary +1.3 %, C3 equal, C5 -0.7 %, the instanced
// batch -4 %: a tenth fewer issued instructions buys nothing, the kernel waits for sector fetches, not for issue slots. Off.
#ifndef ECHO_PREPARE_RAYS
#define ECHO_PREPARE_RAYS 0
#endif
constexpr int kStagedFloat4 = 3; // staging slot per thread: the 32-byte ray + its prepared constants
constexpr int kLeafVote = ECHO_LEAF_VOTE;                      // run the primitive tests once this many lanes have one pending

// BoxBound4.Intersect for one lane with hardware min/max. Bit-identical to slab() whenever no operand is NaN, which holds
// when the three reciprocal direction components and the origin are finite (a NaN needs 0 * inf or inf - inf); the sign
// of a zero result may differ, which no comparison downstream can observe.
ECHO_DEVICE float slab_finite(float minX, float minY, float minZ, float maxX, float maxY, float maxZ, vec3 origin, vec3 directionR)
{
	float x0 = (minX - origin.x) * directionR.x, x1 = (maxX - origin.x) * directionR.x;
	float y0 = (minY - origin.y) * directionR.y, y1 = (maxY - origin.y) * directionR.y;
	float z0 = (minZ - origin.z) * directionR.z, z1 = (maxZ - origin.z) * directionR.z;

	float far = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
	float near = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));

	far *= 1.00000024f;
	return ((far >= near) & (far >= 0.0f)) ? near : kInfinity;
}

// one 32-byte sector per lane per instruction (LDG.E.256, sm_100+): a 128-byte node costs 4 sector reads instead of the 8
// half-sector reads of LDG.128 — the L1 data pipe was the busiest unit of the first persistent kernel (78 %, profiles/r1c).
struct __align__(32) float8
{
	float v[8];
};

ECHO_DEVICE float8 ldg256(const void* pointer)
{
	float8 r;
	asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		: "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(pointer));
	return r;
}

ECHO_DEVICE float8 ldg256_hint(const void* pointer, unsigned long long policy)
{
	float8 r;
	asm volatile("ld.global.nc.L2::cache_hint.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
		: "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(pointer), "l"(policy));
	return r;
}

ECHO_DEVICE bool finite_bits(float v) { return (__float_as_uint(v) & 0x7F800000u) != 0x7F800000u; }

// The work counter of a persistent launch is a pair: [0] = the next unreserved ray, [1] = CTAs that have left. Every kernel
// built on persistent_traverse ends with this call; the last CTA to leave puts both words back to zero, so the pair can serve
// the next launch on the same stream with no memset in between and no host-side bookkeeping of which launch still owns it
// (trace.cu ray_counters: one pair per device and stream, launches on one stream are ordered).
ECHO_DEVICE void persistent_finish(unsigned long long* __restrict__ nextRay)
{
	__syncthreads();

	if (threadIdx.x == 0)
	{
		__threadfence();

		if (atomicAdd(nextRay + 1, 1ull) == (unsigned long long)gridDim.x - 1ull)
		{
			nextRay[0] = 0ull;
			nextRay[1] = 0ull;
			__threadfence();
		}
	}
}

// IO concept:
//   const float4* ray_pointer(uint32_t index)   -> 32 contiguous bytes: origin.xyz direction.x | direction.yz limit ignore
//   void store_closest(uint32_t index, bool hit, uint32_t token, float distance, vec2 uv, float limit)
//   void store_any(uint32_t index, bool occluded)
//   float4* prepared_pointer(uint32_t index)    -> 16 bytes of scratch owned by this launch until ray `index` stores its result (closest-hit
//                                                  kernels: the ray's slot of the hit buffer); never called by any-hit kernels
//
// Control flow is "if-if": one warp-wide loop whose body is a fixed sequence of predicated stages, so all 32 lanes
// reconverge at every stage (a per-lane while loop with continue/break leaves the lanes of a warp scattered over the
// loop body: measured 5.2 active threads per instruction, profiles/r1b). Stages per iteration:
//   E  finish + fetch  lanes whose stack ran empty store their result, take the ray staged for them in shared memory and
//                      stage the next ray of the warp's pool with cp.async (global -> shared, L1 bypassed), so the DRAM
//                      latency of the ray stream never stalls the warp
//   B  node visit      lanes whose current node has no slots left pop the next un-culled node and run the 4 slab tests
//   C  slot scan       the Push calls of that node in reference order, up to the first primitive (which becomes pending)
//   D  primitive test  only when enough lanes have a primitive pending (or nobody can do anything else), then C again
//
// INST = true adds instanced packs (echo_instanced.cuh): a TokenType.Instance leaf is "tested" in stage D by saving a frame
// (parent ray, node, slot position, stack base), moving the ray into the pack's space and making the pack's root the top of
// the stack; when a lane's part of the stack above its base runs empty in stage B the frame is popped and the suspended node
// visit is redone from the saved slot. The IO then also provides
//   uint32_t load_ignore_layers(uint32_t index, uint32_t* tokens)             -> count of the ignore hierarchy's instance layers
//   void store_hit_layers(uint32_t index, bool hit, const uint32_t* tokens, uint32_t count)
// INST = false compiles to exactly the single-pack kernel.
template<int STACK, bool ANY, bool INST = false, class IO>
ECHO_DEVICE void persistent_traverse(const DeviceScene& scene, IO& io, uint32_t count, unsigned long long* __restrict__ nextRay, float4* stagedRays)
{
	const unsigned int lane = threadIdx.x & 31u;
	const unsigned int lanesBelow = (1u << lane) - 1u;

#if ECHO_RAY_POLICY
	unsigned long long rayPolicy;
	asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(rayPolicy));
#endif
#if ECHO_NODE_POLICY
	unsigned long long nodePolicy;
	asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(nodePolicy));
#endif

	constexpr bool PREPARE = ECHO_PREPARE_RAYS != 0 && !ANY;
	float4* stagedSlot = stagedRays + threadIdx.x * kStagedFloat4;
	const uint32_t stagedAddress = (uint32_t)__cvta_generic_to_shared(stagedSlot);

	// warp-uniform pool of reserved ray indices [poolNext, poolEnd). Pool size: kPool for big batches; for small ones (late
	// wavefront bounces) just enough that every resident warp gets a share, never less than one ray per lane.
	uint32_t poolNext = 0u, poolEnd = 0u;
	bool exhausted = false;
	const uint32_t warps = gridDim.x * (blockDim.x >> 5);
	uint32_t poolSize = (count / (warps * 4u) + 31u) & ~31u;
	poolSize = poolSize < 32u ? 32u : (poolSize > (uint32_t)kPool ? (uint32_t)kPool : poolSize);

	// per-lane ray state
	bool haveRay = false, staged = false;
	uint32_t rayIndex = 0u, stagedIndex = 0u;
	vec3 origin = { 0, 0, 0 }, direction = { 0, 0, 0 }, directionR = { 0, 0, 0 };
	uint32_t orders = 0u, ignore = ECHO_TOKEN_EMPTY, bestToken = ECHO_TOKEN_EMPTY;
	float limit = 0.0f, best = 0.0f; // best: TraceQuery.distance (closest) / OccludeQuery.travel (any)
	vec2 bestUV = { 0.0f, 0.0f };
	bool finite = true;

	// per-lane node state: entry distances and tokens of the current node's children IN VISIT ORDER, scan position 0..4
	float hit0 = 0, hit1 = 0, hit2 = 0, hit3 = 0;
	uint32_t child0 = 0, child1 = 0, child2 = 0, child3 = 0;
	int position = 4, next = 0;
	uint32_t leaf = ECHO_TOKEN_EMPTY;

	// Traversal stack {token, entry distance bits}. The newest entry lives in registers (`top`): the nearest child pushed by
	// a node visit is popped by the very next one, so most pushes and pops never touch local memory (per-thread local
	// arrays are interleaved at 4-byte granularity: every 8-byte entry costs two 32-byte sectors in the L1 data pipe).
	// The next ECHO_SHARED_STACK entries live in shared memory, only deeper ones in local memory: local memory is cached in L1,
	// where the node stream keeps evicting it — the pop's wait for its local load was 14 % of all stall samples of the C2 kernel,
	// as much as the wait for the node itself (SASS-level ncu, profiles/README.md). Column layout: entry k of thread t at
	// [k * blockDim + t], consecutive lanes on consecutive 8-byte words.
	uint2 stack[STACK];
	uint2 top = make_uint2(0u, 0u);
	bool haveTop = false;
#if ECHO_SHARED_STACK > 0
	__shared__ uint2 sharedStack[ECHO_SHARED_STACK * kTraverseBlock];
	uint2* const sharedColumn = sharedStack + threadIdx.x;
#endif

	auto spill = [&](uint2 entry) // stack[next++] = entry
	{
#if ECHO_SHARED_STACK > 0
		if (next < ECHO_SHARED_STACK) sharedColumn[next * kTraverseBlock] = entry;
		else stack[next - ECHO_SHARED_STACK] = entry;
#else
		stack[next] = entry;
#endif
		++next;
	};

	auto unspill = [&]() -> uint2 // stack[--next]
	{
		--next;
#if ECHO_SHARED_STACK > 0
		if (next < ECHO_SHARED_STACK) return sharedColumn[next * kTraverseBlock];
		return stack[next - ECHO_SHARED_STACK];
#else
		return stack[next];
#endif
	};

	// instancing state of the lane's ray (INST only)
	const PackView rootPack = INST ? load_pack(scene, 0u) : PackView{ 0u, 0u, 0u, 0u };
	PackView pack = rootPack;
	uint32_t level = 0u, packIndex = 0u, currentNode = 0u, ignoreCount = 0u, hitCount = 0u;
	int base = 0, resume = 0;   // stack base of the current layer; slot at which a resumed node visit continues
	bool ignoreHere = true;     // query.ignore's instance layers == query.current's
	InstanceFrame frames[INST ? ECHO_MAX_INSTANCE_LAYERS : 1];
	uint32_t current[INST ? ECHO_MAX_INSTANCE_LAYERS : 1], ignoreLayers[INST ? ECHO_MAX_INSTANCE_LAYERS : 1], hitLayers[INST ? ECHO_MAX_INSTANCE_LAYERS : 1];

	auto set_ray = [&](vec3 newOrigin, vec3 newDirection)
	{
		origin = newOrigin;
		direction = newDirection;
		directionR = { rcp(direction.x), rcp(direction.y), rcp(direction.z) }; // Ray.cs:23
		orders = (directionR.x > 0.0f ? 1u : 0u) | (directionR.y > 0.0f ? 2u : 0u) | (directionR.z > 0.0f ? 4u : 0u) | 8u;
		finite = finite_bits(directionR.x) && finite_bits(directionR.y) && finite_bits(directionR.z)
			&& finite_bits(origin.x) && finite_bits(origin.y) && finite_bits(origin.z);
	};

	// The Push calls of the current node, in order, up to the first primitive (:200-216 / :296-312). `best` only changes in
	// a primitive test, so the remaining slots are classified at once: bit k of `valid` = slot k passes the distance test,
	// `nodes` = it is a branch, `leaves` = it is a primitive that is not the ignored triangle (GeometryCollection.cs:93-94).
	auto scan_slots = [&]()
	{
		uint32_t pending = 0xFu << position & 0xFu;
		uint32_t valid = ((!(hit0 >= best)) ? 1u : 0u) | ((!(hit1 >= best)) ? 2u : 0u) | ((!(hit2 >= best)) ? 4u : 0u) | ((!(hit3 >= best)) ? 8u : 0u);
		uint32_t nodes = (token_type(child0) == 0u ? 1u : 0u) | (token_type(child1) == 0u ? 2u : 0u) | (token_type(child2) == 0u ? 4u : 0u) | (token_type(child3) == 0u ? 8u : 0u);
		uint32_t ignored = (child0 == ignore && token_type(child0) == ECHO_TOKEN_TYPE_TRIANGLE ? 1u : 0u) | (child1 == ignore && token_type(child1) == ECHO_TOKEN_TYPE_TRIANGLE ? 2u : 0u)
			| (child2 == ignore && token_type(child2) == ECHO_TOKEN_TYPE_TRIANGLE ? 4u : 0u) | (child3 == ignore && token_type(child3) == ECHO_TOKEN_TYPE_TRIANGLE ? 8u : 0u);
		if (INST && !ignoreHere) ignored = 0u;

		valid &= pending;
		uint32_t leaves = valid & ~nodes & ~ignored;
		uint32_t first = leaves ? (uint32_t)(__ffs(leaves) - 1) : 4u; // first primitive to test, in visit order
		uint32_t pushes = valid & nodes & ((1u << first) - 1u);            // branches pushed before it

#define ECHO_PUSH(bit, childK, hitK)                                   \
		if (pushes & (bit))                                            \
		{                                                              \
			ECHO_CHECK(scene, next < STACK, CHECK_STACK);              \
			if (haveTop) spill(top);                          \
			top = make_uint2((childK), __float_as_uint(hitK));         \
			haveTop = true;                                            \
		}

		ECHO_PUSH(1u, child0, hit0)
		ECHO_PUSH(2u, child1, hit1)
		ECHO_PUSH(4u, child2, hit2)
		ECHO_PUSH(8u, child3, hit3)
#undef ECHO_PUSH

		if (first < 4u) leaf = first == 0u ? child0 : (first == 1u ? child1 : (first == 2u ? child2 : child3));
		position = first < 4u ? (int)first + 1 : 4;

#if ECHO_PREFETCH & 1
		// the node this lane pops next is known now, a whole iteration before its 128-byte line is needed
		if (pushes != 0u) asm volatile("prefetch.global.L1 [%0];" :: "l"(scene.nodes + ((size_t)(INST ? pack.nodeOffset : 0u) + token_index(top.x)) * 8));
#endif
#if ECHO_PREFETCH & 2
		// so is the primitive it tests next (it waits for the vote)
		if (first < 4u && token_type(leaf) == ECHO_TOKEN_TYPE_TRIANGLE)
			asm volatile("prefetch.global.L1 [%0];" :: "l"(scene.triHot + ((size_t)(INST ? pack.triangleOffset : 0u) + token_index(leaf)) * 3));
#endif
	};

	int waited = 0; // warp-uniform: iterations some lane has been waiting for its primitive test
#if ECHO_FETCH_VOTE > 0
	int fetchWaited = 0; // warp-uniform: iterations some idle lane has been waiting to take its staged ray
#endif

	while (true)
	{
		// ---- E: finish rays whose traversal ran out of work ----
		if (haveRay && leaf == ECHO_TOKEN_EMPTY && position == 4 && next == 0 && !haveTop && (!INST || level == 0u))
		{
			if (ANY) io.store_any(rayIndex, false);
			else
			{
				io.store_closest(rayIndex, best < limit, bestToken, best, bestUV, limit);
				if constexpr (INST) io.store_hit_layers(rayIndex, best < limit, hitLayers, hitCount);
			}

			haveRay = false;
		}

		// ---- E: idle lanes take their staged ray (already in shared memory) ----
#if ECHO_FETCH_VOTE > 0
		const unsigned int fetchReady = __ballot_sync(0xFFFFFFFFu, !haveRay && staged);
		const unsigned int stillWorking = __ballot_sync(0xFFFFFFFFu, haveRay);
		fetchWaited = fetchReady != 0u ? fetchWaited + 1 : 0;
		const bool fetchNow = __popc(fetchReady) >= ECHO_FETCH_VOTE || fetchWaited > ECHO_FETCH_MAX_WAIT || stillWorking == 0u;
		if (fetchNow) fetchWaited = 0;
		if (fetchNow && !haveRay && staged)
#else
		if (!haveRay && staged)
#endif
		{
			asm volatile("cp.async.wait_all;" ::: "memory");
			float4 a = stagedSlot[0], b = stagedSlot[1];
			staged = false;

			rayIndex = stagedIndex;
			origin = { a.x, a.y, a.z };
			direction = { a.w, b.x, b.y };
			limit = b.z;
			ignore = __float_as_uint(b.w);
			best = limit;
			bestToken = ECHO_TOKEN_EMPTY;

			if constexpr (INST)
			{
				ignoreCount = io.load_ignore_layers(rayIndex, ignoreLayers);
				ignoreHere = ignoreCount == 0u;
				level = 0u;
				base = 0;
				resume = 0;
				packIndex = 0u;
				pack = rootPack;
				hitCount = 0u;
			}

			if (!positive(limit)) // PreparedScene.Trace / Occlude guard, PreparedScene.cs:69,84
			{
				if (ANY) io.store_any(rayIndex, false);
				else
				{
					io.store_closest(rayIndex, false, ECHO_TOKEN_EMPTY, limit, bestUV, limit);
					if constexpr (INST) io.store_hit_layers(rayIndex, false, hitLayers, 0u);
				}
			}
			else
			{
				if constexpr (PREPARE)
				{
					float4 prepared = stagedSlot[2]; // computed with the rest of the ray's pool, see the pool reservation below
					directionR = { prepared.x, prepared.y, prepared.z };
					orders = __float_as_uint(prepared.w) & 15u;
					finite = (__float_as_uint(prepared.w) & 16u) != 0u;
				}
				else
				{
					directionR = { rcp(direction.x), rcp(direction.y), rcp(direction.z) }; // Ray.cs:23
					orders = (directionR.x > 0.0f ? 1u : 0u) | (directionR.y > 0.0f ? 2u : 0u) | (directionR.z > 0.0f ? 4u : 0u) | 8u;
					finite = finite_bits(directionR.x) && finite_bits(directionR.y) && finite_bits(directionR.z)
						&& finite_bits(origin.x) && finite_bits(origin.y) && finite_bits(origin.z);
				}

				haveRay = true;
				top = make_uint2(0u, 0u); // NewNodeToken(0), entry distance 0
				haveTop = true;
				next = 0;
				position = 4;
			}
		}

		// ---- E: lanes with an empty staging slot reserve the next ray of the pool and start its copy ----
		unsigned int empty = __ballot_sync(0xFFFFFFFFu, !staged);

		if (empty != 0u && !(exhausted && poolNext >= poolEnd))
		{
			uint32_t wanted = (uint32_t)__popc(empty);
			uint32_t rank = (uint32_t)__popc(empty & lanesBelow);
			uint32_t available = poolEnd - poolNext;
			uint32_t index = poolNext + rank;
			bool got = !staged && rank < available;

			if (available < wanted && !exhausted)
			{
				// the pool cannot serve everyone: reserve the next kPool rays with one global atomic
				unsigned long long base64 = 0ull;
				if (lane == 0u) base64 = atomicAdd(nextRay, (unsigned long long)poolSize);
				base64 = __shfl_sync(0xFFFFFFFFu, base64, 0);
				exhausted = base64 >= (unsigned long long)count;

				uint32_t base = exhausted ? count : (uint32_t)base64;
				uint32_t end = (unsigned long long)base + poolSize < (unsigned long long)count ? base + poolSize : count;

				if (!staged && !got)
				{
					index = base + (rank - available);
					got = index < end;
				}

				poolNext = base + (wanted - available);
				if (poolNext > end) poolNext = end;
				poolEnd = end;

				if constexpr (PREPARE)
				{
					// the whole new pool, every lane busy: ray k -> {1 / direction, order bits | finite << 4} in its hit slot
					for (uint32_t k = base + lane; k < end; k += 32u)
					{
						const float4* source = io.ray_pointer(k);
						float4 a = __ldcg(source), b = __ldcg(source + 1); // L1 bypassed: the rays themselves are staged through L2 later
						vec3 r = { rcp(a.w), rcp(b.x), rcp(b.y) }; // Ray.cs:23
						uint32_t flags = (r.x > 0.0f ? 1u : 0u) | (r.y > 0.0f ? 2u : 0u) | (r.z > 0.0f ? 4u : 0u) | 8u;
						if (finite_bits(r.x) && finite_bits(r.y) && finite_bits(r.z) && finite_bits(a.x) && finite_bits(a.y) && finite_bits(a.z)) flags |= 16u;
						__stcg(io.prepared_pointer(k), make_float4(r.x, r.y, r.z, __uint_as_float(flags)));
					}

					__threadfence(); // the cp.async of ANY lane of this warp may fetch what another lane just wrote
					__syncwarp();
				}
			}
			else poolNext += wanted < available ? wanted : available;

			if (got)
			{
				const float4* source = io.ray_pointer(index);
#if ECHO_RAY_POLICY
				asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(stagedAddress), "l"(source), "l"(rayPolicy) : "memory");
				asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" :: "r"(stagedAddress + 16u), "l"(source + 1), "l"(rayPolicy) : "memory");
#else
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(stagedAddress), "l"(source) : "memory");
				asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(stagedAddress + 16u), "l"(source + 1) : "memory");
#endif
				if constexpr (PREPARE) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(stagedAddress + 32u), "l"(io.prepared_pointer(index)) : "memory");
				asm volatile("cp.async.commit_group;" ::: "memory");
				stagedIndex = index;
				staged = true;
			}
		}

		if (__ballot_sync(0xFFFFFFFFu, haveRay || staged) == 0u) break; // `staged` is only false for good once the pool ran dry

		// ---- B: node visit ----
		bool visit = false;
		uint32_t nodeToken = 0u;

		if (haveRay && leaf == ECHO_TOKEN_EMPTY && position == 4)
		{
			while (haveTop || next > (INST ? base : 0)) // pop; skip entries the closest hit has already passed (:144-146)
			{
				uint2 entry = top;
				if (haveTop) haveTop = false;
				else entry = unspill();

				if (ANY || !(__uint_as_float(entry.y) >= best))
				{
					nodeToken = entry.x;
					visit = true;
					break;
				}
			}

			if (INST && !visit && level > 0u)
			{
				// the instanced pack's accelerator returned: back to the parent's space (PreparedInstance.cs:58-60 / :78-80),
				// then the rest of the parent's suspended node visit
				const InstanceFrame& frame = frames[--level];
				set_ray(frame.origin, frame.direction);
				best = ANY ? frame.travel : best * __ldg(instance_data(scene, frame.instance) + 6).y;
				packIndex = frame.pack;
				pack = load_pack(scene, packIndex);
				base = (int)frame.base;
				nodeToken = frame.node;
				resume = (int)frame.position;
				ignoreHere = layers_match(ignoreLayers, ignoreCount, current, level);
				visit = true;
			}
		}

		if (visit)
		{
			ECHO_CHECK(scene, (INST ? pack.nodeOffset : 0u) + token_index(nodeToken) < scene.nodeCount, CHECK_NODE);
			const float4* nodeData = scene.nodes + ((size_t)(INST ? pack.nodeOffset : 0u) + token_index(nodeToken)) * 8;
#if ECHO_NODE_POLICY
			float8 q0 = ldg256_hint(nodeData + 0, nodePolicy), q1 = ldg256_hint(nodeData + 2, nodePolicy), q2 = ldg256_hint(nodeData + 4, nodePolicy), q3 = ldg256_hint(nodeData + 6, nodePolicy);
#else
			float8 q0 = ldg256(nodeData + 0), q1 = ldg256(nodeData + 2), q2 = ldg256(nodeData + 4), q3 = ldg256(nodeData + 6);
#endif
			if (INST) currentNode = nodeToken;
			// q0 = minX[4] minY[4], q1 = minZ[4] maxX[4], q2 = maxY[4] maxZ[4], q3 = axisMajor axisMinor0 axisMinor1 token4[4] pad

			float t0, t1, t2, t3;

			if (finite)
			{
				t0 = slab_finite(q0.v[0], q0.v[4], q1.v[0], q1.v[4], q2.v[0], q2.v[4], origin, directionR);
				t1 = slab_finite(q0.v[1], q0.v[5], q1.v[1], q1.v[5], q2.v[1], q2.v[5], origin, directionR);
				t2 = slab_finite(q0.v[2], q0.v[6], q1.v[2], q1.v[6], q2.v[2], q2.v[6], origin, directionR);
				t3 = slab_finite(q0.v[3], q0.v[7], q1.v[3], q1.v[7], q2.v[3], q2.v[7], origin, directionR);
			}
			else
			{
				t0 = slab(q0.v[0], q0.v[4], q1.v[0], q1.v[4], q2.v[0], q2.v[4], origin, directionR);
				t1 = slab(q0.v[1], q0.v[5], q1.v[1], q1.v[5], q2.v[1], q2.v[5], origin, directionR);
				t2 = slab(q0.v[2], q0.v[6], q1.v[2], q1.v[6], q2.v[2], q2.v[6], origin, directionR);
				t3 = slab(q0.v[3], q0.v[7], q1.v[3], q1.v[7], q2.v[3], q2.v[7], origin, directionR);
			}

			uint32_t token0 = __float_as_uint(q3.v[3]), token1 = __float_as_uint(q3.v[4]), token2 = __float_as_uint(q3.v[5]), token3 = __float_as_uint(q3.v[6]);

			// Visit order (QuadBoundingVolumeHierarchy.cs:151-198) as three conditional swaps: inside the first pair when
			// orders[axisMinor0], inside the second pair when orders[axisMinor1], and the pairs themselves when orders[axisMajor].
			bool swap0 = (orders >> __float_as_int(q3.v[1])) & 1u;
			bool swap1 = (orders >> __float_as_int(q3.v[2])) & 1u;
			bool swapPairs = (orders >> __float_as_int(q3.v[0])) & 1u;
#if ECHO_ANY_UNORDERED
			// An occlusion query's answer does not depend on the visit order (OccludeImpl returns on ANY hit within travel), so the
			// any-hit kernels take the children as stored and skip the three order swaps: the same flags bit for bit (the whole GPU
			// suite runs on this build), +1.9 % on the C2 occlusion batch, +11 % on the secondary-ray one, +1.6 % on C3 (r2f A/B). The
			// visit COUNTERS of the roofline come from the one-thread-per-query kernels, which keep the reference's order.
			if (ANY) swap0 = swap1 = swapPairs = false;
#endif

			float a0 = swap0 ? t1 : t0, a1 = swap0 ? t0 : t1, b0 = swap1 ? t3 : t2, b1 = swap1 ? t2 : t3;
			uint32_t c0 = swap0 ? token1 : token0, c1 = swap0 ? token0 : token1, d0 = swap1 ? token3 : token2, d1 = swap1 ? token2 : token3;

			hit0 = swapPairs ? b0 : a0; hit1 = swapPairs ? b1 : a1; hit2 = swapPairs ? a0 : b0; hit3 = swapPairs ? a1 : b1;
			child0 = swapPairs ? d0 : c0; child1 = swapPairs ? d1 : c1; child2 = swapPairs ? c0 : d0; child3 = swapPairs ? c1 : d1;
			position = INST ? resume : 0;
			resume = 0;
		}

		// ---- C: slot scan ----
		if (haveRay && leaf == ECHO_TOKEN_EMPTY && position < 4) scan_slots();

		// ---- D: primitive tests, once enough lanes wait for one (or no lane could use another node visit instead) ----
		unsigned int pending = __ballot_sync(0xFFFFFFFFu, leaf != ECHO_TOKEN_EMPTY);
		unsigned int working = __ballot_sync(0xFFFFFFFFu, haveRay);

		waited = pending != 0u ? waited + 1 : 0;

		if (pending != 0u && (__popc(pending) >= kLeafVote || pending == working || waited > ECHO_LEAF_MAX_WAIT))
		{
			waited = 0;

			if (leaf != ECHO_TOKEN_EMPTY)
			{
				bool occluded = false;

				if (INST && token_type(leaf) == ECHO_TOKEN_TYPE_INSTANCE)
				{
					if (level < ECHO_MAX_INSTANCE_LAYERS)
					{
						// query.current.Push(token); instances[token.Index].Trace / Occlude (GeometryCollection.cs:123-131,160-168)
						uint32_t instance = pack.instanceOffset + token_index(leaf);
						ECHO_CHECK(scene, instance < scene.instanceCount, CHECK_INSTANCE);
						const float4* data = instance_data(scene, instance);
						float4 scales = __ldg(data + 6);

						InstanceFrame& frame = frames[level];
						frame.origin = origin;
						frame.direction = direction;
						frame.travel = best;
						frame.node = currentNode;
						frame.position = (uint32_t)position;
						frame.pack = packIndex;
						frame.instance = instance;
						current[level++] = leaf;

						// TransformForward + the scaled distance, PreparedInstance.cs:51,105-111
						vec3 localOrigin = multiply_point(data, origin);
						vec3 localDirection = multiply_direction(data, direction) * scales.y;
						set_ray(localOrigin, localDirection);
						best *= scales.x;

						packIndex = __float_as_uint(scales.z);
						pack = load_pack(scene, packIndex);
						ECHO_CHECK(scene, next < STACK, CHECK_STACK);
						if (haveTop) spill(top); // the parent's entries all live below the new base
						frame.base = (uint32_t)base;
						base = next;
						top = make_uint2(0u, 0u); // the pack's NewNodeToken(0), entry distance 0
						haveTop = true;
						position = 4; // the rest of this node's slots wait in the frame
						ignoreHere = layers_match(ignoreLayers, ignoreCount, current, level);
					}
				}
				else if (token_type(leaf) == ECHO_TOKEN_TYPE_TRIANGLE)
				{
					ECHO_CHECK(scene, (INST ? pack.triangleOffset : 0u) + token_index(leaf) < scene.triangleCount, CHECK_TRIANGLE);
					const float4* data = scene.triHot + ((size_t)(INST ? pack.triangleOffset : 0u) + token_index(leaf)) * 3;
					float4 a = __ldg(data), b = __ldg(data + 1), c = __ldg(data + 2);

					if (ANY) occluded = triangle_occlude({ a.x, a.y, a.z }, { b.x, b.y, b.z }, { c.x, c.y, c.z }, origin, direction, best);
					else
					{
						vec2 uv;
						float d = triangle_intersect({ a.x, a.y, a.z }, { b.x, b.y, b.z }, { c.x, c.y, c.z }, origin, direction, uv);

						if (!(d >= best)) // GeometryCollection.cs:99
						{
							best = d;
							bestToken = leaf;
							bestUV = uv;

							if (INST)
							{
								hitCount = level;
								for (uint32_t k = 0; k < level; k++) hitLayers[k] = current[k];
							}
						}
					}
				}
				else
				{
					ECHO_CHECK(scene, (INST ? pack.sphereOffset : 0u) + token_index(leaf) < scene.sphereCount, CHECK_SPHERE);
					float4 sphere = __ldg(scene.spheres + (INST ? pack.sphereOffset : 0u) + token_index(leaf));
					bool findFar = leaf == ignore && (!INST || ignoreHere);

					if (ANY) occluded = sphere_occlude(sphere, origin, direction, best, findFar);
					else
					{
						vec2 uv;
						float d = sphere_intersect(sphere, origin, direction, uv, findFar);

						if (!(d >= best)) // GeometryCollection.cs:115
						{
							best = d;
							bestToken = leaf;
							bestUV = uv;

							if (INST)
							{
								hitCount = level;
								for (uint32_t k = 0; k < level; k++) hitLayers[k] = current[k];
							}
						}
					}
				}

				leaf = ECHO_TOKEN_EMPTY;

				if (ANY && occluded)
				{
					io.store_any(rayIndex, true);
					haveRay = false;
					haveTop = false;
					position = 4;
					next = 0;
					if (INST) { level = 0u; base = 0; }
				}
				else if (position < 4) scan_slots(); // the rest of this node's Push calls
			}
		}
	}
}

} // namespace echo
