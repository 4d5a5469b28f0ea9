// upg_band.cu -- K1b `band_sample`: isosurface-equivalent sampling of the land and navigation fields.
//
// The reference fills the whole meshing grid through the terrain functor (MarchingCubes::updateByPosition,
// externals/cage/sources/libcore/mesh/marchingCubes.cpp:149-165) and dualmc then reads of every sample only
//   * its SIGN (cell codes `>= iso`, the manifold test: dualmc.tpp:12-32, 37-86), and
//   * its VALUE where it is an end point of a grid edge whose two signs differ (edge interpolation, dualmc.tpp:91-188).
// land = shape - elevation / 10 and navigation = shape - max(elevation, 0) / 10 (sources/modesConfiguration.cpp:248-279) with
// elevation / 10 inside a known interval [lo, hi] (elevation_interval10, upg_planet_host.h), so wherever the shape value lies
// outside [lo, hi] (by a margin) the sign of both fields is the sign the shape alone dictates, and the 4-12 noise octaves of
// the elevation function cannot change what the extraction sees.  This kernel therefore
//   1. reads the shape value of every grid node from the water field (= the shape; K1 wrote it in a first, cheap launch),
//   2. classifies the node as surely-positive / surely-negative / undecided,
//   3. evaluates the elevation EXACTLY (same code, same bits as the dense kernel) for every node that is undecided or has one of
//      its six grid neighbours in a different class -- both end points of every edge that can possibly change sign -- and
//   4. stores `shape - hi` / `shape - lo` (a value of the settled sign, at least `margin` away from zero) everywhere else.
// The meshes extracted from these samples are bit-identical to the ones from the dense field (tests/test_gpu_mesh.py compares
// both paths; tests/test_gpu_field.py checks properties 3 and 4 on the samples themselves for all shapes x elevation modes).
//
// Work distribution: the nodes that need the elevation are a thin shell (3 % of a 1000^3 sphere grid), scattered over the
// warps that scan the rows.  Each block scans 32 grid rows, 1024 nodes (four per thread) at a time, and PUSHES the nodes to evaluate into a
// shared-memory ring (ballot + warp prefix); whenever 256 are queued the whole block evaluates one node per thread, so the
// expensive code always runs with full warps and exists once in the instruction stream.
#include "../../include/upg.h"
#include "upg_internal.h"

namespace upg
{
	namespace
	{
		constexpr int BAND_BLOCK = 256;
		constexpr int BAND_WARPS = BAND_BLOCK / 32;
		constexpr int BAND_ROWS = 32; // grid rows per block
		constexpr int BAND_NODES = 4; // consecutive nodes per thread and scan step (one 16-byte load per row touched)
		constexpr uint32_t BAND_STEP = BAND_BLOCK * BAND_NODES;
		constexpr uint32_t BAND_RING = 2 * BAND_STEP; // < BAND_BLOCK entries survive a drain, one scan step adds <= BAND_STEP
		constexpr uint32_t BAND_MAX_RES = 2048; // upg_planet_create's limit: the coordinate table below

		struct BandLaunch
		{
			const float *water; // device, rows x res floats: the shape value of every node
			float *land; // may be null
			float *navigation; // may be null
			uint32_t rows; // zcount * res
			uint32_t zcount; // z-slices in this launch
			uint32_t z0; // global index of the first one
			uint32_t vector; // rows are 16-byte aligned in all three arrays: float4 loads / stores
			float lo, hi, margin; // elevation / 10 lies in [lo, hi] (containing 0)
			uint32_t *status; // bit 0: NaN sample
			unsigned long long *evaluated; // statistics: nodes whose elevation was evaluated
		};

		// class of a node: +1 / -1 = the shape alone settles the sign of land and navigation, 0 = undecided
		__device__ __forceinline__ int band_class(float w, float lo, float hi, float margin)
		{
			// NaN compares false twice -> undecided -> evaluated exactly -> flagged like the dense kernel would
			return (w - hi > margin) ? 1 : (w - lo < -margin) ? -1 : 0;
		}

		// four consecutive samples of a row starting at node x (a multiple of 4); nodes beyond the row read as `pad`
		__device__ __forceinline__ void band_load4(const float *p, uint32_t x, uint32_t res, bool vector, float pad, float out[BAND_NODES])
		{
			if (vector && x + BAND_NODES <= res)
			{
				const float4 t = *reinterpret_cast<const float4 *>(p);
				out[0] = t.x, out[1] = t.y, out[2] = t.z, out[3] = t.w;
			}
			else
			{
#pragma unroll
				for (uint32_t k = 0; k < BAND_NODES; k++)
					out[k] = x + k < res ? p[k] : pad;
			}
		}

		template<int ELEV>
		__global__ void __launch_bounds__(BAND_BLOCK) band_sample_kernel(const __grid_constant__ PlanetParams P, const __grid_constant__ BandLaunch L)
		{
			__shared__ uint32_t ring[BAND_RING];
			__shared__ float coord[BAND_MAX_RES]; // grid index -> world coordinate (the same on the three axes)
			__shared__ uint32_t warpCount[BAND_WARPS];
			const uint32_t res = (uint32_t)P.grid.res;
			const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
			grad_table_init<BAND_BLOCK>();
			for (uint32_t i = tid; i < res; i += BAND_BLOCK)
				coord[i] = grid_coord(P.grid, i);
			__syncthreads();
			const uint32_t row0 = blockIdx.x * BAND_ROWS;
			const uint32_t rowCount = min((uint32_t)BAND_ROWS, L.rows - row0);
			const uint32_t xSteps = (res + BAND_STEP - 1) / BAND_STEP;
			const uint32_t steps = rowCount * xSteps;
			const float lo = L.lo, hi = L.hi, margin = L.margin;
			const bool vector = L.vector != 0;
			const int64_t plane = (int64_t)res * res;
			uint32_t head = 0, count = 0; // ring state, identical in every thread
			uint32_t evaluated = 0;
			for (uint32_t step = 0; step < steps; step++)
			{
				const uint32_t r = step / xSteps;
				const uint32_t x = (step - r * xSteps) * BAND_STEP + tid * BAND_NODES;
				const uint32_t row = row0 + r;
				const uint32_t zl = row / res;
				const uint32_t y = row - zl * res;
				bool need[BAND_NODES] = { false, false, false, false };
				if (x < res)
				{
					const uint64_t idx = (uint64_t)row * res + x;
					const float *w = L.water + idx;
					float v[BAND_NODES], t[BAND_NODES];
					int c[BAND_NODES];
					bool differs[BAND_NODES];
					band_load4(w, x, res, vector, 0.0f, v);
#pragma unroll
					for (int k = 0; k < BAND_NODES; k++)
						c[k] = band_class(v[k], lo, hi, margin);
					// an edge to a node of another class may change sign: both of its end points are evaluated
					const int left = x > 0 ? band_class(w[-1], lo, hi, margin) : c[0];
					const int right = x + BAND_NODES < res ? band_class(w[BAND_NODES], lo, hi, margin) : c[BAND_NODES - 1];
#pragma unroll
					for (int k = 0; k < BAND_NODES; k++)
					{
						const int l = k == 0 ? left : c[k - 1];
						// the right neighbour of the last node of the row does not exist; c[k + 1] of a padded node is never used
						const int rr = k == BAND_NODES - 1 ? right : (x + k + 1 < res ? c[k + 1] : c[k]);
						differs[k] = l != c[k] || rr != c[k];
					}
					if (y > 0)
					{
						band_load4(w - res, x, res, vector, 0.0f, t);
#pragma unroll
						for (int k = 0; k < BAND_NODES; k++)
							differs[k] |= band_class(t[k], lo, hi, margin) != c[k];
					}
					if (y + 1 < res)
					{
						band_load4(w + res, x, res, vector, 0.0f, t);
#pragma unroll
						for (int k = 0; k < BAND_NODES; k++)
							differs[k] |= band_class(t[k], lo, hi, margin) != c[k];
					}
					if (zl > 0)
					{
						band_load4(w - plane, x, res, vector, 0.0f, t);
#pragma unroll
						for (int k = 0; k < BAND_NODES; k++)
							differs[k] |= band_class(t[k], lo, hi, margin) != c[k];
					}
					if (zl + 1 < L.zcount)
					{
						band_load4(w + plane, x, res, vector, 0.0f, t);
#pragma unroll
						for (int k = 0; k < BAND_NODES; k++)
							differs[k] |= band_class(t[k], lo, hi, margin) != c[k];
					}
					float fill[BAND_NODES];
					bool any = false;
#pragma unroll
					for (int k = 0; k < BAND_NODES; k++)
					{
						need[k] = x + k < res && (c[k] == 0 || differs[k]);
						any |= need[k];
						fill[k] = c[k] > 0 ? v[k] - hi : v[k] - lo;
					}
					if (!any && vector && x + BAND_NODES <= res)
					{
						const float4 f4 = make_float4(fill[0], fill[1], fill[2], fill[3]);
						if (L.land)
							*reinterpret_cast<float4 *>(L.land + idx) = f4;
						if (L.navigation)
							*reinterpret_cast<float4 *>(L.navigation + idx) = f4;
					}
					else
					{
#pragma unroll
						for (int k = 0; k < BAND_NODES; k++)
							if (x + k < res && !need[k])
							{
								if (L.land)
									L.land[idx + k] = fill[k];
								if (L.navigation)
									L.navigation[idx + k] = fill[k];
							}
					}
				}
				// push the nodes to evaluate into the ring
				uint32_t votes[BAND_NODES], mine = 0;
#pragma unroll
				for (int k = 0; k < BAND_NODES; k++)
				{
					votes[k] = __ballot_sync(0xffffffffu, need[k]);
					mine += __popc(votes[k]);
				}
				if (lane == 0)
					warpCount[warp] = mine;
				__syncthreads();
				uint32_t before = 0, total = 0;
#pragma unroll
				for (uint32_t k = 0; k < BAND_WARPS; k++)
				{
					const uint32_t n = warpCount[k];
					before += k < warp ? n : 0u;
					total += n;
				}
				uint32_t slot = head + count + before;
#pragma unroll
				for (int k = 0; k < BAND_NODES; k++)
				{
					if (need[k])
						ring[(slot + __popc(votes[k] & ((1u << lane) - 1u))) & (BAND_RING - 1u)] = r * res + x + k;
					slot += __popc(votes[k]);
				}
				count += total;
				__syncthreads();
				// drain: full blocks of work, plus whatever is left after the last scan step
				const bool last = step + 1 == steps;
				while (count >= BAND_BLOCK || (last && count > 0))
				{
					const uint32_t n = min(count, (uint32_t)BAND_BLOCK);
					if (tid < n)
					{
						const uint32_t e = ring[(head + tid) & (BAND_RING - 1u)];
						const uint32_t er = e / res;
						const uint32_t ex = e - er * res;
						const uint32_t erow = row0 + er;
						const uint32_t ezl = erow / res;
						const uint32_t ey = erow - ezl * res;
						const uint64_t idx = (uint64_t)erow * res + ex;
						const V3 pos = v3(coord[ex], coord[ey], coord[L.z0 + ezl]);
						const float base = L.water[idx];
						// the dense kernel's expressions (upg_field.cuh, FD_ALL)
						const float e10 = elevation_raw<ELEV>(P, pos) / 10.0f;
						const float land = base - e10;
						const float nav = base - fmax_(e10, 0.0f);
						bool bad = false;
						if (L.land)
						{
							L.land[idx] = land;
							bad |= isnan_(land);
						}
						if (L.navigation)
						{
							L.navigation[idx] = nav;
							bad |= isnan_(nav);
						}
						if (bad)
							atomicOr(L.status, 1u);
					}
					head += n;
					count -= n;
					evaluated += n;
				}
			}
			if (L.evaluated && tid == 0 && evaluated)
				atomicAdd(L.evaluated, (unsigned long long)evaluated);
		}

		template<int ELEV>
		void band_launch(const PlanetParams &P, const BandLaunch &L, cudaStream_t stream)
		{
			const uint32_t blocks = (L.rows + BAND_ROWS - 1) / BAND_ROWS;
			band_sample_kernel<ELEV><<<blocks, BAND_BLOCK, 0, stream>>>(P, L);
		}
	}

	bool band_sampling_applies(const Planet *planet)
	{
		// a constant elevation needs no noise at all: the dense kernel is already a plain store
		return planet->sampling == UPG_SAMPLING_BAND && planet->params.elevation != EL_NONE;
	}

	// land / navigation over z-slices [za, zb) given the water field of the same slices; asynchronous on `stream`
	int band_sample(const Planet *planet, int za, int zb, const float *water, float *land, float *navigation, unsigned long long *evaluated, cudaStream_t stream)
	{
		const PlanetParams &P = planet->params;
		const uint32_t res = (uint32_t)P.grid.res;
		if (zb <= za || (!land && !navigation))
			return UPG_OK;
		double lo = 0, hi = 0;
		elevation_interval10(P.elevation, &lo, &hi);
		BandLaunch L;
		L.water = water;
		L.land = land;
		L.navigation = navigation;
		L.zcount = (uint32_t)(zb - za);
		L.rows = L.zcount * res;
		L.z0 = (uint32_t)za;
		L.lo = (float)lo;
		L.hi = (float)hi;
		// the float subtraction of the field and the one of the classification round independently (relative 2^-24 of values up to
		// the box diagonal ~3500 + |hi|): 1/64 is two orders of magnitude above that and ten times the near-zero threshold 1e-3
		L.margin = 0.015625f;
		L.vector = res % 4 == 0 && ((uintptr_t)water & 15) == 0 && ((uintptr_t)land & 15) == 0 && ((uintptr_t)navigation & 15) == 0;
		L.status = planet->status;
		L.evaluated = evaluated;
		switch (P.elevation)
		{
			case EL_SIMPLE:
				band_launch<EL_SIMPLE>(P, L, stream);
				break;
			case EL_LEGACY:
				band_launch<EL_LEGACY>(P, L, stream);
				break;
			case EL_LAKES:
				band_launch<EL_LAKES>(P, L, stream);
				break;
			case EL_ISLANDS:
				band_launch<EL_ISLANDS>(P, L, stream);
				break;
			case EL_CRATERS:
				band_launch<EL_CRATERS>(P, L, stream);
				break;
			default:
				return fail(UPG_ERR_ARGUMENT, "band sampling does not apply to elevation %d", P.elevation);
		}
		const cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess)
			return cuda_fail(e, "band_sample_kernel launch");
		return UPG_OK;
	}
}
