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
#include <cstdlib>

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
			const float2 *rowMinMax; // optional: (min, max) of every row of `water` (shape_rows_kernel); NaN when the row holds one
			float *land; // may be null
			float *navigation; // may be null
			uint32_t rows; // zcount * res
			uint32_t zcount; // z-slices in this launch
			uint32_t z0; // global index of the first one
			uint32_t vector; // rows are 16-byte aligned in all three arrays: float4 loads / stores
			uint32_t fill; // store `shape - hi` / `shape - lo` on the nodes whose elevation is not evaluated (0: leave them untouched)
			// optional bit fields of both outputs (SignBits, upg_internal.h), zero-filled by the caller; bitPitch = 0: none
			uint32_t *landSigns, *landNear, *navSigns, *navNear;
			uint32_t bitPitch;
			float nearTol;
			float lo, hi, margin; // elevation / 10 lies in [lo, hi] (containing 0)
			uint32_t *status; // bit 0: NaN sample
			unsigned long long *evaluated; // statistics: nodes whose elevation was evaluated
		};

		// class of a node: +1 / -1 = the shape alone settles the sign of land and navigation, 0 = undecided.  A function of the node's
		// own value only, so every thread that looks at a node sees the same class.
		__device__ __forceinline__ int band_class(float w, float above, float below)
		{
			// NaN compares false twice -> undecided -> evaluated exactly -> flagged like the dense kernel would
			return w > above ? 1 : w < below ? -1 : 0;
		}
		// three-input minimum / maximum that return NaN when any input is NaN (FMNMX3.NAN)
		__device__ __forceinline__ float band_min3(float a, float b, float c)
		{
			float r;
			asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
			return r;
		}
		__device__ __forceinline__ float band_max3(float a, float b, float c)
		{
			float r;
			asm("max.NaN.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
			return r;
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
			const float lo = L.lo, hi = L.hi;
			const float above = L.hi + L.margin, below = L.lo - L.margin; // class thresholds
			const bool vector = L.vector != 0;
			const int64_t plane = (int64_t)res * res;
			uint32_t head = 0, count = 0; // ring state, identical in every thread
			uint32_t evaluated = 0;
			// the scan position advances without divisions: r = row inside the block, xs = 1024-node step inside the row
			uint32_t r = 0, xs = 0;
			uint32_t zl = row0 / res;
			uint32_t y = row0 - zl * res;
			for (uint32_t step = 0; step < steps; step++)
			{
				const uint32_t x = xs * BAND_STEP + tid * BAND_NODES;
				const uint32_t row = row0 + r;
				// Whole rows first.  Every value the node-level test below looks at -- a node, its x neighbours, the same x in the rows y +- 1
				// and z +- 1 -- lies in five rows; when the minimum over those rows is above the band (or the maximum below it) every node
				// of this row is settled on that side with no neighbour of another class: exactly what the node-level test would find,
				// for a handful of instructions per step instead of hundreds.  Block-uniform (every thread reads the same five pairs), so
				// the barriers inside the skipped part are skipped by all.  Three quarters of the rows of a 1000^3 sphere go this way.
				int rowSide = 0; // +1 / -1: settled row
				if (L.rowMinMax && !L.fill)
				{
					const float2 m0 = L.rowMinMax[row];
					const float2 mYm = y > 0 ? L.rowMinMax[row - 1] : m0, mYp = y + 1 < res ? L.rowMinMax[row + 1] : m0;
					const float2 mZm = zl > 0 ? L.rowMinMax[row - res] : m0, mZp = zl + 1 < L.zcount ? L.rowMinMax[row + res] : m0;
					const float mn = band_min3(band_min3(m0.x, mYm.x, mYp.x), mZm.x, mZp.x);
					const float mx = band_max3(band_max3(m0.y, mYm.y, mYp.y), mZm.y, mZp.y);
					rowSide = mn > above ? 1 : mx < below ? -1 : 0;
				}
				uint32_t total = 0;
				if (rowSide != 0)
				{
					// the sign words of this step's nodes: all ones on the positive side (nodes beyond the row stay 0), the zero fill otherwise
					const uint32_t wx = xs * (BAND_STEP / 32u) + tid;
					if (rowSide > 0 && L.bitPitch && tid < BAND_STEP / 32u && wx * 32u < res)
					{
						const uint32_t bits = res - wx * 32u;
						const uint32_t word = bits >= 32u ? 0xffffffffu : (1u << bits) - 1u;
						const uint64_t at = (uint64_t)row * L.bitPitch + wx;
						if (L.landSigns)
							L.landSigns[at] = word;
						if (L.navSigns)
							L.navSigns[at] = word;
					}
				}
				else
				{
				bool need[BAND_NODES] = { false, false, false, false };
				uint32_t positive = 0; // bit k: node k is settled on the positive side
				bool any = false;
				if (x < res)
				{
					const uint64_t idx = (uint64_t)row * res + x;
					const float *w = L.water + idx;
					float v[BAND_NODES], ym[BAND_NODES], yp[BAND_NODES], zm[BAND_NODES], zp[BAND_NODES];
					band_load4(w, x, res, vector, 0.0f, v);
#pragma unroll
					for (int k = 1; k < BAND_NODES; k++)
						v[k] = x + k < res ? v[k] : v[0]; // nodes beyond the row: copies of a real one, never stored
					// the six neighbours; a neighbour that does not exist is replaced by the node itself
					const float left = x > 0 ? w[-1] : v[0];
					const float right = x + BAND_NODES < res ? w[BAND_NODES] : v[BAND_NODES - 1];
#pragma unroll
					for (int k = 0; k < BAND_NODES; k++)
						ym[k] = yp[k] = zm[k] = zp[k] = v[k];
					if (y > 0)
						band_load4(w - res, x, res, vector, v[0], ym);
					if (y + 1 < res)
						band_load4(w + res, x, res, vector, v[0], yp);
					if (zl > 0)
						band_load4(w - plane, x, res, vector, v[0], zm);
					if (zl + 1 < L.zcount)
						band_load4(w + plane, x, res, vector, v[0], zp);
					// nearly always the 22 values lie on one side of the band: two NaN-propagating reductions settle the four nodes
					float mn = band_min3(band_min3(v[0], v[1], v[2]), band_min3(v[3], left, right), band_min3(ym[0], ym[1], ym[2]));
					mn = band_min3(mn, band_min3(ym[3], yp[0], yp[1]), band_min3(yp[2], yp[3], zm[0]));
					mn = band_min3(mn, band_min3(zm[1], zm[2], zm[3]), band_min3(zp[0], zp[1], zp[2]));
					mn = band_min3(mn, zp[3], zp[3]);
					float mx = band_max3(band_max3(v[0], v[1], v[2]), band_max3(v[3], left, right), band_max3(ym[0], ym[1], ym[2]));
					mx = band_max3(mx, band_max3(ym[3], yp[0], yp[1]), band_max3(yp[2], yp[3], zm[0]));
					mx = band_max3(mx, band_max3(zm[1], zm[2], zm[3]), band_max3(zp[0], zp[1], zp[2]));
					mx = band_max3(mx, zp[3], zp[3]);
					const uint32_t inRow = res - x >= BAND_NODES ? 0xfu : (1u << (res - x)) - 1u;
					if (mn > above)
						positive = inRow;
					else if (!(mx < below))
					{
						// near the band: node by node.  An edge to a node of another class may change sign: both of its end points are evaluated
						int c[BAND_NODES];
#pragma unroll
						for (int k = 0; k < BAND_NODES; k++)
							c[k] = band_class(v[k], above, below);
						const int cl = band_class(left, above, below), cr = band_class(right, above, below);
#pragma unroll
						for (int k = 0; k < BAND_NODES; k++)
						{
							const int l = k == 0 ? cl : c[k - 1];
							const int rr = k == BAND_NODES - 1 ? cr : c[k + 1];
							const bool differs = l != c[k] || rr != c[k] || band_class(ym[k], above, below) != c[k] || band_class(yp[k], above, below) != c[k] || band_class(zm[k], above, below) != c[k] || band_class(zp[k], above, below) != c[k];
							need[k] = ((inRow >> k) & 1u) && (c[k] == 0 || differs);
							any |= need[k];
							if (((inRow >> k) & 1u) && !need[k] && c[k] > 0)
								positive |= 1u << k;
						}
					}
					if (L.fill)
					{
						if (!any && vector && inRow == 0xfu)
						{
							const float off = positive ? hi : lo; // all four on the same side: an edge between two settled nodes of different sides would have made both `need`
							const float4 f4 = make_float4(v[0] - off, v[1] - off, v[2] - off, v[3] - off);
							if (L.land)
								*reinterpret_cast<float4 *>(L.land + idx) = f4;
							if (L.navigation)
								*reinterpret_cast<float4 *>(L.navigation + idx) = f4;
						}
						else
						{
#pragma unroll
							for (int k = 0; k < BAND_NODES; k++)
								if (((inRow >> k) & 1u) && !need[k])
								{
									const float f = v[k] - (((positive >> k) & 1u) ? hi : lo);
									if (L.land)
										L.land[idx + k] = f;
									if (L.navigation)
										L.navigation[idx + k] = f;
								}
						}
					}
				}
				if (L.bitPitch)
				{
					// the sign bits of the settled nodes (the same for land and navigation): lane l holds the nibble of nodes 4 l .. 4 l + 3 of
					// the warp's 128; three shuffles gather eight nibbles into lanes 0, 8, 16, 24.  The nodes to evaluate contribute their
					// bit after the evaluation (drain below, atomicOr: same block, ordered by the barrier).
					uint32_t word = positive;
					word |= __shfl_down_sync(0xffffffffu, word, 1) << 4;
					word |= __shfl_down_sync(0xffffffffu, word, 2) << 8;
					word |= __shfl_down_sync(0xffffffffu, word, 4) << 16;
					const uint32_t wx = (xs * BAND_STEP + warp * 32u * BAND_NODES) / 32u + (lane >> 3);
					if ((lane & 7u) == 0 && wx * 32u < res)
					{
						const uint64_t at = (uint64_t)row * L.bitPitch + wx;
						if (L.landSigns)
							L.landSigns[at] = word;
						if (L.navSigns)
							L.navSigns[at] = word;
					}
				}
				// push the nodes to evaluate into the ring
				uint32_t votes[BAND_NODES] = { 0, 0, 0, 0 }, mine = 0;
				if (__any_sync(0xffffffffu, any))
				{
#pragma unroll
					for (int k = 0; k < BAND_NODES; k++)
					{
						votes[k] = __ballot_sync(0xffffffffu, need[k]);
						mine += __popc(votes[k]);
					}
				}
				if (lane == 0)
					warpCount[warp] = mine;
				__syncthreads();
				uint32_t before = 0;
#pragma unroll
				for (uint32_t k = 0; k < BAND_WARPS; k++)
				{
					const uint32_t n = warpCount[k];
					before += k < warp ? n : 0u;
					total += n;
				}
				if (mine)
				{
					uint32_t slot = head + count + before;
#pragma unroll
					for (int k = 0; k < BAND_NODES; k++)
					{
						if (need[k])
							ring[(slot + __popc(votes[k] & ((1u << lane) - 1u))) & (BAND_RING - 1u)] = r * res + x + k;
						slot += __popc(votes[k]);
					}
				}
				} // node-level scan of this step
				// next scan position
				if (++xs == xSteps)
				{
					xs = 0;
					r++;
					if (++y == res)
					{
						y = 0;
						zl++;
					}
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
						// run-time guard of the premise: every evaluated elevation must lie inside the interval the classification assumed
						// (elevation_interval10: numeric suprema of the generators + 2 %).  One outside it means a settled node somewhere
						// may carry the wrong sign: status bit 1 makes the entry point redo the call with dense sampling.
						if (!(e10 >= lo && e10 <= hi) && !isnan_(e10))
							atomicOr(L.status, 2u);
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
						if (L.bitPitch)
						{
							const uint64_t at = (uint64_t)erow * L.bitPitch + (ex >> 5);
							const uint32_t bit = 1u << (ex & 31u);
							if (L.landSigns)
							{
								if (land >= 0.0f)
									atomicOr(L.landSigns + at, bit);
								if (abs_(land) < L.nearTol)
									atomicOr(L.landNear + at, bit);
							}
							if (L.navSigns)
							{
								if (nav >= 0.0f)
									atomicOr(L.navSigns + at, bit);
								if (abs_(nav) < L.nearTol)
									atomicOr(L.navNear + at, bit);
							}
						}
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
		return planet->sampling == UPG_SAMPLING_BAND && planet->params.elevation != EL_NONE && !force_dense_sampling();
	}

	// land / navigation over z-slices [za, zb) given the water field of the same slices; asynchronous on `stream`
	int band_sample(const Planet *planet, int za, int zb, const float *water, float *land, float *navigation, unsigned long long *evaluated, cudaStream_t stream, const SignBits *landBits, const SignBits *navigationBits, bool fill, const float2 *rowMinMax)
	{
		const PlanetParams &P = planet->params;
		const uint32_t res = (uint32_t)P.grid.res;
		if (zb <= za || (!land && !navigation))
			return UPG_OK;
		double lo = 0, hi = 0;
		elevation_interval10(P.elevation, &lo, &hi);
		BandLaunch L;
		L.water = water;
		L.rowMinMax = rowMinMax;
		L.land = land;
		L.navigation = navigation;
		L.zcount = (uint32_t)(zb - za);
		L.rows = L.zcount * res;
		L.z0 = (uint32_t)za;
		// UPG_BAND_INTERVAL_SCALE (test hook): shrinks the assumed interval so that the run-time guard of the kernel fires
		static const double scale = std::getenv("UPG_BAND_INTERVAL_SCALE") ? std::atof(std::getenv("UPG_BAND_INTERVAL_SCALE")) : 1.0;
		L.lo = (float)(lo * scale);
		L.hi = (float)(hi * scale);
		// the float subtraction of the field and the one of the classification round independently (relative 2^-24 of values up to
		// the box diagonal ~3500 + |hi|): 1/32 is two orders of magnitude above that and keeps every stored value at least 1/64 from zero,
		// fifteen times the near-zero threshold 1e-3
		L.margin = 0.03125f;
		L.vector = res % 4 == 0 && ((uintptr_t)water & 15) == 0 && ((uintptr_t)land & 15) == 0 && ((uintptr_t)navigation & 15) == 0;
		L.status = status_word(planet);
		L.evaluated = evaluated;
		L.fill = fill ? 1u : 0u;
		L.landSigns = land && landBits ? landBits->signs : nullptr;
		L.landNear = land && landBits ? landBits->near : nullptr;
		L.navSigns = navigation && navigationBits ? navigationBits->signs : nullptr;
		L.navNear = navigation && navigationBits ? navigationBits->near : nullptr;
		L.bitPitch = L.landSigns ? landBits->pitch : L.navSigns ? navigationBits->pitch : 0u;
		L.nearTol = NEAR_ZERO_TOLERANCE;
		if (L.landSigns && L.navSigns && landBits->pitch != navigationBits->pitch)
			return fail(UPG_ERR_ARGUMENT, "bit fields of different pitch");
		if (!fill && ((land && !L.landSigns) || (navigation && !L.navSigns)))
			return fail(UPG_ERR_ARGUMENT, "band sampling without fill values needs the bit fields");
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
