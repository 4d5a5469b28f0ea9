// upg_field.cuh -- K1 `sdf_sample`: dense evaluation of one terrain field over the meshing grid.
//
// Replaces MarchingCubes::updateByPosition (externals/cage/sources/libcore/mesh/marchingCubes.cpp:
// 149-165) driven by terrainSdfLand / Water / Navigation (sources/modesConfiguration.cpp:248-279):
// a serial z/y/x triple loop with one indirect call per sample becomes one thread per sample.
// Layout of the output is the reference's: dens[(z * res + y) * res + x], float32.
// The kernel is compute bound (about 10^3 FP32/INT32 instructions per 4-byte store), so the only
// memory rule that matters is that a warp's 32 stores are consecutive in x (coalesced).
#pragma once

#include "upg_terrain.cuh"

namespace upg
{
	constexpr int FIELD_BLOCK = 256;

	// everything one launch needs besides the planet
	struct FieldLaunch
	{
		float *out; // device, `count` floats (FD_ALL: the land field, may be null)
		float *outWater = nullptr; // FD_ALL only: the water field, may be null
		float *outNavigation = nullptr; // FD_ALL only: the navigation field, may be null
		const float *positions; // device, count x 3 floats, or nullptr = grid mode
		uint64_t count; // samples in this launch
		uint32_t z0; // first z-slice (grid mode)
		int32_t field; // Field
		uint32_t *status; // device status word
		bool contract = false; // dense sampling entry points only: the planet's UPG_ARITH_CONTRACT build may serve this launch
		// row-wise water launches only (shape_rows_kernel): optional sign / near-zero bit fields of the samples, one bit per node,
		// bit (x & 31) of word (row * bitPitch + (x >> 5)); see SignBits in upg_internal.h
		uint32_t *signBits = nullptr;
		uint32_t *nearBits = nullptr;
		uint32_t bitPitch = 0;
		float nearTol = 0;
		// row-wise water launches only: optional minimum / maximum of every grid row (two floats per row: min, max; NaN when the row
		// holds one): the band sampler skips the rows whose whole neighbourhood lies on one side of the elevation band (upg_band.cu)
		float *rowMinMax = nullptr;
	};

	// runtime field selection (uniform across the grid) over the compile-time shape / elevation pair
	template<int SHAPE, int ELEV>
	UPG_HD float terrain_field_rt(const PlanetParams &P, int field, V3 pos)
	{
		switch (field)
		{
			case FD_WATER:
				return terrain_field<SHAPE, ELEV, FD_WATER>(P, pos);
			case FD_ELEVATION:
				return terrain_field<SHAPE, ELEV, FD_ELEVATION>(P, pos);
			case FD_ELEVATION_RAW:
				return terrain_field<SHAPE, ELEV, FD_ELEVATION_RAW>(P, pos);
			default:
			{
				// land / navigation share everything but the final clamp (sources/modesConfiguration.cpp:248-279)
				const float base = shape_sdf<SHAPE>(P, pos);
				const float elev = elevation_raw<ELEV>(P, pos) / 10.0f;
				return field == FD_NAVIGATION ? base - fmax_(elev, 0.0f) : base - elev;
			}
		}
	}

#if defined(__CUDACC__)
	// ---- craters: the Voronoi points of a block's neighbourhood, computed once per block -------------------------------------------------
	// Voronoi::evaluate (sources/voronoi.cpp:45-91) generates the feature points of the 27 cells around the sample from an integer hash
	// chain -- 81 murmur mixes and 54 points per sample, ~7 000 instructions, nearly all of the `craters` elevation.  The 256 samples of
	// a block are consecutive nodes of one grid row (8 units apart on the 500^3 grid against cells of 300 units): their neighbourhoods
	// overlap almost entirely.  So the block computes the points of the UNION of its samples' neighbourhoods once (each thread hashes
	// a cell or two) into shared memory, and every sample walks its own 27 cells in the reference's z / y / x / point order over that
	// table: the same points, the same distances, the same strict-less comparisons -- bit-identical -- for ~1/30 of the hashing.
	// Blocks whose samples span too many cells (the block that wraps from the end of one slice to the start of the next) and Voronoi
	// instances with more than two points per cell take the per-sample search.
	// NOT part of the default build: written after the round's GPU budget was spent, never run on a B200.  -DUPG_CRATERS_TABLE=1 compiles it
	// in (make -C csrc EXTRA_NVFLAGS=-DUPG_CRATERS_TABLE=1); without it the craters kernels are the per-sample search the parity tests ran against.
#ifndef UPG_CRATERS_TABLE
#define UPG_CRATERS_TABLE 0
#endif
	constexpr int CRATER_MAX_CELLS = 160;
	constexpr int CRATER_MAX_POINTS = 2;
	struct CraterTable
	{
		float x[CRATER_MAX_CELLS * CRATER_MAX_POINTS], y[CRATER_MAX_CELLS * CRATER_MAX_POINTS], z[CRATER_MAX_CELLS * CRATER_MAX_POINTS];
		int32_t lo[3], hi[3]; // cell range of the block's samples (base cells), inclusive
	};

	// all threads of the block call this (barriers inside); `live` = the thread has a sample.  Returns false (for the whole block) when the
	// table does not apply; otherwise `nearest` = what voronoi_nearest<1, false> returns for `pos`.
	__device__ __forceinline__ bool craters_block_nearest(const VoronoiInst &v, V3 pos, bool live, CraterTable &T, V3 &nearest)
	{
		const V3 pf = pos * v.frequency;
		const I3 base = I3{ (int32_t)pf.x, (int32_t)pf.y, (int32_t)pf.z };
		if (threadIdx.x == 0)
		{
			T.lo[0] = T.lo[1] = T.lo[2] = 0x7fffffff;
			T.hi[0] = T.hi[1] = T.hi[2] = (int32_t)0x80000000;
		}
		__syncthreads();
		// warp-level extremes first, one shared atomic per warp and bound
		const int32_t big = 0x7fffffff, small = (int32_t)0x80000000;
		const int32_t c[3] = { base.x, base.y, base.z };
#pragma unroll
		for (int a = 0; a < 3; a++)
		{
			const int32_t mn = __reduce_min_sync(0xffffffffu, live ? c[a] : big), mx = __reduce_max_sync(0xffffffffu, live ? c[a] : small);
			if ((threadIdx.x & 31u) == 0)
			{
				atomicMin(&T.lo[a], mn);
				atomicMax(&T.hi[a], mx);
			}
		}
		__syncthreads();
		// the union of the 27-cell neighbourhoods: [lo - 1, hi + 1] per axis
		const int32_t ox = T.lo[0] - 1, oy = T.lo[1] - 1, oz = T.lo[2] - 1;
		const int64_t nx = (int64_t)T.hi[0] - T.lo[0] + 3, ny = (int64_t)T.hi[1] - T.lo[1] + 3, nz = (int64_t)T.hi[2] - T.lo[2] + 3;
		if (v.pointsPerCell > (uint32_t)CRATER_MAX_POINTS || nx * ny * nz > CRATER_MAX_CELLS || nx <= 0 || ny <= 0 || nz <= 0)
			return false; // uniform: every thread read the same bounds
		const int32_t cells = (int32_t)(nx * ny * nz);
		for (int32_t k = (int32_t)threadIdx.x; k < cells; k += (int32_t)blockDim.x)
		{
			const int32_t cx = k % (int32_t)nx, cy = (k / (int32_t)nx) % (int32_t)ny, cz = k / (int32_t)(nx * ny);
			const I3 cell = I3{ ox + cx, oy + cy, oz + cz };
			I3 s = voronoi_mix(v, cell); // genCell: s = mix(cell); per point s = mix(s) (sources/voronoi.cpp:28-41)
			for (uint32_t i = 0; i < v.pointsPerCell; i++)
			{
				s = voronoi_mix(v, s);
				const V3 p = voronoi_point(v, s, cell);
				T.x[k * CRATER_MAX_POINTS + i] = p.x;
				T.y[k * CRATER_MAX_POINTS + i] = p.y;
				T.z[k * CRATER_MAX_POINTS + i] = p.z;
			}
		}
		__syncthreads();
		// voronoi_nearest<1, false>: z, y, x, point order; strict less, so the earlier point keeps a tie
		float best = INFINITY;
		nearest = v3(0, 0, 0);
		if (live)
		{
			for (int32_t dz = -1; dz < 2; dz++)
				for (int32_t dy = -1; dy < 2; dy++)
				{
					const int32_t row = ((base.z + dz - oz) * (int32_t)ny + (base.y + dy - oy)) * (int32_t)nx + (base.x - 1 - ox);
#pragma unroll
					for (int32_t dx = 0; dx < 3; dx++)
						for (uint32_t i = 0; i < v.pointsPerCell; i++)
						{
							const int32_t at = (row + dx) * CRATER_MAX_POINTS + (int32_t)i;
							const V3 p = v3(T.x[at], T.y[at], T.z[at]);
							const float d = distance_squared(p, pos);
							if (d < best)
							{
								best = d;
								nearest = p;
							}
						}
				}
		}
		return true;
	}

	// status word: bit 0 = a NaN sample was produced (the reference throws "invalid ... sdf value",
	// sources/modesConfiguration.cpp:255-256; the host layer turns the flag into UPG_ERR_INVALID_VALUE)
	// ARITH = the arithmetic build of the translation unit (UPG_ARITH_BUILD): part of the kernel's TYPE, so that the exact and the contracted
	// build of one (shape, elevation) pair are two symbols (two weak template instantiations with one name would be merged by the linker)
	template<int SHAPE, int ELEV, int ARITH = UPG_ARITH_BUILD>
	__global__ void __launch_bounds__(FIELD_BLOCK) sdf_sample_kernel(const __grid_constant__ PlanetParams P, const __grid_constant__ FieldLaunch L)
	{
		grad_table_init<FIELD_BLOCK>(); // shared-memory gradient table of the Perlin / OpenSimplex generators (upg_noise.cuh)
		__syncthreads();
		const uint64_t idx = (uint64_t)blockIdx.x * FIELD_BLOCK + threadIdx.x;
		// grid launches of the `craters` mode keep every thread of the block until the shared Voronoi table is built
		constexpr bool CRATERS = (UPG_CRATERS_TABLE != 0) && ELEV == EL_CRATERS;
		const bool cratersGrid = CRATERS && !L.positions;
		const bool live = idx < L.count;
		if (!live && !cratersGrid)
			return;
		V3 pos;
		if (!live)
			pos = v3(0, 0, 0);
		else if (L.positions)
		{
			// explicit positions (parity probes, water-vertex culling of sources/meshGeneration.cpp:62-68)
			pos = v3(L.positions[idx * 3 + 0], L.positions[idx * 3 + 1], L.positions[idx * 3 + 2]);
		}
		else
		{
			const uint32_t res = (uint32_t)P.grid.res;
			const uint32_t row = (uint32_t)(idx / res);
			const uint32_t x = (uint32_t)(idx - (uint64_t)row * res);
			const uint32_t zl = row / res;
			const uint32_t y = row - zl * res;
			pos = v3(grid_coord(P.grid, x), grid_coord(P.grid, y), grid_coord(P.grid, L.z0 + zl));
		}
		// One inlined copy of the shape and of the elevation function serves every field (the instruction cache is the scarce
		// resource of this kernel); the field only selects which of the two run and how they combine
		// (sources/modesConfiguration.cpp:230-279).
		const int field = L.field;
		float base = 0, elev = 0;
		const bool needElevation = field != FD_WATER && field != FD_ELEVATION;
		V3 nearest = v3(0, 0, 0);
		bool haveNearest = false;
		if constexpr (CRATERS)
		{
			__shared__ CraterTable craterTable;
			if (cratersGrid && needElevation) // block-uniform
				haveNearest = craters_block_nearest(P.el.voronoi, pos, live, craterTable, nearest);
			if (!live)
				return;
		}
		if (field != FD_ELEVATION_RAW)
			base = shape_sdf<SHAPE>(P, pos);
		if (needElevation)
			elev = haveNearest ? elevation_raw_given<ELEV>(P, pos, 0.0f, nullptr, nullptr, &nearest) : elevation_raw<ELEV>(P, pos);
		if (field == FD_ALL)
		{
			// land, water and navigation of the same sample share the shape and the elevation: the generator asks for all three
			// (sources/generator.cpp:432-439), so they are produced by one evaluation and up to three coalesced stores
			const float e10 = elev / 10.0f;
			const float land = base - e10;
			const float nav = base - fmax_(e10, 0.0f);
			bool bad = false;
			if (L.out)
			{
				L.out[idx] = land;
				bad |= isnan_(land);
			}
			if (L.outWater)
			{
				L.outWater[idx] = base;
				bad |= isnan_(base);
			}
			if (L.outNavigation)
			{
				L.outNavigation[idx] = nav;
				bad |= isnan_(nav);
			}
			if (bad)
				atomicOr(L.status, 1u);
			return;
		}
		float v;
		switch (field)
		{
			case FD_WATER:
				v = base;
				break;
			case FD_ELEVATION:
				v = base * 10.0f;
				break;
			case FD_ELEVATION_RAW:
				v = elev;
				break;
			case FD_NAVIGATION:
				v = base - fmax_(elev / 10.0f, 0.0f);
				break;
			default:
				v = base - elev / 10.0f;
				break;
		}
		if (isnan_(v))
			atomicOr(L.status, 1u);
		L.out[idx] = v;
	}

	// K1 over whole z-slices for the elevation modes that start with a Value fBm of the position (lakes, islands, legacy): a thread owns
	// one (x, y) COLUMN of the grid and walks COLUMN_Z slices of it, so that the fBm's per-octave interpolation planes live in
	// registers between samples (ValueFbmColumn, upg_noise.cuh: bit-identical, 8 hashes + 7 lerps per octave become 1 lerp) and the
	// x / y coordinates, their IEEE divisions and the index arithmetic are paid once per column.  z is uniform over the block, so
	// "the column left its lattice cell" is a uniform branch.
	constexpr int COLUMN_TILE_X = 32, COLUMN_TILE_Y = FIELD_BLOCK / COLUMN_TILE_X; // columns per block
#ifndef UPG_COLUMN_Z
#define UPG_COLUMN_Z 64
#endif
	constexpr int COLUMN_Z = UPG_COLUMN_Z;
#ifndef UPG_COLUMN_PERLIN
#define UPG_COLUMN_PERLIN 1 // also cache the Perlin mountain mask along the column
#endif
#ifndef UPG_COLUMN_TERRACES
#define UPG_COLUMN_TERRACES 0 // A/B: also cache the terrace octaves (57 more floats per thread)
#endif
#ifndef UPG_COLUMN_MIN_BLOCKS
#define UPG_COLUMN_MIN_BLOCKS 5 // resident blocks per SM the register allocation aims at (A/B in profiles/)
#endif
	// the samples are written once and never read by this kernel: UPG_K1_STORE = 1 marks the stores streaming (st.global.cs, evict
	// first), 0 = plain stores.  The variant was built (VERDICT r1 item 3 asked where the 1.4 GB of DRAM reads of a store-only kernel come
	// from) but NOT measured: the round's GPU budget ended first.  The kernel is issue bound at 9 % of the DRAM throughput
	// (profiles/r02_l_ncu_summary.md), so nothing is expected from it either way.
#ifndef UPG_K1_STORE
#define UPG_K1_STORE 0
#endif
	__device__ __forceinline__ void store_sample(float *p, float v)
	{
#if UPG_K1_STORE == 1
		__stcs(p, v);
#else
		*p = v;
#endif
	}

	template<int SHAPE, int ELEV, int ARITH = UPG_ARITH_BUILD>
	__global__ void __launch_bounds__(FIELD_BLOCK, UPG_COLUMN_MIN_BLOCKS) sdf_column_kernel(const __grid_constant__ PlanetParams P, const __grid_constant__ FieldLaunch L)
	{
		__shared__ float coordZ[COLUMN_Z];
		const uint32_t res = (uint32_t)P.grid.res;
		const uint32_t zcount = (uint32_t)(L.count / ((uint64_t)res * res));
		const uint32_t zBegin = blockIdx.z * COLUMN_Z;
		const uint32_t zEnd = min(zBegin + (uint32_t)COLUMN_Z, zcount);
		grad_table_init<FIELD_BLOCK>();
		if (threadIdx.x < zEnd - zBegin)
			coordZ[threadIdx.x] = grid_coord(P.grid, L.z0 + zBegin + threadIdx.x);
		__syncthreads();
		// a block owns COLUMN_TILE_X x COLUMN_TILE_Y columns and a warp an 8 x 4 tile of them: the mountain masks are coherent in SPACE, so a
		// compact warp footprint (32 x 16 world units instead of 128 x 4 on the 1000^3 grid) crosses their borders less often, and a
		// warp's 32 stores per slice are still four full 32-byte sectors
		const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#if UPG_COLUMN_LINE // A/B: a warp = 32 consecutive x of one row
		const uint32_t x = blockIdx.x * COLUMN_TILE_X + lane;
		const uint32_t y = blockIdx.y * COLUMN_TILE_Y + warp;
#else
#ifndef UPG_COLUMN_WARP_W
#define UPG_COLUMN_WARP_W 8 // A/B: 16 = 16 x 2 tiles (64-byte store segments, more mask borders crossed)
#endif
		constexpr uint32_t WW = UPG_COLUMN_WARP_W, WH = 32u / WW, WPR = COLUMN_TILE_X / WW; // warp tile, warps per block row
		const uint32_t x = blockIdx.x * COLUMN_TILE_X + (warp % WPR) * WW + (lane % WW);
		const uint32_t y = blockIdx.y * COLUMN_TILE_Y + (warp / WPR) * WH + (lane / WW);
#endif
		if (x >= res || y >= res)
			return;
		const float px = grid_coord(P.grid, x), py = grid_coord(P.grid, y);
		const int field = L.field;
		ValueFbmColumn<4> column;
#if UPG_COLUMN_PERLIN
		PerlinColumn maskColumn; // the mountain mask of lakes / islands (elevation_mountains)
#endif
#if UPG_COLUMN_TERRACES
		PerlinFbmColumn<3> terraceColumn; // the three Perlin octaves behind the terrace mask
#endif
		// NaN bookkeeping of the stored samples: one NaN-propagating maximum per store (FMNMX.NAN) instead of compare + select + or
		float nanAcc = 0.0f;
		uint64_t idx = ((uint64_t)zBegin * res + y) * res + x;
		const uint64_t slice = (uint64_t)res * res;
		for (uint32_t zl = zBegin; zl < zEnd; zl++, idx += slice)
		{
			const V3 pos = v3(px, py, coordZ[zl - zBegin]);
			float base = 0;
			if (field != FD_ELEVATION_RAW)
				base = shape_sdf<SHAPE>(P, pos);
			const float first = value_fbm_column<4>(P.el.n[0], pos.x, pos.y, pos.z, column);
#if UPG_COLUMN_PERLIN && UPG_COLUMN_TERRACES
			const float elev = elevation_raw_given<ELEV>(P, pos, first, &maskColumn, &terraceColumn);
#elif UPG_COLUMN_PERLIN
			const float elev = elevation_raw_given<ELEV>(P, pos, first, &maskColumn);
#else
			const float elev = elevation_raw_given<ELEV>(P, pos, first);
#endif
			// the expressions of sdf_sample_kernel
			if (field == FD_ALL)
			{
				const float e10 = div_tol(elev, 10.0f);
				const float land = base - e10;
				const float nav = base - fmax_(e10, 0.0f);
				if (L.out)
				{
					store_sample(L.out + idx, land);
					asm("max.NaN.f32 %0, %0, %1;" : "+f"(nanAcc) : "f"(land));
				}
				if (L.outWater)
				{
					store_sample(L.outWater + idx, base);
					asm("max.NaN.f32 %0, %0, %1;" : "+f"(nanAcc) : "f"(base));
				}
				if (L.outNavigation)
				{
					store_sample(L.outNavigation + idx, nav);
					asm("max.NaN.f32 %0, %0, %1;" : "+f"(nanAcc) : "f"(nav));
				}
			}
			else
			{
				const float v = field == FD_ELEVATION_RAW ? elev : field == FD_NAVIGATION ? base - fmax_(div_tol(elev, 10.0f), 0.0f) : base - div_tol(elev, 10.0f);
				asm("max.NaN.f32 %0, %0, %1;" : "+f"(nanAcc) : "f"(v));
				L.out[idx] = v;
			}
		}
		if (isnan_(nanAcc))
			atomicOr(L.status, 1u);
	}

	// The water field (= the bare shape, terrainSdfWater) of whole grid rows: the shape functions cost a few dozen instructions, so
	// the index arithmetic of the generic kernel above (a 64-bit division and three IEEE divisions per sample) would dominate.  Here
	// the grid -> world coordinates come from a per-block table (same grid_coord(), same bits) and a block walks SHAPE_ROWS rows.
	// The band sampler (upg_band.cu) reads the shape of every node from this field.
	constexpr int SHAPE_ROWS = 16;
	constexpr int SHAPE_MAX_RES = 2048; // upg_planet_create's limit
	template<int SHAPE>
	__global__ void __launch_bounds__(FIELD_BLOCK) shape_rows_kernel(const __grid_constant__ PlanetParams P, const __grid_constant__ FieldLaunch L)
	{
		__shared__ float coord[SHAPE_MAX_RES];
		const uint32_t res = (uint32_t)P.grid.res;
		grad_table_init<FIELD_BLOCK>(); // the asteroid shape samples a noise
		// per-row minimum / maximum (L.rowMinMax): every warp leaves its part of a row as order-preserving integer keys (one redux.sync
		// each), the parts are combined once all rows of the block are done
		__shared__ uint32_t partMin[SHAPE_ROWS][FIELD_BLOCK / 32], partMax[SHAPE_ROWS][FIELD_BLOCK / 32];
		__shared__ uint32_t rowHasNan; // bit r: row r of the block holds a NaN
		for (uint32_t i = threadIdx.x; i < res; i += FIELD_BLOCK)
			coord[i] = grid_coord(P.grid, i);
		if (threadIdx.x == 0)
			rowHasNan = 0;
		__syncthreads();
		const uint32_t rows = (uint32_t)(L.count / res);
		const uint32_t row0 = blockIdx.x * SHAPE_ROWS;
		const uint32_t rowEnd = min(row0 + (uint32_t)SHAPE_ROWS, rows);
		bool bad = false;
		for (uint32_t row = row0; row < rowEnd; row++)
		{
			const uint32_t zl = row / res;
			const float cy = coord[row - zl * res], cz = coord[L.z0 + zl];
			float *out = L.out + (uint64_t)row * res;
			float rowMin = __int_as_float(0x7f800000), rowMax = __int_as_float(0xff800000); // +inf / -inf
			bool rowNan = false;
			// every warp stays whole through the loop (x0 is warp uniform): its 32 samples are one word of the bit fields
			for (uint32_t x0 = threadIdx.x & ~31u; x0 < res; x0 += FIELD_BLOCK)
			{
				const uint32_t x = x0 + (threadIdx.x & 31u);
				float v = -1.0f;
				if (x < res)
				{
					v = shape_sdf<SHAPE>(P, v3(coord[x], cy, cz));
					bad |= isnan_(v);
					out[x] = v;
					rowMin = fminf(rowMin, v);
					rowMax = fmaxf(rowMax, v);
					rowNan |= isnan_(v); // fminf / fmaxf drop NaNs: a row with a NaN sample must never be skipped
				}
				if (L.signBits)
				{
					// the extraction's inside test (`>= iso`, dualmc.tpp:12-32) and its near-zero statistic, taken here once
					const uint32_t sign = __ballot_sync(0xffffffffu, x < res && v >= 0.0f);
					const uint32_t near = __ballot_sync(0xffffffffu, x < res && abs_(v) < L.nearTol);
					if ((threadIdx.x & 31u) == 0)
					{
						const uint64_t word = (uint64_t)row * L.bitPitch + (x0 >> 5);
						L.signBits[word] = sign;
						if (near)
							L.nearBits[word] = near;
					}
				}
			}
			if (L.rowMinMax) // uniform
			{
				// float -> unsigned key of the same order, one redux.sync per bound
				const uint32_t ua = __float_as_uint(rowMin), ub = __float_as_uint(rowMax);
				const uint32_t ka = __reduce_min_sync(0xffffffffu, (ua & 0x80000000u) ? ~ua : (ua | 0x80000000u));
				const uint32_t kb = __reduce_max_sync(0xffffffffu, (ub & 0x80000000u) ? ~ub : (ub | 0x80000000u));
				const bool anyNan = __any_sync(0xffffffffu, rowNan);
				if ((threadIdx.x & 31u) == 0)
				{
					partMin[row - row0][threadIdx.x >> 5] = ka;
					partMax[row - row0][threadIdx.x >> 5] = kb;
					if (anyNan)
						atomicOr(&rowHasNan, 1u << (row - row0));
				}
			}
		}
		if (L.rowMinMax)
		{
			__syncthreads();
			if (threadIdx.x < rowEnd - row0)
			{
				uint32_t ka = partMin[threadIdx.x][0], kb = partMax[threadIdx.x][0];
				for (int w = 1; w < FIELD_BLOCK / 32; w++)
				{
					ka = min(ka, partMin[threadIdx.x][w]);
					kb = max(kb, partMax[threadIdx.x][w]);
				}
				const float lo = __uint_as_float((ka & 0x80000000u) ? (ka & 0x7fffffffu) : ~ka), hi = __uint_as_float((kb & 0x80000000u) ? (kb & 0x7fffffffu) : ~kb);
				const float nan = __int_as_float(0x7fc00000);
				const bool hasNan = (rowHasNan >> threadIdx.x) & 1u;
				reinterpret_cast<float2 *>(L.rowMinMax)[row0 + threadIdx.x] = hasNan ? make_float2(nan, nan) : make_float2(lo, hi);
			}
		}
		if (bad)
			atomicOr(L.status, 1u);
	}

	// host-callable launcher type; one per (elevation, shape range) translation unit, see upg_field_inst.cu
	typedef bool (*FieldLauncher)(int shape, const PlanetParams &P, const FieldLaunch &L, void *stream);
#endif
}
