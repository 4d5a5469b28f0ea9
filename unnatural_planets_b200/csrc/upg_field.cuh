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
		float *out; // device, `count` floats
		const float *positions; // device, count x 3 floats, or nullptr = grid mode
		uint64_t count; // samples in this launch
		uint32_t z0; // first z-slice (grid mode)
		int32_t field; // Field
		uint32_t *status; // device status word
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
	// status word: bit 0 = a NaN sample was produced (the reference throws "invalid ... sdf value",
	// sources/modesConfiguration.cpp:255-256; the host layer turns the flag into UPG_ERR_INVALID_VALUE)
	template<int SHAPE, int ELEV>
	__global__ void __launch_bounds__(FIELD_BLOCK) sdf_sample_kernel(const __grid_constant__ PlanetParams P, const __grid_constant__ FieldLaunch L)
	{
		grad_table_init<FIELD_BLOCK>(); // shared-memory gradient table of the Perlin / OpenSimplex generators (upg_noise.cuh)
		__syncthreads();
		const uint64_t idx = (uint64_t)blockIdx.x * FIELD_BLOCK + threadIdx.x;
		if (idx >= L.count)
			return;
		V3 pos;
		if (L.positions)
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
		const float v = terrain_field_rt<SHAPE, ELEV>(P, L.field, pos);
		if (isnan_(v))
			atomicOr(L.status, 1u);
		L.out[idx] = v;
	}

	// host-callable launcher type; one per (elevation, shape range) translation unit, see upg_field_inst.cu
	typedef bool (*FieldLauncher)(int shape, const PlanetParams &P, const FieldLaunch &L, void *stream);
#endif
}
