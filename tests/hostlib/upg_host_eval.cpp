// upg_host_eval.cpp -- TEST-ONLY host compilation of the product's device headers
// (unnatural_planets_b200/csrc/upg_{common,noise,terrain,planet_host}), so that the CPU test-suite can
// exercise the very same source the CUDA kernels are built from, against oracle/_ref and the golden
// fixtures, without a GPU.  This library is never loaded by the product; libupg.so has no CPU path.
#include <cstdint>
#include <cstring>

#include "upg_field.cuh"
#include "upg_planet_host.h"

using namespace upg;

namespace
{
	template<int S>
	struct ShapeDispatch
	{
		static float run(int shape, const PlanetParams &P, V3 p)
		{
			if (shape == S)
				return shape_sdf<S>(P, p);
			return ShapeDispatch<S + 1>::run(shape, P, p);
		}
	};
	template<>
	struct ShapeDispatch<SH_COUNT>
	{
		static float run(int, const PlanetParams &, V3) { return NAN; }
	};
	template<int E>
	struct ElevDispatch
	{
		static float run(int elev, const PlanetParams &P, V3 p)
		{
			if (elev == E)
				return elevation_raw<E>(P, p);
			return ElevDispatch<E + 1>::run(elev, P, p);
		}
	};
	template<>
	struct ElevDispatch<EL_COUNT>
	{
		static float run(int, const PlanetParams &, V3) { return NAN; }
	};

	float eval_field(const PlanetParams &P, int field, V3 pos)
	{
		switch (field)
		{
			case FD_WATER:
				return ShapeDispatch<0>::run(P.shape, P, pos);
			case FD_ELEVATION:
				return ShapeDispatch<0>::run(P.shape, P, pos) * 10.0f;
			case FD_ELEVATION_RAW:
				return ElevDispatch<0>::run(P.elevation, P, pos);
			default:
			{
				const float base = ShapeDispatch<0>::run(P.shape, P, pos);
				const float elev = ElevDispatch<0>::run(P.elevation, P, pos) / 10.0f;
				return field == FD_NAVIGATION ? base - fmax_(elev, 0.0f) : base - elev;
			}
		}
	}
}

extern "C"
{
	struct upgh_planet
	{
		PlanetParams params;
		SeedLog log;
	};

	upgh_planet *upgh_create(int shape, int elevation, int coloring, int poles, int flowers, uint64_t s0, uint64_t s1, float boxSize, int res)
	{
		upgh_planet *p = new upgh_planet();
		std::memset(&p->params, 0, sizeof(p->params));
		p->params.shape = shape;
		p->params.elevation = elevation;
		p->params.coloring = coloring;
		p->params.poles = poles;
		p->params.flowers = flowers;
		set_grid(p->params, boxSize, res);
		PlanetBuilder b;
		b.thread.s[0] = s0;
		b.thread.s[1] = s1;
		b.shapeStatics(p->params);
		b.elevationStatics(p->params);
		p->log = b.log;
		return p;
	}
	void upgh_destroy(upgh_planet *p)
	{
		delete p;
	}
	int upgh_seed_log(const upgh_planet *p, uint32_t *out, int cap)
	{
		for (int i = 0; i < p->log.count && i < cap; i++)
			out[i] = p->log.seeds[i];
		return p->log.count;
	}
	void upgh_points(const upgh_planet *p, int field, const float *pos, uint64_t n, float *out)
	{
		for (uint64_t i = 0; i < n; i++)
			out[i] = eval_field(p->params, field, v3(pos[i * 3 + 0], pos[i * 3 + 1], pos[i * 3 + 2]));
	}
	void upgh_grid(const upgh_planet *p, int field, int z0, int z1, float *out)
	{
		const GridConsts &g = p->params.grid;
		const uint32_t res = (uint32_t)g.res;
		for (uint32_t z = (uint32_t)z0; z < (uint32_t)z1; z++)
			for (uint32_t y = 0; y < res; y++)
				for (uint32_t x = 0; x < res; x++)
					*out++ = eval_field(p->params, field, v3(grid_coord(g, x), grid_coord(g, y), grid_coord(g, z)));
	}
	void upgh_grid_coords(const upgh_planet *p, float *out)
	{
		for (int i = 0; i < p->params.grid.res; i++)
			out[i] = grid_coord(p->params.grid, (uint32_t)i);
	}
	// one noise call with explicit configuration: type / fractal / octaves are runtime here, so go through a small table
	float upgh_noise3(int type, int fractal, int octaves, int32_t seed, float frequency, float lacunarity, float gain, float x, float y, float z)
	{
		NoiseInst n = make_noise((uint32_t)seed, frequency, octaves, gain, lacunarity);
#define UPGH_CASE(T, F, O) \
	if (type == T && fractal == F && octaves == O) \
		return noise3<T, F, O>(n, x, y, z);
#define UPGH_TYPE(T) \
	UPGH_CASE(T, FT_NONE, 3) \
	UPGH_CASE(T, FT_FBM, 2) UPGH_CASE(T, FT_FBM, 3) UPGH_CASE(T, FT_FBM, 4) UPGH_CASE(T, FT_RIDGED, 2) UPGH_CASE(T, FT_RIDGED, 3) UPGH_CASE(T, FT_RIDGED, 4) UPGH_CASE(T, FT_RIDGED, 6) UPGH_CASE(T, FT_PINGPONG, 3)
		UPGH_TYPE(NT_SIMPLEX)
		UPGH_TYPE(NT_SIMPLEX_REDUCED)
		UPGH_TYPE(NT_PERLIN)
		UPGH_TYPE(NT_CUBIC)
		UPGH_TYPE(NT_VALUE)
		return NAN;
	}
}
