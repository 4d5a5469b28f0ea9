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

	// elevation through the column path of the dense kernel (csrc/upg_field.cuh: sdf_column_kernel): the first Value fBm from a
	// ValueFbmColumn that lives as long as the (x, y) column
	template<int E>
	struct ElevColumnDispatch
	{
		static float run(int elev, const PlanetParams &P, V3 p, ValueFbmColumn<4> &c, PerlinColumn &m, PerlinFbmColumn<3> &tc)
		{
			if (elev == E)
			{
				float first = 0;
				if (ElevationStartsWithValueFbm<E>::value)
					first = value_fbm_column<4>(P.el.n[0], p.x, p.y, p.z, c);
				return elevation_raw_given<E>(P, p, first, &m, &tc);
			}
			return ElevColumnDispatch<E + 1>::run(elev, P, p, c, m, tc);
		}
	};
	template<>
	struct ElevColumnDispatch<EL_COUNT>
	{
		static float run(int, const PlanetParams &, V3, ValueFbmColumn<4> &, PerlinColumn &, PerlinFbmColumn<3> &) { return NAN; }
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
		ColoringConsts coloring;
		float cellular[1024];
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
		coloring_statics(b, p->params, p->coloring);
		default_cellular_table(p->cellular);
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
	// the bounds of elevation / 10 the band sampler (csrc/upg_band.cu) classifies with
	void upgh_elevation_interval10(int elevation, double *lo10, double *hi10)
	{
		elevation_interval10(elevation, lo10, hi10);
	}
	// terrainSdfElevationRaw over the grid, evaluated column by column like sdf_column_kernel does; zOrder: 0 ascending, 1 descending,
	// 2 = a zig-zag with repeats (the cache must not care)
	void upgh_grid_elevation_columns(const upgh_planet *p, int z0, int z1, int zOrder, float *out)
	{
		const GridConsts &g = p->params.grid;
		const uint32_t res = (uint32_t)g.res;
		const int n = z1 - z0;
		int stride = 7; // of the zig-zag: coprime with n, so that every slice is visited once
		for (bool ok = false; !ok; stride += ok ? 0 : 2)
		{
			int a = stride, b = n;
			while (b)
			{
				const int t = a % b;
				a = b;
				b = t;
			}
			ok = a == 1;
		}
		for (uint32_t y = 0; y < res; y++)
			for (uint32_t x = 0; x < res; x++)
			{
				ValueFbmColumn<4> c;
				PerlinColumn m;
				PerlinFbmColumn<3> tc;
				for (int k = 0; k < n; k++)
				{
					int zi = zOrder == 0 ? k : zOrder == 1 ? n - 1 - k : ((k * stride) % n);
					const uint32_t z = (uint32_t)(z0 + zi);
					out[((uint64_t)zi * res + y) * res + x] = ElevColumnDispatch<0>::run(p->params.elevation, p->params, v3(grid_coord(g, x), grid_coord(g, y), grid_coord(g, z)), c, m, tc);
				}
			}
	}
	// the same for the first `rows` grid rows of every slice only, ascending z (big grids on the host)
	void upgh_grid_elevation_columns_rows(const upgh_planet *p, int z0, int z1, int rows, float *out)
	{
		const GridConsts &g = p->params.grid;
		const uint32_t res = (uint32_t)g.res;
		for (uint32_t y = 0; y < (uint32_t)rows; y++)
			for (uint32_t x = 0; x < res; x++)
			{
				ValueFbmColumn<4> c;
				PerlinColumn m;
				PerlinFbmColumn<3> tc;
				for (int z = z0; z < z1; z++)
					out[((uint64_t)(z - z0) * res + y) * res + x] = ElevColumnDispatch<0>::run(p->params.elevation, p->params, v3(grid_coord(g, x), grid_coord(g, y), grid_coord(g, (uint32_t)z)), c, m, tc);
			}
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

	// terrainTile on explicit position / normal pairs, same 13 output columns as oracle/ref_harness.cpp upref_tiles
	void upgh_tiles(const upgh_planet *p, int purpose, const float *pos, const float *nrm, uint64_t n, float *out)
	{
		const PlanetParams &P = p->params;
		for (uint64_t i = 0; i < n; i++)
		{
			const V3 position = v3(pos[i * 3 + 0], pos[i * 3 + 1], pos[i * 3 + 2]);
			const V3 normal = v3(nrm[i * 3 + 0], nrm[i * 3 + 1], nrm[i * 3 + 2]);
			Tile t = tile_init(position, normal, purpose);
			t.elevation = eval_field(P, purpose == MP_LAND ? FD_ELEVATION : FD_ELEVATION_RAW, position);
			float elevs[8];
			for (int k = 0; k < 8; k++)
				elevs[k] = eval_field(P, FD_ELEVATION, slope_probe(position, normal, k));
			t.slope = slope_from_probes(elevs);
			switch (P.coloring)
			{
				case CL_DEFAULT:
					coloring_run<CL_DEFAULT>(p->coloring, p->cellular, t);
					break;
				case CL_BARREN:
					coloring_run<CL_BARREN>(p->coloring, p->cellular, t);
					break;
				default:
					coloring_run<CL_DEBUG>(p->coloring, p->cellular, t);
					break;
			}
			float *o = out + i * 13;
			o[0] = t.albedo.x;
			o[1] = t.albedo.y;
			o[2] = t.albedo.z;
			o[3] = t.roughness;
			o[4] = t.metallic;
			o[5] = t.height;
			o[6] = t.elevation;
			o[7] = t.slope;
			o[8] = t.temperature;
			o[9] = t.precipitation;
			o[10] = t.opacity;
			o[11] = (float)t.biome;
			o[12] = (float)t.type;
		}
	}
	void upgh_color(int what, const float *in3, float *out3)
	{
		const V3 v = v3(in3[0], in3[1], in3[2]);
		V3 r = v;
		switch (what)
		{
			case 0:
				r = rgb_to_hsv(v);
				break;
			case 1:
				r = hsv_to_rgb(v);
				break;
			case 2:
				r = rgb_to_hsluv(v);
				break;
			case 3:
				r = hsluv_to_rgb(v);
				break;
		}
		out3[0] = r.x;
		out3[1] = r.y;
		out3[2] = r.z;
	}
	void upgh_color_interpolate(const float *a3, const float *b3, float f, float *out3)
	{
		const V3 r = interpolate_color(v3(a3[0], a3[1], a3[2]), v3(b3[0], b3[1], b3[2]), f);
		out3[0] = r.x;
		out3[1] = r.y;
		out3[2] = r.z;
	}
}
