// upg_planet_host.h -- HOST-side preparation of the per-planet parameter block (upg_planet.h):
// the seed stream, the per-mode noise configurations and the handful of shape constants the
// reference derives once per process (function-local statics).  Pure C++ (no CUDA), used by the
// C-ABI layer (upg_api.cu) and by the host unit tests.
//
// Seed semantics: the reference has no seed flag.  Every noise seed is the low
// 32 bits of successive outputs of a xoroshiro128** generator that is a COPY of the main thread's
// generator taken at the first noiseSeed() call (sources/math.cpp:73-77; externals/cage/sources/
// libcore/math/random.cpp:27-38).  `fibers` and `asteroid` additionally draw shape constants from
// the thread generator itself before that copy is made (sources/sdf.cpp:133,342).  The order in
// which seeds are consumed is the order in which terrainPreseed() (sources/modesConfiguration.cpp:
// 288-306) first passes each static-local declaration.  upg_planet_desc::seed is that generator state.
#pragma once

#include <cstdint>
#include <cstring>

#include "upg_terrain.cuh"

namespace upg
{
	// xoroshiro128** exactly as cage::RandomGenerator (random.cpp:27-51)
	struct Xoroshiro
	{
		uint64_t s[2];
		static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
		uint64_t next()
		{
			const uint64_t s0 = s[0];
			uint64_t s1 = s[1];
			const uint64_t result = rotl(s0 * 5, 7) * 9;
			s1 ^= s0;
			s[0] = rotl(s0, 24) ^ s1 ^ (s1 << 16);
			s[1] = rotl(s1, 37);
			return result;
		}
		double chanceDouble() { return (double)(next() >> 11) * 0x1.0p-53; }
		float chance() { return (float)chanceDouble(); }
		float range(float lo, float hi) { return lerp_(lo, hi, chance()); } // randomRange(Real, Real): interpolate(min, max, c)
		// randomRange(double, double) (random.cpp:78-85): the chance is narrowed to Real by interpolate(double, double, Real)
		double rangeDouble(double lo, double hi)
		{
			const float f = (float)chanceDouble();
			return lo + (hi - lo) * f;
		}
	};

	inline float fractal_bounding(int octaves, float gain)
	{
		const float g = gain < 0 ? -gain : gain;
		float amp = g;
		float total = 1.0f;
		for (int i = 1; i < octaves; i++)
		{
			total += amp;
			amp *= g;
		}
		return 1 / total;
	}

	// cage defaults: octaves 3, lacunarity 2, gain 0.5, frequency 1 (include/cage-core/noiseFunction.h:59-72)
	inline NoiseInst make_noise(uint32_t seed, float frequency = 1.0f, int octaves = 3, float gain = 0.5f, float lacunarity = 2.0f)
	{
		NoiseInst n;
		n.seed = (int32_t)seed;
		n.frequency = frequency;
		n.lacunarity = lacunarity;
		n.gain = gain;
		n.bounding = fractal_bounding(octaves, gain);
		return n;
	}

	inline VoronoiInst make_voronoi(uint32_t seed, uint32_t pointsPerCell, float cellSize)
	{
		VoronoiInst v;
		const uint32_t h1 = hash32(seed), h2 = hash32(h1), h3 = hash32(h2);
		v.seedHashes[0] = (int32_t)h1;
		v.seedHashes[1] = (int32_t)h2;
		v.seedHashes[2] = (int32_t)h3;
		v.pointsPerCell = pointsPerCell;
		v.cellSize = cellSize;
		v.frequency = 1 / cellSize;
		return v;
	}

	inline float degs_to_rads(float degs) { return degs * 3.14159274101257324f / 180; }

	inline TriangleConsts make_triangle(V3 a, V3 b, V3 c)
	{
		TriangleConsts t;
		t.v[0] = a;
		t.v[1] = b;
		t.v[2] = c;
		t.normal = normalize(cross(b - a, c - a)); // Triangle::normal(), geometry.h:87
		t.d = -dot(a, t.normal); // Plane(point, normal), shapes.cpp:131
		for (int i = 0; i < 3; i++)
		{
			// makeSegment(v[i], v[i+1]) -> Line(a, b - a, 0, 1).normalize(), shapes.cpp:55-96
			const V3 o = t.v[i];
			const V3 dir = t.v[(i + 1) % 3] - o;
			t.segOrigin[i] = o;
			t.segDir[i] = normalize(dir);
			t.segMax[i] = 1.0f * length(dir);
		}
		return t;
	}

	struct SeedLog
	{
		uint32_t seeds[128];
		int count = 0;
	};

	struct PlanetBuilder
	{
		Xoroshiro thread; // the "main thread" generator
		Xoroshiro stream; // noiseSeed()'s private copy
		bool streamStarted = false;
		SeedLog log;

		uint32_t noiseSeed()
		{
			if (!streamStarted)
			{
				stream = thread;
				streamStarted = true;
			}
			const uint32_t s = (uint32_t)stream.next();
			if (log.count < 128)
				log.seeds[log.count++] = s;
			return s;
		}

		void shapeStatics(PlanetParams &P)
		{
			ShapeConsts &S = P.sh;
			S.hexagonNormal = normalize(v3(1, 1, -1));
			S.knotInvLen = length(v3(1, 1, 0.3f));
			for (int k = 0; k < 3; k++)
				S.h3oCenters[k] = quat_rotate(quat_euler(0.0f, 0.0f, degs_to_rads((float)(k * 120))), v3(680, 0, 0));
			{
				const float radius = 900;
				const V3 top = v3(0, radius, 0);
				S.prism = make_triangle(top, quat_rotate(quat_euler(0.0f, 0.0f, degs_to_rads(120)), top), quat_rotate(quat_euler(0.0f, 0.0f, degs_to_rads(-120)), top));
			}
			{
				const V3 c[4] = { v3(1, 1, 1), v3(1, -1, -1), v3(-1, 1, -1), v3(-1, -1, 1) };
				S.tetra[0] = make_triangle(c[1], c[3], c[2]);
				S.tetra[1] = make_triangle(c[0], c[2], c[3]);
				S.tetra[2] = make_triangle(c[0], c[3], c[1]);
				S.tetra[3] = make_triangle(c[0], c[1], c[2]);
			}
			S.fibersOffset = v3(0, 0, 0);
			S.asteroid = make_noise(0);
			if (P.shape == SH_FIBERS)
			{
				const float x = thread.range(-100, 100);
				const float y = thread.range(-100, 100);
				const float z = thread.range(-100, 100);
				S.fibersOffset = v3(x, y, z);
			}
			if (P.shape == SH_ASTEROID)
			{
				const float freq = (float)thread.rangeDouble(0.65, 0.75); // literals are doubles at sources/sdf.cpp:342
				S.asteroid = make_noise(noiseSeed(), freq, 2, 0.3f);
			}
		}

		void elevationStatics(PlanetParams &P)
		{
			ElevationConsts &E = P.el;
			for (int i = 0; i < 4; i++)
				E.n[i] = make_noise(0);
			E.voronoi = make_voronoi(0, 1, 10);
			switch (P.elevation)
			{
				case EL_NONE:
					break;
				case EL_SIMPLE:
					E.n[0] = make_noise(noiseSeed(), 0.0005f, 6, 0.4f);
					break;
				case EL_LEGACY:
					E.n[0] = make_noise(noiseSeed(), 0.0005f, 4);
					E.n[1] = make_noise(noiseSeed(), 1.0f, 4);
					break;
				case EL_LAKES:
				case EL_ISLANDS:
					E.n[0] = make_noise(noiseSeed(), 0.0013f, 4);
					E.n[1] = make_noise(noiseSeed(), 0.0015f);
					E.n[2] = make_noise(noiseSeed(), 0.001f, 4, -0.4f, 1.5f);
					E.n[3] = make_noise(noiseSeed(), 0.002f, 3, 0.3f);
					break;
				case EL_CRATERS:
					E.voronoi = make_voronoi(noiseSeed(), 2, 300);
					E.n[0] = make_noise(noiseSeed(), 1.0f);
					break;
			}
		}
	};

	inline void set_grid(PlanetParams &P, float boxSize, int res)
	{
		// Aabb(Vec3(boxSize * -0.5), Vec3(boxSize * 0.5)), sources/meshGeneration.cpp:36
		const float a = boxSize * -0.5f, b = boxSize * 0.5f;
		P.grid.res = res;
		P.grid.boxA = a;
		P.grid.boxSpan = b - a;
		P.grid.resMinus5 = (float)res - 5.0f;
	}
}
