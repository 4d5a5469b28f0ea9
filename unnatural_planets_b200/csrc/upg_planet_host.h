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

#include <cmath>
#include <cstdint>
#include <cstring>

#include "upg_coloring.cuh"
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

	// cellular-noise variants keep their distance / return kind in the template arguments of the call site; the
	// instance only carries seed / frequency / fractal parameters
	inline NoiseInst make_cellular(uint32_t seed, float frequency, int octaves = 3) { return make_noise(seed, frequency, octaves); }

	// The colouring statics, in the order terrainPreseed() first passes their declarations (sources/modesConfiguration.cpp:
	// 288-306 -> sources/coloringModes.cpp:905-966): a LAND tile through the chosen colouring function, then a WATER tile.
	// Every generate* function declares its statics before any early return, so the order only depends on the colouring mode.
	inline void coloring_statics(PlanetBuilder &b, const PlanetParams &P, ColoringConsts &C)
	{
		std::memset(&C, 0, sizeof(C));
		C.poles = P.poles;
		C.flowers = P.flowers;
		const NoiseInst none = make_noise(0);
		NoiseInst *all = &C.beach;
		for (NoiseInst *n = all; n <= &C.waterHueShift; n++)
			*n = none;
		C.boulderCenters = C.flowerCluster = C.flowerCenter = make_voronoi(0, 1, 10);

		auto precipitation = [&]() { C.precp = make_noise(b.noiseSeed(), 0.001f, 2); };
		auto temperature = [&]()
		{
			C.temp = make_noise(b.noiseSeed(), 0.0003f, 2, 0.4f);
			C.polar = make_noise(b.noiseSeed(), 0.007f, 3);
		};
		auto bedrock = [&]()
		{
			const uint32_t seed = b.noiseSeed();
			C.bedrockDepth = make_noise(b.noiseSeed(), 0.0031f);
			C.bedrockFreq = make_noise(b.noiseSeed(), 0.013f);
			C.bedrockCracks = make_cellular(seed, 0.017f);
			C.bedrockValue = make_cellular(seed, 0.017f);
			C.bedrockSaturation = make_noise(b.noiseSeed(), 0.0022f, 2);
		};
		auto cliffs = [&]() { C.cliffsCracks = make_noise(b.noiseSeed(), 0.047f, 2); };
		auto mica = [&]()
		{
			C.micaMask = make_cellular(b.noiseSeed(), 0.02f);
			C.micaCracks = make_cellular(b.noiseSeed(), 0.05f);
		};
		auto dirt = [&]()
		{
			C.dirtHeight = make_noise(b.noiseSeed(), 0.05f, 3, 0.4f, 2.3f);
			C.dirtCracks = make_noise(b.noiseSeed(), 0.07f, 2);
			C.dirtCracksMask = make_noise(b.noiseSeed(), 0.02f, 4);
		};
		auto sand = [&]()
		{
			C.sandHeightScale = make_noise(b.noiseSeed(), 0.001f, 3);
			C.sandHeight = make_noise(b.noiseSeed(), 0.01f, 3, 0.7f);
			C.sandColor = make_noise(b.noiseSeed(), 0.0008f);
			C.sandHue = make_noise(b.noiseSeed(), 0.02f, 2);
		};
		auto grass = [&]()
		{
			const uint32_t seed = b.noiseSeed();
			C.grassHue = make_noise(b.noiseSeed(), 0.01f, 2);
			C.grassPatchesCells = make_cellular(seed, 0.1f);
			C.grassPatchesDist = make_cellular(seed, 0.1f);
			C.grassScratches = make_cellular(b.noiseSeed(), 0.3f);
			C.beach = make_noise(b.noiseSeed(), 0.035f); // beachMask's static, first reached from the grass mask (:511)
		};
		auto boulders = [&]()
		{
			C.boulderThreshold = make_noise(b.noiseSeed(), 0.01f);
			C.boulderCenters = make_voronoi(b.noiseSeed(), 2, 200);
			C.boulderSize = make_noise(b.noiseSeed(), 0.09f);
			C.boulderHue = make_noise(b.noiseSeed(), 0.3f);
			C.boulderValue = make_noise(b.noiseSeed(), 0.5f);
			C.boulderScratches = make_noise(b.noiseSeed(), 0.5f);
		};
		auto flowers = [&]()
		{
			C.flowerCluster = make_voronoi(b.noiseSeed(), 1, 200);
			C.flowerCenter = make_voronoi(b.noiseSeed(), 1, 35);
			C.flowerSize = make_noise(b.noiseSeed(), 0.35f);
			C.flowerColor = make_noise(b.noiseSeed(), 0.02f);
			C.flowerHue1 = make_noise(b.noiseSeed(), 0.005f);
			C.flowerHue2 = make_noise(b.noiseSeed(), 0.4f);
		};
		auto ice = [&]()
		{
			const uint32_t seed = b.noiseSeed();
			C.iceTempOffset = make_noise(b.noiseSeed(), 0.035f);
			C.iceFreq = make_noise(b.noiseSeed(), 0.01f, 3);
			C.iceThickness = make_cellular(seed, 0.05f, 2);
			C.iceCracks = make_cellular(seed, 0.05f, 2);
		};
		auto snow = [&]()
		{
			C.snowElevOffset = make_noise(b.noiseSeed(), 0.055f);
			C.snowThickness = make_noise(b.noiseSeed(), 0.05f, 3);
		};
		auto water = [&]()
		{
			C.waterColorSwitch = make_noise(b.noiseSeed(), 0.0015f, 2);
			C.waterHueShift = make_noise(b.noiseSeed(), 0.004f, 3);
		};

		switch (P.coloring)
		{
			case CL_DEFAULT:
				precipitation();
				temperature();
				bedrock();
				cliffs();
				mica();
				dirt();
				sand();
				grass();
				boulders();
				flowers();
				ice();
				snow();
				water();
				break;
			case CL_BARREN:
				bedrock();
				cliffs();
				mica();
				boulders();
				water();
				break;
			case CL_DEBUG:
				precipitation();
				temperature();
				break;
		}

		C.hsluvSandTarget = rgb_to_hsluv(v3(187, 104, 50) / 255);
		C.hsluvWaterDeep = rgb_to_hsluv(v3(40, 63, 168) / 255);
		C.hsluvWaterShallow = rgb_to_hsluv(v3(39, 154, 186) / 255);
		C.hsluvFlowerA = rgb_to_hsluv(v3(0.5f, 0, 0.35f));
		C.hsluvFlowerB = rgb_to_hsluv(v3(0.5f, 0.4f, 0));
	}

	// the declared substitute for FastNoiseLite's RandVecs3D constant table (same construction as
	// oracle/up_noise.h upo_randvecs3d): 256 unit vectors from a splitmix64 stream, w = 0
	inline void default_cellular_table(float *table /* 1024 */)
	{
		for (int e = 0; e < 256; e++)
		{
			uint64_t s = 0x5eedc0de00000000ull + (uint64_t)e;
			auto next = [&]()
			{
				uint64_t z = (s += 0x9e3779b97f4a7c15ull);
				z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
				z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
				return z ^ (z >> 31);
			};
			const double u = (double)(next() >> 11) * (1.0 / 9007199254740992.0);
			const double v = (double)(next() >> 11) * (1.0 / 9007199254740992.0);
			const double zz = 2.0 * u - 1.0;
			const double rr = std::sqrt(1.0 - zz * zz);
			const double ph = 6.283185307179586476925286766559 * v;
			table[e * 4 + 0] = (float)(rr * std::cos(ph));
			table[e * 4 + 1] = (float)(rr * std::sin(ph));
			table[e * 4 + 2] = (float)zz;
			table[e * 4 + 3] = 0.0f;
		}
	}

	inline void set_grid(PlanetParams &P, float boxSize, int res)
	{
		// Aabb(Vec3(boxSize * -0.5), Vec3(boxSize * 0.5)), sources/meshGeneration.cpp:36
		const float a = boxSize * -0.5f, b = boxSize * 0.5f;
		P.grid.res = res;
		P.grid.boxA = a;
		P.grid.boxSpan = b - a;
		P.grid.resMinus5 = (float)res - 5.0f;
	}

	// Bounds of elevation / 10 (the term the land and navigation fields subtract from the shape, sources/modesConfiguration.cpp:
	// 248-279, meshElevationRatio = 10) over ALL positions, per elevation mode (sources/elevationModes.cpp), widened to contain 0
	// so that the same pair also bounds max(elevation, 0) / 10.  Used by the band sampler (upg_band.cu) to decide where the sign
	// of a sample is already settled by the shape alone.  Derivation: every single-octave generator is bounded by S (Value is a
	// convex combination of hashes in [-1, 1); the Perlin / OpenSimplex2 suprema are 0.99999 / 0.99997, tools/noise_bounds.cpp;
	// S carries 2 % on top), an fBm sum by S (the amplitudes are normalised by `bounding`), a ridged sum by 2S - 1; the rest is
	// the monotone pieces of each mode followed through.  tests/test_host_terrain.py samples every mode against these numbers.
	inline void elevation_interval10(int elevation, double *lo10, double *hi10)
	{
		const double S = 1.02;
		const double F = 2 * S - 1;
		double a = 0, b = 0;
		switch (elevation)
		{
			case EL_NONE:
				a = b = 100;
				break;
			case EL_SIMPLE:
			{
				// a = -ridged + 0.3; a = (a * 1.3 - 0.35)^3 + 0.1; 100 - a * 1000
				const double tmin = (0.3 - F) * 1.3 - 0.35, tmax = (0.3 + F) * 1.3 - 0.35;
				a = 100 - (tmax * tmax * tmax + 0.1) * 1000;
				b = 100 - (tmin * tmin * tmin + 0.1) * 1000;
				break;
			}
			case EL_LEGACY:
				// a = fbm + 0.11; a < 0 ? -(-a)^0.85 : a^1.7; a * 2500
				a = -std::pow(S - 0.11, 0.85) * 2500;
				b = std::pow(S + 0.11, 1.7) * 2500;
				break;
			case EL_LAKES:
			case EL_ISLANDS:
			{
				// land = l / (|l| + 0.17) + 0.15 with l in [-1, 1], times 150; the mountains add smoothMax(0, max(ridge, terraces), 50)
				// which lies in [0, max(ridge, terraces) + 12.5]; ridge <= (F - 0.1)^1.6 * 1000, terraces <= (floor((S + 0.1) * 2.5) + 1) * 250
				const double ridge = std::pow(F - 0.1, 1.6) * 1000;
				const double terraces = (std::floor((S + 0.1) * 2.5) + 1) * 250;
				a = (-1 / 1.17 + 0.15) * 150;
				b = (1 / 1.17 + 0.15) * 150 + (ridge > terraces ? ridge : terraces) + 12.5;
				break;
			}
			case EL_CRATERS:
			{
				// h = (x^4 - 0.9 x^5 - 0.1) * 10 on x in [0, 1]: -1 at 0, maximum at x = 8 / 9; 200 + h * 190
				const double x = 8.0 / 9.0;
				a = 200 + (-0.1 * 10) * 190;
				b = 200 + (x * x * x * x - 0.9 * x * x * x * x * x - 0.1) * 10 * 190;
				break;
			}
		}
		const double slop = (b - a) * 0.01 + 0.5;
		a = (a - slop) / 10;
		b = (b + slop) / 10;
		*lo10 = a < 0 ? a : 0;
		*hi10 = b > 0 ? b : 0;
	}
}
