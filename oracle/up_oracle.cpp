/*
 * up_oracle.cpp -- TEST INFRASTRUCTURE: the CPU oracle for the generation hot path, stages 1 (field sampling).
 *
 * A plain, scalar, single-file RESTATEMENT of the reference's algorithm, written from a reading of the
 * reference sources; every function cites the file:line it follows.  It shares no code with the product
 * (unnatural_planets_b200/csrc) and is only ever called by tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * PINNING: validated bit-for-bit against oracle/_ref (the reference's own translation units compiled verbatim)
 * and against the committed golden vectors in tests/golden/, which oracle/_ref produced
 * (tests/test_oracle_port.py).  Noise VALUES remain "parity unpinned" against upstream FastNoiseLite, whose source
 * is absent (see up_noise.h).
 *
 * Build: g++ -O3 -ffp-contract=off (no FMA, like the reference's baseline x86-64 build).
 */
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "up_noise.h"

namespace
{
	// ------------------------------------------------------------------------------------------------------
	// cage scalar / vector semantics (externals/cage/sources/include/cage-core/math.h)
	// ------------------------------------------------------------------------------------------------------
	typedef float real;
	struct vec2
	{
		real v[2];
		real &operator[](int i) { return v[i]; }
		real operator[](int i) const { return v[i]; }
	};
	struct vec3
	{
		real v[3];
		real &operator[](int i) { return v[i]; }
		real operator[](int i) const { return v[i]; }
	};
	struct vec4
	{
		real v[4];
		real &operator[](int i) { return v[i]; }
		real operator[](int i) const { return v[i]; }
	};

	inline vec2 V2(real a, real b) { return vec2{ { a, b } }; }
	inline vec3 V3(real a, real b, real c) { return vec3{ { a, b, c } }; }
	inline vec3 V3s(real a) { return vec3{ { a, a, a } }; }
	inline vec4 V4(real a, real b, real c, real d) { return vec4{ { a, b, c, d } }; }

	// math.h:1021-1033 -- ternaries, not fmin/fmax
	inline real rmin(real a, real b) { return a < b ? a : b; }
	inline real rmax(real a, real b) { return a > b ? a : b; }
	inline real rclamp(real v, real a, real b) { return rmin(rmax(v, a), b); }
	inline real rsat(real v) { return rclamp(v, 0, 1); }
	inline real rabs(real a) { return a < 0 ? -a : a; } // math.h:888-892
	inline int rsign(real x) { return x < 0 ? -1 : x > 0 ? 1 : 0; } // math.h:806-809
	inline real rmod(real l, real r) { return l - (r * (real)(int32_t)(l / r)); } // math.h:442-446
	inline real rmix(real a, real b, real f) { return a + (b - a) * f; } // interpolate, math.h:1173-1177
	inline real rsmooth(real x) { return x * x * (3 - 2 * x); } // math.h:798-801
	inline real rsmoother(real x) { return x * x * x * (x * (x * 6 - 15) + 10); } // math.h:802-805
	const real PI = (real)3.141592653589793238; // Real::Pi()
	inline real degs(real d) { return d * PI / 180; } // Rads(Degs), math.h:781
	inline real atan2_cage(real x, real y) { return std::atan2(y, x); } // scalars.cpp:84-87

#define VOP(T, N) \
	inline T operator+(T a, T b) \
	{ \
		T r; \
		for (int i = 0; i < N; i++) \
			r[i] = a[i] + b[i]; \
		return r; \
	} \
	inline T operator-(T a, T b) \
	{ \
		T r; \
		for (int i = 0; i < N; i++) \
			r[i] = a[i] - b[i]; \
		return r; \
	} \
	inline T operator*(T a, T b) \
	{ \
		T r; \
		for (int i = 0; i < N; i++) \
			r[i] = a[i] * b[i]; \
		return r; \
	} \
	inline T operator*(T a, real s) \
	{ \
		T r; \
		for (int i = 0; i < N; i++) \
			r[i] = a[i] * s; \
		return r; \
	} \
	inline T operator*(real s, T a) \
	{ \
		T r; \
		for (int i = 0; i < N; i++) \
			r[i] = s * a[i]; \
		return r; \
	} \
	inline T operator/(T a, real s) \
	{ \
		T r; \
		for (int i = 0; i < N; i++) \
			r[i] = a[i] / s; \
		return r; \
	} \
	inline T operator-(T a, real s) \
	{ \
		T r; \
		for (int i = 0; i < N; i++) \
			r[i] = a[i] - s; \
		return r; \
	} \
	inline T vabs(T a) \
	{ \
		T r; \
		for (int i = 0; i < N; i++) \
			r[i] = rabs(a[i]); \
		return r; \
	} \
	inline T vmax(T a, real s) \
	{ \
		T r; \
		for (int i = 0; i < N; i++) \
			r[i] = rmax(a[i], s); \
		return r; \
	} \
	inline real dot(T a, T b) \
	{ \
		real sum = 0; /* math.h:915-925: products summed left to right starting from zero */ \
		for (int i = 0; i < N; i++) \
			sum += a[i] * b[i]; \
		return sum; \
	} \
	inline real len2(T a) { return dot(a, a); } \
	inline real len(T a) { return std::sqrt(dot(a, a)); } \
	inline T unit(T a) { return a / len(a); }
	VOP(vec2, 2)
	VOP(vec3, 3)
	VOP(vec4, 4)
#undef VOP
	inline real dist2(vec3 a, vec3 b) { return len2(a - b); }
	inline real dist(vec3 a, vec3 b) { return std::sqrt(dist2(a, b)); }
	inline vec3 cross(vec3 l, vec3 r) { return V3(l[1] * r[2] - l[2] * r[1], l[2] * r[0] - l[0] * r[2], l[0] * r[1] - l[1] * r[0]); }

	// Quat(pitch, yaw, roll) and Quat * Vec3 (libcore/math/vectors.cpp:118-133, 189-193)
	inline vec4 quatEuler(real pitch, real yaw, real roll)
	{
		const vec3 e = V3(pitch, yaw, roll) * (real)0.5;
		real c[3], s[3];
		for (int i = 0; i < 3; i++)
		{
			c[i] = std::cos(e[i]);
			s[i] = std::sin(e[i]);
		}
		vec4 q;
		q[0] = s[0] * c[1] * c[2] - c[0] * s[1] * s[2];
		q[1] = c[0] * s[1] * c[2] + s[0] * c[1] * s[2];
		q[2] = c[0] * c[1] * s[2] - s[0] * s[1] * c[2];
		q[3] = c[0] * c[1] * c[2] + s[0] * s[1] * s[2];
		return q / len(q);
	}
	inline vec3 quatApply(vec4 q, vec3 r)
	{
		const vec3 u = V3(q[0], q[1], q[2]);
		const vec3 t = cross(u, r) * (real)2;
		return r + t * q[3] + cross(u, t);
	}

	// ------------------------------------------------------------------------------------------------------
	// sources/math.cpp
	// ------------------------------------------------------------------------------------------------------
	inline real smoothMin(real a, real b, real k) // :28-33
	{
		const real h = rsat((b - a) / k * (real)0.5 + (real)0.5);
		return rmix(b, a, h) - k * h * (1 - h);
	}
	inline real smoothMax(real a, real b, real k) { return -smoothMin(-a, -b, k); } // :35-38
	inline real terrace(real x, real steepness) // :18-26
	{
		const real f = std::floor(x);
		const real t = rsmoother(rsat((x - f) * rmax(steepness, 1))) + f;
		return rmix(x, t, rsat(steepness));
	}

	// ------------------------------------------------------------------------------------------------------
	// random stream (externals/cage/sources/libcore/math/random.cpp:27-51, 64-85; sources/math.cpp:73-77)
	// ------------------------------------------------------------------------------------------------------
	struct Rng
	{
		uint64_t a, b;
		uint64_t next()
		{
			const uint64_t s0 = a;
			uint64_t s1 = b;
			const uint64_t x = s0 * 5;
			const uint64_t out = ((x << 7) | (x >> 57)) * 9;
			s1 ^= s0;
			a = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
			b = (s1 << 37) | (s1 >> 27);
			return out;
		}
		double chanceD() { return (double)(next() >> 11) * 0x1.0p-53; }
		real chance() { return (real)chanceD(); }
		real rangeReal(real lo, real hi) { return rmix(lo, hi, chance()); }
		double rangeDouble(double lo, double hi) // interpolate(double, double, Real): the chance is narrowed to float
		{
			const real c = (real)chanceD();
			return lo + (hi - lo) * (double)c;
		}
	};

	// ------------------------------------------------------------------------------------------------------
	// planet state: what the reference keeps in function-local statics
	// ------------------------------------------------------------------------------------------------------
	struct Tri
	{
		vec3 p[3];
	};
	struct Voronoi
	{
		int32_t seedHashes[3];
		uint32_t pointsPerCell;
		real cellSize, frequency;
	};

	enum
	{
		S_ASTEROID, S_BELT, S_BOWL, S_BOX, S_BUNNY, S_CAPSULE, S_CUBE, S_DISK, S_DOUBLETORUS, S_FIBERS, S_GEAR, S_H2O, S_H3O, S_H4O, S_HEMISPHERES, S_HEXAGON, S_HEXAGONALPRISM, S_INSIDECUBE, S_KNOT, S_MANDELBULB, S_MOBIUSSTRIP, S_MONKEYHEAD, S_OCTAHEDRON, S_PIPE, S_SPHERE, S_SQUARE, S_TETRAHEDRON, S_TORUS, S_TORUSCROSS, S_TRIANGULARPRISM, S_TUBE, S_TWISTEDHEXAGONALPRISM, S_TWISTEDPLANE, S_WORMHOLE, S_COUNT
	};
	enum
	{
		E_NONE, E_SIMPLE, E_LEGACY, E_LAKES, E_ISLANDS, E_CRATERS, E_COUNT
	};

	struct Planet
	{
		int shape, elevation;
		Rng thread; // the main thread's generator
		Rng seeds; // noiseSeed()'s private copy (sources/math.cpp:75)
		bool seedsStarted;
		std::vector<uint32_t> seedLog;
		// shape statics
		vec3 fibersOffset;
		upo_noise_cfg asteroidNoise;
		// elevation statics, in declaration order
		upo_noise_cfg en[4];
		Voronoi craters;

		uint32_t noiseSeed()
		{
			if (!seedsStarted)
			{
				seeds = thread;
				seedsStarted = true;
			}
			const uint32_t s = (uint32_t)seeds.next();
			seedLog.push_back(s);
			return s;
		}
	};

	upo_noise_cfg noiseCfg(uint32_t seed, int type, int fractal, real frequency, int octaves = 3, real gain = 0.5f, real lacunarity = 2)
	{
		upo_noise_cfg c = upo_noise_default(seed);
		c.type = type;
		c.fractal = fractal;
		c.frequency = frequency;
		c.octaves = octaves;
		c.gain = gain;
		c.lacunarity = lacunarity;
		upo_noise_finish(&c);
		return c;
	}

	inline uint32_t murmurFinal(uint32_t k) // cage hash(), math.h:1231-1239
	{
		k ^= k >> 16;
		k *= 0x85ebca6b;
		k ^= k >> 13;
		k *= 0xc2b2ae35;
		k ^= k >> 16;
		return k;
	}

	Voronoi makeVoronoi(uint32_t seed, uint32_t ppc, real cellSize) // sources/voronoi.cpp:17
	{
		Voronoi v;
		v.seedHashes[0] = (int32_t)murmurFinal(seed);
		v.seedHashes[1] = (int32_t)murmurFinal(murmurFinal(seed));
		v.seedHashes[2] = (int32_t)murmurFinal(murmurFinal(murmurFinal(seed)));
		v.pointsPerCell = ppc;
		v.cellSize = cellSize;
		v.frequency = 1 / cellSize;
		return v;
	}

	// terrainPreseed (sources/modesConfiguration.cpp:288-306): statics are initialised in the order control first
	// passes them: shape function first, then the elevation function (terrainSdfLand, :248-253)
	void preseed(Planet &P)
	{
		P.fibersOffset = V3(0, 0, 0);
		if (P.shape == S_FIBERS) // sources/sdf.cpp:133 -- three draws from the THREAD generator
		{
			const real x = P.thread.rangeReal(-100, 100);
			const real y = P.thread.rangeReal(-100, 100);
			const real z = P.thread.rangeReal(-100, 100);
			P.fibersOffset = V3(x, y, z);
		}
		if (P.shape == S_ASTEROID) // sources/sdf.cpp:335-345 -- frequency from the thread generator, then the first noiseSeed()
		{
			const real freq = (real)P.thread.rangeDouble(0.65, 0.75);
			P.asteroidNoise = noiseCfg(P.noiseSeed(), UPO_NOISE_SIMPLEX, UPO_FRACTAL_FBM, freq, 2, 0.3f);
		}
		switch (P.elevation)
		{
			case E_SIMPLE: // sources/elevationModes.cpp:16-26
				P.en[0] = noiseCfg(P.noiseSeed(), UPO_NOISE_SIMPLEX, UPO_FRACTAL_RIDGED, 0.0005f, 6, 0.4f);
				break;
			case E_LEGACY: // :36-55
				P.en[0] = noiseCfg(P.noiseSeed(), UPO_NOISE_VALUE, UPO_FRACTAL_FBM, 0.0005f, 4);
				P.en[1] = noiseCfg(P.noiseSeed(), UPO_NOISE_VALUE, UPO_FRACTAL_FBM, 1, 4);
				break;
			case E_LAKES: // :131-140 then :70-100
			case E_ISLANDS: // :153-162 then :70-100
				P.en[0] = noiseCfg(P.noiseSeed(), UPO_NOISE_VALUE, UPO_FRACTAL_FBM, 0.0013f, 4);
				P.en[1] = noiseCfg(P.noiseSeed(), UPO_NOISE_PERLIN, UPO_FRACTAL_NONE, 0.0015f);
				P.en[2] = noiseCfg(P.noiseSeed(), UPO_NOISE_SIMPLEX, UPO_FRACTAL_RIDGED, 0.001f, 4, -0.4f, 1.5f);
				P.en[3] = noiseCfg(P.noiseSeed(), UPO_NOISE_PERLIN, UPO_FRACTAL_FBM, 0.002f, 3, 0.3f);
				break;
			case E_CRATERS: // :175-190
				P.craters = makeVoronoi(P.noiseSeed(), 2, 300);
				P.en[0] = noiseCfg(P.noiseSeed(), UPO_NOISE_VALUE, UPO_FRACTAL_NONE, 1);
				break;
			default:
				break;
		}
	}

	inline real noise(const upo_noise_cfg &c, vec3 p) { return upo_noise3(&c, p[0], p[1], p[2]); }

	// ------------------------------------------------------------------------------------------------------
	// cage primitives (externals/cage/sources/libcore/signedDistanceFunctions.cpp)
	// ------------------------------------------------------------------------------------------------------
	real primPlane(vec3 pos, vec3 normal, real d) // :8-13
	{
		const vec3 c = normal * d;
		return dot(normal, pos - c);
	}
	real primSphere(vec3 pos, real radius) { return len(pos) - radius; } // :15-18
	real primCapsule(vec3 pos, real prolong, real radius) // :20-26
	{
		vec3 p = pos;
		p[2] += prolong * (real)0.5;
		p[2] -= rclamp(p[2], 0, prolong);
		return len(p) - radius;
	}
	real primCylinder(vec3 pos, real halfHeight, real radius) // :28-32
	{
		const vec2 d = vabs(V2(len(V2(pos[0], pos[1])), pos[2])) - V2(radius, halfHeight);
		return rmin(rmax(d[0], d[1]), 0) + len(vmax(d, 0));
	}
	real primBox(vec3 pos, vec3 radius) // :34-38
	{
		const vec3 p = vabs(pos) - radius;
		return len(vmax(p, 0)) + rmin(rmax(p[0], rmax(p[1], p[2])), 0);
	}

	// closest point machinery (externals/cage/sources/libcore/geometry/geometry.cpp:638-679, shapes.cpp:55-135)
	struct PlaneEq
	{
		vec3 n;
		real d;
	};
	PlaneEq planeOf(const Tri &t) // Plane(Triangle) -> Plane(t[0], t.normal())
	{
		PlaneEq p;
		p.n = unit(cross(t.p[1] - t.p[0], t.p[2] - t.p[0]));
		p.d = -dot(t.p[0], p.n);
		return p;
	}
	vec3 closestPlane(vec3 pt, const PlaneEq &pl) // geometry.cpp:675-679
	{
		const vec3 origin = pl.n * -pl.d;
		return pt - dot(pl.n, pt - origin) * pl.n;
	}
	vec3 closestSegment(vec3 a, vec3 b, vec3 p) // makeSegment + closestPoint(Vec3, Line): geometry.cpp:638-643
	{
		const vec3 raw = b - a;
		const vec3 dir = unit(raw);
		const real maximum = 1 * len(raw); // Line::normalize scales [0, 1] by the un-normalised length
		real d = dot(p - a, dir);
		d = rclamp(d, 0, maximum);
		return a + dir * d;
	}
	vec3 closestTriangle(vec3 source, const Tri &t) // geometry.cpp:645-673
	{
		const vec3 p = closestPlane(source, planeOf(t));
		{
			const vec3 a = t.p[0] - p, b = t.p[1] - p, c = t.p[2] - p;
			const vec3 u = cross(a, b), v = cross(b, c), w = cross(c, a);
			if (dot(u, v) > 0 && dot(u, w) > 0)
				return p;
		}
		const vec3 pab = closestSegment(t.p[0], t.p[1], p);
		const vec3 pbc = closestSegment(t.p[1], t.p[2], p);
		const vec3 pca = closestSegment(t.p[2], t.p[0], p);
		const real dab = dist2(p, pab), dbc = dist2(p, pbc), dca = dist2(p, pca);
		const real dm = rmin(rmin(dab, dbc), dca);
		if (dm == dab)
			return pab;
		if (dm == dbc)
			return pbc;
		return pca;
	}
	real distTriangle(vec3 a, const Tri &t) { return dist(closestTriangle(a, t), a); } // geometry.cpp:67-70

	real primTetrahedron(vec3 pos, real radius) // :40-63
	{
		const vec3 c[4] = { V3(1, 1, 1), V3(1, -1, -1), V3(-1, 1, -1), V3(-1, -1, 1) };
		const Tri tris[4] = { { { c[1], c[3], c[2] } }, { { c[0], c[2], c[3] } }, { { c[0], c[3], c[1] } }, { { c[0], c[1], c[2] } } };
		const vec3 p = pos / radius;
		real mad = -INFINITY;
		for (int i = 0; i < 4; i++)
		{
			const PlaneEq pl = planeOf(tris[i]);
			const vec3 k = closestPlane(p, pl);
			const real d = distTriangle(p, tris[i]) * (real)rsign(dot(pl.n, p - k));
			mad = rmax(mad, d);
		}
		return mad * radius;
	}
	real primOctahedron(vec3 pos, real radius) // :65-81
	{
		const vec3 p = vabs(pos);
		const real m = p[0] + p[1] + p[2] - radius;
		vec3 q;
		if ((real)3.0 * p[0] < m)
			q = p;
		else if ((real)3.0 * p[1] < m)
			q = V3(p[1], p[2], p[0]);
		else if ((real)3.0 * p[2] < m)
			q = V3(p[2], p[0], p[1]);
		else
			return m * (real)0.57735027;
		const real k = rclamp((real)0.5 * (q[2] - q[1] + radius), 0, radius);
		return len(V3(q[0], q[1] - radius + k, q[2] - k));
	}
	real primHexPrism(vec3 pos, real halfHeight, real radius) // :83-90
	{
		const vec3 k = V3((real)-0.8660254, (real)0.5, (real)0.57735);
		vec3 p = vabs(pos);
		const vec2 k2 = V2(k[0], k[1]);
		const vec2 sub = (real)2.0 * rmin(dot(k2, V2(p[0], p[1])), 0) * k2;
		p = p - V3(sub[0], sub[1], 0);
		const vec2 d = V2(len(V2(p[0], p[1]) - V2(rclamp(p[0], -k[2] * radius, k[2] * radius), radius)) * (real)rsign(p[1] - radius), p[2] - halfHeight);
		return rmin(rmax(d[0], d[1]), 0) + len(vmax(d, 0));
	}
	real primTorus(vec3 pos, real major, real minor) // :92-96
	{
		const vec2 q = V2(len(V2(pos[0], pos[1])) - major, pos[2]);
		return len(q) - minor;
	}
	real primKnot(vec3 pos, real scale, real k) // :98-115
	{
		const real TAU = PI * 2;
		const real norm = (real)(2 / 14.4);
		scale *= norm;
		vec3 p = pos / scale;
		const real r = len(V2(p[0], p[1]));
		real a = atan2_cage(p[0], p[1]);
		const real oa = k * a;
		a = rmod(a, (real)0.001 * TAU) - (real)0.001 * TAU / 2;
		p[0] = r * std::cos(a) - 5;
		p[1] = r * std::sin(a);
		const vec2 p2 = std::cos(oa) * V2(p[0], p[2]) + std::sin(oa) * V2(-p[2], p[0]);
		p[0] = rabs(p2[0]) - (real)1.5;
		p[2] = p2[1];
		return (len(p) - (real)0.7) * scale;
	}

	// ------------------------------------------------------------------------------------------------------
	// neural shapes (sources/sdf.cpp:243-324); weights are generated data (tools/gen_neural_weights.py)
	// ------------------------------------------------------------------------------------------------------
	struct NeuralWeights
	{
		float first[4][4][4];
		float hidden[2][4][4][16];
		float bias[2][4][4];
		float out[4][4];
		float outBias;
		float xSign;
		float div[2];
	};
#define UPG_CONST static const
#include "up_neural_weights.inc"
#undef UPG_CONST

	vec4 matVec(const float *m, vec4 r) // Mat4 * Vec4, libcore/math/matrices.cpp:156-165
	{
		vec4 res = V4(0, 0, 0, 0);
		for (int i = 0; i < 4; i++)
			for (int j = 0; j < 4; j++)
				res[i] += r[j] * m[j * 4 + i];
		return res;
	}
	vec4 sinv(vec4 a)
	{
		for (int i = 0; i < 4; i++)
			a[i] = std::sin(a[i]);
		return a;
	}
	real neural(const NeuralWeights &w, vec3 p)
	{
		if (len(p) > 1)
			return len(p) - (real)0.8;
		vec4 f[4], g[4];
		for (int b = 0; b < 4; b++)
		{
			const vec4 wy = V4(w.first[b][0][0], w.first[b][0][1], w.first[b][0][2], w.first[b][0][3]);
			const vec4 wz = V4(w.first[b][1][0], w.first[b][1][1], w.first[b][1][2], w.first[b][1][3]);
			const vec4 wx = V4(w.first[b][2][0], w.first[b][2][1], w.first[b][2][2], w.first[b][2][3]);
			const vec4 wb = V4(w.first[b][3][0], w.first[b][3][1], w.first[b][3][2], w.first[b][3][3]);
			vec4 a = p[1] * wy + p[2] * wz;
			a = w.xSign < 0 ? a - p[0] * wx : a + p[0] * wx;
			f[b] = sinv(a + wb);
		}
		for (int layer = 0; layer < 2; layer++)
		{
			for (int o = 0; o < 4; o++)
			{
				vec4 acc = matVec(w.hidden[layer][o][0], f[0]) + matVec(w.hidden[layer][o][1], f[1]) + matVec(w.hidden[layer][o][2], f[2]) + matVec(w.hidden[layer][o][3], f[3]);
				acc = acc + V4(w.bias[layer][o][0], w.bias[layer][o][1], w.bias[layer][o][2], w.bias[layer][o][3]);
				g[o] = sinv(acc) / w.div[layer] + f[o];
			}
			for (int o = 0; o < 4; o++)
				f[o] = g[o];
		}
		real r = dot(f[0], V4(w.out[0][0], w.out[0][1], w.out[0][2], w.out[0][3]));
		for (int o = 1; o < 4; o++)
			r = r + dot(f[o], V4(w.out[o][0], w.out[o][1], w.out[o][2], w.out[o][3]));
		return r + w.outBias;
	}

	// ------------------------------------------------------------------------------------------------------
	// the 34 base shapes (sources/sdf.cpp)
	// ------------------------------------------------------------------------------------------------------
	vec3 swapYZ(vec3 p) { return V3(p[0], p[2], p[1]); }

	real shapeMobius(vec3 pos, real radius, real majorAxis, real minorAxis) // :91-113
	{
		const real planeRotation = atan2_cage(pos[0], pos[2]);
		const vec2 xz = V2(pos[0], pos[2]);
		const vec2 proj = V2(dot(xz, unit(xz)), pos[1]);
		const vec2 point = proj + V2(-radius, 0);
		real a = atan2_cage(point[0], point[1]);
		a += planeRotation / 2;
		const vec2 p = V2(std::cos(a), std::sin(a)) * len(point);
		const vec2 d = vabs(p) - V2(majorAxis, minorAxis);
		return len(vmax(d, 0)) + rmin(rmax(d[0], d[1]), 0);
	}

	real shapeGear(vec3 pos_) // :180-204
	{
		const real scale = 2;
		const vec3 pos = swapYZ(pos_) / scale;
		const vec2 p = V2(pos[0], pos[2]);
		const uint32_t teeths = 9;
		const real angle = (PI * 2) / (real)teeths;
		const uint32_t sector = (uint32_t)(atan2_cage(p[0], p[1]) / angle + (real)teeths + (real)0.5) % teeths;
		const real ra = (real)sector * -angle;
		const real s = std::sin(ra), c = std::cos(ra);
		const vec2 q = V2(c * p[0] - s * p[1], s * p[0] + c * p[1]);
		const real b = len(vmax(vabs(q - V2(700, 0)) - V2(150, 70), 0));
		const real cc = rabs(len(p) - 450) - 150;
		const real h = rabs(pos[1]) - 30;
		return (smoothMax(rmin(b, cc), h, 50) - 50) * scale;
	}

	real shapeMandelbulb(vec3 pos_) // :206-232
	{
		if (len2(pos_) < (real)1e-3)
			return 0;
		const vec3 pos = pos_ * (real)0.0004;
		volatile int power = 3; // keep pow() a run-time call, as in the reference (cage::pow lives in another TU)
		const real Power = (real)power;
		vec3 z = pos;
		real dr = 1.0f, r = 0.0f;
		for (uint32_t i = 0; i < 20; i++)
		{
			r = len(z);
			if (r > (real)4.0)
				break;
			real theta = std::acos(z[2] / r);
			real phi = atan2_cage(z[1], z[0]);
			dr = std::pow(r, Power - (real)1.0) * Power * dr + (real)1.0;
			const real zr = std::pow(r, Power);
			theta = theta * Power;
			phi = phi * Power;
			z = zr * V3(std::sin(theta) * std::cos(phi), std::sin(phi) * std::sin(theta), std::cos(theta));
			z = z + pos;
		}
		const real value = (real)0.5 * std::log(r) * r / dr;
		return (value + (real)0.1) * 150;
	}

	vec3 hemiRotate(vec3 p) // :376-380
	{
		const real s = (real)0.707106781186547524400844362104;
		return V3((p[0] - p[1]) * s, (p[0] + p[1]) * s, p[2]);
	}

	real shape(const Planet &P, vec3 pos)
	{
		switch (P.shape)
		{
			case S_HEXAGON: // :10-13   Plane(Vec3(), n): d = -dot(0, n)
			{
				const vec3 n = unit(V3(1, 1, -1));
				return primPlane(pos, n, -dot(V3(0, 0, 0), n));
			}
			case S_SQUARE: // :15-18
			{
				const vec3 n = V3(0, 0, -1);
				return primPlane(pos, n, -dot(V3(0, 0, 0), n));
			}
			case S_SPHERE: // :20-23
				return primSphere(pos, 1100);
			case S_CAPSULE: // :25-28
				return primCapsule(swapYZ(pos), 2500, 600);
			case S_TUBE: // :30-33
				return primCylinder(swapYZ(pos), 10000, 800);
			case S_DISK: // :35-38
				return primCylinder(pos, 200, 600) - 100;
			case S_BOX: // :40-43
				return primBox(pos, V3(550, 1000, 550)) - 200;
			case S_CUBE: // :45-48
				return primBox(pos, V3s(450)) - 100;
			case S_TETRAHEDRON: // :50-53
				return primTetrahedron(pos, 600) - 100;
			case S_OCTAHEDRON: // :55-58
				return primOctahedron(pos, 800) - 100;
			case S_TRIANGULARPRISM: // :60-74
			{
				const real height = 300, radius = 900;
				const vec3 top = V3(0, radius, 0);
				const Tri t = { { top, quatApply(quatEuler(0, 0, degs(120)), top), quatApply(quatEuler(0, 0, degs(-120)), top) } };
				vec3 p = swapYZ(pos);
				p[2] = rmax(rabs(p[2]) - height * (real)0.5, 0);
				return distTriangle(p, t) - 100;
			}
			case S_HEXAGONALPRISM: // :76-79
				return primHexPrism(swapYZ(pos), 300, 400) - 100;
			case S_TORUS: // :81-84
				return primTorus(pos, 950, 530);
			case S_KNOT: // :86-90
			{
				const vec3 scale = V3(1, 1, (real)0.3);
				return primKnot(pos * scale, 1600, (real)1.5) / len(scale);
			}
			case S_MOBIUSSTRIP: // :115-118
				return shapeMobius(swapYZ(pos), 1000, 400, 40) - 200;
			case S_FIBERS: // :120-139
			{
				const real scale = (real)0.0007;
				const vec3 p = pos * scale + P.fibersOffset;
				const real gs = (real)3.23;
				const vec3 q = p * gs;
				const vec3 a = V3(std::sin(q[0]), std::sin(q[1]), std::sin(q[2]));
				const vec3 b = V3(std::cos(q[2]), std::cos(q[0]), std::cos(q[1]));
				const real g = rabs(dot(a, b) + (real)1.6) / gs - (real)0.2;
				const real v = g * (real)0.7 / scale;
				const real d = rmax(len(pos) - 1500, 0) * (real)0.5;
				return v + d;
			}
			case S_H2O: // :141-149
			{
				const real scale = (real)0.75;
				const vec3 p = pos * scale;
				const real h1 = primSphere(p - V3(-550, 300, 0), 450);
				const real h2 = primSphere(p - V3(+550, 300, 0), 450);
				const real o = primSphere(p - V3(0, -100, 0), 650);
				return smoothMin(o, rmin(h1, h2), 150) / scale;
			}
			case S_H3O: // :151-163
			{
				const real scale = (real)0.75;
				const vec3 p = pos * scale;
				real hs[3];
				for (int k = 0; k < 3; k++)
					hs[k] = primSphere(p - quatApply(quatEuler(0, 0, degs((real)(120 * k))), V3(680, 0, 0)), 450);
				const real o = primSphere(p, 650);
				const real h = rmin(rmin(hs[0], hs[1]), hs[2]);
				return smoothMin(o, h, 150) / scale;
			}
			case S_H4O: // :165-178
			{
				const real scale = (real)0.75;
				const vec3 p = pos * scale;
				const real h0 = primSphere(p - V3(-550, +400, 0), 450);
				const real h1 = primSphere(p - V3(+550, +400, 0), 450);
				const real h2 = primSphere(p - V3(0, -400, -550), 450);
				const real h3 = primSphere(p - V3(0, -400, +550), 450);
				const real o = primSphere(p, 650);
				const real h = rmin(rmin(h0, h1), rmin(h2, h3));
				return smoothMin(o, h, 150) / scale;
			}
			case S_GEAR:
				return shapeGear(pos);
			case S_MANDELBULB:
				return shapeMandelbulb(pos);
			case S_TWISTEDHEXAGONALPRISM: // :234-241
			{
				const real angle = pos[1] * (real)0.001;
				vec3 p = V3(pos[0], pos[2], 0);
				p = quatApply(quatEuler(0, 0, angle), p);
				p = V3(p[0], pos[1], p[1]);
				return primHexPrism(swapYZ(p), 900, 650) - 200;
			}
			case S_BUNNY: // :313-317
			{
				const real scale = (real)0.0005;
				return neural(g_bunnyWeights, swapYZ(pos) * scale) / scale;
			}
			case S_MONKEYHEAD: // :319-323
			{
				const real scale = (real)0.0005;
				return neural(g_monkeyWeights, swapYZ(pos) * scale) / scale;
			}
			case S_DOUBLETORUS: // :325-332
				return smoothMin(primTorus(V3(pos[1], pos[2], pos[0]), 900, 400), primTorus(V3(pos[1], pos[0], pos[2]), 900, 400), 200);
			case S_TORUSCROSS: // :334-341
			{
				const real ma = 750, mi = 400;
				return smoothMin(primTorus(V3(pos[1] - ma, pos[2], pos[0]), ma, mi), primTorus(V3(pos[1] + ma, pos[0], pos[2]), ma, mi), 200);
			}
			case S_BOWL: // :343-346
				return 2000 - dist(pos, V3(0, 0, -3000));
			case S_INSIDECUBE: // :348-351
				return -(primBox(pos, V3s(1300)) - 200);
			case S_ASTEROID: // :353-372
			{
				const real l = len(pos);
				if (l < 1)
					return -100;
				const vec3 d = pos / l;
				const real n = noise(P.asteroidNoise, d);
				return l + n * 600 - 1300;
			}
			case S_PIPE: // :374-385
			{
				vec3 p = pos;
				if (p[0] < 0)
				{
					p[0] *= -1;
					std::swap(p[1], p[2]);
				}
				p[1] = rabs(p[1]) - 800;
				std::swap(p[1], p[2]);
				return primTorus(p, 800, 400);
			}
			case S_TWISTEDPLANE: // :387-392
			{
				const real rot = 1 - (1 - rmin(1, rabs(pos[2]) / 1000)) * (1 - rmin(1, rabs(pos[2]) / 1000));
				const real angle = degs((real)rsign(pos[2]) * rot * 35);
				return pos[1] * std::cos(angle) + pos[0] * std::sin(angle);
			}
			case S_BELT: // :394-400
			{
				const real Radius = 1800;
				if (pos[0] > 0)
					return primCylinder(swapYZ(pos), 10000, Radius) * -1;
				return Radius - rabs(pos[2]);
			}
			case S_HEMISPHERES: // :402-416
			{
				const vec3 p = V3(rabs(pos[0]), rabs(pos[1]), pos[2]);
				const vec3 q = hemiRotate(p + V3(-2000, 0, 0));
				const real c = primCylinder(V3(q[1], q[2], q[0]), 10000, 200);
				const real s1 = primSphere(q, 1000);
				const real s2 = primSphere(hemiRotate(p + V3(0, -2000, 0)), 1000);
				return smoothMin(c, rmin(s1, s2), 100);
			}
			case S_WORMHOLE: // :418-423
			{
				const real x = rmin(len(V2(pos[0], pos[2])) - 1400, 0);
				return 700 - len(V2(x, pos[1]));
			}
		}
		return NAN;
	}

	// ------------------------------------------------------------------------------------------------------
	// Voronoi (sources/voronoi.cpp:19-80)
	// ------------------------------------------------------------------------------------------------------
	struct ivec3
	{
		int32_t v[3];
	};
	inline int32_t wadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
	ivec3 voronoiMix(const Voronoi &V, ivec3 s) // :19-24
	{
		const ivec3 a = { { s.v[1], s.v[2], s.v[0] } };
		ivec3 r;
		for (int i = 0; i < 3; i++)
			r.v[i] = wadd(wadd(wadd(V.seedHashes[i], s.v[i]), a.v[i]), (int32_t)murmurFinal((uint32_t)s.v[i]));
		return r;
	}
	struct VPoint
	{
		vec3 p;
		real d;
	};
	// all 27 * ppc candidate points in generation order (z, y, x outer to inner), optionally projected; sorted by distance
	void voronoiEvaluate(const Voronoi &V, vec3 position, const vec3 *normal, vec3 out[4]) // :44-80
	{
		std::vector<VPoint> pts;
		const vec3 pf = position * V.frequency;
		const ivec3 cell = { { (int32_t)pf[0], (int32_t)pf[1], (int32_t)pf[2] } };
		for (int32_t z = -1; z < 2; z++)
			for (int32_t y = -1; y < 2; y++)
				for (int32_t x = -1; x < 2; x++)
				{
					const ivec3 c = { { cell.v[0] + x, cell.v[1] + y, cell.v[2] + z } };
					ivec3 s = voronoiMix(V, c);
					for (uint32_t i = 0; i < V.pointsPerCell; i++)
					{
						s = voronoiMix(V, s);
						const vec3 g = V3((real)(s.v[0] % 65536), (real)(s.v[1] % 65536), (real)(s.v[2] % 65536)) / 65535; // :26
						VPoint vp;
						vp.p = (g + V3((real)c.v[0], (real)c.v[1], (real)c.v[2])) * V.cellSize;
						pts.push_back(vp);
					}
				}
		if (normal)
		{
			PlaneEq pl;
			pl.n = *normal;
			pl.d = -dot(position, *normal);
			for (VPoint &p : pts)
				p.p = closestPlane(p.p, pl);
		}
		for (VPoint &p : pts)
			p.d = dist2(p.p, position);
		std::partial_sort(pts.begin(), pts.begin() + 4, pts.end(), [](const VPoint &a, const VPoint &b) { return a.d < b.d; });
		for (int i = 0; i < 4; i++)
			out[i] = pts[i].p;
	}

	// ------------------------------------------------------------------------------------------------------
	// elevation modes (sources/elevationModes.cpp)
	// ------------------------------------------------------------------------------------------------------
	real mountains(const Planet &P, vec3 pos, real land) // :68-123
	{
		const real cover = 1 - rsat(land * (real)-0.1);
		if (cover < (real)1e-7)
			return land;
		const real mask = noise(P.en[1], pos);
		const real rm = rsmooth(rsat(mask * +7 - (real)0.3));
		const real tm = rsmooth(rsat(mask * -7 - (real)1.5));
		real ridge = noise(P.en[2], pos);
		ridge = rmax(ridge - (real)0.1, 0);
		ridge = std::pow(ridge, (real)1.6);
		ridge *= rm * cover;
		ridge *= 1000;
		real terraces = noise(P.en[3], pos);
		terraces = rmax(terraces + (real)0.1, 0) * (real)2.5;
		terraces = terrace(terraces, 4);
		terraces *= tm * cover;
		terraces *= 250;
		return land + smoothMax(0, rmax(ridge, terraces), 50);
	}
	real landMass(const Planet &P, vec3 pos, real exponent) // :143-149 / :165-171
	{
		real land = noise(P.en[0], pos) * (real)0.5 + (real)0.5;
		land = rsat(land);
		land = 1 - std::pow(land, exponent);
		land = land * 2 - 1;
		land = land / (rabs(land) + (real)0.17) + (real)0.15;
		land *= 150;
		return mountains(P, pos, land);
	}
	real elevation(const Planet &P, vec3 pos)
	{
		switch (P.elevation)
		{
			case E_NONE: // :9-12
				return 100;
			case E_SIMPLE: // :28-31
			{
				real a = noise(P.en[0], pos);
				a = -a + (real)0.3;
				a = std::pow(a * (real)1.3 - (real)0.35, (real)3) + (real)0.1;
				return 100 - a * 1000;
			}
			case E_LEGACY: // :56-64
			{
				const real scale = noise(P.en[0], pos) * (real)0.0005 + (real)0.0015;
				real a = noise(P.en[1], pos * scale);
				a += (real)0.11;
				if (a < 0)
					a = -std::pow(-a, (real)0.85);
				else
					a = std::pow(a, (real)1.7);
				return a * 2500;
			}
			case E_LAKES:
				return landMass(P, pos, (real)1.24);
			case E_ISLANDS:
				return landMass(P, pos, (real)0.83);
			case E_CRATERS: // :191-197
			{
				vec3 near4[4];
				voronoiEvaluate(P.craters, pos, nullptr, near4);
				const vec3 c = near4[0];
				const real s = 100 + noise(P.en[0], c) * 50;
				const real d = dist(pos, c);
				const real x = rsat(d / s);
				const real h = (std::pow(x, (real)4) - (real)0.9 * std::pow(x, (real)5) - (real)0.1) * 10;
				return 200 + h * 190;
			}
		}
		return NAN;
	}

	// sources/modesConfiguration.cpp:227-279, meshElevationRatio = 10
	enum
	{
		F_LAND = 0, F_WATER = 1, F_NAVIGATION = 2, F_ELEVATION_RAW = 3, F_ELEVATION = 4
	};
	real field(const Planet &P, int which, vec3 pos)
	{
		switch (which)
		{
			case F_WATER:
				return shape(P, pos);
			case F_ELEVATION:
				return shape(P, pos) * 10;
			case F_ELEVATION_RAW:
				return elevation(P, pos);
			case F_NAVIGATION:
			{
				const real base = shape(P, pos);
				const real elev = elevation(P, pos) / 10;
				return base - rmax(elev, 0);
			}
			default:
			{
				const real base = shape(P, pos);
				const real elev = elevation(P, pos) / 10;
				return base - elev;
			}
		}
	}

	// MarchingCubesCreateConfig::position (externals/cage/sources/libcore/mesh/marchingCubes.cpp:247-257)
	inline real gridCoord(real boxSize, int res, uint32_t i)
	{
		const real a = boxSize * (real)-0.5, b = boxSize * (real)0.5;
		const real f = ((real)i - 2) / ((real)res - 5);
		return (b - a) * f + a;
	}
}

extern "C"
{
	void *upo_planet_create(int shapeId, int elevationId, uint64_t s0, uint64_t s1)
	{
		if (shapeId < 0 || shapeId >= S_COUNT || elevationId < 0 || elevationId >= E_COUNT)
			return nullptr;
		Planet *P = new Planet();
		P->shape = shapeId;
		P->elevation = elevationId;
		P->thread.a = s0;
		P->thread.b = s1;
		P->seedsStarted = false;
		preseed(*P);
		return P;
	}
	void upo_planet_destroy(void *p)
	{
		delete (Planet *)p;
	}
	int upo_seed_log(void *p, uint32_t *out, int cap)
	{
		const Planet *P = (const Planet *)p;
		for (size_t i = 0; i < P->seedLog.size() && (int)i < cap; i++)
			out[i] = P->seedLog[i];
		return (int)P->seedLog.size();
	}
	void upo_points(void *p, int which, const float *pos, uint64_t n, float *out)
	{
		const Planet *P = (const Planet *)p;
		for (uint64_t i = 0; i < n; i++)
			out[i] = field(*P, which, V3(pos[i * 3 + 0], pos[i * 3 + 1], pos[i * 3 + 2]));
	}
	// MarchingCubes::updateByPosition (marchingCubes.cpp:149-165) over slices [z0, z1); threads > 1 only splits the z loop
	void upo_grid(void *p, int which, float boxSize, int res, int z0, int z1, float *out, int threads)
	{
		const Planet *P = (const Planet *)p;
		std::atomic<int> next{ z0 };
		auto work = [&]()
		{
			for (;;)
			{
				const int z = next++;
				if (z >= z1)
					break;
				float *o = out + (size_t)(z - z0) * res * res;
				for (int y = 0; y < res; y++)
					for (int x = 0; x < res; x++)
						*o++ = field(*P, which, V3(gridCoord(boxSize, res, x), gridCoord(boxSize, res, y), gridCoord(boxSize, res, z)));
			}
		};
		std::vector<std::thread> pool;
		for (int t = 1; t < threads; t++)
			pool.emplace_back(work);
		work();
		for (auto &t : pool)
			t.join();
	}
	void upo_grid_coords(float boxSize, int res, float *out)
	{
		for (int i = 0; i < res; i++)
			out[i] = gridCoord(boxSize, res, (uint32_t)i);
	}
	float upo_noise_eval(const upo_noise_cfg *c, float x, float y, float z)
	{
		return upo_noise3(c, x, y, z);
	}
	const float *upo_cellular_table(void)
	{
		return upo_randvecs3d();
	}
}
