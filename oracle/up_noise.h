/*
 * up_noise.h -- TEST INFRASTRUCTURE (oracle). Never included by the product path.
 *
 * CPU restatement of the coherent-noise arithmetic that the reference reaches through
 * cage's NoiseFunction (externals/cage/sources/libcore/noiseFunction.cpp:92-127), i.e. the
 * third-party header Auburn/FastNoiseLite `Cpp/FastNoiseLite.h`, pinned by the reference at
 * commit 785f37a9ad76e283586a379675085f2063ae03f7.  That submodule is ABSENT from
 * /root/reference, so this file restates the library's *published* algorithm:
 * integer hash lattice, Value / ValueCubic / Perlin / OpenSimplex2 / OpenSimplex2S /
 * Cellular single-octave generators, and the FBm / Ridged / PingPong fractal drivers.
 *
 * PARITY STATUS: "parity unpinned" for noise VALUES -- the reference's own noise test
 * (externals/cage/sources/test-core/noiseFunction.cpp:18-157) asserts nothing, and the
 * pinned source cannot be fetched.  Structural checks (range, continuity across lattice
 * cells, hash known-answers derived by hand) live in tests/test_oracle_port.py; tools/fastnoise_pin.py
 * diff-reports this file against the real header generator by generator the day it is available (make -C oracle ref FASTNOISE=...).
 * The Cellular generator needs upstream's 256-entry RandVecs3D constant table, which cannot
 * be reconstructed; `upo_randvecs3d()` below is a DECLARED SUBSTITUTE (deterministic unit
 * vectors from a splitmix64 stream).  Every Cellular-dependent result is therefore
 * "GPU == oracle, oracle != upstream bit-for-bit".
 *
 * Plain C99-compatible; all arithmetic is IEEE binary32, compile with -ffp-contract=off.
 */
#ifndef UP_ORACLE_NOISE_H
#define UP_ORACLE_NOISE_H

#include <math.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum values follow cage's NoiseTypeEnum order (include/cage-core/noiseFunction.h:8-43) */
enum { UPO_NOISE_SIMPLEX = 0, UPO_NOISE_SIMPLEX_REDUCED = 1, UPO_NOISE_CELLULAR = 2, UPO_NOISE_PERLIN = 3, UPO_NOISE_CUBIC = 4, UPO_NOISE_VALUE = 5 };
enum { UPO_FRACTAL_NONE = 0, UPO_FRACTAL_FBM = 1, UPO_FRACTAL_RIDGED = 2, UPO_FRACTAL_PINGPONG = 3 };
enum { UPO_DIST_EUCLIDEAN = 0, UPO_DIST_EUCLIDEAN_SQ = 1, UPO_DIST_MANHATTAN = 2, UPO_DIST_HYBRID = 3 };
enum { UPO_OP_CELL = 0, UPO_OP_DISTANCE = 1, UPO_OP_DISTANCE2 = 2, UPO_OP_ADD = 3, UPO_OP_SUBTRACT = 4, UPO_OP_MULTIPLY = 5, UPO_OP_DIVIDE = 6 };

typedef struct upo_noise_cfg
{
	int32_t type; /* UPO_NOISE_* */
	int32_t fractal; /* UPO_FRACTAL_* */
	int32_t distance; /* UPO_DIST_* */
	int32_t operation; /* UPO_OP_* */
	int32_t seed;
	int32_t octaves;
	float lacunarity;
	float gain;
	float frequency;
	float bounding; /* derived: 1 / sum_{i<octaves} |gain|^i */
} upo_noise_cfg;

/* cage defaults: include/cage-core/noiseFunction.h:59-72 */
static inline upo_noise_cfg upo_noise_default(uint32_t seed)
{
	upo_noise_cfg c;
	c.type = UPO_NOISE_SIMPLEX;
	c.fractal = UPO_FRACTAL_NONE;
	c.distance = UPO_DIST_EUCLIDEAN;
	c.operation = UPO_OP_CELL;
	c.seed = (int32_t)seed;
	c.octaves = 3;
	c.lacunarity = 2.0f;
	c.gain = 0.5f;
	c.frequency = 1.0f;
	c.bounding = 0.0f;
	return c;
}

/* fractal bounding: amplitude normaliser, uses |gain| */
static inline float upo_fractal_bounding(int32_t octaves, float gain)
{
	float g = fabsf(gain);
	float amp = g;
	float total = 1.0f;
	for (int32_t i = 1; i < octaves; i++)
	{
		total += amp;
		amp *= g;
	}
	return 1.0f / total;
}

static inline void upo_noise_finish(upo_noise_cfg *c)
{
	c->bounding = upo_fractal_bounding(c->octaves, c->gain);
}

/* ---- integer lattice -------------------------------------------------------------------- */

#define UPO_PX 501125321
#define UPO_PY 1136930381
#define UPO_PZ 1720413743

/* all integer arithmetic is two's-complement wrap-around on 32 bits */
static inline int32_t upo_imul(int32_t a, int32_t b)
{
	return (int32_t)((uint32_t)a * (uint32_t)b);
}
static inline int32_t upo_iadd(int32_t a, int32_t b)
{
	return (int32_t)((uint32_t)a + (uint32_t)b);
}
static inline int32_t upo_isub(int32_t a, int32_t b)
{
	return (int32_t)((uint32_t)a - (uint32_t)b);
}

static inline int32_t upo_hash3(int32_t seed, int32_t xp, int32_t yp, int32_t zp)
{
	return upo_imul(seed ^ xp ^ yp ^ zp, 0x27d4eb2d);
}
static inline int32_t upo_hash2(int32_t seed, int32_t xp, int32_t yp)
{
	return upo_imul(seed ^ xp ^ yp, 0x27d4eb2d);
}

/* lattice value in [-1, 1): square the hash, fold the low bits up, scale by 2^-31 */
static inline int32_t upo_valbits(int32_t h)
{
	h = upo_imul(h, h);
	h ^= (int32_t)((uint32_t)h << 19);
	return h;
}
static inline float upo_val3(int32_t seed, int32_t xp, int32_t yp, int32_t zp)
{
	return (float)upo_valbits(upo_hash3(seed, xp, yp, zp)) * (1.0f / 2147483648.0f);
}
static inline float upo_val2(int32_t seed, int32_t xp, int32_t yp)
{
	return (float)upo_valbits(upo_hash2(seed, xp, yp)) * (1.0f / 2147483648.0f);
}

/* gradient index: 64 directions, 4 floats per entry -> byte-less index h & (63<<2) */
static inline int32_t upo_gradindex3(int32_t h)
{
	h ^= h >> 15; /* arithmetic shift */
	return h & (63 << 2);
}

/* the 64-entry gradient set: the 12 cube-edge directions, five times over, then 4 fillers */
static inline void upo_grad3(int32_t idx4, float *gx, float *gy, float *gz)
{
	static const signed char edges[12][3] = {
		{ 0, 1, 1 }, { 0, -1, 1 }, { 0, 1, -1 }, { 0, -1, -1 },
		{ 1, 0, 1 }, { -1, 0, 1 }, { 1, 0, -1 }, { -1, 0, -1 },
		{ 1, 1, 0 }, { -1, 1, 0 }, { 1, -1, 0 }, { -1, -1, 0 },
	};
	static const signed char fill[4][3] = { { 1, 1, 0 }, { 0, -1, 1 }, { -1, 1, 0 }, { 0, -1, -1 } };
	const int32_t e = idx4 >> 2;
	const signed char *g = e < 60 ? edges[e % 12] : fill[e - 60];
	*gx = (float)g[0];
	*gy = (float)g[1];
	*gz = (float)g[2];
}

static inline float upo_gradcoord3(int32_t seed, int32_t xp, int32_t yp, int32_t zp, float xd, float yd, float zd)
{
#ifdef UPO_WORST_CASE_GRADIENTS
	/* tools/noise_bounds.cpp: the largest value ANY of the gradients (two non-zero components of magnitude 1) can give at this offset,
	 * i.e. the sum of the two largest |components|: turns the Perlin / OpenSimplex2 generators into tight upper bounds of themselves */
	(void)seed, (void)xp, (void)yp, (void)zp;
	const float ax = xd < 0 ? -xd : xd, ay = yd < 0 ? -yd : yd, az = zd < 0 ? -zd : zd;
	const float mn = ax < ay ? (ax < az ? ax : az) : (ay < az ? ay : az);
	return (ax + ay + az) - mn;
#else
	float gx, gy, gz;
	upo_grad3(upo_gradindex3(upo_hash3(seed, xp, yp, zp)), &gx, &gy, &gz);
	return xd * gx + yd * gy + zd * gz;
#endif
}

static inline int32_t upo_fastfloor(float f)
{
	return f >= 0 ? (int32_t)f : (int32_t)f - 1;
}
static inline int32_t upo_fastround(float f)
{
	return f >= 0 ? (int32_t)(f + 0.5f) : (int32_t)(f - 0.5f);
}
static inline float upo_lerp(float a, float b, float t)
{
	return a + t * (b - a);
}
static inline float upo_hermite(float t)
{
	return t * t * (3 - 2 * t);
}
static inline float upo_quintic(float t)
{
	return t * t * t * (t * (t * 6 - 15) + 10);
}
static inline float upo_cubiclerp(float a, float b, float c, float d, float t)
{
	const float p = (d - c) - (a - b);
	return t * t * t * p + t * t * ((a - b) - p) + t * (c - a) + b;
}

/* ---- single-octave generators ----------------------------------------------------------- */

static inline float upo_single_value3(int32_t seed, float x, float y, float z)
{
	int32_t x0 = upo_fastfloor(x), y0 = upo_fastfloor(y), z0 = upo_fastfloor(z);
	const float xs = upo_hermite((float)(x - (float)x0));
	const float ys = upo_hermite((float)(y - (float)y0));
	const float zs = upo_hermite((float)(z - (float)z0));
	x0 = upo_imul(x0, UPO_PX);
	y0 = upo_imul(y0, UPO_PY);
	z0 = upo_imul(z0, UPO_PZ);
	const int32_t x1 = upo_iadd(x0, UPO_PX), y1 = upo_iadd(y0, UPO_PY), z1 = upo_iadd(z0, UPO_PZ);
	const float xf00 = upo_lerp(upo_val3(seed, x0, y0, z0), upo_val3(seed, x1, y0, z0), xs);
	const float xf10 = upo_lerp(upo_val3(seed, x0, y1, z0), upo_val3(seed, x1, y1, z0), xs);
	const float xf01 = upo_lerp(upo_val3(seed, x0, y0, z1), upo_val3(seed, x1, y0, z1), xs);
	const float xf11 = upo_lerp(upo_val3(seed, x0, y1, z1), upo_val3(seed, x1, y1, z1), xs);
	const float yf0 = upo_lerp(xf00, xf10, ys);
	const float yf1 = upo_lerp(xf01, xf11, ys);
	return upo_lerp(yf0, yf1, zs);
}

static inline float upo_single_value2(int32_t seed, float x, float y)
{
	int32_t x0 = upo_fastfloor(x), y0 = upo_fastfloor(y);
	const float xs = upo_hermite((float)(x - (float)x0));
	const float ys = upo_hermite((float)(y - (float)y0));
	x0 = upo_imul(x0, UPO_PX);
	y0 = upo_imul(y0, UPO_PY);
	const int32_t x1 = upo_iadd(x0, UPO_PX), y1 = upo_iadd(y0, UPO_PY);
	const float xf0 = upo_lerp(upo_val2(seed, x0, y0), upo_val2(seed, x1, y0), xs);
	const float xf1 = upo_lerp(upo_val2(seed, x0, y1), upo_val2(seed, x1, y1), xs);
	return upo_lerp(xf0, xf1, ys);
}

static inline float upo_single_perlin3(int32_t seed, float x, float y, float z)
{
	int32_t x0 = upo_fastfloor(x), y0 = upo_fastfloor(y), z0 = upo_fastfloor(z);
	const float xd0 = (float)(x - (float)x0), yd0 = (float)(y - (float)y0), zd0 = (float)(z - (float)z0);
	const float xd1 = xd0 - 1, yd1 = yd0 - 1, zd1 = zd0 - 1;
	const float xs = upo_quintic(xd0), ys = upo_quintic(yd0), zs = upo_quintic(zd0);
	x0 = upo_imul(x0, UPO_PX);
	y0 = upo_imul(y0, UPO_PY);
	z0 = upo_imul(z0, UPO_PZ);
	const int32_t x1 = upo_iadd(x0, UPO_PX), y1 = upo_iadd(y0, UPO_PY), z1 = upo_iadd(z0, UPO_PZ);
	const float xf00 = upo_lerp(upo_gradcoord3(seed, x0, y0, z0, xd0, yd0, zd0), upo_gradcoord3(seed, x1, y0, z0, xd1, yd0, zd0), xs);
	const float xf10 = upo_lerp(upo_gradcoord3(seed, x0, y1, z0, xd0, yd1, zd0), upo_gradcoord3(seed, x1, y1, z0, xd1, yd1, zd0), xs);
	const float xf01 = upo_lerp(upo_gradcoord3(seed, x0, y0, z1, xd0, yd0, zd1), upo_gradcoord3(seed, x1, y0, z1, xd1, yd0, zd1), xs);
	const float xf11 = upo_lerp(upo_gradcoord3(seed, x0, y1, z1, xd0, yd1, zd1), upo_gradcoord3(seed, x1, y1, z1, xd1, yd1, zd1), xs);
	const float yf0 = upo_lerp(xf00, xf10, ys);
	const float yf1 = upo_lerp(xf01, xf11, ys);
	return upo_lerp(yf0, yf1, zs) * 0.964921414852142333984375f;
}

static inline float upo_single_cubic3(int32_t seed, float x, float y, float z)
{
	int32_t x1 = upo_fastfloor(x), y1 = upo_fastfloor(y), z1 = upo_fastfloor(z);
	const float xs = (float)(x - (float)x1), ys = (float)(y - (float)y1), zs = (float)(z - (float)z1);
	x1 = upo_imul(x1, UPO_PX);
	y1 = upo_imul(y1, UPO_PY);
	z1 = upo_imul(z1, UPO_PZ);
	int32_t xi[4], yi[4], zi[4];
	xi[0] = upo_isub(x1, UPO_PX);
	xi[1] = x1;
	xi[2] = upo_iadd(x1, UPO_PX);
	xi[3] = upo_iadd(x1, (int32_t)((uint32_t)UPO_PX << 1));
	yi[0] = upo_isub(y1, UPO_PY);
	yi[1] = y1;
	yi[2] = upo_iadd(y1, UPO_PY);
	yi[3] = upo_iadd(y1, (int32_t)((uint32_t)UPO_PY << 1));
	zi[0] = upo_isub(z1, UPO_PZ);
	zi[1] = z1;
	zi[2] = upo_iadd(z1, UPO_PZ);
	zi[3] = upo_iadd(z1, (int32_t)((uint32_t)UPO_PZ << 1));
	float zr[4];
	for (int k = 0; k < 4; k++)
	{
		float yr[4];
		for (int j = 0; j < 4; j++)
			yr[j] = upo_cubiclerp(upo_val3(seed, xi[0], yi[j], zi[k]), upo_val3(seed, xi[1], yi[j], zi[k]), upo_val3(seed, xi[2], yi[j], zi[k]), upo_val3(seed, xi[3], yi[j], zi[k]), xs);
		zr[k] = upo_cubiclerp(yr[0], yr[1], yr[2], yr[3], ys);
	}
	return upo_cubiclerp(zr[0], zr[1], zr[2], zr[3], zs) * (1 / (1.5f * 1.5f * 1.5f));
}

/* OpenSimplex2, 3-D: two offset cubic lattices, up to two contributions from each */
static inline float upo_single_simplex3(int32_t seed, float x, float y, float z)
{
	int32_t i = upo_fastround(x), j = upo_fastround(y), k = upo_fastround(z);
	float x0 = (float)(x - (float)i), y0 = (float)(y - (float)j), z0 = (float)(z - (float)k);

	int32_t xNSign = (int32_t)(-1.0f - x0) | 1;
	int32_t yNSign = (int32_t)(-1.0f - y0) | 1;
	int32_t zNSign = (int32_t)(-1.0f - z0) | 1;

	float ax0 = (float)xNSign * -x0;
	float ay0 = (float)yNSign * -y0;
	float az0 = (float)zNSign * -z0;

	i = upo_imul(i, UPO_PX);
	j = upo_imul(j, UPO_PY);
	k = upo_imul(k, UPO_PZ);

	float value = 0;
	float a = (0.6f - x0 * x0) - (y0 * y0 + z0 * z0);

	for (int l = 0;; l++)
	{
		if (a > 0)
			value += (a * a) * (a * a) * upo_gradcoord3(seed, i, j, k, x0, y0, z0);

		if (ax0 >= ay0 && ax0 >= az0)
		{
			float b = a + ax0 + ax0;
			if (b > 1)
			{
				b -= 1;
				value += (b * b) * (b * b) * upo_gradcoord3(seed, upo_isub(i, upo_imul(xNSign, UPO_PX)), j, k, x0 + (float)xNSign, y0, z0);
			}
		}
		else if (ay0 > ax0 && ay0 >= az0)
		{
			float b = a + ay0 + ay0;
			if (b > 1)
			{
				b -= 1;
				value += (b * b) * (b * b) * upo_gradcoord3(seed, i, upo_isub(j, upo_imul(yNSign, UPO_PY)), k, x0, y0 + (float)yNSign, z0);
			}
		}
		else
		{
			float b = a + az0 + az0;
			if (b > 1)
			{
				b -= 1;
				value += (b * b) * (b * b) * upo_gradcoord3(seed, i, j, upo_isub(k, upo_imul(zNSign, UPO_PZ)), x0, y0, z0 + (float)zNSign);
			}
		}

		if (l == 1)
			break;

		ax0 = 0.5f - ax0;
		ay0 = 0.5f - ay0;
		az0 = 0.5f - az0;

		x0 = (float)xNSign * ax0;
		y0 = (float)yNSign * ay0;
		z0 = (float)zNSign * az0;

		a += (0.75f - ax0) - (ay0 + az0);

		i = upo_iadd(i, (xNSign >> 1) & UPO_PX);
		j = upo_iadd(j, (yNSign >> 1) & UPO_PY);
		k = upo_iadd(k, (zNSign >> 1) & UPO_PZ);

		xNSign = -xNSign;
		yNSign = -yNSign;
		zNSign = -zNSign;

		seed = ~seed;
	}
	return value * 32.69428253173828125f;
}

/* OpenSimplex2S ("smooth"), 3-D: floor-based, kernel 0.75 - r^2, 2 fixed + up to 8 optional terms */
static inline float upo_s2s_term(float a, int32_t seed, int32_t xp, int32_t yp, int32_t zp, float dx, float dy, float dz)
{
	return (a * a) * (a * a) * upo_gradcoord3(seed, xp, yp, zp, dx, dy, dz);
}

static inline float upo_single_simplexs3(int32_t seed, float x, float y, float z)
{
	int32_t i = upo_fastfloor(x), j = upo_fastfloor(y), k = upo_fastfloor(z);
	const float xi = (float)(x - (float)i), yi = (float)(y - (float)j), zi = (float)(z - (float)k);

	i = upo_imul(i, UPO_PX);
	j = upo_imul(j, UPO_PY);
	k = upo_imul(k, UPO_PZ);
	const int32_t seed2 = upo_iadd(seed, 1293373);

	const int32_t xNMask = (int32_t)(-0.5f - xi);
	const int32_t yNMask = (int32_t)(-0.5f - yi);
	const int32_t zNMask = (int32_t)(-0.5f - zi);

	const float x0 = xi + (float)xNMask;
	const float y0 = yi + (float)yNMask;
	const float z0 = zi + (float)zNMask;
	const float a0 = 0.75f - x0 * x0 - y0 * y0 - z0 * z0;
	float value = upo_s2s_term(a0, seed, upo_iadd(i, xNMask & UPO_PX), upo_iadd(j, yNMask & UPO_PY), upo_iadd(k, zNMask & UPO_PZ), x0, y0, z0);

	const float x1 = xi - 0.5f;
	const float y1 = yi - 0.5f;
	const float z1 = zi - 0.5f;
	const float a1 = 0.75f - x1 * x1 - y1 * y1 - z1 * z1;
	value += upo_s2s_term(a1, seed2, upo_iadd(i, UPO_PX), upo_iadd(j, UPO_PY), upo_iadd(k, UPO_PZ), x1, y1, z1);

	const float xAFlipMask0 = (float)((xNMask | 1) << 1) * x1;
	const float yAFlipMask0 = (float)((yNMask | 1) << 1) * y1;
	const float zAFlipMask0 = (float)((zNMask | 1) << 1) * z1;
	const float xAFlipMask1 = (float)(-2 - (xNMask << 2)) * x1 - 1.0f;
	const float yAFlipMask1 = (float)(-2 - (yNMask << 2)) * y1 - 1.0f;
	const float zAFlipMask1 = (float)(-2 - (zNMask << 2)) * z1 - 1.0f;

	const int32_t PX2 = (int32_t)((uint32_t)UPO_PX << 1);
	const int32_t PY2 = (int32_t)((uint32_t)UPO_PY << 1);
	const int32_t PZ2 = (int32_t)((uint32_t)UPO_PZ << 1);

	int skip5 = 0;
	const float a2 = xAFlipMask0 + a0;
	if (a2 > 0)
	{
		value += upo_s2s_term(a2, seed, upo_iadd(i, ~xNMask & UPO_PX), upo_iadd(j, yNMask & UPO_PY), upo_iadd(k, zNMask & UPO_PZ), x0 - (float)(xNMask | 1), y0, z0);
	}
	else
	{
		const float a3 = yAFlipMask0 + zAFlipMask0 + a0;
		if (a3 > 0)
			value += upo_s2s_term(a3, seed, upo_iadd(i, xNMask & UPO_PX), upo_iadd(j, ~yNMask & UPO_PY), upo_iadd(k, ~zNMask & UPO_PZ), x0, y0 - (float)(yNMask | 1), z0 - (float)(zNMask | 1));
		const float a4 = xAFlipMask1 + a1;
		if (a4 > 0)
		{
			value += upo_s2s_term(a4, seed2, upo_iadd(i, xNMask & PX2), upo_iadd(j, UPO_PY), upo_iadd(k, UPO_PZ), (float)(xNMask | 1) + x1, y1, z1);
			skip5 = 1;
		}
	}

	int skip9 = 0;
	const float a6 = yAFlipMask0 + a0;
	if (a6 > 0)
	{
		value += upo_s2s_term(a6, seed, upo_iadd(i, xNMask & UPO_PX), upo_iadd(j, ~yNMask & UPO_PY), upo_iadd(k, zNMask & UPO_PZ), x0, y0 - (float)(yNMask | 1), z0);
	}
	else
	{
		const float a7 = xAFlipMask0 + zAFlipMask0 + a0;
		if (a7 > 0)
			value += upo_s2s_term(a7, seed, upo_iadd(i, ~xNMask & UPO_PX), upo_iadd(j, yNMask & UPO_PY), upo_iadd(k, ~zNMask & UPO_PZ), x0 - (float)(xNMask | 1), y0, z0 - (float)(zNMask | 1));
		const float a8 = yAFlipMask1 + a1;
		if (a8 > 0)
		{
			value += upo_s2s_term(a8, seed2, upo_iadd(i, UPO_PX), upo_iadd(j, yNMask & PY2), upo_iadd(k, UPO_PZ), x1, (float)(yNMask | 1) + y1, z1);
			skip9 = 1;
		}
	}

	int skipD = 0;
	const float aA = zAFlipMask0 + a0;
	if (aA > 0)
	{
		value += upo_s2s_term(aA, seed, upo_iadd(i, xNMask & UPO_PX), upo_iadd(j, yNMask & UPO_PY), upo_iadd(k, ~zNMask & UPO_PZ), x0, y0, z0 - (float)(zNMask | 1));
	}
	else
	{
		const float aB = xAFlipMask0 + yAFlipMask0 + a0;
		if (aB > 0)
			value += upo_s2s_term(aB, seed, upo_iadd(i, ~xNMask & UPO_PX), upo_iadd(j, ~yNMask & UPO_PY), upo_iadd(k, zNMask & UPO_PZ), x0 - (float)(xNMask | 1), y0 - (float)(yNMask | 1), z0);
		const float aC = zAFlipMask1 + a1;
		if (aC > 0)
		{
			value += upo_s2s_term(aC, seed2, upo_iadd(i, UPO_PX), upo_iadd(j, UPO_PY), upo_iadd(k, zNMask & PZ2), x1, y1, (float)(zNMask | 1) + z1);
			skipD = 1;
		}
	}

	if (!skip5)
	{
		const float a5 = yAFlipMask1 + zAFlipMask1 + a1;
		if (a5 > 0)
			value += upo_s2s_term(a5, seed2, upo_iadd(i, UPO_PX), upo_iadd(j, yNMask & PY2), upo_iadd(k, zNMask & PZ2), x1, (float)(yNMask | 1) + y1, (float)(zNMask | 1) + z1);
	}
	if (!skip9)
	{
		const float a9 = xAFlipMask1 + zAFlipMask1 + a1;
		if (a9 > 0)
			value += upo_s2s_term(a9, seed2, upo_iadd(i, xNMask & PX2), upo_iadd(j, UPO_PY), upo_iadd(k, zNMask & PZ2), (float)(xNMask | 1) + x1, y1, (float)(zNMask | 1) + z1);
	}
	if (!skipD)
	{
		const float aD = xAFlipMask1 + yAFlipMask1 + a1;
		if (aD > 0)
			value += upo_s2s_term(aD, seed2, upo_iadd(i, xNMask & PX2), upo_iadd(j, yNMask & PY2), upo_iadd(k, UPO_PZ), (float)(xNMask | 1) + x1, (float)(yNMask | 1) + y1, z1);
	}

	return value * 9.046026385208288f;
}

/*
 * DECLARED SUBSTITUTE for upstream's RandVecs3D (256 unit vectors).  Entry e is generated from
 * a splitmix64 stream seeded with 0x5eedc0de00000000 + e: z uniform in [-1,1], azimuth uniform,
 * which is uniform on the sphere.  Computed in double, rounded once to float.  The product side
 * receives this table as DATA (uploaded once); it never calls this function.
 */
static inline uint64_t upo_splitmix64(uint64_t *s)
{
	uint64_t z = (*s += 0x9e3779b97f4a7c15ull);
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
	return z ^ (z >> 31);
}

static inline const float *upo_randvecs3d(void)
{
	static float table[256 * 4];
	static int ready = 0;
	if (!ready)
	{
		for (int e = 0; e < 256; e++)
		{
			uint64_t s = 0x5eedc0de00000000ull + (uint64_t)e;
			const double u = (double)(upo_splitmix64(&s) >> 11) * (1.0 / 9007199254740992.0);
			const double v = (double)(upo_splitmix64(&s) >> 11) * (1.0 / 9007199254740992.0);
			const double zz = 2.0 * u - 1.0;
			const double rr = sqrt(1.0 - zz * zz);
			const double ph = 6.283185307179586476925286766559 * v;
			table[e * 4 + 0] = (float)(rr * cos(ph));
			table[e * 4 + 1] = (float)(rr * sin(ph));
			table[e * 4 + 2] = (float)zz;
			table[e * 4 + 3] = 0.0f;
		}
		ready = 1;
	}
	return table;
}

static inline float upo_single_cellular3(const upo_noise_cfg *c, int32_t seed, float x, float y, float z)
{
	const float *rv = upo_randvecs3d();
	const int32_t xr = upo_fastround(x), yr = upo_fastround(y), zr = upo_fastround(z);
	float d0 = 1e10f, d1 = 1e10f;
	int32_t closest = 0;
	const float jitter = 0.39614353f * 1.0f; /* default cellular jitter modifier */

	int32_t xp = upo_imul(xr - 1, UPO_PX);
	const int32_t ypBase = upo_imul(yr - 1, UPO_PY);
	const int32_t zpBase = upo_imul(zr - 1, UPO_PZ);
	for (int32_t xi = xr - 1; xi <= xr + 1; xi++)
	{
		int32_t yp = ypBase;
		for (int32_t yi = yr - 1; yi <= yr + 1; yi++)
		{
			int32_t zp = zpBase;
			for (int32_t zi = zr - 1; zi <= zr + 1; zi++)
			{
				const int32_t h = upo_hash3(seed, xp, yp, zp);
				const int32_t idx = h & (255 << 2);
				const float vx = (float)((float)xi - x) + rv[idx] * jitter;
				const float vy = (float)((float)yi - y) + rv[idx | 1] * jitter;
				const float vz = (float)((float)zi - z) + rv[idx | 2] * jitter;
				float nd;
				switch (c->distance)
				{
					case UPO_DIST_MANHATTAN:
						nd = fabsf(vx) + fabsf(vy) + fabsf(vz);
						break;
					case UPO_DIST_HYBRID:
						nd = (fabsf(vx) + fabsf(vy) + fabsf(vz)) + (vx * vx + vy * vy + vz * vz);
						break;
					default:
						nd = vx * vx + vy * vy + vz * vz;
						break;
				}
				{
					const float mn = d1 < nd ? d1 : nd;
					d1 = mn > d0 ? mn : d0;
				}
				if (nd < d0)
				{
					d0 = nd;
					closest = h;
				}
				zp = upo_iadd(zp, UPO_PZ);
			}
			yp = upo_iadd(yp, UPO_PY);
		}
		xp = upo_iadd(xp, UPO_PX);
	}

	if (c->distance == UPO_DIST_EUCLIDEAN && c->operation >= UPO_OP_DISTANCE)
	{
		d0 = sqrtf(d0);
		if (c->operation >= UPO_OP_DISTANCE2)
			d1 = sqrtf(d1);
	}

	switch (c->operation)
	{
		case UPO_OP_CELL:
			return (float)closest * (1 / 2147483648.0f);
		case UPO_OP_DISTANCE:
			return d0 - 1;
		case UPO_OP_DISTANCE2:
			return d1 - 1;
		case UPO_OP_ADD:
			return (d1 + d0) * 0.5f - 1;
		case UPO_OP_SUBTRACT:
			return d1 - d0 - 1;
		case UPO_OP_MULTIPLY:
			return d1 * d0 * 0.5f - 1;
		case UPO_OP_DIVIDE:
			return d0 / d1 - 1;
		default:
			return 0;
	}
}

static inline float upo_single3(const upo_noise_cfg *c, int32_t seed, float x, float y, float z)
{
	switch (c->type)
	{
		case UPO_NOISE_SIMPLEX:
			return upo_single_simplex3(seed, x, y, z);
		case UPO_NOISE_SIMPLEX_REDUCED:
			return upo_single_simplexs3(seed, x, y, z);
		case UPO_NOISE_CELLULAR:
			return upo_single_cellular3(c, seed, x, y, z);
		case UPO_NOISE_PERLIN:
			return upo_single_perlin3(seed, x, y, z);
		case UPO_NOISE_CUBIC:
			return upo_single_cubic3(seed, x, y, z);
		case UPO_NOISE_VALUE:
			return upo_single_value3(seed, x, y, z);
		default:
			return 0;
	}
}

static inline float upo_pingpong(float t)
{
	t -= (float)((int32_t)(t * 0.5f) * 2);
	return t < 1 ? t : 2 - t;
}

/* GetNoise(x, y, z) as cage calls it: frequency scale, domain rotation for the simplex types, fractal driver */
static inline float upo_noise3(const upo_noise_cfg *c, float x, float y, float z)
{
	x *= c->frequency;
	y *= c->frequency;
	z *= c->frequency;
	if (c->type == UPO_NOISE_SIMPLEX || c->type == UPO_NOISE_SIMPLEX_REDUCED)
	{
		const float R3 = (float)(2.0 / 3.0);
		const float r = (x + y + z) * R3;
		x = r - x;
		y = r - y;
		z = r - z;
	}
	if (c->fractal == UPO_FRACTAL_NONE)
		return upo_single3(c, c->seed, x, y, z);

	int32_t seed = c->seed;
	float sum = 0;
	float amp = c->bounding;
	for (int32_t o = 0; o < c->octaves; o++)
	{
		float n = upo_single3(c, seed, x, y, z);
		seed = upo_iadd(seed, 1);
		switch (c->fractal)
		{
			case UPO_FRACTAL_FBM:
				sum += n * amp;
				/* weighted strength is 0 (library default, cage never sets it): amp *= lerp(1, ., 0) == amp */
				break;
			case UPO_FRACTAL_RIDGED:
				n = fabsf(n);
				sum += (n * -2 + 1) * amp;
				break;
			default: /* ping-pong, strength 2 (library default) */
				n = upo_pingpong((n + 1) * 2.0f);
				sum += (n - 0.5f) * 2 * amp;
				break;
		}
		x *= c->lacunarity;
		y *= c->lacunarity;
		z *= c->lacunarity;
		amp *= c->gain;
	}
	return sum;
}

/* GetNoise(x, y): only the Value generator is reachable from the hot path (coloringModes.cpp:48, evaluate(Real) -> (x, 0)) */
static inline int upo_noise2_supported(const upo_noise_cfg *c)
{
	return c->type == UPO_NOISE_VALUE;
}

static inline float upo_noise2(const upo_noise_cfg *c, float x, float y)
{
	x *= c->frequency;
	y *= c->frequency;
	if (c->fractal == UPO_FRACTAL_NONE)
		return upo_single_value2(c->seed, x, y);
	int32_t seed = c->seed;
	float sum = 0;
	float amp = c->bounding;
	for (int32_t o = 0; o < c->octaves; o++)
	{
		float n = upo_single_value2(seed, x, y);
		seed = upo_iadd(seed, 1);
		switch (c->fractal)
		{
			case UPO_FRACTAL_FBM:
				sum += n * amp;
				break;
			case UPO_FRACTAL_RIDGED:
				n = fabsf(n);
				sum += (n * -2 + 1) * amp;
				break;
			default:
				n = upo_pingpong((n + 1) * 2.0f);
				sum += (n - 0.5f) * 2 * amp;
				break;
		}
		x *= c->lacunarity;
		y *= c->lacunarity;
		amp *= c->gain;
	}
	return sum;
}

#ifdef __cplusplus
}
#endif

#endif /* UP_ORACLE_NOISE_H */
