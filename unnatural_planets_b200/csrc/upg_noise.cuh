// upg_noise.cuh -- GPU coherent-noise library: integer hash lattice, Value / ValueCubic / Perlin /
// OpenSimplex2 / OpenSimplex2S / Cellular generators and the FBm / Ridged / PingPong fractal drivers,
// FP32 with register-resident octave loops.
//
// Replaces, for the hot path, cage's NoiseFunction::evaluate (externals/cage/sources/libcore/
// noiseFunction.cpp:111-127) and what it forwards to: FastNoiseLite::GetNoise of the un-vendored
// submodule Auburn/FastNoiseLite @785f37a9 (configuration mapping: noiseFunction.cpp:14-107).
// Integer work (primes, hash, value bits, gradient index) must be bit-exact; float work follows
// the reference's operation order with no FMA contraction (see upg_common.cuh).
//
// Noise kind / fractal kind / octave count are template parameters: each call site in the
// reference fixes them at compile time (e.g. sources/elevationModes.cpp:131-140), only the seed
// (and, for the asteroid shape, the frequency) are per-planet data.
#pragma once

#include "upg_common.cuh"

// The octave loop of the fractal drivers is kept ROLLED by default: the bodies are hundreds of instructions each and the
// unrolled kernels (130 KB of SASS for sphere+lakes) overflow the 32 KB L1.5 instruction cache, which shows up as
// "no_instruction" issue stalls (profiles/).  -DUPG_UNROLL_OCTAVES restores full unrolling for A/B measurements.
#if defined(UPG_UNROLL_OCTAVES)
#define UPG_OCTAVE_LOOP _Pragma("unroll")
#else
#define UPG_OCTAVE_LOOP _Pragma("unroll 1")
#endif

namespace upg
{
	// values follow cage's enums (include/cage-core/noiseFunction.h:8-43)
	enum NoiseType : int
	{
		NT_SIMPLEX = 0, // OpenSimplex2
		NT_SIMPLEX_REDUCED = 1, // OpenSimplex2S
		NT_CELLULAR = 2,
		NT_PERLIN = 3,
		NT_CUBIC = 4, // ValueCubic
		NT_VALUE = 5,
	};
	enum FractalType : int
	{
		FT_NONE = 0,
		FT_FBM = 1,
		FT_RIDGED = 2,
		FT_PINGPONG = 3,
	};
	enum CellDistance : int
	{
		CD_EUCLIDEAN = 0,
		CD_EUCLIDEAN_SQ = 1,
		CD_MANHATTAN = 2,
		CD_HYBRID = 3,
	};
	enum CellOperation : int
	{
		CO_CELL = 0,
		CO_DISTANCE = 1,
		CO_DISTANCE2 = 2,
		CO_ADD = 3,
		CO_SUBTRACT = 4,
		CO_MULTIPLY = 5,
		CO_DIVIDE = 6,
	};

	// per-planet data of one noise instance (the compile-time part lives in the template arguments)
	struct NoiseInst
	{
		int32_t seed;
		float frequency;
		float lacunarity;
		float gain;
		float bounding; // 1 / sum |gain|^i, computed on the host with the same float loop
	};

	constexpr int32_t PRIME_X = 501125321;
	constexpr int32_t PRIME_Y = 1136930381;
	constexpr int32_t PRIME_Z = 1720413743;

	UPG_HD int32_t imul(int32_t a, int32_t b) { return (int32_t)((uint32_t)a * (uint32_t)b); }
	UPG_HD int32_t iadd(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
	UPG_HD int32_t isub(int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); }
	UPG_HD int32_t ishl(int32_t a, int s) { return (int32_t)((uint32_t)a << s); }

	UPG_HD int32_t lattice_hash(int32_t seed, int32_t xp, int32_t yp, int32_t zp) { return imul(seed ^ xp ^ yp ^ zp, 0x27d4eb2d); }
	UPG_HD int32_t lattice_hash(int32_t seed, int32_t xp, int32_t yp) { return imul(seed ^ xp ^ yp, 0x27d4eb2d); }

	// Lattice value BEFORE the library's final scale by 2^-31.  The scale is a power of two, so it commutes exactly with every
	// rounding of the interpolation that follows (no intermediate comes near the subnormal range: the unscaled corner values
	// are integers): the generators below interpolate the unscaled values and apply VALUE_SCALE once at the end, which is
	// bit-identical to scaling all 8 (Value) or 64 (ValueCubic) corners first and saves that many multiplies.
	constexpr float VALUE_SCALE = 1.0f / 2147483648.0f;
	UPG_HD float value_bits_from_hash(int32_t h)
	{
		h = imul(h, h);
		h ^= ishl(h, 19);
		return (float)h;
	}
	UPG_HD float value_from_hash(int32_t h) { return value_bits_from_hash(h) * VALUE_SCALE; }
	UPG_HD float val_coord(int32_t seed, int32_t xp, int32_t yp, int32_t zp) { return value_from_hash(lattice_hash(seed, xp, yp, zp)); }
	UPG_HD float val_coord(int32_t seed, int32_t xp, int32_t yp) { return value_from_hash(lattice_hash(seed, xp, yp)); }

	// Gradient dot product.  The library indexes a 64-entry table of directions (the 12 cube-edge
	// vectors five times over, then (1,1,0) (0,-1,1) (-1,1,0) (0,-1,-1)) with bits 2..7 of h ^ (h >> 15)
	// and evaluates xd*gx + yd*gy + zd*gz left to right.
	//
	// Entry e: c = e % 12 (fillers 60..63 -> entries 8, 1, 9, 3); group c >> 2 says which component is
	// zero, bits 0 and 1 of c are the signs of the two others.
	UPG_HD void grad_entry(uint32_t e, float &gx, float &gy, float &gz)
	{
		const uint32_t c = e < 60u ? e % 12u : ((0x3918u >> ((e - 60u) * 4u)) & 15u);
		const uint32_t grp = c >> 2; // 0: (0,s0,s1)  1: (s0,0,s1)  2: (s0,s1,0)
		const float s0 = (c & 1u) ? -1.0f : 1.0f, s1 = (c & 2u) ? -1.0f : 1.0f;
		gx = grp == 0u ? 0.0f : s0;
		gy = grp == 0u ? s0 : grp == 1u ? 0.0f : s1;
		gz = grp == 2u ? 0.0f : s1;
	}

	// Device form: decoding (group, signs, fillers) from the hash costs ~20 ALU-pipe instructions per lattice corner, and the
	// ALU pipe (LOP3/SHF/ISETP/FSEL) is what bounds the sampling kernel (profiles/r01_a).  The table lives in shared memory
	// instead: 64 entries x 8 copies of a float4, copy = lane & 7, so that the 8 lanes of every quarter-warp phase of an
	// LDS.128 hit 8 different bank groups whatever their entries are (conflict free).  One SHF + LOP3 + IMAD + LDS.128, then
	// three multiplies and two adds on the FMA pipe -- the very expression the library evaluates.
	// Every kernel that evaluates gradient noise calls grad_table_init() (+ __syncthreads()) before its first sample.
#if defined(__CUDACC__)
	static __shared__ float4 upg_grad_lds[64 * 8];
	// the 64 entries as literals (= grad_entry(e), checked by tests/test_host_terrain.py through the host build)
	static __constant__ float4 upg_grad_const[64] = {
		{ 0, 1, 1, 0 }, { 0, -1, 1, 0 }, { 0, 1, -1, 0 }, { 0, -1, -1, 0 },
		{ 1, 0, 1, 0 }, { -1, 0, 1, 0 }, { 1, 0, -1, 0 }, { -1, 0, -1, 0 },
		{ 1, 1, 0, 0 }, { -1, 1, 0, 0 }, { 1, -1, 0, 0 }, { -1, -1, 0, 0 },
		{ 0, 1, 1, 0 }, { 0, -1, 1, 0 }, { 0, 1, -1, 0 }, { 0, -1, -1, 0 },
		{ 1, 0, 1, 0 }, { -1, 0, 1, 0 }, { 1, 0, -1, 0 }, { -1, 0, -1, 0 },
		{ 1, 1, 0, 0 }, { -1, 1, 0, 0 }, { 1, -1, 0, 0 }, { -1, -1, 0, 0 },
		{ 0, 1, 1, 0 }, { 0, -1, 1, 0 }, { 0, 1, -1, 0 }, { 0, -1, -1, 0 },
		{ 1, 0, 1, 0 }, { -1, 0, 1, 0 }, { 1, 0, -1, 0 }, { -1, 0, -1, 0 },
		{ 1, 1, 0, 0 }, { -1, 1, 0, 0 }, { 1, -1, 0, 0 }, { -1, -1, 0, 0 },
		{ 0, 1, 1, 0 }, { 0, -1, 1, 0 }, { 0, 1, -1, 0 }, { 0, -1, -1, 0 },
		{ 1, 0, 1, 0 }, { -1, 0, 1, 0 }, { 1, 0, -1, 0 }, { -1, 0, -1, 0 },
		{ 1, 1, 0, 0 }, { -1, 1, 0, 0 }, { 1, -1, 0, 0 }, { -1, -1, 0, 0 },
		{ 0, 1, 1, 0 }, { 0, -1, 1, 0 }, { 0, 1, -1, 0 }, { 0, -1, -1, 0 },
		{ 1, 0, 1, 0 }, { -1, 0, 1, 0 }, { 1, 0, -1, 0 }, { -1, 0, -1, 0 },
		{ 1, 1, 0, 0 }, { -1, 1, 0, 0 }, { 1, -1, 0, 0 }, { -1, -1, 0, 0 },
		{ 1, 1, 0, 0 }, { 0, -1, 1, 0 }, { -1, 1, 0, 0 }, { 0, -1, -1, 0 },
	};
#endif
	template<int BLOCK_THREADS>
	UPG_HD void grad_table_init()
	{
#if defined(__CUDA_ARCH__)
#pragma unroll
		for (uint32_t i = threadIdx.x; i < 64u * 8u; i += BLOCK_THREADS)
			upg_grad_lds[i] = upg_grad_const[i >> 3];
#endif
	}
	// the gradient a lattice hash selects
	UPG_HD void grad_vec(int32_t h, float &gx, float &gy, float &gz)
	{
#if defined(__CUDA_ARCH__)
		const uint32_t t = ((uint32_t)h ^ ((uint32_t)(h >> 15))) & (63u << 2);
		const float4 g = upg_grad_lds[t * 2u + (threadIdx.x & 7u)];
		gx = g.x, gy = g.y, gz = g.z;
#else
		h ^= h >> 15;
		grad_entry(((uint32_t)h >> 2) & 63u, gx, gy, gz);
#endif
	}
	UPG_HD float grad_dot(int32_t h, float xd, float yd, float zd)
	{
		float gx, gy, gz;
		grad_vec(h, gx, gy, gz);
		return (xd * gx + yd * gy) + zd * gz;
	}
	UPG_HD float grad_coord(int32_t seed, int32_t xp, int32_t yp, int32_t zp, float xd, float yd, float zd) { return grad_dot(lattice_hash(seed, xp, yp, zp), xd, yd, zd); }

	// The library's FastFloor / FastRound (note floor(-2.0) = -3 and the NaN case: the comparison is false, so the second branch).  On
	// the device the select is a predicated add: convert, compare "less or unordered", decrement -- three instructions where the
	// compiler's rendering of the ternary took five (5.5 % of the dense sampling kernel's instructions were these, profiles/r01_zl).
	UPG_HD int32_t fast_floor(float f)
	{
#if defined(__CUDA_ARCH__)
		int32_t i;
		asm("{\n\t.reg .pred p;\n\tcvt.rzi.s32.f32 %0, %1;\n\tsetp.ltu.f32 p, %1, 0f00000000;\n\t@p add.s32 %0, %0, -1;\n\t}" : "=r"(i) : "f"(f));
		return i;
#else
		return f >= 0 ? (int32_t)f : (int32_t)f - 1;
#endif
	}
	// FastRound: f >= 0 ? (int)(f + 0.5f) : (int)(f - 0.5f).  On the device the select becomes the sign bit of f moved onto the 0.5 (one
	// LOP3): same values for every input -- -0.0f takes the second branch here and the first in the library, both give 0; NaN converts to 0.
	UPG_HD int32_t fast_round(float f)
	{
#if defined(__CUDA_ARCH__)
		return (int32_t)(f + __uint_as_float((__float_as_uint(f) & 0x80000000u) | 0x3f000000u));
#else
		return f >= 0 ? (int32_t)(f + 0.5f) : (int32_t)(f - 0.5f);
#endif
	}
	UPG_HD float nlerp(float a, float b, float t) { return a + t * (b - a); }
	UPG_HD float interp_hermite(float t) { return t * t * (3 - 2 * t); }
	UPG_HD float interp_quintic(float t) { return t * t * t * (t * (t * 6 - 15) + 10); }
	UPG_HD float cubic_lerp(float a, float b, float c, float d, float t)
	{
		const float p = (d - c) - (a - b);
		return t * t * t * p + t * t * ((a - b) - p) + t * (c - a) + b;
	}

	// ---- single-octave generators --------------------------------------------------------------

	UPG_HD float single_value(int32_t seed, float x, float y, float z)
	{
		int32_t x0 = fast_floor(x), y0 = fast_floor(y), z0 = fast_floor(z);
		const float xs = interp_hermite(x - (float)x0);
		const float ys = interp_hermite(y - (float)y0);
		const float zs = interp_hermite(z - (float)z0);
		x0 = imul(x0, PRIME_X);
		y0 = imul(y0, PRIME_Y);
		z0 = imul(z0, PRIME_Z);
		const int32_t x1 = iadd(x0, PRIME_X), y1 = iadd(y0, PRIME_Y), z1 = iadd(z0, PRIME_Z);
		// hoist the seed ^ y ^ z part of the eight hashes
		const int32_t s00 = seed ^ y0 ^ z0, s10 = seed ^ y1 ^ z0, s01 = seed ^ y0 ^ z1, s11 = seed ^ y1 ^ z1;
		// unscaled corner values, one exact power-of-two scale at the end (see VALUE_SCALE)
		const float xf00 = nlerp(value_bits_from_hash(imul(s00 ^ x0, 0x27d4eb2d)), value_bits_from_hash(imul(s00 ^ x1, 0x27d4eb2d)), xs);
		const float xf10 = nlerp(value_bits_from_hash(imul(s10 ^ x0, 0x27d4eb2d)), value_bits_from_hash(imul(s10 ^ x1, 0x27d4eb2d)), xs);
		const float xf01 = nlerp(value_bits_from_hash(imul(s01 ^ x0, 0x27d4eb2d)), value_bits_from_hash(imul(s01 ^ x1, 0x27d4eb2d)), xs);
		const float xf11 = nlerp(value_bits_from_hash(imul(s11 ^ x0, 0x27d4eb2d)), value_bits_from_hash(imul(s11 ^ x1, 0x27d4eb2d)), xs);
		const float yf0 = nlerp(xf00, xf10, ys);
		const float yf1 = nlerp(xf01, xf11, ys);
		return nlerp(yf0, yf1, zs) * VALUE_SCALE;
	}

	// 2-D value noise: what cage's evaluate(Real) reaches as GetNoise(x, 0) (noiseFunction.cpp:111-115)
	UPG_HD float single_value2(int32_t seed, float x, float y)
	{
		int32_t x0 = fast_floor(x), y0 = fast_floor(y);
		const float xs = interp_hermite(x - (float)x0);
		const float ys = interp_hermite(y - (float)y0);
		x0 = imul(x0, PRIME_X);
		y0 = imul(y0, PRIME_Y);
		const int32_t x1 = iadd(x0, PRIME_X), y1 = iadd(y0, PRIME_Y);
		const float xf0 = nlerp(val_coord(seed, x0, y0), val_coord(seed, x1, y0), xs);
		const float xf1 = nlerp(val_coord(seed, x0, y1), val_coord(seed, x1, y1), xs);
		return nlerp(xf0, xf1, ys);
	}

	UPG_HD float single_perlin(int32_t seed, float x, float y, float z)
	{
		int32_t x0 = fast_floor(x), y0 = fast_floor(y), z0 = fast_floor(z);
		const float xd0 = x - (float)x0, yd0 = y - (float)y0, zd0 = z - (float)z0;
		const float xd1 = xd0 - 1, yd1 = yd0 - 1, zd1 = zd0 - 1;
		const float xs = interp_quintic(xd0), ys = interp_quintic(yd0), zs = interp_quintic(zd0);
		x0 = imul(x0, PRIME_X);
		y0 = imul(y0, PRIME_Y);
		z0 = imul(z0, PRIME_Z);
		const int32_t x1 = iadd(x0, PRIME_X), y1 = iadd(y0, PRIME_Y), z1 = iadd(z0, PRIME_Z);
		const int32_t s00 = seed ^ y0 ^ z0, s10 = seed ^ y1 ^ z0, s01 = seed ^ y0 ^ z1, s11 = seed ^ y1 ^ z1;
		const float xf00 = nlerp(grad_dot(imul(s00 ^ x0, 0x27d4eb2d), xd0, yd0, zd0), grad_dot(imul(s00 ^ x1, 0x27d4eb2d), xd1, yd0, zd0), xs);
		const float xf10 = nlerp(grad_dot(imul(s10 ^ x0, 0x27d4eb2d), xd0, yd1, zd0), grad_dot(imul(s10 ^ x1, 0x27d4eb2d), xd1, yd1, zd0), xs);
		const float xf01 = nlerp(grad_dot(imul(s01 ^ x0, 0x27d4eb2d), xd0, yd0, zd1), grad_dot(imul(s01 ^ x1, 0x27d4eb2d), xd1, yd0, zd1), xs);
		const float xf11 = nlerp(grad_dot(imul(s11 ^ x0, 0x27d4eb2d), xd0, yd1, zd1), grad_dot(imul(s11 ^ x1, 0x27d4eb2d), xd1, yd1, zd1), xs);
		const float yf0 = nlerp(xf00, xf10, ys);
		const float yf1 = nlerp(xf01, xf11, ys);
		return nlerp(yf0, yf1, zs) * 0.964921414852142333984375f;
	}

	// Warp-cooperative form of the x stage (device only, UPG_CUBIC_COOP).  The colouring kernel evaluates ValueCubic at frequencies
	// whose lattice cells are 6 - 1000 world units wide while the 32 texels of a warp are an 8 x 4 tile of neighbours ~1 unit apart
	// (upg_bake.cu: TileOrder): the lanes of a warp sit in one cell, or in two or three, so the 64 lattice values -- and therefore
	// the four polynomial coefficients p, (a - b) - p, c - a, b that cubic_lerp() derives from each row of four of them -- are the
	// same for all lanes of a cell.  Per distinct cell, lanes 0..15 compute one row each (4 hashes instead of 64 per lane) into a
	// per-warp shared-memory block, and the lanes of that cell read the rows back as broadcast LDS.128 and evaluate
	// ((ttt * p + tt * q) + t * r) + b: the expression and operation order of cubic_lerp(), bit for bit (~50 instructions per
	// distinct cell + ~200 once, against 650 per lane; profiles/r01_zl_summary.md).  Warps that are not whole, and lanes whose cell
	// found none of the CUBIC_COOP_CELLS slots, take the per-lane path below.
#if !defined(UPG_CUBIC_COOP)
#define UPG_CUBIC_COOP 1
#endif
#if defined(__CUDACC__) && UPG_CUBIC_COOP
	// one block of 16 rows per distinct cell of the warp (CUBIC_COOP_CELLS of them: 2 KB per warp); warps beyond CUBIC_COOP_WARPS of a
	// block go per lane.  The colouring kernel runs 12 warps per block (upg_bake.cu: COLOR_BLOCK).
	constexpr int CUBIC_COOP_WARPS = 12;
	constexpr int CUBIC_COOP_CELLS = 8; // the 8 x 4 texel tiles of a bake need 1 - 6 at the frequencies of the colouring noises
	static __shared__ float4 upg_cubic_rows[CUBIC_COOP_WARPS][CUBIC_COOP_CELLS][16];
#endif

	UPG_HD float single_cubic(int32_t seed, float x, float y, float z)
	{
		int32_t x1 = fast_floor(x), y1 = fast_floor(y), z1 = fast_floor(z);
		const float xs = x - (float)x1, ys = y - (float)y1, zs = z - (float)z1;
		x1 = imul(x1, PRIME_X);
		y1 = imul(y1, PRIME_Y);
		z1 = imul(z1, PRIME_Z);
		const int32_t xi0 = isub(x1, PRIME_X), xi2 = iadd(x1, PRIME_X), xi3 = iadd(x1, ishl(PRIME_X, 1));
		const int32_t yb = isub(y1, PRIME_Y), zb = isub(z1, PRIME_Z);
#if defined(__CUDA_ARCH__) && UPG_CUBIC_COOP
		if (__activemask() == 0xffffffffu && (threadIdx.x >> 5) < (uint32_t)CUBIC_COOP_WARPS)
		{
			const uint32_t full = 0xffffffffu;
			const uint32_t lane = threadIdx.x & 31u;
			float4(*cells)[16] = upg_cubic_rows[threadIdx.x >> 5];
			// STAGE: one round per distinct lattice cell among the lanes (the primes are odd: equal products <=> equal cells); `todo` and
			// `cell` are warp-uniform.  Lanes 0..15 hash one row of the leader's cell each; the lanes of that cell note its slot.
			uint32_t todo = full;
			int cell = 0, mySlot = -1;
			do
			{
				const int leader = __ffs((int)todo) - 1;
				const int32_t ux = __shfl_sync(full, x1, leader), uy = __shfl_sync(full, y1, leader), uz = __shfl_sync(full, z1, leader);
				const bool mine = (x1 == ux) & (y1 == uy) & (z1 == uz);
				const uint32_t group = __ballot_sync(full, mine);
				if (mine)
					mySlot = cell;
				if (lane < 16u)
				{
					const int32_t s = seed ^ iadd(isub(uy, PRIME_Y), imul((int32_t)(lane & 3u), PRIME_Y)) ^ iadd(isub(uz, PRIME_Z), imul((int32_t)(lane >> 2), PRIME_Z));
					const float a = value_bits_from_hash(imul(s ^ isub(ux, PRIME_X), 0x27d4eb2d)), b = value_bits_from_hash(imul(s ^ ux, 0x27d4eb2d));
					const float c = value_bits_from_hash(imul(s ^ iadd(ux, PRIME_X), 0x27d4eb2d)), d = value_bits_from_hash(imul(s ^ iadd(ux, ishl(PRIME_X, 1)), 0x27d4eb2d));
					const float p = (d - c) - (a - b);
					cells[cell][lane] = make_float4(p, (a - b) - p, c - a, b);
				}
				todo &= ~group;
			} while (todo && ++cell < CUBIC_COOP_CELLS);
			__syncwarp();
			// EVALUATE once: every served lane reads the 16 rows of its cell (lanes of one cell: one broadcast) -- cubic_lerp()'s
			// ((ttt * p + tt * q) + t * r) + b per row, then the y and z stages as below
			float result = 0;
			if (mySlot >= 0)
			{
				const float4 *rows = cells[mySlot];
				const float tt = xs * xs, ttt = tt * xs;
				float zr[4];
#pragma unroll
				for (int k = 0; k < 4; k++)
				{
					float yr[4];
#pragma unroll
					for (int j = 0; j < 4; j++)
					{
						const float4 q = rows[k * 4 + j];
						yr[j] = ttt * q.x + tt * q.y + xs * q.z + q.w;
					}
					zr[k] = cubic_lerp(yr[0], yr[1], yr[2], yr[3], ys);
				}
				result = (cubic_lerp(zr[0], zr[1], zr[2], zr[3], zs) * VALUE_SCALE) * (1 / (1.5f * 1.5f * 1.5f));
			}
			__syncwarp(); // the rows are rewritten by the next octave
			// lanes whose cell found no slot (incoherent positions: upg_tiles on arbitrary points) go per lane below
			if (mySlot >= 0)
				return result;
		}
#endif
		float zr[4];
		int32_t zp = zb;
#pragma unroll
		for (int k = 0; k < 4; k++)
		{
			float yr[4];
			int32_t yp = yb;
#pragma unroll
			for (int j = 0; j < 4; j++)
			{
				const int32_t s = seed ^ yp ^ zp;
				yr[j] = cubic_lerp(value_bits_from_hash(imul(s ^ xi0, 0x27d4eb2d)), value_bits_from_hash(imul(s ^ x1, 0x27d4eb2d)), value_bits_from_hash(imul(s ^ xi2, 0x27d4eb2d)), value_bits_from_hash(imul(s ^ xi3, 0x27d4eb2d)), xs);
				yp = iadd(yp, PRIME_Y);
			}
			zr[k] = cubic_lerp(yr[0], yr[1], yr[2], yr[3], ys);
			zp = iadd(zp, PRIME_Z);
		}
		// corners unscaled, one exact power-of-two scale here (see VALUE_SCALE)
		return (cubic_lerp(zr[0], zr[1], zr[2], zr[3], zs) * VALUE_SCALE) * (1 / (1.5f * 1.5f * 1.5f));
	}

	// OpenSimplex2, 3-D: two half-offset cubic lattices (seed and ~seed), <= 2 kernel terms from each
	UPG_HD float single_simplex(int32_t seed, float x, float y, float z)
	{
		int32_t i = fast_round(x), j = fast_round(y), k = fast_round(z);
		float x0 = x - (float)i, y0 = y - (float)j, z0 = z - (float)k;

		int32_t xNSign = (int32_t)(-1.0f - x0) | 1;
		int32_t yNSign = (int32_t)(-1.0f - y0) | 1;
		int32_t zNSign = (int32_t)(-1.0f - z0) | 1;

		float ax0 = (float)xNSign * -x0;
		float ay0 = (float)yNSign * -y0;
		float az0 = (float)zNSign * -z0;

		i = imul(i, PRIME_X);
		j = imul(j, PRIME_Y);
		k = imul(k, PRIME_Z);

		float value = 0;
		float a = (0.6f - x0 * x0) - (y0 * y0 + z0 * z0);

#pragma unroll
		for (int l = 0; l < 2; l++)
		{
			if (a > 0)
				value += (a * a) * (a * a) * grad_coord(seed, i, j, k, x0, y0, z0);

			// the library's three-way branch on the dominant axis, as selects feeding ONE gradient term: the branch is
			// lane-divergent almost everywhere, so the warp used to run all three bodies back to back
			{
				const bool xm = ax0 >= ay0 && ax0 >= az0;
				const bool ym = !xm && (ay0 > ax0 && ay0 >= az0);
				const bool zm = !xm && !ym;
				const float am = xm ? ax0 : ym ? ay0 : az0;
				float b = a + am + am;
				if (b > 1)
				{
					b -= 1;
					const int32_t ii = xm ? isub(i, imul(xNSign, PRIME_X)) : i;
					const int32_t jj = ym ? isub(j, imul(yNSign, PRIME_Y)) : j;
					const int32_t kk = zm ? isub(k, imul(zNSign, PRIME_Z)) : k;
					const float xx = xm ? x0 + (float)xNSign : x0;
					const float yy = ym ? y0 + (float)yNSign : y0;
					const float zz = zm ? z0 + (float)zNSign : z0;
					value += (b * b) * (b * b) * grad_coord(seed, ii, jj, kk, xx, yy, zz);
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

			i = iadd(i, (xNSign >> 1) & PRIME_X);
			j = iadd(j, (yNSign >> 1) & PRIME_Y);
			k = iadd(k, (zNSign >> 1) & PRIME_Z);

			xNSign = -xNSign;
			yNSign = -yNSign;
			zNSign = -zNSign;

			seed = ~seed;
		}
		return value * 32.69428253173828125f;
	}

	// OpenSimplex2S, 3-D: floor-based, kernel 0.75 - r^2; two fixed terms + up to eight optional ones
	UPG_HD float s2s_term(float a, int32_t seed, int32_t xp, int32_t yp, int32_t zp, float dx, float dy, float dz) { return (a * a) * (a * a) * grad_coord(seed, xp, yp, zp, dx, dy, dz); }

	UPG_HD float single_simplex_s(int32_t seed, float x, float y, float z)
	{
		int32_t i = fast_floor(x), j = fast_floor(y), k = fast_floor(z);
		const float xi = x - (float)i, yi = y - (float)j, zi = z - (float)k;

		i = imul(i, PRIME_X);
		j = imul(j, PRIME_Y);
		k = imul(k, PRIME_Z);
		const int32_t seed2 = iadd(seed, 1293373);

		const int32_t xNMask = (int32_t)(-0.5f - xi);
		const int32_t yNMask = (int32_t)(-0.5f - yi);
		const int32_t zNMask = (int32_t)(-0.5f - zi);

		const float x0 = xi + (float)xNMask;
		const float y0 = yi + (float)yNMask;
		const float z0 = zi + (float)zNMask;
		const float a0 = 0.75f - x0 * x0 - y0 * y0 - z0 * z0;
		float value = s2s_term(a0, seed, iadd(i, xNMask & PRIME_X), iadd(j, yNMask & PRIME_Y), iadd(k, zNMask & PRIME_Z), x0, y0, z0);

		const float x1 = xi - 0.5f;
		const float y1 = yi - 0.5f;
		const float z1 = zi - 0.5f;
		const float a1 = 0.75f - x1 * x1 - y1 * y1 - z1 * z1;
		value += s2s_term(a1, seed2, iadd(i, PRIME_X), iadd(j, PRIME_Y), iadd(k, PRIME_Z), x1, y1, z1);

		const float xAFlipMask0 = (float)((xNMask | 1) << 1) * x1;
		const float yAFlipMask0 = (float)((yNMask | 1) << 1) * y1;
		const float zAFlipMask0 = (float)((zNMask | 1) << 1) * z1;
		const float xAFlipMask1 = (float)(-2 - (xNMask << 2)) * x1 - 1.0f;
		const float yAFlipMask1 = (float)(-2 - (yNMask << 2)) * y1 - 1.0f;
		const float zAFlipMask1 = (float)(-2 - (zNMask << 2)) * z1 - 1.0f;

		const int32_t PX2 = ishl(PRIME_X, 1), PY2 = ishl(PRIME_Y, 1), PZ2 = ishl(PRIME_Z, 1);
		const float xs = (float)(xNMask | 1), ys = (float)(yNMask | 1), zs = (float)(zNMask | 1);

		bool skip5 = false;
		const float a2 = xAFlipMask0 + a0;
		if (a2 > 0)
		{
			value += s2s_term(a2, seed, iadd(i, ~xNMask & PRIME_X), iadd(j, yNMask & PRIME_Y), iadd(k, zNMask & PRIME_Z), x0 - xs, y0, z0);
		}
		else
		{
			const float a3 = yAFlipMask0 + zAFlipMask0 + a0;
			if (a3 > 0)
				value += s2s_term(a3, seed, iadd(i, xNMask & PRIME_X), iadd(j, ~yNMask & PRIME_Y), iadd(k, ~zNMask & PRIME_Z), x0, y0 - ys, z0 - zs);
			const float a4 = xAFlipMask1 + a1;
			if (a4 > 0)
			{
				value += s2s_term(a4, seed2, iadd(i, xNMask & PX2), iadd(j, PRIME_Y), iadd(k, PRIME_Z), xs + x1, y1, z1);
				skip5 = true;
			}
		}

		bool skip9 = false;
		const float a6 = yAFlipMask0 + a0;
		if (a6 > 0)
		{
			value += s2s_term(a6, seed, iadd(i, xNMask & PRIME_X), iadd(j, ~yNMask & PRIME_Y), iadd(k, zNMask & PRIME_Z), x0, y0 - ys, z0);
		}
		else
		{
			const float a7 = xAFlipMask0 + zAFlipMask0 + a0;
			if (a7 > 0)
				value += s2s_term(a7, seed, iadd(i, ~xNMask & PRIME_X), iadd(j, yNMask & PRIME_Y), iadd(k, ~zNMask & PRIME_Z), x0 - xs, y0, z0 - zs);
			const float a8 = yAFlipMask1 + a1;
			if (a8 > 0)
			{
				value += s2s_term(a8, seed2, iadd(i, PRIME_X), iadd(j, yNMask & PY2), iadd(k, PRIME_Z), x1, ys + y1, z1);
				skip9 = true;
			}
		}

		bool skipD = false;
		const float aA = zAFlipMask0 + a0;
		if (aA > 0)
		{
			value += s2s_term(aA, seed, iadd(i, xNMask & PRIME_X), iadd(j, yNMask & PRIME_Y), iadd(k, ~zNMask & PRIME_Z), x0, y0, z0 - zs);
		}
		else
		{
			const float aB = xAFlipMask0 + yAFlipMask0 + a0;
			if (aB > 0)
				value += s2s_term(aB, seed, iadd(i, ~xNMask & PRIME_X), iadd(j, ~yNMask & PRIME_Y), iadd(k, zNMask & PRIME_Z), x0 - xs, y0 - ys, z0);
			const float aC = zAFlipMask1 + a1;
			if (aC > 0)
			{
				value += s2s_term(aC, seed2, iadd(i, PRIME_X), iadd(j, PRIME_Y), iadd(k, zNMask & PZ2), x1, y1, zs + z1);
				skipD = true;
			}
		}

		if (!skip5)
		{
			const float a5 = yAFlipMask1 + zAFlipMask1 + a1;
			if (a5 > 0)
				value += s2s_term(a5, seed2, iadd(i, PRIME_X), iadd(j, yNMask & PY2), iadd(k, zNMask & PZ2), x1, ys + y1, zs + z1);
		}
		if (!skip9)
		{
			const float a9 = xAFlipMask1 + zAFlipMask1 + a1;
			if (a9 > 0)
				value += s2s_term(a9, seed2, iadd(i, xNMask & PX2), iadd(j, PRIME_Y), iadd(k, zNMask & PZ2), xs + x1, y1, zs + z1);
		}
		if (!skipD)
		{
			const float aD = xAFlipMask1 + yAFlipMask1 + a1;
			if (aD > 0)
				value += s2s_term(aD, seed2, iadd(i, xNMask & PX2), iadd(j, yNMask & PY2), iadd(k, PRIME_Z), xs + x1, ys + y1, z1);
		}

		return value * 9.046026385208288f;
	}

	// Cellular: 27 neighbouring cells, one jittered feature point each.  `rv` = the 256 x float4 jitter
	// direction table handed to the planet at creation (see upg.h: upstream's table is not
	// reconstructible, the oracle declares a substitute and the product receives it as data).
	// the 27-cell scan of SingleCellular: two smallest distances and the hash of the closest feature point
	struct CellularScan
	{
		float d0, d1;
		int32_t closest;
	};
	template<int DIST>
	UPG_HD CellularScan cellular_scan(const float *rv, int32_t seed, float x, float y, float z)
	{
		const int32_t xr = fast_round(x), yr = fast_round(y), zr = fast_round(z);
		float d0 = 1e10f, d1 = 1e10f;
		int32_t closest = 0;
		const float jitter = 0.39614353f * 1.0f;

		int32_t xp = imul(xr - 1, PRIME_X);
		const int32_t ypBase = imul(yr - 1, PRIME_Y);
		const int32_t zpBase = imul(zr - 1, PRIME_Z);
#pragma unroll 1
		for (int32_t xi = xr - 1; xi <= xr + 1; xi++)
		{
			int32_t yp = ypBase;
#pragma unroll 1
			for (int32_t yi = yr - 1; yi <= yr + 1; yi++)
			{
				int32_t zp = zpBase;
				for (int32_t zi = zr - 1; zi <= zr + 1; zi++)
				{
					const int32_t h = lattice_hash(seed, xp, yp, zp);
					const int32_t idx = h & (255 << 2);
					const float vx = ((float)xi - x) + rv[idx] * jitter;
					const float vy = ((float)yi - y) + rv[idx | 1] * jitter;
					const float vz = ((float)zi - z) + rv[idx | 2] * jitter;
					float nd;
					if (DIST == CD_MANHATTAN)
						nd = abs_(vx) + abs_(vy) + abs_(vz);
					else if (DIST == CD_HYBRID)
						nd = (abs_(vx) + abs_(vy) + abs_(vz)) + (vx * vx + vy * vy + vz * vz);
					else
						nd = vx * vx + vy * vy + vz * vz;
					d1 = fmax_(fmin_(d1, nd), d0);
					if (nd < d0)
					{
						d0 = nd;
						closest = h;
					}
					zp = iadd(zp, PRIME_Z);
				}
				yp = iadd(yp, PRIME_Y);
			}
			xp = iadd(xp, PRIME_X);
		}
		return CellularScan{ d0, d1, closest };
	}

	// the return kinds of SingleCellular over one scan
	template<int DIST, int OP>
	UPG_HD float cellular_result(CellularScan c)
	{
		float d0 = c.d0, d1 = c.d1;
		if (DIST == CD_EUCLIDEAN && OP >= CO_DISTANCE)
		{
			d0 = sqrtf(d0);
			if (OP >= CO_DISTANCE2)
				d1 = sqrtf(d1);
		}

		switch (OP)
		{
			case CO_CELL:
				return (float)c.closest * (1 / 2147483648.0f);
			case CO_DISTANCE:
				return d0 - 1;
			case CO_DISTANCE2:
				return d1 - 1;
			case CO_ADD:
				return (d1 + d0) * 0.5f - 1;
			case CO_SUBTRACT:
				return d1 - d0 - 1;
			case CO_MULTIPLY:
				return d1 * d0 * 0.5f - 1;
			case CO_DIVIDE:
				return d0 / d1 - 1;
			default:
				return 0;
		}
	}

	template<int DIST, int OP>
	UPG_HD float single_cellular(const float *rv, int32_t seed, float x, float y, float z)
	{
		return cellular_result<DIST, OP>(cellular_scan<DIST>(rv, seed, x, y, z));
	}

	// ---- dispatch + fractal drivers -----------------------------------------------------------------

	template<int TYPE>
	UPG_HD float noise_single(int32_t seed, float x, float y, float z)
	{
		if (TYPE == NT_SIMPLEX)
			return single_simplex(seed, x, y, z);
		if (TYPE == NT_SIMPLEX_REDUCED)
			return single_simplex_s(seed, x, y, z);
		if (TYPE == NT_PERLIN)
			return single_perlin(seed, x, y, z);
		if (TYPE == NT_CUBIC)
			return single_cubic(seed, x, y, z);
		return single_value(seed, x, y, z);
	}

	// The same generators as real (non-inlined) functions.  The colouring kernel calls ~40 noise instances from 13 layers;
	// inlining every octave of every call site made it 49 k SASS instructions (784 KB), 25x the L1.5 instruction cache, and
	// ncu showed the warps waiting for instruction fetch ("no_instruction") most of the time.  One body per generator keeps the
	// hot code resident.  Templates have vague linkage, so there is one copy per library, and none where unused.
#if defined(__CUDACC__)
#define UPG_OUTLINED __host__ __device__ __noinline__
#else
#define UPG_OUTLINED inline
#endif
	template<int TYPE>
	UPG_OUTLINED float noise_single_outlined(int32_t seed, float x, float y, float z)
	{
		return noise_single<TYPE>(seed, x, y, z);
	}
	// one body per distance function; the return kind is a handful of instructions at the call site
	template<int DIST>
	UPG_OUTLINED CellularScan cellular_scan_outlined(const float *rv, int32_t seed, float x, float y, float z)
	{
		return cellular_scan<DIST>(rv, seed, x, y, z);
	}
	template<int DIST, int OP>
	UPG_HD float single_cellular_outlined(const float *rv, int32_t seed, float x, float y, float z)
	{
		return cellular_result<DIST, OP>(cellular_scan_outlined<DIST>(rv, seed, x, y, z));
	}

	UPG_HD float ping_pong(float t)
	{
		t -= (float)((int32_t)(t * 0.5f) * 2);
		return t < 1 ? t : 2 - t;
	}

	template<int FRACTAL>
	UPG_HD float fractal_term(float n)
	{
		if (FRACTAL == FT_RIDGED)
			return abs_(n) * -2 + 1; // library: noise = |n|; (noise * -2 + 1)
		if (FRACTAL == FT_PINGPONG)
			return (ping_pong((n + 1) * 2.0f) - 0.5f) * 2;
		return n;
	}

	// GetNoise(x, y, z): frequency scale, domain rotation for the two simplex kinds, fractal loop.
	// weighted strength stays at the library default 0 (cage never sets it) => amp *= lerp(1, ., 0) == amp.
	template<int TYPE, int FRACTAL, int OCTAVES, bool OUTLINED = false>
	UPG_HD float noise3(const NoiseInst &n, float x, float y, float z)
	{
		// coordinates that reach a floor / round are single rounded operations in every build (mul_rn & co, upg_common.cuh)
		x = mul_rn(x, n.frequency);
		y = mul_rn(y, n.frequency);
		z = mul_rn(z, n.frequency);
		if (TYPE == NT_SIMPLEX || TYPE == NT_SIMPLEX_REDUCED)
		{
			const float R3 = (float)(2.0 / 3.0);
			const float r = mul_rn(add_rn(add_rn(x, y), z), R3);
			x = sub_rn(r, x);
			y = sub_rn(r, y);
			z = sub_rn(r, z);
		}
		if (FRACTAL == FT_NONE)
			return OUTLINED ? noise_single_outlined<TYPE>(n.seed, x, y, z) : noise_single<TYPE>(n.seed, x, y, z);
		int32_t seed = n.seed;
		float sum = 0;
		float amp = n.bounding;
		UPG_OCTAVE_LOOP
		for (int o = 0; o < OCTAVES; o++)
		{
			const float v = OUTLINED ? noise_single_outlined<TYPE>(seed, x, y, z) : noise_single<TYPE>(seed, x, y, z);
			seed = iadd(seed, 1);
			sum += fractal_term<FRACTAL>(v) * amp;
			x = mul_rn(x, n.lacunarity);
			y = mul_rn(y, n.lacunarity);
			z = mul_rn(z, n.lacunarity);
			amp *= n.gain;
		}
		return sum;
	}

	template<int TYPE, int FRACTAL, int OCTAVES, bool OUTLINED = false>
	UPG_HD float noise3(const NoiseInst &n, V3 p)
	{
		return noise3<TYPE, FRACTAL, OCTAVES, OUTLINED>(n, p.x, p.y, p.z);
	}

	// ---- Value fBm along a column of samples ---------------------------------------------------------------------------
	// single_value() interpolates x first, then y, then z: its last step is nlerp(yf0, yf1, zs) with yf0 / yf1 functions of (x, y) and
	// of the lattice planes z0 / z0 + 1 alone.  Samples that share x and y and step through z (a column of the meshing grid) therefore
	// share yf0 / yf1 for as long as they stay between the same two planes -- 24 to 192 consecutive samples for the octaves of the
	// `lakes` land noise on a 1000^3 grid.  The state below keeps both per octave; a sample then costs one floor, one Hermite weight
	// and one nlerp per octave instead of eight hashes and seven nlerps, and every operation that is executed is the one
	// single_value() executes, on the same operands: bit-identical sums (tests/test_host_terrain.py).
	template<int OCTAVES>
	struct ValueFbmColumn
	{
		int32_t z0[OCTAVES]; // lattice plane below the last sample, per octave
		float yf0[OCTAVES], yf1[OCTAVES]; // single_value()'s yf0 / yf1 (unscaled) for those planes
		bool fresh = true; // nothing cached yet
	};

	// yf of single_value() for one lattice plane (zp = z lattice index * PRIME_Z)
	UPG_HD float value_plane(int32_t seed, float x, float y, int32_t zp)
	{
		int32_t x0 = fast_floor(x), y0 = fast_floor(y);
		const float xs = interp_hermite(x - (float)x0);
		const float ys = interp_hermite(y - (float)y0);
		x0 = imul(x0, PRIME_X);
		y0 = imul(y0, PRIME_Y);
		const int32_t x1 = iadd(x0, PRIME_X), y1 = iadd(y0, PRIME_Y);
		const int32_t s0 = seed ^ y0 ^ zp, s1 = seed ^ y1 ^ zp;
		const float xf0 = nlerp(value_bits_from_hash(imul(s0 ^ x0, 0x27d4eb2d)), value_bits_from_hash(imul(s0 ^ x1, 0x27d4eb2d)), xs);
		const float xf1 = nlerp(value_bits_from_hash(imul(s1 ^ x0, 0x27d4eb2d)), value_bits_from_hash(imul(s1 ^ x1, 0x27d4eb2d)), xs);
		return nlerp(xf0, xf1, ys);
	}

	// == noise3<NT_VALUE, FT_FBM, OCTAVES>(n, x, y, z), for successive samples of one (x, y) column in any z order
	template<int OCTAVES>
	UPG_HD float value_fbm_column(const NoiseInst &n, float x, float y, float z, ValueFbmColumn<OCTAVES> &c)
	{
		x = mul_rn(x, n.frequency);
		y = mul_rn(y, n.frequency);
		z = mul_rn(z, n.frequency);
		int32_t seed = n.seed;
		float sum = 0;
		float amp = n.bounding;
#pragma unroll
		for (int o = 0; o < OCTAVES; o++)
		{
			const int32_t z0 = fast_floor(z);
			if (c.fresh || z0 != c.z0[o])
			{
				const int32_t zp = imul(z0, PRIME_Z);
				// one plane up: the old upper plane is the new lower one
				c.yf0[o] = (!c.fresh && z0 == iadd(c.z0[o], 1)) ? c.yf1[o] : value_plane(seed, x, y, zp);
				c.yf1[o] = value_plane(seed, x, y, iadd(zp, PRIME_Z));
				c.z0[o] = z0;
			}
			const float zs = interp_hermite(z - (float)z0);
			const float v = nlerp(c.yf0[o], c.yf1[o], zs) * VALUE_SCALE;
			seed = iadd(seed, 1);
			sum += fractal_term<FT_FBM>(v) * amp;
			x = mul_rn(x, n.lacunarity);
			y = mul_rn(y, n.lacunarity);
			z = mul_rn(z, n.lacunarity);
			amp *= n.gain;
		}
		c.fresh = false;
		return sum;
	}

	// ---- Perlin along a column of samples ----------------------------------------------------------------------------------
	// single_perlin()'s corner term is (xd * gx + yd * gy) + zd * gz.  Along a z column the hash, the gradient and the bracket are fixed for
	// as long as the samples stay between two lattice planes; what remains per sample is zd * gz + bracket for the eight corners, the
	// seven lerps (x and y weights are fixed too) and one quintic weight.  Same operations on the same operands as single_perlin().
	struct PerlinColumn
	{
		int32_t z0; // lattice plane below the last sample
		float xy[8], gz[8]; // bracket and z component per corner: 0..3 lower plane (x0y0, x1y0, x0y1, x1y1), 4..7 upper plane
		float xs, ys; // quintic weights of x and y
		bool fresh = true;
	};

	UPG_HD void perlin_plane(int32_t seed, float x, float y, int32_t zp, float *xy4, float *gz4, float &xs, float &ys)
	{
		int32_t x0 = fast_floor(x), y0 = fast_floor(y);
		const float xd0 = x - (float)x0, yd0 = y - (float)y0;
		const float xd1 = xd0 - 1, yd1 = yd0 - 1;
		xs = interp_quintic(xd0);
		ys = interp_quintic(yd0);
		x0 = imul(x0, PRIME_X);
		y0 = imul(y0, PRIME_Y);
		const int32_t x1 = iadd(x0, PRIME_X), y1 = iadd(y0, PRIME_Y);
		const int32_t s0 = seed ^ y0 ^ zp, s1 = seed ^ y1 ^ zp;
		float gx, gy;
		grad_vec(imul(s0 ^ x0, 0x27d4eb2d), gx, gy, gz4[0]);
		xy4[0] = xd0 * gx + yd0 * gy;
		grad_vec(imul(s0 ^ x1, 0x27d4eb2d), gx, gy, gz4[1]);
		xy4[1] = xd1 * gx + yd0 * gy;
		grad_vec(imul(s1 ^ x0, 0x27d4eb2d), gx, gy, gz4[2]);
		xy4[2] = xd0 * gx + yd1 * gy;
		grad_vec(imul(s1 ^ x1, 0x27d4eb2d), gx, gy, gz4[3]);
		xy4[3] = xd1 * gx + yd1 * gy;
	}

	// == single_perlin(seed, x, y, z) for successive samples of one (x, y) column in any z order
	UPG_HD float perlin_column(int32_t seed, float x, float y, float z, PerlinColumn &c)
	{
		const int32_t z0 = fast_floor(z);
		if (c.fresh || z0 != c.z0)
		{
			const int32_t zp = imul(z0, PRIME_Z);
			if (!c.fresh && z0 == iadd(c.z0, 1))
			{
#pragma unroll
				for (int k = 0; k < 4; k++)
				{
					c.xy[k] = c.xy[4 + k];
					c.gz[k] = c.gz[4 + k];
				}
			}
			else
				perlin_plane(seed, x, y, zp, c.xy, c.gz, c.xs, c.ys);
			perlin_plane(seed, x, y, iadd(zp, PRIME_Z), c.xy + 4, c.gz + 4, c.xs, c.ys);
			c.z0 = z0;
			c.fresh = false;
		}
		const float zd0 = z - (float)z0;
		const float zd1 = zd0 - 1;
		const float zs = interp_quintic(zd0);
		const float xf00 = nlerp(c.xy[0] + zd0 * c.gz[0], c.xy[1] + zd0 * c.gz[1], c.xs);
		const float xf10 = nlerp(c.xy[2] + zd0 * c.gz[2], c.xy[3] + zd0 * c.gz[3], c.xs);
		const float xf01 = nlerp(c.xy[4] + zd1 * c.gz[4], c.xy[5] + zd1 * c.gz[5], c.xs);
		const float xf11 = nlerp(c.xy[6] + zd1 * c.gz[6], c.xy[7] + zd1 * c.gz[7], c.xs);
		const float yf0 = nlerp(xf00, xf10, c.ys);
		const float yf1 = nlerp(xf01, xf11, c.ys);
		return nlerp(yf0, yf1, zs) * 0.964921414852142333984375f;
	}

	// == noise3<NT_PERLIN, FT_FBM, OCTAVES>(n, x, y, z) for successive samples of one (x, y) column: one PerlinColumn per octave
	template<int OCTAVES>
	struct PerlinFbmColumn
	{
		PerlinColumn octave[OCTAVES];
	};
	template<int OCTAVES>
	UPG_HD float perlin_fbm_column(const NoiseInst &n, float x, float y, float z, PerlinFbmColumn<OCTAVES> &c)
	{
		x = mul_rn(x, n.frequency);
		y = mul_rn(y, n.frequency);
		z = mul_rn(z, n.frequency);
		int32_t seed = n.seed;
		float sum = 0;
		float amp = n.bounding;
#pragma unroll
		for (int o = 0; o < OCTAVES; o++)
		{
			const float v = perlin_column(seed, x, y, z, c.octave[o]);
			seed = iadd(seed, 1);
			sum += fractal_term<FT_FBM>(v) * amp;
			x = mul_rn(x, n.lacunarity);
			y = mul_rn(y, n.lacunarity);
			z = mul_rn(z, n.lacunarity);
			amp *= n.gain;
		}
		return sum;
	}

	// cellular variant (needs the jitter table and the distance / return kind)
	template<int DIST, int OP, int FRACTAL, int OCTAVES>
	UPG_HD float noise3_cellular(const float *rv, const NoiseInst &n, float x, float y, float z)
	{
		x = mul_rn(x, n.frequency);
		y = mul_rn(y, n.frequency);
		z = mul_rn(z, n.frequency);
		if (FRACTAL == FT_NONE)
			return single_cellular_outlined<DIST, OP>(rv, n.seed, x, y, z);
		int32_t seed = n.seed;
		float sum = 0;
		float amp = n.bounding;
#pragma unroll 1
		for (int o = 0; o < OCTAVES; o++)
		{
			const float v = single_cellular_outlined<DIST, OP>(rv, seed, x, y, z);
			seed = iadd(seed, 1);
			sum += fractal_term<FRACTAL>(v) * amp;
			x = mul_rn(x, n.lacunarity);
			y = mul_rn(y, n.lacunarity);
			z = mul_rn(z, n.lacunarity);
			amp *= n.gain;
		}
		return sum;
	}

	// TWO cellular noises that share seed, frequency, fractal parameters and position but differ in the return kind -- the reference
	// builds such pairs from one `static const uint32 seed` (bedrock cracks / value, grass patches, ice: sources/coloringModes.cpp:221-240,
	// 480-509, 767-790) -- scan the same 27 cells per octave: one scan serves both.  a = noise3_cellular<DIST, OP_A, ...>(rv, n, ...),
	// b = noise3_cellular<DIST, OP_B, ...>(rv, n, ...), bit for bit.
	template<int DIST, int OP_A, int OP_B, int FRACTAL, int OCTAVES>
	UPG_HD void noise3_cellular_pair(const float *rv, const NoiseInst &n, float x, float y, float z, float &a, float &b)
	{
		x = mul_rn(x, n.frequency);
		y = mul_rn(y, n.frequency);
		z = mul_rn(z, n.frequency);
		if (FRACTAL == FT_NONE)
		{
			const CellularScan c = cellular_scan_outlined<DIST>(rv, n.seed, x, y, z);
			a = cellular_result<DIST, OP_A>(c);
			b = cellular_result<DIST, OP_B>(c);
			return;
		}
		int32_t seed = n.seed;
		float sumA = 0, sumB = 0;
		float amp = n.bounding;
#pragma unroll 1
		for (int o = 0; o < OCTAVES; o++)
		{
			const CellularScan c = cellular_scan_outlined<DIST>(rv, seed, x, y, z);
			seed = iadd(seed, 1);
			sumA += fractal_term<FRACTAL>(cellular_result<DIST, OP_A>(c)) * amp;
			sumB += fractal_term<FRACTAL>(cellular_result<DIST, OP_B>(c)) * amp;
			x = mul_rn(x, n.lacunarity);
			y = mul_rn(y, n.lacunarity);
			z = mul_rn(z, n.lacunarity);
			amp *= n.gain;
		}
		a = sumA;
		b = sumB;
	}

	// GetNoise(x, 0) for the Value kind, no fractal (sources/coloringModes.cpp:36-48)
	UPG_HD float noise1_value(const NoiseInst &n, float x)
	{
		return single_value2(n.seed, mul_rn(x, n.frequency), mul_rn(0.0f, n.frequency));
	}
}
