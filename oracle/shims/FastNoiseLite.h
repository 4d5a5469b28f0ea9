/*
 * FastNoiseLite.h -- TEST INFRASTRUCTURE (oracle shim).
 *
 * Stand-in for the un-vendored submodule Auburn/FastNoiseLite (pinned 785f37a9...), so that
 * cage's `libcore/noiseFunction.cpp` (which does `#include <FastNoiseLite.h>`, line 3) can be
 * compiled VERBATIM from /root/reference into oracle/_ref.  Only the members cage touches
 * (noiseFunction.cpp:14-127) exist.  The arithmetic lives in ../up_noise.h (a restatement of
 * the library's published algorithm; see the parity note there).
 */
#ifndef UP_ORACLE_FASTNOISELITE_SHIM_H
#define UP_ORACLE_FASTNOISELITE_SHIM_H

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../up_noise.h"

class FastNoiseLite;
/* every noise object the reference constructs, in construction order (== noiseSeed() order per SURVEY A.6) */
/* defined once in ref_harness.cpp (a plain global: one per loaded copy of the library) */
std::vector<const FastNoiseLite *> &upo_shim_registry();

class FastNoiseLite
{
public:
	enum NoiseType
	{
		NoiseType_OpenSimplex2,
		NoiseType_OpenSimplex2S,
		NoiseType_Cellular,
		NoiseType_Perlin,
		NoiseType_ValueCubic,
		NoiseType_Value
	};
	enum FractalType
	{
		FractalType_None,
		FractalType_FBm,
		FractalType_Ridged,
		FractalType_PingPong
	};
	enum CellularDistanceFunction
	{
		CellularDistanceFunction_Euclidean,
		CellularDistanceFunction_EuclideanSq,
		CellularDistanceFunction_Manhattan,
		CellularDistanceFunction_Hybrid
	};
	enum CellularReturnType
	{
		CellularReturnType_CellValue,
		CellularReturnType_Distance,
		CellularReturnType_Distance2,
		CellularReturnType_Distance2Add,
		CellularReturnType_Distance2Sub,
		CellularReturnType_Distance2Mul,
		CellularReturnType_Distance2Div
	};

	explicit FastNoiseLite(int seed = 1337)
	{
		cfg = upo_noise_default((uint32_t)seed);
		cfg.frequency = 0.01f; /* library default, cage always overrides */
		cfg.distance = UPO_DIST_EUCLIDEAN_SQ;
		cfg.operation = UPO_OP_DISTANCE;
		upo_noise_finish(&cfg);
		upo_shim_registry().push_back(this);
	}
	FastNoiseLite(const FastNoiseLite &) = delete;
	FastNoiseLite &operator=(const FastNoiseLite &) = delete;

	void SetFrequency(float f) { cfg.frequency = f; }
	void SetNoiseType(NoiseType t)
	{
		static const int map[] = { UPO_NOISE_SIMPLEX, UPO_NOISE_SIMPLEX_REDUCED, UPO_NOISE_CELLULAR, UPO_NOISE_PERLIN, UPO_NOISE_CUBIC, UPO_NOISE_VALUE };
		cfg.type = map[(int)t];
	}
	void SetFractalType(FractalType t) { cfg.fractal = (int)t; }
	void SetFractalOctaves(int o)
	{
		cfg.octaves = o;
		upo_noise_finish(&cfg);
	}
	void SetFractalLacunarity(float l) { cfg.lacunarity = l; }
	void SetFractalGain(float g)
	{
		cfg.gain = g;
		upo_noise_finish(&cfg);
	}
	void SetCellularDistanceFunction(CellularDistanceFunction d) { cfg.distance = (int)d; }
	void SetCellularReturnType(CellularReturnType r) { cfg.operation = (int)r; }

	float GetNoise(float x, float y) const
	{
		if (!upo_noise2_supported(&cfg))
		{
			std::fprintf(stderr, "oracle FastNoiseLite shim: 2-D noise of type %d is not on the hot path and is not restated\n", cfg.type);
			std::abort();
		}
		return upo_noise2(&cfg, x, y);
	}
	float GetNoise(float x, float y, float z) const { return upo_noise3(&cfg, x, y, z); }

	upo_noise_cfg cfg;
};

#endif
