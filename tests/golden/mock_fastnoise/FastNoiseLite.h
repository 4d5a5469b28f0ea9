/*
 * TEST FIXTURE -- NOT Auburn/FastNoiseLite.  A stand-in with the real header's public interface (the members cage touches,
 * externals/cage/sources/libcore/noiseFunction.cpp:14-127) and its PRIVATE constant-table struct, used by tests/test_fastnoise_pin.py to
 * exercise the pin hook (`make -C oracle ref FASTNOISE=<header>`, tools/fastnoise_pin.py) in a container that does not have the real
 * header.  The arithmetic delegates to the oracle's restatement, so the hook must report zero difference for every generator; the
 * table it exposes is deliberately NOT the declared substitute, so the test can tell that the tool extracted it from "upstream".
 */
#ifndef MOCK_FASTNOISELITE_H
#define MOCK_FASTNOISELITE_H

#include "up_noise.h"

class FastNoiseLite
{
public:
	enum NoiseType { NoiseType_OpenSimplex2, NoiseType_OpenSimplex2S, NoiseType_Cellular, NoiseType_Perlin, NoiseType_ValueCubic, NoiseType_Value };
	enum FractalType { FractalType_None, FractalType_FBm, FractalType_Ridged, FractalType_PingPong };
	enum CellularDistanceFunction { CellularDistanceFunction_Euclidean, CellularDistanceFunction_EuclideanSq, CellularDistanceFunction_Manhattan, CellularDistanceFunction_Hybrid };
	enum CellularReturnType { CellularReturnType_CellValue, CellularReturnType_Distance, CellularReturnType_Distance2, CellularReturnType_Distance2Add, CellularReturnType_Distance2Sub, CellularReturnType_Distance2Mul, CellularReturnType_Distance2Div };

	explicit FastNoiseLite(int seed = 1337)
	{
		cfg = upo_noise_default((uint32_t)seed);
		cfg.frequency = 0.01f;
		cfg.distance = UPO_DIST_EUCLIDEAN_SQ;
		cfg.operation = UPO_OP_DISTANCE;
		upo_noise_finish(&cfg);
	}
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
	float GetNoise(float x, float y) const { return upo_noise2_supported(&cfg) ? upo_noise2(&cfg, x, y) : 0.0f; }
	float GetNoise(float x, float y, float z) const { return upo_noise3(&cfg, x, y, z); }

private:
	template<typename T>
	struct Lookup
	{
		static const T RandVecs3D[1024];
	};
	upo_noise_cfg cfg;
};

// entry e = (cos, sin, 0, 0) of angle e: recognisable, unit length, not the substitute table
template<typename T>
const T FastNoiseLite::Lookup<T>::RandVecs3D[1024] = {
#define M4(e) (T)__builtin_cos((double)(e)), (T)__builtin_sin((double)(e)), (T)0, (T)0
#define M16(e) M4(e), M4(e + 1), M4(e + 2), M4(e + 3)
#define M64(e) M16(e), M16(e + 4), M16(e + 8), M16(e + 12)
	M64(0), M64(16), M64(32), M64(48), M64(64), M64(80), M64(96), M64(112), M64(128), M64(144), M64(160), M64(176), M64(192), M64(208), M64(224), M64(240)
#undef M64
#undef M16
#undef M4
};

#endif
