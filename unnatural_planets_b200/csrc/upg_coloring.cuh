// upg_coloring.cuh -- the per-tile colouring pipeline of the texel-baking stage (K6).
//
// Replaces terrainTile -> coloringDefault / coloringBarren / coloringDebug and the layer functions
// they run (sources/modesConfiguration.cpp:281-286; sources/coloringModes.cpp:23-966): one `Tile`
// (sources/planets.h:63-83) in, albedo / roughness / metallic / height / opacity / biome / type out.
// Every expression keeps the reference's operation order and literal widths (a `double` literal next
// to a Real is narrowed to float first; `interpolate(1.0, 0.8, f)` is the one double-typed overload).
//
// What is NOT computed here: tile.elevation and the eight terrainSdfElevation probes of generateSlope
// (coloringModes.cpp:122-151).  They depend on the base shape (and, for water / navigation tiles, on
// the elevation mode), so they are evaluated by K1's explicit-position mode (upg_field.cuh) over the
// positions `slope_probe` lists, and handed to the colouring as nine numbers.  That keeps this
// kernel independent of shape and elevation mode: 3 colourings x 2 purposes instead of 34 x 6 x 3 x 2.
//
// Deviation from the reference, declared: colorDeviation() in generateSand (coloringModes.cpp:456)
// draws three numbers from the calling THREAD's random generator, so the reference's own sand texels
// differ from run to run.  Here the three draws come from a xoroshiro128** generator
// whose state is a hash of the tile position (texel_rng_state) -- deterministic, order independent.
#pragma once

#include "upg_color.cuh"
#include "upg_planet.h"
#include "upg_terrain.cuh"

// All warps of a colouring block run the same layer at the same time (a barrier between layers): the instruction stream of a
// tile walks through ~180 KB of code, 6x the L1.5 instruction cache, and warps that drift apart each pay their own fetches
// from L2 (ncu: `no_instruction` was the dominant stall, profiles/r01_d).  In step, one fetch feeds every warp of the SM.
#if defined(__CUDA_ARCH__)
#define UPG_LAYER_SYNC() __syncthreads()
#else
#define UPG_LAYER_SYNC() ((void)0)
#endif

// A layer's early return (`if (bf < 1e-7) return;`) taken by SOME lanes of a warp saves nothing -- the warp walks the rest of the layer
// anyway -- but leaves it with a partial mask, and the warp-cooperative noise paths (upg_noise.cuh: single_cubic) need whole warps.
// The layers below therefore keep a per-lane `on` flag, leave early only when NO lane of the warp is on (UPG_WARP_ANY, a vote), let the
// off lanes compute values nobody reads, and apply their result under `if (on)`.  On the host the vote is the flag itself: the
// reference's control flow.
#if defined(__CUDA_ARCH__)
#define UPG_WARP_ANY(flag) (__any_sync(__activemask(), (flag)) != 0)
#else
#define UPG_WARP_ANY(flag) (flag)
#endif

#if defined(__CUDACC__)
// the layers are real (non-inlined) device functions: the colouring kernel stays a few thousand instructions per layer
// instead of one giant inlined body; they only exist in the translation unit that launches the kernel
#define UPG_LAYER static __device__ __noinline__
#define UPG_COLORING __device__ __forceinline__
#else
#define UPG_LAYER inline
#define UPG_COLORING inline
#endif

namespace upg
{
	// sources/planets.h:18-56
	enum Biome : int32_t
	{
		BI_BARE = 0,
		BI_TUNDRA,
		BI_TAIGA,
		BI_SHRUBLAND,
		BI_GRASSLAND,
		BI_TEMPERATE_SEASONAL_FOREST,
		BI_TEMPERATE_RAIN_FOREST,
		BI_DESERT,
		BI_SAVANNA,
		BI_TROPICAL_SEASONAL_FOREST,
		BI_TROPICAL_RAIN_FOREST,
		BI_WATER,
		BI_TOTAL
	};
	enum TerrainType : int32_t
	{
		TT_ROAD = 0,
		TT_FLAT = 1,
		TT_ROUGH = 2,
		TT_CLIFFS = 3,
		TT_WATER = 4,
		TT_TOTAL
	};
	enum MeshPurpose : int32_t
	{
		MP_UNDEFINED = 0,
		MP_LAND = 1,
		MP_WATER = 2,
		MP_NAVIGATION = 3
	};

	// every noise / voronoi instance of coloringModes.cpp, named <layer><static>; seeds are assigned by
	// PlanetBuilder::coloringStatics in static-initialisation order
	struct ColoringConsts
	{
		NoiseInst beach; // :36-43  Value
		NoiseInst precp; // :61-70  Cubic FBm 2
		NoiseInst temp; // :84-94  Simplex FBm 2
		NoiseInst polar; // :95-104 Value FBm 3
		NoiseInst bedrockDepth, bedrockFreq, bedrockCracks, bedrockValue, bedrockSaturation; // :204-250
		NoiseInst cliffsCracks; // :265-274 SimplexReduced Ridged 2
		NoiseInst micaMask, micaCracks; // :297-317
		NoiseInst dirtHeight, dirtCracks, dirtCracksMask; // :340-371
		NoiseInst sandHeightScale, sandHeight, sandColor, sandHue; // :406-444
		NoiseInst grassHue, grassPatchesCells, grassPatchesDist, grassScratches; // :469-509
		NoiseInst boulderThreshold, boulderSize, boulderHue, boulderValue, boulderScratches; // :537-584
		NoiseInst flowerSize, flowerColor, flowerHue1, flowerHue2; // :674-705
		NoiseInst iceTempOffset, iceFreq, iceThickness, iceCracks; // :749-790
		NoiseInst snowElevOffset, snowThickness; // :821-838
		NoiseInst waterColorSwitch, waterHueShift; // :619-638
		VoronoiInst boulderCenters; // :545-552 cell 200, 2 points
		VoronoiInst flowerCluster, flowerCenter; // :660-673 cell 200 / 35
		// colorRgbToHsluv of the constant end points of interpolateColor, evaluated once on the host by the same
		// double-precision routine (upg_color.cuh)
		V3 hsluvSandTarget; // (187, 104, 50) / 255              :455
		V3 hsluvWaterDeep, hsluvWaterShallow; // (40, 63, 168) / 255, (39, 154, 186) / 255   :643
		V3 hsluvFlowerA, hsluvFlowerB; // (0.5, 0, 0.35), (0.5, 0.4, 0)   :727
		int32_t poles, flowers;
	};

	struct Tile
	{
		V3 position, normal, albedo;
		float roughness, metallic, height, elevation, slope, temperature, precipitation, opacity;
		int32_t biome, type, purpose;
	};

	UPG_HD Tile tile_init(V3 position, V3 normal, int purpose)
	{
		Tile t;
		t.position = position;
		t.normal = normal;
		t.albedo = v3(0, 0, 0);
		t.roughness = t.metallic = t.height = t.elevation = t.slope = t.temperature = t.precipitation = 0;
		t.opacity = 1;
		t.biome = BI_TOTAL;
		t.type = TT_TOTAL;
		t.purpose = purpose;
		return t;
	}

	// ---- sources/math.cpp -----------------------------------------------------------------------
	UPG_HD float rescale_(float v, float ia, float ib, float oa, float ob) { return (v - ia) / (ib - ia) * (ob - oa) + oa; }
	UPG_HD float sharp_edge(float v, float p = 0.05f) { return rescale_(clamp_(v, 0.5f - p, 0.5f + p), 0.5f - p, 0.5f + p, 0, 1); }
	UPG_HD V3 any_perpendicular(V3 a)
	{
		V3 b = v3(1, 0, 0);
		if (abs_(dot(a, b)) > 0.9f)
			b = v3(0, 1, 0);
		return normalize(cross(a, b));
	}
	UPG_HD V3 lerp3(V3 a, V3 b, float f) { return a + (b - a) * f; }
	// interpolate(double, double, Real) (include/cage-core/math.h:1172-1178), narrowed where it meets a Real
	UPG_HD float lerp_double(double a, double b, float f) { return (float)(a + (b - a) * (double)f); }

	// Degs(Rads) / Rads(Degs) (math.h:781-782)
	UPG_HD float rads_to_degs(float r) { return r * 180 / 3.14159274101257324f; }
	UPG_HD float degs_to_rads_(float d) { return d * 3.14159274101257324f / 180; }

	// coloringModes.cpp:23-34
	UPG_HD float steepness_mask(float slopeRads, float thresholdDegs, float smoothingDegs)
	{
		const float slope = rads_to_degs(slopeRads);
		const float r = (thresholdDegs + smoothingDegs - slope) / (2 * smoothingDegs);
		return saturate_(r);
	}
	UPG_HD float range_mask(float value, float zeroAt, float oneAt) { return saturate_((value - zeroAt) / (oneAt - zeroAt)); }

	// the eight probe offsets of generateSlope (:122-133), k = 0..7; probe 8 is the tile position itself
	UPG_HD V3 slope_probe(V3 position, V3 normal, int k)
	{
		const float radius = 0.5f;
		V3 a = any_perpendicular(normal);
		V3 b = cross(normal, a);
		a = a * radius;
		b = b * radius;
		const float div = (float)(1 / 1.4142135623730951); // Real div = 1 / sqrt(2): cage's sqrt(int) returns double
		const V3 c = (a + b) * div;
		const V3 d = (a - b) * div;
		switch (k)
		{
			case 0:
				return position + a;
			case 1:
				return position + b;
			case 2:
				return position - a;
			case 3:
				return position - b;
			case 4:
				return position + c;
			case 5:
				return position + d;
			case 6:
				return position - c;
			case 7:
				return position - d;
			default:
				return position;
		}
	}

	// generateSlope (:134-151) over the eight terrainSdfElevation values
	UPG_HD float slope_from_probes(const float *elevs)
	{
		float e1 = elevs[0], e2 = elevs[0];
		for (int i = 0; i < 8; i++)
		{
			e1 = fmin_(e1, elevs[i]);
			e2 = fmax_(e2, elevs[i]);
		}
		const float md = (e2 - e1) * 0.1f;
		return atanf(md / 0.5f);
	}

	// ---- the stand-in for the thread RNG behind colorDeviation (see the header comment) ----------------------
	UPG_HD uint64_t splitmix64_(uint64_t z)
	{
		z += 0x9e3779b97f4a7c15ull;
		z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
		z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
		return z ^ (z >> 31);
	}
	UPG_HD uint32_t float_bits(float f)
	{
#if defined(__CUDA_ARCH__)
		return __float_as_uint(f);
#else
		uint32_t u;
		memcpy(&u, &f, 4);
		return u;
#endif
	}
	UPG_HD void texel_rng_state(V3 p, uint64_t *s)
	{
		const uint64_t k = (uint64_t)float_bits(p.x) | ((uint64_t)float_bits(p.y) << 32);
		s[0] = splitmix64_(k) | 1ull;
		s[1] = splitmix64_(s[0] ^ (uint64_t)float_bits(p.z)) | 1ull;
	}
	// RandomGenerator::next / randomChance (externals/cage/sources/libcore/math/random.cpp:27-51)
	UPG_HD float xoroshiro_chance(uint64_t *s)
	{
		const uint64_t s0 = s[0];
		uint64_t s1 = s[1];
		const uint64_t t = s0 * 5;
		const uint64_t result = ((t << 7) | (t >> 57)) * 9;
		s1 ^= s0;
		s[0] = ((s0 << 24) | (s0 >> 40)) ^ s1 ^ (s1 << 16);
		s[1] = (s1 << 37) | (s1 >> 27);
		return (float)((double)(result >> 11) * 0x1.0p-53);
	}

	// nearest projected feature point, as a real function (three call sites)
#if defined(__CUDACC__)
	static __device__ __noinline__ V3 voronoi_nearest1(const VoronoiInst &v, V3 position, V3 normal)
#else
	inline V3 voronoi_nearest1(const VoronoiInst &v, V3 position, V3 normal)
#endif
	{
		V3 r;
		voronoi_nearest<1, true>(v, position, normal, &r);
		return r;
	}

	// ---- layers -------------------------------------------------------------------------------------------------
	// beachMask (:34-49)
	UPG_HD float beach_mask(const ColoringConsts &C, const Tile &t)
	{
		if (t.biome == BI_WATER)
			return 0;
		return range_mask(t.elevation + noise1_value(C.beach, t.elevation) * 20, 10, 15);
	}

	// generatePrecipitation (:59-80)
	UPG_LAYER void generate_precipitation(const ColoringConsts &C, Tile &t)
	{
		float p = noise3<NT_CUBIC, FT_FBM, 2, true>(C.precp, t.position) * 0.5f + 0.5f;
		p = saturate_(p);
		p = smootherstep_(p);
		p = smootherstep_(p);
		p = powf(fmax_(p + 0.2f, 0), 3) * 0.6f;
		p *= 400;
		p += fmax_(120 - abs_(t.elevation), 0) * 0.2f;
		p = fmax_(p, 0);
		t.precipitation = p;
	}

	// generateTemperature (:82-120)
	UPG_LAYER void generate_temperature(const ColoringConsts &C, Tile &t)
	{
		float v = noise3<NT_SIMPLEX, FT_FBM, 2, true>(C.temp, t.position) * 0.5f + 0.5f;
		v = saturate_(v);
		v = smoothstep_(v);
		v = v * 2 - 1;
		if (C.poles)
		{
			float polar = abs_(atanf(t.position.y / length(v2(t.position.x, t.position.z)))) / 3.14159274101257324f * 2;
			polar = powf(polar, 1.7f);
			polar += noise3<NT_VALUE, FT_FBM, 3, true>(C.polar, t.position) * 0.1f;
			v += 0.6f - polar * 3.2f;
		}
		t.temperature = 15 + v * 30 - fmax_(t.elevation, 0) * 0.02f;
	}

	// generateBiome (:153-190)
	UPG_HD void generate_biome(Tile &t)
	{
		if (t.elevation < 0)
			t.biome = BI_WATER;
		else if (t.temperature > 20)
		{
			if (t.precipitation < 40)
				t.biome = BI_DESERT;
			else if (t.precipitation < 130)
				t.biome = BI_SAVANNA;
			else if (t.precipitation < 230)
				t.biome = BI_TROPICAL_SEASONAL_FOREST;
			else
				t.biome = BI_TROPICAL_RAIN_FOREST;
		}
		else if (t.temperature > 5)
		{
			if (t.precipitation < 30)
				t.biome = BI_SHRUBLAND;
			else if (t.precipitation < 100)
				t.biome = BI_GRASSLAND;
			else if (t.precipitation < 200)
				t.biome = BI_TEMPERATE_SEASONAL_FOREST;
			else
				t.biome = BI_TEMPERATE_RAIN_FOREST;
		}
		else if (t.temperature > -5)
		{
			if (t.precipitation < 30)
				t.biome = BI_BARE;
			else
				t.biome = BI_TAIGA;
		}
		else if (t.precipitation < 10)
			t.biome = BI_BARE;
		else
			t.biome = BI_TUNDRA;
	}

	// generateType (:192-200); `slope > Degs(20)` compares in radians (math.h:378-381)
	UPG_HD void generate_type(Tile &t)
	{
		if (t.elevation < 0)
			t.type = TT_WATER;
		else if (t.slope > degs_to_rads_(20))
			t.type = TT_CLIFFS;
		else
			t.type = TT_FLAT;
	}

	// generateBedrock (:202-261)
	UPG_LAYER void generate_bedrock(const ColoringConsts &C, const float *rv, Tile &t)
	{
		const float depth = sqr_(noise3<NT_CUBIC, FT_NONE, 3, true>(C.bedrockDepth, t.position) * 0.5f + 0.501f) * 2;
		const float freq = noise3<NT_CUBIC, FT_NONE, 3, true>(C.bedrockFreq, t.position) * 0.15f + 1;
		const V3 pf = t.position * freq;
		// cracksNoise and valueNoise share their seed and frequency (:221-240): one scan of the 27 cells serves both (noise3_cellular_pair)
		float cracksNoise, value;
		noise3_cellular_pair<CD_HYBRID, CO_SUBTRACT, CO_CELL, FT_NONE, 3>(rv, C.bedrockCracks, pf.x, pf.y, pf.z, cracksNoise, value);
		const float cracks = saturate_(powf(cracksNoise * 0.5f + 0.5f, 0.7f));
		const float saturation = saturate_(noise3<NT_CUBIC, FT_FBM, 2, true>(C.bedrockSaturation, t.position) * 0.35f + 0.3f);
		const V3 hsv = v3(0.07f, saturation, saturate_(value * 0.1f + 0.7f) * (cracks * 0.4f + 0.6f));
		t.albedo = hsv_to_rgb(hsv);
		t.roughness = lerp_(0.9f, value * -0.15f + 0.65f, cracks);
		t.height = cracks * depth;
	}

	// generateCliffs (:263-293)
	UPG_LAYER void generate_cliffs(const ColoringConsts &C, Tile &t)
	{
		const float bf = steepness_mask(t.slope, 19, 4);
		if (bf > 0.99999f)
			return;
		V3 hsv = rgb_to_hsv(t.albedo);
		hsv.x = lerp_((float)(155.0 / 255.0), hsv.x, sharp_edge(bf, 0.005f));
		hsv.y *= bf * 0.4f + 0.6f;
		hsv.z = lerp_(0.7f, hsv.y, bf * 0.8f + 0.2f);
		t.albedo = hsv_to_rgb(hsv);
		t.roughness *= bf * 0.2f + 0.8f;
		{
			const float cracks = sharp_edge(saturate_(noise3<NT_SIMPLEX_REDUCED, FT_RIDGED, 2, true>(C.cliffsCracks, t.position) * 2 - 1.1f), 0.2f);
			t.height *= lerp_double(1.0, 0.8, cracks);
			t.albedo = t.albedo * lerp_double(1.0, 0.9, cracks);
			t.roughness = lerp_(t.roughness, 0.9f, cracks);
		}
	}

	// generateMica (:295-336)
	UPG_LAYER void generate_mica(const ColoringConsts &C, const float *rv, Tile &t)
	{
		if (!C.flowers)
			return;
		const float bf = saturate_((noise3_cellular<CD_HYBRID, CO_MULTIPLY, FT_NONE, 3>(rv, C.micaMask, t.position.x, t.position.y, t.position.z) - 0.98f) * 10);
		if (bf < 1e-7f)
			return;
		const float cracks = sharp_edge(saturate_((noise3_cellular<CD_MANHATTAN, CO_DIVIDE, FT_NONE, 3>(rv, C.micaCracks, t.position.x, t.position.y, t.position.z) + 0.6f)));
		const V3 color = lerp3(v3(122, 90, 88) / 255, v3(184, 209, 187) / 255, cracks);
		const float roughness = cracks * 0.3f + 0.4f;
		const float metallic = 1;
		const float height = t.height * 0.5f;
		t.albedo = lerp3(t.albedo, color, bf);
		t.roughness = lerp_(t.roughness, roughness, bf);
		t.metallic = lerp_(t.metallic, metallic, bf);
		t.height = lerp_(t.height, height, bf);
	}

	// generateDirt (:338-402)
	UPG_LAYER void generate_dirt(const ColoringConsts &C, Tile &t)
	{
		// bf = (a factor in [0, 1]) * steepness: when the steepness mask alone is under the threshold so is bf, and the three Perlin octaves
		// behind the other factor are not needed (same early return as the reference's, taken earlier)
		const float steepness = steepness_mask(t.slope, 20, 5);
		bool on = !(steepness < 1e-7f);
		if (!UPG_WARP_ANY(on))
			return;
		float height = noise3<NT_PERLIN, FT_FBM, 3, true>(C.dirtHeight, t.position) * 0.2f + 0.5f;
		const float bf = sharp_edge(saturate_(height - t.height + 0.4f)) * steepness;
		on = on && !(bf < 1e-7f);
		if (!UPG_WARP_ANY(on))
			return;
		V3 color = v3(168, 94, 28) / 255;
		{
			const float saturation = range_mask(t.precipitation, 0, 50);
			V3 hsv = rgb_to_hsv(color);
			hsv.y *= saturation * 0.7f;
			hsv.z *= 0.8f;
			color = hsv_to_rgb(hsv);
		}
		float roughness = 0.8f;
		const float metallic = 0;
		{
			float cracks = sharp_edge(saturate_(noise3<NT_SIMPLEX_REDUCED, FT_RIDGED, 2, true>(C.dirtCracks, t.position) * 2 - 1.2f), 0.15f);
			cracks *= sqr_(smoothstep_(saturate_(noise3<NT_CUBIC, FT_FBM, 4, true>(C.dirtCracksMask, t.position) * 0.5f + 0.5f))) * 0.9f + 0.1f;
			cracks *= range_mask(t.precipitation, 70, 20) * 0.9f + 0.1f;
			height = lerp_(height, height * 0.5f, cracks);
			color = lerp3(color, v3(0.1f, 0.1f, 0.1f), cracks);
			roughness = lerp_(roughness, 0.9f, cracks);
		}
		if (!on)
			return;
		t.albedo = lerp3(t.albedo, color, bf);
		t.roughness = lerp_(t.roughness, roughness, bf);
		t.metallic = lerp_(t.metallic, metallic, bf);
		t.height = lerp_(t.height, height, bf);
	}

	// generateSand (:404-465)
	UPG_LAYER void generate_sand(const ColoringConsts &C, Tile &t)
	{
		const float bf = range_mask(t.temperature, 24, 28) * steepness_mask(t.slope, 19, 10);
		const bool on = !(bf < 1e-7f);
		if (!UPG_WARP_ANY(on))
			return;
		const float heightScale = noise3<NT_SIMPLEX, FT_FBM, 3, true>(C.sandHeightScale, t.position) * 0.3f + 1;
		const float height = (noise3<NT_SIMPLEX, FT_RIDGED, 3, true>(C.sandHeight, t.position * heightScale) * 0.2f) * (range_mask(t.precipitation, 100, 50) * 0.4f + 0.6f) + 0.5f;
		const float colorShift = smootherstep_(smootherstep_(noise3<NT_PERLIN, FT_NONE, 3, true>(C.sandColor, t.position) * 0.5f + 0.5f));
		const float hueShift = noise3<NT_CUBIC, FT_FBM, 2, true>(C.sandHue, t.position) * 0.1f;
		V3 color = color_hue_shift(v3(189, 174, 152) / 255, hueShift);
		color = interpolate_hsluv(rgb_to_hsluv(color), C.hsluvSandTarget, colorShift);
		{
			uint64_t s[2];
			texel_rng_state(t.position, s);
			const float r1 = xoroshiro_chance(s), r2 = xoroshiro_chance(s), r3 = xoroshiro_chance(s);
			color = color_deviation(color, 0.08f, v3(r1, r2, r3));
		}
		color = hsv_to_rgb(rgb_to_hsv(color) * v3(1, 0.7f, 0.8f));
		const float roughness = (heightScale - 1) * 0.7f + 0.55f + hueShift;
		const float metallic = 0;
		if (!on)
			return;
		t.albedo = lerp3(t.albedo, color, bf);
		t.roughness = lerp_(t.roughness, roughness, bf);
		t.metallic = lerp_(t.metallic, metallic, bf);
		t.height = lerp_(t.height, height, bf);
	}

	// generateGrass (:467-533)
	UPG_LAYER void generate_grass(const ColoringConsts &C, const float *rv, Tile &t)
	{
		const float bf = range_mask(t.temperature, 35, 25) * range_mask(t.precipitation, 15, 35) * steepness_mask(t.slope, 22, 3) * beach_mask(C, t);
		if (bf < 1e-7f)
			return;
		const float dryness = clamp_(t.temperature - (t.precipitation + 100) * 30 / 400, 0, 5) / 5;
		const float hueShiftBase = noise3<NT_PERLIN, FT_FBM, 2, true>(C.grassHue, t.position);
		const float hueShift = hueShiftBase * 0.09f - dryness * 0.1f;
		V3 color = v3(114, 187, 77) / 255;
		color = color_hue_shift(color, hueShift);
		const V3 p = t.position;
		float patchCell, patchDist; // one seed, one frequency (:480-509): one scan
		noise3_cellular_pair<CD_HYBRID, CO_CELL, CO_SUBTRACT, FT_NONE, 3>(rv, C.grassPatchesCells, p.x, p.y, p.z, patchCell, patchDist);
		const float patchesMask = sharp_edge(patchCell * 0.5f + 0.2f);
		const float patches = powf(patchDist * 0.5f + 0.5f, 0.5f) * patchesMask * (1 - dryness * 0.8f);
		color = color * (1 - patches * 0.09f);
		const float scratches = saturate_(noise3_cellular<CD_HYBRID, CO_MULTIPLY, FT_NONE, 3>(rv, C.grassScratches, p.x, p.y, p.z) - 0.2f) * (1 - dryness * 0.5f);
		color = color * (1 - scratches * 0.1f);
		const float roughness = 0.6f + dryness * 0.1f + patches * 0.1f + scratches * 0.1f - hueShiftBase * 0.1f;
		const float metallic = 0;
		const float height = lerp_(t.height, 0.5f, 0.7f) + (0.5f - dryness) * 0.2f + patches * 0.05f - scratches * 0.02f + hueShiftBase * 0.1f + 0.1f;
		t.albedo = lerp3(t.albedo, color, bf);
		t.roughness = lerp_(t.roughness, roughness, bf);
		t.metallic = lerp_(t.metallic, metallic, bf);
		t.height = lerp_(t.height, height, bf);
	}

	// generateBoulders (:535-615)
	UPG_LAYER void generate_boulders(const ColoringConsts &C, Tile &t)
	{
		if (!C.flowers)
			return;
		if (noise3<NT_PERLIN, FT_NONE, 3, true>(C.boulderThreshold, t.position) < 0.15f)
			return;
		const V3 center = voronoi_nearest1(C.boulderCenters, t.position, t.normal);
		const float dist = distance(center, t.position);
		const float size = smootherstep_(smootherstep_(saturate_(noise3<NT_PERLIN, FT_NONE, 3, true>(C.boulderSize, t.position) * 0.5f + 0.5f))) * 3 + 3;
		const float bf = range_mask(size - dist, 0, 0.5f);
		if (bf < 1e-7f)
			return;
		const float hueShift = noise3<NT_PERLIN, FT_NONE, 3, true>(C.boulderHue, t.position) * 0.12f;
		const float valueShift = noise3<NT_PERLIN, FT_NONE, 3, true>(C.boulderValue, t.position) * 0.18f;
		V3 color = rgb_to_hsv(v3(154, 154, 154) / 255);
		color.x = fmod_(color.x + hueShift + 1, 1.0f);
		color.z = saturate_(color.z + valueShift);
		color = hsv_to_rgb(color);
		const float roughness = noise3<NT_PERLIN, FT_NONE, 3, true>(C.boulderScratches, t.position) * 0.1f + 0.4f;
		const float metallic = 0;
		const float height = 1 - sqr_(dist / size) * 0.5f;
		t.albedo = lerp3(t.albedo, color, bf);
		t.roughness = lerp_(t.roughness, roughness, bf);
		t.metallic = lerp_(t.metallic, metallic, bf);
		t.height = lerp_(t.height, height, bf);
	}

	// generateWater (:617-656)
	UPG_LAYER void generate_water(const ColoringConsts &C, Tile &t)
	{
		{
			const float colorSwitch = noise3<NT_PERLIN, FT_FBM, 2, true>(C.waterColorSwitch, t.position) * 0.5f + 0.5f;
			const float hueShift = noise3<NT_CUBIC, FT_FBM, 3, true>(C.waterHueShift, t.position) * 0.09f;
			t.albedo = color_hue_shift(interpolate_hsluv(C.hsluvWaterDeep, C.hsluvWaterShallow, colorSwitch), hueShift);
		}
		{
			const float d1 = 1 - sqr_(range_mask(t.elevation, -25, 3));
			const float d2 = rescale_(range_mask(t.elevation, 0, -100), 0, 1, 0.7f, 1);
			t.opacity = d1 * d2;
		}
		{
			const float deep = smoothstep_(range_mask(t.elevation, 3, -30));
			t.metallic = deep;
		}
	}

	// generateFlowers (:658-744)
	UPG_LAYER void generate_flowers(const ColoringConsts &C, Tile &t)
	{
		if (!C.flowers)
			return;
		const bool waterlily = t.purpose != MP_LAND;
		if ((t.biome == BI_WATER) != waterlily)
			return;
		if (range_mask(abs_(t.temperature - 18), 12, 7) < 1e-5f)
			return;
		if (!waterlily && range_mask(abs_(t.precipitation - 200), 100, 70) < 1e-5f)
			return;
		const V3 center = voronoi_nearest1(C.flowerCenter, t.position, t.normal);
		if (distance_squared(center, t.position) > sqr_(4.0f))
			return;
		const V3 cluster = voronoi_nearest1(C.flowerCluster, t.position, t.normal);
		if (distance_squared(center, cluster) > sqr_(20.0f))
			return;
		const float dist = distance(center, t.position);
		const float size = smootherstep_(smootherstep_(saturate_(noise3<NT_PERLIN, FT_NONE, 3, true>(C.flowerSize, t.position) * 0.5f + 0.5f))) * 2 + 1.5f;
		const float bf = range_mask(size - dist, 0, 0.1f);
		if (bf < 1e-7f)
			return;
		const V3 baseColor = waterlily ? v3(0.2f, 0.14f, 0) : interpolate_hsluv(C.hsluvFlowerA, C.hsluvFlowerB, sharp_edge(noise3<NT_PERLIN, FT_NONE, 3, true>(C.flowerColor, cluster) + 0.4f, 0.2f));
		const V3 color = color_hue_shift(baseColor, noise3<NT_PERLIN, FT_NONE, 3, true>(C.flowerHue1, t.position) * 0.1f + noise3<NT_PERLIN, FT_NONE, 3, true>(C.flowerHue2, t.position) * 0.1f);
		const float roughness = 0.8f;
		const float metallic = waterlily ? 0.85f : 0;
		const float height = 0.7f + sqr_(dist / size) * 0.2f;
		t.albedo = lerp3(t.albedo, color, bf);
		t.roughness = lerp_(t.roughness, roughness, bf);
		t.metallic = lerp_(t.metallic, metallic, bf);
		t.height = lerp_(t.height, height, bf);
		t.opacity = lerp_(t.opacity, 1, bf);
	}

	// generateIce (:746-817)
	UPG_LAYER void generate_ice(const ColoringConsts &C, const float *rv, Tile &t)
	{
		const float tempOff = noise3<NT_VALUE, FT_NONE, 3, true>(C.iceTempOffset, t.position) * 1.5f;
		const float bf = sharp_edge(range_mask(t.temperature + tempOff, 0, -3)) * (1 - beach_mask(C, t));
		if (bf < 1e-7f)
			return;
		if (bf > 0.1f)
		{
			if (t.type != TT_CLIFFS)
				t.type = TT_ROUGH;
		}
		const float freq = noise3<NT_VALUE, FT_FBM, 3, true>(C.iceFreq, t.position) * 0.08f + 1;
		const V3 pf = t.position * freq;
		float thicknessNoise, crackNoise; // one seed, one frequency, two octaves each (:767-790): one scan per octave
		noise3_cellular_pair<CD_HYBRID, CO_CELL, CO_SUBTRACT, FT_FBM, 2>(rv, C.iceThickness, pf.x, pf.y, pf.z, thicknessNoise, crackNoise);
		const float thickness = sharp_edge(saturate_(thicknessNoise * 0.5f + 0.8f), 0.1f) * 0.6f + 0.4f;
		const float crack = 1 - powf(crackNoise * 0.5f + 0.5f, 0.3f);
		V3 color = max3(v3(122, 162, 164) / 255 - crack * 0.3f, 0);
		color = interpolate_color(t.albedo, color, thickness);
		float roughness = 0.15f + crack * 0.6f;
		roughness = lerp_(t.roughness, roughness, thickness);
		const float height = thickness * 0.5f * (1 - crack * 0.8f);
		t.albedo = lerp3(t.albedo, color, bf);
		t.roughness = lerp_(t.roughness, roughness, bf);
		t.metallic = lerp_(t.metallic, 0, bf);
		t.height = lerp_(t.height, height, bf);
		t.opacity = lerp_(t.opacity, t.opacity * 0.9f + thickness * 0.5f, bf);
	}

	// generateSnow (:819-855)
	UPG_LAYER void generate_snow(const ColoringConsts &C, Tile &t)
	{
		float bf = sharp_edge(range_mask(t.elevation + noise3<NT_VALUE, FT_NONE, 3, true>(C.snowElevOffset, t.position) * 20, 220, 240) * steepness_mask(t.slope, 25, 5));
		// The reference evaluates the thickness noise (three ValueCubic octaves, 192 lattice taps) before it tests bf.  The factor applied
		// below lies in [0, 1], so a bf under the threshold here stays under it: the same early return, taken before the noise -- which is
		// nearly always, snow needs 220 m of elevation (identical tiles: tests/test_host_coloring.py).  A NaN bf compares false and goes on.
		bool on = !(bf < 1e-7f);
		if (!UPG_WARP_ANY(on))
			return;
		const float thickness = noise3<NT_CUBIC, FT_FBM, 3, true>(C.snowThickness, t.position) * 0.5f + 0.5f;
		bf *= saturate_(thickness * 0.5f + 0.7f);
		on = on && !(bf < 1e-7f);
		if (!on)
			return;
		const V3 color = v3(230, 230, 230) / 255;
		const float roughness = thickness * 0.3f + 0.2f;
		const float metallic = 0;
		const float height = 0.6f + t.height * 0.1f + thickness * 0.3f;
		t.albedo = lerp3(t.albedo, color, bf);
		t.roughness = lerp_(t.roughness, roughness, bf);
		t.metallic = lerp_(t.metallic, metallic, bf);
		t.height = lerp_(t.height, height, bf);
	}

	// generateVisualization (:857-893)
	UPG_HD void generate_visualization(Tile &t)
	{
		switch (t.biome)
		{
			case BI_BARE:
				t.albedo = v3(187, 187, 187) / 255;
				break;
			case BI_TUNDRA:
				t.albedo = v3(175, 226, 255) / 255;
				break;
			case BI_TAIGA:
				t.albedo = v3(193, 225, 221) / 255;
				break;
			case BI_SHRUBLAND:
				t.albedo = v3(200, 110, 62) / 255;
				break;
			case BI_GRASSLAND:
				t.albedo = v3(239, 252, 143) / 255;
				break;
			case BI_TEMPERATE_SEASONAL_FOREST:
				t.albedo = v3(151, 183, 102) / 255;
				break;
			case BI_TEMPERATE_RAIN_FOREST:
				t.albedo = v3(118, 169, 92) / 255;
				break;
			case BI_DESERT:
				t.albedo = v3(221, 187, 76) / 255;
				break;
			case BI_SAVANNA:
				t.albedo = v3(253, 213, 120) / 255;
				break;
			case BI_TROPICAL_SEASONAL_FOREST:
				t.albedo = v3(161, 151, 0) / 255;
				break;
			case BI_TROPICAL_RAIN_FOREST:
				t.albedo = v3(50, 122, 30) / 255;
				break;
			case BI_WATER:
				t.albedo = v3(0, 0, 255) / 255;
				break;
			default:
				t.albedo = v3(NAN, NAN, NAN);
				break;
		}
		t.roughness = 0.7f;
	}

	// generateFinalization (:895-902)
	UPG_HD void generate_finalization(Tile &t)
	{
		t.albedo = v3(saturate_(t.albedo.x), saturate_(t.albedo.y), saturate_(t.albedo.z));
		t.roughness = saturate_(t.roughness);
		t.metallic = saturate_(t.metallic);
		t.height = saturate_(t.height);
		t.opacity = saturate_(t.opacity);
	}

	// coloringDefault / coloringBarren / coloringDebug (:905-966).  t.elevation and t.slope are already set.
	template<int COLORING>
	UPG_COLORING void coloring_run(const ColoringConsts &C, const float *rv, Tile &t)
	{
		// t.purpose is uniform over a launch, so every thread of a block meets the same sequence of barriers
		if (COLORING == CL_DEFAULT)
		{
			generate_precipitation(C, t);
			UPG_LAYER_SYNC();
			generate_temperature(C, t);
			generate_biome(t);
			generate_type(t);
			UPG_LAYER_SYNC();
			if (t.purpose == MP_WATER)
			{
				generate_water(C, t);
			}
			else
			{
				generate_bedrock(C, rv, t);
				UPG_LAYER_SYNC();
				generate_cliffs(C, t);
				UPG_LAYER_SYNC();
				generate_mica(C, rv, t);
				UPG_LAYER_SYNC();
				generate_dirt(C, t);
				UPG_LAYER_SYNC();
				generate_sand(C, t);
				UPG_LAYER_SYNC();
				generate_grass(C, rv, t);
				UPG_LAYER_SYNC();
				generate_boulders(C, t);
			}
			UPG_LAYER_SYNC();
			generate_flowers(C, t);
			UPG_LAYER_SYNC();
			generate_ice(C, rv, t);
			UPG_LAYER_SYNC();
			if (t.purpose != MP_WATER)
				generate_snow(C, t);
			generate_finalization(t);
		}
		else if (COLORING == CL_BARREN)
		{
			if (t.purpose == MP_WATER)
			{
				t.biome = BI_WATER;
				t.type = TT_WATER;
				generate_water(C, t);
			}
			else
			{
				t.biome = BI_BARE;
				t.type = TT_FLAT;
				generate_bedrock(C, rv, t);
				UPG_LAYER_SYNC();
				generate_cliffs(C, t);
				UPG_LAYER_SYNC();
				generate_mica(C, rv, t);
				UPG_LAYER_SYNC();
				generate_boulders(C, t);
			}
			generate_finalization(t);
		}
		else
		{
			generate_precipitation(C, t);
			UPG_LAYER_SYNC();
			generate_temperature(C, t);
			generate_biome(t);
			generate_type(t);
			generate_visualization(t);
			if (t.purpose == MP_WATER)
				t.opacity = 0.5f;
			generate_finalization(t);
		}
	}
}
