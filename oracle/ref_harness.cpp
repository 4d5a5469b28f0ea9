/*
 * ref_harness.cpp -- TEST INFRASTRUCTURE (oracle/_ref).
 *
 * C entry points over the REFERENCE'S OWN translation units, which oracle/Makefile compiles
 * verbatim from /root/reference into oracle/_ref/libup_ref.so (nothing is copied into this
 * repository).  The only restated arithmetic underneath is FastNoiseLite (shims/FastNoiseLite.h).
 * Used to (1) validate the oracle restatement, (2) generate tests/golden/ fixtures
 * (tests/golden/make_golden.py), (3) serve as the "reference" CPU baseline in bench.py.
 * The product never links or loads this.
 *
 * One loaded copy of the library == one planet configuration: the reference keeps its noise
 * objects in function-local statics (e.g. sources/elevationModes.cpp:131-140), initialised once.
 * oracle/ref.py loads a private copy of the .so per configuration.
 */
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

// With -DUPREF_UPSTREAM_FASTNOISE (oracle/Makefile: `make ref FASTNOISE=/path/FastNoiseLite.h`) the include below resolves to the REAL
// Auburn/FastNoiseLite header instead of shims/FastNoiseLite.h: the pin hook for the day the un-vendored submodule is available
// (tools/fastnoise_pin.py diff-reports the restatement against it and extracts the Cellular table).  Its constant tables are private.
#if defined(UPREF_UPSTREAM_FASTNOISE)
#define private public
#include <FastNoiseLite.h>
#undef private
#include "up_noise.h" // upo_noise_cfg of the harness' C interface
#else
#include <FastNoiseLite.h> // the shim (registry of constructed noise objects)
#endif

#include <cage-core/color.h>
#include <cage-core/config.h>
#include <cage-core/geometry.h>
#include <cage-core/image.h>
#include <cage-core/imageAlgorithms.h>
#include <cage-core/marchingCubes.h>
#include <cage-core/mesh.h>
#include <cage-core/meshAlgorithms.h>
#include <cage-core/random.h>

#include <dualmc.h>

#include "planets.h" // -iquote $(REF)/sources

namespace unnatural
{
	// sources/modesConfiguration.cpp:230-315
	Real terrainSdfElevation(const Vec3 &pos);
	Real terrainSdfElevationRaw(const Vec3 &pos);
	Real terrainSdfLand(const Vec3 &pos);
	Real terrainSdfWater(const Vec3 &pos);
	Real terrainSdfNavigation(const Vec3 &pos);
	void terrainTile(Tile &tile);
	void terrainPreseed();
	void terrainApplyConfig();
	// sources/textures.cpp:86-96
	void generateTexturesLand(const Holder<Mesh> &renderMesh, uint32 width, uint32 height, Holder<Image> &albedo, Holder<Image> &special, Holder<Image> &heightMap);
	void generateTexturesWater(const Holder<Mesh> &renderMesh, uint32 width, uint32 height, Holder<Image> &albedo, Holder<Image> &special, Holder<Image> &heightMap);
	// sources/math.cpp:73-77
	uint32 noiseSeed();

	// globals the reference defines in generator.cpp / doodads.cpp etc. (not compiled here)
	std::vector<Tile> tiles;
}

using namespace cage;
using namespace unnatural;

#if !defined(UPREF_UPSTREAM_FASTNOISE)
std::vector<const FastNoiseLite *> &upo_shim_registry()
{
	static std::vector<const FastNoiseLite *> r;
	return r;
}
#endif

namespace
{
	thread_local std::string lastError;

	template<class F>
	int guarded(F &&f)
	{
		try
		{
			f();
			return 0;
		}
		catch (const cage::Exception &e)
		{
			lastError = std::string("cage::Exception: ") + (const char *)e.message;
		}
		catch (const std::exception &e)
		{
			lastError = std::string("std::exception: ") + e.what();
		}
		catch (...)
		{
			lastError = "unknown exception";
		}
		return 1;
	}

	using Sdf = Real (*)(const Vec3 &);
	Sdf pickField(int what)
	{
		switch (what)
		{
			case 0:
				return &terrainSdfWater; // == the base shape
			case 1:
				return &terrainSdfElevationRaw;
			case 2:
				return &terrainSdfLand;
			case 3:
				return &terrainSdfNavigation;
			case 4:
				return &terrainSdfElevation;
		}
		return nullptr;
	}

	// position-keyed RNG state for the colorDeviation() draws (see upref_tiles_seeded)
	uint64_t splitmix64(uint64_t z)
	{
		z += 0x9e3779b97f4a7c15ull;
		z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
		z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
		return z ^ (z >> 31);
	}
	void texelRngState(const float *p, uint64_t *s)
	{
		uint32_t b[3];
		std::memcpy(b, p, 12);
		s[0] = splitmix64((uint64_t)b[0] | ((uint64_t)b[1] << 32)) | 1ull;
		s[1] = splitmix64(s[0] ^ (uint64_t)b[2]) | 1ull;
	}

	MarchingCubesCreateConfig gridConfig(float boxSize, int res)
	{
		// sources/meshGeneration.cpp:34-37
		MarchingCubesCreateConfig cfg;
		cfg.box = Aabb(Vec3(boxSize * -0.5), Vec3(boxSize * 0.5));
		cfg.resolution = Vec3i(res);
		return cfg;
	}
}

extern "C"
{
	const char *upref_last_error()
	{
		return lastError.c_str();
	}

	// configure + preseed exactly as main.cpp:17-54 / generator.cpp:424-431 would, with an explicit RNG state
	int upref_init(const char *shape, const char *elevation, const char *coloring, int poles, int flowers, uint64_t s0, uint64_t s1)
	{
		return guarded(
			[&]()
			{
				configSetString("unnatural-planets/shape/mode", shape);
				configSetString("unnatural-planets/elevation/mode", elevation);
				configSetString("unnatural-planets/coloring/mode", coloring);
				configSetBool("unnatural-planets/poles/enable", !!poles);
				configSetBool("unnatural-planets/flowers/enable", !!flowers);
				detail::randomGenerator() = RandomGenerator(s0, s1);
				terrainApplyConfig();
				terrainPreseed();
			});
	}

	// state of the calling thread's RNG (to check how many draws the preseed consumed)
	void upref_rng_state(uint64_t *s)
	{
		s[0] = detail::randomGenerator().s[0];
		s[1] = detail::randomGenerator().s[1];
	}
	void upref_rng_set(uint64_t s0, uint64_t s1)
	{
		detail::randomGenerator() = RandomGenerator(s0, s1);
	}

	// KAT helpers over cage's RandomGenerator (test-core/random.cpp:37-79)
	void upref_rng_kat(uint64_t s0, uint64_t s1, int n, uint64_t *raw)
	{
		RandomGenerator g(s0, s1);
		for (int i = 0; i < n; i++)
			raw[i] = g.next();
	}
	void upref_rng_range3(uint64_t s0, uint64_t s1, float a, float b, float *out3, int32_t *out3i, int ia, int ib)
	{
		RandomGenerator g(s0, s1);
		const Vec3i vi = g.randomRange3i(ia, ib);
		const Vec3 v = g.randomRange3(a, b);
		for (int i = 0; i < 3; i++)
		{
			out3[i] = v[i].value;
			out3i[i] = vi[i];
		}
	}

#if !defined(UPREF_UPSTREAM_FASTNOISE)
	int upref_noise_count()
	{
		return (int)upo_shim_registry().size();
	}
	void upref_noise_get(int i, upo_noise_cfg *out)
	{
		*out = upo_shim_registry()[i]->cfg;
	}
	// direct access to one registered noise object (3-D and the 1-D overload cage maps to (x, 0))
	void upref_noise_eval(int i, const float *pos, int n, float *out)
	{
		const FastNoiseLite *f = upo_shim_registry()[i];
		for (int k = 0; k < n; k++)
			out[k] = f->GetNoise(pos[k * 3 + 0], pos[k * 3 + 1], pos[k * 3 + 2]);
	}
#else
	// the real library keeps no registry of its objects
	int upref_noise_count() { return 0; }
	void upref_noise_get(int, upo_noise_cfg *) {}
	void upref_noise_eval(int, const float *, int, float *) {}
#endif
	// 1 when this library was built against the real Auburn/FastNoiseLite header, 0 for the restatement (shims/FastNoiseLite.h)
	int upref_fastnoise_is_upstream()
	{
#if defined(UPREF_UPSTREAM_FASTNOISE)
		return 1;
#else
		return 0;
#endif
	}
	// one FastNoiseLite object configured the way cage configures it (libcore/noiseFunction.cpp:97-107), evaluated at n points: the
	// common ground on which the restatement and the real header are compared generator by generator (tools/fastnoise_pin.py)
	void upref_fastnoise_eval(const upo_noise_cfg *c, const float *pos, int n, int dims, float *out)
	{
		FastNoiseLite f(c->seed);
		static const FastNoiseLite::NoiseType types[] = { FastNoiseLite::NoiseType_OpenSimplex2, FastNoiseLite::NoiseType_OpenSimplex2S, FastNoiseLite::NoiseType_Cellular, FastNoiseLite::NoiseType_Perlin, FastNoiseLite::NoiseType_ValueCubic, FastNoiseLite::NoiseType_Value };
		f.SetFrequency(c->frequency);
		f.SetNoiseType(types[c->type]);
		f.SetFractalType((FastNoiseLite::FractalType)c->fractal);
		f.SetFractalOctaves(c->octaves);
		f.SetFractalLacunarity(c->lacunarity);
		f.SetFractalGain(c->gain);
		f.SetCellularDistanceFunction((FastNoiseLite::CellularDistanceFunction)c->distance);
		f.SetCellularReturnType((FastNoiseLite::CellularReturnType)c->operation);
		for (int k = 0; k < n; k++)
			out[k] = dims == 2 ? f.GetNoise(pos[k * 3 + 0], pos[k * 3 + 1]) : f.GetNoise(pos[k * 3 + 0], pos[k * 3 + 1], pos[k * 3 + 2]);
	}
	// the 256 x 4 jitter table of the Cellular generator this library evaluates with: upstream's constant, or the declared substitute
	void upref_cellular_table(float *out1024)
	{
#if defined(UPREF_UPSTREAM_FASTNOISE)
		for (int i = 0; i < 1024; i++)
			out1024[i] = FastNoiseLite::Lookup<float>::RandVecs3D[i];
#else
		const float *t = upo_randvecs3d();
		for (int i = 0; i < 1024; i++)
			out1024[i] = t[i];
#endif
	}

	int upref_eval(int what, const float *pos, int n, float *out)
	{
		return guarded(
			[&]()
			{
				Sdf f = pickField(what);
				for (int i = 0; i < n; i++)
					out[i] = f(Vec3(pos[i * 3 + 0], pos[i * 3 + 1], pos[i * 3 + 2])).value;
			});
	}

	// grid positions only (externals/cage/sources/libcore/mesh/marchingCubes.cpp:247-257)
	void upref_grid_positions(float boxSize, int res, float *out /* res*3: x(i) y(i) z(i) share one axis */)
	{
		const MarchingCubesCreateConfig cfg = gridConfig(boxSize, res);
		for (int i = 0; i < res; i++)
			out[i] = cfg.position(i, 0, 0)[0].value;
	}

	// marchingCubes.cpp:149-165 driven with the reference field functions; threads > 1 splits z (same arithmetic per sample)
	int upref_grid(int what, float boxSize, int res, int z0, int z1, float *out /* (z1-z0)*res*res */, int threads)
	{
		return guarded(
			[&]()
			{
				Sdf f = pickField(what);
				const MarchingCubesCreateConfig cfg = gridConfig(boxSize, res);
				if (threads <= 1 && z0 == 0 && z1 == res)
				{
					Holder<MarchingCubes> cubes = newMarchingCubes(cfg);
					struct Bound
					{
						Sdf f;
						Real call(const Vec3 &p) { return f(p); }
					} bound{ f };
					Delegate<Real(const Vec3 &)> d;
					d.bind<Bound, &Bound::call>(&bound);
					cubes->updateByPosition(d);
					std::memcpy(out, cubes->densities().data(), sizeof(float) * (size_t)res * res * res);
					return;
				}
				std::atomic<int> next{ z0 };
				std::atomic<bool> failed{ false };
				auto work = [&]()
				{
					try
					{
						for (;;)
						{
							const int z = next++;
							if (z >= z1)
								break;
							float *o = out + (size_t)(z - z0) * res * res;
							for (int y = 0; y < res; y++)
								for (int x = 0; x < res; x++)
									*o++ = f(cfg.position(x, y, z)).value;
						}
					}
					catch (...)
					{
						failed = true;
					}
				};
				std::vector<std::thread> pool;
				for (int t = 1; t < threads; t++)
					pool.emplace_back(work);
				work();
				for (auto &t : pool)
					t.join();
				if (failed)
					CAGE_THROW_ERROR(Exception, "sampling failed");
			});
	}

	// raw dualmc output (externals/cage/externals/dualmc/dualmc/include/dualmc.tpp:222-249), iso 0
	struct upref_buf
	{
		void *ptr;
		uint64_t count;
	};
	void upref_free(void *p)
	{
		std::free(p);
	}

	int upref_dualmc(const float *dens, int rx, int ry, int rz, float iso, int manifold, upref_buf *vertices /* float3 */, upref_buf *quads /* int4 */)
	{
		return guarded(
			[&]()
			{
				dualmc::DualMC<float> mc;
				std::vector<dualmc::Vertex> v;
				std::vector<dualmc::Quad> q;
				mc.build(dens, rx, ry, rz, iso, !!manifold, false, v, q);
				vertices->count = v.size();
				vertices->ptr = std::malloc(sizeof(float) * 3 * v.size() + 16);
				std::memcpy(vertices->ptr, v.data(), sizeof(float) * 3 * v.size());
				quads->count = q.size();
				quads->ptr = std::malloc(sizeof(int32_t) * 4 * q.size() + 16);
				std::memcpy(quads->ptr, q.data(), sizeof(int32_t) * 4 * q.size());
			});
	}

	// MarchingCubes::makeMesh (marchingCubes.cpp:176-245) and the tail of meshGenerateGeneric (sources/meshGeneration.cpp:40-43)
	// stages: bit0 = makeMesh with clip (as the generator does), bit1 = meshRemoveDisconnected, bit2 = meshFlipNormals;
	// bit3 = makeMesh WITHOUT clip/merge is not offered by cage (clip=false merges close vertices instead), see cfg.clip
	int upref_make_mesh(const float *dens, float boxSize, int res, int clip, int stages, upref_buf *positions, upref_buf *normals, upref_buf *indices)
	{
		return guarded(
			[&]()
			{
				MarchingCubesCreateConfig cfg = gridConfig(boxSize, res);
				cfg.clip = !!clip;
				Holder<MarchingCubes> cubes = newMarchingCubes(cfg);
				cubes->densities(PointerRange<const Real>((const Real *)dens, (const Real *)dens + (size_t)res * res * res));
				Holder<Mesh> poly = cubes->makeMesh();
				if (stages & 2)
					meshRemoveDisconnected(+poly);
				if (stages & 4)
					meshFlipNormals(+poly);
				const uint32 nv = poly->verticesCount();
				const uint32 ni = poly->indicesCount();
				positions->count = nv;
				positions->ptr = std::malloc(sizeof(float) * 3 * nv + 16);
				std::memcpy(positions->ptr, poly->positions().data(), sizeof(float) * 3 * nv);
				normals->count = poly->normals().size();
				normals->ptr = std::malloc(sizeof(float) * 3 * nv + 16);
				std::memcpy(normals->ptr, poly->normals().data(), sizeof(float) * 3 * poly->normals().size());
				indices->count = ni;
				indices->ptr = std::malloc(sizeof(uint32_t) * ni + 16);
				std::memcpy(indices->ptr, poly->indices().data(), sizeof(uint32_t) * ni);
			});
	}

	// ---- the serial geometry between stage 2 and stage 3, as the generator runs it (sources/meshGeneration.cpp:173-197): meshChunking
	// (maxSurfaceArea 800000) and meshUnwrap (xatlas; texelsPerUnit 1.35 / 0.3 in debug builds, padding 6).  Out of the product's scope
	// (SURVEY.md 8f n4); here so that the GPU bake can be checked on REAL chunk atlases instead of a synthetic chart.
	static std::vector<Holder<Mesh>> heldChunks;
	int upref_chunking(const float *pos, const float *nrm, uint32_t nv, const uint32_t *idx, uint32_t ni, float maxSurfaceArea, uint32_t *chunkCount)
	{
		return guarded(
			[&]()
			{
				Holder<Mesh> m = newMesh();
				m->positions(PointerRange<const Vec3>((const Vec3 *)pos, (const Vec3 *)pos + nv));
				m->normals(PointerRange<const Vec3>((const Vec3 *)nrm, (const Vec3 *)nrm + nv));
				m->indices(PointerRange<const uint32>(idx, idx + ni));
				MeshChunkingConfig cfg;
				cfg.maxSurfaceArea = maxSurfaceArea;
				cfg.parallelize = false;
				auto r = meshChunking(+m, cfg);
				heldChunks.clear();
				for (auto &it : r)
					heldChunks.push_back(std::move(it));
				*chunkCount = (uint32_t)heldChunks.size();
			});
	}
	// chunk i as meshChunking left it (before any unwrap): its sizes; upref_chunk_copy(i, pos, nrm, NULL, idx) hands out the arrays
	void upref_chunk_size(uint32_t i, uint32_t *nverts, uint32_t *nidx)
	{
		*nverts = heldChunks[i]->verticesCount();
		*nidx = heldChunks[i]->indicesCount();
	}
	// meshClip(Mesh *, const Aabb &) (meshAlgorithms.cpp:1056-1114) of a copy of the mesh against box6 = { a.x, a.y, a.z, b.x, b.y, b.z }: the step the
	// chunking recursion repeats.  The result is held as chunk 0 (upref_chunk_size / upref_chunk_copy).
	int upref_mesh_clip_box(const float *pos, const float *nrm, uint32_t nv, const uint32_t *idx, uint32_t ni, const float *box6, uint32_t *nverts, uint32_t *nidx)
	{
		return guarded(
			[&]()
			{
				Holder<Mesh> m = newMesh();
				m->positions(PointerRange<const Vec3>((const Vec3 *)pos, (const Vec3 *)pos + nv));
				m->normals(PointerRange<const Vec3>((const Vec3 *)nrm, (const Vec3 *)nrm + nv));
				m->indices(PointerRange<const uint32>(idx, idx + ni));
				meshClip(+m, Aabb(Vec3(box6[0], box6[1], box6[2]), Vec3(box6[3], box6[4], box6[5])));
				heldChunks.clear();
				heldChunks.push_back(std::move(m));
				*nverts = heldChunks[0]->verticesCount();
				*nidx = heldChunks[0]->indicesCount();
			});
	}
	// meshUnwrap(const Holder<Mesh> &) of sources/meshGeneration.cpp:184-197 on chunk i; returns the atlas resolution
	int upref_chunk_unwrap(uint32_t i, float texelsPerUnit, uint32_t padding, uint32_t *resolution, uint32_t *nverts, uint32_t *nidx)
	{
		return guarded(
			[&]()
			{
				MeshUnwrapConfig cfg;
				cfg.maxChartIterations = 10;
				cfg.maxChartBoundaryLength = 300;
				cfg.chartRoundness = 0.3;
				cfg.texelsPerUnit = texelsPerUnit;
				cfg.padding = padding;
				*resolution = meshUnwrap(+heldChunks[i], cfg);
				*nverts = heldChunks[i]->verticesCount();
				*nidx = heldChunks[i]->indicesCount();
			});
	}
	void upref_chunk_copy(uint32_t i, float *pos, float *nrm, float *uv, uint32_t *idx)
	{
		const Mesh *m = +heldChunks[i];
		std::memcpy(pos, m->positions().data(), sizeof(Vec3) * m->verticesCount());
		std::memcpy(nrm, m->normals().data(), sizeof(Vec3) * m->normals().size());
		if (uv)
			std::memcpy(uv, m->uvs().data(), sizeof(Vec2) * m->uvs().size());
		std::memcpy(idx, m->indices().data(), sizeof(uint32) * m->indicesCount());
	}

	// terrainTile (sources/modesConfiguration.cpp:281-286) on explicit position/normal pairs
	// out per tile: albedo rgb, roughness, metallic, height, elevation, slope(rad), temperature, precipitation, opacity, biome, type  (13 floats)
	int upref_tiles(int purpose /* 1 land, 2 water, 3 navigation */, const float *pos, const float *nrm, int n, float *out)
	{
		return guarded(
			[&]()
			{
				for (int i = 0; i < n; i++)
				{
					Tile t;
					t.position = Vec3(pos[i * 3 + 0], pos[i * 3 + 1], pos[i * 3 + 2]);
					t.normal = Vec3(nrm[i * 3 + 0], nrm[i * 3 + 1], nrm[i * 3 + 2]);
					t.meshPurpose = (MeshPurposeEnum)purpose;
					terrainTile(t);
					float *o = out + (size_t)i * 13;
					o[0] = t.albedo[0].value;
					o[1] = t.albedo[1].value;
					o[2] = t.albedo[2].value;
					o[3] = t.roughness.value;
					o[4] = t.metallic.value;
					o[5] = t.height.value;
					o[6] = t.elevation.value;
					o[7] = Real(t.slope).value;
					o[8] = t.temperature.value;
					o[9] = t.precipitation.value;
					o[10] = t.opacity.value;
					o[11] = (float)(int)t.biome;
					o[12] = (float)(int)t.type;
				}
			});
	}

	// As upref_tiles, but the calling thread's random generator -- which colorDeviation() in generateSand draws from
	// (sources/coloringModes.cpp:456, sources/math.cpp:40-45), making the reference's sand texels differ from run to
	// run -- is re-seeded before every tile from a hash of the tile position.  The product uses the same
	// position-keyed state (unnatural_planets_b200/csrc/upg_coloring.cuh texel_rng_state); the reference code itself
	// runs unmodified, only its RNG *input* is pinned.
	int upref_tiles_seeded(int purpose, const float *pos, const float *nrm, int n, float *out)
	{
		int rc = 0;
		for (int i = 0; i < n && rc == 0; i++)
		{
			uint64_t s[2];
			texelRngState(pos + (size_t)i * 3, s);
			detail::randomGenerator() = RandomGenerator(s[0], s[1]);
			rc = upref_tiles(purpose, pos + (size_t)i * 3, nrm + (size_t)i * 3, 1, out + (size_t)i * 13);
		}
		return rc;
	}

	// generateTexturesLand / Water (sources/textures.cpp:86-96): rasterise + colour + 7 dilations + u8
	int upref_bake(int water, const float *pos, const float *nrm, const float *uv, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint8_t *albedo, uint8_t *special, uint8_t *height)
	{
		return guarded(
			[&]()
			{
				Holder<Mesh> msh = newMesh();
				msh->positions(PointerRange<const Vec3>((const Vec3 *)pos, (const Vec3 *)pos + nverts));
				msh->normals(PointerRange<const Vec3>((const Vec3 *)nrm, (const Vec3 *)nrm + nverts));
				msh->uvs(PointerRange<const Vec2>((const Vec2 *)uv, (const Vec2 *)uv + nverts));
				msh->indices(PointerRange<const uint32>(idx, idx + nidx));
				Holder<Image> a, s, hm;
				if (water)
					generateTexturesWater(msh, w, h, a, s, hm);
				else
					generateTexturesLand(msh, w, h, a, s, hm);
				std::memcpy(albedo, a->rawViewU8().data(), a->rawViewU8().size());
				std::memcpy(special, s->rawViewU8().data(), s->rawViewU8().size());
				std::memcpy(height, hm->rawViewU8().data(), hm->rawViewU8().size());
			});
	}

	// Generator<Water>::generate (sources/textures.cpp:45-82) re-assembled from the reference's own building blocks
	// (imageFill, meshGenerateTexture, terrainTile, imageDilation, imageConvert) with ONE difference: the thread RNG is
	// re-seeded from the tile position before every terrainTile call (see upref_tiles_seeded), which makes generateSand's
	// colorDeviation reproducible.  tests/ check that without sand (barren / debug colouring) this equals upref_bake bit for bit.
	int upref_bake_seeded(int water, const float *pos, const float *nrm, const float *uv, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint8_t *albedo, uint8_t *special, uint8_t *height, uint64_t *covered)
	{
		return guarded(
			[&]()
			{
				Holder<Mesh> msh = newMesh();
				msh->positions(PointerRange<const Vec3>((const Vec3 *)pos, (const Vec3 *)pos + nverts));
				msh->normals(PointerRange<const Vec3>((const Vec3 *)nrm, (const Vec3 *)nrm + nverts));
				msh->uvs(PointerRange<const Vec2>((const Vec2 *)uv, (const Vec2 *)uv + nverts));
				msh->indices(PointerRange<const uint32>(idx, idx + nidx));
				struct Gen
				{
					Mesh *mesh;
					Holder<Image> a, s, hm;
					bool water;
					void pixel(const Vec2i &xy, const Vec3i &indices, const Vec3 &weights)
					{
						Tile tile;
						tile.position = mesh->positionAt(indices, weights);
						tile.normal = mesh->normalAt(indices, weights);
						uint64_t st[2];
						texelRngState((const float *)&tile.position, st);
						detail::randomGenerator() = RandomGenerator(st[0], st[1]);
						tile.meshPurpose = water ? MeshPurposeEnum::Water : MeshPurposeEnum::Land;
						terrainTile(tile);
						if (water)
							a->set(xy, Vec4(tile.albedo, tile.opacity));
						else
							a->set(xy, tile.albedo);
						s->set(xy, Vec2(tile.roughness, tile.metallic));
						hm->set(xy, tile.height);
					}
				} gen;
				gen.mesh = +msh;
				gen.water = !!water;
				gen.a = newImage();
				gen.a->initialize(w, h, water ? 4 : 3, ImageFormatEnum::Float);
				if (water)
					imageFill(+gen.a, Vec4::Nan());
				else
					imageFill(+gen.a, Vec3::Nan());
				gen.s = newImage();
				gen.s->initialize(w, h, 2, ImageFormatEnum::Float);
				imageFill(+gen.s, Vec2::Nan());
				gen.hm = newImage();
				gen.hm->initialize(w, h, 1, ImageFormatEnum::Float);
				imageFill(+gen.hm, Real::Nan());
				MeshGenerateTextureConfig cfg;
				cfg.width = w;
				cfg.height = h;
				cfg.generator.bind<Gen, &Gen::pixel>(&gen);
				meshGenerateTexture(+msh, cfg);
				if (covered)
				{
					uint64_t c = 0;
					const float *hv = (const float *)gen.hm->rawViewFloat().data();
					for (size_t i = 0; i < (size_t)w * h; i++)
						c += hv[i] == hv[i];
					*covered = c;
				}
				imageDilation(+gen.a, 7, true);
				imageDilation(+gen.s, 7, true);
				imageDilation(+gen.hm, 7, true);
				imageConvert(+gen.a, ImageFormatEnum::U8);
				imageConvert(+gen.s, ImageFormatEnum::U8);
				imageConvert(+gen.hm, ImageFormatEnum::U8);
				std::memcpy(albedo, gen.a->rawViewU8().data(), gen.a->rawViewU8().size());
				std::memcpy(special, gen.s->rawViewU8().data(), gen.s->rawViewU8().size());
				std::memcpy(height, gen.hm->rawViewU8().data(), gen.hm->rawViewU8().size());
			});
	}

	// meshGenerateTexture coverage only (externals/cage/sources/libcore/mesh/meshAlgorithms.cpp:1676-1729):
	// owner[y*w+x] = index of the LAST triangle whose scan covers the texel (0xffffffff if none), weights of that visit
	int upref_raster(const float *pos, const float *uv, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t *owner, float *weights /* w*h*3 */, uint64_t *callbacks)
	{
		return guarded(
			[&]()
			{
				Holder<Mesh> msh = newMesh();
				msh->positions(PointerRange<const Vec3>((const Vec3 *)pos, (const Vec3 *)pos + nverts));
				msh->uvs(PointerRange<const Vec2>((const Vec2 *)uv, (const Vec2 *)uv + nverts));
				msh->indices(PointerRange<const uint32>(idx, idx + nidx));
				struct Rec
				{
					uint32_t w, h;
					uint32_t *owner;
					float *weights;
					uint64_t calls = 0;
					const uint32_t *idx;
					uint32_t nidx;
					void pixel(const Vec2i &xy, const Vec3i &ids, const Vec3 &wg)
					{
						calls++;
						if (xy[0] < 0 || xy[1] < 0 || (uint32_t)xy[0] >= w || (uint32_t)xy[1] >= h)
							return;
						const size_t o = (size_t)xy[1] * w + xy[0];
						// recover the triangle index: the rasteriser walks triangles in order, track via ids match
						owner[o] = cur;
						weights[o * 3 + 0] = wg[0].value;
						weights[o * 3 + 1] = wg[1].value;
						weights[o * 3 + 2] = wg[2].value;
					}
					uint32_t cur = 0;
				} rec{ w, h, owner, weights };
				for (size_t i = 0; i < (size_t)w * h; i++)
					owner[i] = 0xffffffffu;
				// one triangle at a time so that the callback knows the triangle index
				for (uint32_t t = 0; t < nidx / 3; t++)
				{
					Holder<Mesh> one = newMesh();
					one->positions(msh->positions());
					one->uvs(msh->uvs());
					const uint32 tri[3] = { idx[t * 3 + 0], idx[t * 3 + 1], idx[t * 3 + 2] };
					one->indices(tri);
					rec.cur = t;
					MeshGenerateTextureConfig cfg;
					cfg.width = w;
					cfg.height = h;
					cfg.generator.bind<Rec, &Rec::pixel>(&rec);
					meshGenerateTexture(+one, cfg);
				}
				*callbacks = rec.calls;
			});
	}

	// colour KATs (externals/cage/sources/libcore/color.cpp)
	void upref_color(int what, const float *in3, float *out3)
	{
		const Vec3 v(in3[0], in3[1], in3[2]);
		Vec3 r;
		switch (what)
		{
			case 0:
				r = colorRgbToHsv(v);
				break;
			case 1:
				r = colorHsvToRgb(v);
				break;
			case 2:
				r = colorRgbToHsluv(v);
				break;
			case 3:
				r = colorHsluvToRgb(v);
				break;
		}
		out3[0] = r[0].value;
		out3[1] = r[1].value;
		out3[2] = r[2].value;
	}
	void upref_color_interpolate(const float *a3, const float *b3, float f, float *out3)
	{
		const Vec3 r = interpolateColor(Vec3(a3[0], a3[1], a3[2]), Vec3(b3[0], b3[1], b3[2]), f);
		out3[0] = r[0].value;
		out3[1] = r[1].value;
		out3[2] = r[2].value;
	}

	// imageDilation(float image, rounds, useNan) + imageConvert(U8) (imageAlgorithms.cpp:504-527, 100-113)
	int upref_dilate(float *img, uint32_t w, uint32_t h, uint32_t ch, uint32_t rounds, int useNan, uint8_t *u8out /* may be null */)
	{
		return guarded(
			[&]()
			{
				Holder<Image> im = newImage();
				im->importRaw(PointerRange<const char>((const char *)img, (const char *)img + sizeof(float) * (size_t)w * h * ch), w, h, ch, ImageFormatEnum::Float);
				imageDilation(+im, rounds, !!useNan);
				std::memcpy(img, im->rawViewFloat().data(), sizeof(float) * (size_t)w * h * ch);
				if (u8out)
				{
					imageConvert(+im, ImageFormatEnum::U8);
					std::memcpy(u8out, im->rawViewU8().data(), (size_t)w * h * ch);
				}
			});
	}
	// the texture tail the generator applies to every chunk's U8 images before export (sources/generator.cpp:135-139):
	// imageConvertSpecialToGltfPbr (imageAlgorithms.cpp:296-302) and imageConvertHeigthToNormal(img, strength) (:216-246)
	int upref_texture_tail(const uint8_t *special /* w*h*2 */, const uint8_t *height /* w*h*1 */, uint32_t w, uint32_t h, float strength, uint8_t *pbr /* w*h*3 */, uint8_t *normal /* w*h*3 */)
	{
		return guarded(
			[&]()
			{
				Holder<Image> sp = newImage();
				sp->importRaw(PointerRange<const char>((const char *)special, (const char *)special + (size_t)w * h * 2), w, h, 2, ImageFormatEnum::U8);
				imageConvertSpecialToGltfPbr(+sp);
				if (sp->channels() != 3 || sp->format() != ImageFormatEnum::U8)
					CAGE_THROW_ERROR(Exception, "unexpected pbr image layout");
				std::memcpy(pbr, sp->rawViewU8().data(), (size_t)w * h * 3);
				Holder<Image> hm = newImage();
				hm->importRaw(PointerRange<const char>((const char *)height, (const char *)height + (size_t)w * h), w, h, 1, ImageFormatEnum::U8);
				imageConvertHeigthToNormal(+hm, strength);
				if (hm->channels() != 3 || hm->format() != ImageFormatEnum::U8)
					CAGE_THROW_ERROR(Exception, "unexpected normal image layout");
				std::memcpy(normal, hm->rawViewU8().data(), (size_t)w * h * 3);
			});
	}
	// the vertex culling of meshGenerateBaseWater (sources/meshGeneration.cpp:62-95) applied to a given mesh: that translation unit
	// drags the whole generator in, so its few lines are re-stated here over the reference's own terrainSdfElevationRaw,
	// meshConvertToIndexed and meshRemoveInvalid
	int upref_water_cull(const float *pos, const float *nrm, uint32_t nverts, const uint32_t *idx, uint32_t nidx, upref_buf *positions, upref_buf *normals, upref_buf *indices)
	{
		return guarded(
			[&]()
			{
				Holder<Mesh> poly = newMesh();
				poly->positions(PointerRange<const Vec3>((const Vec3 *)pos, (const Vec3 *)pos + nverts));
				poly->normals(PointerRange<const Vec3>((const Vec3 *)nrm, (const Vec3 *)nrm + nverts));
				poly->indices(PointerRange<const uint32>(idx, idx + nidx));
				meshConvertToIndexed(+poly);
				std::vector<bool> valid;
				valid.reserve(poly->verticesCount());
				for (const Vec3 &p : poly->positions())
					valid.push_back(unnatural::terrainSdfElevationRaw(p) < 0.1);
				for (uint32 j = 0; j < 2; j++)
				{
					std::vector<bool> valid2 = valid;
					const uint32 cnt = poly->indicesCount();
					const auto inds = poly->indices();
					for (uint32 i = 0; i < cnt; i += 3)
					{
						const uint32 is[3] = { inds[i + 0], inds[i + 1], inds[i + 2] };
						if (valid[is[0]] || valid[is[1]] || valid[is[2]])
							valid2[is[0]] = valid2[is[1]] = valid2[is[2]] = true;
					}
					std::swap(valid, valid2);
				}
				{
					std::vector<Vec3> ps(poly->positions().begin(), poly->positions().end());
					for (uint32 i = 0; i < ps.size(); i++)
						if (!valid[i])
							ps[i] = Vec3::Nan();
					poly->positions(ps);
				}
				meshRemoveInvalid(+poly);
				const uint32 nv = poly->verticesCount();
				const uint32 ni = poly->indicesCount();
				positions->count = nv;
				positions->ptr = std::malloc(sizeof(float) * 3 * nv + 16);
				std::memcpy(positions->ptr, poly->positions().data(), sizeof(float) * 3 * nv);
				normals->count = poly->normals().size();
				normals->ptr = std::malloc(sizeof(float) * 3 * nv + 16);
				std::memcpy(normals->ptr, poly->normals().data(), sizeof(float) * 3 * poly->normals().size());
				indices->count = ni;
				indices->ptr = std::malloc(sizeof(uint32_t) * ni + 16);
				std::memcpy(indices->ptr, poly->indices().data(), sizeof(uint32_t) * ni);
			});
	}
}
