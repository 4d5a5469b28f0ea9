// shim_test_entry.cpp -- TEST INFRASTRUCTURE: C entry points over the compiled patch (gpu.cpp) and over the reference's own functions in
// the same process image, for tests/test_shim.py (ctypes).  Mirrors what the patched generateEntry does (sources/generator.cpp:424-439).
#include <cstring>
#include <string>

#include <cage-core/config.h>
#include <cage-core/image.h>
#include <cage-core/ini.h>
#include <cage-core/mesh.h>
#include <cage-core/random.h>

using namespace cage;

namespace unnatural
{
	void terrainApplyConfig();
	void terrainPreseed();
	void terrainGpuCreate();
	void applySeedOption(const Holder<Ini> &cmd);
	Holder<Mesh> meshGenerateBaseLand();
	Holder<Mesh> meshGenerateBaseWater();
	Holder<Mesh> meshGenerateBaseNavigation();
	void generateTexturesLandGpu(const Holder<Mesh> &, uint32, uint32, Holder<Image> &, Holder<Image> &, Holder<Image> &);
	void generateTexturesWaterGpu(const Holder<Mesh> &, uint32, uint32, Holder<Image> &, Holder<Image> &, Holder<Image> &);
	// the reference's own (oracle/_ref: sources/textures.cpp compiled verbatim)
	void generateTexturesLand(const Holder<Mesh> &, uint32, uint32, Holder<Image> &, Holder<Image> &, Holder<Image> &);
	void generateTexturesWater(const Holder<Mesh> &, uint32, uint32, Holder<Image> &, Holder<Image> &, Holder<Image> &);
}
using namespace unnatural;

namespace
{
	std::string lastError;
	Holder<Mesh> held[3];

	template<class F>
	int guarded(F f)
	{
		try
		{
			f();
			return 0;
		}
		catch (const cage::Exception &e)
		{
			lastError = (const char *)e.message;
			return 1;
		}
		catch (...)
		{
			lastError = "unknown exception";
			return 2;
		}
	}

	Holder<Mesh> meshFrom(const float *pos, const float *nrm, const float *uv, uint32 nv, const uint32 *idx, uint32 ni)
	{
		Holder<Mesh> m = newMesh();
		m->positions(PointerRange<const Vec3>((const Vec3 *)pos, (const Vec3 *)pos + nv));
		m->normals(PointerRange<const Vec3>((const Vec3 *)nrm, (const Vec3 *)nrm + nv));
		m->uvs(PointerRange<const Vec2>((const Vec2 *)uv, (const Vec2 *)uv + nv));
		m->indices(PointerRange<const uint32>(idx, idx + ni));
		return m;
	}

	void copyOut(const Holder<Image> &img, uint8_t *out)
	{
		const auto v = img->rawViewU8();
		std::memcpy(out, v.data(), v.size());
	}
}

extern "C"
{
	const char *upgshim_last_error(void) { return lastError.c_str(); }

	// main.cpp's applyConfiguration + the patched head of generateEntry: options -> config variables, --seed, terrainApplyConfig,
	// terrainGpuCreate BEFORE terrainPreseed
	int upgshim_setup(const char *shape, const char *elevation, const char *coloring, int poles, int flowers, const char *seedOption, int grid)
	{
		return guarded(
			[&]()
			{
				configSetString("unnatural-planets/shape/mode", shape);
				configSetString("unnatural-planets/elevation/mode", elevation);
				configSetString("unnatural-planets/coloring/mode", coloring);
				configSetBool("unnatural-planets/poles/enable", !!poles);
				configSetBool("unnatural-planets/flowers/enable", !!flowers);
				configSetUint32("unnatural-planets/grid/resolution", (uint32)grid);
				Holder<Ini> cmd = newIni();
				const char *args[3] = { "unnatural-planets", "--seed", seedOption }; // unnatural-planets -s sphere ... --seed a,b
				cmd->parseCmd(3, args);
				applySeedOption(cmd);
				terrainApplyConfig();
				terrainGpuCreate();
				terrainPreseed();
			});
	}

	void upgshim_rng_state(uint64_t *s)
	{
		s[0] = detail::randomGenerator().s[0];
		s[1] = detail::randomGenerator().s[1];
	}

	// field 0 land, 1 water, 2 navigation: runs the PATCHED meshGenerateBase*; sizes first, then the arrays
	int upgshim_base_mesh(int field, uint32_t *nverts, uint32_t *nidx)
	{
		return guarded(
			[&]()
			{
				held[field] = field == 0 ? meshGenerateBaseLand() : field == 1 ? meshGenerateBaseWater() : meshGenerateBaseNavigation();
				*nverts = held[field]->verticesCount();
				*nidx = held[field]->indicesCount();
			});
	}
	void upgshim_base_mesh_copy(int field, float *pos, float *nrm, uint32_t *idx)
	{
		const Mesh *m = +held[field];
		std::memcpy(pos, m->positions().data(), sizeof(Vec3) * m->verticesCount());
		std::memcpy(nrm, m->normals().data(), sizeof(Vec3) * m->verticesCount());
		std::memcpy(idx, m->indices().data(), sizeof(uint32) * m->indicesCount());
	}

	// generateTextures{Land,Water}: which = 0 the patch (GPU), 1 the reference's own body (CPU) -- same signature, same Holder<Image> results
	int upgshim_textures(int which, int water, const float *pos, const float *nrm, const float *uv, uint32_t nv, const uint32_t *idx, uint32_t ni, uint32_t w, uint32_t h, uint8_t *albedo, uint8_t *special, uint8_t *height)
	{
		return guarded(
			[&]()
			{
				Holder<Mesh> mesh = meshFrom(pos, nrm, uv, nv, idx, ni);
				Holder<Image> a, s, hm;
				if (which == 0)
					(water ? generateTexturesWaterGpu : generateTexturesLandGpu)(mesh, w, h, a, s, hm);
				else
					(water ? generateTexturesWater : generateTexturesLand)(mesh, w, h, a, s, hm);
				if (a->format() != ImageFormatEnum::U8 || a->channels() != (water ? 4u : 3u) || a->width() != w || a->height() != h)
					CAGE_THROW_ERROR(Exception, "unexpected image format");
				copyOut(a, albedo);
				copyOut(s, special);
				copyOut(hm, height);
			});
	}
}
