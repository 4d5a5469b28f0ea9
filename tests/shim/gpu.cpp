// sources/gpu.cpp -- THE PATCH a maintainer adds to unnatural-planets to run its generation core on libupg.so.
//
// A new translation unit of the generator, compiled here against the reference's own headers (cage-core, sources/planets.h) and linked
// with the reference's objects (oracle/_ref) and libupg.so by tests/shim/Makefile; tests/test_shim.py drives it through the C entry
// points at the bottom.  It REPLACES, signature for signature:
//     Holder<Mesh> meshGenerateBaseLand() / Water() / Navigation()                     sources/meshGeneration.cpp:47-107
//     void generateTexturesLand / Water(const Holder<Mesh> &, uint32, uint32, Holder<Image> &, Holder<Image> &, Holder<Image> &)
//                                                                                      sources/textures.cpp:86-96
// (the patched build drops those bodies from meshGeneration.cpp / textures.cpp) and adds
//     void terrainGpuCreate()     called by generateEntry right BEFORE terrainPreseed()   sources/generator.cpp:430
//     the --seed option           sources/main.cpp:17-54 (applyConfiguration), see applySeedOption() below
// Everything else of the generator stays as it is: the mode tables, terrainSdf* / terrainTile for the remaining CPU callers,
// simplification, chunking, unwrapping, file export.
#include <mutex>
#include <string>

#include <cage-core/config.h>
#include <cage-core/image.h>
#include <cage-core/ini.h>
#include <cage-core/logger.h>
#include <cage-core/math.h>
#include <cage-core/memoryBuffer.h>
#include <cage-core/pointerRangeHolder.h>
#include <cage-core/mesh.h>
#include <cage-core/random.h>
#include <cage-core/string.h>

#include "upg.h"

using namespace cage;

namespace unnatural
{
	namespace
	{
		// the generator's configuration variables are global, string-keyed objects (include/cage-core/config.h:41-73): these refer to the
		// same values main.cpp:19-35 sets and modesConfiguration.cpp:22-26 reads
		const ConfigString configShapeMode("unnatural-planets/shape/mode");
		const ConfigString configElevationMode("unnatural-planets/elevation/mode");
		const ConfigString configColoringMode("unnatural-planets/coloring/mode");
		const ConfigBool configPolesEnable("unnatural-planets/poles/enable");
		const ConfigBool configFlowersEnable("unnatural-planets/flowers/enable");
		// new options: the reference hard-codes the grid (sources/meshGeneration.cpp:19-28) and has no seed
		const ConfigUint32 configGridResolution("unnatural-planets/grid/resolution", CAGE_DEBUG_BOOL ? 110 : 500);
		const ConfigString configGpuDevices("unnatural-planets/gpu/devices", "0");

		constexpr Real boxSize = 4000; // sources/meshGeneration.cpp:19

		upg_planet *gpuPlanet = nullptr;

		void upgCheck(int rc)
		{
			if (rc != UPG_OK)
			{
				CAGE_LOG_THROW(upg_last_error());
				CAGE_THROW_ERROR(Exception, "gpu generation core failed"); // no exception crosses the C boundary: cage's is thrown here
			}
		}

		std::once_flag baseOnce;
		int gpuHomeDevice = 0; // first device of unnatural-planets/gpu/devices
		upg_mesh baseMeshes[3]; // UPG_FIELD_LAND, UPG_FIELD_WATER, UPG_FIELD_NAVIGATION
		int baseRc = UPG_OK;
		std::string baseError;

		// The three base meshes are three concurrent tasks (sources/generator.cpp:432-439); the first to arrive generates all three in
		// one call (one fused sampling pass serves the three fields), the others pick theirs up.
		Holder<Mesh> takeBaseMesh(int field, const char *name)
		{
			CAGE_LOG(SeverityEnum::Info, "generator", Stringizer() + "generating base " + name + " mesh");
			std::call_once(baseOnce,
				[]()
				{
					baseRc = upg_field_meshes_clean(gpuPlanet, 7u, UPG_MESH_GENERATOR, baseMeshes);
					if (baseRc != UPG_OK)
						baseError = upg_last_error();
				});
			if (baseRc != UPG_OK)
			{
				CAGE_LOG_THROW(baseError.c_str());
				if (baseRc == UPG_ERR_EMPTY_MESH)
					CAGE_THROW_ERROR(Exception, "generated empty base mesh"); // sources/meshGeneration.cpp:51-52, 104-105
				CAGE_THROW_ERROR(Exception, "gpu generation core failed");
			}
			upg_mesh &m = baseMeshes[field];
			Holder<Mesh> poly = newMesh();
			if (m.index_count)
			{
				poly->positions(PointerRange<const Vec3>((const Vec3 *)m.positions, (const Vec3 *)m.positions + m.vertex_count)); // mesh.h:34-36
				poly->normals(PointerRange<const Vec3>((const Vec3 *)m.normals, (const Vec3 *)m.normals + m.vertex_count));
				poly->indices(PointerRange<const uint32>(m.indices, m.indices + m.index_count));
			}
			upg_mesh_free(&m);
			return poly;
		}

		void generateTextures(bool water, const Holder<Mesh> &renderMesh, uint32 width, uint32 height, Holder<Image> &albedo, Holder<Image> &special, Holder<Image> &heightMap)
		{
			// Image exposes its pixels read-only (rawViewU8() is const, image.h:67) but adopts a filled buffer: importRaw(MemoryBuffer &&)
			const uint32 ch = water ? 4 : 3;
			MemoryBuffer a((uintPtr)width * height * ch), s((uintPtr)width * height * 2), h((uintPtr)width * height);
			upgCheck(upg_bake(gpuPlanet, water ? 1 : 0, (const float *)renderMesh->positions().data(), (const float *)renderMesh->normals().data(), (const float *)renderMesh->uvs().data(), renderMesh->verticesCount(), renderMesh->indices().data(), renderMesh->indicesCount(), width, height, (uint8_t *)a.data(), (uint8_t *)s.data(), (uint8_t *)h.data()));
			albedo = newImage();
			albedo->importRaw(std::move(a), width, height, ch, ImageFormatEnum::U8);
			special = newImage();
			special->importRaw(std::move(s), width, height, 2, ImageFormatEnum::U8);
			heightMap = newImage();
			heightMap->importRaw(std::move(h), width, height, 1, ImageFormatEnum::U8);
		}
	}

	// ---- sources/main.cpp, inside applyConfiguration() before terrainApplyConfig(): the seed option the reference lacks --
	// `--seed a,b` (two non-zero 64-bit words = the xoroshiro128** state) or `--seed n` (one word, expanded by splitmix64); without the
	// option the generator keeps seeding from std::random_device (externals/cage/sources/libcore/math/random.cpp:187-194).
	void applySeedOption(const Holder<Ini> &cmd)
	{
		const String text = cmd->cmdString('x', "seed", "");
		if (text.empty())
			return;
		uint64 s[2];
		const uint32 comma = find(text, ',');
		if (comma != m)
		{
			s[0] = toUint64(subString(text, 0, comma));
			s[1] = toUint64(subString(text, comma + 1, m));
		}
		else
		{
			uint64 x = toUint64(text);
			for (uint64 &w : s)
			{
				x += 0x9e3779b97f4a7c15ull;
				uint64 z = x;
				z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
				z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
				w = z ^ (z >> 31);
			}
		}
		if (s[0] == 0 || s[1] == 0)
			CAGE_THROW_ERROR(Exception, "seed words must be non-zero");
		detail::randomGenerator() = RandomGenerator(s[0], s[1]);
		CAGE_LOG(SeverityEnum::Info, "configuration", Stringizer() + "seed: " + s[0] + ", " + s[1]);
	}

	// ---- sources/generator.cpp:430, right BEFORE terrainPreseed(): the planet takes the state of the generator that noiseSeed() is about
	// to copy (sources/math.cpp:73-77) and draws the same shape constants and noise seeds in the same order --------------------------------
	void terrainGpuCreate()
	{
		upg_planet_desc d;
		upg_planet_desc_default(&d);
		d.shape = upg_shape_id(String(configShapeMode).c_str()); // same name tables, same order (sources/modesConfiguration.cpp:78-113)
		d.elevation = upg_elevation_id(String(configElevationMode).c_str());
		d.coloring = upg_coloring_id(String(configColoringMode).c_str());
		if (d.shape < 0 || d.elevation < 0 || d.coloring < 0)
			CAGE_THROW_ERROR(Exception, "unknown mode configuration"); // `random` is resolved by chooseShapeFunction() before this point (:118-123)
		d.poles = configPolesEnable;
		d.flowers = configFlowersEnable;
		d.seed[0] = detail::randomGenerator().s[0];
		d.seed[1] = detail::randomGenerator().s[1];
		d.box_size = boxSize.value;
		d.grid_resolution = (sint32)(uint32)configGridResolution;
		d.mesh_sampling = UPG_SAMPLING_BAND;
		// "0" or "0,1,2,3": the GPUs of this node that share the planet (one process, upg_planet_desc::device_count)
		String devices = configGpuDevices;
		d.device_count = 0;
		while (!devices.empty() && d.device_count < UPG_MAX_DEVICES)
			d.devices[d.device_count++] = toSint32(split(devices, ","));
		d.device = d.device_count ? d.devices[0] : 0;
		gpuHomeDevice = d.device;
		upgCheck(upg_planet_create(&d, &gpuPlanet));
	}

	Holder<Mesh> meshGenerateBaseLand() { return takeBaseMesh(UPG_FIELD_LAND, "land"); } // replaces sources/meshGeneration.cpp:47-55
	Holder<Mesh> meshGenerateBaseWater() { return takeBaseMesh(UPG_FIELD_WATER, "water"); } // :57-98
	Holder<Mesh> meshGenerateBaseNavigation() { return takeBaseMesh(UPG_FIELD_NAVIGATION, "navigation"); } // :100-107

	// replaces meshSplit (sources/meshGeneration.cpp:173-182): cage's meshChunking(maxSurfaceArea 800000) on the device.  Same chunks in the same
	// order (first half before second half, depth first; parts the clip left empty stay in the list).  Compiled here against cage's headers;
	// the device side has not run on hardware yet, so the test entry does not call it.
	Holder<PointerRange<Holder<Mesh>>> meshSplitGpu(const Holder<Mesh> &mesh)
	{
		upg_mesh *chunks = nullptr;
		uint32_t count = 0;
		upgCheck(upg_mesh_chunking(gpuHomeDevice, (const float *)mesh->positions().data(), (const float *)mesh->normals().data(), mesh->verticesCount(), mesh->indices().data(), mesh->indicesCount(), 800000.0f, &chunks, &count));
		PointerRangeHolder<Holder<Mesh>> result;
		result.reserve(count);
		for (uint32_t i = 0; i < count; i++)
		{
			const upg_mesh &m = chunks[i];
			Holder<Mesh> poly = newMesh();
			if (m.index_count)
			{
				poly->positions(PointerRange<const Vec3>((const Vec3 *)m.positions, (const Vec3 *)m.positions + m.vertex_count));
				poly->normals(PointerRange<const Vec3>((const Vec3 *)m.normals, (const Vec3 *)m.normals + m.vertex_count));
				poly->indices(PointerRange<const uint32>(m.indices, m.indices + m.index_count));
			}
			result.push_back(std::move(poly));
		}
		upg_mesh_chunks_free(chunks, count);
		return result;
	}

	void generateTexturesLandGpu(const Holder<Mesh> &renderMesh, uint32 width, uint32 height, Holder<Image> &albedo, Holder<Image> &special, Holder<Image> &heightMap)
	{
		generateTextures(false, renderMesh, width, height, albedo, special, heightMap); // replaces sources/textures.cpp:86-90
	}
	void generateTexturesWaterGpu(const Holder<Mesh> &renderMesh, uint32 width, uint32 height, Holder<Image> &albedo, Holder<Image> &special, Holder<Image> &heightMap)
	{
		generateTextures(true, renderMesh, width, height, albedo, special, heightMap); // replaces :92-96
	}
#if !defined(UPG_SHIM_TEST)
	// in the patched generator these two ARE generateTexturesLand / generateTexturesWater (textures.cpp's bodies are dropped); the test
	// build keeps the reference's own definitions next to them (it compares the two), hence the suffix there
	void generateTexturesLand(const Holder<Mesh> &m, uint32 w, uint32 h, Holder<Image> &a, Holder<Image> &s, Holder<Image> &hm) { generateTexturesLandGpu(m, w, h, a, s, hm); }
	void generateTexturesWater(const Holder<Mesh> &m, uint32 w, uint32 h, Holder<Image> &a, Holder<Image> &s, Holder<Image> &hm) { generateTexturesWaterGpu(m, w, h, a, s, hm); }
#endif
}
