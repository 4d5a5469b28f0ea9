/*
 * upg.h -- C ABI of the B200 generation core ("unnatural-planets GPU"), libupg.so.
 *
 * The reference generator (unnatural-worlds/unnatural-planets) has no FFI: its plug-in surface for the
 * hot path is a pair of function-pointer types selected by name,
 *     using TerrainFunctor  = Real (*)(const Vec3 &);      sources/modesConfiguration.cpp:29-31
 *     using ColoringFunctor = void (*)(Tile &tile);        sources/modesConfiguration.cpp:33-34
 * consumed by two drivers whose C++ signatures stay unchanged in a patched build
 * (see INTEGRATION.md for the shim a maintainer would add):
 *     Holder<Mesh> meshGenerateBaseLand / Water / Navigation();          sources/meshGeneration.cpp:47,57,100
 *     void generateTexturesLand / Water(const Holder<Mesh>&, uint32 w, uint32 h,
 *                                       Holder<Image>&, Holder<Image>&, Holder<Image>&);   sources/textures.cpp:86,92
 * Each entry point below cites the reference interface it stands in for.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success and a non-zero
 * UPG_ERR_* code on failure with a message in upg_last_error() (thread local).  No exception
 * crosses this boundary (the reference throws cage::Exception, include/cage-core/core.h:263-274;
 * the host shim re-throws).  All entry points are re-entrant: the reference calls the three field
 * pipelines concurrently from three pool threads (sources/generator.cpp:432-439); every call uses
 * its own CUDA stream and a planet is immutable after creation.
 * There is NO CPU execution path: without a usable CUDA device every compute call fails with
 * UPG_ERR_CUDA.
 */
#ifndef UPG_H_INCLUDED
#define UPG_H_INCLUDED

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UPG_API_VERSION 3

enum
{
	UPG_OK = 0,
	UPG_ERR_ARGUMENT = 1, /* unknown mode name / id, null pointer, resolution <= 5 (marchingCubes.cpp:21-23) or > UPG_MAX_GRID_RESOLUTION */
	UPG_ERR_CUDA = 2, /* no device, launch or allocation failure */
	UPG_ERR_INVALID_VALUE = 3, /* a NaN sample: "invalid land sdf value" (sources/modesConfiguration.cpp:255-256) */
	UPG_ERR_EMPTY_MESH = 4 /* "generated empty base land mesh" (sources/meshGeneration.cpp:51-52) */
};

/* terrainSdf* of sources/modesConfiguration.cpp:230-279 */
enum
{
	UPG_FIELD_LAND = 0, /* terrainSdfLand:       shape - elevation / 10 */
	UPG_FIELD_WATER = 1, /* terrainSdfWater:      shape */
	UPG_FIELD_NAVIGATION = 2, /* terrainSdfNavigation: shape - max(elevation, 0) / 10 */
	UPG_FIELD_ELEVATION_RAW = 3, /* terrainSdfElevationRaw: elevation */
	UPG_FIELD_ELEVATION = 4 /* terrainSdfElevation:  shape * 10 */
};

/*
 * How upg_field_meshes* sample the land / navigation grids before the extraction.  dualmc reads of a sample only its sign, and its
 * value where it ends a grid edge whose signs differ (dualmc.tpp:12-32, 91-188); elevation / 10 is bounded per elevation mode, so far
 * from the surface the shape alone settles the sign.  BAND evaluates the elevation function exactly on every node that can end a
 * sign-changing edge and stores a value of the settled sign elsewhere: same meshes bit for bit, a few percent of the noise
 * evaluations.  DENSE evaluates every node like MarchingCubes::updateByPosition does.  upg_field_samples* are always dense.
 */
enum
{
	UPG_SAMPLING_BAND = 0,
	UPG_SAMPLING_DENSE = 1
};

/*
 * Arithmetic of the dense sampling kernel (K1, upg_field_samples* / upg_fields_samples_device).  The reference is built for baseline
 * x86-64, which has no fused multiply-add (externals/cage/cmake/cage_build_configuration.cmake sets no -march).  EXACT evaluates
 * every + - * / sqrt in the reference's order without contraction: those operations and all integer work are bit-identical to the
 * CPU.  CONTRACT lets the compiler fuse a * b + c inside the noise interpolation, the gradient dot products and the shape
 * expressions (fewer issue slots per sample); the integer lattice (hashes, cell indices: the products that feed a floor / round are
 * rounded as before), the grid coordinates and the final shape - elevation / 10 stay exact.  Values then agree with EXACT within the
 * tolerance BASELINE.json states (1e-4 relative on the SDF; measured in tests/test_gpu_field.py).  The mesh entry points, the bake
 * and upg_field_points always run EXACT: they are the validator.
 */
enum
{
	UPG_ARITH_EXACT = 0,
	UPG_ARITH_CONTRACT = 1
};

#define UPG_MAX_DEVICES 8
/* largest grid resolution a planet accepts: the isosurface stage indexes cells with 32 bits, (res - 1)^3 < 2^32 */
#define UPG_MAX_GRID_RESOLUTION 1626

typedef struct upg_planet upg_planet;

/*
 * One planet == one process run of the reference: the mode names chosen by applyConfiguration
 * (sources/main.cpp:17-54 -> terrainApplyConfig, sources/modesConfiguration.cpp:308-315) plus the state of
 * the random generator that noiseSeed() copies (sources/math.cpp:73-77).  The reference seeds that generator
 * from std::random_device and has no seed option; `seed` is the injection point SURVEY.md F4 asks for.
 */
typedef struct upg_planet_desc
{
	uint32_t struct_size; /* sizeof(upg_planet_desc), for ABI evolution */
	int32_t shape; /* index into the reference's shapeModeNames order; see upg_shape_id() */
	int32_t elevation; /* index into elevationModeNames */
	int32_t coloring; /* index into tileModeNames */
	int32_t poles; /* unnatural-planets/poles/enable */
	int32_t flowers; /* unnatural-planets/flowers/enable */
	uint64_t seed[2]; /* xoroshiro128** state (both non-zero, random.cpp:14-17) */
	float box_size; /* meshing box side, reference constant 4000 (sources/meshGeneration.cpp:19) */
	int32_t grid_resolution; /* samples per axis, reference constant 500 (110 in debug builds) (:20-28) */
	int32_t device; /* CUDA device ordinal (the planet's home device: single-device entry points run here) */
	int32_t mesh_sampling; /* UPG_SAMPLING_*: how the mesh entry points (upg_field_meshes*) fill the grid; the meshes are identical */
	int32_t arithmetic; /* UPG_ARITH_*: see above */
	/* Multi-GPU inside ONE process (the generator is a single process with pool threads, sources/generator.cpp:432-439): with
	 * device_count > 1 the entry points marked "multi-device" below split their work over devices[0 .. device_count) -- z-slabs for
	 * sampling + extraction, texel-row tiles for baking -- from one host thread per device; results are bit-identical to one device.
	 * devices[0] must equal `device`.  device_count 0 or 1 = `device` alone. */
	int32_t device_count;
	int32_t devices[UPG_MAX_DEVICES];
} upg_planet_desc;

/* fills the reference's defaults: shape "random" is NOT resolved here (pass an explicit id), elevation lakes,
 * coloring default, poles/flowers off (sources/main.cpp:19-44), box 4000, grid 500, device 0, the seed pair of
 * cage's own RNG test (externals/cage/sources/test-core/random.cpp:38) */
void upg_planet_desc_default(upg_planet_desc *desc);

/* name <-> id in the order of the reference's tables (sources/modesConfiguration.cpp:78-113, 176-183, 207-211);
 * -1 / NULL when unknown (the reference throws "unknown shape mode configuration", :131-132) */
int upg_shape_id(const char *name);
int upg_elevation_id(const char *name);
int upg_coloring_id(const char *name);
const char *upg_shape_name(int id);
const char *upg_elevation_name(int id);
const char *upg_coloring_name(int id);
int upg_shape_count(void);

/* replaces terrainApplyConfig + terrainPreseed (sources/modesConfiguration.cpp:288-315): fixes the modes, draws the
 * shape constants and every noise seed in static-initialisation order, uploads nothing yet (parameters travel as
 * kernel arguments) */
int upg_planet_create(const upg_planet_desc *desc, upg_planet **out);
void upg_planet_destroy(upg_planet *planet);

/* The ASYNCHRONOUS device-pointer entry points (upg_field_samples_device, upg_fields_samples_device, upg_fields_samples_band_device)
 * cannot report a NaN sample when they return; they raise a sticky flag on the planet instead.  This call synchronises `cuda_stream`,
 * returns UPG_ERR_INVALID_VALUE if the flag was raised since the last query ("invalid land sdf value", sources/modesConfiguration.cpp:
 * 255-256) and clears it.  Every synchronous entry point checks (and owns) its own flag: concurrent calls never see each other's. */
int upg_planet_async_status(upg_planet *planet, void *cuda_stream);

/*
 * Scratch memory.  Per-call device arrays (sample fields, isosurface / clean-up work arrays, float images) come from ONE library-owned
 * stream-ordered pool per device, mesh results and upload staging from a cache of pinned host blocks; both keep what a call
 * releases so that the next planet of a batch finds it.  upg_reserve grows the pool of `device` to `device_bytes` and pre-creates
 * `pinned_blocks` pinned host blocks of `pinned_block_bytes` each, so that the FIRST planet does not pay for the growth (the driver
 * calls cost tens of milliseconds, more than a 500^3 planet); upg_trim gives everything cached back (device pool down to `keep_bytes`,
 * all pinned blocks).  Neither touches the device's default memory pool.
 */
int upg_reserve(int device, uint64_t device_bytes, uint64_t pinned_block_bytes, int pinned_blocks);
int upg_trim(int device, uint64_t keep_bytes);
/* bytes the device pool holds from the driver / high-water mark of bytes in use / bytes of cached pinned host blocks */
int upg_memory_stats(int device, uint64_t *device_reserved, uint64_t *device_used_high, uint64_t *pinned_cached);

/* the 32-bit seeds in the order noiseSeed() handed them out; returns how many were drawn (for parity checks) */
int upg_planet_seed_log(const upg_planet *planet, uint32_t *out, int capacity);

/*
 * K1.  Replaces MarchingCubes::updateByPosition (externals/cage/sources/libcore/mesh/marchingCubes.cpp:149-165)
 * bound to one of the terrainSdf* functions: samples z-slices [z0, z1) of the res^3 grid, layout
 * out[((z - z0) * res + y) * res + x] exactly as MarchingCubesCreateConfig::index (:259-265); the result can be
 * handed to MarchingCubes::densities(PointerRange<const Real>) (include/cage-core/marchingCubes.h:16).
 * `host_out` is host memory (pinned or pageable), (z1 - z0) * res * res floats.
 */
int upg_field_samples(upg_planet *planet, int field, int z0, int z1, float *host_out);

/* same, into DEVICE memory on `cuda_stream` (a cudaStream_t, NULL = default stream), asynchronous: for callers that
 * keep the field on the GPU (bench.py's device-resident metric, the isosurface stage) */
int upg_field_samples_device(upg_planet *planet, int field, int z0, int z1, float *device_out, void *cuda_stream);

/* land, water and navigation of the same slices from ONE evaluation of the shape and the elevation function per sample
 * (the three fields differ only in how the two combine, sources/modesConfiguration.cpp:248-279, and the generator asks for
 * all three: sources/generator.cpp:432-439).  Each output is bit-identical to the single-field call; any pointer may be NULL. */
int upg_fields_samples_device(upg_planet *planet, int z0, int z1, float *device_land, float *device_water, float *device_navigation, void *cuda_stream);

/* the samples the mesh entry points extract from under UPG_SAMPLING_BAND (whatever the planet's own setting), same layout and
 * arguments: water is the dense shape field; land / navigation are EXACT (bit-identical to upg_fields_samples_device) on both end
 * points of every grid edge inside [z0, z1) whose signs differ -- in fact on every node whose sign the shape value alone does not
 * settle, and on the six neighbours of such nodes -- and carry a value of the correct sign, at least 1/64 from zero, elsewhere.
 * `evaluated_nodes` (may be NULL; host memory, written after a stream synchronisation inside the call) receives the number of nodes
 * whose elevation was evaluated.  Elevation "none" falls back to the dense kernel (nothing to skip). */
int upg_fields_samples_band_device(upg_planet *planet, int z0, int z1, float *device_land, float *device_water, float *device_navigation, void *cuda_stream, uint64_t *evaluated_nodes);

/* one terrainSdf* value per explicit position (n x 3 floats): the per-vertex terrainSdfElevationRaw loop of the water
 * mesh culling (sources/meshGeneration.cpp:62-68) and the parity probes */
int upg_field_points(upg_planet *planet, int field, const float *host_positions, uint64_t n, float *host_out);

/*
 * The triangle mesh of one field as MarchingCubes::makeMesh builds it up to and including the vertex normals
 * (externals/cage/sources/libcore/mesh/marchingCubes.cpp:176-227): dualmc::DualMC<float>::build(iso 0, manifold, shared
 * vertices), grid -> world, quad split by the shorter diagonal, degenerate-triangle filter, area-weighted normals --
 * with the reference's vertex, quad and triangle ORDER.  The remaining clean-up of makeMesh / meshGenerateGeneric
 * (removeNonManifoldTriangles, meshClip, meshRemoveDisconnected, meshFlipNormals: marchingCubes.cpp:237-240,
 * sources/meshGeneration.cpp:41-42) stays with the caller (cage functions over these arrays).
 * All arrays are pinned host memory owned by the library until upg_mesh_free().
 */
typedef struct upg_mesh
{
	float *positions; /* vertex_count x 3, world units                 -> Mesh::positions(PointerRange<const Vec3>) */
	float *normals; /* vertex_count x 3                              -> Mesh::normals(...) */
	uint32_t *indices; /* index_count (3 per triangle)                  -> Mesh::indices(...) */
	float *dmc_vertices; /* vertex_count x 3, dualmc::Vertex in GRID units (parity with dualmc's own output) */
	uint32_t *dmc_quads; /* quad_count x 4, dualmc::Quad */
	uint32_t vertex_count;
	uint32_t index_count;
	uint32_t quad_count;
	uint32_t reserved;
	uint64_t near_zero_samples; /* samples with |value| < 1e-3: where sign / topology parity with a CPU run is not promised */
	void *owner; /* internal */
	uint64_t owner_bytes; /* internal */
} upg_mesh;

/* K1 + K2..K4.  Replaces the body of meshGenerateGeneric<FNC>() up to makeMesh's normals (sources/meshGeneration.cpp:32-40):
 * samples `field` over the planet's grid on the device and extracts the isosurface without the samples ever leaving HBM.
 * UPG_ERR_EMPTY_MESH when no triangle results (the reference throws "generated empty base land mesh", :51-52). */
int upg_field_mesh(upg_planet *planet, int field, upg_mesh *out);

/* The three base meshes of a planet in one call (replaces the three concurrent tasks meshGenerateBaseLand / Water / Navigation,
 * sources/generator.cpp:432-439): the fields of `field_mask` (bit 0 land, bit 1 water, bit 2 navigation) are sampled by one
 * fused K1 pass and extracted concurrently.  out3[field] is filled for every requested field (others zeroed); each mesh is
 * bit-identical to upg_field_mesh(field) and is released with upg_mesh_free. */
int upg_field_meshes(upg_planet *planet, unsigned field_mask, upg_mesh *out3);

/*
 * K9: the clean-up the reference applies to every extracted mesh, on the device, with the reference's vertex and triangle order.
 * `stages` is a combination of:
 */
enum
{
	/* the tail of MarchingCubes::makeMesh (externals/cage/sources/libcore/mesh/marchingCubes.cpp:237-240): removeNonManifoldTriangles
	 * (:28-96) and meshClip(mesh, box) (mesh/meshAlgorithms.cpp:1056-1114) including its removeUnusedVertices / meshRemoveInvalid /
	 * meshMergeCloseVertices({}) tail */
	UPG_MESH_MAKEMESH_TAIL = 1,
	UPG_MESH_REMOVE_DISCONNECTED = 2, /* meshRemoveDisconnected (meshAlgorithms.cpp:1313-1333; sources/meshGeneration.cpp:41) */
	UPG_MESH_FLIP_NORMALS = 4, /* meshFlipNormals (meshAlgorithms.cpp:313-326; sources/meshGeneration.cpp:42) */
	UPG_MESH_WATER_CULL = 8, /* meshGenerateBaseWater's vertex culling (sources/meshGeneration.cpp:62-95); water field only */
	UPG_MESH_GENERATOR = 15 /* everything: out3[] = what meshGenerateBaseLand / Water / Navigation return */
};

/* upg_field_meshes + K9 without the meshes leaving HBM in between: with UPG_MESH_GENERATOR the three results are the complete
 * bodies of meshGenerateBaseLand / meshGenerateBaseWater / meshGenerateBaseNavigation (sources/meshGeneration.cpp:47-107).
 * dmc_vertices / dmc_quads of a cleaned mesh are NULL.  UPG_ERR_EMPTY_MESH for an empty land or navigation mesh (the reference
 * throws for those two only); an empty water mesh is a valid result. */
int upg_field_meshes_clean(upg_planet *planet, unsigned field_mask, unsigned stages, upg_mesh *out3);

/* K9 alone over a caller's triangle mesh in host memory (positions / normals: vertex_count x 3, indices: 3 per triangle);
 * `box_size` is the clip box edge (the generator's 4000).  The water culling needs the planet: not available here. */
int upg_mesh_clean(int device, const float *positions, const float *normals, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, float box_size, unsigned stages, upg_mesh *out);

/* meshChunking of cage (externals/cage/sources/libcore/mesh/meshAlgorithms.cpp:1484-1619) as the generator's meshSplit calls it
 * (sources/meshGeneration.cpp:173-182, maxSurfaceArea 800000): the mesh is split recursively -- longest axis of the bounding box, the best of
 * seven candidate planes by the surface area below each, meshClip of a copy against the two (20 % enlarged) half boxes -- until every part's
 * surface area is below `max_surface_area`.  `*chunks` = the parts in the reference's order (first half before second half, depth first), each
 * with positions, normals and indices (dmc_* NULL); parts the clip left empty stay in the list with zero counts, as in the reference.  An empty
 * input mesh gives zero chunks.  Free with upg_mesh_chunks_free.  Host arrays in, host arrays out; SURVEY.md 8f n4.
 * STATUS (round 2): written after the round's GPU time was spent -- compiled, never run on a B200; see DESIGN.md "not verified on hardware". */
int upg_mesh_chunking(int device, const float *positions, const float *normals, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, float max_surface_area, upg_mesh **chunks, uint32_t *chunk_count);
void upg_mesh_chunks_free(upg_mesh *chunks, uint32_t chunk_count);

/*
 * Multi-GPU sharding of K1 + K2..K4 by z-slab ("grid sampling shards naturally by z-slab").  dualmc's scan over z in [0, res-2)
 * (dualmc.tpp:401) is split into `slab_count` contiguous ranges (upg_slab_range); slab k samples its slices plus a 3-slice halo on
 * each side ON ITS OWN DEVICE (recomputed, not exchanged) and returns the quads / triangles its scan iterations emit and the
 * vertices they create.  Vertex ids in `indices` / `dmc_quads` are RELATIVE to the slab's first vertex, as two's-complement
 * uint32 (ids below zero refer to the last vertices of slab k-1): add the total vertex_count of slabs 0..k-1 (mod 2^32) and
 * concatenate the arrays of all slabs in order -- the result is bit-identical to upg_field_mesh on one device, with the
 * reference's vertex / quad / triangle order.  The only exchange a multi-GPU caller needs is that count and the final gather.
 */
void upg_slab_range(int grid_resolution, int slab, int slab_count, int *z0, int *z1);
int upg_field_mesh_slab(upg_planet *planet, int field, int slab, int slab_count, upg_mesh *out);

/* all requested fields of one slab from ONE fused K1 pass over the slab's slices (upg_field_meshes restricted to slab k): each
 * out3[field] equals upg_field_mesh_slab(field, slab, slab_count) */
int upg_field_meshes_slab(upg_planet *planet, unsigned field_mask, int slab, int slab_count, upg_mesh *out3);

/* the same with the result left in HBM, for callers that gather the slab chunks device to device (NCCL over NVLink) instead of
 * through host memory.  Pointers are device memory on `device`, valid until upg_device_mesh_free(); the work that filled them has
 * completed when the call returns. */
typedef struct upg_device_mesh
{
	float *positions; /* vertex_count x 3 */
	float *normals; /* vertex_count x 3 */
	uint32_t *indices; /* index_count, relative vertex ids like upg_field_mesh_slab */
	float *dmc_vertices; /* vertex_count x 3 */
	uint32_t *dmc_quads; /* quad_count x 4 */
	uint32_t vertex_count;
	uint32_t index_count;
	uint32_t quad_count;
	int32_t device;
	uint64_t near_zero_samples;
	void *owner; /* internal */
} upg_device_mesh;
int upg_field_meshes_slab_device(upg_planet *planet, unsigned field_mask, int slab, int slab_count, upg_device_mesh *out3);
void upg_device_mesh_free(upg_device_mesh *mesh); /* stream-ordered on the device's default stream: finish your own reads first */

/* K2..K4 alone over caller-provided samples (host memory, layout [z][y][x]): MarchingCubes::densities(values) followed by
 * makeMesh() (include/cage-core/marchingCubes.h:16,27).  `full_resolution` / `box_size` define the grid -> world mapping
 * (MarchingCubesCreateConfig::position); pass rx = ry = rz = full_resolution for a whole grid. */
int upg_isosurface(const float *host_samples, int rx, int ry, int rz, float box_size, int full_resolution, int device, upg_mesh *out);

void upg_mesh_free(upg_mesh *mesh);

/* Host memory (pageable or pinned) -> device memory of `device`, complete when the call returns: the multi-threaded staged copy the
 * host-buffer entry points use for their mesh arguments (pageable memory moves at ~11 GB/s through cudaMemcpy, at link speed through
 * this), for callers that feed the *_device entry points from host arrays (unnatural_planets_b200/sharded.py).  Several arrays in one
 * call share the staging: `count` triples of (device_dst[i], host_src[i], bytes[i]), count <= 16. */
int upg_upload(int device, void *const *device_dst, const void *const *host_src, const uint64_t *bytes, int count, void *cuda_stream);

/*
 * ---- texel baking (K5..K7) ------------------------------------------------------------------------------------
 */

/* MeshPurposeEnum (sources/planets.h:49-55): which terrainSdf* feeds tile.elevation and which layers run */
enum
{
	UPG_PURPOSE_LAND = 1,
	UPG_PURPOSE_WATER = 2,
	UPG_PURPOSE_NAVIGATION = 3
};

/* what terrainTile() leaves in a Tile (sources/planets.h:63-83), the fields the generator reads afterwards */
typedef struct upg_tile
{
	float albedo[3];
	float roughness;
	float metallic;
	float height;
	float elevation;
	float slope; /* radians */
	float temperature;
	float precipitation;
	float opacity;
	float biome; /* TerrainBiomeEnum as a number */
	float type; /* TerrainTypeEnum as a number */
} upg_tile;

/*
 * The Cellular generator of FastNoiseLite indexes a 256 x 4 float constant table (RandVecs3D).  The library source is an
 * un-vendored submodule of the reference (externals/cage/externals/fastnoise), so the table is DATA here: a planet starts
 * with the declared substitute (unit vectors from a fixed splitmix64 stream, identical to oracle/up_noise.h) and an
 * integration that has the real header passes upstream's table once per planet.  1024 floats, entry e at [4e .. 4e+2].
 */
int upg_planet_set_cellular_table(upg_planet *planet, const float *table1024);

/* terrainTile() (sources/modesConfiguration.cpp:281-286) for n explicit position / unit-normal pairs: the per-vertex loop of
 * generateTileProperties (sources/tileProperties.cpp:258-265, purpose NAVIGATION) and the parity probe of the colouring. */
int upg_tiles(upg_planet *planet, int purpose, const float *host_positions, const float *host_normals, uint64_t n, upg_tile *host_out);

/*
 * K5 + K6 + K7.  Replaces Generator<Water>::generate (sources/textures.cpp:45-82) behind generateTexturesLand / Water
 * (:86-96): meshGenerateTexture's scan-line coverage with "the last triangle wins" (externals/cage/sources/libcore/mesh/
 * meshAlgorithms.cpp:1676-1729), one terrainTile per covered texel (Generator::pixel, textures.cpp:24-43), seven NaN-aware
 * dilation rounds per image (imageDilation(img, 7, true), libcore/image/imageAlgorithms.cpp:349-527) and the truncating
 * 8-bit quantisation of imageConvert(U8) (image.cpp:244-262).
 * Mesh arrays are what Mesh::positions() / normals() / uvs() / indices() expose (include/cage-core/mesh.h:34-43,86-88).
 * Outputs are tightly packed rows, top row first as cage stores them: albedo w*h*3 (land) or w*h*4 (water: rgb + opacity),
 * special w*h*2 (roughness, metallic), height w*h*1.  Uncovered texels that no dilation round reaches are 0.
 */
int upg_bake(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, uint32_t width, uint32_t height, uint8_t *albedo, uint8_t *special, uint8_t *height_map);

/* Rows [row0, row1) of the same textures: the unit of multi-GPU sharding ("texture baking shards by texel-row tiles").  Texels
 * are independent; the seven dilation rounds reach 7 rows, so the call bakes a band with a 7-row apron internally and returns
 * exactly the rows a whole-image upg_bake would.  Outputs hold (row1 - row0) rows. */
int upg_bake_rows(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, uint32_t width, uint32_t height, uint32_t row0, uint32_t row1, uint8_t *albedo, uint8_t *special, uint8_t *height_map);

/* same with the mesh and the outputs in DEVICE memory, on `cuda_stream` (synchronises that stream before returning): for callers
 * that keep the mesh on the GPU (bench.py's device-resident metric) */
int upg_bake_device(upg_planet *planet, int water, const float *device_positions, const float *device_normals, const float *device_uvs, uint32_t vertex_count, const uint32_t *device_indices, uint32_t index_count, uint32_t width, uint32_t height, uint32_t row0, uint32_t row1, uint8_t *device_albedo, uint8_t *device_special, uint8_t *device_height_map, void *cuda_stream);

/*
 * K5..K8: the bake plus the texture tail the generator applies to every chunk before it writes the PNGs
 * (sources/generator.cpp:123-145, 175-197): imageConvertSpecialToGltfPbr (externals/cage/sources/libcore/image/
 * imageAlgorithms.cpp:296-302) -> `pbr` w*h*3 = (0, roughness, metallic), and imageConvertHeigthToNormal(heightMap, 1)
 * (:202-246) -> `normal_map` w*h*3, a 3x3 Sobel over the U8 height with clamped borders.  `albedo`, `pbr` and `normal_map` are
 * the payloads of <chunk>-albedo.png, -pbr.png and -normal.png.  The *_rows / *_device forms shard by texel rows like upg_bake_rows
 * (the band grows to 8 rows of apron: the Sobel reads one finished row beyond the caller's).
 */
int upg_bake_export(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, uint32_t width, uint32_t height, uint8_t *albedo, uint8_t *pbr, uint8_t *normal_map);
int upg_bake_export_rows(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, uint32_t width, uint32_t height, uint32_t row0, uint32_t row1, uint8_t *albedo, uint8_t *pbr, uint8_t *normal_map);
int upg_bake_export_device(upg_planet *planet, int water, const float *device_positions, const float *device_normals, const float *device_uvs, uint32_t vertex_count, const uint32_t *device_indices, uint32_t index_count, uint32_t width, uint32_t height, uint32_t row0, uint32_t row1, uint8_t *device_albedo, uint8_t *device_pbr, uint8_t *device_normal_map, void *cuda_stream);

/* K8 alone over host U8 images: `special` (w*h*2) -> `pbr` (w*h*3) and / or `height_map` (w*h*1) -> `normal_map` (w*h*3) with
 * imageConvertHeigthToNormal's `strength` (> 0; the generator passes 1).  Either pair may be NULL. */
int upg_texture_tail(int device, const uint8_t *special, const uint8_t *height_map, uint32_t width, uint32_t height, float strength, uint8_t *pbr, uint8_t *normal_map);

/*
 * K10: PNG encoding of an 8-bit image on the device -- what Image::exportFile(".png") does for the three textures of a chunk right
 * after the texture tail (sources/generator.cpp:135-139, 187-191; externals/cage/sources/libcore/image/png.cpp -> libpng + zlib).
 * `pixels`: width x height x channels bytes, tightly packed rows, host OR device memory of `device` (channels 1 grey, 2 grey + alpha,
 * 3 RGB, 4 RGBA).  `out` (host memory, `capacity` >= upg_png_bound()) receives a complete PNG file of *size bytes: signature, IHDR,
 * one IDAT chunk per 8 KiB deflate block, IEND.  Scanlines are filtered with libpng's minimum-sum heuristic; the zlib stream uses fixed
 * Huffman codes over an 8 KiB window (faster and a little larger than zlib's default level).  Any PNG decoder returns the pixels bit
 * for bit; the file bytes are not those of libpng.
 */
uint64_t upg_png_bound(uint32_t width, uint32_t height, uint32_t channels);
int upg_png_encode(int device, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t channels, uint8_t *out, uint64_t capacity, uint64_t *size);

/* K5..K8 + K10 in one call: what a chunk task does between meshUnwrap and the end of its three exportFile calls (sources/generator.cpp:
 * 131-139): bake, texture tail, and the PNG FILES of albedo (RGB / RGBA for water), pbr and normal map.  out3[k] (host memory, capacity3[k]
 * >= upg_png_bound(width, height, channels)) receives file k, size3[k] its length.  The images exist uncompressed only in HBM. */
int upg_bake_export_png(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, uint32_t width, uint32_t height, uint8_t *const out3[3], const uint64_t capacity3[3], uint64_t size3[3]);

/* coverage statistics of the last upg_bake on this thread: texels written by the rasteriser (= terrainTile evaluations;
 * the reference evaluates every (triangle, texel) visit, this library only the winning one) */
uint64_t upg_bake_last_covered_texels(void);

/* K5 alone, for parity: owner[y*w+x] = index of the LAST triangle whose scan covers the texel (0xffffffff = none) and the
 * barycentric weights of that visit (meshAlgorithms.cpp:1719-1724); `callbacks` = number of (triangle, texel) visits */
int upg_raster(int device, const float *uvs, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, uint32_t width, uint32_t height, uint32_t *owner, float *weights /* w*h*3 */, uint64_t *callbacks);

/* K7 alone: imageDilation(img, rounds, use_nan) on a float image (in place) and, when u8_out != NULL, imageConvert(U8)
 * of the result (imageAlgorithms.cpp:504-527, 100-113).  channels 1..4. */
int upg_dilate(int device, float *image, uint32_t width, uint32_t height, uint32_t channels, uint32_t rounds, int use_nan, uint8_t *u8_out);

/* measurement helper for bench.py: the FP32 FMA ceiling (TFLOP/s, FMA = 2 flop) of `device`, measured live with a
 * dependent-free FFMA kernel -- the roofline denominator of the compute-bound kernels (K1 sampling, K6 texel baking);
 * MEASURED_PEAKS.json only carries HBM and tensor-core peaks */
int upg_measure_fp32_peak(int device, float *tflops_out);

/* bytes of per-planet input a sampling launch carries from the host to the device: the planet's parameter block (mode ids, seeds,
 * prepared shape constants) and the launch descriptor travel as __grid_constant__ kernel arguments -- a generator has no other per-step
 * input (bench.py declares them as the host -> device bytes of a sampling step) */
uint32_t upg_kernel_param_bytes(void);

/* thread-local description of the last failure on this thread ("" if none) */
const char *upg_last_error(void);

#ifdef __cplusplus
}
#endif

#endif /* UPG_H_INCLUDED */
