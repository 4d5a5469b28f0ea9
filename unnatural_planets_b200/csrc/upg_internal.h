// upg_internal.h -- host-side objects shared by the translation units of libupg.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <vector>

#include "../../include/upg.h"
#include "upg_field.cuh"
#include "upg_mem.h"
#include "upg_planet_host.h"

namespace upg
{
	int fail(int code, const char *fmt, ...);
	int cuda_fail(cudaError_t e, const char *what);

	// what upg_planet* points to; immutable after upg_planet_create except for the device status word
	struct Planet
	{
		PlanetParams params;
		ColoringConsts coloring;
		SeedLog seedLog;
		float *cellularTable = nullptr; // device: 1024 floats (FastNoiseLite RandVecs3D or its declared substitute)
		float boxSize = 4000;
		int device = 0;
		int sampling = 0; // UPG_SAMPLING_* of include/upg.h
		int arithmetic = 0; // UPG_ARITH_* of include/upg.h: which build of the dense K1 kernels the sampling entry points launch
		int deviceCount = 1; // devices that share the planet's work in the multi-device entry points
		int devices[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; // devices[0] == device
		// multi-device planets: one single-device planet per further device (same modes, seeds and table; own status words and
		// cellular table in that device's memory); children[0] is the planet itself
		Planet *children[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };
		// device: STATUS_SLOTS + 1 words, bit 0 of each = "a NaN sample was produced".  Word 0 is the sticky flag of the ASYNCHRONOUS
		// device-pointer entry points (read and cleared by upg_planet_async_status); words 1.. are handed out one per synchronous
		// call (StatusSlot), so that concurrent calls on one planet -- the generator's three field tasks and its chunk bakes,
		// sources/generator.cpp:432-439, :159 -- never report or clear each other's flag.
		uint32_t *status = nullptr;
		std::atomic<uint64_t> statusSlots{ 0 }; // bit i: word i + 1 is in use
	};
	constexpr int STATUS_SLOTS = 64;

	// the status word kernels of the current call flag NaNs in: the calling thread's slot, or the planet's sticky word
	uint32_t *status_word(const Planet *planet);
	// one per synchronous entry point: takes a free word of the planet, zeroes it on `s` and makes it the calling thread's word until
	// the object dies.  Worker threads of the same call adopt() the parent's word.
	struct StatusSlot
	{
		Planet *planet = nullptr;
		int index = -1;
		uint32_t *previous = nullptr;
		bool installed = false;
		int open(Planet *p, cudaStream_t s);
		void adopt(uint32_t *word);
		uint32_t *word() const;
		~StatusSlot();
	};

	// a private pair of streams per API call: calls arrive concurrently from the reference's pool threads
	// (sources/generator.cpp:432-439, :159) and must not serialise on the default stream
	struct CallStreams
	{
		cudaStream_t compute = nullptr;
		cudaStream_t copy = nullptr;
		int open()
		{
			cudaError_t e = cudaStreamCreateWithFlags(&compute, cudaStreamNonBlocking);
			if (e == cudaSuccess)
				e = cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking);
			if (e != cudaSuccess)
			{
				close();
				return cuda_fail(e, "cudaStreamCreate");
			}
			return 0;
		}
		void close()
		{
			if (compute)
				cudaStreamDestroy(compute);
			if (copy)
				cudaStreamDestroy(copy);
			compute = copy = nullptr;
		}
	};

	// device-side result of the isosurface stage (upg_dmc.cu); arrays live in the stream-ordered pool
	struct DmcResult
	{
		float *rawVertices = nullptr; // dualmc vertices in grid units (x, y, z)
		float *positions = nullptr; // world positions
		float *normals = nullptr; // area-weighted, normalised
		uint32_t *indices = nullptr; // 3 per triangle
		uint32_t *quads = nullptr; // 4 per quad: dualmc's own output, before the triangle split
		uint32_t vertexCount = 0, quadCount = 0, triangleCount = 0;
		// slab extraction (dmc_extract with an owned z range): the blocks of the arrays above that belong to the slab; `quads`
		// and `indices` of those blocks hold vertex ids relative to firstVertex (two's complement)
		bool slab = false;
		uint32_t firstVertex = 0, ownedVertices = 0, firstQuad = 0, ownedQuads = 0, firstTriangle = 0, ownedTriangles = 0;
		uint64_t nearZeroSamples = 0; // samples with |value| < 1e-3: where topology parity is not promised
	};
	// What the dense passes of the extraction need of a sample field, as bits (one per node, rows of `pitch` words, bit x & 31 of word
	// x >> 5, at least one all-zero padding word behind the last sample of a row): `signs` = sample >= 0 (getCellCode's inside test),
	// `near` = |sample| < 1e-3 (DmcResult::nearZeroSamples).  Written by the samplers (shape_rows_kernel, band_sample_kernel) so that
	// the cell-code pass reads 1 bit instead of 4 bytes per node and the band sampler need not store what nobody reads.
	struct SignBits
	{
		uint32_t *signs = nullptr;
		uint32_t *near = nullptr;
		uint32_t pitch = 0;
	};
	inline uint32_t sign_bits_pitch(int res) { return (uint32_t)((res + 31) / 32 + 1); }
	constexpr float NEAR_ZERO_TOLERANCE = 1e-3f;

	// zOffset: global z of local slice 0; [ownZ0, ownZ1): global scan iterations (dualmc's z loop) whose quads / created vertices /
	// triangles are reported (defaults = the whole volume)
	int dmc_extract(const float *dens, int rx, int ry, int rz, float boxSize, int fullRes, cudaStream_t s, DmcResult &out, int zOffset = 0, int ownZ0 = 0, int ownZ1 = 0x7fffffff, const SignBits *bits = nullptr);
	void dmc_release(DmcResult &r, cudaStream_t s);

	// K9 (upg_meshfix.cu): the clean-up chain of makeMesh / meshGenerateGeneric / meshGenerateBaseWater over device arrays; `stages` are
	// the UPG_MESH_* bits of include/upg.h.  `out` receives fresh pool arrays (positions, normals, indices; no dualmc arrays).
	int mesh_clean(Planet *planet, const float *positions, const float *normals, uint32_t verts, const uint32_t *indices, uint32_t tris, float boxSize, unsigned stages, cudaStream_t s, DmcResult &out, const float *clipBox6 = nullptr);
	// internal stage of mesh_clean: meshClip(mesh, box) ALONE with the box of `clipBox6` (a.x a.y a.z b.x b.y b.z) -- clip, removeUnusedVertices,
	// meshRemoveInvalid, meshMergeCloseVertices({}) (meshAlgorithms.cpp:1056-1114) -- as meshChunking applies it to every half
	constexpr unsigned MESH_STAGE_CLIP_BOX = 0x100;
	// meshChunking (meshAlgorithms.cpp:1522-1619) over a device mesh; leaves in the reference's order (upg_meshfix.cu)
	int mesh_chunking(const float *positions, const float *normals, uint32_t verts, const uint32_t *indices, uint32_t tris, float maxSurfaceArea, cudaStream_t s, std::vector<DmcResult> &leaves);

	// K1b (upg_band.cu): land / navigation of slices [za, zb) from the water (= shape) field of the same slices, exact wherever the
	// isosurface can see the value; asynchronous on `stream`.  `evaluated`: optional device counter of evaluated nodes.
	bool band_sampling_applies(const Planet *planet);
	// internal status code of check_status: an evaluated elevation was outside elevation_interval10 (band_sample_kernel's guard).  The
	// mesh entry points catch it and repeat the call with dense sampling; it never reaches the caller.
	constexpr int UPG_ERR_BAND_INTERVAL = 1000;
	bool force_dense_sampling();
	void set_force_dense_sampling(bool on);
	// With `landBits` / `navigationBits` (zero-filled by the caller) the kernel also writes the sign / near-zero bits of both fields and,
	// unless `fill`, stores sample VALUES only where it evaluated them (all that dmc_build reads).
	int band_sample(const Planet *planet, int za, int zb, const float *water, float *land, float *navigation, unsigned long long *evaluated, cudaStream_t stream, const SignBits *landBits = nullptr, const SignBits *navigationBits = nullptr, bool fill = true, const float2 *rowMinMax = nullptr);

	// cached pinned host blocks (upg_api.cu) and the staged upload of pageable host arrays built on them
	cudaError_t pinned_block_acquire(size_t bytes, void **out, size_t *capacity);
	void pinned_block_release(void *ptr, size_t capacity);
	struct HostArray
	{
		void *dst; // device
		const void *src; // host, pageable or pinned
		size_t bytes;
	};
	int upload_host_arrays(const HostArray *arrays, int n, cudaStream_t stream, void **staging, size_t *stagingCapacity);

	// Waits for a stream.  cudaStreamSynchronize spins; with one process per GPU (bench.py --gpus 8: eight ranks, each with its three
	// extraction threads and two bake threads) the spinning threads outnumber the host cores and starve the threads that launch work,
	// so when several GPUs are visible the wait goes through an event created with cudaEventBlockingSync and the thread sleeps instead.
	// UPG_BLOCKING_SYNC=0 / 1 overrides the choice.
	cudaError_t stream_sync(cudaStream_t s);

	bool trace_enabled(); // UPG_TRACE=1
	double now_ms();

	int launch_field(const Planet *planet, const FieldLaunch &L, cudaStream_t stream);

	// ---- pieces of the single-device entry points the multi-device layer (upg_multi.cu) composes -------------------------------
	// upg_field_meshes_slab_device for an internal planet pointer (slabCount 0 = the whole grid)
	int field_meshes_slab_device(Planet *planet, unsigned field_mask, int slab, int slabCount, upg_device_mesh *out3);
	// device-side result -> one pinned host block + the ABI struct (synchronises `s`)
	int mesh_to_host(DmcResult &r, cudaStream_t s, upg_mesh *out);
	// ids[i] += add (mod 2^32) over n ids on `s`: slab-relative vertex ids -> ids of the stitched mesh
	int rebase_ids(uint32_t *ids, uint64_t n, uint32_t add, cudaStream_t s);
	// rows [row0, row1) of a bake with the mesh in device memory of the CURRENT device; outputs are host or device memory and hold
	// (row1 - row0) rows (upg_bake.cu); synchronises `cuda_stream`
	int bake_core(upg_planet *planet, int water, const float *dpos, const float *dnrm, const float *duv, uint32_t nverts, const uint32_t *didx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t row0, uint32_t row1, uint8_t *dAlbedo, uint8_t *dSpecial, uint8_t *dHeight, uint8_t *dPbr, uint8_t *dNormal, float normalStrength, void *cuda_stream);
	uint64_t bake_thread_covered();
	void bake_set_thread_covered(uint64_t n);
	int check_bake_mesh(const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h);

	// multi-device layer (upg_multi.cu): planets created with upg_planet_desc::device_count > 1
	int multi_setup(Planet *planet, const upg_planet_desc *desc);
	void multi_teardown(Planet *planet);
	int multi_field_meshes(Planet *planet, unsigned field_mask, unsigned stages, upg_mesh *out3);
	int multi_bake(Planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t row0, uint32_t row1, uint8_t *albedo, uint8_t *special, uint8_t *heightMap, uint8_t *pbr, uint8_t *normalMap);
	int check_status(Planet *planet, cudaStream_t stream, const char *what);
}
