// upg_api.cu -- the C ABI (include/upg.h) over the CUDA kernels.  Host logic only: mode tables, seed
// stream, parameter marshalling, stream / memory management, error codes.  No compute happens on the CPU.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <atomic>
#include <vector>

#include "../../include/upg.h"
#include "upg_internal.h"

namespace upg
{
	// ---- error plumbing ---------------------------------------------------------------------------
	static thread_local std::string t_lastError;

	int fail(int code, const char *fmt, ...)
	{
		char buf[512];
		va_list ap;
		va_start(ap, fmt);
		vsnprintf(buf, sizeof(buf), fmt, ap);
		va_end(ap);
		t_lastError = buf;
		return code;
	}

	int cuda_fail(cudaError_t e, const char *what)
	{
		if (e == cudaErrorMemoryAllocation && last_failed_request().device >= 0)
		{
			const FailedRequest f = last_failed_request();
			last_failed_request().device = -1; // reported once
			return fail(UPG_ERR_CUDA, "CUDA error in %s: %s (%s); last request the pool of device %d could not serve: %llu bytes (pool holds %llu, %llu in use; device %llu free of %llu)", what, cudaGetErrorName(e), cudaGetErrorString(e), f.device,
				(unsigned long long)f.bytes, (unsigned long long)f.reserved, (unsigned long long)f.used, (unsigned long long)f.deviceFree, (unsigned long long)f.deviceTotal);
		}
		return fail(UPG_ERR_CUDA, "CUDA error in %s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
	}

	// ---- mode tables, in the reference's order (sources/modesConfiguration.cpp:78-113, 176-183, 207-211) ----
	static const char *const shapeNames[SH_COUNT] = {
		"asteroid", "belt", "bowl", "box", "bunny", "capsule", "cube", "disk", "doubletorus", "fibers", "gear", "h2o", "h3o", "h4o", "hemispheres", "hexagon", "hexagonalprism", "insidecube", "knot", "mandelbulb", "mobiusstrip", "monkeyhead", "octahedron", "pipe", "sphere", "square", "tetrahedron", "torus", "toruscross", "triangularprism", "tube", "twistedhexagonalprism", "twistedplane", "wormhole",
	};
	static const char *const elevationNames[EL_COUNT] = { "none", "simple", "legacy", "lakes", "islands", "craters" };
	static const char *const coloringNames[CL_COUNT] = { "default", "barren", "debug" };

	static int lookup(const char *const *table, int n, const char *name)
	{
		if (!name)
			return -1;
		for (int i = 0; i < n; i++)
			if (std::strcmp(table[i], name) == 0)
				return i;
		return -1;
	}

	// ---- K1 dispatch: one launcher per (elevation, shape range) translation unit ----------------------------
#define UPG_DECL_LAUNCHERS(E) \
	bool field_launch_e##E##_p0(int, const PlanetParams &, const FieldLaunch &, void *); \
	bool field_launch_e##E##_p1(int, const PlanetParams &, const FieldLaunch &, void *); \
	bool field_launch_e##E##_p2(int, const PlanetParams &, const FieldLaunch &, void *); \
	bool field_launch_e##E##_p3(int, const PlanetParams &, const FieldLaunch &, void *); \
	bool field_launchc_e##E##_p0(int, const PlanetParams &, const FieldLaunch &, void *); \
	bool field_launchc_e##E##_p1(int, const PlanetParams &, const FieldLaunch &, void *); \
	bool field_launchc_e##E##_p2(int, const PlanetParams &, const FieldLaunch &, void *); \
	bool field_launchc_e##E##_p3(int, const PlanetParams &, const FieldLaunch &, void *);
	UPG_DECL_LAUNCHERS(0)
	UPG_DECL_LAUNCHERS(1)
	UPG_DECL_LAUNCHERS(2)
	UPG_DECL_LAUNCHERS(3)
	UPG_DECL_LAUNCHERS(4)
	UPG_DECL_LAUNCHERS(5)
#define UPG_LIST_LAUNCHERS(E) { field_launch_e##E##_p0, field_launch_e##E##_p1, field_launch_e##E##_p2, field_launch_e##E##_p3 }
	static const FieldLauncher fieldLaunchers[EL_COUNT][4] = {
		UPG_LIST_LAUNCHERS(0), UPG_LIST_LAUNCHERS(1), UPG_LIST_LAUNCHERS(2), UPG_LIST_LAUNCHERS(3), UPG_LIST_LAUNCHERS(4), UPG_LIST_LAUNCHERS(5),
	};
	// the same kernels compiled with FMA contraction (UPG_ARITH_CONTRACT): dense grid launches of the sampling entry points only
#define UPG_LIST_LAUNCHERS_C(E) { field_launchc_e##E##_p0, field_launchc_e##E##_p1, field_launchc_e##E##_p2, field_launchc_e##E##_p3 }
	static const FieldLauncher fieldLaunchersContract[EL_COUNT][4] = {
		UPG_LIST_LAUNCHERS_C(0), UPG_LIST_LAUNCHERS_C(1), UPG_LIST_LAUNCHERS_C(2), UPG_LIST_LAUNCHERS_C(3), UPG_LIST_LAUNCHERS_C(4), UPG_LIST_LAUNCHERS_C(5),
	};

	int launch_field(const Planet *planet, const FieldLaunch &L, cudaStream_t stream)
	{
		if (L.count == 0)
			return UPG_OK;
		const PlanetParams &P = planet->params;
		// the water field is the bare shape: its grid launches take the row-wise kernel that lives with the elevation-free instantiations
		const bool waterRows = L.field == FD_WATER && !L.positions;
		const int table = waterRows ? (int)EL_NONE : P.elevation;
		const bool contract = L.contract && planet->arithmetic == UPG_ARITH_CONTRACT && !waterRows && !L.positions;
		bool done = false;
		for (int part = 0; part < 4 && !done; part++)
			done = (contract ? fieldLaunchersContract : fieldLaunchers)[table][part](P.shape, P, L, (void *)stream);
		if (!done)
			return fail(UPG_ERR_ARGUMENT, "no kernel for shape %d elevation %d", P.shape, P.elevation);
		const cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess)
			return cuda_fail(e, "sdf_sample_kernel launch");
		return UPG_OK;
	}

	static thread_local uint32_t *t_statusWord = nullptr;
	static thread_local bool t_forceDense = false; // set while an entry point repeats a call whose band premise failed

	bool force_dense_sampling() { return t_forceDense; }
	void set_force_dense_sampling(bool on) { t_forceDense = on; }

	uint32_t *status_word(const Planet *planet)
	{
		return t_statusWord ? t_statusWord : planet->status;
	}

	int StatusSlot::open(Planet *p, cudaStream_t s)
	{
		planet = p;
		// a free bit of the mask; more than STATUS_SLOTS concurrent calls on one planet wait for one to finish
		while (index < 0)
		{
			uint64_t mask = p->statusSlots.load();
			if (~mask == 0)
			{
				std::this_thread::yield();
				continue;
			}
			const int bit = __builtin_ctzll(~mask);
			if (p->statusSlots.compare_exchange_weak(mask, mask | (1ull << bit)))
				index = bit;
		}
		const cudaError_t e = cudaMemsetAsync(p->status + 1 + index, 0, sizeof(uint32_t), s);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaMemsetAsync(status)");
		adopt(p->status + 1 + index);
		return UPG_OK;
	}

	void StatusSlot::adopt(uint32_t *word)
	{
		previous = t_statusWord;
		t_statusWord = word;
		installed = true;
	}

	uint32_t *StatusSlot::word() const
	{
		return index >= 0 ? planet->status + 1 + index : nullptr;
	}

	StatusSlot::~StatusSlot()
	{
		if (installed)
			t_statusWord = previous;
		if (index >= 0)
			planet->statusSlots.fetch_and(~(1ull << index));
	}

	int check_status(Planet *planet, cudaStream_t stream, const char *what)
	{
		uint32_t st = 0;
		uint32_t *word = status_word(planet);
		cudaError_t e = cudaMemcpyAsync(&st, word, sizeof(st), cudaMemcpyDeviceToHost, stream);
		if (e == cudaSuccess)
			e = stream_sync(stream);
		if (e != cudaSuccess)
			return cuda_fail(e, what);
		if (st & 3u)
		{
			cudaMemsetAsync(word, 0, sizeof(uint32_t), stream);
			stream_sync(stream);
			if (st & 1u)
				return fail(UPG_ERR_INVALID_VALUE, "invalid sdf value (NaN) in %s", what);
			return fail(UPG_ERR_BAND_INTERVAL, "an elevation left the interval band sampling assumes (%s)", what);
		}
		return UPG_OK;
	}

	static bool valid_field(int field)
	{
		return field >= 0 && field < FD_COUNT;
	}
}

using namespace upg;

extern "C"
{
	const char *upg_last_error(void)
	{
		return t_lastError.c_str();
	}

	void upg_planet_desc_default(upg_planet_desc *d)
	{
		if (!d)
			return;
		std::memset(d, 0, sizeof(*d));
		d->struct_size = sizeof(*d);
		d->shape = SH_SPHERE;
		d->elevation = EL_LAKES;
		d->coloring = CL_DEFAULT;
		d->poles = 0;
		d->flowers = 0;
		d->seed[0] = 1346519564496ull;
		d->seed[1] = 42245614964156ull;
		d->box_size = 4000.0f;
		d->grid_resolution = 500;
		d->device = 0;
	}

	uint32_t upg_kernel_param_bytes(void)
	{
		return (uint32_t)(sizeof(PlanetParams) + sizeof(FieldLaunch));
	}

	int upg_shape_id(const char *name)
	{
		return lookup(shapeNames, SH_COUNT, name);
	}
	int upg_elevation_id(const char *name)
	{
		return lookup(elevationNames, EL_COUNT, name);
	}
	int upg_coloring_id(const char *name)
	{
		return lookup(coloringNames, CL_COUNT, name);
	}
	const char *upg_shape_name(int id)
	{
		return id >= 0 && id < SH_COUNT ? shapeNames[id] : nullptr;
	}
	const char *upg_elevation_name(int id)
	{
		return id >= 0 && id < EL_COUNT ? elevationNames[id] : nullptr;
	}
	const char *upg_coloring_name(int id)
	{
		return id >= 0 && id < CL_COUNT ? coloringNames[id] : nullptr;
	}
	int upg_shape_count(void)
	{
		return SH_COUNT;
	}

	int upg_planet_create(const upg_planet_desc *d, upg_planet **out)
	{
		if (!d || !out)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		*out = nullptr;
		if (d->struct_size != sizeof(upg_planet_desc))
			return fail(UPG_ERR_ARGUMENT, "upg_planet_desc::struct_size mismatch (%u, expected %zu)", d->struct_size, sizeof(upg_planet_desc));
		if (d->shape < 0 || d->shape >= SH_COUNT)
			return fail(UPG_ERR_ARGUMENT, "unknown shape mode configuration (%d)", d->shape);
		if (d->elevation < 0 || d->elevation >= EL_COUNT)
			return fail(UPG_ERR_ARGUMENT, "unknown elevation mode configuration (%d)", d->elevation);
		if (d->coloring < 0 || d->coloring >= CL_COUNT)
			return fail(UPG_ERR_ARGUMENT, "unknown coloring mode configuration (%d)", d->coloring);
		// (5: marchingCubes.cpp:21-23; upper bound: the isosurface stage indexes cells with 32 bits, upg_dmc.cu)
		if (d->grid_resolution <= 5 || d->grid_resolution > UPG_MAX_GRID_RESOLUTION)
			return fail(UPG_ERR_ARGUMENT, "grid resolution %d outside (5, %d]", d->grid_resolution, UPG_MAX_GRID_RESOLUTION);
		if (d->seed[0] == 0 || d->seed[1] == 0)
			return fail(UPG_ERR_ARGUMENT, "seed words must be non-zero");
		if (!(d->box_size > 0))
			return fail(UPG_ERR_ARGUMENT, "box size must be positive");
		if (d->mesh_sampling != UPG_SAMPLING_BAND && d->mesh_sampling != UPG_SAMPLING_DENSE)
			return fail(UPG_ERR_ARGUMENT, "unknown mesh sampling %d", d->mesh_sampling);
		if (d->arithmetic != UPG_ARITH_EXACT && d->arithmetic != UPG_ARITH_CONTRACT)
			return fail(UPG_ERR_ARGUMENT, "unknown arithmetic mode %d", d->arithmetic);
		if (d->device_count < 0 || d->device_count > UPG_MAX_DEVICES)
			return fail(UPG_ERR_ARGUMENT, "device count %d outside [0, %d]", d->device_count, UPG_MAX_DEVICES);
		if (d->device_count > 0 && d->devices[0] != d->device)
			return fail(UPG_ERR_ARGUMENT, "devices[0] (%d) must be the planet's home device (%d)", d->devices[0], d->device);
		for (int i = 0; i < d->device_count; i++)
			for (int j = 0; j < i; j++)
				if (d->devices[i] == d->devices[j])
					return fail(UPG_ERR_ARGUMENT, "device %d listed twice", d->devices[i]);

		Planet *p = new (std::nothrow) Planet();
		if (!p)
			return fail(UPG_ERR_ARGUMENT, "out of memory");
		std::memset(&p->params, 0, sizeof(p->params));
		p->params.shape = d->shape;
		p->params.elevation = d->elevation;
		p->params.coloring = d->coloring;
		p->params.poles = d->poles;
		p->params.flowers = d->flowers;
		set_grid(p->params, d->box_size, d->grid_resolution);
		p->boxSize = d->box_size;
		p->device = d->device;
		p->sampling = d->mesh_sampling;
		p->arithmetic = d->arithmetic;
		p->deviceCount = d->device_count > 1 ? d->device_count : 1;
		p->devices[0] = d->device;
		for (int i = 1; i < p->deviceCount; i++)
			p->devices[i] = d->devices[i];
		// UPG_MESH_SAMPLING=dense|band overrides the descriptor (A/B timing of the two ways to fill the grid)
		if (const char *env = std::getenv("UPG_MESH_SAMPLING"))
			p->sampling = std::strcmp(env, "dense") == 0 ? UPG_SAMPLING_DENSE : std::strcmp(env, "band") == 0 ? UPG_SAMPLING_BAND : p->sampling;

		// terrainPreseed order (sources/modesConfiguration.cpp:288-306): shape statics, then elevation statics, ...
		PlanetBuilder b;
		b.thread.s[0] = d->seed[0];
		b.thread.s[1] = d->seed[1];
		b.shapeStatics(p->params);
		b.elevationStatics(p->params);
		coloring_statics(b, p->params, p->coloring);
		p->seedLog = b.log;

		// the device side: fails loudly when there is no GPU (no CPU fallback exists)
		cudaError_t e = cudaSetDevice(p->device);
		if (e == cudaSuccess)
		{
			// the library's own stream-ordered pool of this device (upg_mem.h): created here so that a planet on an unusable device
			// fails at creation, and never the device's default pool (other allocators of the process keep theirs as it is)
			cudaMemPool_t pool = nullptr;
			e = device_pool(p->device, &pool);
		}
		if (e == cudaSuccess)
			e = cudaMalloc(&p->status, sizeof(uint32_t) * (STATUS_SLOTS + 1));
		if (e == cudaSuccess)
			e = cudaMemset(p->status, 0, sizeof(uint32_t) * (STATUS_SLOTS + 1));
		if (e == cudaSuccess)
			e = cudaMalloc(&p->cellularTable, sizeof(float) * 1024);
		if (e == cudaSuccess)
		{
			float table[1024];
			default_cellular_table(table);
			e = cudaMemcpy(p->cellularTable, table, sizeof(table), cudaMemcpyHostToDevice);
		}
		if (e != cudaSuccess)
		{
			if (p->status)
				cudaFree(p->status);
			if (p->cellularTable)
				cudaFree(p->cellularTable);
			delete p;
			return cuda_fail(e, "upg_planet_create");
		}
		p->children[0] = p;
		if (p->deviceCount > 1)
		{
			const int rc = multi_setup(p, d);
			if (rc != UPG_OK)
			{
				upg_planet_destroy(reinterpret_cast<upg_planet *>(p));
				return rc;
			}
		}
		*out = reinterpret_cast<upg_planet *>(p);
		return UPG_OK;
	}

	void upg_planet_destroy(upg_planet *planet)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p)
			return;
		multi_teardown(p);
		cudaSetDevice(p->device);
		if (p->status)
			cudaFree(p->status);
		if (p->cellularTable)
			cudaFree(p->cellularTable);
		delete p;
	}

	int upg_planet_set_cellular_table(upg_planet *planet, const float *table1024)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || !table1024)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		cudaError_t e = cudaSuccess;
		for (int i = 0; i < p->deviceCount && e == cudaSuccess; i++)
		{
			Planet *c = p->children[i];
			e = cudaSetDevice(c->device);
			if (e == cudaSuccess)
				e = cudaMemcpy(c->cellularTable, table1024, sizeof(float) * 1024, cudaMemcpyHostToDevice);
		}
		if (e != cudaSuccess)
			return cuda_fail(e, "upg_planet_set_cellular_table");
		return UPG_OK;
	}

	int upg_planet_async_status(upg_planet *planet, void *cuda_stream)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		uint32_t st = 0;
		cudaStream_t s = (cudaStream_t)cuda_stream;
		e = cudaMemcpyAsync(&st, p->status, sizeof(st), cudaMemcpyDeviceToHost, s);
		if (e == cudaSuccess)
			e = stream_sync(s);
		if (e != cudaSuccess)
			return cuda_fail(e, "upg_planet_async_status");
		if (st & 3u)
		{
			cudaMemsetAsync(p->status, 0, sizeof(uint32_t), s);
			stream_sync(s);
			if (st & 1u)
				return fail(UPG_ERR_INVALID_VALUE, "invalid sdf value (NaN) in an asynchronous sampling call");
			return fail(UPG_ERR_INVALID_VALUE, "upg_fields_samples_band_device: an elevation left the interval band sampling assumes; use the dense entry point for this planet");
		}
		return UPG_OK;
	}

	int upg_planet_seed_log(const upg_planet *planet, uint32_t *out, int capacity)
	{
		const Planet *p = reinterpret_cast<const Planet *>(planet);
		if (!p)
			return 0;
		for (int i = 0; i < p->seedLog.count && i < capacity && out; i++)
			out[i] = p->seedLog.seeds[i];
		return p->seedLog.count;
	}

	int upg_field_samples_device(upg_planet *planet, int field, int z0, int z1, float *device_out, void *cuda_stream)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || !device_out)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		const int res = p->params.grid.res;
		if (!valid_field(field) || z0 < 0 || z1 > res || z0 > z1)
			return fail(UPG_ERR_ARGUMENT, "bad field %d or slice range [%d, %d)", field, z0, z1);
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		FieldLaunch L;
		L.out = device_out;
		L.positions = nullptr;
		L.count = (uint64_t)(z1 - z0) * res * res;
		L.z0 = (uint32_t)z0;
		L.field = field;
		L.status = status_word(p);
		L.contract = true;
		return launch_field(p, L, (cudaStream_t)cuda_stream);
	}

	int upg_fields_samples_device(upg_planet *planet, int z0, int z1, float *device_land, float *device_water, float *device_navigation, void *cuda_stream)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || (!device_land && !device_water && !device_navigation))
			return fail(UPG_ERR_ARGUMENT, "null argument");
		const int res = p->params.grid.res;
		if (z0 < 0 || z1 > res || z0 > z1)
			return fail(UPG_ERR_ARGUMENT, "bad slice range [%d, %d)", z0, z1);
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		FieldLaunch L;
		L.out = device_land;
		L.outWater = device_water;
		L.outNavigation = device_navigation;
		L.positions = nullptr;
		L.count = (uint64_t)(z1 - z0) * res * res;
		L.z0 = (uint32_t)z0;
		L.field = FD_ALL;
		L.status = status_word(p);
		L.contract = true;
		return launch_field(p, L, (cudaStream_t)cuda_stream);
	}

	// the samples an extraction needs of slices [za, zb): dense (one fused evaluation per node) or, with `band`, the water field
	// densely followed by K1b (upg_band.cu), which needs `water` even when the caller does not ask for that field
	static int sample_for_extraction(Planet *p, bool band, int za, int zb, float *land, float *water, float *navigation, unsigned long long *evaluatedDevice, cudaStream_t stream, const SignBits *bits3 = nullptr, cudaEvent_t waterDone = nullptr)
	{
		const int res = p->params.grid.res;
		FieldLaunch L;
		L.positions = nullptr;
		L.count = (uint64_t)(zb - za) * res * res;
		L.z0 = (uint32_t)za;
		L.status = status_word(p);
		if (!band || (!land && !navigation))
		{
			L.out = land;
			L.outWater = water;
			L.outNavigation = navigation;
			L.field = FD_ALL;
			return launch_field(p, L, stream);
		}
		L.out = water;
		L.field = FD_WATER;
		if (bits3 && bits3[FD_WATER].signs)
		{
			L.signBits = bits3[FD_WATER].signs;
			L.nearBits = bits3[FD_WATER].near;
			L.bitPitch = bits3[FD_WATER].pitch;
			L.nearTol = NEAR_ZERO_TOLERANCE;
		}
		// the water launch also leaves the minimum / maximum of every row: the band sampler skips the rows whose neighbourhood is settled
		float2 *rowMinMax = nullptr;
		if (bits3 && device_alloc(&rowMinMax, sizeof(float2) * (uint64_t)(zb - za) * res, stream) != cudaSuccess)
		{
			cudaGetLastError();
			rowMinMax = nullptr; // optional
		}
		// all-ones = NaN = "not settled": a row the water launch should ever leave unwritten is scanned node by node
		if (rowMinMax && cudaMemsetAsync(rowMinMax, 0xff, sizeof(float2) * (uint64_t)(zb - za) * res, stream) != cudaSuccess)
		{
			cudaGetLastError();
			cudaFreeAsync(rowMinMax, stream);
			rowMinMax = nullptr;
		}
		L.rowMinMax = reinterpret_cast<float *>(rowMinMax);
		int rc = launch_field(p, L, stream);
		// the water field (and its bit fields) is complete here: an extraction of it may start beside the band kernel, which only reads it
		if (rc == UPG_OK && waterDone && cudaEventRecord(waterDone, stream) != cudaSuccess)
			rc = cuda_fail(cudaGetLastError(), "cudaEventRecord(water)");
		// with bit fields the extraction reads the signs from them and values only where the elevation was evaluated: no fill stores
		if (rc == UPG_OK)
			rc = band_sample(p, za, zb, water, land, navigation, evaluatedDevice, stream, bits3 ? &bits3[FD_LAND] : nullptr, bits3 ? &bits3[FD_NAVIGATION] : nullptr, bits3 == nullptr, rowMinMax);
		if (rowMinMax)
			cudaFreeAsync(rowMinMax, stream);
		return rc;
	}

	// zero-filled sign / near-zero bit fields (SignBits) of the fields in `mask` for `slices` z-slices: one stream-ordered allocation
	static int alloc_sign_bits(int res, int slices, unsigned mask, cudaStream_t stream, SignBits bits3[3], uint32_t **block)
	{
		*block = nullptr;
		const uint32_t pitch = sign_bits_pitch(res);
		const uint64_t words = (uint64_t)slices * res * pitch;
		int fields = 0;
		for (int f = 0; f < 3; f++)
			fields += (mask >> f) & 1u;
		if (!fields || !words)
			return UPG_OK;
		cudaError_t e = device_alloc(block, sizeof(uint32_t) * words * 2 * fields, stream);
		if (e == cudaSuccess)
			e = cudaMemsetAsync(*block, 0, sizeof(uint32_t) * words * 2 * fields, stream);
		if (e != cudaSuccess)
			return cuda_fail(e, "device_alloc(sign bits)");
		uint32_t *at = *block;
		for (int f = 0; f < 3; f++)
			if (mask & (1u << f))
			{
				bits3[f].signs = at;
				bits3[f].near = at + words;
				bits3[f].pitch = pitch;
				at += words * 2;
			}
		return UPG_OK;
	}

	// one field of slices [za, zb) for an extraction: land / navigation go through the band sampler when the planet asks for it
	static int sample_one_for_extraction(Planet *p, int field, int za, int zb, float *dens, cudaStream_t stream, SignBits *bits, uint32_t **bitBlock)
	{
		const int res = p->params.grid.res;
		const uint64_t count = (uint64_t)(zb - za) * res * res;
		*bitBlock = nullptr;
		if ((field == FD_LAND || field == FD_NAVIGATION) && band_sampling_applies(p))
		{
			SignBits bits3[3];
			int rc = alloc_sign_bits(res, zb - za, 1u << field, stream, bits3, bitBlock);
			if (rc != UPG_OK)
				return rc;
			float *shape = nullptr;
			const cudaError_t e = device_alloc(&shape, sizeof(float) * count, stream);
			if (e != cudaSuccess)
				return cuda_fail(e, "device_alloc(shape field)");
			rc = sample_for_extraction(p, true, za, zb, field == FD_LAND ? dens : nullptr, shape, field == FD_NAVIGATION ? dens : nullptr, nullptr, stream, bits3);
			cudaFreeAsync(shape, stream);
			*bits = bits3[field];
			return rc;
		}
		FieldLaunch L;
		L.out = dens;
		L.positions = nullptr;
		L.count = count;
		L.z0 = (uint32_t)za;
		L.field = field;
		L.status = status_word(p);
		return launch_field(p, L, stream);
	}

	int upg_fields_samples_band_device(upg_planet *planet, int z0, int z1, float *device_land, float *device_water, float *device_navigation, void *cuda_stream, uint64_t *evaluated_nodes)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || (!device_land && !device_water && !device_navigation))
			return fail(UPG_ERR_ARGUMENT, "null argument");
		const int res = p->params.grid.res;
		if (z0 < 0 || z1 > res || z0 > z1)
			return fail(UPG_ERR_ARGUMENT, "bad slice range [%d, %d)", z0, z1);
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		cudaStream_t stream = (cudaStream_t)cuda_stream;
		const uint64_t count = (uint64_t)(z1 - z0) * res * res;
		const bool band = p->params.elevation != EL_NONE && (device_land || device_navigation);
		float *scratch = nullptr;
		unsigned long long *counter = nullptr;
		if (band && !device_water && count)
		{
			e = device_alloc(&scratch, sizeof(float) * count, stream);
			if (e != cudaSuccess)
				return cuda_fail(e, "device_alloc(shape field)");
		}
		if (evaluated_nodes)
		{
			*evaluated_nodes = band ? 0 : count;
			e = device_alloc(&counter, sizeof(unsigned long long), stream);
			if (e == cudaSuccess)
				e = cudaMemsetAsync(counter, 0, sizeof(unsigned long long), stream);
			if (e != cudaSuccess)
			{
				if (scratch)
					cudaFreeAsync(scratch, stream);
				return cuda_fail(e, "device_alloc(counter)");
			}
		}
		int rc = count ? sample_for_extraction(p, band, z0, z1, device_land, device_water ? device_water : scratch, device_navigation, counter, stream) : UPG_OK;
		if (counter)
		{
			unsigned long long n = 0;
			e = cudaMemcpyAsync(&n, counter, sizeof(n), cudaMemcpyDeviceToHost, stream);
			if (e == cudaSuccess)
				e = stream_sync(stream);
			if (e != cudaSuccess && rc == UPG_OK)
				rc = cuda_fail(e, "upg_fields_samples_band_device");
			if (band)
				*evaluated_nodes = n;
			cudaFreeAsync(counter, stream);
		}
		if (scratch)
			cudaFreeAsync(scratch, stream);
		return rc;
	}

	int upg_field_samples(upg_planet *planet, int field, int z0, int z1, float *host_out)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || !host_out)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		const int res = p->params.grid.res;
		if (!valid_field(field) || z0 < 0 || z1 > res || z0 > z1)
			return fail(UPG_ERR_ARGUMENT, "bad field %d or slice range [%d, %d)", field, z0, z1);
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");

		// slab pipeline: sample slab k on the compute stream while slab k-1 drains to the host on the copy stream
		const uint64_t sliceElems = (uint64_t)res * res;
		const int slabSlices = (int)((((uint64_t)64 << 20) / sizeof(float) + sliceElems - 1) / sliceElems); // ~64 MiB slabs
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		StatusSlot slot; // this call's own NaN flag
		rc = slot.open(p, cs.compute);
		if (rc)
		{
			cs.close();
			return rc;
		}
		float *buf[2] = { nullptr, nullptr };
		cudaEvent_t sampled[2] = { nullptr, nullptr }, drained[2] = { nullptr, nullptr };
		for (int i = 0; i < 2 && e == cudaSuccess; i++)
		{
			e = device_alloc(&buf[i], sizeof(float) * sliceElems * slabSlices, cs.compute);
			if (e == cudaSuccess)
				e = cudaEventCreateWithFlags(&sampled[i], cudaEventDisableTiming);
			if (e == cudaSuccess)
				e = cudaEventCreateWithFlags(&drained[i], cudaEventDisableTiming);
		}
		int slab = 0;
		for (int z = z0; z < z1 && e == cudaSuccess && rc == UPG_OK; z += slabSlices, slab++)
		{
			const int b = slab & 1;
			const int ze = z + slabSlices < z1 ? z + slabSlices : z1;
			if (slab >= 2)
				e = cudaStreamWaitEvent(cs.compute, drained[b], 0);
			if (e != cudaSuccess)
				break;
			FieldLaunch L;
			L.out = buf[b];
			L.positions = nullptr;
			L.count = (uint64_t)(ze - z) * sliceElems;
			L.z0 = (uint32_t)z;
			L.field = field;
			L.status = status_word(p);
			L.contract = true;
			rc = launch_field(p, L, cs.compute);
			if (rc)
				break;
			e = cudaEventRecord(sampled[b], cs.compute);
			if (e == cudaSuccess)
				e = cudaStreamWaitEvent(cs.copy, sampled[b], 0);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(host_out + (uint64_t)(z - z0) * sliceElems, buf[b], sizeof(float) * L.count, cudaMemcpyDeviceToHost, cs.copy);
			if (e == cudaSuccess)
				e = cudaEventRecord(drained[b], cs.copy);
		}
		if (e == cudaSuccess)
			e = stream_sync(cs.copy);
		if (e == cudaSuccess)
			e = stream_sync(cs.compute);
		for (int i = 0; i < 2; i++)
		{
			if (buf[i])
				cudaFreeAsync(buf[i], cs.compute);
			if (sampled[i])
				cudaEventDestroy(sampled[i]);
			if (drained[i])
				cudaEventDestroy(drained[i]);
		}
		if (rc == UPG_OK && e != cudaSuccess)
			rc = cuda_fail(e, "upg_field_samples");
		if (rc == UPG_OK)
			rc = check_status(p, cs.compute, "upg_field_samples");
		cs.close();
		return rc;
	}

	int upg_field_points(upg_planet *planet, int field, const float *host_positions, uint64_t n, float *host_out)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || (n && (!host_positions || !host_out)))
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (!valid_field(field))
			return fail(UPG_ERR_ARGUMENT, "bad field %d", field);
		if (n == 0)
			return UPG_OK;
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		StatusSlot slot; // this call's own NaN flag
		rc = slot.open(p, cs.compute);
		if (rc)
		{
			cs.close();
			return rc;
		}
		float *dpos = nullptr, *dout = nullptr;
		e = device_alloc(&dpos, sizeof(float) * 3 * n, cs.compute);
		if (e == cudaSuccess)
			e = device_alloc(&dout, sizeof(float) * n, cs.compute);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(dpos, host_positions, sizeof(float) * 3 * n, cudaMemcpyHostToDevice, cs.compute);
		if (e == cudaSuccess)
		{
			FieldLaunch L;
			L.out = dout;
			L.positions = dpos;
			L.count = n;
			L.z0 = 0;
			L.field = field;
			L.status = status_word(p);
			rc = launch_field(p, L, cs.compute);
		}
		if (e == cudaSuccess && rc == UPG_OK)
			e = cudaMemcpyAsync(host_out, dout, sizeof(float) * n, cudaMemcpyDeviceToHost, cs.compute);
		if (e == cudaSuccess)
			e = stream_sync(cs.compute);
		if (dpos)
			cudaFreeAsync(dpos, cs.compute);
		if (dout)
			cudaFreeAsync(dout, cs.compute);
		if (rc == UPG_OK && e != cudaSuccess)
			rc = cuda_fail(e, "upg_field_points");
		if (rc == UPG_OK)
			rc = check_status(p, cs.compute, "upg_field_points");
		cs.close();
		return rc;
	}
}

// ---- isosurface entry points ------------------------------------------------------------------------------------
namespace upg
{
	// Pinned host blocks are expensive to create (cudaMallocHost of ~100 MB costs tens of milliseconds, more than the sampling
	// of a 500^3 field), and the generator asks for one mesh after another: blocks handed back by upg_mesh_free are kept
	// (up to 2 GiB in total) and reused for the next mesh of a similar size.
	namespace
	{
		struct PinnedBlock
		{
			void *ptr;
			size_t bytes;
		};
		std::mutex g_pinnedMutex;
		PinnedBlock g_pinnedFree[32];
		int g_pinnedCount = 0;
		size_t g_pinnedBytes = 0;

		// Requests are rounded up to 2^k or 1.5 * 2^k (at least 1 MiB) so that the meshes of successive planets, which never have the
		// same size twice, land on the same few block sizes: a batch of 64 different planets paid up to 960 ms per planet in
		// cudaMallocHost / cudaFreeHost before this rounding (bench.py batch leg), against ~10 ms of GPU work.
		size_t pinned_bucket(size_t bytes)
		{
			size_t b = (size_t)1 << 20;
			while (b < bytes)
			{
				if (b + b / 2 >= bytes)
					return b + b / 2;
				b <<= 1;
			}
			return b;
		}

		cudaError_t pinned_acquire(size_t bytes, void **out, size_t *capacity)
		{
			const size_t bucket = pinned_bucket(bytes);
			{
				std::lock_guard<std::mutex> lock(g_pinnedMutex);
				int best = -1;
				for (int i = 0; i < g_pinnedCount; i++)
					if (g_pinnedFree[i].bytes >= bytes && g_pinnedFree[i].bytes <= bucket * 8 && (best < 0 || g_pinnedFree[i].bytes < g_pinnedFree[best].bytes))
						best = i;
				if (best >= 0)
				{
					*out = g_pinnedFree[best].ptr;
					*capacity = g_pinnedFree[best].bytes;
					g_pinnedBytes -= g_pinnedFree[best].bytes;
					g_pinnedFree[best] = g_pinnedFree[--g_pinnedCount];
					return cudaSuccess;
				}
			}
			*capacity = bucket;
			return cudaMallocHost(out, bucket);
		}

		void pinned_release(void *ptr, size_t bytes)
		{
			{
				std::lock_guard<std::mutex> lock(g_pinnedMutex);
				if (g_pinnedCount < 32 && g_pinnedBytes + bytes <= ((size_t)2 << 30))
				{
					g_pinnedFree[g_pinnedCount++] = PinnedBlock{ ptr, bytes };
					g_pinnedBytes += bytes;
					return;
				}
			}
			cudaFreeHost(ptr);
		}

	}

	cudaError_t pinned_block_acquire(size_t bytes, void **out, size_t *capacity) { return pinned_acquire(bytes, out, capacity); }
	static cudaError_t pinned_reserve(size_t bytes, int count)
	{
		const size_t bucket = pinned_bucket(bytes);
		for (int i = 0; i < count; i++)
		{
			void *ptr = nullptr;
			const cudaError_t e = cudaMallocHost(&ptr, bucket);
			if (e != cudaSuccess)
				return e;
			pinned_release(ptr, bucket);
		}
		return cudaSuccess;
	}
	static void pinned_trim()
	{
		std::lock_guard<std::mutex> lock(g_pinnedMutex);
		for (int i = 0; i < g_pinnedCount; i++)
			cudaFreeHost(g_pinnedFree[i].ptr);
		g_pinnedCount = 0;
		g_pinnedBytes = 0;
	}
	static size_t pinned_cached()
	{
		std::lock_guard<std::mutex> lock(g_pinnedMutex);
		return g_pinnedBytes;
	}
	void pinned_block_release(void *ptr, size_t capacity) { pinned_release(ptr, capacity); }

	// Pageable host memory -> device.  cudaMemcpyAsync from pageable memory is staged by the driver on the calling thread (~11 GB/s
	// measured on the bench box: 7.4 ms for the 81 MB mesh of a 4096^2 bake, as long as half of the colouring it feeds); here worker
	// threads copy 4 MiB slices into a cached pinned block and enqueue the DMA of each slice as soon as it is staged.  Memory that is
	// already pinned / registered goes straight to the copy engine.  The block is returned through `staging` and must be released by
	// the caller (pinned_block_release) after the stream has passed the copies.
	int upload_host_arrays(const HostArray *arrays, int n, cudaStream_t stream, void **staging, size_t *stagingCapacity)
	{
		*staging = nullptr;
		*stagingCapacity = 0;
		size_t total = 0;
		bool pageable[16] = {};
		if (n > 16)
			return fail(UPG_ERR_ARGUMENT, "too many arrays");
		for (int i = 0; i < n; i++)
		{
			cudaPointerAttributes attr;
			const cudaError_t q = cudaPointerGetAttributes(&attr, arrays[i].src);
			pageable[i] = q != cudaSuccess || attr.type == cudaMemoryTypeUnregistered;
			if (q != cudaSuccess)
				cudaGetLastError();
			if (pageable[i])
				total += (arrays[i].bytes + 255) & ~(size_t)255;
		}
		cudaError_t e = cudaSuccess;
		for (int i = 0; i < n && e == cudaSuccess; i++)
			if (!pageable[i] && arrays[i].bytes)
				e = cudaMemcpyAsync(arrays[i].dst, arrays[i].src, arrays[i].bytes, cudaMemcpyHostToDevice, stream);
		if (e != cudaSuccess)
			return cuda_fail(e, "upload");
		if (total == 0)
			return UPG_OK;
		if (total < ((size_t)8 << 20))
		{
			// small: the driver's own staging is as good
			for (int i = 0; i < n && e == cudaSuccess; i++)
				if (pageable[i] && arrays[i].bytes)
					e = cudaMemcpyAsync(arrays[i].dst, arrays[i].src, arrays[i].bytes, cudaMemcpyHostToDevice, stream);
			return e == cudaSuccess ? UPG_OK : cuda_fail(e, "upload");
		}
		char *block = nullptr;
		e = pinned_acquire(total, (void **)&block, stagingCapacity);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaMallocHost(staging)");
		*staging = block;
		// slices in order of the arrays
		struct Slice
		{
			const char *src;
			char *dst;
			char *stage;
			size_t bytes;
		};
		std::vector<Slice> slices;
		const size_t sliceBytes = (size_t)4 << 20;
		size_t at = 0;
		for (int i = 0; i < n; i++)
			if (pageable[i])
			{
				for (size_t o = 0; o < arrays[i].bytes; o += sliceBytes)
					slices.push_back(Slice{ (const char *)arrays[i].src + o, (char *)arrays[i].dst + o, block + at + o, arrays[i].bytes - o < sliceBytes ? arrays[i].bytes - o : sliceBytes });
				at += (arrays[i].bytes + 255) & ~(size_t)255;
			}
		std::atomic<size_t> next{ 0 };
		std::atomic<int> failed{ 0 };
		int device = 0;
		cudaGetDevice(&device);
		auto work = [&]()
		{
			cudaSetDevice(device);
			for (size_t k = next.fetch_add(1); k < slices.size(); k = next.fetch_add(1))
			{
				std::memcpy(slices[k].stage, slices[k].src, slices[k].bytes);
				if (cudaMemcpyAsync(slices[k].dst, slices[k].stage, slices[k].bytes, cudaMemcpyHostToDevice, stream) != cudaSuccess)
					failed.store(1);
			}
		};
		// up to six copy threads, fewer when the host cores are shared by one process per GPU (bench.py --gpus 8: 8 ranks x 2 bakes)
		unsigned workers = std::thread::hardware_concurrency();
		int devices = 1;
		if (cudaGetDeviceCount(&devices) != cudaSuccess || devices < 1)
			devices = 1;
		workers /= (unsigned)devices;
		workers = workers > 6 ? 6 : workers < 1 ? 1 : workers;
		if (workers > slices.size())
			workers = (unsigned)slices.size();
		// Staged uploads take turns (per process: the copy threads and the host -> device link are shared by every call).  Two bakes
		// arriving together -- the generator's chunk tasks -- otherwise split the link, BOTH meshes arrive late, the colourings run back
		// to back on the GPU after that and both sets of images come back at the end.  One after the other, the first call's colouring
		// starts after half the time, the second upload hides behind it, and the first download hides behind the second colouring.
		// UPG_UPLOAD_TURNS=0 switches it off (A/B in profiles/).
		static std::mutex uploadTurn[64]; // one per device: uploads to different GPUs use different links and run side by side
		static const bool turns = !(std::getenv("UPG_UPLOAD_TURNS") && std::getenv("UPG_UPLOAD_TURNS")[0] == '0');
		std::unique_lock<std::mutex> turn(uploadTurn[device & 63], std::defer_lock);
		if (turns)
			turn.lock();
		std::vector<std::thread> pool;
		for (unsigned t = 1; t < workers; t++)
			pool.emplace_back(work);
		work();
		for (auto &t : pool)
			t.join();
		if (turns)
			turn.unlock();
		if (failed.load())
			return cuda_fail(cudaGetLastError(), "upload (staged)");
		return UPG_OK;
	}

	namespace
	{
		struct SleepingEvent
		{
			cudaEvent_t event = nullptr;
			int device = -1;
			~SleepingEvent()
			{
				if (event)
					cudaEventDestroy(event);
			}
		};
		bool blocking_sync_wanted()
		{
			if (const char *env = std::getenv("UPG_BLOCKING_SYNC"))
				return env[0] != '0';
			int devices = 1;
			if (cudaGetDeviceCount(&devices) != cudaSuccess)
			{
				cudaGetLastError();
				return false;
			}
			return devices >= 4;
		}
	}

	cudaError_t stream_sync(cudaStream_t s)
	{
		static const bool blocking = blocking_sync_wanted();
		if (!blocking)
			return cudaStreamSynchronize(s);
		thread_local SleepingEvent ev;
		int device = 0;
		cudaError_t e = cudaGetDevice(&device);
		if (e != cudaSuccess)
			return e;
		if (!ev.event || ev.device != device)
		{
			if (ev.event)
				cudaEventDestroy(ev.event);
			ev.event = nullptr;
			e = cudaEventCreateWithFlags(&ev.event, cudaEventBlockingSync | cudaEventDisableTiming);
			if (e != cudaSuccess)
			{
				// sleeping is an optimisation, never a reason to fail a call: without the event (the driver reports every resource it
				// cannot create for one as cudaErrorMemoryAllocation) wait the plain way
				ev.event = nullptr;
				cudaGetLastError();
				return cudaStreamSynchronize(s);
			}
			ev.device = device;
		}
		e = cudaEventRecord(ev.event, s);
		if (e != cudaSuccess)
		{
			cudaGetLastError();
			return cudaStreamSynchronize(s);
		}
		return cudaEventSynchronize(ev.event);
	}

	// UPG_TRACE=1: wall-clock phases of the host-buffer entry points on stderr (profiling aid)
	bool trace_enabled()
	{
		static const bool on = std::getenv("UPG_TRACE") != nullptr;
		return on;
	}
	double now_ms()
	{
		timespec ts;
		clock_gettime(CLOCK_MONOTONIC, &ts);
		return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
	}

	// Repeats `call` with dense sampling when band sampling's run-time guard fired (UPG_ERR_BAND_INTERVAL, upg_band.cu): the premise of
	// the shortcut failed for this planet, the reference's way of filling the grid does not depend on it.
	template<class F>
	static int with_dense_retry(F call)
	{
		int rc = call();
		if (rc == UPG_ERR_BAND_INTERVAL)
		{
			if (trace_enabled())
				std::fprintf(stderr, "[upg] band sampling: an elevation left its interval, repeating the call with dense sampling\n");
			set_force_dense_sampling(true);
			rc = call();
			set_force_dense_sampling(false);
		}
		return rc;
	}

	// copies a device-side DmcResult into one pinned host block and fills the ABI struct
	int mesh_to_host(DmcResult &r, cudaStream_t s, upg_mesh *out)
	{
		// a slab result only hands out the blocks it owns (DmcResult::first* / owned*); a whole volume owns everything
		const uint64_t nv = r.ownedVertices, nq = r.quads ? r.ownedQuads : 0, ni = (uint64_t)r.ownedTriangles * 3;
		const uint64_t v0 = (uint64_t)r.firstVertex * 3, q0 = (uint64_t)r.firstQuad * 4, i0 = (uint64_t)r.firstTriangle * 3;
		const uint64_t bytes = sizeof(float) * nv * 9 + sizeof(uint32_t) * (ni + nq * 4);
		char *block = nullptr;
		size_t capacity = 0;
		cudaError_t e = pinned_acquire(bytes ? bytes : 16, (void **)&block, &capacity);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaMallocHost(mesh)");
		char *p = block;
		out->positions = (float *)p;
		p += sizeof(float) * nv * 3;
		out->normals = (float *)p;
		p += sizeof(float) * nv * 3;
		out->dmc_vertices = (float *)p;
		p += sizeof(float) * nv * 3;
		out->indices = (uint32_t *)p;
		p += sizeof(uint32_t) * ni;
		out->dmc_quads = nq ? (uint32_t *)p : nullptr;
		if (nv)
		{
			e = cudaMemcpyAsync(out->positions, r.positions + v0, sizeof(float) * nv * 3, cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess)
				e = cudaMemcpyAsync(out->normals, r.normals + v0, sizeof(float) * nv * 3, cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess && r.rawVertices)
				e = cudaMemcpyAsync(out->dmc_vertices, r.rawVertices + v0, sizeof(float) * nv * 3, cudaMemcpyDeviceToHost, s);
		}
		if (!r.rawVertices)
			out->dmc_vertices = nullptr; // a cleaned mesh (K9) has no dualmc counterpart
		if (e == cudaSuccess && ni)
			e = cudaMemcpyAsync(out->indices, r.indices + i0, sizeof(uint32_t) * ni, cudaMemcpyDeviceToHost, s);
		if (e == cudaSuccess && nq)
			e = cudaMemcpyAsync(out->dmc_quads, r.quads + q0, sizeof(uint32_t) * nq * 4, cudaMemcpyDeviceToHost, s);
		if (e == cudaSuccess)
			e = stream_sync(s);
		if (e != cudaSuccess)
		{
			pinned_release(block, capacity);
			std::memset(out, 0, sizeof(*out));
			return cuda_fail(e, "mesh download");
		}
		out->vertex_count = (uint32_t)nv;
		out->index_count = (uint32_t)ni;
		out->quad_count = (uint32_t)nq;
		out->reserved = 0;
		out->near_zero_samples = r.nearZeroSamples;
		out->owner = block;
		out->owner_bytes = capacity;
		return UPG_OK;
	}
}

extern "C"
{
	static int field_mesh_once(upg_planet *planet, int field, upg_mesh *out)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || !out)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		std::memset(out, 0, sizeof(*out));
		if (!valid_field(field))
			return fail(UPG_ERR_ARGUMENT, "bad field %d", field);
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		const int res = p->params.grid.res;
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		StatusSlot slot; // this call's own NaN flag
		rc = slot.open(p, cs.compute);
		if (rc)
		{
			cs.close();
			return rc;
		}
		const bool trace = trace_enabled();
		const double t0 = trace ? now_ms() : 0;
		double t1 = 0, t2 = 0;
		float *dens = nullptr;
		e = device_alloc(&dens, sizeof(float) * (uint64_t)res * res * res, cs.compute);
		if (e != cudaSuccess)
			rc = cuda_fail(e, "device_alloc(field)");
		SignBits bits;
		uint32_t *bitBlock = nullptr;
		if (rc == UPG_OK)
			rc = sample_one_for_extraction(p, field, 0, res, dens, cs.compute, &bits, &bitBlock);
		if (trace)
		{
			stream_sync(cs.compute);
			t1 = now_ms();
		}
		DmcResult r;
		if (rc == UPG_OK)
			rc = dmc_extract(dens, res, res, res, p->boxSize, res, cs.compute, r, 0, 0, 0x7fffffff, bits.signs ? &bits : nullptr);
		if (bitBlock)
			cudaFreeAsync(bitBlock, cs.compute);
		if (rc == UPG_OK)
			rc = check_status(p, cs.compute, "upg_field_mesh");
		if (rc == UPG_OK && r.triangleCount == 0)
			rc = fail(UPG_ERR_EMPTY_MESH, "generated empty mesh");
		if (trace)
		{
			stream_sync(cs.compute);
			t2 = now_ms();
		}
		if (rc == UPG_OK)
			rc = mesh_to_host(r, cs.compute, out);
		if (trace)
			std::fprintf(stderr, "[upg] field_mesh(%d) res %d: sample %.2f ms, extract %.2f ms, download %.2f ms (%u vertices, %u triangles)\n", field, res, t1 - t0, t2 - t1, now_ms() - t2, r.vertexCount, r.triangleCount);
		dmc_release(r, cs.compute);
		if (dens)
			cudaFreeAsync(dens, cs.compute);
		stream_sync(cs.compute);
		cs.close();
		return rc;
	}

	// hands a device-side result over to the caller as it is (arrays stay in HBM, owned until upg_device_mesh_free)
	static void mesh_to_device_struct(DmcResult &r, int device, upg_device_mesh *out)
	{
		const uint64_t v0 = (uint64_t)r.firstVertex * 3, q0 = (uint64_t)r.firstQuad * 4, i0 = (uint64_t)r.firstTriangle * 3;
		out->positions = r.positions ? r.positions + v0 : nullptr;
		out->normals = r.normals ? r.normals + v0 : nullptr;
		out->dmc_vertices = r.rawVertices ? r.rawVertices + v0 : nullptr;
		out->indices = r.indices ? r.indices + i0 : nullptr;
		out->dmc_quads = r.quads ? r.quads + q0 : nullptr;
		out->vertex_count = r.ownedVertices;
		out->index_count = r.ownedTriangles * 3;
		out->quad_count = r.quads ? r.ownedQuads : 0;
		out->device = device;
		out->near_zero_samples = r.nearZeroSamples;
		void **own = new void *[5]{ r.positions, r.normals, r.rawVertices, r.indices, r.quads };
		out->owner = own;
		r = DmcResult(); // ownership moved
	}

	// slabCount == 0: the whole grid.  Exactly one of hostOut3 / deviceOut3 is given.
	int upg_field_mesh(upg_planet *planet, int field, upg_mesh *out)
	{
		return with_dense_retry([&]() { return field_mesh_once(planet, field, out); });
	}

	static int field_meshes_once(upg_planet *planet, unsigned field_mask, unsigned stages, int slab, int slabCount, upg_mesh *out3, upg_device_mesh *deviceOut3);
	static int field_meshes_impl(upg_planet *planet, unsigned field_mask, unsigned stages, int slab, int slabCount, upg_mesh *out3, upg_device_mesh *deviceOut3 = nullptr)
	{
		return with_dense_retry([&]() { return field_meshes_once(planet, field_mask, stages, slab, slabCount, out3, deviceOut3); });
	}

	static int field_meshes_once(upg_planet *planet, unsigned field_mask, unsigned stages, int slab, int slabCount, upg_mesh *out3, upg_device_mesh *deviceOut3)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || (!out3 && !deviceOut3))
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (out3)
			std::memset(out3, 0, sizeof(upg_mesh) * 3);
		if (deviceOut3)
			std::memset(deviceOut3, 0, sizeof(upg_device_mesh) * 3);
		if (slabCount < 0 || (slabCount > 0 && (slab < 0 || slab >= slabCount)))
			return fail(UPG_ERR_ARGUMENT, "bad slab %d of %d", slab, slabCount);
		if (slabCount > 0 && stages)
			return fail(UPG_ERR_ARGUMENT, "the clean-up works on whole meshes, not on slabs");
		if (stages & ~(unsigned)UPG_MESH_GENERATOR)
			return fail(UPG_ERR_ARGUMENT, "bad clean-up stages %#x", stages);
		if ((field_mask & 7u) == 0 || (field_mask & ~7u))
			return fail(UPG_ERR_ARGUMENT, "bad field mask %#x (bit 0 land, bit 1 water, bit 2 navigation)", field_mask);
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		const int res = p->params.grid.res;
		// a slab samples its scan iterations plus three slices of halo on each side (see upg_field_mesh_slab)
		int z0 = 0, z1 = res, za = 0, zb = res;
		if (slabCount > 0)
		{
			upg_slab_range(res, slab, slabCount, &z0, &z1);
			if (z0 == z1)
				return UPG_OK; // more slabs than scan iterations: an empty share
			za = z0 - 3 > 0 ? z0 - 3 : 0;
			zb = z1 + 3 < res ? z1 + 3 : res;
		}
		const uint64_t count = (uint64_t)(zb - za) * res * res;
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		StatusSlot slot; // this call's own NaN flag
		rc = slot.open(p, cs.compute);
		if (rc)
		{
			cs.close();
			return rc;
		}
		const bool trace = trace_enabled();
		const double t0 = trace ? now_ms() : 0;
		float *dens[3] = { nullptr, nullptr, nullptr };
		for (int f = 0; f < 3 && rc == UPG_OK; f++)
			if (field_mask & (1u << f))
			{
				e = device_alloc(&dens[f], sizeof(float) * count, cs.compute);
				if (e != cudaSuccess)
					rc = cuda_fail(e, "device_alloc(field)");
			}
		// band sampling reads the shape from the water field: a scratch copy when the caller does not want that mesh
		const bool band = band_sampling_applies(p) && (dens[FD_LAND] || dens[FD_NAVIGATION]);
		float *shapeScratch = nullptr;
		if (rc == UPG_OK && band && !dens[FD_WATER])
		{
			e = device_alloc(&shapeScratch, sizeof(float) * count, cs.compute);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "device_alloc(shape field)");
		}
		// shape and elevation are evaluated ONCE per sample for all requested fields (and the elevation only where the isosurface
		// can see it, UPG_SAMPLING_BAND)
		SignBits bits3[3];
		uint32_t *bitBlock = nullptr;
		if (rc == UPG_OK && band)
			rc = alloc_sign_bits(res, zb - za, field_mask, cs.compute, bits3, &bitBlock);
		// The water extraction starts as soon as the water launch is done, beside the band kernel of land / navigation (which only READS
		// the water field), and the workers are released without a host synchronisation: their streams wait for the sampling on the
		// device, the NaN status of the sampling is checked after they are done.  Round 1 reverted this ordering because one run lost a
		// vertex; the cause was the union-find flatten of K9, in every ordering (profiles/r02_flake.md: 0 deviations in 1200 planets
		// since).  UPG_EARLY_WATER=0 restores the conservative ordering.
		static const bool earlyWater = !(std::getenv("UPG_EARLY_WATER") && std::getenv("UPG_EARLY_WATER")[0] == '0');
		// UPG_DEBUG_WORKERS=<bits>: 1 = one worker at a time, 2 = every worker waits (host side) for the sampling before its first call
		static const unsigned long debugWorkers = std::getenv("UPG_DEBUG_WORKERS") ? std::strtoul(std::getenv("UPG_DEBUG_WORKERS"), nullptr, 0) : 0;
		static std::mutex debugWorkerMutex;
		cudaEvent_t waterDone = nullptr;
		if (rc == UPG_OK && earlyWater && band && dens[FD_WATER])
		{
			e = cudaEventCreateWithFlags(&waterDone, cudaEventDisableTiming);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "cudaEventCreate");
		}
		if (rc == UPG_OK)
			rc = sample_for_extraction(p, band, za, zb, dens[FD_LAND], dens[FD_WATER] ? dens[FD_WATER] : shapeScratch, dens[FD_NAVIGATION], nullptr, cs.compute, band ? bits3 : nullptr, waterDone);
		if (shapeScratch)
			cudaFreeAsync(shapeScratch, cs.compute);
		cudaEvent_t sampled = nullptr;
		if (rc == UPG_OK)
		{
			e = cudaEventCreateWithFlags(&sampled, cudaEventDisableTiming);
			if (e == cudaSuccess)
				e = cudaEventRecord(sampled, cs.compute);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "cudaEventRecord");
		}
		if (rc == UPG_OK && !earlyWater)
			rc = check_status(p, cs.compute, "upg_field_meshes");
		const double t1 = trace ? now_ms() : 0;
		// the three extractions are independent: one host thread and one stream pair each, like the generator's three tasks
		// (sources/generator.cpp:432-439); their sparse passes and downloads overlap
		int rcs[3] = { UPG_OK, UPG_OK, UPG_OK };
		std::string errs[3];
		if (rc == UPG_OK)
		{
			std::thread workers[3];
			for (int f = 0; f < 3; f++)
			{
				if (!dens[f])
					continue;
				workers[f] = std::thread(
					[&, f]()
					{
						int r = UPG_OK;
						std::unique_lock<std::mutex> oneAtATime(debugWorkerMutex, std::defer_lock);
						if (debugWorkers & 1)
							oneAtATime.lock();
						StatusSlot workerSlot;
						workerSlot.adopt(slot.word()); // the workers belong to this call
						cudaError_t ee = cudaSetDevice(p->device);
						CallStreams ws;
						if (ee != cudaSuccess)
							r = cuda_fail(ee, "cudaSetDevice");
						if (r == UPG_OK)
							r = ws.open();
						if (r == UPG_OK)
						{
							ee = cudaStreamWaitEvent(ws.compute, (f == FD_WATER && waterDone) ? waterDone : sampled, 0);
							if (ee == cudaSuccess && (debugWorkers & 2))
								ee = stream_sync(ws.compute);
							if (ee != cudaSuccess)
								r = cuda_fail(ee, "cudaStreamWaitEvent");
						}
						DmcResult d;
						const double w0 = trace ? now_ms() : 0;
						double w1 = 0, w2 = 0;
						if (r == UPG_OK)
							r = dmc_extract(dens[f], res, res, zb - za, p->boxSize, res, ws.compute, d, za, slabCount > 0 ? z0 : 0, slabCount > 0 ? z1 : 0x7fffffff, bits3[f].signs ? &bits3[f] : nullptr);
						if (trace)
						{
							stream_sync(ws.compute);
							w1 = now_ms();
						}
						if (r == UPG_OK && stages)
						{
							// K9 on the arrays still in HBM; the water culling belongs to the water mesh alone (sources/meshGeneration.cpp:57-98)
							DmcResult c;
							r = mesh_clean(p, d.positions, d.normals, d.vertexCount, d.indices, d.triangleCount, p->boxSize, f == FD_WATER ? stages : (stages & ~(unsigned)UPG_MESH_WATER_CULL), ws.compute, c);
							c.nearZeroSamples = d.nearZeroSamples;
							dmc_release(d, ws.compute);
							d = c;
							if (r == UPG_OK)
								r = check_status(p, ws.compute, "upg_field_meshes_clean");
						}
						// the reference throws for an empty land / navigation mesh only (sources/meshGeneration.cpp:51-52, 104-105)
						if (r == UPG_OK && slabCount == 0 && d.triangleCount == 0 && !(stages && f == FD_WATER))
							r = fail(UPG_ERR_EMPTY_MESH, "generated empty base %s mesh", f == FD_LAND ? "land" : f == FD_WATER ? "water" : "navigation");
						if (trace)
						{
							stream_sync(ws.compute);
							w2 = now_ms();
						}
						if (r == UPG_OK && out3)
							r = mesh_to_host(d, ws.compute, &out3[f]);
						else if (r == UPG_OK)
						{
							ee = stream_sync(ws.compute);
							if (ee != cudaSuccess)
								r = cuda_fail(ee, "extraction");
							else
								mesh_to_device_struct(d, p->device, &deviceOut3[f]);
						}
						if (trace)
							std::fprintf(stderr, "[upg]   field %d: extract %.2f ms, clean-up %.2f ms, download %.2f ms (%u vertices, %u triangles)\n", f, w1 - w0, w2 - w1, now_ms() - w2, d.vertexCount, d.triangleCount);
						dmc_release(d, ws.compute);
						if (ws.compute)
							stream_sync(ws.compute);
						ws.close();
						rcs[f] = r;
						if (r != UPG_OK)
							errs[f] = upg_last_error();
					});
			}
			for (int f = 0; f < 3; f++)
				if (workers[f].joinable())
					workers[f].join();
			for (int f = 0; f < 3 && rc == UPG_OK; f++)
				if (rcs[f] != UPG_OK)
					rc = fail(rcs[f], "%s", errs[f].c_str());
			if (rc == UPG_OK && earlyWater)
				rc = check_status(p, cs.compute, "upg_field_meshes");
		}
		if (waterDone)
			cudaEventDestroy(waterDone);
		if (trace)
			std::fprintf(stderr, "[upg] field_meshes(mask %#x) res %d: %s sampling %.2f ms, extraction + download %.2f ms\n", field_mask, res, band ? "band" : "dense fused", t1 - t0, now_ms() - t1);
		if (sampled)
			cudaEventDestroy(sampled);
		for (int f = 0; f < 3; f++)
			if (dens[f])
				cudaFreeAsync(dens[f], cs.compute);
		if (bitBlock)
			cudaFreeAsync(bitBlock, cs.compute);
		stream_sync(cs.compute);
		cs.close();
		if (rc != UPG_OK)
			for (int f = 0; f < 3; f++)
			{
				if (out3)
					upg_mesh_free(&out3[f]);
				if (deviceOut3)
					upg_device_mesh_free(&deviceOut3[f]);
			}
		return rc;
	}

	int upg_field_meshes(upg_planet *planet, unsigned field_mask, upg_mesh *out3)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (p && p->deviceCount > 1)
			return multi_field_meshes(p, field_mask, 0, out3);
		return field_meshes_impl(planet, field_mask, 0, 0, 0, out3);
	}

	int upg_field_meshes_slab(upg_planet *planet, unsigned field_mask, int slab, int slab_count, upg_mesh *out3)
	{
		if (slab_count < 1)
			return fail(UPG_ERR_ARGUMENT, "bad slab %d of %d", slab, slab_count);
		return field_meshes_impl(planet, field_mask, 0, slab, slab_count, out3);
	}

	int upg_field_meshes_slab_device(upg_planet *planet, unsigned field_mask, int slab, int slab_count, upg_device_mesh *out3)
	{
		if (slab_count < 1)
			return fail(UPG_ERR_ARGUMENT, "bad slab %d of %d", slab, slab_count);
		return field_meshes_impl(planet, field_mask, 0, slab, slab_count, nullptr, out3);
	}

	void upg_device_mesh_free(upg_device_mesh *mesh)
	{
		if (!mesh)
			return;
		if (mesh->owner)
		{
			void **own = reinterpret_cast<void **>(mesh->owner);
			int previous = 0;
			cudaGetDevice(&previous);
			cudaSetDevice(mesh->device);
			for (int i = 0; i < 5; i++)
				if (own[i])
					cudaFreeAsync(own[i], nullptr);
			cudaSetDevice(previous);
			delete[] own;
		}
		std::memset(mesh, 0, sizeof(*mesh));
	}

	int upg_field_meshes_clean(upg_planet *planet, unsigned field_mask, unsigned stages, upg_mesh *out3)
	{
		if (stages == 0)
			return fail(UPG_ERR_ARGUMENT, "no clean-up stage requested");
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (p && p->deviceCount > 1)
			return multi_field_meshes(p, field_mask, stages, out3);
		return field_meshes_impl(planet, field_mask, stages, 0, 0, out3);
	}

	int upg_mesh_clean(int device, const float *positions, const float *normals, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, float box_size, unsigned stages, upg_mesh *out)
	{
		if (!out)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		std::memset(out, 0, sizeof(*out));
		if ((!positions || !normals) && vertex_count)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (!indices && index_count)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (index_count % 3 != 0)
			return fail(UPG_ERR_ARGUMENT, "the clean-up requires a triangle mesh");
		if (stages == 0 || (stages & ~(unsigned)(UPG_MESH_MAKEMESH_TAIL | UPG_MESH_REMOVE_DISCONNECTED | UPG_MESH_FLIP_NORMALS)))
			return fail(UPG_ERR_ARGUMENT, "bad clean-up stages %#x (the water culling needs a planet: upg_field_meshes_clean)", stages);
		for (uint32_t i = 0; i < index_count; i++)
			if (indices[i] >= vertex_count)
				return fail(UPG_ERR_ARGUMENT, "index %u out of range (%u vertices)", indices[i], vertex_count);
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		cudaStream_t s = cs.compute;
		float *dp = nullptr, *dn = nullptr;
		uint32_t *di = nullptr;
		e = device_alloc(&dp, sizeof(float) * 3 * (vertex_count ? vertex_count : 1), s);
		if (e == cudaSuccess)
			e = device_alloc(&dn, sizeof(float) * 3 * (vertex_count ? vertex_count : 1), s);
		if (e == cudaSuccess)
			e = device_alloc(&di, sizeof(uint32_t) * (index_count ? index_count : 1), s);
		if (e == cudaSuccess && vertex_count)
			e = cudaMemcpyAsync(dp, positions, sizeof(float) * 3 * vertex_count, cudaMemcpyHostToDevice, s);
		if (e == cudaSuccess && vertex_count)
			e = cudaMemcpyAsync(dn, normals, sizeof(float) * 3 * vertex_count, cudaMemcpyHostToDevice, s);
		if (e == cudaSuccess && index_count)
			e = cudaMemcpyAsync(di, indices, sizeof(uint32_t) * index_count, cudaMemcpyHostToDevice, s);
		if (e != cudaSuccess)
			rc = cuda_fail(e, "upload mesh");
		DmcResult r;
		if (rc == UPG_OK)
			rc = mesh_clean(nullptr, dp, dn, vertex_count, di, index_count / 3, box_size, stages, s, r);
		if (rc == UPG_OK)
			rc = mesh_to_host(r, s, out);
		dmc_release(r, s);
		if (dp)
			cudaFreeAsync(dp, s);
		if (dn)
			cudaFreeAsync(dn, s);
		if (di)
			cudaFreeAsync(di, s);
		stream_sync(s);
		cs.close();
		return rc;
	}

	int upg_mesh_chunking(int device, const float *positions, const float *normals, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, float max_surface_area, upg_mesh **chunks, uint32_t *chunk_count)
	{
		if (!chunks || !chunk_count)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		*chunks = nullptr;
		*chunk_count = 0;
		if ((!positions || !normals) && vertex_count)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (!indices && index_count)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		if (index_count % 3 != 0)
			return fail(UPG_ERR_ARGUMENT, "mesh chunking requires triangles mesh");
		if (!(max_surface_area > 0))
			return fail(UPG_ERR_ARGUMENT, "mesh chunking requires positive maxSurfaceArea");
		if (index_count == 0)
			return UPG_OK; // meshChunking: `if (msh->facesCount() == 0) return {};`
		for (uint32_t i = 0; i < index_count; i++)
			if (indices[i] >= vertex_count)
				return fail(UPG_ERR_ARGUMENT, "index %u out of range (%u vertices)", indices[i], vertex_count);
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		cudaStream_t s = cs.compute;
		float *dp = nullptr, *dn = nullptr;
		uint32_t *di = nullptr;
		e = device_alloc(&dp, sizeof(float) * 3 * (size_t)vertex_count, s);
		if (e == cudaSuccess)
			e = device_alloc(&dn, sizeof(float) * 3 * (size_t)vertex_count, s);
		if (e == cudaSuccess)
			e = device_alloc(&di, sizeof(uint32_t) * (size_t)index_count, s);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(dp, positions, sizeof(float) * 3 * (size_t)vertex_count, cudaMemcpyHostToDevice, s);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(dn, normals, sizeof(float) * 3 * (size_t)vertex_count, cudaMemcpyHostToDevice, s);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(di, indices, sizeof(uint32_t) * (size_t)index_count, cudaMemcpyHostToDevice, s);
		if (e != cudaSuccess)
			rc = cuda_fail(e, "upload mesh");
		std::vector<DmcResult> leaves;
		if (rc == UPG_OK)
			rc = mesh_chunking(dp, dn, vertex_count, di, index_count / 3, max_surface_area, s, leaves);
		upg_mesh *list = nullptr;
		if (rc == UPG_OK)
		{
			list = static_cast<upg_mesh *>(std::calloc(leaves.size() ? leaves.size() : 1, sizeof(upg_mesh)));
			if (!list)
				rc = fail(UPG_ERR_ARGUMENT, "out of host memory");
		}
		for (size_t i = 0; i < leaves.size() && rc == UPG_OK; i++)
			rc = mesh_to_host(leaves[i], s, &list[i]);
		for (auto &l : leaves)
			dmc_release(l, s);
		if (dp)
			cudaFreeAsync(dp, s);
		if (dn)
			cudaFreeAsync(dn, s);
		if (di)
			cudaFreeAsync(di, s);
		stream_sync(s);
		cs.close();
		if (rc != UPG_OK)
		{
			if (list)
				upg_mesh_chunks_free(list, (uint32_t)leaves.size());
			return rc;
		}
		*chunks = list;
		*chunk_count = (uint32_t)leaves.size();
		return UPG_OK;
	}

	void upg_mesh_chunks_free(upg_mesh *chunks, uint32_t chunk_count)
	{
		if (!chunks)
			return;
		for (uint32_t i = 0; i < chunk_count; i++)
			upg_mesh_free(&chunks[i]);
		std::free(chunks);
	}

	void upg_slab_range(int grid_resolution, int slab, int slab_count, int *z0, int *z1)
	{
		// dualmc's z loop runs over [0, res - 2) (dualmc.tpp:401); split it evenly, earlier slabs take the remainder
		const int total = grid_resolution > 2 ? grid_resolution - 2 : 0;
		const int base = slab_count > 0 ? total / slab_count : 0, rem = slab_count > 0 ? total % slab_count : 0;
		const int a = slab * base + (slab < rem ? slab : rem);
		if (z0)
			*z0 = a;
		if (z1)
			*z1 = a + base + (slab < rem ? 1 : 0);
	}

	static int field_mesh_slab_once(upg_planet *planet, int field, int slab, int slab_count, upg_mesh *out);
	int upg_field_mesh_slab(upg_planet *planet, int field, int slab, int slab_count, upg_mesh *out)
	{
		return with_dense_retry([&]() { return field_mesh_slab_once(planet, field, slab, slab_count, out); });
	}

	static int field_mesh_slab_once(upg_planet *planet, int field, int slab, int slab_count, upg_mesh *out)
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || !out)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		std::memset(out, 0, sizeof(*out));
		if (!valid_field(field) || slab_count < 1 || slab < 0 || slab >= slab_count)
			return fail(UPG_ERR_ARGUMENT, "bad field %d or slab %d of %d", field, slab, slab_count);
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		const int res = p->params.grid.res;
		int z0 = 0, z1 = 0;
		upg_slab_range(res, slab, slab_count, &z0, &z1);
		if (z0 == z1)
			return UPG_OK; // more slabs than scan iterations: an empty share
		// Samples the slab needs: scan iteration z touches cells z-1, z; a cell's manifold test looks one cell further and a
		// cell spans two slices.  Vertices of cell z0-1 may have been created by iteration z0-1 (lower neighbour's), and the
		// normals of the vertices created here also gather the faces of iteration z1: three slices of halo on each side,
		// RECOMPUTED here (0.6 % extra samples per slab at 1024^3 / 8) instead of exchanged.
		const int a = z0 - 3 > 0 ? z0 - 3 : 0;
		const int b = z1 + 3 < res ? z1 + 3 : res;
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		StatusSlot slot; // this call's own NaN flag
		rc = slot.open(p, cs.compute);
		if (rc)
		{
			cs.close();
			return rc;
		}
		float *dens = nullptr;
		const uint64_t count = (uint64_t)(b - a) * res * res;
		e = device_alloc(&dens, sizeof(float) * count, cs.compute);
		if (e != cudaSuccess)
			rc = cuda_fail(e, "device_alloc(slab)");
		SignBits bits;
		uint32_t *bitBlock = nullptr;
		if (rc == UPG_OK)
			rc = sample_one_for_extraction(p, field, a, b, dens, cs.compute, &bits, &bitBlock);
		DmcResult r;
		if (rc == UPG_OK)
			rc = dmc_extract(dens, res, res, b - a, p->boxSize, res, cs.compute, r, a, z0, z1, bits.signs ? &bits : nullptr);
		if (bitBlock)
			cudaFreeAsync(bitBlock, cs.compute);
		if (rc == UPG_OK)
			rc = check_status(p, cs.compute, "upg_field_mesh_slab");
		if (rc == UPG_OK)
			rc = mesh_to_host(r, cs.compute, out);
		dmc_release(r, cs.compute);
		if (dens)
			cudaFreeAsync(dens, cs.compute);
		stream_sync(cs.compute);
		cs.close();
		return rc;
	}

	int upg_isosurface(const float *host_samples, int rx, int ry, int rz, float box_size, int full_resolution, int device, upg_mesh *out)
	{
		if (!host_samples || !out)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		std::memset(out, 0, sizeof(*out));
		if (rx <= 2 || ry <= 2 || rz <= 2 || full_resolution <= 5)
			return fail(UPG_ERR_ARGUMENT, "bad volume size %d x %d x %d (resolution %d)", rx, ry, rz, full_resolution);
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		const uint64_t n = (uint64_t)rx * ry * rz;
		float *dens = nullptr;
		e = device_alloc(&dens, sizeof(float) * n, cs.compute);
		if (e == cudaSuccess)
			e = cudaMemcpyAsync(dens, host_samples, sizeof(float) * n, cudaMemcpyHostToDevice, cs.compute);
		if (e != cudaSuccess)
			rc = cuda_fail(e, "upload samples");
		DmcResult r;
		if (rc == UPG_OK)
			rc = dmc_extract(dens, rx, ry, rz, box_size, full_resolution, cs.compute, r);
		if (rc == UPG_OK)
			rc = mesh_to_host(r, cs.compute, out);
		dmc_release(r, cs.compute);
		if (dens)
			cudaFreeAsync(dens, cs.compute);
		stream_sync(cs.compute);
		cs.close();
		return rc;
	}

	int upg_upload(int device, void *const *device_dst, const void *const *host_src, const uint64_t *bytes, int count, void *cuda_stream)
	{
		if (!device_dst || !host_src || !bytes || count < 0 || count > 16)
			return fail(UPG_ERR_ARGUMENT, "bad argument");
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		HostArray arrays[16];
		for (int i = 0; i < count; i++)
			arrays[i] = HostArray{ device_dst[i], host_src[i], (size_t)bytes[i] };
		void *staging = nullptr;
		size_t capacity = 0;
		int rc = upload_host_arrays(arrays, count, (cudaStream_t)cuda_stream, &staging, &capacity);
		e = stream_sync((cudaStream_t)cuda_stream);
		if (staging)
			pinned_block_release(staging, capacity);
		if (rc == UPG_OK && e != cudaSuccess)
			rc = cuda_fail(e, "upg_upload");
		return rc;
	}

	int upg_reserve(int device, uint64_t device_bytes, uint64_t pinned_block_bytes, int pinned_blocks)
	{
		if (pinned_blocks < 0 || pinned_blocks > 32)
			return fail(UPG_ERR_ARGUMENT, "pinned block count %d outside [0, 32]", pinned_blocks);
		cudaError_t e = cudaSetDevice(device);
		if (e == cudaSuccess && device_bytes)
			e = device_pool_reserve(device, device_bytes);
		if (e == cudaSuccess && pinned_blocks && pinned_block_bytes)
			e = pinned_reserve((size_t)pinned_block_bytes, pinned_blocks);
		if (e != cudaSuccess)
			return cuda_fail(e, "upg_reserve");
		return UPG_OK;
	}

	int upg_trim(int device, uint64_t keep_bytes)
	{
		cudaError_t e = cudaSetDevice(device);
		if (e == cudaSuccess)
			e = cudaDeviceSynchronize(); // stream-ordered frees of finished calls become visible to the trim
		if (e == cudaSuccess)
			e = device_pool_trim(device, keep_bytes);
		if (e != cudaSuccess)
			return cuda_fail(e, "upg_trim");
		pinned_trim();
		return UPG_OK;
	}

	int upg_memory_stats(int device, uint64_t *device_reserved, uint64_t *device_used_high, uint64_t *pinned_cached_bytes)
	{
		uint64_t r = 0, u = 0;
		const cudaError_t e = device_pool_stats(device, &r, &u);
		if (e != cudaSuccess)
			return cuda_fail(e, "upg_memory_stats");
		if (device_reserved)
			*device_reserved = r;
		if (device_used_high)
			*device_used_high = u;
		if (pinned_cached_bytes)
			*pinned_cached_bytes = pinned_cached();
		return UPG_OK;
	}

	void upg_mesh_free(upg_mesh *mesh)
	{
		if (!mesh)
			return;
		if (mesh->owner)
			pinned_release(mesh->owner, (size_t)mesh->owner_bytes);
		std::memset(mesh, 0, sizeof(*mesh));
	}
}

namespace upg
{
	int field_meshes_slab_device(Planet *planet, unsigned field_mask, int slab, int slabCount, upg_device_mesh *out3)
	{
		return field_meshes_impl(reinterpret_cast<upg_planet *>(planet), field_mask, 0, slab, slabCount, nullptr, out3);
	}
}
