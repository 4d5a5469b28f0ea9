// upg_mem.h -- device scratch memory of libupg.so: ONE library-owned stream-ordered pool per device.
//
// Every per-call scratch array (the res^3 sample fields, the isosurface and clean-up work arrays, the float images of a bake)
// comes from here.  The pool is created by the library (cudaMemPoolCreate) instead of re-configuring the device's default pool:
// co-resident allocators (torch's caching allocator in sharded.py / bench.py, the generator's own CUDA code if it ever has any)
// keep their default-pool behaviour, and what the library caches is visible and controllable through upg_reserve / upg_trim
// (include/upg.h).  Freed blocks stay in the pool (release threshold = everything) so that the next call of a batch finds them.
//
// UPG_POISON=<byte, hex> fills every allocation with that byte before it is handed out (debugging aid: a kernel that reads
// scratch memory it never wrote then produces garbage deterministically instead of depending on what the pool held before;
// tests/test_gpu_mesh.py runs the mesh entry points under two different poisons and compares the meshes bit for bit).
#pragma once

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace upg
{
	// stream-ordered allocation from the library's pool of the CURRENT device
	cudaError_t device_alloc_bytes(void **p, size_t bytes, cudaStream_t s);
	template<class T>
	inline cudaError_t device_alloc(T **p, size_t bytes, cudaStream_t s)
	{
		return device_alloc_bytes(reinterpret_cast<void **>(p), bytes, s);
	}
	inline cudaError_t device_free(void *p, cudaStream_t s) { return cudaFreeAsync(p, s); }

	// the calling thread's last request the pool could not serve, even after handing its cached blocks back (device_alloc retries once
	// after cudaMemPoolTrimTo): what was asked for next to what the pool and the device held -- part of the error text
	struct FailedRequest
	{
		uint64_t bytes = 0, reserved = 0, used = 0, deviceFree = 0, deviceTotal = 0;
		int device = -1;
	};
	FailedRequest &last_failed_request();

	// the pool of `device` (created on first use; nullptr + error when the device is unusable)
	cudaError_t device_pool(int device, cudaMemPool_t *pool);
	// grows the pool of `device` to at least `bytes` of cached memory / gives everything cached back to the driver
	cudaError_t device_pool_reserve(int device, uint64_t bytes);
	cudaError_t device_pool_trim(int device, uint64_t keepBytes);
	// bytes the pool currently holds from the driver, and the high-water mark of bytes handed out
	cudaError_t device_pool_stats(int device, uint64_t *reserved, uint64_t *usedHigh);
}
