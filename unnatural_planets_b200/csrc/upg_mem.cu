// upg_mem.cu -- the library-owned stream-ordered memory pools (see upg_mem.h)
#include "upg_mem.h"

#include <cstdlib>
#include <mutex>

namespace upg
{
	namespace
	{
		constexpr int MAX_DEVICES = 64;
		std::mutex g_poolMutex;
		cudaMemPool_t g_pools[MAX_DEVICES] = {};

		// UPG_POISON=a5 -> 0xa5; unset -> -1
		int poison_byte()
		{
			static const int value = []()
			{
				const char *env = std::getenv("UPG_POISON");
				if (!env || !env[0])
					return -1;
				return (int)(std::strtoul(env, nullptr, 16) & 0xffu);
			}();
			return value;
		}
	}

	FailedRequest &last_failed_request()
	{
		static thread_local FailedRequest request;
		return request;
	}

	cudaError_t device_pool(int device, cudaMemPool_t *pool)
	{
		*pool = nullptr;
		if (device < 0 || device >= MAX_DEVICES)
			return cudaErrorInvalidDevice;
		std::lock_guard<std::mutex> lock(g_poolMutex);
		if (!g_pools[device])
		{
			cudaMemPoolProps props = {};
			props.allocType = cudaMemAllocationTypePinned;
			props.handleTypes = cudaMemHandleTypeNone;
			props.location.type = cudaMemLocationTypeDevice;
			props.location.id = device;
			cudaMemPool_t created = nullptr;
			cudaError_t e = cudaMemPoolCreate(&created, &props);
			if (e != cudaSuccess)
				return e;
			// keep freed blocks: the per-call scratch (up to 3-4 x res^3 floats) is re-used by the next call instead of going back to
			// the driver at every synchronisation; upg_trim hands it back on request
			uint64_t threshold = UINT64_MAX;
			e = cudaMemPoolSetAttribute(created, cudaMemPoolAttrReleaseThreshold, &threshold);
			if (e != cudaSuccess)
			{
				cudaMemPoolDestroy(created);
				return e;
			}
			// UPG_POOL_REUSE_OFF=<bits>: switch off the pool's cross-stream reuse policies (1 follow event dependencies, 2 opportunistic,
			// 4 internal dependencies) -- debugging aid for ordering problems between the streams of concurrent calls
			if (const char *env = std::getenv("UPG_POOL_REUSE_OFF"))
			{
				const unsigned long off = std::strtoul(env, nullptr, 0);
				int zero = 0;
				if (off & 1)
					cudaMemPoolSetAttribute(created, cudaMemPoolReuseFollowEventDependencies, &zero);
				if (off & 2)
					cudaMemPoolSetAttribute(created, cudaMemPoolReuseAllowOpportunistic, &zero);
				if (off & 4)
					cudaMemPoolSetAttribute(created, cudaMemPoolReuseAllowInternalDependencies, &zero);
			}
			g_pools[device] = created;
		}
		*pool = g_pools[device];
		return cudaSuccess;
	}

	cudaError_t device_alloc_bytes(void **p, size_t bytes, cudaStream_t s)
	{
		*p = nullptr;
		int device = 0;
		cudaError_t e = cudaGetDevice(&device);
		if (e != cudaSuccess)
			return e;
		cudaMemPool_t pool = nullptr;
		e = device_pool(device, &pool);
		if (e != cudaSuccess)
			return e;
		e = cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, s);
		if (e == cudaErrorMemoryAllocation)
		{
			// what a caching allocator does before it gives up: let the frees queued on this stream take effect, hand every cached
			// block back to the driver (blocks the other streams of concurrent calls hold cannot serve this one) and ask once more
			cudaGetLastError();
			cudaStreamSynchronize(s);
			cudaMemPoolTrimTo(pool, 0);
			e = cudaMallocFromPoolAsync(p, bytes ? bytes : 1, pool, s);
			if (e != cudaSuccess)
			{
				// the failed request next to what the pool and the device held, for the error text (cuda_fail)
				FailedRequest &f = last_failed_request();
				f.bytes = bytes;
				f.device = device;
				f.reserved = f.used = f.deviceFree = f.deviceTotal = 0;
				cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &f.reserved);
				cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &f.used);
				size_t freeBytes = 0, totalBytes = 0;
				if (cudaMemGetInfo(&freeBytes, &totalBytes) == cudaSuccess)
				{
					f.deviceFree = freeBytes;
					f.deviceTotal = totalBytes;
				}
				cudaGetLastError();
			}
		}
		if (e != cudaSuccess)
			return e;
		const int poison = poison_byte();
		if (poison >= 0)
			e = cudaMemsetAsync(*p, poison, bytes ? bytes : 1, s);
		return e;
	}

	cudaError_t device_pool_reserve(int device, uint64_t bytes)
	{
		int previous = 0;
		cudaError_t e = cudaGetDevice(&previous);
		if (e == cudaSuccess)
			e = cudaSetDevice(device);
		if (e != cudaSuccess)
			return e;
		uint64_t reserved = 0, used = 0;
		e = device_pool_stats(device, &reserved, &used);
		if (e == cudaSuccess && reserved < bytes)
		{
			// one block of the missing size, freed again at once: the pool keeps it (release threshold)
			void *p = nullptr;
			cudaMemPool_t pool = nullptr;
			e = device_pool(device, &pool);
			if (e == cudaSuccess)
				e = cudaMallocFromPoolAsync(&p, bytes - reserved, pool, nullptr);
			if (e == cudaSuccess)
				e = cudaFreeAsync(p, nullptr);
			if (e == cudaSuccess)
				e = cudaStreamSynchronize(nullptr);
		}
		cudaSetDevice(previous);
		return e;
	}

	cudaError_t device_pool_trim(int device, uint64_t keepBytes)
	{
		cudaMemPool_t pool = nullptr;
		cudaError_t e = device_pool(device, &pool);
		if (e == cudaSuccess)
			e = cudaMemPoolTrimTo(pool, (size_t)keepBytes);
		return e;
	}

	cudaError_t device_pool_stats(int device, uint64_t *reserved, uint64_t *usedHigh)
	{
		*reserved = *usedHigh = 0;
		cudaMemPool_t pool = nullptr;
		cudaError_t e = device_pool(device, &pool);
		if (e == cudaSuccess)
			e = cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, reserved);
		if (e == cudaSuccess)
			e = cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemHigh, usedHigh);
		return e;
	}
}
