// upg_multi.cu -- one planet on SEVERAL GPUs of one node from ONE host process (upg_planet_desc::device_count > 1).
//
// The reference generator is a single process whose pool threads run the three field pipelines and the chunk bakes
// (sources/generator.cpp:432-439, :159); this layer lets that one process spread a planet over the GPUs of an 8 x B200 box through the
// same C calls -- no MPI, no torch.distributed (unnatural_planets_b200/sharded.py stays as the one-process-per-GPU variant over NCCL).
// The path shards without a data-path collective:
//
//   base meshes  every device samples and extracts ONE z-slab of the grid (upg_field_meshes_slab_device: own slices + a recomputed
//                3-slice halo, one fused sampling for the three fields).  The host adds up the slabs' vertex / quad / triangle counts
//                (a few words per device) -- the only "exchange" --
//                * raw meshes (upg_field_meshes): every device rebases its chunk's vertex ids and copies its chunk STRAIGHT INTO THE
//                  FINAL pinned host arrays at its offset, over its own PCIe link: no gather to one GPU, no second copy;
//                * cleaned meshes (upg_field_meshes_clean): K9's connected components are global, so the chunks of a field meet in
//                  the HBM of ONE device -- peer copies over NVLink (cudaMemcpyPeerAsync, pool memory mapped to the peers) -- but each
//                  FIELD on a different device, so the three clean-ups and the three downloads run on three GPUs / three PCIe links.
//   bake         texel rows split evenly; every device uploads 1 / N of the mesh arrays over its own PCIe link, fetches the other
//                parts from its peers over NVLink, bakes its rows (+ the dilation apron, upg_bake.cu) and writes them straight into
//                the caller's host images.
// Results are bit-identical to one device (tests/test_gpu_multi.py).
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/upg.h"
#include "upg_internal.h"

namespace upg
{
	namespace
	{
		// all threads of a fan-out meet here (C++17 has no std::barrier)
		struct Rendezvous
		{
			std::mutex m;
			std::condition_variable cv;
			int expected, waiting = 0, generation = 0;
			explicit Rendezvous(int n) : expected(n) {}
			void wait()
			{
				std::unique_lock<std::mutex> lock(m);
				const int g = generation;
				if (++waiting == expected)
				{
					waiting = 0;
					generation++;
					cv.notify_all();
				}
				else
					cv.wait(lock, [&] { return generation != g; });
			}
		};

		struct Errors
		{
			std::mutex m;
			int code = UPG_OK;
			std::string text;
			void set(int rc)
			{
				if (rc == UPG_OK)
					return;
				std::lock_guard<std::mutex> lock(m);
				if (code == UPG_OK)
				{
					code = rc;
					text = upg_last_error();
				}
			}
			bool failed()
			{
				std::lock_guard<std::mutex> lock(m);
				return code != UPG_OK;
			}
			int result()
			{
				return code == UPG_OK ? UPG_OK : fail(code, "%s", text.c_str());
			}
		};

		// How the devices exchange data inside a call.
		//   peer  (two devices): NVLink copies between pool allocations (cudaMemcpyPeerAsync, the pools mapped to the peer).  Measured and
		//         checked bit for bit on 2 x B200 (profiles/r02_g_multi_probe_2gpu.log, tests/test_gpu_multi.py).
		//   host  (more than two): every device talks to pinned host memory over its OWN PCIe link and nothing is mapped to a peer.  The
		//         8-GPU run of round 2 (profiles/r02_m_multi_gpu8.md) showed the peer variant failing at 4 and 8 devices on the SECOND call
		//         of a 1024^3 planet -- cudaErrorMemoryAllocation inside a child's extraction with > 170 GB free, every pool mapped to 3 / 7
		//         peers -- while the small grids of the tests passed on 3, 4 and 8 devices; the GPU budget ended before the cause was found.
		//         The host variant uses nothing but cudaMemcpyAsync between device and pinned memory; it costs a few ms per call (two
		//         PCIe crossings of a ~100 MB mesh instead of one NVLink hop) and is the default beyond two devices until the peer variant
		//         is understood.  UPG_MULTI_EXCHANGE=peer|host overrides the choice.
		bool peer_exchange(int devices)
		{
			static const int forced = []()
			{
				const char *env = std::getenv("UPG_MULTI_EXCHANGE");
				if (!env)
					return 0;
				return std::strcmp(env, "peer") == 0 ? 1 : std::strcmp(env, "host") == 0 ? 2 : 0;
			}();
			return forced == 1 || (forced == 0 && devices <= 2);
		}

		// runs fn(i) on n host threads (i = 0 on the calling thread)
		template<class F>
		void fan_out(int n, F fn)
		{
			std::vector<std::thread> pool;
			for (int i = 1; i < n; i++)
				pool.emplace_back([&fn, i]() { fn(i); });
			fn(0);
			for (auto &t : pool)
				t.join();
		}
	}

	int multi_setup(Planet *p, const upg_planet_desc *desc)
	{
		// one single-device planet per further device: same modes, same seed stream (so the same seed table), same grid
		for (int i = 1; i < p->deviceCount; i++)
		{
			upg_planet_desc d = *desc;
			d.device = p->devices[i];
			d.device_count = 0;
			upg_planet *child = nullptr;
			const int rc = upg_planet_create(&d, &child);
			if (rc != UPG_OK)
				return rc;
			p->children[i] = reinterpret_cast<Planet *>(child);
			// overrides taken from the environment at creation must agree across the devices
			p->children[i]->sampling = p->sampling;
		}
		// the pools of all devices readable / writable by their peers: the chunk gather and the mesh exchange are NVLink copies
		// between pool allocations.  Devices without peer access fall back to copies staged by the driver (correct, slower).
		for (int i = 0; i < p->deviceCount && peer_exchange(p->deviceCount); i++)
		{
			cudaMemPool_t pool = nullptr;
			cudaError_t e = cudaSetDevice(p->devices[i]);
			if (e == cudaSuccess)
				e = device_pool(p->devices[i], &pool);
			if (e != cudaSuccess)
				return cuda_fail(e, "multi-device setup");
			for (int j = 0; j < p->deviceCount; j++)
			{
				if (i == j)
					continue;
				int can = 0;
				if (cudaDeviceCanAccessPeer(&can, p->devices[j], p->devices[i]) != cudaSuccess || !can)
				{
					cudaGetLastError();
					continue;
				}
				cudaMemAccessDesc access = {};
				access.location.type = cudaMemLocationTypeDevice;
				access.location.id = p->devices[j];
				access.flags = cudaMemAccessFlagsProtReadWrite;
				if (cudaMemPoolSetAccess(pool, &access, 1) != cudaSuccess)
					cudaGetLastError();
			}
		}
		cudaSetDevice(p->device);
		return UPG_OK;
	}

	void multi_teardown(Planet *p)
	{
		for (int i = 1; i < 8; i++)
			if (p->children[i])
			{
				upg_planet_destroy(reinterpret_cast<upg_planet *>(p->children[i]));
				p->children[i] = nullptr;
			}
	}

	int multi_field_meshes(Planet *p, unsigned field_mask, unsigned stages, upg_mesh *out3)
	{
		if (!out3)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		std::memset(out3, 0, sizeof(upg_mesh) * 3);
		if (stages & ~(unsigned)UPG_MESH_GENERATOR)
			return fail(UPG_ERR_ARGUMENT, "bad clean-up stages %#x", stages);
		if ((field_mask & 7u) == 0 || (field_mask & ~7u))
			return fail(UPG_ERR_ARGUMENT, "bad field mask %#x (bit 0 land, bit 1 water, bit 2 navigation)", field_mask);
		const int N = p->deviceCount;
		const bool trace = trace_enabled();
		const double t0 = trace ? now_ms() : 0;
		// ---- phase 1: every device samples and extracts its z-slab; the chunks stay in its HBM -------------------------------------
		std::vector<upg_device_mesh> chunks((size_t)N * 3);
		std::memset(chunks.data(), 0, sizeof(upg_device_mesh) * chunks.size());
		Errors errors;
		fan_out(N, [&](int d) { errors.set(field_meshes_slab_device(p->children[d], field_mask, d, N, &chunks[(size_t)d * 3])); });
		const double t1 = trace ? now_ms() : 0;
		// ---- the only exchange: per-slab counts -> offsets of every chunk in the stitched mesh --------------------------------------
		struct Offsets
		{
			uint64_t vertex[9], index[9], quad[9]; // prefix sums over the devices (N <= 8), [N] = total
		};
		Offsets off[3];
		for (int f = 0; f < 3; f++)
		{
			off[f].vertex[0] = off[f].index[0] = off[f].quad[0] = 0;
			uint64_t near = 0;
			for (int d = 0; d < N; d++)
			{
				const upg_device_mesh &c = chunks[(size_t)d * 3 + f];
				off[f].vertex[d + 1] = off[f].vertex[d] + c.vertex_count;
				off[f].index[d + 1] = off[f].index[d] + c.index_count;
				off[f].quad[d + 1] = off[f].quad[d] + c.quad_count;
				near += c.near_zero_samples;
			}
			out3[f].near_zero_samples = near;
			if (!errors.failed() && (off[f].vertex[N] >= 0xffffffffull || off[f].index[N] >= 0xffffffffull))
				errors.set(fail(UPG_ERR_ARGUMENT, "stitched mesh too large"));
		}
		if (!errors.failed() && stages == 0)
		{
			// ---- raw meshes: every device copies its chunks straight into the final pinned host arrays, over its own PCIe link ---------
			char *blocks[3] = { nullptr, nullptr, nullptr };
			size_t capacity[3] = { 0, 0, 0 };
			for (int f = 0; f < 3 && !errors.failed(); f++)
			{
				if (!(field_mask & (1u << f)))
					continue;
				const uint64_t nv = off[f].vertex[N], ni = off[f].index[N], nq = off[f].quad[N];
				if (ni == 0)
				{
					errors.set(fail(UPG_ERR_EMPTY_MESH, "generated empty base %s mesh", f == FD_LAND ? "land" : f == FD_WATER ? "water" : "navigation"));
					break;
				}
				const uint64_t bytes = sizeof(float) * nv * 9 + sizeof(uint32_t) * (ni + nq * 4);
				const cudaError_t e = pinned_block_acquire(bytes ? bytes : 16, (void **)&blocks[f], &capacity[f]);
				if (e != cudaSuccess)
				{
					errors.set(cuda_fail(e, "cudaMallocHost(mesh)"));
					break;
				}
				char *at = blocks[f];
				out3[f].positions = (float *)at;
				at += sizeof(float) * nv * 3;
				out3[f].normals = (float *)at;
				at += sizeof(float) * nv * 3;
				out3[f].dmc_vertices = (float *)at;
				at += sizeof(float) * nv * 3;
				out3[f].indices = (uint32_t *)at;
				at += sizeof(uint32_t) * ni;
				out3[f].dmc_quads = nq ? (uint32_t *)at : nullptr;
				out3[f].vertex_count = (uint32_t)nv;
				out3[f].index_count = (uint32_t)ni;
				out3[f].quad_count = (uint32_t)nq;
				out3[f].owner = blocks[f];
				out3[f].owner_bytes = capacity[f];
			}
			if (!errors.failed())
				fan_out(N, [&](int d)
					{
						cudaError_t e = cudaSetDevice(p->devices[d]);
						cudaStream_t s = nullptr;
						if (e == cudaSuccess)
							e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
						int rc = e == cudaSuccess ? UPG_OK : cuda_fail(e, "cudaStreamCreate");
						for (int f = 0; f < 3 && rc == UPG_OK; f++)
						{
							const upg_device_mesh &c = chunks[(size_t)d * 3 + f];
							if (!(field_mask & (1u << f)) || c.vertex_count + c.index_count == 0)
								continue;
							const uint32_t base = (uint32_t)off[f].vertex[d];
							rc = rebase_ids(c.indices, c.index_count, base, s);
							if (rc == UPG_OK && c.dmc_quads)
								rc = rebase_ids(c.dmc_quads, (uint64_t)c.quad_count * 4, base, s);
							if (rc != UPG_OK)
								break;
							const uint64_t v0 = off[f].vertex[d] * 3;
							if (c.vertex_count)
							{
								e = cudaMemcpyAsync(out3[f].positions + v0, c.positions, sizeof(float) * 3 * c.vertex_count, cudaMemcpyDeviceToHost, s);
								if (e == cudaSuccess)
									e = cudaMemcpyAsync(out3[f].normals + v0, c.normals, sizeof(float) * 3 * c.vertex_count, cudaMemcpyDeviceToHost, s);
								if (e == cudaSuccess)
									e = cudaMemcpyAsync(out3[f].dmc_vertices + v0, c.dmc_vertices, sizeof(float) * 3 * c.vertex_count, cudaMemcpyDeviceToHost, s);
							}
							if (e == cudaSuccess && c.index_count)
								e = cudaMemcpyAsync(out3[f].indices + off[f].index[d], c.indices, sizeof(uint32_t) * c.index_count, cudaMemcpyDeviceToHost, s);
							if (e == cudaSuccess && c.quad_count)
								e = cudaMemcpyAsync(out3[f].dmc_quads + off[f].quad[d] * 4, c.dmc_quads, sizeof(uint32_t) * 4 * c.quad_count, cudaMemcpyDeviceToHost, s);
							if (e != cudaSuccess)
								rc = cuda_fail(e, "mesh download");
						}
						if (s)
						{
							e = stream_sync(s);
							if (rc == UPG_OK && e != cudaSuccess)
								rc = cuda_fail(e, "mesh download");
							cudaStreamDestroy(s);
						}
						errors.set(rc);
					});
		}
		else if (!errors.failed())
		{
			// ---- cleaned meshes: the chunks of field f meet on device f mod N (NVLink), K9 and the download run there ------------------
			int fields[3], nf = 0;
			for (int f = 0; f < 3; f++)
				if (field_mask & (1u << f))
					fields[nf++] = f;
			fan_out(nf, [&](int k)
				{
					const int f = fields[k];
					const int home = f % N;
					Planet *hp = p->children[home];
					const uint64_t nv = off[f].vertex[N], ni = off[f].index[N];
					int rc = UPG_OK;
					if (ni == 0)
					{
						// the reference throws for an empty land / navigation mesh only (sources/meshGeneration.cpp:51-52, 104-105)
						if (f != FD_WATER)
							rc = fail(UPG_ERR_EMPTY_MESH, "generated empty base %s mesh", f == FD_LAND ? "land" : "navigation");
						errors.set(rc);
						return;
					}
					cudaError_t e = cudaSetDevice(hp->device);
					CallStreams cs;
					if (e != cudaSuccess)
						rc = cuda_fail(e, "cudaSetDevice");
					if (rc == UPG_OK)
						rc = cs.open();
					StatusSlot slot;
					if (rc == UPG_OK)
						rc = slot.open(hp, cs.compute);
					float *pos = nullptr, *nrm = nullptr;
					uint32_t *idx = nullptr;
					if (rc == UPG_OK)
					{
						e = device_alloc(&pos, sizeof(float) * 3 * nv, cs.compute);
						if (e == cudaSuccess)
							e = device_alloc(&nrm, sizeof(float) * 3 * nv, cs.compute);
						if (e == cudaSuccess)
							e = device_alloc(&idx, sizeof(uint32_t) * ni, cs.compute);
						if (e != cudaSuccess)
							rc = cuda_fail(e, "device_alloc(gathered mesh)");
					}
					void *stage = nullptr; // host variant: the field's chunks meet in one pinned block
					size_t stageCapacity = 0;
					if (rc == UPG_OK && !peer_exchange(N))
					{
						const size_t vb = sizeof(float) * 3 * (size_t)nv, ib = sizeof(uint32_t) * (size_t)ni;
						e = pinned_block_acquire(vb * 2 + ib, &stage, &stageCapacity);
						if (e != cudaSuccess)
							rc = cuda_fail(e, "cudaMallocHost(mesh exchange)");
						char *hp0 = (char *)stage, *hn0 = hp0 + vb, *hi0 = hn0 + vb;
						// every chunk down its own device's link ...
						std::vector<cudaStream_t> down((size_t)N, nullptr);
						for (int d = 0; d < N && rc == UPG_OK; d++)
						{
							const upg_device_mesh &c = chunks[(size_t)d * 3 + f];
							if (c.vertex_count + c.index_count == 0)
								continue;
							e = cudaSetDevice(c.device);
							if (e == cudaSuccess)
								e = cudaStreamCreateWithFlags(&down[d], cudaStreamNonBlocking);
							const size_t v0b = sizeof(float) * 3 * (size_t)off[f].vertex[d];
							if (e == cudaSuccess && c.vertex_count)
								e = cudaMemcpyAsync(hp0 + v0b, c.positions, sizeof(float) * 3 * (size_t)c.vertex_count, cudaMemcpyDeviceToHost, down[d]);
							if (e == cudaSuccess && c.vertex_count)
								e = cudaMemcpyAsync(hn0 + v0b, c.normals, sizeof(float) * 3 * (size_t)c.vertex_count, cudaMemcpyDeviceToHost, down[d]);
							if (e == cudaSuccess && c.index_count)
								e = cudaMemcpyAsync(hi0 + sizeof(uint32_t) * (size_t)off[f].index[d], c.indices, sizeof(uint32_t) * (size_t)c.index_count, cudaMemcpyDeviceToHost, down[d]);
							if (e != cudaSuccess)
								rc = cuda_fail(e, "mesh chunk download");
						}
						for (int d = 0; d < N; d++)
							if (down[d])
							{
								cudaSetDevice(chunks[(size_t)d * 3 + f].device);
								e = cudaStreamSynchronize(down[d]);
								if (e != cudaSuccess && rc == UPG_OK)
									rc = cuda_fail(e, "mesh chunk download");
								cudaStreamDestroy(down[d]);
							}
						e = cudaSetDevice(hp->device);
						if (e != cudaSuccess && rc == UPG_OK)
							rc = cuda_fail(e, "cudaSetDevice");
						// ... and the whole field up the home device's link
						if (rc == UPG_OK)
						{
							e = cudaMemcpyAsync(pos, hp0, vb, cudaMemcpyHostToDevice, cs.compute);
							if (e == cudaSuccess)
								e = cudaMemcpyAsync(nrm, hn0, vb, cudaMemcpyHostToDevice, cs.compute);
							if (e == cudaSuccess)
								e = cudaMemcpyAsync(idx, hi0, ib, cudaMemcpyHostToDevice, cs.compute);
							if (e != cudaSuccess)
								rc = cuda_fail(e, "gathered mesh upload");
						}
						for (int d = 0; d < N && rc == UPG_OK; d++)
							rc = rebase_ids(idx + off[f].index[d], chunks[(size_t)d * 3 + f].index_count, (uint32_t)off[f].vertex[d], cs.compute);
					}
					for (int d = 0; d < N && rc == UPG_OK && peer_exchange(N); d++)
					{
						const upg_device_mesh &c = chunks[(size_t)d * 3 + f];
						const uint64_t v0 = off[f].vertex[d] * 3;
						if (c.vertex_count)
						{
							e = cudaMemcpyPeerAsync(pos + v0, hp->device, c.positions, c.device, sizeof(float) * 3 * c.vertex_count, cs.compute);
							if (e == cudaSuccess)
								e = cudaMemcpyPeerAsync(nrm + v0, hp->device, c.normals, c.device, sizeof(float) * 3 * c.vertex_count, cs.compute);
						}
						if (e == cudaSuccess && c.index_count)
							e = cudaMemcpyPeerAsync(idx + off[f].index[d], hp->device, c.indices, c.device, sizeof(uint32_t) * c.index_count, cs.compute);
						if (e != cudaSuccess)
							rc = cuda_fail(e, "cudaMemcpyPeerAsync(mesh chunk)");
						if (rc == UPG_OK)
							rc = rebase_ids(idx + off[f].index[d], c.index_count, (uint32_t)off[f].vertex[d], cs.compute);
					}
					DmcResult cleaned;
					if (rc == UPG_OK)
						rc = mesh_clean(hp, pos, nrm, (uint32_t)nv, idx, (uint32_t)(ni / 3), p->boxSize, f == FD_WATER ? stages : (stages & ~(unsigned)UPG_MESH_WATER_CULL), cs.compute, cleaned);
					if (rc == UPG_OK)
						rc = check_status(hp, cs.compute, "upg_field_meshes_clean");
					if (rc == UPG_OK && cleaned.triangleCount == 0 && f != FD_WATER)
						rc = fail(UPG_ERR_EMPTY_MESH, "generated empty base %s mesh", f == FD_LAND ? "land" : "navigation");
					const uint64_t near = out3[f].near_zero_samples;
					if (rc == UPG_OK)
					{
						cleaned.nearZeroSamples = near;
						rc = mesh_to_host(cleaned, cs.compute, &out3[f]);
					}
					dmc_release(cleaned, cs.compute);
					if (pos)
						device_free(pos, cs.compute);
					if (nrm)
						device_free(nrm, cs.compute);
					if (idx)
						device_free(idx, cs.compute);
					if (cs.compute)
						stream_sync(cs.compute);
					if (stage)
						pinned_block_release(stage, stageCapacity);
					cs.close();
					errors.set(rc);
				});
		}
		// the slab chunks go back to their pools
		for (auto &c : chunks)
			upg_device_mesh_free(&c);
		cudaSetDevice(p->device);
		if (trace)
			std::fprintf(stderr, "[upg] multi-device field_meshes(mask %#x, stages %#x) over %d devices: slabs %.2f ms, stitch %s %.2f ms\n", field_mask, stages, N, t1 - t0, stages ? "(gather + clean-up + download)" : "(downloads)", now_ms() - t1);
		if (errors.failed())
		{
			for (int f = 0; f < 3; f++)
				upg_mesh_free(&out3[f]);
			return errors.result();
		}
		return UPG_OK;
	}

	int multi_bake(Planet *p, int water, const float *positions, const float *normals, const float *uvs, uint32_t nverts, const uint32_t *idx, uint32_t nidx, uint32_t w, uint32_t h, uint32_t row0, uint32_t row1, uint8_t *albedo, uint8_t *special, uint8_t *heightMap, uint8_t *pbr, uint8_t *normalMap)
	{
		const int N = p->deviceCount;
		const bool trace = trace_enabled();
		const double t0 = trace ? now_ms() : 0;
		// the four mesh arrays, cut into N parts at element boundaries
		struct Part
		{
			const char *host;
			size_t bytes; // whole array
			size_t element;
		};
		const Part arrays[4] = { { (const char *)uvs, sizeof(float) * 2 * (size_t)nverts, 8 }, { (const char *)idx, sizeof(uint32_t) * (size_t)nidx, 4 }, { (const char *)positions, sizeof(float) * 3 * (size_t)nverts, 12 }, { (const char *)normals, sizeof(float) * 3 * (size_t)nverts, 12 } };
		auto cut = [&](int a, int part) -> size_t
		{
			const size_t elements = arrays[a].bytes / arrays[a].element;
			return (elements * (size_t)part / (size_t)N) * arrays[a].element;
		};
		std::vector<char *> dev((size_t)N * 4, nullptr); // dev[d * 4 + a]: array a in the memory of device d
		std::vector<cudaStream_t> streams((size_t)N, nullptr);
		std::vector<cudaEvent_t> uploaded((size_t)N, nullptr);
		std::vector<uint64_t> covered((size_t)N, 0);
		std::vector<double> tUp((size_t)N, 0), tGather((size_t)N, 0);
		Errors errors;
		Rendezvous allocated(N), sliceUp(N), done(N);
		const int ach = water ? 4 : 3;
		const bool peer = peer_exchange(N);
		fan_out(N, [&](int d)
			{
				Planet *c = p->children[d];
				int rc = UPG_OK;
				cudaError_t e = cudaSetDevice(c->device);
				if (e == cudaSuccess)
					e = cudaStreamCreateWithFlags(&streams[d], cudaStreamNonBlocking);
				if (e == cudaSuccess)
					e = cudaEventCreateWithFlags(&uploaded[d], cudaEventDisableTiming);
				for (int a = 0; a < 4 && e == cudaSuccess; a++)
					e = device_alloc(&dev[(size_t)d * 4 + a], arrays[a].bytes, streams[d]);
				if (e != cudaSuccess)
					rc = cuda_fail(e, "multi-device bake setup");
				errors.set(rc);
				allocated.wait(); // every device's arrays exist (their addresses are read by the peers below)
				// own part of every array from the host, over this device's PCIe link
				void *staging = nullptr;
				size_t stagingCapacity = 0;
				if (!errors.failed())
				{
					HostArray mine[4];
					for (int a = 0; a < 4; a++)
					{
						// host variant: the whole mesh up this device's own link, nothing to fetch from the peers
						const size_t b0 = peer ? cut(a, d) : 0, b1 = peer ? cut(a, d + 1) : arrays[a].bytes;
						mine[a] = HostArray{ dev[(size_t)d * 4 + a] + b0, arrays[a].host + b0, b1 - b0 };
					}
					rc = upload_host_arrays(mine, 4, streams[d], &staging, &stagingCapacity);
					if (rc == UPG_OK && cudaEventRecord(uploaded[d], streams[d]) != cudaSuccess)
						rc = cuda_fail(cudaGetLastError(), "cudaEventRecord");
					errors.set(rc);
				}
				if (trace)
					tUp[d] = now_ms();
				sliceUp.wait(); // every device's upload is enqueued and its event recorded
				// the other parts from the peers, over NVLink, ordered after each peer's own upload
				if (!errors.failed())
				{
					for (int k = 1; k < N && rc == UPG_OK && peer; k++)
					{
						const int o = (d + k) % N; // start with the next device: the N pulls spread over the links
						e = cudaStreamWaitEvent(streams[d], uploaded[o], 0);
						for (int a = 0; a < 4 && e == cudaSuccess; a++)
						{
							const size_t b0 = cut(a, o), b1 = cut(a, o + 1);
							if (b1 > b0)
								e = cudaMemcpyPeerAsync(dev[(size_t)d * 4 + a] + b0, c->device, dev[(size_t)o * 4 + a] + b0, p->devices[o], b1 - b0, streams[d]);
						}
						if (e != cudaSuccess)
							rc = cuda_fail(e, "cudaMemcpyPeerAsync(mesh part)");
					}
					if (trace && rc == UPG_OK)
					{
						stream_sync(streams[d]);
						tGather[d] = now_ms();
					}
				}
				// own rows; the outputs are the caller's host images (the core copies with cudaMemcpyDefault)
				if (!errors.failed() && rc == UPG_OK)
				{
					const uint32_t rows = row1 - row0;
					const uint32_t r0 = row0 + (uint32_t)((uint64_t)rows * d / N), r1 = row0 + (uint32_t)((uint64_t)rows * (d + 1) / N);
					const size_t skip = (size_t)(r0 - row0) * w;
					rc = bake_core(reinterpret_cast<upg_planet *>(c), water, (const float *)dev[(size_t)d * 4 + 2], (const float *)dev[(size_t)d * 4 + 3], (const float *)dev[(size_t)d * 4 + 0], nverts, (const uint32_t *)dev[(size_t)d * 4 + 1], nidx, w, h, r0, r1,
						albedo + skip * ach, special ? special + skip * 2 : nullptr, heightMap ? heightMap + skip : nullptr, pbr ? pbr + skip * 3 : nullptr, normalMap ? normalMap + skip * 3 : nullptr, 1.f, streams[d]);
					covered[d] = bake_thread_covered();
				}
				errors.set(rc);
				if (streams[d])
					stream_sync(streams[d]);
				if (staging)
					pinned_block_release(staging, stagingCapacity);
				done.wait(); // nobody frees its arrays while a peer may still be reading them
				for (int a = 0; a < 4; a++)
					if (dev[(size_t)d * 4 + a])
						device_free(dev[(size_t)d * 4 + a], streams[d]);
				if (streams[d])
				{
					stream_sync(streams[d]);
					cudaStreamDestroy(streams[d]);
				}
				if (uploaded[d])
					cudaEventDestroy(uploaded[d]);
			});
		uint64_t total = 0;
		for (int d = 0; d < N; d++)
			total += covered[d];
		bake_set_thread_covered(total);
		cudaSetDevice(p->device);
		if (trace)
			std::fprintf(stderr, "[upg] multi-device bake %s %ux%u rows [%u, %u) over %d devices: own part uploaded after %.2f ms, parts of the peers after %.2f ms, call %.2f ms\n", water ? "water" : "land", w, h, row0, row1, N, tUp[0] - t0, tGather[0] - t0, now_ms() - t0);
		return errors.result();
	}
}
