// upg_png.cu -- K10: PNG encoding of the baked 8-bit textures on the device (third item).
//
// Replaces, for the chunk textures, Image::exportFile / exportBuffer(".png") as the generator's chunk tasks call it right after the
// texture tail (sources/generator.cpp:135-139, 187-191; cage: libcore/image/png.cpp -> libpng + zlib, both un-vendored submodules of
// the reference).  With the bake at a few milliseconds per 4096^2 chunk, cage's single-threaded deflate of three such images is the
// wall clock of a chunk; here the image never leaves the GPU uncompressed.
//
// A PNG is a container around ONE zlib stream of the filtered scanlines; nothing in the format says who compresses how, only that a
// decoder gets the pixels back (the parity bar of this kernel: decode(encode(image)) == image, tests/test_gpu_png.py decodes with
// zlib / PIL).  The encoder is laid out for the GPU:
//   png_filter    one block per scanline: the five PNG filter types are tried with libpng's heuristic (minimum sum of absolute
//                 differences) and the best is written, type byte first -- rows only read the RAW row above, so all rows go at once;
//   png_deflate   the filtered stream is cut into 8 KiB chunks; ONE THREAD compresses one chunk into ONE deflate block with the fixed
//                 Huffman code (greedy LZ77 over a 1024-entry hash table in shared memory, matches confined to the chunk), closes it
//                 with an empty stored block -- that byte-aligns the stream, so chunks concatenate without bit shifting (the pigz
//                 trick) -- and leaves the chunk's Adler-32 partial sums.  ~6000 independent chunks per 4096^2 x 3 image;
//   png_layout    one block: exclusive scan of the chunk sizes, Adler-32 of the whole stream from the partial sums;
//   png_emit      every chunk becomes its OWN IDAT chunk (the format allows any number): length, "IDAT", bytes, CRC-32 of that chunk
//                 alone -- no CRC has to be combined across chunks.  One block per chunk copies and checksums.
// Compression is what a fixed-code, 8 KiB-window deflate gives (measured in the tests / bench next to zlib); the point is throughput.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <mutex>

#include "../../include/upg.h"
#include "upg_internal.h"

namespace upg
{
	namespace
	{
		constexpr uint32_t PNG_CHUNK = 8192; // filtered bytes per deflate block / IDAT chunk
		constexpr uint32_t PNG_OUT_STRIDE = PNG_CHUNK + PNG_CHUNK / 8 + 64; // worst case 9 bits per literal + block framing, 4-byte aligned
		constexpr uint32_t PNG_HASH = 1024;
		constexpr int PNG_DEFLATE_THREADS = 64; // per block: 64 x 2 KiB hash tables in shared memory
		constexpr uint32_t PNG_HEADER = 8 + 25; // signature + IHDR chunk

		__constant__ uint32_t c_lenCode[259]; // match length 3..258 -> (bits << 8) | count: fixed-Huffman code (bit-reversed) + extra bits
		__constant__ uint32_t c_crcTable[256];

		__device__ __forceinline__ int paeth(int a, int b, int c)
		{
			const int p = a + b - c;
			const int pa = abs(p - a), pb = abs(p - b), pc = abs(p - c);
			return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
		}

		// ---- scanline filters (PNG spec 9.2), heuristic of libpng's PNG_ALL_FILTERS: smallest sum of |signed filtered byte| ----------
		__global__ void __launch_bounds__(256) png_filter_kernel(const uint8_t *__restrict__ pixels, uint32_t rowBytes, uint32_t bpp, uint32_t rows, uint8_t *__restrict__ filtered)
		{
			const uint32_t r = blockIdx.x;
			if (r >= rows)
				return;
			const uint8_t *cur = pixels + (size_t)r * rowBytes;
			const uint8_t *up = r ? cur - rowBytes : nullptr;
			uint32_t sums[5] = { 0, 0, 0, 0, 0 };
			for (uint32_t i = threadIdx.x; i < rowBytes; i += 256)
			{
				const int x = cur[i];
				const int a = i >= bpp ? cur[i - bpp] : 0;
				const int b = up ? up[i] : 0;
				const int c = (up && i >= bpp) ? up[i - bpp] : 0;
				const int f[5] = { x, x - a, x - b, x - ((a + b) >> 1), x - paeth(a, b, c) };
#pragma unroll
				for (int k = 0; k < 5; k++)
				{
					const int v = f[k] & 255;
					sums[k] += v < 128 ? v : 256 - v;
				}
			}
			__shared__ uint32_t red[5][8];
			__shared__ uint32_t choice;
#pragma unroll
			for (int k = 0; k < 5; k++)
			{
				uint32_t s = sums[k];
				for (int o = 16; o > 0; o >>= 1)
					s += __shfl_down_sync(0xffffffffu, s, o);
				if ((threadIdx.x & 31) == 0)
					red[k][threadIdx.x >> 5] = s;
			}
			__syncthreads();
			if (threadIdx.x == 0)
			{
				uint32_t best = 0, bestSum = 0xffffffffu;
				for (int k = 0; k < 5; k++)
				{
					uint32_t s = 0;
					for (int w = 0; w < 8; w++)
						s += red[k][w];
					if (s < bestSum)
					{
						bestSum = s;
						best = k;
					}
				}
				choice = best;
			}
			__syncthreads();
			const uint32_t type = choice;
			uint8_t *out = filtered + (size_t)r * (rowBytes + 1);
			if (threadIdx.x == 0)
				out[0] = (uint8_t)type;
			for (uint32_t i = threadIdx.x; i < rowBytes; i += 256)
			{
				const int x = cur[i];
				const int a = i >= bpp ? cur[i - bpp] : 0;
				const int b = up ? up[i] : 0;
				const int c = (up && i >= bpp) ? up[i - bpp] : 0;
				const int v = type == 0 ? x : type == 1 ? x - a : type == 2 ? x - b : type == 3 ? x - ((a + b) >> 1) : x - paeth(a, b, c);
				out[1 + i] = (uint8_t)v;
			}
		}

		// ---- deflate, one thread per chunk -----------------------------------------------------------------------------------------
		struct BitWriter
		{
			uint32_t *out; // 4-byte aligned
			uint64_t acc = 0;
			uint32_t bits = 0, words = 0;
			__device__ __forceinline__ void put(uint32_t value, uint32_t count)
			{
				acc |= (uint64_t)value << bits;
				bits += count;
				if (bits >= 32)
				{
					out[words++] = (uint32_t)acc;
					acc >>= 32;
					bits -= 32;
				}
			}
			// pads to a byte boundary, appends `tail` whole bytes given LSB first in `value`, returns the size in bytes
			__device__ __forceinline__ uint32_t finish()
			{
				bits = (bits + 7u) & ~7u;
				if (bits >= 32)
				{
					out[words++] = (uint32_t)acc;
					acc >>= 32;
					bits -= 32;
				}
				const uint32_t bytes = words * 4 + bits / 8;
				if (bits)
					out[words++] = (uint32_t)acc; // the unused high bytes of the last word are never copied out
				return bytes;
			}
		};

		__device__ __forceinline__ uint32_t reverse_bits(uint32_t v, uint32_t n) { return __brev(v) >> (32 - n); }

		__device__ __forceinline__ void put_literal(BitWriter &w, uint32_t v)
		{
			// fixed code (RFC 1951 3.2.6): 0..143 -> 8 bits 0x30 + v, 144..255 -> 9 bits 0x190 + (v - 144); Huffman codes go in MSB first
			if (v < 144)
				w.put(reverse_bits(0x30 + v, 8), 8);
			else
				w.put(reverse_bits(0x190 + (v - 144), 9), 9);
		}

		__device__ __forceinline__ void put_match(BitWriter &w, uint32_t length, uint32_t distance)
		{
			const uint32_t lc = c_lenCode[length];
			w.put(lc >> 8, lc & 255u);
			const uint32_t d1 = distance - 1;
			uint32_t code, eb = 0, extra = 0;
			if (d1 < 4)
				code = d1;
			else
			{
				const uint32_t t = 31u - (uint32_t)__clz((int)d1);
				code = 2 * t + ((d1 >> (t - 1)) & 1u);
				eb = t - 1;
				extra = d1 & ((1u << eb) - 1u);
			}
			w.put(reverse_bits(code, 5) | (extra << 5), 5 + eb); // 5-bit fixed distance code, then the extra bits LSB first
		}

		__global__ void __launch_bounds__(PNG_DEFLATE_THREADS) png_deflate_kernel(const uint8_t *__restrict__ filtered, uint64_t total, uint32_t chunks, uint8_t *__restrict__ scratch, uint32_t *__restrict__ sizes, uint32_t *__restrict__ adlerParts)
		{
			extern __shared__ uint16_t table[]; // PNG_DEFLATE_THREADS x PNG_HASH entries (128 KiB: dynamic, opted in by the launcher)
			const uint32_t chunk = blockIdx.x * PNG_DEFLATE_THREADS + threadIdx.x;
			if (chunk >= chunks)
				return;
			uint16_t *hash = table + (size_t)threadIdx.x * PNG_HASH;
			for (uint32_t i = 0; i < PNG_HASH; i += 2)
				*reinterpret_cast<uint32_t *>(hash + i) = 0;
			const uint64_t begin = (uint64_t)chunk * PNG_CHUNK;
			const uint32_t n = (uint32_t)(total - begin < PNG_CHUNK ? total - begin : PNG_CHUNK);
			const uint8_t *in = filtered + begin;
			BitWriter w;
			w.out = reinterpret_cast<uint32_t *>(scratch + (size_t)chunk * PNG_OUT_STRIDE);
			w.put(2u, 3); // BFINAL = 0, BTYPE = 01 (fixed Huffman), both LSB first
			uint32_t s1 = 0, s2 = 0; // Adler-32 partial sums of this chunk: sum x_i, sum (n - i) x_i (reduced below: n <= 8192 keeps them in 32 bits)
			uint32_t pos = 0;
			while (pos < n)
			{
				uint32_t length = 0, distance = 0;
				if (pos + 3 <= n)
				{
					const uint32_t key = (uint32_t)in[pos] | ((uint32_t)in[pos + 1] << 8) | ((uint32_t)in[pos + 2] << 16);
					const uint32_t h = (key * 2654435761u) >> 22;
					const uint32_t cand = hash[h]; // position + 1, 0 = empty
					hash[h] = (uint16_t)(pos + 1);
					if (cand)
					{
						const uint32_t at = cand - 1;
						const uint32_t limit = n - pos < 258u ? n - pos : 258u;
						uint32_t m = 0;
						while (m < limit && in[at + m] == in[pos + m])
							m++;
						if (m >= 3)
						{
							length = m;
							distance = pos - at;
						}
					}
				}
				if (length)
				{
					put_match(w, length, distance);
					for (uint32_t k = 0; k < length; k++)
					{
						const uint32_t x = in[pos + k];
						s1 += x;
						s2 += (n - (pos + k)) * x;
					}
					pos += length;
				}
				else
				{
					const uint32_t x = in[pos];
					put_literal(w, x);
					s1 += x;
					s2 += (n - pos) * x;
					pos++;
				}
				if ((pos & 1023u) < 258u) // keep the sums small (255 * 8192 * 1024 would overflow): a cheap reduction every ~1 KiB
					s2 %= 65521u;
			}
			w.put(0u, 7); // end of block: symbol 256 = seven zero bits
			// an empty stored block aligns the stream to a byte: BFINAL (the last chunk ends the stream), BTYPE = 00, pad, LEN = 0, NLEN = 0xffff
			w.put(chunk + 1 == chunks ? 1u : 0u, 3);
			w.bits = (w.bits + 7u) & ~7u;
			w.put(0u, 16);
			w.put(0xffffu, 16);
			sizes[chunk] = w.finish();
			adlerParts[chunk * 2 + 0] = s1 % 65521u;
			adlerParts[chunk * 2 + 1] = s2 % 65521u;
		}

		// ---- layout: offsets of the IDAT chunks, Adler-32 of the stream, total file size -----------------------------------------------
		// offsets[c] = file offset of IDAT chunk c; the first carries the 2-byte zlib header in front, the last the 4-byte Adler-32 behind
		__global__ void __launch_bounds__(1024) png_layout_kernel(const uint32_t *__restrict__ sizes, const uint32_t *__restrict__ adlerParts, uint32_t chunks, uint64_t total, uint64_t *__restrict__ offsets, uint32_t *__restrict__ adlerOut, uint64_t *__restrict__ fileSize)
		{
			__shared__ uint64_t carry;
			__shared__ uint64_t warpSums[32];
			if (threadIdx.x == 0)
				carry = PNG_HEADER;
			__syncthreads();
			for (uint32_t base = 0; base < chunks; base += 1024)
			{
				const uint32_t c = base + threadIdx.x;
				// bytes this IDAT chunk occupies in the file: length + type + data + crc
				uint64_t mine = 0;
				if (c < chunks)
					mine = 12ull + sizes[c] + (c == 0 ? 2u : 0u) + (c + 1 == chunks ? 4u : 0u);
				uint64_t incl = mine;
				for (int o = 1; o < 32; o <<= 1)
				{
					const uint64_t t = __shfl_up_sync(0xffffffffu, incl, o);
					if ((threadIdx.x & 31) >= (uint32_t)o)
						incl += t;
				}
				if ((threadIdx.x & 31) == 31)
					warpSums[threadIdx.x >> 5] = incl;
				__syncthreads();
				uint64_t before = carry;
				for (uint32_t wv = 0; wv < (threadIdx.x >> 5); wv++)
					before += warpSums[wv];
				if (c < chunks)
					offsets[c] = before + incl - mine;
				__syncthreads();
				if (threadIdx.x == 1023)
					carry = before + incl;
				__syncthreads();
			}
			if (threadIdx.x == 0)
			{
				// Adler-32 over the whole filtered stream from the per-chunk sums: a' = a + s1, b' = b + n * a + s2 (mod 65521)
				uint64_t a = 1, b = 0;
				for (uint32_t c = 0; c < chunks; c++)
				{
					const uint64_t begin = (uint64_t)c * PNG_CHUNK;
					const uint64_t n = total - begin < PNG_CHUNK ? total - begin : PNG_CHUNK;
					b = (b + n * a + adlerParts[c * 2 + 1]) % 65521u;
					a = (a + adlerParts[c * 2 + 0]) % 65521u;
				}
				*adlerOut = (uint32_t)((b << 16) | a);
				*fileSize = carry + 12; // + IEND
			}
		}

		__device__ __forceinline__ uint32_t crc_byte(uint32_t crc, uint32_t byte, const uint32_t *table) { return table[(crc ^ byte) & 255u] ^ (crc >> 8); }

		// ---- emit: one block per IDAT chunk ----------------------------------------------------------------------------------------------
		// CRC-32 is linear: the CRC of the chunk is the xor of the CRCs of its bytes each followed by the right number of zeros; simpler
		// here: thread 0 runs the table CRC over the <= 9.3 KiB of the chunk while the other threads copy (the copy hides it)
		__global__ void __launch_bounds__(128) png_emit_kernel(const uint8_t *__restrict__ scratch, const uint32_t *__restrict__ sizes, const uint64_t *__restrict__ offsets, const uint32_t *__restrict__ adler, uint32_t chunks, uint8_t *__restrict__ file)
		{
			__shared__ uint32_t table[256];
			for (uint32_t i = threadIdx.x; i < 256; i += 128)
				table[i] = c_crcTable[i];
			__syncthreads();
			const uint32_t c = blockIdx.x;
			if (c >= chunks)
				return;
			const uint32_t n = sizes[c];
			const uint32_t head = c == 0 ? 2u : 0u, tail = c + 1 == chunks ? 4u : 0u;
			const uint32_t dataLen = head + n + tail;
			uint8_t *out = file + offsets[c];
			const uint8_t *src = scratch + (size_t)c * PNG_OUT_STRIDE;
			uint8_t *data = out + 8;
			for (uint32_t i = threadIdx.x; i < n; i += 128)
				data[head + i] = src[i];
			if (threadIdx.x == 0)
			{
				out[0] = (uint8_t)(dataLen >> 24), out[1] = (uint8_t)(dataLen >> 16), out[2] = (uint8_t)(dataLen >> 8), out[3] = (uint8_t)dataLen;
				out[4] = 'I', out[5] = 'D', out[6] = 'A', out[7] = 'T';
				uint32_t crc = 0xffffffffu;
				crc = crc_byte(crc, 'I', table), crc = crc_byte(crc, 'D', table), crc = crc_byte(crc, 'A', table), crc = crc_byte(crc, 'T', table);
				if (head)
				{
					data[0] = 0x78, data[1] = 0x01; // zlib header: deflate, 32 KiB window, no preset dictionary, fastest (CMF * 256 + FLG divisible by 31)
					crc = crc_byte(crc, 0x78, table), crc = crc_byte(crc, 0x01, table);
				}
				for (uint32_t i = 0; i < n; i++)
					crc = crc_byte(crc, src[i], table);
				if (tail)
				{
					const uint32_t a = *adler;
					const uint8_t t4[4] = { (uint8_t)(a >> 24), (uint8_t)(a >> 16), (uint8_t)(a >> 8), (uint8_t)a };
					for (int k = 0; k < 4; k++)
					{
						data[head + n + k] = t4[k];
						crc = crc_byte(crc, t4[k], table);
					}
				}
				crc = ~crc;
				uint8_t *e = data + dataLen;
				e[0] = (uint8_t)(crc >> 24), e[1] = (uint8_t)(crc >> 16), e[2] = (uint8_t)(crc >> 8), e[3] = (uint8_t)crc;
			}
		}

		uint32_t host_crc_table[256];
		uint32_t host_crc(const uint8_t *p, size_t n)
		{
			uint32_t crc = 0xffffffffu;
			for (size_t i = 0; i < n; i++)
				crc = host_crc_table[(crc ^ p[i]) & 255u] ^ (crc >> 8);
			return ~crc;
		}

		cudaError_t upload_tables()
		{
			static std::once_flag once[64];
			static cudaError_t result[64];
			int device = 0;
			cudaError_t e = cudaGetDevice(&device);
			if (e != cudaSuccess)
				return e;
			std::call_once(once[device & 63],
				[&]()
				{
					for (uint32_t i = 0; i < 256; i++)
					{
						uint32_t c = i;
						for (int k = 0; k < 8; k++)
							c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
						host_crc_table[i] = c;
					}
					// length 3..258 -> symbol 257..285 + extra bits (RFC 1951 3.2.5), fixed code lengths 7 (256..279) / 8 (280..287)
					static const uint16_t base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
					static const uint8_t extraBits[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
					uint32_t table[259] = {};
					for (uint32_t len = 3; len <= 258; len++)
					{
						int s = 28;
						while (base[s] > len)
							s--;
						if (len == 258)
							s = 28;
						else if (s == 28)
							s = 27; // 227..257 belong to symbol 284
						const uint32_t symbol = 257 + s;
						uint32_t code, nbits;
						if (symbol < 280)
							code = symbol - 256, nbits = 7;
						else
							code = 0xc0 + (symbol - 280), nbits = 8;
						uint32_t rev = 0;
						for (uint32_t k = 0; k < nbits; k++)
							rev |= ((code >> k) & 1u) << (nbits - 1 - k);
						const uint32_t bits = rev | ((len - base[s]) << nbits);
						table[len] = (bits << 8) | (nbits + extraBits[s]);
					}
					cudaError_t r = cudaMemcpyToSymbol(c_lenCode, table, sizeof(table));
					if (r == cudaSuccess)
						r = cudaMemcpyToSymbol(c_crcTable, host_crc_table, sizeof(host_crc_table));
					result[device & 63] = r;
				});
			return result[device & 63];
		}

		void put_be32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)(v >> 24), p[1] = (uint8_t)(v >> 16), p[2] = (uint8_t)(v >> 8), p[3] = (uint8_t)v; }
	}

	uint64_t png_bound(uint32_t w, uint32_t h, uint32_t channels)
	{
		const uint64_t total = (uint64_t)h * ((uint64_t)w * channels + 1);
		const uint64_t chunks = (total + PNG_CHUNK - 1) / PNG_CHUNK;
		return PNG_HEADER + chunks * (PNG_OUT_STRIDE + 12ull) + 6 + 12;
	}

	// `pixels`: device memory on the current device, tightly packed rows.  `file`: device memory, png_bound() bytes.  Asynchronous on `s`;
	// *fileSizeDevice receives the size.  Scratch comes from (and returns to) the pool on `s`.
	int png_encode_device(const uint8_t *pixels, uint32_t w, uint32_t h, uint32_t channels, uint8_t *file, uint64_t *fileSizeDevice, cudaStream_t s)
	{
		if (!pixels || !file || !fileSizeDevice || w == 0 || h == 0 || channels < 1 || channels > 4)
			return fail(UPG_ERR_ARGUMENT, "bad PNG image %u x %u x %u", w, h, channels);
		cudaError_t e = upload_tables();
		if (e != cudaSuccess)
			return cuda_fail(e, "PNG tables");
		const uint32_t rowBytes = w * channels;
		const uint64_t total = (uint64_t)h * (rowBytes + 1ull);
		const uint64_t chunks64 = (total + PNG_CHUNK - 1) / PNG_CHUNK;
		if (chunks64 > 0x7fffffffull)
			return fail(UPG_ERR_ARGUMENT, "image too large for the PNG encoder");
		const uint32_t chunks = (uint32_t)chunks64;
		uint8_t *filtered = nullptr, *scratch = nullptr;
		uint32_t *sizes = nullptr, *adlerParts = nullptr, *adler = nullptr;
		uint64_t *offsets = nullptr;
		e = device_alloc(&filtered, total, s);
		if (e == cudaSuccess)
			e = device_alloc(&scratch, (size_t)chunks * PNG_OUT_STRIDE, s);
		if (e == cudaSuccess)
			e = device_alloc(&sizes, sizeof(uint32_t) * chunks, s);
		if (e == cudaSuccess)
			e = device_alloc(&adlerParts, sizeof(uint32_t) * 2 * chunks, s);
		if (e == cudaSuccess)
			e = device_alloc(&adler, sizeof(uint32_t), s);
		if (e == cudaSuccess)
			e = device_alloc(&offsets, sizeof(uint64_t) * chunks, s);
		if (e == cudaSuccess)
		{
			// signature + IHDR (13 bytes: width, height, bit depth 8, colour type, compression 0, filter 0, no interlace) + IEND are host work
			uint8_t head[PNG_HEADER] = { 0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a, 0, 0, 0, 13, 'I', 'H', 'D', 'R' };
			put_be32(head + 16, w);
			put_be32(head + 20, h);
			head[24] = 8;
			head[25] = channels == 1 ? 0 : channels == 2 ? 4 : channels == 3 ? 2 : 6;
			head[26] = head[27] = head[28] = 0;
			put_be32(head + 29, host_crc(head + 12, 17));
			e = cudaMemcpyAsync(file, head, PNG_HEADER, cudaMemcpyHostToDevice, s); // pageable source: staged before the call returns
		}
		if (e == cudaSuccess)
		{
			png_filter_kernel<<<h, 256, 0, s>>>(pixels, rowBytes, channels, h, filtered);
			constexpr size_t hashBytes = sizeof(uint16_t) * PNG_HASH * PNG_DEFLATE_THREADS;
			e = cudaFuncSetAttribute(png_deflate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hashBytes);
			if (e == cudaSuccess)
				png_deflate_kernel<<<(chunks + PNG_DEFLATE_THREADS - 1) / PNG_DEFLATE_THREADS, PNG_DEFLATE_THREADS, hashBytes, s>>>(filtered, total, chunks, scratch, sizes, adlerParts);
			png_layout_kernel<<<1, 1024, 0, s>>>(sizes, adlerParts, chunks, total, offsets, adler, fileSizeDevice);
			png_emit_kernel<<<chunks, 128, 0, s>>>(scratch, sizes, offsets, adler, chunks, file);
			e = cudaGetLastError();
		}
		void *all[] = { filtered, scratch, sizes, adlerParts, adler, offsets };
		for (void *p : all)
			if (p)
				device_free(p, s);
		if (e != cudaSuccess)
			return cuda_fail(e, "PNG encoder");
		return UPG_OK;
	}

	// the 12 bytes behind the last IDAT chunk
	static const uint8_t pngIend[12] = { 0, 0, 0, 0, 'I', 'E', 'N', 'D', 0xae, 0x42, 0x60, 0x82 };
}

using namespace upg;

extern "C"
{
	uint64_t upg_png_bound(uint32_t width, uint32_t height, uint32_t channels)
	{
		return png_bound(width, height, channels);
	}

	int upg_png_encode(int device, const uint8_t *pixels, uint32_t width, uint32_t height, uint32_t channels, uint8_t *out, uint64_t capacity, uint64_t *size)
	{
		if (!pixels || !out || !size)
			return fail(UPG_ERR_ARGUMENT, "null argument");
		*size = 0;
		if (width == 0 || height == 0 || channels < 1 || channels > 4)
			return fail(UPG_ERR_ARGUMENT, "bad PNG image %u x %u x %u", width, height, channels);
		cudaError_t e = cudaSetDevice(device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		int rc = cs.open();
		if (rc)
			return rc;
		cudaStream_t s = cs.compute;
		const uint64_t bytes = (uint64_t)width * height * channels, bound = png_bound(width, height, channels);
		uint8_t *dpix = nullptr, *dfile = nullptr;
		uint64_t *dsize = nullptr;
		// pixels in host memory are uploaded (pageable memory through the staged copy the bake uses); device memory is used in place
		cudaPointerAttributes attr;
		const bool onDevice = cudaPointerGetAttributes(&attr, pixels) == cudaSuccess && (attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
		if (!onDevice)
			cudaGetLastError();
		void *staging = nullptr;
		size_t stagingCapacity = 0;
		if (!onDevice)
		{
			e = device_alloc(&dpix, bytes, s);
			if (e == cudaSuccess)
			{
				const HostArray a[1] = { { dpix, pixels, (size_t)bytes } };
				rc = upload_host_arrays(a, 1, s, &staging, &stagingCapacity);
			}
		}
		if (e == cudaSuccess)
			e = device_alloc(&dfile, bound, s);
		if (e == cudaSuccess)
			e = device_alloc(&dsize, sizeof(uint64_t), s);
		if (e != cudaSuccess && rc == UPG_OK)
			rc = cuda_fail(e, "device_alloc(PNG)");
		if (rc == UPG_OK)
			rc = png_encode_device(onDevice ? pixels : dpix, width, height, channels, dfile, dsize, s);
		const bool trace = trace_enabled();
		const double t0 = trace ? now_ms() : 0;
		uint64_t fileSize = 0;
		if (rc == UPG_OK)
		{
			e = cudaMemcpyAsync(&fileSize, dsize, sizeof(fileSize), cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess)
				e = stream_sync(s);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "PNG encoder");
		}
		const double t1 = trace ? now_ms() : 0;
		if (trace)
			std::fprintf(stderr, "[upg] png %u x %u x %u: upload + filter + deflate + layout + emit %.2f ms\n", width, height, channels, t1 - t0);
		if (rc == UPG_OK && fileSize > capacity)
			rc = fail(UPG_ERR_ARGUMENT, "PNG output buffer too small: %llu bytes needed, %llu given (upg_png_bound)", (unsigned long long)fileSize, (unsigned long long)capacity);
		if (rc == UPG_OK)
		{
			e = cudaMemcpyAsync(out, dfile, fileSize - 12, cudaMemcpyDefault, s);
			if (e == cudaSuccess)
				e = stream_sync(s);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "PNG download");
			else
			{
				std::memcpy(out + fileSize - 12, pngIend, 12);
				*size = fileSize;
			}
			if (trace)
				std::fprintf(stderr, "[upg] png %u x %u x %u: %.2f MB file, download %.2f ms (after the encode kernels were waited for)\n", width, height, channels, fileSize * 1e-6, now_ms() - t1);
		}
		if (staging)
		{
			stream_sync(s);
			pinned_block_release(staging, stagingCapacity);
		}
		void *all[] = { dpix, dfile, dsize };
		for (void *p : all)
			if (p)
				device_free(p, s);
		stream_sync(s);
		cs.close();
		return rc;
	}

	// K5..K8 + K10: the complete texture export of one chunk.  The three images never exist uncompressed in host memory: bake and texture
	// tail write them to HBM, the encoder compresses them there, the files come down the link (about half the bytes of the raw images).
	int upg_bake_export_png(upg_planet *planet, int water, const float *positions, const float *normals, const float *uvs, uint32_t vertex_count, const uint32_t *indices, uint32_t index_count, uint32_t width, uint32_t height, uint8_t *const out3[3], const uint64_t capacity3[3], uint64_t size3[3])
	{
		Planet *p = reinterpret_cast<Planet *>(planet);
		if (!p || !positions || !normals || !out3 || !capacity3 || !size3 || !out3[0] || !out3[1] || !out3[2])
			return fail(UPG_ERR_ARGUMENT, "null argument");
		size3[0] = size3[1] = size3[2] = 0;
		int rc = check_bake_mesh(uvs, vertex_count, indices, index_count, width, height);
		if (rc)
			return rc;
		cudaError_t e = cudaSetDevice(p->device);
		if (e != cudaSuccess)
			return cuda_fail(e, "cudaSetDevice");
		CallStreams cs;
		rc = cs.open();
		if (rc)
			return rc;
		cudaStream_t s = cs.compute;
		const uint32_t channels[3] = { water ? 4u : 3u, 3u, 3u };
		const uint64_t texels = (uint64_t)width * height;
		float *dpos = nullptr, *dnrm = nullptr, *duv = nullptr;
		uint32_t *didx = nullptr;
		uint8_t *image[3] = { nullptr, nullptr, nullptr }, *file[3] = { nullptr, nullptr, nullptr };
		uint64_t *dsize = nullptr;
		e = device_alloc(&dpos, sizeof(float) * 3 * (size_t)vertex_count, s);
		if (e == cudaSuccess)
			e = device_alloc(&dnrm, sizeof(float) * 3 * (size_t)vertex_count, s);
		if (e == cudaSuccess)
			e = device_alloc(&duv, sizeof(float) * 2 * (size_t)vertex_count, s);
		if (e == cudaSuccess)
			e = device_alloc(&didx, sizeof(uint32_t) * (size_t)index_count, s);
		for (int k = 0; k < 3 && e == cudaSuccess; k++)
		{
			e = device_alloc(&image[k], texels * channels[k], s);
			if (e == cudaSuccess)
				e = device_alloc(&file[k], png_bound(width, height, channels[k]), s);
		}
		if (e == cudaSuccess)
			e = device_alloc(&dsize, sizeof(uint64_t) * 3, s);
		if (e != cudaSuccess)
			rc = cuda_fail(e, "device_alloc(export)");
		void *staging = nullptr;
		size_t stagingCapacity = 0;
		if (rc == UPG_OK)
		{
			const HostArray arrays[4] = { { duv, uvs, sizeof(float) * 2 * (size_t)vertex_count }, { didx, indices, sizeof(uint32_t) * (size_t)index_count }, { dpos, positions, sizeof(float) * 3 * (size_t)vertex_count }, { dnrm, normals, sizeof(float) * 3 * (size_t)vertex_count } };
			rc = upload_host_arrays(arrays, 4, s, &staging, &stagingCapacity);
		}
		// albedo, pbr and normal map in HBM (bake_core's outputs may be device memory)
		if (rc == UPG_OK)
			rc = bake_core(planet, water, dpos, dnrm, duv, vertex_count, didx, index_count, width, height, 0, height, image[0], nullptr, nullptr, image[1], image[2], 1.f, s);
		for (int k = 0; k < 3 && rc == UPG_OK; k++)
			rc = png_encode_device(image[k], width, height, channels[k], file[k], dsize + k, s);
		uint64_t sizes[3] = { 0, 0, 0 };
		if (rc == UPG_OK)
		{
			e = cudaMemcpyAsync(sizes, dsize, sizeof(sizes), cudaMemcpyDeviceToHost, s);
			if (e == cudaSuccess)
				e = stream_sync(s);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "PNG encoder");
		}
		for (int k = 0; k < 3 && rc == UPG_OK; k++)
			if (sizes[k] > capacity3[k])
				rc = fail(UPG_ERR_ARGUMENT, "PNG output buffer %d too small: %llu bytes needed, %llu given (upg_png_bound)", k, (unsigned long long)sizes[k], (unsigned long long)capacity3[k]);
		for (int k = 0; k < 3 && rc == UPG_OK; k++)
		{
			e = cudaMemcpyAsync(out3[k], file[k], sizes[k] - 12, cudaMemcpyDefault, s);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "PNG download");
		}
		if (rc == UPG_OK)
		{
			e = stream_sync(s);
			if (e != cudaSuccess)
				rc = cuda_fail(e, "PNG download");
		}
		for (int k = 0; k < 3 && rc == UPG_OK; k++)
		{
			std::memcpy(out3[k] + sizes[k] - 12, pngIend, 12);
			size3[k] = sizes[k];
		}
		if (staging)
		{
			stream_sync(s);
			pinned_block_release(staging, stagingCapacity);
		}
		void *all[] = { dpos, dnrm, duv, didx, image[0], image[1], image[2], file[0], file[1], file[2], dsize };
		for (void *q : all)
			if (q)
				device_free(q, s);
		stream_sync(s);
		cs.close();
		return rc;
	}
}
