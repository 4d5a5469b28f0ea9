// upg_peak.cu -- measurement helper: the FP32 FMA issue ceiling of the device the library runs on.
// MEASURED_PEAKS.json holds HBM and bf16-tensor peaks only; K1 / K6 are FP32 SIMT kernels, so bench.py measures
// their roofline denominator live with this kernel (dependent-free FFMA chains, 8 accumulators per thread).
#include <cuda_runtime.h>

#include "../../include/upg.h"
#include "upg_internal.h"

namespace upg
{
	__global__ void __launch_bounds__(256) fma_peak_kernel(float *out, float a, float b, int iters)
	{
		float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
		for (int i = 0; i < iters; i++)
		{
#pragma unroll
			for (int u = 0; u < 16; u++)
			{
				x0 = __fmaf_rn(x0, a, b);
				x1 = __fmaf_rn(x1, a, b);
				x2 = __fmaf_rn(x2, a, b);
				x3 = __fmaf_rn(x3, a, b);
				x4 = __fmaf_rn(x4, a, b);
				x5 = __fmaf_rn(x5, a, b);
				x6 = __fmaf_rn(x6, a, b);
				x7 = __fmaf_rn(x7, a, b);
			}
		}
		out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
	}
}

extern "C" int upg_measure_fp32_peak(int device, float *tflops_out)
{
	using namespace upg;
	if (!tflops_out)
		return fail(UPG_ERR_ARGUMENT, "null argument");
	cudaError_t e = cudaSetDevice(device);
	if (e != cudaSuccess)
		return cuda_fail(e, "cudaSetDevice");
	cudaDeviceProp prop;
	e = cudaGetDeviceProperties(&prop, device);
	if (e != cudaSuccess)
		return cuda_fail(e, "cudaGetDeviceProperties");
	const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
	float *out = nullptr;
	e = cudaMalloc(&out, sizeof(float) * blocks * threads);
	if (e != cudaSuccess)
		return cuda_fail(e, "cudaMalloc");
	cudaEvent_t t0, t1;
	cudaEventCreate(&t0);
	cudaEventCreate(&t1);
	float best = 0;
	for (int rep = 0; rep < 6; rep++)
	{
		cudaEventRecord(t0);
		fma_peak_kernel<<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
		cudaEventRecord(t1);
		e = cudaEventSynchronize(t1);
		if (e != cudaSuccess)
			break;
		float ms = 0;
		cudaEventElapsedTime(&ms, t0, t1);
		const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
		const float tf = (float)(flops / (ms * 1e-3) / 1e12);
		if (rep > 0 && tf > best)
			best = tf;
	}
	cudaEventDestroy(t0);
	cudaEventDestroy(t1);
	cudaFree(out);
	if (e != cudaSuccess)
		return cuda_fail(e, "fma_peak_kernel");
	*tflops_out = best;
	return UPG_OK;
}
