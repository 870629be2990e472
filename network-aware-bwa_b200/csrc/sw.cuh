// sw.cuh -- K5 placeholder, replaced below in this round.
#pragma once
#include <cuda_runtime.h>
#include "../../include/bwa_gpu.h"
namespace bwagpu {
static int sw_batch(cudaStream_t, const uint8_t *, int64_t, int, const bwa_gpu_sw_job_t *, bwa_gpu_sw_res_t *,
                    int (*fail)(const char *, ...))
{
	return fail("bwa_gpu_mate_sw: kernel not built yet");
}
}
