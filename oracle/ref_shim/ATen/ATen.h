// ORACLE build shim (test infrastructure): the reference kernel header
// MOTR/models/ops/src/cuda/ms_deform_im2col_cuda.cuh includes <ATen/ATen.h>, <ATen/cuda/CUDAContext.h>
// and <THC/THCAtomics.cuh> but its forward kernel uses nothing from them; the backward kernels only
// need atomicAdd(float/double), native on sm_100. Empty stand-ins keep libtorch out of oracle/_ref.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
