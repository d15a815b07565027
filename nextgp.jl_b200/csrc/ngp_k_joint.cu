// ngp_k_joint.cu — ONE instantiation of ngp::joint_kernel<k> (per-locus tuple sampler) per translation unit: compile with -DNGP_JK=2..8.
#include "ngp_joint.cuh"
#include "ngp_kernels.h"

#ifndef NGP_JK
#error "compile with -DNGP_JK=2..8"
#endif
#define NGP_CAT2_(a, b) a##b
#define NGP_CAT2(a, b) NGP_CAT2_(a, b)

extern "C" const void* NGP_CAT2(ngp_kptr_joint_, NGP_JK)(void) { return (const void*)ngp::joint_kernel<NGP_JK>; }
