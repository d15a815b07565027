// ngp_k_gibbs.cu — ONE instantiation of ngp::gibbs_kernel per translation unit: compile with -DNGP_KB={16,32,64} -DNGP_KV={0..5}
// (block size; variant: 0 plain, 1 instrumented, 2 timing experiments, 3 per-marker "literal" sweep, 4 tuple sweep, 5 all ranks of a sharded chain on one device as one grid).  See ngp_kernels.h.
#include "ngp_sweep.cuh"
#include "ngp_kernels.h"

#ifndef NGP_KB
#error "compile with -DNGP_KB=16|32|64 -DNGP_KV=0..4"
#endif
#define NGP_CAT3_(a, b, c) a##b##_##c
#define NGP_CAT3(a, b, c) NGP_CAT3_(a, b, c)

extern "C" const void* NGP_CAT3(ngp_kptr_gibbs_, NGP_KB, NGP_KV)(void)
{
#if NGP_KV == 5
    return (const void*)ngp::gibbs_group_kernel<NGP_KB, true>;
#elif NGP_KV == 6
    return (const void*)ngp::gibbs_kernel<NGP_KB, false, false, false, false, true>;
#elif NGP_KV == 7
    return (const void*)ngp::gibbs_kernel<NGP_KB, false, false, false, false, false, true>;
#elif NGP_KV == 8
    return (const void*)ngp::gibbs_group_kernel<NGP_KB, false>;
#elif NGP_KV == 9
    return (const void*)ngp::gibbs_kernel<NGP_KB, false, false, false, false, false, false, true>;
#elif NGP_KV == 10
    return (const void*)ngp::gibbs_kernel<NGP_KB, false, false, false, false, true, false, true>;
#elif NGP_KV == 11
    return (const void*)ngp::gibbs_kernel<NGP_KB, false, false, false, false, true, false, false, true>;
#else
    return (const void*)ngp::gibbs_kernel<NGP_KB, NGP_KV == NGP_KV_PROF, NGP_KV == NGP_KV_DBG, NGP_KV == NGP_KV_LIT, NGP_KV == NGP_KV_TUP>;
#endif
}
