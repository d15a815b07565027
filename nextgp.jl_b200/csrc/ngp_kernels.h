// ngp_kernels.h — the sweep kernels are compiled one instantiation per translation unit (ngp_k_gibbs.cu with -DNGP_KB / -DNGP_KV,
// ngp_k_joint.cu with -DNGP_JK) so that the library builds in parallel and a change recompiles only what it touches; each unit
// exports the address of its kernel's host stub through one C function, which ngp_api.cu looks up in the tables below.
#pragma once

// variants of ngp::gibbs_kernel<B, PROF, DBG, LIT, TUP> (ngp_sweep.cuh)
enum { NGP_KV_PLAIN = 0, NGP_KV_PROF = 1, NGP_KV_DBG = 2, NGP_KV_LIT = 3, NGP_KV_TUP = 4, NGP_KV_GROUP = 5, NGP_KV_BIGR = 6, NGP_KV_R = 7, NGP_KV_GROUPB = 8, NGP_KV_SHARD = 9, NGP_KV_SHARD_BIGR = 10, NGP_KV_BIGR1 = 11, NGP_KV_COUNT = 12 };   // 5: ngp::gibbs_group_kernel<B>; 6: plain sweep with up to 2048 rows per CTA for blocks of 32 / 64; 7: blocked sweep of BayesR sets; 8: shard group on the blocked sweep; 9 / 10: blocked sweep (plain / up to 2048 rows) of one rank of a row-sharded chain over several GPUs; 11: as 6 on a refetch ring of ONE tile whose dots the 8 dot warps share

#define NGP_KDECL(B, V) extern "C" const void* ngp_kptr_gibbs_##B##_##V(void);
#define NGP_KDECL_B(B) NGP_KDECL(B, 0) NGP_KDECL(B, 1) NGP_KDECL(B, 2) NGP_KDECL(B, 3) NGP_KDECL(B, 4) NGP_KDECL(B, 5) NGP_KDECL(B, 6) NGP_KDECL(B, 7) NGP_KDECL(B, 8) NGP_KDECL(B, 9) NGP_KDECL(B, 10) NGP_KDECL(B, 11)
NGP_KDECL_B(16) NGP_KDECL_B(32) NGP_KDECL_B(64)
#undef NGP_KDECL_B
#undef NGP_KDECL
#define NGP_JDECL(K) extern "C" const void* ngp_kptr_joint_##K(void);
NGP_JDECL(2) NGP_JDECL(3) NGP_JDECL(4) NGP_JDECL(5) NGP_JDECL(6) NGP_JDECL(7) NGP_JDECL(8)
#undef NGP_JDECL

static inline const void* ngp_gibbs_kernel(int B, int variant)
{
#define NGP_KROW(B) {ngp_kptr_gibbs_##B##_0, ngp_kptr_gibbs_##B##_1, ngp_kptr_gibbs_##B##_2, ngp_kptr_gibbs_##B##_3, ngp_kptr_gibbs_##B##_4, ngp_kptr_gibbs_##B##_5, ngp_kptr_gibbs_##B##_6, ngp_kptr_gibbs_##B##_7, ngp_kptr_gibbs_##B##_8, ngp_kptr_gibbs_##B##_9, ngp_kptr_gibbs_##B##_10, ngp_kptr_gibbs_##B##_11}
    typedef const void* (*fn_t)(void);
    static const fn_t tab[3][NGP_KV_COUNT] = {NGP_KROW(16), NGP_KROW(32), NGP_KROW(64)};
#undef NGP_KROW
    return tab[B == 64 ? 2 : B == 32 ? 1 : 0][variant]();
}

static inline const void* ngp_joint_kernel(int k)
{
    typedef const void* (*fn_t)(void);
    static const fn_t tab[7] = {ngp_kptr_joint_2, ngp_kptr_joint_3, ngp_kptr_joint_4, ngp_kptr_joint_5, ngp_kptr_joint_6, ngp_kptr_joint_7, ngp_kptr_joint_8};
    return tab[(k < 2 ? 2 : k > 8 ? 8 : k) - 2]();
}
