/* mdl_b200_selftest.h -- test infrastructure, NOT part of the reference-facing surface: self-tests, descriptor probes
 * and microbenchmarks of the tcgen05 / TMEM conventions (csrc/umma.cuh) the fused kernels use.  Built into its own
 * library, libmdl_b200_selftest.so (csrc/umma_selftest.cu), so that the product library carries none of it. */
#ifndef MDL_B200_SELFTEST_H
#define MDL_B200_SELFTEST_H
#include "mdl_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* ---- tensor-core self-test: D[128,N] = A[128,K] . B[N,K]^T through the same
 * tcgen05/TMEM conventions (umma.cuh) the fused kernels use.  split=0: plain
 * TF32 (operands truncated by the hardware); split=1: 3xTF32 (fp32-faithful).
 * No reference counterpart: test infrastructure for the kernels above. ---- */
MDL_API int mdl_selftest_umma(const float* A, const float* B, float* D, int32_t N, int32_t K,
                              int32_t split, void* stream);
/* descriptor-field probe used while bringing up umma.cuh: same staging layout, caller-chosen LBO/SBO */
MDL_API int mdl_selftest_umma_ex(const float* A, const float* B, float* D, int32_t N, int32_t K,
                                 int32_t split, int32_t lbo_a, int32_t sbo_a, int32_t lbo_b, int32_t sbo_b,
                                 void* stream);
/* same product with A staged in tensor memory (tcgen05.st) and B in shared memory */
MDL_API int mdl_selftest_umma_ts(const float* A, const float* B, float* D, int32_t N, int32_t K,
                                 int32_t split, void* stream);
/* microbenchmark (development aid): cycles of `nstores` tcgen05.st of `width` columns per warp, with `mma_count`
 * 128x128x8 MMAs issued concurrently; out = 18 x int64 (per-warp cycles, [16] MMA issue, [17] MMA complete) */
MDL_API int mdl_selftest_tmem_st_bench(long long* out, int32_t nwarps, int32_t nstores, int32_t width,
                                       int32_t mma_count, int32_t wait_each, void* stream);
/* layout probe: raw shared-memory images of both operand tiles + descriptor fields, nmma K=8 MMAs */
MDL_API int mdl_selftest_umma_probe(const float* rawA, int32_t a_floats, const float* rawB, int32_t b_floats,
                                    float* D, int32_t N, int32_t lbo_a, int32_t sbo_a, int32_t lbo_b,
                                    int32_t sbo_b, int32_t a_mn, int32_t b_mn, int32_t nmma, int32_t step_a,
                                    int32_t step_b, int32_t layout_a, int32_t layout_b, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDL_B200_SELFTEST_H */
