// Host side of the TMA-fed marching kernels: tensor maps over plane-SoA vectors.
// cuTensorMapEncodeTiled is fetched through the runtime (cudaGetDriverEntryPoint):
// the library does not link libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string>

typedef CUresult (*ksfd_tmap_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                        const cuuint64_t *, const cuuint64_t *,
                                        const cuuint32_t *, const cuuint32_t *,
                                        CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline ksfd_tmap_encode_fn ksfd_tmap_encoder()
{
    static ksfd_tmap_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
                cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (ksfd_tmap_encode_fn)p;
        else
            cudaGetLastError();
    }
    return fn;
}

// The three box shapes of a tile (tma_march.cuh): centre (TX,TY), y strip (TX,2), x strip
// (2,TY); in 2-D (n1 == 1) the boxes are one row high and the y strip is unused.
// base: first double of the buffer; nfp = fields per plane * planes of the buffer.
// Returns an empty string or an error message.
static inline std::string ksfd_make_tmaps(CUtensorMap out[3], const double *base, long long n0,
                                          long long n1, long long nfp, int nc, int TX, int TY)
{
    ksfd_tmap_encode_fn enc = ksfd_tmap_encoder();
    if (!enc) return "cuTensorMapEncodeTiled is not available";
    const cuuint64_t dims[3] = {(cuuint64_t)n0, (cuuint64_t)n1, (cuuint64_t)nfp};
    const cuuint64_t strides[2] = {(cuuint64_t)n0 * 8, (cuuint64_t)n0 * n1 * 8};
    const cuuint32_t es[3] = {1, 1, 1};
    const int two_d = n1 == 1;
    const cuuint32_t bx[3] = {(cuuint32_t)TX, (cuuint32_t)TX, 2};
    const cuuint32_t by[3] = {(cuuint32_t)(two_d ? 1 : TY), (cuuint32_t)(two_d ? 1 : 2),
                              (cuuint32_t)(two_d ? 1 : TY)};
    for (int s = 0; s < 3; ++s) {
        const cuuint32_t box[3] = {bx[s], by[s], (cuuint32_t)nc};
        const CUresult r = enc(&out[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, (void *)base, dims,
                               strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE,
                               s == 2 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                      : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS)
            return "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r);
    }
    return std::string();
}
