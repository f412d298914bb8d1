// sharded.cu -- multi-GPU row-sharded gbmv with an x halo over NVLink peer memory.
#include "common.cuh"

extern "C" int bmb200_halo_create(bmb200_handle_t h, int64_t max_halo, void *ipc_handle_out)
{
    (void)max_halo; (void)ipc_handle_out;
    if (!h) return -1;
    snprintf(h->err, sizeof(h->err), "halo exchange not built yet");
    return BMB200_ERR_CUDA;
}
extern "C" int bmb200_halo_connect(bmb200_handle_t h, int rank, int nranks, const void *ipc_left, const void *ipc_right)
{
    (void)rank; (void)nranks; (void)ipc_left; (void)ipc_right;
    if (!h) return -1;
    return BMB200_ERR_CUDA;
}
extern "C" int bmb200_halo_destroy(bmb200_handle_t h)
{
    if (!h) return -1;
    return 0;
}
extern "C" int bmb200_dgbmv_sharded(bmb200_handle_t h, int64_t n_global, int64_t c0, int64_t c1, int64_t kl,
                                    int64_t ku, double alpha, const double *dA_local, int64_t lda,
                                    const double *dx_local, double beta, double *dy_local)
{
    (void)n_global; (void)c0; (void)c1; (void)kl; (void)ku; (void)alpha; (void)dA_local; (void)lda; (void)dx_local;
    (void)beta; (void)dy_local;
    if (!h) return -1;
    return BMB200_ERR_CUDA;
}
