// gbtrf_blocked.cu -- wide-band partial-pivot LU (panel + trailing update).  See gbtrf.cu for the contract.
#include "common.cuh"

int bmb_gbtrf_blocked(bmb200_ctx *h, i64 m, i64 n, i64 kl, i64 ku, double *dAB, i64 ldab, i64 *d_ipiv)
{
    (void)m; (void)n; (void)kl; (void)ku; (void)dAB; (void)ldab; (void)d_ipiv;
    snprintf(h->err, sizeof(h->err), "dgbtrf: band (%lld,%lld) exceeds the shared-memory window kernel; "
             "blocked path not built yet", (long long)kl, (long long)ku);
    return BMB200_ERR_CUDA;
}
