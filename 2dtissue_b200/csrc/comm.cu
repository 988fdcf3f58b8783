// comm.cu — multi-GPU slabs: NCCL halo + migration exchange.  Filled in below.
#include "t2d_internal.h"

extern "C" {
int t2d_comm_unique_id(uint8_t id[T2D_UNIQUE_ID_BYTES])
{
    (void)id;
    return -1;
}
int t2d_comm_init(t2d_ctx* ctx, int rank, int world, const uint8_t id[T2D_UNIQUE_ID_BYTES], const double* cuts)
{
    (void)ctx; (void)rank; (void)world; (void)id; (void)cuts;
    return -1;
}
int t2d_comm_destroy(t2d_ctx* ctx)
{
    (void)ctx;
    return -1;
}
}
