// Multi-GPU plumbing: one process per GPU, one partition per process; halo values move by
// NCCL send/recv over NVLink and residual scalars by all-reduce.  NCCL is dlopen'ed
// (libnccl.so.2, the copy torch already loaded when the host is Python) so that the library
// loads on boxes without NCCL and single-GPU use never touches it.
#include "state.h"

extern "C" {

int cfdl_comm_unique_id(uint8_t id[128]) {
  (void)id;
  return cfdl::fail(CFDL_ERR_UNSUPPORTED, "cfdl_comm_unique_id: multi-GPU path not built yet");
}
int cfdl_comm_init(cfdl_handle h, const uint8_t id[128], int32_t rank, int32_t nranks) {
  (void)h; (void)id; (void)rank; (void)nranks;
  return cfdl::fail(CFDL_ERR_UNSUPPORTED, "cfdl_comm_init: multi-GPU path not built yet");
}
int cfdl_set_interfaces(cfdl_handle h, int32_t nnbr, const int32_t* nbr_rank, const int32_t* send_ptr, const int32_t* send_cells,
                        const int32_t* recv_ptr, const int32_t* recv_halos, int64_t ne_global, int32_t owns_ref_cell) {
  (void)h; (void)nnbr; (void)nbr_rank; (void)send_ptr; (void)send_cells; (void)recv_ptr; (void)recv_halos; (void)ne_global; (void)owns_ref_cell;
  return cfdl::fail(CFDL_ERR_UNSUPPORTED, "cfdl_set_interfaces: multi-GPU path not built yet");
}

}  // extern "C"
