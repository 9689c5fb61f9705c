// quickrank_b200 — multi-GPU plumbing.  NCCL is resolved with dlopen at run time so that the
// single-GPU path has no link-time dependency on it.
#include "qr_comm.cuh"

#include <dlfcn.h>

namespace qr {

struct Comm {
  int rank = 0, world = 1;
};

void comm_destroy(Comm *c) { delete c; }

int comm_allreduce_sum_f64(Comm *, double *, size_t, cudaStream_t) {
  set_error("multi-GPU support is not built yet");
  return QR_ECOMM;
}
int comm_allreduce_max_u64(Comm *, unsigned long long *, size_t, cudaStream_t) {
  set_error("multi-GPU support is not built yet");
  return QR_ECOMM;
}
int comm_reduce_tasks(qr_ctx *, uint32_t, bool) {
  set_error("multi-GPU support is not built yet");
  return QR_ECOMM;
}
int comm_leaf_values(qr_ctx *, uint32_t) {
  set_error("multi-GPU support is not built yet");
  return QR_ECOMM;
}

}  // namespace qr

extern "C" {

int qr_comm_unique_id(unsigned char id[QR_COMM_ID_BYTES]) {
  (void) id;
  qr::set_error("multi-GPU support is not built yet");
  return QR_ECOMM;
}

int qr_ctx_comm_init(qr_ctx *ctx, const unsigned char id[QR_COMM_ID_BYTES], int rank, int world) {
  (void) ctx; (void) id; (void) rank; (void) world;
  qr::set_error("multi-GPU support is not built yet");
  return QR_ECOMM;
}

}
