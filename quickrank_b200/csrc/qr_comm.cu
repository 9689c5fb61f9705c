// quickrank_b200 — multi-GPU plumbing: one process per GPU, documents sharded by query, NCCL over
// NVLink for the only exchange steps the path has:
//   * once at start: the global list of distinct feature values (thresholds must be identical on
//     every rank) and the per-bin document counts of the whole dataset;
//   * per tree: the maximum |pseudo-response| (common fixed-point scale);
//   * per growth round: the freshly built per-bin histograms (int64 sums + uint32 counts) of the
//     round's nodes and their squares sums, so that every rank scans the same totals and makes the
//     same split decisions without any broadcast;
//   * per tree: per-leaf (sum lambda, sum weight); per evaluation: the sum of per-query NDCG.
// Histogram sums are integers, so the all-reduced values — and therefore every split — do not
// depend on the number of ranks or on NCCL's reduction order.
//
// NCCL is resolved with dlopen at run time (libnccl.so.2; inside a PyTorch process this is the
// copy torch already loaded), so the single-GPU path has no link-time dependency on it.
#include "qr_comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>

#include <cfloat>

#include "qr_task.cuh"

namespace qr {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;

static int load_nccl() {
  if (g_nccl.handle) return QR_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) { set_error("cannot load NCCL (libnccl.so.2): %s", dlerror()); return QR_ECOMM; }
#define QR_SYM(field, name)                                                       \
  *(void **) (&g_nccl.field) = dlsym(h, name);                                    \
  if (!g_nccl.field) { set_error("NCCL symbol %s not found", name); return QR_ECOMM; }
  QR_SYM(GetUniqueId, "ncclGetUniqueId");
  QR_SYM(CommInitRank, "ncclCommInitRank");
  QR_SYM(CommDestroy, "ncclCommDestroy");
  QR_SYM(AllReduce, "ncclAllReduce");
  QR_SYM(AllGather, "ncclAllGather");
  QR_SYM(GroupStart, "ncclGroupStart");
  QR_SYM(GroupEnd, "ncclGroupEnd");
  QR_SYM(GetErrorString, "ncclGetErrorString");
#undef QR_SYM
  g_nccl.handle = h;
  return QR_OK;
}

#define QR_NCCL(expr)                                                                        \
  do {                                                                                       \
    ncclResult_t _r = (expr);                                                                \
    if (_r != ncclSuccess) {                                                                 \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r)); \
      return QR_ECOMM;                                                                       \
    }                                                                                        \
  } while (0)

struct Comm {
  int rank = 0, world = 1;
  ncclComm_t nccl = nullptr;
  long long *d_sq_limbs = nullptr;   // [max_tasks][3] 43-bit limbs of the squares sums
  double *d_scratch = nullptr;       // small host<->device staging
};

int comm_rank(const Comm *c) { return c ? c->rank : 0; }
int comm_world(const Comm *c) { return c ? c->world : 1; }

int comm_create(const unsigned char id[QR_COMM_ID_BYTES], int rank, int world, Comm **out) {
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) { set_error("bad rank %d / world %d", rank, world); return QR_EINVAL; }
  QR_TRY(load_nccl());
  Comm *c = new Comm();
  c->rank = rank;
  c->world = world;
  ncclUniqueId uid;
  static_assert(sizeof(uid.internal) == QR_COMM_ID_BYTES, "NCCL unique id size");
  memcpy(uid.internal, id, QR_COMM_ID_BYTES);
  ncclResult_t r = g_nccl.CommInitRank(&c->nccl, world, uid, rank);
  if (r != ncclSuccess) {
    set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
    delete c;
    return QR_ECOMM;
  }
  *out = c;
  return QR_OK;
}

void comm_destroy(Comm *c) {
  if (!c) return;
  if (c->d_sq_limbs) cudaFree(c->d_sq_limbs);
  if (c->d_scratch) cudaFree(c->d_scratch);
  if (c->nccl) g_nccl.CommDestroy(c->nccl);
  delete c;
}

int comm_allreduce_sum_f64(Comm *c, double *buf, size_t count, cudaStream_t st) {
  QR_NCCL(g_nccl.AllReduce(buf, buf, count, ncclFloat64, ncclSum, c->nccl, st));
  return QR_OK;
}
int comm_allreduce_max_u64(Comm *c, unsigned long long *buf, size_t count, cudaStream_t st) {
  QR_NCCL(g_nccl.AllReduce(buf, buf, count, ncclUint64, ncclMax, c->nccl, st));
  return QR_OK;
}
int comm_allreduce_sum_u32(Comm *c, uint32_t *buf, size_t count, cudaStream_t st) {
  QR_NCCL(g_nccl.AllReduce(buf, buf, count, ncclUint32, ncclSum, c->nccl, st));
  return QR_OK;
}
int comm_allreduce_sum_u64(Comm *c, unsigned long long *buf, size_t count, cudaStream_t st) {
  QR_NCCL(g_nccl.AllReduce(buf, buf, count, ncclUint64, ncclSum, c->nccl, st));
  return QR_OK;
}
int comm_allgather_bytes(Comm *c, const void *send, void *recv, size_t bytes_per_rank, cudaStream_t st) {
  QR_NCCL(g_nccl.AllGather(send, recv, bytes_per_rank, ncclUint8, c->nccl, st));
  return QR_OK;
}

// every rank contributes `n` values (host); returns all ranks' values concatenated, rank-major
int comm_allgather_host(Comm *c, const void *send, size_t bytes, std::vector<unsigned char> *out, cudaStream_t st) {
  unsigned char *d_send = nullptr, *d_recv = nullptr;
  const size_t padded = std::max<size_t>(bytes, 1);
  QR_CUDA(cudaMalloc((void **) &d_send, padded));
  QR_CUDA(cudaMalloc((void **) &d_recv, padded * c->world));
  QR_CUDA(cudaMemcpyAsync(d_send, send, bytes, cudaMemcpyHostToDevice, st));
  int rc = comm_allgather_bytes(c, d_send, d_recv, padded, st);
  if (rc == QR_OK) {
    out->resize(padded * c->world);
    cudaError_t e = cudaMemcpyAsync(out->data(), d_recv, padded * c->world, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { set_error("allgather copy failed: %s", cudaGetErrorString(e)); rc = QR_ECUDA; }
  }
  cudaFree(d_send);
  cudaFree(d_recv);
  return rc;
}

// squares: per-slice 128-bit partials -> per-task total -> three 43-bit limbs (summable as int64)
__global__ void sq_pack_kernel(const NodeTask *__restrict__ tasks, const ulonglong2 *__restrict__ sq128,
                               long long *limbs) {
  const uint32_t task = blockIdx.x;
  if (threadIdx.x != 0) return;
  const NodeTask t = tasks[task];
  U128 tot{0ull, 0ull};
  for (uint32_t i = 0; i < t.hist_nblk; ++i) { const ulonglong2 v = sq128[t.hist_blk0 + i]; u128_add(tot, v.x, v.y); }
  const unsigned long long m43 = (1ull << 43) - 1ull;
  limbs[task * 3 + 0] = (long long) (tot.lo & m43);
  limbs[task * 3 + 1] = (long long) (((tot.lo >> 43) | (tot.hi << 21)) & m43);
  limbs[task * 3 + 2] = (long long) (tot.hi >> 22);
}

// recombine the all-reduced limbs into one 128-bit partial per task (finalize_kernel then reads a
// single slice per task)
__global__ void sq_unpack_kernel(NodeTask *tasks, const long long *__restrict__ limbs, ulonglong2 *sq128,
                                 uint32_t slot0) {
  const uint32_t task = blockIdx.x;
  if (threadIdx.x != 0) return;
  U128 tot{0ull, 0ull};
  const unsigned long long l0 = (unsigned long long) limbs[task * 3 + 0];
  const unsigned long long l1 = (unsigned long long) limbs[task * 3 + 1];
  const unsigned long long l2 = (unsigned long long) limbs[task * 3 + 2];
  u128_add(tot, l0, 0ull);
  u128_add(tot, l1 << 43, l1 >> 21);           // l1 * 2^43
  u128_add(tot, 0ull, l2 << 22);               // l2 * 2^86 = (l2 << 22) * 2^64
  sq128[slot0 + task] = make_ulonglong2(tot.lo, tot.hi);
  tasks[task].hist_blk0 = slot0 + task;
  tasks[task].hist_nblk = 1;
}

int comm_reduce_tasks(qr_ctx *ctx, uint32_t k, bool root) {
  Comm *c = ctx->comm;
  cudaStream_t st = ctx->stream;
  if (ctx->exact) { set_error("reference-order accumulation is single-GPU only"); return QR_ECOMM; }
  if (!c->d_sq_limbs) QR_CUDA(cudaMalloc((void **) &c->d_sq_limbs, (size_t) ctx->max_tasks * 3 * sizeof(long long)));
  sq_pack_kernel<<<k, 32, 0, st>>>(ctx->d_tasks, ctx->d_sq128, c->d_sq_limbs);
  ctx->launches++;
  QR_CUDA(cudaGetLastError());
  QR_NCCL(g_nccl.GroupStart());
  for (uint32_t j = 0; j < k; ++j) {
    const int slot = ctx->h_tasks[j].slotB;
    unsigned long long *hs = ctx->d_hist_sum + (size_t) slot * ctx->ncells;
    uint32_t *hc = ctx->d_hist_cnt + (size_t) slot * ctx->ncells;
    QR_NCCL(g_nccl.AllReduce(hs, hs, ctx->ncells, ncclInt64, ncclSum, c->nccl, st));
    // the root's counts come from the table all-reduced once at start
    if (!(root && ctx->d_root_cnt != nullptr))
      QR_NCCL(g_nccl.AllReduce(hc, hc, ctx->ncells, ncclUint32, ncclSum, c->nccl, st));
  }
  QR_NCCL(g_nccl.AllReduce(c->d_sq_limbs, c->d_sq_limbs, (size_t) k * 3, ncclInt64, ncclSum, c->nccl, st));
  QR_NCCL(g_nccl.GroupEnd());
  // totals go to the tail of the partials array, one slice per task
  const uint32_t slot0 = ctx->max_slices - ctx->max_tasks;
  sq_unpack_kernel<<<k, 32, 0, st>>>(ctx->d_tasks, c->d_sq_limbs, ctx->d_sq128, slot0);
  ctx->launches++;
  QR_CUDA(cudaGetLastError());
  return QR_OK;
}

__global__ void leaf_values_kernel(const double2 *__restrict__ leafsum, const unsigned long long *__restrict__ leafn,
                                   uint32_t nleaves, bool newton, double *leafval) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nleaves) return;
  const double2 s = leafsum[i];
  if (newton) leafval[i] = s.y >= DBL_EPSILON ? s.x / s.y : 0.0;   // rt.cc:200
  else leafval[i] = s.x / (double) leafn[i];                       // rt.cc:178
}

int comm_leaf_values(qr_ctx *ctx, uint32_t nleaves) {
  Comm *c = ctx->comm;
  cudaStream_t st = ctx->stream;
  QR_NCCL(g_nccl.AllReduce(ctx->d_leafsum, ctx->d_leafsum, (size_t) nleaves * 2, ncclFloat64, ncclSum, c->nccl, st));
  // global leaf sizes for the MART mean
  std::vector<unsigned long long> n(nleaves);
  for (uint32_t k = 0; k < nleaves; ++k) n[k] = ctx->nodes[ctx->leaves[k]].res.n;
  unsigned long long *d_n = nullptr;
  QR_CUDA(cudaMalloc((void **) &d_n, std::max<size_t>(nleaves, 1) * sizeof(unsigned long long)));
  QR_CUDA(cudaMemcpyAsync(d_n, n.data(), nleaves * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  leaf_values_kernel<<<(nleaves + 63) / 64, 64, 0, st>>>(ctx->d_leafsum, d_n, nleaves, ctx->lambda, ctx->d_leafval);
  ctx->launches++;
  QR_CUDA(cudaGetLastError());
  QR_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_n);
  return QR_OK;
}

}  // namespace qr

extern "C" {

int qr_comm_unique_id(unsigned char id[QR_COMM_ID_BYTES]) {
  if (!id) { qr::set_error("qr_comm_unique_id: null id"); return QR_EINVAL; }
  QR_TRY(qr::load_nccl());
  ncclUniqueId uid;
  ncclResult_t r = qr::g_nccl.GetUniqueId(&uid);
  if (r != ncclSuccess) { qr::set_error("ncclGetUniqueId failed: %s", qr::g_nccl.GetErrorString(r)); return QR_ECOMM; }
  memcpy(id, uid.internal, QR_COMM_ID_BYTES);
  return QR_OK;
}

}
