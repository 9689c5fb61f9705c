// quickrank_b200 — multi-GPU plumbing: one process per GPU, documents sharded by query.  The exchange
// steps the path has:
//   * once at start: the global list of distinct feature values (thresholds must be identical on
//     every rank) and the per-bin document counts of the whole dataset (NCCL);
//   * per tree: the maximum |pseudo-response| (common fixed-point scale), per-leaf (sum lambda, sum
//     weight), the sum of per-query NDCG (NCCL);
//   * per growth round — the only exchange on the critical path: the freshly built per-bin histograms
//     (int64 sums + uint32 counts) of the round's nodes and their exact squares sums, so that every rank
//     scans the same totals and makes the same split decisions without any broadcast.  This one travels
//     through PEER MEMORY (every rank maps the other ranks' histogram pools over CUDA IPC): fused into
//     the split-scan kernel for small rounds (finalize_kernel<false, true>, qr_tree_kernels.cuh), or by
//     the stand-alone in-place reduce-scatter + all-gather kernel below (peer_reduce_kernel) for wide
//     ones; grouped NCCL all-reduces remain as the fallback when the devices cannot map each other.
// Histogram sums are integers, so the totals — and therefore every split — do not depend on the number
// of ranks, on the exchange path or on the order of the additions.
//
// NCCL is resolved with dlopen at run time (libnccl.so.2; inside a PyTorch process this is the
// copy torch already loaded), so the single-GPU path has no link-time dependency on it.
#include "qr_comm.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <string>

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

// ---- peer-memory exchange (NVLink / NVSwitch) -------------------------------------------------
// Every rank exports its histogram pool and a small signal buffer with CUDA IPC; peers map them, so a
// kernel on one GPU can load and store the other GPUs' histograms directly.
struct PeerSignals {
  uint32_t flags[kMaxPeers];         // flags[p]: last barrier epoch announced by rank p (stored by p)
  uint32_t magic;                    // mapping check at set-up
  uint32_t pad[23];
  ulonglong2 sq_local[1];            // [max_tasks] this rank's exact squares sums of the current round
};
struct PeerTable {
  unsigned long long *sum[kMaxPeers];  // histogram pools (sums), own pool at [rank]
  uint32_t *cnt[kMaxPeers];            // histogram pools (counts)
  PeerSignals *sig[kMaxPeers];
  ulonglong2 *sq[kMaxPeers];           // squares partials (d_sq128)
};

struct Comm {
  int rank = 0, world = 1;
  ncclComm_t nccl = nullptr;         // shared with the other contexts of this process that use the same id
  std::string key;
  long long *d_sq_limbs = nullptr;   // [max_tasks][3] 43-bit limbs of the squares sums
  double *d_scratch = nullptr;       // small host<->device staging
  unsigned long long *d_leafn = nullptr;   // [maxleaves] global leaf sizes (MART mean)
  size_t leafn_cap = 0;
  // peer-memory path
  bool peer_ok = false;
  PeerTable peers{};
  PeerSignals *d_sig = nullptr;      // own signal buffer (exported)
  uint32_t *d_done = nullptr;        // block counter of peer_reduce_kernel
  uint32_t epoch = 0;                // two barrier epochs per reduce launch
  size_t mail_off = 0;               // byte offset of the candidate mailbox inside the signal buffer
  uint32_t mail_tasks = 0;
  void *opened[4 * kMaxPeers] = {nullptr};
  int nopened = 0;
};

int comm_rank(const Comm *c) { return c ? c->rank : 0; }
int comm_world(const Comm *c) { return c ? c->world : 1; }

// NCCL communicators are process-level plumbing (bootstrap + transport set-up take seconds): contexts
// created with the same id in one process share one communicator, and an idle communicator is kept until
// a different id is asked for.  (Single caller thread, like everything else here; the contexts of one
// process issue their collectives one after the other, in the same order on every rank.)
struct SharedNccl { ncclComm_t comm = nullptr; int refs = 0, rank = 0, world = 0; };
static std::map<std::string, SharedNccl> g_comms;

int comm_create(const unsigned char id[QR_COMM_ID_BYTES], int rank, int world, Comm **out) {
  *out = nullptr;
  if (world < 1 || rank < 0 || rank >= world) { set_error("bad rank %d / world %d", rank, world); return QR_EINVAL; }
  QR_TRY(load_nccl());
  const std::string key(reinterpret_cast<const char *>(id), QR_COMM_ID_BYTES);
  auto it = g_comms.find(key);
  if (it != g_comms.end() && (it->second.rank != rank || it->second.world != world)) {
    set_error("communicator id already in use in this process as rank %d of %d", it->second.rank, it->second.world);
    return QR_EINVAL;
  }
  if (it == g_comms.end()) {
    for (auto jt = g_comms.begin(); jt != g_comms.end();) {   // idle communicators of other ids
      if (jt->second.refs == 0) { g_nccl.CommDestroy(jt->second.comm); jt = g_comms.erase(jt); }
      else ++jt;
    }
    ncclUniqueId uid;
    static_assert(sizeof(uid.internal) == QR_COMM_ID_BYTES, "NCCL unique id size");
    memcpy(uid.internal, id, QR_COMM_ID_BYTES);
    SharedNccl sh;
    sh.rank = rank;
    sh.world = world;
    ncclResult_t r = g_nccl.CommInitRank(&sh.comm, world, uid, rank);
    if (r != ncclSuccess) {
      set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
      return QR_ECOMM;
    }
    it = g_comms.emplace(key, sh).first;
  }
  it->second.refs++;
  Comm *c = new Comm();
  c->rank = rank;
  c->world = world;
  c->nccl = it->second.comm;
  c->key = key;
  *out = c;
  return QR_OK;
}

void comm_destroy(Comm *c) {
  if (!c) return;
  for (int i = 0; i < c->nopened; ++i) cudaIpcCloseMemHandle(c->opened[i]);
  if (c->d_sig) cudaFree(c->d_sig);
  if (c->d_done) cudaFree(c->d_done);
  if (c->d_leafn) cudaFree(c->d_leafn);
  if (c->d_sq_limbs) cudaFree(c->d_sq_limbs);
  if (c->d_scratch) cudaFree(c->d_scratch);
  auto it = g_comms.find(c->key);
  if (it != g_comms.end() && it->second.refs > 0) it->second.refs--;   // the communicator stays cached
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

// ------------------------------------------------------------------------------------------
// All-reduce of a growth round's freshly built histograms over peer memory, in ONE kernel per rank:
//   0. block 0 totals this rank's exact squares partials into its signal buffer, then tells every
//      peer "my histogram kernel of this round has finished" (the kernel is stream-ordered after it);
//   1. every block waits for the same announcement from all peers (barrier A);
//   2. reduce-scatter + all-gather in place: rank r owns the cells [ncells*r/W, ncells*(r+1)/W) of every
//      task's slot; it loads them from all W pools, adds (64-bit integers and counts: exact, any order)
//      and stores the totals into all W pools.  In phase 2 nobody else reads or writes those cells;
//   3. the last block to finish tells the peers and waits for theirs (barrier B): when the kernel ends,
//      every pool holds the global histograms and no peer is still reading this rank's partials, so
//      the split scan that follows — and the next round's histogram kernel — need no further fence.
// Per rank and round this moves 2 x (W-1)/W x k x H bytes over NVLink and costs two flag round trips;
// the NCCL path it replaces was 2k+1 grouped all-reduces plus two pack/unpack launches.
// ------------------------------------------------------------------------------------------
constexpr uint32_t kPeerThreads = 256;

__global__ void __launch_bounds__(kPeerThreads)
peer_reduce_kernel(const PeerTable pt, int rank, int world, const NodeTask *__restrict__ tasks, uint32_t k,
                   uint32_t ncells, uint32_t epoch_a, int with_counts, ulonglong2 *sq128, uint32_t *done,
                   uint32_t *host_err, const __grid_constant__ TaskPack pack) {
  if (pack.n) tasks = pack.t;
  __shared__ uint32_t s_last;
  PeerSignals *mine = pt.sig[rank];
  const uint32_t tid = threadIdx.x;
  if (blockIdx.x == 0) {
    // one warp per task: the slices' 128-bit partials are loaded in parallel and folded with shuffles
    const uint32_t lane = tid & 31u;
    for (uint32_t j = tid >> 5; j < k; j += kPeerThreads / 32u) {
      const uint32_t blk0 = tasks[j].hist_blk0, nblk = tasks[j].hist_nblk;
      U128 tot{0ull, 0ull};
      for (uint32_t i = lane; i < nblk; i += 32u) { const ulonglong2 v = sq128[blk0 + i]; u128_add(tot, v.x, v.y); }
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ol = __shfl_xor_sync(0xffffffffu, tot.lo, o);
        const unsigned long long oh = __shfl_xor_sync(0xffffffffu, tot.hi, o);
        u128_add(tot, ol, oh);
      }
      if (lane == 0) {
        volatile unsigned long long *o = reinterpret_cast<volatile unsigned long long *>(&mine->sq_local[j]);
        o[0] = tot.lo; o[1] = tot.hi;
      }
    }
    __syncthreads();
    if (tid < (uint32_t) world && tid != (uint32_t) rank) st_flag(&pt.sig[tid]->flags[rank], epoch_a);
  }
  // barrier A
  if (tid < (uint32_t) world && tid != (uint32_t) rank) wait_flag_or_report(&mine->flags[tid], epoch_a, host_err);
  __syncthreads();

  const uint32_t c0 = (uint32_t) ((unsigned long long) ncells * (uint32_t) rank / (uint32_t) world);
  const uint32_t c1 = (uint32_t) ((unsigned long long) ncells * ((uint32_t) rank + 1u) / (uint32_t) world);
  const uint32_t span = c1 - c0;
  const uint32_t total = k * span;
  for (uint32_t idx = blockIdx.x * kPeerThreads + tid; idx < total; idx += gridDim.x * kPeerThreads) {
    const uint32_t j = idx / span, i = c0 + (idx - j * span);
    const size_t off = (size_t) build_slot(tasks[j]) * ncells + i;
    // every pool's copy in flight together (as in scan_pub_kernel: left to itself ptxas adds each value as it arrives,
    // which serialises the NVLink round trips)
    unsigned long long v[kMaxPeers];
    uint32_t m[kMaxPeers];
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      v[p] = 0ull; m[p] = 0u;
      if (p < world) {
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v[p]) : "l"(pt.sum[p] + off) : "memory");
        if (with_counts) asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(m[p]) : "l"(pt.cnt[p] + off) : "memory");
      }
    }
    static_assert(kMaxPeers == 8, "the operand list below names eight pools");
    asm volatile("" : "+l"(v[0]), "+l"(v[1]), "+l"(v[2]), "+l"(v[3]), "+l"(v[4]), "+l"(v[5]), "+l"(v[6]), "+l"(v[7]),
                      "+r"(m[0]), "+r"(m[1]), "+r"(m[2]), "+r"(m[3]), "+r"(m[4]), "+r"(m[5]), "+r"(m[6]), "+r"(m[7]));
    unsigned long long s = 0ull;
    uint32_t n = 0u;
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) { s += v[p]; n += m[p]; }
#pragma unroll
    for (int p = 0; p < kMaxPeers; ++p) {
      if (p < world) {
        pt.sum[p][off] = s;
        if (with_counts) pt.cnt[p][off] = n;
      }
    }
  }
  // squares: every rank adds the same W totals, so all of them hold the same exact value; the total goes
  // to the task's first slice and the others are zeroed (finalize_kernel sums the task's slices)
  if (blockIdx.x == 0) {
    for (uint32_t j = tid; j < k; j += kPeerThreads) {
      const NodeTask t = tasks[j];
      U128 tot{0ull, 0ull};
#pragma unroll
      for (int p = 0; p < kMaxPeers; ++p) {
        if (p < world) {
          const volatile unsigned long long *v = reinterpret_cast<const volatile unsigned long long *>(&pt.sig[p]->sq_local[j]);
          const unsigned long long lo = v[0], hi = v[1];
          u128_add(tot, lo, hi);
        }
      }
      sq128[t.hist_blk0] = make_ulonglong2(tot.lo, tot.hi);
      for (uint32_t i = 1; i < t.hist_nblk; ++i) sq128[t.hist_blk0 + i] = make_ulonglong2(0ull, 0ull);
    }
  }
  // barrier B, by the last block of this rank
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    s_last = atomicAdd(done, 1u) == gridDim.x - 1u ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) *done = 0u;   // for the next launch (stream-ordered after this one)
  if (tid < (uint32_t) world && tid != (uint32_t) rank) {
    st_flag(&pt.sig[tid]->flags[rank], epoch_a + 1u);
    wait_flag_or_report(&mine->flags[tid], epoch_a + 1u, host_err);
  }
}

struct PeerHello {
  cudaIpcMemHandle_t sum, cnt, sig, sq;
  int ok;
  int device;
};

// Exports this rank's pools, maps the peers', checks every mapping and agrees with all ranks on whether
// the peer-memory path is used (all ranks must take the same path: its barriers involve everyone).
int comm_setup_peers(qr_ctx *ctx) {
  Comm *c = ctx->comm;
  if (!c || ctx->d_hist_sum == nullptr) return QR_OK;
  cudaStream_t st = ctx->stream;
  const char *env = getenv("QR_PEER_REDUCE");
  int ok = (env == nullptr || atoi(env) != 0) && c->world <= kMaxPeers;
  const char *why = ok ? "" : (c->world > kMaxPeers ? "more than 8 ranks" : "QR_PEER_REDUCE=0");
  // signal buffer: flags | squares of the current round | candidate mailbox of the feature-sliced exchange
  // ([2 parities][kMaxPeers sources][max_tasks][2 children] records of kMailWords words, zero = never written)
  c->mail_off = (sizeof(PeerSignals) + (size_t) ctx->max_tasks * sizeof(ulonglong2) + 255) & ~(size_t) 255;
  c->mail_tasks = ctx->max_tasks;
  const size_t mail_bytes = (size_t) 2 * kMaxPeers * ctx->max_tasks * 2 * kMailWords * sizeof(unsigned long long);
  const size_t sig_bytes = std::max<size_t>((size_t) 2 << 20, c->mail_off + mail_bytes);
  QR_CUDA(cudaMalloc((void **) &c->d_sig, sig_bytes));
  QR_CUDA(cudaMalloc((void **) &c->d_done, sizeof(uint32_t)));
  QR_CUDA(cudaMemsetAsync(c->d_sig, 0, sig_bytes, st));
  QR_CUDA(cudaMemsetAsync(c->d_done, 0, sizeof(uint32_t), st));
  // marks the peers will look for through their mappings (the pools are scratch: slots are cleared before use)
  const uint32_t magic = 0x51b20000u + (uint32_t) c->rank;
  const unsigned long long magic64 = 0x51b2000051b20000ull + (unsigned long long) c->rank;
  QR_CUDA(cudaMemcpyAsync(&c->d_sig->magic, &magic, sizeof(magic), cudaMemcpyHostToDevice, st));
  QR_CUDA(cudaMemcpyAsync(ctx->d_hist_sum, &magic64, sizeof(magic64), cudaMemcpyHostToDevice, st));
  QR_CUDA(cudaMemcpyAsync(ctx->d_hist_cnt, &magic, sizeof(magic), cudaMemcpyHostToDevice, st));
  QR_CUDA(cudaMemcpyAsync(ctx->d_sq128, &magic64, sizeof(magic64), cudaMemcpyHostToDevice, st));
  QR_CUDA(cudaStreamSynchronize(st));
  PeerHello hello;
  memset(&hello, 0, sizeof(hello));
  hello.device = ctx->device;
  if (ok) {
    if (cudaIpcGetMemHandle(&hello.sum, ctx->d_hist_sum) != cudaSuccess ||
        cudaIpcGetMemHandle(&hello.cnt, ctx->d_hist_cnt) != cudaSuccess ||
        cudaIpcGetMemHandle(&hello.sig, c->d_sig) != cudaSuccess ||
        cudaIpcGetMemHandle(&hello.sq, ctx->d_sq128) != cudaSuccess) {
      ok = 0; why = "cudaIpcGetMemHandle failed";
      cudaGetLastError();
    }
  }
  hello.ok = ok;
  std::vector<unsigned char> all;
  QR_TRY(comm_allgather_host(c, &hello, sizeof(hello), &all, st));
  const PeerHello *hs = reinterpret_cast<const PeerHello *>(all.data());
  for (int p = 0; p < c->world; ++p) if (!hs[p].ok && ok) { ok = 0; why = "a peer cannot export its memory"; }
  if (ok) {
    c->peers.sum[c->rank] = ctx->d_hist_sum;
    c->peers.cnt[c->rank] = ctx->d_hist_cnt;
    c->peers.sig[c->rank] = c->d_sig;
    c->peers.sq[c->rank] = ctx->d_sq128;
    for (int p = 0; p < c->world && ok; ++p) {
      if (p == c->rank) continue;
      void *ptrs[4] = {nullptr, nullptr, nullptr, nullptr};
      const cudaIpcMemHandle_t *hd[4] = {&hs[p].sum, &hs[p].cnt, &hs[p].sig, &hs[p].sq};
      for (int i = 0; i < 4 && ok; ++i) {
        if (cudaIpcOpenMemHandle(&ptrs[i], *hd[i], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          ok = 0; why = "cudaIpcOpenMemHandle failed (no peer access between the devices?)";
          cudaGetLastError();
        } else {
          c->opened[c->nopened++] = ptrs[i];
        }
      }
      if (!ok) break;
      c->peers.sum[p] = (unsigned long long *) ptrs[0];
      c->peers.cnt[p] = (uint32_t *) ptrs[1];
      c->peers.sig[p] = (PeerSignals *) ptrs[2];
      c->peers.sq[p] = (ulonglong2 *) ptrs[3];
      unsigned long long m64 = 0, m64b = 0;
      uint32_t m32a = 0, m32b = 0;
      if (cudaMemcpy(&m64, c->peers.sum[p], sizeof(m64), cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(&m32a, c->peers.cnt[p], sizeof(m32a), cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(&m32b, &c->peers.sig[p]->magic, sizeof(m32b), cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(&m64b, c->peers.sq[p], sizeof(m64b), cudaMemcpyDeviceToHost) != cudaSuccess ||
          m64b != 0x51b2000051b20000ull + (unsigned long long) p ||
          m64 != 0x51b2000051b20000ull + (unsigned long long) p || m32a != 0x51b20000u + (uint32_t) p ||
          m32b != 0x51b20000u + (uint32_t) p) {
        ok = 0; why = "a mapped peer buffer does not show the peer's mark";
        cudaGetLastError();
      }
    }
  }
  // agreement: one rank's failure turns the path off everywhere
  unsigned long long bad = ok ? 0ull : 1ull, *d_bad = nullptr;
  QR_CUDA(cudaMalloc((void **) &d_bad, sizeof(bad)));
  QR_CUDA(cudaMemcpyAsync(d_bad, &bad, sizeof(bad), cudaMemcpyHostToDevice, st));
  QR_TRY(comm_allreduce_max_u64(c, d_bad, 1, st));
  unsigned long long any_bad = 1;
  QR_CUDA(cudaMemcpyAsync(&any_bad, d_bad, sizeof(any_bad), cudaMemcpyDeviceToHost, st));
  QR_CUDA(cudaStreamSynchronize(st));
  cudaFree(d_bad);
  c->peer_ok = any_bad == 0ull;
  if (!c->peer_ok) {
    if (!ok && env == nullptr)
      fprintf(stderr, "quickrank_b200: rank %d: peer-memory histogram exchange unavailable (%s); using NCCL all-reduces\n",
              c->rank, why);
    for (int i = 0; i < c->nopened; ++i) cudaIpcCloseMemHandle(c->opened[i]);
    c->nopened = 0;
  }
  return QR_OK;
}

int comm_transport(const Comm *c) { return c == nullptr ? 0 : (c->peer_ok ? 2 : 1); }

// the view finalize_kernel needs to add the peers' staging slots itself (one barrier epoch per round)
void comm_peer_view(qr_ctx *ctx, bool with_counts, PeerView *pv) {
  Comm *c = ctx->comm;
  memset(pv, 0, sizeof(*pv));
  for (int p = 0; p < c->world; ++p) {
    pv->sum[p] = c->peers.sum[p];
    pv->cnt[p] = c->peers.cnt[p];
    pv->sq[p] = c->peers.sq[p] + ctx->round_sq_off;
    pv->peer_flags[p] = c->peers.sig[p]->flags;
  }
  pv->flags = c->d_sig->flags;
  pv->rank = c->rank;
  pv->world = c->world;
  pv->epoch = ++c->epoch;
  pv->with_counts = with_counts ? 1 : 0;
  if (ctx->sliced) {
    // contiguous feature ranges, ascending with the rank: the first maximum over ranks is the first over features
    pv->f_lo = (uint32_t) ((size_t) ctx->F * (size_t) c->rank / (size_t) c->world);
    pv->f_hi = (uint32_t) ((size_t) ctx->F * (size_t) (c->rank + 1) / (size_t) c->world);
    for (int p = 0; p < c->world; ++p)
      pv->mail[p] = reinterpret_cast<unsigned long long *>(reinterpret_cast<unsigned char *>(c->peers.sig[p]) + c->mail_off);
    pv->mail_tasks = c->mail_tasks;
  }
}

static int peer_reduce_tasks(qr_ctx *ctx, uint32_t k, bool root) {
  Comm *c = ctx->comm;
  const int with_counts = (root && ctx->d_root_cnt != nullptr) ? 0 : 1;
  const uint32_t span = ctx->ncells / (uint32_t) c->world + 1u;
  const uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>(2u * 148u, (k * span + kPeerThreads - 1) / kPeerThreads));
  const uint32_t epoch_a = c->epoch + 1u;
  c->epoch += 2u;
  peer_reduce_kernel<<<grid, kPeerThreads, 0, ctx->stream>>>(c->peers, c->rank, c->world, ctx->d_tasks, k, ctx->ncells,
                                                             epoch_a, with_counts, ctx->d_sq128 + ctx->round_sq_off, c->d_done,
                                                             ctx->d_err_mapped, ctx->pack);
  ctx->launches++;
  QR_CUDA(cudaGetLastError());
  return QR_OK;
}

int comm_reduce_tasks(qr_ctx *ctx, uint32_t k, bool root) {
  Comm *c = ctx->comm;
  cudaStream_t st = ctx->stream;
  if (ctx->exact) { set_error("reference-order accumulation is single-GPU only"); return QR_ECOMM; }
  if (c->peer_ok) return peer_reduce_tasks(ctx, k, root);
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

__global__ void leaf_values_kernel(const longlong2 *__restrict__ leafsum, const unsigned long long *__restrict__ leafn,
                                   uint32_t nleaves, bool newton, const int *__restrict__ qexp, double *leafval) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nleaves) return;
  leafval[i] = leaf_value_of(leafsum[i], leafn[i], newton, qexp);   // exact integer sums: the same on every rank
}

int comm_leaf_values(qr_ctx *ctx, uint32_t nleaves) {
  Comm *c = ctx->comm;
  cudaStream_t st = ctx->stream;
  QR_NCCL(g_nccl.AllReduce(ctx->d_leafsum, ctx->d_leafsum, (size_t) nleaves * 2, ncclInt64, ncclSum, c->nccl, st));
  // global leaf sizes for the MART mean
  std::vector<unsigned long long> n(std::max<uint32_t>(nleaves, 1));
  for (uint32_t k = 0; k < nleaves; ++k) n[k] = ctx->nodes[ctx->leaves[k]].res.n;
  if (c->leafn_cap < n.size()) {
    if (c->d_leafn) cudaFree(c->d_leafn);
    c->leafn_cap = std::max<size_t>(n.size(), 256);
    QR_CUDA(cudaMalloc((void **) &c->d_leafn, c->leafn_cap * sizeof(unsigned long long)));
  }
  QR_CUDA(cudaMemcpyAsync(c->d_leafn, n.data(), nleaves * sizeof(unsigned long long), cudaMemcpyHostToDevice, st));
  leaf_values_kernel<<<(nleaves + 63) / 64, 64, 0, st>>>(reinterpret_cast<const longlong2 *>(ctx->d_leafsum), c->d_leafn, nleaves, ctx->lambda, ctx->d_qexp, ctx->d_leafval);
  ctx->launches++;
  QR_CUDA(cudaGetLastError());
  QR_CUDA(cudaStreamSynchronize(st));   // `n` is pageable: the copy above must have left it
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
