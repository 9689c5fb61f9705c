// Development micro-benchmark #2: inner-loop variants of the fixed-point shared-memory histogram.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o scripts/hist_mb2 scripts/hist_mb2.cu
// Data shaped like bench.py's (256-level skewed features); identity (root) and gathered (child) lists.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <random>
#include <cmath>
#include <algorithm>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int FPP = 16;
constexpr int B = 256;
constexpr int CELLS = FPP * B;

__device__ __forceinline__ uint4 rotate_bytes(uint4 v, uint32_t rb) {
  if (rb & 4u) { uint32_t t = v.x; v.x = v.y; v.y = v.z; v.z = v.w; v.w = t; }
  if (rb & 8u) { uint32_t t = v.x; v.x = v.z; v.z = t; t = v.y; v.y = v.w; v.w = t; }
  const uint32_t sel = 0x3210u + 0x1111u * (rb & 3u);
  uint4 r;
  r.x = __byte_perm(v.x, v.y, sel); r.y = __byte_perm(v.y, v.z, sel);
  r.z = __byte_perm(v.z, v.w, sel); r.w = __byte_perm(v.w, v.x, sel);
  return r;
}
__device__ __forceinline__ uint32_t ext(const uint4 &v, int j) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  return (w[j >> 2] >> ((j & 3) * 8)) & 0xffu;
}

// ---- baseline: production-like loop (feature-major cells, add-rotation, hot bypass via branches) ----
template <bool COUNT>
__device__ __forceinline__ void add_row_base(const uint4 &row, const uint4 &hotx, long long q, uint32_t rot,
                                             uint32_t *s_lo, int32_t *s_hi, uint32_t *s_cnt) {
  const uint32_t qlo = (uint32_t) q;
  const int32_t qhi = (int32_t) (q >> 32);
#pragma unroll
  for (int h0 = 0; h0 < 16; h0 += 8) {
    uint32_t cell[8], old[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint32_t slot = (h0 + j + rot) & 15;
      const bool on = ext(hotx, h0 + j) != 0u;
      cell[j] = on ? slot * B + ext(row, h0 + j) : 0xffffffffu;
      if (on) old[j] = atomicAdd(s_lo + cell[j], qlo);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (cell[j] != 0xffffffffu) {
        const int32_t carry = (old[j] + qlo) < old[j];
        atomicAdd(s_hi + cell[j], qhi + carry);
        if (COUNT) atomicAdd(s_cnt + cell[j], 1u);
      }
    }
  }
}

template <bool COUNT, bool GATHER>
__global__ void __launch_bounds__(256, 4)
k_base(const uint4 *__restrict__ panels, size_t N, const long long *__restrict__ lamq, const uint32_t *__restrict__ ids,
       uint32_t n, uint32_t dpb, unsigned long long *gsum, uint32_t *gcnt, const uint4 *__restrict__ hot_rows) {
  extern __shared__ __align__(16) unsigned char sm[];
  uint32_t *s_lo = (uint32_t *) sm;
  int32_t *s_hi = (int32_t *) (s_lo + CELLS);
  uint32_t *s_cnt = (uint32_t *) (s_hi + CELLS);
  for (int i = threadIdx.x; i < CELLS; i += 256) { s_lo[i] = 0; s_hi[i] = 0; if (COUNT) s_cnt[i] = 0; }
  __syncthreads();
  const uint32_t p = blockIdx.y;
  const uint4 *prow = panels + (size_t) p * N;
  const uint32_t begin = blockIdx.x * dpb, end = min(n, begin + dpb);
  const uint32_t rot = threadIdx.x & 15, rotb = rot;
  const uint4 hot = rotate_bytes(hot_rows[p], rotb);
  long long totq = 0; uint32_t totn = 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += 512) {
    const uint32_t i1 = i + 256;
    const bool has1 = i1 < end;
    const uint32_t d0 = GATHER ? ids[i] : i;
    const uint32_t d1 = has1 ? (GATHER ? ids[i1] : i1) : d0;
    const uint4 raw0 = prow[d0], raw1 = prow[d1];
    const long long q0 = lamq[d0], q1 = lamq[d1];
    totq += q0 + (has1 ? q1 : 0ll); totn += has1 ? 2u : 1u;
    const uint4 r0 = rotate_bytes(raw0, rotb), r1 = rotate_bytes(raw1, rotb);
    const uint4 x0 = make_uint4(r0.x ^ hot.x, r0.y ^ hot.y, r0.z ^ hot.z, r0.w ^ hot.w);
    const uint4 x1 = make_uint4(r1.x ^ hot.x, r1.y ^ hot.y, r1.z ^ hot.z, r1.w ^ hot.w);
    add_row_base<COUNT>(r0, x0, q0, rot, s_lo, s_hi, s_cnt);
    if (has1) add_row_base<COUNT>(r1, x1, q1, rot, s_lo, s_hi, s_cnt);
  }
  __shared__ long long s_tq[8]; __shared__ uint32_t s_tn[8];
  for (int o = 16; o > 0; o >>= 1) { totq += __shfl_xor_sync(~0u, totq, o); totn += __shfl_xor_sync(~0u, totn, o); }
  if ((threadIdx.x & 31) == 0) { s_tq[threadIdx.x >> 5] = totq; s_tn[threadIdx.x >> 5] = totn; }
  __syncthreads();
  long long bq = 0; uint32_t bn = 0;
  for (int w = 0; w < 8; ++w) { bq += s_tq[w]; bn += s_tn[w]; }
  unsigned long long *gs = gsum + (size_t) p * CELLS;
  uint32_t *gc = gcnt + (size_t) p * CELLS;
  const uint8_t *hotb = (const uint8_t *) (hot_rows + p);
  for (uint32_t slot = threadIdx.x >> 5; slot < 16; slot += 8) {
    long long oq = 0; uint32_t on = 0;
    for (uint32_t b = threadIdx.x & 31; b < B; b += 32) {
      const uint32_t i = slot * B + b;
      const long long v = ((long long) s_hi[i] << 32) + (long long) s_lo[i];
      const uint32_t cn = COUNT ? s_cnt[i] : 0u;
      if (v != 0) atomicAdd(gs + i, (unsigned long long) v);
      if (COUNT && cn) atomicAdd(gc + i, cn);
      oq += v; on += cn;
    }
    for (int o = 16; o > 0; o >>= 1) { oq += __shfl_xor_sync(~0u, oq, o); on += __shfl_xor_sync(~0u, on, o); }
    if ((threadIdx.x & 31) == 0) {
      const uint32_t hb = hotb[slot];
      atomicAdd(gs + slot * B + hb, (unsigned long long) (bq - oq));
      if (COUNT) atomicAdd(gc + slot * B + hb, bn - on);
    }
  }
}

// ---- lean: bin-major cells (x*16 + slot), XOR rotation, predicated PTX atomics, no branches ----
// permute the 16 bytes: result byte j = input byte (j ^ r)
__device__ __forceinline__ uint4 xor_permute(uint4 v, uint32_t r, uint32_t sel) {
  if (r & 4u) { uint32_t t = v.x; v.x = v.y; v.y = t; t = v.z; v.z = v.w; v.w = t; }
  if (r & 8u) { uint32_t t = v.x; v.x = v.z; v.z = t; t = v.y; v.y = v.w; v.w = t; }
  uint4 o;
  o.x = __byte_perm(v.x, 0, sel); o.y = __byte_perm(v.y, 0, sel);
  o.z = __byte_perm(v.z, 0, sel); o.w = __byte_perm(v.w, 0, sel);
  return o;
}

// MODE 0: returning lo + hi(+carry) [+cnt];  MODE 1: three non-returning words (lo, sum of qlo>>16, hi) [+cnt]
template <bool COUNT, int MODE, int H>
__device__ __forceinline__ void add_row_lean(const uint4 &x, long long q, uint32_t cinc, unsigned char *rbp /* smem base | rot*4 */) {
  const uint32_t qlo = (uint32_t) q;
  const uint32_t qhi = (uint32_t) (q >> 32);
  constexpr uint32_t HI = CELLS * 4, MID = CELLS * 8, CNT = MODE == 1 ? CELLS * 12 : CELLS * 8;
  const uint32_t rb = (uint32_t) (uintptr_t) rbp;   // low bits only; used for the xor
#pragma unroll
  for (int h0 = 0; h0 < 16; h0 += H) {
    unsigned char *addr[H];
    uint32_t old[H];
#pragma unroll
    for (int j = 0; j < H; ++j) {
      const uint32_t xb = ext(x, h0 + j);
      addr[j] = rbp + ((xb << 6) + ((rb ^ (uint32_t) ((h0 + j) * 4)) - rb));
      if (MODE == 0) old[j] = atomicAdd((uint32_t *) addr[j], qlo);
      else atomicAdd((uint32_t *) addr[j], qlo);
    }
#pragma unroll
    for (int j = 0; j < H; ++j) {
      if (MODE == 0) {
        const uint32_t carry = (old[j] + qlo) < old[j];
        atomicAdd((uint32_t *) (addr[j] + HI), qhi + carry);
      } else {
        atomicAdd((uint32_t *) (addr[j] + MID), qlo >> 16);
        atomicAdd((uint32_t *) (addr[j] + HI), qhi);
      }
      if (COUNT) atomicAdd((uint32_t *) (addr[j] + CNT), cinc);
    }
  }
}

template <bool COUNT, int MODE, bool GATHER, int PIPE, int MINB, int T>
__global__ void __launch_bounds__(T, MINB)
k_lean(const uint4 *__restrict__ panels, size_t N, const long long *__restrict__ lamq, const uint32_t *__restrict__ ids,
       uint32_t n, uint32_t dpb, unsigned long long *gsum, uint32_t *gcnt, const uint4 *__restrict__ hot_rows) {
  extern __shared__ __align__(1024) unsigned char sm[];
  constexpr int WORDS = (MODE == 1 ? 3 : 2) + (COUNT ? 1 : 0);
  constexpr int H = MODE == 0 ? 8 : 8;
  uint32_t *s = (uint32_t *) sm;
  for (int i = threadIdx.x; i < CELLS * WORDS; i += T) s[i] = 0;
  __syncthreads();
  const uint32_t p = blockIdx.y;
  const uint4 *prow = panels + (size_t) p * N;
  const uint32_t begin = blockIdx.x * dpb, end = min(n, begin + dpb);
  const uint32_t rot = threadIdx.x & 15;
  const uint32_t sel = 0x3210u ^ (0x1111u * (rot & 3u));
  unsigned char *rbp = sm + rot * 4u;
  if (PIPE == 0) {
    for (uint32_t i = begin + threadIdx.x; i < end; i += 2 * T) {
      const uint32_t i1 = i + T;
      const bool has1 = i1 < end;
      const uint32_t d0 = GATHER ? ids[i] : i;
      const uint32_t d1 = has1 ? (GATHER ? ids[i1] : i1) : d0;
      const uint4 raw0 = prow[d0], raw1 = prow[d1];
      const long long q0 = lamq[d0], q1 = has1 ? lamq[d1] : 0ll;
      const uint4 x0 = xor_permute(raw0, rot, sel), x1 = xor_permute(raw1, rot, sel);
      add_row_lean<COUNT, MODE, H>(x0, q0, 1u, rbp);
      add_row_lean<COUNT, MODE, H>(x1, q1, has1 ? 1u : 0u, rbp);
    }
  } else {
    // software pipeline: rows of iteration t+1 are in flight while iteration t updates shared memory
    uint32_t i = begin + threadIdx.x;
    uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
    long long q0 = 0, q1 = 0;
    bool v0 = i < end, v1 = i + T < end;
    uint32_t nd0 = 0, nd1 = 0;   // document ids of the NEXT iteration
    if (v0) { const uint32_t d = GATHER ? ids[i] : i; c0 = prow[d]; q0 = lamq[d]; }
    if (v1) { const uint32_t d = GATHER ? ids[i + T] : i + T; c1 = prow[d]; q1 = lamq[d]; }
    bool w0 = i + 2 * T < end, w1 = i + 3 * T < end;
    if (w0) nd0 = GATHER ? ids[i + 2 * T] : i + 2 * T;
    if (w1) nd1 = GATHER ? ids[i + 3 * T] : i + 3 * T;
    while (v0) {
      uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
      long long nq0 = 0, nq1 = 0;
      if (w0) { n0 = prow[nd0]; nq0 = lamq[nd0]; }
      if (w1) { n1 = prow[nd1]; nq1 = lamq[nd1]; }
      i += 2 * T;
      const bool z0 = i + 2 * T < end, z1 = i + 3 * T < end;
      if (z0) nd0 = GATHER ? ids[i + 2 * T] : i + 2 * T;
      if (z1) nd1 = GATHER ? ids[i + 3 * T] : i + 3 * T;
      const uint4 x0 = xor_permute(c0, rot, sel), x1 = xor_permute(c1, rot, sel);
      add_row_lean<COUNT, MODE, H>(x0, q0, 1u, rbp);
      add_row_lean<COUNT, MODE, H>(x1, q1, v1 ? 1u : 0u, rbp);
      c0 = n0; c1 = n1; q0 = nq0; q1 = nq1; v0 = w0; v1 = w1; w0 = z0; w1 = z1;
    }
  }
  __syncthreads();
  unsigned long long *gs = gsum + (size_t) p * CELLS;
  uint32_t *gc = gcnt + (size_t) p * CELLS;
  const uint32_t *s_lo = s, *s_hi = s + CELLS, *s_mid = s + 2 * CELLS, *s_cnt = s + (MODE == 1 ? 3 : 2) * CELLS;
  // flush: consecutive threads take consecutive shared cells (cell = bin*16 + slot)
  for (uint32_t i = threadIdx.x; i < CELLS; i += T) {
    const uint32_t slot = i & 15, bin = i >> 4;
    long long v;
    if (MODE == 0) v = ((long long) (int32_t) s_hi[i] << 32) + (long long) s_lo[i];
    else {
      const unsigned long long base = (unsigned long long) s_mid[i] << 16;
      const unsigned long long L = base + (uint32_t) (s_lo[i] - (uint32_t) base);
      v = ((long long) (int32_t) s_hi[i] << 32) + (long long) L;
    }
    const uint32_t cn = COUNT ? s_cnt[i] : 0u;
    if (v != 0) atomicAdd(gs + slot * B + bin, (unsigned long long) v);
    if (COUNT && cn) atomicAdd(gc + slot * B + bin, cn);
  }
}

int main(int argc, char **argv) {
  const size_t N = argc > 1 ? atol(argv[1]) : 1000000;
  const int P = 9;
  const double frac = argc > 2 ? atof(argv[2]) : 0.3;
  std::mt19937_64 rng(1);
  std::vector<uint8_t> h((size_t) P * N * 16);
  std::uniform_real_distribution<double> U(0, 1);
  for (int p = 0; p < P; ++p)
    for (size_t d = 0; d < N; ++d)
      for (int j = 0; j < 16; ++j) {
        int f = p * 16 + j;
        double v = pow(U(rng), 1 + f % 3);
        uint8_t b = (uint8_t) llround(255 * v);
        if (f % 20 == 19 || f >= 136) b = 0;
        h[((size_t) p * N + d) * 16 + j] = b;
      }
  // hot bins: most frequent bin per feature
  std::vector<uint8_t> hot((size_t) P * 16, 0);
  for (int p = 0; p < P; ++p)
    for (int j = 0; j < 16; ++j) {
      std::vector<uint32_t> c(256, 0);
      for (size_t d = 0; d < N; d += 7) c[h[((size_t) p * N + d) * 16 + j]]++;
      hot[p * 16 + j] = (uint8_t) (std::max_element(c.begin(), c.end()) - c.begin());
    }
  std::vector<long long> hq(N);
  for (size_t d = 0; d < N; ++d) hq[d] = (long long) ((U(rng) - 0.5) * (double) (1ll << 42));
  // gathered list: ascending random 30% subset
  std::vector<uint32_t> hid;
  for (size_t d = 0; d < N; ++d) if (U(rng) < frac) hid.push_back((uint32_t) d);
  const uint32_t NG = (uint32_t) hid.size();
  uint4 *d_p, *d_hot; long long *d_q; unsigned long long *d_sum; uint32_t *d_cnt, *d_ids;
  CK(cudaMalloc(&d_p, h.size())); CK(cudaMalloc(&d_q, N * 8)); CK(cudaMalloc(&d_hot, hot.size()));
  CK(cudaMalloc(&d_ids, hid.size() * 4));
  CK(cudaMalloc(&d_sum, (size_t) P * CELLS * 8)); CK(cudaMalloc(&d_cnt, (size_t) P * CELLS * 4));
  CK(cudaMemcpy(d_p, h.data(), h.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_q, hq.data(), N * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_hot, hot.data(), hot.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_ids, hid.data(), hid.size() * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  std::vector<unsigned long long> ref[2], got((size_t) P * CELLS);
  std::vector<uint32_t> refc[2], gotc((size_t) P * CELLS);

#define SETSMEM(k, bytes) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes))
  auto run = [&](const char *name, int variant, bool gather, bool count, uint32_t slices) {
    const uint32_t n = gather ? NG : (uint32_t) N;
    uint32_t dpb = (n + slices - 1) / slices; dpb = (dpb + 255) & ~255u;
    dim3 grid((n + dpb - 1) / dpb, P);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaMemset(d_sum, 0, (size_t) P * CELLS * 8)); CK(cudaMemset(d_cnt, 0, (size_t) P * CELLS * 4));
      cudaEventRecord(e0);
#define LT(K, WORDS, T) do { SETSMEM(K, CELLS * 4 * (WORDS)); K<<<grid, T, CELLS * 4 * (WORDS)>>>(d_p, N, d_q, d_ids, n, dpb, d_sum, d_cnt, d_hot); } while (0)
#define L(K, WORDS) do { SETSMEM(K, CELLS * 4 * (WORDS)); K<<<grid, 256, CELLS * 4 * (WORDS)>>>(d_p, N, d_q, d_ids, n, dpb, d_sum, d_cnt, d_hot); } while (0)
#define LK(MODE, PIPE, MINB, T, W) do { \
        if (gather) { if (count) LT((k_lean<true, MODE, true, PIPE, MINB, T>), W + 1, T); else LT((k_lean<false, MODE, true, PIPE, MINB, T>), W, T); } \
        else { if (count) LT((k_lean<true, MODE, false, PIPE, MINB, T>), W + 1, T); else LT((k_lean<false, MODE, false, PIPE, MINB, T>), W, T); } } while (0)
      switch (variant) {
        case 0: if (gather) { if (count) L((k_base<true, true>), 3); else L((k_base<false, true>), 2); }
                else { if (count) L((k_base<true, false>), 3); else L((k_base<false, false>), 2); } break;
        case 1: LK(0, 1, 4, 256, 2); break;
        case 2: LK(0, 1, 2, 512, 2); break;
        case 3: LK(0, 1, 1, 1024, 2); break;
        case 4: LK(0, 1, 2, 256, 2); break;
        case 5: LK(0, 1, 1, 512, 2); break;
      }
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    CK(cudaMemcpy(got.data(), d_sum, got.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gotc.data(), d_cnt, gotc.size() * 4, cudaMemcpyDeviceToHost));
    const char *ok;
    if (variant == 0) { ref[gather] = got; if (count) refc[gather] = gotc; ok = "(ref)"; }
    else ok = (got == ref[gather] && (!count || gotc == refc[gather])) ? "OK" : "MISMATCH";
    const double upd = (double) n * 136;
    printf("%-28s %s %s slices=%3u grid=%4u x %d  %8.3f ms  %7.2f Gupd/s %s\n", name, gather ? "gather" : "ident ",
           count ? "cnt  " : "nocnt", slices, grid.x, P, best, upd / best / 1e6, ok);
  };
  const char *names[] = {"V0 production-like", "V1 lean 256x4", "V2 lean 512x2", "V3 lean 1024x1", "V4 lean 256x2", "V5 lean 512x1"};
  for (int gather = 0; gather < 2; ++gather)
    for (int count = 1; count >= 0; --count)
      for (uint32_t slices : {66u, 33u, 16u})
        for (int v = 0; v < 6; ++v) run(names[v], v, gather, count, slices);
  return 0;
}
