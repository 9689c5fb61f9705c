// Development micro-benchmark: shared-memory histogram accumulation variants on B200.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o hist_mb scripts/hist_microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <random>
#include <cmath>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int FPP = 16;
constexpr int B = 256;           // bins per feature
constexpr int CELLS = FPP * B;   // per panel

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

// V0: stream only
__global__ void __launch_bounds__(256) k_stream(const uint4 *panels, size_t N, const long long *lamq, uint32_t n, uint32_t dpb, unsigned long long *out) {
  const uint4 *prow = panels + (size_t) blockIdx.y * N;
  uint32_t begin = blockIdx.x * dpb, end = min(n, begin + dpb);
  unsigned long long acc = 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += 256) {
    uint4 r = prow[i];
    acc += r.x + r.y + r.z + r.w + (unsigned long long) lamq[i];
  }
  if (acc == 0x1234567) out[0] = acc;
}

// V1: CAS64 sum + 32-bit count
template <bool ROT, bool COUNT>
__global__ void __launch_bounds__(256) k_cas64(const uint4 *panels, size_t N, const long long *lamq, uint32_t n, uint32_t dpb, unsigned long long *gsum, uint32_t *gcnt) {
  extern __shared__ unsigned char sm[];
  unsigned long long *s_sum = (unsigned long long *) sm;
  uint32_t *s_cnt = (uint32_t *) (s_sum + CELLS);
  for (int i = threadIdx.x; i < CELLS; i += 256) { s_sum[i] = 0; s_cnt[i] = 0; }
  __syncthreads();
  const uint4 *prow = panels + (size_t) blockIdx.y * N;
  uint32_t begin = blockIdx.x * dpb, end = min(n, begin + dpb);
  const uint32_t rot = ROT ? (threadIdx.x & 15) : 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += 256) {
    uint4 row = rotate_bytes(prow[i], rot);
    unsigned long long q = (unsigned long long) lamq[i];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t slot = (j + rot) & 15;
      uint32_t cell = slot * B + ext(row, j);
      atomicAdd(s_sum + cell, q);
      if (COUNT) atomicAdd(s_cnt + cell, 1u);
    }
  }
  __syncthreads();
  unsigned long long *gs = gsum + (size_t) blockIdx.y * CELLS;
  uint32_t *gc = gcnt + (size_t) blockIdx.y * CELLS;
  for (int i = threadIdx.x; i < CELLS; i += 256) {
    if (s_sum[i]) atomicAdd(gs + i, s_sum[i]);
    if (COUNT && s_cnt[i]) atomicAdd(gc + i, s_cnt[i]);
  }
}

// V2: two 32-bit limbs (lo with carry detection) + separate count
template <bool ROT, bool COUNT>
__global__ void __launch_bounds__(256) k_limb2(const uint4 *panels, size_t N, const long long *lamq, uint32_t n, uint32_t dpb, unsigned long long *gsum, uint32_t *gcnt) {
  extern __shared__ unsigned char sm[];
  uint32_t *s_lo = (uint32_t *) sm;
  int32_t *s_hi = (int32_t *) (s_lo + CELLS);
  uint32_t *s_cnt = (uint32_t *) (s_hi + CELLS);
  for (int i = threadIdx.x; i < CELLS; i += 256) { s_lo[i] = 0; s_hi[i] = 0; s_cnt[i] = 0; }
  __syncthreads();
  const uint4 *prow = panels + (size_t) blockIdx.y * N;
  uint32_t begin = blockIdx.x * dpb, end = min(n, begin + dpb);
  const uint32_t rot = ROT ? (threadIdx.x & 15) : 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += 256) {
    uint4 row = rotate_bytes(prow[i], rot);
    long long q = lamq[i];
    const uint32_t qlo = (uint32_t) q;
    const int32_t qhi = (int32_t) (q >> 32);
    uint32_t cells[16], olds[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t slot = (j + rot) & 15;
      cells[j] = slot * B + ext(row, j);
      olds[j] = atomicAdd(s_lo + cells[j], qlo);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int32_t carry = (olds[j] + qlo) < olds[j];
      atomicAdd(s_hi + cells[j], qhi + carry);
      if (COUNT) atomicAdd(s_cnt + cells[j], 1u);
    }
  }
  __syncthreads();
  unsigned long long *gs = gsum + (size_t) blockIdx.y * CELLS;
  uint32_t *gc = gcnt + (size_t) blockIdx.y * CELLS;
  for (int i = threadIdx.x; i < CELLS; i += 256) {
    long long v = ((long long) s_hi[i] << 32) + (long long) s_lo[i];
    if (v) atomicAdd(gs + i, (unsigned long long) v);
    if (COUNT && s_cnt[i]) atomicAdd(gc + i, s_cnt[i]);
  }
}

// V3: two limbs, count packed into the hi word (block-local: dpb <= 4096, |q| < 2^37)
template <bool ROT>
__global__ void __launch_bounds__(256) k_limb2_packed(const uint4 *panels, size_t N, const long long *lamq, uint32_t n, uint32_t dpb, unsigned long long *gsum, uint32_t *gcnt) {
  extern __shared__ unsigned char sm[];
  uint32_t *s_lo = (uint32_t *) sm;
  int32_t *s_hi = (int32_t *) (s_lo + CELLS);
  for (int i = threadIdx.x; i < CELLS; i += 256) { s_lo[i] = 0; s_hi[i] = 0; }
  __syncthreads();
  const uint4 *prow = panels + (size_t) blockIdx.y * N;
  uint32_t begin = blockIdx.x * dpb, end = min(n, begin + dpb);
  const uint32_t rot = ROT ? (threadIdx.x & 15) : 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += 256) {
    uint4 row = rotate_bytes(prow[i], rot);
    long long q = lamq[i];
    const uint32_t qlo = (uint32_t) q;
    const int32_t qhi = (int32_t) (q >> 32);
    uint32_t cells[16], olds[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t slot = (j + rot) & 15;
      cells[j] = slot * B + ext(row, j);
      olds[j] = atomicAdd(s_lo + cells[j], qlo);
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int32_t carry = (olds[j] + qlo) < olds[j];
      atomicAdd(s_hi + cells[j], (qhi + carry) * 8192 + 1);
    }
  }
  __syncthreads();
  unsigned long long *gs = gsum + (size_t) blockIdx.y * CELLS;
  uint32_t *gc = gcnt + (size_t) blockIdx.y * CELLS;
  for (int i = threadIdx.x; i < CELLS; i += 256) {
    const int32_t w = s_hi[i];
    const uint32_t cnt = (uint32_t) w & 8191u;
    const int32_t hi = (w - (int32_t) cnt) >> 13;
    long long v = ((long long) hi << 32) + (long long) s_lo[i];
    if (cnt) { atomicAdd(gs + i, (unsigned long long) v); atomicAdd(gc + i, cnt); }
  }
}

// V4: one warp per (feature pair?) -- per-thread-private bins are too big; instead: 32-bit count only
template <bool ROT>
__global__ void __launch_bounds__(256) k_count_only(const uint4 *panels, size_t N, uint32_t n, uint32_t dpb, uint32_t *gcnt) {
  extern __shared__ unsigned char sm[];
  uint32_t *s_cnt = (uint32_t *) sm;
  for (int i = threadIdx.x; i < CELLS; i += 256) s_cnt[i] = 0;
  __syncthreads();
  const uint4 *prow = panels + (size_t) blockIdx.y * N;
  uint32_t begin = blockIdx.x * dpb, end = min(n, begin + dpb);
  const uint32_t rot = ROT ? (threadIdx.x & 15) : 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += 256) {
    uint4 row = rotate_bytes(prow[i], rot);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t slot = (j + rot) & 15;
      atomicAdd(s_cnt + slot * B + ext(row, j), 1u);
    }
  }
  __syncthreads();
  uint32_t *gc = gcnt + (size_t) blockIdx.y * CELLS;
  for (int i = threadIdx.x; i < CELLS; i += 256) if (s_cnt[i]) atomicAdd(gc + i, s_cnt[i]);
}

// V5: float32 x2 (hi/lo float split, "double-float" accumulate) -- native fp32 smem atomics
template <bool ROT>
__global__ void __launch_bounds__(256) k_f32(const uint4 *panels, size_t N, const long long *lamq, uint32_t n, uint32_t dpb, float *gout) {
  extern __shared__ unsigned char sm[];
  float *s = (float *) sm;
  for (int i = threadIdx.x; i < CELLS; i += 256) s[i] = 0.f;
  __syncthreads();
  const uint4 *prow = panels + (size_t) blockIdx.y * N;
  uint32_t begin = blockIdx.x * dpb, end = min(n, begin + dpb);
  const uint32_t rot = ROT ? (threadIdx.x & 15) : 0;
  for (uint32_t i = begin + threadIdx.x; i < end; i += 256) {
    uint4 row = rotate_bytes(prow[i], rot);
    float q = (float) lamq[i];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      uint32_t slot = (j + rot) & 15;
      atomicAdd(s + slot * B + ext(row, j), q);
    }
  }
  __syncthreads();
  float *g = gout + (size_t) blockIdx.y * CELLS;
  for (int i = threadIdx.x; i < CELLS; i += 256) if (s[i] != 0.f) atomicAdd(g + i, s[i]);
}

int main(int argc, char **argv) {
  const size_t N = argc > 1 ? atol(argv[1]) : 1000000;
  const int P = 9;
  std::mt19937_64 rng(1);
  std::vector<uint8_t> h((size_t) P * N * 16);
  std::uniform_real_distribution<double> U(0, 1);
  for (int p = 0; p < P; ++p)
    for (size_t d = 0; d < N; ++d)
      for (int j = 0; j < 16; ++j) {
        int f = p * 16 + j;
        double u = U(rng);
        double v = pow(u, 1 + f % 3);
        uint8_t b = (uint8_t) llround(255 * v);
        if (f % 20 == 19) b = 0;
        h[((size_t) p * N + d) * 16 + j] = b;
      }
  std::vector<long long> hq(N);
  for (size_t d = 0; d < N; ++d) hq[d] = (long long) ((U(rng) - 0.5) * (double) (1ll << 37));
  uint4 *d_p; long long *d_q; unsigned long long *d_sum; uint32_t *d_cnt; float *d_f;
  CK(cudaMalloc(&d_p, h.size())); CK(cudaMalloc(&d_q, N * 8));
  CK(cudaMalloc(&d_sum, (size_t) P * CELLS * 8)); CK(cudaMalloc(&d_cnt, (size_t) P * CELLS * 4)); CK(cudaMalloc(&d_f, (size_t) P * CELLS * 4));
  CK(cudaMemcpy(d_p, h.data(), h.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_q, hq.data(), N * 8, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  // reference result from V1
  std::vector<unsigned long long> ref((size_t) P * CELLS), got((size_t) P * CELLS);
  std::vector<uint32_t> refc((size_t) P * CELLS), gotc((size_t) P * CELLS);

  auto run = [&](const char *name, int variant, uint32_t dpb, bool check) {
    dim3 grid((unsigned) ((N + dpb - 1) / dpb), P);
    float best = 1e9;
    for (int rep = 0; rep < 6; ++rep) {
      CK(cudaMemset(d_sum, 0, (size_t) P * CELLS * 8)); CK(cudaMemset(d_cnt, 0, (size_t) P * CELLS * 4)); CK(cudaMemset(d_f, 0, (size_t) P * CELLS * 4));
      cudaEventRecord(e0);
      switch (variant) {
        case 0: k_stream<<<grid, 256>>>(d_p, N, d_q, (uint32_t) N, dpb, d_sum); break;
        case 1: k_cas64<true, true><<<grid, 256, CELLS * 12>>>(d_p, N, d_q, (uint32_t) N, dpb, d_sum, d_cnt); break;
        case 2: k_cas64<false, true><<<grid, 256, CELLS * 12>>>(d_p, N, d_q, (uint32_t) N, dpb, d_sum, d_cnt); break;
        case 3: k_cas64<true, false><<<grid, 256, CELLS * 12>>>(d_p, N, d_q, (uint32_t) N, dpb, d_sum, d_cnt); break;
        case 4: k_limb2<true, true><<<grid, 256, CELLS * 12>>>(d_p, N, d_q, (uint32_t) N, dpb, d_sum, d_cnt); break;
        case 5: k_limb2<false, true><<<grid, 256, CELLS * 12>>>(d_p, N, d_q, (uint32_t) N, dpb, d_sum, d_cnt); break;
        case 6: k_limb2<true, false><<<grid, 256, CELLS * 12>>>(d_p, N, d_q, (uint32_t) N, dpb, d_sum, d_cnt); break;
        case 7: k_limb2_packed<true><<<grid, 256, CELLS * 8>>>(d_p, N, d_q, (uint32_t) N, dpb, d_sum, d_cnt); break;
        case 8: k_limb2_packed<false><<<grid, 256, CELLS * 8>>>(d_p, N, d_q, (uint32_t) N, dpb, d_sum, d_cnt); break;
        case 9: k_count_only<true><<<grid, 256, CELLS * 4>>>(d_p, N, (uint32_t) N, dpb, d_cnt); break;
        case 10: k_count_only<false><<<grid, 256, CELLS * 4>>>(d_p, N, (uint32_t) N, dpb, d_cnt); break;
        case 11: k_f32<true><<<grid, 256, CELLS * 4>>>(d_p, N, d_q, (uint32_t) N, dpb, d_f); break;
      }
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    const char *ok = "";
    if (check) {
      CK(cudaMemcpy(got.data(), d_sum, got.size() * 8, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(gotc.data(), d_cnt, gotc.size() * 4, cudaMemcpyDeviceToHost));
      if (variant == 1) { ref = got; refc = gotc; ok = "(ref)"; }
      else ok = (got == ref && (gotc == refc || variant == 3 || variant == 6)) ? "OK" : "MISMATCH";
    }
    double upd = (double) N * P * 16;
    printf("%-34s dpb=%6u grid=%5u x %d  %8.3f ms  %7.2f Gupd/s  %7.1f GB/s(bins) %s\n", name, dpb, grid.x, P, best,
           upd / best / 1e6, (double) N * P * 16 / best / 1e6, ok);
  };
  for (uint32_t dpb : {4096u, 16384u}) {
    run("V0 stream only", 0, dpb, false);
    run("V1 cas64+cnt rot", 1, dpb, true);
    run("V1 cas64+cnt norot", 2, dpb, true);
    run("V1 cas64 nocnt rot", 3, dpb, true);
    run("V2 limb2+cnt rot", 4, dpb, true);
    run("V2 limb2+cnt norot", 5, dpb, true);
    run("V2 limb2 nocnt rot", 6, dpb, true);
    if (dpb <= 4096) { run("V3 limb2 packed rot", 7, dpb, true); run("V3 limb2 packed norot", 8, dpb, true); }
    run("V4 count only rot", 9, dpb, false);
    run("V4 count only norot", 10, dpb, false);
    run("V5 f32 atomics rot", 11, dpb, false);
  }
  return 0;
}
