// probe.cu -- two small hardware probes (not on the hot path):
//   mz_probe_umma      does a UMMA whose A descriptor starts `row_shift` rows into a TMA-swizzled tile
//                      read the rows the linear address arithmetic says it should?  (This is what lets
//                      the conv kernel serve all nine filter taps from ONE halo tile.)
//   mz_probe_mma_rate  SM cycles per 128 x n x 16 tcgen05.mma when nothing else limits the pipe.
#include <vector>

#include "kernels.cuh"

namespace mz {

constexpr int kProbeRows = 384;
constexpr int kProbeN = 64;

struct ProbeParams {
  CUtensorMap tmA, tmB;
  float* out;  // [128][64]
  int kc, row_shift, bo_mode;
};

__global__ void __launch_bounds__(128, 1) probe_umma_kernel(const __grid_constant__ ProbeParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  const int row_bytes = p.kc * 2;
  const uint32_t a_smem = base;
  const uint32_t b_smem = base + kProbeRows * 128;  // A region sized for kc = 64
  const uint32_t bar_full = b_smem + kProbeN * 128;
  const uint32_t bar_done = bar_full + 8;
  const uint32_t tmem_slot = bar_done + 32;
  volatile uint32_t* tmem_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_done, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 64);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_gen;

  if (threadIdx.x == 0) {
    mbar_expect_tx(bar_full, (kProbeRows + kProbeN) * row_bytes);
    tma_load_2d(a_smem, &p.tmA, bar_full, 0, 0);
    tma_load_2d(a_smem + 192 * row_bytes, &p.tmA, bar_full, 0, 192);
    tma_load_2d(b_smem, &p.tmB, bar_full, 0, 0);
    mbar_wait(bar_full, 0);
    tc_fence_after();
    const uint32_t lt = umma_layout_type(p.kc);
    const uint32_t sbo = 8u * row_bytes;
    const uint32_t idesc = umma_idesc_bf16(128, kProbeN);
    for (int ks = 0; ks < p.kc / 16; ++ks) {
      const uint32_t a_addr = a_smem + p.row_shift * row_bytes + ks * 32;
      const uint32_t bo = p.bo_mode ? ((a_addr >> 7) & 7u) : 0u;
      umma_bf16(tmem_base, umma_smem_desc(a_addr, sbo, lt, bo), umma_smem_desc(b_smem + ks * 32, sbo, lt, 0), idesc,
                ks != 0);
    }
    umma_commit(bar_done);
  }
  mbar_wait(bar_done, 0);
  __syncwarp();
  tc_fence_after();
  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
  for (int n0 = 0; n0 < kProbeN; n0 += 16) {
    uint32_t v[16];
    tmem_ld16(taddr + n0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) p.out[(warp * 32 + lane) * kProbeN + n0 + i] = __uint_as_float(v[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
  }
}

struct RateParams {
  float* out;  // cycles per MMA, one per CTA
  int n, kc, iters, distinct_a, distinct_d, a_row_shift;
  int gap_cycles, commit_in_gap;  // issuer idles gap_cycles (optionally after a tcgen05.commit) between bursts of 8 x iters_per_burst
  int burst_iters;
  uint32_t tmem_cols;
};

__global__ void __launch_bounds__(128, 1) probe_rate_kernel(const RateParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  // layout: A [8 tiles][128 rows][128 B] = 128 KB, B [256 rows][128 B] = 32 KB, barrier, tmem slot
  const uint32_t a_smem = base;
  const uint32_t b_smem = base + 8 * 16384;
  const uint32_t bar_done = b_smem + 32768;
  const uint32_t tmem_slot = bar_done + 32;
  volatile uint32_t* tmem_gen = reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < (8 * 16384 + 32768) / 4; i += blockDim.x)
    reinterpret_cast<uint32_t*>(gen)[i] = 0x3c003c00u;  // bf16 pairs of a small normal value
  fence_proxy_async_smem();
  if (threadIdx.x == 0) {
    mbar_init(bar_done, 1);
    mbar_init(bar_done + 16, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_gen;
  if (warp == 1) {
    // warp-uniform issue loop: 8 fully unrolled UMMAs per iteration, descriptors precomputed
    const int row_bytes = p.kc * 2;
    const uint32_t lt = umma_layout_type(p.kc);
    const uint32_t sbo = 8u * row_bytes;
    const uint32_t idesc = umma_idesc_bf16(128, p.n);
    const uint32_t hi = static_cast<uint32_t>(umma_smem_desc(0, sbo, lt, 0) >> 32);
    const uint32_t lo0 = static_cast<uint32_t>(umma_smem_desc(0, sbo, lt, 0));
    const uint32_t b_lo = lo0 + (b_smem >> 4);
    const uint32_t a_lo = lo0 + ((a_smem + static_cast<uint32_t>(p.a_row_shift * row_bytes)) >> 4);
    const uint32_t a_step = p.distinct_a > 1 ? (16384u >> 4) : 0u;
    const uint32_t d_step = p.distinct_d > 1 ? static_cast<uint32_t>(p.n) : 0u;
    const bool leader = elect_one();
    const uint32_t dynmask = (p.iters & 1) ? 0xffffffffu : static_cast<uint32_t>(p.a_row_shift >> 8);  // runtime 0
    const long long t0 = clock64();
    int until_gap = p.burst_iters;
    for (int it = 0; it < p.iters; ++it) {
      const uint32_t dyn = static_cast<uint32_t>(it) & dynmask;
      if (p.burst_iters > 0 && --until_gap < 0) {  // what the issuer of a real kernel does between two bursts
        until_gap = p.burst_iters - 1;
        if (p.commit_in_gap && leader) umma_commit(bar_done + 16);
        const long long g0 = clock64();
        while (clock64() - g0 < p.gap_cycles) {
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        // dyn (a runtime zero unless iters is odd) keeps the descriptor words from being folded into loop-invariant
        // registers: each UMMA is then preceded by the uniform adds that feed it, as in the convolution's issue loop
        const uint64_t adesc = (static_cast<uint64_t>(hi) << 32) | (a_lo + (j & 1) * a_step + 2 * (j >> 1) + j * dyn);
        const uint64_t bdesc = (static_cast<uint64_t>(hi) << 32) | (b_lo + 2 * (j >> 1) + (j >> 1) * dyn);
        if (leader) umma_acc(tmem_base + (j & 1) * d_step, adesc, bdesc, idesc);
      }
    }
    if (leader) umma_commit(bar_done);
    mbar_wait(bar_done, 0);
    const long long t1 = clock64();
    if (leader) p.out[blockIdx.x] = static_cast<float>(t1 - t0) / static_cast<float>(p.iters * 8);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

}  // namespace mz

using namespace mz;

extern "C" {

int mz_probe_umma(int32_t kc, int32_t row_shift, int32_t base_offset_mode, float* max_abs_err_out) {
  MZ_REQUIRE(kc == 16 || kc == 32 || kc == 64, "probe: kc must be 16, 32 or 64");
  MZ_REQUIRE(row_shift >= 0 && row_shift + 128 <= kProbeRows, "probe: row_shift out of range");
  MZ_REQUIRE(max_abs_err_out, "probe: null output");
  std::vector<__nv_bfloat16> hA(static_cast<size_t>(kProbeRows) * kc), hB(static_cast<size_t>(kProbeN) * kc);
  std::vector<float> fA(hA.size()), fB(hB.size());
  uint32_t s = 12345u;
  auto rnd = [&]() {
    s = s * 1664525u + 1013904223u;
    return static_cast<float>(static_cast<int>((s >> 24) % 17) - 8) * 0.125f;  // exact in bf16
  };
  for (size_t i = 0; i < hA.size(); ++i) {
    fA[i] = rnd();
    hA[i] = __float2bfloat16_rn(fA[i]);
  }
  for (size_t i = 0; i < hB.size(); ++i) {
    fB[i] = rnd();
    hB[i] = __float2bfloat16_rn(fB[i]);
  }
  __nv_bfloat16 *dA = nullptr, *dB = nullptr;
  float* dOut = nullptr;
  MZ_CUDA(cudaMalloc(&dA, hA.size() * 2));
  MZ_CUDA(cudaMalloc(&dB, hB.size() * 2));
  MZ_CUDA(cudaMalloc(&dOut, sizeof(float) * 128 * kProbeN));
  MZ_CUDA(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  MZ_CUDA(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  MZ_CUDA(cudaMemset(dOut, 0xff, sizeof(float) * 128 * kProbeN));
  ProbeParams p;
  memset(&p, 0, sizeof(p));
  p.out = dOut;
  p.kc = kc;
  p.row_shift = row_shift;
  p.bo_mode = base_offset_mode;
  const CUtensorMapSwizzle swz =
      kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  int rc;
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(kc), kProbeRows};
    const uint64_t strides[1] = {static_cast<uint64_t>(kc) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kc), 192};
    rc = encode_tmap(&p.tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, strides, box, swz);
  }
  if (rc == MZ_OK) {
    const uint64_t dims[2] = {static_cast<uint64_t>(kc), kProbeN};
    const uint64_t strides[1] = {static_cast<uint64_t>(kc) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(kc), kProbeN};
    rc = encode_tmap(&p.tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, strides, box, swz);
  }
  if (rc == MZ_OK) {
    const int smem = 1024 + kProbeRows * 128 + kProbeN * 128 + 64;
    cudaError_t e = cudaFuncSetAttribute(probe_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e == cudaSuccess) {
      probe_umma_kernel<<<1, 128, smem>>>(p);
      e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) rc = cuda_fail(e, "probe_umma_kernel", __FILE__, __LINE__);
  }
  if (rc == MZ_OK) {
    std::vector<float> out(128 * kProbeN);
    cudaMemcpy(out.data(), dOut, out.size() * sizeof(float), cudaMemcpyDeviceToHost);
    float worst = 0.f;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < kProbeN; ++n) {
        float ref = 0.f;
        for (int k = 0; k < kc; ++k) ref += fA[static_cast<size_t>(row_shift + m) * kc + k] * fB[static_cast<size_t>(n) * kc + k];
        float d = fabsf(out[m * kProbeN + n] - ref);
        if (!(d == d)) d = 1e30f;  // NaN
        if (d > worst) worst = d;
      }
    *max_abs_err_out = worst;
  }
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dOut);
  return rc;
}

static int g_gap_cycles = 0, g_commit_in_gap = 0, g_burst_iters = 0;
int mz_probe_set_gap(int32_t burst_iters, int32_t gap_cycles, int32_t commit_in_gap) {
  g_burst_iters = burst_iters;
  g_gap_cycles = gap_cycles;
  g_commit_in_gap = commit_in_gap;
  return MZ_OK;
}

int mz_probe_mma_rate(int32_t n, int32_t kc, int32_t iters, int32_t ctas, int32_t distinct_a, int32_t distinct_d,
                      int32_t a_row_shift, float* cycles_per_mma_out) {
  MZ_REQUIRE(a_row_shift >= 0 && a_row_shift <= 64, "probe: a_row_shift must be in [0, 64]");
  MZ_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0, "probe: n must be a multiple of 16 in [16, 256]");
  MZ_REQUIRE(kc == 16 || kc == 32 || kc == 64, "probe: kc must be 16, 32 or 64");
  MZ_REQUIRE(iters > 0 && iters <= (1 << 20) && ctas > 0 && ctas <= 4096, "probe: bad iters/ctas");
  MZ_REQUIRE(distinct_a == 1 || distinct_a == 2 || distinct_a == 4 || distinct_a == 8, "probe: distinct_a must be 1, 2, 4 or 8");
  MZ_REQUIRE(distinct_d >= 1 && distinct_d * n <= 512, "probe: distinct_d * n must fit 512 TMEM columns");
  MZ_REQUIRE(cycles_per_mma_out, "probe: null output");
  float* dOut = nullptr;
  MZ_CUDA(cudaMalloc(&dOut, sizeof(float) * ctas));
  RateParams p;
  p.out = dOut;
  p.n = n;
  p.kc = kc;
  p.iters = iters;
  p.distinct_a = distinct_a;
  p.distinct_d = distinct_d;
  p.a_row_shift = a_row_shift;
  p.gap_cycles = g_gap_cycles;
  p.commit_in_gap = g_commit_in_gap;
  p.burst_iters = g_burst_iters;
  p.tmem_cols = 32;
  while (p.tmem_cols < static_cast<uint32_t>(n * distinct_d)) p.tmem_cols <<= 1;
  const int smem = 1024 + 8 * 16384 + 32768 + 128;
  int rc = MZ_OK;
  cudaError_t e = cudaFuncSetAttribute(probe_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e == cudaSuccess) {
    probe_rate_kernel<<<ctas, 128, smem>>>(p);
    e = cudaGetLastError();
  }
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) rc = cuda_fail(e, "probe_rate_kernel", __FILE__, __LINE__);
  if (rc == MZ_OK) {
    std::vector<float> out(ctas);
    cudaMemcpy(out.data(), dOut, sizeof(float) * ctas, cudaMemcpyDeviceToHost);
    double sum = 0;
    for (float v : out) sum += v;
    *cycles_per_mma_out = static_cast<float>(sum / ctas);
  }
  cudaFree(dOut);
  return rc;
}

}  // extern "C"
