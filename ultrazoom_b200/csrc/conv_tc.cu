// conv_tc.cu -- 3x3 / pad 1 / stride 1 convolution as an implicit GEMM on the 5th-gen tensor cores.
//
//   D[128 pixels, N] += A[128 pixels, K = (tap, cin)] * W[N, K]^T        (fp16 | bf16 operands -> fp32 in TMEM)
//
// Replaces InvertedBottleneck.conv1/conv2 (reference model.py:742-748,773-778) and
// SubpixelConv2d.conv (model.py:902-909) with their elementwise successors fused in the epilogue.
//
// Work unit ("patch"): `rows` consecutive image rows x 128 pixels of one image; one fp32
// accumulator (128 TMEM lanes x N columns) per row.  Persistent CTAs walk the patches.
//
// Warp roles (256 threads):
//   warp 0 (one lane)  TMA producer: per K chunk ONE activation halo tile
//                      [(rows+2) image rows][130 pixels][kc channels] (TMA zero-fills the conv padding),
//                      then nine weight tiles [N][kc] -- one per filter tap.
//   warp 1 (one lane)  MMA issuer: for each tap the A operand is the SAME halo tile addressed through a
//                      shared-memory descriptor shifted by (dy*130 + dx) pixel rows, so activations are
//                      read from L2 once per chunk instead of nine times.
//   warp 2             TMEM allocator.
//   warps 4..7         epilogue: tcgen05.ld -> FiLM scale/shift + SiLU -> bf16 | residual add | pixel-shuffle.
//                      With two TMEM stages the epilogue of patch i overlaps the MMAs of patch i+1.
#include "kernels.cuh"

namespace mz {

constexpr int kTileW = 128;
constexpr int kThreads = 256;
constexpr int kMaxSmem = 232448;  // 227 KB opt-in limit per CTA on sm_100

struct TcParams {
  CUtensorMap tmA;  // activations: (cin_p, W, H, B) bf16
  CUtensorMap tmB;  // weights:     (cin_p, n_pad, 9) bf16
  EpiParams epi;
  int kc, n_chunks;
  int rows, acc_stages, acc_stride;
  int halo_mode;
  int a_stages, b_stages;
  int a_stage_bytes, b_stage_bytes;
  int a_tx_bytes, b_tx_bytes;
  int pw;  // pixel rows per image row inside an A stage
  int tiles_x, tiles_y, n_units, n_rounds;
  int cluster;       // CTAs per cluster sharing the weight stream by TMA multicast (1, 2 or 4)
  int b_slice_rows;  // n_pad / cluster: weight rows each CTA loads and multicasts per stage
  uint32_t idesc;
  uint32_t tmem_cols;
};

struct SmemPlan {
  uint32_t a, b, bars, tmem_ptr, total;
};

__host__ __device__ inline SmemPlan plan_smem(const TcParams& p) {
  SmemPlan s;
  s.a = 0;
  s.b = s.a + p.a_stages * p.a_stage_bytes;
  s.bars = s.b + p.b_stages * p.b_stage_bytes;
  const uint32_t nbars = 2 * p.a_stages + 2 * p.b_stages + 4;
  s.tmem_ptr = s.bars + nbars * 8;
  s.total = s.tmem_ptr + 16;
  return s;
}

template <int MODE, int KSTEPS>
__global__ void __launch_bounds__(kThreads, 1) conv_tc_kernel(const __grid_constant__ TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);
  const SmemPlan sp = plan_smem(p);

  const uint32_t a_base = base + sp.a;
  const uint32_t b_base = base + sp.b;
  const uint32_t bar_a_full = base + sp.bars;
  const uint32_t bar_a_empty = bar_a_full + 8 * p.a_stages;
  const uint32_t bar_b_full = bar_a_empty + 8 * p.a_stages;
  const uint32_t bar_b_empty = bar_b_full + 8 * p.b_stages;
  const uint32_t bar_acc_full = bar_b_empty + 8 * p.b_stages;
  const uint32_t bar_acc_empty = bar_acc_full + 16;
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen_base + sp.tmem_ptr);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(bar_a_full + 8 * i, 1);
      mbar_init(bar_a_empty + 8 * i, 1);
    }
    for (int i = 0; i < p.b_stages; ++i) {
      mbar_init(bar_b_full + 8 * i, 1);
      mbar_init(bar_b_empty + 8 * i, p.cluster);  // one tcgen05.commit arrive per CTA sharing the stage
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_acc_full + 8 * i, 1);
      mbar_init(bar_acc_empty + 8 * i, 4);  // one arrive per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(base + sp.tmem_ptr, p.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();  // peers' barriers must be initialised before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  const uint32_t cta_rank = p.cluster > 1 ? cluster_ctarank() : 0u;
  const uint16_t cta_mask = static_cast<uint16_t>((1u << p.cluster) - 1u);

  const int row_bytes = p.kc * 2;
  const int units_per_img = p.tiles_x * p.tiles_y;

  if (warp == 0) {
    // =============================== TMA producer ===============================
    // The whole warp walks the (uniform) loop so that addresses stay in uniform registers; lane 0 issues.
    uint32_t a_it = 0, b_it = 0;
    const uint32_t per_dx = (p.rows + 2) * kTileW * row_bytes;
    for (int round = 0; round < p.n_rounds; ++round) {
      // every CTA of a cluster walks the same number of rounds (the weight stream is shared); a CTA whose unit
      // index runs past the end recomputes the last unit and its epilogue stores nothing
      const int unit = min(round * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x), p.n_units - 1);
      const int b = unit / units_per_img;
      const int rem = unit - b * units_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x0 = tx * kTileW, y0 = ty * p.rows;
      for (int c = 0; c < p.n_chunks; ++c) {
        const uint32_t sa = a_it % p.a_stages, pa = (a_it / p.a_stages) & 1u;
        mbar_wait(bar_a_empty + 8 * sa, pa ^ 1u);
        const uint32_t dstA = a_base + sa * p.a_stage_bytes;
        if (lane == 0) {
          mbar_expect_tx(bar_a_full + 8 * sa, p.a_tx_bytes);
          if (p.halo_mode == 1) {
            for (int dx = 0; dx < 3; ++dx)
              tma_load_4d(dstA + dx * per_dx, &p.tmA, bar_a_full + 8 * sa, c * p.kc, x0 + dx - 1, y0 - 1, b);
          } else {
            tma_load_4d(dstA, &p.tmA, bar_a_full + 8 * sa, c * p.kc, x0 - 1, y0 - 1, b);
          }
        }
        ++a_it;
        for (int tap = 0; tap < 9; ++tap) {
          const uint32_t sb = b_it % p.b_stages, pb = (b_it / p.b_stages) & 1u;
          mbar_wait(bar_b_empty + 8 * sb, pb ^ 1u);
          if (lane == 0) {
            mbar_expect_tx(bar_b_full + 8 * sb, p.b_tx_bytes);
            if (p.cluster > 1)
              tma_load_3d_mcast(b_base + sb * p.b_stage_bytes + cta_rank * p.b_slice_rows * row_bytes, &p.tmB,
                                bar_b_full + 8 * sb, c * p.kc, cta_rank * p.b_slice_rows, tap, cta_mask);
            else
              tma_load_3d(b_base + sb * p.b_stage_bytes, &p.tmB, bar_b_full + 8 * sb, c * p.kc, 0, tap);
          }
          ++b_it;
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer ===============================
    // Uniform loop over the whole warp, lane 0 issues tcgen05.mma / tcgen05.commit.  Descriptors are built once:
    // the high word is constant, the low word (start address >> 4) advances by plain 32-bit adds.
    const uint32_t lt = umma_layout_type(p.kc);
    const uint32_t sbo = 8u * row_bytes;
    const uint32_t desc_hi = static_cast<uint32_t>(umma_smem_desc(0, sbo, lt, 0) >> 32);
    const uint32_t desc_lo0 = static_cast<uint32_t>(umma_smem_desc(0, sbo, lt, 0));
    // tap (dy, dx) and accumulator row r start at dy * DY + dx * DX + r * RP sixteen-byte units into the A stage
    uint32_t DY, DX, RP;
    if (p.halo_mode == 1) {
      DX = ((p.rows + 2) * kTileW * row_bytes) >> 4;
      DY = (kTileW * row_bytes) >> 4;
      RP = DY;
    } else {
      DX = row_bytes >> 4;
      DY = (p.pw * row_bytes) >> 4;
      RP = DY;
    }
    const bool leader = elect_one();  // the same lane issues every tcgen05.mma and tcgen05.commit
    uint32_t a_it = 0, b_it = 0, acc_it = 0;
    for (int round = 0; round < p.n_rounds; ++round) {
      const uint32_t as = acc_it % p.acc_stages, pacc = (acc_it / p.acc_stages) & 1u;
      mbar_wait(bar_acc_empty + 8 * as, pacc ^ 1u);
      tc_fence_after();
      const uint32_t d_base = tmem_base + as * p.rows * p.acc_stride;
      for (int c = 0; c < p.n_chunks; ++c) {
        const uint32_t sa = a_it % p.a_stages, pa = (a_it / p.a_stages) & 1u;
        mbar_wait(bar_a_full + 8 * sa, pa);
        const uint32_t a_lo_stage = desc_lo0 + ((a_base + sa * p.a_stage_bytes) >> 4);
        uint32_t first = c == 0 ? 0u : 1u;  // accumulate flag of the first UMMA of this tap
        for (int dy = 0; dy < 3; ++dy) {
          for (int dx = 0; dx < 3; ++dx) {
            const uint32_t sb = b_it % p.b_stages, pb = (b_it / p.b_stages) & 1u;
            mbar_wait(bar_b_full + 8 * sb, pb);
            tc_fence_after();
            const uint32_t b_lo = desc_lo0 + ((b_base + sb * p.b_stage_bytes) >> 4);
            const uint32_t a_lo_tap = a_lo_stage + dy * DY + dx * DX;
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
              const uint64_t bdesc = (static_cast<uint64_t>(desc_hi) << 32) | (b_lo + 2 * ks);
              uint32_t a_lo = a_lo_tap + 2 * ks;
              uint32_t d = d_base;
              const uint32_t accum = ks == 0 ? first : 1u;
              for (int r = 0; r < p.rows; ++r) {
                const uint64_t adesc = (static_cast<uint64_t>(desc_hi) << 32) | a_lo;
                if (leader) umma_bf16(d, adesc, bdesc, p.idesc, accum);
                a_lo += RP;
                d += p.acc_stride;
              }
            }
            if (leader) {
              if (p.cluster > 1)
                umma_commit_mcast(bar_b_empty + 8 * sb, cta_mask);
              else
                umma_commit(bar_b_empty + 8 * sb);
            }
            first = 1u;
            ++b_it;
          }
        }
        if (leader) umma_commit(bar_a_empty + 8 * sa);
        ++a_it;
      }
      if (leader) umma_commit(bar_acc_full + 8 * as);
      ++acc_it;
    }
    __syncwarp();
  } else if (warp >= 4) {
    // =============================== epilogue ===============================
    const int q = warp - 4;  // TMEM lane quarter this warp may read (== warp % 4)
    uint32_t acc_it = 0;
    for (int round = 0; round < p.n_rounds; ++round) {
      const int unit_raw = round * static_cast<int>(gridDim.x) + static_cast<int>(blockIdx.x);
      const bool unit_ok = unit_raw < p.n_units;  // CTA-uniform
      const int unit = min(unit_raw, p.n_units - 1);
      const int b = unit / units_per_img;
      const int rem = unit - b * units_per_img;
      const int ty = rem / p.tiles_x, tx = rem - ty * p.tiles_x;
      const int x = tx * kTileW + q * 32 + lane;
      const int y0 = ty * p.rows;
      const uint32_t as = acc_it % p.acc_stages, pacc = (acc_it / p.acc_stages) & 1u;
      mbar_wait(bar_acc_full + 8 * as, pacc);
      __syncwarp();
      tc_fence_after();
      for (int r = 0; r < p.rows; ++r) {
        const int y = y0 + r;
        if (y >= p.epi.H || !unit_ok) break;  // warp-uniform
        const bool ok = x < p.epi.W;
        const uint32_t taddr =
            tmem_base + (static_cast<uint32_t>(q * 32) << 16) + (as * p.rows + r) * p.acc_stride;
        if (MODE == 2) {
          float acc[48];
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            if (j * 16 < p.epi.n_pad) {  // warp-uniform
              tmem_ld16(taddr + j * 16, v);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[j * 16 + i] = __uint_as_float(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[j * 16 + i] = 0.f;
            }
          }
          if (ok) epi_head<48>(p.epi, b, y, x, acc);
        } else {
          for (int n0 = 0; n0 < p.epi.n_pad; n0 += 16) {
            uint32_t v[16];
            tmem_ld16(taddr + n0, v);
            tmem_ld_wait();
            if (ok) {
              float acc[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[i] = __uint_as_float(v[i]);
              constexpr int M01 = MODE == 2 ? 0 : MODE;
              epi_store16<M01>(p.epi, b, y, x, n0, acc);
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * as);
      ++acc_it;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.cluster > 1) cluster_sync_all();  // no CTA may exit while a peer can still multicast into it
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
static int pick_kc(int cin_p) { return cin_p % 64 == 0 ? 64 : (cin_p % 32 == 0 ? 32 : 16); }

static uint32_t pow2_cols(uint32_t c) {
  uint32_t v = 32;
  while (v < c) v <<= 1;
  return v;
}

static void fill_geometry(TcParams& p, int cin_p, int kc, int rows, int acc_stages, int halo_mode, int a_stages,
                          int b_stages) {
  p.kc = kc;
  p.n_chunks = cin_p / kc;
  p.rows = rows;
  p.acc_stages = acc_stages;
  p.acc_stride = ((p.epi.n_pad + 31) / 32) * 32;
  p.halo_mode = halo_mode;
  p.a_stages = a_stages;
  p.b_stages = b_stages;
  p.pw = halo_mode == 1 ? kTileW : kTileW + 2;
  const int a_bytes = (halo_mode == 1 ? 3 : 1) * (rows + 2) * p.pw * kc * 2;
  p.a_tx_bytes = a_bytes;
  p.a_stage_bytes = ((a_bytes + 1023) / 1024) * 1024;
  p.b_tx_bytes = p.epi.n_pad * kc * 2;
  p.b_stage_bytes = ((p.b_tx_bytes + 1023) / 1024) * 1024;
  p.tmem_cols = pow2_cols(static_cast<uint32_t>(acc_stages * rows * p.acc_stride));
}

static bool fits(const TcParams& p) {
  return p.acc_stages * p.rows * p.acc_stride <= 512 && plan_smem(p).total + 1024 <= static_cast<uint32_t>(kMaxSmem);
}

int launch_conv_tc(const ConvArgs& a, const ConvTcTune& tune, int device, cudaStream_t s) {
  const EpiParams& e = a.epi;
  MZ_REQUIRE(e.B > 0 && e.H > 0 && e.W > 0, "conv: empty input (B %d, H %d, W %d)", e.B, e.H, e.W);
  MZ_REQUIRE(a.cin_p > 0 && a.cin_p % 16 == 0, "conv: cin_p must be a positive multiple of 16, %d given", a.cin_p);
  MZ_REQUIRE(e.n_pad >= 16 && e.n_pad % 16 == 0 && e.n_pad <= 256,
             "conv: n_pad must be a multiple of 16 in [16, 256], %d given", e.n_pad);
  MZ_REQUIRE(e.mode >= 0 && e.mode <= 2, "conv: bad epilogue mode %d", e.mode);
  MZ_REQUIRE(e.mode != 2 || e.n_pad <= 48, "head conv: n_pad must be <= 48, %d given", e.n_pad);
  MZ_REQUIRE(tune.halo_mode >= 0 && tune.halo_mode <= 1, "conv: bad halo_mode %d", tune.halo_mode);
  MZ_REQUIRE(tune.cluster == 0 || tune.cluster == 1 || tune.cluster == 2 || tune.cluster == 4,
             "conv: cluster must be 0 (auto), 1, 2 or 4, %d given", tune.cluster);
  MZ_REQUIRE(tune.kc == 0 || ((tune.kc == 16 || tune.kc == 32 || tune.kc == 64) && a.cin_p % tune.kc == 0),
             "conv: kc %d does not divide cin_p %d (or is not 16/32/64)", tune.kc, a.cin_p);

  TcParams p;
  memset(&p, 0, sizeof(p));
  p.epi = e;

  // ---- choose the patch geometry: largest patch that keeps two TMEM stages and fits shared memory ----
  const int acc_stride = ((e.n_pad + 31) / 32) * 32;
  bool found = false;
  const int kc_first = tune.kc ? tune.kc : pick_kc(a.cin_p);
  for (int acc_stages = tune.acc_stages ? tune.acc_stages : 2; acc_stages >= 1 && !found; --acc_stages) {
    int rmax = 512 / (acc_stages * acc_stride);
    if (rmax > 4) rmax = 4;
    if (rmax > e.H) rmax = e.H;
    if (tune.rows) rmax = tune.rows;
    for (int rows = rmax; rows >= 1 && !found; --rows) {
      for (int kc = kc_first; kc >= 16 && !found; kc >>= 1) {
        for (int bs = tune.b_stages ? tune.b_stages : 4; bs >= 2 && !found; --bs) {
          fill_geometry(p, a.cin_p, kc, rows, acc_stages, tune.halo_mode, tune.a_stages ? tune.a_stages : 2, bs);
          if (fits(p)) found = true;
          if (tune.b_stages) break;
        }
        if (tune.kc) break;
      }
      if (tune.rows) break;
    }
    if (tune.acc_stages) break;
  }
  if (!found) {
    set_error("conv: no tcgen05 configuration fits (cin_p %d, n_pad %d, rows %d, acc_stages %d, kc %d, halo_mode %d)",
              a.cin_p, e.n_pad, tune.rows, tune.acc_stages, tune.kc, tune.halo_mode);
    return MZ_ERR_UNSUPPORTED;
  }

  // cluster size: share the weight stream between k CTAs when the slices keep whole 8-row swizzle atoms
  {
    int k = tune.cluster ? tune.cluster : 2;
    while (k > 1 && (e.n_pad % k != 0 || (e.n_pad / k) % 8 != 0)) k >>= 1;
    p.cluster = k;
    p.b_slice_rows = e.n_pad / k;
  }
  p.tiles_x = ceil_div(e.W, kTileW);
  p.tiles_y = ceil_div(e.H, p.rows);
  const long long n_units = static_cast<long long>(e.B) * p.tiles_x * p.tiles_y;
  MZ_REQUIRE(n_units < (1LL << 31), "conv: too many patches (%lld)", n_units);
  p.n_units = static_cast<int>(n_units);
  p.idesc = e.bf16 ? umma_idesc_bf16(128, e.n_pad) : umma_idesc_f16(128, e.n_pad);
  const CUtensorMapDataType tdt = e.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;

  const CUtensorMapSwizzle swz =
      p.kc == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.kc == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  {
    const uint64_t dims[4] = {static_cast<uint64_t>(a.cin_p), static_cast<uint64_t>(e.W), static_cast<uint64_t>(e.H),
                              static_cast<uint64_t>(e.B)};
    const uint64_t strides[3] = {static_cast<uint64_t>(a.cin_p) * 2, static_cast<uint64_t>(e.W) * a.cin_p * 2,
                                 static_cast<uint64_t>(e.H) * e.W * a.cin_p * 2};
    const uint32_t box[4] = {static_cast<uint32_t>(p.kc), static_cast<uint32_t>(p.pw),
                             static_cast<uint32_t>(p.rows + 2), 1u};
    int rc = encode_tmap(&p.tmA, tdt, 4, const_cast<uint16_t*>(a.in), dims, strides, box, swz);
    if (rc != MZ_OK) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(a.cin_p), static_cast<uint64_t>(e.n_pad), 9};
    const uint64_t strides[2] = {static_cast<uint64_t>(a.cin_p) * 2, static_cast<uint64_t>(e.n_pad) * a.cin_p * 2};
    const uint32_t box[3] = {static_cast<uint32_t>(p.kc), static_cast<uint32_t>(p.b_slice_rows), 1u};
    int rc = encode_tmap(&p.tmB, tdt, 3, const_cast<uint16_t*>(a.w), dims, strides, box, swz);
    if (rc != MZ_OK) return rc;
  }

  const uint32_t smem = plan_smem(p).total + 1024;
  int sms = sm_count(device);
  if (sms <= 0) sms = 148;
  const int k = p.cluster;
  int grid = p.n_units < sms ? p.n_units : sms;
  if (tune.max_ctas > 0 && grid > tune.max_ctas) grid = tune.max_ctas;
  grid = ceil_div(grid, k) * k;  // whole clusters (surplus CTAs recompute the last patch without storing)
  if (grid > sms) grid = (sms / k) * k;
  p.n_rounds = ceil_div(p.n_units, grid);

  auto launch = [&](auto kern) -> int {
    MZ_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmem));
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = k;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = k > 1 ? 1 : 0;
    MZ_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
    return MZ_OK;
  };
  const int ks = p.kc / 16;
  if (e.mode == 0) return ks == 4 ? launch(conv_tc_kernel<0, 4>) : (ks == 2 ? launch(conv_tc_kernel<0, 2>) : launch(conv_tc_kernel<0, 1>));
  if (e.mode == 1) return ks == 4 ? launch(conv_tc_kernel<1, 4>) : (ks == 2 ? launch(conv_tc_kernel<1, 2>) : launch(conv_tc_kernel<1, 1>));
  return ks == 4 ? launch(conv_tc_kernel<2, 4>) : (ks == 2 ? launch(conv_tc_kernel<2, 2>) : launch(conv_tc_kernel<2, 1>));
}

}  // namespace mz
