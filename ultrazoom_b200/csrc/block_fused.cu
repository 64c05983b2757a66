// block_fused.cu -- one encoder block as ONE kernel: conv1 (3x3, C -> hC) -> control FiLM -> SiLU -> conv2 (3x3, hC -> C)
// -> ResidualConnection, for the 48-channel models (hC = 96).  Replaces InvertedBottleneck.forward + EncoderBlock's skip
// (reference model.py:773-778, 507-511, 789-792) with the hidden tensor never leaving the SM: SURVEY.md 8(f) rank 2.
//
// Why: run as two kernels, the hidden tensor (2 hC bytes per pixel) is written and read back every block -- 384 of the
// 960 bytes per pixel a 48-channel block moves -- and that round trip is what makes cfg2 HBM-bound.
//
// Shape of the solution
//   * A CTA PAIR (cluster of 2, tcgen05 cta_group::2, M = 256): each CTA owns one 126-pixel-wide column strip of the
//     image and both filter banks are split between the two CTAs along N (41.5 KB + 41.5 KB per CTA instead of
//     83 KB + 83 KB, which do not fit one SM beside any tile).  Every UMMA is issued by the leader CTA for both strips.
//   * The pair walks a SEGMENT of T output rows top to bottom, one image row per step ("rolling" over the rows):
//       conv1(i):  hidden row i = sum over 3 x 3 taps of the three most recent zb rows (a 4-deep ring of TMA-loaded
//                  rows of 130 pixels) -> TMEM accumulator (128 lanes x 96 columns, two stages);
//       epilogue-1 (4 warps): TMEM -> FiLM + SiLU -> 16-bit, zeroed outside the image -> written straight into shared
//                  memory in the K-major swizzled layout conv2's A descriptors read (two buffers);
//       conv2(i):  ONE pass over hidden row i with the three vertical taps STACKED along N (N = 3 x 48 = 144): the
//                  UMMA adds the row's contribution to the accumulators of output rows i-1, i, i+1 at once, which are
//                  adjacent 48-column blocks of a TMEM ring -- so a hidden row is consumed the moment it exists and
//                  never has to be kept (nor recomputed: the only redundant work is 2 hidden rows per segment);
//       epilogue-2 (4 warps): a finished output row -> + fp32 residual tile (TMA load) -> fp32 stream + 16-bit shadow
//                  -> TMA store; zeroes the accumulator block for its next use.
//     The ring has NB = 4 blocks plus 2 overflow blocks, so that the stacked window never wraps (a split window would
//     need differently split filter banks in the two CTAs): rows whose block index is 0 or 1 collect the contributions
//     of the previous lap in the overflow blocks, and epilogue-2 adds the two parts.  The block of an output row is
//     its image row mod 4, so the result does not depend on batch size, segment height or grid size bit for bit.
//   * zb is read with a halo from neighbouring strips and segments, so the block writes its 16-bit output to a SECOND
//     buffer (ping-pong between blocks); the fp32 stream is updated in place (no halo there).
//
// HBM traffic per pixel and block: read zb 2C + zf 4C, write zf 4C + zb 2C = 12C = 576 B at C = 48 (was 960).
#include <stdlib.h>

#include "kernels.cuh"

namespace mz {

namespace fb {
constexpr int kC = 48, kHC = 96;      // channels / hidden channels this kernel is built for
constexpr int kStrip = 126;           // output pixels per strip (a 128-pixel hidden tile minus the conv2 halo)
constexpr int kThreads = 384;         // warps: 0 TMA, 1 MMA issuer, 2 TMEM allocator, 3 spare, 4-7 epilogue-1, 8-11 epilogue-2
constexpr int kZbStages = 4, kHidBufs = 2, kAcc1Stages = 2, kNB = 4, kRingBlocks = kNB + 2;
constexpr uint32_t kW1Tile = 48 * 32;                 // [N/2 = 48 rows][16 ch] 32-byte swizzle
constexpr uint32_t kW1Bytes = 27 * kW1Tile;           // [dy][sub][dx] tiles
constexpr uint32_t kW2Tile = 72 * 64;                 // [144/2 = 72 stacked rows][32 ch] 64-byte swizzle
constexpr uint32_t kW2Bytes = 9 * kW2Tile;            // [dx][chunk] tiles
constexpr uint32_t kZbSub = 4352;                     // 130 px x 32 B = 4160, padded to the 256-byte swizzle period
constexpr uint32_t kZbRow = 3 * kZbSub, kZbRowTx = 3 * 130 * 32;
constexpr uint32_t kHidChunk = 128 * 64, kHidBuf = 3 * kHidChunk;
constexpr uint32_t kStageZ = 3 * 32 * 64, kStageO = 3 * 32 * 32, kStageWarp = kStageZ + kStageO;   // per epilogue-2 warp

constexpr uint32_t oW1 = 0;
constexpr uint32_t oW2 = oW1 + kW1Bytes;
constexpr uint32_t oZb = oW2 + kW2Bytes;                              // 82,944 (1024-aligned)
constexpr uint32_t oHid = oZb + kZbStages * kZbRow;                   // 135,168
constexpr uint32_t oStage = (oHid + kHidBufs * kHidBuf + 256 + 1023) & ~1023u;   // (+256: the dx = 2 tap reads two rows past a tile)
constexpr uint32_t oFilm = oStage + 4 * kStageWarp;
constexpr uint32_t oBars = oFilm + 2 * kHC * 4;
constexpr uint32_t kNumBars = 1 + 2 * kZbStages + 2 * kAcc1Stages + 2 * kHidBufs + 2 * kNB + 12;
constexpr uint32_t oTmemPtr = oBars + kNumBars * 8;
constexpr uint32_t kSmemBytes = oTmemPtr + 16 + 1024;                 // + alignment slack
static_assert(kSmemBytes <= 232448, "fused block: shared memory plan exceeds 227 KB");
constexpr uint32_t kAcc1Col = 0, kRingCol = kAcc1Stages * kHC;        // TMEM columns: conv1 stages, then the conv2 ring
constexpr uint32_t kTmemCols = 512;
static_assert(kRingCol + kRingBlocks * kC <= kTmemCols, "fused block: TMEM plan exceeds 512 columns");
}  // namespace fb

struct FbParams {
  CUtensorMap tmA;     // zb in : (48, W, H, B) 16-bit, box (16, 130, 1, 1), 32-byte swizzle
  CUtensorMap tmW1;    // conv1 : (48, 96, 9)   16-bit, box (16, 48, 3),     32-byte swizzle
  CUtensorMap tmW2;    // conv2 : (96, 144, 3)  16-bit (taps stacked [dy2; dy1; dy0] per dx), box (32, 72, 1), 64-byte swizzle
  CUtensorMap tmZ[2];  // zf     : (48, W, H, B) fp32,   box (16, 32 | 30, 1, 1), 64-byte swizzle
  CUtensorMap tmO[2];  // zb out : (48, W, H, B) 16-bit, box (16, 32 | 30, 1, 1), 32-byte swizzle
  const float* film;   // [B][2][96] (scale row = 1 + gamma, shift row) or nullptr
  unsigned int* sat;
  int B, H, W, bf16;
  int T;               // output rows per segment
  int n_sp, n_seg, n_units;   // strip pairs per row, segments per image, units = B * n_sp * n_seg
  uint32_t idesc1, idesc2;
  int dbg;            // timing experiments only (WRONG results): 1 epilogue-2 drains without residual / stores, 2 epilogue-1 skips
                      // the FiLM + SiLU math, 4 no conv1 UMMAs, 8 no conv2 UMMAs
  long long* prof;    // optional role timers (clock64 ticks), [cluster][16]; nullptr = off
};

#define FB_TIMED(slot, stmt)                         \
  do {                                               \
    if (prof_on) {                                   \
      const long long _t = clock64();                \
      stmt;                                          \
      tick[slot] += clock64() - _t;                  \
    } else {                                         \
      stmt;                                          \
    }                                                \
  } while (0)

__global__ void __launch_bounds__(fb::kThreads, 1) block_fused_kernel(const __grid_constant__ FbParams p) {
  using namespace fb;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen_base = smem_raw + (base - raw);

  // barriers (same offsets in both CTAs of the pair)
  const uint32_t bars = base + oBars;
  const uint32_t bar_w_full = bars;                                   // leader: both CTAs' filter banks have landed
  const uint32_t bar_zb_full = bar_w_full + 8;                        // [4] leader: a zb row of both strips has landed
  const uint32_t bar_zb_empty = bar_zb_full + 8 * kZbStages;          // [4] local : conv1 no longer reads the row
  const uint32_t bar_acc1_full = bar_zb_empty + 8 * kZbStages;        // [2] local : a hidden row's accumulator is complete
  const uint32_t bar_acc1_empty = bar_acc1_full + 8 * kAcc1Stages;    // [2] leader: epilogue-1 of both CTAs has read it
  const uint32_t bar_hid_full = bar_acc1_empty + 8 * kAcc1Stages;     // [2] leader: the hidden row is in shared memory (both CTAs)
  const uint32_t bar_hid_empty = bar_hid_full + 8 * kHidBufs;         // [2] local : conv2 no longer reads the buffer
  const uint32_t bar_acc2_full = bar_hid_empty + 8 * kHidBufs;        // [4] local : an output row's accumulator is complete
  const uint32_t bar_acc2_empty = bar_acc2_full + 8 * kNB;            // [4] leader: epilogue-2 of both CTAs has read + zeroed it
  const uint32_t bar_res = bar_acc2_empty + 8 * kNB;                  // [4][3] local: residual slice j of an epilogue-2 warp
  volatile uint32_t* tmem_ptr_gen = reinterpret_cast<volatile uint32_t*>(gen_base + oTmemPtr);

  const int warp = __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmW1);
    tma_prefetch_desc(&p.tmW2);
    tma_prefetch_desc(&p.tmZ[0]);
    tma_prefetch_desc(&p.tmO[0]);
    mbar_init(bar_w_full, 1);
    for (int i = 0; i < kZbStages; ++i) {
      mbar_init(bar_zb_full + 8 * i, 1);
      mbar_init(bar_zb_empty + 8 * i, 1);
    }
    for (int i = 0; i < kAcc1Stages; ++i) {
      mbar_init(bar_acc1_full + 8 * i, 1);
      mbar_init(bar_acc1_empty + 8 * i, 8);   // four epilogue-1 warps of each CTA
    }
    for (int i = 0; i < kHidBufs; ++i) {
      mbar_init(bar_hid_full + 8 * i, 8);
      mbar_init(bar_hid_empty + 8 * i, 1);
    }
    for (int i = 0; i < kNB; ++i) {
      mbar_init(bar_acc2_full + 8 * i, 1);
      mbar_init(bar_acc2_empty + 8 * i, 8);   // four epilogue-2 warps of each CTA
    }
    for (int i = 0; i < 12; ++i) mbar_init(bar_res + 8 * i, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc2(base + oTmemPtr, kTmemCols);
    tmem_relinquish2();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_gen;
  if (warp >= 8) {  // every conv2 UMMA accumulates: the ring starts from zero (afterwards epilogue-2 re-zeroes what it reads)
    for (uint32_t col = 0; col < kRingBlocks * kC; col += 16)
      tmem_zero16(tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) + kRingCol + col);
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers are initialised and both rings are zero before anything is signalled / issued
  tc_fence_after();
  griddep_launch_dependents();

  const uint32_t cta_rank = cluster_ctarank();
  const int n_clusters = static_cast<int>(gridDim.x >> 1), cluster_id = static_cast<int>(blockIdx.x >> 1);
  const int units_per_img = p.n_sp * p.n_seg;
  const int T = p.T, steps = T + 2, rows_per_unit = T + 4;
  const bool prof_on = p.prof != nullptr;
  long long tick[7] = {0, 0, 0, 0, 0, 0, 0};
  const long long t_role0 = prof_on ? clock64() : 0;

  // unit -> (image, first output row, first output pixel of THIS CTA's strip)
  auto unit_geom = [&](int unit, int& b, int& y0, int& x0) {
    b = unit / units_per_img;
    const int rem = unit - b * units_per_img;
    const int sp = rem / p.n_seg, sg = rem - sp * p.n_seg;
    y0 = sg * T;
    x0 = (2 * sp + static_cast<int>(cta_rank)) * kStrip;
  };

  if (warp == 0) {
    // =============================== TMA producer (each CTA for its own strip) ===============================
    if (lane == 0) {
      // this CTA's halves of the two filter banks; the bytes of both CTAs complete on the leader's barrier
      if (cta_rank == 0) mbar_expect_tx(bar_w_full, 2 * (kW1Bytes + kW2Bytes));
      const uint32_t wfull = mapa_u32(bar_w_full, 0);
      for (int dy = 0; dy < 3; ++dy)
        for (int sub = 0; sub < 3; ++sub)  // box = 16 channels x 48 rows x the three taps dx of filter row dy
          tma2_load_3d(base + oW1 + (dy * 3 + sub) * 3 * kW1Tile, &p.tmW1, wfull, sub * 16, cta_rank * 48, dy * 3);
      for (int dx = 0; dx < 3; ++dx)
        for (int c = 0; c < 3; ++c)
          tma2_load_3d(base + oW2 + (dx * 3 + c) * kW2Tile, &p.tmW2, wfull, c * 32, cta_rank * 72, dx);
    }
    griddep_wait();  // zb of the previous block is valid from here on
    uint32_t cnt = 0;  // zb rows loaded so far (ring position)
    for (int unit = cluster_id; unit < p.n_units; unit += n_clusters) {
      int b, y0, x0;
      unit_geom(unit, b, y0, x0);
      for (int rho = 0; rho < rows_per_unit; ++rho, ++cnt) {
        const uint32_t st = cnt % kZbStages, par = (cnt / kZbStages) & 1u;
        FB_TIMED(0, mbar_wait(bar_zb_empty + 8 * st, par ^ 1u));
        if (lane == 0) {
          if (cta_rank == 0) mbar_expect_tx(bar_zb_full + 8 * st, 2 * kZbRowTx);
          const uint32_t full = mapa_u32(bar_zb_full + 8 * st, 0);
          const uint32_t dst = base + oZb + st * kZbRow;
          for (int sub = 0; sub < 3; ++sub) tma2_load_4d(dst + sub * kZbSub, &p.tmA, full, sub * 16, x0 - 2, y0 - 2 + rho, b);
        }
      }
    }
  } else if (warp == 1 && cta_rank == 0) {
    // =============================== MMA issuer (leader CTA, for both strips) ===============================
    const bool leader = elect_one();
    // descriptor words: conv1 operands use the 32-byte swizzle (16-channel sub-tiles), conv2 operands the 64-byte one
    const uint32_t hi32 = static_cast<uint32_t>(umma_smem_desc(0, 8 * 32, umma_layout_type(16), 0) >> 32);
    const uint32_t hi64 = static_cast<uint32_t>(umma_smem_desc(0, 8 * 64, umma_layout_type(32), 0) >> 32);
    const uint32_t lo0 = static_cast<uint32_t>(umma_smem_desc(0, 8 * 32, umma_layout_type(16), 0));  // (start = 0: LBO bit only)
    auto lo = [&](uint32_t addr) { return lo0 + ((addr & 0x3FFFFu) >> 4); };
    mbar_wait(bar_w_full, 0);
    tc_fence_after();
    uint32_t zcnt = 0;    // ring position of zb row rho = 0 of the current unit
    uint32_t a1 = 0;      // conv1 accumulator uses so far
    uint32_t hcnt = 0;    // hidden rows consumed so far
    uint32_t pos = 0;     // conv2 ring position of output row rho' = 0 of the current unit
    uint32_t opened = 0;  // ring positions < opened have been claimed from epilogue-2
    for (int unit = cluster_id; unit < p.n_units; unit += n_clusters) {
      {
        // Ring block of an output row = image row mod 4, whatever the batch, segment or round: rows in blocks 0 / 1 sum
        // two partial accumulators (overflow blocks), so tying the block to the image row makes the fp32 association --
        // hence every bit of the result -- independent of how the work was cut.  Skipped positions are handed to
        // epilogue-2 as junk rows.
        int b, y0, x0;
        unit_geom(unit, b, y0, x0);
        const uint32_t lead = (static_cast<uint32_t>(y0 + 2) - pos) & 3u;
        for (uint32_t s2 = 0; s2 < lead; ++s2, ++pos) {
          while (opened <= pos) {
            FB_TIMED(3, mbar_wait(bar_acc2_empty + 8 * (opened % kNB), ((opened / kNB) & 1u) ^ 1u));
            ++opened;
          }
          if (leader) umma2_commit_mcast(bar_acc2_full + 8 * (pos % kNB), 3);
          __syncwarp();
        }
      }
      for (int i = 0; i <= steps; ++i) {
        if (i < steps) {
          // ---- conv1(i): hidden row i from zb rows rho = i, i+1, i+2 ----
          for (int k = (i == 0 ? 0 : 2); k < 3; ++k) {
            const uint32_t c = zcnt + i + k;
            FB_TIMED(0, mbar_wait(bar_zb_full + 8 * (c % kZbStages), (c / kZbStages) & 1u));
          }
          const uint32_t st1 = a1 % kAcc1Stages;
          FB_TIMED(1, mbar_wait(bar_acc1_empty + 8 * st1, ((a1 / kAcc1Stages) & 1u) ^ 1u));
          tc_fence_after();
          const uint32_t d1 = tmem_base + kAcc1Col + st1 * kHC;
          if (leader) {
            if (!(p.dbg & 4))
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              const uint32_t arow = base + oZb + ((zcnt + i + dy) % kZbStages) * kZbRow;
#pragma unroll
              for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
                for (int sub = 0; sub < 3; ++sub) {
                  const uint64_t ad = (static_cast<uint64_t>(hi32) << 32) | lo(arow + sub * kZbSub + dx * 32);
                  const uint64_t bd = (static_cast<uint64_t>(hi32) << 32) | lo(base + oW1 + ((dy * 3 + sub) * 3 + dx) * kW1Tile);
                  if (dy == 0 && dx == 0 && sub == 0)
                    umma2_bf16(d1, ad, bd, p.idesc1, 0u);
                  else
                    umma2_acc(d1, ad, bd, p.idesc1);
                }
              }
            }
            umma2_commit_mcast(bar_acc1_full + 8 * st1, 3);
            umma2_commit_mcast(bar_zb_empty + 8 * ((zcnt + i) % kZbStages), 3);  // the oldest of the three rows is done
            if (i == steps - 1) {  // the last step also releases the two rows nobody will read again
              umma2_commit_mcast(bar_zb_empty + 8 * ((zcnt + i + 1) % kZbStages), 3);
              umma2_commit_mcast(bar_zb_empty + 8 * ((zcnt + i + 2) % kZbStages), 3);
            }
          }
          __syncwarp();
          ++a1;
        }
        if (i >= 1) {
          // ---- conv2(j): hidden row j adds to output rows rho' = j, j+1, j+2 (ring positions pos + j ...) ----
          const int j = i - 1;
          const uint32_t hb = hcnt % kHidBufs;
          FB_TIMED(2, mbar_wait(bar_hid_full + 8 * hb, (hcnt / kHidBufs) & 1u));
          while (opened <= pos + j + 2) {  // claim the blocks this window opens (drained + zeroed by epilogue-2)
            FB_TIMED(3, mbar_wait(bar_acc2_empty + 8 * (opened % kNB), ((opened / kNB) & 1u) ^ 1u));
            ++opened;
          }
          tc_fence_after();
          const uint32_t d2 = tmem_base + kRingCol + ((pos + j) % kNB) * kC;
          if (leader) {
            if (!(p.dbg & 8))
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
              for (int c = 0; c < 3; ++c) {
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                  const uint64_t ad = (static_cast<uint64_t>(hi64) << 32) | lo(base + oHid + hb * kHidBuf + c * kHidChunk + dx * 64 + ks * 32);
                  const uint64_t bd = (static_cast<uint64_t>(hi64) << 32) | lo(base + oW2 + (dx * 3 + c) * kW2Tile + ks * 32);
                  umma2_acc(d2, ad, bd, p.idesc2);
                }
              }
            }
            umma2_commit_mcast(bar_hid_empty + 8 * hb, 3);
            umma2_commit_mcast(bar_acc2_full + 8 * ((pos + j) % kNB), 3);  // output row rho' = j has all its contributions
            if (j == steps - 1) {  // end of the segment: the two rows below it only ever receive junk -- hand them over too
              umma2_commit_mcast(bar_acc2_full + 8 * ((pos + j + 1) % kNB), 3);
              umma2_commit_mcast(bar_acc2_full + 8 * ((pos + j + 2) % kNB), 3);
            }
          }
          __syncwarp();
          ++hcnt;
        }
      }
      zcnt += rows_per_unit;
      pos += rows_per_unit;
    }
  } else if (warp >= 4 && warp < 8) {
    // =============================== epilogue-1: conv1 accumulator -> FiLM + SiLU -> hidden row in shared memory ===============================
    griddep_wait();  // the FiLM table is written by an earlier kernel of the stream
    const int q = warp & 3;
    float* film_s = reinterpret_cast<float*>(gen_base + oFilm);
    int film_b = -1;
    float amax = 0.f;
    uint32_t a1 = 0;
    const uint32_t px_row = q * 32 + lane;  // row of the hidden tile this thread writes: hidden pixel x0 - 1 + px_row
    for (int unit = cluster_id; unit < p.n_units; unit += n_clusters) {
      int b, y0, x0;
      unit_geom(unit, b, y0, x0);
      if (b != film_b) {
        named_bar_sync(1, 128);  // nobody still reads the previous image's rows
        for (int i = threadIdx.x - 128; i < 2 * kHC; i += 128) {
          float v = i < kHC ? 1.f : 0.f;
          if (p.film != nullptr) v = __ldg(p.film + static_cast<size_t>(b) * 2 * kHC + i);
          film_s[i] = 0.5f * v;  // SiLU's v / 2 folded into scale and shift (exact): see silu_h
        }
        named_bar_sync(1, 128);
        film_b = b;
      }
      const int hx = x0 - 1 + static_cast<int>(px_row);
      const bool px_ok = hx >= 0 && hx < p.W;
      for (int i = 0; i < steps; ++i, ++a1) {
        const uint32_t st1 = a1 % kAcc1Stages, hb = a1 % kHidBufs;
        const int hy = y0 - 1 + i;
        const bool keep = px_ok && hy >= 0 && hy < p.H;  // conv2 zero-pads the HIDDEN tensor: outside the image it is 0, not SiLU(shift)
        FB_TIMED(0, mbar_wait(bar_acc1_full + 8 * st1, (a1 / kAcc1Stages) & 1u));
        FB_TIMED(1, mbar_wait(bar_hid_empty + 8 * hb, ((a1 / kHidBufs) & 1u) ^ 1u));
        __syncwarp();
        tc_fence_after();
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kAcc1Col + st1 * kHC;
        const uint32_t hrow = base + oHid + hb * kHidBuf + px_row * 64;
#pragma unroll
        for (int c = 0; c < 3; ++c) {  // 32 channels = one 64-byte row of chunk tile c
          uint32_t v[32];
          FB_TIMED(2, { tmem_ld32(taddr + c * 32, v); tmem_ld_wait(); });
          const long long t_math = prof_on ? clock64() : 0;
          const float4* sc = reinterpret_cast<const float4*>(film_s + c * 32);
          const float4* sh = reinterpret_cast<const float4*>(film_s + kHC + c * 32);
          uint32_t o[16];
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const float4 g = sc[kk], h = sh[kk];
            float a0 = silu_h(fmaf(__uint_as_float(v[4 * kk + 0]), g.x, h.x));
            float a1v = silu_h(fmaf(__uint_as_float(v[4 * kk + 1]), g.y, h.y));
            float a2 = silu_h(fmaf(__uint_as_float(v[4 * kk + 2]), g.z, h.z));
            float a3 = silu_h(fmaf(__uint_as_float(v[4 * kk + 3]), g.w, h.w));
            if (!keep) a0 = a1v = a2 = a3 = 0.f;
            if (p.dbg & 2) a0 = a1v = a2 = a3 = __uint_as_float(v[4 * kk]);
            amax = fmaxf(fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1v))), fmaxf(fabsf(a2), fabsf(a3)));
            o[2 * kk] = pack_op2(p.bf16, a0, a1v);
            o[2 * kk + 1] = pack_op2(p.bf16, a2, a3);
          }
          const uint32_t crow = hrow + c * kHidChunk;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)
            sts128(crow + swz_chunk(px_row, ch, 64) * 16, o[4 * ch], o[4 * ch + 1], o[4 * ch + 2], o[4 * ch + 3]);
          if (prof_on) tick[3] += clock64() - t_math;
        }
        FB_TIMED(4, {
          tc_fence_before();
          fence_proxy_async_smem();  // the UMMAs (async proxy) read what these threads wrote
          __syncwarp();
        });
        FB_TIMED(5, {
          if (lane == 0) {
            mbar_arrive_remote(mapa_u32(bar_acc1_empty + 8 * st1, 0));
            mbar_arrive_remote(mapa_u32(bar_hid_full + 8 * hb, 0));
          }
        });
      }
    }
    if (!p.bf16 && p.sat != nullptr && !(amax <= MZ_F16_MAX)) *p.sat = 1u;
  } else if (warp >= 8) {
    // =============================== epilogue-2: finished output row -> + residual -> zf, zb -> TMA store ===============================
    // Per row: read the accumulator block (plus its overflow part) into registers, zero it and hand it back to the
    // issuer AT ONCE; then the row is finished in three 16-channel slices, each with its own staging slot (fp32 box +
    // 16-bit box) and residual barrier: slice j of the NEXT row is requested as soon as slice j's stores of this row
    // have been read (one slice later: bulk_wait_read<1>), so residual loads run a row ahead and neither their latency
    // nor the store read-out is on the row's critical path.
    griddep_wait();  // the residual stream is written by earlier kernels of the stream
    const int q = warp & 3;
    const uint32_t st_z = base + oStage + q * kStageWarp, st_o = st_z + kStageZ;
    const uint32_t my_res = bar_res + 8 * (q * 3);  // one barrier per slice slot
    const CUtensorMap* tmZ = &p.tmZ[q == 3 ? 1 : 0];  // the last warp's box is 30 pixels: a strip is 126 wide
    const CUtensorMap* tmO = &p.tmO[q == 3 ? 1 : 0];
    const uint32_t slice_bytes = (q == 3 ? 30u : 32u) * 16 * 4;
    float amax = 0.f;
    const bool lane_live = !(q == 3 && lane >= 30);  // the last two pixel rows of a 128-pixel tile belong to the next strip
    uint32_t pos = 0, rpar = 0;
    bool preloaded = false;    // slices 0 and 1 of the current row's residual were requested during the previous row ...
    bool pending_last = false;  // ... and slice 2 still has to be (its slot was being stored from until now)
    for (int unit = cluster_id; unit < p.n_units; unit += n_clusters) {
      int b, y0, x0;
      unit_geom(unit, b, y0, x0);
      const int xw = x0 + q * 32;  // first output pixel of this warp
      const int lead = static_cast<int>((static_cast<uint32_t>(y0 + 2) - pos) & 3u);  // junk positions: see the issuer
      for (int rho = -lead; rho < rows_per_unit; ++rho, ++pos) {
        const int y = y0 - 2 + rho;
        const bool store = rho >= 2 && rho < T + 2 && y < p.H && xw < p.W && !(p.dbg & 1);  // (warp-uniform)
        const uint32_t k = pos % kNB, par = (pos / kNB) & 1u;
        if (store && lane == 0) {
          if (!preloaded) {  // first stored row of a segment: nothing was requested ahead
            FB_TIMED(2, bulk_wait_read<0>());
            for (int j = 0; j < 3; ++j) {
              mbar_expect_tx(my_res + 8 * j, slice_bytes);
              tma_load_4d(st_z + j * 2048, tmZ, my_res + 8 * j, j * 16, xw, y, b);
            }
          } else if (pending_last) {
            FB_TIMED(2, bulk_wait_read<0>());
            mbar_expect_tx(my_res + 16, slice_bytes);
            tma_load_4d(st_z + 2 * 2048, tmZ, my_res + 16, 32, xw, y, b);
          }
        }
        FB_TIMED(0, mbar_wait(bar_acc2_full + 8 * k, par));
        __syncwarp();
        tc_fence_after();
        const long long t_body = prof_on ? clock64() : 0;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + kRingCol;
        const uint32_t t_main = lane_base + k * kC, t_ovf = lane_base + (kNB + k) * kC;
        const bool has_ovf = k < 2;  // rows in blocks 0 / 1 collected the previous lap's contributions in the overflow blocks
        uint32_t v[48];
        if (store) {
          uint32_t(&v0)[32] = *reinterpret_cast<uint32_t(*)[32]>(&v[0]);
          uint32_t(&v1)[16] = *reinterpret_cast<uint32_t(*)[16]>(&v[32]);
          tmem_ld32(t_main, v0);
          tmem_ld16(t_main + 32, v1);
          if (has_ovf) {
            uint32_t w0[32], w1[16];
            tmem_ld32(t_ovf, w0);
            tmem_ld16(t_ovf + 32, w1);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __uint_as_float(w0[e]));
#pragma unroll
            for (int e = 0; e < 16; ++e) v[32 + e] = __float_as_uint(__uint_as_float(v[32 + e]) + __uint_as_float(w1[e]));
          } else {
            tmem_ld_wait();
          }
        }
#pragma unroll
        for (int n0 = 0; n0 < kC; n0 += 16) {
          tmem_zero16(t_main + n0);
          if (has_ovf) tmem_zero16(t_ovf + n0);
        }
        tmem_st_wait();  // the zeroes are in TMEM before the issuer may accumulate into the block again
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(mapa_u32(bar_acc2_empty + 8 * k, 0));
        if (prof_on) tick[3] += clock64() - t_body;
        if (store) {
          const bool next_store = rho + 1 < T + 2 && y + 1 < p.H;  // the next position stores too (same strip, next row)
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            FB_TIMED(1, mbar_wait(my_res + 8 * j, rpar));
            const uint32_t zrow = st_z + j * 2048 + lane * 64;
            uint32_t o[8], addr[4];
            float4 z[4];
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // (all four loads first: the accessors are volatile, i.e. ordered)
              addr[kk] = zrow + swz_chunk(lane, kk, 64) * 16;
              z[kk] = lds128f(addr[kk]);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
              z[kk].x += __uint_as_float(v[16 * j + 4 * kk + 0]);
              z[kk].y += __uint_as_float(v[16 * j + 4 * kk + 1]);
              z[kk].z += __uint_as_float(v[16 * j + 4 * kk + 2]);
              z[kk].w += __uint_as_float(v[16 * j + 4 * kk + 3]);
              sts128(addr[kk], __float_as_uint(z[kk].x), __float_as_uint(z[kk].y), __float_as_uint(z[kk].z), __float_as_uint(z[kk].w));
              if (lane_live) amax = fmaxf(fmaxf(amax, fmaxf(fabsf(z[kk].x), fabsf(z[kk].y))), fmaxf(fabsf(z[kk].z), fabsf(z[kk].w)));
              o[2 * kk] = pack_op2(p.bf16, z[kk].x, z[kk].y);
              o[2 * kk + 1] = pack_op2(p.bf16, z[kk].z, z[kk].w);
            }
            const uint32_t orow = st_o + j * 1024 + lane * 32;
            sts128(orow + swz_chunk(lane, 0, 32) * 16, o[0], o[1], o[2], o[3]);
            sts128(orow + swz_chunk(lane, 1, 32) * 16, o[4], o[5], o[6], o[7]);
            FB_TIMED(4, {
              fence_proxy_async_smem();
              __syncwarp();
            });
            const long long t_tail = prof_on ? clock64() : 0;
            if (lane == 0) {
              tma_store_4d(tmZ, st_z + j * 2048, j * 16, xw, y, b);
              tma_store_4d(tmO, st_o + j * 1024, j * 16, xw, y, b);
              bulk_commit();
              if (j >= 1 && next_store) {  // slot j-1: its stores (one group back) have been read -> next row's slice
                bulk_wait_read<1>();
                mbar_expect_tx(my_res + 8 * (j - 1), slice_bytes);
                tma_load_4d(st_z + (j - 1) * 2048, tmZ, my_res + 8 * (j - 1), (j - 1) * 16, xw, y + 1, b);
              }
            }
            if (prof_on) tick[5] += clock64() - t_tail;
          }
          rpar ^= 1u;
          preloaded = next_store;
          pending_last = next_store;
        } else {
          preloaded = pending_last = false;
        }
      }
    }
    if (lane == 0) bulk_wait_all();
    if (!p.bf16 && p.sat != nullptr && !(amax <= MZ_F16_MAX)) *p.sat = 1u;
  }

  if (prof_on && lane == 0 && cta_rank == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 8)) {
    const int role = warp == 0 ? 0 : (warp == 1 ? 1 : (warp == 4 ? 2 : 3));
    long long* dst = p.prof + (static_cast<size_t>(cluster_id) * 4 + role) * 8;
    for (int i = 0; i < 7; ++i) dst[i] = tick[i];
    dst[7] = clock64() - t_role0;
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA may exit while its peer can still signal into it or read its operands
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, fb::kTmemCols);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
bool fused_block_applies(int Cp, int hCp, int zb_pitch) { return Cp == fb::kC && hCp == fb::kHC && zb_pitch == fb::kC; }

struct FusedLaunchImpl {
  FbParams p;
  int grid;
};
static_assert(sizeof(FusedLaunchImpl) <= sizeof(ConvLaunch::storage), "ConvLaunch::storage too small for the fused block");

// Stack conv2's packed bank [tap = dy*3+dx][48][96] into [dx][(2-dy)*48 + n][96]: per filter column dx the three vertical
// taps as one 144-row B matrix whose column blocks are, in ascending order, the output rows i-1 (dy = 2), i (dy = 1),
// i+1 (dy = 0) a hidden row i feeds.  Device-to-device, nine contiguous copies.
int stack_conv2_bank(const uint16_t* packed, uint16_t* stacked, cudaStream_t s) {
  const size_t tap = static_cast<size_t>(fb::kC) * fb::kHC;
  for (int dy = 0; dy < 3; ++dy)
    for (int dx = 0; dx < 3; ++dx)
      MZ_CUDA(cudaMemcpyAsync(stacked + (static_cast<size_t>(dx) * 3 + (2 - dy)) * tap, packed + (static_cast<size_t>(dy) * 3 + dx) * tap,
                              tap * sizeof(uint16_t), cudaMemcpyDeviceToDevice, s));
  return MZ_OK;
}

int prepare_block_fused(const FusedBlockArgs& a, int device, ConvLaunch* out) {
  out->valid = false;
  FusedLaunchImpl& L = *reinterpret_cast<FusedLaunchImpl*>(out->storage);
  MZ_REQUIRE(a.B > 0 && a.H > 0 && a.W > 0, "fused block: empty input (B %d, H %d, W %d)", a.B, a.H, a.W);
  MZ_REQUIRE(a.zb_in && a.zb_out && a.zf && a.w1 && a.w2s, "fused block: null pointer");
  MZ_REQUIRE(a.zb_in != a.zb_out, "fused block: the 16-bit output needs its own buffer (the input is read with a halo)");
  FbParams& p = L.p;
  memset(&p, 0, sizeof(p));
  p.film = a.film;
  p.sat = a.sat;
  p.B = a.B;
  p.H = a.H;
  p.W = a.W;
  p.bf16 = a.bf16;
  int sms = sm_count(device);
  if (sms <= 0) sms = 148;
  const int n_clusters_max = sms / 2;
  const int n_strips = ceil_div(a.W, fb::kStrip);
  p.n_sp = ceil_div(n_strips, 2);
  // segment height: the T that minimises rounds x (T + 2) steps over the persistent clusters (2 redundant hidden rows per
  // segment against the quantisation of units over 74 clusters)
  long long best_cost = -1;
  int best_T = a.H;
  const int t_forced = a.seg_rows > 0 ? (a.seg_rows < a.H ? a.seg_rows : a.H) : 0;
  const int t_min = t_forced ? t_forced : (a.H < 8 ? a.H : 8), t_max = t_forced ? t_forced : a.H;
  for (int T = t_min; T <= t_max; ++T) {
    const long long units = static_cast<long long>(a.B) * p.n_sp * ceil_div(a.H, T);
    const long long clusters = units < n_clusters_max ? units : n_clusters_max;
    const long long cost = ceil_div(static_cast<int>(units), static_cast<int>(clusters)) * static_cast<long long>(T + 2);
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best_T = T;
    }
  }
  p.T = best_T;
  p.n_seg = ceil_div(a.H, p.T);
  const long long n_units = static_cast<long long>(a.B) * p.n_sp * p.n_seg;
  MZ_REQUIRE(n_units < (1LL << 30), "fused block: too many units (%lld)", n_units);
  p.n_units = static_cast<int>(n_units);
  int clusters = p.n_units < n_clusters_max ? p.n_units : n_clusters_max;
  if (a.max_ctas > 0 && clusters > a.max_ctas / 2) clusters = a.max_ctas / 2 > 0 ? a.max_ctas / 2 : 1;
  L.grid = 2 * clusters;
  p.idesc1 = a.bf16 ? umma_idesc_bf16(256, fb::kHC) : umma_idesc_f16(256, fb::kHC);
  p.idesc2 = a.bf16 ? umma_idesc_bf16(256, 3 * fb::kC) : umma_idesc_f16(256, 3 * fb::kC);
  const CUtensorMapDataType tdt = a.bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const uint64_t W = a.W, H = a.H, B = a.B, C = fb::kC;
  int rc;
  {
    const uint64_t dims[4] = {C, W, H, B};
    const uint64_t st[3] = {C * 2, W * C * 2, H * W * C * 2};
    const uint32_t box[4] = {16u, 130u, 1u, 1u};
    if ((rc = encode_tmap(&p.tmA, tdt, 4, const_cast<uint16_t*>(a.zb_in), dims, st, box, CU_TENSOR_MAP_SWIZZLE_32B)) != MZ_OK) return rc;
    for (int v = 0; v < 2; ++v) {
      const uint32_t bo[4] = {16u, v ? 30u : 32u, 1u, 1u};
      if ((rc = encode_tmap(&p.tmO[v], tdt, 4, a.zb_out, dims, st, bo, CU_TENSOR_MAP_SWIZZLE_32B)) != MZ_OK) return rc;
      const uint64_t st32[3] = {C * 4, W * C * 4, H * W * C * 4};
      if ((rc = encode_tmap(&p.tmZ[v], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a.zf, dims, st32, bo, CU_TENSOR_MAP_SWIZZLE_64B)) != MZ_OK) return rc;
    }
  }
  {
    const uint64_t dims[3] = {C, static_cast<uint64_t>(fb::kHC), 9};
    const uint64_t st[2] = {C * 2, static_cast<uint64_t>(fb::kHC) * C * 2};
    const uint32_t box[3] = {16u, 48u, 3u};
    if ((rc = encode_tmap(&p.tmW1, tdt, 3, const_cast<uint16_t*>(a.w1), dims, st, box, CU_TENSOR_MAP_SWIZZLE_32B)) != MZ_OK) return rc;
  }
  {
    const uint64_t K = fb::kHC, N = 3 * fb::kC;
    const uint64_t dims[3] = {K, N, 3};
    const uint64_t st[2] = {K * 2, N * K * 2};
    const uint32_t box[3] = {32u, 72u, 1u};
    if ((rc = encode_tmap(&p.tmW2, tdt, 3, const_cast<uint16_t*>(a.w2s), dims, st, box, CU_TENSOR_MAP_SWIZZLE_64B)) != MZ_OK) return rc;
  }
  {
    const char* d = getenv("MZ_FB_DBG");
    p.dbg = d ? atoi(d) : 0;
  }
  MZ_CUDA(cudaFuncSetAttribute(block_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(fb::kSmemBytes)));
  {
    static const bool verbose = getenv("MZ_VERBOSE") != nullptr;
    if (verbose)
      fprintf(stderr, "[mz fused block] B %d H %d W %d | T %d segs %d strip pairs %d units %d | grid %d (clusters of 2) smem %u\n",
              a.B, a.H, a.W, p.T, p.n_seg, p.n_sp, p.n_units, L.grid, fb::kSmemBytes);
  }
  out->valid = true;
  return MZ_OK;
}

int run_block_fused(ConvLaunch& launch, cudaStream_t s) {
  MZ_REQUIRE(launch.valid, "fused block: launch was not prepared");
  FusedLaunchImpl& L = *reinterpret_cast<FusedLaunchImpl*>(launch.storage);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(L.grid);
  cfg.blockDim = dim3(fb::kThreads);
  cfg.dynamicSmemBytes = fb::kSmemBytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  attr[na].id = cudaLaunchAttributeClusterDimension;
  attr[na].val.clusterDim.x = 2;
  attr[na].val.clusterDim.y = 1;
  attr[na].val.clusterDim.z = 1;
  ++na;
  static const bool no_pdl = getenv("MZ_NO_PDL") != nullptr;
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  FbParams q = L.p;
  static const bool prof = getenv("MZ_FB_PROF") != nullptr;
  static long long* g_prof = nullptr;
  static int prof_left = 3;  // print the first few launches only
  const bool do_prof = prof && prof_left > 0;
  if (do_prof) {
    if (!g_prof) MZ_CUDA(cudaMalloc(&g_prof, sizeof(long long) * 128 * 32));
    MZ_CUDA(cudaMemsetAsync(g_prof, 0, sizeof(long long) * 128 * 32, s));
    q.prof = g_prof;
    cfg.numAttrs = 1;  // (no dependent launch while timing)
  }
  void* args[1] = {&q};
  MZ_CUDA(cudaLaunchKernelExC(&cfg, reinterpret_cast<const void*>(block_fused_kernel), args));
  if (do_prof) {
    --prof_left;
    MZ_CUDA(cudaStreamSynchronize(s));
    const int nc = L.grid / 2;
    std::vector<long long> h(static_cast<size_t>(nc) * 32);
    MZ_CUDA(cudaMemcpy(h.data(), g_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    double m[32] = {0};
    for (int c = 0; c < nc; ++c)
      for (int i = 0; i < 32; ++i) m[i] += static_cast<double>(h[static_cast<size_t>(c) * 32 + i]) / nc;
    fprintf(stderr,
            "[mz fused prof] T %d units %d clusters %d | producer: wait_zb_empty %.0f total %.0f | issuer: wait_zb_full %.0f "
            "wait_acc1_empty %.0f wait_hid_full %.0f wait_acc2_empty %.0f total %.0f | epi1: wait_acc1_full %.0f wait_hid_empty "
            "%.0f tmem_ld %.0f math+sts %.0f fences %.0f arrives %.0f total %.0f | epi2: wait_acc2_full %.0f wait_residual %.0f "
            "wait_store_read %.0f body(incl. residual wait) %.0f st_wait+fences %.0f arrive+stores %.0f total %.0f\n",
            L.p.T, L.p.n_units, nc, m[0], m[7], m[8], m[9], m[10], m[11], m[15], m[16], m[17], m[18], m[19], m[20], m[21], m[23],
            m[24], m[25], m[26], m[27], m[28], m[29], m[31]);
  }
  return MZ_OK;
}

}  // namespace mz
