// small_kernels.cu -- the bandwidth-bound kernels around the conv stack:
//   * bicubic zoom            (Upsample(mode="bicubic"),  reference model.py:71,156)
//   * stem + layout change    (FanOutProjection,          reference model.py:212-242)
//   * FiLM coefficient table  (control module, restated -- SURVEY.md Appendix C)
//   * SIMT direct 3x3 conv    (diagnostic twin of the tcgen05 kernel; same epilogues)
#include <math.h>

#include "kernels.cuh"

namespace mz {

// ----------------------------------------------------------------------------------------------
// bicubic phase table (host) -- follows ATen's upsample_bicubic2d arithmetic in fp32:
//   src = scale * (dst + 0.5) - 0.5 with scale = 1/r, idx = floor(src), t = src - idx,
//   w = {cc2(t+1), cc1(t), cc1(1-t), cc2(2-t)}, A = -0.75.
// ----------------------------------------------------------------------------------------------
static float cc1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
static float cc2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

void make_bicubic_table(int r, BicubicTable* t) {
  const float A = -0.75f;
  const float scale = 1.0f / static_cast<float>(r);
  t->r = r;
  for (int p = 0; p < 4; ++p) {
    t->off[p] = 0;
    for (int k = 0; k < 4; ++k) t->w[p][k] = 0.f;
  }
  for (int p = 0; p < r && p < 4; ++p) {
    const float src = scale * (static_cast<float>(p) + 0.5f) - 0.5f;
    const float fl = floorf(src);
    const float tt = src - fl;
    t->off[p] = static_cast<int>(fl);
    t->w[p][0] = cc2(tt + 1.f, A);
    t->w[p][1] = cc1(tt, A);
    const float x2 = 1.f - tt;
    t->w[p][2] = cc1(x2, A);
    t->w[p][3] = cc2(x2 + 1.f, A);
  }
}

// ----------------------------------------------------------------------------------------------
// bicubic zoom: one thread -> VEC consecutive HR pixels of one HR row of one plane.
// Reads hit L1/L2 (each LR pixel is reused r*r*16/… times); the kernel is bound by the HR write.
// ----------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) bicubic_kernel(const float* __restrict__ x, float* __restrict__ y, int planes,
                                                      int H, int W, BicubicTable bt) {
  // thread -> one LR column `lx` of one HR row: produces R consecutive HR pixels [lx*R, lx*R+R)
  const int WR = W * R, HR = H * R;
  const long long total = static_cast<long long>(planes) * HR * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int lx = static_cast<int>(idx % W);
    const long long t = idx / W;
    const int oy = static_cast<int>(t % HR);
    const int pl = static_cast<int>(t / HR);
    const float* plane = x + static_cast<size_t>(pl) * H * W;
    const int py = oy % R;
    const int by = oy / R + bt.off[py] - 1;
    // the R phases need LR columns lx-2 .. lx+2
    float v[5];
#pragma unroll
    for (int m = 0; m < 5; ++m) v[m] = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int yy = min(max(by + k, 0), H - 1);
      const float* row = plane + static_cast<size_t>(yy) * W;
      const float wy = bt.w[py][k];
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        const int xx = min(max(lx - 2 + m, 0), W - 1);
        v[m] = fmaf(__ldg(row + xx), wy, v[m]);
      }
    }
    float o[R];
#pragma unroll
    for (int p = 0; p < R; ++p) {
      // taps start at lx + off[p] - 1  ->  v index (off[p] + 1) .. (off[p] + 4)
      const int s = bt.off[p] + 1;
      float a = 0.f;
#pragma unroll
      for (int m = 0; m < 4; ++m) a = fmaf(s == 0 ? v[m] : v[m + 1], bt.w[p][m], a);
      o[p] = a;
    }
    float* dst = y + (static_cast<size_t>(pl) * HR + oy) * WR + static_cast<size_t>(lx) * R;
    if (R == 4) {
      *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    } else if (R == 2) {
      *reinterpret_cast<float2*>(dst) = make_float2(o[0], o[1]);
    } else {
#pragma unroll
      for (int p = 0; p < R; ++p) dst[p] = o[p];
    }
  }
}

int launch_bicubic(const float* x, float* y, int planes, int H, int W, int r, cudaStream_t s) {
  MZ_REQUIRE(r == 2 || r == 3 || r == 4, "Upscale ratio must be either 2, 3, or 4, %d given.", r);
  MZ_REQUIRE(planes > 0 && H > 0 && W > 0, "bicubic: empty input (planes %d, H %d, W %d)", planes, H, W);
  BicubicTable bt;
  make_bicubic_table(r, &bt);
  const long long total = static_cast<long long>(planes) * H * r * W;
  const int block = 256;
  long long blocks = (total + block - 1) / block;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  if (r == 2)
    bicubic_kernel<2><<<static_cast<unsigned>(blocks), block, 0, s>>>(x, y, planes, H, W, bt);
  else if (r == 3)
    bicubic_kernel<3><<<static_cast<unsigned>(blocks), block, 0, s>>>(x, y, planes, H, W, bt);
  else
    bicubic_kernel<4><<<static_cast<unsigned>(blocks), block, 0, s>>>(x, y, planes, H, W, bt);
  MZ_CUDA(cudaGetLastError());
  return MZ_OK;
}

// ----------------------------------------------------------------------------------------------
// stem: x (B,3,H,W) fp32 NCHW -> zf (B,H,W,Cp) fp32 and zb (B,H,W,Cp) bf16; z = W x + bias,
// channels >= C are written as zero (w/bias are zero-padded to Cp by the caller).
// One thread -> 8 channels of one pixel (16-byte bf16 store, 2 x 16-byte fp32 stores).
// ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                   const float* __restrict__ bias, float* __restrict__ zf,
                                                   uint16_t* __restrict__ zb, int bf16, int B, int H, int W, int Cp,
                                                   int Cz) {
  const int groups = Cz / 8;
  const long long npix = static_cast<long long>(B) * H * W;
  const long long total = npix * groups;
  const size_t plane = static_cast<size_t>(H) * W;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    const long long pix = idx / groups;
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (g * 8 < Cp) {  // channels beyond Cp exist only in the 16-bit shadow (zero padding up to its pitch)
      const long long b = pix / static_cast<long long>(plane);
      const size_t sp = static_cast<size_t>(pix - b * static_cast<long long>(plane));
      const float* xb = x + static_cast<size_t>(b) * 3 * plane + sp;
      const float r0 = __ldg(xb), r1 = __ldg(xb + plane), r2 = __ldg(xb + 2 * plane);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = g * 8 + i;
        o[i] = fmaf(__ldg(w + n * 3 + 2), r2, fmaf(__ldg(w + n * 3 + 1), r1, fmaf(__ldg(w + n * 3), r0, __ldg(bias + n))));
      }
      if (zf != nullptr) {
        float4* f = reinterpret_cast<float4*>(zf + static_cast<size_t>(pix) * Cp + g * 8);
        f[0] = make_float4(o[0], o[1], o[2], o[3]);
        f[1] = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
    if (zf == nullptr) {  // split stream: z16 = [hi | lo], pitch 2 * Cp (Cz == Cp here)
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) split_op2(bf16, o[2 * i], o[2 * i + 1], hi[i], lo[i]);
      uint16_t* dst = zb + static_cast<size_t>(pix) * 2 * Cp + g * 8;
      st_global_v4(dst, hi[0], hi[1], hi[2], hi[3]);
      st_global_v4(dst + Cp, lo[0], lo[1], lo[2], lo[3]);
      continue;
    }
    st_global_v4(zb + static_cast<size_t>(pix) * Cz + g * 8, pack_op2(bf16, o[0], o[1]), pack_op2(bf16, o[2], o[3]),
                 pack_op2(bf16, o[4], o[5]), pack_op2(bf16, o[6], o[7]));
  }
}

int launch_stem(const float* x, const float* w, const float* bias, float* zf, uint16_t* zb, int bf16, int B, int H,
                int W, int Cp, int zb_pitch, cudaStream_t s) {
  MZ_REQUIRE(Cp > 0 && Cp % 8 == 0, "stem: padded channel count must be a multiple of 8, %d given", Cp);
  MZ_REQUIRE(B > 0 && H > 0 && W > 0, "stem: empty input");
  const int Cz = (zb_pitch && zf != nullptr) ? zb_pitch : Cp;
  MZ_REQUIRE(Cz >= Cp && Cz % 8 == 0, "stem: zb pitch %d must be a multiple of 8 and >= %d", Cz, Cp);
  const long long total = static_cast<long long>(B) * H * W * (Cz / 8);
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  stem_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(x, w, bias, zf, zb, bf16, B, H, W, Cp, Cz);
  MZ_CUDA(cudaGetLastError());
  return MZ_OK;
}

// ----------------------------------------------------------------------------------------------
// FiLM table: film[l][b][0][n] = 1 + gamma, film[l][b][1][n] = beta, g = Linear_l(c_b);
// gamma = g[n], beta = g[hC + n]; padded channels n >= hC get scale 1, shift 0.
// ----------------------------------------------------------------------------------------------
__global__ void film_kernel(const float* __restrict__ c, int c_rows, const float* __restrict__ w,
                            const float* __restrict__ bias, float* __restrict__ film, int L, int B, int F, int hC,
                            int hCp) {
  const long long total = static_cast<long long>(L) * B * hCp;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int n = static_cast<int>(idx % hCp);
    const long long t = idx / hCp;
    const int b = static_cast<int>(t % B);
    const int l = static_cast<int>(t / B);
    float scale = 1.f, shift = 0.f;
    if (n < hC) {
      const float* cb = c + static_cast<size_t>(c_rows == 1 ? 0 : b) * F;
      const float* wg = w + (static_cast<size_t>(l) * 2 * hC + n) * F;
      const float* wb = w + (static_cast<size_t>(l) * 2 * hC + hC + n) * F;
      float g = bias[static_cast<size_t>(l) * 2 * hC + n];
      float be = bias[static_cast<size_t>(l) * 2 * hC + hC + n];
      for (int f = 0; f < F; ++f) {
        g = fmaf(cb[f], wg[f], g);
        be = fmaf(cb[f], wb[f], be);
      }
      scale = 1.f + g;
      shift = be;
    }
    float* dst = film + (static_cast<size_t>(l) * B + b) * 2 * hCp;
    dst[n] = scale;
    dst[hCp + n] = shift;
  }
}

int launch_film(const float* c, int c_rows, const float* w, const float* b, float* film, int L, int B, int F, int hC,
                int hCp, cudaStream_t s) {
  MZ_REQUIRE(c_rows == 1 || c_rows == B, "Batch size of c (%d) must match x (%d).", c_rows, B);
  MZ_REQUIRE(L > 0 && B > 0 && F > 0 && hC > 0 && hCp >= hC, "film: bad shape");
  const long long total = static_cast<long long>(L) * B * hCp;
  long long blocks = (total + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  film_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(c, c_rows, w, b, film, L, B, F, hC, hCp);
  MZ_CUDA(cudaGetLastError());
  return MZ_OK;
}

// ----------------------------------------------------------------------------------------------
// SIMT direct convolution (diagnostic).  Same operands (16-bit NHWC activations, 16-bit [tap][n][k]
// weights), fp32 accumulation, same epilogues -- so that a tcgen05 result can be bisected against
// it on the GPU.  One thread -> 16 output channels of one pixel (modes 0/1) or the whole head.
// ----------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(128) conv_simt_kernel(ConvArgs a) {
  const EpiParams& p = a.epi;
  const int groups = MODE == 2 ? 1 : p.n_pad / 16;
  const long long total = static_cast<long long>(p.B) * p.H * p.W * groups;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(idx % groups);
    long long t = idx / groups;
    const int x = static_cast<int>(t % p.W);
    t /= p.W;
    const int y = static_cast<int>(t % p.H);
    const int b = static_cast<int>(t / p.H);
    constexpr int NACC = MODE == 2 ? 48 : 16;
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.f;
    const int n0 = g * 16;
    const int nlim = MODE == 2 ? p.n_pad : 16;
    for (int ky = 0; ky < 3; ++ky) {
      const int yy = y + ky - 1;
      if (yy < 0 || yy >= p.H) continue;
      for (int kx = 0; kx < 3; ++kx) {
        const int xx = x + kx - 1;
        if (xx < 0 || xx >= p.W) continue;
        const uint16_t* src = a.in + ((static_cast<size_t>(b) * p.H + yy) * p.W + xx) * (a.in_pitch ? a.in_pitch : a.cin_p);
        const uint16_t* wt = a.w + (static_cast<size_t>(ky * 3 + kx) * p.n_pad + n0) * a.cin_p;
        for (int k = 0; k < a.cin_p; ++k) {
          const float v = op_to_float(p.bf16, src[k]);
#pragma unroll
          for (int i = 0; i < NACC; ++i)
            if (i < nlim) acc[i] = fmaf(v, op_to_float(p.bf16, wt[static_cast<size_t>(i) * a.cin_p + k]), acc[i]);
        }
      }
    }
    if (MODE == 2) {
      float h[48];
#pragma unroll
      for (int i = 0; i < 48; ++i) h[i] = i < NACC ? acc[i] : 0.f;
      epi_head<48>(p, b, y, x, h);
    } else {
      float h[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) h[i] = acc[i];
      constexpr int M01 = MODE == 2 ? 0 : MODE;  // (epilogue mode 0, 1 or 3)
      epi_store16<M01>(p, b, y, x, n0, h,
                       p.film != nullptr ? p.film + static_cast<size_t>(b) * 2 * p.n_pad : nullptr);
    }
  }
}

int launch_conv_simt(const ConvArgs& a, cudaStream_t s) {
  const EpiParams& p = a.epi;
  MZ_REQUIRE(p.n_pad % 16 == 0 && p.n_pad > 0, "conv: n_pad must be a positive multiple of 16, %d given", p.n_pad);
  MZ_REQUIRE(p.mode != 2 || p.n_pad <= 48, "head conv: n_pad must be <= 48, %d given", p.n_pad);
  const long long total = static_cast<long long>(p.B) * p.H * p.W * (p.mode == 2 ? 1 : p.n_pad / 16);
  long long blocks = (total + 127) / 128;
  if (blocks > 148LL * 64) blocks = 148LL * 64;
  const unsigned gb = static_cast<unsigned>(blocks);
  if (p.mode == 0)
    conv_simt_kernel<0><<<gb, 128, 0, s>>>(a);
  else if (p.mode == 1)
    conv_simt_kernel<1><<<gb, 128, 0, s>>>(a);
  else if (p.mode == 3)
    conv_simt_kernel<3><<<gb, 128, 0, s>>>(a);
  else
    conv_simt_kernel<2><<<gb, 128, 0, s>>>(a);
  MZ_CUDA(cudaGetLastError());
  return MZ_OK;
}

}  // namespace mz
