// unet_ops.cu -- the operators of the reference's 0.3.0 U-Net that the flat (0.2.x) path does not have (SURVEY.md 8(f)
// rank 3), on NHWC fp32 feature maps, fp32 arithmetic throughout:
//   * AdaptiveResidualMix  (reference model.py:795-839): beta = sigmoid(conv1x1([x ; z])), w = sigmoid(alpha) * beta,
//                           out = (1 - w) x + w z                                   -> gather-GEMM, K = 2C, fused gate epilogue
//   * PixelCrush           (model.py:842-882): f x f / stride f convolution, f in {2, 3, 4}, no bias
//                                                                                   -> gather-GEMM, K = f f Cin
//   * QualityAssessor      (model.py:1004-1032): conv3x3 (pad 1, bias) -> global average pool -> (B, F)
//                                                                                   -> gather-GEMM, K = 9 C, pooled epilogue
//   * mid-network SubpixelConv2d's PixelShuffle on NHWC (model.py:911,928; Decoder :569-571): pure re-indexing
//   * Decoder.crop_feature_maps (model.py:650-689): centre crop / zero pad to a target size
// The 3x3 convolutions of these blocks (InvertedBottleneck, SubpixelConv2d.conv) run on the tcgen05 kernel of conv_tc.cu.
// One tiled SIMT GEMM serves the three gather forms in exact fp32: M = output pixels, N = output channels, K as above;
// 64 x 64 x 16 tiles, 256 threads, a 4 x 4 register tile per thread.  For the mix and the crush it is the exact TWIN
// (math = MZ_MATH_FP32) of the tcgen05 kind::tf32 kernel of unet_tc.cu, which is the default (MZ_MATH_TF32) and runs at
// the HBM roofline; the quality head (a 3x3 convolution reduced to B x F numbers) only exists in this form.
#include "kernels.cuh"

namespace mz {

struct GemmArgs {
  const float* a0;   // x | input feature map (B,H,W,pitch_in)
  const float* a1;   // z (mix) or nullptr
  const float* wt;   // weights as [N][K] fp32 (conv.weight as the reference stores it)
  const float* bias; // [N] or nullptr (pool epilogue adds it once per image)
  float* out;        // (M, pitch_out) fp32 | (B, N) pooled
  uint16_t* out16;   // optional 16-bit shadow of out (pitch_out) or nullptr
  int bf16;
  int M, N, K;
  int B, H, W, C;    // input geometry (C = input channels)
  int Ho, Wo, f;     // output geometry / crush factor
  int pitch_in, pitch_out;
  float gate;        // sigmoid(alpha) (mix)
  float inv_hw;      // 1 / (H W) (pool)
};

constexpr int kBM = 64, kBN = 64, kBK = 16;

// AMODE: 0 mix ([x ; z] along K), 1 crush (f x f patch gather), 2 conv3x3 pad 1.  EMODE: 0 gated mix, 1 store, 2 pooled sum.
template <int AMODE, int EMODE>
__global__ void __launch_bounds__(256) gather_gemm_kernel(const GemmArgs g) {
  __shared__ float As[kBK][kBM + 4];
  __shared__ float Bs[kBK][kBN + 4];
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * kBM, n0 = blockIdx.y * kBN;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each a 4 x 4 tile: rows ty*4.., columns tx*4..
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // loader mapping: A tile 64 x 16 = 1024 elements, 4 per thread: row = tid / 4, k = (tid % 4) * 4 .. +3
  const int lr = tid >> 2, lk = (tid & 3) * 4;
  const int m = m0 + lr;
  int pb = 0, py = 0, px = 0;
  if (m < g.M) {
    const int hw = g.Ho * g.Wo;
    pb = m / hw;
    const int rem = m - pb * hw;
    py = rem / g.Wo;
    px = rem - py * g.Wo;
  }
  auto load_a = [&](int k) -> float {
    if (m >= g.M || k >= g.K) return 0.f;
    if (AMODE == 0) {
      return k < g.C ? __ldg(g.a0 + static_cast<size_t>(m) * g.pitch_in + k)
                     : __ldg(g.a1 + static_cast<size_t>(m) * g.pitch_in + (k - g.C));
    } else if (AMODE == 1) {
      const int t = k / g.C, c = k - t * g.C;
      const int i = t / g.f, j = t - i * g.f;
      const int y = py * g.f + i, x = px * g.f + j;
      return __ldg(g.a0 + ((static_cast<size_t>(pb) * g.H + y) * g.W + x) * g.pitch_in + c);
    } else {
      const int t = k / g.C, c = k - t * g.C;
      const int y = py + t / 3 - 1, x = px + t % 3 - 1;
      if (y < 0 || y >= g.H || x < 0 || x >= g.W) return 0.f;
      return __ldg(g.a0 + ((static_cast<size_t>(pb) * g.H + y) * g.W + x) * g.pitch_in + c);
    }
  };
  // B tile 16 x 64: k = tid / 16, n = (tid % 16) * 4 .. +3
  const int bk = tid >> 4, bn = (tid & 15) * 4;

  for (int k0 = 0; k0 < g.K; k0 += kBK) {
#pragma unroll
    for (int e = 0; e < 4; ++e) As[lk + e][lr] = load_a(k0 + lk + e);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int k = k0 + bk, n = n0 + bn + e;
      Bs[bk][bn + e] = (k < g.K && n < g.N) ? __ldg(g.wt + static_cast<size_t>(n) * g.K + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + ty * 4 + i;
    if (mm >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      if (EMODE == 0) {
        const float xv = __ldg(g.a0 + static_cast<size_t>(mm) * g.pitch_in + n);
        const float zv = __ldg(g.a1 + static_cast<size_t>(mm) * g.pitch_in + n);
        const float beta = 1.f / (1.f + expf(-acc[i][j]));
        const float w = g.gate * beta;
        const float o = (1.f - w) * xv + w * zv;  // (same expression as the reference: model.py:837)
        g.out[static_cast<size_t>(mm) * g.pitch_out + n] = o;
        if (g.out16) g.out16[static_cast<size_t>(mm) * g.pitch_out + n] =
            g.bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(o)) : __half_as_ushort(__float2half_rn(o));
      } else if (EMODE == 1) {
        const float o = acc[i][j] + (g.bias ? __ldg(g.bias + n) : 0.f);
        g.out[static_cast<size_t>(mm) * g.pitch_out + n] = o;
        if (g.out16) g.out16[static_cast<size_t>(mm) * g.pitch_out + n] =
            g.bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(o)) : __half_as_ushort(__float2half_rn(o));
      } else {
        const int b = mm / (g.Ho * g.Wo);
        atomicAdd(g.out + static_cast<size_t>(b) * g.N + n, acc[i][j] * g.inv_hw);
      }
    }
  }
}

template <int AMODE, int EMODE>
static int launch_gemm(const GemmArgs& g, cudaStream_t s) {
  const dim3 grid(static_cast<unsigned>(ceil_div(g.M, kBM)), static_cast<unsigned>(ceil_div(g.N, kBN)));
  gather_gemm_kernel<AMODE, EMODE><<<grid, 256, 0, s>>>(g);
  MZ_CUDA(cudaGetLastError());
  return MZ_OK;
}

// pooled epilogue: out[b][n] starts from the bias
__global__ void fill_bias_kernel(float* out, const float* bias, int B, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * N) out[i] = bias ? bias[i % N] : 0.f;
}

// PixelShuffle on NHWC: out[b, h r + i, w r + j, c] = in[b, h, w, c r r + i r + j]   (model.py:911,928 index rule).
// One thread per (input pixel, output channel): its r r input values are consecutive (a warp reads 32 r r consecutive
// floats; 16-byte loads when VEC), and for each (i, j) the warp writes 32 consecutive channels of one output pixel.
template <int R, bool VEC>
__global__ void __launch_bounds__(256) pixel_shuffle_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                                 uint16_t* __restrict__ out16, int bf16, int B, int H, int W, int C,
                                                                 int pitch_in, int pitch_out) {
  const long long total = static_cast<long long>(B) * H * W * C;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    const long long pix = idx / C;
    const int w = static_cast<int>(pix % W);
    const long long t = pix / W;
    const int h = static_cast<int>(t % H);
    const int b = static_cast<int>(t / H);
    const float* src = in + static_cast<size_t>(pix) * pitch_in + c * (R * R);
    float v[R * R];
    if (VEC) {
#pragma unroll
      for (int e = 0; e < R * R / 4; ++e) {
        const float4 q = __ldg(reinterpret_cast<const float4*>(src) + e);
        v[4 * e] = q.x, v[4 * e + 1] = q.y, v[4 * e + 2] = q.z, v[4 * e + 3] = q.w;
      }
    } else {
#pragma unroll
      for (int e = 0; e < R * R; ++e) v[e] = __ldg(src + e);
    }
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
      for (int j = 0; j < R; ++j) {
        const size_t o = ((static_cast<size_t>(b) * H * R + h * R + i) * (static_cast<size_t>(W) * R) + w * R + j) * pitch_out + c;
        out[o] = v[i * R + j];
        if (out16)
          out16[o] = bf16 ? __bfloat16_as_ushort(__float2bfloat16_rn(v[i * R + j])) : __half_as_ushort(__float2half_rn(v[i * R + j]));
      }
  }
}

// Decoder.crop_feature_maps (model.py:650-689): centre crop (start = (h - th) / 2) or zero pad (top = (th - h) / 2) per axis.
// V = 4: channels in 16-byte groups (C and the pitches multiples of 4, aligned bases); V = 1: any layout.
template <int V>
__global__ void __launch_bounds__(256) crop_pad_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int H, int W,
                                                            int C, int tH, int tW, int pitch_in, int pitch_out) {
  const int oy0 = H > tH ? (H - tH) / 2 : -((tH - H) / 2);  // input row of output row 0
  const int ox0 = W > tW ? (W - tW) / 2 : -((tW - W) / 2);
  const int Cv = C / V;
  const long long total = static_cast<long long>(B) * tH * tW * Cv;
  for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % Cv) * V;
    long long t = idx / Cv;
    const int x = static_cast<int>(t % tW);
    t /= tW;
    const int y = static_cast<int>(t % tH);
    const int b = static_cast<int>(t / tH);
    const int iy = y + oy0, ix = x + ox0;
    const bool inside = iy >= 0 && iy < H && ix >= 0 && ix < W;
    const size_t src = ((static_cast<size_t>(b) * H + iy) * W + ix) * pitch_in + c;
    const size_t dst = ((static_cast<size_t>(b) * tH + y) * tW + x) * pitch_out + c;
    if (V == 4) {
      const float4 v = inside ? __ldg(reinterpret_cast<const float4*>(in + src)) : make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(out + dst) = v;
    } else {
      out[dst] = inside ? __ldg(in + src) : 0.f;
    }
  }
}

static unsigned blocks_for(long long total) {
  long long b = (total + 255) / 256;
  if (b > 148LL * 64) b = 148LL * 64;
  return static_cast<unsigned>(b < 1 ? 1 : b);
}

}  // namespace mz

using namespace mz;

extern "C" {

int mz_adaptive_mix(const float* x_dev, const float* z_dev, const float* wt_dev, float alpha_logit, float* out_dev,
                    void* out16_dev, int64_t npix, int32_t C, int32_t pitch, int32_t operand_dtype, int32_t math, void* stream) {
  MZ_REQUIRE(x_dev && z_dev && wt_dev && out_dev, "adaptive_mix: null pointer");
  MZ_REQUIRE(npix > 0 && npix < (1LL << 31) && C > 0 && pitch >= C, "adaptive_mix: bad shape (npix %lld, C %d, pitch %d)",
             static_cast<long long>(npix), C, pitch);
  MZ_REQUIRE(dtype_ok(operand_dtype), "operand_dtype must be MZ_DTYPE_F16 or MZ_DTYPE_BF16, %d given", operand_dtype);
  MZ_REQUIRE(math == MZ_MATH_TF32 || math == MZ_MATH_FP32, "math must be MZ_MATH_TF32 or MZ_MATH_FP32, %d given", math);
  MZ_REQUIRE(out_dev != x_dev && out_dev != z_dev, "adaptive_mix: the output needs its own buffer");
  if (math == MZ_MATH_TF32) {
    MZ_REQUIRE(seg_gemm_tc_applies(x_dev, z_dev, wt_dev, out_dev, out16_dev, C, C, 2 * C, pitch, pitch),
               "adaptive_mix (tf32): pointers must be 16-byte aligned, C and pitch multiples of 4 (8 with a 16-bit shadow)");
    SgArgs a;
    memset(&a, 0, sizeof(a));
    a.a0 = x_dev;
    a.a1 = z_dev;
    a.wt = wt_dev;
    a.out = out_dev;
    a.out16 = out16_dev;
    a.bf16 = operand_dtype == MZ_DTYPE_BF16;
    a.mix = 1;
    a.gate = 1.f / (1.f + expf(-alpha_logit));
    a.C = a.N = C;
    a.f = 1;
    a.rows = 1;
    a.wo = npix;
    a.pitch_in = a.pitch_out = pitch;
    return launch_seg_gemm_tc(a, static_cast<cudaStream_t>(stream));
  }
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.a0 = x_dev;
  g.a1 = z_dev;
  g.wt = wt_dev;
  g.out = out_dev;
  g.out16 = static_cast<uint16_t*>(out16_dev);
  g.bf16 = operand_dtype == MZ_DTYPE_BF16;
  g.M = static_cast<int>(npix);
  g.N = C;
  g.K = 2 * C;
  g.C = C;
  g.pitch_in = g.pitch_out = pitch;
  g.gate = 1.f / (1.f + expf(-alpha_logit));
  return launch_gemm<0, 0>(g, static_cast<cudaStream_t>(stream));
}

int mz_pixel_crush(const float* in_dev, const float* wt_dev, float* out_dev, void* out16_dev, int32_t B, int32_t H, int32_t W,
                   int32_t Cin, int32_t Cout, int32_t factor, int32_t pitch_in, int32_t pitch_out, int32_t operand_dtype,
                   int32_t math, void* stream) {
  MZ_REQUIRE(in_dev && wt_dev && out_dev, "pixel_crush: null pointer");
  MZ_REQUIRE(factor == 2 || factor == 3 || factor == 4, "Crush factor must be either 2, 3, or 4, %d given.", factor);
  MZ_REQUIRE(Cin > 0, "Input channels must be greater than 0.");
  MZ_REQUIRE(Cout > 0, "Output channels must be greater than 0.");
  MZ_REQUIRE(B > 0 && H >= factor && W >= factor && pitch_in >= Cin && pitch_out >= Cout, "pixel_crush: bad shape");
  MZ_REQUIRE(dtype_ok(operand_dtype), "operand_dtype must be MZ_DTYPE_F16 or MZ_DTYPE_BF16, %d given", operand_dtype);
  MZ_REQUIRE(math == MZ_MATH_TF32 || math == MZ_MATH_FP32, "math must be MZ_MATH_TF32 or MZ_MATH_FP32, %d given", math);
  if (math == MZ_MATH_TF32) {
    MZ_REQUIRE(seg_gemm_tc_applies(in_dev, nullptr, wt_dev, out_dev, out16_dev, Cin, Cout, factor * factor * Cin, pitch_in, pitch_out),
               "pixel_crush (tf32): pointers must be 16-byte aligned, Cin and the pitches multiples of 4 (8 with a 16-bit shadow)");
    SgArgs a;
    memset(&a, 0, sizeof(a));
    a.a0 = in_dev;
    a.wt = wt_dev;
    a.out = out_dev;
    a.out16 = out16_dev;
    a.bf16 = operand_dtype == MZ_DTYPE_BF16;
    a.C = Cin;
    a.N = Cout;
    a.f = factor;
    a.H = H;
    a.W = W;
    a.Ho = H / factor;
    a.rows = static_cast<long long>(B) * a.Ho;
    a.wo = W / factor;
    a.in_rows = static_cast<long long>(B) * H;
    a.pitch_in = pitch_in;
    a.pitch_out = pitch_out;
    return launch_seg_gemm_tc(a, static_cast<cudaStream_t>(stream));
  }
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.a0 = in_dev;
  g.wt = wt_dev;
  g.out = out_dev;
  g.out16 = static_cast<uint16_t*>(out16_dev);
  g.bf16 = operand_dtype == MZ_DTYPE_BF16;
  g.B = B;
  g.H = H;
  g.W = W;
  g.C = Cin;
  g.f = factor;
  g.Ho = H / factor;
  g.Wo = W / factor;
  const long long M = static_cast<long long>(B) * g.Ho * g.Wo;
  MZ_REQUIRE(M < (1LL << 31), "pixel_crush: too many pixels");
  g.M = static_cast<int>(M);
  g.N = Cout;
  g.K = factor * factor * Cin;
  g.pitch_in = pitch_in;
  g.pitch_out = pitch_out;
  return launch_gemm<1, 1>(g, static_cast<cudaStream_t>(stream));
}

int mz_quality_assessor(const float* in_dev, const float* wt_dev, const float* bias_dev, float* out_dev, int32_t B, int32_t H,
                        int32_t W, int32_t C, int32_t F, int32_t pitch_in, void* stream) {
  MZ_REQUIRE(in_dev && wt_dev && out_dev, "quality_assessor: null pointer");
  MZ_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && F > 0 && pitch_in >= C, "quality_assessor: bad shape");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  fill_bias_kernel<<<ceil_div(B * F, 128), 128, 0, s>>>(out_dev, bias_dev, B, F);
  MZ_CUDA(cudaGetLastError());
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.a0 = in_dev;
  g.wt = wt_dev;
  g.out = out_dev;
  g.B = B;
  g.H = g.Ho = H;
  g.W = g.Wo = W;
  g.C = C;
  const long long M = static_cast<long long>(B) * H * W;
  MZ_REQUIRE(M < (1LL << 31), "quality_assessor: too many pixels");
  g.M = static_cast<int>(M);
  g.N = F;
  g.K = 9 * C;
  g.pitch_in = pitch_in;
  g.inv_hw = 1.f / (static_cast<float>(H) * static_cast<float>(W));
  // (a 64-pixel tile must not straddle two images for the per-image pooled sum: the kernel resolves the image per row)
  return launch_gemm<2, 2>(g, s);
}

int mz_pixel_shuffle_nhwc(const float* in_dev, float* out_dev, void* out16_dev, int32_t B, int32_t H, int32_t W, int32_t C,
                          int32_t r, int32_t pitch_in, int32_t pitch_out, int32_t operand_dtype, void* stream) {
  MZ_REQUIRE(in_dev && out_dev, "pixel_shuffle: null pointer");
  MZ_REQUIRE(r == 2 || r == 3 || r == 4, "Upscale ratio must be either 2, 3, or 4, %d given.", r);
  MZ_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && pitch_in >= C * r * r && pitch_out >= C, "pixel_shuffle: bad shape");
  MZ_REQUIRE(dtype_ok(operand_dtype), "operand_dtype must be MZ_DTYPE_F16 or MZ_DTYPE_BF16, %d given", operand_dtype);
  const long long total = static_cast<long long>(B) * H * W * C;
  const bool vec = (reinterpret_cast<uintptr_t>(in_dev) & 15u) == 0 && pitch_in % 4 == 0 && r != 3;
  const unsigned nb = blocks_for(total);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  uint16_t* o16 = static_cast<uint16_t*>(out16_dev);
  const int bf = operand_dtype == MZ_DTYPE_BF16;
  if (r == 2 && vec) pixel_shuffle_nhwc_kernel<2, true><<<nb, 256, 0, s>>>(in_dev, out_dev, o16, bf, B, H, W, C, pitch_in, pitch_out);
  else if (r == 2) pixel_shuffle_nhwc_kernel<2, false><<<nb, 256, 0, s>>>(in_dev, out_dev, o16, bf, B, H, W, C, pitch_in, pitch_out);
  else if (r == 3) pixel_shuffle_nhwc_kernel<3, false><<<nb, 256, 0, s>>>(in_dev, out_dev, o16, bf, B, H, W, C, pitch_in, pitch_out);
  else if (vec) pixel_shuffle_nhwc_kernel<4, true><<<nb, 256, 0, s>>>(in_dev, out_dev, o16, bf, B, H, W, C, pitch_in, pitch_out);
  else pixel_shuffle_nhwc_kernel<4, false><<<nb, 256, 0, s>>>(in_dev, out_dev, o16, bf, B, H, W, C, pitch_in, pitch_out);
  MZ_CUDA(cudaGetLastError());
  return MZ_OK;
}

int mz_crop_feature_maps(const float* in_dev, float* out_dev, int32_t B, int32_t H, int32_t W, int32_t C, int32_t target_h,
                         int32_t target_w, int32_t pitch_in, int32_t pitch_out, void* stream) {
  MZ_REQUIRE(in_dev && out_dev, "crop_feature_maps: null pointer");
  MZ_REQUIRE(B > 0 && H > 0 && W > 0 && C > 0 && target_h > 0 && target_w > 0 && pitch_in >= C && pitch_out >= C,
             "crop_feature_maps: bad shape");
  const bool vec = ((reinterpret_cast<uintptr_t>(in_dev) | reinterpret_cast<uintptr_t>(out_dev)) & 15u) == 0 && C % 4 == 0 &&
                   pitch_in % 4 == 0 && pitch_out % 4 == 0;
  const long long total = static_cast<long long>(B) * target_h * target_w * (vec ? C / 4 : C);
  if (vec)
    crop_pad_nhwc_kernel<4><<<blocks_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(in_dev, out_dev, B, H, W, C, target_h,
                                                                                             target_w, pitch_in, pitch_out);
  else
    crop_pad_nhwc_kernel<1><<<blocks_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(in_dev, out_dev, B, H, W, C, target_h,
                                                                                             target_w, pitch_in, pitch_out);
  MZ_CUDA(cudaGetLastError());
  return MZ_OK;
}

}  // extern "C"
