// P1: letterbox + normalise (rtm_letterbox).
//
// Reproduces, per frame, what ultralytics' LetterBox(auto=False, scaleup=True, center=True)
// followed by BasePredictor.preprocess do on the way into detector.py:100-111:
//   cv2.resize(img, new_unpad, INTER_LINEAR)          8-bit path of OpenCV: 11-bit fixed-point
//                                                     horizontal taps, then
//                                                     ((b0*(S0>>4))>>16 + (b1*(S1>>4))>>16 + 2) >> 2
//   cv2.copyMakeBorder(..., BORDER_CONSTANT, 114)     top/left = round(pad - 0.1)
//   BGR -> RGB, HWC -> CHW, cast, / 255               division in float32, rounded once to the
//                                                     output type (what torch does for half / bf16)
// The 1080p -> 640x360 case lands exactly on source pixels (every third row and pixel, zero
// second tap) and 720p -> 640x360 on the 2x2 box mean; both fall out of the general
// fixed-point formula, so there is one code path.  Taps with zero weight are not loaded.
//
// Mapping: one thread produces 8 horizontally consecutive output pixels of all three
// channel planes (three 16-byte stores for bf16/f16), so stores are fully coalesced; the
// source rows are read through the read-only path.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>
#include <type_traits>
#include <stdlib.h>
#include <string.h>

#include "rtm_common.cuh"

namespace {

constexpr int kPix = 8;  // output pixels per thread
constexpr int kCoefBits = 11;

struct LetterboxArgs {
  const uint8_t* frames;
  int src_h, src_w;
  long long row_stride, frame_stride;
  void* out;
  int out_h, out_w;
  int new_h, new_w, top, left;
  double scale_x, scale_y;  // src / new, as cv::resize derives them (1 / inv_scale)
  bool resize;              // false when the source already has the unpadded size
};

// OpenCV resize coordinate + fixed-point taps for one destination index.  Horizontally the
// fraction is reset at the borders; vertically the weights are kept and the rows clipped
// (cv::resize computes xofs/alpha and yofs/beta that way) - it matters only when up-scaling.
template <bool HORIZONTAL>
__device__ __forceinline__ void linear_tap(int d, double scale, int ssize, int* s0, int* s1, int* c0, int* c1) {
  float f = static_cast<float>(__dsub_rn(__dmul_rn(d + 0.5, scale), 0.5));  // no FMA contraction
  int s = static_cast<int>(floorf(f));
  f -= static_cast<float>(s);
  if (HORIZONTAL) {
    if (s < 0) {
      s = 0;
      f = 0.f;
    }
    if (s >= ssize - 1) {
      s = ssize - 1;
      f = 0.f;
    }
    *s0 = s;
    *s1 = min(s + 1, ssize - 1);
  } else {
    *s0 = min(max(s, 0), ssize - 1);
    *s1 = min(max(s + 1, 0), ssize - 1);
  }
  *c0 = __float2int_rn((1.f - f) * static_cast<float>(1 << kCoefBits));
  *c1 = __float2int_rn(f * static_cast<float>(1 << kCoefBits));
}

template <typename T>
__device__ __forceinline__ T from_u8(int v);
template <>
__device__ __forceinline__ float from_u8<float>(int v) {
  return __fdiv_rn(static_cast<float>(v), 255.f);
}
template <>
__device__ __forceinline__ __half from_u8<__half>(int v) {
  return __float2half_rn(__fdiv_rn(static_cast<float>(v), 255.f));
}
template <>
__device__ __forceinline__ __nv_bfloat16 from_u8<__nv_bfloat16>(int v) {
  return __float2bfloat16_rn(__fdiv_rn(static_cast<float>(v), 255.f));
}

template <typename T>
struct alignas(sizeof(T) * kPix) OutPack {
  T v[kPix];
};

template <typename T>
__global__ void __launch_bounds__(256) letterbox_kernel(const LetterboxArgs a) {
  // u8 -> T(v / 255) for the 256 possible values, one IEEE division each instead of one per pixel
  __shared__ T s_lut[256];
  s_lut[threadIdx.x] = from_u8<T>(threadIdx.x);
  __syncthreads();
  const int groups_per_row = a.out_w / kPix;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups_per_row * a.out_h) return;
  const int b = blockIdx.y;
  const int oy = g / groups_per_row;
  const int ox0 = (g - oy * groups_per_row) * kPix;
  const uint8_t* frame = a.frames + static_cast<long long>(b) * a.frame_stride;

  int px[3][kPix];  // [channel in RGB order][pixel]
  const int ry = oy - a.top;
  if (ry < 0 || ry >= a.new_h) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int e = 0; e < kPix; ++e) px[c][e] = 114;
  } else {
    int y0 = ry, y1 = ry, b0 = 1 << kCoefBits, b1 = 0;
    if (a.resize) linear_tap<false>(ry, a.scale_y, a.src_h, &y0, &y1, &b0, &b1);
    const uint8_t* r0 = frame + static_cast<long long>(y0) * a.row_stride;
    const uint8_t* r1 = frame + static_cast<long long>(y1) * a.row_stride;
#pragma unroll
    for (int e = 0; e < kPix; ++e) {
      const int rx = ox0 + e - a.left;
      if (rx < 0 || rx >= a.new_w) {
        px[0][e] = px[1][e] = px[2][e] = 114;
        continue;
      }
      int x0 = rx, x1 = rx, a0 = 1 << kCoefBits, a1 = 0;
      if (a.resize) linear_tap<true>(rx, a.scale_x, a.src_w, &x0, &x1, &a0, &a1);
#pragma unroll
      for (int c = 0; c < 3; ++c) {  // c: B, G, R in the source
        int v;
        if (!a.resize) {
          v = __ldg(r0 + x0 * 3 + c);
        } else {
          int s0 = a0 * __ldg(r0 + x0 * 3 + c);
          if (a1) s0 += a1 * __ldg(r0 + x1 * 3 + c);
          int s1 = 0;
          if (b1) {
            s1 = a0 * __ldg(r1 + x0 * 3 + c);
            if (a1) s1 += a1 * __ldg(r1 + x1 * 3 + c);
          }
          v = (((b0 * (s0 >> 4)) >> 16) + ((b1 * (s1 >> 4)) >> 16) + 2) >> 2;
          v = min(max(v, 0), 255);
        }
        px[2 - c][e] = v;  // BGR -> RGB
      }
    }
  }
  T* out = static_cast<T*>(a.out) + (static_cast<long long>(b) * 3 * a.out_h + oy) * a.out_w + ox0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    OutPack<T> p;
#pragma unroll
    for (int e = 0; e < kPix; ++e) p.v[e] = s_lut[px[c][e]];
    *reinterpret_cast<OutPack<T>*>(out + static_cast<long long>(c) * a.out_h * a.out_w) = p;
  }
}

// 3 : 1 decimation (1080p -> 640 x 360 inside 640 x 640, the BASELINE configuration): OpenCV's
// fixed-point bilinear lands exactly on source pixel (3 rx + 1, 3 ry + 1) with unit weight, so the
// resize is a strided gather.  A thread produces 16 consecutive output pixels: their 48 source bytes
// lie in a 138-byte span that starts 3 bytes into a 16-byte aligned block (9 * 16 g + 3), fetched as
// nine aligned 16-byte loads; every byte is then picked with a compile-time word index and shift.
// The host checks the geometry (taps, alignment) before choosing this kernel.
constexpr int kDecPix = 16;
template <typename T>
__global__ void __launch_bounds__(256) letterbox_decimate3_kernel(const LetterboxArgs a) {
  __shared__ T s_lut[256];
  s_lut[threadIdx.x] = from_u8<T>(threadIdx.x);
  __syncthreads();
  const int groups_per_row = a.out_w / kDecPix;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups_per_row * a.out_h) return;
  const int b = blockIdx.y;
  const int oy = g / groups_per_row;
  const int ox0 = (g - oy * groups_per_row) * kDecPix;
  const int ry = oy - a.top, rx0 = ox0 - a.left;
  T* out = static_cast<T*>(a.out) + (static_cast<long long>(b) * 3 * a.out_h + oy) * a.out_w + ox0;
  const long long plane = static_cast<long long>(a.out_h) * a.out_w;
  using Pack = OutPack<T>;  // 8 values
  if (ry < 0 || ry >= a.new_h || rx0 < 0 || rx0 >= a.new_w) {  // whole groups are inside or outside (host-checked)
    Pack p;
#pragma unroll
    for (int e = 0; e < kPix; ++e) p.v[e] = s_lut[114];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      *reinterpret_cast<Pack*>(out + c * plane) = p;
      *reinterpret_cast<Pack*>(out + c * plane + kPix) = p;
    }
    return;
  }
  const uint8_t* row = a.frames + static_cast<long long>(b) * a.frame_stride + static_cast<long long>(3 * ry + 1) * a.row_stride;
  const uint4* src = reinterpret_cast<const uint4*>(row + 9 * rx0);  // 9 * 16 g: 16-byte aligned; first tap at byte 3
  uint32_t w[36];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const uint4 v = __ldg(src + k);
    w[4 * k] = v.x;
    w[4 * k + 1] = v.y;
    w[4 * k + 2] = v.z;
    w[4 * k + 3] = v.w;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {  // c: B, G, R in the source -> plane 2 - c
    Pack lo, hi;
#pragma unroll
    for (int e = 0; e < kDecPix; ++e) {
      const int o = 3 + 9 * e + c;  // byte offset inside the span: compile-time
      const T val = s_lut[(w[o >> 2] >> ((o & 3) * 8)) & 0xffu];
      if (e < kPix) lo.v[e] = val;
      else hi.v[e - kPix] = val;
    }
    *reinterpret_cast<Pack*>(out + (2 - c) * plane) = lo;
    *reinterpret_cast<Pack*>(out + (2 - c) * plane + kPix) = hi;
  }
}

// The same gather for 16-bit outputs, twice as wide: a thread produces 32 consecutive output pixels from nine aligned
// 32-byte loads (LDG.E.256: every lane owns whole sectors - the 16-byte version asked L1 for 32 half-sectors per
// instruction, ncu: l1tex at 75 % of its peak) and writes each plane as two 32-byte stores (whole sectors again; the
// 16-byte stores sent every sector to L2 twice, half filled).  u8 -> T(v / 255) without the table: v * 0x3b808081
// (the float next to 1 / 255) rounds to the same bf16 / f16 as the IEEE quotient for all 256 values
// (tests/test_gpu_detect.py checks the kernel against cv2 + torch bit for bit), so a byte costs one PRMT that plants it
// in the mantissa of 2^23, one FFMA and half a pack instruction, and no shared-memory lookup with random bank conflicts.
constexpr int kWidePix = 32;

__device__ __forceinline__ void ldg256_stream(const void* p, uint32_t* w) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* w) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
               "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
               : "memory");
}
template <typename T>
__device__ __forceinline__ uint32_t pack_unit_pair(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack_unit_pair<__nv_bfloat16>(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack_unit_pair<__half>(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// byte k (compile-time) of word x as a float in [0, 1]: rn(v * c), c = 0x3b808081
template <int K>
__device__ __forceinline__ float unit_from_byte(uint32_t x) {
  const float c = __uint_as_float(0x3b808081u);
  const float t = __uint_as_float(__byte_perm(x, 0x4B000000u, 0x7650 | K));  // 2^23 + v, exact
  return __fmaf_rn(t, c, -8388608.f * c);                                       // 2^23 * c is exact: = rn(v * c)
}

template <typename T>
__global__ void __launch_bounds__(128) letterbox_decimate3_wide_kernel(const LetterboxArgs a) {
  static_assert(sizeof(T) == 2, "16-bit outputs only");
  const int groups_per_row = a.out_w / kWidePix;
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups_per_row * a.out_h) return;
  const int b = blockIdx.y;
  const int oy = g / groups_per_row;
  const int ox0 = (g - oy * groups_per_row) * kWidePix;
  const int ry = oy - a.top, rx0 = ox0 - a.left;
  T* out = static_cast<T*>(a.out) + (static_cast<long long>(b) * 3 * a.out_h + oy) * a.out_w + ox0;
  const long long plane = static_cast<long long>(a.out_h) * a.out_w;
  if (ry < 0 || ry >= a.new_h || rx0 < 0 || rx0 >= a.new_w) {  // whole groups are inside or outside (host-checked)
    const float pad = unit_from_byte<0>(114u);
    uint32_t p[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) p[e] = pack_unit_pair<T>(pad, pad);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      stg256(out + c * plane, p);
      stg256(out + c * plane + 16, p);
    }
    return;
  }
  const uint8_t* row = a.frames + static_cast<long long>(b) * a.frame_stride + static_cast<long long>(3 * ry + 1) * a.row_stride;
  const uint8_t* src = row + 9 * rx0;  // 9 * 32 g: 32-byte aligned; first tap at byte 3
  uint32_t w[72];
#pragma unroll
  for (int k = 0; k < 9; ++k) ldg256_stream(src + 32 * k, w + 8 * k);
#pragma unroll
  for (int c = 0; c < 3; ++c) {  // c: B, G, R in the source -> plane 2 - c
    uint32_t o[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int o0 = 3 + 9 * (2 * e) + c, o1 = 3 + 9 * (2 * e + 1) + c;  // byte offsets inside the span: compile-time
      float lo, hi;
      switch (o0 & 3) {
        case 0: lo = unit_from_byte<0>(w[o0 >> 2]); break;
        case 1: lo = unit_from_byte<1>(w[o0 >> 2]); break;
        case 2: lo = unit_from_byte<2>(w[o0 >> 2]); break;
        default: lo = unit_from_byte<3>(w[o0 >> 2]); break;
      }
      switch (o1 & 3) {
        case 0: hi = unit_from_byte<0>(w[o1 >> 2]); break;
        case 1: hi = unit_from_byte<1>(w[o1 >> 2]); break;
        case 2: hi = unit_from_byte<2>(w[o1 >> 2]); break;
        default: hi = unit_from_byte<3>(w[o1 >> 2]); break;
      }
      o[e] = pack_unit_pair<T>(lo, hi);
    }
    stg256(out + (2 - c) * plane, o);
    stg256(out + (2 - c) * plane + 16, o + 8);
  }
}

// General ratios, 4-byte aligned rows: a block owns a tile of 16 output rows x 128 output columns.
//   * the horizontal taps of the 128 columns are computed once per block (shared memory) instead of once per pixel:
//     byte offset of the left tap in its row, and the weight pair packed for DP2A;
//   * lane l owns columns l, l + 32, l + 64, l + 96 of two rows, so the lanes of a warp read NEIGHBOURING source
//     pixels (the 8-pixels-per-thread kernel sent every byte load to 12 different cache lines);
//   * the 6 bytes of a tap pair (B G R B G R at 3 x0) come from three aligned 32-bit loads and two funnel shifts; one
//     PRMT puts the two bytes of a channel side by side and one DP2A forms a0 p0 + a1 p1 (cv::resize's horizontal pass);
//   * 16-bit outputs are converted like the 3 : 1 kernel (exact for all 256 values), float goes through the table.
// The arithmetic is the arithmetic of letterbox_kernel; unit weights (no resize, borders) reproduce the source byte.
constexpr int kTileCols = 128, kTileRows = 16;  // 4 columns per lane, 2 rows per warp

// The three aligned words that hold bytes 3 x0 .. 3 x0 + 5 of a row, for the four columns of a lane at once (all loads
// are issued before any is used: the kernel lives on the loads it keeps in flight).  Words past the end of a row only
// ever meet zero weights, so reading on into the next row is harmless; CLAMP (the last source row of the last frame,
// where "on" would leave the buffer) replaces them by the row's last word.
template <bool CLAMP>
__device__ __forceinline__ void tap_words(const uint8_t* row, const uint32_t (&wo)[4], uint32_t last_word, uint32_t (&w)[4][3]) {
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (CLAMP) {
      w[k][0] = __ldg(reinterpret_cast<const uint32_t*>(row + wo[k]));
      w[k][1] = __ldg(reinterpret_cast<const uint32_t*>(row + min(wo[k] + 4u, last_word)));
      w[k][2] = __ldg(reinterpret_cast<const uint32_t*>(row + min(wo[k] + 8u, last_word)));
    } else {
      const uint32_t* p = reinterpret_cast<const uint32_t*>(row + wo[k]);
      w[k][0] = __ldg(p);
      w[k][1] = __ldg(p + 1);
      w[k][2] = __ldg(p + 2);
    }
  }
}
// cv::resize's horizontal pass on the words of one column: a0 p0 + a1 p1 per channel (B, G, R)
__device__ __forceinline__ void tap_dot(const uint32_t (&w)[3], uint32_t sh, uint32_t wt, uint32_t (&h)[3]) {
  const uint32_t lo = __funnelshift_r(w[0], w[1], sh), hi = __funnelshift_r(w[1], w[2], sh);  // bytes 0..3, 4..7 from 3 x0
  h[0] = __dp2a_lo(wt, __byte_perm(lo, hi, 0x3030), 0u);                                      // B: bytes 0, 3
  h[1] = __dp2a_lo(wt, __byte_perm(lo, hi, 0x4141), 0u);                                      // G: bytes 1, 4
  h[2] = __dp2a_lo(wt, __byte_perm(lo, hi, 0x5252), 0u);                                      // R: bytes 2, 5
}
template <typename T>
__device__ __forceinline__ T tile_unit(uint32_t v, const T* lut) {  // T(v / 255), v <= 255
  if constexpr (sizeof(T) == 4) {
    return lut[v];
  } else {
    const float f = unit_from_byte<0>(v);
    if constexpr (std::is_same<T, __half>::value) return __float2half_rn(f);
    else return __float2bfloat16_rn(f);
  }
}

template <typename T>
__global__ void __launch_bounds__(256, 4) letterbox_tile_kernel(const LetterboxArgs a) {
  __shared__ int s_off[kTileCols];       // 3 * x0, or -1 for a padding column
  __shared__ uint32_t s_wt[kTileCols];   // a0 | a1 << 16 (0 for a padding column)
  __shared__ T s_lut[sizeof(T) == 4 ? 256 : 1];
  const int tid = threadIdx.x;
  const int col0 = blockIdx.x * kTileCols, row0 = blockIdx.y * kTileRows, b = blockIdx.z;
  if (sizeof(T) == 4) s_lut[tid] = from_u8<T>(tid);
  if (tid < kTileCols) {
    const int rx = col0 + tid - a.left;
    int off = -1;
    uint32_t wt = 0;
    if (rx >= 0 && rx < a.new_w && col0 + tid < a.out_w) {
      int x0 = rx, x1 = rx, a0 = 1 << kCoefBits, a1 = 0;
      if (a.resize) linear_tap<true>(rx, a.scale_x, a.src_w, &x0, &x1, &a0, &a1);
      off = 3 * x0;  // a1 != 0 implies x1 == x0 + 1 (the fraction is reset where the taps are clamped)
      wt = static_cast<uint32_t>(a0) | (static_cast<uint32_t>(a1) << 16);
    }
    s_off[tid] = off;
    s_wt[tid] = wt;
  }
  __syncthreads();
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t last_word = static_cast<uint32_t>(a.row_stride) - 4u;  // byte offset of the row's last whole word
  const uint8_t* frame = a.frames + static_cast<long long>(b) * a.frame_stride;
  const long long plane = static_cast<long long>(a.out_h) * a.out_w;
  const T pad = tile_unit<T>(114u, s_lut);
#pragma unroll
  for (int rr = 0; rr < 2; ++rr) {
    const int oy = row0 + 2 * warp + rr;
    if (oy >= a.out_h) break;
    const int ry = oy - a.top;
    T* out = static_cast<T*>(a.out) + (static_cast<long long>(b) * 3 * a.out_h + oy) * a.out_w + col0 + lane;
    if (ry < 0 || ry >= a.new_h) {  // padding row
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (col0 + lane + 32 * k < a.out_w) out[32 * k] = out[plane + 32 * k] = out[2 * plane + 32 * k] = pad;
      continue;
    }
    int y0 = ry, y1 = ry, b0 = 1 << kCoefBits, b1 = 0;
    if (a.resize) linear_tap<false>(ry, a.scale_y, a.src_h, &y0, &y1, &b0, &b1);
    const uint8_t* r0 = frame + static_cast<long long>(y0) * a.row_stride;
    const uint8_t* r1 = frame + static_cast<long long>(y1) * a.row_stride;
    const uint32_t m0 = static_cast<uint32_t>(b0) << 16, m1 = static_cast<uint32_t>(b1) << 16;  // (b x) >> 16 = umulhi(b << 16, x)
    uint32_t wo[4], sh[4], wt[4];
    bool live[4];  // a pixel of the resized image (not a padding column, not past the output row)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int col = lane + 32 * k;
      const int off = s_off[col];
      live[k] = off >= 0;
      wt[k] = s_wt[col];
      wo[k] = live[k] ? static_cast<uint32_t>(off) & ~3u : 0u;
      sh[k] = (static_cast<uint32_t>(off) & 3u) * 8u;
    }
    uint32_t w0[4][3], w1[4][3];
    const bool last_rows = b == static_cast<int>(gridDim.z) - 1 && max(y0, y1) == a.src_h - 1;
    if (!last_rows) {
      tap_words<false>(r0, wo, last_word, w0);
      if (b1) tap_words<false>(r1, wo, last_word, w1);
    } else {
      tap_words<true>(r0, wo, last_word, w0);
      if (b1) tap_words<true>(r1, wo, last_word, w1);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      uint32_t h0[3], h1[3] = {0u, 0u, 0u};
      tap_dot(w0[k], sh[k], wt[k], h0);
      if (b1) tap_dot(w1[k], sh[k], wt[k], h1);
      if (col0 + lane + 32 * k < a.out_w) {
        // cv::resize's vertical pass.  No clamp: with a0 + a1 <= 2049 and b0 + b1 <= 2049 (each weight is a rounding of
        // f * 2048 or (1 - f) * 2048) the sum is at most (2049 * 32655) >> 16 = 1020, i.e. 255 after the final shift
        const T v0 = tile_unit<T>((__umulhi(m0, h0[0] >> 4) + __umulhi(m1, h1[0] >> 4) + 2u) >> 2, s_lut);
        const T v1 = tile_unit<T>((__umulhi(m0, h0[1] >> 4) + __umulhi(m1, h1[1] >> 4) + 2u) >> 2, s_lut);
        const T v2 = tile_unit<T>((__umulhi(m0, h0[2] >> 4) + __umulhi(m1, h1[2] >> 4) + 2u) >> 2, s_lut);
        out[2 * plane + 32 * k] = live[k] ? v0 : pad;  // BGR -> RGB planes
        out[plane + 32 * k] = live[k] ? v1 : pad;
        out[32 * k] = live[k] ? v2 : pad;
      }
    }
  }
}

// host copy of linear_tap (same IEEE operations) to recognise the pure 3 : 1 decimation
bool taps_are_decimate3(int dst, double scale, int ssize, bool horizontal) {
  for (int d = 0; d < dst; ++d) {
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    int s = static_cast<int>(floorf(f));
    f -= static_cast<float>(s);
    if (horizontal && (s < 0 || s >= ssize - 1)) return false;
    const int c1 = static_cast<int>(nearbyintf(f * 2048.f)), c0 = static_cast<int>(nearbyintf((1.f - f) * 2048.f));
    if (s != 3 * d + 1 || c1 != 0 || c0 != 2048) return false;
  }
  return true;
}

}  // namespace

namespace {

int launch_letterbox(const uint8_t* frames, int num_streams, int src_h, int src_w, int64_t row_stride, int64_t frame_stride,
                     void* out, int out_dtype, int out_h, int out_w, int new_h, int new_w, int top, int left,
                     rtm_cuda_stream stream) {
  RTM_REQUIRE(frames && out, "rtm_letterbox: null pointer");
  RTM_REQUIRE(num_streams > 0 && src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0, "rtm_letterbox: bad shape");
  RTM_REQUIRE(out_w % kPix == 0, "rtm_letterbox: out_w %d must be a multiple of %d", out_w, kPix);
  RTM_REQUIRE(row_stride >= 3ll * src_w && frame_stride >= row_stride * src_h, "rtm_letterbox: strides too small");
  RTM_REQUIRE((reinterpret_cast<uintptr_t>(out) & 31) == 0, "rtm_letterbox: out must be 32-byte aligned");
  RTM_REQUIRE(new_w >= 1 && new_h >= 1 && left >= 0 && top >= 0 && left + new_w <= out_w && top + new_h <= out_h,
              "rtm_letterbox: resized image %dx%d at (%d, %d) does not fit the %dx%d output", new_w, new_h, left, top, out_w,
              out_h);
  LetterboxArgs a;
  a.frames = frames;
  a.src_h = src_h;
  a.src_w = src_w;
  a.row_stride = row_stride;
  a.frame_stride = frame_stride;
  a.out = out;
  a.out_h = out_h;
  a.out_w = out_w;
  a.new_w = new_w;
  a.new_h = new_h;
  a.left = left;
  a.top = top;
  a.resize = !(a.new_w == src_w && a.new_h == src_h);
  // cv::resize: inv_scale = dsize / ssize; scale = 1. / inv_scale
  a.scale_x = 1.0 / (static_cast<double>(a.new_w) / src_w);
  a.scale_y = 1.0 / (static_cast<double>(a.new_h) / src_h);
  const int groups = (out_w / kPix) * out_h;
  dim3 grid((groups + 255) / 256, num_streams);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  rtm::ProfileScope prof(RTM_K_LETTERBOX, s);
  // 3 : 1 decimation fast path (RTM_LETTERBOX_IMPL=direct turns it off)
  static const bool allow_fast = !(getenv("RTM_LETTERBOX_IMPL") && strcmp(getenv("RTM_LETTERBOX_IMPL"), "direct") == 0);
  if (allow_fast && a.resize && src_w == 3 * new_w && src_h == 3 * new_h && out_w % kDecPix == 0 && left % kDecPix == 0 &&
      new_w % kDecPix == 0 && (reinterpret_cast<uintptr_t>(frames) & 15) == 0 && row_stride % 16 == 0 && frame_stride % 16 == 0 &&
      row_stride >= 9ll * new_w && taps_are_decimate3(new_w, a.scale_x, src_w, true) &&
      taps_are_decimate3(new_h, a.scale_y, src_h, false)) {
    static const bool allow_wide = !(getenv("RTM_LETTERBOX_IMPL") && strcmp(getenv("RTM_LETTERBOX_IMPL"), "narrow") == 0);
    if (allow_wide && out_dtype != RTM_F32 && out_w % kWidePix == 0 && left % kWidePix == 0 && new_w % kWidePix == 0 &&
        (reinterpret_cast<uintptr_t>(frames) & 31) == 0 && row_stride % 32 == 0 && frame_stride % 32 == 0) {
      const int wgroups = (out_w / kWidePix) * out_h;
      dim3 wgrid((wgroups + 127) / 128, num_streams);
      if (out_dtype == RTM_F16) letterbox_decimate3_wide_kernel<__half><<<wgrid, 128, 0, s>>>(a);
      else if (out_dtype == RTM_BF16) letterbox_decimate3_wide_kernel<__nv_bfloat16><<<wgrid, 128, 0, s>>>(a);
      else RTM_REQUIRE(false, "rtm_letterbox: unknown out_dtype %d", out_dtype);
      RTM_LAUNCH_CHECK("letterbox_decimate3_wide_kernel");
      return RTM_OK;
    }
    const int dgroups = (out_w / kDecPix) * out_h;
    dim3 dgrid((dgroups + 255) / 256, num_streams);
    switch (out_dtype) {
      case RTM_F32:
        letterbox_decimate3_kernel<float><<<dgrid, 256, 0, s>>>(a);
        break;
      case RTM_F16:
        letterbox_decimate3_kernel<__half><<<dgrid, 256, 0, s>>>(a);
        break;
      case RTM_BF16:
        letterbox_decimate3_kernel<__nv_bfloat16><<<dgrid, 256, 0, s>>>(a);
        break;
      default:
        RTM_REQUIRE(false, "rtm_letterbox: unknown out_dtype %d", out_dtype);
    }
    RTM_LAUNCH_CHECK("letterbox_decimate3_kernel");
    return RTM_OK;
  }
  // tile kernel (RTM_LETTERBOX_IMPL=pixels turns it off): needs rows that start on 4-byte boundaries
  static const bool allow_tile = !(getenv("RTM_LETTERBOX_IMPL") && strcmp(getenv("RTM_LETTERBOX_IMPL"), "pixels") == 0);
  if (allow_tile && (reinterpret_cast<uintptr_t>(frames) & 3) == 0 && row_stride % 4 == 0 && frame_stride % 4 == 0 &&
      row_stride < (1ll << 31) && row_stride >= 12) {
    dim3 tgrid((out_w + kTileCols - 1) / kTileCols, (out_h + kTileRows - 1) / kTileRows, num_streams);
    switch (out_dtype) {
      case RTM_F32:
        letterbox_tile_kernel<float><<<tgrid, 256, 0, s>>>(a);
        break;
      case RTM_F16:
        letterbox_tile_kernel<__half><<<tgrid, 256, 0, s>>>(a);
        break;
      case RTM_BF16:
        letterbox_tile_kernel<__nv_bfloat16><<<tgrid, 256, 0, s>>>(a);
        break;
      default:
        RTM_REQUIRE(false, "rtm_letterbox: unknown out_dtype %d", out_dtype);
    }
    RTM_LAUNCH_CHECK("letterbox_tile_kernel");
    return RTM_OK;
  }
  switch (out_dtype) {
    case RTM_F32:
      letterbox_kernel<float><<<grid, 256, 0, s>>>(a);
      break;
    case RTM_F16:
      letterbox_kernel<__half><<<grid, 256, 0, s>>>(a);
      break;
    case RTM_BF16:
      letterbox_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(a);
      break;
    default:
      RTM_REQUIRE(false, "rtm_letterbox: unknown out_dtype %d", out_dtype);
  }
  RTM_LAUNCH_CHECK("letterbox_kernel");
  return RTM_OK;
}

}  // namespace

extern "C" int rtm_letterbox(const uint8_t* frames, int32_t num_streams, int32_t src_h, int32_t src_w,
                             int64_t row_stride, int64_t frame_stride, void* out, int32_t out_dtype,
                             int32_t out_h, int32_t out_w, rtm_cuda_stream stream) {
  RTM_REQUIRE(src_h > 0 && src_w > 0 && out_h > 0 && out_w > 0, "rtm_letterbox: bad shape");
  // LetterBox.__call__ (auto=False, scaleup=True, center=True): r = min(H/h0, W/w0);
  // new_unpad = round(w0*r), round(h0*r); dw, dh = (W - new_w)/2, (H - new_h)/2;
  // top = round(dh - 0.1), left = round(dw - 0.1)
  const double r = fmin(static_cast<double>(out_h) / src_h, static_cast<double>(out_w) / src_w);
  const int new_w = static_cast<int>(nearbyint(src_w * r));  // Python round(): half to even
  const int new_h = static_cast<int>(nearbyint(src_h * r));
  const int left = static_cast<int>(nearbyint((out_w - new_w) / 2.0 - 0.1));
  const int top = static_cast<int>(nearbyint((out_h - new_h) / 2.0 - 0.1));
  return launch_letterbox(frames, num_streams, src_h, src_w, row_stride, frame_stride, out, out_dtype, out_h, out_w, new_h,
                          new_w, top, left, stream);
}

extern "C" int rtm_letterbox_ex(const uint8_t* frames, int32_t num_streams, int32_t src_h, int32_t src_w,
                                int64_t row_stride, int64_t frame_stride, void* out, int32_t out_dtype,
                                int32_t out_h, int32_t out_w, int32_t new_h, int32_t new_w, int32_t top, int32_t left,
                                rtm_cuda_stream stream) {
  return launch_letterbox(frames, num_streams, src_h, src_w, row_stride, frame_stride, out, out_dtype, out_h, out_w, new_h,
                          new_w, top, left, stream);
}
