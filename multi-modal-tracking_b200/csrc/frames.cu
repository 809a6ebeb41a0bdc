// Per-frame glue around the network forward, on the device.
//
// The reference does this on the host, one sequence at a time, and synchronises the GPU every frame:
//   sample_target            lib/train/data/processing_utils.py:15-83   crop window, zero padding, cv.resize
//   Preprocessor_Multimodal  lib/test/tracker/tracker_utils.py:37-48    JET colour map (infrared), /255, mean/std, CHW
//   MixFormer.track          lib/test/tracker/asymmetric_shared_ce.py:99-103,134-140   pred box -> frame coordinates
//   clip_box                 lib/utils/box_ops.py:155-164
// Here B sequences advance together: the tracker state [B,4] (float64, like the Python floats it replaces) lives in
// HBM, one kernel turns the uint8 frames into the normalised fp32 crops the backbone embeds, one kernel folds the
// predicted boxes back into the state.  No host round trip per frame.
//
// Bit-exactness: the crop window is float64 arithmetic with round-half-even (`round()`), the resize is OpenCV's
// fixed-point INTER_LINEAR for uint8 (resize.cpp: 11-bit coefficient pairs from float32 fractions, horizontal pass in
// int32, vertical pass ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2), BGR2GRAY is its 15-bit fixed point; the
// uint8 crops therefore equal the reference's byte for byte (tests/test_frames_gpu.py, oracle/frame_oracle.py).
// Both are HBM/L2-bound gathers.  The crop runs as two launches: a tiny one per image for the window geometry and the
// S tap entries (all the float64 work), then the gather - thread = output column, 8 rows per block, the three channel
// planes written coalesced.
#include "common.cuh"
#include "../../include/mmt_b200.h"

namespace mmt {

struct CropGeom {
  int crop_sz, x1, y1, xa, xb, ya, yb;   // window origin and the frame range [xa,xb) x [ya,yb) that is copied
  int H, W, pitch;
  int valid;
  double scale;                          // 1 / (S / crop_sz), as cv::resize computes it
};

// processing_utils.py:31-49, float64 with explicitly rounded operations (no contraction)
__device__ inline void crop_geometry(const double* st, double factor, int H, int W, CropGeom& g) {
  const double x = st[0], y = st[1], w = st[2], h = st[3];
  const double side = __dmul_rn(sqrt(__dmul_rn(w, h)), factor);
  const double c = ceil(side);
  g.valid = (c >= 1.0 && c < 65536.0) ? 1 : 0;      // "Too small bounding box." in the reference; NaN fails both tests
                                                    // (tap tables hold 16-bit source indices)
  g.crop_sz = g.valid ? static_cast<int>(c) : 1;
  const double half = __dmul_rn(static_cast<double>(g.crop_sz), 0.5);
  g.x1 = static_cast<int>(rint(__dsub_rn(__dadd_rn(x, __dmul_rn(0.5, w)), half)));    // round(): half to even
  g.y1 = static_cast<int>(rint(__dsub_rn(__dadd_rn(y, __dmul_rn(0.5, h)), half)));
  const int x2 = g.x1 + g.crop_sz, y2 = g.y1 + g.crop_sz;
  g.xa = max(g.x1, 0);
  g.xb = x2 - max(x2 - W + 1, 0);                   // the reference's "+ 1": the last column is dropped with the overhang
  g.ya = max(g.y1, 0);
  g.yb = y2 - max(y2 - H + 1, 0);
  g.H = H; g.W = W;
}

// source index and 11-bit weight pair of destination index d (resize.cpp set-up loop, horizontal flavour when
// `horizontal`: negative index -> (0, weight 0); last column -> single tap)
__device__ __forceinline__ void linear_tap(int d, double scale, int ssize, bool horizontal, int& s0, int& s1, int& w0,
                                           int& w1) {
  float f = static_cast<float>(__dsub_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), 0.5));
  int s = static_cast<int>(floorf(f));
  f = __fsub_rn(f, static_cast<float>(s));
  if (horizontal) {
    if (s < 0) { f = 0.f; s = 0; }
    if (s + 1 >= ssize) {                           // dx >= xmax: D[dx] = S[sx] * ONE
      s0 = s1 = min(s, ssize - 1);
      w0 = 2048; w1 = 0;
      return;
    }
    s0 = s; s1 = s + 1;
  } else {
    s0 = min(max(s, 0), ssize - 1);
    s1 = min(max(s + 1, 0), ssize - 1);
  }
  w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));      // saturate_cast<short>(cvRound(.)); |.| <= 2048
  w1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

// Tap tables: entry d of image `img` = {s0 | s1 << 16, w0 | w1 << 16} for the horizontal pass followed by the same
// for the vertical pass (the crop is square and both passes share the scale, but their border rules differ).
struct TapPair { uint32_t hs, hw, vs, vw; };

// one block per image: window geometry (thread 0), then the S tap entries
__global__ void __launch_bounds__(128)
frame_geom_kernel(const int* __restrict__ dims, const double* __restrict__ state, const uint8_t* __restrict__ active,
                  int B, double factor, int S, CropGeom* __restrict__ geom, TapPair* __restrict__ taps,
                  double* __restrict__ resize_factor, const uint8_t* __restrict__ jet_lut, float* __restrict__ norm_tab) {
  const int img = blockIdx.x;
  const int b = img % B;
  if (img == 0) {
    // ((v / 255) - mean) / std for every byte value, fp32 true divisions (tracker_utils.py:44-47): [c][v], followed by
    // the same through the JET table ([c][gray]) - the gather kernel turns two divisions per value into one lookup
    const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
      const int c = i >> 8, v = i & 255;
      norm_tab[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), mean[c]), sd[c]);
      if (jet_lut)
        norm_tab[768 + i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(jet_lut[3 * v + c]), 255.0f), mean[c]), sd[c]);
    }
  }
  if (active && !active[b]) return;
  __shared__ CropGeom g;
  if (threadIdx.x == 0) {
    crop_geometry(state + 4 * b, factor, dims[3 * img], dims[3 * img + 1], g);
    g.pitch = dims[3 * img + 2];
    const double inv = __ddiv_rn(static_cast<double>(S), static_cast<double>(g.crop_sz));
    g.scale = __ddiv_rn(1.0, inv);
    if (img < B && resize_factor) resize_factor[b] = inv;       // output_sz / crop_sz
    geom[img] = g;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < S; d += blockDim.x) {
    int s0, s1, w0, w1;
    TapPair t;
    linear_tap(d, g.scale, g.crop_sz, true, s0, s1, w0, w1);
    t.hs = static_cast<uint32_t>(s0) | (static_cast<uint32_t>(s1) << 16);
    t.hw = static_cast<uint32_t>(w0) | (static_cast<uint32_t>(w1) << 16);
    linear_tap(d, g.scale, g.crop_sz, false, s0, s1, w0, w1);
    t.vs = static_cast<uint32_t>(s0) | (static_cast<uint32_t>(s1) << 16);
    t.vw = static_cast<uint32_t>(w0) | (static_cast<uint32_t>(w1) << 16);
    taps[static_cast<size_t>(img) * S + d] = t;
  }
}

// grid (row groups, images); a block walks ROWS output rows, thread = output column (strided when S > blockDim)
constexpr int CROP_ROWS = 8;

__global__ void __launch_bounds__(512)
frame_crop_kernel(const uint8_t* const* __restrict__ frames, const CropGeom* __restrict__ geom,
                  const TapPair* __restrict__ taps, const uint8_t* __restrict__ active, int B, unsigned jet_mask, int S,
                  const uint8_t* __restrict__ jet_lut, const float* __restrict__ norm_tab, float* __restrict__ out,
                  uint8_t* __restrict__ out_u8) {
  const int img = blockIdx.y;              // m * B + b
  const int b = img % B, m = img / B;
  if (active && !active[b]) return;
  __shared__ uint8_t lut[768];
  __shared__ float ntab[768];              // this image's value -> normalised fp32 table ([c][v] or [c][gray])
  const bool jet = (jet_mask >> m) & 1u;
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    ntab[i] = norm_tab[(jet ? 768 : 0) + i];
    if (jet) lut[i] = jet_lut[i];
  }
  __syncthreads();
  const CropGeom g = geom[img];
  const TapPair* __restrict__ tp = taps + static_cast<size_t>(img) * S;
  const uint8_t* __restrict__ base = frames[img];
  const int row0 = blockIdx.x * CROP_ROWS;
  for (int dx = threadIdx.x; dx < S; dx += blockDim.x) {
    const TapPair tx = tp[dx];
    const int sx0 = tx.hs & 0xffff, sx1 = tx.hs >> 16, a0 = tx.hw & 0xffff, a1 = tx.hw >> 16;
    // frame columns of the two taps; outside [xa,xb) the padded crop is zero
    const int fx0 = g.x1 + sx0, fx1 = g.x1 + sx1;
    const bool vx0 = g.valid && fx0 >= g.xa && fx0 < g.xb, vx1 = g.valid && fx1 >= g.xa && fx1 < g.xb;
#pragma unroll 2
    for (int r = 0; r < CROP_ROWS; ++r) {
      const int dy = row0 + r;
      if (dy >= S) break;
      const TapPair ty = tp[dy];             // warp-uniform
      const int sy0 = ty.vs & 0xffff, sy1 = ty.vs >> 16, b0 = ty.vw & 0xffff, b1 = ty.vw >> 16;
      const int fy0 = g.y1 + sy0, fy1 = g.y1 + sy1;
      const bool vy0 = fy0 >= g.ya && fy0 < g.yb, vy1 = fy1 >= g.ya && fy1 < g.yb;
      const uint8_t* r0 = base + static_cast<size_t>(fy0) * g.pitch;
      const uint8_t* r1 = base + static_cast<size_t>(fy1) * g.pitch;
      int v[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int p00 = (vy0 && vx0) ? r0[fx0 * 3 + c] : 0, p01 = (vy0 && vx1) ? r0[fx1 * 3 + c] : 0;
        const int p10 = (vy1 && vx0) ? r1[fx0 * 3 + c] : 0, p11 = (vy1 && vx1) ? r1[fx1 * 3 + c] : 0;
        const int S0 = p00 * a0 + p01 * a1, S1 = p10 * a0 + p11 * a1;
        v[c] = (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
        v[c] = min(max(v[c], 0), 255);
      }
      int key[3] = {v[0], v[1], v[2]};       // index into the normalisation table per channel
      if (jet) {
        // cv2.applyColorMap on a 3-channel image: BGR2GRAY with channel 0 in the 'B' slot, then the 256-entry table
        const int gray = (v[0] * 3735 + v[1] * 19235 + v[2] * 9798 + (1 << 14)) >> 15;
        key[0] = key[1] = key[2] = gray;
        if (out_u8) {
          const uint8_t* e = lut + 3 * gray;
          v[0] = e[0]; v[1] = e[1]; v[2] = e[2];
        }
      }
      const int pix = dy * S + dx;
      if (out_u8) {
        uint8_t* o = out_u8 + (static_cast<size_t>(img) * S * S + pix) * 3;
        o[0] = static_cast<uint8_t>(v[0]); o[1] = static_cast<uint8_t>(v[1]); o[2] = static_cast<uint8_t>(v[2]);
      }
      if (out) {
        float* o = out + static_cast<size_t>(img) * 3 * S * S + pix;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[static_cast<size_t>(c) * S * S] = ntab[c * 256 + key[c]];
      }
    }
  }
}

// asymmetric_shared_ce.py:99-103 + map_box_back :134-140 + clip_box box_ops.py:155-164; one thread per sequence
__global__ void track_update_kernel(const float* __restrict__ pred, const double* __restrict__ resize_factor,
                                    const int* __restrict__ dims, double* __restrict__ state, double* __restrict__ log,
                                    const uint8_t* __restrict__ active, int B, int search_size, double margin) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double st[4] = {state[4 * b], state[4 * b + 1], state[4 * b + 2], state[4 * b + 3]};
  if (!active || active[b]) {
    const double rf = resize_factor[b];
    // pred_boxes.mean(0) * search_size / resize_factor: fp32 tensor arithmetic (one box per sequence)
    const float rff = static_cast<float>(rf), ss = static_cast<float>(search_size);
    double p[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) p[k] = static_cast<double>(__fdiv_rn(__fmul_rn(pred[4 * b + k], ss), rff));
    const double cx_prev = __dadd_rn(st[0], __dmul_rn(0.5, st[2])), cy_prev = __dadd_rn(st[1], __dmul_rn(0.5, st[3]));
    const double half_side = __ddiv_rn(__dmul_rn(0.5, static_cast<double>(search_size)), rf);
    const double cx = __dadd_rn(p[0], __dsub_rn(cx_prev, half_side));
    const double cy = __dadd_rn(p[1], __dsub_rn(cy_prev, half_side));
    double x1 = __dsub_rn(cx, __dmul_rn(0.5, p[2])), y1 = __dsub_rn(cy, __dmul_rn(0.5, p[3]));
    const double w = p[2], h = p[3];
    const double H = static_cast<double>(dims[3 * b]), W = static_cast<double>(dims[3 * b + 1]);
    double x2 = __dadd_rn(x1, w), y2 = __dadd_rn(y1, h);
    x1 = fmin(fmax(0.0, x1), W - margin);
    x2 = fmin(fmax(margin, x2), W);
    y1 = fmin(fmax(0.0, y1), H - margin);
    y2 = fmin(fmax(margin, y2), H);
    st[0] = x1; st[1] = y1;
    st[2] = fmax(margin, __dsub_rn(x2, x1));
    st[3] = fmax(margin, __dsub_rn(y2, y1));
#pragma unroll
    for (int k = 0; k < 4; ++k) state[4 * b + k] = st[k];
  }
  if (log) {
#pragma unroll
    for (int k = 0; k < 4; ++k) log[4 * b + k] = st[k];
  }
}

// lib/test/tracker/mixformer_convmae_online.py:99,105-113 (same in mixformer_vit_online.py): pred_score =
// sigmoid(logit) (fp32), max_pred_score *= decay, the online-template candidate is replaced when
// pred_score > 0.5 and pred_score > max_pred_score.  One thread per sequence; max_score is the float64 Python float.
__global__ void online_score_update_kernel(const float* __restrict__ logits, double* __restrict__ max_score,
                                           uint8_t* __restrict__ take, const uint8_t* __restrict__ active, int B,
                                           double decay) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  uint8_t t = 0;
  if (!active || active[b]) {
    const float sc = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-logits[b])));      // torch.sigmoid, fp32
    const double m = __dmul_rn(max_score[b], decay);
    const double s = static_cast<double>(sc);
    t = (s > 0.5 && s > m) ? 1 : 0;
    max_score[b] = t ? s : m;
  }
  take[b] = t;
}

// Preprocessor_*.process on the device (lib/test/tracker/tracker_utils.py:24-48): uint8 HWC crops [n, S, S, 3] ->
// ((x / 255) - mean) / std as fp32 [n, 3, S, S]; images whose bit is set in jet_bits (bit = image / per_mod) go through
// cv2.applyColorMap(JET) first.  The reference uploads the uint8 crop and converts on the GPU as well
// (`torch.tensor(img_arr).cuda().float()`), so this is its H2D volume: 1 byte per value.
__global__ void __launch_bounds__(256)
preprocess_u8_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, int n_img, int pix_per_img, int per_mod,
                     unsigned jet_mask, const uint8_t* __restrict__ jet_lut) {
  __shared__ float ntab[2][768];
  const float mean[3] = {0.485f, 0.456f, 0.406f}, sd[3] = {0.229f, 0.224f, 0.225f};
  for (int i = threadIdx.x; i < 768; i += blockDim.x) {
    const int c = i >> 8, v = i & 255;
    ntab[0][i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v), 255.0f), mean[c]), sd[c]);
    ntab[1][i] = jet_mask ? __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(jet_lut[3 * v + c]), 255.0f), mean[c]), sd[c])
                          : 0.f;
  }
  __syncthreads();
  const size_t total = static_cast<size_t>(n_img) * pix_per_img;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int img = static_cast<int>(i / pix_per_img);
    const int pix = static_cast<int>(i - static_cast<size_t>(img) * pix_per_img);
    const uint8_t* p = in + i * 3;
    const int v0 = p[0], v1 = p[1], v2 = p[2];
    const bool jet = (jet_mask >> (img / per_mod)) & 1u;
    float* o = out + static_cast<size_t>(img) * 3 * pix_per_img + pix;
    if (jet) {
      const int gray = (v0 * 3735 + v1 * 19235 + v2 * 9798 + (1 << 14)) >> 15;
      o[0] = ntab[1][gray]; o[pix_per_img] = ntab[1][256 + gray]; o[2 * static_cast<size_t>(pix_per_img)] = ntab[1][512 + gray];
    } else {
      o[0] = ntab[0][v0]; o[pix_per_img] = ntab[0][256 + v1]; o[2 * static_cast<size_t>(pix_per_img)] = ntab[0][512 + v2];
    }
  }
}

}  // namespace mmt

extern "C" int mmt_preprocess_u8(const unsigned char* crops_u8, float* out, int n_img, int size, int per_mod,
                                 unsigned jet_mask, const unsigned char* jet_lut_dev, void* stream) {
  MMT_CHECK_ARG(crops_u8 && out && n_img > 0 && size > 0 && per_mod > 0);
  MMT_CHECK_ARG(jet_mask == 0u || jet_lut_dev != nullptr);
  const size_t total = static_cast<size_t>(n_img) * size * size;
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  mmt::preprocess_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(crops_u8, out, n_img, size * size, per_mod,
                                                                                 jet_mask, jet_lut_dev);
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_online_score_update(const float* logits, double* max_score_dev, unsigned char* take_dev,
                                       const unsigned char* active_dev, int B, double decay, void* stream) {
  MMT_CHECK_ARG(logits && max_score_dev && take_dev && B > 0);
  mmt::online_score_update_kernel<<<mmt::cdiv(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      logits, max_score_dev, take_dev, active_dev, B, decay);
  MMT_RETURN_LAST_ERROR();
}

extern "C" long long mmt_frame_crop_workspace_bytes(int B, int n_mod, int out_sz) {
  return static_cast<long long>(B) * n_mod * (sizeof(mmt::CropGeom) + sizeof(mmt::TapPair) * static_cast<long long>(out_sz)) +
         2 * 768 * static_cast<long long>(sizeof(float));
}

extern "C" int mmt_frame_crop(const void* const* frames_dev, const int* dims_dev, const double* state_dev,
                              const unsigned char* active_dev, int B, int n_mod, unsigned jet_mask, double factor,
                              int out_sz, const unsigned char* jet_lut_dev, float* out, unsigned char* out_u8,
                              double* resize_factor_dev, void* workspace, long long workspace_bytes, void* stream) {
  MMT_CHECK_ARG(frames_dev && dims_dev && state_dev && (out || out_u8) && workspace);
  MMT_CHECK_ARG(B > 0 && n_mod > 0 && n_mod <= 8 && out_sz > 0 && out_sz <= 4096 && factor > 0.0);
  MMT_CHECK_ARG(static_cast<long long>(B) * n_mod <= 65535);
  MMT_CHECK_ARG(jet_mask == 0u || jet_lut_dev != nullptr);
  MMT_CHECK_ARG(workspace_bytes >= mmt_frame_crop_workspace_bytes(B, n_mod, out_sz));
  MMT_CHECK_ARG((reinterpret_cast<uintptr_t>(workspace) & 15u) == 0);
  const int n_img = B * n_mod;
  // workspace: [n_img] TapPair tables of out_sz entries (16-byte records), 2 x [3][256] fp32 normalisation tables, then
  // [n_img] CropGeom
  mmt::TapPair* taps = reinterpret_cast<mmt::TapPair*>(workspace);
  float* norm_tab = reinterpret_cast<float*>(taps + static_cast<size_t>(n_img) * out_sz);
  mmt::CropGeom* geom = reinterpret_cast<mmt::CropGeom*>(norm_tab + 2 * 768);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  mmt::frame_geom_kernel<<<n_img, 128, 0, st>>>(dims_dev, state_dev, active_dev, B, factor, out_sz, geom, taps,
                                                resize_factor_dev, jet_mask ? jet_lut_dev : nullptr, norm_tab);
  const dim3 grid(mmt::cdiv(out_sz, mmt::CROP_ROWS), n_img);
  const int threads = out_sz <= 512 ? ((out_sz + 31) / 32) * 32 : 256;     // thread = output column
  mmt::frame_crop_kernel<<<grid, threads, 0, st>>>(reinterpret_cast<const uint8_t* const*>(frames_dev), geom, taps,
                                                   active_dev, B, jet_mask, out_sz, jet_lut_dev, norm_tab, out,
                                                   out_u8);
  MMT_RETURN_LAST_ERROR();
}

extern "C" int mmt_track_update(const float* pred_cxcywh, const double* resize_factor_dev, const int* dims_dev,
                                double* state_dev, double* log_dev, const unsigned char* active_dev, int B,
                                int search_size, double margin, void* stream) {
  MMT_CHECK_ARG(pred_cxcywh && resize_factor_dev && dims_dev && state_dev && B > 0 && search_size > 0);
  mmt::track_update_kernel<<<mmt::cdiv(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      pred_cxcywh, resize_factor_dev, dims_dev, state_dev, log_dev, active_dev, B, search_size, margin);
  MMT_RETURN_LAST_ERROR();
}
