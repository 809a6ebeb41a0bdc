/*
 * mmt_b200.h - C ABI of libmmt_b200.so: the sm_100a kernels behind the per-frame network forward of the
 * MixViT RGB / RGB-T trackers (reference: LZ-QWQ/Multi-modal-Tracking).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise;
 *   - caller allocates every output and workspace; no hidden allocation, no hidden synchronisation;
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream;
 *   - return value: 0 on success, a cudaError_t value for launch failures, or MMT_ERR_* (>= 1000001)
 *     for rejected arguments.  Nothing is printed.  (The reference's native ops raise through
 *     AT_ASSERTM / only printf launch errors: external/PreciseRoIPooling/pytorch/prroi_pool/src/
 *     prroi_pooling_gpu.c:42, lib/models/mixformer_vit_rgbt/deformable_attention/ops/src/cuda/
 *     ms_deform_im2col_cuda.cuh:948-952; the Python wrapper turns a non-zero status into RuntimeError.)
 *   - "T" buffers are bf16 when `bf16 != 0` and fp32 otherwise (the fp32 mode is the parity mode).
 *   - all token / pixel tensors are row-major [rows, channels] (NHWC for feature maps).
 *
 * Each entry point cites the reference code whose arithmetic it replaces (paths relative to the
 * reference root).
 */
#ifndef MMT_B200_H
#define MMT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMT_ACT_NONE 0
#define MMT_ACT_GELU 1 /* erf GELU (nn.GELU default), timm Mlp: lib/models/mixformer_vit/mixformer.py:123 */
#define MMT_ACT_RELU 2 /* head conv+BN+ReLU lib/models/mixformer_cvt/head.py:7-20; fusion FFN */

/* Version / capability probe (host only). Returns the ABI version; *sm gets the compiled SM (100). */
int mmt_abi_version(int* sm);

/* Programmatic dependent launch (host only; returns the previous setting).  The GEMM, attention and LayerNorm-statistics
 * kernels run their prologue (barrier initialisation, tensor-memory allocation, descriptor prefetch) before a
 * griddepcontrol.wait and are launched with cudaLaunchAttributeProgrammaticStreamSerialization, so that prologue overlaps
 * the tail of the previous kernel of the stream; results are unaffected.  enable = 0 launches them plainly (A/B). */
int mmt_config_pdl(int enable);

/* Clusters of four CTAs for the big backbone GEMMs (host only; returns the previous setting): two CTA pairs work on two
 * 256-row blocks of the same 256-column tile and share the weight tile by TMA multicast (3/4 of the L2 -> SM operand bytes
 * of the CTA-pair launch).  Same arithmetic per output element - results are bit-identical.  OFF by default (measured
 * 4-15 % slower per launch on B200: clusters of four leave SMs idle and the per-SM operand ingest is unchanged). */
int mmt_config_cluster4(int enable);

/* SM budget of the small-GEMM tile rule (host only; returns the previous setting; 0 = the whole GPU).  Small problems (one
 * sequence: 452 rows) get the narrowest tile whose grid still fits ONE wave of `sms` CTAs.  A caller that runs two independent
 * chains of such GEMMs on two streams (the two modality backbones) sets half the SM count so that the chains run side by side
 * instead of alternating - every CTA of the GEMM kernel owns its SM.  Tile width never changes a result (same K order). */
int mmt_config_small_gemm_sms(int sms);

/*
 * out[M,N] = act(A[M,K] * W[N,K]^T + bias[N]) + rowadd[row % period][N] + resid[M,N]
 * A, W bf16 row-major (K contiguous, lda/ldw in elements, multiples of 8, 16-byte aligned base);
 * bias / rowadd / resid fp32 or NULL; out bf16 (out_fp32 = 0) or fp32, may alias resid.
 * tcgen05 / TMEM / TMA kernel.  max_ctas <= 0 means "all SMs".
 * Replaces nn.Linear / 1x1 nn.Conv2d / (with mmt_im2col3x3) 3x3 nn.Conv2d+BN on the path:
 *   qkv, proj, fc1, fc2      lib/models/mixformer_vit/mixformer.py:45-47,56,74,123
 *   patch-embed projection   lib/models/mixformer_vit/mixformer.py:25-33,192-203 (rowadd = pos_embed)
 *   fusion projections       lib/models/mixformer_vit_rgbt/deformable_attention/deformable_encoder_lnspecific.py:127-160
 *   head convs (BN folded)   lib/models/mixformer_cvt/head.py:7-20,159-198
 */
int mmt_gemm_bf16(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, int act,
                  const float* resid, int ldr, const float* rowadd, int rowadd_period, void* out, int ldo,
                  int out_fp32, int max_ctas, void* stream);

/*
 * mmt_gemm_bf16 with the transformer block's LayerNorm folded into the two GEMMs around it, so that the normalised
 * rows never exist in HBM (Block.forward lib/models/mixformer_vit/mixformer.py:126-129: x + attn(norm1(x)),
 * x + mlp(norm2(x)); modality-specific norms mixformer_shared.py:143-159 = one call per modality's row range).
 * With LN(x) = (x - mu) * rs * gamma + beta, Linear(LN(x)) = rs * (x W'^T - mu * colsum) + bias' where
 * W' = W * diag(gamma), colsum[n] = sum_k W'[n,k], bias' = bias + W beta are prepared once on the host:
 *   consumer side (ln_stats != NULL; bf16 output): A holds the RAW residual rows in bf16, W / bias are W' / bias';
 *     ln_stats [ln_slots][ln_stride][2] fp32 (slot-major; ln_stride >= M rows between slots, row 0 = first row of A) =
 *     per-row partial (sum, sum of squares) over the K columns, summed here in slot order (deterministic); the epilogue
 *     computes out = rs * (acc - mu * colsum[n]) + bias[n], then `act`.
 *   producer side (xb_out, stats_out != NULL; fp32 output, N % 128 == 0, ldo == N): besides out (= act(AW^T + bias) +
 *     resid) the call leaves xb_out [M, ld_xb] = bf16(out) and stats_out [N/128][stats_stride][2] = partial (sum, sum of
 *     squares) of the out rows (one slot per 128 columns, same slot-major layout), i.e. what the next consumer needs.  Fused into the epilogue of the CTA-pair kernel; for launch
 *     shapes that take another kernel the same outputs are produced by mmt_rowstats_cast after the GEMM.
 * Everything else as mmt_gemm_bf16 (which is this function with the four extra pointers NULL).
 */
int mmt_gemm_bf16_ex(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, int act,
                     const float* resid, int ldr, const float* rowadd, int rowadd_period, void* out, int ldo,
                     int out_fp32, int max_ctas, const float* ln_stats, int ln_slots, int ln_stride, float ln_eps,
                     const float* colsum, void* xb_out, int ld_xb, float* stats_out, int stats_stride, void* stream);

/* Stand-alone producer of the folded LayerNorm: xb [rows, ld_xb] = bf16(x), stats [slots = C/128][slot_stride][2] = the
 * partial (sum, sum of squares) of the fp32 row per 128-column slot, accumulated in the same order as the fused GEMM
 * epilogue accumulates them (bit-identical statistics whichever producer ran). */
int mmt_rowstats_cast(const float* x, int rows, int C, void* xb, int ld_xb, float* stats, int slots, int slot_stride,
                      void* stream);

/* Same contract, fp32 operands and fp32 FMA accumulation (parity mode; SIMT kernel). */
int mmt_gemm_f32(const float* A, int lda, const float* W, int ldw, int M, int N, int K, const float* bias, int act,
                 const float* resid, int ldr, const float* rowadd, int rowadd_period, float* out, int ldo,
                 void* stream);

/*
 * Patch matrix of a Conv2d(Cin, C, P, stride P): img fp32 NCHW [B,Cin,H,W] -> rows of length Cin*P*P
 * (k = c*P*P + ky*P + kx == weight.view(C,-1) order) written at row b*tok_per_seq + tok_off + patch.
 * Reference: PatchEmbed.forward lib/models/mixformer_vit/mixformer.py:28-33 (the conv itself is then
 * mmt_gemm_* with rowadd = pos_embed, :192-203).
 */
int mmt_patchify(const float* img, void* out, int B, int Cin, int H, int W, int P, int tok_off, int tok_per_seq,
                 int out_bf16, void* stream);

/*
 * LayerNorm over the last dim of fp32 rows [rows, C]; parameter set (g1,b1) is used for rows with
 * ((row / period) & 1) == 1 (period <= 0 or g1 == NULL: always (g0,b0)).  Writes an fp32 and/or a bf16 copy.
 * Reference: nn.LayerNorm(eps=1e-6) in Block.forward lib/models/mixformer_vit/mixformer.py:126-129,259;
 * modality-specific norms lib/models/mixformer_vit_rgbt/mixformer_shared.py:143-157 and
 * deformable_attention/deformable_encoder_lnspecific.py:143-160 (eps 1e-5).
 */
int mmt_layernorm(const float* x, int rows, int C, float eps, const float* g0, const float* b0, const float* g1,
                  const float* b1, int period, float* out_f32, void* out_bf16, void* stream);

/*
 * GroupNorm(G, C) of fp32 NHWC rows [B, HW, C] (statistics over HW x C/G per sample and group).
 * Output row of (b, r) is b*out_seq_rows + out_row_off + r (out_seq_rows <= 0: dense, = HW), so that one modality's
 * map lands in its half of the [B, 2*HW, C] fusion token tensor.
 * Reference: nn.GroupNorm(32, .) after the fusion 1x1 convs lib/models/mixformer_vit_rgbt/fusion_utils.py:252-268.
 */
int mmt_groupnorm(const float* x, int B, int HW, int C, int G, float eps, const float* gamma, const float* beta,
                  float* out_f32, void* out_bf16, int out_seq_rows, int out_row_off, void* stream);

/*
 * ConvMAE stem (lib/models/mixformer_convmae/mixformer_online.py).  mmt_layernorm_act: channel LayerNorm of fp32 NHWC
 * rows + optional exact GELU (PatchEmbed.forward :45-50; CBlock norms :172-189 with gelu = 0); seg_rows > 0 re-maps
 * output row r to (r / seg_rows) * out_seq_rows + out_row_off + r % seg_rows (token order [t | ot | s]).
 * mmt_patchify2x2: rows of a Conv2d(C, E, 2, stride 2) patch matrix, k = (ky*2 + kx)*C + c.  mmt_dwconv5x5: depthwise
 * Conv2d(E, E, 5, padding 2, groups E) + bias (CBlock.attn :172), weights fp32 [25, E] tap-major.
 */
int mmt_layernorm_act(const float* x, int rows, int C, float eps, const float* gamma, const float* beta, int gelu,
                      float* out_f32, void* out_bf16, int seg_rows, int out_seq_rows, int out_row_off, void* stream);
int mmt_patchify2x2(const float* x, int B, int H, int W, int C, void* out, int out_bf16, void* stream);
int mmt_dwconv5x5(const void* in, const float* w, const float* bias, int B, int H, int W, int E, void* out, int is_bf16,
                  void* stream);

/* dst[s*rows_per_seq + r, :] = T(src[s*seq_stride + row_off + r, :]): e.g. the search tokens of every sequence
 * out of the [template, online template, search] token layout (torch.split, mixformer.py:208). */
int mmt_copy_rows(const float* src, int seq_stride, int row_off, int rows_per_seq, int nseq, int C, void* dst,
                  int out_bf16, void* stream);

/*
 * Staging of the bimodal deformable-attention inputs from the fusion tokens src fp32 [B, 2L, C]
 * (first L tokens RGB, last L TIR):  out_val[B*2L, C] = T(src);  out_q[B*L, 2C] = T(cat_c(src_v + pos_v, src_i + pos_i)).
 * Reference: query_bimodal / with_pos_embed, ms_deform_attn_bimodal.py:97-99, deformable_encoder_lnspecific.py:151.
 */
int mmt_fusion_prep(const float* src, const float* pos, int B, int L, int C, void* out_val, void* out_q, int out_bf16,
                    void* stream);

/*
 * MultiScaleDeformableAttention forward, reference tensor layout (value [N,S,M,D] T; sampling_loc
 * [N,Lq,M,L,P,2] fp32 in [0,1]; attn_weight [N,Lq,M,L,P] fp32; out [N,Lq,M*D] T).  level_hw_host is a HOST array
 * of L (H,W) pairs.  Replaces MSDA.ms_deform_attn_forward (deformable_attention/ops/src/cuda/
 * ms_deform_attn_cuda.cu:20-80, ms_deform_im2col_cuda.cuh:237-299); no batch % im2col_step restriction.
 */
int mmt_msda_fwd(const void* value, const int* level_hw_host, const float* sampling_loc, const float* attn_weight,
                 void* out, int N, int S, int M, int D, int L, int Lq, int P, int is_bf16, void* stream);

/*
 * Fused bimodal form used in the RGB-T fusion encoder: value T [B, 2*H*W, M*64]; offw fp32 [B*H*W, ld] holding the
 * raw projection output [M*2*P*2 sampling offsets | M*2*P attention logits] per position; reference points,
 * offset normalisation, softmax over the 2*P samples and sampling are fused; the (identical) result is written for
 * the RGB and the TIR query rows.  Reference: MSDeformAttn_Bimodal.forward ms_deform_attn_bimodal.py:83-130.
 */
int mmt_msda_bimodal_fwd(const void* value, const float* offw, int ld_offw, void* out, int B, int H, int W, int M,
                         int D, int P, int is_bf16, void* stream);

/*
 * im2col of Conv2d(k=3, pad=1) on NHWC maps with nearest upsampling (+ optional add of a second map) folded in:
 * in(b,y,x,:) = src1[b, y/s1, x/s1, :] (+ src2[b, y/s2, x/s2, :]);  out[(b*H+y)*W+x, (ky*3+kx)*C + c].
 * Reference: F.interpolate + add + conv in Pyramid_Corner_Predictor.get_score_map lib/models/mixformer_cvt/head.py:159-198.
 */
int mmt_im2col3x3(const void* src1, int ld1, int s1, const void* src2, int ld2, int s2, int B, int H, int W, int C,
                  void* out, int is_bf16, void* stream);

/*
 * Conv2d(C -> N, k=3, pad=1) + bias + activation on a bf16 NHWC map in [B*H*W, ld_in] as an IMPLICIT GEMM on the
 * tensor cores: the K loop walks 9 taps x channel chunks and every A slice is one 4-D TMA box of the map shifted by
 * the tap offset (out-of-image pixels arrive as zeros = the zero padding); no im2col matrix exists.  Wt [N, 9*C]
 * holds the filter as (ky, kx, c) with eval-BatchNorm folded in.  out [B*H*W, ldo] bf16 or fp32.
 * Reference: conv() lib/models/mixformer_cvt/head.py:7-20 as used in get_score_map :159-198.
 */
int mmt_conv3x3_bf16(const void* in, int ld_in, int B, int H, int W, int C, const void* Wt, int ldw, int N,
                     const float* bias, int act, void* out, int ldo, int out_fp32, void* stream);

/* out(b,y,x,:) = src1[b, y/s1, x/s1, :] (+ src2[b, y/s2, x/s2, :]), bf16 NHWC, dense out [B*H*W, C]:
 * F.interpolate(scale_factor=2|4) + add of the pyramid head (head.py:166-178), the input of the next 3x3 conv. */
int mmt_upsample_add(const void* src1, int ld1, int s1, const void* src2, int ld2, int s2, int B, int H, int W, int C,
                     void* out, void* stream);

/*
 * Corner decode for both corners: score = conv5_1x1(x4) + up4(a3) + up2(a4), softmax over S*S, soft-argmax,
 * xyxy / img_sz and box_xyxy_to_cxcywh.  score_maps (fp32 [B,2,S*S], raw logits) may be NULL.  a3_* / a4_* NULL: the
 * plain Corner_Predictor (score = conv5_1x1(x4) only, head.py:23-94).
 * Reference: head.py:181,198-212 (coords :138-145), lib/utils/box_ops.py:27-31, forward_box_head mixformer.py:325-338.
 */
int mmt_corner_decode(const void* x4_tl, const void* x4_br, int ld4x, int C4, const float* w5_tl, const float* w5_br,
                      float b5_tl, float b5_br, const void* a3_tl, const void* a3_br, int lda3, const void* a4_tl,
                      const void* a4_br, int lda4, int B, int S, float stride_px, float img_sz, float* score_maps,
                      float* xyxy, float* cxcywh, int is_bf16, void* stream);

/*
 * Asymmetric mixed attention over packed qkv rows [rows, 3C] (q | k | v, head-major 64-wide slices).
 * tiles_dev: DEVICE array of n_tiles records of 16 int32:
 *   {q_row0, q_rows(<=128), out_row0, nseg(<=3), k_row0[3], k_len[3], k_buf[3], pad[3]}; key segment i reads rows
 *   [k_row0[i], +k_len[i]) of qkv0 (k_buf 0, rows0 rows) or qkv1 (k_buf 1, rows1 rows, e.g. cached template K/V).
 *   max_keys = largest total key count of any tile (fp32 mode sizing).  out T [.., ldo], head h at column h*64.
 *   bf16 mode: tcgen05 / TMEM kernel fed by TMA straight from the qkv buffer (csrc/attention_tc.cu).
 * Reference: Attention.forward / forward_test lib/models/mixformer_vit/mixformer.py:51-93; cross-modal
 * lib/models/mixformer_vit_rgbt/asymmetric_shared.py:55-104, asymmetric_shared_ce.py:146-200.
 */
int mmt_mixattn_fwd(const void* qkv0, int rows0, const void* qkv1, int rows1, int ld, int C, int heads,
                    const int* tiles_dev, int n_tiles, int max_keys, void* out, int ldo, float scale, int is_bf16,
                    void* stream);

/*
 * Candidate-elimination scores [B, 2*Ls] = mean over heads of mean over the 2*Lt template rows of
 * softmax_{2Ls}([q_mt_V; q_mt_I] [k_s_V; k_s_I]^T * scale), from the modality-major qkv buffer [2B, n_tok, 3C].
 * partial_ws: fp32 workspace [B * heads * (2*Lt/32) * 2*Ls].  fp32 arithmetic in both modes, deterministic.
 * Reference: asymmetric_shared_ce.py:202-205 (attn_t2s) and :91-92 (mean(dim=2).mean(dim=1)).
 */
int mmt_ce_scores(const void* qkv, int ld, int C, int heads, int B, int n_tok, int Lt, int Ls, float scale,
                  float* partial_ws, float* scores, int is_bf16, void* stream);

/*
 * The same scores with the template rows (queries) and the search rows (keys) in two buffers of the same row stride:
 * qbuf [2B, q_seq_rows, 3C] holds every sequence-modality's template rows first in its block, qkv [2B, n_tok, 3C] the
 * search rows from row k_row_off of its block.  mmt_ce_scores is the call with one buffer (q_seq_rows = n_tok,
 * k_row_off = Lt); the cached-template path (template q/k/v computed once per template update, search tokens only per
 * frame) passes the template cache and the search-only buffer (k_row_off = 0).  Same kernels, same summation order.
 */
int mmt_ce_scores_split(const void* qbuf, int q_seq_rows, const void* qkv, int n_tok, int k_row_off, int ld, int C,
                        int heads, int B, int Lt, int Ls, float scale, float* partial_ws, float* scores, int is_bf16,
                        void* stream);

/*
 * Per (sequence, modality) descending sort of the Ls scores; order int32 [2B, Ls] (sorted local indices),
 * gidx_keep fp32 [2B, keep] / gidx_removed fp32 [2B, Ls-keep] = gidx_in gathered in score order.  All [2B, .]
 * tensors are modality-major (row m*B + b).  Reference: get_token_from_attn asymmetric_shared_ce.py:22-46.
 */
int mmt_ce_topk(const float* scores, int B, int Ls, int keep, const float* gidx_in, float* gidx_keep,
                float* gidx_removed, int* order, void* stream);

/* x_out[s, :Lt] = x[s, :Lt]; x_out[s, Lt+i] = x[s, Lt + order[s][i]], i < keep (tokens.gather, :41-44). */
int mmt_ce_gather_tokens(const float* x, int nseq, int n_tok, int Lt, const int* order, int Ls, int keep, float* x_out,
                         int C, void* stream);

/* out[s*Ls0 + int(gidx[s][i]), :] = T(x[s, Lt+i, :]), zeros elsewhere (_recover_search :427-447). */
int mmt_ce_recover(const float* x, int nseq, int n_tok, int Lt, const float* gidx, int Lk, int Ls0, void* out, int C,
                   int out_bf16, void* stream);

/* out[r, :C] = a[r], out[r, C:] = b[r]: channel concatenation of two NHWC maps (RGBT_Fusion_Cat, fusion_utils.py:106). */
int mmt_concat_cols(const void* a, const void* b, int rows, int C, void* out, int is_bf16, void* stream);

/* rois[b] = (b, xyxy[b] * scale), fp32 [B,5]: the SPM's target_roi (score_decoder.py:37-44). */
int mmt_spm_rois(const float* xyxy, int B, float scale, float* rois, void* stream);

/*
 * Precise RoI pooling forward: rois fp32 [R,5] = (batch_idx, x0, y0, x1, y1).  channels_last = 0: feat NCHW,
 * out [R,C,PH,PW] (the reference op's layout, prroi_pooling_gpu.c:22-44); channels_last = 1: feat NHWC, out
 * [R, PH*PW, C] (token layout used by the SPM head).  Kernel restated from prroi_pooling_gpu_impl.cu:37-106,149-212.
 */
int mmt_prroi_fwd(const float* feat, const float* rois, float* out, int R, int C, int H, int W, int PH, int PW,
                  float spatial_scale, int channels_last, void* stream);

/*
 * Frame side of the per-frame loop, batched on the device.
 *
 * mmt_frame_crop: for n_mod x B uint8 HWC (3-channel) frames - image m*B + b; frames_dev is a DEVICE array of device
 * pointers, dims_dev a DEVICE int32 [n_mod*B][3] = (H, W, row pitch in bytes) - extract the square window of side
 * ceil(sqrt(w*h) * factor) centred on the sequence's box (state_dev: DEVICE float64 [B][4] = x, y, w, h, shared by the
 * modalities), zero-pad it, resize it to out_sz x out_sz with OpenCV's fixed-point INTER_LINEAR, apply the JET colour
 * map for the modalities whose bit is set in jet_mask (jet_lut_dev: DEVICE uint8 [256][3]), and write
 * ((x/255) - mean) / std as fp32 [n_mod][B][3][S][S] (out) and/or the uint8 crop [n_mod][B][S][S][3] (out_u8).
 * resize_factor_dev (DEVICE float64 [B], may be NULL) receives out_sz / crop side.  active_dev (DEVICE uint8 [B] or NULL)
 * selects the sequences that are processed; the others keep their previous outputs.  A degenerate box (side < 1, where
 * the reference raises "Too small bounding box.") yields an all-zero crop.  workspace: caller-allocated, see
 * mmt_frame_crop_workspace_bytes; two launches (window geometry + tap tables, then the gather).
 * Replaces sample_target lib/train/data/processing_utils.py:15-83 (cv.copyMakeBorder + cv.resize) and
 * Preprocessor_Multimodal.process lib/test/tracker/tracker_utils.py:37-48 (cv2.applyColorMap, normalisation, H2D).
 *
 * mmt_track_update: state <- clip_box(map_box_back(pred * search_size / resize_factor)), margin in pixels, frame sizes
 * from dims_dev rows 0..B-1; the new state is also written to log_dev (float64 [B][4], may be NULL: e.g. row t of the
 * sequence's result table).  Replaces the host arithmetic and the per-frame `.tolist()` synchronisation of
 * MixFormer.track lib/test/tracker/asymmetric_shared_ce.py:99-103,134-140 and clip_box lib/utils/box_ops.py:155-164.
 */
int mmt_frame_crop(const void* const* frames_dev, const int* dims_dev, const double* state_dev,
                   const unsigned char* active_dev, int B, int n_mod, unsigned jet_mask, double factor, int out_sz,
                   const unsigned char* jet_lut_dev, float* out, unsigned char* out_u8, double* resize_factor_dev,
                   void* workspace, long long workspace_bytes, void* stream);
/* bytes of 16-byte aligned DEVICE workspace mmt_frame_crop needs (per-image window geometry + resize tap tables) */
long long mmt_frame_crop_workspace_bytes(int B, int n_mod, int out_sz);
/*
 * Preprocessor on the device: uint8 HWC crops [n_img, size, size, 3] -> fp32 [n_img, 3, size, size] = ((x/255) - mean)/std;
 * image i belongs to modality i / per_mod, modalities whose bit is set in jet_mask get cv2.applyColorMap(JET) first
 * (jet_lut_dev: DEVICE uint8 [256][3]).  Replaces Preprocessor_wo_mask.process / Preprocessor_Multimodal.process
 * lib/test/tracker/tracker_utils.py:24-48 after their `torch.tensor(img_arr).cuda()` upload.
 */
int mmt_preprocess_u8(const unsigned char* crops_u8, float* out, int n_img, int size, int per_mod, unsigned jet_mask,
                      const unsigned char* jet_lut_dev, void* stream);

/*
 * Online (SPM) trackers: score bookkeeping of the online-template candidate, per sequence:
 *   s = sigmoid(logits[b]) (fp32); max_score[b] *= decay; take[b] = s > 0.5 && s > max_score[b]; if take: max_score[b] = s.
 * take_dev (DEVICE uint8 [B]) is then the `active` mask of the mmt_frame_crop call that refreshes the candidate crop.
 * Replaces MixFormerOnline.track lib/test/tracker/mixformer_convmae_online.py:99,105-113 (pred_score.item() host sync).
 */
int mmt_online_score_update(const float* logits, double* max_score_dev, unsigned char* take_dev,
                            const unsigned char* active_dev, int B, double decay, void* stream);
int mmt_track_update(const float* pred_cxcywh, const double* resize_factor_dev, const int* dims_dev, double* state_dev,
                     double* log_dev, const unsigned char* active_dev, int B, int search_size, double margin,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMT_B200_H */
