/* detsam2.h — C ABI of libdetsam2.so, the sm_100a kernel library behind the Det-SAM2 / SAM 2.1
 * video-predictor hot path (Hiera encoder -> memory attention -> mask decoder -> memory encoder
 * -> post-processing).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - every function returns 0 on success, a negative DS2_E_* code on a bad argument, or a
 *     positive cudaError_t when a launch fails.  No exceptions cross the ABI;
 *   - activations are token-major ("channels last"): [batch, H*W, C];
 *   - bf16 buffers are passed as void*; f32 as float*.
 *
 * The reference has exactly one FFI on this path (pybind11
 * `sam2._C.get_connected_componnets`, /root/reference/sam2/csrc/connected_components.cu:213-289);
 * `ds2_connected_components` replaces it one-for-one.  Every other entry point replaces a torch
 * library call made by the reference's Python modules; the file:line each one stands in for is
 * cited next to it.
 */
#ifndef DETSAM2_H
#define DETSAM2_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DS2_OK 0
#define DS2_E_ARG (-1)      /* invalid argument / unsupported shape */
#define DS2_E_ALIGN (-2)    /* pointer or leading dimension not aligned as required */
#define DS2_E_DRIVER (-3)   /* cuTensorMapEncodeTiled unavailable or failed */
#define DS2_E_NODEVICE (-4) /* no CUDA device */

/* ---- library / device ---------------------------------------------------------------------- */
int ds2_version(void);
/* number of kernels launched by this library since load (bench.py `gpu_launches`) */
int64_t ds2_launch_count(void);
const char* ds2_last_error(void);
int ds2_device_sm_count(void);

/* ---- GEMM: out = epilogue(A[M,K] @ W[N,K]^T) -------------------------------------------------
 * replaces nn.Linear / 1x1 nn.Conv2d / im2col'd convs everywhere on the path
 * (e.g. sam2/modeling/backbones/hieradet.py:49-50,75-80, memory_attention.py:47-48,
 *  sam/transformer.py:232-238, memory_encoder.py:98-101).
 * epilogue, in order:  v = acc + bias[n];  v = act(v);  v *= gamma[n];
 *                      rotary on column pairs (2i,2i+1) inside [rope_col0, rope_col1);
 *                      v += residual[row or row % res_row_mod][n];  store f32 and/or bf16.      */
typedef struct ds2_gemm_args {
  const void* A;   /* bf16 [M, K], row pitch lda elements (lda % 8 == 0, 16-byte aligned base) */
  const void* W;   /* bf16 [N, K], row pitch ldw elements */
  int64_t lda, ldw;
  int32_t M, N, K;
  const float* bias;     /* [N] or NULL */
  const float* gamma;    /* [N] or NULL */
  const float* residual; /* f32, row pitch ldr, or NULL */
  int64_t ldr;
  int32_t res_row_mod;   /* 0: residual row = row;  >0: row % res_row_mod (broadcast over batch) */
  int32_t act;           /* 0 none, 1 ReLU, 2 GELU(erf) */
  float* out_f32;        /* row pitch ldc, or NULL */
  void* out_bf16;        /* row pitch ldc_bf16, or NULL */
  int64_t ldc, ldc_bf16;
  /* rotary epilogue (RoPEAttention, sam/transformer.py:311-363; position_encoding.py:193-220) */
  const float* rope_cs;  /* [128][rope_period][2] (cos, sin), pair-major (position fastest), or NULL */
  int32_t rope_col0, rope_col1;   /* rotate columns c in [col0, col1); pair = ((c-col0) % 256)/2 */
  int32_t rope_period;            /* table rows; position = (row % rope_rows_per_batch) % period */
  int32_t rope_rows_per_batch;    /* rows per batch item */
  int32_t rope_row_limit;         /* rotate only rows with (row % rows_per_batch) < limit */
  int32_t impl;                   /* 0 = tcgen05 (product path: CTA-pair kernel for M >= 512, else single-CTA),
                                     1 = SIMT debug kernel, 2 = single-CTA with row-per-thread stores,
                                     3 = force the CTA-pair kernel, 4 = force the single-CTA kernel */
  /* axial form of the same table (takes precedence over rope_cs): [64][rope_side][2] (cos, sin) of ONE grid coordinate;
   * pair j < 64 rotates by x = position % side, pair j >= 64 by y = position / side with the frequencies of pair j - 64
   * (compute_axial_cis, position_encoding.py:173-182); requires rope_period == rope_side^2.  Staged in shared memory. */
  const float* rope_axial;
  int32_t rope_side;
} ds2_gemm_args;
int ds2_gemm(const ds2_gemm_args* args, void* stream);

/* ---- flash attention, one 256-wide head (memory attention) -----------------------------------
 * replaces F.scaled_dot_product_attention in RoPEAttention.forward
 * (sam2/modeling/sam/transformer.py:343-361) for MemoryAttentionLayer self- and cross-attention
 * (memory_attention.py:58-81).   out[b] = softmax(scale * q[b] k[b]^T) v[b]
 * q [B, Lq, 256], k [B, Lk, 256], v [B, Lk, DV], out [B, Lq, DV], DV in {64, 256}, all bf16.
 * ld*: row pitch in elements; bs*: batch pitch in elements.                                     */
typedef struct ds2_flash_args {
  const void* q; const void* k; const void* v; void* out;
  int64_t ldq, ldk, ldv, ldo;
  int64_t bsq, bsk, bsv, bso;
  int32_t B, Lq, Lk, DV;
  float scale;
  int32_t impl; /* 0 = tcgen05, 1 = SIMT debug kernel */
  /* Optional caller-owned scratch for the DV = 64 kernel (at least ds2_flash_workspace_bytes(B, Lq, DV) bytes, zeroed
   * once by the caller, private to one stream): every (object, query tile) item is then computed as two key halves
   * combined in a fixed order, which lets the launch run the items of its partial last wave as two CTAs each without
   * changing a bit of the result.  NULL: one pass per item.  The library keeps no state of its own between calls. */
  void* workspace;
  int64_t workspace_bytes;
  int32_t impl_flags; /* tests only: bit 0 = never split an item over two CTAs, bit 1 = always split */
} ds2_flash_args;
int ds2_flash_attn(const ds2_flash_args* args, void* stream);
int64_t ds2_flash_workspace_bytes(int32_t B, int32_t Lq, int32_t DV);

/* Tuning aid: barrier-stall cycle counters of flash launches made with impl == 8 (summed over CTAs):
 * [0] K-tile wait [1] V-tile wait [2] P wait [3] MMA-warp cycles [4] S wait [5] O wait [6] softmax-warp
 * cycles [7] CTAs.  Synchronises the device.  No reference counterpart. */
int ds2_debug_flash_stalls(unsigned long long* out16, int reset);   /* 16 counters */
/* Tuning aid: phase timestamps (cycles since kernel start) of CTA 0 of the last windowed-attention launch made
 * with DS2_WIN_DBG=1 in the environment.  Synchronises the device.  No reference counterpart. */
int ds2_debug_win_times(long long* out16);

/* ---- generic multi-head attention with optional window addressing (Hiera, mask decoder) ------
 * replaces F.scaled_dot_product_attention in MultiScaleAttention.forward
 * (backbones/hieradet.py:57-82, window_partition/unpartition backbones/utils.py:16-63) and in
 * sam/transformer.py:214-284 (Attention).  See ds2_mha_args in the kernel section below.        */
typedef struct ds2_mha_args {
  const void* q; const void* k; const void* v; void* out;  /* bf16 */
  int64_t q_tok_stride, k_tok_stride, v_tok_stride, o_tok_stride; /* elements between tokens */
  int64_t q_bs, k_bs, v_bs, o_bs;                                  /* elements between batches */
  int32_t B, H, D;       /* batch, heads, head_dim (<= 128); head h at column offset h*D */
  int32_t Lq, Lk;        /* tokens per batch item when window == 0 */
  /* window mode (window > 0): tokens live on a [Hm, Wm] raster per batch item; attention is
   * restricted to window x window tiles (zero padding participates, hieradet.py:145-148).  With
   * q_pool = 1 the query of each 2x2 cell is the element-wise max over the cell
   * (hieradet.py:65-68) and out has (Hm/2)*(Wm/2) tokens.                                       */
  int32_t window, Hm, Wm, q_pool;
  int32_t Lk_valid;      /* keys >= Lk_valid are masked (0 = all valid) */
  float scale;
  /* bf16 [H*D] rows substituted for zero-padded window tokens: a padded token is x = 0 after
   * norm1, so its q/k/v are the qkv bias (hieradet.py:145-148, backbones/utils.py:28-32)        */
  const void* pad_q; const void* pad_k; const void* pad_v;
} ds2_mha_args;
int ds2_mha(const ds2_mha_args* args, void* stream);

/* ---- normalisation / elementwise ------------------------------------------------------------- */
/* LayerNorm over the last dim (nn.LayerNorm; LayerNorm2d sam2_utils.py:150-162 is the same maths
 * on channels-last data).  y = LN(x)*w + b; optional GELU; optional second output y + pos.      */
typedef struct ds2_ln_args {
  const float* x; int64_t ldx; int32_t rows, C;
  const float* w; const float* b; float eps;
  int32_t act;               /* 0 none, 2 GELU */
  float* out_f32; void* out_bf16; int64_t ldo;
  const float* pos; int32_t pos_row_mod; /* second output = y + pos[row % pos_row_mod] */
  void* out2_bf16;
  const void* x_bf16;        /* alternative bf16 input (x == NULL) */
} ds2_ln_args;
int ds2_layernorm(const ds2_ln_args* args, void* stream);

/* y = a*alpha + b*beta (b row-broadcast with modulo), to f32 and/or bf16 */
int ds2_axpby(const float* a, const float* b, int64_t rows, int32_t C, int32_t b_row_mod,
              float alpha, float beta, float* out_f32, void* out_bf16, void* stream);
int ds2_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream);
int ds2_cast_bf16_f32(const void* x, float* y, int64_t n, void* stream);
/* 2x2 max-pool on a token-major [B, Hm, Wm, C] f32 map (hieradet.py:14-29 do_pool) */
int ds2_maxpool2x2(const float* x, float* y, int32_t B, int32_t Hm, int32_t Wm, int32_t C,
                   void* stream);
/* y[b, 2i+di, 2j+dj, :] = top[b, i, j, :] + lat[b, 2i+di, 2j+dj, :]  (FpnNeck top-down,
 * backbones/image_encoder.py:116-127, nearest) */
int ds2_upsample2x_add(const float* top, const float* lat, float* y, int32_t B, int32_t Hm,
                       int32_t Wm, int32_t C, void* stream);

/* ---- convolution helpers ---------------------------------------------------------------------- */
/* im2col for the patch embedding: fp16 NCHW frame [3, S, S] -> bf16 [ (S/4)^2, Kpad ] rows with
 * (c, ky, kx) ordering, k7 s4 p3 (backbones/utils.py:66-96). */
int ds2_im2col_patch(const void* frame_f16, void* out_bf16, int32_t S, int32_t Kpad, void* stream);
/* im2col k3 s2 p1 on channels-last bf16 [B, Hi, Wi, C] -> [B*Ho*Wo, 9*C] ((ky,kx,c) ordering) */
int ds2_im2col_k3s2(const void* x_bf16, void* out_bf16, int32_t B, int32_t Hi, int32_t Wi,
                    int32_t C, void* stream);
/* depth-wise 7x7 pad 3 on channels-last f32 [B, Hm, Wm, C] (memory_encoder.py:76-82) */
int ds2_dwconv7(const float* x, const float* w /*[C,49]*/, const float* bias, float* y, int32_t B,
                int32_t Hm, int32_t Wm, int32_t C, void* stream);

/* ---- memory-encoder mask path (memory_encoder.py:17-58, sam2_base.py:692-743) ------------------
 * stage 1: low-res logits [B, Sl, Sl] -> (bilinear x4, sigmoid or >0, *scale + bias) -> conv3x3 s2
 *          (1->4) -> LN2d -> GELU -> bf16 [B, 2Sl, 2Sl, 4].  Never materialises the 4Sl x 4Sl mask. */
int ds2_maskds_stage1(const float* lowres, int32_t B, int32_t Sl, int32_t binarize, float scale,
                      float bias, const float* w /*[4,1,3,3]*/, const float* b, const float* ln_w,
                      const float* ln_b, void* out_bf16, void* stream);
/* stage 2: direct conv3x3 s2 Cin->Cout + LN2d + GELU on channels-last bf16 */
int ds2_maskds_conv(const void* x_bf16, int32_t B, int32_t Hi, int32_t Wi, int32_t Cin,
                    int32_t Cout, const float* w /*[Cout,Cin,3,3]*/, const float* b,
                    const float* ln_w, const float* ln_b, void* out_bf16, void* stream);

/* ---- mask decoder pieces (sam/mask_decoder.py:163-247, sam2_base.py:342-397) ------------------ */
/* pixel-shuffle of a ConvTranspose2d(k2,s2) GEMM result + skip + LN2d + GELU:
 * g [B*Hm*Wm, 4*C] (col = (dy*2+dx)*C + c) -> y bf16 [B, 2Hm, 2Wm, C];  skip f32 [(2Hm*2Wm), C]  */
int ds2_upscale1(const float* g, const float* bias, const float* skip, const float* ln_w,
                 const float* ln_b, void* y_bf16, int32_t B, int32_t Hm, int32_t Wm, int32_t C,
                 void* stream);
/* second ConvT result g [B*Hm*Wm, 4*C] + skip -> GELU -> dot with hyper[B, M, C] ->
 * masks f32 [B, M, 2Hm, 2Wm] */
int ds2_upscale2_masks(const float* g, const float* bias, const float* skip, const float* hyper,
                       float* masks, int32_t B, int32_t Hm, int32_t Wm, int32_t C, int32_t M,
                       void* stream);
/* batched 3-layer MLP on small inputs: y[i] = W3 act(W2 act(W1 x[idx[i]] + b1) + b2) + b3
 * (hypernetwork / IoU / object-score / obj_ptr heads, sam2_utils.py:121-145).
 * nmlp independent weight sets laid out back to back; item i uses weight set i % nmlp.
 * Weights are INPUT-major: w1 [nmlp][din][dh], w2 [nmlp][dh][dh], w3 [nmlp][dh][dout] (the transpose of
 * nn.Linear.weight), dh, dout <= 256.                                                          */
typedef struct ds2_mlp3_args {
  const float* x; int64_t ldx; const int32_t* gather; /* optional row indices */
  int32_t rows, nmlp, din, dh, dout;
  const float* w1; const float* b1; const float* w2; const float* b2; const float* w3;
  const float* b3;
  int32_t sigmoid_out;
  float* y; int64_t ldy;
} ds2_mlp3_args;
int ds2_mlp3(const ds2_mlp3_args* args, void* stream);

/* SAM-head selection epilogue (sam2_base.py:342-397, mask_decoder.py:143-158,249-296):
 * from all_masks [B,4,S,S], ious [B,4], object score [B], tokens [B,4,C] choose the output mask
 * (multimask: argmax IoU over masks 1..3; single: mask 0 with the stability fallback), apply the
 * NO_OBJ_SCORE gate, and emit low_res [B,S,S], best index [B], chosen token [B,C].              */
int ds2_sam_select(const float* all_masks, const float* ious, const float* obj_score,
                   const float* mask_tokens, int32_t B, int32_t S, int32_t C, int32_t multimask,
                   float stab_delta, float stab_thresh, float* low_res, float* iou_out,
                   int32_t* best_idx, float* token_out, void* stream);
/* obj_ptr = lambda*ptr + (1-lambda)*no_obj_ptr, lambda = [score > 0] (sam2_base.py:376-387) */
int ds2_objptr_mix(float* ptr, const float* obj_score, const float* no_obj_ptr, int32_t B,
                   int32_t C, void* stream);

/* ---- memory bank (sam2_base.py:564-648) -------------------------------------------------------
 * gathers one memory frame into the step's key-input / value buffers:
 *   kin[b, row0 + t, :] = bf16(mem[b, t, :] + pos[t, :] + tpos[:]),  val[b, row0+t, :] = mem[b,t,:] */
int ds2_bank_gather(const void* mem_bf16, const float* pos, const float* tpos, void* kin_bf16,
                    void* val_bf16, int32_t B, int32_t T, int32_t C, int64_t dst_bs, int32_t row0,
                    void* stream);
/* object-pointer tokens: ptr f32 [B, 256] -> 4 tokens of 64; kin = bf16(ptr + tpos), val = bf16(ptr) */
int ds2_bank_ptr(const float* ptr, const float* tpos, void* kin_bf16, void* val_bf16, int32_t B,
                 int64_t dst_bs, int32_t row0, void* stream);

/* ---- prompt encoder / token assembly (sam/prompt_encoder.py:73-95,134-171; mask_decoder.py:163-186)
 * tokens f32 [B, n_out + P + 1, 256] = [out_tokens (obj-score, IoU, mask tokens) ; point embeddings ;
 * padding point].  coords f32 [B,P,2] in model pixels, labels int32 [B,P] (-1 pad, 0/1 clicks, 2/3 box). */
int ds2_prompt_tokens(const float* coords, const int32_t* labels, int32_t B, int32_t P,
                      const float* gauss /*[2,128]*/, const float* point_emb /*[4,256]*/,
                      const float* not_a_point /*[256]*/, const float* out_tokens /*[n_out,256]*/,
                      int32_t n_out, float image_size, float* tokens, void* stream);
/* object-pointer bank tokens with the temporal PE computed in-kernel (sam2_base.py:588-648,
 * sam2_utils.py:69-79): tpos = W[64,256] . sine1d(dist_norm, 256) + bias                        */
int ds2_bank_ptr_pe(const float* ptr, float dist_norm, const float* w, const float* bias, void* kin_bf16,
                    void* val_bf16, int32_t B, int64_t dst_bs, int32_t row0, void* stream);
/* whole-bank assembly from DEVICE-resident tables (graph-replay friendly: every launch argument is
 * static, the per-step variation lives in the tables).  frame_src[f] -> bf16 [B,T,C] memory of stored
 * frame f, frame_tpos[f] = row of tpos_table ([num_maskmem, C]); ptr_src[j] -> f32 [B,256] pointer,
 * ptr_dist[j] = signed temporal distance / (max_obj_ptrs - 1).  Output rows: nf*T memory tokens
 * then 4*np pointer tokens; kin/val bf16 [B, nf*T + 4*np, C].  Replaces the torch.cat / flatten /
 * permute / get_1d_sine_pe / obj_ptr_tpos_proj chain of sam2_base.py:564-650.                    */
int ds2_bank_assemble(const void* const* frame_src, const int32_t* frame_tpos, int32_t nf,
                      const float* const* ptr_src, const float* ptr_dist, int32_t np, const float* pos,
                      const float* tpos_table, const float* ptr_w, const float* ptr_bias, void* kin_bf16,
                      void* val_bf16, int32_t B, int32_t T, int32_t C, void* stream);
/* memory-encoder tail (sam2_base.py:733-741, sam2_video_predictor.py:1337):
 * out bf16 [B,T,C] = x + (1 - [score_b > 0]) * no_obj_embed                                      */
int ds2_memenc_finish(const float* x, const float* score, const float* no_obj_embed, void* out_bf16,
                      int32_t B, int32_t T, int32_t C, void* stream);

/* ---- post-processing --------------------------------------------------------------------------
 * drop-in for sam2._C.get_connected_componnets (csrc/connected_components.cu:213-282):
 * 8-connected components of a uint8 [N,1,H,W] mask; labels/counts int32 [N,1,H,W].              */
int ds2_connected_components(const uint8_t* mask, int32_t* labels, int32_t* counts, int32_t N,
                             int32_t H, int32_t W, void* stream);
/* fill_holes_in_mask_scores (sam2/utils/misc.py:365-393), fused: in-place on f32 [N,H,W] */
int ds2_fill_holes(float* scores, int32_t* labels_ws, int32_t* counts_ws, int32_t N, int32_t H,
                   int32_t W, int32_t max_area, void* stream);
/* F.interpolate(bilinear, align_corners=False) on f32 [N, Hi, Wi] -> [N, Ho, Wo]
 * (sam2_video_predictor.py:630-635, sam2_base.py:355-360) */
int ds2_resize_bilinear(const float* x, float* y, int32_t N, int32_t Hi, int32_t Wi, int32_t Ho,
                        int32_t Wo, void* stream);
/* (x > 0) bit-packed, 8 pixels per byte, little-endian bit order (det_sam2_RT.py:396-399) */
int ds2_threshold_pack(const float* x, uint8_t* bits, int64_t n, void* stream);
/* The result hand-off one step further: thresholds N masks [N, H, W] f32 at > 0 and emits
 * (a) bits: [N, H, ceil(W/8)] bytes, bit e of byte j of a row = pixel 8j+e (numpy.packbits bitorder "little"), and/or
 * (b) stats: uint64 [N][3] = {area, sum of x, sum of y} over the set pixels, i.e. the raw moments m00, m10, m01 that
 * Det-SAM2's post-processor obtains with cv2.moments on the host (postprocess_det_sam2.py:331-343); the centroid is
 * (sum_x / area, sum_y / area).  Integer arithmetic, bit-reproducible.  Either output may be NULL.             */
/* ---- dense mask prompts (add_new_mask, svp:527-600 -> SAM2Base._use_mask_as_output sam2_base.py:399-448) ------ */
/* y[B, S/4, S/4] = F.interpolate(x * scale + bias, scale 1/4, mode="bilinear", antialias=True, align_corners=False)
 * for x [B, S, S] f32 (the low-resolution logits that stand in for SAM's output, sam2_base.py:407-413)           */
int ds2_downsample4_aa(const float* x, float* y, int32_t B, int32_t S, float scale, float bias, void* stream);
/* mask_downsample (Conv2d 1->1 k4 s4, sam2_base.py:185-187,419) + PromptEncoder.mask_downscaling without its final
 * 1x1 conv (prompt_encoder.py:52-60: Conv 1->4 k2 s2, LayerNorm2d, GELU, Conv 4->16 k2 s2, LayerNorm2d, GELU):
 * mask [B, S, S] f32 -> bf16 [B * (S/16)^2, 16] token-major features; the 16 -> 256 conv is a ds2_gemm.
 * w0 [4][1][2][2], w3 [16][4][2][2] in Conv2d layout.
 * wds == bds == NULL: `mask` is already at the prompt encoder's input size ([B, S, S] with S = 4 x the embedding
 * side: the low-resolution logits of a previous decode fed back as a dense prompt by a refinement click,
 * sam2_base.py:306-329, svp:463-480); only mask_downscaling runs and the output has B * (S/4)^2 rows.            */
int ds2_mask_prompt_embed(const float* mask, int32_t B, int32_t S, const float* wds, const float* bds,
                          const float* w0, const float* b0, const float* ln0_w, const float* ln0_b,
                          const float* w3, const float* b3, const float* ln3_w, const float* ln3_b,
                          void* out_bf16, void* stream);

int ds2_mask_pack_stats(const float* x, uint8_t* bits, uint64_t* stats, int32_t N, int32_t H, int32_t W,
                        void* stream);

/* ---- frame ingest ------------------------------------------------------------------------
 * load_video_frames, ndarray branches (sam2/utils/misc.py:336-359): per frame `cv2.resize(frame_rgb, (S, S)) / 255.0`
 * stored into the fp16 `images` tensor, then `images -= mean; images /= std` in fp16.
 * src_u8: N RGB frames uint8 [Hv][Wv][3] (row pitch and frame stride in bytes);  dst_f16: fp16 [N][3][S][S].
 * The resize is OpenCV's 8-bit INTER_LINEAR reproduced bit for bit (11-bit fixed-point weights; the exact 2x
 * decimation is a 2x2 box mean; equal sizes copy), see oracle/resize_oracle.py.  lut_3x256: uint16 [3][256], the fp16
 * bit pattern of the normalised value of byte v in channel c, built by the caller with the reference's arithmetic
 * (the three roundings have only 3 x 256 distinct inputs).                                                        */
int ds2_ingest_frames(const uint8_t* src_u8, int32_t N, int32_t Hv, int32_t Wv, int64_t pitch_bytes,
                      int64_t frame_stride_bytes, const uint16_t* lut_3x256, void* dst_f16, int32_t S, void* stream);

/* ---- detector (YOLOv8) pre-processing on the device — SURVEY.md 8f rank 1 -----------------------------------------
 * replaces, for the frames Det-SAM2 hands to its detector (det_sam2_inference/det_sam2_RT.py:201-245), the HOST work of
 * ultralytics 8.2.82 (third party, requirements.txt:134): LetterBox (cv2.resize INTER_LINEAR to new_w x new_h, grey
 * border 114 to Wd x Hd), BGR->RGB, upload, cast, / 255.  src as for ds2_ingest_frames (uint8 RGB already in HBM);
 * dst planar RGB [N][3][Hd][Wd], fp32 (dst_is_f32) or fp16; lut_256 = the 256 values v / 255 in that dtype.  The
 * resized frame occupies rows [top, top + new_h), columns [left, left + new_w); byte-exact with cv2.              */
int ds2_letterbox_frames(const uint8_t* src_u8, int32_t N, int32_t Hv, int32_t Wv, int64_t pitch_bytes,
                         int64_t frame_stride_bytes, const void* lut_256, void* dst, int32_t dst_is_f32, int32_t Hd,
                         int32_t Wd, int32_t new_h, int32_t new_w, int32_t top, int32_t left, int32_t pad_value,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DETSAM2_H */
