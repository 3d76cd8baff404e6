/* ctx_b200.h — C ABI of the B200-native detection hot path (libctx_b200.so).
 *
 * Drop-in boundary for Ze-Yang/Context-Transformer's inference path.  Every entry point is
 * `extern "C"`, takes plain pointers / sizes / a `cudaStream_t` passed as `void*`, returns an
 * `int` status (0 = OK) and never prints or aborts; `ctx_last_error()` returns a thread-local
 * message for the last non-zero status.  Unless a name ends in `_host`, pointers are DEVICE
 * pointers owned by the caller (the Python host passes `tensor.data_ptr()`), the call is
 * asynchronous on `stream`, and the callee allocates nothing (scratch comes in as `workspace`).
 *
 * Each declaration cites the reference interface (file:line under the reference repo) it
 * replaces.  INTEGRATION.md shows the reference-side binding for each.
 */
#ifndef CTX_B200_H_
#define CTX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CTX_OK 0
#define CTX_ERR_INVALID 1     /* bad argument (null pointer, negative size, unsupported geometry) */
#define CTX_ERR_CUDA 2        /* a CUDA runtime / driver call failed */
#define CTX_ERR_WORKSPACE 3   /* workspace too small */
#define CTX_ERR_UNSUPPORTED 4 /* valid request that this build does not implement */

#define CTX_F32 0
#define CTX_BF16 1
#define CTX_F16 2

/* ---- library ------------------------------------------------------------------------------ */
int ctx_version(void);
const char* ctx_last_error(void);
/* Number of kernel launches issued by this library since process start (bench.py gpu_launches). */
unsigned long long ctx_launch_count(void);

/* ---- Detect.forward : layers/functions/detection.py:18-55 + utils/box_utils.py:184-202 ----- */
/* boxes_out[B,P,4] = decode(loc, priors, var);  scores_out[B,P,1+C] = cat(obj0, obj1*conf).   */
int ctx_detect_forward(const float* loc, const float* conf, const float* obj, const float* priors,
                       int batch, int num_priors, int num_fg_classes, float var0, float var1,
                       float* boxes_out, float* scores_out, void* stream);

/* ---- fused post-processing : test.py:133-161 (scale, per-class threshold + NMS, top-k) ------ */
typedef struct CtxPostParams {
  int batch, num_priors, num_fg_classes;
  float var0, var1;
  int scale_per_image;      /* 0: scale[4] shared, 1: scale[batch,4]                (test.py:123-124,136) */
  float score_thresh;       /* candidates are score >  score_thresh                 (test.py:143) */
  float nms_thresh;         /* IoU threshold, +1 pixel convention                   (test.py:152) */
  int suppress_on_equal;    /* 1: cpu_nms ">=" (cpu_nms.pyx:65), 0: gpu_nms ">" (nms_kernel.cu:71) */
  int nms_method;           /* 0 hard NMS; 1/2/3: cpu_soft_nms method 1 (linear) / 2 (gaussian) / 0 (hard) */
  float soft_sigma, soft_threshold; /* cpu_soft_nms sigma, threshold (Nt = nms_thresh)   (cpu_nms.pyx:70) */
  int max_per_image;        /* 200 upstream; <=0 disables the cut                   (test.py:96,155-161) */
  int max_out;              /* record capacity per image (>= max_per_image; ties may exceed it upstream) */
} CtxPostParams;
size_t ctx_postprocess_workspace_bytes(int batch, int num_priors, int num_fg_classes);
/* records[B,max_out,6] = x1,y1,x2,y2,score,class (class asc, score desc); counts[B] = number of
 * detections the reference would keep (may exceed max_out -> truncated); prior_idx[B,max_out]
 * (optional) = index of the prior each record came from. */
int ctx_detect_postprocess(const float* loc, const float* conf, const float* obj, const float* priors,
                           const float* scale, const CtxPostParams* p,
                           float* records, int* counts, int* prior_idx,
                           void* workspace, size_t workspace_bytes, void* stream);

/* ---- NMS : utils/nms_wrapper.py:23-31, utils/nms/gpu_nms.pyx:16-31, cpu_nms.pyx:17-68 -------- */
size_t ctx_nms_workspace_bytes(int n);
/* dets[n,5] device (x1,y1,x2,y2,score).  keep_out[n] device: indices into dets, best score first. */
int ctx_nms_device(const float* dets, int n, float thresh, int suppress_on_equal,
                   int* keep_out, int* num_out, void* workspace, size_t workspace_bytes, void* stream);
/* Host-pointer convenience with nms_wrapper.nms semantics (blocking). */
int ctx_nms_host(const float* dets_host, int n, float thresh, int suppress_on_equal,
                 int* keep_host, int* num_out_host, int device_id);
/* Legacy signature, literal drop-in for utils/nms/gpu_nms.hpp:1-2 (rows pre-sorted by the caller,
 * host pointers, blocking; errors are recorded in ctx_last_error() instead of printed). */
void _nms(int* keep_out, int* num_out, const float* boxes_host, int boxes_num, int boxes_dim,
          float nms_overlap_thresh, int device_id);
/* cpu_soft_nms : utils/nms/cpu_nms.pyx:70-163.  In-place on boxes_host[n,5]; *n_out = N_final. */
int ctx_soft_nms_host(float* boxes_host, int n, float sigma, float Nt, float threshold,
                      unsigned method, int* n_out, int device_id);

/* ---- conv / pool program : models/RFB_Net_vgg.py BasicConv :7-22, vgg :323-343, heads :387-416 */
typedef struct CtxOutSeg {   /* where output channels [c_begin, c_end) of a conv go */
  void* ptr;
  int c_begin, c_end;
  long long img_stride;      /* elements between images */
  int pix_stride;            /* elements between output pixels (row-major oy*Wo+ox) */
  int ch_offset;             /* element offset of channel c_begin inside a pixel */
  int dtype;                 /* CTX_F32 / CTX_BF16 / CTX_F16 */
} CtxOutSeg;

typedef struct CtxConvParams {
  int N, H, W, Cin;          /* NHWC input; Cin channels are consumed ...                       */
  int in_cstride, in_coffset;/* ... starting at channel in_coffset of a pixel of in_cstride channels */
  int Cout, KH, KW, stride, pad_h, pad_w, dil;
  int Ho, Wo;
  int relu;
  int pool2;                 /* 1: the epilogue also applies MaxPool2d(2, 2) (vgg 'M', RFB_Net_vgg.py:327-328): the output
                              * segment describes the POOLED [N, Ho/2, Wo/2, Cout] tensor (tensor-core path only)        */
  int relu_channels;         /* with relu != 0: ReLU only on output channels < relu_channels (0 = all) — fused entry convs */
  int in_dtype;
  int in_nchw;               /* 1: `in` is the raw fp32 NCHW network input [N,3,H,W] (RFBNet.forward x, :210) and this is the
                              * 3x3 / stride 1 / pad 1 stem conv: the tensor-core kernel builds the 27-value patches itself   */
  const void* in;
  const void* weight;        /* SIMT path: fp32 [KH*KW*Cin][Cout]; TC path: 16-bit [Cout_pad][KH*KW][Cin_pad] */
  const float* bias;         /* [Cout] (BatchNorm folded in) or NULL */
  const void* residual;      /* optional NHWC tensor added before ReLU (RFB shortcut, :59-61) */
  int res_dtype, res_cstride, res_coffset;
  int nseg;
  CtxOutSeg seg[3];
  /* split mode ("fp32x3": fp32 arithmetic emulated on the tensor cores; ctx_conv2d_x3_* only, 0 / NULL elsewhere).  Every 16-bit
   * tensor is a PAIR of fp16 planes of identical geometry, value = hi + lo (hi = fp16(v), lo = fp16(v - hi)): */
  int split;
  const void* in_lo;         /* lo plane of `in` (not used by the fp32 NCHW stem input) */
  const void* residual_lo;   /* lo plane of `residual` */
  void* out_lo;              /* lo plane of seg[0] when seg[0] is 16-bit */
  const float* out_scale;    /* [Cout]: the accumulator is multiplied by out_scale[c] before the bias — undoes the per-channel
                              * power-of-two scaling that keeps the fp16 weight planes in the normal range */
} CtxConvParams;

typedef struct CtxPoolParams { /* nn.MaxPool2d on NHWC views (vgg 'M'/'C'/pool5, conf pool :242-244) */
  int N, H, W, C, Ho, Wo, k, stride, pad;
  int dtype;
  const void* in; long long in_img_stride; int in_pix_stride;
  void* out; long long out_img_stride; int out_pix_stride;
  const void* in_lo; void* out_lo;   /* split mode (fp16 hi/lo planes, see CtxConvParams::split): the window maximum is taken on hi + lo */
} CtxPoolParams;

int ctx_conv2d_simt(const CtxConvParams* p, void* stream);          /* fp32-accumulate CUDA-core path */
int ctx_conv2d_tc_supported(const CtxConvParams* p);                /* 1 if the tcgen05 path takes it  */
/* tcgen05/TMA implicit-GEMM path.  The plan owns the TMA descriptors (pointers are baked in).   */
int ctx_conv2d_tc_plan_create(const CtxConvParams* p, void** plan_out);
/* Same, with the tiling chosen by the caller: n_tiles_n = number of output-channel tiles (0: ceil(Cout/256)), cluster = 1 | 2
 * CTAs per MMA (0: default), a_mode = -1 rule of thumb | 0 im2col gather | 1 TMA pixel patches | 2 halo patches (3x3: the 8x16
 * tile's neighbourhood staged once, nine shifted descriptors) | 3 halo patches with two CTAs per SM (tiles <= 128 wide) | 4 halo patches with the
 * layer's whole weight tensor resident in shared memory (single N tile, <= 2 channel blocks), commit_group = K-steps per
 * tcgen05.commit (0: by tile width | 1 | 2 | 4).  Outputs are bit-identical for every setting; ctx_prog_autotune() picks per
 * layer by measurement.  info8 = {tile width, N tiles, cluster, A mode (0 gather, 1 TMA, 2 stem, 3 halo, 4 halo x 2 CTAs/SM, 5 halo with resident weights), ring stages, grid,
 * commit group, TMA patch TW * 1000 + TH}. */
int ctx_conv2d_tc_plan_create_tuned(const CtxConvParams* p, int n_tiles_n, int cluster, int a_mode, int commit_group, void** plan_out);
int ctx_conv2d_tc_plan_info(void* plan, int* info8);
int ctx_conv2d_tc_plan_run(void* plan, void* stream);
void ctx_conv2d_tc_plan_destroy(void* plan);
/* conv1_1 -> ReLU -> conv1_2 -> ReLU [-> MaxPool2d(2,2)] of vgg() (models/RFB_Net_vgg.py:323-343, base.0 .. base.4) in ONE kernel:
 * the 64-channel full-resolution activation between the two convs never exists in HBM.  `conv12` describes conv1_2 exactly as for
 * ctx_conv2d_tc_plan_create (3x3 / stride 1 / pad 1, Cin = 64, Cout <= 64, one 16-bit output segment, ReLU, optional pool2); its `in`
 * is ignored.  stem_in = the raw fp32 NCHW network input [N,3,H,W] (W % 4 == 0), stem_weight = conv1_1 packed as for the in_nchw
 * stem conv (16-bit [64][64]: 27 values k = (ky*3+kx)*3+ci, then zeros), stem_bias = [64] fp32.  Results are bit-identical to the two
 * separate tensor-core convs.  The plan is run / destroyed with ctx_conv2d_tc_plan_run / _destroy (info8 A mode 6). */
int ctx_conv2d_stem2_supported(const CtxConvParams* conv12);
int ctx_conv2d_stem2_plan_create(const CtxConvParams* conv12, const float* stem_in, const void* stem_weight, const float* stem_bias, void** plan_out);
/* development aid: device buffer (>= 8 * 64 * 6 int64) receiving clock64() stamps of CTA 0 of the fused kernel's roles; NULL disables */
void ctx_debug_set_conv_timeline(void* device_buffer);
/* fp32 emulated on the tensor cores (precision 'fp32x3'): activations and weights are fp16 hi/lo plane pairs (p->split = 1),
 * weights [Cout_pad16][2 planes: hi, lo][KH*KW][Cin_pad64] scaled per output channel by a power of two (p->out_scale undoes it).
 * Per 64-channel K-step of one filter tap the kernel chains lo*Whi + hi*Wlo + hi*Whi (12 tcgen05.mma) into a fresh TMEM
 * accumulator and adds the result to an fp32 running sum in registers: the tensor core's fp32 adder truncates
 * (profiles/r2_acc_probe.txt), short chains + round-to-nearest adds keep the result within fp32 accuracy of the exact conv.
 * n_tiles_n = number of output-channel tiles (0: ceil(Cout/128)). */
int ctx_conv2d_x3_supported(const CtxConvParams* p);
int ctx_conv2d_x3_plan_create(const CtxConvParams* p, int n_tiles_n, void** plan_out);
int ctx_conv2d_x3_plan_run(void* plan, void* stream);
int ctx_conv2d_x3_plan_info(void* plan, int* info8);
void ctx_conv2d_x3_plan_destroy(void* plan);
int ctx_maxpool2d_nhwc(const CtxPoolParams* p, void* stream);
/* x[N,3,H,W] fp32 NCHW (RFBNet.forward input, :210) -> NHWC of dtype */
int ctx_nchw_to_nhwc(const float* in, void* out, int N, int C, int H, int W, int out_dtype, void* stream);
/* BaseTransform (data/data_augment.py:224-266) for images that already have the network's size (cv2.resize to the same size
 * is a copy): img[N,H,W,3] uint8 (cv2 channel order) -> x[N,3,H,W] fp32 = float(img) - means[c].  means3 is a HOST pointer to
 * three floats ((104,117,123), test.py:87).  Other sizes: ctx_base_transform_resize. */
int ctx_base_transform(const unsigned char* img_hwc, float* out_chw, int N, int H, int W, const float* means3, void* stream);
/* BaseTransform with the resize (data/data_augment.py:257-261): img[src_h,src_w,3] uint8 -> cv2.resize(.., (size, size),
 * INTER_LINEAR) -> float - means -> out[3,size,size].  OpenCV's 8-bit bilinear kernel (fixed point, 11-bit coefficients,
 * imgproc/resize.cpp) restated operation for operation; one image per call (test.py:122-126 transforms one image at a time). */
int ctx_base_transform_resize(const unsigned char* img_hwc, int src_h, int src_w, float* out_chw, int size, const float* means3, void* stream);
/* x[N,3,H,W] fp32 -> 3x3/pad-1 patches [N,H,W,64] 16-bit (27 values, channel = (ky*3+kx)*3+ci, + 37 zeros): turns the
 * Cin = 3 stem conv (vgg() base.0, RFB_Net_vgg.py:331) into a K = 64 tensor-core GEMM */
int ctx_nchw_to_patch27(const float* in, void* out, int N, int H, int W, int out_dtype, void* stream);

/* ---- Context-Transformer : models/RFB_Net_vgg.py:253-271 (+ :273-285 output activation) ------ */
typedef struct CtxAttnParams {
  int batch, num_priors, num_pooled, dim;      /* P, Pk, C_src (60 transfer / 15 incre) */
  int num_novel;                               /* rows of OBJ_Target (20 / 5) */
  int incre;                                   /* 1: conf = cat(fc_base(conf)+conf, novel) */
  int apply_softmax;                           /* eval mode (:279-285) */
  const float* conf;                           /* [B,P,dim] raw conf head output */
  const float* pooled;                         /* [B,Pk,dim] spatially max-pooled conf */
  const float *theta_w, *theta_b, *phi_w, *phi_b, *g_w, *g_b;   /* [dim,dim], [dim] */
  const float *fc_base_w, *fc_base_b;          /* incre only */
  const float* Wz;                             /* [dim] */
  const float* obj_target_w;                   /* [num_novel, dim] */
  float scale;
  int use_tensor_cores;                        /* 0: fp32 CUDA-core kernel; 1: tcgen05 kernel, fp16 logits (|dconf| ~6e-3 max /
                                                * 2e-5 mean: the 16-bit engine modes); 2: tcgen05, fp16 hi/lo split logits (4e-5);
                                                * 3: tcgen05, hi/lo split logits AND hi/lo split P and V (fp32-grade output: the
                                                * 'fp32x3' engine mode, 1e-4 parity bar) */
  void* workspace;                             /* ctx_attention_workspace_bytes(p) bytes, 1024-byte aligned */
  size_t workspace_bytes;
  float* out;                                  /* [B,P, incre ? dim+num_novel : num_novel] */
} CtxAttnParams;
size_t ctx_attention_workspace_bytes(const CtxAttnParams* p);   /* reads batch, num_priors, num_pooled, dim, use_tensor_cores */
int ctx_attention_forward(const CtxAttnParams* p, void* stream);
/* development aid: device buffer (>= 512 int64) receiving clock64() stamps of CTA (0,0) of the tensor-core kernel; NULL disables */
void ctx_debug_set_attention_timeline(void* device_buffer);
/* row softmax over the last dim (C <= 128) — output activation :279-285 for obj / non-'ours' conf */
int ctx_softmax_lastdim(const float* in, float* out, long long rows, int cols, void* stream);

/* ---- op program: a recorded list of the ops above, replayed with one call ------------------- */
int ctx_prog_create(void** prog_out);
int ctx_prog_add_conv_simt(void* prog, const CtxConvParams* p);
int ctx_prog_add_conv_tc(void* prog, const CtxConvParams* p);
int ctx_prog_add_conv_x3(void* prog, const CtxConvParams* p);
int ctx_prog_add_conv_stem2(void* prog, const CtxConvParams* conv12, const float* stem_in, const void* stem_weight, const float* stem_bias);
int ctx_prog_add_pool(void* prog, const CtxPoolParams* p);
int ctx_prog_add_nchw_to_nhwc(void* prog, const float* in, void* out, int N, int C, int H, int W, int out_dtype);
int ctx_prog_add_nchw_to_patch27(void* prog, const float* in, void* out, int N, int H, int W, int out_dtype);
int ctx_prog_add_attention(void* prog, const CtxAttnParams* p);
int ctx_prog_add_softmax(void* prog, const float* in, float* out, long long rows, int cols);
/* Ops added after this call belong to `lane` (0..7); the next op added first waits for everything issued so far on the
 * lanes in `wait_mask` (bit l = lane l).  Lanes are the independent chains of the forward (the branches of an RFB block,
 * RFB_Net_vgg.py:48-54,93-101; the per-level heads, :238-248): the captured graph runs them side by side, serial replay
 * (ctx_prog_run_range) ignores them — program order must therefore already be a valid execution order. */
int ctx_prog_set_lane(void* prog, int lane, unsigned wait_mask);
/* time every tensor-core conv of the program under its candidate tilings on `stream` and keep the fastest (blocking) */
int ctx_prog_autotune(void* prog, void* stream, int reps);
int ctx_prog_conv_config(void* prog, int op_index, int* info8);   /* zeros for ops that are not tensor-core convs */
int ctx_prog_num_ops(void* prog);
int ctx_prog_run(void* prog, void* stream);
/* capture the op list into a CUDA graph on `stream` (non-default); later ctx_prog_run calls replay it */
int ctx_prog_instantiate_graph(void* prog, void* stream);
/* run ops [first, last) — used by bench.py/ncu to time one layer class */
int ctx_prog_run_range(void* prog, int first, int last, void* stream);
/* the same range `reps` times back to back (per-kernel timing of small ops without a host-language call per launch) */
int ctx_prog_run_range_repeat(void* prog, int first, int last, int reps, void* stream);
void ctx_prog_destroy(void* prog);

/* ---- training targets : utils/box_utils.py:83-156, multibox_loss_combined.py:88-96 ----------- */
/* Batched match(): truths[B,max_obj,6] (x1,y1,x2,y2,label,weight; rows >= num_obj[b] ignored).
 * Outputs loc_t[B,P,4], conf_t[B,P,2], obj_t[B,P] (uint8); optional best_truth_idx[B,P] and
 * best_truth_overlap[B,P] (the `overlap` argument of match(), box_utils.py:115-116).          */
int ctx_match_encode(const float* truths, const int* num_obj, int max_obj, const float* priors,
                     int batch, int num_priors, float threshold, float var0, float var1,
                     float* loc_t, float* conf_t, unsigned char* obj_t, int* best_truth_idx,
                     float* best_truth_overlap, void* stream);
/* rank[b,p] = position of p in the descending stable sort of loss[b,:] (the reference's two sorts). */
size_t ctx_rank_workspace_bytes(int batch, int num_priors);
int ctx_hard_negative_rank(const float* loss, int batch, int num_priors, int* rank,
                           void* workspace, size_t workspace_bytes, void* stream);

/* MultiBoxLoss_combined.forward + its backward, layers/modules/multibox_loss_combined.py:76-122 (after match(): loc_t, conf_t,
 * obj_t from ctx_match_encode).
 * ctx_loss_mining: mining[B,P] = CE(obj_p, 0) with positives / ignored boxes zeroed (:88-90) — the input of
 *   ctx_hard_negative_rank — and num_pos_w[B] (fp64) = sum of the mixup weights of the positives (:77; the caller truncates to
 *   integer and forms num_neg = min(3 num_pos, P - 1), :94).
 * ctx_loss_forward_backward: sums3 (fp64) = {loss_box_reg, loss_cls, loss_obj} BEFORE the division by N (:119-122), and the
 *   gradients of those sums: dloc[B,P,4] (of sums3[0]), dconf[B,P,C] and dobj_cls[B,P,2] (of sums3[1]), dobj_obj[B,P,2] (of
 *   sums3[2]); rows outside pos | neg are zero.  rank[B,P] from ctx_hard_negative_rank, num_neg[B] int64. */
int ctx_loss_mining(const float* obj_p, const float* conf_t, const unsigned char* obj_t, int batch, int num_priors,
                    float* mining, double* num_pos_w, void* stream);
int ctx_loss_forward_backward(const float* loc_p, const float* conf_p, const float* obj_p, const float* loc_t, const float* conf_t,
                              const unsigned char* obj_t, const int* rank, const long long* num_neg, int batch, int num_priors,
                              int num_fg_classes, double* sums3, float* dloc, float* dconf, float* dobj_cls, float* dobj_obj,
                              void* stream);

/* Box algebra helpers of utils/box_utils.py as stand-alone device ops (the match kernel uses the same device functions):
 * point_form :5-14 (cx,cy,w,h -> corners), jaccard :50-68 (out[na,nb], corner-form boxes, no +1), encode :135-156. */
int ctx_point_form(const float* boxes, int n, float* out, void* stream);
int ctx_jaccard(const float* box_a, int na, const float* box_b, int nb, float* out, void* stream);
int ctx_encode(const float* matched, const float* priors, int n, float var0, float var1, float* out, void* stream);

/* OBJ(Target) prototype initialisation, train.py:252-286 (init_reweight).  feat[B,P,dim] = model(x, init=True) (raw conf
 * features), conf_t[B,P,2] = the match() labels (ctx_match_encode).  Adds every positive prior's L2-normalised feature row
 * to sums[num_fg, dim] (fp64, caller-zeroed before the first batch) and its class to counts[num_fg]; call once per batch.
 * ctx_prototype_finalize writes out[num_fg - first_class, dim] = normalise(mean of the class's rows) for classes
 * first_class.. (15 for the 'incre' setting, train.py:281-282); a class with no sample yields NaN, as upstream. */
int ctx_prototype_accumulate(const float* feat, const float* conf_t, int batch, int num_priors, int dim, int num_fg,
                             double* sums, int* counts, void* stream);
int ctx_prototype_finalize(const double* sums, const int* counts, int num_fg, int dim, int first_class, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CTX_B200_H_ */
