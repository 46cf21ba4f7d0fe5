/*
 * sfron_b200.h — C ABI of the B200-native SFR-on hot path (libsfron_b200.so).
 *
 * The reference (K1nght/Unified-Unlearning-w-Remain-Geometry) has no FFI: the
 * hot path is inline Python between loss.backward() and the next forward.  Each
 * entry point below replaces one stock-torch op sequence of that path; the
 * reference file:line it replaces is cited per function (paths relative to the
 * reference root).
 *
 * Conventions
 *   - plain C, no torch types: raw DEVICE pointers, element counts, a CUDA stream
 *     handle (cudaStream_t passed as void*; NULL = legacy default stream).
 *   - every call is asynchronous on `stream`, re-entrant across streams, and keeps
 *     no pointer past return.  The caller (torch) owns every buffer.
 *   - return value: 0 = SFR_OK, negative = SFR_ERR_* (argument error),
 *     positive = a cudaError_t raised by the launch.  No exceptions cross the ABI.
 *   - all vector base pointers must be 16-byte aligned (SFR_ERR_ALIGN otherwise);
 *     `n` is arbitrary (ragged tails are handled inside the kernels); n == 0 is a
 *     no-op that returns SFR_OK.
 *   - floating point is IEEE fp32, round-to-nearest, no fast-math, no implicit FMA
 *     contraction; FMA is used exactly where torch's CPU kernels use it
 *     (see DESIGN.md "Arithmetic contract").
 */
#ifndef SFRON_B200_H
#define SFRON_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFR_ABI_VERSION 3

#if defined(__GNUC__)
#define SFR_API __attribute__((visibility("default")))
#else
#define SFR_API
#endif

/* ---- return codes --------------------------------------------------------- */
#define SFR_OK 0
#define SFR_ERR_NULL (-1)     /* a required pointer is NULL                      */
#define SFR_ERR_ALIGN (-2)    /* a vector pointer is not 16-byte aligned         */
#define SFR_ERR_ARG (-3)      /* an enum / size / flag argument is out of range  */
#define SFR_ERR_NO_DEVICE (-4)/* no usable sm_100 CUDA device / context          */

/* ---- element types of gradient streams ------------------------------------ */
#define SFR_F32 0
#define SFR_BF16 1

/* Masks are always 1 byte per element with values 0/1 (torch.bool / torch.uint8 storage).
 * The int64 0/1 form of the SalUn top-k masks (runners/diffusion.py:1026-1030) is produced
 * by the format exporter at save time, never on the hot path. */

typedef void* sfr_stream_t; /* cudaStream_t */

/* Library / device introspection (no GPU work). */
SFR_API int sfr_abi_version(void);
SFR_API const char* sfr_error_string(int code);
/* Number of SMs of the current device and max resident CTAs the library sizes its
 * persistent grids for; returns SFR_ERR_NO_DEVICE without a GPU. */
SFR_API int sfr_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ===========================================================================
 * K1  Fisher-diagonal accumulation
 *   acc[i] <- acc[i] + (g_b[i] * g_b[i]) / divisor      for b = 0 .. rows-1, in order
 * replaces  `F[name] += param.grad.data.cpu()**2 / len(loader)`
 *   Classification/unlearn/sfron.py:288-291,315-318
 *   DDPM/runners/diffusion.py:1277-1281,1342-1346   (with clip: see `clip_*`)
 *   DiT/generate_fisher.py:236-239,276-279
 *   SD/train-scripts/generate_fisher.py:73-76,123-126
 * and the per-sample FIM  `F += tmp_i**2 / |D|`  DDPM/runners/diffusion.py:337-344
 * (rows = samples, row_stride = elements between consecutive rows of g).
 *
 * If `clip_sumsq` != NULL every g is first multiplied by
 *   coef = min(1, clip_max_norm / (sqrt(*clip_sumsq) + 1e-6))
 * (torch.nn.utils.clip_grad_norm_, DDPM/runners/diffusion.py:1270-1275: the DDPM
 * Fisher is of the CLIPPED batch gradient).  *clip_sumsq is a device double that
 * sfr_masked_sumsq produced earlier on the same stream.
 * Arithmetic is bit-exact with the reference's CPU fp32 sequence.
 * ======================================================================== */
SFR_API int sfr_fisher_accum(float* acc, const void* g, int g_dtype, int64_t rows,
                     int64_t row_stride, int64_t n, float divisor,
                     const double* clip_sumsq, float clip_max_norm,
                     sfr_stream_t stream);

/* SalUn saliency accumulation (the input of the K2b top-k mask):  acc[i] <- acc[i] + g[i] * coef
 * replaces  `clip_grad_norm_(...)` + `gradients[name] += param.grad.data.cpu()`
 *   DDPM/runners/diffusion.py:985-994 (clipped), Classification/unlearn/salun.py:163-169 (unclipped)
 * coef as in sfr_fisher_accum when clip_sumsq != NULL, else 1 (no multiply).  Bit-exact (mul, then add). */
SFR_API int sfr_grad_accum(float* acc, const void* g, int g_dtype, int64_t n,
                   const double* clip_sumsq, float clip_max_norm, sfr_stream_t stream);

/* ===========================================================================
 * K2a  ratio saliency mask
 *   mask[i] = ((ff[i] + eps) / (rf[i] + eps)) >= threshold ;  *zero_count += #(mask == 0)
 * replaces  sfron.py:325-334, DDPM/generate_fisher_mask.py:39-45,
 *           DiT/generate_mask.py:31-39, SD/train-scripts/generate_fisher_mask.py:39-45
 * (eps = 1e-15 there).  `zero_count` (device u64, may be NULL) is ADDED to, so the
 * caller zeroes it once per mask ("Total sparsity" print of the reference).
 * Bit-exact (IEEE fp32 add / divide / compare).
 * ======================================================================== */
SFR_API int sfr_ratio_mask(const float* ff, const float* rf, int64_t n, float threshold,
                   float eps, uint8_t* mask, unsigned long long* zero_count,
                   sfr_stream_t stream);

/* Same, for up to SFR_MAX_THRESHOLDS thresholds in ONE pass over ff/rf
 * (DiT/generate_mask.py:25-46 re-reads both Fishers per threshold).
 * mask t is written at  masks + t * mask_stride  (bytes; multiple of 16),
 * zero_counts[t] (device, may be NULL) is added to. */
#define SFR_MAX_THRESHOLDS 8
SFR_API int sfr_ratio_mask_multi(const float* ff, const float* rf, int64_t n,
                         const float* thresholds_host, int n_thresholds, float eps,
                         uint8_t* masks, int64_t mask_stride,
                         unsigned long long* zero_counts, sfr_stream_t stream);

/* ===========================================================================
 * K2b  exact global top-k selection (two-pass radix / histogram select)
 * replaces  `ranks = argsort(argsort(-cat(|g|))); mask = ranks < int(N*ratio)`
 *   DDPM/runners/diffusion.py:1009-1034, Classification/unlearn/salun.py:170-193
 *   (and the k-th order statistic of SD/train-scripts/proximal_gradient.py:161-165).
 *
 * Keys.  key(x) = 0 for NaN, else bits(|x|) + 1  (monotone in |x|; NaN ranks last and
 * is never selected, as in torch's ascending sort of -|x|).
 *   SFR_KEY_ABS   : x = a[i]                       (b ignored)
 *   SFR_KEY_RATIO : x = (a[i] + eps) / (b[i] + eps)  (top-k by Fisher ratio; extension)
 *   SFR_KEY_ABSDIFF: x = a[i] - b[i]                 (SD/train-scripts/proximal_gradient.py:158-159)
 *
 * Protocol (all on one stream, no host synchronisation needed):
 *   memset(state, 0), state->k = k                      (sfr_select_init)
 *   sfr_select_hist(pass = 0)   -> bins[0 .. SFR_SELECT_BINS0)    (+= local counts)
 *   [all-reduce(sum) bins over ranks]
 *   sfr_select_scan(pass = 0)   -> state: 15-bit prefix, k remaining inside that bin
 *   sfr_select_hist(pass = 1)   -> bins[0 .. SFR_SELECT_BINS1)    (only keys matching prefix)
 *     or sfr_select_hist1_mask  -> the same, plus the provisional mask (see below)
 *   [all-reduce(sum) bins over ranks]
 *   sfr_select_scan(pass = 1)   -> state: thr_key, count_gt, count_eq, tie_budget
 *   sfr_select_apply            -> mask
 * Selection contract: every key > thr_key is selected; of the count_eq keys == thr_key
 * exactly tie_budget = k - count_gt are selected, lowest flat index first (= a STABLE
 * argsort; the reference's default argsort leaves the choice among ties unspecified).
 * `tie_base` = number of keys == thr_key that precede this shard (0 on one GPU).
 * Masks are bit-exact with the stable-sort reference.
 * ======================================================================== */
#define SFR_KEY_ABS 0
#define SFR_KEY_RATIO 1
#define SFR_KEY_ABSDIFF 2 /* x = a[i] - b[i]  (proximal gradient: |theta - theta0|) */
#define SFR_SELECT_BINS0 32768 /* key bits [30:16] */
#define SFR_SELECT_BINS1 65536 /* key bits [15:0]  */
/* u64 words every bins_dev buffer must hold: the histogram bins plus the scan's per-CTA partial sums and
 * ticket counter.  Only bins[0 .. SFR_SELECT_BINS0 / BINS1) take part in the multi-GPU all-reduce. */
#define SFR_SELECT_BINS_ALLOC (65536 + 128)

typedef struct sfr_select_state {
  unsigned long long k;          /* in: number of elements to select (global)        */
  unsigned long long k_in_bin;   /* after scan 0: rank wanted inside the chosen bin  */
  unsigned long long count_gt;   /* after scan 1: #keys >  thr_key (global)          */
  unsigned long long count_eq;   /* after scan 1: #keys == thr_key (global)          */
  unsigned long long tie_budget; /* after scan 1: #keys == thr_key to select         */
  uint32_t prefix;               /* after scan 0: chosen value of key bits [30:16]   */
  uint32_t thr_key;              /* after scan 1: the k-th largest key               */
  uint32_t select_all;           /* k >= n_valid: everything (non-NaN included) goes */
  uint32_t select_none;          /* k == 0                                           */
  unsigned long long reserved[3];
} sfr_select_state;

SFR_API int sfr_select_init(sfr_select_state* state_dev, unsigned long long* bins_dev,
                    unsigned long long k, sfr_stream_t stream);
/* scratch_dev (>= sfr_select_scratch_elems(n) u64) is required for pass 1, which stages the keys
 * matching the chosen prefix there for the tie handling of sfr_select_apply; unused by pass 0. */
SFR_API int sfr_select_hist(const float* a, const float* b, int key_mode, float eps,
                    int64_t n, int pass, const sfr_select_state* state_dev,
                    unsigned long long* bins_dev, unsigned long long* scratch_dev,
                    sfr_stream_t stream);
/* Pass 1 that ALSO writes a provisional mask (1 where key[30:16] > the chosen prefix, else 0): final for
 * every element except the staged candidates.  A following sfr_select_apply on the SAME mask and scratch then
 * finishes the mask from the candidate list alone — the select reads the vector twice instead of three times
 * (4 + 4 + 1 B/elem).  If a candidate region overflows, or apply is given another mask buffer, apply falls
 * back to its full streaming pass; results are identical either way. */
SFR_API int sfr_select_hist1_mask(const float* a, const float* b, int key_mode, float eps,
                    int64_t n, const sfr_select_state* state_dev,
                    unsigned long long* bins_dev, unsigned long long* scratch_dev,
                    uint8_t* mask, sfr_stream_t stream);
SFR_API int sfr_select_scan(int pass, sfr_select_state* state_dev,
                    unsigned long long* bins_dev, sfr_stream_t stream);
/* tie_base_dev: device u64 (NULL = 0).  scratch_dev: the SAME device u64 array of at least
 * sfr_select_scratch_elems(n) elements that pass 1 filled (per-chunk tie counts, per-block tie
 * bases, and the staged candidates; read only when ties must be ordered).  sfr_select_apply may be
 * repeated after one pass 1 (it re-derives the tie counts each time). */
SFR_API int64_t sfr_select_scratch_elems(int64_t n);
SFR_API int sfr_select_apply(const float* a, const float* b, int key_mode, float eps,
                     int64_t n, const sfr_select_state* state_dev,
                     const unsigned long long* tie_base_dev,
                     unsigned long long* scratch_dev, uint8_t* mask,
                     sfr_stream_t stream);

/* ===========================================================================
 * Clip norm:  *out += sum_i (g[i] * mask[i])^2          (device double)
 * replaces the norm half of torch.nn.utils.clip_grad_norm_ after `grad *= mask`
 *   sfron.py:201-205, DDPM/runners/diffusion.py:1126-1136, DiT/forget.py:289-298.
 * mask may be NULL (remain step / Fisher clip).  Accumulates in double, so the norm
 * is at least as accurate as torch's fp32 norm-of-norms (|rel diff| <= ~1e-7).
 * ======================================================================== */
SFR_API int sfr_masked_sumsq(const void* g, int g_dtype, const uint8_t* mask, int64_t n,
                     double* out, sfr_stream_t stream);

/* ===========================================================================
 * K3  fused saliency-masked fast/slow update — ONE pass over the shard:
 *   g' = g * mask                (param.grad *= mask[name])           [SFR_F_MASK]
 *   g' = g' * coef               (clip_grad_norm_)                    [clip_sumsq != NULL]
 *   optimizer step on (p, m, v)  torch.optim.{SGD(momentum), Adam, AdamW}
 *   slow / EMA weights           EMAHelper.update | update_ema | update_parameters
 *   g  = 0                       optimizer.zero_grad() for flat-view grads [SFR_F_ZERO_GRAD]
 * replaces
 *   mask·grad   sfron.py:201-204, runners/diffusion.py:1126-1129, DiT/forget.py:289-292,
 *               SD nsfw_removal.py:157-160 (intended behaviour), gradient_ascent.py:94-99
 *   clip        sfron.py:205, runners/diffusion.py:1131-1136,1169-1174, DiT/forget.py:293-298
 *   step        sfron.py:206,222 (SGD m=.9 wd=5e-4), runners/diffusion.py:1138,1176 (Adam),
 *               DiT/forget.py:299,320 (AdamW wd=0), SD nsfw_removal.py:162,173 (Adam)
 *   slow/EMA    DDPM/models/ema.py:17-24, DiT/forget.py:52-62,322, sfron.py:30-37,255-257
 * Arithmetic follows torch 2.11 _single_tensor_{adam,sgd} op by op;
 * results agree with the CPU reference to <= 1e-6 relative.
 * ======================================================================== */
#define SFR_OPT_SGD 0
#define SFR_OPT_ADAM 1  /* L2 weight decay folded into the gradient          */
#define SFR_OPT_ADAMW 2 /* decoupled weight decay                            */

#define SFR_EMA_NONE 0
#define SFR_EMA_DDPM 1     /* s = (1-mu)*p + mu*s          DDPM/models/ema.py:22-24  (ema_a = mu)   */
#define SFR_EMA_DIT 2      /* s = s*d + (1-d)*p            DiT/forget.py:62          (ema_a = d)    */
#define SFR_EMA_SLOWFAST 3 /* p = (1-b)*s + b*p ; s = p    sfron.py:30-37,126-127,255-257 (ema_a = b) */

#define SFR_F_MASK 1u            /* multiply g by mask before clipping (SFR-on order)  */
#define SFR_F_MASK_AFTER_CLIP 2u /* SalUn-DDPM order: clip first (norm of UNMASKED g), then mask
                                    runners/diffusion.py:579-590                        */
#define SFR_F_ZERO_GRAD 4u       /* write zeros back to g                               */
#define SFR_F_SGD_FIRST_STEP 8u  /* momentum buffer does not exist yet: buf = g         */
#define SFR_F_WRITE_BF16 16u     /* also write bf16(p) to p_bf16 (working copy)         */
#define SFR_F_REUSE_CONSTS 32u   /* consts_scratch_dev already holds THIS optimizer step's scalars: an earlier
                                    launch of the same step on another range of the vector prepared them (a shard
                                    updated in several pieces, e.g. pipelined against the backward pass).  The step
                                    counter is not advanced and the scalars are not recomputed. */

typedef struct sfr_update_args {
  int32_t opt;       /* SFR_OPT_*  */
  int32_t ema_mode;  /* SFR_EMA_*  */
  uint32_t flags;    /* SFR_F_*    */
  int32_t g_dtype;   /* SFR_F32 | SFR_BF16 */
  int64_t step;      /* Adam/AdamW: 1-based step count AFTER the increment (state['step']) */
  double lr;
  double beta1;        /* Adam beta1                                      */
  double beta2;        /* Adam beta2                                      */
  double eps;          /* Adam eps                                        */
  double weight_decay;
  double momentum;     /* SGD                                             */
  double dampening;    /* SGD                                             */
  double ema_a;        /* mu | decay | beta, see SFR_EMA_*                */
  double clip_max_norm;/* used when clip_sumsq != NULL                    */
  /* Learning-rate schedule on the device (optional; both NULL = use `lr`): the step's learning rate is
   * lr_table_dev[*lr_index_dev], read by the scalar-prep kernel, so that a captured launch follows the schedule at
   * every replay (the classification loop steps a CosineAnnealingLR each iteration: sfron.py:172-174,259).  The
   * caller advances *lr_index_dev between iterations (any stream-ordered increment).  Requires consts_scratch_dev
   * in sfr_fused_update. */
  const double* lr_table_dev;
  const long long* lr_index_dev;
} sfr_update_args;

/* consts_scratch_dev (optional, >= 128 bytes of 16-byte-aligned device memory): a one-thread prep kernel
 * leaves the step-dependent scalars and the clip coefficient there, so the main kernel reads plain
 * floats instead of deriving the coefficient from *clip_sumsq in front of every CTA (~3.5 % faster when
 * clipping).  step_counter_dev (optional, requires the scratch): a device int64 that the prep kernel
 * increments and uses INSTEAD of args->step — the optimizer step then lives on the device and advances
 * at every replay of a captured launch (CUDA graphs).  Both NULL: everything by value, args->step. */
#define SFR_UPDATE_CONSTS_BYTES 128
SFR_API int sfr_fused_update(float* p, void* g, float* m, float* v, const uint8_t* mask,
                     float* ema, void* p_bf16, int64_t n,
                     const sfr_update_args* args, const double* clip_sumsq,
                     long long* step_counter_dev, void* consts_scratch_dev,
                     sfr_stream_t stream);

/* Clip norm + K3 in ONE cooperative launch, for small vectors (ResNet-18, an 8-way shard): zero *sumsq, accumulate
 * sum((g*mask)^2) [mask only with SFR_F_MASK: the SalUn order clips the unmasked gradient], grid barrier, then the
 * update of sfr_fused_update with coef = min(1, args->clip_max_norm / (sqrt(*sumsq) + 1e-6)).  Replaces four
 * stream-ordered launches (memset, sfr_masked_sumsq, the scalar prep kernel, the update); the second read of g is
 * served by the L2 when the vector fits.  *sumsq (device double) is left holding the norm's square, as after
 * sfr_masked_sumsq.  step_counter_dev as in sfr_fused_update.  Same arithmetic, same results.  Single-GPU only: a
 * sharded vector needs its norm summed across ranks between the two phases. */
SFR_API int sfr_clipped_update(float* p, void* g, float* m, float* v, const uint8_t* mask,
                     float* ema, void* p_bf16, int64_t n, const sfr_update_args* args,
                     double* sumsq, long long* step_counter_dev, sfr_stream_t stream);

/* EMA / slow-weight pass alone (frozen parameters that only the reference's EMA
 * loops touch, e.g. DiT pos_embed: DiT/forget.py:58-62). */
SFR_API int sfr_ema_update(const float* p, float* ema, int64_t n, int ema_mode, double ema_a,
                   sfr_stream_t stream);

/* ===========================================================================
 * Consumers next to the hot path
 *
 * EWC / Selective-Amnesia penalty — DDPM/runners/diffusion.py:424-433 (sa_forget):
 *   per step, per tensor:  _loss = fisher * (param - params_mle)**2 ; loss += lmbda * _loss.sum()
 * One pass adds autograd's gradient of that term, (lmbda*F) * (2*(p - p_star)), into g and
 * accumulates the penalty value lmbda * sum(F * (p - p_star)^2) into *penalty (device double, may be NULL).
 *
 * Proximal-gradient shrink — SD/train-scripts/proximal_gradient.py:151-183:
 *   threshold = k-th smallest |theta - theta0|: run the K2b select with key mode SFR_KEY_ABSDIFF and
 *   k' = n - k + 1 (k-th smallest = k'-th largest), read it with sfr_select_threshold_value, then
 *   sfr_soft_threshold: d = p - p0 ; d > thr: d -= thr ; d < -thr: d += thr ; else 0 ; p = d + p0.
 * ======================================================================== */
SFR_API int sfr_ewc_penalty(const float* p, const float* p_star, const float* fisher, float* g,
                    int64_t n, float lambda, double* penalty, sfr_stream_t stream);
SFR_API int sfr_select_threshold_value(const sfr_select_state* state_dev, float* out_dev,
                               sfr_stream_t stream);
SFR_API int sfr_soft_threshold(float* p, const float* p0, int64_t n, const float* threshold_dev,
                       sfr_stream_t stream);

/* ===========================================================================
 * Flat-gradient capture: gather `count` gradient tensors into the flat vector
 *   flat[offsets[t] .. offsets[t] + sizes[t]) = src[t][0 .. sizes[t])
 * (one launch instead of one copy per named_parameter).
 * srcs/offsets/sizes are DEVICE arrays of length `count`.
 * ======================================================================== */
SFR_API int sfr_gather_segments(float* flat, const void* const* srcs_dev,
                        const int64_t* offsets_dev, const int64_t* sizes_dev,
                        int32_t count, int src_dtype, int64_t total, sfr_stream_t stream);

/* ===========================================================================
 * Cross-GPU exchange fused with the kernels it feeds (one process per GPU, NVLink peer memory)
 * replaces  torch.nn.DataParallel's reduce_add of the per-GPU gradients onto GPU 0 and its weight
 *           re-broadcast before the next forward
 *   DiT/forget.py:193,285-322 ; DiT/generate_fisher.py:173,216-291 ;
 *   DDPM/runners/diffusion.py:110,1060,1126-1180,1270-1281
 * The flat vector is sharded: rank r owns elements [lo, lo + n_local), lo a multiple of 16.
 *
 * sfr_peer_buf describes ONE symmetric buffer as this rank sees it: ptrs[r] is the device address at
 * which rank r's copy of the buffer is mapped into THIS process (ptrs[rank] is the local copy), and
 * `multicast` is the NVLS multicast address of the same buffer (NULL if the fabric has none).  The
 * caller obtains them from its allocator (torch.distributed._symmetric_memory, cuMem* + cuMulticast*,
 * or cudaIpc*); the library only dereferences them.  Transport per call:
 *   SFR_XP_P2P       loads from / stores to the `world` mapped pointers; gradients are summed in rank
 *                    order 0..world-1 in fp32 (deterministic: equals a sequential sum)
 *   SFR_XP_MULTIMEM  multimem.ld_reduce / multimem.st on the multicast address: the NVSwitch sums the
 *                    gradients (fp32 accumulation; bf16 results are rounded to bf16) and replicates the
 *                    weight stores
 *   SFR_XP_TMA       the same mapped pointers and the same rank-order fp32 sum as SFR_XP_P2P, but the NVLink
 *                    traffic is moved by TMA bulk copies (cp.async.bulk + mbarrier) through a shared-memory
 *                    ring / staging buffers, off the load-store pipeline that streams the shard's local
 *                    state — so HBM time and NVLink time overlap inside one kernel.  sfr_peer_fused_update
 *                    uses it for both directions or not at all (g_transport == bc_transport)
 * `average` != 0 divides the sum by `world` (true division), the mean DataParallel's loss takes over
 * the global batch.
 *
 * Ordering is the caller's: sfr_peer_barrier on the same stream before the first kernel that reads
 * peer gradients (every rank has finished writing them) and after the last kernel that reads them or
 * pushes weights (nobody overwrites a gradient that is still being read; every weight store has landed).
 * ======================================================================== */
#define SFR_MAX_PEERS 8
#define SFR_XP_P2P 1
#define SFR_XP_MULTIMEM 2
#define SFR_XP_TMA 3

typedef struct sfr_peer_buf {
  void* ptrs[SFR_MAX_PEERS];
  void* multicast;
} sfr_peer_buf;

typedef struct sfr_peer_geom {
  int32_t world;
  int32_t rank;
  int64_t lo;      /* first element of this rank's shard in the full vector (multiple of 16) */
  int64_t n_local; /* elements in this rank's shard (ragged only at the end of the vector)   */
} sfr_peer_geom;

/* Cross-GPU barrier on `stream`, which also all-reduces up to 8 doubles: every rank contributes
 * vals_dev[0..nvals) and receives in sums_dev[j] the sum over ranks IN RANK ORDER (identical bits on
 * every rank) — the clip norm's all-reduce (one double) costs no extra launch.
 * `pad`: a symmetric buffer of sfr_peer_pad_bytes() bytes, zero-filled on every rank before first use and
 * used by ONE stream at a time.  The epoch lives in the pad, so captured launches replay correctly.
 * A rank that waits longer than timeout_ns (0 = 30 s) sets word 1 of its pad (u64) to 1 and returns:
 * the caller checks that status; the GPU never hangs on a dead peer. */
SFR_API int64_t sfr_peer_pad_bytes(void);
SFR_API int sfr_peer_barrier(const sfr_peer_buf* pad, int world, int rank, const double* vals_dev,
                     double* sums_dev, int nvals, uint64_t timeout_ns, sfr_stream_t stream);

/* Reduce-scatter fused with K1 and the clip norm.  g: every rank's FULL flat gradient (dtype g_dtype).
 * For this rank's shard, gbar = sum_r g_r [/ world], then any of (NULL = skip):
 *   g_red[i]      = gbar[i]                                (fp32 reduced shard, local, n_local elements)
 *   fisher_acc[i] += gbar[i]**2 / fisher_divisor           (K1; same rounding sequence as sfr_fisher_accum)
 *   *sumsq        += sum_i (gbar[i] * mask[i])**2          (mask may be NULL; same as sfr_masked_sumsq)
 * mask / g_red / fisher_acc are LOCAL shard pointers (element 0 = vector element lo). */
/* max_ctas (both calls below): upper bound on the CTAs of the launch, 0 = as many as the kernel wants.  A small
 * bound (16-32) lets an exchange kernel run BESIDE other work — the backward pass on another stream — instead of
 * taking every SM; the TMA transport keeps NVLink busy from few CTAs. */
SFR_API int sfr_peer_reduce(const sfr_peer_buf* g, int g_dtype, const sfr_peer_geom* geom, int transport,
                    int average, float* g_red, const uint8_t* mask, double* sumsq,
                    float* fisher_acc, float fisher_divisor, int max_ctas, sfr_stream_t stream);

/* K3 on this rank's shard with the exchange on both sides.  Gradient source: `g_red` (local fp32 shard
 * left by sfr_peer_reduce — the clipped steps, whose norm must be known first) or, when g != NULL, the
 * peers' full gradients reduced on the fly (g_transport, average).  p / m / v / mask / ema are local
 * shard pointers.  Every updated weight is also pushed into all ranks' full-vector buffers: bc_f32
 * (fp32 weights; NULL = none) and/or bc_bf16 (bf16 working copy; NULL = none) via bc_transport.
 * args->flags: SFR_F_MASK | SFR_F_MASK_AFTER_CLIP | SFR_F_SGD_FIRST_STEP | SFR_F_REUSE_CONSTS only.  Everything else as
 * sfr_fused_update (same arithmetic: the two share their per-element code). */
SFR_API int sfr_peer_fused_update(float* p, const float* g_red, const sfr_peer_buf* g, int g_dtype,
                    int g_transport, int average, float* m, float* v, const uint8_t* mask,
                    float* ema, const sfr_peer_buf* bc_f32, const sfr_peer_buf* bc_bf16,
                    int bc_transport, const sfr_peer_geom* geom, const sfr_update_args* args,
                    const double* clip_sumsq, long long* step_counter_dev,
                    void* consts_scratch_dev, int max_ctas, sfr_stream_t stream);

/* All-gather alone: push this rank's shard (src_local, n_local elements of elem_bytes = 2 | 4) into every
 * rank's full-vector buffer `dst` at element offset lo. */
SFR_API int sfr_peer_broadcast(const void* src_local, const sfr_peer_buf* dst, int elem_bytes,
                    const sfr_peer_geom* geom, int transport, sfr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SFRON_B200_H */
