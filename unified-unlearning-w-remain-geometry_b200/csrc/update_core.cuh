// update_core.cuh — the per-element arithmetic of K3 (optimizer step + slow/EMA weights) and the scalars
// it needs, shared by update.cu (shard-local launch) and peer.cu (the same update fused with the
// cross-GPU gradient reduction and weight broadcast).
//
// Arithmetic contract (torch 2.11 CPU kernels, probed op by op):
//   add(b, alpha)   -> fma(b, alpha, a)          lerp(w<.5) -> fma(end-start, w, start)
//   addcmul(value)  -> fma(value*t1, t2, self)   addcdiv    -> self + (value*t1)/t2   (no fma)
//   mul / div / sqrt / add with a Python scalar -> the scalar rounded to fp32 first.
#pragma once

#include <math.h>

#include "common.cuh"

namespace sfr {

// ------------------------------------------------------------------------- K3 constants
struct UpdateConsts {
  float neg_lr;         // SGD     : fp32(-lr)
  float wd;             // SGD/Adam: fp32(weight_decay)      grad.add(param, alpha=wd)
  float decay_mul;      // AdamW   : fp32(1 - lr*wd)         param.mul_(1 - lr*wd)
  float lerp_w;         // Adam    : fp32(1 - beta1)         exp_avg.lerp_(grad, 1-beta1)
  float beta2;          //           fp32(beta2)             exp_avg_sq.mul_(beta2)
  float one_m_beta2;    //           fp32(1 - beta2)         .addcmul_(grad, grad, value=1-beta2)
  float bc2_sqrt;       //           fp32((1 - beta2**t)**0.5)
  float eps;            //           fp32(eps)
  float neg_step_size;  //           fp32(-(lr / (1 - beta1**t)))
  float momentum;       // SGD     : fp32(momentum)
  float one_m_damp;     // SGD     : fp32(1 - dampening)
  float ema_c1;         // mode-dependent, see ema_step
  float ema_c2;
  float max_norm;
  uint32_t flags;
  int has_wd;
  int has_momentum;
};

// Step-dependent scalars that a one-thread prep kernel leaves in device scratch (when the caller
// provides one): Adam's bias corrections from the DEVICE step counter (graph replay), SGD's
// first-step flag, and the clip coefficient.  Everything the kernel needs to decide WHICH loads to
// issue stays in the by-value UpdateConsts; these values are first used in the arithmetic, after the
// data loads are in flight, so reading them never stalls the start of a CTA (a coefficient derived
// in-kernel from *clip_sumsq does: load -> double sqrt -> divide in front of every short-lived CTA
// costs ~3.5 %, measured).
struct DevConsts {
  float neg_step_size;
  float bc2_sqrt;
  float clip_coef;
  uint32_t sgd_first_step;
  float neg_lr;       // the learning-rate-dependent scalars, valid when has_lr (device-side schedule)
  float decay_mul;
  uint32_t has_lr;
};

// Scalars a kernel must take from the device copy instead of its by-value UpdateConsts.
template <int OPT>
__device__ __forceinline__ void apply_dev_consts(UpdateConsts& c, const DevConsts* c_dev) {
  if constexpr (OPT == SFR_OPT_SGD) {
    // only SGD's momentum-buffer init changes WHICH loads are issued
    if (c_dev->sgd_first_step) c.flags |= SFR_F_SGD_FIRST_STEP; else c.flags &= ~SFR_F_SGD_FIRST_STEP;
    if (c_dev->has_lr) c.neg_lr = c_dev->neg_lr;
  } else {
    c.neg_step_size = c_dev->neg_step_size;
    c.bc2_sqrt = c_dev->bc2_sqrt;
    if (c_dev->has_lr) c.decay_mul = c_dev->decay_mul;
  }
}

// The step's learning rate: from the device-side schedule when one is attached.
__device__ __forceinline__ void resolve_lr(sfr_update_args& a) {
  if (a.lr_table_dev != nullptr) a.lr = a.lr_table_dev[*a.lr_index_dev];
}

// ---- slow / EMA weights ----------------------------------------------------------------
// Returns the new slow value; may rewrite p (SLOWFAST).
template <int EMA>
__device__ __forceinline__ float ema_step(float& p, float s, const UpdateConsts& c) {
  if constexpr (EMA == SFR_EMA_DDPM) {
    // shadow = (1.0 - mu) * param + mu * shadow         DDPM/models/ema.py:22-24
    return __fadd_rn(__fmul_rn(c.ema_c1, p), __fmul_rn(c.ema_c2, s));
  } else if constexpr (EMA == SFR_EMA_DIT) {
    // ema.mul_(decay).add_(param, alpha=1 - decay)      DiT/forget.py:62
    return __fmaf_rn(p, c.ema_c2, __fmul_rn(s, c.ema_c1));
  } else if constexpr (EMA == SFR_EMA_SLOWFAST) {
    // p = (1 - beta) * p_prev + beta * p ; p_prev = copy(p)   sfron.py:126-127,30-37,255-257
    p = __fadd_rn(__fmul_rn(c.ema_c1, s), __fmul_rn(c.ema_c2, p));
    return p;
  } else {
    return s;
  }
}

// ---- one parameter element ---------------------------------------------------------------
template <int OPT>
__device__ __forceinline__ void opt_step(float& p, float g, float& m, float& v,
                                         const UpdateConsts& c) {
  if constexpr (OPT == SFR_OPT_SGD) {
    // torch/optim/sgd.py _single_tensor_sgd
    if (c.has_wd) g = __fmaf_rn(p, c.wd, g);  // grad.add(param, alpha=wd)
    if (c.has_momentum) {
      if (c.flags & SFR_F_SGD_FIRST_STEP) {
        m = g;  // buf = clone(grad)
      } else {
        m = __fmaf_rn(g, c.one_m_damp, __fmul_rn(m, c.momentum));  // buf.mul_(mom).add_(grad, alpha=1-damp)
      }
      g = m;
    }
    p = __fmaf_rn(g, c.neg_lr, p);  // param.add_(grad, alpha=-lr)
  } else {
    // torch/optim/adam.py _single_tensor_adam (capturable=False, amsgrad=False)
    if (c.has_wd) {
      if constexpr (OPT == SFR_OPT_ADAMW) {
        p = __fmul_rn(p, c.decay_mul);  // param.mul_(1 - lr*wd)
      } else {
        g = __fmaf_rn(p, c.wd, g);  // grad.add(param, alpha=wd)
      }
    }
    // exp_avg.lerp_(grad, w):  |w| < 0.5 ? fma(diff, w, start) : fma(diff, w-1, end)
    const float diff = __fsub_rn(g, m);
    m = fabsf(c.lerp_w) < 0.5f ? __fmaf_rn(diff, c.lerp_w, m)
                               : __fmaf_rn(diff, __fsub_rn(c.lerp_w, 1.0f), g);
    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
    v = __fmaf_rn(__fmul_rn(c.one_m_beta2, g), g, __fmul_rn(v, c.beta2));
    // denom = (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), c.bc2_sqrt), c.eps);
    // param.addcdiv_(exp_avg, denom, value=-step_size)
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(c.neg_step_size, m), denom));
  }
}

template <int OPT, int EMA>
__device__ __forceinline__ void update_one(float& p, float g, float& m, float& v, float& s,
                                           float maskf, float coef, const UpdateConsts& c) {
  if (c.flags & SFR_F_MASK) g = __fmul_rn(g, maskf);              // param.grad *= mask
  g = __fmul_rn(g, coef);                                           // clip: grad.mul_(coef)
  if (c.flags & SFR_F_MASK_AFTER_CLIP) g = __fmul_rn(g, maskf);   // SalUn-DDPM order
  opt_step<OPT>(p, g, m, v, c);
  s = ema_step<EMA>(p, s, c);
}

__host__ __device__ inline void fill_ema_consts(UpdateConsts& c, int ema_mode, double a) {
  // Python computes (1 - a) in double; torch rounds each scalar to fp32 when the op runs.
  if (ema_mode == SFR_EMA_DDPM) {         // (1.0 - mu) * p + mu * s
    c.ema_c1 = (float)(1.0 - a);
    c.ema_c2 = (float)a;
  } else if (ema_mode == SFR_EMA_DIT) {   // s.mul_(d).add_(p, alpha=1 - d)
    c.ema_c1 = (float)a;
    c.ema_c2 = (float)(1.0 - a);
  } else if (ema_mode == SFR_EMA_SLOWFAST) {  // (1 - b) * prev + b * p
    c.ema_c1 = (float)(1.0 - a);
    c.ema_c2 = (float)a;
  } else {
    c.ema_c1 = c.ema_c2 = 0.f;
  }
}

// Scalars exactly as torch's Python forms them (double), rounded to fp32 where the ATen kernel would
// round them.  __host__ __device__: the device twin serves graph-replayable launches (device pow() is
// within 2 ulp of glibc's in double, far below the fp32 rounding that follows).
__host__ __device__ inline UpdateConsts make_update_consts(const sfr_update_args& a, int64_t step_i,
                                                           bool has_momentum) {
  UpdateConsts c{};
  c.flags = a.flags;
  c.has_wd = a.weight_decay != 0.0;
  c.has_momentum = has_momentum;
  c.max_norm = (float)a.clip_max_norm;
  c.wd = (float)a.weight_decay;
  if (a.opt == SFR_OPT_SGD) {
    c.neg_lr = (float)(-a.lr);
    c.momentum = (float)a.momentum;
    c.one_m_damp = (float)(1.0 - a.dampening);
  } else {
    const double step = (double)step_i;
    const double bc1 = 1.0 - pow(a.beta1, step);        // 1 - beta1 ** step
    const double bc2 = 1.0 - pow(a.beta2, step);        // 1 - beta2 ** step
    const double step_size = a.lr / bc1;                // lr / bias_correction1
    c.neg_step_size = (float)(-step_size);
    c.bc2_sqrt = (float)pow(bc2, 0.5);                  // bias_correction2 ** 0.5
    c.lerp_w = (float)(1.0 - a.beta1);
    c.beta2 = (float)a.beta2;
    c.one_m_beta2 = (float)(1.0 - a.beta2);
    c.eps = (float)a.eps;
    c.decay_mul = (float)(1.0 - a.lr * a.weight_decay);
  }
  fill_ema_consts(c, a.ema_mode, a.ema_a);
  return c;
}


// One-thread prep kernel (update.cu): leaves the step-dependent scalars and the clip coefficient in
// `out` (DevConsts); increments the device step counter when one is given (CUDA-graph replay).
void launch_update_consts(const sfr_update_args& a, bool has_momentum, long long* step_counter,
                          const double* clip_sumsq, void* consts_scratch, cudaStream_t s);

}  // namespace sfr
