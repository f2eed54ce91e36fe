// EWC Fisher accumulation / consolidation / quadratic penalty (+ its gradient), and a fused AdamW
// step, over flat fp32 state.  All kernels are pure streaming (12-20 B per parameter): float4
// accesses, grid sized to the SM count, one launch per <=32 model tensors (pointers travel in the
// kernel parameter block, so nothing is copied or retained).
#include "common.cuh"

using namespace nv;

namespace {

constexpr int MT = 32;  // tensors per launch

struct TensorTable {
  const float* src[MT];   // per-tensor device pointer (theta_i or grad_i)
  float* dst[MT];         // per-tensor output pointer (penalty_bwd only)
  int64_t off[MT];        // offset of tensor i in the flat state
  int64_t n[MT];
};

__device__ __forceinline__ bool vec_ok(const void* a, const void* b, const void* c) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

// fisher[off+k] += scale * g[k]^2
__global__ void __launch_bounds__(256) fisher_accum_kernel(TensorTable tb, float* __restrict__ fisher, float scale) {
  const int ti = blockIdx.y;
  const float* __restrict__ g = tb.src[ti];
  float* __restrict__ f = fisher + tb.off[ti];
  const int64_t n = tb.n[ti];
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok(g, f, nullptr)) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      float4 gv = reinterpret_cast<const float4*>(g)[i];
      float4 fv = reinterpret_cast<float4*>(f)[i];
      fv.x = fmaf(scale * gv.x, gv.x, fv.x);
      fv.y = fmaf(scale * gv.y, gv.y, fv.y);
      fv.z = fmaf(scale * gv.z, gv.z, fv.z);
      fv.w = fmaf(scale * gv.w, gv.w, fv.w);
      reinterpret_cast<float4*>(f)[i] = fv;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) f[i] = fmaf(scale * g[i], g[i], f[i]);
  } else {
    for (int64_t i = tid; i < n; i += stride) f[i] = fmaf(scale * g[i], g[i], f[i]);
  }
}

__global__ void __launch_bounds__(256) axpby_kernel(float* __restrict__ v, const float* __restrict__ w, int64_t n,
                                                    float a, float b) {
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok(v, w, nullptr)) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      float4 x = reinterpret_cast<float4*>(v)[i];
      if (w) {
        float4 y = reinterpret_cast<const float4*>(w)[i];
        x = make_float4(a * x.x + b * y.x, a * x.y + b * y.y, a * x.z + b * y.z, a * x.w + b * y.w);
      } else {
        x = make_float4(a * x.x, a * x.y, a * x.z, a * x.w);
      }
      reinterpret_cast<float4*>(v)[i] = x;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) v[i] = w ? a * v[i] + b * w[i] : a * v[i];
  } else {
    for (int64_t i = tid; i < n; i += stride) v[i] = w ? a * v[i] + b * w[i] : a * v[i];
  }
}

// *out += coef * sum F*(theta-star)^2
__global__ void __launch_bounds__(256) penalty_fwd_kernel(TensorTable tb, const float* __restrict__ fisher,
                                                          const float* __restrict__ star, float coef,
                                                          float* __restrict__ out) {
  const int ti = blockIdx.y;
  const float* __restrict__ th = tb.src[ti];
  const float* __restrict__ f = fisher + tb.off[ti];
  const float* __restrict__ st = star + tb.off[ti];
  const int64_t n = tb.n[ti];
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  double acc = 0.0;
  if (vec_ok(th, f, st)) {
    const int64_t n4 = n >> 2;
    float part = 0.f;
    int cnt = 0;
    for (int64_t i = tid; i < n4; i += stride) {
      float4 t = reinterpret_cast<const float4*>(th)[i];
      float4 fv = reinterpret_cast<const float4*>(f)[i];
      float4 s = reinterpret_cast<const float4*>(st)[i];
      float dx = t.x - s.x, dy = t.y - s.y, dz = t.z - s.z, dw = t.w - s.w;
      part += fv.x * dx * dx + fv.y * dy * dy + fv.z * dz * dz + fv.w * dw * dw;
      if (++cnt == 16) { acc += part; part = 0.f; cnt = 0; }
    }
    acc += part;
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
      float d = th[i] - st[i];
      acc += (double)(f[i] * d * d);
    }
  } else {
    for (int64_t i = tid; i < n; i += stride) {
      float d = th[i] - st[i];
      acc += (double)(f[i] * d * d);
    }
  }
  acc = warp_sum(acc);
  __shared__ double part_s[8];
  if ((threadIdx.x & 31) == 0) part_s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += part_s[i];
    if (t != 0.0) atomicAdd(out, (float)(t * (double)coef));
  }
}

// grad[k] += gscale * coef2 * F*(theta-star)
__global__ void __launch_bounds__(256) penalty_bwd_kernel(TensorTable tb, const float* __restrict__ fisher,
                                                          const float* __restrict__ star, float coef2,
                                                          const float* __restrict__ gscale) {
  const int ti = blockIdx.y;
  const float* __restrict__ th = tb.src[ti];
  float* __restrict__ gr = tb.dst[ti];
  const float* __restrict__ f = fisher + tb.off[ti];
  const float* __restrict__ st = star + tb.off[ti];
  const int64_t n = tb.n[ti];
  const float k = coef2 * (gscale ? __ldg(gscale) : 1.f);
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok(th, f, st) && vec_ok(gr, nullptr, nullptr)) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) {
      float4 t = reinterpret_cast<const float4*>(th)[i];
      float4 fv = reinterpret_cast<const float4*>(f)[i];
      float4 s = reinterpret_cast<const float4*>(st)[i];
      float4 g = reinterpret_cast<float4*>(gr)[i];
      g.x = fmaf(k * fv.x, t.x - s.x, g.x);
      g.y = fmaf(k * fv.y, t.y - s.y, g.y);
      g.z = fmaf(k * fv.z, t.z - s.z, g.z);
      g.w = fmaf(k * fv.w, t.w - s.w, g.w);
      reinterpret_cast<float4*>(gr)[i] = g;
    }
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) gr[i] = fmaf(k * f[i], th[i] - st[i], gr[i]);
  } else {
    for (int64_t i = tid; i < n; i += stride) gr[i] = fmaf(k * f[i], th[i] - st[i], gr[i]);
  }
}

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, float lr, float b1, float b2, float eps,
                                          float wd, float bc1, float bc2_sqrt, float gscale) {
  const float gi = g * gscale;
  float pi = p;
  pi *= (1.f - lr * wd);                      // decoupled weight decay
  const float mi = fmaf(b1, m, (1.f - b1) * gi);
  const float vi = fmaf(b2, v, (1.f - b2) * gi * gi);
  m = mi;
  v = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p = pi - (lr / bc1) * (mi / denom);
}

// 28 B per parameter (p, g, m, v read; p, m, v written): 128-bit accesses, two independent vectors in flight per thread
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, int64_t n,
                                                    float lr, float b1, float b2, float eps, float wd, float bc1,
                                                    float bc2_sqrt, float gscale, const int32_t* __restrict__ step_dev) {
  if (step_dev) {          // step count kept on the device (CUDA-graph replays): bias corrections computed here
    const float st = (float)__ldg(step_dev);
    bc1 = 1.f - powf(b1, st);
    bc2_sqrt = sqrtf(1.f - powf(b2, st));
  }
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok(p, g, m) && vec_ok(v, nullptr, nullptr)) {
    const int64_t n4 = n >> 2;
    float4* p4 = reinterpret_cast<float4*>(p);
    const float4* g4 = reinterpret_cast<const float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    int64_t i = tid;
    for (; i + stride < n4; i += 2 * stride) {
      float4 pa = p4[i], ga = g4[i], ma = m4[i], va = v4[i];
      float4 pb = p4[i + stride], gb = g4[i + stride], mb = m4[i + stride], vb = v4[i + stride];
      adamw_one(pa.x, ga.x, ma.x, va.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pa.y, ga.y, ma.y, va.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pa.z, ga.z, ma.z, va.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pa.w, ga.w, ma.w, va.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pb.x, gb.x, mb.x, vb.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pb.y, gb.y, mb.y, vb.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pb.z, gb.z, mb.z, vb.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pb.w, gb.w, mb.w, vb.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      p4[i] = pa; m4[i] = ma; v4[i] = va;
      p4[i + stride] = pb; m4[i + stride] = mb; v4[i + stride] = vb;
    }
    for (; i < n4; i += stride) {
      float4 pa = p4[i], ga = g4[i], ma = m4[i], va = v4[i];
      adamw_one(pa.x, ga.x, ma.x, va.x, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pa.y, ga.y, ma.y, va.y, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pa.z, ga.z, ma.z, va.z, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      adamw_one(pa.w, ga.w, ma.w, va.w, lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
      p4[i] = pa; m4[i] = ma; v4[i] = va;
    }
    for (int64_t k = (n4 << 2) + tid; k < n; k += stride) adamw_one(p[k], g[k], m[k], v[k], lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
  } else {
    for (int64_t k = tid; k < n; k += stride) adamw_one(p[k], g[k], m[k], v[k], lr, b1, b2, eps, wd, bc1, bc2_sqrt, gscale);
  }
}

// Synaptic Intelligence running importance (reference ewc.py:342-352):  W += -g * (theta - p_old);  p_old = theta
__global__ void __launch_bounds__(256) si_update_kernel(TensorTable tb, float* __restrict__ W, float* __restrict__ pold) {
  const int ti = blockIdx.y;
  const float* __restrict__ th = tb.src[ti];
  const float* __restrict__ g = tb.dst[ti];       // (read-only here: the table's second pointer column)
  float* __restrict__ w = W + tb.off[ti];
  float* __restrict__ po = pold + tb.off[ti];
  const int64_t n = tb.n[ti];
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n; i += stride) {
    const float t = th[i];
    if (g) {                                        // a parameter without a gradient keeps W AND p_old (ewc.py:347)
      w[i] = w[i] + (-g[i]) * (t - po[i]);
      po[i] = t;
    }
  }
}

// Synaptic Intelligence consolidation (reference ewc.py:354-366):
//   omega += W / ((theta - p_old)^2 + damping);  W = 0;  p_old = theta
__global__ void __launch_bounds__(256) si_register_kernel(TensorTable tb, float* __restrict__ W, float* __restrict__ pold,
                                                          float* __restrict__ omega, float damping) {
  const int ti = blockIdx.y;
  const float* __restrict__ th = tb.src[ti];
  float* __restrict__ w = W + tb.off[ti];
  float* __restrict__ po = pold + tb.off[ti];
  float* __restrict__ om = omega + tb.off[ti];
  const int64_t n = tb.n[ti];
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = tid; i < n; i += stride) {
    const float t = th[i];
    const float d = t - po[i];
    om[i] = om[i] + w[i] / (d * d + damping);
    w[i] = 0.f;
    po[i] = t;
  }
}

// flat[off_i + k] = src_i[k]
__global__ void __launch_bounds__(256) flat_gather_kernel(TensorTable tb, float* __restrict__ flat) {
  const int ti = blockIdx.y;
  const float* __restrict__ g = tb.src[ti];
  float* __restrict__ f = flat + tb.off[ti];
  const int64_t n = tb.n[ti];
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  if (vec_ok(g, f, nullptr)) {
    const int64_t n4 = n >> 2;
    for (int64_t i = tid; i < n4; i += stride) reinterpret_cast<float4*>(f)[i] = reinterpret_cast<const float4*>(g)[i];
    for (int64_t i = (n4 << 2) + tid; i < n; i += stride) f[i] = g[i];
  } else {
    for (int64_t i = tid; i < n; i += stride) f[i] = g[i];
  }
}

inline dim3 table_grid(const TensorTable& tb, int cnt) {
  int64_t maxn = 1;
  for (int i = 0; i < cnt; ++i) maxn = imax(maxn, tb.n[i]);
  int gx = (int)imax(1, imin(cdiv(maxn, 256 * 4 * 4), (kSMs * 8) / cnt + 1));
  return dim3(gx, cnt);
}

// Walks `ntensors` host-side table entries in chunks of at most MT LIVE tensors and calls launch(table, count) per
// chunk.  One running index crosses the chunks, so skipped entries (NULL first pointer when `skip_null`, or zero
// elements) can neither be visited twice nor shift the flat offsets of what follows.  `second` may be NULL.
// `offsets` (nullable) overrides the running offsets (a flat layout with padded slots).
template <typename Launch>
int for_each_chunk(const float* const* first, float* const* second, const int64_t* numel, int ntensors, bool skip_null,
                   bool second_optional, Launch&& launch, const int64_t* offsets = nullptr) {
  int64_t off = 0;
  int i = 0;
  while (i < ntensors) {
    TensorTable tb;
    int cnt = 0;
    for (; i < ntensors && cnt < MT; ++i) {
      if (numel[i] < 0) return NERVECL_EINVAL;
      if (numel[i] > 0) {                          // (an empty tensor may legitimately have a NULL data pointer)
        if (!first[i] && !skip_null) return NERVECL_EINVAL;
        if (second && !second[i] && !second_optional) return NERVECL_EINVAL;
      }
      if (first[i] && numel[i] > 0) {
        tb.src[cnt] = first[i];
        tb.dst[cnt] = second ? second[i] : nullptr;
        tb.off[cnt] = offsets ? offsets[i] : off;
        tb.n[cnt] = numel[i];
        ++cnt;
      }
      off += numel[i];
    }
    if (cnt == 0) continue;
    int rc = launch(tb, cnt);
    if (rc) return rc;
  }
  return NERVECL_OK;
}

}  // namespace

NV_API int nervecl_ewc_fisher_accum(float* fisher, const float* const* grads_host, const int64_t* numel_host,
                                    int ntensors, float scale, nervecl_stream_t stream) {
  if (!fisher || !grads_host || !numel_host || ntensors <= 0) return NERVECL_EINVAL;
  // a NULL grad (param.grad is None, ewc.py:140) is skipped; its slot of the flat Fisher keeps its value
  return for_each_chunk(grads_host, nullptr, numel_host, ntensors, true, false, [&](const TensorTable& tb, int cnt) {
    fisher_accum_kernel<<<table_grid(tb, cnt), 256, 0, as_stream(stream)>>>(tb, fisher, scale);
    return launch_status();
  });
}

NV_API int nervecl_ewc_axpby(float* v, const float* w, int64_t n, float a, float b, nervecl_stream_t stream) {
  if (!v || n <= 0) return NERVECL_EINVAL;
  int blocks = (int)imax(1, imin(cdiv(n, 256 * 16), kSMs * 8));
  axpby_kernel<<<blocks, 256, 0, as_stream(stream)>>>(v, w, n, a, b);
  return launch_status();
}

NV_API int nervecl_ewc_penalty_fwd(const float* const* theta_host, const int64_t* numel_host, int ntensors,
                                   const float* fisher, const float* star, float coef, float* out,
                                   nervecl_stream_t stream) {
  if (!theta_host || !numel_host || !fisher || !star || !out || ntensors <= 0) return NERVECL_EINVAL;
  return for_each_chunk(theta_host, nullptr, numel_host, ntensors, false, false, [&](const TensorTable& tb, int cnt) {
    penalty_fwd_kernel<<<table_grid(tb, cnt), 256, 0, as_stream(stream)>>>(tb, fisher, star, coef, out);
    return launch_status();
  });
}

NV_API int nervecl_ewc_penalty_bwd(const float* const* theta_host, float* const* grad_host,
                                   const int64_t* numel_host, int ntensors, const float* fisher, const float* star,
                                   float coef2, const float* gscale, nervecl_stream_t stream) {
  if (!theta_host || !grad_host || !numel_host || !fisher || !star || ntensors <= 0) return NERVECL_EINVAL;
  return for_each_chunk(theta_host, grad_host, numel_host, ntensors, false, false, [&](const TensorTable& tb, int cnt) {
    penalty_bwd_kernel<<<table_grid(tb, cnt), 256, 0, as_stream(stream)>>>(tb, fisher, star, coef2, gscale);
    return launch_status();
  });
}

NV_API int nervecl_flat_gather(const float* const* src_host, const int64_t* numel_host, const int64_t* offset_host,
                               int ntensors, float* flat, nervecl_stream_t stream) {
  if (!src_host || !numel_host || !offset_host || !flat || ntensors <= 0) return NERVECL_EINVAL;
  return for_each_chunk(src_host, nullptr, numel_host, ntensors, true, false, [&](const TensorTable& tb, int cnt) {
    flat_gather_kernel<<<table_grid(tb, cnt), 256, 0, as_stream(stream)>>>(tb, flat);
    return launch_status();
  }, offset_host);
}

NV_API int nervecl_si_update(const float* const* theta_host, const float* const* grad_host, const int64_t* numel_host,
                             int ntensors, float* W, float* p_old, nervecl_stream_t stream) {
  if (!theta_host || !grad_host || !numel_host || !W || !p_old || ntensors <= 0) return NERVECL_EINVAL;
  return for_each_chunk(theta_host, const_cast<float* const*>(grad_host), numel_host, ntensors, false, true,
                        [&](const TensorTable& tb, int cnt) {
    si_update_kernel<<<table_grid(tb, cnt), 256, 0, as_stream(stream)>>>(tb, W, p_old);
    return launch_status();
  });
}

NV_API int nervecl_si_register(const float* const* theta_host, const int64_t* numel_host, int ntensors, float* W,
                               float* p_old, float* omega, float damping, nervecl_stream_t stream) {
  if (!theta_host || !numel_host || !W || !p_old || !omega || ntensors <= 0) return NERVECL_EINVAL;
  return for_each_chunk(theta_host, nullptr, numel_host, ntensors, false, false, [&](const TensorTable& tb, int cnt) {
    si_register_kernel<<<table_grid(tb, cnt), 256, 0, as_stream(stream)>>>(tb, W, p_old, omega, damping);
    return launch_status();
  });
}

NV_API int nervecl_adamw_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                              float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                              float grad_scale, const int32_t* step_dev, nervecl_stream_t stream) {
  if (!param || !grad || !exp_avg || !exp_avg_sq || n <= 0 || (step < 1 && !step_dev)) return NERVECL_EINVAL;
  float bc1 = 1.f - powf(beta1, (float)step);
  float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  // one resident wave: 8 blocks of 256 threads per SM, every thread streams two 128-bit vectors per array per trip
  int blocks = (int)imax(1, imin(cdiv(n, 256 * 8), kSMs * 8));
  adamw_kernel<<<blocks, 256, 0, as_stream(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps,
                                                      weight_decay, bc1, bc2_sqrt, grad_scale, step_dev);
  return launch_status();
}
