// Hidden-layer activations of the fp32 per-layer path and their derivatives in terms of the activation's VALUE.
#pragma once

// hidden-layer activations (VAEB_ACT_*: 1 tanh -- the reference's VAEB.py:246,254 --, 2 sigmoid, 3 ReLU: the
// alternatives of Report/replication/replic.tex:73-82) and their derivatives in terms of the activation's VALUE
__device__ __forceinline__ float act_fwd(float a, int act) {
  return act == 1 ? tanhf(a) : (act == 2 ? 1.0f / (1.0f + expf(-a)) : (act == 3 ? fmaxf(a, 0.f) : a));
}
__device__ __forceinline__ float act_bwd(float h, int act) {
  return act == 1 ? 1.0f - h * h : (act == 2 ? h * (1.0f - h) : (act == 3 ? (h > 0.f ? 1.0f : 0.f) : 1.0f));
}
