// Latent-layer kernels of the AEVB step: everything between the two wide hidden layers, where
// the contractions are too thin (N = 2Z, K = Z) for tiles and the work is latency-bound.
//   latent_fwd : mu, ls = h_e.W4+b4, h_e.W5+b5 (VAEB.py:248-249); eps, z = mu+exp(.5 ls) eps
//                (VAEB.py:41-47); KL / LA row terms (VAEB.py:343, :322-325);
//                h_d = tanh(z.W1+b1) (VAEB.py:254)            -- one warp per datapoint
//   latent_bwd : dz = da1.W1^T; dmu, dls; da3 = (dmu.W4^T + dls.W5^T)*(1-h_e^2);
//                per-row bound and its deterministic total    -- one warp per datapoint
//   small_wgrad: gW1,gb1 = [z|1]^T.da1 ; gW4,gb4,gW5,gb5 = [h_e|1]^T.[dmu|dls]
// Optional bf16 hi/lo mirrors of h_d and da3 feed the tcgen05 GEMMs of the wide layers.
#include <cuda_bf16.h>

#include "launchers.h"
#include "philox.cuh"

namespace {

constexpr int NWARPS = 16;          // 512 threads per block
constexpr int NTHREADS = NWARPS * 32;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ void store_split(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t o, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[o] = h;
  if (lo) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// W45T = [W4^T ; W5^T] ([2Z, H], H contiguous): both latent heads and their backward read the
// head weights with the hidden index contiguous (coalesced over lanes).
__global__ void __launch_bounds__(256)
transpose_heads_kernel(const float* __restrict__ W4, const float* __restrict__ W5, int H, int Z,
                       float* __restrict__ w45t) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * Z * H) return;
  const int o = i / H, k = i % H;
  w45t[i] = (o < Z) ? W4[(size_t)k * Z + o] : W5[(size_t)k * Z + (o - Z)];
}

// A block owns RB datapoints.  Phase 1: the 2Z head pre-activations; warps split (chunk of 8
// outputs) x (slice of the hidden index), every weight is loaded once per block and reused for
// the RB rows.  Phase 2: reparameterisation + row terms (one thread per (row, j)).  Phase 3:
// decoder hidden layer, one thread per hidden unit, W1 column held in registers across the rows.
template <int RB>
__global__ void __launch_bounds__(NTHREADS)
latent_fwd_kernel(const float* __restrict__ h_e, int rows, int H, const float* __restrict__ w45t,
                  const float* __restrict__ b4, const float* __restrict__ b5, const float* __restrict__ W1,
                  const float* __restrict__ b1, int Z, int L, int la, EpsSource src, float* __restrict__ mu,
                  float* __restrict__ ls, float* __restrict__ eps, float* __restrict__ z,
                  float* __restrict__ row_aux, float* __restrict__ h_d, __nv_bfloat16* __restrict__ hd_hi,
                  __nv_bfloat16* __restrict__ hd_lo, int ld_mirror, int KS) {
  extern __shared__ float sm[];
  const int Z2 = 2 * Z;
  float* part = sm;                         // [KS][RB][2Z]
  float* out = part + KS * RB * Z2;         // [RB][2Z] mu | ls
  float* zs = out + RB * Z2;                // [RB][Z]  z of the current sample
  float* terms = zs + RB * Z;               // [RB][Z]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * RB;
  const int nchunks = (Z2 + 7) / 8;
  for (int task = warp; task < nchunks * KS; task += NWARPS) {
    const int c0 = (task % nchunks) * 8, sl = task / nchunks;
    const int nq = min(8, Z2 - c0);
    const int k_lo = (int)(((long)H * sl) / KS), k_hi = (int)(((long)H * (sl + 1)) / KS);
    float acc[RB][8];
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[r][q] = 0.f;
    for (int k = k_lo + lane; k < k_hi; k += 32) {
      float wv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) wv[q] = (q < nq) ? w45t[(size_t)(c0 + q) * H + k] : 0.f;
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        const float hv = (m0 + r < rows) ? h_e[(size_t)(m0 + r) * H + k] : 0.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[r][q] = fmaf(hv, wv[q], acc[r][q]);
      }
    }
#pragma unroll
    for (int r = 0; r < RB; ++r)
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float t = warp_sum(acc[r][q]);
        if (lane == 0 && q < nq) part[(sl * RB + r) * Z2 + c0 + q] = t;
      }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RB * Z2; i += NTHREADS) {
    const int o = i % Z2;
    float t = (o < Z) ? b4[o] : b5[o - Z];
    for (int sl = 0; sl < KS; ++sl) t += part[sl * RB * Z2 + i];
    out[i] = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < RB * Z; i += NTHREADS) {
    const int r = i / Z, j = i % Z, m = m0 + r;
    float term = 0.f;
    if (m < rows) {
      const float am = out[r * Z2 + j], al = out[r * Z2 + Z + j];
      mu[(size_t)m * Z + j] = am;
      ls[(size_t)m * Z + j] = al;
      if (!la) term = 0.5f * (1.0f + al - am * am - expf(al));
    }
    terms[i] = term;
  }
  for (int l = 0; l < L; ++l) {
    for (int i = threadIdx.x; i < RB * Z; i += NTHREADS) {
      const int r = i / Z, j = i % Z, m = m0 + r;
      if (m < rows) {
        const float am = out[r * Z2 + j], al = out[r * Z2 + Z + j];
        const size_t o2 = ((size_t)l * rows + m) * Z + j;
        const float e = src.injected
                            ? src.injected[o2]
                            : philox_normal1(src.seed, src.stream, src.step, (uint32_t)l,
                                             (uint64_t)((src.row_offset + m) * Z + j));
        const float zv = am + expf(0.5f * al) * e;
        eps[o2] = e;
        z[o2] = zv;
        zs[i] = zv;
        if (la) terms[i] += (-0.5f * zv * zv + 0.5f * al + 0.5f * e * e) / (float)L;
      }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < H; n += NTHREADS) {
      float a[RB];
      const float bn = b1[n];
#pragma unroll
      for (int r = 0; r < RB; ++r) a[r] = bn;
      for (int j = 0; j < Z; ++j) {
        const float wv = W1[(size_t)j * H + n];
#pragma unroll
        for (int r = 0; r < RB; ++r) a[r] = fmaf(zs[r * Z + j], wv, a[r]);
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (m0 + r < rows) {
          const size_t rr = (size_t)l * rows + m0 + r;
          const float hv = tanhf(a[r]);
          h_d[rr * H + n] = hv;
          if (hd_hi) store_split(hd_hi, hd_lo, rr * ld_mirror + n, hv);
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x < RB && m0 + threadIdx.x < rows) {
    float t = 0.f;
    for (int j = 0; j < Z; ++j) t += terms[threadIdx.x * Z + j];
    row_aux[m0 + threadIdx.x] = t;
  }
}

// deterministic sum of v[0..n) by one block (fixed strided partials + fixed tree)
__device__ float block_total(const float* __restrict__ v, int n, float* red) {
  float t = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += __ldcg(v + i);   // written by other blocks: bypass L1
  t = warp_sum(t);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) red[wid] = t;
  __syncthreads();
  float s = 0.f;
  if (wid == 0) {
    s = (lane < (int)(blockDim.x >> 5)) ? red[lane] : 0.f;
    s = warp_sum(s);
  }
  return s;
}

// A block owns RB datapoints: dz (warps split latent chunks x hidden slices), dmu/dls, da3 (one
// thread per hidden unit, head weights held in registers across the rows), the per-row bound; the
// last block to finish totals the per-row bounds in a fixed order.
template <int RB>
__global__ void __launch_bounds__(NTHREADS)
latent_bwd_kernel(const float* __restrict__ da1, const float* __restrict__ W1, const float* __restrict__ w45t,
                  const float* __restrict__ h_e, const float* __restrict__ z, const float* __restrict__ eps,
                  const float* __restrict__ mu, const float* __restrict__ ls, int rows, int H, int Z, int L, int la,
                  float w, float* __restrict__ dmu, float* __restrict__ dls, float* __restrict__ da3,
                  __nv_bfloat16* __restrict__ da3_hi, __nv_bfloat16* __restrict__ da3_lo, int ld_mirror,
                  const float* __restrict__ partial, int n_tiles, const float* __restrict__ row_aux,
                  float* __restrict__ per_row, unsigned int* __restrict__ counter, float* __restrict__ base_out,
                  float mult, const float* __restrict__ tprior, int n_tprior, float div,
                  float* __restrict__ scalar_out, int NSL) {
  extern __shared__ float sm[];
  __shared__ float red[NWARPS];
  __shared__ int is_last;
  float* part = sm;                    // [NSL][RB][Z]
  float* dm = part + NSL * RB * Z;     // [RB][Z] dmu
  float* dl = dm + RB * Z;             // [RB][Z] dls
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * RB;
  const float s = w / (float)L;
  const int nchunks = (Z + 7) / 8;
  for (int i = threadIdx.x; i < RB * Z; i += NTHREADS) { dm[i] = 0.f; dl[i] = 0.f; }
  for (int l = 0; l < L; ++l) {
    // dz[r][j] = sum_n da1[(l,m0+r), n] W1[j, n]
    for (int task = warp; task < nchunks * NSL; task += NWARPS) {
      const int c0 = (task % nchunks) * 8, sl = task / nchunks;
      const int nq = min(8, Z - c0);
      const int n_lo = (int)(((long)H * sl) / NSL), n_hi = (int)(((long)H * (sl + 1)) / NSL);
      float acc[RB][8];
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[r][q] = 0.f;
      for (int n = n_lo + lane; n < n_hi; n += 32) {
        float wv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) wv[q] = (q < nq) ? W1[(size_t)(c0 + q) * H + n] : 0.f;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
          const float dv = (m0 + r < rows) ? da1[((size_t)l * rows + m0 + r) * H + n] : 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) acc[r][q] = fmaf(dv, wv[q], acc[r][q]);
        }
      }
#pragma unroll
      for (int r = 0; r < RB; ++r)
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float t = warp_sum(acc[r][q]);
          if (lane == 0 && q < nq) part[(sl * RB + r) * Z + c0 + q] = t;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < RB * Z; i += NTHREADS) {
      const int r = i / Z, j = i % Z, m = m0 + r;
      if (m < rows) {
        float d = 0.f;
        for (int sl = 0; sl < NSL; ++sl) d += part[sl * RB * Z + i];
        const size_t o2 = ((size_t)l * rows + m) * Z + j;
        if (la) d -= s * z[o2];
        const float sd = expf(0.5f * ls[(size_t)m * Z + j]);
        dm[i] += d;
        dl[i] += d * (0.5f * sd * eps[o2]);
      }
    }
    __syncthreads();
  }
  for (int i = threadIdx.x; i < RB * Z; i += NTHREADS) {
    const int r = i / Z, j = i % Z, m = m0 + r;
    if (m < rows) {
      const float lsv = ls[(size_t)m * Z + j], muv = mu[(size_t)m * Z + j];
      float a = dm[i], b = dl[i];
      if (la) {
        b += w * 0.5f;
      } else {
        a -= w * muv;
        b += w * 0.5f * (1.0f - expf(lsv));
      }
      dm[i] = a;
      dl[i] = b;
      dmu[(size_t)m * Z + j] = a;
      dls[(size_t)m * Z + j] = b;
    }
  }
  __syncthreads();
  // da3[m,n] = (sum_j dmu_j W4[n,j] + dls_j W5[n,j]) * (1 - h_e^2)
  for (int n = threadIdx.x; n < H; n += NTHREADS) {
    float a[RB];
#pragma unroll
    for (int r = 0; r < RB; ++r) a[r] = 0.f;
    for (int j = 0; j < Z; ++j) {
      const float w4 = w45t[(size_t)j * H + n], w5 = w45t[(size_t)(Z + j) * H + n];
#pragma unroll
      for (int r = 0; r < RB; ++r) a[r] = fmaf(dm[r * Z + j], w4, fmaf(dl[r * Z + j], w5, a[r]));
    }
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      const int m = m0 + r;
      if (m < rows) {
        const float hv = h_e[(size_t)m * H + n];
        const float v = a[r] * (1.0f - hv * hv);
        da3[(size_t)m * H + n] = v;
        if (da3_hi) store_split(da3_hi, da3_lo, (size_t)m * ld_mirror + n, v);
      }
    }
  }
  // per-datapoint bound: (1/L) sum_l sum_tiles partial + row_aux   (one warp per row, from the top)
  for (int r = NWARPS - 1 - warp; r < RB; r += NWARPS) {
    const int m = m0 + r;
    if (m < rows) {
      float t = 0.f;
      for (int l = 0; l < L; ++l) {
        const float* p = partial + ((size_t)l * rows + m) * n_tiles;
        for (int q = lane; q < n_tiles; q += 32) t += p[q];
      }
      t = warp_sum(t);
      if (lane == 0) per_row[m] = t / (float)L + row_aux[m];
    }
  }
  // the last block to finish totals the rows in a fixed order (deterministic)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  const float base = block_total(per_row, rows, red);
  float tp = 0.f;
  if (tprior) { __syncthreads(); tp = block_total(tprior, n_tprior, red); }
  if (threadIdx.x == 0) {
    *base_out = base;
    if (scalar_out) *scalar_out = (mult * base + tp) / div;
    *counter = 0u;
  }
}

// Thin weight gradients.  A block owns 64 outputs; its 4 thread groups each take a quarter of the
// row chunk (independent loads, unrolled) and are summed through shared memory in a fixed order.
//   part 0: gW1[j,n] (j<Z) and gb1[n] (j==Z)          = sum_r [z|1][r,j] * da1[r,n]        ((Z+1)*H outputs)
//   part 1: gW4[k,j], gW5[k,j] (k<H), gb4, gb5 (k==H) = sum_m [h_e|1][m,k] * dmu/dls[m,j]  ((H+1)*Z outputs x2)
__global__ void __launch_bounds__(256)
small_wgrad_kernel(const float* __restrict__ z, const float* __restrict__ da1, int R, const float* __restrict__ h_e,
                   const float* __restrict__ dmu, const float* __restrict__ dls, int rows, int H, int Z,
                   int rows_per_chunk, float* __restrict__ gW1, float* __restrict__ gb1, float* __restrict__ gW4,
                   float* __restrict__ gb4, float* __restrict__ gW5, float* __restrict__ gb5,
                   float* __restrict__ scratch) {
  __shared__ float red[2][4][64];
  const int nA = (Z + 1) * H, nB = (H + 1) * Z;
  const int oi = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int i = blockIdx.x * 64 + oi;
  const int chunk = blockIdx.y, nchunks = gridDim.y;
  float a4 = 0.f, a5 = 0.f;
  if (i < nA) {
    const int j = i / H, n = i % H;
    const int Lr = R / rows;
    const int c0 = chunk * rows_per_chunk * Lr, c1 = min(R, c0 + rows_per_chunk * Lr);
    const int per = (c1 - c0 + 3) / 4;
    const int r0 = c0 + grp * per, r1 = min(c1, r0 + per);
    if (j < Z) {
#pragma unroll 8
      for (int r = r0; r < r1; ++r) a4 = fmaf(z[(size_t)r * Z + j], da1[(size_t)r * H + n], a4);
    } else {
#pragma unroll 8
      for (int r = r0; r < r1; ++r) a4 += da1[(size_t)r * H + n];
    }
  } else if (i < nA + nB) {
    const int o = i - nA;
    const int k = o / Z, j = o % Z;
    const int c0 = chunk * rows_per_chunk, c1 = min(rows, c0 + rows_per_chunk);
    const int per = (c1 - c0 + 3) / 4;
    const int r0 = c0 + grp * per, r1 = min(c1, r0 + per);
    if (k < H) {
#pragma unroll 8
      for (int r = r0; r < r1; ++r) {
        const float hv = h_e[(size_t)r * H + k];
        a4 = fmaf(hv, dmu[(size_t)r * Z + j], a4);
        a5 = fmaf(hv, dls[(size_t)r * Z + j], a5);
      }
    } else {
#pragma unroll 8
      for (int r = r0; r < r1; ++r) { a4 += dmu[(size_t)r * Z + j]; a5 += dls[(size_t)r * Z + j]; }
    }
  }
  red[0][grp][oi] = a4;
  red[1][grp][oi] = a5;
  __syncthreads();
  if (grp != 0) return;
  a4 = (red[0][0][oi] + red[0][1][oi]) + (red[0][2][oi] + red[0][3][oi]);
  a5 = (red[1][0][oi] + red[1][1][oi]) + (red[1][2][oi] + red[1][3][oi]);
  if (i < nA) {
    const int j = i / H, n = i % H;
    if (nchunks > 1) scratch[(size_t)chunk * (nA + 2 * nB) + i] = a4;
    else if (j < Z) gW1[(size_t)j * H + n] = a4;
    else gb1[n] = a4;
  } else if (i < nA + nB) {
    const int o = i - nA;
    const int k = o / Z, j = o % Z;
    if (nchunks > 1) {
      scratch[(size_t)chunk * (nA + 2 * nB) + nA + o] = a4;
      scratch[(size_t)chunk * (nA + 2 * nB) + nA + nB + o] = a5;
    } else if (k < H) {
      gW4[(size_t)k * Z + j] = a4;
      gW5[(size_t)k * Z + j] = a5;
    } else {
      gb4[j] = a4;
      gb5[j] = a5;
    }
  }
}

__global__ void __launch_bounds__(256)
small_wgrad_reduce_kernel(const float* __restrict__ scratch, int nchunks, int H, int Z, float* __restrict__ gW1,
                          float* __restrict__ gb1, float* __restrict__ gW4, float* __restrict__ gb4,
                          float* __restrict__ gW5, float* __restrict__ gb5) {
  const int nA = (Z + 1) * H, nB = (H + 1) * Z, tot = nA + 2 * nB;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tot) return;
  float a = 0.f;
  for (int c = 0; c < nchunks; ++c) a += scratch[(size_t)c * tot + i];
  if (i < nA) {
    const int j = i / H, n = i % H;
    if (j < Z) gW1[(size_t)j * H + n] = a; else gb1[n] = a;
  } else {
    const int o = (i - nA) % nB;
    float* gW = (i - nA) < nB ? gW4 : gW5;
    float* gb = (i - nA) < nB ? gb4 : gb5;
    const int k = o / Z, j = o % Z;
    if (k < H) gW[(size_t)k * Z + j] = a; else gb[j] = a;
  }
}

}  // namespace

cudaError_t launch_transpose_heads(cudaStream_t st, int64_t* launches, const float* W4, const float* W5, int H, int Z,
                                   float* w45t) {
  transpose_heads_kernel<<<(2 * Z * H + 255) / 256, 256, 0, st>>>(W4, W5, H, Z, w45t);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_latent_fwd(cudaStream_t st, int64_t* launches, const float* h_e, int rows, int H,
                              const float* w45t, const float* b4, const float* b5, const float* W1, const float* b1,
                              int Z, int L, int la, EpsSource src, float* mu, float* ls, float* eps, float* z,
                              float* row_aux, float* h_d, void* hd_hi, void* hd_lo, int ld_mirror) {
  const int nchunks = (2 * Z + 7) / 8;
  const int KS = max(1, NWARPS / nchunks);
  ++*launches;
  if (rows <= 512) {
    const size_t smem = (size_t)((KS + 1) * 2 * Z + 2 * Z) * sizeof(float);
    latent_fwd_kernel<1><<<rows, NTHREADS, smem, st>>>(h_e, rows, H, w45t, b4, b5, W1, b1, Z, L, la, src, mu, ls, eps,
                                                      z, row_aux, h_d, (__nv_bfloat16*)hd_hi, (__nv_bfloat16*)hd_lo,
                                                      ld_mirror, KS);
  } else {
    constexpr int RB = 8;
    const size_t smem = (size_t)RB * ((KS + 1) * 2 * Z + 2 * Z) * sizeof(float);
    latent_fwd_kernel<RB><<<(rows + RB - 1) / RB, NTHREADS, smem, st>>>(
        h_e, rows, H, w45t, b4, b5, W1, b1, Z, L, la, src, mu, ls, eps, z, row_aux, h_d, (__nv_bfloat16*)hd_hi,
        (__nv_bfloat16*)hd_lo, ld_mirror, KS);
  }
  return cudaGetLastError();
}

cudaError_t launch_latent_bwd(cudaStream_t st, int64_t* launches, const float* da1, const float* W1,
                              const float* w45t, const float* h_e, const float* z, const float* eps, const float* mu,
                              const float* ls, int rows, int H, int Z, int L, int la, float w, float* dmu, float* dls,
                              float* da3, void* da3_hi, void* da3_lo, int ld_mirror, const float* partial,
                              int n_tiles, const float* row_aux, float* per_row, unsigned int* counter,
                              float* base_out, float mult, const float* tprior, int n_tprior, float div,
                              float* scalar_out) {
  const int nchunks = (Z + 7) / 8;
  const int NSL = max(1, (NWARPS - 2) / nchunks);
  ++*launches;
  if (rows <= 512) {
    const size_t smem = (size_t)(NSL + 2) * Z * sizeof(float);
    latent_bwd_kernel<1><<<rows, NTHREADS, smem, st>>>(
        da1, W1, w45t, h_e, z, eps, mu, ls, rows, H, Z, L, la, w, dmu, dls, da3, (__nv_bfloat16*)da3_hi,
        (__nv_bfloat16*)da3_lo, ld_mirror, partial, n_tiles, row_aux, per_row, counter, base_out, mult, tprior,
        n_tprior, div, scalar_out, NSL);
  } else {
    constexpr int RB = 8;
    const size_t smem = (size_t)RB * (NSL + 2) * Z * sizeof(float);
    latent_bwd_kernel<RB><<<(rows + RB - 1) / RB, NTHREADS, smem, st>>>(
        da1, W1, w45t, h_e, z, eps, mu, ls, rows, H, Z, L, la, w, dmu, dls, da3, (__nv_bfloat16*)da3_hi,
        (__nv_bfloat16*)da3_lo, ld_mirror, partial, n_tiles, row_aux, per_row, counter, base_out, mult, tprior,
        n_tprior, div, scalar_out, NSL);
  }
  return cudaGetLastError();
}

int small_wgrad_chunks(int rows) { return rows <= 512 ? 1 : (rows + 255) / 256; }
size_t small_wgrad_scratch_elems(int rows, int H, int Z) {
  const int c = small_wgrad_chunks(rows);
  return c > 1 ? (size_t)c * ((size_t)(Z + 1) * H + 2 * (size_t)(H + 1) * Z) : 0;
}

cudaError_t launch_small_wgrad(cudaStream_t st, int64_t* launches, const float* z, const float* da1, int R,
                               const float* h_e, const float* dmu, const float* dls, int rows, int H, int Z,
                               float* gW1, float* gb1, float* gW4, float* gb4, float* gW5, float* gb5,
                               float* scratch) {
  const int tot = (Z + 1) * H + (H + 1) * Z;
  const int chunks = small_wgrad_chunks(rows);
  const int rpc = chunks > 1 ? 256 : rows;
  dim3 grid((tot + 63) / 64, chunks);
  small_wgrad_kernel<<<grid, 256, 0, st>>>(z, da1, R, h_e, dmu, dls, rows, H, Z, rpc, gW1, gb1, gW4, gb4, gW5, gb5,
                                          scratch);
  ++*launches;
  if (chunks > 1) {
    const int all = (Z + 1) * H + 2 * (H + 1) * Z;
    small_wgrad_reduce_kernel<<<(all + 255) / 256, 256, 0, st>>>(scratch, chunks, H, Z, gW1, gb1, gW4, gb4, gW5, gb5);
    ++*launches;
  }
  return cudaGetLastError();
}
