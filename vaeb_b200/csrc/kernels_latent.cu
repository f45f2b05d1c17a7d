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

constexpr int MAX_WARPS = 16;

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

// acc[q] += sum_{k = lane, lane+32, ...} a[k] * Wm[k*ldw + c0 + q]   (q < nq <= 8), 4 k's in flight
__device__ __forceinline__ void dot8_strided(const float* __restrict__ a, int K, const float* __restrict__ Wm,
                                             int ldw, int c0, int nq, int lane, float acc[8]) {
  for (int k0 = lane; k0 < K; k0 += 128) {
    float av[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) av[u] = (k0 + 32 * u < K) ? a[k0 + 32 * u] : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int k = min(k0 + 32 * u, K - 1);
      const float* wr = Wm + (size_t)k * ldw + c0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        if (q < nq) acc[q] = fmaf(av[u], wr[q], acc[q]);
    }
  }
}

// One block per datapoint.  Warps split the 2Z head outputs in chunks of 8 (phase 1); warp 0 does the
// reparameterisation (phase 2); all threads share the decoder hidden layer (phase 3).
__global__ void __launch_bounds__(MAX_WARPS * 32)
latent_fwd_kernel(const float* __restrict__ h_e, int rows, int H, const float* __restrict__ W4,
                  const float* __restrict__ b4, const float* __restrict__ W5, const float* __restrict__ b5,
                  const float* __restrict__ W1, const float* __restrict__ b1, int Z, int L, int la, EpsSource src,
                  float* __restrict__ mu, float* __restrict__ ls, float* __restrict__ eps, float* __restrict__ z,
                  float* __restrict__ row_aux, float* __restrict__ h_d, __nv_bfloat16* __restrict__ hd_hi,
                  __nv_bfloat16* __restrict__ hd_lo, int ld_mirror) {
  extern __shared__ float sm[];
  float* out = sm;            // [2Z] mu | ls
  float* zs = sm + 2 * Z;     // [Z] z of the current sample
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int m = blockIdx.x;
  const float* hr = h_e + (size_t)m * H;
  const int cpm = (Z + 7) / 8;                 // chunks per head matrix
  for (int c = warp; c < 2 * cpm; c += nwarps) {
    const int mat = c / cpm, c0 = (c % cpm) * 8, nq = min(8, Z - c0);
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    dot8_strided(hr, H, mat ? W5 : W4, Z, c0, nq, lane, acc);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float t = warp_sum(acc[q]);
      if (lane == 0 && q < nq) out[mat * Z + c0 + q] = t + (mat ? b5[c0 + q] : b4[c0 + q]);
    }
  }
  __syncthreads();
  float term = 0.f, la_acc = 0.f;
  if (warp == 0) {
    for (int j = lane; j < Z; j += 32) {
      const float am = out[j], al = out[Z + j];
      mu[(size_t)m * Z + j] = am;
      ls[(size_t)m * Z + j] = al;
      if (!la) term += 0.5f * (1.0f + al - am * am - expf(al));
    }
  }
  for (int l = 0; l < L; ++l) {
    const size_t r = (size_t)l * rows + m;
    if (warp == 0) {
      for (int j = lane; j < Z; j += 32) {
        const float am = out[j], al = out[Z + j];
        const size_t o2 = r * Z + j;
        const float e = src.injected
                            ? src.injected[o2]
                            : philox_normal1(src.seed, src.stream, src.step, (uint32_t)l,
                                             (uint64_t)((src.row_offset + m) * Z + j));
        const float zv = am + expf(0.5f * al) * e;
        eps[o2] = e;
        z[o2] = zv;
        zs[j] = zv;
        la_acc += -0.5f * zv * zv + 0.5f * al + 0.5f * e * e;
      }
    }
    __syncthreads();
    for (int n = threadIdx.x; n < H; n += blockDim.x) {
      float a = b1[n];
      for (int j = 0; j < Z; ++j) a = fmaf(zs[j], W1[(size_t)j * H + n], a);
      const float hv = tanhf(a);
      h_d[r * H + n] = hv;
      if (hd_hi) store_split(hd_hi, hd_lo, r * ld_mirror + n, hv);
    }
    __syncthreads();
  }
  if (warp == 0) {
    if (la && L > 0) term = la_acc / (float)L;
    term = warp_sum(term);
    if (lane == 0) row_aux[m] = term;
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

// One block per datapoint: dz by warps (chunks of 8 latent dims), dmu/dls by warp 0, da3 and the
// bound by all threads; the last block to finish totals the per-row bounds in a fixed order.
__global__ void __launch_bounds__(MAX_WARPS * 32)
latent_bwd_kernel(const float* __restrict__ da1, const float* __restrict__ W1, const float* __restrict__ W4,
                  const float* __restrict__ W5, const float* __restrict__ h_e, const float* __restrict__ z,
                  const float* __restrict__ eps, const float* __restrict__ mu, const float* __restrict__ ls,
                  int rows, int H, int Z, int L, int la, float w, float* __restrict__ dmu, float* __restrict__ dls,
                  float* __restrict__ da3, __nv_bfloat16* __restrict__ da3_hi, __nv_bfloat16* __restrict__ da3_lo,
                  int ld_mirror, const float* __restrict__ partial, int n_tiles, const float* __restrict__ row_aux,
                  float* __restrict__ per_row, unsigned int* __restrict__ counter, float* __restrict__ base_out,
                  float mult, const float* __restrict__ tprior, int n_tprior, float div,
                  float* __restrict__ scalar_out) {
  extern __shared__ float sm[];
  __shared__ float red[MAX_WARPS];
  __shared__ int is_last;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  float* dm = sm;          // [Z] dmu
  float* dl = dm + Z;      // [Z] dls
  float* dzs = dl + Z;     // [Z] dz of the current sample
  const int m = blockIdx.x;
  const float s = w / (float)L;
  for (int j = threadIdx.x; j < Z; j += blockDim.x) { dm[j] = 0.f; dl[j] = 0.f; }
  __syncthreads();
  for (int l = 0; l < L; ++l) {
    const size_t r = (size_t)l * rows + m;
    const float* dr = da1 + r * H;
    // dz[j] = sum_n da1[r,n] W1[j,n]: W1 rows are contiguous in n, 4 n's in flight per lane
    for (int c0 = warp * 8; c0 < Z; c0 += nwarps * 8) {
      const int nq = min(8, Z - c0);
      float acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.f;
      for (int n0 = lane; n0 < H; n0 += 128) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int n = n0 + 32 * u;
          if (n < H) {
            const float dv = dr[n];
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q < nq) acc[q] = fmaf(dv, W1[(size_t)(c0 + q) * H + n], acc[q]);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const float t = warp_sum(acc[q]);
        if (lane == 0 && q < nq) dzs[c0 + q] = t;
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < Z; j += blockDim.x) {
      const size_t o2 = r * Z + j;
      float d = dzs[j];
      if (la) d -= s * z[o2];
      const float sd = expf(0.5f * ls[(size_t)m * Z + j]);
      dm[j] += d;
      dl[j] += d * (0.5f * sd * eps[o2]);
    }
    __syncthreads();
  }
  for (int j = threadIdx.x; j < Z; j += blockDim.x) {
    const float lsv = ls[(size_t)m * Z + j], muv = mu[(size_t)m * Z + j];
    float a = dm[j], b = dl[j];
    if (la) {
      b += w * 0.5f;
    } else {
      a -= w * muv;
      b += w * 0.5f * (1.0f - expf(lsv));
    }
    dm[j] = a;
    dl[j] = b;
    dmu[(size_t)m * Z + j] = a;
    dls[(size_t)m * Z + j] = b;
  }
  __syncthreads();
  // da3[m,n] = (sum_j dmu_j W4[n,j] + dls_j W5[n,j]) * (1 - h_e^2)
  for (int n = threadIdx.x; n < H; n += blockDim.x) {
    float a = 0.f;
    const float* w4 = W4 + (size_t)n * Z;
    const float* w5 = W5 + (size_t)n * Z;
    for (int j = 0; j < Z; ++j) a = fmaf(dm[j], w4[j], fmaf(dl[j], w5[j], a));
    const float hv = h_e[(size_t)m * H + n];
    const float v = a * (1.0f - hv * hv);
    da3[(size_t)m * H + n] = v;
    if (da3_hi) store_split(da3_hi, da3_lo, (size_t)m * ld_mirror + n, v);
  }
  // per-datapoint bound: (1/L) sum_l sum_tiles partial + row_aux   (last warp)
  if (warp == nwarps - 1) {
    float t = 0.f;
    for (int l = 0; l < L; ++l) {
      const float* p = partial + ((size_t)l * rows + m) * n_tiles;
      for (int q = lane; q < n_tiles; q += 32) t += p[q];
    }
    t = warp_sum(t);
    if (lane == 0) per_row[m] = t / (float)L + row_aux[m];
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

cudaError_t launch_latent_fwd(cudaStream_t st, int64_t* launches, const float* h_e, int rows, int H, const float* W4,
                              const float* b4, const float* W5, const float* b5, const float* W1, const float* b1,
                              int Z, int L, int la, EpsSource src, float* mu, float* ls, float* eps, float* z,
                              float* row_aux, float* h_d, void* hd_hi, void* hd_lo, int ld_mirror) {
  const size_t smem = (size_t)3 * Z * sizeof(float);
  const int nw = max(4, min(MAX_WARPS, 2 * ((Z + 7) / 8)));
  latent_fwd_kernel<<<rows, nw * 32, smem, st>>>(
      h_e, rows, H, W4, b4, W5, b5, W1, b1, Z, L, la, src, mu, ls, eps, z, row_aux, h_d, (__nv_bfloat16*)hd_hi,
      (__nv_bfloat16*)hd_lo, ld_mirror);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t launch_latent_bwd(cudaStream_t st, int64_t* launches, const float* da1, const float* W1, const float* W4,
                              const float* W5, const float* h_e, const float* z, const float* eps, const float* mu,
                              const float* ls, int rows, int H, int Z, int L, int la, float w, float* dmu, float* dls,
                              float* da3, void* da3_hi, void* da3_lo, int ld_mirror, const float* partial,
                              int n_tiles, const float* row_aux, float* per_row, unsigned int* counter,
                              float* base_out, float mult, const float* tprior, int n_tprior, float div,
                              float* scalar_out) {
  const size_t smem = (size_t)3 * Z * sizeof(float);
  const int nw = max(4, min(MAX_WARPS, (Z + 7) / 8 + 1));
  latent_bwd_kernel<<<rows, nw * 32, smem, st>>>(
      da1, W1, W4, W5, h_e, z, eps, mu, ls, rows, H, Z, L, la, w, dmu, dls, da3, (__nv_bfloat16*)da3_hi,
      (__nv_bfloat16*)da3_lo, ld_mirror, partial, n_tiles, row_aux, per_row, counter, base_out, mult, tprior,
      n_tprior, div, scalar_out);
  ++*launches;
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
