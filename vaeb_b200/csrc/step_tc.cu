// The tensor-core single-launch AEVB step (see step_tc.cuh): VAEB.py:408-415 at M <= 128 rows, Bernoulli decoder, L = 1.
//
// Work decomposition.  132 CTAs = 33 clusters of 4 (one CTA per SM).  A phase is a list of items:
//   * cluster items -- the skinny activation GEMMs out[128 x N] = act[128 x K] . W[K x N]: one N tile per cluster, the
//     contraction split four ways over the cluster's CTAs (k chunks of 64), partial accumulators read from TMEM and
//     reduce-scattered by rows through distributed shared memory (st.shared::cluster), each CTA finishing 32 rows with
//     the layer's fused epilogue (tanh / reparameterisation + KL / Bernoulli log-likelihood + delta / tanh' / dz);
//   * CTA items -- the weight-gradient GEMMs gW[128 features x N] = act^T . delta (K = the minibatch, zero padded to
//     128): one output tile per CTA, Adagrad with the prior (VAEB.py:389-390,426-444) in the epilogue, which also
//     rewrites the bf16 mirrors of the weights it just updated.
// Operands.  The A side of every GEMM is produced in software: fp32 activations are read from L2 (ld.global.cg),
// split into bf16 hi + lo and stored into shared memory in the UMMA K-major SWIZZLE_128B image (the decoder hidden
// layer and the encoder's dh_e are not even read -- they are recomputed there from z resp. [dmu|dls] with FFMA, K = Z).
// The B side of the activation GEMMs is a straight 16-byte copy of a pre-swizzled mirror blob.  One thread issues
// hi*hi + hi*lo + lo*hi per k step (bf16x3: the fp32 parity tier), commits to an mbarrier; 16 warps read TMEM.
// Phases of one update (a grid barrier after each):
//   P1 enc1 (+ the W4/W5 update of the previous step on the spare cluster)   P2 heads + reparam + KL
//   P3 dec1 (recomputed) + dec2 + log-lik + da2                              P4 dgrad -> da1
//   P5 dz -> dmu, dls | W2,b2 update | the bound                             P6 W3,b3 update (dh_e recomputed) | W1,b1 update
#include <cuda_bf16.h>

#include <algorithm>

#include "common.cuh"
#include "philox.cuh"
#include "step_tc.cuh"
#include "tc_common.cuh"

namespace st2 {

constexpr int TBA = MP * 128;                      // bytes of one 64-wide k chunk of an A tile (hi or lo half)
constexpr int A_MAXCH = 4;
constexpr int SM_A = 0;                            // A_MAXCH chunks x (hi, lo) x 16 KB = 128 KB
constexpr int SM_B = A_MAXCH * 2 * TBA;            // B tiles: up to 32 KB (48 KB for the heads), scratch above
constexpr int SM_B_BYTES = 57344;
constexpr int SM_SCR = SM_B + 32768;               // 24 KB of producer scratch inside the B region
constexpr int SM_RECV = SM_B + SM_B_BYTES;         // [CL][32 rows][<= 48 cols] fp32 partials from the cluster
constexpr int SM_RECV_BYTES = CL * 32 * 48 * 4;
constexpr int SM_MISC = SM_RECV + SM_RECV_BYTES;   // mbarrier, TMEM slot, small reductions
constexpr int SMEM_BYTES = SM_MISC + 1024 + 1024;  // + slack for the 1024-byte alignment of the base
constexpr uint32_t TMEM_COLS = 64;

struct Ctx {
  uint8_t* sm;
  uint32_t tmem;
  uint64_t* mma_bar;
  uint32_t mma_phase;
  int rank, cid, ncl, warp, lane;
  long long* trace; int tn;            // optional sub-phase stamps (cluster 0, rank 0, thread 0 of the last step)
};

// ---- PTX helpers -------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_idx() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ long long gtime() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define ST2_TRACE(c, id) do { if ((c).trace && threadIdx.x == 0 && (c).tn < 60) { (c).trace[(c).tn * 2] = (id); (c).trace[(c).tn * 2 + 1] = gtime(); ++(c).tn; } } while (0)
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
// eight fp32 -> eight bf16 hi + eight bf16 lo (lo = the rounding residual: hi + lo carries 16 mantissa bits)
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    h[q] = pack2(v[2 * q], v[2 * q + 1]);
    l[q] = pack2(v[2 * q] - __uint_as_float(h[q] << 16), v[2 * q + 1] - __uint_as_float(h[q] & 0xffff0000u));
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// branch-free tanh, relative error < 5e-6 (as the bf16x3 layer kernels, tc_layers.cu): 1 - 2/(e^{2|x|}+1) from
// ex2.approx / rcp.approx for |x| >= 0.1, the odd Taylor polynomial through x^9 below
__device__ __forceinline__ float tanh_fast(float x) {
  const float ax = fabsf(x), x2 = x * x;
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(ax * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  const float big = copysignf(fmaf(-2.0f, r, 1.0f), x);
  const float poly = fmaf(x * x2, fmaf(x2, fmaf(x2, fmaf(x2, 0.021869488536155203f, -0.053968253968253971f),
                                                0.13333333333333333f), -0.33333333333333331f), x);
  return ax < 0.1f ? poly : big;
}
__device__ __forceinline__ float softplusf_(float a) { return fmaxf(a, 0.f) + log1pf(expf(-fabsf(a))); }
__device__ __forceinline__ float sigmoidf_(float a) { return 1.0f / (1.0f + expf(-a)); }

// byte offset of the 16-byte unit (row r, k octet cu) inside one K-major SWIZZLE_128B chunk tile
__device__ __forceinline__ uint32_t unit_off(int r, int cu) { return (uint32_t)r * 128u + (((uint32_t)cu ^ ((uint32_t)r & 7u)) << 4); }

// ---- operand producers -------------------------------------------------------------------------------------------
// Every producer is two-phase: ALL global loads of a batch are issued before the first shared-memory store, so a
// thread pays one L2 round trip per batch instead of one per 32 bytes (stores through generic pointers would otherwise
// order the loads behind them).  A unit = eight consecutive k of one tile row = one 16-byte piece of the hi and of
// the lo half of the tile.
__device__ __forceinline__ void put_unit(uint8_t* d, int half_bytes, const float* v) {
  uint4 hi, lo;
  split8(v, hi, lo);
  *reinterpret_cast<uint4*>(d) = hi;
  *reinterpret_cast<uint4*>(d + half_bytes) = lo;
}

// K-major tile whose rows are the SOURCE rows (activation GEMMs): src[rows][ld] fp32, k in [k0, k0 + 64 nch) clipped to
// kmax (a multiple of 8), rows >= rows_valid and k >= kmax zero filled.  Chunk ci: hi at base + ci*2*TBA, lo TBA later.
struct RowsSrc {
  const float* src; int ld, rows_valid, k0, kmax, per_row;
  __device__ __forceinline__ void load(int u, float* v) const {
    const int r = u / per_row, ku = u - r * per_row;
    const int k = k0 + (ku >> 3) * 64 + (ku & 7) * 8;
    if (r < rows_valid && k < kmax) {
      const float* s = src + (size_t)r * ld + k;
      const float4 a = __ldcg(reinterpret_cast<const float4*>(s)), b = __ldcg(reinterpret_cast<const float4*>(s + 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
  }
  __device__ __forceinline__ uint8_t* dst(uint8_t* base, int u) const {
    const int r = u / per_row, ku = u - r * per_row;
    return base + (size_t)(ku >> 3) * 2 * TBA + unit_off(r, ku & 7);
  }
};

// K-major tile whose rows are FEATURES and whose contraction index is the minibatch (weight gradients): element
// (feature f0 + i, batch b) = src[b][f0 + i]; feature == ones_f reads as 1 (bias gradient), features >= fmax and
// batch rows >= M are zero.  Two k chunks (128 batch rows).  TR rows per tile, TR * 128 bytes per chunk half.
struct TSrc {
  const float* src; int TR, ld, M, f0, fmax, ones_f;
  __device__ __forceinline__ void load(int u, float* v) const {
    const int bo = u / TR, f = f0 + (u - bo * TR);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int b = bo * 8 + e;
      float t = 0.f;
      if (b < M) {
        if (f == ones_f) t = 1.0f;
        else if (f < fmax) t = __ldcg(src + (size_t)b * ld + f);
      }
      v[e] = t;
    }
  }
  __device__ __forceinline__ uint8_t* dst(uint8_t* base, int u) const {
    const int bo = u / TR, i = u - bo * TR;
    return base + (size_t)(bo >> 3) * 2 * (TR * 128) + unit_off(i, bo & 7);
  }
};

// A tile from RowsSrc (2 nch units per thread) + the pre-swizzled B blob of a mirror (16-byte pieces)
__device__ __forceinline__ void stage_rows_blob(uint8_t* sm, const RowsSrc& a, int nch, const uint8_t* __restrict__ blob,
                                                int blob_bytes) {
  uint8_t* bd = sm + SM_B;
  uint4 bl[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = (threadIdx.x + i * NT) * 16;
    bl[i] = o < blob_bytes ? __ldcg(reinterpret_cast<const uint4*>(blob + o)) : make_uint4(0, 0, 0, 0);
  }
  const int n_units = MP * nch * 8;
  bool blob_stored = false;
  for (int u0 = threadIdx.x; u0 < n_units; u0 += 4 * NT) {
    float v[4][8];
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (u0 + b * NT < n_units) a.load(u0 + b * NT, v[b]);
    if (!blob_stored) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int o = (threadIdx.x + i * NT) * 16;
        if (o < blob_bytes) *reinterpret_cast<uint4*>(bd + o) = bl[i];
      }
      blob_stored = true;
    }
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (u0 + b * NT < n_units) put_unit(a.dst(sm + SM_A, u0 + b * NT), TBA, v[b]);
  }
  if (!blob_stored) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = (threadIdx.x + i * NT) * 16;
      if (o < blob_bytes) *reinterpret_cast<uint4*>(bd + o) = bl[i];
    }
  }
  for (int o = (threadIdx.x + 4 * NT) * 16; o < blob_bytes; o += NT * 16)
    *reinterpret_cast<uint4*>(bd + o) = __ldcg(reinterpret_cast<const uint4*>(blob + o));
}

// transposed tiles of a weight-gradient GEMM: A (128 features: 4 units per thread) and optionally B (<= 48 features)
__device__ __forceinline__ void stage_T(uint8_t* sm, const TSrc& a, const TSrc* b) {
  float va[4][8], vb[2][8];
  const int nb = b ? b->TR * 16 : 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) a.load(threadIdx.x + i * NT, va[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if ((int)threadIdx.x + i * NT < nb) b->load(threadIdx.x + i * NT, vb[i]);
#pragma unroll
  for (int i = 0; i < 4; ++i) put_unit(a.dst(sm + SM_A, threadIdx.x + i * NT), TBA, va[i]);
#pragma unroll
  for (int i = 0; i < 2; ++i)
    if ((int)threadIdx.x + i * NT < nb) put_unit(b->dst(sm + SM_B, threadIdx.x + i * NT), b->TR * 128, vb[i]);
}

__device__ __forceinline__ void copy_blob(uint8_t* dst, const uint8_t* __restrict__ src, int bytes) {
  for (int o = threadIdx.x * 16; o < bytes; o += NT * 16)
    *reinterpret_cast<uint4*>(dst + o) = __ldcg(reinterpret_cast<const uint4*>(src + o));
}

// ---- MMA ---------------------------------------------------------------------------------------------------------
// acc[128 x N] = sum over nch chunks / k16_total k steps of A.B^T in bf16x3.  Called by every thread of the CTA after
// the producers; returns when the accumulator is complete (nch == 0: nothing is issued).
__device__ __forceinline__ void mma_run(Ctx& c, int nch, int k16_total, int TBB, int N) {
  tc::fence_proxy_async();            // this thread's shared-memory stores -> visible to the tensor core (async proxy)
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (nch <= 0) return;
  if (threadIdx.x == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(MP, N, 0, 0);
    const uint32_t a0 = tc::smem_u32(c.sm + SM_A), b0 = tc::smem_u32(c.sm + SM_B);
    uint32_t acc = 0;
    int k16 = 0;
    for (int ci = 0; ci < nch; ++ci) {
      const uint32_t a = a0 + (uint32_t)ci * 2u * TBA, b = b0 + (uint32_t)ci * 2u * (uint32_t)TBB;
#pragma unroll
      for (int k = 0; k < 4; ++k, ++k16) {
        if (k16 >= k16_total) break;
        const uint64_t ah = tc::desc_kmajor(a, k), al = tc::desc_kmajor(a + TBA, k);
        const uint64_t bh = tc::desc_kmajor(b, k), bl = tc::desc_kmajor(b + (uint32_t)TBB, k);
        tc::umma_bf16(c.tmem, ah, bh, idesc, acc);
        tc::umma_bf16(c.tmem, ah, bl, idesc, 1u);
        tc::umma_bf16(c.tmem, al, bh, idesc, 1u);
        acc = 1u;
      }
    }
    tc::umma_commit(c.mma_bar);
  }
  tc::mbar_wait(c.mma_bar, c.mma_phase);
  c.mma_phase ^= 1u;
  tc::tc_fence_after();
}

// Partial accumulator [128 x N] -> the four CTAs of the cluster by row quarter: rows 32q..32q+31 go to CTA q, slot
// `rank` of its receive buffer.  `have` == false sends zeros (this CTA had no k chunk).  Then a cluster barrier.
__device__ __forceinline__ void reduce_scatter(Ctx& c, int N, bool have) {
  const int q = c.warp & 3;
  const uint32_t recv_local = tc::smem_u32(c.sm + SM_RECV);
  const uint32_t dst0 = mapa(recv_local, (uint32_t)q);
  for (int u = c.warp >> 2; u < N / 8; u += 4) {
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (have) {
      tmem_ld8(c.tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * 8), v);
      tc::tmem_ld_wait();
    }
    const uint32_t d = dst0 + (uint32_t)(((c.rank * 32 + c.lane) * N + u * 8) * 4);
    st_cluster_v4(d, v[0], v[1], v[2], v[3]);
    st_cluster_v4(d + 16, v[4], v[5], v[6], v[7]);
  }
  tc::tc_fence_before();
  cluster_sync();
}
__device__ __forceinline__ float recv_sum(const float* recv, int N, int row, int col) {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CL; ++k) s += recv[(k * 32 + row) * N + col];
  return s;
}

struct Hyper { float lr, eps, prior, p2; };
__device__ __forceinline__ float adagrad_step(float* P, float* ada, size_t o, float g, const Hyper& hy) {
  const float p = __ldcg(P + o), a0 = __ldcg(ada + o);
  g -= hy.prior * p;                                    // VAEB.py:389-390
  const float a = a0 + g * g;                           // VAEB.py:439
  float np_ = p + hy.lr * g / (sqrtf(a) + hy.eps);      // VAEB.py:441
  if (hy.p2 != 0.f) np_ -= hy.p2 * p * p;               // VAEBfullbayes.py:183-184
  P[o] = np_;
  ada[o] = a;
  return np_;
}
// Adagrad on the (up to) eight parameters one epilogue thread owns: every load is issued before the first store
// (one L2 round trip per unit), 16-byte accesses when the eight are contiguous and aligned.  nv = the new values.
__device__ __forceinline__ void adagrad8(float* P, float* ada, const size_t* off, const bool* ok, const float* g,
                                         const Hyper& hy, float* nv) {
  bool vec = (off[0] & 3) == 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) vec = vec && ok[e] && off[e] == off[0] + e;
  float p[8], a[8];
  if (vec) {
    const float4 p0 = __ldcg(reinterpret_cast<const float4*>(P + off[0])), p1 = __ldcg(reinterpret_cast<const float4*>(P + off[0] + 4));
    const float4 a0 = __ldcg(reinterpret_cast<const float4*>(ada + off[0])), a1 = __ldcg(reinterpret_cast<const float4*>(ada + off[0] + 4));
    p[0] = p0.x; p[1] = p0.y; p[2] = p0.z; p[3] = p0.w; p[4] = p1.x; p[5] = p1.y; p[6] = p1.z; p[7] = p1.w;
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      p[e] = ok[e] ? __ldcg(P + off[e]) : 0.f;
      a[e] = ok[e] ? __ldcg(ada + off[e]) : 0.f;
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float gg = g[e] - hy.prior * p[e];             // VAEB.py:389-390
    a[e] += gg * gg;                                      // VAEB.py:439
    float q = p[e] + hy.lr * gg / (sqrtf(a[e]) + hy.eps); // VAEB.py:441
    if (hy.p2 != 0.f) q -= hy.p2 * p[e] * p[e];           // VAEBfullbayes.py:183-184
    nv[e] = q;
  }
  if (vec) {
    *reinterpret_cast<float4*>(P + off[0]) = make_float4(nv[0], nv[1], nv[2], nv[3]);
    *reinterpret_cast<float4*>(P + off[0] + 4) = make_float4(nv[4], nv[5], nv[6], nv[7]);
    *reinterpret_cast<float4*>(ada + off[0]) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(ada + off[0] + 4) = make_float4(a[4], a[5], a[6], a[7]);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (ok[e]) { P[off[e]] = nv[e]; ada[off[e]] = a[e]; }
  }
}
// one element of a mirror: tile of TR rows, KC chunks per tile; (n, k) -> hi and lo halves
__device__ __forceinline__ void mirror_put(uint8_t* m, int TR, int KC, int n, int k, float v) {
  const int TB = TR * 128;
  const int r = n % TR;
  uint8_t* d = m + ((size_t)(n / TR) * KC + (k >> 6)) * 2 * TB + tc::sw128_offset(r, k & 63);
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(d) = h;
  *reinterpret_cast<__nv_bfloat16*>(d + TB) = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ---- grid barrier (monotonic counter, release / acquire at gpu scope) --------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned long long* bar, unsigned long long target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(bar) : "memory");
    unsigned long long v;
    uint32_t spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(bar) : "memory");
      if (++spins > 20000000u) __trap();     // a lost CTA must abort the launch, not hang the GPU
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// =================================================================================================================
// cluster items
// =================================================================================================================
// P1: h_e[:, 16t .. 16t+15] = tanh(x.W3 + b3)                                               VAEB.py:246
__device__ __noinline__ void item_enc1(Ctx& c, const Params& p, const float* x, float* he, int t) {
  const int c0 = c.rank * p.KD / CL, c1 = (c.rank + 1) * p.KD / CL, nch = c1 - c0;
  constexpr int TB = TR_ENC1 * 128;
  stage_rows_blob(c.sm, RowsSrc{x, p.D, p.M, c0 * 64, p.D, nch * 8}, nch,
                  p.m_enc1 + ((size_t)t * p.KD + c0) * 2 * TB, nch * 2 * TB);
  mma_run(c, nch, min(nch * 4, (p.D - c0 * 64 + 15) / 16), TB, TR_ENC1);
  reduce_scatter(c, TR_ENC1, nch > 0);
  const float* recv = reinterpret_cast<const float*>(c.sm + SM_RECV);
  {
    const int row = threadIdx.x >> 4, col = threadIdx.x & 15;
    const int gr = c.rank * 32 + row, j = t * 16 + col;
    if (gr < p.M) {
      const float v = recv_sum(recv, TR_ENC1, row, col);
      he[(size_t)gr * p.HP + j] = j < p.H ? tanhf(v + __ldcg(p.P + p.ob3 + j)) : 0.f;
    }
  }
  cluster_sync();
}

// P2: (mu, ls) = h_e.[W4|W5] + b, eps, z = mu + exp(ls/2) eps, the KL / L^A row term        VAEB.py:248-249,41-47,343,322-325
__device__ __noinline__ void item_heads(Ctx& c, const Params& p, const float* he, uint32_t step) {
  ST2_TRACE(c, 20);
  const int c0 = c.rank * p.KH / CL, c1 = (c.rank + 1) * p.KH / CL, nch = c1 - c0;
  const int TB = p.NH * 128, Z = p.Z, N = p.NH;
  stage_rows_blob(c.sm, RowsSrc{he, p.HP, p.M, c0 * 64, p.HP, nch * 8}, nch, p.m_heads + (size_t)c0 * 2 * TB, nch * 2 * TB);
  ST2_TRACE(c, 21);
  mma_run(c, nch, nch * 4, TB, N);
  ST2_TRACE(c, 22);
  reduce_scatter(c, N, nch > 0);
  ST2_TRACE(c, 23);
  const float* recv = reinterpret_cast<const float*>(c.sm + SM_RECV);
  float* term = reinterpret_cast<float*>(c.sm + SM_B);          // [32][Z] (the B tiles are dead)
  for (int it = threadIdx.x; it < 32 * Z; it += NT) {
    const int row = it / Z, j = it - row * Z;
    const int gr = c.rank * 32 + row;
    float tv = 0.f;
    if (gr < p.M) {
      const float am = recv_sum(recv, N, row, j) + __ldcg(p.P + p.ob4 + j);
      const float al = recv_sum(recv, N, row, Z + j) + __ldcg(p.P + p.ob5 + j);
      const size_t o = (size_t)gr * Z + j;
      const float e = p.eps_inj ? __ldcg(p.eps_inj + o)
                                : philox_normal1(p.seed, VAEB_STREAM_TRAIN, step, 0u, (uint64_t)((p.row_offset + gr) * Z + j));
      const float zv = am + expf(0.5f * al) * e;
      p.mu[o] = am; p.ls[o] = al; p.eps[o] = e; p.z[o] = zv;
      tv = p.la ? (-0.5f * zv * zv + 0.5f * al + 0.5f * e * e) : 0.5f * (1.0f + al - am * am - expf(al));
    }
    term[it] = tv;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int gr = c.rank * 32 + threadIdx.x;
    if (gr < p.M) {
      float s = 0.f;
      for (int j = 0; j < Z; ++j) s += term[threadIdx.x * Z + j];
      p.aux[gr] = s;
    }
  }
  ST2_TRACE(c, 24);
  cluster_sync();
  ST2_TRACE(c, 25);
}

// P3: h_d = tanh(z.W1 + b1) recomputed into the A tile, a = h_d.W2 + b2, x a - softplus(a), da2 = w (x - sigmoid a)
//                                                                                           VAEB.py:254,263,311
__device__ __noinline__ void item_dec2(Ctx& c, const Params& p, const float* x, int t) {
  const int c0 = c.rank * p.KH / CL, c1 = (c.rank + 1) * p.KH / CL, nch = c1 - c0;
  constexpr int TB = TR_DEC2 * 128;
  const int Z = p.Z, H = p.H, M = p.M;
  float* zs = reinterpret_cast<float*>(c.sm + SM_SCR);          // [MP][Z]
  float* ws = zs + MP * Z;                                      // [Z][64]
  float* bs = ws + Z * 64;                                      // [64]
  {
    // z and the mirror blob: all loads first (one L2 round trip), then the shared-memory stores
    float zr[6]; uint4 bl[4];
    const uint8_t* blob = p.m_dec2 + ((size_t)t * p.KH + c0) * 2 * TB;
    const int bbytes = nch * 2 * TB;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      const int e = threadIdx.x + i * NT;
      zr[i] = e < M * Z ? __ldcg(p.z + e) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = (threadIdx.x + i * NT) * 16;
      bl[i] = o < bbytes ? __ldcg(reinterpret_cast<const uint4*>(blob + o)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i)
      if ((int)threadIdx.x + i * NT < MP * Z) zs[threadIdx.x + i * NT] = zr[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int o = (threadIdx.x + i * NT) * 16;
      if (o < bbytes) *reinterpret_cast<uint4*>(c.sm + SM_B + o) = bl[i];
    }
  }
  for (int ci = 0; ci < nch; ++ci) {
    const int k0 = (c0 + ci) * 64;
    __syncthreads();
    {
      float wr[3]; float br = 0.f;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const int e = threadIdx.x + i * NT;
        const int q = e >> 6, kk = e & 63;
        wr[i] = (e < Z * 64 && k0 + kk < H) ? __ldcg(p.P + p.oW1 + (size_t)q * H + k0 + kk) : 0.f;
      }
      if (threadIdx.x < 64 && k0 + threadIdx.x < H) br = __ldcg(p.P + p.ob1 + k0 + threadIdx.x);
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if ((int)threadIdx.x + i * NT < Z * 64) ws[threadIdx.x + i * NT] = wr[i];
      if (threadIdx.x < 64) bs[threadIdx.x] = br;
    }
    __syncthreads();
    for (int u = threadIdx.x; u < M * 8; u += NT) {
      const int r = u >> 3, cu = u & 7;
      float v[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = bs[cu * 8 + e];
      for (int q = 0; q < Z; ++q) {
        const float zq = zs[r * Z + q];
        const float4 w0 = *reinterpret_cast<const float4*>(ws + q * 64 + cu * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(ws + q * 64 + cu * 8 + 4);
        v[0] = fmaf(zq, w0.x, v[0]); v[1] = fmaf(zq, w0.y, v[1]); v[2] = fmaf(zq, w0.z, v[2]); v[3] = fmaf(zq, w0.w, v[3]);
        v[4] = fmaf(zq, w1.x, v[4]); v[5] = fmaf(zq, w1.y, v[5]); v[6] = fmaf(zq, w1.z, v[6]); v[7] = fmaf(zq, w1.w, v[7]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (k0 + cu * 8 + e < H) ? tanh_fast(v[e]) : 0.f;
      if (t == 0) {                                              // the cluster of tile 0 publishes h_d for P4 / P5
        float* o = p.hd + (size_t)r * p.HP + k0 + cu * 8;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
      uint4 hi, lo;
      split8(v, hi, lo);
      uint8_t* d = c.sm + SM_A + (size_t)ci * 2 * TBA + unit_off(r, cu);
      *reinterpret_cast<uint4*>(d) = hi;
      *reinterpret_cast<uint4*>(d + TBA) = lo;
    }
  }
  for (int u = M * 8 * nch + threadIdx.x; u < MP * 8 * nch; u += NT) {      // rows >= M of the A tile: zeros
    const int r = M + (u - M * 8 * nch) / (8 * nch), ku = (u - M * 8 * nch) % (8 * nch);
    uint8_t* d = c.sm + SM_A + (size_t)(ku >> 3) * 2 * TBA + unit_off(r, ku & 7);
    *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(d + TBA) = make_uint4(0, 0, 0, 0);
  }
  mma_run(c, nch, nch * 4, TB, TR_DEC2);
  reduce_scatter(c, TR_DEC2, nch > 0);
  const float* recv = reinterpret_cast<const float*>(c.sm + SM_RECV);
  float* term = reinterpret_cast<float*>(c.sm + SM_B);          // [32][32]
  for (int it = threadIdx.x; it < 32 * TR_DEC2; it += NT) {
    const int row = it >> 5, col = it & 31;
    const int gr = c.rank * 32 + row, n = t * TR_DEC2 + col;
    float tv = 0.f;
    if (gr < M && n < p.D) {
      const float a = recv_sum(recv, TR_DEC2, row, col) + __ldcg(p.P + p.ob2 + n);
      const float xv = __ldcg(x + (size_t)gr * p.D + n);
      p.da2[(size_t)gr * p.D + n] = p.w * (xv - sigmoidf_(a));
      tv = xv * a - softplusf_(a);
    }
    term[it] = tv;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int gr = c.rank * 32 + threadIdx.x;
    if (gr < M) {
      float s = 0.f;
      for (int j = 0; j < TR_DEC2; ++j) s += term[threadIdx.x * TR_DEC2 + j];
      p.partial[(size_t)gr * p.n_tiles3 + t] = s;
    }
  }
  cluster_sync();
}

// P4: da1[:, 16t..] = (da2.W2^T) * (1 - h_d^2)                                              T.grad, VAEB.py:397
__device__ __noinline__ void item_dgrad(Ctx& c, const Params& p, int t) {
  ST2_TRACE(c, 40);
  const int c0 = c.rank * p.KD / CL, c1 = (c.rank + 1) * p.KD / CL, nch = c1 - c0;
  constexpr int TB = TR_DGRAD * 128;
  stage_rows_blob(c.sm, RowsSrc{p.da2, p.D, p.M, c0 * 64, p.D, nch * 8}, nch,
                  p.m_dgrad + ((size_t)t * p.KD + c0) * 2 * TB, nch * 2 * TB);
  ST2_TRACE(c, 41);
  mma_run(c, nch, min(nch * 4, (p.D - c0 * 64 + 15) / 16), TB, TR_DGRAD);
  ST2_TRACE(c, 42);
  reduce_scatter(c, TR_DGRAD, nch > 0);
  ST2_TRACE(c, 43);
  const float* recv = reinterpret_cast<const float*>(c.sm + SM_RECV);
  {
    const int row = threadIdx.x >> 4, col = threadIdx.x & 15;
    const int gr = c.rank * 32 + row, j = t * 16 + col;
    if (gr < p.M) {
      const size_t o = (size_t)gr * p.HP + j;
      const float hv = __ldcg(p.hd + o);
      p.da1[o] = recv_sum(recv, TR_DGRAD, row, col) * (1.0f - hv * hv);
    }
  }
  ST2_TRACE(c, 44);
  cluster_sync();
  ST2_TRACE(c, 45);
}

// P5: dz = da1.W1^T -> dmu, dls (SURVEY.md 8a backward formulas)
__device__ __noinline__ void item_dz(Ctx& c, const Params& p) {
  const int c0 = c.rank * p.KH / CL, c1 = (c.rank + 1) * p.KH / CL, nch = c1 - c0;
  const int TB = p.NZ * 128, Z = p.Z, N = p.NZ;
  stage_rows_blob(c.sm, RowsSrc{p.da1, p.HP, p.M, c0 * 64, p.HP, nch * 8}, nch, p.m_dz + (size_t)c0 * 2 * TB, nch * 2 * TB);
  mma_run(c, nch, nch * 4, TB, N);
  reduce_scatter(c, N, nch > 0);
  const float* recv = reinterpret_cast<const float*>(c.sm + SM_RECV);
  for (int it = threadIdx.x; it < 32 * Z; it += NT) {
    const int row = it / Z, j = it - row * Z;
    const int gr = c.rank * 32 + row;
    if (gr < p.M) {
      const size_t o = (size_t)gr * Z + j;
      const float lsv = __ldcg(p.ls + o), ev = __ldcg(p.eps + o);
      float d = recv_sum(recv, N, row, j);
      if (p.la) d -= p.w * __ldcg(p.z + o);
      float a = d, b = d * (0.5f * expf(0.5f * lsv) * ev);
      if (p.la) {
        b += p.w * 0.5f;
      } else {
        a -= p.w * __ldcg(p.mu + o);
        b += p.w * 0.5f * (1.0f - expf(lsv));
      }
      p.dd[(size_t)gr * 2 * Z + j] = a;
      p.dd[(size_t)gr * 2 * Z + Z + j] = b;
    }
  }
  cluster_sync();
}

// =================================================================================================================
// CTA items: weight gradients with Adagrad (+ prior) in the epilogue                        VAEB.py:397,389-390,426-444
// =================================================================================================================
// epilogue driver: fn(row in the 128-row tile, first column of an 8-column unit, the eight sums)
template <class F>
__device__ __forceinline__ void wgrad_epilogue(Ctx& c, int N, F fn) {
  const int q = c.warp & 3;
  for (int u = c.warp >> 2; u < N / 8; u += 4) {
    float v[8];
    tmem_ld8(c.tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(u * 8), v);
    tc::tmem_ld_wait();
    fn(q * 32 + c.lane, u * 8, v);
  }
  tc::tc_fence_before();
}

// W2, b2 <- Adagrad([h_d | 1]^T . da2); tile = 128 hidden units x 32 pixels
__device__ __noinline__ void item_wg2(Ctx& c, const Params& p, const Hyper& hy, int mt, int nt) {
  {
    const TSrc b{p.da2, TR_DEC2, p.D, p.M, nt * TR_DEC2, p.D, -1};
    stage_T(c.sm, TSrc{p.hd, MP, p.HP, p.M, mt * MP, p.H, p.H}, &b);
  }
  mma_run(c, 2, (p.M + 15) / 16, TR_DEC2 * 128, TR_DEC2);
  const int H = p.H, D = p.D;
  wgrad_epilogue(c, TR_DEC2, [&](int row, int col0, const float* v) {
    const int i = mt * MP + row, n0 = nt * TR_DEC2 + col0;
    if (i > H || n0 >= D) return;
    size_t off[8]; bool ok[8]; float nv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ok[e] = n0 + e < D;
      off[e] = (i < H ? (size_t)p.oW2 + (size_t)i * D : (size_t)p.ob2) + n0 + e;
    }
    adagrad8(p.P, p.ada, off, ok, v, hy, nv);
    if (i < H) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (ok[e]) mirror_put(p.m_dec2, TR_DEC2, p.KH, n0 + e, i, nv[e]);
        else nv[e] = 0.f;
      // dgrad mirror: row = hidden unit i, k = pixel: eight consecutive k = one 16-byte unit (n0 is a multiple of 8)
      constexpr int TB = TR_DGRAD * 128;
      put_unit(p.m_dgrad + ((size_t)(i / TR_DGRAD) * p.KD + (n0 >> 6)) * 2 * TB + unit_off(i % TR_DGRAD, (n0 & 63) >> 3), TB, nv);
    }
  });
}

// W3, b3 <- Adagrad([x | 1]^T . da3), da3 = ([dmu|dls].[W4|W5]^T) * (1 - h_e^2) recomputed into the B tile;
// tile = 128 pixels x 32 hidden units
__device__ __noinline__ void item_wg3(Ctx& c, const Params& p, const Hyper& hy, const float* x, const float* he, int mt, int nt) {
  const int Z = p.Z, Z2 = 2 * p.Z, H = p.H, M = p.M, D = p.D;
  stage_T(c.sm, TSrc{x, MP, D, M, mt * MP, D, D}, nullptr);
  float* dds = reinterpret_cast<float*>(c.sm + SM_SCR);         // [M][2Z]  (scratch region + receive buffer: 48 KB)
  float* w45 = dds + MP * Z2;                                   // [32][2Z + 1]
  for (int i = threadIdx.x; i < M * Z2; i += NT) dds[i] = __ldcg(p.dd + i);
  for (int i = threadIdx.x; i < 32 * Z2; i += NT) {
    const int jj = i / Z2, q = i - jj * Z2;
    const int j = nt * 32 + jj;
    float t = 0.f;
    if (j < H) t = q < Z ? __ldcg(p.P + p.oW4 + (size_t)j * Z + q) : __ldcg(p.P + p.oW5 + (size_t)j * Z + q - Z);
    w45[jj * (Z2 + 1) + q] = t;
  }
  __syncthreads();
  {
    constexpr int TB = 32 * 128;
    const int jj = threadIdx.x & 31, bo = threadIdx.x >> 5;      // 32 hidden units x 16 batch octets
    const int j = nt * 32 + jj;
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int b = bo * 8 + e;
      float t = 0.f;
      if (b < M && j < H) {
        float s = 0.f;
        for (int q = 0; q < Z2; ++q) s = fmaf(dds[b * Z2 + q], w45[jj * (Z2 + 1) + q], s);
        const float hv = __ldcg(he + (size_t)b * p.HP + j);
        t = s * (1.0f - hv * hv);
      }
      v[e] = t;
    }
    uint4 hi, lo;
    split8(v, hi, lo);
    uint8_t* d = c.sm + SM_B + (size_t)(bo >> 3) * 2 * TB + unit_off(jj, bo & 7);
    *reinterpret_cast<uint4*>(d) = hi;
    *reinterpret_cast<uint4*>(d + TB) = lo;
  }
  mma_run(c, 2, (M + 15) / 16, 32 * 128, 32);
  wgrad_epilogue(c, 32, [&](int row, int col0, const float* v) {
    const int i = mt * MP + row, j0 = nt * 32 + col0;
    if (i > D || j0 >= H) return;
    size_t off[8]; bool ok[8]; float nv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      ok[e] = j0 + e < H;
      off[e] = (i < D ? (size_t)p.oW3 + (size_t)i * H : (size_t)p.ob3) + j0 + e;
    }
    adagrad8(p.P, p.ada, off, ok, v, hy, nv);
    if (i < D) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (ok[e]) mirror_put(p.m_enc1, TR_ENC1, p.KD, j0 + e, i, nv[e]);
    }
  });
}

// W1, b1 <- Adagrad([z | 1]^T . da1), computed transposed: tile = 128 hidden units x (Z + 1) latent columns
__device__ __noinline__ void item_wg1(Ctx& c, const Params& p, const Hyper& hy, int mt) {
  const int Z = p.Z, H = p.H, N = p.NZ;
  {
    const TSrc b{p.z, N, Z, p.M, 0, Z, Z};
    stage_T(c.sm, TSrc{p.da1, MP, p.HP, p.M, mt * MP, H, -1}, &b);
  }
  mma_run(c, 2, (p.M + 15) / 16, N * 128, N);
  wgrad_epilogue(c, N, [&](int row, int col0, const float* v) {
    const int i = mt * MP + row;
    if (i >= H || col0 > Z) return;
    size_t off[8]; bool ok[8]; float nv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int q = col0 + e;
      ok[e] = q <= Z;
      off[e] = q < Z ? (size_t)p.oW1 + (size_t)q * H + i : (size_t)p.ob1 + i;
    }
    adagrad8(p.P, p.ada, off, ok, v, hy, nv);
#pragma unroll
    for (int e = 0; e < 8; ++e)
      if (col0 + e < Z) mirror_put(p.m_dz, N, p.KH, col0 + e, i, nv[e]);
  });
}

// W4, b4, W5, b5 <- Adagrad([h_e | 1]^T . [dmu | dls]); tile = 128 hidden units x 2Z
__device__ __noinline__ void item_wg45(Ctx& c, const Params& p, const Hyper& hy, const float* he, int mt) {
  const int Z = p.Z, H = p.H, N = p.NH;
  {
    const TSrc b{p.dd, N, 2 * Z, p.M, 0, 2 * Z, -1};
    stage_T(c.sm, TSrc{he, MP, p.HP, p.M, mt * MP, H, H}, &b);
  }
  mma_run(c, 2, (p.M + 15) / 16, N * 128, N);
  wgrad_epilogue(c, N, [&](int row, int col0, const float* v) {
    const int i = mt * MP + row;
    if (i > H || col0 >= 2 * Z) return;
    size_t off[8]; bool ok[8]; float nv[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int cc = col0 + e;
      ok[e] = cc < 2 * Z;
      const int j = cc < Z ? cc : cc - Z;
      if (i < H) off[e] = (size_t)(cc < Z ? p.oW4 : p.oW5) + (size_t)i * Z + j;
      else off[e] = (size_t)(cc < Z ? p.ob4 : p.ob5) + j;
    }
    adagrad8(p.P, p.ada, off, ok, v, hy, nv);
    if (i < H) {
#pragma unroll
      for (int e = 0; e < 8; ++e)
        if (ok[e]) mirror_put(p.m_heads, N, p.KH, col0 + e, i, nv[e]);
    }
  });
}

// the bound of step s: fixed-order sum of the row partials (VAEB.py:340-344) / Mg
__device__ __noinline__ void item_bound(Ctx& c, const Params& p, int s) {
  float* red = reinterpret_cast<float*>(c.sm + SM_MISC + 256);
  float t = 0.f;
  for (int r = threadIdx.x; r < p.M; r += NT) {
    float rs = __ldcg(p.aux + r);
    for (int q = 0; q < p.n_tiles3; ++q) rs += __ldcg(p.partial + (size_t)r * p.n_tiles3 + q);
    t += rs;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
  __syncthreads();
  if (c.lane == 0) red[c.warp] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    float b = 0.f;
    for (int w = 0; w < NT / 32; ++w) b += red[w];
    p.scalars[s] = p.bmult * b / p.Mg;
  }
  __syncthreads();
}

// =================================================================================================================
__global__ void __launch_bounds__(NT, 1) step_tc_kernel(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  Ctx c;
  c.sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  c.mma_bar = reinterpret_cast<uint64_t*>(c.sm + SM_MISC);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(c.sm + SM_MISC + 16);
  c.mma_phase = 0;
  c.rank = (int)cluster_ctarank(); c.cid = (int)cluster_idx(); c.ncl = (int)cluster_count();
  c.warp = threadIdx.x >> 5; c.lane = threadIdx.x & 31;
  c.trace = nullptr; c.tn = 0;
  if (threadIdx.x == 0) {
    tc::mbar_init(c.mma_bar, 1);
    tc::fence_barrier_init();
  }
  if (c.warp == 1) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  c.tmem = *tmem_slot;
  cluster_sync();                                   // every CTA of the cluster is resident before any remote store

  const int G = gridDim.x;
  unsigned long long target = p.bar_base;
  const Hyper hy{p.lr, p.ada_eps, p.prior, p.p2};
  const bool rec = p.timing != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  const int n1 = p.HP / 16;                         // enc1 / dgrad tiles
  const int n3 = p.n_tiles3;                        // dec2 tiles
  const int m_h1 = (p.H + 1 + MP - 1) / MP;         // 128-row tiles over the hidden units + the ones row
  const int m_h = (p.H + MP - 1) / MP;
  const int m_d1 = (p.D + 1 + MP - 1) / MP;
  const int n_h32 = p.HP / 32;
  // a phase's item list: cluster items first, then CTA items in groups of four (one per rank of a cluster)
  auto n_groups = [](int n_cta_items) { return (n_cta_items + CL - 1) / CL; };

  for (int s = 0; s < p.n_steps; ++s) {
    const float* x = p.batch_order ? p.x_base + (size_t)__ldg(p.batch_order + s) * p.M * p.D : p.x_direct;
    float* he = p.he + (size_t)(s & 1) * MP * p.HP;              // double buffered: the W4/W5 update of step s runs in P1 of s+1
    const float* he_prev = p.he + (size_t)((s & 1) ^ 1) * MP * p.HP;
    long long* tm = rec ? p.timing + (size_t)s * (N_PHASES + 1) : nullptr;
    if (tm) tm[0] = gtime();
    if (p.timing && blockIdx.x == 0 && s == p.n_steps - 1) c.trace = p.timing + (size_t)p.n_steps * (N_PHASES + 1);

    // ---- P1: encoder hidden layer | W4, W5 update of the previous step ---------------------------------------
    {
      const int n45 = s > 0 ? m_h1 : 0;
      for (int it = c.cid; it < n1 + n_groups(n45); it += c.ncl) {
        if (it < n1) { item_enc1(c, p, x, he, it); continue; }
        const int i = (it - n1) * CL + c.rank;
        if (i < n45) item_wg45(c, p, hy, he_prev, i);
      }
    }
    ST2_TRACE(c, 81);
    grid_barrier(p.bar, target += G);
    if (tm) tm[1] = gtime();
    ST2_TRACE(c, 91);
    // ---- P2: latent heads ------------------------------------------------------------------------------------
    if (c.cid == 0) item_heads(c, p, he, p.step0 + (uint32_t)s);
    ST2_TRACE(c, 82);
    grid_barrier(p.bar, target += G);
    if (tm) tm[2] = gtime();
    ST2_TRACE(c, 92);
    // ---- P3: decoder + log-likelihood ------------------------------------------------------------------------
    for (int it = c.cid; it < n3; it += c.ncl) item_dec2(c, p, x, it);
    ST2_TRACE(c, 83);
    grid_barrier(p.bar, target += G);
    if (tm) tm[3] = gtime();
    ST2_TRACE(c, 93);
    // ---- P4: back through the decoder output layer ---------------------------------------------------------------
    for (int it = c.cid; it < n1; it += c.ncl) item_dgrad(c, p, it);
    ST2_TRACE(c, 84);
    grid_barrier(p.bar, target += G);
    if (tm) tm[4] = gtime();
    ST2_TRACE(c, 94);
    // ---- P5: dz | W2 update | the bound ----------------------------------------------------------------------------
    {
      const int n2 = m_h1 * n3;
      for (int it = c.cid; it < 1 + n_groups(n2 + 1); it += c.ncl) {
        if (it == 0) { item_dz(c, p); continue; }
        const int i = (it - 1) * CL + c.rank;
        if (i < n2) item_wg2(c, p, hy, i / n3, i % n3);
        else if (i == n2) item_bound(c, p, s);
      }
    }
    ST2_TRACE(c, 85);
    grid_barrier(p.bar, target += G);
    if (tm) tm[5] = gtime();
    ST2_TRACE(c, 95);
    // ---- P6: W3 update | W1 update ---------------------------------------------------------------------------------
    {
      const int n3w = m_d1 * n_h32;
      for (int it = c.cid; it < n_groups(n3w + m_h); it += c.ncl) {
        const int i = it * CL + c.rank;
        if (i < n3w) item_wg3(c, p, hy, x, he, i / n_h32, i % n_h32);
        else if (i < n3w + m_h) item_wg1(c, p, hy, i - n3w);
      }
    }
    ST2_TRACE(c, 86);
    grid_barrier(p.bar, target += G);
    if (tm) tm[6] = gtime();
    ST2_TRACE(c, 96);
  }
  // ---- tail: the W4, W5 update of the last step (nothing else runs: no barrier needed after it) ---------------------
  {
    const float* he_last = p.he + (size_t)((p.n_steps - 1) & 1) * MP * p.HP;
    for (int it = c.cid; it < n_groups(m_h1); it += c.ncl) {
      const int i = it * CL + c.rank;
      if (i < m_h1) item_wg45(c, p, hy, he_last, i);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  cluster_sync();                                   // no CTA leaves while a peer may still write its shared memory
  if (c.warp == 1) tc::tmem_dealloc(c.tmem, TMEM_COLS);
}

// ---- mirrors from the fp32 master parameters (after set_tensors / load / an update by another path) ----------------
struct MirrorArgs {
  const float* P; int64_t oW3, oW4, oW5, oW1, oW2;
  uint8_t *m_enc1, *m_heads, *m_dec2, *m_dgrad, *m_dz;
  int D, H, Z, HP, KD, KH, NH, NZ;
};
__global__ void __launch_bounds__(256) build_mirrors_kernel(MirrorArgs a) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int D = a.D, H = a.H, Z = a.Z;
  switch (blockIdx.y) {
    case 0: {   // enc1: n = hidden (HP rows), k = pixel (KD chunks)
      const int K = a.KD * 64;
      if (i >= (int64_t)a.HP * K) return;
      const int n = (int)(i / K), k = (int)(i % K);
      mirror_put(a.m_enc1, TR_ENC1, a.KD, n, k, (n < H && k < D) ? a.P[a.oW3 + (size_t)k * H + n] : 0.f);
      break;
    }
    case 1: {   // heads: n = [W4 cols | W5 cols] (NH rows), k = hidden
      const int K = a.KH * 64;
      if (i >= (int64_t)a.NH * K) return;
      const int n = (int)(i / K), k = (int)(i % K);
      float v = 0.f;
      if (k < H && n < 2 * Z) v = n < Z ? a.P[a.oW4 + (size_t)k * Z + n] : a.P[a.oW5 + (size_t)k * Z + n - Z];
      mirror_put(a.m_heads, a.NH, a.KH, n, k, v);
      break;
    }
    case 2: {   // dec2: n = pixel (tiles of 32), k = hidden
      const int K = a.KH * 64, NR = (D + TR_DEC2 - 1) / TR_DEC2 * TR_DEC2;
      if (i >= (int64_t)NR * K) return;
      const int n = (int)(i / K), k = (int)(i % K);
      mirror_put(a.m_dec2, TR_DEC2, a.KH, n, k, (n < D && k < H) ? a.P[a.oW2 + (size_t)k * D + n] : 0.f);
      break;
    }
    case 3: {   // dgrad: n = hidden (HP rows), k = pixel
      const int K = a.KD * 64;
      if (i >= (int64_t)a.HP * K) return;
      const int n = (int)(i / K), k = (int)(i % K);
      mirror_put(a.m_dgrad, TR_DGRAD, a.KD, n, k, (n < H && k < D) ? a.P[a.oW2 + (size_t)n * D + k] : 0.f);
      break;
    }
    default: {  // dz: n = latent (NZ rows), k = hidden
      const int K = a.KH * 64;
      if (i >= (int64_t)a.NZ * K) return;
      const int n = (int)(i / K), k = (int)(i % K);
      mirror_put(a.m_dz, a.NZ, a.KH, n, k, (n < Z && k < H) ? a.P[a.oW1 + (size_t)n * H + k] : 0.f);
      break;
    }
  }
}

}  // namespace st2

// =====================================================================================================================
// host side
// =====================================================================================================================
bool step_tc_supported(const vaeb_handle* h, int rows) {
  const int e = h->cfg.estimator;
  if (h->steptc.unavailable || h->steptc_off) return false;
  return (e == VAEB_EST_LB || e == VAEB_EST_LA) && !h->cont && h->L == 1 && h->world == 1 &&
         h->cfg.precision != VAEB_PREC_BF16 && h->optimizer == VAEB_OPT_ADAGRAD && rows >= 1 && rows <= st2::MP &&
         (h->D % 8) == 0 && (h->H % 4) == 0 && h->D >= 64 && h->H >= 64 && h->D <= 1024 && h->H <= 1024 && h->Z >= 1 &&
         h->Z <= 24;
}

void step_tc_free(StepTcState& s) {
  void* ptrs[] = {s.bar, s.m_enc1, s.m_heads, s.m_dec2, s.m_dgrad, s.m_dz, s.he, s.hd, s.da2, s.da1, s.dd, s.mu, s.ls,
                  s.eps, s.z, s.partial, s.aux, s.d_order, s.d_timing};
  for (void* q : ptrs)
    if (q) cudaFree(q);
  s = StepTcState();
}

static int step_tc_init(vaeb_handle* h) {
  using namespace st2;
  StepTcState& s = h->steptc;
  const int D = h->D, H = h->H, Z = h->Z;
  const int HP = (H + 63) / 64 * 64, KD = (D + 63) / 64, KH = HP / 64;
  const int NH = (2 * Z + 15) / 16 * 16, NZ = (Z + 1 + 15) / 16 * 16;
  const int n3 = (D + TR_DEC2 - 1) / TR_DEC2;
  VAEB_CUDA(cudaFuncSetAttribute(step_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  // how many clusters of four fit at once: the grid must be co-resident (grid barriers)
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(CL * 64); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = SMEM_BYTES;
  cudaLaunchAttribute at{};
  at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
  cfg.attrs = &at; cfg.numAttrs = 1;
  int ncl = 0;
  VAEB_CUDA(cudaOccupancyMaxActiveClusters(&ncl, step_tc_kernel, &cfg));
  if (ncl < 8) { s.unavailable = true; return VAEB_OK; }
  s.n_cta = CL * std::min(ncl, 33);
  auto alloc = [](void** q, size_t bytes) -> cudaError_t {
    cudaError_t e = cudaMalloc(q, bytes);
    if (e == cudaSuccess) e = cudaMemset(*q, 0, bytes);
    return e;
  };
  VAEB_CUDA(alloc((void**)&s.bar, sizeof(unsigned long long)));
  VAEB_CUDA(alloc((void**)&s.m_enc1, (size_t)HP * KD * 64 * 4));
  VAEB_CUDA(alloc((void**)&s.m_dgrad, (size_t)HP * KD * 64 * 4));
  VAEB_CUDA(alloc((void**)&s.m_heads, (size_t)NH * KH * 64 * 4));
  VAEB_CUDA(alloc((void**)&s.m_dec2, (size_t)n3 * TR_DEC2 * KH * 64 * 4));
  VAEB_CUDA(alloc((void**)&s.m_dz, (size_t)NZ * KH * 64 * 4));
  VAEB_CUDA(alloc((void**)&s.he, (size_t)2 * MP * HP * 4));
  VAEB_CUDA(alloc((void**)&s.hd, (size_t)MP * HP * 4));
  VAEB_CUDA(alloc((void**)&s.da1, (size_t)MP * HP * 4));
  VAEB_CUDA(alloc((void**)&s.da2, (size_t)MP * D * 4));
  VAEB_CUDA(alloc((void**)&s.dd, (size_t)MP * 2 * Z * 4));
  VAEB_CUDA(alloc((void**)&s.mu, (size_t)MP * Z * 4));
  VAEB_CUDA(alloc((void**)&s.ls, (size_t)MP * Z * 4));
  VAEB_CUDA(alloc((void**)&s.eps, (size_t)MP * Z * 4));
  VAEB_CUDA(alloc((void**)&s.z, (size_t)MP * Z * 4));
  VAEB_CUDA(alloc((void**)&s.partial, (size_t)MP * n3 * 4));
  VAEB_CUDA(alloc((void**)&s.aux, (size_t)MP * 4));
  s.ready = true;
  return VAEB_OK;
}

int step_tc_launch(vaeb_handle* h, const int* d_order, const float* d_xrows, int rows, int n_steps, const float* d_eps,
                   int slot0, long long* d_timing) {
  using namespace st2;
  StepTcState& s = h->steptc;
  const Layout& l = h->lay;
  const int D = h->D, H = h->H, Z = h->Z;
  if (!s.ready) {
    VAEB_TRY(step_tc_init(h));
    if (s.unavailable) { vaeb_set_error("step_tc: thread-block clusters of 4 are not available on this device"); return VAEB_ESTATE; }
  }
  Params p{};
  p.D = D; p.H = H; p.Z = Z; p.M = rows;
  p.HP = (H + 63) / 64 * 64; p.KD = (D + 63) / 64; p.KH = p.HP / 64;
  p.NH = (2 * Z + 15) / 16 * 16; p.NZ = (Z + 1 + 15) / 16 * 16;
  p.n_tiles3 = (D + TR_DEC2 - 1) / TR_DEC2;
  p.la = h->cfg.estimator == VAEB_EST_LA ? 1 : 0;
  const bool fb = h->cfg.variant == VAEB_VARIANT_FULLBAYES;
  p.w = fb ? 1.0f / (float)rows : 1.0f;
  p.lr = h->cfg.learning_rate; p.ada_eps = h->cfg.adagrad_eps;
  p.prior = fb ? 0.f : h->cfg.prior_scale;
  p.p2 = fb ? h->cfg.learning_rate * 1e-6f : 0.f;
  p.P = h->d_params; p.ada = h->d_ada;
  p.oW3 = l.off[l.iW3]; p.oW4 = l.off[l.iW4]; p.oW5 = l.off[l.iW5]; p.oW1 = l.off[l.iW1]; p.oW2 = l.off[l.iW2];
  p.ob3 = l.off[l.ib3]; p.ob4 = l.off[l.ib4]; p.ob5 = l.off[l.ib5]; p.ob1 = l.off[l.ib1]; p.ob2 = l.off[l.ib2];
  p.m_enc1 = s.m_enc1; p.m_heads = s.m_heads; p.m_dec2 = s.m_dec2; p.m_dgrad = s.m_dgrad; p.m_dz = s.m_dz;
  p.x_base = (d_order && d_xrows) ? d_xrows : h->d_x; p.batch_order = d_order; p.x_direct = d_xrows;
  p.eps_inj = d_eps;
  p.seed = h->cfg.seed; p.step0 = h->step; p.row_offset = 0;
  p.he = s.he; p.hd = s.hd; p.da2 = s.da2; p.da1 = s.da1; p.dd = s.dd; p.mu = s.mu; p.ls = s.ls; p.eps = s.eps; p.z = s.z;
  p.partial = s.partial; p.aux = s.aux;
  p.scalars = h->d_scalars + slot0; p.Mg = (float)rows; p.bmult = 1.0f;
  p.n_steps = n_steps;
  p.bar = s.bar; p.bar_base = s.bar_count;
  p.timing = d_timing;
  if (!s.mirrors_valid) {
    MirrorArgs a{h->d_params, p.oW3, p.oW4, p.oW5, p.oW1, p.oW2, s.m_enc1, s.m_heads, s.m_dec2, s.m_dgrad, s.m_dz,
                 D, H, Z, p.HP, p.KD, p.KH, p.NH, p.NZ};
    int64_t most = (int64_t)p.HP * p.KD * 64;
    most = std::max<int64_t>(most, (int64_t)p.n_tiles3 * TR_DEC2 * p.KH * 64);
    most = std::max<int64_t>(most, (int64_t)p.NH * p.KH * 64);
    build_mirrors_kernel<<<dim3((unsigned)((most + 255) / 256), 5), 256, 0, h->stream>>>(a);
    VAEB_CUDA(cudaGetLastError());
    ++h->launches;
    s.mirrors_valid = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(s.n_cta); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = SMEM_BYTES; cfg.stream = h->stream;
  cudaLaunchAttribute at[2]{};
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
  cfg.attrs = at; cfg.numAttrs = 2;
  VAEB_CUDA(cudaLaunchKernelEx(&cfg, step_tc_kernel, p));
  s.bar_count += (unsigned long long)s.n_cta * (unsigned long long)N_PHASES * (unsigned long long)n_steps;
  ++h->launches;
  h->step += (uint32_t)n_steps;
  h->grads_have_prior = false;
  return VAEB_OK;
}
