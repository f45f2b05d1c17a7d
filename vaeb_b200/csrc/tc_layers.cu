// The wide layers of the AEVB step on the 5th-generation tensor cores: tcgen05.mma with bf16
// operands staged by TMA (128B swizzle), fp32 accumulation in TMEM, fused epilogues read back
// with tcgen05.ld.  NS = 1: plain bf16 operands (1e-2 tier).  NS = 2: every operand is carried as
// bf16 hi + bf16 lo and each k-step issues hi*hi + hi*lo + lo*hi into the same accumulator
// ("bf16x3": ~2^-17 relative per product, the fp32 parity tier on tensor cores).
//   enc1      h_e  = tanh(x.W3 + b3)                       VAEB.py:246     A K-major,  B MN-major
//   dec2      a    = h_d.W2 + b2 -> Bernoulli log-lik, da  VAEB.py:263,311 A K-major,  B MN-major
//   dgrad     da1  = (da.W2^T) * (1 - h_d^2)               T.grad :397     A K-major,  B K-major
//   wgrad     gW   = [act|1]^T . delta (bias row included) T.grad :397     A MN-major, B MN-major
// Warp roles (640 threads): w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w4-19 epilogue (four
// warps per TMEM lane quarter, each owning a quarter of the tile's columns: at M = 100 the whole
// layer is 8-13 CTAs, so the elementwise epilogue needs all the threads it can get).
#include <cuda_bf16.h>

#include "common.cuh"
#include "launchers.h"
#include "tc_common.cuh"
#include "tc_layers.h"

namespace {

constexpr int BM = 128, BK = 64;
constexpr int EPI_WARPS = 16, TC_THREADS = 128 + EPI_WARPS * 32;

__device__ __forceinline__ float softplusf_(float a) { return fmaxf(a, 0.f) + log1pf(expf(-fabsf(a))); }
__device__ __forceinline__ float sigmoidf_(float a) { return 1.0f / (1.0f + expf(-a)); }
__device__ __forceinline__ void put_split(__nv_bfloat16* hi, __nv_bfloat16* lo, size_t o, float v) {
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  hi[o] = h;
  if (lo) lo[o] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// ---- epilogues: one thread owns one accumulator row, 16 columns per call --------------------
struct EpiTanh {            // out[row, col] = tanh(acc + bias[col])
  const float* bias; float* out; int ld;
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ void chunk(int row, bool ok, int col0, int N, const float* v) {
    if (!ok) return;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < N) out[(size_t)row * ld + col0 + j] = tanhf(v[j] + bias[col0 + j]);
  }
  __device__ __forceinline__ void end(int, bool, int, int) {}
};

struct EpiBernoulliTc {     // VAEB.py:263,311: term = x*a - softplus(a); da = scale*(x - sigmoid(a))
  const float* bias; const float* x; int ldx; int x_div; int x_mod; float scale;
  __nv_bfloat16* da_hi; __nv_bfloat16* da_lo; int ldda; float* partial;
  float acc;
  __device__ __forceinline__ void begin() { acc = 0.f; }
  __device__ __forceinline__ void chunk(int row, bool ok, int col0, int N, const float* v) {
    if (!ok) return;
    const float* xr = x + (size_t)((row / x_div) % x_mod) * ldx;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int c = col0 + j;
      if (c < N) {
        const float a = v[j] + bias[c];
        const float xv = xr[c];
        acc += xv * a - softplusf_(a);
        if (da_hi) put_split(da_hi, da_lo, (size_t)row * ldda + c, scale * (xv - sigmoidf_(a)));
      }
    }
  }
  __device__ __forceinline__ void end(int row, bool ok, int tile_n, int n_tiles) {
    if (ok) partial[(size_t)row * n_tiles + tile_n] = acc;
  }
};

struct EpiWgradTc {         // rows < Hreal -> gW[Hreal, N]; row == Hreal (the ones column of A) -> gb
  float* gW; float* gb; int Hreal; int ld;
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ void chunk(int row, bool ok, int col0, int N, const float* v) {
    if (!ok) return;
    float* dst = row < Hreal ? gW + (size_t)row * ld : gb;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < N) dst[col0 + j] = v[j];
  }
  __device__ __forceinline__ void end(int, bool, int, int) {}
};

struct EpiDgradTanh {       // out = acc * (1 - h^2)
  const float* h; float* out; int ld;
  __device__ __forceinline__ void begin() {}
  __device__ __forceinline__ void chunk(int row, bool ok, int col0, int N, const float* v) {
    if (!ok) return;
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (col0 + j < N) {
        const float hv = h[(size_t)row * ld + col0 + j];
        out[(size_t)row * ld + col0 + j] = v[j] * (1.0f - hv * hv);
      }
  }
  __device__ __forceinline__ void end(int, bool, int, int) {}
};

template <int BN, int NS>
struct LayerSmem {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = NS * (A_BYTES + B_BYTES);
  static constexpr int STAGES = (200 * 1024 / STAGE_BYTES) > 6 ? 6 : (200 * 1024 / STAGE_BYTES);
  static constexpr int TOTAL = STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(STAGES >= 2, "tile too large");
};

struct LayerMaps {            // hi/lo tensor maps of both operands (lo unused when NS == 1)
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

template <int BN, bool A_MN, bool B_MN, int NS, class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_layer_kernel(const __grid_constant__ LayerMaps maps, Epi epi, int M, int N, int K, int a_row_off) {
  using S = LayerSmem<BN, NS>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + S::STAGES * S::STAGE_BYTES);
  uint64_t* empty = full + S::STAGES;
  uint64_t* tmem_full = empty + S::STAGES;
  uint32_t* tmem_slot = (uint32_t*)(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int nkb = (K + BK - 1) / BK;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&maps.a_hi);
    tc::tma_prefetch_desc(&maps.b_hi);
    if (NS == 2) { tc::tma_prefetch_desc(&maps.a_lo); tc::tma_prefetch_desc(&maps.b_lo); }
    for (int s = 0; s < S::STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    tc::mbar_init(tmem_full, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S::STAGES;
      tc::mbar_wait(&empty[s], ((kb / S::STAGES) & 1) ^ 1);
      uint8_t* base = smem + s * S::STAGE_BYTES;
      tc::mbar_expect_tx(&full[s], S::STAGE_BYTES);
#pragma unroll
      for (int sp = 0; sp < NS; ++sp) {
        uint8_t* a = base + sp * S::A_BYTES;
        uint8_t* b = base + NS * S::A_BYTES + sp * S::B_BYTES;
        const CUtensorMap* ta = sp ? &maps.a_lo : &maps.a_hi;
        const CUtensorMap* tb = sp ? &maps.b_lo : &maps.b_hi;
        if (A_MN) {   // A stored [K rows, M contiguous]: the row offset applies to the K coordinate
          for (int g = 0; g < BM / 64; ++g) tc::tma_load_2d(a + g * 8192, ta, &full[s], m0 + g * 64, a_row_off + kb * BK);
        } else {      // A stored [M rows, K contiguous]
          tc::tma_load_2d(a, ta, &full[s], kb * BK, a_row_off + m0);
        }
        if (B_MN) {
          for (int g = 0; g < BN / 64; ++g) tc::tma_load_2d(b + g * 8192, tb, &full[s], n0 + g * 64, kb * BK);
        } else {
          tc::tma_load_2d(b, tb, &full[s], kb * BK, n0);
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer =====
    constexpr uint32_t idesc = tc::make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % S::STAGES;
      tc::mbar_wait(&full[s], (kb / S::STAGES) & 1);
      tc::tc_fence_after();
      const uint32_t a = tc::smem_u32(smem + s * S::STAGE_BYTES);
      const uint32_t b = a + NS * S::A_BYTES;
#pragma unroll
      for (int k = 0; k < BK / 16; ++k) {
        const uint64_t dah = A_MN ? tc::desc_mnmajor(a, k, 8192u) : tc::desc_kmajor(a, k);
        const uint64_t dbh = B_MN ? tc::desc_mnmajor(b, k, 8192u) : tc::desc_kmajor(b, k);
        tc::umma_bf16(tmem_base, dah, dbh, idesc, (kb | k) != 0 ? 1u : 0u);
        if (NS == 2) {
          const uint64_t dal = A_MN ? tc::desc_mnmajor(a + S::A_BYTES, k, 8192u) : tc::desc_kmajor(a + S::A_BYTES, k);
          const uint64_t dbl = B_MN ? tc::desc_mnmajor(b + S::B_BYTES, k, 8192u) : tc::desc_kmajor(b + S::B_BYTES, k);
          tc::umma_bf16(tmem_base, dah, dbl, idesc, 1u);
          tc::umma_bf16(tmem_base, dal, dbh, idesc, 1u);
        }
      }
      tc::umma_commit(&empty[s]);
    }
    tc::umma_commit(tmem_full);
  } else if (warp >= 4) {
    // ===== epilogue: warp e = warp-4 reads lane quarter e%4, column slice e/4 of the tile =====
    const int q = warp & 3, cs = (warp - 4) >> 2;
    constexpr int SLICE = BN / (EPI_WARPS / 4);
    const int row = m0 + q * 32 + lane;
    const bool ok = row < M;
    epi.begin();
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int c = cs * SLICE; c < (cs + 1) * SLICE; c += 16) {
      float v[16];
      tc::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
      tc::tmem_ld_wait();
      if (n0 + c < N) epi.chunk(row, ok, n0 + c, N, v);
    }
    epi.end(row, ok, blockIdx.x * (EPI_WARPS / 4) + cs, gridDim.x * (EPI_WARPS / 4));
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int BN, bool A_MN, bool B_MN, int NS, class Epi>
cudaError_t launch_layer(cudaStream_t st, const LayerMaps& maps, const Epi& epi, int M, int N, int K, int a_row_off) {
  using S = LayerSmem<BN, NS>;
  auto kfn = tc_layer_kernel<BN, A_MN, B_MN, NS, Epi>;
  static bool attr_done = false;   // per instantiation
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  kfn<<<grid, TC_THREADS, S::TOTAL, st>>>(maps, epi, M, N, K, a_row_off);
  return cudaGetLastError();
}

template <bool A_MN, bool B_MN, class Epi>
cudaError_t dispatch_layer(cudaStream_t st, int ns, int bn, const LayerMaps& maps, const Epi& epi, int M, int N, int K,
                           int a_row_off) {
  if (ns == 2) {
    if (bn == 128) return launch_layer<128, A_MN, B_MN, 2, Epi>(st, maps, epi, M, N, K, a_row_off);
    return launch_layer<64, A_MN, B_MN, 2, Epi>(st, maps, epi, M, N, K, a_row_off);
  }
  if (bn == 128) return launch_layer<128, A_MN, B_MN, 1, Epi>(st, maps, epi, M, N, K, a_row_off);
  return launch_layer<64, A_MN, B_MN, 1, Epi>(st, maps, epi, M, N, K, a_row_off);
}

// fp32 [rows, cols] (leading dim ld_src) -> bf16 hi (/lo) mirrors [rows, ld_dst]; column `ones_col`
// (>= cols) is set to 1 so that a weight-gradient GEMM over the mirror also yields the bias gradient.
__global__ void __launch_bounds__(256)
split_matrix_kernel(const float* __restrict__ src, int64_t rows, int cols, int ld_src, __nv_bfloat16* __restrict__ hi,
                    __nv_bfloat16* __restrict__ lo, int ld_dst, int ones_col) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * ld_dst) return;
  const int64_t r = i / ld_dst;
  const int c = (int)(i % ld_dst);
  float v = 0.f;
  if (c < cols) v = src[r * ld_src + c];
  else if (c == ones_col) v = 1.0f;
  put_split(hi, lo, (size_t)i, v);
}

struct MirrorSeg { const float* src; __nv_bfloat16* hi; __nv_bfloat16* lo; int rows, cols, ld; };
__global__ void __launch_bounds__(256)
mirror_weights_kernel(MirrorSeg s0, MirrorSeg s1) {
  const MirrorSeg& s = blockIdx.y == 0 ? s0 : s1;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s.src == nullptr || i >= (int64_t)s.rows * s.cols) return;
  const int r = (int)(i / s.cols), c = (int)(i % s.cols);
  put_split(s.hi, s.lo, (size_t)r * s.ld + c, s.src[i]);
}

}  // namespace

// ---- host API (tc_layers.h) -------------------------------------------------------------------
cudaError_t tc_split_matrix(cudaStream_t st, int64_t* launches, const float* src, int64_t rows, int cols, int ld_src,
                            void* hi, void* lo, int ld_dst, int ones_col) {
  const int64_t n = rows * ld_dst;
  split_matrix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, rows, cols, ld_src, (__nv_bfloat16*)hi,
                                                                    (__nv_bfloat16*)lo, ld_dst, ones_col);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t tc_mirror_weights(cudaStream_t st, int64_t* launches, const float* w3, void* w3h, void* w3l, int D, int H,
                              int ldh, const float* w2, void* w2h, void* w2l, int ldd) {
  MirrorSeg s0{w3, (__nv_bfloat16*)w3h, (__nv_bfloat16*)w3l, D, H, ldh};
  MirrorSeg s1{w2, (__nv_bfloat16*)w2h, (__nv_bfloat16*)w2l, H, D, ldd};
  const int64_t n = (int64_t)D * H;
  mirror_weights_kernel<<<dim3((unsigned)((n + 255) / 256), 2), 256, 0, st>>>(s0, s1);
  ++*launches;
  return cudaGetLastError();
}

static int make_pair(CUtensorMap* hi, CUtensorMap* lo, const void* bh, const void* bl, uint64_t rows, uint64_t cols,
                     uint64_t stride, uint32_t box_rows) {
  VAEB_TRY(vaeb_make_tmap_bf16(hi, bh, rows, cols, stride, box_rows));
  if (bl) VAEB_TRY(vaeb_make_tmap_bf16(lo, bl, rows, cols, stride, box_rows));
  else *lo = *hi;
  return VAEB_OK;
}

int tc_build_maps(TcMaps* m, const TcBuffers& b, int rows_data, int R, int rows, int D, int H, int bn) {
  // enc1: A = x mirror [rows_data, D] K-major (the ones column at D stays out of the map), B = W3 [D, H] MN-major
  LayerMaps* e1 = reinterpret_cast<LayerMaps*>(m->enc1);
  VAEB_TRY(make_pair(&e1->a_hi, &e1->a_lo, b.xh, b.xl, rows_data, D, b.ldx, BM));
  VAEB_TRY(make_pair(&e1->b_hi, &e1->b_lo, b.w3h, b.w3l, D, H, b.ldh, 64));
  // dec2: A = h_d mirror [R, H] K-major, B = W2 [H, D] MN-major
  LayerMaps* d2 = reinterpret_cast<LayerMaps*>(m->dec2);
  VAEB_TRY(make_pair(&d2->a_hi, &d2->a_lo, b.hdh, b.hdl, R, H, b.ldh, BM));
  VAEB_TRY(make_pair(&d2->b_hi, &d2->b_lo, b.w2h, b.w2l, H, D, b.ldd, 64));
  // dgrad h_d: A = da2 mirror [R, D] K-major, B = W2 [H, D] K-major (N = H rows)
  LayerMaps* dg = reinterpret_cast<LayerMaps*>(m->dgrad);
  VAEB_TRY(make_pair(&dg->a_hi, &dg->a_lo, b.da2h, b.da2l, R, D, b.ldd, BM));
  VAEB_TRY(make_pair(&dg->b_hi, &dg->b_lo, b.w2h, b.w2l, H, D, b.ldd, (uint32_t)bn));
  // wgrad W2: A = h_d mirror [R, H+1] MN-major (ones column -> bias row), B = da2 mirror [R, D] MN-major
  LayerMaps* w2 = reinterpret_cast<LayerMaps*>(m->wgrad2);
  VAEB_TRY(make_pair(&w2->a_hi, &w2->a_lo, b.hdh, b.hdl, R, H + 1, b.ldh, 64));
  VAEB_TRY(make_pair(&w2->b_hi, &w2->b_lo, b.da2h, b.da2l, R, D, b.ldd, 64));
  // wgrad W3: A = x mirror [rows_data, D+1] MN-major, B = da3 mirror [rows, H] MN-major
  LayerMaps* w3 = reinterpret_cast<LayerMaps*>(m->wgrad3);
  VAEB_TRY(make_pair(&w3->a_hi, &w3->a_lo, b.xh, b.xl, rows_data, D + 1, b.ldx, 64));
  VAEB_TRY(make_pair(&w3->b_hi, &w3->b_lo, b.da3h, b.da3l, rows, H, b.ldh, 64));
  return VAEB_OK;
}

static_assert(sizeof(LayerMaps) == TC_LAYER_MAPS_BYTES, "TcMaps storage size");

cudaError_t tc_enc1(cudaStream_t st, int64_t* launches, const TcMaps& m, int ns, int bn, int rows, int D, int H,
                    int x_row_off, const float* b3, float* h_e) {
  EpiTanh epi{b3, h_e, H};
  ++*launches;
  return dispatch_layer<false, true>(st, ns, bn, *reinterpret_cast<const LayerMaps*>(m.enc1), epi, rows, H, D,
                                     x_row_off);
}

cudaError_t tc_dec2_bernoulli(cudaStream_t st, int64_t* launches, const TcMaps& m, int ns, int bn, int R, int H, int D,
                              const float* b2, const float* x, int x_div, int x_mod, float scale, void* da_hi,
                              void* da_lo, int ldda, float* partial, int* n_tiles) {
  EpiBernoulliTc epi{b2, x, D, x_div, x_mod, scale, (__nv_bfloat16*)da_hi, (__nv_bfloat16*)da_lo, ldda, partial, 0.f};
  *n_tiles = ((D + bn - 1) / bn) * (EPI_WARPS / 4);
  ++*launches;
  return dispatch_layer<false, true>(st, ns, bn, *reinterpret_cast<const LayerMaps*>(m.dec2), epi, R, D, H, 0);
}

cudaError_t tc_dgrad_hd(cudaStream_t st, int64_t* launches, const TcMaps& m, int ns, int bn, int R, int D, int H,
                        const float* h_d, float* da1) {
  EpiDgradTanh epi{h_d, da1, H};
  ++*launches;
  return dispatch_layer<false, false>(st, ns, bn, *reinterpret_cast<const LayerMaps*>(m.dgrad), epi, R, H, D, 0);
}

cudaError_t tc_wgrad2(cudaStream_t st, int64_t* launches, const TcMaps& m, int ns, int bn, int R, int H, int D,
                      float* gW2, float* gb2) {
  EpiWgradTc epi{gW2, gb2, H, D};
  ++*launches;
  return dispatch_layer<true, true>(st, ns, bn, *reinterpret_cast<const LayerMaps*>(m.wgrad2), epi, H + 1, D, R, 0);
}

cudaError_t tc_wgrad3(cudaStream_t st, int64_t* launches, const TcMaps& m, int ns, int bn, int rows, int D, int H,
                      int x_row_off, float* gW3, float* gb3) {
  EpiWgradTc epi{gW3, gb3, D, H};
  ++*launches;
  return dispatch_layer<true, true>(st, ns, bn, *reinterpret_cast<const LayerMaps*>(m.wgrad3), epi, D + 1, H, rows,
                                    x_row_off);
}
