// fp32 FFMA tiled GEMM with fused prologue/epilogue hooks -- the VAEB_PREC_FP32 dense layers.
// C(m,n) = epi( sum_k A(m,k) * B(k,n) ), every operand addressable transposed, so the five
// GEMM shapes of one AEVB step (VAEB.py:246,257 forward; T.grad of them, VAEB.py:397) reuse
// one kernel.  256 threads, BK = 16, register-prefetch double buffering, float4 LDS.
#pragma once
#include <cuda_runtime.h>
#include "activations.cuh"
#include <cstdint>

struct GemmOperands {
  const float* A;    // pass 0
  const float* B;
  const float* A2;   // pass 1 (TWO_PASS: C = A.B + A2.B2)
  const float* B2;   // pass 1, or the second B of DUAL_B (two accumulators sharing A)
  int lda, ldb;
  int M, N, K;
};

// --- epilogue functors ---------------------------------------------------------------
// kRowReduce: operator() returns a per-element term that the kernel sums over the tile's
// columns (deterministic order) and writes to partial[m * n_col_tiles + col_tile].

struct EpiStore {  // out = acc
  static constexpr bool kRowReduce = false;
  float* out; int ld;
  __device__ __forceinline__ float operator()(int m, int n, float acc, float) const {
    out[(size_t)m * ld + n] = acc; return 0.f;
  }
};

struct EpiBiasAct {  // out = f(acc + bias[n]); act: 0 identity, 1 tanh, 2 sigmoid, 3 ReLU
  static constexpr bool kRowReduce = false;
  const float* bias; float* out; int ld; int act;
  __device__ __forceinline__ float operator()(int m, int n, float acc, float) const {
    out[(size_t)m * ld + n] = act_fwd(acc + bias[n], act); return 0.f;
  }
};

struct EpiWgrad {  // rows < H go to gW[H,N], the ones-row (m == H) is the bias gradient
  static constexpr bool kRowReduce = false;
  float* gW; float* gb; int H; int ld;
  __device__ __forceinline__ float operator()(int m, int n, float acc, float) const {
    if (m < H) gW[(size_t)m * ld + n] = acc; else gb[n] = acc;
    return 0.f;
  }
};

struct EpiMulOneMinusSq {  // out = acc * f'(h): backprop through the hidden activation (tanh: 1 - h^2)
  static constexpr bool kRowReduce = false;
  const float* h; float* out; int ld; int act;
  __device__ __forceinline__ float operator()(int m, int n, float acc, float) const {
    const float hv = h[(size_t)m * ld + n];
    out[(size_t)m * ld + n] = acc * act_bwd(hv, act); return 0.f;
  }
};

__device__ __forceinline__ float softplusf(float a) {  // log(1+e^a), stable
  return fmaxf(a, 0.f) + log1pf(expf(-fabsf(a)));
}
__device__ __forceinline__ float sigmoidf(float a) { return 1.0f / (1.0f + expf(-a)); }

// Bernoulli decoder head (VAEB.py:263,311): term = x*a - softplus(a); d_a = scale*(x - sigmoid(a)).
// The x row of GEMM row r is (r / x_div) % x_mod: training rows are l*M+m (x_div=1,x_mod=M),
// importance-sampling rows are i*L+l (x_div=L).
struct EpiBernoulli {
  static constexpr bool kRowReduce = true;
  const float* bias; const float* x; int ldx; int x_div; int x_mod;
  float scale; float* da; int ldda;  // da == nullptr: evaluation only
  __device__ __forceinline__ float operator()(int m, int n, float acc, float) const {
    const float a = acc + bias[n];
    const float xv = x[(size_t)((m / x_div) % x_mod) * ldx + n];
    if (da) da[(size_t)m * ldda + n] = scale * (xv - sigmoidf(a));
    return xv * a - softplusf(a);
  }
};

// Gaussian decoder head (VAEB.py:257-258,306-307): mu_x = sigmoid(a), lv = second accumulator.
struct EpiGaussian {
  static constexpr bool kRowReduce = true;
  const float* bias; const float* bias2; const float* x; int ldx; int x_div; int x_mod;
  float scale; float* da; float* dlv; int ldda;
  __device__ __forceinline__ float operator()(int m, int n, float acc, float acc2) const {
    const float a = acc + bias[n];
    const float lv = acc2 + bias2[n];
    const float xv = x[(size_t)((m / x_div) % x_mod) * ldx + n];
    const float mu = sigmoidf(a);
    const float d = xv - mu;
    const float r = d * expf(-lv);
    if (da) {
      da[(size_t)m * ldda + n] = scale * r * mu * (1.0f - mu);
      dlv[(size_t)m * ldda + n] = scale * (-0.5f + 0.5f * d * r);
    }
    return -0.91893853320467274178f - 0.5f * lv - 0.5f * d * r;
  }
};

// AE baselines: logpdf.bernoulli with 1e-7 inside both logs (degenerate-vae/logpdf.py:85-86, ae.py:62):
// term = x log(P+e) + (1-x) log(1-P+e), P = sigmoid(a); d_a = (x/(P+e) - (1-x)/(1-P+e)) P (1-P)
struct EpiBernoulliClamp {
  static constexpr bool kRowReduce = true;
  const float* bias; const float* x; int ldx; float* da; int ldda;
  __device__ __forceinline__ float operator()(int m, int n, float acc, float) const {
    const float a = acc + bias[n];
    const float xv = x[(size_t)m * ldx + n];
    const float P = sigmoidf(a), Q = sigmoidf(-a);           // Q = 1 - P without cancellation
    if (da) da[(size_t)m * ldda + n] = (xv / (P + 1e-7f) - (1.0f - xv) / (Q + 1e-7f)) * P * Q;
    return xv * logf(P + 1e-7f) + (1.0f - xv) * logf(Q + 1e-7f);
  }
};
// vanilla AE (vanilla-ae/ae.py:66-72): term = -(x - P)^2 (the row sums give -se); d_a = 2 (x - P) P (1-P)
struct EpiSquaredError {
  static constexpr bool kRowReduce = true;
  const float* bias; const float* x; int ldx; float* da; int ldda;
  __device__ __forceinline__ float operator()(int m, int n, float acc, float) const {
    const float a = acc + bias[n];
    const float xv = x[(size_t)m * ldx + n];
    const float P = sigmoidf(a), d = xv - P;
    if (da) da[(size_t)m * ldda + n] = 2.0f * d * P * sigmoidf(-a);
    return -d * d;
  }
};

// reconstruct (VAEB.py:282-292): y += sigmoid(a)/n, lv += lv/n
struct EpiReconAccum {
  static constexpr bool kRowReduce = false;
  const float* bias; const float* bias2; float* y; float* lvout; int ld; float inv_n; int first;
  __device__ __forceinline__ float operator()(int m, int n, float acc, float acc2) const {
    const size_t o = (size_t)m * ld + n;
    const float yv = sigmoidf(acc + bias[n]) * inv_n;
    y[o] = first ? yv : y[o] + yv;
    if (lvout) { const float l = (acc2 + bias2[n]) * inv_n; lvout[o] = first ? l : lvout[o] + l; }
    return 0.f;
  }
};

// --- kernel ------------------------------------------------------------------------------
template <int BM, int BN, int TM, int TN, bool TA, bool TB, bool ONES_ROW, bool DUAL_B, bool TWO_PASS, class Epi>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_f32_kernel(const GemmOperands g, const Epi epi, float* __restrict__ partial) {
  constexpr int BK = 16;
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int TXN = BN / TN;  // threads along n (16 for both tile shapes used)
  static_assert(TXN == 16, "row reduction assumes 16 threads along n");
  static_assert((BM * BK) % NT == 0 && (BN * BK) % NT == 0, "loader shape");
  constexpr int A_PER = BM * BK / NT, B_PER = BN * BK / NT;
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  __shared__ __align__(16) float Bs2[DUAL_B ? BK : 1][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % TXN, ty = tid / TXN;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  float acc[TM][TN], acc2[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) { acc[i][j] = 0.f; acc2[i][j] = 0.f; }

  float ra[A_PER], rb[B_PER], rb2[B_PER];

  auto load_tile = [&](const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ Bd, int kt) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int e = tid + i * NT;
      int r, k;
      if (TA) { k = e / BM; r = e % BM; } else { r = e / BK; k = e % BK; }
      const int gm = m0 + r, gk = kt + k;
      float v = 0.f;
      if (gm < g.M && gk < g.K) {
        if (ONES_ROW && gm == g.M - 1) v = 1.0f;
        else v = TA ? A[(size_t)gk * g.lda + gm] : A[(size_t)gm * g.lda + gk];
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + i * NT;
      int n, k;
      if (TB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      const int gn = n0 + n, gk = kt + k;
      float v = 0.f, v2 = 0.f;
      if (gn < g.N && gk < g.K) {
        const size_t o = TB ? (size_t)gn * g.ldb + gk : (size_t)gk * g.ldb + gn;
        v = B[o];
        if (DUAL_B) v2 = Bd[o];
      }
      rb[i] = v; rb2[i] = v2;
    }
  };
  auto store_tile = [&]() {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int e = tid + i * NT;
      int r, k;
      if (TA) { k = e / BM; r = e % BM; } else { r = e / BK; k = e % BK; }
      As[k][r] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + i * NT;
      int n, k;
      if (TB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      Bs[k][n] = rb[i];
      if (DUAL_B) Bs2[k][n] = rb2[i];
    }
  };

  const int ntiles = (g.K + BK - 1) / BK;
#pragma unroll 1
  for (int pass = 0; pass < (TWO_PASS ? 2 : 1); ++pass) {
    const float* A = (TWO_PASS && pass) ? g.A2 : g.A;
    const float* B = (TWO_PASS && pass) ? g.B2 : g.B;
    const float* Bd = DUAL_B ? g.B2 : nullptr;
    load_tile(A, B, Bd, 0);
#pragma unroll 1
    for (int t = 0; t < ntiles; ++t) {
      store_tile();
      __syncthreads();
      if (t + 1 < ntiles) load_tile(A, B, Bd, (t + 1) * BK);
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        float a[TM], b[TN], b2[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
        for (int j = 0; j < TN; ++j) { b[j] = Bs[k][tx * TN + j]; if (DUAL_B) b2[j] = Bs2[k][tx * TN + j]; }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) {
            acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
            if (DUAL_B) acc2[i][j] = fmaf(a[i], b2[j], acc2[i][j]);
          }
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int m = m0 + ty * TM + i;
    float s = 0.f;
    if (m < g.M) {
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int n = n0 + tx * TN + j;
        if (n < g.N) s += epi(m, n, acc[i][j], acc2[i][j]);
      }
    }
    if (Epi::kRowReduce) {
      // all 32 lanes take part: a warp covers two rows (16 lanes each)
      s += __shfl_xor_sync(0xffffffffu, s, 8);
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (tx == 0 && m < g.M) partial[(size_t)m * gridDim.x + blockIdx.x] = s;
    }
  }
}
