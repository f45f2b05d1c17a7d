// Importance-sampled log p(x) on the 5th-generation tensor cores.
//
// One persistent kernel, one CTA per SM.  A work tile is 128 samples of ONE test point:
//   z = mu + exp(.5 ls)*eps (Philox keyed by the global (point, sample, j))      -- producer warps
//   h = tanh(z.W1 + b1), bf16, written straight into the UMMA K-major SWIZZLE_128B layout in shared
//       memory: the decoder hidden layer never exists in HBM                       -- producer warps
//   a = h.W2 : tcgen05.mma (kind::f16, bf16 x bf16 -> fp32 in TMEM), W2^T streamed through a TMA
//       ring of [192 n x 64 k] boxes, output swept in 192-column chunks, two TMEM accumulators so
//       the epilogue of chunk c overlaps the MMAs of chunk c+1                     -- TMA + MMA warps
//   log w = sum_d x_d a_d - softplus(a_d) + log p(z) - log q(z|x), per sample; then the tile's
//       (max, sum exp) pair -> partial[tile]                                       -- epilogue warps
// A finishing kernel folds the per-tile pairs of a point into log p^(x) = logsumexp - log L.
// bf16 operands: the 1e-2 tier of the north star; the fp32 estimator (api.cu) stays the parity tier.
#include <cuda_bf16.h>

#include "common.cuh"
#include "philox.cuh"
#include "tc_common.cuh"

namespace istc {

constexpr int BM = 128, BK = 64, NC = 192, STAGES = 3;
constexpr int KB_MAX = 8;                       // hidden units padded to <= 512
constexpr int ZMAX = 20;
constexpr int DMAX = 1024;
constexpr int PROD_WARPS = 8, EPI_WARPS = 8;
constexpr int THREADS = (4 + PROD_WARPS + EPI_WARPS) * 32;
constexpr int A_BLOCK = BM * 128;               // bytes of one 64-wide k block of the A tile
constexpr int B_STAGE = NC * 128;
constexpr int TMEM_COLS = 512, ACC_STRIDE = 256;

struct Smem {                                   // offsets from a 1024-byte aligned base
  static constexpr int A = 0;
  static constexpr int B = A + KB_MAX * A_BLOCK;
  static constexpr int ZS = B + STAGES * B_STAGE;            // [128][ZMAX] fp32
  static constexpr int AUX = ZS + BM * ZMAX * 4;             // [2][128]
  static constexpr int XB = AUX + 2 * BM * 4;                // [DMAX] float2 {b2[col], x[col]}; col >= D: {-1e30, 0}
  static constexpr int ROWSUM = XB + DMAX * 8;               // [2][128]
  static constexpr int RED = ROWSUM + 2 * BM * 4;            // [8]
  static constexpr int BARS = RED + 64;                      // mbarriers
  static constexpr int TOTAL = BARS + 256 + 1024;
};
static_assert(Smem::TOTAL <= 232448, "shared memory budget");

struct Params {
  int n_points, L, D, H, Z, KB;                 // KB = ceil(H / 64)
  int tiles_per_point, n_chunks, tail_cols;     // output chunks of 192 columns, the last one `tail_cols` wide (multiple of 16)
  const float* x;                               // [n_points, D]
  const float* mu; const float* ls;             // [n_points, Z]
  const float* W1; const float* b1; const float* b2;
  const float* eps_inj;                         // [n_points, L, Z] or nullptr
  uint64_t seed; int64_t row_offset;
  float2* partial;                              // [n_points * tiles_per_point] (max, sum exp)
  float* logw_out;                              // nullptr or [n_points * L]
};

__device__ long long* g_is_dbg = nullptr;   // optional clock64 trace of CTA 0: [role][64]
#define IS_STAMP(role, idx) do { if (g_is_dbg && blockIdx.x == 0 && (idx) < 64) g_is_dbg[(role) * 64 + (idx)] = clock64(); } while (0)

__device__ __forceinline__ float tanh_approx(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// 16 accumulator columns of one row: rs += x*a - softplus(a), a = acc + b2; {b2, x} pairs at shared address xb
__device__ __forceinline__ void fold16(const float (&v)[16], uint32_t xb, float& rs0, float& rs1) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float bq, xq;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(bq), "=f"(xq) : "r"(xb + 8u * j));
    const float a = v[j] + bq;
    // softplus(a) = max(a,0) + ln2 * log2(1 + 2^(-|a| log2 e))
    float t, lg;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fabsf(a) * -1.4426950408889634f));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(1.0f + t));
    const float sp = fmaf(lg, 0.6931471805599453f, fmaxf(a, 0.f));
    if (j & 1) rs1 += fmaf(xq, a, -sp); else rs0 += fmaf(xq, a, -sp);
  }
}

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

__global__ void __launch_bounds__(THREADS, 1)
is_tc_kernel(const __grid_constant__ CUtensorMap map_full, const __grid_constant__ CUtensorMap map_tail, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + Smem::BARS);
  uint64_t* b_full = bars;                  // [STAGES]
  uint64_t* b_empty = bars + STAGES;        // [STAGES]
  uint64_t* a_full = bars + 2 * STAGES;
  uint64_t* a_empty = a_full + 1;
  uint64_t* acc_full = a_empty + 1;         // [2]
  uint64_t* acc_empty = acc_full + 2;       // [2]
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + 2);
  float* zs = (float*)(smem + Smem::ZS);
  float* aux_s = (float*)(smem + Smem::AUX);
  float2* xb = (float2*)(smem + Smem::XB);
  float* rowsum_s = (float*)(smem + Smem::ROWSUM);
  float* red_s = (float*)(smem + Smem::RED);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.n_points * p.tiles_per_point;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&map_full);
    tc::tma_prefetch_desc(&map_tail);
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(&b_full[s], 1); tc::mbar_init(&b_empty[s], 1); }
    tc::mbar_init(a_full, PROD_WARPS);
    tc::mbar_init(a_empty, 1);
    for (int b = 0; b < 2; ++b) { tc::mbar_init(&acc_full[b], 1); tc::mbar_init(&acc_empty[b], EPI_WARPS); }
    tc::fence_barrier_init();
  }
  if (warp == 2) tc::tmem_alloc(tmem_slot, TMEM_COLS);
  // columns past D: b2 = -1e30 and x = 0 make x*a - softplus(a) exactly 0 (no per-element bounds test)
  for (int i = threadIdx.x; i < DMAX; i += THREADS) xb[i] = make_float2(i < p.D ? p.b2[i] : -1e30f, 0.f);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA: stream W2^T boxes [chunk rows x 64 k] =====
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        for (int c = 0; c < p.n_chunks; ++c) {
          const bool tail = c == p.n_chunks - 1 && p.tail_cols != NC;
          const uint32_t bytes = (uint32_t)(tail ? p.tail_cols : NC) * 128u;
          for (int kb = 0; kb < p.KB; ++kb, ++it) {
            const int s = it % STAGES;
            tc::mbar_wait(&b_empty[s], ((it / STAGES) & 1) ^ 1);
            tc::mbar_expect_tx(&b_full[s], bytes);
            tc::tma_load_2d(smem + Smem::B + s * B_STAGE, tail ? &map_tail : &map_full, &b_full[s], kb * BK, c * NC);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t it = 0, acc_it = 0, tile_it = 0;
      const uint32_t a_base = tc::smem_u32(smem + Smem::A);
      for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
        tc::mbar_wait(a_full, tile_it & 1);
        tc::tc_fence_after();
        IS_STAMP(1, 2 * tile_it);
        for (int c = 0; c < p.n_chunks; ++c, ++acc_it) {
          const bool tail = c == p.n_chunks - 1 && p.tail_cols != NC;
          const uint32_t idesc = tc::make_idesc_bf16(BM, tail ? p.tail_cols : NC, 0, 0);
          const uint32_t buf = acc_it & 1;
          tc::mbar_wait(&acc_empty[buf], ((acc_it >> 1) & 1) ^ 1);
          tc::tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * ACC_STRIDE;
          for (int kb = 0; kb < p.KB; ++kb, ++it) {
            const int s = it % STAGES;
            tc::mbar_wait(&b_full[s], (it / STAGES) & 1);
            tc::tc_fence_after();
            const uint32_t a = a_base + kb * A_BLOCK;
            const uint32_t b = tc::smem_u32(smem + Smem::B + s * B_STAGE);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              tc::umma_bf16(d_tmem, tc::desc_kmajor(a, k), tc::desc_kmajor(b, k), idesc, (kb | k) != 0 ? 1u : 0u);
            tc::umma_commit(&b_empty[s]);
          }
          tc::umma_commit(&acc_full[buf]);
        }
        tc::umma_commit(a_empty);        // every MMA that reads this A tile has completed
        IS_STAMP(1, 2 * tile_it + 1);
      }
    }
  } else if (warp >= 4 && warp < 4 + PROD_WARPS) {
    // ===== producers: z, aux, then h = tanh(z.W1 + b1) as the bf16 A tile =====
    const int pw = warp - 4, pt = threadIdx.x - 128;           // pt in [0, 256)
    const int Z = p.Z, H = p.H;
    const int k0 = pw * 64 + 2 * lane;                         // this thread's two hidden units
    float w1a[ZMAX], w1b[ZMAX];
#pragma unroll
    for (int j = 0; j < ZMAX; ++j) {
      w1a[j] = (j < Z && k0 < H) ? p.W1[(size_t)j * H + k0] : 0.f;
      w1b[j] = (j < Z && k0 + 1 < H) ? p.W1[(size_t)j * H + k0 + 1] : 0.f;
    }
    const float b1a = k0 < H ? p.b1[k0] : 0.f, b1b = k0 + 1 < H ? p.b1[k0 + 1] : 0.f;
    uint32_t tile_it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const int pi = t / p.tiles_per_point, l0 = (t - pi * p.tiles_per_point) * BM;
      tc::mbar_wait(a_empty, (tile_it & 1) ^ 1);               // the previous tile's MMAs are done with A (and zs)
      if (pt == 0) IS_STAMP(0, 3 * tile_it);
      // --- z of every sample row: one task = (row, 4 consecutive latent dims) = one Philox call
      float* aux_t = aux_s + (tile_it & 1) * BM;
      {
        constexpr int G = ZMAX / 4;
        const bool quad = (Z & 3) == 0 && !p.eps_inj;          // element groups line up with Philox groups
        for (int i = pt; i < BM * G; i += PROD_WARPS * 32) {
          const int r = i / G, j0 = (i - r * G) * 4;
          const int l = l0 + r;
          float nrm[4] = {0.f, 0.f, 0.f, 0.f};
          if (quad && j0 < Z)
            philox_normal4(p.seed, VAEB_STREAM_IS, 0u, (uint32_t)l, (uint64_t)((p.row_offset + pi) * Z + j0) >> 2, nrm);
          float4 zq;
          float* zp = &zq.x;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int j = j0 + u;
            float zv = 0.f;
            if (j < Z) {
              float e;
              if (p.eps_inj) e = l < p.L ? p.eps_inj[((size_t)pi * p.L + l) * Z + j] : 0.f;
              else if (quad) e = nrm[u];
              else e = philox_normal1(p.seed, VAEB_STREAM_IS, 0u, (uint32_t)l, (uint64_t)((p.row_offset + pi) * Z + j));
              zv = __ldg(p.mu + (size_t)pi * Z + j) + expf(0.5f * __ldg(p.ls + (size_t)pi * Z + j)) * e;
            }
            zp[u] = zv;
          }
          *reinterpret_cast<float4*>(zs + r * ZMAX + j0) = zq;
        }
      }
      named_bar(1, PROD_WARPS * 32);
      if (pt < BM) {                                           // aux[r] = sum_j -z^2/2 + ls/2 + eps^2/2, eps = (z-mu)/sd
        float a = 0.f;
        for (int j = 0; j < Z; ++j) {
          const float lsv = __ldg(p.ls + (size_t)pi * Z + j), zv = zs[pt * ZMAX + j];
          const float e = (zv - __ldg(p.mu + (size_t)pi * Z + j)) * expf(-0.5f * lsv);
          a += -0.5f * zv * zv + 0.5f * lsv + 0.5f * e * e;
        }
        aux_t[pt] = a;
      }
      if (pt == 0) IS_STAMP(0, 3 * tile_it + 1);
      // --- the A tile: warp pw fills k block pw (128 rows x 64 hidden units)
      if (pw < p.KB) {
        uint8_t* ablk = smem + Smem::A + pw * A_BLOCK;
#pragma unroll 2
        for (int r = 0; r < BM; ++r) {
          float ha = b1a, hb = b1b;
          const float4* zr = reinterpret_cast<const float4*>(zs + r * ZMAX);
#pragma unroll
          for (int q = 0; q < ZMAX / 4; ++q) {
            const float4 zq = zr[q];
            ha = fmaf(zq.x, w1a[4 * q], ha); hb = fmaf(zq.x, w1b[4 * q], hb);
            ha = fmaf(zq.y, w1a[4 * q + 1], ha); hb = fmaf(zq.y, w1b[4 * q + 1], hb);
            ha = fmaf(zq.z, w1a[4 * q + 2], ha); hb = fmaf(zq.z, w1b[4 * q + 2], hb);
            ha = fmaf(zq.w, w1a[4 * q + 3], ha); hb = fmaf(zq.w, w1b[4 * q + 3], hb);
          }
          const __nv_bfloat162 hv = __floats2bfloat162_rn(tanh_approx(ha), tanh_approx(hb));
          *reinterpret_cast<__nv_bfloat162*>(ablk + tc::sw128_offset(r, 2 * lane)) = hv;
        }
      }
      tc::fence_proxy_async();                                 // generic-proxy writes -> visible to the MMA
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(a_full);
      if (pt == 0) IS_STAMP(0, 3 * tile_it + 2);
    }
  } else if (warp >= 4 + PROD_WARPS) {
    // ===== epilogue: TMEM -> x*a - softplus(a) row sums -> log w -> tile (max, sum exp) =====
    const int e = warp - 4 - PROD_WARPS, q = warp & 3, ch = e >> 2;   // TMEM lane quarter = warp % 4
    const int et = threadIdx.x - (4 + PROD_WARPS) * 32;                // [0, 256)
    const int row = q * 32 + lane;
    uint32_t acc_it = 0, tile_it = 0;
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x, ++tile_it) {
      const int pi = t / p.tiles_per_point, l0 = (t - pi * p.tiles_per_point) * BM;
      for (int i = et; i < p.D; i += EPI_WARPS * 32) xb[i].y = p.x[(size_t)pi * p.D + i];
      named_bar(2, EPI_WARPS * 32);
      const uint32_t xb_addr = tc::smem_u32(xb);
      float rs0 = 0.f, rs1 = 0.f;
      for (int c = 0; c < p.n_chunks; ++c, ++acc_it) {
        const bool tail = c == p.n_chunks - 1 && p.tail_cols != NC;
        const int width = tail ? p.tail_cols : NC;
        const uint32_t buf = acc_it & 1;
        tc::mbar_wait(&acc_full[buf], (acc_it >> 1) & 1);
        tc::tc_fence_after();
        if (et == 0) IS_STAMP(2, 2 * acc_it);
        const uint32_t taddr = tmem_base + buf * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
        // this warp's half of the chunk, 16 columns per TMEM load; the next load is in flight while
        // the current 16 columns are folded
        const int half = (width / 2 + 15) & ~15;
        const int c_lo = ch * half, c_hi = min(width, c_lo + half);
        float va[16], vb[16];
        if (c_lo < c_hi) tc::tmem_ld16(taddr + (uint32_t)c_lo, va);
        for (int cc = c_lo; cc < c_hi; cc += 32) {
          tc::tmem_ld_wait();
          if (cc + 16 < c_hi) tc::tmem_ld16(taddr + (uint32_t)(cc + 16), vb);
          fold16(va, xb_addr + (uint32_t)(c * NC + cc) * 8u, rs0, rs1);
          if (cc + 16 < c_hi) {
            tc::tmem_ld_wait();
            if (cc + 32 < c_hi) tc::tmem_ld16(taddr + (uint32_t)(cc + 32), va);
            fold16(vb, xb_addr + (uint32_t)(c * NC + cc + 16) * 8u, rs0, rs1);
          }
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&acc_empty[buf]);
        if (et == 0) IS_STAMP(2, 2 * acc_it + 1);
      }
      const float rs = rs0 + rs1;
      rowsum_s[ch * BM + row] = rs;
      named_bar(2, EPI_WARPS * 32);
      if (ch == 0) {
        const int l = l0 + row;
        const bool valid = l < p.L;
        const float lw = rowsum_s[row] + rowsum_s[BM + row] + aux_s[(tile_it & 1) * BM + row];
        if (valid && p.logw_out) p.logw_out[(size_t)pi * p.L + l] = lw;
        float m = valid ? lw : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float se = valid ? expf(lw - m) : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        if (lane == 0) { red_s[2 * q] = m; red_s[2 * q + 1] = se; }
        named_bar(3, 4 * 32);
        if (et == 0) {
          float M = -INFINITY;
          for (int w = 0; w < 4; ++w) M = fmaxf(M, red_s[2 * w]);
          float S = 0.f;
          for (int w = 0; w < 4; ++w) if (red_s[2 * w] > -INFINITY) S += red_s[2 * w + 1] * expf(red_s[2 * w] - M);
          p.partial[t] = make_float2(M, S);
        }
        named_bar(3, 4 * 32);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) tc::tmem_dealloc(tmem_base, TMEM_COLS);
}

// log p^(x_i) = log sum_t S_t exp(M_t - M) + M - log L   over the tiles of point i
__global__ void is_tc_finish_kernel(const float2* __restrict__ partial, int n_points, int tiles_per_point, int L,
                                    float* __restrict__ logp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  float M = -INFINITY;
  for (int t = 0; t < tiles_per_point; ++t) M = fmaxf(M, partial[(size_t)i * tiles_per_point + t].x);
  float S = 0.f;
  for (int t = 0; t < tiles_per_point; ++t) {
    const float2 v = partial[(size_t)i * tiles_per_point + t];
    if (v.x > -INFINITY) S += v.y * expf(v.x - M);
  }
  logp[i] = M + logf(S) - logf((float)L);
}

// W2 [H, D] fp32 -> W2^T [D, KP] bf16 (k contiguous, zero padded to KP = 64*KB)
__global__ void is_tc_prep_kernel(const float* __restrict__ W2, int H, int D, int KP, __nv_bfloat16* __restrict__ w2t) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)D * KP) return;
  const int n = (int)(i / KP), k = (int)(i % KP);
  w2t[i] = __float2bfloat16_rn(k < H ? W2[(size_t)k * D + n] : 0.f);
}

}  // namespace istc

bool is_tc_supported(const vaeb_handle* h) {
  return h->cfg.precision == VAEB_PREC_BF16 && !h->cont && h->H <= 64 * istc::KB_MAX && h->Z <= istc::ZMAX &&
         h->D <= istc::DMAX && h->D >= 16;
}

// mu, ls: device [n, Z] (fp32 encoder already run); d_x device [n, D]; logp_out device [n]; logw_out device or nullptr
int is_tc_run(vaeb_handle* h, const float* d_x, const float* d_mu, const float* d_ls, int n, int L, const float* d_eps,
              int64_t row_offset, float* d_logp, float* d_logw) {
  using namespace istc;
  const Layout& l = h->lay;
  const int D = h->D, H = h->H, Z = h->Z;
  const int KB = (H + 63) / 64, KP = KB * 64;
  IsTcState& s = h->istc;
  cudaStream_t st = h->stream;
  if (!s.w2t) {
    VAEB_CUDA(cudaMalloc(&s.w2t, (size_t)D * KP * 2));
    VAEB_CUDA(cudaFuncSetAttribute(is_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::TOTAL));
    VAEB_CUDA(cudaDeviceGetAttribute(&s.n_sm, cudaDevAttrMultiProcessorCount, h->cfg.device));
    const int n_chunks = (D + NC - 1) / NC;
    int tail = D - (n_chunks - 1) * NC;
    tail = (tail + 15) & ~15;
    s.n_chunks = n_chunks; s.tail_cols = tail;
    VAEB_TRY(vaeb_make_tmap_bf16((CUtensorMap*)s.map_full, s.w2t, (uint64_t)D, (uint64_t)KP, (uint64_t)KP, NC));
    VAEB_TRY(vaeb_make_tmap_bf16((CUtensorMap*)s.map_tail, s.w2t, (uint64_t)D, (uint64_t)KP, (uint64_t)KP,
                                 (uint32_t)tail));
  }
  // the weights may have changed since the last call: refresh the bf16 transpose (0.8 MB)
  {
    const int64_t tot = (int64_t)D * KP;
    is_tc_prep_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(h->d_params + l.off[l.iW2], H, D, KP,
                                                                    (__nv_bfloat16*)s.w2t);
    ++h->launches;
    VAEB_CUDA(cudaGetLastError());
  }
  const int tpp = (L + BM - 1) / BM;
  const int64_t n_tiles = (int64_t)n * tpp;
  VAEB_REQUIRE(n_tiles < ((int64_t)1 << 31), "too many importance-sampling tiles in one call");
  if (n_tiles > s.partial_cap) {
    VAEB_CUDA(cudaStreamSynchronize(st));
    if (s.partial) VAEB_CUDA(cudaFree(s.partial));
    s.partial = nullptr;
    VAEB_CUDA(cudaMalloc(&s.partial, (size_t)n_tiles * sizeof(float2)));
    s.partial_cap = n_tiles;
  }
  Params p{};
  p.n_points = n; p.L = L; p.D = D; p.H = H; p.Z = Z; p.KB = KB;
  p.tiles_per_point = tpp; p.n_chunks = s.n_chunks; p.tail_cols = s.tail_cols;
  p.x = d_x; p.mu = d_mu; p.ls = d_ls;
  p.W1 = h->d_params + l.off[l.iW1]; p.b1 = h->d_params + l.off[l.ib1]; p.b2 = h->d_params + l.off[l.ib2];
  p.eps_inj = d_eps; p.seed = h->cfg.seed; p.row_offset = row_offset;
  p.partial = (float2*)s.partial; p.logw_out = d_logw;
  const int grid = (int)std::min<int64_t>(n_tiles, s.n_sm);
  is_tc_kernel<<<grid, THREADS, Smem::TOTAL, st>>>(*(const CUtensorMap*)s.map_full, *(const CUtensorMap*)s.map_tail, p);
  ++h->launches;
  VAEB_CUDA(cudaGetLastError());
  is_tc_finish_kernel<<<(n + 127) / 128, 128, 0, st>>>((const float2*)s.partial, n, tpp, L, d_logp);
  ++h->launches;
  VAEB_CUDA(cudaGetLastError());
  return VAEB_OK;
}

extern "C" int vaeb_is_tc_debug(long long* d_buf) {
  cudaMemcpyToSymbol(istc::g_is_dbg, &d_buf, sizeof(d_buf));
  return 0;
}
