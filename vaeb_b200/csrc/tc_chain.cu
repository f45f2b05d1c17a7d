// The activation chain of the large-batch AEVB step (config C3) as ONE persistent tcgen05 kernel.
//   enc1 -> enc2 (+ reparameterisation, KL) -> dec1 -> dec2 (+ Bernoulli log-lik, delta) -> dgrad h_d -> dz (+ dmu, dls)
//   -> dgrad h_e                                                     VAEB.py:245-265 (forward), :396-399 (T.grad)
// Every one of these seven layers is `out[rows, N] = epilogue(A[rows, K] . B)` with a K-major activation operand, and
// row block r of a layer needs only row block r of the layer before it.  So the layers are not separate launches
// (17 per update in round 1, each paying launch + prologue + the TMA -> MMA -> commit -> TMEM-read chain + a tail)
// but ITEMS of one launch: item = (layer, 128-row block, column tile), listed layer by layer, dealt round-robin to one
// CTA per SM.  A CTA keeps its operand ring, its two TMEM accumulators and its warp roles (TMA producer, MMA issuer,
// 16 epilogue warps) across items and layers; the only synchronisation between layers is a per-(layer, row block)
// arrival counter in global memory: the epilogue warps of a tile `red.release` it after their stores, the TMA
// producer of a consuming item `ld.acquire`s it before it loads the A operand (fence.proxy.async on both sides: the
// mirrors are written through the generic proxy and read through the async proxy).  With >= 148 items per layer the
// producers of an item finished a wave earlier, so nobody waits; with few row blocks (2048 rows per GPU in the 8-GPU
// data-parallel split) the wait is one item, not a launch.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_epilogues.cuh"
#include "tc_layers.h"

namespace {

constexpr int CH_MAX_LAYERS = 7;
// warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator + publisher, 3..18 epilogue (19 warps: 104 registers per thread)
constexpr int CH_THREADS = 96 + EPI_WARPS * 32;
enum ChainKind { CK_TANH = 0, CK_HEADS = 1, CK_BERN = 2, CK_DGRAD = 3, CK_DZ = 4 };

struct alignas(64) ChainLayer {
  LayerMaps maps;
  int M, N, K;            // rows, output columns, contraction extent
  int bn;                 // UMMA N of this layer (64 / 128 / 256)
  int b_mn;               // B operand MN-major (3-D box) or K-major (2-D box)
  int a_row_off;          // row offset of the A operand (resident data set)
  int tiles_n;            // column tiles per row block
  int first_item;         // index of this layer's first item in the launch-wide item list
  int kind, epi;          // epilogue kind and index into the epilogue table of that kind
  int dep;                // layer whose row block must be complete before A is loaded (-1: none)
  int dep_count;          // arrivals per row block of that layer (one per column tile, by the publisher warp)
};

struct ChainArgs {
  ChainLayer layer[CH_MAX_LAYERS];
  EpiTanh tanh_[2];             // enc1, dec1
  EpiHeads heads;               // enc2
  EpiBernoulliTc bern;          // dec2
  EpiDgradTanh dgrad[2];        // dgrad h_d, dgrad h_e
  EpiDzPrep dz;
  int n_layers, n_items;
  unsigned int* ready;          // [n_layers][ready_stride] arrival counters, monotonic over launches
  int ready_stride;
  unsigned int epoch;           // this launch's number (1, 2, ...): a row block is complete at epoch * dep_count
  int stages, stage_bytes;      // operand ring of this launch
  long long* stamps;            // debug (VAEB_CHAIN_STAMPS): per item {dependency met, accumulator complete, stores done, SM}
  int throttle;                 // measurement switch (VAEB_CHAIN_THROTTLE): the MMA thread waits for a stage's MMAs before the next stage
  int ld32;                     // measurement switch (VAEB_CHAIN_LD32): 32 accumulator columns per TMEM read in the epilogue
  int diag;                     // debug (VAEB_CHAIN_DIAG): the MMA thread also waits for each item's accumulator and stamps it
};

// Operand ring: a stage holds NS x (A 128 x 64 | B bn_max x 64) bf16; the ring takes 192 KB whatever bn_max is, so a launch
// with narrower tiles (few row blocks: more, smaller items) gets a deeper ring -- with K = 784 an item is 13 stages and
// its latency is the number of L2 round trips: 13 / depth.
constexpr int CH_A_BYTES = BM * BK * 2;                 // 16 KB: 128 rows x 64 bf16
constexpr int CH_RING_BYTES = 192 * 1024;
constexpr int CH_MAX_STAGES = 8;
constexpr int CH_SMEM_TOTAL = CH_RING_BYTES + 1024 + 256;
inline int chain_stage_bytes(int ns, int bn_max) { return ns * (CH_A_BYTES + bn_max * BK * 2); }
inline int chain_stages(int ns, int bn_max) {
  const int s = CH_RING_BYTES / chain_stage_bytes(ns, bn_max);
  return s > CH_MAX_STAGES ? CH_MAX_STAGES : s;
}

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long gtimer() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async;" ::: "memory"); }

// ---- CTA pair (cta_group::2): one UMMA of M = 256 spans two SMs of a TPC; each CTA stages its own 128 rows of A and HALF of
// the B tile, so a k block costs a CTA 32 KB x NS instead of 48 KB x NS (L2 -> shared-memory traffic and ring depth both
// gain: 3 stages instead of 2 in bf16x3).  The even CTA of the pair issues the MMAs; TMA loads of both CTAs complete on
// ITS full barrier; its commits are multicast to the barriers of both.
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_slot, uint32_t ncols) {   // one full warp, in both CTAs
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::smem_u32(smem_slot)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_2(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(tc::smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the even CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(m), "r"(tc::smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(tc::smem_u32(smem_dst)), "l"(m), "r"(tc::smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_even_cta(uint64_t* bar) {      // the barrier of the pair's even CTA
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(tc::smem_u32(bar) & PEER_BIT_MASK) : "memory");
}

__device__ __forceinline__ const ChainLayer& layer_of(const ChainArgs& a, int g) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < CH_MAX_LAYERS; ++i)
    if (i < a.n_layers && g >= a.layer[i].first_item) l = i;
  return a.layer[l];
}

// One tile's epilogue for one warp: lane quarter q, column slice cs of BN / 4 columns (the body of the persistent
// layer kernel's epilogue, tc_layers.cu).  Returns after the warp's last store.
template <class Epi, int BN, bool PAIR = false>
__device__ __forceinline__ void chain_epi_tile(Epi& epi, uint32_t acc, uint64_t* tmem_full, uint64_t* tmem_empty,
                                               uint32_t use, int tm, int tn, int tiles_n, int M, int N, int q, int cs,
                                               int lane) {
  constexpr int SLICE = BN / (EPI_WARPS / 4);
  constexpr int NCH = SLICE / 16;
  const int n0 = tn * BN;
  const int row = tm * BM + q * 32 + lane;
  const bool ok = row < M;
  epi.begin();
  // (no global read before the accumulator is complete: until then the row block this item depends on -- the operand
  // the epilogue re-reads was written by another CTA in this very launch -- may still be in flight)
  tc::mbar_wait(tmem_full, use & 1);
  tc::tc_fence_after();
  // the TMEM read of chunk i+1 and (PREFETCH) its global operand are in flight while chunk i is finished
  float v[2][16];
  uint32_t pw[2][8];
  bool ph[2] = {false, false};
  tc::tmem_ld16(acc + (uint32_t)(cs * SLICE), v[0]);
  if constexpr (Epi::PREFETCH) ph[0] = (n0 + cs * SLICE < N) && epi.preload(row, ok, n0 + cs * SLICE, N, pw[0]);
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const int c = cs * SLICE + 16 * i;
    tc::tmem_ld_wait();                               // chunk i is in registers
    if (i + 1 < NCH) {
      tc::tmem_ld16(acc + (uint32_t)(c + 16), v[(i + 1) & 1]);
      if constexpr (Epi::PREFETCH)
        ph[(i + 1) & 1] = (n0 + c + 16 < N) && epi.preload(row, ok, n0 + c + 16, N, pw[(i + 1) & 1]);
    } else {                                          // last read of this warp: hand the accumulator back early
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_even_cta(tmem_empty); else tc::mbar_arrive(tmem_empty);
      }
    }
    if (n0 + c < N) {
      if constexpr (Epi::PREFETCH) epi.chunk(row, ok, n0 + c, N, v[i & 1], ph[i & 1] ? pw[i & 1] : nullptr);
      else epi.chunk(row, ok, n0 + c, N, v[i & 1]);
    }
  }
  epi.end(row, ok, tn * (EPI_WARPS / 4) + cs, tiles_n * (EPI_WARPS / 4));
}

// the same with 32 accumulator columns per TMEM read (BN >= 128): half as many tcgen05.ld per tile
template <class Epi, int BN, bool PAIR = false>
__device__ __forceinline__ void chain_epi_tile32(Epi& epi, uint32_t acc, uint64_t* tmem_full, uint64_t* tmem_empty,
                                                 uint32_t use, int tm, int tn, int tiles_n, int M, int N, int q, int cs,
                                                 int lane) {
  constexpr int SLICE = BN / (EPI_WARPS / 4);
  constexpr int NCH = SLICE / 32;
  const int n0 = tn * BN;
  const int row = tm * BM + q * 32 + lane;
  const bool ok = row < M;
  epi.begin();
  tc::mbar_wait(tmem_full, use & 1);
  tc::tc_fence_after();
#pragma unroll 1
  for (int i = 0; i < NCH; ++i) {
    const int c = cs * SLICE + 32 * i;
    float v[32];
    tc::tmem_ld32(acc + (uint32_t)c, v);
    tc::tmem_ld_wait();
    if (i == NCH - 1) {
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_even_cta(tmem_empty); else tc::mbar_arrive(tmem_empty);
      }
    }
    if (n0 + c < N) epi.chunk(row, ok, n0 + c, N, v);
    if (n0 + c + 16 < N) epi.chunk(row, ok, n0 + c + 16, N, v + 16);
  }
  epi.end(row, ok, tn * (EPI_WARPS / 4) + cs, tiles_n * (EPI_WARPS / 4));
}

template <bool PAIR, class Epi>
__device__ __forceinline__ void chain_epi(Epi epi, int bn, uint32_t acc, uint64_t* tmem_full, uint64_t* tmem_empty,
                                          uint32_t use, int tm, int tn, int tiles_n, int M, int N, int q, int cs,
                                          int lane) {
  if (bn < 0)                 // (measurement switch: 32 columns per TMEM read, 256-wide tiles)
    chain_epi_tile32<Epi, 256, PAIR>(epi, acc, tmem_full, tmem_empty, use, tm, tn, tiles_n, M, N, q, cs, lane);
  else if (PAIR || bn == 256)      // the pair form runs its wide layers 256 wide only
    chain_epi_tile<Epi, 256, PAIR>(epi, acc, tmem_full, tmem_empty, use, tm, tn, tiles_n, M, N, q, cs, lane);
  else if (bn == 128) chain_epi_tile<Epi, 128, PAIR>(epi, acc, tmem_full, tmem_empty, use, tm, tn, tiles_n, M, N, q, cs, lane);
  else chain_epi_tile<Epi, 64, PAIR>(epi, acc, tmem_full, tmem_empty, use, tm, tn, tiles_n, M, N, q, cs, lane);
}

template <int NS, bool PAIR>
__global__ void __launch_bounds__(CH_THREADS, 1)
tc_chain_kernel(const __grid_constant__ ChainArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(smem + CH_RING_BYTES);
  uint64_t* empty = full + CH_MAX_STAGES;
  uint64_t* tmem_full = empty + CH_MAX_STAGES;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]
  uint32_t* tmem_slot = (uint32_t*)(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t ACC_COLS = 256, TMEM_COLS = 512;
  // PAIR: the two CTAs of a cluster walk the same item list; an item covers the row blocks 2 * tp + {0, 1}
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int n_workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < args.n_layers; ++l) {
      tc::tma_prefetch_desc(&args.layer[l].maps.a_hi);
      tc::tma_prefetch_desc(&args.layer[l].maps.b_hi);
      if (NS == 2) { tc::tma_prefetch_desc(&args.layer[l].maps.a_lo); tc::tma_prefetch_desc(&args.layer[l].maps.b_lo); }
    }
    for (int s = 0; s < CH_MAX_STAGES; ++s) { tc::mbar_init(&full[s], 1); tc::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&tmem_full[b], 1);
      tc::mbar_init(&tmem_empty[b], PAIR ? 2 * EPI_WARPS : EPI_WARPS);   // PAIR: the epilogue warps of both CTAs (even CTA's barrier)
    }
    tc::fence_barrier_init();
  }
  if (PAIR) cluster_sync_all();                      // the peer's barriers exist before anything is sent to them
  if (warp == 2) { if (PAIR) tmem_alloc2(tmem_slot, TMEM_COLS); else tc::tmem_alloc(tmem_slot, TMEM_COLS); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    const int stages = args.stages;
    const uint32_t stage_bytes = (uint32_t)args.stage_bytes;
    int s = 0;                                         // ring slot and its phase, carried across items
    uint32_t ph = 0;
    for (int g = worker; g < args.n_items; g += n_workers) {
      const ChainLayer& L = layer_of(args, g);
      const int tile = g - L.first_item;
      const int tp = tile / L.tiles_n, tn = tile - tp * L.tiles_n;
      const int tm = PAIR ? 2 * tp + rank : tp;
      // PAIR: this CTA stages its half of the B tile (64-wide layers: the odd CTA's half lies outside the matrix -> zeros)
      const int bh = PAIR ? (L.bn == 64 ? 64 : L.bn / 2) : L.bn;
      const int m0 = tm * BM, n0 = tn * L.bn + rank * bh;
      const int nkb = (L.K + BK - 1) / BK;
      const uint32_t b_bytes = (uint32_t)bh * (BK * 2);
      const uint32_t stage_tx = (PAIR ? 2 : 1) * NS * (CH_A_BYTES + b_bytes);
      const int b_mn = L.b_mn, arow = L.a_row_off + m0;
      const CUtensorMap* ta_hi = &L.maps.a_hi;
      const CUtensorMap* ta_lo = &L.maps.a_lo;
      const CUtensorMap* tb_hi = &L.maps.b_hi;
      const CUtensorMap* tb_lo = &L.maps.b_lo;
      if (L.dep >= 0 && m0 < L.M) {
        // the row block of the producing layer: every column tile has been stored and published
        const unsigned int* flag = args.ready + (size_t)L.dep * args.ready_stride + tm;
        const unsigned int target = args.epoch * (unsigned int)L.dep_count;
        uint32_t spins = 0;
        while ((int)(ld_acquire_u32(flag) - target) < 0) {
          if (++spins > 200000000u) __trap();          // a wrong dependency must trap, not hang the GPU
        }
        fence_proxy_async_global();                    // generic-proxy stores (acquired above) -> this thread's TMA reads
      }
      if (args.stamps && m0 < L.M) args.stamps[8 * (size_t)(PAIR ? 2 * g + rank : g)] = gtimer();
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait(&empty[s], ph ^ 1);
        uint8_t* base = smem + (uint32_t)s * stage_bytes;
        uint8_t* b = base + NS * CH_A_BYTES;
        if constexpr (PAIR) {
          if (rank == 0) tc::mbar_expect_tx(&full[s], stage_tx);      // the bytes of both CTAs land on the even CTA's barrier
          tma_load_2d_pair(base, ta_hi, &full[s], kb * BK, arow);
          if (NS == 2) tma_load_2d_pair(base + CH_A_BYTES, ta_lo, &full[s], kb * BK, arow);
          if (b_mn) {
            tma_load_3d_pair(b, tb_hi, &full[s], 0, kb * BK, n0 / 64);
            if (NS == 2) tma_load_3d_pair(b + b_bytes, tb_lo, &full[s], 0, kb * BK, n0 / 64);
          } else {
            tma_load_2d_pair(b, tb_hi, &full[s], kb * BK, n0);
            if (NS == 2) tma_load_2d_pair(b + b_bytes, tb_lo, &full[s], kb * BK, n0);
          }
        } else {
          tc::mbar_expect_tx(&full[s], stage_tx);
          tc::tma_load_2d(base, ta_hi, &full[s], kb * BK, arow);
          if (NS == 2) tc::tma_load_2d(base + CH_A_BYTES, ta_lo, &full[s], kb * BK, arow);
          if (b_mn) {                                    // one box: bn / 64 column groups
            tc::tma_load_3d(b, tb_hi, &full[s], 0, kb * BK, n0 / 64);
            if (NS == 2) tc::tma_load_3d(b + b_bytes, tb_lo, &full[s], 0, kb * BK, n0 / 64);
          } else {
            tc::tma_load_2d(b, tb_hi, &full[s], kb * BK, n0);
            if (NS == 2) tc::tma_load_2d(b + b_bytes, tb_lo, &full[s], kb * BK, n0);
          }
        }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1 && lane == 0 && rank == 0) {
    // ===== MMA issuer (PAIR: the even CTA issues for both) =====
    // One thread issues every MMA of the CTA (~8 cycles per instruction): the shared-memory descriptors of a stage are
    // (constant bits | address >> 4), so a k step is one add per operand, whatever the layer's layouts are.
    const int stages = args.stages;
    const uint32_t stage_bytes = (uint32_t)args.stage_bytes;
    int s = 0;
    uint32_t ph = 0, lt = 0;                           // ring slot / phase; items done by this CTA
    const uint32_t smem0 = tc::smem_u32(smem);
    for (int g = worker; g < args.n_items; g += n_workers, ++lt) {
      const ChainLayer& L = layer_of(args, g);
      const int nkb = (L.K + BK - 1) / BK;
      const int bh = PAIR ? (L.bn == 64 ? 64 : L.bn / 2) : L.bn;
      const uint32_t b_bytes = (uint32_t)bh * (BK * 2);
      const uint32_t idesc = tc::make_idesc_bf16(PAIR ? 2 * BM : BM, PAIR ? 2 * bh : bh, 0, L.b_mn);
      // descriptor bits without the address: K-major (LBO 16 B, SBO 1024 B), MN-major (LBO = 8192 B between 64-wide groups)
      const uint64_t a_bits = tc::make_smem_desc(0u, 16u, 1024u);
      const uint64_t b_bits = L.b_mn ? tc::make_smem_desc(0u, 8192u, 1024u) : a_bits;
      const uint64_t b_step = L.b_mn ? (2048u >> 4) : (32u >> 4);     // one UMMA_K = 16 slice further, in 16-byte units
      const uint32_t buf = lt & 1, use = lt >> 1;
      tc::mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);           // the epilogue drained this accumulator
      tc::tc_fence_after();
      const uint32_t acc = tmem_base + buf * ACC_COLS;
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        if (kb == 0 && args.stamps) args.stamps[8 * (size_t)(PAIR ? 2 * g + rank : g) + 6] = gtimer();
        const uint32_t a = smem0 + (uint32_t)s * stage_bytes;
        const uint32_t b = a + NS * CH_A_BYTES;
        uint64_t dah = a_bits | (uint64_t)((a & 0x3FFFFu) >> 4), dbh = b_bits | (uint64_t)((b & 0x3FFFFu) >> 4);
        uint64_t dal = a_bits | (uint64_t)(((a + CH_A_BYTES) & 0x3FFFFu) >> 4);
        uint64_t dbl = b_bits | (uint64_t)(((b + b_bytes) & 0x3FFFFu) >> 4);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          if constexpr (PAIR) {
            if (k == 0) umma_bf16_2(acc, dah, dbh, idesc, kb != 0 ? 1u : 0u);
            else umma_bf16_2(acc, dah, dbh, idesc, 1u);
            if (NS == 2) {
              umma_bf16_2(acc, dah, dbl, idesc, 1u);
              umma_bf16_2(acc, dal, dbh, idesc, 1u);
            }
          } else {
            if (k == 0) tc::umma_bf16(acc, dah, dbh, idesc, kb != 0 ? 1u : 0u);
            else tc::umma_bf16(acc, dah, dbh, idesc, 1u);
            if (NS == 2) {
              tc::umma_bf16(acc, dah, dbl, idesc, 1u);
              tc::umma_bf16(acc, dal, dbh, idesc, 1u);
            }
          }
          if (NS == 2) { dal += 2; dbl += b_step; }
          dah += 2; dbh += b_step;
        }
        if constexpr (PAIR) umma_commit_2(&empty[s]); else tc::umma_commit(&empty[s]);
        if (args.throttle) tc::mbar_wait(&empty[s], ph);      // (measurement) at most one stage of MMAs in the tensor pipe's queue
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      if constexpr (PAIR) umma_commit_2(&tmem_full[buf]); else tc::umma_commit(&tmem_full[buf]);
      if (args.stamps && args.diag) {                  // debug: when this item's MMAs really completed
        tc::mbar_wait(&tmem_full[buf], use & 1);
        args.stamps[8 * (size_t)(PAIR ? 2 * g + rank : g) + 7] = gtimer();
      }
    }
  } else if (warp >= 3) {
    // ===== epilogue: warp w reads TMEM lane quarter w % 4 (the hardware's rule), column slice (w - 3) / 4 of the tile =====
    const int q = warp & 3, cs = (warp - 3) >> 2;
    uint32_t lt = 0;
    for (int g = worker; g < args.n_items; g += n_workers, ++lt) {
      const ChainLayer& L = layer_of(args, g);
      const int tile = g - L.first_item;
      const int tp = tile / L.tiles_n, tn = tile - tp * L.tiles_n;
      const int tm = PAIR ? 2 * tp + rank : tp;
      const uint32_t buf = lt & 1, use = lt >> 1;
      const uint32_t acc = tmem_base + buf * ACC_COLS + ((uint32_t)(q * 32) << 16);
      uint64_t* tf = &tmem_full[buf];
      uint64_t* te = &tmem_empty[buf];
      const size_t sg = 8 * (size_t)(PAIR ? 2 * g + rank : g);
      if (args.stamps && warp == 4 && lane == 0) {     // debug: when this item's accumulator was complete
        args.stamps[sg + 4] = gtimer();
        tc::mbar_wait(tf, use & 1);
        args.stamps[sg + 1] = gtimer();
      }
      switch (L.kind) {
        case CK_TANH:
          chain_epi<PAIR>(args.tanh_[L.epi], (args.ld32 && L.bn == 256) ? -256 : L.bn, acc, tf, te, use, tm, tn, L.tiles_n, L.M, L.N, q, cs, lane);
          break;
        case CK_HEADS: {                                 // enc2 and dz are always 64 wide
          EpiHeads e = args.heads;
          chain_epi_tile<EpiHeads, 64, PAIR>(e, acc, tf, te, use, tm, tn, L.tiles_n, L.M, L.N, q, cs, lane);
          break;
        }
        case CK_BERN:
          chain_epi<PAIR>(args.bern, (args.ld32 && L.bn == 256) ? -256 : L.bn, acc, tf, te, use, tm, tn, L.tiles_n, L.M, L.N, q, cs, lane);
          break;
        case CK_DGRAD:
          chain_epi<PAIR>(args.dgrad[L.epi], (args.ld32 && L.bn == 256) ? -256 : L.bn, acc, tf, te, use, tm, tn, L.tiles_n, L.M, L.N, q, cs, lane);
          break;
        default: {
          EpiDzPrep e = args.dz;
          chain_epi_tile<EpiDzPrep, 64, PAIR>(e, acc, tf, te, use, tm, tn, L.tiles_n, L.M, L.N, q, cs, lane);
          break;
        }
      }
      // this warp's stores of the tile are issued: tell the publisher warp and go on with the next tile (the fences that
      // make the stores visible to other CTAs wait for the stores to drain -- ~2 us per tile when every epilogue warp
      // paid for them itself)
      if (args.stamps && warp == 4 && lane == 0) { args.stamps[sg + 2] = gtimer(); args.stamps[sg + 3] = blockIdx.x; }
      asm volatile("barrier.cta.arrive %0, %1;" ::"r"(1u + (lt & 3u)), "r"((uint32_t)(EPI_WARPS * 32 + 32)) : "memory");
    }
  } else if (warp == 2) {
    // ===== publisher: a tile is complete when its 16 epilogue warps have arrived; one gpu-scope fence + one release
    // per tile then cover the stores of all of them (the cumulativity a grid barrier relies on) =====
    uint32_t lt = 0;
    for (int g = worker; g < args.n_items; g += n_workers, ++lt) {
      const ChainLayer& L = layer_of(args, g);
      const int tile = g - L.first_item;
      const int tp = tile / L.tiles_n;
      const int tm = PAIR ? 2 * tp + rank : tp;        // (a row block past the end lands in the spare counter slot)
      asm volatile("barrier.cta.sync %0, %1;" ::"r"(1u + (lt & 3u)), "r"((uint32_t)(EPI_WARPS * 32 + 32)) : "memory");
      __threadfence();
      fence_proxy_async_global();                      // generic-proxy stores -> ordered before other CTAs' TMA reads
      if (lane == 0) {
        red_release_add(args.ready + (size_t)(&L - args.layer) * args.ready_stride + tm, 1u);
        if (args.stamps) args.stamps[8 * (size_t)(PAIR ? 2 * g + rank : g) + 5] = gtimer();
      }
      __syncwarp();
    }
  }
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  tc::tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();                      // the even CTA's MMAs read the odd CTA's shared memory: leave together
  if (warp == 2) { if (PAIR) tmem_dealloc2(tmem_base, TMEM_COLS); else tc::tmem_dealloc(tmem_base, TMEM_COLS); }
}

void fill_layer(ChainLayer& l, const unsigned char* maps, int M, int N, int K, int bn, int b_mn, int a_row_off, int kind,
                int epi, int dep) {
  l.maps = *reinterpret_cast<const LayerMaps*>(maps);
  l.M = M; l.N = N; l.K = K; l.bn = bn; l.b_mn = b_mn; l.a_row_off = a_row_off;
  l.tiles_n = (N + bn - 1) / bn;
  l.kind = kind; l.epi = epi; l.dep = dep; l.dep_count = 0; l.first_item = 0;
}

}  // namespace

int tc_chain_ready_elems(int rows) { return CH_MAX_LAYERS * ((rows + BM - 1) / BM + 1); }

// One launch: the seven activation layers of forward + backward for `rows` rows (L = 1).  `bn` is the UMMA N of the
// wide layers (tc_build_maps built the maps for it); enc2 and dz are 64 wide.  `ready` holds tc_chain_ready_elems(rows)
// counters, zeroed whenever the shapes change; `epoch` is the number of launches since then (1, 2, ...).
cudaError_t tc_chain_step(cudaStream_t st, int64_t* launches, const TcMaps& m, int ns, int bn, int rows, int D, int H,
                          int Z, int la, int x_row_off, const float* b3, const float* b4, const float* b5, const float* b1,
                          const float* b2, const EpsSource& src, float scale, float w, float* mu, float* ls, float* eps,
                          float* z, float* dmu, float* dls, float* aux_part, int* n_aux, float* partial, int* n_tiles,
                          const TcBuffers& b, const void* xm_hi, const void* xm_lo, unsigned int* ready,
                          unsigned int epoch, int n_sm, int pair) {
  ChainArgs a;              // ~5 KB: filled per call, passed by value
  const int fast = ns == 1 ? 1 : 0;
  const int tiles_m = (rows + BM - 1) / BM;
  __nv_bfloat16 *heh = (__nv_bfloat16*)b.heh, *hel = (__nv_bfloat16*)b.hel, *hdh = (__nv_bfloat16*)b.hdh,
                *hdl = (__nv_bfloat16*)b.hdl, *d1h = (__nv_bfloat16*)b.d1h, *d1l = (__nv_bfloat16*)b.d1l,
                *da3h = (__nv_bfloat16*)b.da3h, *da3l = (__nv_bfloat16*)b.da3l;
  // layer 0 enc1: h_e = tanh(x.W3 + b3)                                   VAEB.py:246
  fill_layer(a.layer[0], m.enc1, rows, H, D, bn, 1, x_row_off, CK_TANH, 0, -1);
  static const int epi_dbg = getenv("VAEB_EPI_DBG") ? atoi(getenv("VAEB_EPI_DBG")) : 0;     // measurement switch
  a.tanh_[0] = EpiTanh{b3, nullptr, H, heh, hel, b.ldh, fast, epi_dbg};
  // layer 1 enc2: (mu, ls) = h_e.[W4|W5] + b, eps, z, KL / L^A row term    VAEB.py:248-249, 41-47, 343, 322-325
  fill_layer(a.layer[1], m.enc2, rows, 2 * Z, H, 64, 1, 0, CK_HEADS, 0, 0);
  a.heads = EpiHeads{b4, b5, Z, la, src, mu, ls, eps, z, (__nv_bfloat16*)b.zh, (__nv_bfloat16*)b.zl, b.ldz, aux_part, 0.f};
  *n_aux = a.layer[1].tiles_n * (EPI_WARPS / 4);
  // layer 2 dec1: h_d = tanh(z.W1 + b1)                                   VAEB.py:254
  fill_layer(a.layer[2], m.dec1, rows, H, Z, bn, 1, 0, CK_TANH, 1, 1);
  a.tanh_[1] = EpiTanh{b1, nullptr, H, hdh, hdl, b.ldh, fast, epi_dbg};
  // layer 3 dec2: a = h_d.W2 + b2, Bernoulli log-likelihood, delta        VAEB.py:263, 311
  fill_layer(a.layer[3], m.dec2, rows, D, H, bn, 1, 0, CK_BERN, 0, 2);
  a.bern = EpiBernoulliTc{b2, nullptr, D, 1, rows, scale, (__nv_bfloat16*)b.da2h, (__nv_bfloat16*)b.da2l, b.ldd, partial,
                          (const __nv_bfloat16*)xm_hi, (const __nv_bfloat16*)xm_lo, b.ldx, x_row_off, 0.f};
  *n_tiles = a.layer[3].tiles_n * (EPI_WARPS / 4);
  // layer 4 dgrad h_d: da1 = (da2.W2^T) * (1 - h_d^2)                      T.grad, VAEB.py:397
  fill_layer(a.layer[4], m.dgrad, rows, H, D, bn, 0, 0, CK_DGRAD, 0, 3);
  a.dgrad[0] = EpiDgradTanh{nullptr, nullptr, H, d1h, d1l, b.ldh, hdh, hdl};
  // layer 5 dz: dz = da1.W1^T -> dmu, dls and their mirror
  fill_layer(a.layer[5], m.dz, rows, Z, H, 64, 0, 0, CK_DZ, 0, 4);
  a.dz = EpiDzPrep{z, eps, mu, ls, Z, la, w, dmu, dls, (__nv_bfloat16*)b.ddh, (__nv_bfloat16*)b.ddl, b.ldq};
  // layer 6 dgrad h_e: da3 = ([dmu|dls].[W4^T;W5^T]) * (1 - h_e^2)
  fill_layer(a.layer[6], m.dhe, rows, H, 2 * Z, bn, 1, 0, CK_DGRAD, 1, 5);
  a.dgrad[1] = EpiDgradTanh{nullptr, nullptr, H, da3h, da3l, b.ldh, heh, hel};
  a.n_layers = 7;
  int items = 0;
  const int tiles_mw = pair ? (tiles_m + 1) / 2 : tiles_m;      // row blocks (pairs of them) per item column
  for (int l = 0; l < a.n_layers; ++l) {
    a.layer[l].first_item = items;
    items += tiles_mw * a.layer[l].tiles_n;
    if (a.layer[l].dep >= 0) a.layer[l].dep_count = a.layer[a.layer[l].dep].tiles_n;
  }
  a.n_items = items;
  a.ready = ready;
  a.ready_stride = tiles_m + 1;
  a.epoch = epoch;

  int bn_max = 64;
  for (int l = 0; l < a.n_layers; ++l) bn_max = a.layer[l].bn > bn_max ? a.layer[l].bn : bn_max;
  if (pair) bn_max /= 2;                                         // a CTA of a pair stages half of the B tile
  a.stages = chain_stages(ns, bn_max);
  a.stage_bytes = chain_stage_bytes(ns, bn_max);

  cudaLaunchConfig_t cfg{};
  int grid = items < n_sm ? items : n_sm;
  if (pair) { const int workers = items < n_sm / 2 ? items : n_sm / 2; grid = 2 * workers; }
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(CH_THREADS);
  cfg.dynamicSmemBytes = CH_SMEM_TOTAL;
  cfg.stream = st;
  cudaLaunchAttribute attr[2]{};
  int na = 0;
  static const bool no_pdl = getenv("VAEB_NO_PDL") != nullptr;
  if (!no_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  if (pair) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  ++*launches;
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(tc_chain_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_TOTAL);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_chain_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_TOTAL);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_chain_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_TOTAL);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_chain_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_TOTAL);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  auto launch = [&]() -> cudaError_t {
    if (pair) return ns == 2 ? cudaLaunchKernelEx(&cfg, tc_chain_kernel<2, true>, a) : cudaLaunchKernelEx(&cfg, tc_chain_kernel<1, true>, a);
    return ns == 2 ? cudaLaunchKernelEx(&cfg, tc_chain_kernel<2, false>, a) : cudaLaunchKernelEx(&cfg, tc_chain_kernel<1, false>, a);
  };
  a.stamps = nullptr;
  a.diag = getenv("VAEB_CHAIN_DIAG") ? 1 : 0;
  a.throttle = getenv("VAEB_CHAIN_THROTTLE") ? 1 : 0;
  a.ld32 = (getenv("VAEB_CHAIN_LD32") && getenv("VAEB_CHAIN_LD32")[0] == '0') ? 0 : 1;   // measured: 256-wide tiles 5-30 % shorter epilogues
  static const char* stamp_path = getenv("VAEB_CHAIN_STAMPS");      // debug: dump the per-item time stamps of every launch
  if (stamp_path) {
    static long long* d_st = nullptr; static int cap = 0;
    const int slots = pair ? 2 * items : items;
    if (slots > cap) { if (d_st) cudaFree(d_st); cudaMalloc((void**)&d_st, (size_t)slots * 8 * sizeof(long long)); cap = slots; }
    cudaMemsetAsync(d_st, 0, (size_t)slots * 8 * sizeof(long long), st);
    a.stamps = d_st;
    cudaError_t e = launch();
    if (e != cudaSuccess) return e;
    cudaStreamSynchronize(st);
    std::vector<long long> hst((size_t)slots * 8);
    cudaMemcpy(hst.data(), d_st, hst.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    FILE* f = fopen(stamp_path, "wb");
    if (f) {
      int hdr[2 + 2 * CH_MAX_LAYERS] = {slots, a.n_layers};
      for (int l = 0; l < a.n_layers; ++l) { hdr[2 + 2 * l] = (pair ? 2 : 1) * a.layer[l].first_item; hdr[3 + 2 * l] = a.layer[l].tiles_n; }
      fwrite(hdr, sizeof(int), 2 + 2 * CH_MAX_LAYERS, f);
      fwrite(hst.data(), sizeof(long long), hst.size(), f);
      fclose(f);
    }
    return cudaSuccess;
  }
  return launch();
}
