#define PMV_PDL_FAMILY 2
// tcgen05 / TMEM / TMA pooling attention, backward (bf16 operands, fp32 accumulation).
//
// Two kernels, both structured like the forward kernel (attn_tc.cu) and both free of the [Nq, Nk] score
// matrix; the scores and probabilities are recomputed per 128 x 128 tile from the saved log-sum-exp:
//
//   dQ kernel  (CTA = 128 query rows, walks the key tiles)            lanes = queries
//       S  = Q' K'^T            SS MMA      P  = exp2(S c - lse2)          softmax warps, fp32
//       dP = dO V^T             SS MMA      dS = P (dP - delta) scale  ->  bf16 into TMEM (over S)
//       dQ' += dS K'            TS MMA (A = dS from TMEM, B = K' read MN-major from the same smem tile)
//     also produces delta[q] = sum_c dO[q,c] O_pre[q,c] for the second kernel and folds the residual-pooling
//     path (dQ'[:, :96] += dO for rows >= 1) into its epilogue.
//
//   dK/dV kernel (CTA = 128 keys x a chunk of query tiles)             lanes = keys  (everything transposed)
//       S^T  = K' Q'^T          SS MMA      P^T  = exp2(S^T c - lse2[q])
//       dP^T = V dO^T           SS MMA      dS^T = P^T (dP^T - delta[q]) scale -> bf16 into TMEM
//       dV += P^T dO            TS MMA (B = dO read MN-major)          dK += dS^T Q'[:, :96]   TS MMA
//     partial sums of the query chunks are added into fp32 workspaces with red.global.add.
//
// Recomputing S / dP in both kernels costs 1.4x the tensor work of a fused kernel but needs no shared-memory
// round trip for P / dS and no per-tile atomics.  Operand tiles are staged as 64-column 128B-swizzle blocks
// (+ one 32-column 64B-swizzle block for the bias columns when kd = 160); a tile is read K-major or MN-major
// through the matrix descriptor only (layouts pinned by tests/test_tcgen05_probe.py).
#include "tc_common.cuh"

#ifdef PMV_ATTN_TRACE
// Debug build only (scripts/attn_trace.py bwd): per-CTA cycle stamps of the dQ kernel's phases.
constexpr int BTRACE_CTAS = 1024, BTRACE_SLOTS = 24;
__device__ long long pmv_attn_bwd_trace_buf[BTRACE_CTAS * BTRACE_SLOTS];
#define BTRACE(slot)                                                                                                  \
  do {                                                                                                                \
    const int cta_ = blockIdx.y * gridDim.x + blockIdx.x;                                                             \
    if (cta_ < BTRACE_CTAS) pmv_attn_bwd_trace_buf[cta_ * BTRACE_SLOTS + (slot)] = clock64();                         \
  } while (0)
extern "C" int pmv_debug_attn_bwd_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, pmv_attn_bwd_trace_buf, sizeof(long long) * BTRACE_CTAS * BTRACE_SLOTS) == cudaSuccess ? 0 : 1;
}
// dK/dV kernel: 4096 CTAs x 32 slots (scripts/attn_trace_dkv.py)
constexpr int KTRACE_CTAS = 4096, KTRACE_SLOTS = 32;
__device__ long long pmv_attn_dkv_trace_buf[KTRACE_CTAS * KTRACE_SLOTS];
#define KTRACE(slot)                                                                                                  \
  do {                                                                                                                \
    const int cta_ = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;                                  \
    if (cta_ < KTRACE_CTAS) pmv_attn_dkv_trace_buf[cta_ * KTRACE_SLOTS + (slot)] = clock64();                         \
  } while (0)
#define KTRACE_VAL(slot, val)                                                                                         \
  do {                                                                                                                \
    const int cta_ = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;                                  \
    if (cta_ < KTRACE_CTAS) pmv_attn_dkv_trace_buf[cta_ * KTRACE_SLOTS + (slot)] = (long long)(val);                  \
  } while (0)
extern "C" int pmv_debug_attn_dkv_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, pmv_attn_dkv_trace_buf, sizeof(long long) * KTRACE_CTAS * KTRACE_SLOTS) == cudaSuccess ? 0 : 1;
}
#else
#define BTRACE(slot) do { } while (0)
#define KTRACE(slot) do { } while (0)
#define KTRACE_VAL(slot, val) do { } while (0)
#endif

namespace {

constexpr int HD = PMV_HEAD_DIM;
constexpr int BT = 128;  // tile edge (queries and keys)
constexpr int THREADS = 384;  // warp 0 TMA, 1 MMA, 2 TMEM allocation, 4..11 softmax (two warps per TMEM lane quarter)
constexpr int SM_WARPS = 8;
constexpr int X96_BYTES = 2 * BT * 128;  // a [128 x 96] operand staged as two 64-column blocks (second half-used)

struct BwdGeom {
  int B, heads, Nq, Nk;
  float scale;
  int residual;
  int64_t ld_qk;
  int chunks;  // dK/dV kernel: the query tiles are split evenly over gridDim.z = chunks CTAs per (key tile, batch, head)
};

template <int KD> struct BCfg {
  static constexpr int QK_BYTES = BT * KD * 2;
  static constexpr int STAGE_BYTES = QK_BYTES + X96_BYTES;
  static constexpr int SMEM_BYTES = 3 * STAGE_BYTES + 1024 /*align*/ + 2048 /*lse, delta*/ + 256;
};

// K-major descriptor of k-step ks (16 columns) inside a [rows x KD] tile, starting `half` * 64 rows into the tile
// (128B-swizzle blocks: 64 rows = 8192 B; the 64B-swizzle block of the bias columns: 64 rows = 4096 B)
// ONE thread issues every MMA of a CTA (18 - 26 per half-tile here), and building two 64-bit descriptors from a byte
// address per MMA (shift, mask, or, for both words) made that thread the bottleneck of both kernels: the softmax warps
// waited 0.55 us per half for it (scripts/attn_trace_bwd.py).  The high word of a descriptor depends only on the layout
// and the low word is (address >> 4) | (LBO >> 4) << 16, which is additive in the address: a tile's low words are made
// once and each MMA adds a constant.
struct KDesc { uint32_t lo128, lo64; };  // K-major [rows x KD] tile: 128B-swizzle blocks / the 64B-swizzle bias block
__device__ __forceinline__ uint32_t desc_lo(uint64_t d) { return (uint32_t)d; }
__device__ __forceinline__ uint32_t desc_hi(uint64_t d) { return (uint32_t)(d >> 32); }
__device__ __forceinline__ uint64_t desc_join(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }
__device__ __forceinline__ KDesc kdesc(uint32_t base) {
  KDesc d;
  d.lo128 = desc_lo(tc::make_smem_desc(base, 16, 1024, tc::SWIZZLE_128B));
  d.lo64 = desc_lo(tc::make_smem_desc(base + 32768, 16, 512, tc::SWIZZLE_64B));
  return d;
}
template <int KD> __device__ __forceinline__ uint64_t kmajor_desc(KDesc d, int ks, int half = 0) {
  if (ks < 8)
    return desc_join(d.lo128 + (uint32_t)((ks >> 2) * 1024 + half * 512 + (ks & 3) * 2), desc_hi(tc::make_smem_desc(0, 16, 1024, tc::SWIZZLE_128B)));
  return desc_join(d.lo64 + (uint32_t)(half * 256 + (ks - 8) * 2), desc_hi(tc::make_smem_desc(0, 16, 512, tc::SWIZZLE_64B)));
}
// MN-major reads of the same tiles (B operand of the "retire" products): 16-row k-steps are 2048 B (1024 B) apart
struct MDesc { uint32_t lo128, lo64; };
__device__ __forceinline__ MDesc mdesc(uint32_t base) {
  MDesc d;
  d.lo128 = desc_lo(tc::make_smem_desc(base, 16384, 1024, tc::SWIZZLE_128B));
  d.lo64 = desc_lo(tc::make_smem_desc(base + 32768, 8192, 512, tc::SWIZZLE_64B));
  return d;
}
__device__ __forceinline__ uint64_t mn128_desc(MDesc d, int kk) { return desc_join(d.lo128 + (uint32_t)kk * 128, desc_hi(tc::make_smem_desc(0, 16384, 1024, tc::SWIZZLE_128B))); }
__device__ __forceinline__ uint64_t mn64_desc(MDesc d, int kk) { return desc_join(d.lo64 + (uint32_t)kk * 64, desc_hi(tc::make_smem_desc(0, 8192, 512, tc::SWIZZLE_64B))); }

__device__ __forceinline__ void load_tile_qk(uint8_t* dst, const CUtensorMap* m128, const CUtensorMap* m64, int kd, int row0,
                                             int bh, uint64_t* bar) {
  tc::tma_load_3d(dst, m128, 0, row0, bh, bar);
  tc::tma_load_3d(dst + 16384, m128, 64, row0, bh, bar);
  if (kd == 160) tc::tma_load_3d(dst + 32768, m64, 128, row0, bh, bar);
}

// Software pipeline shared by both kernels.  A 128-wide tile of the "other" axis (keys in the dQ kernel, queries in
// the dK/dV kernel) is processed as two halves of 64; the score-like accumulators S and dP are double-buffered in
// TMEM (2 x 64 columns each), so the tensor core computes S / dP of half h+1 while the softmax warps turn half h
// into P / dS, and the products that consume P / dS of half h are issued right after.  Before this the chain
// MMA -> softmax -> MMA ran serially per tile (profiles/r01_kernels_after.txt: 80-230 TFLOP/s).
constexpr int HK = 64;
// where the 16-element k-step ks of a half sits after the softmax warps packed it to bf16: warp-half 0 (elements 0..31)
// writes columns [0, 16) of the S buffer, warp-half 1 (elements 32..63) columns [32, 48) — each behind its own reads
__device__ __forceinline__ uint32_t packed_col(int ks) { return ks < 2 ? ks * 8 : 32 + (ks - 2) * 8; }

// delta[bh, n] = sum_c dO[b, n, head, c] * O_pre[b, n, head, c]: 4 lanes per (token, head) row, 24 channels (three
// 16-byte loads per tensor) each, 8 rows per warp: both tensors are read once, fully coalesced.
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ o_pre,
                                                         float* __restrict__ delta, int B, int heads, int Nq) {
  pdl_wait();
  const int64_t rows = (int64_t)B * Nq * heads;  // (b, n, head) in memory order
  const int sub = threadIdx.x & 3;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2; r < ((rows + 7) & ~(int64_t)7); r += ((int64_t)gridDim.x * blockDim.x) >> 2) {
    float acc = 0.f;
    if (r < rows) {
      const uint4* a = reinterpret_cast<const uint4*>(dout + r * HD + sub * 24);
      const uint4* b = reinterpret_cast<const uint4*>(o_pre + r * HD + sub * 24);
#pragma unroll
      for (int v = 0; v < 3; ++v) {
        const uint4 x = __ldg(a + v), y = __ldg(b + v);
        const __nv_bfloat162* x2 = reinterpret_cast<const __nv_bfloat162*>(&x);
        const __nv_bfloat162* y2 = reinterpret_cast<const __nv_bfloat162*>(&y);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          acc = fmaf(__low2float(x2[i]), __low2float(y2[i]), acc);
          acc = fmaf(__high2float(x2[i]), __high2float(y2[i]), acc);
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (sub == 0 && r < rows) {
      const int head = (int)(r % heads);
      const int64_t bn = r / heads;
      const int n = (int)(bn % Nq);
      const int64_t b_ = bn / Nq;
      delta[(b_ * heads + head) * Nq + n] = acc;
    }
  }
}

// ================================================================================================ dQ kernel
template <int KD>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ2,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmK2,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const bf16* __restrict__ dout, const float* __restrict__ lse,
                   const float* __restrict__ delta_in, bf16* __restrict__ dq_aug, BwdGeom g) {
  using Cfg = BCfg<KD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                    // Q' tile, then dO tile
  uint8_t* sdO = smem + Cfg::QK_BYTES;
  uint8_t* sStage = smem + Cfg::STAGE_BYTES;  // 2 x (K' tile, V tile)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * Cfg::STAGE_BYTES + 2048);
  uint64_t* q_full = bars;        // [1]
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;  // [2]  S and dP of a half are in TMEM buffer b
  uint64_t* ds_full = bars + 7;   // [2]  dS of buffer b written (8 warps)
  uint64_t* dq_final = bars + 9;  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) BTRACE(0);
  const int bh = blockIdx.y;
  const int bidx = bh / g.heads, head = bh - bidx * g.heads;
  const int q0 = blockIdx.x * BT;
  const int ntiles = (g.Nk + BT - 1) / BT;

  if (warp == 0 && lane == 0) {
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&kv_full[i], 1);
      tc::mbar_init(&kv_empty[i], 1);
      tc::mbar_init(&sdp_full[i], 1);
      tc::mbar_init(&ds_full[i], SM_WARPS);
    }
    tc::mbar_init(dq_final, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) { tc::tmem_alloc(tmem_slot, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_s = tmem_base, tmem_dp = tmem_base + 128, tmem_dq = tmem_base + 256;  // S / dP: buffer b at + 64 b
  pdl_wait();  // the prologue above overlaps the previous kernel's tail
  if (threadIdx.x == 0) BTRACE(1);

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, Cfg::STAGE_BYTES);
      load_tile_qk(sQ, &tmQ, &tmQ2, KD, q0, bh, q_full);
      tc::tma_load_3d(sdO, &tmdO, head * HD, q0, bidx, q_full);
      tc::tma_load_3d(sdO + 16384, &tmdO, head * HD + 64, q0, bidx, q_full);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        tc::mbar_wait(&kv_empty[st], ((j >> 1) & 1) ^ 1);
        uint8_t* sK = sStage + st * Cfg::STAGE_BYTES;
        uint8_t* sV = sK + Cfg::QK_BYTES;
        tc::mbar_expect_tx(&kv_full[st], Cfg::STAGE_BYTES);
        load_tile_qk(sK, &tmK, &tmK2, KD, j * BT, bh, &kv_full[st]);
        tc::tma_load_3d(sV, &tmV, 0, j * BT, bh, &kv_full[st]);
        tc::tma_load_3d(sV + 16384, &tmV, 64, j * BT, bh, &kv_full[st]);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const KDesc dQt = kdesc(tc::smem_u32(sQ)), ddO = kdesc(tc::smem_u32(sdO));
      const uint32_t idesc_128 = tc::make_idesc_bf16(BT, 128, false, true);
      const uint32_t idesc_32 = tc::make_idesc_bf16(BT, 32, false, true);
      tc::mbar_wait(q_full, 0);
      tc::tc_fence_after();
      int hh = 0;  // live halves issued so far
      bool have_prev = false, first_dq = true;
      int p_b = 0, p_hh = 0, p_nks = 0, p_half = 0, p_st = 0;
      bool p_last = false;
      MDesc p_sk{0, 0};
      // dQ' += dS K' for the half recorded in p_*: A = dS (TMEM, packed bf16), B = K' rows of that half read MN-major
      auto retire = [&]() {
        tc::mbar_wait(&ds_full[p_b], (uint32_t)((p_hh >> 1) & 1));
        tc::tc_fence_after();
        const uint32_t a0 = tmem_s + p_b * HK;
        MDesc d = p_sk;  // advanced to the half's first k-step: a k-step is 128 (64) descriptor units
        d.lo128 += (uint32_t)(4 * p_half) * 128;
        d.lo64 += (uint32_t)(4 * p_half) * 64;
        auto one = [&](int ks, uint32_t acc) {
          tc::umma_ts(tmem_dq, a0 + packed_col(ks), mn128_desc(d, ks), idesc_128, acc);
          if (KD == 160) tc::umma_ts(tmem_dq + 128, a0 + packed_col(ks), mn64_desc(d, ks), idesc_32, acc);
        };
        if (p_nks == 4) {  // full half: straight-line issue (the loop form cost ~80 ns per MMA on the single issuing thread)
          one(0, first_dq ? 0u : 1u);
          one(1, 1u);
          one(2, 1u);
          one(3, 1u);
        } else {
          for (int ks = 0; ks < p_nks; ++ks) one(ks, (first_dq && ks == 0) ? 0u : 1u);
        }
        first_dq = false;
        if (p_last) tc::umma_commit(&kv_empty[p_st]);
      };
      for (int j = 0; j < ntiles; ++j) {
        const int st = j & 1;
        tc::mbar_wait(&kv_full[st], (j >> 1) & 1);
        tc::tc_fence_after();
        const int nvalid = min(BT, g.Nk - j * BT);
        const uint32_t sk_addr = tc::smem_u32(sStage + st * Cfg::STAGE_BYTES);
        const KDesc dK = kdesc(sk_addr), dV = kdesc(sk_addr + Cfg::QK_BYTES);
        const MDesc mK = mdesc(sk_addr);
        for (int half = 0; half < 2; ++half) {
          const int nv = min(HK, nvalid - HK * half);
          if (nv <= 0) break;
          const int n16 = (nv + 15) & ~15;
          const int b = hh & 1;
          const uint32_t idesc_s = tc::make_idesc_bf16(BT, n16, false, false);
          if (hh == 2) BTRACE(21);
#pragma unroll
          for (int ks = 0; ks < KD / 16; ++ks)  // S = Q' K'^T (keys of this half)
            tc::umma_ss(tmem_s + b * HK, kmajor_desc<KD>(dQt, ks), kmajor_desc<KD>(dK, ks, half), idesc_s, ks > 0);
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)  // dP = dO V^T
            tc::umma_ss(tmem_dp + b * HK, kmajor_desc<128>(ddO, ks), kmajor_desc<128>(dV, ks, half), idesc_s, ks > 0);
          tc::umma_commit(&sdp_full[b]);
          if (hh == 2) BTRACE(22);
          if (have_prev) retire();
          if (hh == 2) BTRACE(23);
          have_prev = true;
          p_b = b; p_hh = hh; p_nks = n16 >> 4; p_half = half; p_st = st; p_sk = mK;
          p_last = (half == 1) || (nvalid <= HK);
          ++hh;
        }
      }
      if (have_prev) retire();
      tc::umma_commit(dq_final);
    }
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int wh = (warp - 4) >> 2;  // which 32 elements of a half this warp converts
    const int row = qd * 32 + lane;
    const int n = q0 + row;
    const bool rvalid = n < g.Nq;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const float c = g.scale * 1.4426950408889634f;
    const int64_t ld_o = (int64_t)g.heads * HD;
    const int64_t ooff = ((int64_t)bidx * g.Nq + n) * ld_o + head * HD;
    // delta = rowsum(dO * O_pre) comes from attn_delta_kernel (coalesced; it was 18 % of this kernel when every thread
    // fetched its own two 192-byte rows)
    float delta_s = 0.f, lse2 = 0.f;  // delta * scale, lse * log2(e)
    if (rvalid) {
      delta_s = delta_in[(int64_t)bh * g.Nq + n] * g.scale;
      lse2 = lse[(int64_t)bh * g.Nq + n] * 1.4426950408889634f;
    }
    int hh = 0;
    for (int j = 0; j < ntiles; ++j) {
      const int nvalid = min(BT, g.Nk - j * BT);
      for (int half = 0; half < 2; ++half) {
        const int nv = min(HK, nvalid - HK * half);
        if (nv <= 0) break;
        const int b = hh & 1;
        tc::mbar_wait(&sdp_full[b], (uint32_t)((hh >> 1) & 1));
        tc::tc_fence_after();
        if (warp == 4 && lane == 0 && hh < 8) BTRACE(2 + 2 * hh);
        if (wh * 32 < nv) {
          uint32_t s[32], dp[32], pk[16];
          tc::tmem_ld32(tmem_s + lane_addr + b * HK + wh * 32, s);
          tc::tmem_ld32(tmem_dp + lane_addr + b * HK + wh * 32, dp);
          tc::tmem_ld_wait();
          // No masking of keys >= Nk or queries >= Nq: TMA zero-fills those rows of K', V, Q' and dO, so a stray dS column
          // multiplies a zero K' row in dQ' += dS K', and a stray query row has S = dP = delta = 0 (its dQ' is not stored).
          // 4 instructions per score: FFMA, MUFU.EX2, FFMA, FMUL (+ the pack on the integer ALU).
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float d[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float p = tc::fast_ex2(fmaf(__uint_as_float(s[2 * i + h]), c, -lse2));
              d[h] = p * fmaf(__uint_as_float(dp[2 * i + h]), g.scale, -delta_s);
            }
            pk[i] = tc::pack_bf16x2_alu(d[0], d[1]);
          }
          tc::tmem_st16(tmem_s + lane_addr + b * HK + wh * 32, pk);
          tc::tmem_st_wait();
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&ds_full[b]);
        if (warp == 4 && lane == 0 && hh < 8) BTRACE(3 + 2 * hh);
        ++hh;
      }
    }
    // epilogue: dQ' (+ dO on the first 96 columns for rows >= 1: residual pooling) -> bf16
    tc::mbar_wait(dq_final, 0);
    tc::tc_fence_after();
    if (warp == 4 && lane == 0) BTRACE(18);
    bf16* dqp = dq_aug + ((int64_t)bh * g.Nq + n) * g.ld_qk;
    const bool add_do = g.residual && n >= 1;
#pragma unroll 1
    for (int ch = wh; ch < KD / 32; ch += 2) {
      uint32_t o[32];
      tc::tmem_ld32(tmem_dq + lane_addr + ch * 32, o);
      tc::tmem_ld_wait();
      if (rvalid) {
#pragma unroll
        for (int v8 = 0; v8 < 4; ++v8) {
          const int col = ch * 32 + v8 * 8;
          float f[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(o[v8 * 8 + i]);
          if (add_do && col < HD) {
            // dO of this row from the swizzled smem tile the TMA brought in (block = 64 columns, 16-byte chunks XOR row & 7)
            const uint4 a = *reinterpret_cast<const uint4*>(sdO + (col >> 6) * 16384 + row * 128 + ((((col & 63) >> 3) ^ (row & 7)) << 4));
            const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
#pragma unroll
            for (int i = 0; i < 4; ++i) { f[2 * i] += __low2float(a2[i]); f[2 * i + 1] += __high2float(a2[i]); }
          }
          uint4 pk;
          __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
          __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
          pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
          pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
          *reinterpret_cast<uint4*>(dqp + col) = pk;
        }
      }
    }
    if (warp == 4 && lane == 0) BTRACE(19);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) { tc::tc_fence_after(); tc::tmem_dealloc(tmem_base, 512); if (lane == 0) BTRACE(20); }
}

// ================================================================================================ dK / dV kernel
template <int KD>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ2,
                    const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmK2,
                    const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                    const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV,
                    const float* __restrict__ lse, const float* __restrict__ delta, BwdGeom g) {
  using Cfg = BCfg<KD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sK = smem;  // K' tile then V tile (loaded once)
  uint8_t* sV = smem + Cfg::QK_BYTES;
  uint8_t* sStage = smem + Cfg::STAGE_BYTES;  // 2 x (Q' tile, dO tile)
  float* s_lse = reinterpret_cast<float*>(smem + 3 * Cfg::STAGE_BYTES);  // [2][128] lse * log2e
  float* s_delta = s_lse + 256;                                           // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * Cfg::STAGE_BYTES + 2048);
  uint64_t* k_full = bars;       // [1]
  uint64_t* q_full = bars + 1;   // [2]
  uint64_t* q_empty = bars + 3;  // [2]
  uint64_t* sdp_full = bars + 5;  // [2]
  uint64_t* pds_full = bars + 7;  // [2]  P^T and dS^T of buffer b written (8 warps)
  uint64_t* final_bar = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int bidx = bh / g.heads, head = bh - bidx * g.heads;
  const int k0 = blockIdx.x * BT;
  const int total_qt = (g.Nq + BT - 1) / BT;
  const int qt_begin = (int)(((int64_t)blockIdx.z * total_qt) / g.chunks);
  const int qt_end = (int)(((int64_t)(blockIdx.z + 1) * total_qt) / g.chunks);
  const int nq_tiles = qt_end - qt_begin;  // >= 1: chunks <= total_qt
#ifdef PMV_ATTN_TRACE
  if (threadIdx.x == 0) {
    KTRACE(0);
    unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    KTRACE_VAL(29, smid); KTRACE_VAL(30, gt);
  }
#endif

  if (warp == 0 && lane == 0) {
    tc::mbar_init(k_full, 1);
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&q_full[i], 1);
      tc::mbar_init(&q_empty[i], 1);
      tc::mbar_init(&sdp_full[i], 1);
      tc::mbar_init(&pds_full[i], SM_WARPS);
    }
    tc::mbar_init(final_bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) { tc::tmem_alloc(tmem_slot, 512); tc::tmem_relinquish(); }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_st = tmem_base, tmem_dpt = tmem_base + 128, tmem_dv = tmem_base + 256, tmem_dk = tmem_base + 384;
  if (threadIdx.x == 0) KTRACE(1);
  pdl_wait();  // the prologue above overlaps the previous kernel's tail

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(k_full, Cfg::STAGE_BYTES);
      load_tile_qk(sK, &tmK, &tmK2, KD, k0, bh, k_full);
      tc::tma_load_3d(sV, &tmV, 0, k0, bh, k_full);
      tc::tma_load_3d(sV + 16384, &tmV, 64, k0, bh, k_full);
      for (int i = 0; i < nq_tiles; ++i) {
        const int st = i & 1;
        const int q0 = (qt_begin + i) * BT;
        tc::mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
        uint8_t* sQ = sStage + st * Cfg::STAGE_BYTES;
        uint8_t* sdO = sQ + Cfg::QK_BYTES;
        tc::mbar_expect_tx(&q_full[st], Cfg::STAGE_BYTES);
        load_tile_qk(sQ, &tmQ, &tmQ2, KD, q0, bh, &q_full[st]);
        tc::tma_load_3d(sdO, &tmdO, head * HD, q0, bidx, &q_full[st]);
        tc::tma_load_3d(sdO + 16384, &tmdO, head * HD + 64, q0, bidx, &q_full[st]);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const KDesc dKt = kdesc(tc::smem_u32(sK)), dVt = kdesc(tc::smem_u32(sV));
      tc::mbar_wait(k_full, 0);
      tc::tc_fence_after();
      KTRACE(2);
      const uint32_t idesc_o = tc::make_idesc_bf16(BT, HD, false, true);  // N = 96 channels, B read MN-major
      int hh = 0;
      bool have_prev = false, first_acc = true;
      int p_b = 0, p_hh = 0, p_nks = 0, p_half = 0, p_st = 0;
      bool p_last = false;
      MDesc p_sq{0, 0}, p_sdo{0, 0};
      // dV += P^T dO ; dK += dS^T Q'[:, :96] for the recorded half (B tiles read MN-major: 16-query steps are 2048 B apart)
      auto retire = [&]() {
        tc::mbar_wait(&pds_full[p_b], (uint32_t)((p_hh >> 1) & 1));
        tc::tc_fence_after();
        const uint32_t ap = tmem_st + p_b * HK, ad = tmem_dpt + p_b * HK;
        MDesc d_do = p_sdo, d_q = p_sq;
        d_do.lo128 += (uint32_t)(4 * p_half) * 128;
        d_q.lo128 += (uint32_t)(4 * p_half) * 128;
        auto one = [&](int ks, uint32_t acc) {
          tc::umma_ts(tmem_dv, ap + packed_col(ks), mn128_desc(d_do, ks), idesc_o, acc);
          tc::umma_ts(tmem_dk, ad + packed_col(ks), mn128_desc(d_q, ks), idesc_o, acc);
        };
        if (p_nks == 4) {  // full half: straight-line issue
          one(0, first_acc ? 0u : 1u);
          one(1, 1u);
          one(2, 1u);
          one(3, 1u);
        } else {
          for (int ks = 0; ks < p_nks; ++ks) one(ks, (first_acc && ks == 0) ? 0u : 1u);
        }
        first_acc = false;
        if (p_last) tc::umma_commit(&q_empty[p_st]);
      };
      for (int i = 0; i < nq_tiles; ++i) {
        const int st = i & 1;
        const int q0 = (qt_begin + i) * BT;
        tc::mbar_wait(&q_full[st], (i >> 1) & 1);
        tc::tc_fence_after();
        if (i < 3) KTRACE(3 + i);
        const int nqv = min(BT, g.Nq - q0);
        const uint32_t sq_addr = tc::smem_u32(sStage + st * Cfg::STAGE_BYTES);
        const KDesc dQs = kdesc(sq_addr), dOs = kdesc(sq_addr + Cfg::QK_BYTES);
        const MDesc mQ = mdesc(sq_addr), mdO = mdesc(sq_addr + Cfg::QK_BYTES);
        for (int half = 0; half < 2; ++half) {
          const int nv = min(HK, nqv - HK * half);
          if (nv <= 0) break;
          const int n16 = (nv + 15) & ~15;
          const int b = hh & 1;
          const uint32_t idesc_s = tc::make_idesc_bf16(BT, n16, false, false);
          if (hh == 2) KTRACE(6);
#pragma unroll
          for (int ks = 0; ks < KD / 16; ++ks)  // S^T = K' Q'^T (queries of this half)
            tc::umma_ss(tmem_st + b * HK, kmajor_desc<KD>(dKt, ks), kmajor_desc<KD>(dQs, ks, half), idesc_s, ks > 0);
#pragma unroll
          for (int ks = 0; ks < HD / 16; ++ks)  // dP^T = V dO^T
            tc::umma_ss(tmem_dpt + b * HK, kmajor_desc<128>(dVt, ks), kmajor_desc<128>(dOs, ks, half), idesc_s, ks > 0);
          tc::umma_commit(&sdp_full[b]);
          if (hh == 2) KTRACE(7);
          if (have_prev) retire();
          if (hh == 2) KTRACE(8);
          have_prev = true;
          p_b = b; p_hh = hh; p_nks = n16 >> 4; p_half = half; p_st = st; p_sq = mQ; p_sdo = mdO;
          p_last = (half == 1) || (nqv <= HK);
          ++hh;
        }
      }
      if (have_prev) retire();
      tc::umma_commit(final_bar);
    }
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int wh = (warp - 4) >> 2;
    const int row = qd * 32 + lane;  // key row of this thread
    const int tid128 = threadIdx.x - 128;  // 0..255; the first 128 stage lse / delta
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const float c = g.scale * 1.4426950408889634f;
    int hh = 0;
    for (int i = 0; i < nq_tiles; ++i) {
      const int q0 = (qt_begin + i) * BT;
      const int nqv = min(BT, g.Nq - q0);
      float* lse_t = s_lse + (i & 1) * 128;
      float* del_t = s_delta + (i & 1) * 128;
      if (tid128 < 128) {
        const int n = q0 + tid128;
        const bool ok = n < g.Nq;
        lse_t[tid128] = ok ? lse[(int64_t)bh * g.Nq + n] * 1.4426950408889634f : 0.f;
        del_t[tid128] = ok ? delta[(int64_t)bh * g.Nq + n] * g.scale : 0.f;  // delta * scale
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight softmax warps only
      for (int half = 0; half < 2; ++half) {
        const int nv = min(HK, nqv - HK * half);
        if (nv <= 0) break;
        const int b = hh & 1;
        tc::mbar_wait(&sdp_full[b], (uint32_t)((hh >> 1) & 1));
        tc::tc_fence_after();
        if (warp == 4 && lane == 0 && hh < 8) KTRACE(9 + 2 * hh);
        if (wh * 32 < nv) {
          uint32_t s[32], dp[32], pk[16], dk[16];
          tc::tmem_ld32(tmem_st + lane_addr + b * HK + wh * 32, s);
          tc::tmem_ld32(tmem_dpt + lane_addr + b * HK + wh * 32, dp);
          tc::tmem_ld_wait();
          // lse / delta of the 32 queries of this warp half: 16-byte broadcast loads (one per four scores, not two per score)
          const float4* l4 = reinterpret_cast<const float4*>(lse_t + half * HK + wh * 32);
          const float4* d4 = reinterpret_cast<const float4*>(del_t + half * HK + wh * 32);
#pragma unroll
          for (int e4 = 0; e4 < 8; ++e4) {
            const float4 lv = l4[e4], dv = d4[e4];
            const float ls[4] = {lv.x, lv.y, lv.z, lv.w}, ds[4] = {dv.x, dv.y, dv.z, dv.w};
            float p[4], d[4];
            // no masking of queries >= Nq: their Q' and dO rows are zero-filled by TMA, so P^T dO and dS^T Q' gain nothing
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              p[h] = tc::fast_ex2(fmaf(__uint_as_float(s[4 * e4 + h]), c, -ls[h]));
              d[h] = p[h] * fmaf(__uint_as_float(dp[4 * e4 + h]), g.scale, -ds[h]);
            }
            pk[2 * e4] = tc::pack_bf16x2_alu(p[0], p[1]);
            pk[2 * e4 + 1] = tc::pack_bf16x2_alu(p[2], p[3]);
            dk[2 * e4] = tc::pack_bf16x2_alu(d[0], d[1]);
            dk[2 * e4 + 1] = tc::pack_bf16x2_alu(d[2], d[3]);
          }
          tc::tmem_st16(tmem_st + lane_addr + b * HK + wh * 32, pk);
          tc::tmem_st16(tmem_dpt + lane_addr + b * HK + wh * 32, dk);
          tc::tmem_st_wait();
        }
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&pds_full[b]);
        if (warp == 4 && lane == 0 && hh < 8) KTRACE(10 + 2 * hh);
        ++hh;
      }
    }
    if (warp == 4 && lane == 0) KTRACE(25);
    tc::mbar_wait(final_bar, 0);
    tc::tc_fence_after();
    if (warp == 4 && lane == 0) KTRACE(26);
    // dV / dK tiles leave through shared memory: [128 keys x 32 channels] fp32 boxes in the 128B-swizzle layout (a thread
    // owns a key row; 16-byte chunk j of row r sits at chunk j ^ (r & 7), so the eight lanes of a quarter warp hit eight
    // different bank groups), then one bulk tensor reduce-add (several chunks) or store (one chunk) per box.  Per-thread
    // red.global.add.v4 / st.global.v4 with a 384-byte stride between lanes cost 3.2 / 5.5 us per CTA: every warp
    // instruction is 32 separate line requests (scripts/attn_trace_dkv.py).  Keys >= Nk are clipped by the tensor map.
    uint8_t* sOut = sStage;  // the Q' / dO ring is idle: every MMA has retired
#pragma unroll 1
    for (int ch = wh; ch < 3; ch += 2) {
      uint32_t a[32], b[32];
      tc::tmem_ld32(tmem_dv + lane_addr + ch * 32, a);
      tc::tmem_ld32(tmem_dk + lane_addr + ch * 32, b);
      tc::tmem_ld_wait();
      uint8_t* rv = sOut + ch * 16384 + row * 128;
      uint8_t* rk = sOut + (3 + ch) * 16384 + row * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int off = (j ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(rv + off) = make_uint4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
        *reinterpret_cast<uint4*>(rk + off) = make_uint4(b[4 * j], b[4 * j + 1], b[4 * j + 2], b[4 * j + 3]);
      }
    }
    tc::fence_proxy_async();  // generic-proxy writes -> visible to the bulk operations
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (warp == 4 && tc::elect_one()) {
#pragma unroll
      for (int ch = 0; ch < 3; ++ch) {
        if (g.chunks == 1) {
          tc::tma_store_3d(&tmDV, sOut + ch * 16384, ch * 32, k0, bh);
          tc::tma_store_3d(&tmDK, sOut + (3 + ch) * 16384, ch * 32, k0, bh);
        } else {
          tc::tma_reduce_add_3d(&tmDV, sOut + ch * 16384, ch * 32, k0, bh);
          tc::tma_reduce_add_3d(&tmDK, sOut + (3 + ch) * 16384, ch * 32, k0, bh);
        }
      }
      tc::bulk_commit_group();
      tc::bulk_wait_group_read0();  // shared memory is released when the CTA exits
    }
  }
  if (warp == 4 && lane == 0) KTRACE(27);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
#ifdef PMV_ATTN_TRACE
    if (lane == 0) {
      KTRACE(28);
      unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
      KTRACE_VAL(31, gt);
    }
#endif
  }
}

template <typename T>
__global__ void __launch_bounds__(256) cast_rows96_kernel(const float* __restrict__ src, T* __restrict__ dst, int64_t rows, int64_t ld) {
  pdl_wait();
  const int64_t total = rows * (HD / 4);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / (HD / 4);
    const int c4 = (int)(i - r * (HD / 4));
    float v[4];
    load4(src + r * HD + c4 * 4, v);
    store4(dst + r * ld + c4 * 4, v);
  }
}

// How many CTAs share the query tiles of one (key tile, batch, head) in the dK/dV kernel.  Cycle stamps
// (scripts/attn_trace_dkv.py, profiles/r02_attn_dkv_trace.md): a CTA pays ~2 us until its first S^T / dP^T is ready and
// ~3.5 us at the end (final commit, 98 KB of red.global.add per CTA, TMEM release) around ~2.05 us (kd = 128) / ~2.3 us
// (kd = 160) per 128-query tile, one CTA per SM at a time.  The first version aimed at >= 3 CTAs per SM and paid that fixed
// cost 3 - 6 times per SM (mid-stage: 512 CTAs of 4 query tiles, 49 us; one chunk: 128 CTAs of 13 tiles).  Pick the chunk count
// that minimises  rounds x (fixed + tiles per chunk x per-tile).
int dkv_chunks(int q_tiles, int64_t base_ctas, int kd) {
  static int forced = -1;
  if (forced < 0) {
    const char* ev = getenv("PMV_ATTN_DKV_CHUNKS");
    forced = ev != nullptr ? atoi(ev) : 0;
  }
  if (forced > 0) return forced < q_tiles ? forced : q_tiles;
  const double per_tile = kd == 128 ? 2.05 : 2.3;
  double best = 1e30;
  int best_c = 1;
  for (int c = 1; c <= q_tiles && c <= 64; ++c) {
    const int64_t rounds = (base_ctas * c + 147) / 148;
    const double fixed = c == 1 ? 4.5 : 5.8;  // plain stores instead of reductions
    const double t = (double)rounds * (fixed + (double)((q_tiles + c - 1) / c) * per_tile);
    if (t < best - 1e-9) { best = t; best_c = c; }
  }
  return best_c;
}

template <int KD>
int launch_bwd(const void* q_aug, const void* k_aug, int64_t ld_qk, const void* v, int64_t ld_v, const void* o_pre,
               const void* dout, const float* lse, void* dq_aug, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, float* ws,
               BwdGeom g, cudaStream_t stream) {
  using Cfg = BCfg<KD>;
  const uint64_t BH = (uint64_t)g.B * g.heads;
  CUtensorMap tmQ, tmQ2, tmK, tmK2, tmV, tmdO;
  int rc;
  if ((rc = pmv_make_tensor_map_3d(&tmQ, q_aug, 2, KD, g.Nq, BH, ld_qk, (uint64_t)g.Nq * ld_qk, 64, BT, 1, 128))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmK, k_aug, 2, KD, g.Nk, BH, ld_qk, (uint64_t)g.Nk * ld_qk, 64, BT, 1, 128))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmQ2, q_aug, 2, KD, g.Nq, BH, ld_qk, (uint64_t)g.Nq * ld_qk, 32, BT, 1, 64))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmK2, k_aug, 2, KD, g.Nk, BH, ld_qk, (uint64_t)g.Nk * ld_qk, 32, BT, 1, 64))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmV, v, 2, HD, g.Nk, BH, ld_v, (uint64_t)g.Nk * ld_v, 64, BT, 1, 128))) return rc;
  const uint64_t ld_o = (uint64_t)g.heads * HD;
  if ((rc = pmv_make_tensor_map_3d(&tmdO, dout, 2, ld_o, g.Nq, (uint64_t)g.B, ld_o, (uint64_t)g.Nq * ld_o, 64, BT, 1, 128))) return rc;

  const int64_t krows = (int64_t)BH * g.Nk;
  float* dk_ws = ws;
  float* dv_ws = ws + krows * HD;
  float* delta = ws + 2 * krows * HD;
  CUtensorMap tmDK, tmDV;  // fp32 accumulators [BH][Nk][96] as [128 keys x 32 channels] boxes
  if ((rc = pmv_make_tensor_map_3d(&tmDK, dk_ws, 4, HD, g.Nk, BH, HD, (uint64_t)g.Nk * HD, 32, BT, 1, 128))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmDV, dv_ws, 4, HD, g.Nk, BH, HD, (uint64_t)g.Nk * HD, 32, BT, 1, 128))) return rc;

  auto kq = attn_bwd_dq_kernel<KD>;
  auto kkv = attn_bwd_dkv_kernel<KD>;
  static bool attr_set = false;
  if (!attr_set) {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    PMV_CHECK_CUDA(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int q_tiles = (g.Nq + BT - 1) / BT, k_tiles = (g.Nk + BT - 1) / BT;
  g.ld_qk = ld_qk;
  {
    const int64_t rows = (int64_t)g.B * g.Nq * g.heads;
    int64_t dblocks = ceil_div64(rows * 4, 256);
    if (dblocks > 148 * 16) dblocks = 148 * 16;
    pmv_launch(attn_delta_kernel, (unsigned)dblocks, 256, 0, stream, (const bf16*)dout, (const bf16*)o_pre, delta, g.B, g.heads, g.Nq);
  }
  pmv_launch(kq, dim3((unsigned)q_tiles, (unsigned)BH), THREADS, Cfg::SMEM_BYTES, stream, 
      tmQ, tmQ2, tmK, tmK2, tmV, tmdO, (const bf16*)dout, lse, delta, (bf16*)dq_aug, g);
  const int chunks = dkv_chunks(q_tiles, (int64_t)k_tiles * (int64_t)BH, KD);
  g.chunks = chunks;
  // several chunks add their partial dK / dV into the workspace (cleared here, between the two kernels' launches: the dQ
  // kernel does not touch it); a single chunk stores
  if (chunks > 1) PMV_CHECK_CUDA(cudaMemsetAsync(ws, 0, (size_t)(2 * krows * HD) * sizeof(float), stream));
  pmv_launch(kkv, dim3((unsigned)k_tiles, (unsigned)BH, (unsigned)chunks), THREADS, Cfg::SMEM_BYTES, stream, 
      tmQ, tmQ2, tmK, tmK2, tmV, tmdO, tmDK, tmDV, lse, delta, g);
  int64_t cblocks = ceil_div64(krows * (HD / 4), 256);
  if (cblocks > 148 * 8) cblocks = 148 * 8;
  if (dk != nullptr) pmv_launch(cast_rows96_kernel<bf16>, (unsigned)cblocks, 256, 0, stream, dk_ws, (bf16*)dk, krows, ld_dk);
  if (dv != nullptr) pmv_launch(cast_rows96_kernel<bf16>, (unsigned)cblocks, 256, 0, stream, dv_ws, (bf16*)dv, krows, ld_dv);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

}  // namespace

int64_t attn_tc_bwd_workspace_floats(int B, int heads, int Nq, int Nk) {
  return (int64_t)2 * B * heads * Nk * HD + (int64_t)B * heads * Nq;
}

int attn_tc_bwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v, const void* o_pre,
                const void* dout, const float* lse, void* dq_aug, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv, float* ws,
                int B, int heads, int Nq, int Nk, float scale, int residual, cudaStream_t stream) {
  PMV_CHECK_ARG(kd == 128 || kd == 160, "attention bwd(tc): kd must be 128 or 160 (got %d)", kd);
  PMV_CHECK_ARG(ld_qk % 8 == 0 && ld_v % 8 == 0 && ld_dk % 4 == 0 && ld_dv % 4 == 0, "attention bwd(tc): bad row strides");
  BwdGeom g{B, heads, Nq, Nk, scale, residual, ld_qk, 1};
  if (kd == 128) return launch_bwd<128>(q_aug, k_aug, ld_qk, v, ld_v, o_pre, dout, lse, dq_aug, dk, ld_dk, dv, ld_dv, ws, g, stream);
  return launch_bwd<160>(q_aug, k_aug, ld_qk, v, ld_v, o_pre, dout, lse, dq_aug, dk, ld_dk, dv, ld_dv, ws, g, stream);
}
