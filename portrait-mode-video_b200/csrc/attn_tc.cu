#define PMV_PDL_FAMILY 1
// tcgen05 / TMEM / TMA pooling attention, forward (bf16 operands, fp32 softmax and accumulation):
//     O = softmax(scale * Q' K'^T) V  (+ q for rows >= 1: residual pooling), head-merged store.
// The decomposed relative-position bias rides in the augmented columns of Q'/K' (relpos.cu), so the score
// tile that leaves the tensor core is already biased; the [Nq, Nk] score matrix of the reference
// (attention.py:412-446) lives only in tensor memory, 128 x 128 at a time.
//
// One CTA = 128 query rows of one (batch, head).  256 threads:
//   warp 0      TMA producer : Q' once, then K'/V tiles of 128 keys through a 2-stage mbarrier ring
//   warp 1      MMA issuer   : S_j = Q' K'_j^T  (SS, K-major, 128 x 128 x kd)  into TMEM buffer j&1,
//                              O += P_j V_j     (TS: P read from TMEM, V MN-major from smem, 128 x 96 x 128)
//   warp 2      TMEM allocator
//   warps 4..7  softmax      : one thread per query row; tcgen05.ld S -> online softmax with lazy rescaling
//                              (O is touched only when the running max grows by > 2^8) -> bf16 P written back
//                              over S with tcgen05.st -> epilogue (O / l, residual, lse)
// S_{j+1} is issued before the MMA warp waits for P_j, so the tensor core computes the next score tile while
// the softmax warps work on the current one.
//
// Shared-memory operand layouts (pinned by tests/test_tcgen05_probe.py):
//   Q', K' : K-major; columns [0,128) as two 64-wide 128B-swizzle blocks, columns [128,160) (kd = 160 only) as
//            one 32-wide 64B-swizzle block.
//   V      : MN-major (channel axis contiguous), three 32-channel 64B-swizzle groups of [128 keys x 64 B].
//   P      : tensor memory, bf16 pairs packed in 32-bit columns (column c = keys 2c, 2c+1).
#include "tc_common.cuh"

#ifdef PMV_ATTN_TRACE
// Debug build only (scripts/attn_trace.py): per-CTA cycle stamps of the forward kernel's phases.
constexpr int TRACE_CTAS = 1024, TRACE_SLOTS = 24;
__device__ long long pmv_attn_trace_buf[TRACE_CTAS * TRACE_SLOTS];
#define TRACE(slot)                                                                                                   \
  do {                                                                                                                \
    const int cta_ = blockIdx.y * gridDim.x + blockIdx.x;                                                             \
    if (cta_ < TRACE_CTAS) pmv_attn_trace_buf[cta_ * TRACE_SLOTS + (slot)] = clock64();                               \
  } while (0)
extern "C" int pmv_debug_attn_trace(long long* host_out) {
  return cudaMemcpyFromSymbol(host_out, pmv_attn_trace_buf, sizeof(long long) * TRACE_CTAS * TRACE_SLOTS) == cudaSuccess ? 0 : 1;
}
#else
#define TRACE(slot) do { } while (0)
#endif

namespace {

constexpr int HD = PMV_HEAD_DIM;
constexpr int BQ = 128, BKV = 128;
constexpr int THREADS = 256;
constexpr int V_BYTES = 3 * BKV * 64;  // 24576

struct TcAttnGeom {
  int B, heads, Nq, Nk;
  float scale;
  int residual;
};

template <int KD> struct ACfg {
  static constexpr int QK_BYTES = BQ * KD * 2;  // 32768 (kd 128) / 40960 (kd 160)
  static constexpr int STAGE_BYTES = QK_BYTES + V_BYTES;
  // K'/V ring depth: with two stages the load of tile j+2 can only start when P_j V_j has retired, and the softmax warps
  // then wait ~0.4 us per tile for it (scripts/attn_trace.py); three stages fit for kd = 128 (205 KB), not for kd = 160
  static constexpr int NST = KD == 128 ? 3 : 2;
  static constexpr int SMEM_BYTES = QK_BYTES + NST * STAGE_BYTES + 1024 + 128 + 512;  // alignment slack, barriers, row scales
};

template <int KD>
__global__ void __launch_bounds__(THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ2,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmK2,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                   const __grid_constant__ CUtensorMap tmOpre, bf16* __restrict__ out, bf16* __restrict__ out_pre,
                   float* __restrict__ lse, TcAttnGeom g) {
  using Cfg = ACfg<KD>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;
  uint8_t* sStage = smem + Cfg::QK_BYTES;
  constexpr int NST = Cfg::NST;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::QK_BYTES + NST * Cfg::STAGE_BYTES);
  uint64_t* q_full = bars;         // [1]
  uint64_t* kv_full = bars + 1;    // [NST <= 3]
  uint64_t* kv_empty = bars + 4;   // [NST <= 3]
  uint64_t* s_full = bars + 7;     // [2]
  uint64_t* p_full = bars + 9;     // [2]
  uint64_t* o_done = bars + 11;    // [1]  arrives after every P_j V_j
  uint64_t* o_final = bars + 12;   // [1]  arrives once, after the last P V (unambiguous phase for the epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  float* s_inv = reinterpret_cast<float*>(bars + 16);  // [128] 1 / row sum, softmax warps -> epilogue

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int q0 = blockIdx.x * BQ;
  const int ntiles = (g.Nk + BKV - 1) / BKV;
#ifdef PMV_ATTN_TRACE
  if (threadIdx.x == 0) {
    TRACE(0);
    unsigned smid; asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
    const int cta_ = blockIdx.y * gridDim.x + blockIdx.x;
    if (cta_ < TRACE_CTAS) { pmv_attn_trace_buf[cta_ * TRACE_SLOTS + 21] = smid; pmv_attn_trace_buf[cta_ * TRACE_SLOTS + 19] = (long long)gt; }
  }
#endif

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&tmQ);
    tc::tma_prefetch_desc(&tmK);
    tc::tma_prefetch_desc(&tmV);
    tc::mbar_init(q_full, 1);
    for (int i = 0; i < NST; ++i) {
      tc::mbar_init(&kv_full[i], 1);
      tc::mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      tc::mbar_init(&s_full[i], 1);
      tc::mbar_init(&p_full[i], 4);
    }
    tc::mbar_init(o_done, 1);
    tc::mbar_init(o_final, 1);
    tc::fence_barrier_init();
  }
  if (warp == 2) {
    tc::tmem_alloc(tmem_slot, 512);
    tc::tmem_relinquish();
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;
  if (threadIdx.x == 0) TRACE(1);
  pdl_wait();  // the prologue above overlaps the previous kernel's tail

  if (warp == 0) {
    if (tc::elect_one()) {
      tc::mbar_expect_tx(q_full, Cfg::QK_BYTES);
      tc::tma_load_3d(sQ, &tmQ, 0, q0, bh, q_full);
      tc::tma_load_3d(sQ + 16384, &tmQ, 64, q0, bh, q_full);
      if (KD == 160) tc::tma_load_3d(sQ + 32768, &tmQ2, 128, q0, bh, q_full);
      TRACE(2);
      for (int j = 0; j < ntiles; ++j) {
        const int st = j % NST;
        tc::mbar_wait(&kv_empty[st], ((j / NST) & 1) ^ 1);
        uint8_t* sK = sStage + st * Cfg::STAGE_BYTES;
        uint8_t* sV = sK + Cfg::QK_BYTES;
        tc::mbar_expect_tx(&kv_full[st], Cfg::STAGE_BYTES);
        tc::tma_load_3d(sK, &tmK, 0, j * BKV, bh, &kv_full[st]);
        tc::tma_load_3d(sK + 16384, &tmK, 64, j * BKV, bh, &kv_full[st]);
        if (KD == 160) tc::tma_load_3d(sK + 32768, &tmK2, 128, j * BKV, bh, &kv_full[st]);
#pragma unroll
        for (int c = 0; c < 3; ++c) tc::tma_load_3d(sV + c * 8192, &tmV, c * 32, j * BKV, bh, &kv_full[st]);
      }
    }
  } else if (warp == 1) {
    if (tc::elect_one()) {
      const uint32_t sq_addr = tc::smem_u32(sQ);
      const uint32_t hi128 = (uint32_t)(tc::make_smem_desc(0, 16, 1024, tc::SWIZZLE_128B) >> 32);
      const uint32_t hi64 = (uint32_t)(tc::make_smem_desc(0, 16, 512, tc::SWIZZLE_64B) >> 32);
      const uint32_t lq128 = (uint32_t)tc::make_smem_desc(sq_addr, 16, 1024, tc::SWIZZLE_128B);
      const uint32_t lq64 = (uint32_t)tc::make_smem_desc(sq_addr + 32768, 16, 512, tc::SWIZZLE_64B);
      auto issue_s = [&](int j) {
        const int st = j % NST;
        tc::mbar_wait(&kv_full[st], (j / NST) & 1);
        tc::tc_fence_after();
        if (j < 1) TRACE(4 + j);
        const int nvalid = min(BKV, g.Nk - j * BKV);
        const uint32_t idesc = tc::make_idesc_bf16(BQ, (nvalid + 15) & ~15, false, false);
        const uint32_t sk_addr = tc::smem_u32(sStage + st * Cfg::STAGE_BYTES);
        const uint32_t tmem_s = tmem_base + (uint32_t)(j & 1) * 128;
        // descriptor low words are additive in the address (>> 4): one per tile, a constant per MMA (the single issuing
        // thread is on the critical path of the first score tile and of the last P V)
        const uint32_t lk128 = (uint32_t)tc::make_smem_desc(sk_addr, 16, 1024, tc::SWIZZLE_128B);
        const uint32_t lk64 = (uint32_t)tc::make_smem_desc(sk_addr + 32768, 16, 512, tc::SWIZZLE_64B);
#pragma unroll
        for (int ks = 0; ks < KD / 16; ++ks) {
          uint64_t da, db;
          if (ks < 8) {
            const uint32_t off = (uint32_t)(ks >> 2) * 1024 + (uint32_t)(ks & 3) * 2;
            da = ((uint64_t)hi128 << 32) | (lq128 + off);
            db = ((uint64_t)hi128 << 32) | (lk128 + off);
          } else {
            const uint32_t off = (uint32_t)(ks - 8) * 2;
            da = ((uint64_t)hi64 << 32) | (lq64 + off);
            db = ((uint64_t)hi64 << 32) | (lk64 + off);
          }
          tc::umma_ss(tmem_s, da, db, idesc, ks > 0);
        }
        tc::umma_commit(&s_full[j & 1]);
      };
      tc::mbar_wait(q_full, 0);
      tc::tc_fence_after();
      TRACE(3);
      issue_s(0);
      const uint32_t idesc_pv = tc::make_idesc_bf16(BQ, HD, false, true);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);
        tc::mbar_wait(&p_full[j & 1], (j >> 1) & 1);
        tc::tc_fence_after();
        const int nvalid = min(BKV, g.Nk - j * BKV);
        const int nks = (nvalid + 15) >> 4;
        const uint32_t sv_addr = tc::smem_u32(sStage + (j % NST) * Cfg::STAGE_BYTES + Cfg::QK_BYTES);
        const uint32_t tmem_p = tmem_base + (uint32_t)(j & 1) * 128;
        const uint64_t dv0 = tc::make_smem_desc(sv_addr, 8192, 512, tc::SWIZZLE_64B);  // + 64 per 16-key step (1024 B)
        if (nks == 8) {  // full tile: straight-line issue
#pragma unroll
          for (int ks = 0; ks < 8; ++ks) tc::umma_ts(tmem_o, tmem_p + ks * 8, dv0 + (uint64_t)(ks * 64), idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
        } else {
          for (int ks = 0; ks < nks; ++ks) tc::umma_ts(tmem_o, tmem_p + ks * 8, dv0 + (uint64_t)(ks * 64), idesc_pv, (j > 0 || ks > 0) ? 1u : 0u);
        }
        tc::umma_commit(&kv_empty[j % NST]);
        tc::umma_commit(o_done);
        if (j == ntiles - 1) tc::umma_commit(o_final);
      }
    }
  } else if (warp >= 4) {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const float c = g.scale * 1.4426950408889634f;  // scores are exponentiated in base 2
    float m_used = 0.f, l_run = 0.f;
    for (int j = 0; j < ntiles; ++j) {
      const int b = j & 1;
      tc::mbar_wait(&s_full[b], (j >> 1) & 1);
      tc::tc_fence_after();
      if (warp == 4 && lane == 0 && j < 4) TRACE(8 + j);
      const int nvalid = min(BKV, g.Nk - j * BKV);
      const int nchunks = (nvalid + 31) >> 5;
      const uint32_t tmem_s = tmem_base + lane_addr + (uint32_t)b * 128;
      uint32_t s[4][32];
#pragma unroll
      for (int ch = 0; ch < 4; ++ch)
        if (ch < nchunks) tc::tmem_ld32(tmem_s + ch * 32, s[ch]);
      tc::tmem_ld_wait();
      if (warp == 4 && lane == 0 && j == 1) TRACE(22);
      float mx = -INFINITY;
      if (nvalid == BKV) {  // full tile (all but the last): no key masking, 1 instruction per score instead of 3
        float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch)
#pragma unroll
          for (int i = 0; i < 32; ++i) m4[i & 3] = fmaxf(m4[i & 3], __uint_as_float(s[ch][i]));
        mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
      } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          if (ch < nchunks) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              float v = __uint_as_float(s[ch][i]);
              if (ch * 32 + i >= nvalid) v = -INFINITY;
              s[ch][i] = __float_as_uint(v);
              mx = fmaxf(mx, v);
            }
          }
        }
      }
      if (j == 0) {
        m_used = mx;
      } else {
        const bool grow = (mx - m_used) * c > 8.0f;  // lazy rescale: keep the old reference max while exp2 stays <= 2^8
        if (__any_sync(0xffffffffu, grow)) {
          tc::mbar_wait(o_done, (j - 1) & 1);  // P_{j-1} V_{j-1} has landed in O
          tc::tc_fence_after();
          const float alpha = grow ? tc::fast_ex2((m_used - mx) * c) : 1.0f;
#pragma unroll
          for (int ch = 0; ch < 3; ++ch) {
            uint32_t o[32];
            tc::tmem_ld32(tmem_o + lane_addr + ch * 32, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tc::tmem_st32(tmem_o + lane_addr + ch * 32, o);
          }
          tc::tmem_st_wait();
          l_run *= alpha;
          if (grow) m_used = mx;
        }
      }
      const float mc = m_used * c;
      float sum = 0.f;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        if (ch < nchunks) {
          uint32_t pk[16];
          const float2 c2 = make_float2(c, c), nmc2 = make_float2(-mc, -mc);
          float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {  // packed fp32 pairs: one FFMA2 + two MUFU + one FADD2 + one pack per two scores
            const float2 t = __ffma2_rn(make_float2(__uint_as_float(s[ch][2 * i]), __uint_as_float(s[ch][2 * i + 1])), c2, nmc2);
            const float2 pr = make_float2(tc::fast_ex2(t.x), tc::fast_ex2(t.y));
            sum2 = __fadd2_rn(sum2, pr);
            pk[i] = tc::pack_bf16x2_alu(pr.x, pr.y);
          }
          sum += sum2.x + sum2.y;
          tc::tmem_st16(tmem_s + ch * 16, pk);
        }
      }
      l_run += sum;
      if (warp == 4 && lane == 0 && j == 1) TRACE(23);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&p_full[b]);
      if (warp == 4 && lane == 0 && j < 4) TRACE(12 + j);
    }
    // ---------------- hand the row scale to the epilogue (all eight warps), log-sum-exp
    tc::mbar_wait(o_final, 0);
    tc::tc_fence_after();
    if (warp == 4 && lane == 0) TRACE(16);
    s_inv[row] = 1.0f / l_run;
    if (lse != nullptr && q0 + row < g.Nq) lse[(int64_t)bh * g.Nq + q0 + row] = m_used * g.scale + logf(l_run);
  }

  // ---------------- epilogue: O / l (+ residual q), head-merged bf16 tiles, two bulk tensor stores.
  // All eight warps take part (a warp reaches the TMEM lanes 32 (w % 4) .. + 31): warps 4..7 take channels [0, 48), the
  // TMA / MMA / allocator warps, idle by now, channels [48, 96) - one warp per scheduler needed 1.5 us for the 96
  // channels of its rows (scripts/attn_trace.py).  The rows are staged as dense [128][96] bf16 tiles in the idle K'/V
  // ring; direct 16-byte stores from one thread per row touch 32 different lines per warp instruction.
  tc::tc_fence_before();
  __syncthreads();  // O complete (the softmax warps waited for o_final), s_inv written
  tc::tc_fence_after();
  {
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(qd * 32) << 16;
    const int c_base = warp >= 4 ? 0 : 48;
    const int n = q0 + row;
    const float inv = s_inv[row];
    const int bidx = bh / g.heads, head = bh - bidx * g.heads;
    const bool add_q = g.residual && n >= 1;
    uint8_t* sOut = sStage;
    uint8_t* sPre = sStage + BQ * HD * 2;
    uint32_t o32[32], o16[16];
    tc::tmem_ld32(tmem_o + lane_addr + c_base, o32);
    tc::tmem_ld16(tmem_o + lane_addr + c_base + 32, o16);
    tc::tmem_ld_wait();
    if (warp == 4 && lane == 0) TRACE(5);
#pragma unroll
    for (int v8 = 0; v8 < 6; ++v8) {  // 8 channels = one 16-byte chunk of the swizzled Q' row
      const int col = c_base + v8 * 8;
      float f[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(v8 < 4 ? o32[(v8 & 3) * 8 + i] : o16[(v8 & 1) * 8 + i]) * inv;
      if (out_pre != nullptr) {  // pre-residual output, kept for backward (delta = rowsum(dO * O_attn))
        uint4 pk;
        __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
        __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
        pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
        pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
        *reinterpret_cast<uint4*>(sPre + row * (HD * 2) + col * 2) = pk;
      }
      if (add_q) {
        const int blk = col >> 6, cc = (col & 63) >> 3;
        const uint4 qv = *reinterpret_cast<const uint4*>(sQ + blk * 16384 + row * 128 + ((cc ^ (row & 7)) << 4));
        const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&qv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          f[2 * i] += __low2float(q2[i]);
          f[2 * i + 1] += __high2float(q2[i]);
        }
      }
      uint4 pk;
      __nv_bfloat162 t0 = __floats2bfloat162_rn(f[0], f[1]), t1 = __floats2bfloat162_rn(f[2], f[3]);
      __nv_bfloat162 t2 = __floats2bfloat162_rn(f[4], f[5]), t3 = __floats2bfloat162_rn(f[6], f[7]);
      pk.x = *reinterpret_cast<uint32_t*>(&t0); pk.y = *reinterpret_cast<uint32_t*>(&t1);
      pk.z = *reinterpret_cast<uint32_t*>(&t2); pk.w = *reinterpret_cast<uint32_t*>(&t3);
      *reinterpret_cast<uint4*>(sOut + row * (HD * 2) + col * 2) = pk;
    }
    tc::fence_proxy_async();  // generic-proxy writes -> visible to the bulk stores
    if (warp == 4 && lane == 0) TRACE(6);
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 4 && tc::elect_one()) {
      tc::tma_store_3d(&tmO, sOut, head * HD, q0, bidx);
      if (out_pre != nullptr) tc::tma_store_3d(&tmOpre, sPre, head * HD, q0, bidx);
      tc::bulk_commit_group();
      TRACE(7);
      tc::bulk_wait_group_read0();  // shared memory is released when the CTA exits
      TRACE(17);
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc::tc_fence_after();
    tc::tmem_dealloc(tmem_base, 512);
#ifdef PMV_ATTN_TRACE
    if (lane == 0) {
      TRACE(18);
      unsigned long long gt; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(gt));
      const int cta_ = blockIdx.y * gridDim.x + blockIdx.x;
      if (cta_ < TRACE_CTAS) pmv_attn_trace_buf[cta_ * TRACE_SLOTS + 20] = (long long)gt;
    }
#endif
  }
}

template <int KD>
int launch(const void* q_aug, const void* k_aug, int64_t ld_qk, const void* v, int64_t ld_v, void* out, void* out_pre, float* lse,
           const TcAttnGeom& g, cudaStream_t stream) {
  using Cfg = ACfg<KD>;
  const uint64_t BH = (uint64_t)g.B * g.heads;
  CUtensorMap tmQ, tmQ2, tmK, tmK2, tmV;
  int rc;
  if ((rc = pmv_make_tensor_map_3d(&tmQ, q_aug, 2, KD, g.Nq, BH, ld_qk, (uint64_t)g.Nq * ld_qk, 64, BQ, 1, 128))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmK, k_aug, 2, KD, g.Nk, BH, ld_qk, (uint64_t)g.Nk * ld_qk, 64, BKV, 1, 128))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmQ2, q_aug, 2, KD, g.Nq, BH, ld_qk, (uint64_t)g.Nq * ld_qk, 32, BQ, 1, 64))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmK2, k_aug, 2, KD, g.Nk, BH, ld_qk, (uint64_t)g.Nk * ld_qk, 32, BKV, 1, 64))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmV, v, 2, HD, g.Nk, BH, ld_v, (uint64_t)g.Nk * ld_v, 32, BKV, 1, 64))) return rc;
  CUtensorMap tmO, tmOpre;
  const uint64_t ld_o = (uint64_t)g.heads * HD;
  if ((rc = pmv_make_tensor_map_3d(&tmO, out, 2, ld_o, g.Nq, (uint64_t)g.B, ld_o, (uint64_t)g.Nq * ld_o, HD, BQ, 1, 0))) return rc;
  if ((rc = pmv_make_tensor_map_3d(&tmOpre, out_pre != nullptr ? out_pre : out, 2, ld_o, g.Nq, (uint64_t)g.B, ld_o, (uint64_t)g.Nq * ld_o, HD,
                                   BQ, 1, 0)))
    return rc;
  auto kern = attn_tc_fwd_kernel<KD>;
  static bool attr_set = false;
  if (!attr_set) {
    PMV_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((unsigned)((g.Nq + BQ - 1) / BQ), (unsigned)BH);
  pmv_launch(kern, grid, THREADS, Cfg::SMEM_BYTES, stream, tmQ, tmQ2, tmK, tmK2, tmV, tmO, tmOpre, (bf16*)out, (bf16*)out_pre, lse, g);
  PMV_CHECK_LAUNCH();
  return PMV_OK;
}

}  // namespace

int attn_tc_fwd(const void* q_aug, const void* k_aug, int64_t ld_qk, int kd, const void* v, int64_t ld_v, void* out,
                void* out_pre, float* lse, int B, int heads, int Nq, int Nk, float scale, int residual, cudaStream_t stream) {
  PMV_CHECK_ARG(kd == 128 || kd == 160, "attention(tc): kd must be 128 or 160 (got %d)", kd);
  PMV_CHECK_ARG(ld_qk % 8 == 0 && ld_v % 8 == 0, "attention(tc): row strides must be multiples of 8 elements");
  PMV_CHECK_ARG(((uintptr_t)q_aug & 15) == 0 && ((uintptr_t)k_aug & 15) == 0 && ((uintptr_t)v & 15) == 0 && ((uintptr_t)out & 15) == 0,
                "attention(tc): operands must be 16-byte aligned");
  TcAttnGeom g{B, heads, Nq, Nk, scale, residual};
  if (kd == 128) return launch<128>(q_aug, k_aug, ld_qk, v, ld_v, out, out_pre, lse, g, stream);
  return launch<160>(q_aug, k_aug, ld_qk, v, ld_v, out, out_pre, lse, g, stream);
}
