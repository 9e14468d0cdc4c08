// attn_tc.cuh -- windowed relative-position multi-head self-attention (attentions.py:165-196) on the 5th-generation
// tensor cores (tcgen05 + TMEM + TMA), sm_100a only.  One CTA = 128 query rows of one head of one utterance.
//
//   S  = Q K^T            128 x 64 key tile, fp32 in TMEM                       (attentions.py:172, scores)
//   Sr = Q Ek^T           128 x 16 (9 relative offsets, zero padded), once      (:173-177, rel_logits before the skew)
//   s_ij = (S_ij + [|j-i|<=W] Sr_i[j-i+W]) / sqrt(dk);  keys >= len masked      (:178,183: -1e4 fill == exp -> 0 in fp32)
//   P  = exp(s - m)       online softmax, running max refreshed lazily          (:190 softmax)
//   O += P V + Pband Ev   128 x dk, fp32 in TMEM                                (:192-196, output + relative values)
//   out = O / l
//
// Operand format = the split-bf16 planes of the tensor-core convs: x = hi + lo to ~2^-18, three MMAs per K16 slice
// (lo*hi + hi*lo + hi*hi).  Q, K: K-major 128B-swizzled TMA tiles of the qkv planes (channels-last, so a head is a
// channel range and no transposition is needed).  V is consumed as an MN-major B operand straight from the same
// channels-last tiles.  P never leaves the SM: the softmax warps write it back into TMEM as packed bf16 (hi and lo
// planes, tcgen05.st) and the P V MMAs take their A operand from TMEM.  The relative-position terms ride on the same
// tensor-core path: Q Ek^T is one extra N=16 MMA group per CTA, and P_band Ev one extra K=16 MMA per diagonal key tile.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread MMA issuer,
// warps 2..5 = softmax (one query row per thread, no shuffles), O rescaling and the epilogue.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vtts {

constexpr int ATC_BM = 128;        // query rows per CTA (UMMA M)
constexpr int ATC_KT = 64;         // keys per tile
constexpr int ATC_NS = 2;          // K/V ring depth
constexpr int ATC_THREADS = 192;
constexpr int ATC_RELP = 16;       // relative-offset slots (2W+1 <= 16)
constexpr int ATC_RS = 13;         // floats per row of the per-row band scratch in shared memory
constexpr float ATC_LAZY = 6.0f;   // the running max is refreshed only when a tile exceeds it by more than this (natural log units)

// TMEM columns (fp32 cells): S double buffer, P double buffer (bf16 pairs: hi 32 cols + lo 32 cols), O, Sr, Pband (hi 8 + lo 8) x 2
constexpr int ATC_COL_S = 0, ATC_COL_P = 128, ATC_COL_O = 256, ATC_COL_SR = 384, ATC_COL_PB = 400, ATC_TMEM_COLS = 512;

struct AttnTcParams {
  CUtensorMap q_hi, q_lo;      // qkv planes [rows][ld], box 64 channels x 128 rows
  CUtensorMap kv_hi, kv_lo;    // same planes, box 64 channels x 64 rows
  CUtensorMap rk_hi, rk_lo;    // relative-key table   [16][128] bf16 (rows >= 2W+1 and channels >= dk are zero), box 64 x 16
  CUtensorMap rv_hi, rv_lo;    // relative-value table [16][128]
  float* out;                  // fp32 rows [.][ldo] (or null)
  __nv_bfloat16* p_hi;         // split-bf16 planes of the output for a tensor-core consumer (or null)
  __nv_bfloat16* p_lo;
  int ldo, ldp;
  int n_heads, window;
  int koff, voff;              // channel offsets of the K and V sections inside a qkv row (H, 2H)
};

constexpr int atc_chunks(int dk) { return (dk + 63) / 64; }
constexpr int atc_smem_bytes(int dk) {
  return 2 * atc_chunks(dk) * ATC_BM * 128                 // Q hi/lo
         + ATC_NS * 2 * 2 * atc_chunks(dk) * ATC_KT * 128  // K and V hi/lo per stage
         + 2 * 2 * atc_chunks(dk) * ATC_RELP * 128         // relative key / value tables hi/lo
         + 2 * ATC_BM * ATC_RS * 4                         // per-row band scratch (Sr values, band P values)
         + 1024 /*alignment slack*/ + 256 /*barriers*/;
}

__device__ __forceinline__ uint32_t umma_idesc_bf16_ex(int n, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(b_mn_major & 1) << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(ATC_BM >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ void sts_f32(uint32_t saddr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(saddr), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 a, __nv_bfloat16 b) {
  return (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
}

template <int DK>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_tc_kernel(const __grid_constant__ AttnTcParams ap, const int* __restrict__ lens, const int* __restrict__ offs) {
  static_assert(DK % 32 == 0 && DK >= 32 && DK <= 128, "head dim must be 32/64/96/128");
  constexpr int NC = atc_chunks(DK);                 // 64-channel chunks of a head (the last one may be half used)
  constexpr int LASTW = DK - 64 * (NC - 1);          // channels used of the last chunk
  constexpr int Q_TILE = ATC_BM * 128, KV_TILE = ATC_KT * 128, R_TILE = ATC_RELP * 128;
  constexpr int STAGE_BYTES = 2 * 2 * NC * KV_TILE;  // K hi/lo + V hi/lo
  PDL_LAUNCH();
  const int b = blockIdx.z, head = blockIdx.y;
  const int q0 = blockIdx.x * ATC_BM;
  // lens/offs are final before the graph that contains this kernel starts (host copies or the previous phase): an idle
  // CTA leaves before it allocates anything
  const int len = lens[b];
  if (q0 >= len) return;
  const long base = offs[b];
  const int W = ap.window, nrel = 2 * W + 1;

  extern __shared__ uint8_t atc_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(atc_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem;                                   // [plane][chunk][128 x 128 B]
  uint8_t* sKV = sQ + 2 * NC * Q_TILE;                  // [stage][K: plane, chunk | V: plane, chunk][64 x 128 B]
  uint8_t* sRK = sKV + ATC_NS * STAGE_BYTES;            // [plane][chunk][16 x 128 B]
  uint8_t* sRV = sRK + 2 * NC * R_TILE;
  float* sSr = reinterpret_cast<float*>(sRV + 2 * NC * R_TILE);   // [128][RS]  q_i . Ek[m]
  float* sPb = sSr + ATC_BM * ATC_RS;                             // [128][RS]  band probabilities of the current tile
  uint64_t* bars = reinterpret_cast<uint64_t*>(sPb + ATC_BM * ATC_RS);
  uint64_t* q_full = bars;              // Q tile + relative tables landed
  uint64_t* kv_full = bars + 1;         // [NS]
  uint64_t* kv_empty = kv_full + ATC_NS;
  uint64_t* s_full = kv_empty + ATC_NS; // [2] S tile complete in TMEM
  uint64_t* p_full = s_full + 2;        // [2] P tile written by all 128 softmax threads
  uint64_t* pv_done = p_full + 2;       // the P V MMAs of a tile have retired
  uint64_t* sr_full = pv_done + 1;      // Sr complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sr_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nt = (len + ATC_KT - 1) / ATC_KT;

  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < ATC_NS; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_full[s], 128); }
    mbar_init(pv_done, 1);
    mbar_init(sr_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ap.q_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ap.q_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ap.kv_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&ap.kv_lo)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)ATC_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      // the relative-position tables are constants: requested before the dependency wait
      mbar_expect_tx(q_full, (uint32_t)(2 * NC * Q_TILE + 4 * NC * R_TILE));
      for (int c = 0; c < NC; ++c) {
        tma_load_2d(sRK + c * R_TILE, &ap.rk_hi, c * 64, 0, q_full);
        tma_load_2d(sRK + (NC + c) * R_TILE, &ap.rk_lo, c * 64, 0, q_full);
        tma_load_2d(sRV + c * R_TILE, &ap.rv_hi, c * 64, 0, q_full);
        tma_load_2d(sRV + (NC + c) * R_TILE, &ap.rv_lo, c * 64, 0, q_full);
      }
      PDL_WAIT();
      timeline_stamp_t(-31);
      const int qch = head * DK;
      for (int c = 0; c < NC; ++c) {
        tma_load_2d(sQ + c * Q_TILE, &ap.q_hi, qch + c * 64, (int)base + q0, q_full);
        tma_load_2d(sQ + (NC + c) * Q_TILE, &ap.q_lo, qch + c * 64, (int)base + q0, q_full);
      }
      for (int t = 0; t < nt; ++t) {
        const int st = t % ATC_NS;
        if (t >= ATC_NS) mbar_wait(&kv_empty[st], ((t / ATC_NS) - 1) & 1);
        uint8_t* kb = sKV + st * STAGE_BYTES;
        uint8_t* vb = kb + 2 * NC * KV_TILE;
        mbar_expect_tx(&kv_full[st], (uint32_t)STAGE_BYTES);
        const int row = (int)base + t * ATC_KT;
        for (int c = 0; c < NC; ++c) {
          tma_load_2d(kb + c * KV_TILE, &ap.kv_hi, ap.koff + qch + c * 64, row, &kv_full[st]);
          tma_load_2d(kb + (NC + c) * KV_TILE, &ap.kv_lo, ap.koff + qch + c * 64, row, &kv_full[st]);
        }
        for (int c = 0; c < NC; ++c) {
          tma_load_2d(vb + c * KV_TILE, &ap.kv_hi, ap.voff + qch + c * 64, row, &kv_full[st]);
          tma_load_2d(vb + (NC + c) * KV_TILE, &ap.kv_lo, ap.voff + qch + c * 64, row, &kv_full[st]);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (one thread)
    if (lane == 0) {
      PDL_WAIT();
      const uint32_t id_s = umma_idesc_bf16_ex(ATC_KT, 0), id_r = umma_idesc_bf16_ex(ATC_RELP, 0);
      // S[buf] = Q K(t)^T : three MMAs per K16 slice
      auto issue_qk = [&](int t) {
        const int st = t % ATC_NS;
        mbar_wait(&kv_full[st], (t / ATC_NS) & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + ATC_COL_S + (t & 1) * ATC_KT;
        const uint32_t kb = smem_u32(sKV + st * STAGE_BYTES);
        uint32_t acc = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint64_t qh = umma_desc_sw128(smem_u32(sQ + c * Q_TILE)), ql = umma_desc_sw128(smem_u32(sQ + (NC + c) * Q_TILE));
          const uint64_t kh = umma_desc_sw128(kb + c * KV_TILE), kl = umma_desc_sw128(kb + (NC + c) * KV_TILE);
          const int nk = (c == NC - 1 ? LASTW : 64) / 16;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kk < nk) {
              const uint64_t adv = (uint64_t)((kk * 32) >> 4);
              umma_bf16(d, ql + adv, kh + adv, id_s, acc);
              umma_bf16(d, qh + adv, kl + adv, id_s, 1u);
              umma_bf16(d, qh + adv, kh + adv, id_s, 1u);
              acc = 1u;
            }
          }
        }
        umma_commit(&s_full[t & 1]);
      };
      mbar_wait(q_full, 0);
      timeline_stamp_t(-32);
      tc_fence_after();
      {   // Sr = Q Ek^T (N = 16)
        const uint32_t d = tmem_base + ATC_COL_SR;
        uint32_t acc = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint64_t qh = umma_desc_sw128(smem_u32(sQ + c * Q_TILE)), ql = umma_desc_sw128(smem_u32(sQ + (NC + c) * Q_TILE));
          const uint64_t eh = umma_desc_sw128(smem_u32(sRK + c * R_TILE)), el = umma_desc_sw128(smem_u32(sRK + (NC + c) * R_TILE));
          const int nk = (c == NC - 1 ? LASTW : 64) / 16;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kk < nk) {
              const uint64_t adv = (uint64_t)((kk * 32) >> 4);
              umma_bf16(d, ql + adv, eh + adv, id_r, acc);
              umma_bf16(d, qh + adv, el + adv, id_r, 1u);
              umma_bf16(d, qh + adv, eh + adv, id_r, 1u);
              acc = 1u;
            }
          }
        }
        umma_commit(sr_full);
      }
      issue_qk(0);
      for (int t = 0; t < nt; ++t) {
        if (t + 1 < nt) issue_qk(t + 1);            // overlaps the softmax of tile t
        mbar_wait(&p_full[t & 1], (t >> 1) & 1);
        tc_fence_after();
        const int st = t % ATC_NS;
        const uint32_t vb = smem_u32(sKV + st * STAGE_BYTES + 2 * NC * KV_TILE);
        const uint32_t pa = tmem_base + ATC_COL_P + (t & 1) * 64;       // hi: +0..31, lo: +32..63 (packed bf16 pairs)
        const int k0 = t * ATC_KT;
        const bool band = (k0 + ATC_KT - 1 + W >= q0) && (k0 - W <= q0 + ATC_BM - 1);
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int n = (c == NC - 1) ? LASTW : 64;
          const uint32_t id_v = umma_idesc_bf16_ex(n, 1);
          const uint32_t d = tmem_base + ATC_COL_O + c * 64;
          const uint64_t vh = umma_desc_sw128(vb + c * KV_TILE), vl = umma_desc_sw128(vb + (NC + c) * KV_TILE);
#pragma unroll
          for (int kk = 0; kk < ATC_KT / 16; ++kk) {
            const uint64_t adv = (uint64_t)((kk * 16 * 128) >> 4);      // 16 keys = 16 rows of 128 bytes
            umma_bf16_ts(d, pa + 32 + kk * 8, vh + adv, id_v, (t | kk) ? 1u : 0u);
            umma_bf16_ts(d, pa + kk * 8, vl + adv, id_v, 1u);
            umma_bf16_ts(d, pa + kk * 8, vh + adv, id_v, 1u);
          }
          if (band) {      // + P_band Ev (K = 16 relative offsets)
            const uint32_t pb = tmem_base + ATC_COL_PB + (t & 1) * 16;   // hi 8 cols, lo 8 cols
            const uint64_t eh = umma_desc_sw128(smem_u32(sRV + c * R_TILE)), el = umma_desc_sw128(smem_u32(sRV + (NC + c) * R_TILE));
            umma_bf16_ts(d, pb + 8, eh, id_v, 1u);
            umma_bf16_ts(d, pb, el, id_v, 1u);
            umma_bf16_ts(d, pb, eh, id_v, 1u);
          }
        }
        umma_commit(&kv_empty[st]);
        umma_commit(pv_done);
      }
    }
  } else {
    // ------------------------------------------------------------------ softmax / correction / epilogue: one row per thread
    const int quad = warp & 3;
    const int row = quad * 32 + lane;                // row of the query tile == TMEM lane
    const int qi = q0 + row;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    const float scale = 1.0f / sqrtf((float)DK);
    const float L2E = 1.4426950408889634f;
    const uint32_t mySr = smem_u32(sSr + row * ATC_RS);     // (explicit shared-space accesses: generic LD/ST otherwise)
    const uint32_t myPb = smem_u32(sPb + row * ATC_RS);
    PDL_WAIT();
    {   // relative-key logits of this row -> shared memory (indexed by a run-time offset below)
      mbar_wait(sr_full, 0);
      tc_fence_after();
      uint32_t r[16];
      tmem_ld_x16(tmem_base + lane_sel + ATC_COL_SR, r);
      tmem_wait_ld();
#pragma unroll
      for (int m = 0; m < ATC_RS; ++m) sts_f32(mySr + 4 * m, __uint_as_float(r[m]) * scale);
    }
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < nt; ++t) {
      const int bsel = t & 1;
      const int k0 = t * ATC_KT;
      mbar_wait(&s_full[bsel], (t >> 1) & 1);
      if (threadIdx.x == 64) timeline_stamp_t(-40 - t);
      tc_fence_after();
      float s[ATC_KT];
      {
        uint32_t r0[16], r1[16], r2[16], r3[16];           // all four loads in flight, one wait
        const uint32_t sa = tmem_base + lane_sel + ATC_COL_S + bsel * ATC_KT;
        tmem_ld_x16(sa, r0); tmem_ld_x16(sa + 16, r1); tmem_ld_x16(sa + 32, r2); tmem_ld_x16(sa + 48, r3);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          s[i] = __uint_as_float(r0[i]) * scale; s[16 + i] = __uint_as_float(r1[i]) * scale;
          s[32 + i] = __uint_as_float(r2[i]) * scale; s[48 + i] = __uint_as_float(r3[i]) * scale;
        }
      }
      // does the +-W band of any row of this CTA / of this warp cross the key tile?
      const bool band_cta = (k0 + ATC_KT - 1 + W >= q0) && (k0 - W <= q0 + ATC_BM - 1);
      const int w_lo = q0 + quad * 32, w_hi = w_lo + 31;
      const bool band_warp = (k0 + ATC_KT - 1 + W >= w_lo) && (k0 - W <= w_hi);
      const int moff = k0 - qi + W;                  // relative slot of key column c is c + moff
      if (band_warp) {
#pragma unroll
        for (int c = 0; c < ATC_KT; ++c) {
          const int m = c + moff;
          if ((unsigned)m < (unsigned)nrel) s[c] += lds_f32(mySr + 4 * m);
        }
      }
      const int kvalid = len - k0;                   // columns >= kvalid are beyond the utterance
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < ATC_KT; ++c) {
        if (c >= kvalid) s[c] = -INFINITY;
        mx = fmaxf(mx, s[c]);
      }
      // lazily refreshed running max: O and l are rescaled only when the tile max exceeds it by more than ATC_LAZY
      float alpha = 1.f;
      bool need = false;
      if (t == 0) {
        m_run = mx;
      } else if (mx > m_run + ATC_LAZY) {
        alpha = exp2f((m_run - mx) * L2E);
        m_run = mx;
        need = true;
      }
      // pv_done completes one phase per key tile and a parity wait can only tell the current phase from the one before it:
      // every thread therefore observes EVERY phase, in order -- here when O has to be rescaled, otherwise just before this
      // tile's P is handed over (by then the P V MMAs of tile t-1 have normally retired, so the wait costs nothing)
      bool pv_seen = (t == 0);
      if (__any_sync(0xffffffffu, need)) {           // warp-uniform: tcgen05.ld/st are warp collectives
        mbar_wait(pv_done, (t - 1) & 1);             // the P V MMAs of tile t-1 have retired: O is quiescent
        pv_seen = true;
        tc_fence_after();
        l_run *= alpha;
#pragma unroll
        for (int n0 = 0; n0 < DK; n0 += 16) {
          uint32_t r[16];
          tmem_ld_x16(tmem_base + lane_sel + ATC_COL_O + n0, r);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
          tmem_st_x16(tmem_base + lane_sel + ATC_COL_O + n0, r);
        }
      }
      const float mneg = -m_run * L2E;
      float lsum = 0.f;
      if (band_cta) {
#pragma unroll
        for (int m = 0; m < ATC_RS; ++m) sts_f32(myPb + 4 * m, 0.f);
      }
      {
        uint32_t ph[ATC_KT / 2], pl[ATC_KT / 2];
#pragma unroll
        for (int c = 0; c < ATC_KT; c += 2) {
          const float p0 = exp2f(fmaf(s[c], L2E, mneg)), p1 = exp2f(fmaf(s[c + 1], L2E, mneg));
          lsum += p0 + p1;
          if (band_warp) {
            const int m0 = c + moff, m1 = c + 1 + moff;
            if ((unsigned)m0 < (unsigned)nrel) sts_f32(myPb + 4 * m0, p0);
            if ((unsigned)m1 < (unsigned)nrel) sts_f32(myPb + 4 * m1, p1);
          }
          split_bf16_pair(p0, p1, ph[c >> 1], pl[c >> 1]);
        }
        l_run += lsum;
        const uint32_t pa = tmem_base + lane_sel + ATC_COL_P + bsel * 64;
#pragma unroll
        for (int n0 = 0; n0 < ATC_KT / 2; n0 += 16) {
          uint32_t r[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = ph[n0 + i];
          tmem_st_x16(pa + n0, r);
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = pl[n0 + i];
          tmem_st_x16(pa + 32 + n0, r);
        }
      }
      if (band_cta) {
        uint32_t bh[8], bl[8];
#pragma unroll
        for (int m = 0; m < 16; m += 2) {
          const float p0 = m < ATC_RS ? lds_f32(myPb + 4 * m) : 0.f, p1 = m + 1 < ATC_RS ? lds_f32(myPb + 4 * (m + 1)) : 0.f;
          split_bf16_pair(p0, p1, bh[m >> 1], bl[m >> 1]);
        }
        const uint32_t pb = tmem_base + lane_sel + ATC_COL_PB + bsel * 16;
        tmem_st_x8(pb, bh);
        tmem_st_x8(pb + 8, bl);
      }
      if (!pv_seen) mbar_wait(pv_done, (t - 1) & 1);
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(&p_full[bsel]);
      if (threadIdx.x == 64) timeline_stamp_t(-60 - t);
    }
    // ---- epilogue: O / l -> fp32 rows and/or split-bf16 planes
    mbar_wait(pv_done, (nt - 1) & 1);
    if (threadIdx.x == 64) timeline_stamp_t(-36);
    tc_fence_after();
    const float inv = 1.f / l_run;
    const bool rowok = qi < len;
    const long orow = base + qi;
#pragma unroll
    for (int n0 = 0; n0 < DK; n0 += 32) {
      uint32_t ra[16], rb[16];
      tmem_ld_x16(tmem_base + lane_sel + ATC_COL_O + n0, ra);
      tmem_ld_x16(tmem_base + lane_sel + ATC_COL_O + n0 + 16, rb);
      tmem_wait_ld();
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
      const uint32_t (&r)[16] = hh ? rb : ra;
      const int nb = n0 + 16 * hh;
      if (rowok) {
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]) * inv;
        if (ap.out) {
          float* o = ap.out + orow * (long)ap.ldo + head * DK + nb;
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(o + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        if (ap.p_hi) {
          __align__(16) __nv_bfloat16 hb[16], lb[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) split_bf16(v[i], hb[i], lb[i]);
          __nv_bfloat16* ph = ap.p_hi + orow * (long)ap.ldp + head * DK + nb;
          __nv_bfloat16* pl = ap.p_lo + orow * (long)ap.ldp + head * DK + nb;
          *reinterpret_cast<uint4*>(ph) = *reinterpret_cast<const uint4*>(hb);
          *reinterpret_cast<uint4*>(ph + 8) = *reinterpret_cast<const uint4*>(hb + 8);
          *reinterpret_cast<uint4*>(pl) = *reinterpret_cast<const uint4*>(lb);
          *reinterpret_cast<uint4*>(pl + 8) = *reinterpret_cast<const uint4*>(lb + 8);
        }
      }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 64) timeline_stamp_t(-37);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)ATC_TMEM_COLS) : "memory");
  }
}

}  // namespace vtts
