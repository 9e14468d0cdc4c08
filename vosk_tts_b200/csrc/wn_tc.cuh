// wn_tc.cuh -- one WaveNet layer of the coupling flows (modules.py:155-175) as ONE tcgen05 kernel, sm_100a only:
//
//   a    = conv_k(x) + bias + cond[b]                      gated dilated conv, H -> 2H channels     (in_layers[i], cond_layer)
//   acts = tanh(a_t) * sigmoid(a_s)                         fused_add_tanh_sigmoid_multiply          (commons.py:100-107)
//   rs   = W_rs acts + b_rs                                 1x1 conv, H -> 2H (last layer: H)         (res_skip_layers[i])
//   x    = (x + rs[:H]) * mask ;  skip += rs[H:]            (last layer: skip += rs)
//
// The separate kernels (gated conv, then 1x1) hand `acts` over through HBM/L2 and pay a kernel boundary in between -- at
// batch 1 that is ~10 us of a ~22 us layer (16 layers per utterance).  Here a thread-block cluster of WN_NCL = 4 CTAs owns one
// 128-row tile: CTA c computes 2H/4 interleaved gate columns of the first GEMM (tcgen05, split-bf16 operands, accumulator
// in TMEM), gates them into H/4 channels of `acts`, and ALL-GATHERS that slice -- already split into bf16 hi/lo and laid
// out as a 128-byte-swizzled K-major UMMA operand -- into the shared memory of all four CTAs through distributed shared
// memory (st.shared::cluster), where it aliases the drained activation ring.  Each CTA then runs its quarter of the
// second GEMM straight from that tile (second TMEM accumulator) and finishes it: residual update of x (fp32 + the planes
// the next layer's TMA reads) or the skip accumulation.  `acts` never touches global memory.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vtts {

constexpr int WN_NCL = 4;          // CTAs of a cluster (along the output channels)
constexpr int WN_THREADS = 192;    // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogues
constexpr int WN_AST = 3, WN_WST = 4;

struct WnParams {
  CUtensorMap a_hi, a_lo;          // planes of x [rows][H], box 64 x 128
  CUtensorMap win_hi, win_lo;      // in-conv weights [k][2H][H] (gate-interleaved rows), box 64 x (2H/4)
  CUtensorMap wrx_hi, wrx_lo;      // res weights  [H][H] (rows = output channels), box 64 x BN2   (unused on the last layer)
  CUtensorMap wrs_hi, wrs_lo;      // skip weights [H][H], box 64 x BN2
  const float* bias_in;            // [2H] interleaved
  const float* cond;               // per-utterance conditioning rows [B][cond_ld] (or null)
  const float* bias_rx;            // [H] (null on the last layer)
  const float* bias_rs;            // [H]
  float* x;                        // hidden state rows [.][H], updated in place (unused on the last layer)
  __nv_bfloat16* xp_hi;            // planes of the updated x for the next layer
  __nv_bfloat16* xp_lo;
  float* skip;                     // skip accumulator rows [.][H]
  __nv_bfloat16* sp_hi;            // planes of skip (last layer: input of the coupling layer's post conv) or null
  __nv_bfloat16* sp_lo;
  int cond_ld, k, dil, pad;
  int first;                       // 1: skip = rs (no accumulation)
};

template <int H>
constexpr int wn_smem_bytes() {
  return WN_AST * 2 * 128 * 128 + WN_WST * 2 * (2 * H / WN_NCL) * 128 + 1024 + 256 + (2 * H / WN_NCL) * 4 + 128 * 4;
}

template <int H, bool LAST>
__global__ void __launch_bounds__(WN_THREADS, 1)
wn_layer_tc_kernel(const __grid_constant__ WnParams wp, const int* __restrict__ lens, const int* __restrict__ offs) {
  static_assert(H % 64 == 0 && H / 64 <= WN_AST && (H / WN_NCL) % 16 == 0, "hidden width not supported by the fused WN layer");
  constexpr int BN1 = 2 * H / WN_NCL;                 // gate columns of the first GEMM per CTA (96 for H = 192)
  constexpr int NA = BN1 / 2;                         // acts channels this CTA produces (48)
  constexpr int BN2 = (LAST ? H : 2 * H) / WN_NCL;    // columns of the second GEMM per CTA (96; last layer 48)
  constexpr int NCH = H / 64;                         // 64-channel chunks of x / acts
  constexpr int A_BYTES = 128 * 128;                  // one activation plane tile
  constexpr int W_BYTES = BN1 * 128;                  // one in-conv weight plane tile
  static_assert(2 * NCH * A_BYTES <= WN_AST * 2 * A_BYTES, "acts tile must fit the activation ring");
  static_assert(NCH <= WN_WST && BN2 <= BN1, "res/skip weight tiles must fit the weight ring slots");
  PDL_LAUNCH();
  const int b = blockIdx.z;
  const int t0 = blockIdx.x * 128;
  const uint32_t crank = cluster_rank();              // == blockIdx.y (cluster dims (1, 4, 1))

  extern __shared__ uint8_t wn_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wn_smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_w = smem + WN_AST * 2 * A_BYTES;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem_w + WN_WST * 2 * W_BYTES);
  uint64_t* a_empty = a_full + WN_AST;
  uint64_t* w_full = a_empty + WN_AST;
  uint64_t* w_empty = w_full + WN_WST;
  uint64_t* d1_full = w_empty + WN_WST;
  uint64_t* d2_full = d1_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d2_full + 1);
  float* bias1_s = reinterpret_cast<float*>(tmem_slot + 4);      // [BN1]
  float* bias2_s = bias1_s + BN1;                                // [BN2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsteps = NCH * wp.k;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < WN_AST; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < WN_WST; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
    mbar_init(d1_full, 1);
    mbar_init(d2_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wp.a_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wp.a_lo)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wp.win_hi)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&wp.win_lo)) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto issue_w = [&](int s) {      // in-conv weight tile of k-step s = (chunk c, tap j)
    const int c = s / wp.k, j = s - c * wp.k;
    const int wst = s % WN_WST;
    uint8_t* wb = smem_w + wst * 2 * W_BYTES;
    mbar_expect_tx(&w_full[wst], 2 * W_BYTES);
    tma_load_2d(wb, &wp.win_hi, c * 64, j * 2 * H + (int)crank * BN1, &w_full[wst]);
    tma_load_2d(wb + W_BYTES, &wp.win_lo, c * 64, j * 2 * H + (int)crank * BN1, &w_full[wst]);
  };
  const int w_pre = min(WN_WST, nsteps);
  if (warp == 0 && lane == 0)
    for (int i = 0; i < w_pre; ++i) issue_w(i);        // weights are constants: requested before the dependency wait
  PDL_WAIT();
  const int L = lens[b];
  const long base = offs[b];
  const bool active = t0 < L;                          // uniform over the cluster (its CTAs share the row tile)

  if (!active) {
    // nothing to compute, and no peer will touch this CTA: only the prefetched weight tiles have to land before it exits
    if (warp == 0 && lane == 0)
      for (int i = 0; i < w_pre; ++i) mbar_wait(&w_full[i], 0);
  } else if (warp == 0) {
    // -------------------------------------------------------------------- TMA producer
    if (lane == 0) {
      for (int s = 0; s < nsteps; ++s) {
        const int c = s / wp.k, j = s - c * wp.k;
        const int ast = s % WN_AST, use = s / WN_AST;
        if (use > 0) mbar_wait(&a_empty[ast], (use - 1) & 1);
        uint8_t* ab = smem + ast * 2 * A_BYTES;
        mbar_expect_tx(&a_full[ast], 2 * A_BYTES);
        const int row = (int)base + t0 - wp.pad + j * wp.dil;
        tma_load_2d(ab, &wp.a_hi, c * 64, row, &a_full[ast]);
        tma_load_2d(ab + A_BYTES, &wp.a_lo, c * 64, row, &a_full[ast]);
        if (s >= w_pre) {
          const int wst = s % WN_WST, wuse = s / WN_WST;
          mbar_wait(&w_empty[wst], (wuse - 1) & 1);
          issue_w(s);
        }
      }
      // res/skip weights into the drained weight ring (chunk c -> slot c): requested as soon as the first GEMM has retired,
      // they land while the epilogue warps gate and gather
      mbar_wait(d1_full, 0);
      const bool resx = !LAST && (int)crank * BN2 < H;            // this CTA's columns are the residual half
      const CUtensorMap* mh = resx ? &wp.wrx_hi : &wp.wrs_hi;
      const CUtensorMap* ml = resx ? &wp.wrx_lo : &wp.wrs_lo;
      const int r0 = ((int)crank * BN2) % H;
      for (int c = 0; c < NCH; ++c) {
        uint8_t* wb = smem_w + c * 2 * W_BYTES;
        mbar_expect_tx(&w_full[c], 2u * (uint32_t)BN2 * 128u);
        tma_load_2d(wb, mh, c * 64, r0, &w_full[c]);
        tma_load_2d(wb + W_BYTES, ml, c * 64, r0, &w_full[c]);
      }
    }
    __syncwarp();
    cluster_sync_all();      // (1)
    cluster_sync_all();      // (2)
  } else if (warp == 1) {
    // -------------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(BN1);
      for (int s = 0; s < nsteps; ++s) {
        const int ast = s % WN_AST, wst = s % WN_WST;
        mbar_wait(&a_full[ast], (s / WN_AST) & 1);
        mbar_wait(&w_full[wst], (s / WN_WST) & 1);
        tc_fence_after();
        const uint32_t abase = smem_u32(smem + ast * 2 * A_BYTES), wbase = smem_u32(smem_w + wst * 2 * W_BYTES);
        const uint64_t ahi = umma_desc_sw128(abase), alo = umma_desc_sw128(abase + A_BYTES);
        const uint64_t bhi = umma_desc_sw128(wbase), blo = umma_desc_sw128(wbase + W_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t adv = (uint64_t)((kk * 32) >> 4);
          umma_bf16(tmem_base, alo + adv, bhi + adv, idesc, (s | kk) ? 1u : 0u);
          umma_bf16(tmem_base, ahi + adv, blo + adv, idesc, 1u);
          umma_bf16(tmem_base, ahi + adv, bhi + adv, idesc, 1u);
        }
        umma_commit(&w_empty[wst]);
        umma_commit(&a_empty[ast]);
      }
      umma_commit(d1_full);
    }
    __syncwarp();
    cluster_sync_all();      // (1)
    cluster_sync_all();      // (2) the acts tile of the whole cluster sits in this CTA's activation ring
    if (lane == 0) {
      asm volatile("fence.proxy.async;" ::: "memory");              // generic-proxy (DSMEM) writes -> tensor-core reads
      const uint32_t idesc2 = umma_idesc_bf16(BN2);
      const uint32_t d2 = tmem_base + 128;
      const uint32_t act = smem_u32(smem);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int uses = (nsteps - c + WN_WST - 1) / WN_WST;         // times slot c was filled by the first GEMM
        mbar_wait(&w_full[c], uses & 1);
        tc_fence_after();
        const uint32_t wbase = smem_u32(smem_w + c * 2 * W_BYTES);
        const uint64_t ahi = umma_desc_sw128(act + c * A_BYTES), alo = umma_desc_sw128(act + (NCH + c) * A_BYTES);
        const uint64_t bhi = umma_desc_sw128(wbase), blo = umma_desc_sw128(wbase + W_BYTES);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t adv = (uint64_t)((kk * 32) >> 4);
          umma_bf16(d2, alo + adv, bhi + adv, idesc2, (c | kk) ? 1u : 0u);
          umma_bf16(d2, ahi + adv, blo + adv, idesc2, 1u);
          umma_bf16(d2, ahi + adv, bhi + adv, idesc2, 1u);
        }
      }
      umma_commit(d2_full);
    }
  } else {
    // -------------------------------------------------------------------- epilogue warps: one tile row per thread
    const int quad = warp & 3;
    const int row = quad * 32 + lane;
    const int et = threadIdx.x - 64;
    const uint32_t lane_sel = (uint32_t)(quad * 32) << 16;
    for (int i = et; i < BN1; i += 128) {
      const int cc = (int)crank * BN1 + i;
      float bv = wp.bias_in[cc];
      if (wp.cond) bv += wp.cond[(long)b * wp.cond_ld + cc];
      bias1_s[i] = bv;
    }
    for (int i = et; i < BN2; i += 128) {
      const int g = (int)crank * BN2 + i;
      bias2_s[i] = LAST ? wp.bias_rs[g] : (g < H ? wp.bias_rx[g] : wp.bias_rs[g - H]);
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 64) timeline_stamp_t(-71);
    mbar_wait(d1_full, 0);
    if (threadIdx.x == 64) timeline_stamp_t(-72);
    tc_fence_after();
    // ---- gate: this row's BN1 interleaved (tanh, sigmoid) columns -> NA channels of acts as packed bf16 hi / lo pairs
    uint32_t ah[NA / 2], al[NA / 2];
#pragma unroll
    for (int n0 = 0; n0 < BN1; n0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + lane_sel + (uint32_t)n0, v);
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float t0v = v[i] + bias1_s[n0 + i], s0v = v[i + 1] + bias1_s[n0 + i + 1];
        const float t1v = v[i + 2] + bias1_s[n0 + i + 2], s1v = v[i + 3] + bias1_s[n0 + i + 3];
        const float a0 = tanhf(t0v) * (1.f / (1.f + expf(-s0v)));
        const float a1 = tanhf(t1v) * (1.f / (1.f + expf(-s1v)));
        split_bf16_pair(a0, a1, ah[(n0 + i) >> 2], al[(n0 + i) >> 2]);
      }
    }
    tc_fence_before();
    if (threadIdx.x == 64) timeline_stamp_t(-73);
    cluster_sync_all();      // (1) every CTA of the cluster has retired its first GEMM: the activation rings are free
    if (threadIdx.x == 64) timeline_stamp_t(-74);
    // ---- all-gather: this CTA's NA channels of this row into the acts tile of every CTA of the cluster.  Tile layout =
    // what TMA would produce for a K-major 128B-swizzled box: [plane][chunk of 64 channels][row][128 bytes], the 16-byte
    // unit j of a row stored at unit j ^ (row & 7)
    {
      const uint32_t local = smem_u32(smem);
#pragma unroll
      for (int dst = 0; dst < WN_NCL; ++dst) {
        uint32_t rbase;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbase) : "r"(local), "r"(dst));
#pragma unroll
        for (int u = 0; u < NA / 8; ++u) {
          const int U = (int)crank * (NA / 8) + u;                   // 8-channel unit index within the H acts channels
          const uint32_t off = (uint32_t)((U >> 3) * A_BYTES + row * 128 + (((U & 7) ^ (row & 7)) << 4));
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + off), "r"(ah[4 * u]), "r"(ah[4 * u + 1]),
                       "r"(ah[4 * u + 2]), "r"(ah[4 * u + 3])
                       : "memory");
          asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(rbase + (uint32_t)(NCH * A_BYTES) + off), "r"(al[4 * u]),
                       "r"(al[4 * u + 1]), "r"(al[4 * u + 2]), "r"(al[4 * u + 3])
                       : "memory");
        }
      }
    }
    asm volatile("fence.proxy.async;" ::: "memory");
    if (threadIdx.x == 64) timeline_stamp_t(-75);
    cluster_sync_all();      // (2) all slices have landed everywhere; nobody writes into a peer after this point
    if (threadIdx.x == 64) timeline_stamp_t(-76);
    // ---- second epilogue: residual update of x / skip accumulation for this CTA's BN2 columns
    const int t = t0 + row;
    const bool rowok = t < L;
    const long orow = base + t;
    const int g0 = (int)crank * BN2;                                  // first res_skip column of this CTA
    const bool resx = !LAST && g0 < H;
    float* dstf = resx ? (wp.x + orow * (long)H + g0) : (wp.skip + orow * (long)H + (LAST ? g0 : g0 - H));
    const bool addres = resx || !wp.first;
    __nv_bfloat16* ph = resx ? wp.xp_hi : wp.sp_hi;
    __nv_bfloat16* pl = resx ? wp.xp_lo : wp.sp_lo;
    const long poff = orow * (long)H + (resx || LAST ? g0 : g0 - H);
    mbar_wait(d2_full, 0);
    if (threadIdx.x == 64) timeline_stamp_t(-77);
    tc_fence_after();
#pragma unroll
    for (int n0 = 0; n0 < BN2; n0 += 16) {
      float v[16];
      tmem_ld16(tmem_base + lane_sel + 128u + (uint32_t)n0, v);
      if (rowok) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 r4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (addres) r4 = *reinterpret_cast<const float4*>(dstf + n0 + i);
          v[i] += bias2_s[n0 + i] + r4.x; v[i + 1] += bias2_s[n0 + i + 1] + r4.y;
          v[i + 2] += bias2_s[n0 + i + 2] + r4.z; v[i + 3] += bias2_s[n0 + i + 3] + r4.w;
          *reinterpret_cast<float4*>(dstf + n0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
        if (ph) {
          uint32_t hb[8], lb[8];
#pragma unroll
          for (int i = 0; i < 16; i += 2) split_bf16_pair(v[i], v[i + 1], hb[i >> 1], lb[i >> 1]);
          *reinterpret_cast<uint4*>(ph + poff + n0) = make_uint4(hb[0], hb[1], hb[2], hb[3]);
          *reinterpret_cast<uint4*>(ph + poff + n0 + 8) = make_uint4(hb[4], hb[5], hb[6], hb[7]);
          *reinterpret_cast<uint4*>(pl + poff + n0) = make_uint4(lb[0], lb[1], lb[2], lb[3]);
          *reinterpret_cast<uint4*>(pl + poff + n0 + 8) = make_uint4(lb[4], lb[5], lb[6], lb[7]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 64) timeline_stamp_t(-78);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

}  // namespace vtts
