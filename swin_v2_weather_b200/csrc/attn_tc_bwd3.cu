// tcgen05 windowed cosine attention, backward, third generation: ONE pass over the logits, warp-specialised.
//
// Reference math: swinv2_global.py:300-318 (cosine logits, clamped scale, CPB bias, shift mask, softmax, PV) differentiated.
//
// Work item = (sample, window, head), persistent CTAs (one per SM) loop over items.  16 warps (4 per scheduler, 128 registers
// each -- a 17th "control" warp would put 5 warps on one scheduler and cap every thread at 96 registers):
//   all warps    thread-per-key-row softmax gradient and the output epilogues
//   thread 0     additionally issues every TMA load and every tcgen05.mma of the CTA, at the points where its warp would
//                otherwise wait for the same mbarrier as everybody else; everything it issues is asynchronous
// Key-major orientation, so that each product is computed once (5 MMA chains per 128-key tile instead of the 7 of the
// two-sweep kernel, and one exp per logit instead of two):
//   S^T_u = K^_u Q^T , dP^T_u = V_u dO^T            (keys on TMEM lanes, queries on columns; fp32)
//   P^T_u = 2^(scale*S^T - lse_q) ; dS^T_u = P^T_u o (dP^T_u - D_q)           [compute warps; D = <dO, O> from the pre-pass]
//   dV_u  = P^T_u dO      A = P^T_u packed bf16 in TENSOR MEMORY (TS mode, never touches shared memory)
//   dK^_u = dS^T_u Q^     A = dS^T_u from shared memory, K-major
//   dQ^_t = dS_t K^       A = the SAME dS^T bytes read MN-major (rows = keys = k dimension): no transposed copy, no recompute
// The L2-normalisation Jacobians need <k^_j, dk^_j> = scale * sum_i dS_ij cos_ij -- exactly the per-thread partial sums the
// softmax phase already forms for d(logit_scale) -- so the dk epilogue needs no extra reduction; <q^_i, dq^_i> is reduced
// over the four column groups of a row through shared memory.
//
// TMEM columns (512):  S^T [0,176)  dP^T [176,352)  P^T bf16 [416,504);  once a tile's softmax is done its dV / dK^ accumulators
// reuse [192,288) / [288,384) and, after the last tile, dQ^_0 / dQ^_1 reuse [0,96) / [96,192).  S^T of the next key tile is issued
// while the compute warps are still draining dV / dK^ of the current one.
// Shared memory: dS^T 60.5 KB + four operand buffers (Q^, K^, V, dO; 33 KB each, 64B-swizzled TMA boxes as in the second
// generation).  Buffers are released by MMA completion, in the order V, (Q^, dO), K^; the next item's K^ is loaded into the
// buffer V just left (K^/V swap buffers every item), so its operands arrive while this item's tail is still computing.
#include <stdlib.h>
#include "attn_tc.cuh"

namespace swinb200 {

__device__ long long* g_phase_buf3 = nullptr;

template <int D>
struct Bwd3Smem {
  static constexpr int kDSCS = kMaxLP * 16;                 // dS^T tile: [22 chunks of 8 queries][176 key rows][16 B]
  static constexpr int kDS = (kMaxLP / 8) * kDSCS;          // 61,952
  static constexpr int kTile = (D / 32) * kCS64;            // one operand: 3 x [176 rows x 64 B]
  static constexpr int kOffDS = 0;                          // first: MMA over-reads past its end land in the operand buffers
  static constexpr int kOffOp = kDS;
  static constexpr int kOffTok = kOffOp + 4 * kTile;        // [2][176] token indices (current / next item)
  static constexpr int kOffLse = kOffTok + 2 * kMaxLP * 4;  // log2-domain LSE per query (+inf for pad queries)
  static constexpr int kOffDv = kOffLse + kMaxLP * 4;       // D = <dO, O> per query
  static constexpr int kOffDot = kOffDv + kMaxLP * 4;       // [2 tile parities][4 groups][128 rows]: sum_i dS_ij cos_ij partials
  static constexpr int kOffRed = kOffDot + 2 * 4 * 128 * 4; // [2 tile parities][4 groups][128 rows]: <q^, dQ^> partials
  static constexpr int kOffDsc = kOffRed + 2 * 4 * 128 * 4; // per-head d(scale) partial sums of this CTA
  static constexpr int kOffBar = kOffDsc + 128;
  static constexpr int kBytes = kOffBar + 128;
  static_assert(kDS % 512 == 0 && kTile % 512 == 0, "64B-swizzled operand tiles need 512-byte alignment");
  static_assert(kBytes <= 227 * 1024, "shared memory budget");
};

constexpr int kB3Compute = 512;                  // 16 warps
constexpr int kB3Threads = kB3Compute;
constexpr uint32_t kColST = 0, kColDPT = 176, kColDV = 192, kColDK = 288, kColDQ = 0;   // dQ^_1 ends at 192: dV starts there

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint4 pack8f(const float* v) {
  uint4 r;
  r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]);
  r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
  return r;
}

template <int D, bool kProf>
__global__ void __launch_bounds__(kB3Threads, 1)
attn_tc_bwd3_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __grid_constant__ CUtensorMap tm_do,
                    const float* __restrict__ Dpre, const __nv_bfloat16* __restrict__ qkv, const float* __restrict__ inv_norm,
                    const float* __restrict__ scale_p, const float* __restrict__ bias, const __nv_bfloat16* __restrict__ d_o,
                    const float* __restrict__ lse, __nv_bfloat16* __restrict__ dqkv, float* __restrict__ dscale,
                    float* __restrict__ dbias, const AttnGeom g, const uint32_t kColP) {
  using SM = Bwd3Smem<D>;
  static_assert(D == 96, "column split of the epilogues assumes head_dim 96 (4 groups x 24 columns)");
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr int kPieces = D / 8;                 // 16-byte pieces per operand row
  constexpr int kBoxes = D / 32;                 // TMA boxes (32-channel chunks) per operand
  constexpr int kEpiCols = D / 4;                // output columns per thread in the epilogues (24)
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* sDS = smem + SM::kOffDS;
  int* tokbuf0 = reinterpret_cast<int*>(smem + SM::kOffTok);
  float* lse2 = reinterpret_cast<float*>(smem + SM::kOffLse);
  float* Dv = reinterpret_cast<float*>(smem + SM::kOffDv);
  float* dotk = reinterpret_cast<float*>(smem + SM::kOffDot);
  float* red = reinterpret_cast<float*>(smem + SM::kOffRed);
  float* dsc_heads = reinterpret_cast<float*>(smem + SM::kOffDsc);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* full = bars;            // [4] operand (0 Q^, 1 K^, 2 V, 3 dO) of the next item to start has landed
  uint64_t* sbar = bars + 4;        // S^T / dP^T of a key tile are in tensor memory
  uint64_t* pbar = bars + 5;        // P^T (TMEM) and dS^T (smem) of the tile are written              (16 warps)
  uint64_t* obar = bars + 6;        // dV / dK^ accumulators of the tile are complete
  uint64_t* ebar = bars + 7;        // ... and have been read by the epilogue                            (16 warps)
  uint64_t* qbar = bars + 8;        // dQ^ accumulators are complete
  uint64_t* eqbar = bars + 9;       // ... and have been read                                            (16 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = g.L, LP = g.LP, C = g.C, C3 = 3 * g.C;
  const int ntiles = (LP > 128) ? 2 : 1;
  const bool shifted = (g.s0 > 0) || (g.s1 > 0);
  const int nitems = g.B * g.nW * g.heads;
  const int first = blockIdx.x;

  auto op_ptr = [&](int buf) { return smem + SM::kOffOp + buf * SM::kTile; };
  auto item_is_box = [&](int item) {   // a window that wraps around the cyclic shift is not one box of the tensor
    const int ww_all = (item / g.heads) % g.nW;
    const int wh = ww_all / g.nWw, ww = ww_all - wh * g.nWw;
    return !((g.s0 > 0 && (wh + 1) * g.Wh + g.s0 > g.H) || (g.s1 > 0 && (ww + 1) * g.Ww + g.s1 > g.W));
  };

  // ---- one-time set-up -----------------------------------------------------------------------------------------------
  if (tid == 0) {
    prefetch_tmap(&tm_qkv);
    prefetch_tmap(&tm_do);
    for (int i = 0; i < 4; ++i) mbar_init(&full[i], 1);
    mbar_init(sbar, 1);
    mbar_init(pbar, 16);
    mbar_init(obar, 1);
    mbar_init(ebar, 16);
    mbar_init(qbar, 1);
    mbar_init(eqbar, 16);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  // pad rows [L, LP) of the four operand buffers stay zero for the whole kernel (loads only ever write rows < L)
  for (int i = tid; i < (LP - L) * kPieces * 4; i += kB3Threads) {
    const int op = i / ((LP - L) * kPieces);
    const int rem = i - op * (LP - L) * kPieces;
    const int c = rem / (LP - L), r = L + rem % (LP - L);
    *reinterpret_cast<uint4*>(op_ptr(op) + opnd_off(r, c)) = make_uint4(0, 0, 0, 0);
  }
  if (tid < 32) dsc_heads[tid] = 0.f;
  auto fill_tok = [&](int item, int* tk) {          // compute threads
    const int ww = (item / g.heads) % g.nW;
    const int bb = item / (g.heads * g.nW);
    for (int n = tid; n < LP; n += kB3Compute) {
      int rr;
      tk[n] = (n < L) ? win_token(g, bb, ww, n, rr) : -1;
    }
  };
  auto fill_rows = [&](int item, const int* tk) {   // lse2 / Dv of `item` (compute threads)
    const int hd = item % g.heads;
    const int ww = (item / g.heads) % g.nW;
    const int bb = item / (g.heads * g.nW);
    for (int n = tid; n < LP; n += kB3Compute) {
      lse2[n] = (n < L) ? lse[(((size_t)bb * g.nW + ww) * g.heads + hd) * L + n] * kLog2e : INFINITY;
      Dv[n] = (n < L) ? Dpre[(size_t)tk[n] * g.heads + hd] : 0.f;
    }
  };
  // operand `role` (0 Q^, 1 K^, 2 V, 3 dO) of the (window, head) with token table `tk` -> buffer `buf`, by the compute threads
  auto gather = [&](int role, int buf, const int* tk, int hd) {
    unsigned char* dst = op_ptr(buf);
    for (int i = tid; i < L * kPieces; i += kB3Compute) {
      const int n = i / kPieces, c = i - n * kPieces;
      const __nv_bfloat16* src = (role < 3) ? qkv + (size_t)tk[n] * C3 + role * C + hd * D + c * 8
                                            : d_o + (size_t)tk[n] * C + hd * D + c * 8;
      cp_async16(dst + opnd_off(n, c), src);
    }
  };
  if (first < nitems) fill_tok(first, tokbuf0);
  __syncthreads();
  if (first < nitems) {
    fill_rows(first, tokbuf0);
    if (!item_is_box(first)) {
      const int hd = first % g.heads;
      gather(0, 0, tokbuf0, hd);
      gather(1, 1, tokbuf0, hd);
      gather(2, 2, tokbuf0, hd);
      gather(3, 3, tokbuf0, hd);
      cp_async_wait_all();
      fence_proxy_async_smem();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0 && first < nitems && !item_is_box(first))
    for (int i = 0; i < 4; ++i) mbar_arrive(&full[i]);

  const uint32_t idesc_s = umma_idesc_bf16(128, LP, false, false);   // [128 keys x LP queries] = A(k-major) B(k-major)^T
  const uint32_t idesc_o = umma_idesc_bf16(128, D, false, true);     // [128 x D] = A(k-major or TMEM) B(n-major)
  const uint32_t idesc_q = umma_idesc_bf16(128, D, true, true);      // [128 x D] = A(m-major) B(n-major)
  const uint32_t ds0 = smem_u32(sDS);

  // ---- control state (meaningful in thread 0 only) ---------------------------------------------------------------------
  const bool is_ctrl = (tid == 0);
  uint32_t cph_full = 0, cph_p = 0, cph_e = 0, cph_eq = 0;
  auto tma_operand = [&](int role, int buf, int item) {    // role: 0 Q^, 1 K^, 2 V, 3 dO
    const int hd = item % g.heads;
    const int ww_all = (item / g.heads) % g.nW;
    const int bb = item / (g.heads * g.nW);
    const int wh = ww_all / g.nWw, ww = ww_all - wh * g.nWw;
    mbar_arrive_expect_tx(&full[role], (uint32_t)kBoxes * 64u * (uint32_t)L);
    const CUtensorMap* tm = (role < 3) ? &tm_qkv : &tm_do;
    const int chunk0 = ((role < 3) ? role * C + hd * D : hd * D) / 32;
#pragma unroll
    for (int c = 0; c < kBoxes; ++c)
      tma_load_5d(op_ptr(buf) + c * kCS64, tm, &full[role], 0, chunk0 + c, ww * g.Ww + g.s1, wh * g.Wh + g.s0, bb);
  };
  if (is_ctrl && first < nitems && item_is_box(first)) {
    tma_operand(1, 1, first);
    tma_operand(0, 0, first);
    tma_operand(2, 2, first);
    tma_operand(3, 3, first);
  }
  __syncwarp();
  {
    // ================================================ compute warps ================================================
    const int grp = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;                    // row inside the current 128-row tile == TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    const int nchunks = LP / 16;
    const int c_begin = (nchunks * grp / 4) * 16, c_end = (nchunks * (grp + 1) / 4) * 16;   // this group's query columns
    const int ecol = grp * kEpiCols;                      // this group's output columns in the epilogues
    uint32_t ph_s = 0, ph_o = 0, ph_q = 0;
    int kb = 1, vb = 2;
    // per-phase cycle accounting of thread 0 (bring-up aid; the kProf = false instantiation carries none of it)
    long long ph_acc[kProf ? 8 : 1];
#pragma unroll
    for (int i = 0; i < (kProf ? 8 : 1); ++i) ph_acc[i] = 0;
    long long ph_t = kProf ? clock64() : 0;
#define SWB_ACC(i) do { if (kProf && tid == 0) { const long long now_ = clock64(); ph_acc[kProf ? (i) : 0] += now_ - ph_t; ph_t = now_; } } while (0)

    int it = 0;
    for (int item = first; item < nitems; item += gridDim.x, ++it) {
      int* tok = tokbuf0 + (it & 1) * kMaxLP;
      int* tok_next = tokbuf0 + ((it & 1) ^ 1) * kMaxLP;
      const int head = item % g.heads;
      const int w = (item / g.heads) % g.nW;
      const int item_next = item + gridDim.x;
      const bool has_next = item_next < nitems;
      const bool next_gather = has_next && !item_is_box(item_next);
      const int head_next = item_next % g.heads;
      if (has_next) fill_tok(item_next, tok_next);
      int label_split = LP;
      if (shifted) {
        const int wh = w / g.nWw;
        if (g.s0 > 0) {
          const int first_row = g.H - g.s0 - wh * g.Wh;
          label_split = first_row <= 0 ? 0 : (first_row >= g.Wh ? LP : first_row * g.Ww);
        } else {
          label_split = 0;
        }
      }
      const bool plain = (bias == nullptr) && !(label_split > 0 && label_split < L);
      const float scale = scale_p[head];
      const float scale_l2 = scale * kLog2e;
      named_bar_sync(1, kB3Compute);      // this item's lse2 / Dv / tok (written during the previous item) are visible
      const bool next_box = has_next && !next_gather;
      const uint32_t q0 = smem_u32(op_ptr(0)), k0 = smem_u32(op_ptr(kb)), v0 = smem_u32(op_ptr(vb)), g0 = smem_u32(op_ptr(3));
      auto issue_st = [&](int u) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16_ss(tmem_base + kColST, opnd_kmajor(k0, k, u * 128), opnd_kmajor(q0, k, 0), idesc_s, k > 0);
      };
      auto issue_dpt = [&](int u) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          umma_bf16_ss(tmem_base + kColDPT, opnd_kmajor(v0, k, u * 128), opnd_kmajor(g0, k, 0), idesc_s, k > 0);
      };
      if (is_ctrl) {
        mbar_wait(&full[1], cph_full, 800);
        mbar_wait(&full[0], cph_full, 801);
        if (it > 0) { mbar_wait(eqbar, cph_eq, 802); cph_eq ^= 1; }     // dQ^ columns [0,192) of the previous item drained
        tc_fence_after();
        issue_st(0);
        mbar_wait(&full[2], cph_full, 803);
        mbar_wait(&full[3], cph_full, 804);
        cph_full ^= 1;
        if (it > 0) { mbar_wait(ebar, cph_e, 805); cph_e ^= 1; }         // dV / dK^ columns of the previous item's last tile drained
        tc_fence_after();
        issue_dpt(0);
        umma_commit(sbar);
      }
      __syncwarp();
      SWB_ACC(0);
      float dsc_acc = 0.f;

      for (int u = 0; u < ntiles; ++u) {
        const bool last = (u == ntiles - 1);
        const int jk = u * 128 + r;                         // key slot of this thread
        const bool key_ok = jk < L;
        const bool warp_rows = u * 128 + quarter * 32 < LP;   // warp-uniform: this warp's lanes hold rows of the padded window
        mbar_wait(sbar, ph_s, 820 + u); ph_s ^= 1;
        tc_fence_after();
        SWB_ACC(1);
        if (last) {                       // the last MMA that reads V is complete: the next item's K^ goes where V was
          if (next_gather) gather(1, vb, tok_next, head_next);
          else if (is_ctrl && next_box) tma_operand(1, vb, item_next);
          __syncwarp();
        }
        // ---- P^T, dS^T of this thread's key row over the group's query columns ------------------------------------------
        float part = 0.f;                                   // sum_i dS_ij cos_ij
        if (warp_rows) {
          const int key_label = (jk >= label_split) ? 1 : 0;
          const bool row_exists = jk < LP;
          for (int c0 = c_begin; c0 < c_end; c0 += 16) {
            uint32_t sv[16], pv[16];
            tmem_ld_32x16(t_lane + kColST + c0, sv);
            tmem_ld_32x16(t_lane + kColDPT + c0, pv);
            tmem_ld_wait();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {          // two halves of 8 queries keep the live set small
              const int cb = c0 + hh * 8;
              float ls[8], dd[8];
              *reinterpret_cast<float4*>(&ls[0]) = *reinterpret_cast<const float4*>(&lse2[cb]);
              *reinterpret_cast<float4*>(&ls[4]) = *reinterpret_cast<const float4*>(&lse2[cb + 4]);
              *reinterpret_cast<float4*>(&dd[0]) = *reinterpret_cast<const float4*>(&Dv[cb]);
              *reinterpret_cast<float4*>(&dd[4]) = *reinterpret_cast<const float4*>(&Dv[cb + 4]);
              float pp[8], ds[8];
              if (plain) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float cosv = as_f(sv[hh * 8 + j]);
                  const float p = ex2_approx(fmaf(cosv, scale_l2, -ls[j]));     // pad queries: lse2 = +inf -> p = 0
                  pp[j] = key_ok ? p : 0.f;                                     // pad keys: cos = 0 but p != 0 -> force zero
                  ds[j] = key_ok ? p * (as_f(pv[hh * 8 + j]) - dd[j]) : 0.f;
                  part = fmaf(ds[j], cosv, part);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const int qi = cb + j;
                  const float cosv = as_f(sv[hh * 8 + j]);
                  float sl = cosv * scale_l2;
                  if (bias != nullptr && key_ok && qi < L) sl += __ldg(bias + ((size_t)head * L + qi) * L + jk) * kLog2e;
                  if (((qi >= label_split) ? 1 : 0) != key_label) sl += -100.0f * kLog2e;
                  const float p = key_ok ? ex2_approx(sl - ls[j]) : 0.f;
                  pp[j] = p;
                  ds[j] = (key_ok && qi < L) ? p * (as_f(pv[hh * 8 + j]) - dd[j]) : 0.f;
                  part = fmaf(ds[j], cosv, part);
                  if (dbias != nullptr && key_ok && qi < L) atomicAdd(dbias + ((size_t)head * L + qi) * L + jk, ds[j]);
                }
              }
              tmem_st_32x4(t_lane + kColP + cb / 2, pack8f(pp));               // 8 queries = 4 packed columns of P^T
              if (row_exists) *reinterpret_cast<uint4*>(sDS + (cb / 8) * SM::kDSCS + jk * 16) = pack8f(ds);
            }
          }
          tmem_st_wait();
        }
        dotk[((u & 1) * 4 + grp) * 128 + r] = part;
        if (key_ok) dsc_acc += part;        // rows beyond the window read garbage cosines (0 * NaN would poison the sum)
        fence_proxy_async_smem();       // dS^T (generic-proxy stores) -> visible to the tensor core's async-proxy reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(pbar);
        if (is_ctrl) {
          mbar_wait(pbar, cph_p, 808); cph_p ^= 1;
          tc_fence_after();
          for (int k = 0; k < LP / 16; ++k)       // dV_u = P^T_u dO   (A from tensor memory; dO read n-major: rows = queries = k)
            umma_bf16_ts(tmem_base + kColDV, tmem_base + kColP + k * 8, opnd_mnmajor(g0, k), idesc_o, k > 0);
          for (int k = 0; k < LP / 16; ++k)       // dK^_u = dS^T_u Q^
            umma_bf16_ss(tmem_base + kColDK, umma_desc_nosw(ds0 + u * 128 * 16 + 2 * k * SM::kDSCS, SM::kDSCS, 128),
                         opnd_mnmajor(q0, k), idesc_o, k > 0);
          umma_commit(obar);
          if (!last) {
            issue_st(u + 1);                      // runs while the warps drain dV_u / dK^_u
          } else {
            for (int t = 0; t < ntiles; ++t)      // dQ^_t = dS_t K^   (A = dS^T read m-major, K^ read n-major; k = keys)
              for (int k = 0; k < LP / 16; ++k)
                umma_bf16_ss(tmem_base + kColDQ + t * D, umma_desc_nosw(ds0 + t * 16 * SM::kDSCS + k * 256, 128, SM::kDSCS),
                             opnd_mnmajor(k0, k), idesc_q, k > 0);
            umma_commit(qbar);
          }
        }
        __syncwarp();
        SWB_ACC(2);

        // ---- dV_u / dK^_u epilogue: this thread owns 24 of the 96 columns of its key row ----------------------------------
        mbar_wait(obar, ph_o, 830 + u); ph_o ^= 1;
        tc_fence_after();
        SWB_ACC(3);
        if (last && has_next) {
          fill_rows(item_next, tok_next);                   // every warp is past its last read of lse2 / Dv
          if (next_gather) { gather(0, 0, tok_next, head_next); gather(3, 3, tok_next, head_next); }
          else if (is_ctrl && next_box) { tma_operand(0, 0, item_next); tma_operand(3, 3, item_next); }   // Q^ / dO served their last MMA
          __syncwarp();
        }
        if (u * 128 + quarter * 32 < L) {                   // warp-uniform: the warp has real key rows
          // k^ pieces are requested first (shared memory, or L2 once the K^ buffer is being recycled), then dV is drained and
          // stored while they are in flight, then dK^
          uint4 kraw[kEpiCols / 8];
          float ink = 0.f;
          int tk = 0;
          if (key_ok) {
            tk = tok[jk];
            ink = __ldg(inv_norm + (size_t)tk * 2 * g.heads + g.heads + head);
#pragma unroll
            for (int i = 0; i < kEpiCols / 8; ++i)
              kraw[i] = !last ? *reinterpret_cast<const uint4*>(op_ptr(kb) + opnd_off(jk, grp * (kEpiCols / 8) + i))
                              : __ldg(reinterpret_cast<const uint4*>(qkv + (size_t)tk * C3 + C + head * D + ecol + i * 8));
          }
          __nv_bfloat16* dst = dqkv + (size_t)tk * C3 + head * D + ecol;
          {
            uint32_t av[kEpiCols];
#pragma unroll
            for (int i = 0; i < kEpiCols / 8; ++i)
              tmem_ld_32x8(t_lane + kColDV + ecol + i * 8, *reinterpret_cast<uint32_t(*)[8]>(&av[i * 8]));
            tmem_ld_wait();
            if (key_ok) {
#pragma unroll
              for (int i = 0; i < kEpiCols / 8; ++i)
                *reinterpret_cast<uint4*>(dst + 2 * C + i * 8) = pack8f(reinterpret_cast<const float*>(&av[i * 8]));
            }
          }
          {
            uint32_t ak[kEpiCols];
#pragma unroll
            for (int i = 0; i < kEpiCols / 8; ++i)
              tmem_ld_32x8(t_lane + kColDK + ecol + i * 8, *reinterpret_cast<uint32_t(*)[8]>(&ak[i * 8]));
            tmem_ld_wait();
            if (key_ok) {
              const float* dk4 = dotk + (u & 1) * 4 * 128 + r;
              const float dot = (dk4[0] + dk4[128]) + (dk4[256] + dk4[384]);     // sum_i dS_ij cos_ij = <k^_j, dK^_j>
              const float ks = ink * scale;
#pragma unroll
              for (int i = 0; i < kEpiCols / 8; ++i) {
                float kh[8], ok[8];
                const uint32_t wds[4] = {kraw[i].x, kraw[i].y, kraw[i].z, kraw[i].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  kh[2 * e] = __uint_as_float(wds[e] << 16);
                  kh[2 * e + 1] = __uint_as_float(wds[e] & 0xffff0000u);
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) ok[e] = ks * fmaf(-kh[e], dot, as_f(ak[i * 8 + e]));   // dk = inv_norm scale (dK^ - k^ <k^, dK^>)
                *reinterpret_cast<uint4*>(dst + C + i * 8) = pack8f(ok);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(ebar);
        if (is_ctrl && !last) {
          mbar_wait(ebar, cph_e, 806); cph_e ^= 1;      // dV_u / dK^_u drained by every warp: dP^T may overwrite them
          tc_fence_after();
          issue_dpt(u + 1);
          umma_commit(sbar);
        }
        __syncwarp();
        SWB_ACC(4);
      }

      // ---- dQ^ epilogue ------------------------------------------------------------------------------------------------------
      mbar_wait(qbar, ph_q, 840); ph_q ^= 1;
      tc_fence_after();
      SWB_ACC(5);
      if (next_gather) gather(2, kb, tok_next, head_next);        // K^ served its last MMA: the next item's V goes there
      else if (is_ctrl && next_box) tma_operand(2, kb, item_next);
      __syncwarp();
      for (int t = 0; t < ntiles; ++t) {
        const int n = t * 128 + r;
        const bool row_ok = n < L;
        const bool warp_has = t * 128 + quarter * 32 < L;
        uint32_t aq[kEpiCols];
        uint4 qraw[kEpiCols / 8];
        float partq = 0.f, inq = 0.f;
        int tk = 0;
        auto q_at = [&](int i, int e) {     // element e of the i-th 8-column piece of q^
          const uint32_t wd = (e >> 1) == 0 ? qraw[i].x : (e >> 1) == 1 ? qraw[i].y : (e >> 1) == 2 ? qraw[i].z : qraw[i].w;
          return __uint_as_float((e & 1) ? (wd & 0xffff0000u) : (wd << 16));
        };
        if (warp_has) {
          if (row_ok) {
            tk = tok[n];
            inq = __ldg(inv_norm + (size_t)tk * 2 * g.heads + head);
#pragma unroll
            for (int i = 0; i < kEpiCols / 8; ++i)
              qraw[i] = __ldg(reinterpret_cast<const uint4*>(qkv + (size_t)tk * C3 + head * D + ecol + i * 8));
          }
#pragma unroll
          for (int i = 0; i < kEpiCols / 8; ++i)
            tmem_ld_32x8(t_lane + kColDQ + t * D + ecol + i * 8, *reinterpret_cast<uint32_t(*)[8]>(&aq[i * 8]));
          tmem_ld_wait();
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < kEpiCols / 8; ++i)
#pragma unroll
              for (int e = 0; e < 8; ++e) partq = fmaf(q_at(i, e), as_f(aq[i * 8 + e]), partq);
          }
        }
        red[((t & 1) * 4 + grp) * 128 + r] = partq;
        named_bar_sync(2, kB3Compute);
        if (warp_has && row_ok) {
          const float* rq = red + (t & 1) * 4 * 128 + r;
          const float dot = (rq[0] + rq[128]) + (rq[256] + rq[384]);          // <q^_i, dQ^_i>
          const float qs = inq * scale;
          __nv_bfloat16* dst = dqkv + (size_t)tk * C3 + head * D + ecol;
#pragma unroll
          for (int i = 0; i < kEpiCols / 8; ++i) {
            float oq[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) oq[e] = qs * fmaf(-q_at(i, e), dot, as_f(aq[i * 8 + e]));   // dq = inv_norm scale (dQ^ - q^ <q^, dQ^>)
            *reinterpret_cast<uint4*>(dst + i * 8) = pack8f(oq);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(eqbar);
      SWB_ACC(6);
      if (next_gather) {                                     // the gathered operands of the next item are complete
        cp_async_wait_all();
        fence_proxy_async_smem();
        named_bar_sync(3, kB3Compute);
        if (tid == 0)
          for (int i = 0; i < 4; ++i) mbar_arrive(&full[i]);
      }
      dsc_acc = warp_sum(dsc_acc);
      if (lane == 0) atomicAdd(&dsc_heads[head], dsc_acc);
      const int tmp = kb; kb = vb; vb = tmp;
      SWB_ACC(7);
    }
    named_bar_sync(1, kB3Compute);
    if (tid < g.heads) {
      const float v = dsc_heads[tid];
      if (v != 0.f) atomicAdd(dscale + tid, v);
    }
    if (kProf && g_phase_buf3 != nullptr && tid == 0 && blockIdx.x < 4096)
      for (int i = 0; i < 8; ++i) g_phase_buf3[blockIdx.x * 16 + i] = ph_acc[i];
#undef SWB_ACC
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

__global__ void __launch_bounds__(256) attn_rowdot3_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o,
                                                           float* __restrict__ out, long long n_pairs, int d) {
  // D[t, h] = <dO[t, h, :], O[t, h, :]>, four lanes per (token, head)
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long pair = gid >> 2;
  const int q = (int)(gid & 3);
  float acc = 0.f;
  if (pair < n_pairs) {
    const __nv_bfloat16* a = o + pair * d;
    const __nv_bfloat16* b = d_o + pair * d;
    for (int c = q * 8; c < d; c += 32) {
      float a8[8], b8[8];
      ld8(a + c, a8);
      ld8(b + c, b8);
#pragma unroll
      for (int e = 0; e < 8; ++e) acc = fmaf(a8[e], b8[e], acc);
    }
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (q == 0 && pair < n_pairs) out[pair] = acc;
}

static bool g_prof3 = false;
int attn_set_phase_buffer3(long long* buf) {
  SWB_CUDA(cudaMemcpyToSymbol(g_phase_buf3, &buf, sizeof(buf)));
  g_prof3 = buf != nullptr;
  return SWINB200_OK;
}

int attn_tcgen05_bwd3(const void* qkv, const float* inv_norm, const float* scale, const float* bias, const void* o, const void* d_o,
                      const float* lse, void* dqkv, float* dscale, float* dbias, float* ws, const AttnGeom& g, cudaStream_t stream) {
  using SM = Bwd3Smem<96>;
  static bool configured = false;
  if (!configured) {
    SWB_CUDA(cudaFuncSetAttribute(attn_tc_bwd3_kernel<96, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
    SWB_CUDA(cudaFuncSetAttribute(attn_tc_bwd3_kernel<96, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
    configured = true;
  }
  CUtensorMap tm_qkv, tm_do;
  if (int e = attn_make_window_tmap(&tm_qkv, qkv, g.B, g.H, g.W, 3 * g.C, g.Wh, g.Ww)) return e;
  if (int e = attn_make_window_tmap(&tm_do, d_o, g.B, g.H, g.W, g.C, g.Wh, g.Ww)) return e;
  const long long n_pairs = (long long)g.B * g.H * g.W * g.heads;
  attn_rowdot3_kernel<<<(unsigned)((n_pairs * 4 + 255) / 256), 256, 0, stream>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)d_o, ws,
                                                                                 n_pairs, g.C / g.heads);
  SWB_LAUNCH_CHECK();
  const int grid = min(g.B * g.nW * g.heads, sm_count());
  static int colp_env = -1;
  if (colp_env < 0) { const char* e = getenv("SWINB200_BWD3_COLP"); colp_env = e ? atoi(e) : 416; }
  const uint32_t colp = (uint32_t)colp_env;
  if (g_prof3)
    attn_tc_bwd3_kernel<96, true><<<grid, kB3Threads, SM::kBytes, stream>>>(tm_qkv, tm_do, ws, (const __nv_bfloat16*)qkv, inv_norm, scale,
                                                                            bias, (const __nv_bfloat16*)d_o, lse, (__nv_bfloat16*)dqkv,
                                                                            dscale, dbias, g, colp);
  else
    attn_tc_bwd3_kernel<96, false><<<grid, kB3Threads, SM::kBytes, stream>>>(tm_qkv, tm_do, ws, (const __nv_bfloat16*)qkv, inv_norm, scale,
                                                                             bias, (const __nv_bfloat16*)d_o, lse, (__nv_bfloat16*)dqkv,
                                                                             dscale, dbias, g, colp);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

}  // namespace swinb200
