// tcgen05 windowed cosine attention, forward, third generation: persistent CTAs, operands by TMA window boxes, double-buffered.
//
// Reference math: swinv2_global.py:300-318 -- S = scale * q^ k^T (+ CPB bias) (+ shift mask), P = softmax(S), O = P v; the
// roll / window_partition / window_reverse copies (:89-119, 446-478) are folded into the TMA box coordinates and the scatter.
//
// Work item = (sample, window, head); one CTA per SM loops over its items.  9 warps:
//   warps 0..7   compute: warp w owns the 32 query rows of TMEM lane quarter (w & 3) of query tile (w >> 2); thread = row
//   warp 8       control: runs convergently, one elected lane issues every TMA load and tcgen05.mma
// Per item (query-major):  S_t = Q^_t K^T (both 128-query tiles into tensor memory) -> two passes over S in tensor memory
// (row max, then p = 2^(s - max) written back IN PLACE as packed bf16) -> O_t = P_t V with A = P_t from tensor memory (TS) ->
// O / rowsum -> bf16 rows parked in shared memory -> whole-row 192-byte stores to the un-rolled token order; LSE saved.
// TMEM columns per tile t (base 184 t):  S [0,176)  P bf16 in place [0,88)  O [88,184)   (368 of 512 columns).
// Shared memory: two stages of (Q^, K^, V), each operand 3 x [176 rows x 64 B] 64B-swizzled boxes.  The loads of item i+1 are
// issued when item i-1 has released its stage, i.e. they are in flight for the whole of item i.  The output rows are parked in
// the Q^ buffer of the item's own stage (Q^ is dead once S is complete); K^ / V pad rows [L, 176) are zeroed once and never
// written again, so padded keys contribute exact zeros.
#include <stdlib.h>
#include "attn_tc.cuh"

namespace swinb200 {

constexpr int kF3Compute = 256;                  // 8 compute warps
constexpr int kF3Threads = kF3Compute + 32;      // + the control warp
constexpr int kF3CtrlWarp = kF3Compute / 32;
constexpr uint32_t kF3TileCols = 184;            // TMEM columns per query tile: S 176 (P in place) + 8 more for O [88,184)

template <int D>
struct Fwd3Smem {
  static constexpr int kTile = (D / 32) * kCS64;            // one operand: 3 x [176 rows x 64 B]
  static constexpr int kStage = 3 * kTile;                  // Q^, K^, V
  static constexpr int kOffTok = 2 * kStage;                // [2][176] token indices
  static constexpr int kOffBar = kOffTok + 2 * kMaxLP * 4;
  static constexpr int kBytes = kOffBar + 128;
  static_assert(kTile % 512 == 0, "64B-swizzled operand tiles need 512-byte alignment");
  static_assert(kMaxLP * kRowPitch <= kTile || 162 * kRowPitch <= kTile, "output staging must fit in the Q^ buffer");
  static_assert(kBytes <= 227 * 1024, "shared memory budget");
};

__device__ __forceinline__ bool elect_one_f3() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void named_bar_f3(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ long long* g_phase_buf_f3 = nullptr;     // bring-up: per-CTA cycles per phase (tools/attn_phases.py)

template <int D>
__global__ void __launch_bounds__(kF3Threads, 1)
attn_tc_fwd3_kernel(const __grid_constant__ CUtensorMap tm_qkv, const __nv_bfloat16* __restrict__ qkv,
                    const float* __restrict__ scale_p, const float* __restrict__ bias, __nv_bfloat16* __restrict__ o,
                    float* __restrict__ lse, const AttnGeom g) {
  using SM = Fwd3Smem<D>;
  constexpr float kLog2e = 1.4426950408889634f;
  constexpr int kPieces = D / 8;
  constexpr int kBoxes = D / 32;
  extern __shared__ __align__(1024) unsigned char smem[];
  int* tokbuf0 = reinterpret_cast<int*>(smem + SM::kOffTok);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::kOffBar);
  uint64_t* full = bars;            // [2] the stage's three operands have landed
  uint64_t* sbar = bars + 2;        // S of both query tiles is in tensor memory
  uint64_t* pbar = bars + 3;        // P written back by every compute warp                     (8 warps)
  uint64_t* obar = bars + 4;        // O of both tiles is complete
  uint64_t* ebar = bars + 5;        // output rows stored, stage released                       (8 warps)
  uint64_t* tbar = bars + 6;        // O is in registers: tensor memory may take the next item's S (8 warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int L = g.L, LP = g.LP, C = g.C, C3 = 3 * g.C;
  const int ntiles = (L > 128) ? 2 : 1;
  const int nitems = g.B * g.nW * g.heads;
  const int first = blockIdx.x, stride = gridDim.x;
  const bool is_compute = warp < kF3CtrlWarp;

  auto op_ptr = [&](int stage, int op) { return smem + stage * SM::kStage + op * SM::kTile; };   // op: 0 Q^, 1 K^, 2 V
  auto item_is_box = [&](int item) {
    const int ww_all = (item / g.heads) % g.nW;
    const int wh = ww_all / g.nWw, ww = ww_all - wh * g.nWw;
    return !((g.s0 > 0 && (wh + 1) * g.Wh + g.s0 > g.H) || (g.s1 > 0 && (ww + 1) * g.Ww + g.s1 > g.W));
  };
  auto fill_tok = [&](int item, int* tk) {          // compute threads
    const int ww = (item / g.heads) % g.nW;
    const int bb = item / (g.heads * g.nW);
    for (int n = tid; n < LP; n += kF3Compute) {
      int rr;
      tk[n] = (n < L) ? win_token(g, bb, ww, n, rr) : -1;
    }
  };
  // all three operands of `item` -> `stage`, per-thread 16-byte copies (windows that wrap around the cyclic shift)
  auto gather_item = [&](int item, int stage, const int* tk) {
    const int hd = item % g.heads;
    for (int i = tid; i < 3 * L * kPieces; i += kF3Compute) {
      const int op = i / (L * kPieces);
      const int rem = i - op * L * kPieces;
      const int n = rem / kPieces, c = rem - n * kPieces;
      cp_async16(op_ptr(stage, op) + opnd_off(n, c), qkv + (size_t)tk[n] * C3 + op * C + hd * D + c * 8);
    }
  };

  // ---- one-time set-up -----------------------------------------------------------------------------------------------
  if (tid == 0) {
    prefetch_tmap(&tm_qkv);
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_init(sbar, 1);
    mbar_init(pbar, kF3Compute / 32);
    mbar_init(obar, 1);
    mbar_init(ebar, kF3Compute / 32);
    mbar_init(tbar, kF3Compute / 32);
    fence_barrier_init();
  }
  if (warp == kF3CtrlWarp) tmem_alloc(tmem_slot, 512);
  for (int i = tid; i < (LP - L) * kPieces * 6; i += kF3Threads) {      // zero pad rows [L, LP) of every operand buffer
    const int buf = i / ((LP - L) * kPieces);
    const int rem = i - buf * (LP - L) * kPieces;
    const int c = rem / (LP - L), r = L + rem % (LP - L);
    *reinterpret_cast<uint4*>(op_ptr(buf / 3, buf % 3) + opnd_off(r, c)) = make_uint4(0, 0, 0, 0);
  }
  if (is_compute) {
    if (first < nitems) fill_tok(first, tokbuf0);
    if (first + stride < nitems) fill_tok(first + stride, tokbuf0 + kMaxLP);
  }
  __syncthreads();
  if (is_compute) {
    bool any = false;
    if (first < nitems && !item_is_box(first)) { gather_item(first, 0, tokbuf0); any = true; }
    if (first + stride < nitems && !item_is_box(first + stride)) { gather_item(first + stride, 1, tokbuf0 + kMaxLP); any = true; }
    if (any) {
      cp_async_wait_all();
      fence_proxy_async_smem();
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (tid == 0) {
    if (first < nitems && !item_is_box(first)) mbar_arrive(&full[0]);
    if (first + stride < nitems && !item_is_box(first + stride)) mbar_arrive(&full[1]);
  }
  const uint32_t idesc_s = umma_idesc_bf16(128, LP, false, false);   // [128 queries x LP keys] = A(k-major) B(k-major)^T
  const uint32_t idesc_o = umma_idesc_bf16(128, D, false, true);     // [128 x D] = A(TMEM) B(n-major)

  if (!is_compute) {
    // =========================================== control: TMA + MMA issue ===========================================
    auto tma_item = [&](int item, int stage) {        // whole warp; one lane issues
      const int hd = item % g.heads;
      const int ww_all = (item / g.heads) % g.nW;
      const int bb = item / (g.heads * g.nW);
      const int wh = ww_all / g.nWw, ww = ww_all - wh * g.nWw;
      if (elect_one_f3()) {
        mbar_arrive_expect_tx(&full[stage], 3u * kBoxes * 64u * (uint32_t)L);
#pragma unroll
        for (int op = 0; op < 3; ++op)
#pragma unroll
          for (int c = 0; c < kBoxes; ++c)
            tma_load_5d(op_ptr(stage, op) + c * kCS64, &tm_qkv, &full[stage], 0, (op * C + hd * D) / 32 + c, ww * g.Ww + g.s1,
                        wh * g.Wh + g.s0, bb);
      }
      __syncwarp();
    };
    if (first < nitems && item_is_box(first)) tma_item(first, 0);
    if (first + stride < nitems && item_is_box(first + stride)) tma_item(first + stride, 1);
    const int nk = LP / 16;
    uint32_t ph_p = 0, ph_e = 0, ph_t = 0;
    int it = 0;
    for (int item = first; item < nitems; item += stride, ++it) {
      const int s = it & 1;
      // previous item: its O sits in registers -> S of this item may overwrite tensor memory while those rows are still
      // being parked and stored
      if (it > 0) { mbar_wait(tbar, ph_t, 903); ph_t ^= 1; }
      mbar_wait(&full[s], (uint32_t)((it >> 1) & 1), 901);
      tc_fence_after();
      const uint32_t q0 = smem_u32(op_ptr(s, 0)), k0 = smem_u32(op_ptr(s, 1)), v0 = smem_u32(op_ptr(s, 2));
      if (elect_one_f3()) {
        for (int t = 0; t < ntiles; ++t)
#pragma unroll
          for (int k = 0; k < D / 16; ++k)
            umma_bf16_ss(tmem_base + t * kF3TileCols, opnd_kmajor(q0, k, t * 128), opnd_kmajor(k0, k, 0), idesc_s, k > 0);
        umma_commit(sbar);
      }
      __syncwarp();
      if (it > 0) {
        mbar_wait(ebar, ph_e, 900); ph_e ^= 1;           // previous item: output rows stored, its stage released
        const int item_next = item + stride;
        if (item_next < nitems && item_is_box(item_next)) tma_item(item_next, s ^ 1);
      }
      mbar_wait(pbar, ph_p, 902); ph_p ^= 1;
      tc_fence_after();
      if (elect_one_f3()) {
        const uint64_t bv = opnd_mnmajor(v0, 0);
        for (int t = 0; t < ntiles; ++t) {
#pragma unroll
          for (int k = 0; k < kMaxLP / 16; ++k)     // O_t = P_t V   (A = P from tensor memory, V read n-major: rows = keys = k)
            if (k < nk)
              umma_bf16_ts(tmem_base + t * kF3TileCols + 88, tmem_base + t * kF3TileCols + k * 8, bv + (uint64_t)(k * 64), idesc_o, k > 0);
        }
        umma_commit(obar);
      }
      __syncwarp();
    }
  } else {
    // ================================================ compute warps ================================================
    const int t = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int n = t * 128 + r;                       // this thread's query slot
    const bool row_ok = n < L;
    const bool warp_rows = (t < ntiles) && (t * 128 + quarter * 32 < L);     // warp-uniform: the warp has real query rows
    const uint32_t t_s = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)t * kF3TileCols;
    uint32_t ph_s = 0, ph_o = 0;
    int it = 0;
    long long* const prof = g_phase_buf_f3;
    long long ph_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long ph_t = clock64();
#define F3_ACC(i) do { if (prof != nullptr && tid == 0) { const long long now_ = clock64(); ph_acc[i] += now_ - ph_t; ph_t = now_; } } while (0)
    for (int item = first; item < nitems; item += stride, ++it) {
      const int s = it & 1;
      int* tok = tokbuf0 + s * kMaxLP;
      const int head = item % g.heads;
      const int w = (item / g.heads) % g.nW;
      const int b = item / (g.heads * g.nW);
      int label_split = LP;
      if ((g.s0 > 0) || (g.s1 > 0)) {
        const int wh = w / g.nWw;
        if (g.s0 > 0) {
          const int first_row = g.H - g.s0 - wh * g.Wh;
          label_split = first_row <= 0 ? 0 : (first_row >= g.Wh ? LP : first_row * g.Ww);
        } else {
          label_split = 0;
        }
      }
      const bool plain = (bias == nullptr) && !(label_split > 0 && label_split < L);
      const float scale_l2 = scale_p[head] * kLog2e;
      F3_ACC(6);
      mbar_wait(sbar, ph_s, 910); ph_s ^= 1;
      F3_ACC(0);
      tc_fence_after();
      float row_sum = 0.f, row_max = -INFINITY, cos_sum = 0.f;      // cos_sum = sum_j p_j cos_j
      if (warp_rows) {
        if (plain) {
          // Row maximum.  Cosines are bounded by 1, so for scale * log2(e) < 60 the fixed bound scale * 1 is a safe softmax
          // offset: the smallest possible p = 2^(-2 scale log2 e) > 2^-120 is still a normal fp32 / bf16 number, the row sum
          // is at least the p of the row's best key, and the common factor cancels in O = P V / rowsum and in the LSE.  That
          // saves the first of the two passes over S in tensor memory (the kernel's bound).  Larger scales (the clamp allows
          // up to 100) take the exact maximum: pad keys hold exact zeros, including them can only raise it.
          float mx = 1.0f;
          if (scale_l2 >= 60.0f) {
          uint32_t va[16], vb[16];
          mx = -INFINITY;
          tmem_ld_32x16(t_s, va);
          for (int c0 = 0; c0 < LP; c0 += 32) {
            tmem_ld_wait();
            const bool has_b = c0 + 16 < LP;
            if (has_b) tmem_ld_32x16(t_s + c0 + 16, vb);
#pragma unroll
            for (int j = 0; j < 16; ++j) mx = fmaxf(mx, as_f(va[j]));
            if (has_b) {
              tmem_ld_wait();
              if (c0 + 32 < LP) tmem_ld_32x16(t_s + c0 + 32, va);
#pragma unroll
              for (int j = 0; j < 16; ++j) mx = fmaxf(mx, as_f(vb[j]));
            }
          }
          }
          row_max = mx * scale_l2;
          const float neg_m = -row_max;
          // pass 2: p = 2^(scale*cos - max) -> packed bf16 over the columns already consumed; fp32 row sum.  The next 16
          // columns are requested before the current ones are processed (the stores go to columns [8k, 8k+8), the load in
          // flight reads [16k+16, 16k+32): never the same)
          auto softmax16 = [&](const uint32_t (&v)[16], int c0) {
            float p[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) p[j] = ex2_approx(fmaf(as_f(v[j]), scale_l2, neg_m));
            if (c0 + 16 > L) {
#pragma unroll
              for (int j = 0; j < 16; ++j) if (c0 + j >= L) p[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              row_sum += p[j];
              cos_sum = fmaf(p[j], as_f(v[j]), cos_sum);
            }
            tmem_st_32x8(t_s + c0 / 2, pack8(p, 0), pack8(p, 8));
          };
          uint32_t va[16], vb[16];
          tmem_ld_32x16(t_s, va);
          for (int c0 = 0; c0 < LP; c0 += 32) {
            tmem_ld_wait();
            const bool has_b = c0 + 16 < LP;
            if (has_b) tmem_ld_32x16(t_s + c0 + 16, vb);
            softmax16(va, c0);
            if (has_b) {
              tmem_ld_wait();
              if (c0 + 32 < LP) tmem_ld_32x16(t_s + c0 + 32, va);
              softmax16(vb, c0 + 16);
            }
          }
        } else {
          // continuous position bias and / or the shifted-window mask (-100 across region labels)
          const float* brow = (bias != nullptr && row_ok) ? bias + ((size_t)head * L + n) * L : nullptr;
          const int my_label = (n >= label_split) ? 1 : 0;
          for (int c0 = 0; c0 < LP; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x16(t_s + c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int key = c0 + j;
              float sv = as_f(v[j]) * scale_l2;
              if (brow != nullptr && key < L) sv += brow[key] * kLog2e;
              if (((key >= label_split) ? 1 : 0) != my_label) sv += -100.0f * kLog2e;
              if (key < L) row_max = fmaxf(row_max, sv);
            }
          }
          for (int c0 = 0; c0 < LP; c0 += 16) {
            uint32_t v[16];
            tmem_ld_32x16(t_s + c0, v);
            tmem_ld_wait();
            float p[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int key = c0 + j;
              float sv = as_f(v[j]) * scale_l2;
              if (brow != nullptr && key < L) sv += brow[key] * kLog2e;
              if (((key >= label_split) ? 1 : 0) != my_label) sv += -100.0f * kLog2e;
              p[j] = (key < L && row_ok) ? ex2_approx(sv - row_max) : 0.f;
              row_sum += p[j];
              if (key < L && row_ok) cos_sum = fmaf(p[j], as_f(v[j]), cos_sum);
            }
            tmem_st_32x8(t_s + c0 / 2, pack8(p, 0), pack8(p, 8));
          }
        }
        if (row_ok) {
          const size_t ri = (((size_t)b * g.nW + w) * g.heads + head) * L + n;
          lse[ri] = (row_max + log2f(row_sum)) * 0.6931471805599453f;
          // second plane: E_P[cos] of the row.  The backward accumulates d(scale) = sum dS (cos - E_P[cos]); the row sum of dS
          // is zero in exact arithmetic, so this changes nothing but removes the first-order sensitivity to D = <dO, O>
          // being formed from the bf16-rounded O.
          lse[(size_t)g.B * g.nW * g.heads * L + ri] = cos_sum / row_sum;
        }
        tmem_st_wait();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pbar);
      F3_ACC(1);

      // ---- O / rowsum -> bf16 rows parked in the (dead) Q^ buffer -> whole-row stores ------------------------------------
      mbar_wait(obar, ph_o, 911); ph_o ^= 1;
      F3_ACC(2);
      tc_fence_after();
      unsigned char* stage_rows = op_ptr(s, 0);
      if (warp_rows) {
        const float inv = 1.0f / row_sum;
        float ov[D];
#pragma unroll
        for (int c0 = 0; c0 < D; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(t_s + 88 + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) ov[c0 + j] = as_f(v[j]) * inv;
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tbar);
        if (row_ok) park_row<D>(stage_rows, n, ov);
      } else {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tbar);
      }
      F3_ACC(3);
      named_bar_f3(1, kF3Compute);
      F3_ACC(4);
      // per-thread 16-byte pieces of whole 192-byte rows.  (Three TMA box stores per item -- the inverse of the operand loads --
      // were measured slower: the stage cannot be released before the copy engine has read the parked tile, 2.9 k cycles
      // per item against 1.8 k for this loop.)
      scatter_rows<D>(stage_rows, L, tok, 0, o, C, head * D, tid, kF3Compute);
      // the stage is released by the barrier below; its next user is the item after next: token table and, for a window
      // that wraps around the shift, the operand gather (TMA boxes are issued by the control warp after the same barrier)
      const int item_nn = item + 2 * stride;
      const bool nn_gather = item_nn < nitems && !item_is_box(item_nn);
      named_bar_f3(2, kF3Compute);                    // every thread has finished reading the parked rows / this tok table
      F3_ACC(5);
      if (item_nn < nitems) fill_tok(item_nn, tok);
      if (nn_gather) {
        named_bar_f3(3, kF3Compute);
        gather_item(item_nn, s, tok);
        cp_async_wait_all();
      }
      fence_proxy_async_smem();                       // parked rows / gathered operands (generic proxy) before async-proxy users
      __syncwarp();
      if (nn_gather) {
        named_bar_f3(3, kF3Compute);
        if (tid == 0) mbar_arrive(&full[s]);
      }
      if (lane == 0) mbar_arrive(ebar);
    }
    if (prof != nullptr && tid == 0 && blockIdx.x < 4096)
      for (int i = 0; i < 8; ++i) prof[blockIdx.x * 16 + i] = ph_acc[i];
#undef F3_ACC
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kF3CtrlWarp) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int attn_set_phase_buffer_f3(long long* buf) {
  SWB_CUDA(cudaMemcpyToSymbol(g_phase_buf_f3, &buf, sizeof(buf)));
  return SWINB200_OK;
}

int attn_tcgen05_fwd3(const void* qkv, const float* scale, const float* bias, void* o, float* lse, const AttnGeom& g, cudaStream_t stream) {
  using SM = Fwd3Smem<96>;
  static bool configured = false;
  if (!configured) {
    SWB_CUDA(cudaFuncSetAttribute(attn_tc_fwd3_kernel<96>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kBytes));
    configured = true;
  }
  CUtensorMap tm_qkv;
  if (int e = attn_make_window_tmap(&tm_qkv, qkv, g.B, g.H, g.W, 3 * g.C, g.Wh, g.Ww)) return e;
  const int grid = min(g.B * g.nW * g.heads, sm_count());
  attn_tc_fwd3_kernel<96><<<grid, kF3Threads, SM::kBytes, stream>>>(tm_qkv, (const __nv_bfloat16*)qkv, scale, bias, (__nv_bfloat16*)o, lse, g);
  SWB_LAUNCH_CHECK();
  return SWINB200_OK;
}

}  // namespace swinb200
