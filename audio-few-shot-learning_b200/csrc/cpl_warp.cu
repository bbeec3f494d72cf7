// CPL loss, one WARP per episode (same closed form and reference as cpl.cu: CPL_Loss.forward /
// similarity_sampling, loops/loss.py:118-165), for the shapes whose W normalised prototypes fit a 5 KB
// shared-memory slice per warp (5-way at Dp <= 256).
//
// Lane l owns Dp/32 columns of every row.  Query rows stream through registers kB at a time; the kB*(W+1)
// per-lane partials of a batch (W dot products with the normalised prototypes + the row's squared norm) are
// combined by one transposing butterfly, which leaves lane u*(W+1)+w with <p^_w, q_u>, so the division by the
// norm and the temperature happens once per matrix entry.  The W x Nq similarity matrix lives in the warp's
// shared-memory slice; the masked log-sum-exp of row i runs on lane i.  The backward builds dL/dC column by
// column (lane j), then walks the query rows a second time (L2 hits) to store dQ and accumulate dP in registers -
// or ONCE, when the forward saved C and 1/|q| for it (afsl_cpl_fwd_save_f32 / afsl_cpl_bwd_saved_f32, what the autograd
// op launches: 0.58 -> 0.71 of the HBM roofline at Dp = 256; 128 registers for a fourth CTA per SM spill 628 bytes).
// Without a sampled-negative mask (M >= per-class count, the reference's default M = 5 at 5 queries per class)
// every row of class w sees the same negatives, so their exp-sum S_w is computed once per class and the
// row statistics / dL/dC follow from it with O(W) work per lane instead of O(Nq) (same value, regrouped sums).
// No CTA-wide barrier anywhere: warps are independent, 4 per CTA, one episode each.
#include "cpl.cuh"
#include "warp_rows.cuh"

namespace afsl {
namespace {

using namespace warp_rows;

constexpr float kCosEps = 1e-8f;

__host__ __device__ inline int round4(int n) { return (n + 3) & ~3; }
// per-warp slice: p^ [W*D] | C [W*Nq] | G/T pairs [W*Nqp*2] | G/T/|q| pairs [W*Nqp*2] | rmax, rsum [Nq] | 1/|q| [Nqp] | labels [Nq]
__host__ __device__ inline int slice_words(int kWD, int W, int Nq, int Nqp) {
  return kWD + round4(W * Nq) + 2 * round4(2 * W * Nqp) + 2 * round4(Nq) + round4(Nqp) + round4(Nq);
}
__device__ __forceinline__ f32x2 lds_pair(const float* p) {
  f32x2 r;
  asm volatile("ld.shared.b64 %0, [%1];" : "=l"(r) : "r"(smem_u32(p)));
  return r;
}



template <int kW, int kV, int kB, bool kBwd, bool kSaved>
__global__ void __launch_bounds__(kCtaThreads, kBwd ? (kV == 8 ? 3 : kV == 4 ? 4 : 5) : (kV == 8 ? 4 : kV == 4 ? 6 : 8)) cpl_warp_kernel(const CplParams p) {
  extern __shared__ __align__(16) float smem_raw[];
  constexpr int kH = kV / 2, kD = kV * 32, kC = kW + 1, kVals = kB * kC, kN = pow2_ceil(kVals);
  static_assert(kVals <= 32, "a batch of rows must fit one value per lane");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int e = blockIdx.x * kWarpsPerCta + warp;
  if (e >= p.E) return;
  const int Nq = p.Nq, words = (Nq + 31) >> 5;
  const int nq_pad = (Nq + kB - 1) / kB * kB;
  float* sp = smem_raw + (size_t)warp * slice_words(kW * kD, kW, Nq, nq_pad);   // [kW*kD] normalised prototypes
  float* sim = sp + kW * kD;                                            // [kW*Nq] C
  float* gw2 = sim + round4(kW * Nq);                                   // [kW*nq_pad] pairs (G/T, G/T), zero past Nq
  float* gj2 = gw2 + round4(2 * kW * nq_pad);                           // [kW*nq_pad] pairs G/T/|q_j|
  float* rmax = gj2 + round4(2 * kW * nq_pad);                          // [Nq]  (default keep: the row's own softmax term - 1)
  float* rsum = rmax + round4(Nq);                                      // [Nq]
  float* qinv = rsum + round4(Nq);                                      // [nq_pad] 1/|q_j|, negative flags the eps clamp
  int* lab = reinterpret_cast<int*>(qinv + round4(nq_pad));             // [Nq]
  const int u_l = lane / kC, c_l = lane - u_l * kC;
  const bool lane_valid = lane < kVals;
  const uint32_t* keep_e = p.keep ? p.keep + (size_t)e * Nq * words : nullptr;

  const float* qry = p.queries + (size_t)e * Nq * kD;
  const int total = kBwd ? 2 * nq_pad : nq_pad;          // the backward walks the rows twice (once with saved similarities)
  auto row_ptr = [&](int t) -> const float* {
    const int r = t < nq_pad ? t : t - nq_pad;
    return qry + (size_t)min(r, Nq - 1) * kD;
  };

  // ------------------------------------------------------------------ normalised prototypes -> shared
  float pinv[kW];
  {
    f32x2 x[kW][kH];
#pragma unroll
    for (int w = 0; w < kW; ++w) load_row<kV>(p.protos + ((size_t)e * kW + w) * kD, lane, x[w]);
    for (int k = lane; k < Nq; k += 32) lab[k] = p.labels[(size_t)e * Nq + k];
    float ss[pow2_ceil(kW)];
#pragma unroll
    for (int w = 0; w < pow2_ceil(kW); ++w) ss[w] = 0.f;
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      f32x2 a = 0ull;
#pragma unroll
      for (int j = 0; j < kH; ++j) a = fma2(x[w][j], x[w][j], a);
      ss[w] = sum2(a);
    }
    const float tot = reduce_values<pow2_ceil(kW)>(ss, lane);
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      const float nrm = sqrtf(__shfl_sync(kFull, tot, w));
      const float r = 1.f / fmaxf(nrm, kCosEps);
      pinv[w] = nrm > kCosEps ? 1.f / nrm : -1.f / kCosEps;   // negative flags the clamp
      const f32x2 r2 = pack2(r, r);
#pragma unroll
      for (int j = 0; j < kH; ++j) x[w][j] = mul2(x[w][j], r2);
      sts_row<kV>(sp + w * kD, lane, x[w]);
    }
  }
  f32x2 buf[kB][kH];
#pragma unroll
  for (int u = 0; u < kB; ++u) load_row<kV>(row_ptr((kSaved ? nq_pad : 0) + u), lane, buf[u]);
  if (kSaved) {
    // the forward's similarity matrix and reciprocal norms: no dot-product walk over the queries
    const float* gs = p.sim_in + (size_t)e * kW * Nq;
    for (int k = lane; k < kW * Nq; k += 32) sim[k] = gs[k];
    for (int k = lane; k < nq_pad; k += 32) qinv[k] = k < Nq ? p.qinv_in[(size_t)e * Nq + k] : 1.f;
  }
  __syncwarp();

  // ------------------------------------------------------------------ pass 1: C[w,j] = <p^_w, q^_j> / T
  for (int t0 = 0; t0 < (kSaved ? 0 : nq_pad); t0 += kB) {
    float part[kN];
#pragma unroll
    for (int u = 0; u < kB; ++u) {
      f32x2 a = 0ull;
#pragma unroll
      for (int j = 0; j < kH; ++j) a = fma2(buf[u][j], buf[u][j], a);
      part[u * kC + kW] = sum2(a);
    }
#pragma unroll
    for (int w = 0; w < kW; ++w) {
      f32x2 pw[kH];
      lds_row<kV>(sp + w * kD, lane, pw);
#pragma unroll
      for (int u = 0; u < kB; ++u) {
        f32x2 a0 = 0ull, a1 = 0ull;
#pragma unroll
        for (int j = 0; j < kH; ++j) {
          if (j & 1) a1 = fma2(pw[j], buf[u][j], a1); else a0 = fma2(pw[j], buf[u][j], a0);
        }
        part[u * kC + w] = kH > 1 ? sum2(add2(a0, a1)) : sum2(a0);
      }
    }
#pragma unroll
    for (int u = 0; u < kB; ++u)
      if (t0 + u + kB < total) load_row<kV>(row_ptr(t0 + u + kB), lane, buf[u]);
#pragma unroll
    for (int k = kVals; k < kN; ++k) part[k] = 0.f;
    const float tot = reduce_values<kN>(part, lane);
    const float nrm = sqrtf(__shfl_sync(kFull, tot, u_l * kC + kW));
    const int j_l = t0 + u_l;
    if (lane_valid && j_l < Nq) {
      if (c_l < kW) sim[c_l * Nq + j_l] = __fdiv_rn(__fdiv_rn(tot, fmaxf(nrm, kCosEps)), p.temperature);
      else qinv[j_l] = nrm > kCosEps ? 1.f / nrm : -1.f / kCosEps;
    } else if (lane_valid && c_l == kW) {
      qinv[j_l] = 1.f;        // padding row of the last batch
    }
  }
  __syncwarp();
  if (!kBwd && p.sim_out) {
    float* gs = p.sim_out + (size_t)e * kW * Nq;
    for (int k = lane; k < kW * Nq; k += 32) gs[k] = sim[k];
    for (int k = lane; k < Nq; k += 32) p.qinv_out[(size_t)e * Nq + k] = qinv[k];
  }

  const float inv_t = 1.f / p.temperature;
  // g = d_loss / Nq^2 ; the stored pairs are g * dL/dC / T (and that times 1/|q_j|)
  const float g = kBwd ? p.d_loss[e] * (float)(1.0 / (double)Nq) / (float)Nq : 0.f;
  float loss_acc = 0.f;
  if (!keep_e) {
    // ---------------------------------------------------------------- default keep: j == i or another class
    // negatives of class w = all queries of other classes: max M_w and S_w = sum exp(C[w,j] - M_w) once per class
    float mneg[kW], sneg[kW], rcls[kW];
#pragma unroll
    for (int w = 0; w < kW; ++w) { mneg[w] = -INFINITY; sneg[w] = 0.f; rcls[w] = 0.f; }
    for (int j0 = 0; j0 < Nq; j0 += 32) {
      const int j = j0 + lane;
      const int lj = j < Nq ? lab[j] : -1;
#pragma unroll
      for (int w = 0; w < kW; ++w)
        if (j < Nq && lj != w) mneg[w] = fmaxf(mneg[w], sim[w * Nq + j]);
    }
#pragma unroll
    for (int w = 0; w < kW; ++w) mneg[w] = warp_max(mneg[w]);
    for (int j0 = 0; j0 < Nq; j0 += 32) {
      const int j = j0 + lane;
      const int lj = j < Nq ? lab[j] : -1;
#pragma unroll
      for (int w = 0; w < kW; ++w)
        if (j < Nq && lj != w) sneg[w] += expf(sim[w * Nq + j] - mneg[w]);
    }
#pragma unroll
    for (int w = 0; w < kW; ++w) sneg[w] = warp_sum(sneg[w]);
    // row i (lane i): m_i = max(M_w, C[w,i]) ; sum_i = exp(C[w,i] - m_i) + S_w exp(M_w - m_i)
    for (int i0 = 0; i0 < Nq; i0 += 32) {
      const int i = i0 + lane;
      const int y = i < Nq ? lab[i] : -1;
      const bool valid = y >= 0 && y < kW;
      float mw = mneg[0], sw = sneg[0];
#pragma unroll
      for (int v = 1; v < kW; ++v) { mw = y == v ? mneg[v] : mw; sw = y == v ? sneg[v] : sw; }
      const float cii = valid ? sim[y * Nq + i] : 0.f;
      const float m = fmaxf(mw, cii);
      const float own = expf(cii - m), rest = expf(mw - m);      // exp(-inf) = 0 when the class has no negatives
      const float se = own + (sw > 0.f ? sw * rest : 0.f);
      if (valid) loss_acc += -((cii - m) - logf(se));
      if (kBwd) {
        if (valid) rmax[i] = own / se - 1.f;                      // the row's own term of dL/dC[y_i, i]
        const float r = valid ? rest / se : 0.f;                  // its weight on every negative column
#pragma unroll
        for (int v = 0; v < kW; ++v) rcls[v] += y == v ? r : 0.f;
      }
    }
    if (kBwd) {
#pragma unroll
      for (int w = 0; w < kW; ++w) rcls[w] = warp_sum(rcls[w]);
      __syncwarp();
      // dL/dC[w,j] = sum_{i in w} exp(C[w,j] - m_i)/sum_i = exp(C[w,j] - M_w) * R_w for a negative column,
      //            = own term of row j                                              for lab_j == w
      for (int j0 = 0; j0 < nq_pad; j0 += 32) {
        const int j = j0 + lane;
        if (j < nq_pad) {
          const int lj = j < Nq ? lab[j] : -1;
          const float qa = j < Nq ? fabsf(qinv[j]) : 0.f;
#pragma unroll
          for (int w = 0; w < kW; ++w) {
            float t = 0.f;
            if (j < Nq) t = lj == w ? rmax[j] : expf(sim[w * Nq + j] - mneg[w]) * rcls[w];
            const float a = g * t * inv_t;
            reinterpret_cast<float2*>(gw2)[w * nq_pad + j] = make_float2(a, a);
            reinterpret_cast<float2*>(gj2)[w * nq_pad + j] = make_float2(a * qa, a * qa);
          }
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- sampled negatives: masked LSE of row i on lane i
    for (int i0 = 0; i0 < Nq; i0 += 32) {
      const int i = i0 + lane;
      const bool active = i < Nq;
      const int ic = active ? i : Nq - 1;
      const int y = lab[ic];
      const bool valid = active && y >= 0 && y < kW;
      const float* row = sim + (valid ? y : 0) * Nq;
      float m = -INFINITY, se = 0.f;
      for (int j0 = 0; j0 < Nq; j0 += 32) {
        const uint32_t word = keep_e[(size_t)ic * words + (j0 >> 5)];
        const int jn = min(32, Nq - j0);
        for (int jj = 0; jj < jn; ++jj)
          if ((word >> jj) & 1u) m = fmaxf(m, row[j0 + jj]);
      }
      for (int j0 = 0; j0 < Nq; j0 += 32) {
        const uint32_t word = keep_e[(size_t)ic * words + (j0 >> 5)];
        const int jn = min(32, Nq - j0);
        for (int jj = 0; jj < jn; ++jj)
          if ((word >> jj) & 1u) se += expf(row[j0 + jj] - m);
      }
      if (valid) {
        loss_acc += -((row[i] - m) - logf(se));
        if (kBwd) { rmax[i] = m; rsum[i] = se; }
      } else if (active && kBwd) {   // label without a prototype: row contributes nothing
        rmax[i] = 0.f; rsum[i] = 1.f;
      }
    }
    if (kBwd) {
      __syncwarp();
      // dL/dC[w,j] = sum_{i: y_i = w} ( keep_ij * softmax_i[j] - [j == i] ), column j on lane j
      for (int j0 = 0; j0 < nq_pad; j0 += 32) {
        const int j = j0 + lane;
        const bool active = j < Nq;
        const int jc = active ? j : Nq - 1;
        float cj[kW], acc[kW];
#pragma unroll
        for (int w = 0; w < kW; ++w) { cj[w] = sim[w * Nq + jc]; acc[w] = 0.f; }
        for (int i = 0; i < Nq; ++i) {          // members of every class in ascending order
          const int w = lab[i];
          if (w < 0 || w >= kW) continue;       // warp-uniform
          float c = cj[0];
#pragma unroll
          for (int v = 1; v < kW; ++v) c = w == v ? cj[v] : c;
          const bool k = (keep_e[(size_t)i * words + (jc >> 5)] >> (jc & 31)) & 1u;
          float t = k ? expf(c - rmax[i]) / rsum[i] : 0.f;
          if (jc == i) t -= 1.f;
#pragma unroll
          for (int v = 0; v < kW; ++v) acc[v] += w == v ? t : 0.f;
        }
        if (j < nq_pad) {
          const float qa = active ? fabsf(qinv[j]) : 0.f;
#pragma unroll
          for (int w = 0; w < kW; ++w) {
            const float a = active ? g * acc[w] * inv_t : 0.f;
            reinterpret_cast<float2*>(gw2)[w * nq_pad + j] = make_float2(a, a);
            reinterpret_cast<float2*>(gj2)[w * nq_pad + j] = make_float2(a * qa, a * qa);
          }
        }
      }
    }
  }
  if (!kBwd) {
    const float tot = warp_sum(loss_acc);
    // (1/Nq) * NLLLoss(mean): loops/loss.py:131
    if (lane == 0) p.loss[e] = (float)(1.0 / (double)Nq) * (tot / (float)Nq);
    return;
  }
  __syncwarp();

  // ------------------------------------------------------------------ pass 2: dQ rows out, dP^ in registers
  // queries:    dq^ = 1/T sum_w G[w,j] p^_w ;  dq = (dq^ - q^ <q^,dq^>) / |q|
  // prototypes: dp^ = 1/T sum_j G[w,j] q^_j (ascending j) ;  dp = (dp^ - p^ <p^,dp^>) / |p|
  f32x2 dPh[kW][kH];
#pragma unroll
  for (int w = 0; w < kW; ++w)
#pragma unroll
    for (int j = 0; j < kH; ++j) dPh[w][j] = 0ull;
  float* dq = p.d_queries + (size_t)e * Nq * kD;
  for (int t0 = nq_pad; t0 < total; t0 += kB) {
#pragma unroll
    for (int u = 0; u < kB; ++u) {
      const int j = t0 - nq_pad + u;
      const bool live = j < Nq;
      const float qi = qinv[j];
      const bool clamped = qi < 0.f;
      const float inv = fabsf(qi);
      const f32x2 inv2 = pack2(inv, inv);
      f32x2 qh[kH], out[kH];
#pragma unroll
      for (int k = 0; k < kH; ++k) { qh[k] = mul2(buf[u][k], inv2); out[k] = 0ull; }
#pragma unroll
      for (int w = 0; w < kW; ++w) {
        // packed (G/T, G/T) and (G/T/|q_j|, ..) pairs; zero in the padding columns of the last batch
        const f32x2 a2 = lds_pair(gw2 + 2 * (w * nq_pad + j)), b2 = lds_pair(gj2 + 2 * (w * nq_pad + j));
        f32x2 pw[kH];
        lds_row<kV>(sp + w * kD, lane, pw);
#pragma unroll
        for (int k = 0; k < kH; ++k) {
          out[k] = fma2(a2, pw[k], out[k]);
          dPh[w][k] = fma2(b2, buf[u][k], dPh[w][k]);
        }
      }
      f32x2 a = 0ull;
#pragma unroll
      for (int k = 0; k < kH; ++k) a = fma2(out[k], qh[k], a);
      float dot = warp_sum(sum2(a));
      if (clamped) dot = 0.f;  // norm clamped to eps: the denominator is a constant
      const f32x2 nd2 = pack2(-dot, -dot);
#pragma unroll
      for (int k = 0; k < kH; ++k) out[k] = mul2(fma2(nd2, qh[k], out[k]), inv2);
      if (live) store_row<kV>(dq + (size_t)j * kD, lane, out);
      if (t0 + u + kB < total) load_row<kV>(row_ptr(t0 + u + kB), lane, buf[u]);
    }
  }
#pragma unroll
  for (int w = 0; w < kW; ++w) {
    f32x2 ph[kH];
    lds_row<kV>(sp + w * kD, lane, ph);
    f32x2 a = 0ull;
#pragma unroll
    for (int k = 0; k < kH; ++k) a = fma2(dPh[w][k], ph[k], a);
    float dot = warp_sum(sum2(a));
    const bool clamped = pinv[w] < 0.f;
    if (clamped) dot = 0.f;
    const float inv = fabsf(pinv[w]);
    const f32x2 nd2 = pack2(-dot, -dot), inv2 = pack2(inv, inv);
#pragma unroll
    for (int k = 0; k < kH; ++k) dPh[w][k] = mul2(fma2(nd2, ph[k], dPh[w][k]), inv2);
    store_row<kV>(p.d_protos + ((size_t)e * kW + w) * kD, lane, dPh[w]);
  }
}

using KernelFn = void (*)(const CplParams);

template <int kW, int kV>
void variant(bool bwd, bool saved, KernelFn& fn) {
  constexpr int kB = 32 / (kW + 1) < 6 ? 32 / (kW + 1) : 6;
  fn = bwd ? (saved ? cpl_warp_kernel<kW, kV, kB, true, true> : cpl_warp_kernel<kW, kV, kB, true, false>)
           : cpl_warp_kernel<kW, kV, kB, false, false>;
}

bool pick_variant(int W, int D, bool bwd, bool saved, KernelFn& fn) {
#define AFSL_WV(W_, D_)                      \
  if (W == W_ && D == D_) {                  \
    variant<W_, D_ / 32>(bwd, saved, fn);    \
    return true;                             \
  }
  AFSL_WV(5, 256) AFSL_WV(5, 128) AFSL_WV(5, 64)
#undef AFSL_WV
  return false;
}

}  // namespace

static size_t warp_smem_bytes(int Nq, int W, int D) {
  const int kb = 32 / (W + 1) < 6 ? 32 / (W + 1) : 6;
  const int nq_pad = (Nq + kb - 1) / kb * kb;
  return (size_t)kWarpsPerCta * slice_words(W * D, W, Nq, nq_pad) * sizeof(float);
}

bool cpl_warp_supported(int Nq, int W, int D) {
  KernelFn fn = nullptr;
  return Nq > 0 && pick_variant(W, D, false, false, fn) && warp_smem_bytes(Nq, W, D) <= 64 * 1024;
}

int launch_cpl_warp(const CplParams& p, bool bwd, cudaStream_t stream, const char* name, bool* handled) {
  *handled = false;
  KernelFn fn = nullptr;
  if (!pick_variant(p.W, p.D, bwd, bwd && p.sim_in != nullptr, fn)) return AFSL_OK;
  const size_t bytes = warp_smem_bytes(p.Nq, p.W, p.D);
  if (bytes > 64 * 1024) return AFSL_OK;            // very long query lists: the CTA-per-episode kernel takes them
  *handled = true;
  if (int rc = opt_in_smem(fn, bytes, name)) return rc;
  const int grid = (p.E + kWarpsPerCta - 1) / kWarpsPerCta;
  fn<<<grid, kCtaThreads, bytes, stream>>>(p);
  AFSL_CHECK_LAUNCH(name);
  return AFSL_OK;
}

}  // namespace afsl
