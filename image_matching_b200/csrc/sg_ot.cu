// Log-domain Sinkhorn optimal transport and mutual-nearest-neighbour match selection.
// Reference: superglue/models/superglue_test.py:141-170 (log_sinkhorn_iterations, log_optimal_transport),
// :268-285 (match selection).  HBM/L2-bound streaming kernels: the (N+1)x(M+1) couplings matrix is never
// materialised -- the bin row/column are the scalar alpha -- and every sweep re-reads S (L2-resident when
// the caller micro-batches pairs).  LSE uses the same max-shifted two-pass form as torch.logsumexp.
#include <algorithm>
#include <type_traits>
#include "kernels.cuh"
#include "tc_common.cuh"

namespace b200m {

struct PairDims { int n, m; float norm, mu_bin, nu_bin; };

__device__ __forceinline__ PairDims pair_dims(const OtParams& p, int b) {
  PairDims d;
  d.n = p.counts0 ? p.counts0[b] : p.N;
  d.m = p.counts1 ? p.counts1[b] : p.M;
  d.norm = -logf((float)d.m + (float)d.n);          // norm = -(ms + ns).log()
  d.mu_bin = logf((float)d.m) + d.norm;             // log_mu[-1] = ns.log() + norm  (ns = #columns)
  d.nu_bin = logf((float)d.n) + d.norm;             // log_nu[-1] = ms.log() + norm  (ms = #rows)
  return d;
}

__global__ void ot_init_kernel(float* u, float* v, int total) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) { u[i] = 0.f; v[i] = 0.f; }
}

void launch_ot_init(LaunchCtx& ctx, const OtParams& p) {
  ProfScope prof__(ctx, "ot_init");
  int total = p.B * p.ld_uv;
  launch_pdl(ctx, kPdlPost, ot_init_kernel, dim3(cdiv(total, 256)), dim3(256), 0, p.u, p.v, total);
  B200M_LAUNCH_CHECK(ctx, "ot_init");
}

// u = log_mu - logsumexp(Z + v, dim=2): one warp per row (row n is the dustbin row).
__global__ void __launch_bounds__(256) ot_row_kernel(OtParams p) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const PairDims d = pair_dims(p, b);
  if (i > d.n || d.n == 0 || d.m == 0) return;
  const float* v = p.v + (size_t)b * p.ld_uv;
  const float* srow = p.S + (size_t)b * p.strideS + (size_t)i * p.ldS;
  const bool bin_row = (i == d.n);
  float mx = -INFINITY;
  for (int j = lane; j <= d.m; j += 32) {
    float c = (bin_row || j == d.m) ? p.alpha : srow[j];
    mx = fmaxf(mx, c + v[j]);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j <= d.m; j += 32) {
    float c = (bin_row || j == d.m) ? p.alpha : srow[j];
    sum += expf((c + v[j]) - mx);
  }
  sum = warp_sum(sum);
  if (lane == 0) p.u[(size_t)b * p.ld_uv + i] = (bin_row ? d.mu_bin : d.norm) - (logf(sum) + mx);
}

void launch_ot_row_update(LaunchCtx& ctx, const OtParams& p) {
  ProfScope prof__(ctx, "ot_row_update");
  dim3 grid(cdiv(p.N + 1, 8), p.B);
  ot_row_kernel<<<grid, 256, 0, ctx.stream>>>(p);
  B200M_LAUNCH_CHECK(ctx, "ot_row");
}

// v = log_nu - logsumexp(Z + u, dim=1): a block owns 32 columns, its 8 warps stride over the rows with
// coalesced 128 B row reads; per-column partial max / sum are combined through shared memory.
__global__ void __launch_bounds__(256) ot_col_kernel(OtParams p) {
  __shared__ float red[8][33];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const PairDims d = pair_dims(p, b);
  if (d.n == 0 || d.m == 0) return;               // uniform per block
  if (blockIdx.x * 32 > d.m) return;              // uniform per block
  const float* u = p.u + (size_t)b * p.ld_uv;
  const float* S = p.S + (size_t)b * p.strideS;
  const bool active = j <= d.m;
  const bool bin_col = (j == d.m);
  // one streaming pass with a running (max, sum) pair per column -- halves the L2 traffic of the column sweep;
  // the partial pairs of the 8 warps are merged exactly like a flash-softmax split
  float mx = -INFINITY, sum = 0.f;
  if (active)
    for (int i = w; i <= d.n; i += 8) {
      const float c = (bin_col || i == d.n) ? p.alpha : S[(size_t)i * p.ldS + j];
      const float x = c + u[i];
      if (x > mx) { sum = sum * expf(mx - x) + 1.f; mx = x; }
      else sum += expf(x - mx);
    }
  __shared__ float reds[8][33];
  red[w][lane] = mx;
  reds[w][lane] = sum;
  __syncthreads();
  if (w == 0 && active) {
    float M = red[0][lane];
#pragma unroll
    for (int k = 1; k < 8; ++k) M = fmaxf(M, red[k][lane]);
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += red[k][lane] == -INFINITY ? 0.f : reds[k][lane] * expf(red[k][lane] - M);
    p.v[(size_t)b * p.ld_uv + j] = (bin_col ? d.nu_bin : d.norm) - (logf(tot) + M);
  }
}

void launch_ot_col_update(LaunchCtx& ctx, const OtParams& p) {
  ProfScope prof__(ctx, "ot_col_update");
  dim3 grid(cdiv(p.M + 1, 32), p.B);
  ot_col_kernel<<<grid, 256, 0, ctx.stream>>>(p);
  B200M_LAUNCH_CHECK(ctx, "ot_col");
}

// ---------------------------------------------------------------------------------------------------------------
// Fused Sinkhorn iteration for M <= 1024 columns: ONE launch per iteration over all pairs, S is read from HBM once per
// iteration (the unfused pair of kernels above reads it three times), 5 instructions and 1 exp per matrix element.
//
// No running maxima: both logsumexps are shifted by RIGOROUS upper bounds that follow from the previous half-step,
//     u_i = log_mu_i - LSE_j'(c_ij' + v_j')  =>  c_ij + u_i <= log_mu_i - v_j <= mu_bin - v_j        (column shift)
//     v_j = log_nu_j - LSE_i'(c_i'j + u_i')  =>  c_ij + v_j <= log_nu_j - u_i <= nu_bin - u_i(prev)  (row shift)
// (mu_bin / nu_bin are the largest log-marginals), so no term can overflow; the shifts are lowered by kOtHeadroom = 60
// nats so that only entries 147 nats (e^-147 of a marginal) below the bound flush to zero -- a column / row whose
// every entry is that small is clamped to the smallest normal sum.  The first iteration (v = 0 did not come from a
// column update) takes the exact row maximum instead.  With a shift that depends on the column (row) only, partial
// sums of different rows simply ADD: the cross-warp / cross-CTA merges need no exps at all.
//
//   A CTA (8 warps) owns rows_per_cta consecutive rows of one pair (sized so that the whole grid is one resident wave:
//   6 CTAs x 176 rows per pair at 64 pairs), a warp an eighth of them.  Rows stream in through a per-warp
//   double-buffered cp.async.bulk ring (requested before the prologue).  A row (<= 1024 floats) is turned into
//   z_ij = (c_ij + v_j) log2(e) in 32 registers per lane; row sum: 2^(z + r_i) -> u_i; column sums: 2^(z + q_i) with
//   q_i = (u_i - mu_bin + 60) log2(e), accumulated in 32 registers per lane; the 8 warps' sums are added through shared
//   memory and written as ONE partial row; the LAST CTA of a pair to finish (atomic ticket) folds the <= 17 partial rows
//   into the new v.
constexpr int kOtFusedMaxM = 1024;
constexpr int kOtWideMaxM = 4096;        // ot_iter_wide_kernel (rows kept in shared memory)
constexpr int kOtMaxParts = 64;          // CTAs (= partial rows) per pair, upper bound (register-resident kernel)
constexpr int kOtWideMaxParts = 48;      // same for the wide kernels (few pairs of many rows: one CTA per SM)
constexpr int kOtCtasPerSm = 2;   // (3 per SM = 80 registers with spills: measured 1.71 -> 2.11 ms)
constexpr int kOtRing = 2;                // rows in flight per warp (3 measured no faster)
constexpr int kOtRowRingBytes = 8 * kOtRing * kOtFusedMaxM * 4;                             // 64 KB
constexpr int kOtSmemBytes = kOtRowRingBytes + (kOtFusedMaxM + 4) * 4 + 8 * kOtRing * 8;
constexpr float kOtNegBig = -1.0e30f;    // masked entries: 2^(anything this small) == 0
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kOtHeadroom = 60.f;      // nats
constexpr float kOtTiny = 1.17549435e-38f;

// CTAs per pair: the whole grid (parts x pairs) should be resident at once -- one wave, no tail, and the per-CTA
// prologue / merge cost is paid as few times as possible -- but never fewer than 2 rows per warp.  Few pairs (the
// reference caller's batch of ONE, superpoint_glue_test.py:66) get up to 64 CTAs each: with 20 (the former bound) a warp
// walked 7 rows one after the other and an iteration took 19 us at one pair.
int ot_parts_cap(int pairs) {            // what the scratch buffer is sized for
  const int by_slots = 320 / (pairs > 0 ? pairs : 1);
  return by_slots > kOtMaxParts ? kOtMaxParts : (by_slots < 20 ? 20 : by_slots);
}
int ot_fused_parts(int pairs, int N, int num_sms) {
  const int slots = num_sms * kOtCtasPerSm;
  int parts = slots / (pairs > 0 ? pairs : 1);
  parts = parts < 1 ? 1 : parts;
  const int cap = ot_parts_cap(pairs);
  parts = parts > cap ? cap : parts;
  const int by_rows = cdiv(N + 1, 16);
  return parts > by_rows ? by_rows : parts;
}

__device__ __forceinline__ float ot_ex2(float x) {    // 2^x, one MUFU op (ex2.approx.ftz: rel. error 2^-22)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256, kOtCtasPerSm) ot_iter_kernel(OtParams p, float* __restrict__ partials,
                                                                    int* __restrict__ tickets, int max_parts,
                                                                    int ld_part, int rows_per_cta, int first) {
  // dynamic shared memory: per-warp row ring (8 warps x 2 rows x 4 KB; later reused as the column-sum exchange),
  // v log2(e) (4 KB + pad), mbarriers
  extern __shared__ __align__(128) uint8_t ot_smem[];
  float* rows_s = reinterpret_cast<float*>(ot_smem);                                   // [8][kOtRing][1024]
  float* v2_s = reinterpret_cast<float*>(ot_smem + kOtRowRingBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(ot_smem + kOtRowRingBytes + (kOtFusedMaxM + 4) * 4);   // [8][kOtRing]
  __shared__ int s_ticket;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const PairDims d = pair_dims(p, b);
  const int parts = cdiv(d.n + 1, rows_per_cta);
  if (d.n == 0 || d.m == 0 || (int)blockIdx.x >= parts) {   // uniform per block
    pdl_wait();                                             // (no CTA leaves before the predecessor completed: the chain of
    return;                                                 // PDL waits stays transitive even if every CTA takes this exit)
  }
  // ---- this warp's rows stream in through kOtRing 4 KB buffers (cp.async.bulk + mbarrier): the first ones are requested
  // before the prologue so their DRAM latency hides behind it
  const int rows_per_warp = rows_per_cta >> 3;            // rows_per_cta is a multiple of 8
  const int row0 = blockIdx.x * rows_per_cta + warp * rows_per_warp;
  const int row_end = min(row0 + rows_per_warp, d.n + 1);
  const float* S = p.S + (size_t)b * p.strideS;
  float* wrow = rows_s + warp * kOtRing * kOtFusedMaxM;
  const uint32_t row_bytes = (uint32_t)p.ldS * 4;
  auto request = [&](int i) {                             // lane 0 only; the dustbin row (i == n) is not stored anywhere
    if (i < row_end && i < d.n) {
      uint64_t* bar = &bars[warp * kOtRing + (i - row0) % kOtRing];
      tc::mbar_expect_tx(bar, row_bytes);
      tc::bulk_load(wrow + ((i - row0) % kOtRing) * kOtFusedMaxM, S + (size_t)i * p.ldS, row_bytes, bar);
    }
  };
  // PDL: S does not change between iterations -- the first rows are requested while the PREVIOUS iteration's last CTAs
  // still fold their partial sums; u, v, partials and tickets are touched only behind pdl_wait().  This kernel triggers
  // AFTER its wait, so when a successor starts, everything up to this kernel's predecessor has completed (the first
  // iteration, whose predecessors wrote S, waits before it requests anything).
  if (first) pdl_wait();
  if (lane == 0) {
    for (int q = 0; q < kOtRing; ++q) tc::mbar_init(&bars[warp * kOtRing + q], 1);
    tc::fence_barrier_init();
    tc::fence_proxy_async();
    for (int q = 0; q < kOtRing; ++q) request(row0 + q);
  }
  if (!first) pdl_wait();
  pdl_trigger();
  __syncwarp();
  // ---- prologue: v of the previous iteration (zeros before the first one), in log2 units
  float* vrow = p.v + (size_t)b * p.ld_uv;
  for (int j = threadIdx.x; j < kOtFusedMaxM + 4; j += 256) v2_s[j] = j <= d.m ? vrow[j] * kLog2e : 0.f;
  __syncthreads();
  const float a2 = p.alpha * kLog2e;
  const float zbin = a2 + v2_s[d.m];                      // dustbin column entry of every row, log2 units
  const float mu_bin2 = d.mu_bin * kLog2e, head2 = kOtHeadroom * kLog2e;
  // columns of this lane: 4*lane + 128*k + {0..3}.  Groups k < kfull are valid in every lane, group kfull is the ragged
  // one (columns >= m masked), groups above it are skipped (warp-uniform conditions).
  const int kfull = d.m >> 7;
  float cs[32];
#pragma unroll
  for (int q = 0; q < 32; ++q) cs[q] = 0.f;
  float bs = 0.f;                                         // dustbin column sum (same in every lane)
  float* u = p.u + (size_t)b * p.ld_uv;
  auto run_rows = [&](auto full_tag) {
  constexpr bool FULL = decltype(full_tag)::value;   // every column group valid in every lane: straight-line code
  for (int i = row0; i < row_end; ++i) {
    const bool bin_row = (i == d.n);
    const int slot = (i - row0) % kOtRing;
    const float* srow = wrow + slot * kOtFusedMaxM + 4 * lane;
    // row shift r_i (log2 units): 60 nats above the negated bound nu_bin - u_i(prev)
    const float u_prev = u[i];
    if (!bin_row) tc::mbar_wait(&bars[warp * kOtRing + slot], ((i - row0) / kOtRing) & 1);
    float z[32];                                          // z_ij = (c_ij + v_j) log2(e)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (FULL || k <= kfull) {
        const int j = 4 * lane + 128 * k;
        float4 t = make_float4(p.alpha, p.alpha, p.alpha, p.alpha);
        if (!bin_row && (FULL || j < d.m)) t = *reinterpret_cast<const float4*>(srow + 128 * k);   // rows padded to ldS
        const float4 vv = *reinterpret_cast<const float4*>(v2_s + j);
        z[4 * k] = fmaf(t.x, kLog2e, vv.x); z[4 * k + 1] = fmaf(t.y, kLog2e, vv.y);
        z[4 * k + 2] = fmaf(t.z, kLog2e, vv.z); z[4 * k + 3] = fmaf(t.w, kLog2e, vv.w);
        if (!FULL && k == kfull) {
          z[4 * k] = j < d.m ? z[4 * k] : kOtNegBig;
          z[4 * k + 1] = j + 1 < d.m ? z[4 * k + 1] : kOtNegBig;
          z[4 * k + 2] = j + 2 < d.m ? z[4 * k + 2] : kOtNegBig;
          z[4 * k + 3] = j + 3 < d.m ? z[4 * k + 3] : kOtNegBig;
        }
      }
    }
    __syncwarp();                           // every lane has its copy of the row: the buffer can take the next one
    if (lane == 0) request(i + kOtRing);
    // ---- u_i = log_mu_i - logsumexp_j(c_ij + v_j), dustbin column included
    float r, rn;                            // the shift in log2 units and in nats
    if (first) {                            // warp-uniform: exact maximum (v = 0 carries no bound yet)
      float mx = zbin;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (FULL || k <= kfull)
          mx = fmaxf(mx, fmaxf(fmaxf(z[4 * k], z[4 * k + 1]), fmaxf(z[4 * k + 2], z[4 * k + 3])));
      r = -warp_max(mx);
      rn = r * kLn2;
    } else {
      rn = kOtHeadroom - (d.nu_bin - u_prev);
      r = rn * kLog2e;
    }
    const float ebin = ot_ex2(zbin + r);
    float s4[4] = {lane == 0 ? ebin : 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (FULL || k <= kfull) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          z[4 * k + e] = ot_ex2(z[4 * k + e] + r);        // e_ij = 2^(z_ij + r_i): kept for the column sums below
          s4[e] += z[4 * k + e];
        }
      }
    const float sum = fmaxf(warp_sum((s4[0] + s4[1]) + (s4[2] + s4[3])), kOtTiny);
    const float ui = (bin_row ? d.mu_bin : d.norm) - (logf(sum) - rn);
    if (lane == 0) u[i] = ui;
    // ---- column sums: 2^(y_ij - (mu_bin - v_j - 60)) log2 e) = 2^(z_ij + q_i), q_i = (u_i - mu_bin + 60) log2 e.
    // The exponent differs from the row term's only by the per-row constant q_i - r_i, so the second exp of every
    // element is one multiply-add with f_i = 2^(q_i - r_i) (clamped: no inf * 0): ONE MUFU op per matrix element.
    const float qi = fmaf(ui, kLog2e, head2 - mu_bin2);
    const float fi = ot_ex2(fminf(qi - r, 126.f));
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (FULL || k <= kfull) {
#pragma unroll
        for (int e = 0; e < 4; ++e) cs[4 * k + e] = fmaf(z[4 * k + e], fi, cs[4 * k + e]);
      }
    bs = fmaf(ebin, fi, bs);
  }
  };
  if (kfull >= 8) run_rows(std::true_type{});
  else run_rows(std::false_type{});
  // ---- add the 8 warps' column sums through shared memory (the exchange aliases the row ring: every warp must be
  // done with its rows first)
  __syncthreads();
  float* xch = rows_s + warp * (kOtFusedMaxM + 4);          // [8][1028]
#pragma unroll
  for (int k = 0; k < 8; ++k)
    *reinterpret_cast<float4*>(xch + 4 * lane + 128 * k) = make_float4(cs[4 * k], cs[4 * k + 1], cs[4 * k + 2], cs[4 * k + 3]);
  if (lane == 0) xch[kOtFusedMaxM] = bs;
  __syncthreads();
  float* prow = partials + ((size_t)b * max_parts + blockIdx.x) * ld_part;
  for (int j = threadIdx.x; j <= d.m; j += 256) {
    const int jj = j == d.m ? kOtFusedMaxM : j;           // the dustbin column sits in slot 1024 of the exchange
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += rows_s[w * (kOtFusedMaxM + 4) + jj];
    prow[j] = t;
  }
  __threadfence();                       // publish this CTA's partial row before taking a ticket
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[b], 1);
  __syncthreads();
  if (s_ticket != parts - 1) return;
  // ---- the last CTA of the pair folds the partial rows into v_j = log_nu_j - logsumexp_i(c_ij + u_i)
  //      = log_nu_j - (ln(sum_j) + mu_bin - v_j(old) - 60)
  __threadfence();
  // (thread = four adjacent columns, 16 partial rows in flight per thread; the partial rows are added in CTA order)
  const int ngroups = (d.m + 4) >> 2;                        // float4 groups covering columns 0 .. m (ld_part % 4 == 0)
  for (int g = threadIdx.x; g < ngroups; g += 256) {
    const float4* col = reinterpret_cast<const float4*>(partials + (size_t)b * max_parts * ld_part) + g;
    const size_t pitch4 = (size_t)(ld_part >> 2);
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int q0 = 0; q0 < parts; q0 += 16) {
      float4 tq[16];
#pragma unroll
      for (int q = 0; q < 16; ++q)
        tq[q] = q0 + q < parts ? __ldcg(col + (size_t)(q0 + q) * pitch4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int q = 0; q < 16; ++q) { t[0] += tq[q].x; t[1] += tq[q].y; t[2] += tq[q].z; t[3] += tq[q].w; }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * g + e;
      if (j <= d.m) {
        const float lse = logf(fmaxf(t[e], kOtTiny)) + ((d.mu_bin - vrow[j]) - kOtHeadroom);
        vrow[j] = (j == d.m ? d.nu_bin : d.norm) - lse;
      }
    }
  }
  if (threadIdx.x == 0) tickets[b] = 0;    // ready for the next iteration (next launch)
}

// ---------------------------------------------------------------------------------------------------------------
// The same fused iteration for 1024 < M <= 4096 columns (BASELINE configs 3 and 5: 2048 / 4096 keypoints).  A row no
// longer fits a lane's registers next to the column sums, so the row terms e_ij = 2^(z_ij + r_i) are written back IN
// PLACE into the row's shared-memory buffer (a lane only ever touches its own columns) and re-read for the column
// sums; the registers hold only the column sums.  TWO warps share a row (each owns one half of the columns: its own
// bulk-copy ring, its own 4G column sums) and exchange their partial row sums through shared memory and a 64-thread
// named barrier, so that 12 (M <= 4096) or 16 (M <= 2048) warps fit next to the ring instead of 6 / 8: with one or
// two warps per scheduler the kernel was bound by instruction latency, not by HBM.
// G = 128-column groups per WARP (a row has 2G), NW warps per CTA, one CTA per SM.
template <int G, int NW>
__global__ void __launch_bounds__(NW * 32, 1) ot_iter_wide_kernel(OtParams p, float* __restrict__ partials,
                                                                  int* __restrict__ tickets, int max_parts,
                                                                  int ld_part, int rows_per_cta, int first) {
  constexpr int HALF = 128 * G, MAXM = 2 * HALF, NT = NW * 32, NTEAM = NW / 2;
  extern __shared__ __align__(128) uint8_t ot_smem[];
  float* rows_s = reinterpret_cast<float*>(ot_smem);                                   // [NW][kOtRing][HALF]
  float* v2_s = rows_s + NW * kOtRing * HALF;                                          // [MAXM + 4]
  float* psum = v2_s + MAXM + 4;                                                       // [2][NW] partial row sums
  float* pmax = psum + 2 * NW;                                                         // [NW]   partial row maxima
  uint64_t* bars = reinterpret_cast<uint64_t*>(pmax + NW);                             // [NW][kOtRing]
  __shared__ int s_ticket;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int team = warp >> 1, side = warp & 1;
  const PairDims d = pair_dims(p, b);
  const int parts = cdiv(d.n + 1, rows_per_cta);
  if (d.n == 0 || d.m == 0 || (int)blockIdx.x >= parts) {   // uniform per block
    pdl_wait();                                             // (no CTA leaves before the predecessor completed: the chain of
    return;                                                 // PDL waits stays transitive even if every CTA takes this exit)
  }
  const int rows_per_team = rows_per_cta / NTEAM;         // rows_per_cta is a multiple of NTEAM
  const int row0 = blockIdx.x * rows_per_cta + team * rows_per_team;
  const int row_end = min(row0 + rows_per_team, d.n + 1);
  const int coff = side * HALF;                           // first column of this warp
  const int my_m = max(0, min(HALF, d.m - coff));         // valid columns of this warp
  const int my_ld = max(0, min(HALF, p.ldS - coff));      // floats of a (padded) row this warp loads
  const float* S = p.S + (size_t)b * p.strideS + coff;
  float* wrow = rows_s + warp * kOtRing * HALF;
  const uint32_t row_bytes = (uint32_t)my_ld * 4;
  auto request = [&](int i) {                             // lane 0 only; the dustbin row (i == n) is not stored anywhere
    if (i < row_end && i < d.n && my_ld > 0) {
      uint64_t* bar = &bars[warp * kOtRing + (i - row0) % kOtRing];
      tc::mbar_expect_tx(bar, row_bytes);
      tc::bulk_load(wrow + ((i - row0) % kOtRing) * HALF, S + (size_t)i * p.ldS, row_bytes, bar);
    }
  };
  // PDL: S does not change between iterations -- the first rows are requested while the PREVIOUS iteration's last CTAs
  // still fold their partial sums; u, v, partials and tickets are touched only behind pdl_wait().  This kernel triggers
  // AFTER its wait, so when a successor starts, everything up to this kernel's predecessor has completed (the first
  // iteration, whose predecessors wrote S, waits before it requests anything).
  if (first) pdl_wait();
  if (lane == 0) {
    for (int q = 0; q < kOtRing; ++q) tc::mbar_init(&bars[warp * kOtRing + q], 1);
    tc::fence_barrier_init();
    tc::fence_proxy_async();
    for (int q = 0; q < kOtRing; ++q) request(row0 + q);
  }
  if (!first) pdl_wait();
  pdl_trigger();
  __syncwarp();
  float* vrow = p.v + (size_t)b * p.ld_uv;
  for (int j0 = threadIdx.x; j0 < MAXM + 4; j0 += 8 * NT) {      // eight loads in flight per thread, then the stores
    float t8[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = j0 + q * NT;
      t8[q] = j <= d.m ? vrow[j] * kLog2e : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int j = j0 + q * NT;
      if (j < MAXM + 4) v2_s[j] = t8[q];
    }
  }
  __syncthreads();
  const float a2 = p.alpha * kLog2e;
  const float zbin = a2 + v2_s[d.m];
  const float mu_bin2 = d.mu_bin * kLog2e, head2 = kOtHeadroom * kLog2e;
  const int kfull = my_m >> 7;                            // this warp's groups k < kfull are full, group kfull is ragged (if any)
  const bool ragged = (d.m & 127) != 0;                   // block-uniform: both warps of a team take the same variant
  const int kend = kfull + ((my_m & 127) ? 1 : 0);        // groups to visit (warp-uniform), <= G
  const float* v2w = v2_s + coff;
  float cs[4 * G];
#pragma unroll
  for (int q = 0; q < 4 * G; ++q) cs[q] = 0.f;
  float bs = 0.f;
  float* u = p.u + (size_t)b * p.ld_uv;
  auto team_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(team + 1) : "memory"); };
  auto run_rows = [&](auto ragged_tag) {
  constexpr bool RAGGED = decltype(ragged_tag)::value;    // m a multiple of 128 (2048, 4096): no masking code at all
  for (int i = row0; i < row_end; ++i) {
    const bool bin_row = (i == d.n);
    const int slot = (i - row0) % kOtRing;
    float* srow = wrow + slot * HALF + 4 * lane;          // this lane's columns: coff + 4*lane + 128*k + {0..3}
    const float u_prev = u[i];
    if (!bin_row && my_ld > 0) tc::mbar_wait(&bars[warp * kOtRing + slot], ((i - row0) / kOtRing) & 1);
    // z_ij = (c_ij + v_j) log2(e) for group k (masked columns: -1e30 -> 2^z == 0)
    auto zgroup = [&](int k, float* z) {
      const int j = 4 * lane + 128 * k;                   // column inside this warp's half
      float4 t = make_float4(p.alpha, p.alpha, p.alpha, p.alpha);
      if (!bin_row && (!RAGGED || j < my_m)) t = *reinterpret_cast<const float4*>(srow + 128 * k);   // rows are padded to ldS
      const float4 vv = *reinterpret_cast<const float4*>(v2w + j);
      z[0] = fmaf(t.x, kLog2e, vv.x); z[1] = fmaf(t.y, kLog2e, vv.y);
      z[2] = fmaf(t.z, kLog2e, vv.z); z[3] = fmaf(t.w, kLog2e, vv.w);
      if (RAGGED && k == kfull) {                         // the ragged group (warp-uniform branch)
        z[0] = j < my_m ? z[0] : kOtNegBig; z[1] = j + 1 < my_m ? z[1] : kOtNegBig;
        z[2] = j + 2 < my_m ? z[2] : kOtNegBig; z[3] = j + 3 < my_m ? z[3] : kOtNegBig;
      }
    };
    float r, rn;
    if (first) {                                          // block-uniform: exact maximum (v = 0 carries no bound yet)
      float mx = zbin;
#pragma unroll 8
      for (int k = 0; k < G; ++k)
        if (k < kend) {
          float z[4];
          zgroup(k, z);
          mx = fmaxf(mx, fmaxf(fmaxf(z[0], z[1]), fmaxf(z[2], z[3])));
        }
      mx = warp_max(mx);
      if (lane == 0) pmax[warp] = mx;
      team_sync();
      r = -fmaxf(pmax[2 * team], pmax[2 * team + 1]);
      rn = r * kLn2;
    } else {
      rn = kOtHeadroom - (d.nu_bin - u_prev);
      r = rn * kLog2e;
    }
    const float ebin = ot_ex2(zbin + r);
    float s4[4] = {(lane == 0 && side == 0) ? ebin : 0.f, 0.f, 0.f, 0.f};
    // (unrolled: the independent groups are most of the latency hiding there is)
#pragma unroll 8
    for (int k = 0; k < G; ++k)
      if (k < kend) {
        float z[4];
        zgroup(k, z);
#pragma unroll
        for (int e = 0; e < 4; ++e) { z[e] = ot_ex2(z[e] + r); s4[e] += z[e]; }
        *reinterpret_cast<float4*>(srow + 128 * k) = make_float4(z[0], z[1], z[2], z[3]);   // e_ij, in place
      }
    const float part = warp_sum((s4[0] + s4[1]) + (s4[2] + s4[3]));
    float* ps = psum + ((i - row0) & 1) * NW;             // double-buffered: the partner is at most one barrier behind
    if (lane == 0) ps[warp] = part;
    team_sync();
    const float sum = fmaxf(ps[2 * team] + ps[2 * team + 1], kOtTiny);     // same order in both warps: same u_i
    const float ui = (bin_row ? d.mu_bin : d.norm) - (logf(sum) - rn);
    if (lane == 0 && side == 0) u[i] = ui;
    const float qi = fmaf(ui, kLog2e, head2 - mu_bin2);
    const float fi = ot_ex2(fminf(qi - r, 126.f));
#pragma unroll
    for (int k = 0; k < G; ++k)
      if (k < kend) {
        const float4 e4 = *reinterpret_cast<const float4*>(srow + 128 * k);
        cs[4 * k] = fmaf(e4.x, fi, cs[4 * k]); cs[4 * k + 1] = fmaf(e4.y, fi, cs[4 * k + 1]);
        cs[4 * k + 2] = fmaf(e4.z, fi, cs[4 * k + 2]); cs[4 * k + 3] = fmaf(e4.w, fi, cs[4 * k + 3]);
      }
    bs = fmaf(ebin, fi, bs);
    tc::fence_proxy_async();              // generic-proxy writes of this buffer are ordered before the next bulk copy
    __syncwarp();
    if (lane == 0) request(i + kOtRing);
  }
  };
  if (ragged) run_rows(std::true_type{});
  else run_rows(std::false_type{});
  // ---- add the teams' column sums through shared memory (the exchange aliases the row ring)
  __syncthreads();
  float* xch = rows_s + team * (MAXM + 4) + coff;           // [NTEAM][MAXM + 4]
#pragma unroll
  for (int k = 0; k < G; ++k)
    *reinterpret_cast<float4*>(xch + 4 * lane + 128 * k) = make_float4(cs[4 * k], cs[4 * k + 1], cs[4 * k + 2], cs[4 * k + 3]);
  if (lane == 0 && side == 0) rows_s[team * (MAXM + 4) + MAXM] = bs;
  __syncthreads();
  float* prow = partials + ((size_t)b * max_parts + blockIdx.x) * ld_part;
  for (int j = threadIdx.x; j <= d.m; j += NT) {
    const int jj = j == d.m ? MAXM : j;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < NTEAM; ++w) t += rows_s[w * (MAXM + 4) + jj];
    prow[j] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_ticket = atomicAdd(&tickets[b], 1);
  __syncthreads();
  if (s_ticket != parts - 1) return;
  __threadfence();
  // (up to 48 partial rows of up to 4097 columns: float4 columns, eight loads in flight per thread, summed in part order)
  const float4* pbase = reinterpret_cast<const float4*>(partials + (size_t)b * max_parts * ld_part);
  const int ld4 = ld_part >> 2;
  for (int j4 = threadIdx.x; 4 * j4 <= d.m; j4 += NT) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q0 = 0; q0 < parts; q0 += 8) {
      float4 x[8];
#pragma unroll
      for (int qq = 0; qq < 8; ++qq)
        x[qq] = q0 + qq < parts ? __ldcg(pbase + (size_t)(q0 + qq) * ld4 + j4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int qq = 0; qq < 8; ++qq) { t.x += x[qq].x; t.y += x[qq].y; t.z += x[qq].z; t.w += x[qq].w; }
    }
    const float tt[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int j = 4 * j4 + e;
      if (j <= d.m) {
        const float lse = logf(fmaxf(tt[e], kOtTiny)) + ((d.mu_bin - vrow[j]) - kOtHeadroom);
        vrow[j] = (j == d.m ? d.nu_bin : d.norm) - lse;
      }
    }
  }
  if (threadIdx.x == 0) tickets[b] = 0;
}

template <int G, int NW>
constexpr int ot_wide_smem_bytes() { return (NW * kOtRing * 128 * G + 2 * 128 * G + 4 + 3 * NW) * 4 + NW * kOtRing * 8 + 16; }

bool ot_fused_supported(const OtParams& p) {
  return p.M <= kOtWideMaxM && p.ldS % 4 == 0 && p.strideS % 4 == 0 && (reinterpret_cast<uintptr_t>(p.S) & 15) == 0;
}

// floats of scratch for `pairs` pairs: per-pair tickets + one set of partial rows
size_t ot_fused_scratch_floats(int pairs, int N, int M) {
  return (size_t)round_up(pairs, 64) + (size_t)pairs * (M > kOtFusedMaxM ? kOtWideMaxParts : ot_parts_cap(pairs)) * round_up(M + 1, 4);
}

// `iters` full Sinkhorn iterations (u update then v update) starting from u = v = 0 (launch_ot_init); leaves u and v in
// p.u / p.v.  One launch per iteration over all pairs.
void launch_ot_sinkhorn_fused(LaunchCtx& ctx, const OtParams& p, int iters, float* scratch, int num_sms) {
  if (iters <= 0) return;
  static SmemOptIn opt_a, opt_b, opt_c;
  opt_a.ensure(ot_iter_kernel, kOtSmemBytes);
  opt_b.ensure(ot_iter_wide_kernel<8, 16>, ot_wide_smem_bytes<8, 16>());
  opt_c.ensure(ot_iter_wide_kernel<16, 12>, ot_wide_smem_bytes<16, 12>());
  const int wide = p.M <= kOtFusedMaxM ? 0 : (p.M <= 2048 ? 1 : 2);
  int max_parts = std::max(1, std::min(kOtWideMaxParts, num_sms / std::max(p.B, 1)));   // one CTA per SM, one wave
  max_parts = std::min(max_parts, cdiv(p.N + 1, 24));
  if (!wide) max_parts = ot_fused_parts(p.B, p.N, num_sms);
  const int ld_part = round_up(p.M + 1, 4);
  // a multiple of the 8 warps (register-resident kernel) or of the 6 / 8 row teams (wide kernels)
  const int rows_per_cta = round_up(cdiv(p.N + 1, max_parts), wide ? 24 : 8);
  int* tickets = reinterpret_cast<int*>(scratch);
  float* partials = scratch + round_up(p.B, 64);
  cudaMemsetAsync(tickets, 0, sizeof(int) * p.B, ctx.stream);
  for (int it = 0; it < iters; ++it) {
    ProfScope prof__(ctx, "ot_iter_fused");
    dim3 grid(max_parts, p.B);
    if (wide == 0)
      launch_pdl(ctx, kPdlOt, ot_iter_kernel, grid, dim3(256), kOtSmemBytes, p, partials, tickets, max_parts, ld_part,
                 rows_per_cta, (int)(it == 0));
    else if (wide == 1)
      launch_pdl(ctx, kPdlOt, ot_iter_wide_kernel<8, 16>, grid, dim3(512), ot_wide_smem_bytes<8, 16>(), p, partials, tickets,
                 max_parts, ld_part, rows_per_cta, (int)(it == 0));
    else
      launch_pdl(ctx, kPdlOt, ot_iter_wide_kernel<16, 12>, grid, dim3(384), ot_wide_smem_bytes<16, 12>(), p, partials, tickets,
                 max_parts, ld_part, rows_per_cta, (int)(it == 0));
    B200M_LAUNCH_CHECK(ctx, "ot_iter_fused");
  }
}

// Z = couplings + u + v - norm, dense (stage API; full sizes)
__global__ void ot_write_Z_kernel(OtParams p, float* __restrict__ Z) {
  const int b = blockIdx.z, i = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > p.M) return;
  const PairDims d = pair_dims(p, b);
  float c = (i == p.N || j == p.M) ? p.alpha : p.S[(size_t)b * p.strideS + (size_t)i * p.ldS + j];
  float z = ((c + p.u[(size_t)b * p.ld_uv + i]) + p.v[(size_t)b * p.ld_uv + j]) - d.norm;
  Z[((size_t)b * (p.N + 1) + i) * (p.M + 1) + j] = z;
}

void launch_ot_write_Z(LaunchCtx& ctx, const OtParams& p, float* Z) {
  dim3 grid(cdiv(p.M + 1, 256), p.N + 1, p.B);
  ot_write_Z_kernel<<<grid, 256, 0, ctx.stream>>>(p, Z);
  B200M_LAUNCH_CHECK(ctx, "ot_write_Z");
}

// ---- argmax over rows / columns of Z[:, :-1, :-1]  (scores.max(2), scores.max(1); first index on ties)
struct ZSource {
  const float* S; int ldS; long long strideS; const float* u; const float* v; int ld_uv;
  const float* Z; int ldZ; long long strideZ;   // dense alternative
};
__device__ __forceinline__ float z_at(const ZSource& z, int b, int i, int j, float norm) {
  if (z.Z) return z.Z[(size_t)b * z.strideZ + (size_t)i * z.ldZ + j];
  float c = z.S[(size_t)b * z.strideS + (size_t)i * z.ldS + j];
  return ((c + z.u[(size_t)b * z.ld_uv + i]) + z.v[(size_t)b * z.ld_uv + j]) - norm;
}

__global__ void __launch_bounds__(256) row_argmax_kernel(ZSource z, const int* counts0, const int* counts1,
                                                         int N, int M, int* __restrict__ idx0,
                                                         float* __restrict__ max0, int ld) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  const int b = blockIdx.y;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  const int n = counts0 ? counts0[b] : N, m = counts1 ? counts1[b] : M;
  if (i >= n || m == 0) return;
  const float norm = -logf((float)m + (float)n);
  float best = -INFINITY;
  int bj = 0x7fffffff;
  for (int j = lane; j < m; j += 32) {
    float val = z_at(z, b, i, j, norm);
    if (val > best) { best = val; bj = j; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, best, o);
    int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ov > best || (ov == best && oj < bj)) { best = ov; bj = oj; }
  }
  // (an all-NaN row never beats -inf and keeps the sentinel: report index 0 so match_select stays in bounds)
  if (lane == 0) { idx0[(size_t)b * ld + i] = bj == 0x7fffffff ? 0 : bj; max0[(size_t)b * ld + i] = best; }
}

__global__ void __launch_bounds__(256) col_argmax_kernel(ZSource z, const int* counts0, const int* counts1,
                                                         int N, int M, int* __restrict__ idx1, int ld) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  __shared__ float rv[8][33];
  __shared__ int ri[8][33];
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int n = counts0 ? counts0[b] : N, m = counts1 ? counts1[b] : M;
  if (n == 0 || blockIdx.x * 32 >= m) return;   // uniform per block
  const float norm = -logf((float)m + (float)n);
  float best = -INFINITY;
  int bi = 0x7fffffff;
  if (j < m)
    for (int i = w; i < n; i += 8) {
      float val = z_at(z, b, i, j, norm);
      if (val > best) { best = val; bi = i; }
    }
  rv[w][lane] = best;
  ri[w][lane] = bi;
  __syncthreads();
  if (w == 0 && j < m) {
    for (int k = 1; k < 8; ++k) {
      float ov = rv[k][lane];
      int oi = ri[k][lane];
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    idx1[(size_t)b * ld + j] = bi == 0x7fffffff ? 0 : bi;
  }
}

static void launch_argmax_common(LaunchCtx& ctx, const ZSource& z, const int* c0, const int* c1, int B, int N,
                                 int M, int* idx0, float* max0, int* idx1, int ld) {
  ProfScope prof__(ctx, "argmax");
  if (N <= 0 || M <= 0) return;
  dim3 g0(cdiv(N, 8), B);
  launch_pdl(ctx, kPdlPost, row_argmax_kernel, dim3(g0), dim3(256), 0, z, c0, c1, N, M, idx0, max0, ld);
  B200M_LAUNCH_CHECK(ctx, "row_argmax");
  dim3 g1(cdiv(M, 32), B);
  launch_pdl(ctx, kPdlPost, col_argmax_kernel, dim3(g1), dim3(256), 0, z, c0, c1, N, M, idx1, ld);
  B200M_LAUNCH_CHECK(ctx, "col_argmax");
}

void launch_ot_argmax(LaunchCtx& ctx, const OtParams& p, int* idx0, float* max0, int* idx1) {
  ZSource z{p.S, p.ldS, p.strideS, p.u, p.v, p.ld_uv, nullptr, 0, 0};
  launch_argmax_common(ctx, z, p.counts0, p.counts1, p.B, p.N, p.M, idx0, max0, idx1, p.ld_uv);
}

void launch_dense_argmax(LaunchCtx& ctx, const float* Z, int B, int N, int M, int* idx0, float* max0, int* idx1,
                         int ld) {
  ZSource z{nullptr, 0, 0, nullptr, nullptr, 0, Z, M + 1, (long long)(N + 1) * (M + 1)};
  launch_argmax_common(ctx, z, nullptr, nullptr, B, N, M, idx0, max0, idx1, ld);
}

// mutual check + exp + threshold (:270-278); indices widened to int64 at the boundary
__global__ void match_select_kernel(const int* __restrict__ idx0, const float* __restrict__ max0,
                                    const int* __restrict__ idx1, int ld, const int* counts0, const int* counts1,
                                    int N, int M, float thr, long long* __restrict__ matches0,
                                    long long* __restrict__ matches1, float* __restrict__ ms0,
                                    float* __restrict__ ms1) {
  pdl_trigger();   // PDL (common.cuh): the next kernel may be scheduled; nothing is read or written before the wait
  pdl_wait();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = counts0 ? counts0[b] : N, m = counts1 ? counts1[b] : M;
  const int* i0 = idx0 + (size_t)b * ld;
  const int* i1 = idx1 + (size_t)b * ld;
  const float* mx = max0 + (size_t)b * ld;
  const bool empty = (n == 0 || m == 0);
  if (t < N) {
    long long mt = -1;
    float sc = 0.f;
    if (!empty && t < n) {
      int j = i0[t];
      bool mutual = (i1[j] == t);
      sc = mutual ? expf(mx[t]) : 0.f;
      if (mutual && sc > thr) mt = j;
    }
    matches0[(size_t)b * N + t] = mt;
    ms0[(size_t)b * N + t] = sc;
  }
  if (t < M) {
    long long mt = -1;
    float sc = 0.f;
    if (!empty && t < m) {
      int i = i1[t];
      bool mutual1 = (i0[i] == t);
      // mscores0[i] and valid0[i] of the row this column points at
      bool mutual0 = (i1[i0[i]] == i);
      float s0 = mutual0 ? expf(mx[i]) : 0.f;
      sc = mutual1 ? s0 : 0.f;
      if (mutual1 && mutual0 && s0 > thr) mt = i;
    }
    matches1[(size_t)b * M + t] = mt;
    ms1[(size_t)b * M + t] = sc;
  }
}

void launch_match_select(LaunchCtx& ctx, const int* idx0, const float* max0, const int* idx1, int ld,
                         const int* counts0, const int* counts1, int B, int N, int M, float thr,
                         long long* matches0, long long* matches1, float* ms0, float* ms1) {
  ProfScope prof__(ctx, "match_select");
  int T = N > M ? N : M;
  if (T <= 0) return;
  dim3 grid(cdiv(T, 256), B);
  launch_pdl(ctx, kPdlPost, match_select_kernel, dim3(grid), dim3(256), 0, idx0, max0, idx1, ld, counts0, counts1, N, M, thr, matches0,
                                                    matches1, ms0, ms1);
  B200M_LAUNCH_CHECK(ctx, "match_select");
}

// ---- multi-GPU wire format of the final gather (image_matching_b200/dist.py): ONE int32 buffer per rank,
// [pair][0][N] = match index (int64 -> int32), [pair][1][N] = matching score bits; pairs >= B_valid are padding (-1 / 0)
__global__ void pack_match_wire_kernel(const long long* __restrict__ matches, const float* __restrict__ scores,
                                       int B_valid, int N, int ld, int* __restrict__ wire, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t b = i / N, t = i - b * N;
  const bool ok = (int)b < B_valid;
  wire[(b * 2) * N + t] = ok ? (int)matches[b * ld + t] : -1;
  wire[(b * 2 + 1) * N + t] = ok ? __float_as_int(scores[b * ld + t]) : 0;
}
// gathered wire (ranks x bmax pairs) -> (n_pairs, N) int64 matches + fp32 scores; rank r owns `base + (r < rem)` pairs
__global__ void unpack_match_wire_kernel(const int* __restrict__ wire, int world, int bmax, int base, int rem, int N,
                                         long long* __restrict__ matches, float* __restrict__ scores, size_t total) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const size_t g = i / N, t = i - g * N;                 // g = global pair index
  // contiguous shards, remainder pairs on the first ranks (dist.shard_range)
  const size_t big = (size_t)rem * (base + 1);
  const size_t r = g < big ? g / (base + 1) : rem + (g - big) / (base > 0 ? base : 1);
  const size_t lo = r < (size_t)rem ? r * (base + 1) : big + (r - rem) * base;
  const size_t src = (r * bmax + (g - lo)) * 2;
  matches[i] = wire[src * N + t];
  scores[i] = __int_as_float(wire[(src + 1) * N + t]);
}
void launch_pack_match_wire(LaunchCtx& ctx, const long long* matches, const float* scores, int B_valid, int B_wire,
                            int N, int ld, int* wire) {
  const size_t total = (size_t)B_wire * N;
  if (!total) return;
  ProfScope prof__(ctx, "match_wire");
  pack_match_wire_kernel<<<(unsigned)cdivz(total, 256), 256, 0, ctx.stream>>>(matches, scores, B_valid, N, ld, wire, total);
  B200M_LAUNCH_CHECK(ctx, "pack_match_wire");
}
void launch_unpack_match_wire(LaunchCtx& ctx, const int* wire, int world, int bmax, int n_pairs, int N,
                              long long* matches, float* scores) {
  const size_t total = (size_t)n_pairs * N;
  if (!total) return;
  ProfScope prof__(ctx, "match_wire");
  unpack_match_wire_kernel<<<(unsigned)cdivz(total, 256), 256, 0, ctx.stream>>>(wire, world, bmax, n_pairs / world,
                                                                               n_pairs % world, N, matches, scores, total);
  B200M_LAUNCH_CHECK(ctx, "unpack_match_wire");
}

}  // namespace b200m
