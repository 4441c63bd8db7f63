// C ABI of libb200match.so: handle, weight packing and the orchestration of the kernel pipeline.
// See include/b200m.h for the reference file:line each entry point replaces.
#include <algorithm>
#include <map>
#include <string>
#include <vector>
#include <cstring>
#include <cmath>
#include <cstdarg>

#include "../../include/b200m.h"
#include "kernels.cuh"

using namespace b200m;

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

struct HostTensor {
  std::vector<int64_t> shape;
  std::vector<float> data;
  size_t numel() const { size_t n = 1; for (auto s : shape) n *= (size_t)s; return n; }
};

struct ConvLayer {      // packed conv on C4-planar activations
  int cin = 0, cout = 0, cout_pad = 0, ks = 3;
  size_t w_off = 0, b_off = 0;   // float offsets into the device weight arena
  int nb = 0;                    // tensor-core path: output channels per CTA tile (0 = no tc weights)
  size_t tc_w_off = 0;
  std::vector<float> bias_host;  // folded bias [cout_pad] (copied into the kernel parameters of small layers)
};
struct Linear {         // packed [N][K] row-major weight + bias
  int N = 0, K = 0;
  size_t w_off = 0, b_off = 0;
  size_t w_hi_off = 0, w_lo_off = 0;   // fp16 hi / lo*2048 planes of the same matrix (tensor-core path)
};

constexpr int kHeads = 4;
constexpr int kSpMicroBatch = 64;   // images per SuperPoint micro-batch (bounds the activation arena; 64 images x 20 tiles
                                    // of the 60x80 layers = 8.65 waves of 148 CTAs, against 2.16 at 16 images)

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Arena {          // bump allocator over the caller's workspace
  char* base; size_t size; size_t off = 0; bool ok = true;
  Arena(void* p, size_t n) : base((char*)p), size(n) {}
  template <typename T> T* take(size_t count) {
    off = align_up(off, 256);
    size_t bytes = count * sizeof(T);
    if (!base || off + bytes > size) { ok = false; off += bytes; return nullptr; }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
};

int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

}  // namespace

struct b200m_handle {
  b200m_config cfg;
  int device = 0;
  std::map<std::string, HostTensor> tensors;
  bool packed_sp = false, packed_sg = false;   // which half b200m_pack found weights for
  float* d_w = nullptr;          // device weight arena
  // SuperPoint
  size_t conv1_w = 0, conv1_b = 0;
  std::vector<float> stem_host;  // [9][64] weights | [64] bias of the first conv (kernel-parameter copy for the fused stem)
  ConvLayer c1b, c2a, c2b, c3a, c3b, c4a, c4b, heads, pb, db;
  // SuperGlue
  std::vector<Linear> kenc;
  struct Gnn {
    Linear qkv, merge, mlp1, mlp2;
    size_t fused_w_off = 0, fused_b_off = 0;   // tc_gnn.cu weight stream / bias block (D = 128 only)
    bool fused = false, fused_has_qkv = false;
  };
  std::vector<Gnn> gnn;
  Linear final_proj;
  float bin_score = 1.f;
  long long launches = 0;
  const char* err_where = nullptr;
  Profiler prof;
  std::vector<ProfRecord> prof_recs;
  bool use_tc = true;            // tcgen05 fp16x3 convolutions (B200M_CONV_IMPL=simt selects the fp32 CUDA-core path)
  bool use_tc_attn = true;       // tcgen05 flash attention (B200M_ATTN_IMPL=simt selects the fp32 CUDA-core kernel)
  bool use_tc_gemm = true;       // tcgen05 linear layers (B200M_GEMM_IMPL=simt selects the fp32 CUDA-core GEMM)
  bool use_fused_stem = true;    // first conv computed inside the second conv's kernel (B200M_STEM_IMPL=unfused: two kernels)
  bool use_fused_gnn = true;     // fused merge/mlp/residual/q|k|v layer kernel (B200M_GNN_IMPL=unfused: four GEMM launches)
  // precision experiment, never set in the product: B200M_SINGLE=desc,gemm,gnn,attn (any subset) drops the lo planes'
  // products in the descriptor head / SuperGlue GEMMs / fused layer / attention (profiles/r02_single_product_experiment.md)
  bool single_desc = false, single_gemm = false, single_gnn = false, single_attn = false;
  int num_sms = 148;
  int sp_micro_batch = 0;        // > 0: B200M_SP_MICROBATCH override of the images per SuperPoint micro-batch
  // CUDA-graph replay of whole forward calls (see with_graph below)
  struct GraphEntry {
    std::vector<uint64_t> key;
    cudaGraphExec_t exec = nullptr;
    long long launches = 0;
    unsigned long long last_use = 0;
    bool dead = false;           // capture failed once: always run eagerly
  };
  std::vector<GraphEntry> graphs;
  bool use_graphs = true;        // B200M_GRAPHS=0 disables
  cudaStream_t gstream = nullptr;
  cudaEvent_t g_in = nullptr, g_out = nullptr;
  unsigned long long graph_clock = 0;
  long long graph_replays = 0;
  // small-batch latency (the reference caller's batch of one pair, superpoint_glue_test.py:66):
  int pdl_max_pairs = 8;         // calls of at most this many pairs / images launch with PDL (B200M_PDL_MAX_PAIRS)
  bool pdl_now = false;          // this call's setting (make_ctx)
  int sp_dual_max = 16;          // Matching.forward of at most this many pairs runs the two images' SuperPoint passes
                                 // concurrently (forked side stream, second workspace; B200M_SP_DUAL_MAX, 0 = never).
                                 // Measured (ms per call, off -> on): 1 pair 1.68 -> 1.50, 8 pairs 3.94 -> 3.75,
                                 // 16 pairs 7.24 -> 7.01, 64 pairs 27.04 -> 26.95 (not worth the second workspace)
  cudaStream_t side_stream = nullptr;
  cudaEvent_t fork_ev = nullptr, join_ev = nullptr;
};

namespace {

// Every entry point runs on the handle's device and leaves the caller's current device untouched.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

LaunchCtx make_ctx(b200m_handle* h, void* stream) {
  LaunchCtx c;
  c.stream = (cudaStream_t)stream;
  c.counter = &h->launches;
  c.err_where = &h->err_where;
  c.err = cudaSuccess;
  c.prof = &h->prof;
  c.pdl = h->pdl_now;
  return c;
}

int finish(b200m_handle* h, LaunchCtx& ctx) {
  if (ctx.err != cudaSuccess)
    return fail(B200M_ERR_CUDA, "CUDA launch failed in %s: %s", h->err_where ? h->err_where : "?",
                cudaGetErrorString(ctx.err));
  return B200M_OK;
}

// ------------------------------------------------------------------ CUDA-graph replay of a forward call
// A forward call is ~60-120 kernel launches whose arguments are a pure function of the entry point's arguments (shapes,
// device pointers, workspace) and of the handle's packed weights.  The second time an entry point sees the SAME
// arguments, its launch sequence is captured into a CUDA graph (on a handle-owned stream: torch's default stream is the
// legacy stream, which cannot be captured) and from then on replayed with one cudaGraphLaunch: no per-launch CPU work
// (tensor-map encoding, argument marshalling) and back-to-back kernel scheduling on the device.  The reference's caller
// is a batch_size = 1 loop (superpoint_glue_test.py:65-78), where the ~100 launches of one pair are launch-bound.
// Semantics of include/b200m.h are unchanged: work is ordered after everything already enqueued on the caller's stream
// and before everything enqueued later (event hand-over), outputs land in the caller's buffers.  Eager fallback: first
// sight of a key, per-launch profiling, a caller stream that is itself being captured, B200M_GRAPHS=0, capture failure.
void drop_graphs(b200m_handle* h) {
  for (auto& e : h->graphs)
    if (e.exec) cudaGraphExecDestroy(e.exec);
  h->graphs.clear();
}

template <typename Body>
int with_graph(b200m_handle* h, void* user_stream, std::vector<uint64_t> key, Body body) {
  cudaStream_t us = (cudaStream_t)user_stream;
  if (!h->use_graphs || h->prof.enabled) return body(user_stream);
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(us, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) {
    cudaGetLastError();
    return body(user_stream);
  }
  b200m_handle::GraphEntry* e = nullptr;
  for (auto& g : h->graphs)
    if (g.key == key) { e = &g; break; }
  if (!e) {                                 // first sight: remember the key, run eagerly
    if (h->graphs.size() >= 8) {            // evict the least recently used entry
      size_t lru = 0;
      for (size_t i = 1; i < h->graphs.size(); ++i)
        if (h->graphs[i].last_use < h->graphs[lru].last_use) lru = i;
      if (h->graphs[lru].exec) cudaGraphExecDestroy(h->graphs[lru].exec);
      h->graphs.erase(h->graphs.begin() + lru);
    }
    b200m_handle::GraphEntry ne;
    ne.key = std::move(key);
    ne.last_use = ++h->graph_clock;
    h->graphs.push_back(std::move(ne));
    return body(user_stream);
  }
  e->last_use = ++h->graph_clock;
  if (e->dead) return body(user_stream);
  if (!h->gstream) {
    if (cudaStreamCreateWithFlags(&h->gstream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->g_in, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->g_out, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      h->use_graphs = false;
      return body(user_stream);
    }
  }
  if (!e->exec) {                           // second sight: capture
    const long long l0 = h->launches;
    if (cudaStreamBeginCapture(h->gstream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      e->dead = true;
      return body(user_stream);
    }
    const int rc = body((void*)h->gstream);
    cudaGraph_t g = nullptr;
    const cudaError_t ee = cudaStreamEndCapture(h->gstream, &g);
    if (rc != B200M_OK || ee != cudaSuccess || !g || cudaGraphInstantiate(&e->exec, g, 0) != cudaSuccess) {
      if (g) cudaGraphDestroy(g);
      cudaGetLastError();
      e->exec = nullptr;
      e->dead = true;
      h->launches = l0;
      return rc != B200M_OK ? rc : body(user_stream);
    }
    cudaGraphDestroy(g);
    e->launches = h->launches - l0;
    h->launches = l0;
  }
  cudaError_t err = cudaEventRecord(h->g_in, us);
  if (err == cudaSuccess) err = cudaStreamWaitEvent(h->gstream, h->g_in, 0);
  if (err == cudaSuccess) err = cudaGraphLaunch(e->exec, h->gstream);
  if (err == cudaSuccess) err = cudaEventRecord(h->g_out, h->gstream);
  if (err == cudaSuccess) err = cudaStreamWaitEvent(us, h->g_out, 0);
  if (err != cudaSuccess) return fail(B200M_ERR_CUDA, "CUDA graph replay failed: %s", cudaGetErrorString(err));
  h->launches += e->launches;
  ++h->graph_replays;
  return B200M_OK;
}

inline uint64_t kp(const void* p) { return (uint64_t)reinterpret_cast<uintptr_t>(p); }
inline uint64_t ki(long long a, long long b = 0) { return ((uint64_t)(uint32_t)a << 32) | (uint32_t)b; }

// ------------------------------------------------------------------ packing helpers
struct Packer {
  b200m_handle* h;
  std::vector<float> host;
  std::string err;
  const HostTensor* get(const std::string& name, std::initializer_list<int64_t> shape) {
    auto it = h->tensors.find(name);
    if (it == h->tensors.end()) { if (err.empty()) err = "missing tensor " + name; return nullptr; }
    const HostTensor& t = it->second;
    std::vector<int64_t> want(shape);
    // accept trailing singleton dims (Conv1d weights are [out,in,1])
    std::vector<int64_t> got = t.shape;
    while (got.size() > want.size() && got.back() == 1) got.pop_back();
    if (got != want) { if (err.empty()) err = "bad shape for " + name; return nullptr; }
    return &t;
  }
  size_t alloc(size_t n) {
    size_t off = align_up(host.size(), 64);
    host.resize(off + n, 0.f);
    return off;
  }
  // conv/linear weight [out][in*k*k] and bias folded with an optional BatchNorm (eval, eps 1e-5)
  bool folded(const std::string& conv, const std::string& bn, int out, int in_kk, std::vector<double>& w,
              std::vector<double>& b, std::initializer_list<int64_t> wshape) {
    const HostTensor* W = get(conv + ".weight", wshape);
    const HostTensor* Bv = get(conv + ".bias", {out});
    if (!W || !Bv) return false;
    w.assign(W->data.begin(), W->data.end());
    b.assign(Bv->data.begin(), Bv->data.end());
    // a conv without BatchNorm tensors is used as is: the "official" SuperPoint variant
    // (superglue/models/superpoint.py:95-202) has the same topology with no normalisation layers
    if (!bn.empty() && h->tensors.count(bn + ".weight")) {
      const HostTensor* g = get(bn + ".weight", {out});
      const HostTensor* be = get(bn + ".bias", {out});
      const HostTensor* mu = get(bn + ".running_mean", {out});
      const HostTensor* var = get(bn + ".running_var", {out});
      if (!g || !be || !mu || !var) return false;
      for (int o = 0; o < out; ++o) {
        double s = (double)g->data[o] / std::sqrt((double)var->data[o] + 1e-5);
        for (int k = 0; k < in_kk; ++k) w[(size_t)o * in_kk + k] *= s;
        b[o] = (b[o] - (double)mu->data[o]) * s + (double)be->data[o];
      }
    }
    return true;
  }
  // append a conv layer from already-folded [cout][cin][ks][ks] weights
  ConvLayer pack_conv(const std::vector<double>& w, const std::vector<double>& b, int cout, int cin, int ks) {
    ConvLayer L;
    L.cin = cin; L.cout = cout; L.ks = ks; L.cout_pad = round_up(cout, 64);
    const int taps = ks * ks, nchunks = cin / 8, ncb = L.cout_pad / 64;
    L.w_off = alloc((size_t)ncb * nchunks * taps * 8 * 64);
    L.b_off = alloc(L.cout_pad);
    for (int cb = 0; cb < ncb; ++cb)
      for (int cc = 0; cc < nchunks; ++cc)
        for (int t = 0; t < taps; ++t)
          for (int ci = 0; ci < 8; ++ci)
            for (int co = 0; co < 64; ++co) {
              int o = cb * 64 + co, i = cc * 8 + ci;
              float v = o < cout ? (float)w[((size_t)o * cin + i) * taps + t] : 0.f;
              host[L.w_off + ((((size_t)cb * nchunks + cc) * taps + t) * 8 + ci) * 64 + co] = v;
            }
    for (int o = 0; o < cout; ++o) host[L.b_off + o] = (float)b[o];
    L.bias_host.assign(host.begin() + L.b_off, host.begin() + L.b_off + L.cout_pad);
    return L;
  }
  void pack_tc(ConvLayer& L, const std::vector<double>& w) {   // 3x3 / 1x1 layers with cin % 32 == 0
    L.nb = L.cout_pad >= 128 ? 128 : 64;
    L.tc_w_off = alloc(tc_conv_weight_floats(L.cin, L.cout_pad, L.nb, L.ks));
    tc_conv_pack_weights(w.data(), L.cout, L.cin, L.cout_pad, L.nb, L.ks, host.data() + L.tc_w_off);
  }
  Linear pack_linear(const std::vector<double>& w, const std::vector<double>& b, int N, int K, int Kpad,
                     const int* row_perm = nullptr, const int* col_perm = nullptr) {
    Linear L;
    L.N = N; L.K = Kpad;
    L.w_off = alloc((size_t)N * Kpad);
    L.b_off = alloc(N);
    for (int r = 0; r < N; ++r) {
      int sr = row_perm ? row_perm[r] : r;
      for (int c = 0; c < K; ++c) {
        int scol = col_perm ? col_perm[c] : c;
        host[L.w_off + (size_t)r * Kpad + c] = (float)w[(size_t)sr * K + scol];
      }
      host[L.b_off + r] = (float)b[sr];
    }
    // fp16 hi / lo*2048 planes of the same matrix for the tensor-core GEMM (halves stored in the float arena)
    L.w_hi_off = alloc(((size_t)N * Kpad + 1) / 2);
    L.w_lo_off = alloc(((size_t)N * Kpad + 1) / 2);
    gemm_pack_fp16_planes(host.data() + L.w_off, (size_t)N * Kpad, host.data() + L.w_hi_off, host.data() + L.w_lo_off);
    return L;
  }
};

bool has_prefix(const b200m_handle* h, const std::string& prefix) {
  auto it = h->tensors.lower_bound(prefix);
  return it != h->tensors.end() && it->first.compare(0, prefix.size(), prefix) == 0;
}

int pack_superpoint(b200m_handle* h, Packer& P) {
  const int D = h->cfg.descriptor_dim;
  std::vector<double> w, b;
  const std::string sp = "superpoint.";
  // ---- SuperPoint (unet_parts.py:10-48; superpoint_test.py:70-84)
  if (!P.folded(sp + "inc.conv.conv.0", sp + "inc.conv.conv.1", 64, 9, w, b, {64, 1, 3, 3}))
    return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
  h->conv1_w = P.alloc(9 * 64);
  h->conv1_b = P.alloc(64);
  for (int t = 0; t < 9; ++t)
    for (int o = 0; o < 64; ++o) P.host[h->conv1_w + t * 64 + o] = (float)w[(size_t)o * 9 + t];
  for (int o = 0; o < 64; ++o) P.host[h->conv1_b + o] = (float)b[o];
  h->stem_host.assign(P.host.begin() + h->conv1_w, P.host.begin() + h->conv1_w + 9 * 64);
  h->stem_host.insert(h->stem_host.end(), P.host.begin() + h->conv1_b, P.host.begin() + h->conv1_b + 64);
  auto conv3 = [&](const std::string& conv, const std::string& bn, int cin, int cout, ConvLayer& L) -> bool {
    if (!P.folded(sp + conv, sp + bn, cout, cin * 9, w, b, {cout, cin, 3, 3})) return false;
    L = P.pack_conv(w, b, cout, cin, 3);
    P.pack_tc(L, w);
    return true;
  };
  bool ok = conv3("inc.conv.conv.3", "inc.conv.conv.4", 64, 64, h->c1b) &&
            conv3("down1.mpconv.1.conv.0", "down1.mpconv.1.conv.1", 64, 64, h->c2a) &&
            conv3("down1.mpconv.1.conv.3", "down1.mpconv.1.conv.4", 64, 64, h->c2b) &&
            conv3("down2.mpconv.1.conv.0", "down2.mpconv.1.conv.1", 64, 128, h->c3a) &&
            conv3("down2.mpconv.1.conv.3", "down2.mpconv.1.conv.4", 128, 128, h->c3b) &&
            conv3("down3.mpconv.1.conv.0", "down3.mpconv.1.conv.1", 128, 128, h->c4a) &&
            conv3("down3.mpconv.1.conv.3", "down3.mpconv.1.conv.4", 128, 128, h->c4b);
  if (!ok) return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
  {  // detector + descriptor 3x3 heads share their input x4 -> one conv with 512 output channels
    std::vector<double> wa, ba, wd, bd;
    if (!P.folded(sp + "convPa", sp + "bnPa", 256, 128 * 9, wa, ba, {256, 128, 3, 3}) ||
        !P.folded(sp + "convDa", sp + "bnDa", 256, 128 * 9, wd, bd, {256, 128, 3, 3}))
      return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
    wa.insert(wa.end(), wd.begin(), wd.end());
    ba.insert(ba.end(), bd.begin(), bd.end());
    h->heads = P.pack_conv(wa, ba, 512, 128, 3);
    P.pack_tc(h->heads, wa);
  }
  if (!P.folded(sp + "convPb", sp + "bnPb", 65, 256, w, b, {65, 256, 1, 1}))
    return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
  h->pb = P.pack_conv(w, b, 65, 256, 1);
  P.pack_tc(h->pb, w);
  if (!P.folded(sp + "convDb", sp + "bnDb", D, 256, w, b, {D, 256, 1, 1}))
    return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
  h->db = P.pack_conv(w, b, D, 256, 1);
  P.pack_tc(h->db, w);
  return B200M_OK;
}

int pack_superglue(b200m_handle* h, Packer& P) {
  const int D = h->cfg.descriptor_dim;
  std::vector<double> w, b;
  // ---- SuperGlue (superglue_test.py:204-219)
  const std::string sg = "superglue.";
  {
    const HostTensor* bs = P.get(sg + "bin_score", {});
    if (!bs) return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
    h->bin_score = bs->data[0];
  }
  h->kenc.clear();
  {
    std::vector<int> ch = {3};
    for (int i = 0; i < h->cfg.n_kenc; ++i) ch.push_back(h->cfg.kenc[i]);
    ch.push_back(D);
    int idx = 0;
    for (size_t i = 1; i < ch.size(); ++i) {
      bool last = (i + 1 == ch.size());
      std::string conv = sg + "kenc.encoder." + std::to_string(idx);
      std::string bn = last ? std::string() : sg + "kenc.encoder." + std::to_string(idx + 1);
      if (!P.folded(conv, bn, ch[i], ch[i - 1], w, b, {ch[i], ch[i - 1]}))
        return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
      h->kenc.push_back(P.pack_linear(w, b, ch[i], ch[i - 1], round_up(ch[i - 1], 4)));
      idx += last ? 1 : 3;
    }
  }
  const int d = D / kHeads;
  std::vector<int> perm(D);   // head-major position -> reference channel (c = dd*heads + h)
  for (int hh = 0; hh < kHeads; ++hh)
    for (int dd = 0; dd < d; ++dd) perm[hh * d + dd] = dd * kHeads + hh;
  h->gnn.clear();
  for (int l = 0; l < h->cfg.n_gnn_layers; ++l) {
    std::string p = sg + "gnn.layers." + std::to_string(l);
    b200m_handle::Gnn G;
    std::vector<double> wq, bq, wk, bk, wv, bv;
    if (!P.folded(p + ".attn.proj.0", "", D, D, wq, bq, {D, D}) ||
        !P.folded(p + ".attn.proj.1", "", D, D, wk, bk, {D, D}) ||
        !P.folded(p + ".attn.proj.2", "", D, D, wv, bv, {D, D}))
      return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
    {  // fused q|k|v projection with de-interleaved heads
      std::vector<double> wcat, bcat;
      for (auto* src : {&wq, &wk, &wv})
        for (int r = 0; r < D; ++r)
          wcat.insert(wcat.end(), src->begin() + (size_t)perm[r] * D, src->begin() + (size_t)(perm[r] + 1) * D);
      for (auto* src : {&bq, &bk, &bv})
        for (int r = 0; r < D; ++r) bcat.push_back((*src)[perm[r]]);
      G.qkv = P.pack_linear(wcat, bcat, 3 * D, D, D);
    }
    if (!P.folded(p + ".attn.merge", "", D, D, w, b, {D, D})) return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
    G.merge = P.pack_linear(w, b, D, D, D, nullptr, perm.data());
    if (!P.folded(p + ".mlp.0", p + ".mlp.1", 2 * D, 2 * D, w, b, {2 * D, 2 * D}))
      return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
    G.mlp1 = P.pack_linear(w, b, 2 * D, 2 * D, 2 * D);
    if (!P.folded(p + ".mlp.3", "", D, 2 * D, w, b, {D, 2 * D})) return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
    G.mlp2 = P.pack_linear(w, b, D, 2 * D, 2 * D);
    h->gnn.push_back(G);
  }
  if (D == 128)   // weight streams of the fused layer kernel: layer l carries layer l+1's q|k|v projection
    for (int l = 0; l < h->cfg.n_gnn_layers; ++l) {
      b200m_handle::Gnn& G = h->gnn[l];
      const b200m_handle::Gnn* next = l + 1 < h->cfg.n_gnn_layers ? &h->gnn[l + 1] : nullptr;
      G.fused_w_off = P.alloc(gnn_fused_weight_floats(next != nullptr));
      G.fused_b_off = P.alloc(768);
      gnn_fused_pack_weights(P.host.data() + G.merge.w_off, P.host.data() + G.mlp1.w_off, P.host.data() + G.mlp2.w_off,
                             next ? P.host.data() + next->qkv.w_off : nullptr, P.host.data() + G.fused_w_off);
      // bias block [256 mlp1' | 128 mlp2 | 384 next q|k|v]; the merge bias only ever enters through the mlp's first
      // layer, so it is folded into that bias: b_1' = b_1 + W_1[:, D:2D] b_merge
      float* bb = P.host.data() + G.fused_b_off;
      const float* W1 = P.host.data() + G.mlp1.w_off;
      const float* bm = P.host.data() + G.merge.b_off;
      for (int o = 0; o < 256; ++o) {
        double acc = P.host[G.mlp1.b_off + o];
        for (int k = 0; k < 128; ++k) acc += (double)W1[(size_t)o * 256 + 128 + k] * (double)bm[k];
        bb[o] = (float)acc;
      }
      std::copy_n(P.host.data() + G.mlp2.b_off, 128, bb + 256);
      if (next) std::copy_n(P.host.data() + next->qkv.b_off, 384, bb + 384);
      G.fused = true;
      G.fused_has_qkv = next != nullptr;
    }
  if (!P.folded(sg + "final_proj", "", D, D, w, b, {D, D})) return fail(B200M_ERR_WEIGHTS, "%s", P.err.c_str());
  h->final_proj = P.pack_linear(w, b, D, D, D);
  return B200M_OK;
}

// A handle may carry only one half (SuperPoint used alone by superpoint_flann_test.py:52-61 /
// datasets/GlueSparse.py:18-39, or SuperGlue fed with external features): pack what is present.
int do_pack(b200m_handle* h, cudaStream_t stream) {
  Packer P{h, {}, {}};
  h->packed_sp = h->packed_sg = false;
  const bool want_sp = has_prefix(h, "superpoint."), want_sg = has_prefix(h, "superglue.");
  if (!want_sp && !want_sg) return fail(B200M_ERR_WEIGHTS, "no tensors were set before b200m_pack");
  if (want_sp) { int rc = pack_superpoint(h, P); if (rc) return rc; }
  if (want_sg) { int rc = pack_superglue(h, P); if (rc) return rc; }

  drop_graphs(h);                 // captured launches carry pointers into the old weight arena
  if (h->d_w) { cudaFree(h->d_w); h->d_w = nullptr; }
  if (cudaMalloc(&h->d_w, P.host.size() * sizeof(float)) != cudaSuccess)
    return fail(B200M_ERR_CUDA, "cudaMalloc of %zu weight bytes failed", P.host.size() * sizeof(float));
  cudaError_t e = cudaMemcpyAsync(h->d_w, P.host.data(), P.host.size() * sizeof(float), cudaMemcpyHostToDevice, stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
  if (e != cudaSuccess) return fail(B200M_ERR_CUDA, "weight upload failed: %s", cudaGetErrorString(e));
  h->packed_sp = want_sp;
  h->packed_sg = want_sg;
  return B200M_OK;
}

// ------------------------------------------------------------------ SuperPoint pipeline
struct SpDims {
  int H, W, H2, W2, H3, W3, hc, wc, H8, W8, cand_cap, dpad;
};
SpDims sp_dims(const b200m_handle* h, int H, int W) {
  SpDims d;
  d.H = H; d.W = W; d.H2 = H / 2; d.W2 = W / 2; d.H3 = d.H2 / 2; d.W3 = d.W2 / 2;
  d.hc = d.H3 / 2; d.wc = d.W3 / 2; d.H8 = d.hc * 8; d.W8 = d.wc * 8;
  int r = h->cfg.nms_radius + 1;
  d.cand_cap = next_pow2(std::max(1024, cdiv(d.H8, r) * cdiv(d.W8, r)));
  d.dpad = round_up(h->cfg.descriptor_dim, 64);
  return d;
}

struct SpWs {
  float *p0, *p1, *p0_lo, *p1_lo, *semi, *draw, *sumsq, *heat, *imgf;
  unsigned long long* keys;
  unsigned char* nms_scratch;
  int *cand_counts, *overflow;      // overflow[0]: candidate list overflow, overflow[1]: fp16 activation overflow
  size_t p0_img, p1_img, semi_img, draw_img, heat_img;
  int desc_ncb;                     // column blocks of the descriptor head = partial sums of squares per pixel
};
// images per SuperPoint micro-batch: kSpMicroBatch at 640x480, fewer for larger images so the arena stays ~8 GB
int sp_micro_batch(const b200m_handle* h, int H, int W) {
  if (h->sp_micro_batch > 0) return h->sp_micro_batch;
  const long long px = (long long)H * W;
  const long long mb = (long long)kSpMicroBatch * 640 * 480 / std::max(px, 1LL);
  return (int)std::min<long long>(kSpMicroBatch, std::max<long long>(8, mb));
}

bool sp_carve(const b200m_handle* h, const SpDims& d, int mb, Arena& A, SpWs& w) {
  w.p0_img = (size_t)64 * d.H * d.W;
  w.p1_img = std::max((size_t)64 * d.H2 * d.W2, (size_t)128 * d.H3 * d.W3);
  w.semi_img = (size_t)128 * d.hc * d.wc;
  w.draw_img = (size_t)d.dpad * d.hc * d.wc;
  w.desc_ncb = h->use_tc ? std::max(1, d.dpad / (d.dpad >= 128 ? 128 : 64)) : 1;
  w.heat_img = (size_t)d.H8 * d.W8;
  w.p0 = A.take<float>(w.p0_img * mb);
  w.p1 = A.take<float>(w.p1_img * mb);
  w.p0_lo = h->use_tc ? A.take<float>(w.p0_img * mb) : nullptr;
  w.p1_lo = h->use_tc ? A.take<float>(w.p1_img * mb) : nullptr;
  w.semi = A.take<float>(w.semi_img * mb);
  w.draw = A.take<float>(w.draw_img * mb);
  w.sumsq = A.take<float>((size_t)w.desc_ncb * d.hc * d.wc * mb);
  w.heat = A.take<float>(w.heat_img * mb);
  w.imgf = A.take<float>((size_t)d.H * d.W * mb);       // fp32 copy of a uint8 micro-batch (b200m_*_u8 entry points)
  w.keys = A.take<unsigned long long>((size_t)d.cand_cap * mb);
  w.nms_scratch = A.take<unsigned char>(nms_scratch_bytes(mb, d.H8, d.W8));
  w.cand_counts = A.take<int>(mb + 2);
  w.overflow = w.cand_counts ? w.cand_counts + mb : nullptr;
  return A.ok;
}

void run_conv(b200m_handle* h, LaunchCtx& ctx, const ConvLayer& L, const float* in, int in_c4_total, int in_c4_off,
              float* out, int out_c4_total, int n, int H, int W, bool relu, bool pool) {
  ConvParams p;
  p.in = in; p.in_c4_total = in_c4_total; p.in_c4_off = in_c4_off; p.cin = L.cin;
  p.wpk = h->d_w + L.w_off; p.bias = h->d_w + L.b_off;
  p.out = out; p.out_c4_total = out_c4_total; p.out_c4_off = 0; p.cout_pad = L.cout_pad;
  p.n = n; p.H = H; p.W = W; p.relu = relu ? 1 : 0;
  launch_conv(ctx, p, L.ks, pool);
}

// One 3x3 / 1x1 layer on the tensor cores.  in/out are (hi, lo) plane pairs; out_lo == nullptr -> full-precision output
// in out_hi.  A declined launch (no cuTensorMapEncodeTiled entry point, tensor-map or attribute failure) is an ERROR
// surfaced through finish(): the output buffer would otherwise be consumed unwritten.
void run_conv_tc(b200m_handle* h, LaunchCtx& ctx, const ConvLayer& L, const float* in_hi, const float* in_lo,
                 float* out_hi, float* out_lo, int out_c4_total, int n, int H, int W, bool pool, int* overflow,
                 bool relu = true, int in_c8_total = 0, int in_c8_off = 0, float* sumsq = nullptr,
                 int single_from_cb = 1 << 30) {
  TcConvParams p;
  p.sumsq = sumsq;
  p.single_from_cb = single_from_cb;
  p.in_hi = in_hi; p.in_lo = in_lo; p.wpk = h->d_w + L.tc_w_off; p.bias = h->d_w + L.b_off;
  p.out_hi = out_hi; p.out_lo = out_lo; p.out_c4_total = out_c4_total; p.out_c4_off = 0; p.overflow = overflow;
  p.cin = L.cin; p.cout_pad = L.cout_pad; p.nb = L.nb; p.n = n; p.H = H; p.W = W; p.relu = relu ? 1 : 0;
  p.pool = pool ? 1 : 0; p.ks = L.ks; p.in_c8_total = in_c8_total; p.in_c8_off = in_c8_off;
  if (L.cout_pad <= 128 && (int)L.bias_host.size() == L.cout_pad) {
    p.bias_in_params = 1;
    memcpy(p.bias_c, L.bias_host.data(), sizeof(float) * L.cout_pad);
  }
  if (!launch_tc_conv(ctx, p, h->num_sms) && ctx.err == cudaSuccess) {
    ctx.err = cudaErrorLaunchFailure;
    *ctx.err_where = "tc_conv (launch declined; B200M_CONV_IMPL=simt selects the fp32 CUDA-core path)";
  }
}

// encoder + heads for `n` images (n <= micro-batch): fills w.semi (C4, 32 groups) and w.draw (C4, dpad/4 groups)
// images: fp32 pixels, or (images_u8 != null) the raw 8-bit pixels -- normalised inside the fused stem's patch load; only
// the unfused / CUDA-core paths still need the fp32 copy (w.imgf)
void sp_dense(b200m_handle* h, LaunchCtx& ctx, const SpDims& d, const SpWs& w, const float* images, int n,
              const uint8_t* images_u8 = nullptr) {
  auto fp32_images = [&]() -> const float* {
    if (!images_u8) return images;
    launch_u8_to_unit_f32(ctx, images_u8, w.imgf, (size_t)n * d.H * d.W);
    return w.imgf;
  };
  // the C4 buffers are addressed per image with each layer's own channel-group count, so the ping-pong
  // buffers are simply re-interpreted per layer
  if (h->use_tc) {
    int* ovf = w.overflow ? w.overflow + 1 : nullptr;
    // fp16 hi/lo planes, C8-planar: channel units per image = C / 8
    // first conv (1 -> 64) fused into the second one's operand producer: no full-resolution 64-channel map in HBM
    bool stem = h->use_fused_stem;
    if (stem) {
      TcConvParams p;
      const ConvLayer& L = h->c1b;
      p.in_hi = nullptr; p.in_lo = nullptr; p.wpk = h->d_w + L.tc_w_off; p.bias = h->d_w + L.b_off;
      p.out_hi = w.p1; p.out_lo = w.p1_lo; p.out_c4_total = 8; p.out_c4_off = 0; p.overflow = ovf;
      p.cin = L.cin; p.cout_pad = L.cout_pad; p.nb = L.nb; p.n = n; p.H = d.H; p.W = d.W; p.relu = 1; p.pool = 1; p.ks = 3;
      p.img = images_u8 ? nullptr : images;
      p.img_u8 = images_u8;
      memcpy(p.c1, h->stem_host.data(), sizeof(p.c1));
      p.bias_in_params = 1;
      memcpy(p.bias_c, L.bias_host.data(), sizeof(float) * L.cout_pad);
      stem = launch_tc_conv(ctx, p, h->num_sms);
    }
    if (!stem) {
      images = fp32_images();
      launch_conv1_direct(ctx, images, h->d_w + h->conv1_w, h->d_w + h->conv1_b, w.p0, w.p0_lo, n, d.H, d.W);
      run_conv_tc(h, ctx, h->c1b, w.p0, w.p0_lo, w.p1, w.p1_lo, 8, n, d.H, d.W, true, ovf);      // -> 64 x H2 x W2
    }
    run_conv_tc(h, ctx, h->c2a, w.p1, w.p1_lo, w.p0, w.p0_lo, 8, n, d.H2, d.W2, false, ovf);
    run_conv_tc(h, ctx, h->c2b, w.p0, w.p0_lo, w.p1, w.p1_lo, 8, n, d.H2, d.W2, true, ovf);      // -> 64 x H3 x W3
    run_conv_tc(h, ctx, h->c3a, w.p1, w.p1_lo, w.p0, w.p0_lo, 16, n, d.H3, d.W3, false, ovf);    // 128 ch
    run_conv_tc(h, ctx, h->c3b, w.p0, w.p0_lo, w.p1, w.p1_lo, 16, n, d.H3, d.W3, true, ovf);     // -> 128 x hc x wc
    run_conv_tc(h, ctx, h->c4a, w.p1, w.p1_lo, w.p0, w.p0_lo, 16, n, d.hc, d.wc, false, ovf);
    run_conv_tc(h, ctx, h->c4b, w.p0, w.p0_lo, w.p1, w.p1_lo, 16, n, d.hc, d.wc, false, ovf);    // x4
    // cPa | cDa (512 ch = four 128-channel column blocks; the experiment's single-product part is cDa = blocks 2, 3)
    run_conv_tc(h, ctx, h->heads, w.p1, w.p1_lo, w.p0, w.p0_lo, 64, n, d.hc, d.wc, false, ovf, true, 0, 0, nullptr,
                h->single_desc ? 2 : 1 << 30);
    // 1x1 heads on the same pipeline (one tap per K block), full fp32 C4-planar outputs
    run_conv_tc(h, ctx, h->pb, w.p0, w.p0_lo, w.semi, nullptr, 32, n, d.hc, d.wc, false, ovf, false, 64, 0);
    // descriptor head: raw descriptors + per-pixel sums of squares (the L2 normalisation is applied by the sampler)
    run_conv_tc(h, ctx, h->db, w.p0, w.p0_lo, w.draw, nullptr, d.dpad / 4, n, d.hc, d.wc, false, ovf, false, 64, 32,
                w.sumsq, h->single_desc ? 0 : 1 << 30);
    return;
  } else {
    images = fp32_images();
    launch_conv1_direct(ctx, images, h->d_w + h->conv1_w, h->d_w + h->conv1_b, w.p0, nullptr, n, d.H, d.W);
    run_conv(h, ctx, h->c1b, w.p0, 16, 0, w.p1, 16, n, d.H, d.W, true, true);       // -> 64 x H2 x W2
    run_conv(h, ctx, h->c2a, w.p1, 16, 0, w.p0, 16, n, d.H2, d.W2, true, false);
    run_conv(h, ctx, h->c2b, w.p0, 16, 0, w.p1, 16, n, d.H2, d.W2, true, true);     // -> 64 x H3 x W3
    run_conv(h, ctx, h->c3a, w.p1, 16, 0, w.p0, 32, n, d.H3, d.W3, true, false);    // 128 ch
    run_conv(h, ctx, h->c3b, w.p0, 32, 0, w.p1, 32, n, d.H3, d.W3, true, true);     // -> 128 x hc x wc
    run_conv(h, ctx, h->c4a, w.p1, 32, 0, w.p0, 32, n, d.hc, d.wc, true, false);
    run_conv(h, ctx, h->c4b, w.p0, 32, 0, w.p1, 32, n, d.hc, d.wc, true, false);    // x4
    run_conv(h, ctx, h->heads, w.p1, 32, 0, w.p0, 128, n, d.hc, d.wc, true, false); // cPa | cDa (512 ch)
  }
  run_conv(h, ctx, h->pb, w.p0, 128, 0, w.semi, 32, n, d.hc, d.wc, false, false); // semi (65 of 128 ch)
  run_conv(h, ctx, h->db, w.p0, 128, 64, w.draw, d.dpad / 4, n, d.hc, d.wc, false, false);
  launch_c4_sumsq(ctx, w.draw, d.dpad / 4, 0, w.sumsq, h->cfg.descriptor_dim, n, d.hc, d.wc);
}

int sp_forward_impl(b200m_handle* h, void* stream, const void* images_any, bool images_u8, int n_images, int H, int W,
                    float* keypoints, float* scores, float* descriptors, int* counts, int cap,
                    float* semi_out, float* desc_out, float* tok_out, int tok_ld, size_t tok_img_stride,
                    void* ws, size_t ws_bytes) {
  if (!h->packed_sp) return fail(B200M_ERR_WEIGHTS, "SuperPoint weights are not packed (b200m_set_tensor + b200m_pack)");
  if (n_images <= 0) return B200M_OK;
  if (H < 8 || W < 8) return fail(B200M_ERR_INVALID, "image smaller than 8x8");
  const int D = h->cfg.descriptor_dim;
  SpDims d = sp_dims(h, H, W);
  const int mb = std::min(n_images, sp_micro_batch(h, H, W));
  Arena A(ws, ws_bytes);
  SpWs w;
  if (!sp_carve(h, d, mb, A, w)) return fail(B200M_ERR_WORKSPACE, "SuperPoint workspace too small: need %zu bytes", A.off);
  DeviceGuard dev_guard__(h->device);
  h->pdl_now = n_images <= h->pdl_max_pairs;
  LaunchCtx ctx = make_ctx(h, stream);
  cudaMemsetAsync(w.overflow, 0, 2 * sizeof(int), ctx.stream);
  for (int i0 = 0; i0 < n_images; i0 += mb) {
    const int n = std::min(mb, n_images - i0);
    // 8-bit input: SSHIDataset.py:26-29's normalisation (pixel / 255, rounded to fp32) happens in the stem's patch load
    if (images_u8) sp_dense(h, ctx, d, w, nullptr, n, static_cast<const uint8_t*>(images_any) + (size_t)i0 * H * W);
    else sp_dense(h, ctx, d, w, static_cast<const float*>(images_any) + (size_t)i0 * H * W, n);
    if (semi_out)
      launch_c4_to_nchw(ctx, w.semi, 32, 0, 65, semi_out + (size_t)i0 * 65 * d.hc * d.wc, n, d.hc, d.wc, false);
    if (desc_out)
      launch_c4_to_nchw(ctx, w.draw, d.dpad / 4, 0, D, desc_out + (size_t)i0 * D * d.hc * d.wc, n, d.hc, d.wc, true);
    if (keypoints) {
      launch_softmax_heat(ctx, w.semi, 32, w.heat, n, d.hc, d.wc);
      cudaMemsetAsync(w.cand_counts, 0, sizeof(int) * mb, ctx.stream);
      launch_nms_candidates(ctx, w.heat, nullptr, n, d.H8, d.W8, h->cfg.nms_radius, h->cfg.keypoint_threshold,
                            h->cfg.remove_borders, w.keys, w.cand_counts, d.cand_cap, w.overflow, w.nms_scratch);
      launch_select_keypoints(ctx, w.keys, w.cand_counts, d.cand_cap, n, d.W8, h->cfg.max_keypoints,
                              keypoints + (size_t)i0 * cap * 2, scores + (size_t)i0 * cap, counts + i0, cap);
      launch_sample_descriptors(ctx, w.draw, d.dpad / 4, D, n, d.hc, d.wc, keypoints + (size_t)i0 * cap * 2, counts + i0,
                                cap, h->cfg.align_corners, descriptors ? descriptors + (size_t)i0 * D * cap : nullptr,
                                tok_out ? tok_out + (size_t)i0 * tok_img_stride : nullptr, tok_ld, tok_img_stride,
                                w.sumsq, w.desc_ncb);
    }
  }
  if (keypoints) launch_apply_flags(ctx, w.overflow, counts, n_images);
  return finish(h, ctx);
}

size_t sp_ws_bytes(const b200m_handle* h, int n_images, int H, int W) {
  SpDims d = sp_dims(h, H, W);
  Arena A(nullptr, 0);
  SpWs w;
  sp_carve(h, d, std::max(1, std::min(n_images, sp_micro_batch(h, H, W))), A, w);
  return A.off + 256;
}

// ------------------------------------------------------------------ SuperGlue pipeline
struct SgWs {
  int Np, ldS, ld_uv;
  size_t rows;           // 2 * B * Np
  float *X, *QKV, *QKV_lo, *VT, *VT_lo, *MSG, *HID, *IN4, *S, *u, *v, *max0, *ot_part;
  float* XP;             // [x ; merged message] as fp16 operand planes (hi [rows][2D] halves, then lo): the A operand of
                         // the layer GEMMs when the layers run as tc_gemm launches (D != 128); null on the fused path
  int *idx0, *idx1;
};
// pairs per Sinkhorn micro-batch: their score matrices share the L2
int ot_chunk_pairs(int B, int N, int ldS) {
  const size_t pair_bytes = (size_t)std::max(N, 1) * ldS * sizeof(float);
  return std::max(1, std::min(B, (int)((size_t)(64u << 20) / std::max<size_t>(pair_bytes, 1))));
}
size_t ot_part_floats(int B, int N, int M, int ldS) {
  return ot_fused_scratch_floats(B, N, M);   // the fused path sweeps every pair in one launch
}

bool sg_carve(const b200m_handle* h, int B, int N, int M, Arena& A, SgWs& w) {
  const int D = h->cfg.descriptor_dim;
  w.Np = round_up(std::max(std::max(N, M), 1), 64);
  w.rows = (size_t)2 * B * w.Np;
  w.ldS = round_up(std::max(M, 1), 4);
  w.ld_uv = round_up(std::max(N, M) + 1, 4);
  w.X = A.take<float>(w.rows * 2 * D);
  w.QKV = A.take<float>(w.rows * 3 * D);
  w.QKV_lo = h->use_tc_attn ? A.take<float>(w.rows * 3 * D) : nullptr;
  w.VT = h->use_tc_attn ? A.take<float>(w.rows * D) : nullptr;
  w.VT_lo = h->use_tc_attn ? A.take<float>(w.rows * D) : nullptr;
  w.MSG = A.take<float>(w.rows * D);
  w.HID = A.take<float>(w.rows * 2 * D);
  // (widths whose head size the tensor-core attention kernel covers: 16, 32, 64)
  w.XP = (h->use_tc_attn && h->use_tc_gemm && !(h->use_fused_gnn && D == 128) && (D == 64 || D == 128 || D == 256))
             ? A.take<float>(w.rows * 2 * D) : nullptr;
  w.IN4 = A.take<float>(w.rows * 4);
  w.S = A.take<float>((size_t)B * std::max(N, 1) * w.ldS);
  w.u = A.take<float>((size_t)B * w.ld_uv);
  w.v = A.take<float>((size_t)B * w.ld_uv);
  w.max0 = A.take<float>((size_t)B * w.ld_uv);
  w.ot_part = A.take<float>(ot_part_floats(B, N, M, w.ldS));
  w.idx0 = A.take<int>((size_t)B * w.ld_uv);
  w.idx1 = A.take<int>((size_t)B * w.ld_uv);
  return A.ok;
}

// operand-plane input (A_hi / A_lo, lda_p halves) and the additional plane copy of an fp32 result (P_hi / P_lo, ldp)
struct PlaneIO {
  const void* A_hi = nullptr; const void* A_lo = nullptr; int lda_p = 0;
  void* P_hi = nullptr; void* P_lo = nullptr; int ldp = 0;
};

void run_linear(b200m_handle* h, LaunchCtx& ctx, const Linear& L, const float* A, int lda, float* C, int ldc,
                size_t M, bool relu, bool accumulate, float* C_lo = nullptr, float* VT = nullptr,
                float* VT_lo = nullptr, int vt_col0 = 0, int vt_np = 1, float lo_scale = 1.f,   // C_lo set => fp16 plane outputs
                const PlaneIO* pio = nullptr) {
  GemmParams p;
  if (pio) {
    p.A_hi_p = pio->A_hi; p.A_lo_p = pio->A_lo; p.lda_p = pio->lda_p;
    p.P_hi = pio->P_hi; p.P_lo = pio->P_lo; p.ldp = pio->ldp;
  }
  p.A = A; p.lda = lda; p.strideA = 0;
  p.Bw = h->d_w + L.w_off; p.ldb = L.K; p.strideB = 0;
  p.C = C; p.ldc = ldc; p.strideC = 0;
  p.bias = h->d_w + L.b_off;
  p.M = (int)M; p.N = L.N; p.K = L.K; p.batch = 1;
  p.alpha = 1.f; p.relu = relu ? 1 : 0; p.accumulate = accumulate ? 1 : 0;
  p.C_lo = C_lo; p.out_f16 = C_lo ? 1 : 0;
  p.VT = VT; p.VT_lo = VT_lo; p.vt_col0 = vt_col0; p.vt_np = vt_np; p.lo_scale = lo_scale;
  p.single = h->single_gemm ? 1 : 0;
  if (h->use_tc_gemm && launch_tc_gemm(ctx, p, h->d_w + L.w_hi_off, h->d_w + L.w_lo_off, h->num_sms)) return;
  if (pio) {                 // plane operands / outputs exist only on the tensor-core path: no CUDA-core fallback
    if (ctx.err == cudaSuccess) {
      ctx.err = cudaErrorLaunchFailure;
      *ctx.err_where = "tc_gemm with operand planes (launch declined)";
    }
    return;
  }
  launch_gemm(ctx, p);
}

// keypoint encoder on one side: X[:, :D] += MLP([x_norm, y_norm, score])   (X already holds the descriptors)
void sg_kenc(b200m_handle* h, LaunchCtx& ctx, const SgWs& w, int side, const float* kpts, const float* scores,
             int B, int N, int Himg, int Wimg) {
  const int D = h->cfg.descriptor_dim;
  const size_t rows = (size_t)B * w.Np;
  float* in4 = w.IN4 + (size_t)side * rows * 4;
  // normalize_keypoints (:63-70): center = size/2, scaling = max(W,H) * 0.7
  const float cx = (float)Wimg / 2.f, cy = (float)Himg / 2.f;
  const float scale = (float)std::max(Wimg, Himg) * 0.7f;
  launch_kenc_input(ctx, kpts, scores, B, N, w.Np, cx, cy, scale, in4);
  // ping-pong through the (currently unused) QKV / HID buffers of this side
  float* t0 = w.QKV + (size_t)side * rows * 3 * D;
  float* t1 = w.HID + (size_t)side * rows * 2 * D;
  const float* cur = in4;
  int ld = 4;
  for (size_t i = 0; i < h->kenc.size(); ++i) {
    const Linear& L = h->kenc[i];
    bool last = (i + 1 == h->kenc.size());
    if (last) {
      run_linear(h, ctx, L, cur, ld, w.X + (size_t)side * rows * 2 * D, 2 * D, rows, false, true);
    } else {
      float* dst = (i & 1) ? t1 : t0;
      run_linear(h, ctx, L, cur, ld, dst, L.N, rows, true, false);
      cur = dst;
      ld = L.N;
    }
  }
}

void sg_gnn(b200m_handle* h, LaunchCtx& ctx, const SgWs& w, int B, const int* c0, const int* c1, int N, int M,
            int l_begin, int l_end) {
  const int D = h->cfg.descriptor_dim;
  if (h->use_fused_gnn && h->use_tc_attn && h->use_tc_gemm && D == 128 && w.Np % 8 == 0 && l_begin < l_end &&
      h->gnn[l_begin].fused) {
    // attention planes -> fused merge/mlp/residual/next-q|k|v kernel: two launches per layer, the message, the hidden
    // layer and the concatenation never reach HBM.  MSG doubles as the attention's fp16 hi / lo output planes.
    __half* att_hi = reinterpret_cast<__half*>(w.MSG);
    __half* att_lo = att_hi + w.rows * D;
    run_linear(h, ctx, h->gnn[l_begin].qkv, w.X, 2 * D, w.QKV, 3 * D, w.rows, false, false, w.QKV_lo, w.VT, w.VT_lo,
               2 * D, w.Np);
    for (int l = l_begin; l < l_end; ++l) {
      const b200m_handle::Gnn& G = h->gnn[l];
      const bool cross = h->cfg.gnn_cross[l] != 0;
      bool ok = launch_tc_attention(ctx, w.QKV, w.QKV_lo, w.VT, w.VT_lo, nullptr, B, w.Np, D, kHeads, c0, c1, N, M,
                                    cross, att_hi, att_lo, h->single_attn);
      GnnFusedParams p;
      p.wts = reinterpret_cast<const uint8_t*>(h->d_w + G.fused_w_off);
      p.bias = h->d_w + G.fused_b_off;
      p.X = w.X; p.ldx = 2 * D;
      p.qkv_hi = reinterpret_cast<__half*>(w.QKV); p.qkv_lo = reinterpret_cast<__half*>(w.QKV_lo);
      p.vt_hi = reinterpret_cast<__half*>(w.VT); p.vt_lo = reinterpret_cast<__half*>(w.VT_lo); p.vt_np = w.Np;
      p.rows = (int)w.rows;
      p.nt4 = (l + 1 < l_end && G.fused_has_qkv) ? 3 : 0;
      p.overflow = nullptr;
      p.single = h->single_gnn ? 1 : 0;
      ok = ok && launch_tc_gnn_layer(ctx, p, att_hi, att_lo, h->num_sms);
      if (!ok && ctx.err == cudaSuccess) {
        ctx.err = cudaErrorLaunchFailure;
        *ctx.err_where = "fused GNN layer (launch declined)";
      }
    }
    return;
  }
  if (w.XP && h->use_tc_attn && h->use_tc_gemm && l_begin < l_end) {
    // Layers as tc_gemm launches (D != 128: the fused kernel's operands do not fit one SM) with every A
    // operand in fp16 hi / lo*2048 PLANE format, written by the producing kernel's epilogue: the GEMMs skip their
    // in-kernel split (at D = 256 the kernel is bound by shared-memory bandwidth, and the split moves 64 of the 208 KB
    // a pipeline stage moves).  The planes hold exactly the values the in-kernel split derives from the fp32 data, so
    // the results are bit-identical to the fp32-input path.
    //   XP = [x ; merged message] planes, MSG = attention output planes, HID = hidden-layer planes, X[:, :D] fp32 residual
    __half* xp_hi = reinterpret_cast<__half*>(w.XP);
    __half* xp_lo = xp_hi + w.rows * 2 * D;
    __half* att_hi = reinterpret_cast<__half*>(w.MSG);
    __half* att_lo = att_hi + w.rows * D;
    __half* hid_hi = reinterpret_cast<__half*>(w.HID);
    __half* hid_lo = hid_hi + w.rows * 2 * D;
    launch_split_planes(ctx, w.X, 2 * D, xp_hi, xp_lo, 2 * D, w.rows, D);
    bool ok = true;
    for (int l = l_begin; l < l_end && ok; ++l) {
      const b200m_handle::Gnn& G = h->gnn[l];
      const bool cross = h->cfg.gnn_cross[l] != 0;
      PlaneIO x_in;   x_in.A_hi = xp_hi; x_in.A_lo = xp_lo; x_in.lda_p = 2 * D;
      run_linear(h, ctx, G.qkv, w.X, 2 * D, w.QKV, 3 * D, w.rows, false, false, w.QKV_lo, w.VT, w.VT_lo, 2 * D, w.Np, 1.f,
                 &x_in);
      ok = launch_tc_attention(ctx, w.QKV, w.QKV_lo, w.VT, w.VT_lo, nullptr, B, w.Np, D, kHeads, c0, c1, N, M, cross,
                               att_hi, att_lo, h->single_attn);
      PlaneIO att_in; att_in.A_hi = att_hi; att_in.A_lo = att_lo; att_in.lda_p = D;
      run_linear(h, ctx, G.merge, nullptr, D, reinterpret_cast<float*>(xp_hi + D), 2 * D, w.rows, false, false,
                 reinterpret_cast<float*>(xp_lo + D), nullptr, nullptr, 0, 1, 2048.f, &att_in);       // -> XP[:, D:2D]
      run_linear(h, ctx, G.mlp1, nullptr, 2 * D, reinterpret_cast<float*>(hid_hi), 2 * D, w.rows, true, false,
                 reinterpret_cast<float*>(hid_lo), nullptr, nullptr, 0, 1, 2048.f, &x_in);            // relu(bn(W1 [x;msg]))
      PlaneIO hid_io; hid_io.A_hi = hid_hi; hid_io.A_lo = hid_lo; hid_io.lda_p = 2 * D;
      hid_io.P_hi = xp_hi; hid_io.P_lo = xp_lo; hid_io.ldp = 2 * D;
      run_linear(h, ctx, G.mlp2, nullptr, 2 * D, w.X, 2 * D, w.rows, false, true, nullptr, nullptr, nullptr, 0, 1, 1.f,
                 &hid_io);                                                                          // x += W2 hid (+ planes)
    }
    if (!ok && ctx.err == cudaSuccess) {
      ctx.err = cudaErrorLaunchFailure;
      *ctx.err_where = "tensor-core attention (launch declined)";
    }
    return;
  }
  for (int l = l_begin; l < l_end; ++l) {
    const b200m_handle::Gnn& G = h->gnn[l];
    const bool cross = h->cfg.gnn_cross[l] != 0;
    bool done = false;
    if (h->use_tc_attn) {
      run_linear(h, ctx, G.qkv, w.X, 2 * D, w.QKV, 3 * D, w.rows, false, false, w.QKV_lo, w.VT, w.VT_lo, 2 * D, w.Np);
      done = launch_tc_attention(ctx, w.QKV, w.QKV_lo, w.VT, w.VT_lo, w.MSG, B, w.Np, D, kHeads, c0, c1, N, M, cross,
                                 nullptr, nullptr, h->single_attn);
    }
    if (!done) {
      run_linear(h, ctx, G.qkv, w.X, 2 * D, w.QKV, 3 * D, w.rows, false, false);
      launch_attention(ctx, w.QKV, w.MSG, B, w.Np, D, kHeads, c0, c1, N, M, cross);
    }
    run_linear(h, ctx, G.merge, w.MSG, D, w.X + D, 2 * D, w.rows, false, false);     // message -> X[:, D:2D]
    run_linear(h, ctx, G.mlp1, w.X, 2 * D, w.HID, 2 * D, w.rows, true, false);       // relu(bn(W1 [x;msg]))
    run_linear(h, ctx, G.mlp2, w.HID, 2 * D, w.X, 2 * D, w.rows, false, true);       // x += W2 hid
  }
}

void sg_scores(b200m_handle* h, LaunchCtx& ctx, const SgWs& w, int B, int N, int M) {
  const int D = h->cfg.descriptor_dim;
  run_linear(h, ctx, h->final_proj, w.X, 2 * D, w.MSG, D, w.rows, false, false);
  GemmParams p;
  p.A = w.MSG; p.lda = D; p.strideA = (long long)w.Np * D;
  p.Bw = w.MSG + (size_t)B * w.Np * D; p.ldb = D; p.strideB = (long long)w.Np * D;
  p.C = w.S; p.ldc = w.ldS; p.strideC = (long long)N * w.ldS;
  p.bias = nullptr; p.M = N; p.N = M; p.K = D; p.batch = B;
  p.alpha = 1.f / sqrtf((float)D); p.relu = 0; p.accumulate = 0;
  if (h->use_tc_gemm && w.QKV_lo && D % 64 == 0) {
    // scores on the tensor cores: side 1's projected descriptors once more as fp16 hi / lo*2048 operand planes (the
    // q|k|v buffers are free now), then ONE batched fp16x3 GEMM  S_b = mdesc0_b mdesc1_b^T / sqrt(D)  over all pairs
    const size_t side = (size_t)B * w.Np;
    run_linear(h, ctx, h->final_proj, w.X + side * 2 * D, 2 * D, w.QKV, D, side, false, false, w.QKV_lo, nullptr, nullptr,
               0, 1, 2048.f);
    p.batch_rows_a = w.Np; p.batch_rows_b = w.Np;
    p.single = h->single_gemm ? 1 : 0;
    if (launch_tc_gemm(ctx, p, w.QKV, w.QKV_lo, h->num_sms)) return;
  }
  launch_gemm(ctx, p);
}

OtParams ot_params(b200m_handle* h, const SgWs& w, const float* S, int ldS, long long strideS, int B, int N, int M,
                   const int* c0, const int* c1) {
  OtParams p;
  p.S = S; p.ldS = ldS; p.strideS = strideS; p.u = w.u; p.v = w.v; p.ld_uv = w.ld_uv;
  p.counts0 = c0; p.counts1 = c1; p.B = B; p.N = N; p.M = M; p.alpha = h->bin_score;
  return p;
}

// Sinkhorn sweeps in micro-batches of pairs whose score matrices fit the L2 together
void sg_sinkhorn(b200m_handle* h, LaunchCtx& ctx, const OtParams& all, int iters, float* partials) {
  const int chunk = ot_chunk_pairs(all.B, all.N, all.ldS);
  const bool fused = partials && ot_fused_supported(all);
  launch_ot_init(ctx, all);
  if (fused) {
    // one launch per iteration over ALL pairs: the kernel is a DRAM-streaming pass (S is read once per iteration,
    // rows arrive through per-warp bulk-copy rings), so there is nothing to gain from L2-sized micro-batches
    launch_ot_sinkhorn_fused(ctx, all, iters, partials, h->num_sms);
    return;
  }
  for (int b0 = 0; b0 < all.B; b0 += chunk) {
    OtParams p = all;
    p.B = std::min(chunk, all.B - b0);
    p.S = all.S + (size_t)b0 * all.strideS;
    p.u = all.u + (size_t)b0 * all.ld_uv;
    p.v = all.v + (size_t)b0 * all.ld_uv;
    if (p.counts0) p.counts0 += b0;
    if (p.counts1) p.counts1 += b0;
    for (int it = 0; it < iters; ++it) {
      launch_ot_row_update(ctx, p);
      launch_ot_col_update(ctx, p);
    }
  }
}

int sg_core(b200m_handle* h, LaunchCtx& ctx, const SgWs& w, const float* kpts0, const float* scores0,
            const int* c0, const float* kpts1, const float* scores1, const int* c1, int B, int N, int M,
            int H0, int W0, int H1, int W1, int64_t* matches0, int64_t* matches1, float* ms0, float* ms1) {
  sg_kenc(h, ctx, w, 0, kpts0, scores0, B, N, H0, W0);
  sg_kenc(h, ctx, w, 1, kpts1, scores1, B, M, H1, W1);
  sg_gnn(h, ctx, w, B, c0, c1, N, M, 0, h->cfg.n_gnn_layers);
  sg_scores(h, ctx, w, B, N, M);
  OtParams p = ot_params(h, w, w.S, w.ldS, (long long)N * w.ldS, B, N, M, c0, c1);
  sg_sinkhorn(h, ctx, p, h->cfg.sinkhorn_iterations, w.ot_part);
  launch_ot_argmax(ctx, p, w.idx0, w.max0, w.idx1);
  launch_match_select(ctx, w.idx0, w.max0, w.idx1, w.ld_uv, c0, c1, B, N, M, h->cfg.match_threshold,
                      (long long*)matches0, (long long*)matches1, ms0, ms1);
  return 0;
}

}  // namespace

// =================================================================== exported C ABI
extern "C" {

const char* b200m_last_error(void) { return g_err.c_str(); }
int b200m_version(void) { return 100; }

int b200m_create(const b200m_config* cfg, int device, b200m_handle** out) {
  if (!cfg || !out) return fail(B200M_ERR_INVALID, "null argument");
  const int D = cfg->descriptor_dim;
  if (!(D == 64 || D == 128 || D == 256)) return fail(B200M_ERR_INVALID, "descriptor_dim must be 64, 128 or 256 (got %d)", D);
  if (cfg->nms_radius < 0 || cfg->nms_radius > 4)   // reference asserts nms_radius >= 0 (superpoint_test.py:9)
    return fail(B200M_ERR_INVALID, "nms_radius must be in [0,4] (got %d)", cfg->nms_radius);
  if (cfg->n_kenc < 1 || cfg->n_kenc > B200M_MAX_KENC) return fail(B200M_ERR_INVALID, "bad keypoint_encoder length");
  for (int i = 0; i < cfg->n_kenc; ++i)
    if (cfg->kenc[i] <= 0 || cfg->kenc[i] % 4 || cfg->kenc[i] > 2 * D)
      return fail(B200M_ERR_INVALID, "keypoint_encoder widths must be multiples of 4 and <= 2*descriptor_dim");
  if (cfg->n_gnn_layers < 0 || cfg->n_gnn_layers > B200M_MAX_GNN) return fail(B200M_ERR_INVALID, "bad GNN layer count");
  if (cfg->sinkhorn_iterations < 0) return fail(B200M_ERR_INVALID, "negative sinkhorn_iterations");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
    return fail(B200M_ERR_CUDA, "CUDA device %d not available (%d devices)", device, ndev);
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, device);
  if (prop.major != 10)
    return fail(B200M_ERR_CUDA, "libb200match is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  b200m_handle* h = new b200m_handle();
  h->cfg = *cfg;
  h->device = device;
  h->num_sms = prop.multiProcessorCount;
  const char* impl = getenv("B200M_CONV_IMPL");
  h->use_tc = !(impl && strcmp(impl, "simt") == 0);
  impl = getenv("B200M_ATTN_IMPL");
  h->use_tc_attn = !(impl && strcmp(impl, "simt") == 0);
  impl = getenv("B200M_GEMM_IMPL");
  h->use_tc_gemm = !(impl && strcmp(impl, "simt") == 0);
  impl = getenv("B200M_STEM_IMPL");
  h->use_fused_stem = !(impl && strcmp(impl, "unfused") == 0);
  impl = getenv("B200M_GNN_IMPL");
  h->use_fused_gnn = !(impl && strcmp(impl, "unfused") == 0);
  impl = getenv("B200M_SINGLE");
  if (impl && impl[0] && !kSingleExp) {
    delete h;
    return fail(B200M_ERR_INVALID, "B200M_SINGLE needs a library built with EXTRA=-DB200M_SINGLE_EXPERIMENT");
  }
  if (impl) {
    h->single_desc = strstr(impl, "desc") != nullptr;
    h->single_gemm = strstr(impl, "gemm") != nullptr;
    h->single_gnn = strstr(impl, "gnn") != nullptr;
    h->single_attn = strstr(impl, "attn") != nullptr;
  }
  impl = getenv("B200M_GRAPHS");
  h->use_graphs = !(impl && strcmp(impl, "0") == 0);
  impl = getenv("B200M_SP_MICROBATCH");
  if (impl && atoi(impl) > 0) h->sp_micro_batch = std::min(atoi(impl), 256);
  impl = getenv("B200M_PDL_MAX_PAIRS");
  if (impl) h->pdl_max_pairs = atoi(impl);
  impl = getenv("B200M_SP_DUAL_MAX");
  if (impl) h->sp_dual_max = atoi(impl);
  *out = h;
  return B200M_OK;
}

void b200m_destroy(b200m_handle* h) {
  if (!h) return;
  DeviceGuard dev_guard__(h->device);
  drop_graphs(h);
  if (h->g_in) cudaEventDestroy(h->g_in);
  if (h->g_out) cudaEventDestroy(h->g_out);
  if (h->gstream) cudaStreamDestroy(h->gstream);
  if (h->fork_ev) cudaEventDestroy(h->fork_ev);
  if (h->join_ev) cudaEventDestroy(h->join_ev);
  if (h->side_stream) cudaStreamDestroy(h->side_stream);
  if (h->d_w) cudaFree(h->d_w);
  delete h;
}

int b200m_set_tensor(b200m_handle* h, const char* name, const float* data_host, const int64_t* shape, int ndim) {
  if (!h || !name || ndim < 0 || (ndim > 0 && !shape)) return fail(B200M_ERR_INVALID, "null argument");
  HostTensor t;
  for (int i = 0; i < ndim; ++i) t.shape.push_back(shape[i]);
  size_t n = t.numel();
  if (n && !data_host) return fail(B200M_ERR_INVALID, "null data for %s", name);
  t.data.assign(data_host, data_host + n);
  h->tensors[name] = std::move(t);
  h->packed_sp = h->packed_sg = false;
  return B200M_OK;
}

int b200m_pack(b200m_handle* h, void* stream) {
  if (!h) return fail(B200M_ERR_INVALID, "null handle");
  DeviceGuard dev_guard__(h->device);
  return do_pack(h, (cudaStream_t)stream);
}

int b200m_debug_conv_layer(b200m_handle* h, int layer, int use_tc, const float* in, float* out, int n, int H,
                           int W, void* stream) {
  if (!h || !h->packed_sp || !in || !out) return fail(B200M_ERR_INVALID, "bad argument / SuperPoint not packed");
  DeviceGuard dev_guard0__(h->device);
  const ConvLayer* Ls[8] = {&h->c1b, &h->c2a, &h->c2b, &h->c3a, &h->c3b, &h->c4a, &h->c4b, &h->heads};
  const bool pools[8] = {true, false, true, false, true, false, false, false};
  if (layer < 0 || layer >= 8) return fail(B200M_ERR_INVALID, "layer must be in [0,8)");
  const ConvLayer& L = *Ls[layer];
  const bool pool = pools[layer];
  const int Ho = pool ? H / 2 : H, Wo = pool ? W / 2 : W;
  const size_t in_f = (size_t)n * L.cin * H * W, out_f = (size_t)n * L.cout_pad * Ho * Wo;
  float* buf = nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMallocAsync(&buf, (3 * in_f + 2 * out_f) * sizeof(float), st) != cudaSuccess)
    return fail(B200M_ERR_CUDA, "scratch allocation failed");
  float *a = buf, *a_hi = a + in_f, *a_lo = a_hi + in_f, *o_hi = a_lo + in_f, *o_lo = o_hi + out_f;
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  if (use_tc) {
    launch_nchw_to_c8_split(ctx, in, L.cin, a_hi, a_lo, n, H, W);
    const bool split_out = layer != 7;
    run_conv_tc(h, ctx, L, a_hi, a_lo, o_hi, split_out ? o_lo : nullptr, split_out ? L.cout_pad / 8 : L.cout_pad / 4, n,
                H, W, pool, nullptr);
    if (split_out) launch_c8_to_nchw(ctx, o_hi, o_lo, L.cout_pad / 8, L.cout, out, n, Ho, Wo);
    else launch_c4_to_nchw(ctx, o_hi, L.cout_pad / 4, 0, L.cout, out, n, Ho, Wo, false);
  } else {
    launch_nchw_to_c4(ctx, in, L.cin, a, L.cin / 4, n, H, W);
    run_conv(h, ctx, L, a, L.cin / 4, 0, o_hi, L.cout_pad / 4, n, H, W, true, pool);
    launch_c4_to_nchw(ctx, o_hi, L.cout_pad / 4, 0, L.cout, out, n, Ho, Wo, false);
  }
  cudaFreeAsync(buf, st);
  return finish(h, ctx);
}

int b200m_debug_attention(b200m_handle* h, const float* qkv, float* msg, int B, int Np, int n0, int n1, int cross,
                          int use_tc, void* stream) {
  if (!h || !qkv || !msg) return fail(B200M_ERR_INVALID, "null argument");
  const int D = h->cfg.descriptor_dim;
  const size_t rows = (size_t)2 * B * Np, n = rows * 3 * D;
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  if (use_tc) {
    float* planes = nullptr;      // fp16 planes: hi, lo (rows x 3D) and V^T hi, lo (rows x D), stored in a float buffer
    const size_t nv = rows * D;
    if (cudaMallocAsync(&planes, (n + nv) * sizeof(float), ctx.stream) != cudaSuccess)
      return fail(B200M_ERR_CUDA, "scratch allocation failed");
    float *q_hi = planes, *q_lo = planes + n / 2, *vt_hi = planes + n, *vt_lo = planes + n + nv / 2;
    launch_qkv_to_f16_planes(ctx, qkv, q_hi, q_lo, vt_hi, vt_lo, 2 * B, Np, D);
    bool ok = launch_tc_attention(ctx, q_hi, q_lo, vt_hi, vt_lo, msg, B, Np, D, kHeads, nullptr, nullptr, n0, n1,
                                  cross != 0);
    cudaFreeAsync(planes, ctx.stream);
    if (!ok) return fail(B200M_ERR_CUDA, "tcgen05 attention launch refused");
  } else {
    launch_attention(ctx, qkv, msg, B, Np, D, kHeads, nullptr, nullptr, n0, n1, cross != 0);
  }
  return finish(h, ctx);
}

long long b200m_launch_count(const b200m_handle* h) { return h ? h->launches : 0; }
long long b200m_graph_replay_count(const b200m_handle* h) { return h ? h->graph_replays : 0; }

int b200m_profile_begin(b200m_handle* h, int max_records) {
  if (!h || max_records <= 0) return fail(B200M_ERR_INVALID, "bad argument");
  h->prof_recs.assign((size_t)max_records, ProfRecord{nullptr, nullptr, nullptr});
  h->prof.recs = h->prof_recs.data();
  h->prof.cap = max_records;
  h->prof.n = 0;
  h->prof.enabled = true;
  return B200M_OK;
}

int b200m_profile_end(b200m_handle* h, char* json, size_t json_cap) {
  if (!h || !json || json_cap < 3) return fail(B200M_ERR_INVALID, "bad argument");
  h->prof.enabled = false;
  DeviceGuard dev_guard__(h->device);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return fail(B200M_ERR_CUDA, "sync failed: %s", cudaGetErrorString(e));
  std::map<std::string, std::pair<double, long long>> acc;
  for (int i = 0; i < h->prof.n; ++i) {
    ProfRecord& r = h->prof_recs[i];
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.start, r.stop) == cudaSuccess) {
      auto& a = acc[r.name];
      a.first += ms;
      a.second += 1;
    }
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
  }
  cudaGetLastError();
  std::string out = "{";
  bool first = true;
  for (auto& kv : acc) {
    char buf[256];
    snprintf(buf, sizeof(buf), "%s\"%s\": {\"ms\": %.6f, \"launches\": %lld}", first ? "" : ", ", kv.first.c_str(),
             kv.second.first, kv.second.second);
    out += buf;
    first = false;
  }
  out += "}";
  h->prof.n = 0;
  if (out.size() + 1 > json_cap) return fail(B200M_ERR_INVALID, "profile buffer too small");
  memcpy(json, out.c_str(), out.size() + 1);
  return B200M_OK;
}

int b200m_keypoint_capacity(const b200m_handle* h, int H, int W) {
  if (!h) return 0;
  if (h->cfg.max_keypoints >= 0) return h->cfg.max_keypoints;
  return sp_dims(h, H, W).cand_cap;
}

size_t b200m_superpoint_workspace_bytes(const b200m_handle* h, int n_images, int H, int W) {
  return h ? sp_ws_bytes(h, n_images, H, W) : 0;
}

int b200m_superpoint_forward(b200m_handle* h, const float* images, int n_images, int H, int W, float* keypoints,
                             float* scores, float* descriptors, int* counts, int cap, void* ws, size_t ws_bytes,
                             void* stream) {
  if (!h || !images || !keypoints || !scores || !counts) return fail(B200M_ERR_INVALID, "null argument");
  if (cap < b200m_keypoint_capacity(h, H, W)) return fail(B200M_ERR_INVALID, "keypoint capacity %d too small", cap);
  DeviceGuard dev_guard__(h->device);
  return with_graph(h, stream, {1, kp(images), ki(n_images, cap), ki(H, W), kp(keypoints), kp(scores), kp(descriptors),
                                kp(counts), kp(ws), (uint64_t)ws_bytes},
                    [&](void* st) {
                      return sp_forward_impl(h, st, images, false, n_images, H, W, keypoints, scores, descriptors, counts,
                                             cap, nullptr, nullptr, nullptr, 0, 0, ws, ws_bytes);
                    });
}

int b200m_superpoint_forward_u8(b200m_handle* h, const uint8_t* images, int n_images, int H, int W, float* keypoints,
                                float* scores, float* descriptors, int* counts, int cap, void* ws, size_t ws_bytes,
                                void* stream) {
  if (!h || !images || !keypoints || !scores || !counts) return fail(B200M_ERR_INVALID, "null argument");
  if (cap < b200m_keypoint_capacity(h, H, W)) return fail(B200M_ERR_INVALID, "keypoint capacity %d too small", cap);
  DeviceGuard dev_guard__(h->device);
  return with_graph(h, stream, {2, kp(images), ki(n_images, cap), ki(H, W), kp(keypoints), kp(scores), kp(descriptors),
                                kp(counts), kp(ws), (uint64_t)ws_bytes},
                    [&](void* st) {
                      return sp_forward_impl(h, st, images, true, n_images, H, W, keypoints, scores, descriptors, counts,
                                             cap, nullptr, nullptr, nullptr, 0, 0, ws, ws_bytes);
                    });
}

int b200m_superpoint_dense(b200m_handle* h, const float* images, int n_images, int H, int W, float* semi,
                           float* desc, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !images) return fail(B200M_ERR_INVALID, "null argument");
  return sp_forward_impl(h, stream, images, false, n_images, H, W, nullptr, nullptr, nullptr, nullptr, 0, semi, desc,
                         nullptr, 0, 0, ws, ws_bytes);
}

int b200m_detector_post(b200m_handle* h, const float* semi, int n_images, int hc, int wc, float* heat, float* nms,
                        float* keypoints, float* scores, int* counts, int cap, void* ws, size_t ws_bytes,
                        void* stream) {
  if (!h || !semi) return fail(B200M_ERR_INVALID, "null argument");
  SpDims d = sp_dims(h, hc * 8, wc * 8);
  if (keypoints && cap < b200m_keypoint_capacity(h, hc * 8, wc * 8))
    return fail(B200M_ERR_INVALID, "keypoint capacity %d too small", cap);
  Arena A(ws, ws_bytes);
  float* semi_c4 = A.take<float>((size_t)n_images * 128 * hc * wc);
  float* heat_ws = A.take<float>((size_t)n_images * d.H8 * d.W8);
  unsigned long long* keys = A.take<unsigned long long>((size_t)n_images * d.cand_cap);
  int* cc = A.take<int>(n_images + 2);
  unsigned char* nms_scr = A.take<unsigned char>(nms_scratch_bytes(n_images, d.H8, d.W8));
  if (!A.ok) return fail(B200M_ERR_WORKSPACE, "detector_post workspace too small: need %zu bytes", A.off);
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  launch_nchw_to_c4(ctx, semi, 65, semi_c4, 32, n_images, hc, wc);
  float* hp = heat ? heat : heat_ws;
  launch_softmax_heat(ctx, semi_c4, 32, hp, n_images, hc, wc);
  cudaMemsetAsync(cc, 0, sizeof(int) * (n_images + 2), ctx.stream);
  launch_nms_candidates(ctx, hp, nms, n_images, d.H8, d.W8, h->cfg.nms_radius, h->cfg.keypoint_threshold,
                        h->cfg.remove_borders, keypoints ? keys : nullptr, cc, d.cand_cap, cc + n_images, nms_scr);
  if (keypoints)
    launch_select_keypoints(ctx, keys, cc, d.cand_cap, n_images, d.W8, h->cfg.max_keypoints, keypoints, scores,
                            counts, cap);
  if (keypoints) launch_apply_flags(ctx, cc + n_images, counts, n_images);
  return finish(h, ctx);
}

int b200m_sample_descriptors(b200m_handle* h, const float* keypoints, const int* counts, const float* desc,
                             int n_images, int hc, int wc, int cap, float* descriptors, void* stream) {
  if (!h || !keypoints || !desc || !descriptors) return fail(B200M_ERR_INVALID, "null argument");
  const int D = h->cfg.descriptor_dim;
  // the stage API receives the reference-layout (n,D,h,w) map; the kernel wants C4-planar
  DeviceGuard dev_guard0__(h->device);
  float* tmp = nullptr;
  if (cudaMallocAsync(&tmp, (size_t)n_images * D * hc * wc * sizeof(float), (cudaStream_t)stream) != cudaSuccess)
    return fail(B200M_ERR_CUDA, "scratch allocation failed");
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  launch_nchw_to_c4(ctx, desc, D, tmp, D / 4, n_images, hc, wc);
  launch_sample_descriptors(ctx, tmp, D / 4, D, n_images, hc, wc, keypoints, counts, cap, h->cfg.align_corners,
                            descriptors, nullptr, 0, 0);
  cudaFreeAsync(tmp, (cudaStream_t)stream);
  return finish(h, ctx);
}

int b200m_knn_ratio_match(b200m_handle* h, const float* desc0, const float* desc1, const int* counts0,
                          const int* counts1, int B, int N, int M, float ratio, int64_t* matches, float* dist1,
                          float* dist2, void* stream) {
  if (!h || !desc0 || !desc1 || !matches || !dist1 || !dist2) return fail(B200M_ERR_INVALID, "null argument");
  if (M <= 0 && N > 0) return fail(B200M_ERR_INVALID, "empty train set");
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  if (!launch_knn_ratio(ctx, desc0, desc1, counts0, counts1, B, h->cfg.descriptor_dim, N, M, ratio,
                        (long long*)matches, dist1, dist2))
    return fail(B200M_ERR_INVALID, "descriptor_dim %d not supported by the matcher", h->cfg.descriptor_dim);
  return finish(h, ctx);
}

int b200m_estimate_affine_partial(b200m_handle* h, const float* kpts0, const float* kpts1, const int64_t* matches0,
                                  const int* counts0, int B, int N, int M, double ransac_reproj_threshold,
                                  int max_iters, double confidence, int refine_iters, double* matrices,
                                  uint8_t* inlier0, int* info, void* stream) {
  if (!h || !kpts0 || !kpts1 || !matches0 || !matrices || !inlier0 || !info)
    return fail(B200M_ERR_INVALID, "null argument");
  if (B < 0 || N < 0 || M < 0) return fail(B200M_ERR_INVALID, "negative size");
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  if (!launch_ransac_affine_partial(ctx, kpts0, kpts1, (const long long*)matches0, counts0, B, N, M,
                                    ransac_reproj_threshold, max_iters, confidence, refine_iters > 0, matrices,
                                    inlier0, info))
    return fail(B200M_ERR_INVALID, "N = %d keypoints per image exceed the estimator's shared-memory capacity", N);
  return finish(h, ctx);
}

int b200m_warp_affine(b200m_handle* h, const void* src, int dtype, int B, int src_h, int src_w,
                      const double* matrices, void* dst, int dst_h, int dst_w, void* stream) {
  if (!h || !src || !matrices || !dst) return fail(B200M_ERR_INVALID, "null argument");
  if (src_h <= 0 || src_w <= 0 || src_h > 32767 || src_w > 32767)
    return fail(B200M_ERR_INVALID, "source size %dx%d outside [1, 32767] (cv2's short coordinate maps)", src_h, src_w);
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  if (!launch_warp_affine(ctx, src, dtype, B, src_h, src_w, matrices, dst, dst_h, dst_w))
    return fail(B200M_ERR_INVALID, "dtype %d not supported (0 = uint8, 1 = float32, 2 = float64)", dtype);
  return finish(h, ctx);
}

size_t b200m_superglue_workspace_bytes(const b200m_handle* h, int B, int N, int M) {
  if (!h) return 0;
  Arena A(nullptr, 0);
  SgWs w;
  sg_carve(h, std::max(B, 1), N, M, A, w);
  // stage APIs also stage a dense Z-sized / S-sized scratch
  return A.off + 1024;
}

int b200m_superglue_forward(b200m_handle* h, const float* kpts0, const float* scores0, const float* desc0,
                            const int* counts0, const float* kpts1, const float* scores1, const float* desc1,
                            const int* counts1, int B, int N, int M, int H0, int W0, int H1, int W1,
                            int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1, void* ws,
                            size_t ws_bytes, void* stream) {
  if (!h) return fail(B200M_ERR_INVALID, "null handle");
  if (!h->packed_sg) return fail(B200M_ERR_WEIGHTS, "SuperGlue weights are not packed (b200m_set_tensor + b200m_pack)");
  if (B <= 0) return B200M_OK;
  DeviceGuard dev_guard__(h->device);
  return with_graph(h, stream, {3, kp(kpts0), kp(scores0), kp(desc0), kp(counts0), kp(kpts1), kp(scores1), kp(desc1),
                                kp(counts1), ki(B, N), ki(M, H0), ki(W0, H1), ki(W1), kp(matches0), kp(matches1),
                                kp(mscores0), kp(mscores1), kp(ws), (uint64_t)ws_bytes},
                    [&](void* stream) -> int {
  h->pdl_now = B <= h->pdl_max_pairs;
  LaunchCtx ctx = make_ctx(h, stream);
  if (N == 0 || M == 0) {   // superglue_test.py:235-242
    launch_match_select(ctx, nullptr, nullptr, nullptr, 0, nullptr, nullptr, B, N, M, 0.f, (long long*)matches0,
                        (long long*)matches1, mscores0, mscores1);
    return finish(h, ctx);
  }
  const int D = h->cfg.descriptor_dim;
  Arena A(ws, ws_bytes);
  SgWs w;
  if (!sg_carve(h, B, N, M, A, w)) return fail(B200M_ERR_WORKSPACE, "SuperGlue workspace too small: need %zu bytes", A.off);
  const size_t rows = (size_t)B * w.Np;
  launch_bcn_to_tokens(ctx, desc0, B, D, N, w.X, w.Np, 2 * D);
  launch_bcn_to_tokens(ctx, desc1, B, D, M, w.X + rows * 2 * D, w.Np, 2 * D);
  sg_core(h, ctx, w, kpts0, scores0, counts0, kpts1, scores1, counts1, B, N, M, H0, W0, H1, W1, matches0,
          matches1, mscores0, mscores1);
  return finish(h, ctx);
                    });
}

int b200m_keypoint_encode(b200m_handle* h, const float* kpts, const float* scores, const float* desc, int B, int N,
                          int H, int W, float* out, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !h->packed_sg) return fail(B200M_ERR_WEIGHTS, "SuperGlue weights are not packed");
  const int D = h->cfg.descriptor_dim;
  Arena A(ws, ws_bytes);
  SgWs w;
  if (!sg_carve(h, B, N, N, A, w)) return fail(B200M_ERR_WORKSPACE, "workspace too small: need %zu bytes", A.off);
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  launch_bcn_to_tokens(ctx, desc, B, D, N, w.X, w.Np, 2 * D);
  sg_kenc(h, ctx, w, 0, kpts, scores, B, N, H, W);
  launch_tokens_to_bcn(ctx, w.X, w.Np, 2 * D, out, B, D, N);
  return finish(h, ctx);
}

int b200m_gnn(b200m_handle* h, const float* desc0, const float* desc1, const int* counts0, const int* counts1,
              int B, int N, int M, int layer_begin, int layer_end, float* out0, float* out1, void* ws,
              size_t ws_bytes, void* stream) {
  if (!h || !h->packed_sg) return fail(B200M_ERR_WEIGHTS, "SuperGlue weights are not packed");
  if (layer_begin < 0 || layer_end > h->cfg.n_gnn_layers || layer_begin > layer_end)
    return fail(B200M_ERR_INVALID, "bad layer range");
  const int D = h->cfg.descriptor_dim;
  Arena A(ws, ws_bytes);
  SgWs w;
  if (!sg_carve(h, B, N, M, A, w)) return fail(B200M_ERR_WORKSPACE, "workspace too small: need %zu bytes", A.off);
  const size_t rows = (size_t)B * w.Np;
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  launch_bcn_to_tokens(ctx, desc0, B, D, N, w.X, w.Np, 2 * D);
  launch_bcn_to_tokens(ctx, desc1, B, D, M, w.X + rows * 2 * D, w.Np, 2 * D);
  sg_gnn(h, ctx, w, B, counts0, counts1, N, M, layer_begin, layer_end);
  launch_tokens_to_bcn(ctx, w.X, w.Np, 2 * D, out0, B, D, N);
  launch_tokens_to_bcn(ctx, w.X + rows * 2 * D, w.Np, 2 * D, out1, B, D, M);
  return finish(h, ctx);
}

int b200m_score_matrix(b200m_handle* h, const float* desc0, const float* desc1, int B, int N, int M, float* S,
                       void* ws, size_t ws_bytes, void* stream) {
  if (!h || !h->packed_sg) return fail(B200M_ERR_WEIGHTS, "SuperGlue weights are not packed");
  const int D = h->cfg.descriptor_dim;
  Arena A(ws, ws_bytes);
  SgWs w;
  if (!sg_carve(h, B, N, M, A, w)) return fail(B200M_ERR_WORKSPACE, "workspace too small: need %zu bytes", A.off);
  const size_t rows = (size_t)B * w.Np;
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  launch_bcn_to_tokens(ctx, desc0, B, D, N, w.X, w.Np, 2 * D);
  launch_bcn_to_tokens(ctx, desc1, B, D, M, w.X + rows * 2 * D, w.Np, 2 * D);
  sg_scores(h, ctx, w, B, N, M);
  // compact (B,N,ldS) -> (B,N,M)
  cudaMemcpy2DAsync(S, (size_t)M * sizeof(float), w.S, (size_t)w.ldS * sizeof(float), (size_t)M * sizeof(float),
                    (size_t)B * N, cudaMemcpyDeviceToDevice, ctx.stream);
  return finish(h, ctx);
}

int b200m_sinkhorn(b200m_handle* h, const float* S, int B, int N, int M, int iters, float* Z, void* ws,
                   size_t ws_bytes, void* stream) {
  if (!h || !S || !Z) return fail(B200M_ERR_INVALID, "null argument");
  Arena A(ws, ws_bytes);
  SgWs w;
  w.ld_uv = round_up(std::max(N, M) + 1, 4);
  w.u = A.take<float>((size_t)B * w.ld_uv);
  w.v = A.take<float>((size_t)B * w.ld_uv);
  w.ot_part = A.take<float>(ot_part_floats(B, N, M, M));
  if (!A.ok) return fail(B200M_ERR_WORKSPACE, "workspace too small: need %zu bytes", A.off);
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  OtParams p = ot_params(h, w, S, M, (long long)N * M, B, N, M, nullptr, nullptr);
  sg_sinkhorn(h, ctx, p, iters, w.ot_part);
  launch_ot_write_Z(ctx, p, Z);
  return finish(h, ctx);
}

int b200m_match_select(b200m_handle* h, const float* Z, int B, int N, int M, int64_t* matches0, int64_t* matches1,
                       float* mscores0, float* mscores1, void* ws, size_t ws_bytes, void* stream) {
  if (!h || !Z) return fail(B200M_ERR_INVALID, "null argument");
  Arena A(ws, ws_bytes);
  const int ld = round_up(std::max(N, M) + 1, 4);
  float* max0 = A.take<float>((size_t)B * ld);
  int* idx0 = A.take<int>((size_t)B * ld);
  int* idx1 = A.take<int>((size_t)B * ld);
  if (!A.ok) return fail(B200M_ERR_WORKSPACE, "workspace too small: need %zu bytes", A.off);
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  launch_dense_argmax(ctx, Z, B, N, M, idx0, max0, idx1, ld);
  launch_match_select(ctx, idx0, max0, idx1, ld, nullptr, nullptr, B, N, M, h->cfg.match_threshold,
                      (long long*)matches0, (long long*)matches1, mscores0, mscores1);
  return finish(h, ctx);
}

int b200m_resize_linear_u8(b200m_handle* h, const uint8_t* src, int B, int src_h, int src_w, uint8_t* dst, int dst_h,
                           int dst_w, void* stream) {
  if (!h || !src || !dst) return fail(B200M_ERR_INVALID, "null argument");
  if (B < 0 || src_h <= 0 || src_w <= 0 || dst_h <= 0 || dst_w <= 0) return fail(B200M_ERR_INVALID, "bad image size");
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  launch_resize_linear_u8(ctx, src, B, src_h, src_w, dst, dst_h, dst_w);
  return finish(h, ctx);
}

int b200m_pack_match_wire(b200m_handle* h, const int64_t* matches0, const float* mscores0, int B_valid, int B_wire,
                          int N, int ld, int32_t* wire, void* stream) {
  if (!h || !wire || (B_valid > 0 && (!matches0 || !mscores0))) return fail(B200M_ERR_INVALID, "null argument");
  if (B_valid < 0 || B_wire < B_valid || N < 0 || ld < N) return fail(B200M_ERR_INVALID, "bad wire shape");
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  launch_pack_match_wire(ctx, (const long long*)matches0, mscores0, B_valid, B_wire, N, ld, wire);
  return finish(h, ctx);
}

int b200m_unpack_match_wire(b200m_handle* h, const int32_t* wire, int world, int B_wire, int n_pairs, int N,
                            int64_t* matches0, float* mscores0, void* stream) {
  if (!h || !wire || !matches0 || !mscores0) return fail(B200M_ERR_INVALID, "null argument");
  if (world <= 0 || n_pairs < 0 || N < 0 || (long long)B_wire * world < n_pairs)
    return fail(B200M_ERR_INVALID, "bad wire shape");
  DeviceGuard dev_guard__(h->device);
  LaunchCtx ctx = make_ctx(h, stream);
  launch_unpack_match_wire(ctx, wire, world, B_wire, n_pairs, N, (long long*)matches0, mscores0);
  return finish(h, ctx);
}

size_t b200m_matching_workspace_bytes(const b200m_handle* h, int B, int H, int W) {
  if (!h) return 0;
  int cap = b200m_keypoint_capacity(h, H, W);
  // (small batches: one SuperPoint workspace per image side, the two passes run concurrently)
  return align_up(sp_ws_bytes(h, B, H, W), 256) * (B <= h->sp_dual_max ? 2 : 1) +
         b200m_superglue_workspace_bytes(h, B, cap, cap) + 256;
}

static int matching_forward_impl(b200m_handle* h, const void* image0, const void* image1, bool images_u8, int B, int H,
                                 int W, float* keypoints0, float* scores0, float* descriptors0, int* counts0,
                                 float* keypoints1, float* scores1, float* descriptors1, int* counts1, int cap,
                                 int64_t* matches0, int64_t* matches1, float* mscores0, float* mscores1, void* ws,
                                 size_t ws_bytes, void* stream) {
  if (!h || !image0 || !image1) return fail(B200M_ERR_INVALID, "null argument");
  if (!h->packed_sp || !h->packed_sg) return fail(B200M_ERR_WEIGHTS, "both SuperPoint and SuperGlue weights must be packed");
  if (B <= 0) return B200M_OK;
  if (cap < b200m_keypoint_capacity(h, H, W) || cap <= 0) return fail(B200M_ERR_INVALID, "keypoint capacity %d too small", cap);
  DeviceGuard dev_guard0__(h->device);
  return with_graph(h, stream, {images_u8 ? 5u : 4u, kp(image0), kp(image1), ki(B, cap), ki(H, W), kp(keypoints0),
                                kp(scores0), kp(descriptors0), kp(counts0), kp(keypoints1), kp(scores1), kp(descriptors1),
                                kp(counts1), kp(matches0), kp(matches1), kp(mscores0), kp(mscores1), kp(ws),
                                (uint64_t)ws_bytes},
                    [&](void* stream) -> int {
  const int D = h->cfg.descriptor_dim;
  const size_t sp_bytes = align_up(sp_ws_bytes(h, B, H, W), 256);
  // Few pairs (the reference caller's loop runs ONE pair per call): a single image leaves most of the chip idle in the
  // 1/4- and 1/8-resolution layers and in the detector post-processing, so the two images' SuperPoint passes run
  // side by side -- image 1 on a forked stream with its own workspace (inside the captured graph: a parallel branch).
  bool dual = B <= h->sp_dual_max && !h->prof.enabled &&
              ws_bytes >= 2 * sp_bytes + b200m_superglue_workspace_bytes(h, B, cap, cap);   // (else: one after the other)
  if (dual && !h->side_stream) {
    if (cudaStreamCreateWithFlags(&h->side_stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->fork_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      h->sp_dual_max = 0;
      dual = false;
    }
  }
  const size_t sp_total = dual ? 2 * sp_bytes : sp_bytes;
  if (ws_bytes < sp_total) return fail(B200M_ERR_WORKSPACE, "matching workspace too small");
  Arena A((char*)ws + sp_total, ws_bytes - sp_total);
  SgWs w;
  if (!sg_carve(h, B, cap, cap, A, w))
    return fail(B200M_ERR_WORKSPACE, "matching workspace too small: need %zu bytes", sp_bytes + A.off);
  const size_t rows = (size_t)B * w.Np;
  // token rows in [cap, Np) are never written by the sampler but are read (masked) as attention keys,
  // so they must be finite
  if (w.Np != cap) cudaMemsetAsync(w.X, 0, w.rows * 2 * D * sizeof(float), (cudaStream_t)stream);
  // SuperPoint writes token-major descriptors straight into X[:, :D] of its side
  void* stream1 = stream;
  if (dual) {
    cudaError_t fe = cudaEventRecord(h->fork_ev, (cudaStream_t)stream);
    if (fe == cudaSuccess) fe = cudaStreamWaitEvent(h->side_stream, h->fork_ev, 0);
    if (fe != cudaSuccess) return fail(B200M_ERR_CUDA, "stream fork failed: %s", cudaGetErrorString(fe));
    stream1 = (void*)h->side_stream;
  }
  int rc = sp_forward_impl(h, stream, image0, images_u8, B, H, W, keypoints0, scores0, descriptors0, counts0, cap, nullptr,
                           nullptr, w.X, 2 * D, (size_t)w.Np * 2 * D, ws, sp_bytes);
  const int rc1 = sp_forward_impl(h, stream1, image1, images_u8, B, H, W, keypoints1, scores1, descriptors1, counts1, cap,
                                  nullptr, nullptr, w.X + rows * 2 * D, 2 * D, (size_t)w.Np * 2 * D,
                                  (char*)ws + (dual ? sp_bytes : 0), sp_bytes);
  if (dual) {                      // join (also on errors: a captured side stream must not be left dangling)
    cudaError_t je = cudaEventRecord(h->join_ev, h->side_stream);
    if (je == cudaSuccess) je = cudaStreamWaitEvent((cudaStream_t)stream, h->join_ev, 0);
    if (je != cudaSuccess && !rc && !rc1) return fail(B200M_ERR_CUDA, "stream join failed: %s", cudaGetErrorString(je));
  }
  if (rc) return rc;
  if (rc1) return rc1;
  DeviceGuard dev_guard__(h->device);
  h->pdl_now = B <= h->pdl_max_pairs;
  LaunchCtx ctx = make_ctx(h, stream);
  sg_core(h, ctx, w, keypoints0, scores0, counts0, keypoints1, scores1, counts1, B, cap, cap, H, W, H, W, matches0,
          matches1, mscores0, mscores1);
  return finish(h, ctx);
                    });
}

int b200m_matching_forward(b200m_handle* h, const float* image0, const float* image1, int B, int H, int W,
                           float* keypoints0, float* scores0, float* descriptors0, int* counts0, float* keypoints1,
                           float* scores1, float* descriptors1, int* counts1, int cap, int64_t* matches0,
                           int64_t* matches1, float* mscores0, float* mscores1, void* ws, size_t ws_bytes,
                           void* stream) {
  return matching_forward_impl(h, image0, image1, false, B, H, W, keypoints0, scores0, descriptors0, counts0, keypoints1,
                               scores1, descriptors1, counts1, cap, matches0, matches1, mscores0, mscores1, ws, ws_bytes,
                               stream);
}

int b200m_matching_forward_u8(b200m_handle* h, const uint8_t* image0, const uint8_t* image1, int B, int H, int W,
                              float* keypoints0, float* scores0, float* descriptors0, int* counts0, float* keypoints1,
                              float* scores1, float* descriptors1, int* counts1, int cap, int64_t* matches0,
                              int64_t* matches1, float* mscores0, float* mscores1, void* ws, size_t ws_bytes,
                              void* stream) {
  return matching_forward_impl(h, image0, image1, true, B, H, W, keypoints0, scores0, descriptors0, counts0, keypoints1,
                               scores1, descriptors1, counts1, cap, matches0, matches1, mscores0, mscores1, ws, ws_bytes,
                               stream);
}

}  // extern "C"
