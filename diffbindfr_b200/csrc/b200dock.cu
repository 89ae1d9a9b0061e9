// libb200dock: handle, workspace, weight upload and the per-step orchestration behind the C ABI
// declared in include/b200dock.h.  Everything is launched on the caller's stream with fixed-size
// grids that read their extents from device memory, so one evaluation needs no host round trip.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>   // header-only; ranges are emitted only when B200DOCK_NVTX=1 (nsys / ncu --nvtx)

#include "common.cuh"
#include "graph.cuh"
#include "embed.cuh"
#include "conv.cuh"
#include "tc_common.cuh"
#include "conv_fused.cuh"
#include "conv_fused2.cuh"
#include "heads.cuh"
#include "pose.cuh"
#include "mdn.cuh"
#include "mdn_enc.cuh"
#include "mdn_feat.cuh"
#include "assemble.cuh"
#include "vina.cuh"

#define CK(call)                                                                      \
  do {                                                                                \
    cudaError_t _e = (call);                                                          \
    if (_e != cudaSuccess) {                                                          \
      h->err = std::string(#call) + ": " + cudaGetErrorString(_e);                    \
      return B200_ERR_CUDA;                                                           \
    }                                                                                 \
  } while (0)
#define FAIL(code, msg) do { h->err = (msg); return (code); } while (0)

namespace {

struct Buf {
  void* p = nullptr; size_t cap = 0;
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct ConvWs {                 // per edge family: lig, atom, al, la, tor, sc
  int T = 0; int cap = 0; int z_max = 0;
  Buf counts, seg, gpad, es, ed, eaux, emb, sh, H1, Zt, msg, agg, part;
};

struct ConvW {                  // views into the device weight blob
  const float *W1t, *b1, *W2p; LnParams ln;
  int n_cols = 0;
  const __half *W1h16 = nullptr, *W1l16 = nullptr, *W2h16 = nullptr, *W2l16 = nullptr; float inv_s1 = 1.f, inv_s2 = 1.f;
};

// conv kernels: 0 exact fp32 SIMT (192-column units), 5 fused tcgen05 single CTA, 6 fused tcgen05 CTA pairs + fused scatter (144),
// 10 the pair kernel re-pipelined over two A buffers with 96-column units (no tile-transition bubble)
inline bool kernel_known(int k) { return k == 0 || k == 5 || k == 6 || k == 11; }
inline int variant_of_kernel(int k) { return k == 0 ? 0 : 1; }
inline bool kernel_keeps_msg(int k) { return k == 0 || k == 5; }

}  // namespace

struct B200Handle {
  int device = 0;
  std::string err;
  B200Config cfg;
  std::vector<std::vector<int32_t>> hold_i; std::vector<std::vector<float>> hold_f;
  DevPlan dplans[B200_N_PLANS];
  int variant = 0;              // slot of c_plans this handle's plans live in
  std::vector<void*> plan_allocs;
  int* d_tor_cg_ijk = nullptr; float* d_tor_cg_val = nullptr;
  float* d_blob = nullptr; __half* d_w16 = nullptr; size_t blob_n = 0; std::vector<int64_t> off; bool weights = false;
  ConvW convw[26];
  // workspace
  ConvWs cw[6];
  Buf pre[6], h_lig, h_atom, jmax_lig, jmax_atom, centre, cmsg, s_tr, s_rot, s_tor, s_sc, atom14, errflag;
  Buf c_temb, c_trs, c_rotn, c_torn, c_scn, temb_steps;
  Buf mdn_w, mdn_A, mdn_B; bool mdn_weights = false;
  Buf enc_w, enc_ws; std::vector<int64_t> enc_off; bool enc_weights = false;
  // host-batch path
  Buf pinned_in, dev_in, dev_noise, dev_lig_out, dev_a14_out, pinned_out;
  int64_t launches = 0;
  int64_t last_counts[5] = {0, 0, 0, 0, 0};
  bool profiling = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tp_events;
  std::vector<cudaEvent_t> event_pool; size_t event_used = 0;
  B200Batch last_batch; bool have_last = false;
  int n_sms = 148;
  int debug_layers = 6;
  int tp_grid = 148;
  std::vector<float> cg_dense; int atom14_group[21 * 14];
  // side stream: independent small kernels (graph families, ligand vs pocket node updates, centre head) run concurrently
  bool nvtx = false, warp_node_update = false;
  int* host_meta = nullptr; bool deferred_check = false; size_t feat_smem = 48 * 1024, vina_smem = 48 * 1024;
  Buf ex_in, ex_pin, ex_out[32];   // batch assembly: staged base batch (device / pinned) and the expanded arrays
  Buf trace; bool trace_on = false;
  cudaStream_t side = nullptr; cudaEvent_t ev_fork = nullptr, ev_join = nullptr; bool use_side = true;
};

namespace {

struct NvtxRange {               // RAII range on the calling thread: marks the phases of one evaluation in a timeline
  bool on;
  NvtxRange(const B200Handle* h, const char* name) : on(h->nvtx) { if (on) nvtxRangePushA(name); }
  ~NvtxRange() { if (on) nvtxRangePop(); }
};

int ensure(B200Handle* h, Buf& b, size_t bytes) {
  if (bytes <= b.cap) return B200_OK;
  if (b.p) CK(cudaFree(b.p));
  b.p = nullptr; b.cap = 0;
  size_t want = bytes + bytes / 8 + 256;
  CK(cudaMalloc(&b.p, want));
  b.cap = want;
  return B200_OK;
}
#define ENS(buf, bytes) do { int _r = ensure(h, buf, (size_t)(bytes)); if (_r) return _r; } while (0)

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
inline int grid_for(long long n, int threads, int max_blocks) {
  long long g = (n + threads - 1) / threads;
  if (g < 1) g = 1;
  if (g > max_blocks) g = max_blocks;
  return (int)g;
}

template <typename T>
int upload(B200Handle* h, const T* src, size_t n, T** dst) {
  CK(cudaMalloc((void**)dst, n * sizeof(T) + 16));
  h->plan_allocs.push_back(*dst);
  if (n) CK(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
  return B200_OK;
}

int plan_of_layer(int l) { return l < 3 ? l : 3; }

cudaEvent_t get_event(B200Handle* h) {
  if (h->event_used == h->event_pool.size()) {
    cudaEvent_t e; cudaEventCreate(&e); h->event_pool.push_back(e);
  }
  return h->event_pool[h->event_used++];
}

int setup_workspace(B200Handle* h, const B200Batch& b) {
  if (b.B <= 0 || b.N_l <= 0 || b.N_a <= 0 || b.N_r <= 0) FAIL(B200_ERR_INVALID, "empty batch");
  if (b.max_lig_atoms > POSE_MAX_ATOMS) FAIL(B200_ERR_INVALID, "ligand larger than 256 heavy atoms is not supported");
  auto r128 = [](long long x) { return (int)(((x + 127) / 128) * 128 + 128); };
  long long caps[6];
  caps[0] = 33LL * b.N_l + b.E_b;
  caps[1] = std::min<long long>(64LL * b.N_a, b.atom_pairs);
  caps[2] = caps[3] = b.cross_pairs;
  caps[4] = 32LL * b.n_tor;
  caps[5] = 32LL * b.n_sc;
  for (int c = 0; c < 6; ++c) caps[c] += 32LL * b.B;          // every graph's edge range is padded to a multiple of 32 slots
  int Ts[6] = {b.N_l, b.N_a, b.N_l, b.N_a, b.n_tor, b.n_sc};
  int zmax[6] = {624, 624, 624, 624, 144, 144};
  for (int c = 0; c < 6; ++c) {
    ConvWs& w = h->cw[c];
    w.T = Ts[c]; w.cap = r128(caps[c]); w.z_max = zmax[c];
    ENS(w.counts, (size_t)(w.T + 1) * 4); ENS(w.seg, (size_t)(w.T + 2) * 4); ENS(w.gpad, (size_t)(b.B + 1) * 4);
    ENS(w.es, (size_t)w.cap * 4); ENS(w.ed, (size_t)w.cap * 4);
    if (c == 0) ENS(w.eaux, (size_t)w.cap * 4);
    ENS(w.emb, (size_t)w.cap * NSC * 4); ENS(w.sh, (size_t)w.cap * 9 * 4);
    if (h->cfg.conv_kernel == 0) {
      ENS(w.H1, (size_t)w.cap * KP * 4);
      ENS(w.Zt, (size_t)w.cap * w.z_max * 4);
    }
    if (kernel_keeps_msg(h->cfg.conv_kernel)) ENS(w.msg, (size_t)w.cap * HS * 4);
    ENS(w.agg, (size_t)(w.T + 1) * HS * 4); ENS(w.part, (size_t)(w.cap / 32) * 2 * HS * 4);
  }
  for (int m = 0; m < 6; ++m) ENS(h->pre[m], (size_t)b.B * NSC * 4);
  ENS(h->h_lig, (size_t)b.N_l * HS * 4); ENS(h->h_atom, (size_t)b.N_a * HS * 4);
  ENS(h->jmax_lig, (size_t)b.N_l * 4); ENS(h->jmax_atom, (size_t)b.N_a * 4);
  ENS(h->centre, (size_t)b.B * 3 * 4); ENS(h->cmsg, (size_t)b.N_l * 12 * 4);
  ENS(h->s_tr, (size_t)b.B * 3 * 4); ENS(h->s_rot, (size_t)b.B * 3 * 4);
  ENS(h->s_tor, (size_t)(b.n_tor + 1) * 4); ENS(h->s_sc, (size_t)(b.n_sc + 1) * 4);
  ENS(h->atom14, (size_t)b.N_r * 14 * 3 * 4);
  ENS(h->errflag, 16);
  ENS(h->c_temb, (size_t)b.B * SIG * 4); ENS(h->c_trs, (size_t)b.B * 4); ENS(h->c_rotn, (size_t)b.B * 4);
  ENS(h->c_torn, (size_t)(b.n_tor + 1) * 4); ENS(h->c_scn, (size_t)(b.n_sc + 1) * 4);
  return B200_OK;
}

EdgeMlp edge_mlp(const B200Handle* h, int section, int n_bond, int n_sigma) {
  EdgeMlp m; m.w = h->d_blob + h->off[section]; m.n_bond = n_bond; m.n_sigma = n_sigma; return m;
}

template <int KIND>
int build_graph(B200Handle* h, const GraphArgs& G, ConvWs& w, cudaStream_t st) {
  if (w.T == 0) return B200_OK;
  k_graph_count<KIND><<<grid_for((long long)w.T * 32, 256, 148 * 8), 256, 0, st>>>(G, w.T, w.counts.as<int>());
  k_scan_aligned<KIND><<<1, 1024, 0, st>>>(G, w.T, w.cap - 128, w.counts.as<int>(), w.seg.as<int>(), w.gpad.as<int>(),
                                           h->errflag.as<int>());
  k_graph_fill<KIND><<<grid_for((long long)w.T * 32, 256, 148 * 8), 256, 0, st>>>(G, w.T, w.seg.as<int>(), w.counts.as<int>(), w.cap,
                                                               w.es.as<int>(), w.ed.as<int>(),
                                                               KIND == G_LIG ? w.eaux.as<int>() : nullptr);
  h->launches += 3;
  return B200_OK;
}

template <int KIND>
void edge_feat(B200Handle* h, const B200Batch& b, ConvWs& w, EdgeMlp mlp, const float* pre, float stop, cudaStream_t st) {
  if (w.T == 0) return;
  EdgeFeatArgs A{};
  A.n_edges = w.seg.as<int>() + w.T; A.es = w.es.as<int>(); A.ed = w.ed.as<int>(); A.eaux = w.eaux.as<int>();
  if (KIND == G_LIG) { A.pos_s = b.lig_pos; A.pos_d = b.lig_pos; A.batch_for_pre = b.lig_batch; }
  if (KIND == G_ATOM) { A.pos_s = b.rec_atm_pos; A.pos_d = b.rec_atm_pos; A.batch_for_pre = b.atom_batch; }
  if (KIND == G_AL) { A.pos_s = b.lig_pos; A.pos_d = b.rec_atm_pos; A.batch_for_pre = b.lig_batch; }
  if (KIND == G_LA) { A.pos_s = b.rec_atm_pos; A.pos_d = b.lig_pos; A.batch_for_pre = b.lig_batch; }
  if (KIND == G_TOR) { A.pos_s = b.lig_pos; A.pos_d = b.lig_pos; A.bonds = b.tor_bonds; }
  if (KIND == G_SC) { A.pos_s = b.rec_atm_pos; A.pos_d = b.rec_atm_pos; A.bonds = b.sc_bonds; }
  A.pre = pre; A.lig_edge_feat = b.lig_edge_feat; A.mlp = mlp; A.stop = stop;
  A.emb = w.emb.as<float>(); A.sh = w.sh.as<float>();
  A.cg_ijk = h->d_tor_cg_ijk; A.cg_val = h->d_tor_cg_val;
  for (int i = 0; i < 4; ++i) A.cg_off[i] = h->cfg.tor_cg_off[i];
  size_t smem = (size_t)((mlp.n_bond + SIG) * NSC + NSC * NSC + 2 * NSC + SIG) * 4;
  k_edge_feat<KIND><<<148 * 4, 128, smem, st>>>(A);
  h->launches += 1;
}

int launch_tp(B200Handle* h, const ConvLaunch& L, const Fused16Extra& F, cudaStream_t st) {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (h->profiling) { e0 = get_event(h); e1 = get_event(h); cudaEventRecord(e0, st); }
  int rc = B200_OK;
  const int k = h->cfg.conv_kernel;
  if (k == 6) {
    static const int dbg = getenv("B200DOCK_DBG") ? atoi(getenv("B200DOCK_DBG")) : 0;   // honoured by the -DB200DOCK_TRACE build only
    ConvLaunch L2 = L; L2.dbg = dbg;
    rc = launch_conv_fused16x2(L2, F, h->tp_grid, st);
  }
  else if (k == 11) {
    static const int dbg = getenv("B200DOCK_DBG") ? atoi(getenv("B200DOCK_DBG")) : 0;
    ConvLaunch L2 = L; L2.dbg = dbg;
    rc = launch_conv_fused16x2(L2, F, h->tp_grid, st, true);
  }
  else if (k == 5) rc = launch_conv_fused16(L, F, h->tp_grid, st);
  else {
    k_conv_prologue<<<h->n_sms, PRO_THREADS, PRO_SMEM, st>>>(L);
    k_conv_tp_simt<<<h->n_sms, TP_THREADS, TP_SMEM, st>>>(L);
    h->launches += 1;
  }
  h->launches += 1;
  if (kernel_keeps_msg(k)) {      // per-edge messages -> per-node sums (the CTA-pair kernel does this in its epilogue)
    MsgScatterLaunch M{};
    M.n = L.n;
    for (int i = 0; i < L.n; ++i) {
      const ConvArgs& C = L.c[i];
      M.c[i] = MsgScatterArgs{C.n_edges, C.es, C.seg, C.counts, C.msg, C.agg, C.part, h->dplans[C.cgp].out_dim};
    }
    k_msg_scatter<<<h->n_sms * 4, 256, 0, st>>>(M);
    h->launches += 1;
  }
  if (h->profiling) { cudaEventRecord(e1, st); h->tp_events.push_back({e0, e1}); }
  if (rc) { char m[64]; snprintf(m, sizeof m, "tcgen05 conv launch failed (%d)", rc); FAIL(B200_ERR_CUDA, m); }
  return B200_OK;
}

ConvArgs conv_args(B200Handle* h, ConvWs& w, int widx, int plan, const float* tabA, const float* tabB, int mode,
                   const int* bonds, int sh_stride, Fused16Extra& X, int slot) {
  ConvArgs C{};
  const ConvW& cw = h->convw[widx];
  X.W1hi[slot] = cw.W1h16; X.W1lo[slot] = cw.W1l16; X.W2hi[slot] = cw.W2h16; X.W2lo[slot] = cw.W2l16;
  X.w2_rows[slot] = (uint64_t)cw.n_cols + 144;
  C.n_edges = w.seg.as<int>() + w.T; C.es = w.es.as<int>(); C.ed = w.ed.as<int>();
  C.emb = w.emb.as<float>(); C.sh = w.sh.as<float>(); C.sh_stride = sh_stride;
  C.tabA = tabA; C.tabB = tabB; C.bonds = bonds; C.mode = mode;
  C.plan = h->variant * B200_N_PLANS + plan; C.cgp = plan;
  C.W1t = cw.W1t; C.b1 = cw.b1; C.W2p = cw.W2p;
  C.inv_s1 = cw.inv_s1; C.inv_s2 = cw.inv_s2;
  C.H1 = w.H1.as<float>(); C.Zt = w.Zt.as<float>(); C.msg = w.msg.as<float>();
  C.seg = w.seg.as<int>(); C.counts = w.counts.as<int>(); C.agg = w.agg.as<float>(); C.part = w.part.as<float>();
  return C;
}

AggSrc agg_src(const ConvWs& w) { return AggSrc{w.seg.as<int>(), w.counts.as<int>(), w.agg.as<float>(), w.part.as<float>()}; }

// The conv plans live in __constant__ memory, one slot per unit-width variant and device; a slot's content is a pure function of
// the static network spec, so it is written once (first handle of that variant on the device) and never changes afterwards:
// handles with different conv kernels, on any stream or thread, cannot disturb each other's in-flight kernels.
static std::mutex g_plan_mutex;
static bool g_plan_done[64][B200_PLAN_VARIANTS] = {};
static bool g_cg_done[64] = {};

int bind_plans(B200Handle* h) {
  std::lock_guard<std::mutex> lock(g_plan_mutex);
  const int dev = h->device & 63;
  if (!g_plan_done[dev][h->variant]) {
    CK(cudaMemcpyToSymbol(c_plans, h->dplans, sizeof(DevPlan) * B200_N_PLANS, sizeof(DevPlan) * B200_N_PLANS * h->variant));
    g_plan_done[dev][h->variant] = true;
  }
  if (!g_cg_done[dev]) {
    CK(cudaMemcpyToSymbol(c_cg_dense, h->cg_dense.data(), h->cg_dense.size() * sizeof(float)));
    CK(cudaMemcpyToSymbol(c_atom14_group, h->atom14_group, sizeof(int) * 21 * 14));
    g_cg_done[dev] = true;
  }
  return B200_OK;
}

// fork: the side stream continues from the current point of `st`; join: `st` waits for everything queued on the side stream
static inline cudaStream_t side_fork(B200Handle* h, cudaStream_t st) {
  if (!h->use_side) return st;
  cudaEventRecord(h->ev_fork, st);
  cudaStreamWaitEvent(h->side, h->ev_fork, 0);
  return h->side;
}
static inline void side_join(B200Handle* h, cudaStream_t st) {
  if (!h->use_side) return;
  cudaEventRecord(h->ev_join, h->side);
  cudaStreamWaitEvent(st, h->ev_join, 0);
}

int score_device(B200Handle* h, const B200Batch& b, const B200Cond& c, float* tr, float* rot, float* tor, float* sc,
                 cudaStream_t st) {
  if (!h->weights) FAIL(B200_ERR_STATE, "weights not loaded");
  NvtxRange r_score(h, "b200dock.score_network");
  const float* W = h->d_blob;
  const std::vector<int64_t>& off = h->off;
  // ---- sigma pre-activations
  {
    PreArgs P{};
    int secs[6] = {B200_W_LIG_NODE, B200_W_LIG_EDGE, B200_W_ATOM_EMB, B200_W_ATOM_EDGE, B200_W_LA_EDGE, B200_W_CENTER_EDGE};
    int ins[6] = {59, 74, 0, 64, 64, 64};
    int sig_off[6] = {27, 10, 48, 0, 0, 0};
    for (int m = 0; m < 6; ++m) {
      const float* rec = W + off[secs[m]];
      if (m == 2) { P.w0t[m] = rec + 86 * NSC; P.b0[m] = nullptr; }           // scalar_lin, no bias
      else { P.w0t[m] = rec; P.b0[m] = rec + (size_t)ins[m] * NSC; }
      P.sig_off[m] = sig_off[m]; P.out[m] = h->pre[m].as<float>();
    }
    P.n = 6;
    k_graph_pre<<<dim3(b.B, 6), 64, 0, st>>>(P, c.time_emb, b.B);
    h->launches += 1;
  }
  cudaStream_t s2 = side_fork(h, st);   // from here to the first conv layer: pocket-side work on the side stream, ligand-side on the caller's
  k_lig_node_embed<<<grid_for(b.N_l, 128, 148 * 4), 128, 0, st>>>(b.lig_node, b.lig_batch, b.N_l, W + off[B200_W_LIG_NODE],
                                                                   h->pre[0].as<float>(), h->h_lig.as<float>());
  k_atom_node_embed<<<grid_for(b.N_a, 128, 148 * 4), 128, 0, s2>>>(b.pocket_feat, b.atom_batch, b.N_a, W + off[B200_W_ATOM_EMB],
                                                                    h->pre[2].as<float>(), h->h_atom.as<float>());
  h->launches += 2;
  // ---- graphs
  GraphArgs G{};
  G.lig_pos = b.lig_pos; G.lig_batch = b.lig_batch; G.lig_ptr = b.lig_ptr; G.N_l = b.N_l;
  G.atom_pos = b.rec_atm_pos; G.atom_batch = b.atom_batch; G.atom_ptr = b.atom_ptr; G.N_a = b.N_a;
  G.pocket_feat = b.pocket_feat; G.tr_sigma = c.tr_sigma;
  G.bond_ptr = b.bond_ptr; G.bond_dst = b.bond_dst; G.bond_eid = b.bond_eid;
  G.tor_bonds = b.tor_bonds; G.n_tor = b.n_tor; G.sc_bonds = b.sc_bonds; G.n_sc = b.n_sc;
  G.tor_ptr = b.tor_ptr; G.sc_ptr = b.sc_ptr; G.B = b.B;
  G.lig_jmax = h->jmax_lig.as<int>(); G.atom_jmax = h->jmax_atom.as<int>();
  k_radius_cap<<<grid_for(b.N_l, 128, 148 * 8), 128, 0, st>>>(b.lig_pos, b.lig_batch, b.lig_ptr, b.N_l, 25.0f, 33, h->jmax_lig.as<int>());
  k_radius_cap<<<grid_for(b.N_a, 128, 148 * 8), 128, 0, s2>>>(b.rec_atm_pos, b.atom_batch, b.atom_ptr, b.N_a, 16.0f, 1001, h->jmax_atom.as<int>());
  h->launches += 2;
  int rc;
  {   // the four graph families are independent: pocket-side chains on the side stream, ligand-side chains on the caller's
    if ((rc = build_graph<G_ATOM>(h, G, h->cw[1], s2))) return rc;
    if ((rc = build_graph<G_LIG>(h, G, h->cw[0], st))) return rc;
    if ((rc = build_graph<G_LA>(h, G, h->cw[3], s2))) return rc;
    if ((rc = build_graph<G_AL>(h, G, h->cw[2], st))) return rc;
    edge_feat<G_ATOM>(h, b, h->cw[1], edge_mlp(h, B200_W_ATOM_EDGE, 0, 32), h->pre[3].as<float>(), 4.0f, s2);
    edge_feat<G_LIG>(h, b, h->cw[0], edge_mlp(h, B200_W_LIG_EDGE, 10, 32), h->pre[1].as<float>(), 5.0f, st);
    edge_feat<G_LA>(h, b, h->cw[3], edge_mlp(h, B200_W_LA_EDGE, 0, 32), h->pre[4].as<float>(), 32.0f, s2);
    edge_feat<G_AL>(h, b, h->cw[2], edge_mlp(h, B200_W_LA_EDGE, 0, 32), h->pre[4].as<float>(), 32.0f, st);
    side_join(h, st);
  }
  // ---- six interaction layers
  float* hl = h->h_lig.as<float>(); float* ha = h->h_atom.as<float>();
  for (int l = 0; l < h->debug_layers; ++l) {
    NvtxRange r_layer(h, "b200dock.interaction_layer");
    const int plan = plan_of_layer(l);
    ConvLaunch L{};
    Fused16Extra X{};
    L.n = 4; L.trace = h->trace_on ? h->trace.as<long long>() : nullptr;
    L.c[0] = conv_args(h, h->cw[0], 0 * 6 + l, plan, hl, hl, 0, nullptr, 9, X, 0);     // lig
    L.c[1] = conv_args(h, h->cw[1], 1 * 6 + l, plan, ha, ha, 0, nullptr, 9, X, 1);     // atom
    L.c[2] = conv_args(h, h->cw[2], 2 * 6 + l, plan, hl, ha, 0, nullptr, 9, X, 2);     // al: target lig, gather atom
    L.c[3] = conv_args(h, h->cw[3], 3 * 6 + l, plan, ha, hl, 0, nullptr, 9, X, 3);     // la: target atom, gather lig
    if ((rc = launch_tp(h, L, X, st))) return rc;
    NodeUpdateArgs U{};
    U.plan = h->variant * B200_N_PLANS + plan;
    U.N = b.N_l; U.h = hl;
    U.src[0] = agg_src(h->cw[0]); U.ln[0] = h->convw[0 * 6 + l].ln;
    U.src[1] = agg_src(h->cw[2]); U.ln[1] = h->convw[2 * 6 + l].ln;
    cudaStream_t s2 = side_fork(h, st);                                  // ligand and pocket node updates are independent
    if (h->warp_node_update) k_node_update_warp<<<grid_for(b.N_l, 8, 148 * 8), 256, 0, st>>>(U);
    else k_node_update<<<cdiv((long long)b.N_l * h->dplans[plan].n_blocks, 128), 128, 0, st>>>(U);   // one thread per (node, irreps block)
    U.N = b.N_a; U.h = ha;
    U.src[0] = agg_src(h->cw[1]); U.ln[0] = h->convw[1 * 6 + l].ln;
    U.src[1] = agg_src(h->cw[3]); U.ln[1] = h->convw[3 * 6 + l].ln;
    if (h->warp_node_update) k_node_update_warp<<<grid_for(b.N_a, 8, 148 * 8), 256, 0, s2>>>(U);
    else k_node_update<<<cdiv((long long)b.N_a * h->dplans[plan].n_blocks, 128), 128, 0, s2>>>(U);
    side_join(h, st);
    h->launches += 2;
  }
  NvtxRange r_heads(h, "b200dock.heads");
  // ---- translation / rotation heads (side stream: independent of the pseudo-torque convs below)
  cudaStream_t s3 = side_fork(h, st);
  {
    k_centroid<<<cdiv(b.B, 64), 64, 0, s3>>>(b.lig_pos, b.lig_ptr, b.B, h->centre.as<float>());
    CenterArgs A{};
    A.lig_pos = b.lig_pos; A.lig_batch = b.lig_batch; A.N_l = b.N_l; A.centre = h->centre.as<float>();
    A.pre = h->pre[5].as<float>(); A.mlp = edge_mlp(h, B200_W_CENTER_EDGE, 0, 32); A.fc = W + off[B200_W_FINAL_FC];
    A.h_lig = hl; A.cmsg = h->cmsg.as<float>(); A.plan = h->variant * B200_N_PLANS + B200_PLAN_FINAL;
    k_center_edge<<<b.N_l, 128, 0, s3>>>(A);
    CenterHeadArgs H{};
    H.cmsg = h->cmsg.as<float>(); H.lig_ptr = b.lig_ptr; H.B = b.B; H.ln = W + off[B200_W_FINAL_LN];
    H.tr_mlp = W + off[B200_W_TR_FINAL]; H.rot_mlp = W + off[B200_W_ROT_FINAL];
    H.time_emb = c.time_emb; H.tr_sigma = c.tr_sigma; H.rot_score_norm = c.rot_score_norm; H.tr = tr; H.rot = rot;
    k_center_head<<<cdiv(b.B, 64), 64, 0, s3>>>(H);
    h->launches += 3;
  }
  // ---- pseudo-torque heads (ligand torsions, side-chain chi)
  // graph + edge features of the side-chain graph on the side stream, of the ligand-torsion graph on the caller's stream;
  // both tensor-product launches stay on the caller's stream (they fill the GPU; their event timing stays clean)
  if (h->cw[5].T) {
    if ((rc = build_graph<G_SC>(h, G, h->cw[5], s3))) return rc;
    edge_feat<G_SC>(h, b, h->cw[5], edge_mlp(h, B200_W_SC_EDGE, 0, 0), nullptr, 4.0f, s3);
  }
  if (h->cw[4].T) {
    if ((rc = build_graph<G_TOR>(h, G, h->cw[4], st))) return rc;
    edge_feat<G_TOR>(h, b, h->cw[4], edge_mlp(h, B200_W_TOR_EDGE, 0, 0), nullptr, 5.0f, st);
  }
  side_join(h, st);
  {   // both pseudo-torque convs in ONE tensor-product launch (the ligand-torsion graph alone fills a quarter of the SMs)
    ConvLaunch L{};
    Fused16Extra X{};
    L.trace = h->trace_on ? h->trace.as<long long>() : nullptr;
    for (int which = 0; which < 2; ++which) {
      ConvWs& w = h->cw[4 + which];
      if (w.T == 0) continue;
      const float* tab = which == 0 ? hl : ha;
      L.c[L.n] = conv_args(h, w, 24 + which, B200_PLAN_TOR, tab, tab, 1, which == 0 ? b.tor_bonds : b.sc_bonds, 8, X, L.n);
      L.n += 1;
    }
    if (L.n) {
      if ((rc = launch_tp(h, L, X, st))) return rc;
    }
    for (int which = 0; which < 2; ++which) {
      ConvWs& w = h->cw[4 + which];
      if (w.T == 0) continue;
      TorHeadArgs T{};
      T.n = w.T; T.plan = h->variant * B200_N_PLANS + B200_PLAN_TOR; T.src = agg_src(w); T.ln = h->convw[24 + which].ln;
      T.mlp = W + off[which == 0 ? B200_W_TOR_FINAL : B200_W_SC_FINAL];
      T.norm2 = which == 0 ? c.tor_score_norm2 : c.sc_tor_score_norm2; T.out = which == 0 ? tor : sc;
      k_tor_head<<<grid_for(w.T, 8, 148 * 8), 256, 0, st>>>(T);
      h->launches += 1;
    }
  }
  CK(cudaGetLastError());
  return B200_OK;
}

__global__ void k_fill_cond(int B, int n_tor, int n_sc, const float* __restrict__ temb, B200Step s, float* c_temb,
                            float* c_trs, float* c_rotn, float* c_torn, float* c_scn) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * SIG) c_temb[i] = temb[i % SIG];
  if (i < B) { c_trs[i] = s.tr_sigma; c_rotn[i] = s.rot_score_norm; }
  if (i < n_tor) c_torn[i] = s.tor_score_norm2;
  if (i < n_sc) c_scn[i] = s.sc_tor_score_norm2;
}

// End-of-call check: ONE stream synchronisation that brings back the capacity flag and the edge counts of the last evaluation
// (six 4-byte copies into a pinned block).  Skipped in deferred mode (b200dock_set_deferred_check): the call then returns with
// all work merely enqueued and the caller runs b200dock_check when it needs the verdict.
int check_errflag(B200Handle* h, cudaStream_t st) {
  if (!h->host_meta) CK(cudaMallocHost((void**)&h->host_meta, 8 * sizeof(int)));
  int* m = h->host_meta;
  for (int i = 0; i < 8; ++i) m[i] = 0;
  CK(cudaMemcpyAsync(&m[0], h->errflag.p, 4, cudaMemcpyDeviceToHost, st));
  const int idx[5] = {0, 1, 2, 4, 5};
  for (int i = 0; i < 5; ++i) {
    ConvWs& w = h->cw[idx[i]];
    if (w.T > 0) CK(cudaMemcpyAsync(&m[1 + i], w.seg.as<int>() + w.T + 1, 4, cudaMemcpyDeviceToHost, st));   // real edges (slots incl. padding at [T])
  }
  CK(cudaStreamSynchronize(st));
  for (int i = 0; i < 5; ++i) h->last_counts[i] = m[1 + i];
  if (m[0]) {
    char msg[128];
    snprintf(msg, sizeof msg, "edge list of graph kind %d overflowed its workspace capacity", m[0] - 1);
    FAIL(B200_ERR_CAPACITY, msg);
  }
  return B200_OK;
}

int sample_device(B200Handle* h, B200Batch& b, const B200Step* steps, int n_steps, const float* time_emb_host,
                  const float* noise, float* lig_traj, float* atom14_out, float* atom14_traj, cudaStream_t st) {
  int rc = setup_workspace(h, b);
  if (rc) return rc;
  ENS(h->temb_steps, (size_t)n_steps * SIG * 4);
  CK(cudaMemcpyAsync(h->temb_steps.p, time_emb_host, (size_t)n_steps * SIG * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(h->errflag.p, 0, 16, st));
  const size_t nstride = (size_t)6 * b.B + b.n_tor + b.n_sc;
  const int maxn = std::max(std::max(b.B * SIG, b.n_tor), b.n_sc);
  for (int s = 0; s < n_steps; ++s) {
    NvtxRange r_step(h, "b200dock.denoising_step");
    k_fill_cond<<<cdiv(maxn, 256), 256, 0, st>>>(b.B, b.n_tor, b.n_sc, h->temb_steps.as<float>() + (size_t)s * SIG, steps[s],
                                                 h->c_temb.as<float>(), h->c_trs.as<float>(), h->c_rotn.as<float>(),
                                                 h->c_torn.as<float>(), h->c_scn.as<float>());
    h->launches += 1;
    B200Cond c{h->c_temb.as<float>(), h->c_trs.as<float>(), h->c_rotn.as<float>(), h->c_torn.as<float>(), h->c_scn.as<float>()};
    rc = score_device(h, b, c, h->s_tr.as<float>(), h->s_rot.as<float>(), h->s_tor.as<float>(), h->s_sc.as<float>(), st);
    if (rc) return rc;
    const float* z = noise + (size_t)s * nstride;
    PoseArgs P{};
    P.B = b.B; P.lig_pos = b.lig_pos; P.lig_ptr = b.lig_ptr; P.tor_bonds = b.tor_bonds; P.tor_ptr = b.tor_ptr;
    P.rot_mask = b.rot_mask; P.rot_mask_off = b.rot_mask_off;
    P.tr_score = h->s_tr.as<float>(); P.rot_score = h->s_rot.as<float>(); P.tor_score = h->s_tor.as<float>();
    P.z_tr = z; P.z_rot = z + 3 * b.B; P.z_tor = z + 6 * b.B; P.st = steps[s];
    P.lig_traj_out = lig_traj ? lig_traj + (size_t)s * b.N_l * 3 : nullptr;
    cudaStream_t s2 = side_fork(h, st);           // ligand pose update and side-chain rebuild touch disjoint data
    k_lig_pose_update<<<b.B, 32, 0, st>>>(P);
    SideChainArgs S{};
    S.N_r = b.N_r; S.N_a = b.N_a; S.sequence = b.sequence; S.bb_t = b.backbone_transl; S.bb_R = b.backbone_rots;
    S.default_frame = b.default_frame; S.rigid_pos = b.rigid_group_pos; S.torsion_angle = b.torsion_angle;
    S.sc_index = b.sc_index; S.atom14_mask = b.atom14_mask; S.atom_slot = b.atom_slot;
    S.sc_score = h->s_sc.as<float>(); S.z_sc = z + 6 * b.B + b.n_tor; S.st = steps[s]; S.apply_update = 1;
    S.atom14 = (s == n_steps - 1 && atom14_out) ? atom14_out : h->atom14.as<float>();
    S.atom14_traj = atom14_traj ? atom14_traj + (size_t)s * b.N_r * 42 : nullptr;
    k_sidechain_update<<<cdiv(b.N_r, 64), 64, 0, s2>>>(S);
    k_gather_atoms<<<grid_for(b.N_a, 256, 148 * 4), 256, 0, s2>>>(S.atom14, b.atom_slot, b.N_a, b.rec_atm_pos);
    side_join(h, st);
    h->launches += 3;
  }
  CK(cudaGetLastError());
  h->last_batch = b; h->have_last = true;
  return B200_OK;
}

}  // namespace

extern "C" {

const char* b200dock_version(void) { return "b200dock 0.1 (sm_100a)"; }

const char* b200dock_last_error(const B200Handle* h) { return h ? h->err.c_str() : "null handle"; }

int b200dock_create(const B200Config* cfg, int device, B200Handle** out) {
  if (!cfg || !out) return B200_ERR_INVALID;
  B200Handle* h = new B200Handle();
  *out = h;
  h->device = device;
  h->cfg = *cfg;
  if (!kernel_known(cfg->conv_kernel)) FAIL(B200_ERR_INVALID, "unknown conv_kernel (0 exact fp32 SIMT, 5 fused tcgen05, 6 fused tcgen05 on CTA pairs, 11 pair kernel with a gather/convert warpgroup)");
  h->variant = variant_of_kernel(cfg->conv_kernel);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  h->n_sms = prop.multiProcessorCount;
  h->tp_grid = h->n_sms;
  if (const char* g = getenv("B200DOCK_NO_SIDE")) h->use_side = atoi(g) == 0;
  if (const char* g = getenv("B200DOCK_NVTX")) h->nvtx = atoi(g) != 0;
  CK(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
  CK(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
  CK(cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  if (const char* g = getenv("B200DOCK_TP_GRID")) { int v = atoi(g); if (v > 0 && v <= h->n_sms) h->tp_grid = v; }
  for (int p = 0; p < B200_N_PLANS; ++p) {
    const B200ConvPlan& s = cfg->plans[p];
    DevPlan& d = h->dplans[p];
    memset(&d, 0, sizeof d);
    d.n_paths = s.n_paths;
    for (int i = 0; i < B200_MAX_PATHS; ++i) d.paths[i] = s.paths[i];
    d.in_dim = s.in_dim; d.sh_dim = s.sh_dim; d.out_dim = s.out_dim; d.z_numel = s.z_numel; d.n_cols = s.n_cols;
    d.n_blocks = s.n_blocks;
    for (int i = 0; i < B200_MAX_BLOCKS; ++i) d.blocks[i] = s.blocks[i];
    if (p < B200_PLAN_TOR)                            // k_node_update: one thread per (node, irreps block), float4 rows
      for (int i = 0; i < s.n_blocks; ++i) {
        const B200Block& bl = s.blocks[i];
        if (!((bl.dim == 1 && bl.mul == 48) || (bl.dim == 3 && bl.mul == 12)) || (bl.off & 3))
          FAIL(B200_ERR_INVALID, "conv layer output irreps must be blocks of 48 scalars / 12 vectors at 16-byte aligned offsets");
      }
    d.n_chunks = s.n_chunks;
    if (s.n_chunks > B200_MAX_CHUNKS) FAIL(B200_ERR_INVALID, "too many weight chunks in a conv plan");
    for (int i = 0; i < s.n_chunks; ++i) { d.chunk_col[i] = s.chunk_col[i]; d.chunk_n[i] = s.chunk_n[i]; d.chunk_path[i] = s.chunk_path[i]; }
  }
  {
    std::vector<float>& dense = h->cg_dense;
    dense.assign((size_t)B200_N_PLANS * B200_MAX_PATHS * 45, 0.0f);
    for (int p = 0; p < B200_N_PLANS; ++p) {
      const B200ConvPlan& s = cfg->plans[p];
      for (int q = 0; q < s.n_paths; ++q) {
        const B200Path& pa = s.paths[q];
        for (int c = pa.cg_off; c < pa.cg_off + pa.cg_n; ++c) {
          int ijk = s.cg_ijk[c];
          int i = ijk & 255, j = (ijk >> 8) & 255, k = (ijk >> 16) & 255;
          if (i < 3 && j < 5 && k < 3) dense[((size_t)p * B200_MAX_PATHS + q) * 45 + (i * 5 + j) * 3 + k] = s.cg_val[c];
        }
      }
    }
  }
  memcpy(h->atom14_group, cfg->atom14_group, sizeof(int) * 21 * 14);
  int rc;
  if ((rc = bind_plans(h))) return rc;
  if ((rc = upload(h, cfg->tor_cg_ijk, (size_t)cfg->tor_cg_off[3], &h->d_tor_cg_ijk))) return rc;
  if ((rc = upload(h, cfg->tor_cg_val, (size_t)cfg->tor_cg_off[3], &h->d_tor_cg_val))) return rc;
  CK(cudaFuncSetAttribute(k_conv_prologue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PRO_SMEM));
  CK(cudaFuncSetAttribute(k_conv_tp_simt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM));
  if (tc_init() || conv_fused_init() || conv_fused2_init()) FAIL(B200_ERR_CUDA, "tcgen05 conv kernel attribute setup failed");
  h->cfg.atom14_group = nullptr; h->cfg.tor_cg_ijk = nullptr; h->cfg.tor_cg_val = nullptr;
  return B200_OK;
}

void b200dock_destroy(B200Handle* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  auto fr = [](Buf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; };
  for (auto& w : h->cw) { fr(w.counts); fr(w.seg); fr(w.gpad); fr(w.es); fr(w.ed); fr(w.eaux); fr(w.emb); fr(w.sh); fr(w.H1); fr(w.Zt); fr(w.msg); fr(w.agg); fr(w.part); }
  for (auto& b : h->pre) fr(b);
  Buf* all[] = {&h->h_lig, &h->h_atom, &h->jmax_lig, &h->jmax_atom, &h->centre, &h->cmsg, &h->s_tr, &h->s_rot, &h->s_tor,
                &h->s_sc, &h->atom14, &h->errflag, &h->c_temb, &h->c_trs, &h->c_rotn, &h->c_torn, &h->c_scn,
                &h->temb_steps, &h->mdn_w, &h->mdn_A, &h->mdn_B, &h->enc_w, &h->enc_ws, &h->trace, &h->dev_in, &h->dev_noise, &h->dev_lig_out, &h->dev_a14_out};
  for (Buf* b : all) fr(*b);
  if (h->pinned_in.p) cudaFreeHost(h->pinned_in.p);
  if (h->pinned_out.p) cudaFreeHost(h->pinned_out.p);
  if (h->host_meta) cudaFreeHost(h->host_meta);
  if (h->ex_pin.p) cudaFreeHost(h->ex_pin.p);
  if (h->ex_in.p) cudaFree(h->ex_in.p);
  for (auto& b : h->ex_out) if (b.p) cudaFree(b.p);
  for (void* p : h->plan_allocs) cudaFree(p);
  if (h->d_blob) cudaFree(h->d_blob);
  if (h->d_w16) cudaFree(h->d_w16);
  for (auto e : h->event_pool) cudaEventDestroy(e);
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  delete h;
}

int b200dock_load_weights(B200Handle* h, const float* blob, size_t n, const int64_t* offsets, int n_sections) {
  if (!h || !blob || !offsets) return B200_ERR_INVALID;
  if (n_sections != B200_W_N_SECTIONS) FAIL(B200_ERR_INVALID, "unexpected number of weight sections");
  CK(cudaSetDevice(h->device));
  if (h->d_blob) CK(cudaFree(h->d_blob));
  CK(cudaMalloc((void**)&h->d_blob, n * sizeof(float) + 1024));
  CK(cudaMemcpy(h->d_blob, blob, n * sizeof(float), cudaMemcpyHostToDevice));
  h->blob_n = n;
  h->off.assign(offsets, offsets + n_sections);
  for (int i = 0; i < 26; ++i) {
    int plan = i < 24 ? plan_of_layer(i % 6) : B200_PLAN_TOR;
    const B200ConvPlan& P = h->cfg.plans[plan];
    int nirr = 0, nsc = 0;
    for (int b = 0; b < P.n_blocks; ++b) { nirr += P.blocks[b].mul; if (P.blocks[b].bias_off >= 0) nsc += P.blocks[b].mul; }
    const float* r = h->d_blob + h->off[B200_W_CONV0 + i];
    ConvW& w = h->convw[i];
    w.W1t = r; r += 144 * 144;
    w.b1 = r; r += 144;
    w.W2p = r; r += (size_t)P.n_cols * KP;
    w.ln.shift = r; r += nirr;
    w.ln.weight = r; r += nirr;
    w.ln.bias = r; r += nsc;
    if ((size_t)(r - h->d_blob) > n) FAIL(B200_ERR_INVALID, "weight blob too small for its section table");
    if (((uintptr_t)w.W2p & 15) != 0) FAIL(B200_ERR_INVALID, "conv record is not 16-byte aligned");
    w.n_cols = P.n_cols;
  }
  if (h->cfg.conv_kernel >= 5) {   // fp16 hi/lo copies of W1p / W2p with exact power-of-two scales
    size_t tot = 0;
    for (int i = 0; i < 26; ++i) tot += ((size_t)h->convw[i].n_cols + 144 + 192) * KH;
    if (h->d_w16) CK(cudaFree(h->d_w16));
    CK(cudaMalloc((void**)&h->d_w16, 2 * tot * sizeof(__half) + 1024));
    float* tmp = nullptr; float* d_max = nullptr;
    CK(cudaMalloc((void**)&tmp, (size_t)192 * KP * sizeof(float)));
    CK(cudaMalloc((void**)&d_max, 2 * sizeof(float)));
    size_t pos = 0;
    for (int i = 0; i < 26; ++i) {
      ConvW& w = h->convw[i];
      const size_t n2 = (size_t)w.n_cols * KP, r2 = (size_t)w.n_cols + 144;
      k_build_w1p_f32<<<120, 256>>>(w.W1t, w.b1, tmp);
      CK(cudaMemset(d_max, 0, 2 * sizeof(float)));
      k_absmax<<<148, 256>>>(tmp, (size_t)192 * KP, d_max);
      k_absmax<<<148 * 4, 256>>>(w.W2p, n2, d_max + 1);
      float mx[2];
      CK(cudaMemcpy(mx, d_max, sizeof mx, cudaMemcpyDeviceToHost));
      auto scale_of = [](float m) { int ex; frexpf(m > 0 ? m : 1.0f, &ex); return ldexpf(1.0f, 10 - ex); };   // max -> [2^9, 2^10)
      const float s1 = scale_of(mx[0]), s2 = scale_of(mx[1]);
      __half* h1 = h->d_w16 + pos; __half* l1 = h1 + (size_t)192 * KH;
      __half* h2 = l1 + (size_t)192 * KH; __half* l2 = h2 + r2 * KH;
      k_build_w16<<<148, 256>>>(tmp, 192, 192, s1, h1, l1);
      k_build_w16<<<148 * 4, 256>>>(w.W2p, w.n_cols, (int)r2, s2, h2, l2);
      w.W1h16 = h1; w.W1l16 = l1; w.W2h16 = h2; w.W2l16 = l2; w.inv_s1 = 1.0f / s1; w.inv_s2 = 1.0f / s2;
      pos += 2 * (size_t)192 * KH + 2 * r2 * KH;
    }
    CK(cudaDeviceSynchronize());
    CK(cudaFree(tmp)); CK(cudaFree(d_max));
  }
  h->weights = true;
  return B200_OK;
}

int b200dock_score(B200Handle* h, const B200Batch* batch, const B200Cond* cond, float* tr, float* rot, float* tor,
                   float* sc, void* stream) {
  if (!h || !batch || !cond) return B200_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  h->launches = 0;
  int rc = setup_workspace(h, *batch);
  if (rc) return rc;
  CK(cudaMemsetAsync(h->errflag.p, 0, 16, st));
  rc = score_device(h, *batch, *cond, tr, rot, tor ? tor : h->s_tor.as<float>(), sc ? sc : h->s_sc.as<float>(), st);
  if (rc) return rc;
  h->last_batch = *batch; h->have_last = true;
  return h->deferred_check ? B200_OK : check_errflag(h, st);
}

int b200dock_sample(B200Handle* h, B200Batch* batch, const B200Step* steps, int n_steps, const float* time_emb,
                    const float* noise, float* lig_traj, float* atom14_out, float* atom14_traj, void* stream) {
  if (!h || !batch || !steps || !time_emb || !noise || n_steps <= 0) return B200_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  h->launches = 0;
  int rc = sample_device(h, *batch, steps, n_steps, time_emb, noise, lig_traj, atom14_out, atom14_traj, st);
  if (rc) return rc;
  return h->deferred_check ? B200_OK : check_errflag(h, st);
}

int b200dock_set_deferred_check(B200Handle* h, int on) {
  if (!h) return B200_ERR_INVALID;
  h->deferred_check = on != 0;
  return B200_OK;
}

int b200dock_check(B200Handle* h, void* stream) {
  if (!h) return B200_ERR_INVALID;
  if (!h->errflag.p) return B200_OK;
  CK(cudaSetDevice(h->device));
  return check_errflag(h, (cudaStream_t)stream);
}

int b200dock_sample_host(B200Handle* h, const B200Batch* hb, const B200Step* steps, int n_steps, const float* time_emb,
                         const float* noise, float* lig_out, float* atom14_out, uint64_t* h2d_bytes,
                         uint64_t* d2h_bytes, void* stream) {
  if (!h || !hb || !steps || !time_emb || !noise || !lig_out || !atom14_out || n_steps <= 0) return B200_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  h->launches = 0;
  const B200Batch& b = *hb;
  // ---- lay all inputs out in one pinned arena, one H2D copy
  struct Item { const void* src; size_t bytes; size_t off; };
  std::vector<Item> items;
  size_t total = 0;
  auto add = [&](const void* p, size_t bytes) { size_t o = total; items.push_back({p, bytes, o}); total += (bytes + 255) & ~(size_t)255; return o; };
  const size_t nstride = (size_t)6 * b.B + b.n_tor + b.n_sc;
  size_t o_lig_node = add(b.lig_node, (size_t)b.N_l * 27 * 4), o_lig_pos = add(b.lig_pos, (size_t)b.N_l * 12),
         o_lig_ptr = add(b.lig_ptr, (size_t)(b.B + 1) * 4), o_lig_batch = add(b.lig_batch, (size_t)b.N_l * 4),
         o_bond_ptr = add(b.bond_ptr, (size_t)(b.N_l + 1) * 4), o_bond_dst = add(b.bond_dst, (size_t)b.E_b * 4),
         o_bond_eid = add(b.bond_eid, (size_t)b.E_b * 4), o_ef = add(b.lig_edge_feat, (size_t)b.E_b * 40),
         o_tb = add(b.tor_bonds, (size_t)b.n_tor * 8), o_tp = add(b.tor_ptr, (size_t)(b.B + 1) * 4),
         o_rm = add(b.rot_mask, (size_t)b.rot_mask_bytes), o_rmo = add(b.rot_mask_off, (size_t)b.n_tor * 8),
         o_pf = add(b.pocket_feat, (size_t)b.N_a * 20), o_rp = add(b.rec_atm_pos, (size_t)b.N_a * 12),
         o_ap = add(b.atom_ptr, (size_t)(b.B + 1) * 4), o_ab = add(b.atom_batch, (size_t)b.N_a * 4),
         o_as = add(b.atom_slot, (size_t)b.N_a * 4), o_rptr = add(b.res_ptr, (size_t)(b.B + 1) * 4),
         o_m14 = add(b.atom14_mask, (size_t)b.N_r * 14), o_seq = add(b.sequence, (size_t)b.N_r * 4),
         o_bt = add(b.backbone_transl, (size_t)b.N_r * 12), o_bR = add(b.backbone_rots, (size_t)b.N_r * 36),
         o_df = add(b.default_frame, (size_t)b.N_r * 8 * 64), o_rg = add(b.rigid_group_pos, (size_t)b.N_r * 14 * 12),
         o_ta = add(b.torsion_angle, (size_t)b.N_r * 20), o_scb = add(b.sc_bonds, (size_t)b.n_sc * 8),
         o_sci = add(b.sc_index, (size_t)b.N_r * 16), o_scp = add(b.sc_ptr, (size_t)(b.B + 1) * 4), o_noise = add(noise, nstride * n_steps * 4);
  if (total > h->pinned_in.cap) {
    if (h->pinned_in.p) CK(cudaFreeHost(h->pinned_in.p));
    h->pinned_in.p = nullptr; h->pinned_in.cap = 0;
    CK(cudaMallocHost(&h->pinned_in.p, total + total / 8));
    h->pinned_in.cap = total + total / 8;
  }
  ENS(h->dev_in, total);
  char* pin = (char*)h->pinned_in.p;
  for (const Item& it : items) if (it.bytes) memcpy(pin + it.off, it.src, it.bytes);
  CK(cudaMemcpyAsync(h->dev_in.p, pin, total, cudaMemcpyHostToDevice, st));
  char* d = (char*)h->dev_in.p;
  B200Batch db = b;
  db.lig_node = (const float*)(d + o_lig_node); db.lig_pos = (float*)(d + o_lig_pos);
  db.lig_ptr = (const int*)(d + o_lig_ptr); db.lig_batch = (const int*)(d + o_lig_batch);
  db.bond_ptr = (const int*)(d + o_bond_ptr); db.bond_dst = (const int*)(d + o_bond_dst);
  db.bond_eid = (const int*)(d + o_bond_eid); db.lig_edge_feat = (const float*)(d + o_ef);
  db.tor_bonds = (const int*)(d + o_tb); db.tor_ptr = (const int*)(d + o_tp);
  db.rot_mask = (const uint8_t*)(d + o_rm); db.rot_mask_off = (const int64_t*)(d + o_rmo);
  db.pocket_feat = (const int*)(d + o_pf); db.rec_atm_pos = (float*)(d + o_rp);
  db.atom_ptr = (const int*)(d + o_ap); db.atom_batch = (const int*)(d + o_ab); db.atom_slot = (const int*)(d + o_as);
  db.res_ptr = (const int*)(d + o_rptr); db.atom14_mask = (const uint8_t*)(d + o_m14); db.sequence = (const int*)(d + o_seq);
  db.backbone_transl = (const float*)(d + o_bt); db.backbone_rots = (const float*)(d + o_bR);
  db.default_frame = (const float*)(d + o_df); db.rigid_group_pos = (const float*)(d + o_rg);
  db.torsion_angle = (float*)(d + o_ta); db.sc_bonds = (const int*)(d + o_scb); db.sc_index = (const int*)(d + o_sci); db.sc_ptr = (const int*)(d + o_scp);
  ENS(h->dev_a14_out, (size_t)b.N_r * 42 * 4);
  int rc = sample_device(h, db, steps, n_steps, time_emb, (const float*)(d + o_noise), nullptr,
                         h->dev_a14_out.as<float>(), nullptr, st);
  if (rc) return rc;
  const size_t out_bytes = (size_t)b.N_l * 12 + (size_t)b.N_r * 42 * 4;
  if (out_bytes > h->pinned_out.cap) {
    if (h->pinned_out.p) CK(cudaFreeHost(h->pinned_out.p));
    h->pinned_out.p = nullptr; h->pinned_out.cap = 0;
    CK(cudaMallocHost(&h->pinned_out.p, out_bytes + out_bytes / 8));
    h->pinned_out.cap = out_bytes + out_bytes / 8;
  }
  char* po = (char*)h->pinned_out.p;
  CK(cudaMemcpyAsync(po, db.lig_pos, (size_t)b.N_l * 12, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(po + (size_t)b.N_l * 12, h->dev_a14_out.p, (size_t)b.N_r * 42 * 4, cudaMemcpyDeviceToHost, st));
  rc = check_errflag(h, st);   // synchronises the stream
  if (rc) return rc;
  memcpy(lig_out, po, (size_t)b.N_l * 12);
  memcpy(atom14_out, po + (size_t)b.N_l * 12, (size_t)b.N_r * 42 * 4);
  if (h2d_bytes) *h2d_bytes = total + (uint64_t)n_steps * SIG * 4;
  if (d2h_bytes) *d2h_bytes = out_bytes + 4;
  return B200_OK;
}

int b200dock_mdn_load_weights(B200Handle* h, const float* blob, size_t n) {
  if (!h || !blob) return B200_ERR_INVALID;
  const size_t want = (size_t)MDN_H * MDN_H * 2 + MDN_H + (size_t)MDN_H * 30 + 30;
  if (n != want) FAIL(B200_ERR_INVALID, "unexpected MDN weight blob size");
  CK(cudaSetDevice(h->device));
  ENS(h->mdn_w, n * sizeof(float));
  CK(cudaMemcpy(h->mdn_w.p, blob, n * sizeof(float), cudaMemcpyHostToDevice));
  h->mdn_weights = true;
  return B200_OK;
}

int b200dock_mdn_score(B200Handle* h, const B200MdnBatch* mb, float dist_threshold, float* score, void* stream) {
  if (!h || !mb || !score) return B200_ERR_INVALID;
  if (!h->mdn_weights) FAIL(B200_ERR_STATE, "MDN weights not loaded");
  if (mb->B <= 0 || mb->N_l <= 0 || mb->N_r <= 0) FAIL(B200_ERR_INVALID, "empty MDN batch");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  ENS(h->mdn_A, (size_t)mb->N_l * MDN_H * 4); ENS(h->mdn_B, (size_t)mb->N_r * MDN_H * 4);
  const float* W = h->mdn_w.as<float>();
  const float* Wl = W; const float* Wr = W + MDN_H * MDN_H; const float* bias = Wr + MDN_H * MDN_H;
  const float* W30 = bias + MDN_H; const float* b30 = W30 + MDN_H * 30;
  k_mdn_project<<<grid_for(mb->N_l, 1, 148 * 8), MDN_H, 0, st>>>(mb->lig_s, mb->N_l, Wl, nullptr, h->mdn_A.as<float>());
  k_mdn_project<<<grid_for(mb->N_r, 1, 148 * 8), MDN_H, 0, st>>>(mb->pro_s, mb->N_r, Wr, bias, h->mdn_B.as<float>());
  MdnArgs M{};
  M.B = mb->B; M.A = h->mdn_A.as<float>(); M.Bm = h->mdn_B.as<float>(); M.lig_pos = mb->lig_pos; M.lig_ptr = mb->lig_ptr;
  M.xyz_full = mb->xyz_full; M.res_ptr = mb->res_ptr; M.W30t = W30; M.b30 = b30; M.thr = dist_threshold; M.score = score;
  k_mdn_pairs<<<grid_for(mb->B, 1, 148 * 4), 256, 0, st>>>(M);
  h->launches += 3;
  CK(cudaGetLastError());
  return B200_OK;
}

int b200dock_mdn_load_encoder_weights(B200Handle* h, const float* blob, size_t n, const int64_t* offsets, int n_offsets) {
  if (!h || !blob || !offsets) return B200_ERR_INVALID;
  if (n_offsets != B200_MDN_ENC_SECTIONS) FAIL(B200_ERR_INVALID, "unexpected number of MDN encoder weight sections");
  for (int i = 0; i < n_offsets; ++i)
    if (offsets[i] < -1 || offsets[i] >= (int64_t)n) FAIL(B200_ERR_INVALID, "MDN encoder weight offset out of range");
  CK(cudaSetDevice(h->device));
  ENS(h->enc_w, n * sizeof(float));
  CK(cudaMemcpy(h->enc_w.p, blob, n * sizeof(float), cudaMemcpyHostToDevice));
  h->enc_off.assign(offsets, offsets + n_offsets);
  h->enc_weights = true;
  return B200_OK;
}

namespace {
struct EncCtx {
  B200Handle* h; cudaStream_t st; const float* W;
  const float* sec(int i) const { return h->enc_off[i] < 0 ? nullptr : W + h->enc_off[i]; }
  void linear(const float* in, int ld_in, int rows, int K, int O, int wsec, int bsec, int act, const float* res, float* out) const {
    LinArgs A{in, ld_in, rows, K, O, sec(wsec), bsec >= 0 ? sec(bsec) : nullptr, act, res, O, out, O};
    k_enc_linear<<<grid_for(rows, ENC_ROWS, 148 * 8), ENC_THREADS, (size_t)ENC_ROWS * K * 4, st>>>(A);
    h->launches++;
  }
  // GVP with weight sections [base, base+4); segments given by the caller
  void gvp(GvpArgs A, int base, int scalar_act, int vector_act) const {
    A.h = A.vi > A.vo ? A.vi : A.vo;
    A.wh = sec(base); A.ws_t = sec(base + 1); A.ws_b = sec(base + 2); A.wv = sec(base + 3);
    A.scalar_act = scalar_act; A.vector_act = vector_act;
    const size_t sm = (size_t)ENC_ROWS * (A.si + A.h + 3 * A.vi + 3 * A.h) * 4;
    k_gvp<<<grid_for(A.rows, ENC_ROWS, 148 * 8), ENC_THREADS, sm, st>>>(A);
    h->launches++;
  }
  void ln(GvpLnArgs A, int base) const {
    A.w = sec(base); A.b = sec(base + 1);
    k_gvp_ln<<<grid_for(A.rows, 4, 148 * 8), 128, 0, st>>>(A);
    h->launches++;
  }
};
GvpArgs gvp_plain(int rows, int si, int vi, int so, int vo, const float* s, const float* v, float* os, float* ov) {
  GvpArgs A{};
  A.rows = rows; A.si = si; A.vi = vi; A.so = so; A.vo = vo;
  A.s[0] = GvpSeg{s, nullptr, si}; A.v[0] = GvpSeg{v, nullptr, vi};
  A.out_s = os; A.out_v = ov;
  return A;
}
}  // namespace

int b200dock_mdn_encode(B200Handle* h, const B200MdnGraph* g, float* pro_s, float* lig_s, void* stream) {
  if (!h || !g || !pro_s || !lig_s) return B200_ERR_INVALID;
  if (!h->enc_weights) FAIL(B200_ERR_STATE, "MDN encoder weights not loaded");
  if (g->N_r <= 0 || g->N_l <= 0 || g->E_p < 0 || g->E_l < 0) FAIL(B200_ERR_INVALID, "empty MDN graph");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const size_t Nl = g->N_l, El = g->E_l > 0 ? g->E_l : 1, Nr = g->N_r, Ep = g->E_p > 0 ? g->E_p : 1;
  // workspace carve-up (floats)
  size_t need = 0;
  auto take = [&](size_t n) { size_t o = need; need += (n + 63) & ~(size_t)63; return o; };
  // ligand
  const size_t o_x = take(Nl * 128), o_x2 = take(Nl * 128), o_qkv = take(Nl * 384), o_hn = take(Nl * 128), o_t256 = take(Nl * 256);
  const size_t o_e = take(El * 128), o_e2 = take(El * 128), o_ep = take(El * 128), o_al = take(El * 128), o_ax = take(El * 4), o_et = take(El * 256);
  // pocket
  const size_t o_s0 = take(Nr * 40), o_v0 = take(Nr * 9), o_s = take(Nr * 128), o_v = take(Nr * 48), o_sb = take(Nr * 128), o_vb = take(Nr * 48);
  const size_t o_ds = take(Nr * 128), o_dv = take(Nr * 48), o_fs = take(Nr * 512), o_fv = take(Nr * 96);
  const size_t o_es0 = take(Ep * 21), o_ev0 = take(Ep * 3), o_es = take(Ep * 32), o_ev = take(Ep * 3);
  const size_t o_ms = take(Ep * 128), o_mv = take(Ep * 48), o_ms2 = take(Ep * 128), o_mv2 = take(Ep * 48);
  ENS(h->enc_ws, need * 4);
  float* ws = h->enc_ws.as<float>();
  EncCtx C{h, st, h->enc_w.as<float>()};

  // ======================= ligand graph transformer (GraphTransformer_Block.py:413-424)
  float *x = ws + o_x, *x2 = ws + o_x2, *e = ws + o_e, *e2 = ws + o_e2;
  C.linear(g->lig_node_s, 89, g->N_l, 89, 128, 0, 1, 0, nullptr, x);
  if (g->E_l > 0) C.linear(g->lig_edge_s, 20, g->E_l, 20, 128, 2, 3, 0, nullptr, e);
  for (int l = 0; l < 6; ++l) {
    const int b = 4 + 14 * l;
    const bool fin = l == 5;
    C.linear(x, 128, g->N_l, 128, 384, b, b + 1, 0, nullptr, ws + o_qkv);                       // BN1 folded: Q | K | V
    if (g->E_l > 0) {
      C.linear(e, 128, g->E_l, 128, 128, b + 2, b + 3, 0, nullptr, ws + o_ep);                   // BN1 folded: edge projection
      k_gt_edge<<<grid_for(g->E_l, 8, 148 * 8), 256, 0, st>>>(ws + o_qkv, ws + o_ep, g->lig_row, g->lig_col, g->E_l, ws + o_al, ws + o_ax);
      h->launches++;
    }
    k_gt_node<<<grid_for(g->N_l, 1, 148 * 8), 128, 0, st>>>(ws + o_qkv, ws + o_ax, g->lig_row, g->lig_perm, g->lig_ptr, g->N_l, ws + o_hn);
    h->launches++;
    C.linear(ws + o_hn, 128, g->N_l, 128, 128, b + 4, b + 5, 0, x, x2);                          // x2 = x + O_node(h)
    C.linear(x2, 128, g->N_l, 128, 256, b + 6, b + 7, 1, nullptr, ws + o_t256);                  // BN2 folded, SiLU
    C.linear(ws + o_t256, 256, g->N_l, 256, 128, b + 8, -1, 0, x2, fin ? lig_s : x);             // x = x2 + MLP
    if (!fin && g->E_l > 0) {
      C.linear(ws + o_al, 128, g->E_l, 128, 128, b + 9, b + 10, 0, e, e2);
      C.linear(e2, 128, g->E_l, 128, 256, b + 11, b + 12, 1, nullptr, ws + o_et);
      C.linear(ws + o_et, 256, g->E_l, 256, 128, b + 13, -1, 0, e2, e);
    }
  }

  // ======================= pocket GVP embedding (GVP_Block.py:63-79)
  float *s = ws + o_s, *v = ws + o_v, *sb = ws + o_sb, *vb = ws + o_vb;
  {
    GvpLnArgs L{}; L.rows = g->N_r; L.ns = 40; L.nv = 3; L.s = g->pro_node_s; L.v = g->pro_node_v; L.emb = C.sec(88); L.seq = g->pro_seq; L.ns0 = 9;
    L.out_s = ws + o_s0; L.out_v = ws + o_v0;
    C.ln(L, 89);
    C.gvp(gvp_plain(g->N_r, 40, 3, 128, 16, ws + o_s0, ws + o_v0, s, v), 91, 0, 0);
  }
  if (g->E_p > 0) {
    GvpLnArgs L{}; L.rows = g->E_p; L.ns = 21; L.nv = 1; L.s = g->pro_edge_s; L.v = g->pro_edge_v; L.out_s = ws + o_es0; L.out_v = ws + o_ev0;
    C.ln(L, 95);
    C.gvp(gvp_plain(g->E_p, 21, 1, 32, 1, ws + o_es0, ws + o_ev0, ws + o_es, ws + o_ev), 97, 0, 0);
  }
  for (int l = 0; l < 3; ++l) {
    const int b = 101 + 24 * l;
    if (g->E_p > 0) {
      GvpArgs M{};
      M.rows = g->E_p; M.si = 288; M.vi = 33; M.so = 128; M.vo = 16;
      M.s[0] = GvpSeg{s, g->pro_src, 128}; M.s[1] = GvpSeg{ws + o_es, nullptr, 32}; M.s[2] = GvpSeg{s, g->pro_dst, 128};
      M.v[0] = GvpSeg{v, g->pro_src, 16};  M.v[1] = GvpSeg{ws + o_ev, nullptr, 1};  M.v[2] = GvpSeg{v, g->pro_dst, 16};
      M.out_s = ws + o_ms; M.out_v = ws + o_mv;
      C.gvp(M, b, 1, 1);
      C.gvp(gvp_plain(g->E_p, 128, 16, 128, 16, ws + o_ms, ws + o_mv, ws + o_ms2, ws + o_mv2), b + 4, 1, 1);
      C.gvp(gvp_plain(g->E_p, 128, 16, 128, 16, ws + o_ms2, ws + o_mv2, ws + o_ms, ws + o_mv), b + 8, 0, 0);
    }
    k_seg_mean<<<grid_for(g->N_r, 1, 148 * 8), 128, 0, st>>>(ws + o_ms, 128, g->pro_perm, g->pro_ptr, g->N_r, ws + o_ds);
    k_seg_mean<<<grid_for(g->N_r, 1, 148 * 8), 64, 0, st>>>(ws + o_mv, 48, g->pro_perm, g->pro_ptr, g->N_r, ws + o_dv);
    h->launches += 2;
    { GvpLnArgs L{}; L.rows = g->N_r; L.ns = 128; L.nv = 16; L.s = s; L.v = v; L.ds = ws + o_ds; L.dv = ws + o_dv; L.out_s = sb; L.out_v = vb; C.ln(L, b + 12); }
    C.gvp(gvp_plain(g->N_r, 128, 16, 512, 32, sb, vb, ws + o_fs, ws + o_fv), b + 14, 1, 1);
    C.gvp(gvp_plain(g->N_r, 512, 32, 128, 16, ws + o_fs, ws + o_fv, ws + o_ds, ws + o_dv), b + 18, 0, 0);
    { GvpLnArgs L{}; L.rows = g->N_r; L.ns = 128; L.nv = 16; L.s = sb; L.v = vb; L.ds = ws + o_ds; L.dv = ws + o_dv; L.out_s = s; L.out_v = v; C.ln(L, b + 22); }
  }
  { GvpLnArgs L{}; L.rows = g->N_r; L.ns = 128; L.nv = 16; L.s = s; L.v = v; L.out_s = sb; L.out_v = vb; C.ln(L, 173); }
  C.gvp(gvp_plain(g->N_r, 128, 16, 128, 0, sb, vb, pro_s, nullptr), 175, 1, 0);
  CK(cudaGetLastError());
  return B200_OK;
}

int b200dock_mdn_featurize(B200Handle* h, const B200MdnFeat* f, void* stream) {
  if (!h || !f) return B200_ERR_INVALID;
  if (f->B <= 0 || f->N_r <= 0 || f->topk <= 0 || f->topk > 32 || f->max_res <= 0) FAIL(B200_ERR_INVALID, "bad MDN featuriser sizes (topk must be 1..32)");
  const size_t smem = (size_t)f->max_res * (3 * 8 + 6 * 4) + 8 * 32 * 4;
  if (smem > 220 * 1024) FAIL(B200_ERR_INVALID, "pocket too large for the MDN featuriser (more than ~4600 residues)");
  CK(cudaSetDevice(h->device));
  if (smem > h->feat_smem) {
    CK(cudaFuncSetAttribute(k_mdn_featurize, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->feat_smem = smem;
  }
  MdnFeatArgs A{};
  A.B = f->B; A.topk = f->topk; A.res_ptr = f->res_ptr; A.edge_ptr = (const long long*)f->edge_ptr;
  A.atom14 = f->atom14; A.mask = f->atom14_mask; A.bb_sincos = f->bb_sincos;
  A.node_s = f->node_s; A.node_v = f->node_v; A.edge_src = f->edge_src; A.edge_dst = f->edge_dst;
  A.edge_s = f->edge_s; A.edge_v = f->edge_v; A.node_ptr = f->node_ptr;
  k_mdn_featurize<<<f->B, 256, smem, (cudaStream_t)stream>>>(A);
  h->launches += 1;
  CK(cudaGetLastError());
  return B200_OK;
}

int b200dock_vina(B200Handle* h, const B200Vina* v, void* stream) {
  if (!h || !v) return B200_ERR_INVALID;
  if (v->n_pose <= 0 || v->n_lig <= 0 || v->n_lig > VINA_MAX_LIG || v->n_tors < 0 || v->n_tors > VINA_MAX_TORS || v->n_rec <= 0 ||
      v->n_rec > 10000 || v->root < 0 || v->root >= v->n_lig)
    FAIL(B200_ERR_INVALID, "bad sizes for the error-correction stage (n_lig <= 128, n_tors <= 58, n_rec <= 10000)");
  if (!v->lig_xyz || !v->lig_radius || !v->lig_flags || !v->rec_xyz || !v->rec_radius || !v->rec_flags || !v->pair_ptr || !v->out_energy ||
      (v->n_tors && (!v->tors_axis || !v->tors_mask)) || (v->mode == 1 && !v->out_xyz))
    FAIL(B200_ERR_INVALID, "null pointer in B200Vina");
  CK(cudaSetDevice(h->device));
  const size_t smem = vina_smem_bytes(v->n_rec);
  if (smem > h->vina_smem) {
    CK(cudaFuncSetAttribute(k_vina, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    h->vina_smem = smem;
  }
  VinaArgs A{};
  A.n_pose = v->n_pose; A.n_lig = v->n_lig; A.n_rec = v->n_rec; A.n_tors = v->n_tors; A.root = v->root; A.max_steps = v->max_steps;
  A.mode = v->mode; A.rec_pose_stride = v->rec_pose_stride;
  A.lig_xyz = v->lig_xyz; A.lig_R = v->lig_radius; A.lig_flags = v->lig_flags;
  A.rec_xyz = v->rec_xyz; A.rec_R = v->rec_radius; A.rec_flags = v->rec_flags;
  A.tors_axis = v->tors_axis; A.tors_mask = v->tors_mask; A.pair_ptr = v->pair_ptr; A.pair_idx = v->pair_idx; A.n_rot = v->n_rot;
  A.out_xyz = v->out_xyz; A.out_energy = v->out_energy; A.out_terms = v->out_terms; A.out_stats = v->out_stats;
  k_vina<<<v->n_pose, VINA_THREADS, smem, (cudaStream_t)stream>>>(A);
  h->launches += 1;
  CK(cudaGetLastError());
  return B200_OK;
}

int b200dock_expand_host(B200Handle* h, const B200Batch* base, const B200Expand* ex, B200Batch* out, void* stream) {
  if (!h || !base || !ex || !out || !ex->src_graph || ex->B_out <= 0) return B200_ERR_INVALID;
  if (ex->randomize && !ex->stream_id) FAIL(B200_ERR_INVALID, "randomize needs one RNG stream id per output graph");
  CK(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  const B200Batch& b = *base;
  const int Bo = ex->B_out;
  for (int g = 0; g < Bo; ++g)
    if (ex->src_graph[g] < 0 || ex->src_graph[g] >= b.B) FAIL(B200_ERR_INVALID, "src_graph out of range");
  // ---- per-entity offsets of the base graphs (host arrays are readable here)
  enum { EL, EE, ET, EM, EA, ER, ER14, ES, NENT };
  std::vector<long long> off[NENT];
  for (auto& v : off) v.assign(b.B + 1, 0);
  for (int g = 0; g <= b.B; ++g) {
    off[EL][g] = b.lig_ptr[g]; off[EE][g] = b.bond_ptr[b.lig_ptr[g]]; off[ET][g] = b.tor_ptr[g];
    off[EA][g] = b.atom_ptr[g]; off[ER][g] = b.res_ptr[g]; off[ER14][g] = 14LL * b.res_ptr[g]; off[ES][g] = b.sc_ptr[g];
  }
  for (int g = 0; g < b.B; ++g)
    off[EM][g + 1] = off[EM][g] + (long long)(b.tor_ptr[g + 1] - b.tor_ptr[g]) * (b.lig_ptr[g + 1] - b.lig_ptr[g]);
  // ---- small per-output-graph tables: new_ptr [Bo+1], old_start [Bo], delta [Bo] per entity, + int64 delta of the mask offsets
  const size_t per = (size_t)(Bo + 1) + 2 * (size_t)Bo;
  std::vector<int> tab(NENT * per);
  std::vector<long long> d64(Bo);
  long long tot[NENT];
  for (int e = 0; e < NENT; ++e) {
    int* np = tab.data() + e * per; int* os = np + Bo + 1; int* dl = os + Bo;
    long long run = 0;
    for (int g = 0; g < Bo; ++g) {
      const int sg = ex->src_graph[g];
      np[g] = (int)run; os[g] = (int)off[e][sg]; dl[g] = (int)(run - off[e][sg]);
      if (e == EM) d64[g] = run - off[e][sg];
      run += off[e][sg + 1] - off[e][sg];
    }
    np[Bo] = (int)run; tot[e] = run;
    if (run > 0x7fffffffLL) FAIL(B200_ERR_INVALID, "expanded batch too large for 32-bit indices");
  }
  B200Batch o{};
  o.B = Bo; o.N_l = (int)tot[EL]; o.E_b = (int)tot[EE]; o.n_tor = (int)tot[ET]; o.N_a = (int)tot[EA]; o.N_r = (int)tot[ER]; o.n_sc = (int)tot[ES];
  o.rot_mask_bytes = tot[EM];
  o.max_lig_atoms = 0; o.cross_pairs = 0; o.atom_pairs = 0;
  for (int g = 0; g < Bo; ++g) {
    const int sg = ex->src_graph[g];
    const long long nl = b.lig_ptr[sg + 1] - b.lig_ptr[sg], na = b.atom_ptr[sg + 1] - b.atom_ptr[sg];
    o.max_lig_atoms = std::max<int>(o.max_lig_atoms, (int)nl);
    o.cross_pairs += nl * na; o.atom_pairs += na * na;
  }
  if (o.max_lig_atoms > POSE_MAX_ATOMS) FAIL(B200_ERR_INVALID, "ligand larger than 256 heavy atoms is not supported");
  // ---- stage the base batch + tables through one pinned arena, one H2D copy
  struct Item { const void* src; size_t bytes; size_t offp; };
  std::vector<Item> items;
  size_t total = 0;
  auto add = [&](const void* p, size_t bytes) { size_t ofs = total; items.push_back({p, bytes, ofs}); total += (bytes + 255) & ~(size_t)255; return ofs; };
  const int one_past = (int)tot[EE];
  const size_t i_lig_node = add(b.lig_node, (size_t)b.N_l * 108), i_lig_pos = add(b.lig_pos, (size_t)b.N_l * 12),
               i_bond_ptr = add(b.bond_ptr, (size_t)(b.N_l + 1) * 4), i_bond_dst = add(b.bond_dst, (size_t)b.E_b * 4),
               i_bond_eid = add(b.bond_eid, (size_t)b.E_b * 4), i_ef = add(b.lig_edge_feat, (size_t)b.E_b * 40),
               i_tb = add(b.tor_bonds, (size_t)b.n_tor * 8), i_rm = add(b.rot_mask, (size_t)b.rot_mask_bytes),
               i_rmo = add(b.rot_mask_off, (size_t)b.n_tor * 8), i_pf = add(b.pocket_feat, (size_t)b.N_a * 20),
               i_rp = add(b.rec_atm_pos, (size_t)b.N_a * 12), i_as = add(b.atom_slot, (size_t)b.N_a * 4),
               i_m14 = add(b.atom14_mask, (size_t)b.N_r * 14), i_seq = add(b.sequence, (size_t)b.N_r * 4),
               i_bt = add(b.backbone_transl, (size_t)b.N_r * 12), i_bR = add(b.backbone_rots, (size_t)b.N_r * 36),
               i_df = add(b.default_frame, (size_t)b.N_r * 512), i_rg = add(b.rigid_group_pos, (size_t)b.N_r * 168),
               i_ta = add(b.torsion_angle, (size_t)b.N_r * 20), i_scb = add(b.sc_bonds, (size_t)b.n_sc * 8),
               i_sci = add(b.sc_index, (size_t)b.N_r * 16), i_tab = add(tab.data(), tab.size() * 4), i_d64 = add(d64.data(), d64.size() * 8),
               i_sid = add(ex->stream_id, ex->stream_id ? (size_t)Bo * 8 : 0), i_last = add(&one_past, 4);
  if (total > h->ex_pin.cap) {
    if (h->ex_pin.p) CK(cudaFreeHost(h->ex_pin.p));
    h->ex_pin.p = nullptr; h->ex_pin.cap = 0;
    CK(cudaMallocHost(&h->ex_pin.p, total + total / 8));
    h->ex_pin.cap = total + total / 8;
  }
  ENS(h->ex_in, total);
  char* pin = (char*)h->ex_pin.p;
  for (const Item& it : items) if (it.bytes) memcpy(pin + it.offp, it.src, it.bytes);
  CK(cudaMemcpyAsync(h->ex_in.p, pin, total, cudaMemcpyHostToDevice, st));
  const char* d = (const char*)h->ex_in.p;
  auto tabp = [&](int e, int which) { return reinterpret_cast<const int*>(d + i_tab) + e * per + (which == 0 ? 0 : (which == 1 ? Bo + 1 : 2 * Bo + 1)); };
  // ---- output arrays
  int slot = 0;
  auto outbuf = [&](size_t bytes) -> void* { Buf& bf = h->ex_out[slot++]; if (ensure(h, bf, bytes ? bytes : 4)) return nullptr; return bf.p; };
  ExpandLaunch L{};
  L.B_out = Bo;
  bool alloc_fail = false;
  auto job = [&](const void* src, size_t row_bytes, int ent, int fix_ent, unsigned fix_mask, int set_graph, bool bytes_mode, bool is64) -> void* {
    const long long rows = bytes_mode ? tot[ent] : tot[ent];
    void* dst = outbuf((size_t)rows * (bytes_mode ? 1 : row_bytes) + 16);
    if (!dst) { alloc_fail = true; return nullptr; }
    ExpandJob& J = L.j[L.n++];
    J.src = src; J.dst = dst; J.words = bytes_mode ? 1 : (int)(row_bytes / 4); J.byte_rows = bytes_mode ? 1 : 0;
    J.new_ptr = tabp(ent, 0); J.old_start = tabp(ent, 1);
    J.delta = fix_ent >= 0 ? tabp(fix_ent, 2) : nullptr;
    J.delta64 = is64 ? reinterpret_cast<const long long*>(d + i_d64) : nullptr;
    J.fix_mask = fix_mask; J.set_graph = set_graph; J.n_rows = (int)rows;
    return dst;
  };
  o.lig_node = (const float*)job(d + i_lig_node, 108, EL, -1, 0, 0, false, false);
  o.lig_pos = (float*)job(d + i_lig_pos, 12, EL, -1, 0, 0, false, false);
  o.lig_batch = (const int*)job(nullptr, 4, EL, -1, 0, 1, false, false);
  o.bond_ptr = (const int*)job(d + i_bond_ptr, 4, EL, EE, 1u, 0, false, false);
  o.bond_dst = (const int*)job(d + i_bond_dst, 4, EE, EL, 1u, 0, false, false);
  o.bond_eid = (const int*)job(d + i_bond_eid, 4, EE, EE, 1u, 0, false, false);
  o.lig_edge_feat = (const float*)job(d + i_ef, 40, EE, -1, 0, 0, false, false);
  o.tor_bonds = (const int*)job(d + i_tb, 8, ET, EL, 3u, 0, false, false);
  o.rot_mask_off = (const int64_t*)job(d + i_rmo, 8, ET, -1, 0, 0, false, true);
  o.rot_mask = (const uint8_t*)job(d + i_rm, 1, EM, -1, 0, 0, true, false);
  o.pocket_feat = (const int*)job(d + i_pf, 20, EA, -1, 0, 0, false, false);
  o.rec_atm_pos = (float*)job(d + i_rp, 12, EA, -1, 0, 0, false, false);
  o.atom_batch = (const int*)job(nullptr, 4, EA, -1, 0, 1, false, false);
  o.atom_slot = (const int*)job(d + i_as, 4, EA, ER14, 1u, 0, false, false);
  o.atom14_mask = (const uint8_t*)job(d + i_m14, 1, ER14, -1, 0, 0, true, false);
  o.sequence = (const int*)job(d + i_seq, 4, ER, -1, 0, 0, false, false);
  o.backbone_transl = (const float*)job(d + i_bt, 12, ER, -1, 0, 0, false, false);
  o.backbone_rots = (const float*)job(d + i_bR, 36, ER, -1, 0, 0, false, false);
  o.default_frame = (const float*)job(d + i_df, 512, ER, -1, 0, 0, false, false);
  o.rigid_group_pos = (const float*)job(d + i_rg, 168, ER, -1, 0, 0, false, false);
  o.torsion_angle = (float*)job(d + i_ta, 20, ER, -1, 0, 0, false, false);
  o.sc_index = (const int*)job(d + i_sci, 16, ER, ES, 15u, 0, false, false);
  o.sc_bonds = (const int*)job(d + i_scb, 8, ES, EA, 3u, 0, false, false);
  if (alloc_fail) return B200_ERR_CUDA;
  o.lig_ptr = tabp(EL, 0); o.tor_ptr = tabp(ET, 0); o.atom_ptr = tabp(EA, 0); o.res_ptr = tabp(ER, 0); o.sc_ptr = tabp(ES, 0);
  k_expand<<<dim3(148 * 2, L.n), 256, 0, st>>>(L);
  CK(cudaMemcpyAsync(const_cast<int*>(o.bond_ptr) + o.N_l, d + i_last, 4, cudaMemcpyDeviceToDevice, st));   // CSR end
  h->launches = 1;
  if (ex->randomize) {
    const unsigned long long* sid = reinterpret_cast<const unsigned long long*>(d + i_sid);
    LigInitArgs LA{Bo, o.lig_pos, o.lig_ptr, o.tor_bonds, o.tor_ptr, o.rot_mask, (const long long*)o.rot_mask_off, sid,
                   (unsigned long long)ex->seed, ex->tr_sigma_max};
    k_lig_init<<<Bo, 32, 0, st>>>(LA);
    ChiInitArgs CA{o.N_r, Bo, o.res_ptr, o.torsion_angle, o.sc_index, sid, (unsigned long long)ex->seed};
    k_chi_init<<<cdiv(o.N_r, 128), 128, 0, st>>>(CA);
    // atom14 / rec_atm_pos from the new chi angles (build_pdb_from_template, prot_math.py:243-291)
    ENS(h->atom14, (size_t)o.N_r * 42 * 4);
    SideChainArgs S{};
    S.N_r = o.N_r; S.N_a = o.N_a; S.sequence = o.sequence; S.bb_t = o.backbone_transl; S.bb_R = o.backbone_rots;
    S.default_frame = o.default_frame; S.rigid_pos = o.rigid_group_pos; S.torsion_angle = o.torsion_angle;
    S.sc_index = o.sc_index; S.atom14_mask = o.atom14_mask; S.atom_slot = o.atom_slot; S.apply_update = 0;
    S.atom14 = h->atom14.as<float>();
    k_sidechain_update<<<cdiv(o.N_r, 64), 64, 0, st>>>(S);
    k_gather_atoms<<<grid_for(o.N_a, 256, 148 * 4), 256, 0, st>>>(S.atom14, o.atom_slot, o.N_a, o.rec_atm_pos);
    h->launches += 4;
  }
  CK(cudaGetLastError());
  *out = o;
  return B200_OK;
}

int b200dock_last_edge_counts(B200Handle* h, int64_t counts[5]) {
  if (!h || !counts) return B200_ERR_INVALID;
  for (int i = 0; i < 5; ++i) counts[i] = h->last_counts[i];
  return B200_OK;
}

int b200dock_last_launch_count(const B200Handle* h, int64_t* n) {
  if (!h || !n) return B200_ERR_INVALID;
  *n = h->launches;
  return B200_OK;
}

int b200dock_set_profiling(B200Handle* h, int enable) {
  if (!h) return B200_ERR_INVALID;
  h->profiling = enable != 0;
  h->tp_events.clear(); h->event_used = 0;
  return B200_OK;
}

int b200dock_tp_kernel_time_ms(B200Handle* h, double* ms, int64_t* launches) {
  if (!h || !ms || !launches) return B200_ERR_INVALID;
  double tot = 0;
  for (auto& pr : h->tp_events) {
    float t = 0;
    CK(cudaEventSynchronize(pr.second));
    CK(cudaEventElapsedTime(&t, pr.first, pr.second));
    tot += t;
  }
  *ms = tot; *launches = (int64_t)h->tp_events.size();
  h->tp_events.clear(); h->event_used = 0;
  return B200_OK;
}

int b200dock_debug_tap(B200Handle* h, int what, int arg, void* host_out, size_t cap_bytes, size_t* n_bytes) {
  if (!h || !host_out || !n_bytes || !h->have_last) return B200_ERR_INVALID;
  CK(cudaSetDevice(h->device));
  CK(cudaDeviceSynchronize());
  const B200Batch& b = h->last_batch;
  const void* src = nullptr; size_t bytes = 0;
  if (what == B200_TAP_H_LIG) { src = h->h_lig.p; bytes = (size_t)b.N_l * HS * 4; }
  else if (what == 7) { src = h->trace.p; bytes = h->trace.p ? (size_t)B200_TRACE_WORDS * 8 : 0; }
  else if (what == B200_TAP_H_ATOM) { src = h->h_atom.p; bytes = (size_t)b.N_a * HS * 4; }
  else if (what == B200_TAP_EDGES) {
    if (arg < 0 || arg > 5) return B200_ERR_INVALID;
    ConvWs& w = h->cw[arg];
    int E = 0;
    if (w.T > 0) CK(cudaMemcpy(&E, w.seg.as<int>() + w.T, 4, cudaMemcpyDeviceToHost));
    bytes = (size_t)E * 8;
    if (bytes > cap_bytes) FAIL(B200_ERR_INVALID, "tap buffer too small");
    std::vector<int> s(E), d(E);
    if (E) { CK(cudaMemcpy(s.data(), w.es.p, (size_t)E * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(d.data(), w.ed.p, (size_t)E * 4, cudaMemcpyDeviceToHost)); }
    int* o = (int*)host_out;
    size_t n_real = 0;
    for (int i = 0; i < E; ++i) if (s[i] >= 0) { o[2 * n_real] = s[i]; o[2 * n_real + 1] = d[i]; ++n_real; }   // es = -1: alignment padding
    *n_bytes = n_real * 8;
    return B200_OK;
  } else if (what == B200_TAP_CONV_BUF) {
    int conv = arg / 16, which = arg % 16;
    if (conv < 0 || conv > 5) return B200_ERR_INVALID;
    ConvWs& w = h->cw[conv];
    int E = 0;
    if (w.T > 0) CK(cudaMemcpy(&E, w.seg.as<int>() + w.T, 4, cudaMemcpyDeviceToHost));
    size_t Ep = (size_t)((E + 127) / 128) * 128;
    if (which == 0) { src = w.emb.p; bytes = (size_t)E * NSC * 4; }
    else if (which == 1) { src = w.sh.p; bytes = (size_t)E * (conv >= 4 ? 8 : 9) * 4; }
    else if (which == 2 && w.H1.p) { src = w.H1.p; bytes = Ep * KP * 4; }
    else if (which == 3 && w.Zt.p) { src = w.Zt.p; bytes = Ep * w.z_max * 4; }
    else if (which == 4 && w.msg.p) { src = w.msg.p; bytes = Ep * HS * 4; }
    else if (which == 5) { src = w.seg.p; bytes = (size_t)(w.T + 2) * 4; }
    else if (which == 7) { src = w.counts.p; bytes = (size_t)w.T * 4; }
    else if (which == 8) { src = w.es.p; bytes = Ep * 4; }
    else if (which == 6) { src = h->cmsg.p; bytes = (size_t)b.N_l * 12 * 4; }
    else return B200_ERR_INVALID;
  } else return B200_ERR_INVALID;
  if (bytes > cap_bytes) FAIL(B200_ERR_INVALID, "tap buffer too small");
  CK(cudaMemcpy(host_out, src, bytes, cudaMemcpyDeviceToHost));
  *n_bytes = bytes;
  return B200_OK;
}

int b200dock_debug_set(B200Handle* h, int key, int value) {
  if (!h) return B200_ERR_INVALID;
  if (key == 0) { h->debug_layers = value; return B200_OK; }
  if (key == 2) { h->warp_node_update = value != 0; return B200_OK; }
  if (key == 1) {   // wait-cycle accounting of the fused conv kernel (mode 5): 148 CTAs x 32 counters, accumulated over launches
    CK(cudaSetDevice(h->device));
    ENS(h->trace, (size_t)B200_TRACE_WORDS * 8);
    CK(cudaMemset(h->trace.p, 0, (size_t)B200_TRACE_WORDS * 8));
    h->trace_on = value != 0;
    return B200_OK;
  }
  return B200_ERR_INVALID;
}

}  // extern "C"
