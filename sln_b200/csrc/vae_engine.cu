// Host-side orchestration + C ABI of the VAE-graph hot path (see include/sln_b200.h).
// One call = one whole encoder/decoder forward or backward: a fixed sequence of launches on the caller's stream,
// no allocation, no synchronisation -> capturable in a CUDA graph.
//
// Reference being replaced: models/graph.py:57-143 (GraphTripleConv[Net]), models/Sg2ScVAE_model.py:115-188
// (encoder / decoder / forward), utils.py:12-33 (losses), train.py:82-84 (Adam).
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <unordered_map>
#include <vector>

#include "../../include/sln_b200.h"
#include "vae_kernels.cuh"
#include "tc_gemm.cuh"

namespace sln {

namespace {

// contraction engine of the MLP stage: 1 = tcgen05 3xTF32 tiles (default), 0 = FP32 SIMT tiles (A/B checks, odd shapes)
int g_engine = 1;
// debug knob: sln_set_engine(16 | mask) keeps tcgen05 only for the call sites in `mask` (1 fwd, 2 bwd_w, 4 bwd_x_plain, 8 bwd_x_masked)
int g_tc_mask = 15;
enum { SITE_FWD = 1, SITE_BWD_W = 2, SITE_BWD_X = 4, SITE_BWD_XM = 8 };
// debug knob (env SLN_SKINNY, default 7): narrow-head kernels per call site (1 fwd, 2 bwd_w, 4 bwd_x)
int g_skinny = -1;
inline bool use_skinny(int site) {
  if (g_skinny < 0) { const char* e = getenv("SLN_SKINNY"); g_skinny = e ? atoi(e) : 7; }
  return (g_skinny & site) != 0;
}
inline bool use_tc(int M, int N, int K, int site) { return g_engine == 1 && (g_tc_mask & site) && tc::tc_eligible(M, N, K); }

constexpr int kMaxLayers = 32;

struct Lin {
  const float* W; const float* b; float* dW; float* db; int in, out;
  const float* pf;   // pre-split, pre-tiled image of W  [out][in]  (forward B operand; tc::PackedB), or null
  const float* pb;   // ... of W^T [in][out] (backward-data B operand), or null
};
struct PackJob { const float* W; float* out; int N, K, ldw, transposed; };
struct Blk {  // Linear [+ BatchNorm1d] [+ ReLU]   (reference graph.py:10-27)
  Lin lin;
  int has_bn, relu;
  const float* gamma; const float* beta; float* dgamma; float* dbeta;
  float* rm; float* rv; long long* nbt;
};
struct BlkState {  // per-instance saved tensors (workspace)
  float* y;  // pre-BN output [M, out]
  int M;
  float *mean, *rstd, *scale, *shift;  // BN statistics / lazy affine [out]
  float *p, *q, *r;                    // BN-backward coefficients [out]
  float* partial;                      // [2 * ceil(M/64) * out]
  unsigned* counter;
  float* g;  // masked gradient w.r.t. the post-activation output [M, out] (backward scratch)
};

struct Dims {
  int E, D, H, Z, L, Lw, obj_w, attr_w, box_w, ang_w, box_dim, n_angle;
  int norm, training, num_preds;
  float eps, momentum;
};

struct Model {
  const float* emb[7];
  float* demb[7];
  Blk box_emb;
  Blk enc[kMaxLayers][4], dec[kMaxLayers][4];
  Blk bmv[2], amv[2], box_mean, box_var, angle_mean, angle_var, box_net[2], angle_net[2];
};

// Weight-gradient contractions (dW = dy^T X) only feed the optimizer, so they leave the dependency chain of the backward
// pass: they are forked onto a side stream (event record / wait, capturable in a CUDA graph as a parallel branch) and run
// concurrently with the data-gradient chain, filling the SMs that the 62..128-CTA grids of the chain leave idle.
// Hazards: a dW kernel READS a masked-gradient buffer `g` that a later chain kernel overwrites (the g buffers are shared
// by all layers), so every chain kernel that writes such a buffer first waits for the last side-stream reader of it.
constexpr int kLeafStreams = 3;
constexpr int kSideEvents = 256;
struct Side {
  cudaStream_t st = nullptr;
  cudaStream_t leaf[kLeafStreams];   // the embedding-table gradients at the tail of a backward call: small independent kernels, one stream each
  int leaf_next = 0;
  unsigned leaf_used = 0u;
  cudaEvent_t ev[kSideEvents];   // ring: far more than one backward call takes between a reader's record and the matching wait
  int next = 0;
  bool ready = false, active = false;
  std::unordered_map<const void*, cudaEvent_t> readers;
  cudaEvent_t take() { cudaEvent_t e = ev[next]; next = (next + 1) % kSideEvents; return e; }
};
thread_local Side g_side;
int g_side_enabled = -1;

struct Ctx {
  cudaStream_t st;
  Dims dm;
  bool side = false;   // weight gradients go to the side stream during this call
  // SyncBatchNorm (sln_vae_desc.bn_sync): table pointer, first slot of this call (direction / encoder|decoder) and the workspace's
  // counter base, so that a layer's slot is slot_off + (its counter - cbase)
  const sln_bn_sync* sync = nullptr;
  int slot_off = 0;
  const unsigned* cbase = nullptr;
};
inline void ctx_sync(Ctx& c, const sln_vae_desc* d, const unsigned* cbase, int which, int direction) {
  c.sync = (d->bn_sync && d->norm == 1 && d->training) ? (const sln_bn_sync*)d->bn_sync : nullptr;
  c.slot_off = (direction * 2 + which) * SLN_BN_SYNC_SLOTS;
  c.cbase = cbase;
}

// Decide whether this call uses the side stream (creating it on first use; never created while the caller is capturing).
void side_begin(Ctx& c) {
  c.side = false;
  if (g_side_enabled < 0) { const char* e = getenv("SLN_SIDE_STREAM"); g_side_enabled = (e && e[0] == '0') ? 0 : 1; }
  if (!g_side_enabled) return;
  Side& sd = g_side;
  if (!sd.ready) {
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(c.st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { (void)cudaGetLastError(); return; }
    if (cudaStreamCreateWithFlags(&sd.st, cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); return; }
    for (int i = 0; i < kLeafStreams; ++i)
      if (cudaStreamCreateWithFlags(&sd.leaf[i], cudaStreamNonBlocking) != cudaSuccess) { (void)cudaGetLastError(); return; }
    for (int i = 0; i < kSideEvents; ++i)
      if (cudaEventCreateWithFlags(&sd.ev[i], cudaEventDisableTiming) != cudaSuccess) { (void)cudaGetLastError(); return; }
    sd.ready = true;
  }
  sd.readers.clear();
  sd.active = false;
  sd.leaf_next = 0; sd.leaf_used = 0u;
  c.side = true;
}
// stream for a leaf kernel (reads chain results, feeds only the optimizer, nothing overwrites its inputs before side_end)
cudaStream_t side_fork_leaf(const Ctx& c) {
  if (!c.side) return c.st;
  Side& sd = g_side;
  const int k = sd.leaf_next++ % kLeafStreams;
  cudaEvent_t e = sd.take();
  cudaEventRecord(e, c.st);
  cudaStreamWaitEvent(sd.leaf[k], e, 0);
  sd.leaf_used |= 1u << k;
  return sd.leaf[k];
}
// stream for a weight-gradient launch that reads buffer `g`: forks from the chain (everything issued so far is visible)
cudaStream_t side_fork(const Ctx& c) {
  if (!c.side) return c.st;
  Side& sd = g_side;
  cudaEvent_t e = sd.take();
  cudaEventRecord(e, c.st);
  cudaStreamWaitEvent(sd.st, e, 0);
  sd.active = true;
  return sd.st;
}
void side_read_done(const Ctx& c, const void* g) {
  if (!c.side) return;
  Side& sd = g_side;
  cudaEvent_t e = sd.take();
  cudaEventRecord(e, sd.st);
  sd.readers[g] = e;
}
// the chain is about to overwrite buffer `g`
void side_before_write(const Ctx& c, const void* g) {
  if (!c.side || g == nullptr) return;
  Side& sd = g_side;
  auto it = sd.readers.find(g);
  if (it == sd.readers.end()) return;
  cudaStreamWaitEvent(c.st, it->second, 0);
  sd.readers.erase(it);
}
// the chain waits for everything issued to the leaf streams so far (a branch of the chain that ran there rejoins it)
int side_join_leaf(const Ctx& c) {
  if (!c.side) return SLN_OK;
  Side& sd = g_side;
  for (int k = 0; k < kLeafStreams; ++k) {
    if (!(sd.leaf_used >> k & 1u)) continue;
    cudaEvent_t e = sd.take();
    SLN_CUDA_TRY(cudaEventRecord(e, sd.leaf[k]));
    SLN_CUDA_TRY(cudaStreamWaitEvent(c.st, e, 0));
  }
  sd.leaf_used = 0u;
  return SLN_OK;
}
// join: every weight gradient of this call is complete before anything the caller enqueues next
int side_end(Ctx& c) {
  if (!c.side) return SLN_OK;
  Side& sd = g_side;
  if (sd.active) {
    cudaEvent_t e = sd.take();
    SLN_CUDA_TRY(cudaEventRecord(e, sd.st));
    SLN_CUDA_TRY(cudaStreamWaitEvent(c.st, e, 0));
  }
  for (int k = 0; k < kLeafStreams; ++k) {
    if (!(sd.leaf_used >> k & 1u)) continue;
    cudaEvent_t e = sd.take();
    SLN_CUDA_TRY(cudaEventRecord(e, sd.leaf[k]));
    SLN_CUDA_TRY(cudaStreamWaitEvent(c.st, e, 0));
  }
  sd.leaf_used = 0u; sd.leaf_next = 0;
  sd.readers.clear();
  sd.active = false;
  c.side = false;
  return SLN_OK;
}

int make_dims(const sln_vae_desc* d, Dims* o) {
  SLN_CHECK_ARG(d != nullptr, "null model descriptor");
  SLN_CHECK_ARG(d->embedding_dim > 0 && d->embedding_dim % 4 == 0, "embedding_dim must be a positive multiple of 4 (got %d)", d->embedding_dim);
  SLN_CHECK_ARG(d->n_layers >= 1 && d->n_layers <= kMaxLayers, "gconv_num_layers must be in [1,%d] (got %d)", kMaxLayers, d->n_layers);
  SLN_CHECK_ARG(d->norm == 0 || d->norm == 1, "norm must be 0 ('none') or 1 ('batch')");
  SLN_CHECK_ARG(d->box_dim == 6 || d->box_dim == 4, "box_dim must be 6 or 4");
  SLN_CHECK_ARG(d->n_angle > 0 && d->num_objs > 0 && d->num_preds > 0 && d->num_attrs > 0, "vocabulary sizes must be positive");
  o->E = d->embedding_dim; o->D = 2 * o->E; o->H = 4 * o->E; o->Z = o->E;
  if (d->gconv_dim_override > 0) o->D = d->gconv_dim_override;      // standalone GraphTripleConv only
  if (d->gconv_hidden_override > 0) o->H = d->gconv_hidden_override;
  SLN_CHECK_ARG(o->D % 4 == 0 && o->H % 4 == 0, "gconv dims must be multiples of 4 (D=%d, H=%d)", o->D, o->H);
  o->L = d->n_layers; o->Lw = d->recurrent ? 1 : d->n_layers;
  o->obj_w = o->E * 3 / 4; o->attr_w = o->E / 4; o->box_w = o->E * 3 / 4; o->ang_w = o->E / 4;
  o->box_dim = d->box_dim; o->n_angle = d->n_angle;
  o->norm = d->norm; o->training = d->training; o->num_preds = d->num_preds;
  o->eps = d->bn_eps; o->momentum = d->bn_momentum;
  return SLN_OK;
}

int count_params(const Dims& dm) {
  int bn = dm.norm ? 2 : 0;
  int n = 7;
  n += 2;                              // box_embeddings
  n += 2 * dm.Lw * 4 * (2 + bn);       // enc + dec gconv
  n += 4 * (2 + bn);                   // box_mean_var, angle_mean_var
  n += 4 * 2;                          // box_mean/var, angle_mean/var
  n += (2 + bn) + 2 + (2 + bn) + 2;    // box_net, angle_net
  return n;
}
int count_bn(const Dims& dm) { return dm.norm ? (2 * dm.Lw * 4 + 4 + 2) : 0; }

struct TableReader {
  const void* const* params; void* const* grads; void* const* bn; int ip, ib; int norm;
  float* pack_base = nullptr;             // packed weight images (sln_vae_desc::packed_weights) or null
  size_t pack_off = 0;                    // floats; advanced even when pack_base is null (size query)
  std::vector<PackJob>* jobs = nullptr;   // filled by sln_vae_pack_weights
  Blk take(int in, int out, bool with_bn, bool relu) {
    Blk b; memset(&b, 0, sizeof(b));
    b.lin.in = in; b.lin.out = out; b.relu = relu ? 1 : 0;
    b.lin.W = (const float*)params[ip]; b.lin.dW = grads ? (float*)grads[ip] : nullptr; ++ip;
    if (in >= 32 && out >= 32 && in % 4 == 0) {   // the shapes the tensor-core path accepts (tc_eligible) get packed images
      const size_t nf = tc::packed_weight_floats(out, in), nb = tc::packed_weight_floats(in, out);
      if (pack_base) { b.lin.pf = pack_base + pack_off; b.lin.pb = pack_base + pack_off + nf; }
      if (jobs && pack_base) {
        jobs->push_back(PackJob{b.lin.W, pack_base + pack_off, out, in, in, 0});
        jobs->push_back(PackJob{b.lin.W, pack_base + pack_off + nf, in, out, in, 1});
      }
      pack_off += nf + nb;
    }
    b.lin.b = (const float*)params[ip]; b.lin.db = grads ? (float*)grads[ip] : nullptr; ++ip;
    if (with_bn && norm) {
      b.has_bn = 1;
      b.gamma = (const float*)params[ip]; b.dgamma = grads ? (float*)grads[ip] : nullptr; ++ip;
      b.beta = (const float*)params[ip]; b.dbeta = grads ? (float*)grads[ip] : nullptr; ++ip;
      if (bn) { b.rm = (float*)bn[ib]; b.rv = (float*)bn[ib + 1]; b.nbt = (long long*)bn[ib + 2]; }
      ib += 3;
    }
    return b;
  }
};

void take_gconv(TableReader& tr, const Dims& dm, Blk* blk) {
  blk[0] = tr.take(3 * dm.D, dm.H, true, true);
  blk[1] = tr.take(dm.H, 2 * dm.H + dm.D, true, true);
  blk[2] = tr.take(dm.H, dm.H, true, true);
  blk[3] = tr.take(dm.H, dm.D, true, true);
}

int parse_model(const Dims& dm, const void* const* params, void* const* grads, void* const* bn, Model* m, const void* packed = nullptr,
                std::vector<PackJob>* jobs = nullptr, size_t* packed_floats = nullptr) {
  SLN_CHECK_ARG(params != nullptr, "null parameter table");
  memset(m, 0, sizeof(Model));
  for (int i = 0; i < 7; ++i) { m->emb[i] = (const float*)params[i]; m->demb[i] = grads ? (float*)grads[i] : nullptr; }
  TableReader tr{params, grads, bn, 7, 0, dm.norm};
  tr.pack_base = (float*)packed; tr.jobs = jobs;
  m->box_emb = tr.take(dm.box_dim, dm.box_w, false, false);
  for (int l = 0; l < dm.Lw; ++l) take_gconv(tr, dm, m->enc[l]);
  for (int l = dm.Lw; l < dm.L; ++l) for (int k = 0; k < 4; ++k) m->enc[l][k] = m->enc[0][k];
  for (int l = 0; l < dm.Lw; ++l) take_gconv(tr, dm, m->dec[l]);
  for (int l = dm.Lw; l < dm.L; ++l) for (int k = 0; k < 4; ++k) m->dec[l][k] = m->dec[0][k];
  m->bmv[0] = tr.take(dm.D, dm.H, true, true);
  m->bmv[1] = tr.take(dm.H, dm.D, true, true);
  m->amv[0] = tr.take(dm.D, dm.H, true, true);
  m->amv[1] = tr.take(dm.H, dm.D, true, true);
  m->box_mean = tr.take(dm.D, dm.box_w, false, false);
  m->box_var = tr.take(dm.D, dm.box_w, false, false);
  m->angle_mean = tr.take(dm.D, dm.ang_w, false, false);
  m->angle_var = tr.take(dm.D, dm.ang_w, false, false);
  m->box_net[0] = tr.take(dm.D + dm.attr_w, dm.H, true, true);
  m->box_net[1] = tr.take(dm.H, dm.box_dim, false, false);
  m->angle_net[0] = tr.take(dm.D, dm.H, true, true);
  m->angle_net[1] = tr.take(dm.H, dm.n_angle, false, false);
  if (tr.ip != count_params(dm)) { set_error("internal: parameter table walk mismatch (%d vs %d)", tr.ip, count_params(dm)); return SLN_EINVAL; }
  if (packed_floats) *packed_floats = tr.pack_off;
  return SLN_OK;
}

// all weight images of a model in a few launches: the job table travels by value in the kernel parameters (graph-capturable)
constexpr int kPackJobsPerLaunch = 96;
struct PackJobs { PackJob job[kPackJobsPerLaunch]; int unit_start[kPackJobsPerLaunch + 1]; int n; };
__global__ void __launch_bounds__(256) k_pack_jobs(const PackJobs jobs) {
  int j = 0;
  while (j + 1 < jobs.n && (int)blockIdx.x >= jobs.unit_start[j + 1]) ++j;
  const PackJob& jb = jobs.job[j];
  const int unit = blockIdx.x - jobs.unit_start[j], kchunks = ceil_div(jb.K, 32);
  const int n32 = unit / kchunks, kc = unit - n32 * kchunks;
  const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
  const int n = n32 * 32 + r, k = kc * 32 + q * 4;
  float v[4], h[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const bool ok = n < jb.N && k + e < jb.K;
    v[e] = ok ? __ldg(jb.transposed ? jb.W + (size_t)(k + e) * jb.ldw + n : jb.W + (size_t)n * jb.ldw + k + e) : 0.f;
    h[e] = tc::tf32_hi(v[e]);
  }
  float* u = jb.out + (size_t)unit * tc::PACK_UNIT_FLOATS;
  const uint32_t off = tc::sw128(r, q * 4) / 4;
  *reinterpret_cast<float4*>(u + off) = make_float4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<float4*>(u + 1024 + off) = make_float4(v[0] - h[0], v[1] - h[1], v[2] - h[2], v[3] - h[3]);
}

// ---------------------------------------------------------------- workspace plans
struct CounterPool { unsigned* base; int n, cap; };

void plan_state(Arena& ar, CounterPool& cp, int M, int out, BlkState* s) {
  s->M = M;
  s->y = ar.take<float>((size_t)M * out);
  s->mean = ar.take<float>(out); s->rstd = ar.take<float>(out); s->scale = ar.take<float>(out); s->shift = ar.take<float>(out);
  s->p = ar.take<float>(out); s->q = ar.take<float>(out); s->r = ar.take<float>(out);
  s->partial = ar.take<float>((size_t)2 * max_row_tiles(M) * out);
  s->counter = cp.base + (size_t)cp.n * kCounterStride; cp.n++;
  s->g = nullptr;
}

void plan_graph(Arena& ar, int O, int T, Graph* g, int** deg, int** err) {
  g->O = O; g->T = T;
  g->s_idx = ar.take<int>(T); g->p_idx = ar.take<int>(T); g->o_idx = ar.take<int>(T);
  g->row_ptr = ar.take<int>(O + 1); g->ent = ar.take<int>((size_t)2 * T);
  g->cnt = ar.take<float>(O); g->cursor = ar.take<int>(O);
  *deg = ar.take<int>(O); *err = ar.take<int>(1);
}

struct GconvScratch { float *g1, *g2, *g3, *g4, *dpooled, *dcat[2]; };
void plan_gconv_scratch(Arena& ar, const Dims& dm, int O, int T, GconvScratch* s) {
  s->g1 = ar.take<float>((size_t)T * dm.H);
  s->g2 = ar.take<float>((size_t)T * (2 * dm.H + dm.D));
  s->g3 = ar.take<float>((size_t)O * dm.H);
  s->g4 = ar.take<float>((size_t)O * dm.D);
  s->dpooled = ar.take<float>((size_t)O * dm.H);
  s->dcat[0] = ar.take<float>((size_t)T * 3 * dm.D);
  s->dcat[1] = ar.take<float>((size_t)T * 3 * dm.D);
}

struct NetPlan {
  Graph g; int* deg; int* err;
  int *objs32, *attrs32, *angles32;
  float *obj0, *pred0;
  BlkState st[kMaxLayers][4];
  float* pooled[kMaxLayers];
  GconvScratch sc;
  float* dobj0;
  // encoder heads
  BlkState bmv[2], amv[2];
  float *tmpA, *tmpA2, *tmpB;
  // decoder heads
  BlkState box_net0, angle_net0;
  float *logits, *logp, *dlogits, *tmpC;
  CounterPool cp;
  size_t bytes;
};

constexpr int kCounterCap = (4 * kMaxLayers + 16) * kCounterStride;
static_assert(kCounterCap <= SLN_BN_SYNC_SLOTS, "one sync slot per BatchNorm counter");

// which: 0 encoder, 1 decoder, 2 single standalone layer
void make_plan(const Dims& dm, int O, int T, int which, void* ws, NetPlan* p) {
  Arena ar(ws, (size_t)-1);
  memset(p, 0, sizeof(NetPlan));
  p->cp.base = ar.take<unsigned>(kCounterCap); p->cp.n = 0; p->cp.cap = kCounterCap;
  plan_graph(ar, O, T, &p->g, &p->deg, &p->err);
  p->objs32 = ar.take<int>(O); p->attrs32 = ar.take<int>(O); p->angles32 = ar.take<int>(O);
  p->obj0 = ar.take<float>((size_t)O * dm.D);
  p->pred0 = ar.take<float>((size_t)T * dm.D);
  const int L = which == 2 ? 1 : dm.L;
  for (int l = 0; l < L; ++l) {
    plan_state(ar, p->cp, T, dm.H, &p->st[l][0]);
    plan_state(ar, p->cp, T, 2 * dm.H + dm.D, &p->st[l][1]);
    plan_state(ar, p->cp, O, dm.H, &p->st[l][2]);
    plan_state(ar, p->cp, O, dm.D, &p->st[l][3]);
    p->pooled[l] = ar.take<float>((size_t)O * dm.H);
  }
  plan_gconv_scratch(ar, dm, O, T, &p->sc);
  for (int l = 0; l < L; ++l) { p->st[l][0].g = p->sc.g1; p->st[l][1].g = p->sc.g2; p->st[l][2].g = p->sc.g3; p->st[l][3].g = p->sc.g4; }
  p->dobj0 = ar.take<float>((size_t)O * dm.D);
  if (which == 0) {
    plan_state(ar, p->cp, O, dm.H, &p->bmv[0]); plan_state(ar, p->cp, O, dm.D, &p->bmv[1]);
    plan_state(ar, p->cp, O, dm.H, &p->amv[0]); plan_state(ar, p->cp, O, dm.D, &p->amv[1]);
    p->bmv[0].g = ar.take<float>((size_t)O * dm.H); p->bmv[1].g = ar.take<float>((size_t)O * dm.D);
    p->amv[0].g = ar.take<float>((size_t)O * dm.H); p->amv[1].g = ar.take<float>((size_t)O * dm.D);
    p->tmpA = ar.take<float>((size_t)O * dm.D); p->tmpB = ar.take<float>((size_t)O * dm.D); p->tmpA2 = ar.take<float>((size_t)O * dm.D);
  } else if (which == 1) {
    plan_state(ar, p->cp, O, dm.H, &p->box_net0); plan_state(ar, p->cp, O, dm.H, &p->angle_net0);
    p->box_net0.g = ar.take<float>((size_t)O * dm.H); p->angle_net0.g = ar.take<float>((size_t)O * dm.H);
    p->logits = ar.take<float>((size_t)O * dm.n_angle); p->logp = ar.take<float>((size_t)O * dm.n_angle);
    p->dlogits = ar.take<float>((size_t)O * dm.n_angle);
    p->tmpC = ar.take<float>((size_t)O * (dm.D + dm.attr_w));
  }
  p->bytes = ar.off;
}

// ---------------------------------------------------------------- building blocks
// algorithmic bytes of the avg-pool gather-reduce: read new_s,new_o (2*T*H fp32) + CSR (2T entries + O+1 offsets + O counts),
// write pooled (O*H fp32)   
double pool_bytes(int O, int T, int H) { return 4.0 * (2.0 * T * H + 2.0 * T + 2.0 * O + 1.0 + (double)O * H); }
int norm_mode(const Ctx& c, const Blk& b) { return !b.has_bn ? NORM_NONE : (c.dm.training ? NORM_BN_TRAIN : NORM_BN_EVAL); }

MatView block_out(const Blk& b, const BlkState& s) {
  return make_view(s.y, b.lin.out, s.M, b.lin.out, b.has_bn ? s.scale : nullptr, b.has_bn ? s.shift : nullptr, b.relu);
}
MatView slice_cols(const MatView& v, int off, int width) {
  return make_view(v.p + off, v.ld, v.rows, width, v.scale ? v.scale + off : nullptr, v.shift ? v.shift + off : nullptr, v.relu);
}
MatView weight_view(const Lin& l) { return make_view(l.W, l.in, l.out, l.in); }

template <class AOp>
int block_fwd(const Ctx& c, const AOp& A, int M, const Blk& b, BlkState& s, float* out = nullptr, int ldo = 0, const tc::TileSel* sel = nullptr,
              bool skip_eval_prep = false) {
  EpiStore epi; memset(&epi, 0, sizeof(epi));
  epi.C = out ? out : s.y; epi.ldc = out ? ldo : b.lin.out; epi.bias = b.lin.b;
  int mode = norm_mode(c, b);
  if (mode == NORM_BN_TRAIN) {
    BnFwdFin& f = epi.fin;
    f.enabled = 1; f.partial = s.partial; f.counter = s.counter; f.gamma = b.gamma; f.beta = b.beta;
    f.running_mean = b.rm; f.running_var = b.rv; f.nbt = b.nbt;
    f.mean = s.mean; f.rstd = s.rstd; f.scale = s.scale; f.shift = s.shift;
    f.eps = c.dm.eps; f.momentum = c.dm.momentum; f.M = M;
    if (c.sync) { f.sync = c.sync; f.slot0 = c.slot_off + (int)(s.counter - c.cbase); }
  } else if (mode == NORM_BN_EVAL && !skip_eval_prep) {
    SLN_CHECK_ARG(b.rm && b.rv, "eval-mode BatchNorm needs running statistics");
    k_bn_eval_prep<<<ceil_div(b.lin.out, 128), 128, 0, c.st>>>(b.gamma, b.beta, b.rm, b.rv, c.dm.eps, b.lin.out, s.mean, s.rstd, s.scale, s.shift);
    SLN_TRY(check_launch("bn_eval_prep"));
  }
  if (use_skinny(1) && mode == NORM_NONE && b.lin.out <= 32 && b.lin.in <= 1024)   // narrow heads: latency problems, not contractions
    return launch_skinny_fwd(c.st, A, b.lin.W, b.lin.b, M, b.lin.out, b.lin.in, epi.C, epi.ldc);
  if (use_tc(M, b.lin.out, b.lin.in, SITE_FWD) && A.vec_ok() && weight_view(b.lin).vec_ok()) {
    tc::TcEpiStore te{epi.C, epi.ldc, epi.bias, epi.fin};
    if (b.lin.pf) {   // pre-split weight image: the B tiles arrive by cp.async.bulk
      tc::PackedB pb{b.lin.pf, ceil_div(b.lin.in, tc::BK)};
      return tc::launch_tc<true, true>(c.st, A, pb, te, M, b.lin.out, b.lin.in, false, "linear_fwd_tc_packed", PROF_GEMM_FWD, sel ? *sel : tc::TileSel());
    }
    if (sel) { set_error("internal: column-tile subsets need the packed tensor-core path"); return SLN_EINVAL; }
    return tc::launch_tc<true, true>(c.st, A, weight_view(b.lin), te, M, b.lin.out, b.lin.in, false, "linear_fwd_tc", PROF_GEMM_FWD);
  }
  return launch_gemm<true, true>(c.st, A, weight_view(b.lin), epi, M, b.lin.out, b.lin.in, false, "linear_fwd", PROF_GEMM_FWD);
}

DyView blk_dy(const Ctx& c, const Blk& b, const BlkState& s) {
  int mode = norm_mode(c, b);
  if (mode == NORM_NONE) return make_dy(s.g, b.lin.out, s.M, b.lin.out);
  if (mode == NORM_BN_EVAL) return make_dy(s.g, b.lin.out, s.M, b.lin.out, nullptr, 0, s.p);
  return make_dy(s.g, b.lin.out, s.M, b.lin.out, s.y, b.lin.out, s.p, s.q, s.r);
}
BnBwdFin blk_fin(const Ctx& c, const Blk& b, const BlkState& s) {
  BnBwdFin f; memset(&f, 0, sizeof(f));
  f.mode = norm_mode(c, b); f.partial = s.partial; f.counter = s.counter; f.gamma = b.gamma;
  f.mean = s.mean; f.rstd = s.rstd; f.scale = s.scale; f.p = s.p; f.q = s.q; f.r = s.r;
  f.dgamma = b.dgamma; f.dbeta = b.dbeta; f.dbias = b.lin.db; f.M = s.M;
  if (c.sync && f.mode == NORM_BN_TRAIN) { f.sync = c.sync; f.slot0 = c.slot_off + (int)(s.counter - c.cbase); }
  return f;
}
ActInfo blk_act(const Blk& b, const BlkState& s) {
  ActInfo a; memset(&a, 0, sizeof(a));
  a.has_act = 1; a.y = s.y; a.ldy = b.lin.out;
  if (b.has_bn) { a.scale = s.scale; a.shift = s.shift; a.mean = s.mean; a.rstd = s.rstd; }
  return a;
}

// dW[out,in] += dy^T X     (split along the sample dimension, RED.ADD); runs on the side stream (see Side)
// head_bias: also reduce the Linear bias gradient (layers without a following activation: their bias gradient is not
// produced by a BatchNorm-backward finalisation)
template <class XOp>
int bwd_w(const Ctx& c, const DyView& dy, const XOp& X, const Lin& lin, int M, bool head_bias = false) {
  if (!lin.dW) return head_bias && lin.db ? launch_embed_bwd(c.st, dy.g, dy.ldg, nullptr, 0, nullptr, M, lin.out, lin.db, 1) : SLN_OK;
  const bool skinny = use_skinny(2) && lin.out <= 32 && dy.p == nullptr;   // narrow layer with a plain gradient [M, out]
  if (head_bias && !skinny && lin.db) SLN_TRY(launch_embed_bwd(c.st, dy.g, dy.ldg, nullptr, 0, nullptr, M, lin.out, lin.db, 1));
  cudaStream_t st = side_fork(c);
  int rc;
  if (skinny) {
    rc = launch_skinny_bwd_w(st, dy.g, dy.ldg, lin.out, X, M, lin.in, lin.dW, lin.in, false, head_bias ? lin.db : nullptr);
  } else if (use_tc(lin.out, lin.in, M, SITE_BWD_W) && dy.vec_ok() && X.vec_ok()) {
    tc::TcEpiAtomic te{lin.dW, lin.in};
    rc = tc::launch_tc<false, false>(st, dy, X, te, lin.out, lin.in, M, true, "linear_bwd_w_tc", PROF_GEMM_BWD_W);
  } else {
    EpiAtomic epi{lin.dW, lin.in};
    rc = launch_gemm<false, false>(st, dy, X, epi, lin.out, lin.in, M, true, "linear_bwd_w", PROF_GEMM_BWD_W);
  }
  side_read_done(c, dy.g);
  return rc;
}
// dX[M,in] = dy W
int bwd_x_plain(const Ctx& c, const DyView& dy, const Lin& lin, int M, float* dX, int ldx) {
  if (use_skinny(4) && lin.out <= 32 && dy.p == nullptr) {
    SmallKSrc src{dy.g, dy.ldg, lin.out, lin.W, lin.in, nullptr, 0};
    ActInfo none; memset(&none, 0, sizeof(none));
    BnBwdFin nofin; memset(&nofin, 0, sizeof(nofin));
    return launch_prep(c.st, src, none, dX, ldx, nofin, M, lin.in, "skinny_bwd_x");
  }
  EpiStore epi; memset(&epi, 0, sizeof(epi));
  epi.C = dX; epi.ldc = ldx;
  if (use_tc(M, lin.in, lin.out, SITE_BWD_X) && dy.vec_ok() && weight_view(lin).vec_ok()) {
    tc::TcEpiStore te{dX, ldx, nullptr, epi.fin};
    if (lin.pb) {
      tc::PackedB pb{lin.pb, ceil_div(lin.out, tc::BK)};
      return tc::launch_tc<true, true>(c.st, dy, pb, te, M, lin.in, lin.out, false, "linear_bwd_x_tc_packed", PROF_GEMM_BWD_X);
    }
    return tc::launch_tc<true, false>(c.st, dy, weight_view(lin), te, M, lin.in, lin.out, false, "linear_bwd_x_tc", PROF_GEMM_BWD_X);
  }
  return launch_gemm<true, false>(c.st, dy, weight_view(lin), epi, M, lin.in, lin.out, false, "linear_bwd_x", PROF_GEMM_BWD_X);
}
// prev.g = relu_mask(prev) ? (dy W + add) : 0, with the BN-backward reduction of `prev` fused in the epilogue.
int bwd_x_masked(const Ctx& c, const DyView& dy, const Lin& lin, int M, const Blk& pb, BlkState& ps, const float* add, int ldadd) {
  EpiMaskReduce epi; memset(&epi, 0, sizeof(epi));
  epi.G = ps.g; epi.ldg = pb.lin.out; epi.add = add; epi.ldadd = ldadd;
  epi.yprev = ps.y; epi.ldy = pb.lin.out;
  if (pb.has_bn) { epi.scale = ps.scale; epi.shift = ps.shift; epi.mean = ps.mean; epi.rstd = ps.rstd; }
  epi.fin = blk_fin(c, pb, ps);
  SLN_CHECK_ARG(lin.in == pb.lin.out, "internal: masked backward expects matching widths (%d vs %d)", lin.in, pb.lin.out);
  side_before_write(c, ps.g);
  if (use_skinny(4) && lin.out <= 32 && dy.p == nullptr) {
    SmallKSrc src{dy.g, dy.ldg, lin.out, lin.W, lin.in, add, ldadd};
    return launch_prep(c.st, src, blk_act(pb, ps), ps.g, pb.lin.out, epi.fin, M, lin.in, "skinny_bwd_x_masked");
  }
  if (use_tc(M, lin.in, lin.out, SITE_BWD_XM) && dy.vec_ok() && weight_view(lin).vec_ok()) {
    tc::TcEpiMaskReduce te{epi.G, epi.ldg, epi.add, epi.ldadd, epi.yprev, epi.ldy, epi.scale, epi.shift, epi.mean, epi.rstd, epi.fin};
    if (lin.pb) {
      tc::PackedB pb{lin.pb, ceil_div(lin.out, tc::BK)};
      return tc::launch_tc<true, true>(c.st, dy, pb, te, M, lin.in, lin.out, false, "linear_bwd_x_masked_tc_packed", PROF_GEMM_BWD_X);
    }
    return tc::launch_tc<true, false>(c.st, dy, weight_view(lin), te, M, lin.in, lin.out, false, "linear_bwd_x_masked_tc", PROF_GEMM_BWD_X);
  }
  return launch_gemm<true, false>(c.st, dy, weight_view(lin), epi, M, lin.in, lin.out, false, "linear_bwd_x_masked", PROF_GEMM_BWD_X);
}

// sln_vae_desc.graph_ws: the decoder borrows the graph (CSR, int32 index arrays) that this step's encoder call built in ITS workspace
inline bool borrow_graph(const Dims& dm, const sln_vae_desc* d, int O, int T, NetPlan& p) {
  if (!d->graph_ws) return false;
  static thread_local NetPlan enc;
  make_plan(dm, O, T, 0, const_cast<void*>(d->graph_ws), &enc);
  p.g = enc.g;
  p.objs32 = enc.objs32; p.attrs32 = enc.attrs32;
  return true;
}

int graph_prep(const Ctx& c, NetPlan& p, const int64_t* triples_or_edges, int stride3, bool clear_err = true) {
  const Graph& g = p.g;
  SLN_CUDA_TRY(cudaMemsetAsync(p.deg, 0, sizeof(int) * g.O, c.st));
  SLN_CUDA_TRY(cudaMemsetAsync(g.cursor, 0, sizeof(int) * g.O, c.st));
  if (clear_err) SLN_CUDA_TRY(cudaMemsetAsync(p.err, 0, sizeof(int), c.st));
  (void)stride3;
  if (g.T > 0) {
    k_split_triples<<<ceil_div(g.T, 256), 256, 0, c.st>>>((const long long*)triples_or_edges, g.T, g.O, c.dm.num_preds, g.s_idx, g.p_idx, g.o_idx, p.deg, p.err);
    SLN_TRY(check_launch("split_triples"));
  }
  k_scan_deg<<<1, 1024, 0, c.st>>>(p.deg, g.O, g.row_ptr, g.cnt);
  SLN_TRY(check_launch("scan_deg"));
  if (g.T > 0) {
    k_fill_csr<<<ceil_div(g.T, 256), 256, 0, c.st>>>(g.s_idx, g.o_idx, g.T, g.row_ptr, g.cursor, g.ent);
    SLN_TRY(check_launch("fill_csr"));
    k_sort_rows<<<ceil_div(g.O, 128), 128, 0, c.st>>>(g.row_ptr, g.O, g.ent);
    SLN_TRY(check_launch("sort_rows"));
  }
  return SLN_OK;
}

// edges [T,2] variant for the standalone layer: same kernel with a 2-wide stride
__global__ void k_split_edges(const long long* __restrict__ edges, int T, int O, int* s_idx, int* o_idx, int* deg, int* err) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  long long s = edges[(size_t)t * 2 + 0], o = edges[(size_t)t * 2 + 1];
  if (s < 0 || s >= O || o < 0 || o >= O) { atomicOr(err, 8); s = 0; o = 0; }
  s_idx[t] = (int)s; o_idx[t] = (int)o;
  atomicAdd(deg + s, 1);
  atomicAdd(deg + o, 1);
}
int graph_prep_edges(cudaStream_t st, Graph& g, int* deg, int* err, const int64_t* edges) {
  SLN_CUDA_TRY(cudaMemsetAsync(deg, 0, sizeof(int) * g.O, st));
  SLN_CUDA_TRY(cudaMemsetAsync(g.cursor, 0, sizeof(int) * g.O, st));
  SLN_CUDA_TRY(cudaMemsetAsync(err, 0, sizeof(int), st));
  if (g.T > 0) {
    k_split_edges<<<ceil_div(g.T, 256), 256, 0, st>>>((const long long*)edges, g.T, g.O, g.s_idx, g.o_idx, deg, err);
    SLN_TRY(check_launch("split_edges"));
  }
  k_scan_deg<<<1, 1024, 0, st>>>(deg, g.O, g.row_ptr, g.cnt);
  SLN_TRY(check_launch("scan_deg"));
  if (g.T > 0) {
    k_fill_csr<<<ceil_div(g.T, 256), 256, 0, st>>>(g.s_idx, g.o_idx, g.T, g.row_ptr, g.cursor, g.ent);
    SLN_TRY(check_launch("fill_csr"));
    k_sort_rows<<<ceil_div(g.O, 128), 128, 0, st>>>(g.row_ptr, g.O, g.ent);
    SLN_TRY(check_launch("sort_rows"));
  }
  return SLN_OK;
}

int gather_rows(const Ctx& c, const float* table, int ldt, const int* idx, int n, int width, float* out, int ldo, int off) {
  if (n <= 0) return SLN_OK;
  dim3 blk(32, 8);
  k_gather_rows<<<ceil_div(n, 8), blk, 0, c.st>>>(table, ldt, idx, n, width, out, ldo, off);
  return check_launch("gather_rows");
}
int embed_bwd(const Ctx& c, const float* a, int lda, const float* b, int ldb, const int* idx, int n, int width, float* tg, int rows) {
  return launch_embed_bwd(side_fork_leaf(c), a, lda, b, ldb, idx, n, width, tg, rows);   // tail of a backward call: see Side::leaf
}

// ---------------------------------------------------------------- one GraphTripleConv layer
int gconv_fwd(const Ctx& c, const Graph& g, const Blk* blk, BlkState* st, float* pooled, const MatView& obj_in,
              const MatView& pred_in, MatView* obj_out, MatView* pred_out) {
  const int D = c.dm.D, H = c.dm.H, O = g.O, T = g.T;
  GatherCat gc{obj_in, pred_in, g.s_idx, g.o_idx, D, T, 3 * D};
  SLN_TRY(block_fwd(c, gc, T, blk[0], st[0]));
  // net1's second Linear writes [new_s | new_p | new_o] (2H + D columns).  At configs[1] that is 5 x 31 = 155 tiles of 128 x 128 on
  // 148 SMs: a 7-CTA second wave doubles the kernel's time.  Only new_s / new_o feed the pooling and net2; new_p is needed by the
  // NEXT layer.  So the s / o column tiles run on the chain (124 CTAs: one wave) and the p tiles as a parallel branch that overlaps
  // the pooling and net2 and rejoins at the end of the layer.  Per-column-block BatchNorm state makes the two launches independent.
  MatView a1 = block_out(blk[0], st[0]);
  const int tm = ceil_div(T, tc::BM);
  const bool split = c.side && blk[1].lin.pf && use_tc(T, blk[1].lin.out, blk[1].lin.in, SITE_FWD) && a1.vec_ok() && H % 128 == 0 && D % 128 == 0 &&
                     blk[1].lin.out == 2 * H + D && tm * ((2 * H + D) / 128) > kNumSMs && tm * (2 * H / 128) <= kNumSMs;
  if (split) {
    const tc::TileSel so{2 * H / 128, H / 128, D / 128, 128}, sp{D / 128, 0, H / 128, 128};
    SLN_TRY(block_fwd(c, a1, T, blk[1], st[1], nullptr, 0, &so));
    Ctx cp = c;
    cp.side = false;
    cp.st = side_fork_leaf(c);
    SLN_TRY(block_fwd(cp, a1, T, blk[1], st[1], nullptr, 0, &sp, true));
  } else {
    SLN_TRY(block_fwd(c, a1, T, blk[1], st[1]));
  }
  MatView a2 = block_out(blk[1], st[1]);
  if (O > 0) {
    ProfScope prof(c.st, PROF_POOL, pool_bytes(O, T, H));
    launch_pool_fwd(c.st, a2, g.row_ptr, g.ent, O, H, D, pooled);
    SLN_TRY(check_launch("pool_fwd"));
  }
  SLN_TRY(block_fwd(c, make_view(pooled, H, O, H), O, blk[2], st[2]));
  SLN_TRY(block_fwd(c, block_out(blk[2], st[2]), O, blk[3], st[3]));
  *obj_out = block_out(blk[3], st[3]);
  *pred_out = slice_cols(a2, H, D);
  if (split) SLN_TRY(side_join_leaf(c));
  return SLN_OK;
}

// Backward of one layer.  Precondition: st[3].g holds the masked gradient of the layer's obj output and its (p,q,r)
// are finalised.  dpred_next: gradient w.r.t. the layer's pred output (null = 0).  Writes dcat [T,3D] = gradient w.r.t.
// the virtual [obj[s] | pred | obj[o]] input.
int gconv_bwd(const Ctx& c, const Graph& g, const Blk* blk, BlkState* st, const float* pooled, float* dpooled, const MatView& obj_in,
              const MatView& pred_in, const float* dpred_next, int ld_dpred, float* dcat) {
  const int D = c.dm.D, H = c.dm.H, O = g.O, T = g.T;
  // net2, second Linear
  DyView dy4 = blk_dy(c, blk[3], st[3]);
  SLN_TRY(bwd_w(c, dy4, block_out(blk[2], st[2]), blk[3].lin, O));
  SLN_TRY(bwd_x_masked(c, dy4, blk[3].lin, O, blk[2], st[2], nullptr, 0));
  // net2, first Linear
  DyView dy3 = blk_dy(c, blk[2], st[2]);
  SLN_TRY(bwd_w(c, dy3, make_view(pooled, H, O, H), blk[2].lin, O));
  SLN_TRY(bwd_x_plain(c, dy3, blk[2].lin, O, dpooled, H));
  // pooling backward + BN-backward reduction of net1's second Linear
  PoolBwdSrc src{dpooled, g.cnt, g.s_idx, g.o_idx, dpred_next, ld_dpred, H, D};
  side_before_write(c, st[1].g);
  SLN_TRY(launch_prep(c.st, src, blk_act(blk[1], st[1]), st[1].g, 2 * H + D, blk_fin(c, blk[1], st[1]), T, 2 * H + D, "pool_bwd_prep"));
  // net1, second Linear
  DyView dy2 = blk_dy(c, blk[1], st[1]);
  SLN_TRY(bwd_w(c, dy2, block_out(blk[0], st[0]), blk[1].lin, T));
  SLN_TRY(bwd_x_masked(c, dy2, blk[1].lin, T, blk[0], st[0], nullptr, 0));
  // net1, first Linear
  DyView dy1 = blk_dy(c, blk[0], st[0]);
  GatherCat gc{obj_in, pred_in, g.s_idx, g.o_idx, D, T, 3 * D};
  SLN_TRY(bwd_w(c, dy1, gc, blk[0].lin, T));
  SLN_TRY(bwd_x_plain(c, dy1, blk[0].lin, T, dcat, 3 * D));
  return SLN_OK;
}

// gradient of the node inputs from dcat: either masked against the previous layer's activation (+ its BN reduction)
// or plain (first layer: the inputs are embeddings).
int node_gather(const Ctx& c, const Graph& g, const float* dcat, const Blk* pblk, BlkState* pst, float* plain_out) {
  const int D = c.dm.D;
  NodeGatherSrc src{dcat, 3 * D, D, g.row_ptr, g.ent};
  if (pblk) side_before_write(c, pst->g);
  if (pblk) return launch_prep(c.st, src, blk_act(*pblk, *pst), pst->g, D, blk_fin(c, *pblk, *pst), g.O, D, "node_gather_prep");
  ActInfo none; memset(&none, 0, sizeof(none));
  BnBwdFin nofin; memset(&nofin, 0, sizeof(nofin));
  return launch_prep(c.st, src, none, plain_out, D, nofin, g.O, D, "node_gather_plain");
}

int gconv_net_fwd(const Ctx& c, NetPlan& p, Blk (*blk)[4], MatView obj, MatView pred, MatView* obj_f) {
  for (int l = 0; l < c.dm.L; ++l) {
    MatView no, np;
    SLN_TRY(gconv_fwd(c, p.g, blk[l], p.st[l], p.pooled[l], obj, pred, &no, &np));
    obj = no; pred = np;
  }
  *obj_f = obj;
  return SLN_OK;
}

MatView layer_obj_in(const Ctx& c, NetPlan& p, Blk (*blk)[4], int l) {
  if (l == 0) return make_view(p.obj0, c.dm.D, p.g.O, c.dm.D);
  return block_out(blk[l - 1][3], p.st[l - 1][3]);
}
MatView layer_pred_in(const Ctx& c, NetPlan& p, Blk (*blk)[4], int l) {
  if (l == 0) return make_view(p.pred0, c.dm.D, p.g.T, c.dm.D);
  return slice_cols(block_out(blk[l - 1][1], p.st[l - 1][1]), c.dm.H, c.dm.D);
}

// Precondition: p.st[L-1][3].g / (p,q,r) ready.  On return p.dobj0 [O,D] = gradient of the layer-0 node inputs,
// *dpred0 points at the [T, ld 3D] slice holding the gradient of the layer-0 predicate inputs.
int gconv_net_bwd(const Ctx& c, NetPlan& p, Blk (*blk)[4], const float** dpred0, int* ld_dpred0) {
  const float* dpred_next = nullptr;
  const int D = c.dm.D;
  for (int l = c.dm.L - 1; l >= 0; --l) {
    float* dcat = p.sc.dcat[l & 1];
    SLN_TRY(gconv_bwd(c, p.g, blk[l], p.st[l], p.pooled[l], p.sc.dpooled, layer_obj_in(c, p, blk, l), layer_pred_in(c, p, blk, l),
                      dpred_next, 3 * D, dcat));
    if (l > 0) SLN_TRY(node_gather(c, p.g, dcat, &blk[l - 1][3], &p.st[l - 1][3], nullptr));
    else SLN_TRY(node_gather(c, p.g, dcat, nullptr, nullptr, p.dobj0));
    dpred_next = dcat + D;
  }
  *dpred0 = dpred_next; *ld_dpred0 = 3 * D;
  return SLN_OK;
}

int check_ws(const NetPlan& p, const void* ws, size_t ws_bytes) {
  SLN_CHECK_ARG(ws != nullptr, "null workspace");
  SLN_CHECK_ARG((uintptr_t)ws % 256 == 0, "workspace must be 256-byte aligned");
  if (ws_bytes < p.bytes) { set_error("workspace too small: %zu < %zu bytes", ws_bytes, p.bytes); return SLN_EWORKSPACE; }
  return SLN_OK;
}
int to_i32(const Ctx& c, const int64_t* src, int n, int* dst, int limit, int* err, int bit) {
  if (n <= 0) return SLN_OK;
  k_i64_to_i32<<<ceil_div(n, 256), 256, 0, c.st>>>((const long long*)src, n, dst, limit, err, bit);
  return check_launch("i64_to_i32");
}
int check_dims(int64_t O, int64_t T) {
  SLN_CHECK_ARG(O >= 1 && O < (1ll << 30), "O out of range: %lld", (long long)O);
  SLN_CHECK_ARG(T >= 0 && T < (1ll << 30), "T out of range: %lld", (long long)T);
  return SLN_OK;
}

}  // namespace
}  // namespace sln

using namespace sln;

extern "C" {

size_t sln_vae_packed_bytes(const sln_vae_desc* d) {
  Dims dm;
  if (make_dims(d, &dm)) return 0;
  std::vector<const void*> fake((size_t)count_params(dm), (const void*)(uintptr_t)256);
  static thread_local Model m;
  size_t floats = 0;
  if (parse_model(dm, fake.data(), nullptr, nullptr, &m, nullptr, nullptr, &floats)) return 0;
  return floats * sizeof(float);
}

int launch_pack_jobs(cudaStream_t st, const std::vector<PackJob>& jobs, size_t floats) {
  ProfScope prof(st, PROF_MISC, 12.0 * (double)floats);
  for (size_t j0 = 0; j0 < jobs.size(); j0 += kPackJobsPerLaunch) {
    PackJobs pj; memset(&pj, 0, sizeof(pj));
    pj.n = (int)std::min((size_t)kPackJobsPerLaunch, jobs.size() - j0);
    int units = 0;
    for (int j = 0; j < pj.n; ++j) {
      pj.job[j] = jobs[j0 + j];
      pj.unit_start[j] = units;
      units += (int)(ceil_div64(pj.job[j].N, 128) * 4 * ceil_div64(pj.job[j].K, 32));
    }
    pj.unit_start[pj.n] = units;
    k_pack_jobs<<<units, 256, 0, st>>>(pj);
    SLN_TRY(check_launch("pack_weights"));
  }
  return SLN_OK;
}

int sln_vae_pack_weights(const sln_vae_desc* d, const void* const* params, void* packed, size_t packed_bytes, void* stream) {
  Dims dm;
  SLN_TRY(make_dims(d, &dm));
  SLN_CHECK_ARG(params && packed && (uintptr_t)packed % 16 == 0, "null or misaligned pointer");
  static thread_local Model m;
  std::vector<PackJob> jobs;
  size_t floats = 0;
  SLN_TRY(parse_model(dm, params, nullptr, nullptr, &m, packed, &jobs, &floats));
  if (packed_bytes < floats * sizeof(float)) { set_error("packed weight buffer too small: %zu < %zu bytes", packed_bytes, floats * sizeof(float)); return SLN_EWORKSPACE; }
  return launch_pack_jobs((cudaStream_t)stream, jobs, floats);
}

int sln_vae_num_params(const sln_vae_desc* d) { Dims dm; if (make_dims(d, &dm)) return -1; return count_params(dm); }
int sln_vae_num_bn(const sln_vae_desc* d) { Dims dm; if (make_dims(d, &dm)) return -1; return count_bn(dm); }

size_t sln_vae_workspace_bytes(const sln_vae_desc* d, int64_t O, int64_t T, int which) {
  Dims dm;
  if (make_dims(d, &dm) || check_dims(O, T)) return 0;
  static thread_local NetPlan p;
  make_plan(dm, (int)O, (int)T, which, nullptr, &p);
  return p.bytes;
}

size_t sln_bn_sync_recv_bytes(int32_t world) {
  return (size_t)4 * SLN_BN_SYNC_SLOTS * (size_t)(world > 0 ? world : 1) * SLN_BN_SYNC_COLS * 3 * sizeof(double);
}
size_t sln_bn_sync_flag_bytes(void) { return (size_t)4 * SLN_BN_SYNC_SLOTS * sizeof(uint32_t); }

int64_t sln_vae_index_flag_offset(const sln_vae_desc* d, int64_t O, int64_t T, int which) {
  Dims dm;
  if (make_dims(d, &dm) || check_dims(O, T)) return -1;
  static thread_local NetPlan p;
  make_plan(dm, (int)O, (int)T, which, nullptr, &p);
  return (int64_t)((uintptr_t)p.err - 256);   // dry-run arenas hand out fake addresses 256 + offset (common.cuh Arena::take)
}

int sln_vae_encoder_fwd(const sln_vae_desc* d, const void* const* params, void* const* bn_bufs, const int64_t* objs,
                        const int64_t* triples, const float* boxes, const int64_t* angles, const int64_t* attributes, int64_t O64,
                        int64_t T64, float* mu, float* logvar, void* ws, size_t ws_bytes, void* stream) {
  Ctx c; c.st = (cudaStream_t)stream;
  SLN_TRY(make_dims(d, &c.dm)); SLN_TRY(check_dims(O64, T64));
  SLN_CHECK_ARG(objs && boxes && angles && attributes && mu && logvar && (triples || T64 == 0), "null input/output pointer");
  const Dims& dm = c.dm; const int O = (int)O64, T = (int)T64;
  static thread_local Model m; static thread_local NetPlan p;
  SLN_TRY(parse_model(dm, params, nullptr, bn_bufs, &m, d->packed_weights));
  make_plan(dm, O, T, 0, ws, &p);
  ctx_sync(c, d, p.cp.base, 0, 0);
  SLN_TRY(check_ws(p, ws, ws_bytes));
  SLN_CUDA_TRY(cudaMemsetAsync(p.cp.base, 0, sizeof(unsigned) * kCounterCap, c.st));
  SLN_CUDA_TRY(cudaMemsetAsync(p.err, 0, sizeof(int), c.st));
  // Two independent preparations run as parallel branches and meet before the first layer: the CSR build (chain) and the int32
  // index copies + embedding gathers of the node features (leaf stream).  (Re-splitting the weight images on a third branch was
  // measured too: no gain — the bandwidth kernel fills every SM slot and the small kernels queue behind it.)
  side_begin(c);
  Ctx ce = c;
  ce.side = false;
  if (c.side) ce.st = side_fork_leaf(c);
  SLN_TRY(to_i32(ce, objs, O, p.objs32, d->num_objs, p.err, SLN_IDX_OBJS));
  SLN_TRY(to_i32(ce, attributes, O, p.attrs32, d->num_attrs, p.err, SLN_IDX_ATTRS));
  SLN_TRY(to_i32(ce, angles, O, p.angles32, d->n_angle, p.err, SLN_IDX_ANGLES));
  // obj_vecs = [obj_emb | attr_emb | box_linear | angle_emb]   (Sg2ScVAE_model.py:121-129)
  SLN_TRY(gather_rows(ce, m.emb[0], dm.obj_w, p.objs32, O, dm.obj_w, p.obj0, dm.D, 0));
  SLN_TRY(gather_rows(ce, m.emb[1], dm.attr_w, p.attrs32, O, dm.attr_w, p.obj0, dm.D, dm.obj_w));
  {
    dim3 blk(32, 8);
    k_small_linear<<<ceil_div(O, 8), blk, 0, ce.st>>>(boxes, O, dm.box_dim, m.box_emb.lin.W, m.box_emb.lin.b, dm.box_w, p.obj0, dm.D, dm.obj_w + dm.attr_w);
    SLN_TRY(check_launch("box_embeddings"));
  }
  SLN_TRY(gather_rows(ce, m.emb[2], dm.ang_w, p.angles32, O, dm.ang_w, p.obj0, dm.D, dm.obj_w + dm.attr_w + dm.box_w));
  SLN_TRY(graph_prep(c, p, triples, 1, false));
  SLN_TRY(gather_rows(c, m.emb[3], dm.D, p.g.p_idx, T, dm.D, p.pred0, dm.D, 0));
  SLN_TRY(side_join_leaf(c));
  MatView obj_f;
  SLN_TRY(gconv_net_fwd(c, p, m.enc, make_view(p.obj0, dm.D, O, dm.D), make_view(p.pred0, dm.D, T, dm.D), &obj_f));
  // heads (Sg2ScVAE_model.py:134-143)
  // The box and the angle branch are independent 4-kernel chains on obj_f (disjoint states, disjoint column ranges of mu / logvar):
  // the angle branch runs on the side stream (a parallel branch of a captured graph), joined before the call returns.
  Ctx ca = c;
  ca.side = false;
  if (c.side) ca.st = side_fork(c);
  BlkState dummy; memset(&dummy, 0, sizeof(dummy)); dummy.M = O;
  SLN_TRY(block_fwd(ca, obj_f, O, m.amv[0], p.amv[0]));
  SLN_TRY(block_fwd(ca, block_out(m.amv[0], p.amv[0]), O, m.amv[1], p.amv[1]));
  MatView ha = block_out(m.amv[1], p.amv[1]);
  SLN_TRY(block_fwd(ca, ha, O, m.angle_mean, dummy, mu + dm.box_w, dm.Z));
  SLN_TRY(block_fwd(ca, ha, O, m.angle_var, dummy, logvar + dm.box_w, dm.Z));
  SLN_TRY(block_fwd(c, obj_f, O, m.bmv[0], p.bmv[0]));
  SLN_TRY(block_fwd(c, block_out(m.bmv[0], p.bmv[0]), O, m.bmv[1], p.bmv[1]));
  MatView hb = block_out(m.bmv[1], p.bmv[1]);
  SLN_TRY(block_fwd(c, hb, O, m.box_mean, dummy, mu, dm.Z));
  SLN_TRY(block_fwd(c, hb, O, m.box_var, dummy, logvar, dm.Z));
  return side_end(c);
}

int sln_vae_encoder_bwd(const sln_vae_desc* d, const void* const* params, void* const* grads, const float* boxes, const float* d_mu,
                        const float* d_logvar, int64_t O64, int64_t T64, void* ws, size_t ws_bytes, void* stream) {
  Ctx c; c.st = (cudaStream_t)stream;
  SLN_TRY(make_dims(d, &c.dm)); SLN_TRY(check_dims(O64, T64));
  SLN_CHECK_ARG(grads && boxes && d_mu && d_logvar, "null pointer");
  const Dims& dm = c.dm; const int O = (int)O64, T = (int)T64;
  static thread_local Model m; static thread_local NetPlan p;
  SLN_TRY(parse_model(dm, params, grads, nullptr, &m, d->packed_weights));
  make_plan(dm, O, T, 0, ws, &p);
  ctx_sync(c, d, p.cp.base, 0, 1);
  SLN_TRY(check_ws(p, ws, ws_bytes));
  side_begin(c);
  const int L = dm.L;
  MatView obj_f = block_out(m.enc[L - 1][3], p.st[L - 1][3]);
  MatView hb = block_out(m.bmv[1], p.bmv[1]), ha = block_out(m.amv[1], p.amv[1]);
  MatView hb1 = block_out(m.bmv[0], p.bmv[0]), ha1 = block_out(m.amv[0], p.amv[0]);
  // --- angle branch: independent of the box branch up to the kernel that adds tmpB, so it runs on a leaf stream (its weight
  // gradients fork from there to the side stream as usual) and rejoins the chain below
  Ctx ca = c;
  if (c.side) ca.st = side_fork_leaf(c);
  DyView dmu_a = make_dy(d_mu + dm.box_w, dm.Z, O, dm.ang_w), dlv_a = make_dy(d_logvar + dm.box_w, dm.Z, O, dm.ang_w);
  SLN_TRY(bwd_w(ca, dmu_a, ha, m.angle_mean.lin, O, true));
  SLN_TRY(bwd_w(ca, dlv_a, ha, m.angle_var.lin, O, true));
  SLN_TRY(bwd_x_plain(ca, dmu_a, m.angle_mean.lin, O, p.tmpA2, dm.D));
  SLN_TRY(bwd_x_masked(ca, dlv_a, m.angle_var.lin, O, m.amv[1], p.amv[1], p.tmpA2, dm.D));
  DyView dya2 = blk_dy(ca, m.amv[1], p.amv[1]);
  SLN_TRY(bwd_w(ca, dya2, ha1, m.amv[1].lin, O));
  SLN_TRY(bwd_x_masked(ca, dya2, m.amv[1].lin, O, m.amv[0], p.amv[0], nullptr, 0));
  DyView dya1 = blk_dy(ca, m.amv[0], p.amv[0]);
  SLN_TRY(bwd_w(ca, dya1, obj_f, m.amv[0].lin, O));
  // --- box branch
  DyView dmu_b = make_dy(d_mu, dm.Z, O, dm.box_w), dlv_b = make_dy(d_logvar, dm.Z, O, dm.box_w);
  SLN_TRY(bwd_w(c, dmu_b, hb, m.box_mean.lin, O, true));
  SLN_TRY(bwd_w(c, dlv_b, hb, m.box_var.lin, O, true));
  SLN_TRY(bwd_x_plain(c, dmu_b, m.box_mean.lin, O, p.tmpA, dm.D));
  SLN_TRY(bwd_x_masked(c, dlv_b, m.box_var.lin, O, m.bmv[1], p.bmv[1], p.tmpA, dm.D));
  DyView dyb2 = blk_dy(c, m.bmv[1], p.bmv[1]);
  SLN_TRY(bwd_w(c, dyb2, hb1, m.bmv[1].lin, O));
  SLN_TRY(bwd_x_masked(c, dyb2, m.bmv[1].lin, O, m.bmv[0], p.bmv[0], nullptr, 0));
  DyView dyb1 = blk_dy(c, m.bmv[0], p.bmv[0]);
  SLN_TRY(bwd_w(c, dyb1, obj_f, m.bmv[0].lin, O));
  SLN_TRY(bwd_x_plain(c, dyb1, m.bmv[0].lin, O, p.tmpB, dm.D));
  SLN_TRY(side_join_leaf(c));
  // both branches meet at the last gconv layer's node output
  SLN_TRY(bwd_x_masked(c, dya1, m.amv[0].lin, O, m.enc[L - 1][3], p.st[L - 1][3], p.tmpB, dm.D));
  // --- graph conv stack
  const float* dpred0; int ldp;
  SLN_TRY(gconv_net_bwd(c, p, m.enc, &dpred0, &ldp));
  // --- embeddings (Sg2ScVAE_model.py:121-129)
  SLN_TRY(embed_bwd(c, p.dobj0, dm.D, nullptr, 0, p.objs32, O, dm.obj_w, m.demb[0], d->num_objs));
  SLN_TRY(embed_bwd(c, p.dobj0 + dm.obj_w, dm.D, nullptr, 0, p.attrs32, O, dm.attr_w, m.demb[1], d->num_attrs));
  {
    const int off = dm.obj_w + dm.attr_w;
    // dW [box_w, box_dim] = dy^T boxes: the SMALL side is the input (6), so boxes play the role of the narrow operand
    if (m.box_emb.lin.dW)
      SLN_TRY(launch_skinny_bwd_w(side_fork_leaf(c), boxes, dm.box_dim, dm.box_dim, make_view(p.dobj0 + off, dm.D, O, dm.box_w), O, dm.box_w,
                                  m.box_emb.lin.dW, dm.box_dim, true, nullptr));
    if (m.box_emb.lin.db) SLN_TRY(launch_embed_bwd(side_fork_leaf(c), p.dobj0 + off, dm.D, nullptr, 0, nullptr, O, dm.box_w, m.box_emb.lin.db, 1));
  }
  SLN_TRY(embed_bwd(c, p.dobj0 + dm.obj_w + dm.attr_w + dm.box_w, dm.D, nullptr, 0, p.angles32, O, dm.ang_w, m.demb[2], d->n_angle));
  SLN_TRY(embed_bwd(c, dpred0, ldp, nullptr, 0, p.g.p_idx, T, dm.D, m.demb[3], d->num_preds));
  return side_end(c);
}

int sln_vae_decoder_fwd(const sln_vae_desc* d, const void* const* params, void* const* bn_bufs, const float* z, const int64_t* objs,
                        const int64_t* triples, const int64_t* attributes, int64_t O64, int64_t T64, float* boxes_pred,
                        float* angles_pred, void* ws, size_t ws_bytes, void* stream) {
  Ctx c; c.st = (cudaStream_t)stream;
  SLN_TRY(make_dims(d, &c.dm)); SLN_TRY(check_dims(O64, T64));
  SLN_CHECK_ARG(z && objs && attributes && boxes_pred && angles_pred && (triples || T64 == 0), "null input/output pointer");
  const Dims& dm = c.dm; const int O = (int)O64, T = (int)T64;
  static thread_local Model m; static thread_local NetPlan p;
  SLN_TRY(parse_model(dm, params, nullptr, bn_bufs, &m, d->packed_weights));
  make_plan(dm, O, T, 1, ws, &p);
  ctx_sync(c, d, p.cp.base, 1, 0);
  SLN_TRY(check_ws(p, ws, ws_bytes));
  SLN_CUDA_TRY(cudaMemsetAsync(p.cp.base, 0, sizeof(unsigned) * kCounterCap, c.st));
  if (borrow_graph(dm, d, O, T, p)) {                    // the encoder validated the indices; this call's flag stays clear
    SLN_CUDA_TRY(cudaMemsetAsync(p.err, 0, sizeof(int), c.st));
  } else {
    SLN_TRY(graph_prep(c, p, triples, 1));
    SLN_TRY(to_i32(c, objs, O, p.objs32, d->num_objs, p.err, SLN_IDX_OBJS));
    SLN_TRY(to_i32(c, attributes, O, p.attrs32, d->num_attrs, p.err, SLN_IDX_ATTRS));
  }
  // obj_vecs = [obj_emb_dc | attr_emb_dc | z]   (Sg2ScVAE_model.py:150-159, decoder_cat)
  SLN_TRY(gather_rows(c, m.emb[4], dm.obj_w, p.objs32, O, dm.obj_w, p.obj0, dm.D, 0));
  SLN_TRY(gather_rows(c, m.emb[5], dm.attr_w, p.attrs32, O, dm.attr_w, p.obj0, dm.D, dm.obj_w));
  SLN_TRY(gather_rows(c, z, dm.Z, nullptr, O, dm.Z, p.obj0, dm.D, dm.obj_w + dm.attr_w));
  SLN_TRY(gather_rows(c, m.emb[6], dm.D, p.g.p_idx, T, dm.D, p.pred0, dm.D, 0));
  side_begin(c);
  MatView obj_f;
  SLN_TRY(gconv_net_fwd(c, p, m.dec, make_view(p.obj0, dm.D, O, dm.D), make_view(p.pred0, dm.D, T, dm.D), &obj_f));
  // box_net on [obj_f | attr_vecs], angle_net on obj_f   (Sg2ScVAE_model.py:166-171)
  Concat2 cat{obj_f, make_view(p.obj0 + dm.obj_w, dm.D, O, dm.attr_w), O, dm.D + dm.attr_w};
  // box_net runs on the side stream while angle_net + log-softmax run on the chain (independent branches on obj_f)
  Ctx cb = c;
  cb.side = false;
  if (c.side) cb.st = side_fork(c);
  BlkState dummy; memset(&dummy, 0, sizeof(dummy)); dummy.M = O;
  SLN_TRY(block_fwd(cb, cat, O, m.box_net[0], p.box_net0));
  SLN_TRY(block_fwd(cb, block_out(m.box_net[0], p.box_net0), O, m.box_net[1], dummy, boxes_pred, dm.box_dim));
  SLN_TRY(block_fwd(c, obj_f, O, m.angle_net[0], p.angle_net0));
  SLN_TRY(block_fwd(c, block_out(m.angle_net[0], p.angle_net0), O, m.angle_net[1], dummy, p.logits, dm.n_angle));
  k_log_softmax_fwd<<<ceil_div(O, 8), 256, 0, c.st>>>(p.logits, O, dm.n_angle, angles_pred);
  SLN_TRY(check_launch("log_softmax_fwd"));
  SLN_CUDA_TRY(cudaMemcpyAsync(p.logp, angles_pred, sizeof(float) * (size_t)O * dm.n_angle, cudaMemcpyDeviceToDevice, c.st));
  return side_end(c);
}

int sln_vae_decoder_bwd(const sln_vae_desc* d, const void* const* params, void* const* grads, const float* d_boxes, const float* d_angles,
                        int angles_are_logits, float* d_z, int64_t O64, int64_t T64, void* ws, size_t ws_bytes, void* stream) {
  Ctx c; c.st = (cudaStream_t)stream;
  SLN_TRY(make_dims(d, &c.dm)); SLN_TRY(check_dims(O64, T64));
  SLN_CHECK_ARG(grads && d_boxes && d_angles, "null pointer");
  const Dims& dm = c.dm; const int O = (int)O64, T = (int)T64;
  static thread_local Model m; static thread_local NetPlan p;
  SLN_TRY(parse_model(dm, params, grads, nullptr, &m, d->packed_weights));
  make_plan(dm, O, T, 1, ws, &p);
  ctx_sync(c, d, p.cp.base, 1, 1);
  SLN_TRY(check_ws(p, ws, ws_bytes));
  (void)borrow_graph(dm, d, O, T, p);
  side_begin(c);
  const int L = dm.L;
  MatView obj_f = block_out(m.dec[L - 1][3], p.st[L - 1][3]);
  // angle_net's output layer (+ log-softmax backward) is independent of box_net's two layers: a parallel branch (leaf stream)
  Ctx ca = c;
  if (c.side) ca.st = side_fork_leaf(c);
  const float* dlogits = d_angles;
  if (!angles_are_logits) {
    k_log_softmax_bwd<<<ceil_div(O, 8), 256, 0, ca.st>>>(d_angles, p.logp, O, dm.n_angle, p.dlogits);
    SLN_TRY(check_launch("log_softmax_bwd"));
    dlogits = p.dlogits;
  }
  DyView dyl = make_dy(dlogits, dm.n_angle, O, dm.n_angle);
  SLN_TRY(bwd_w(ca, dyl, block_out(m.angle_net[0], p.angle_net0), m.angle_net[1].lin, O, true));
  SLN_TRY(bwd_x_masked(ca, dyl, m.angle_net[1].lin, O, m.angle_net[0], p.angle_net0, nullptr, 0));
  // box_net
  DyView dyb = make_dy(d_boxes, dm.box_dim, O, dm.box_dim);
  SLN_TRY(bwd_w(c, dyb, block_out(m.box_net[0], p.box_net0), m.box_net[1].lin, O, true));
  SLN_TRY(bwd_x_masked(c, dyb, m.box_net[1].lin, O, m.box_net[0], p.box_net0, nullptr, 0));
  DyView dyb0 = blk_dy(c, m.box_net[0], p.box_net0);
  Concat2 cat{obj_f, make_view(p.obj0 + dm.obj_w, dm.D, O, dm.attr_w), O, dm.D + dm.attr_w};
  SLN_TRY(bwd_w(c, dyb0, cat, m.box_net[0].lin, O));
  const int ldc = dm.D + dm.attr_w;
  SLN_TRY(bwd_x_plain(c, dyb0, m.box_net[0].lin, O, p.tmpC, ldc));
  SLN_TRY(side_join_leaf(c));
  DyView dya0 = blk_dy(c, m.angle_net[0], p.angle_net0);
  SLN_TRY(bwd_w(c, dya0, obj_f, m.angle_net[0].lin, O));
  SLN_TRY(bwd_x_masked(c, dya0, m.angle_net[0].lin, O, m.dec[L - 1][3], p.st[L - 1][3], p.tmpC, ldc));
  // graph conv stack
  const float* dpred0; int ldp;
  SLN_TRY(gconv_net_bwd(c, p, m.dec, &dpred0, &ldp));
  // embeddings + z
  SLN_TRY(embed_bwd(c, p.dobj0, dm.D, nullptr, 0, p.objs32, O, dm.obj_w, m.demb[4], d->num_objs));
  SLN_TRY(embed_bwd(c, p.dobj0 + dm.obj_w, dm.D, p.tmpC + dm.D, ldc, p.attrs32, O, dm.attr_w, m.demb[5], d->num_attrs));
  SLN_TRY(embed_bwd(c, dpred0, ldp, nullptr, 0, p.g.p_idx, T, dm.D, m.demb[6], d->num_preds));
  if (d_z) SLN_TRY(gather_rows(c, p.dobj0 + dm.obj_w + dm.attr_w, dm.D, nullptr, O, dm.Z, d_z, dm.Z, 0));
  return side_end(c);
}

// ---------------------------------------------------------------- standalone GraphTripleConv layer
static int parse_layer(const Dims& dm, const void* const* lp, void* const* lg, void* const* lbn, Blk* blk) {
  SLN_CHECK_ARG(lp != nullptr, "null layer parameter table");
  TableReader tr{lp, lg, lbn, 0, 0, dm.norm};
  take_gconv(tr, dm, blk);
  return SLN_OK;
}

int sln_gconv_layer_fwd(const sln_vae_desc* d, const void* const* layer_params, void* const* layer_bn_bufs, const float* obj_vecs,
                        const float* pred_vecs, const int64_t* edges, int64_t O64, int64_t T64, float* new_obj, float* new_pred,
                        void* ws, size_t ws_bytes, void* stream) {
  Ctx c; c.st = (cudaStream_t)stream;
  SLN_TRY(make_dims(d, &c.dm)); SLN_TRY(check_dims(O64, T64));
  SLN_CHECK_ARG(obj_vecs && new_obj && (T64 == 0 || (pred_vecs && edges && new_pred)), "null pointer");
  const Dims& dm = c.dm; const int O = (int)O64, T = (int)T64;
  Blk blk[4]; static thread_local NetPlan p;
  SLN_TRY(parse_layer(dm, layer_params, nullptr, layer_bn_bufs, blk));
  make_plan(dm, O, T, 2, ws, &p);
  SLN_TRY(check_ws(p, ws, ws_bytes));
  SLN_CUDA_TRY(cudaMemsetAsync(p.cp.base, 0, sizeof(unsigned) * kCounterCap, c.st));
  SLN_TRY(graph_prep_edges(c.st, p.g, p.deg, p.err, edges));
  MatView no, np;
  SLN_TRY(gconv_fwd(c, p.g, blk, p.st[0], p.pooled[0], make_view(obj_vecs, dm.D, O, dm.D), make_view(pred_vecs, dm.D, T, dm.D), &no, &np));
  dim3 b(32, 8);
  k_materialize<<<ceil_div(O, 8), b, 0, c.st>>>(no, new_obj, dm.D);
  SLN_TRY(check_launch("materialize_obj"));
  if (T > 0) {
    k_materialize<<<ceil_div(T, 8), b, 0, c.st>>>(np, new_pred, dm.D);
    SLN_TRY(check_launch("materialize_pred"));
  }
  return SLN_OK;
}

int sln_gconv_layer_bwd(const sln_vae_desc* d, const void* const* layer_params, void* const* layer_grads, const float* obj_vecs,
                        const float* pred_vecs, const float* d_new_obj, const float* d_new_pred, int64_t O64, int64_t T64, float* d_obj,
                        float* d_pred, void* ws, size_t ws_bytes, void* stream) {
  Ctx c; c.st = (cudaStream_t)stream;
  SLN_TRY(make_dims(d, &c.dm)); SLN_TRY(check_dims(O64, T64));
  SLN_CHECK_ARG(obj_vecs && d_new_obj && d_obj, "null pointer");
  const Dims& dm = c.dm; const int O = (int)O64, T = (int)T64;
  Blk blk[4]; static thread_local NetPlan p;
  SLN_TRY(parse_layer(dm, layer_params, layer_grads, nullptr, blk));
  make_plan(dm, O, T, 2, ws, &p);
  SLN_TRY(check_ws(p, ws, ws_bytes));
  side_begin(c);
  // masked gradient of the node output + BN-backward reduction of net2's last Linear
  PlainSrc src{d_new_obj, dm.D};
  side_before_write(c, p.st[0][3].g);
  SLN_TRY(launch_prep(c.st, src, blk_act(blk[3], p.st[0][3]), p.st[0][3].g, dm.D, blk_fin(c, blk[3], p.st[0][3]), O, dm.D, "out_prep"));
  float* dcat = p.sc.dcat[0];
  SLN_TRY(gconv_bwd(c, p.g, blk, p.st[0], p.pooled[0], p.sc.dpooled, make_view(obj_vecs, dm.D, O, dm.D), make_view(pred_vecs, dm.D, T, dm.D),
                    d_new_pred, dm.D, dcat));
  SLN_TRY(node_gather(c, p.g, dcat, nullptr, nullptr, d_obj));
  if (d_pred && T > 0) SLN_TRY(gather_rows(c, dcat + dm.D, 3 * dm.D, nullptr, T, dm.D, d_pred, dm.D, 0));
  return side_end(c);
}

// ---------------------------------------------------------------- pooling stage alone
size_t sln_gconv_pool_workspace_bytes(int64_t O, int64_t T) {
  Arena ar(nullptr, 0);
  Graph g; int *deg, *err;
  plan_graph(ar, (int)O, (int)T, &g, &deg, &err);
  return ar.off;
}
int sln_csr_build(const int64_t* edges, int64_t edge_stride, int64_t O, int64_t T, void* ws, size_t ws_bytes, void* stream) {
  SLN_TRY(check_dims(O, T));
  SLN_CHECK_ARG(edge_stride == 2, "edges must be a contiguous [T,2] int64 tensor");
  SLN_CHECK_ARG(ws && (uintptr_t)ws % 256 == 0 && ws_bytes >= sln_gconv_pool_workspace_bytes(O, T), "workspace missing, misaligned or too small");
  Arena ar(ws, ws_bytes);
  Graph g; int *deg, *err;
  plan_graph(ar, (int)O, (int)T, &g, &deg, &err);
  return graph_prep_edges((cudaStream_t)stream, g, deg, err, edges);
}
int sln_csr_pointers(void* ws, int64_t O, int64_t T, const int32_t** row_ptr, const int32_t** ent) {
  Arena ar(ws, (size_t)-1);
  Graph g; int *deg, *err;
  plan_graph(ar, (int)O, (int)T, &g, &deg, &err);
  *row_ptr = g.row_ptr; *ent = g.ent;
  return SLN_OK;
}
int sln_gconv_pool_fwd(const float* new_t_vecs, int64_t O, int64_t T, int32_t H, int32_t Dout, float* pooled, const void* ws,
                       size_t ws_bytes, void* stream) {
  SLN_TRY(check_dims(O, T));
  SLN_CHECK_ARG(new_t_vecs && pooled && ws, "null pointer");
  SLN_CHECK_ARG(H > 0 && Dout >= 0, "bad feature sizes");
  Arena ar((void*)ws, ws_bytes);
  Graph g; int *deg, *err;
  plan_graph(ar, (int)O, (int)T, &g, &deg, &err);
  MatView a2 = make_view(new_t_vecs, 2 * H + Dout, (int)T, 2 * H + Dout);
  ProfScope prof((cudaStream_t)stream, PROF_POOL, pool_bytes((int)O, (int)T, H));
  launch_pool_fwd((cudaStream_t)stream, a2, g.row_ptr, g.ent, (int)O, H, Dout, pooled);
  return check_launch("pool_fwd");
}

// ---------------------------------------------------------------- contraction primitive (tests / roofline sweeps)
int sln_set_engine(int engine) {
  int prev = g_engine;
  if (engine == 0 || engine == 1) { g_engine = engine; g_tc_mask = 15; }
  else if (engine >= 16 && engine < 32) { g_engine = 1; g_tc_mask = engine & 15; }
  return prev;
}

int sln_contract(const float* A, int64_t lda, int32_t a_rc, const float* B, int64_t ldb, int32_t b_rc, float* C, int64_t ldc, int64_t M,
                 int64_t N, int64_t K, int32_t accumulate, int32_t engine, void* stream) {
  SLN_CHECK_ARG(A && B && C && M >= 0 && N >= 0 && K >= 0 && M < (1ll << 30) && N < (1ll << 30) && K < (1ll << 30), "bad argument");
  SLN_CHECK_ARG(engine == 0 || engine == 1, "engine must be 0 (fp32 SIMT) or 1 (tcgen05 3xTF32)");
  cudaStream_t st = (cudaStream_t)stream;
  const int m = (int)M, n = (int)N, k = (int)K;
  MatView a = a_rc ? make_view(A, (int)lda, m, k) : make_view(A, (int)lda, k, m);
  MatView b = b_rc ? make_view(B, (int)ldb, n, k) : make_view(B, (int)ldb, k, n);
  const bool tcp = engine == 1;
  if (tcp) SLN_CHECK_ARG(tc::tc_eligible(m, n, k) && a.vec_ok() && b.vec_ok(),
                         "tcgen05 engine needs M >= 64, N >= 32, K >= 32, 16-byte aligned operands and leading dimensions % 4 == 0");
#define SLN_DISPATCH(ARC, BRC)                                                                                                   \
  do {                                                                                                                           \
    if (accumulate) {                                                                                                            \
      if (tcp) { tc::TcEpiAtomic e{C, (int)ldc}; return tc::launch_tc<ARC, BRC>(st, a, b, e, m, n, k, true, "contract_tc", PROF_MISC); } \
      EpiAtomic e{C, (int)ldc}; return launch_gemm<ARC, BRC>(st, a, b, e, m, n, k, true, "contract", PROF_MISC);                 \
    } else {                                                                                                                     \
      EpiStore e; memset(&e, 0, sizeof(e)); e.C = C; e.ldc = (int)ldc;                                                           \
      if (tcp) { tc::TcEpiStore te{C, (int)ldc, nullptr, e.fin}; return tc::launch_tc<ARC, BRC>(st, a, b, te, m, n, k, false, "contract_tc", PROF_MISC); } \
      return launch_gemm<ARC, BRC>(st, a, b, e, m, n, k, false, "contract", PROF_MISC);                                          \
    }                                                                                                                            \
  } while (0)
  if (a_rc && b_rc) SLN_DISPATCH(true, true);
  else if (a_rc && !b_rc) SLN_DISPATCH(true, false);
  else if (!a_rc && b_rc) SLN_DISPATCH(false, true);
  else SLN_DISPATCH(false, false);
#undef SLN_DISPATCH
  return SLN_OK;
}

// tuning aid (not part of the public header): SM-clock trace of the last tcgen05 contraction launched by sln_contract
int sln_debug_tc_trace(long long* out16) {   /* 48 slots */
  SLN_CUDA_TRY(cudaDeviceSynchronize());
  SLN_CUDA_TRY(cudaMemcpyFromSymbol(out16, tc::g_tc_trace, sizeof(long long) * 48));   // all zero unless built with -DSLN_TC_TRACE
  return SLN_OK;
}

// ---------------------------------------------------------------- reparameterisation, losses, Adam
int sln_reparam_fwd(const float* mu, const float* logvar, const float* eps, int64_t n, float* z, void* stream) {
  SLN_CHECK_ARG(mu && logvar && eps && z && n >= 0, "bad argument");
  if (n == 0) return SLN_OK;
  k_reparam_fwd<<<(int)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(mu, logvar, eps, (int)n, z);
  return check_launch("reparam_fwd");
}
int sln_reparam_bwd(const float* d_z, const float* logvar, const float* eps, int64_t n, float* d_mu, float* d_logvar, void* stream) {
  SLN_CHECK_ARG(d_z && logvar && eps && d_mu && d_logvar && n >= 0, "bad argument");
  if (n == 0) return SLN_OK;
  k_reparam_bwd<<<(int)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(d_z, logvar, eps, (int)n, d_mu, d_logvar);
  return check_launch("reparam_bwd");
}

static int vae_loss_impl(const float* boxes_pred, const float* boxes_gt, int32_t box_dim, const float* angles_pred, const int64_t* angles_gt,
                 int32_t n_angle, const float* mu, const float* logvar, int32_t Z, float kl_weight, const float* kl_weight_dev, int64_t O, float* losses, float* d_boxes,
                 float* d_angles, int32_t angles_grad_is_logits, float* d_mu, float* d_logvar, void* scratch, size_t scratch_bytes, void* stream) {
  SLN_CHECK_ARG(boxes_pred && boxes_gt && angles_pred && angles_gt && losses && scratch, "null pointer");
  SLN_CHECK_ARG(O >= 1, "O must be >= 1");
  SLN_CHECK_ARG((mu == nullptr) == (logvar == nullptr), "mu and logvar must both be given or both be null");
  int blocks = (int)ceil_div64(O, 64);
  SLN_CHECK_ARG(scratch_bytes >= 16 + (size_t)12 * blocks, "loss scratch too small");
  LossArgs a;
  a.boxes_pred = boxes_pred; a.boxes_gt = boxes_gt; a.BD = box_dim; a.logp = angles_pred; a.angles_gt = (const long long*)angles_gt;
  a.NA = n_angle; a.mu = mu; a.logvar = logvar; a.Z = Z; a.kl_weight = kl_weight; a.kl_weight_dev = kl_weight_dev; a.O = (int)O;
  a.logits_grad = angles_grad_is_logits;
  a.d_boxes = d_boxes; a.d_logits = d_angles; a.d_mu = d_mu; a.d_logvar = d_logvar;
  a.counter = (unsigned*)scratch; a.partial = (float*)((char*)scratch + 16); a.losses = losses;
  k_vae_loss<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("vae_loss");
}

int sln_vae_loss(const float* boxes_pred, const float* boxes_gt, int32_t box_dim, const float* angles_pred, const int64_t* angles_gt,
                 int32_t n_angle, const float* mu, const float* logvar, int32_t Z, float kl_weight, int64_t O, float* losses, float* d_boxes,
                 float* d_angles, int32_t angles_grad_is_logits, float* d_mu, float* d_logvar, void* scratch, size_t scratch_bytes, void* stream) {
  return vae_loss_impl(boxes_pred, boxes_gt, box_dim, angles_pred, angles_gt, n_angle, mu, logvar, Z, kl_weight, nullptr, O, losses, d_boxes,
                       d_angles, angles_grad_is_logits, d_mu, d_logvar, scratch, scratch_bytes, stream);
}
int sln_vae_loss_dyn(const float* boxes_pred, const float* boxes_gt, int32_t box_dim, const float* angles_pred, const int64_t* angles_gt,
                     int32_t n_angle, const float* mu, const float* logvar, int32_t Z, const float* kl_weight_dev, int64_t O, float* losses,
                     float* d_boxes, float* d_angles, int32_t angles_grad_is_logits, float* d_mu, float* d_logvar, void* scratch,
                     size_t scratch_bytes, void* stream) {
  SLN_CHECK_ARG(kl_weight_dev != nullptr, "null kl_weight_dev");
  return vae_loss_impl(boxes_pred, boxes_gt, box_dim, angles_pred, angles_gt, n_angle, mu, logvar, Z, 0.f, kl_weight_dev, O, losses, d_boxes,
                       d_angles, angles_grad_is_logits, d_mu, d_logvar, scratch, scratch_bytes, stream);
}

static int adam_impl(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, const float* lr_dev, float beta1,
                     float beta2, float eps, float weight_decay, float grad_scale, int64_t* step, int32_t advance_step, const float* guard,
                     void* stream) {
  SLN_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step && n >= 0, "bad argument");
  SLN_CHECK_ARG(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, "arenas must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (advance_step) {
    k_inc_step<<<1, 1, 0, st>>>((long long*)step, guard);
    SLN_TRY(check_launch("adam_inc_step"));
  }
  if (n == 0) return SLN_OK;
  long long threads = ceil_div64(n, 4);
  k_adam<<<(int)ceil_div64(threads, 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, (long long)n, lr, beta1, beta2, eps, weight_decay,
                                                       grad_scale, (const long long*)step, lr_dev, guard);
  return check_launch("adam");
}
int sln_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, float grad_scale, int64_t* step, int32_t advance_step, void* stream) {
  return adam_impl(params, grads, exp_avg, exp_avg_sq, n, lr, nullptr, beta1, beta2, eps, weight_decay, grad_scale, step, advance_step, nullptr, stream);
}
int sln_adam_step_dyn(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, int64_t n, const float* lr_dev, float beta1,
                      float beta2, float eps, float weight_decay, float grad_scale, int64_t* step, int32_t advance_step, const float* guard_loss,
                      void* stream) {
  SLN_CHECK_ARG(lr_dev != nullptr, "null lr_dev");
  return adam_impl(params, grads, exp_avg, exp_avg_sq, n, 0.f, lr_dev, beta1, beta2, eps, weight_decay, grad_scale, step, advance_step, guard_loss, stream);
}

}  // extern "C"
