// prog.cu — op program: the detector's forward pass recorded once as a flat list of kernel
// launches (conv / pool / layout / attention / softmax) and replayed with one C call per step.
//
// The reference walks nn.ModuleLists in Python every forward (models/RFB_Net_vgg.py:190-286,
// ~300 framework calls per image); here the host builds the list once per (batch, size) and a step
// is `ctx_prog_run`.  `ctx_prog_instantiate_graph` additionally captures the list into a CUDA graph
// so that a step is a single cudaGraphLaunch (the late-pyramid layers are launch-latency bound).
#include "common.cuh"

#include <stdlib.h>
#include <array>
#include <map>
#include <mutex>
#include <vector>

namespace ctx {
int conv_simt_launch(const CtxConvParams* p, cudaStream_t st);
int maxpool_launch(const CtxPoolParams* p, cudaStream_t st);
int nchw_to_nhwc_launch(const float* in, void* out, int N, int C, int H, int W, int dtype, cudaStream_t st);
int softmax_launch(const float* in, float* out, long long rows, int cols, cudaStream_t st);
int patch27_launch(const float* in, void* out, int N, int H, int W, int dtype, cudaStream_t st);
int attention_simt_launch(const CtxAttnParams* p, cudaStream_t st);

enum OpKind { OP_CONV_SIMT, OP_CONV_TC, OP_CONV_X3, OP_POOL, OP_NCHW2NHWC, OP_PATCH27, OP_ATTN, OP_SOFTMAX };

constexpr int MAX_LANES = 8;

struct Op {
  OpKind kind;
  int lane = 0;                  // stream the op is captured on (graph mode); serial replay ignores lanes
  unsigned wait_mask = 0;        // lanes whose work issued so far must finish before this op starts
  CtxConvParams conv;
  CtxPoolParams pool;
  CtxAttnParams attn;
  void* tc_plan = nullptr;
  void* x3_plan = nullptr;
  bool fixed_plan = false;       // the tensor-core plan has no alternative tilings (fused conv1_1 + conv1_2)
  struct { const float* in; void* out; int N, C, H, W, dtype; } cvt;
  struct { const float* in; float* out; long long rows; int cols; } sm;
};

struct Prog {
  std::vector<Op> ops;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int launches_per_run = 0;      // kernels inside the captured graph (for ctx_launch_count)
  int cur_lane = 0;
  unsigned pending_wait = 0;
  cudaStream_t lane_stream[MAX_LANES] = {};
  std::vector<cudaEvent_t> events;
};

static void push_op(Prog* pr, Op& op) {
  op.lane = pr->cur_lane;
  op.wait_mask = pr->pending_wait & ~(1u << pr->cur_lane);
  pr->pending_wait = 0;
  pr->ops.push_back(op);
}

static int run_op(Op& op, cudaStream_t st) {
  switch (op.kind) {
    case OP_CONV_SIMT: return conv_simt_launch(&op.conv, st);
    case OP_CONV_TC: return ctx_conv2d_tc_plan_run(op.tc_plan, st);
    case OP_CONV_X3: return ctx_conv2d_x3_plan_run(op.x3_plan, st);
    case OP_POOL: return maxpool_launch(&op.pool, st);
    case OP_NCHW2NHWC: return nchw_to_nhwc_launch(op.cvt.in, op.cvt.out, op.cvt.N, op.cvt.C, op.cvt.H, op.cvt.W, op.cvt.dtype, st);
    case OP_PATCH27: return patch27_launch(op.cvt.in, op.cvt.out, op.cvt.N, op.cvt.H, op.cvt.W, op.cvt.dtype, st);
    case OP_ATTN: return attention_simt_launch(&op.attn, st);
    case OP_SOFTMAX: return softmax_launch(op.sm.in, op.sm.out, op.sm.rows, op.sm.cols, st);
  }
  set_error("prog: corrupt op kind");
  return CTX_ERR_INVALID;
}

// Tilings found by ctx_prog_autotune, per conv geometry, for the life of the process: recompiling a program after a
// parameter change (load_state_dict, normalize(), another batch of the same size) re-uses them instead of timing ~30
// candidates per layer again.  Key: everything the candidate timings depend on (shapes, strides, dtypes, epilogue kind).
using TuneKey = std::array<int, 22>;
struct TuneVal { int n, cl, amode, cg; };
static std::map<TuneKey, TuneVal>& tune_cache() { static std::map<TuneKey, TuneVal> m; return m; }
static std::mutex& tune_mutex() { static std::mutex m; return m; }
static TuneKey tune_key(const CtxConvParams& c) {
  return TuneKey{c.N, c.H, c.W, c.Cin, c.in_cstride, c.in_coffset % 64, c.Cout, c.KH, c.KW, c.stride, c.pad_h, c.pad_w, c.dil, c.relu,
                 c.pool2, c.relu_channels, c.in_dtype, c.nseg, c.seg[0].dtype, c.residual ? 1 : 0, c.seg[0].pix_stride, c.res_cstride};
}

static void drop_graph(Prog* pr) {
  if (pr->exec) { cudaGraphExecDestroy(pr->exec); pr->exec = nullptr; }
  if (pr->graph) { cudaGraphDestroy(pr->graph); pr->graph = nullptr; }
}
}  // namespace ctx

using namespace ctx;

extern "C" int ctx_prog_create(void** prog_out) {
  CTX_REQUIRE(prog_out, "ctx_prog_create: null output");
  *prog_out = new Prog();
  return CTX_OK;
}

extern "C" void ctx_prog_destroy(void* prog) {
  if (!prog) return;
  Prog* pr = (Prog*)prog;
  drop_graph(pr);
  for (Op& op : pr->ops)
    if (op.tc_plan) ctx_conv2d_tc_plan_destroy(op.tc_plan);
  for (Op& op : pr->ops)
    if (op.x3_plan) ctx_conv2d_x3_plan_destroy(op.x3_plan);
  for (cudaEvent_t e : pr->events) cudaEventDestroy(e);
  for (cudaStream_t s : pr->lane_stream) if (s) cudaStreamDestroy(s);
  delete pr;
}

#define PROG_OR_FAIL(prog) \
  CTX_REQUIRE(prog, "prog: null handle"); \
  Prog* pr = (Prog*)prog; \
  drop_graph(pr)

extern "C" int ctx_prog_add_conv_simt(void* prog, const CtxConvParams* p) {
  PROG_OR_FAIL(prog);
  CTX_REQUIRE(p, "ctx_prog_add_conv_simt: null params");
  Op op{}; op.kind = OP_CONV_SIMT; op.conv = *p;
  push_op(pr, op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_conv_tc(void* prog, const CtxConvParams* p) {
  PROG_OR_FAIL(prog);
  CTX_REQUIRE(p, "ctx_prog_add_conv_tc: null params");
  Op op{}; op.kind = OP_CONV_TC; op.conv = *p;
  int rc = ctx_conv2d_tc_plan_create(p, &op.tc_plan);
  if (rc) return rc;
  push_op(pr, op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_conv_stem2(void* prog, const CtxConvParams* p, const float* stem_in, const void* stem_weight, const float* stem_bias) {
  PROG_OR_FAIL(prog);
  CTX_REQUIRE(p, "ctx_prog_add_conv_stem2: null params");
  Op op{}; op.kind = OP_CONV_TC; op.conv = *p; op.fixed_plan = true;
  int rc = ctx_conv2d_stem2_plan_create(p, stem_in, stem_weight, stem_bias, &op.tc_plan);
  if (rc) return rc;
  push_op(pr, op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_conv_x3(void* prog, const CtxConvParams* p) {
  PROG_OR_FAIL(prog);
  CTX_REQUIRE(p, "ctx_prog_add_conv_x3: null params");
  Op op{}; op.kind = OP_CONV_X3; op.conv = *p;
  int rc = ctx_conv2d_x3_plan_create(p, 0, &op.x3_plan);
  if (rc) return rc;
  push_op(pr, op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_pool(void* prog, const CtxPoolParams* p) {
  PROG_OR_FAIL(prog);
  CTX_REQUIRE(p, "ctx_prog_add_pool: null params");
  Op op{}; op.kind = OP_POOL; op.pool = *p;
  push_op(pr, op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_nchw_to_nhwc(void* prog, const float* in, void* out, int N, int C, int H, int W, int out_dtype) {
  PROG_OR_FAIL(prog);
  Op op{}; op.kind = OP_NCHW2NHWC;
  op.cvt.in = in; op.cvt.out = out; op.cvt.N = N; op.cvt.C = C; op.cvt.H = H; op.cvt.W = W; op.cvt.dtype = out_dtype;
  push_op(pr, op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_nchw_to_patch27(void* prog, const float* in, void* out, int N, int H, int W, int out_dtype) {
  PROG_OR_FAIL(prog);
  Op op{}; op.kind = OP_PATCH27;
  op.cvt.in = in; op.cvt.out = out; op.cvt.N = N; op.cvt.C = 3; op.cvt.H = H; op.cvt.W = W; op.cvt.dtype = out_dtype;
  push_op(pr, op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_attention(void* prog, const CtxAttnParams* p) {
  PROG_OR_FAIL(prog);
  CTX_REQUIRE(p, "ctx_prog_add_attention: null params");
  Op op{}; op.kind = OP_ATTN; op.attn = *p;
  push_op(pr, op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_softmax(void* prog, const float* in, float* out, long long rows, int cols) {
  PROG_OR_FAIL(prog);
  Op op{}; op.kind = OP_SOFTMAX; op.sm.in = in; op.sm.out = out; op.sm.rows = rows; op.sm.cols = cols;
  push_op(pr, op);
  return CTX_OK;
}

extern "C" int ctx_prog_set_lane(void* prog, int lane, unsigned wait_mask) {
  CTX_REQUIRE(prog, "ctx_prog_set_lane: null handle");
  CTX_REQUIRE(lane >= 0 && lane < MAX_LANES && (wait_mask >> MAX_LANES) == 0, "ctx_prog_set_lane: lane %d / mask %u out of range", lane, wait_mask);
  Prog* pr = (Prog*)prog;
  pr->cur_lane = lane;
  pr->pending_wait |= wait_mask;
  return CTX_OK;
}

extern "C" int ctx_prog_conv_config(void* prog, int op_index, int* info6) {          // info6: 8 ints (ctx_conv2d_tc_plan_info)
  CTX_REQUIRE(prog && info6, "ctx_prog_conv_config: null argument");
  Prog* pr = (Prog*)prog;
  CTX_REQUIRE(op_index >= 0 && op_index < (int)pr->ops.size(), "ctx_prog_conv_config: bad op index %d", op_index);
  for (int i = 0; i < 8; ++i) info6[i] = 0;
  if (pr->ops[op_index].kind == OP_CONV_X3) return ctx_conv2d_x3_plan_info(pr->ops[op_index].x3_plan, info6);
  if (pr->ops[op_index].kind != OP_CONV_TC) return CTX_OK;
  return ctx_conv2d_tc_plan_info(pr->ops[op_index].tc_plan, info6);
}

// Per-layer tile selection by measurement: every tensor-core conv is timed in place (its real buffers; an op only ever
// reads its inputs and rewrites its own outputs, so running it out of order is harmless) under each candidate tiling —
// N-tile count (wave quantisation on 148 SMs vs. operand reuse), CTA pairs, TMA patches vs. im2col gather — and the
// fastest plan is kept.  All candidates produce bit-identical outputs.
extern "C" int ctx_prog_autotune(void* prog, void* stream, int reps) {
  PROG_OR_FAIL(prog);
  cudaStream_t st = (cudaStream_t)stream;
  if (reps < 1) reps = 3;
  cudaEvent_t e0, e1;
  CTX_CUDA_TRY(cudaEventCreate(&e0));
  CTX_CUDA_TRY(cudaEventCreate(&e1));
  int rc = CTX_OK;
  // optional trace of every measurement (development aid): CTX_AUTOTUNE_LOG=<path>
  const char* log_path = getenv("CTX_AUTOTUNE_LOG");
  FILE* log = log_path && log_path[0] ? fopen(log_path, "a") : nullptr;
  int op_index = -1;
  auto trace = [&](const Op& op, const int* info, float ms, const char* tag) {
    if (!log) return;
    fprintf(log, "op %d  %dx%dx%d->%d k%dx%d s%d d%d  bn %d ntn %d cl %d amode %d stages %d grid %d commit %d patch %dx%d  %.2f us %s\n", op_index,
            op.conv.H, op.conv.W, op.conv.Cin, op.conv.Cout, op.conv.KH, op.conv.KW, op.conv.stride, op.conv.dil, info[0], info[1], info[2],
            info[3], info[4], info[5], info[6], info[7] / 1000, info[7] % 1000, 1000.f * ms / reps, tag);
  };
  auto time_plan = [&](void* plan, float* ms) -> int {
    int r = ctx_conv2d_tc_plan_run(plan, st);                  // warm-up (descriptor fetch, L2)
    if (r) return r;
    if (cudaEventRecord(e0, st) != cudaSuccess) return CTX_ERR_CUDA;
    for (int i = 0; i < reps && !r; ++i) r = ctx_conv2d_tc_plan_run(plan, st);
    if (r) return r;
    if (cudaEventRecord(e1, st) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess) { set_error("autotune: timing failed: %s", cudaGetErrorString(cudaGetLastError())); return CTX_ERR_CUDA; }
    cudaEventElapsedTime(ms, e0, e1);
    return CTX_OK;
  };
  for (Op& op : pr->ops) {
    ++op_index;
    if (op.kind != OP_CONV_TC || op.conv.in_nchw || op.fixed_plan) continue;
    const TuneKey tkey = tune_key(op.conv);
    {
      std::lock_guard<std::mutex> lock(tune_mutex());
      auto it = tune_cache().find(tkey);
      if (it != tune_cache().end()) {
        void* cand = nullptr;
        const TuneVal& v = it->second;
        if (ctx_conv2d_tc_plan_create_tuned(&op.conv, v.n, v.cl, v.amode, v.cg, &cand) == CTX_OK) {
          ctx_conv2d_tc_plan_destroy(op.tc_plan);
          op.tc_plan = cand;
          continue;
        }
      }
    }
    float best_ms = 0.f;
    if ((rc = time_plan(op.tc_plan, &best_ms))) break;
    int base[8];
    ctx_conv2d_tc_plan_info(op.tc_plan, base);
    trace(op, base, best_ms, "default");
    const int n0 = base[1];
    int best_n = 0, best_cl = 0, best_amode = -1;             // arguments that produced the current best plan
    std::vector<long> seen;                                   // (tile width, cluster, A mode, commit group) already measured
    auto key_of = [](const int* info) { return ((long)info[0] * 100 + info[2] * 10 + info[3]) * 10 + info[6]; };
    seen.push_back(key_of(base));
    auto consider = [&](int n, int cl, int amode, int cg) -> bool {          // true if the candidate became the best plan
      void* cand = nullptr;
      if (ctx_conv2d_tc_plan_create_tuned(&op.conv, n, cl, amode, cg, &cand) != CTX_OK) return false;
      int info[8];
      ctx_conv2d_tc_plan_info(cand, info);
      const long key = key_of(info);
      bool dup = info[0] < 32;
      for (long k : seen) dup = dup || k == key;
      if (dup) { ctx_conv2d_tc_plan_destroy(cand); return false; }
      seen.push_back(key);
      float ms = 0.f;
      rc = time_plan(cand, &ms);
      const bool better = !rc && ms < best_ms * 0.97f;
      if (!rc) trace(op, info, ms, better ? "better" : "");
      if (better) { ctx_conv2d_tc_plan_destroy(op.tc_plan); op.tc_plan = cand; best_ms = ms; }
      else ctx_conv2d_tc_plan_destroy(cand);
      return better;
    };
    // pass 1: tile width x CTA pairs x A-operand mode, commit group by rule; pass 2: commit group on the winner
    for (int dn = 0; dn < 6 && !rc; ++dn) {
      const int n = dn < 4 ? n0 + dn : n0 * (dn == 4 ? 2 : 3);
      for (int amode = -1; amode <= 5 && !rc; ++amode)
        for (int cl = 1; cl <= 2 && !rc; ++cl)
          if (consider(n, cl, amode, 0)) { best_n = n; best_cl = cl; best_amode = amode; }
    }
    int best_cg = 0;
    for (int cg = 1; cg <= 4 && !rc; cg *= 2)
      if (consider(best_n, best_cl, best_amode, cg)) best_cg = cg;
    if (rc) break;
    {
      std::lock_guard<std::mutex> lock(tune_mutex());
      tune_cache()[tkey] = TuneVal{best_n, best_cl, best_amode, best_cg};
    }
  }
  if (log) fclose(log);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return rc;
}

extern "C" int ctx_prog_num_ops(void* prog) { return prog ? (int)((Prog*)prog)->ops.size() : 0; }

extern "C" int ctx_prog_run_range(void* prog, int first, int last, void* stream) {
  CTX_REQUIRE(prog, "ctx_prog_run_range: null handle");
  Prog* pr = (Prog*)prog;
  CTX_REQUIRE(first >= 0 && last <= (int)pr->ops.size() && first <= last, "ctx_prog_run_range: bad range [%d,%d)", first, last);
  for (int i = first; i < last; ++i) {
    int rc = run_op(pr->ops[i], (cudaStream_t)stream);
    if (rc) return rc;
  }
  return CTX_OK;
}

// the same range `reps` times back to back from native code: per-kernel timing of small ops without the host language's
// per-call overhead (a Python -> ctypes call costs ~8 us, more than most of the late-pyramid kernels run)
extern "C" int ctx_prog_run_range_repeat(void* prog, int first, int last, int reps, void* stream) {
  CTX_REQUIRE(prog, "ctx_prog_run_range_repeat: null handle");
  Prog* pr = (Prog*)prog;
  CTX_REQUIRE(first >= 0 && last <= (int)pr->ops.size() && first <= last && reps >= 0, "ctx_prog_run_range_repeat: bad range [%d,%d) x %d", first, last, reps);
  for (int r = 0; r < reps; ++r)
    for (int i = first; i < last; ++i) {
      int rc = run_op(pr->ops[i], (cudaStream_t)stream);
      if (rc) return rc;
    }
  return CTX_OK;
}

extern "C" int ctx_prog_instantiate_graph(void* prog, void* stream) {
  CTX_REQUIRE(prog, "ctx_prog_instantiate_graph: null handle");
  Prog* pr = (Prog*)prog;
  drop_graph(pr);
  cudaStream_t st = (cudaStream_t)stream;
  CTX_REQUIRE(st != nullptr, "ctx_prog_instantiate_graph: needs a non-default stream to capture on");
  // Lanes: ops the host put on lane L > 0 (independent RFB branches, the per-level heads) are captured on their own
  // stream, forked from / joined to lane 0 with events, so the graph is a DAG and small kernels run side by side.
  for (Op& op : pr->ops)
    if (op.lane > 0 && !pr->lane_stream[op.lane]) CTX_CUDA_TRY(cudaStreamCreateWithFlags(&pr->lane_stream[op.lane], cudaStreamNonBlocking));
  size_t ev_used = 0;
  auto next_event = [&](cudaEvent_t* ev) -> cudaError_t {
    if (ev_used == pr->events.size()) {
      cudaEvent_t e;
      cudaError_t rc = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
      if (rc != cudaSuccess) return rc;
      pr->events.push_back(e);
    }
    *ev = pr->events[ev_used++];
    return cudaSuccess;
  };
  CTX_CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const unsigned long long before = ctx_launch_count();
  int rc = CTX_OK;
  cudaError_t ce = cudaSuccess;
  unsigned active = 1u;                              // lanes that are part of the capture and not yet joined back
  auto lane_st = [&](int lane) { return lane == 0 ? st : pr->lane_stream[lane]; };
  auto make_wait = [&](int waiter, int signal) {
    cudaEvent_t ev;
    if (ce == cudaSuccess) ce = next_event(&ev);
    if (ce == cudaSuccess) ce = cudaEventRecord(ev, lane_st(signal));
    if (ce == cudaSuccess) ce = cudaStreamWaitEvent(lane_st(waiter), ev, 0);
  };
  for (Op& op : pr->ops) {
    unsigned mask = op.wait_mask & active;
    if (!(active & (1u << op.lane))) { mask |= 1u; active |= 1u << op.lane; }     // a lane enters the capture by waiting on lane 0
    for (int l = 0; l < MAX_LANES; ++l)
      if ((mask >> l) & 1u) make_wait(op.lane, l);
    if (ce != cudaSuccess) break;
    rc = run_op(op, lane_st(op.lane));
    if (rc) break;
  }
  for (int l = 1; l < MAX_LANES && ce == cudaSuccess && !rc; ++l)
    if ((active >> l) & 1u) make_wait(0, l);         // every forked lane joins lane 0 before the capture ends
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(st, &g);
  if (rc) { if (g) cudaGraphDestroy(g); return rc; }
  if (ce != cudaSuccess) { if (g) cudaGraphDestroy(g); set_error("ctx_prog_instantiate_graph: lane fork/join failed: %s", cudaGetErrorString(ce)); return CTX_ERR_CUDA; }
  if (e != cudaSuccess) { set_error("ctx_prog_instantiate_graph: capture failed: %s", cudaGetErrorString(e)); return CTX_ERR_CUDA; }
  pr->graph = g;
  pr->launches_per_run = (int)(ctx_launch_count() - before);
  e = cudaGraphInstantiate(&pr->exec, g, 0);
  if (e != cudaSuccess) { drop_graph(pr); set_error("ctx_prog_instantiate_graph: instantiate failed: %s", cudaGetErrorString(e)); return CTX_ERR_CUDA; }
  return CTX_OK;
}

extern "C" int ctx_prog_run(void* prog, void* stream) {
  CTX_REQUIRE(prog, "ctx_prog_run: null handle");
  Prog* pr = (Prog*)prog;
  if (pr->exec) {
    CTX_CUDA_TRY(cudaGraphLaunch(pr->exec, (cudaStream_t)stream));
    count_launch(pr->launches_per_run);
    return CTX_OK;
  }
  return ctx_prog_run_range(prog, 0, (int)pr->ops.size(), stream);
}
