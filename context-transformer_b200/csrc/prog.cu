// prog.cu — op program: the detector's forward pass recorded once as a flat list of kernel
// launches (conv / pool / layout / attention / softmax) and replayed with one C call per step.
//
// The reference walks nn.ModuleLists in Python every forward (models/RFB_Net_vgg.py:190-286,
// ~300 framework calls per image); here the host builds the list once per (batch, size) and a step
// is `ctx_prog_run`.  `ctx_prog_instantiate_graph` additionally captures the list into a CUDA graph
// so that a step is a single cudaGraphLaunch (the late-pyramid layers are launch-latency bound).
#include "common.cuh"

#include <vector>

namespace ctx {
int conv_simt_launch(const CtxConvParams* p, cudaStream_t st);
int maxpool_launch(const CtxPoolParams* p, cudaStream_t st);
int nchw_to_nhwc_launch(const float* in, void* out, int N, int C, int H, int W, int dtype, cudaStream_t st);
int softmax_launch(const float* in, float* out, long long rows, int cols, cudaStream_t st);
int patch27_launch(const float* in, void* out, int N, int H, int W, int dtype, cudaStream_t st);
int attention_simt_launch(const CtxAttnParams* p, cudaStream_t st);

enum OpKind { OP_CONV_SIMT, OP_CONV_TC, OP_POOL, OP_NCHW2NHWC, OP_PATCH27, OP_ATTN, OP_SOFTMAX };

struct Op {
  OpKind kind;
  CtxConvParams conv;
  CtxPoolParams pool;
  CtxAttnParams attn;
  void* tc_plan = nullptr;
  struct { const float* in; void* out; int N, C, H, W, dtype; } cvt;
  struct { const float* in; float* out; long long rows; int cols; } sm;
};

struct Prog {
  std::vector<Op> ops;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int launches_per_run = 0;      // kernels inside the captured graph (for ctx_launch_count)
};

static int run_op(Op& op, cudaStream_t st) {
  switch (op.kind) {
    case OP_CONV_SIMT: return conv_simt_launch(&op.conv, st);
    case OP_CONV_TC: return ctx_conv2d_tc_plan_run(op.tc_plan, st);
    case OP_POOL: return maxpool_launch(&op.pool, st);
    case OP_NCHW2NHWC: return nchw_to_nhwc_launch(op.cvt.in, op.cvt.out, op.cvt.N, op.cvt.C, op.cvt.H, op.cvt.W, op.cvt.dtype, st);
    case OP_PATCH27: return patch27_launch(op.cvt.in, op.cvt.out, op.cvt.N, op.cvt.H, op.cvt.W, op.cvt.dtype, st);
    case OP_ATTN: return attention_simt_launch(&op.attn, st);
    case OP_SOFTMAX: return softmax_launch(op.sm.in, op.sm.out, op.sm.rows, op.sm.cols, st);
  }
  set_error("prog: corrupt op kind");
  return CTX_ERR_INVALID;
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
  pr->ops.push_back(op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_conv_tc(void* prog, const CtxConvParams* p) {
  PROG_OR_FAIL(prog);
  CTX_REQUIRE(p, "ctx_prog_add_conv_tc: null params");
  Op op{}; op.kind = OP_CONV_TC; op.conv = *p;
  int rc = ctx_conv2d_tc_plan_create(p, &op.tc_plan);
  if (rc) return rc;
  pr->ops.push_back(op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_pool(void* prog, const CtxPoolParams* p) {
  PROG_OR_FAIL(prog);
  CTX_REQUIRE(p, "ctx_prog_add_pool: null params");
  Op op{}; op.kind = OP_POOL; op.pool = *p;
  pr->ops.push_back(op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_nchw_to_nhwc(void* prog, const float* in, void* out, int N, int C, int H, int W, int out_dtype) {
  PROG_OR_FAIL(prog);
  Op op{}; op.kind = OP_NCHW2NHWC;
  op.cvt.in = in; op.cvt.out = out; op.cvt.N = N; op.cvt.C = C; op.cvt.H = H; op.cvt.W = W; op.cvt.dtype = out_dtype;
  pr->ops.push_back(op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_nchw_to_patch27(void* prog, const float* in, void* out, int N, int H, int W, int out_dtype) {
  PROG_OR_FAIL(prog);
  Op op{}; op.kind = OP_PATCH27;
  op.cvt.in = in; op.cvt.out = out; op.cvt.N = N; op.cvt.C = 3; op.cvt.H = H; op.cvt.W = W; op.cvt.dtype = out_dtype;
  pr->ops.push_back(op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_attention(void* prog, const CtxAttnParams* p) {
  PROG_OR_FAIL(prog);
  CTX_REQUIRE(p, "ctx_prog_add_attention: null params");
  Op op{}; op.kind = OP_ATTN; op.attn = *p;
  pr->ops.push_back(op);
  return CTX_OK;
}

extern "C" int ctx_prog_add_softmax(void* prog, const float* in, float* out, long long rows, int cols) {
  PROG_OR_FAIL(prog);
  Op op{}; op.kind = OP_SOFTMAX; op.sm.in = in; op.sm.out = out; op.sm.rows = rows; op.sm.cols = cols;
  pr->ops.push_back(op);
  return CTX_OK;
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

extern "C" int ctx_prog_instantiate_graph(void* prog, void* stream) {
  CTX_REQUIRE(prog, "ctx_prog_instantiate_graph: null handle");
  Prog* pr = (Prog*)prog;
  drop_graph(pr);
  cudaStream_t st = (cudaStream_t)stream;
  CTX_REQUIRE(st != nullptr, "ctx_prog_instantiate_graph: needs a non-default stream to capture on");
  CTX_CUDA_TRY(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const unsigned long long before = ctx_launch_count();
  int rc = CTX_OK;
  for (Op& op : pr->ops) { rc = run_op(op, st); if (rc) break; }
  cudaGraph_t g = nullptr;
  cudaError_t e = cudaStreamEndCapture(st, &g);
  if (rc) { if (g) cudaGraphDestroy(g); return rc; }
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
