// Internal definitions shared by the engine translation units (engine.cu: handle, sampling, inference-side C ABI;
// train.cu: the training step).  Not part of the public interface.
#pragma once
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "kernels.cuh"
#include "plan.h"

namespace ds {
int tc_error_flag();
void init_pointwise_attrs();
}
using namespace ds;

int fail(int code, const char* fmt, ...);      // engine.cu: records the message for ds_last_error()
#define CK(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e_ = (call);                                                                             \
    if (e_ != cudaSuccess)                                                                               \
      return fail(DS_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

enum { S_SQRT_AC = 0, S_SQRT_1MAC, S_SQRT_RECIP, S_SQRT_RECIPM1, S_COEF1, S_COEF2, S_SIGMA, S_AC, S_LW, S_NEG_1MAC,
       S_NEG_RECIPM1, S_ZERO, S_ONE, S_COUNT };

struct TrainState;
void train_state_destroy(TrainState* t);      // train.cu

struct ds_handle {
  ds_config cfg;
  Plan plan;
  bool taps = false;
  bool bf16_mode = false, use_tc = false;
  bool gnt = false;      // fused GroupNorm convs run the channels-on-lanes kernel (weights stored row-permuted)
  bool gnt_plain = false;   // ... and so do the plain GEMMs with N % 128 == 0 (fuse_level 3)
  size_t esz = 4;
  cudaStream_t own_stream = nullptr;
  std::map<std::string, std::vector<float>> host_w;
  bool committed = false;
  // packed weights
  char* warena = nullptr;
  std::vector<size_t> w_off;
  float* varena = nullptr;
  std::vector<size_t> v_off;
  // fp32 conditioning path
  float *time_w1 = nullptr, *time_b1 = nullptr, *time_w3 = nullptr, *time_b3 = nullptr, *time_wall = nullptr,
        *time_ball = nullptr, *time_table = nullptr, *sin_freq = nullptr;
  float *ctx_wall = nullptr, *ctx_ball = nullptr, *ctx_table = nullptr;
  int ctx_rows = 0, ctx_batch = 0;
  bool ctx_shared = false, ctx_set = false;
  std::vector<float*> kv_w;
  float* xctx = nullptr;
  int xctx_batch = 0, xctx_cap = 0;
  // activations
  int cap_scenes = 0, rows_cap = 0;
  std::vector<void*> bufs;
  std::vector<TcGemmPlan*> tc;
  std::vector<LnGemmPlan*> lnp;      // fused GEMM + LayerNorm ops
  std::vector<AttnQkvPlan*> atp;     // fused LayerNorm + to_qkv + linear-attention ops
  float* attn_cs = nullptr;          // [n wmats][384] column sums of the gain-folded to_qkv weights (rows of unused matrices are 0)
  int* t_dev = nullptr;
  float* x_state = nullptr;      // [cap, N, d] running sample
  float* x_tmp = nullptr;        // [cap, N, d] scratch (q_sample / host staging)
  int64_t* t64_tmp = nullptr;
  float* loss_parts = nullptr;
  // schedule
  int T = 0, mean_type = DS_MEAN_V;
  std::vector<float> sched_host[S_COUNT];
  float* sched_dev[S_COUNT] = {nullptr};
  StepCoef* coef_dev = nullptr;
  int coef_cap = 0;
  StepState* state_dev = nullptr;
  int64_t launches = 0;
  int t_uniform = 0;          // 1 while every entry of t_dev is the same (inside the sampling loop)
  // instantiated step graph, reused across ds_sample_loop calls whose captured parameters agree (batch, flags,
  // injected-buffer pointers); seed / scene offset / coefficients live in device memory and are not captured
  struct GraphKey {
    int batch = -1, clip = 0, num_partial = 0, plan_gen = 0;
    const void *noise = nullptr, *partial = nullptr, *partial_noise = nullptr;
    cudaStream_t stream = nullptr;
    bool operator==(const GraphKey& o) const {
      return batch == o.batch && clip == o.clip && num_partial == o.num_partial && plan_gen == o.plan_gen &&
             noise == o.noise && partial == o.partial && partial_noise == o.partial_noise && stream == o.stream;
    }
  } gkey;
  cudaGraphExec_t gexec = nullptr;
  int64_t g_launches = 0;     // kernel launches inside one replay
  int plan_gen = 0;           // bumped whenever buffers / tensor maps / tables are rebuilt
  int64_t graph_builds = 0;
  TrainState* train = nullptr;      // lazily created by ds_train_* (train.cu)
};
void drop_graph(ds_handle* h);

