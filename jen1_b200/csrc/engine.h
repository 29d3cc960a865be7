// Host-side runtime of the JEN-1 denoiser: weight packing, conditioning caches, the UNet graph walk,
// workspace arena and CUDA-graph capture.  Exposed through the C ABI in include/jen1_b200.h (capi.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/jen1_b200.h"
#include "kernels.h"

namespace jen1 {

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
};

struct DConv {          // one packed conv / linear: W [ntaps][Cin][Cout] (+ bias)
  void* w = nullptr;    // engine dtype unless w_f32
  float* bias = nullptr;
  int Cin = 0, Cout = 0, ntaps = 1;
  bool w_f32 = false;
  void* wu = nullptr;   // tcgen05 blob stream (conv_umma.cu), bf16 engines only
  int u_nphase = 1, u_tpp = 1, u_wtap_phase = 0, u_wtap_step = 1;  // tap order the blob stream was packed for
};
struct DNorm {
  float* gamma = nullptr;
  float* beta = nullptr;
};
struct DRes {  // ResnetBlock1d (reference blocks.py:168-231)
  DNorm gn1, gn2;
  DConv c1, c2, co;  // co: to_out 1x1 (fused into c2's launch as a second K segment; its bias is folded into c2.bias)
  bool has_out = false;
  int cin = 0, cout = 0;
  int gn_real_c = 0;  // block1's input carries zero-padded channels: real channel count for the GroupNorm size
  int64_t film_off = 0;
};
struct DAttn {     // Attention (reference blocks.py:383-437) with the LayerNorm affines folded into the projections
  DConv qkv;       // self: [C -> 3C] (q | k | v);  cross: [C -> C] (q only)
  DConv out;
  int64_t kvc_off = 0;  // cross: column offset of this layer in the K/V caches
};
struct DTrBlock {
  DAttn self, cross;
  DConv ff1, ff2;
};
struct DTransformer {  // Transformer1d (reference blocks.py:497-537)
  DNorm gn;
  DConv conv;
  std::vector<DTrBlock> blocks;
  int C = 0;
};
struct DDown {
  DConv down;
  int factor = 1;
  std::vector<DRes> blocks;
  bool has_tr = false;
  DTransformer tr;
};
struct DUp {
  std::vector<DRes> blocks;
  bool has_tr = false;
  DTransformer tr;
  DConv up;
  int factor = 1;
};

struct Act {  // channels-last activation [Bt][L][C] in the arena
  void* ptr = nullptr;
  int Bt = 0, L = 0, C = 0;
  long long* stats = nullptr;  // GroupNorm statistics [Bt][FG][2], fixed-point accumulators (common.cuh)
  int FG = 0;
  float* rowpart = nullptr;  // LayerNorm partials [Bt][L][rp_nct][2]
  int rp_nct = 0;
  bool f32 = false;
};

struct CtlBlock {  // per-call control data, device resident (written by a tiny kernel from launch parameters)
  int step;
  int cond_row[256];
  unsigned char drop[128];
};

class Engine {
 public:
  Engine(const Jen1ModelDesc& d, int device, int dtype);
  ~Engine();

  int load_tensor(const char* name, const float* data, const int64_t* shape, int ndim);
  int finalize();
  size_t workspace_bytes(int B, int T);
  int reserve(int B, int T);
  int set_context(const float* emb, const float* mask, int B, int S, cudaStream_t st);
  int set_timesteps(const int64_t* t_host, int n, cudaStream_t st);
  int forward(const float* x, const float* cc, const int32_t* cond_rows, const uint8_t* drop, int B, int T,
              int causal, float emb_scale, int scale_cfg, float phi, float* out, cudaStream_t st);
  int sample_begin(const float* coef_host, int S, const float* cc, int B, int T, int causal, float emb_scale,
                   int scale_cfg, float phi, int objective, int use_graph, cudaStream_t st);
  int sample_step(int step, float* x, const float* noise, const uint8_t* drop, cudaStream_t st);
  int debug_tensor(const char* name, float* host_out, int64_t capacity, int64_t* shape3);
  int attention(const void* qkv, void* out, int B, int N, int H, int d, int causal, int impl, cudaStream_t st);

  const char* last_error() const { return err_.c_str(); }
  int64_t launch_count() const { return launches_; }
  int64_t weight_bytes() const { return step_weight_bytes_; }
  int64_t umma_launch_count() const { return umma_launches_; }
  int64_t umma_attn_launch_count() const { return umma_attn_launches_; }
  int64_t fused_tr_launch_count() const { return fused_tr_launches_; }

 private:
  // ---- errors
  int fail(const std::string& m);
  bool ck(cudaError_t e, const char* what);

  // ---- weights
  const HostTensor& ht(const std::string& name);
  void* wmalloc(size_t bytes);
  float* upload_f32(const std::vector<float>& v);
  void* upload_w(const std::vector<float>& v, bool f32);
  DConv pack_conv(const std::string& prefix, bool transposed = false, bool count = true, int cin_pad = 0);
  DConv pack_linear_raw(const std::vector<float>& w, const std::vector<float>* bias, int O, int I, bool f32,
                        bool count = true);
  DNorm pack_norm(const std::string& prefix, int pad_to = 0);
  DRes pack_res(const std::string& prefix, int cin, int cout, int cin_pad = 0);
  DTransformer pack_transformer(const std::string& prefix, int C, int layers);
  DAttn pack_attention(const std::string& prefix, int C, bool cross);

  // ---- arena
  void* aalloc(size_t bytes);
  Act new_act(int Bt, int L, int C, bool f32 = false);
  void add_stats(Act& a);
  void* salloc(size_t bytes);  // from the per-forward statistics zone (zeroed by one memset before the launch chain)
  bool begin_stats_zone(cudaStream_t st);
  void add_rowpart(Act& a, int nct);
  size_t esz() const { return dtype_ == JEN1_DTYPE_F32 ? 4 : 2; }

  // ---- op launchers
  struct ConvOpts {
    int ntaps = 1, in_stride = 1, shift0 = 0, shift_step = 1, wtap0 = 0, wtap_phase = 0, wtap_step = 1;
    int nphase = 1, Lm = 0, out_stride = 1, out_off0 = 0, out_off_phase = 0, Lout = 0;
    int mode = PRO_AFFINE, G = 0;
    float eps = 1e-5f;
    const DNorm* norm = nullptr;
    const float* film = nullptr;
    int act = ACT_NONE, epi_act = ACT_NONE;
    int gn_real_c = 0;
    const Act* res = nullptr;
    bool want_stats = false, want_rowpart = false, out_f32 = false;
  };
  Act conv_op(const DConv& W, int Bout, const Act& a0, const Act* a1, float scale1, const ConvOpts& o,
              const DConv* W2 = nullptr, const Act* r0 = nullptr, const Act* r1 = nullptr, float rscale1 = 1.f);
  void conv_into(const DConv& W, int Bout, const Act& a0, const ConvOpts& o, void* into);
  Act conv_build(const DConv& W, int Bout, const Act& a0, const Act* a1, float scale1, const ConvOpts& o,
                 const DConv* W2, const Act* r0, const Act* r1, float rscale1, void* into);
  bool run_conv(const ConvParams& p, bool act_f32, bool w_f32, bool out_f32);
  void pack_umma(DConv& c, const std::vector<float>& tap_cin_cout, bool transposed);
  Act resblock(const DRes& R, const Act& x, const Act* skip, float sscale, int groups, bool causal, int Bout,
               bool out_f32);
  Act transformer(const DTransformer& Tr, const Act& x, bool causal, int Bout);
  Act transformer_fused(const DTransformer& Tr, const Act& x, bool causal, int Bout);
  Act attention_core(const Act& q, const Act* kvself, const DAttn* cross, int C, bool causal);
  bool unet(const Act& xpk, const Act& ccpk, int B, int B2, int T, bool causal, Act* y);
  void tap(const char* name, const Act& a);
  bool upload_ctl(const CtlBlock& c, cudaStream_t st);
  bool ensure_arena(size_t bytes);
  bool pack_inputs(const float* x, int B, int T, Act* xpk, bool with_cc, const float* cc, Act* ccpk);

  Jen1ModelDesc d_;
  int device_, dtype_;
  std::string err_;
  bool finalized_ = false;
  std::map<std::string, HostTensor> host_;
  // finalize() scratch: all MappingToScaleShift linears / all cross-attention to_kv, concatenated
  std::vector<float> film_acc_w_, film_acc_b_, kvc_acc_w_, kvc_acc_b_;
  std::vector<void*> wallocs_;
  int64_t weight_total_bytes_ = 0, step_weight_bytes_ = 0;
  int64_t launches_ = 0;

  // model
  DRes to_in_, to_out_, mid_pre_, mid_post_;
  DTransformer mid_tr_;
  bool mid_has_tr_ = false;
  std::vector<DDown> downs_;
  std::vector<DUp> ups_;
  int64_t film_total_ = 0, kvc_total_ = 0;
  int Fm_ = 0, E_ = 0, tdim_ = 0;
  int cc_pad_ = 0;  // stored channel count of the packed input-concat conditioning (multiple of 8 on the tcgen05 path)
  // conditioning networks (fp32)
  DConv to_time_, map0_, map2_, film_lin_, to_tok_;
  float *tw_map_ = nullptr, *tw_tok_ = nullptr;
  DConv kvc_lin_;  // [E -> kvc_total] all cross-attention to_kv with norm_context folded
  // caches
  void* kv_fixed_ = nullptr;  // [ctx_len][kvc_total]
  void* kv_cond_ = nullptr;   // [Bc*S][kvc_total]
  size_t kv_cond_cap_ = 0;
  float* ctx_mask_ = nullptr;
  size_t ctx_mask_cap_ = 0;
  bool ctx_has_mask_ = false;
  int ctx_B_ = 0, ctx_S_ = 0;
  float* ctx_rowpart_ = nullptr;
  size_t ctx_rowpart_cap_ = 0;
  // time tables
  int tt_n_ = 0, tt_cap_ = 0;
  int64_t* tt_t_ = nullptr;
  float *tt_tfm_ = nullptr, *tt_tft_ = nullptr, *tt_m1_ = nullptr, *tt_m2_ = nullptr, *tt_map_ = nullptr,
        *tt_film_ = nullptr, *tt_tok_ = nullptr, *tt_tokrp_ = nullptr;
  void* tt_kv_ = nullptr;
  // control
  CtlBlock* d_ctl_ = nullptr;
  // tcgen05 path
  bool use_umma_ = false, use_pdl_ = true, use_umma_attn_ = true, use_fused_tr_ = true;
  int64_t fused_tr_launches_ = 0;
  int64_t umma_attn_launches_ = 0;
  int num_sms_ = 148;
  int64_t umma_launches_ = 0;
  long long* timeline_ = nullptr;
  int tl_ops_ = 0, tr_tl_n_ = 0;
  void dump_timeline(cudaStream_t st);
  // arena
  char* arena_ = nullptr;
  size_t arena_cap_ = 0, arena_off_ = 0;
  char* szone_ = nullptr;          // statistics zone of the forward in progress (inside the arena)
  size_t szone_off_ = 0, szone_cap_ = 0, szone_need_ = 0;  // szone_need_: bytes the last dry run asked for
  bool dry_ = false;
  cudaStream_t st_ = nullptr;
  bool ok_ = true;
  // debug taps
  bool debug_ = false, trace_ = false;
  int op_index_ = 0, at_index_ = 0;
  std::map<std::string, Act> taps_;
  // sampler state
  struct Sampler {
    bool active = false;
    int S = 0, B = 0, T = 0, causal = 0, scale_cfg = 0, objective = 0, use_graph = 0;
    float emb_scale = 1.f, phi = 0.7f;
    float* coef = nullptr;
    size_t coef_cap = 0;
    std::vector<float> coef_h;  // host copy of the per-step scalars (is-last flag validates a NULL noise pointer)
    float* cc_copy = nullptr;  // persistent packed concat-cond lives at the arena base
    Act ccpk;
    size_t arena_base = 0;
    size_t szone_need = 0;  // statistics-zone bytes of one step (from the dry run of sample_begin)
    cudaGraphExec_t exec = nullptr;
    float* g_x = nullptr;
    const float* g_noise = nullptr;
    int64_t launches_per_step = 0, umma_per_step = 0, umma_attn_per_step = 0, fused_tr_per_step = 0;
    struct Sig {  // everything a captured step graph depends on
      int B, T, causal, scale_cfg, objective, use_graph, ctx_B, ctx_S, ctx_has_mask;
      float emb_scale, phi;
      const void *arena, *coef, *tt_film, *kv_cond;
    } sig = {};
  } smp_;
};

}  // namespace jen1
