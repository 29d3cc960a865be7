// Host runtime of the JEN-1 denoiser engine (see engine.h).  Reference citations are to
// /root/reference paths as listed in SURVEY.md section 8a.
#include "engine.h"

#include <cuda_bf16.h>

#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace jen1 {

namespace {

__global__ void set_ctl_kernel(CtlBlock* dst, const CtlBlock v) {
  const int* s = reinterpret_cast<const int*>(&v);
  int* d = reinterpret_cast<int*>(dst);
  for (int i = threadIdx.x; i < (int)(sizeof(CtlBlock) / sizeof(int)); i += blockDim.x) d[i] = s[i];
}

inline int cdivi(int a, int b) { return (a + b - 1) / b; }

struct EngineError : std::runtime_error {
  using std::runtime_error::runtime_error;
};

}  // namespace

// ============================================================================================ basics
Engine::Engine(const Jen1ModelDesc& d, int device, int dtype) : d_(d), device_(device), dtype_(dtype) {}

Engine::~Engine() {
  cudaSetDevice(device_);
  if (smp_.exec) cudaGraphExecDestroy(smp_.exec);
  for (void* p : wallocs_) cudaFree(p);
  void* bufs[] = {kv_cond_, ctx_mask_, ctx_rowpart_, tt_t_, tt_tfm_, tt_tft_, tt_m1_, tt_m2_, tt_map_, tt_film_,
                  tt_tok_, tt_tokrp_, tt_kv_, d_ctl_, arena_, smp_.coef, timeline_};
  for (void* p : bufs)
    if (p) cudaFree(p);
}

int Engine::fail(const std::string& m) {
  if (ok_) err_ = m;  // keep the first failure of a call
  ok_ = false;
  return 1;
}

bool Engine::ck(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  if (ok_) fail(std::string(what) + ": " + cudaGetErrorString(e));
  return false;
}

// ============================================================================================ weights
int Engine::load_tensor(const char* name, const float* data, const int64_t* shape, int ndim) {
  ok_ = true;
  if (finalized_) return fail("load_tensor after finalize");
  HostTensor t;
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    t.shape.push_back(shape[i]);
    n *= shape[i];
  }
  t.data.assign(data, data + n);
  host_[name] = std::move(t);
  return 0;
}

const HostTensor& Engine::ht(const std::string& name) {
  auto it = host_.find(name);
  if (it == host_.end()) throw EngineError("state_dict tensor missing: " + name);
  return it->second;
}

void* Engine::wmalloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) throw EngineError("cudaMalloc failed for weights");
  wallocs_.push_back(p);
  weight_total_bytes_ += (int64_t)bytes;
  return p;
}

float* Engine::upload_f32(const std::vector<float>& v) {
  float* p = (float*)wmalloc(v.size() * sizeof(float));
  if (cudaMemcpy(p, v.data(), v.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
    throw EngineError("cudaMemcpy failed (f32 upload)");
  return p;
}

void* Engine::upload_w(const std::vector<float>& v, bool f32) {
  if (f32 || dtype_ == JEN1_DTYPE_F32) return upload_f32(v);
  std::vector<__nv_bfloat16> h(v.size());
  for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16_rn(v[i]);
  void* p = wmalloc(h.size() * 2);
  if (cudaMemcpy(p, h.data(), h.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess)
    throw EngineError("cudaMemcpy failed (bf16 upload)");
  return p;
}

// Conv1d weight [Cout][Cin][k] (or ConvTranspose1d [Cin][Cout][k]) -> [k][Cin][Cout]
DConv Engine::pack_conv(const std::string& prefix, bool transposed, bool count, int cin_pad) {
  const HostTensor& w = ht(prefix + ".weight");
  const HostTensor& b = ht(prefix + ".bias");
  if (w.shape.size() != 3) throw EngineError("conv weight must be 3-D: " + prefix);
  DConv c;
  const int d0 = (int)w.shape[0], d1 = (int)w.shape[1], k = (int)w.shape[2];
  c.Cout = transposed ? d1 : d0;
  const int cin_real = transposed ? d0 : d1;
  c.Cin = cin_pad > cin_real ? cin_pad : cin_real;  // extra input channels (zero weights) match zero-padded inputs
  c.ntaps = k;
  std::vector<float> pk((size_t)k * c.Cin * c.Cout, 0.0f);
  for (int o = 0; o < c.Cout; ++o)
    for (int i = 0; i < cin_real; ++i)
      for (int t = 0; t < k; ++t) {
        const float v = transposed ? w.data[((size_t)i * c.Cout + o) * k + t] : w.data[((size_t)o * cin_real + i) * k + t];
        pk[((size_t)t * c.Cin + i) * c.Cout + o] = v;
      }
  c.w = upload_w(pk, false);
  c.bias = upload_f32(b.data);
  if (count) step_weight_bytes_ += (int64_t)k * cin_real * c.Cout * (int64_t)esz();
  pack_umma(c, pk, transposed);
  return c;
}

// Second copy of a bf16 weight in the tcgen05 blob order (see conv_umma.cu).  A ConvTranspose1d(k=2f, stride f)
// is consumed as f output phases of two taps {z, z+f}; everything else as one phase of `ntaps` taps.
void Engine::pack_umma(DConv& c, const std::vector<float>& pk, bool transposed) {
  if (!use_umma_ || c.w_f32 || c.Cout % 128 != 0) return;
  if (transposed) {
    const int f = c.ntaps / 2;
    c.u_nphase = f;
    c.u_tpp = 2;
    c.u_wtap_phase = 1;
    c.u_wtap_step = f;
  } else {
    c.u_nphase = 1;
    c.u_tpp = c.ntaps;
    c.u_wtap_phase = 0;
    c.u_wtap_step = 1;
  }
  std::vector<uint16_t> blob(conv_umma_packed_elems(c.Cin, c.Cout, c.u_nphase * c.u_tpp));
  conv_umma_pack(pk.data(), c.Cin, c.Cout, c.u_nphase, c.u_tpp, 0, c.u_wtap_phase, c.u_wtap_step, blob.data());
  c.wu = wmalloc(blob.size() * 2);
  if (cudaMemcpy(c.wu, blob.data(), blob.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess)
    throw EngineError("cudaMemcpy failed (tcgen05 weight upload)");
}

// Linear weight [O][I] -> [1][I][O]
DConv Engine::pack_linear_raw(const std::vector<float>& w, const std::vector<float>* bias, int O, int I, bool f32,
                              bool count) {
  DConv c;
  c.Cin = I;
  c.Cout = O;
  c.ntaps = 1;
  c.w_f32 = f32 || dtype_ == JEN1_DTYPE_F32;
  std::vector<float> pk((size_t)I * O);
  for (int o = 0; o < O; ++o)
    for (int i = 0; i < I; ++i) pk[(size_t)i * O + o] = w[(size_t)o * I + i];
  c.w = upload_w(pk, f32);
  if (bias) c.bias = upload_f32(*bias);
  if (count) step_weight_bytes_ += (int64_t)pk.size() * (int64_t)(c.w_f32 ? 4 : 2);
  pack_umma(c, pk, false);
  return c;
}

DNorm Engine::pack_norm(const std::string& prefix, int pad_to) {
  DNorm n;
  std::vector<float> g = ht(prefix + ".weight").data, b = ht(prefix + ".bias").data;
  if (pad_to > (int)g.size()) {  // zero affine on zero-padded channels
    g.resize(pad_to, 0.0f);
    b.resize(pad_to, 0.0f);
  }
  n.gamma = upload_f32(g);
  n.beta = upload_f32(b);
  return n;
}

DRes Engine::pack_res(const std::string& p, int cin, int cout, int cin_pad) {
  DRes r;
  r.cout = cout;
  if (cin_pad > cin) r.gn_real_c = cin;
  r.gn1 = pack_norm(p + ".block1.groupnorm", cin_pad);
  r.c1 = pack_conv(p + ".block1.project.conv", false, true, cin_pad);
  if (cin_pad > cin) cin = cin_pad;
  r.cin = cin;
  r.gn2 = pack_norm(p + ".block2.groupnorm");
  const HostTensor& fw = ht(p + ".to_scale_shift.to_scale_shift.1.weight");
  const HostTensor& fb = ht(p + ".to_scale_shift.to_scale_shift.1.bias");
  r.film_off = (int64_t)film_acc_b_.size();
  film_acc_w_.insert(film_acc_w_.end(), fw.data.begin(), fw.data.end());
  film_acc_b_.insert(film_acc_b_.end(), fb.data.begin(), fb.data.end());
  r.has_out = host_.count(p + ".to_out.conv.weight") > 0;
  if (r.has_out != (cin != cout)) throw EngineError("to_out presence mismatch at " + p);
  if (r.has_out) {
    // fold to_out's bias into block2's bias: both are added to the same accumulator
    HostTensor& b2 = host_[p + ".block2.project.conv.bias"];
    const HostTensor& bo = ht(p + ".to_out.conv.bias");
    for (size_t i = 0; i < b2.data.size(); ++i) b2.data[i] += bo.data[i];
    r.co = pack_conv(p + ".to_out.conv", false, true, cin_pad);
  }
  r.c2 = pack_conv(p + ".block2.project.conv");
  if (r.c1.Cin != cin || r.c1.Cout != cout || r.c2.Cin != cout) throw EngineError("resblock shape mismatch at " + p);
  return r;
}

DAttn Engine::pack_attention(const std::string& p, int C, bool cross) {
  const int ctxC = cross ? E_ : C;
  const HostTensor &gn = ht(p + ".norm.weight"), &bn = ht(p + ".norm.bias");
  const HostTensor &gc = ht(p + ".norm_context.weight"), &bc = ht(p + ".norm_context.bias");
  const HostTensor &wq = ht(p + ".to_q.weight"), &wkv = ht(p + ".to_kv.weight");
  if ((int)wq.shape[0] != C || (int)wq.shape[1] != C || (int)wkv.shape[0] != 2 * C || (int)wkv.shape[1] != ctxC)
    throw EngineError("attention shape mismatch at " + p);
  // LayerNorm affine folded into the projection (exact algebra):  W (n*g + b) = (W diag(g)) n + W b
  std::vector<float> q2((size_t)C * C), qb(C);
  for (int o = 0; o < C; ++o) {
    double acc = 0.0;
    for (int i = 0; i < C; ++i) {
      q2[(size_t)o * C + i] = wq.data[(size_t)o * C + i] * gn.data[i];
      acc += (double)wq.data[(size_t)o * C + i] * (double)bn.data[i];
    }
    qb[o] = (float)acc;
  }
  std::vector<float> kv2((size_t)2 * C * ctxC), kvb(2 * C);
  for (int o = 0; o < 2 * C; ++o) {
    double acc = 0.0;
    for (int i = 0; i < ctxC; ++i) {
      kv2[(size_t)o * ctxC + i] = wkv.data[(size_t)o * ctxC + i] * gc.data[i];
      acc += (double)wkv.data[(size_t)o * ctxC + i] * (double)bc.data[i];
    }
    kvb[o] = (float)acc;
  }
  DAttn a;
  if (!cross) {
    std::vector<float> w(q2);
    w.insert(w.end(), kv2.begin(), kv2.end());
    std::vector<float> b(qb);
    b.insert(b.end(), kvb.begin(), kvb.end());
    a.qkv = pack_linear_raw(w, &b, 3 * C, C, false);
  } else {
    a.qkv = pack_linear_raw(q2, &qb, C, C, false);
    a.kvc_off = (int64_t)kvc_acc_b_.size();
    kvc_acc_w_.insert(kvc_acc_w_.end(), kv2.begin(), kv2.end());
    kvc_acc_b_.insert(kvc_acc_b_.end(), kvb.begin(), kvb.end());
  }
  a.out = pack_linear_raw(ht(p + ".attention.to_out.weight").data, &ht(p + ".attention.to_out.bias").data, C, C, false);
  return a;
}

DTransformer Engine::pack_transformer(const std::string& p, int C, int layers) {
  DTransformer t;
  t.C = C;
  t.gn = pack_norm(p + ".group_norm");
  t.conv = pack_conv(p + ".conv1d.conv");
  const int mid = C * d_.attention_multiplier;
  for (int j = 0; j < layers; ++j) {
    const std::string q = p + ".blocks." + std::to_string(j);
    DTrBlock b;
    b.self = pack_attention(q + ".attention", C, false);
    b.cross = pack_attention(q + ".cross_attention", C, true);
    b.ff1 = pack_linear_raw(ht(q + ".feed_forward.0.weight").data, &ht(q + ".feed_forward.0.bias").data, mid, C, false);
    b.ff2 = pack_linear_raw(ht(q + ".feed_forward.2.weight").data, &ht(q + ".feed_forward.2.bias").data, C, mid, false);
    t.blocks.push_back(b);
  }
  return t;
}

int Engine::finalize() {
  ok_ = true;
  if (finalized_) return fail("finalize called twice");
  if (cudaSetDevice(device_) != cudaSuccess) return fail("cudaSetDevice failed (no CUDA device? there is no CPU fallback)");
  film_acc_w_.clear();
  film_acc_b_.clear();
  kvc_acc_w_.clear();
  kvc_acc_b_.clear();
  {
    const char* impl = getenv("JEN1_CONV_IMPL");  // "generic" forces the fp32-FMA kernel everywhere (A/B testing)
    const char* pdl = getenv("JEN1_PDL");
    use_umma_ = dtype_ == JEN1_DTYPE_BF16 && !(impl && strcmp(impl, "generic") == 0);
    use_pdl_ = !(pdl && strcmp(pdl, "0") == 0);
    const char* at = getenv("JEN1_ATTN_IMPL");  // "fma" forces the fp32-FMA attention core (A/B testing)
    use_umma_attn_ = !(at && strcmp(at, "fma") == 0);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device_) == cudaSuccess) {
      num_sms_ = prop.multiProcessorCount;
      if (prop.major != 10) use_umma_ = false;  // tcgen05 needs sm_100
    }
    if (use_umma_ && conv_umma_init() != cudaSuccess) return fail("tcgen05 path initialisation failed");
    // shared-memory opt-ins of the attention kernels are per device as well (one engine per GPU in one process)
    if (attention_init() != cudaSuccess) return fail("attention kernel initialisation failed");
    if (use_umma_ && (attn_umma_init() != cudaSuccess || attn_flash_init() != cudaSuccess))
      return fail("tcgen05 attention initialisation failed");
    const char* ftr = getenv("JEN1_FUSED_TR");  // "1": run every Transformer1d as ONE fused launch (tr_umma.cu)
    use_fused_tr_ = ftr && strcmp(ftr, "1") == 0;
    if (use_umma_ && use_fused_tr_ && tr_umma_init() != cudaSuccess) return fail("fused transformer initialisation failed");
  }
  try {
    const int nl = d_.num_layers;
    if (nl < 1 || nl > JEN1_MAX_LEVELS) throw EngineError("bad num_layers");
    Fm_ = d_.channels * d_.context_features_multiplier;
    E_ = d_.context_embedding_features;
    tdim_ = d_.channels + 1;
    auto lc = [&](int i) { return d_.channels * d_.multipliers[i]; };
    for (int i = 0; i <= nl; ++i) {
      const int c = lc(i);
      const int gs = c / 32;
      if (c % 32 != 0 || gs < 1 || gs > 64 || 64 % gs != 0)
        throw EngineError("level channel counts must be 32*2^k (<= 2048): got " + std::to_string(c));
    }
    // conditioning networks, fp32 (reference model.py:75-89, 286-291; utils/module.py:58-79)
    to_time_ = pack_linear_raw(ht("to_time.0.1.weight").data, &ht("to_time.0.1.bias").data, Fm_, tdim_, true, false);
    map0_ = pack_linear_raw(ht("to_mapping.0.weight").data, &ht("to_mapping.0.bias").data, Fm_, Fm_, true, false);
    map2_ = pack_linear_raw(ht("to_mapping.2.weight").data, &ht("to_mapping.2.bias").data, Fm_, Fm_, true, false);
    to_tok_ = pack_linear_raw(ht("to_time_embedding.0.1.weight").data, &ht("to_time_embedding.0.1.bias").data, E_,
                              tdim_, true, false);
    tw_map_ = upload_f32(ht("to_time.0.0.weights").data);
    tw_tok_ = upload_f32(ht("to_time_embedding.0.0.weights").data);

    cc_pad_ = use_umma_ ? (d_.context_channels + 7) / 8 * 8 : d_.context_channels;
    if (use_umma_ && d_.in_channels % 8 != 0) cc_pad_ = d_.context_channels;  // rows would not be 16-byte aligned anyway
    to_in_ = pack_res("to_in.block", d_.in_channels + d_.context_channels, lc(0), d_.in_channels + cc_pad_);
    for (int i = 0; i < nl; ++i) {
      DDown D;
      const std::string p = "downsamples." + std::to_string(i);
      D.factor = d_.factors[i];
      D.down = pack_conv(p + ".downsample.conv");
      if (D.down.ntaps != 2 * D.factor + 1) throw EngineError("downsample kernel size mismatch at " + p);
      for (int j = 0; j < d_.num_blocks[i]; ++j)
        D.blocks.push_back(pack_res(p + ".blocks." + std::to_string(j), lc(i + 1), lc(i + 1)));
      D.has_tr = d_.attentions[i] > 0;
      if (D.has_tr) D.tr = pack_transformer(p + ".transformer", lc(i + 1), d_.attentions[i]);
      downs_.push_back(std::move(D));
    }
    mid_pre_ = pack_res("bottleneck.pre_block", lc(nl), lc(nl));
    mid_has_tr_ = d_.attentions[nl] > 0;
    if (mid_has_tr_) mid_tr_ = pack_transformer("bottleneck.transformer", lc(nl), d_.attentions[nl]);
    mid_post_ = pack_res("bottleneck.post_block", lc(nl), lc(nl));
    for (int u = 0; u < nl; ++u) {
      const int i = nl - 1 - u;
      DUp U;
      const std::string p = "upsamples." + std::to_string(u);
      U.factor = d_.factors[i];
      const int nb = d_.num_blocks[i] + (d_.attentions[i] > 0 ? 1 : 0);
      for (int j = 0; j < nb; ++j) U.blocks.push_back(pack_res(p + ".blocks." + std::to_string(j), 2 * lc(i + 1), lc(i + 1)));
      U.has_tr = d_.attentions[i] > 0;
      if (U.has_tr) U.tr = pack_transformer(p + ".transformer", lc(i + 1), d_.attentions[i]);
      U.up = pack_conv(p + ".upsample", /*transposed=*/U.factor != 1);
      if (U.up.ntaps != (U.factor == 1 ? 3 : 2 * U.factor)) throw EngineError("upsample kernel size mismatch at " + p);
      ups_.push_back(std::move(U));
    }
    to_out_ = pack_res("to_out.block", lc(0), d_.out_channels);

    // FiLM: all MappingToScaleShift linears (reference blocks.py:148-165) as one [Fm -> sum 2C] fp32 GEMM
    film_total_ = (int64_t)film_acc_b_.size();
    film_lin_ = pack_linear_raw(film_acc_w_, &film_acc_b_, (int)film_total_, Fm_, true, false);
    // cross-attention K/V projections of the context, all layers as one [E -> sum 2C] GEMM
    kvc_total_ = (int64_t)kvc_acc_b_.size();
    if (kvc_total_ > 0) kvc_lin_ = pack_linear_raw(kvc_acc_w_, &kvc_acc_b_, (int)kvc_total_, E_, false, false);

    // control block + weight-only cache: K/V of the learned null embedding (reference utils/module.py:20-33)
    if (cudaMalloc(&d_ctl_, sizeof(CtlBlock)) != cudaSuccess) throw EngineError("cudaMalloc ctl");
    cudaMemset(d_ctl_, 0, sizeof(CtlBlock));
    const int ctx_len = d_.context_embedding_max_length + 1;
    if (kvc_total_ > 0) {
      const HostTensor& fe = ht("fixed_embedding.embedding.weight");
      if ((int)fe.shape[0] != ctx_len || (int)fe.shape[1] != E_) throw EngineError("fixed_embedding shape mismatch");
      float* fe_d = upload_f32(fe.data);
      float* rp = (float*)wmalloc((size_t)ctx_len * 2 * sizeof(float));
      kv_fixed_ = wmalloc((size_t)ctx_len * kvc_total_ * esz());
      ok_ = true;
      st_ = nullptr;
      dry_ = false;
      ck(launch_rowstats<float>(fe_d, rp, ctx_len, E_, nullptr), "rowstats(fixed)");
      Act a;
      a.ptr = fe_d;
      a.Bt = 1;
      a.L = ctx_len;
      a.C = E_;
      a.rowpart = rp;
      a.rp_nct = 1;
      a.f32 = true;
      ConvOpts o;
      o.mode = PRO_ROWNORM;
      o.Lm = o.Lout = ctx_len;
      conv_into(kvc_lin_, 1, a, o, kv_fixed_);
      if (!ok_) throw EngineError(err_);
      if (cudaDeviceSynchronize() != cudaSuccess) throw EngineError("kernel failure while building the fixed-embedding K/V cache");
    }
  } catch (const std::exception& e) {
    return fail(e.what());
  }
  film_acc_w_ = std::vector<float>();
  film_acc_b_ = std::vector<float>();
  kvc_acc_w_ = std::vector<float>();
  kvc_acc_b_ = std::vector<float>();
  host_.clear();
  finalized_ = true;
  return 0;
}

// ============================================================================================ arena / acts
void* Engine::aalloc(size_t bytes) {
  const size_t a = (arena_off_ + 255) & ~(size_t)255;
  arena_off_ = a + bytes;
  if (dry_) return (void*)(uintptr_t)(a + 256);  // never dereferenced
  if (arena_off_ > arena_cap_) {
    fail("workspace arena overflow (call jen1_engine_reserve with the largest B, T first)");
    return arena_;
  }
  return arena_ + a;
}

Act Engine::new_act(int Bt, int L, int C, bool f32) {
  Act a;
  a.Bt = Bt;
  a.L = L;
  a.C = C;
  a.f32 = f32 || dtype_ == JEN1_DTYPE_F32;
  a.ptr = aalloc((size_t)Bt * L * C * (a.f32 ? 4 : 2));
  return a;
}

// GroupNorm partial statistics are kept per "fine group": 32 per tensor (covers GN(8), GN(32), GN(1) and the
// 4+4 split of GN(8) over a channel concat); narrow tensors (C | 64) keep one whole-tensor group (GN(1) only).
static int fine_groups(int C) { return (C % 32 == 0) ? 32 : ((C > 0 && 64 % C == 0) ? 1 : 0); }

// GroupNorm statistics accumulators live in one contiguous zone per forward so that a single memset (one graph node)
// zeroes them before the launch chain starts; the dry run sizes the zone.
void* Engine::salloc(size_t bytes) {
  const size_t a = (szone_off_ + 15) & ~(size_t)15;
  szone_off_ = a + bytes;
  if (dry_) return (void*)(uintptr_t)(a + 256);
  if (szone_off_ > szone_cap_) {
    fail("internal: statistics zone overflow");
    return szone_;
  }
  return szone_ + a;
}
bool Engine::begin_stats_zone(cudaStream_t st) {
  szone_cap_ = (szone_need_ + 255) & ~(size_t)255;
  szone_ = (char*)aalloc(szone_cap_);
  szone_off_ = 0;
  if (!ok_) return false;
  return ck(cudaMemsetAsync(szone_, 0, szone_cap_, st), "memset(statistics zone)");
}
void Engine::add_stats(Act& a) {
  a.FG = fine_groups(a.C);
  a.stats = (long long*)salloc((size_t)a.Bt * a.FG * 2 * sizeof(long long));
}
void Engine::add_rowpart(Act& a, int nct) {
  a.rp_nct = nct;
  a.rowpart = (float*)aalloc((size_t)a.Bt * a.L * nct * 2 * sizeof(float));
}

bool Engine::ensure_arena(size_t bytes) {
  if (bytes <= arena_cap_) return true;
  cudaDeviceSynchronize();
  if (smp_.exec) {
    cudaGraphExecDestroy(smp_.exec);
    smp_.exec = nullptr;
  }
  smp_.active = false;
  if (arena_) cudaFree(arena_);
  arena_ = nullptr;
  arena_cap_ = 0;
  if (!ck(cudaMalloc((void**)&arena_, bytes), "cudaMalloc(workspace)")) return false;
  arena_cap_ = bytes;
  return true;
}

void Engine::tap(const char* name, const Act& a) {
  if (!dry_ && debug_) taps_[name] = a;
}

bool Engine::upload_ctl(const CtlBlock& c, cudaStream_t st) {
  set_ctl_kernel<<<1, 128, 0, st>>>(d_ctl_, c);
  ++launches_;
  return ck(cudaGetLastError(), "set_ctl");
}

// ============================================================================================ conv launcher
bool Engine::run_conv(const ConvParams& p, bool act_f32, bool w_f32, bool out_f32) {
  if (dry_) return true;
  if (!ok_) return false;
  if ((size_t)2 * p.seg[0].Cin * sizeof(float) > 24 * 1024) {
    fail("conv input channel count too large for the generic kernel");
    return false;
  }
  cudaError_t e;
  if (dtype_ == JEN1_DTYPE_F32 || (act_f32 && w_f32 && out_f32)) {
    if (!(act_f32 && w_f32 && out_f32)) {
      fail("internal: mixed storage types in fp32 mode");
      return false;
    }
    e = launch_conv_generic<float, float, float>(p, st_);
  } else if (!act_f32 && !w_f32 && !out_f32) {
    e = launch_conv_generic<bf16, bf16, bf16>(p, st_);
  } else if (act_f32 && !w_f32 && !out_f32) {
    e = launch_conv_generic<float, bf16, bf16>(p, st_);
  } else if (!act_f32 && !w_f32 && out_f32) {
    e = launch_conv_generic<bf16, bf16, float>(p, st_);
  } else {
    fail("internal: unsupported storage type combination");
    return false;
  }
  ++launches_;
  return ck(e, "conv launch");
}

static ConvSrc make_src(const Act* a, float scale) {
  ConvSrc s;
  memset(&s, 0, sizeof(s));
  s.bmod = 1;
  s.scale = 1.f;
  if (a) {
    s.ptr = a->ptr;
    s.stats = a->stats;
    s.C = a->C;
    s.FG = a->FG;
    s.bmod = a->Bt;
    s.scale = scale;
  }
  return s;
}

// Builds the parameter block; if `into` is non-null the output goes there instead of a fresh arena tensor.
Act Engine::conv_op(const DConv& W, int Bout, const Act& a0, const Act* a1, float scale1, const ConvOpts& o,
                    const DConv* W2, const Act* r0, const Act* r1, float rscale1) {
  return conv_build(W, Bout, a0, a1, scale1, o, W2, r0, r1, rscale1, nullptr);
}
void Engine::conv_into(const DConv& W, int Bout, const Act& a0, const ConvOpts& o, void* into) {
  conv_build(W, Bout, a0, nullptr, 1.f, o, nullptr, nullptr, nullptr, 1.f, into);
}

Act Engine::conv_build(const DConv& W, int Bout, const Act& a0, const Act* a1, float scale1, const ConvOpts& o,
                       const DConv* W2, const Act* r0, const Act* r1, float rscale1, void* into) {
  ConvParams p;
  memset(&p, 0, sizeof(p));
  ConvSeg& s = p.seg[0];
  s.s[0] = make_src(&a0, 1.f);
  s.s[1] = make_src(a1, scale1);
  s.w = W.w;
  s.Cin = a0.C + (a1 ? a1->C : 0);
  s.L = a0.L;
  s.ntaps = o.ntaps;
  s.in_stride = o.in_stride;
  s.shift0 = o.shift0;
  s.shift_step = o.shift_step;
  s.wtap0 = o.wtap0;
  s.wtap_phase = o.wtap_phase;
  s.wtap_step = o.wtap_step;
  p.nseg = 1;
  Act out;
  if (s.Cin != W.Cin || (a1 && a1->L != a0.L) || (a1 && a1->f32 != a0.f32)) {
    fail("internal: conv input shape mismatch (Cin " + std::to_string(s.Cin) + " vs " + std::to_string(W.Cin) + ")");
    return out;
  }
  if (W2) {
    ConvSeg& t = p.seg[1];
    t.s[0] = make_src(r0, 1.f);
    t.s[1] = make_src(r1, rscale1);
    t.w = W2->w;
    t.Cin = r0->C + (r1 ? r1->C : 0);
    t.L = r0->L;
    t.ntaps = 1;
    t.in_stride = 1;
    t.shift_step = 1;
    t.wtap_step = 1;
    p.nseg = 2;
    if (t.Cin != W2->Cin || W2->Cout != W.Cout || r0->L != o.Lout || r0->f32 != a0.f32) {
      fail("internal: residual conv shape mismatch");
      return out;
    }
  }
  p.mode = o.mode;
  p.G = o.G;
  p.gn_real_c = o.gn_real_c;
  p.eps = o.eps;
  if (o.norm) {
    p.gamma = o.norm->gamma;
    p.beta = o.norm->beta;
  }
  p.film = o.film;
  p.film_stride = (int)film_total_;
  p.cond_row = d_ctl_ ? d_ctl_->cond_row : nullptr;
  p.act = o.act;
  p.rowpart = a0.rowpart;
  p.rp_nct = a0.rp_nct;
  p.ln_eps = 1e-5f;
  if (o.mode == PRO_ROWNORM && (!a0.rowpart || a1)) {
    fail("internal: row-norm prologue without row statistics");
    return out;
  }
  if (o.G > 0) {  // GroupNorm group boundaries must coincide with the producers' fine-group boundaries
    const int Ct = s.Cin;
    if (Ct % o.G != 0 || o.G > 32) {
      fail("GroupNorm groups do not divide channels");
      return out;
    }
    const int cpg = Ct / o.G;
    int off = 0;
    for (int k = 0; k < 2; ++k) {
      const ConvSrc& sr = s.s[k];
      if (sr.C > 0) {
        if (!sr.stats || sr.FG <= 0 || sr.C % sr.FG != 0) {
          fail("internal: GroupNorm prologue without producer statistics");
          return out;
        }
        const int gs = sr.C / sr.FG;
        for (int g = 0; g < o.G; ++g) {
          const int lo = std::max(g * cpg, off), hi = std::min((g + 1) * cpg, off + sr.C);
          if (hi > lo && (((lo - off) % gs) || ((hi - off) % gs))) {
            fail("GroupNorm group boundaries are not aligned with the statistics granularity");
            return out;
          }
        }
      }
      off += sr.C;
    }
  }
  p.B = Bout;
  p.Lm = o.Lm;
  p.nphase = o.nphase;
  p.out_stride = o.out_stride;
  p.out_off0 = o.out_off0;
  p.out_off_phase = o.out_off_phase;
  p.Lout = o.Lout;
  p.Cout = W.Cout;
  p.bias = W.bias;
  p.epi_act = o.epi_act;
  if (o.res) {
    p.res = o.res->ptr;
    p.res_bmod = o.res->Bt;
    if (o.res->C != W.Cout || o.res->L != o.Lout || o.res->f32 != a0.f32) {
      fail("internal: residual shape mismatch");
      return out;
    }
  } else {
    p.res_bmod = 1;
  }
  const int TNc = conv_generic_col_tile();
  const bool out_is_f32 = o.out_f32 || dtype_ == JEN1_DTYPE_F32;
  if (o.want_stats && fine_groups(W.Cout) == 0) {
    fail("statistics requested for an unsupported channel count (need C % 32 == 0 or C | 64)");
    return out;
  }
  if (o.want_stats) p.FGo = fine_groups(W.Cout);
  // kernel selection: tcgen05 path for bf16 storage when the shape fits, else the generic fp32-FMA kernel
  UmmaPlan plan;
  memset(&plan, 0, sizeof(plan));
  if (use_umma_ && !a0.f32 && W.wu && (!W2 || W2->wu) && W.u_nphase == o.nphase && W.u_tpp == o.ntaps &&
      (o.nphase == 1 ? (o.wtap0 == 0 && o.wtap_step == 1) : (o.wtap0 == 0 && o.wtap_phase == W.u_wtap_phase && o.wtap_step == W.u_wtap_step))) {
    p.out = (void*)1;  // planning only looks at which of out / out_ncl is set
    plan = conv_umma_plan(p, o.want_stats, num_sms_);
    p.out = nullptr;
  }
  if (into) {
    out.ptr = into;
    out.Bt = Bout;
    out.L = o.Lout;
    out.C = W.Cout;
    out.f32 = out_is_f32;
  } else {
    out = new_act(Bout, o.Lout, W.Cout, o.out_f32);
  }
  if (o.want_stats) {
    add_stats(out);
    p.stats_out = out.stats;
  }
  if (o.want_rowpart) {
    add_rowpart(out, plan.ok ? plan.m_tiles : cdivi(W.Cout, TNc));
    p.rowpart_out = out.rowpart;
  }
  p.out = out.ptr;
  if (debug_ && !dry_) {
    char nm[32];
    snprintf(nm, sizeof nm, "op%03d", op_index_);
    taps_[nm] = out;
    if (trace_)
      fprintf(stderr, "[jen1] op%03d %s B=%d Lm=%d Lout=%d Cin=%d(+%d) Cout=%d taps=%d stride=%d phases=%d G=%d mode=%d | NT=%d tiles=%dx%d splitk=%d stages=%d\n",
              op_index_, plan.ok ? "umma   " : "generic", Bout, o.Lm, o.Lout, s.Cin, W2 ? W2->Cin : 0, W.Cout, o.ntaps,
              o.in_stride, o.nphase, o.G, o.mode, plan.NT, plan.n_tiles, plan.m_tiles, plan.splitk, plan.stages);
    ++op_index_;
  }
  if (plan.ok) {
    if (dry_) return out;
    if (!ok_) return out;
    long long* tl = (timeline_ && tl_ops_ < 1024) ? timeline_ + (size_t)(tl_ops_++) * 32 : nullptr;
    cudaError_t e = launch_conv_umma(p, plan, W.wu, W2 ? W2->wu : nullptr, out.f32, use_pdl_, st_, tl);
    ++launches_;
    ++umma_launches_;
    ck(e, "tcgen05 conv launch");
    return out;
  }
  run_conv(p, a0.f32, W.w_f32 || dtype_ == JEN1_DTYPE_F32, out.f32);
  return out;
}

// ============================================================================================ layers
// ResnetBlock1d (reference blocks.py:219-231): block1 -> FiLM block2 -> + to_out(x); the 1x1 to_out conv rides
// along as a second K segment of block2's launch, an identity residual is added in its epilogue.
Act Engine::resblock(const DRes& R, const Act& x, const Act* skip, float sscale, int groups, bool causal, int Bout,
                     bool out_f32) {
  ConvOpts o1;
  o1.ntaps = 3;
  o1.shift0 = causal ? -2 : -1;
  o1.Lm = o1.Lout = x.L;
  o1.mode = PRO_AFFINE;
  o1.G = groups;
  o1.eps = 1e-5f;
  o1.norm = &R.gn1;
  o1.act = ACT_SILU;
  o1.want_stats = true;
  o1.gn_real_c = R.gn_real_c;
  Act h = conv_op(R.c1, Bout, x, skip, sscale, o1);
  ConvOpts o2 = o1;
  o2.gn_real_c = 0;
  o2.norm = &R.gn2;
  o2.film = tt_film_ + R.film_off;
  o2.want_stats = !out_f32;
  o2.out_f32 = out_f32;
  if (R.has_out) return conv_op(R.c2, Bout, h, nullptr, 1.f, o2, &R.co, &x, skip, sscale);
  if (skip) {
    fail("internal: concat input with identity residual");
    return h;
  }
  o2.res = &x;
  return conv_op(R.c2, Bout, h, nullptr, 1.f, o2);
}

Act Engine::attention_core(const Act& q, const Act* kvself, const DAttn* cross, int C, bool causal) {
  AttnParams p;
  memset(&p, 0, sizeof(p));
  const int H = d_.attention_heads;
  p.q = q.ptr;
  p.q_ld = q.C;
  p.q_off = 0;
  p.B2 = q.Bt;
  p.N = q.L;
  p.H = H;
  p.d = C / H;
  p.C = C;
  p.scale = (float)std::pow((double)p.d, -0.5);
  p.cond_row = d_ctl_->cond_row;
  if (!cross) {
    p.M = q.L;
    p.Bc = q.Bt;
    p.causal = causal ? 1 : 0;
    p.kv = kvself->ptr;
    p.kv_ld = kvself->C;
    p.k_off = C;
    p.v_off = 2 * C;
  } else {
    p.cross = 1;
    p.M = ctx_S_ + 1;
    p.Bc = ctx_B_;
    p.kv_cond = kv_cond_;
    p.kv_fixed = kv_fixed_;
    p.kv_time = tt_kv_;
    p.kvc_ld = (int)kvc_total_;
    p.kvc_off = (int)cross->kvc_off;
    p.drop = d_ctl_->drop;
    p.mask = ctx_has_mask_ ? ctx_mask_ : nullptr;
  }
  Act out = new_act(q.Bt, q.L, C);
  p.out = out.ptr;
  if (debug_ && !dry_) {
    char nm[32];
    snprintf(nm, sizeof nm, "at%03d", at_index_++);
    taps_[nm] = out;
  }
  if (dry_ || !ok_) return out;
  if (C % H != 0 || p.d > 128) {
    fail("attention head dimension must divide channels and be <= 128");
    return out;
  }
  cudaError_t e;
  if (use_umma_ && use_umma_attn_ && attn_umma_supported(p)) {
    e = launch_attention_umma(p, use_pdl_, st_);
    ++umma_attn_launches_;
  } else if (use_umma_ && use_umma_attn_ && attn_flash_supported(p)) {  // more than 256 keys: key-tiled online softmax
    e = launch_attention_flash(p, use_pdl_, st_);
    ++umma_attn_launches_;
  } else {
    e = (dtype_ == JEN1_DTYPE_F32) ? launch_attention<float>(p, st_) : launch_attention<bf16>(p, st_);
  }
  ++launches_;
  ck(e, "attention launch");
  return out;
}

// Fused Transformer1d: builds the op list of tr_umma.cu (same op order, same intermediate tensors and the same debug tap
// names as the unfused chain below, so the A/B harness compares them op by op).
Act Engine::transformer_fused(const DTransformer& Tr, const Act& x, bool causal, int Bout) {
  const int C = Tr.C, N = x.L, H = d_.attention_heads;
  TrParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.B2 = Bout;
  tp.N = N;
  tp.C = C;
  tp.H = H;
  tp.d = C / H;
  tp.Bc = ctx_B_;
  tp.causal = causal ? 1 : 0;
  tp.CS = (C >= 512 && conv_umma_max_cluster() >= 16) ? 16 : 8;
  tp.scale = (float)std::pow((double)tp.d, -0.5);
  tp.gn_eps = 1e-6f;
  tp.gn_stats = x.stats;
  tp.gn_gamma = Tr.gn.gamma;
  tp.gn_beta = Tr.gn.beta;
  tp.kv_cond = (const bf16*)kv_cond_;
  tp.kv_fixed = (const bf16*)kv_fixed_;
  tp.kv_time = (const bf16*)tt_kv_;
  tp.kvc_ld = (int)kvc_total_;
  tp.drop = d_ctl_->drop;
  tp.mask = ctx_has_mask_ ? ctx_mask_ : nullptr;
  tp.cond_row = d_ctl_->cond_row;
  int n = 0;
  auto tap_as = [&](const char* pref, int& counter, const Act& a) {
    if (debug_ && !dry_) {
      char nm[32];
      snprintf(nm, sizeof nm, "%s%03d", pref, counter);
      taps_[nm] = a;
    }
    if (debug_ && !dry_) ++counter;
  };
  auto gemm = [&](const DConv& W, const Act& src, int pro, bool gelu, const Act* res, int Cout, bool stats) -> Act {
    Act dst = new_act(Bout, N, Cout);
    TrOp& o = tp.ops[n++];
    o.type = TR_GEMM;
    o.src = (const bf16*)src.ptr;
    o.src_ld = src.C;
    o.src_bmod = src.Bt;
    o.K = src.C;
    o.pro = pro;
    o.w = (const bf16*)W.wu;
    o.bias = W.bias;
    o.Cout = Cout;
    o.gelu = gelu ? 1 : 0;
    o.res = res ? (const bf16*)res->ptr : nullptr;
    o.res_ld = res ? res->C : 0;
    o.dst = (bf16*)dst.ptr;
    o.dst_ld = Cout;
    o.stats = stats ? 1 : 0;
    tap_as("op", op_index_, dst);
    return dst;
  };
  auto attn = [&](const Act& q, const Act* kvself, const DAttn* cross) -> Act {
    Act ao = new_act(Bout, N, C);
    TrOp& o = tp.ops[n++];
    o.type = TR_ATTN;
    o.q = (const bf16*)q.ptr;
    o.q_ld = q.C;
    o.ao = (bf16*)ao.ptr;
    if (!cross) {
      o.cross = 0;
      o.M = N;
      o.kv = (const bf16*)kvself->ptr;
      o.kv_ld = kvself->C;
      o.k_off = C;
      o.v_off = 2 * C;
    } else {
      o.cross = 1;
      o.M = ctx_S_ + 1;
      o.kvc_off = (int)cross->kvc_off;
    }
    tap_as("at", at_index_, ao);
    return ao;
  };
  Act t = gemm(Tr.conv, x, TRP_GN, false, nullptr, C, false);
  for (const DTrBlock& blk : Tr.blocks) {
    Act qkv = gemm(blk.self.qkv, t, TRP_LN, false, nullptr, 3 * C, false);
    Act ao = attn(qkv, &qkv, nullptr);
    Act t1 = gemm(blk.self.out, ao, TRP_RAW, false, &t, C, false);
    Act q = gemm(blk.cross.qkv, t1, TRP_LN, false, nullptr, C, false);
    Act ao2 = attn(q, nullptr, &blk.cross);
    Act t2 = gemm(blk.cross.out, ao2, TRP_RAW, false, &t1, C, false);
    Act hdn = gemm(blk.ff1, t2, TRP_RAW, true, nullptr, C, false);
    t = gemm(blk.ff2, hdn, TRP_RAW, false, &t2, C, false);
  }
  Act out = gemm(Tr.conv, t, TRP_RAW, false, nullptr, C, true);
  add_stats(out);
  tp.stats_out = out.stats;
  tp.FGo = out.FG;
  tp.n_ops = n;
  if (dry_ || !ok_) return out;
  if (timeline_ && tr_tl_n_ < 16) tp.timeline = timeline_ + 1024 * 32 + (size_t)(tr_tl_n_++) * TR_MAX_OPS * 8;
  cudaError_t e = launch_tr_umma(tp, use_pdl_, st_);
  ++launches_;
  ++fused_tr_launches_;
  ck(e, "fused transformer launch");
  return out;
}

// Transformer1d (reference blocks.py:528-537) with TransformerBlock (:483-489).  Tokens are the channels-last
// rows themselves, so the two rearranges are free and nn.Linear == 1x1 conv.
Act Engine::transformer(const DTransformer& Tr, const Act& x, bool causal, int Bout) {
  const int C = Tr.C, N = x.L;
  // ---- fused path: the whole Transformer1d as ONE launch (tr_umma.cu); the chain below is the fallback (fp32 engine,
  //      shapes outside the fused kernel's envelope, JEN1_FUSED_TR=0)
  if (use_umma_ && use_fused_tr_ && !x.f32 && x.stats && x.FG == 32 && d_.attention_multiplier == 1 && Tr.conv.wu &&
      x.Bt >= 1 && (Bout == x.Bt || Bout == 2 * x.Bt) &&
      tr_umma_supported(N, C, d_.attention_heads, ctx_S_ + 1, (int)Tr.blocks.size()) &&
      tr_umma_smem_bytes(N, C, d_.attention_heads, std::max(N, ctx_S_ + 1)) <= (size_t)227 * 1024) {
    bool packed = true;
    for (const DTrBlock& b : Tr.blocks)
      packed = packed && b.self.qkv.wu && b.self.out.wu && b.cross.qkv.wu && b.cross.out.wu && b.ff1.wu && b.ff2.wu;
    if (packed) return transformer_fused(Tr, x, causal, Bout);
  }
  ConvOpts oin;
  oin.Lm = oin.Lout = N;
  oin.G = 32;
  oin.eps = 1e-6f;
  oin.norm = &Tr.gn;
  oin.want_rowpart = true;
  Act t = conv_op(Tr.conv, Bout, x, nullptr, 1.f, oin);
  ConvOpts oln;
  oln.Lm = oln.Lout = N;
  oln.mode = PRO_ROWNORM;
  for (size_t bi = 0; bi < Tr.blocks.size(); ++bi) {
    const DTrBlock& blk = Tr.blocks[bi];
    Act qkv = conv_op(blk.self.qkv, Bout, t, nullptr, 1.f, oln);
    Act ao = attention_core(qkv, &qkv, nullptr, C, causal);
    ConvOpts oo;
    oo.Lm = oo.Lout = N;
    oo.res = &t;
    oo.want_rowpart = true;
    Act t1 = conv_op(blk.self.out, Bout, ao, nullptr, 1.f, oo);
    Act q = conv_op(blk.cross.qkv, Bout, t1, nullptr, 1.f, oln);
    Act ao2 = attention_core(q, nullptr, &blk.cross, C, false);
    ConvOpts oo2;
    oo2.Lm = oo2.Lout = N;
    oo2.res = &t1;
    Act t2 = conv_op(blk.cross.out, Bout, ao2, nullptr, 1.f, oo2);
    ConvOpts of1;
    of1.Lm = of1.Lout = N;
    of1.epi_act = ACT_GELU;
    Act hdn = conv_op(blk.ff1, Bout, t2, nullptr, 1.f, of1);
    ConvOpts of2;
    of2.Lm = of2.Lout = N;
    of2.res = &t2;
    of2.want_rowpart = bi + 1 < Tr.blocks.size();
    t = conv_op(blk.ff2, Bout, hdn, nullptr, 1.f, of2);
  }
  ConvOpts oz;
  oz.Lm = oz.Lout = N;
  oz.want_stats = true;
  return conv_op(Tr.conv, Bout, t, nullptr, 1.f, oz);  // the same 1x1 conv applied a second time
}

// UNet1d.forward (reference model.py:225-265).  Rows [0,B) and [B,2B) of the CFG batch share x, t and the
// concat conditioning, so everything before the first cross-attention is computed once for B rows and read by
// both halves (`Bt`/`bmod`).
bool Engine::unet(const Act& xpk, const Act& ccpk, int B, int B2, int T, bool causal, Act* y) {
  const int nl = d_.num_layers;
  const int G = d_.resnet_groups;
  const float sscale = d_.use_skip_scale ? (float)std::pow(2.0, -0.5) : 1.0f;
  int Bcur = B;
  Act x = resblock(to_in_, xpk, &ccpk, 1.f, 1, false, Bcur, false);  // Patcher: GN groups=1, never causal
  tap("to_in", x);
  const Act a0 = x;
  std::vector<int> Ls(nl + 1);
  Ls[0] = T;
  std::vector<std::vector<Act>> skips(nl);
  for (int i = 0; i < nl; ++i) {
    const DDown& D = downs_[i];
    const int f = D.factor;
    ConvOpts o;
    o.ntaps = 2 * f + 1;
    o.in_stride = f;
    o.shift0 = causal ? -2 * f : -f;
    o.Lm = o.Lout = cdivi(x.L, f);
    o.want_stats = true;
    x = conv_op(D.down, Bcur, x, nullptr, 1.f, o);
    Ls[i + 1] = x.L;
    for (const DRes& R : D.blocks) {
      x = resblock(R, x, nullptr, 1.f, G, causal, Bcur, false);
      skips[i].push_back(x);
    }
    if (D.has_tr) {
      Bcur = B2;
      x = transformer(D.tr, x, causal, Bcur);
      skips[i].push_back(x);
    }
    char nm[32];
    snprintf(nm, sizeof nm, "down%d", i);
    tap(nm, x);
  }
  x = resblock(mid_pre_, x, nullptr, 1.f, G, causal, Bcur, false);
  if (mid_has_tr_) {
    Bcur = B2;
    x = transformer(mid_tr_, x, causal, Bcur);
  }
  x = resblock(mid_post_, x, nullptr, 1.f, G, causal, Bcur, false);
  tap("mid", x);
  for (int u = 0; u < nl; ++u) {
    const int i = nl - 1 - u;
    const DUp& U = ups_[u];
    for (const DRes& R : U.blocks) {
      if (skips[i].empty()) {
        fail("internal: skip stack underflow");
        return false;
      }
      const Act s = skips[i].back();
      skips[i].pop_back();
      if (s.L != x.L) {
        fail("internal: skip length mismatch");
        return false;
      }
      x = resblock(R, x, &s, sscale, G, causal, Bcur, false);
    }
    if (U.has_tr) x = transformer(U.tr, x, causal, Bcur);
    const int f = U.factor;
    const int target = Ls[i];
    ConvOpts o;
    o.want_stats = true;
    if (i == 0) {
      o.res = &a0;  // `x += skips_list.pop()` (model.py:261) folded into the last up-conv's epilogue
      Bcur = B2;    // to_out always sees the full CFG batch
    }
    if (f == 1) {  // plain nn.Conv1d k3 p1, never causal (blocks.py:72-75)
      o.ntaps = 3;
      o.shift0 = -1;
      o.Lm = o.Lout = x.L;
      if (x.L != target) {
        fail("up-conv output length does not match the skip length");
        return false;
      }
    } else {  // ConvTranspose1d(k=2f, s=f, p=f/2+f%2, op=f%2) as f output phases of two taps each; the centre
              // crop of UpsampleBlock1d.add_skip (blocks.py:732-734, utils/module.py:186-204) is an output offset
      const int pad = f / 2 + f % 2;
      const int full = f * x.L;
      const int dcrop = full - target;
      if (dcrop < 0 || (i == 0 && dcrop != 0)) {
        fail("up-conv output shorter than the skip (or final add length mismatch)");
        return false;
      }
      o.nphase = f;
      o.ntaps = 2;
      o.shift0 = 0;
      o.shift_step = -1;
      o.wtap0 = 0;
      o.wtap_phase = 1;
      o.wtap_step = f;
      o.Lm = x.L + 1;
      o.out_stride = f;
      o.out_off0 = -pad - dcrop / 2;
      o.out_off_phase = 1;
      o.Lout = target;
    }
    x = conv_op(U.up, Bcur, x, nullptr, 1.f, o);
    char nm[32];
    snprintf(nm, sizeof nm, "up%d", u);
    tap(i == 0 ? "pre_out" : nm, x);  // the last up-conv already carries the `x += skip` of model.py:261
  }
  *y = resblock(to_out_, x, nullptr, 1.f, 1, false, B2, true);  // Unpatcher, fp32 channels-last output
  return ok_;
}

bool Engine::pack_inputs(const float* x, int B, int T, Act* xpk, bool with_cc, const float* cc, Act* ccpk) {
  auto one = [&](const float* src, int C, int Cp, Act* a) {
    *a = new_act(B, T, Cp);
    a->FG = 1;
    a->stats = (long long*)salloc((size_t)B * 2 * sizeof(long long));
    if (dry_ || !ok_) return;
    cudaError_t e = (dtype_ == JEN1_DTYPE_F32) ? launch_pack_ncl<float>(src, (float*)a->ptr, a->stats, B, C, Cp, T, st_)
                                               : launch_pack_ncl<bf16>(src, (bf16*)a->ptr, a->stats, B, C, Cp, T, st_);
    ++launches_;
    ck(e, "pack launch");
  };
  if (with_cc) one(cc, d_.context_channels, cc_pad_, ccpk);
  one(x, d_.in_channels, d_.in_channels, xpk);
  return ok_;
}

// ============================================================================================ caches
int Engine::set_timesteps(const int64_t* t_host, int n, cudaStream_t st) {
  ok_ = true;
  szone_ = nullptr;  // no statistics zone outside a forward
  if (!finalized_) return fail("engine not finalized");
  if (n < 1) return fail("set_timesteps: n must be >= 1");
  cudaSetDevice(device_);
  ok_ = true;
  if (n > tt_cap_) {
    cudaDeviceSynchronize();
    void** bufs[] = {(void**)&tt_t_, (void**)&tt_tfm_, (void**)&tt_tft_, (void**)&tt_m1_, (void**)&tt_m2_,
                     (void**)&tt_map_, (void**)&tt_film_, (void**)&tt_tok_, (void**)&tt_tokrp_, &tt_kv_};
    for (void** b : bufs) {
      if (*b) cudaFree(*b);
      *b = nullptr;
    }
    const size_t N = (size_t)n;
    bool g = true;
    g &= ck(cudaMalloc((void**)&tt_t_, N * 8), "malloc");
    g &= ck(cudaMalloc((void**)&tt_tfm_, N * tdim_ * 4), "malloc");
    g &= ck(cudaMalloc((void**)&tt_tft_, N * tdim_ * 4), "malloc");
    g &= ck(cudaMalloc((void**)&tt_m1_, N * Fm_ * 4), "malloc");
    g &= ck(cudaMalloc((void**)&tt_m2_, N * Fm_ * 4), "malloc");
    g &= ck(cudaMalloc((void**)&tt_map_, N * Fm_ * 4), "malloc");
    g &= ck(cudaMalloc((void**)&tt_film_, N * (size_t)film_total_ * 4), "malloc");
    g &= ck(cudaMalloc((void**)&tt_tok_, N * E_ * 4), "malloc");
    g &= ck(cudaMalloc((void**)&tt_tokrp_, N * 2 * 4), "malloc");
    g &= ck(cudaMalloc(&tt_kv_, N * (size_t)std::max<int64_t>(kvc_total_, 1) * esz()), "malloc");
    if (!g) return 1;
    tt_cap_ = n;
    if (smp_.exec) {  // table pointers are baked into a captured graph
      cudaGraphExecDestroy(smp_.exec);
      smp_.exec = nullptr;
    }
  }
  tt_n_ = n;
  st_ = st;
  dry_ = false;
  if (!ck(cudaMemcpyAsync(tt_t_, t_host, (size_t)n * 8, cudaMemcpyHostToDevice, st), "memcpy t")) return 1;
  ck(launch_time_features(tt_t_, tw_map_, tt_tfm_, n, d_.channels / 2, st), "time features");
  ck(launch_time_features(tt_t_, tw_tok_, tt_tft_, n, d_.channels / 2, st), "time features");
  launches_ += 2;
  auto rows = [&](float* p, int C) {
    Act a;
    a.ptr = p;
    a.Bt = 1;
    a.L = n;
    a.C = C;
    a.f32 = true;
    return a;
  };
  ConvOpts og;
  og.Lm = og.Lout = n;
  og.epi_act = ACT_GELU;
  og.out_f32 = true;
  // mapping = to_mapping(to_time(t)) (reference model.py:204-223): Linear,GELU | Linear,GELU,Linear,GELU
  conv_into(to_time_, 1, rows(tt_tfm_, tdim_), og, tt_m1_);
  conv_into(map0_, 1, rows(tt_m1_, Fm_), og, tt_m2_);
  conv_into(map2_, 1, rows(tt_m2_, Fm_), og, tt_map_);
  // all 56 MappingToScaleShift: Linear(SiLU(mapping)) (reference blocks.py:156-165)
  ConvOpts ofl;
  ofl.Lm = ofl.Lout = n;
  ofl.act = ACT_SILU;
  ofl.out_f32 = true;
  conv_into(film_lin_, 1, rows(tt_map_, Fm_), ofl, tt_film_);
  // time token = GELU(Linear(features)) (reference model.py:286-291, 315-316) and its K/V row in every layer
  conv_into(to_tok_, 1, rows(tt_tft_, tdim_), og, tt_tok_);
  if (kvc_total_ > 0) {
    ck(launch_rowstats<float>(tt_tok_, tt_tokrp_, n, E_, st), "rowstats(tok)");
    ++launches_;
    Act tok = rows(tt_tok_, E_);
    tok.rowpart = tt_tokrp_;
    tok.rp_nct = 1;
    ConvOpts ok;
    ok.Lm = ok.Lout = n;
    ok.mode = PRO_ROWNORM;
    conv_into(kvc_lin_, 1, tok, ok, tt_kv_);
  }
  return ok_ ? 0 : 1;
}

int Engine::set_context(const float* emb, const float* mask, int B, int S, cudaStream_t st) {
  ok_ = true;
  szone_ = nullptr;
  if (!finalized_) return fail("engine not finalized");
  if (B < 1 || B > 128) return fail("set_context: batch must be in [1, 128]");
  if (S < 1 || S > d_.context_embedding_max_length) return fail("set_context: context length exceeds context_embedding_max_length");
  cudaSetDevice(device_);
  ok_ = true;
  const size_t rows = (size_t)B * S;
  const size_t need = rows * (size_t)std::max<int64_t>(kvc_total_, 1) * esz();
  if (need > kv_cond_cap_ || rows * 2 * 4 > ctx_rowpart_cap_ || rows * 4 > ctx_mask_cap_) {
    cudaDeviceSynchronize();
    if (kv_cond_) cudaFree(kv_cond_);
    if (ctx_rowpart_) cudaFree(ctx_rowpart_);
    if (ctx_mask_) cudaFree(ctx_mask_);
    kv_cond_ = nullptr;
    ctx_rowpart_ = nullptr;
    ctx_mask_ = nullptr;
    bool g = ck(cudaMalloc(&kv_cond_, need), "malloc kv_cond");
    g &= ck(cudaMalloc((void**)&ctx_rowpart_, rows * 2 * 4), "malloc");
    g &= ck(cudaMalloc((void**)&ctx_mask_, rows * 4), "malloc");
    if (!g) return 1;
    kv_cond_cap_ = need;
    ctx_rowpart_cap_ = rows * 2 * 4;
    ctx_mask_cap_ = rows * 4;
    if (smp_.exec) {
      cudaGraphExecDestroy(smp_.exec);
      smp_.exec = nullptr;
    }
  }
  ctx_B_ = B;
  ctx_S_ = S;
  ctx_has_mask_ = mask != nullptr;
  st_ = st;
  dry_ = false;
  if (mask && !ck(cudaMemcpyAsync(ctx_mask_, mask, rows * 4, cudaMemcpyDeviceToDevice, st), "memcpy mask")) return 1;
  if (kvc_total_ > 0) {
    ck(launch_rowstats<float>(emb, ctx_rowpart_, (int)rows, E_, st), "rowstats(ctx)");
    ++launches_;
    Act a;
    a.ptr = (void*)emb;
    a.Bt = 1;
    a.L = (int)rows;
    a.C = E_;
    a.rowpart = ctx_rowpart_;
    a.rp_nct = 1;
    a.f32 = true;
    ConvOpts o;
    o.Lm = o.Lout = (int)rows;
    o.mode = PRO_ROWNORM;
    conv_into(kvc_lin_, 1, a, o, kv_cond_);
  }
  return ok_ ? 0 : 1;
}

// ============================================================================================ forward
size_t Engine::workspace_bytes(int B, int T) {
  if (!finalized_) return 0;
  const bool ok0 = ok_;
  const std::string e0 = err_;
  const size_t off0 = arena_off_;
  const int cb = ctx_B_, cs = ctx_S_;
  if (ctx_S_ == 0) ctx_S_ = d_.context_embedding_max_length;
  ctx_B_ = B;
  dry_ = true;
  ok_ = true;
  arena_off_ = 0;
  szone_off_ = 0;
  Act xpk, ccpk, y;
  pack_inputs(nullptr, B, T, &xpk, true, nullptr, &ccpk);
  unet(xpk, ccpk, B, 2 * B, T, false, &y);
  szone_need_ = szone_off_ + 256;
  const size_t need = arena_off_ + szone_need_ + 8192;
  dry_ = false;
  arena_off_ = off0;
  ctx_B_ = cb;
  ctx_S_ = cs;
  const bool shape_ok = ok_;
  ok_ = ok0;
  if (!shape_ok) return 0;
  err_ = e0;
  return need;
}

int Engine::reserve(int B, int T) {
  ok_ = true;
  if (!finalized_) return fail("engine not finalized");
  cudaSetDevice(device_);
  ok_ = true;
  const size_t need = workspace_bytes(B, T);
  if (need == 0) return fail("reserve: unsupported shape: " + err_);
  return ensure_arena(need) ? 0 : 1;
}

int Engine::forward(const float* x, const float* cc, const int32_t* cond_rows, const uint8_t* drop, int B, int T,
                    int causal, float emb_scale, int scale_cfg, float phi, float* out, cudaStream_t st) {
  ok_ = true;
  if (!finalized_) return fail("engine not finalized");
  cudaSetDevice(device_);
  ok_ = true;
  if (B < 1 || B > 128 || T < 1) return fail("forward: bad batch / length");
  if (kvc_total_ > 0 && ctx_B_ != B) return fail("forward: jen1_engine_set_context was not called for this batch size");
  if (tt_n_ < 1) return fail("forward: jen1_engine_set_timesteps was not called");
  const bool cfg = emb_scale != 1.0f;
  const int B2 = cfg ? 2 * B : B;
  CtlBlock c;
  memset(&c, 0, sizeof(c));
  for (int r = 0; r < B2; ++r) {
    const int row = cond_rows ? cond_rows[r % B] : 0;
    if (row < 0 || row >= tt_n_) return fail("forward: conditioning row out of range");
    c.cond_row[r] = row;
  }
  const size_t need = workspace_bytes(B, T);
  if (need == 0) return fail("forward: unsupported shape: " + err_);
  if (!ensure_arena(need)) return 1;
  smp_.active = false;
  st_ = st;
  dry_ = false;
  debug_ = true;
  op_index_ = 0;
  at_index_ = 0;
  trace_ = getenv("JEN1_TRACE") != nullptr;
  if (getenv("JEN1_TIMELINE") && !timeline_) {
    cudaMalloc((void**)&timeline_, (1024 * 32 + 16 * TR_MAX_OPS * 8) * sizeof(long long));
  }
  if (timeline_) cudaMemsetAsync(timeline_, 0, (1024 * 32 + 16 * TR_MAX_OPS * 8) * sizeof(long long), st);
  tl_ops_ = 0;
  tr_tl_n_ = 0;
  taps_.clear();
  arena_off_ = 0;
  if (!upload_ctl(c, st)) return 1;
  if (drop && !ck(cudaMemcpyAsync(d_ctl_->drop, drop, (size_t)B, cudaMemcpyDeviceToDevice, st), "memcpy drop")) return 1;
  Act xpk, ccpk, y;
  if (!begin_stats_zone(st)) return 1;
  if (!pack_inputs(x, B, T, &xpk, true, cc, &ccpk)) return 1;
  if (!unet(xpk, ccpk, B, B2, T, causal != 0, &y)) return 1;
  tap("y", y);
  debug_ = false;
  SamplerParams sp;
  memset(&sp, 0, sizeof(sp));
  sp.y = (const float*)y.ptr;
  sp.B = B;
  sp.C = d_.out_channels;
  sp.L = T;
  sp.cfg = cfg ? 1 : 0;
  sp.emb_scale = emb_scale;
  sp.scale_cfg = scale_cfg;
  sp.phi = phi;
  sp.one_minus_phi = (float)(1.0 - (double)phi);
  sp.mode = 0;
  sp.pred_out = out;
  ck(launch_sampler(sp, st), "sampler launch");
  ++launches_;
  if (timeline_ && ok_) dump_timeline(st);
  return ok_ ? 0 : 1;
}

// Debugging aid (JEN1_TIMELINE): per-op phase clocks of CTA (0,0,0) of every tcgen05 conv launch.
void Engine::dump_timeline(cudaStream_t st) {
  std::vector<long long> h(1024 * 32);
  cudaStreamSynchronize(st);
  cudaMemcpy(h.data(), timeline_, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
  long long prev_end = 0, sum_gap = 0, sum_body = 0;
  for (int i = 0; i < tl_ops_ && i < 1024; ++i) {
    const long long* t = &h[(size_t)i * 32];
    if (t[0] == 0) continue;
    // cycles relative to the return of griddepcontrol.wait; "gap" = ns between the previous op's CTA-0 end and
    // this op's wait return; "body" = ns from wait return to CTA-0 end
    const long long gap = prev_end ? t[11] - prev_end : 0;
    sum_gap += gap;
    sum_body += t[12] - t[11];
    fprintf(stderr, "[jen1-tl] u%03d gap_ns %lld body_ns %lld | cyc: early %lld stats %lld coef %lld panels %lld accfull %lld cluster %lld end %lld part %lld eploop %lld epstats %lld | mma first_a %lld issued %lld\n",
            i, gap, t[12] - t[11], t[2] - t[0], t[3] - t[2], t[4] - t[2], t[5] - t[2], t[6] - t[2],
            t[7] ? t[7] - t[2] : 0, t[8] - t[2], t[16] ? t[16] - t[2] : 0, t[13] ? t[13] - t[2] : 0,
            t[14] ? t[14] - t[2] : 0, t[9] - t[2], t[10] - t[2]);
    prev_end = t[12];
  }
  {  // fused transformer launches: per op (us from the kernel's first op): start, barrier passed, panel/tiles staged,
     // first accumulator ready (attention: softmax done), op body done, fence done
    std::vector<long long> ht((size_t)16 * TR_MAX_OPS * 8);
    cudaMemcpy(ht.data(), timeline_ + 1024 * 32, ht.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    for (int k = 0; k < tr_tl_n_ && k < 16; ++k) {
      const long long* t = &ht[(size_t)k * TR_MAX_OPS * 8];
      const long long t0 = t[0];
      for (int oi = 0; oi < TR_MAX_OPS && t[oi * 8] != 0; ++oi) {
        const long long* o = t + oi * 8;
        fprintf(stderr, "[jen1-trtl] tr%02d op%02d start %.2f wait %.2f staged %.2f acc %.2f body %.2f fence %.2f | mma: panel seen %.2f issued %.2f\n", k, oi,
                (o[0] - t0) / 1965.0, (o[1] - t0) / 1965.0, (o[2] - t0) / 1965.0, (o[3] - t0) / 1965.0, (o[4] - t0) / 1965.0,
                (o[5] - t0) / 1965.0, o[6] ? (o[6] - t0) / 1965.0 : 0.0, o[7] ? (o[7] - t0) / 1965.0 : 0.0);
      }
    }
  }
  fprintf(stderr, "[jen1-tl] total: %d tcgen05 ops, sum gap %.1f us, sum body %.1f us\n", tl_ops_, sum_gap / 1e3, sum_body / 1e3);
}

// ============================================================================================ sampler
int Engine::sample_begin(const float* coef_host, int S, const float* cc, int B, int T, int causal, float emb_scale,
                         int scale_cfg, float phi, int objective, int use_graph, cudaStream_t st) {
  ok_ = true;
  if (getenv("JEN1_TIMELINE") && !timeline_) {
    cudaMalloc((void**)&timeline_, (1024 * 32 + 16 * TR_MAX_OPS * 8) * sizeof(long long));
    cudaMemset(timeline_, 0, (1024 * 32 + 16 * TR_MAX_OPS * 8) * sizeof(long long));
  }
  if (!finalized_) return fail("engine not finalized");
  cudaSetDevice(device_);
  ok_ = true;
  if (B < 1 || B > 128 || T < 1 || S < 1) return fail("sample_begin: bad arguments");
  if (d_.in_channels != d_.out_channels) return fail("sampling needs in_channels == out_channels");
  if (kvc_total_ > 0 && ctx_B_ != B) return fail("sample_begin: set_context was not called for this batch size");
  if (tt_n_ < S) return fail("sample_begin: set_timesteps must provide one conditioning row per step");
  const size_t need = workspace_bytes(B, T);
  if (need == 0) return fail("sample_begin: unsupported shape: " + err_);
  if (!ensure_arena(need)) return 1;
  if ((size_t)S * 8 * 4 > smp_.coef_cap) {
    cudaDeviceSynchronize();
    if (smp_.coef) cudaFree(smp_.coef);
    smp_.coef = nullptr;
    if (!ck(cudaMalloc((void**)&smp_.coef, (size_t)S * 8 * 4), "malloc coef")) return 1;
    smp_.coef_cap = (size_t)S * 8 * 4;
  }
  smp_.coef_h.assign(coef_host, coef_host + (size_t)S * 8);
  if (!ck(cudaMemcpyAsync(smp_.coef, smp_.coef_h.data(), (size_t)S * 8 * 4, cudaMemcpyHostToDevice, st), "memcpy coef")) return 1;
  // A captured step graph stays valid across sample() calls as long as everything baked into it is unchanged:
  // shapes and mode flags (they select kernels / plans) and the buffers the nodes point at (arena, tables, caches
  // -- their owners destroy the graph when they reallocate).  Re-capturing 263 nodes costs ~10 ms per call.
  Sampler::Sig sig;
  memset(&sig, 0, sizeof(sig));
  sig.B = B; sig.T = T; sig.causal = causal; sig.scale_cfg = scale_cfg; sig.objective = objective; sig.use_graph = use_graph;
  sig.emb_scale = emb_scale; sig.phi = phi; sig.ctx_B = ctx_B_; sig.ctx_S = ctx_S_; sig.ctx_has_mask = ctx_has_mask_ ? 1 : 0;
  sig.arena = arena_; sig.coef = smp_.coef; sig.tt_film = tt_film_; sig.kv_cond = kv_cond_;
  if (smp_.exec && memcmp(&sig, &smp_.sig, sizeof(sig)) != 0) {
    cudaGraphExecDestroy(smp_.exec);
    smp_.exec = nullptr;
  }
  smp_.sig = sig;
  smp_.S = S;
  smp_.B = B;
  smp_.T = T;
  smp_.causal = causal;
  smp_.emb_scale = emb_scale;
  smp_.scale_cfg = scale_cfg;
  smp_.phi = phi;
  smp_.objective = objective;
  smp_.use_graph = use_graph;
  // the concat conditioning is step-invariant: pack it once at the base of the arena
  st_ = st;
  dry_ = false;
  arena_off_ = 0;
  smp_.ccpk = new_act(B, T, cc_pad_);
  smp_.ccpk.FG = 1;
  smp_.ccpk.stats = (long long*)aalloc((size_t)B * 2 * sizeof(long long));  // persistent: outside the per-step zone
  if (!ck(cudaMemsetAsync(smp_.ccpk.stats, 0, (size_t)B * 2 * sizeof(long long), st), "memset(cc statistics)")) return 1;
  cudaError_t e = (dtype_ == JEN1_DTYPE_F32)
                      ? launch_pack_ncl<float>(cc, (float*)smp_.ccpk.ptr, smp_.ccpk.stats, B, d_.context_channels, cc_pad_, T, st)
                      : launch_pack_ncl<bf16>(cc, (bf16*)smp_.ccpk.ptr, smp_.ccpk.stats, B, d_.context_channels, cc_pad_, T, st);
  ++launches_;
  if (!ck(e, "pack(cc)")) return 1;
  smp_.arena_base = arena_off_;
  smp_.szone_need = szone_need_;
  smp_.active = true;
  if (!smp_.exec) {
    smp_.g_x = nullptr;
    smp_.g_noise = nullptr;
  }
  return 0;
}

int Engine::sample_step(int step, float* x, const float* noise, const uint8_t* drop, cudaStream_t st) {
  ok_ = true;
  if (!smp_.active) return fail("sample_step without sample_begin");
  cudaSetDevice(device_);
  ok_ = true;
  if (step < 0 || step >= smp_.S) return fail("sample_step: step out of range");
  // the update mixes sigma * noise unless this row is flagged as the last step (coef[step][7]): a NULL noise pointer
  // is only legal there (graph and non-graph paths alike)
  if (noise == nullptr && smp_.coef_h[(size_t)step * 8 + 7] == 0.0f)
    return fail("sample_step: noise must not be NULL unless the step's coefficient row is flagged as last");
  const int B = smp_.B, T = smp_.T;
  const bool cfg = smp_.emb_scale != 1.0f;
  const int B2 = cfg ? 2 * B : B;
  CtlBlock c;
  memset(&c, 0, sizeof(c));
  c.step = step;
  for (int r = 0; r < B2; ++r) c.cond_row[r] = step;
  if (!upload_ctl(c, st)) return 1;
  if (drop && !ck(cudaMemcpyAsync(d_ctl_->drop, drop, (size_t)B, cudaMemcpyDeviceToDevice, st), "memcpy drop")) return 1;

  auto body = [&]() -> bool {
    arena_off_ = smp_.arena_base;
    tl_ops_ = 0;
    Act xpk, y, unused;
    szone_need_ = smp_.szone_need;
    if (!begin_stats_zone(st)) return false;
    if (!pack_inputs(x, B, T, &xpk, false, nullptr, &unused)) return false;
    if (!unet(xpk, smp_.ccpk, B, B2, T, smp_.causal != 0, &y)) return false;
    SamplerParams sp;
    memset(&sp, 0, sizeof(sp));
    sp.y = (const float*)y.ptr;
    sp.B = B;
    sp.C = d_.out_channels;
    sp.L = T;
    sp.cfg = cfg ? 1 : 0;
    sp.emb_scale = smp_.emb_scale;
    sp.scale_cfg = smp_.scale_cfg;
    sp.phi = smp_.phi;
    sp.one_minus_phi = (float)(1.0 - (double)smp_.phi);
    sp.mode = 1;
    sp.x = x;
    sp.noise = noise;
    sp.x_out = x;
    sp.coef = smp_.coef;
    sp.step = &d_ctl_->step;
    sp.objective = smp_.objective;
    ck(launch_sampler(sp, st), "sampler launch");
    ++launches_;
    return ok_;
  };

  st_ = st;
  dry_ = false;
  if (!smp_.use_graph) return body() ? 0 : 1;

  const bool noise_ok = (noise == smp_.g_noise) || noise == nullptr;
  if (!smp_.exec || smp_.g_x != x || !noise_ok) {
    if (smp_.exec) {
      cudaGraphExecDestroy(smp_.exec);
      smp_.exec = nullptr;
    }
    if (noise == nullptr) return fail("sample_step: the first graph-captured step needs a noise buffer");
    cudaGraph_t graph = nullptr;
    const int64_t l0 = launches_, u0 = umma_launches_, a0 = umma_attn_launches_, f0 = fused_tr_launches_;
    if (!ck(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal), "begin capture")) return 1;
    const bool good = body();
    cudaError_t e = cudaStreamEndCapture(st, &graph);
    if (!good || e != cudaSuccess) {
      if (graph) cudaGraphDestroy(graph);
      if (good) ck(e, "end capture");
      return 1;
    }
    smp_.launches_per_step = launches_ - l0;
    smp_.umma_per_step = umma_launches_ - u0;
    smp_.umma_attn_per_step = umma_attn_launches_ - a0;
    smp_.fused_tr_per_step = fused_tr_launches_ - f0;
    umma_attn_launches_ = a0;
    fused_tr_launches_ = f0;
    launches_ = l0;
    umma_launches_ = u0;
    e = cudaGraphInstantiate(&smp_.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (!ck(e, "graph instantiate")) return 1;
    smp_.g_x = x;
    smp_.g_noise = noise;
  }
  if (!ck(cudaGraphLaunch(smp_.exec, st), "graph launch")) return 1;
  launches_ += smp_.launches_per_step;
  umma_launches_ += smp_.umma_per_step;
  umma_attn_launches_ += smp_.umma_attn_per_step;
  fused_tr_launches_ += smp_.fused_tr_per_step;
  if (timeline_) {
    const char* e = getenv("JEN1_TIMELINE_STEP");
    if (e && atoi(e) == step) dump_timeline(st);
  }
  return 0;
}

// ============================================================================================ attention operator
// Stand-alone AttentionBase.forward (reference blocks.py:355-380) on a packed bf16 [B][N][3 * H * d] (q | k | v) tensor:
// the kernels' own entry point for parity tests and the large-N tensor-pipe evidence.  impl: 0 = tcgen05 (single-tile
// kernel up to 256 keys, key-tiled online-softmax kernel beyond), 1 = fp32-FMA core, 2 = force the key-tiled kernel.
int Engine::attention(const void* qkv, void* out, int B, int N, int H, int d, int causal, int impl, cudaStream_t st) {
  ok_ = true;
  if (!finalized_) return fail("engine not finalized");
  if (dtype_ != JEN1_DTYPE_BF16) return fail("attention operator: bf16 engines only");
  if (B < 1 || N < 1 || H < 1 || d < 1) return fail("attention operator: bad shape");
  cudaSetDevice(device_);
  AttnParams p;
  memset(&p, 0, sizeof(p));
  const int C = H * d;
  p.q = qkv;
  p.q_ld = 3 * C;
  p.q_off = 0;
  p.B2 = B;
  p.Bc = B;
  p.N = N;
  p.M = N;
  p.H = H;
  p.d = d;
  p.C = C;
  p.scale = (float)std::pow((double)d, -0.5);
  p.causal = causal ? 1 : 0;
  p.kv = qkv;
  p.kv_ld = 3 * C;
  p.k_off = C;
  p.v_off = 2 * C;
  p.cond_row = d_ctl_->cond_row;
  p.out = out;
  cudaError_t e;
  if (impl == 1) {
    if (d > 128 || (d & 7)) return fail("attention operator: head dim must be a multiple of 8, <= 128");
    e = launch_attention<bf16>(p, st);
  } else if (impl == 2 || !attn_umma_supported(p)) {
    if (!use_umma_ || !attn_flash_supported(p)) return fail("attention operator: shape not supported by the tcgen05 kernels");
    e = launch_attention_flash(p, false, st);
    ++umma_attn_launches_;
  } else {
    if (!use_umma_) return fail("attention operator: tcgen05 path unavailable");
    e = launch_attention_umma(p, false, st);
    ++umma_attn_launches_;
  }
  ++launches_;
  ck(e, "attention operator launch");
  return ok_ ? 0 : 1;
}

// ============================================================================================ debug
int Engine::debug_tensor(const char* name, float* host_out, int64_t capacity, int64_t* shape3) {
  ok_ = true;
  auto it = taps_.find(name);
  if (it == taps_.end()) return fail(std::string("no such tap: ") + name);
  cudaSetDevice(device_);
  const Act& a = it->second;
  const int64_t n = (int64_t)a.Bt * a.L * a.C;
  shape3[0] = a.Bt;
  shape3[1] = a.L;
  shape3[2] = a.C;
  if (n > capacity) return fail("debug_tensor: buffer too small");
  if (cudaDeviceSynchronize() != cudaSuccess) return fail("debug_tensor: device error");
  if (a.f32) {
    if (cudaMemcpy(host_out, a.ptr, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess) return fail("memcpy");
  } else {
    std::vector<__nv_bfloat16> h(n);
    if (cudaMemcpy(h.data(), a.ptr, n * 2, cudaMemcpyDeviceToHost) != cudaSuccess) return fail("memcpy");
    for (int64_t i = 0; i < n; ++i) host_out[i] = __bfloat162float(h[i]);
  }
  return 0;
}

}  // namespace jen1
