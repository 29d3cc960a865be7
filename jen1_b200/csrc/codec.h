// Encodec (SEANet) decoder engine -- see codec.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <initializer_list>
#include <map>
#include <string>
#include <vector>

#include "../../include/jen1_b200.h"
#include "conv_params.h"

namespace jen1 {

class CodecDecoder {
 public:
  // encoder = false: SEANetDecoder (latent -> audio); true: SEANetEncoder + residual vector quantizer (audio -> latent)
  CodecDecoder(const Jen1CodecDesc& d, int device, int strict, bool encoder = false, int n_q = 0, int codebook_size = 0);
  ~CodecDecoder();

  int load_tensor(const char* name, const float* data, const int64_t* shape, int ndim);
  int finalize();
  size_t workspace_bytes(int B, int T);
  int reserve(int B, int T);
  int decode(const float* latent, float* audio, int B, int T, cudaStream_t st);
  // audio [N][channels][L] -> latent [N][dimension][T] (continuous), codes [n_q][N][T] (or null), quantized [N][dimension][T]
  // (or null); T = ceil(L / hop)
  int encode(const float* audio, float* latent, int32_t* codes, float* quantized, int N, int L, cudaStream_t st);
  size_t encode_workspace_bytes(int N, int L);
  int quantize(const float* latent, int32_t* codes, float* quantized, int N, int T, cudaStream_t st);
  bool is_encoder() const { return encoder_; }

  const char* last_error() const { return err_.c_str(); }
  int64_t launch_count() const { return launches_; }
  int64_t weight_bytes() const { return weight_bytes_; }
  int hop() const { return hop_; }
  int lstm_cluster() const { return CS_; }
  int64_t tf32_launch_count() const { return tf32_launches_; }
  int64_t lstm_tc_launch_count() const { return lstm_tc_launches_; }

 private:
  struct HostTensor {
    std::vector<int64_t> shape;
    std::vector<float> data;
  };
  struct ConvW {  // tap-major fp32 weights [k][Cin][Cout] + bias + the GroupNorm(1, Cout) affine applied by consumers
    const float* w = nullptr;
    const float* wT = nullptr;  // [k][Cout][Cin] copy for the TF32 kernel
    const float* bias = nullptr;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    int cin = 0, cout = 0, k = 0;
  };
  struct LstmW {
    const float* wih = nullptr;   // [H][4H] (transposed for the k=1 tap-GEMM)
    const float* wih_T = nullptr; // [4H][H] as loaded (TF32 kernel layout)
    const float* bias = nullptr;  // b_ih + b_hh
    const uint4* whh = nullptr;   // fp16, packed per cluster rank (codec.cu LstmParams)
    const float* whh_f32 = nullptr;  // [4H][H] as loaded (the tensor-core kernel builds its register fragments from it)
  };
  struct Stage {
    int ratio = 1;
    ConvW up, res1, res2, shortcut;  // encoder: `up` is the strided down conv that follows the resblock
  };
  // an activation: raw values + (optionally) the statistics / affine of the GroupNorm its consumers must apply
  struct Act {
    float* ptr = nullptr;
    long long* stats = nullptr;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    int C = 0, L = 0, Lstore = 0, row0 = 0, FG = 1;
  };

  int fail(const std::string& m);
  bool ck(cudaError_t e, const char* what);
  const HostTensor* get(const std::string& name, std::initializer_list<int64_t> shape);
  float* upload(const std::vector<float>& v);
  bool make_conv(const std::string& prefix, const char* conv, const char* norm, int cin, int cout, int k, bool transposed,
                 ConvW* out);
  size_t lstm_smem() const;
  float* falloc(size_t n);
  long long* salloc(int B, int FG);
  void fill_src(ConvParams& p, const Act& in, const Act* in2, int act);
  cudaError_t launch_conv(const ConvParams& p, cudaStream_t st);
  Act conv(const Act& in, const Act* in2, int act, const ConvW& W, int pad_left, bool reflect, bool want_stats, cudaStream_t st,
           int stride = 1);
  void walk_encoder(const float* audio, float* latent, int32_t* codes, float* quantized, int N, int L, cudaStream_t st);
  int finalize_encoder();
  int load_lstm();
  int finish_finalize();
  Act convtr(const Act& in, const Act* in2, int act, const ConvW& W, int r, cudaStream_t st);
  Act last_conv(const Act& in, const Act* in2, const ConvW& W, int pad_left, cudaStream_t st);
  void walk(const float* latent, float* audio, int B, int T, cudaStream_t st);
  Act lstm_stack(const Act& y0, int B, int T, cudaStream_t st);
  bool launch_lstm_tc(int l, const float* gx, int gxT, int gx_t0, float* hout, float* cstate, int t0, int t1, int B, int T,
                      cudaStream_t st);
  static constexpr int kLstmChunks = 8;

  Jen1CodecDesc d_;
  int device_;
  bool encoder_ = false;
  int n_q_ = 0, K_ = 0, cin0_ = 0;      // quantizer stages, codebook size, padded channel count of the audio input
  const float* codebooks_ = nullptr;   // [n_q][K][dimension]
  const float* enorm_ = nullptr;       // [n_q][K] squared norms
  bool finalized_ = false, dry_ = false, ok_ = true, strict_ = false;
  std::string err_;
  std::string lstm_prefix_ = "model.1.lstm.";
  std::map<std::string, HostTensor> host_;
  std::vector<void*> owned_;
  ConvW first_, last_;
  std::vector<LstmW> lstm_;
  std::vector<Stage> stages_;
  int H_ = 0, hop_ = 1, CS_ = 0, U_ = 0, B_ = 0;
  uint8_t* arena_ = nullptr;
  size_t arena_bytes_ = 0, off_ = 0, soff_ = 0, stats_bytes_need_ = 0;
  int64_t launches_ = 0, weight_bytes_ = 0, tf32_launches_ = 0, lstm_tc_launches_ = 0;
  bool lstm_smem_kernel_ = false, lstm_no_overlap_ = false;
  cudaStream_t side_ = nullptr;
  std::vector<cudaEvent_t> ev_;
};

}  // namespace jen1
