// Internal model representation: raw state-dict tensors, host-packed tables, device pointers.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "gemm.cuh"

namespace swc {

struct RawTensor {
  std::vector<int64_t> shape;
  int dtype;                   // 0 f32, 1 i32
  std::vector<float> f;
  std::vector<int32_t> i;
  int64_t numel() const { int64_t n = 1; for (auto d : shape) n *= d; return n; }
};

// one packed table: fp32 on the host; uploaded either as fp32 or (GEMM weights in bf16 mode) as bf16
struct Packed {
  std::vector<float> host;
  bool as_act_type = false;    // true: stored in the activation dtype of the model (GEMM operand)
  void* dev = nullptr;
};

struct LinearW {               // W [w_rows, K] row-major + optional bias[N]
  const void* w = nullptr;
  const void* w3 = nullptr;    // bf16x3 mode: [w_rows, 3K] bf16 = (hi | lo | hi) planes of the fp32 weight, else null
  const float* bias = nullptr;
  int N = 0, w_rows = 0, K = 0;
};

struct LayerW {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  LinearW qkv, out, fc1, fc2;
};

struct ResUnitW {
  const float *a0, *b0, *fu0, *fd0, *a2, *b2, *fu2, *fd2;
  LinearW conv7, conv1;
  int dilation;
};

struct VocosBlockW {
  const float *dw_w, *dw_b, *ln_g, *ln_b, *gamma;
  LinearW pw1, pw2;
};

struct Model {
  int precision = 0;           // SWC_PRECISION_*
  int device = -1;
  int num_sms = 148;
  bool packed = false, uploaded = false;
  std::map<std::string, RawTensor> raw;
  std::map<std::string, Packed> tab;

  // architecture (config/SimWhisperCodec.yaml)
  int d_model = 768, heads = 12, ffn = 3072, n_enc = 0, n_dec = 0;
  int hidden = 512, latent = 32, stack = 4;
  int voc_dim = 512, voc_inter = 4096, n_voc = 0, n_fft = 640, hop = 160, mel_bins = 80;
  int mel_pitch = 128;         // channel-last mel rows are padded to 128 channels
  FsqConst fsq;

  // resolved device views
  LinearW conv1, conv2;
  std::vector<LayerW> enc_layers, dec_layers;
  const float *enc_ln_g, *enc_ln_b, *dec_ln_g, *dec_ln_b;
  LinearW dn_in, dn_latent, up_from, up_stacked;
  ResUnitW dn_res[3], up_res[3];
  LinearW deconv1_even, deconv1_odd, deconv2;
  LinearW voc_embed, voc_head;
  const float *voc_norm_g, *voc_norm_b, *voc_final_g, *voc_final_b;
  std::vector<VocosBlockW> voc_blocks;
  const float *w_dft, *w_melfb, *w_idft, *win_sq;
  const void* w_dft6 = nullptr;    // [416][6*448] split operand (act dtype) for the tensor-core DFT
  const void* w_idft3 = nullptr;   // [n_fft][3*NP] split operand (act dtype) for the tensor-core iDFT

  int act_type() const { return precision == 1 ? 1 : 0; }   // activation storage: 0 fp32, 1 bf16
  bool x3() const { return precision == 2; }                // fp32 activations, dense contractions as 3-product split-bf16 tcgen05 GEMMs
  std::vector<void*> extra_allocs;                          // device buffers outside the slab (w3 planes)
};

int pack_model(Model& m);       // host only
int upload_model(Model& m, int device);
void free_model(Model& m);

}  // namespace swc

struct swc_model {
  swc::Model m;
};
