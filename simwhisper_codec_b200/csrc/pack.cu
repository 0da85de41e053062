// Host-side weight packing: reference state_dict (711 tensors) -> the tables the kernels read.
//   * old-style weight norm folded once (w = g * v / ||v||_(in,k)), reference modules.py:30-31
//   * q/k/v projections fused into one [2304,768] operand, q pre-scaled by head_dim^-0.5 (exact: 2^-3)
//   * Conv1d / ConvTranspose1d weights reordered to [out][tap][in] so a convolution is a GEMM whose
//     A operand is the activation shifted by a per-tap row offset (gemm.cuh)
//   * frame stack / unstack (einops 'b d (t s) -> b (d s) t', modules.py:541,624) folded into the
//     column order of in_proj / row order of to_stacked
//   * Hann-windowed DFT (400 -> 201 bins), slaney mel filterbank, windowed inverse DFT (321 -> 640)
// Everything is computed on the host in double and rounded once, so it can be unit-tested without a GPU.
#include <cmath>
#include <cstring>

#include "model.h"

namespace swc {

namespace {

const double kPi = 3.14159265358979323846;

// round-to-nearest-even to bf16 precision, kept as float (exactly representable, so the bf16 upload is lossless)
static float bf16_round(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7f800000u) == 0x7f800000u) return f;
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  std::memcpy(&f, &u, 4);
  return f;
}

struct Packer {
  Model& m;
  std::string err;
  explicit Packer(Model& mm) : m(mm) {}

  const RawTensor* raw(const std::string& k, std::initializer_list<int64_t> shape) {
    auto it = m.raw.find(k);
    if (it == m.raw.end()) { if (err.empty()) err = "missing state_dict key: " + k; return nullptr; }
    const RawTensor& t = it->second;
    std::vector<int64_t> want(shape);
    if (t.shape != want) { if (err.empty()) err = "bad shape for state_dict key: " + k; return nullptr; }
    return &t;
  }
  std::vector<float>& put(const std::string& name, size_t n, bool act) {
    Packed& p = m.tab[name];
    p.host.assign(n, 0.0f);
    p.as_act_type = act;
    return p.host;
  }
  void copy_vec(const std::string& name, const std::string& key, int64_t n) {
    const RawTensor* t = raw(key, {n});
    auto& o = put(name, (size_t)n, false);
    if (t) std::memcpy(o.data(), t->f.data(), sizeof(float) * n);
  }
  // plain Linear [N,K]
  void linear(const std::string& name, const std::string& key, int N, int K, bool bias) {
    const RawTensor* w = raw(key + ".weight", {N, K});
    auto& o = put(name + ".w", (size_t)N * K, true);
    if (w) std::memcpy(o.data(), w->f.data(), sizeof(float) * (size_t)N * K);
    if (bias) copy_vec(name + ".b", key + ".bias", N);
  }
  // folded weight-norm conv weight as double [cout][cin][k]
  std::vector<double> wn(const std::string& key, int cout, int cin, int k) {
    std::vector<double> w((size_t)cout * cin * k, 0.0);
    const RawTensor* v = raw(key + ".weight_v", {cout, cin, k});
    const RawTensor* g = raw(key + ".weight_g", {cout, 1, 1});
    if (!v || !g) return w;
    for (int o = 0; o < cout; ++o) {
      double nrm = 0.0;
      const float* vo = v->f.data() + (size_t)o * cin * k;
      for (int i = 0; i < cin * k; ++i) nrm += (double)vo[i] * vo[i];
      nrm = std::sqrt(nrm);
      const double sc = (double)g->f[o] / nrm;
      for (int i = 0; i < cin * k; ++i) w[(size_t)o * cin * k + i] = vo[i] * sc;
    }
    return w;
  }
  void layer(const std::string& name, const std::string& key) {
    const int D = m.d_model, F = m.ffn;
    copy_vec(name + ".ln1.g", key + ".self_attn_layer_norm.weight", D);
    copy_vec(name + ".ln1.b", key + ".self_attn_layer_norm.bias", D);
    copy_vec(name + ".ln2.g", key + ".final_layer_norm.weight", D);
    copy_vec(name + ".ln2.b", key + ".final_layer_norm.bias", D);
    const RawTensor* q = raw(key + ".self_attn.q_proj.weight", {D, D});
    const RawTensor* k = raw(key + ".self_attn.k_proj.weight", {D, D});
    const RawTensor* v = raw(key + ".self_attn.v_proj.weight", {D, D});
    const RawTensor* qb = raw(key + ".self_attn.q_proj.bias", {D});
    const RawTensor* vb = raw(key + ".self_attn.v_proj.bias", {D});
    auto& w = put(name + ".qkv.w", (size_t)3 * D * D, true);
    auto& b = put(name + ".qkv.b", (size_t)3 * D, false);
    if (q && k && v && qb && vb) {
      const float scale = 1.0f / std::sqrt((float)(D / m.heads));   // 0.125 for head_dim 64
      for (size_t i = 0; i < (size_t)D * D; ++i) {
        w[i] = q->f[i] * scale;
        w[(size_t)D * D + i] = k->f[i];
        w[(size_t)2 * D * D + i] = v->f[i];
      }
      for (int i = 0; i < D; ++i) { b[i] = qb->f[i] * scale; b[D + i] = 0.0f; b[2 * D + i] = vb->f[i]; }
    }
    linear(name + ".out", key + ".self_attn.out_proj", D, D, true);
    linear(name + ".fc1", key + ".fc1", F, D, true);
    linear(name + ".fc2", key + ".fc2", D, F, true);
  }
  void res_units(const std::string& name, const std::string& key) {
    const int H = m.hidden;
    for (int i = 0; i < 3; ++i) {
      const std::string n = name + ".res" + std::to_string(i), k = key + ".res_blocks." + std::to_string(i) + ".block.";
      for (int j : {0, 2}) {
        const std::string a = n + ".act" + std::to_string(j), ka = k + std::to_string(j);
        copy_vec(a + ".alpha", ka + ".act.alpha", H);
        copy_vec(a + ".beta", ka + ".act.beta", H);
        const RawTensor* fu = raw(ka + ".upsample.filter", {1, 1, 12});
        const RawTensor* fd = raw(ka + ".downsample.lowpass.filter", {1, 1, 12});
        auto& ou = put(a + ".fu", 12, false);
        auto& od = put(a + ".fd", 12, false);
        if (fu && fd) for (int t = 0; t < 12; ++t) { ou[t] = fu->f[t]; od[t] = fd->f[t]; }
      }
      {   // dilated k7 conv: [out][tap][in]
        auto w = wn(k + "1", H, H, 7);
        auto& o = put(n + ".conv7.w", (size_t)H * 7 * H, true);
        for (int oc = 0; oc < H; ++oc)
          for (int c = 0; c < H; ++c)
            for (int t = 0; t < 7; ++t) o[((size_t)oc * 7 + t) * H + c] = (float)w[((size_t)oc * H + c) * 7 + t];
        copy_vec(n + ".conv7.b", k + "1.bias", H);
      }
      {
        auto w = wn(k + "3", H, H, 1);
        auto& o = put(n + ".conv1.w", (size_t)H * H, true);
        for (size_t t = 0; t < (size_t)H * H; ++t) o[t] = (float)w[t];
        copy_vec(n + ".conv1.b", k + "3.bias", H);
      }
    }
  }
};

// slaney mel scale (transformers.audio_utils.hertz_to_mel / mel_to_hertz, mel_scale="slaney")
double hz_to_mel(double f) {
  return f >= 1000.0 ? 15.0 + std::log(f / 1000.0) * (27.0 / std::log(6.4)) : 3.0 * f / 200.0;
}
double mel_to_hz(double mel) {
  return mel >= 15.0 ? 1000.0 * std::exp((std::log(6.4) / 27.0) * (mel - 15.0)) : 200.0 * mel / 3.0;
}

}  // namespace

int pack_model(Model& m) {
  Packer P(m);
  // ---- architecture from tensor inventory
  m.n_enc = m.n_dec = m.n_voc = 0;
  while (m.raw.count("acoustic_encoder.layers." + std::to_string(m.n_enc) + ".fc1.weight")) ++m.n_enc;
  while (m.raw.count("acoustic_decoder.layers." + std::to_string(m.n_dec) + ".fc1.weight")) ++m.n_dec;
  while (m.raw.count("vocos.backbone.convnext." + std::to_string(m.n_voc) + ".gamma")) ++m.n_voc;
  SWC_REQUIRE(m.n_enc > 0 && m.n_dec > 0 && m.n_voc > 0, "state_dict has no encoder/decoder/vocos layers (%d/%d/%d)", m.n_enc, m.n_dec, m.n_voc);
  const int D = m.d_model, MP = m.mel_pitch, MB = m.mel_bins, H = m.hidden;

  // ---- encoder stem: Conv1d weights [out][in][k] -> [out][k][in(pad)]
  if (const RawTensor* w = P.raw("acoustic_encoder.conv1.weight", {D, MB, 3})) {
    auto& o = P.put("enc.conv1.w", (size_t)D * 3 * MP, true);
    for (int oc = 0; oc < D; ++oc)
      for (int c = 0; c < MB; ++c)
        for (int t = 0; t < 3; ++t) o[((size_t)oc * 3 + t) * MP + c] = w->f[((size_t)oc * MB + c) * 3 + t];
  }
  P.copy_vec("enc.conv1.b", "acoustic_encoder.conv1.bias", D);
  if (const RawTensor* w = P.raw("acoustic_encoder.conv2.weight", {D, D, 3})) {
    auto& o = P.put("enc.conv2.w", (size_t)D * 3 * D, true);
    for (int oc = 0; oc < D; ++oc)
      for (int c = 0; c < D; ++c)
        for (int t = 0; t < 3; ++t) o[((size_t)oc * 3 + t) * D + c] = w->f[((size_t)oc * D + c) * 3 + t];
  }
  P.copy_vec("enc.conv2.b", "acoustic_encoder.conv2.bias", D);
  for (int i = 0; i < m.n_enc; ++i) P.layer("enc.L" + std::to_string(i), "acoustic_encoder.layers." + std::to_string(i));
  P.copy_vec("enc.ln.g", "acoustic_encoder.layer_norm.weight", D);
  P.copy_vec("enc.ln.b", "acoustic_encoder.layer_norm.bias", D);

  // ---- down-sampler: stacked channel d*4+s -> operand column s*768+d (4 consecutive token rows)
  {
    const int S = m.stack;
    auto w = P.wn("downsample.in_proj", H, D * S, 1);
    auto& o = P.put("dn.in.w", (size_t)H * D * S, true);
    for (int oc = 0; oc < H; ++oc)
      for (int d = 0; d < D; ++d)
        for (int s = 0; s < S; ++s) o[(size_t)oc * D * S + s * D + d] = (float)w[(size_t)oc * D * S + d * S + s];
    P.copy_vec("dn.in.b", "downsample.in_proj.bias", H);
    P.res_units("dn", "downsample");
    auto wl = P.wn("downsample.to_latent", m.latent, H, 1);
    auto& ol = P.put("dn.latent.w", (size_t)m.latent * H, true);
    for (size_t t = 0; t < ol.size(); ++t) ol[t] = (float)wl[t];
    P.copy_vec("dn.latent.b", "downsample.to_latent.bias", m.latent);
  }
  // ---- FSQ constants (reference quantizer.py:129-140, 155, 161), fp32 arithmetic like torch
  if (const RawTensor* lv = P.raw("quantizer.fsqs.0.num_levels", {1, 4, 1})) {
    const RawTensor* bs = P.raw("quantizer.fsqs.0.dim_base_index", {1, 4, 1});
    const float one_minus_eps = (float)(1.0 - 1e-3);
    for (int d = 0; d < 4 && bs; ++d) {
      const int L = lv->i[d];
      m.fsq.levels[d] = L;
      m.fsq.base[d] = bs->i[d];
      m.fsq.scale[d] = ((float)(L - 1) / 2.0f) * one_minus_eps;
      m.fsq.offset[d] = (L % 2 == 0) ? 0.5f : 0.0f;
      m.fsq.shift[d] = tanf(m.fsq.offset[d] / m.fsq.scale[d]);
      m.fsq.half[d] = (float)(L / 2);
    }
    auto& o = P.put("fsq.const", 16, false);
    for (int d = 0; d < 4; ++d) { o[d] = m.fsq.scale[d]; o[4 + d] = m.fsq.offset[d]; o[8 + d] = m.fsq.shift[d]; o[12 + d] = m.fsq.half[d]; }
  }
  // ---- up-sampler
  {
    const int S = m.stack;
    auto wf = P.wn("upsample.from_latent", H, m.latent, 1);
    auto& of = P.put("up.from.w", (size_t)H * m.latent, false);       // stays fp32 (K=32 SIMT GEMM)
    for (size_t t = 0; t < of.size(); ++t) of[t] = (float)wf[t];
    P.copy_vec("up.from.b", "upsample.from_latent.bias", H);
    P.res_units("up", "upsample");
    auto w = P.wn("upsample.to_stacked", D * S, H, 1);
    auto& o = P.put("up.stacked.w", (size_t)D * S * H, true);
    auto& ob = P.put("up.stacked.b", (size_t)D * S, false);
    const RawTensor* b = P.raw("upsample.to_stacked.bias", {D * S});
    for (int d = 0; d < D; ++d)
      for (int s = 0; s < S; ++s) {
        for (int c = 0; c < H; ++c) o[((size_t)s * D + d) * H + c] = (float)w[((size_t)d * S + s) * H + c];
        if (b) ob[s * D + d] = b->f[d * S + s];
      }
  }
  // ---- decoder
  for (int i = 0; i < m.n_dec; ++i) P.layer("dec.L" + std::to_string(i), "acoustic_decoder.layers." + std::to_string(i));
  P.copy_vec("dec.ln.g", "acoustic_decoder.layer_norm.weight", D);
  P.copy_vec("dec.ln.b", "acoustic_decoder.layer_norm.bias", D);
  // ConvTranspose1d weights are [in][out][k]:  z[2t+k] += W1[:, :, k]^T h[t]
  //   even rows 2t   = W1[..,2]^T h[t-1] + W1[..,0]^T h[t]      odd rows 2t+1 = W1[..,1]^T h[t]
  if (const RawTensor* w = P.raw("acoustic_decoder.deconv1.weight", {D, D, 3})) {
    auto& oe = P.put("dec.deconv1e.w", (size_t)D * 2 * D, true);
    auto& oo = P.put("dec.deconv1o.w", (size_t)D * D, true);
    for (int oc = 0; oc < D; ++oc)
      for (int c = 0; c < D; ++c) {
        const float* src = w->f.data() + ((size_t)c * D + oc) * 3;
        oe[(size_t)oc * 2 * D + c] = src[2];
        oe[(size_t)oc * 2 * D + D + c] = src[0];
        oo[(size_t)oc * D + c] = src[1];
      }
  }
  P.copy_vec("dec.deconv1.b", "acoustic_decoder.deconv1.bias", D);
  //   o[u] = sum_k W2[:, :, k]^T z[u-k]; output channels padded 80 -> 128 (zero rows) so the result is
  //   directly the padded channel-last operand of the Vocos embed convolution
  if (const RawTensor* w = P.raw("acoustic_decoder.deconv2.weight", {D, MB, 3})) {
    auto& o = P.put("dec.deconv2.w", (size_t)MP * 3 * D, true);
    for (int oc = 0; oc < MB; ++oc)
      for (int c = 0; c < D; ++c)
        for (int t = 0; t < 3; ++t) o[((size_t)oc * 3 + t) * D + c] = w->f[((size_t)c * MB + oc) * 3 + t];
    auto& ob = P.put("dec.deconv2.b", (size_t)MP, false);
    if (const RawTensor* b = P.raw("acoustic_decoder.deconv2.bias", {MB})) for (int i = 0; i < MB; ++i) ob[i] = b->f[i];
  }
  // ---- Vocos
  {
    const int V = m.voc_dim, I = m.voc_inter;
    if (const RawTensor* w = P.raw("vocos.backbone.embed.weight", {V, MB, 7})) {
      auto& o = P.put("voc.embed.w", (size_t)V * 7 * MP, true);
      for (int oc = 0; oc < V; ++oc)
        for (int c = 0; c < MB; ++c)
          for (int t = 0; t < 7; ++t) o[((size_t)oc * 7 + t) * MP + c] = w->f[((size_t)oc * MB + c) * 7 + t];
    }
    P.copy_vec("voc.embed.b", "vocos.backbone.embed.bias", V);
    P.copy_vec("voc.norm.g", "vocos.backbone.norm.weight", V);
    P.copy_vec("voc.norm.b", "vocos.backbone.norm.bias", V);
    for (int i = 0; i < m.n_voc; ++i) {
      const std::string n = "voc.B" + std::to_string(i), k = "vocos.backbone.convnext." + std::to_string(i);
      if (const RawTensor* w = P.raw(k + ".dwconv.weight", {V, 1, 7})) {
        auto& o = P.put(n + ".dw.w", (size_t)7 * V, false);
        for (int c = 0; c < V; ++c)
          for (int t = 0; t < 7; ++t) o[(size_t)t * V + c] = w->f[(size_t)c * 7 + t];
      }
      P.copy_vec(n + ".dw.b", k + ".dwconv.bias", V);
      P.copy_vec(n + ".ln.g", k + ".norm.weight", V);
      P.copy_vec(n + ".ln.b", k + ".norm.bias", V);
      P.copy_vec(n + ".gamma", k + ".gamma", V);
      P.linear(n + ".pw1", k + ".pwconv1", I, V, true);
      P.linear(n + ".pw2", k + ".pwconv2", V, I, true);
    }
    P.copy_vec("voc.final.g", "vocos.backbone.final_layer_norm.weight", V);
    P.copy_vec("voc.final.b", "vocos.backbone.final_layer_norm.bias", V);
    // head: interleave (log-magnitude_j, phase_j) rows; pad 642 -> 704 zero rows (their spectrum is (1, 0): finite, and it meets zero iDFT columns)
    const int NB = m.n_fft / 2 + 1, NO = 2 * NB, NP = (NO + 63) / 64 * 64;   // 642 -> 704: whole 64-wide K slabs for the iDFT GEMM
    if (const RawTensor* w = P.raw("vocos.head.out.weight", {NO, V})) {
      const RawTensor* b = P.raw("vocos.head.out.bias", {NO});
      auto& o = P.put("voc.head.w", (size_t)NP * V, true);
      auto& ob = P.put("voc.head.b", (size_t)NP, false);
      for (int j = 0; j < NB && b; ++j) {
        std::memcpy(&o[(size_t)(2 * j) * V], &w->f[(size_t)j * V], sizeof(float) * V);
        std::memcpy(&o[(size_t)(2 * j + 1) * V], &w->f[(size_t)(NB + j) * V], sizeof(float) * V);
        ob[2 * j] = b->f[j];
        ob[2 * j + 1] = b->f[NB + j];
      }
    }
    // windowed inverse real DFT as a [640][656] operand on interleaved (Re_j, Im_j)  
    if (const RawTensor* win = P.raw("vocos.head.istft.window", {m.n_fft})) {
      const int N = m.n_fft;
      auto& o = P.put("voc.idft.w", (size_t)N * NP, false);
      auto& w2 = P.put("voc.win_sq", (size_t)N, false);
      for (int n = 0; n < N; ++n) {
        const double wn = (double)win->f[n] / N;
        for (int j = 0; j < NB; ++j) {
          const double cj = (j == 0 || j == N / 2) ? 1.0 : 2.0;
          const double ang = 2.0 * kPi * (double)(((long long)j * n) % N) / N;
          o[(size_t)n * NP + 2 * j] = (float)(wn * cj * std::cos(ang));
          o[(size_t)n * NP + 2 * j + 1] = (j == 0 || j == N / 2) ? 0.0f : (float)(-wn * cj * std::sin(ang));
        }
        w2[n] = win->f[n] * win->f[n];
      }
      // bf16 tensor-core form of the same operand: w = w1 + w2 (+ 2^-17 relative), rows [w1 | w2 | w1] so that one GEMM
      // over the taps (s1, s1, s2) of a split spectrum s = s1 + s2 accumulates s1 w1 + s1 w2 + s2 w1 in fp32
      auto& o3 = P.put("voc.idft.w3", (size_t)N * 3 * NP, true);
      for (int n = 0; n < N; ++n)
        for (int k = 0; k < NP; ++k) {
          const float w = o[(size_t)n * NP + k];
          const float w1 = bf16_round(w), w2r = bf16_round(w - w1);
          o3[(size_t)n * 3 * NP + k] = w1;
          o3[(size_t)n * 3 * NP + NP + k] = w2r;
          o3[(size_t)n * 3 * NP + 2 * NP + k] = w1;
        }
    }
  }
  // ---- log-mel tables (reference feature_extractor.py:50-58, 92-101)
  {
    const int NF = 400, NBIN = 201, NROW = 416, KP = 208;
    auto& o = P.put("mel.dft.w", (size_t)NROW * NF, false);
    for (int k = 0; k < NBIN; ++k)
      for (int n = 0; n < NF; ++n) {
        const double win = 0.5 - 0.5 * std::cos(2.0 * kPi * n / NF);    // periodic Hann
        const double ang = 2.0 * kPi * (double)(((long long)k * n) % NF) / NF;
        o[(size_t)(2 * k) * NF + n] = (float)(win * std::cos(ang));
        o[(size_t)(2 * k + 1) * NF + n] = (float)(-win * std::sin(ang));
      }
    // bf16 tensor-core form: w = w1 + w2 + w3 (three bf16 terms = 24 mantissa bits); rows hold the six products' W factors
    // [w1 | w2 | w1 | w3 | w2 | w1] (each zero-padded 400 -> 448 columns) against the frame planes (a1,a1,a2,a1,a2,a3)
    {
      const int KT = 448;
      auto& o6 = P.put("mel.dft.w6", (size_t)NROW * 6 * KT, true);
      static const int which[6] = {0, 1, 0, 2, 1, 0};
      for (int r = 0; r < NROW; ++r)
        for (int n = 0; n < NF; ++n) {
          const float w = o[(size_t)r * NF + n];
          float part[3];
          part[0] = bf16_round(w);
          part[1] = bf16_round(w - part[0]);
          part[2] = bf16_round(w - part[0] - part[1]);
          for (int t = 0; t < 6; ++t) o6[(size_t)r * 6 * KT + (size_t)t * KT + n] = part[which[t]];
        }
    }
    auto& fb = P.put("mel.fb.w", (size_t)MB * KP, false);
    const double mlo = hz_to_mel(0.0), mhi = hz_to_mel(8000.0);
    std::vector<double> hz(MB + 2);
    for (int i = 0; i < MB + 2; ++i) hz[i] = mel_to_hz(mlo + (mhi - mlo) * i / (MB + 1));
    for (int mi = 0; mi < MB; ++mi) {
      const double enorm = 2.0 / (hz[mi + 2] - hz[mi]);
      for (int k = 0; k < NBIN; ++k) {
        const double f = 8000.0 * k / (NBIN - 1);
        const double down = (f - hz[mi]) / (hz[mi + 1] - hz[mi]);
        const double up = (hz[mi + 2] - f) / (hz[mi + 2] - hz[mi + 1]);
        const double v = std::fmax(0.0, std::fmin(down, up)) * enorm;
        fb[(size_t)mi * KP + k] = (float)v;
      }
    }
  }
  SWC_REQUIRE(P.err.empty(), "%s", P.err.c_str());
  m.packed = true;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// upload + view resolution
// ---------------------------------------------------------------------------------------------
namespace {
uint16_t f32_to_bf16_rn(float f) {
  uint32_t u;
  std::memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
}  // namespace

static void* g_slab_unused = nullptr;

int upload_model(Model& m, int device) {
  SWC_REQUIRE(m.packed, "upload before pack");
  DeviceGuard guard(device);            // the caller's current device is restored on every return path
  SWC_REQUIRE(guard.ok(), "cudaSetDevice(%d) failed", device);
  cudaDeviceProp prop;
  SWC_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  SWC_REQUIRE(prop.major == 10, "libswc is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
  m.num_sms = prop.multiProcessorCount;
  m.device = device;
  const bool bf = m.precision == 1;
  // stored as bf16: every GEMM operand in bf16 mode; in bf16x3 mode only the two pre-split DFT operands (their entries are
  // bf16-exact terms w1, w2, w3 of the fp32 tables, consumed by the plain tensor-core GEMM)
  auto as_bf16 = [&](const std::string& name, const Packed& p) {
    return p.as_act_type && (bf || (m.x3() && (name == "mel.dft.w6" || name == "voc.idft.w3")));
  };
  size_t total = 0;
  for (auto& kv : m.tab) {
    const size_t bytes = kv.second.host.size() * (as_bf16(kv.first, kv.second) ? 2 : 4);
    total += (bytes + 255) / 256 * 256;
  }
  char* slab = nullptr;
  SWC_CHECK_CUDA(cudaMalloc(&slab, total));
  m.tab["__slab__"].dev = slab;
  size_t off = 0;
  std::vector<uint16_t> tmp;
  for (auto& kv : m.tab) {
    if (kv.first == "__slab__") continue;
    Packed& p = kv.second;
    p.dev = slab + off;
    if (as_bf16(kv.first, p)) {
      tmp.resize(p.host.size());
      for (size_t i = 0; i < tmp.size(); ++i) tmp[i] = f32_to_bf16_rn(p.host[i]);
      SWC_CHECK_CUDA(cudaMemcpy(p.dev, tmp.data(), tmp.size() * 2, cudaMemcpyHostToDevice));
      off += (tmp.size() * 2 + 255) / 256 * 256;
    } else {
      SWC_CHECK_CUDA(cudaMemcpy(p.dev, p.host.data(), p.host.size() * 4, cudaMemcpyHostToDevice));
      off += (p.host.size() * 4 + 255) / 256 * 256;
    }
    if (!(m.x3() && p.as_act_type)) std::vector<float>().swap(p.host);     // bf16x3: L() below still needs the fp32 rows
  }
  m.raw.clear();

  auto F = [&](const std::string& n) -> const float* { return (const float*)m.tab.at(n).dev; };
  auto L = [&](const std::string& n, int N, int rows, int K, bool bias = true) {
    LinearW l;
    l.w = m.tab.at(n + ".w").dev;
    l.bias = bias ? F(n + ".b") : nullptr;
    l.N = N; l.w_rows = rows; l.K = K;
    if (m.x3()) {      // split every weight row into bf16 planes: (hi | lo | hi), the operand of the three-product GEMM
      Packed& pw = m.tab.at(n + ".w");
      if ((long long)pw.host.size() == (long long)rows * K) {
        std::vector<uint16_t> w3((size_t)rows * 3 * K);
        for (int r = 0; r < rows; ++r) {
          for (int k = 0; k < K; ++k) {
            const float v = pw.host[(size_t)r * K + k];
            const uint16_t hi = f32_to_bf16_rn(v);
            uint32_t hb = (uint32_t)hi << 16;
            float hf;
            std::memcpy(&hf, &hb, 4);
            const uint16_t lo = f32_to_bf16_rn(v - hf);
            uint16_t* o = w3.data() + (size_t)r * 3 * K;
            o[k] = hi; o[K + k] = lo; o[2 * K + k] = hi;
          }
        }
        void* dev = nullptr;
        if (cudaMalloc(&dev, w3.size() * 2) == cudaSuccess &&
            cudaMemcpy(dev, w3.data(), w3.size() * 2, cudaMemcpyHostToDevice) == cudaSuccess) {
          m.extra_allocs.push_back(dev);
          l.w3 = dev;
        }
      }
      std::vector<float>().swap(pw.host);
    }
    return l;
  };
  const int D = m.d_model, MP = m.mel_pitch, H = m.hidden, V = m.voc_dim, I = m.voc_inter;
  m.conv1 = L("enc.conv1", D, D, 3 * MP);
  m.conv2 = L("enc.conv2", D, D, 3 * D);
  auto layers = [&](const std::string& pfx, int n, std::vector<LayerW>& out) {
    out.resize(n);
    for (int i = 0; i < n; ++i) {
      const std::string p = pfx + ".L" + std::to_string(i);
      LayerW& l = out[i];
      l.ln1_g = F(p + ".ln1.g"); l.ln1_b = F(p + ".ln1.b"); l.ln2_g = F(p + ".ln2.g"); l.ln2_b = F(p + ".ln2.b");
      l.qkv = L(p + ".qkv", 3 * D, 3 * D, D);
      l.out = L(p + ".out", D, D, D);
      l.fc1 = L(p + ".fc1", m.ffn, m.ffn, D);
      l.fc2 = L(p + ".fc2", D, D, m.ffn);
    }
  };
  layers("enc", m.n_enc, m.enc_layers);
  layers("dec", m.n_dec, m.dec_layers);
  m.enc_ln_g = F("enc.ln.g"); m.enc_ln_b = F("enc.ln.b"); m.dec_ln_g = F("dec.ln.g"); m.dec_ln_b = F("dec.ln.b");
  auto res = [&](const std::string& pfx, ResUnitW* r) {
    const int dil[3] = {1, 3, 9};
    for (int i = 0; i < 3; ++i) {
      const std::string p = pfx + ".res" + std::to_string(i);
      r[i].a0 = F(p + ".act0.alpha"); r[i].b0 = F(p + ".act0.beta"); r[i].fu0 = F(p + ".act0.fu"); r[i].fd0 = F(p + ".act0.fd");
      r[i].a2 = F(p + ".act2.alpha"); r[i].b2 = F(p + ".act2.beta"); r[i].fu2 = F(p + ".act2.fu"); r[i].fd2 = F(p + ".act2.fd");
      r[i].conv7 = L(p + ".conv7", H, H, 7 * H);
      r[i].conv1 = L(p + ".conv1", H, H, H);
      r[i].dilation = dil[i];
    }
  };
  m.dn_in = L("dn.in", H, H, D * m.stack);
  res("dn", m.dn_res);
  m.dn_latent = L("dn.latent", m.latent, m.latent, H);
  m.up_from = L("up.from", H, H, m.latent);
  res("up", m.up_res);
  m.up_stacked = L("up.stacked", D * m.stack, D * m.stack, H);
  m.deconv1_even = L("dec.deconv1e", D, D, 2 * D, false);
  m.deconv1_odd = L("dec.deconv1o", D, D, D, false);
  m.deconv1_even.bias = m.deconv1_odd.bias = F("dec.deconv1.b");
  m.deconv2 = L("dec.deconv2", MP, MP, 3 * D);
  m.voc_embed = L("voc.embed", V, V, 7 * MP);
  m.voc_norm_g = F("voc.norm.g"); m.voc_norm_b = F("voc.norm.b");
  m.voc_final_g = F("voc.final.g"); m.voc_final_b = F("voc.final.b");
  m.voc_blocks.resize(m.n_voc);
  for (int i = 0; i < m.n_voc; ++i) {
    const std::string p = "voc.B" + std::to_string(i);
    VocosBlockW& b = m.voc_blocks[i];
    b.dw_w = F(p + ".dw.w"); b.dw_b = F(p + ".dw.b"); b.ln_g = F(p + ".ln.g"); b.ln_b = F(p + ".ln.b"); b.gamma = F(p + ".gamma");
    b.pw1 = L(p + ".pw1", I, I, V);
    b.pw2 = L(p + ".pw2", V, V, I);
  }
  const int NP = ((m.n_fft + 2) + 63) / 64 * 64;
  m.voc_head = L("voc.head", NP, NP, V);
  m.w_idft = F("voc.idft.w"); m.win_sq = F("voc.win_sq");
  m.w_idft3 = m.tab.count("voc.idft.w3") ? m.tab["voc.idft.w3"].dev : nullptr;
  m.w_dft = F("mel.dft.w"); m.w_melfb = F("mel.fb.w");
  m.w_dft6 = m.tab.count("mel.dft.w6") ? m.tab["mel.dft.w6"].dev : nullptr;
  for (auto& kv : m.tab) std::vector<float>().swap(kv.second.host);
  m.uploaded = true;
  (void)g_slab_unused;
  return 0;
}

void free_model(Model& m) {
  auto it = m.tab.find("__slab__");
  if (it != m.tab.end() && it->second.dev) cudaFree(it->second.dev);
  for (void* p : m.extra_allocs) cudaFree(p);
  m.extra_allocs.clear();
  m.tab.clear();
  m.raw.clear();
  m.uploaded = false;
}

}  // namespace swc
