// Internal stage functions on channel-last buffers (pipeline.cu) and the workspace arena.
#pragma once
#include <algorithm>

#include "kernels.cuh"
#include "model.h"

namespace swc {

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0, largest = 0;   // largest: biggest single allocation (bounds any GEMM operand)
  bool overflow = false;
  void* alloc(long long bytes) {
    const size_t a = (off + 255) & ~(size_t)255;
    off = a + (size_t)bytes;
    peak = std::max(peak, off);
    largest = std::max(largest, (size_t)bytes);
    if (!base) return nullptr;
    if (off > cap) { overflow = true; return nullptr; }
    return base + a;
  }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
  int check() const {
    if (base && overflow) { set_error("workspace too small: this call needs at least %zu bytes (got %zu)", peak, cap); return -3; }
    return 0;
  }
};

struct Ctx {
  const Model* m = nullptr;
  cudaStream_t s = nullptr;
  Arena ws;
  bool dry = false;          // size the workspace only, launch nothing
  bool force_simt = false;   // debugging: run bf16 operands through the SIMT kernels
  float* hidden_out = nullptr;        // encoder only: (n_layers + 1, nb, 768, T) fp32 channels-first, the input of every layer and the
                                      // final LayerNorm output, rows >= len zeroed (reference modules.py:344-371 output_hidden_states)
  const RaggedTable* rag = nullptr;   // bf16 mode: run the transformer stacks on packed valid tokens (host-known lengths)
};

int transformer_stack(Ctx& c, const std::vector<LayerW>& layers, float* h, const long long* lens, int nb, int T,
                      const RaggedTable* rag = nullptr);
int encoder_cl(Ctx& c, const void* mel_cl, const long long* enc_lens, int nb, int Tm, void* enc_cl);
int downsample_fsq(Ctx& c, const void* enc_cl, const long long* code_lens, int nb, int T4, int* codes, float* zq_cf,
                   float* latent_cf, float* zq_cl);
int upsample_cl(Ctx& c, const float* zq_cl, int nb, int Tc, float* h);
int decoder_cl(Ctx& c, float* h, const long long* lens, int nb, int T, void* mel_cl);
// dense: Tv frames of each of the nb items; vt: packed items of different lengths (see pipeline.cu); wav rows are wav_stride
// floats apart (0 = 160 Tv); extra_rows only enlarges the buffers of a sizing run
int vocos_cl(Ctx& c, const void* mel_cl, int nb, int Tv, float* wav, const RaggedTable* vt = nullptr, long long wav_stride = 0,
             int extra_rows = 0);
int mel_frontend(Ctx& c, const float* wav, long long wav_stride, int wav_cols, const long long* lens, int nb,
                 float* mel_cf, void* mel_cl, long long* mel_lens);
int tokenize_chain(Ctx& c, const float* wav, long long wav_stride, int wav_cols, const long long* lens, int nb,
                   int* codes, float* zq_cf, long long* codes_lens);
int detokenize_chain(Ctx& c, const float* zq_cl, const long long* code_lens, int nb, int Tc, float* wav,
                     long long* out_lens);
int lens_affine_pub(Ctx& c, const long long* in, long long* out, int n, long long mul, long long add, long long div);

}  // namespace swc
