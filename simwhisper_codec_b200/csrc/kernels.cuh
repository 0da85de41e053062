// Launch wrappers of the bandwidth-bound / small kernels (kernels.cu, attention_simt.cu).
// dtype codes: 0 = fp32, 1 = bf16.  All pointers are device pointers; nothing here synchronises.
#pragma once
#include "gemm.cuh"

namespace swc {

// LayerNorm over the last dim (C = 768 or 512) of channel-last fp32 rows, optionally fused with the residual
// add that precedes it:  h' = in + delta (written to h_out, which may alias `in`), out = LN(h').
// Output has t_out >= t_in rows per batch; rows t >= t_in or t >= lens[b] (if lens) are written as zero.
int layernorm(const float* in, const float* delta, float* h_out, void* out, int out_type, const float* gamma,
              const float* beta, float eps, int nb, int t_in, int t_out, int C, const long long* lens, cudaStream_t s);

// Vocos ConvNeXt front half: depthwise conv k7 (pad 3) + bias + LayerNorm(eps) over C=512 on x (+ delta) fp32
// (nb,T,C).  With delta, the updated stream x + delta is also written to x_out (must not alias x).
int dwconv7_ln(const float* x, const float* delta, float* x_out, const float* w7c /*[7][C]*/, const float* bias,
               const float* gamma, const float* beta, float eps, void* out, int out_type, int nb, int T, int C,
               cudaStream_t s);

// Anti-aliased SnakeBeta along time of channel-last (nb,T,C): 2x up (12 taps) -> snake -> 2x down.
int aa_snake(const void* in, int in_type, void* out, int out_type, const float* taps_up, const float* taps_dn,
             const float* alpha_log, const float* beta_log, int nb, int T, int C, cudaStream_t s);

// FSQ on channels-first fp32 latents (module-level quantizer.forward).
int fsq_encode_cf(const float* latent_cf, const long long* lens, int nb, int T, const FsqConst& c,
                  float* zq_cf, int* codes, float* zq_cl, cudaStream_t s);
// codes (G=8, nb, T) int32/int64 -> zq (masked beyond lens)
int fsq_decode(const void* codes, int codes_i64, const long long* lens, int nb, int T, const FsqConst& c,
               float* zq_cf, float* zq_cl, cudaStream_t s);

// layout conversion: channels-first fp32 (nb,C,T) <-> channel-last (nb,T_rows,Cpitch)
int cf_to_cl(const float* in, void* out, int out_type, int nb, int C, int T, int t_rows, int c_pitch, cudaStream_t s);
// lens (optional): rows t >= lens[b] are written as zero
int cl_to_cf(const void* in, int in_type, float* out, int nb, int C, int T, long long in_batch_stride, int c_pitch, cudaStream_t s,
             const long long* lens = nullptr);

// log-mel helpers
int mel_pad(const float* wav, long long wav_stride, int wav_cols, const long long* lens, int nb,
            float* padded /*(nb, 480400)*/, long long* mel_lens, float* item_max, cudaStream_t s);
// framed + split signal for the tensor-core DFT: frames (nb, 3000, 3, 448) bf16 with x = a1 + a2 + a3 per sample
// (columns 400..447 of every plane are zero); reads the reflect-padded signal written by mel_pad
int mel_frames_split(const float* padded, int nb, bf16* frames, cudaStream_t s);
int mel_finalize(const float* logmel /*(nb,3000,80)*/, const float* item_max, int nb, float* mel_cf,
                 void* mel_cl, int cl_type, int cl_pitch, cudaStream_t s);

// iSTFT overlap-add ("same" padding): frames (nb,T,640) fp32 -> wav (nb,160T) fp32
// wav rows are wav_stride floats apart (>= 160 T)
int istft_ola(const float* frames, const float* win_sq /*[640]*/, int nb, int T, float* wav, long long wav_stride, cudaStream_t s);

// ---- ragged (packed valid tokens) transformer path: per-item token counts known on the host --------------------
constexpr int kMaxRagged = 256;      // items per ragged call: the table travels as a 2 KB kernel parameter (callers split larger batches)
// bf16x3 mode: fp32 rows (nb, rows, cols; row / batch strides in elements) -> bf16 planes (nb, rows, 2 cols) = (hi | lo),
// hi = bf16(x), lo = bf16(x - hi).  cols % 8 == 0.
int split_bf16_planes(const float* in, long long row_stride, long long batch_stride, int nb, int rows, int cols, bf16* planes,
                      cudaStream_t s);

// (rows, hi | lo) bf16 planes of `cols` columns each -> fp32 rows hi + lo
int merge_bf16_planes(const bf16* planes, long long rows, int cols, float* out, cudaStream_t s);

struct RaggedTable {          // passed by value to kernels (2 KB)
  int nb = 0;
  int t_max = 0;              // longest item
  int total = 0;              // sum of len
  int len[kMaxRagged];
  int off[kMaxRagged + 1];    // off[b] = first packed row of item b
};
// packed[off[b] + t] = padded[b, t] for t < len[b]   (fp32 rows of C channels; padded has t_pad rows per item)
int pack_rows(const float* padded, float* packed, const RaggedTable& tab, int t_pad, int C, cudaStream_t s);
// padded[b, t] = t < len[b] ? packed[off[b] + t] : 0 for t < t_pad   (dtype 0 fp32 / 1 bf16)
int unpack_rows(const void* packed, void* padded, int dtype, const RaggedTable& tab, int t_pad, int C, cudaStream_t s);

// packed rows with zeroed gaps: packed[off[b] + t] = padded[b, t] for t < len[b], rows off[b] + len[b] .. off[b + 1] - 1 = 0
// (rows of row_bytes bytes, a multiple of 16; off[] may leave gaps between items)
int pack_rows_gap(const void* padded, void* packed, const RaggedTable& tab, int t_pad, int row_bytes, cudaStream_t s);
// dwconv7 + LayerNorm and the iSTFT overlap-add over packed items (item b = rows [off[b], off[b] + len[b]); rows outside an
// item are zero padding, exactly as at the ends of a dense (nb, T) batch)
int dwconv7_ln_ragged(const float* x, const float* w7c, const float* bias, const float* gamma, const float* beta, float eps,
                      void* out, int out_type, const RaggedTable& tab, int C, cudaStream_t s);
int istft_ola_ragged(const float* frames, const float* win_sq, const RaggedTable& tab, float* wav, long long wav_stride, cudaStream_t s);

// fp32 SIMT flash attention over fused qkv rows (nb*T, 3*H*64): q pre-scaled. Keys >= lens[b] masked.
int attention_simt(const void* qkv, int type, void* out, const long long* lens, int nb, int T, int H, cudaStream_t s);
// bf16x3 mode: fp32-class flash attention on tcgen05 (attention_tc_x3.cu) over the (hi | lo) bf16 planes of the fp32 qkv rows
// (rows of 2 x 3*H*64 bf16); the output rows are (hi | lo) planes of H*64 columns each — the operand of the out_proj GEMM
int attention_tc_x3(const bf16* planes, bf16* out_planes, const long long* lens, int nb, int T, int H, int num_sms, cudaStream_t s);
// the same over packed rows (planes and out hold tab.total rows; item b = rows [off[b], off[b] + len[b]))
int attention_tc_x3_ragged(const bf16* planes, bf16* out_planes, const RaggedTable& tab, int H, int num_sms, cudaStream_t s);
// bf16 flash attention on tcgen05 / TMEM / TMA (attention_tc.cu), same contract
int attention_tc(const bf16* qkv, bf16* out, const long long* lens, int nb, int T, int H, int num_sms, cudaStream_t s);
// the same kernel over packed rows: item b owns rows [off[b], off[b] + len[b]) of qkv / out; no padded rows exist
int attention_tc_ragged(const bf16* qkv, bf16* out, const RaggedTable& tab, int H, int num_sms, cudaStream_t s);

// misc
int fill_f32(float* p, float v, long long n, cudaStream_t s);
int convert_f32_to_bf16(const float* in, bf16* out, long long n, cudaStream_t s);

}  // namespace swc
