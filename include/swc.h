/*
 * swc.h — C ABI of the B200-native SimWhisper-Codec hot path (libswc.so).
 *
 * The reference (ZhangXinWhut/SimWhisper-Codec) is pure Python and has no FFI layer; the boundary
 * this library replaces is the set of torch modules behind `audiocodec/model.py::AudioCodec`.
 * Each entry point below names the reference interface it stands in for.  The Python host side
 * (simwhisper_codec_b200/audiocodec/model.py) binds these with ctypes and mirrors the reference's
 * class/method names; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - every data pointer is a DEVICE pointer unless the name says `host`;
 *   - "cf" tensors are the reference's channels-first fp32 layout (B, C, T), contiguous;
 *   - lengths are device int64[B] (the reference passes LongTensors on the module's device);
 *   - `stream` is a cudaStream_t; nothing here allocates, frees or synchronises after
 *     swc_model_finalize(): the caller owns all buffers, scratch comes from `workspace`;
 *   - return 0 on success, negative on error; swc_last_error() returns the message (the Python
 *     wrapper raises RuntimeError with it).  There is no CPU fallback: without a CUDA device every
 *     compute entry point fails.
 */
#ifndef SWC_H_
#define SWC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct swc_model swc_model;

/* BF16X3: fp32 activations and epilogues as in FP32, dense contractions on the tensor cores as three bf16 products
   (a_hi w_hi + a_hi w_lo + a_lo w_hi, fp32 accumulation: relative error ~2^-17); attention, mel and iSTFT stay fp32 */
enum { SWC_PRECISION_FP32 = 0, SWC_PRECISION_BF16 = 1, SWC_PRECISION_BF16X3 = 2 };
enum { SWC_DTYPE_F32 = 0, SWC_DTYPE_I32 = 1 };

/* stages, for swc_workspace_bytes() */
enum {
  SWC_STAGE_MEL = 0,
  SWC_STAGE_ENCODER = 1,
  SWC_STAGE_DOWNSAMPLE = 2,
  SWC_STAGE_QUANTIZER = 3,
  SWC_STAGE_UPSAMPLE = 4,
  SWC_STAGE_DECODER = 5,
  SWC_STAGE_VOCOS = 6,
  SWC_STAGE_TOKENIZE = 7,    /* mel + encoder + downsample + FSQ */
  SWC_STAGE_DETOKENIZE = 8,  /* FSQ decode + upsample + decoder + Vocos/iSTFT */
  SWC_STAGE_FORWARD = 9      /* AudioCodec.forward: mel features in, audio out */
};

int swc_version(void);
const char* swc_last_error(void);

/* ---- model lifecycle: replaces AudioCodec.__init__ + load_state_dict (reference
 *      audiocodec/model.py:15-57, 375-396).  Tensors are passed under their reference
 *      state_dict keys (711 entries; old-style weight-norm weight_g/weight_v pairs included). ---- */
int swc_model_create(swc_model** out, int precision);
int swc_model_set_tensor(swc_model* m, const char* key, const void* host_data, int dtype,
                         const int64_t* shape, int ndim);
/* fold weight-norm, fuse QKV (+ pre-scale q), reorder conv taps, build DFT/iDFT/mel tables (host) */
int swc_model_pack(swc_model* m);
/* copy one packed table to host memory as fp32 (test hook; valid after swc_model_pack) */
int64_t swc_model_packed_numel(const swc_model* m, const char* name);
int swc_model_get_packed(const swc_model* m, const char* name, float* host_out, int64_t numel);
/* swc_model_pack() if needed, then upload to `device` (cudaMalloc happens only here) */
int swc_model_finalize(swc_model* m, int device);
void swc_model_destroy(swc_model* m);

/* scratch needed by one call of `stage` with `batch` items of `frames` time steps at the stage's
 * INPUT rate (mel frames for ENCODER/FORWARD, tokens for DOWNSAMPLE/DECODER, code frames for
 * QUANTIZER/UPSAMPLE/DETOKENIZE, mel frames for VOCOS; ignored for MEL/TOKENIZE = 3000) */
size_t swc_workspace_bytes(const swc_model* m, int stage, int batch, int frames);

/* ---- MelFeatureExtractor.__call__ (reference audiocodec/nn/feature_extractor.py:136-245):
 *      wav (B, wav_cols) fp32 with row stride wav_stride; samples >= lengths[b] read as 0; always
 *      30 s => mel_cf (B,80,3000), mel_lens[b] = ceil(len/160). ---- */
int swc_mel(const swc_model* m, const float* wav, int64_t wav_stride, int wav_cols, const int64_t* lengths,
            int batch, float* mel_cf, int64_t* mel_lens, void* workspace, size_t ws_bytes, void* stream);

/* ---- OmniAudioEncoder.forward (reference audiocodec/nn/modules.py:287-376):
 *      mel_cf (B,80,Tm) -> enc_cf (B,768,ceil(Tm/2)), out_lens = mel_lens // 2 ---- */
int swc_encoder(const swc_model* m, const float* mel_cf, const int64_t* mel_lens, int batch, int mel_frames,
                float* enc_cf, int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream);
/* the same with output_hidden_states=True (modules.py:344-371): hidden_cf (n_layers + 1, B, 768, ceil(T/2)) fp32 receives the
   input of each of the 12 layers and the final LayerNorm output, frames >= length zeroed, channels-first like enc_cf */
int swc_encoder_hidden(const swc_model* m, const float* mel_cf, const int64_t* mel_lens, int batch, int mel_frames,
                       float* enc_cf, int64_t* out_lens, float* hidden_cf, void* workspace, size_t ws_bytes, void* stream);

/* ---- FrameStackDownConv.forward (modules.py:519-550): (B,768,T) -> latent (B,32,ceil(T/4)) ---- */
int swc_downsample(const swc_model* m, const float* x_cf, const int64_t* lens, int batch, int frames,
                   float* latent_cf, int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream);

/* ---- GroupFiniteScalarQuantizer.forward / .decode (reference audiocodec/nn/quantizer.py:273-317) ---- */
int swc_quantize(const swc_model* m, const float* latent_cf, const int64_t* lens, int batch, int frames,
                 float* zq_cf, int32_t* codes /* (8,B,T) */, void* stream);
int swc_dequantize(const swc_model* m, const void* codes, int codes_are_int64, const int64_t* lens, int batch,
                   int frames, float* zq_cf, void* stream);

/* ---- FrameStackUpConv.forward (modules.py:601-631): (B,32,T') -> (B,768,4T') ---- */
int swc_upsample(const swc_model* m, const float* zq_cf, const int64_t* lens, int batch, int frames,
                 float* y_cf, int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream);

/* ---- OmniAudioDecoder.forward (modules.py:437-474): (B,768,T) -> (B,80,2T) ---- */
int swc_decoder(const swc_model* m, const float* x_cf, const int64_t* lens, int batch, int frames,
                float* mel_cf, int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream);

/* ---- Vocos.forward (modules.py:1569-1573, 1492-1504, 1064-1082, 831-886): (B,80,Tv) -> wav (B,160 Tv) ---- */
int swc_vocos(const swc_model* m, const float* mel_cf, const int64_t* lens, int batch, int frames,
              float* wav, int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream);

/* ---- AudioCodec.inference_tokenize (reference audiocodec/model.py:167-210), fused on device:
 *      wav -> codes (8,B,375) int32, optional zq_cf (B,32,375), codes_lens[b]. ---- */
int swc_tokenize(const swc_model* m, const float* wav, int64_t wav_stride, int wav_cols, const int64_t* lengths,
                 int batch, int32_t* codes, float* zq_cf, int64_t* codes_lens, void* workspace,
                 size_t ws_bytes, void* stream);

/* ---- AudioCodec.inference_detokenize (model.py:212-242): codes (8,B,T') -> wav (B, 1280 T').
 *      T' is explicit: the un-masked convolutions see the zero padding up to T' (SURVEY 3.3). ---- */
int swc_detokenize(const swc_model* m, const void* codes, int codes_are_int64, const int64_t* lens, int batch,
                   int code_frames, float* wav, int64_t* out_lens, void* workspace, size_t ws_bytes,
                   void* stream);

/* ---- the same two chains when the caller also knows the lengths on the HOST (AudioCodec.encode()/decode() do:
 *      model.py:262-268, 327-333 build them from Python lists).  In bf16 mode the two transformer stacks then run on the
 *      packed valid tokens only (padded tokens of a window are skipped), and swc_detokenize_ragged runs Vocos only over
 *      each item's valid frames plus its receptive-field halo, all items packed into one batch of rows.  Codes, code lengths
 *      and the first out_lens[b] = 1280 lens[b] samples of every item are bit-identical to swc_tokenize / swc_detokenize;
 *      samples of wav beyond an item's out_lens[b] are NOT written by the ragged form (the padded form computes them from
 *      the padding, as the reference does; its callers discard them: model.py:352-356).  host_lengths[b] must equal
 *      lengths[b]; a call takes at most swc_max_ragged() items (the length tables travel as kernel parameters): larger
 *      batches are an error, the caller splits them (AudioCodec does). ---- */
int swc_max_ragged(void);
int swc_tokenize_ragged(const swc_model* m, const float* wav, int64_t wav_stride, int wav_cols, const int64_t* lengths,
                        const int64_t* host_lengths, int batch, int32_t* codes, float* zq_cf, int64_t* codes_lens,
                        void* workspace, size_t ws_bytes, void* stream);
int swc_detokenize_ragged(const swc_model* m, const void* codes, int codes_are_int64, const int64_t* lens,
                          const int64_t* host_lens, int batch, int code_frames, float* wav, int64_t* out_lens,
                          void* workspace, size_t ws_bytes, void* stream);

/* ---- AudioCodec.forward (model.py:112-165): mel_cf (B,80,Tm) -> audio (B, 160*8*ceil(ceil(Tm/2)/4)) ---- */
int swc_forward(const swc_model* m, const float* mel_cf, const int64_t* mel_lens, int batch, int mel_frames,
                float* wav, int64_t* out_lens, int32_t* codes /* optional (8,B,Tc) */, void* workspace,
                size_t ws_bytes, void* stream);

/* ---- launch accounting: every kernel launch is counted per class (0 tcgen05 GEMM, 1 SIMT GEMM,
 *      2 attention, 3 LayerNorm, 4 dwconv+LayerNorm, 5 anti-aliased snake, 6 other); with
 *      enable_timing != 0 each launch is also bracketed by CUDA events on its stream.
 *      swc_profile() resets; swc_profile_read() synchronises on the recorded events and returns
 *      per-class milliseconds and launch counts (arrays of >= 7 entries). ---- */
void swc_profile(int enable_timing);
int swc_profile_read(double* ms_per_class, int64_t* launches_per_class, int n);

/* ---- low-level operator hooks used by the unit tests (GEMM back ends, attention) ---- */
int swc_test_gemm(int backend /*0 simt fp32, 1 simt bf16, 2 tcgen05 bf16 (gen 1), 3 gen 2 single-CTA, 4 gen 2 CTA pairs*/, const void* A, const void* W,
                  const float* bias, void* out, int out_bf16, int M, int N, int K, int act, void* stream);
/* tuning hook: which tcgen05 GEMM generation the pipeline uses (0 gen 1, 1 gen 2 single-CTA, 2 gen 2 CTA pairs = default) */
void swc_set_gemm_variant(int variant);
int swc_test_attention(int backend /*0 simt fp32, 1 simt bf16, 2 mma.sync bf16, 3 tcgen05 bf16*/, const void* qkv, void* out,
                       const int64_t* lens, int batch, int T, int heads, void* stream);
/* profiling hook (tools/attn_trace.py): enable != 0 arms per-phase clock64 stamps of CTA 0 for the next tcgen05 attention
   launch; enable == 0 synchronises, copies up to n stamps (3 x 4096 slots) to host_out and disarms.  0 on success. */
int swc_debug_attn_trace(int enable, long long* host_out, int n);

#ifdef __cplusplus
}
#endif
#endif /* SWC_H_ */
