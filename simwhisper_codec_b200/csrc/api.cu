// extern "C" surface of libswc.so (include/swc.h).  Module-level entry points take the reference's
// channels-first fp32 tensors, convert to the channel-last activation layout, and run the stage.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../include/swc.h"
#include "pipeline.h"

namespace swc {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

struct ProfState {
  bool timing = false;
  long long launches[KC_COUNT] = {0};
  struct Rec { int cls; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t cur = nullptr;
};
// per host thread: each thread that drives a stream sees its own launch counters and timers (no shared mutable state)
static thread_local ProfState g_prof;
static cudaEvent_t prof_event() {
  if (!g_prof.pool.empty()) { cudaEvent_t e = g_prof.pool.back(); g_prof.pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
void prof_begin(int cls, cudaStream_t s) {
  g_prof.launches[cls]++;
  if (!g_prof.timing) return;
  g_prof.cur = prof_event();
  cudaEventRecord(g_prof.cur, s);
}
void prof_end(int cls, cudaStream_t s) {
  if (!g_prof.timing || !g_prof.cur) return;
  cudaEvent_t b = prof_event();
  cudaEventRecord(b, s);
  g_prof.recs.push_back({cls, g_prof.cur, b});
  g_prof.cur = nullptr;
}
}  // namespace swc

using namespace swc;

namespace {

size_t esz(int t) { return t == 0 ? 4 : 2; }

int make_ctx(const swc_model* mm, void* ws, size_t ws_bytes, void* stream, Ctx& c) {
  SWC_REQUIRE(mm != nullptr, "null model");
  SWC_REQUIRE(mm->m.uploaded, "model is not finalized on a CUDA device (swc_model_finalize); there is no CPU path");
  c.m = &mm->m;
  c.s = (cudaStream_t)stream;
  c.ws.base = (char*)ws;
  c.ws.cap = ws_bytes;
  c.dry = false;
  const char* f = getenv("SWC_FORCE_SIMT");
  c.force_simt = f && f[0] == '1';
  SWC_REQUIRE(((uintptr_t)ws & 255) == 0, "workspace must be 256-byte aligned");
  return 0;
}

// a device pointer handed to an entry point must live on the model's device (the kernels and tensor maps are issued there)
int check_on_device(int device, const void* p, const char* what) {
  if (!p) return 0;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return 0; }
  SWC_REQUIRE(a.type != cudaMemoryTypeDevice || a.device == device, "%s lives on device %d but the model was finalized on device %d",
              what, a.device, device);
  return 0;
}

// run `fn` against either the real arena or a dry arena (workspace sizing)
template <typename Fn>
size_t dry_size(const swc_model* mm, Fn fn) {
  Ctx c;
  c.m = &mm->m;
  c.dry = true;
  fn(c);
  // bf16x3 mode: every tensor-core GEMM splits its fp32 operand into bf16 planes of the same byte size on top of whatever is
  // live at that point; the operand is itself one allocation of this pass, so the largest one bounds the extra space
  return c.ws.peak + 256 + (mm->m.x3() ? c.ws.largest + 512 : 0);
}

// ---- stage bodies shared by the real call and the dry sizing -------------------------------------
int stage_mel(Ctx& c, const float* wav, long long stride, int cols, const long long* lens, int B, float* mel_cf, long long* mel_lens) {
  return mel_frontend(c, wav, stride, cols, lens, B, mel_cf, nullptr, mel_lens);
}

int stage_encoder(Ctx& c, const float* mel_cf, const long long* mel_lens, int B, int Tm, float* enc_cf, long long* out_lens) {
  const Model& m = *c.m;
  const int at = m.act_type(), T = (Tm + 1) / 2, T4 = (T + 3) / 4 * 4;
  void* mel_cl = c.ws.alloc((long long)B * Tm * m.mel_pitch * esz(at));
  void* enc_cl = c.ws.alloc((long long)B * T4 * m.d_model * esz(at));
  long long* enc_lens = (long long*)c.ws.alloc(B * 8);
  SWC_TRY(c.ws.check());
  if (!c.dry) SWC_TRY(cf_to_cl(mel_cf, mel_cl, at, B, m.mel_bins, Tm, Tm, m.mel_pitch, c.s));
  SWC_TRY(lens_affine_pub(c, mel_lens, enc_lens, B, 1, 0, 2));
  SWC_TRY(lens_affine_pub(c, enc_lens, out_lens, B, 1, 0, 1));
  SWC_TRY(encoder_cl(c, mel_cl, enc_lens, B, Tm, enc_cl));
  if (!c.dry) SWC_TRY(cl_to_cf(enc_cl, at, enc_cf, B, m.d_model, T, (long long)T4 * m.d_model, m.d_model, c.s));
  return 0;
}

int stage_downsample(Ctx& c, const float* x_cf, const long long* lens, int B, int T, float* latent_cf, long long* out_lens) {
  const Model& m = *c.m;
  const int at = m.act_type(), T4 = (T + 3) / 4 * 4;
  void* x_cl = c.ws.alloc((long long)B * T4 * m.d_model * esz(at));
  long long* code_lens = (long long*)c.ws.alloc(B * 8);
  SWC_TRY(c.ws.check());
  if (!c.dry) SWC_TRY(cf_to_cl(x_cf, x_cl, at, B, m.d_model, T, T4, m.d_model, c.s));
  SWC_TRY(lens_affine_pub(c, lens, code_lens, B, 1, 3, 4));
  SWC_TRY(lens_affine_pub(c, code_lens, out_lens, B, 1, 0, 1));
  return downsample_fsq(c, x_cl, code_lens, B, T4, nullptr, nullptr, latent_cf, nullptr);
}

int stage_upsample(Ctx& c, const float* zq_cf, const long long* lens, int B, int Tc, float* y_cf, long long* out_lens) {
  const Model& m = *c.m;
  float* zq_cl = (float*)c.ws.alloc((long long)B * Tc * m.latent * 4);
  float* h = (float*)c.ws.alloc((long long)B * 4 * Tc * m.d_model * 4);
  SWC_TRY(c.ws.check());
  if (!c.dry) SWC_TRY(cf_to_cl(zq_cf, zq_cl, 0, B, m.latent, Tc, Tc, m.latent, c.s));
  SWC_TRY(lens_affine_pub(c, lens, out_lens, B, 4, 0, 1));
  SWC_TRY(upsample_cl(c, zq_cl, B, Tc, h));
  if (!c.dry) SWC_TRY(cl_to_cf(h, 0, y_cf, B, m.d_model, 4 * Tc, (long long)4 * Tc * m.d_model, m.d_model, c.s));
  return 0;
}

int stage_decoder(Ctx& c, const float* x_cf, const long long* lens, int B, int T, float* mel_cf, long long* out_lens) {
  const Model& m = *c.m;
  const int at = m.act_type();
  float* h = (float*)c.ws.alloc((long long)B * T * m.d_model * 4);
  void* mel_cl = c.ws.alloc((long long)B * 2 * T * m.mel_pitch * esz(at));
  SWC_TRY(c.ws.check());
  if (!c.dry) SWC_TRY(cf_to_cl(x_cf, h, 0, B, m.d_model, T, T, m.d_model, c.s));
  SWC_TRY(lens_affine_pub(c, lens, out_lens, B, 2, 0, 1));
  SWC_TRY(decoder_cl(c, h, lens, B, T, mel_cl));
  if (!c.dry) SWC_TRY(cl_to_cf(mel_cl, at, mel_cf, B, m.mel_bins, 2 * T, (long long)2 * T * m.mel_pitch, m.mel_pitch, c.s));
  return 0;
}

int stage_vocos(Ctx& c, const float* mel_cf, const long long* lens, int B, int Tv, float* wav, long long* out_lens) {
  const Model& m = *c.m;
  const int at = m.act_type();
  void* mel_cl = c.ws.alloc((long long)B * Tv * m.mel_pitch * esz(at));
  SWC_TRY(c.ws.check());
  if (!c.dry) SWC_TRY(cf_to_cl(mel_cf, mel_cl, at, B, m.mel_bins, Tv, Tv, m.mel_pitch, c.s));
  SWC_TRY(lens_affine_pub(c, lens, out_lens, B, 160, 0, 1));
  return vocos_cl(c, mel_cl, B, Tv, wav);
}

int stage_detokenize(Ctx& c, const void* codes, int i64, const long long* lens, int B, int Tc, float* wav, long long* out_lens) {
  const Model& m = *c.m;
  float* zq_cl = (float*)c.ws.alloc((long long)B * Tc * m.latent * 4);
  SWC_TRY(c.ws.check());
  if (!c.dry) SWC_TRY(fsq_decode(codes, i64, lens, B, Tc, m.fsq, nullptr, zq_cl, c.s));
  return detokenize_chain(c, zq_cl, lens, B, Tc, wav, out_lens);
}

int stage_forward(Ctx& c, const float* mel_cf, const long long* mel_lens, int B, int Tm, float* wav, long long* out_lens, int* codes) {
  const Model& m = *c.m;
  const int at = m.act_type(), T = (Tm + 1) / 2, T4 = (T + 3) / 4 * 4, Tc = T4 / 4;
  void* mel_cl = c.ws.alloc((long long)B * Tm * m.mel_pitch * esz(at));
  void* enc_cl = c.ws.alloc((long long)B * T4 * m.d_model * esz(at));
  float* zq_cl = (float*)c.ws.alloc((long long)B * Tc * m.latent * 4);
  long long* enc_lens = (long long*)c.ws.alloc(B * 8);
  long long* code_lens = (long long*)c.ws.alloc(B * 8);
  SWC_TRY(c.ws.check());
  if (!c.dry) SWC_TRY(cf_to_cl(mel_cf, mel_cl, at, B, m.mel_bins, Tm, Tm, m.mel_pitch, c.s));
  SWC_TRY(lens_affine_pub(c, mel_lens, enc_lens, B, 1, 0, 2));
  SWC_TRY(lens_affine_pub(c, enc_lens, code_lens, B, 1, 3, 4));
  SWC_TRY(encoder_cl(c, mel_cl, enc_lens, B, Tm, enc_cl));
  SWC_TRY(downsample_fsq(c, enc_cl, code_lens, B, T4, codes, nullptr, nullptr, zq_cl));
  return detokenize_chain(c, zq_cl, code_lens, B, Tc, wav, out_lens);
}

}  // namespace

extern "C" {

int swc_version(void) { return 100; }
const char* swc_last_error(void) { return swc::g_err; }

int swc_model_create(swc_model** out, int precision) {
  SWC_REQUIRE(out != nullptr, "null output pointer");
  SWC_REQUIRE(precision == SWC_PRECISION_FP32 || precision == SWC_PRECISION_BF16 || precision == SWC_PRECISION_BF16X3, "unknown precision %d", precision);
  *out = new swc_model();
  (*out)->m.precision = precision;
  return 0;
}

int swc_model_set_tensor(swc_model* mm, const char* key, const void* host_data, int dtype, const int64_t* shape, int ndim) {
  SWC_REQUIRE(mm && key && host_data && shape, "null argument");
  SWC_REQUIRE(!mm->m.packed, "model already packed");
  RawTensor t;
  t.dtype = dtype;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  if (dtype == SWC_DTYPE_F32) { t.f.resize(n); std::memcpy(t.f.data(), host_data, n * 4); }
  else if (dtype == SWC_DTYPE_I32) { t.i.resize(n); std::memcpy(t.i.data(), host_data, n * 4); }
  else { set_error("unsupported dtype %d for %s", dtype, key); return -1; }
  mm->m.raw[key] = std::move(t);
  return 0;
}

int swc_model_pack(swc_model* mm) {
  SWC_REQUIRE(mm != nullptr, "null model");
  if (mm->m.packed) return 0;
  return pack_model(mm->m);
}

int64_t swc_model_packed_numel(const swc_model* mm, const char* name) {
  if (!mm || !mm->m.packed) return -1;
  auto it = mm->m.tab.find(name);
  if (it == mm->m.tab.end()) return -1;
  return (int64_t)it->second.host.size();
}

int swc_model_get_packed(const swc_model* mm, const char* name, float* host_out, int64_t numel) {
  SWC_REQUIRE(mm && mm->m.packed && !mm->m.uploaded, "packed tables are only readable between pack and finalize");
  auto it = mm->m.tab.find(name);
  SWC_REQUIRE(it != mm->m.tab.end(), "no packed table named %s", name);
  SWC_REQUIRE((int64_t)it->second.host.size() == numel, "size mismatch for %s", name);
  std::memcpy(host_out, it->second.host.data(), numel * 4);
  return 0;
}

int swc_model_finalize(swc_model* mm, int device) {
  SWC_REQUIRE(mm != nullptr, "null model");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  SWC_REQUIRE(e == cudaSuccess && n > 0, "no CUDA device available: %s (libswc has no CPU fallback)", cudaGetErrorString(e));
  SWC_TRY(swc_model_pack(mm));
  return upload_model(mm->m, device);
}

void swc_model_destroy(swc_model* mm) {
  if (!mm) return;
  free_model(mm->m);
  delete mm;
}

size_t swc_workspace_bytes(const swc_model* mm, int stage, int B, int frames) {
  if (!mm) return 0;
  switch (stage) {
    case SWC_STAGE_MEL: return dry_size(mm, [&](Ctx& c) { stage_mel(c, nullptr, 0, 0, nullptr, B, nullptr, nullptr); });
    case SWC_STAGE_ENCODER: return dry_size(mm, [&](Ctx& c) { stage_encoder(c, nullptr, nullptr, B, frames, nullptr, nullptr); });
    case SWC_STAGE_DOWNSAMPLE: return dry_size(mm, [&](Ctx& c) { stage_downsample(c, nullptr, nullptr, B, frames, nullptr, nullptr); });
    case SWC_STAGE_QUANTIZER: return 256;
    case SWC_STAGE_UPSAMPLE: return dry_size(mm, [&](Ctx& c) { stage_upsample(c, nullptr, nullptr, B, frames, nullptr, nullptr); });
    case SWC_STAGE_DECODER: return dry_size(mm, [&](Ctx& c) { stage_decoder(c, nullptr, nullptr, B, frames, nullptr, nullptr); });
    case SWC_STAGE_VOCOS: return dry_size(mm, [&](Ctx& c) { stage_vocos(c, nullptr, nullptr, B, frames, nullptr, nullptr); });
    case SWC_STAGE_TOKENIZE: return dry_size(mm, [&](Ctx& c) { tokenize_chain(c, nullptr, 0, 0, nullptr, B, nullptr, nullptr, nullptr); });
    case SWC_STAGE_DETOKENIZE: return dry_size(mm, [&](Ctx& c) { stage_detokenize(c, nullptr, 0, nullptr, B, frames, nullptr, nullptr); });
    case SWC_STAGE_FORWARD: return dry_size(mm, [&](Ctx& c) { stage_forward(c, nullptr, nullptr, B, frames, nullptr, nullptr, nullptr); });
  }
  return 0;
}

#define SWC_ENTER()                                  \
  Ctx c;                                             \
  SWC_TRY(make_ctx(m, workspace, ws_bytes, stream, c)); \
  SWC_REQUIRE(batch > 0, "empty batch");             \
  DeviceGuard guard(m->m.device);                    \
  SWC_REQUIRE(guard.ok(), "cannot make device %d current", m->m.device); \
  SWC_TRY(check_on_device(m->m.device, workspace, "workspace"))

int swc_mel(const swc_model* m, const float* wav, int64_t wav_stride, int wav_cols, const int64_t* lengths, int batch,
            float* mel_cf, int64_t* mel_lens, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, wav, "wav"));
  return stage_mel(c, wav, wav_stride, wav_cols, (const long long*)lengths, batch, mel_cf, (long long*)mel_lens);
}

int swc_encoder(const swc_model* m, const float* mel_cf, const int64_t* mel_lens, int batch, int mel_frames, float* enc_cf,
                int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, mel_cf, "mel_cf"));
  SWC_REQUIRE(mel_frames > 0, "encoder: no frames");
  return stage_encoder(c, mel_cf, (const long long*)mel_lens, batch, mel_frames, enc_cf, (long long*)out_lens);
}

int swc_encoder_hidden(const swc_model* m, const float* mel_cf, const int64_t* mel_lens, int batch, int mel_frames, float* enc_cf,
                       int64_t* out_lens, float* hidden_cf, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, mel_cf, "mel_cf"));
  SWC_REQUIRE(mel_frames > 0 && hidden_cf != nullptr, "encoder_hidden: no frames or no output buffer");
  c.hidden_out = hidden_cf;
  return stage_encoder(c, mel_cf, (const long long*)mel_lens, batch, mel_frames, enc_cf, (long long*)out_lens);
}

int swc_downsample(const swc_model* m, const float* x_cf, const int64_t* lens, int batch, int frames, float* latent_cf,
                   int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, x_cf, "x_cf"));
  SWC_REQUIRE(frames > 0, "downsample: no frames");
  return stage_downsample(c, x_cf, (const long long*)lens, batch, frames, latent_cf, (long long*)out_lens);
}

int swc_quantize(const swc_model* m, const float* latent_cf, const int64_t* lens, int batch, int frames, float* zq_cf,
                 int32_t* codes, void* stream) {
  SWC_REQUIRE(m && m->m.uploaded, "model is not finalized on a CUDA device");
  DeviceGuard guard(m->m.device);
  SWC_REQUIRE(guard.ok(), "cannot make device %d current", m->m.device);
  SWC_TRY(check_on_device(m->m.device, latent_cf, "latent"));
  return fsq_encode_cf(latent_cf, (const long long*)lens, batch, frames, m->m.fsq, zq_cf, codes, nullptr, (cudaStream_t)stream);
}

int swc_dequantize(const swc_model* m, const void* codes, int codes_are_int64, const int64_t* lens, int batch, int frames,
                   float* zq_cf, void* stream) {
  SWC_REQUIRE(m && m->m.uploaded, "model is not finalized on a CUDA device");
  DeviceGuard guard(m->m.device);
  SWC_REQUIRE(guard.ok(), "cannot make device %d current", m->m.device);
  SWC_TRY(check_on_device(m->m.device, codes, "codes"));
  return fsq_decode(codes, codes_are_int64, (const long long*)lens, batch, frames, m->m.fsq, zq_cf, nullptr, (cudaStream_t)stream);
}

int swc_upsample(const swc_model* m, const float* zq_cf, const int64_t* lens, int batch, int frames, float* y_cf,
                 int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, zq_cf, "zq_cf"));
  return stage_upsample(c, zq_cf, (const long long*)lens, batch, frames, y_cf, (long long*)out_lens);
}

int swc_decoder(const swc_model* m, const float* x_cf, const int64_t* lens, int batch, int frames, float* mel_cf,
                int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, x_cf, "x_cf"));
  return stage_decoder(c, x_cf, (const long long*)lens, batch, frames, mel_cf, (long long*)out_lens);
}

int swc_vocos(const swc_model* m, const float* mel_cf, const int64_t* lens, int batch, int frames, float* wav,
              int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, mel_cf, "mel_cf"));
  return stage_vocos(c, mel_cf, (const long long*)lens, batch, frames, wav, (long long*)out_lens);
}

int swc_tokenize(const swc_model* m, const float* wav, int64_t wav_stride, int wav_cols, const int64_t* lengths, int batch,
                 int32_t* codes, float* zq_cf, int64_t* codes_lens, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, wav, "wav"));
  return tokenize_chain(c, wav, wav_stride, wav_cols, (const long long*)lengths, batch, codes, zq_cf, (long long*)codes_lens);
}

int swc_detokenize(const swc_model* m, const void* codes, int codes_are_int64, const int64_t* lens, int batch, int code_frames,
                   float* wav, int64_t* out_lens, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, codes, "codes"));
  SWC_REQUIRE(code_frames > 0, "detokenize: no code frames");
  return stage_detokenize(c, codes, codes_are_int64, (const long long*)lens, batch, code_frames, wav, (long long*)out_lens);
}

// host-known lengths -> packed-token table (bf16 mode).  Returns false if the batch does not qualify (the callers then run
// the padded path, which gives identical results).
static bool build_ragged(RaggedTable& tab, const int64_t* host_lens, int batch, int mode /*0 samples -> tokens, 1 code frames -> tokens*/,
                         int limit) {
  if (!host_lens || batch <= 0) return false;
  tab.nb = batch; tab.total = 0; tab.t_max = 0;
  for (int b = 0; b < batch; ++b) {
    long long l = host_lens[b] < 0 ? 0 : host_lens[b];
    int tok;
    if (mode == 0) {
      if (l > limit) l = limit;                 // samples, clamped like mel_pad does
      tok = (int)(((l + 159) / 160) / 2);      // mel_len // 2 (reference modules.py:322)
    } else {
      if (l > limit) l = limit;                 // code frames, clamped to the decode pad length T'
      tok = (int)(4 * l);                       // up-sampler: len * 4 (reference modules.py:626)
    }
    tab.len[b] = tok;
    tab.off[b] = tab.total;
    tab.total += tok;
    tab.t_max = tok > tab.t_max ? tok : tab.t_max;
  }
  tab.off[batch] = tab.total;
  return tab.total > 0;
}

int swc_max_ragged(void) { return kMaxRagged; }

int swc_tokenize_ragged(const swc_model* m, const float* wav, int64_t wav_stride, int wav_cols, const int64_t* lengths,
                        const int64_t* host_lengths, int batch, int32_t* codes, float* zq_cf, int64_t* codes_lens,
                        void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, wav, "wav"));
  SWC_REQUIRE(batch <= kMaxRagged, "swc_tokenize_ragged: %d items exceed swc_max_ragged() = %d; split the batch", batch, kMaxRagged);
  RaggedTable tab;
  if ((m->m.act_type() == 1 || m->m.x3()) && build_ragged(tab, host_lengths, batch, 0, wav_cols < 480000 ? wav_cols : 480000)) c.rag = &tab;
  return tokenize_chain(c, wav, wav_stride, wav_cols, (const long long*)lengths, batch, codes, zq_cf, (long long*)codes_lens);
}

int swc_detokenize_ragged(const swc_model* m, const void* codes, int codes_are_int64, const int64_t* lens,
                          const int64_t* host_lens, int batch, int code_frames, float* wav, int64_t* out_lens, void* workspace,
                          size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, codes, "codes"));
  SWC_REQUIRE(code_frames > 0, "detokenize: no code frames");
  SWC_REQUIRE(batch <= kMaxRagged, "swc_detokenize_ragged: %d items exceed swc_max_ragged() = %d; split the batch", batch, kMaxRagged);
  RaggedTable tab;
  if ((m->m.act_type() == 1 || m->m.x3()) && build_ragged(tab, host_lens, batch, 1, code_frames)) c.rag = &tab;
  return stage_detokenize(c, codes, codes_are_int64, (const long long*)lens, batch, code_frames, wav, (long long*)out_lens);
}

int swc_forward(const swc_model* m, const float* mel_cf, const int64_t* mel_lens, int batch, int mel_frames, float* wav,
                int64_t* out_lens, int32_t* codes, void* workspace, size_t ws_bytes, void* stream) {
  SWC_ENTER();
  SWC_TRY(check_on_device(m->m.device, mel_cf, "mel_cf"));
  return stage_forward(c, mel_cf, (const long long*)mel_lens, batch, mel_frames, wav, (long long*)out_lens, codes);
}

void swc_profile(int enable_timing) {
  g_prof.timing = enable_timing != 0;
  for (int i = 0; i < KC_COUNT; ++i) g_prof.launches[i] = 0;
  for (auto& r : g_prof.recs) { g_prof.pool.push_back(r.a); g_prof.pool.push_back(r.b); }
  g_prof.recs.clear();
}

int swc_profile_read(double* ms_per_class, int64_t* launches_per_class, int n) {
  SWC_REQUIRE(n >= KC_COUNT, "swc_profile_read: need room for %d classes", (int)KC_COUNT);
  for (int i = 0; i < n; ++i) { ms_per_class[i] = 0.0; launches_per_class[i] = 0; }
  for (auto& r : g_prof.recs) {
    SWC_CHECK_CUDA(cudaEventSynchronize(r.b));
    float ms = 0.f;
    SWC_CHECK_CUDA(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_per_class[r.cls] += ms;
    g_prof.pool.push_back(r.a);
    g_prof.pool.push_back(r.b);
  }
  g_prof.recs.clear();
  for (int i = 0; i < KC_COUNT; ++i) launches_per_class[i] = g_prof.launches[i];
  return 0;
}

int swc_test_gemm(int backend, const void* A, const void* W, const float* bias, void* out, int out_bf16, int M, int N, int K,
                  int act, void* stream) {
  GemmDesc d{};
  d.A = A; d.a_row_stride = K; d.a_batch_stride = 0; d.a_rows = M; d.a_cols = K; d.m_rows = M; d.nb = 1;
  d.n_taps = 1; d.tap_k = K;
  d.W = W; d.N = N; d.w_rows = N;
  d.epi.bias = bias; d.epi.act = act; d.epi.out = out; d.epi.out_row_stride = N; d.epi.out_row_mul = 1; d.epi.nb = 1;
  if (backend == 0) return gemm_simt(d, EPI_STORE, 0, out_bf16, (cudaStream_t)stream);
  if (backend == 1) return gemm_simt(d, EPI_STORE, 1, out_bf16, (cudaStream_t)stream);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // 2: first-generation kernel, 3: TMA-store epilogue, 4: CTA pairs + TMA-store epilogue
  const int saved = get_gemm_variant();
  set_gemm_variant(backend - 2);
  const int rc = gemm_tc(d, EPI_STORE, out_bf16, sms, (cudaStream_t)stream);
  set_gemm_variant(saved);
  return rc;
}

void swc_set_gemm_variant(int variant) { set_gemm_variant(variant); }

int swc_test_attention(int backend, const void* qkv, void* out, const int64_t* lens, int batch, int T, int heads, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  if (backend == 0) return attention_simt(qkv, 0, out, (const long long*)lens, batch, T, heads, s);
  if (backend == 1) return attention_simt(qkv, 1, out, (const long long*)lens, batch, T, heads, s);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (backend == 2) {
    // bf16x3 tcgen05 kernel behind an fp32 interface: split qkv into (hi | lo) planes, run, merge the output planes
    const long long rows = (long long)batch * T;
    const int D = heads * 64;
    bf16 *pin = nullptr, *pout = nullptr;
    SWC_CHECK_CUDA(cudaMalloc(&pin, (size_t)rows * 6 * D * sizeof(bf16)));
    SWC_CHECK_CUDA(cudaMalloc(&pout, (size_t)rows * 2 * D * sizeof(bf16)));
    int rc = split_bf16_planes((const float*)qkv, 3 * D, (long long)T * 3 * D, batch, T, 3 * D, pin, s);
    if (rc == 0) rc = attention_tc_x3(pin, pout, (const long long*)lens, batch, T, heads, sms, s);
    if (rc == 0) rc = merge_bf16_planes(pout, rows, D, (float*)out, s);
    cudaStreamSynchronize(s);
    cudaFree(pin);
    cudaFree(pout);
    return rc;
  }
  return attention_tc((const bf16*)qkv, (bf16*)out, (const long long*)lens, batch, T, heads, sms, s);
}

}  // extern "C"
