// Stage orchestration: which kernels run, on which channel-last buffers, for each reference module.
// Everything is enqueued on the caller's stream; scratch comes from a bump arena over the caller's
// workspace.  The same code runs in "dry" mode (no launches) to size the workspace.
#include "pipeline.h"

#include <algorithm>
#include <cstdlib>

namespace swc {

namespace {

__global__ void lens_affine_kernel(const long long* in, long long* out, int n, long long mul, long long add, long long div) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    long long v = in[i];
    if (v < 0) v = 0;
    out[i] = (v * mul + add) / div;
  }
}

int lens_affine(Ctx& c, const long long* in, long long* out, int n, long long mul, long long add, long long div) {
  if (c.dry || !out) return 0;
  ProfScope ps(KC_MISC, c.s);
  lens_affine_kernel<<<ceil_div(n, 128), 128, 0, c.s>>>(in, out, n, mul, add, div);
  SWC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

size_t esz(int type) { return type == 0 ? 4 : 2; }

GemmDesc base_desc(const void* A, long long a_row_stride, long long a_batch_stride, int a_rows, int a_cols,
                   int m_rows, int nb, const LinearW& w) {
  GemmDesc d{};
  d.A = A; d.a_row_stride = a_row_stride; d.a_batch_stride = a_batch_stride; d.a_rows = a_rows; d.a_cols = a_cols;
  d.m_rows = m_rows; d.nb = nb;
  d.n_taps = 1; d.tap_row[0] = 0; d.tap_col[0] = 0; d.tap_k = w.K;
  d.W = w.w; d.W3 = w.w3; d.N = w.N; d.w_rows = w.w_rows;
  d.epi.bias = w.bias;
  d.epi.out_row_mul = 1; d.epi.out_row_off = 0;
  d.epi.nb = nb;
  return d;
}

void set_out(GemmDesc& d, void* out, long long row_stride, long long batch_stride) {
  d.epi.out = out; d.epi.out_row_stride = row_stride; d.epi.out_batch_stride = batch_stride;
}

// bf16x3 mode: may the producer GEMM `a` write its result as bf16 planes for the consumer GEMM `b`?  Both must take the
// three-product path and the producer must run on the pair kernel (the only one with the planes epilogue).
bool x3_planes(const Ctx& c, const LinearW& a, const LinearW& b) {
  return c.m->x3() && !c.force_simt && !c.dry && a.w3 != nullptr && b.w3 != nullptr && get_gemm_variant() == 2 && a.N % 64 == 0 &&
         a.N >= 256 && a.K % 64 == 0 && b.K == a.N;
}

// bf16x3 mode: may a row-wise producer (LayerNorm, dwconv + LayerNorm) write its result directly as the (hi | lo) planes
// the three-product GEMM `w` reads?  (Same arithmetic as the split pass it replaces: bit-identical results.)
bool x3_rowwise_planes(const Ctx& c, const LinearW& w) {
  return c.m->x3() && !c.force_simt && w.w3 != nullptr && w.K % 64 == 0 && w.N % 8 == 0;
}

// dispatch on the model precision: bf16 operands go to the tcgen05 kernel, fp32 to the SIMT kernel.
// bf16x3 mode (fp32 activations): the fp32 operand is split into two bf16 planes (hi | lo) and the contraction runs on the
// tensor cores as three products a_hi w_hi + a_hi w_lo + a_lo w_hi with fp32 accumulation: every tap of the problem
// becomes three taps (the A plane is a column offset, the packed weight holds the planes hi | lo | hi along K).
int run_gemm(Ctx& c, const GemmDesc& d, int kind, int a_type, int out_type) {
  const bool x3 = c.m->x3() && a_type == 0 && !c.force_simt && d.W3 != nullptr && d.tap_k % 64 == 0 && d.a_cols % 8 == 0 &&
                  3 * d.n_taps <= kMaxTaps && d.N % 8 == 0 && d.a_row_stride % 4 == 0 && d.a_batch_stride % 4 == 0;
  if (!c.dry) SWC_REQUIRE(x3 || (!d.a_planes && !d.epi.out_planes), "run_gemm: plane operands need the three-product path");
  if (x3) {
    const size_t mark = c.ws.mark();
    bf16* planes = d.a_planes ? (bf16*)d.A : (bf16*)c.ws.alloc((long long)d.nb * d.a_rows * d.a_cols * 4);
    SWC_TRY(c.ws.check());
    int rc = 0;
    if (!c.dry) {
      if (!d.a_planes)
        rc = split_bf16_planes((const float*)d.A, d.a_row_stride, d.a_batch_stride, d.nb, d.a_rows, d.a_cols, planes, c.s);
      if (rc == 0) {
        GemmDesc e = d;
        e.A = planes; e.a_cols = 2 * d.a_cols; e.a_planes = 0;
        if (!d.a_planes) { e.a_row_stride = 2ll * d.a_cols; e.a_batch_stride = 2ll * d.a_cols * d.a_rows; }
        e.W = d.W3; e.W3 = nullptr;
        e.n_taps = 3 * d.n_taps;
        for (int j = 0; j < 3; ++j)
          for (int t = 0; t < d.n_taps; ++t) {
            e.tap_row[j * d.n_taps + t] = d.tap_row[t];
            e.tap_col[j * d.n_taps + t] = d.tap_col[t] + (j == 2 ? d.a_cols : 0);
          }
        rc = gemm_tc(e, kind, out_type, c.m->num_sms, c.s);
      }
    }
    c.ws.release(mark);
    return rc;
  }
  if (c.dry) return 0;
  if (a_type == 1 && !c.force_simt) return gemm_tc(d, kind, out_type, c.m->num_sms, c.s);
  return gemm_simt(d, kind, a_type, out_type, c.s);
}

// bf16 mode: bf16 rows in and out (tcgen05).  bf16x3 mode with the qkv planes of the pair GEMM: the three-product tcgen05
// kernel, whose output is again (hi | lo) planes (`*out_planes` = true).  Otherwise fp32 SIMT.
int run_attention(Ctx& c, const void* qkv, void* out, const long long* lens, int nb, int T, bool qkv_planes, bool* out_planes) {
  *out_planes = false;
  if (c.dry) return 0;
  const int at = c.m->act_type();
  if (at == 1 && !c.force_simt) return attention_tc((const bf16*)qkv, (bf16*)out, lens, nb, T, c.m->heads, c.m->num_sms, c.s);
  if (c.m->x3() && !c.force_simt && qkv_planes) {
    *out_planes = true;
    return attention_tc_x3((const bf16*)qkv, (bf16*)out, lens, nb, T, c.m->heads, c.m->num_sms, c.s);
  }
  return attention_simt(qkv, at, out, lens, nb, T, c.m->heads, c.s);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// transformer stack shared by encoder and decoder (reference modules.py:214-232), h fp32 (nb,T,768)
// ------------------------------------------------------------------------------------------------
// Each sub-layer's output GEMM accumulates straight into the fp32 residual stream: h += out_proj(...) / h += fc2(...)
// (bf16 mode: the epilogue's TMA store is a reduce-add performed by the L2; fp32 mode: the SIMT epilogue reads and
// writes h in place).  LayerNorm then only reads h and writes the normalised operand of the next GEMM.
// With `rag` the rows of h are the packed valid tokens of all items (rag->total rows) and attention runs on the packed
// layout; every other kernel is row-wise, so the result per token is bit-identical to the padded layout.
int transformer_stack(Ctx& c, const std::vector<LayerW>& layers, float* h, const long long* lens, int nb, int T,
                      const RaggedTable* rag) {
  const Model& m = *c.m;
  const int D = m.d_model, at = m.act_type();
  const int gelu = at == 1 ? 2 : 1;
  const long long rows = rag ? (long long)rag->total : (long long)nb * T;
  if (rag) { nb = 1; T = rag->total; }      // row-wise kernels see one long item
  const size_t mark = c.ws.mark();
  void* xn = c.ws.alloc(rows * D * esz(at));
  void* qkv = c.ws.alloc(rows * 3 * D * esz(at));
  void* ao = c.ws.alloc(rows * D * esz(at));
  void* ff = c.ws.alloc(rows * m.ffn * esz(at));
  SWC_TRY(c.ws.check());
  for (size_t li = 0; li < layers.size() && !c.dry; ++li) {
    const LayerW& L = layers[li];
    // bf16x3 mode: the qkv GEMM writes the (hi | lo) planes the three-product attention kernel reads
    const bool qkv_planes = c.m->x3() && !c.force_simt && L.qkv.w3 != nullptr && get_gemm_variant() == 2;
    if (c.hidden_out && !rag)      // the layer's input, masked and channels-first
      SWC_TRY(cl_to_cf(h, 0, c.hidden_out + (long long)li * nb * D * T, nb, D, T, (long long)T * D, D, c.s, lens));
    const bool xn1_planes = x3_rowwise_planes(c, L.qkv);
    SWC_TRY(layernorm(h, nullptr, nullptr, xn, xn1_planes ? 2 : at, L.ln1_g, L.ln1_b, 1e-5f, nb, T, T, D, nullptr, c.s));
    {
      GemmDesc d = base_desc(xn, xn1_planes ? 2 * D : D, 0, (int)rows, D, (int)rows, 1, L.qkv);
      d.a_planes = xn1_planes;
      set_out(d, qkv, qkv_planes ? 6 * D : 3 * D, 0);
      d.epi.out_planes = qkv_planes;
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, qkv_planes ? 1 : at));
    }
    bool ao_planes = false;      // bf16x3: the attention output is already the (hi | lo) operand of the out_proj GEMM
    if (rag && m.x3()) {
      SWC_REQUIRE(qkv_planes, "bf16x3: the packed-token path needs the pair kernel's planes epilogue for qkv");
      SWC_TRY(attention_tc_x3_ragged((const bf16*)qkv, (bf16*)ao, *rag, m.heads, m.num_sms, c.s));
      ao_planes = true;
    } else if (rag) {
      SWC_TRY(attention_tc_ragged((const bf16*)qkv, (bf16*)ao, *rag, m.heads, m.num_sms, c.s));
    } else {
      SWC_TRY(run_attention(c, qkv, ao, lens, nb, T, qkv_planes, &ao_planes));
    }
    {
      GemmDesc d = base_desc(ao, ao_planes ? 2 * D : D, 0, (int)rows, D, (int)rows, 1, L.out);
      d.a_planes = ao_planes;
      set_out(d, h, D, 0);
      d.epi.residual = h; d.epi.res_row_stride = D;
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, 0));
    }
    const bool xn2_planes = x3_rowwise_planes(c, L.fc1);
    SWC_TRY(layernorm(h, nullptr, nullptr, xn, xn2_planes ? 2 : at, L.ln2_g, L.ln2_b, 1e-5f, nb, T, T, D, nullptr, c.s));
    // bf16x3 mode: fc1 hands its GELU output to fc2 as the two bf16 planes fc2's three-product GEMM reads (same bytes as
    // the fp32 hidden, no separate split pass over the widest operand of the layer)
    const bool planes = x3_planes(c, L.fc1, L.fc2);
    {
      GemmDesc d = base_desc(xn, xn2_planes ? 2 * D : D, 0, (int)rows, D, (int)rows, 1, L.fc1);
      d.a_planes = xn2_planes;
      set_out(d, ff, planes ? 2 * m.ffn : m.ffn, 0);
      d.epi.act = gelu;
      d.epi.out_planes = planes;
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, planes ? 1 : at));
    }
    {
      GemmDesc d = base_desc(ff, planes ? 2 * m.ffn : m.ffn, 0, (int)rows, m.ffn, (int)rows, 1, L.fc2);
      d.a_planes = planes;
      set_out(d, h, D, 0);
      d.epi.residual = h; d.epi.res_row_stride = D;
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, 0));
    }
  }
  c.ws.release(mark);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// encoder (reference modules.py:287-376).  mel_cl (nb,Tm,128) act type -> enc_cl (nb,T4,768) act type
// with T = ceil(Tm/2) tokens, T4 = T rounded up to a multiple of 4 (zero rows), rows >= lens zeroed.
// ------------------------------------------------------------------------------------------------
int encoder_cl(Ctx& c, const void* mel_cl, const long long* enc_lens, int nb, int Tm, void* enc_cl) {
  const Model& m = *c.m;
  const int D = m.d_model, MP = m.mel_pitch, at = m.act_type();
  const int T = (Tm + 1) / 2, T4 = (T + 3) / 4 * 4;
  const size_t mark = c.ws.mark();
  void* stem = c.ws.alloc((long long)nb * 2 * T * D * esz(at));
  float* h = (float*)c.ws.alloc((long long)nb * T * D * 4);
  SWC_TRY(c.ws.check());
  if (!c.dry) {
    if (Tm & 1) SWC_CHECK_CUDA(cudaMemsetAsync(stem, 0, (size_t)nb * 2 * T * D * esz(at), c.s));
    {   // conv1: k3, pad 1, no activation
      GemmDesc d = base_desc(mel_cl, MP, (long long)Tm * MP, Tm, MP, Tm, nb, m.conv1);
      d.n_taps = 3; d.tap_k = MP;
      for (int k = 0; k < 3; ++k) { d.tap_row[k] = k - 1; d.tap_col[k] = 0; }
      set_out(d, stem, D, (long long)2 * T * D);
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, at));
    }
    {   // conv2: k3, stride 2, pad 1 on the (row-pair) view of the stem: token j reads frames 2j-1, 2j, 2j+1
      GemmDesc d = base_desc(stem, 2 * D, (long long)2 * T * D, T, 2 * D, T, nb, m.conv2);
      d.n_taps = 3; d.tap_k = D;
      d.tap_row[0] = -1; d.tap_col[0] = D;
      d.tap_row[1] = 0; d.tap_col[1] = 0;
      d.tap_row[2] = 0; d.tap_col[2] = D;
      set_out(d, h, D, (long long)T * D);
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, 0));
    }
  }
  const bool ragged = (at == 1 || (m.x3() && get_gemm_variant() == 2)) && !c.force_simt && c.rag != nullptr && c.rag->nb == nb && c.rag->total > 0 && c.rag->t_max <= T;
  float* hp = nullptr;
  void* xnp = nullptr;
  if ((at == 1 || m.x3()) && (c.dry || ragged)) {       // packed residual stream and packed normalised output
    hp = (float*)c.ws.alloc((long long)nb * T * D * 4);
    xnp = c.ws.alloc((long long)nb * T * D * esz(at));
    SWC_TRY(c.ws.check());
  }
  if (ragged) {
    SWC_TRY(pack_rows(h, hp, *c.rag, T, D, c.s));
    SWC_TRY(transformer_stack(c, m.enc_layers, hp, nullptr, nb, T, c.rag));
    SWC_TRY(layernorm(hp, nullptr, nullptr, xnp, at, m.enc_ln_g, m.enc_ln_b, 1e-5f, 1, c.rag->total, c.rag->total, D, nullptr, c.s));
    SWC_TRY(unpack_rows(xnp, enc_cl, at, *c.rag, T4, D, c.s));      // rows >= len (and T..T4) are zero, as the masked LayerNorm writes them
  } else {
    SWC_TRY(transformer_stack(c, m.enc_layers, h, enc_lens, nb, T));
    if (!c.dry) SWC_TRY(layernorm(h, nullptr, nullptr, enc_cl, at, m.enc_ln_g, m.enc_ln_b, 1e-5f, nb, T, T4, D, enc_lens, c.s));
    if (!c.dry && c.hidden_out)
      SWC_TRY(cl_to_cf(enc_cl, at, c.hidden_out + (long long)m.enc_layers.size() * nb * D * T, nb, D, T, (long long)T4 * D, D, c.s));
  }
  c.ws.release(mark);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// residual units shared by the down- and up-sampler (reference modules.py:37-49); x fp32 (nb,Tc,512)
// ------------------------------------------------------------------------------------------------
static int residual_units(Ctx& c, const ResUnitW* R, float* x, bf16* x_bf_last, int nb, int Tc) {
  const Model& m = *c.m;
  const int H = m.hidden, at = m.act_type();
  const long long rows = (long long)nb * Tc;
  const size_t mark = c.ws.mark();
  void* a = c.ws.alloc(rows * H * esz(at));
  void* cbuf = c.ws.alloc(rows * H * esz(at));
  SWC_TRY(c.ws.check());
  for (int i = 0; i < 3 && !c.dry; ++i) {
    const ResUnitW& r = R[i];
    SWC_TRY(aa_snake(x, 0, a, at, r.fu0, r.fd0, r.a0, r.b0, nb, Tc, H, c.s));
    {
      GemmDesc d = base_desc(a, H, (long long)Tc * H, Tc, H, Tc, nb, r.conv7);
      d.n_taps = 7; d.tap_k = H;
      for (int k = 0; k < 7; ++k) { d.tap_row[k] = (k - 3) * r.dilation; d.tap_col[k] = 0; }
      set_out(d, cbuf, H, (long long)Tc * H);
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, at));
    }
    SWC_TRY(aa_snake(cbuf, at, a, at, r.fu2, r.fd2, r.a2, r.b2, nb, Tc, H, c.s));
    {
      GemmDesc d = base_desc(a, H, 0, (int)rows, H, (int)rows, 1, r.conv1);
      set_out(d, x, H, 0);
      d.epi.residual = x; d.epi.res_row_stride = H;
      if (i == 2) d.epi.out2 = x_bf_last;
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, 0));
    }
  }
  c.ws.release(mark);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// down-sampler + FSQ (reference modules.py:519-550, quantizer.py:273-290)
// enc_cl (nb,T4,768): T4 multiple of 4, rows beyond the real T are zero (the reference zero-pads).
// ------------------------------------------------------------------------------------------------
int downsample_fsq(Ctx& c, const void* enc_cl, const long long* code_lens, int nb, int T4, int* codes, float* zq_cf,
                   float* latent_cf, float* zq_cl) {
  const Model& m = *c.m;
  const int D = m.d_model, H = m.hidden, at = m.act_type();
  const int Tc = T4 / 4;
  const size_t mark = c.ws.mark();
  float* x = (float*)c.ws.alloc((long long)nb * Tc * H * 4);
  bf16* xb = at == 1 ? (bf16*)c.ws.alloc((long long)nb * Tc * H * 2) : nullptr;
  SWC_TRY(c.ws.check());
  if (!c.dry) {
    GemmDesc d = base_desc(enc_cl, 4 * D, (long long)T4 * D, Tc, 4 * D, Tc, nb, m.dn_in);
    set_out(d, x, H, (long long)Tc * H);
    SWC_TRY(run_gemm(c, d, EPI_STORE, at, 0));
  }
  SWC_TRY(residual_units(c, m.dn_res, x, xb, nb, Tc));
  if (!c.dry) {
    GemmDesc d = base_desc(at == 1 ? (const void*)xb : (const void*)x, H, (long long)Tc * H, Tc, H, Tc, nb, m.dn_latent);
    d.epi.lens = code_lens; d.epi.codes = codes; d.epi.zq_cf = zq_cf; d.epi.latent_cf = latent_cf; d.epi.zq_cl = zq_cl;
    d.epi.fsq = m.fsq;
    SWC_TRY(run_gemm(c, d, EPI_FSQ, at, 0));
  }
  c.ws.release(mark);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// up-sampler (reference modules.py:601-631): zq_cl fp32 (nb,Tc,32) -> h fp32 (nb,4Tc,768)
// ------------------------------------------------------------------------------------------------
int upsample_cl(Ctx& c, const float* zq_cl, int nb, int Tc, float* h) {
  const Model& m = *c.m;
  const int D = m.d_model, H = m.hidden, at = m.act_type();
  const size_t mark = c.ws.mark();
  float* x = (float*)c.ws.alloc((long long)nb * Tc * H * 4);
  bf16* xb = at == 1 ? (bf16*)c.ws.alloc((long long)nb * Tc * H * 2) : nullptr;
  SWC_TRY(c.ws.check());
  if (!c.dry) {   // from_latent: K = 32, always the fp32 SIMT kernel
    GemmDesc d = base_desc(zq_cl, m.latent, 0, nb * Tc, m.latent, nb * Tc, 1, m.up_from);
    set_out(d, x, H, 0);
    SWC_TRY(gemm_simt(d, EPI_STORE, 0, 0, c.s));
  }
  SWC_TRY(residual_units(c, m.up_res, x, xb, nb, Tc));
  if (!c.dry) {   // to_stacked with rows ordered (s, d): output row t' of width 3072 == 4 token rows of 768
    GemmDesc d = base_desc(at == 1 ? (const void*)xb : (const void*)x, H, 0, nb * Tc, H, nb * Tc, 1, m.up_stacked);
    set_out(d, h, 4 * D, 0);
    SWC_TRY(run_gemm(c, d, EPI_STORE, at, 0));
  }
  c.ws.release(mark);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// decoder (reference modules.py:437-474): h fp32 (nb,T,768) (destroyed) -> mel_cl (nb,2T,128) act type
// ------------------------------------------------------------------------------------------------
int decoder_cl(Ctx& c, float* h, const long long* lens, int nb, int T, void* mel_cl) {
  const Model& m = *c.m;
  const int D = m.d_model, MP = m.mel_pitch, at = m.act_type();
  const size_t mark = c.ws.mark();
  const bool ragged = (at == 1 || (m.x3() && get_gemm_variant() == 2)) && !c.force_simt && c.rag != nullptr && c.rag->nb == nb && c.rag->total > 0 && c.rag->t_max <= T;
  float* hp = nullptr;
  void* xnp = nullptr;
  if ((at == 1 || m.x3()) && (c.dry || ragged)) {
    hp = (float*)c.ws.alloc((long long)nb * T * D * 4);
    xnp = c.ws.alloc((long long)nb * T * D * esz(at));
    SWC_TRY(c.ws.check());
  }
  if (ragged) {
    SWC_TRY(pack_rows(h, hp, *c.rag, T, D, c.s));
    SWC_TRY(transformer_stack(c, m.dec_layers, hp, nullptr, nb, T, c.rag));
  } else {
    SWC_TRY(transformer_stack(c, m.dec_layers, h, lens, nb, T));
  }
  void* y = c.ws.alloc((long long)nb * T * D * esz(at));
  void* z = c.ws.alloc((long long)nb * 2 * T * D * esz(at));
  SWC_TRY(c.ws.check());
  if (!c.dry) {
    if (ragged) {
      SWC_TRY(layernorm(hp, nullptr, nullptr, xnp, at, m.dec_ln_g, m.dec_ln_b, 1e-5f, 1, c.rag->total, c.rag->total, D, nullptr, c.s));
      SWC_TRY(unpack_rows(xnp, y, at, *c.rag, T, D, c.s));
    } else {
      SWC_TRY(layernorm(h, nullptr, nullptr, y, at, m.dec_ln_g, m.dec_ln_b, 1e-5f, nb, T, T, D, lens, c.s));
    }
    {   // deconv1 even output rows 2t: taps h[t-1] (k=2), h[t] (k=0)
      GemmDesc d = base_desc(y, D, (long long)T * D, T, D, T, nb, m.deconv1_even);
      d.n_taps = 2; d.tap_k = D;
      d.tap_row[0] = -1; d.tap_row[1] = 0; d.tap_col[0] = d.tap_col[1] = 0;
      set_out(d, z, D, (long long)2 * T * D);
      d.epi.out_row_mul = 2; d.epi.out_row_off = 0;
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, at));
    }
    {   // odd output rows 2t+1: tap h[t] (k=1)
      GemmDesc d = base_desc(y, D, (long long)T * D, T, D, T, nb, m.deconv1_odd);
      set_out(d, z, D, (long long)2 * T * D);
      d.epi.out_row_mul = 2; d.epi.out_row_off = 1;
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, at));
    }
    {   // deconv2 (k3, stride 1), first 2T outputs: o[u] = sum_k W2_k^T z[u-k]
      GemmDesc d = base_desc(z, D, (long long)2 * T * D, 2 * T, D, 2 * T, nb, m.deconv2);
      d.n_taps = 3; d.tap_k = D;
      for (int k = 0; k < 3; ++k) { d.tap_row[k] = -k; d.tap_col[k] = 0; }
      set_out(d, mel_cl, MP, (long long)2 * T * MP);
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, at));
    }
  }
  c.ws.release(mark);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Vocos backbone + ISTFT head (reference modules.py:1492-1504, 1229-1248, 1064-1082, 831-886)
// mel_cl (nb,Tv,128) act type -> wav fp32 (nb,160 Tv).  No length masking anywhere.
// vt (optional): PACKED items.  mel_cl then holds vt->total rows, item b = rows [off[b], off[b] + len[b]) with at least three
// zero rows between consecutive items (pack_rows_gap), every GEMM / LayerNorm runs over all rows as one batch, the depthwise
// convolutions and the overlap-add take the item geometry from the table, and item b's 160 len[b] samples go to
// wav + b * wav_stride.  Sizing (dry) runs pass extra_rows for the gap rows of the packed form.
// ------------------------------------------------------------------------------------------------
int vocos_cl(Ctx& c, const void* mel_cl, int nb, int Tv, float* wav, const RaggedTable* vt, long long wav_stride, int extra_rows) {
  const Model& m = *c.m;
  if (wav_stride <= 0) wav_stride = 160ll * Tv;
  const int V = m.voc_dim, I = m.voc_inter, MP = m.mel_pitch, at = m.act_type();
  const long long rows = vt ? (long long)vt->total : (long long)nb * Tv + extra_rows;
  const int NP = m.voc_head.N;
  const size_t mark = c.ws.mark();
  float* x = (float*)c.ws.alloc(rows * V * 4);       // fp32 residual stream; pwconv2 accumulates into it (x += gamma * (...))
  void* y = c.ws.alloc(rows * V * esz(at));
  // the 4096-wide hidden, the spectrum and the frames are never live together: share one region
  const size_t big = std::max<size_t>((size_t)rows * I * esz(at), (size_t)rows * (NP + m.n_fft) * 4);
  char* region = (char*)c.ws.alloc(big);
  SWC_TRY(c.ws.check());
  if (!c.dry) {
    void* g = region;
    float* e = (float*)region;     // embed output (fp32) lives in the region until the first LayerNorm
    {
      // packed: one batch of `rows` rows; the taps of an item's first / last rows land in the zero rows between the items
      GemmDesc d = vt ? base_desc(mel_cl, MP, 0, (int)rows, MP, (int)rows, 1, m.voc_embed)
                      : base_desc(mel_cl, MP, (long long)Tv * MP, Tv, MP, Tv, nb, m.voc_embed);
      d.n_taps = 7; d.tap_k = MP;
      for (int k = 0; k < 7; ++k) { d.tap_row[k] = k - 3; d.tap_col[k] = 0; }
      set_out(d, e, V, vt ? 0 : (long long)Tv * V);
      SWC_TRY(run_gemm(c, d, EPI_STORE, at, 0));
    }
    SWC_TRY(layernorm(e, nullptr, nullptr, x, 0, m.voc_norm_g, m.voc_norm_b, 1e-6f, 1, (int)rows, (int)rows, V, nullptr, c.s));
    for (const VocosBlockW& B : m.voc_blocks) {
      const bool y_planes = x3_rowwise_planes(c, B.pw1);
      if (vt) SWC_TRY(dwconv7_ln_ragged(x, B.dw_w, B.dw_b, B.ln_g, B.ln_b, 1e-6f, y, y_planes ? 2 : at, *vt, V, c.s));
      else SWC_TRY(dwconv7_ln(x, nullptr, nullptr, B.dw_w, B.dw_b, B.ln_g, B.ln_b, 1e-6f, y, y_planes ? 2 : at, nb, Tv, V, c.s));
      const bool planes = x3_planes(c, B.pw1, B.pw2);      // bf16x3 mode: the 4096-wide hidden goes to pwconv2 as bf16 planes
      {
        GemmDesc d = base_desc(y, y_planes ? 2 * V : V, 0, (int)rows, V, (int)rows, 1, B.pw1);
        d.a_planes = y_planes;
        set_out(d, g, planes ? 2 * I : I, 0);
        d.epi.act = at == 1 ? 2 : 1;
        d.epi.out_planes = planes;
        SWC_TRY(run_gemm(c, d, EPI_STORE, at, planes ? 1 : at));
      }
      {
        GemmDesc d = base_desc(g, planes ? 2 * I : I, 0, (int)rows, I, (int)rows, 1, B.pw2);
        d.a_planes = planes;
        set_out(d, x, V, 0);
        d.epi.gamma = B.gamma;
        d.epi.residual = x; d.epi.res_row_stride = V;
        SWC_TRY(run_gemm(c, d, EPI_STORE, at, 0));
      }
    }
    SWC_TRY(layernorm(x, nullptr, nullptr, y, at, m.voc_final_g, m.voc_final_b, 1e-6f, 1, (int)rows, (int)rows, V, nullptr, c.s));
    float* S = (float*)region;
    float* frames = S + rows * NP;
    const bool tc_idft = (at == 1 || m.x3()) && !c.force_simt && m.w_idft3 != nullptr;
    {   // head GEMM with exp/clip/sincos epilogue -> interleaved complex spectrum (fp32, or split bf16 planes s1|s2)
      GemmDesc d = base_desc(y, V, 0, (int)rows, V, (int)rows, 1, m.voc_head);
      set_out(d, S, NP, 0);
      if (tc_idft) d.epi.out2 = (bf16*)S;
      SWC_TRY(run_gemm(c, d, EPI_HEAD, at, 0));
    }
    if (tc_idft) {
      // windowed inverse real DFT on the tensor cores with fp32-class accuracy: three bf16 products
      // s1 w1 + s1 w2 + s2 w1 accumulated in fp32 (taps select the s plane, the packed operand holds w1|w2|w1)
      LinearW wi; wi.w = m.w_idft3; wi.bias = nullptr; wi.N = m.n_fft; wi.w_rows = m.n_fft; wi.K = NP;
      GemmDesc d = base_desc(S, 2 * NP, 0, (int)rows, 2 * NP, (int)rows, 1, wi);
      d.n_taps = 3; d.tap_k = NP;
      d.tap_row[0] = d.tap_row[1] = d.tap_row[2] = 0;
      d.tap_col[0] = 0; d.tap_col[1] = 0; d.tap_col[2] = NP;
      set_out(d, frames, m.n_fft, 0);
      SWC_TRY(run_gemm(c, d, EPI_STORE, 1, 0));
    } else {   // fp32 parity mode: the same contraction as an fp32 FFMA GEMM
      LinearW wi; wi.w = m.w_idft; wi.bias = nullptr; wi.N = m.n_fft; wi.w_rows = m.n_fft; wi.K = NP;
      GemmDesc d = base_desc(S, NP, 0, (int)rows, NP, (int)rows, 1, wi);
      set_out(d, frames, m.n_fft, 0);
      SWC_TRY(gemm_simt(d, EPI_STORE, 0, 0, c.s));
    }
    if (vt) SWC_TRY(istft_ola_ragged(frames, m.win_sq, *vt, wav, wav_stride, c.s));
    else SWC_TRY(istft_ola(frames, m.win_sq, nb, Tv, wav, wav_stride, c.s));
  }
  c.ws.release(mark);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// log-mel (reference feature_extractor.py:86-112, 136-245): wav -> mel_cf fp32 / mel_cl act type
// ------------------------------------------------------------------------------------------------
int mel_frontend(Ctx& c, const float* wav, long long wav_stride, int wav_cols, const long long* lens, int nb,
                 float* mel_cf, void* mel_cl, long long* mel_lens) {
  const Model& m = *c.m;
  const size_t mark = c.ws.mark();
  float* padded = (float*)c.ws.alloc((long long)nb * 480400 * 4);
  float* power = (float*)c.ws.alloc((long long)nb * 3000 * 208 * 4);
  float* logmel = (float*)c.ws.alloc((long long)nb * 3000 * 80 * 4);
  float* item_max = (float*)c.ws.alloc((long long)nb * 4);
  const bool tc_dft = (m.act_type() == 1 || m.x3()) && !c.force_simt && m.w_dft6 != nullptr;
  bf16* frames = (m.act_type() == 1 || m.x3()) ? (bf16*)c.ws.alloc((long long)nb * 3000 * 3 * 448 * 2) : nullptr;
  SWC_TRY(c.ws.check());
  if (!c.dry) {
    SWC_TRY(mel_pad(wav, wav_stride, wav_cols, lens, nb, padded, mel_lens, item_max, c.s));
    if (tc_dft) {
      // windowed DFT on the tensor cores at fp32 accuracy: sample and operand are split into three bf16 terms and the six
      // significant products a1w1 + a1w2 + a2w1 + a1w3 + a2w2 + a3w1 accumulate in fp32 (one 6-tap GEMM, K = 6 x 448)
      SWC_TRY(mel_frames_split(padded, nb, frames, c.s));
      LinearW w; w.w = m.w_dft6; w.bias = nullptr; w.N = 416; w.w_rows = 416; w.K = 448;
      GemmDesc d = base_desc(frames, 3 * 448, (long long)3000 * 3 * 448, 3000, 3 * 448, 3000, nb, w);
      d.n_taps = 6; d.tap_k = 448;
      const int plane[6] = {0, 0, 1, 0, 1, 2};
      for (int t = 0; t < 6; ++t) { d.tap_row[t] = 0; d.tap_col[t] = plane[t] * 448; }
      set_out(d, power, 208, (long long)3000 * 208);
      SWC_TRY(gemm_tc(d, EPI_POWER, 0, m.num_sms, c.s));
    } else {   // frames are rows of the padded signal at stride 160; Hann window folded into the DFT operand
      LinearW w; w.w = m.w_dft; w.bias = nullptr; w.N = 416; w.w_rows = 416; w.K = 400;
      GemmDesc d = base_desc(padded, 160, 480400, 3000, 400, 3000, nb, w);
      set_out(d, power, 208, (long long)3000 * 208);
      SWC_TRY(gemm_simt(d, EPI_POWER, 0, 0, c.s));
    }
    {
      LinearW w; w.w = m.w_melfb; w.bias = nullptr; w.N = 80; w.w_rows = 80; w.K = 208;
      GemmDesc d = base_desc(power, 208, (long long)3000 * 208, 3000, 208, 3000, nb, w);
      set_out(d, logmel, 80, (long long)3000 * 80);
      d.epi.item_max = item_max;
      SWC_TRY(gemm_simt(d, EPI_LOGMEL, 0, 0, c.s));
    }
    SWC_TRY(mel_finalize(logmel, item_max, nb, mel_cf, mel_cl, m.act_type(), m.mel_pitch, c.s));
  }
  c.ws.release(mark);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// fused chains
// ------------------------------------------------------------------------------------------------
int tokenize_chain(Ctx& c, const float* wav, long long wav_stride, int wav_cols, const long long* lens, int nb,
                   int* codes, float* zq_cf, long long* codes_lens) {
  const Model& m = *c.m;
  const int at = m.act_type(), Tm = 3000, T = 1500, T4 = 1500;
  const size_t mark = c.ws.mark();
  void* mel_cl = c.ws.alloc((long long)nb * Tm * m.mel_pitch * esz(at));
  void* enc_cl = c.ws.alloc((long long)nb * T4 * m.d_model * esz(at));
  long long* mel_lens = (long long*)c.ws.alloc(nb * 8);
  long long* enc_lens = (long long*)c.ws.alloc(nb * 8);
  long long* code_lens = (long long*)c.ws.alloc(nb * 8);
  SWC_TRY(c.ws.check());
  SWC_TRY(mel_frontend(c, wav, wav_stride, wav_cols, lens, nb, nullptr, mel_cl, mel_lens));
  SWC_TRY(lens_affine(c, mel_lens, enc_lens, nb, 1, 0, 2));
  SWC_TRY(lens_affine(c, enc_lens, code_lens, nb, 1, 3, 4));
  SWC_TRY(lens_affine(c, code_lens, codes_lens, nb, 1, 0, 1));
  SWC_TRY(encoder_cl(c, mel_cl, enc_lens, nb, Tm, enc_cl));
  SWC_TRY(downsample_fsq(c, enc_cl, code_lens, nb, T4, codes, zq_cf, nullptr, nullptr));
  (void)T;
  c.ws.release(mark);
  return 0;
}

int detokenize_chain(Ctx& c, const float* zq_cl, const long long* code_lens, int nb, int Tc, float* wav,
                     long long* out_lens) {
  const Model& m = *c.m;
  const int at = m.act_type(), T = 4 * Tc, Tv = 8 * Tc;
  const size_t mark = c.ws.mark();
  float* h = (float*)c.ws.alloc((long long)nb * T * m.d_model * 4);
  void* mel_cl = c.ws.alloc((long long)nb * Tv * m.mel_pitch * esz(at));
  long long* tok_lens = (long long*)c.ws.alloc(nb * 8);
  SWC_TRY(c.ws.check());
  SWC_TRY(lens_affine(c, code_lens, tok_lens, nb, 4, 0, 1));
  SWC_TRY(lens_affine(c, code_lens, out_lens, nb, 1280, 0, 1));
  SWC_TRY(upsample_cl(c, zq_cl, nb, Tc, h));
  SWC_TRY(decoder_cl(c, h, tok_lens, nb, T, mel_cl));
  // Vocos has no length masking (reference modules.py:1492-1504) but its receptive field is finite: an output sample
  // of frame t reads frames <= t + 1 of the last block, each of the 24 depthwise convolutions and the embedding add 3:
  // everything a valid sample depends on lies below frame 8 len + 77.  When the caller passed host lengths, item b
  // computes only need(b) = min(Tv, 8 len + 80) frames: the items' rows are packed back to back (eight zero rows between
  // them for the embedding convolution's taps), every GEMM of the backbone runs once over all packed rows, and the
  // depthwise convolutions / the overlap-add see each item's cut as a sequence end, which can reach valid samples only
  // from frame 8 len + 5 on.  Valid samples are bit-identical to the full-length computation.
  const bool packed = c.rag != nullptr && c.rag->nb == nb && (at == 1 || m.x3());
  constexpr int kGap = 8;
  void* mel_pk = nullptr;
  if (c.dry || packed) {
    mel_pk = c.ws.alloc((long long)nb * (Tv + kGap) * m.mel_pitch * esz(at));
    SWC_TRY(c.ws.check());
  }
  if (c.dry) {
    SWC_TRY(vocos_cl(c, mel_cl, nb, Tv, wav, nullptr, 0, nb * kGap));
  } else if (packed) {
    // receptive field of a valid sample: 3 frames per depthwise convolution and for the embedding, + 1 for the overlap-add,
    // rounded up to 8 (24 ConvNeXt blocks: 77 -> 80); derived from the model so another depth cannot silently corrupt frames
    const int kHalo = (3 * ((int)m.voc_blocks.size() + 1) + 2 + 7) / 8 * 8;
    RaggedTable vt;
    vt.nb = nb; vt.total = 0; vt.t_max = 0;
    for (int b = 0; b < nb; ++b) {
      const int need = std::min(Tv, (2 * c.rag->len[b] + kHalo + 7) / 8 * 8);      // len = tokens = 4 x code frames
      vt.len[b] = need;
      vt.off[b] = vt.total;
      vt.total += need + (b + 1 < nb ? kGap : 0);
      vt.t_max = std::max(vt.t_max, need);
    }
    vt.off[nb] = vt.total;
    SWC_TRY(pack_rows_gap(mel_cl, mel_pk, vt, Tv, m.mel_pitch * (int)esz(at), c.s));
    SWC_TRY(vocos_cl(c, mel_pk, nb, Tv, wav, &vt, 160ll * Tv));
  } else {
    SWC_TRY(vocos_cl(c, mel_cl, nb, Tv, wav));
  }
  c.ws.release(mark);
  return 0;
}

int lens_affine_pub(Ctx& c, const long long* in, long long* out, int n, long long mul, long long add, long long div) {
  return lens_affine(c, in, out, n, mul, add, div);
}

}  // namespace swc
