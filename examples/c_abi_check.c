/* Plain-C consumer of include/swc.h: proves the boundary is a C ABI (no C++ or torch types in the signatures) and shows the
 * call sequence a host in any language binds: create -> set tensors -> pack/finalize -> workspace query -> stage calls.
 * Built and run by tests/test_packing_cpu.py::test_header_is_plain_c_and_links (gcc, no GPU needed): without weights the
 * model cannot be packed, so this only exercises the lifecycle and the error path; the compute entry points are referenced so
 * that a missing export fails at link time.
 *   gcc -std=c99 -Wall -Werror -Iinclude examples/c_abi_check.c -o c_abi_check simwhisper_codec_b200/libswc.so */
#include <stdio.h>
#include <string.h>

#include "swc.h"

int main(void) {
  swc_model* m = NULL;
  if (swc_version() <= 0) return 1;
  if (swc_model_create(&m, SWC_PRECISION_BF16X3) != 0 || m == NULL) return 2;
  /* a tensor under a reference state_dict key (audiocodec/model.py load_state_dict): accepted as data, checked at pack time */
  {
    const float ones[4] = {1.f, 1.f, 1.f, 1.f};
    const int64_t shape[1] = {4};
    if (swc_model_set_tensor(m, "vocos.backbone.norm.weight", ones, SWC_DTYPE_F32, shape, 1) != 0) return 3;
  }
  /* one tensor is not a model: packing must fail with a message, not crash */
  if (swc_model_pack(m) == 0) return 4;
  if (strlen(swc_last_error()) == 0) return 5;
  printf("pack without weights -> \"%s\"\n", swc_last_error());
  /* the compute entry points exist with C linkage (addresses taken, not called: no GPU here) */
  {
    const void* fns[] = {(const void*)swc_mel, (const void*)swc_encoder, (const void*)swc_downsample, (const void*)swc_quantize,
                         (const void*)swc_dequantize, (const void*)swc_upsample, (const void*)swc_decoder, (const void*)swc_vocos,
                         (const void*)swc_tokenize, (const void*)swc_detokenize, (const void*)swc_forward,
                         (const void*)swc_tokenize_ragged, (const void*)swc_detokenize_ragged, (const void*)swc_workspace_bytes};
    size_t i;
    for (i = 0; i < sizeof(fns) / sizeof(fns[0]); ++i)
      if (fns[i] == NULL) return 6;
  }
  swc_model_destroy(m);
  puts("ok");
  return 0;
}
